#!/bin/bash
# end-of-round check: -m gpu suite, smoke, bench (ours + reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-fin}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
FD_NO_PDL=1 timeout 300 python tools/profile_step.py > gpurun_out/${T}_warm_kernel_times.txt 2>/dev/null; echo "warm rc=$?"
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log
