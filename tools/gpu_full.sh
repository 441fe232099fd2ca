#!/bin/bash
# full GPU check: whole -m gpu suite, smoke, bench (ours + reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-f}
rm -f gpurun_out/${T}_summary.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_summary.txt
cat gpurun_out/${T}_summary.txt; tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log
