#!/bin/bash
# round-2 full GPU check: -m gpu suite, smoke, bench (ours + reference arm), warm kernel times, rows, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r2}
S=gpurun_out/${T}_summary.txt
rm -f $S
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> $S
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?" >> $S
timeout 300 python tools/profile_step.py > gpurun_out/${T}_warm_kernel_times.txt 2>&1; echo "warm rc=$?" >> $S
timeout 600 python tools/bench_rows.py > gpurun_out/${T}_rows.jsonl 2> gpurun_out/${T}_rows.err; echo "rows rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?" >> $S
cat $S; tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench.json
