#!/bin/bash
# GPU run A: new kernels first (isolated processes), then parity, then old suite, then bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -x -q -m gpu -k "pw_conv" > gpurun_out/a_pw.log 2>&1; echo "pw rc=$?" >> gpurun_out/a_summary.txt
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -k "dwconv or resize" > gpurun_out/a_dw.log 2>&1; echo "dw/resize rc=$?" >> gpurun_out/a_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -s -k "not pw_conv and not dwconv and not resize_bilinear" > gpurun_out/a_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/a_summary.txt
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity_r2.py > gpurun_out/a_old.log 2>&1; echo "old rc=$?" >> gpurun_out/a_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_summary.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?" >> gpurun_out/a_summary.txt
cat gpurun_out/a_summary.txt
tail -5 gpurun_out/a_pw.log gpurun_out/a_dw.log gpurun_out/a_parity.log gpurun_out/a_old.log
