cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_layers.py -x -q -m gpu 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2c_ncu_bench.log 2>&1; echo "ncu rc=$?"; grep "^==ERROR" gpurun_out/r2c_launches_bench.csv | head -3; grep -vc "^==" gpurun_out/r2c_launches_bench.csv
