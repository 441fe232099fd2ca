#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/c_summary.txt
timeout 600 python -m pytest tests/test_gpu_ssd_model.py -q -m gpu -s > gpurun_out/c_ssd.log 2>&1; echo "ssd rc=$?" >> gpurun_out/c_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -s > gpurun_out/c_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/c_summary.txt
timeout 300 python tools/profile_layers.py mbv3 256 > gpurun_out/c_prof_mbv3.txt 2>&1; echo "prof mbv3 rc=$?" >> gpurun_out/c_summary.txt
timeout 300 python tools/profile_layers.py ssd 16 > gpurun_out/c_prof_ssd.txt 2>&1; echo "prof ssd rc=$?" >> gpurun_out/c_summary.txt
cat gpurun_out/c_summary.txt
