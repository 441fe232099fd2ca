#!/bin/bash
# Runs the tcgen05 probes one per process with a timeout each; output -> gpurun_out/probe.log
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version --format=csv >> $LOG 2>&1
for args in "conv 0 2 15 15" "conv 0 4 60 60" "conv 0 4 30 30" \
            "wgrad 0 2 15 15" "wgrad 0 4 60 60" "wgrad 0 4 30 30" "conv 0 64 60 60" "conv 0 64 30 30" "conv 0 64 15 15" \
            "wgrad 0 64 60 60" "wgrad 0 64 30 30" "wgrad 0 64 15 15"; do
  echo "=== $args" >> $LOG
  timeout 120 python tools/gpu_probe.py $args >> $LOG 2>&1
  echo "exit=$?" >> $LOG
done
tail -n 80 $LOG
