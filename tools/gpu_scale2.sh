#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
python bench.py --steps 20 --warmup 5 --no-extra-legs --no-cpu-baseline > gpurun_out/s_1gpu.json 2> gpurun_out/s_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extra-legs > gpurun_out/s_${N}gpu.json 2> gpurun_out/s_${N}gpu.err
python - <<PY
import json
a=json.load(open("gpurun_out/s_1gpu.json")); b=json.load(open("gpurun_out/s_${N}gpu.json"))
print("1 gpu", a["value"], a["ms_per_step"], "e2e_u8", a["e2e_u8"]["value"])
print("${N} gpu", b["value"], b["ms_per_step"], "eff", b["value"]/a["value"]/${N}, b["collective"], "status", b["comm_status"], "e2e_u8", b["e2e_u8"]["value"], "launches", b["launches_per_step"])
PY
tail -3 gpurun_out/s_${N}gpu.err
