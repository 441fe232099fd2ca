#!/bin/bash
# round-2 ncu captures: --set full of the top kernels (one eager step each), raw metric pages as CSV
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r2n}
KS='regex:conv3x3_tc_kernel|wgrad3x3_tc_kernel|resblock_chain_kernel|conv3x3_wide_kernel|conv3x3_wide_chain_kernel|wgrad3x3_wide_kernel|stem_fwd_tc_kernel|stem_fwd_bf16_kernel|stem_wgrad_bf16|sepblock_fwd_kernel|pw_gemm_kernel|dwconv|mbv3_stem_tc'
timeout 300 python tools/ncu_r2.py all > gpurun_out/${T}_plain.log 2>&1; echo "plain rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KS" -c 70 -o gpurun_out/${T}_full -f python tools/ncu_r2.py all > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2> gpurun_out/${T}_raw.err; echo "raw rc=$?"
# SASS-level stall samples of the dominant kernels (source page), then drop the report: gpurun_out/ is capped at 64 MiB
for k in conv3x3_tc_kernel conv3x3_wide_kernel wgrad3x3_wide_kernel; do
  ncu -i gpurun_out/${T}_full.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/${T}_source_$k.csv 2>/dev/null
done
rm -f gpurun_out/${T}_full.ncu-rep
ls -la gpurun_out/${T}_*
tail -3 gpurun_out/${T}_ncu.log
