cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:mbv3_dw_kernel|pw_gemm_kernel|mbv3_se_kernel|mbv3_scale_kernel|mbv3_head_kernel|sepblock_fwd" -c 60 -o gpurun_out/r2m_full -f python tools/ncu_r2.py infer > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2m_full.ncu-rep --page raw --csv > gpurun_out/r2m_raw.csv 2>/dev/null
ncu -i gpurun_out/r2m_full.ncu-rep --page source --csv -k regex:mbv3_dw_kernel -c 1 > gpurun_out/r2m_source_dw.csv 2>/dev/null
rm -f gpurun_out/r2m_full.ncu-rep
ls -la gpurun_out/r2m_*
