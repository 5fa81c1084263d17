cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
WK_GEMM_TAILSPLIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f -o gpurun_out/ncu_gemm_f32_r02m_n32768 python tools/gemm_time.py f32 32768 2>&1 | tail -4
ls -la gpurun_out/*.ncu-rep
