cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int_tc.py tests/test_gpu_complex.py -x -q 2>&1 | tail -4
timeout 600 python tools/gemm_int_sweep.py gpurun_out/sweep_gemm_int_r02c > gpurun_out/sweep_int_c.log 2>&1
cat gpurun_out/sweep_gemm_int_r02c.md | head -20
