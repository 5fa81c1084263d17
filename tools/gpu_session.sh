cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_splitk.py tests/test_gpu_fullsize.py tests/test_gpu_complex.py -x -q 2>&1 | tail -3
python tools/gemm_sweep.py gpurun_out/sweep_gemm_r02 4096 8192 16384 > gpurun_out/sweep_gemm_r02.log 2>&1
cat gpurun_out/sweep_gemm_r02.md
