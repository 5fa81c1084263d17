cd $GRAFT_REPO_ROOT
S="128 128 128 256 256 256 384 384 384 512 512 512 256 1024 256"
python tools/gemm_small_time.py f32 $S
python tools/gemm_small_time.py f64 $S
WK_GEMM_PATH=simt python tools/gemm_small_time.py f32 512 512 512 768 768 768
WK_GEMM_PATH=simt python tools/gemm_small_time.py f64 768 768 768
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_xor.py tests/test_gpu_complex.py -x -q 2>&1 | tail -3
