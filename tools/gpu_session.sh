cd $GRAFT_REPO_ROOT
WK_GEMM_TRACE=1 python tools/gemm_trace.py 256 256 256 1024 1024 1024 2>&1 | grep "wk trace" | grep -v "cta  *[1-9][0-9]* *:" 
python tools/gemm_small_time.py f32 256 256 256 512 512 512 1024 1024 1024 1536 1536 1536 2048 2048 2048 256 4096 4096 64 512 4096 128 1024 8192
python tools/gemm_time.py f32 4096 16384
timeout 800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_splitk.py tests/test_gpu_linear.py tests/test_gpu_fullsize.py tests/test_gpu_xor.py -x -q 2>&1 | tail -4
