cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_simt_mid.py tests/test_gpu_parity.py tests/test_gpu_linear.py tests/test_gpu_xor.py tests/test_gpu_splitk.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
S32="128 128 128 256 256 256 384 384 384 512 512 512 640 640 640 768 768 768 1024 1024 1024 256 1024 256 256 1024 1024 128 4096 512"
S64="128 128 128 256 256 256 384 384 384 512 512 512 768 768 768 1024 1024 1024 256 1024 256 256 1024 1024"
(echo "# round 2c: after the mid-tile (32 x 64, cp.async ring) SIMT kernel; automatic routing"; python tools/gemm_small_time.py f32 $S32; python tools/gemm_small_time.py f64 $S64) | tee gpurun_out/small_gemm_r02c.txt
python tools/linear_time.py 2>&1 | tail -8
