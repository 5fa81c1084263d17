cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( echo "# compute-sanitizer --tool memcheck over the GPU parity tests (all but the full-size ones), HEAD of round 2";
  timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_int_tc.py tests/test_gpu_splitk.py tests/test_gpu_unaligned.py tests/test_gpu_simt_mid.py tests/test_gpu_complex.py tests/test_gpu_multi_tensor.py tests/test_gpu_xor.py tests/test_gpu_stream_kernels.py tests/test_gpu_linear.py -m gpu -q -k "not full_size" 2>&1 | tail -25 ) | tee gpurun_out/sanitizer_memcheck_r02.txt
