cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_xor.py tests/test_gpu_multi_tensor.py -x -q 2>&1 | tail -15
python tools/linear_time.py 256 4096 4096 2>&1 | tail -4
python tools/linear_time.py 2048 2048 2048 2>&1 | tail -4
python tools/linear_time.py 8192 8192 8192 2>&1 | tail -4
