cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int_tc.py -m gpu -x -q -k graph 2>&1 | tail -15
