cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 0 1 2; do echo "== WK_UNIFORM_CONV=$c"; WK_UNIFORM_CONV=$c WK_SWEEP_ONLY=uniform python tools/stream_sweep.py gpurun_out/sweep_uni_c$c 27 2>&1 | tail -2; done
for c in 1 2; do WK_UNIFORM_CONV=$c timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_kernels.py -m gpu -x -q -k "uniform or random" 2>&1 | tail -2; done
