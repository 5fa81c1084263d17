set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./build/mma_peak_i8 > gpurun_out/mma_peak_i8_r02.txt 2>&1; cat gpurun_out/mma_peak_i8_r02.txt
# ncu: HEAD f32 kernel at N=32768, tail split off (its in-kernel wait for sibling splits traps under ncu's replay)
WK_GEMM_TAILSPLIT=0 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -o gpurun_out/ncu_gemm_f32_r02_n32768 python tools/gemm_time.py f32 32768 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
cat > /tmp/i8time.py <<'P'
import sys, ctypes as C, numpy as np
sys.path.insert(0, ".")
import wekua_b200 as wk
ctx = wk.Context.init([0]); pipe = wk.Pipeline.init(ctx.command_queues[0])
n = 8192
a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), np.int8) for _ in range(3))
wk.tensor.random.uniform(pipe, a, 42); wk.tensor.random.uniform(pipe, b, 43)
for _ in range(4):
    wk.blas.gemm(pipe, None, a, 0, b, 1, None, c)
pipe.wait_and_cleanup()
P
ncu --set full --clock-control none --import-source on -k regex:gemm_u8_kernel -s 1 -c 1 -o gpurun_out/ncu_gemm_i8_r02_n8192 python /tmp/i8time.py > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_r02b.log
cat gpurun_out/pytest_r02b.log
