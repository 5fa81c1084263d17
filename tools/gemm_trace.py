"""WK_GEMM_TRACE=1 python tools/gemm_trace.py M N K [M N K ...]: two f32 gemm calls per shape; the library prints the
per-CTA pipeline milestones (entry, setup done, first TMA, first raw tile, first MMA, last MMA, accumulator full, epilogue
done, exit) of each launch to stderr -- where a launch-bound problem spends its time."""
import sys

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
dims = [int(v) for v in sys.argv[1:]]
for i in range(0, len(dims), 3):
    M, N, K = dims[i:i + 3]
    a, b, c = (wk.Tensor.alloc(ctx, pipe, s, np.float32) for s in ((M, K), (K, N), (M, N)))
    wk.tensor.random.uniform(pipe, a, 42, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    for _ in range(2):
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    pipe.wait_and_cleanup()
    for t in (a, b, c):
        t.release(pipe)
