cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_int_tc.py -x -q 2>&1 | tail -3
python - <<'P'
import sys, ctypes as C, numpy as np
sys.path.insert(0, ".")
import wekua_b200 as wk
ctx = wk.Context.init([0]); pipe = wk.Pipeline.init(ctx.command_queues[0]); lib = wk.capi.lib()
def ev():
    e = C.c_void_p(); wk.capi.check(lib.wk_event_record(pipe.q, C.byref(e))); return e
for dtype, n in ((np.int16, 8192), (np.int32, 8192), (np.int64, 4096), (np.int64, 8192), (np.uint64, 8192)):
    a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
    wk.tensor.random.uniform(pipe, a, 42); wk.tensor.random.uniform(pipe, b, 43)
    for oa, ob in ((0, 0), (0, 1)):
        for _ in range(2): wk.blas.gemm(pipe, None, a, oa, b, ob, None, c)
        pipe.wait_and_cleanup()
        l0 = wk.capi.launch_count(); e0 = ev()
        for _ in range(3): wk.blas.gemm(pipe, None, a, oa, b, ob, None, c)
        e1 = ev(); lib.wk_event_wait(e1); ms = C.c_float(); lib.wk_event_elapsed_ms(e0, e1, C.byref(ms))
        print(f"{np.dtype(dtype).name} N={n} {'NT'[oa]}{'NT'[ob]}: {2*n**3*3/(ms.value*1e-3)/1e12:.1f} Top/s ({ms.value/3:.3f} ms, {(wk.capi.launch_count()-l0)//3} launches)", flush=True)
    for t in (a, b, c): t.release(pipe)
P
