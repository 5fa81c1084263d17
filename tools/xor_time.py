"""steps/s of the XOR training step (BASELINE config 3), eager launches vs one CUDA-graph replay per step"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402
from tests.test_gpu_xor import _build  # noqa: E402

ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
inputs, expected, seq, layers, cache, opt = _build(wk, ctx, pipe, np.float32, (42, 43), False)
lc = cache.get_layer_cache(0)


def step():
    out = layers.forward(pipe, inputs, lc)
    wk.nn.mse(pipe, out, expected, cache, calculate_derivative=True)
    layers.backward(pipe, lc, inputs, None)
    opt.step(pipe, cache)


for _ in range(20):
    step()
pipe.wait_and_cleanup()
n = 2000
t = time.perf_counter()
for _ in range(n):
    step()
pipe.wait_and_cleanup()
eager = n / (time.perf_counter() - t)
pipe.begin_capture()
step()
g = pipe.end_capture()
for _ in range(20):
    g.launch(pipe)
pipe.wait_and_cleanup()
t = time.perf_counter()
for _ in range(n):
    g.launch(pipe)
pipe.wait_and_cleanup()
graph = n / (time.perf_counter() - t)
print(f"xor step: eager {eager:.0f} steps/s ({1e6 / eager:.1f} us/step), graph {graph:.0f} steps/s ({1e6 / graph:.1f} us/step), "
      f"{g.num_kernels} kernels per step")
