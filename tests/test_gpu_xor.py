"""-m gpu: BASELINE config 3 -- examples/xor_neural_network.zig end to end through the CUDA backend
(Linear + Sigmoid + MSE + GD, 300 steps, batch 4), step-by-step against the op-by-op oracle run."""
import math

import numpy as np
import pytest

from tests import gpu_helpers as gh
from tests.xor_reference import INPUTS, TARGETS, OracleXor

pytestmark = pytest.mark.gpu


def _build(wk, ctx, pipe, dtype, seeds, fused):
    nn = wk.nn
    inputs = wk.Tensor.alloc(ctx, pipe, (4, 2), dtype)
    expected = wk.Tensor.alloc(ctx, pipe, (4, 1), dtype)
    wk.tensor.memory.read_from_buffer(pipe, inputs, INPUTS.astype(dtype))
    wk.tensor.memory.read_from_buffer(pipe, expected, TARGETS.astype(dtype))
    seq = nn.Sequential.init()
    act = nn.Sigmoid.init()
    seq.append(nn.Linear.init(ctx, pipe, 2, 10, act, dtype=dtype, seed=seeds[0], fused=fused))
    seq.append(nn.Linear.init(ctx, pipe, 10, 1, act, dtype=dtype, seed=seeds[1], fused=fused))
    layers = seq.layer()
    cache = nn.Cache.init(ctx, pipe, 4, [layers])
    opt = nn.GD.init(None, lr=1)
    return inputs, expected, seq, layers, cache, opt


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fused", [False, True])
def test_xor_training_matches_oracle(oracle, dtype, fused):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    seeds = (42, 43)
    inputs, expected, seq, layers, cache, opt = _build(wk, ctx, pipe, dtype, seeds, fused)
    ref = OracleXor(oracle, gh.oracle_dev(oracle), dtype, seeds)
    # same PRNG on both sides: initial weights must agree bit for bit (uniform.cl restated on both)
    for lin, W in zip(seq.layers, ref.W):
        np.testing.assert_array_equal(gh.to_np(lin.weights[0]), W.to_host())
    layer_cache = cache.get_layer_cache(0)
    eps = np.finfo(dtype).eps
    for it in range(300):
        output = layers.forward(pipe, inputs, layer_cache)
        wk.nn.mse(pipe, output, expected, cache, calculate_derivative=True)
        layers.backward(pipe, layer_cache, inputs, None)
        opt.step(pipe, cache)
        ref.step()
        if it in (0, 1, 9, 99, 299):
            # exp() differs by ulps between CUDA and glibc; the error compounds slowly over GD steps
            rtol = eps * 64 * (it + 1)
            for lin, W, b in zip(seq.layers, ref.W, ref.b):
                np.testing.assert_allclose(gh.to_np(lin.weights[0]), W.to_host(), rtol=rtol, atol=rtol)
                np.testing.assert_allclose(gh.to_np(lin.bias[0]), b.to_host(), rtol=rtol, atol=rtol)
    out = gh.to_np(layers.forward(pipe, inputs, layer_cache)).reshape(-1)
    np.testing.assert_allclose(out, ref.forward().to_host().reshape(-1), rtol=1e-3, atol=1e-4)
    assert np.all(np.abs(out - TARGETS.reshape(-1)) < 0.2), out  # it learned XOR
    err = wk.nn.mse(pipe, layers.forward(pipe, inputs, layer_cache), expected, cache, calculate_derivative=False, want_error=True)
    # SURVEY Q2, reproduced on purpose: sigmoid ran over the padded buffer, so the padded column of the [4,1]
    # error tensor holds (0 - 0.5)^2 and math.sum (sum.cl:33-35) adds it.  The value must equal the reference's.
    oracle.mse(ref.forward(), ref.t, ref.err, None)
    ref_err = float(oracle.mean(ref.err))
    assert math.isfinite(float(err)) and abs(float(err) - ref_err) <= 1e-4 * max(1.0, abs(ref_err))
    # closed form of that value: for a [4,1] tensor math.sum adds the first 4 elements of the pitch-2 BUFFER
    # (basic.zig:150-152,193-202), i.e. err[0], pad, err[1], pad with pad = (0 - sigmoid(0))^2 = 0.25, and mean divides
    # by the 4 logical elements
    e = (out.astype(np.float64) - TARGETS.reshape(-1)) ** 2
    assert abs(float(err) - (e[0] + e[1] + 0.5) / 4) < 1e-4
    opt.deinit(pipe)
    cache.deinit(pipe)
    seq.deinit(pipe)


def test_cpp_host_mirror_xor_example_matches_python_path(oracle):
    """examples/xor_neural_network.cpp (the reference's example against include/wekua.hpp) issues the same C-ABI call
    sequence as the Python mirror: with the same weight seeds its final outputs are the same numbers, and it learns XOR."""
    import os
    import re
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "bin", "xor_neural_network")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "examples")], check=True, capture_output=True)
    r = subprocess.run([exe, "42"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    m = re.search(r"output:\s+\[([^\]]+)\]", r.stdout)
    got = np.array([float(v) for v in m.group(1).split(",")])
    assert np.all(np.abs(got - TARGETS.reshape(-1)) < 0.2), r.stdout
    assert int(re.search(r"kernel launches: (\d+)", r.stdout).group(1)) > 300 * 10

    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    inputs, expected, seq, layers, cache, opt = _build(wk, ctx, pipe, np.float32, (42, 43), False)
    layer_cache = cache.get_layer_cache(0)
    for _ in range(300):
        output = layers.forward(pipe, inputs, layer_cache)
        wk.nn.mse(pipe, output, expected, cache, calculate_derivative=True)
        layers.backward(pipe, layer_cache, inputs, None)
        opt.step(pipe, cache)
    out = gh.to_np(layers.forward(pipe, inputs, layer_cache)).reshape(-1)
    np.testing.assert_allclose(got, out, rtol=0, atol=2e-6)  # printed with 6 decimals
    opt.deinit(pipe)
    cache.deinit(pipe)
    seq.deinit(pipe)


def test_training_step_as_cuda_graph_matches_eager(oracle):
    """one XOR training step (forward + mse + backward + GD, ~22 launches) captured once and replayed as a CUDA graph
    leaves the same weights, bit for bit, as launching every op eagerly"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    steps = 60

    def make():
        inputs, expected, seq, layers, cache, opt = _build(wk, ctx, pipe, np.float32, (42, 43), False)
        layer_cache = cache.get_layer_cache(0)

        def step():
            output = layers.forward(pipe, inputs, layer_cache)
            wk.nn.mse(pipe, output, expected, cache, calculate_derivative=True)
            layers.backward(pipe, layer_cache, inputs, None)
            opt.step(pipe, cache)
        return step, seq, cache, opt

    step_a, seq_a, cache_a, opt_a = make()
    for _ in range(steps):
        step_a()
    want = [gh.to_np(l.weights[0]) for l in seq_a.layers] + [gh.to_np(l.bias[0]) for l in seq_a.layers]

    step_b, seq_b, cache_b, opt_b = make()
    step_b()  # eager warm-up (first-call attribute setup happens outside the capture)
    pipe.wait_and_cleanup()
    pipe.begin_capture()
    step_b()
    graph = pipe.end_capture()
    assert graph.num_kernels >= 15
    launches0 = wk.capi.launch_count()
    for _ in range(steps - 1):
        graph.launch(pipe)
    pipe.wait_and_cleanup()
    assert wk.capi.launch_count() - launches0 == (steps - 1) * graph.num_kernels
    got = [gh.to_np(l.weights[0]) for l in seq_b.layers] + [gh.to_np(l.bias[0]) for l in seq_b.layers]
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    graph.release()
    for opt, cache, seq in ((opt_a, cache_a, seq_a), (opt_b, cache_b, seq_b)):
        opt.deinit(pipe)
        cache.deinit(pipe)
        seq.deinit(pipe)
