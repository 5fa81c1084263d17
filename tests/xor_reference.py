"""The XOR example (examples/xor_neural_network.zig:51-138) run op-by-op on the ORACLE: the exact launch sequence of
Linear.forward / loss.mse / Linear.backward / GD.step (linear.zig:480-678, mse.zig:63-132, gd.zig:55-94) on the
reference-layout tensors.  Test infrastructure only."""
import math

import numpy as np

INPUTS = np.array([[1, 1], [0, 1], [1, 0], [0, 0]], dtype=np.float32)
TARGETS = np.array([[0], [1], [1], [0]], dtype=np.float32)
SIZES = [(2, 10), (10, 1)]


class OracleXor:
    def __init__(self, oracle, dev, dtype, seeds, batch=4):
        o = self.o = oracle
        self.dev, self.dtype = dev, dtype
        T = lambda shape: o.OTensor(dev, dtype, shape)  # noqa: E731
        self.x = T((batch, 2)).read_from(INPUTS)
        self.t = T((batch, 1)).read_from(TARGETS)
        self.W, self.b, self.out, self.sens, self.dact, self.gW, self.gb = [], [], [], [], [], [], []
        for (n_in, n_out), seed in zip(SIZES, seeds):
            lim = math.sqrt(6.0 / (n_in + n_out))
            self.W.append(T((n_out, n_in)).uniform(seed, -lim, lim))
            self.b.append(T((n_out,)))
            self.out.append(T((batch, n_out)))
            self.sens.append(T((batch, n_out)).fill(1))
            self.dact.append(T((batch, n_out)))
            self.gW.append(T((n_out, n_in)))
            self.gb.append(T((n_out,)))
        self.err = T((batch, 1))

    def forward(self):
        o, inp = self.o, self.x
        for W, b, out in zip(self.W, self.b, self.out):
            o.gemm(None, inp, 0, W, 1, None, out, packed=True)
            o.bias(out, b)
            o.unary(out, "sigmoid")
            inp = out
        return inp

    def step(self):
        o = self.o
        out = self.forward()
        o.mse(out, self.t, self.err, self.sens[-1])
        for i in (1, 0):
            o.sigmoid_dev(self.out[i], self.dact[i])
            o.hadamard(self.sens[i], self.dact[i])
            prev = self.out[i - 1] if i >= 1 else self.x
            o.gemm(None, self.sens[i], 1, prev, 0, None, self.gW[i], packed=True)
            o.bias_step(self.sens[i], self.gb[i])
            if i >= 1:
                o.gemm(None, self.sens[i], 0, self.W[i], 0, None, self.sens[i - 1], packed=True)
        for W, gW, b, gb in zip(self.W, self.gW, self.b, self.gb):
            o.axpy(gW, -1.0, W)  # GD lr = 1 -> alpha = -1 -> SUBSTRACT kernel
            o.axpy(gb, -1.0, b)
