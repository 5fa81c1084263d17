"""The C++ host mirror (include/wekua.hpp) and the reference's programs rewritten against it compile and link against
the in-tree CUDA library (no GPU needed: nothing is executed)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_self_contained():
    src = '#include "wekua.hpp"\nint main() { return wekua::core::types::getTypeIndex<float>() == 8 ? 0 : 1; }\n'
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_examples_and_benchmarks_build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for exe in ("xor_neural_network", "bench_gemm", "bench_axpy"):
        assert os.path.exists(os.path.join(ROOT, "build", "bin", exe))
