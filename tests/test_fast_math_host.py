"""-m "not gpu": the branch-free transcendental replacements of wekua_b200/csrc/common.cuh (f64 tanh / sigmoid / cosh / tan /
sin / cos, f32 sin / cos / tan / cosh), checked WITHOUT a GPU: the device functions are plain fma / add / mul sequences, so their
text is compiled for the host with a small shim (CUDA intrinsics -> libm, MUFU.RCP64H -> a 18-bit reciprocal, the libdevice
slow paths -> libm) and compared with long-double libm over the ranges the GPU tests use.  This pins the ALGORITHM (constants,
polynomial coefficients, operation order); tests/test_gpu_stream_kernels.py pins the device result."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = r'''
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
static double rcp_approx(double d) {  // MUFU.RCP64H: reads the high word, returns ~20 bits; modelled with 18
    uint64_t u; memcpy(&u, &d, 8); u &= 0xffffffff00000000ull; double t; memcpy(&t, &u, 8);
    double r = 1.0 / t; if (!(std::fabs(r) >= 2.3e-308)) r = 0; memcpy(&u, &r, 8); u &= ~((1ull << 34) - 1); memcpy(&r, &u, 8); return r; }
static float rcp_approx_f(float d) { return 1.0f / d; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline int __double2loint(double t) { int64_t i; memcpy(&i, &t, 8); return (int)(int32_t)(i & 0xffffffff); }
static inline int __double2hiint(double t) { int64_t i; memcpy(&i, &t, 8); return (int)(int32_t)(i >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
using std::fma; using std::fabs; using std::copysign;
'''

MAIN = r'''
static double worst[10];
static void upd(int k, long double got, long double want) {
    if (want == 0 || !std::isfinite((double)want)) return;
    const double e = (double)fabsl((got - want) / want);
    if (e > worst[k]) worst[k] = e;
}
int main() {
    srand(1);
    for (int i = 0; i < 3000000; i++) {
        const double u = rand() / (double)RAND_MAX;
        double x;
        switch (i % 4) { case 0: x = (u * 2 - 1) * 1.5; break; case 1: x = (u * 2 - 1) * 25; break;
                         case 2: x = ldexp(u, -(rand() % 60)); break; default: x = (u * 2 - 1) * 700; }
        upd(0, wk_tanh_f64(x), tanhl(x));
        const long double sg = 1.0L / (1.0L + expl(-(long double)x));
        if (sg > 1e-290L) upd(1, wk_sigmoid_f64(x), sg);
        upd(2, wk_cosh_f64(x), coshl(x));
        const double y = (i % 4 == 3) ? (u * 2 - 1) * 99000.0 : x;   // trig: up to the libdevice hand-over
        upd(3, wk_tan_f64(y), tanl(y));
        if (std::fabs(y) <= 100) { upd(4, wk_sin_f64(y), sinl(y)); upd(5, wk_cos_f64(y), cosl(y)); }
        const float f = (float)y;
        upd(6, wk_tan_f32(f), tan((double)f));
        if (std::fabs(y) <= 100) { upd(7, wk_sin_f32(f), sin((double)f)); upd(8, wk_cos_f32(f), cos((double)f)); }
        if (std::fabs(x) < 89.0) upd(9, wk_cosh_f32((float)x), cosh((double)(float)x));
    }
    for (int k = 0; k < 10; k++) printf("%.6g\n", worst[k]);
    printf("%d\n", (int)(wk_tanh_f64(INFINITY) == 1.0 && wk_tanh_f64(-INFINITY) == -1.0 && wk_tanh_f64(1e-300) == 1e-300
                         && wk_sigmoid_f64(800.0) == 1.0 && wk_sigmoid_f64(-800.0) < 1e-300 && std::isinf(wk_cosh_f64(711.0))
                         && wk_cosh_f64(0.0) == 1.0 && std::isnan(wk_tanh_f64(NAN)) && std::isnan(wk_sigmoid_f64(NAN))
                         && std::signbit(wk_tan_f64(-0.0)) && std::signbit(wk_sin_f64(-0.0)) && wk_cos_f64(0.0) == 1.0
                         && std::signbit(wk_sin_f32(-0.0f)) && std::signbit(wk_tan_f32(-0.0f)) && wk_cos_f32(0.0f) == 1.0f
                         && wk_cosh_f32(0.0f) == 1.0f && std::isinf(wk_cosh_f32(89.5f)) && std::isinf(wk_cosh_f32(-INFINITY))
                         && std::isfinite(wk_cosh_f32(89.4f)) && std::isnan(wk_cosh_f32(NAN))));
    return 0;
}
'''


def _device_text():
    src = open(os.path.join(ROOT, "wekua_b200", "csrc", "common.cuh")).read()
    start = src.index("#define WK_EXP_ROW(sc)")
    end = src.index("#endif", start)
    body = src[start:end]
    body = body.replace("static __device__ __noinline__", "static").replace("__device__ __forceinline__", "static inline")
    body = body.replace("static __constant__", "static const")
    body = body.replace('asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));', "rc = rcp_approx(den);")
    body = re.sub(r'asm\("rcp\.approx\.ftz\.f32 %0, %1;" : "=f"\(rc\) : "f"\((\w+)\)\);', r"rc = rcp_approx_f(\1);", body)
    assert "asm" not in body, "a device-only instruction of common.cuh has no host model in this test"
    return body


def test_fast_math_matches_libm(tmp_path):
    cpp = tmp_path / "fm.cpp"
    cpp.write_text(SHIM + _device_text() + MAIN)
    exe = tmp_path / "fm"
    r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", str(exe), str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300).stdout.split()
    worst = [float(v) for v in out[:10]]
    e64, e32 = 2.220446049250313e-16, 1.1920928955078125e-07
    names = ["tanh", "sigmoid", "cosh", "tan", "sin", "cos", "tanf", "sinf", "cosf", "coshf"]
    limits = [2 * e64, 2 * e64, 2 * e64, 3 * e64, 2 * e64, 2 * e64, 4 * e32, 2.5 * e32, 2.5 * e32, 2.5 * e32]
    for n, w, lim in zip(names, worst, limits):
        assert w <= lim, (n, w / (e64 if not n.endswith("f") else e32))
    assert out[10] == "1", "limit / NaN / signed-zero behaviour"
