#!/usr/bin/env python
"""bench.py -- headline benchmark of the wekua dense-BLAS hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n SIZE] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs 4/5): f32 GEMM, C = A.B (NN), square N = 32768, rows of A and C sharded over the
N GPUs of the box (N = 1: the whole product on one GPU), synthetic U[-1,1) matrices from the reference's own
counter-based PRNG (seeds 42/43).  One step = one blas.gemm over the resident operands; `value` = 2*N^3*steps /
time, whole job, CUDA-event timed on the queue's stream, max over ranks.  `e2e` is the same product through the
public API from HOST buffers (pinned H2D of A and B and D2H of C inside the timed region).  The `also` list carries
the other configs the metric names (f32/f64 GEMM N = 16384, f32/f64 axpy on 2^28 elements) with their own rooflines.

--impl reference times the reference's CPU implementation of the same product (the C restatement of
gemm_pack.cl + gemm_nxn_pack.cl built -O3 with OpenMP, all host cores) on a bounded sample (N = 2048 by default).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------ peaks
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


FP64_DMMA_PEAK_TFLOPS = 37.04  # profiles/mma_peak_fp64_r01.txt
# tcgen05.mma kind::tf32, MMA-only loop (operands resident in smem, TMEM accumulators) on this pool's B200,
# tools/mma_peak_tf32.cu -> profiles/mma_peak_tf32_r01.txt: 900.6-903.2 TF/s over ~3 s (power-capped clock), 1045 burst
TF32_MMA_PEAK_TFLOPS = {"sustained": 903.2, "burst": 1044.9}
# dram__bytes_read.sum + dram__bytes_write.sum per launch: read from the committed `ncu --set full` summaries under
# profiles/ (tools/ncu_summary.py output), matched by KERNEL NAME so a summary of an older kernel revision is not quoted
NCU_SUMMARIES = {
    # key: (profile file, regex the "== <kernel name>" header must match)
    "gemm_f32_n32768_1gpu": ("ncu_gemm_f32_r02m_n32768.txt", r"gemm_tf32x3_kernel<2, *(false|0), *(false|0)>"),
    "axpy_f32_2^28": ("ncu_axpy_f32_r02q.txt", r"map_vec_kernel<float, *2, *AxpyF"),
    "axpy_f64_2^28": ("ncu_axpy_f64_r02q.txt", r"map_vec_kernel<double, *2, *AxpyF"),
}
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_traffic_bytes(key):
    """dram read + write bytes of the first launch in the summary whose kernel name matches; None if there is no such
    capture in the tree (the roofline then carries traffic = null rather than a stale constant)"""
    import re

    if key not in NCU_SUMMARIES:
        return None
    fname, pat = NCU_SUMMARIES[key]
    path = os.path.join(ROOT, "profiles", fname)
    if not os.path.exists(path):
        return None
    inside, got = False, {}
    for line in open(path):
        if line.startswith("== "):
            if inside and len(got) == 2:
                break
            inside, got = re.search(pat, line) is not None, {}
            continue
        if not inside:
            continue
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in _UNIT:
            got[f[0]] = float(f[1].replace(",", "")) * _UNIT[f[2]]
    return sum(got.values()) if len(got) == 2 else None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx, self.lines, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1])
                mx = float(f[2])
            except ValueError:
                continue
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than one sample: take everything we saw
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ helpers
def ev_record(wk, pipe):
    ev = C.c_void_p()
    wk.capi.check(wk.capi.lib().wk_event_record(pipe.q, C.byref(ev)))
    return ev


def ev_ms(wk, a, b):
    ms = C.c_float()
    wk.capi.check(wk.capi.lib().wk_event_wait(b))
    wk.capi.check(wk.capi.lib().wk_event_elapsed_ms(a, b, C.byref(ms)))
    return float(ms.value)


def timed(wk, pipe, fn, steps, warmup, barrier=None):
    """W untimed steps, then exactly K steps bracketed by sync (+ barrier) on both sides; device time via events"""
    for _ in range(warmup):
        fn()
    pipe.wait_and_cleanup()
    if barrier:
        barrier()
    t0 = time.time()
    l0 = wk.capi.launch_count()
    e0 = ev_record(wk, pipe)
    for _ in range(steps):
        fn()
    e1 = ev_record(wk, pipe)
    ms = ev_ms(wk, e0, e1)
    pipe.wait_and_cleanup()
    launches = wk.capi.launch_count() - l0
    if barrier:
        barrier()
    t1 = time.time()
    for e in (e0, e1):
        wk.capi.lib().wk_event_release(e)
    return ms, launches, t0, t1


def pinned_array(wk, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    wk.capi.check(wk.capi.lib().wk_host_alloc(n, C.byref(p)))
    buf = (C.c_char * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def numa_local_host_memory(torch, local_rank):
    """Place this process's future pinned host buffers (and, where the cpuset allows, its threads) on the NUMA node the
    GPU hangs off: with one process per GPU the default policy puts every rank's staging memory wherever the process
    happens to run, and half of the GPUs then DMA across the socket interconnect.  Best effort: returns what was done."""
    if os.environ.get("WK_NUMA", "1") == "0":
        return {"numa": "off"}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"numa": "single-node", "bdf": bdf}
        done = {"numa": node, "bdf": bdf}
        libc = C.CDLL(None, use_errno=True)
        mask = C.c_ulong(1 << node)
        rc = libc.syscall(238, 1, C.byref(mask), 65)  # set_mempolicy(MPOL_PREFERRED, {node})
        done["mempolicy"] = "preferred" if rc == 0 else f"errno {C.get_errno()}"
        try:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
                done["cpus"] = len(allowed)
        except OSError:
            pass
        return done
    except Exception as e:  # no sysfs entry, no such attribute, syscall refused: keep the default placement
        return {"numa": f"unavailable ({type(e).__name__})"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_gemm(n, dtype, steps, warmup, vw_f32=None):
    """the restated reference CPU path: PackedTensors + gemm_pack.cl x2 + gemm_nxn_pack.cl per call
    (benchmark/gemm.zig:150-204 re-packs on every call), -O3, OpenMP over work-items, all host cores."""
    from oracle import pyoracle as o

    fast = True
    lib = o.lib(fast=fast)
    if vw_f32 is None:
        vw_f32 = 16 if o._cpu_has("avx512f") else 8
    dev = o.device("cpu", vw_f32)
    a = o.OTensor(dev, dtype, (n, n), fast=fast).uniform(42, -1, 1)
    b = o.OTensor(dev, dtype, (n, n), fast=fast).uniform(43, -1, 1)
    c = o.OTensor(dev, dtype, (n, n), fast=fast)
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        o.gemm(None, a, 0, b, 0, None, c, packed=True, fast=fast)
        if i >= warmup:
            times.append(time.perf_counter() - t)
    sec = float(np.mean(times))
    return {"tflops": 2.0 * n ** 3 / sec / 1e12, "sec_per_call": sec, "cores": int(lib.wko_num_threads()), "n": n,
            "vw": vw_f32, "tile": 2 << c.layout.gemm_algorithm}


def cpu_reference_axpy(n, dtype, reps=3):
    from oracle import pyoracle as o

    dev = o.device("cpu", 16 if o._cpu_has("avx512f") else 8)
    x = o.OTensor(dev, dtype, (n,), fast=True).uniform(42)
    y = o.OTensor(dev, dtype, (n,), fast=True).uniform(43)
    o.axpy(x, 0.5, y, fast=True)
    t = time.perf_counter()
    for _ in range(reps):
        o.axpy(x, 0.5, y, fast=True)
    sec = (time.perf_counter() - t) / reps
    return 3.0 * n * np.dtype(dtype).itemsize / sec / 1e9


def workload_config(N, g, gather):
    """the `config` object: the SAME for our arm and for --impl reference (the driver compares them)"""
    return {"workload": f"f32 GEMM NN N={N}, rows of A/C sharded over {g} GPU(s), B replicated "
                        f"(BASELINE config 5); gather={gather}",
            "N": N, "parallelism": f"rowshard{g}", "l2": "inputs_exceed_l2"}


def host_threads():
    """Threads the CPU arm may use.  torch.distributed.run exports OMP_NUM_THREADS=1 into every rank; the reference arm
    is one process on the box's host cores, so that setting is undone BEFORE the OpenMP runtime of the oracle library
    starts (it reads the variable once, at load)."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    return max(1, n)


REF_SAMPLE_SIZES = (8192, 6144, 4096, 3072, 2048, 1024)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.pop("OMP_PROC_BIND", None)
    from oracle import pyoracle as o

    lib = o.lib(fast=True)
    if hasattr(lib, "wko_set_num_threads"):
        lib.wko_set_num_threads(threads)
    # bounded sample: the largest N whose (warmup + steps) calls fit the time budget at the rate a probe call measures,
    # so the arm ends within about a minute whatever the core count (N = 32768 itself would be ~150 s per call)
    budget = float(os.environ.get("WK_REF_BUDGET_S", 60.0))
    warmup = max(args.warmup, 1)
    probe = cpu_reference_gemm(1024, np.float32, 2, 1)
    rate = probe["tflops"] * 1e12
    cands = [c_ for c_ in REF_SAMPLE_SIZES if c_ <= args.ref_n] or [args.ref_n]  # --ref-n caps the sample
    n = cands[-1]
    for cand in cands:
        # big problems run ~1.3x faster per flop than the 1024 probe at best; stay conservative and take the probe rate
        if (warmup + args.steps) * 2.0 * cand ** 3 / rate <= budget:
            n = cand
            break
    r = cpu_reference_gemm(n, np.float32, args.steps, warmup)
    sample = (f"f32 NN GEMM N={n} (a {n}^3 sub-problem of the N={args.n} workload, sized to a {budget:.0f} s budget), restated "
              f"gemm_pack.cl + gemm_nxn_pack.cl, {r['tile']}x{r['tile']} tiles, vector width {r['vw']}, re-packed every call")
    g = int(os.environ.get("WORLD_SIZE", args.gpus))
    line = {
        "impl": "reference", "metric": "gemm_f32_tflops", "value": r["tflops"], "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec_per_call"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.n, g, args.gather if g > 1 else "none"),
        "cpu_baseline": {"value": r["tflops"], "unit": "TFLOP/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["tflops"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def gemm_roofline(tflops, dtype, peaks, sustained, traffic=None):
    bf16 = peaks["bf16_tflops_sustained" if sustained else "bf16_tflops"]
    extra = {}
    if dtype == "f32":
        # MEASURED_PEAKS.json has no TF32 entry, and bf16/2/3 sits BELOW what this kernel reaches (frac > 1: at equal
        # power TF32 MMAs clock higher than cuBLAS bf16), so the denominator is the tensor pipe itself, measured
        which = "sustained" if sustained else "burst"
        peak = TF32_MMA_PEAK_TFLOPS[which] / 3.0
        basis = (f"tcgen05.mma kind::tf32 MMA-only loop {which} {TF32_MMA_PEAK_TFLOPS[which]:.0f} TF/s measured on this pool's "
                 "B200 (tools/mma_peak_tf32.cu, profiles/mma_peak_tf32_r01.txt) / 3 (3xTF32 passes per useful flop); "
                 "MEASURED_PEAKS.json has no TF32 entry")
        extra = {"peak_from_bf16": bf16 / 6.0, "frac_of_bf16_derived": tflops / (bf16 / 6.0),
                 "bf16_basis": f"{peaks['_source']} cuBLAS bf16 {which} {bf16:.0f} TF/s / 2 / 3"}
    else:
        peak = FP64_DMMA_PEAK_TFLOPS
        basis = ("DMMA m16n8k16 register-only loop measured on this pool's B200 by tools/mma_peak.cu "
                 "(profiles/mma_peak_fp64_r01.txt; MEASURED_PEAKS.json has no FP64 entry; spec 40 TF/s)")
    return {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak,
            "traffic": traffic, "peak_basis": basis, **extra}


def hbm_roofline(gbs, peaks, traffic=None):
    # the denominator is the copy bandwidth MEASURED_PEAKS.json records (a driver copy); a streaming kernel launched as one chunk
    # per CTA can exceed it (frac > 1), so the fraction of the 8 TB/s HBM3e specification is carried next to it
    return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": traffic, "peak_basis": f"{peaks['_source']} copy bandwidth", "frac_of_hbm3e_spec_8000": gbs / 8000.0}


def bench_axpy(wk, ctx, pipe, dtype, n, steps, warmup, peaks, barrier=None, reduce_max=None, world=1):
    """benchmark/axpy.zig:119-162: alternating axpy(x, a, y) / axpy(y, a, x) on N-element vectors.  With world > 1 every
    rank streams its OWN n-element vectors (the path shards with no exchange step: weak scaling), timed between barriers
    as the max over ranks; value = all ranks' bytes / that time."""
    x = wk.Tensor.alloc(ctx, pipe, (n,), dtype)
    y = wk.Tensor.alloc(ctx, pipe, (n,), dtype)
    wk.tensor.random.uniform(pipe, x, 42)
    wk.tensor.random.uniform(pipe, y, 43)
    state = {"i": 0}
    alphas = np.random.default_rng(1234).uniform(-1, 1, 64) / np.sqrt(2)  # |alpha| < 1 keeps values finite

    def step():
        i = state["i"]
        a = float(alphas[i % 64])
        if i % 2 == 0:
            wk.blas.axpy(pipe, x, a, y)
        else:
            wk.blas.axpy(pipe, y, a, x)
        state["i"] = i + 1

    ms, launches, _, _ = timed(wk, pipe, step, steps, warmup, barrier)
    if reduce_max is not None:
        ms = reduce_max(ms)
    bytes_per = 3.0 * n * np.dtype(dtype).itemsize
    gbs = bytes_per * steps / (ms * 1e-3) / 1e9
    x.release(pipe)
    y.release(pipe)
    name = "f32" if np.dtype(dtype) == np.float32 else "f64"
    per_rank = f", per GPU x {world} GPUs" if world > 1 else ""
    return {"metric": f"axpy_{name}_gbs", "value": gbs * world, "unit": "GB/s", "ms_per_step": ms / steps, "dtype": name,
            "n_gpus": world, "scaling": "weak",
            "config": {"workload": f"{name} axpy, 2^{int(np.log2(n))} elements{per_rank}, alternating x/y (benchmark/axpy.zig)",
                       "l2": "inputs_exceed_l2"},
            "gpu_launches": launches * world,
            "roofline": hbm_roofline(gbs, peaks, ncu_traffic_bytes(f"axpy_{name}_2^{int(np.log2(n))}"))}


def bench_gemm_graph(wk, ctx, pipe, dtype, m, n, k, peaks, reps=20, launches=10):
    """small problems (BASELINE config 1: 1024^3): `reps` gemm calls captured in one CUDA graph, the graph replayed
    `launches` times between two events -- kernel time without the host's per-call overhead, as tools/gemm_small_time.py"""
    a, b, c = (wk.Tensor.alloc(ctx, pipe, s_, dtype) for s_ in ((m, k), (k, n), (m, n)))
    wk.tensor.random.uniform(pipe, a, 42, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    for _ in range(3):
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    pipe.wait_and_cleanup()
    pipe.begin_capture()
    for _ in range(reps):
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    graph = pipe.end_capture()
    for _ in range(3):
        graph.launch(pipe)
    pipe.wait_and_cleanup()
    l0 = wk.capi.launch_count()
    e0 = ev_record(wk, pipe)
    for _ in range(launches):
        graph.launch(pipe)
    e1 = ev_record(wk, pipe)
    ms = ev_ms(wk, e0, e1)
    for e in (e0, e1):
        wk.capi.lib().wk_event_release(e)
    graph.release()
    for t in (a, b, c):
        t.release(pipe)
    us = ms * 1e3 / (reps * launches)
    tf = 2.0 * m * n * k / (us * 1e-6) / 1e12
    name = "f32" if np.dtype(dtype) == np.float32 else "f64"
    r = gemm_roofline(tf, name, peaks, sustained=False)
    r.update({"us_per_call": us, "how": f"{reps} calls per CUDA graph x {launches} replays, L2-resident operands (the problem is 12 MiB)"})
    return r


# tcgen05.mma kind::i8 (u8 x u8 -> s32), MMA-only loop on this pool's B200: tools/mma_peak_i8.cu -> profiles/mma_peak_i8_r02.txt
I8_MMA_PEAK_TOPS = {"sustained": 4036.7, "burst": 4534.9}
BYTE_GEMMS = {1: 1, 2: 3, 4: 10, 8: 36}  # byte-plane products per element product: W (W + 1) / 2


def bench_gemm_int(wk, ctx, pipe, dtype, n, steps, warmup):
    """integer GEMM (src/blas/gemm.zig:834-874 for i8..u64; gemm_nxn_gpu.cl:82-319): exact mod 2^bits on tcgen05.mma kind::i8
    over byte planes (csrc/gemm_i8_tc.cu).  Tera-ops/s = 2 N^3 / t of the ELEMENT product; the roofline is the measured
    kind::i8 MMA ceiling divided by the byte GEMMs one element product costs (1 / 3 / 10 / 36)."""
    a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
    wk.tensor.random.uniform(pipe, a, 42)
    wk.tensor.random.uniform(pipe, b, 43)
    ms, launches, _, _ = timed(wk, pipe, lambda: wk.blas.gemm(pipe, None, a, 0, b, 1, None, c), steps, warmup)
    for t in (a, b, c):
        t.release(pipe)
    tops = 2.0 * n ** 3 * steps / (ms * 1e-3) / 1e12
    w = np.dtype(dtype).itemsize
    peak = I8_MMA_PEAK_TOPS["burst"] / BYTE_GEMMS[w]
    return {"bound": "tensor", "achieved": tops, "peak": peak, "unit": "Top/s", "frac": tops / peak, "traffic": None,
            "ms_per_step": ms / steps, "gpu_launches": launches,
            "workload": f"{np.dtype(dtype).name} GEMM NT N={n}: {BYTE_GEMMS[w]} byte GEMM(s) on tcgen05 kind::i8, bit-exact mod 2^{8 * w}",
            "peak_basis": f"kind::i8 MMA-only loop burst {I8_MMA_PEAK_TOPS['burst']:.0f} Top/s (profiles/mma_peak_i8_r02.txt) / {BYTE_GEMMS[w]}"}


def bench_gemm_single(wk, ctx, pipe, dtype, n, steps, warmup, peaks, op_a=0, op_b=0):
    a = wk.Tensor.alloc(ctx, pipe, (n, n), dtype)
    b = wk.Tensor.alloc(ctx, pipe, (n, n), dtype)
    c = wk.Tensor.alloc(ctx, pipe, (n, n), dtype)
    wk.tensor.random.uniform(pipe, a, 42, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    ms, launches, _, _ = timed(wk, pipe, lambda: wk.blas.gemm(pipe, None, a, op_a, b, op_b, None, c), steps, warmup)
    tf = 2.0 * n ** 3 * steps / (ms * 1e-3) / 1e12
    for t in (a, b, c):
        t.release(pipe)
    name = "f32" if np.dtype(dtype) == np.float32 else "f64"
    return {"metric": f"gemm_{name}_tflops", "value": tf, "unit": "TFLOP/s", "ms_per_step": ms / steps, "dtype": name,
            "config": {"workload": f"{name} GEMM {'NT'[op_a]}{'NT'[op_b]} N={n}, 1 GPU", "l2": "inputs_exceed_l2"},
            "gpu_launches": launches, "roofline": gemm_roofline(tf, name, peaks, sustained=True)}


def freivalds_gather_check(wk, ctx, pipe, a_blk, b, c_full, N, rows, row0, rank, world, rowshard):
    """Every rank proves that the C it holds after the fused all-gather is A.B -- all of it, not only its own rows.
    W [N, N/256] has random signs in rows [256 b, 256 b + 256) of column b and zeros elsewhere, so C.W sums every 256-column
    tile strip of C exactly once; it is compared with A.(B.W), whose rows come from the rank that owns them (exchanged over
    the host).  Short sums keep the f32 rounding of the check itself small (the tensor core's fp32 accumulation truncates, so a
    32768-term checksum drifts by K.eps.|sum|): the tolerance below is ~10x smaller than ONE typical entry of C (|c| ~
    sqrt(N/9) = 60), so a single lost element -- let alone a 128-byte row segment or a chunk -- is caught."""
    rng = np.random.default_rng(99)
    nb = (N + 255) // 256
    W = np.zeros((N, nb), dtype=np.float32)
    signs = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=N)
    W[np.arange(N), np.arange(N) // 256] = signs
    w = wk.Tensor.alloc(ctx, pipe, (N, nb), np.float32)
    y = wk.Tensor.alloc(ctx, pipe, (N, nb), np.float32)
    z = wk.Tensor.alloc(ctx, pipe, (rows, nb), np.float32)
    t = wk.Tensor.alloc(ctx, pipe, (N, nb), np.float32)
    wk.tensor.memory.read_from_buffer(pipe, w, W)
    wk.blas.gemm(pipe, None, b, 0, w, 0, None, y)          # Y = B.W
    wk.blas.gemm(pipe, None, a_blk, 0, y, 0, None, z)      # my rows of A.Y
    wk.blas.gemm(pipe, None, c_full, 0, w, 0, None, t)     # C.W from MY copy of the gathered C
    zh = wk.tensor.memory.to_numpy(pipe, z).astype(np.float64)
    th = wk.tensor.memory.to_numpy(pipe, t).astype(np.float64)
    for x in (w, y, z, t):
        x.release(pipe)
    z_all = np.concatenate(rowshard.exchange(zh, world), axis=0)  # rank order = row order
    eps = float(np.finfo(np.float32).eps)
    bound = 0.5 * N * eps * float(np.abs(z_all).max())  # A.Y is the one long (K = N) accumulation left in the check
    err = float(np.abs(th - z_all).max())
    ok = bool(np.all(np.isfinite(th)) and err <= bound and float(np.abs(z_all).max()) > 1.0)
    oks = rowshard.exchange((ok, err, bound), world)
    return {"ok": bool(all(o_[0] for o_ in oks)),
            "method": "Freivalds per 256-column strip: C.W == A.(B.W), W = random signs block-diagonal over the strips, every rank's own copy of C",
            "max_err": max(o_[1] for o_ in oks), "bound": min(o_[2] for o_ in oks)}


def host_link_ceiling(g, h2d_bytes, d2h_bytes):
    """the measured host<->device copy ceiling of this box class (tools/pcie_peak.py -> profiles/pcie_peak_r02_g{g}.json): time
    the step's copies alone would take, one direction after the other at the all-ranks rates, and both at once"""
    path = os.path.join(ROOT, "profiles", f"pcie_peak_r02_g{g}.json")
    if not os.path.exists(path):
        return None
    d = json.loads(open(path).read().strip().splitlines()[-1])
    serial_ms = (h2d_bytes / (d["h2d"]["aggregate_gbs"] * 1e9) + d2h_bytes / (d["d2h"]["aggregate_gbs"] * 1e9)) * 1e3
    both_ms = (h2d_bytes + d2h_bytes) / (d["both"]["aggregate_gbs"] * 1e9) * 1e3
    return {"h2d_gbs_per_gpu": d["h2d"]["per_gpu_gbs_each_direction"], "d2h_gbs_per_gpu": d["d2h"]["per_gpu_gbs_each_direction"],
            "both_gbs_per_gpu_each_way": d["both"]["per_gpu_gbs_each_direction"], "copies_alone_ms_serial": serial_ms,
            "copies_alone_ms_concurrent": both_ms, "source": f"profiles/pcie_peak_r02_g{g}.json"}


def run_ours(args):
    import torch  # plumbing only: torch.distributed rendezvous / barrier / max-reduce, never on the compute path

    import wekua_b200 as wk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = None
    if world > 1:
        torch.cuda.set_device(local_rank)
        numa = numa_local_host_memory(torch, local_rank)
        if os.environ.get("WK_DEBUG"):
            print(f"[bench] rank {rank}: host memory placement {numa}", file=sys.stderr, flush=True)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run", file=sys.stderr)
        args.gpus = world
    peaks = measured_peaks()
    ctx = wk.Context.init([local_rank])
    pipe = wk.Pipeline.init(ctx.command_queues[0])
    N = args.n
    g = world
    rows = N // g  # BASELINE config 5: contiguous row blocks of A and C, B replicated
    row0 = rank * rows
    if rank == g - 1:
        rows = N - row0

    def barrier():
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # ---------------- resident operands: A row block, full B, full C (the gather target, IPC-mapped on every peer)
    from wekua_b200 import rowshard

    dt = np.float32
    a = wk.Tensor.alloc(ctx, pipe, (rows, N), dt)
    b = wk.Tensor.alloc(ctx, pipe, (N, N), dt)
    shard = rowshard.RowShardedC(ctx, pipe, N, N, dt, rank, world)
    c_full = shard.c
    c_blk = shard.block(pipe)
    # reference PRNG (uniform.cl); every rank seeds its own row block of A, B is identical everywhere
    wk.tensor.random.uniform(pipe, a, 42 + 1000 * rank, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    pipe.wait_and_cleanup()

    gather = args.gather if world > 1 else "none"
    c_torch = None
    if world > 1:
        class _Cai:
            __cuda_array_interface__ = {"shape": (N * c_full.row_pitch,), "typestr": "<f4", "data": (c_full.buffer, False),
                                        "version": 3}
        c_torch = torch.as_tensor(_Cai(), device=f"cuda:{local_rank}")

    def make_step(mode):
        def step():
            if mode == "fused":  # the GEMM epilogue stores every C tile to all peers over NVLink: no collective
                shard.gemm(pipe, None, a, 0, b, 0, None)
            else:
                wk.blas.gemm(pipe, None, a, 0, b, 0, None, c_blk)
                if mode == "nccl":
                    pipe.command_queue.finish()
                    chunk = (N // g) * c_full.row_pitch
                    torch.distributed.all_gather_into_tensor(c_torch[: chunk * g], c_torch[rank * chunk:(rank + 1) * chunk])
                    torch.cuda.synchronize()
        return step

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=f"cuda:{local_rank}", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    step = make_step(gather)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, t0, t1 = timed(wk, pipe, step, args.steps, args.warmup, barrier)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms = max_over_ranks(ms)
    if world > 1:
        lt = torch.tensor([launches], device=f"cuda:{local_rank}", dtype=torch.int64)
        torch.distributed.all_reduce(lt)
        launches = int(lt.item())
    flops = 2.0 * N * N * N
    tflops = flops * args.steps / (ms * 1e-3) / 1e12
    per_gpu_tflops = tflops / g

    # the other two ways of finishing the step (reported beside the headline), and a check that the gathered C is the
    # same matrix on every rank: the deterministic two-stage sum of the full C must agree bit for bit
    variants, gather_check = {}, None
    if world > 1:
        for mode in ("none", "nccl", "fused"):
            if mode == gather:
                continue
            vms, _, _, _ = timed(wk, pipe, make_step(mode), max(2, args.steps // 2), 2, barrier)
            vms = max_over_ranks(vms)
            variants[mode] = flops * max(2, args.steps // 2) / (vms * 1e-3) / 1e12
        wk.capi.check(wk.capi.lib().wk_memset_zero(pipe.q, c_full.ptr, c_full.size))
        pipe.wait_and_cleanup()
        barrier()
        make_step("fused")()
        pipe.wait_and_cleanup()
        barrier()
        gather_check = freivalds_gather_check(wk, ctx, pipe, a, b, c_full, N, rows, row0, rank, world, rowshard)

    # ---------------- e2e: same product through the public API from HOST buffers (pinned), copies inside the region.
    # The user-level recipe for overlap in wekua's model is "several command queues in one context" + event wait lists
    # (core/context.zig:145-178, pipeline.zig:35-45): queue 1 uploads B then row panels of A, queue 0 multiplies each
    # panel as soon as its upload event fires, queue 2 downloads each finished panel of C.
    e2e = None
    if not args.no_e2e and world > 1:
        # N > 1: B is the same matrix in every process's host memory, so each rank uploads only its 1/g share of B's rows
        # over PCIe and pushes it into the peers' B over NVLink (rowshard.ReplicatedOperand) while its row panels of A
        # follow on the upload queue; host-side barriers (a gloo group: no device synchronisation) fence "every share has
        # landed everywhere" and "everybody has finished reading B" around the products.
        gloo = torch.distributed.new_group(backend="gloo")
        hostbar = lambda: torch.distributed.barrier(group=gloo)
        ctx4 = wk.Context.init([local_rank] * 4)
        p_mm, p_up, p_dn, p_push = (wk.Pipeline.init(q) for q in ctx4.command_queues)
        b4 = wk.Tensor.wrap(ctx4, p_mm, (N, N), dt, b.buffer)  # the same HBM, as a tensor of the 4-queue context
        # B lives in HBM as NJ dense column panels: panel j is uploaded (1/g per rank), pushed to the peers, fenced and
        # multiplied while the later panels are still crossing PCIe -- the step is upload-bound, so what matters is that
        # little work is left after the LAST piece has arrived.  A (the rank's row block) follows B's first panel.
        NJ = int(os.environ.get("WK_E2E_NJ_MULTI", 8))
        if N % (2 * NJ) or b4.row_pitch != N:
            NJ = 1
        nc = N // NJ
        rep = rowshard.ReplicatedOperand(ctx4, p_mm, b4, rank, world, col_panels=NJ)
        ha, pa = pinned_array(wk, (rows, N), dt)
        pins = [pa]
        hbs, hcs = [], []
        for _ in range(NJ):
            hb_, pb_ = pinned_array(wk, (rep.rows, nc), dt)
            hc_, pc_ = pinned_array(wk, (rows, nc), dt)
            hbs.append(hb_); hcs.append(hc_); pins += [pb_, pc_]
        P = 512
        blk = np.random.default_rng(7).uniform(-1, 1, (P, N)).astype(dt)  # A rows repeat blk, B rows repeat blk reversed
        for r0 in range(0, rows, P):
            ha[r0:r0 + P] = blk[: min(P, rows - r0)]
        brev = blk[::-1]
        for r0 in range(0, rep.rows, P):  # global row r of B is brev[r % P]
            idx = (rep.row0 + r0 + np.arange(min(P, rep.rows - r0))) % P
            for jj in range(NJ):
                hbs[jj][r0:r0 + len(idx)] = brev[idx, jj * nc:(jj + 1) * nc]
        n_panels = max(1, min(int(os.environ.get("WK_E2E_PANELS", 8)), rows // 4096))
        bounds = [rows * i // n_panels for i in range(n_panels + 1)]
        es = np.dtype(dt).itemsize
        a_pan = [wk.Tensor.wrap(ctx4, p_mm, (bounds[i + 1] - bounds[i], N), dt, a.buffer + bounds[i] * a.row_pitch * es)
                 for i in range(n_panels)]
        c_pan = [[wk.Tensor.wrap(ctx4, p_mm, (bounds[i + 1] - bounds[i], nc), dt,
                                 c_blk.buffer + (bounds[i] * c_blk.row_pitch + jj * nc) * es, row_pitch=c_blk.row_pitch)
                  for i in range(n_panels)] for jj in range(NJ)]
        rfb, wtb = wk.tensor.memory.read_from_buffer, wk.tensor.memory.write_to_buffer

        def e2e_step():
            p_mm.wait_and_cleanup()  # my products of the previous step have read B ...
            hostbar()                # ... and so have everybody else's: B may be overwritten
            pushed, a_ev = [], []
            for jj in range(NJ):
                rep.upload_share(p_up, hbs[jj], jj)
                p_push.wait_for(p_up.record_event())
                rep.push_share(p_push, jj)
                pushed.append(p_push.record_event())
                if jj == 0:
                    for i in range(n_panels):
                        rfb(p_up, a_pan[i], ha[bounds[i]:bounds[i + 1]])
                        a_ev.append(p_up.record_event())
            for jj in range(NJ):
                wk.capi.check(wk.capi.lib().wk_event_wait(pushed[jj]))  # my share of panel jj is in every peer's B ...
                hostbar()                                                # ... and every peer's share is in mine
                for i in range(n_panels):
                    if jj == 0:
                        p_mm.wait_for(a_ev[i])
                    wk.blas.gemm(p_mm, None, a_pan[i], 0, rep.panels[jj], 0, None, c_pan[jj][i])
                    p_dn.wait_for(p_mm.record_event())
                    wtb(p_dn, c_pan[jj][i], hcs[jj][bounds[i]:bounds[i + 1]].reshape(-1))
            p_mm.wait_for(p_dn.record_event())  # the step ends when the last panel of C is on the host
            p_up.wait_for(p_mm.record_event())

        class _All4:
            command_queue = p_mm.command_queue
            q = p_mm.q

            @staticmethod
            def wait_and_cleanup():
                for p_ in (p_up, p_push, p_mm, p_dn):
                    p_.wait_and_cleanup()

        e_steps = max(1, min(args.steps, 10))
        ems, _, _, _ = timed(wk, _All4, e2e_step, e_steps, 1, barrier)
        ems = max_over_ranks(ems)
        for (i_, j_) in ((0, 0), (rows - 1, N - 1), (rows // 2, 17)):  # the product did reach the host
            bcol = np.tile(brev[:, j_].astype(np.float64), (N + P - 1) // P)[:N]
            want = float(ha[i_].astype(np.float64) @ bcol)
            got_ = float(hcs[j_ // nc][i_, j_ % nc])
            assert abs(got_ - want) <= 1e-3 * max(1.0, abs(want)), (rank, i_, j_, got_, want)
        e2e = {"value": flops * e_steps / (ems * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": int(2 * N * N * 4), "d2h_bytes_per_step": int(N * N * 4),
               "nvlink_bytes_per_step": int((g - 1) * N * N * 4), "host_memory": numa,
               "steps": e_steps, "ms_per_step": ems / e_steps,
               "how": (f"per rank, 4 queues on its device: B as {NJ} column panels, each uploaded 1/{g} per rank from pinned host "
                       f"memory and pushed to the {g - 1} peers over NVLink (IPC-mapped copies); the rank's A block ({n_panels} row "
                       "panel(s)) follows B's first panel; per column panel a host barrier, then multiply and download C")}
        barrier()
        rep.release(p_mm)
        for p in pins:
            wk.capi.lib().wk_host_free(p)
        ctx4.deinit()
    elif not args.no_e2e:
        ctx3 = wk.Context.init([local_rank] * 3)
        p_mm, p_up, p_dn = (wk.Pipeline.init(q) for q in ctx3.command_queues)
        NJ = int(os.environ.get("WK_E2E_NJ", 2)) if N % 512 == 0 else 1  # column panels of B and C (dense host arrays of their own)
        nc = N // NJ
        ha, pa = pinned_array(wk, (rows, N), dt)
        hbs, hcs, pins = [], [], [pa]
        for _ in range(NJ):
            hb_, pb_ = pinned_array(wk, (N, nc), dt)
            hc_, pc_ = pinned_array(wk, (rows, nc), dt)
            hbs.append(hb_); hcs.append(hc_); pins += [pb_, pc_]
        # random host data (constant operands would toggle fewer bits, draw less power and clock higher than real work)
        blk = np.random.default_rng(7).uniform(-1, 1, (min(rows, 512), N)).astype(dt)
        for r0 in range(0, rows, blk.shape[0]):
            ha[r0:r0 + blk.shape[0]] = blk[: min(blk.shape[0], rows - r0)]
        for jj in range(NJ):
            for r0 in range(0, N, blk.shape[0]):
                hbs[jj][r0:r0 + blk.shape[0]] = blk[::-1, jj * nc:(jj + 1) * nc][: min(blk.shape[0], N - r0)]
        n_panels = max(1, min(int(os.environ.get("WK_E2E_PANELS", 16)), rows // 1024))
        bounds = [rows * i // n_panels for i in range(n_panels + 1)]
        es = np.dtype(dt).itemsize
        a_pan = [wk.Tensor.wrap(ctx3, p_mm, (bounds[i + 1] - bounds[i], N), dt, a.buffer + bounds[i] * a.row_pitch * es)
                 for i in range(n_panels)]
        b_pan = [wk.Tensor.wrap(ctx3, p_mm, (N, nc), dt, b.buffer + jj * nc * es, row_pitch=b.row_pitch) for jj in range(NJ)]
        c_pan = [[wk.Tensor.wrap(ctx3, p_mm, (bounds[i + 1] - bounds[i], nc), dt,
                                 c_blk.buffer + (bounds[i] * c_blk.row_pitch + jj * nc) * es, row_pitch=c_blk.row_pitch)
                  for i in range(n_panels)] for jj in range(NJ)]
        # B's later column panels are uploaded in k-slices interleaved with the panels of A
        kb = [N * i // n_panels for i in range(n_panels + 1)]
        b_slices = [[wk.Tensor.wrap(ctx3, p_mm, (kb[i + 1] - kb[i], nc), dt, b.buffer + (kb[i] * b.row_pitch + jj * nc) * es,
                                    row_pitch=b.row_pitch) for i in range(n_panels)] for jj in range(NJ)]

        def e2e_step():
            rfb, wtb = wk.tensor.memory.read_from_buffer, wk.tensor.memory.write_to_buffer
            rfb(p_up, b_pan[0], hbs[0])
            for i in range(n_panels):
                rfb(p_up, a_pan[i], ha[bounds[i]:bounds[i + 1]])
                p_mm.wait_for(p_up.record_event())
                wk.blas.gemm(p_mm, None, a_pan[i], 0, b_pan[0], 0, None, c_pan[0][i])
                p_dn.wait_for(p_mm.record_event())
                wtb(p_dn, c_pan[0][i], hcs[0][bounds[i]:bounds[i + 1]].reshape(-1))
                for jj in range(1, NJ):
                    rfb(p_up, b_slices[jj][i], hbs[jj][kb[i]:kb[i + 1]])
            for jj in range(1, NJ):
                p_mm.wait_for(p_up.record_event())
                for i in range(n_panels):
                    wk.blas.gemm(p_mm, None, a_pan[i], 0, b_pan[jj], 0, None, c_pan[jj][i])
                    p_dn.wait_for(p_mm.record_event())
                    wtb(p_dn, c_pan[jj][i], hcs[jj][bounds[i]:bounds[i + 1]].reshape(-1))
            # the step ends when the last panel of C is on the host; the next step may not overwrite A/B before that
            p_mm.wait_for(p_dn.record_event())
            p_up.wait_for(p_mm.record_event())

        class _All:  # timed() records its events on the compute queue and synchronises all three
            command_queue = p_mm.command_queue
            q = p_mm.q

            @staticmethod
            def wait_and_cleanup():
                for p_ in (p_up, p_mm, p_dn):
                    p_.wait_and_cleanup()

        e_steps = max(1, min(args.steps, 10))
        ems, _, _, _ = timed(wk, _All, e2e_step, e_steps, 1, barrier)
        ems = max_over_ranks(ems)
        for (i_, j_) in ((0, 0), (rows - 1, N - 1), (rows // 2, 17)):  # the product did reach the host
            want = float(ha[i_].astype(np.float64) @ hbs[j_ // nc][:, j_ % nc].astype(np.float64))
            got_ = float(hcs[j_ // nc][i_, j_ % nc])
            assert abs(got_ - want) <= 1e-3 * max(1.0, abs(want)), (i_, j_, got_, want)
        e2e = {"value": flops * e_steps / (ems * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": int((rows * N + N * N) * 4 * g), "d2h_bytes_per_step": int(N * N * 4),
               "steps": e_steps, "ms_per_step": ems / e_steps,
               "how": (f"3 queues on one device: {NJ} column panels of B x {n_panels} row panels of A uploaded, multiplied "
                       "and downloaded in a pipeline")}
        for p in pins:
            wk.capi.lib().wk_host_free(p)
        ctx3.deinit()
    for t_ in (a, b):
        t_.release(pipe)
    if world > 1:
        barrier()  # nobody unmaps a peer's C while that peer may still be storing into it
    shard.release(pipe)

    # ---------------- secondary configs + cpu baseline (rank 0, N = 1 only)
    also, cpu = [], None
    layer_step = None
    roof_also = {}  # lives INSIDE `roofline` (the driver keeps that object; top-level extras are dropped)

    def keep(key, r_):
        rf = dict(r_["roofline"])
        rf.pop("peak_basis", None), rf.pop("bf16_basis", None)
        rf.update({"ms_per_step": r_["ms_per_step"], "workload": r_["config"]["workload"]})
        roof_also[key] = rf

    if world == 1 and not args.quick:
        s2 = max(3, min(args.steps, 10))
        for key, dt_, n_, st_ in (("gemm_f32_n16384", np.float32, 16384, s2),
                                  ("gemm_f64_n16384", np.float64, 16384 if not args.small else 4096, max(3, s2 // 2))):
            r_ = bench_gemm_single(wk, ctx, pipe, dt_, n_, st_, 3, peaks)
            also.append(r_)
            keep(key, r_)
        for key, dt_ in (("axpy_f32_2p28", np.float32), ("axpy_f64_2p28", np.float64)):
            r_ = bench_axpy(wk, ctx, pipe, dt_, 1 << 28, 50, 5, peaks)
            also.append(r_)
            keep(key, r_)
        # (the streaming kernels run before the power-hungry GEMM extras below: right after a power-capped burst the chip clocks lower)
        # the other HBM-bound kernels of a layer step, same method as tools/stream_sweep.py:
        # 2^27 elements per operand, CUDA events, algorithmic bytes / time against the measured copy bandwidth
        import importlib.util

        spec = importlib.util.spec_from_file_location("wk_stream_sweep", os.path.join(ROOT, "tools", "stream_sweep.py"))
        sw = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(sw)
        rows_, _ = sw.sweep(27, only=("scal,hadamard (math.dot),sum,sum (device scalar),dot_reduce,sigmoid,tanh,sigmoid_dev,tanh_dev,"
                                      "act_backward(sigmoid),bias_add,bias_step,mse (+dev),gdm,adagrad,rmsprop,adam"),
                            ctx=ctx, pipe=pipe, reps=10, warm=3, verbose=False)
        layer_step = [{"op": r_["op"], "dtype": r_["dtype"], "bytes_per_elem": r_["bytes_per_elem"], "gbs": round(r_["gbs"], 1),
                       "frac": round(r_["gbs"] / peaks["hbm_gbs"], 3)} for r_ in rows_]
        roof_also["layer_step_streaming"] = {
            "bound": "hbm", "unit": "frac of measured copy bandwidth", "n": 1 << 27, "peak": peaks["hbm_gbs"],
            "frac": {f"{r_['op']}_{r_['dtype']}": r_["frac"] for r_ in layer_step},
            "frac_min": min(r_["frac"] for r_ in layer_step)}
        # BASELINE config 1 (the reference's own benchmark size) and a Linear-layer shape, kernel time under a CUDA graph
        roof_also["gemm_f32_1024"] = bench_gemm_graph(wk, ctx, pipe, np.float32, 1024, 1024, 1024, peaks)
        roof_also["gemm_f64_1024"] = bench_gemm_graph(wk, ctx, pipe, np.float64, 1024, 1024, 1024, peaks)
        roof_also["gemm_f32_256x4096x4096"] = bench_gemm_graph(wk, ctx, pipe, np.float32, 256, 4096, 4096, peaks)
        # BASELINE config 4: transposed operands and alpha/beta at N = 8192 (all four op pairs x (0.75, 0.5))
        sweep = {}
        for name_, dt_, st_ in (("f32", np.float32, 30), ("f64", np.float64, 4)):
            a_, b_, c_ = (wk.Tensor.alloc(ctx, pipe, (8192, 8192), dt_) for _ in range(3))
            wk.tensor.random.uniform(pipe, a_, 42, -1, 1)
            wk.tensor.random.uniform(pipe, b_, 43, -1, 1)
            # bring the chip to its sustained (power-capped) clock first: the four variants are compared with each other
            timed(wk, pipe, lambda: wk.blas.gemm(pipe, None, a_, 0, b_, 0, None, c_), 60 if name_ == "f32" else 8, 1)
            for opn, oa, ob in (("NN", 0, 0), ("NT", 0, 1), ("TN", 1, 0), ("TT", 1, 1)):
                wk.tensor.random.uniform(pipe, c_, 44, -1, 1)
                ms_, _, _, _ = timed(wk, pipe, lambda: wk.blas.gemm(pipe, 0.75, a_, oa, b_, ob, 0.5, c_), st_, 3)
                tf_ = 2.0 * 8192 ** 3 * st_ / (ms_ * 1e-3) / 1e12
                sweep[f"{name_}_{opn}"] = round(tf_, 2)
            for t_ in (a_, b_, c_):
                t_.release(pipe)
        pk32, pk64 = TF32_MMA_PEAK_TFLOPS["sustained"] / 3.0, FP64_DMMA_PEAK_TFLOPS
        roof_also["gemm_n8192_ops_alpha_beta"] = {
            "bound": "tensor", "unit": "TFLOP/s", "tflops": sweep,
            "frac_min_f32": min(v for k_, v in sweep.items() if k_.startswith("f32")) / pk32,
            "frac_min_f64": min(v for k_, v in sweep.items() if k_.startswith("f64")) / pk64,
            "workload": "C = 0.75 op(A) op(B) + 0.5 C, N = 8192, four transpose pairs (BASELINE config 4)"}
        # integer GEMM (8 of the 10 real dtypes): all on the tensor cores (kind::i8 over byte planes)
        for key, dt_, n_ in (("gemm_i8_n8192", np.int8, 8192), ("gemm_i16_n8192", np.int16, 8192), ("gemm_i32_n8192", np.int32, 8192),
                             ("gemm_i64_n4096", np.int64, 4096)):
            roof_also[key] = bench_gemm_int(wk, ctx, pipe, dt_, n_, 5, 3)
    if world > 1 and not args.quick:  # "AXPY HBM GB/s at 1/2/4/8 B200": every rank streams its own vectors
        for key, dt_ in (("axpy_f32_2p28", np.float32), ("axpy_f64_2p28", np.float64)):
            r_ = bench_axpy(wk, ctx, pipe, dt_, 1 << 28, 50, 5, peaks, barrier, max_over_ranks, world)
            also.append(r_)
            keep(key, r_)
    if world == 1 and rank == 0 and not args.no_cpu:
        r = cpu_reference_gemm(args.ref_n, np.float32, 3, 1)
        cpu = {"value": r["tflops"], "unit": "TFLOP/s", "cores": r["cores"], "kind": "port",
               "sample": (f"f32 NN GEMM N={args.ref_n} sub-problem, restated gemm_pack.cl + gemm_nxn_pack.cl "
                          f"({r['tile']}x{r['tile']} tiles, vw {r['vw']}), OpenMP, {r['sec_per_call']:.2f} s/call"),
               "axpy_f32_gbs": cpu_reference_axpy(1 << 26, np.float32)}

    if rank == 0 and e2e is not None:
        hl = host_link_ceiling(g, e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"])
        if hl is not None:
            hl["step_ms_over_copies_alone"] = e2e["ms_per_step"] / min(hl["copies_alone_ms_serial"], hl["copies_alone_ms_concurrent"])
            e2e["host_link"] = hl
    if rank == 0:
        line = {
            "metric": "gemm_f32_tflops", "value": tflops, "unit": "TFLOP/s", "n_gpus": g, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(N, g, gather),
            "arith": "3xTF32 tcgen05, fp32 accumulate", "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {**gemm_roofline(per_gpu_tflops, "f32", peaks, sustained=True,
                                         traffic=ncu_traffic_bytes("gemm_f32_n32768_1gpu") if (g == 1 and N == 32768) else None),
                         "also": roof_also},
            "cpu_baseline": cpu, "also": also,
        }
        if layer_step is not None:
            line["layer_step_streaming"] = {"n": 1 << 27, "peak_gbs": peaks["hbm_gbs"], "peak_basis": f"{peaks['_source']} copy bandwidth",
                                            "kernels": layer_step}
        if world > 1:
            line["gather_variants_tflops"] = variants
            line["gather_check"] = gather_check
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--ref-n", type=int, default=8192, dest="ref_n")
    ap.add_argument("--gather", default="fused", choices=["none", "nccl", "fused"])
    ap.add_argument("--quick", action="store_true", help="headline line only (no secondary configs)")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
