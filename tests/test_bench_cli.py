"""-m "not gpu": the driver-facing scripts at least parse and import here (no GPU): bench.py's CLI, the streaming sweep
module bench.py imports for `layer_step_streaming`, and the reference arm's JSON contract on a tiny sample."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_help_and_sweep_import():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "--impl" in r.stdout and "--gpus" in r.stdout
    spec = importlib.util.spec_from_file_location("wk_stream_sweep_t", os.path.join(ROOT, "tools", "stream_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert callable(mod.sweep) and callable(mod.main)


def test_reference_arm_contract_line():
    """bench.py --impl reference on a small sub-problem: one JSON line with the keys the driver reads"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-n", "256"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gemm_f32_tflops" and line["unit"] == "TFLOP/s"
    assert line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_env_uses_all_cores_and_our_config():
    """torch.distributed.run exports OMP_NUM_THREADS=1 (round 1: the arm then ran on one core and hit the driver's limit);
    rank 0 must undo that, size its sample to the time budget, and print the SAME config object as our arm; other ranks
    print nothing"""
    env = dict(os.environ, OMP_NUM_THREADS="1", WORLD_SIZE="2", RANK="0", LOCAL_RANK="0", WK_REF_BUDGET_S="4")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    ncpu = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == ncpu
    sys.path.insert(0, ROOT)
    import bench

    assert line["config"] == bench.workload_config(32768, 2, "fused")
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 2
    assert line["ms_per_step"] * 3 < 4000 * 2  # the sample was sized to the budget
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=60, cwd=ROOT, env=dict(env, RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_bench_reads_roofline_inputs_from_committed_profiles():
    """roofline.traffic comes from the committed ncu summary whose kernel NAME matches the HEAD kernel (not a constant), and the
    e2e host-link ceiling from the committed PCIe measurements"""
    sys.path.insert(0, ROOT)
    import bench

    t = bench.ncu_traffic_bytes("gemm_f32_n32768_1gpu")
    assert t is not None and 1.5e11 < t < 1.9e11  # 163 GB read + 4.6 GB written per launch at N = 32768
    for key in ("axpy_f32_2^28", "axpy_f64_2^28"):
        assert bench.ncu_traffic_bytes(key) is not None
    assert bench.ncu_traffic_bytes("no_such_kernel") is None
    hl = bench.host_link_ceiling(8, 2 * 32768 ** 2 * 4, 32768 ** 2 * 4)
    assert hl is not None and 60 < hl["copies_alone_ms_serial"] < 150 and hl["source"].endswith("pcie_peak_r02_g8.json")
    assert bench.host_link_ceiling(3, 1, 1) is None  # no measurement for 3 GPUs: no ceiling claimed
