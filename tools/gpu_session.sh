cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_r02h.log; cat gpurun_out/pytest_r02h.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02h.json 2>/dev/null; tail -c 250 gpurun_out/bench_ref_r02h.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; tail -c 200 gpurun_out/bench_r02h.json; tail -3 gpurun_out/bench_r02h.err
WK_GEMM_TAILSPLIT=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02h.csv python bench.py --steps 3 --warmup 3 --quick --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches_r02h.csv | cut -c1-200
timeout 600 python tools/gemm_int_sweep.py gpurun_out/sweep_gemm_int_r02h > gpurun_out/sweep_int_h.log 2>&1; head -20 gpurun_out/sweep_gemm_int_r02h.md
