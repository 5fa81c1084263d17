cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/pcie_peak.py > gpurun_out/pcie_peak_g1.json 2>/dev/null; cat gpurun_out/pcie_peak_g1.json
cp gpurun_out/pcie_peak_g1.json profiles/pcie_peak_r02_g1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r02e.log; cat gpurun_out/pytest_r02e.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02e.json 2>gpurun_out/bench_ref_r02e.err; tail -c 400 gpurun_out/bench_ref_r02e.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; tail -c 300 gpurun_out/bench_r02e.json; tail -3 gpurun_out/bench_r02e.err
WK_GEMM_TAILSPLIT=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02e.csv python bench.py --steps 3 --warmup 3 --quick --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/launches_r02e.csv
