# One gpurun call = one run of this script on a fresh B200 box (see tools/README.md); this is the round's validation session.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>/dev/null; tail -c 250 gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
WK_GEMM_TAILSPLIT=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches.csv | cut -c1-200
