cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_r02j.log; cat gpurun_out/pytest_r02j.log
python tools/stream_sweep.py gpurun_out/sweep_stream_r02 27 > gpurun_out/sweep_stream_r02.log 2>&1; grep -c . gpurun_out/sweep_stream_r02.md
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err; tail -c 300 gpurun_out/bench_r02j.json; tail -3 gpurun_out/bench_r02j.err
