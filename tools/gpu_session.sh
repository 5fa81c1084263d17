set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int_tc.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_int_tc.log
cat gpurun_out/pytest_int_tc.log | tail -5
timeout 600 python tools/gemm_int_sweep.py gpurun_out/sweep_gemm_int_r02b > gpurun_out/sweep_int_b.log 2>&1
tail -3 gpurun_out/sweep_int_b.log
# ncu: HEAD f32 kernel at N=32768 (plain launch for the tail split: ncu cannot replay cooperative cluster launches)
WK_GEMM_SPLITK_COOP=0 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -o gpurun_out/ncu_gemm_f32_r02_n32768 python tools/gemm_time.py f32 32768 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
