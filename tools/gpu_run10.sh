#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_splitk.py tests/test_gpu_parity.py tests/test_gpu_linear.py tests/test_gpu_xor.py tests/test_gpu_complex.py -m gpu -q -x --tb=short 2>&1 | tail -30 > gpurun_out/pytest_splitk.log; tail -6 gpurun_out/pytest_splitk.log
for mode in auto off; do
  for dt in f32 f64; do
    if [ $mode = off ]; then export WK_GEMM_SPLITK=1; else unset WK_GEMM_SPLITK; fi
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_${dt}_${mode}.csv -k regex:'gemm_tf32x3|gemm_dmma' \
       python tools/gemm_time.py $dt 256 512 1024 1536 2048 > /dev/null 2>&1
    echo "== $dt splitk=$mode (kernel us, last launch of each size)"
    python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/small_${dt}_${mode}.csv")) if len(r)>5 and r[0].isdigit()]
# 13 launches per size (3 warm + 10 timed)
vals=[float(r[-1]) for r in rows]
per=13
for i,n in enumerate((256,512,1024,1536,2048)):
    v=vals[i*per:(i+1)*per]
    if v: print(n, "min %.1f us  median %.1f us" % (min(v)/1000 if max(v)>1000 else min(v), sorted(v)[len(v)//2]/1000 if max(v)>1000 else sorted(v)[len(v)//2]), rows[i*per][-2])
PY
  done
done
unset WK_GEMM_SPLITK
timeout 200 python tools/linear_time.py 256 4096 4096 2>&1 | tail -3
