cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
run() {
  echo "== gpus=$G $*"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps 5 --warmup 3 --quick --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d.get('gather_variants_tflops'), d.get('gather_check', {}).get('ok'), d['clocks']['sm_mhz'])
"
}
run WK_GEMM_PEER_BULK=0
run WK_GEMM_PEER_BULK=1
