cd $GRAFT_REPO_ROOT
for cfg in "2 16" "4 16" "4 8" "8 8" "8 4"; do set -- $cfg; echo "NJ=$1 PANELS=$2"; WK_E2E_NJ=$1 WK_E2E_PANELS=$2 python bench.py --steps 5 --warmup 3 --quick --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['e2e']['ms_per_step'],1))"; done
