for v in "2 1" "1 1" "2 0" "1 0"; do set -- $v
TDB_WGRAD_LAG=$1 TDB_MAIN_PRIO=$2 python bench.py --steps 20 --warmup 3 --skip-cpu --no-dedup-probe 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lag $1 mainprio $2: ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
