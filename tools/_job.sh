python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])"
