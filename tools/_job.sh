# default GPU-box job: the driver's own test command, smoke, one bench line
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | cut -c1-600
