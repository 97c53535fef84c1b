mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --no-extras > gpurun_out/ab.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/ab.json')); print('bench', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks']['sm_mhz'])"
