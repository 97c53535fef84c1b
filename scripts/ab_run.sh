mkdir -p gpurun_out
python -m pytest tests/test_transforms_gpu.py -q -s -k "rq_spline" 2>&1 | grep -E "passed|failed|^FAILED|spline n|inverse:"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-extras > gpurun_out/ab5_bench.json 2>gpurun_out/ab5_bench.err; python -c "import json; d=json.load(open('gpurun_out/ab5_bench.json')); print('bench', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['precision_check'], d['kernel_classes'])"
