mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; python -c "import json; d=json.load(open('gpurun_out/r02_bench_final.json')); print('bench', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks']['sm_mhz'], {k:(v['ms'],v['tflops']) for k,v in d['kernel_classes'].items()}); print(d['sweep_one_gpu']); print(d['per_config'])"
