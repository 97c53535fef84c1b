mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_transforms_gpu.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
for v in flowcompare_b200/libflowcompare_b200.so flowcompare_b200/libfc_e0.so flowcompare_b200/libfc_base.so; do
FLOWCOMPARE_B200_LIB=$PWD/$v python bench.py --no-extras > gpurun_out/ab.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/ab.json')); print('$v', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'], {k:(v['ms'],v['tflops']) for k,v in d['kernel_classes'].items()})"
done; done
