mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_transforms_gpu.py tests/test_sampling_gpu.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
for v in flowcompare_b200/libflowcompare_b200.so flowcompare_b200/libfc_base.so; do
FC_PROFILE_DUMP=$PWD/gpurun_out/dump_$(basename $v .so).txt FLOWCOMPARE_B200_LIB=$PWD/$v timeout 300 python bench.py --no-extras > gpurun_out/ab.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/ab.json')); print('$v', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'], {k:(v['ms'],v['tflops']) for k,v in d['kernel_classes'].items()})"
done; done
python scripts/shape_dump.py gpurun_out/dump_libflowcompare_b200.txt
python scripts/shape_dump.py gpurun_out/dump_libfc_base.txt
