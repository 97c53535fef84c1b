mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "inner_loop or full_depth or deterministic" 2>&1 | tail -2
for v in main nostaged main nostaged; do
  if [ $v = main ]; then unset FLOWCOMPARE_B200_LIB; else export FLOWCOMPARE_B200_LIB=flowcompare_b200/variants/lib_$v.so; fi
  python bench.py --no-extras --no-cpu-baseline --steps 3 > gpurun_out/cpl_$v.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/cpl_$v.json')); print('$v', d['value'], d['kernel_classes']['gemm_tcgen05_3x']['ms'])"
done
