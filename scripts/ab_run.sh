mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or deterministic" > gpurun_out/ab1_pytest.log 2>&1; tail -3 gpurun_out/ab1_pytest.log
echo "== warp arrive"; python scripts/gemm_bench.py 65536 2>&1 | tee gpurun_out/ab1_gemm_warp.log | cut -d'|' -f3
echo "== thread arrive"; FLOWCOMPARE_B200_LIB=flowcompare_b200/variants/lib_thrarrive.so python scripts/gemm_bench.py 65536 2>&1 | tee gpurun_out/ab1_gemm_thr.log | cut -d'|' -f3
python bench.py --no-extras --no-cpu-baseline > gpurun_out/ab1_bench_warp.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/ab1_bench_warp.json')); print('warp', d['value'], d['roofline']['achieved'])"
FLOWCOMPARE_B200_LIB=flowcompare_b200/variants/lib_thrarrive.so python bench.py --no-extras --no-cpu-baseline > gpurun_out/ab1_bench_thr.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/ab1_bench_thr.json')); print('thr', d['value'], d['roofline']['achieved'])"
