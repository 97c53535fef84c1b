mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or deterministic or inner_loop_matches" 2>&1 | tail -2
echo "tma store on"; python scripts/gemm_time.py 2>&1 | tail -1
echo "tma store off"; FC_TC_TMASTORE=0 python scripts/gemm_time.py 2>&1 | tail -1
python bench.py --no-extras --no-cpu-baseline --steps 3 > gpurun_out/tma_on.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/tma_on.json')); print('on ', d['value'], d['kernel_classes']['gemm_tcgen05_3x'])"
FC_TC_TMASTORE=0 python bench.py --no-extras --no-cpu-baseline --steps 3 > gpurun_out/tma_off.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/tma_off.json')); print('off', d['value'], d['kernel_classes']['gemm_tcgen05_3x'])"
