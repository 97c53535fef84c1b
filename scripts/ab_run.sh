mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or deterministic" 2>&1 | tail -2
python scripts/determinism_check.py 2>&1 | tail -3
python scripts/gemm_time.py 2>&1 | tail -1
for v in nofence single mmaonly koepi koconv noload; do FLOWCOMPARE_B200_LIB=flowcompare_b200/variants/lib_$v.so python scripts/gemm_time.py 2>&1 | tail -1; done
FLOWCOMPARE_B200_LIB=flowcompare_b200/variants/lib_timers.so python scripts/tc_phases.py 2>&1 | tail -6 | head -3
