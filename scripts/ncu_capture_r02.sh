# Round-2 evidence run (one gpurun call): GPU tests, the default bench, then the ncu launch list and full captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
CMD="python bench.py --ncu --steps 1 --batch 16 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
python scripts/gemm_one.py 65536 512 512 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_gemm_tc_r02 python scripts/gemm_one.py 65536 512 512 1 > gpurun_out/r02_ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|knn_norms|edgeconv_gather_max_kernel|ln_stats_kernel" -c 10 -f -o gpurun_out/prof_small_r02 $CMD > gpurun_out/r02_ncu_s.log 2>&1
ls -la gpurun_out/
