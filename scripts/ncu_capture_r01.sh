CMD="python bench.py --ncu --steps 1 --batch 16"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/launches_r01d.csv $CMD > gpurun_out/ncu_list2.log 2>&1
python scripts/gemm_one.py 65536 512 512 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_gemm_tc_r01d python scripts/gemm_one.py 65536 512 512 1 > gpurun_out/ncu_g2.log 2>&1
python scripts/attn_bench.py 64 1024 1250 > gpurun_out/attn_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_attn_tc_r01d python scripts/attn_bench.py 64 1024 1250 > gpurun_out/ncu_a2.log 2>&1
ncu --set full --clock-control none -k regex:"knn_kernel|edgeconv_gather_max_kernel|ln_stats_kernel" -c 8 -f -o gpurun_out/prof_small_r01d $CMD > gpurun_out/ncu_s2.log 2>&1
ls -la gpurun_out/*r01d*
