# Round-2 evidence run (one gpurun call).  Every ncu command runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --ncu --steps 1 --batch 16 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
for shape in "65536 512 512 1" "65536 256 256 1" "65536 512 300 0"; do
  tag=$(echo $shape | tr ' ' '_')
  python scripts/gemm_one.py $shape > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_gemm_tc_r02_$tag python scripts/gemm_one.py $shape > gpurun_out/r02_ncu_g_$tag.log 2>&1
done
python scripts/attn_bench.py 64 1024 1250 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:attention_tc_kernel -s 12 -c 1 -f -o gpurun_out/prof_attn_tc_r02 python scripts/attn_bench.py 64 1024 1250 > gpurun_out/r02_ncu_a.log 2>&1
ncu --set full --clock-control none -k regex:"knn_dist_kernel|knn_select_kernel|norms_kernel|edgeconv_gather_max_kernel|ln_stats_kernel" -c 16 -f -o gpurun_out/prof_small_r02 $CMD > gpurun_out/r02_ncu_s.log 2>&1
ls -la gpurun_out/ | grep -E "prof_|launches_r02"
