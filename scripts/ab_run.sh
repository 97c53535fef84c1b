mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
FC_PROFILE_DUMP=$PWD/gpurun_out/final_dump.txt timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; python -c "import json; d=json.load(open('gpurun_out/r02_bench_final.json')); print('bench', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
python scripts/shape_dump.py gpurun_out/final_dump.txt > gpurun_out/final_shapes.txt; head -3 gpurun_out/final_shapes.txt
CMD="python bench.py --ncu --steps 1 --batch 16 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
for shape in "65536 512 512 1" "65536 256 256 1" "65536 512 300 0"; do
  tag=$(echo $shape | tr ' ' '_')
  python scripts/gemm_one.py $shape > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_gemm_tc_r02_$tag python scripts/gemm_one.py $shape > gpurun_out/r02_ncu_g_$tag.log 2>&1
done
ls -la gpurun_out/ | grep -E "prof_gemm|launches_r02"
