# Same-box A/B of two builds of the library (FLOWCOMPARE_B200_LIB picks the .so): A = TC_FUSE_WLO=0, B = TC_FUSE_WLO=1 (default build).
# Build A first:  scripts/build_variant.sh fuse0 gemm_tc "-DTC_FUSE_WLO=0"   (output of the last run: profiles/r02_fuse_wlo_ab.txt)
mkdir -p gpurun_out
A=$PWD/flowcompare_b200/variants/lib_fuse0.so; B=$PWD/flowcompare_b200/libflowcompare_b200.so
i=0
for L in $A $B $B $A $A $B; do
  i=$((i+1))
  FLOWCOMPARE_B200_LIB=$L timeout 40 python bench.py --no-extras --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/abx_$i.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/abx_$i.json')); print('$i', '$(basename $L)', d['value'], d['ms_per_step'], d['kernel_classes']['gemm_tcgen05_3x']['ms'], d['kernel_classes']['cross_attention']['ms'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/abx.log
done
