#!/bin/bash
# TEST INFRASTRUCTURE ONLY -- compiles the REFERENCE's own pointops CUDA kernels, from the sources where they lie under
# /root/reference (nothing is copied), into oracle/_ref/libpointops_ref.so (git-ignored; travels to the GPU box).
#   K1 knnquery_heap   lib/pointops/src/knnquery_heap/knnquery_heap_cuda_kernel.cu
#   K2/K3 FPS, gather  lib/pointops/src/sampling/sampling_cuda_kernel.cu
#   K5/K6 3-NN, interp lib/pointops/src/interpolation/interpolation_cuda_kernel.cu
#   K4 grouping        lib/pointops/src/grouping/grouping_cuda_kernel.cu
# Only the *_cuda_kernel.cu files (kernels + extern "C" raw-pointer launchers) are built; the *_cuda.cpp at::Tensor wrappers
# need <THC/THC.h> (gone from modern torch) and are not.  The kernel headers include two torch headers only to declare those
# wrappers; by default they resolve to the two-line stand-ins in oracle/ref_shim (seconds instead of minutes per file);
# FC_REF_REAL_TORCH_HEADERS=1 uses the installed torch headers instead (same object code, ~4 min).
# The reference has no sm_100 target: the same sources are simply compiled for sm_100a here.
set -e
REF=${FLOWCOMPARE_REFERENCE:-/root/reference}
SRC=$REF/models/scene_seg_PAConv/lib/pointops/src
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref/libpointops_ref.so
[ -d "$SRC" ] || { echo "reference tree not found at $REF: keeping any prebuilt $OUT"; exit 0; }
mkdir -p "$HERE/_ref"
FILES="$SRC/knnquery_heap/knnquery_heap_cuda_kernel.cu $SRC/sampling/sampling_cuda_kernel.cu $SRC/interpolation/interpolation_cuda_kernel.cu $SRC/grouping/grouping_cuda_kernel.cu"
if [ -f "$OUT" ] && [ -z "$(find $FILES "$0" -newer "$OUT" 2>/dev/null)" ]; then echo "$OUT is up to date"; exit 0; fi
if [ "$FC_REF_REAL_TORCH_HEADERS" = "1" ]; then
  INC=$(python -c "import torch.utils.cpp_extension as c; print(' '.join('-I' + p for p in c.include_paths()))")
else
  INC="-I$HERE/ref_shim"
fi
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Wno-deprecated-gpu-targets -w -shared -Xcompiler -fPIC $INC -o "$OUT" $FILES
echo "built $OUT"
