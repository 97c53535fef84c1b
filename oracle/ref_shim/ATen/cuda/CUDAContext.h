// TEST INFRASTRUCTURE ONLY: see oracle/ref_shim/torch/serialize/tensor.h
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
using std::max;
using std::min;
