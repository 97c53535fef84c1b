// TEST INFRASTRUCTURE ONLY (oracle/build_ref_pointops.sh): stands in for <torch/serialize/tensor.h> when the reference's
// pointops *_cuda_kernel.cu files are compiled into oracle/_ref.  Those kernel files use torch only to DECLARE the
// at::Tensor wrappers that live in the *_cuda.cpp files (which are not compiled: they include <THC/THC.h>, removed from
// torch >= 1.11); the kernels and their extern "C" raw-pointer launchers are torch-free.
#pragma once
namespace at { class Tensor; }
