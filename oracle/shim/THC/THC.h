// ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for the THC headers that torch >= 1.11 no longer ships: the reference's
// maskrcnn_benchmark/csrc/cuda/*.cu include <THC/THC.h>, <THC/THCAtomics.cuh> and <THC/THCDeviceUtils.cuh> for exactly three
// things -- THCudaCheck, atomicAdd on floats (a CUDA builtin) and THCCeilDiv -- which are restated here so that the reference
// kernel file compiles UNMODIFIED (oracle/Makefile, target ref_gpu).
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#define THCudaCheck(expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t thc_shim_err = (expr);                                                                      \
    if (thc_shim_err != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(thc_shim_err)); \
  } while (0)
