// ORACLE / TEST INFRASTRUCTURE ONLY (see THC.h in this directory).
#pragma once
template <typename T> __host__ __device__ __forceinline__ T THCCeilDiv(T a, T b) { return (a + b - 1) / b; }
