// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// The reference's own ROIAlignRotated3D CUDA kernels, compiled UNMODIFIED where they lie
// (/root/reference/maskrcnn_benchmark/csrc/cuda/ROIAlignRotated3D_cuda.cu, #included below through the -I path of oracle/Makefile,
// <THC/*> resolved to oracle/shim/THC), behind a C ABI on raw device pointers so that the GPU parity tests can run them beside the
// product kernel on the same inputs.  Output: oracle/_ref/libroialign3d_ref.so (sm_100a; git-ignored, travels to the GPU box).
#include <ATen/ATen.h>
#include <torch/types.h>

// The reference dispatches with AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...): torch >= 2.1 only accepts a ScalarType there.  The
// macro is redefined for THIS translation unit to the float instantiation (the only one the glue below uses); the reference file
// itself stays untouched.
#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  {                                                 \
    using scalar_t = float;                         \
    (__VA_ARGS__)();                                \
  }
#include "cuda/ROIAlignRotated3D_cuda.cu"

extern "C" {
// input [batch][C][H][W][Z], rois [n][8] (batch, cw, ch, cz, w, h, z, theta in degrees) -> out [n][C][ph][pw][pz]
int ref_roi_align_rotated_3d_forward(const float *input, long batch, long C, long H, long W, long Z, const float *rois, long n, float scale, int ph, int pw, int pz,
                                     int sampling, float *out) {
  try {
    auto opt = at::TensorOptions().dtype(at::kFloat).device(at::kCUDA);
    auto in = at::from_blob(const_cast<float *>(input), {batch, C, H, W, Z}, opt);
    auto r = at::from_blob(const_cast<float *>(rois), {n, 8}, opt);
    auto o = ROIAlignRotated3D_forward_cuda(in, r, scale, ph, pw, pz, sampling);
    cudaMemcpy(out, o.data_ptr<float>(), sizeof(float) * o.numel(), cudaMemcpyDeviceToDevice);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
  } catch (...) { return 2; }
}
// grad [n][C][ph][pw][pz] -> d_input [batch][C][H][W][Z]
int ref_roi_align_rotated_3d_backward(const float *grad, const float *rois, long n, float scale, int ph, int pw, int pz, long batch, long C, long H, long W, long Z,
                                      int sampling, float *d_input) {
  try {
    auto opt = at::TensorOptions().dtype(at::kFloat).device(at::kCUDA);
    auto g = at::from_blob(const_cast<float *>(grad), {n, C, ph, pw, pz}, opt);
    auto r = at::from_blob(const_cast<float *>(rois), {n, 8}, opt);
    auto d = ROIAlignRotated3D_backward_cuda(g, r, scale, ph, pw, pz, (int)batch, (int)C, (int)H, (int)W, (int)Z, sampling);
    cudaMemcpy(d_input, d.data_ptr<float>(), sizeof(float) * d.numel(), cudaMemcpyDeviceToDevice);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
  } catch (...) { return 2; }
}
}
