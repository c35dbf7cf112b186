// RPN head and anchor generation on the sparse feature maps (the first consumers of the backbone's outputs).
//   reference: maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:81-131 (RPNHead: three 1x1 Conv2d on [1, C, n, 1] = GEMMs over the n feature
//   rows), maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:88-104 (grid_anchors).
// Sizes at the B470 building: 2,312 feature rows x 128 channels over the four rpn maps, 4 yaws x 2 class groups -> 8 logits + 56 box
// deltas per row, 9,248 anchors: 57 MFLOP and 1.3 MB of rows -- a launch-latency problem, not a throughput one.  Hence ONE kernel for
// the whole head: conv + ReLU + both output layers per tile of 8 rows, hidden activations never leave shared memory, exact fp32 on
// the CUDA cores (the reference runs these layers in fp32 and the proposals are ranked by the logits: no operand rounding here).
#include "../../include/scn_b200.h"
#include "common.cuh"

namespace scn {
namespace {
constexpr int kHeadRows = 8;
// w_*: Conv2d weights [out][in] (kernel 1x1).  logits [n][nCls], reg [n][nReg].
__global__ void __launch_bounds__(128) k_rpn_head(const float *__restrict__ x, long n, int C, const float *__restrict__ wc, const float *__restrict__ bc,
                                                  const float *__restrict__ wl, const float *__restrict__ bl, int nCls, const float *__restrict__ wr,
                                                  const float *__restrict__ br, int nReg, float *__restrict__ logits, float *__restrict__ reg) {
  extern __shared__ float sm[]; // xs[kHeadRows][C], ts[kHeadRows][C]
  float *xs = sm, *ts = sm + kHeadRows * C;
  const long r0 = (long)blockIdx.x * kHeadRows;
  const int rows = (int)min((long)kHeadRows, n - r0);
  for (int i = threadIdx.x; i < kHeadRows * C; i += blockDim.x) xs[i] = i < rows * C ? __ldg(x + r0 * C + i) : 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { // hidden channel c of every row of the tile
    float acc[kHeadRows];
    const float b = bc ? __ldg(bc + c) : 0.f;
#pragma unroll
    for (int r = 0; r < kHeadRows; r++) acc[r] = b;
    const float *w = wc + (long)c * C;
    for (int k = 0; k < C; k += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w + k));
#pragma unroll
      for (int r = 0; r < kHeadRows; r++) {
        const float4 v = *reinterpret_cast<const float4 *>(xs + r * C + k);
        acc[r] = fmaf(w4.x, v.x, acc[r]); acc[r] = fmaf(w4.y, v.y, acc[r]); acc[r] = fmaf(w4.z, v.z, acc[r]); acc[r] = fmaf(w4.w, v.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; r++) ts[r * C + c] = fmaxf(acc[r], 0.f); // F.relu
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nCls + nReg; o += blockDim.x) {
    const bool isCls = o < nCls;
    const int oo = isCls ? o : o - nCls;
    const float *w = (isCls ? wl : wr) + (long)oo * C;
    const float *bp = isCls ? bl : br;
    float acc[kHeadRows];
    const float b = bp ? __ldg(bp + oo) : 0.f;
#pragma unroll
    for (int r = 0; r < kHeadRows; r++) acc[r] = b;
    for (int k = 0; k < C; k += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w + k));
#pragma unroll
      for (int r = 0; r < kHeadRows; r++) {
        const float4 v = *reinterpret_cast<const float4 *>(ts + r * C + k);
        acc[r] = fmaf(w4.x, v.x, acc[r]); acc[r] = fmaf(w4.y, v.y, acc[r]); acc[r] = fmaf(w4.z, v.z, acc[r]); acc[r] = fmaf(w4.w, v.w, acc[r]);
      }
    }
    for (int r = 0; r < rows; r++) {
      if (isCls) logits[(r0 + r) * nCls + oo] = acc[r];
      else reg[(r0 + r) * nReg + oo] = acc[r];
    }
  }
}
// anchors[(row * A + a)][7] = [loc_xyz / voxel_scale * stride, 0, 0, 0, 0] + base[a]   (yx_zb boxes: xc, yc, z_bot, y_size, x_size, z_size, yaw)
// The reference evaluates (loc.float() / voxel_scale) * stride and then adds: three separately rounded fp32 operations.
__global__ void k_grid_anchors(const long *__restrict__ loc, long n, float voxelScale, float s0, float s1, float s2, const float *__restrict__ base, int A,
                               float *__restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n * A; i += (long)gridDim.x * blockDim.x) {
    const long row = i / A;
    const int a = (int)(i - row * A);
    const float c[3] = {__fmul_rn(__fdiv_rn((float)loc[row * 4 + 0], voxelScale), s0), __fmul_rn(__fdiv_rn((float)loc[row * 4 + 1], voxelScale), s1),
                        __fmul_rn(__fdiv_rn((float)loc[row * 4 + 2], voxelScale), s2)};
    float *o = out + i * 7;
    const float *b = base + a * 7;
#pragma unroll
    for (int j = 0; j < 7; j++) o[j] = __fadd_rn(j < 3 ? c[j] : 0.f, __ldg(b + j));
  }
}
} // namespace
} // namespace scn

extern "C" {
int scn_rpn_head_forward(const float *feats, long n_rows, int n_planes, const float *w_conv, const float *b_conv, const float *w_cls, const float *b_cls,
                         int n_cls, const float *w_reg, const float *b_reg, int n_reg, float *logits, float *reg, void *stream) {
  SCN_CHECK(n_rows >= 0 && n_planes > 0 && n_planes % 4 == 0 && n_planes <= 1024 && n_cls >= 0 && n_reg >= 0, "RPN head shapes");
  if (n_rows == 0) return 0;
  SCN_CHECK(feats && w_conv && w_cls && w_reg && logits && reg, "RPN head: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = 2 * scn::kHeadRows * (size_t)n_planes * sizeof(float);
  if (smem > 48 * 1024) SCN_CUDA(cudaFuncSetAttribute(scn::k_rpn_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scn::k_rpn_head<<<scn::cdiv(n_rows, scn::kHeadRows), 128, smem, scn::LS(s)>>>(feats, n_rows, n_planes, w_conv, b_conv, w_cls, b_cls, n_cls, w_reg, b_reg, n_reg,
                                                                                logits, reg);
  SCN_CUDA(cudaGetLastError());
  return 0;
}
int scn_rpn_grid_anchors(const long *locations, long n_rows, float voxel_scale, const float stride[3], const float *base_anchors, int n_anchors, float *anchors,
                         void *stream) {
  SCN_CHECK(n_rows >= 0 && n_anchors > 0 && voxel_scale > 0.f, "anchor arguments");
  if (n_rows == 0) return 0;
  SCN_CHECK(locations && base_anchors && anchors, "anchors: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  scn::k_grid_anchors<<<scn::stream_grid(n_rows * n_anchors, 256), 256, 0, scn::LS(s)>>>(locations, n_rows, voxel_scale, stride[0], stride[1], stride[2], base_anchors,
                                                                                         n_anchors, anchors);
  SCN_CUDA(cudaGetLastError());
  return 0;
}
}
