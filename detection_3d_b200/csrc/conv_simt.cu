// fp32 CUDA-core gather-GEMM kernels: the exact-fp32 path (tolerance anchor for the tensor-core
// path, and the only path for channel counts the tcgen05 kernel does not take, e.g. Cin = 9).
//
//  * conv_plan_simt : output-stationary.  A CTA owns 64 output sites (in spatial order) x 64 output
//    channels and loops over the filter offsets that have at least one neighbour inside the tile,
//    gathering input rows by the plan's neighbour table.  One store per output element, no
//    read-modify-write, no atomics (replaces dConvolution_KMxKN_forwardA/B + RULEBOOKITERATOR,
//    SCN/CUDA/Convolution.cu:57-203, RuleBookIterator.h:15-32).
//  * conv_list_simt : rule-list driven (gather rows by pairs[i].src, write rows pairs[i].dst) for
//    deconvolution, where every fine row has one parent (SCN/CUDA/Deconvolution.cu).
#include "common.cuh"

namespace scn {

constexpr int TM = 64, TN = 64, KC = 16, APAD = 4;

template <bool VEC>
__device__ __forceinline__ void load_a_chunk(const float *__restrict__ in, int row, int Cin, int c0, int tid, float (*A_s)[TM + APAD]) {
  const int rr = tid >> 2, c4 = (tid & 3) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (row >= 0) {
    const float *src = in + (long)row * Cin + c0 + c4;
    if (VEC) {
      if (c0 + c4 < Cin) { float4 t = __ldg(reinterpret_cast<const float4 *>(src)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) if (c0 + c4 + j < Cin) v[j] = __ldg(src + j);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) A_s[c4 + j][rr] = v[j];
}
template <bool VEC>
__device__ __forceinline__ void load_b_chunk(const float *__restrict__ Wk, int Cin, int Cout, int c0, int n0, int tid, float (*B_s)[TN]) {
  const int kk = tid >> 4, n4 = (tid & 15) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 + kk < Cin) {
    const float *src = Wk + (long)(c0 + kk) * Cout + n0 + n4;
    if (VEC) {
      if (n0 + n4 < Cout) { float4 t = __ldg(reinterpret_cast<const float4 *>(src)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) if (n0 + n4 + j < Cout) v[j] = __ldg(src + j);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) B_s[kk][n4 + j] = v[j];
}
__device__ __forceinline__ void mma_chunk(const float (*A_s)[TM + APAD], const float (*B_s)[TN], int tx, int ty, float acc[4][4]) {
#pragma unroll
  for (int kk = 0; kk < KC; kk++) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = A_s[kk][ty * 4 + i];
    float4 bv = *reinterpret_cast<const float4 *>(&B_s[kk][tx * 4]);
    b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) conv_plan_simt(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ W,
                                                      const int *__restrict__ nbr, const int *__restrict__ outRow, int nOut, int K, int Cin, int Cout,
                                                      const float *__restrict__ bias) {
  __shared__ __align__(16) float A_s[KC][TM + APAD];
  __shared__ __align__(16) float B_s[KC][TN];
  __shared__ int s_ids[TM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int k = 0; k < K; k++) {
    int id = -1;
    if (tid < TM && p0 + tid < nOut) id = nbr ? __ldg(nbr + nbr_index(p0 + tid, k, K)) : p0 + tid; // no plan: dense rows (NetworkInNetwork)
    __syncthreads(); // previous offset's readers of s_ids / tiles are done
    if (tid < TM) s_ids[tid] = id;
    if (!__syncthreads_or(id >= 0)) continue;
    const int myrow = s_ids[tid >> 2];
    const float *Wk = W + (long)k * Cin * Cout;
    for (int c0 = 0; c0 < Cin; c0 += KC) {
      load_a_chunk<VEC>(in, myrow, Cin, c0, tid, A_s);
      load_b_chunk<VEC>(Wk, Cin, Cout, c0, n0, tid, B_s);
      __syncthreads();
      mma_chunk(A_s, B_s, tx, ty, acc);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int p = p0 + ty * 4 + i;
    if (p >= nOut) continue;
    float *dst = out + (long)(outRow ? __ldg(outRow + p) : p) * Cout + n0 + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = n0 + tx * 4 + j;
      if (c < Cout) dst[j] = acc[i][j] + (bias ? __ldg(bias + c) : 0.f);
    }
  }
}

// Rule-list driven: tile = 64 consecutive pairs of one list.
template <bool VEC>
__global__ void __launch_bounds__(256) conv_list_simt(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ W,
                                                      const int2 *__restrict__ pairs, const int *__restrict__ off, int K, int Cin, int Cout,
                                                      int srcIsY, int accumulate, int onlyList) {
  __shared__ __align__(16) float A_s[KC][TM + APAD];
  __shared__ __align__(16) float B_s[KC][TN];
  __shared__ int s_src[TM], s_dst[TM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // locate (list, first pair) of this tile
  int t = blockIdx.x, L = 0, start = 0, len = 0;
  if (onlyList >= 0) { L = onlyList; start = off[L]; len = off[L + 1] - start; }
  else {
    for (L = 0; L < K; L++) {
      start = off[L]; len = off[L + 1] - start;
      int tiles = (len + TM - 1) / TM;
      if (t < tiles) break;
      t -= tiles;
    }
    if (L == K) return;
  }
  const int i0 = t * TM, n0 = blockIdx.y * TN;
  if (tid < TM) {
    int2 pr = make_int2(-1, -1);
    if (i0 + tid < len) pr = __ldg(pairs + start + i0 + tid);
    s_src[tid] = srcIsY ? pr.y : pr.x;
    s_dst[tid] = srcIsY ? pr.x : pr.y;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  const int myrow = s_src[tid >> 2];
  const float *Wk = W + (long)L * Cin * Cout;
  for (int c0 = 0; c0 < Cin; c0 += KC) {
    load_a_chunk<VEC>(in, myrow, Cin, c0, tid, A_s);
    load_b_chunk<VEC>(Wk, Cin, Cout, c0, n0, tid, B_s);
    __syncthreads();
    mma_chunk(A_s, B_s, tx, ty, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int d = s_dst[ty * 4 + i];
    if (d < 0) continue;
    float *dst = out + (long)d * Cout + n0 + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = n0 + tx * 4 + j;
      if (c < Cout) dst[j] = accumulate ? dst[j] + acc[i][j] : acc[i][j];
    }
  }
}

int launch_conv_plan_simt(const float *in, float *out, const float *W, const int *nbr, const int *outRow, int nOut, int K, int Cin, int Cout,
                          const float *bias, cudaStream_t s) {
  if (nOut == 0) return 0;
  ++g_counters[kCntSimtLaunch];
  dim3 grid(cdiv(nOut, TM), cdiv(Cout, TN));
  if (Cin % 4 == 0 && Cout % 4 == 0)
    conv_plan_simt<true><<<grid, 256, 0, LS(s)>>>(in, out, W, nbr, outRow, nOut, K, Cin, Cout, bias);
  else
    conv_plan_simt<false><<<grid, 256, 0, LS(s)>>>(in, out, W, nbr, outRow, nOut, K, Cin, Cout, bias);
  SCN_CUDA(cudaGetLastError());
  return 0;
}

// offHost: list offsets on the host (K+1).  singleParent: every dst row occurs in exactly one rule.
int launch_conv_list_simt(const float *in, float *out, const float *W, const int2 *pairs, const int *d_off, const int *offHost, int K, int Cin,
                          int Cout, int srcIsY, int singlePass, cudaStream_t s) {
  const bool vec = Cin % 4 == 0 && Cout % 4 == 0;
  ++g_counters[kCntSimtLaunch];
  if (singlePass) {
    long tiles = 0;
    for (int L = 0; L < K; L++) tiles += cdiv(offHost[L + 1] - offHost[L], TM);
    if (!tiles) return 0;
    dim3 grid((unsigned)tiles, cdiv(Cout, TN));
    if (vec) conv_list_simt<true><<<grid, 256, 0, LS(s)>>>(in, out, W, pairs, d_off, K, Cin, Cout, srcIsY, 0, -1);
    else conv_list_simt<false><<<grid, 256, 0, LS(s)>>>(in, out, W, pairs, d_off, K, Cin, Cout, srcIsY, 0, -1);
  } else { // a dst row may occur once per list: lists run back to back, read-modify-write
    for (int L = 0; L < K; L++) {
      int len = offHost[L + 1] - offHost[L];
      if (!len) continue;
      dim3 grid(cdiv(len, TM), cdiv(Cout, TN));
      if (vec) conv_list_simt<true><<<grid, 256, 0, LS(s)>>>(in, out, W, pairs, d_off, K, Cin, Cout, srcIsY, 1, L);
      else conv_list_simt<false><<<grid, 256, 0, LS(s)>>>(in, out, W, pairs, d_off, K, Cin, Cout, srcIsY, 1, L);
    }
  }
  SCN_CUDA(cudaGetLastError());
  return 0;
}

} // namespace scn
