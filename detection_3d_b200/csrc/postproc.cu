// The steps either side of the backbone path: RPN post-processing and the input voxeliser.
//   reference: maskrcnn_benchmark/modeling/rpn/inference_3d.py:82-163 (sigmoid, top-k, box decode, rotated 3-D NMS, post top-n),
//              maskrcnn_benchmark/modeling/box_coder_3d.py:38-65 + second/pytorch/core/box_torch_ops.py:51-88 (decode),
//              maskrcnn_benchmark/structures/boxlist_ops_3d.py:14-61 + second/pytorch/core/box_torch_ops.py:489-514 (boxlist_nms_3d, rotate_nms_3d),
//              utils3d/rotate_nms_3d_torch.py:7-84 (boxes_iou_3d = rotated BEV IoU x IoU along z),
//              second/core/non_max_suppression/nms_gpu.py:166-420,548-667 (rotated IoU of two rectangles, numba CUDA),
//              second/core/non_max_suppression/nms_cpu.py:32-44 (rotate_nms_3d_cc: greedy suppression, spconv's rotate_non_max_suppression_cpu),
//              data3d/suncg_utils/suncg_dataset.py:115-177 (voxeliser: affine map, offset, bounds mask, truncation to integer voxels).
// The reference does these steps on three devices: torch ops on the GPU, a numba kernel for the BEV IoU with host round trips either
// side (utils3d/rotate_nms_3d_torch.py:63-75), and the greedy loop in C++ on the host.  Here the whole chain stays on the device: one
// ranking kernel (counting rank, deterministic: ties -> lower index first), one decode kernel, one pairwise kernel that writes the
// suppression bit matrix, one single-CTA sweep over it.  9,248 anchors -> 1,500 candidates -> 2.25 M pairs at the B470 building.
#include "../../include/scn_b200.h"
#include "common.cuh"
#include <algorithm>

namespace scn {
namespace {

// ------------------------------------------------------------------ ranking (torch.topk(sorted=True) semantics)
// rank[i] = number of elements that come before element i in descending order (ties: lower index first).
constexpr int kRankTile = 1024;
__global__ void __launch_bounds__(256) k_rank_desc(const float *__restrict__ v, long n, int sigmoid, long k, float *__restrict__ outVal, long *__restrict__ outIdx) {
  __shared__ float tile[kRankTile];
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  auto val = [&](long j) { const float x = __ldg(v + j); return sigmoid ? 1.f / (1.f + expf(-x)) : x; };
  const float mine = i < n ? val(i) : 0.f;
  long rank = 0;
  for (long base = 0; base < n; base += kRankTile) {
    __syncthreads();
    for (int t = threadIdx.x; t < kRankTile; t += blockDim.x) tile[t] = base + t < n ? val(base + t) : -INFINITY;
    __syncthreads();
    const int cnt = (int)min((long)kRankTile, n - base);
    if (i < n) {
      int r = 0;
      if (base + cnt <= i) { // every j of this tile is < i: ties count
#pragma unroll 8
        for (int t = 0; t < cnt; t++) r += tile[t] >= mine;
      } else if (base > i) { // every j > i: strict
#pragma unroll 8
        for (int t = 0; t < cnt; t++) r += tile[t] > mine;
      } else {
        for (int t = 0; t < cnt; t++) r += (tile[t] > mine) || (tile[t] == mine && base + t < i);
      }
      rank += r;
    }
  }
  if (i < n && rank < k) {
    outVal[rank] = mine;
    outIdx[rank] = i;
  }
}

// ------------------------------------------------------------------ box decode (BoxCoder3D.decode, smooth_dim or exp)
struct DecodeParams { float w[7]; float clip; int smooth; };
__global__ void k_box_decode(const float *__restrict__ enc, const float *__restrict__ anchors, const long *__restrict__ idx, long n, DecodeParams P, float *__restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long src = idx ? idx[i] : i;
  float t[7], a[7];
#pragma unroll
  for (int c = 0; c < 7; c++) { t[c] = enc[src * 7 + c] / P.w[c]; a[c] = anchors[src * 7 + c]; }
#pragma unroll
  for (int c = 3; c < 6; c++) t[c] = fminf(t[c], P.clip);
  const float diag = sqrtf(a[4] * a[4] + a[3] * a[3]);
  float o[7];
  o[0] = t[0] * diag + a[0];
  o[1] = t[1] * diag + a[1];
  o[2] = t[2] * a[5] + a[2];
#pragma unroll
  for (int c = 3; c < 6; c++) o[c] = P.smooth ? (t[c] + 1.f) * a[c] : expf(t[c]) * a[c];
  const float yaw = t[6] + a[6];
  const float pi = 3.14159265358979323846f;
  o[6] = yaw - floorf(yaw / pi + 0.5f) * pi; // limit_period(yaw, 0.5, pi)
#pragma unroll
  for (int c = 0; c < 7; c++) out[i * 7 + c] = o[c];
}

// ------------------------------------------------------------------ rotated rectangles in the ground plane
// [cx, cy, dx, dy, angle] -> corners, clockwise when angle is positive (rbbox_to_corners, nms_gpu.py:355-377; the same corners
// as center_to_corner_box2d, box_np_ops.py:374-394).
struct Quad { float x[4], y[4]; };
__device__ __forceinline__ Quad corners_of(float cx, float cy, float dx, float dy, float ang) {
  float s, c;
  sincosf(ang, &s, &c);
  const float hx = dx * 0.5f, hy = dy * 0.5f;
  const float lx[4] = {-hx, -hx, hx, hx}, ly[4] = {-hy, hy, hy, -hy};
  Quad q;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    q.x[i] = c * lx[i] + s * ly[i] + cx;
    q.y[i] = -s * lx[i] + c * ly[i] + cy;
  }
  return q;
}
// Area of the intersection of two convex quadrilaterals: A clipped against the four half planes of B (Sutherland-Hodgman; the
// polygon stays ordered, so no vertex sort is needed), then the shoelace formula.
__device__ float quad_intersection_area(const Quad &A, const Quad &B) {
  float px[10], py[10], qx[10], qy[10];
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { px[i] = A.x[i]; py[i] = A.y[i]; }
  // orientation of B (sign of its doubled area)
  float orient = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) { const int j = (i + 1) & 3; orient += B.x[i] * B.y[j] - B.x[j] * B.y[i]; }
  const float sgn = orient >= 0.f ? 1.f : -1.f;
  for (int e = 0; e < 4 && n > 0; e++) {
    const float ax = B.x[e], ay = B.y[e], ex = B.x[(e + 1) & 3] - ax, ey = B.y[(e + 1) & 3] - ay;
    int m = 0;
    float sx = px[n - 1], sy = py[n - 1];
    float sd = sgn * (ex * (sy - ay) - ey * (sx - ax));
    for (int i = 0; i < n; i++) {
      const float cx = px[i], cy = py[i];
      const float cd = sgn * (ex * (cy - ay) - ey * (cx - ax));
      if ((cd >= 0.f) != (sd >= 0.f)) { // the edge s -> c crosses the line
        const float t = sd / (sd - cd);
        qx[m] = sx + t * (cx - sx);
        qy[m] = sy + t * (cy - sy);
        m++;
      }
      if (cd >= 0.f) { qx[m] = cx; qy[m] = cy; m++; }
      sx = cx; sy = cy; sd = cd;
    }
    n = m;
    for (int i = 0; i < n; i++) { px[i] = qx[i]; py[i] = qy[i]; }
  }
  if (n < 3) return 0.f;
  float a2 = 0.f;
  for (int i = 0; i < n; i++) { const int j = i + 1 == n ? 0 : i + 1; a2 += (px[i] - px[0]) * (py[j] - py[0]) - (px[j] - px[0]) * (py[i] - py[0]); }
  return fabsf(a2) * 0.5f;
}
// devRotateIoUEval (nms_gpu.py:548-570): rbox1 = the QUERY box (second argument of rotate_iou_gpu_eval), rbox2 = the box
__device__ __forceinline__ float criterion_iou(float inter, float area1, float area2, float b2dx, float b2dy, int criterion) {
  if (criterion == -1) return inter / (area1 + area2 - inter);
  if (criterion == 0) return inter / area1;
  if (criterion == 1) return inter / area2;
  if (criterion == 2) {
    const bool thin = fminf(b2dx, b2dy) / fmaxf(b2dx, b2dy) < 0.25f;
    return thin ? inter / (area2 + fmaxf(0.f, area1 * 0.5f - inter)) : inter / (area1 + area2 - inter);
  }
  return inter;
}
struct IouParams { float augTY, augTZ, augAY, augAZ; int criterion, onlyXY; };
// boxes_iou_3d (utils3d/rotate_nms_3d_torch.py:22-84): iou[t][a] = BEV IoU(target t, anchor a) * IoU of their z extents.
// Boxes are yx_zb rows [x, y, z_bottom, size_y, size_x, size_z, yaw]; the BEV rectangle is [x, y, size_y, size_x, yaw] (:64-65).
__global__ void __launch_bounds__(128) k_boxes_iou_3d(const float *__restrict__ T, long nT, const float *__restrict__ A, long nA, IouParams P, float *__restrict__ iou) {
  const long a = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long t = blockIdx.y;
  if (a >= nA || t >= nT) return;
  float tb[7], ab[7];
#pragma unroll
  for (int c = 0; c < 7; c++) { tb[c] = __ldg(T + t * 7 + c); ab[c] = __ldg(A + a * 7 + c); }
  tb[3] = fmaxf(tb[3], P.augTY); ab[3] = fmaxf(ab[3], P.augAY);
  tb[5] = fmaxf(tb[5], P.augTZ); ab[5] = fmaxf(ab[5], P.augAZ);
  // check_same_boxes (nms_gpu.py:657-667): identical rectangles are forced to 1
  bool same = true;
  const int bev[5] = {0, 1, 3, 4, 6};
#pragma unroll
  for (int c = 0; c < 5; c++) same = same && fabsf(tb[bev[c]] - ab[bev[c]]) < 1e-6f;
  float v;
  if (same) v = 1.f;
  else {
    const Quad qa = corners_of(ab[0], ab[1], ab[3], ab[4], ab[6]), qt = corners_of(tb[0], tb[1], tb[3], tb[4], tb[6]);
    const float inter = quad_intersection_area(qa, qt);
    // rotate_iou_gpu_eval(targets, anchors): boxes = targets, query = anchors -> rbox1 = anchor, rbox2 = target
    v = criterion_iou(inter, ab[3] * ab[4], tb[3] * tb[4], tb[3], tb[4], P.criterion);
  }
  if (!P.onlyXY) { // iou_one_dim (:7-20): overlap / common of [z, z + size]; may be negative
    const float t0 = tb[2], t1 = tb[2] + tb[5], a0 = ab[2], a1 = ab[2] + ab[5];
    v *= (fminf(a1, t1) - fmaxf(a0, t0)) / (fmaxf(a1, t1) - fminf(a0, t0));
  }
  iou[t * nA + a] = v;
}

// ------------------------------------------------------------------ rotate_nms_3d
// Candidates sorted by descending score.  bit (i, j), j > i, is set when box i suppresses box j: their 3-D IoU (criterion -1, no
// thickness augmentation) is positive AND the IoU of their BEV polygons reaches the threshold (rotate_nms_3d_cc passes the 3-D IoU
// matrix as the pre-filter argument of spconv's rotate_non_max_suppression_cpu, nms_cpu.py:35-43).
__global__ void __launch_bounds__(64) k_nms_mask(const float *__restrict__ boxes, int n, float thresh, unsigned long long *__restrict__ mask, int words) {
  const int i = blockIdx.y * 64 + threadIdx.x, jw = blockIdx.x;
  __shared__ float sb[64][7];
  {
    const int j = jw * 64 + threadIdx.x;
    for (int c = 0; c < 7; c++) sb[threadIdx.x][c] = j < n ? boxes[(long)j * 7 + c] : 0.f;
  }
  __syncthreads();
  if (i >= n) return;
  unsigned long long bits = 0;
  if (jw * 64 + 63 > i) {
    float b[7];
    for (int c = 0; c < 7; c++) b[c] = boxes[(long)i * 7 + c];
    const Quad qi = corners_of(b[0], b[1], b[3], b[4], b[6]);
    const float areaI = b[3] * b[4];
    for (int t = 0; t < 64; t++) {
      const int j = jw * 64 + t;
      if (j <= i || j >= n) continue;
      const float *o = sb[t];
      bool same = true;
      const int bev[5] = {0, 1, 3, 4, 6};
      for (int c = 0; c < 5; c++) same = same && fabsf(b[bev[c]] - o[bev[c]]) < 1e-6f;
      const Quad qj = corners_of(o[0], o[1], o[3], o[4], o[6]);
      const float inter = quad_intersection_area(qj, qi), areaJ = o[3] * o[4];
      const float iou2 = same ? 1.f : inter / (areaI + areaJ - inter);
      const float i0 = b[2], i1 = b[2] + b[5], j0 = o[2], j1 = o[2] + o[5];
      const float iouz = (fminf(j1, i1) - fmaxf(j0, i0)) / (fmaxf(j1, i1) - fminf(j0, i0));
      if (!(iou2 * iouz > 0.f)) continue;
      const float poly = inter / (areaI + areaJ - inter); // polygon intersection over polygon union
      if (poly >= thresh) bits |= 1ull << t;
    }
  }
  mask[(long)i * words + jw] = bits;
}
// One CTA: greedy sweep in score order, 64 candidates at a time.  Within a block of 64 one thread resolves who survives from the
// block's diagonal mask words (register work); then every thread ORs the rows of the survivors into its word of the `removed`
// bitmap.  keep[] = indices of the survivors (into the caller's boxes), at most postMax; stops as soon as postMax are found.
__global__ void __launch_bounds__(256) k_nms_sweep(const unsigned long long *__restrict__ mask, int n, int words, const long *__restrict__ sortedIdx, long postMax,
                                                   long *__restrict__ keep, long *__restrict__ nKeep) {
  extern __shared__ unsigned long long removed[];
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_keptBits;
  __shared__ long s_kept;
  for (int w = threadIdx.x; w < words; w += blockDim.x) removed[w] = 0;
  if (threadIdx.x == 0) s_kept = 0;
  __syncthreads();
  for (int b = 0; b < words; b++) {
    if (threadIdx.x < 64) {
      const int row = b * 64 + threadIdx.x;
      s_diag[threadIdx.x] = row < n ? mask[(long)row * words + b] : 0ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long alive = ~removed[b], kb = 0;
      const int left = n - b * 64;
      if (left < 64) alive &= (1ull << left) - 1ull;
      long k = s_kept;
      for (int t = 0; t < 64 && k < postMax; t++)
        if ((alive >> t) & 1ull) {
          kb |= 1ull << t;
          alive &= ~s_diag[t];
          keep[k++] = sortedIdx ? sortedIdx[b * 64 + t] : (long)(b * 64 + t);
        }
      s_keptBits = kb;
      s_kept = k;
    }
    __syncthreads();
    if (s_kept >= postMax) break;
    const unsigned long long kb = s_keptBits;
    for (int w = b + 1 + threadIdx.x; w < words; w += blockDim.x) {
      unsigned long long acc = removed[w];
      for (unsigned long long rest = kb; rest; rest &= rest - 1) acc |= mask[(long)(b * 64 + __ffsll((long long)rest) - 1) * words + w];
      removed[w] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *nKeep = s_kept;
}
__global__ void k_gather_boxes(const float *__restrict__ boxes, const long *__restrict__ idx, long n, float *__restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n * 7) out[i] = boxes[idx[i / 7] * 7 + i % 7];
}

// ------------------------------------------------------------------ voxeliser
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double from_ordered_bits(unsigned long long o) {
  const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
struct Affine { double m[9]; };
__device__ __forceinline__ void affine_apply(const float *__restrict__ xyz, long i, const Affine &M, double a[3]) {
  const double x = xyz[i * 3], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
#pragma unroll
  for (int c = 0; c < 3; c++) a[c] = x * M.m[c] + y * M.m[3 + c] + z * M.m[6 + c]; // row vector times matrix (np.matmul(a, m))
}
// per-coordinate minimum and maximum of the transformed points: stats[0..2] = min, [3..5] = max (ordered bit patterns)
__global__ void __launch_bounds__(256) k_voxel_extent(const float *__restrict__ xyz, long n, Affine M, unsigned long long *stats) {
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double a[3];
    affine_apply(xyz, i, M, a);
#pragma unroll
    for (int c = 0; c < 3; c++) { lo[c] = fmin(lo[c], a[c]); hi[c] = fmax(hi[c], a[c]); }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    for (int o = 16; o; o >>= 1) { lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o)); hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(stats + c, ordered_bits(lo[c])); atomicMax(stats + 3 + c, ordered_bits(hi[c])); }
  }
}
struct VoxelParams { Affine M; double offset[3]; double scale; double full[3]; int nFeat, xyzInFeats; long batchIdx; int locCols; };
constexpr int kVoxBlock = 1024;
__device__ __forceinline__ bool voxel_of(const float *xyz, long i, const VoxelParams &P, double a[3]) {
  affine_apply(xyz, i, P.M, a);
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 3; c++) { a[c] += P.offset[c]; ok = ok && a[c] >= 0.0 && a[c] < P.full[c]; }
  return ok;
}
__global__ void __launch_bounds__(kVoxBlock) k_voxel_count(const float *__restrict__ xyz, long n, VoxelParams P, int *__restrict__ blockCnt) {
  const long i = blockIdx.x * (long)kVoxBlock + threadIdx.x;
  double a[3];
  const int ok = i < n && voxel_of(xyz, i, P, a);
  const int c = __syncthreads_count(ok);
  if (threadIdx.x == 0) blockCnt[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) k_block_scan(int *cnt, int nBlocks, long *total) { // exclusive scan of the block counts, one CTA
  __shared__ long s_w[32];
  __shared__ long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nBlocks; base += 1024) {
    const int i = base + threadIdx.x;
    const long v = i < nBlocks ? cnt[i] : 0;
    long incl = v;
    for (int o = 1; o < 32; o <<= 1) { const long t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      long w = s_w[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) { const long t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
      s_w[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const long excl = s_carry + s_w[threadIdx.x >> 5] + incl - v;
    if (i < nBlocks) cnt[i] = (int)excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(kVoxBlock) k_voxel_write(const float *__restrict__ xyz, const float *__restrict__ feats, long n, VoxelParams P,
                                                            const int *__restrict__ blockOff, long *__restrict__ locs, float *__restrict__ featsOut) {
  __shared__ int s_w[kVoxBlock / 32];
  const long i = blockIdx.x * (long)kVoxBlock + threadIdx.x;
  double a[3];
  const int ok = i < n && voxel_of(xyz, i, P, a);
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_w[wid] = __popc(bal);
  __syncthreads();
  if (wid == 0) {
    int v = s_w[lane], incl = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    s_w[lane] = incl - v;
  }
  __syncthreads();
  if (!ok) return;
  const long dst = blockOff[blockIdx.x] + s_w[wid] + __popc(bal & ((1u << lane) - 1));
#pragma unroll
  for (int c = 0; c < 3; c++) locs[dst * P.locCols + c] = (long)a[c]; // torch .long(): truncation (values are >= 0)
  if (P.locCols == 4) locs[dst * 4 + 3] = P.batchIdx;
  for (int c = 0; c < P.nFeat; c++) featsOut[dst * P.nFeat + c] = (P.xyzInFeats && c < 3) ? (float)(a[c] / P.scale) : feats[i * P.nFeat + c];
}

} // namespace
} // namespace scn

extern "C" {

int scn_top_k_descending(const float *values, long n, int apply_sigmoid, long k, float *out_values, long *out_indices, void *stream) {
  SCN_CHECK(n >= 0 && k >= 0 && k <= n, "top-k: k must be in [0, n]");
  if (n == 0 || k == 0) return 0;
  SCN_CHECK(values && out_values && out_indices, "top-k: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  scn::k_rank_desc<<<scn::cdiv(n, 256), 256, 0, scn::LS(s)>>>(values, n, apply_sigmoid, k, out_values, out_indices);
  SCN_CUDA(cudaGetLastError());
  return 0;
}

int scn_box_decode_3d(const float *encodings, const float *anchors, const long *indices, long n, const float weights[7], float clip, int smooth_dim, float *boxes,
                      void *stream) {
  if (n == 0) return 0;
  SCN_CHECK(encodings && anchors && boxes && weights && n > 0, "box decode: bad arguments");
  scn::DecodeParams P;
  for (int c = 0; c < 7; c++) P.w[c] = weights[c];
  P.clip = clip;
  P.smooth = smooth_dim;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  scn::k_box_decode<<<scn::cdiv(n, 128), 128, 0, scn::LS(s)>>>(encodings, anchors, indices, n, P, boxes);
  SCN_CUDA(cudaGetLastError());
  return 0;
}

int scn_boxes_iou_3d(const float *targets, long n_targets, const float *anchors, long n_anchors, const float aug_thickness[4], int criterion, int only_xy, float *iou,
                     void *stream) {
  if (n_targets == 0 || n_anchors == 0) return 0;
  SCN_CHECK(targets && anchors && iou && n_targets > 0 && n_anchors > 0 && n_targets < 65536, "boxes_iou_3d: bad arguments (at most 65535 targets per call)");
  scn::IouParams P{0.f, 0.f, 0.f, 0.f, criterion, only_xy};
  if (aug_thickness) { P.augTY = aug_thickness[0]; P.augTZ = aug_thickness[1]; P.augAY = aug_thickness[2]; P.augAZ = aug_thickness[3]; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  scn::k_boxes_iou_3d<<<dim3(scn::cdiv(n_anchors, 128), (unsigned)n_targets), 128, 0, scn::LS(s)>>>(targets, n_targets, anchors, n_anchors, P, iou);
  SCN_CUDA(cudaGetLastError());
  return 0;
}

int scn_rotate_nms_3d(const float *boxes, const float *scores, long n, long pre_max_size, long post_max_size, float iou_threshold, long *keep, long *n_keep,
                      void *stream) {
  SCN_CHECK(n >= 0 && keep && n_keep, "rotate_nms_3d: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) { SCN_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(long), s)); return 0; }
  SCN_CHECK(boxes && scores, "rotate_nms_3d: null boxes / scores");
  const long m = pre_max_size > 0 ? std::min(n, pre_max_size) : n;
  SCN_CHECK(m <= 16384, "rotate_nms_3d: at most 16384 candidates after pre_max_size");
  if (post_max_size <= 0 || post_max_size > m) post_max_size = m;
  const int words = scn::cdiv(m, 64);
  float *sScore = nullptr, *sBoxes = nullptr;
  long *sIdx = nullptr;
  unsigned long long *mask = nullptr;
  SCN_CUDA(cudaMallocAsync((void **)&sScore, m * 4, s));
  SCN_CUDA(cudaMallocAsync((void **)&sIdx, m * 8, s));
  SCN_CUDA(cudaMallocAsync((void **)&sBoxes, m * 7 * 4, s));
  SCN_CUDA(cudaMallocAsync((void **)&mask, (size_t)m * words * 8, s));
  scn::k_rank_desc<<<scn::cdiv(n, 256), 256, 0, scn::LS(s)>>>(scores, n, 0, m, sScore, sIdx); // torch.topk(scores, k = pre_max_size)
  scn::k_gather_boxes<<<scn::cdiv(m * 7, 256), 256, 0, scn::LS(s)>>>(boxes, sIdx, m, sBoxes);
  scn::k_nms_mask<<<dim3(words, words), 64, 0, scn::LS(s)>>>(sBoxes, (int)m, iou_threshold, mask, words);
  scn::k_nms_sweep<<<1, 256, words * 8, scn::LS(s)>>>(mask, (int)m, words, sIdx, post_max_size, keep, n_keep);
  SCN_CUDA(cudaGetLastError());
  cudaFreeAsync(sScore, s);
  cudaFreeAsync(sIdx, s);
  cudaFreeAsync(sBoxes, s);
  cudaFreeAsync(mask, s);
  return 0;
}

int scn_voxelize_extent(const float *xyz, long n, const double matrix[9], double extent_min[3], double extent_max[3], void *stream) {
  SCN_CHECK(xyz && matrix && n > 0, "voxelize: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long *stats = nullptr;
  SCN_CUDA(cudaMallocAsync((void **)&stats, 6 * 8, s));
  SCN_CUDA(cudaMemsetAsync(stats, 0xff, 3 * 8, s));
  SCN_CUDA(cudaMemsetAsync(stats + 3, 0, 3 * 8, s));
  scn::Affine M;
  for (int i = 0; i < 9; i++) M.m[i] = matrix[i];
  scn::k_voxel_extent<<<std::min(scn::cdiv(n, 256), 148 * 8), 256, 0, scn::LS(s)>>>(xyz, n, M, stats);
  unsigned long long h[6];
  SCN_CUDA(cudaMemcpyAsync(h, stats, sizeof h, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(stats, s);
  for (int c = 0; c < 3; c++) { extent_min[c] = scn::from_ordered_bits(h[c]); extent_max[c] = scn::from_ordered_bits(h[3 + c]); }
  return 0;
}

int scn_voxelize(const float *xyz, const float *feats, long n, int n_feat, const double matrix[9], const double offset[3], double scale, const double full_scale[3],
                 int xyz_in_feats, long batch_index, int loc_cols, long *locs, float *feats_out, long *n_kept, void *stream) {
  SCN_CHECK(xyz && matrix && offset && full_scale && locs && n_kept && n >= 0 && (loc_cols == 3 || loc_cols == 4) && (n_feat == 0 || (feats && feats_out)),
            "voxelize: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) { *n_kept = 0; return 0; }
  scn::VoxelParams P;
  for (int i = 0; i < 9; i++) P.M.m[i] = matrix[i];
  for (int c = 0; c < 3; c++) { P.offset[c] = offset[c]; P.full[c] = full_scale[c]; }
  P.scale = scale;
  P.nFeat = n_feat;
  P.xyzInFeats = xyz_in_feats;
  P.batchIdx = batch_index;
  P.locCols = loc_cols;
  const int nb = scn::cdiv(n, scn::kVoxBlock);
  int *cnt = nullptr;
  long *total = nullptr;
  SCN_CUDA(cudaMallocAsync((void **)&cnt, (size_t)nb * 4, s));
  SCN_CUDA(cudaMallocAsync((void **)&total, 8, s));
  scn::k_voxel_count<<<nb, scn::kVoxBlock, 0, scn::LS(s)>>>(xyz, n, P, cnt);
  scn::k_block_scan<<<1, 1024, 0, scn::LS(s)>>>(cnt, nb, total);
  scn::k_voxel_write<<<nb, scn::kVoxBlock, 0, scn::LS(s)>>>(xyz, feats, n, P, cnt, locs, feats_out);
  SCN_CUDA(cudaGetLastError());
  SCN_CUDA(cudaMemcpyAsync(n_kept, total, 8, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(cnt, s);
  cudaFreeAsync(total, s);
  return 0;
}
}
