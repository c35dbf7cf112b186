// tcgen05 tensor-core gather-GEMM for sm_100a: the output-stationary sparse convolution
//   out[row(p)] = sum_k  in[nbr[p][k]] @ W[k]
// with fp32 accumulators in TMEM across ALL filter offsets (one store per output element, no
// read-modify-write, no atomics), A rows gathered with 16-byte cp.async into 128B-swizzled
// shared memory, W[k] streamed with cp.async.bulk from a pre-swizzled image, MMAs issued by one
// thread (tcgen05.mma kind::tf32, M=128, N=Cout, K=8).
//
// Persistent kernel, one CTA per SM, warp-specialised:
//   warps 0-3  epilogue   (tcgen05.ld -> registers -> global rows; warp w owns TMEM lanes 32w..)
//   warps 4-7  A producer (gather rows; 8 threads per 128-byte row chunk, coalesced)
//   warp  8    MMA issuer (lane 0) + TMEM allocation
//   warp  9    B loader   (lane 0, bulk copies)
// A CTA owns a "supertile" of T = min(4, 512/Cout) tiles of 128 output sites whose accumulators
// live in TMEM simultaneously, so each weight slice W[k][32 channels] is fetched once per
// supertile instead of once per tile (L2 -> SM traffic of B is 1/T of A's).
//
// Replaces dConvolution_KMxKN_forwardA/B (SCN/CUDA/Convolution.cu:57-203: SIMT tiles, fp64
// accumulators, one launch + one blocking H2D rule copy per filter offset).
#include "common.cuh"
#include <stdlib.h>

namespace scn {

constexpr int kTileM = 128;
constexpr int kAStageBytes = kTileM * 128; // 128 rows x 32 tf32
constexpr int kInflight = 5;               // A stages a producer thread keeps in flight
constexpr int kThreads = 320;

struct TcParams {
  const float *in;
  float *out;
  const float *wimg;
  const float *bias;
  const int *nbr;
  const int *outRow;
  const unsigned long long *tileMask;
  const int *tileW; // optional: weight slice per tile (deconvolution plans; then K == 1 and T == 1)
  int nOut, K, Cin, Cout, nTiles, T, nSuper, SA, SB;
  int dbg;    // developer switches (SCN_TC_DBG): 1 = skip A gathers, 2 = skip B copies, 4 = skip MMAs
  int kSplit; // > 1: the filter offsets of a supertile are split over kSplit CTAs, epilogue accumulates atomically
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t srcBytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(dTmem),
      "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ unsigned long long range_mask(int lo, int hi) { // bits [lo, hi)
  unsigned long long a = hi >= 64 ? ~0ull : ((1ull << hi) - 1ull);
  return a & ~((1ull << lo) - 1ull);
}

// ------------------------------------------------------------------ kernel
// Work item = (supertile of T <= 2 tiles, range of filter offsets).  Accumulators are double
// buffered in TMEM (2 x T x Cout <= 512 columns) so the epilogue of item i overlaps the main loop
// of item i+1.
__global__ void __launch_bounds__(kThreads, 1) conv_plan_tc(TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nc = P.Cin / 32;                 // K-atoms per filter offset
  const int nWork = P.nSuper * P.kSplit;     // work items
  const int bStageBytes = P.Cout * 128;      // Cout rows x 32 tf32
  unsigned char *sA = smem;
  unsigned char *sB = sA + (size_t)P.SA * kAStageBytes;
  float *sEpi = reinterpret_cast<float *>(sB + (size_t)P.SB * bStageBytes); // 4 warps x 32 rows x 36 floats
  int *sIds = reinterpret_cast<int *>(sEpi + 4 * 32 * 36);                   // T*128*K neighbour ids of the current work item
  const int idsPerItem = P.T * kTileM * P.K;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sIds + (size_t)idsPerItem);
  uint64_t *aFull = bars, *aEmpty = bars + P.SA, *bFull = bars + 2 * P.SA, *bEmpty = bars + 2 * P.SA + P.SB;
  uint64_t *accFull = bars + 2 * P.SA + 2 * P.SB, *accEmpty = accFull + 2;
  uint64_t *idsFull = accEmpty + 2;
  uint32_t *tmemSlot = reinterpret_cast<uint32_t *>(idsFull + 2);

  if (tid == 0) {
    for (int i = 0; i < P.SA; i++) { mbar_init(smem_u32(aFull + i), 1); mbar_init(smem_u32(aEmpty + i), 1); }
    for (int i = 0; i < P.SB; i++) { mbar_init(smem_u32(bFull + i), 1); mbar_init(smem_u32(bEmpty + i), 1); }
    for (int i = 0; i < 2; i++) { mbar_init(smem_u32(accFull + i), 1); mbar_init(smem_u32(accEmpty + i), 4); mbar_init(smem_u32(idsFull + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemSlot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemSlot;
  const int accCols = P.T * P.Cout; // columns of one accumulator stage

  if (warp < 4) {
    // ============================ epilogue ============================
    float *stg = sEpi + warp * (32 * 36);
    int it = 0;
    for (int wi = blockIdx.x; wi < nWork; wi += gridDim.x, it++) {
      const int st = wi / P.kSplit, part = wi % P.kSplit;
      const unsigned long long kmask = range_mask(P.K * part / P.kSplit, P.K * (part + 1) / P.kSplit);
      const int a = it & 1;
      mbar_wait(smem_u32(accFull + a), (it >> 1) & 1);
      tc_fence_after();
      for (int t = 0; t < P.T; t++) {
        const int tile = st * P.T + t;
        if (tile >= P.nTiles) break;
        const bool started = (__ldg(P.tileMask + tile) & kmask) != 0ull;
        if (!started && P.kSplit > 1) continue; // nothing to add
        const int p = tile * kTileM + warp * 32 + lane;
        const int myRow = p < P.nOut ? __ldg(P.outRow + p) : -1;
        for (int c0 = 0; c0 < P.Cout; c0 += 32) {
          uint32_t v[32];
          if (started) {
            tmem_ld32(tmemBase + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * accCols + t * P.Cout + c0), v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = 0u;
          }
          // stage the 32x32 block so that each store instruction writes whole 128-byte row segments
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4 *>(stg + lane * 36 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          __syncwarp();
          const int cc = (lane & 7) * 4;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.bias && part == 0) bv = __ldg(reinterpret_cast<const float4 *>(P.bias + c0 + cc));
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int r = i * 4 + (lane >> 3);
            const int row = __shfl_sync(0xffffffffu, myRow, r);
            float4 o = *reinterpret_cast<const float4 *>(stg + r * 36 + cc);
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            if (row >= 0) {
              float *dst = P.out + (size_t)row * P.Cout + c0 + cc;
              if (P.kSplit == 1) *reinterpret_cast<float4 *>(dst) = o;
              else { atomicAdd(dst, o.x); atomicAdd(dst + 1, o.y); atomicAdd(dst + 2, o.z); atomicAdd(dst + 3, o.w); } // pre-zeroed by the launcher
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(accEmpty + a));
    }
  } else if (warp < 8) {
    // ============================ A producer ============================
    // Each of the 4 producer warps gathers WHOLE stages (128 rows x 128 B) on its own: warp w owns
    // the stages n with n % 4 == w, so four gathers proceed independently and every warp keeps two
    // of its own stages in flight (SA = 8 ring slots; slot = n % SA is owned by warp slot % 4).
    const int pw = warp - 4;
    const int ptid = tid - 128;
    const int chunk = lane & 7, rsub = lane >> 3; // lane -> 16-byte chunk of rows rsub, rsub+4, ...
    const uint32_t idsBytes = (uint32_t)idsPerItem * 4u;
    auto issue_ids = [&](int wi) {
      const int st = wi / P.kSplit;
      const uint32_t bar = smem_u32(idsFull);
      mbar_arrive_expect_tx(bar, idsBytes);
      bulk_g2s(smem_u32(sIds), P.nbr + (size_t)st * idsPerItem, idsBytes, bar);
    };
    if (ptid == 0 && (int)blockIdx.x < nWork) issue_ids(blockIdx.x);
    uint32_t n = 0;                 // global stage counter (all warps count every stage)
    uint32_t pend0 = 0, pend1 = 0;  // ring slots (+1) of this warp's gathers that are not yet published
    int it = 0;
    for (int wi = blockIdx.x; wi < nWork; wi += gridDim.x, it++) {
      const int st = wi / P.kSplit, part = wi % P.kSplit;
      const int kLo = P.K * part / P.kSplit, kHi = P.K * (part + 1) / P.kSplit;
      const unsigned long long kmask = range_mask(kLo, kHi);
      unsigned long long m[2], uni = 0;
      for (int t = 0; t < 2; t++) { m[t] = (t < P.T && st * P.T + t < P.nTiles) ? (__ldg(P.tileMask + st * P.T + t) & kmask) : 0ull; uni |= m[t]; }
      mbar_wait(smem_u32(idsFull), it & 1);
      for (int k = kLo; k < kHi; k++) {
        if (!((uni >> k) & 1ull)) continue;
        for (int c = 0; c < nc; c++) {
#pragma unroll
          for (int t = 0; t < 2; t++) {
            if (!((m[t] >> k) & 1ull)) continue;
            const uint32_t mine = n & 3u, slot = n % (uint32_t)P.SA, round = n / (uint32_t)P.SA;
            n++;
            if (mine != (uint32_t)pw) continue;
            // publish the older of my two gathers before starting a third
            if (pend0 && pend1) {
              cp_async_wait<1>();
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(smem_u32(aFull + (pend0 - 1)));
              pend0 = pend1; pend1 = 0;
            }
            mbar_wait(smem_u32(aEmpty + slot), (round & 1u) ^ 1u);
            const uint32_t sbase = smem_u32(sA + (size_t)slot * kAStageBytes);
            const int *ids = sIds + (size_t)(t * kTileM) * P.K + k;
            if (!(P.dbg & 1)) {
#pragma unroll 8
              for (int i = 0; i < 32; i++) {
                const int row = i * 4 + rsub;
                const int id = ids[row * P.K];
                const float *src = P.in + (size_t)(id >= 0 ? id : 0) * P.Cin + c * 32 + chunk * 4;
                cp_async16(sbase + row * 128 + ((chunk ^ (row & 7)) << 4), src, id >= 0 ? 16u : 0u);
              }
            }
            cp_async_commit();
            if (!pend0) pend0 = slot + 1; else pend1 = slot + 1;
          }
        }
      }
      // all gathers of this item are issued: once every producer warp got here the id buffer is free
      // and the next item's ids stream in while the last stages drain
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (ptid == 0 && wi + (int)gridDim.x < nWork) issue_ids(wi + gridDim.x);
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (pend0) mbar_arrive(smem_u32(aFull + (pend0 - 1)));
        if (pend1) mbar_arrive(smem_u32(aFull + (pend1 - 1)));
      }
      pend0 = pend1 = 0;
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, K-major both, N=Cout, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.Cout >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      uint32_t aStage = 0, aPhase = 0, bStage = 0, bPhase = 0;
      int it = 0;
      for (int wi = blockIdx.x; wi < nWork; wi += gridDim.x, it++) {
        const int st = wi / P.kSplit, part = wi % P.kSplit;
        const unsigned long long kmask = range_mask(P.K * part / P.kSplit, P.K * (part + 1) / P.kSplit);
        unsigned long long m[2], uni = 0;
        for (int t = 0; t < 2; t++) { m[t] = (t < P.T && st * P.T + t < P.nTiles) ? (__ldg(P.tileMask + st * P.T + t) & kmask) : 0ull; uni |= m[t]; }
        const int a = it & 1;
        mbar_wait(smem_u32(accEmpty + a), ((it >> 1) & 1) ^ 1); // epilogue has drained this accumulator stage
        tc_fence_after();
        uint32_t started = 0;
        for (int k = 0; k < P.K; k++) {
          if (!((uni >> k) & 1ull)) continue;
          for (int c = 0; c < nc; c++) {
            mbar_wait(smem_u32(bFull + bStage), bPhase);
            tc_fence_after();
            const uint64_t bDesc = smem_desc_sw128(smem_u32(sB + (size_t)bStage * bStageBytes));
#pragma unroll
            for (int t = 0; t < 2; t++) {
              if (!((m[t] >> k) & 1ull)) continue;
              mbar_wait(smem_u32(aFull + aStage), aPhase);
              if (!(P.dbg & 8)) tc_fence_after();
              const uint64_t aDesc = smem_desc_sw128(smem_u32(sA + (size_t)aStage * kAStageBytes));
              const uint32_t d = tmemBase + (uint32_t)(a * accCols + t * P.Cout);
              if (!(P.dbg & 4)) {
#pragma unroll
                for (int j = 0; j < 4; j++) // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle atom
                  tc_mma_tf32(d, aDesc + (uint64_t)(j * 2), bDesc + (uint64_t)(j * 2), idesc, ((started >> t) & 1u) | (j > 0));
              }
              started |= 1u << t;
              if (P.dbg & 16) mbar_arrive(smem_u32(aEmpty + aStage)); else tc_commit(smem_u32(aEmpty + aStage));
              if (++aStage == (uint32_t)P.SA) { aStage = 0; aPhase ^= 1; }
            }
            tc_commit(smem_u32(bEmpty + bStage));
            if (++bStage == (uint32_t)P.SB) { bStage = 0; bPhase ^= 1; }
          }
        }
        tc_commit(smem_u32(accFull + a));
      }
    }
  } else {
    // ============================ B loader ============================
    if (lane == 0) {
      uint32_t bStage = 0, bPhase = 0;
      for (int wi = blockIdx.x; wi < nWork; wi += gridDim.x) {
        const int st = wi / P.kSplit, part = wi % P.kSplit;
        const unsigned long long kmask = range_mask(P.K * part / P.kSplit, P.K * (part + 1) / P.kSplit);
        unsigned long long uni = 0;
        for (int t = 0; t < P.T; t++) if (st * P.T + t < P.nTiles) uni |= __ldg(P.tileMask + st * P.T + t) & kmask;
        for (int k = 0; k < P.K; k++) {
          if (!((uni >> k) & 1ull)) continue;
          for (int c = 0; c < nc; c++) {
            mbar_wait(smem_u32(bEmpty + bStage), bPhase ^ 1);
            const uint32_t bar = smem_u32(bFull + bStage);
            if (P.dbg & 2) { mbar_arrive(bar); }
            else {
              mbar_arrive_expect_tx(bar, (uint32_t)bStageBytes);
              bulk_g2s(smem_u32(sB + (size_t)bStage * bStageBytes), P.wimg + ((size_t)(P.tileW ? __ldg(P.tileW + st) : k) * nc + c) * P.Cout * 32, (uint32_t)bStageBytes, bar);
            }
            if (++bStage == (uint32_t)P.SB) { bStage = 0; bPhase ^= 1; }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(512u) : "memory");
  }
}

// W [K][Cin][Cout] fp32 -> per (k, 32-channel atom) the exact shared-memory image of the B operand:
// Cout rows x 128 bytes, 16-byte chunks XOR-swizzled by (row & 7), values rounded to TF32 (rna).
__global__ void k_prep_wimg(const float *__restrict__ W, float *__restrict__ img, int K, int Cin, int Cout) {
  const long n = (long)K * Cin * Cout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long t = i / Cout;
    const int ci = (int)(t % Cin), k = (int)(t / Cin);
    const int c = ci >> 5, j = ci & 31, chunk = j >> 2, within = j & 3;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(W[i]));
    img[((long)k * (Cin >> 5) + c) * Cout * 32 + (long)co * 32 + ((chunk ^ (co & 7)) << 2) + within] = __uint_as_float(r);
  }
}

int tc_available() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) major = 0;
    cached = major == 10;
  }
  return cached;
}

__global__ void k_pad_rows(const float *__restrict__ in, float *__restrict__ out, long n, int C, int Cp) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n * Cp; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    out[i] = c < C ? __ldg(in + (i / Cp) * C + c) : 0.f;
  }
}
int launch_conv_plan_tc(const float *in, float *out, const float *W, const int *nbr, const int *outRow, const unsigned long long *tileMask,
                        int nOut, int K, int Cin, int Cout, const float *bias, int mathMode, cudaStream_t s, const int *tileW, int nWeights,
                        long nInRows) {
  if (nOut == 0) return 0;
  if (Cin % 32 != 0) { // e.g. the 9-channel input convolution: zero-pad rows and weight slices to 32 channels
    const int Cp = (Cin + 31) / 32 * 32;
    float *xp = nullptr, *wp = nullptr;
    SCN_CUDA(cudaMallocAsync((void **)&xp, (size_t)nInRows * Cp * 4, s));
    SCN_CUDA(cudaMallocAsync((void **)&wp, (size_t)nWeights * Cp * Cout * 4, s));
    k_pad_rows<<<stream_grid(nInRows * Cp, 256), 256, 0, LS(s)>>>(in, xp, nInRows, Cin, Cp);
    SCN_CUDA(cudaMemsetAsync(wp, 0, (size_t)nWeights * Cp * Cout * 4, s));
    SCN_CUDA(cudaMemcpy2DAsync(wp, (size_t)Cp * Cout * 4, W, (size_t)Cin * Cout * 4, (size_t)Cin * Cout * 4, nWeights, cudaMemcpyDeviceToDevice, s));
    int r = launch_conv_plan_tc(xp, out, wp, nbr, outRow, tileMask, nOut, K, Cp, Cout, bias, mathMode, s, tileW, nWeights, nInRows);
    cudaFreeAsync(xp, s);
    cudaFreeAsync(wp, s);
    return r;
  }
  SCN_CHECK(Cout % 16 == 0 && Cout >= 16 && Cout <= 256 && K <= 64, "tcgen05 path: unsupported channel counts");
  TcParams P;
  P.in = in; P.out = out; P.bias = bias; P.nbr = nbr; P.outRow = outRow; P.tileMask = tileMask; P.tileW = tileW;
  P.nOut = nOut; P.K = K; P.Cin = Cin; P.Cout = Cout;
  P.nTiles = cdiv(nOut, kTileM);
  // supertile height: 2 tiles share every weight slice when the level is large enough for >= 2
  // work items per SM (accumulators double-buffered: 2 x T x Cout <= 512 TMEM columns); small levels
  // use 1-tile items and split the filter offsets over CTAs so the whole chip works on them.
  const int Tmax = Cout <= 128 ? 2 : 1;
  P.T = (Tmax == 2 && P.nTiles >= 2 * kSMs * 2 && !tileW && K <= 32) ? 2 : 1;
  static int envT = -1, envSA = -1, envSB = -1, envDbg = 0;
  if (envT < 0) {
    envT = getenv("SCN_TC_T") ? atoi(getenv("SCN_TC_T")) : 0;
    envSA = getenv("SCN_TC_SA") ? atoi(getenv("SCN_TC_SA")) : 0;
    envSB = getenv("SCN_TC_SB") ? atoi(getenv("SCN_TC_SB")) : 0;
    envDbg = getenv("SCN_TC_DBG") ? atoi(getenv("SCN_TC_DBG")) : 0;
  }
  P.dbg = envDbg;
  if (envT > 0 && envT <= Tmax && !tileW) P.T = envT;
  P.nSuper = cdiv(P.nTiles, P.T);
  P.kSplit = 1;
  if (P.nSuper < kSMs / 2 && !tileW) P.kSplit = std::max(1, std::min(K, kSMs / P.nSuper));
  // weight-slice ring: ~48 KB deep (a slice is only Cout x 128 B, and one is needed per filter offset)
  P.SB = envSB > 0 ? envSB : std::max(2, std::min(12, (48 * 1024) / (Cout * 128)));
  size_t fixed = (size_t)P.SB * Cout * 128 + 4 * 32 * 36 * 4 + (size_t)P.T * kTileM * K * 4 + 64 * 8 + 16;
  P.SA = (int)std::min<size_t>(envSA > 0 ? envSA : 8, (227 * 1024 - fixed) / kAStageBytes);
  P.SA = P.SA >= 8 ? 8 : 4; // ring slots are statically owned by the 4 producer warps
  SCN_CHECK((size_t)P.SA * kAStageBytes + fixed <= 227 * 1024, "tcgen05 path: shared memory budget exceeded");
  size_t smem = (size_t)P.SA * kAStageBytes + fixed;
  if (P.kSplit > 1) SCN_CUDA(cudaMemsetAsync(out, 0, (size_t)nOut * Cout * 4, s));
  float *wimg = nullptr;
  SCN_CUDA(cudaMallocAsync((void **)&wimg, (size_t)nWeights * Cin * Cout * 4, s));
  P.wimg = wimg;
  k_prep_wimg<<<stream_grid((long)nWeights * Cin * Cout, 256), 256, 0, LS(s)>>>(W, wimg, nWeights, Cin, Cout);
  static bool attr = false;
  if (!attr) {
    SCN_CUDA(cudaFuncSetAttribute(conv_plan_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = std::min(P.nSuper * P.kSplit, kSMs);
  conv_plan_tc<<<grid, kThreads, smem, LS(s)>>>(P);
  SCN_CUDA(cudaGetLastError());
  cudaFreeAsync(wimg, s);
  return 0;
}

} // namespace scn
