// Developer probe: how fast can one SM fill shared memory with GATHERED rows (random 256-byte rows of a table), by mechanism?
//   0 cp.async.cg 16 B (what conv_plan_tc uses)      1 cp.async.ca 16 B          2 ld.global.nc.v4 + st.shared.v4
//   3 cp.async.bulk 128 B per copy (TMA engine, linear destination)             4 cp.async.bulk 256 B per copy
// Persistent CTAs (2 per SM, 128 gather threads each), ring of S slots x 32 KB (= 128 rows x 256 B), table 300 MB (HBM) or 48 MB (L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_paths_probe gather_paths_probe.cu && ./gather_paths_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int kRows = 128, kRowBytes = 256, kSlot = kRows * kRowBytes;
template <int MODE, int S>
__global__ void __launch_bounds__(128, 2) k_gather(const unsigned char *table, const int *ids, long nIds, int iters, unsigned *sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[S];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, chunk = lane & 7, rsub = lane >> 3;
  if (tid == 0) { for (int s = 0; s < S; s++) mbar_init(smem_u32(full + s), MODE >= 3 ? 1 : 128); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long base = ((long)blockIdx.x * 9973 * kRows) % (nIds - (long)kRows * (iters + 1));
  unsigned acc = 0;
  for (int it = 0; it < iters + S; it++) {
    const int s = it % S;
    if (it >= S) { // slot s was filled S iterations ago: wait for it, "consume" it
      if (MODE == 0 || MODE == 1) { asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory"); }
      else if (MODE >= 3) mbar_wait(smem_u32(full + s), ((it / S) - 1) & 1);
      acc += smem[s * kSlot + tid * 16];
      __syncthreads();
    }
    if (it < iters) {
      const int *my = ids + base + (long)it * kRows;
      const uint32_t sb = smem_u32(smem) + s * kSlot;
      if (MODE <= 2) {
        const int id = my[warp * 32 + lane]; // 32 rows per warp; 8 lanes copy one 128-byte half row per instruction
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int r = __shfl_sync(0xffffffffu, id, i * 4 + rsub), row = warp * 32 + i * 4 + rsub;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const unsigned char *src = table + (size_t)r * kRowBytes + h * 128 + chunk * 16;
            const uint32_t dst = sb + h * (kRows * 128) + row * 128 + ((chunk ^ (row & 7)) << 4);
            if (MODE == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            else if (MODE == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            else {
              uint4 v;
              asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src));
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
          }
        }
        if (MODE <= 1) asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
        if (tid == 0) mbar_expect(smem_u32(full + s), kSlot);
        __syncthreads();
        const int r = my[tid]; // one row per thread
        if (MODE == 3) { bulk_g2s(sb + tid * 128, table + (size_t)r * kRowBytes, 128, smem_u32(full + s)); bulk_g2s(sb + kRows * 128 + tid * 128, table + (size_t)r * kRowBytes + 128, 128, smem_u32(full + s)); }
        else bulk_g2s(sb + tid * 256, table + (size_t)r * kRowBytes, 256, smem_u32(full + s));
      }
    } else if (MODE <= 1) asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (acc == 0xdeadbeef) *sink = acc;
}
template <int MODE, int S>
static void run(const char *name, const unsigned char *table, const int *ids, long nIds, unsigned *sink, const char *where) {
  const int iters = 400, grid = 148 * 2;
  const size_t smem = (size_t)S * kSlot;
  CK(cudaFuncSetAttribute(k_gather<MODE, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e9f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a));
    k_gather<MODE, S><<<grid, 128, smem>>>(table, ids, nIds, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double bytes = (double)grid * iters * kSlot;
  printf("%-34s S=%d %-4s %7.3f ms  %6.2f TB/s  %5.1f B/clk/SM (at 1.965 GHz)\n", name, S, where, best, bytes / best / 1e9, bytes / (best * 1e-3) / 148 / 1.965e9);
}
int main() {
  unsigned *sink; CK(cudaMalloc(&sink, 4));
  for (int pass = 0; pass < 2; pass++) {
    const long nRows = pass == 0 ? 1171875 : 187500; // 300 MB / 48 MB of 256-byte rows
    unsigned char *table; CK(cudaMalloc(&table, nRows * kRowBytes)); CK(cudaMemset(table, 1, nRows * kRowBytes));
    const long nIds = 4 << 20;
    std::vector<int> h(nIds);
    // spatially coherent like the real gathers: runs of ~8 neighbouring rows at random bases
    unsigned long long x = 88172645463325252ull;
    for (long i = 0; i < nIds; i++) { if ((i & 7) == 0) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; } h[i] = (int)(((x % (nRows - 64)) + (i & 7) * 3) % nRows); }
    int *ids; CK(cudaMalloc(&ids, nIds * 4)); CK(cudaMemcpy(ids, h.data(), nIds * 4, cudaMemcpyHostToDevice));
    const char *where = pass == 0 ? "HBM" : "L2";
    run<0, 2>("cp.async.cg 16B", table, ids, nIds, sink, where);
    run<0, 3>("cp.async.cg 16B", table, ids, nIds, sink, where);
    run<1, 3>("cp.async.ca 16B", table, ids, nIds, sink, where);
    run<2, 3>("ld.global.nc + st.shared 16B", table, ids, nIds, sink, where);
    run<3, 3>("cp.async.bulk 128B/copy", table, ids, nIds, sink, where);
    run<4, 3>("cp.async.bulk 256B/copy", table, ids, nIds, sink, where);
    CK(cudaFree(table)); CK(cudaFree(ids));
  }
  return 0;
}
