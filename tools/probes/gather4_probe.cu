// Developer probe: does cp.async.bulk.tensor ... tile::gather4 deliver swizzled rows, what does an
// out-of-range row give, and how fast is it compared with 16-byte cp.async gathers?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_probe gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *tm, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// ---- correctness: one CTA gathers 128 rows x 32 floats (one 16 KB swizzle atom) and dumps smem
__global__ void k_check(const __grid_constant__ CUtensorMap tm, const int *ids, float *dump, int col) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = -7.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) mbar_expect(smem_u32(&bar), 128 * 128);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    gather4(smem_u32(smem) + l * 512, &tm, col, ids[l * 4], ids[l * 4 + 1], ids[l * 4 + 2], ids[l * 4 + 3], smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) dump[i] = reinterpret_cast<float *>(smem)[i];
}

// ---- throughput: persistent CTAs, S stages of 16 KB, one warp issues gather4, one thread recycles
template <int S>
__global__ void __launch_bounds__(64) k_bw_tma(const __grid_constant__ CUtensorMap tm, const int *ids, long nIds, int iters, long long *cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[S];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < S; s++) mbar_init(smem_u32(full + s), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  long long t0 = clock64();
  long base = ((long)blockIdx.x * 9973 * 128) % (nIds - 128);
  for (int it = 0; it < iters + S; it++) {
    const int s = it % S;
    if (it >= S) mbar_wait(smem_u32(full + s), ((it / S) - 1) & 1);
    if (it < iters) {
      if (lane == 0) mbar_expect(smem_u32(full + s), 128 * 128);
      __syncwarp();
      const int *p = ids + base + lane * 4;
      gather4(smem_u32(smem) + s * 16384 + lane * 512, &tm, (it & 3) * 32, p[0], p[1], p[2], p[3], smem_u32(full + s));
      if ((it & 3) == 3) base = (base + 128 * 148) % (nIds - 128);
    }
  }
  if (lane == 0) cyc[blockIdx.x] = clock64() - t0;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t n) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory"); }
template <int S>
__global__ void __launch_bounds__(128) k_bw_cpasync(const float *in, const int *ids, long nIds, int iters, long long *cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, chunk = tid & 7, rg = tid >> 3;
  long long t0 = clock64();
  long base = ((long)blockIdx.x * 9973 * 128) % (nIds - 128);
  for (int it = 0; it < iters; it++) {
    const int s = it % S;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int row = rg * 8 + i;
      const int id = ids[base + row];
      cp_async16(smem_u32(smem) + s * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4), in + (size_t)(id >= 0 ? id : 0) * 128 + (it & 3) * 32 + chunk * 4, id >= 0 ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
    if ((it & 3) == 3) base = (base + 128 * 148) % (nIds - 128);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main(int argc, char **argv) {
  const long R = argc > 1 ? atol(argv[1]) : 1155656;
  const int C = 128;
  float *d;
  CK(cudaMalloc(&d, (size_t)R * C * 4));
  std::vector<float> h((size_t)R * C);
  for (long r = 0; r < R; r++) for (int c = 0; c < C; c++) h[r * C + c] = (float)(r % 65536) + c / 1024.f;
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qr));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

  for (int boxRows : {1, 4}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)boxRows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("boxRows=%d encode -> %d\n", boxRows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    // correctness
    std::vector<int> ids(128);
    for (int i = 0; i < 128; i++) ids[i] = (int)((i * 7919L + 13) % R);
    ids[5] = -1; ids[6] = (int)R; ids[7] = (int)R + 100; ids[9] = 0x7fffffff;
    int *dids; float *ddump;
    CK(cudaMalloc(&dids, 512)); CK(cudaMalloc(&ddump, 16384));
    CK(cudaMemcpy(dids, ids.data(), 512, cudaMemcpyHostToDevice));
    const int col = 64;
    k_check<<<1, 128, 16384>>>(tm, dids, ddump, col);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k_check failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> dump(4096);
    CK(cudaMemcpy(dump.data(), ddump, 16384, cudaMemcpyDeviceToHost));
    int bad = 0, zeroRows = 0;
    for (int row = 0; row < 128; row++) {
      bool oob = ids[row] < 0 || ids[row] >= R;
      for (int ch = 0; ch < 8; ch++) for (int w = 0; w < 4; w++) {
        float got = dump[row * 32 + ((ch ^ (row & 7)) << 2) + w];
        float want = oob ? 0.f : (float)(ids[row] % 65536) + (col + ch * 4 + w) / 1024.f;
        if (got != want) { if (bad < 8) printf("  row %d (id %d) ch %d w %d: got %g want %g\n", row, ids[row], ch, w, got, want); bad++; }
      }
      zeroRows += oob;
    }
    printf("boxRows=%d check: %d mismatches (%d out-of-range rows expected zero)\n", boxRows, bad, zeroRows);
    if (bad) continue;

    // throughput
    const long nIds = 1 << 22;
    std::vector<int> hid(nIds);
    uint64_t st = 88172645463325252ull;
    for (long i = 0; i < nIds; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; hid[i] = (int)(st % (uint64_t)R); }
    int *dbig; long long *dcyc;
    CK(cudaMalloc(&dbig, nIds * 4)); CK(cudaMalloc(&dcyc, 148 * 8));
    CK(cudaMemcpy(dbig, hid.data(), nIds * 4, cudaMemcpyHostToDevice));
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto launch) {
      launch(); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double bytes = 148.0 * iters * 16384;
      printf("  %-28s %.3f ms  %.1f GB/s  (%.1f B/clk/SM at 1.9 GHz)\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.9);
    };
    CK(cudaFuncSetAttribute(k_bw_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_bw_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_bw_tma<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_bw_cpasync<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_bw_cpasync<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    run("tma gather4, 4 stages", [&] { k_bw_tma<4><<<148, 64, 4 * 16384>>>(tm, dbig, nIds, iters, dcyc); });
    run("tma gather4, 8 stages", [&] { k_bw_tma<8><<<148, 64, 8 * 16384>>>(tm, dbig, nIds, iters, dcyc); });
    run("tma gather4, 12 stages", [&] { k_bw_tma<12><<<148, 64, 12 * 16384>>>(tm, dbig, nIds, iters, dcyc); });
    run("cp.async16 128 thr, 4 stages", [&] { k_bw_cpasync<4><<<148, 128, 4 * 16384>>>(d, dbig, nIds, iters, dcyc); });
    run("cp.async16 128 thr, 8 stages", [&] { k_bw_cpasync<8><<<148, 128, 8 * 16384>>>(d, dbig, nIds, iters, dcyc); });
    cudaFree(dbig); cudaFree(dcyc); cudaFree(dids); cudaFree(ddump);
  }
  return 0;
}
