// How fast can 148 persistent CTAs write GEMM output tiles?  (measurement tooling for the tcgen05 GEMM epilogue)
// Every CTA owns 128-row x 256-column tiles of a row-major [M, N] output and writes them with one of several store
// mechanisms; reported: GB/s of payload.  nvcc -arch=sm_100a -O3 -o scripts/store_bench.bin scripts/store_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: warp instruction = 4 rows x 128 B (16 B per lane)            [fp32-like, the staged epilogue]
// mode 1: warp instruction = 4 rows x 64 B (8 B per lane)              [bf16-like, the staged epilogue]
// mode 2: warp instruction = 32 rows x 16 B (row per lane)             [the old direct epilogue]
// mode 3: warp instruction = 1 row x 512 B (16 B per lane, fully linear inside the row)
// mode 4: like 0 with st.global.cs (streaming)   mode 5: like 0 with st.global.L1::no_allocate
template <int ES>   // element size 4 or 2
__global__ void __launch_bounds__(256) stg_kernel(uint8_t* out, int M, int N, int mode, int iters) {
  const int tiles_m = M / 128, tiles_n = N / 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = (size_t)N * ES;                         // bytes per row
  unsigned long long pol = 0;
  if (mode == 6) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (mode == 7) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  for (int it = 0; it < iters; ++it)
    for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
      const int m0 = (t % tiles_m) * 128, n0 = (t / tiles_m) * 256;
      uint8_t* base = out + (size_t)m0 * ld + (size_t)n0 * ES;
      const int q = warp & 3, half = warp >> 2;             // 32-row quarter, alternate 32-column chunks
      for (int c = half; c < 8; c += 2) {
        uint8_t* cb = base + (size_t)(q * 32) * ld + (size_t)c * 32 * ES;
        if (mode == 2) {
          uint8_t* p = cb + (size_t)lane * ld;
          for (int j = 0; j < 32 * ES; j += 16) *(uint4*)(p + j) = make_uint4(it, t, j, lane);
        } else if (mode == 3) {
          // 32 rows x (32*ES) bytes as linear rows: one row per iteration would be < 512 B; emulate with whole-tile rows
          for (int r = 0; r < 32; ++r) {
            uint8_t* p = cb + (size_t)r * ld;
            if (lane * 16 < 32 * ES) *(uint4*)(p + lane * 16) = make_uint4(it, t, r, lane);
          }
        } else {
          const int rg = lane >> 3, cg = lane & 7;
          for (int i = 0; i < 8; ++i) {
            uint8_t* p = cb + (size_t)(i * 4 + rg) * ld + cg * 4 * ES;
            if (ES == 4) {
              if (mode == 4) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(it), "r"(t), "r"(i), "r"(lane) : "memory");
              else if (mode == 5) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(it), "r"(t), "r"(i), "r"(lane) : "memory");
              else if (mode == 6 || mode == 7) asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(it), "r"(t), "r"(i), "r"(lane), "l"(pol) : "memory");
              else *(uint4*)p = make_uint4(it, t, i, lane);
            } else {
              *(uint2*)p = make_uint2(it + t, i + lane);
            }
          }
        }
      }
    }
}

// bulk stores: each epilogue warp fills a [32 rows][32*ES bytes] staging tile in shared memory, then lane 0 issues one
// cp.async.bulk.global.shared::cta per row (mode 10) -- or one 2-D TMA tensor store per chunk (mode 11)
template <int ES>
__global__ void __launch_bounds__(256) bulk_kernel(uint8_t* out, int M, int N, int mode, int iters,
                                                   const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(128) uint8_t stg_raw[];
  uint8_t (*stg)[2][32 * 128] = (uint8_t (*)[2][32 * 128])stg_raw;
  const int tiles_m = M / 128, tiles_n = N / 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = (size_t)N * ES;
  const int rowbytes = 32 * ES;
  int buf = 0;
  for (int it = 0; it < iters; ++it)
    for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
      const int m0 = (t % tiles_m) * 128, n0 = (t / tiles_m) * 256;
      const int q = warp & 3, half = warp >> 2;
      for (int c = half; c < 8; c += 2) {
        uint8_t* s = stg[warp][buf];
        // wait until the bulk group that last read this buffer is done (at most 1 other group in flight)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        for (int j = 0; j < rowbytes; j += 16) *(uint4*)(s + lane * rowbytes + j) = make_uint4(it, t, j, lane);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (mode == 10) {
            for (int r = 0; r < 32; ++r) {
              uint8_t* g = out + (size_t)(m0 + q * 32 + r) * ld + (size_t)(n0 + c * 32) * ES;
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s + r * rowbytes)), "r"(rowbytes) : "memory");
            }
          } else {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm),
                         "r"(smem_u32(s)), "r"(n0 + c * 32), "r"(m0 + q * 32) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
      }
    }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  const int shapes[3][2] = {{6656, 1024}, {6656, 4096}, {26624, 4096}};
  uint8_t* out;
  const size_t ARENA = (size_t)3 << 30;      // 3 GiB: every pass writes a region that left the L2 long ago
  CK(cudaMalloc(&out, ARENA));
  CK(cudaMemset(out, 0, ARENA));
  CK(cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(cudaFuncSetAttribute(bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(cudaFuncSetAttribute(stg_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(cudaFuncSetAttribute(stg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int smem_extra = 0; smem_extra <= 160 * 1024; smem_extra += 160 * 1024)
  for (int rotate = 1; rotate < 2; ++rotate)
  for (int si = 0; si < 2; ++si) {
    const int M = shapes[si][0], N = shapes[si][1];
    for (int es = 2; es <= 4; es += 2) {
      const double bytes = (double)M * N * es;
      const int modes[] = {0, 1, 2, 11};
      for (int mode : modes) {
        if ((mode == 1 && es == 4) || ((mode == 0 || mode == 4 || mode == 5 || mode == 6 || mode == 7) && es == 2)) continue;
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)N * es};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int iters = 1;
        const size_t region = ((size_t)bytes + 4095) / 4096 * 4096;
        const int nreg = (int)(ARENA / region);
        int pass = 0;
        std::vector<CUtensorMap> tms(nreg < 64 ? nreg : 64);
        for (size_t k = 0; k < tms.size(); ++k)
          enc(&tms[k], es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out + k * region, dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        auto launch = [&]() {
          const int k = rotate ? (pass++ % (int)tms.size()) : 0;
          uint8_t* o = out + (size_t)k * region;
          if (mode >= 10) {
            if (es == 4) bulk_kernel<4><<<148, 256, 65536 + smem_extra>>>(o, M, N, mode, iters, tms[k]);
            else bulk_kernel<2><<<148, 256, 65536 + smem_extra>>>(o, M, N, mode, iters, tms[k]);
          } else {
            if (es == 4) stg_kernel<4><<<148, 256, smem_extra>>>(o, M, N, mode, iters);
            else stg_kernel<2><<<148, 256, smem_extra>>>(o, M, N, mode, iters);
          }
        };
        launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int k = 0; k < 20; ++k) launch();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / (20 * iters);
        printf("smem+%3dK %s M=%d N=%d es=%d mode=%2d  %8.2f us per pass  %7.1f GB/s\n", smem_extra >> 10, rotate ? "L2-miss" : "L2-hit ", M, N, es, mode, us, bytes / us / 1e3);
      }
    }
  }
  return 0;
}
