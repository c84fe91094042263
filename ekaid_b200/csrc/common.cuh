// Shared helpers for the ekaid_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>

typedef __nv_bfloat16 bf16;
typedef __half f16;

// Error codes returned through the C ABI (0 = ok); see include/ekaid_b200.h
#define EK_OK 0
#define EK_ERR_SHAPE -1
#define EK_ERR_ALIGN -2
#define EK_ERR_ARCH -3
#define EK_ERR_CUDA -4
#define EK_ERR_UNSUPPORTED -5

void ek_set_error(const char* fmt, ...);

#define EK_CHECK_LAUNCH()                                                        \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ek_set_error("%s:%d launch failed: %s", __FILE__, __LINE__,                \
                   cudaGetErrorString(e__));                                     \
      return EK_ERR_CUDA;                                                        \
    }                                                                            \
  } while (0)

#define EK_REQUIRE(cond, code, ...)                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      ek_set_error(__VA_ARGS__);                                                 \
      return code;                                                               \
    }                                                                            \
  } while (0)

static inline int ek_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch.  Every kernel is launched with the programmatic-stream-serialization attribute and
// starts with ek_pdl_prologue() = griddepcontrol.wait, which blocks until the previous kernel of the stream has
// completed and flushed its memory.  No global memory is touched before the wait, so ordering is exactly stream
// order; what overlaps with the previous kernel's tail is launch latency and, for the GEMM, barrier / TMEM /
// tensor-map set-up.  Measured on the training step: 5.33 -> 5.15 ms.  An explicit early trigger
// (griddepcontrol.launch_dependents at kernel entry) was tried and rejected: the step got slower (5.73 ms) and fp32
// parity broke, so dependents are released by the implicit trigger at grid completion only.
// EKAID_B200_PDL=0 / ekaid_set_pdl(0) turn the attribute off (the wait becomes a no-op).
__device__ __forceinline__ void ek_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ek_pdl_prologue() { ek_pdl_wait(); }
int ek_pdl_enabled();
// EKAID_B200_MAX_CARVEOUT=1: on the first launch of every kernel ask for the maximum shared-memory carve-out (an SM
// cannot change its L1 / shared-memory split while CTAs are resident, so kernels with one common split can share SMs).
// Off by default: the small streaming kernels lose more from the smaller L1 than the overlap gains (4.22 vs 4.15 ms).
void ek_prepare_kernel(const void* fn);

template <typename... KArgs, typename... Args>
inline cudaError_t ek_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             Args&&... args) {
  ek_prepare_kernel((const void*)kernel);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = ek_pdl_enabled();
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` must hold >= 32 floats of shared memory; all threads must call
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  return r;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<f16>(f16 v) { return __half2float(v); }
// fp32 pair -> packed fp16x2, round to nearest, saturating to +-65504 instead of overflowing to infinity
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&t;
}
// 16-bit storage format chosen at run time: 0 = bf16, 1 = fp16 (saturating)
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, int fmt) { return fmt ? pack_f16x2_sat(lo, hi) : pack_bf16x2(lo, hi); }
__device__ __forceinline__ void ek_store16(bf16* dst, float v, int fmt) {
  *(unsigned short*)dst = (unsigned short)(pack16x2(v, 0.f, fmt) & 0xFFFFu);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ f16 from_f32<f16>(float v) {
  const unsigned short b = (unsigned short)(pack_f16x2_sat(v, 0.f) & 0xFFFFu);
  return *(const f16*)&b;
}
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// V consecutive elements <-> registers with explicit 8/16-byte accesses (pointers aligned to V * sizeof(T); V in {4, 8})
template <typename T, int V> __device__ __forceinline__ void store_vec(T* dst, const float* v);
template <> __device__ __forceinline__ void store_vec<float, 4>(float* dst, const float* v) {
  *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store_vec<float, 8>(float* dst, const float* v) {
  *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
  *(float4*)(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store_vec<bf16, 4>(bf16* dst, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 pk;
  pk.x = *(uint32_t*)&a; pk.y = *(uint32_t*)&b;
  *(uint2*)dst = pk;
}
template <> __device__ __forceinline__ void store_vec<bf16, 8>(bf16* dst, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *(uint32_t*)&a; pk.y = *(uint32_t*)&b; pk.z = *(uint32_t*)&c; pk.w = *(uint32_t*)&d;
  *(uint4*)dst = pk;
}
template <> __device__ __forceinline__ void store_vec<f16, 4>(f16* dst, const float* v) {
  uint2 pk;
  pk.x = pack_f16x2_sat(v[0], v[1]); pk.y = pack_f16x2_sat(v[2], v[3]);
  *(uint2*)dst = pk;
}
template <> __device__ __forceinline__ void store_vec<f16, 8>(f16* dst, const float* v) {
  uint4 pk;
  pk.x = pack_f16x2_sat(v[0], v[1]); pk.y = pack_f16x2_sat(v[2], v[3]);
  pk.z = pack_f16x2_sat(v[4], v[5]); pk.w = pack_f16x2_sat(v[6], v[7]);
  *(uint4*)dst = pk;
}
template <typename T, int V> __device__ __forceinline__ void load_vec(const T* src, float* v);
template <> __device__ __forceinline__ void load_vec<f16, 4>(const f16* src, float* v) {
  const uint2 raw = *(const uint2*)src;
  const float2 a = __half22float2(*(const __half2*)&raw.x), b = __half22float2(*(const __half2*)&raw.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> __device__ __forceinline__ void load_vec<float, 4>(const float* src, float* v) {
  const float4 a = *(const float4*)src;
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load_vec<bf16, 4>(const bf16* src, float* v) {
  const uint2 raw = *(const uint2*)src;
  const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&raw.x), b = __bfloat1622float2(*(const __nv_bfloat162*)&raw.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Counter-based dropout RNG (splitmix64 finaliser): the mask of element `idx` at dropout site `site` for the step
// whose seed lives in device memory (CUDA-graph safe) -- regenerated, never stored, identical in forward and backward.
struct EkDrop {
  const unsigned long long* seed;   // device pointer, nullptr = dropout off
  unsigned int site;
  float p;
};
__device__ __forceinline__ unsigned int ek_rand32(unsigned long long seed, unsigned int site, unsigned long long idx) {
  unsigned long long z = seed + (unsigned long long)site * 0x9E3779B97F4A7C15ull + idx * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 32);
}
// One 64-bit draw decides four consecutive elements (16 bits each, p resolved to 2^-16): element idx of a site uses
// lane idx & 3 of the draw of group idx >> 2.  Kernels that own aligned runs of elements hash once per group.
__device__ __forceinline__ unsigned long long ek_draw64(unsigned long long seedv, unsigned int site, unsigned long long grp) {
  unsigned long long z = seedv + (unsigned long long)site * 0x9E3779B97F4A7C15ull + grp * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// multiplier of element idx: 0 (dropped) or 1/(1-p) (kept); 1 when dropout is off
__device__ __forceinline__ float ek_drop_mult(const EkDrop& d, unsigned long long seedv, unsigned long long idx) {
  if (d.seed == nullptr || d.p <= 0.f) return 1.f;
  const unsigned int thr = (unsigned int)(d.p * 65536.0f);
  const unsigned long long z = ek_draw64(seedv, d.site, idx >> 2);
  return ((unsigned int)(z >> (16 * (idx & 3))) & 0xFFFFu) >= thr ? 1.f / (1.f - d.p) : 0.f;
}
// multipliers of the V consecutive elements e0 .. e0+V-1 (V = 2: e0 even; V % 4 == 0: one draw per aligned group of four)
template <int V>
__device__ __forceinline__ void ek_drop_multv(const EkDrop& d, unsigned long long seedv, unsigned long long e0, float* out) {
  if (d.seed == nullptr || d.p <= 0.f) {
#pragma unroll
    for (int k = 0; k < V; ++k) out[k] = 1.f;
    return;
  }
  const unsigned int thr = (unsigned int)(d.p * 65536.0f);
  const float keep = 1.f / (1.f - d.p);
  if (V == 2 && (e0 & 1) == 0) {
    const unsigned long long z = ek_draw64(seedv, d.site, e0 >> 2) >> (16 * (e0 & 3));
    out[0] = ((unsigned int)z & 0xFFFFu) >= thr ? keep : 0.f;
    out[1] = ((unsigned int)(z >> 16) & 0xFFFFu) >= thr ? keep : 0.f;
  } else if (V % 4 == 0 && (e0 & 3) == 0) {
#pragma unroll
    for (int g = 0; g < V / 4; ++g) {
      const unsigned long long z = ek_draw64(seedv, d.site, (e0 >> 2) + g);
#pragma unroll
      for (int k = 0; k < 4; ++k) out[4 * g + k] = ((unsigned int)(z >> (16 * k)) & 0xFFFFu) >= thr ? keep : 0.f;
    }
  } else {
#pragma unroll
    for (int k = 0; k < V; ++k) out[k] = ek_drop_mult(d, seedv, e0 + k);
  }
}
__device__ __forceinline__ unsigned long long ek_seed(const EkDrop& d) { return d.seed ? *d.seed : 0ull; }
// four independent 16-bit lanes from one 64-bit draw: multipliers of elements 4*group .. 4*group+3 (p resolved to 2^-16)
__device__ __forceinline__ void ek_drop_mult4(const EkDrop& d, unsigned long long seedv, unsigned long long group,
                                              float out[4]) {
  if (d.seed == nullptr || d.p <= 0.f) { out[0] = out[1] = out[2] = out[3] = 1.f; return; }
  unsigned long long z = seedv + (unsigned long long)d.site * 0x9E3779B97F4A7C15ull + group * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const unsigned int thr = (unsigned int)(d.p * 65536.0f);
  const float keep = 1.f / (1.f - d.p);
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = ((unsigned int)(z >> (16 * k)) & 0xFFFFu) >= thr ? keep : 0.f;
}
