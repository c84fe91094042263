// Element-wise / reduction kernels of the fusion half of the path and shared utilities
// (reference models/modules.py:233-310, models/relation_encoder.py:19-29, utils/mimic_utils.py:119-149).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- casts / copies
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long long lds, bf16* __restrict__ dst,
                                     long long ldd, long long rows, int cols, int fmt) {
  ek_pdl_prologue();
  const long long total = rows * (cols / 4);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / (cols / 4);
    const int c = (int)(e % (cols / 4)) * 4;
    const float4 v = *(const float4*)(src + r * lds + c);
    uint2 pk;
    pk.x = pack16x2(v.x, v.y, fmt);
    pk.y = pack16x2(v.z, v.w, fmt);
    *(uint2*)(dst + r * ldd + c) = pk;
  }
}
// test hook: the dropout multipliers (0 or 1 / (1 - p)) of elements [0, n) at one site, exactly as every kernel of the
// library derives them from (seed, site, element index) -- lets a test hand the SAME masks to the oracle
__global__ void drop_mask_kernel(EkDrop dr, long long n, float* __restrict__ out) {
  ek_pdl_prologue();
  const unsigned long long seedv = ek_seed(dr);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    out[e] = ek_drop_mult(dr, seedv, (unsigned long long)e);
}

// fp32 operand -> three bf16 planes for the split-precision tensor-core product of the fp32 parity path:
//   x = hi + lo + eps,  hi = bf16(x), lo = bf16(x - hi), |eps| <= 2^-17 |x|
//   A B^T ~= Ah Bh^T + Al Bh^T + Ah Bl^T      (the dropped Al Bl^T term is <= 2^-16 of the product)
// which is ONE bf16 GEMM over a contraction axis three times as long: A' = [Al | Ah | Ah], B' = [Bh | Bl | Bh].
// The two small products come FIRST: the tensor core's accumulator adds with truncation, an error proportional to the
// running sum, so the correction terms are accumulated while the sum is still 2^-9 of its final size.
// pattern 0 = (lo, hi, hi) for the A operand, 1 = (hi, lo, hi) for B.  along_rows = 0: the contraction axis is the
// column axis of src [rows, cols] -> dst [rows, 3 cols]; 1: it is the row axis -> dst [3 rows, cols].
__global__ void split3_bf16_kernel(const float* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd,
                                   long long rows, int cols, int pattern, int along_rows) {
  ek_pdl_prologue();
  const long long total = rows * (cols / 4);
  const long long plane = along_rows ? rows * ldd : (long long)cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / (cols / 4);
    const int c = (int)(e % (cols / 4)) * 4;
    const float4 v = *(const float4*)(src + r * lds + c);
    const float x[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = __bfloat162float(__float2bfloat16_rn(x[i]));
      l[i] = x[i] - h[i];                       // exact in fp32
    }
    uint2 ph, pl;
    ph.x = pack16x2(h[0], h[1], 0); ph.y = pack16x2(h[2], h[3], 0);
    pl.x = pack16x2(l[0], l[1], 0); pl.y = pack16x2(l[2], l[3], 0);
    bf16* d = dst + r * ldd + c;
    *(uint2*)d = pattern ? ph : pl;
    *(uint2*)(d + plane) = pattern ? pl : ph;
    *(uint2*)(d + 2 * plane) = ph;
  }
}
// Up to CM_MAX strided 2-D blocks in one launch (blockIdx.y = block): fp32 -> bf16 casts (mode 0), fp32 -> fp32 copies
// (mode 1), raw 16-byte copies of `cols` BYTES per row (mode 2: the captured step's input staging), fp32 -> fp16 casts
// (mode 3, saturating) or fp16 -> bf16 conversions (mode 4: the backward's copy of a forward activation).
constexpr int CM_MAX = 16;
struct CastMany {
  const void* src[CM_MAX];
  void* dst[CM_MAX];
  long long lds[CM_MAX], ldd[CM_MAX], rows[CM_MAX];
  int cols[CM_MAX];
  int mode[CM_MAX];
};
__global__ void cast_many_kernel(CastMany t) {
  ek_pdl_prologue();
  const int i = blockIdx.y;
  const long long rows = t.rows[i];
  const int cols = t.cols[i], mode = t.mode[i];
  if (mode == 2) {
    const int c16 = cols / 16;
    const long long total = rows * c16;
    const uint8_t* src = (const uint8_t*)t.src[i];
    uint8_t* dst = (uint8_t*)t.dst[i];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      const long long r = e / c16;
      const int c = (int)(e % c16) * 16;
      *(uint4*)(dst + r * t.ldd[i] + c) = *(const uint4*)(src + r * t.lds[i] + c);
    }
    return;
  }
  if (mode == 4) {
    const int c8n = cols / 8;
    const long long total = rows * c8n;
    const f16* src = (const f16*)t.src[i];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      const long long r = e / c8n;
      const int c = (int)(e % c8n) * 8;
      const uint4 raw = *(const uint4*)(src + r * t.lds[i] + c);
      const float2 a = __half22float2(*(const __half2*)&raw.x), b = __half22float2(*(const __half2*)&raw.y);
      const float2 cc = __half22float2(*(const __half2*)&raw.z), d = __half22float2(*(const __half2*)&raw.w);
      uint4 pk;
      pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(b.x, b.y); pk.z = pack_bf16x2(cc.x, cc.y); pk.w = pack_bf16x2(d.x, d.y);
      *(uint4*)((bf16*)t.dst[i] + r * t.ldd[i] + c) = pk;
    }
    return;
  }
  const int c4n = cols / 4;
  const long long total = rows * c4n;
  const float* src = (const float*)t.src[i];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / c4n;
    const int c = (int)(e % c4n) * 4;
    const float4 v = *(const float4*)(src + r * t.lds[i] + c);
    if (mode == 0 || mode == 3) {
      uint2 pk;
      pk.x = pack16x2(v.x, v.y, mode == 3);
      pk.y = pack16x2(v.z, v.w, mode == 3);
      *(uint2*)((bf16*)t.dst[i] + r * t.ldd[i] + c) = pk;
    } else {
      *(float4*)((float*)t.dst[i] + r * t.ldd[i] + c) = v;
    }
  }
}
__global__ void copy_f32_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                                long long rows, int cols) {
  ek_pdl_prologue();
  const long long total = rows * (cols / 4);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / (cols / 4);
    const int c = (int)(e % (cols / 4)) * 4;
    *(float4*)(dst + r * ldd + c) = *(const float4*)(src + r * lds + c);
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, long long lds, float* __restrict__ dst,
                                     long long ldd, long long rows, int cols, float scale) {
  ek_pdl_prologue();
  const long long total = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cols;
    const int c = (int)(e % cols);
    dst[r * ldd + c] = scale * __bfloat162float(src[r * lds + c]);
  }
}

// ---------------------------------------------------------------- column sums (bias gradients), deterministic, one launch
// part[p, n] = sum_{m in chunk p} scale[m] * src[m, n]; the last CTA of a column block to finish (ticket counter) adds the
// partials in fixed order p = 0..nparts-1 and leaves the counter at zero for the next call.
// Block = 8 warps; a warp reads one row segment of VEC*32 consecutive columns with 16-byte loads (VEC = 8 bf16 / 4 fp32),
// the 8 warps take 8 different rows per iteration.  grid = (column blocks, row chunks).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ src, long long ld, long long M, int N, const float* __restrict__ rowscale,
              float* part, int nparts, unsigned int* tickets, float* __restrict__ out, int vec_ok) {
  ek_pdl_prologue();
  constexpr int CB = VEC * 32;             // columns per block
  __shared__ float red[8][CB + 1];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n0 = blockIdx.x * CB + lane * VEC;
  const long long rows_per = (M + nparts - 1) / nparts;
  const long long r0 = (long long)blockIdx.y * rows_per;
  const long long r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
  float s[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) s[v] = 0.f;
  if (vec_ok && n0 + VEC <= N) {
#pragma unroll 4
    for (long long m = r0 + ty; m < r1; m += 8) {
      const float sc = rowscale ? rowscale[m] : 1.f;
      const uint4 raw = *(const uint4*)(src + m * ld + n0);
      const T* pv = (const T*)&raw;
#pragma unroll
      for (int v = 0; v < VEC; ++v) s[v] = fmaf(to_f32<T>(pv[v]), sc, s[v]);
    }
  } else {
    for (long long m = r0 + ty; m < r1; m += 8) {
      const float sc = rowscale ? rowscale[m] : 1.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (n0 + v < N) s[v] = fmaf(to_f32<T>(src[m * ld + n0 + v]), sc, s[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) red[ty][lane * VEC + v] = s[v];
  __syncthreads();
  for (int c = threadIdx.x; c < CB; c += 256) {
    const int n = blockIdx.x * CB + c;
    if (n < N) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][c];
      part[(size_t)blockIdx.y * N + n] = t;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(tickets + blockIdx.x, 1u);
    is_last = (ticket == (unsigned int)nparts - 1u);
    if (is_last) tickets[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (is_last) {
    // final pass, all 8 warps: warp ty adds partials ty, ty+8, ... for its lane's VEC columns (independent loads, one
    // L2 round trip), then the 8 sub-sums are added in fixed order through shared memory -> still deterministic
    __threadfence();
    float t[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) t[v] = 0.f;
#pragma unroll 8
    for (int p = ty; p < nparts; p += 8) {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (n0 + v < N) t[v] += __ldcg(part + (size_t)p * N + n0 + v);
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[ty][lane * VEC + v] = t[v];
    __syncthreads();
    for (int c = threadIdx.x; c < CB; c += 256) {
      const int n = blockIdx.x * CB + c;
      if (n < N) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += red[k][c];
        out[n] = a;
      }
    }
  }
}

// Several column sums over the same number of rows in ONE launch (blockIdx.z = job): the bias gradients of a relation
// layer (query, key, self_weights, out-projection).  Same scheme as colsum_kernel with 128 columns per block for both
// operand types; job j uses tickets [j*1024, (j+1)*1024) and partials [(CSM_MAX + j*64*... see launcher].
constexpr int CSM_MAX = 8;
struct ColsumMany {
  const void* src[CSM_MAX];
  float* out[CSM_MAX];
  long long ld[CSM_MAX];
  int N[CSM_MAX];
  int is_bf16[CSM_MAX];
  int vec_ok[CSM_MAX];
  long long part_off[CSM_MAX];     // float offset of the job's partial sums in the workspace
};
__global__ void __launch_bounds__(256)
colsum_many_kernel(ColsumMany t, long long M, int nparts, float* workspace) {
  ek_pdl_prologue();
  constexpr int VEC = 4, CB = 128;
  __shared__ float red[8][CB + 1];
  __shared__ int is_last;
  const int j = blockIdx.z;
  const int N = t.N[j];
  if (blockIdx.x * CB >= N) return;                    // this job has fewer column blocks than the widest one
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n0 = blockIdx.x * CB + lane * VEC;
  const long long ld = t.ld[j];
  unsigned int* tickets = (unsigned int*)workspace + (size_t)j * 1024;
  float* part = workspace + t.part_off[j];
  const long long rows_per = (M + nparts - 1) / nparts;
  const long long r0 = (long long)blockIdx.y * rows_per;
  const long long r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
  float s[VEC] = {0.f, 0.f, 0.f, 0.f};
  if (t.is_bf16[j]) {
    const bf16* src = (const bf16*)t.src[j];
    if (t.vec_ok[j] && n0 + VEC <= N) {
#pragma unroll 4
      for (long long m = r0 + ty; m < r1; m += 8) {
        float x[4];
        load_vec<bf16, 4>(src + m * ld + n0, x);
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] += x[v];
      }
    } else {
      for (long long m = r0 + ty; m < r1; m += 8)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (n0 + v < N) s[v] += __bfloat162float(src[m * ld + n0 + v]);
    }
  } else {
    const float* src = (const float*)t.src[j];
    if (t.vec_ok[j] && n0 + VEC <= N) {
#pragma unroll 4
      for (long long m = r0 + ty; m < r1; m += 8) {
        const float4 x = *(const float4*)(src + m * ld + n0);
        s[0] += x.x; s[1] += x.y; s[2] += x.z; s[3] += x.w;
      }
    } else {
      for (long long m = r0 + ty; m < r1; m += 8)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (n0 + v < N) s[v] += src[m * ld + n0 + v];
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) red[ty][lane * VEC + v] = s[v];
  __syncthreads();
  for (int c = threadIdx.x; c < CB; c += 256) {
    const int n = blockIdx.x * CB + c;
    if (n < N) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[k][c];
      part[(size_t)blockIdx.y * N + n] = a;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(tickets + blockIdx.x, 1u);
    is_last = (ticket == (unsigned int)nparts - 1u);
    if (is_last) tickets[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float a[VEC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int p = ty; p < nparts; p += 8)
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (n0 + v < N) a[v] += __ldcg(part + (size_t)p * N + n0 + v);
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[ty][lane * VEC + v] = a[v];
    __syncthreads();
    for (int c = threadIdx.x; c < CB; c += 256) {
      const int n = blockIdx.x * CB + c;
      if (n < N) {
        float b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) b += red[k][c];
        t.out[j][n] = b;
      }
    }
  }
}

// ---------------------------------------------------------------- row flags for quirk Q10
__global__ void row_zero_flags_kernel(const float* __restrict__ X, long long M, int D, uint8_t* __restrict__ flags) {
  ek_pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += X[row * D + c];
  s = warp_sum(s);
  if (lane == 0) flags[row] = (s == 0.f) ? 1 : 0;
}

// ---------------------------------------------------------------- per-sample row sums (backward of the q broadcast)
// out[b, c] = sum_{s < S} sum_{n < N} (flags[row] ? 0 : src[row, c]),  row = (s*B + b)*N + n
template <typename T>
__global__ void __launch_bounds__(256)
group_rowsum_kernel(const T* __restrict__ src, long long ld, int N, int B, int S, int D,
                    const uint8_t* __restrict__ flags, float* __restrict__ out) {
  ek_pdl_prologue();
  // 64 columns x 4 row lanes per CTA: every lane sums every fourth row of the sample (independent loads in flight
  // instead of one serial chain of S*N dependent ones), fixed-order combine -> deterministic
  __shared__ float part[4][64];
  const int b = blockIdx.x;
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  const int R = S * N;
  float acc = 0.f;
  if (c < D) {
#pragma unroll 4
    for (int i = rl; i < R; i += 4) {
      const int s_ = i / N, n = i - s_ * N;
      const long long row = ((long long)s_ * B + b) * N + n;
      if (!(flags && flags[row])) acc += to_f32<T>(src[row * ld + c]);
    }
  }
  part[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && c < D) out[(size_t)b * D + c] = ((part[0][cl] + part[1][cl]) + part[2][cl]) + part[3][cl];
}

// ---------------------------------------------------------------- graph combine + difference (modules.py:233-250)
// X3: [2*BN, D] (bef rows, then aft rows).  mode 0: identity, 1: 'all' (three fp32 products, quirk Q1),
// 2: 'i+s' ((x+x)/2).  Outputs Xc fp32 [2BN, D] and the concat buffer CAT[2BN, 3D]:
// [:, 0:D] = Xc, [:, D:2D] = diff (same diff for the bef and the aft row of a pair).
template <typename T>
__global__ void combine_diff_fwd_kernel(const float* __restrict__ X3, long long BN, int D, int mode, float c1, float c2,
                                        float c3, float* __restrict__ Xc, T* __restrict__ CAT) {
  ek_pdl_prologue();
  const long long total = BN * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / D;
    const int c = (int)(e % D);
    float xb = X3[e], xa = X3[total + e];
    if (mode == 1) {
      xb = c1 * xb + c2 * xb + c3 * xb;
      xa = c1 * xa + c2 * xa + c3 * xa;
    } else if (mode == 2) {
      xb = (xb + xb) / 2.f;
      xa = (xa + xa) / 2.f;
    }
    const float d = xa - xb;
    Xc[e] = xb;
    Xc[total + e] = xa;
    CAT[r * 3 * D + c] = from_f32<T>(xb);
    CAT[(BN + r) * 3 * D + c] = from_f32<T>(xa);
    CAT[r * 3 * D + D + c] = from_f32<T>(d);
    CAT[(BN + r) * 3 * D + D + c] = from_f32<T>(d);
  }
}
// dX3 = csum * (dXc_direct + dCAT[:, 0:D] -/+ (dCAT_bef[:, D:2D] + dCAT_aft[:, D:2D]))
__global__ void combine_diff_bwd_kernel(const float* __restrict__ dXc, const float* __restrict__ dCAT, long long BN,
                                        int D, int mode, float c1, float c2, float c3, float* __restrict__ dX3) {
  ek_pdl_prologue();
  const long long total = BN * D;
  float cs = 1.f;
  if (mode == 1) cs = c1 + c2 + c3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / D;
    const int c = (int)(e % D);
    const float dd = dCAT[r * 3 * D + D + c] + dCAT[(BN + r) * 3 * D + D + c];
    const float gb = dXc[e] + dCAT[r * 3 * D + c] - dd;
    const float ga = dXc[total + e] + dCAT[(BN + r) * 3 * D + c] + dd;
    if (mode == 1) {
      dX3[e] = c1 * gb + c2 * gb + c3 * gb;
      dX3[total + e] = c1 * ga + c2 * ga + c3 * ga;
    } else {
      dX3[e] = cs * gb;
      dX3[total + e] = cs * ga;
    }
  }
}

// ---------------------------------------------------------------- gated fusion (modules.py:278-288)
// pre [M, 2D] fp32: [:, 0:D] = context pre-activation, [:, D:2D] = gate pre-activation (biases included)
template <typename T>
__global__ void gate_fwd_kernel(const float* __restrict__ pre, long long M, int D, T* __restrict__ ctx,
                                T* __restrict__ gate, T* __restrict__ CAT, EkDrop dc, EkDrop dg) {
  ek_pdl_prologue();
  // four consecutive columns per thread (D % 4 == 0, checked by the launcher)
  const long long total = M * D / 4;
  const unsigned long long sc = ek_seed(dc), sg = ek_seed(dg);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t * 4;
    const long long r = e / D;
    const int c = (int)(e % D);
    const float4 pc = *(const float4*)(pre + r * 2 * D + c);
    const float4 pg = *(const float4*)(pre + r * 2 * D + D + c);
    const float cv[4] = {tanhf(pc.x), tanhf(pc.y), tanhf(pc.z), tanhf(pc.w)};
    const float gv[4] = {sigmoidf_(pg.x), sigmoidf_(pg.y), sigmoidf_(pg.z), sigmoidf_(pg.w)};
    // train mode: Dropout(0.5) on the tanh output and, independently, on the sigmoid output (modules.py:279,281)
    float mc[4], mg[4];
    ek_drop_multv<4>(dc, sc, (unsigned long long)e, mc);
    ek_drop_multv<4>(dg, sg, (unsigned long long)e, mg);
    float xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = (gv[k] * mg[k]) * (cv[k] * mc[k]);
    store_vec<T, 4>(ctx + e, cv);
    store_vec<T, 4>(gate + e, gv);
    store_vec<T, 4>(CAT + r * 3 * D + 2 * D + c, xs);
  }
}
// dXs = dCAT[:, 2D:3D] (fp32) -> dpre [M, 2D] (T)
template <typename T>
__global__ void gate_bwd_kernel(const float* __restrict__ dCAT, const T* __restrict__ ctx, const T* __restrict__ gate,
                                long long M, int D, T* __restrict__ dpre, EkDrop dc, EkDrop dg) {
  ek_pdl_prologue();
  const long long total = M * D / 4;
  const unsigned long long sc = ek_seed(dc), sg = ek_seed(dg);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t * 4;
    const long long r = e / D;
    const int c = (int)(e % D);
    const float4 dv = *(const float4*)(dCAT + r * 3 * D + 2 * D + c);
    const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
    float mc[4], mg[4];
    ek_drop_multv<4>(dc, sc, (unsigned long long)e, mc);
    ek_drop_multv<4>(dg, sg, (unsigned long long)e, mg);
    float cv[4], gv[4], dc4[4], dg4[4];
    load_vec<T, 4>(ctx + e, cv);
    load_vec<T, 4>(gate + e, gv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = dd[k] * mc[k] * mg[k];
      dc4[k] = d * gv[k] * (1.f - cv[k] * cv[k]);
      dg4[k] = d * cv[k] * gv[k] * (1.f - gv[k]);
    }
    store_vec<T, 4>(dpre + r * 2 * D + c, dc4);
    store_vec<T, 4>(dpre + r * 2 * D + D + c, dg4);
  }
}

// ---------------------------------------------------------------- attention pooling (modules.py:302-308)
// att[row] = sigmoid(e[row,:] . w + b)   one warp per row
__global__ void att_score_kernel(const float* __restrict__ E, long long M, int dim, const float* __restrict__ w,
                                 const float* __restrict__ b, float* __restrict__ att) {
  ek_pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) s = fmaf(E[row * dim + c], w[c], s);
  s = warp_sum(s);
  if (lane == 0) att[row] = sigmoidf_(s + b[0]);
}
// attended[g, c] = sum_n att[g*N + n] * Xc[g*N + n, c]
__global__ void att_pool_kernel(const float* __restrict__ att, const float* __restrict__ Xc, int N, int D,
                                float* __restrict__ attended) {
  ek_pdl_prologue();
  const int g = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s = fmaf(att[(size_t)g * N + n], Xc[((size_t)g * N + n) * D + c], s);
  attended[(size_t)g * D + c] = s;
}
// backward, one warp per row:  d_att = Xc[row,:] . dA[g,:] (+ d_attw[row]);  dpre = d_att * att * (1 - att)
//   dXc[row,:] = att[row] * dA[g,:] ; dE[row,:] = dpre * w * (e > 0)   (T) ; dpre_out[row] = dpre
template <typename T>
__global__ void att_pool_bwd_kernel(const float* __restrict__ dA, const float* __restrict__ dattw,
                                    const float* __restrict__ att, const float* __restrict__ Xc,
                                    const float* __restrict__ E, const float* __restrict__ w, long long M, int N, int D,
                                    int dim, float* __restrict__ dXc, T* __restrict__ dE, float* __restrict__ dpre_out,
                                    float escale) {
  ek_pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const long long g = row / N;
  const float a = att[row];
  float s = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float da = dA[g * D + c];
    s = fmaf(Xc[row * D + c], da, s);
    dXc[row * D + c] = a * da;
  }
  s = warp_sum(s);
  if (dattw) s += dattw[row];
  const float dp = s * a * (1.f - a);
  if (lane == 0) dpre_out[row] = dp;
  // escale = 1/(1-p) of the embed Dropout(0.5) in train mode (E > 0 implies the element was kept)
  for (int c = lane; c < dim; c += 32)
    dE[row * dim + c] = from_f32<T>(E[row * dim + c] > 0.f ? dp * w[c] * escale : 0.f);
}

// ---------------------------------------------------------------- process_matrix (mimic_utils.py:119-149)
// labels: float64 [B, S, S]; out fp32 [B, N, N, L]: plane c = (label == c+1)
__global__ void onehot_adj_kernel(const double* __restrict__ labels, int S, int N, int L, long long total,
                                  float* __restrict__ out) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / ((long long)N * N);
    const int ij = (int)(e % ((long long)N * N));
    const int i = ij / N, j = ij % N;
    const double v = labels[(b * S + i) * S + j];
    float* o = out + e * L;
    for (int c = 0; c < L; ++c) o[c] = (v == (double)(c + 1)) ? 1.f : 0.f;
  }
}

// the same from the loader's compact int8 label matrices
__global__ void onehot_adj_i8_kernel(const int8_t* __restrict__ labels, int S, int N, int L, long long total,
                                     float* __restrict__ out) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / ((long long)N * N);
    const int ij = (int)(e % ((long long)N * N));
    const int i = ij / N, j = ij % N;
    const int v = labels[(b * S + i) * S + j];
    float* o = out + e * L;
    for (int c = 0; c < L; ++c) o[c] = (v == c + 1) ? 1.f : 0.f;
  }
}

// ---------------------------------------------------------------- spatial adjacency labels from ROI boxes
// bbox_relation_type / reverse_type / get_adj_matrix ("feature extraction/ana_bbox_generator.py":213-259,266-302,
// 320-335): boxes f64 [B, N, 4] (xmin, ymin, xmax, ymax) -> labels f64 [B, S, S] (the HDF5 `image_adj_matrix` layout the
// loader hands to process_matrix; rows / columns >= N are 0).  The reference evaluates the pair (i, j) for i <= j and
// stores reverse_type at (j, i); one thread per entry does the same on its own pair, in double like the reference.
__device__ __forceinline__ int box_relation_type(const double* a, const double* b, double far) {
  if (a[0] < b[0] && a[1] < b[1] && a[2] > b[2] && a[3] > b[3]) return 1;
  if (a[0] > b[0] && a[1] > b[1] && a[2] < b[2] && a[3] < b[3]) return 2;
  const double iw = fmax(fmin(a[2], b[2]) - fmax(a[0], b[0]) + 1., 0.);
  const double ih = fmax(fmin(a[3], b[3]) - fmax(a[1], b[1]) + 1., 0.);
  // __dmul_rn / __dadd_rn: products and sums rounded separately like the reference's Python floats (no FMA contraction),
  // so the IoU >= 0.5 and distance >= far decisions see the same doubles
  const double inter = __dmul_rn(iw, ih);
  const double uni = __dadd_rn(__dadd_rn(__dmul_rn(a[2] - a[0] + 1., a[3] - a[1] + 1.),
                                         __dmul_rn(b[2] - b[0] + 1., b[3] - b[1] + 1.)), -inter);
  if (inter / uni >= 0.5) return 3;
  const double dx = (b[2] + b[0]) / 2 - (a[2] + a[0]) / 2, dy = (b[3] + b[1]) / 2 - (a[3] + a[1]) / 2;
  if (sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) >= far) return 0;
  double ang = atan2(dy, dx) / 3.141592653589793 * 180.;
  if (ang < 0) ang += 360.;
  return (int)ceil(ang / 45.) + 3;
}
__global__ void spatial_labels_kernel(const double* __restrict__ boxes, int N, int S, double far, long long total,
                                      double* __restrict__ labels) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / ((long long)S * S);
    const int ij = (int)(e % ((long long)S * S));
    const int i = ij / S, j = ij % S;
    int t = 0;
    if (i < N && j < N) {
      const double* bb = boxes + b * N * 4;
      double lo[4], hi[4];
      const int a = min(i, j), c = max(i, j);
#pragma unroll
      for (int q = 0; q < 4; ++q) { lo[q] = bb[a * 4 + q]; hi[q] = bb[c * 4 + q]; }
      t = box_relation_type(lo, hi, far);
      if (i > j) t = (t == 1) ? 2 : (t == 2) ? 1 : (t >= 8) ? t - 4 : (t >= 4) ? t + 4 : t;   // reverse_type
    }
    labels[e] = (double)t;
  }
}

// ---------------------------------------------------------------- semantic adjacency labels from detected classes
// get_semantic_adj ("feature extraction/combine_dicts.py":106-151): classes int32 [B, T] (anatomy ids, then disease ids
// offset by the anatomy count; ncls = background) -> int8 labels [B, S, S].  For the pair (a, b) = (min, max) of (i, j):
//   1 if both classes belong to the same organ group of the knowledge graph and one is an anatomy class, the other a disease
//     class; then, if both classes are among the co-occurrence diseases, max(co-occurrence[a-th, b-th], that value);
// the result is written at (i, j) and (j, i) alike.  The dictionaries of the reference arrive as per-class tables.
__global__ void semantic_labels_kernel(const int* __restrict__ cls, int T, int S, int ncls, const int* __restrict__ group,
                                       const unsigned char* __restrict__ in_ana, const unsigned char* __restrict__ in_di,
                                       const int* __restrict__ small_idx, const int* __restrict__ small_adj, int ns,
                                       long long total, signed char* __restrict__ labels) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / ((long long)S * S);
    const int ij = (int)(e % ((long long)S * S));
    const int i = ij / S, j = ij % S;
    int v = 0;
    if (i < T && j < T) {
      const int ca = cls[b * T + min(i, j)], cb = cls[b * T + max(i, j)];
      if (ca >= 0 && cb >= 0 && ca < ncls && cb < ncls) {
        if (group[ca] == group[cb] && ((in_ana[ca] && in_di[cb]) || (in_ana[cb] && in_di[ca]))) v = 1;
        const int sa = small_idx[ca], sb = small_idx[cb];
        if (sa >= 0 && sb >= 0) v = max(small_adj[sa * ns + sb], v);
      }
    }
    labels[e] = (signed char)v;
  }
}

// ---------------------------------------------------------------- Adam (utils/utils.py:96-99 -> torch.optim.Adam)
__device__ __forceinline__ void adam_one(float& pv, float gr, float& mv, float& vv, float lr_bc1, float rs_bc2, float b1,
                                         float b2, float eps, float wd) {
  if (wd != 0.f) gr = fmaf(wd, pv, gr);
  mv = b1 * mv + (1.f - b1) * gr;
  vv = b2 * vv + (1.f - b2) * gr * gr;
  const float denom = sqrtf(vv) * rs_bc2 + eps;
  pv = pv - lr_bc1 * (mv / denom);
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                            const float* __restrict__ pow_state) {
  ek_pdl_prologue();
  // pow_state = {b1^t, b2^t} lives in device memory so a captured CUDA graph stays valid across steps
  const float bc1 = 1.f - pow_state[0], bc2 = 1.f - pow_state[1];
  const float lr_bc1 = lr / bc1, rs_bc2 = 1.f / sqrtf(bc2);
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  const long long n4 = vec ? n / 4 : 0;
  // Four 16-byte groups per thread and pass, all 16 loads issued before the first use: the update that runs next to the
  // question-path BPTT is capped at one CTA per SM (step.py, EKAID_B200_BG_CTAS), and with one group per pass those 256
  // threads kept only 16 KB per SM in flight -- 3.9-4.3 TB/s on 1 GB of traffic (profiles/r02b_launches_summary.txt).
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; e + (U - 1) * stride < n4; e += U * stride) {
    float4 pv[U], mv[U], vv[U], gr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pv[u] = ((float4*)p)[e + u * stride];
      mv[u] = ((float4*)m)[e + u * stride];
      vv[u] = ((float4*)v)[e + u * stride];
      gr[u] = ((const float4*)g)[e + u * stride];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      adam_one(pv[u].x, gr[u].x, mv[u].x, vv[u].x, lr_bc1, rs_bc2, b1, b2, eps, wd);
      adam_one(pv[u].y, gr[u].y, mv[u].y, vv[u].y, lr_bc1, rs_bc2, b1, b2, eps, wd);
      adam_one(pv[u].z, gr[u].z, mv[u].z, vv[u].z, lr_bc1, rs_bc2, b1, b2, eps, wd);
      adam_one(pv[u].w, gr[u].w, mv[u].w, vv[u].w, lr_bc1, rs_bc2, b1, b2, eps, wd);
      ((float4*)m)[e + u * stride] = mv[u];
      ((float4*)v)[e + u * stride] = vv[u];
      ((float4*)p)[e + u * stride] = pv[u];
    }
  }
  for (; e < n4; e += stride) {
    float4 pv = ((float4*)p)[e], mv = ((float4*)m)[e], vv = ((float4*)v)[e];
    const float4 gr = ((const float4*)g)[e];
    adam_one(pv.x, gr.x, mv.x, vv.x, lr_bc1, rs_bc2, b1, b2, eps, wd);
    adam_one(pv.y, gr.y, mv.y, vv.y, lr_bc1, rs_bc2, b1, b2, eps, wd);
    adam_one(pv.z, gr.z, mv.z, vv.z, lr_bc1, rs_bc2, b1, b2, eps, wd);
    adam_one(pv.w, gr.w, mv.w, vv.w, lr_bc1, rs_bc2, b1, b2, eps, wd);
    ((float4*)m)[e] = mv;
    ((float4*)v)[e] = vv;
    ((float4*)p)[e] = pv;
  }
  for (long long e = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float pv = p[e], mv = m[e], vv = v[e];
    adam_one(pv, g[e], mv, vv, lr_bc1, rs_bc2, b1, b2, eps, wd);
    m[e] = mv;
    v[e] = vv;
    p[e] = pv;
  }
}

__global__ void adam_advance_kernel(float* pow_state, float b1, float b2) {
  ek_pdl_prologue();
  pow_state[0] *= b1;
  pow_state[1] *= b2;
}

// ---------------------------------------------------------------- train-mode dropout helpers
// VQ[m, 0:D] = drop(X[m,:]), VQ[m, D:D+Dq] = drop(flag[m] ? 0 : qv[(m / N) % B, :])
// = Dropout(0.2)(q_expand_v_cat(q, v))   (relation_encoder.py:19-29 + fc.py:25-32); element index m*(D+Dq) + c
template <typename T>
__global__ void build_vq_kernel(const float* __restrict__ X, const float* __restrict__ qv,
                                const uint8_t* __restrict__ flags, long long M, int N, int B, int D, int Dq,
                                T* __restrict__ VQ, EkDrop dr, bf16* __restrict__ VQB) {
  ek_pdl_prologue();
  // 8 consecutive columns per thread (D, Dq multiples of 8): one row lookup, 2 x 16-byte loads, one 16-byte store
  const int W = D + Dq;
  const int W8 = W / 8;
  const long long total = M * W8;
  const unsigned long long sd = ek_seed(dr);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / W8;
    const int c = (int)(t % W8) * 8;
    float v[8];
    if (c < D) {
      const float4 a = *(const float4*)(X + m * D + c), b = *(const float4*)(X + m * D + c + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if (flags[m]) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    } else {
      const float* q = qv + (size_t)((m / N) % B) * Dq + (c - D);
      const float4 a = *(const float4*)q, b = *(const float4*)(q + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    const unsigned long long e0 = (unsigned long long)m * W + c;
    float mk[8];
    ek_drop_multv<8>(dr, sd, e0, mk);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= mk[k];
    store_vec<T, 8>(VQ + m * W + c, v);
    if (VQB) store_vec<bf16, 8>(VQB + m * W + c, v);      // the backward's bf16 copy (wgrad operand), same values
  }
}
// out = sum_k mult_k(idx) * in_k   (k < nin <= 3; mult_k = 1 when site k has p = 0);  idx = m*C + c
// optional fp32 output (optionally accumulated into) and/or operand-type output
template <typename TI, typename TO, int V>
__global__ void drop_combine_kernel(int nin, const TI* __restrict__ in0, const TI* __restrict__ in1,
                                    const TI* __restrict__ in2, long long ldi, EkDrop d0, EkDrop d1, EkDrop d2,
                                    long long M, int C, float* __restrict__ outf, long long ldf, int accumulate,
                                    TO* __restrict__ outT, long long ldo) {
  ek_pdl_prologue();
  // V consecutive columns per thread (V = 4 needs C, pitches % 4 == 0; the compiler merges the accesses)
  const int CV = C / V;
  const long long total = M * CV;
  const unsigned long long s0 = ek_seed(d0), s1 = ek_seed(d1), s2 = ek_seed(d2);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / CV;
    const int c = (int)(t % CV) * V;
    const unsigned long long e0 = (unsigned long long)m * C + c;
    float v[V], mk[V], x[V];
    ek_drop_multv<V>(d0, s0, e0, mk);
    if constexpr (V == 4) load_vec<TI, 4>(in0 + m * ldi + c, x);
    else x[0] = to_f32<TI>(in0[m * ldi + c]);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = x[k] * mk[k];
    if (nin > 1) {
      ek_drop_multv<V>(d1, s1, e0, mk);
      if constexpr (V == 4) load_vec<TI, 4>(in1 + m * ldi + c, x);
      else x[0] = to_f32<TI>(in1[m * ldi + c]);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] += x[k] * mk[k];
    }
    if (nin > 2) {
      ek_drop_multv<V>(d2, s2, e0, mk);
      if constexpr (V == 4) load_vec<TI, 4>(in2 + m * ldi + c, x);
      else x[0] = to_f32<TI>(in2[m * ldi + c]);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] += x[k] * mk[k];
    }
    if (outf) {
      if (accumulate) {
        if constexpr (V == 4) load_vec<float, 4>(outf + m * ldf + c, x);
        else x[0] = outf[m * ldf + c];
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] += x[k];
      }
      if constexpr (V == 4) store_vec<float, 4>(outf + m * ldf + c, v);
      else outf[m * ldf + c] = v[0];
    }
    if (outT) {
      if constexpr (V == 4) store_vec<TO, 4>(outT + m * ldo + c, v);
      else outT[m * ldo + c] = from_f32<TO>(v[0]);
    }
  }
}
// out0 = mult_0 * in, out1 = mult_1 * in (two independent dropout sites on the same tensor: the query and key inputs of
// a relation layer, graph_att_layer.py:77,89 through fc.py:25-32); index m*C + c; 4 columns per thread.
template <typename T>
__global__ void drop_fanout_kernel(const T* __restrict__ in, long long ldi, EkDrop d0, EkDrop d1, long long M, int C,
                                   T* __restrict__ out0, T* __restrict__ out1, long long ldo, bf16* __restrict__ out0B,
                                   bf16* __restrict__ out1B) {
  ek_pdl_prologue();
  const int CV = C / 4;
  const long long total = M * CV;
  const unsigned long long s0 = ek_seed(d0), s1 = ek_seed(d1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / CV;
    const int c = (int)(t % CV) * 4;
    const unsigned long long e0 = (unsigned long long)m * C + c;
    float x[4], m0[4], m1[4], a[4], b[4];
    load_vec<T, 4>(in + m * ldi + c, x);
    ek_drop_multv<4>(d0, s0, e0, m0);
    ek_drop_multv<4>(d1, s1, e0, m1);
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] = x[k] * m0[k]; b[k] = x[k] * m1[k]; }
    store_vec<T, 4>(out0 + m * ldo + c, a);
    store_vec<T, 4>(out1 + m * ldo + c, b);
    if (out0B) {                                         // the backward's bf16 copies (wgrad operands)
      store_vec<bf16, 4>(out0B + m * ldo + c, a);
      store_vec<bf16, 4>(out1B + m * ldo + c, b);
    }
  }
}

// ---------------------------------------------------------------- legacy weight_norm(dim=None) (fc.py:33-34)
// w = v * g / ||v||_F.  Two launches forward (partials; scale) and two backward, all deterministic.
constexpr int WN_BLOCKS = 128;
__global__ void wn_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                  float* __restrict__ part) {
  ek_pdl_prologue();     // part[blk] = sum a*b over the block's slice
  __shared__ float red[32];
  float s = 0.f;
  const long long n4 = ((((uintptr_t)a | (uintptr_t)b) & 15) == 0) ? n / 4 : 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 x = ((const float4*)a)[e], y = ((const float4*)b)[e];
    s = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, s))));
  }
  for (long long e = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    s = fmaf(a[e], b[e], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__device__ __forceinline__ float wn_total(const float* __restrict__ part, float* red) {
  float s = (threadIdx.x < WN_BLOCKS) ? part[threadIdx.x] : 0.f;
  return block_sum(s, red);
}
__global__ void wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ part,
                                long long n, float* __restrict__ w, float* __restrict__ norm_out) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const float nrm = sqrtf(wn_total(part, red));
  const float sc = g[0] / nrm;
  if (blockIdx.x == 0 && threadIdx.x == 0) norm_out[0] = nrm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    w[e] = v[e] * sc;
}
// dv = (g/n) dw - (g dot / n^3) v ;  dg = dot / n ;  dot = sum dw*v
__global__ void wn_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                              const float* __restrict__ norm, const float* __restrict__ part, long long n,
                              float* __restrict__ dv, float* __restrict__ dg) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const float dot = wn_total(part, red);
  const float nrm = norm[0], gv = g[0];
  const float c1 = gv / nrm, c2 = gv * dot / (nrm * nrm * nrm);
  if (blockIdx.x == 0 && threadIdx.x == 0) dg[0] = dot / nrm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    dv[e] = c1 * dw[e] - c2 * v[e];
}

// Batched form: the same arithmetic for up to WN_MANY tensors in two launches (blockIdx.y = tensor).  All
// weight-normalised matrices of the relation encoders are independent of the activations, so the step normalises them
// together up front and differentiates them together at the end.
constexpr int WN_MANY = 16;
struct WnMany {
  const float* a[WN_MANY];      // forward: v            backward: dw
  const float* b[WN_MANY];      // forward: v            backward: v
  const float* g[WN_MANY];
  float* out[WN_MANY];          // forward: w            backward: dv
  float* out1[WN_MANY];         // forward: unused       backward: dg (1 element)
  long long n[WN_MANY];
};
__global__ void wn_many_partial_kernel(WnMany t, float* __restrict__ part) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const int i = blockIdx.y;
  const float* __restrict__ a = t.a[i];
  const float* __restrict__ b = t.b[i];
  const long long n = t.n[i];
  float s = 0.f;
  const long long n4 = ((((uintptr_t)a | (uintptr_t)b) & 15) == 0) ? n / 4 : 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 x = ((const float4*)a)[e], y = ((const float4*)b)[e];
    s = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, s))));
  }
  for (long long e = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    s = fmaf(a[e], b[e], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[(size_t)i * WN_BLOCKS + blockIdx.x] = s;
}
__global__ void wn_many_scale_kernel(WnMany t, const float* __restrict__ part, float* __restrict__ norms) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const int i = blockIdx.y;
  const long long n = t.n[i];
  if ((long long)blockIdx.x * blockDim.x * 4 >= n && blockIdx.x != 0) return;      // nothing to do for this block
  const float nrm = sqrtf(wn_total(part + (size_t)i * WN_BLOCKS, red));
  const float sc = t.g[i][0] / nrm;
  if (blockIdx.x == 0 && threadIdx.x == 0) norms[i] = nrm;
  const float* __restrict__ v = t.b[i];
  float* __restrict__ w = t.out[i];
  const long long n4 = ((((uintptr_t)v | (uintptr_t)w) & 15) == 0) ? n / 4 : 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 x = ((const float4*)v)[e];
    ((float4*)w)[e] = make_float4(x.x * sc, x.y * sc, x.z * sc, x.w * sc);
  }
  for (long long e = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    w[e] = v[e] * sc;
}
__global__ void wn_many_bwd_kernel(WnMany t, const float* __restrict__ part, const float* __restrict__ norms) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const int i = blockIdx.y;
  const long long n = t.n[i];
  if ((long long)blockIdx.x * blockDim.x * 4 >= n && blockIdx.x != 0) return;
  const float dot = wn_total(part + (size_t)i * WN_BLOCKS, red);
  const float nrm = norms[i], gv = t.g[i][0];
  const float c1 = gv / nrm, c2 = gv * dot / (nrm * nrm * nrm);
  if (blockIdx.x == 0 && threadIdx.x == 0) t.out1[i][0] = dot / nrm;
  const float* __restrict__ dw = t.a[i];
  const float* __restrict__ v = t.b[i];
  float* __restrict__ dv = t.out[i];
  const long long n4 = ((((uintptr_t)dw | (uintptr_t)v | (uintptr_t)dv) & 15) == 0) ? n / 4 : 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 x = ((const float4*)dw)[e], y = ((const float4*)v)[e];
    ((float4*)dv)[e] = make_float4(c1 * x.x - c2 * y.x, c1 * x.y - c2 * y.y, c1 * x.z - c2 * y.z, c1 * x.w - c2 * y.w);
  }
  for (long long e = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    dv[e] = c1 * dw[e] - c2 * v[e];
}

// y[m, n] = x[m, :] . W[n, :] + b[n]  for a handful of outputs (fc1: 6 change classes, modules.py:312) -- one warp per (m, n)
__global__ void small_linear_kernel(const float* __restrict__ x, long long ldx, int M, int K, const float* __restrict__ W,
                                    const float* __restrict__ b, int N, float* __restrict__ y) {
  ek_pdl_prologue();
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= M * N) return;
  const int m = pair / N, n = pair % N, lane = threadIdx.x & 31;
  const float* xr = x + (size_t)m * ldx;
  const float* wr = W + (size_t)n * K;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(xr[k], __ldg(wr + k), s);
  s = warp_sum(s);
  if (lane == 0) y[(size_t)m * N + n] = s + (b ? b[n] : 0.f);
}

// out[0] = sum_k coef[k] * <a_k, w_k>  (w_k = NULL: plain sum of a_k); k < 5.  The training objective of
// train_mimic.py:246-247 with the decoder's gradient given as fixed cotangents: one CTA, fixed summation order.
struct WsumArgs {
  const float* a[5];
  const float* w[5];
  long long n[5];
  float coef[5];
};
__global__ void __launch_bounds__(256)
weighted_sums_kernel(WsumArgs t, int count, float* part, unsigned int* ticket, float* __restrict__ out) {
  ek_pdl_prologue();
  __shared__ float red[32];
  __shared__ int is_last;
  float total = 0.f;
  for (int k = 0; k < count; ++k) {
    const float* __restrict__ a = t.a[k];
    const float* __restrict__ w = t.w[k];
    float s = 0.f;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < t.n[k]; e += (long long)gridDim.x * blockDim.x)
      s = w ? fmaf(a[e], w[e], s) : s + a[e];
    total = fmaf(t.coef[k], s, total);
  }
  total = block_sum(total, red);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = total;
    __threadfence();
    const unsigned int tk = atomicAdd(ticket, 1u);
    is_last = (tk == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {       // fixed summation order: deterministic
    __threadfence();
    float s = 0.f;
    for (unsigned int p = 0; p < gridDim.x; ++p) s += __ldcg(part + p);
    out[0] = s;
  }
}

// backward of weighted_sums: grad_k[e] = g * coef_k * (w_k ? w_k[e] : 1), all k in one launch (blockIdx.y = k)
__global__ void weighted_sums_bwd_kernel(WsumArgs t, const float* __restrict__ g) {
  ek_pdl_prologue();
  const int k = blockIdx.y;
  const float s = g[0] * t.coef[k];
  float* __restrict__ o = const_cast<float*>(t.a[k]);
  const float* __restrict__ w = t.w[k];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < t.n[k]; e += (long long)gridDim.x * blockDim.x)
    o[e] = w ? s * w[e] : s;
}
// input_attended = attended_2 - attended_1 (modules.py:309): attended [2B, D] (main rows, then reference rows) -> ia [B, D]
__global__ void head_fwd_kernel(const float* __restrict__ attended, long long BD, float* __restrict__ ia) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < BD; e += (long long)gridDim.x * blockDim.x)
    ia[e] = attended[BD + e] - attended[e];
}
// gradients reaching the five views the module returns -> the two tensors the fusion stage produced, in ONE launch:
// d_att = [d_att_bef ; d_att_aft],  d_attended = [d_a1 - d_ia ; d_a2 + d_ia]   (any input may be NULL = zero)
__global__ void head_bwd_kernel(const float* __restrict__ d_att_bef, const float* __restrict__ d_att_aft,
                                const float* __restrict__ d_a1, const float* __restrict__ d_a2,
                                const float* __restrict__ d_ia, long long BN, long long BD, float* __restrict__ d_att,
                                float* __restrict__ d_attended) {
  ek_pdl_prologue();
  const long long total = 2 * BN + 2 * BD;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    if (e < BN) d_att[e] = d_att_bef ? d_att_bef[e] : 0.f;
    else if (e < 2 * BN) d_att[e] = d_att_aft ? d_att_aft[e - BN] : 0.f;
    else if (e < 2 * BN + BD) {
      const long long i = e - 2 * BN;
      d_attended[i] = (d_a1 ? d_a1[i] : 0.f) - (d_ia ? d_ia[i] : 0.f);
    } else {
      const long long i = e - 2 * BN - BD;
      d_attended[BD + i] = (d_a2 ? d_a2[i] : 0.f) + (d_ia ? d_ia[i] : 0.f);
    }
  }
}

__global__ void rng_advance_kernel(unsigned long long* seed) {
  ek_pdl_prologue(); *seed = *seed * 6364136223846793005ull + 1442695040888963407ull; }

inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int ek_drop_mask_launch(EkDrop dr, long long n, float* out, cudaStream_t st) {
  if (n <= 0) return EK_OK;
  ek_launch(drop_mask_kernel, grid_for(n), 256, 0, st, dr, n, out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_split3_bf16_launch(const float* src, long long lds, bf16* dst, long long ldd, long long rows, int cols,
                          int pattern, int along_rows, cudaStream_t st) {
  EK_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0,
             EK_ERR_ALIGN, "split3_bf16: cols/pitch must be multiples of 4 and pointers aligned");
  if (rows * cols == 0) return EK_OK;
  ek_launch(split3_bf16_kernel, grid_for(rows * cols / 4), 256, 0, st, src, lds, dst, ldd, rows, cols, pattern,
            along_rows);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_cast_f32_bf16_launch(const float* src, long long lds, bf16* dst, long long ldd, long long rows, int cols,
                            int fmt, cudaStream_t st) {
  EK_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0,
             EK_ERR_ALIGN, "cast_f32_bf16: cols/pitch must be multiples of 4 and pointers aligned");
  if (rows * cols == 0) return EK_OK;
  ek_launch(cast_f32_bf16_kernel, grid_for(rows * cols / 4), 256, 0, st, src, lds, dst, ldd, rows, cols, fmt);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
// count <= 16 blocks; all arrays are host arrays read at call time.  mode: 0 fp32->bf16, 1 fp32->fp32, 2 bytes (cols = bytes
// per row, multiple of 16; pitches in bytes).  Modes 0/1: cols and pitches (elements) multiples of 4, 16-byte aligned.
int ek_cast_many_launch(int count, const void* const* src, const long long* lds, void* const* dst, const long long* ldd,
                        const long long* rows, const int* cols, const int* mode, cudaStream_t st) {
  EK_REQUIRE(count >= 1 && count <= CM_MAX, EK_ERR_SHAPE, "cast_many: count=%d not in [1,%d]", count, CM_MAX);
  CastMany t = {};
  long long most = 0;
  for (int i = 0; i < count; ++i) {
    const int unit = mode[i] == 2 ? 16 : (mode[i] == 4 ? 8 : 4);
    EK_REQUIRE(mode[i] >= 0 && mode[i] <= 4 && cols[i] % unit == 0 && lds[i] % unit == 0 && ldd[i] % unit == 0 &&
                   ((uintptr_t)src[i] & 15) == 0 && ((uintptr_t)dst[i] & ((mode[i] == 0 || mode[i] == 3) ? 7 : 15)) == 0,
               EK_ERR_ALIGN, "cast_many: block %d: cols/pitches must be multiples of %d and pointers aligned", i, unit);
    t.src[i] = src[i]; t.dst[i] = dst[i]; t.lds[i] = lds[i]; t.ldd[i] = ldd[i]; t.rows[i] = rows[i]; t.cols[i] = cols[i];
    t.mode[i] = mode[i];
    const long long work = rows[i] * (cols[i] / unit);
    if (work > most) most = work;
  }
  if (most == 0) return EK_OK;
  int gx = (int)((most + 255) / 256);
  if (gx > 148 * 4) gx = 148 * 4;
  ek_launch(cast_many_kernel, dim3(gx, count), 256, 0, st, t);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_copy_f32_launch(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols,
                       cudaStream_t st) {
  EK_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0,
             EK_ERR_ALIGN, "copy_f32: cols/pitch must be multiples of 4 and pointers aligned");
  if (rows * cols == 0) return EK_OK;
  ek_launch(copy_f32_kernel, grid_for(rows * cols / 4), 256, 0, st, src, lds, dst, ldd, rows, cols);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_cast_bf16_f32_launch(const bf16* src, long long lds, float* dst, long long ldd, long long rows, int cols,
                            float scale, cudaStream_t st) {
  if (rows * cols == 0) return EK_OK;
  ek_launch(cast_bf16_f32_kernel, grid_for(rows * cols), 256, 0, st, src, lds, dst, ldd, rows, cols, scale);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

// workspace: >= 64 * N floats
int ek_colsum_launch(int is_bf16, const void* src, long long ld, long long M, int N, const float* rowscale, float* out,
                     float* workspace, cudaStream_t st) {
  int nparts = (int)((M + 31) / 32);       // enough CTAs to cover the machine even for the narrow (N = 1024) case
  if (nparts > 64) nparts = 64;
  if (nparts < 1) nparts = 1;
  // workspace: [1024 ticket counters, one per 32-column block, zero between calls] [64 * N partial sums].  The split is
  // the same for every N so that calls with different N can share one workspace.
  EK_REQUIRE(N <= 32 * 1024, EK_ERR_SHAPE, "colsum: N=%d > 32768", N);
  unsigned int* tickets = (unsigned int*)workspace;
  float* part = workspace + 1024;
  const int es = is_bf16 ? 2 : 4;
  const int vec_ok = (((uintptr_t)src & 15) == 0 && (ld * es) % 16 == 0) ? 1 : 0;
  if (is_bf16)
    ek_launch(colsum_kernel<bf16, 8>, dim3(ek_div_up(N, 256), nparts), 256, 0, st, (const bf16*)src, ld, M, N, rowscale, part,
              nparts, tickets, out, vec_ok);
  else
    ek_launch(colsum_kernel<float, 4>, dim3(ek_div_up(N, 128), nparts), 256, 0, st, (const float*)src, ld, M, N, rowscale,
              part, nparts, tickets, out, vec_ok);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

// workspace >= count * (1024 + 64 * Nmax) floats (Nmax = widest job); the first count*1024 words are ticket counters
// (zero between calls).  All jobs sum over the same M rows.
int ek_colsum_many_launch(int count, const int* is_bf16, const void* const* src, const long long* ld, long long M,
                          const int* N, float* const* out, float* workspace, cudaStream_t st) {
  EK_REQUIRE(count >= 1 && count <= CSM_MAX, EK_ERR_SHAPE, "colsum_many: count=%d not in [1,%d]", count, CSM_MAX);
  ColsumMany t = {};
  int nmax = 0;
  for (int j = 0; j < count; ++j) {
    EK_REQUIRE(N[j] > 0 && N[j] <= 32 * 1024, EK_ERR_SHAPE, "colsum_many: N[%d]=%d", j, N[j]);
    if (N[j] > nmax) nmax = N[j];
  }
  int nparts = (int)((M + 31) / 32);
  if (nparts > 64) nparts = 64;
  if (nparts < 1) nparts = 1;
  for (int j = 0; j < count; ++j) {
    const int es = is_bf16[j] ? 2 : 4;
    t.src[j] = src[j]; t.out[j] = out[j]; t.ld[j] = ld[j]; t.N[j] = N[j]; t.is_bf16[j] = is_bf16[j];
    t.vec_ok[j] = (((uintptr_t)src[j] & 15) == 0 && (ld[j] * es) % 16 == 0) ? 1 : 0;
    t.part_off[j] = (long long)count * 1024 + (long long)j * 64 * nmax;
  }
  ek_launch(colsum_many_kernel, dim3(ek_div_up(nmax, 128), nparts, count), 256, 0, st, t, M, nparts, workspace);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_row_zero_flags_launch(const float* X, long long M, int D, uint8_t* flags, cudaStream_t st) {
  ek_launch(row_zero_flags_kernel, ek_div_up(M, 8), 256, 0, st, X, M, D, flags);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_group_rowsum_launch(int is_bf16, const void* src, long long ld, int N, int B, int S, int D, const uint8_t* flags,
                           float* out, cudaStream_t st) {
  dim3 grid(B, ek_div_up(D, 64));
  if (is_bf16) ek_launch(group_rowsum_kernel<bf16>, grid, 256, 0, st, (const bf16*)src, ld, N, B, S, D, flags, out);
  else ek_launch(group_rowsum_kernel<float>, grid, 256, 0, st, (const float*)src, ld, N, B, S, D, flags, out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_combine_diff_fwd_launch(int is_bf16, const float* X3, long long BN, int D, int mode, float c1, float c2,
                               float c3, float* Xc, void* CAT, cudaStream_t st) {
  if (is_bf16)
    ek_launch(combine_diff_fwd_kernel<bf16>, grid_for(BN * D), 256, 0, st, X3, BN, D, mode, c1, c2, c3, Xc, (bf16*)CAT);
  else
    ek_launch(combine_diff_fwd_kernel<float>, grid_for(BN * D), 256, 0, st, X3, BN, D, mode, c1, c2, c3, Xc, (float*)CAT);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_combine_diff_bwd_launch(const float* dXc, const float* dCAT, long long BN, int D, int mode, float c1, float c2,
                               float c3, float* dX3, cudaStream_t st) {
  ek_launch(combine_diff_bwd_kernel, grid_for(BN * D), 256, 0, st, dXc, dCAT, BN, D, mode, c1, c2, c3, dX3);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_gate_fwd_launch(int is_bf16, const float* pre, long long M, int D, void* ctx, void* gate, void* CAT, EkDrop dc,
                       EkDrop dg, cudaStream_t st) {
  EK_REQUIRE(D % 4 == 0 && ((uintptr_t)pre & 15) == 0, EK_ERR_ALIGN, "gate_fwd: D=%d must be a multiple of 4, pre 16-byte aligned", D);
  if (is_bf16)
    ek_launch(gate_fwd_kernel<bf16>, grid_for(M * D / 4), 256, 0, st, pre, M, D, (bf16*)ctx, (bf16*)gate, (bf16*)CAT, dc, dg);
  else
    ek_launch(gate_fwd_kernel<float>, grid_for(M * D / 4), 256, 0, st, pre, M, D, (float*)ctx, (float*)gate, (float*)CAT, dc, dg);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_gate_bwd_launch(int is_bf16, const float* dCAT, const void* ctx, const void* gate, long long M, int D,
                       void* dpre, EkDrop dc, EkDrop dg, cudaStream_t st) {
  EK_REQUIRE(D % 4 == 0 && ((uintptr_t)dCAT & 15) == 0, EK_ERR_ALIGN, "gate_bwd: D=%d must be a multiple of 4, dCAT 16-byte aligned", D);
  if (is_bf16)
    ek_launch(gate_bwd_kernel<bf16>, grid_for(M * D / 4), 256, 0, st, dCAT, (const bf16*)ctx, (const bf16*)gate, M, D, (bf16*)dpre,
                                                           dc, dg);
  else
    ek_launch(gate_bwd_kernel<float>, grid_for(M * D / 4), 256, 0, st, dCAT, (const float*)ctx, (const float*)gate, M, D,
                                                            (float*)dpre, dc, dg);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_att_pool_fwd_launch(const float* E, long long M, int N, int D, int dim, const float* w, const float* b,
                           const float* Xc, float* att, float* attended, cudaStream_t st) {
  ek_launch(att_score_kernel, ek_div_up(M, 8), 256, 0, st, E, M, dim, w, b, att);
  EK_CHECK_LAUNCH();
  ek_launch(att_pool_kernel, dim3((unsigned)(M / N), ek_div_up(D, 128)), 128, 0, st, att, Xc, N, D, attended);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_att_pool_bwd_launch(int is_bf16, const float* dA, const float* dattw, const float* att, const float* Xc,
                           const float* E, const float* w, long long M, int N, int D, int dim, float* dXc, void* dE,
                           float* dpre, float escale, cudaStream_t st) {
  if (is_bf16)
    ek_launch(att_pool_bwd_kernel<bf16>, ek_div_up(M, 8), 256, 0, st, dA, dattw, att, Xc, E, w, M, N, D, dim, dXc, (bf16*)dE,
                                                                dpre, escale);
  else
    ek_launch(att_pool_bwd_kernel<float>, ek_div_up(M, 8), 256, 0, st, dA, dattw, att, Xc, E, w, M, N, D, dim, dXc,
                                                                 (float*)dE, dpre, escale);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_spatial_labels_launch(const double* boxes, int B, int N, int S, double lx, double ly, double* labels,
                             cudaStream_t st) {
  const long long total = (long long)B * S * S;
  if (total == 0) return EK_OK;
  ek_launch(spatial_labels_kernel, grid_for(total), 256, 0, st, boxes, N, S, (lx + ly) / 3., total, labels);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_semantic_labels_launch(const int* cls, int B, int T, int S, int ncls, const int* group, const unsigned char* in_ana,
                              const unsigned char* in_di, const int* small_idx, const int* small_adj, int ns,
                              signed char* labels, cudaStream_t st) {
  const long long total = (long long)B * S * S;
  if (total == 0) return EK_OK;
  ek_launch(semantic_labels_kernel, grid_for(total), 256, 0, st, cls, T, S, ncls, group, in_ana, in_di, small_idx,
            small_adj, ns, total, labels);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_onehot_adj_launch(const double* labels, int B, int S, int N, int L, float* out, cudaStream_t st) {
  const long long total = (long long)B * N * N;
  if (total == 0) return EK_OK;
  ek_launch(onehot_adj_kernel, grid_for(total), 256, 0, st, labels, S, N, L, total, out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_onehot_adj_i8_launch(const int8_t* labels, int B, int S, int N, int L, float* out, cudaStream_t st) {
  const long long total = (long long)B * N * N;
  if (total == 0) return EK_OK;
  ek_launch(onehot_adj_i8_kernel, grid_for(total), 256, 0, st, labels, S, N, L, total, out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_adam_launch(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                   float wd, const float* pow_state, int max_ctas, cudaStream_t st) {
  if (n == 0) return EK_OK;
  // one thread per FOUR 16-byte groups (the kernel's unrolled pass): a segment smaller than the capped grid x 4 groups
  // would otherwise run entirely in the one-group remainder loop, at two CTAs per SM (122 registers) -- measured 35 us
  // instead of 27 us for the question-path segment that closes the step (profiles/r02b_ncu_adam_embed.csv)
  int grid = grid_for(n / 16 + 1);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;      // background mode: leave SM slots to concurrent kernels
  static bool carveout_set = false;
  if (!carveout_set) {
    // An SM cannot change its L1 / shared-memory split while CTAs are resident.  This kernel needs no shared memory, but
    // it runs next to kernels that need almost all of it (the GRU recurrence, the GEMMs): ask for the maximum carve-out
    // so their CTAs can join an SM that already hosts ours.
    cudaFuncSetAttribute(adam_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  ek_launch(adam_kernel, grid, 256, 0, st, p, g, m, v, n, lr, b1, b2, eps, wd, pow_state);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_adam_advance_launch(float* pow_state, float b1, float b2, cudaStream_t st) {
  ek_launch(adam_advance_kernel, 1, 1, 0, st, pow_state, b1, b2);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_build_vq_launch(int is_bf16, const float* X, const float* qv, const uint8_t* flags, long long M, int N, int B,
                       int D, int Dq, void* VQ, EkDrop dr, void* VQB, cudaStream_t st) {
  EK_REQUIRE(D % 8 == 0 && Dq % 8 == 0, EK_ERR_SHAPE, "build_vq: D=%d Dq=%d must be multiples of 8", D, Dq);
  const int g = grid_for(M * ((D + Dq) / 8));
  if (is_bf16 == 2) ek_launch(build_vq_kernel<f16>, g, 256, 0, st, X, qv, flags, M, N, B, D, Dq, (f16*)VQ, dr, (bf16*)VQB);
  else if (is_bf16) ek_launch(build_vq_kernel<bf16>, g, 256, 0, st, X, qv, flags, M, N, B, D, Dq, (bf16*)VQ, dr, (bf16*)VQB);
  else ek_launch(build_vq_kernel<float>, g, 256, 0, st, X, qv, flags, M, N, B, D, Dq, (float*)VQ, dr, (bf16*)nullptr);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
template <int V>
static void drop_combine_dispatch(int in_bf16, int out_bf16, int nin, const void* in0, const void* in1, const void* in2,
                                  long long ldi, EkDrop d0, EkDrop d1, EkDrop d2, long long M, int C, float* outf,
                                  long long ldf, int accumulate, void* outT, long long ldo, cudaStream_t st) {
  const int g = grid_for(M * (C / V));
  if (in_bf16 == 2 && out_bf16 == 2)
    ek_launch(drop_combine_kernel<f16, f16, V>, g, 256, 0, st, nin, (const f16*)in0, (const f16*)in1, (const f16*)in2, ldi,
                                                       d0, d1, d2, M, C, outf, ldf, accumulate, (f16*)outT, ldo);
  else if (in_bf16 == 2 && out_bf16 == 0)
    ek_launch(drop_combine_kernel<f16, float, V>, g, 256, 0, st, nin, (const f16*)in0, (const f16*)in1, (const f16*)in2, ldi,
                                                         d0, d1, d2, M, C, outf, ldf, accumulate, (float*)outT, ldo);
  else if (in_bf16 == 0 && out_bf16 == 2)
    ek_launch(drop_combine_kernel<float, f16, V>, g, 256, 0, st, nin, (const float*)in0, (const float*)in1, (const float*)in2,
                                                         ldi, d0, d1, d2, M, C, outf, ldf, accumulate, (f16*)outT, ldo);
  else if (in_bf16 && out_bf16)
    ek_launch(drop_combine_kernel<bf16, bf16, V>, g, 256, 0, st, nin, (const bf16*)in0, (const bf16*)in1, (const bf16*)in2, ldi,
                                                         d0, d1, d2, M, C, outf, ldf, accumulate, (bf16*)outT, ldo);
  else if (in_bf16)
    ek_launch(drop_combine_kernel<bf16, float, V>, g, 256, 0, st, nin, (const bf16*)in0, (const bf16*)in1, (const bf16*)in2, ldi,
                                                          d0, d1, d2, M, C, outf, ldf, accumulate, (float*)outT, ldo);
  else if (out_bf16)
    ek_launch(drop_combine_kernel<float, bf16, V>, g, 256, 0, st, nin, (const float*)in0, (const float*)in1, (const float*)in2,
                                                          ldi, d0, d1, d2, M, C, outf, ldf, accumulate, (bf16*)outT, ldo);
  else
    ek_launch(drop_combine_kernel<float, float, V>, g, 256, 0, st, nin, (const float*)in0, (const float*)in1,
                                                           (const float*)in2, ldi, d0, d1, d2, M, C, outf, ldf,
                                                           accumulate, (float*)outT, ldo);
}
int ek_drop_combine_launch(int in_bf16, int out_bf16, int nin, const void* in0, const void* in1, const void* in2,
                           long long ldi, EkDrop d0, EkDrop d1, EkDrop d2, long long M, int C, float* outf,
                           long long ldf, int accumulate, void* outT, long long ldo, cudaStream_t st) {
  const uintptr_t ptrs = (uintptr_t)in0 | (uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)outf | (uintptr_t)outT;
  const bool v4 = (C % 4 == 0) && (ldi % 4 == 0) && (!outf || ldf % 4 == 0) && (!outT || ldo % 4 == 0) && (ptrs & 15) == 0;
  if (v4) drop_combine_dispatch<4>(in_bf16, out_bf16, nin, in0, in1, in2, ldi, d0, d1, d2, M, C, outf, ldf, accumulate, outT, ldo, st);
  else drop_combine_dispatch<1>(in_bf16, out_bf16, nin, in0, in1, in2, ldi, d0, d1, d2, M, C, outf, ldf, accumulate, outT, ldo, st);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_drop_fanout_launch(int is_bf16, const void* in, long long ldi, EkDrop d0, EkDrop d1, long long M, int C, void* out0,
                          void* out1, long long ldo, void* out0B, void* out1B, cudaStream_t st) {
  const uintptr_t ptrs = (uintptr_t)in | (uintptr_t)out0 | (uintptr_t)out1;
  EK_REQUIRE(C % 4 == 0 && ldi % 4 == 0 && ldo % 4 == 0 && (ptrs & 15) == 0, EK_ERR_ALIGN,
             "drop_fanout: C=%d and the pitches must be multiples of 4, pointers 16-byte aligned", C);
  const int g = grid_for(M * (C / 4));
  if (is_bf16 == 2)
    ek_launch(drop_fanout_kernel<f16>, g, 256, 0, st, (const f16*)in, ldi, d0, d1, M, C, (f16*)out0, (f16*)out1, ldo,
                                                    (bf16*)out0B, (bf16*)out1B);
  else if (is_bf16)
    ek_launch(drop_fanout_kernel<bf16>, g, 256, 0, st, (const bf16*)in, ldi, d0, d1, M, C, (bf16*)out0, (bf16*)out1, ldo,
                                                     (bf16*)out0B, (bf16*)out1B);
  else
    ek_launch(drop_fanout_kernel<float>, g, 256, 0, st, (const float*)in, ldi, d0, d1, M, C, (float*)out0, (float*)out1, ldo,
                                                      (bf16*)nullptr, (bf16*)nullptr);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_rng_advance_launch(unsigned long long* seed, cudaStream_t st) {
  ek_launch(rng_advance_kernel, 1, 1, 0, st, seed);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

// workspace: WN_BLOCKS (128) floats
int ek_wn_fwd_launch(const float* v, const float* g, long long n, float* w, float* norm_out, float* workspace,
                     cudaStream_t st) {
  ek_launch(wn_partial_kernel, WN_BLOCKS, 256, 0, st, v, v, n, workspace);
  EK_CHECK_LAUNCH();
  ek_launch(wn_scale_kernel, grid_for(n), 256, 0, st, v, g, workspace, n, w, norm_out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_small_linear_launch(const float* x, long long ldx, int M, int K, const float* W, const float* b, int N, float* y,
                           cudaStream_t st) {
  EK_REQUIRE(M > 0 && N > 0 && K > 0, EK_ERR_SHAPE, "small_linear: bad shape M=%d N=%d K=%d", M, N, K);
  ek_launch(small_linear_kernel, ek_div_up((long long)M * N, 8), 256, 0, st, x, ldx, M, K, W, b, N, y);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
// workspace: 1 ticket word (zero between calls) + 64 partial sums
int ek_weighted_sums_launch(int count, const float* const* a, const float* const* w, const long long* n,
                            const float* coef, float* out, float* workspace, cudaStream_t st) {
  EK_REQUIRE(count >= 1 && count <= 5, EK_ERR_SHAPE, "weighted_sums: count=%d not in [1,5]", count);
  WsumArgs t = {};
  for (int k = 0; k < count; ++k) { t.a[k] = a[k]; t.w[k] = w[k]; t.n[k] = n[k]; t.coef[k] = coef[k]; }
  ek_launch(weighted_sums_kernel, 64, 256, 0, st, t, count, workspace + 1, (unsigned int*)workspace, out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_weighted_sums_bwd_launch(int count, float* const* out, const float* const* w, const long long* n, const float* coef,
                                const float* g, cudaStream_t st) {
  EK_REQUIRE(count >= 1 && count <= 5, EK_ERR_SHAPE, "weighted_sums_bwd: count=%d not in [1,5]", count);
  WsumArgs t = {};
  for (int k = 0; k < count; ++k) { t.a[k] = out[k]; t.w[k] = w[k]; t.n[k] = n[k]; t.coef[k] = coef[k]; }
  ek_launch(weighted_sums_bwd_kernel, dim3(32, count), 256, 0, st, t, g);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_head_fwd_launch(const float* attended, long long BD, float* ia, cudaStream_t st) {
  ek_launch(head_fwd_kernel, grid_for(BD), 256, 0, st, attended, BD, ia);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_head_bwd_launch(const float* d_att_bef, const float* d_att_aft, const float* d_a1, const float* d_a2,
                       const float* d_ia, long long BN, long long BD, float* d_att, float* d_attended, cudaStream_t st) {
  ek_launch(head_bwd_kernel, grid_for(2 * BN + 2 * BD), 256, 0, st, d_att_bef, d_att_aft, d_a1, d_a2, d_ia, BN, BD, d_att,
                                              d_attended);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
// norms: [count] floats kept for the backward; workspace: count * 128 floats
int ek_wn_fwd_many_launch(int count, const float* const* v, const float* const* g, const long long* n, float* const* w,
                          float* norms, float* workspace, cudaStream_t st) {
  EK_REQUIRE(count >= 1 && count <= WN_MANY, EK_ERR_SHAPE, "wn_fwd_many: count=%d not in [1,%d]", count, WN_MANY);
  WnMany t = {};
  for (int i = 0; i < count; ++i) { t.a[i] = v[i]; t.b[i] = v[i]; t.g[i] = g[i]; t.out[i] = w[i]; t.n[i] = n[i]; }
  ek_launch(wn_many_partial_kernel, dim3(WN_BLOCKS, count), 256, 0, st, t, workspace);
  EK_CHECK_LAUNCH();
  ek_launch(wn_many_scale_kernel, dim3(296, count), 256, 0, st, t, workspace, norms);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_wn_bwd_many_launch(int count, const float* const* dw, const float* const* v, const float* const* g,
                          const float* norms, const long long* n, float* const* dv, float* const* dg, float* workspace,
                          cudaStream_t st) {
  EK_REQUIRE(count >= 1 && count <= WN_MANY, EK_ERR_SHAPE, "wn_bwd_many: count=%d not in [1,%d]", count, WN_MANY);
  WnMany t = {};
  for (int i = 0; i < count; ++i) {
    t.a[i] = dw[i]; t.b[i] = v[i]; t.g[i] = g[i]; t.out[i] = dv[i]; t.out1[i] = dg[i]; t.n[i] = n[i];
  }
  ek_launch(wn_many_partial_kernel, dim3(WN_BLOCKS, count), 256, 0, st, t, workspace);
  EK_CHECK_LAUNCH();
  ek_launch(wn_many_bwd_kernel, dim3(296, count), 256, 0, st, t, workspace, norms);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_wn_bwd_launch(const float* dw, const float* v, const float* g, const float* norm, long long n, float* dv,
                     float* dg, float* workspace, cudaStream_t st) {
  ek_launch(wn_partial_kernel, WN_BLOCKS, 256, 0, st, dw, v, n, workspace);
  EK_CHECK_LAUNCH();
  ek_launch(wn_bwd_kernel, grid_for(n), 256, 0, st, dw, v, g, norm, workspace, n, dv, dg);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
