// Skinny GEMM for the answer decoder's per-step products: C[M <= 64, N] = A[M, K] * op(B), 16-bit operands, fp32 accumulate.
//
// One decode step multiplies the 64 (batch) rows of a recurrent state by L2-resident weights five times in a row
// (dynamic_speaker_change_pos.py:94-131).  A 128-row tcgen05 tile is half empty there and its fixed costs (TMEM allocation,
// mbarrier pipeline, tensor-map fetch, 10-warp CTA, staged epilogue) are ~10 us per launch -- as much as the whole product
// should take.  These products are bound by how fast the weights stream out of L2, so this kernel is the lean form:
// 4 warps, warp-level mma.sync m16n8k16 (one 16-row slab per warp), a 3-stage cp.async ring of 64-wide K chunks,
// grid = (N / BN) x splits with BN in {16, 32, 64} chosen so that ~all SMs pull weights at once; split-K partial sums of a
// plain fp32 output meet in fp32 atomics on a zeroed C.  Epilogue: bias, addend, activation, fp32 and / or 16-bit output.
//   B_MN = 0: B stored [N, K] (K contiguous; forward: nn.Linear weight)   B_MN = 1: B stored [K, N] (dgrad)
#include "common.cuh"
#include "epilogue.cuh"
#include <cstdlib>

namespace {

constexpr int SK_BM = 64;
constexpr int SK_BK = 64;
constexpr int SK_STAGES = 3;
constexpr int SK_AP = SK_BK + 8;           // A / K-major B row pitch in elements (144 B: conflict-free ldmatrix)

__device__ __forceinline__ uint32_t sk_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sk_u32(p)));
}
__device__ __forceinline__ void sk_ldsm_x4_t(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sk_u32(p)));
}
template <bool F16>
__device__ __forceinline__ void sk_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sk_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sk_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void sk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void sk_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int BN, int B_MN>
struct SkCfg {
  static constexpr int BP = B_MN ? BN + 8 : SK_AP;                 // B row pitch (elements)
  static constexpr int B_ROWS = B_MN ? SK_BK : BN;
  static constexpr int A_ELEMS = SK_BM * SK_AP;
  static constexpr int B_ELEMS = B_ROWS * BP;
  static constexpr int SMEM = SK_STAGES * (A_ELEMS + B_ELEMS) * 2;
};

template <int BN, int B_MN, bool F16>
__global__ void __launch_bounds__(128)
gemm_skinny_kernel(const bf16* __restrict__ A, long long lda, const bf16* __restrict__ B, long long ldb, int M, int N, int K,
                   int kper, EkEpilogue ep, int atomic) {
  using C = SkCfg<BN, B_MN>;
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t sk_raw[];
  bf16* As = (bf16*)sk_raw;                                   // [STAGES][64][AP]
  bf16* Bs = As + SK_STAGES * C::A_ELEMS;                     // [STAGES][B_ROWS][BP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.y * kper;
  const int k_end = min(K, k_begin + kper);
  const int nchunks = (k_end - k_begin + SK_BK - 1) / SK_BK;

  auto load_stage = [&](int chunk, int stage) {
    const int k0 = k_begin + chunk * SK_BK;
    bf16* as = As + stage * C::A_ELEMS;
    bf16* bs = Bs + stage * C::B_ELEMS;
    // A chunk: 64 rows x 8 sixteen-byte pieces
#pragma unroll
    for (int it = 0; it < (SK_BM * 8) / 128; ++it) {
      const int e = tid + it * 128;
      const int r = e >> 3, c = (e & 7) * 8;
      bf16* dst = as + r * SK_AP + c;
      if (r < M && k0 + c < k_end) sk_cp16(dst, A + (long long)r * lda + k0 + c);
      else *(uint4*)dst = make_uint4(0, 0, 0, 0);
    }
    if (B_MN == 0) {
      // B chunk: BN rows (n) x 8 pieces along k
      for (int e = tid; e < BN * 8; e += 128) {
        const int r = e >> 3, c = (e & 7) * 8;
        bf16* dst = bs + r * C::BP + c;
        if (n0 + r < N && k0 + c < k_end) sk_cp16(dst, B + (long long)(n0 + r) * ldb + k0 + c);
        else *(uint4*)dst = make_uint4(0, 0, 0, 0);
      }
    } else {
      // B chunk: 64 rows (k) x BN / 8 pieces along n
      constexpr int PPR = BN / 8;
      for (int e = tid; e < SK_BK * PPR; e += 128) {
        const int r = e / PPR, c = (e % PPR) * 8;
        bf16* dst = bs + r * C::BP + c;
        if (k0 + r < k_end && n0 + c < N) sk_cp16(dst, B + (long long)(k0 + r) * ldb + n0 + c);
        else *(uint4*)dst = make_uint4(0, 0, 0, 0);
      }
    }
  };

  constexpr int NT = BN / 8;
  float acc[NT][4];
#pragma unroll
  for (int a = 0; a < NT; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;

#pragma unroll
  for (int s = 0; s < SK_STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    sk_commit();
  }
#pragma unroll 1
  for (int ch = 0; ch < nchunks; ++ch) {
    sk_wait<SK_STAGES - 2>();
    __syncthreads();
    {
      const int nx = ch + SK_STAGES - 1;
      if (nx < nchunks) load_stage(nx, nx % SK_STAGES);
      sk_commit();
    }
    const bf16* as = As + (ch % SK_STAGES) * C::A_ELEMS;
    const bf16* bs = Bs + (ch % SK_STAGES) * C::B_ELEMS;
#pragma unroll
    for (int kt = 0; kt < SK_BK / 16; ++kt) {
      uint32_t af[4];
      sk_ldsm_x4(af, as + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SK_AP + kt * 16 + (lane >> 4) * 8);
      if (NT >= 2) {
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t bfr[4];
          if (B_MN == 0)      // B(k, n) = Bs[n][k]: stored [n][k] -> plain ldmatrix
            sk_ldsm_x4(bfr, bs + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * C::BP + kt * 16 + ((lane >> 3) & 1) * 8);
          else                // stored [k][n] -> transposing ldmatrix
            sk_ldsm_x4_t(bfr, bs + (kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * C::BP + np * 16 + (lane >> 4) * 8);
          sk_mma<F16>(acc[2 * np], af, bfr[0], bfr[1]);
          sk_mma<F16>(acc[2 * np + 1], af, bfr[2], bfr[3]);
        }
      }
    }
  }
  sk_wait<0>();

  // epilogue: this lane holds rows warp*16 + (lane >> 2) + 8*hh, columns n0 + nt*8 + 2*(lane & 3) + {0, 1}
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int r = warp * 16 + (lane >> 2) + hh * 8;
    if (r >= M) continue;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int n = n0 + nt * 8 + 2 * (lane & 3);
      if (n >= N) continue;
      float v0 = acc[nt][2 * hh], v1 = acc[nt][2 * hh + 1];
      const bool two = n + 1 < N;
      if (atomic) {
        atomicAdd(ep.C + (long long)r * ep.ldc + n, v0);
        if (two) atomicAdd(ep.C + (long long)r * ep.ldc + n + 1, v1);
        continue;
      }
      if (ep.bias) { v0 += __ldg(ep.bias + n); if (two) v1 += __ldg(ep.bias + n + 1); }
      if (ep.addend) {
        v0 += ep.addend[(long long)r * ep.ldadd + n];
        if (two) v1 += ep.addend[(long long)r * ep.ldadd + n + 1];
      }
      v0 = ek_act(v0, ep.act);
      v1 = ek_act(v1, ep.act);
      if (ep.C) {
        float* pc = ep.C + (long long)r * ep.ldc + n;
        if (two && (((uintptr_t)pc & 7) == 0)) *(float2*)pc = make_float2(v0, v1);
        else { pc[0] = v0; if (two) pc[1] = v1; }
      }
      if (ep.Cb) {
        bf16* pb = ep.Cb + (long long)r * ep.ldcb + n;
        if (two && (((uintptr_t)pb & 3) == 0)) *(uint32_t*)pb = pack16x2(v0, v1, ep.cb_fmt);
        else { ek_store16(pb, v0, ep.cb_fmt); if (two) ek_store16(pb + 1, v1, ep.cb_fmt); }
      }
    }
  }
}

template <int BN, int B_MN, bool F16>
int sk_launch(const bf16* A, long long lda, const bf16* B, long long ldb, int M, int N, int K, int splits,
              const EkEpilogue& ep, cudaStream_t st) {
  using C = SkCfg<BN, B_MN>;
  auto kern = gemm_skinny_kernel<BN, B_MN, F16>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gemm_skinny: cannot set smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int kper = ((K + splits - 1) / splits + SK_BK - 1) / SK_BK * SK_BK;
  splits = (K + kper - 1) / kper;
  if (splits > 1) {
    cudaError_t e = cudaMemset2DAsync(ep.C, (size_t)ep.ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gemm_skinny: memset failed: %s", cudaGetErrorString(e));
  }
  dim3 grid((N + BN - 1) / BN, splits);
  ek_launch(kern, grid, 128, C::SMEM, st, A, lda, B, ldb, M, N, K, kper, ep, splits > 1 ? 1 : 0);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

template <int B_MN, bool F16>
int sk_dispatch_bn(int bn, const bf16* A, long long lda, const bf16* B, long long ldb, int M, int N, int K, int splits,
                   const EkEpilogue& ep, cudaStream_t st) {
  switch (bn) {
    case 16: return sk_launch<16, B_MN, F16>(A, lda, B, ldb, M, N, K, splits, ep, st);
    case 32: return sk_launch<32, B_MN, F16>(A, lda, B, ldb, M, N, K, splits, ep, st);
    default: return sk_launch<64, B_MN, F16>(A, lda, B, ldb, M, N, K, splits, ep, st);
  }
}

int sk_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

// 1 when the skinny kernel takes this product (M <= 64, A row-major [M, K], same 16-bit format on both operands, 16-byte
// aligned rows, epilogue = bias / addend / activation / C / Cb only), 0 when it belongs to the tcgen05 kernel.
int ek_gemm_skinny_ok(int transA, int transB, int M, int N, int K, const void* A, long long lda, const void* B, long long ldb,
                      const EkEpilogue& ep, int fmt) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("EKAID_B200_SKINNY");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled || transA || M > SK_BM || M < 1 || N < 8 || K < 8) return 0;
  if ((fmt & 1) != ((fmt >> 1) & 1)) return 0;                          // one MMA takes one 16-bit format
  if ((K % 8) || (lda % 8) || (ldb % 8) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return 0;
  if (transB && (N % 8)) return 0;
  if (ep.rowb || ep.drop.seed || ep.Cb2 || ep.cb_n1 || ep.cb2_n0 || ep.C2 || ep.add_n1) return 0;
  if (!ep.C && !ep.Cb) return 0;
  return 1;
}

static long long g_skinny_launches = 0;
long long ek_gemm_skinny_count() { return g_skinny_launches; }

int ek_gemm_skinny_launch(int transB, int M, int N, int K, const bf16* A, long long lda, const bf16* B, long long ldb,
                          const EkEpilogue& ep, int fmt, int splits_req, cudaStream_t st) {
  ++g_skinny_launches;
  // split-K only for a plain fp32 output (partial sums are added onto a zeroed C)
  const bool plain = ep.C && !ep.Cb && !ep.bias && !ep.addend && ep.act == EK_ACT_NONE;
  // cost model: every CTA streams (64 + BN) x Kper 16-bit elements out of L2 at ~80 GB/s (one SM's share of the L2
  // bandwidth); CTAs beyond the SM count run in a second wave; split-K pays a memset and the atomics
  const int sms = sk_num_sms();
  int best_bn = 64, best_s = 1;
  double best = 1e30;
  const int cand_bn[3] = {64, 32, 16};
  const int cand_s[6] = {1, 2, 4, 8, 13, 26};
  for (int bi = 0; bi < 3; ++bi)
    for (int si = 0; si < 6; ++si) {
      const int bn = cand_bn[bi], s = cand_s[si];
      if (s > 1 && (!plain || splits_req == 1 || K / s < 2 * SK_BK)) continue;
      const long long ctas = (long long)((N + bn - 1) / bn) * s;
      const double waves = (double)((ctas + sms - 1) / sms);
      const double kper = (double)K / s;
      const double us = waves * ((SK_BM + bn) * kper * 2.0 / 80e3 + 0.6) + (s > 1 ? 1.5 : 0.0);
      if (us < best) { best = us; best_bn = bn; best_s = s; }
    }
  const bool f16 = (fmt & 1) != 0;
  if (transB) {
    if (f16) return sk_dispatch_bn<1, true>(best_bn, A, lda, B, ldb, M, N, K, best_s, ep, st);
    return sk_dispatch_bn<1, false>(best_bn, A, lda, B, ldb, M, N, K, best_s, ep, st);
  }
  if (f16) return sk_dispatch_bn<0, true>(best_bn, A, lda, B, ldb, M, N, K, best_s, ep, st);
  return sk_dispatch_bn<0, false>(best_bn, A, lda, B, ldb, M, N, K, best_s, ep, st);
}
