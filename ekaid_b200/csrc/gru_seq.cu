// Persistent GRU recurrence of the question path (language_model.py:106-115 forward_all, and its BPTT) on the bf16
// tensor-core path.  One launch runs all L time steps instead of L x (GEMM + cell kernel):
//
//   * the hidden state is split over the grid: CTA c owns GU = 8 hidden units j0..j0+7 and keeps the matching slice
//     of W_hh resident in shared memory for the whole sequence (forward: the 3 x 8 gate rows, 48 KB; backward: the
//     8 columns, as 8 K-major rows of W_hh^T, 48 KB);
//   * per step every CTA streams the full broadcast operand (h_{t-1} [B,H] forward, dgh_{t+1} [B,3H] backward; bf16,
//     L2 resident) through a double-buffered cp.async ring, multiplies it with its weight slice on warp-level
//     mma.sync m16n8k16 (fp32 accumulate; M = batch is tiny, so this is a skinny GEMM and tcgen05's 128-row tile
//     would be 50 % empty), applies the GRU cell (or its derivative) to its own units and writes its slice of the
//     outputs;
//   * a grid-wide barrier (one atomic counter in global memory, release/acquire) separates the steps.  The grid is
//     H / 8 = 128 CTAs of one CTA per SM, so all CTAs are co-resident on a B200; a barrier wait is bounded and traps
//     instead of hanging.
//
// Arithmetic is that of gru_cell_fwd/bwd_kernel + the bf16 GEMMs they were paired with (question.cu), so the two
// formulations agree to fp32 summation order.
#include "common.cuh"

namespace {

constexpr int GU = 8;             // hidden units per CTA
constexpr int KC = 512;           // k elements of the broadcast operand per staged chunk
constexpr int RB = 64;            // batch rows per pass (4 MMA row tiles)
constexpr int AP = KC + 8;        // chunk row pitch in elements (1040 B: ldmatrix rows land in distinct 16-byte slots)
constexpr int THREADS = 256;      // 8 warps = 4 row tiles x 2 k halves

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16-byte async copy through L2 only (the source was written by other CTAs earlier in this kernel); nbytes = 0 zero-fills
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, int nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_u32(dst)), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// All CTAs of the grid arrive; returns when `target` arrivals have been counted since the counter was zeroed.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                       // this CTA's global writes are visible before the arrival
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(bar) < target) {
      if (clock64() - t0 > (1ll << 33)) {  // seconds: a CTA of the grid never became resident
        printf("gru_seq: grid barrier timeout (block %d, target %u)\n", blockIdx.x, target);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// stage rows [rb0, rb0+RB) x columns [k0, k0+KC) of the row-major bf16 matrix Ag [rows, K] into buf [RB][AP]
__device__ __forceinline__ void load_chunk(bf16* buf, const bf16* Ag, int rows, int K, int rb0, int k0) {
  constexpr int CPR = KC / 8;              // 16-byte pieces per row
  for (int e = threadIdx.x; e < RB * CPR; e += THREADS) {
    const int r = e / CPR, ch = e % CPR;
    const int row = rb0 + r;
    const bool ok = row < rows;
    const bf16* src = Ag + (size_t)(ok ? row : 0) * K + k0 + ch * 8;
    cp_async16_zfill(buf + r * AP + ch * 8, src, ok ? 16 : 0);
  }
}

// acc[nt][:] += A[rb0 + 16*mt .. +16, :] . Ws[nt*8 .. +8, :]^T  for this warp's k half of every chunk.
// Ws is [NT*8][wp] bf16, K-major (row n = output column n).  All threads of the CTA must call.
template <int NT>
__device__ __forceinline__ void skinny_gemm(const bf16* Ag, int rows, int K, int rb0, bf16* As, const bf16* Ws, int wp,
                                            float (&acc)[NT][4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, kh = warp >> 2;
  const int nchunks = K / KC;
  load_chunk(As, Ag, rows, K, rb0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load_chunk(As + ((c + 1) & 1) * (RB * AP), Ag, rows, K, rb0, (c + 1) * KC);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* a_base = As + (c & 1) * (RB * AP) + (mt * 16 + (lane & 15)) * AP + kh * (KC / 2) + (lane >> 4) * 8;
    const bf16* w_base = Ws + (size_t)(lane >> 2) * wp + c * KC + kh * (KC / 2) + (lane & 3) * 2;
#pragma unroll 4
    for (int ks = 0; ks < KC / 2 / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4(a, a_base + ks * 16);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const bf16* w = w_base + (size_t)nt * 8 * wp + ks * 16;
        mma_bf16_16816(acc[nt], a, *(const uint32_t*)w, *(const uint32_t*)(w + 8));
      }
    }
    __syncthreads();                       // the buffer may be refilled by the next iteration's prefetch
  }
}

// sum the two k halves and lay the [RB x NT*8] result out in shared memory (row pitch NT*8+1 floats)
template <int NT>
__device__ __forceinline__ void reduce_to_smem(float (&acc)[NT][4], float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, kh = warp >> 2;
  constexpr int OP = NT * 8 + 1;
  const int r0 = mt * 16 + (lane >> 2), c0 = (lane & 3) * 2;
  if (kh == 0) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      out[r0 * OP + nt * 8 + c0] = acc[nt][0];
      out[r0 * OP + nt * 8 + c0 + 1] = acc[nt][1];
      out[(r0 + 8) * OP + nt * 8 + c0] = acc[nt][2];
      out[(r0 + 8) * OP + nt * 8 + c0 + 1] = acc[nt][3];
    }
  }
  __syncthreads();
  if (kh == 1) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      out[r0 * OP + nt * 8 + c0] += acc[nt][0];
      out[r0 * OP + nt * 8 + c0 + 1] += acc[nt][1];
      out[(r0 + 8) * OP + nt * 8 + c0] += acc[nt][2];
      out[(r0 + 8) * OP + nt * 8 + c0 + 1] += acc[nt][3];
    }
  }
  __syncthreads();
}

constexpr int EPT = RB * GU / THREADS;     // (row, unit) elements per thread per pass = 2

// ------------------------------------------------------------------------------------------------ forward
// gi [L*B, 3H] fp32 (= x W_ih^T + b_ih, time-major rows t*B + b), Whh [3H, H] bf16, bhh [3H].
// Hs [L*B, H] fp32, HsT [(L+1)*B, H] bf16 with block 0 = h_{-1} = 0 (written by the caller), gates [L, B, 4H] = (r, z, n, gh_n).
__global__ void __launch_bounds__(THREADS, 1)
gru_seq_fwd_kernel(const float* __restrict__ gi, const bf16* __restrict__ Whh, const float* __restrict__ bhh, int B,
                   int H, int L, float* __restrict__ Hs, bf16* HsT, float* __restrict__ gates, unsigned int* bar) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int NT = 3 * GU / 8;
  constexpr int OP = NT * 8 + 1;
  const int wp = H + 8;
  bf16* Ws = (bf16*)smraw;                                  // [3*GU][wp]   row g*GU + u = W_hh[g*H + j0 + u, :]
  bf16* As = Ws + (size_t)3 * GU * wp;                      // 2 x [RB][AP]
  float* ghs = (float*)(As + 2 * RB * AP);                  // [RB][OP]
  float* hprev = ghs + RB * OP;                             // [B][GU] this CTA's slice of h_{t-1} (fp32)
  const int j0 = blockIdx.x * GU;
  const int tid = threadIdx.x;
  {
    const int cpr = H / 8;
    for (int e = tid; e < 3 * GU * cpr; e += THREADS) {
      const int n = e / cpr, ch = e % cpr;
      const int g = n / GU, u = n % GU;
      cp_async16_zfill(Ws + (size_t)n * wp + ch * 8, Whh + ((size_t)g * H + j0 + u) * H + ch * 8, 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    for (int e = tid; e < B * GU; e += THREADS) hprev[e] = 0.f;
    __syncthreads();
  }
  unsigned int arrivals = 0;
  for (int t = 0; t < L; ++t) {
    for (int rb0 = 0; rb0 < B; rb0 += RB) {
      // this thread's (row, unit) elements of the pass; input gates prefetched ahead of the GEMM
      float gir[EPT], giz[EPT], gin[EPT];
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * THREADS;
        const int b = rb0 + e / GU, j = j0 + e % GU;
        if (b < B) {
          const float* gp = gi + ((size_t)t * B + b) * 3 * H + j;
          gir[i] = __ldg(gp); giz[i] = __ldg(gp + H); gin[i] = __ldg(gp + 2 * H);
        } else {
          gir[i] = giz[i] = gin[i] = 0.f;
        }
      }
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      if (t > 0) skinny_gemm<NT>(HsT + (size_t)t * B * H, B, H, rb0, As, Ws, wp, acc);     // h_{t-1} W_hh^T (h_{-1} = 0)
      reduce_to_smem<NT>(acc, ghs);
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * THREADS;
        const int bl = e / GU, u = e % GU;
        const int b = rb0 + bl, j = j0 + u;
        if (b < B) {
          const float ghr = ghs[bl * OP + u] + __ldg(bhh + j);
          const float ghz = ghs[bl * OP + GU + u] + __ldg(bhh + H + j);
          const float ghn = ghs[bl * OP + 2 * GU + u] + __ldg(bhh + 2 * H + j);
          const float r = sigmoidf_(gir[i] + ghr);
          const float z = sigmoidf_(giz[i] + ghz);
          const float n = tanhf(gin[i] + r * ghn);
          const float hp = hprev[b * GU + u];
          const float hv = (1.f - z) * n + z * hp;
          hprev[b * GU + u] = hv;
          const size_t row = (size_t)t * B + b;
          Hs[row * H + j] = hv;
          HsT[(row + B) * H + j] = __float2bfloat16_rn(hv);
          float* gs = gates + row * 4 * H + j;
          gs[0] = r; gs[H] = z; gs[2 * H] = n; gs[3 * H] = ghn;
        }
      }
      __syncthreads();                     // ghs is rewritten by the next pass
    }
    if (t + 1 < L) {
      arrivals += gridDim.x;
      grid_barrier(bar, arrivals);         // every CTA's slice of h_t is in HsT before anyone starts step t+1
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward (BPTT)
// dHs [L*B, H]: gradient reaching h_t from outside the recurrence.  Per step (t = L-1 .. 0), for the CTA's units k:
//   dh = dHs[t] + dh_{t+1} * z_{t+1} + dgh_{t+1} . W_hh[:, k];   cell derivative -> dgi[t], dgh[t] (fp32 and bf16 copies).
__global__ void __launch_bounds__(THREADS, 1)
gru_seq_bwd_kernel(const float* __restrict__ dHs, const float* __restrict__ gates, const float* __restrict__ Hs,
                   const bf16* __restrict__ Whh, int B, int H, int L, float* __restrict__ dgi, float* __restrict__ dgh,
                   bf16* __restrict__ dgiT, bf16* dghT, unsigned int* bar) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int NT = GU / 8;
  constexpr int OP = NT * 8 + 1;
  const int K = 3 * H;
  const int wp = K + 8;
  bf16* Ws = (bf16*)smraw;                                  // [GU][wp]   row u = W_hh[:, j0 + u]
  bf16* As = Ws + (size_t)GU * wp;                          // 2 x [RB][AP]
  float* cs = (float*)(As + 2 * RB * AP);                   // [RB][OP]   dgh_{t+1} W_hh for this CTA's units
  float* dhz = cs + RB * OP;                                // [B][GU]    dh_{t+1} * z_{t+1}
  const int j0 = blockIdx.x * GU;
  const int tid = threadIdx.x;
  for (int c = tid; c < K; c += THREADS) {
    const uint4 v = *(const uint4*)(Whh + (size_t)c * H + j0);       // 8 consecutive columns of row c
    const bf16* pv = (const bf16*)&v;
#pragma unroll
    for (int u = 0; u < GU; ++u) Ws[(size_t)u * wp + c] = pv[u];
  }
  for (int e = tid; e < B * GU; e += THREADS) dhz[e] = 0.f;
  __syncthreads();
  unsigned int arrivals = 0;
  for (int t = L - 1; t >= 0; --t) {
    for (int rb0 = 0; rb0 < B; rb0 += RB) {
      float gr[EPT], gz[EPT], gn[EPT], gg[EPT], dd[EPT], hp[EPT];
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * THREADS;
        const int b = rb0 + e / GU, j = j0 + e % GU;
        if (b < B) {
          const size_t row = (size_t)t * B + b;
          const float* gs = gates + row * 4 * H + j;
          gr[i] = __ldg(gs); gz[i] = __ldg(gs + H); gn[i] = __ldg(gs + 2 * H); gg[i] = __ldg(gs + 3 * H);
          dd[i] = __ldg(dHs + row * H + j);
          hp[i] = t > 0 ? __ldg(Hs + (row - B) * H + j) : 0.f;
        } else {
          gr[i] = gz[i] = gn[i] = gg[i] = dd[i] = hp[i] = 0.f;
        }
      }
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      if (t < L - 1) skinny_gemm<NT>(dghT + (size_t)(t + 1) * B * K, B, K, rb0, As, Ws, wp, acc);
      reduce_to_smem<NT>(acc, cs);
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * THREADS;
        const int bl = e / GU, u = e % GU;
        const int b = rb0 + bl, j = j0 + u;
        if (b < B) {
          const float r = gr[i], z = gz[i], n = gn[i], ghn = gg[i];
          const float d = dd[i] + dhz[b * GU + u] + cs[bl * OP + u];
          const float dn = d * (1.f - z);
          const float dz = d * (hp[i] - n);
          const float dnp = dn * (1.f - n * n);
          const float drp = dnp * ghn * r * (1.f - r);
          const float dzp = dz * z * (1.f - z);
          dhz[b * GU + u] = d * z;
          const size_t o = ((size_t)t * B + b) * K + j;
          dgi[o] = drp; dgi[o + H] = dzp; dgi[o + 2 * H] = dnp;
          dgh[o] = drp; dgh[o + H] = dzp; dgh[o + 2 * H] = dnp * r;
          dgiT[o] = __float2bfloat16_rn(drp); dgiT[o + H] = __float2bfloat16_rn(dzp);
          dgiT[o + 2 * H] = __float2bfloat16_rn(dnp);
          dghT[o] = __float2bfloat16_rn(drp); dghT[o + H] = __float2bfloat16_rn(dzp);
          dghT[o + 2 * H] = __float2bfloat16_rn(dnp * r);
        }
      }
      __syncthreads();
    }
    if (t > 0) {
      arrivals += gridDim.x;
      grid_barrier(bar, arrivals);         // dgh_t of every unit is in dghT before anyone starts step t-1
    }
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int check_shape(const char* who, int B, int H, int L) {
  EK_REQUIRE(B > 0 && L > 0 && H > 0, EK_ERR_SHAPE, "%s: bad shape B=%d H=%d L=%d", who, B, H, L);
  EK_REQUIRE(H % KC == 0, EK_ERR_UNSUPPORTED, "%s: H=%d must be a multiple of %d", who, H, KC);
  EK_REQUIRE(H / GU <= sm_count(), EK_ERR_UNSUPPORTED,
             "%s: H/%d = %d CTAs must be co-resident (one per SM, %d SMs)", who, GU, H / GU, sm_count());
  return EK_OK;
}

}  // namespace

int ek_gru_seq_fwd_launch(const float* gi, const bf16* Whh, const float* bhh, int B, int H, int L, float* Hs, bf16* HsT,
                          float* gates, unsigned int* bar, cudaStream_t st) {
  int rc = check_shape("gru_seq_fwd", B, H, L);
  if (rc) return rc;
  const size_t smem = (size_t)3 * GU * (H + 8) * 2 + (size_t)2 * RB * AP * 2 + (size_t)RB * (3 * GU + 1) * 4 +
                      (size_t)B * GU * 4;
  EK_REQUIRE(smem <= 227 * 1024, EK_ERR_UNSUPPORTED, "gru_seq_fwd: B=%d H=%d needs %zu bytes of shared memory", B, H, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(gru_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gru_seq_fwd: cannot set smem attr: %s", cudaGetErrorString(e));
    attr = smem;
  }
  cudaMemsetAsync(bar, 0, sizeof(unsigned int), st);
  ek_launch(gru_seq_fwd_kernel, H / GU, THREADS, smem, st, gi, Whh, bhh, B, H, L, Hs, HsT, gates, bar);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_gru_seq_bwd_launch(const float* dHs, const float* gates, const float* Hs, const bf16* Whh, int B, int H, int L,
                          float* dgi, float* dgh, bf16* dgiT, bf16* dghT, unsigned int* bar, cudaStream_t st) {
  int rc = check_shape("gru_seq_bwd", B, H, L);
  if (rc) return rc;
  const size_t smem = (size_t)GU * (3 * H + 8) * 2 + (size_t)2 * RB * AP * 2 + (size_t)RB * (GU + 1) * 4 +
                      (size_t)B * GU * 4;
  EK_REQUIRE(smem <= 227 * 1024, EK_ERR_UNSUPPORTED, "gru_seq_bwd: B=%d H=%d needs %zu bytes of shared memory", B, H, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(gru_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gru_seq_bwd: cannot set smem attr: %s", cudaGetErrorString(e));
    attr = smem;
  }
  cudaMemsetAsync(bar, 0, sizeof(unsigned int), st);
  ek_launch(gru_seq_bwd_kernel, H / GU, THREADS, smem, st, dHs, gates, Hs, Whh, B, H, L, dgi, dgh, dgiT, dghT, bar);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
