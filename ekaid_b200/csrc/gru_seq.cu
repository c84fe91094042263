// Persistent GRU recurrence of the question path (language_model.py:106-115 forward_all, and its BPTT) on the bf16
// tensor-core path.  One launch runs all L time steps instead of L x (GEMM + cell kernel):
//
//   * the hidden state is split over the grid: CTA c owns GU = 16 hidden units; four neighbouring CTAs form a
//     thread-block cluster that owns 64 units together.  Inside a cluster the reduction dimension of the per-step
//     product is split four ways: CTA r multiplies ITS quarter of the broadcast operand (h_{t-1} [B,H] forward,
//     dgh_{t+1} [B,3H] backward; bf16, L2 resident) with the matching quarter of the cluster's W_hh slice, which stays
//     in its shared memory for the whole sequence (~100 KB), on warp-level mma.sync m16n8k16 (fp32 accumulate; M =
//     batch is tiny, a tcgen05 128-row tile would be half empty).  So each SM pulls only a quarter of the operand per
//     step, and the grid is 64 CTAs: 84 SMs stay free for whatever runs next to the recurrence.
//   * the partial sums are pushed into the owner CTA's shared memory through distributed shared memory
//     (st.shared::cluster), one cluster barrier later the owner applies the GRU cell (or its derivative) to its 16 units
//     and writes its slice of the outputs;
//   * a grid-wide barrier (one atomic counter in global memory, release/acquire) separates the steps.  All 64 CTAs are
//     co-resident (checked with cudaOccupancyMaxActiveClusters); a barrier wait is bounded and traps instead of hanging.
//
// Arithmetic is that of gru_cell_fwd/bwd_kernel + the bf16 GEMMs they were paired with (question.cu), so the two
// formulations agree to fp32 summation order.
#include <cstdlib>
#include "common.cuh"

namespace {

// Template parameters of the kernels: GU = hidden units per CTA, CS = CTAs per cluster (= split of the reduction
// dimension).  <16, 4>: 64 CTAs, least L2 traffic, leaves 84 SMs free (backward, which runs next to the optimizer).
// <8, 2>: 128 CTAs, half the mma.sync work per SM (forward, which is on the critical path of the step).
constexpr int KC = 256;           // k elements of the broadcast operand per staged chunk
constexpr int RB = 64;            // batch rows per pass (4 MMA row tiles)
constexpr int AP = KC + 8;        // chunk row pitch in elements (528 B: ldmatrix rows land in distinct 16-byte slots)
constexpr int THREADS = 256;      // 8 warps = 4 row tiles x 2 column halves

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
__device__ __forceinline__ void mma_f16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
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

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory location of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(s_u32(p)), "r"(rank));
  return a;
}
__device__ __forceinline__ void dsmem_st2(uint32_t addr, float x, float y) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

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

// Leaves the barrier words zero for the next launch: every CTA checks out after its last barrier wait (word 1), the last
// one clears both words.  A cudaMemsetAsync in front of every launch did the same, as a memset node between two kernel
// nodes on the serial question path of the captured step.
__device__ __forceinline__ void grid_barrier_retire(unsigned int* bar) {
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(bar + 1, 1u) == gridDim.x - 1) {
      bar[0] = 0u;
      bar[1] = 0u;
      __threadfence();
    }
  }
}

// stage rows [rb0, rb0+RB) x columns [k0, k0+KC) of the row-major bf16 matrix Ag [rows, ld] into buf [RB][AP]
__device__ __forceinline__ void load_chunk(bf16* buf, const bf16* Ag, int rows, int ld, int rb0, int k0) {
  constexpr int CPR = KC / 8;              // 16-byte pieces per row
  for (int e = threadIdx.x; e < RB * CPR; e += THREADS) {
    const int r = e / CPR, ch = e % CPR;
    const int row = rb0 + r;
    const bool ok = row < rows;
    const bf16* src = Ag + (size_t)(ok ? row : 0) * ld + k0 + ch * 8;
    cp_async16_zfill(buf + r * AP + ch * 8, src, ok ? 16 : 0);
  }
}

// acc[j][:] += A[rb0 + 16*mt .. +16, kbeg .. kbeg + nchunks*KC) . Ws[(nh*NTW + j)*8 .. +8, 0 .. nchunks*KC)^T
// warp = (mt = warp & 3, nh = warp >> 2).  Ws is [2*NTW*8][wp] bf16, K-major.  All threads of the CTA must call.
template <int NTW, bool F16 = false>     // F16: both operands hold IEEE fp16 (forward: h in [-1, 1], weights), else bf16
__device__ __forceinline__ void skinny_gemm(const bf16* Ag, int rows, int ld, int rb0, int kbeg, int nchunks, bf16* As,
                                            const bf16* Ws, int wp, float (&acc)[NTW][4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, nh = warp >> 2;
  load_chunk(As, Ag, rows, ld, rb0, kbeg);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load_chunk(As + ((c + 1) & 1) * (RB * AP), Ag, rows, ld, rb0, kbeg + (c + 1) * KC);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* a_base = As + (c & 1) * (RB * AP) + (mt * 16 + (lane & 15)) * AP + (lane >> 4) * 8;
    const bf16* w_base = Ws + (size_t)(nh * NTW * 8 + (lane >> 2)) * wp + c * KC + (lane & 3) * 2;
#pragma unroll 4
    for (int ks = 0; ks < KC / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4(a, a_base + ks * 16);
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const bf16* w = w_base + (size_t)j * 8 * wp + ks * 16;
        if (F16) mma_f16_16816(acc[j], a, *(const uint32_t*)w, *(const uint32_t*)(w + 8));
        else mma_bf16_16816(acc[j], a, *(const uint32_t*)w, *(const uint32_t*)(w + 8));
      }
    }
    __syncthreads();                       // the buffer may be refilled by the next iteration's prefetch
  }
}

// Push this warp's partial sums to the owners: cluster column n (0 .. CS*OW) belongs to CTA n / OW, local column
// n % OW; the owner keeps one [RB][OW + 1] fp32 slab per source rank in `recv`.
template <int NTW, int OW>
__device__ __forceinline__ void push_partials(float (&acc)[NTW][4], float* recv, uint32_t my_rank) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, nh = warp >> 2;
  constexpr int RP = OW + 2;               // slab row pitch (even: 8-byte aligned pairs)
  const int r0 = mt * 16 + (lane >> 2);
#pragma unroll
  for (int j = 0; j < NTW; ++j) {
    const int n = (nh * NTW + j) * 8 + (lane & 3) * 2;       // cluster column of acc[j][0]
    const uint32_t owner = n / OW;
    const int cl = n % OW;
    float* slot = recv + ((size_t)my_rank * RB + r0) * RP + cl;
    const uint32_t a0 = dsmem_addr(slot, owner);
    dsmem_st2(a0, acc[j][0], acc[j][1]);
    dsmem_st2(a0 + 8 * RP * 4, acc[j][2], acc[j][3]);
  }
}

// ------------------------------------------------------------------------------------------------ forward
// gi [L*B, 3H] fp32 (= x W_ih^T + b_ih, time-major rows t*B + b), Whh [3H, H] bf16, bhh [3H].
// Hs [L*B, H] fp32, HsT [(L+1)*B, H] bf16 with block 0 = h_{-1} = 0 (written by the caller), gates [L, B, 4H] = (r, z, n, gh_n).
// F16: Whh and HsT hold IEEE fp16 (h is bounded by 1, so fp16 only adds mantissa bits); HsB (optional): a bf16 copy of
// HsT for the backward pass, whose wgrad pairs it with bf16 gradients (one MMA takes one 16-bit format for both operands)
template <int GU, int CS, bool F16>
__global__ void __launch_bounds__(THREADS, 1)
gru_seq_fwd_kernel(const float* __restrict__ gi, const bf16* __restrict__ Whh, const float* __restrict__ bhh, int B,
                   int H, int L, float* __restrict__ Hs, bf16* HsT, float* __restrict__ gates, unsigned int* bar,
                   bf16* __restrict__ HsB) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int CU = GU * CS;              // hidden units per cluster
  constexpr int EPT = RB * GU / THREADS;   // (row, unit) elements per thread per pass
  constexpr int OW = 3 * GU;               // columns owned per CTA: (gate, unit)
  constexpr int NTW = CS * OW / 8 / 2;     // n-tiles per warp (two column halves)
  static_assert((CS * OW / 8) % 2 == 0, "the cluster's columns split into two warp halves");
  constexpr int RP = OW + 2;
  const int KS = H / CS;                   // this CTA's slice of the reduction dimension
  const int wp = KS + 8;
  bf16* Ws = (bf16*)smraw;                                  // [CS*OW][wp]: row = owner*OW + g*GU + u
  const int nbuf = (KS / KC > 1) ? 2 : 1;
  bf16* As = Ws + (size_t)CS * OW * wp;                     // nbuf x [RB][AP]
  float* recv = (float*)(As + nbuf * RB * AP);              // [CS][RB][RP] partial sums pushed by the cluster
  float* hprev = recv + CS * RB * RP;                       // [B][GU] this CTA's slice of h_{t-1} (fp32)
  const uint32_t rank = cluster_rank();
  const int jc0 = (blockIdx.x / CS) * CU;                   // first unit of the cluster
  const int j0 = jc0 + (int)rank * GU;                      // first unit of this CTA
  const int tid = threadIdx.x;
  {
    const int cpr = KS / 8;
    for (int e = tid; e < CS * OW * cpr; e += THREADS) {
      const int n = e / cpr, ch = e % cpr;
      const int owner = n / OW, g = (n % OW) / GU, u = n % GU;
      cp_async16_zfill(Ws + (size_t)n * wp + ch * 8,
                       Whh + ((size_t)g * H + jc0 + owner * GU + u) * H + (size_t)rank * KS + ch * 8, 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    for (int e = tid; e < B * GU; e += THREADS) hprev[e] = 0.f;
    for (int e = tid; e < CS * RB * RP; e += THREADS) recv[e] = 0.f;
    __syncthreads();
  }
  cluster_sync();                          // every CTA of the cluster is running before anyone pushes into it
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
      if (t > 0) {                         // h_{-1} = 0: the first step has no recurrent term (recv stays zero)
        float acc[NTW][4];
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        skinny_gemm<NTW, F16>(HsT + (size_t)t * B * H, B, H, rb0, (int)rank * KS, KS / KC, As, Ws, wp, acc);
        push_partials<NTW, OW>(acc, recv, rank);
        cluster_sync();
      }
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * THREADS;
        const int bl = e / GU, u = e % GU;
        const int b = rb0 + bl, j = j0 + u;
        if (b < B) {
          float ghr = __ldg(bhh + j), ghz = __ldg(bhh + H + j), ghn = __ldg(bhh + 2 * H + j);
#pragma unroll
          for (int s = 0; s < CS; ++s) {
            const float* rs = recv + ((size_t)s * RB + bl) * RP;
            ghr += rs[u]; ghz += rs[GU + u]; ghn += rs[2 * GU + u];
          }
          const float r = sigmoidf_(gir[i] + ghr);
          const float z = sigmoidf_(giz[i] + ghz);
          const float n = tanhf(gin[i] + r * ghn);
          const float hp = hprev[b * GU + u];
          const float hv = (1.f - z) * n + z * hp;
          hprev[b * GU + u] = hv;
          const size_t row = (size_t)t * B + b;
          Hs[row * H + j] = hv;
          if (F16) ((f16*)HsT)[(row + B) * H + j] = from_f32<f16>(hv);
          else HsT[(row + B) * H + j] = __float2bfloat16_rn(hv);
          if (HsB) HsB[(row + B) * H + j] = __float2bfloat16_rn(hv);
          float* gs = gates + row * 4 * H + j;
          gs[0] = r; gs[H] = z; gs[2 * H] = n; gs[3 * H] = ghn;
        }
      }
      if (B > RB) cluster_sync();          // recv is rewritten by the next pass of this step
    }
    if (t + 1 < L) {
      arrivals += gridDim.x;
      grid_barrier(bar, arrivals);         // every CTA's slice of h_t is in HsT before anyone starts step t+1
    }
  }
  cluster_sync();                          // no CTA leaves while a peer may still push into it
  grid_barrier_retire(bar);
}

// ------------------------------------------------------------------------------------------------ backward (BPTT)
// dHs [L*B, H]: gradient reaching h_t from outside the recurrence.  Per step (t = L-1 .. 0), for the CTA's units k:
//   dh = dHs[t] + dh_{t+1} * z_{t+1} + dgh_{t+1} . W_hh[:, k];   cell derivative -> dgi[t], dgh[t] (fp32 and bf16 copies).
template <int GU, int CS>
__global__ void __launch_bounds__(THREADS, 1)
gru_seq_bwd_kernel(const float* __restrict__ dHs, const float* __restrict__ gates, const float* __restrict__ Hs,
                   const bf16* __restrict__ Whh, int B, int H, int L, float* __restrict__ dgi, float* __restrict__ dgh,
                   bf16* __restrict__ dgiT, bf16* dghT, unsigned int* bar) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int CU = GU * CS;
  constexpr int EPT = RB * GU / THREADS;
  constexpr int OW = GU;
  constexpr int NTW = CS * OW / 8 / 2;
  static_assert((CS * OW / 8) % 2 == 0, "the cluster's columns split into two warp halves");
  constexpr int RP = OW + 2;
  const int K = 3 * H;
  const int KS = K / CS;
  const int wp = KS + 8;
  bf16* Ws = (bf16*)smraw;                                  // [CU][wp]: row n = W_hh[rank*KS .. +KS, jc0 + n]
  bf16* As = Ws + (size_t)CU * wp;                          // 2 x [RB][AP]
  float* recv = (float*)(As + 2 * RB * AP);                 // [CS][RB][RP]
  float* dhz = recv + CS * RB * RP;                         // [B][GU]    dh_{t+1} * z_{t+1}
  const uint32_t rank = cluster_rank();
  const int jc0 = (blockIdx.x / CS) * CU;
  const int j0 = jc0 + (int)rank * GU;
  const int tid = threadIdx.x;
  for (int e = tid; e < KS * (CU / 8); e += THREADS) {
    const int cl = e / (CU / 8), c8 = e % (CU / 8);
    const uint4 v = *(const uint4*)(Whh + ((size_t)rank * KS + cl) * H + jc0 + c8 * 8);      // 8 consecutive columns
    const bf16* pv = (const bf16*)&v;
#pragma unroll
    for (int u = 0; u < 8; ++u) Ws[(size_t)(c8 * 8 + u) * wp + cl] = pv[u];
  }
  for (int e = tid; e < B * GU; e += THREADS) dhz[e] = 0.f;
  for (int e = tid; e < CS * RB * RP; e += THREADS) recv[e] = 0.f;
  __syncthreads();
  cluster_sync();
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
      if (t < L - 1) {
        float acc[NTW][4];
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        skinny_gemm<NTW>(dghT + (size_t)(t + 1) * B * K, B, K, rb0, (int)rank * KS, KS / KC, As, Ws, wp, acc);
        push_partials<NTW, OW>(acc, recv, rank);
        cluster_sync();
      }
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * THREADS;
        const int bl = e / GU, u = e % GU;
        const int b = rb0 + bl, j = j0 + u;
        if (b < B) {
          const float r = gr[i], z = gz[i], n = gn[i], ghn = gg[i];
          float d = dd[i] + dhz[b * GU + u];
#pragma unroll
          for (int s = 0; s < CS; ++s) d += recv[((size_t)s * RB + bl) * RP + u];
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
      if (B > RB) cluster_sync();
    }
    if (t > 0) {
      arrivals += gridDim.x;
      grid_barrier(bar, arrivals);         // dgh_t of every unit is in dghT before anyone starts step t-1
    }
  }
  cluster_sync();
  grid_barrier_retire(bar);
}

// EKAID_B200_GRU_MEMSET=1: clear the barrier words in front of every launch as well (the kernels leave them zero)
bool barrier_memset() {
  static const bool on = [] { const char* e = getenv("EKAID_B200_GRU_MEMSET"); return e && atoi(e) != 0; }();
  return on;
}

int check_shape(const char* who, int B, int H, int L) {
  EK_REQUIRE(B > 0 && L > 0 && H > 0, EK_ERR_SHAPE, "%s: bad shape B=%d H=%d L=%d", who, B, H, L);
  EK_REQUIRE(H % 1024 == 0, EK_ERR_UNSUPPORTED, "%s: H=%d must be a multiple of 1024", who, H);
  return EK_OK;
}

template <int CS, typename Kern, typename... Args>
int launch_clustered(const char* who, Kern kern, int grid, size_t smem, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "%s: cannot set smem attr (%zu bytes): %s", who, smem, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = ek_pdl_enabled();
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // the grid barrier needs every cluster resident at the same time
  int max_clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "%s: cudaOccupancyMaxActiveClusters: %s", who, cudaGetErrorString(e));
  EK_REQUIRE(max_clusters >= grid / CS, EK_ERR_UNSUPPORTED, "%s: %d clusters of %d CTAs needed, %d fit on this device", who,
             grid / CS, CS, max_clusters);
  cfg.numAttrs = 2;
  e = cudaLaunchKernelEx(&cfg, kern, args...);
  EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "%s: launch failed: %s", who, cudaGetErrorString(e));
  return EK_OK;
}

template <int GU, int CS>
int fwd_launch(const float* gi, const bf16* Whh, const float* bhh, int B, int H, int L, float* Hs, bf16* HsT, float* gates,
               unsigned int* bar, int f16_ops, bf16* HsB, cudaStream_t st) {
  const size_t nbuf = (H / CS / KC > 1) ? 2 : 1;
  const size_t smem = (size_t)CS * 3 * GU * (H / CS + 8) * 2 + nbuf * RB * AP * 2 +
                      (size_t)CS * RB * (3 * GU + 2) * 4 + (size_t)B * GU * 4;
  EK_REQUIRE(smem <= 227 * 1024, EK_ERR_UNSUPPORTED, "gru_seq_fwd: B=%d H=%d needs %zu bytes of shared memory", B, H, smem);
  if (barrier_memset()) cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st);
  if (f16_ops)
    return launch_clustered<CS>("gru_seq_fwd", gru_seq_fwd_kernel<GU, CS, true>, H / GU, smem, st, gi, Whh, bhh, B, H, L, Hs,
                                HsT, gates, bar, HsB);
  return launch_clustered<CS>("gru_seq_fwd", gru_seq_fwd_kernel<GU, CS, false>, H / GU, smem, st, gi, Whh, bhh, B, H, L, Hs,
                              HsT, gates, bar, HsB);
}
template <int GU, int CS>
int bwd_launch(const float* dHs, const float* gates, const float* Hs, const bf16* Whh, int B, int H, int L, float* dgi,
               float* dgh, bf16* dgiT, bf16* dghT, unsigned int* bar, cudaStream_t st) {
  const size_t smem = (size_t)GU * CS * (3 * H / CS + 8) * 2 + (size_t)2 * RB * AP * 2 + (size_t)CS * RB * (GU + 2) * 4 +
                      (size_t)B * GU * 4;
  EK_REQUIRE(smem <= 227 * 1024, EK_ERR_UNSUPPORTED, "gru_seq_bwd: B=%d H=%d needs %zu bytes of shared memory", B, H, smem);
  if (barrier_memset()) cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st);
  return launch_clustered<CS>("gru_seq_bwd", gru_seq_bwd_kernel<GU, CS>, H / GU, smem, st, dHs, gates, Hs, Whh, B, H, L,
                              dgi, dgh, dgiT, dghT, bar);
}

}  // namespace

// variant: 0 = default (forward <8,2>, backward <16,4>), 1 = <8,2>, 2 = <16,4>   (EKAID_B200_GRU_VARIANT, for measurements)
static int gru_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EKAID_B200_GRU_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int ek_gru_seq_fwd_launch(const float* gi, const bf16* Whh, const float* bhh, int B, int H, int L, float* Hs, bf16* HsT,
                          float* gates, unsigned int* bar, int f16_ops, bf16* HsB, cudaStream_t st) {
  int rc = check_shape("gru_seq_fwd", B, H, L);
  if (rc) return rc;
  if (gru_variant() == 2) rc = fwd_launch<16, 4>(gi, Whh, bhh, B, H, L, Hs, HsT, gates, bar, f16_ops, HsB, st);
  else rc = fwd_launch<8, 2>(gi, Whh, bhh, B, H, L, Hs, HsT, gates, bar, f16_ops, HsB, st);
  if (rc) return rc;
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_gru_seq_bwd_launch(const float* dHs, const float* gates, const float* Hs, const bf16* Whh, int B, int H, int L,
                          float* dgi, float* dgh, bf16* dgiT, bf16* dghT, unsigned int* bar, cudaStream_t st) {
  int rc = check_shape("gru_seq_bwd", B, H, L);
  if (rc) return rc;
  if (gru_variant() == 1) rc = bwd_launch<8, 2>(dHs, gates, Hs, Whh, B, H, L, dgi, dgh, dgiT, dghT, bar, st);
  else rc = bwd_launch<16, 4>(dHs, gates, Hs, Whh, B, H, L, dgi, dgh, dgiT, dghT, bar, st);
  if (rc) return rc;
  EK_CHECK_LAUNCH();
  return EK_OK;
}
