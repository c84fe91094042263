// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared-memory ring ->
// tcgen05.mma (single elected thread, fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue.
//
//   C[M,N] = op(A) * op(B),  A(m,k), B(k,n) bf16;  fp32 accumulate;  fused epilogue (epilogue.cuh)
//
// Operand layouts (both served straight from the row-major tensors the rest of the path keeps in HBM,
// no transposed copies):
//   A_MN = 0: A stored [M, K] (K contiguous)   -> K-major UMMA operand      (forward, dgrad)
//   A_MN = 1: A stored [K, M] (M contiguous)   -> MN-major UMMA operand     (wgrad: dY^T)
//   B_MN = 0: B stored [N, K] (K contiguous)   -> K-major                   (forward: weight [out,in])
//   B_MN = 1: B stored [K, N] (N contiguous)   -> MN-major                  (dgrad: weight, wgrad: X)
//
// Persistent: grid = min(#tiles, #SMs); warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..9 = epilogue (two per TMEM lane quarter, alternate 32-column chunks).  Tile 128 x BN x 64, BN in {64,128,256}.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "epilogue.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

// kernel `flags` argument
constexpr int GF_A_F16 = 1;        // A operand holds IEEE fp16 (default bf16); kind::f16 takes either format per operand
constexpr int GF_B_F16 = 2;        // B operand holds IEEE fp16
// measurement-only ablations (ekaid_gemm_debug): results are garbage, timings isolate one pipeline role
constexpr int GF_NOSTORE = 16;     // epilogue drains TMEM but skips the fused epilogue (no global loads / stores)
constexpr int GF_NOTMA = 32;       // producer arrives on the full barriers without loading
constexpr int GF_NOMMA = 64;       // issuer commits without issuing MMAs
constexpr int GF_NOEPI = 256;      // epilogue does not even read TMEM
constexpr int GF_EARLY = 128;      // griddepcontrol.launch_dependents once this CTA has set up (experiment)

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GSTAMP(slot) do { if (dbg) dbg[(size_t)blockIdx.x * 32 + (slot)] = gtime_ns(); } while (0)

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  int spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins == 1024) t0 = clock64();
    if (spins > 1024 && (clock64() - t0) > 4000000000LL) {
      printf("ekaid gemm_tc: mbarrier wait timeout (block %d thread %d parity %u)\n", blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tmap, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, and each destination CTA's
// mbarrier (same CTA-relative offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* tmap, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// ---- cta_group::2 (CTA pair): one MMA of M = 256 spans both CTAs; each CTA stages its 128 rows of A and HALF of B
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA load whose completion is signalled on `bar_cluster_addr`, an mbarrier of either CTA of the pair
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* tmap, uint32_t bar_cluster_addr, void* dst, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}
// One lane of a converged warp, chosen by the hardware.  Unlike `lane == 0`, ptxas knows that exactly one lane runs the
// guarded region, so the uniform-datapath instructions inside it (UTMALDG, UTCHMMA, UTCBAR) are emitted directly instead
// of inside an ELECT / R2UR.BROADCAST / BRA.U.ANY loop per instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

template <int BN, int A_MN, int B_MN, int CG = 1>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;      // cta_group::2: each CTA stages half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CG == 2) ? (BN == 256 ? 6 : 8) : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // two accumulator buffers
  static constexpr int EPI_BYTES = 4 * 32 * 36 * 4;                     // transposition buffers of the 4 epilogue warps
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
};

// Direct epilogue of one 32-column chunk: this lane owns one accumulator row and stores 32 consecutive columns.
__device__ __forceinline__ void epi_direct_chunk(const EkEpilogue& ep, const uint32_t* r, long long m, int nb, int N,
                                                 int vec_ok, const float* rowb_ptr) {
  if ((vec_ok & 1) && nb + 32 <= N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (ep.bias) {
              if (vec_ok & 2) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b4 = __ldg((const float4*)(ep.bias + nb + j));
                  v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __ldg(ep.bias + nb + j);
              }
            }
            if (ep.drop.seed) {
              float mk[32];
              ek_drop_multv<32>(ep.drop, ek_seed(ep.drop), (unsigned long long)m * ep.dropN + ep.dropOff + nb, mk);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= mk[j];
            }
            if (ep.addend) {
              const float* ap = ep.addend + m * ep.ldadd + nb;
              if (vec_ok & 4) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 a4 = *(const float4*)(ap + j);
                  v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += ap[j];
              }
            }
            if (rowb_ptr) {
              if (vec_ok & 8) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 a4 = __ldg((const float4*)(rowb_ptr + nb + j));
                  v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __ldg(rowb_ptr + nb + j);
              }
            }
            if (ep.act != EK_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = ek_act(v[j], ep.act);
            }
            if (ep.C) {
              float* cp = ep.C + m * ep.ldc + nb;
#pragma unroll
              for (int j = 0; j < 32; j += 4) *(float4*)(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            // 16-bit outputs: bf16 or saturating fp16 per output, each limited to its column range (warp-uniform tests:
            // cb_n1 and cb2_n0 are multiples of the 32-column chunk)
            if (ep.Cb && (ep.cb_n1 == 0 || nb < ep.cb_n1)) {
              bf16* cp = ep.Cb + m * ep.ldcb + nb;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                pk.x = pack16x2(v[j], v[j + 1], ep.cb_fmt); pk.y = pack16x2(v[j + 2], v[j + 3], ep.cb_fmt);
                pk.z = pack16x2(v[j + 4], v[j + 5], ep.cb_fmt); pk.w = pack16x2(v[j + 6], v[j + 7], ep.cb_fmt);
                *(uint4*)(cp + j) = pk;
              }
            }
            if (ep.Cb2 && nb >= ep.cb2_n0) {
              bf16* cp = ep.Cb2 + m * ep.ldcb2 + (nb - ep.cb2_n0);
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                pk.x = pack16x2(v[j], v[j + 1], ep.cb2_fmt); pk.y = pack16x2(v[j + 2], v[j + 3], ep.cb2_fmt);
                pk.z = pack16x2(v[j + 4], v[j + 5], ep.cb2_fmt); pk.w = pack16x2(v[j + 6], v[j + 7], ep.cb2_fmt);
                *(uint4*)(cp + j) = pk;
              }
            }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {         // cold path (static indexing keeps r[] in registers)
      const int n = nb + j;
      if (n < N) ek_epilogue_store(ep, m, n, __uint_as_float(r[j]));
    }
  }
}

// CL = 1: independent CTAs.  CL = 2: thread-block cluster of two CTAs on neighbouring M tiles of the SAME N tile; each
// loads its own A tile and HALF of the shared B tile, multicast into both CTAs' shared memory -> 1.5x fewer bytes pulled
// from L2 per flop (the single-CTA 128x256 tile is L2-bandwidth bound at ~85 flop/B).
template <int BN, int A_MN, int B_MN, int CL, int CG = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                    int K, EkEpilogue ep, int vec_ok, int splits, int tail_from, int flags, unsigned long long* dbg) {
  using C = Cfg<BN, A_MN, B_MN, CG>;
  static_assert(CL == 1 || BN >= 128, "the shared B tile must split into two TMA boxes");
  static_assert(CG == 1 || CL == 2, "cta_group::2 needs a cluster of two CTAs");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  // work unit = (M tile group of CL tiles, N tile, K split); every CTA of a cluster walks the same unit sequence
  const int num_mg = (num_m + CL - 1) / CL;
  const int num_tiles = num_mg * num_n;
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per = (num_kb + splits - 1) / splits;
  // Work units [0, tail_from) are whole tiles.  When the last wave of tiles would leave most SMs idle, the launcher sets
  // tail_from < num_tiles and the remaining tiles are issued as half-width units (BN/2 columns each): the last wave
  // then costs about half a tile time on twice as many SMs.  (splits == 1 and CL == 1 in that mode.)
  const int num_units = (tail_from < num_tiles) ? tail_from + 2 * (num_tiles - tail_from) : num_tiles * splits;
  auto decode = [&](int unit, int& tile, int& ncol0, int& bn_eff) {
    if (unit < tail_from || tail_from >= num_tiles) {
      tile = unit % num_tiles;
      ncol0 = (tile / num_mg) * BN;
      bn_eff = BN;
    } else {
      const int hidx = unit - tail_from;
      tile = tail_from + (hidx >> 1);
      ncol0 = (tile / num_mg) * BN + (hidx & 1) * (BN / 2);
      bn_eff = BN / 2;
    }
  };
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / CL;
  const int unit_stride = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CG == 2 ? 1 : CL);   // one commit-arrive per MMA issuer that reads this slot
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], CG == 2 ? 16 : 8);   // one arrive per epilogue warp (of both CTAs under cta_group::2)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // peer barriers initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) GSTAMP(0);
  ek_pdl_wait();     // barriers, TMEM and tensor maps are set up; from here on we touch the previous kernel's data
  if (threadIdx.x == 0) GSTAMP(1);
  if (flags & GF_EARLY) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < num_units; unit += unit_stride) {
        int tile, n0, bn_eff;
        decode(unit, tile, n0, bn_eff);
        const int kb0 = (tail_from < num_tiles) ? 0 : (unit / num_tiles) * kb_per;
        const int kb1 = min(kb0 + kb_per, num_kb);
        const int m0 = ((tile % num_mg) * CL + (int)crank) * BM;
        const bool half_w = bn_eff != BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const int k0 = kb * BK;
          if (flags & GF_NOTMA) {
            if (CG == 1 || crank == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (unit == unit0 && kb == kb0) GSTAMP(2);
          if constexpr (CG == 2) {
            // cta_group::2: both CTAs' boxes complete on the LEADER's barrier (it alone waits, its MMA spans both CTAs)
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            if (A_MN == 0) {
              tma_load_2d_cg2(&tmA, lbar, sa, k0, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d_cg2(&tmA, lbar, sa + i * (64 * BK * 2), m0 + 64 * i, k0);
            }
            const int nh0 = n0 + (int)crank * (BN / 2);          // this CTA's half of the B tile
            if (B_MN == 0) {
              tma_load_2d_cg2(&tmB, lbar, sb, k0, nh0);          // box {64 k, BN/2 n}
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i) tma_load_2d_cg2(&tmB, lbar, sb + i * (64 * BK * 2), nh0 + 64 * i, k0);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], half_w ? C::A_BYTES + C::B_BYTES / 2 : C::STAGE_BYTES);
          if (A_MN == 0) {
            tma_load_2d(&tmA, &full_bar[stage], sa, k0, m0);                 // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)                                   // box {64 m, 64 k}
              tma_load_2d(&tmA, &full_bar[stage], sa + i * (64 * BK * 2), m0 + 64 * i, k0);
          }
          if (CL == 1) {
            if (B_MN == 0) {
              if (BN == 256) {                                                  // two boxes {64 k, 128 n}
                tma_load_2d(&tmB, &full_bar[stage], sb, k0, n0);
                if (!half_w) tma_load_2d(&tmB, &full_bar[stage], sb + (BN / 2) * BK * 2, k0, n0 + BN / 2);
              } else {
                tma_load_2d(&tmB, &full_bar[stage], sb, k0, n0);             // box {64 k, BN n}
              }
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)                                 // box {64 n, 64 k}
                if (!half_w || i < BN / 128)
                  tma_load_2d(&tmB, &full_bar[stage], sb + i * (64 * BK * 2), n0 + 64 * i, k0);
            }
          } else {
            // this CTA fetches half of the shared B tile and multicasts it to both CTAs of the pair
            if (B_MN == 0) {                                                    // box {64 k, BN/2 n}
              tma_load_2d_mc(&tmB, &full_bar[stage], sb + crank * (BN / 2 * BK * 2), k0, n0 + (int)crank * (BN / 2), 3);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i) {                              // boxes {64 n, 64 k}
                const int bi = (int)crank * (BN / 128) + i;
                tma_load_2d_mc(&tmB, &full_bar[stage], sb + bi * (64 * BK * 2), n0 + 64 * bi, k0, 3);
              }
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if ((CG == 1 || crank == 0) && elect_one()) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): f32 accum, bf16 x bf16
      // a_format / b_format (bits 7-9 / 10-12): 0 = fp16, 1 = bf16 -- chosen per operand, so a bf16 gradient can meet an
      // fp16 activation or weight in one MMA
      const uint32_t idesc = (1u << 4) | ((flags & GF_A_F16) ? 0u : (1u << 7)) | ((flags & GF_B_F16) ? 0u : (1u << 10)) |
                             ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t idesc_half = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(BN >> 4) << 17);      // N = BN / 2
      for (int unit = unit0; unit < num_units; unit += unit_stride, ++it) {
        const int kb0 = (tail_from < num_tiles) ? 0 : (unit / num_tiles) * kb_per;
        const int kb1 = min(kb0 + kb_per, num_kb);
        const uint32_t idesc_u = (tail_from < num_tiles && unit >= tail_from) ? idesc_half : idesc;
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && kb == kb0) GSTAMP(3);
          if (flags & GF_NOMMA) {
            if constexpr (CG == 2) tc_commit_cg2(&empty_bar[stage], 3);
            else if (CL == 1) tc_commit(&empty_bar[stage]);
            else tc_commit_mc(&empty_bar[stage], 3);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 8-row groups 1024 B apart, k-step = 32 B inside the swizzle span.
            // MN-major: 64-element column chunks 8 KB apart (LBO), 8-k-row groups 1024 B apart (SBO),
            //           k-step = 16 rows * 128 B.
            const uint64_t adesc = A_MN ? make_smem_desc(sa + k * (UMMA_K * 128), 64 * BK * 2, 1024)
                                        : make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(sb + k * (UMMA_K * 128), 64 * BK * 2, 1024)
                                        : make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            if constexpr (CG == 2) tc_mma_bf16_cg2(tmem_d, adesc, bdesc, idesc_u, ((kb - kb0) | k) ? 1u : 0u);
            else tc_mma_bf16(tmem_d, adesc, bdesc, idesc_u, ((kb - kb0) | k) ? 1u : 0u);
          }
          // frees the smem slot when these MMAs retire (in both CTAs of a pair: the peer multicasts into it too)
          if constexpr (CG == 2) tc_commit_cg2(&empty_bar[stage], 3);
          else if (CL == 1) tc_commit(&empty_bar[stage]);
          else tc_commit_mc(&empty_bar[stage], 3);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) tc_commit_cg2(&tfull_bar[buf], 3);   // accumulator complete -> both CTAs' epilogues
        else tc_commit(&tfull_bar[buf]);         // accumulator complete -> epilogue
        if (it < 6) GSTAMP(4 + 2 * it);          // all MMAs of work unit `it` issued
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2..9)
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;            // which of the two warps of this quarter
    float* stg = (float*)(smem + C::STAGES * C::STAGE_BYTES + 256) + (warp & 3) * (32 * 36);   // split-K path, half 0 only
    int it = 0;
    for (int unit = unit0; unit < num_units; unit += unit_stride, ++it) {
      int tile, n0, bn_eff;
      decode(unit, tile, n0, bn_eff);
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = ((tile % num_mg) * CL + (int)crank) * BM;
      mbar_wait(&tfull_bar[buf], acc_phase);
      tc_fence_after();
      if (warp == 2 && lane == 0 && it < 6) GSTAMP(16 + it);     // accumulator of work unit `it` complete
      const long long mlane0 = (long long)m0 + q * 32;
      if (flags & GF_NOEPI) {
      } else if (splits > 1 && half == 1) {
        // split-K partial sums are reduced by the first warp of each quarter (one staging buffer per quarter)
      } else if (splits > 1) {
        // Each lane owns accumulator row (q*32 + lane) in TMEM.  A 32x32 chunk is transposed through shared memory (16-byte
        // accesses, pitch 36 floats: conflict-free both ways) so that one warp instruction stores 4 rows x 128 contiguous
        // bytes instead of 32 scattered 16-byte pieces.
        const long long mlane = (long long)m0 + q * 32 + lane;
        const float* rowb_lane = nullptr;
        if (ep.rowb && mlane < M) {
          if (ep.rowflag && ep.rowflag[mlane]) rowb_lane = ep.rowb_alt;
          else rowb_lane = ep.rowb + (long long)((mlane / ep.rowb_div) % ep.rowb_mod) * ep.ldrowb;
        }
        const unsigned long long dseed = ep.drop.seed ? ek_seed(ep.drop) : 0ull;
        int rows_here = (int)((long long)M - ((long long)m0 + q * 32));
        rows_here = rows_here < 0 ? 0 : (rows_here > 32 ? 32 : rows_here);
        const int rg = lane >> 3, c4 = (lane & 7) * 4;        // this lane's row group / 4-column group after transposition
  #pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int nb = n0 + c * 32;
          if (nb >= N) break;                      // warp-uniform
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), r);
          tc_wait_ld();
          if ((vec_ok == 15) && nb + 32 <= N) {
  #pragma unroll
            for (int j = 0; j < 32; j += 4)
              *(float4*)(stg + lane * 36 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            __syncwarp();
            const int n = nb + c4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias) bv = __ldg((const float4*)(ep.bias + n));
  #pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + rg;
              const unsigned long long pr = ep.rowb ? __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)rowb_lane, rr) : 0ull;
              if (rr < rows_here) {
                const long long mm = (long long)m0 + q * 32 + rr;
                float4 v = *(const float4*)(stg + rr * 36 + c4);
                if (splits > 1) {                  // split-K partial sums: fp32 reduction in L2
                  float* cp = ep.C + mm * ep.ldc + n;
                  // one 16-byte reduction instead of four 4-byte ones: the L2 atomic units are the bottleneck of split-K
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(v.x), "f"(v.y), "f"(v.z),
                               "f"(v.w)
                               : "memory");
                } else {
                  v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                  if (ep.drop.seed) {
                    const unsigned long long e0 = (unsigned long long)mm * ep.dropN + ep.dropOff + n;
                    float mk[4];
                    ek_drop_multv<4>(ep.drop, dseed, e0, mk);
                    v.x *= mk[0]; v.y *= mk[1]; v.z *= mk[2]; v.w *= mk[3];
                  }
                  if (ep.addend) {
                    const float4 a4 = *(const float4*)(ep.addend + mm * ep.ldadd + n);
                    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
                  }
                  if (ep.rowb) {
                    const float4 a4 = __ldg((const float4*)((const float*)(uintptr_t)pr + n));
                    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
                  }
                  if (ep.act != EK_ACT_NONE) {
                    v.x = ek_act(v.x, ep.act); v.y = ek_act(v.y, ep.act); v.z = ek_act(v.z, ep.act); v.w = ek_act(v.w, ep.act);
                  }
                  if (ep.C) *(float4*)(ep.C + mm * ep.ldc + n) = v;
                  if (ep.Cb && (ep.cb_n1 == 0 || n < ep.cb_n1)) {
                    uint2 pk;
                    pk.x = pack16x2(v.x, v.y, ep.cb_fmt);
                    pk.y = pack16x2(v.z, v.w, ep.cb_fmt);
                    *(uint2*)(ep.Cb + mm * ep.ldcb + n) = pk;
                  }
                  if (ep.Cb2 && n >= ep.cb2_n0) {
                    uint2 pk;
                    pk.x = pack16x2(v.x, v.y, ep.cb2_fmt);
                    pk.y = pack16x2(v.z, v.w, ep.cb2_fmt);
                    *(uint2*)(ep.Cb2 + mm * ep.ldcb2 + (n - ep.cb2_n0)) = pk;
                  }
                }
              }
            }
            __syncwarp();
          } else if (mlane < M) {
            // cold path (unaligned operands / partial last chunk): this lane's row, element by element
  #pragma unroll 1
            for (int j = 0; j < 32; ++j) {
              const int n = nb + j;
              if (n < N) {
                if (splits > 1) atomicAdd(ep.C + mlane * ep.ldc + n, __uint_as_float(r[j]));
                else ek_epilogue_store(ep, mlane, n, __uint_as_float(r[j]));
              }
            }
          }
        }
      } else {
        // direct path, software-pipelined: the TMEM load of the next chunk is in flight while this one is stored.
        // Two warps share each TMEM lane quarter and take alternate 32-column chunks.
        const long long m = mlane0 + lane;
        const bool row_ok = m < M;
        const float* rowb_ptr = nullptr;
        if (ep.rowb && row_ok) {
          if (ep.rowflag && ep.rowflag[m]) rowb_ptr = ep.rowb_alt;
          else rowb_ptr = ep.rowb + (long long)((m / ep.rowb_div) % ep.rowb_mod) * ep.ldrowb;
        }
        int nch = (N - n0 + 31) / 32;
        nch = nch > bn_eff / 32 ? bn_eff / 32 : nch;
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
        uint32_t ra[32], rb[32];
        int c = half;
        if (c < nch) tc_ld32(tbase + c * 32, ra);
#pragma unroll 1
        for (; c < nch; c += 4) {
          tc_wait_ld();
          if (c + 2 < nch) tc_ld32(tbase + (c + 2) * 32, rb);
          if (row_ok && !(flags & GF_NOSTORE)) epi_direct_chunk(ep, ra, m, n0 + c * 32, N, vec_ok, rowb_ptr);
          if (c + 2 < nch) {
            tc_wait_ld();
            if (c + 4 < nch) tc_ld32(tbase + (c + 4) * 32, ra);
            if (row_ok && !(flags & GF_NOSTORE)) epi_direct_chunk(ep, rb, m, n0 + (c + 2) * 32, N, vec_ok, rowb_ptr);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));   // the leader issues the MMAs
        else mbar_arrive(&tempty_bar[buf]);
      }
      if (warp == 2 && lane == 0 && it < 6) GSTAMP(5 + 2 * it);  // epilogue of work unit `it` done (this warp)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) GSTAMP(31);
  if (CL > 1) cluster_sync_all();        // no CTA leaves while its peer may still multicast into it / arrive on it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
  }
}

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = (size_t)k.ptr;
    h ^= (size_t)k.rows * 1000003u + (size_t)k.cols * 10007u + (size_t)k.ld * 131u + (size_t)k.box_rows;
    return h;
  }
};

// 2-D bf16 row-major [rows, cols] with pitch ld; box = {64 cols, box_rows}, SWIZZLE_128B, zero OOB fill.
int make_tmap(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return EK_OK; }
  }
  PFN_encodeTiled enc = get_encode();
  EK_REQUIRE(enc != nullptr, EK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  EK_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld % 8) == 0, EK_ERR_ALIGN,
             "gemm_tc: operand must be 16-byte aligned with pitch %% 8 == 0 (ptr=%p ld=%lld)", ptr, ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EK_REQUIRE(r == CUDA_SUCCESS, EK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
             rows, cols, ld);
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return EK_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int g_dbg_flags = 0;                       // ekaid_gemm_debug: measurement-only ablation flags (GF_NO*) / experiments
unsigned long long* g_dbg_ts = nullptr;    // device buffer [grid][32] of globaltimer stamps, or null

template <int BN, int A_MN, int B_MN, int CL, int CG = 1>
int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EkEpilogue& ep, int vec_ok,
               int splits, int tail_halving, int flags, cudaStream_t stream) {
  flags |= g_dbg_flags;
  unsigned long long* dbg = g_dbg_ts;
  using C = Cfg<BN, A_MN, B_MN, CG>;
  static bool attr_set = false;
  auto kern = gemm_bf16_tc_kernel<BN, A_MN, B_MN, CL, CG>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gemm_tc: cannot set smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ek_div_up(ek_div_up(M, BM), CL) * ek_div_up(N, BN);
  const int units = tiles * splits;
  const int slots = num_sms() / CL;
  const int grid = CL * (units < slots ? units : slots);
  // Tail halving: a last wave of R <= slots/2 whole tiles would keep most SMs idle for a full tile time; issue those
  // tiles as 2R half-width units instead (kernel: decode()).  Only for the 256-wide single-CTA configuration.
  int tail_from = tiles;
  if (tail_halving && CL == 1 && BN == 256 && splits == 1 && tiles > slots) {
    const int rem = tiles % slots;
    if (rem > 0 && 2 * rem <= slots) tail_from = tiles - rem;
  }
  if (CL == 1) {
    ek_launch(kern, grid, NUM_THREADS, C::SMEM_BYTES, stream, ta, tb, M, N, K, ep, vec_ok, splits, tail_from, flags, dbg);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, ep, vec_ok, splits, tail_from, flags, dbg);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gemm_tc: cluster launch failed: %s", cudaGetErrorString(e));
  }
  EK_CHECK_LAUNCH();
  return EK_OK;
}

template <int A_MN, int B_MN>
int launch_bn(int bn, int cl, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EkEpilogue& ep,
              int vec_ok, int splits, int th, int flags, cudaStream_t stream) {
  if (cl == 3) {     // cta_group::2 pair
    if (bn == 128) return launch_cfg<128, A_MN, B_MN, 2, 2>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
    return launch_cfg<256, A_MN, B_MN, 2, 2>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
  }
  if (cl == 2) {
    if (bn == 128) return launch_cfg<128, A_MN, B_MN, 2>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
    return launch_cfg<256, A_MN, B_MN, 2>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
  }
  switch (bn) {
    case 64: return launch_cfg<64, A_MN, B_MN, 1>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
    case 128: return launch_cfg<128, A_MN, B_MN, 1>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
    default: return launch_cfg<256, A_MN, B_MN, 1>(ta, tb, M, N, K, ep, vec_ok, splits, th, flags, stream);
  }
}

}  // namespace

// transA: A stored [K, M];  transB: B stored [K, N]  (see file header)
void ek_gemm_debug(int flags, unsigned long long* ts) {
  g_dbg_flags = flags;
  g_dbg_ts = ts;
}

// fmt: bit 0 = A holds fp16 (else bf16), bit 1 = B holds fp16
int ek_gemm_bf16_tc_launch(int transA, int transB, int M, int N, int K, const bf16* A, long long lda, const bf16* B,
                           long long ldb, const EkEpilogue& ep, int force_bn, int splits, int fmt, cudaStream_t stream) {
  const int flags = fmt & (GF_A_F16 | GF_B_F16);
  EK_REQUIRE(M > 0 && N > 0 && K > 0, EK_ERR_SHAPE, "gemm_tc: bad shape M=%d N=%d K=%d", M, N, K);
  // force_bn = 1000 + width selects the CTA-pair (cluster of 2, multicast B) variant of that width.  It is NOT the
  // default: measured on B200 it is no faster than independent CTAs (L2 already merges the two CTAs' requests for the
  // same B tile), and the lock-step coupling costs ~10-40 % on some shapes (profiles/r01_notes.md).
  bool want_cluster = false, want_cg2 = false;
  if (force_bn > 2000) { want_cg2 = true; force_bn -= 2000; }          // 2000 + width: cta_group::2 pair
  else if (force_bn > 1000) { want_cluster = true; force_bn -= 1000; }
  // Split-K candidates: plain fp32 outputs, or "C += A B" (addend aliases C, nothing else in the epilogue) where the
  // partial sums are reduced straight onto the existing values.
  const bool acc_alias = ep.addend && ep.addend == ep.C && ep.ldadd == ep.ldc;
  const bool plain_out = ep.C && !ep.Cb && !ep.Cb2 && !ep.bias && (!ep.addend || acc_alias) && !ep.rowb &&
                         ep.act == EK_ACT_NONE && !ep.drop.seed;
  const int nkb = ek_div_up(K, BK);
  int bn = 256;
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) {
    bn = force_bn;
  } else if (plain_out && splits != 1 && nkb >= 16 && N >= 64) {
    // widest tile (best flop/byte from L2) whose tiles x possible splits still fill the machine
    const int cand[3] = {256, 128, 64};
    bn = 64;
    for (int i = 0; i < 3; ++i) {
      const int c = cand[i];
      if (c > 64 && N < c) continue;
      const long long tiles = (long long)ek_div_up(M, BM) * ek_div_up(N, c);
      if (tiles * (nkb / 4) >= (long long)(0.85 * num_sms()) || tiles >= num_sms()) { bn = c; break; }
    }
  } else {
    double best = -1;
    const int cand[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      const int c = cand[i];
      const long long tiles = (long long)ek_div_up(M, BM) * ek_div_up(N, c);
      double waves = (double)((tiles + num_sms() - 1) / num_sms());
      const long long rem = tiles % num_sms();
      if (c == 256 && tiles > num_sms() && rem > 0 && 2 * rem <= num_sms()) waves -= 0.44;   // tail halving (launch_cfg)
      // useful fraction of issued MMA work: (real N columns / padded) * (tiles / slots); wide tiles are
      // ~10% more efficient per flop (fewer A re-reads, shorter epilogue share)
      const double useful = (double)N / ((double)ek_div_up(N, c) * c) * (double)tiles / (waves * num_sms());
      const double score = useful * (c == 256 ? 1.0 : (c == 128 ? 0.72 : 0.5));   // measured: scripts/gemm_bench.py
      if (score > best) { best = score; bn = c; }
    }
  }
  // CTA pairs (cluster of 2 along M sharing the B tile): opt-in, needs two M tiles and a >= 128 wide tile
  int cl = (want_cluster && bn >= 128 && ek_div_up(M, BM) >= 2) ? 2 : 1;
  if (want_cg2 && bn >= 128 && ek_div_up(M, BM) >= 2) cl = 3;
  CUtensorMap ta, tb;
  int rc;
  if (!transA) rc = make_tmap(&ta, A, M, K, lda, BM);                 // [M rows, K cols], box {64, 128}
  else rc = make_tmap(&ta, A, K, M, lda, BK);                         // [K rows, M cols], box {64, 64}
  if (rc) return rc;
  if (!transB) rc = make_tmap(&tb, B, N, K, ldb, (cl >= 2 || bn == 256) ? bn / 2 : bn);   // [N rows, K cols], box {64, bn or bn/2}
  else rc = make_tmap(&tb, B, K, N, ldb, BK);                         // [K rows, N cols], box {64, 64}
  if (rc) return rc;
  // bit 0: 16-byte vector stores possible; bits 1..3: vector loads of bias / addend / row-broadcast operands
  int vec_ok = 15;
  if (ep.C && (((uintptr_t)ep.C & 15) || (ep.ldc & 3))) vec_ok &= ~1;
  if (ep.Cb && (((uintptr_t)ep.Cb & 15) || (ep.ldcb & 7))) vec_ok &= ~1;
  if (ep.Cb2 && (((uintptr_t)ep.Cb2 & 15) || (ep.ldcb2 & 7))) vec_ok &= ~1;
  if (ep.bias && ((uintptr_t)ep.bias & 15)) vec_ok &= ~2;
  if (ep.addend && (((uintptr_t)ep.addend & 15) || (ep.ldadd & 3))) vec_ok &= ~4;
  if (ep.rowb && (((uintptr_t)ep.rowb & 15) || (ep.ldrowb & 3) || ((uintptr_t)ep.rowb_alt & 15))) vec_ok &= ~8;
  // split-K (auto when splits == 0): too few tiles to fill the chip and a splittable epilogue
  const int num_kb = nkb;
  if (splits <= 0) {
    splits = 1;
    const long long tiles = (long long)ek_div_up(M, BM) * ek_div_up(N, bn);   // (pairs of tiles fill two SMs)
    if (plain_out && tiles * 2 <= num_sms() && num_kb >= 16) {
      splits = (int)(num_sms() / tiles);
      if (splits > num_kb / 4) splits = num_kb / 4;
      if (splits < 1) splits = 1;
    }
  }
  static int th = -1;
  if (th < 0) {
    const char* e = getenv("EKAID_B200_TAIL_HALVING");
    th = (e && e[0] == '0') ? 0 : 1;
  }
  EkEpilogue epk = ep;
  if (splits > 1) {
    EK_REQUIRE(plain_out, EK_ERR_UNSUPPORTED, "gemm_tc: split-K needs a plain fp32 (or C += AB) epilogue");
    const int kb_per = ek_div_up(num_kb, splits);
    splits = ek_div_up(num_kb, kb_per);        // no empty splits
    if (splits > 1) {
      if (!acc_alias) {
        cudaError_t e = cudaMemset2DAsync(ep.C, (size_t)ep.ldc * 4, 0, (size_t)N * 4, (size_t)M, stream);
        EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "gemm_tc: memset failed: %s", cudaGetErrorString(e));
      }
      epk.addend = nullptr;                    // partial sums are atomically added onto C (zeroed or pre-existing)
    }
  }
  if (!transA && !transB) return launch_bn<0, 0>(bn, cl, ta, tb, M, N, K, epk, vec_ok, splits, th, flags, stream);
  if (!transA && transB) return launch_bn<0, 1>(bn, cl, ta, tb, M, N, K, epk, vec_ok, splits, th, flags, stream);
  if (transA && !transB) return launch_bn<1, 0>(bn, cl, ta, tb, M, N, K, epk, vec_ok, splits, th, flags, stream);
  return launch_bn<1, 1>(bn, cl, ta, tb, M, N, K, epk, vec_ok, splits, th, flags, stream);
}
