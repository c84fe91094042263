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
constexpr int GF_NOGST = 512;      // fused epilogue runs but its global stores are skipped
constexpr int GF_NOSTAGE = 1024;   // no shared-memory staging round trip (stores garbage)
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
// explicit shared-space 16-byte accesses (the staging pointer is derived from an aligned-up generic pointer, for which
// the compiler would otherwise emit generic ST.E / LD.E -- measured 2x slower than the uncoalesced epilogue it replaced)
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

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
  static constexpr int EPI_BYTES = 8 * 32 * 32 * 4;                     // one swizzled 32x32 fp32 staging tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
};

// ---- the fused epilogue on one staged 32x32 chunk, in the transposed layout (this lane: rows i*4 + rg, cols n..n+3) ----
// The epilogue is ISSUE bound: a 128x256 tile is 8192 float4 row pieces, and every instruction spent per piece costs
// ~0.15 us per tile.  A loop that tests run-time feature flags and nullable pointers per piece compiled to ~90
// instructions (7 us per tile, as long as the K = 1024 main loop).  So the common combinations are compiled separately
// (OUT: bit 0 = fp32 C, bit 1 = Cb, bit 2 = Cb2; ADD = addend) with pointers resolved once per chunk and bumped
// unconditionally; everything else takes the general forms below.
struct EpiChunk {
  uint32_t stg;           // this warp's staging tile (shared-space address)
  int rg, cg;             // row group / 16-byte column group of this lane
  int nvalid;             // row groups of this lane inside the matrix
  long long row0;         // first row of this lane
  int n, nb;              // this lane's first column, the chunk's first column
  float* pc;              // fp32 destination of this lane's first element (C, or C2 for chunks at / beyond c_n1)
  long long sc;           // its stride per row group (4 rows)
};

__device__ __forceinline__ void epi_load8(const EpiChunk& k, float4* vt) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + k.rg;
    vt[i] = lds128(k.stg + (uint32_t)(rr * 128 + ((k.cg ^ (rr & 7)) << 4)));
  }
}

template <int OUT, bool ADD, bool DROP = false, bool ROWB = false>
__device__ __forceinline__ void epi_rows_fast(const EkEpilogue& ep, const EpiChunk& k, float4 bv, const float4* add4,
                                              unsigned long long dseed = 0ull, const float* rowb_lane = nullptr) {
  float4 vt[8];
  epi_load8(k, vt);
  float* pc = k.pc;
  bf16* pb = ep.Cb + k.row0 * ep.ldcb + k.n;
  bf16* pb2 = ep.Cb2 + k.row0 * ep.ldcb2 + (k.n - ep.cb2_n0);
  const long long sc = k.sc, sb = 4 * ep.ldcb, sb2 = 4 * ep.ldcb2;
  const int fb = ep.cb_fmt, fb2 = ep.cb2_fmt;
  // dropout counter of this lane's first element; one 64-bit draw covers the four columns (n is a multiple of 4)
  unsigned long long e0 = DROP ? (unsigned long long)k.row0 * ep.dropN + ep.dropOff + k.n : 0ull;
  const unsigned long long es = DROP ? 4ull * (unsigned long long)ep.dropN : 0ull;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long pr = 0ull;
    if (ROWB) pr = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)rowb_lane, i * 4 + k.rg);
    if (i < k.nvalid) {
      float4 v = vt[i];
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      if (DROP) {
        float mk[4];
        ek_drop_multv<4>(ep.drop, dseed, e0, mk);
        v.x *= mk[0]; v.y *= mk[1]; v.z *= mk[2]; v.w *= mk[3];
      }
      if (ADD) { v.x += add4[i].x; v.y += add4[i].y; v.z += add4[i].z; v.w += add4[i].w; }
      if (ROWB) {
        const float4 a4 = __ldg((const float4*)((const float*)(uintptr_t)pr + k.n));
        v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
      }
      if (OUT & 1) *(float4*)pc = v;
      if (OUT & 2) *(uint2*)pb = make_uint2(pack16x2(v.x, v.y, fb), pack16x2(v.z, v.w, fb));
      if (OUT & 4) *(uint2*)pb2 = make_uint2(pack16x2(v.x, v.y, fb2), pack16x2(v.z, v.w, fb2));
    }
    if (OUT & 1) pc += sc;
    if (OUT & 2) pb += sb;
    if (OUT & 4) pb2 += sb2;
    if (DROP) e0 += es;
  }
}

// split-K partial sums: fp32 reduction in L2, one 16-byte reduction per four columns
__device__ __forceinline__ void epi_rows_red(const EkEpilogue& ep, const EpiChunk& k) {
  float4 vt[8];
  epi_load8(k, vt);
  float* pc = ep.C + k.row0 * ep.ldc + k.n;
  const long long sc = 4 * ep.ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < k.nvalid)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pc), "f"(vt[i].x), "f"(vt[i].y), "f"(vt[i].z),
                   "f"(vt[i].w)
                   : "memory");
    pc += sc;
  }
}

// four consecutive columns of one row -> C (fp32) / Cb / Cb2 (16-bit, per-output format and column range)
__device__ __forceinline__ void epi_store4(const EkEpilogue& ep, long long mm, int n, int nb, float4 v) {
  if (ep.C2 && nb >= ep.c_n1) *(float4*)(ep.C2 + mm * ep.ldc2 + (n - ep.c_n1)) = v;
  else if (ep.C) *(float4*)(ep.C + mm * ep.ldc + n) = v;
  if (ep.Cb && (ep.cb_n1 == 0 || nb < ep.cb_n1)) {
    uint2 pk;
    pk.x = pack16x2(v.x, v.y, ep.cb_fmt);
    pk.y = pack16x2(v.z, v.w, ep.cb_fmt);
    *(uint2*)(ep.Cb + mm * ep.ldcb + n) = pk;
  }
  if (ep.Cb2 && nb >= ep.cb2_n0) {
    uint2 pk;
    pk.x = pack16x2(v.x, v.y, ep.cb2_fmt);
    pk.y = pack16x2(v.z, v.w, ep.cb2_fmt);
    *(uint2*)(ep.Cb2 + mm * ep.ldcb2 + (n - ep.cb2_n0)) = pk;
  }
}

// general form: every epilogue feature, one row group per iteration of a ROLLED loop (one copy of the dropout hash and
// of the transcendental activations in the instruction stream)
__device__ __forceinline__ void epi_rows_general(const EkEpilogue& ep, const EpiChunk& k, float4 bv, const float* rowb_lane,
                                              unsigned long long dseed) {
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + k.rg;
    const unsigned long long pr =
        ep.rowb ? __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)rowb_lane, rr) : 0ull;
    if (i < k.nvalid) {
      const long long mm = k.row0 + 4 * i;
      float4 v = lds128(k.stg + (uint32_t)(rr * 128 + ((k.cg ^ (rr & 7)) << 4)));
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      if (ep.drop.seed) {
        float mk[4];
        ek_drop_multv<4>(ep.drop, dseed, (unsigned long long)mm * ep.dropN + ep.dropOff + k.n, mk);
        v.x *= mk[0]; v.y *= mk[1]; v.z *= mk[2]; v.w *= mk[3];
      }
      if (ep.addend && (ep.add_n1 == 0 || k.nb < ep.add_n1)) {
        const float4 a4 = *(const float4*)(ep.addend + mm * ep.ldadd + k.n);
        v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
      }
      if (ep.rowb) {
        const float4 a4 = __ldg((const float4*)((const float*)(uintptr_t)pr + k.n));
        v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
      }
      if (ep.act != EK_ACT_NONE) {
        v.x = ek_act(v.x, ep.act); v.y = ek_act(v.y, ep.act); v.z = ek_act(v.z, ep.act); v.w = ek_act(v.w, ep.act);
      }
      epi_store4(ep, mm, k.n, k.nb, v);
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
    // TMEM -> registers (tcgen05.ld 32x32b: lane = accumulator row) -> per-warp 4 KB staging tile in shared memory
    // (16-byte chunks XOR-swizzled by row: conflict-free both ways) -> read back TRANSPOSED, so that one warp
    // instruction touches 4 rows x 128 contiguous bytes.  The fused epilogue runs in that layout: bias / addend /
    // row-broadcast loads and the C / Cb / Cb2 stores are all fully coalesced (a row-per-lane store of the same data
    // costs 32 cache-line transactions per instruction and made the epilogue -- 6 us per 128x256 bf16 tile, 13 us per
    // fp32 tile -- as long as the whole K = 1024 main loop: profiles/r02_gemm_probe.json).  Two warps share each TMEM
    // lane quarter and take alternate 32-column chunks; the next chunk's TMEM load is in flight while this one is stored.
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;            // which of the two warps of this quarter
    const uint32_t stg = smem_u32(smem + C::STAGES * C::STAGE_BYTES + 256) + (uint32_t)(warp - 2) * 4096u;
    const int rg = lane >> 3, cg = lane & 7;     // transposed layout: row group / 4-column group of this lane
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
      const long long mrow0 = (long long)m0 + q * 32;            // first accumulator row of this warp's quarter
      const long long mlane = mrow0 + lane;
      // row-broadcast operand: pointer of this lane's OWN row; the transposed loop fetches it by shuffle
      const float* rowb_lane = nullptr;
      if (ep.rowb && mlane < M) {
        if (ep.rowflag && ep.rowflag[mlane]) rowb_lane = ep.rowb_alt;
        else rowb_lane = ep.rowb + (long long)((mlane / ep.rowb_div) % ep.rowb_mod) * ep.ldrowb;
      }
      const unsigned long long dseed = ep.drop.seed ? ek_seed(ep.drop) : 0ull;
      // compiled-in fast forms cover: bias, addend, dropout, row broadcast, the common output sets; no activation
      const int fastmask = (ep.act == EK_ACT_NONE) ? 1 : 0;
      const bool lean = fastmask != 0;
      int rows_here = (int)((long long)M - mrow0);
      rows_here = rows_here < 0 ? 0 : (rows_here > 32 ? 32 : rows_here);
      int nch = (N - n0 + 31) / 32;
      nch = nch > bn_eff / 32 ? bn_eff / 32 : nch;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      auto release_acc = [&]() {                 // this warp has read everything it needs from the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));   // the leader issues the MMAs
          else mbar_arrive(&tempty_bar[buf]);
        }
      };
      uint32_t r[32];
      int c = half;
      if ((flags & GF_NOEPI) || c >= nch) {
        release_acc();
      } else {
        tc_ld32(tbase + c * 32, r);
#pragma unroll 1
        for (; c < nch; c += 2) {
          const int nb = n0 + c * 32;
          const bool fast = (vec_ok == 15) && nb + 32 <= N && !(flags & GF_NOSTORE);
          const int n = nb + cg * 4;
          // operand prefetch in the transposed layout, before waiting for the accumulator chunk
          float4 add4[8];
          const bool add_here = ep.addend && (ep.add_n1 == 0 || nb < ep.add_n1);     // (add_n1, c_n1: multiples of 32)
          if (fast && lean && add_here && splits == 1) {
            const float* pa = ep.addend + (mrow0 + rg) * ep.ldadd + n;
            const long long sa = 4 * ep.ldadd;
            const int nvalid = (rows_here - rg + 3) >> 2;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              add4[i] = (i < nvalid) ? *(const float4*)pa : make_float4(0.f, 0.f, 0.f, 0.f);
              pa += sa;
            }
          }
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (fast && ep.bias && splits == 1) bv = __ldg((const float4*)(ep.bias + n));
          const bool st_on = dbg && warp == 2 && lane == 0 && it == 0 && c < 4;
          const int sb = 22 + (c >> 1) * 4;           // slots 22..25 (chunk 0), 26..29 (chunk 2)
          if (st_on) dbg[(size_t)blockIdx.x * 32 + sb] = gtime_ns();
          tc_wait_ld();
          if (st_on) dbg[(size_t)blockIdx.x * 32 + sb + 1] = gtime_ns();
          if (fast) {
            if (!(flags & GF_NOSTAGE))
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(stg + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), __uint_as_float(r[4 * j]),
                     __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            __syncwarp();
            if (c + 2 < nch) tc_ld32(tbase + (c + 2) * 32, r);      // next chunk's TMEM load overlaps the stores below
            else release_acc();
            if (st_on) dbg[(size_t)blockIdx.x * 32 + sb + 2] = gtime_ns();
            {
              EpiChunk k;
              k.stg = stg; k.rg = rg; k.cg = cg; k.nvalid = (rows_here - rg + 3) >> 2; k.row0 = mrow0 + rg; k.n = n; k.nb = nb;
              if (ep.C2 && nb >= ep.c_n1) { k.pc = ep.C2 + k.row0 * ep.ldc2 + (n - ep.c_n1); k.sc = 4 * ep.ldc2; }
              else { k.pc = ep.C + k.row0 * ep.ldc + n; k.sc = 4 * ep.ldc; }
              if (flags & GF_NOGST) {
              } else if (splits > 1) {
                epi_rows_red(ep, k);
              } else {
                // outputs this chunk really writes (Cb / Cb2 may be limited to a column range)
                const int om = fastmask == 0 ? 0
                             : ((ep.C ? 1 : 0) | ((ep.Cb && (ep.cb_n1 == 0 || nb < ep.cb_n1)) ? 2 : 0) |
                                ((ep.Cb2 && nb >= ep.cb2_n0) ? 4 : 0));
                const int feat = (add_here ? 1 : 0) | (ep.drop.seed ? 2 : 0) | (ep.rowb ? 4 : 0);
                switch (om * 8 + feat) {
                  case 8: epi_rows_fast<1, false>(ep, k, bv, add4); break;
                  case 9: epi_rows_fast<1, true>(ep, k, bv, add4); break;
                  case 10: epi_rows_fast<1, false, true>(ep, k, bv, add4, dseed); break;
                  case 11: epi_rows_fast<1, true, true>(ep, k, bv, add4, dseed); break;
                  case 16: epi_rows_fast<2, false>(ep, k, bv, add4); break;
                  case 17: epi_rows_fast<2, true>(ep, k, bv, add4); break;
                  case 18: epi_rows_fast<2, false, true>(ep, k, bv, add4, dseed); break;
                  case 20: epi_rows_fast<2, false, false, true>(ep, k, bv, add4, 0ull, rowb_lane); break;
                  case 24: epi_rows_fast<3, false>(ep, k, bv, add4); break;
                  case 25: epi_rows_fast<3, true>(ep, k, bv, add4); break;
                  case 28: epi_rows_fast<3, false, false, true>(ep, k, bv, add4, 0ull, rowb_lane); break;
                  case 32: epi_rows_fast<4, false>(ep, k, bv, add4); break;
                  case 48: epi_rows_fast<6, false>(ep, k, bv, add4); break;
                  case 52: epi_rows_fast<6, false, false, true>(ep, k, bv, add4, 0ull, rowb_lane); break;
                  default: epi_rows_general(ep, k, bv, rowb_lane, dseed); break;
                }
              }
            }
            if (st_on) dbg[(size_t)blockIdx.x * 32 + sb + 3] = gtime_ns();
            __syncwarp();                          // staging tile is rewritten by the next chunk
          } else {
            // cold path (unaligned operands / partial last chunk): staged like the fast path so that the registers are
            // free again, then this lane walks its own row element by element in a ROLLED loop (one copy of the scalar
            // epilogue in the instruction stream instead of 32)
            if (!(flags & GF_NOSTORE)) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(stg + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), __uint_as_float(r[4 * j]),
                       __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
              __syncwarp();
              if (mlane < M) {
#pragma unroll 1
                for (int j = 0; j < 32; ++j) {
                  const int nn = nb + j;
                  if (nn >= N) break;
                  float a;
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a)
                               : "r"(stg + (uint32_t)(lane * 128 + (((j >> 2) ^ (lane & 7)) << 4) + ((j & 3) << 2)))
                               : "memory");
                  if (splits > 1) atomicAdd(ep.C + mlane * ep.ldc + nn, a);
                  else ek_epilogue_store(ep, mlane, nn, a);
                }
              }
            }
            __syncwarp();
            if (c + 2 < nch) tc_ld32(tbase + (c + 2) * 32, r);
            else release_acc();
          }
        }
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

// gemm_skinny.cu: the M <= 64 form (answer decoder's per-step products)
int ek_gemm_skinny_ok(int transA, int transB, int M, int N, int K, const void* A, long long lda, const void* B, long long ldb,
                      const EkEpilogue& ep, int fmt);
int ek_gemm_skinny_launch(int transB, int M, int N, int K, const bf16* A, long long lda, const bf16* B, long long ldb,
                          const EkEpilogue& ep, int fmt, int splits_req, cudaStream_t st);

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
  // at most 64 rows: half of a 128-row tcgen05 tile would be empty and its fixed costs dominate -> the lean warp-MMA form
  if (force_bn == 0 && g_dbg_flags == 0 && ek_gemm_skinny_ok(transA, transB, M, N, K, A, lda, B, ldb, ep, fmt))
    return ek_gemm_skinny_launch(transB, M, N, K, A, lda, B, ldb, ep, fmt, splits, stream);
  // force_bn = 1000 + width selects the CTA-pair (cluster of 2, multicast B) variant of that width.  It is NOT the
  // default: measured on B200 it is no faster than independent CTAs (L2 already merges the two CTAs' requests for the
  // same B tile), and the lock-step coupling costs ~10-40 % on some shapes (profiles/r01_notes.md).
  bool want_cluster = false, want_cg2 = false;
  if (force_bn > 2000) { want_cg2 = true; force_bn -= 2000; }          // 2000 + width: cta_group::2 pair
  else if (force_bn > 1000) { want_cluster = true; force_bn -= 1000; }
  // Split-K candidates: plain fp32 outputs, or "C += A B" (addend aliases C, nothing else in the epilogue) where the
  // partial sums are reduced straight onto the existing values.
  const bool acc_alias = ep.addend && ep.addend == ep.C && ep.ldadd == ep.ldc;
  const bool plain_out = ep.C && !ep.C2 && !ep.Cb && !ep.Cb2 && !ep.bias && (!ep.addend || acc_alias) && !ep.rowb &&
                         ep.act == EK_ACT_NONE && !ep.drop.seed && ep.add_n1 == 0;
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
  // CTA pairs (cta_group::2, 256 x 256 per pair, each CTA stages half of B) by default where they measured faster with
  // the coalesced epilogue (profiles/r02_gemm_probe.json): long K loops or very many tiles; an even number of M tiles
  // keeps every pair full.  EKAID_B200_CG2=0 turns the rule off.
  static int cg2_auto = -1;
  if (cg2_auto < 0) {
    const char* e = getenv("EKAID_B200_CG2");
    cg2_auto = (e && e[0] == '0') ? 0 : 1;
  }
  if (cg2_auto && force_bn == 0 && cl == 1 && bn == 256 && N >= 256 && (ek_div_up(M, BM) % 2) == 0) {
    const long long tiles = (long long)ek_div_up(M, BM) * ek_div_up(N, 256);
    if ((K >= 2048 && tiles >= 100) || tiles >= 12LL * num_sms()) cl = 3;
  }
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
  if (ep.C2 && (((uintptr_t)ep.C2 & 15) || (ep.ldc2 & 3))) vec_ok &= ~1;
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
