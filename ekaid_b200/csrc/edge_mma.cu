// bf16 path of the per-image edge kernels on tensor cores (warp-level mma.sync m16n8k16, fp32 accumulate).
// The per-image problems are tiny GEMMs (52 x 208 x 1024 and 52 x 52 x 256), far below a tcgen05 128-row tile,
// and the kernels are bound by the Z / X traffic, so the legacy warp MMA is the right tool here: operands are
// staged once in shared memory (cp.async for the 16-byte-aligned Z rows), fragments come from ldmatrix.
//
//   aggregate fwd : out[i, c]   = sum_{(h,j)} P[i,(h,j)] Z[(h,j), c] + b ;  Xout = Xin + relu(2 out)
//   aggregate bwd : dZ[(h,j),c] = sum_i P[i,(h,j)] dout[i,c] ;  dPpart[slice][i,(h,j)] = sum_{c in slice} dout[i,c] Z[(h,j),c]
//
// P is split into bf16 hi + lo parts in the forward so the attention weights keep 16 significant bits.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int NC = 128;          // output columns per CTA
constexpr int ZS = NC + 8;       // shared-memory row pitch (elements) of Z / dout tiles: 272 B = odd multiple of 16 B
constexpr int MAXCH = 256;       // (h,j) rows staged per chunk

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
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
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// stage rows k0..k0+kc of the (h,j)-indexed Z matrix, columns [c0, c0+NC), into Zs[kc_pad][ZS]
__device__ __forceinline__ void load_z_chunk(bf16* Zs, const bf16* __restrict__ QKZ, long long ld, int D, int g, int N,
                                             int Kn, int HK, int k0, int kc_pad, int c0, int zoff) {
  // thread -> (row kk, 16-byte piece ch); a thread's rows are kstep apart, so (h, j) = (k / Kn, k % Kn) is divided out once
  // and stepped from row to row (a division per copy was a third of the instructions the 16-warp aggregate backward issued)
  const int tid = threadIdx.x;
  const int ch = tid % (NC / 8);
  const int kstep = blockDim.x / (NC / 8);          // block sizes are multiples of NC / 8 = 16
  const bool cok = c0 + ch * 8 < D;
  int k = k0 + tid / (NC / 8);
  int h = k / Kn, j = k - h * Kn;
  for (int kk = tid / (NC / 8); kk < kc_pad; kk += kstep, k += kstep) {
    bf16* dst = Zs + kk * ZS + ch * 8;
    if (k < HK && cok)
      cp_async16(dst, QKZ + ((size_t)g * N + j) * ld + (size_t)zoff + (size_t)h * D + c0 + ch * 8);
    else
      *(uint4*)dst = make_uint4(0, 0, 0, 0);
    j += kstep;
    while (j >= Kn) { j -= Kn; ++h; }
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// F16: Z (at Zsrc, pitch ldz, head 0 at column zoff) and the attention weights (plane 2 of Phl) hold IEEE fp16 -- one
// MMA per product instead of the bf16 path's hi + lo pair, and 11 significant bits in Z; XoutT is then written as fp16 too.
template <int MR, bool F16>   // padded query rows per CTA pass: 64 or 128
__global__ void __launch_bounds__(256, (F16 && MR == 64) ? 3 : 1)
agg_fwd_mma_kernel(const float* __restrict__ P, const bf16* __restrict__ QKZ, long long ld, int D,
                   const float* __restrict__ b_out, const float* __restrict__ Xin, int N, int Kn, int H,
                   float* __restrict__ Xout, bf16* __restrict__ XoutT, long long ldt, uint8_t* __restrict__ mask,
                   int kchunk, EkDrop dr, const bf16* __restrict__ Phl, long long plane, int zoff) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  const unsigned long long sd = ek_seed(dr);
  const bool don = dr.seed != nullptr && dr.p > 0.f;          // dropout constants of the epilogue (common.cuh rule)
  const unsigned int dthr = (unsigned int)(dr.p * 65536.0f);
  const float dkeep = 1.f / (1.f - dr.p);
  const int PS = kchunk + 8;                        // P row pitch (elements)
  bf16* Zs = (bf16*)smraw;                          // [kchunk][ZS]
  bf16* Phi = Zs + (size_t)kchunk * ZS;             // [MR][PS]
  bf16* Plo = Phi + (size_t)MR * PS;                // [MR][PS]
  const int g = blockIdx.x, c0 = blockIdx.y * NC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HK = H * Kn;
  constexpr int ITEMS = (MR / 16) * 2;              // (16-row tile, 64-column half)
  constexpr int IPW = ITEMS / 8;                    // items per warp
  for (int r0 = 0; r0 < N; r0 += MR) {
    float acc[IPW][8][4];
#pragma unroll
    for (int a = 0; a < IPW; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    for (int k0 = 0; k0 < HK; k0 += kchunk) {
      const int kc = min(kchunk, HK - k0);
      const int kc_pad = (kc + 15) & ~15;
      __syncthreads();
      load_z_chunk(Zs, QKZ, ld, D, g, N, Kn, HK, k0, kc_pad, c0, zoff);
      if (F16) {
        // fp16 attention weights (plane 2 of the softmax kernel's 16-bit planes): one plane, pure async copies
        const int cpr = kc_pad / 8;
        for (int e = tid; e < MR * cpr; e += 256) {
          const int i = e / cpr, ch = e % cpr;
          bf16* dst = Phi + i * PS + ch * 8;
          if (r0 + i < N && ch * 8 < kc)
            cp_async16(dst, Phl + 2 * plane + ((size_t)g * N + r0 + i) * HK + k0 + ch * 8);
          else
            *(uint4*)dst = make_uint4(0, 0, 0, 0);
        }
      } else if (Phl) {
        // attention weights pre-split into bf16 hi/lo planes by the softmax kernel: pure async copies
        const int cpr = kc_pad / 8;                  // 16-byte chunks per row
        for (int e = tid; e < 2 * MR * cpr; e += 256) {
          const int pl = e / (MR * cpr), rem = e % (MR * cpr);
          const int i = rem / cpr, ch = rem % cpr;
          bf16* dst = (pl ? Plo : Phi) + i * PS + ch * 8;
          if (r0 + i < N && ch * 8 < kc)             // HK % 8 == 0 (checked by the launcher): chunks never straddle kc
            cp_async16(dst, Phl + pl * plane + ((size_t)g * N + r0 + i) * HK + k0 + ch * 8);
          else
            *(uint4*)dst = make_uint4(0, 0, 0, 0);
        }
      } else {
        for (int i = warp; i < MR; i += 8) {         // one row per warp pass: all loads of a row are independent
          const bool rok = r0 + i < N;
          const float* Pr = P + ((size_t)g * N + r0 + i) * HK + k0;
#pragma unroll
          for (int r = 0; r < MAXCH / 32; ++r) {
            const int kk = lane + 32 * r;
            if (kk < kc_pad) {
              const float p = (rok && kk < kc) ? Pr[kk] : 0.f;
              const bf16 hi = __float2bfloat16_rn(p);
              Phi[i * PS + kk] = hi;
              Plo[i * PS + kk] = __float2bfloat16_rn(p - __bfloat162float(hi));
            }
          }
        }
      }
      cp_async_wait_all();
      __syncthreads();
#pragma unroll
      for (int it = 0; it < IPW; ++it) {
        const int item = warp + it * 8;
        const int mt = item >> 1, nh = item & 1;
        for (int kt = 0; kt < kc_pad / 16; ++kt) {
          uint32_t ah[4], al[4];
          const int arow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int acol = kt * 16 + (lane >> 4) * 8;
          ldsm_x4(ah, Phi + arow * PS + acol);
          if (!F16) ldsm_x4(al, Plo + arow * PS + acol);
#pragma unroll
          for (int np = 0; np < 4; ++np) {           // pairs of 8-column tiles
            uint32_t bfr[4];
            const int brow = kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
            const int bcol = nh * 64 + np * 16 + (lane >> 4) * 8;
            ldsm_x4_t(bfr, Zs + brow * ZS + bcol);
            if (F16) {
              mma_f16_16816(acc[it][2 * np], ah, bfr[0], bfr[1]);
              mma_f16_16816(acc[it][2 * np + 1], ah, bfr[2], bfr[3]);
            } else {
              mma_bf16_16816(acc[it][2 * np], ah, bfr[0], bfr[1]);
              mma_bf16_16816(acc[it][2 * np], al, bfr[0], bfr[1]);
              mma_bf16_16816(acc[it][2 * np + 1], ah, bfr[2], bfr[3]);
              mma_bf16_16816(acc[it][2 * np + 1], al, bfr[2], bfr[3]);
            }
          }
        }
      }
    }
    // epilogue: + b_out, doubled ReLU, residual.  Row bases are computed once per thread row, and all Xin / bias loads
    // of an item are issued before the first use (they were the kernel's main stall).
#pragma unroll
    for (int it = 0; it < IPW; ++it) {
      const int item = warp + it * 8;
      const int mt = item >> 1, nh = item & 1;
      const int colb = c0 + nh * 64 + 2 * (lane & 3);          // column of nt = 0
      size_t rbase[2], tbase[2];                             // element offsets of this thread's two rows in Xout / XoutT
      bool rok[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int i = r0 + mt * 16 + (lane >> 2) + hh * 8;
        rok[hh] = i < N;
        rbase[hh] = ((size_t)g * N + (rok[hh] ? i : 0)) * D + colb;
        tbase[hh] = ((size_t)g * N + (rok[hh] ? i : 0)) * (size_t)ldt + colb;
      }
      // two halves of four column tiles: the loads of a half are all in flight before its first use, and the live
      // operand registers stay at 24 (three CTAs per SM need <= 85 registers per thread)
#pragma unroll
      for (int nh4 = 0; nh4 < 2; ++nh4) {
      float2 bo[8], xi[2][8];
#pragma unroll
      for (int nt = 4 * nh4; nt < 4 * nh4 + 4; ++nt) {
        const bool cok = colb + nt * 8 < D;
        bo[nt] = cok ? *(const float2*)(b_out + colb + nt * 8) : make_float2(0.f, 0.f);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          xi[hh][nt] = (cok && rok[hh] && Xin) ? *(const float2*)(Xin + rbase[hh] + nt * 8) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int nt = 4 * nh4; nt < 4 * nh4 + 4; ++nt) {
        if (colb + nt * 8 >= D) continue;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (!rok[hh]) continue;
          const size_t idx = rbase[hh] + nt * 8;
          const float o0 = acc[it][nt][2 * hh] + bo[nt].x, o1 = acc[it][nt][2 * hh + 1] + bo[nt].y;
          float x0, x1, p0 = 0.f, p1 = 0.f;
          if (mask == nullptr) {
            // plain attention output of GraphSelfAttentionLayer.forward: no doubling, dropout or ReLU
            x0 = xi[hh][nt].x + o0;
            x1 = xi[hh][nt].y + o1;
          } else {
            // train mode: Dropout(0.2) on the doubled output before the ReLU (graph_att.py:103-104)
            // (idx is even -- D % 8 == 0, even column -- so both elements come from one draw: ek_drop_multv<2>'s rule with
            // the threshold hoisted and without its odd-index fallback, which doubled the size of this epilogue)
            float dm0 = 1.f, dm1 = 1.f;
            if (don) {
              const unsigned long long z = ek_draw64(sd, dr.site, idx >> 2) >> (16 * (idx & 3));
              dm0 = ((unsigned int)z & 0xFFFFu) >= dthr ? dkeep : 0.f;
              dm1 = ((unsigned int)(z >> 16) & 0xFFFFu) >= dthr ? dkeep : 0.f;
            }
            p0 = (o0 + o0) * dm0;
            p1 = (o1 + o1) * dm1;
            x0 = xi[hh][nt].x + fmaxf(p0, 0.f);
            x1 = xi[hh][nt].y + fmaxf(p1, 0.f);
          }
          *(float2*)(Xout + idx) = make_float2(x0, x1);
          if (XoutT)          // (the operand copy may have its own pitch; no 64-bit `idx / D` per element to find its row)
            *(uint32_t*)(XoutT + tbase[hh] + nt * 8) = F16 ? pack_f16x2_sat(x0, x1) : pack_bf16x2(x0, x1);
          if (mask) {
            uchar2 mk;
            mk.x = p0 > 0.f ? 1 : 0;
            mk.y = p1 > 0.f ? 1 : 0;
            *(uchar2*)(mask + idx) = mk;
          }
        }
      }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int MR>   // padded query rows (all of them are staged): 64 or 128
__global__ void __launch_bounds__(256)
agg_bwd_mma_kernel(const float* __restrict__ dXout, const uint8_t* __restrict__ mask, const float* __restrict__ P,
                   const bf16* __restrict__ QKZ, long long ld, int D, int N, int Kn, int H, bf16* __restrict__ dQKZ,
                   float* __restrict__ dOut, float* __restrict__ dPpart, int kchunk, float gscale,
                   const bf16* __restrict__ Phl) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int PS = kchunk + 8;
  bf16* dOs = (bf16*)smraw;                         // [MR][ZS]   dout tile (i, c)
  bf16* Zs = dOs + (size_t)MR * ZS;                 // [kchunk][ZS]
  bf16* Ps = Zs + (size_t)kchunk * ZS;              // [MR][PS]   P (i, (h,j)) chunk
  const int g = blockIdx.x, slice = blockIdx.y, c0 = slice * NC;
  const int G = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HK = H * Kn;
  // dout = 2 * mask * dX  -> global (fp32, for the b_out column sum) and shared (bf16 operand)
#pragma unroll 4
  for (int e = tid; e < MR * (NC / 4); e += 256) {
    const int i = e / (NC / 4), c = (e % (NC / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < N && c0 + c < D) {
      const size_t idx = ((size_t)g * N + i) * D + c0 + c;
      const float4 d = *(const float4*)(dXout + idx);
      const uchar4 m = *(const uchar4*)(mask + idx);
      v = make_float4(m.x ? gscale * d.x : 0.f, m.y ? gscale * d.y : 0.f, m.z ? gscale * d.z : 0.f,
                      m.w ? gscale * d.w : 0.f);
      *(float4*)(dOut + idx) = v;
    }
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *(uint32_t*)&a;
    pk.y = *(uint32_t*)&b;
    *(uint2*)(dOs + i * ZS + c) = pk;
  }
  float* dPg = dPpart + ((size_t)slice * G + g) * N * HK;
  for (int k0 = 0; k0 < HK; k0 += kchunk) {
    const int kc = min(kchunk, HK - k0);
    const int kc_pad = (kc + 15) & ~15;
    __syncthreads();
    load_z_chunk(Zs, QKZ, ld, D, g, N, Kn, HK, k0, kc_pad, c0, 2 * D);
    if (Phl) {
      const int cpr = kc_pad / 8;
      for (int e = tid; e < MR * cpr; e += 256) {
        const int i = e / cpr, ch = e % cpr;
        bf16* dst = Ps + i * PS + ch * 8;
        if (i < N && ch * 8 < kc) cp_async16(dst, Phl + ((size_t)g * N + i) * HK + k0 + ch * 8);
        else *(uint4*)dst = make_uint4(0, 0, 0, 0);
      }
    } else {
      for (int i = warp; i < MR; i += 8) {
        const bool rok = i < N;
        const float* Pr = P + ((size_t)g * N + i) * HK + k0;
#pragma unroll
        for (int r = 0; r < MAXCH / 32; ++r) {
          const int kk = lane + 32 * r;
          if (kk < kc_pad) Ps[i * PS + kk] = __float2bfloat16_rn((rok && kk < kc) ? Pr[kk] : 0.f);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- dZ[(h,j), c] = sum_i P[i,(h,j)] dout[i,c] : rows = (h,j) tiles, round-robin over warps
    for (int mt = warp; mt < kc_pad / 16; mt += 8) {
      float acc[16][4];
#pragma unroll
      for (int a = 0; a < 16; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll
      for (int kt = 0; kt < MR / 16; ++kt) {
        uint32_t af[4];
        // A = P^T: stored [i (k)][(h,j) (m)] -> transposed ldmatrix
        const int arow = kt * 16 + (lane >> 4) * 8 + (lane & 7);
        const int acol = mt * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(af, Ps + arow * PS + acol);
#pragma unroll
        for (int np = 0; np < 8; ++np) {
          uint32_t bfr[4];
          const int brow = kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int bcol = np * 16 + (lane >> 4) * 8;
          ldsm_x4_t(bfr, dOs + brow * ZS + bcol);
          mma_bf16_16816(acc[2 * np], af, bfr[0], bfr[1]);
          mma_bf16_16816(acc[2 * np + 1], af, bfr[2], bfr[3]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int kk = mt * 16 + (lane >> 2) + hh * 8;
        if (kk >= kc) continue;
        const int k = k0 + kk, h = k / Kn, j = k - h * Kn;
        bf16* rowp = dQKZ + ((size_t)g * N + j) * ld + 2 * (size_t)D + (size_t)h * D + c0 + 2 * (lane & 3);
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
          if (c0 + nt * 8 + 2 * (lane & 3) < D)
            *(__nv_bfloat162*)(rowp + nt * 8) = __floats2bfloat162_rn(acc[nt][2 * hh], acc[nt][2 * hh + 1]);
      }
    }
    // ---- dPpart[i, (h,j)] = sum_c dout[i,c] Z[(h,j),c] : items = (16-row tile of i, half of the (h,j) tiles)
    const int ntiles = kc_pad / 8;
    const int nhalf = (ntiles + 1) / 2;              // <= 16
    for (int item = warp; item < (MR / 16) * 2; item += 8) {
      const int mt = item >> 1, nb = (item & 1) * nhalf;
      const int ncnt = min(nhalf, ntiles - nb);
      float acc[16][4];
#pragma unroll
      for (int a = 0; a < 16; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll
      for (int kt = 0; kt < NC / 16; ++kt) {
        uint32_t af[4];
        const int arow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int acol = kt * 16 + (lane >> 4) * 8;
        ldsm_x4(af, dOs + arow * ZS + acol);
#pragma unroll
        for (int np = 0; np < 8; ++np) {
          if (2 * np < ncnt) {
            // B(k = c, n = (h,j)) = Z[(h,j)][c]: stored [n][k] -> plain ldmatrix
            uint32_t bfr[4];
            int nrow = (nb + 2 * np) * 8 + (lane >> 4) * 8 + (lane & 7);
            if (nrow >= kc_pad) nrow = kc_pad - 1;   // odd tile count: second tile unused
            const int kcol = kt * 16 + ((lane >> 3) & 1) * 8;
            ldsm_x4(bfr, Zs + nrow * ZS + kcol);
            mma_bf16_16816(acc[2 * np], af, bfr[0], bfr[1]);
            if (2 * np + 1 < ncnt) mma_bf16_16816(acc[2 * np + 1], af, bfr[2], bfr[3]);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        if (nt >= ncnt) continue;
        const int kk = (nb + nt) * 8 + 2 * (lane & 3);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int i = mt * 16 + (lane >> 2) + hh * 8;
          if (i >= N) continue;
          float* dst = dPg + (size_t)i * HK + k0 + kk;
          if (kk < kc) dst[0] = acc[nt][2 * hh];
          if (kk + 1 < kc) dst[1] = acc[nt][2 * hh + 1];
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// aggregate backward, one CTA per image (N <= 64, all (h,j) rows in one chunk, D % 128 == 0, P available as bf16).
// The CTA walks the D/128 column slices itself: P is staged once, Z slices are double-buffered with cp.async (the next
// slice streams in while this one is multiplied), and dP = sum_c dout[i,c] Z[(h,j),c] stays in registers across the
// slices -- no [slices, G, N, H*K] partial tensor, no re-reading of P, and loads overlap the MMAs.
// NW = warps per CTA.  With 8 warps (two per scheduler, 238 registers) the kernel was bound by the latency of its own
// dependent instructions: 27 % issue slots busy, 7.4 cycles between two instructions of a warp, stall_wait /
// short_scoreboard on top (profiles/r02_ncu_full_aggbwdimg.csv).  16 warps split every warp's accumulators in two (dZ: 64
// instead of 128 columns per item, dP: a quarter instead of half of the (h,j) tiles), which fits 128 registers and gives
// every scheduler four warps to choose from; shared memory and global traffic are unchanged.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
agg_bwd_img_kernel(const float* __restrict__ dXout, const uint8_t* __restrict__ mask, const bf16* __restrict__ QKZ,
                   long long ld, int D, int N, int Kn, int H, bf16* __restrict__ dQKZ, float* __restrict__ dOut,
                   float* __restrict__ dP, float gscale, const bf16* __restrict__ Phl) {
  ek_pdl_prologue();
  constexpr int MR = 64;
  constexpr int NTHR = NW * 32;
  constexpr int NSPL = NW / 8;                      // a dZ item covers NC / NSPL columns of one 16-row (h,j) tile
  constexpr int ZC = NC / NSPL;
  constexpr int NPART = NW / 4;                     // warps that share one 16-row query tile of dP
  constexpr int ACCN = 32 / NPART;                  // (h,j) 8-column tiles of dP per warp (HKP < 256: at most 31 in all)
  extern __shared__ __align__(16) uint8_t smraw[];
  const int HK = H * Kn;
  const int HKP = (HK + 15) & ~15;
  const int PS = HKP + 8;
  bf16* Ps = (bf16*)smraw;                          // [MR][PS]      P (i, (h,j)), staged once
  bf16* Zs0 = Ps + (size_t)MR * PS;                 // 2 x [HKP][ZS] Z slice, double-buffered
  float* raw = (float*)(Zs0 + (size_t)2 * HKP * ZS);   // [MR][NC]   dX slice as it comes from memory
  uint8_t* mk = (uint8_t*)(raw + MR * NC);          // [MR][NC]      ReLU mask slice
  bf16* dOs = (bf16*)(mk + MR * NC);                // [MR][ZS]      dout slice (bf16 operand)
  const int g = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = D / NC;

  auto issue_slice = [&](int s) {
    const int c0 = s * NC;
    load_z_chunk(Zs0 + (size_t)(s & 1) * HKP * ZS, QKZ, ld, D, g, N, Kn, HK, 0, HKP, c0, 2 * D);
    for (int e = tid; e < MR * (NC / 4); e += NTHR) {          // 16-byte pieces of the fp32 rows
      const int i = e / (NC / 4), c = (e % (NC / 4)) * 4;
      float* dst = raw + i * NC + c;
      if (i < N) cp_async16(dst, dXout + ((size_t)g * N + i) * D + c0 + c);
      else *(float4*)dst = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int e = tid; e < MR * (NC / 16); e += NTHR) {
      const int i = e / (NC / 16), c = (e % (NC / 16)) * 16;
      uint8_t* dst = mk + i * NC + c;
      if (i < N) cp_async16(dst, mask + ((size_t)g * N + i) * D + c0 + c);
      else *(uint4*)dst = make_uint4(0, 0, 0, 0);
    }
  };

  {
    const int cpr = HKP / 8;
    for (int e = tid; e < MR * cpr; e += NTHR) {
      const int i = e / cpr, ch = e % cpr;
      bf16* dst = Ps + i * PS + ch * 8;
      if (i < N && ch * 8 < HK) cp_async16(dst, Phl + ((size_t)g * N + i) * HK + ch * 8);
      else *(uint4*)dst = make_uint4(0, 0, 0, 0);
    }
  }
  issue_slice(0);

  // this warp's share of dP: 16 query rows x 1/NPART of the (h,j) tiles, accumulated over all slices
  const int ntiles = HKP / 8;
  const int nper = (ntiles + NPART - 1) / NPART;    // <= ACCN
  const int pmt = warp / NPART, pnb = (warp % NPART) * nper;
  const int pcnt = max(0, min(nper, ntiles - pnb));
  float accp[ACCN][4];
#pragma unroll
  for (int a = 0; a < ACCN; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) accp[a][c] = 0.f;

  for (int s = 0; s < S; ++s) {
    const int c0 = s * NC;
    const bf16* Zs = Zs0 + (size_t)(s & 1) * HKP * ZS;
    cp_async_wait_all();
    __syncthreads();                                 // slice s has landed; everybody is done with slice s-1
    // dout = gscale * mask * dX  -> global (fp32, for the b_out column sum) and shared (bf16 operand)
    for (int e = tid; e < MR * (NC / 4); e += NTHR) {
      const int i = e / (NC / 4), c = (e % (NC / 4)) * 4;
      const float4 d = *(const float4*)(raw + i * NC + c);
      const uchar4 m = *(const uchar4*)(mk + i * NC + c);
      const float4 v = make_float4(m.x ? gscale * d.x : 0.f, m.y ? gscale * d.y : 0.f, m.z ? gscale * d.z : 0.f,
                                   m.w ? gscale * d.w : 0.f);
      if (i < N) *(float4*)(dOut + ((size_t)g * N + i) * D + c0 + c) = v;
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *(uint32_t*)&a;
      pk.y = *(uint32_t*)&b;
      *(uint2*)(dOs + i * ZS + c) = pk;
    }
    __syncthreads();
    if (s + 1 < S) issue_slice(s + 1);               // streams in while this slice is multiplied
    // ---- dZ[(h,j), c] = sum_i P[i,(h,j)] dout[i,c]
    // (CTA-uniform trip counts and no warp-dependent branch around the ldmatrix / mma instructions below: the compiler
    // cannot see that `warp` is uniform inside a warp and would fence every one of them with WARPSYNC / BSSY otherwise; a
    // warp without an item in the last round multiplies tile 0 again and stores nothing)
    const int nitems = (HKP / 16) * NSPL;
    for (int it0 = 0; it0 < nitems; it0 += NW) {
      const bool live = it0 + warp < nitems;
      const int item = live ? it0 + warp : 0;
      const int mt = item / NSPL, zc0 = (item % NSPL) * ZC;      // 16 (h,j) rows x columns [zc0, zc0 + ZC) of the slice
      float acc[2 * (ZC / 16)][4];
#pragma unroll
      for (int a = 0; a < 2 * (ZC / 16); ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll
      for (int kt = 0; kt < MR / 16; ++kt) {
        uint32_t af[4];
        const int arow = kt * 16 + (lane >> 4) * 8 + (lane & 7);
        const int acol = mt * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(af, Ps + arow * PS + acol);
#pragma unroll
        for (int np = 0; np < ZC / 16; ++np) {
          uint32_t bfr[4];
          const int brow = kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int bcol = zc0 + np * 16 + (lane >> 4) * 8;
          ldsm_x4_t(bfr, dOs + brow * ZS + bcol);
          mma_bf16_16816(acc[2 * np], af, bfr[0], bfr[1]);
          mma_bf16_16816(acc[2 * np + 1], af, bfr[2], bfr[3]);
        }
      }
      // the two (h,j) rows this thread owns in the tile: one division each, outside the column loop
      bf16* rowp[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = mt * 16 + (lane >> 2) + hh * 8;
        const int h = k / Kn, j = k - h * Kn;
        rowp[hh] = (live && k < HK) ? dQKZ + ((size_t)g * N + j) * ld + 2 * (size_t)D + (size_t)h * D + c0 + zc0 + 2 * (lane & 3)
                                    : nullptr;
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (rowp[hh] == nullptr) continue;
#pragma unroll
        for (int nt = 0; nt < 2 * (ZC / 16); ++nt)
          *(__nv_bfloat162*)(rowp[hh] + nt * 8) = __floats2bfloat162_rn(acc[nt][2 * hh], acc[nt][2 * hh + 1]);
      }
    }
    // ---- dP[i, (h,j)] += sum_{c in slice} dout[i,c] Z[(h,j),c]
#pragma unroll
    for (int kt = 0; kt < NC / 16; ++kt) {
      uint32_t af[4];
      const int arow = pmt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int acol = kt * 16 + (lane >> 4) * 8;
      ldsm_x4(af, dOs + arow * ZS + acol);
#pragma unroll
      for (int np = 0; np < ACCN / 2; ++np) {          // all tile pairs, unconditionally (see above): tiles >= pcnt
        uint32_t bfr[4];                                //  read a clamped row and are never stored
        int nrow = (pnb + 2 * np) * 8 + (lane >> 4) * 8 + (lane & 7);
        if (nrow >= HKP) nrow = HKP - 1;
        const int kcol = kt * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(bfr, Zs + nrow * ZS + kcol);
        mma_bf16_16816(accp[2 * np], af, bfr[0], bfr[1]);
        mma_bf16_16816(accp[2 * np + 1], af, bfr[2], bfr[3]);
      }
    }
  }
  float* dPg = dP + (size_t)g * N * HK;
#pragma unroll
  for (int nt = 0; nt < ACCN; ++nt) {
    if (nt >= pcnt) continue;
    const int kk = (pnb + nt) * 8 + 2 * (lane & 3);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int i = pmt * 16 + (lane >> 2) + hh * 8;
      if (i >= N) continue;
      float* dst = dPg + (size_t)i * HK + kk;
      if (kk < HK) dst[0] = accp[nt][2 * hh];
      if (kk + 1 < HK) dst[1] = accp[nt][2 * hh + 1];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// scores + mask/bias + softmax on tensor cores: one CTA per (image, head), one warp per 16 query rows
// ------------------------------------------------------------------------------------------------
constexpr float NEG_MASK_MMA = -9e15f;

template <int NT>   // 8-column key tiles held per warp: 8 (Kn <= 64) or 16 (Kn <= 128)
__global__ void __launch_bounds__(256)
softmax_fwd_mma_kernel(const bf16* __restrict__ QKZ, long long ld, int D, const float* __restrict__ cond,
                       const float* __restrict__ lbias, const float* __restrict__ gbias, int N, int Kn, int H,
                       float* __restrict__ P, int MR, bf16* __restrict__ Phl, long long plane, int p16) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int g = blockIdx.x, h = blockIdx.y;
  const int dh = D / H;
  const int DS = dh + 8;                         // row pitch (elements); dh % 16 == 0 -> conflict-free ldmatrix
  constexpr int KR = NT * 8;
  constexpr int TS = KR + 1;                     // pitch of the fp32 tile (odd: rows fall on different banks)
  bf16* Qs = (bf16*)smraw;                       // [MR][DS]
  bf16* Ks = Qs + (size_t)MR * DS;               // [KR][DS]
  float* Ts = (float*)(Ks + (size_t)KR * DS);    // [MR][TS]: additive score terms on the way in, probabilities on the way out
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = dh / 8;
#pragma unroll 1
  for (int e = tid; e < (MR + KR) * chunks; e += 256) {
    const int r = e / chunks, ch = e % chunks;
    const bool isq = r < MR;
    const int row = isq ? r : r - MR;
    bf16* dst = (isq ? Qs + row * DS : Ks + row * DS) + ch * 8;
    const bool ok = isq ? (row < N) : (row < Kn);
    if (ok) cp_async16(dst, QKZ + ((size_t)g * N + row) * ld + (isq ? 0 : D) + h * dh + ch * 8);
    else *(uint4*)dst = make_uint4(0, 0, 0, 0);
  }
  // Additive term of every score element, fetched while the Q / K tiles are still in flight, in ONE rolled loop with
  // coalesced reads:   score = masked ? -9e15 : scale * qk + (geometry bias + label bias);   masked (cond <= 0) = -inf here.
  // (The reference adds the label bias to the -9e15 of a masked edge as well; in fp32 that sum IS -9e15 for any bias
  // below 2.7e8.)  The previous form -- 96 guarded scalar loads per thread in fully unrolled loops -- made this kernel
  // 6.7 k instructions long and instruction-fetch bound (34 % of warp samples `stall_no_inst`, profiles/r02_notes.md).
  const size_t total = (size_t)N * Kn;
  // (a warp per row, lanes over the keys: no index division; all loads of a row are issued before the first use)
#pragma unroll 1
  for (int i = warp; i < N; i += 8) {
    const size_t gr = (size_t)g * total + (size_t)i * Kn;
    float a[KR / 32], cv[KR / 32];
#pragma unroll
    for (int r = 0; r < KR / 32; ++r) {
      const int j = lane + 32 * r;
      a[r] = 0.f;
      cv[r] = 1.f;
      if (j < Kn) {
        if (gbias) a[r] = gbias[(gr + j) * H + h];
        if (lbias) a[r] += lbias[gr + j];
        if (cond) cv[r] = cond[gr + j];
      }
    }
#pragma unroll
    for (int r = 0; r < KR / 32; ++r) {
      const int j = lane + 32 * r;
      if (j < Kn) Ts[i * TS + j] = (cv[r] > 0.f) ? a[r] : -INFINITY;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int mt = warp;
  if (mt * 16 < N) {
    float acc[NT][4];
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 2
    for (int kt = 0; kt < dh / 16; ++kt) {
      uint32_t af[4];
      ldsm_x4(af, Qs + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * DS + kt * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bfr[4];
        // B(k = d, n = j) = K[j][d]: stored [n][k] -> plain ldmatrix
        ldsm_x4(bfr, Ks + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * DS + kt * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16_16816(acc[2 * np], af, bfr[0], bfr[1]);
        mma_bf16_16816(acc[2 * np + 1], af, bfr[2], bfr[3]);
      }
    }
    const float scale = 1.0f / sqrtf((float)dh);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int i = mt * 16 + (lane >> 2) + hh * 8;
      const bool rok = i < N;
      float* Tr = Ts + i * TS + 2 * (lane & 3);
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = nt * 8 + 2 * (lane & 3) + c;
          float sv = -INFINITY;
          if (rok && j < Kn) {
            const float a = Tr[nt * 8 + c];
            sv = (a == -INFINITY) ? NEG_MASK_MMA : scale * acc[nt][2 * hh + c] + a;
          }
          acc[nt][2 * hh + c] = sv;
          mx = fmaxf(mx, sv);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = nt * 8 + 2 * (lane & 3) + c;
          const float ev = (rok && j < Kn) ? expf(acc[nt][2 * hh + c] - mx) : 0.f;
          acc[nt][2 * hh + c] = ev;
          sum += ev;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (rok) {
        const float inv = 1.f / sum;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (nt * 8 + 2 * (lane & 3) + c < Kn) Tr[nt * 8 + c] = acc[nt][2 * hh + c] * inv;
      }
    }
  }
  __syncthreads();
  // probabilities out: P fp32 + the 16-bit planes the aggregation kernels stage (bf16 hi + lo = 16 significant bits for the
  // backward, fp16 for the fp16 forward aggregation), one rolled loop, consecutive threads on consecutive keys of a row
#pragma unroll 1
  for (int i = warp; i < N; i += 8) {
    const size_t orow = (((size_t)g * N + i) * H + h) * Kn;
#pragma unroll
    for (int r = 0; r < KR / 32; ++r) {
      const int j = lane + 32 * r;
      if (j < Kn) {
        const float pv = Ts[i * TS + j];
        const size_t o = orow + j;
        P[o] = pv;
        if (Phl) {
          const bf16 hi = __float2bfloat16_rn(pv);
          Phl[o] = hi;
          Phl[plane + o] = __float2bfloat16_rn(pv - __bfloat162float(hi));
          if (p16) ((f16*)Phl)[2 * plane + o] = from_f32<f16>(pv);
        }
      }
    }
  }
}

// backward of scores/softmax on tensor cores: one CTA per (image, head).
//   ds = P * (dP - sum_j P dP), dP = sum_slices dPpart;  outputs dgbias / dlbias_part;  ds_m = cond > 0 ? ds : 0
//   dQ = scale * ds_m K,  dK = scale * ds_m^T Q      (ds_m kept as bf16 hi + lo)
template <int MR>   // padded rows for both queries and keys: 64 or 128
__global__ void __launch_bounds__(256, MR == 64 ? 2 : 1)      // MR = 64: 86 KB of shared memory, two CTAs per SM
softmax_bwd_mma_kernel(const float* __restrict__ P, const float* __restrict__ dPpart, int nslices,
                       const bf16* __restrict__ QKZ, long long ld, int D, const float* __restrict__ cond, int N, int Kn,
                       int H, bf16* __restrict__ dQKZ, float* __restrict__ dlbias_part, float* __restrict__ dgbias) {
  ek_pdl_prologue();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int g = blockIdx.x, h = blockIdx.y, G = gridDim.x;
  const int dh = D / H, HK = H * Kn;
  const int DS = dh + 8;
  constexpr int SS = MR + 8;
  bf16* Qs = (bf16*)smraw;                       // [MR][DS]
  bf16* Ks = Qs + (size_t)MR * DS;               // [MR][DS]
  bf16* Shi = Ks + (size_t)MR * DS;              // [MR][SS]  ds_m (i, j)
  bf16* Slo = Shi + (size_t)MR * SS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = dh / 8;
  for (int e = tid; e < 2 * MR * chunks; e += 256) {
    const int r = e / chunks, ch = e % chunks;
    const bool isq = r < MR;
    const int row = isq ? r : r - MR;
    bf16* dst = (isq ? Qs + row * DS : Ks + row * DS) + ch * 8;
    const bool ok = isq ? (row < N) : (row < Kn);
    if (ok) cp_async16(dst, QKZ + ((size_t)g * N + row) * ld + (isq ? 0 : D) + h * dh + ch * 8);
    else *(uint4*)dst = make_uint4(0, 0, 0, 0);
  }
  const size_t total = (size_t)N * Kn;
  {
    // ds = P * (dP - <P, dP>) per query row; a warp owns rows warp, warp+8, ...  A ROLLED loop over the rows (two in flight):
    // the fully unrolled form (all rows' guarded loads hoisted in front of the reductions, the slice loop unrolled inside)
    // compiled to 14 k instructions -- 576 LDG, 4 k IMAD of address arithmetic -- and the kernel was instruction-fetch
    // bound (38 % of warp samples `stall_no_inst`, profiles/r02_notes.md).
    constexpr int NR = (MR + 31) / 32;         // key columns per lane
    const size_t sstride = (size_t)G * N * HK; // one slice of dPpart
#pragma unroll 2
    for (int i0 = 0; i0 < MR; i0 += 8) {         // (uniform trip count: the warp reductions below need no reconvergence fence)
      const int i = i0 + warp;
      const bool iok = i < N;
      const size_t row = (size_t)g * N + (iok ? i : 0);
      const float* Pr = P + (row * H + h) * Kn;
      const float* dPr = dPpart + row * HK + (size_t)h * Kn;
      const float* cr = cond ? cond + (size_t)g * total + (size_t)(iok ? i : 0) * Kn : nullptr;
      float pv[NR], dpv[NR], cv[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int j = lane + 32 * r;
        float pj = 0.f, dp = 0.f, cj = 1.f;
        if (iok && j < Kn) {
          pj = Pr[j];
#pragma unroll 1
          for (int sidx = 0; sidx < nslices; ++sidx) dp += dPr[(size_t)sidx * sstride + j];
          if (cr) cj = cr[j];
        }
        pv[r] = pj; dpv[r] = dp; cv[r] = cj;
      }
      float dot = 0.f;
#pragma unroll
      for (int r = 0; r < NR; ++r) dot = fmaf(pv[r], dpv[r], dot);
      dot = warp_sum(dot);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int j = lane + 32 * r;
        if (j < MR) {
          float ds = 0.f;
          if (iok && j < Kn) {
            ds = pv[r] * (dpv[r] - dot);
            const size_t ge = (size_t)g * total + (size_t)i * Kn + j;
            if (dgbias) dgbias[ge * H + h] = ds;
            if (dlbias_part) dlbias_part[(size_t)h * G * total + ge] = ds;
            if (cond && !(cv[r] > 0.f)) ds = 0.f;
          }
          const bf16 hi = __float2bfloat16_rn(ds);
          Shi[i * SS + j] = hi;
          Slo[i * SS + j] = __float2bfloat16_rn(ds - __bfloat162float(hi));
        }
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  // items: (which: dQ/dK, 16-row tile, 128-column block of d); every item has MR/16 k-steps
  const int nblk = dh / 128 + ((dh % 128) ? 1 : 0);
  const int per = (MR / 16) * nblk;
  // which (dQ / dK) and the round are CTA-uniform loop variables, and a warp without an item in the last round recomputes
  // item 0 and stores nothing: no warp-dependent branch around the ldmatrix / mma instructions (the compiler cannot see
  // that `warp` is uniform inside a warp and fenced them with WARPSYNC / BSSY: 61 of them for 128 HMMA)
#pragma unroll 1
  for (int which = 0; which < 2; ++which)
#pragma unroll 1
  for (int it0 = 0; it0 < per; it0 += 8) {
    const bool live = it0 + warp < per;
    const int pitem = live ? it0 + warp : 0;
    const int mt = pitem / nblk, nb = pitem % nblk;
    float acc[16][4];
#pragma unroll
    for (int a = 0; a < 16; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    const bf16* Bs = which == 0 ? Ks : Qs;       // dQ = ds K ; dK = ds^T Q  -> B(k, n = d) stored [k][n]: trans
#pragma unroll
    for (int kt = 0; kt < MR / 16; ++kt) {
      uint32_t ah[4], al[4];
      if (which == 0) {
        const int off = (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SS + kt * 16 + (lane >> 4) * 8;
        ldsm_x4(ah, Shi + off);
        ldsm_x4(al, Slo + off);
      } else {
        const int off = (kt * 16 + (lane >> 4) * 8 + (lane & 7)) * SS + mt * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(ah, Shi + off);
        ldsm_x4_t(al, Slo + off);
      }
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        // (columns beyond dh -- only when dh is no multiple of 128 -- read column block 0 and are not stored)
        const int ncol = (nb * 128 + np * 16 < dh) ? nb * 128 + np * 16 : 0;
        uint32_t bfr[4];
        ldsm_x4_t(bfr, Bs + (kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * DS + ncol + (lane >> 4) * 8);
        mma_bf16_16816(acc[2 * np], ah, bfr[0], bfr[1]);
        mma_bf16_16816(acc[2 * np], al, bfr[0], bfr[1]);
        mma_bf16_16816(acc[2 * np + 1], ah, bfr[2], bfr[3]);
        mma_bf16_16816(acc[2 * np + 1], al, bfr[2], bfr[3]);
      }
    }
    // rows of dK beyond Kn are zero by construction (ds columns j >= Kn are zero), rows >= N are not written
    bf16* dst = dQKZ + (size_t)g * N * ld + (which == 0 ? 0 : D) + h * dh;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const int d = nb * 128 + nt * 8 + 2 * (lane & 3);
      if (d >= dh) continue;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = mt * 16 + (lane >> 2) + hh * 8;
        if (live && r < N)
          *(__nv_bfloat162*)(dst + (size_t)r * ld + d) =
              __floats2bfloat162_rn(scale * acc[nt][2 * hh], scale * acc[nt][2 * hh + 1]);
      }
    }
  }
}

template <typename K>
int set_smem(K kern, size_t smem, size_t& configured, const char* what) {
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ek_set_error("%s: smem %zu: %s", what, smem, cudaGetErrorString(e)); return EK_ERR_CUDA; }
    configured = smem;
  }
  return EK_OK;
}

}  // namespace

// returns EK_ERR_UNSUPPORTED when the shape does not fit (caller falls back to the SIMT template)
// Z16 (optional): the Z blocks as IEEE fp16, [G*N, H*D] with pitch ldz16; needs the fp16 plane of Phl (3 planes)
int ek_agg_fwd_mma_launch(const float* P, const bf16* QKZ, long long ld, int D, const float* b_out, const float* Xin,
                          int G, int N, int Kn, int H, float* Xout, bf16* XoutT, long long ldt, uint8_t* mask,
                          EkDrop dr, const bf16* Phl, const bf16* Z16, long long ldz16, cudaStream_t st) {
  if ((D % 8) || (ld % 8) || (ldt % 2) || ((uintptr_t)QKZ & 15)) return EK_ERR_UNSUPPORTED;
  if (Z16 && ((ldz16 % 8) || ((uintptr_t)Z16 & 15) || !Phl || (H * Kn) % 8 || ((uintptr_t)Phl & 15))) return EK_ERR_UNSUPPORTED;
  const int HK = H * Kn;
  const int HKp = (HK + 15) & ~15;
  const int kchunk = HKp < MAXCH ? HKp : MAXCH;
  const int MR = N <= 64 ? 64 : 128;
  if (Phl && ((HK % 8) || ((uintptr_t)Phl & 15))) Phl = nullptr;      // planes need 16-byte-aligned rows
  const long long plane = (long long)G * N * HK;
  // fp16 forward: one weight plane, and (h, j) chunks of about half the rows -- 46 KB instead of 112 KB per CTA, so three
  // CTAs share an SM (the kernel is latency bound at 2 CTAs = 16 warps: 33 % stall_wait, 12 % issuing, profiles/r02_notes.md)
  int kc = kchunk;
  if (Z16 && HKp > 128) kc = ((HKp / 2) + 15) & ~15;
  const size_t smem = Z16 ? ((size_t)kc * ZS + (size_t)MR * (kc + 8)) * sizeof(bf16)
                          : ((size_t)kchunk * ZS + 2 * (size_t)MR * (kchunk + 8)) * sizeof(bf16);
  dim3 grid(G, ek_div_up(D, NC));
  static size_t c64 = 0, c128 = 0;
  static size_t h64 = 0, h128 = 0;
  if (Z16 && MR == 64) {
    int rc = set_smem(agg_fwd_mma_kernel<64, true>, smem, h64, "agg_fwd_mma");
    if (rc) return rc;
    ek_launch(agg_fwd_mma_kernel<64, true>, grid, 256, smem, st, P, Z16, ldz16, D, b_out, Xin, N, Kn, H, Xout, XoutT, ldt, mask,
                                                          kc, dr, Phl, plane, 0);
  } else if (Z16) {
    int rc = set_smem(agg_fwd_mma_kernel<128, true>, smem, h128, "agg_fwd_mma");
    if (rc) return rc;
    ek_launch(agg_fwd_mma_kernel<128, true>, grid, 256, smem, st, P, Z16, ldz16, D, b_out, Xin, N, Kn, H, Xout, XoutT, ldt, mask,
                                                           kc, dr, Phl, plane, 0);
  } else if (MR == 64) {
    int rc = set_smem(agg_fwd_mma_kernel<64, false>, smem, c64, "agg_fwd_mma");
    if (rc) return rc;
    ek_launch(agg_fwd_mma_kernel<64, false>, grid, 256, smem, st, P, QKZ, ld, D, b_out, Xin, N, Kn, H, Xout, XoutT, ldt, mask,
                                                           kchunk, dr, Phl, plane, 2 * D);
  } else {
    int rc = set_smem(agg_fwd_mma_kernel<128, false>, smem, c128, "agg_fwd_mma");
    if (rc) return rc;
    ek_launch(agg_fwd_mma_kernel<128, false>, grid, 256, smem, st, P, QKZ, ld, D, b_out, Xin, N, Kn, H, Xout, XoutT, ldt, mask,
                                                            kchunk, dr, Phl, plane, 2 * D);
  }
  EK_CHECK_LAUNCH();
  return EK_OK;
}

// one CTA per image with dP kept in registers (a single dP slice) when the shape allows it
static size_t agg_bwd_img_smem(int N, int Kn, int H) {
  const int HKP = (H * Kn + 15) & ~15;
  return ((size_t)64 * (HKP + 8) + (size_t)2 * HKP * ZS + (size_t)64 * ZS) * sizeof(bf16) + (size_t)64 * NC * 5;
}
int ek_agg_bwd_img_ok(int D, int N, int Kn, int H, int have_phl) {
  return have_phl && N <= 64 && D % NC == 0 && D / NC >= 2 && (H * Kn) % 8 == 0 && (D % 16) == 0 &&
         agg_bwd_img_smem(N, Kn, H) <= 227 * 1024;
}

int ek_agg_bwd_mma_launch(const float* dXout, const uint8_t* mask, const float* P, const bf16* QKZ, long long ld, int D,
                          int G, int N, int Kn, int H, bf16* dQKZ, float* dOut, float* dPpart, float gscale,
                          const bf16* Phl, cudaStream_t st) {
  if ((D % 8) || (ld % 8) || ((uintptr_t)QKZ & 15) || ((uintptr_t)dQKZ & 3) || N > 128) return EK_ERR_UNSUPPORTED;
  if (Phl && (((H * Kn) % 8) || ((uintptr_t)Phl & 15))) Phl = nullptr;
  if (ek_agg_bwd_img_ok(D, N, Kn, H, Phl != nullptr) && !((uintptr_t)dXout & 15) && !((uintptr_t)mask & 15)) {
    const size_t smem = agg_bwd_img_smem(N, Kn, H);
    static size_t cimg8 = 0, cimg16 = 0;
    // 16 warps per CTA by default; EKAID_B200_AGG_BWD_WARPS=8 selects the 8-warp form (A/B measurements)
    static const int nwarps = [] { const char* e = getenv("EKAID_B200_AGG_BWD_WARPS"); return (e && atoi(e) == 8) ? 8 : 16; }();
    if (nwarps == 16) {
      int rc = set_smem(agg_bwd_img_kernel<16>, smem, cimg16, "agg_bwd_img");
      if (rc) return rc;
      ek_launch(agg_bwd_img_kernel<16>, G, 512, smem, st, dXout, mask, QKZ, ld, D, N, Kn, H, dQKZ, dOut, dPpart, gscale, Phl);
    } else {
      int rc = set_smem(agg_bwd_img_kernel<8>, smem, cimg8, "agg_bwd_img");
      if (rc) return rc;
      ek_launch(agg_bwd_img_kernel<8>, G, 256, smem, st, dXout, mask, QKZ, ld, D, N, Kn, H, dQKZ, dOut, dPpart, gscale, Phl);
    }
    EK_CHECK_LAUNCH();
    return EK_OK;
  }
  const int HK = H * Kn;
  const int HKp = (HK + 15) & ~15;
  const int kchunk = HKp < MAXCH ? HKp : MAXCH;
  const int MR = N <= 64 ? 64 : 128;
  const size_t smem = ((size_t)MR * ZS + (size_t)kchunk * ZS + (size_t)MR * (kchunk + 8)) * sizeof(bf16);
  dim3 grid(G, ek_div_up(D, NC));
  static size_t c64 = 0, c128 = 0;
  if (MR == 64) {
    int rc = set_smem(agg_bwd_mma_kernel<64>, smem, c64, "agg_bwd_mma");
    if (rc) return rc;
    ek_launch(agg_bwd_mma_kernel<64>, grid, 256, smem, st, dXout, mask, P, QKZ, ld, D, N, Kn, H, dQKZ, dOut, dPpart, kchunk,
                                                    gscale, Phl);
  } else {
    int rc = set_smem(agg_bwd_mma_kernel<128>, smem, c128, "agg_bwd_mma");
    if (rc) return rc;
    ek_launch(agg_bwd_mma_kernel<128>, grid, 256, smem, st, dXout, mask, P, QKZ, ld, D, N, Kn, H, dQKZ, dOut, dPpart, kchunk,
                                                     gscale, Phl);
  }
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_softmax_fwd_mma_launch(const bf16* QKZ, long long ld, int D, const float* cond, const float* lbias,
                              const float* gbias, int G, int N, int Kn, int H, float* P, bf16* Phl, int p16,
                              cudaStream_t st) {
  const long long plane = (long long)G * N * H * Kn;
  if ((D % H) || ((D / H) % 16) || (ld % 8) || ((uintptr_t)QKZ & 15) || N > 128 || Kn > 128) return EK_ERR_UNSUPPORTED;
  const int dh = D / H;
  const int MR = ((N + 15) / 16) * 16;
  const int NT = Kn <= 64 ? 8 : 16;
  const size_t smem = (size_t)(MR + NT * 8) * (dh + 8) * sizeof(bf16) + (size_t)MR * (NT * 8 + 1) * sizeof(float);
  static size_t c8 = 0, c16 = 0;
  dim3 grid(G, H);
  if (NT == 8) {
    int rc = set_smem(softmax_fwd_mma_kernel<8>, smem, c8, "softmax_fwd_mma");
    if (rc) return rc;
    ek_launch(softmax_fwd_mma_kernel<8>, grid, 256, smem, st, QKZ, ld, D, cond, lbias, gbias, N, Kn, H, P, MR, Phl, plane, p16);
  } else {
    int rc = set_smem(softmax_fwd_mma_kernel<16>, smem, c16, "softmax_fwd_mma");
    if (rc) return rc;
    ek_launch(softmax_fwd_mma_kernel<16>, grid, 256, smem, st, QKZ, ld, D, cond, lbias, gbias, N, Kn, H, P, MR, Phl, plane, p16);
  }
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_softmax_bwd_mma_launch(const float* P, const float* dPpart, int nslices, const bf16* QKZ, long long ld, int D,
                              const float* cond, int G, int N, int Kn, int H, bf16* dQKZ, float* dlbias_part,
                              float* dgbias, cudaStream_t st) {
  if ((D % H) || ((D / H) % 16) || (ld % 8) || ((uintptr_t)QKZ & 15) || ((uintptr_t)dQKZ & 3) || N > 128)
    return EK_ERR_UNSUPPORTED;
  const int dh = D / H;
  const int MR = N <= 64 ? 64 : 128;
  const size_t smem = ((size_t)2 * MR * (dh + 8) + (size_t)2 * MR * (MR + 8)) * sizeof(bf16);
  static size_t c64 = 0, c128 = 0;
  dim3 grid(G, H);
  if (MR == 64) {
    int rc = set_smem(softmax_bwd_mma_kernel<64>, smem, c64, "softmax_bwd_mma");
    if (rc) return rc;
    ek_launch(softmax_bwd_mma_kernel<64>, grid, 256, smem, st, P, dPpart, nslices, QKZ, ld, D, cond, N, Kn, H, dQKZ,
                                                         dlbias_part, dgbias);
  } else {
    int rc = set_smem(softmax_bwd_mma_kernel<128>, smem, c128, "softmax_bwd_mma");
    if (rc) return rc;
    ek_launch(softmax_bwd_mma_kernel<128>, grid, 256, smem, st, P, dPpart, nslices, QKZ, ld, D, cond, N, Kn, H, dQKZ,
                                                          dlbias_part, dgbias);
  }
  EK_CHECK_LAUNCH();
  return EK_OK;
}
