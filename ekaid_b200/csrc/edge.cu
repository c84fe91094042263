// Fused per-image edge kernels of the relation-aware graph attention
// (reference models/graph_att_layer.py:60-178, models/graph_att.py:76-104, utils/mimic_utils.py:152-208).
//
// Inputs come from ONE projection GEMM per relation: QKZ[row, 0:D] = query, [D:2D] = key,
// [2D + h*D : 2D + (h+1)*D] = Z_h = self_feat * W_out2[:, h*D:(h+1)*D]^T   (re-association of quirk Q3:
// sum_h (P_h V) W_h^T == sum_h P_h (V W_h^T)).
//
//   adj_prep      : one-hot / float adjacency [G,N,N,L] -> mask source cond[g,i,j] = sum_c adj[g,j,i,c] and
//                   label bias lbias[g,i,j] = sum_c adj[g,j,i,c] * w[c]   (direction 1 = transposed, quirk Q2/Q5)
//   geom_bias     : boxes -> log(max(relu(W_p . posemb(i,j) + b_p), 1e-6)) per head (quirk Q7, Q13), the
//                   [B,N,K,64] fp64 position embedding is never materialised
//   edge_softmax  : scores = Q_h K_h^T / sqrt(dh) (+gbias) ; where(cond>0, s, -9e15) + lbias ; softmax_j (Q6)
//   edge_aggregate: out = sum_h P_h Z_h + b_out ; X_out = X_in + relu(2 out)   (Q2 doubling, Q1 residual)
// and their backward counterparts.
#include "common.cuh"

namespace {

constexpr float NEG_MASK = -9e15f;

// ------------------------------------------------------------------------------------------------
// adjacency -> (cond, lbias)
// ------------------------------------------------------------------------------------------------
__global__ void adj_prep_fwd_kernel(const float* __restrict__ adj0, const float* __restrict__ adj1, int g_split,
                                    const float* __restrict__ w, int N, int Kn, int L, float* __restrict__ cond,
                                    float* __restrict__ lbias) {
  ek_pdl_prologue();
  const int g = blockIdx.x;
  const float* adj = (g < g_split) ? adj0 + (size_t)g * N * N * L : adj1 + (size_t)(g - g_split) * N * N * L;
  for (int e = threadIdx.x; e < N * Kn; e += blockDim.x) {
    const int i = e / Kn, j = e % Kn;
    const float* a = adj + ((size_t)j * N + i) * L;     // transposed: adj[g, j, i, :]
    float s = 0.f, b = 0.f;
    for (int c = 0; c < L; ++c) {
      const float v = __ldg(a + c);
      s += v;
      b = fmaf(v, __ldg(w + c), b);
    }
    cond[(size_t)g * N * Kn + e] = s;
    lbias[(size_t)g * N * Kn + e] = b;
  }
}

// dw_part[g, c] = sum_{i,j} adj[g,j,i,c] * dlbias[g,i,j]      (one pass: all labels of an edge are consecutive in memory)
__global__ void adj_prep_bwd_kernel(const float* __restrict__ adj0, const float* __restrict__ adj1, int g_split,
                                    const float* __restrict__ dlbias_part, int nparts, int N, int Kn, int L,
                                    float* __restrict__ dw_part) {
  ek_pdl_prologue();
  constexpr int LC = 16;                   // labels per sweep (the reference has 11 spatial / 3 semantic labels)
  __shared__ float red[8][LC];
  const int g = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* adj = (g < g_split) ? adj0 + (size_t)g * N * N * L : adj1 + (size_t)(g - g_split) * N * N * L;
  for (int c0 = 0; c0 < L; c0 += LC) {
    const int lc = min(LC, L - c0);
    float acc[LC];
#pragma unroll
    for (int c = 0; c < LC; ++c) acc[c] = 0.f;
    for (int e = threadIdx.x; e < N * Kn; e += blockDim.x) {
      const int i = e / Kn, j = e % Kn;
      float dl = 0.f;
      for (int p = 0; p < nparts; ++p) dl += dlbias_part[((size_t)p * gridDim.x + g) * N * Kn + e];
      const float* ap = adj + ((size_t)j * N + i) * L + c0;
#pragma unroll
      for (int c = 0; c < LC; ++c)
        if (c < lc) acc[c] = fmaf(__ldg(ap + c), dl, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < LC; ++c) {
      const float v = warp_sum(acc[c]);
      if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < lc) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
      dw_part[(size_t)g * L + c0 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

// The same two kernels on the loader's INTEGER label matrices (int8 [*, S, S], label 0 = no edge, c+1 = plane c of
// process_matrix, utils/mimic_utils.py:119-149): a one-hot row sums to (1 <= label <= L) and its dot product with the
// bias table is w[label - 1], so the fp32 one-hot planes (4 * L bytes per edge) never have to exist.
__global__ void adj_labels_fwd_kernel(const int8_t* __restrict__ lab0, const int8_t* __restrict__ lab1, int g_split, int S,
                                      const float* __restrict__ w, int N, int Kn, int L, float* __restrict__ cond,
                                      float* __restrict__ lbias) {
  ek_pdl_prologue();
  const int g = blockIdx.x;
  const int8_t* lab = (g < g_split) ? lab0 + (size_t)g * S * S : lab1 + (size_t)(g - g_split) * S * S;
  for (int e = threadIdx.x; e < N * Kn; e += blockDim.x) {
    const int i = e / Kn, j = e % Kn;
    const int l = lab[(size_t)j * S + i];               // transposed: adjacency entry (j, i)   (graph_att.py:76, Q2)
    const bool on = l >= 1 && l <= L;
    cond[(size_t)g * N * Kn + e] = on ? 1.f : 0.f;
    lbias[(size_t)g * N * Kn + e] = on ? __ldg(w + l - 1) : 0.f;
  }
}
// dw_part[g, c] = sum over the edges of image g with label c+1 of sum_p dlbias_part[p, g, i, j]
__global__ void adj_labels_bwd_kernel(const int8_t* __restrict__ lab0, const int8_t* __restrict__ lab1, int g_split, int S,
                                      const float* __restrict__ dlbias_part, int nparts, int N, int Kn, int L,
                                      float* __restrict__ dw_part) {
  ek_pdl_prologue();
  constexpr int LC = 16;
  __shared__ float red[8][LC];
  const int g = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int8_t* lab = (g < g_split) ? lab0 + (size_t)g * S * S : lab1 + (size_t)(g - g_split) * S * S;
  for (int c0 = 0; c0 < L; c0 += LC) {
    const int lc = min(LC, L - c0);
    float acc[LC];
#pragma unroll
    for (int c = 0; c < LC; ++c) acc[c] = 0.f;
    for (int e = threadIdx.x; e < N * Kn; e += blockDim.x) {
      const int i = e / Kn, j = e % Kn;
      float dl = 0.f;
      for (int p = 0; p < nparts; ++p) dl += dlbias_part[((size_t)p * gridDim.x + g) * N * Kn + e];
      const int l = (int)lab[(size_t)j * S + i] - 1 - c0;
#pragma unroll
      for (int c = 0; c < LC; ++c) acc[c] += (c == l) ? dl : 0.f;
    }
#pragma unroll
    for (int c = 0; c < LC; ++c) {
      const float v = warp_sum(acc[c]);
      if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < lc) {
      float t = 0.f;
      for (int w_ = 0; w_ < (int)(blockDim.x >> 5); ++w_) t += red[w_][threadIdx.x];
      dw_part[(size_t)g * L + c0 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// geometry bias (implicit relation)
// ------------------------------------------------------------------------------------------------
// position features of (row r, col c) as the reference computes them in fp64:
//   (log max(|cx_r - cx_c| / w_r, 1e-3), log max(|cy_r - cy_c| / h_r, 1e-3), log(w_r / w_c), log(h_r / h_c))
__device__ __forceinline__ void pos_feats(const double* __restrict__ bb, int r, int c, double g[4]) {
  const double x0r = bb[r * 4 + 0], y0r = bb[r * 4 + 1], x1r = bb[r * 4 + 2], y1r = bb[r * 4 + 3];
  const double x0c = bb[c * 4 + 0], y0c = bb[c * 4 + 1], x1c = bb[c * 4 + 2], y1c = bb[c * 4 + 3];
  const double wr = x1r - x0r + 1.0, hr = y1r - y0r + 1.0;
  const double wc = x1c - x0c + 1.0, hc = y1c - y0c + 1.0;
  const double cxr = 0.5 * (x0r + x1r), cyr = 0.5 * (y0r + y1r);
  const double cxc = 0.5 * (x0c + x1c), cyc = 0.5 * (y0c + y1c);
  double dx = fabs((cxr - cxc) / wr);
  double dy = fabs((cyr - cyc) / hr);
  dx = dx < 1e-3 ? 1e-3 : dx;
  dy = dy < 1e-3 ? 1e-3 : dy;
  g[0] = log(dx);
  g[1] = log(dy);
  g[2] = log(wr / wc);
  g[3] = log(hr / hc);
}

// 64-d embedding of one pair: for component q (0..3), t (0..7): emb[q*16 + t] = sin(100 g_q / dim_t),
// emb[q*16 + 8 + t] = cos(same); dim_t = 1000^(t/8) rounded to fp32 (utils/mimic_utils.py:195-197).
// Computes f[h] = W[h,:] . emb + b[h] for h < H (H <= 8), optionally returns emb.
template <int MAXH>
__device__ __forceinline__ void pair_pos_fc(const double g[4], const float* __restrict__ Wp, const float* __restrict__ bp,
                                            int H, const float* __restrict__ dim_t, float f[MAXH], float* emb_out,
                                            const EkDrop& dr, unsigned long long seedv, unsigned long long pair_idx,
                                            bool fast_trig) {
#pragma unroll
  for (int h = 0; h < MAXH; ++h) f[h] = (h < H) ? bp[h] : 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    // train mode: Dropout(0.2) on the 64-d embedding before pair_pos_fc1 (fc.py:25-32); one 64-bit draw per 4 elements
    float ms[8], mc[8];
    ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 0, ms);
    ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 1, ms + 4);
    ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 2, mc);
    ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 3, mc + 4);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const double a = (100.0 * g[q]) / (double)dim_t[t];
      double sv, cv;
      if (fast_trig) {      // bf16 path: fp32 sincosf (full-range reduction, ~2 ulp) on the fp64-computed argument
        float sf, cf;
        sincosf((float)a, &sf, &cf);
        sv = sf;
        cv = cf;
      } else {
        sincos(a, &sv, &cv);
      }
      // reference casts the fp64 embedding to fp32 (graph_att_layer.py:115)
      const float s = (float)sv * ms[t];
      const float c = (float)cv * mc[t];
      if (emb_out) { emb_out[q * 16 + t] = s; emb_out[q * 16 + 8 + t] = c; }
#pragma unroll
      for (int h = 0; h < MAXH; ++h)
        if (h < H) f[h] = fmaf(Wp[h * 64 + q * 16 + t], s, fmaf(Wp[h * 64 + q * 16 + 8 + t], c, f[h]));
    }
  }
}

// gbias[g, i, j, h] for query i, key j.  Q13: pair index scrambling when N != Kn.
__global__ void geom_bias_fwd_kernel(const double* __restrict__ bb0, const double* __restrict__ bb1, int g_split,
                                     const float* __restrict__ Wp, const float* __restrict__ bp,
                                     const float* __restrict__ dim_t, int N, int Kn, int H,
                                     float* __restrict__ gbias, EkDrop dr, float* __restrict__ emb_cache,
                                     int fast_trig) {
  ek_pdl_prologue();
  extern __shared__ float sW[];      // H*64 + H, then 8 wave lengths
  float* sDim = sW + H * 65;
  for (int e = threadIdx.x; e < H * 64; e += blockDim.x) sW[e] = Wp[e];
  for (int e = threadIdx.x; e < H; e += blockDim.x) sW[H * 64 + e] = bp[e];
  if (threadIdx.x < 8) sDim[threadIdx.x] = dim_t[threadIdx.x];
  __syncthreads();
  const int g = blockIdx.x;
  const double* bb = (g < g_split) ? bb0 + (size_t)g * N * 4 : bb1 + (size_t)(g - g_split) * N * 4;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < N * Kn; e += gridDim.y * blockDim.x) {
    const int r = e / N, c = e % N;        // flat index i*Kn + j reinterpreted over [Kn, N]
    double gq[4];
    pos_feats(bb, r, c, gq);
    float f[8];
    if (emb_cache) {      // keep the (dropped) embedding for the backward pass: 256 B per pair instead of 32 sincos
      float emb[64];
      pair_pos_fc<8>(gq, sW, sW + H * 64, H, sDim, f, emb, dr, ek_seed(dr), (unsigned long long)g * N * Kn + e,
                     fast_trig != 0);
      float4* dst = (float4*)(emb_cache + ((size_t)g * N * Kn + e) * 64);
#pragma unroll
      for (int k = 0; k < 16; ++k) dst[k] = make_float4(emb[4 * k], emb[4 * k + 1], emb[4 * k + 2], emb[4 * k + 3]);
    } else {
      pair_pos_fc<8>(gq, sW, sW + H * 64, H, sDim, f, nullptr, dr, ek_seed(dr), (unsigned long long)g * N * Kn + e,
                     fast_trig != 0);
    }
    for (int h = 0; h < H; ++h) {
      const float v = fmaxf(fmaxf(f[h], 0.f), 1e-6f);
      gbias[((size_t)g * N * Kn + e) * H + h] = logf(v);
    }
  }
}

// bf16-path forward: box differences and ratios in fp64 (they cancel), everything after that in fp32 -- logf, the
// wave-length division as a multiplication by 100/dim_t, sin/cos as a two-term Cody-Waite reduction to [-pi, pi] + the
// SFU intrinsic (abs error 4e-7 there).  ~3x fewer instructions than the fp64 formulation the fp32 parity path keeps.
__device__ __forceinline__ void fast_sincos(float x, float* s, float* c) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831855f, x);
  r = fmaf(-k, -1.7484555e-7f, r);
  __sincosf(r, s, c);
}
__global__ void __launch_bounds__(128)
geom_bias_fwd_fast_kernel(const double* __restrict__ bb0, const double* __restrict__ bb1, int g_split,
                          const float* __restrict__ Wp, const float* __restrict__ bp, const float* __restrict__ dim_t,
                          int N, int Kn, int H, float* __restrict__ gbias, EkDrop dr, float* __restrict__ emb_cache) {
  ek_pdl_prologue();
  extern __shared__ float sW[];      // H*64 weights, H biases, then 8 x (100 / wave length)
  float* sK = sW + H * 65;
  for (int e = threadIdx.x; e < H * 64; e += blockDim.x) sW[e] = Wp[e];
  for (int e = threadIdx.x; e < H; e += blockDim.x) sW[H * 64 + e] = bp[e];
  if (threadIdx.x < 8) sK[threadIdx.x] = (float)(100.0 / (double)dim_t[threadIdx.x]);
  __syncthreads();
  const int g = blockIdx.x;
  const double* bb = (g < g_split) ? bb0 + (size_t)g * N * 4 : bb1 + (size_t)(g - g_split) * N * 4;
  const unsigned long long seedv = ek_seed(dr);
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < N * Kn; e += gridDim.y * blockDim.x) {
    const int r = e / N, c = e % N;        // flat index i*Kn + j reinterpreted over [Kn, N] (Q13)
    float gq[4];
    {
      const double x0r = bb[r * 4 + 0], y0r = bb[r * 4 + 1], x1r = bb[r * 4 + 2], y1r = bb[r * 4 + 3];
      const double x0c = bb[c * 4 + 0], y0c = bb[c * 4 + 1], x1c = bb[c * 4 + 2], y1c = bb[c * 4 + 3];
      const double wr = x1r - x0r + 1.0, hr = y1r - y0r + 1.0;
      const double wc = x1c - x0c + 1.0, hc = y1c - y0c + 1.0;
      const float iwr = 1.f / (float)wr, ihr = 1.f / (float)hr;
      const float dx = fmaxf(fabsf((float)(0.5 * (x0r + x1r) - 0.5 * (x0c + x1c)) * iwr), 1e-3f);
      const float dy = fmaxf(fabsf((float)(0.5 * (y0r + y1r) - 0.5 * (y0c + y1c)) * ihr), 1e-3f);
      gq[0] = logf(dx);
      gq[1] = logf(dy);
      gq[2] = logf((float)wr / (float)wc);
      gq[3] = logf((float)hr / (float)hc);
    }
    const unsigned long long pair_idx = (unsigned long long)g * N * Kn + e;
    float f[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) f[h] = (h < H) ? sW[H * 64 + h] : 0.f;
    float4* dst = emb_cache ? (float4*)(emb_cache + pair_idx * 64) : nullptr;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float ms[8], mc[8];
      ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 0, ms);
      ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 1, ms + 4);
      ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 2, mc);
      ek_drop_mult4(dr, seedv, pair_idx * 16 + q * 4 + 3, mc + 4);
      float es[8], ec[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float sv, cv;
        fast_sincos(gq[q] * sK[t], &sv, &cv);
        es[t] = sv * ms[t];
        ec[t] = cv * mc[t];
#pragma unroll
        for (int h = 0; h < 8; ++h)
          if (h < H) f[h] = fmaf(sW[h * 64 + q * 16 + t], es[t], fmaf(sW[h * 64 + q * 16 + 8 + t], ec[t], f[h]));
      }
      if (dst) {
        dst[q * 4 + 0] = make_float4(es[0], es[1], es[2], es[3]);
        dst[q * 4 + 1] = make_float4(es[4], es[5], es[6], es[7]);
        dst[q * 4 + 2] = make_float4(ec[0], ec[1], ec[2], ec[3]);
        dst[q * 4 + 3] = make_float4(ec[4], ec[5], ec[6], ec[7]);
      }
    }
    for (int h = 0; h < H; ++h) {
      const float v = fmaxf(fmaxf(f[h], 0.f), 1e-6f);
      gbias[pair_idx * H + h] = logf(v);
    }
  }
}

// part[g * GB_SPLIT + y, h*65 + k]: k < 64 -> dWp[h,k], k = 64 -> dbp[h]   (blockIdx.y = y takes every GB_SPLIT-th tile).
// Two phases per tile of 128 pairs: (1) one thread per pair recomputes the embedding and df = dgbias / f (f > 1e-6),
// parks both in shared memory; (2) one thread per (h,k) output accumulates over the tile.  No shuffles, no atomics.
constexpr int GB_TILE = 128;
constexpr int GB_SPLIT = 4;
__global__ void __launch_bounds__(GB_TILE)
geom_bias_bwd_kernel(const double* __restrict__ bb0, const double* __restrict__ bb1, int g_split,
                     const float* __restrict__ Wp, const float* __restrict__ bp, const float* __restrict__ dim_t, int N,
                     int Kn, int H, const float* __restrict__ dgbias, float* __restrict__ part, EkDrop dr,
                     const float* __restrict__ emb_cache, int fast_trig) {
  ek_pdl_prologue();
  extern __shared__ float sm[];      // weights H*65 | 8 wave lengths | emb [GB_TILE][65] | df [GB_TILE][8]
  float* sW = sm;
  float* sDim = sm + H * 65;
  float* sE = sDim + 8;
  float* sDf = (float*)(((uintptr_t)(sE + GB_TILE * 65) + 15) & ~(uintptr_t)15);      // 16-byte aligned rows of 8
  const int tid = threadIdx.x;
  if (tid < 8) sDim[tid] = dim_t[tid];
  for (int e = tid; e < H * 64; e += GB_TILE) sW[e] = Wp[e];
  for (int e = tid; e < H; e += GB_TILE) sW[H * 64 + e] = bp[e];
  const int g = blockIdx.x;
  const double* bb = (g < g_split) ? bb0 + (size_t)g * N * 4 : bb1 + (size_t)(g - g_split) * N * 4;
  const int total = N * Kn;
  const int nout = H * 65;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};         // outputs tid, tid+128, ... (H <= 8 -> <= 520 outputs)
  float kacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // cached path: column tid % 64 of every head
  float bacc = 0.f;                                  // cached path: bias gradient of head tid (threads 0..7)
  __syncthreads();
  for (int t0 = blockIdx.y * GB_TILE; t0 < total; t0 += gridDim.y * GB_TILE) {
    const int e = t0 + tid;
    float df[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) df[h] = 0.f;
    if (emb_cache) {
      // the tile's cached embeddings are one contiguous [<=128 x 64] block: coalesced 16-byte loads into sE
      const int rows = min(GB_TILE, total - t0);
      const float4* src = (const float4*)(emb_cache + ((size_t)g * total + t0) * 64);
      for (int i = tid; i < GB_TILE * 16; i += GB_TILE) {
        const int row = i >> 4, c4 = (i & 15) * 4;
        const float4 v = (row < rows) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float* d = sE + row * 65 + c4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
      sE[tid * 65 + 64] = 1.f;                         // bias column
      __syncthreads();
      if (e < total) {
        for (int h = 0; h < H; ++h) {
          float a = sW[H * 64 + h];
#pragma unroll 16
          for (int k = 0; k < 64; ++k) a = fmaf(sW[h * 64 + k], sE[tid * 65 + k], a);
          df[h] = (a > 1e-6f) ? dgbias[((size_t)g * total + e) * H + h] / a : 0.f;
        }
      }
#pragma unroll
      for (int h = 0; h < 8; ++h) sDf[tid * 8 + h] = df[h];
      __syncthreads();
      // thread (k = tid % 64, hf = tid / 64) owns column k of every head over every second pair: one E load and two
      // 16-byte df loads per 8 FMAs (the old (h,k)-per-thread form needed two loads per FMA); bias column: threads 0..7
      {
        const int k = tid & 63, hf = tid >> 6;
#pragma unroll 4
        for (int p = hf; p < GB_TILE; p += 2) {
          const float ev = sE[p * 65 + k];
          const float4 d0 = *(const float4*)(sDf + p * 8), d1 = *(const float4*)(sDf + p * 8 + 4);
          kacc[0] = fmaf(d0.x, ev, kacc[0]); kacc[1] = fmaf(d0.y, ev, kacc[1]);
          kacc[2] = fmaf(d0.z, ev, kacc[2]); kacc[3] = fmaf(d0.w, ev, kacc[3]);
          kacc[4] = fmaf(d1.x, ev, kacc[4]); kacc[5] = fmaf(d1.y, ev, kacc[5]);
          kacc[6] = fmaf(d1.z, ev, kacc[6]); kacc[7] = fmaf(d1.w, ev, kacc[7]);
        }
        if (tid < 8) {
          float bs = bacc;
          for (int p = 0; p < GB_TILE; ++p) bs += sDf[p * 8 + tid];
          bacc = bs;
        }
      }
      __syncthreads();
      continue;
    }
    float emb[64];
    if (e < total) {
      float f[8];
      {
        const int r = e / N, c = e % N;
        double gq[4];
        pos_feats(bb, r, c, gq);
        pair_pos_fc<8>(gq, sW, sW + H * 64, H, sDim, f, emb, dr, ek_seed(dr), (unsigned long long)g * total + e,
                       fast_trig != 0);
      }
      for (int h = 0; h < H; ++h)
        df[h] = (f[h] > 1e-6f) ? dgbias[((size_t)g * total + e) * H + h] / f[h] : 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k) emb[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) sE[tid * 65 + k] = emb[k];
    sE[tid * 65 + 64] = 1.f;                           // bias column
#pragma unroll
    for (int h = 0; h < 8; ++h) sDf[tid * 8 + h] = df[h];
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const int o = tid + a * GB_TILE;
      if (o < nout) {
        const int h = o / 65, k = o % 65;
        float s = acc[a];
        for (int p = 0; p < GB_TILE; ++p) s = fmaf(sDf[p * 8 + h], sE[p * 65 + k], s);
        acc[a] = s;
      }
    }
    __syncthreads();
  }
  float* pout = part + ((size_t)g * gridDim.y + blockIdx.y) * nout;
  if (emb_cache) {
    // combine the two pair-halves through shared memory (sE is free now), then write [h][65]
    __syncthreads();
    float* sK = sE;                                   // [2][64][8]
#pragma unroll
    for (int h = 0; h < 8; ++h) sK[((tid >> 6) * 64 + (tid & 63)) * 8 + h] = kacc[h];
    __syncthreads();
    for (int o = tid; o < H * 64; o += GB_TILE) {
      const int h = o / 64, k = o % 64;
      pout[h * 65 + k] = sK[k * 8 + h] + sK[(64 + k) * 8 + h];
    }
    if (tid < H) pout[tid * 65 + 64] = bacc;
    return;
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    const int o = tid + a * GB_TILE;
    if (o < nout) pout[o] = acc[a];
  }
}

// ------------------------------------------------------------------------------------------------
// scores + mask/bias + softmax   (one CTA per (image, head))
// ------------------------------------------------------------------------------------------------
constexpr int SC_DC = 64;       // dh chunk staged in shared memory

template <typename T>
__global__ void __launch_bounds__(256)
edge_softmax_fwd_kernel(const T* __restrict__ QKZ, long long ld, int D, const float* __restrict__ cond,
                        const float* __restrict__ lbias, const float* __restrict__ gbias, int N, int Kn, int H,
                        float* __restrict__ P) {
  ek_pdl_prologue();
  extern __shared__ float smf[];
  const int g = blockIdx.x, h = blockIdx.y;
  const int dh = D / H;
  float* S = smf;                              // [N][Kn]
  float* Qs = S + N * Kn;                      // [N][SC_DC+1]
  float* Ks = Qs + N * (SC_DC + 1);            // [Kn][SC_DC+1]
  const T* Qg = QKZ + (size_t)g * N * ld + h * dh;
  const T* Kg = QKZ + (size_t)g * N * ld + D + h * dh;
  const int tid = threadIdx.x;
  const int total = N * Kn;
  // each thread owns entries e = tid + r*256
  constexpr int MAXR = 64;                     // supports N*Kn <= 16384 (N = Kn = 126 -> 15876)
  float acc[MAXR];
  const int nr = (total + 255) / 256;
#pragma unroll
  for (int r = 0; r < MAXR; ++r) acc[r] = 0.f;
  for (int d0 = 0; d0 < dh; d0 += SC_DC) {
    __syncthreads();
    for (int e = tid; e < N * SC_DC; e += 256) {
      const int i = e / SC_DC, d = e % SC_DC;
      Qs[i * (SC_DC + 1) + d] = (d0 + d < dh) ? to_f32<T>(Qg[(size_t)i * ld + d0 + d]) : 0.f;
    }
    for (int e = tid; e < Kn * SC_DC; e += 256) {
      const int j = e / SC_DC, d = e % SC_DC;
      Ks[j * (SC_DC + 1) + d] = (d0 + d < dh) ? to_f32<T>(Kg[(size_t)j * ld + d0 + d]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      if (r < nr) {
        const int e = tid + r * 256;
        if (e < total) {
          const int i = e / Kn, j = e % Kn;
          const float* qp = Qs + i * (SC_DC + 1);
          const float* kp = Ks + j * (SC_DC + 1);
          float a = acc[r];
#pragma unroll 16
          for (int d = 0; d < SC_DC; ++d) a = fmaf(qp[d], kp[d], a);
          acc[r] = a;
        }
      }
    }
  }
  const float scale = 1.0f / sqrtf((float)dh);
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if (r < nr) {
      const int e = tid + r * 256;
      if (e < total) {
        float s = scale * acc[r];
        const size_t ge = (size_t)g * total + e;
        if (gbias) s += gbias[ge * H + h];
        if (cond) s = (cond[ge] > 0.f) ? s : NEG_MASK;
        if (lbias) s += lbias[ge];
        S[e] = s;
      }
    }
  }
  __syncthreads();
  // softmax over j, one warp per row i
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < N; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < Kn; j += 32) mx = fmaxf(mx, S[i * Kn + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Kn; j += 32) {
      const float ev = expf(S[i * Kn + j] - mx);
      S[i * Kn + j] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float* Pr = P + (((size_t)g * N + i) * H + h) * Kn;
    for (int j = lane; j < Kn; j += 32) Pr[j] = S[i * Kn + j] * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// aggregation + output bias + doubled ReLU residual  (one CTA per (image, 128-column slice))
// ------------------------------------------------------------------------------------------------
constexpr int AG_COLS = 128;
constexpr int AG_ROWS = 32;       // query rows per pass (P tile staged in smem)

template <typename T>
__global__ void __launch_bounds__(256)
edge_aggregate_fwd_kernel(const float* __restrict__ P, const T* __restrict__ QKZ, long long ld, int D,
                          const float* __restrict__ b_out, const float* __restrict__ Xin, int N, int Kn, int H,
                          float* __restrict__ Xout, T* __restrict__ XoutT, long long ldt,
                          uint8_t* __restrict__ mask, EkDrop dr) {
  ek_pdl_prologue();
  extern __shared__ float smf[];
  const unsigned long long sd = ek_seed(dr);
  float* Ps = smf;                               // [AG_ROWS][H*Kn]
  const int g = blockIdx.x;
  const int c0 = blockIdx.y * AG_COLS;
  const int tid = threadIdx.x;
  const int cl = tid & (AG_COLS - 1);            // column within slice
  const int half = tid >> 7;                     // 0/1: which half of the row tile
  const int HK = H * Kn;
  const T* Zg = QKZ + (size_t)g * N * ld + 2 * D + c0 + cl;
  const bool col_ok = (c0 + cl) < D;
  const float bo = col_ok ? b_out[c0 + cl] : 0.f;
  for (int i0 = 0; i0 < N; i0 += AG_ROWS) {
    const int rows = min(AG_ROWS, N - i0);
    __syncthreads();
    for (int e = tid; e < rows * HK; e += 256) Ps[e] = P[((size_t)g * N + i0) * HK + e];
    __syncthreads();
    float acc[AG_ROWS / 2];
#pragma unroll
    for (int r = 0; r < AG_ROWS / 2; ++r) acc[r] = 0.f;
    if (col_ok) {
      for (int h = 0; h < H; ++h) {
        for (int j = 0; j < Kn; ++j) {
          const float z = to_f32<T>(Zg[(size_t)j * ld + (size_t)h * D]);
          const float* pp = Ps + h * Kn + j;
#pragma unroll
          for (int r = 0; r < AG_ROWS / 2; ++r) {
            const int i = half * (AG_ROWS / 2) + r;
            if (i < rows) acc[r] = fmaf(pp[i * HK], z, acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < AG_ROWS / 2; ++r) {
        const int i = half * (AG_ROWS / 2) + r;
        if (i < rows) {
          const size_t row = (size_t)g * N + i0 + i;
          const float o = acc[r] + bo;
          float xn;
          if (mask == nullptr) {
            // plain attention output of GraphSelfAttentionLayer.forward (graph_att_layer.py:164-178): no doubling / ReLU
            xn = o + (Xin ? Xin[row * D + c0 + cl] : 0.f);
          } else {
            // train mode: Dropout(0.2) on the doubled output before the ReLU (graph_att.py:103-104)
            const float o2 = (o + o) * ek_drop_mult(dr, sd, row * D + c0 + cl);
            xn = (Xin ? Xin[row * D + c0 + cl] : 0.f) + fmaxf(o2, 0.f);
            mask[row * D + c0 + cl] = o2 > 0.f ? 1 : 0;
          }
          Xout[row * D + c0 + cl] = xn;
          if (XoutT) XoutT[row * ldt + c0 + cl] = from_f32<T>(xn);
        }
      }
    }
  }
}

// backward of the aggregation for one (image, column slice):
//   dout = 2 * mask * dXout                      (written to dOut for the b_out column-sum)
//   dZ[j,h,c] = sum_i P[i,h,j] dout[i,c]         (written into dQKZ[:, 2D + h*D + c])
//   dPpart[slice][g,i,h,j] = sum_{c in slice} dout[i,c] Z[j,h,c]
template <typename T>
__global__ void __launch_bounds__(256)
edge_aggregate_bwd_kernel(const float* __restrict__ dXout, const uint8_t* __restrict__ mask,
                          const float* __restrict__ P, const T* __restrict__ QKZ, long long ld, int D, int N, int Kn,
                          int H, T* __restrict__ dQKZ, float* __restrict__ dOut, float* __restrict__ dPpart,
                          float gscale) {
  ek_pdl_prologue();
  extern __shared__ float smf[];
  const int g = blockIdx.x;
  const int slice = blockIdx.y;
  const int c0 = slice * AG_COLS;
  const int tid = threadIdx.x;
  const int HK = H * Kn;
  float* dO = smf;                               // [N][AG_COLS+1]
  float* Zs = dO + N * (AG_COLS + 1);            // [32][AG_COLS+1]  (chunk of (h,j) rows)
  float* Ps = Zs + 32 * (AG_COLS + 1);           // [N][32]  P[i, chunk]
  // stage dout tile
  for (int e = tid; e < N * AG_COLS; e += 256) {
    const int i = e / AG_COLS, c = e % AG_COLS;
    float v = 0.f;
    if (c0 + c < D) {
      const size_t idx = ((size_t)g * N + i) * D + c0 + c;
      v = mask[idx] ? gscale * dXout[idx] : 0.f;      // gscale = 2 / (1 - p_dropout)
      dOut[idx] = v;
    }
    dO[i * (AG_COLS + 1) + c] = v;
  }
  const int cl = tid & (AG_COLS - 1);
  const int half = tid >> 7;
  const int warp = tid >> 5, lane = tid & 31;
  float* dPg = dPpart + ((size_t)slice * gridDim.x + g) * N * HK;
  for (int k0 = 0; k0 < HK; k0 += 32) {          // chunk of 32 (h,j) pairs
    const int kc = min(32, HK - k0);
    __syncthreads();
    for (int e = tid; e < kc * AG_COLS; e += 256) {
      const int kk = e / AG_COLS, c = e % AG_COLS;
      const int hj = k0 + kk, h = hj / Kn, j = hj % Kn;
      Zs[kk * (AG_COLS + 1) + c] =
          (c0 + c < D) ? to_f32<T>(QKZ[((size_t)g * N + j) * ld + 2 * D + (size_t)h * D + c0 + c]) : 0.f;
    }
    for (int e = tid; e < N * 32; e += 256) {
      const int i = e / 32, kk = e % 32;
      Ps[e] = (kk < kc) ? P[((size_t)g * N + i) * HK + k0 + kk] : 0.f;
    }
    __syncthreads();
    // dZ: thread (cl, half) handles kk = half, half+2, ...
    if (c0 + cl < D) {
      for (int kk = half; kk < kc; kk += 2) {
        float a = 0.f;
        for (int i = 0; i < N; ++i) a = fmaf(Ps[i * 32 + kk], dO[i * (AG_COLS + 1) + cl], a);
        const int hj = k0 + kk, h = hj / Kn, j = hj % Kn;
        dQKZ[((size_t)g * N + j) * ld + 2 * D + (size_t)h * D + c0 + cl] = from_f32<T>(a);
      }
    }
    // dP partial: warp w handles rows i = w, w+8, ...; lane = kk
    for (int i = warp; i < N; i += 8) {
      float a = 0.f;
      if (lane < kc) {
        const float* dr = dO + i * (AG_COLS + 1);
        const float* zr = Zs + lane * (AG_COLS + 1);
#pragma unroll 8
        for (int c = 0; c < AG_COLS; ++c) a = fmaf(dr[c], zr[c], a);
        dPg[(size_t)i * HK + k0 + lane] = a;
      }
    }
  }
}

// backward of scores/softmax for one (image, head):
//   dP = sum_slices dPpart ; ds = P * (dP - sum_j P dP)
//   dgbias[g,i,j,h] = ds ; dlbias_part[h][g,i,j] = ds ; score path: ds_m = cond > 0 ? ds : 0
//   dQ[i,:] = scale * sum_j ds_m[i,j] K[j,:] ; dK[j,:] = scale * sum_i ds_m[i,j] Q[i,:]
template <typename T>
__global__ void __launch_bounds__(256)
edge_softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dPpart, int nslices,
                        const T* __restrict__ QKZ, long long ld, int D, const float* __restrict__ cond, int N, int Kn,
                        int H, T* __restrict__ dQKZ, float* __restrict__ dlbias_part, float* __restrict__ dgbias) {
  ek_pdl_prologue();
  extern __shared__ float smf[];
  const int g = blockIdx.x, h = blockIdx.y;
  const int G = gridDim.x;
  const int dh = D / H;
  const int HK = H * Kn;
  float* dS = smf;                               // [N][Kn]
  float* Ts = dS + N * Kn;                       // [max(N,Kn)][SC_DC+1] staging of K or Q chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t total = (size_t)N * Kn;
  for (int i = warp; i < N; i += 8) {
    const float* Pr = P + (((size_t)g * N + i) * H + h) * Kn;
    float dot = 0.f;
    for (int j = lane; j < Kn; j += 32) {
      float dp = 0.f;
      for (int s = 0; s < nslices; ++s) dp += dPpart[(((size_t)s * G + g) * N + i) * HK + h * Kn + j];
      dS[i * Kn + j] = dp;
      dot = fmaf(Pr[j], dp, dot);
    }
    dot = warp_sum(dot);
    for (int j = lane; j < Kn; j += 32) {
      const float ds = Pr[j] * (dS[i * Kn + j] - dot);
      const size_t ge = (size_t)g * total + (size_t)i * Kn + j;
      if (dgbias) dgbias[ge * H + h] = ds;
      if (dlbias_part) dlbias_part[(size_t)h * G * total + ge] = ds;
      dS[i * Kn + j] = (cond && !(cond[ge] > 0.f)) ? 0.f : ds;
    }
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  const T* Qg = QKZ + (size_t)g * N * ld + h * dh;
  const T* Kg = QKZ + (size_t)g * N * ld + D + h * dh;
  T* dQg = dQKZ + (size_t)g * N * ld + h * dh;
  T* dKg = dQKZ + (size_t)g * N * ld + D + h * dh;
  for (int d0 = 0; d0 < dh; d0 += SC_DC) {
    // dQ chunk: stage K chunk
    __syncthreads();
    for (int e = tid; e < Kn * SC_DC; e += 256) {
      const int j = e / SC_DC, d = e % SC_DC;
      Ts[j * (SC_DC + 1) + d] = (d0 + d < dh) ? to_f32<T>(Kg[(size_t)j * ld + d0 + d]) : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < N * SC_DC; e += 256) {
      const int i = e / SC_DC, d = e % SC_DC;
      if (d0 + d < dh) {
        float a = 0.f;
        for (int j = 0; j < Kn; ++j) a = fmaf(dS[i * Kn + j], Ts[j * (SC_DC + 1) + d], a);
        dQg[(size_t)i * ld + d0 + d] = from_f32<T>(scale * a);
      }
    }
    // dK chunk: stage Q chunk
    __syncthreads();
    for (int e = tid; e < N * SC_DC; e += 256) {
      const int i = e / SC_DC, d = e % SC_DC;
      Ts[i * (SC_DC + 1) + d] = (d0 + d < dh) ? to_f32<T>(Qg[(size_t)i * ld + d0 + d]) : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < N * SC_DC; e += 256) {          // rows j >= Kn of dK are zero (keys clamped, Q9)
      const int j = e / SC_DC, d = e % SC_DC;
      if (d0 + d < dh) {
        float a = 0.f;
        if (j < Kn)
          for (int i = 0; i < N; ++i) a = fmaf(dS[i * Kn + j], Ts[i * (SC_DC + 1) + d], a);
        dKg[(size_t)j * ld + d0 + d] = from_f32<T>(scale * a);
      }
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers (called from api.cu)
// ------------------------------------------------------------------------------------------------
int ek_adj_prep_fwd_launch(const float* adj0, const float* adj1, int g_split, const float* w, int G, int N, int Kn,
                           int L, float* cond, float* lbias, cudaStream_t st) {
  ek_launch(adj_prep_fwd_kernel, G, 256, 0, st, adj0, adj1, g_split, w, N, Kn, L, cond, lbias);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_adj_prep_bwd_launch(const float* adj0, const float* adj1, int g_split, const float* dlbias_part, int nparts,
                           int G, int N, int Kn, int L, float* dw_part, cudaStream_t st) {
  ek_launch(adj_prep_bwd_kernel, G, 256, 0, st, adj0, adj1, g_split, dlbias_part, nparts, N, Kn, L, dw_part);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_adj_labels_fwd_launch(const int8_t* lab0, const int8_t* lab1, int g_split, int S, const float* w, int G, int N,
                             int Kn, int L, float* cond, float* lbias, cudaStream_t st) {
  EK_REQUIRE(N <= S && Kn <= N, EK_ERR_SHAPE, "adj_labels: N=%d Kn=%d S=%d", N, Kn, S);
  ek_launch(adj_labels_fwd_kernel, G, 256, 0, st, lab0, lab1, g_split, S, w, N, Kn, L, cond, lbias);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_adj_labels_bwd_launch(const int8_t* lab0, const int8_t* lab1, int g_split, int S, const float* dlbias_part,
                             int nparts, int G, int N, int Kn, int L, float* dw_part, cudaStream_t st) {
  EK_REQUIRE(N <= S && Kn <= N, EK_ERR_SHAPE, "adj_labels: N=%d Kn=%d S=%d", N, Kn, S);
  ek_launch(adj_labels_bwd_kernel, G, 256, 0, st, lab0, lab1, g_split, S, dlbias_part, nparts, N, Kn, L, dw_part);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

int ek_geom_bias_fwd_launch(const double* bb0, const double* bb1, int g_split, const float* Wp, const float* bp,
                            const float* dim_t, int G, int N, int Kn, int H, float* gbias, EkDrop dr, float* emb_cache,
                            int fast_trig, cudaStream_t st) {
  EK_REQUIRE(H <= 8, EK_ERR_UNSUPPORTED, "geom_bias: H=%d > 8", H);
  dim3 grid(G, ek_div_up(N * Kn, 128 * 4));
  if (fast_trig)
    ek_launch(geom_bias_fwd_fast_kernel, grid, 128, (H * 65 + 8) * sizeof(float), st, bb0, bb1, g_split, Wp, bp, dim_t, N,
              Kn, H, gbias, dr, emb_cache);
  else
    ek_launch(geom_bias_fwd_kernel, grid, 128, (H * 65 + 8) * sizeof(float), st, bb0, bb1, g_split, Wp, bp, dim_t, N, Kn, H,
              gbias, dr, emb_cache, fast_trig);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_geom_bias_bwd_parts() { return GB_SPLIT; }
int ek_geom_bias_bwd_launch(const double* bb0, const double* bb1, int g_split, const float* Wp, const float* bp,
                            const float* dim_t, int G, int N, int Kn, int H, const float* dgbias, float* part,
                            EkDrop dr, const float* emb_cache, int fast_trig, cudaStream_t st) {
  EK_REQUIRE(H <= 8, EK_ERR_UNSUPPORTED, "geom_bias: H=%d > 8", H);
  const size_t smem = (H * 65 + 8 + GB_TILE * 65 + GB_TILE * 8) * sizeof(float) + 16;
  ek_launch(geom_bias_bwd_kernel, dim3(G, GB_SPLIT), GB_TILE, smem, st, bb0, bb1, g_split, Wp, bp, dim_t, N, Kn, H, dgbias,
            part, dr, emb_cache, fast_trig);
  EK_CHECK_LAUNCH();
  return EK_OK;
}

template <typename T>
static int edge_softmax_fwd_t(const T* QKZ, long long ld, int D, const float* cond, const float* lbias,
                              const float* gbias, int G, int N, int Kn, int H, float* P, cudaStream_t st) {
  EK_REQUIRE(N * Kn <= 64 * 256, EK_ERR_UNSUPPORTED, "edge_softmax: N*K = %d too large", N * Kn);
  const size_t smem = ((size_t)N * Kn + (size_t)(N + Kn) * (SC_DC + 1)) * sizeof(float);
  static size_t configured = 0;
  auto kern = edge_softmax_fwd_kernel<T>;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "edge_softmax: smem %zu: %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  ek_launch(kern, dim3(G, H), 256, smem, st, QKZ, ld, D, cond, lbias, gbias, N, Kn, H, P);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_softmax_fwd_mma_launch(const bf16* QKZ, long long ld, int D, const float* cond, const float* lbias,
                              const float* gbias, int G, int N, int Kn, int H, float* P, bf16* Phl, int p16,
                              cudaStream_t st);
int ek_softmax_bwd_mma_launch(const float* P, const float* dPpart, int nslices, const bf16* QKZ, long long ld, int D,
                              const float* cond, int G, int N, int Kn, int H, bf16* dQKZ, float* dlbias_part,
                              float* dgbias, cudaStream_t st);

int ek_edge_softmax_fwd_launch(int is_bf16, const void* QKZ, long long ld, int D, const float* cond,
                               const float* lbias, const float* gbias, int G, int N, int Kn, int H, float* P,
                               void* Phl, cudaStream_t st) {
  if (is_bf16) {
    // is_bf16 == 3: Phl has a third plane that receives the attention weights as IEEE fp16
    const int rc = ek_softmax_fwd_mma_launch((const bf16*)QKZ, ld, D, cond, lbias, gbias, G, N, Kn, H, P, (bf16*)Phl,
                                             is_bf16 == 3 ? 1 : 0, st);
    if (rc != EK_ERR_UNSUPPORTED) return rc;
    if (Phl) { ek_set_error("edge_softmax: bf16 planes requested but the tensor-core kernel does not take this shape"); return EK_ERR_UNSUPPORTED; }
  }
  return is_bf16 ? edge_softmax_fwd_t<bf16>((const bf16*)QKZ, ld, D, cond, lbias, gbias, G, N, Kn, H, P, st)
                 : edge_softmax_fwd_t<float>((const float*)QKZ, ld, D, cond, lbias, gbias, G, N, Kn, H, P, st);
}

template <typename T>
static int edge_aggregate_fwd_t(const float* P, const T* QKZ, long long ld, int D, const float* b_out,
                                const float* Xin, int G, int N, int Kn, int H, float* Xout, T* XoutT, long long ldt,
                                uint8_t* mask, EkDrop dr, cudaStream_t st) {
  const size_t smem = (size_t)AG_ROWS * H * Kn * sizeof(float);
  static size_t configured = 0;
  auto kern = edge_aggregate_fwd_kernel<T>;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "edge_aggregate: smem %zu: %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  ek_launch(kern, dim3(G, ek_div_up(D, AG_COLS)), 256, smem, st, P, QKZ, ld, D, b_out, Xin, N, Kn, H, Xout, XoutT, ldt, mask,
                                                          dr);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_agg_fwd_mma_launch(const float* P, const bf16* QKZ, long long ld, int D, const float* b_out, const float* Xin,
                          int G, int N, int Kn, int H, float* Xout, bf16* XoutT, long long ldt, uint8_t* mask,
                          EkDrop dr, const bf16* Phl, const bf16* Z16, long long ldz16, cudaStream_t st);
int ek_agg_bwd_mma_launch(const float* dXout, const uint8_t* mask, const float* P, const bf16* QKZ, long long ld, int D,
                          int G, int N, int Kn, int H, bf16* dQKZ, float* dOut, float* dPpart, float gscale,
                          const bf16* Phl, cudaStream_t st);

int ek_edge_aggregate_fwd_launch(int is_bf16, const float* P, const void* QKZ, long long ld, int D, const float* b_out,
                                 const float* Xin, int G, int N, int Kn, int H, float* Xout, void* XoutT,
                                 long long ldt, uint8_t* mask, EkDrop dr, const void* Phl, const void* Z16,
                                 long long ldz16, cudaStream_t st) {
  if (is_bf16) {   // tensor-core kernel (edge_mma.cu); SIMT template only for shapes it does not take
    const int rc = ek_agg_fwd_mma_launch(P, (const bf16*)QKZ, ld, D, b_out, Xin, G, N, Kn, H, Xout, (bf16*)XoutT, ldt,
                                         mask, dr, (const bf16*)Phl, (const bf16*)Z16, ldz16, st);
    if (rc != EK_ERR_UNSUPPORTED) return rc;
    if (Z16) { ek_set_error("edge_aggregate_fwd: fp16 Z requested but the tensor-core kernel does not take this shape"); return EK_ERR_UNSUPPORTED; }
  }
  return is_bf16 ? edge_aggregate_fwd_t<bf16>(P, (const bf16*)QKZ, ld, D, b_out, Xin, G, N, Kn, H, Xout, (bf16*)XoutT,
                                              ldt, mask, dr, st)
                 : edge_aggregate_fwd_t<float>(P, (const float*)QKZ, ld, D, b_out, Xin, G, N, Kn, H, Xout,
                                               (float*)XoutT, ldt, mask, dr, st);
}

int ek_edge_num_slices(int D) { return ek_div_up(D, AG_COLS); }
int ek_agg_bwd_img_ok(int D, int N, int Kn, int H, int have_phl);
// number of dP partial slices ekaid_edge_aggregate_bwd writes (and ekaid_edge_softmax_bwd has to add up) for this call
int ek_edge_bwd_slices(int is_bf16, int D, int N, int Kn, int H, int have_phl) {
  if (is_bf16 && ek_agg_bwd_img_ok(D, N, Kn, H, have_phl)) return 1;
  return ek_edge_num_slices(D);
}

template <typename T>
static int edge_aggregate_bwd_t(const float* dXout, const uint8_t* mask, const float* P, const T* QKZ, long long ld,
                                int D, int G, int N, int Kn, int H, T* dQKZ, float* dOut, float* dPpart, float gscale,
                                cudaStream_t st) {
  const size_t smem = ((size_t)N * (AG_COLS + 1) + 32 * (AG_COLS + 1) + (size_t)N * 32) * sizeof(float);
  static size_t configured = 0;
  auto kern = edge_aggregate_bwd_kernel<T>;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "edge_aggregate_bwd: smem %zu: %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  ek_launch(kern, dim3(G, ek_div_up(D, AG_COLS)), 256, smem, st, dXout, mask, P, QKZ, ld, D, N, Kn, H, dQKZ, dOut, dPpart,
                                                          gscale);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_edge_aggregate_bwd_launch(int is_bf16, const float* dXout, const uint8_t* mask, const float* P, const void* QKZ,
                                 long long ld, int D, int G, int N, int Kn, int H, void* dQKZ, float* dOut,
                                 float* dPpart, float gscale, const void* Phl, cudaStream_t st) {
  if (is_bf16) {
    const int rc = ek_agg_bwd_mma_launch(dXout, mask, P, (const bf16*)QKZ, ld, D, G, N, Kn, H, (bf16*)dQKZ, dOut,
                                         dPpart, gscale, (const bf16*)Phl, st);
    if (rc != EK_ERR_UNSUPPORTED) return rc;
  }
  return is_bf16 ? edge_aggregate_bwd_t<bf16>(dXout, mask, P, (const bf16*)QKZ, ld, D, G, N, Kn, H, (bf16*)dQKZ, dOut,
                                              dPpart, gscale, st)
                 : edge_aggregate_bwd_t<float>(dXout, mask, P, (const float*)QKZ, ld, D, G, N, Kn, H, (float*)dQKZ,
                                               dOut, dPpart, gscale, st);
}

template <typename T>
static int edge_softmax_bwd_t(const float* P, const float* dPpart, int nslices, const T* QKZ, long long ld, int D,
                              const float* cond, int G, int N, int Kn, int H, T* dQKZ, float* dlbias_part,
                              float* dgbias, cudaStream_t st) {
  const int mx = N > Kn ? N : Kn;
  const size_t smem = ((size_t)N * Kn + (size_t)mx * (SC_DC + 1)) * sizeof(float);
  static size_t configured = 0;
  auto kern = edge_softmax_bwd_kernel<T>;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EK_REQUIRE(e == cudaSuccess, EK_ERR_CUDA, "edge_softmax_bwd: smem %zu: %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  ek_launch(kern, dim3(G, H), 256, smem, st, P, dPpart, nslices, QKZ, ld, D, cond, N, Kn, H, dQKZ, dlbias_part, dgbias);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_edge_softmax_bwd_launch(int is_bf16, const float* P, const float* dPpart, int nslices, const void* QKZ,
                               long long ld, int D, const float* cond, int G, int N, int Kn, int H, void* dQKZ,
                               float* dlbias_part, float* dgbias, cudaStream_t st) {
  if (is_bf16) {
    const int rc = ek_softmax_bwd_mma_launch(P, dPpart, nslices, (const bf16*)QKZ, ld, D, cond, G, N, Kn, H,
                                             (bf16*)dQKZ, dlbias_part, dgbias, st);
    if (rc != EK_ERR_UNSUPPORTED) return rc;
  }
  return is_bf16 ? edge_softmax_bwd_t<bf16>(P, dPpart, nslices, (const bf16*)QKZ, ld, D, cond, G, N, Kn, H,
                                            (bf16*)dQKZ, dlbias_part, dgbias, st)
                 : edge_softmax_bwd_t<float>(P, dPpart, nslices, (const float*)QKZ, ld, D, cond, G, N, Kn, H,
                                             (float*)dQKZ, dlbias_part, dgbias, st);
}
