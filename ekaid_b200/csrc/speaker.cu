// Answer decoder (DynamicSpeaker / DynamicCore, models/dynamic_speaker_change_pos.py:94-131,182-240,287-357) and the
// masked language-model criterion (utils/utils.py:204-216): the per-step kernels around the tcgen05 GEMMs.
//
// One decode step = five dense products (gemm_tc.cu; M = batch rows, weights L2-resident) and the kernels below:
//   dec_lstm_fwd/bwd   LSTMCell point-wise part: gate pre-activations arrive as up to three partial products
//                      (+ a per-token table row for the word-embedding part in eval mode) plus both bias vectors
//   dec_att_fwd/bwd    one CTA per sample: module attention (weight_fc + softmax, weighted sum of bef/diff/aft),
//                      position branch (pos1 ReLU/Dropout -> weight_pos -> Dropout(0.5) -> softmax -> pos2)
//   dec_gate_fwd/bwd   gated_att_feat = sigmoid(gate2x(.)) * att_feat
//   dec_drop_op / dec_relu_drop_bwd   Dropout on an activation that feeds a GEMM, and its backward through ReLU
//   dec_embed          relu(Embedding[token]) (+ Dropout) as a zero-padded GEMM operand
//   dec_token          greedy sampling step on the device: log-softmax, arg-max, `unfinished` bookkeeping, next token
//                      (replaces the .sum() == 0 host synchronisation of :213 / :354)
//   dec_nll            log-softmax + masked NLL + its gradient in one pass over the logits of all steps
//   dec_outer_small    sum_r a[r, i] b[r, j] for the three tiny layers' weight gradients (i < 16)
// Operand buffers are written in the GEMM operand type: opf = 0 fp32 (parity path), 1 bf16 (tensor-core path).
#include "common.cuh"

namespace {

__device__ __forceinline__ void st_op(void* base, long long idx, float v, int opf) {
  if (opf) ((bf16*)base)[idx] = __float2bfloat16_rn(v);
  else ((float*)base)[idx] = v;
}
__device__ __forceinline__ float ld_op(const void* base, long long idx, int opf) {
  return opf ? __bfloat162float(((const bf16*)base)[idx]) : ((const float*)base)[idx];
}

// ---------------------------------------------------------------- word embedding operand
// out[r, j] = dropout(relu(emb[token(r), j])) for j < We, 0 for We <= j < ldo;  r = t * B + b, token = seq[b * sb + (t0 + t) * st]
__global__ void dec_embed_kernel(const long long* __restrict__ seq, long long sb, long long st, int t0, int B, int rows,
                                 const float* __restrict__ emb, int V, int We, void* __restrict__ out, long long ldo, int opf,
                                 EkDrop dr, int* __restrict__ err) {
  ek_pdl_prologue();
  const unsigned long long seedv = ek_seed(dr);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)rows * ldo;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / ldo;
    const int j = (int)(e % ldo);
    float v = 0.f;
    if (j < We) {
      const int t = (int)(r / B), b = (int)(r % B);
      long long tok = seq[b * sb + (long long)(t0 + t) * st];
      if (tok < 0 || tok >= V) {
        if (err) atomicExch(err, 1);
        tok = 0;
      }
      v = fmaxf(emb[tok * We + j], 0.f) * ek_drop_mult(dr, seedv, (unsigned long long)(r + (long long)t0 * B) * We + j);
    }
    st_op(out, r * ldo + j, v, opf);
  }
}

// ---------------------------------------------------------------- LSTM cell
// pre[b, g*R + u] = sum_k src_k[b * ld_k + g*R + u] (+ tbl[token_b, g*R + u]) + b1[g*R+u] + b2[g*R+u],  gate order i, f, g, o
// (torch.nn.LSTMCell).  Writes the activated gates (backward), c, h (fp32), h as operand (h_op) and, optionally, the
// dropped output F.dropout(h) as operand (out_op, dynamic_speaker_change_pos.py:126).
__global__ void dec_lstm_fwd_kernel(const float* __restrict__ s0, long long l0, const float* __restrict__ s1, long long l1,
                                    const float* __restrict__ s2, long long l2, const float* __restrict__ tbl,
                                    const long long* __restrict__ tok, const float* __restrict__ b1,
                                    const float* __restrict__ b2, const float* __restrict__ c_prev, int B, int R,
                                    float* __restrict__ gates, float* __restrict__ c_out, float* __restrict__ h_out,
                                    void* __restrict__ h_op, long long ldh, void* __restrict__ out_op, long long ldo, int opf,
                                    EkDrop dr, unsigned long long drop_base) {
  ek_pdl_prologue();
  const unsigned long long seedv = ek_seed(dr);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * R; e += gridDim.x * blockDim.x) {
    const int b = e / R, u = e % R;
    float pre[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = g * R + u;
      float v = b1[col] + b2[col];
      if (s0) v += s0[b * l0 + col];
      if (s1) v += s1[b * l1 + col];
      if (s2) v += s2[b * l2 + col];
      if (tbl) v += tbl[tok[b] * (long long)(4 * R) + col];
      pre[g] = v;
    }
    const float i = sigmoidf_(pre[0]), f = sigmoidf_(pre[1]), g = tanhf(pre[2]), o = sigmoidf_(pre[3]);
    const float cp = c_prev ? c_prev[e] : 0.f;
    const float c = f * cp + i * g;
    const float h = o * tanhf(c);
    float* gp = gates + (long long)b * 4 * R + u;
    gp[0] = i; gp[R] = f; gp[2 * R] = g; gp[3 * R] = o;
    c_out[e] = c;
    h_out[e] = h;
    if (h_op) st_op(h_op, (long long)b * ldh + u, h, opf);
    if (out_op) st_op(out_op, (long long)b * ldo + u, h * ek_drop_mult(dr, seedv, drop_base + e), opf);
  }
}

// dh = dh_a * dropout-mask (the dropped output's gradient) + dh_b + dh_c; dc = dc_in + dh o (1 - tanh(c)^2)
// -> pre-activation gradients as operand (dpre_op, row pitch ldp) and fp32 (dpre_f, for the bias sums), dc_prev
__global__ void dec_lstm_bwd_kernel(const float* __restrict__ dh_a, long long lda, EkDrop dr, unsigned long long drop_base,
                                    const float* __restrict__ dh_b, long long ldb, const float* __restrict__ dh_c,
                                    long long ldc, const float* __restrict__ dc_in, const float* __restrict__ gates,
                                    const float* __restrict__ c, const float* __restrict__ c_prev, int B, int R,
                                    void* __restrict__ dpre_op, long long ldp, int opf, float* __restrict__ dpre_f,
                                    float* __restrict__ dc_out) {
  ek_pdl_prologue();
  const unsigned long long seedv = ek_seed(dr);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * R; e += gridDim.x * blockDim.x) {
    const int b = e / R, u = e % R;
    float dh = 0.f;
    if (dh_a) dh += dh_a[b * lda + u] * ek_drop_mult(dr, seedv, drop_base + e);
    if (dh_b) dh += dh_b[b * ldb + u];
    if (dh_c) dh += dh_c[b * ldc + u];
    const float* gp = gates + (long long)b * 4 * R + u;
    const float i = gp[0], f = gp[R], g = gp[2 * R], o = gp[3 * R];
    const float tc = tanhf(c[e]);
    const float dc = (dc_in ? dc_in[e] : 0.f) + dh * o * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[e] : 0.f;
    const float d[4] = {dc * g * i * (1.f - i), dc * cp * f * (1.f - f), dc * i * (1.f - g * g), dh * tc * o * (1.f - o)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st_op(dpre_op, (long long)b * ldp + k * R + u, d[k], opf);
      if (dpre_f) dpre_f[(long long)b * 4 * R + k * R + u] = d[k];
    }
    dc_out[e] = dc * f;
  }
}

// ---------------------------------------------------------------- module attention + position branch (one CTA per sample)
constexpr int ATT_THREADS = 256;
constexpr int NPOS = 16;

struct DecAttW {
  const float* Wfc;    // [3, R]
  const float* bfc;    // [3]
  const float* bp1;    // [P]      (pos1 bias; its product with prev_h arrives in p1pre)
  const float* Wwp;    // [16, P]
  const float* bwp;    // [16]
  const float* Wp2;    // [R, 16]
  const float* bp2;    // [R]
};

// h_mod [B,R]; p1pre [B, *] (ld) = prev_h W_pos1^T; feats bef/diff/aft [B,D].
// Writes mw [B,4], pw [B,16], dposd [B,16] (the `output_pos` of the step: weight_pos output after Dropout(0.5)),
// vpos [B,P] (after ReLU + Dropout), att [B,D], and the operand gi2 [B, R + D] = [ppos | att_feat].
__global__ void __launch_bounds__(ATT_THREADS)
dec_att_fwd_kernel(const float* __restrict__ h_mod, const float* __restrict__ p1pre, long long ldp1, DecAttW w,
                   const float* __restrict__ bef, const float* __restrict__ diff, const float* __restrict__ aft, int R,
                   int P, int D, EkDrop dr1, EkDrop dr5, unsigned long long row_base, float* __restrict__ mw_out,
                   float* __restrict__ pw_out, float* __restrict__ dposd_out, float* __restrict__ vpos_out,
                   float* __restrict__ att_out, void* __restrict__ gi2, long long ldg, int opf) {
  ek_pdl_prologue();
  extern __shared__ float sm[];
  float* vpos = sm;             // [P]
  float* sc = sm + P;           // [32] scalars: fc logits 0..2, mw 4..6, dpos 8..23 -> pw
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long s1 = ek_seed(dr1), s5 = ek_seed(dr5);
  const unsigned long long rb = row_base + b;
  for (int j = threadIdx.x; j < P; j += ATT_THREADS) {
    const float v = fmaxf(p1pre[b * ldp1 + j] + w.bp1[j], 0.f) * ek_drop_mult(dr1, s1, rb * P + j);
    vpos[j] = v;
    vpos_out[(long long)b * P + j] = v;
  }
  if (warp < 3) {
    float a = 0.f;
    for (int u = lane; u < R; u += 32) a = fmaf(w.Wfc[warp * R + u], h_mod[(long long)b * R + u], a);
    a = warp_sum(a);
    if (lane == 0) sc[warp] = a + w.bfc[warp];
  }
  __syncthreads();
  for (int k = warp; k < NPOS; k += ATT_THREADS / 32) {
    float a = 0.f;
    for (int j = lane; j < P; j += 32) a = fmaf(w.Wwp[k * P + j], vpos[j], a);
    a = warp_sum(a);
    if (lane == 0) {
      const float v = (a + w.bwp[k]) * ek_drop_mult(dr5, s5, rb * NPOS + k);
      sc[8 + k] = v;
      dposd_out[(long long)b * NPOS + k] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float m = fmaxf(sc[0], fmaxf(sc[1], sc[2]));
    const float e0 = expf(sc[0] - m), e1 = expf(sc[1] - m), e2 = expf(sc[2] - m);
    const float inv = 1.f / (e0 + e1 + e2);
    sc[4] = e0 * inv; sc[5] = e1 * inv; sc[6] = e2 * inv;
    mw_out[b * 4 + 0] = sc[4]; mw_out[b * 4 + 1] = sc[5]; mw_out[b * 4 + 2] = sc[6]; mw_out[b * 4 + 3] = 0.f;
    float mx = sc[8];
    for (int k = 1; k < NPOS; ++k) mx = fmaxf(mx, sc[8 + k]);
    float s = 0.f;
    for (int k = 0; k < NPOS; ++k) { sc[8 + k] = expf(sc[8 + k] - mx); s += sc[8 + k]; }
    for (int k = 0; k < NPOS; ++k) { sc[8 + k] /= s; pw_out[b * NPOS + k] = sc[8 + k]; }
  }
  __syncthreads();
  for (int u = threadIdx.x; u < R; u += ATT_THREADS) {
    float a = w.bp2[u];
#pragma unroll
    for (int k = 0; k < NPOS; ++k) a = fmaf(w.Wp2[u * NPOS + k], sc[8 + k], a);
    st_op(gi2, (long long)b * ldg + u, a, opf);
  }
  const float m0 = sc[4], m1 = sc[5], m2 = sc[6];
  for (int j = threadIdx.x; j < D; j += ATT_THREADS) {
    const long long o = (long long)b * D + j;
    const float a = m0 * bef[o] + m1 * diff[o] + m2 * aft[o];
    att_out[o] = a;
    st_op(gi2, (long long)b * ldg + R + j, a, opf);
  }
}

// dgi2 [B, R + D] fp32 = [dppos | datt from gate1x], datt_g [B, D] (from the gate product).
// Accumulates dbef / ddiff / daft (+=), writes dfc [B,4], ddpos [B,16], dhmod_fc [B,R] (gradient of h_mod through weight_fc)
// and the pos1 pre-activation gradient as operand (dvp_op, pitch ldv).
__global__ void __launch_bounds__(ATT_THREADS)
dec_att_bwd_kernel(const float* __restrict__ dgi2, long long ldg, const float* __restrict__ datt_g, DecAttW w,
                   const float* __restrict__ bef, const float* __restrict__ diff, const float* __restrict__ aft,
                   const float* __restrict__ mw, const float* __restrict__ pw, const float* __restrict__ vpos, int R, int P,
                   int D, EkDrop dr1, EkDrop dr5, unsigned long long row_base, float* __restrict__ dbef,
                   float* __restrict__ ddiff, float* __restrict__ daft, float* __restrict__ dfc_out,
                   float* __restrict__ ddpos_out, float* __restrict__ dhmod_fc, void* __restrict__ dvp_op, long long ldv,
                   int opf) {
  ek_pdl_prologue();
  __shared__ float red[32];
  __shared__ float sc[48];      // dmw 0..2, dfc 4..6, dpw 8..23, ddpos 24..39
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long s5 = ek_seed(dr5);
  const unsigned long long rb = row_base + b;
  const float m0 = mw[b * 4], m1 = mw[b * 4 + 1], m2 = mw[b * 4 + 2];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int j = threadIdx.x; j < D; j += ATT_THREADS) {
    const long long o = (long long)b * D + j;
    const float da = datt_g[o] + dgi2[b * ldg + R + j];
    a0 = fmaf(da, bef[o], a0); a1 = fmaf(da, diff[o], a1); a2 = fmaf(da, aft[o], a2);
    dbef[o] += m0 * da; ddiff[o] += m1 * da; daft[o] += m2 * da;
  }
  a0 = block_sum(a0, red); a1 = block_sum(a1, red); a2 = block_sum(a2, red);
  // dpw[k] = sum_u Wp2[u, k] dppos[u]
  for (int k = warp; k < NPOS; k += ATT_THREADS / 32) {
    float a = 0.f;
    for (int u = lane; u < R; u += 32) a = fmaf(w.Wp2[u * NPOS + k], dgi2[b * ldg + u], a);
    a = warp_sum(a);
    if (lane == 0) sc[8 + k] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float dot = m0 * a0 + m1 * a1 + m2 * a2;
    sc[4] = m0 * (a0 - dot); sc[5] = m1 * (a1 - dot); sc[6] = m2 * (a2 - dot);
    dfc_out[b * 4] = sc[4]; dfc_out[b * 4 + 1] = sc[5]; dfc_out[b * 4 + 2] = sc[6]; dfc_out[b * 4 + 3] = 0.f;
    float pd = 0.f;
    for (int k = 0; k < NPOS; ++k) pd = fmaf(pw[b * NPOS + k], sc[8 + k], pd);
    for (int k = 0; k < NPOS; ++k) {
      const float v = pw[b * NPOS + k] * (sc[8 + k] - pd) * ek_drop_mult(dr5, s5, rb * NPOS + k);
      sc[24 + k] = v;
      ddpos_out[b * NPOS + k] = v;
    }
  }
  __syncthreads();
  for (int u = threadIdx.x; u < R; u += ATT_THREADS)
    dhmod_fc[(long long)b * R + u] = w.Wfc[u] * sc[4] + w.Wfc[R + u] * sc[5] + w.Wfc[2 * R + u] * sc[6];
  const float keep = (dr1.seed && dr1.p > 0.f) ? 1.f / (1.f - dr1.p) : 1.f;
  for (int j = threadIdx.x; j < P; j += ATT_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < NPOS; ++k) a = fmaf(w.Wwp[k * P + j], sc[24 + k], a);
    // vpos > 0  <=>  ReLU active and the element kept by Dropout
    st_op(dvp_op, (long long)b * ldv + j, vpos[(long long)b * P + j] > 0.f ? a * keep : 0.f, opf);
  }
}

// ---------------------------------------------------------------- gate product
// gate = sigmoid(pre) (pre already holds gate2x's bias); gated = gate * att  (operand)
__global__ void dec_gate_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ att, long long n,
                                    float* __restrict__ gate, void* __restrict__ gated, int opf) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float g = sigmoidf_(pre[e]);
    gate[e] = g;
    st_op(gated, e, g * att[e], opf);
  }
}
__global__ void dec_gate_bwd_kernel(const float* __restrict__ dgated, const float* __restrict__ gate,
                                    const float* __restrict__ att, long long n, float* __restrict__ datt_g,
                                    void* __restrict__ dpre, int opf) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float d = dgated[e], g = gate[e];
    datt_g[e] = d * g;
    st_op(dpre, e, d * att[e] * g * (1.f - g), opf);
  }
}

// ---------------------------------------------------------------- dropout on a GEMM operand / backward through ReLU + Dropout
// out[r, j] = x[(xmod ? r % xmod : r) * ldx + j] * mask(base + r * n + j)     (xmod: the same xmod source rows at every step)
__global__ void dec_drop_op_kernel(const float* __restrict__ x, long long ldx, int rows, int n, int xmod, EkDrop dr,
                                   unsigned long long base, void* __restrict__ out, long long ldo, int opf) {
  ek_pdl_prologue();
  const unsigned long long seedv = ek_seed(dr);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)rows * n;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / n;
    const int j = (int)(e % n);
    const long long xr = xmod ? r % xmod : r;
    st_op(out, r * ldo + j, x[xr * ldx + j] * ek_drop_mult(dr, seedv, base + e), opf);
  }
}
// dpre[r, j] = y[r, j] > 0 ? dy[r, j] * keep : 0     (y = Dropout(ReLU(pre)) as stored operand: > 0 <=> active and kept)
__global__ void dec_relu_drop_bwd_kernel(const float* __restrict__ dy, long long ldd, const void* __restrict__ y,
                                         long long ldy, int yf, int rows, int n, float keep, void* __restrict__ out,
                                         long long ldo, int opf, float* __restrict__ out_f, long long ldf) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)rows * n;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / n;
    const int j = (int)(e % n);
    const float v = ld_op(y, r * ldy + j, yf) > 0.f ? dy[r * ldd + j] * keep : 0.f;
    if (out) st_op(out, r * ldo + j, v, opf);
    if (out_f) out_f[r * ldf + j] = v;
  }
}
// acc[b, j] (+)= sum_t x[(t * B + b) * ldx + j] * mask(base + (t * B + b) * n + j)       (gradient of the step-invariant
// core.embed output: every step applies its own Dropout mask to it)
__global__ void dec_masked_sum_t_kernel(const float* __restrict__ x, long long ldx, int T, int B, int n, EkDrop dr,
                                        unsigned long long base, float* __restrict__ acc) {
  ek_pdl_prologue();
  const unsigned long long seedv = ek_seed(dr);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)B * n;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / n;
    const int j = (int)(e % n);
    float a = 0.f;
    for (int t = 0; t < T; ++t) {
      const long long r = (long long)t * B + b;
      a += x[r * ldx + j] * ek_drop_mult(dr, seedv, base + r * n + j);
    }
    acc[e] = a;
  }
}

// ---------------------------------------------------------------- sampling step (dynamic_speaker_change_pos.py:312-355)
// state[0] = 1 while the reference's loop would still be running (it breaks once every sequence has produced token 0);
// unfinished [B] u8.  One CTA; a warp per row: log-softmax over V logits, then either the arg-max (sample_max = 1: first
// maximum, like torch.max) or a multinomial draw from exp(logp / temperature) (sample_max = 0, :341-349; inverse CDF over
// the tokens in index order, one counter-based uniform per (row, step) from the device seed);  it = it * unfinished,
// seq[b, t] = it, seq_logprobs[b, t] = log-prob of the chosen token -- both only while state[0] -- next[b] = it.
__global__ void __launch_bounds__(1024)
dec_token_kernel(const float* __restrict__ logits, long long ldl, int B, int V, int t, int T, long long* __restrict__ seq,
                 float* __restrict__ seq_logp, unsigned char* __restrict__ unfinished, int* __restrict__ state,
                 long long* __restrict__ next_tok, float* __restrict__ logp_out, int multinomial, float inv_temp,
                 EkDrop rng) {
  ek_pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int running = state[0];
  const unsigned long long seedv = ek_seed(rng);
  int any = 0;
  for (int b = warp; b < B; b += nw) {
    const float* lg = logits + b * ldl;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, lg[v]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(lg[v] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (!multinomial) {
      for (int v = lane; v < V; v += 32) {
        float lp = lg[v] - lse;
        if (t == 0 && v == 0) lp = -INFINITY;                 // never sample NULL at the first step (:329-332)
        if (logp_out) logp_out[(long long)b * V + v] = lp;
        if (lp > best) { best = lp; bi = v; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
    } else {
      // weights w_v = exp(logp_v / temperature) in blocks of 32 consecutive tokens; total, then the first token whose
      // inclusive prefix sum reaches u * total
      // (relative to the largest allowed logit, so a small temperature cannot underflow every weight)
      float m2 = -INFINITY;
      for (int v = lane; v < V; v += 32)
        if (!(t == 0 && v == 0)) m2 = fmaxf(m2, lg[v]);
      m2 = warp_max(m2);
      float tot = 0.f;
      for (int v0 = 0; v0 < V; v0 += 32) {
        const int v = v0 + lane;
        float w = 0.f;
        if (v < V && !(t == 0 && v == 0)) w = expf((lg[v] - m2) * inv_temp);
        if (logp_out && v < V) logp_out[(long long)b * V + v] = (t == 0 && v == 0) ? -INFINITY : lg[v] - lse;
        tot += w;
      }
      tot = warp_sum(tot);
      const float u = (float)(ek_rand32(seedv, rng.site, (unsigned long long)b * (T + 1) + t) >> 8) * (1.0f / 16777216.0f);
      const float target = u * tot;
      float carry = 0.f;
      int pick = -1;
      for (int v0 = 0; v0 < V && pick < 0; v0 += 32) {
        const int v = v0 + lane;
        float w = 0.f;
        if (v < V && !(t == 0 && v == 0)) w = expf((lg[v] - m2) * inv_temp);
        float c = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, c, o);
          if (lane >= o) c += n;
        }
        c += carry;
        const unsigned hit = __ballot_sync(0xffffffffu, w > 0.f && c > target);
        if (hit) pick = v0 + __ffs(hit) - 1;
        carry = __shfl_sync(0xffffffffu, c, 31);
      }
      if (pick < 0) {                                         // u * total landed on the rounding of the last prefix sum
        for (int v = V - 1; v >= 0 && pick < 0; --v)
          if (!(t == 0 && v == 0) && expf((lg[v] - m2) * inv_temp) > 0.f) pick = v;
        if (pick < 0) pick = (t == 0) ? 1 : 0;
      }
      bi = pick;
      best = lg[pick] - lse;
    }
    if (lane == 0) {
      int un = (t == 0) ? (bi > 0) : (unfinished[b] && bi > 0);
      const long long it = un ? bi : 0;
      if (running) {
        seq[(long long)b * T + t] = it;
        seq_logp[(long long)b * T + t] = best;
        unfinished[b] = (unsigned char)un;
      } else {
        un = 0;
      }
      next_tok[b] = it;
      any |= un;
    }
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0 && running && !any) state[0] = 0;
}

// ---------------------------------------------------------------- masked NLL over all steps (utils/utils.py:204-216)
// row r = t * B + b (t < T steps that produced logits): logp = log_softmax(logits[r]); target = labels[b, t + 1];
// m = masks[b, t + 1].  mode bit 0: write logp into out[b, t, :] ([B, Tout, V], the `outputs` tensor of _forward);
// bit 1: row_loss[r] = -logp[target] * m; bit 2: dlogits[r, v] = (softmax_v - [v == target]) * m * gscale[0] * inv_msum[0]
// as operand with pitch ldd (padding columns zeroed).
__global__ void dec_nll_kernel(const float* __restrict__ logits, long long ldl, int rows, int B, int V,
                               const long long* __restrict__ labels, long long lsb, const float* __restrict__ masks,
                               long long msb, int mode, float* __restrict__ out, int Tout, float* __restrict__ row_loss,
                               const float* __restrict__ gscale, const float* __restrict__ inv_msum,
                               void* __restrict__ dlogits, long long ldd, int opf, int* __restrict__ err) {
  ek_pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int t = r / B, b = r % B;
    const float* lg = logits + r * ldl;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, lg[v]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(lg[v] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    long long tgt = labels[b * lsb + t + 1];
    if (tgt < 0 || tgt >= V) {
      if (err && lane == 0) atomicExch(err, 1);
      tgt = 0;
    }
    const float m = masks[b * msb + t + 1];
    if (mode & 1)
      for (int v = lane; v < V; v += 32) out[((long long)b * Tout + t) * V + v] = lg[v] - lse;
    if ((mode & 2) && lane == 0) row_loss[r] = -(lg[tgt] - lse) * m;
    if (mode & 4) {
      const float k = m * gscale[0] * inv_msum[0];
      for (int v = lane; v < ldd; v += 32) {
        float d = 0.f;
        if (v < V) d = (expf(lg[v] - lse) - (v == tgt ? 1.f : 0.f)) * k;
        st_op(dlogits, r * ldd + v, d, opf);
      }
    }
  }
}
// backward of log_softmax for callers that differentiate the log-probabilities themselves:
// dlogits[r, v] = dlogp[b, t, v] - exp(logp[b, t, v]) * sum_v' dlogp[b, t, v']      (operand, pitch ldd, padding zeroed)
__global__ void dec_lsm_bwd_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp, int rows, int B, int V,
                                   int Tout, void* __restrict__ dlogits, long long ldd, int opf) {
  ek_pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int t = r / B, b = r % B;
    const long long o = ((long long)b * Tout + t) * V;
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += dlogp[o + v];
    s = warp_sum(s);
    for (int v = lane; v < ldd; v += 32)
      st_op(dlogits, r * ldd + v, v < V ? dlogp[o + v] - expf(logp[o + v]) * s : 0.f, opf);
  }
}
// one CTA: res[0] = sum row_loss / sum mask, res[1] = 1 / sum mask   (mask summed over masks[b, 1 .. T])
__global__ void __launch_bounds__(1024)
dec_nll_reduce_kernel(const float* __restrict__ row_loss, int rows, const float* __restrict__ masks, long long msb, int B,
                      int T, float* __restrict__ res) {
  ek_pdl_prologue();
  __shared__ float red[32];
  float a = 0.f, m = 0.f;
  if (row_loss)
    for (int r = threadIdx.x; r < rows; r += blockDim.x) a += row_loss[r];
  for (int e = threadIdx.x; e < B * T; e += blockDim.x) m += masks[(e / T) * msb + (e % T) + 1];
  a = block_sum(a, red);
  m = block_sum(m, red);
  if (threadIdx.x == 0) {
    res[0] = a / m;
    res[1] = 1.f / m;
  }
}

// ---------------------------------------------------------------- tiny weight gradients
// part[c][i, j] = sum_{r in chunk c} a[r * lda + i] * b[r * ldb + j]   (i < m <= 16; transpose_out: stored as [j, i]).
// The rows are split over gridDim.y chunks; the caller adds the chunks up with the deterministic column-sum kernel.
__global__ void dec_outer_small_kernel(const float* __restrict__ a, long long lda, int m, const float* __restrict__ bm,
                                       long long ldb, int n, int rows, float* __restrict__ part, int transpose_out) {
  ek_pdl_prologue();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  __shared__ float as[64][16];
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int rb = r0; rb < r1; rb += 64) {
    const int nr = min(64, r1 - rb);
    __syncthreads();
    for (int e = threadIdx.x; e < nr * m; e += blockDim.x) as[e / m][e % m] = a[(long long)(rb + e / m) * lda + e % m];
    __syncthreads();
    if (j < n)
      for (int r = 0; r < nr; ++r) {
        const float bv = bm[(long long)(rb + r) * ldb + j];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < m) acc[i] = fmaf(as[r][i], bv, acc[i]);
      }
  }
  if (j >= n) return;
  float* p = part + (long long)blockIdx.y * m * n;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < m) {
      if (transpose_out) p[(long long)j * m + i] = acc[i];
      else p[(long long)i * n + j] = acc[i];
    }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

int ek_dec_embed_launch(const long long* seq, long long sb, long long st, int t0, int B, int rows, const float* emb, int V,
                        int We, void* out, long long ldo, int opf, EkDrop dr, int* err, cudaStream_t s) {
  EK_REQUIRE(ldo >= We && rows % B == 0, EK_ERR_SHAPE, "dec_embed: ldo=%lld We=%d rows=%d B=%d", ldo, We, rows, B);
  ek_launch(dec_embed_kernel, grid_for((long long)rows * ldo), 256, 0, s, seq, sb, st, t0, B, rows, emb, V, We, out, ldo, opf,
            dr, err);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_lstm_fwd_launch(const float* s0, long long l0, const float* s1, long long l1, const float* s2, long long l2,
                           const float* tbl, const long long* tok, const float* b1, const float* b2, const float* c_prev,
                           int B, int R, float* gates, float* c_out, float* h_out, void* h_op, long long ldh, void* out_op,
                           long long ldo, int opf, EkDrop dr, unsigned long long drop_base, cudaStream_t s) {
  ek_launch(dec_lstm_fwd_kernel, grid_for((long long)B * R), 256, 0, s, s0, l0, s1, l1, s2, l2, tbl, tok, b1, b2, c_prev, B,
            R, gates, c_out, h_out, h_op, ldh, out_op, ldo, opf, dr, drop_base);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_lstm_bwd_launch(const float* dh_a, long long lda, EkDrop dr, unsigned long long drop_base, const float* dh_b,
                           long long ldb, const float* dh_c, long long ldc, const float* dc_in, const float* gates,
                           const float* c, const float* c_prev, int B, int R, void* dpre_op, long long ldp, int opf,
                           float* dpre_f, float* dc_out, cudaStream_t s) {
  ek_launch(dec_lstm_bwd_kernel, grid_for((long long)B * R), 256, 0, s, dh_a, lda, dr, drop_base, dh_b, ldb, dh_c, ldc,
            dc_in, gates, c, c_prev, B, R, dpre_op, ldp, opf, dpre_f, dc_out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_att_fwd_launch(const float* h_mod, const float* p1pre, long long ldp1, const float* const* w7, const float* bef,
                          const float* diff, const float* aft, int B, int R, int P, int D, EkDrop dr1, EkDrop dr5,
                          unsigned long long row_base, float* mw, float* pw, float* dposd, float* vpos, float* att,
                          void* gi2, long long ldg, int opf, cudaStream_t s) {
  DecAttW w = {w7[0], w7[1], w7[2], w7[3], w7[4], w7[5], w7[6]};
  ek_launch(dec_att_fwd_kernel, B, ATT_THREADS, (size_t)(P + 32) * sizeof(float), s, h_mod, p1pre, ldp1, w, bef, diff, aft,
            R, P, D, dr1, dr5, row_base, mw, pw, dposd, vpos, att, gi2, ldg, opf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_att_bwd_launch(const float* dgi2, long long ldg, const float* datt_g, const float* const* w7, const float* bef,
                          const float* diff, const float* aft, const float* mw, const float* pw, const float* vpos, int B,
                          int R, int P, int D, EkDrop dr1, EkDrop dr5, unsigned long long row_base, float* dbef, float* ddiff,
                          float* daft, float* dfc, float* ddpos, float* dhmod_fc, void* dvp_op, long long ldv, int opf,
                          cudaStream_t s) {
  DecAttW w = {w7[0], w7[1], w7[2], w7[3], w7[4], w7[5], w7[6]};
  ek_launch(dec_att_bwd_kernel, B, ATT_THREADS, 0, s, dgi2, ldg, datt_g, w, bef, diff, aft, mw, pw, vpos, R, P, D, dr1, dr5,
            row_base, dbef, ddiff, daft, dfc, ddpos, dhmod_fc, dvp_op, ldv, opf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_gate_fwd_launch(const float* pre, const float* att, long long n, float* gate, void* gated, int opf,
                           cudaStream_t s) {
  ek_launch(dec_gate_fwd_kernel, grid_for(n), 256, 0, s, pre, att, n, gate, gated, opf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_gate_bwd_launch(const float* dgated, const float* gate, const float* att, long long n, float* datt_g, void* dpre,
                           int opf, cudaStream_t s) {
  ek_launch(dec_gate_bwd_kernel, grid_for(n), 256, 0, s, dgated, gate, att, n, datt_g, dpre, opf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_lsm_bwd_launch(const float* dlogp, const float* logp, int rows, int B, int V, int Tout, void* dlogits,
                          long long ldd, int opf, cudaStream_t s) {
  if (rows == 0) return EK_OK;
  ek_launch(dec_lsm_bwd_kernel, grid_for((long long)rows * 32), 256, 0, s, dlogp, logp, rows, B, V, Tout, dlogits, ldd, opf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_drop_op_launch(const float* x, long long ldx, int rows, int n, int xmod, EkDrop dr, unsigned long long base,
                          void* out, long long ldo, int opf, cudaStream_t s) {
  ek_launch(dec_drop_op_kernel, grid_for((long long)rows * n), 256, 0, s, x, ldx, rows, n, xmod, dr, base, out, ldo, opf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_relu_drop_bwd_launch(const float* dy, long long ldd, const void* y, long long ldy, int yf, int rows, int n,
                                float keep, void* out, long long ldo, int opf, float* out_f, long long ldf, cudaStream_t s) {
  ek_launch(dec_relu_drop_bwd_kernel, grid_for((long long)rows * n), 256, 0, s, dy, ldd, y, ldy, yf, rows, n, keep, out, ldo,
            opf, out_f, ldf);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_masked_sum_t_launch(const float* x, long long ldx, int T, int B, int n, EkDrop dr, unsigned long long base,
                               float* acc, cudaStream_t s) {
  ek_launch(dec_masked_sum_t_kernel, grid_for((long long)B * n), 256, 0, s, x, ldx, T, B, n, dr, base, acc);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_token_launch(const float* logits, long long ldl, int B, int V, int t, int T, long long* seq, float* seq_logp,
                        unsigned char* unfinished, int* state, long long* next_tok, float* logp_out, int multinomial,
                        float temperature, EkDrop rng, cudaStream_t s) {
  EK_REQUIRE(!multinomial || (temperature > 0.f && rng.seed), EK_ERR_SHAPE, "dec_token: sampling needs temperature > 0 and a seed");
  ek_launch(dec_token_kernel, 1, 1024, 0, s, logits, ldl, B, V, t, T, seq, seq_logp, unfinished, state, next_tok, logp_out,
            multinomial, multinomial ? 1.0f / temperature : 1.0f, rng);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_nll_launch(const float* logits, long long ldl, int rows, int B, int V, const long long* labels, long long lsb,
                      const float* masks, long long msb, int mode, float* out, int Tout, float* row_loss,
                      const float* gscale, const float* inv_msum, void* dlogits, long long ldd, int opf, int* err,
                      cudaStream_t s) {
  if (rows == 0) return EK_OK;
  ek_launch(dec_nll_kernel, grid_for((long long)rows * 32), 256, 0, s, logits, ldl, rows, B, V, labels, lsb, masks, msb, mode,
            out, Tout, row_loss, gscale, inv_msum, dlogits, ldd, opf, err);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_nll_reduce_launch(const float* row_loss, int rows, const float* masks, long long msb, int B, int T, float* res,
                             cudaStream_t s) {
  ek_launch(dec_nll_reduce_kernel, 1, 1024, 0, s, row_loss, rows, masks, msb, B, T, res);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_dec_outer_small_launch(const float* a, long long lda, int m, const float* b, long long ldb, int n, int rows,
                              float* part, int nchunks, int transpose_out, cudaStream_t s) {
  EK_REQUIRE(m >= 1 && m <= 16 && nchunks >= 1, EK_ERR_SHAPE, "dec_outer_small: m=%d (1..16) chunks=%d", m, nchunks);
  ek_launch(dec_outer_small_kernel, dim3((n + 127) / 128, nchunks), 128, 0, s, a, lda, m, b, ldb, n, rows, part,
            transpose_out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
