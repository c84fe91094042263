// Question path kernels (reference models/language_model.py:48-53, :106-115, :127-156).
// All sequence buffers are TIME-MAJOR: row = l * B + b, so each GRU step reads/writes a contiguous slab.
#include "common.cuh"

namespace {

// E[l*B + b, 0:ed] = emb[q[b,l]], E[l*B + b, ed:2ed] = emb_[q[b,l]]     (language_model.py:48-53)
template <typename T>
__global__ void embed_gather_kernel(const long long* __restrict__ q, const float* __restrict__ emb,
                                    const float* __restrict__ emb2, int B, int L, int ed, T* __restrict__ E) {
  ek_pdl_prologue();
  const int row = blockIdx.x;                 // l*B + b
  const int l = row / B, b = row % B;
  const long long tok = q[(size_t)b * L + l];
  for (int c = threadIdx.x; c < 2 * ed; c += blockDim.x) {
    const float v = (c < ed) ? emb[tok * ed + c] : emb2[tok * ed + (c - ed)];
    E[(size_t)row * 2 * ed + c] = from_f32<T>(v);
  }
}
// demb[v, c] = sum over tokens equal to v of dE[row, c]  (c < ed; the second table is frozen).
// One CTA per (vocabulary row, 32-column chunk): warp 0 builds the ordered list of matching rows (ballot compaction),
// four row lanes sum interleaved rows, fixed-order combine.  Deterministic, no atomics; the padding token (most of the
// positions) is spread over ed/32 CTAs instead of serialising one.
__global__ void __launch_bounds__(128)
embed_gather_bwd_kernel(const long long* __restrict__ q, const float* __restrict__ dE, long long ldde, int B, int L,
                        int ed, float* __restrict__ demb, int padding_idx) {
  ek_pdl_prologue();
  if ((int)blockIdx.x == padding_idx) {
    // nn.Embedding(padding_idx=ntoken) (language_model.py:26): that row never receives a gradient, whatever the batch holds
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    if (threadIdx.x < 32 && c < ed) demb[(size_t)blockIdx.x * ed + c] = 0.f;
    return;
  }
  extern __shared__ int rows[];        // up to B*L matching rows
  __shared__ int count;
  __shared__ float part[4][32];
  const int v = blockIdx.x;
  const int total = B * L;
  if (threadIdx.x < 32) {
    // eight token loads in flight per pass (one at a time made this scan -- 40 dependent L2 round trips at B*L = 1280 --
    // the whole 46 us of the kernel, on the serial tail of the step); the list order (ascending row) is unchanged
    constexpr int U = 8;
    int n = 0;
    for (int r0 = 0; r0 < total; r0 += 32 * U) {
      long long tok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {            // unguarded loads of a clamped row: nothing orders them behind a branch
        const int row = min(r0 + u * 32 + (int)threadIdx.x, total - 1);
        const int l = row / B, b = row - l * B;
        tok[u] = q[(size_t)b * L + l];
      }
      bool hit[U];
#pragma unroll
      for (int u = 0; u < U; ++u) hit[u] = (r0 + u * 32 + (int)threadIdx.x < total) && (tok[u] == v);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned m = __ballot_sync(0xffffffffu, hit[u]);
        if (hit[u]) rows[n + __popc(m & ((1u << threadIdx.x) - 1))] = r0 + u * 32 + (int)threadIdx.x;
        n += __popc(m);
      }
    }
    if (threadIdx.x == 0) count = n;
  }
  __syncthreads();
  const int n = count;
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  float s = 0.f;
  if (c < ed) {
    // The reference's questions are zero-padded, and id 0 is an ordinary table row (padding_idx is ntoken): more than half
    // of the B*L positions land on it, ~175 rows per row lane at batch 64.  Sixteen of its loads are in flight at a time
    // (with four, that one row was 40 us of serial L2 round trips); the order of the additions is unchanged.
#pragma unroll 16
    for (int k = rl; k < n; k += 4) s += dE[(size_t)rows[k] * ldde + c];
  }
  part[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && c < ed) demb[(size_t)v * ed + c] = ((part[0][lane] + part[1][lane]) + part[2][lane]) + part[3][lane];
}

// GRU cell (torch.nn.GRU gate order r,z,n).  gi, gh: [B, 3H] (biases included)
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r*gh_n), h = (1-z)*n + z*hprev
// saves (r, z, n, gh_n) in gates [B, 4H]
template <typename T>
__global__ void gru_cell_fwd_kernel(const float* __restrict__ gi, float* __restrict__ gh,
                                    const float* __restrict__ hprev, int B, int H, float* __restrict__ h,
                                    T* __restrict__ hT, float* __restrict__ gates, const float* __restrict__ gh_reset) {
  ek_pdl_prologue();
  const int total = B * H;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int b = e / H, c = e % H;
    const float* gib = gi + (size_t)b * 3 * H;
    float* ghb = gh + (size_t)b * 3 * H;
    const float r = sigmoidf_(gib[c] + ghb[c]);
    const float z = sigmoidf_(gib[H + c] + ghb[H + c]);
    const float ghn = ghb[2 * H + c];
    if (gh_reset) {      // re-arm the accumulator of the next step's split-K GEMM with the bias b_hh
      ghb[c] = gh_reset[c];
      ghb[H + c] = gh_reset[H + c];
      ghb[2 * H + c] = gh_reset[2 * H + c];
    }
    const float n = tanhf(gib[2 * H + c] + r * ghn);
    const float hp = hprev ? hprev[e] : 0.f;
    const float hv = (1.f - z) * n + z * hp;
    h[e] = hv;
    if (hT) hT[e] = from_f32<T>(hv);
    float* gs = gates + (size_t)b * 4 * H;
    gs[c] = r; gs[H + c] = z; gs[2 * H + c] = n; gs[3 * H + c] = ghn;
  }
}
// backward of one step. dh = total gradient wrt h_t.  Outputs dgi [B,3H] (fp32 + T copy), dgh [B,3H] (fp32 + T),
// dhprev [B,H] = dh * z  (the W_hh path is added by the caller's GEMM)
template <typename T>
__global__ void gru_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ gates,
                                    const float* __restrict__ hprev, int B, int H, float* __restrict__ dgi,
                                    float* __restrict__ dgh, T* __restrict__ dgiT, T* __restrict__ dghT,
                                    float* __restrict__ dhprev) {
  ek_pdl_prologue();
  const int total = B * H;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int b = e / H, c = e % H;
    const float* gs = gates + (size_t)b * 4 * H;
    const float r = gs[c], z = gs[H + c], n = gs[2 * H + c], ghn = gs[3 * H + c];
    const float d = dh[e];
    const float hp = hprev ? hprev[e] : 0.f;
    const float dn = d * (1.f - z);
    const float dz = d * (hp - n);
    const float dnp = dn * (1.f - n * n);
    const float drp = dnp * ghn * r * (1.f - r);
    const float dzp = dz * z * (1.f - z);
    const size_t o = (size_t)b * 3 * H;
    dgi[o + c] = drp; dgi[o + H + c] = dzp; dgi[o + 2 * H + c] = dnp;
    dgh[o + c] = drp; dgh[o + H + c] = dzp; dgh[o + 2 * H + c] = dnp * r;
    if (dgiT) { dgiT[o + c] = from_f32<T>(drp); dgiT[o + H + c] = from_f32<T>(dzp); dgiT[o + 2 * H + c] = from_f32<T>(dnp); }
    if (dghT) { dghT[o + c] = from_f32<T>(drp); dghT[o + H + c] = from_f32<T>(dzp); dghT[o + 2 * H + c] = from_f32<T>(dnp * r); }
    dhprev[e] = d * z;
  }
}

// out[row] = A[row,:] . w + b      (W2 of the question self-attention; one warp per row)
template <typename T>
__global__ void rowdot_kernel(const T* __restrict__ A, long long lda, long long M, int K, const float* __restrict__ w,
                              const float* __restrict__ b, float* __restrict__ out) {
  ek_pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < K; c += 32) s = fmaf(to_f32<T>(A[row * lda + c]), w[c], s);
  s = warp_sum(s);
  if (lane == 0) out[row] = s + (b ? b[0] : 0.f);
}

// Question attention pooling incl. quirk Q4 (language_model.py:149-153).
// a: [L*B] time-major logits (a_tm[l*B + b] = atten[b, l]).  S[l, :] = softmax over b.  The reference then
// re-views the contiguous [L,B] buffer as [B,1,L]:  Wt[b, l] = S_flat[b*L + l].
// qv[b, :] = sum_l Wt[b,l] * Hs[l*B + b, :]
__global__ void qpool_softmax_kernel(const float* __restrict__ a, int B, int L, float* __restrict__ S) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const int l = blockIdx.x;
  float mx = -INFINITY;
  for (int b = threadIdx.x; b < B; b += blockDim.x) mx = fmaxf(mx, a[(size_t)l * B + b]);
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int w = 0; w < (blockDim.x + 31) / 32; ++w) mx = fmaxf(mx, red[w]);
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += expf(a[(size_t)l * B + b] - mx);
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int b = threadIdx.x; b < B; b += blockDim.x) S[(size_t)l * B + b] = expf(a[(size_t)l * B + b] - mx) * inv;
}
__global__ void qpool_sum_kernel(const float* __restrict__ S, const float* __restrict__ Hs, int B, int L, int H,
                                 float* __restrict__ qv) {
  ek_pdl_prologue();
  const int b = blockIdx.x;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < H; c += gridDim.y * blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s = fmaf(S[(size_t)b * L + l], Hs[((size_t)l * B + b) * H + c], s);
    qv[(size_t)b * H + c] = s;
  }
}
// backward part 1 (one CTA per (b, l)): dWt[b,l] = dqv[b,:] . Hs[l*B+b,:]  -> dS_flat[b*L + l];
//                                        dHs[l*B+b, :] = Wt[b,l] * dqv[b,:]
__global__ void qpool_bwd1_kernel(const float* __restrict__ dqv, const float* __restrict__ S,
                                  const float* __restrict__ Hs, int B, int L, int H, float* __restrict__ dS,
                                  float* __restrict__ dHs) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const int b = blockIdx.x, l = blockIdx.y;
  const float wt = S[(size_t)b * L + l];
  const size_t row = ((size_t)l * B + b) * H;
  float s = 0.f;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float d = dqv[(size_t)b * H + c];
    s = fmaf(d, Hs[row + c], s);
    dHs[row + c] = wt * d;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) dS[(size_t)b * L + l] = s;
}
// backward part 2 (one CTA per l): softmax over b:  da[l,b] = S * (dS - sum_b S dS)
__global__ void qpool_bwd2_kernel(const float* __restrict__ S, const float* __restrict__ dS, int B, int L,
                                  float* __restrict__ da) {
  ek_pdl_prologue();
  __shared__ float red[32];
  const int l = blockIdx.x;
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s = fmaf(S[(size_t)l * B + b], dS[(size_t)l * B + b], s);
  s = block_sum(s, red);
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    da[(size_t)l * B + b] = S[(size_t)l * B + b] * (dS[(size_t)l * B + b] - s);
}
// dpre[row, c] = da[row] * w2[c] * (1 - a1[row,c]^2)     (tanh + W2 backward, output in GEMM operand type)
template <typename TA, typename T>     // TA: storage of the saved tanh output (fp32 keeps 1 - t^2 exact near saturation)
__global__ void qatt_tanh_bwd_kernel(const float* __restrict__ da, const float* __restrict__ w2,
                                     const TA* __restrict__ a1, long long M, int H, T* __restrict__ dpre,
                                     float* __restrict__ dpre32) {
  ek_pdl_prologue();
  const long long total = M * H;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / H;
    const int c = (int)(e % H);
    const float t = to_f32<TA>(a1[e]);
    const float v = da[r] * w2[c] * (1.f - t * t);
    dpre[e] = from_f32<T>(v);
    if (dpre32) dpre32[e] = v;       // unrounded copy: the bias gradient is a cancelling column sum of these
  }
}
// y += x  (fp32), used to accumulate gradient streams
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  ek_pdl_prologue();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    y[e] += x[e];
}

inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int ek_embed_gather_launch(int is_bf16, const long long* q, const float* emb, const float* emb2, int B, int L, int ed,
                           void* E, cudaStream_t st) {
  if (is_bf16 == 2) ek_launch(embed_gather_kernel<f16>, B * L, 128, 0, st, q, emb, emb2, B, L, ed, (f16*)E);
  else if (is_bf16) ek_launch(embed_gather_kernel<bf16>, B * L, 128, 0, st, q, emb, emb2, B, L, ed, (bf16*)E);
  else ek_launch(embed_gather_kernel<float>, B * L, 128, 0, st, q, emb, emb2, B, L, ed, (float*)E);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_embed_gather_bwd_launch(const long long* q, const float* dE, long long ldde, int B, int L, int ed, int V,
                               float* demb, int padding_idx, cudaStream_t st) {
  ek_launch(embed_gather_bwd_kernel, dim3(V, ek_div_up(ed, 32)), 128, (size_t)B * L * sizeof(int), st, q, dE, ldde, B, L, ed,
                                                                                                  demb, padding_idx);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_gru_cell_fwd_launch(int is_bf16, const float* gi, float* gh, const float* hprev, int B, int H, float* h,
                           void* hT, float* gates, const float* gh_reset, cudaStream_t st) {
  if (is_bf16 == 2)
    ek_launch(gru_cell_fwd_kernel<f16>, grid_for((long long)B * H), 256, 0, st, gi, gh, hprev, B, H, h, (f16*)hT, gates,
                                                                          gh_reset);
  else if (is_bf16)
    ek_launch(gru_cell_fwd_kernel<bf16>, grid_for((long long)B * H), 256, 0, st, gi, gh, hprev, B, H, h, (bf16*)hT, gates,
                                                                           gh_reset);
  else
    ek_launch(gru_cell_fwd_kernel<float>, grid_for((long long)B * H), 256, 0, st, gi, gh, hprev, B, H, h, (float*)hT, gates,
                                                                            gh_reset);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_gru_cell_bwd_launch(int is_bf16, const float* dh, const float* gates, const float* hprev, int B, int H,
                           float* dgi, float* dgh, void* dgiT, void* dghT, float* dhprev, cudaStream_t st) {
  if (is_bf16)
    ek_launch(gru_cell_bwd_kernel<bf16>, grid_for((long long)B * H), 256, 0, st, dh, gates, hprev, B, H, dgi, dgh, (bf16*)dgiT,
                                                                           (bf16*)dghT, dhprev);
  else
    ek_launch(gru_cell_bwd_kernel<float>, grid_for((long long)B * H), 256, 0, st, dh, gates, hprev, B, H, dgi, dgh,
                                                                            (float*)dgiT, (float*)dghT, dhprev);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_rowdot_launch(int is_bf16, const void* A, long long lda, long long M, int K, const float* w, const float* b,
                     float* out, cudaStream_t st) {
  if (is_bf16 == 2) ek_launch(rowdot_kernel<f16>, ek_div_up(M, 8), 256, 0, st, (const f16*)A, lda, M, K, w, b, out);
  else if (is_bf16) ek_launch(rowdot_kernel<bf16>, ek_div_up(M, 8), 256, 0, st, (const bf16*)A, lda, M, K, w, b, out);
  else ek_launch(rowdot_kernel<float>, ek_div_up(M, 8), 256, 0, st, (const float*)A, lda, M, K, w, b, out);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_qpool_fwd_launch(const float* a, const float* Hs, int B, int L, int H, float* S, float* qv, cudaStream_t st) {
  ek_launch(qpool_softmax_kernel, L, 256, 0, st, a, B, L, S);
  EK_CHECK_LAUNCH();
  ek_launch(qpool_sum_kernel, dim3(B, ek_div_up(H, 256)), 256, 0, st, S, Hs, B, L, H, qv);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_qpool_bwd_launch(const float* dqv, const float* S, const float* Hs, int B, int L, int H, float* dS, float* da,
                        float* dHs, cudaStream_t st) {
  ek_launch(qpool_bwd1_kernel, dim3(B, L), 256, 0, st, dqv, S, Hs, B, L, H, dS, dHs);
  EK_CHECK_LAUNCH();
  ek_launch(qpool_bwd2_kernel, L, 256, 0, st, S, dS, B, L, da);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_qatt_tanh_bwd_launch(int is_bf16, const float* da, const float* w2, const void* a1, long long M, int H,
                            void* dpre, float* dpre32, cudaStream_t st) {
  // is_bf16: 0 = fp32 a1 and dpre, 1 = bf16 both, 3 = fp32 a1 with a bf16 dpre (the 16-bit path keeps tanh outputs in fp32)
  if (is_bf16 == 3) ek_launch(qatt_tanh_bwd_kernel<float, bf16>, grid_for(M * H), 256, 0, st, da, w2, (const float*)a1, M, H, (bf16*)dpre, dpre32);
  else if (is_bf16) ek_launch(qatt_tanh_bwd_kernel<bf16, bf16>, grid_for(M * H), 256, 0, st, da, w2, (const bf16*)a1, M, H, (bf16*)dpre, dpre32);
  else ek_launch(qatt_tanh_bwd_kernel<float, float>, grid_for(M * H), 256, 0, st, da, w2, (const float*)a1, M, H, (float*)dpre, dpre32);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
int ek_add_inplace_launch(float* y, const float* x, long long n, cudaStream_t st) {
  if (n == 0) return EK_OK;
  ek_launch(add_inplace_kernel, grid_for(n), 256, 0, st, y, x, n);
  EK_CHECK_LAUNCH();
  return EK_OK;
}
