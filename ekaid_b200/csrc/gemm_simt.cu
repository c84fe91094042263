// fp32 SIMT GEMM with generic operand strides -- the "fp32 path" (1e-4 parity mode) of every dense
// contraction on the EKAID hot path.  C[M,N] = sum_k A(m,k) * B(k,n) with
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]
// so the same kernel serves  y = x W^T (forward), dx = dy W (dgrad) and dW = dy^T x (wgrad).
// Register-tiled (TM x TM per thread), BK = 16, register prefetch of the next k-slab.
#include "common.cuh"
#include "epilogue.cuh"

namespace {

template <int TM>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, long long sam, long long sak,
                const float* __restrict__ B, long long sbk, long long sbn, EkEpilogue ep) {
  ek_pdl_prologue();
  constexpr int BT = 16 * TM;   // block tile (square)
  constexpr int BK = 16;
  constexpr int PER = BT * BK / 256;   // elements per thread per operand per slab
  __shared__ float As[BK][BT + 4];
  __shared__ float Bs[BK][BT + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * BT;
  const int n0 = blockIdx.x * BT;
  const bool a_kfast = (sak == 1);
  const bool b_kfast = (sbk == 1);

  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;

  float ra[PER], rb[PER];
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int idx = tid + p * 256;
      int mm, kk;
      if (a_kfast) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BT; kk = idx / BT; }
      const long long m = m0 + mm;
      const int k = k0 + kk;
      ra[p] = (m < M && k < K) ? __ldg(A + m * sam + (long long)k * sak) : 0.f;
      int nn, kb;
      if (b_kfast) { kb = idx % BK; nn = idx / BK; } else { nn = idx % BT; kb = idx / BT; }
      const int n = n0 + nn;
      const int k2 = k0 + kb;
      rb[p] = (n < N && k2 < K) ? __ldg(B + (long long)k2 * sbk + (long long)n * sbn) : 0.f;
    }
  };
  auto store_slab = [&]() {
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int idx = tid + p * 256;
      int mm, kk;
      if (a_kfast) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BT; kk = idx / BT; }
      As[kk][mm] = ra[p];
      int nn, kb;
      if (b_kfast) { kb = idx % BK; nn = idx / BK; } else { nn = idx % BT; kb = idx / BT; }
      Bs[kb][nn] = rb[p];
    }
  };

  load_slab(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    store_slab();
    __syncthreads();
    if (k0 + BK < K) load_slab(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TM; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N) ek_epilogue_store(ep, m, n, acc[i][j]);
    }
  }
}

}  // namespace

int ek_gemm_f32_launch(int M, int N, int K, const float* A, long long sam, long long sak, const float* B,
                       long long sbk, long long sbn, const EkEpilogue& ep, cudaStream_t stream) {
  EK_REQUIRE(M > 0 && N > 0 && K > 0, EK_ERR_SHAPE, "gemm_f32: bad shape M=%d N=%d K=%d", M, N, K);
  if ((long long)M * N >= 128 * 128 * 64) {
    dim3 grid(ek_div_up(N, 128), ek_div_up(M, 128));
    ek_launch(gemm_f32_kernel<8>, grid, 256, 0, stream, M, N, K, A, sam, sak, B, sbk, sbn, ep);
  } else {
    dim3 grid(ek_div_up(N, 64), ek_div_up(M, 64));
    ek_launch(gemm_f32_kernel<4>, grid, 256, 0, stream, M, N, K, A, sam, sak, B, sbk, sbn, ep);
  }
  EK_CHECK_LAUNCH();
  return EK_OK;
}
