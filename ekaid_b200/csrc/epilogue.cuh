// Epilogue description shared by the SIMT fp32 GEMM and the tcgen05 bf16 GEMM.
#pragma once
#include "common.cuh"

enum EkAct { EK_ACT_NONE = 0, EK_ACT_RELU = 1, EK_ACT_TANH = 2, EK_ACT_SIGMOID = 3 };

// v = (acc (+ bias[n])) * dropout(m,n) (+ addend[m,n]) (+ rowflag[m] ? rowb_alt[n] : rowb[(m / rowb_div) % rowb_mod, n]);
// v = act(v); C[m,n] = v (fp32, optional); Cb[m,n] = bf16(v) (optional)
struct EkEpilogue {
  const float* bias;
  const float* addend;
  long long ldadd;
  const float* rowb;
  long long ldrowb;
  int rowb_div;
  int rowb_mod;
  const uint8_t* rowflag;
  const float* rowb_alt;
  int act;
  EkDrop drop;        // applied to (acc + bias), before addend / row-broadcast / activation; index m*dropN + dropOff + n
  int dropN;
  int dropOff;
  float* C;
  long long ldc;
  bf16* Cb;           // 16-bit output #1 (bf16, or fp16 when cb_fmt = 1), columns n < cb_n1 (0 = all)
  long long ldcb;
  int cb_fmt;
  int cb_n1;
  bf16* Cb2;          // optional 16-bit output #2 (own format), columns n >= cb2_n0, stored at column n - cb2_n0
  long long ldcb2;
  int cb2_fmt;
  int cb2_n0;
  float* C2;          // optional second fp32 output: columns n >= c_n1, stored at column n - c_n1
  long long ldc2;
  int c_n1;
  int add_n1;         // addend only for columns n < add_n1 (0 = all)
};

__device__ __forceinline__ float ek_act(float v, int act) {
  switch (act) {
    case EK_ACT_RELU: return fmaxf(v, 0.f);
    case EK_ACT_TANH: return tanhf(v);
    case EK_ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

// scalar epilogue for element (m, n)
__device__ __forceinline__ void ek_epilogue_store(const EkEpilogue& e, long long m, int n, float acc) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.drop.seed) v *= ek_drop_mult(e.drop, ek_seed(e.drop), (unsigned long long)m * e.dropN + e.dropOff + n);
  if (e.addend && (e.add_n1 == 0 || n < e.add_n1)) v += e.addend[m * e.ldadd + n];
  if (e.rowb) {
    if (e.rowflag && e.rowflag[m]) v += __ldg(e.rowb_alt + n);
    else v += __ldg(e.rowb + (long long)((m / e.rowb_div) % e.rowb_mod) * e.ldrowb + n);
  }
  v = ek_act(v, e.act);
  if (e.C2 && n >= e.c_n1) e.C2[m * e.ldc2 + (n - e.c_n1)] = v;
  else if (e.C) e.C[m * e.ldc + n] = v;
  if (e.Cb && (e.cb_n1 == 0 || n < e.cb_n1)) ek_store16(e.Cb + m * e.ldcb + n, v, e.cb_fmt);
  if (e.Cb2 && n >= e.cb2_n0) ek_store16(e.Cb2 + m * e.ldcb2 + (n - e.cb2_n0), v, e.cb2_fmt);
}
