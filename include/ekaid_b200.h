/* ekaid_b200 -- C ABI of the B200-native EKAID graph+fusion hot path.
 *
 * The reference (Holipori/EKAID) is pure Python: it has no FFI/plugin layer, its "operators" are ATen calls
 * made from nn.Module.forward.  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md).  Each entry point names the reference call site(s) it replaces; paths are relative to
 * /root/reference/model.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless stated;
 *   - the caller owns every buffer (outputs and workspaces); nothing here allocates or synchronises;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); CUDA-graph capturable;
 *   - return value 0 = ok, negative = error (EKAID_ERR_*), message via ekaid_last_error();
 *   - `is_bf16` selects the storage type of GEMM-operand activations: 0 = float, 1 = __nv_bfloat16
 *     (fp32 parity path vs. bf16 tensor-core path); reductions / softmax / residuals are always fp32;
 *   - sequence buffers of the question path are time-major (row = l*B + b);
 *   - "stacked" image batches: G = number of images; images [0, g_split) read adj0/bb0, the rest adj1/bb1
 *     (main and reference image of each pair share weights and are processed in one launch).
 */
#ifndef EKAID_B200_H
#define EKAID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EKAID_OK 0
#define EKAID_ERR_SHAPE -1
#define EKAID_ERR_ALIGN -2
#define EKAID_ERR_ARCH -3
#define EKAID_ERR_CUDA -4
#define EKAID_ERR_UNSUPPORTED -5

#define EKAID_ACT_NONE 0
#define EKAID_ACT_RELU 1
#define EKAID_ACT_TANH 2
#define EKAID_ACT_SIGMOID 3

/* Fused GEMM epilogue:  v = (acc (+ bias[n])) * dropout (+ addend[m,n])
 *                           (+ rowflag[m] ? rowb_alt[n] : rowb[(m / rowb_div) % rowb_mod, n]);  v = act(v);
 *                       C[m,n] = v (fp32, may be NULL);  Cb[m,n] = bf16(v) (may be NULL).
 * rowb/rowflag implement q_expand_v_cat (models/relation_encoder.py:19-29, quirk Q10) without materialising
 * the [B,N,2048] concat: the question half of self_weights is applied once per sample and broadcast per row. */
typedef struct ekaid_epilogue {
  const float* bias;
  const float* addend;
  int64_t ldadd;
  const float* rowb;
  int64_t ldrowb;
  int32_t rowb_div;
  int32_t rowb_mod;
  const uint8_t* rowflag;
  const float* rowb_alt;
  int32_t act;
  /* train-mode dropout applied to (acc + bias), before addend / row broadcast / activation (embed: Linear ->
   * Dropout(0.5) -> ReLU, modules.py:105-111; masked dgrad of a dropped GEMM input): element (m,n) uses counter
   * m*drop_n + drop_off + n of site drop_site; drop_seed = device pointer to the step seed, NULL = off */
  const uint64_t* drop_seed;
  uint32_t drop_site;
  float drop_p;
  int32_t drop_n;
  int32_t drop_off;
  float* C;
  int64_t ldc;
  void* Cb; /* 16-bit output #1: __nv_bfloat16* (cb_fmt 0) or __half* (cb_fmt 1, saturating conversion) */
  int64_t ldcb;
  /* 16-bit outputs may be split by column: Cb receives columns n < cb_n1 (0 = all); Cb2 (optional second 16-bit
   * output, own format) receives columns n >= cb2_n0 at Cb2[m * ldcb2 + n - cb2_n0].  cb_n1 and cb2_n0 are multiples
   * of 32.  Used to keep one tensor in two formats (e.g. Z as bf16 for the backward kernels and as fp16 for the
   * forward aggregation) or to route [query | key] and Z to different buffers from ONE GEMM. */
  int32_t cb_fmt;
  int32_t cb_n1;
  void* Cb2;
  int64_t ldcb2;
  int32_t cb2_fmt;
  int32_t cb2_n0;
  /* the fp32 output may be split by column too: C2 (optional) receives columns n >= c_n1 at C2[m * ldc2 + n - c_n1], C the
   * rest; the addend is applied to columns n < add_n1 only (0 = all).  c_n1 and add_n1 are multiples of 32.  ONE GEMM then
   * yields both halves of d[v | q] = mask * (dSf W_sw) in the backward of self_weights (graph_att.py:80): the node half plus
   * the residual gradient into dX, the question half into its own buffer. */
  float* C2;
  int64_t ldc2;
  int32_t c_n1;
  int32_t add_n1;
} ekaid_epilogue_t;

/* Dropout convention (train mode; eval / p = 0 passes seed = NULL): masks are never stored.  Every dropout site of the
 * reference has a site id; element idx of that site is kept iff hash(*seed, site, idx) >= p * 2^32 and scaled by
 * 1/(1-p).  The seed lives in device memory (ekaid_rng_advance once per step) so a captured CUDA graph draws fresh
 * masks at every replay; forward and backward regenerate identical masks from (seed, site, idx). */

int ekaid_abi_version(void);
const char* ekaid_last_error(void);
/* 0 if the current device is sm_100 (B200), EKAID_ERR_ARCH otherwise */
int ekaid_check_device(void);
/* Programmatic dependent launch between consecutive kernels of a stream (default on; EKAID_B200_PDL=0 disables).
 * Returns the previous setting.  Ordering and results are identical either way; only launch latency overlaps. */
int ekaid_set_pdl(int on);

/* ---- dense contractions ---------------------------------------------------------------------------------
 * C[M,N] = op(A) op(B):  transA = 0: A is [M,K] (lda), 1: A is [K,M];  transB = 0: B is [N,K] (ldb) i.e. an
 * nn.Linear weight, 1: B is [K,N].  Replaces every addmm/mm on the path: modules.py:195-196 (img),
 * graph_att.py:80 (self_weights), graph_att_layer.py:79,90,174 (query, key, linear_out_2),
 * modules.py:278-288 (context/gate), :300-303 (embed), language_model.py:113 (GRU), :138 (W1), and their
 * autograd backward (dgrad: transB=1, wgrad: transA=1,transB=1). */
int ekaid_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int64_t lda, const float* B,
                   int64_t ldb, const ekaid_epilogue_t* ep, void* stream);
/* bf16 operands, fp32 accumulation in TMEM (tcgen05.mma, TMA-fed).  force_bn: 0 = auto, else 64/128/256, or
 * 1128/1256 for the CTA-pair variant (cluster of 2, TMA multicast of the shared weight tile).
 * splits: 0 = auto split-K (plain fp32 C only), 1 = off. Operands 16-byte aligned, pitches multiples of 8. */
int ekaid_gemm_bf16(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B,
                    int64_t ldb, const ekaid_epilogue_t* ep, int force_bn, int splits, void* stream);
/* The same kernel with the 16-bit format chosen per operand (tcgen05 kind::f16 takes fp16 or bf16 for A and for B
 * independently): a_fp16 / b_fp16 = 1 when that operand holds IEEE fp16.  The forward pass keeps weights and bounded
 * activations in fp16 (11 significant bits), gradients stay bf16 (range); wgrad / dgrad GEMMs mix the two. */
int ekaid_gemm_tc(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, int a_fp16, const void* B,
                  int64_t ldb, int b_fp16, const ekaid_epilogue_t* ep, int force_bn, int splits, void* stream);
/* Measurement hook (scripts/gemm_probe.py): flags isolate one pipeline role of the GEMM kernel (16 = no fused epilogue,
 * 32 = no TMA loads, 64 = no MMAs, 256 = no TMEM drain; results are then garbage), 128 = early programmatic-launch
 * trigger; ts = device buffer of grid x 32 uint64 globaltimer stamps or NULL.  0 / NULL restores normal operation. */
int ekaid_gemm_debug(int flags, void* ts);
/* how many products ekaid_gemm_tc has routed to the M <= 64 warp-MMA kernel (gemm_skinny.cu) so far: the answer decoder's
 * per-step products, models/dynamic_speaker_change_pos.py:94-131 (tests; EKAID_B200_SKINNY=0 turns the routing off) */
int ekaid_gemm_skinny_count(void);

/* ---- casts / reductions / glue -------------------------------------------------------------------------- */
int ekaid_cast_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream);
/* fp32 operand -> three bf16 planes (hi, lo) for the split-precision tensor-core product of the fp32 parity path:
 * A B^T ~= Al Bh^T + Ah Bl^T + Ah Bh^T = one bf16 GEMM over a 3x longer contraction axis (every addmm/mm of the reference
 * in fp32, e.g. models/fc.py:25-32, on tcgen05 instead of SIMT FMA).  pattern 0 = planes (lo, hi, hi) [A operand],
 * 1 = (hi, lo, hi) [B operand]; along_rows 0: dst [rows, 3*cols], 1: dst [3*rows, cols]. */
int ekaid_split3_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, int pattern,
                      int along_rows, void* stream);
int ekaid_cast_bf16_f32(const void* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, void* stream);
/* dst = scale * float(src): the receiving side of the bf16 gradient exchange (scale = 1 / world size after a SUM
 * all-reduce, which -- unlike AVG -- NCCL can run inside the NVSwitch) */
int ekaid_cast_bf16_f32_scaled(const void* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, float scale,
                               void* stream);
/* fp32 -> IEEE fp16, round to nearest, saturating at +-65504 (forward operands of the 16-bit path: weights, bounded
 * activations) */
int ekaid_cast_f32_f16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream);
/* the same for up to 16 strided 2-D blocks in ONE launch.  src, lds, dst, ldd, rows, cols, mode are HOST arrays read at
 * call time; mode[i]: 0 = fp32 -> bf16, 1 = fp32 -> fp32, 2 = raw bytes (cols = bytes per row, multiple of 16, pitches in
 * bytes), 3 = fp32 -> fp16, 4 = fp16 -> bf16 (cols multiple of 8; the backward's bf16 copy of a forward activation).  Used for the operand-type copies of a relation encoder's weights and for staging the step's inputs. */
int ekaid_cast_many(int count, const void* const* src, const int64_t* lds, void* const* dst, const int64_t* ldd,
                    const int64_t* rows, const int32_t* cols, const int32_t* mode, void* stream);
int ekaid_copy_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, void* stream);
/* out[n] = sum_m rowscale[m] * src[m,n]  (bias gradients), deterministic, one launch, N <= 32768.  workspace >= 1024 +
 * 64*N floats; the first 1024 words are ticket counters: zero them once, every call leaves them zero again */
int ekaid_colsum(int is_bf16, const void* src, int64_t ld, int64_t M, int N, const float* rowscale, float* out,
                 float* workspace, void* stream);
/* up to 8 column sums over the same M rows in ONE launch.  is_bf16, src, ld, N, out are HOST arrays read at call time.
 * workspace >= count * (1024 + 64 * max N) floats; its first count*1024 words are ticket counters (zero them once). */
int ekaid_colsum_many(int count, const int32_t* is_bf16, const void* const* src, const int64_t* ld, int64_t M,
                      const int32_t* N, float* const* out, float* workspace, void* stream);
int ekaid_add_inplace(float* y, const float* x, int64_t n, void* stream);
/* flags[m] = (sum_c X[m,c] == 0): the mask of q_expand_v_cat (relation_encoder.py:23-27) */
int ekaid_row_zero_flags(const float* X, int64_t M, int D, uint8_t* flags, void* stream);
/* backward of the per-sample question broadcast: out[b,:] = sum over the rows of sample b (S stacked images x N
 * nodes) whose flag is clear */
int ekaid_group_rowsum(int is_bf16, const void* src, int64_t ld, int N, int B, int S, int D, const uint8_t* flags,
                       float* out, void* stream);
/* process_matrix / torch_broadcast_adj_matrix (utils/mimic_utils.py:119-149): float64 labels [B,S,S] ->
 * one-hot fp32 [B,N,N,L] in ONE launch (the reference loops over labels with a host sync each) */
int ekaid_onehot_adj(const double* labels, int B, int S, int N, int L, float* out, void* stream);
/* bbox_relation_type / reverse_type / get_adj_matrix ("feature extraction/ana_bbox_generator.py":266-302,320-335):
 * boxes f64 [B,N,4] (xmin,ymin,xmax,ymax) -> spatial labels f64 [B,S,S] in the HDF5 `image_adj_matrix` layout
 * (0 = far, 1 inside, 2 cover, 3 IoU >= 0.5, 4..11 = 45-degree sector; entry (j,i), j > i, = reverse_type of (i,j);
 * rows / columns >= N are 0).  lx, ly: image extent (1024 x 1024 in the reference); "far" = (lx+ly)/3 */
int ekaid_spatial_labels(const double* boxes, int B, int N, int S, double lx, double ly, double* labels, void* stream);
/* get_semantic_adj ("feature extraction/combine_dicts.py":106-151): detected classes int32 [B,T] (anatomy ids, then disease
 * ids offset by the anatomy count; ncls = background) -> int8 labels [B,S,S] (the HDF5 `semantic_adj_matrix` layout): 1 for an
 * anatomy / disease pair of the same organ group, raised to the disease co-occurrence value where both classes have one.
 * The reference's dictionaries arrive as per-class tables: group [ncls], in_ana / in_di [ncls] (0/1), small_idx [ncls]
 * (-1 = not a co-occurrence disease), small_adj [ns*ns]. */
int ekaid_semantic_labels(const int32_t* classes, int B, int T, int S, int ncls, const int32_t* group, const uint8_t* in_ana,
                          const uint8_t* in_di, const int32_t* small_idx, const int32_t* small_adj, int ns, int8_t* labels,
                          void* stream);

/* ---- relation-aware graph attention (models/graph_att.py:53-106, models/graph_att_layer.py:60-178) ------- */
/* cond[g,i,j] = sum_c adj[g,j,i,c], lbias[g,i,j] = sum_c adj[g,j,i,c] w[c]   (graph_att.py:76,88-92; Q2,Q5) */
int ekaid_adj_prep_fwd(const float* adj0, const float* adj1, int g_split, const float* w, int G, int N, int Kn, int L,
                       float* cond, float* lbias, void* stream);
/* dw_part[g,c] = sum_ij adj[g,j,i,c] * sum_p dlbias_part[p,g,i,j] */
int ekaid_adj_prep_bwd(const float* adj0, const float* adj1, int g_split, const float* dlbias_part, int nparts, int G,
                       int N, int Kn, int L, float* dw_part, void* stream);
/* The same on the loader's integer label matrices (int8 [*, S, S]; 0 = no edge, c+1 = plane c of process_matrix,
 * utils/mimic_utils.py:119-149): cond = (1 <= label <= L), lbias = w[label - 1]; the fp32 one-hot planes are never built.
 * This is what ChangeDetector.forward runs when it is handed label matrices instead of one-hot adjacency. */
int ekaid_adj_labels_fwd(const int8_t* lab0, const int8_t* lab1, int g_split, int S, const float* w, int G, int N, int Kn,
                         int L, float* cond, float* lbias, void* stream);
int ekaid_adj_labels_bwd(const int8_t* lab0, const int8_t* lab1, int g_split, int S, const float* dlbias_part, int nparts,
                         int G, int N, int Kn, int L, float* dw_part, void* stream);
/* process_matrix from int8 labels (for callers that want the reference's one-hot tensors) */
int ekaid_onehot_adj_i8(const int8_t* labels, int B, int S, int N, int L, float* out, void* stream);
/* gbias[g,i,j,h] = log(max(relu(Wp[h,:] . posemb(pair) + bp[h]), 1e-6)) straight from fp64 boxes [G,N,4]
 * (modules.py:162-166, utils/mimic_utils.py:152-208, graph_att_layer.py:113-135; Q7, Q13).
 * dim_t: 8 fp32 wave lengths 1000^(t/8) */
int ekaid_geom_bias_fwd(const double* bb0, const double* bb1, int g_split, const float* Wp, const float* bp,
                        const float* dim_t, int G, int N, int Kn, int H, float* gbias, const uint64_t* seed,
                        uint32_t site, float p, float* emb_cache, int fast_trig, void* stream);
/* part[r, h*65 + k], r < G * ekaid_geom_bias_bwd_parts(): partial sums, k < 64 -> dWp[h,k], k = 64 -> dbp[h]; the caller
 * adds the rows (ekaid_colsum) */
int ekaid_geom_bias_bwd_parts(void);
int ekaid_geom_bias_bwd(const double* bb0, const double* bb1, int g_split, const float* Wp, const float* bp,
                        const float* dim_t, int G, int N, int Kn, int H, const float* dgbias, float* part,
                        const uint64_t* seed, uint32_t site, float p, const float* emb_cache, int fast_trig,
                        void* stream);
/* emb_cache (optional, fp32 [G, N*Kn, 64]): the forward stores the (dropped) 64-d embedding of every pair so the
 * backward does not redo the 32 sincos per pair; fast_trig = 1 uses fp32 sincosf on the fp64 argument (bf16 path). */
/* QKZ [G*N, ld]: cols [0,D) query, [D,2D) key, [2D + h*D, 2D + (h+1)*D) Z_h.  P [G,N,H,Kn] fp32.
 * scores/sqrt(dh) (+gbias) -> where(cond>0, s, -9e15) + lbias -> softmax over keys
 * (graph_att_layer.py:105-157; Q6).  cond / lbias / gbias may be NULL. */
int ekaid_edge_softmax_fwd(int is_bf16, const void* QKZ, int64_t ld, int D, const float* cond, const float* lbias,
                           const float* gbias, int G, int N, int Kn, int H, float* P, void* Phl, void* stream);
/* Phl (optional, bf16 path only, needs H*Kn % 8 == 0): [2, G, N, H*Kn] bf16 = P split into hi and lo planes, which the
 * aggregation kernels stage with 16-byte async copies (pass the same pointer to edge_aggregate_fwd/bwd, or NULL).
 * is_bf16 = 3: Phl is [3, G, N, H*Kn] and plane 2 additionally receives P as IEEE fp16 (for the fp16 forward aggregation). */
/* out = sum_h P_h Z_h + b_out;  Xout = Xin + relu(2 out);  mask = (out > 0)
 * (graph_att_layer.py:164-176 re-associated per Q3, graph_att.py:95-104 (Q2), relation_encoder.py:81,129 (Q1)).
 * XoutT: optional copy of Xout in the operand type with pitch ldt (may be NULL).
 * Xin NULL: no residual.  mask NULL: plain attention output Xout = (Xin +) out, without doubling, dropout or ReLU
 * (GraphSelfAttentionLayer.forward stand-alone, graph_att_layer.py:164-178).
 * Z16 (optional): the Z blocks as IEEE fp16, [G*N, H*D] with pitch ldz16 -- the forward then multiplies fp16 P (plane 2 of
 * a 3-plane Phl) with fp16 Z (11 significant bits each, one MMA per product) and writes XoutT as fp16; QKZ is not read. */
int ekaid_edge_aggregate_fwd(int is_bf16, const float* P, const void* QKZ, int64_t ld, int D, const float* b_out,
                             const float* Xin, int G, int N, int Kn, int H, float* Xout, void* XoutT, int64_t ldt,
                             uint8_t* mask, const uint64_t* seed, uint32_t site, float p, const void* Phl,
                             const void* Z16, int64_t ldz16, void* stream);
int ekaid_edge_num_slices(int D);
/* number of dP partial slices ekaid_edge_aggregate_bwd will write for these arguments (1 when the per-image tensor-core
 * kernel applies: bf16, Phl given, N <= 64, D % 128 == 0; else ekaid_edge_num_slices(D)) */
int ekaid_edge_bwd_slices(int is_bf16, int D, int N, int Kn, int H, int have_phl);
/* dOut [G*N, D] = gscale*mask*dXout (gscale = 2/(1-p)); dQKZ[:, 2D:] = dZ; dPpart [slices, G,N,H,Kn] */
int ekaid_edge_aggregate_bwd(int is_bf16, const float* dXout, const uint8_t* mask, const float* P, const void* QKZ,
                             int64_t ld, int D, int G, int N, int Kn, int H, void* dQKZ, float* dOut, float* dPpart,
                             float gscale, const void* Phl, void* stream);
/* dQKZ[:, 0:2D] = (dQ, dK); dlbias_part [H, G,N,Kn] and dgbias [G,N,Kn,H] may be NULL */
int ekaid_edge_softmax_bwd(int is_bf16, const float* P, const float* dPpart, int nslices, const void* QKZ, int64_t ld,
                           int D, const float* cond, int G, int N, int Kn, int H, void* dQKZ, float* dlbias_part,
                           float* dgbias, void* stream);

/* ---- difference + gated fusion + attention pooling (modules.py:233-310) --------------------------------- */
/* X3 [2*BN, D] (main rows then reference rows).  mode 0 single graph, 1 'all' (Q1), 2 'i+s'.
 * Xc fp32 [2BN, D]; CAT [2BN, 3D] operand type: [:,0:D]=Xc, [:,D:2D]=Xaft-Xbef (modules.py:234-250,297-298) */
int ekaid_combine_diff_fwd(int is_bf16, const float* X3, int64_t BN, int D, int mode, float c1, float c2, float c3,
                           float* Xc, void* CAT, void* stream);
int ekaid_combine_diff_bwd(const float* dXc, const float* dCAT, int64_t BN, int D, int mode, float c1, float c2,
                           float c3, float* dX3, void* stream);
/* pre [M,2D] fp32 = (context | gate) pre-activations -> ctx=tanh, gate=sigmoid, CAT[:,2D:3D] = gate*ctx
 * (modules.py:278-288) */
int ekaid_gate_fwd(int is_bf16, const float* pre, int64_t M, int D, void* ctx, void* gate, void* CAT,
                   const uint64_t* seed, uint32_t site_ctx, uint32_t site_gate, float p, void* stream);
int ekaid_gate_bwd(int is_bf16, const float* dCAT, const void* ctx, const void* gate, int64_t M, int D, void* dpre,
                   const uint64_t* seed, uint32_t site_ctx, uint32_t site_gate, float p, void* stream);
/* att = sigmoid(E w + b) [M];  attended[g,:] = sum_n att[g,n] Xc[g,n,:]   (modules.py:302-308) */
int ekaid_att_pool_fwd(const float* E, int64_t M, int N, int D, int dim, const float* w, const float* b,
                       const float* Xc, float* att, float* attended, void* stream);
int ekaid_att_pool_bwd(int is_bf16, const float* dA, const float* dattw, const float* att, const float* Xc,
                       const float* E, const float* w, int64_t M, int N, int D, int dim, float* dXc, void* dE,
                       float* dpre, float escale, void* stream);

/* ---- question path (models/language_model.py) ------------------------------------------------------------ */
/* E[l*B+b,:] = [emb[q[b,l]] | emb_[q[b,l]]]   (:48-53) */
int ekaid_embed_gather(int is_bf16, const int64_t* q, const float* emb, const float* emb2, int B, int L, int ed,
                       void* E, void* stream);
/* demb[v,:] = sum of dE rows whose token is v; row padding_idx gets zeros (nn.Embedding(padding_idx=ntoken),
 * language_model.py:26: that row is never trained), -1 = no padding row.  Token ids must lie in [0, V). */
int ekaid_embed_gather_bwd(const int64_t* q, const float* dE, int64_t ldde, int B, int L, int ed, int V, float* demb,
                           int padding_idx, void* stream);
/* one GRU step (:106-115): gi, gh [B,3H] incl. biases; gates [B,4H] saves (r,z,n,gh_n).  gh_reset (optional, [3H]):
 * after use gh is overwritten with it, so the next step's split-K "gh += h W_hh^T" GEMM starts from the bias */
int ekaid_gru_cell_fwd(int is_bf16, const float* gi, float* gh, const float* hprev, int B, int H, float* h, void* hT,
                       float* gates, const float* gh_reset, void* stream);
int ekaid_gru_cell_bwd(int is_bf16, const float* dh, const float* gates, const float* hprev, int B, int H, float* dgi,
                       float* dgh, void* dgiT, void* dghT, float* dhprev, void* stream);
/* The whole recurrence of forward_all (:106-115) / its BPTT in ONE persistent launch (bf16 tensor-core path): CTA c owns
 * hidden units 8c..8c+7, keeps its slice of W_hh in shared memory, and a grid-wide barrier separates the time steps.
 * Rows are time-major (t*B + b).  gi [L*B,3H] = x W_ih^T + b_ih; Whh [3H,H] bf16; Hs [L*B,H] fp32; HsT [(L+1)*B,H] bf16
 * whose block 0 (h_{-1} = 0) the caller zeroes; gates [L,B,4H] saves (r,z,n,gh_n); dHs [L*B,H] = gradient reaching h_t from
 * outside the recurrence; dgi/dgh [L*B,3H] fp32 with bf16 copies dgiT/dghT.  barrier_ws: 8 bytes of device memory, ZERO
 * before the first launch that uses them; every launch leaves them zero again (the last CTA out clears both words), so no
 * memset sits in front of the kernel.  One workspace serves launches that are ordered on a stream, not concurrent ones.
 * Requires H % 512 == 0 and H/8 <= number of SMs (EKAID_ERR_UNSUPPORTED otherwise: use the per-step entry points).
 * fp16_ops = 1: Whh and HsT hold IEEE fp16 (the forward operand format of the 16-bit path; h is bounded by 1); HsB
 * (optional): bf16 copy of HsT for the backward wgrad, whose other operand is a bf16 gradient. */
int ekaid_gru_seq_fwd(const float* gi, const void* Whh, const float* bhh, int B, int H, int L, float* Hs, void* HsT,
                      float* gates, void* barrier_ws, int fp16_ops, void* HsB, void* stream);
int ekaid_gru_seq_bwd(const float* dHs, const float* gates, const float* Hs, const void* Whh, int B, int H, int L,
                      float* dgi, float* dgh, void* dgiT, void* dghT, void* barrier_ws, void* stream);
/* out[m] = A[m,:] . w + b[0]   (W2_self_att_q, :142) */
int ekaid_rowdot(int is_bf16, const void* A, int64_t lda, int64_t M, int K, const float* w, const float* b, float* out,
                 void* stream);
/* batch-axis softmax + reinterpret + weighted sum (:149-153, quirk Q4). a,S: [L*B]; Hs [L*B,H]; qv [B,H] */
int ekaid_qpool_fwd(const float* a, const float* Hs, int B, int L, int H, float* S, float* qv, void* stream);
int ekaid_qpool_bwd(const float* dqv, const float* S, const float* Hs, int B, int L, int H, float* dS, float* da,
                    float* dHs, void* stream);
/* dpre = da[r] w2[c] (1 - a1^2)  (backward of tanh(W1 h) . w2).  is_bf16: 0 = fp32 a1 and dpre, 1 = bf16 both, 3 = fp32
 * a1 with a bf16 dpre.  dpre32 (optional): unrounded fp32 copy for the bias gradient's cancelling column sum. */
int ekaid_qatt_tanh_bwd(int is_bf16, const float* da, const float* w2, const void* a1, int64_t M, int H, void* dpre,
                        float* dpre32, void* stream);

/* y[m,n] = x[m,:] . W[n,:] + b[n] for a handful of outputs (fc1, the 6-way change classifier, modules.py:312) */
int ekaid_small_linear(const float* x, int64_t ldx, int M, int K, const float* W, const float* b, int N, float* y,
                       void* stream);
/* out[0] = sum_k coef[k] * <a_k, w_k> (w_k NULL: plain sum), k < count <= 5; a, w, n, coef are HOST arrays.  The step
 * objective of train_mimic.py:246-247 when the decoder's gradient arrives as cotangents; deterministic.  workspace: 65
 * floats, word 0 a ticket counter (zero it once; every call leaves it zero). */
int ekaid_weighted_sums(int count, const float* const* a, const float* const* w, const int64_t* n, const float* coef,
                        float* out, float* workspace, void* stream);
/* backward of ekaid_weighted_sums in one launch: out_k[e] = g[0] * coef_k * (w_k ? w_k[e] : 1); g is a device scalar */
int ekaid_weighted_sums_bwd(int count, float* const* out, const float* const* w, const int64_t* n, const float* coef,
                            const float* g, void* stream);
/* input_attended = attended_2 - attended_1 (modules.py:309) from the stacked attended [2B, D]; BD = B*D */
int ekaid_head_fwd(const float* attended, int64_t BD, float* input_attended, void* stream);
/* gradients of the module's five outputs (att_bef, att_aft, attended_1, attended_2, input_attended; each may be NULL = zero)
 * -> d_att [2*B*N] and d_attended [2B, D] of the fusion stage, one launch; BN = B*N, BD = B*D */
int ekaid_head_bwd(const float* d_att_bef, const float* d_att_aft, const float* d_a1, const float* d_a2, const float* d_ia,
                   int64_t BN, int64_t BD, float* d_att, float* d_attended, void* stream);

/* ---- legacy weight_norm(dim=None) (models/fc.py:33-34): w = v * g / ||v||_F over the whole tensor -------------- */
/* workspace: 128 floats; norm_out: 1 float kept for the backward */
int ekaid_wn_fwd(const float* v, const float* g, int64_t n, float* w, float* norm_out, float* workspace, void* stream);
/* dv = (g/n) dw - (g <dw,v> / n^3) v ; dg = <dw,v> / n */
int ekaid_wn_bwd(const float* dw, const float* v, const float* g, const float* norm, int64_t n, float* dv, float* dg,
                 float* workspace, void* stream);

/* the same for `count` <= 16 tensors in two launches: v, g, w, dw, dv, dg are HOST arrays of `count` device pointers, n a
 * host array of element counts (all read at call time); norms [count] and workspace [count*128] are device memory */
int ekaid_wn_fwd_many(int count, const float* const* v, const float* const* g, const int64_t* n, float* const* w,
                      float* norms, float* workspace, void* stream);
int ekaid_wn_bwd_many(int count, const float* const* dw, const float* const* v, const float* const* g,
                      const float* norms, const int64_t* n, float* const* dv, float* const* dg, float* workspace,
                      void* stream);

/* ---- train-mode dropout helpers -------------------------------------------------------------------------- */
int ekaid_rng_advance(uint64_t* seed, void* stream);
/* VQ [M, D+Dq] = Dropout(cat(X, flag ? 0 : q))  -- the input of self_weights in train mode (relation_encoder.py:19-29,
 * fc.py:25-32; graph_att.py:80), operand type (is_bf16: 0 fp32, 1 bf16, 2 fp16).  VQB (optional): the same values as
 * bf16, the backward's copy (its wgrad pairs this operand with a bf16 gradient). */
int ekaid_build_vq(int is_bf16, const float* X, const float* qv, const uint8_t* flags, int64_t M, int N, int B, int D,
                   int Dq, void* VQ, const uint64_t* seed, uint32_t site, float p, void* VQB, void* stream);
/* out = sum_{k<nin} mult_k * in_k  (mult_k = dropout multiplier of site k, 1 when p_k = 0); index m*C + c.
 * outf (fp32, optional, accumulate != 0 adds to its old value) and/or outT (operand type, optional). */
int ekaid_drop_combine(int in_bf16, int out_bf16, int nin, const void* in0, const void* in1, const void* in2,
                       int64_t ldi, const uint64_t* seed, uint32_t site0, float p0, uint32_t site1, float p1,
                       uint32_t site2, float p2, int64_t M, int C, float* outf, int64_t ldf, int accumulate,
                       void* outT, int64_t ldo, void* stream);
/* out0 = mult(site0) * in, out1 = mult(site1) * in: two independent dropout sites on one tensor in the operand type (the
 * query and key inputs of a relation layer in train mode); index m*C + c.  out0B / out1B (optional, both or none): bf16
 * copies for the backward. */
int ekaid_drop_fanout(int is_bf16, const void* in, int64_t ldi, const uint64_t* seed, uint32_t site0, float p0,
                      uint32_t site1, float p1, int64_t M, int C, void* out0, void* out1, int64_t ldo, void* out0B,
                      void* out1B, void* stream);

/* test hook: out[e] = dropout multiplier (0 or 1/(1-p)) of element e < n at `site` for the current seed -- the same
 * function of (seed, site, index) every kernel uses, so a test can hand identical masks to the oracle */
int ekaid_drop_mask(const uint64_t* seed, uint32_t site, float p, int64_t n, float* out, void* stream);

/* ---- optimizer (utils/utils.py:96-99 -> torch.optim.Adam semantics) -------------------------------------- */
/* pow_state: device float[2] = {beta1^t, beta2^t}; call ekaid_adam_advance once per step before the updates */
int ekaid_adam_advance(float* pow_state, float b1, float b2, void* stream);
/* max_ctas > 0 caps the grid (an update running next to other kernels should not take every SM slot); 0 = full grid */
int ekaid_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                    float wd, const float* pow_state, int max_ctas, void* stream);

/* ---- answer decoder: DynamicSpeaker / DynamicCore (models/dynamic_speaker_change_pos.py:94-131, 182-240, 287-357) and
 * LanguageModelCriterion (utils/utils.py:204-216).  One decode step = dense products through ekaid_gemm_tc / ekaid_gemm_f32
 * plus these kernels.  opf: operand type of the GEMM-operand outputs (0 fp32, 1 bf16).  Dropout sites take
 * (seed, site, p) like the rest of the library; *_base = index of the first element at that site (steps are laid out one
 * after the other: (t * B + b) * width + column). ----------------------------------------------------------------- */
/* out[r, j] = Dropout(relu(emb[seq[b*sb + (t0+t)*st], j])), r = t*B + b, zero for We <= j < ldo (speaker.embed, :157-160).
 * err (optional) is set to 1 when a token id is outside [0, V). */
int ekaid_dec_embed(const int64_t* seq, int64_t sb, int64_t st, int t0, int B, int rows, const float* emb, int V, int We,
                    void* out, int64_t ldo, int opf, const uint64_t* seed, uint32_t site, float p, int32_t* err,
                    void* stream);
/* nn.LSTMCell point-wise part (:103, :123): pre = s0 + s1 + s2 (partial products, each optional) + tbl[tok] (optional table
 * row per token) + b1 + b2; gates [B,4R] (activated i,f,g,o), c_out, h_out fp32; h_op / out_op (optional) = h and
 * F.dropout(h) (:125) as operands. */
int ekaid_dec_lstm_fwd(const float* s0, int64_t l0, const float* s1, int64_t l1, const float* s2, int64_t l2,
                       const float* tbl, const int64_t* tok, const float* b1, const float* b2, const float* c_prev, int B,
                       int R, float* gates, float* c_out, float* h_out, void* h_op, int64_t ldh, void* out_op, int64_t ldo,
                       int opf, const uint64_t* seed, uint32_t site, float p, int64_t drop_base, void* stream);
/* dh = dh_a * mask + dh_b + dh_c (each optional); -> gate pre-activation gradients (operand + optional fp32) and dc_prev */
int ekaid_dec_lstm_bwd(const float* dh_a, int64_t lda, const uint64_t* seed, uint32_t site, float p, int64_t drop_base,
                       const float* dh_b, int64_t ldb, const float* dh_c, int64_t ldc, const float* dc_in,
                       const float* gates, const float* c, const float* c_prev, int B, int R, void* dpre_op, int64_t ldp,
                       int opf, float* dpre_f, float* dc_out, void* stream);
/* module attention + position branch of DynamicCore.forward (:104-119), one CTA per sample.
 * w7 = {weight_fc.W [3,R], weight_fc.b, pos1.b [P], weight_pos.W [16,P], weight_pos.b, pos2.W [R,16], pos2.b} (host array of
 * device pointers); p1pre = prev_h pos1.W^T.  Outputs: mw [B,4], pw [B,16], dposd [B,16] (= output_pos), vpos [B,P],
 * att [B,D], gi2 [B, R+D] = [ppos | att_feat] (operand). */
int ekaid_dec_att_fwd(const float* h_mod, const float* p1pre, int64_t ldp1, const float* const* w7, const float* bef,
                      const float* diff, const float* aft, int B, int R, int P, int D, const uint64_t* seed, uint32_t site1,
                      float p1, uint32_t site5, float p5, int64_t row_base, float* mw, float* pw, float* dposd, float* vpos,
                      float* att, void* gi2, int64_t ldg, int opf, void* stream);
/* backward of the above: dgi2 [B,R+D] fp32, datt_g [B,D]; dbef/ddiff/daft are accumulated (+=) */
int ekaid_dec_att_bwd(const float* dgi2, int64_t ldg, const float* datt_g, const float* const* w7, const float* bef,
                      const float* diff, const float* aft, const float* mw, const float* pw, const float* vpos, int B, int R,
                      int P, int D, const uint64_t* seed, uint32_t site1, float p1, uint32_t site5, float p5,
                      int64_t row_base, float* dbef, float* ddiff, float* daft, float* dfc, float* ddpos, float* dhmod_fc,
                      void* dvp_op, int64_t ldv, int opf, void* stream);
/* gate = sigmoid(pre); gated = gate * att (:121-123) and its backward */
int ekaid_dec_gate_fwd(const float* pre, const float* att, int64_t n, float* gate, void* gated, int opf, void* stream);
int ekaid_dec_gate_bwd(const float* dgated, const float* gate, const float* att, int64_t n, float* datt_g, void* dpre,
                       int opf, void* stream);
/* Dropout on an activation that feeds a GEMM; backward through Dropout(ReLU(.)) given the stored output y */
int ekaid_dec_drop_op(const float* x, int64_t ldx, int rows, int n, int xmod, const uint64_t* seed, uint32_t site, float p,
                      int64_t base, void* out, int64_t ldo, int opf, void* stream);
/* backward of F.log_softmax (:238) for callers that differentiate _forward's log-probabilities themselves */
int ekaid_dec_lsm_bwd(const float* dlogp, const float* logp, int rows, int B, int V, int Tout, void* dlogits, int64_t ldd,
                      int opf, void* stream);
int ekaid_dec_relu_drop_bwd(const float* dy, int64_t ldd, const void* y, int64_t ldy, int yf, int rows, int n, float keep,
                            void* out, int64_t ldo, int opf, float* out_f, int64_t ldf, void* stream);
/* acc[b, j] = sum_t x[(t*B+b), j] * mask(t, b, j): gradient of the step-invariant core.embed output */
int ekaid_dec_masked_sum_t(const float* x, int64_t ldx, int T, int B, int n, const uint64_t* seed, uint32_t site, float p,
                           int64_t base, float* acc, void* stream);
/* one sampling step on the device (:312-355): log-softmax, arg-max (multinomial = 0) or a draw from
 * exp(logp / temperature) (multinomial = 1; uniform from the counter RNG at (seed, site, b*(T+1)+t)), unfinished bookkeeping,
 * next token; state[0] = 1 while the reference's loop would still run (replaces the host synchronisation of :354) */
int ekaid_dec_token(const float* logits, int64_t ldl, int B, int V, int t, int T, int64_t* seq, float* seq_logp,
                    uint8_t* unfinished, int32_t* state, int64_t* next_tok, float* logp_out, int multinomial,
                    float temperature, const uint64_t* seed, uint32_t site, void* stream);
/* log-softmax + masked NLL + gradient over the logits of all steps (utils/utils.py:204-216; train_mimic.py:242):
 * mode bit 0: out[b,t,:] = logp; bit 1: row_loss; bit 2: dlogits (operand, pitch ldd) */
int ekaid_dec_nll(const float* logits, int64_t ldl, int rows, int B, int V, const int64_t* labels, int64_t lsb,
                  const float* masks, int64_t msb, int mode, float* out, int Tout, float* row_loss, const float* gscale,
                  const float* inv_msum, void* dlogits, int64_t ldd, int opf, int32_t* err, void* stream);
/* res[0] = sum(row_loss) / sum(mask), res[1] = 1 / sum(mask), mask = masks[:, 1 .. T] */
int ekaid_dec_nll_reduce(const float* row_loss, int rows, const float* masks, int64_t msb, int B, int T, float* res,
                         void* stream);
/* part[c][i, j] = sum over the rows of chunk c of a[r, i] b[r, j], i < m <= 16 (weight gradients of weight_fc, weight_pos,
 * pos2; part is [nchunks, m*n], stored [j, i] per chunk when transpose_out; ekaid_colsum over the chunks finishes it) */
int ekaid_dec_outer_small(const float* a, int64_t lda, int m, const float* b, int64_t ldb, int n, int rows, float* part,
                          int nchunks, int transpose_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EKAID_B200_H */
