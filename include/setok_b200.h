/* setok_b200.h — C ABI of the B200-native SeTok tokenizer hot path (libsetok_b200.so).
 *
 * Plain pointers and sizes only; no torch types.  Every pointer marked (device) is a CUDA device
 * pointer owned by the caller; buffers are row-major and contiguous unless a leading dimension is
 * passed.  No entry point allocates, synchronises or throws: each enqueues work on `stream` and
 * returns SETOK_OK (0) or a negative setok_status; setok_last_error() gives the message of the last
 * failure on the calling thread.  Distinct streams may be driven from distinct threads.
 *
 * The reference (ChocoWu/SeTok) is pure Python and has no FFI; the seam these entry points sit
 * behind is the duck-typed Python plugin API.  Each entry cites the reference
 * code it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a reference maintainer would add.
 */
#ifndef SETOK_B200_H_
#define SETOK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* setok_stream_t; /* cudaStream_t */

typedef enum {
  SETOK_OK = 0,
  SETOK_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, misaligned buffer            */
  SETOK_ERR_UNSUPPORTED = -2,  /* shape outside what the sm_100a kernels are instantiated for   */
  SETOK_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed (message has the reason)   */
  SETOK_ERR_WORKSPACE = -4     /* workspace smaller than the matching *_workspace_bytes() query  */
} setok_status;

typedef enum { SETOK_F32 = 0, SETOK_BF16 = 1, SETOK_U8 = 2 /* images only: setok_vit_forward_u8 */ } setok_dtype;
typedef enum { SETOK_ACT_NONE = 0, SETOK_ACT_QUICK_GELU = 1, SETOK_ACT_GELU_ERF = 2 } setok_act;

const char* setok_last_error(void);
int setok_abi_version(void);
/* Number of kernels this library has launched since load (process-wide, all streams). */
uint64_t setok_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Primitive: D[M,N] = epilogue(A[M,K] * W[N,K]^T)  — tcgen05/TMEM/TMA GEMM, bf16 in, fp32 accumulate.
 * Replaces every nn.Linear on the path (HF CLIP q/k/v/out_proj/fc1/fc2, src/model/setok/module.py:34-36,
 * :56-58, src/model/setok/tokenizer.py:43, src/model/multimodal_projector/builder.py:37,50-58).
 *   A, W        (device) bf16, leading dimensions lda/ldw in elements (multiples of 8, 16-byte aligned)
 *   D           (device) bf16 or f32 (out_dtype), leading dimension ldd
 *   bias        (device) f32 [N] or NULL
 *   residual    (device) bf16/f32 (residual_dtype) [M, ldr] or NULL; added after the activation
 *   m_dev       (device) optional int32: the live row count (rows >= *m_dev are skipped); M is then the
 *               capacity the buffers were sized for.  This is how ragged (data-dependent K) batches run
 *               without a host sync.
 *   N % 4 == 0; K arbitrary (leading dimensions must still be multiples of 8 elements).
 */
int setok_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd,
                    int out_dtype, const float* bias, const void* residual, int64_t ldr, int residual_dtype,
                    int act, int M, int N, int K, const int32_t* m_dev, setok_stream_t stream);

/* The LayerNorm-fold forms of setok_gemm_bf16 (what SETOK_VIT_LN_FOLD runs inside the tower, see the flag below): a pre-LN
 * transformer's LayerNorm (modeling_clip.py:374,381: layer_norm1 / layer_norm2) is never a pass of its own.
 *   records       (device) f32 [M, 2 + 2*ns], ns = ceil(ln_C / 128): per row {c, r, (s1, s2) x ns}
 *   producing side (ln_in, ln_out, xhat non-NULL; f32 residual stream, no activation, N == ln_C): besides
 *                 D = A W^T + bias + residual it writes xhat [M, ld_xhat] bf16 = (D - c') * r' with (c', r') = the mean and
 *                 inverse standard deviation of the row BEFORE this update (from ln_in), and into ln_out the pair (c', r')
 *                 plus, per 128-column slot, s1 = sum(D - c'), s2 = sum((D - c')^2)
 *   consuming side (ln_in, ln_s non-NULL; bf16 output, act none | quick_gelu): A = xhat, W = bf16(gamma (.) W0),
 *                 bias = W0 beta + b0, ln_s[n] = sum_k float(W[n, k]);  D = act(LayerNorm(x) W0^T + b0) with the row's exact
 *                 mean / variance taken from the record (m = sum s1 / C, var = sum s2 / C - m^2)
 * setok_ln_fold_init writes the first record and xhat of a stream x [rows, C] f32 (exact statistics, c = mean, r = rstd). */
int setok_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd, int out_dtype,
                       const float* bias, const void* residual, int64_t ldr, int residual_dtype, int act, int M, int N, int K,
                       const float* ln_in, float* ln_out, const float* ln_s, void* xhat, int64_t ld_xhat, float ln_eps,
                       int ln_C, setok_stream_t stream);
int setok_ln_fold_init(const float* x, void* xhat, float* records, float eps, int rows, int C, setok_stream_t stream);

/* Batched form: `batch` independent products D_b = epilogue(A_b * W_b^T); operand b starts at base + b * stride
 * (strides in elements).  With w_mn_major != 0, W_b is given as [K, N] row-major (N contiguous) instead of [N, K] —
 * the form P.V takes in attention (W = V: keys x head_dim).  Used for the dense masked attention of the cluster
 * encoder (src/model/setok/module.py:66-70 over all clusters of an image at once). */
int setok_gemm_bf16_batched(const void* A, int64_t lda, int64_t a_batch_stride, const void* W, int64_t ldw,
                            int64_t w_batch_stride, int w_mn_major, void* D, int64_t ldd, int64_t d_batch_stride,
                            int out_dtype, const float* bias, int act, int batch, int M, int N, int K,
                            setok_stream_t stream);

/* Row LayerNorm (eps inside the sqrt, biased variance — torch.nn.LayerNorm).  in/out dtype f32|bf16.
 * gather (device, int32 [rows]) optionally picks the source row: out[r] = LN(in[gather[r]]). */
int setok_layernorm(const void* in, int in_dtype, void* out, int out_dtype, const float* gamma, const float* beta,
                    float eps, int rows, int C, const int32_t* gather, const int32_t* m_dev, setok_stream_t stream);

/* Segmented (block-diagonal / varlen) multi-head self-attention over packed rows.
 * qkv bf16 [rows, 3C] laid out [q | k | v], heads of C/heads; row r attends to rows
 * [seg_off[s], seg_off[s+1]) where s = row_seg[r].  softmax(q k^T * scale) v, fp32 math.
 * Replaces Attention.forward (src/model/setok/module.py:61-73) for the per-cluster encoder
 * (tokenizer.py:150, segments = clusters) and the inter-cluster encoder (tokenizer.py:179, segments =
 * images), and the HF CLIP attention when uniform_T > 0 (segments = images of uniform_T rows,
 * seg_off/row_seg may then be NULL). */
int setok_attention(const void* qkv, void* out, int rows, int C, int heads, float scale,
                    const int32_t* seg_off, const int32_t* row_seg, int uniform_T, const int32_t* m_dev,
                    setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a1+a2: vision tower.  Replaces CLIPVisionTower.forward + feature_select
 * (src/model/setok/clip_encoder.py:50-62, :40-48) and the HF CLIPVisionTransformer it calls
 * (transformers/models/clip/modeling_clip.py:138-220, 261-386, 647-690).
 */
typedef struct {
  const void* w_qkv;   /* bf16 [3C, C]  rows = q_proj | k_proj | v_proj */
  const float* b_qkv;  /* f32 [3C] */
  const void* w_o;     /* bf16 [C, C] */
  const float* b_o;
  const void* w_fc1;   /* bf16 [F, C] */
  const float* b_fc1;
  const void* w_fc2;   /* bf16 [C, F] */
  const float* b_fc2;
  const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;
  /* SETOK_VIT_LN_FOLD only (else NULL): f32 [3C] / [F], s_n = sum_k W'_nk of the gamma-scaled bf16 matrices (see the flag) */
  const float* s_qkv; const float* s_fc1;
} setok_vit_layer;

typedef struct {
  int image_size, patch, hidden, heads, layers, mlp;
  float ln_eps;
  const void* w_patch;   /* bf16 [C, Kp], Kp = round_up(3*patch*patch, 64), columns (c, ky, kx), zero padded;
                            with SETOK_VIT_PATCH_SPLIT: bf16 [C, 3*Kp] = [w_hi | w_hi | w_lo], w = w_hi + w_lo */
  const float* cls;      /* f32 [C] */
  const float* pos;      /* f32 [(N+1), C] */
  const float* pre_ln_g; const float* pre_ln_b;
  const setok_vit_layer* layer;   /* host array of `layers` entries */
  int flags;             /* SETOK_VIT_* bits; 0 = bf16 residual stream */
} setok_vit;
/* Residual stream x of the tower kept in f32 between layers (LayerNorm reads f32, the out_proj / fc2 epilogues add and
 * store f32) instead of being rounded to bf16 after every residual add: 2 x layers fewer bf16 roundings of x. */
#define SETOK_VIT_RESIDUAL_F32 1
/* Patch embedding (modeling_clip.py:209) evaluated as a 3-term bf16 hi/lo split of pixels and weights in one GEMM of
 * K = 3*Kp (pixel.w = hi.w_hi + lo.w_hi + hi.w_lo): its rounding error, which the residual stream carries unchanged
 * through every layer, drops from 2^-9 to ~2^-17 relative for 0.4 % more tower FLOPs. */
#define SETOK_VIT_PATCH_SPLIT 2
/* The layers' LayerNorms folded into the GEMMs around them (needs SETOK_VIT_RESIDUAL_F32 and hidden % 32 == 0): no
 * normalisation pass re-reads the residual stream.  out_proj / fc2 also emit xhat = bf16((x - c) * r) -- x normalised with the
 * row statistics of one sub-layer earlier -- and per-row partial sums; qkv / fc1 multiply xhat by W' = gamma (.) W and finish
 * the LayerNorm exactly in their fp32 epilogue (gemm_tcgen05.cu).  With this flag the caller packs, per layer:
 *   w_qkv = bf16(W_qkv * ln1_g[None, :]), b_qkv = W_qkv ln1_b + b_qkv (f32), s_qkv[n] = sum_k float(w_qkv[n, k]);
 *   w_fc1 = bf16(W_fc1 * ln2_g[None, :]), b_fc1 = W_fc1 ln2_b + b_fc1,       s_fc1[n] = sum_k float(w_fc1[n, k]);
 * ln1_g .. ln2_b are not read. */
#define SETOK_VIT_LN_FOLD 4

size_t setok_vit_workspace_bytes(const setok_vit* vit, int B);
/* images (device) [B,3,H,W] f32|bf16 -> features (device) [B, N(+1), C] f32|bf16.
 * n_layers_run = index into HF hidden_states (select_layer resolved by the caller: -2 -> layers-1).
 * keep_cls: 0 = 'patch' (drop CLS), 1 = 'cls_patch'. */
int setok_vit_forward(const setok_vit* vit, const void* images, int image_dtype, int B, int n_layers_run,
                      int keep_cls, void* features, int feature_dtype, void* workspace, size_t workspace_bytes,
                      setok_stream_t stream);

/* SURVEY 8f row 3 (normalisation part): uint8 pixels straight into the patch embedding.  images (device) uint8 [B,3,H,W]
 * (channels first, already at the tower's resolution); the CLIPImageProcessor arithmetic of transformers 4.46.3 --
 * rescale: float32(float64(u8) * rescale_factor), normalize: (x - mean) / std in float32 -- is applied inside the im2col
 * pass, so the host->device copy carries 1 byte per pixel instead of 4.  lut[v] = float32(v * rescale_factor) is built by
 * the caller in double precision (256 entries).  pos_table may be NULL (then features = hidden_states[...][:, 1:] in
 * feature_dtype) or the (N, C) sincos table (then the output is the f32 position-embedded tensor, as setok_vit_forward_pos). */
typedef struct { float lut[256]; float mean[3]; float std[3]; } setok_u8_norm;
int setok_vit_forward_u8(const setok_vit* vit, const uint8_t* images, const setok_u8_norm* norm, int B, int n_layers_run,
                         int keep_cls, const float* pos_table, void* features, int feature_dtype, void* workspace,
                         size_t workspace_bytes, setok_stream_t stream);

/* Same tower, with feature_select('patch') and the position-embedding add of tokenizer.py:164-169 fused into the last
 * row pass: x_pos (device) f32 [B, N, C] = hidden_states[n_layers_run][:, 1:] + pos_table.  pos_table: (N, C) f32
 * (device), the PositionalEncoding2D table (module.py:118-146).  Feeds setok_dpc_cluster_embedded / setok_head_forward. */
int setok_vit_forward_pos(const setok_vit* vit, const void* images, int image_dtype, int B, int n_layers_run,
                          const float* pos_table, float* x_pos, void* workspace, size_t workspace_bytes,
                          setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a3+a4: 2-D sincos position embedding add + DPC-kNN clustering.  Replaces PositionalEncoding2D
 * (src/model/setok/module.py:105-146, utils.py:5-10), the add at tokenizer.py:164-169 and
 * SetokTokenizer.cluster_dpc_knn (tokenizer.py:78-121).
 *   feats       (device) [B, N, C] f32|bf16, N = h*w
 *   noise       (device) f32 [B, N]: the U[0,1) draw of tokenizer.py:91 (scaled by 1e-6 inside)
 *   token_mask  (device) f32 [B, N] or NULL (tokenizer.py:84-86, 93-94; >0 keeps the token)
 *   x_pos       (device) f32 [B, N, C] out: feats + pos  (the tensor tokenizer.py:168 produces)
 *   idx_cluster (device) int64 [B, N] out;  score f32 [B, N] out
 *   index_down  (device) int64 [B, N] out: first num_clusters[b] entries valid, rest -1
 *   num_clusters(device) int32 [B] out;  offsets int32 [B+1] out (exclusive scan of num_clusters)
 */
size_t setok_dpc_workspace_bytes(int B, int N, int C);
int setok_dpc_cluster(const void* feats, int feat_dtype, const float* noise, const float* token_mask,
                      int B, int h, int w, int C, int k, float threshold, int min_cluster_num,
                      float* x_pos, int64_t* idx_cluster, float* score, int64_t* index_down,
                      int32_t* num_clusters, int32_t* offsets, void* workspace, size_t workspace_bytes,
                      setok_stream_t stream);
/* Same, with the (h*w, C) f32 sincos table supplied by the caller (device).  The Python host passes the
 * table torch computes with the reference's own formula, so x_pos is bit-identical to the reference's;
 * setok_dpc_cluster() builds the table itself (libm, cached per (device, h, w, C)). */
int setok_dpc_cluster_pos(const void* feats, int feat_dtype, const float* pos_table, const float* noise,
                          const float* token_mask, int B, int h, int w, int C, int k, float threshold,
                          int min_cluster_num, float* x_pos, int64_t* idx_cluster, float* score,
                          int64_t* index_down, int32_t* num_clusters, int32_t* offsets, void* workspace,
                          size_t workspace_bytes, setok_stream_t stream);

/* cluster_dpc_knn (tokenizer.py:78-121) on features that already carry the position embedding (the x of
 * tokenizer.py:168, e.g. from setok_vit_forward_pos): x_pos (device) [B, N, C] f32|bf16 is read once, nothing but the
 * clustering outputs is written.  Workspace: setok_dpc_workspace_bytes(B, N, C). */
int setok_dpc_cluster_embedded(const void* x_pos, int dtype, const float* noise, const float* token_mask, int B, int N,
                               int C, int k, float threshold, int min_cluster_num, int64_t* idx_cluster, float* score,
                               int64_t* index_down, int32_t* num_clusters, int32_t* offsets, void* workspace,
                               size_t workspace_bytes, setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a5+a6: group_encoding (tokenizer.py:123-155) + inter_encoder + out (tokenizer.py:179-180) with
 * Block/Attention/Mlp (module.py:29-100).  Consumes the clustering outputs, emits the ragged batch
 * packed as tokens [sum K, C_tok] + offsets (from setok_dpc_cluster).
 */
typedef struct { const void* w_qkv; const float* b_qkv; const void* w_proj; const float* b_proj; } setok_attn;
typedef struct {
  int depth;                     /* number of attention layers sharing norm1 (module.py:86-91) */
  const float* n1_g; const float* n1_b; const float* n2_g; const float* n2_b;
  const setok_attn* attn;        /* host array [depth] */
  const void* w_fc1; const float* b_fc1;   /* bf16 [F, C] */
  const void* w_fc2; const float* b_fc2;   /* bf16 [C, F] */
} setok_block;

typedef struct {
  int hidden, heads, mlp, token_dim;
  setok_block inner, inter;
  const void* w_out; const float* b_out;   /* bf16 [C_tok, C] */
} setok_head;

size_t setok_head_workspace_bytes(const setok_head* head, int B, int N);
/* tokens (device) [B*N (capacity), C_tok] f32|bf16 out: rows [offsets[b], offsets[b+1]) are image b's tokens.
 * group_features (device) f32 [B*N, C] optional out (NULL to skip): the (K, C) tensor of tokenizer.py:153. */
int setok_head_forward(const setok_head* head, const float* x_pos, const int64_t* idx_cluster,
                       const int32_t* num_clusters, const int32_t* offsets, int B, int N,
                       void* tokens, int token_dtype, float* group_features,
                       void* workspace, size_t workspace_bytes, setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a8: mm_in_projector (src/model/multimodal_projector/builder.py:33-64) applied to the packed rows
 * (src/model/setokim_arch.py:210).  n_linear Linear layers with GELU(erf) between them; optional
 * LayerNorm after the first ('_Norm').
 */
typedef struct {
  int n_linear;
  const void* const* w;      /* host array of bf16 [out_i, in_i] */
  const float* const* b;     /* host array of f32 [out_i] */
  const int* dims;           /* host array [n_linear + 1]: in_0, out_0 (= in_1), ... */
  const float* norm_g; const float* norm_b;   /* NULL unless '_Norm' */
} setok_projector;

size_t setok_project_workspace_bytes(const setok_projector* proj, int rows);
int setok_project(const setok_projector* proj, const void* tokens, int token_dtype, int rows_capacity,
                  const int32_t* m_dev, void* out, int out_dtype, void* workspace, size_t workspace_bytes,
                  setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a10: SetokDeTokenizer.forward (src/model/setok/detokenizer.py:101-120) as a consumer of the ragged
 * token batch: mapper_fc_in -> Q-Former (module.py:476-583; learned queries cross-attend the image's K_b tokens every
 * cross_attention_freq layers, text FFN deleted detokenizer.py:94-96) -> decoder_fc_in -> + PositionalEncoding2D ->
 * decoder_depth timm ViT blocks -> decoder_norm.  The reference's padded (B, K_max, C_tok) + attention_masks pair is
 * replaced by packed rows + offsets (varlen cross-attention, no -10000 padding mask); the missing `return` is added.
 */
typedef struct {
  const void* w_qkv; const float* b_qkv;     /* attention.self.{query,key,value} fused: bf16 [3H, H] */
  const void* w_so; const float* b_so;       /* attention.output.dense bf16 [H, H] */
  const float* ln_s_g; const float* ln_s_b;  /* attention.output.LayerNorm */
  int has_cross;                             /* layer_num % cross_attention_freq == 0 (module.py:484-493) */
  const void* w_cq; const float* b_cq;       /* crossattention.self.query bf16 [H, H] */
  const void* w_ckv; const float* b_ckv;     /* crossattention.self.{key,value} fused: bf16 [2H, H] (encoder_width = H) */
  const void* w_co; const float* b_co;       /* crossattention.output.dense */
  const float* ln_c_g; const float* ln_c_b;  /* crossattention.output.LayerNorm */
  const void* w_f1; const float* b_f1;       /* intermediate_query.dense bf16 [I, H], GELU(erf) */
  const void* w_f2; const float* b_f2;       /* output_query.dense bf16 [H, I] */
  const float* ln_f_g; const float* ln_f_b;  /* output_query.LayerNorm */
} setok_qformer_layer;

typedef struct {
  int token_dim, hidden, q_heads, q_inter, q_layers, grid;      /* n_queries = grid * grid */
  int dec_dim, dec_heads, dec_mlp, dec_depth;
  float q_ln_eps, dec_ln_eps;                                   /* 1e-12 (BertConfig), 1e-5 (nn.LayerNorm) */
  const void* w_map_in; const float* b_map_in;                  /* mapper_fc_in bf16 [H, C_tok] */
  const float* mask_tokens;                                     /* f32 [Q, H] */
  const float* emb_ln_g; const float* emb_ln_b;                 /* mapper.embeddings.LayerNorm */
  const setok_qformer_layer* qlayer;                            /* host array [q_layers] */
  const void* w_dec_in; const float* b_dec_in;                  /* decoder_fc_in bf16 [Dd, H] */
  const float* pos;                                             /* f32 [Q, Dd]: PositionalEncoding2D(hidden)(grid, grid)[..., :Dd] */
  const setok_vit_layer* block;                                 /* host array [dec_depth]: timm Block (w_o = attn.proj) */
  const float* norm_g; const float* norm_b;                     /* decoder_norm */
} setok_detok;

size_t setok_detok_workspace_bytes(const setok_detok* detok, int B, int rows_capacity);
/* tokens (device) [rows_capacity, C_tok] f32|bf16, rows [offsets[b], offsets[b+1]) = image b's tokens; offsets (device)
 * int32 [B+1].  out (device) [B, grid*grid, dec_dim] f32|bf16. */
int setok_detok_forward(const setok_detok* detok, const void* tokens, int token_dtype, const int32_t* offsets, int B,
                        int rows_capacity, void* out, int out_dtype, void* workspace, size_t workspace_bytes,
                        setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a9 / SURVEY 8f row 2: the splice of SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal
 * (src/model/setokim_arch.py:241-354) on the device: every IMAGE_TOKEN_INDEX (-200) placeholder of a sample is replaced
 * by the K rows of the next image of the ragged batch, text ids go through the embedding table, labels are filled with
 * IGNORE_INDEX (-100) over image rows (TARGET_TOKEN_INDEX -300 -> -100, :345), sequences are truncated to max_length
 * (:307-310) and padded right or left to the batch maximum (:312-341).  A sample without a placeholder still consumes
 * one image index (:262-269).  Replaces the per-sample Python loop, the boolean-mask compaction and the torch.cat chain.
 *   input_ids      (device) int64 [B, L];  attention_mask (device) uint8 [B, L] or NULL (all valid);  labels int64 [B, L] or NULL
 *   embed          (device) [V, H] f32|bf16: the LLM's embed_tokens weight;  image_rows (device) [*, H], same dtype
 *   image_offsets  (device) int32 [n_images + 1]: rows [o[i], o[i+1]) are image i (RaggedTokens.offsets)
 *   out_cap        row capacity T of the outputs: embeds [B, T, H], labels_out int64 [B, T], mask_out uint8 [B, T],
 *                  pos_out int64 [B, T]; columns >= max_len are padding.  lens int32 [B] and max_len int32 [1] come back
 *                  on the device (no host sync inside); a sequence longer than out_cap is truncated to it.
 *   workspace      setok_splice_workspace_bytes(B, L, out_cap)
 */
size_t setok_splice_workspace_bytes(int B, int L, int out_cap);
/* Two-phase form for exactly-sized outputs: _plan computes lens / max_len (device) and leaves the plan in the workspace;
 * the caller reads max_len once, allocates [B, max_len, ...] outputs and calls _fill with out_cap = max_len and the SAME
 * workspace (sized for that out_cap).  out_cap_limit: truncate to this many columns at most (use a large value for none). */
int setok_splice_plan(const int64_t* input_ids, const uint8_t* attention_mask, int B, int L, const int32_t* image_offsets,
                      int n_images, int max_length, int out_cap_limit, int32_t* lens, int32_t* max_len, void* workspace,
                      size_t workspace_bytes, setok_stream_t stream);
int setok_splice_fill(const int64_t* input_ids, const uint8_t* attention_mask, const int64_t* labels, int B, int L,
                      const void* embed, int dtype, int V, int H, const void* image_rows, const int32_t* image_offsets,
                      int n_images, int pad_left, int out_cap, const int32_t* lens, const int32_t* max_len, void* embeds,
                      int64_t* labels_out, uint8_t* mask_out, int64_t* pos_out, void* workspace, size_t workspace_bytes,
                      setok_stream_t stream);
/* Both phases in one call, for a caller with a known column bound (no host read). */
int setok_splice(const int64_t* input_ids, const uint8_t* attention_mask, const int64_t* labels, int B, int L,
                 const void* embed, int dtype, int V, int H, const void* image_rows, const int32_t* image_offsets,
                 int n_images, int max_length, int pad_left, int out_cap, void* embeds, int64_t* labels_out,
                 uint8_t* mask_out, int64_t* pos_out, int32_t* lens, int32_t* max_len, void* workspace,
                 size_t workspace_bytes, setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * image preprocessing on the device.  Replaces, for a whole batch of decoded images of arbitrary sizes,
 * what the reference does per image on the host: process_images / expand2square (src/mm_utils.py:152-182) followed by
 * CLIPImageProcessor.preprocess of transformers 4.46.3 (PIL bicubic resize of the uint8 image to the shortest edge, center
 * crop); rescale + normalize then happen inside setok_vit_forward_u8.  Integer arithmetic (Pillow's 22-bit fixed-point
 * resampling, restated), bit-exact against PIL.
 *   One descriptor per image.  The host builds the tap tables exactly as Pillow's precompute_coeffs does (double precision,
 *   then fixed point): per axis, S rows of [first source index, tap count, taps[ksize]] for the S output pixels that survive
 *   the crop.  `canvas` = the image after the optional pad-to-square: the source pasted at (pad_x, pad_y) into an Hp x Wp
 *   background-coloured rectangle (never materialised).
 */
typedef struct {
  const uint8_t* src;          /* (device) H x W x 3 uint8, HWC, as decoded */
  int H, W;
  int pad_x, pad_y;            /* where the source sits in the canvas (0, 0 without padding) */
  int y0, y1;                  /* canvas rows [y0, y1) the vertical pass reads (the horizontal pass computes only those) */
  int top, left;               /* crop origin in the resized image (used by the identity passes) */
  int identity_x, identity_y;  /* the pass does not change the size: Pillow skips it */
  int ksize_x, ksize_y;        /* taps per table row */
  int kx_off, ky_off;          /* offsets (int32 words) of this image's horizontal / vertical tables in `tables` */
  long long tmp_off;           /* offset (bytes) of this image's [(y1 - y0), S, 3] intermediate in the workspace */
} setok_resize_desc;

/* descs (device) [B]; max_rows = max_b (y1 - y0); tables (device) int32; background: the pad colour (host, 3 bytes);
 * out (device) uint8 [B, 3, S, S]; workspace_needed = sum_b (y1 - y0) * S * 3 (the caller knows it from the descriptors). */
int setok_preprocess_u8(const setok_resize_desc* descs, int B, int max_rows, const int32_t* tables, int S,
                        const uint8_t* background, uint8_t* out, void* workspace, size_t workspace_bytes,
                        size_t workspace_needed, setok_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * primitives of the head's training path -- what torch.autograd does implicitly for the reference when
 * gradients flow through group_encoding / inter_encoder / out (src/model/setok/tokenizer.py:123-155, 179-180, Block /
 * Attention / Mlp module.py:29-100) and mm_projector (multimodal_projector/builder.py:33-64).  The contractions of the
 * backward pass run on setok_gemm_bf16_batched (dgrad: dX = dY W with W [N, K] as the MN-major operand; wgrad:
 * dW = dY^T X with X [M, K] as the MN-major operand and dY^T from setok_transpose_to_bf16); these are the row kernels in
 * between.  setok_b200/training.py composes them into torch.autograd Functions.
 */
/* out[b][c][r] = bf16(in[b][r][c]); in f32|bf16 with leading dimension ld_in, batch strides in elements */
int setok_transpose_to_bf16(const void* in, int in_dtype, int64_t ld_in, int64_t in_batch_stride, void* out, int64_t ld_out,
                            int64_t out_batch_stride, int rows, int cols, int batch, setok_stream_t stream);
/* out[c] += sum_r in[r][c]  (bias gradients; the caller zeroes `out`) */
int setok_colsum_add(const void* in, int in_dtype, int64_t ld, int rows, int cols, float* out, setok_stream_t stream);
/* LayerNorm backward from the saved input x (f32): dx (f32) written, dgamma / dbeta (f32 [C]) accumulated */
int setok_layernorm_bwd(const float* x, const void* dy, int dy_dtype, const float* gamma, float eps, int rows, int C,
                        float* dx, float* dgamma, float* dbeta, setok_stream_t stream);
/* GELU(erf): act = gelu(pre) as bf16; dpre = dy * gelu'(pre) */
int setok_gelu_fwd(const void* pre, int pre_dtype, void* act_bf16, int64_t n, setok_stream_t stream);
int setok_gelu_bwd(const void* pre, int pre_dtype, const void* dy, int dy_dtype, float* dpre, int64_t n, setok_stream_t stream);
/* Dense masked attention of the cluster encoders, middle step and its backward: rows are N-token images sorted by segment;
 * row r may attend columns [seg_off[s] - image start, seg_off[s+1] - image start), s = row_seg[r].
 *   P = softmax(scale * S) inside the segment, 0 outside (bf16);  dS = scale * P * (dP - sum_k P dP) (bf16) */
int setok_masked_softmax(const float* S, int64_t ldS, const int32_t* seg_off, const int32_t* row_seg, int rows, int N, float scale,
                         void* P_bf16, int64_t ldP, setok_stream_t stream);
int setok_masked_softmax_bwd(const void* P_bf16, int64_t ldP, const float* dP, int64_t lddP, const int32_t* seg_off,
                             const int32_t* row_seg, int rows, int N, float scale, void* dS_bf16, int64_t lddS,
                             setok_stream_t stream);
/* tokenizer.py:141-151 bookkeeping: stable sort of every image's tokens by cluster label (perm, row_seg, seg_off as in
 * setok_head_forward), the per-cluster mean and its backward */
int setok_sort_by_cluster(const int64_t* idx_cluster, const int32_t* num_clusters, const int32_t* offsets, int B, int N,
                          int32_t* perm, int32_t* row_seg, int32_t* seg_off, setok_stream_t stream);
int setok_segment_mean(const float* x, const int32_t* seg_off, int n_segments, int C, float* out, setok_stream_t stream);
int setok_segment_mean_bwd(const float* dg, const int32_t* seg_off, const int32_t* row_seg, int rows, int C, float* dx,
                           setok_stream_t stream);
/* out[r] = index[r] >= 0 ? in[index[r]] : 0   (f32 rows; pads / unpads ragged batches, applies permutations) */
int setok_gather_rows_f32(const float* in, const int32_t* index, int rows_out, int C, float* out, setok_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SETOK_B200_H_ */
