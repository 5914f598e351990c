// api.cu — host-side orchestration behind the C ABI: the vision tower, the clustering head encoders and
// the projector are each one call that enqueues its whole kernel sequence on the caller's stream.
// No allocation, no synchronisation: scratch comes from the caller's workspace, ragged row counts stay
// on the device (offsets[B] is read by the kernels themselves).
#include "common.cuh"
#include "rowops.cuh"

#include <cmath>

using namespace setok;

namespace {

struct VitBufs { bf16* A; float* emb; void* x; bf16 *h, *qkv, *ao, *u; float* rec[2]; };   // x: bf16, or f32 with SETOK_VIT_RESIDUAL_F32

// SETOK_VIT_LN_FOLD: per-row record of the folded LayerNorms {c, r, (s1, s2) x ceil(C / 128)} (gemm_tcgen05.cu)
inline size_t ln_rec_floats(int C) { return 2 + 2 * static_cast<size_t>(ceil_div(C, 128)); }

int vit_carve(const setok_vit* v, int B, Arena& a, VitBufs* o) {
  const int P = (v->image_size / v->patch) * (v->image_size / v->patch);
  const int T = P + 1, C = v->hidden;
  const int Kp = static_cast<int>(round_up(3 * v->patch * v->patch, 64));
  const size_t R = static_cast<size_t>(B) * T;
  o->A = a.take<bf16>(static_cast<size_t>(B) * P * Kp * ((v->flags & SETOK_VIT_PATCH_SPLIT) ? 3 : 1));
  o->emb = a.take<float>(R * C);
  if (v->flags & SETOK_VIT_RESIDUAL_F32) o->x = a.take<float>(R * C);
  else o->x = a.take<bf16>(R * C);
  o->h = a.take<bf16>(R * C);
  o->qkv = a.take<bf16>(R * 3 * C);
  o->ao = a.take<bf16>(R * C);
  o->u = a.take<bf16>(R * v->mlp);
  o->rec[0] = o->rec[1] = nullptr;
  if (v->flags & SETOK_VIT_LN_FOLD) {
    o->rec[0] = a.take<float>(R * ln_rec_floats(C));
    o->rec[1] = a.take<float>(R * ln_rec_floats(C));
  }
  return SETOK_OK;
}

int check_vit(const setok_vit* v) {
  SETOK_REQUIRE(v != nullptr, SETOK_ERR_BAD_ARG, "vit: null config");
  SETOK_REQUIRE(v->image_size > 0 && v->patch > 0 && v->image_size % v->patch == 0, SETOK_ERR_BAD_ARG,
                "vit: image_size %d not a multiple of patch %d", v->image_size, v->patch);
  SETOK_REQUIRE(v->hidden > 0 && v->heads > 0 && v->hidden % v->heads == 0 && v->hidden % 8 == 0, SETOK_ERR_UNSUPPORTED,
                "vit: hidden %d / heads %d unsupported", v->hidden, v->heads);
  SETOK_REQUIRE(v->mlp > 0 && v->mlp % 8 == 0 && v->layers >= 0, SETOK_ERR_UNSUPPORTED, "vit: mlp %d layers %d unsupported", v->mlp, v->layers);
  SETOK_REQUIRE(v->w_patch && v->cls && v->pos && v->pre_ln_g && v->pre_ln_b && (v->layers == 0 || v->layer), SETOK_ERR_BAD_ARG, "vit: null weights");
  SETOK_REQUIRE(!(v->flags & SETOK_VIT_LN_FOLD) || ((v->flags & SETOK_VIT_RESIDUAL_F32) && v->hidden % 32 == 0 && v->mlp % 32 == 0), SETOK_ERR_UNSUPPORTED,
                "vit: SETOK_VIT_LN_FOLD needs SETOK_VIT_RESIDUAL_F32 and hidden / mlp multiples of 32 (hidden %d, mlp %d)", v->hidden, v->mlp);
  return SETOK_OK;
}

// One pre-LN transformer layer on the residual stream x [R, C] (bf16 or f32: xdt) of sequences of T rows:
//   x += W_o MHSA(LN1(x)); x += W_2 act(W_1 LN2(x))      (CLIPEncoderLayer, modeling_clip.py:363-386; timm Block)
int run_preln_layer(const setok_vit_layer& L, void* x, int xdt, bf16* h, bf16* qkv, bf16* ao, bf16* u, int R, int T, int C, int F, int heads,
                    float eps, int act, cudaStream_t stream) {
  const float scale = 1.0f / std::sqrt(static_cast<float>(C / heads));
  SETOK_TRY(launch_layernorm(x, xdt, h, SETOK_BF16, L.ln1_g, L.ln1_b, eps, R, C, nullptr, nullptr, stream));
  SETOK_TRY(launch_gemm(GemmArgs{h, C, L.w_qkv, C, qkv, 3LL * C, SETOK_BF16, L.b_qkv, nullptr, 0, 0, SETOK_ACT_NONE, R, 3 * C, C, nullptr, 0}, stream));
  SETOK_TRY(launch_attention(qkv, ao, R, C, heads, scale, nullptr, nullptr, T, nullptr, stream));
  SETOK_TRY(launch_gemm(GemmArgs{ao, C, L.w_o, C, x, C, xdt, L.b_o, x, C, xdt, SETOK_ACT_NONE, R, C, C, nullptr, 0}, stream));
  SETOK_TRY(launch_layernorm(x, xdt, h, SETOK_BF16, L.ln2_g, L.ln2_b, eps, R, C, nullptr, nullptr, stream));
  SETOK_TRY(launch_gemm(GemmArgs{h, C, L.w_fc1, C, u, F, SETOK_BF16, L.b_fc1, nullptr, 0, 0, act, R, F, C, nullptr, 0}, stream));
  SETOK_TRY(launch_gemm(GemmArgs{u, F, L.w_fc2, F, x, C, xdt, L.b_fc2, x, C, xdt, SETOK_ACT_NONE, R, C, F, nullptr, 0}, stream));
  return SETOK_OK;
}

// The same layer with its two LayerNorms folded into the GEMMs (SETOK_VIT_LN_FOLD): `h` holds xhat of the current stream and
// rec[0] its row records on entry and on exit; out_proj writes the stream, xhat and rec[1], fc2 the stream, xhat and rec[0].
// emit_next = false (the last layer run): fc2 only writes the stream.
int run_preln_layer_fold(const setok_vit_layer& L, float* x, bf16* h, bf16* qkv, bf16* ao, bf16* u, float* const rec[2], int R, int T, int C, int F,
                         int heads, float eps, int act, bool emit_next, cudaStream_t stream) {
  const float scale = 1.0f / std::sqrt(static_cast<float>(C / heads));
  GemmArgs g1{h, C, L.w_qkv, C, qkv, 3LL * C, SETOK_BF16, L.b_qkv, nullptr, 0, 0, SETOK_ACT_NONE, R, 3 * C, C, nullptr, 0};
  g1.ln_in = rec[0]; g1.ln_s = L.s_qkv; g1.ln_eps = eps; g1.ln_C = C;
  SETOK_TRY(launch_gemm(g1, stream));
  SETOK_TRY(launch_attention(qkv, ao, R, C, heads, scale, nullptr, nullptr, T, nullptr, stream));
  GemmArgs g2{ao, C, L.w_o, C, x, C, SETOK_F32, L.b_o, x, C, SETOK_F32, SETOK_ACT_NONE, R, C, C, nullptr, 0};
  g2.ln_in = rec[0]; g2.ln_out = rec[1]; g2.xhat = h; g2.ld_xhat = C; g2.ln_eps = eps; g2.ln_C = C;
  SETOK_TRY(launch_gemm(g2, stream));
  GemmArgs g3{h, C, L.w_fc1, C, u, F, SETOK_BF16, L.b_fc1, nullptr, 0, 0, act, R, F, C, nullptr, 0};
  g3.ln_in = rec[1]; g3.ln_s = L.s_fc1; g3.ln_eps = eps; g3.ln_C = C;
  SETOK_TRY(launch_gemm(g3, stream));
  GemmArgs g4{u, F, L.w_fc2, F, x, C, SETOK_F32, L.b_fc2, x, C, SETOK_F32, SETOK_ACT_NONE, R, C, F, nullptr, 0};
  if (emit_next) { g4.ln_in = rec[1]; g4.ln_out = rec[0]; g4.xhat = h; g4.ld_xhat = C; g4.ln_eps = eps; g4.ln_C = C; }
  SETOK_TRY(launch_gemm(g4, stream));
  return SETOK_OK;
}

struct BlockBufs { bf16 *h, *qkv, *ao, *u; float* S; bf16* P; };

// Block.forward (reference module.py:95-100): depth x [x += Attn_i(norm1(x))], then x += Mlp(norm2(x)).
// x is the fp32 residual stream over packed rows; attention is restricted to each row's segment.
// dense_N > 0: the rows are dense_N-token images sorted by cluster (the per-cluster encoder).  Attention then runs as a
// dense masked attention per image on the tensor cores -- S = Q K^T (batched GEMM), masked softmax over each row's
// cluster range, O = P V (batched GEMM, V as MN-major operand) -- whose cost does not depend on the cluster sizes;
// the warp-per-row kernel is kept for the ragged inter-cluster encoder, where segments are a handful of tokens.
int run_block(const setok_block& blk, int C, int heads, int F, float* x, int rows_cap, const int32_t* m_dev,
              const int32_t* seg_off, const int32_t* row_seg, const BlockBufs& w, cudaStream_t stream, int dense_N = 0) {
  const float scale = 1.0f / std::sqrt(static_cast<float>(C / heads));
  const int hd = C / heads;
  for (int i = 0; i < blk.depth; ++i) {
    const setok_attn& at = blk.attn[i];
    SETOK_TRY(launch_layernorm(x, SETOK_F32, w.h, SETOK_BF16, blk.n1_g, blk.n1_b, 1e-5f, rows_cap, C, nullptr, m_dev, stream));
    SETOK_TRY(launch_gemm(GemmArgs{w.h, C, at.w_qkv, C, w.qkv, 3LL * C, SETOK_BF16, at.b_qkv, nullptr, 0, 0, SETOK_ACT_NONE, rows_cap, 3 * C, C, m_dev, 0}, stream));
    if (dense_N > 0) {
      const int N = dense_N, Bimg = rows_cap / N;
      const int ldS = N, ldP = static_cast<int>(round_up(N, 8));
      for (int h = 0; h < heads; ++h) {
        GemmArgs sg{w.qkv + h * hd, 3LL * C, w.qkv + C + h * hd, 3LL * C, w.S, ldS, SETOK_F32, nullptr, nullptr, 0, 0, SETOK_ACT_NONE, N, N, hd, nullptr, 0};
        sg.batch = Bimg; sg.a_batch_stride = 3LL * C * N; sg.w_batch_stride = 3LL * C * N; sg.d_batch_stride = static_cast<int64_t>(N) * ldS;
        SETOK_TRY(launch_gemm(sg, stream));
        SETOK_TRY(launch_masked_softmax(w.S, w.P, seg_off, row_seg, rows_cap, N, ldS, ldP, scale, stream));
        GemmArgs pv{w.P, ldP, w.qkv + 2 * C + h * hd, 3LL * C, w.ao + h * hd, C, SETOK_BF16, nullptr, nullptr, 0, 0, SETOK_ACT_NONE, N, hd, N, nullptr, 0};
        pv.batch = Bimg; pv.a_batch_stride = static_cast<int64_t>(N) * ldP; pv.w_batch_stride = 3LL * C * N; pv.d_batch_stride = static_cast<int64_t>(N) * C;
        pv.w_mn_major = 1;
        SETOK_TRY(launch_gemm(pv, stream));
      }
    } else {
      SETOK_TRY(launch_attention(w.qkv, w.ao, rows_cap, C, heads, scale, seg_off, row_seg, 0, m_dev, stream));
    }
    SETOK_TRY(launch_gemm(GemmArgs{w.ao, C, at.w_proj, C, x, C, SETOK_F32, at.b_proj, x, C, SETOK_F32, SETOK_ACT_NONE, rows_cap, C, C, m_dev, 0}, stream));
  }
  SETOK_TRY(launch_layernorm(x, SETOK_F32, w.h, SETOK_BF16, blk.n2_g, blk.n2_b, 1e-5f, rows_cap, C, nullptr, m_dev, stream));
  SETOK_TRY(launch_gemm(GemmArgs{w.h, C, blk.w_fc1, C, w.u, F, SETOK_BF16, blk.b_fc1, nullptr, 0, 0, SETOK_ACT_GELU_ERF, rows_cap, F, C, m_dev, 0}, stream));
  SETOK_TRY(launch_gemm(GemmArgs{w.u, F, blk.w_fc2, F, x, C, SETOK_F32, blk.b_fc2, x, C, SETOK_F32, SETOK_ACT_NONE, rows_cap, C, F, m_dev, 0}, stream));
  return SETOK_OK;
}

int check_block(const setok_block& b, const char* name) {
  SETOK_REQUIRE(b.depth >= 0 && (b.depth == 0 || b.attn) && b.n1_g && b.n1_b && b.n2_g && b.n2_b && b.w_fc1 && b.b_fc1 && b.w_fc2 && b.b_fc2,
                SETOK_ERR_BAD_ARG, "head: null weights in %s block", name);
  for (int i = 0; i < b.depth; ++i)
    SETOK_REQUIRE(b.attn[i].w_qkv && b.attn[i].b_qkv && b.attn[i].w_proj && b.attn[i].b_proj, SETOK_ERR_BAD_ARG, "head: null attention weights in %s block layer %d", name, i);
  return SETOK_OK;
}

struct HeadBufs { int32_t *perm, *row_seg, *seg_off, *img_seg; float *xs, *g; bf16* gb; BlockBufs bb; };

void head_carve(const setok_head* hd, int B, int N, Arena& a, HeadBufs* o) {
  const size_t R = static_cast<size_t>(B) * N;
  const int C = hd->hidden;
  o->perm = a.take<int32_t>(R);
  o->row_seg = a.take<int32_t>(R);
  o->seg_off = a.take<int32_t>(R + 1);
  o->img_seg = a.take<int32_t>(R);
  o->xs = a.take<float>(R * C);
  o->g = a.take<float>(R * C);
  o->gb = a.take<bf16>(R * C);
  o->bb.h = a.take<bf16>(R * C);
  o->bb.qkv = a.take<bf16>(R * 3 * C);
  o->bb.ao = a.take<bf16>(R * C);
  o->bb.u = a.take<bf16>(R * hd->mlp);
  o->bb.S = a.take<float>(R * static_cast<size_t>(N));
  o->bb.P = a.take<bf16>(R * round_up(N, 8));
}

}  // namespace

// =================================================================================================
extern "C" size_t setok_vit_workspace_bytes(const setok_vit* vit, int B) {
  if (vit == nullptr || B <= 0 || vit->patch <= 0) return 0;
  Arena a(nullptr, 0);
  VitBufs b;
  vit_carve(vit, B, a, &b);
  return a.off;
}

namespace {
int vit_forward_impl(const setok_vit* v, const void* images, int image_dtype, int B, int n_layers_run, int keep_cls, void* features,
                     int feature_dtype, const float* pos_add, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     const setok_u8_norm* u8norm = nullptr) {
  SETOK_NVTX("setok a1+a2(+a3) vision tower");
  SETOK_TRY(check_vit(v));
  SETOK_REQUIRE(images && features && B > 0, SETOK_ERR_BAD_ARG, "vit_forward: null images/features or B <= 0");
  SETOK_REQUIRE(n_layers_run >= 0 && n_layers_run <= v->layers, SETOK_ERR_BAD_ARG, "vit_forward: n_layers_run %d outside [0, %d]", n_layers_run, v->layers);
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_vit_workspace_bytes(v, B), SETOK_ERR_WORKSPACE, "vit_forward: workspace too small");
  const int G = v->image_size / v->patch, P = G * G, T = P + 1, C = v->hidden, F = v->mlp;
  const int Kp = static_cast<int>(round_up(3 * v->patch * v->patch, 64));
  const int R = B * T;
  const int split = (v->flags & SETOK_VIT_PATCH_SPLIT) ? 1 : 0;
  Arena a(workspace, workspace_bytes);
  VitBufs w;
  vit_carve(v, B, a, &w);

  // embeddings (modeling_clip.py:199-220): conv patch-embed as im2col + GEMM whose epilogue scatters the
  // rows past each image's CLS slot and adds the position embedding; CLS rows; pre_layrnorm.
  if (image_dtype == SETOK_U8) {
    SETOK_REQUIRE(u8norm != nullptr, SETOK_ERR_BAD_ARG, "vit_forward: uint8 images need the normalisation constants (setok_vit_forward_u8)");
    SETOK_TRY(launch_im2col_u8(static_cast<const uint8_t*>(images), u8norm, w.A, B, v->image_size, v->image_size, v->patch, Kp, split, stream));
  } else {
    SETOK_TRY(launch_im2col(images, image_dtype, w.A, B, v->image_size, v->image_size, v->patch, Kp, split, stream));
  }
  const int Ke = split ? 3 * Kp : Kp;     // split: rows [hi | lo | hi] against weight rows [w_hi | w_hi | w_lo]
  SETOK_TRY(launch_gemm(GemmArgs{w.A, Ke, v->w_patch, Ke, w.emb, C, SETOK_F32, nullptr, v->pos, C, SETOK_F32, SETOK_ACT_NONE, B * P, C, Ke, nullptr, P}, stream));
  SETOK_TRY(launch_cls_rows(w.emb, v->cls, v->pos, B, T, C, stream));
  const int xdt = (v->flags & SETOK_VIT_RESIDUAL_F32) ? SETOK_F32 : SETOK_BF16;
  const bool fold = (v->flags & SETOK_VIT_LN_FOLD) != 0;
  if (fold && n_layers_run > 0)   // pre_layrnorm, and the first xhat / row records of the stream it starts, in one pass
    SETOK_TRY(launch_preln_fold_init(w.emb, static_cast<float*>(w.x), w.h, w.rec[0], v->pre_ln_g, v->pre_ln_b, v->ln_eps, R, C, stream));
  else
    SETOK_TRY(launch_layernorm(w.emb, SETOK_F32, w.x, xdt, v->pre_ln_g, v->pre_ln_b, v->ln_eps, R, C, nullptr, nullptr, stream));
  for (int l = 0; l < n_layers_run; ++l) {
    const setok_vit_layer& L = v->layer[l];
    SETOK_REQUIRE(L.w_qkv && L.b_qkv && L.w_o && L.b_o && L.w_fc1 && L.b_fc1 && L.w_fc2 && L.b_fc2, SETOK_ERR_BAD_ARG, "vit_forward: null weights in layer %d", l);
    if (fold) {
      SETOK_REQUIRE(L.s_qkv && L.s_fc1, SETOK_ERR_BAD_ARG, "vit_forward: SETOK_VIT_LN_FOLD without s_qkv / s_fc1 in layer %d", l);
      SETOK_TRY(run_preln_layer_fold(L, static_cast<float*>(w.x), w.h, w.qkv, w.ao, w.u, w.rec, R, T, C, F, v->heads, v->ln_eps, SETOK_ACT_QUICK_GELU,
                                     l + 1 < n_layers_run, stream));
    } else {
      SETOK_REQUIRE(L.ln1_g && L.ln1_b && L.ln2_g && L.ln2_b, SETOK_ERR_BAD_ARG, "vit_forward: null LayerNorm weights in layer %d", l);
      SETOK_TRY(run_preln_layer(L, w.x, xdt, w.h, w.qkv, w.ao, w.u, R, T, C, F, v->heads, v->ln_eps, SETOK_ACT_QUICK_GELU, stream));
    }
  }
  // feature_select (clip_encoder.py:40-48)
  SETOK_TRY(launch_select_rows(w.x, xdt, features, feature_dtype, B, T, keep_cls ? 0 : 1, C, pos_add, stream));
  return SETOK_OK;
}
}  // namespace

extern "C" int setok_vit_forward(const setok_vit* v, const void* images, int image_dtype, int B, int n_layers_run, int keep_cls,
                                 void* features, int feature_dtype, void* workspace, size_t workspace_bytes, setok_stream_t stream) {
  return vit_forward_impl(v, images, image_dtype, B, n_layers_run, keep_cls, features, feature_dtype, nullptr, workspace, workspace_bytes,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int setok_vit_forward_u8(const setok_vit* v, const uint8_t* images, const setok_u8_norm* norm, int B, int n_layers_run, int keep_cls,
                                    const float* pos_table, void* features, int feature_dtype, void* workspace, size_t workspace_bytes,
                                    setok_stream_t stream) {
  SETOK_REQUIRE(norm != nullptr, SETOK_ERR_BAD_ARG, "vit_forward_u8: null normalisation constants");
  for (int c = 0; c < 3; ++c) SETOK_REQUIRE(norm->std[c] != 0.f, SETOK_ERR_BAD_ARG, "vit_forward_u8: std[%d] is zero", c);
  SETOK_REQUIRE(pos_table == nullptr || (keep_cls == 0 && feature_dtype == SETOK_F32 && aligned16(pos_table)), SETOK_ERR_BAD_ARG,
                "vit_forward_u8: the fused position-embedding add needs 'patch' features in f32 and a 16-byte aligned table");
  return vit_forward_impl(v, images, SETOK_U8, B, n_layers_run, keep_cls, features, feature_dtype, pos_table, workspace, workspace_bytes,
                          static_cast<cudaStream_t>(stream), norm);
}

extern "C" int setok_vit_forward_pos(const setok_vit* v, const void* images, int image_dtype, int B, int n_layers_run, const float* pos_table,
                                     float* x_pos, void* workspace, size_t workspace_bytes, setok_stream_t stream) {
  SETOK_REQUIRE(pos_table != nullptr && aligned16(pos_table), SETOK_ERR_BAD_ARG, "vit_forward_pos: pos_table must be a 16-byte aligned (N, C) f32 table");
  return vit_forward_impl(v, images, image_dtype, B, n_layers_run, 0, x_pos, SETOK_F32, pos_table, workspace, workspace_bytes,
                          static_cast<cudaStream_t>(stream));
}

// =================================================================================================
extern "C" size_t setok_head_workspace_bytes(const setok_head* head, int B, int N) {
  if (head == nullptr || B <= 0 || N <= 0) return 0;
  Arena a(nullptr, 0);
  HeadBufs b;
  head_carve(head, B, N, a, &b);
  return a.off;
}

extern "C" int setok_head_forward(const setok_head* hd, const float* x_pos, const int64_t* idx_cluster, const int32_t* num_clusters,
                                  const int32_t* offsets, int B, int N, void* tokens, int token_dtype, float* group_features,
                                  void* workspace, size_t workspace_bytes, setok_stream_t stream_) {
  SETOK_NVTX("setok a5+a6 cluster encoders");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(hd && x_pos && idx_cluster && num_clusters && offsets && tokens, SETOK_ERR_BAD_ARG, "head_forward: null pointer");
  SETOK_REQUIRE(B > 0 && N > 0, SETOK_ERR_BAD_ARG, "head_forward: B=%d N=%d", B, N);
  const int C = hd->hidden, F = hd->mlp, Ct = hd->token_dim;
  SETOK_REQUIRE(C > 0 && hd->heads > 0 && C % hd->heads == 0 && C % 8 == 0 && F % 8 == 0 && Ct % 8 == 0, SETOK_ERR_UNSUPPORTED,
                "head_forward: hidden %d heads %d mlp %d token_dim %d unsupported (multiples of 8 required)", C, hd->heads, F, Ct);
  SETOK_REQUIRE((C / hd->heads) % 8 == 0 && C / hd->heads <= 512, SETOK_ERR_UNSUPPORTED, "head_forward: head_dim %d unsupported", C / hd->heads);
  SETOK_TRY(check_block(hd->inner, "inner"));
  SETOK_TRY(check_block(hd->inter, "inter"));
  SETOK_REQUIRE(hd->w_out && hd->b_out, SETOK_ERR_BAD_ARG, "head_forward: null out projection");
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_head_workspace_bytes(hd, B, N), SETOK_ERR_WORKSPACE, "head_forward: workspace too small");
  const int R = B * N;
  Arena a(workspace, workspace_bytes);
  HeadBufs w;
  head_carve(hd, B, N, a, &w);
  const int32_t* n_tok = offsets + B;   // device scalar: sum_b K_b

  // group_encoding (tokenizer.py:123-155): tokens sorted by cluster so that every cluster is a contiguous
  // segment; the Block's linear layers then run over all B*N rows at once and only attention is segmented.
  SETOK_TRY(launch_sort_by_cluster(idx_cluster, num_clusters, offsets, B, N, w.perm, w.row_seg, w.seg_off, stream));
  SETOK_TRY(launch_gather_rows(x_pos, w.xs, w.perm, R, C, stream));
  const int dense_N = (N % 4 == 0 && N >= 64 && (C / hd->heads) % 8 == 0) ? N : 0;
  SETOK_TRY(run_block(hd->inner, C, hd->heads, F, w.xs, R, nullptr, w.seg_off, w.row_seg, w.bb, stream, dense_N));
  SETOK_TRY(launch_segment_mean(w.xs, w.seg_off, n_tok, R, C, w.g, group_features, stream));       // tokenizer.py:151

  // inter_encoder over each image's K_b cluster tokens (tokenizer.py:179, repair R2), then `out` (:180)
  SETOK_TRY(launch_image_segments(offsets, B, w.img_seg, stream));
  SETOK_TRY(run_block(hd->inter, C, hd->heads, F, w.g, R, n_tok, offsets, w.img_seg, w.bb, stream));
  SETOK_TRY(launch_convert_rows(w.g, SETOK_F32, w.gb, SETOK_BF16, R, C, SETOK_ACT_NONE, n_tok, stream));
  SETOK_TRY(launch_gemm(GemmArgs{w.gb, C, hd->w_out, C, tokens, Ct, token_dtype, hd->b_out, nullptr, 0, 0, SETOK_ACT_NONE, R, Ct, C, n_tok, 0}, stream));
  return SETOK_OK;
}

// =================================================================================================
extern "C" size_t setok_project_workspace_bytes(const setok_projector* p, int rows) {
  if (p == nullptr || rows <= 0 || p->n_linear <= 0) return 0;
  Arena a(nullptr, 0);
  int maxd = 0;
  for (int i = 0; i <= p->n_linear; ++i) maxd = p->dims[i] > maxd ? p->dims[i] : maxd;
  a.take<bf16>(static_cast<size_t>(rows) * maxd);
  a.take<bf16>(static_cast<size_t>(rows) * maxd);
  a.take<float>(static_cast<size_t>(rows) * maxd);
  return a.off;
}

extern "C" int setok_project(const setok_projector* p, const void* tokens, int token_dtype, int rows, const int32_t* m_dev,
                             void* out, int out_dtype, void* workspace, size_t workspace_bytes, setok_stream_t stream_) {
  SETOK_NVTX("setok a8 mm_projector");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(p && tokens && out && rows > 0, SETOK_ERR_BAD_ARG, "project: null pointer or rows <= 0");
  SETOK_REQUIRE(p->n_linear >= 1 && p->w && p->b && p->dims, SETOK_ERR_BAD_ARG, "project: bad projector description");
  for (int i = 0; i <= p->n_linear; ++i) SETOK_REQUIRE(p->dims[i] > 0 && p->dims[i] % 8 == 0, SETOK_ERR_UNSUPPORTED, "project: dim %d must be a positive multiple of 8", p->dims[i]);
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_project_workspace_bytes(p, rows), SETOK_ERR_WORKSPACE, "project: workspace too small");
  int maxd = 0;
  for (int i = 0; i <= p->n_linear; ++i) maxd = p->dims[i] > maxd ? p->dims[i] : maxd;
  Arena a(workspace, workspace_bytes);
  bf16* ping = a.take<bf16>(static_cast<size_t>(rows) * maxd);
  bf16* pong = a.take<bf16>(static_cast<size_t>(rows) * maxd);
  float* tmp = a.take<float>(static_cast<size_t>(rows) * maxd);
  const void* cur = tokens;
  if (token_dtype == SETOK_F32) {
    SETOK_TRY(launch_convert_rows(tokens, SETOK_F32, ping, SETOK_BF16, rows, p->dims[0], SETOK_ACT_NONE, m_dev, stream));
    cur = ping;
  } else {
    SETOK_REQUIRE(token_dtype == SETOK_BF16, SETOK_ERR_BAD_ARG, "project: bad token dtype %d", token_dtype);
  }
  const bool use_norm = p->norm_g != nullptr;
  for (int i = 0; i < p->n_linear; ++i) {
    const int in_d = p->dims[i], out_d = p->dims[i + 1];
    const bool last = i == p->n_linear - 1;
    bf16* dst = (cur == ping) ? pong : ping;
    SETOK_REQUIRE(p->w[i] && p->b[i], SETOK_ERR_BAD_ARG, "project: null weights for linear %d", i);
    if (i == 0 && use_norm) {
      // Linear -> LayerNorm -> (GELU) (builder.py:48-58 with '_Norm')
      SETOK_TRY(launch_gemm(GemmArgs{cur, in_d, p->w[i], in_d, tmp, out_d, SETOK_F32, p->b[i], nullptr, 0, 0, SETOK_ACT_NONE, rows, out_d, in_d, m_dev, 0}, stream));
      if (last) {
        SETOK_TRY(launch_layernorm(tmp, SETOK_F32, out, out_dtype, p->norm_g, p->norm_b, 1e-5f, rows, out_d, nullptr, m_dev, stream));
      } else {
        SETOK_TRY(launch_layernorm(tmp, SETOK_F32, tmp, SETOK_F32, p->norm_g, p->norm_b, 1e-5f, rows, out_d, nullptr, m_dev, stream));
        SETOK_TRY(launch_convert_rows(tmp, SETOK_F32, dst, SETOK_BF16, rows, out_d, SETOK_ACT_GELU_ERF, m_dev, stream));
        cur = dst;
      }
      continue;
    }
    if (last) {
      SETOK_TRY(launch_gemm(GemmArgs{cur, in_d, p->w[i], in_d, out, out_d, out_dtype, p->b[i], nullptr, 0, 0, SETOK_ACT_NONE, rows, out_d, in_d, m_dev, 0}, stream));
    } else {
      SETOK_TRY(launch_gemm(GemmArgs{cur, in_d, p->w[i], in_d, dst, out_d, SETOK_BF16, p->b[i], nullptr, 0, 0, SETOK_ACT_GELU_ERF, rows, out_d, in_d, m_dev, 0}, stream));
      cur = dst;
    }
  }
  return SETOK_OK;
}

// =================================================================================================
namespace {
struct DetokBufs { bf16 *tokb, *enc, *kv, *h, *a, *q, *qkv, *ctx, *u, *x, *hx; float* t32; int32_t* idx; };

void detok_carve(const setok_detok* d, int B, int cap, Arena& a, DetokBufs* o) {
  const size_t R = static_cast<size_t>(B) * d->grid * d->grid;
  const size_t H = d->hidden, Dd = d->dec_dim;
  const size_t W = H > Dd ? H : Dd;
  const size_t U = static_cast<size_t>(d->q_inter > d->dec_mlp ? d->q_inter : d->dec_mlp);
  o->tokb = a.take<bf16>(static_cast<size_t>(cap) * d->token_dim);
  o->enc = a.take<bf16>(static_cast<size_t>(cap) * H);
  o->kv = a.take<bf16>(static_cast<size_t>(cap) * 2 * H);
  o->idx = a.take<int32_t>(R);
  o->h = a.take<bf16>(R * H);
  o->a = a.take<bf16>(R * H);
  o->q = a.take<bf16>(R * H);
  o->qkv = a.take<bf16>(R * 3 * W);
  o->ctx = a.take<bf16>(R * W);
  o->u = a.take<bf16>(R * U);
  o->x = a.take<bf16>(R * Dd);
  o->hx = a.take<bf16>(R * Dd);
  o->t32 = a.take<float>(R * W);
}
}  // namespace

extern "C" size_t setok_detok_workspace_bytes(const setok_detok* d, int B, int rows_capacity) {
  if (d == nullptr || B <= 0 || rows_capacity <= 0 || d->grid <= 0) return 0;
  Arena a(nullptr, 0);
  DetokBufs b;
  detok_carve(d, B, rows_capacity, a, &b);
  return a.off;
}

extern "C" int setok_detok_forward(const setok_detok* d, const void* tokens, int token_dtype, const int32_t* offsets, int B, int cap,
                                   void* out, int out_dtype, void* workspace, size_t workspace_bytes, setok_stream_t stream_) {
  SETOK_NVTX("setok a10 detokenizer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(d && tokens && offsets && out, SETOK_ERR_BAD_ARG, "detok_forward: null pointer");
  SETOK_REQUIRE(B > 0 && cap > 0 && d->grid > 0, SETOK_ERR_BAD_ARG, "detok_forward: B=%d rows_capacity=%d grid=%d", B, cap, d->grid);
  const int Q = d->grid * d->grid, H = d->hidden, I = d->q_inter, Dd = d->dec_dim, Fd = d->dec_mlp, Ct = d->token_dim;
  SETOK_REQUIRE(H > 0 && d->q_heads > 0 && H % d->q_heads == 0 && (H / d->q_heads) % 8 == 0 && H % 8 == 0 && I % 8 == 0 && Ct % 8 == 0,
                SETOK_ERR_UNSUPPORTED, "detok_forward: Q-Former hidden %d heads %d inter %d token_dim %d unsupported", H, d->q_heads, I, Ct);
  SETOK_REQUIRE(Dd > 0 && d->dec_heads > 0 && Dd % d->dec_heads == 0 && (Dd / d->dec_heads) % 8 == 0 && Dd % 8 == 0 && Fd % 8 == 0,
                SETOK_ERR_UNSUPPORTED, "detok_forward: decoder dim %d heads %d mlp %d unsupported", Dd, d->dec_heads, Fd);
  SETOK_REQUIRE(d->w_map_in && d->b_map_in && d->mask_tokens && d->emb_ln_g && d->emb_ln_b && (d->q_layers == 0 || d->qlayer) && d->w_dec_in &&
                d->b_dec_in && d->pos && (d->dec_depth == 0 || d->block) && d->norm_g && d->norm_b, SETOK_ERR_BAD_ARG, "detok_forward: null weights");
  SETOK_REQUIRE(token_dtype == SETOK_F32 || token_dtype == SETOK_BF16, SETOK_ERR_BAD_ARG, "detok_forward: bad token dtype %d", token_dtype);
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_detok_workspace_bytes(d, B, cap), SETOK_ERR_WORKSPACE, "detok_forward: workspace too small");
  const int R = B * Q;
  Arena ar(workspace, workspace_bytes);
  DetokBufs w;
  detok_carve(d, B, cap, ar, &w);
  const int32_t* n_tok = offsets + B;                      // device scalar: sum_b K_b
  const float qscale = 1.0f / std::sqrt(static_cast<float>(H / d->q_heads));

  // detokenizer.py:104: encoder states = mapper_fc_in(tokens), over the packed rows
  const void* tok = tokens;
  if (token_dtype == SETOK_F32) {
    SETOK_TRY(launch_convert_rows(tokens, SETOK_F32, w.tokb, SETOK_BF16, cap, Ct, SETOK_ACT_NONE, n_tok, stream));
    tok = w.tokb;
  }
  SETOK_TRY(launch_gemm(GemmArgs{tok, Ct, d->w_map_in, Ct, w.enc, H, SETOK_BF16, d->b_map_in, nullptr, 0, 0, SETOK_ACT_NONE, cap, H, Ct, n_tok, 0}, stream));
  // BertEmbeddings (module.py:203-205): LayerNorm of the learned queries, replicated per image
  SETOK_TRY(launch_iota_mod(w.idx, R, Q, stream));
  SETOK_TRY(launch_layernorm(d->mask_tokens, SETOK_F32, w.h, SETOK_BF16, d->emb_ln_g, d->emb_ln_b, d->q_ln_eps, R, H, w.idx, nullptr, stream));

  for (int l = 0; l < d->q_layers; ++l) {
    const setok_qformer_layer& L = d->qlayer[l];
    SETOK_REQUIRE(L.w_qkv && L.b_qkv && L.w_so && L.b_so && L.ln_s_g && L.ln_s_b && L.w_f1 && L.b_f1 && L.w_f2 && L.b_f2 && L.ln_f_g && L.ln_f_b,
                  SETOK_ERR_BAD_ARG, "detok_forward: null weights in Q-Former layer %d", l);
    // self-attention among the queries + BertSelfOutput (post-LN, module.py:383-387)
    SETOK_TRY(launch_gemm(GemmArgs{w.h, H, L.w_qkv, H, w.qkv, 3LL * H, SETOK_BF16, L.b_qkv, nullptr, 0, 0, SETOK_ACT_NONE, R, 3 * H, H, nullptr, 0}, stream));
    SETOK_TRY(launch_attention(w.qkv, w.ctx, R, H, d->q_heads, qscale, nullptr, nullptr, Q, nullptr, stream));
    SETOK_TRY(launch_gemm(GemmArgs{w.ctx, H, L.w_so, H, w.t32, H, SETOK_F32, L.b_so, w.h, H, SETOK_BF16, SETOK_ACT_NONE, R, H, H, nullptr, 0}, stream));
    SETOK_TRY(launch_layernorm(w.t32, SETOK_F32, w.a, SETOK_BF16, L.ln_s_g, L.ln_s_b, d->q_ln_eps, R, H, nullptr, nullptr, stream));
    if (L.has_cross) {
      SETOK_REQUIRE(L.w_cq && L.b_cq && L.w_ckv && L.b_ckv && L.w_co && L.b_co && L.ln_c_g && L.ln_c_b, SETOK_ERR_BAD_ARG,
                    "detok_forward: null cross-attention weights in Q-Former layer %d", l);
      // cross-attention to the image's own K_b tokens (module.py:528-544), varlen over the packed rows
      SETOK_TRY(launch_gemm(GemmArgs{w.a, H, L.w_cq, H, w.q, H, SETOK_BF16, L.b_cq, nullptr, 0, 0, SETOK_ACT_NONE, R, H, H, nullptr, 0}, stream));
      SETOK_TRY(launch_gemm(GemmArgs{w.enc, H, L.w_ckv, H, w.kv, 2LL * H, SETOK_BF16, L.b_ckv, nullptr, 0, 0, SETOK_ACT_NONE, cap, 2 * H, H, n_tok, 0}, stream));
      // the tensor-core kernel's last 64-key box of the batch reaches up to 63 rows past the last live row: keep them finite
      SETOK_TRY(launch_zero_tail_rows(w.kv, 2LL * H, n_tok, 64, cap, 2 * H, stream));
      SETOK_TRY(launch_cross_attention(w.q, w.kv, w.ctx, R, Q, H, d->q_heads, qscale, offsets, cap, stream));
      SETOK_TRY(launch_gemm(GemmArgs{w.ctx, H, L.w_co, H, w.t32, H, SETOK_F32, L.b_co, w.a, H, SETOK_BF16, SETOK_ACT_NONE, R, H, H, nullptr, 0}, stream));
      SETOK_TRY(launch_layernorm(w.t32, SETOK_F32, w.a, SETOK_BF16, L.ln_c_g, L.ln_c_b, d->q_ln_eps, R, H, nullptr, nullptr, stream));
    }
    // query FFN (module.py:579-582)
    SETOK_TRY(launch_gemm(GemmArgs{w.a, H, L.w_f1, H, w.u, I, SETOK_BF16, L.b_f1, nullptr, 0, 0, SETOK_ACT_GELU_ERF, R, I, H, nullptr, 0}, stream));
    SETOK_TRY(launch_gemm(GemmArgs{w.u, I, L.w_f2, I, w.t32, H, SETOK_F32, L.b_f2, w.a, H, SETOK_BF16, SETOK_ACT_NONE, R, H, I, nullptr, 0}, stream));
    SETOK_TRY(launch_layernorm(w.t32, SETOK_F32, w.h, SETOK_BF16, L.ln_f_g, L.ln_f_b, d->q_ln_eps, R, H, nullptr, nullptr, stream));
  }

  // detokenizer.py:111-115: decoder_fc_in, + 2-D sincos position embedding
  SETOK_TRY(launch_gemm(GemmArgs{w.h, H, d->w_dec_in, H, w.t32, Dd, SETOK_F32, d->b_dec_in, nullptr, 0, 0, SETOK_ACT_NONE, R, Dd, H, nullptr, 0}, stream));
  SETOK_TRY(launch_add_pos_rows(w.t32, d->pos, w.x, R, Q, Dd, stream));
  // :117-118 pixel_decoder (timm Block = pre-LN ViT layer with GELU(erf)); :120 decoder_norm
  for (int l = 0; l < d->dec_depth; ++l) {
    const setok_vit_layer& L = d->block[l];
    SETOK_REQUIRE(L.w_qkv && L.b_qkv && L.w_o && L.b_o && L.w_fc1 && L.b_fc1 && L.w_fc2 && L.b_fc2 && L.ln1_g && L.ln1_b && L.ln2_g && L.ln2_b,
                  SETOK_ERR_BAD_ARG, "detok_forward: null weights in decoder block %d", l);
    SETOK_TRY(run_preln_layer(L, w.x, SETOK_BF16, w.hx, w.qkv, w.ctx, w.u, R, Q, Dd, Fd, d->dec_heads, d->dec_ln_eps, SETOK_ACT_GELU_ERF, stream));
  }
  SETOK_TRY(launch_layernorm(w.x, SETOK_BF16, out, out_dtype, d->norm_g, d->norm_b, d->dec_ln_eps, R, Dd, nullptr, nullptr, stream));
  return SETOK_OK;
}
