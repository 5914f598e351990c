// splice.cu — the multimodal splice of SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal
// (reference src/model/setokim_arch.py:241-354) as four launches, no host synchronisation:
//   1. splice_count_kernel   one CTA per sample: valid tokens, image placeholders (attention-mask compaction :255-256)
//   2. splice_plan_kernel    one CTA: running image index across the batch (a sample without a placeholder consumes one
//                            image, :262-269), new lengths incl. the K rows of every image, truncation (:307-310), max length
//   3. splice_fill_kernel    one CTA per sample: a block scan turns every valid token into its destination column and
//                            writes labels / mask / position ids and one int32 source descriptor per output column
//   4. splice_gather_kernel  grid-stride row copy by descriptor (embedding-table row, image row, or zero padding) with
//                            16-byte vectors -- the only HBM-heavy part: B x T x H elements in, the same out.
#include "common.cuh"

namespace setok {
namespace {

constexpr long long IMAGE_TOKEN = -200, IGNORE = -100, TARGET_TOKEN = -300;
constexpr int SP_THREADS = 256;

__device__ __forceinline__ bool tok_valid(const uint8_t* mask, long long i) { return mask == nullptr || mask[i] != 0; }

__global__ void __launch_bounds__(SP_THREADS) splice_count_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, int L,
                                                                  int32_t* __restrict__ n_valid, int32_t* __restrict__ n_img) {
  __shared__ int sv[SP_THREADS / 32], si[SP_THREADS / 32];
  const int b = blockIdx.x;
  int v = 0, im = 0;
  for (int t = threadIdx.x; t < L; t += SP_THREADS) {
    const long long i = static_cast<long long>(b) * L + t;
    if (tok_valid(mask, i)) { ++v; im += ids[i] == IMAGE_TOKEN ? 1 : 0; }
  }
  v = __reduce_add_sync(0xffffffffu, v);
  im = __reduce_add_sync(0xffffffffu, im);
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = im; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tv = 0, ti = 0;
    for (int w = 0; w < SP_THREADS / 32; ++w) { tv += sv[w]; ti += si[w]; }
    n_valid[b] = tv; n_img[b] = ti;
  }
}

// img_start[b] = images consumed by samples < b; full_len[b] = text tokens + rows of the sample's images
__global__ void splice_plan_kernel(const int32_t* __restrict__ n_valid, const int32_t* __restrict__ n_img, const int32_t* __restrict__ img_off,
                                   int n_images, int B, int max_length, int out_cap, int32_t* __restrict__ img_start,
                                   int32_t* __restrict__ lens, int32_t* __restrict__ max_len) {
  if (threadIdx.x != 0) return;
  int cur = 0, mx = 0;
  for (int b = 0; b < B; ++b) {
    img_start[b] = cur;
    const int ni = n_img[b];
    int rows = 0;
    for (int i = 0; i < ni; ++i) {
      const int g = cur + i;
      if (g < n_images) rows += img_off[g + 1] - img_off[g];
    }
    cur += ni > 0 ? ni : 1;
    int len = n_valid[b] - ni + rows;
    if (max_length > 0 && len > max_length) len = max_length;
    if (len > out_cap) len = out_cap;
    lens[b] = len;
    mx = len > mx ? len : mx;
  }
  *max_len = mx;
}

__global__ void __launch_bounds__(SP_THREADS) splice_fill_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask,
                                                                 const int64_t* __restrict__ labels, int L, int V, const int32_t* __restrict__ img_off,
                                                                 int n_images, const int32_t* __restrict__ img_start, const int32_t* __restrict__ lens,
                                                                 const int32_t* __restrict__ max_len_p, int pad_left, int out_cap,
                                                                 int64_t* __restrict__ labels_out, uint8_t* __restrict__ mask_out,
                                                                 int64_t* __restrict__ pos_out, int32_t* __restrict__ src) {
  __shared__ int s_text[SP_THREADS], s_img[SP_THREADS];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int len = lens[b], max_len = *max_len_p;
  const int shift = pad_left ? max_len - len : 0;            // :323-331 left padding puts the sequence at the end
  int32_t* srcb = src + static_cast<long long>(b) * out_cap;
  int64_t* lab_o = labels_out + static_cast<long long>(b) * out_cap;
  uint8_t* msk_o = mask_out + static_cast<long long>(b) * out_cap;
  int64_t* pos_o = pos_out + static_cast<long long>(b) * out_cap;
  // defaults: padding everywhere (:316-318)
  for (int t = tid; t < out_cap; t += SP_THREADS) {
    const bool in = t >= shift && t < shift + len;
    srcb[t] = INT32_MIN;                                      // zero row
    lab_o[t] = IGNORE;
    msk_o[t] = in ? 1 : 0;
    pos_o[t] = in ? t - shift : 0;
  }
  // each thread owns a contiguous chunk of the input row; exclusive scans of (valid text tokens, placeholders) give the
  // destination of every token: text before + rows of the images before
  const int chunk = (L + SP_THREADS - 1) / SP_THREADS;
  const int t0 = tid * chunk, t1 = min(L, t0 + chunk);
  int ct = 0, ci = 0;
  for (int t = t0; t < t1; ++t) {
    const long long i = static_cast<long long>(b) * L + t;
    if (tok_valid(mask, i)) { if (ids[i] == IMAGE_TOKEN) ++ci; else ++ct; }
  }
  s_text[tid] = ct; s_img[tid] = ci;
  __syncthreads();
  if (tid == 0) {
    int rt = 0, ri = 0;
    for (int w = 0; w < SP_THREADS; ++w) { const int a = s_text[w], c = s_img[w]; s_text[w] = rt; s_img[w] = ri; rt += a; ri += c; }
  }
  __syncthreads();
  int text_before = s_text[tid], img_before = s_img[tid];
  const int g0 = img_start[b];
  auto rows_before = [&](int k) {                              // rows of this sample's first k images
    const int lo = g0 < n_images ? g0 : n_images, hi = g0 + k < n_images ? g0 + k : n_images;
    return img_off[hi] - img_off[lo];
  };
  for (int t = t0; t < t1; ++t) {
    const long long i = static_cast<long long>(b) * L + t;
    if (!tok_valid(mask, i)) continue;
    const long long id = ids[i];
    const int dst = text_before + rows_before(img_before);
    if (id == IMAGE_TOKEN) {
      const int g = g0 + img_before;
      if (g < n_images) {
        const int r0 = img_off[g], k = img_off[g + 1] - r0;
        for (int r = 0; r < k; ++r)
          if (dst + r < len) srcb[shift + dst + r] = -1 - (r0 + r);       // image row; label stays IGNORE (:297)
      }
      ++img_before;
    } else {
      if (dst < len) {
        srcb[shift + dst] = id >= 0 && id < V ? static_cast<int32_t>(id) : INT32_MIN;
        if (labels != nullptr) { const long long lb = labels[i]; lab_o[shift + dst] = lb == TARGET_TOKEN ? IGNORE : lb; }
      }
      ++text_before;
    }
  }
}

// embeds[b, t, :] = embed[id] | image_rows[r] | 0, by the source descriptors; grid-stride over 16-byte vectors
template <class T>
__global__ void __launch_bounds__(256) splice_gather_kernel(const int32_t* __restrict__ src, const T* __restrict__ embed, const T* __restrict__ img_rows,
                                                            int H, long long rows, T* __restrict__ embeds) {
  const int nvec = H * static_cast<int>(sizeof(T)) / 16;
  const long long total = rows * nvec;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = e / nvec;
    const int vi = static_cast<int>(e % nvec);
    const int sdesc = src[r];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sdesc >= 0) v = __ldg(reinterpret_cast<const uint4*>(embed + static_cast<long long>(sdesc) * H) + vi);
    else if (sdesc != INT32_MIN) v = __ldg(reinterpret_cast<const uint4*>(img_rows + static_cast<long long>(-1 - sdesc) * H) + vi);
    reinterpret_cast<uint4*>(embeds + r * H)[vi] = v;
  }
}

}  // namespace
}  // namespace setok

using namespace setok;

extern "C" size_t setok_splice_workspace_bytes(int B, int L, int out_cap) {
  if (B <= 0 || L <= 0 || out_cap <= 0) return 0;
  Arena a(nullptr, 0);
  a.take<int32_t>(B); a.take<int32_t>(B); a.take<int32_t>(B);      // the plan lives in the first three slices: the same for any out_cap
  a.take<int32_t>(static_cast<size_t>(B) * out_cap);
  return a.off;
}

namespace {
struct SpliceWs { int32_t *n_valid, *n_img, *img_start, *src; };
void splice_carve(int B, int out_cap, Arena& a, SpliceWs* w) {
  w->n_valid = a.take<int32_t>(B); w->n_img = a.take<int32_t>(B); w->img_start = a.take<int32_t>(B);
  w->src = a.take<int32_t>(static_cast<size_t>(B) * out_cap);
}
}  // namespace

// Phase 1: lengths only (lens, max_len on the device).  A caller that wants exactly-sized outputs reads max_len once and
// passes it to setok_splice_fill as out_cap; the workspace carries the plan between the two calls.
extern "C" int setok_splice_plan(const int64_t* input_ids, const uint8_t* attention_mask, int B, int L, const int32_t* image_offsets, int n_images,
                                 int max_length, int out_cap_limit, int32_t* lens, int32_t* max_len, void* workspace, size_t workspace_bytes,
                                 setok_stream_t stream_) {
  SETOK_NVTX("setok a9 splice plan");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(input_ids && image_offsets && lens && max_len, SETOK_ERR_BAD_ARG, "splice_plan: null pointer");
  SETOK_REQUIRE(B > 0 && L > 0 && n_images >= 0 && out_cap_limit > 0, SETOK_ERR_BAD_ARG, "splice_plan: bad shape B=%d L=%d", B, L);
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_splice_workspace_bytes(B, L, 1), SETOK_ERR_WORKSPACE, "splice_plan: workspace too small");
  Arena a(workspace, workspace_bytes);
  SpliceWs w;
  splice_carve(B, 1, a, &w);
  splice_count_kernel<<<B, SP_THREADS, 0, stream>>>(input_ids, attention_mask, L, w.n_valid, w.n_img);
  SETOK_LAUNCH_CHECK();
  splice_plan_kernel<<<1, 32, 0, stream>>>(w.n_valid, w.n_img, image_offsets, n_images, B, max_length, out_cap_limit, w.img_start, lens, max_len);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

// Phase 2: descriptors + row gather into outputs of out_cap columns (lens / max_len / workspace from setok_splice_plan).
extern "C" int setok_splice_fill(const int64_t* input_ids, const uint8_t* attention_mask, const int64_t* labels, int B, int L, const void* embed,
                                 int dtype, int V, int H, const void* image_rows, const int32_t* image_offsets, int n_images, int pad_left,
                                 int out_cap, const int32_t* lens, const int32_t* max_len, void* embeds, int64_t* labels_out, uint8_t* mask_out,
                                 int64_t* pos_out, void* workspace, size_t workspace_bytes, setok_stream_t stream_) {
  SETOK_NVTX("setok a9 splice fill");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(input_ids && embed && image_rows && image_offsets && embeds && labels_out && mask_out && pos_out && lens && max_len,
                SETOK_ERR_BAD_ARG, "splice_fill: null pointer");
  SETOK_REQUIRE(B > 0 && L > 0 && V > 0 && H > 0 && n_images >= 0 && out_cap > 0, SETOK_ERR_BAD_ARG, "splice_fill: bad shape B=%d L=%d V=%d H=%d out_cap=%d", B, L, V, H, out_cap);
  SETOK_REQUIRE(dtype == SETOK_F32 || dtype == SETOK_BF16, SETOK_ERR_BAD_ARG, "splice_fill: bad dtype %d", dtype);
  const int esz = dtype == SETOK_F32 ? 4 : 2;
  SETOK_REQUIRE((H * esz) % 16 == 0 && aligned16(embed) && aligned16(image_rows) && aligned16(embeds), SETOK_ERR_UNSUPPORTED,
                "splice_fill: rows must be 16-byte multiples and 16-byte aligned (H=%d)", H);
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_splice_workspace_bytes(B, L, out_cap), SETOK_ERR_WORKSPACE, "splice_fill: workspace too small");
  Arena a(workspace, workspace_bytes);
  SpliceWs w;
  splice_carve(B, out_cap, a, &w);
  splice_fill_kernel<<<B, SP_THREADS, 0, stream>>>(input_ids, attention_mask, labels, L, V, image_offsets, n_images, w.img_start, lens, max_len,
                                                   pad_left, out_cap, labels_out, mask_out, pos_out, w.src);
  SETOK_LAUNCH_CHECK();
  const long long rows = static_cast<long long>(B) * out_cap;
  const long long vecs = rows * (H * esz / 16);
  long long blocks = (vecs + 255) / 256;
  if (blocks > num_sms() * 16LL) blocks = num_sms() * 16LL;
  if (dtype == SETOK_F32)
    splice_gather_kernel<float><<<static_cast<int>(blocks), 256, 0, stream>>>(w.src, static_cast<const float*>(embed), static_cast<const float*>(image_rows), H, rows, static_cast<float*>(embeds));
  else
    splice_gather_kernel<bf16><<<static_cast<int>(blocks), 256, 0, stream>>>(w.src, static_cast<const bf16*>(embed), static_cast<const bf16*>(image_rows), H, rows, static_cast<bf16*>(embeds));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

// Both phases back to back for a caller with a known column bound (e.g. tokenizer_model_max_length): no host read at all.
extern "C" int setok_splice(const int64_t* input_ids, const uint8_t* attention_mask, const int64_t* labels, int B, int L, const void* embed,
                            int dtype, int V, int H, const void* image_rows, const int32_t* image_offsets, int n_images, int max_length,
                            int pad_left, int out_cap, void* embeds, int64_t* labels_out, uint8_t* mask_out, int64_t* pos_out, int32_t* lens,
                            int32_t* max_len, void* workspace, size_t workspace_bytes, setok_stream_t stream) {
  SETOK_TRY(setok_splice_plan(input_ids, attention_mask, B, L, image_offsets, n_images, max_length, out_cap, lens, max_len, workspace, workspace_bytes, stream));
  return setok_splice_fill(input_ids, attention_mask, labels, B, L, embed, dtype, V, H, image_rows, image_offsets, n_images, pad_left, out_cap, lens,
                           max_len, embeds, labels_out, mask_out, pos_out, workspace, workspace_bytes, stream);
}
