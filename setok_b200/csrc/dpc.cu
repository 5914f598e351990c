// dpc.cu — position-embedding add + DPC-kNN clustering (SetokTokenizer.cluster_dpc_knn,
// reference src/model/setok/tokenizer.py:78-121) for a whole batch, no host synchronisation.
//
// Three launches per batch:
//   1. posadd_sqnorm_kernel   x_pos = feats + pos (fp32), row squared norms           (HBM-bound)
//   2. gram_dist_kernel       D[b] = sqrt(max(n_i + n_j - 2 x_i.x_j, 0)) / sqrt(C)     (fp32 FMA; the
//                             matmul form torch.cdist takes for N > 25); upper-triangular tiles only, mirrored
//   3. dpc_select_kernel      one CTA per image on D[b] (L2-resident): kNN density by bitwise
//                             selection, the column-indexed row-max fill of tokenizer.py:98-99,
//                             score, threshold / top-k fallback, ordered compaction, argmin assignment.
//   4. offsets_scan_kernel    offsets = exclusive scan of per-image cluster counts.
// Every float operation the reference performs as a separate rounding step (mul then add, divide by
// sqrt(C), divide by k) is performed with explicit round-to-nearest intrinsics so nvcc cannot contract
// it into an FMA: the integer outputs depend on comparisons of these floats.
#include "common.cuh"
#include "rowops.cuh"

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace setok {
extern int g_dpc_tensor_gram;
namespace {

// ---- 1. x_pos = feats + pos; sqnorm ------------------------------------------------------------------
template <class TI>
__global__ void __launch_bounds__(256) posadd_sqnorm_kernel(const TI* __restrict__ feats, const float* __restrict__ pos,
                                                            float* __restrict__ x_pos, float* __restrict__ sqn, int B, int N, int C,
                                                            bf16* __restrict__ xh, bf16* __restrict__ xl) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long rows = static_cast<long long>(B) * N;
  for (long long r = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows; r += static_cast<long long>(gridDim.x) * wpb) {
    const int t = static_cast<int>(r % N);
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      float4 v;
      if (sizeof(TI) == 4) {
        v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(feats) + r * C + c);
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(feats) + r * C + c);
        const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y);
        v = make_float4(a.x, a.y, b2.x, b2.y);
      }
      if (pos != nullptr) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * C + c));
        v.x = __fadd_rn(v.x, p.x); v.y = __fadd_rn(v.y, p.y); v.z = __fadd_rn(v.z, p.z); v.w = __fadd_rn(v.w, p.w);
      }
      if (x_pos != nullptr) *reinterpret_cast<float4*>(x_pos + r * C + c) = v;
      if (xh != nullptr) {
        // x = hi + lo + O(2^-17 |x|) with hi, lo both bf16: the Gram then runs on the bf16 tensor cores with exact
        // products (hi.hi + hi.lo + lo.hi + lo.lo) and fp32 accumulation
        const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
        const float2 a = unpack_bf16x2(h0), b2 = unpack_bf16x2(h1);
        *reinterpret_cast<uint2*>(xh + r * C + c) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(xl + r * C + c) = make_uint2(pack_bf16x2(v.x - a.x, v.y - a.y), pack_bf16x2(v.z - b2.x, v.w - b2.y));
      }
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = warp_sum(s);
    if (lane == 0) sqn[r] = s;
  }
}

// ---- 2. pairwise distances -----------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256) gram_dist_kernel(const float* __restrict__ x, const float* __restrict__ sqn,
                                                        float* __restrict__ D, int N, int C, float sqrtC) {
  __shared__ __align__(16) float As[GK][GT + 4];
  __shared__ __align__(16) float Bs[GK][GT + 4];
  const int b = blockIdx.z;
  // D is symmetric by construction (same fp32 operations for (i,j) and (j,i)): only tiles on or above the diagonal
  // are computed, each off-diagonal tile is also written transposed
  if (blockIdx.x < blockIdx.y) return;
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  const float* xb = x + static_cast<long long>(b) * N * C;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < C; k0 += GK) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (k0 + lk < C) {
      if (i0 + lr < N) va = *reinterpret_cast<const float4*>(xb + static_cast<long long>(i0 + lr) * C + k0 + lk);
      if (j0 + lr < N) vb = *reinterpret_cast<const float4*>(xb + static_cast<long long>(j0 + lr) * C + k0 + lk);
    }
    __syncthreads();
    As[lk][lr] = va.x; As[lk + 1][lr] = va.y; As[lk + 2][lr] = va.z; As[lk + 3][lr] = va.w;
    Bs[lk][lr] = vb.x; Bs[lk + 1][lr] = vb.y; Bs[lk + 2][lr] = vb.z; Bs[lk + 3][lr] = vb.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
  }
  const float* nb = sqn + static_cast<long long>(b) * N;
  float* Db = D + static_cast<long long>(b) * N * N;
  const bool mirror = blockIdx.x != blockIdx.y;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= N) continue;
    const float ni = nb[i];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= N) continue;
      const float d2 = fmaf(-2.0f, acc[a][c], __fadd_rn(ni, nb[j]));
      const float d = __fdiv_rn(sqrtf(fmaxf(d2, 0.f)), sqrtC);
      Db[static_cast<long long>(i) * N + j] = d;
      if (mirror) Db[static_cast<long long>(j) * N + i] = d;
    }
  }
}

// G (x_i . x_j, from the tensor-core path) -> D in place, same formula as gram_dist_kernel's epilogue
__global__ void __launch_bounds__(256) gram_to_dist_kernel(float* __restrict__ D, const float* __restrict__ sqn, int N, float sqrtC, long long total) {
  const long long nn = static_cast<long long>(N) * N;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = e / nn;
    const int i = static_cast<int>((e % nn) / N), j = static_cast<int>(e % N);
    const float* nb = sqn + b * N;
    const float d2 = fmaf(-2.0f, D[e], __fadd_rn(nb[i], nb[j]));
    D[e] = __fdiv_rn(sqrtf(fmaxf(d2, 0.f)), sqrtC);
  }
}

// ---- 3. per-image selection ----------------------------------------------------------------------------
constexpr int SEL_THREADS = 512;
constexpr int SEL_MAXN = 1024;

template <int MAXT>
__device__ __forceinline__ void load_row(const float* __restrict__ Drow, int N, int lane, const float* mask_s, float fill,
                                         float (&v)[MAXT]) {
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    const int j = lane + 32 * t;
    float d = 0.f;
    if (j < N) {
      d = Drow[j];
      if (mask_s != nullptr && !(mask_s[j] > 0.f)) d = fill;
    }
    v[t] = d;
  }
}

template <int MAXT>
__global__ void __launch_bounds__(SEL_THREADS) dpc_select_kernel(const float* __restrict__ D, const float* __restrict__ noise,
                                                                 const float* __restrict__ token_mask, int N, int k, float threshold,
                                                                 int min_cluster_num, int64_t* __restrict__ idx_cluster,
                                                                 float* __restrict__ score_out, int64_t* __restrict__ index_down,
                                                                 int32_t* __restrict__ num_clusters) {
  __shared__ float dens[SEL_MAXN], rmax[SEL_MAXN], score[SEL_MAXN], mask_buf[SEL_MAXN];
  __shared__ int cidx[SEL_MAXN];
  __shared__ float red[SEL_THREADS / 32];
  __shared__ int wcount[SEL_THREADS / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = SEL_THREADS / 32;
  const float* Db = D + static_cast<long long>(b) * N * N;
  const float* mask_s = nullptr;
  float fill = 0.f;
  if (token_mask != nullptr) {
    // tokenizer.py:84-86: masked columns are pushed to (global max + 1)
    float mx = -INFINITY;
    for (long long e = tid; e < static_cast<long long>(N) * N; e += SEL_THREADS) mx = fmaxf(mx, Db[e]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    for (int j = tid; j < N; j += SEL_THREADS) mask_buf[j] = token_mask[static_cast<long long>(b) * N + j];
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < NW; ++w) mx = fmaxf(mx, red[w]);
    fill = __fadd_rn(mx, 1.0f);
    mask_s = mask_buf;
    __syncthreads();
  }

  // phase 1: kNN density + row max
  for (int i = warp; i < N; i += NW) {
    float v[MAXT];
    load_row<MAXT>(Db + static_cast<long long>(i) * N, N, lane, mask_s, fill, v);
    float mx = -INFINITY;
    uint32_t bits[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const bool ok = lane + 32 * t < N;
      if (ok) mx = fmaxf(mx, v[t]);
      bits[t] = ok ? __float_as_uint(v[t]) : 0xFFFFFFFFu;
    }
    mx = warp_max(mx);
    // smallest X with |{v <= X}| >= k  (distances are >= +0, so uint order == float order)
    uint32_t lo = 0u, hi = 0x7F800000u;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int t = 0; t < MAXT; ++t) c += bits[t] <= mid ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) hi = mid; else lo = mid + 1u;
    }
    const float kth = __uint_as_float(lo);
    float s = 0.f;
    int less = 0;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      if (bits[t] < lo) { s = __fadd_rn(s, __fmul_rn(v[t], v[t])); ++less; }
    }
    s = warp_sum(s);
    less = __reduce_add_sync(0xffffffffu, less);
    const float kk = __fmul_rn(kth, kth);
    for (int e = less; e < k; ++e) s = __fadd_rn(s, kk);
    if (lane == 0) {
      const float mean = __fdiv_rn(s, static_cast<float>(k));
      float d = __fadd_rn(expf(-mean), __fmul_rn(noise[static_cast<long long>(b) * N + i], 1e-6f));
      if (mask_s != nullptr) d = __fmul_rn(d, mask_s[i] > 0.f ? 1.0f : 0.0f);
      dens[i] = d;
      rmax[i] = mx;
    }
  }
  __syncthreads();

  // phase 2: distance to the nearest higher-density token; where no such relation holds the fill is
  // rowmax[j] (column-indexed: tokenizer.py:98-99 broadcasts dist_max of shape (1,1,N) along the last axis)
  for (int i = warp; i < N; i += NW) {
    float v[MAXT];
    load_row<MAXT>(Db + static_cast<long long>(i) * N, N, lane, mask_s, fill, v);
    const float di = dens[i];
    float pd = INFINITY;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int j = lane + 32 * t;
      if (j < N) pd = fminf(pd, dens[j] > di ? v[t] : rmax[j]);
    }
    pd = warp_min(pd);
    if (lane == 0) {
      const float sc = __fmul_rn(pd, di);
      score[i] = sc;
      score_out[static_cast<long long>(b) * N + i] = sc;
    }
  }
  __syncthreads();

  // phase 3: centres = {score > threshold}; if empty, the top-min_cluster_num scores (tokenizer.py:103-107)
  int any = 0;
  for (int i = tid; i < N; i += SEL_THREADS) any |= score[i] > threshold ? 1 : 0;
  const int total_pass = __syncthreads_or(any);
  int base = 0;
  for (int i0 = 0; i0 < N; i0 += SEL_THREADS) {
    const int i = i0 + tid;
    bool f = false;
    if (i < N) {
      if (total_pass) {
        f = score[i] > threshold;
      } else {
        const float si = score[i];
        int rank = 0;
        for (int j = 0; j < N; ++j) {
          const float sj = score[j];
          rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
        }
        f = rank < min_cluster_num;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wcount[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < NW; ++w) { const int c = wcount[w]; if (w < warp) woff += c; tot += c; }
    if (f) cidx[base + woff + __popc(bal & ((1u << lane) - 1u))] = i;
    base += tot;
    __syncthreads();
  }
  const int K = base;
  if (tid == 0) num_clusters[b] = K;
  for (int c = tid; c < N; c += SEL_THREADS) index_down[static_cast<long long>(b) * N + c] = c < K ? static_cast<int64_t>(cidx[c]) : -1;

  // phase 4: nearest centre (first minimum), then every centre owns its own label (tokenizer.py:111-119)
  for (int j = tid; j < N; j += SEL_THREADS) {
    float best = INFINITY;
    int bi = 0;
    const bool masked = mask_s != nullptr && !(mask_s[j] > 0.f);
    for (int c = 0; c < K; ++c) {
      const float d = masked ? fill : Db[static_cast<long long>(cidx[c]) * N + j];
      if (d < best) { best = d; bi = c; }
    }
    score[j] = __int_as_float(bi);   // reuse smem as the label buffer
  }
  __syncthreads();
  for (int c = tid; c < K; c += SEL_THREADS) score[cidx[c]] = __int_as_float(c);
  __syncthreads();
  for (int j = tid; j < N; j += SEL_THREADS) idx_cluster[static_cast<long long>(b) * N + j] = static_cast<int64_t>(__float_as_int(score[j]));
}

__global__ void offsets_scan_kernel(const int32_t* __restrict__ counts, int B, int32_t* __restrict__ offsets) {
  extern __shared__ int32_t sc[];
  for (int i = threadIdx.x; i < B; i += blockDim.x) sc[i] = counts[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < B; ++i) { offsets[i] = run; run += sc[i]; }
    offsets[B] = run;
  }
}

// ---- host: cached sincos table (fallback when the caller does not pass one) ----------------------------
struct PosKey { int dev, h, w, C; bool operator<(const PosKey& o) const { return std::tie(dev, h, w, C) < std::tie(o.dev, o.h, o.w, o.C); } };
std::mutex g_pos_mu;
std::map<PosKey, float*> g_pos_cache;

}  // namespace

// PositionalEncoding2D (reference module.py:105-146): ch = ceil(C/4)*2, inv_freq_i = 10000^(-2i/ch),
// channel 2i/2i+1 of the first ch <- sin/cos(row * inv_freq_i), of the next ch <- sin/cos(col * inv_freq_i).
int get_pos_table(int h, int w, int C, const float** out, cudaStream_t stream) {
  int dev = 0;
  SETOK_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_pos_mu);
  const PosKey key{dev, h, w, C};
  auto it = g_pos_cache.find(key);
  if (it != g_pos_cache.end()) { *out = it->second; return SETOK_OK; }
  const int ch = static_cast<int>(std::ceil(C / 4.0)) * 2;
  std::vector<float> tab(static_cast<size_t>(h) * w * C);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < C; ++c) {
        const int cc = c < ch ? c : c - ch;
        if (c >= 2 * ch) { tab[(static_cast<size_t>(y) * w + x) * C + c] = 0.f; continue; }
        const float inv = 1.0f / std::pow(10000.0f, static_cast<float>(cc / 2 * 2) / static_cast<float>(ch));
        const float arg = static_cast<float>(c < ch ? y : x) * inv;
        tab[(static_cast<size_t>(y) * w + x) * C + c] = static_cast<float>((cc & 1) ? std::cos(static_cast<double>(arg)) : std::sin(static_cast<double>(arg)));
      }
  float* d = nullptr;
  SETOK_CUDA_OK(cudaMalloc(&d, tab.size() * sizeof(float)));
  SETOK_CUDA_OK(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  SETOK_CUDA_OK(cudaStreamSynchronize(stream));   // first use only: `tab` is a host temporary
  g_pos_cache[key] = d;
  *out = d;
  return SETOK_OK;
}

int g_dpc_tensor_gram = 1;   // 0 forces the fp32-FMA Gram kernel (A/B via setok_debug_set_dpc_tensor_gram)

}  // namespace setok

using namespace setok;

extern "C" void setok_debug_set_dpc_tensor_gram(int on) { g_dpc_tensor_gram = on; }

extern "C" size_t setok_dpc_workspace_bytes(int B, int N, int C) {
  Arena a(nullptr, 0);
  a.take<float>(static_cast<size_t>(B) * N);          // squared norms
  a.take<float>(static_cast<size_t>(B) * N * N);      // distance matrices
  a.take<bf16>(static_cast<size_t>(B) * N * C);       // bf16 split of x_pos: hi
  a.take<bf16>(static_cast<size_t>(B) * N * C);       //                      lo
  return a.off;
}

namespace {
// embedded: `feats` already carries the position embedding (tokenizer.py:168 done upstream): no table, no x_pos output
int dpc_cluster_impl(const void* feats, int feat_dtype, const float* pos_table, bool embedded, const float* noise,
                     const float* token_mask, int B, int h, int w, int C, int k, float threshold,
                     int min_cluster_num, float* x_pos, int64_t* idx_cluster, float* score, int64_t* index_down,
                     int32_t* num_clusters, int32_t* offsets, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream) {
  SETOK_NVTX("setok a3+a4 position embedding + DPC-kNN clustering");
  const int N = h * w;
  SETOK_REQUIRE(feats && noise && (x_pos || embedded) && idx_cluster && score && index_down && num_clusters && offsets, SETOK_ERR_BAD_ARG, "dpc_cluster: null pointer");
  SETOK_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0, SETOK_ERR_BAD_ARG, "dpc_cluster: non-positive shape");
  SETOK_REQUIRE(C % 4 == 0, SETOK_ERR_UNSUPPORTED, "dpc_cluster: C (%d) must be a multiple of 4", C);
  SETOK_REQUIRE(N <= SEL_MAXN, SETOK_ERR_UNSUPPORTED, "dpc_cluster: N (%d) exceeds %d", N, SEL_MAXN);
  SETOK_REQUIRE(k >= 1 && k <= N, SETOK_ERR_BAD_ARG, "dpc_cluster: k (%d) out of range for N=%d (torch.topk would raise)", k, N);
  SETOK_REQUIRE(min_cluster_num >= 1 && min_cluster_num <= N, SETOK_ERR_BAD_ARG, "dpc_cluster: min_cluster_num (%d) out of range for N=%d", min_cluster_num, N);
  SETOK_REQUIRE(aligned16(feats) && aligned16(x_pos), SETOK_ERR_BAD_ARG, "dpc_cluster: feats/x_pos must be 16-byte aligned");
  SETOK_REQUIRE(feat_dtype == SETOK_F32 || feat_dtype == SETOK_BF16, SETOK_ERR_BAD_ARG, "dpc_cluster: bad feature dtype %d", feat_dtype);
  SETOK_REQUIRE(workspace && workspace_bytes >= setok_dpc_workspace_bytes(B, N, C), SETOK_ERR_WORKSPACE, "dpc_cluster: workspace too small");
  Arena a(workspace, workspace_bytes);
  float* sqn = a.take<float>(static_cast<size_t>(B) * N);
  float* D = a.take<float>(static_cast<size_t>(B) * N * N);
  bf16* xh = a.take<bf16>(static_cast<size_t>(B) * N * C);
  bf16* xl = a.take<bf16>(static_cast<size_t>(B) * N * C);
  // Gram on the tensor cores when the shapes allow the batched GEMM (leading dimensions multiples of 8 / 4)
  const bool tensor_gram = (C % 8 == 0) && (N % 4 == 0) && g_dpc_tensor_gram;
  const float* pos = embedded ? nullptr : pos_table;
  if (pos == nullptr && !embedded) SETOK_TRY(get_pos_table(h, w, C, &pos, stream));
  SETOK_REQUIRE(aligned16(pos), SETOK_ERR_BAD_ARG, "dpc_cluster: pos table must be 16-byte aligned");
  if (embedded) x_pos = nullptr;

  if (dpc_fused_supported(N, C, k) && (feat_dtype == SETOK_F32 || feat_dtype == SETOK_BF16)) {
    // N <= 256: one persistent kernel, features read once, distances stay in tensor memory (dpc_fused.cu)
    SETOK_TRY(launch_dpc_fused(feats, feat_dtype, pos, noise, token_mask, B, N, C, k, threshold, min_cluster_num, x_pos, idx_cluster, score,
                               index_down, num_clusters, stream));
    offsets_scan_kernel<<<1, 256, B * sizeof(int32_t), stream>>>(num_clusters, B, offsets);
    SETOK_LAUNCH_CHECK();
    return SETOK_OK;
  }

  // the fp32-FMA Gram fallback reads fp32 rows: with embedded input those are the features themselves
  SETOK_REQUIRE(!embedded || tensor_gram || feat_dtype == SETOK_F32, SETOK_ERR_UNSUPPORTED,
                "dpc_cluster_embedded: bf16 input needs C %% 8 == 0 and N %% 4 == 0 (C=%d N=%d)", C, N);
  const long long rows = static_cast<long long>(B) * N;
  int grid = static_cast<int>((rows + 7) / 8);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  bf16* oh = tensor_gram ? xh : nullptr;
  bf16* ol = tensor_gram ? xl : nullptr;
  if (feat_dtype == SETOK_F32) posadd_sqnorm_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(feats), pos, x_pos, sqn, B, N, C, oh, ol);
  else if (feat_dtype == SETOK_BF16) posadd_sqnorm_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(feats), pos, x_pos, sqn, B, N, C, oh, ol);
  else return fail(SETOK_ERR_BAD_ARG, "dpc_cluster: bad feature dtype %d", feat_dtype);
  SETOK_LAUNCH_CHECK();

  const float sqrtC = static_cast<float>(std::sqrt(static_cast<double>(C)));
  if (tensor_gram) {
    // G = hi hi^T + hi lo^T + lo hi^T + lo lo^T: four batched tcgen05 GEMMs accumulating in fp32 through the residual input
    const bf16* lhs[4] = {xh, xh, xl, xl};
    const bf16* rhs[4] = {xh, xl, xh, xl};
    for (int t = 0; t < 4; ++t) {
      GemmArgs g{lhs[t], C, rhs[t], C, D, N, SETOK_F32, nullptr, t ? D : nullptr, N, SETOK_F32, SETOK_ACT_NONE, N, N, C, nullptr, 0};
      g.batch = B; g.a_batch_stride = static_cast<int64_t>(N) * C; g.w_batch_stride = static_cast<int64_t>(N) * C;
      g.d_batch_stride = static_cast<int64_t>(N) * N; g.r_batch_stride = static_cast<int64_t>(N) * N;
      SETOK_TRY(launch_gemm(g, stream));
    }
    const long long total = static_cast<long long>(B) * N * N;
    long long blocks = (total + 255) / 256;
    if (blocks > num_sms() * 16LL) blocks = num_sms() * 16LL;
    gram_to_dist_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(D, sqn, N, sqrtC, total);
    SETOK_LAUNCH_CHECK();
  } else {
    dim3 ggrid(ceil_div(N, GT), ceil_div(N, GT), B);
    gram_dist_kernel<<<ggrid, 256, 0, stream>>>(embedded ? static_cast<const float*>(feats) : x_pos, sqn, D, N, C, sqrtC);
    SETOK_LAUNCH_CHECK();
  }

  if (N <= 256) dpc_select_kernel<8><<<B, SEL_THREADS, 0, stream>>>(D, noise, token_mask, N, k, threshold, min_cluster_num, idx_cluster, score, index_down, num_clusters);
  else if (N <= 576) dpc_select_kernel<18><<<B, SEL_THREADS, 0, stream>>>(D, noise, token_mask, N, k, threshold, min_cluster_num, idx_cluster, score, index_down, num_clusters);
  else dpc_select_kernel<32><<<B, SEL_THREADS, 0, stream>>>(D, noise, token_mask, N, k, threshold, min_cluster_num, idx_cluster, score, index_down, num_clusters);
  SETOK_LAUNCH_CHECK();

  offsets_scan_kernel<<<1, 256, B * sizeof(int32_t), stream>>>(num_clusters, B, offsets);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}
}  // namespace

extern "C" int setok_dpc_cluster_pos(const void* feats, int feat_dtype, const float* pos_table, const float* noise,
                                     const float* token_mask, int B, int h, int w, int C, int k, float threshold,
                                     int min_cluster_num, float* x_pos, int64_t* idx_cluster, float* score, int64_t* index_down,
                                     int32_t* num_clusters, int32_t* offsets, void* workspace, size_t workspace_bytes,
                                     setok_stream_t stream) {
  return dpc_cluster_impl(feats, feat_dtype, pos_table, false, noise, token_mask, B, h, w, C, k, threshold, min_cluster_num, x_pos,
                          idx_cluster, score, index_down, num_clusters, offsets, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int setok_dpc_cluster_embedded(const void* x_pos_in, int dtype, const float* noise, const float* token_mask, int B, int N, int C,
                                          int k, float threshold, int min_cluster_num, int64_t* idx_cluster, float* score,
                                          int64_t* index_down, int32_t* num_clusters, int32_t* offsets, void* workspace,
                                          size_t workspace_bytes, setok_stream_t stream) {
  return dpc_cluster_impl(x_pos_in, dtype, nullptr, true, noise, token_mask, B, N, 1, C, k, threshold, min_cluster_num, nullptr,
                          idx_cluster, score, index_down, num_clusters, offsets, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int setok_dpc_cluster(const void* feats, int feat_dtype, const float* noise, const float* token_mask, int B, int h, int w,
                                 int C, int k, float threshold, int min_cluster_num, float* x_pos, int64_t* idx_cluster,
                                 float* score, int64_t* index_down, int32_t* num_clusters, int32_t* offsets, void* workspace,
                                 size_t workspace_bytes, setok_stream_t stream) {
  return setok_dpc_cluster_pos(feats, feat_dtype, nullptr, noise, token_mask, B, h, w, C, k, threshold, min_cluster_num, x_pos,
                               idx_cluster, score, index_down, num_clusters, offsets, workspace, workspace_bytes, stream);
}
