// rowops.cu — bandwidth-bound row kernels around the GEMMs: LayerNorm, im2col for the patch embedding,
// CLS rows, feature selection, cluster sort / gather, segment mean.  All are coalesced, vectorised
// (16-byte accesses where the layout allows) and sized in whole waves of the 148 SMs by grid-stride.
#include "common.cuh"
#include "rowops.cuh"

namespace setok {
namespace {

template <class T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(bf16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
};

__device__ __forceinline__ int live_rows(int rows, const int32_t* m_dev) {
  if (m_dev == nullptr) return rows;
  const int m = *m_dev;
  return m < rows ? (m < 0 ? 0 : m) : rows;
}

int g_im2col_rows = 1;   // 0: the element-wise im2col kernels for every patch size (setok_debug_set_im2col_rows; tests compare the two)

// ---- LayerNorm: one warp per row, row cached in registers when C <= 1024 -----------------------
constexpr int LN_MAXV = 8;   // float4 vectors per lane held in registers (C <= 1024)

template <class TI, class TO>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ in, TO* __restrict__ out,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, int rows, int C, const int32_t* __restrict__ gather,
                                                        const int32_t* __restrict__ m_dev) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const int n = live_rows(rows, m_dev);
  const int nvec = C >> 2;
  const bool cached = nvec <= LN_MAXV * 32;
  const float invC = 1.0f / static_cast<float>(C);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
    const long long src = gather ? gather[r] : r;
    const TI* x = in + src * C;
    float4 v[LN_MAXV];
    float s = 0.f;
    if (cached) {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) { v[i] = Vec4<TI>::load(x + vi * 4); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
      }
    } else {
      for (int vi = lane; vi < nvec; vi += 32) { const float4 t = Vec4<TI>::load(x + vi * 4); s += (t.x + t.y) + (t.z + t.w); }
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
    if (cached) {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
          q += (a * a + b * b) + (c * c + d * d);
        }
      }
    } else {
      for (int vi = lane; vi < nvec; vi += 32) {
        const float4 t = Vec4<TI>::load(x + vi * 4);
        const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invC + eps);
    TO* y = out + static_cast<long long>(r) * C;
    if (cached) {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + vi);
          const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + vi);
          float4 o;
          o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
          o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
          Vec4<TO>::store(y + vi * 4, o);
        }
      }
    } else {
      for (int vi = lane; vi < nvec; vi += 32) {
        const float4 t = Vec4<TI>::load(x + vi * 4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + vi);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + vi);
        float4 o;
        o.x = (t.x - mean) * rstd * g.x + b.x; o.y = (t.y - mean) * rstd * g.y + b.y;
        o.z = (t.z - mean) * rstd * g.z + b.z; o.w = (t.w - mean) * rstd * g.w + b.w;
        Vec4<TO>::store(y + vi * 4, o);
      }
    }
  }
}

// ---- LayerNorm fold, first record (SETOK_VIT_LN_FOLD): xhat = bf16((x - mu) * rho) with the row's exact statistics, record
// {c = mu, r = rho, (s1, s2) = (0, C var), 0 ...}: the consuming GEMM's epilogue then finishes LN(x) = (rho' / rho) * xhat ----
__global__ void __launch_bounds__(256) ln_fold_init_kernel(const float* __restrict__ in, bf16* __restrict__ xhat, float* __restrict__ rec,
                                                           float eps, int rows, int C, int ns) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const float invC = 1.0f / static_cast<float>(C);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* x = in + static_cast<long long>(r) * C;
    float s = 0.f;
    for (int vi = lane; vi < nvec; vi += 32) { const float4 t = Vec4<float>::load(x + vi * 4); s += (t.x + t.y) + (t.z + t.w); }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
    for (int vi = lane; vi < nvec; vi += 32) {
      const float4 t = Vec4<float>::load(x + vi * 4);
      const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    q = warp_sum(q);
    const float rho = rsqrtf(q * invC + eps);
    bf16* y = xhat + static_cast<long long>(r) * C;
    for (int vi = lane; vi < nvec; vi += 32) {
      const float4 t = Vec4<float>::load(x + vi * 4);
      Vec4<bf16>::store(y + vi * 4, make_float4((t.x - mean) * rho, (t.y - mean) * rho, (t.z - mean) * rho, (t.w - mean) * rho));
    }
    float2* o = reinterpret_cast<float2*>(rec + static_cast<long long>(r) * (2 + 2 * ns));
    for (int i = lane; i <= ns; i += 32) o[i] = i == 0 ? make_float2(mean, rho) : (i == 1 ? make_float2(0.f, q) : make_float2(0.f, 0.f));
  }
}

// pre_layrnorm and the first fold record in one pass (C <= 1024: the row stays in registers): x = LN(emb) gamma + beta is written
// as the f32 stream and, from the same registers, its own statistics give xhat and the record -- the stream is not re-read.
__global__ void __launch_bounds__(256) preln_fold_init_kernel(const float* __restrict__ in, float* __restrict__ xout, bf16* __restrict__ xhat,
                                                              float* __restrict__ rec, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int rows, int C, int ns) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const float invC = 1.0f / static_cast<float>(C);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* x = in + static_cast<long long>(r) * C;
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) { v[i] = Vec4<float>::load(x + vi * 4); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invC + eps);        // exactly layernorm_kernel's arithmetic: same stream bits
    float* xo = xout + static_cast<long long>(r) * C;
    s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + vi);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + vi);
        v[i].x = (v[i].x - mean) * rstd * g.x + b.x; v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
        v[i].z = (v[i].z - mean) * rstd * g.z + b.z; v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
        Vec4<float>::store(xo + vi * 4, v[i]);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean2 = warp_sum(s) * invC;
    q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float a = v[i].x - mean2, b = v[i].y - mean2, c = v[i].z - mean2, d = v[i].w - mean2;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    q = warp_sum(q);
    const float rho = rsqrtf(q * invC + eps);
    bf16* y = xhat + static_cast<long long>(r) * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) Vec4<bf16>::store(y + vi * 4, make_float4((v[i].x - mean2) * rho, (v[i].y - mean2) * rho, (v[i].z - mean2) * rho, (v[i].w - mean2) * rho));
    }
    float2* o = reinterpret_cast<float2*>(rec + static_cast<long long>(r) * (2 + 2 * ns));
    for (int i = lane; i <= ns; i += 32) o[i] = i == 0 ? make_float2(mean2, rho) : (i == 1 ? make_float2(0.f, q) : make_float2(0.f, 0.f));
  }
}

// ---- patch embedding im2col: images [B,3,H,W] -> A [B*P, Kp] bf16, columns (c, ky, kx), zero pad ----
// split: the row is [hi | lo | hi] (3 * Kp columns) with x = hi + lo (bf16 + bf16): against the weight row
// [w_hi | w_hi | w_lo] one GEMM with K = 3 Kp evaluates x.w = hi.w_hi + lo.w_hi + hi.w_lo, i.e. the patch embedding to
// ~2^-17 instead of the 2^-9 of bf16-rounded pixels and weights (the embedding's rounding error is carried unchanged down
// the residual stream, so it would otherwise set the floor of the whole tower's error)
__device__ __forceinline__ void store_patch_pair(bf16* __restrict__ A, long long row, int Kp, int cp, float v0, float v1, int split) {
  const uint32_t hi = pack_bf16x2(v0, v1);
  if (!split) { *reinterpret_cast<uint32_t*>(A + row * Kp + cp * 2) = hi; return; }
  const float2 hf = unpack_bf16x2(hi);
  const uint32_t lo = pack_bf16x2(v0 - hf.x, v1 - hf.y);
  bf16* r = A + row * (3LL * Kp) + cp * 2;
  *reinterpret_cast<uint32_t*>(r) = hi;
  *reinterpret_cast<uint32_t*>(r + Kp) = lo;
  *reinterpret_cast<uint32_t*>(r + 2 * Kp) = hi;
}
template <class TI>
__global__ void __launch_bounds__(256) im2col_kernel(const TI* __restrict__ img, bf16* __restrict__ A, int B, int H, int W,
                                                     int patch, int Kp, int split) {
  const int gw = W / patch, gh = H / patch;
  const int K = 3 * patch * patch;
  const long long total = static_cast<long long>(B) * gh * gw * (Kp / 2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cp = static_cast<int>(i % (Kp / 2));
    const long long row = i / (Kp / 2);
    const int px = static_cast<int>(row % gw);
    const int py = static_cast<int>((row / gw) % gh);
    const int b = static_cast<int>(row / (static_cast<long long>(gw) * gh));
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = cp * 2 + e;
      if (col < K) {
        const int c = col / (patch * patch);
        const int rem = col % (patch * patch);
        const int ky = rem / patch, kx = rem % patch;
        v[e] = to_f32<TI>(img[((static_cast<long long>(b) * 3 + c) * H + (py * patch + ky)) * W + (px * patch + kx)]);
      } else {
        v[e] = 0.f;
      }
    }
    store_patch_pair(A, row, Kp, cp, v[0], v[1], split);
  }
}

// uint8 pixels: CLIPImageProcessor's rescale + normalize (transformers 4.46.3, numpy float32 arithmetic) fused into the im2col
//   x = lut[u8] (= float32(float64(u8) * rescale_factor), built on the host);  y = (x - mean[c]) / std[c], rounded separately
__global__ void __launch_bounds__(256) im2col_u8_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ A, int B, int H, int W, int patch,
                                                        int Kp, int split, const setok_u8_norm nrm) {
  __shared__ float lut[256];
  lut[threadIdx.x] = nrm.lut[threadIdx.x];
  __syncthreads();
  const int gw = W / patch, gh = H / patch;
  const int K = 3 * patch * patch;
  const long long total = static_cast<long long>(B) * gh * gw * (Kp / 2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cp = static_cast<int>(i % (Kp / 2));
    const long long row = i / (Kp / 2);
    const int px = static_cast<int>(row % gw);
    const int py = static_cast<int>((row / gw) % gh);
    const int b = static_cast<int>(row / (static_cast<long long>(gw) * gh));
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = cp * 2 + e;
      v[e] = 0.f;
      if (col < K) {
        const int c = col / (patch * patch);
        const int rem = col % (patch * patch);
        const int ky = rem / patch, kx = rem % patch;
        const uint8_t u = img[((static_cast<long long>(b) * 3 + c) * H + (py * patch + ky)) * W + (px * patch + kx)];
        v[e] = __fdiv_rn(__fsub_rn(lut[u], nrm.mean[c]), nrm.std[c]);
      }
    }
    store_patch_pair(A, row, Kp, cp, v[0], v[1], split);
  }
}

// Vector form for the patch sizes the path runs (14, 16; compile-time, so the (c, ky, kx) split of a column index is constant
// multiplies): one work item = 8 consecutive columns of one patch row of A -> one aligned 16-byte store per segment
// ([hi | lo | hi] with `split`), a warp writes 512 contiguous bytes per instruction; the 8 pixels come from at most two short
// runs of the image (L1-resident across neighbouring items).  U8: uint8 pixels with the processor's rescale / normalize (as
// im2col_u8_kernel).  Values identical to the element-wise kernels.
template <class TI, bool U8, int PATCH>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const TI* __restrict__ img, bf16* __restrict__ A, int B, int H, int W,
                                                          int Kp, int split, const __grid_constant__ setok_u8_norm nrm) {
  __shared__ float lut[U8 ? 256 : 1];
  if constexpr (U8) {
    lut[threadIdx.x] = nrm.lut[threadIdx.x];
    __syncthreads();
  }
  const int gw = W / PATCH, gh = H / PATCH;
  constexpr int PP = PATCH * PATCH, K = 3 * PP;
  const int per_row = Kp >> 3;
  const long long total = static_cast<long long>(B) * gh * gw * per_row;
  const long long ldA = split ? 3LL * Kp : Kp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % per_row);
    const long long row = i / per_row;
    const int px = static_cast<int>(row % gw);
    const int py = static_cast<int>((row / gw) % gh);
    const int b = static_cast<int>(row / (static_cast<long long>(gw) * gh));
    const TI* src = img + (static_cast<long long>(b) * 3 * H + py * PATCH) * W + px * PATCH;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = 8 * j + e;
      v[e] = 0.f;
      if (col < K) {
        const int c = col / PP, rem = col % PP;
        const int ky = rem / PATCH, kx = rem % PATCH;
        const TI* q = src + (static_cast<long long>(c) * H + ky) * W + kx;
        if constexpr (U8) v[e] = __fdiv_rn(__fsub_rn(lut[*q], nrm.mean[c]), nrm.std[c]);
        else v[e] = to_f32<TI>(*q);
      }
    }
    uint4 hi;
    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
    bf16* dst = A + row * ldA + 8 * j;
    *reinterpret_cast<uint4*>(dst) = hi;
    if (split) {
      const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y), h2 = unpack_bf16x2(hi.z), h3 = unpack_bf16x2(hi.w);
      uint4 lo;
      lo.x = pack_bf16x2(v[0] - h0.x, v[1] - h0.y); lo.y = pack_bf16x2(v[2] - h1.x, v[3] - h1.y);
      lo.z = pack_bf16x2(v[4] - h2.x, v[5] - h2.y); lo.w = pack_bf16x2(v[6] - h3.x, v[7] - h3.y);
      *reinterpret_cast<uint4*>(dst + Kp) = lo;
      *reinterpret_cast<uint4*>(dst + 2 * Kp) = hi;
    }
  }
}

// emb[b*T + 0, :] = cls + pos[0]   (fp32)
__global__ void cls_rows_kernel(float* __restrict__ emb, const float* __restrict__ cls, const float* __restrict__ pos,
                                int B, int T, int C) {
  const int total = B * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / C, c = i % C;
    emb[static_cast<long long>(b) * T * C + c] = cls[c] + pos[c];
  }
}

// features[b, t', :] = x[b*T + skip + t', :] (+ pos[t', :])  (dtype conversion bf16 -> out).  With `pos` the output is
// the position-embedded tensor of tokenizer.py:168 (feature_select and the add fused in one pass).
template <class TI, class TO>
__global__ void __launch_bounds__(256) select_rows_kernel(const TI* __restrict__ x, TO* __restrict__ out, int B, int T, int skip, int C,
                                                          const float* __restrict__ pos) {
  const int To = T - skip;
  const int nvec = C >> 2;
  const long long total = static_cast<long long>(B) * To * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vi = static_cast<int>(i % nvec);
    const long long r = i / nvec;
    const int t = static_cast<int>(r % To);
    const long long b = r / To;
    float4 v = Vec4<TI>::load(x + ((b * T + skip + t) * C + vi * 4));
    if (pos != nullptr) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * C + vi * 4));
      v.x = __fadd_rn(v.x, q.x); v.y = __fadd_rn(v.y, q.y); v.z = __fadd_rn(v.z, q.z); v.w = __fadd_rn(v.w, q.w);
    }
    Vec4<TO>::store(out + (r * C + vi * 4), v);
  }
}

// ---- cluster bookkeeping -------------------------------------------------------------------------
// One CTA per image: stable counting sort of the image's N tokens by cluster label.
//   perm[b*N + p]    = global source row (b*N + t) of the token at sorted position p
//   row_seg[b*N + p] = global cluster id (offsets[b] + label)
//   seg_off[g]       = first sorted row of global cluster g; seg_off[offsets[B]] = B*N
__global__ void __launch_bounds__(256) sort_by_cluster_kernel(const int64_t* __restrict__ idx_cluster,
                                                              const int32_t* __restrict__ num_clusters,
                                                              const int32_t* __restrict__ offsets, int B, int N,
                                                              int32_t* __restrict__ perm, int32_t* __restrict__ row_seg,
                                                              int32_t* __restrict__ seg_off) {
  extern __shared__ int32_t sm[];
  int32_t* lab = sm;            // [N]
  int32_t* cnt = sm + N;        // [N]  counts, then exclusive starts
  const int b = blockIdx.x;
  const int K = num_clusters[b];
  const int goff = offsets[b];
  for (int t = threadIdx.x; t < N; t += blockDim.x) { lab[t] = static_cast<int32_t>(idx_cluster[static_cast<long long>(b) * N + t]); cnt[t] = 0; }
  __syncthreads();
  for (int t = threadIdx.x; t < N; t += blockDim.x) atomicAdd(&cnt[lab[t]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int c = 0; c < K; ++c) { const int n = cnt[c]; cnt[c] = run; run += n; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    int pos = cnt[c];
    seg_off[goff + c] = b * N + pos;
    for (int t = 0; t < N; ++t) {
      if (lab[t] == c) { perm[b * N + pos] = b * N + t; row_seg[b * N + pos] = goff + c; ++pos; }
    }
  }
  if (b == B - 1 && threadIdx.x == 0) seg_off[goff + K] = B * N;
}

// out[r, :] = in[perm[r], :]   (fp32 rows)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          const int32_t* __restrict__ perm, int rows, int C) {
  const int nvec = C >> 2;
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vi = static_cast<int>(i % nvec);
    const long long r = i / nvec;
    reinterpret_cast<float4*>(out)[r * nvec + vi] = reinterpret_cast<const float4*>(in)[static_cast<long long>(perm[r]) * nvec + vi];
  }
}

// out[g, :] = mean over rows [seg_off[g], seg_off[g+1]) of x   (fp32), g < *n_seg_dev
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ seg_off,
                                                           const int32_t* __restrict__ n_seg_dev, int cap, int C,
                                                           float* __restrict__ out, float* __restrict__ out2) {
  const int nvec = C >> 2;
  const int nseg = live_rows(cap, n_seg_dev);
  for (int g = blockIdx.x; g < nseg; g += gridDim.x) {
    const int r0 = seg_off[g], r1 = seg_off[g + 1];
    const float inv = 1.0f / static_cast<float>(r1 - r0);
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = r0; r < r1; ++r) {
        const float4 t = reinterpret_cast<const float4*>(x)[static_cast<long long>(r) * nvec + vi];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
      reinterpret_cast<float4*>(out)[static_cast<long long>(g) * nvec + vi] = s;
      if (out2) reinterpret_cast<float4*>(out2)[static_cast<long long>(g) * nvec + vi] = s;
    }
  }
}

// row_seg[r] = image index for the packed token rows; used by the inter-cluster encoder
__global__ void __launch_bounds__(256) image_segments_kernel(const int32_t* __restrict__ offsets, int B, int32_t* __restrict__ row_seg) {
  const int b = blockIdx.x;
  for (int r = offsets[b] + threadIdx.x; r < offsets[b + 1]; r += blockDim.x) row_seg[r] = b;
}

template <class TI, class TO>
__global__ void __launch_bounds__(256) convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    Vec4<TO>::store(out + i * 4, Vec4<TI>::load(in + i * 4));
}

// out[r, :] = (TO) act(in[r, :]) for r < live rows; act: 0 none, 2 GELU(erf)
template <class TI, class TO>
__global__ void __launch_bounds__(256) convert_rows_kernel(const TI* __restrict__ in, TO* __restrict__ out, int rows, int C, int act,
                                                           const int32_t* __restrict__ m_dev) {
  const int nvec = C >> 2;
  const long long total = static_cast<long long>(live_rows(rows, m_dev)) * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = Vec4<TI>::load(in + i * 4);
    if (act == SETOK_ACT_GELU_ERF) { v.x = act_gelu_erf(v.x); v.y = act_gelu_erf(v.y); v.z = act_gelu_erf(v.z); v.w = act_gelu_erf(v.w); }
    Vec4<TO>::store(out + i * 4, v);
  }
}

// Dense cluster attention, middle step: P[r, :] = softmax(scale * S[r, keys of r's cluster]) and 0 elsewhere.
// Rows are sorted by cluster, so the admissible keys of row r are the contiguous range of its segment.  One warp per row.
__global__ void __launch_bounds__(256) masked_softmax_kernel(const float* __restrict__ S, bf16* __restrict__ P,
                                                             const int32_t* __restrict__ seg_off, const int32_t* __restrict__ row_seg,
                                                             int rows, int N, int ldS, int ldP, float scale_log2) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int img0 = (r / N) * N;
    const int g = row_seg[r];
    const int k0 = seg_off[g] - img0, k1 = seg_off[g + 1] - img0;
    const float* s = S + static_cast<long long>(r) * ldS;
    bf16* p = P + static_cast<long long>(r) * ldP;
    float mx = -INFINITY;
    for (int j = k0 + lane; j < k1; j += 32) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx) * scale_log2;
    float sum = 0.f;
    for (int j = k0 + lane; j < k1; j += 32) sum += exp2f(fmaf(s[j], scale_log2, -mx));
    const float inv = 1.0f / warp_sum(sum);
    for (int j = lane; j < ldP; j += 32) {
      float v = 0.f;
      if (j >= k0 && j < k1) v = exp2f(fmaf(s[j], scale_log2, -mx)) * inv;
      p[j] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void zero_tail_rows_kernel(bf16* __restrict__ buf, long long ld, const int32_t* __restrict__ n_live_dev, int n_rows, int cap, int cols) {
  const int n0 = min(max(*n_live_dev, 0), cap);
  const int n1 = min(n0 + n_rows, cap);
  const int total = (n1 - n0) * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    buf[static_cast<long long>(n0 + i / cols) * ld + i % cols] = __float2bfloat16_rn(0.f);
}

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

template <class TI, bool U8>
bool launch_im2col_rows(const TI* img, bf16* A, int B, int H, int W, int patch, int Kp, int split, const setok_u8_norm* nrm_host, cudaStream_t stream) {
  if ((patch != 14 && patch != 16) || Kp % 8 != 0 || (reinterpret_cast<uintptr_t>(A) % 16) != 0) return false;   // 16-byte stores
  const long long total = static_cast<long long>(B) * (H / patch) * (W / patch) * (Kp / 8);
  const int grid = grid_for(total, 256, 16);
  static const setok_u8_norm none{};
  const setok_u8_norm& nrm = nrm_host ? *nrm_host : none;
  if (patch == 14) im2col_rows_kernel<TI, U8, 14><<<grid, 256, 0, stream>>>(img, A, B, H, W, Kp, split, nrm);
  else im2col_rows_kernel<TI, U8, 16><<<grid, 256, 0, stream>>>(img, A, B, H, W, Kp, split, nrm);
  return true;
}


}  // namespace

int launch_layernorm(const void* in, int in_dtype, void* out, int out_dtype, const float* gamma, const float* beta, float eps,
                     int rows, int C, const int32_t* gather, const int32_t* m_dev, cudaStream_t stream) {
  SETOK_REQUIRE(in && out && gamma && beta, SETOK_ERR_BAD_ARG, "layernorm: null pointer");
  SETOK_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, SETOK_ERR_UNSUPPORTED, "layernorm: rows=%d C=%d (C must be a multiple of 4)", rows, C);
  SETOK_REQUIRE(aligned16(in) && aligned16(out) && aligned16(gamma) && aligned16(beta), SETOK_ERR_BAD_ARG, "layernorm: buffers must be 16-byte aligned");
  const int wpb = 8;
  int grid = ceil_div(rows, wpb);
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
#define LN_CASE(TI, TO) SETOK_CUDA_OK(launch_pdl(layernorm_kernel<TI, TO>, dim3(grid), dim3(wpb * 32), 0, stream, static_cast<const TI*>(in), static_cast<TO*>(out), gamma, beta, eps, rows, C, gather, m_dev))
  if (in_dtype == SETOK_F32 && out_dtype == SETOK_F32) LN_CASE(float, float);
  else if (in_dtype == SETOK_F32 && out_dtype == SETOK_BF16) LN_CASE(float, bf16);
  else if (in_dtype == SETOK_BF16 && out_dtype == SETOK_BF16) LN_CASE(bf16, bf16);
  else if (in_dtype == SETOK_BF16 && out_dtype == SETOK_F32) LN_CASE(bf16, float);
  else return fail(SETOK_ERR_BAD_ARG, "layernorm: bad dtypes %d -> %d", in_dtype, out_dtype);
#undef LN_CASE
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_ln_fold_init(const float* x, void* xhat_bf16, float* rec, float eps, int rows, int C, cudaStream_t stream) {
  SETOK_REQUIRE(C % 4 == 0 && aligned16(x) && aligned16(xhat_bf16) && aligned16(rec), SETOK_ERR_BAD_ARG, "ln_fold_init: C %% 4 / alignment");
  int grid = ceil_div(rows, 8);
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  SETOK_CUDA_OK(launch_pdl(ln_fold_init_kernel, dim3(grid), dim3(256), 0, stream, x, static_cast<bf16*>(xhat_bf16), rec, eps, rows, C, ceil_div(C, 128)));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

// x = LayerNorm(emb) (f32 -> f32) followed by launch_ln_fold_init(x), as one pass when the row fits the register cache
int launch_preln_fold_init(const float* emb, float* x, void* xhat_bf16, float* rec, const float* gamma, const float* beta, float eps, int rows,
                           int C, cudaStream_t stream) {
  if (C % 4 != 0 || (C >> 2) > LN_MAXV * 32) {
    SETOK_TRY(launch_layernorm(emb, SETOK_F32, x, SETOK_F32, gamma, beta, eps, rows, C, nullptr, nullptr, stream));
    return launch_ln_fold_init(x, xhat_bf16, rec, eps, rows, C, stream);
  }
  SETOK_REQUIRE(aligned16(emb) && aligned16(x) && aligned16(xhat_bf16) && aligned16(rec) && aligned16(gamma) && aligned16(beta), SETOK_ERR_BAD_ARG,
                "preln_fold_init: buffers must be 16-byte aligned");
  int grid = ceil_div(rows, 8);
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  SETOK_CUDA_OK(launch_pdl(preln_fold_init_kernel, dim3(grid), dim3(256), 0, stream, emb, x, static_cast<bf16*>(xhat_bf16), rec, gamma, beta, eps, rows, C,
                           ceil_div(C, 128)));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_masked_softmax(const float* S, void* P, const int32_t* seg_off, const int32_t* row_seg, int rows, int N, int ldS, int ldP,
                          float scale, cudaStream_t stream) {
  int grid = ceil_div(rows, 8);
  const int cap = num_sms() * 16;
  if (grid > cap) grid = cap;
  masked_softmax_kernel<<<grid, 256, 0, stream>>>(S, static_cast<bf16*>(P), seg_off, row_seg, rows, N, ldS, ldP, scale * 1.4426950408889634f);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_im2col(const void* images, int image_dtype, void* A, int B, int H, int W, int patch, int Kp, int split, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * (H / patch) * (W / patch) * (Kp / 2);
  const int grid = grid_for(total, 256, 16);
  if (g_im2col_rows && image_dtype == SETOK_F32 && launch_im2col_rows<float, false>(static_cast<const float*>(images), static_cast<bf16*>(A), B, H, W, patch, Kp, split, nullptr, stream)) {
    SETOK_LAUNCH_CHECK();
    return SETOK_OK;
  }
  if (g_im2col_rows && image_dtype == SETOK_BF16 && launch_im2col_rows<bf16, false>(static_cast<const bf16*>(images), static_cast<bf16*>(A), B, H, W, patch, Kp, split, nullptr, stream)) {
    SETOK_LAUNCH_CHECK();
    return SETOK_OK;
  }
  if (image_dtype == SETOK_F32) im2col_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(images), static_cast<bf16*>(A), B, H, W, patch, Kp, split);
  else if (image_dtype == SETOK_BF16) im2col_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(images), static_cast<bf16*>(A), B, H, W, patch, Kp, split);
  else return fail(SETOK_ERR_BAD_ARG, "im2col: bad image dtype %d", image_dtype);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_im2col_u8(const uint8_t* images, const setok_u8_norm* norm, void* A, int B, int H, int W, int patch, int Kp, int split, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * (H / patch) * (W / patch) * (Kp / 2);
  if (g_im2col_rows && launch_im2col_rows<uint8_t, true>(images, static_cast<bf16*>(A), B, H, W, patch, Kp, split, norm, stream)) {
    SETOK_LAUNCH_CHECK();
    return SETOK_OK;
  }
  im2col_u8_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(images, static_cast<bf16*>(A), B, H, W, patch, Kp, split, *norm);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_cls_rows(float* emb, const float* cls, const float* pos, int B, int T, int C, cudaStream_t stream) {
  cls_rows_kernel<<<grid_for(static_cast<long long>(B) * C, 256), 256, 0, stream>>>(emb, cls, pos, B, T, C);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

// idx[r] = r % Q: gather indices that replicate a (Q, C) table for every image of a batch
__global__ void iota_mod_kernel(int32_t* __restrict__ idx, int rows, int Q) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) idx[r] = r % Q;
}

// out[r, :] (bf16) = in[r, :] (f32) + pos[r % Q, :]
__global__ void __launch_bounds__(256) add_pos_rows_kernel(const float* __restrict__ in, const float* __restrict__ pos, bf16* __restrict__ out,
                                                           int rows, int Q, int C) {
  const int nvec = C >> 2;
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vi = static_cast<int>(i % nvec);
    const long long r = i / nvec;
    float4 v = reinterpret_cast<const float4*>(in)[i];
    const float4 q = __ldg(reinterpret_cast<const float4*>(pos + (r % Q) * C + vi * 4));
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    Vec4<bf16>::store(out + (r * C + vi * 4), v);
  }
}

int launch_iota_mod(int32_t* idx, int rows, int Q, cudaStream_t stream) {
  iota_mod_kernel<<<grid_for(rows, 256, 4), 256, 0, stream>>>(idx, rows, Q);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_add_pos_rows(const float* in, const float* pos, void* out_bf16, int rows, int Q, int C, cudaStream_t stream) {
  SETOK_REQUIRE(C % 4 == 0, SETOK_ERR_UNSUPPORTED, "add_pos_rows: C (%d) must be a multiple of 4", C);
  add_pos_rows_kernel<<<grid_for(static_cast<long long>(rows) * (C / 4), 256, 16), 256, 0, stream>>>(in, pos, static_cast<bf16*>(out_bf16), rows, Q, C);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_select_rows(const void* x, int x_dtype, void* out, int out_dtype, int B, int T, int skip, int C, const float* pos, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * (T - skip) * (C / 4);
  const int grid = grid_for(total, 256, 16);
  SETOK_REQUIRE((x_dtype == SETOK_F32 || x_dtype == SETOK_BF16) && (out_dtype == SETOK_F32 || out_dtype == SETOK_BF16), SETOK_ERR_BAD_ARG,
                "select_rows: bad dtype %d -> %d", x_dtype, out_dtype);
#define SETOK_SEL(TI, TO) select_rows_kernel<TI, TO><<<grid, 256, 0, stream>>>(static_cast<const TI*>(x), static_cast<TO*>(out), B, T, skip, C, pos)
  if (x_dtype == SETOK_BF16) { if (out_dtype == SETOK_F32) SETOK_SEL(bf16, float); else SETOK_SEL(bf16, bf16); }
  else { if (out_dtype == SETOK_F32) SETOK_SEL(float, float); else SETOK_SEL(float, bf16); }
#undef SETOK_SEL
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_zero_tail_rows(void* buf_bf16, long long ld, const int32_t* n_live_dev, int n_rows, int cap, int cols, cudaStream_t stream) {
  zero_tail_rows_kernel<<<grid_for(static_cast<long long>(n_rows) * cols, 256), 256, 0, stream>>>(static_cast<bf16*>(buf_bf16), ld, n_live_dev, n_rows, cap, cols);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_sort_by_cluster(const int64_t* idx_cluster, const int32_t* num_clusters, const int32_t* offsets, int B, int N,
                           int32_t* perm, int32_t* row_seg, int32_t* seg_off, cudaStream_t stream) {
  sort_by_cluster_kernel<<<B, 256, 2 * N * sizeof(int32_t), stream>>>(idx_cluster, num_clusters, offsets, B, N, perm, row_seg, seg_off);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_gather_rows(const float* in, float* out, const int32_t* perm, int rows, int C, cudaStream_t stream) {
  gather_rows_kernel<<<grid_for(static_cast<long long>(rows) * (C / 4), 256, 16), 256, 0, stream>>>(in, out, perm, rows, C);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_segment_mean(const float* x, const int32_t* seg_off, const int32_t* n_seg_dev, int cap, int C, float* out,
                        float* out2, cudaStream_t stream) {
  int grid = cap < num_sms() * 8 ? cap : num_sms() * 8;
  segment_mean_kernel<<<grid, 256, 0, stream>>>(x, seg_off, n_seg_dev, cap, C, out, out2);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_image_segments(const int32_t* offsets, int B, int32_t* row_seg, cudaStream_t stream) {
  image_segments_kernel<<<B, 256, 0, stream>>>(offsets, B, row_seg);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_convert(const void* in, int in_dtype, void* out, int out_dtype, long long n, cudaStream_t stream) {
  SETOK_REQUIRE(n % 4 == 0, SETOK_ERR_UNSUPPORTED, "convert: element count must be a multiple of 4");
  const int grid = grid_for(n / 4, 256, 16);
  if (in_dtype == SETOK_F32 && out_dtype == SETOK_BF16) convert_kernel<float, bf16><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<bf16*>(out), n / 4);
  else if (in_dtype == SETOK_BF16 && out_dtype == SETOK_F32) convert_kernel<bf16, float><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(in), static_cast<float*>(out), n / 4);
  else return fail(SETOK_ERR_BAD_ARG, "convert: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

int launch_convert_rows(const void* in, int in_dtype, void* out, int out_dtype, int rows, int C, int act, const int32_t* m_dev,
                        cudaStream_t stream) {
  SETOK_REQUIRE(C % 4 == 0, SETOK_ERR_UNSUPPORTED, "convert_rows: C must be a multiple of 4");
  const int grid = grid_for(static_cast<long long>(rows) * (C / 4), 256, 16);
#define CV_CASE(TI, TO) convert_rows_kernel<TI, TO><<<grid, 256, 0, stream>>>(static_cast<const TI*>(in), static_cast<TO*>(out), rows, C, act, m_dev)
  if (in_dtype == SETOK_F32 && out_dtype == SETOK_BF16) CV_CASE(float, bf16);
  else if (in_dtype == SETOK_BF16 && out_dtype == SETOK_F32) CV_CASE(bf16, float);
  else if (in_dtype == SETOK_BF16 && out_dtype == SETOK_BF16) CV_CASE(bf16, bf16);
  else if (in_dtype == SETOK_F32 && out_dtype == SETOK_F32) CV_CASE(float, float);
  else return fail(SETOK_ERR_BAD_ARG, "convert_rows: bad dtypes %d -> %d", in_dtype, out_dtype);
#undef CV_CASE
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" void setok_debug_set_im2col_rows(int on) { setok::g_im2col_rows = on; }

extern "C" int setok_ln_fold_init(const float* x, void* xhat, float* records, float eps, int rows, int C, setok_stream_t stream) {
  SETOK_REQUIRE(x && xhat && records && rows > 0 && C > 0, SETOK_ERR_BAD_ARG, "ln_fold_init: null pointer or empty shape");
  return setok::launch_ln_fold_init(x, xhat, records, eps, rows, C, static_cast<cudaStream_t>(stream));
}

extern "C" int setok_layernorm(const void* in, int in_dtype, void* out, int out_dtype, const float* gamma, const float* beta,
                               float eps, int rows, int C, const int32_t* gather, const int32_t* m_dev, setok_stream_t stream) {
  return setok::launch_layernorm(in, in_dtype, out, out_dtype, gamma, beta, eps, rows, C, gather, m_dev, static_cast<cudaStream_t>(stream));
}
