// attention.cu — self-attention over packed rows.
//
// Two kernels, chosen by shape (not by backend):
//  * attn_tcgen05_hd64_kernel (attention_tcgen05.cu): uniform-length segments (the ViT: T = N+1 rows per
//    image), head_dim 64, on the tensor cores with S/O in TMEM.
//  * attn_seg_kernel (here): arbitrary ragged segments (clusters of 1..N tokens, images of K_b tokens),
//    head_dim up to 512 (the head uses 2 heads of C/2).  One warp per (row, head), online softmax,
//    16-byte coalesced K/V row reads; fp32 math throughout.  The work is tiny (8*C*sum n_g^2 FLOPs)
//    and irregular, so it stays on CUDA cores by design.
#include "common.cuh"

namespace setok {
namespace {

__device__ __forceinline__ int live_rows_a(int rows, const int32_t* m_dev) {
  if (m_dev == nullptr) return rows;
  const int m = *m_dev;
  return m < rows ? (m < 0 ? 0 : m) : rows;
}

// -------------------------------------------------------------------------------------------------
// ragged segments, any head_dim % 8 == 0, <= 512
// -------------------------------------------------------------------------------------------------
// Queries at q + r*ldq, keys / values at kp / vp + j*ldkv.  Self-attention over a fused qkv buffer passes q = qkv,
// kp = qkv + C, vp = qkv + 2C, ldq = ldkv = 3C.  cross_Q > 0: cross-attention, query row r belongs to image r / cross_Q
// and attends to the packed key rows [seg_off[image], seg_off[image + 1]) (the ragged token batch: no padding mask).
template <int VPL>
__global__ void __launch_bounds__(256) attn_seg_kernel(const bf16* __restrict__ qp, long long ldq, const bf16* __restrict__ kp,
                                                       const bf16* __restrict__ vp, long long ldkv, bf16* __restrict__ out, int rows, int C,
                                                       int heads, float scale, const int32_t* __restrict__ seg_off,
                                                       const int32_t* __restrict__ row_seg, int uniform_T, int cross_Q,
                                                       const int32_t* __restrict__ m_dev) {
  const int lane = threadIdx.x & 31;
  const int n = live_rows_a(rows, m_dev);
  const long long total = static_cast<long long>(n) * heads;
  const int hd = C / heads;
  const int nv = hd >> 3;
  for (long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); wid < total;
       wid += static_cast<long long>(gridDim.x) * (blockDim.x >> 5)) {
    const int r = static_cast<int>(wid / heads);
    const int h = static_cast<int>(wid % heads);
    int s0, s1;
    if (cross_Q > 0) { const int s = r / cross_Q; s0 = seg_off[s]; s1 = seg_off[s + 1]; }
    else if (uniform_T > 0) { s0 = (r / uniform_T) * uniform_T; s1 = s0 + uniform_T; }
    else { const int s = row_seg[r]; s0 = seg_off[s]; s1 = seg_off[s + 1]; }
    float q[VPL][8], o[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
#pragma unroll
      for (int e = 0; e < 8; ++e) { q[i][e] = 0.f; o[i][e] = 0.f; }
      if (vi < nv) {
        const uint4 u = *reinterpret_cast<const uint4*>(qp + r * ldq + h * hd + vi * 8);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        q[i][0] = a.x * scale; q[i][1] = a.y * scale; q[i][2] = b.x * scale; q[i][3] = b.y * scale;
        q[i][4] = c.x * scale; q[i][5] = c.y * scale; q[i][6] = d.x * scale; q[i][7] = d.y * scale;
      }
    }
    float m = -INFINITY, l = 0.f;
    for (int j = s0; j < s1; j += 4) {
      float sc[4];
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int jj = j + u4;
        float part = 0.f;
        if (jj < s1) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int vi = lane + 32 * i;
            if (vi < nv) {
              const uint4 u = *reinterpret_cast<const uint4*>(kp + jj * ldkv + h * hd + vi * 8);
              const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
              part += q[i][0] * a.x + q[i][1] * a.y + q[i][2] * b.x + q[i][3] * b.y + q[i][4] * c.x + q[i][5] * c.y +
                      q[i][6] * d.x + q[i][7] * d.y;
            }
          }
        }
        sc[u4] = part;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) sc[u4] += __shfl_xor_sync(0xffffffffu, sc[u4], off);
      }
      float mnew = m;
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) if (j + u4 < s1) mnew = fmaxf(mnew, sc[u4]);
      const float alpha = expf(m - mnew);
      float p[4], psum = 0.f;
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) { p[u4] = (j + u4 < s1) ? expf(sc[u4] - mnew) : 0.f; psum += p[u4]; }
      l = l * alpha + psum;
      m = mnew;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) o[i][e] *= alpha;
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int jj = j + u4;
        if (jj < s1) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int vi = lane + 32 * i;
            if (vi < nv) {
              const uint4 u = *reinterpret_cast<const uint4*>(vp + jj * ldkv + h * hd + vi * 8);
              const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
              o[i][0] += p[u4] * a.x; o[i][1] += p[u4] * a.y; o[i][2] += p[u4] * b.x; o[i][3] += p[u4] * b.y;
              o[i][4] += p[u4] * c.x; o[i][5] += p[u4] * c.y; o[i][6] += p[u4] * d.x; o[i][7] += p[u4] * d.y;
            }
          }
        }
      }
    }
    const float inv = l > 0.f ? 1.0f / l : 0.f;     // an empty key segment yields zeros
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nv) {
        uint4 u;
        u.x = pack_bf16x2(o[i][0] * inv, o[i][1] * inv); u.y = pack_bf16x2(o[i][2] * inv, o[i][3] * inv);
        u.z = pack_bf16x2(o[i][4] * inv, o[i][5] * inv); u.w = pack_bf16x2(o[i][6] * inv, o[i][7] * inv);
        *reinterpret_cast<uint4*>(out + static_cast<long long>(r) * C + h * hd + vi * 8) = u;
      }
    }
  }
}


// -------------------------------------------------------------------------------------------------
// cross-attention, head_dim 64 (the Q-Former's BERT heads): one thread per query, the image's keys / values stream
// through shared memory as fp32 in chunks of 64 (every lane of a warp reads the same K/V row: 16-byte broadcasts),
// scores in groups of 8 keys, online softmax in the exp2 domain.  Grid (head, image, query block).
// -------------------------------------------------------------------------------------------------
constexpr int XA_HD = 64, XA_KC = 64, XA_THREADS = 192;

__global__ void __launch_bounds__(XA_THREADS) cross_attn_hd64_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv, bf16* __restrict__ out,
                                                                     int Q, int C, float scale_log2, const int32_t* __restrict__ offsets) {
  __shared__ __align__(16) float Ks[XA_KC][XA_HD];
  __shared__ __align__(16) float Vs[XA_KC][XA_HD];
  const int h = blockIdx.x, b = blockIdx.y;
  const int qi = blockIdx.z * XA_THREADS + threadIdx.x;
  const bool live = qi < Q;
  const int s0 = offsets[b], s1 = offsets[b + 1];
  const long long qrow = static_cast<long long>(b) * Q + (live ? qi : 0);
  float qv[XA_HD], o[XA_HD];
#pragma unroll
  for (int c = 0; c < XA_HD / 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(q + qrow * C + h * XA_HD + c * 8);
    const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d2 = unpack_bf16x2(u.w);
    qv[c * 8 + 0] = a.x * scale_log2; qv[c * 8 + 1] = a.y * scale_log2; qv[c * 8 + 2] = b2.x * scale_log2; qv[c * 8 + 3] = b2.y * scale_log2;
    qv[c * 8 + 4] = c2.x * scale_log2; qv[c * 8 + 5] = c2.y * scale_log2; qv[c * 8 + 6] = d2.x * scale_log2; qv[c * 8 + 7] = d2.y * scale_log2;
  }
#pragma unroll
  for (int d = 0; d < XA_HD; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  const long long ldkv = 2LL * C;
  for (int k0 = s0; k0 < s1; k0 += XA_KC) {
    const int nk = min(XA_KC, s1 - k0);
    const int nk8 = (nk + 7) & ~7;
    __syncthreads();
    for (int idx = threadIdx.x; idx < nk8 * (XA_HD / 8); idx += XA_THREADS) {
      const int key = idx / (XA_HD / 8), c = idx % (XA_HD / 8);
      float kf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (key < nk) {
        const bf16* src = kv + (k0 + key) * ldkv + h * XA_HD + c * 8;
        const uint4 uk = *reinterpret_cast<const uint4*>(src);
        const uint4 uv = *reinterpret_cast<const uint4*>(src + C);
        const float2 k0f = unpack_bf16x2(uk.x), k1f = unpack_bf16x2(uk.y), k2f = unpack_bf16x2(uk.z), k3f = unpack_bf16x2(uk.w);
        const float2 v0f = unpack_bf16x2(uv.x), v1f = unpack_bf16x2(uv.y), v2f = unpack_bf16x2(uv.z), v3f = unpack_bf16x2(uv.w);
        kf[0] = k0f.x; kf[1] = k0f.y; kf[2] = k1f.x; kf[3] = k1f.y; kf[4] = k2f.x; kf[5] = k2f.y; kf[6] = k3f.x; kf[7] = k3f.y;
        vf[0] = v0f.x; vf[1] = v0f.y; vf[2] = v1f.x; vf[3] = v1f.y; vf[4] = v2f.x; vf[5] = v2f.y; vf[6] = v3f.x; vf[7] = v3f.y;
      }
      *reinterpret_cast<float4*>(&Ks[key][c * 8]) = make_float4(kf[0], kf[1], kf[2], kf[3]);
      *reinterpret_cast<float4*>(&Ks[key][c * 8 + 4]) = make_float4(kf[4], kf[5], kf[6], kf[7]);
      *reinterpret_cast<float4*>(&Vs[key][c * 8]) = make_float4(vf[0], vf[1], vf[2], vf[3]);
      *reinterpret_cast<float4*>(&Vs[key][c * 8 + 4]) = make_float4(vf[4], vf[5], vf[6], vf[7]);
    }
    __syncthreads();
    for (int kk = 0; kk < nk8; kk += 8) {
      float sc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int d = 0; d < XA_HD; d += 8) {
          const float4 k4 = *reinterpret_cast<const float4*>(&Ks[kk + u][d]);
          const float4 k5 = *reinterpret_cast<const float4*>(&Ks[kk + u][d + 4]);
          a0 = fmaf(qv[d], k4.x, a0); a1 = fmaf(qv[d + 1], k4.y, a1); a0 = fmaf(qv[d + 2], k4.z, a0); a1 = fmaf(qv[d + 3], k4.w, a1);
          a0 = fmaf(qv[d + 4], k5.x, a0); a1 = fmaf(qv[d + 5], k5.y, a1); a0 = fmaf(qv[d + 6], k5.z, a0); a1 = fmaf(qv[d + 7], k5.w, a1);
        }
        sc[u] = kk + u < nk ? a0 + a1 : -INFINITY;
      }
      float mx = sc[0];
#pragma unroll
      for (int u = 1; u < 8; ++u) mx = fmaxf(mx, sc[u]);
      const float m_new = fmaxf(m, mx);
      const float alpha = exp2f(m - m_new);
      float psum = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) { sc[u] = exp2f(sc[u] - m_new); psum += sc[u]; }
      l = l * alpha + psum;
      m = m_new;
#pragma unroll
      for (int d = 0; d < XA_HD; ++d) o[d] *= alpha;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int d = 0; d < XA_HD; d += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(&Vs[kk + u][d]);
          o[d] = fmaf(sc[u], v4.x, o[d]); o[d + 1] = fmaf(sc[u], v4.y, o[d + 1]); o[d + 2] = fmaf(sc[u], v4.z, o[d + 2]); o[d + 3] = fmaf(sc[u], v4.w, o[d + 3]);
        }
      }
    }
  }
  if (live) {
    const float inv = l > 0.f ? 1.0f / l : 0.f;       // an empty key segment yields zeros
    bf16* dst = out + (static_cast<long long>(b) * Q + qi) * C + h * XA_HD;
#pragma unroll
    for (int c = 0; c < XA_HD / 8; ++c) {
      uint4 u;
      u.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv); u.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
      u.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv); u.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
      *reinterpret_cast<uint4*>(dst + c * 8) = u;
    }
  }
}

}  // namespace

int launch_attention(const void* qkv, void* out, int rows, int C, int heads, float scale, const int32_t* seg_off,
                     const int32_t* row_seg, int uniform_T, const int32_t* m_dev, cudaStream_t stream) {
  SETOK_REQUIRE(qkv && out, SETOK_ERR_BAD_ARG, "attention: null pointer");
  SETOK_REQUIRE(rows > 0 && C > 0 && heads > 0 && C % heads == 0, SETOK_ERR_BAD_ARG, "attention: bad shape rows=%d C=%d heads=%d", rows, C, heads);
  SETOK_REQUIRE(aligned16(qkv) && aligned16(out), SETOK_ERR_BAD_ARG, "attention: buffers must be 16-byte aligned");
  const int hd = C / heads;
  SETOK_REQUIRE(hd % 8 == 0 && hd <= 512, SETOK_ERR_UNSUPPORTED, "attention: head_dim %d unsupported (need %%8==0, <=512)", hd);
  if (uniform_T > 0) {
    SETOK_REQUIRE(rows % uniform_T == 0, SETOK_ERR_BAD_ARG, "attention: rows %d not a multiple of uniform_T %d", rows, uniform_T);
  } else {
    SETOK_REQUIRE(seg_off && row_seg, SETOK_ERR_BAD_ARG, "attention: seg_off/row_seg required for ragged segments");
  }
  if (uniform_T > 0 && hd == 64 && m_dev == nullptr) {
    // T <= 257 (the 224^2 tower and shorter): whole score rows in tensor memory; longer sequences: 64-key chunks, online softmax
    if (attention_fullrow_supported(uniform_T)) return launch_attention_fullrow(qkv, out, rows / uniform_T, uniform_T, C, heads, scale, stream);
    return launch_attention_tcgen05(qkv, out, rows / uniform_T, uniform_T, C, heads, scale, stream);
  }
  const long long warps = static_cast<long long>(rows) * heads;
  long long blocks = (warps + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  const bf16* q = static_cast<const bf16*>(qkv);
  if (hd <= 256)
    attn_seg_kernel<1><<<static_cast<int>(blocks), 256, 0, stream>>>(q, 3LL * C, q + C, q + 2 * C, 3LL * C, static_cast<bf16*>(out), rows, C, heads, scale, seg_off, row_seg, uniform_T, 0, m_dev);
  else
    attn_seg_kernel<2><<<static_cast<int>(blocks), 256, 0, stream>>>(q, 3LL * C, q + C, q + 2 * C, 3LL * C, static_cast<bf16*>(out), rows, C, heads, scale, seg_off, row_seg, uniform_T, 0, m_dev);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

// Cross-attention of `rows` = B * Q query rows (q [rows, C]) over the ragged key/value rows kv [*, 2C] (keys in columns
// [0, C), values in [C, 2C)): query row r sees rows [offsets[r / Q], offsets[r / Q + 1]).
int g_cross_tc = 1;   // 0: the CUDA-core kernels only (A/B timing via setok_debug_set_cross_attention_tc)

// kv_rows: row capacity of `kv` (the tensor-core kernel's TMA map is bounded by it); 0 = unknown -> CUDA-core kernels
int launch_cross_attention(const void* q, const void* kv, void* out, int rows, int Q, int C, int heads, float scale,
                           const int32_t* offsets, int kv_rows, cudaStream_t stream) {
  SETOK_REQUIRE(q && kv && out && offsets, SETOK_ERR_BAD_ARG, "cross_attention: null pointer");
  SETOK_REQUIRE(rows > 0 && Q > 0 && rows % Q == 0 && C > 0 && heads > 0 && C % heads == 0, SETOK_ERR_BAD_ARG, "cross_attention: bad shape rows=%d Q=%d C=%d heads=%d", rows, Q, C, heads);
  const int hd = C / heads;
  SETOK_REQUIRE(hd % 8 == 0 && hd <= 512, SETOK_ERR_UNSUPPORTED, "cross_attention: head_dim %d unsupported (need %%8==0, <=512)", hd);
  SETOK_REQUIRE(aligned16(q) && aligned16(kv) && aligned16(out), SETOK_ERR_BAD_ARG, "cross_attention: buffers must be 16-byte aligned");
  if (hd == 64 && g_cross_tc && kv_rows > 0 && C % 8 == 0)
    return launch_cross_attention_tcgen05(q, kv, out, rows / Q, Q, C, heads, scale, offsets, kv_rows, stream);
  if (hd == XA_HD) {
    dim3 grid(heads, rows / Q, ceil_div(Q, XA_THREADS));
    cross_attn_hd64_kernel<<<grid, XA_THREADS, 0, stream>>>(static_cast<const bf16*>(q), static_cast<const bf16*>(kv), static_cast<bf16*>(out), Q, C,
                                                            scale * 1.4426950408889634f, offsets);
    SETOK_LAUNCH_CHECK();
    return SETOK_OK;
  }
  const long long warps = static_cast<long long>(rows) * heads;
  long long blocks = (warps + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  const bf16* k = static_cast<const bf16*>(kv);
  if (hd <= 256)
    attn_seg_kernel<1><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(q), C, k, k + C, 2LL * C, static_cast<bf16*>(out), rows, C, heads, scale, offsets, nullptr, 0, Q, nullptr);
  else
    attn_seg_kernel<2><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(q), C, k, k + C, 2LL * C, static_cast<bf16*>(out), rows, C, heads, scale, offsets, nullptr, 0, Q, nullptr);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" int setok_attention(const void* qkv, void* out, int rows, int C, int heads, float scale, const int32_t* seg_off,
                               const int32_t* row_seg, int uniform_T, const int32_t* m_dev, setok_stream_t stream) {
  return setok::launch_attention(qkv, out, rows, C, heads, scale, seg_off, row_seg, uniform_T, m_dev, static_cast<cudaStream_t>(stream));
}

extern "C" void setok_debug_set_cross_attention_tc(int on) { setok::g_cross_tc = on; }
