// attention.cu — self-attention over packed rows.
//
// Two kernels, chosen by shape (not by backend):
//  * attn_mma_hd64_kernel: uniform-length segments (the ViT: T = N+1 rows per image), head_dim 64.
//    Flash-style single pass, bf16 warp MMA with fp32 softmax statistics, K/V chunks staged in
//    XOR-swizzled shared memory.  (Round-1 implementation; the tcgen05/TMEM version replaces it.)
//  * attn_seg_kernel: arbitrary ragged segments (clusters of 1..N tokens, images of K_b tokens),
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
template <int VPL>
__global__ void __launch_bounds__(256) attn_seg_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int rows, int C,
                                                       int heads, float scale, const int32_t* __restrict__ seg_off,
                                                       const int32_t* __restrict__ row_seg, int uniform_T,
                                                       const int32_t* __restrict__ m_dev) {
  const int lane = threadIdx.x & 31;
  const int n = live_rows_a(rows, m_dev);
  const long long total = static_cast<long long>(n) * heads;
  const int hd = C / heads;
  const int nv = hd >> 3;
  const long long ld = 3LL * C;
  for (long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); wid < total;
       wid += static_cast<long long>(gridDim.x) * (blockDim.x >> 5)) {
    const int r = static_cast<int>(wid / heads);
    const int h = static_cast<int>(wid % heads);
    int s0, s1;
    if (uniform_T > 0) { s0 = (r / uniform_T) * uniform_T; s1 = s0 + uniform_T; }
    else { const int s = row_seg[r]; s0 = seg_off[s]; s1 = seg_off[s + 1]; }
    float q[VPL][8], o[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
#pragma unroll
      for (int e = 0; e < 8; ++e) { q[i][e] = 0.f; o[i][e] = 0.f; }
      if (vi < nv) {
        const uint4 u = *reinterpret_cast<const uint4*>(qkv + r * ld + h * hd + vi * 8);
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
              const uint4 u = *reinterpret_cast<const uint4*>(qkv + jj * ld + C + h * hd + vi * 8);
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
              const uint4 u = *reinterpret_cast<const uint4*>(qkv + jj * ld + 2 * C + h * hd + vi * 8);
              const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
              o[i][0] += p[u4] * a.x; o[i][1] += p[u4] * a.y; o[i][2] += p[u4] * b.x; o[i][3] += p[u4] * b.y;
              o[i][4] += p[u4] * c.x; o[i][5] += p[u4] * c.y; o[i][6] += p[u4] * d.x; o[i][7] += p[u4] * d.y;
            }
          }
        }
      }
    }
    const float inv = 1.0f / l;
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
// uniform segments of T rows, head_dim 64
// -------------------------------------------------------------------------------------------------
constexpr int AQ = 64, AK = 64, HD = 64;

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {   // byte offset in a [rows][64 bf16] swizzled tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void load_tile_64x64(bf16* s, const bf16* g, long long ld, int row0, int row_end, int tid) {
  // 64 rows x 8 chunks of 16 B; rows >= row_end are zero-filled
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, c = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + row < row_end) v = *reinterpret_cast<const uint4*>(g + (row0 + row) * ld + c * 8);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(s) + sw_off(row, c)) = v;
  }
}

__global__ void __launch_bounds__(128) attn_mma_hd64_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int T, int heads,
                                                            int C, float scale_log2) {
  __shared__ __align__(128) bf16 sQ[AQ * HD];
  __shared__ __align__(128) bf16 sK[AK * HD];
  __shared__ __align__(128) bf16 sV[AK * HD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const long long ld = 3LL * C;
  const long long img_row0 = static_cast<long long>(b) * T;
  const bf16* gq = qkv + img_row0 * ld + h * HD;
  const bf16* gk = gq + C;
  const bf16* gv = gq + 2 * C;

  load_tile_64x64(sQ, gq, ld, qt * AQ, T, tid);
  __syncthreads();
  const bool active = qt * AQ + warp * 16 < T;
  uint32_t qf[4][4];
  {
    const uint32_t qbase = smem_u32(sQ);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int chunk = ks * 2 + (lane >> 4);
      ldmatrix_x4(qbase + sw_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
  }
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const uint32_t kbase = smem_u32(sK), vbase = smem_u32(sV);
  const int nchunks = (T + AK - 1) / AK;
  for (int kc = 0; kc < nchunks; ++kc) {
    __syncthreads();
    load_tile_64x64(sK, gk, ld, kc * AK, T, tid);
    load_tile_64x64(sV, gv, ld, kc * AK, T, tid);
    __syncthreads();
    if (!active) continue;
    const int valid = min(AK, T - kc * AK);
    const int nkg = (valid + 15) >> 4;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g < nkg) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int mi = lane >> 3;
          const int row = g * 16 + (mi >> 1) * 8 + (lane & 7);
          const int chunk = ks * 2 + (mi & 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(kbase + sw_off(row, chunk), b0, b1, b2, b3);
          mma_bf16_16816(s[2 * g], qf[ks], b0, b1);
          mma_bf16_16816(s[2 * g + 1], qf[ks], b2, b3);
        }
      }
    }
    // mask + running max
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + 2 * (lane & 3) + (e & 1);
        const float v = key < valid ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
      s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
      ps0 += s[nt][0] + s[nt][1]; ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * a0 + ps0; l1 = l1 * a1 + ps1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= a0; o[dt][1] *= a0; o[dt][2] *= a1; o[dt][3] *= a1; }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g < nkg) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * g][0], s[2 * g][1]);
        pa[1] = pack_bf16x2(s[2 * g][2], s[2 * g][3]);
        pa[2] = pack_bf16x2(s[2 * g + 1][0], s[2 * g + 1][1]);
        pa[3] = pack_bf16x2(s[2 * g + 1][2], s[2 * g + 1][3]);
#pragma unroll
        for (int d2 = 0; d2 < 4; ++d2) {
          const int mi = lane >> 3;
          const int row = g * 16 + (mi & 1) * 8 + (lane & 7);
          const int chunk = 2 * d2 + (mi >> 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(vbase + sw_off(row, chunk), b0, b1, b2, b3);
          mma_bf16_16816(o[2 * d2], pa, b0, b1);
          mma_bf16_16816(o[2 * d2 + 1], pa, b2, b3);
        }
      }
    }
  }
  if (!active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = qt * AQ + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = h * HD + dt * 8 + 2 * (lane & 3);
    if (r0 < T) *reinterpret_cast<uint32_t*>(out + (img_row0 + r0) * C + col) = pack_bf16x2(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(out + (img_row0 + r1) * C + col) = pack_bf16x2(o[dt][2] * i1, o[dt][3] * i1);
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
  if (uniform_T > 0 && hd == HD && m_dev == nullptr) {
    const float scale_log2 = scale * 1.4426950408889634f;
    dim3 grid(ceil_div(uniform_T, AQ), heads, rows / uniform_T);
    attn_mma_hd64_kernel<<<grid, 128, 0, stream>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), uniform_T, heads, C, scale_log2);
    SETOK_LAUNCH_CHECK();
    return SETOK_OK;
  }
  const long long warps = static_cast<long long>(rows) * heads;
  long long blocks = (warps + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  if (hd <= 256)
    attn_seg_kernel<1><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), rows, C, heads, scale, seg_off, row_seg, uniform_T, m_dev);
  else
    attn_seg_kernel<2><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), rows, C, heads, scale, seg_off, row_seg, uniform_T, m_dev);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" int setok_attention(const void* qkv, void* out, int rows, int C, int heads, float scale, const int32_t* seg_off,
                               const int32_t* row_seg, int uniform_T, const int32_t* m_dev, setok_stream_t stream) {
  return setok::launch_attention(qkv, out, rows, C, heads, scale, seg_off, row_seg, uniform_T, m_dev, static_cast<cudaStream_t>(stream));
}
