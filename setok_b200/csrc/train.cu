// train.cu — primitives of the head's TRAINING path: what torch.autograd does implicitly for the
// reference when gradients flow through group_encoding / inter_encoder / out / mm_projector (src/model/setok/tokenizer.py:
// 123-155, 179-180; module.py:29-100; multimodal_projector/builder.py:33-64).  The contractions of the backward pass
// (dgrad = dY W, wgrad = dY^T X, and the four products of the attention backward) reuse the tcgen05 GEMM with its
// MN-major B operand form; this file holds what sits between them: the bf16 transposes that put dY^T / P^T / dS^T in
// K-major form, column sums (bias gradients), LayerNorm / GELU / masked-softmax backward, the segment-mean pair and the
// indexed row gather used for padding ragged token batches.  All bandwidth-bound row kernels: coalesced, 16-byte vectors
// where the layout allows, grid-stride over whole waves.
#include "common.cuh"
#include "rowops.cuh"

namespace setok {
namespace {

inline int tr_grid(long long items, int threads, int waves = 8) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * waves;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

// out[b][c][r] = bf16(in[b][r][c]); 32 x 32 tiles through shared memory (+1 padding), both sides coalesced
template <class TI>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const TI* __restrict__ in, long long ld_in, long long in_bs, bf16* __restrict__ out,
                                                                long long ld_out, long long out_bs, int rows, int cols) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  const int tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  const long long per_batch = static_cast<long long>(tiles_c) * tiles_r;
  for (long long t = blockIdx.x; t < per_batch; t += gridDim.x) {
    const int b = blockIdx.y;
    const int r0 = static_cast<int>(t / tiles_c) * 32, c0 = static_cast<int>(t % tiles_c) * 32;
    const TI* src = in + b * in_bs;
    bf16* dst = out + b * out_bs;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      tile[ty + 8 * i][tx] = (r < rows && c < cols) ? to_f32<TI>(src[static_cast<long long>(r) * ld_in + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;
      if (c < cols && r < rows) dst[static_cast<long long>(c) * ld_out + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}

// out[c] += sum_r in[r][c]: one thread per column within a block of 128 columns x a slab of rows; fp32 atomics across slabs
template <class TI>
__global__ void __launch_bounds__(128) colsum_kernel(const TI* __restrict__ in, long long ld, int rows, int cols, float* __restrict__ out, int slab) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * slab, r1 = min(rows, r0 + slab);
  if (c >= cols) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += to_f32<TI>(in[static_cast<long long>(r) * ld + c]); s1 += to_f32<TI>(in[static_cast<long long>(r + 1) * ld + c]);
    s2 += to_f32<TI>(in[static_cast<long long>(r + 2) * ld + c]); s3 += to_f32<TI>(in[static_cast<long long>(r + 3) * ld + c]);
  }
  for (; r < r1; ++r) s0 += to_f32<TI>(in[static_cast<long long>(r) * ld + c]);
  atomicAdd(out + c, (s0 + s1) + (s2 + s3));
}

// LayerNorm backward, one warp per row (C <= 4096): with xhat = (x - mean) rstd, g = gamma dy:
//   dx = rstd (g - mean(g) - xhat mean(g xhat));  dgamma += dy xhat;  dbeta += dy   (per-block partials, then fp32 atomics)
template <class TD>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const TD* __restrict__ dy, const float* __restrict__ gamma,
                                                            float eps, int rows, int C, float* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
  extern __shared__ float sacc[];                      // [2][C] per-block partial dgamma / dbeta
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const float invC = 1.0f / static_cast<float>(C);
  for (int r = blockIdx.x * wpb + wib; r < rows; r += gridDim.x * wpb) {
    const float* xr = x + static_cast<long long>(r) * C;
    const TD* dyr = dy + static_cast<long long>(r) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invC + eps);
    float sg = 0.f, sgx = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mean) * rstd, g = gamma[c] * to_f32<TD>(dyr[c]);
      sg += g; sgx += g * xh;
    }
    sg = warp_sum(sg) * invC; sgx = warp_sum(sgx) * invC;
    float* dxr = dx + static_cast<long long>(r) * C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mean) * rstd, d = to_f32<TD>(dyr[c]);
      dxr[c] = rstd * (gamma[c] * d - sg - xh * sgx);
      atomicAdd(&sacc[c], d * xh);
      atomicAdd(&sacc[C + c], d);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (sacc[c] != 0.f) atomicAdd(dgamma + c, sacc[c]);
    if (sacc[C + c] != 0.f) atomicAdd(dbeta + c, sacc[C + c]);
  }
}

// GELU(erf) forward (pre -> act, bf16) and backward: d/dx [0.5 x (1 + erf(x / sqrt 2))] = Phi(x) + x phi(x)
template <class TI>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const TI* __restrict__ pre, bf16* __restrict__ act, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = to_f32<TI>(pre[i]);
    act[i] = __float2bfloat16_rn(0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)));
  }
}
template <class TI, class TD>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const TI* __restrict__ pre, const TD* __restrict__ dy, float* __restrict__ dpre, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = to_f32<TI>(pre[i]);
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    dpre[i] = to_f32<TD>(dy[i]) * (cdf + x * pdf);
  }
}

// masked-softmax backward: dS[r, j] = scale P[r, j] (dP[r, j] - sum_k P[r, k] dP[r, k]) inside row r's segment, 0 outside
__global__ void __launch_bounds__(256) masked_softmax_bwd_kernel(const bf16* __restrict__ P, long long ldP, const float* __restrict__ dP, long long lddP,
                                                                 const int32_t* __restrict__ seg_off, const int32_t* __restrict__ row_seg, int rows, int N,
                                                                 float scale, bf16* __restrict__ dS, long long lddS) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int img0 = (r / N) * N;
    const int g = row_seg[r];
    const int k0 = seg_off[g] - img0, k1 = seg_off[g + 1] - img0;
    const bf16* p = P + static_cast<long long>(r) * ldP;
    const float* dp = dP + static_cast<long long>(r) * lddP;
    float dot = 0.f;
    for (int j = k0 + lane; j < k1; j += 32) dot += __bfloat162float(p[j]) * dp[j];
    dot = warp_sum(dot);
    bf16* ds = dS + static_cast<long long>(r) * lddS;
    for (int j = lane; j < lddS; j += 32) {
      float v = 0.f;
      if (j >= k0 && j < k1) v = scale * __bfloat162float(p[j]) * (dp[j] - dot);
      ds[j] = __float2bfloat16_rn(v);
    }
  }
}

// segment mean over sorted rows and its backward (dx[r] = dg[row_seg[r]] / |segment|)
__global__ void __launch_bounds__(256) segment_mean_bwd_kernel(const float* __restrict__ dg, const int32_t* __restrict__ seg_off,
                                                               const int32_t* __restrict__ row_seg, int rows, int C, float* __restrict__ dx) {
  const int nvec = C >> 2;
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vi = static_cast<int>(i % nvec);
    const int r = static_cast<int>(i / nvec);
    const int g = row_seg[r];
    const float inv = 1.0f / static_cast<float>(seg_off[g + 1] - seg_off[g]);
    float4 v = reinterpret_cast<const float4*>(dg)[static_cast<long long>(g) * nvec + vi];
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    reinterpret_cast<float4*>(dx)[i] = v;
  }
}

// out[r] = index[r] >= 0 ? in[index[r]] : 0  (pads / unpads ragged batches; a permutation's inverse is the same call)
__global__ void __launch_bounds__(256) gather_rows_idx_kernel(const float* __restrict__ in, const int32_t* __restrict__ index, int rows_out, int C,
                                                              float* __restrict__ out) {
  const int nvec = C >> 2;
  const long long total = static_cast<long long>(rows_out) * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vi = static_cast<int>(i % nvec);
    const int r = static_cast<int>(i / nvec);
    const int s = index[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0) v = reinterpret_cast<const float4*>(in)[static_cast<long long>(s) * nvec + vi];
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

}  // namespace
}  // namespace setok

using namespace setok;

#define TR_DT_OK(dt) ((dt) == SETOK_F32 || (dt) == SETOK_BF16)

extern "C" int setok_transpose_to_bf16(const void* in, int in_dtype, int64_t ld_in, int64_t in_batch_stride, void* out, int64_t ld_out,
                                       int64_t out_batch_stride, int rows, int cols, int batch, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(in && out && rows > 0 && cols > 0 && batch > 0 && batch <= 65535, SETOK_ERR_BAD_ARG, "transpose: bad arguments");
  SETOK_REQUIRE(TR_DT_OK(in_dtype) && ld_in >= cols && ld_out >= rows, SETOK_ERR_BAD_ARG, "transpose: bad dtype / leading dimensions");
  const long long tiles = static_cast<long long>(ceil_div(rows, 32)) * ceil_div(cols, 32);
  const int gx = static_cast<int>(tiles < 65535LL * 8 ? (tiles < num_sms() * 8 ? tiles : num_sms() * 8) : num_sms() * 8);
  dim3 grid(gx, batch);
  if (in_dtype == SETOK_F32)
    transpose_to_bf16_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), ld_in, in_batch_stride, static_cast<bf16*>(out), ld_out, out_batch_stride, rows, cols);
  else
    transpose_to_bf16_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(in), ld_in, in_batch_stride, static_cast<bf16*>(out), ld_out, out_batch_stride, rows, cols);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_colsum_add(const void* in, int in_dtype, int64_t ld, int rows, int cols, float* out, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(in && out && rows > 0 && cols > 0 && TR_DT_OK(in_dtype) && ld >= cols, SETOK_ERR_BAD_ARG, "colsum: bad arguments");
  const int slab = 256;
  dim3 grid(ceil_div(cols, 128), ceil_div(rows, slab));
  SETOK_REQUIRE(grid.y <= 65535, SETOK_ERR_UNSUPPORTED, "colsum: too many rows");
  if (in_dtype == SETOK_F32) colsum_kernel<float><<<grid, 128, 0, stream>>>(static_cast<const float*>(in), ld, rows, cols, out, slab);
  else colsum_kernel<bf16><<<grid, 128, 0, stream>>>(static_cast<const bf16*>(in), ld, rows, cols, out, slab);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_layernorm_bwd(const float* x, const void* dy, int dy_dtype, const float* gamma, float eps, int rows, int C, float* dx,
                                   float* dgamma, float* dbeta, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(x && dy && gamma && dx && dgamma && dbeta && rows > 0 && C > 0 && TR_DT_OK(dy_dtype), SETOK_ERR_BAD_ARG, "layernorm_bwd: bad arguments");
  SETOK_REQUIRE(C <= 6144, SETOK_ERR_UNSUPPORTED, "layernorm_bwd: C = %d exceeds the shared-memory partials (6144)", C);
  int grid = ceil_div(rows, 8);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  const size_t smem = 2 * static_cast<size_t>(C) * sizeof(float);
  if (dy_dtype == SETOK_F32) layernorm_bwd_kernel<float><<<grid, 256, smem, stream>>>(x, static_cast<const float*>(dy), gamma, eps, rows, C, dx, dgamma, dbeta);
  else layernorm_bwd_kernel<bf16><<<grid, 256, smem, stream>>>(x, static_cast<const bf16*>(dy), gamma, eps, rows, C, dx, dgamma, dbeta);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_gelu_fwd(const void* pre, int pre_dtype, void* act_bf16, int64_t n, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(pre && act_bf16 && n > 0 && TR_DT_OK(pre_dtype), SETOK_ERR_BAD_ARG, "gelu_fwd: bad arguments");
  if (pre_dtype == SETOK_F32) gelu_fwd_kernel<float><<<tr_grid(n, 256, 16), 256, 0, stream>>>(static_cast<const float*>(pre), static_cast<bf16*>(act_bf16), n);
  else gelu_fwd_kernel<bf16><<<tr_grid(n, 256, 16), 256, 0, stream>>>(static_cast<const bf16*>(pre), static_cast<bf16*>(act_bf16), n);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_gelu_bwd(const void* pre, int pre_dtype, const void* dy, int dy_dtype, float* dpre, int64_t n, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(pre && dy && dpre && n > 0 && TR_DT_OK(pre_dtype) && TR_DT_OK(dy_dtype), SETOK_ERR_BAD_ARG, "gelu_bwd: bad arguments");
  const int grid = tr_grid(n, 256, 16);
#define GB(TI, TD) gelu_bwd_kernel<TI, TD><<<grid, 256, 0, stream>>>(static_cast<const TI*>(pre), static_cast<const TD*>(dy), dpre, n)
  if (pre_dtype == SETOK_F32) { if (dy_dtype == SETOK_F32) GB(float, float); else GB(float, bf16); }
  else { if (dy_dtype == SETOK_F32) GB(bf16, float); else GB(bf16, bf16); }
#undef GB
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_masked_softmax(const float* S, int64_t ldS, const int32_t* seg_off, const int32_t* row_seg, int rows, int N, float scale,
                                    void* P_bf16, int64_t ldP, setok_stream_t stream_) {
  SETOK_REQUIRE(S && seg_off && row_seg && P_bf16 && rows > 0 && N > 0 && rows % N == 0 && ldS >= N && ldP >= N, SETOK_ERR_BAD_ARG, "masked_softmax: bad arguments");
  return launch_masked_softmax(S, P_bf16, seg_off, row_seg, rows, N, static_cast<int>(ldS), static_cast<int>(ldP), scale, static_cast<cudaStream_t>(stream_));
}

extern "C" int setok_masked_softmax_bwd(const void* P_bf16, int64_t ldP, const float* dP, int64_t lddP, const int32_t* seg_off, const int32_t* row_seg,
                                        int rows, int N, float scale, void* dS_bf16, int64_t lddS, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(P_bf16 && dP && seg_off && row_seg && dS_bf16 && rows > 0 && N > 0 && rows % N == 0 && ldP >= N && lddP >= N && lddS >= N,
                SETOK_ERR_BAD_ARG, "masked_softmax_bwd: bad arguments");
  int grid = ceil_div(rows, 8);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  masked_softmax_bwd_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(P_bf16), ldP, dP, lddP, seg_off, row_seg, rows, N, scale,
                                                      static_cast<bf16*>(dS_bf16), lddS);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_sort_by_cluster(const int64_t* idx_cluster, const int32_t* num_clusters, const int32_t* offsets, int B, int N, int32_t* perm,
                                     int32_t* row_seg, int32_t* seg_off, setok_stream_t stream_) {
  SETOK_REQUIRE(idx_cluster && num_clusters && offsets && perm && row_seg && seg_off && B > 0 && N > 0, SETOK_ERR_BAD_ARG, "sort_by_cluster: bad arguments");
  return launch_sort_by_cluster(idx_cluster, num_clusters, offsets, B, N, perm, row_seg, seg_off, static_cast<cudaStream_t>(stream_));
}

extern "C" int setok_segment_mean(const float* x, const int32_t* seg_off, int n_segments, int C, float* out, setok_stream_t stream_) {
  SETOK_REQUIRE(x && seg_off && out && n_segments > 0 && C > 0 && C % 4 == 0, SETOK_ERR_BAD_ARG, "segment_mean: bad arguments");
  return launch_segment_mean(x, seg_off, nullptr, n_segments, C, out, nullptr, static_cast<cudaStream_t>(stream_));
}

extern "C" int setok_segment_mean_bwd(const float* dg, const int32_t* seg_off, const int32_t* row_seg, int rows, int C, float* dx, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(dg && seg_off && row_seg && dx && rows > 0 && C > 0 && C % 4 == 0, SETOK_ERR_BAD_ARG, "segment_mean_bwd: bad arguments");
  segment_mean_bwd_kernel<<<tr_grid(static_cast<long long>(rows) * (C / 4), 256, 16), 256, 0, stream>>>(dg, seg_off, row_seg, rows, C, dx);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

extern "C" int setok_gather_rows_f32(const float* in, const int32_t* index, int rows_out, int C, float* out, setok_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(in && index && out && rows_out > 0 && C > 0 && C % 4 == 0, SETOK_ERR_BAD_ARG, "gather_rows: bad arguments");
  gather_rows_idx_kernel<<<tr_grid(static_cast<long long>(rows_out) * (C / 4), 256, 16), 256, 0, stream>>>(in, index, rows_out, C, out);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}
