// attention_tcgen05.cu — ViT multi-head self-attention (head_dim 64, uniform T rows per image) on the
// 5th-generation tensor cores.
//
// One CTA per (image, head, 128-query tile); 192 threads, three CTAs per SM (64 KB smem, 128 TMEM columns each)
// so one CTA's softmax (MUFU-bound) overlaps the others' MMAs and barrier latencies:
//   warp 4      TMA producer: Q tile once, then K/V chunks of 64 keys through a 2-stage mbarrier ring.
//               3-D tensor maps over qkv viewed as [B][T][3C]; rows past T are zero-filled by TMA.
//   warp 5      MMA issuer (one thread), owns the TMEM allocation; the highest warp id in the CTA because the
//               scheduler favours high warp ids and this warp sits on every chunk's critical path:
//                 S = Q K_j^T   128 x nk (nk <= 64) x 64    -> TMEM columns [0, 64)
//                 O += P_j V_j  128 x 64 x nk               -> TMEM columns [64, 128), V as MN-major operand
//   warps 0..3  softmax, one thread per query row: tcgen05.ld S -> online softmax in the exp2 domain ->
//               P (bf16 pairs) back into TMEM over the S columns (tcgen05.st; the P.V MMA takes its A operand from
//               tensor memory) -> rescale O in TMEM when the running max moved -> hand over to the MMA warp.
//               After the last chunk: O / l -> bf16 -> global.
// Tensor-core work is issued in order, so "S_j complete" implies "P_{j-1} V_{j-1} complete": the single S and P
// buffers need no further handshakes.  Latency is hidden across the three CTAs resident per SM.
#include "common.cuh"

#include <mutex>

namespace setok {
namespace {

constexpr int HD = 64, QT = 128, KT = 64;
constexpr int KV_STAGE = KT * HD * 2;          // 8 KiB
constexpr int OFF_Q = 0;                       // 128 x 64 bf16 = 16 KiB
constexpr int KV_STAGES = 2;
constexpr int OFF_K = 16384;                   // KV_STAGES x 8 KiB
constexpr int OFF_V = OFF_K + KV_STAGES * KV_STAGE;
constexpr int OFF_BARS = OFF_V + KV_STAGES * KV_STAGE;
constexpr int ATT_SMEM = OFF_BARS + 256;       // no alignment slack: the dynamic window starts 1 KiB aligned (checked)
constexpr int ATT_THREADS = 192;
constexpr int ATT_TMEM_COLS = 128;
constexpr uint32_t O_COL = KT;
constexpr uint32_t P_COL = 0;

__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; exp2(-inf) = 0, inputs here are <= 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// CROSS = false: self-attention over qkv [B][T][3C] (keys = the image's own T rows).
// CROSS = true : the Q-Former's cross-attention (module.py:283-364): queries q [B][T][C], keys / values the image's own rows
//                [offsets[b], offsets[b+1]) of the packed kv [*, 2C] ([k | v]); a varlen kernel keyed by `offsets`, no padding mask.
template <bool CROSS>
__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_tcgen05_hd64_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_kv, bf16* __restrict__ out,
                         int T, int heads, int C, float scale_log2, long long* __restrict__ dbg, const int32_t* __restrict__ offsets) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
#ifdef SETOK_ATTN_TRACE
  const bool trace = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 1 && blockIdx.z == 1;
#define TRACE(slot) do { if (trace) dbg[slot] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw;
  if ((base & 1023u) != 0u) __trap();   // the 128B-swizzled tiles need a 1 KiB aligned window
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * QT;

  const uint32_t bar0 = base + OFF_BARS;
  const uint32_t q_full = bar0, s_full = bar0 + 8, p_full = bar0 + 16, o_final = bar0 + 24;
  auto kv_full = [&](int s) { return bar0 + 40u + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 40u + 8u * (KV_STAGES + s); };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + OFF_BARS + 112);

  constexpr int W_TMA = 4, W_MMA = 5;
  if (threadIdx.x == 0) TRACE(0);
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_final, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc<ATT_TMEM_COLS>(base + OFF_BARS + 112);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();                                   // qkv comes from the previous kernel of the stream
  if (threadIdx.x == 0) TRACE(1);
  // keys of this CTA: Tk rows starting at row krow0 of batch index kb of the key/value map; columns kcol (K) / vcol (V)
  const int krow0 = CROSS ? offsets[b] : 0;
  const int Tk = CROSS ? offsets[b + 1] - krow0 : T;
  const int kb = CROSS ? 0 : b;
  const int kcol = (CROSS ? 0 : C) + h * HD, vcol = (CROSS ? C : 2 * C) + h * HD;
  const int nchunks = (Tk + KT - 1) / KT;
  if (CROSS && Tk <= 0) {
    // an image without tokens: zeros (no barrier is armed, every role falls through to the common exit)
    if (warp < 4) {
      const int grow = q0 + (warp & 3) * 32 + lane;
      if (grow < T) {
        bf16* dst = out + (static_cast<long long>(b) * T + grow) * C + h * HD;
#pragma unroll
        for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(dst + 8 * g) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  } else

  if (warp == W_TMA) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, QT * HD * 2);
      tma_load_3d(&tm, q_full, base + OFF_Q, h * HD, q0, b);
      for (int j = 0; j < nchunks; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(kv_empty(s), ((j / KV_STAGES) & 1) ^ 1u);
        mbar_arrive_expect_tx(kv_full(s), 2 * KV_STAGE);
        tma_load_3d(&tm_kv, kv_full(s), base + OFF_K + s * KV_STAGE, kcol, krow0 + j * KT, kb);
        tma_load_3d(&tm_kv, kv_full(s), base + OFF_V + s * KV_STAGE, vcol, krow0 + j * KT, kb);
      }
    }
  } else if (warp == W_MMA) {
    {
      // The whole warp runs this role convergently (waits included) and one elected lane issues the tcgen05
      // instructions: with warp-uniform control flow the descriptors live in uniform registers, which the MMA needs,
      // instead of being moved there lane-by-lane before every instruction (this issue path is on the critical chain).
      // all operand descriptors are loop-invariant: build them once
      // (stage s of K/V only shifts the 14-bit address field by s * KV_STAGE / 16)
      const uint64_t dq0 = umma_desc_k_sw128(base + OFF_Q), dq1 = dq0 + 2, dq2 = dq0 + 4, dq3 = dq0 + 6;   // +32 B per k-step
      const uint32_t tp = tmem_base + P_COL;                 // P_j (bf16 pairs) overwrites the first half of the S columns
      const uint64_t dk0 = umma_desc_k_sw128(base + OFF_K), dk1 = dk0 + 2, dk2 = dk0 + 4, dk3 = dk0 + 6;
      const uint64_t dv0 = umma_desc_mn_sw128(base + OFF_V), dv1 = dv0 + 128, dv2 = dv0 + 256, dv3 = dv0 + 384;   // +2048 B
      const uint32_t idesc_pv = umma_idesc_bf16(QT, HD, true);
      const uint32_t idesc_s_full = umma_idesc_bf16(QT, KT);
#define ISSUE_S(J)                                                                                   \
      do {                                                                                             \
        const int j_ = (J);                                                                            \
        const int nk_ = min(KT, ((Tk - j_ * KT) + 15) & ~15);                                          \
        const uint32_t idesc_ = nk_ == KT ? idesc_s_full : umma_idesc_bf16(QT, nk_);                   \
        const int st_ = j_ % KV_STAGES;                                                                \
        mbar_wait(kv_full(st_), (j_ / KV_STAGES) & 1);                                                 \
        tcgen05_fence_after();                                                                         \
        const uint64_t soff_ = static_cast<uint64_t>(st_ * (KV_STAGE >> 4));                           \
        if (elect_one_sync()) {                                                                        \
          umma_f16(tmem_base, dq0, dk0 + soff_, idesc_, 0u);                                           \
          umma_f16(tmem_base, dq1, dk1 + soff_, idesc_, 1u);                                           \
          umma_f16(tmem_base, dq2, dk2 + soff_, idesc_, 1u);                                           \
          umma_f16(tmem_base, dq3, dk3 + soff_, idesc_, 1u);                                           \
          umma_commit(s_full);                                                                         \
        }                                                                                              \
        __syncwarp();                                                                                  \
      } while (0)
      mbar_wait(q_full, 0);
      if (lane == 0) TRACE(2);
      ISSUE_S(0);
      if (lane == 0) TRACE(3);
      for (int j = 0; j < nchunks; ++j) {
        const int nk16 = min(KT, ((Tk - j * KT) + 15) & ~15) >> 4;
        mbar_wait(p_full, j & 1);
        if (lane == 0) TRACE(10 + 4 * j);
        tcgen05_fence_after();
        const int st = j % KV_STAGES;
        const uint64_t soff = static_cast<uint64_t>(st * (KV_STAGE >> 4));
        if (elect_one_sync()) {
          umma_f16_ts(tmem_base + O_COL, tp, dv0 + soff, idesc_pv, j != 0 ? 1u : 0u);      // 16 keys = 8 TMEM columns per step
          if (nk16 > 1) umma_f16_ts(tmem_base + O_COL, tp + 8, dv1 + soff, idesc_pv, 1u);
          if (nk16 > 2) umma_f16_ts(tmem_base + O_COL, tp + 16, dv2 + soff, idesc_pv, 1u);
          if (nk16 > 3) umma_f16_ts(tmem_base + O_COL, tp + 24, dv3 + soff, idesc_pv, 1u);
          umma_commit(kv_empty(st));
          if (j + 1 >= nchunks) umma_commit(o_final);
        }
        __syncwarp();
        if (j + 1 < nchunks) ISSUE_S(j + 1);
        if (lane == 0) TRACE(11 + 4 * j);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                  // query row within the tile == TMEM lane
    const bool warp_live = q0 + q * 32 < T;         // does this warp own any real query row?
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_o = t_s + O_COL;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nchunks; ++j) {
      mbar_wait(s_full, j & 1);
      if (warp == 0 && lane == 0) TRACE(40 + 4 * j);
      tcgen05_fence_after();
      if (warp_live) {
        const int valid = min(KT, Tk - j * KT);       // real keys in this chunk (columns >= valid hold stale data)
        float m_new, alpha, psum;
        // S row (64 fp32) read from TMEM once; the probabilities are packed to bf16 pairs in place and stored back to
        // TMEM over the first 32 S columns, where the P.V MMA reads them as its A operand (no trip through smem)
        uint32_t r[KT];
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_wait();
        if (valid < KT) {
#pragma unroll
          for (int i = 0; i < KT; ++i) if (i >= valid) r[i] = 0xff800000u;   // -inf -> probability 0
        }
        float mx0 = __uint_as_float(r[0]), mx1 = __uint_as_float(r[1]), mx2 = __uint_as_float(r[2]), mx3 = __uint_as_float(r[3]);
#pragma unroll
        for (int i = 4; i < KT; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
        }
        m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2);
        alpha = fast_exp2(m - m_new);
        if (warp == 0 && lane == 0) TRACE(41 + 4 * j);
        float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
#pragma unroll
        for (int i = 0; i < KT; i += 4) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, -m_new));
          const float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
          const float p2 = fast_exp2(fmaf(__uint_as_float(r[i + 2]), scale_log2, -m_new));
          const float p3 = fast_exp2(fmaf(__uint_as_float(r[i + 3]), scale_log2, -m_new));
          ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
          r[i / 2] = pack_bf16x2(p0, p1);               // in place: slots <= i/2+1 were consumed already
          r[i / 2 + 1] = pack_bf16x2(p2, p3);
        }
        psum = (ps0 + ps1) + (ps2 + ps3);
        tmem_st_32x32b_x32(t_s + P_COL, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        l = l * alpha + psum;
        m = m_new;
        if (warp == 0 && lane == 0) TRACE(42 + 4 * j);
        // rescale the running output when any row's maximum moved (S_j complete => P_{j-1}V_{j-1} complete)
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_o + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(t_o + c0, o);
          }
        }
        tmem_st_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 0 && lane == 0) TRACE(43 + 4 * j);
    }
    mbar_wait(o_final, 0);
    if (warp == 0 && lane == 0) TRACE(80);
    tcgen05_fence_after();
    if (warp_live) {
      const float inv = 1.0f / l;
      const int grow = q0 + row;
      bf16* dst = out + (static_cast<long long>(b) * T + grow) * C + h * HD;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_o + c0, o);
        tmem_ld_wait();
        if (grow < T) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c0 + 8 * g) = u;
          }
        }
      }
    }
  }
  if (warp == 0 && lane == 0) TRACE(81);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
  if (threadIdx.x == W_MMA * 32) TRACE(82);
#undef TRACE
#undef ISSUE_S
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

}  // namespace

long long* g_attn_trace = nullptr;   // debug: device buffer of 128 clock64 slots for one CTA's timeline (tools/attn_timeline.py)

// qkv bf16 [B*T, 3C] -> out bf16 [B*T, C]; heads of 64; softmax(q k^T * scale) v per image.
int launch_attention_tcgen05(const void* qkv, void* out, int B, int T, int C, int heads, float scale, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SETOK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tm;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(3 * C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(3 * C) * 2, static_cast<cuuint64_t>(T) * 3 * C * 2};
  cuuint32_t box[3] = {HD, QT, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "attention: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  CUtensorMap tm_kv;
  cuuint32_t box_kv[3] = {HD, KT, 1};
  r = enc(&tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box_kv, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "attention: cuTensorMapEncodeTiled (kv) failed with CUresult %d", (int)r);
  SETOK_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attn_tcgen05_hd64_kernel<false>), ATT_SMEM));
  dim3 grid(ceil_div(T, QT), heads, B);
  SETOK_CUDA_OK(launch_pdl(attn_tcgen05_hd64_kernel<false>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, tm_kv, static_cast<bf16*>(out), T, heads, C,
                           scale * 1.4426950408889634f, g_attn_trace, static_cast<const int32_t*>(nullptr)));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

// Varlen cross-attention on the tensor cores: q bf16 [B*Q, C], kv bf16 [kv_rows (capacity), 2C] = [k | v], query rows of image b
// see kv rows [offsets[b], offsets[b+1]); heads of 64.  Rows of a 64-key chunk past the image's last key belong to the next image
// (or to the zeroed tail, see launch_zero_tail_rows) and are masked: they never reach the softmax, and P = 0 meets finite V.
int launch_cross_attention_tcgen05(const void* q, const void* kv, void* out, int B, int Q, int C, int heads, float scale,
                                   const int32_t* offsets, int kv_rows, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SETOK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tm, tm_kv;
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint64_t qdim[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(Q), static_cast<cuuint64_t>(B)};
  cuuint64_t qstr[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(Q) * C * 2};
  cuuint32_t box[3] = {HD, QT, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(q), qdim, qstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "cross attention: cuTensorMapEncodeTiled (q) failed with CUresult %d", (int)r);
  cuuint64_t kdim[3] = {static_cast<cuuint64_t>(2 * C), static_cast<cuuint64_t>(kv_rows), 1};
  cuuint64_t kstr[2] = {static_cast<cuuint64_t>(2 * C) * 2, static_cast<cuuint64_t>(kv_rows) * 2 * C * 2};
  cuuint32_t box_kv[3] = {HD, KT, 1};
  r = enc(&tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(kv), kdim, kstr, box_kv, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "cross attention: cuTensorMapEncodeTiled (kv) failed with CUresult %d", (int)r);
  SETOK_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attn_tcgen05_hd64_kernel<true>), ATT_SMEM));
  dim3 grid(ceil_div(Q, QT), heads, B);
  SETOK_CUDA_OK(launch_pdl(attn_tcgen05_hd64_kernel<true>, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tm, tm_kv, static_cast<bf16*>(out), Q, heads, C,
                           scale * 1.4426950408889634f, static_cast<long long*>(nullptr), offsets));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" void setok_debug_set_attention_trace(long long* device_buf) { setok::g_attn_trace = device_buf; }
