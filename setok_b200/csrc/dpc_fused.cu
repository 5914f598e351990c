// dpc_fused.cu — position-embedding add + DPC-kNN clustering (reference src/model/setok/tokenizer.py:78-121,
// :164-169) as ONE persistent kernel for images of N <= 256 tokens: the features are read from HBM exactly once and
// the distance matrix never leaves the SM.
//
// One CTA per SM walks images b = blockIdx.x, blockIdx.x + gridDim.x, ...   288 threads:
//   warps 0..7  "row" warps.  (1) convert: read feats (+ pos, L2-resident), write x_pos (optional), split x = hi + lo
//               (bf16 + bf16, |x - hi - lo| <= 2^-17 |x|) straight into the 128B-swizzled K-major operand layout of
//               the tensor cores (3-stage ring, 64 channels per stage), accumulate the fp32 row norms.
//               (2) select: thread i owns token i = TMEM lane i of the Gram accumulator: G -> D in place (the matmul
//               form torch.cdist takes), the k smallest of its row by a register-resident bitonic top-k, density,
//               parent distance with the column-indexed row-max fill (tokenizer.py:98-99), score, centres
//               (threshold or top-min_cluster_num fallback), ordered compaction, nearest-centre labels.
//   warp 8      one thread issues tcgen05.mma: G = hi hi^T + hi lo^T + lo hi^T (+ lo lo^T), M = 128 per accumulator
//               (two accumulators = all 512 TMEM columns for N = 256), N = round_up(N, 16), fp32 accumulation.
// The A and the B operand of every MMA are the SAME shared-memory tile (G = X X^T), so each element is staged once.
// Every float operation the reference performs as a separate rounding step uses explicit round-to-nearest intrinsics.
//
// Label assignment reads D[j][c] (token j's own TMEM row at the centre columns) where the reference reads D[c][j]:
// the two differ only by the fp32 accumulation order inside the tensor core (~1e-7 relative), far below the 2e-4
// decision margin under which the reference's own cdist rounding decides.
#include "common.cuh"

#include <cmath>
#include <mutex>
#include <set>

namespace setok {
int g_dpc_fused = 5;   // 0: multi-kernel path only; 1: fused, 4-term split; 2: fused, 3-term split (lo.lo dropped); +4: norms from the Gram diagonal
namespace {

constexpr int FZ_BK = 64;
#ifndef SETOK_FZ_STAGES
#define SETOK_FZ_STAGES 3
#endif
constexpr int FZ_STAGES = SETOK_FZ_STAGES;
constexpr int FZ_ROWS = 256;
constexpr int FZ_TILE_BYTES = FZ_ROWS * 128;          // 32 KiB: 256 rows x 64 bf16
constexpr int FZ_STAGE_BYTES = 2 * FZ_TILE_BYTES;     // hi tile + lo tile
constexpr int FZ_ROW_WARPS = 8;
constexpr int FZ_ROW_THREADS = 32 * FZ_ROW_WARPS;
constexpr int FZ_THREADS = FZ_ROW_THREADS + 32;
constexpr int FZ_OFF_ARR = FZ_STAGES * FZ_STAGE_BYTES;
// float sqn[256], dens[256], rmax[256], score[256], maskv[256]; int cidx[256]; int wcount[8]; uint cmask[8]; float red[8]
constexpr int FZ_ARR_BYTES = 6 * 256 * 4 + 3 * 8 * 4;
constexpr int FZ_OFF_BAR = FZ_OFF_ARR + FZ_ARR_BYTES;
constexpr int FZ_NUM_BARS = 2 * FZ_STAGES + 2;
constexpr int FZ_SMEM_BYTES = FZ_OFF_BAR + FZ_NUM_BARS * 8 + 16 + 1024;

struct FusedDev {
  const void* feats;
  const float* pos;
  const float* noise;
  const float* token_mask;
  float* x_pos;
  int64_t* idx_cluster;
  float* score;
  int64_t* index_down;
  int32_t* num_clusters;
  int B, N, C, k, min_cluster_num;
  float threshold, sqrtC, inv_sqrtC;   // inv_sqrtC > 0 when sqrt(C) is a power of two (x / 2^e == x * 2^-e exactly)
  int feat_bf16, diag_norm;
};

// 16-byte global load of streamed data
__device__ __forceinline__ uint4 ldg_stream(const void* ptr) {
  uint4 v;
#if defined(SETOK_FZ_LD_NA)
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
#else
  v = *reinterpret_cast<const uint4*>(ptr);
#endif
  return v;
}

__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, %0;" ::"n"(FZ_ROW_THREADS) : "memory"); }

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void cas(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo; b = hi;
}
template <int n>
__device__ __forceinline__ void bitonic_sort_asc(float (&a)[n]) {
#pragma unroll
  for (int k = 2; k <= n; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) cas(a[i], a[l]); else cas(a[l], a[i]);
        }
      }
}
template <int n>
__device__ __forceinline__ void bitonic_merge_asc(float (&a)[n]) {   // a bitonic -> ascending
#pragma unroll
  for (int j = n >> 1; j > 0; j >>= 1)
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const int l = i ^ j;
      if (l > i) cas(a[i], a[l]);
    }
}

// One pass of thread i over its row of the accumulator, 16 columns at a time.
//   CONVERT: G -> D = sqrt(max(n_i + n_j - 2 g, 0)) / sqrt(C), written back to tensor memory
//   SELECT : keep the KSEL smallest (masked columns read as `fill`) in `best`, ascending
// rowmax accumulates the row maximum of what the pass saw (masked view when SELECT, raw D otherwise).
template <int KSEL, bool CONVERT, bool SELECT>
__device__ __forceinline__ void row_pass(uint32_t trow, int N, float ni, const float* sqn_s, const float* mask_s, float fill,
                                         float sqrtC, float inv_sqrtC, float (&best)[KSEL], float& rowmax) {
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(trow + c0, r);
    tmem_ld_wait();
    float d[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int j = c0 + t;
      float dd;
      if (CONVERT) {
        const float d2 = fmaf(-2.0f, __uint_as_float(r[t]), __fadd_rn(ni, sqn_s[j]));
        dd = sqrtf(fmaxf(d2, 0.f));
        dd = inv_sqrtC > 0.f ? __fmul_rn(dd, inv_sqrtC) : __fdiv_rn(dd, sqrtC);
        r[t] = __float_as_uint(dd);
      } else {
        dd = __uint_as_float(r[t]);
      }
      if (mask_s != nullptr && SELECT && !(mask_s[j] > 0.f)) dd = fill;
      if (j < N) rowmax = fmaxf(rowmax, dd); else dd = INFINITY;
      d[t] = dd;
    }
    if (CONVERT) tmem_st_32x32b_x16(trow + c0, r);
    if (SELECT) {
      bitonic_sort_asc<16>(d);
#pragma unroll
      for (int t = 0; t < 16; ++t) best[KSEL - 16 + t] = fminf(best[KSEL - 16 + t], d[15 - t]);
      bitonic_merge_asc<KSEL>(best);
    }
  }
  if (CONVERT) tmem_st_wait();
}

template <int KSEL, int TERMS, bool FBF16>
__global__ void __launch_bounds__(FZ_THREADS, 1) dpc_fused_kernel(FusedDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float* sqn_s = reinterpret_cast<float*>(smem + FZ_OFF_ARR);
  float* dens_s = sqn_s + 256;
  float* rmax_s = dens_s + 256;
  float* score_s = rmax_s + 256;
  float* maskv_s = score_s + 256;
  int* cidx_s = reinterpret_cast<int*>(maskv_s + 256);
  int* wcount_s = cidx_s + 256;
  uint32_t* cmask_s = reinterpret_cast<uint32_t*>(wcount_s + 8);
  float* red_s = reinterpret_cast<float*>(cmask_s + 8);

  const uint32_t bar0 = base + FZ_OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (FZ_STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * FZ_STAGES), tempty_bar = bar0 + 8u * (2 * FZ_STAGES + 1);
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + FZ_OFF_BAR + FZ_NUM_BARS * 8);

  if (threadIdx.x == 0) {
    for (int s = 0; s < FZ_STAGES; ++s) { mbar_init(full_bar(s), FZ_ROW_WARPS); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, FZ_ROW_WARPS);
    fence_mbar_init();
  }
  if (warp == FZ_ROW_WARPS) tmem_alloc<512>(base + FZ_OFF_BAR + FZ_NUM_BARS * 8);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int N = p.N, C = p.C;
  const int MT = N > 128 ? 2 : 1;                       // 128-row accumulators in use
  const int Npad = N < 16 ? 16 : ((N + 15) & ~15);      // MMA N
  const int k_blocks = (C + FZ_BK - 1) / FZ_BK;

  if (warp == FZ_ROW_WARPS) {
    // ------------------------------------------------ MMA issuer ------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, Npad);
      int stage = 0; uint32_t phase = 0, it = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        mbar_wait(tempty_bar, (it & 1u) ^ 1u);           // the row warps are done with the previous image's D
        tcgen05_fence_after();
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t hi = base + stage * FZ_STAGE_BYTES, lo = hi + FZ_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < FZ_BK / 16; ++k) {
            const uint64_t b_hi = umma_desc_k_sw128(hi + k * 32), b_lo = umma_desc_k_sw128(lo + k * 32);
            for (int m = 0; m < MT; ++m) {
              const uint64_t a_hi = umma_desc_k_sw128(hi + m * 16384 + k * 32), a_lo = umma_desc_k_sw128(lo + m * 16384 + k * 32);
              const uint32_t d = tmem_base + static_cast<uint32_t>(m * 256);
              umma_f16(d, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_f16(d, a_hi, b_lo, idesc, 1u);
              umma_f16(d, a_lo, b_hi, idesc, 1u);
              if (TERMS == 4) umma_f16(d, a_lo, b_lo, idesc, 1u);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == FZ_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else {
    // ------------------------------------------------ row warps -------------------------------------------------
    const int i = warp * 32 + lane;                      // this thread's token in the select phase
    const bool active = i < N;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) * 256);
    const int csub = lane & 7, rsub = lane >> 3;
    const int passes = MT * 4;                           // 32 rows per pass
    const bool has_mask = p.token_mask != nullptr;
    int stage = 0; uint32_t phase = 0, it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
      // ---- (1) convert: x = feats + pos -> hi/lo operand tiles, row norms ----
      // Rounds of 4 rows x 8 channels per thread, software-pipelined: the loads of round r+1 are in flight while
      // round r is split and stored.  RPK rounds fill one 64-channel stage.
      const long long img = static_cast<long long>(b) * N * C;
      const int RPK = MT;                                // rounds per k-block (32 rows x 4 per round)
      const int R = k_blocks * RPK;
      float sqA[4] = {0.f, 0.f, 0.f, 0.f}, sqB[4] = {0.f, 0.f, 0.f, 0.f};
      // loads are unconditional (row / channel clamped into the image) so that nothing but the load itself touches the
      // buffer registers between issue and first use; out-of-range units are zeroed when they are consumed
      auto load_round = [&](int rr, uint4 (&fa)[4], uint4 (&fb)[4], float4 (&pa)[4], float4 (&pb)[4]) {
        rr = rr < R ? rr : R - 1;
        const int kb = RPK == 2 ? (rr >> 1) : rr, half = RPK == 2 ? (rr & 1) : 0;
        int ch = kb * FZ_BK + csub * 8;
        ch = ch < C ? ch : C - 8;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int r = (half * 4 + u) * 32 + warp * 4 + rsub;
          r = r < N ? r : N - 1;
          const long long e = img + static_cast<long long>(r) * C + ch;
          if (FBF16) {
            fa[u] = ldg_stream(static_cast<const bf16*>(p.feats) + e);
          } else {
            fa[u] = ldg_stream(static_cast<const float*>(p.feats) + e);
            fb[u] = ldg_stream(static_cast<const float*>(p.feats) + e + 4);
          }
          pa[u] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<long long>(r) * C + ch));
          pb[u] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<long long>(r) * C + ch + 4));
        }
      };
      auto process_round = [&](int rr, const uint4 (&fa)[4], const uint4 (&fb)[4], const float4 (&pa)[4], const float4 (&pb)[4], float (&sq)[4]) {
        if (rr >= R) return;
        const int kb = RPK == 2 ? (rr >> 1) : rr, half = RPK == 2 ? (rr & 1) : 0;
        const int ch = kb * FZ_BK + csub * 8;
        if (half == 0) mbar_wait(empty_bar(stage), phase ^ 1u);
        uint8_t* hi_t = smem + stage * FZ_STAGE_BYTES;
        uint8_t* lo_t = hi_t + FZ_TILE_BYTES;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = (half * 4 + u) * 32 + warp * 4 + rsub;
          const bool valid = r < N && ch < C;
          float v[8];
          if (FBF16) {
            const float2 a = unpack_bf16x2(fa[u].x), c2 = unpack_bf16x2(fa[u].y), e2 = unpack_bf16x2(fa[u].z), g2 = unpack_bf16x2(fa[u].w);
            v[0] = a.x; v[1] = a.y; v[2] = c2.x; v[3] = c2.y; v[4] = e2.x; v[5] = e2.y; v[6] = g2.x; v[7] = g2.y;
          } else {
            v[0] = __uint_as_float(fa[u].x); v[1] = __uint_as_float(fa[u].y); v[2] = __uint_as_float(fa[u].z); v[3] = __uint_as_float(fa[u].w);
            v[4] = __uint_as_float(fb[u].x); v[5] = __uint_as_float(fb[u].y); v[6] = __uint_as_float(fb[u].z); v[7] = __uint_as_float(fb[u].w);
          }
          v[0] = __fadd_rn(v[0], pa[u].x); v[1] = __fadd_rn(v[1], pa[u].y); v[2] = __fadd_rn(v[2], pa[u].z); v[3] = __fadd_rn(v[3], pa[u].w);
          v[4] = __fadd_rn(v[4], pb[u].x); v[5] = __fadd_rn(v[5], pb[u].y); v[6] = __fadd_rn(v[6], pb[u].z); v[7] = __fadd_rn(v[7], pb[u].w);
          if (!valid) {
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = 0.f;
          } else if (p.x_pos != nullptr) {
            float* xo = p.x_pos + img + static_cast<long long>(r) * C + ch;
            *reinterpret_cast<float4*>(xo) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(xo + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          uint32_t h[4], l[4];
          float s = sq[u];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float x0 = v[2 * t], x1 = v[2 * t + 1];
            h[t] = pack_bf16x2(x0, x1);
            const float2 hf = unpack_bf16x2(h[t]);
            l[t] = pack_bf16x2(x0 - hf.x, x1 - hf.y);
            s = fmaf(x0, x0, s);
            s = fmaf(x1, x1, s);
          }
          sq[u] = s;
          const uint32_t off = static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(csub ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(hi_t + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(lo_t + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        if (half == RPK - 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar(stage));
          if (++stage == FZ_STAGES) { stage = 0; phase ^= 1u; }
        }
      };
      {
        uint4 faA[4], fbA[4], faB[4], fbB[4];
        float4 paA[4], pbA[4], paB[4], pbB[4];
        load_round(0, faA, fbA, paA, pbA);
        for (int q = 0; 2 * q < R; ++q) {
          load_round(2 * q + 1, faB, fbB, paB, pbB);
          process_round(2 * q, faA, fbA, paA, pbA, sqA);
          load_round(2 * q + 2, faA, fbA, paA, pbA);
          process_round(2 * q + 1, faB, fbB, paB, pbB, sqB);
        }
      }
      // row norms: RPK == 2: slot A = rows of passes 0..3, slot B = passes 4..7; RPK == 1: both slots hold the same rows
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float sa = sqA[u], sb = sqB[u];
        if (RPK == 1) { sa += sb; sb = 0.f; }
        sa += __shfl_xor_sync(0xffffffffu, sa, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 1);
        sa += __shfl_xor_sync(0xffffffffu, sa, 2); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
        sa += __shfl_xor_sync(0xffffffffu, sa, 4); sb += __shfl_xor_sync(0xffffffffu, sb, 4);
        if (csub == 0) {
          sqn_s[u * 32 + warp * 4 + rsub] = sa;
          if (RPK == 2) sqn_s[(4 + u) * 32 + warp * 4 + rsub] = sb;
        }
      }
      if (has_mask && active) maskv_s[i] = p.token_mask[static_cast<long long>(b) * N + i];
      bar_rows();

      // ---- (2) select on the accumulator ----
      mbar_wait(tfull_bar, it & 1u);
      tcgen05_fence_after();
      if (p.diag_norm) {
        // norms from the Gram diagonal: n_i + n_j - 2 g_ij is then evaluated on ONE consistently rounded matrix, so the
        // accumulation error of the tensor core cancels for near neighbours (where the difference is small) and the
        // diagonal is exactly zero
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + static_cast<uint32_t>(warp * 32), r);
        tmem_ld_wait();
        float gii = 0.f;
#pragma unroll
        for (int t = 0; t < 32; ++t) gii = t == lane ? __uint_as_float(r[t]) : gii;
        sqn_s[i] = gii;
        bar_rows();
      }
      const float* mask_s = has_mask ? maskv_s : nullptr;
      const float ni = active ? sqn_s[i] : 0.f;
      float best[KSEL];
#pragma unroll
      for (int t = 0; t < KSEL; ++t) best[t] = INFINITY;
      float rowmax = -INFINITY, fill = 0.f;
      if (!has_mask) {
        row_pass<KSEL, true, true>(trow, N, ni, sqn_s, nullptr, 0.f, p.sqrtC, p.inv_sqrtC, best, rowmax);
      } else {
        // tokenizer.py:84-86: masked columns are pushed to (global max + 1)
        float gmax = -INFINITY;
        row_pass<KSEL, true, false>(trow, N, ni, sqn_s, nullptr, 0.f, p.sqrtC, p.inv_sqrtC, best, gmax);
        if (!active) gmax = -INFINITY;
        gmax = warp_max(gmax);
        if (lane == 0) red_s[warp] = gmax;
        bar_rows();
        gmax = red_s[0];
#pragma unroll
        for (int w = 1; w < FZ_ROW_WARPS; ++w) gmax = fmaxf(gmax, red_s[w]);
        fill = __fadd_rn(gmax, 1.0f);
        row_pass<KSEL, false, true>(trow, N, ni, sqn_s, mask_s, fill, p.sqrtC, p.inv_sqrtC, best, rowmax);
      }
      // density (tokenizer.py:88-94): exp(-mean of the k smallest squared distances) + 1e-6 * noise
      float di = 0.f;
      {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < KSEL; ++t)
          if (t < p.k) s = __fadd_rn(s, __fmul_rn(best[t], best[t]));
        const float mean = __fdiv_rn(s, static_cast<float>(p.k));
        if (active) {
          di = __fadd_rn(expf(-mean), __fmul_rn(p.noise[static_cast<long long>(b) * N + i], 1e-6f));
          if (has_mask) di = __fmul_rn(di, maskv_s[i] > 0.f ? 1.0f : 0.0f);
          dens_s[i] = di;
          rmax_s[i] = rowmax;
        }
      }
      bar_rows();

      // parent distance (tokenizer.py:96-99): where density[j] > density[i] the distance, elsewhere rowmax[j]
      float pd = INFINITY;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(trow + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int j = c0 + t;
          if (j < N) {
            float dd = __uint_as_float(r[t]);
            if (has_mask && !(maskv_s[j] > 0.f)) dd = fill;
            pd = fminf(pd, dens_s[j] > di ? dd : rmax_s[j]);
          }
        }
      }
      const float sc = __fmul_rn(pd, di);
      if (active) {
        score_s[i] = sc;
        p.score[static_cast<long long>(b) * N + i] = sc;
      }
      // centres = {score > threshold}; none -> the min_cluster_num best scores (tokenizer.py:103-107)
      bool f = active && sc > p.threshold;
      uint32_t bal = __ballot_sync(0xffffffffu, f);
      if (lane == 0) wcount_s[warp] = __popc(bal);
      bar_rows();                                         // also publishes score_s
      int tot = 0;
#pragma unroll
      for (int w = 0; w < FZ_ROW_WARPS; ++w) tot += wcount_s[w];
      if (tot == 0) {
        int rank = 0;
        for (int j = 0; j < N; ++j) {
          const float sj = score_s[j];
          rank += (sj > sc || (sj == sc && j < i)) ? 1 : 0;
        }
        f = active && rank < p.min_cluster_num;
        bal = __ballot_sync(0xffffffffu, f);
        bar_rows();                                       // everyone has read the all-zero counts
        if (lane == 0) wcount_s[warp] = __popc(bal);
        bar_rows();
      }
      int woff = 0, K = 0;
#pragma unroll
      for (int w = 0; w < FZ_ROW_WARPS; ++w) { const int c = wcount_s[w]; if (w < warp) woff += c; K += c; }
      const int mypos = woff + __popc(bal & ((1u << lane) - 1u));
      if (f) cidx_s[mypos] = i;
      if (lane == 0) cmask_s[warp] = bal;
      bar_rows();

      // nearest centre, first minimum in centre order (tokenizer.py:111-113); centres own their label (:117-119)
      int label = 0;
      {
        float bd = INFINITY;
        int cnt = 0;
        const bool masked_i = has_mask && active && !(maskv_s[i] > 0.f);
        for (int c0 = 0; c0 < N; c0 += 16) {
          const uint32_t word = (cmask_s[c0 >> 5] >> (c0 & 16)) & 0xFFFFu;
          if (word == 0u) continue;
          uint32_t r[16];
          tmem_ld_32x32b_x16(trow + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            if ((word >> t) & 1u) {
              const float dd = masked_i ? fill : __uint_as_float(r[t]);
              if (dd < bd) { bd = dd; label = cnt; }
              ++cnt;
            }
          }
        }
        if (f) label = mypos;
      }
      if (active) {
        p.idx_cluster[static_cast<long long>(b) * N + i] = static_cast<int64_t>(label);
        p.index_down[static_cast<long long>(b) * N + i] = i < K ? static_cast<int64_t>(cidx_s[i]) : -1;
      }
      if (warp == 0 && lane == 0) p.num_clusters[b] = K;
      // hand the accumulator back to the MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      bar_rows();                                         // smem arrays are rewritten by the next image's convert/select
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == FZ_ROW_WARPS) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool dpc_fused_supported(int N, int C, int k) { return g_dpc_fused != 0 && N >= 1 && N <= 256 && C % 8 == 0 && k <= 64; }

// One launch for the whole batch; offsets are scanned by the caller (offsets_scan_kernel).
int launch_dpc_fused(const void* feats, int feat_dtype, const float* pos, const float* noise, const float* token_mask, int B, int N, int C,
                     int k, float threshold, int min_cluster_num, float* x_pos, int64_t* idx_cluster, float* score, int64_t* index_down,
                     int32_t* num_clusters, cudaStream_t stream) {
  FusedDev p;
  p.feats = feats; p.pos = pos; p.noise = noise; p.token_mask = token_mask; p.x_pos = x_pos; p.idx_cluster = idx_cluster;
  p.score = score; p.index_down = index_down; p.num_clusters = num_clusters;
  p.B = B; p.N = N; p.C = C; p.k = k; p.min_cluster_num = min_cluster_num; p.threshold = threshold;
  p.sqrtC = static_cast<float>(std::sqrt(static_cast<double>(C)));
  int e = 0;
  const float m = std::frexp(p.sqrtC, &e);
  p.inv_sqrtC = (m == 0.5f) ? 1.0f / p.sqrtC : 0.f;
  p.feat_bf16 = feat_dtype == SETOK_BF16 ? 1 : 0;
  p.diag_norm = (g_dpc_fused & 4) ? 1 : 0;
  using KernelFn = void (*)(FusedDev);
  const bool t3 = (g_dpc_fused & 3) == 2;
  const bool fb = p.feat_bf16 != 0;
#define SETOK_FZ_PICK(KS) (t3 ? (fb ? dpc_fused_kernel<KS, 3, true> : dpc_fused_kernel<KS, 3, false>) \
                              : (fb ? dpc_fused_kernel<KS, 4, true> : dpc_fused_kernel<KS, 4, false>))
  KernelFn fn = k <= 16 ? SETOK_FZ_PICK(16) : (k <= 32 ? SETOK_FZ_PICK(32) : SETOK_FZ_PICK(64));
#undef SETOK_FZ_PICK
  static std::mutex attr_mu;
  static std::set<KernelFn> attr_done;
  {
    std::lock_guard<std::mutex> lk(attr_mu);
    if (!attr_done.count(fn)) {
      SETOK_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM_BYTES));
      attr_done.insert(fn);
    }
  }
  const int grid = B < num_sms() ? B : num_sms();
  fn<<<grid, FZ_THREADS, FZ_SMEM_BYTES, stream>>>(p);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" void setok_debug_set_dpc_fused(int mode) { setok::g_dpc_fused = mode; }
