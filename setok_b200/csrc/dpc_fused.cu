// dpc_fused.cu — position-embedding add + DPC-kNN clustering (reference src/model/setok/tokenizer.py:78-121,
// :164-169) as ONE persistent kernel for images of N <= 256 tokens: the features are read from HBM exactly once and
// the distance matrix never leaves the SM.
//
// One CTA per SM walks images b = blockIdx.x, blockIdx.x + gridDim.x, ...   320 threads:
//   warp 9      TMA producer: raw feature tiles (256 rows x 128 B: 32 fp32 or 64 bf16 channels) into a 2-stage ring.
//               Bulk copies bypass L1 (which the 200 KB of shared memory leave almost no room for) and keep 64 KB in
//               flight per SM.
//   warps 0..7  "row" warps.  (1) convert: raw tile (+ pos, optional) -> x; x_pos out (optional); split x = hi + lo
//               (bf16 + bf16, |x - hi - lo| <= 2^-17 |x|) straight into the 128B-swizzled K-major operand layout of
//               the tensor cores (2-stage ring, 64 channels per stage).
//               (2) select: thread i owns token i = TMEM lane i of the Gram accumulator: G -> D in place (the matmul
//               form torch.cdist takes, row norms from the Gram diagonal), the k smallest of its row by a
//               register-resident bitonic top-k, density, parent distance with the column-indexed row-max fill
//               (tokenizer.py:98-99), score, centres (threshold or top-min_cluster_num fallback), ordered compaction,
//               nearest-centre labels.
//   warp 8      one thread issues tcgen05.mma: G = hi hi^T + hi lo^T + lo hi^T (+ lo lo^T), M = 128 per accumulator
//               (two accumulators = all 512 TMEM columns for N = 256), N = round_up(N, 16), fp32 accumulation.
// The A and the B operand of every MMA are the SAME shared-memory tile (G = X X^T), so each element is staged once.
// Every float operation the reference performs as a separate rounding step uses explicit round-to-nearest intrinsics.
//
// Numerics.  (a) The row norms n_i are the Gram diagonal g_ii, so n_i + n_j - 2 g_ij is evaluated on one consistently
// rounded matrix: the accumulation error of the tensor core cancels where the difference is small (near neighbours) and
// D_ii is exactly 0.  Against float64 ground truth this is ~3x closer than separately summed fp32 norms
// (tools/bench_cluster.py).  (b) Label assignment reads D[j][c] (token j's own TMEM row at the centre columns) where
// the reference reads D[c][j]: the two differ only by the fp32 accumulation order inside the tensor core (~1e-7
// relative), far below the 2e-4 decision margin under which the reference's own cdist rounding decides
// (DESIGN.md 2).
#include "common.cuh"

#include <cmath>

namespace setok {
// 0: multi-kernel path only; 1: fused, IEEE-half hi/lo split, 3 terms (default); 2: fused, bf16 split, 4 terms; 3: fused, bf16
// split, 3 terms (lo.lo dropped)
int g_dpc_fused = 1;
namespace {

constexpr int FZ_BK = 64;                              // channels per operand stage
constexpr int FZ_OP_STAGES = 2;
constexpr int FZ_RAW_STAGES = 3;
constexpr int FZ_TILE_BYTES = 256 * 128;               // 32 KiB: 256 rows x 128 B (64 bf16 | 32 fp32)
constexpr int FZ_OP_STAGE_BYTES = 2 * FZ_TILE_BYTES;   // hi tile + lo tile
constexpr int FZ_ROW_WARPS = 8;
constexpr int FZ_GROUP_WARPS = 4;                      // the row warps convert in two groups that take alternate half-stages
constexpr int FZ_ROW_THREADS = 32 * FZ_ROW_WARPS;
constexpr int FZ_W_MMA = FZ_ROW_WARPS, FZ_W_TMA = FZ_ROW_WARPS + 1;
constexpr int FZ_THREADS = FZ_ROW_THREADS + 64;
constexpr int FZ_OFF_RAW = FZ_OP_STAGES * FZ_OP_STAGE_BYTES;
// The select phase's arrays (float sqn[256], dens[256], rmax[256], score[256], maskv[256]; int cidx[256]; int wcount[8]; uint
// cmask[8]; float red[8]) live on top of the operand ring: they are touched only between the accumulator-complete barrier
// (every MMA that read the ring has retired) and the select phase's last barrier (the next image's converters start after it).
constexpr int FZ_OFF_ARR = 0;
constexpr int FZ_ARR_BYTES = 6 * 256 * 4 + 3 * 8 * 4;
static_assert(FZ_ARR_BYTES <= FZ_OP_STAGE_BYTES, "select arrays must fit the operand ring");
constexpr int FZ_OFF_BAR = FZ_OFF_RAW + FZ_RAW_STAGES * FZ_TILE_BYTES;
constexpr int FZ_HALVES = 2 * FZ_OP_STAGES;           // half-stages (32 channels): the unit of the operand pipeline
constexpr int FZ_NUM_BARS = 2 * FZ_HALVES + 2 * FZ_RAW_STAGES + 2;
constexpr int FZ_SMEM_BYTES = FZ_OFF_BAR + FZ_NUM_BARS * 8 + 16 + 1024;
static_assert(FZ_SMEM_BYTES <= 232448, "dpc_fused: shared memory over the 227 KiB a CTA may use");

struct FusedDev {
  const float* pos;
  const float* noise;
  const float* token_mask;
  float* x_pos;
  int64_t* idx_cluster;
  float* score;
  int64_t* index_down;
  int32_t* num_clusters;
  int B, N, C, k, min_cluster_num, terms, f16;
  int dbg;                             // timing experiments (setok_debug_set_dpc_fused bits 5, 6): 1 converters touch no shared memory, 2 + no TMA loads
  float threshold, sqrtC, inv_sqrtC;   // inv_sqrtC > 0 when sqrt(C) is a power of two (x / 2^e == x * 2^-e exactly)
};

#ifdef SETOK_FZ_TRACE
__device__ unsigned long long* g_fz_trace = nullptr;   // [role][256] timestamps (ns) of CTA 0, first image
#define FZ_TRACE(role, seq) do { if (blockIdx.x == 0 && it == 0 && g_fz_trace != nullptr && (seq) < 256) g_fz_trace[(role) * 256 + (seq)] = globaltimer_ns(); } while (0)
#else
#define FZ_TRACE(role, seq) do { } while (0)
#endif

__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, %0;" ::"n"(FZ_ROW_THREADS) : "memory"); }

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

// 32 columns [c0, c0+32) of thread i's row, already in registers.
//   CONVERT: G -> D = sqrt(max(n_i + n_j - 2 g, 0)) / sqrt(C), written back to tensor memory
//   SELECT : keep the KSEL smallest (masked columns read as `fill`) in `best`, ascending
// rowmax accumulates the row maximum of what the pass saw (masked view when SELECT, raw D otherwise).
template <int KSEL, bool CONVERT, bool SELECT>
__device__ __forceinline__ void row_chunk(uint32_t (&r)[32], uint32_t taddr, int c0, int N, float ni, const float* sqn_s, const float* mask_s,
                                          float fill, float sqrtC, float inv_sqrtC, float (&best)[KSEL], float& rowmax) {
  float d[32];
#pragma unroll
  for (int t4 = 0; t4 < 32; t4 += 4) {
    // norms / mask of four columns per shared-memory load (all lanes read the same address: broadcast)
    float nj[4] = {0.f, 0.f, 0.f, 0.f}, mk[4] = {1.f, 1.f, 1.f, 1.f};
    if (CONVERT) { const float4 q = *reinterpret_cast<const float4*>(sqn_s + c0 + t4); nj[0] = q.x; nj[1] = q.y; nj[2] = q.z; nj[3] = q.w; }
    if (SELECT && mask_s != nullptr) { const float4 q = *reinterpret_cast<const float4*>(mask_s + c0 + t4); mk[0] = q.x; mk[1] = q.y; mk[2] = q.z; mk[3] = q.w; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int t = t4 + e, j = c0 + t;
      float dd;
      if (CONVERT) {
        const float d2 = fmaf(-2.0f, __uint_as_float(r[t]), __fadd_rn(ni, nj[e]));
        // MUFU square root (2^-22 relative): two orders of magnitude below the rounding noise the Gram itself carries
        asm("sqrt.approx.f32 %0, %1;" : "=f"(dd) : "f"(fmaxf(d2, 0.f)));
        dd = inv_sqrtC > 0.f ? __fmul_rn(dd, inv_sqrtC) : __fdiv_rn(dd, sqrtC);
        r[t] = __float_as_uint(dd);
      } else {
        dd = __uint_as_float(r[t]);
      }
      if (SELECT && mask_s != nullptr && !(mk[e] > 0.f)) dd = fill;
      if (j < N) rowmax = fmaxf(rowmax, dd); else dd = INFINITY;
      d[t] = dd;
    }
  }
  if (CONVERT) tmem_st_32x32b_x32(taddr, r);
  if (SELECT) {
    float lo16[16], hi16[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) { lo16[t] = d[t]; hi16[t] = d[16 + t]; }
    bitonic_sort_asc<16>(lo16);              // two independent networks: the scheduler interleaves them
    bitonic_sort_asc<16>(hi16);
#pragma unroll
    for (int t = 0; t < 16; ++t) best[KSEL - 16 + t] = fminf(best[KSEL - 16 + t], lo16[15 - t]);
    bitonic_merge_asc<KSEL>(best);
#pragma unroll
    for (int t = 0; t < 16; ++t) best[KSEL - 16 + t] = fminf(best[KSEL - 16 + t], hi16[15 - t]);
    bitonic_merge_asc<KSEL>(best);
  }
}

// One pass of thread i over its row of the accumulator, 32 columns at a time.
template <int KSEL, bool CONVERT, bool SELECT>
__device__ __forceinline__ void row_pass(uint32_t trow, int N, float ni, const float* sqn_s, const float* mask_s, float fill,
                                         float sqrtC, float inv_sqrtC, float (&best)[KSEL], float& rowmax) {
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(trow + c0, r);
    tmem_ld_wait();
    row_chunk<KSEL, CONVERT, SELECT>(r, trow + c0, c0, N, ni, sqn_s, mask_s, fill, sqrtC, inv_sqrtC, best, rowmax);
  }
  if (CONVERT) tmem_st_wait();
}

template <int KSEL, bool FBF16>
__global__ void __launch_bounds__(FZ_THREADS, 1) dpc_fused_kernel(const __grid_constant__ CUtensorMap tmF, FusedDev p) {
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
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                                  // operand half-stage written (row warps -> MMA)
  auto empty_bar = [&](int s) { return bar0 + 8u * (FZ_HALVES + s); };                   // operand half-stage consumed (MMA -> row warps)
  auto rfull_bar = [&](int s) { return bar0 + 8u * (2 * FZ_HALVES + s); };               // raw tile landed (TMA -> row warps)
  auto rempty_bar = [&](int s) { return bar0 + 8u * (2 * FZ_HALVES + FZ_RAW_STAGES + s); };      // raw tile read (row warps -> TMA)
  const uint32_t tfull_bar = bar0 + 8u * (2 * FZ_HALVES + 2 * FZ_RAW_STAGES), tempty_bar = tfull_bar + 8u;
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + FZ_OFF_BAR + FZ_NUM_BARS * 8);

  if (threadIdx.x == 0) {
    // a half-stage is written by ONE converter group; a raw tile is read by one group (fp32: 32 channels = one half-stage) or
    // by both (bf16: 64 channels = two half-stages)
    for (int s = 0; s < FZ_HALVES; ++s) { mbar_init(full_bar(s), FZ_GROUP_WARPS); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < FZ_RAW_STAGES; ++s) { mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), FBF16 ? FZ_ROW_WARPS : FZ_GROUP_WARPS); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, FZ_ROW_WARPS);
    fence_mbar_init();
  }
  if (warp == FZ_W_TMA && lane == 0) tma_prefetch_desc(&tmF);
  if (warp == FZ_W_MMA) tmem_alloc<512>(base + FZ_OFF_BAR + FZ_NUM_BARS * 8);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int N = p.N, C = p.C;
  const int MT = N > 128 ? 2 : 1;                       // 128-row accumulators in use
  const int Npad = N < 16 ? 16 : ((N + 15) & ~15);      // MMA N
  const int k_blocks = (C + FZ_BK - 1) / FZ_BK;
  constexpr int RAW_CH = FBF16 ? 64 : 32;               // channels per raw tile (128 B rows)
  constexpr int RAW_PER_OP = FZ_BK / RAW_CH;            // raw tiles per operand stage
  const int raw_tiles = k_blocks * RAW_PER_OP;          // per image

  if (warp == FZ_W_TMA) {
    // ------------------------------------------------ TMA producer ----------------------------------------------
    if (lane == 0) {
      int rs = 0; uint32_t rphase = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int t = 0; t < raw_tiles; ++t) {
          mbar_wait(rempty_bar(rs), rphase ^ 1u);
#ifdef SETOK_FZ_TRACE
          { const unsigned it = (b == (int)blockIdx.x) ? 0u : 1u; FZ_TRACE(3, t); }
#endif
          if (p.dbg == 2) { mbar_arrive(rfull_bar(rs)); if (++rs == FZ_RAW_STAGES) { rs = 0; rphase ^= 1u; } continue; }
          mbar_arrive_expect_tx(rfull_bar(rs), FZ_TILE_BYTES);
          // 256-row box starting at the image's first row: for N < 256 the tail rows belong to the next image (or are
          // zero-filled past the end of the tensor) and are discarded by the converters
          tma_load_2d(&tmF, rfull_bar(rs), base + FZ_OFF_RAW + rs * FZ_TILE_BYTES, t * RAW_CH, b * N);
          if (++rs == FZ_RAW_STAGES) { rs = 0; rphase ^= 1u; }
        }
      }
    }
  } else if (warp == FZ_W_MMA) {
    // ------------------------------------------------ MMA issuer ------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = p.f16 ? umma_idesc_f16(128, Npad) : umma_idesc_bf16(128, Npad);
      int stage = 0; uint32_t phase = 0, it = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        mbar_wait(tempty_bar, (it & 1u) ^ 1u);           // the row warps are done with the previous image's D
        tcgen05_fence_after();
        FZ_TRACE(0, 0);
        for (int hk = 0; hk < 2 * k_blocks; ++hk) {
          // half-stage = 32 channels = two k16 steps: the pipeline between the converters and the tensor core is 4 deep
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          FZ_TRACE(0, 1 + 2 * hk);
          const uint32_t hi = base + (stage >> 1) * FZ_OP_STAGE_BYTES + (stage & 1) * 64, lo = hi + FZ_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t b_hi = umma_desc_k_sw128(hi + k * 32), b_lo = umma_desc_k_sw128(lo + k * 32);
            for (int m = 0; m < MT; ++m) {
              const uint64_t a_hi = umma_desc_k_sw128(hi + m * 16384 + k * 32), a_lo = umma_desc_k_sw128(lo + m * 16384 + k * 32);
              const uint32_t d = tmem_base + static_cast<uint32_t>(m * 256);
              umma_f16(d, a_hi, b_hi, idesc, (hk | k) != 0 ? 1u : 0u);
              umma_f16(d, a_hi, b_lo, idesc, 1u);
              umma_f16(d, a_lo, b_hi, idesc, 1u);
              if (p.terms == 4) umma_f16(d, a_lo, b_lo, idesc, 1u);
            }
          }
          umma_commit(empty_bar(stage));
          FZ_TRACE(0, 2 + 2 * hk);
          if (++stage == FZ_HALVES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else {
    // ------------------------------------------------ row warps -------------------------------------------------
    const int i = warp * 32 + lane;                      // this thread's token in the select phase
    const bool active = i < N;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) * 256);
    // convert-phase mapping: the 8 row warps form two groups of 4; group g takes the half-stages hs = g, g + 2, ... (32 channels
    // of all rows each), so that one group's chain (wait raw tile -> shared loads -> split -> operand stores -> proxy fence ->
    // publish) runs under the other's.  A unit = 8 channels of one row; 4 lanes cover a row's 32 channels, 8 rows per warp,
    // 32 rows per pass of a group, MT * 4 units per thread and half-stage.
    const int grp = warp >> 2, gw = warp & 3;
    const int c4 = lane & 3, rsub = lane >> 2;
    constexpr int PR = FZ_GROUP_WARPS * 8;               // rows per pass of one group: a multiple of 8
    const int units = MT * 128 / PR;                     // 4 | 8
    const int r0 = gw * 8 + rsub;                        // this thread's row in pass 0; (row & 7) is the same in every pass
    const int x7 = r0 & 7;
    const bool has_mask = p.token_mask != nullptr;
    const bool has_pos = p.pos != nullptr;
    const int HS = 2 * k_blocks;                         // half-stages per image (even: both groups take HS / 2 of them)
    uint32_t gh_base = 0, it = 0;                        // half-stages of this CTA's earlier images
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it, gh_base += static_cast<uint32_t>(HS)) {
      // ---- (1) convert: raw tile (+ pos) -> hi/lo operand tiles ----
      const long long img = static_cast<long long>(b) * N * C;
#pragma unroll 1
      for (int hs = grp; hs < HS; hs += 2) {
        const uint32_t gh = gh_base + static_cast<uint32_t>(hs);            // position in the CTA-wide half-stage sequence
        const int stage = static_cast<int>(gh & (FZ_HALVES - 1));
        const uint32_t phase = (gh / FZ_HALVES) & 1u;
        const uint32_t t = FBF16 ? (gh >> 1) : gh;                           // raw tile in the CTA-wide sequence
        const int rs = static_cast<int>(t % FZ_RAW_STAGES);
        const uint32_t rphase = (t / FZ_RAW_STAGES) & 1u;
        const int h = hs & 1;
        const int oc = h * 4 + c4;                       // 16-byte chunk of the operand row (8 halves)
        const int ch = (hs >> 1) * FZ_BK + oc * 8;
        // position embedding of this thread's units (L2-resident table): requested before the waits
        float4 pa[8], pb[8];
        if (has_pos) {
          const float* pp = p.pos + (ch < C ? ch : C - 8);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (u < units) {
              int r = r0 + u * PR;
              r = r < N ? r : N - 1;
              pa[u] = __ldg(reinterpret_cast<const float4*>(pp + r * C));
              pb[u] = __ldg(reinterpret_cast<const float4*>(pp + r * C + 4));
            }
          }
        }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (threadIdx.x == 0) FZ_TRACE(1, 3 * hs);
        mbar_wait(rfull_bar(rs), rphase);
        if (threadIdx.x == 0) FZ_TRACE(1, 3 * hs + 1);
        uint8_t* hi_t = smem + (stage >> 1) * FZ_OP_STAGE_BYTES + r0 * 128 + ((oc ^ x7) << 4);
        const uint8_t* raw_t = smem + FZ_OFF_RAW + rs * FZ_TILE_BYTES + r0 * 128;
#pragma unroll
        for (int ub = 0; ub < 8; ub += 4) {
          if (ub < units && p.dbg == 0) {
            // the shared-memory reads of four units first (the compiler cannot move them across the operand stores below)
            uint4 qa[4], qb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (FBF16) {
                qa[u] = *reinterpret_cast<const uint4*>(raw_t + (ub + u) * PR * 128 + ((oc ^ x7) << 4));
              } else {
                qa[u] = *reinterpret_cast<const uint4*>(raw_t + (ub + u) * PR * 128 + (((2 * c4) ^ x7) << 4));
                qb[u] = *reinterpret_cast<const uint4*>(raw_t + (ub + u) * PR * 128 + (((2 * c4 + 1) ^ x7) << 4));
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + (ub + u) * PR;
              float v[8];
              if (FBF16) {
                const float2 a = unpack_bf16x2(qa[u].x), c2 = unpack_bf16x2(qa[u].y), e2 = unpack_bf16x2(qa[u].z), g2 = unpack_bf16x2(qa[u].w);
                v[0] = a.x; v[1] = a.y; v[2] = c2.x; v[3] = c2.y; v[4] = e2.x; v[5] = e2.y; v[6] = g2.x; v[7] = g2.y;
              } else {
                v[0] = __uint_as_float(qa[u].x); v[1] = __uint_as_float(qa[u].y); v[2] = __uint_as_float(qa[u].z); v[3] = __uint_as_float(qa[u].w);
                v[4] = __uint_as_float(qb[u].x); v[5] = __uint_as_float(qb[u].y); v[6] = __uint_as_float(qb[u].z); v[7] = __uint_as_float(qb[u].w);
              }
              if (has_pos) {
                const float4 a4 = pa[ub + u], b4 = pb[ub + u];
                v[0] = __fadd_rn(v[0], a4.x); v[1] = __fadd_rn(v[1], a4.y); v[2] = __fadd_rn(v[2], a4.z); v[3] = __fadd_rn(v[3], a4.w);
                v[4] = __fadd_rn(v[4], b4.x); v[5] = __fadd_rn(v[5], b4.y); v[6] = __fadd_rn(v[6], b4.z); v[7] = __fadd_rn(v[7], b4.w);
              }
              if (!(r < N && ch < C)) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
              } else if (p.x_pos != nullptr) {
                float* xo = p.x_pos + img + static_cast<long long>(r) * C + ch;
                *reinterpret_cast<float4*>(xo) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(xo + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float x0 = v[2 * q], x1 = v[2 * q + 1];
                if (p.f16) {
                  hh[q] = pack_f16x2_sat(x0, x1);
                  const float2 hf = unpack_f16x2(hh[q]);
                  ll[q] = pack_f16x2_sat(x0 - hf.x, x1 - hf.y);
                } else {
                  hh[q] = pack_bf16x2(x0, x1);
                  const float2 hf = unpack_bf16x2(hh[q]);
                  ll[q] = pack_bf16x2(x0 - hf.x, x1 - hf.y);
                }
              }
              *reinterpret_cast<uint4*>(hi_t + (ub + u) * PR * 128) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
              *reinterpret_cast<uint4*>(hi_t + FZ_TILE_BYTES + (ub + u) * PR * 128) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
          }
        }
        // this warp is done with the raw tile (fp32: the tile was this group's alone; bf16: the other group reads its other half)
        __syncwarp();
        if (lane == 0) mbar_arrive(rempty_bar(rs));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
        if (threadIdx.x == 0) FZ_TRACE(1, 3 * hs + 2);
      }

      // ---- (2) select on the accumulator ----
      if (threadIdx.x == 0) FZ_TRACE(2, 0);
      mbar_wait(tfull_bar, it & 1u);
      tcgen05_fence_after();
      if (threadIdx.x == 0) FZ_TRACE(2, 1);
      if (has_mask && active) maskv_s[i] = p.token_mask[static_cast<long long>(b) * N + i];
      {
        // row norms = Gram diagonal (see the header)
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
      const float ni = sqn_s[i];
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
      if (threadIdx.x == 0) FZ_TRACE(2, 2);
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
      {
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(trow + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int t4 = 0; t4 < 32; t4 += 4) {
            const float4 dn = *reinterpret_cast<const float4*>(dens_s + c0 + t4);
            const float4 rm = *reinterpret_cast<const float4*>(rmax_s + c0 + t4);
            float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
            if (has_mask) mk = *reinterpret_cast<const float4*>(maskv_s + c0 + t4);
            const float dnv[4] = {dn.x, dn.y, dn.z, dn.w}, rmv[4] = {rm.x, rm.y, rm.z, rm.w}, mkv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c0 + t4 + e;
              if (j < N) {
                float dd = __uint_as_float(r[t4 + e]);
                if (has_mask && !(mkv[e] > 0.f)) dd = fill;
                pd = fminf(pd, dnv[e] > di ? dd : rmv[e]);
              }
            }
          }
        }
      }
      if (threadIdx.x == 0) FZ_TRACE(2, 3);
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
        for (int c0 = 0; c0 < N; c0 += 32) {
          const uint32_t word = cmask_s[c0 >> 5];
          if (word == 0u) continue;
          uint32_t r[32];
          tmem_ld_32x32b_x32(trow + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) {
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
      if (threadIdx.x == 0) FZ_TRACE(2, 4);
      bar_rows();                                         // smem arrays are rewritten by the next image's select
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == FZ_W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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

bool dpc_fused_supported(int N, int C, int k) { return g_dpc_fused != 0 && N >= 1 && N <= 256 && C % 8 == 0 && k <= 64; }

// One launch for the whole batch; offsets are scanned by the caller (offsets_scan_kernel).  pos may be NULL (features
// already position-embedded), x_pos may be NULL (not wanted).
int launch_dpc_fused(const void* feats, int feat_dtype, const float* pos, const float* noise, const float* token_mask, int B, int N, int C,
                     int k, float threshold, int min_cluster_num, float* x_pos, int64_t* idx_cluster, float* score, int64_t* index_down,
                     int32_t* num_clusters, cudaStream_t stream) {
  FusedDev p;
  p.pos = pos; p.noise = noise; p.token_mask = token_mask; p.x_pos = x_pos; p.idx_cluster = idx_cluster;
  p.score = score; p.index_down = index_down; p.num_clusters = num_clusters;
  p.B = B; p.N = N; p.C = C; p.k = k; p.min_cluster_num = min_cluster_num; p.threshold = threshold;
  // Operand split of the exact Gram (x = hi + lo): IEEE halves carry 11 + 11 significand bits, so hi hi^T + hi lo^T + lo hi^T
  // leaves out only lo lo^T <= 2^-22 |x_i x_j| (below the float32 rounding of the product itself) -- 3 MMAs per k-step and a
  // tighter representation than the bf16 pair (8 + 8 bits, 2^-17 |x|, which needs all 4 terms).  Domain: |x| <= 65504
  // (finite saturation beyond; CLIP features + sincos table are O(1) .. O(1e2)).
  p.f16 = (g_dpc_fused & 3) == 1 ? 1 : 0;
  p.dbg = (g_dpc_fused >> 5) & 3;
  p.terms = (g_dpc_fused & 3) == 2 ? 4 : 3;
  if (g_dpc_fused & 8) p.x_pos = nullptr;    // timing experiments only
  if (g_dpc_fused & 16) p.pos = nullptr;
  p.sqrtC = static_cast<float>(std::sqrt(static_cast<double>(C)));
  int e = 0;
  const float m = std::frexp(p.sqrtC, &e);
  p.inv_sqrtC = (m == 0.5f) ? 1.0f / p.sqrtC : 0.f;
  const bool fb = feat_dtype == SETOK_BF16;

  // features as a 2-D tensor [B*N rows, C]; box = 128 bytes of channels x 256 rows, 128B swizzle, zero fill out of bounds
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SETOK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  CUtensorMap tm;
  const size_t esz = fb ? 2 : 4;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B) * N};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(C) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), 256};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tm, fb ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(feats), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "dpc_fused: cuTensorMapEncodeTiled failed with CUresult %d (B=%d N=%d C=%d)", (int)cr, B, N, C);

  using KernelFn = void (*)(const CUtensorMap, FusedDev);
#define SETOK_FZ_PICK(KS) (fb ? dpc_fused_kernel<KS, true> : dpc_fused_kernel<KS, false>)
  KernelFn fn = k <= 16 ? SETOK_FZ_PICK(16) : (k <= 32 ? SETOK_FZ_PICK(32) : SETOK_FZ_PICK(64));
#undef SETOK_FZ_PICK
  SETOK_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(fn), FZ_SMEM_BYTES));
  const int grid = B < num_sms() ? B : num_sms();
  fn<<<grid, FZ_THREADS, FZ_SMEM_BYTES, stream>>>(tm, p);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" void setok_debug_set_dpc_fused(int mode) { setok::g_dpc_fused = mode; }
#ifdef SETOK_FZ_TRACE
extern "C" void setok_debug_set_fz_trace(void* buf) { cudaMemcpyToSymbol(setok::g_fz_trace, &buf, sizeof buf); }
#endif
