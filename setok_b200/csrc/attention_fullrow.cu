// attention_fullrow.cu — ViT multi-head self-attention for sequences of T <= 257 rows (head_dim 64) with the whole
// score row resident in tensor memory: the 224^2 tower (T = 16*16 + 1 = 257) and every shorter sequence.
//
// Why a second kernel next to attention_tcgen05.cu (any T, 64-key chunks, online softmax): at T = 257 the chunked form
// tiles the problem as 3 x 128 query rows by 5 x 64 keys and executes 1.58x the algorithmic score / exp / P.V work, a third
// of its CTAs carry one live row, and every 64-key chunk pays a S -> softmax -> P.V hand-shake chain (in-kernel timeline:
// ~2800 cycles per chunk, of which ~1300 are barrier hops and MMA issue).  Here:
//   * one persistent CTA per SM walks (image, head) pairs; K, V and Q of a pair are loaded ONCE (TMA, 2-stage ring, the
//     next pair's loads fly during the current pair's math);
//   * two 128-row query tiles (A, B) per pair, TWO softmax threads per row: warp (t, q, kh) owns rows q*32..q*32+31 of tile t and
//     the keys [128 kh, 128 kh + 128); the two halves of a row sit on the same SM sub-partition, so one warp's exponentials
//     (MUFU: 16 results/clk/SM, measured: tools/micro/mufu_rate.cu) run under the other's tensor-memory loads / stores -- a lone
//     warp per sub-partition serialises on every tcgen05.ld / wait / st (155 -> 135 us per layer at B = 256);
//   * S = Q K^T for 256 keys is ONE accumulator of 128 x 256 fp32 = 256 tensor-memory columns per tile (4 MMAs of N = 256):
//     exact two-pass softmax (row maximum, then exp2 / sum) straight from tensor memory in 32-key steps -- no online rescaling,
//     no O correction pass; the halves of a row exchange their maximum and their sum through shared memory + a 64-thread named
//     barrier (both add (half 0) + (half 1): identical bits);
//   * P (bf16 pairs) is written back over score columns the SAME warp has already consumed -- keys [0, 128) -> columns [0, 64),
//     keys [128, 256) -> columns [128, 192) -- and feeds the P.V MMA as its tensor-memory A operand; O lands in columns [192, 256)
//     of the same region (S is dead by then): 2 tiles x 256 columns = all of TMEM;
//   * T = 257 = 256 + 1: the 257th KEY is a 16-column MMA (N = 16: Q . [k_256; 0]^T) into spare columns [80, 96) (scores of keys
//     64..127 of the previous pair, consumed) plus one extra k-step of the P.V product (its P pair at column 64); the 257th QUERY
//     ROW is scored as S^T = K . [q_256; 0]^T (keys along the TMEM lanes, columns [96, 128)), soft-maxed by one auxiliary warp,
//     and its P.V (257 x 64 FMAs) is split over the eight kh = 1 warps on the CUDA cores.  No third query tile, no fifth key chunk.
// Warp roles (608 threads, 92 registers): 0-3 / 4-7 tile A / B keys [0, 128), 8-11 / 12-15 tile A / B keys [128, 256), 16 auxiliary
// (257th row), 17 TMA producer, 18 MMA issuer (highest warp id: the issue arbiter favours it).
#include "common.cuh"

namespace setok {
// 0: always the chunked kernel; 1: whole-row kernel where it wins (225 <= T <= 257: two full query tiles); 2: whole-row kernel
// for every T <= 257 (tests, A/B timing) -- setok_debug_set_attention_fullrow
int g_attn_fullrow = 1;
int g_attn_fullrow_dbg = 0;   // timing experiments only (results are wrong): 1 skip pass 1, 2 no exp2, 4 no P store, 8 no O store
namespace {

constexpr int FR_THREADS = 608;
constexpr int FR_W_AUX = 16, FR_W_TMA = 17, FR_W_MMA = 18;
constexpr int FR_TILE_BYTES = 256 * 128;                 // 256 rows x 64 bf16
constexpr int FR_X_BYTES = 16 * 128;                     // 16-row boxes holding the 257th q / k / v row (+ zero fill)
constexpr int FR_OFF_Q = 0, FR_OFF_K = FR_TILE_BYTES, FR_OFF_V = 2 * FR_TILE_BYTES;
constexpr int FR_OFF_QX = 3 * FR_TILE_BYTES, FR_OFF_KX = FR_OFF_QX + FR_X_BYTES, FR_OFF_VX = FR_OFF_KX + FR_X_BYTES;
constexpr int FR_STAGE_BYTES = 3 * FR_TILE_BYTES + 3 * FR_X_BYTES;     // 102 KiB
constexpr int FR_STAGES = 2;
constexpr int FR_OFF_SLEFT = FR_STAGES * FR_STAGE_BYTES;              // float s_left[256]: scores of row 256 against keys 0..255
constexpr int FR_PLEFT_STRIDE = 264;                                   // floats per buffer: p[0..256], inv_l at [257]
constexpr int FR_OFF_PLEFT = FR_OFF_SLEFT + 256 * 4;                  // float p_left[2][264]
constexpr int FR_OFF_PART = FR_OFF_PLEFT + 2 * FR_PLEFT_STRIDE * 4;   // float part[8][64]
constexpr int FR_OFF_STG = FR_OFF_PART + 8 * 64 * 4;                  // O staging: 16 warps x 16 rows x 64 B (also the partner exchange slots)
constexpr int FR_OFF_BAR = FR_OFF_STG + 16 * 1024;
constexpr int FR_NUM_BARS = 4 * FR_STAGES + 10 + 3;
constexpr int FR_SMEM_BYTES = FR_OFF_BAR + FR_NUM_BARS * 8 + 16 + 1024;
// tensor-memory columns inside a tile's 256-column region
// P of keys [0, 128) over columns [0, 64) and of keys [128, 256) over [128, 192): each half is written by the warp that consumed
// exactly those score columns; the 257th-key / 257th-row extras sit in [64, 128) (scores of keys 64..127, consumed by then)
constexpr uint32_t FR_COL_P1 = 128, FR_COL_O = 192, FR_COL_PX = 64, FR_COL_E = 80, FR_COL_L0 = 96, FR_COL_L1 = 112;

__device__ __forceinline__ float fr_exp2(float x) {      // MUFU.EX2; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifdef SETOK_ATTN_TRACE
__device__ long long* g_fr_trace = nullptr;   // [3 roles][64 slots] clock64 stamps of CTA 0, local pair 2 (tools/attn_timeline.py --fullrow)
#define FR_TRACE(role, slot) do { if (blockIdx.x == 0 && n == 2 && lane == 0 && (warp & 3) == 0 && g_fr_trace != nullptr) g_fr_trace[(role) * 64 + (slot)] = clock64(); } while (0)
#define FR_TRACE_M(slot) do { if (blockIdx.x == 0 && n == 2 && g_fr_trace != nullptr) g_fr_trace[2 * 64 + (slot)] = clock64(); } while (0)
#else
#define FR_TRACE(role, slot) do { } while (0)
#define FR_TRACE_M(slot) do { } while (0)
#endif

struct FrParams {
  bf16* out;
  int T, heads, C, n_pairs;
  float scale_log2;
  int dbg;
};

__global__ void __launch_bounds__(FR_THREADS, 1)
attn_fullrow_hd64_kernel(const __grid_constant__ CUtensorMap tm256, const __grid_constant__ CUtensorMap tm16, FrParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  const int T = p.T, C = p.C;
  const bool leftover = T == 257;                          // one key and one query row beyond the two 128-row tiles
  const int nkeys = T < 256 ? T : 256;                     // keys scored by the S MMA
  const int Nk = (nkeys + 15) & ~15;                       // MMA N
  const int ntiles = T > 128 ? 2 : 1;
  const int n_local = p.n_pairs > static_cast<int>(blockIdx.x) ? (p.n_pairs - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

  float* s_left = reinterpret_cast<float*>(smem + FR_OFF_SLEFT);
  float* p_left = reinterpret_cast<float*>(smem + FR_OFF_PLEFT);
  float* part = reinterpret_cast<float*>(smem + FR_OFF_PART);
  const uint32_t bar0 = base + FR_OFF_BAR;
  // Two load groups per stage with their own full / empty barriers: K + Q (+ the 257th q / k rows) are dead as soon as the
  // pair's score MMAs have been issued, i.e. early in the pair, so the NEXT-BUT-ONE pair's K / Q are in flight almost two
  // pairs ahead of their use; V (+ the 257th v row) lives until the pair's P.V is done and its O has been staged out.
  auto kq_full = [&](int s) { return bar0 + 8u * s; };
  auto kq_empty = [&](int s) { return bar0 + 8u * (FR_STAGES + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (2 * FR_STAGES + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (3 * FR_STAGES + s); };
  auto s_full = [&](int t) { return bar0 + 8u * (4 * FR_STAGES + t); };
  auto p_full = [&](int t) { return bar0 + 8u * (4 * FR_STAGES + 2 + t); };
  auto o_full = [&](int t) { return bar0 + 8u * (4 * FR_STAGES + 4 + t); };
  auto o_read = [&](int t) { return bar0 + 8u * (4 * FR_STAGES + 6 + t); };
  auto e_full = [&](int t) { return bar0 + 8u * (4 * FR_STAGES + 8 + t); };     // the 257th-key / row scores of the NEXT pair are in TMEM
  const uint32_t left_s = bar0 + 8u * (4 * FR_STAGES + 10), left_p = left_s + 8u, left_o = left_s + 16u;
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + FR_OFF_BAR + FR_NUM_BARS * 8);

  if (warp == FR_W_TMA && lane == 0) {
    tma_prefetch_desc(&tm256);
    tma_prefetch_desc(&tm16);
    for (int s = 0; s < FR_STAGES; ++s) {
      mbar_init(kq_full(s), 1); mbar_init(kq_empty(s), leftover ? 2 : 1);            // MMA commit (+ the auxiliary warp)
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), leftover ? 2 : 1);
    }
    for (int t = 0; t < 2; ++t) mbar_init(e_full(t), 1);
    for (int t = 0; t < 2; ++t) { mbar_init(s_full(t), 1); mbar_init(p_full(t), 8); mbar_init(o_full(t), 1); mbar_init(o_read(t), 8); }
    mbar_init(left_s, 4); mbar_init(left_p, 1); mbar_init(left_o, 8);
    fence_mbar_init();
  }
  if (warp == FR_W_MMA) tmem_alloc<512>(base + FR_OFF_BAR + FR_NUM_BARS * 8);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();                                   // qkv comes from the previous kernel of the stream

  if (warp == FR_W_TMA) {
    // ------------------------------------------------ TMA producer ----------------------------------------------------
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int g = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
        const int b = g / p.heads, h = g % p.heads;
        const int st = n & 1;
        const uint32_t sb = base + st * FR_STAGE_BYTES;
        const uint32_t par = ((n >> 1) & 1) ^ 1u;
        mbar_wait(kq_empty(st), par);
        mbar_arrive_expect_tx(kq_full(st), 2 * FR_TILE_BYTES + (leftover ? 2 * FR_X_BYTES : 0));
        tma_load_3d(&tm256, kq_full(st), sb + FR_OFF_K, C + h * 64, 0, b);
        tma_load_3d(&tm256, kq_full(st), sb + FR_OFF_Q, h * 64, 0, b);
        if (leftover) {
          tma_load_3d(&tm16, kq_full(st), sb + FR_OFF_QX, h * 64, 256, b);
          tma_load_3d(&tm16, kq_full(st), sb + FR_OFF_KX, C + h * 64, 256, b);
        }
        mbar_wait(v_empty(st), par);
        mbar_arrive_expect_tx(v_full(st), FR_TILE_BYTES + (leftover ? FR_X_BYTES : 0));
        tma_load_3d(&tm256, v_full(st), sb + FR_OFF_V, 2 * C + h * 64, 0, b);
        if (leftover) tma_load_3d(&tm16, v_full(st), sb + FR_OFF_VX, 2 * C + h * 64, 256, b);
      }
    }
  } else if (warp == FR_W_MMA) {
    // ------------------------------------------------ MMA issuer -------------------------------------------------------
    // The whole warp runs this role convergently (waits included) and one elected lane issues the tcgen05 instructions: with
    // warp-uniform control flow the operand descriptors stay in uniform registers (in the in-kernel timeline a divergent
    // single-lane issuer spent ~90 cycles per tcgen05.mma, three times the 32 cycles a 128x64x16 MMA executes in).
    if (n_local > 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, Nk);
      const uint32_t idesc_x = umma_idesc_bf16(128, 16);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, true);
      const int ksteps = Nk >> 4;
      auto stage_base = [&](int n) { return base + static_cast<uint32_t>((n & 1) * FR_STAGE_BYTES); };
      // S_t(n) = Q_t K^T: 4 k-steps of 16 dims (+32 B inside the 128 B rows)
      auto issue_s = [&](int t, int n, bool last) {
        const uint32_t sb = stage_base(n);
        const uint64_t dq = umma_desc_k_sw128(sb + FR_OFF_Q + t * 16384), dk = umma_desc_k_sw128(sb + FR_OFF_K);
        const uint32_t d = tmem_base + 256u * t;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(s_full(t));
          if (last) umma_commit(kq_empty(n & 1));              // every MMA that reads pair n's K / Q has been issued
        }
        __syncwarp();
      };
      // scores that involve the 257th key / query row of pair n, written into the tail columns of tile t's region
      // (free from the end of the tile's softmax on; read by the softmax threads before the next S overwrites them)
      auto issue_extras = [&](int t, int n) {
        const uint32_t sb = stage_base(n);
        const uint64_t dq = umma_desc_k_sw128(sb + FR_OFF_Q + t * 16384), dkx = umma_desc_k_sw128(sb + FR_OFF_KX);
        const uint64_t dk0 = umma_desc_k_sw128(sb + FR_OFF_K), dk1 = umma_desc_k_sw128(sb + FR_OFF_K + 16384);
        const uint64_t dqx = umma_desc_k_sw128(sb + FR_OFF_QX);
        const uint32_t d = tmem_base + 256u * t;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d + FR_COL_E, dq + 2 * k, dkx + 2 * k, idesc_x, k != 0 ? 1u : 0u);      // rows of tile t . k_256
          if (t == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d + FR_COL_L0, dk0 + 2 * k, dqx + 2 * k, idesc_x, k != 0 ? 1u : 0u);   // keys 0..127 . q_256
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d + FR_COL_L1, dk1 + 2 * k, dqx + 2 * k, idesc_x, k != 0 ? 1u : 0u);   // keys 128..255 . q_256
          }
          umma_commit(e_full(t));
        }
        __syncwarp();
      };
      // O_t(n) = P_t V: P from tensor memory (8 columns = 16 keys per k-step), V as MN-major operand (+2048 B per k-step)
      // O_t(n) = P_t V: P from tensor memory (8 columns = 16 keys per k-step), V as MN-major operand (+2048 B per k-step);
      // commits o_full(t) and, with `last`, hands the pair's V back to the producer.  Commits are issued by the same
      // elected lane as the MMAs they track.
      auto issue_pv = [&](int t, int n, bool last) {
        const uint32_t sb = stage_base(n);
        const uint64_t dv = umma_desc_mn_sw128(sb + FR_OFF_V), dvx = umma_desc_mn_sw128(sb + FR_OFF_VX);
        const uint32_t d = tmem_base + 256u * t;
        if (elect_one_sync()) {
          // P columns: 8 per k-step, keys [0, 128) from column 0, keys [128, 256) from column FR_COL_P1
          if (ksteps == 16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) umma_f16_ts(d + FR_COL_O, d + 8u * k + (k >= 8 ? FR_COL_P1 - 64u : 0u), dv + 128ull * k, idesc_pv, k != 0 ? 1u : 0u);
          } else {
            for (int k = 0; k < ksteps; ++k) umma_f16_ts(d + FR_COL_O, d + 8u * k + (k >= 8 ? FR_COL_P1 - 64u : 0u), dv + 128ull * k, idesc_pv, k != 0 ? 1u : 0u);
          }
          if (leftover) umma_f16_ts(d + FR_COL_O, d + FR_COL_PX, dvx, idesc_pv, 1u);
          umma_commit(o_full(t));
          if (last) umma_commit(v_empty(n & 1));               // every MMA that reads pair n's V has been issued
        }
        __syncwarp();
      };
      // prologue: pair 0's extras, o_full's completion #0 (a bare commit: nothing to drain yet), then S(0)
      mbar_wait(kq_full(0), 0);
      tcgen05_fence_after();
      for (int t = 0; t < ntiles; ++t) {
        if (leftover) issue_extras(t, 0);
        if (elect_one_sync()) umma_commit(o_full(t));
        __syncwarp();
      }
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(o_read(t), 0);
        tcgen05_fence_after();
        issue_s(t, 0, t == ntiles - 1);
      }
      for (int n = 0; n < n_local; ++n) {
        const bool more = n + 1 < n_local;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(p_full(t), n & 1);
          if (lane == 0) FR_TRACE_M(8 * t + 0);
          if (t == 0) mbar_wait(v_full(n & 1), (n >> 1) & 1);
          tcgen05_fence_after();
          issue_pv(t, n, t == ntiles - 1);
          if (lane == 0) FR_TRACE_M(8 * t + 1);
          if (more) {
            if (t == 0) { mbar_wait(kq_full((n + 1) & 1), ((n + 1) >> 1) & 1); tcgen05_fence_after(); }
            if (leftover) issue_extras(t, n + 1);              // behind the P.V in the tensor pipe; its own barrier (e_full)
            mbar_wait(o_read(t), (n + 1) & 1);                 // tile t's threads have drained O(n) and the extras of n + 1
            if (lane == 0) FR_TRACE_M(8 * t + 2);
            tcgen05_fence_after();
            issue_s(t, n + 1, t == ntiles - 1);
            if (lane == 0) FR_TRACE_M(8 * t + 3);
          }
        }
      }
    }
  } else if (warp == FR_W_AUX) {
    // ------------------------------------------------ 257th query row ------------------------------------------------
    if (leftover) {
      for (int n = 0; n < n_local; ++n) {
        const int g = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
        const int b = g / p.heads, h = g % p.heads;
        const uint8_t* sb = smem + (n & 1) * FR_STAGE_BYTES;
        float* pl = p_left + (n & 1) * FR_PLEFT_STRIDE;
        mbar_wait(kq_full(n & 1), (n >> 1) & 1);
        // q_256 . k_256 (row 0 of a 128B-swizzled tile is stored unswizzled): two dims per lane
        const float2 qv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sb + FR_OFF_QX + 4 * lane));
        const float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sb + FR_OFF_KX + 4 * lane));
        const float s_self = warp_sum(fmaf(qv.x, kv.x, qv.y * kv.y));
        if (lane == 0) mbar_arrive(kq_empty(n & 1));        // this warp is done with the pair's K / Q group
        mbar_wait(left_s, n & 1);                            // tile A's threads have stored the 256 scores of this row
        float sv[8];
        float mx = s_self;
#pragma unroll
        for (int i = 0; i < 8; ++i) { sv[i] = s_left[lane + 32 * i]; mx = fmaxf(mx, sv[i]); }
        mx = warp_max(mx) * p.scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e = fr_exp2(fmaf(sv[i], p.scale_log2, -mx));
          pl[lane + 32 * i] = e;
          sum += e;
        }
        const float e_self = fr_exp2(fmaf(s_self, p.scale_log2, -mx));
        sum = warp_sum(sum) + e_self;
        if (lane == 0) { pl[256] = e_self; pl[257] = 1.0f / sum; }
        __syncwarp();
        if (lane == 0) mbar_arrive(left_p);
        mbar_wait(left_o, n & 1);                            // the eight softmax warps have summed their 32-key slices of P.V
        const float* pt = part;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const float2 v = *reinterpret_cast<const float2*>(pt + w * 64 + 2 * lane); o0 += v.x; o1 += v.y; }
        mbar_wait(v_full(n & 1), (n >> 1) & 1);
        const float2 vx = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sb + FR_OFF_VX + 4 * lane));
        const float inv = 1.0f / sum;
        o0 = fmaf(e_self, vx.x, o0) * inv;
        o1 = fmaf(e_self, vx.y, o1) * inv;
        *reinterpret_cast<uint32_t*>(p.out + (static_cast<long long>(b) * T + 256) * C + h * 64 + 2 * lane) = pack_bf16x2(o0, o1);
        __syncwarp();
        if (lane == 0) mbar_arrive(v_empty(n & 1));          // all CUDA-core readers of this pair's V are done
      }
    }
  } else if (warp < 16 && ((warp >> 2) & 1) < ntiles) {
    // ------------------------------------------------ softmax warps ------------------------------------------------
    // Two threads per query row: warp (t, q, kh) owns rows q*32 .. q*32+31 of tile t and the keys [128 kh, 128 kh + 128).  A lone
    // warp per SM sub-partition cannot overlap its own MUFU, tensor-memory and ALU work (tcgen05.ld / wait / st serialise it);
    // with the two halves of a row on the same sub-partition one warp's exponentials run under the other's loads and stores.
    // The halves meet twice per pair through their staging slots and a 64-thread named barrier: row maximum, row sum.
    const int t = (warp >> 2) & 1;                           // query tile
    const int kh = warp >> 3;                                // key half
    const int q = warp & 3;                                  // TMEM lane quarter
    const int row = q * 32 + lane;                           // row within the tile == TMEM lane
    const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 256u * t;
    const float sl2 = p.scale_log2;
    const bool masked = nkeys < 256;
    const int k_lo = 128 * kh;
    const int k_hi = Nk < k_lo + 128 ? Nk : k_lo + 128;
    const int nsteps = k_hi > k_lo ? (k_hi - k_lo + 31) >> 5 : 0;      // 32-key steps of this half
    float* my_x = reinterpret_cast<float*>(smem + FR_OFF_STG + warp * 1024);          // exchange slots alias the O staging
    const float* peer_x = reinterpret_cast<const float*>(smem + FR_OFF_STG + (warp ^ 8) * 1024);
    const int pair_bar = 1 + (warp & 7);                     // named barrier of the two warps that share these rows
    float inv_l = 0.f;
    for (int n = 0; n_local > 0 && n <= n_local; ++n) {
      const int g = static_cast<int>(blockIdx.x) + (n - 1) * static_cast<int>(gridDim.x);     // pair whose O is drained now
      mbar_wait(o_full(t), n & 1);
      FR_TRACE(t, 0);
      tcgen05_fence_after();
      uint4 ov[4];
      if (n > 0) {
        // this half's 32 dims of O(n-1) / l -> bf16, held in registers so that the region can be handed back before anything is stored
        uint32_t o[32];
        tmem_ld_32x32b_x32(treg + FR_COL_O + 32 * kh, o);
        tmem_ld_wait();
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          ov[v4].x = pack_bf16x2(__uint_as_float(o[8 * v4 + 0]) * inv_l, __uint_as_float(o[8 * v4 + 1]) * inv_l);
          ov[v4].y = pack_bf16x2(__uint_as_float(o[8 * v4 + 2]) * inv_l, __uint_as_float(o[8 * v4 + 3]) * inv_l);
          ov[v4].z = pack_bf16x2(__uint_as_float(o[8 * v4 + 4]) * inv_l, __uint_as_float(o[8 * v4 + 5]) * inv_l);
          ov[v4].w = pack_bf16x2(__uint_as_float(o[8 * v4 + 6]) * inv_l, __uint_as_float(o[8 * v4 + 7]) * inv_l);
        }
      }
      float s_x = -INFINITY;
      if (leftover && kh == 0 && n < n_local) {
        // scores against the 257th key (this row) and of the 257th row (this lane's key), parked in the region's spare columns by
        // MMAs that run behind the P.V in the tensor pipe (hence after the O drain: their barrier fires later than o_full)
        mbar_wait(e_full(t), n & 1);
        tcgen05_fence_after();
        s_x = __uint_as_float(tmem_ld_32x32b_x1(treg + FR_COL_E));
        if (t == 0) {
          const float l0 = __uint_as_float(tmem_ld_32x32b_x1(treg + FR_COL_L0));
          const float l1 = __uint_as_float(tmem_ld_32x32b_x1(treg + FR_COL_L1));
          tmem_ld_wait();
          s_left[row] = l0;
          s_left[128 + row] = l1;
        } else {
          tmem_ld_wait();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(o_read(t));                              // the region may take the next S (once all eight warps of the tile are here)
        if (leftover && t == 0 && kh == 0 && n < n_local) mbar_arrive(left_s);
      }
      FR_TRACE(t, 1);
      if (n > 0) {
        // rows -> global through this warp's 1 KiB of staging, 16 rows x 64 B per round: a thread owns a row in tensor memory, but
        // a coalesced store wants 4 lanes on one 64-byte half row; both shared-memory sides are conflict-free (XOR swizzle)
        const int b = g / p.heads, h = g % p.heads;
        uint8_t* stg = smem + FR_OFF_STG + warp * 1024;
#pragma unroll
        for (int rnd = 0; rnd < 2; ++rnd) {
          if ((lane >> 4) == rnd) {
            const int lr = lane & 15;
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(stg + lr * 64 + ((c ^ ((lr >> 1) & 3)) << 4)) = ov[c];
          }
          __syncwarp();
          if (!(p.dbg & 8)) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int lr = it * 8 + (lane >> 2), c = lane & 3;
              const uint4 v = *reinterpret_cast<const uint4*>(stg + lr * 64 + ((c ^ ((lr >> 1) & 3)) << 4));
              const int grow = t * 128 + q * 32 + rnd * 16 + lr;
              if (grow < T) *reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * T + grow) * C + h * 64 + kh * 32 + c * 8) = v;
            }
          }
          __syncwarp();
        }
      }
      if (n == n_local) break;

      FR_TRACE(t, 2);
      mbar_wait(s_full(t), n & 1);
      FR_TRACE(t, 3);
      tcgen05_fence_after();
      uint32_t r[32];
      // pass 1: maximum over this half's keys, then over the row through the partner's slot
      float mx0 = s_x, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (!(p.dbg & 1)) {
#pragma unroll 1
        for (int hs = 0; hs < nsteps; ++hs) {
          const int k0 = k_lo + 32 * hs;
          tmem_ld_32x32b_x32(treg + k0, r);
          tmem_ld_wait();
          if (masked) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (k0 + i >= nkeys) r[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
          }
        }
      }
      my_x[lane] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      named_bar_sync(pair_bar, 64);
      const float m = (p.dbg & 1) ? 8.0f : fmaxf(my_x[lane], peer_x[lane]) * sl2;
      FR_TRACE(t, 4);
      const bool no_exp = (p.dbg & 2) != 0;
      // pass 2: p = exp2(s * scale - m); P (bf16 pairs) goes back over score columns this warp has already consumed (the 16
      // columns of step hs lie inside the columns of step hs / 2 of the same half)
      float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
#pragma unroll 1
      for (int hs = 0; hs < nsteps; ++hs) {
        const int k0 = k_lo + 32 * hs;
        tmem_ld_32x32b_x32(treg + k0, r);
        tmem_ld_wait();
        if (masked) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (k0 + i >= nkeys) r[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float p0 = fmaf(__uint_as_float(r[i]), sl2, -m), p1 = fmaf(__uint_as_float(r[i + 1]), sl2, -m);
          float p2 = fmaf(__uint_as_float(r[i + 2]), sl2, -m), p3 = fmaf(__uint_as_float(r[i + 3]), sl2, -m);
          if (!no_exp) { p0 = fr_exp2(p0); p1 = fr_exp2(p1); p2 = fr_exp2(p2); p3 = fr_exp2(p3); }
          ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
          r[i / 2] = pack_bf16x2(p0, p1);               // in place: slots <= i/2+1 were consumed already
          r[i / 2 + 1] = pack_bf16x2(p2, p3);
        }
        if (!(p.dbg & 4)) tmem_st_32x32b_x16(treg + (kh ? FR_COL_P1 : 0u) + 16 * hs, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
      }
      float l = (ps0 + ps1) + (ps2 + ps3);
      if (leftover && kh == 0) {
        const float p_x = fr_exp2(fmaf(s_x, sl2, -m));
        l += p_x;
        uint32_t px[8] = {pack_bf16x2(p_x, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st_32x32b_x8(treg + FR_COL_PX, px);
      }
      my_x[32 + lane] = l;
      named_bar_sync(pair_bar, 64);
      {
        const float lp = peer_x[32 + lane];
        inv_l = 1.0f / (kh == 0 ? l + lp : lp + l);       // both halves add (half 0) + (half 1) in that order: identical bits
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(t));
      FR_TRACE(t, 5);

      if (leftover && kh == 1) {
        // this warp's 32-key slice of the 257th row's P.V on the CUDA cores: lane <-> dims (2 lane, 2 lane + 1)
        const int sl = warp & 7;
        mbar_wait(v_full(n & 1), (n >> 1) & 1);
        mbar_wait(left_p, n & 1);
        const float* pl = p_left + (n & 1) * FR_PLEFT_STRIDE + 32 * sl;
        const uint8_t* vt = smem + (n & 1) * FR_STAGE_BYTES + FR_OFF_V;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) {
          const int rr = 32 * sl + jj;
          const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vt + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + (lane & 3) * 4));
          const float pj = pl[jj];
          a0 = fmaf(pj, v.x, a0);
          a1 = fmaf(pj, v.y, a1);
        }
        *reinterpret_cast<float2*>(part + sl * 64 + 2 * lane) = make_float2(a0, a1);
        __syncwarp();
        if (lane == 0) mbar_arrive(left_o);
        FR_TRACE(t, 6);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == FR_W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn fr_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

}  // namespace

bool attention_fullrow_supported(int T) {
  if (g_attn_fullrow == 0 || T < 1 || T > 257) return false;
  return g_attn_fullrow == 2 || T >= 225;
}

// qkv bf16 [B*T, 3C] -> out bf16 [B*T, C]; heads of 64; softmax(q k^T * scale) v per image; T <= 257.
int launch_attention_fullrow(const void* qkv, void* out, int B, int T, int C, int heads, float scale, cudaStream_t stream) {
  EncodeTiledFn enc = fr_encode_fn();
  if (!enc) return fail(SETOK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tm256, tm16;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(3 * C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(3 * C) * 2, static_cast<cuuint64_t>(T) * 3 * C * 2};
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t box256[3] = {64, 256, 1}, box16[3] = {64, 16, 1};
  CUresult r = enc(&tm256, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box256, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "attention(fullrow): cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  r = enc(&tm16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box16, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SETOK_ERR_CUDA, "attention(fullrow): cuTensorMapEncodeTiled (16-row box) failed with CUresult %d", (int)r);
  SETOK_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fullrow_hd64_kernel), FR_SMEM_BYTES));
  FrParams p;
  p.out = static_cast<bf16*>(out); p.T = T; p.heads = heads; p.C = C; p.n_pairs = B * heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.dbg = g_attn_fullrow_dbg;
  const int grid = p.n_pairs < num_sms() ? p.n_pairs : num_sms();
  SETOK_CUDA_OK(launch_pdl(attn_fullrow_hd64_kernel, dim3(grid), dim3(FR_THREADS), FR_SMEM_BYTES, stream, tm256, tm16, p));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" void setok_debug_set_attention_fullrow(int on) { setok::g_attn_fullrow = on; }
extern "C" void setok_debug_set_attention_fullrow_dbg(int flags) { setok::g_attn_fullrow_dbg = flags; }
#ifdef SETOK_ATTN_TRACE
extern "C" void setok_debug_set_fullrow_trace(void* buf) { cudaMemcpyToSymbol(setok::g_fr_trace, &buf, sizeof buf); }
#endif
