// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   D[M,N] = epilogue(A[M,K] * W[N,K]^T),  fp32 accumulation in TMEM.
//
// One CTA per SM (CTAs paired into 2-CTA clusters for the cta_group::2 variant, see Cfg), 320 threads:
//   warp 8      TMA producer   (cp.async.bulk.tensor 2-D, 128B swizzle, 4-stage mbarrier ring)
//   warp 9      MMA issuer     (one thread: tcgen05.mma.cta_group::1.kind::f16, 128x256x16 per instr),
//               owns the TMEM allocation (512 columns = two 128x256 fp32 accumulators); highest warp id
//               because the issue arbiter favours high warp ids and this warp feeds the tensor pipe
//   warps 0..7  epilogue       (tcgen05.ld -> per-warp smem transpose -> coalesced bias/act/residual/store);
//               two warps per TMEM lane quarter (one per 128-column half) so each SM sub-partition has two
//               epilogue warps to hide the MUFU / residual-load latency behind
// The two accumulators let the epilogue of tile i overlap the MMAs of tile i+1.  Tiles are walked
// N-fastest so the CTAs running concurrently share A row-blocks through L2 while W (<= 8 MB) stays
// L2-resident.  M may be ragged (TMA zero-fills / the epilogue masks) and may live on the device
// (m_dev) so data-dependent row counts need no host sync.
#include "common.cuh"

namespace setok {
namespace {

#ifndef SETOK_LNF_PF
#define SETOK_LNF_PF 1      // residual chunks requested ahead in the folded transposing epilogue (2 spills 152 bytes; A/B build knob)
#endif
constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int EPI_WARPS = 8;                      // two warps per TMEM lane quarter, each owning half of the 256 columns
constexpr int STG_BYTES_PER_WARP = 32 * 32 * 4;   // 32 rows x 32 fp32 columns
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KiB: this CTA's 128 rows of A

// CG = 1: one CTA computes a 128x256 tile, B tile 256 rows (32 KiB/stage), 4 stages.
// CG = 2: a CTA pair (cta_group::2) computes a 256x256 tile; each CTA stages its own 128 rows of A and HALF of the B
//         tile (128 rows, 16 KiB/stage) -- the tensor cores read the other half from the peer's shared memory -- so
//         the B operand crosses L2->SM once per pair, and the smaller stage buys 6 stages.
template <int CG> struct Cfg {
  static constexpr int STAGES = CG == 2 ? 6 : 4;
  static constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
  static constexpr int OFF_STG = OFF_B + STAGES * B_STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_STG + EPI_WARPS * STG_BYTES_PER_WARP;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;   // + tmem ptr + alignment slack
};

// float4 activation accesses of the transposing epilogue, with the evict-first hint under SETOK_EPI_STREAMING
__device__ __forceinline__ float4 epi_ld4(const float* p) {
#if SETOK_EPI_STREAMING
  return __ldcs(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}
__device__ __forceinline__ void epi_st4(float* p, float4 v) {
#if SETOK_EPI_STREAMING
  __stcs(reinterpret_cast<float4*>(p), v);
#else
  *reinterpret_cast<float4*>(p) = v;
#endif
}

struct GemmDev {
  void* D; long long ldd;
  const float* bias;
  const void* res; long long ldr;
  const int32_t* m_dev;
  int M, N, K;
  int act, out_f32, res_kind, remap_P;
  int batch;
  long long d_batch_stride;   // elements of D's type
  long long r_batch_stride;   // elements of the residual's type
  int w_mn_major;
  int direct;                 // row-owner epilogue without the shared-memory transpose (see the epilogue branch)
  const float* ln_in; float* ln_out; const float* ln_s;   // LayerNorm fold (LNF != 0): row records in / out, column sums of W'
  bf16* xhat; long long ld_xhat;
  float ln_eps, ln_invC; int ln_ns;
  int epi_mode;               // debug (setok_debug_set_gemm_epi_mode): 0 normal; 1 drain only (no transpose / math / stores);
                              // 2 transpose + math, no residual loads / stores; 3 normal minus the residual loads; 4 / 5: drain only and the
                              // producer stages A only / nothing (is the main loop bound by L2->SM operand traffic?)
};

// Epilogue configuration is a template so the per-element code has no run-time branches; -1 = run time
// (the generic instantiation serves the rarely used combinations).
//
// LNF (row-owner epilogue only): the pre-LN transformer's LayerNorms folded into the GEMMs on either side of them, so that the
// residual stream is never re-read by a normalisation pass.  The GEMM that PRODUCES the stream (out_proj / fc2, LNF = 2) also
// writes xhat = bf16((x - c) * r) -- x normalised with the row's statistics of one sub-layer earlier, (c, r), which are known
// before the row is complete and keep xhat O(1) so that its bf16 rounding is the rounding the LayerNorm output would get --
// and, per 128-column half tile, the partial sums s1 = sum(x - c), s2 = sum((x - c)^2).  The GEMM that CONSUMES the normalised
// rows (qkv / fc1, LNF = 1) multiplies xhat by W' = gamma (.) W and finishes the LayerNorm exactly in its fp32 epilogue:
//   m = s1 / C, var = s2 / C - m^2, rho = rsqrt(var + eps):  LN(x) = (rho / r) * xhat + rho * (c - mu),  mu = c + m
//   y_n = (rho / r) * acc_n - rho * m * sum_k W'_nk + (W beta + b)_n
// The partial sums are added in slot order by every reader: bit-deterministic, no atomics.
template <int CG, int ACT, int RES, int OUTF32, int REMAP, bool DIRECT, int LNF = 0>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmDev p) {
  constexpr int STAGES = Cfg<CG>::STAGES, B_STAGE_BYTES = Cfg<CG>::B_STAGE_BYTES, STAGE_BYTES = Cfg<CG>::STAGE_BYTES;
  constexpr int OFF_A = Cfg<CG>::OFF_A, OFF_B = Cfg<CG>::OFF_B, OFF_STG = Cfg<CG>::OFF_STG, OFF_BAR = Cfg<CG>::OFF_BAR;
  constexpr int NUM_BARS = Cfg<CG>::NUM_BARS;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;      // position in the CTA pair; rank 0 issues the MMAs
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();            // the next kernel of the stream may set itself up while this one runs

  const uint32_t bar0 = base + OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + NUM_BARS * 8);

  constexpr int W_TMA = EPI_WARPS, W_MMA = EPI_WARPS + 1;
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    // full: armed by the leader's producer with the byte count of BOTH CTAs' loads; tempty: every epilogue warp of the pair
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS * CG); }
    fence_mbar_init();
  }
  if (warp == W_MMA) {
    if (CG == 2) tmem_alloc_2cta<TMEM_COLS>(base + OFF_BAR + NUM_BARS * 8);
    else tmem_alloc<TMEM_COLS>(base + OFF_BAR + NUM_BARS * 8);
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();                         // operands, residual and the device row count come from earlier kernels of the stream

  int M_eff = p.M;
  if (p.m_dev != nullptr) { int m = *p.m_dev; M_eff = m < p.M ? (m < 0 ? 0 : m) : p.M; }
  constexpr int TILE_M = BM * CG;
  const int tiles_m = (M_eff + TILE_M - 1) / TILE_M;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_per_batch = tiles_m * tiles_n;
  const int num_tiles = tiles_per_batch * p.batch;
  const int k_blocks = (p.K + BK - 1) / BK;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;   // both CTAs of a pair walk the same tiles

  if (warp == W_TMA) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int bt = t / tiles_per_batch, tt = t % tiles_per_batch;
        const int m_blk = tt / tiles_n, n_blk = tt % tiles_n;
        const int a_row = m_blk * TILE_M + static_cast<int>(rank) * BM;
        const int b_row = n_blk * BN + static_cast<int>(rank) * (BN / CG);      // first N index this CTA stages
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + OFF_A + stage * A_STAGE_BYTES, sb = base + OFF_B + stage * B_STAGE_BYTES;
          if (p.epi_mode >= 4) {
            // diagnostic (tools/bench_gemm_modes.py): mode 4 stages A only, mode 5 nothing -- the MMAs run on stale smem, so
            // the time left is the tensor pipe's (plus A's share of the L2 traffic in mode 4)
            if (CG == 2) {
              const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
              if (p.epi_mode == 4) {
                if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * A_STAGE_BYTES);
                tma_load_3d_2sm(&tmA, lead_full, sa, kb * BK, a_row, bt);
              } else if (rank == 0) {
                mbar_arrive(full_bar(stage));
              }
            } else if (p.epi_mode == 4) {
              mbar_arrive_expect_tx(full_bar(stage), A_STAGE_BYTES);
              tma_load_3d(&tmA, full_bar(stage), sa, kb * BK, a_row, bt);
            } else {
              mbar_arrive(full_bar(stage));
            }
          } else if (CG == 2) {
            // both CTAs' bytes are counted on the LEADER's full barrier (the MMA issuer waits there).  The peer cannot
            // run a phase ahead: its empty barrier for this stage fires only after the MMAs that consumed the stage
            // retired, i.e. after the leader's full barrier already flipped.
            const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            tma_load_3d_2sm(&tmA, lead_full, sa, kb * BK, a_row, bt);
            if (p.w_mn_major) {
              // B tile = (BN/CG)/64 atoms of [64 K-rows x 64 N-columns]; global W is [K, N] (N contiguous)
#pragma unroll
              for (int at = 0; at < BN / CG / 64; ++at) tma_load_3d_2sm(&tmB, lead_full, sb + at * 8192, b_row + at * 64, kb * BK, bt);
            } else {
              tma_load_3d_2sm(&tmB, lead_full, sb, kb * BK, b_row, bt);
            }
          } else {
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_3d(&tmA, full_bar(stage), sa, kb * BK, a_row, bt);
            if (p.w_mn_major) {
#pragma unroll
              for (int at = 0; at < BN / CG / 64; ++at) tma_load_3d(&tmB, full_bar(stage), sb + at * 8192, b_row + at * 64, kb * BK, bt);
            } else {
              tma_load_3d(&tmB, full_bar(stage), sb, kb * BK, b_row, bt);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == W_MMA) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM * CG, BN, p.w_mn_major != 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = base + OFF_A + stage * A_STAGE_BYTES;
          const uint32_t b_addr = base + OFF_B + stage * B_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = umma_desc_k_sw128(a_addr + k * UMMA_K * 2);
            // K-major B: +32 B per 16-wide k-step inside the 128 B row; MN-major B: +16 K-rows = 2048 B, atoms 8 KiB apart
            const uint64_t bdesc = p.w_mn_major ? umma_desc_mn_sw128(b_addr + k * 2048, 8192) : umma_desc_k_sw128(b_addr + k * UMMA_K * 2);
            if (CG == 2) umma_f16_2cta(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // smem slot reusable once these MMAs retire (in both CTAs of a pair)
          if (CG == 2) umma_commit_2cta_mc(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CG == 2) umma_commit_2cta_mc(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    static_assert(!DIRECT || (ACT >= 0 && RES >= 0 && OUTF32 >= 0 && REMAP == 0), "row-owner epilogue: specialised instantiations only");
    static_assert(LNF != 1 || DIRECT, "LayerNorm fold, consuming side: row-owner epilogue only");
    static_assert(LNF != 1 || (RES == 0 && OUTF32 == 0), "LayerNorm fold, consuming side: bf16 output, no residual");
    static_assert(LNF != 2 || (RES == 2 && OUTF32 == 1 && ACT == 0), "LayerNorm fold, producing side: f32 residual stream");
    const int act = ACT >= 0 ? ACT : p.act;
    const int res_kind = RES >= 0 ? RES : p.res_kind;
    const bool out_f32 = OUTF32 >= 0 ? (OUTF32 != 0) : (p.out_f32 != 0);
    const int remap_P = REMAP >= 0 ? (REMAP ? p.remap_P : 0) : p.remap_P;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = warp >> 2;                   // which 128-column half of the tile this warp drains
    uint8_t* stg = smem + OFF_STG + warp * STG_BYTES_PER_WARP;
    int acc = 0; uint32_t acc_phase = 0;
    const int j = lane & 7;                       // 16-byte column chunk handled in the coalesced phase
    const int rsub = lane >> 3;                   // row (mod 4) handled in the coalesced phase
    const float* bias = p.bias;
    for (int t = tile0; t < num_tiles; t += tile_step) {
      const int bt = t / tiles_per_batch, tt = t % tiles_per_batch;
      const int m_blk = tt / tiles_n, n_blk = tt % tiles_n;
      const int row0 = m_blk * TILE_M + static_cast<int>(rank) * BM + q * 32;
      const int n0 = n_blk * BN + half * 128;
      const long long dbase = static_cast<long long>(bt) * p.d_batch_stride;
      bool waited = false;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + half * 128);
      if constexpr (DIRECT) {
        // Row-owner epilogue: thread `lane` keeps TMEM lane (= tile row) q*32+lane and writes its 32 columns of every chunk
        // straight from registers with 256-bit stores -- one full 32-byte sector per thread per instruction -- so the tile
        // never crosses shared memory (whose bandwidth the TMA writes and the MMA operand reads already fill) and there
        // is no intra-warp synchronisation.  The next chunk's tcgen05.ld is in flight while this one is processed.
        const int grow = row0 + lane;
        const bool row_ok = grow < M_eff;
        constexpr int RW = RES == 1 ? 16 : 32;                 // residual words per chunk and thread
        constexpr int DW = OUTF32 ? 32 : 16;                   // output words per chunk and thread
        uint32_t rres[RES != 0 ? 2 : 1][RES != 0 ? RW : 1];
        const char* rptr = nullptr;
        if (RES != 0) rptr = static_cast<const char*>(p.res) + (static_cast<long long>(bt) * p.r_batch_stride + static_cast<long long>(grow) * p.ldr + n0) * (RES == 1 ? 2 : 4);
        char* dptr = static_cast<char*>(p.D) + (dbase + static_cast<long long>(grow) * p.ldd + n0) * (OUTF32 ? 4 : 2);
        auto load_res = [&](int ch, int slot) {
          if (RES != 0) {
            const bool ok = row_ok && (n0 + ch * 32 < p.N);
#pragma unroll
            for (int v = 0; v < RW / 8; ++v) {
              if (ok) ld_global_v8(rptr + ch * (RW * 4) + v * 32, &rres[slot][8 * v]);
              else {
#pragma unroll
                for (int e = 0; e < 8; ++e) rres[slot][8 * v + e] = 0u;
              }
            }
          }
        };
        if (RES != 0 && p.epi_mode == 0) { load_res(0, 0); load_res(1, 1); }
        // LayerNorm fold: this row's record (read under the MMAs of the tile): the consuming side finishes the LayerNorm as
        // y = ln_a * acc + ln_b * s_n + t_n, the producing side emits xhat = (x - ln_c) * ln_r and its partial sums
        float ln_a = 1.f, ln_b = 0.f, ln_c = 0.f, ln_r = 1.f, ln_s1 = 0.f, ln_s2 = 0.f;
        if constexpr (LNF != 0) {
          if (row_ok) {
            const float2* rec = reinterpret_cast<const float2*>(p.ln_in + static_cast<long long>(grow) * (2 + 2 * p.ln_ns));
            const float2 pre = rec[0];
            float S1 = 0.f, S2 = 0.f;
            for (int i = 0; i < p.ln_ns; ++i) { const float2 part = rec[1 + i]; S1 += part.x; S2 += part.y; }
            const float m = S1 * p.ln_invC;
            const float var = fmaxf(fmaf(-m, m, S2 * p.ln_invC), 0.f);
            const float rho = rsqrtf(var + p.ln_eps);
            if (LNF == 1) { ln_a = rho / pre.y; ln_b = -rho * m; }
            else { ln_c = pre.x + m; ln_r = rho; }
          }
        }
        char* xptr = nullptr;
        if (LNF == 2) xptr = reinterpret_cast<char*>(p.xhat) + (static_cast<long long>(grow) * p.ld_xhat + n0) * 2;
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        // accumulator chunks: double-buffered (the next tcgen05.ld in flight under this chunk's math and stores) unless the
        // f32 residual ring already takes 64 registers (3 warps share an SM sub-partition: 168 registers per thread)
        constexpr int NB = RES == 2 ? 1 : 2;
        uint32_t ra[NB][32];
        tmem_ld_32x32b_x32(taddr, ra[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int gcol0 = n0 + ch * 32;
          tmem_ld_wait();
          if (NB == 2 && ch + 1 < 4) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, ra[(ch + 1) % NB]);
          uint32_t* a = ra[ch % NB];
          if (gcol0 >= p.N || p.epi_mode == 1) {
            if (NB == 1 && ch + 1 < 4) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, ra[0]);
            continue;
          }
          uint32_t* outw = a;           // results replace the accumulator words in place (word 2c / 4c is consumed before it is rewritten)
          uint32_t xw[LNF == 2 ? 8 : 1];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias + gcol0) + c);
            float v[4];
            if constexpr (LNF == 1) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.ln_s + gcol0) + c);
              v[0] = fmaf(__uint_as_float(a[4 * c]), ln_a, fmaf(ln_b, s4.x, b4.x));
              v[1] = fmaf(__uint_as_float(a[4 * c + 1]), ln_a, fmaf(ln_b, s4.y, b4.y));
              v[2] = fmaf(__uint_as_float(a[4 * c + 2]), ln_a, fmaf(ln_b, s4.z, b4.z));
              v[3] = fmaf(__uint_as_float(a[4 * c + 3]), ln_a, fmaf(ln_b, s4.w, b4.w));
            } else {
              v[0] = __uint_as_float(a[4 * c]) + b4.x; v[1] = __uint_as_float(a[4 * c + 1]) + b4.y;
              v[2] = __uint_as_float(a[4 * c + 2]) + b4.z; v[3] = __uint_as_float(a[4 * c + 3]) + b4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (ACT == SETOK_ACT_QUICK_GELU) v[e] = act_quick_gelu(v[e]);
              else if (ACT == SETOK_ACT_GELU_ERF) v[e] = act_gelu_erf(v[e]);
            }
            if (RES == 1) {
              const float2 lo = unpack_bf16x2(rres[ch & 1][2 * c]), hi = unpack_bf16x2(rres[ch & 1][2 * c + 1]);
              v[0] += lo.x; v[1] += lo.y; v[2] += hi.x; v[3] += hi.y;
            } else if (RES == 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] += __uint_as_float(rres[ch & 1][4 * c + e]);
            }
            if (OUTF32) {
#pragma unroll
              for (int e = 0; e < 4; ++e) outw[4 * c + e] = __float_as_uint(v[e]);
            } else {
              outw[2 * c] = pack_bf16x2(v[0], v[1]);
              outw[2 * c + 1] = pack_bf16x2(v[2], v[3]);
            }
            if constexpr (LNF == 2) {
              float d[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) { d[e] = v[e] - ln_c; ln_s1 += d[e]; ln_s2 = fmaf(d[e], d[e], ln_s2); }
              xw[2 * (c & 3)] = pack_bf16x2(d[0] * ln_r, d[1] * ln_r);
              xw[2 * (c & 3) + 1] = pack_bf16x2(d[2] * ln_r, d[3] * ln_r);
              if ((c & 3) == 3 && row_ok) st_global_v8(xptr + ch * 64 + (c >> 2) * 32, xw);     // 16 bf16 = one 32-byte sector
            }
          }
          if (RES != 0 && ch + 2 < 4 && p.epi_mode == 0) load_res(ch + 2, ch & 1);
          if (row_ok && p.epi_mode != 2) {
#pragma unroll
            for (int v = 0; v < DW / 8; ++v) st_global_v8(dptr + ch * (DW * 4) + v * 32, &outw[8 * v]);
          }
          if (NB == 1 && ch + 1 < 4) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, ra[0]);
        }
        if constexpr (LNF == 2) {
          if (row_ok && n0 < p.N) {
            float2* rec = reinterpret_cast<float2*>(p.ln_out + static_cast<long long>(grow) * (2 + 2 * p.ln_ns));
            rec[1 + (n0 >> 7)] = make_float2(ln_s1, ln_s2);
            if (n0 == 0) rec[0] = make_float2(ln_c, ln_r);
          }
        }
        waited = true;
      } else {
      // Residual of the ViT's out_proj / fc2 (x += ...): requested ahead of its use so that the HBM latency overlaps the
      // MMAs instead of being paid once per 32-column chunk.  bf16: the whole tile (4 chunks, 64 registers) before waiting
      // for the accumulator; f32: a ring of two chunks (64 registers), chunk c + 2 requested when chunk c has been consumed.
      // (one chunk ahead only under the LayerNorm fold, whose row-owner pass needs the registers)
      constexpr int PF = (REMAP == 0 && RES == 1) ? 4 : ((REMAP == 0 && RES == 2) ? (LNF == 2 ? SETOK_LNF_PF : 2) : 0);
      uint2 rt[PF > 0 && RES == 1 ? PF : 1][8];
      float4 rtf[PF > 0 && RES == 2 ? PF : 1][8];
      auto prefetch = [&](int ch, int slot) {
        const int col = n0 + ch * 32 + 4 * j;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row0 + it * 4 + rsub;
          const bool ok = grow < M_eff && col < p.N;
          if (RES == 1) {
            rt[slot][it] = make_uint2(0u, 0u);
            if (ok) rt[slot][it] = *reinterpret_cast<const uint2*>(static_cast<const bf16*>(p.res) + static_cast<long long>(bt) * p.r_batch_stride + static_cast<long long>(grow) * p.ldr + col);
          } else {
            rtf[slot][it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) rtf[slot][it] = epi_ld4(static_cast<const float*>(p.res) + static_cast<long long>(bt) * p.r_batch_stride + static_cast<long long>(grow) * p.ldr + col);
          }
        }
      };
      const int epi_mode = p.epi_mode;
      if (PF > 0 && epi_mode == 0) {
#pragma unroll
        for (int ch = 0; ch < PF; ++ch) prefetch(ch, ch);
      }
      // LayerNorm fold, producing side (HBM-bound shapes keep this epilogue's 128-byte row segments for the f32 stream): the
      // finished chunk goes back into the warp's staging tile and the thread that owns the row takes its sums and xhat from there
      float ln_c = 0.f, ln_r = 1.f, ln_s1 = 0.f, ln_s2 = 0.f;
      const int own_row = row0 + lane;
      const bool own_ok = own_row < M_eff;
      char* xptr = nullptr;
      if constexpr (LNF == 2) {
        if (own_ok) {
          const float2* rec = reinterpret_cast<const float2*>(p.ln_in + static_cast<long long>(own_row) * (2 + 2 * p.ln_ns));
          const float2 pre = rec[0];
          float S1 = 0.f, S2 = 0.f;
          for (int i = 0; i < p.ln_ns; ++i) { const float2 part = rec[1 + i]; S1 += part.x; S2 += part.y; }
          const float m = S1 * p.ln_invC;
          const float var = fmaxf(fmaf(-m, m, S2 * p.ln_invC), 0.f);
          ln_c = pre.x + m;
          ln_r = rsqrtf(var + p.ln_eps);
        }
        xptr = reinterpret_cast<char*>(p.xhat) + (static_cast<long long>(own_row) * p.ld_xhat + n0) * 2;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int gcol0 = n0 + ch * 32;
        if (gcol0 >= p.N) break;
        const int col = gcol0 + 4 * j;
        const bool col_ok = col < p.N;
        // residual prefetch in the coalesced layout: in flight while the accumulator is drained and transposed
        uint2 rb[8];
        float4 rf[8];
        if (PF > 0) {
          constexpr int PFm = PF > 0 ? PF : 1;      // (the branch is dead for PF == 0; keeps the modulo well-formed)
#pragma unroll
          for (int it = 0; it < 8; ++it) { if (RES == 1) rb[it] = rt[ch % PFm][it]; else rf[it] = rtf[ch % PFm][it]; }
          if (PF < 4 && ch + PF < 4 && epi_mode == 0) prefetch(ch + PF, ch % PFm);
        } else if (res_kind == 1) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int grow = row0 + it * 4 + rsub;
            rb[it] = make_uint2(0u, 0u);
            if (grow < M_eff && col_ok) {
              const long long rrow = remap_P > 0 ? (1 + grow % remap_P) : grow;
              rb[it] = *reinterpret_cast<const uint2*>(static_cast<const bf16*>(p.res) + static_cast<long long>(bt) * p.r_batch_stride + rrow * p.ldr + col);
            }
          }
        } else if (res_kind == 2) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int grow = row0 + it * 4 + rsub;
            rf[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < M_eff && col_ok) {
              const long long rrow = remap_P > 0 ? (1 + grow % remap_P) : grow;
              rf[it] = *reinterpret_cast<const float4*>(static_cast<const float*>(p.res) + static_cast<long long>(bt) * p.r_batch_stride + rrow * p.ldr + col);
            }
          }
        }
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
        if (!waited) {
          mbar_wait(tfull_bar(acc), acc_phase);
          tcgen05_fence_after();
          waited = true;
        }
        uint32_t r0[32];
        tmem_ld_32x32b_x32(taddr + ch * 32, r0);
        tmem_ld_wait();
        if (epi_mode == 1 || epi_mode >= 4) continue;
        // transpose through smem: thread `lane` owns tile row q*32+lane, 32 fp32 columns (8 x 16 B, XOR-swizzled)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
              make_uint4(r0[4 * c], r0[4 * c + 1], r0[4 * c + 2], r0[4 * c + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rsub;
          const int grow = row0 + r;
          float4 v = *reinterpret_cast<const float4*>(stg + r * 128 + ((j ^ (r & 7)) << 4));
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          if (act == SETOK_ACT_QUICK_GELU) {
            v.x = act_quick_gelu(v.x); v.y = act_quick_gelu(v.y); v.z = act_quick_gelu(v.z); v.w = act_quick_gelu(v.w);
          } else if (act == SETOK_ACT_GELU_ERF) {
            v.x = act_gelu_erf(v.x); v.y = act_gelu_erf(v.y); v.z = act_gelu_erf(v.z); v.w = act_gelu_erf(v.w);
          }
          if (res_kind == 1) {
            const float2 lo = unpack_bf16x2(rb[it].x), hi = unpack_bf16x2(rb[it].y);
            v.x += lo.x; v.y += lo.y; v.z += hi.x; v.w += hi.y;
          } else if (res_kind == 2) {
            v.x += rf[it].x; v.y += rf[it].y; v.z += rf[it].z; v.w += rf[it].w;
          }
          if (epi_mode == 2) { if (v.x == 1.2345e30f) p.act = 0; continue; }   // keep the math alive, store nothing
          if constexpr (LNF == 2) *reinterpret_cast<float4*>(stg + r * 128 + ((j ^ (r & 7)) << 4)) = v;
          if (grow < M_eff && col_ok) {
            const long long orow = remap_P > 0 ? (grow + grow / remap_P + 1) : grow;
            if (out_f32) {
              epi_st4(static_cast<float*>(p.D) + dbase + orow * p.ldd + col, v);
            } else {
              *reinterpret_cast<uint2*>(static_cast<bf16*>(p.D) + dbase + orow * p.ldd + col) =
                  make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
            }
          }
        }
        __syncwarp();
        if constexpr (LNF == 2) {
          uint32_t xw[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4));
            const float d[4] = {t.x - ln_c, t.y - ln_c, t.z - ln_c, t.w - ln_c};
#pragma unroll
            for (int e = 0; e < 4; ++e) { ln_s1 += d[e]; ln_s2 = fmaf(d[e], d[e], ln_s2); }
            xw[2 * (c & 3)] = pack_bf16x2(d[0] * ln_r, d[1] * ln_r);
            xw[2 * (c & 3) + 1] = pack_bf16x2(d[2] * ln_r, d[3] * ln_r);
            if ((c & 3) == 3 && own_ok) st_global_v8(xptr + ch * 64 + (c >> 2) * 32, xw);
          }
          __syncwarp();
        }
      }
      if constexpr (LNF == 2) {
        if (own_ok && n0 < p.N) {
          float2* rec = reinterpret_cast<float2*>(p.ln_out + static_cast<long long>(own_row) * (2 + 2 * p.ln_ns));
          rec[1 + (n0 >> 7)] = make_float2(ln_s1, ln_s2);
          if (n0 == 0) rec[0] = make_float2(ln_c, ln_r);
        }
      }
      }   // transposing epilogue
      if (!waited) { mbar_wait(tfull_bar(acc), acc_phase); tcgen05_fence_after(); }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));   // the leader's MMA issuer waits on it
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: nobody leaves while the peer may still touch its smem/TMEM
  if (warp == W_MMA) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_2cta<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---- host --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 3-D map over a batch of row-major bf16 matrices: dims {cols, rows, batch}; box {box_cols, box_rows, 1}; 128B swizzle.
int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint64_t batch,
                   uint64_t batch_stride_elems, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(SETOK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstr[2] = {ld_elems * 2, (batch > 1 ? batch_stride_elems : rows * ld_elems) * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SETOK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu batch=%llu)", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, (unsigned long long)batch);
  return SETOK_OK;
}

}  // namespace

int g_gemm_epi_mode = 0;
int g_gemm_epi_direct = -1;   // -1 automatic, 0 / 1 force the transposing / the row-owner epilogue (setok_debug_set_gemm_epi_direct)
int g_gemm_cta_group = 0;   // 0 = automatic; 1 forces single-CTA tiles (debug / A-B timing via setok_debug_set_gemm_cta_group)

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  SETOK_REQUIRE(g.A && g.W && g.D, SETOK_ERR_BAD_ARG, "gemm: null operand");
  SETOK_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, SETOK_ERR_BAD_ARG, "gemm: non-positive shape M=%d N=%d K=%d", g.M, g.N, g.K);
  // K may be ragged (TMA zero-fills the tail of the last 64-wide K block); N only needs 4-column granularity
  SETOK_REQUIRE(g.N % 4 == 0, SETOK_ERR_UNSUPPORTED, "gemm: N (%d) must be a multiple of 4", g.N);
  SETOK_REQUIRE(g.lda % 8 == 0 && g.ldw % 8 == 0 && g.lda >= g.K && g.ldw >= (g.w_mn_major ? g.N : g.K), SETOK_ERR_BAD_ARG,
                "gemm: lda/ldw must be multiples of 8 and cover a row (lda=%lld ldw=%lld K=%d N=%d)", (long long)g.lda, (long long)g.ldw, g.K, g.N);
  SETOK_REQUIRE(g.batch >= 1, SETOK_ERR_BAD_ARG, "gemm: batch must be >= 1");
  if (g.batch > 1) {
    SETOK_REQUIRE(g.a_batch_stride % 8 == 0 && g.w_batch_stride % 8 == 0 && g.d_batch_stride % 4 == 0, SETOK_ERR_BAD_ARG, "gemm: batch strides must keep 16-byte alignment");
    SETOK_REQUIRE(!g.m_dev && g.remap_P == 0 && g.r_batch_stride % 4 == 0, SETOK_ERR_UNSUPPORTED, "gemm: batched form takes no device row count / remap");
  }
  SETOK_REQUIRE(aligned16(g.A) && aligned16(g.W) && aligned16(g.D), SETOK_ERR_BAD_ARG, "gemm: operands must be 16-byte aligned");
  SETOK_REQUIRE(g.ldd % 4 == 0 && g.ldd >= g.N, SETOK_ERR_BAD_ARG, "gemm: ldd (%lld) must be a multiple of 4 and >= N", (long long)g.ldd);
  SETOK_REQUIRE(g.out_dtype == SETOK_F32 || g.out_dtype == SETOK_BF16, SETOK_ERR_BAD_ARG, "gemm: bad out_dtype %d", g.out_dtype);
  SETOK_REQUIRE(g.act >= 0 && g.act <= 2, SETOK_ERR_BAD_ARG, "gemm: bad activation %d", g.act);
  if (g.residual) {
    SETOK_REQUIRE(aligned16(g.residual) && g.ldr % 4 == 0, SETOK_ERR_BAD_ARG, "gemm: residual must be 16-byte aligned, ldr %% 4 == 0");
    SETOK_REQUIRE(g.residual_dtype == SETOK_F32 || g.residual_dtype == SETOK_BF16, SETOK_ERR_BAD_ARG, "gemm: bad residual dtype");
  }
  if (g.bias) SETOK_REQUIRE(aligned16(g.bias), SETOK_ERR_BAD_ARG, "gemm: bias must be 16-byte aligned");

  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, GemmDev);
  const int res_kind = g.residual ? (g.residual_dtype == SETOK_BF16 ? 1 : 2) : 0;
  const int out_f32 = g.out_dtype == SETOK_F32 ? 1 : 0;
  // Row-owner epilogue (no shared-memory transpose): needs whole 32-column chunks and 32-byte aligned rows of the output /
  // residual.  Measured at the ViT-L shapes (tools/bench_gemm_epilogue.py, B200, M = 65792): bf16 outputs gain 2-25 % (qkv
  // -4 %, fc1 -7 %, GELU(erf) fc1 -10 %, bf16-stream out_proj -25 %, fc2 -2 %); float32 outputs written 32 bytes per row and
  // instruction lose on the HBM-bound K = 1024 shape (f32-stream out_proj +25 %, plain f32 output +5 %) and gain 3 % at
  // K = 4096 (fc2): the automatic choice follows those measurements.
  bool direct = g.N % 32 == 0 && (g_gemm_epi_direct == 1 || (g_gemm_epi_direct < 0 && (!out_f32 || (res_kind == 2 && g.K >= 2048))));
  {
    const int de = out_f32 ? 4 : 2, re = res_kind == 1 ? 2 : 4;
    direct = direct && (g.ldd * de) % 32 == 0 && (reinterpret_cast<uintptr_t>(g.D) % 32) == 0 && (g.d_batch_stride * de) % 32 == 0;
    if (g.residual) direct = direct && (g.ldr * re) % 32 == 0 && (reinterpret_cast<uintptr_t>(g.residual) % 32) == 0 && (g.r_batch_stride * re) % 32 == 0;
  }
  // specialised epilogues for the combinations the tokenizer path launches; everything else -> generic
  // CTA pairs whenever there is at least one full 256-row tile; single CTAs for short row counts
  const int cg = (g.M >= 160 && g_gemm_cta_group != 1) ? 2 : 1;
#define SETOK_PICK(A, R, O, P) (cg == 2 ? gemm_bf16_tcgen05_kernel<2, A, R, O, P, false> : gemm_bf16_tcgen05_kernel<1, A, R, O, P, false>)
#define SETOK_PICK_D(A, R, O) (direct ? (cg == 2 ? gemm_bf16_tcgen05_kernel<2, A, R, O, 0, true> : gemm_bf16_tcgen05_kernel<1, A, R, O, 0, true>) : SETOK_PICK(A, R, O, 0))
  KernelFn fn = SETOK_PICK(-1, -1, -1, -1);
  if (g.remap_P == 0) {
    if (g.act == SETOK_ACT_NONE && res_kind == 0 && !out_f32) fn = SETOK_PICK_D(0, 0, 0);              // qkv
    else if (g.act == SETOK_ACT_NONE && res_kind == 1 && !out_f32) fn = SETOK_PICK_D(0, 1, 0);         // bf16-stream out_proj / fc2
    else if (g.act == SETOK_ACT_QUICK_GELU && res_kind == 0 && !out_f32) fn = SETOK_PICK_D(1, 0, 0);   // ViT fc1
    else if (g.act == SETOK_ACT_GELU_ERF && res_kind == 0 && !out_f32) fn = SETOK_PICK_D(2, 0, 0);     // head / projector / decoder fc1
    else if (g.act == SETOK_ACT_NONE && res_kind == 2 && out_f32) fn = SETOK_PICK_D(0, 2, 1);          // f32-stream out_proj / fc2, head proj / fc2
    else if (g.act == SETOK_ACT_NONE && res_kind == 0 && out_f32) fn = SETOK_PICK_D(0, 0, 1);          // out / projector last
    else if (g.act == SETOK_ACT_NONE && res_kind == 1 && out_f32) fn = SETOK_PICK_D(0, 1, 1);          // Q-Former dense + residual -> post-LN
  } else if (g.act == SETOK_ACT_NONE && res_kind == 2 && out_f32) {
    fn = SETOK_PICK(0, 2, 1, 1);                                                                         // patch embedding
  }
  const int lnf = g.ln_in ? (g.ln_out ? 2 : 1) : 0;
  if (lnf != 0) {
    bool ok = g.N % 32 == 0 && g.batch == 1 && g.remap_P == 0 && g.ln_C > 0 && aligned16(g.ln_in);
    const int de = out_f32 ? 4 : 2;
    ok = ok && (g.ldd * de) % 32 == 0 && (reinterpret_cast<uintptr_t>(g.D) % 32) == 0;
    if (lnf == 1) ok = ok && g.ln_s && aligned16(g.ln_s) && res_kind == 0 && !out_f32 && (g.act == SETOK_ACT_NONE || g.act == SETOK_ACT_QUICK_GELU);
    else ok = ok && g.xhat && res_kind == 2 && out_f32 && g.act == SETOK_ACT_NONE && g.N == g.ln_C && (g.ld_xhat * 2) % 32 == 0 &&
              (reinterpret_cast<uintptr_t>(g.xhat) % 32) == 0 && (g.ldr * 4) % 32 == 0 && (reinterpret_cast<uintptr_t>(g.residual) % 32) == 0 &&
              aligned16(g.ln_out) && g.ln_out != g.ln_in;
    SETOK_REQUIRE(ok, SETOK_ERR_UNSUPPORTED, "gemm: LayerNorm fold (%s side) unsupported for this shape / layout (N=%d C=%d act=%d)",
                  lnf == 1 ? "consuming" : "producing", g.N, g.ln_C, g.act);
    // producing side: the automatic choice of the epilogue follows the plain f32-stream GEMMs (transposing at K = 1024, row-owner at K = 4096)
    direct = lnf == 1 || (g_gemm_epi_direct >= 0 ? g_gemm_epi_direct == 1 : g.K >= 2048);
#define SETOK_PICK_LN(A, R, O, F) (cg == 2 ? gemm_bf16_tcgen05_kernel<2, A, R, O, 0, true, F> : gemm_bf16_tcgen05_kernel<1, A, R, O, 0, true, F>)
    if (lnf == 2) fn = direct ? SETOK_PICK_LN(0, 2, 1, 2) : (cg == 2 ? gemm_bf16_tcgen05_kernel<2, 0, 2, 1, 0, false, 2> : gemm_bf16_tcgen05_kernel<1, 0, 2, 1, 0, false, 2>);
    else if (g.act == SETOK_ACT_QUICK_GELU) fn = SETOK_PICK_LN(1, 0, 0, 1);
    else fn = SETOK_PICK_LN(0, 0, 0, 1);
#undef SETOK_PICK_LN
  }
#undef SETOK_PICK_D
#undef SETOK_PICK
  const int smem_bytes = cg == 2 ? Cfg<2>::SMEM_BYTES : Cfg<1>::SMEM_BYTES;
  SETOK_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(fn), smem_bytes));
  CUtensorMap tmA, tmB;
  SETOK_TRY(make_tmap_bf16(&tmA, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda, (uint64_t)g.batch, (uint64_t)g.a_batch_stride, BK, BM));
  if (g.w_mn_major)   // W is [K, N]: box = 64 N-columns (128 B) x 64 K-rows, one per 64-column atom of the B tile
    SETOK_TRY(make_tmap_bf16(&tmB, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldw, (uint64_t)g.batch, (uint64_t)g.w_batch_stride, 64, BK));
  else
    SETOK_TRY(make_tmap_bf16(&tmB, g.W, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldw, (uint64_t)g.batch, (uint64_t)g.w_batch_stride, BK, BN / cg));
  GemmDev p;
  p.D = g.D; p.ldd = g.ldd; p.bias = g.bias; p.res = g.residual; p.ldr = g.ldr; p.m_dev = g.m_dev;
  p.M = g.M; p.N = g.N; p.K = g.K; p.act = g.act; p.out_f32 = g.out_dtype == SETOK_F32;
  p.res_kind = res_kind;
  p.remap_P = g.remap_P;
  p.batch = g.batch; p.d_batch_stride = g.d_batch_stride; p.r_batch_stride = g.r_batch_stride; p.w_mn_major = g.w_mn_major;
  p.epi_mode = g_gemm_epi_mode;
  p.direct = direct ? 1 : 0;
  p.ln_in = g.ln_in; p.ln_out = g.ln_out; p.ln_s = g.ln_s; p.xhat = static_cast<bf16*>(g.xhat); p.ld_xhat = g.ld_xhat;
  p.ln_eps = g.ln_eps; p.ln_invC = g.ln_C > 0 ? 1.0f / static_cast<float>(g.ln_C) : 0.f; p.ln_ns = ceil_div(g.ln_C > 0 ? g.ln_C : 1, 128);
  const int tiles = ceil_div(g.M, BM * cg) * ceil_div(g.N, BN) * g.batch;
  const int max_groups = num_sms() / cg;
  const int grid = (tiles < max_groups ? tiles : max_groups) * cg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  SETOK_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, tmA, tmB, p));
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}

}  // namespace setok

extern "C" int setok_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd, int out_dtype,
                               const float* bias, const void* residual, int64_t ldr, int residual_dtype, int act, int M,
                               int N, int K, const int32_t* m_dev, setok_stream_t stream) {
  setok::GemmArgs g{A, lda, W, ldw, D, ldd, out_dtype, bias, residual, ldr, residual_dtype, act, M, N, K, m_dev, 0};
  return setok::launch_gemm(g, static_cast<cudaStream_t>(stream));
}

extern "C" int setok_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd, int out_dtype,
                                  const float* bias, const void* residual, int64_t ldr, int residual_dtype, int act, int M, int N, int K,
                                  const float* ln_in, float* ln_out, const float* ln_s, void* xhat, int64_t ld_xhat, float ln_eps,
                                  int ln_C, setok_stream_t stream) {
  SETOK_REQUIRE(ln_in != nullptr && (ln_s != nullptr) != (ln_out != nullptr), SETOK_ERR_BAD_ARG,
                "gemm_ln: records plus exactly one of ln_s (consuming side) / ln_out (producing side)");
  setok::GemmArgs g{A, lda, W, ldw, D, ldd, out_dtype, bias, residual, ldr, residual_dtype, act, M, N, K, nullptr, 0};
  g.ln_in = ln_in; g.ln_out = ln_out; g.ln_s = ln_s; g.xhat = xhat; g.ld_xhat = ld_xhat; g.ln_eps = ln_eps; g.ln_C = ln_C;
  return setok::launch_gemm(g, static_cast<cudaStream_t>(stream));
}

extern "C" int setok_gemm_bf16_batched(const void* A, int64_t lda, int64_t a_batch_stride, const void* W, int64_t ldw, int64_t w_batch_stride,
                                       int w_mn_major, void* D, int64_t ldd, int64_t d_batch_stride, int out_dtype, const float* bias,
                                       int act, int batch, int M, int N, int K, setok_stream_t stream) {
  setok::GemmArgs g{A, lda, W, ldw, D, ldd, out_dtype, bias, nullptr, 0, 0, act, M, N, K, nullptr, 0};
  g.batch = batch; g.a_batch_stride = a_batch_stride; g.w_batch_stride = w_batch_stride; g.d_batch_stride = d_batch_stride;
  g.w_mn_major = w_mn_major;
  return setok::launch_gemm(g, static_cast<cudaStream_t>(stream));
}

extern "C" void setok_debug_set_gemm_cta_group(int cg) { setok::g_gemm_cta_group = cg; }
extern "C" void setok_debug_set_gemm_epi_mode(int mode) { setok::g_gemm_epi_mode = mode; }
extern "C" void setok_debug_set_gemm_epi_direct(int v) { setok::g_gemm_epi_direct = v; }
