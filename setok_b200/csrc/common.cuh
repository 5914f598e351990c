// common.cuh — shared host/device helpers for libsetok_b200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "../../include/setok_b200.h"

namespace setok {

// ---------------------------------------------------------------------------------------------
// Host: error reporting (thread-local message), launch accounting
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

#define SETOK_CUDA_OK(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::setok::fail(SETOK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                   \
  } while (0)

// Every kernel launch goes through this: counts the launch and turns launch errors into a status.
#define SETOK_LAUNCH_CHECK()                                                                      \
  do {                                                                                            \
    ::setok::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess)                                                                        \
      return ::setok::fail(SETOK_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                    \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                           \
  } while (0)

#define SETOK_REQUIRE(cond, code, ...)                    \
  do {                                                    \
    if (!(cond)) return ::setok::fail(code, __VA_ARGS__); \
  } while (0)

#define SETOK_TRY(expr)          \
  do {                           \
    int _s = (expr);             \
    if (_s != SETOK_OK) return _s; \
  } while (0)

// NVTX range around each SURVEY.md 8 row's entry point (a no-op unless a profiler is attached): SETOK_NVTX("a1+a2 ...").
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define SETOK_NVTX(name) ::setok::NvtxRange _setok_nvtx_range(name)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int num_sms();                                            // of the current device (cached per device)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (current device, kernel)
int ensure_dynamic_smem(const void* kernel, int bytes);

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Arena {
  uint8_t* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base(static_cast<uint8_t*>(p)), cap(n), off(0) {}
  template <class T>
  T* take(size_t count) {
    size_t bytes = round_up(count * sizeof(T), 256);
    if (base != nullptr && off + bytes > cap) { off = cap + 1; return nullptr; }
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return r;
  }
  bool overflow() const { return base != nullptr && off > cap; }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
template <class T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// x * sigmoid(1.702 x) with sigmoid(z) = 0.5 tanh(z/2) + 0.5: one MUFU (tanh.approx, rel. error ~2^-11) instead of the
// ex2 + rcp pair -- the epilogue of the fc1 GEMM is MUFU/issue-bound, and its output is rounded to bf16 anyway
__device__ __forceinline__ float act_quick_gelu(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// GELU(erf) (nn.GELU: module.py:29-45, multimodal_projector/builder.py:45-59, timm Mlp) as x * sigmoid(2 x p(x^2)) with
// p(z) = c1 + c3 z + c5 z^2 fitted (minimax, x in [-8, 8]; z clamped at 64 beyond, where the sigmoid is saturated) to
// 0.5 x (1 + erf(x / sqrt 2)): |error| <= 2.6e-5 absolute, below the bf16 rounding of every consumer for |y| > 0.01.
// 7 ALU + 2 MUFU (ex2, rcp) per element: the Abramowitz-Stegun 7.1.26 form it replaces (1.5e-7, ~16 ALU + 2 MUFU) made the
// GELU GEMM epilogues issue-bound (detokenizer fc1 at K = 768 ran at 60 % of the qkv GEMM's rate).  The constants carry the
// factor -2 log2(e) of the exponent.
__device__ __forceinline__ float act_gelu_erf(float x) {
  const float z = fminf(x * x, 64.0f);
  float p = fmaf(z, 0.001014265581034124f, -0.10677573829889297f);
  p = fmaf(p, z, -2.301121234893799f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// ---- programmatic dependent launch ------------------------------------------------------------
// Kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (prologue: barrier init, TMEM
// allocation, descriptor prefetch) while the previous kernel of the stream is still draining; pdl_wait() returns once that
// kernel has completed and its writes are visible.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (an error the host sees), never as
// a hung GPU.  The bound is wall time (4 s), far above any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0x3ffu) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();   // no printf here: an ABI call would force spills around every wait
    }
  }
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint32_t bar, uint32_t smem_dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* tmap, uint32_t bar, uint32_t smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// One lane of a fully converged warp (warp-uniform code keeps descriptor arithmetic in the uniform datapath;
// only the tcgen05 instruction itself is predicated on the elected lane).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {            // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 lanes x K/2 32-bit columns, two bf16 per column) comes from
// tensor memory -- used to feed the softmax probabilities straight back into the P.V product.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns, registers -> TMEM
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// one fp32 column of 32 lanes
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
// 32 lanes x 8 consecutive columns, registers -> TMEM
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// named barrier among `nthreads` threads (whole warps) of the CTA; id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per thread per instruction.  The pointer
// must be 32-byte aligned.
#ifndef SETOK_EPI_STREAMING
#define SETOK_EPI_STREAMING 0   // 1: the GEMM epilogues' activation loads / stores carry the evict-first hint (ld/st.global.cs): every activation
#endif                          // is read once by the next kernel, only the operands the concurrent tiles share are worth keeping in L2
#if SETOK_EPI_STREAMING
#define SETOK_CS ".cs"
#else
#define SETOK_CS ""
#endif
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t* r) {
  asm volatile("ld.global" SETOK_CS ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
  asm volatile("st.global" SETOK_CS ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (tile rows of 64 bf16 = 128 B, 8-row
// swizzle atoms 1024 B apart).  Field layout: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, 128-byte-swizzled operand: tile rows are K indices, each row holds 64 contiguous MN elements
// (128 B); 8-row groups 1024 B apart (SBO); LBO = byte stride between 64-element MN atoms (unused for MN extent 64).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes = 16) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, A K-major,
// B K-major unless b_mn_major (bit 16), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool b_mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::f16 with IEEE half operands (A / B format fields 0) -- otherwise as umma_idesc_bf16
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool b_mn_major = false) {
  return (1u << 4) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// two floats -> packed IEEE halves (round to nearest, finite saturation: |x| > 65504 clamps instead of becoming inf)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
  float2 f;
  asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(f.x), "=f"(f.y) : "r"(u));
  return f;
}

// ---- CTA pairs (cta_group::2): cluster helpers, peer-barrier signalling, 2-SM TMA / MMA / TMEM forms -----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  // default (.release.cta) semantics on purpose: a cluster-scope release costs a MEMBAR+ERRBAR per arrive, and the
  // accumulator hand-off this guards is ordered by tcgen05.fence::before_thread_sync / ::after_thread_sync instead
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(const void* tmap, uint32_t cluster_bar_addr, uint32_t smem_dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const void* tmap, uint32_t cluster_bar_addr, uint32_t smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_result_addr) {   // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// issued by the leader CTA only: D (M = 256 across the pair) += A (128 rows per CTA) * B (N/2 rows per CTA)
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the same-offset mbarrier of every CTA in `cta_mask` once the pair's previously issued MMAs completed
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// Internal launch API (host), shared between translation units
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  const void* A; int64_t lda;
  const void* W; int64_t ldw;
  void* D; int64_t ldd; int out_dtype;
  const float* bias;
  const void* residual; int64_t ldr; int residual_dtype;
  int act;
  int M, N, K;
  const int32_t* m_dev;
  int remap_P;   // >0: patch-embed mode: out row r -> r + r/P + 1, residual row -> 1 + r%P (pos-emb broadcast)
  // batched GEMM: `batch` independent problems, operand b at base + b * stride (elements).  bias is shared; no residual.
  int batch = 1;
  int64_t a_batch_stride = 0, w_batch_stride = 0, d_batch_stride = 0, r_batch_stride = 0;
  // W given as [K, N] row-major (N contiguous) instead of [N, K]: the MN-major B operand form (e.g. V in P.V)
  int w_mn_major = 0;
  // LayerNorm folded into the GEMMs around it (SETOK_VIT_LN_FOLD; see the epilogue comment in gemm_tcgen05.cu).  Per-row
  // records of 2 + 2*ns floats, ns = ceil(C / 128): {c, r, (s1, s2) x ns}.
  //   consumer (ln_in && ln_s):  D = act(a_row * (A W'^T) + b_row * ln_s + bias), A = xhat, W' = gamma (.) W, bias = W beta + b
  //   producer (ln_in && ln_out && xhat): D = x_new = acc + bias + residual as usual, plus xhat = bf16((x_new - c) * r) and the
  //                                      row's partial sums of (x_new - c), (x_new - c)^2 over this warp's 128 columns
  const float* ln_in = nullptr;
  float* ln_out = nullptr;
  const float* ln_s = nullptr;
  void* xhat = nullptr; int64_t ld_xhat = 0;
  float ln_eps = 0.f; int ln_C = 0;
};
int launch_gemm(const GemmArgs& g, cudaStream_t stream);
extern int g_pdl;   // 1: launch the tower's kernels with programmatic stream serialization (setok_debug_set_pdl)

// <<<>>> with the programmatic-stream-serialization attribute (the kernel must call pdl_wait() before it touches memory)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// dpc_fused.cu: whole clustering (a3+a4) in one persistent kernel for N <= 256
bool dpc_fused_supported(int N, int C, int k);
int launch_dpc_fused(const void* feats, int feat_dtype, const float* pos, const float* noise, const float* token_mask, int B, int N, int C,
                     int k, float threshold, int min_cluster_num, float* x_pos, int64_t* idx_cluster, float* score, int64_t* index_down,
                     int32_t* num_clusters, cudaStream_t stream);

int launch_layernorm(const void* in, int in_dtype, void* out, int out_dtype, const float* gamma, const float* beta,
                     float eps, int rows, int C, const int32_t* gather, const int32_t* m_dev, cudaStream_t stream);
int launch_attention_tcgen05(const void* qkv, void* out, int B, int T, int C, int heads, float scale, cudaStream_t stream);
bool attention_fullrow_supported(int T);
int launch_attention_fullrow(const void* qkv, void* out, int B, int T, int C, int heads, float scale, cudaStream_t stream);
int launch_attention(const void* qkv, void* out, int rows, int C, int heads, float scale, const int32_t* seg_off,
                     const int32_t* row_seg, int uniform_T, const int32_t* m_dev, cudaStream_t stream);
int launch_cross_attention(const void* q, const void* kv, void* out, int rows, int Q, int C, int heads, float scale,
                           const int32_t* offsets, int kv_rows, cudaStream_t stream);
int launch_cross_attention_tcgen05(const void* q, const void* kv, void* out, int B, int Q, int C, int heads, float scale,
                                   const int32_t* offsets, int kv_rows, cudaStream_t stream);

}  // namespace setok
