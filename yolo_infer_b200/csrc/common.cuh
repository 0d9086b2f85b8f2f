// Shared helpers for liby11_b200 (sm_100a only).
#pragma once
#include <cuda.h>  // CUtensorMap + enums only; the driver entry point is resolved at run time
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/y11.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liby11_b200 is written for sm_100a (Blackwell) only"
#endif

// ------------------------------------------------------------------------------------------ host
void y11_set_error(const char* fmt, ...);

#define Y11_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      y11_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

#define Y11_REQUIRE(cond, ...)                                                                 \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      y11_set_error(__VA_ARGS__);                                                              \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

typedef CUresult (*y11_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

struct y11_engine {
  int device;
  int num_sms;
  y11_encode_tiled_fn encode_tiled;
  int* dev_error_flag;   // device alias of host_error_flag: written by kernels before __trap()
  int* host_error_flag;  // mapped pinned memory (readable after a trapped kernel poisoned the context)
};

static inline int y11_ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: opt a kernel in once for every device it is used
// on (bit d of a per-call-site mask = done for device d; atomic, so engines on several GPUs / threads of one process are
// safe).  The current device is the engine's (y11_create / the caller's cudaSetDevice).
#define Y11_OPT_IN_SMEM(fn, bytes)                                                                         \
  do {                                                                                                     \
    static std::atomic<unsigned long long> _y11_done{0ull};                                                \
    int _y11_dev = 0;                                                                                      \
    Y11_CHECK_CUDA(cudaGetDevice(&_y11_dev));                                                              \
    const unsigned long long _y11_bit = 1ull << (_y11_dev & 63);                                           \
    if (!(_y11_done.load(std::memory_order_acquire) & _y11_bit)) {                                         \
      Y11_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _y11_done.fetch_or(_y11_bit, std::memory_order_release);                                             \
    }                                                                                                      \
  } while (0)

// ---------------------------------------------------------------------------------------- device
#ifdef __CUDACC__
namespace y11 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// SiLU = x*sigmoid(x) = h + h*tanh(h), h = x/2: ONE MUFU op (tanh.approx, max rel. error 2^-11) + FMUL + FFMA.
// The conv epilogue is instruction-issue bound on the 1x1 layers (ncu: 13 instructions per output element with
// __expf/__fdividef, of which 4 FMUL + FSETP were denormal-range fix-ups), so this is the hot scalar of the network.
// Absolute error <= 2.5e-4*|x| for x < 0 (cancellation in 1+tanh), i.e. below one bf16 ulp of the activations it feeds.
__device__ __forceinline__ float silu(float x) {
#ifdef Y11_SILU_EX2
  // x * sigmoid(x) = x / (1 + 2^(-x*log2 e)): two full-rate MUFU ops (EX2, RCP) + FMUL/FADD/FMUL
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return x * r;
#else
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
#endif
}
// magic-number division for the persistent tile loops (exact for n, d < 2^21): q = (n * ceil(2^42/d)) >> 42
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint64_t magic) { return (uint32_t)(((uint64_t)n * magic) >> 42); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issued instruction) -------------
// A pair lives in one 64-bit register (lo = first element).  Every op rounds each element exactly like its scalar form
// (fma.rn / mul.rn / add.rn, no flush to zero), so code that moves from the scalar to the packed form keeps its results bit
// for bit; what it saves is issue slots, which is what bounds the depthwise conv and the conv / stem epilogues.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// one 32-bit word of two bf16 -> the pair as fp32
__device__ __forceinline__ f32x2 f2_from_bf16x2(uint32_t v) { return f2_pack(bf16_lo(v), bf16_hi(v)); }
__device__ __forceinline__ uint32_t f2_to_bf16x2(f32x2 v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
#ifndef Y11_SILU_EX2
// silu() of both elements: FMUL2 + 2 MUFU + FFMA2 instead of 2 x (FMUL + MUFU + FFMA); same bits as silu()
__device__ __forceinline__ f32x2 silu2(f32x2 x) {
  const f32x2 h = f2_mul(x, f2_pack(0.5f, 0.5f));
  float hl, hh, tl, th;
  f2_unpack(h, hl, hh);
  asm("tanh.approx.f32 %0, %1;" : "=f"(tl) : "f"(hl));
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hh));
  return f2_fma(h, f2_pack(tl, th), h);
}
#else
__device__ __forceinline__ f32x2 silu2(f32x2 x) {
  float lo, hi;
  f2_unpack(x, lo, hi);
  return f2_pack(silu(lo), silu(hi));
}
#endif

// ---- mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait WITH a suspend-time hint: the warp is parked by the hardware until the phase completes (or the hint, in ns,
// elapses) instead of spinning.  Without the hint the wait returns almost immediately and the polling loops of the ten
// warps of each CTA ate most of the SM's issue slots (ncu: 34 M of 71 M executed instructions of a 1x1 layer were
// try_wait/branch/clock spin instructions, schedulers ~85 % busy, real work starved - profiles/r01c_summary.md).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (with a flag for the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 1.9 GHz
      if (err_flag) *reinterpret_cast<volatile int*>(err_flag) = code;  // mapped pinned host memory
      __threadfence_system();
      __trap();
    }
  }
}

// ---- TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 in, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with e4m3 operands (kind::f8f6f4: K = 32 bytes per instruction, fp32 accumulate).
__device__ __forceinline__ void umma_e4m3(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, swizzled UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor layout):
// [0,14) addr>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout
__device__ __forceinline__ uint64_t make_umma_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (static_cast<uint64_t>(layout_type) << 61);
}
// Un-swizzled ("interleaved") K-major descriptor: core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
// LBO = byte distance between core matrices adjacent in K, SBO = between core matrices adjacent in M/N (8-row groups).
// The start address only needs 16-byte alignment, which is what lets a conv tap be a plain address offset into a halo tile.
__device__ __forceinline__ uint64_t make_umma_desc_interleaved(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// 16-byte async copy global -> shared (L2 only), zero-filled when src_bytes == 0; completion is reported to an mbarrier by
// cp_async_mbar_arrive (one pending arrival of the barrier's expected count per calling thread, no increment).
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc_bf16_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// kind::f8f6f4 instruction descriptor: D=f32, A=B=e4m3 (format code 0), both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc_e4m3_m128(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// four fp32 -> four e4m3 bytes (byte i = value i), round to nearest even, saturating at +-448
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return (uint32_t)lo | ((uint32_t)hi << 16);
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// Every kernel of the plan is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may be scheduled
// (and run their prologue: barrier init, TMEM alloc, tensor-map prefetch, weight staging) while the previous kernel of the
// stream drains.  pdl_wait() blocks until that previous kernel has completed and its writes are visible; it must precede
// the first access to anything the predecessor wrote.  pdl_trigger() lets the NEXT kernel start being scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- misc -----------------------------------------------------------------------------------
// One lane of a fully converged warp (warp-uniform branch): lets the compiler keep tcgen05/TMA operands in uniform registers.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

}  // namespace y11
#endif
