// Implicit-GEMM convolution on Blackwell 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA in AND out).
//
// Replaces, for every dense conv of the fused YOLO11 network, what the reference executes as
// ultralytics Conv.forward_fuse -> torch conv2d (+SiLU) (+ Bottleneck/PSABlock residual add) (+ Concat copy)
// (SURVEY.md section 8a rows a6-a8, a11).
//
// GEMM view (NHWC bf16 activations, fp32 accumulation):
//   M = output pixels   : a tile is 128 rows = a Tw x Th x Tn box of (w, h, image), chosen per layer to waste few rows
//   N = output channels : BN <= 128 columns per tile; accumulator = 128 TMEM lanes x BN fp32 columns
//   K = (kh, kw, cin)   : one pipeline stage = one filter tap x Cc (16/32/64) input channels
// The A operand is never materialised (no im2col buffer): for each tap the TMA engine loads the SHIFTED
// Tw x Th x Tn x Cc box of the input straight into the canonical K-major swizzled smem layout UMMA reads;
// rows/cols outside the image are zero-filled by TMA, which is exactly the conv zero padding.  Stride-2
// convs address four parity sub-grids of the input (one tensor map each) so they are shifted boxes too.
//
// v2 (round 1, after the first ncu pass showed the one-tile-per-CTA version latency bound - profiles/r01_summary.md):
//   * PERSISTENT: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...; barrier init / TMEM alloc / tensor-map
//     prefetch are paid once per SM instead of once per 128 pixels, and the TMA producer runs ahead across tile borders.
//   * TWO TMEM accumulator stages: the MMA warp starts tile i+1 while the epilogue drains tile i.
//   * 8 epilogue warps (2 per TMEM lane quadrant, each pair splits the columns) so SiLU/bias/convert has 2 warps/SMSP.
//   * Epilogue writes through swizzled shared-memory staging + TMA STORE (cp.async.bulk.tensor ... global.shared::cta):
//     full-sector writes at the channel offset of the destination view (this is how Concat/chunk cost nothing), and
//     ragged tiles are clipped by the TMA unit instead of per-thread predicates.
//
// Warp roles (320 threads): warp0 = TMA producer (1 lane), warp1 = MMA issuer (1 lane), warps2-9 = epilogue;
// warp2 also owns TMEM alloc/dealloc.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ops.h"

#ifdef Y11_TRACE
// debug timeline: CTA 0 records (event, clock) pairs per role into a global buffer (set with y11_debug_set_trace)
static long long* g_trace_buf = nullptr;
extern "C" void y11_debug_set_trace(void* dev_ptr) { g_trace_buf = static_cast<long long*>(dev_ptr); }
#define TRACE(role, ev)                                                            \
  do {                                                                             \
    if (p.trace && blockIdx.x == 0 && trc < 2000) {                                 \
      p.trace[(role) * 4096 + 2 * trc] = (ev);                                     \
      p.trace[(role) * 4096 + 2 * trc + 1] = clock64();                            \
      ++trc;                                                                       \
    }                                                                              \
  } while (0)
#else
#define TRACE(role, ev) do {} while (0)
#endif

namespace {

constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 32;           // small-channel layers have 5-9 KB stages: bytes in flight, not stage count, hide latency
constexpr uint32_t kHeaderBytes = 3072;  // mbarriers + tmem pointer (1 KB) + LSU-mode position table (2 KB)
constexpr uint32_t kSmemBudget = 200 * 1024;
// Tile queue: the producer warp owns the CTA's tile sequence and publishes it to the MMA warp and the epilogue warps through
// a small shared-memory ring guarded by mbarriers.  Default: the static walk blockIdx.x, +gridDim.x, ...  With
// Y11_DYN_TILES=1 every tile after the first comes from a global atomic counter (a CTA that becomes resident late - kernels of
// another stream, NCCL CTAs, the drain of a non-persistent predecessor - then simply takes fewer tiles); measured neutral
// on B200 (see api.cu), so it is off by default.
constexpr int kTileQ = 8;
constexpr uint32_t kTqFullOff = 640, kTqEmptyOff = 704, kTqTileOff = 768;  // inside the first KB of the header

__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ int ld_shared_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_s32(uint32_t addr, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// tcgen05.mma with the two shared-memory descriptors given as (low, high) 32-bit halves: the low word (start address, LBO)
// is what changes between UMMAs, the high word (SBO, version, swizzle mode) is a per-kernel constant.
template <bool kF8>
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  if (kF8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// one TMA pipeline stage = KK UMMAs, +32 B (= 2 descriptor units) inside the swizzle row per UMMA
template <int KK, bool kQ>
__device__ __forceinline__ void tma_issue(uint32_t tmem_acc, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc0,
                                          bool f8) {
  if (kQ && f8) {
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) umma_lohi<true>(tmem_acc, a_lo + 2u * kk, hi, b_lo + 2u * kk, hi, idesc, kk ? 1u : acc0);
  } else {
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) umma_lohi<false>(tmem_acc, a_lo + 2u * kk, hi, b_lo + 2u * kk, hi, idesc, kk ? 1u : acc0);
  }
}
// All K stages of one tile in TMA mode, run by ONE thread: wait for the stage, KK UMMAs, release the stage; accumulator-full
// commit after the last stage.
struct TmaTileArgs {
  uint32_t tmem_acc, ring_lo0, b_off, hi, idesc, stage_units, n_stages, full_bar, empty_bar, accf;
  int k_iters;
  int* err_flag;
  uint32_t bres_lo, b_step;  // resident weights (bres_lo != 0): descriptor low word of K stage 0 and the step between stages
};
template <int KK, bool kQ>
__device__ __forceinline__ void tma_tile_issue(const TmaTileArgs& t, uint32_t stage, uint32_t phase, bool f8) {
  uint32_t a_lo = t.ring_lo0 + stage * t.stage_units, fb = t.full_bar + 8 * stage, eb = t.empty_bar + 8 * stage;
  for (int k = 0; k < t.k_iters; ++k) {
    y11::mbar_wait(fb, phase, t.err_flag, 102);
    y11::tc_fence_after();
    tma_issue<KK, kQ>(t.tmem_acc, a_lo, t.bres_lo ? t.bres_lo + (uint32_t)k * t.b_step : a_lo + t.b_off, t.hi, t.idesc, k != 0, f8);
    y11::umma_commit(eb);  // frees the smem slot once these MMAs retire
    if (++stage == t.n_stages) { stage = 0; phase ^= 1; a_lo = t.ring_lo0; fb = t.full_bar; eb = t.empty_bar; }
    else { a_lo += t.stage_units; fb += 8; eb += 8; }
  }
  y11::umma_commit(t.accf);  // accumulator complete
}
// one 3x3 halo tile with ONE weight chunk per tap (cin = 16 * KK): 9 * KK UMMAs with compile-time tap offsets
template <int KK>
__device__ __forceinline__ void halo3_issue(uint32_t tmem_acc, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t b_step,
                                            uint32_t k_step, uint32_t idesc) {
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t at = a_lo + (uint32_t)((tap / 3) * 10 + tap % 3);
#pragma unroll
    for (int kk = 0; kk < KK; ++kk)
      umma_lohi<false>(tmem_acc, at + (uint32_t)kk * k_step, a_hi, b_lo + 2u * kk, b_hi, idesc, (tap | kk) ? 1u : 0u);
    b_lo += b_step;
  }
}

// TMA-halo tile (cin = 64, round 2): the halo sits in the 128-byte-SWIZZLED K-major layout, one 128-byte row per halo pixel,
// exactly as ONE TMA box load {64 ch, Tw+2, Th+2} writes it.  Tap (kh, kw) is the start address moved by (kh * 10 + kw) ROWS and the
// 8-row groups follow at SBO = 10 rows = 1280 B - neither a multiple of the 1024-byte swizzle atom.  Measured on B200
// (tools/halo_sw_probe.py): tcgen05.mma applies the 128-byte swizzle to the ABSOLUTE shared-memory address bits, so any
// 128-byte-aligned start and any SBO that is a multiple of 128 B read what TMA wrote, with the descriptor's base-offset field left 0
// (setting it to the start address's row phase gives wrong results).  Same K order (tap, then channels) as the other modes.
template <int KK>
__device__ __forceinline__ void halo3_issue_sw(uint32_t tmem_acc, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t b_step,
                                               uint32_t idesc) {
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t at = a_lo + (uint32_t)((tap / 3) * 10 + tap % 3) * 8u;  // 128-byte rows in 16-byte descriptor units
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) umma_lohi<false>(tmem_acc, at + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc, (tap | kk) ? 1u : 0u);
    b_lo += b_step;
  }
}

// 16 accumulator columns of one tile row: +bias, (+pre-activation term), SiLU, (+residual), convert, and the swizzled
// 16-byte stores into the staging row at `dst`; `unit0` = index of the first 16-byte unit of these columns in the row.
template <bool kQ>  // kQ: the e4m3 paths (dequantisation scale, e4m3 stores) are compiled in; false = the bf16 kernels, unchanged
__device__ __forceinline__ void epi16(const ConvTcParams& p, const uint32_t (&v)[16], const uint4 r0, const uint4 r1, int n,
                                      uint32_t dst, uint32_t unit0, uint32_t swz) {
  using namespace y11;
  // Round 2: the 16 columns are handled as 8 packed fp32 pairs (FADD2 / FMUL2 / FFMA2): bias, SiLU and the residual cost half
  // the issue slots; every element goes through the same IEEE operations as before, so the results are bit-identical.
  f32x2 f[8];
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
  if (kQ && p.cscale) {  // fp8 operands: accumulator * (activation scale * weight scale of the channel) + bias
    const float4* c4 = reinterpret_cast<const float4*>(p.cscale + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 bb = __ldg(b4 + i), cc = __ldg(c4 + i);
      f[2 * i + 0] = f2_fma(f2_pack(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])), f2_pack(cc.x, cc.y), f2_pack(bb.x, bb.y));
      f[2 * i + 1] = f2_fma(f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), f2_pack(cc.z, cc.w), f2_pack(bb.z, bb.w));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 bb = __ldg(b4 + i);
      f[2 * i + 0] = f2_add(f2_pack(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])), f2_pack(bb.x, bb.y));
      f[2 * i + 1] = f2_add(f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), f2_pack(bb.z, bb.w));
    }
  }
  const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  if (p.res_pre) {  // up2(W_up . p) of a folded Upsample+Concat: part of the pre-activation sum
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f2_add(f[i], f2_from_bf16x2(rr[i]));
  }
  if (p.act == Y11_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = silu2(f[i]);
  }
  if (p.res && !p.res_pre) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f2_add(f[i], f2_from_bf16x2(rr[i]));
  }
  if (kQ && p.out_esz == 1) {  // e4m3: 16 channels = ONE 16-byte unit of the staging row
    const float o = p.oscale;
    float g[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) f2_unpack(f[i], g[2 * i], g[2 * i + 1]);
    st_shared_v4(dst + (((unit0 >> 1) ^ swz) << 4), pack_e4m3x4(g[0] * o, g[1] * o, g[2] * o, g[3] * o),
                 pack_e4m3x4(g[4] * o, g[5] * o, g[6] * o, g[7] * o), pack_e4m3x4(g[8] * o, g[9] * o, g[10] * o, g[11] * o),
                 pack_e4m3x4(g[12] * o, g[13] * o, g[14] * o, g[15] * o));
  } else if (p.out_f32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a, b, c, d;
      f2_unpack(f[2 * i], a, b);
      f2_unpack(f[2 * i + 1], c, d);
      st_shared_v4(dst + (((2u * unit0 + i) ^ swz) << 4), __float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
    }
  } else {
    st_shared_v4(dst + (((unit0 + 0u) ^ swz) << 4), f2_to_bf16x2(f[0]), f2_to_bf16x2(f[1]), f2_to_bf16x2(f[2]), f2_to_bf16x2(f[3]));
    st_shared_v4(dst + (((unit0 + 1u) ^ swz) << 4), f2_to_bf16x2(f[4]), f2_to_bf16x2(f[5]), f2_to_bf16x2(f[6]), f2_to_bf16x2(f[7]));
  }
}

// ---- class-emit epilogue (ConvTcParams::emit) ----------------------------------------------------------------------------
// The reference takes, per anchor, the maximum of the SIGMOID scores and the first class that attains it.  sigmoid is monotone,
// so the maximum logit gives the score; two different logits share a class rank only when their fp32 sigmoids are equal, which
// needs them closer than 1e-2 (below 8) or both in the saturating range - only then is the accurate sigmoid evaluated here.
__device__ __forceinline__ float emit_sigmoid(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
__device__ __forceinline__ void emit_scan16(const ConvTcParams& p, const uint32_t (&v)[16], int col0, float& best, int& cls) {
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 bb = __ldg(b4 + i);
    const float f[4] = {__uint_as_float(v[4 * i + 0]) + bb.x, __uint_as_float(v[4 * i + 1]) + bb.y,
                        __uint_as_float(v[4 * i + 2]) + bb.z, __uint_as_float(v[4 * i + 3]) + bb.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int col = col0 + 4 * i + k;
      const float x = f[k];
      if (col < p.emit_nc && x > best) {
        // up to 10, logits more than 1e-2 apart have sigmoids >= 7 fp32 ulps apart (sigmoid'(10) = 4.5e-5): only closer pairs
        // and the saturating range need the accurate comparison
        const bool near = x - best <= 1e-2f || best > 10.0f;
        if (!(near && emit_sigmoid(x) == emit_sigmoid(best))) cls = col;  // same sigmoid as an earlier class: that one keeps the rank
        best = x;
      }
    }
  }
}
// called by all 32 lanes of a converged warp; the lanes that list a row of the same image share one atomicAdd
__device__ __forceinline__ void emit_row(const ConvTcParams& p, bool valid, int on, int oh, int ow, float best, int cls) {
  const bool c = valid && best > p.emit_thr;
  const unsigned m = __ballot_sync(0xffffffffu, c);
  if (c) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(m, on);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(p.emit_count + on, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int slot = base + __popc(peers & ((1u << lane) - 1u));
    if (slot < p.emit_cap)
      p.emit_list[(size_t)on * p.emit_cap + slot] = make_int4(p.emit_aoff + oh * p.Wout + ow, cls, __float_as_int(best), 0);
  }
}

// kCpw = accumulator columns an epilogue warp drains per step: 16 (3 CTAs/SM, 64 registers) or 32 ("fat" epilogue: two
// tcgen05.ld in flight, 64-channel store chunks = half the fences / barriers / TMA stores per tile; needs > 64 registers,
// i.e. 2 CTAs/SM).  The ncu source view of the store-heavy 1x1 layers showed the epilogue chain - tcgen05.wait::ld,
// fence.proxy.async, the CTA-wide barrier - as the top stall sites with the issue slots half idle.
template <int kCpw, bool kQ>
__device__ __forceinline__ void conv_tc_body(const ConvTcMaps& maps, const ConvTcParams& p) {
  using namespace y11;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t tiles_base = (smem_base + kHeaderBytes + 1023u) & ~1023u;
  const uint32_t full_bar = smem_base;                    // kMaxStages x 8 B
  const uint32_t empty_bar = smem_base + 8 * kMaxStages;  // kMaxStages x 8 B
  const uint32_t accf_bar = smem_base + 16 * kMaxStages;  // 2 x 8 B  accumulator full  (MMA -> epilogue)
  const uint32_t acce_bar = accf_bar + 16;                // 2 x 8 B  accumulator empty (epilogue -> MMA)
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_raw + 16 * kMaxStages + 48);
  // TMA mode : stages x [A tap box | B tap tile];  halo mode: [resident B, all taps] then stages x [A halo tile]
  //            halo-stream mode: [2 tile buffers x cin/64 swizzled halo chunks] then stages x [B tap tile]
  const uint32_t stage_bytes = p.hstream ? p.b_slot : (p.halo || p.bres) ? p.a_slot : p.a_slot + p.b_slot;
  const uint32_t ring_base = tiles_base + p.b_res_bytes;
  const uint32_t hfull_bar = smem_base + 576, hempty_bar = smem_base + 592;  // halo-stream mode: 2 x halo tile full / empty
  const uint32_t staging_base = ring_base + p.stages * stage_bytes;  // 2 buffers x stg_bytes
  const uint32_t bres_bar = acce_bar + 16;               // halo mode: resident weights landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, (p.halo && !p.halo_tma) ? 32 : 1);   // cp.async halo mode: every producer lane reports its own copies
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(bres_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(hfull_bar + 8 * a, 1); mbar_init(hempty_bar + 8 * a, 1); }
    for (int q = 0; q < kTileQ; ++q) {
      mbar_init(smem_base + kTqFullOff + 8 * q, 1);
      mbar_init(smem_base + kTqEmptyOff + 8 * q, 1 + kEpiWarps);  // consumers: the MMA warp + every epilogue warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf_bar + 8 * a, 1);
      mbar_init(acce_bar + 8 * a, p.epi_warp ? kEpiWarps / 2 : kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.out);
    prefetch_tmap(&maps.outq);
    if (p.stride == 2) {
      prefetch_tmap(&maps.a[1]);
      prefetch_tmap(&maps.a[2]);
      prefetch_tmap(&maps.a[3]);
    }
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), p.tmem_cols);
    tmem_relinquish();
  }
  // LSU mode: per-position table {element offset of the pixel relative to the tile origin, packed (x, y, n)}; positions are
  // the tile's pixels (1x1) or its (Tw+2)x(Th+2) halo (3x3), in the order they are laid out in shared memory
  uint2* pos_tab = reinterpret_cast<uint2*>(smem_raw + 1024);
  if (p.halo) {
    const int HWp = p.Tw + 2 * p.pad, HHp = p.Th + 2 * p.pad;
    for (int pos = threadIdx.x; pos < p.n_pos; pos += kThreads) {
      const int x = pos % HWp, y = (pos / HWp) % HHp, n = pos / (HWp * HHp);
      const int off = ((n * p.Hin + (y - p.pad)) * p.Win + (x - p.pad)) * p.in_ct;
      pos_tab[pos] = make_uint2((uint32_t)off, (uint32_t)(x | (y << 8) | (n << 16)));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t acc_stride = p.tmem_cols >> 1;  // column offset of accumulator stage 1
  // everything above overlapped the tail of the previous kernel (PDL); from here on we touch its output
  pdl_wait();
  pdl_trigger();

  const int k_iters = p.n_kt ? p.n_kt : p.taps * p.chunks_per_tap;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
#ifdef Y11_TRACE
  int trc = 0;
#endif
  const uint32_t tq_full = smem_base + kTqFullOff, tq_empty = smem_base + kTqEmptyOff, tq_tile = smem_base + kTqTileOff;
  // producer side: hand tile `t` (or -1 = no more tiles) to the consumers; qi = running queue index
  auto tq_publish = [&](uint32_t qi, int t) {
    const uint32_t slot = qi & (kTileQ - 1), ph = (qi / kTileQ) & 1u;
    mbar_wait(tq_empty + 8 * slot, ph ^ 1u, p.err_flag, 106);
    if (lane == 0) {
      st_shared_s32(tq_tile + 4 * slot, t);
      mbar_arrive(tq_full + 8 * slot);  // release: the tile index above is visible to whoever observes this phase
    }
    __syncwarp();
  };
  // consumer side (whole warp): next tile of this CTA, -1 when there is none
  auto tq_take = [&](uint32_t qi) -> int {
    const uint32_t slot = qi & (kTileQ - 1), ph = (qi / kTileQ) & 1u;
    mbar_wait(tq_full + 8 * slot, ph, p.err_flag, 107);
    const int t = ld_shared_s32(tq_tile + 4 * slot);
    __syncwarp();
    if (lane == 0) mbar_arrive(tq_empty + 8 * slot);
    return t;
  };
  // producer: index of the tile after the current one (lane 0 asks the global counter; issued early, consumed late)
  auto tq_next = [&](int cur) -> int {
    int nx = 0;
    if (lane == 0) nx = p.tile_counter ? atomicAdd(p.tile_counter, 1) + (int)gridDim.x : cur + (int)gridDim.x;
    return nx;
  };

  if (warp == 0 && p.hstream) {
    // ------------------------------------------------------------------ halo-stream producer (converged warp, one elected lane issues)
    // per tile: the cin/64 swizzled halo chunks of the tile into one of two tile buffers (one barrier), then the 9 x cin/64 weight
    // tiles through the stage ring in the order the MMA warp consumes them (tap, then chunk)
    const int nch = p.chunks_per_tap;
    uint32_t stage = 0, phase = 0, qi = 0;
    for (int tile = blockIdx.x; ; ++qi) {
      tq_publish(qi, tile);
      if (tile < 0) break;
      const int nx_raw = tq_next(tile);
      uint32_t t = tile, q;
      q = fast_div(t, p.mg_ntiles); const int nt = t - q * p.n_tiles; t = q;
      q = fast_div(t, p.mg_tw); const int w0 = (t - q * p.tiles_w) * p.Tw; t = q;
      q = fast_div(t, p.mg_th); const int h0 = (t - q * p.tiles_h) * p.Th; t = q;
      const int n0 = t * p.Tn;
      const uint32_t hs = p.hbufs == 2 ? (qi & 1u) : 0u, hph = p.hbufs == 2 ? ((qi >> 1) & 1u) : (qi & 1u);
      mbar_wait(hempty_bar + 8 * hs, hph ^ 1u, p.err_flag, 108);
      if (lane == 0) TRACE(0, 1);
      if (elect_one()) {
        const uint32_t hb = hfull_bar + 8 * hs;
        mbar_expect_tx(hb, (uint32_t)nch * (uint32_t)p.n_pos * 128u);
        for (int c = 0; c < nch; ++c)
          tma_load_4d(tiles_base + (hs * (uint32_t)nch + (uint32_t)c) * p.a_slot, &maps.a[0], hb, c * 64, w0 - 1, h0 - 1, n0);
      }
      __syncwarp();
      if (lane == 0) TRACE(0, 2);
      for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < nch; ++c) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 101);
          if (elect_one()) {
            const uint32_t fb = full_bar + 8 * stage;
            mbar_expect_tx(fb, p.tx_bytes);
            tma_load_2d(ring_base + stage * stage_bytes, &maps.b, fb, tap * p.cin + c * 64, nt * p.BN);
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      const int nx = __shfl_sync(0xffffffffu, nx_raw, 0);
      tile = nx < total_tiles ? nx : -1;
    }
  } else if (warp == 0 && p.halo) {
    // ------------------------------------------------------------------ LSU producer (whole warp, cp.async)
    if (lane == 0) {  // weights of all taps: loaded once, stay resident
      const int nb = p.taps * p.chunks_per_tap;
      mbar_expect_tx(bres_bar, (uint32_t)(nb * p.BN) * (uint32_t)(p.Cc * 2));
      for (int i = 0; i < nb; ++i) tma_load_2d(tiles_base + i * p.b_slot, &maps.b, bres_bar, i * p.Cc, 0);
    }
    const int ncg = p.cin >> 3;
    const uint32_t plane = p.a_lbo;  // bytes per 8-channel plane
    const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(p.in);
    // Lane l owns positions l, l+32, ... (<= 6 per tile): their source offsets relative to the tile origin and their
    // packed (x, y, n) are per-lane constants in registers.  Consecutive lanes write consecutive 16-byte slots of a plane
    // (conflict-free shared-memory writes - a lane-per-channel-group mapping that coalesces the global side instead
    // measured 35 % slower) and each lane copies the cin/8 channel groups of its pixel with immediate offsets.
    constexpr int kMaxPos = 6;
    int soff[kMaxPos];
    uint32_t meta[kMaxPos];
#pragma unroll
    for (int i = 0; i < kMaxPos; ++i) {
      const int pos = lane + 32 * i;
      const uint2 e = pos < p.n_pos ? pos_tab[pos] : make_uint2(0u, 0xffffffffu);
      soff[i] = (int)e.x;
      meta[i] = e.y;
    }
    uint32_t stage = 0, phase = 0, qi = 0;
    for (int tile = blockIdx.x; ; ++qi) {
      tq_publish(qi, tile);
      if (tile < 0) break;
      const int nx_raw = tq_next(tile);
      uint32_t t = tile, q;
      q = fast_div(t, p.mg_tw); const int w0 = (t - q * p.tiles_w) * p.Tw; t = q;
      q = fast_div(t, p.mg_th); const int h0 = (t - q * p.tiles_h) * p.Th; t = q;
      const int n0 = t * p.Tn;
      const __nv_bfloat16* org = in + ((size_t)n0 * p.Hin * p.Win + (size_t)h0 * p.Win + w0) * p.in_ct;
      mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 101);
      if (lane == 0) TRACE(0, 1);
      if (p.halo_tma) {
        // the whole (Tw+2) x (Th+2) x 64-channel halo as ONE TMA box: 128-byte rows in the swizzled layout, zero fill outside the
        // image - 180 rows per tile instead of 9 x 128 (tap-by-tap TMA mode) or 1440 16-byte cp.async copies
        if (elect_one()) {
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx(fb, (uint32_t)p.n_pos * 128u);
          tma_load_4d(ring_base + stage * stage_bytes, &maps.a[0], fb, 0, w0 - 1, h0 - 1, n0);
        }
        __syncwarp();
        if (lane == 0) TRACE(0, 2);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        const int nx = __shfl_sync(0xffffffffu, nx_raw, 0);
        tile = nx < total_tiles ? nx : -1;
        continue;
      }
      const uint32_t dst = ring_base + stage * stage_bytes + (uint32_t)lane * 16u;
#pragma unroll
      for (int i = 0; i < kMaxPos; ++i) {
        if (meta[i] == 0xffffffffu) continue;  // beyond the tile's positions
        const int iw = w0 + (int)(meta[i] & 0xff) - p.pad, ih = h0 + (int)((meta[i] >> 8) & 0xff) - p.pad;
        const int in_ = n0 + (int)(meta[i] >> 16);
        const bool ok = ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win && in_ < p.B;  // outside: zero fill == conv padding
        const __nv_bfloat16* src = org + (ok ? soff[i] : 0);
        const uint32_t d = dst + (uint32_t)i * 512u, nbytes = ok ? 16u : 0u;
#pragma unroll 4
        for (int cg = 0; cg < ncg; ++cg) cp_async_16(d + cg * plane, src + cg * 8, nbytes);
      }
      cp_async_mbar_arrive(full_bar + 8 * stage);
      if (lane == 0) TRACE(0, 2);
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      const int nx = __shfl_sync(0xffffffffu, nx_raw, 0);
      tile = nx < total_tiles ? nx : -1;
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, one elected lane issues)
    if (p.bres && lane == 0) {  // resident weights: every K stage of the single N tile, once per CTA
      const int nb = p.n_kt ? p.n_kt : p.taps * p.chunks_per_tap;
      mbar_expect_tx(bres_bar, (uint32_t)(nb * p.BN) * (uint32_t)(p.Cc * 2));
      for (int i = 0; i < nb; ++i) tma_load_2d(tiles_base + i * p.b_slot, &maps.b, bres_bar, i * p.Cc, 0);
    }
    uint32_t stage = 0, phase = 0, qi = 0;
    for (int tile = blockIdx.x; ; ++qi) {
      tq_publish(qi, tile);
      if (tile < 0) break;
      const int nx_raw = tq_next(tile);
      uint32_t t = tile, q;
      q = fast_div(t, p.mg_ntiles); const int nt = t - q * p.n_tiles; t = q;
      q = fast_div(t, p.mg_tw); const int w0 = (t - q * p.tiles_w) * p.Tw; t = q;
      q = fast_div(t, p.mg_th); const int h0 = (t - q * p.tiles_h) * p.Th; t = q;
      const int n0 = t * p.Tn;
      for (int k = 0; k < p.n_kt; ++k) {  // compact k = 2 form: only the channel blocks a block tap can touch
        const uint32_t e = p.kt[k];
        const int tap = (int)(e & 3u), c0 = (int)(e >> 2) * 16;
        mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 101);
        if (elect_one()) {
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t a_dst = ring_base + stage * stage_bytes;
          mbar_expect_tx(fb, p.tx_bytes);
          tma_load_4d(a_dst, &maps.a[0], fb, c0, w0 + (tap & 1) - 1, h0 + (tap >> 1) - 1, n0);
          if (!p.bres) tma_load_2d(a_dst + p.a_slot, &maps.b, fb, k * p.Cc, nt * p.BN);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      }
      for (int tap = 0; tap < (p.n_kt ? 0 : p.taps); ++tap) {
        int mi = 0, cw, ch;
        if (p.ksize == 1) {
          cw = w0; ch = h0;
        } else if (p.ksize == 2) {  // taps at {-1, 0} x {-1, 0}: a 3x3 stride-2 conv on a space-to-depth input
          cw = w0 + tap % 2 - 1; ch = h0 + tap / 2 - 1;
        } else if (p.stride == 1) {
          cw = w0 + tap % 3 - 1; ch = h0 + tap / 3 - 1;
        } else {
          const int kh = tap / 3, kw = tap % 3;
          mi = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1);  // input row 2*oy+kh-1 has parity (kh != 1)
          cw = w0 - (kw == 0); ch = h0 - (kh == 0);
        }
        for (int c = 0; c < p.chunks_per_tap; ++c) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 101);
          if (lane == 0) TRACE(0, 1);
          if (elect_one()) {
            const uint32_t fb = full_bar + 8 * stage;
            const uint32_t a_dst = ring_base + stage * stage_bytes;
            mbar_expect_tx(fb, p.tx_bytes);
            tma_load_4d(a_dst, &maps.a[mi], fb, c * p.Cc, cw, ch, n0);
            if (!p.bres) tma_load_2d(a_dst + p.a_slot, &maps.b, fb, tap * p.cin + c * p.Cc, nt * p.BN);
          }
          __syncwarp();
          if (lane == 0) TRACE(0, 2);
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
      const int nx = __shfl_sync(0xffffffffu, nx_raw, 0);
      tile = nx < total_tiles ? nx : -1;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop converged and ONE elected lane issues tcgen05.mma / tcgen05.commit.  (The first
    // version ran the loop inside `if (lane == 0)`: in a divergent region the compiler cannot prove operands warp-uniform
    // and wrapped every UTCHMMA in an ELECT/R2UR/BRA.U.ANY sequence - ~50 instructions and ~500 cycles per MMA, which
    // made the issue thread the bottleneck of every short-K layer: ncu source view, profiles/r01c_summary.md.)
    // Round 2 (timeline of this warp, tools/trace_narrow.py): the generic loops - runtime K-step counts, descriptors rebuilt
    // from constant-bank parameters per stage, the kmask / e4m3 branches inside the K loop - still cost 150-250 cycles per
    // UMMA (2 UMMAs + commit of a 32-channel stage: 500 cycles; the 18 UMMAs of a 3x3 32->16 halo tile: 3300), i.e. the issue
    // thread was slower than the tensor pipe (64 cycles for a 128x128x16 UMMA).  The loops are now specialised on the number
    // of UMMAs per stage (compile-time unrolled, immediate descriptor offsets) and carry the 32-bit descriptor LOW words as
    // running warp-uniform values; the HIGH words are per-kernel constants.  Same issue order -> bit-identical results.
    const bool f8 = kQ && p.in_fp8;
    const uint32_t idesc = f8 ? make_idesc_e4m3_m128(p.BN) : make_idesc_bf16_m128(p.BN);
    const int kk_n = f8 ? p.Cc / 32 : p.Cc / 16;  // one MMA = 32 bytes of K: 16 bf16 or 32 e4m3 elements
    uint32_t stage = 0, phase = 0;  // ring position, advanced incrementally (no div/mod per stage)
    int ti = 0;
    if (p.halo || p.bres) mbar_wait(bres_bar, 0, p.err_flag, 105);
    const uint32_t n_stages = (uint32_t)p.stages;
    const uint32_t stage_units = stage_bytes >> 4;
    // swizzled K-major descriptor (weights everywhere, activations in TMA mode): lo = addr>>4 | LBO(1)<<16, hi = SBO>>4 | version | layout
    const uint32_t sw_hi = (p.sbo >> 4) | (1u << 14) | (p.layout_type << 29);
    const uint32_t ring_lo0 = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
    if (p.hstream) {
      // halo-stream: A = the tile's swizzled halo chunks (tap = row offset of the start address, see halo3_issue_sw), B = one
      // streamed weight tile per (tap, chunk) stage; K order (tap, chunk, 16-channel step) as in the tap-by-tap TMA mode
      const uint32_t a_hi = (p.a_sbo >> 4) | (1u << 14) | (2u << 29);
      const uint32_t halo_lo0 = ((tiles_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_units = p.a_slot >> 4;
      const int nch = p.chunks_per_tap;
      for (;; ++ti) {
        if (tq_take((uint32_t)ti) < 0) break;
        const int as = ti & 1;
        mbar_wait(acce_bar + 8 * as, ((ti >> 1) & 1) ^ 1, p.err_flag, 104);  // epilogue has drained this accumulator stage
        if (lane == 0) TRACE(1, 1);
        const uint32_t tmem_acc = tmem_base + as * acc_stride;
        if (elect_one()) {
          const uint32_t hs = p.hbufs == 2 ? ((uint32_t)ti & 1u) : 0u, hph = p.hbufs == 2 ? (((uint32_t)ti >> 1) & 1u) : ((uint32_t)ti & 1u);
          mbar_wait(hfull_bar + 8 * hs, hph, p.err_flag, 109);
          tc_fence_after();
          const uint32_t h_lo = halo_lo0 + hs * (uint32_t)nch * a_units;
          uint32_t st = stage, ph = phase;
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t toff = (uint32_t)((tap / 3) * 10 + tap % 3) * 8u;
            for (int c = 0; c < nch; ++c) {
              mbar_wait(full_bar + 8 * st, ph, p.err_flag, 102);
              tc_fence_after();
              const uint32_t at = h_lo + (uint32_t)c * a_units + toff;
              const uint32_t b_lo = ring_lo0 + st * stage_units;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_lohi<false>(tmem_acc, at + 2u * kk, a_hi, b_lo + 2u * kk, sw_hi, idesc, (tap | c | kk) ? 1u : 0u);
              umma_commit(empty_bar + 8 * st);  // frees the weight slot once these MMAs retire
              if (++st == n_stages) { st = 0; ph ^= 1; }
            }
          }
          umma_commit(hempty_bar + 8 * hs);   // the halo tile buffer may be refilled
          umma_commit(accf_bar + 8 * as);     // accumulator complete
        }
        __syncwarp();
        if (lane == 0) TRACE(1, 3);
        uint32_t adv = stage + (uint32_t)k_iters;
        while (adv >= n_stages) { adv -= n_stages; phase ^= 1; }
        stage = adv;
      }
    } else if (p.halo) {
      // un-swizzled activation tile: lo = addr>>4 | (LBO>>4)<<16, hi = SBO>>4 | version
      const uint32_t a_hi = (p.a_sbo >> 4) | (1u << 14);
      const uint32_t a_lo0 = ((ring_base & 0x3FFFFu) >> 4) | ((p.a_lbo >> 4) << 16);
      const uint32_t b_lo0 = ((tiles_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_step = p.b_slot >> 4, k_step = (2u * p.a_lbo) >> 4;
      const int cpt = p.chunks_per_tap;
      const int mode = (p.pad && cpt == 1 && !f8) ? kk_n : 0;  // 1 / 2 / 4: specialised 3x3 loop (cin = 16 / 32 / 64)
      uint32_t a_lo = a_lo0, fb = full_bar, eb = empty_bar;
      for (;; ++ti) {
        if (tq_take((uint32_t)ti) < 0) break;
        const int as = ti & 1;
        mbar_wait(acce_bar + 8 * as, ((ti >> 1) & 1) ^ 1, p.err_flag, 104);  // epilogue has drained this accumulator stage
        if (lane == 0) TRACE(1, 1);
        const uint32_t tmem_acc = tmem_base + as * acc_stride;
        mbar_wait(fb, phase, p.err_flag, 102);
        if (lane == 0) TRACE(1, 2);
        fence_async_smem();  // cp.async wrote through the generic proxy; the MMA reads through the async proxy
        tc_fence_after();
        if (elect_one()) {
          // 3x3: tap (kh, kw) = the halo tile shifted by kh rows of (Tw+2) = 10 pixels and kw pixels, in 16-byte units; the next
          // 16 channels = two 8-channel planes further (k_step) / +32 B in the weight row
          if (p.halo_tma) halo3_issue_sw<4>(tmem_acc, (a_lo & 0x3FFFu) | (1u << 16), (p.a_sbo >> 4) | (1u << 14) | (2u << 29), b_lo0, sw_hi, b_step, idesc);
          else if (mode == 2) halo3_issue<2>(tmem_acc, a_lo, a_hi, b_lo0, sw_hi, b_step, k_step, idesc);
          else if (mode == 4) halo3_issue<4>(tmem_acc, a_lo, a_hi, b_lo0, sw_hi, b_step, k_step, idesc);
          else if (mode == 1) halo3_issue<1>(tmem_acc, a_lo, a_hi, b_lo0, sw_hi, b_step, k_step, idesc);
          else if (p.pad) {
            // cin = 48 (the 1.5x-wide scale): three 16-channel weight chunks per tap, each with its own resident weight slot
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t at = a_lo + (uint32_t)((tap / 3) * 10 + tap % 3);
              for (int c = 0; c < cpt; ++c)
                for (int kk = 0; kk < kk_n; ++kk)
                  umma_lohi<false>(tmem_acc, at + (uint32_t)(c * kk_n + kk) * k_step, a_hi,
                                   b_lo0 + (uint32_t)(tap * cpt + c) * b_step + 2u * kk, sw_hi, idesc, (tap | c | kk) != 0);
            }
          } else {
            // 1x1: K = cin in weight chunks of Cc channels
            for (int c = 0; c < cpt; ++c)
              for (int kk = 0; kk < kk_n; ++kk)
                umma_lohi<false>(tmem_acc, a_lo + (uint32_t)(c * kk_n + kk) * k_step, a_hi, b_lo0 + (uint32_t)c * b_step + 2u * kk,
                                 sw_hi, idesc, (c | kk) != 0);
          }
          umma_commit(eb);
          umma_commit(accf_bar + 8 * as);
        }
        __syncwarp();
        if (lane == 0) TRACE(1, 3);
        if (++stage == n_stages) { stage = 0; phase ^= 1; a_lo = a_lo0; fb = full_bar; eb = empty_bar; }
        else { a_lo += stage_units; fb += 8; eb += 8; }
      }
    } else {
      const uint32_t b_off = p.a_slot >> 4;  // weights of a stage follow its activation box
      const int mode = p.kmask_on ? 0 : kk_n;
      for (;; ++ti) {
        if (tq_take((uint32_t)ti) < 0) break;
        const int as = ti & 1;
        mbar_wait(acce_bar + 8 * as, ((ti >> 1) & 1) ^ 1, p.err_flag, 104);  // epilogue has drained this accumulator stage
        if (lane == 0) TRACE(1, 1);
        const uint32_t tmem_acc = tmem_base + as * acc_stride;
        const uint32_t accf = accf_bar + 8 * as;
        // ONE elected lane walks all K stages of the tile (wait, UMMAs, commit): no per-stage elect / warp re-convergence
        // between the last UMMA of a stage and the first of the next - with BN = 256 the pipe's queue does not cover that gap
        if (elect_one()) {
          const uint32_t bres_lo = p.bres ? (((tiles_base & 0x3FFFFu) >> 4) | (1u << 16)) : 0u;
          const TmaTileArgs ta{tmem_acc, ring_lo0, b_off, sw_hi, idesc, stage_units, n_stages, full_bar, empty_bar, accf,
                               k_iters, p.err_flag, bres_lo, p.b_slot >> 4};
          if (mode == 4) tma_tile_issue<4, kQ>(ta, stage, phase, f8);
          else if (mode == 2) tma_tile_issue<2, kQ>(ta, stage, phase, f8);
          else if (mode == 1) tma_tile_issue<1, kQ>(ta, stage, phase, f8);
          else if (mode == 8) tma_tile_issue<8, kQ>(ta, stage, phase, f8);
          else {
            // structurally sparse K (the 2x2 space-to-depth form of a 3x3 stride-2 conv has 7 of 16 all-zero
            // 16-channel blocks): issue only the K steps whose weights are not all zero
            uint32_t st = stage, ph = phase, started = 0u;
            for (int k = 0; k < k_iters; ++k) {
              mbar_wait(full_bar + 8 * st, ph, p.err_flag, 102);
              tc_fence_after();
              const uint32_t a_lo = ring_lo0 + st * stage_units;
              const uint32_t b_lo = bres_lo ? bres_lo + (uint32_t)k * (p.b_slot >> 4) : a_lo + b_off;
              const uint32_t bits = (uint32_t)(p.kmask >> (k * kk_n));
              for (int kk = 0; kk < kk_n; ++kk) {
                if ((bits >> kk) & 1u) {
                  umma_lohi<false>(tmem_acc, a_lo + 2u * kk, sw_hi, b_lo + 2u * kk, sw_hi, idesc, started);
                  started = 1u;
                }
              }
              umma_commit(empty_bar + 8 * st);  // frees the smem slot once these MMAs retire
              if (++st == n_stages) { st = 0; ph ^= 1; }
            }
            umma_commit(accf);  // accumulator complete
          }
        }
        __syncwarp();
        if (lane == 0) TRACE(1, 3);
        // ring position after this tile, the same in every lane
        uint32_t adv = stage + (uint32_t)k_iters;
        while (adv >= n_stages) { adv -= n_stages; phase ^= 1; }
        stage = adv;
      }
    }
  } else if (p.epi_warp) {
    // -------------------------------------------------------------------- epilogue, warp-independent (warps 2..9)
    // Two groups of four warps take ALTERNATE tiles (group g owns accumulator stage g); inside a group every warp drains
    // its own TMEM lane quadrant = 32 tile rows = a (qw x qh x qn) sub-box of the tile, stages it in a private swizzled
    // buffer and issues its own TMA store.  No CTA-wide barrier, no shared leader: the timeline of the CTA-wide version
    // (tools/trace_conv.py) showed ~3500 cycles per 128x32 tile spent in one serial chain - wait accumulator -> tcgen05.ld
    // -> math -> st.shared -> leader waits for the previous store -> named barrier (7 warps idle) -> leader issues the
    // store -> release - with the MMA warp and the TMA producer both waiting on it.  Here eight such chains run
    // independently and the accumulator stage is handed back as soon as its last tcgen05.ld has completed.
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const uint32_t esz = p.out_f32 ? 4u : 2u;
    const uint32_t pitch = (uint32_t)p.cw * esz;           // staging row pitch: 32 / 64 / 128 B
    const uint32_t swz = ((uint32_t)lane / (128u / pitch)) & (pitch / 16u - 1u);  // TMA SWIZZLE_{32,64,128}B phase of row `lane`
    const uint32_t wbuf_bytes = 32u * pitch;
    const uint32_t my_stage = staging_base + (uint32_t)ew * (uint32_t)p.nstg * wbuf_bytes;
    const int n_chunks = (p.BN + p.cw - 1) / p.cw;
    const int steps = p.cw / kCpw;  // kCpw-column steps per chunk
    // position of this warp's 32 rows inside the tile, and of this lane's row (for the residual read)
    const int r0 = quad * 32, r = r0 + lane;
    const int qw0 = r0 % p.Tw, qh0 = (r0 / p.Tw) % p.Th, qn0 = r0 / (p.Tw * p.Th);
    const int tw = r % p.Tw, th = (r / p.Tw) % p.Th, tn = r / (p.Tw * p.Th);
    uint32_t sb = 0;
    int ti = 0;
    for (;; ++ti) {
      const int tile = tq_take((uint32_t)ti);
      if (tile < 0) break;
      if ((ti & 1) != grp) continue;
      uint32_t t = tile, qq;
      qq = fast_div(t, p.mg_ntiles); const int nt = t - qq * p.n_tiles; t = qq;
      qq = fast_div(t, p.mg_tw); const int w0 = (t - qq * p.tiles_w) * p.Tw; t = qq;
      qq = fast_div(t, p.mg_th); const int h0 = (t - qq * p.tiles_h) * p.Th; t = qq;
      const int n0 = t * p.Tn;
      const int ow = w0 + tw, oh = h0 + th, on = n0 + tn;
      const bool valid = ow < p.Wout && oh < p.Hout && on < p.B;
      const size_t pix = (static_cast<size_t>(on) * p.Hout + oh) * p.Wout + ow;
      // residual row: the output pixel itself, or (pre-activation, nearest-upsampled term) pixel (oh/2, ow/2) of a half-size map
      const size_t rpix = p.res_pre ? (static_cast<size_t>(on) * (p.Hout >> 1) + (oh >> 1)) * (p.Wout >> 1) + (ow >> 1) : pix;
      const __nv_bfloat16* res_row = static_cast<const __nv_bfloat16*>(p.res) + rpix * p.res_ct + p.res_co + nt * p.BN;
      mbar_wait(accf_bar + 8 * grp, (ti >> 1) & 1, p.err_flag, 103);
      tc_fence_after();
      const uint32_t taddr = tmem_base + grp * acc_stride + (static_cast<uint32_t>(quad * 32) << 16);
      if (p.emit) {  // class-emit mode: row maximum + class instead of a store
        float best = -INFINITY;
        int cls = 0;
        for (int col = 0; col < p.BN; col += 16) {
          uint32_t va[16];
          tmem_ld16(taddr + col, va);
          tmem_ld_wait();
          emit_scan16(p, va, col, best, cls);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar + 8 * grp);
        emit_row(p, valid, on, oh, ow, best, cls);
        continue;
      }
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t buf = my_stage + sb * wbuf_bytes;
        if (lane == 0) {  // the last store that used this buffer (two chunks ago / the previous one) has finished reading it
          if (p.nstg >= 2) bulk_wait_read1();
          else bulk_wait_read0();
        }
        __syncwarp();
        const uint32_t dst = buf + (uint32_t)lane * pitch;
        for (int st = 0; st < steps; ++st) {
          const int col = c * p.cw + st * kCpw;  // column inside the tile's accumulator
          if (col >= p.BN) break;
          const bool two = kCpw == 32 && col + 16 < p.BN;
          uint4 ra0 = make_uint4(0, 0, 0, 0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
          if (p.res && valid) {  // issued before the TMEM load: independent of it
            ra0 = *reinterpret_cast<const uint4*>(res_row + col);
            ra1 = *reinterpret_cast<const uint4*>(res_row + col + 8);
            if (two) {
              rb0 = *reinterpret_cast<const uint4*>(res_row + col + 16);
              rb1 = *reinterpret_cast<const uint4*>(res_row + col + 24);
            }
          }
          uint32_t va[16], vb[16];
          tmem_ld16(taddr + col, va);
          if (two) tmem_ld16(taddr + col + 16, vb);
          tmem_ld_wait();
          if (c == n_chunks - 1 && col + kCpw >= p.BN) {
            // last TMEM read of this accumulator stage: hand it back to the MMA warp before doing the math
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acce_bar + 8 * grp);
          }
          const int n = nt * p.BN + col;
          epi16<kQ>(p, va, ra0, ra1, n, dst, (uint32_t)(st * kCpw) / 8u, swz);
          if (two) epi16<kQ>(p, vb, rb0, rb1, n + 16, dst, (uint32_t)(st * kCpw + 16) / 8u, swz);
        }
        fence_async_smem();  // my generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          // columns beyond cout (only when cout % cw != 0, single N tile) and pixels beyond the image are clipped by the map
          tma_store_4d(&maps.outq, buf, nt * p.BN + c * p.cw, w0 + qw0, h0 + qh0, n0 + qn0);
          bulk_commit();
        }
        if (p.nstg >= 2) sb ^= 1u;
      }
    }
    if (lane == 0) bulk_wait_all();
  } else {
    // -------------------------------------------------------------------- epilogue (warps 2..9)
    // All 8 warps work on the same chunk of CW output channels: warp quadrant q owns TMEM lanes / tile rows
    // [32q, 32q+32), `half` picks the kCpw-column slice inside the chunk.  One staging buffer pair, one named barrier and
    // one TMA store per chunk (v2 used two independent 4-warp groups with 16-column chunks: twice the barriers/stores).
    const int ew = warp - 2;
    const int half = ew >> 2;  // 0/1: which kCpw columns of a 2*kCpw-column chunk
    const int quad = warp & 3; // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int r = quad * 32 + lane;
    const int tw = r % p.Tw, th = (r / p.Tw) % p.Th, tn = r / (p.Tw * p.Th);
    const bool leader = ew == 0 && lane == 0;
    const uint32_t esz = kQ ? (uint32_t)p.out_esz : (p.out_f32 ? 4u : 2u);
    const int halves = p.cw / kCpw;                        // 1 or 2 warps per row share a chunk
    const uint32_t pitch = (uint32_t)p.cw * esz;           // staging row pitch: 32 / 64 / 128 B
    const uint32_t swz = (r / (128u / pitch)) & (pitch / 16u - 1u);  // TMA SWIZZLE_{32,64,128}B pattern for row r
    const uint32_t stg_bytes = 128u * pitch;
    const uint32_t row_addr = r * pitch;
    const int n_chunks = (p.BN + p.cw - 1) / p.cw;
    const bool active = half < halves;
    int ti = 0;
    uint32_t sb = 0;  // staging ring position
    for (;; ++ti) {
      const int tile = tq_take((uint32_t)ti);
      if (tile < 0) break;
      uint32_t t = tile, qq;
      qq = fast_div(t, p.mg_ntiles); const int nt = t - qq * p.n_tiles; t = qq;
      qq = fast_div(t, p.mg_tw); const int w0 = (t - qq * p.tiles_w) * p.Tw; t = qq;
      qq = fast_div(t, p.mg_th); const int h0 = (t - qq * p.tiles_h) * p.Th; t = qq;
      const int n0 = t * p.Tn;
      const int ow = w0 + tw, oh = h0 + th, on = n0 + tn;
      const bool valid = (r < p.Tw * p.Th * p.Tn) && ow < p.Wout && oh < p.Hout && on < p.B;
      const size_t pix = (static_cast<size_t>(on) * p.Hout + oh) * p.Wout + ow;
      // residual pixel: the output pixel itself, or (pre-activation, nearest-upsampled term) pixel (oh/2, ow/2) of a half-size map
      const size_t rpix = p.res_pre ? (static_cast<size_t>(on) * (p.Hout >> 1) + (oh >> 1)) * (p.Wout >> 1) + (ow >> 1) : pix;
      const __nv_bfloat16* res_row = static_cast<const __nv_bfloat16*>(p.res) + rpix * p.res_ct + p.res_co + nt * p.BN;
      const int as = ti & 1;
      if (lane == 0 && (ew == 0 || ew == 3)) TRACE(2 + (ew == 3), 0);
      mbar_wait(accf_bar + 8 * as, (ti >> 1) & 1, p.err_flag, 103);
      if (lane == 0 && (ew == 0 || ew == 3)) TRACE(2 + (ew == 3), 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * acc_stride + (static_cast<uint32_t>(quad * 32) << 16);
      if (p.emit) {  // class-emit mode: the first warp of each lane quadrant scans all columns of its 32 rows; nothing is stored
        if (half == 0) {
          float best = -INFINITY;
          int cls = 0;
          for (int col = 0; col < p.BN; col += 16) {
            uint32_t va[16];
            tmem_ld16(taddr + col, va);
            tmem_ld_wait();
            emit_scan16(p, va, col, best, cls);
          }
          emit_row(p, valid, on, oh, ow, best, cls);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar + 8 * as);
        continue;
      }
      for (int c = 0; c < n_chunks; ++c) {
        const int col = c * p.cw + half * kCpw;          // column inside the tile's accumulator
        const uint32_t dst = staging_base + sb * stg_bytes + row_addr;
        if (active && col < p.BN) {
          const bool two = kCpw == 32 && col + 16 < p.BN;
          uint4 ra0 = make_uint4(0, 0, 0, 0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
          if (p.res && valid) {  // issued before the TMEM load: independent of it, and a DRAM/L2 round trip long
            ra0 = *reinterpret_cast<const uint4*>(res_row + col);
            ra1 = *reinterpret_cast<const uint4*>(res_row + col + 8);
            if (two) {
              rb0 = *reinterpret_cast<const uint4*>(res_row + col + 16);
              rb1 = *reinterpret_cast<const uint4*>(res_row + col + 24);
            }
          }
          uint32_t va[16], vb[16];
          tmem_ld16(taddr + col, va);
          if (two) tmem_ld16(taddr + col + 16, vb);
          tmem_ld_wait();
          if (lane == 0 && (ew == 0 || ew == 3)) TRACE(2 + (ew == 3), 2);
          const int n = nt * p.BN + col;
          epi16<kQ>(p, va, ra0, ra1, n, dst, (uint32_t)(half * kCpw) / 8u, swz);
          if (two) epi16<kQ>(p, vb, rb0, rb1, n + 16, dst, (uint32_t)(half * kCpw + 16) / 8u, swz);
          fence_async_smem();  // my generic-proxy smem writes -> visible to the TMA (async proxy)
          if (lane == 0 && (ew == 0 || ew == 3)) TRACE(2 + (ew == 3), 3);
        }
        if (leader) bulk_wait_read0();   // the previous store has finished READING its (other) staging buffer
        named_bar_sync(1, kEpiWarps * 32);  // all rows/columns of this chunk staged; other buffer free for the next chunk
        if (leader) {
          // columns beyond cout (only when cout % cw != 0, single N tile) are clipped by the tensor map
          tma_store_4d(&maps.out, staging_base + sb * stg_bytes, nt * p.BN + c * p.cw, w0, h0, n0);
          bulk_commit();
          TRACE(2, 6);
        }
        sb ^= 1u;
      }
      // all tcgen05.ld of this accumulator stage have completed (wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar + 8 * as);
    }
    if (leader) bulk_wait_all();
  }

  if (threadIdx.x == 0 && p.tile_counter) {
    // thread 0 is the producer lane that made this CTA's last request; the last CTA to get here re-arms the counters for the
    // next launch of this op (visible to it through the kernel boundary / griddepcontrol.wait)
    if (atomicAdd(p.tile_counter + 1, 1) == (int)gridDim.x - 1) {
      p.tile_counter[0] = 0;
      p.tile_counter[1] = 0;
      __threadfence();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ================================================================================================================
// 2-CTA variant (round 2): tcgen05.mma.cta_group::2 - a CTA PAIR (cluster of two, one CTA per SM of a TPC) computes a 256-row
// x BN tile.  Each CTA loads its own 128-row activation box and HALF of the weight tile (BN/2 rows), the leader CTA's MMA warp
// issues one UMMA for the pair (M = 256; D rows 0-127 land in the leader's TMEM, 128-255 in the peer's), each CTA drains and
// stores its own rows.  Why: the 1-CTA tiles are bound by SHARED-MEMORY bandwidth, not by the tensor pipe - per K stage of a
// 128 x 256 tile the TMA writes 48 KB and the UMMAs read 48 KB, 96 KB / 128 B/clk = 750 cycles against 512 cycles of math
// (timeline: 676 cycles per stage, 61-72 % tensor-pipe utilisation in ncu).  With the weight tile split across the pair each
// SM writes + reads 2 x 32 KB per stage (500 cycles), and 5 stages of 32 KB fit instead of 3 of 48 KB.
// Scope: TMA producer, bf16, k in {1, 3}, stride 1 / 2, CTA-wide epilogue.  Same per-element K order -> bit-identical results.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_bf16_m256(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

template <int KK>
__device__ __forceinline__ void pair_issue(uint32_t tmem_acc, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) umma2_lohi(tmem_acc, a_lo + 2u * kk, hi, b_lo + 2u * kk, hi, idesc, kk ? 1u : acc0);
}

template <int kCpw>  // accumulator columns an epilogue warp drains per step (32 = fat epilogue, 64-channel store chunks)
__device__ __forceinline__ void conv_tc_pair_body(const ConvTcMaps& maps, const ConvTcParams& p) {
  using namespace y11;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t tiles_base = (smem_base + kHeaderBytes + 1023u) & ~1023u;
  const uint32_t full_bar = smem_base;                    // kMaxStages x 8 B (used in the leader CTA)
  const uint32_t empty_bar = smem_base + 8 * kMaxStages;  // kMaxStages x 8 B (one per CTA)
  const uint32_t accf_bar = smem_base + 16 * kMaxStages;  // 2 x 8 B  accumulator full  (MMA -> epilogue, one per CTA)
  const uint32_t acce_bar = accf_bar + 16;                // 2 x 8 B  accumulator empty (both epilogues -> leader's MMA warp)
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_raw + 16 * kMaxStages + 48);
  const uint32_t stage_bytes = p.a_slot + p.b_slot;
  const uint32_t ring_base = tiles_base;
  const uint32_t staging_base = ring_base + p.stages * stage_bytes;  // 2 buffers x stg_bytes
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs, owns the full / accumulator-empty barriers)

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);   // the leader producer's arrive.expect_tx; both CTAs' TMA loads complete bytes on it
      mbar_init(empty_bar + 8 * s, 1);  // the pair's tcgen05.commit, multicast to both CTAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf_bar + 8 * a, 1);
      mbar_init(acce_bar + 8 * a, 2 * kEpiWarps);  // the epilogue warps of BOTH CTAs
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.out);
    if (p.stride == 2) {
      prefetch_tmap(&maps.a[1]);
      prefetch_tmap(&maps.a[2]);
      prefetch_tmap(&maps.a[3]);
    }
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything of this CTA can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t acc_stride = p.tmem_cols >> 1;
  pdl_wait();
  pdl_trigger();

  const int k_iters = p.taps * p.chunks_per_tap;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_pairs = ((m_tiles + 1) >> 1) * p.n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // pair tile -> (N tile, this CTA's M tile); an odd last M tile leaves the peer with an all-out-of-bounds tile (n0 >= B: the
  // TMA loads zero-fill, the store is clipped, the residual reads are guarded)
  auto decode = [&](int pt, int& nt, int& w0, int& h0, int& n0) {
    uint32_t t = (uint32_t)pt, q;
    q = fast_div(t, p.mg_ntiles); nt = (int)(t - q * p.n_tiles); t = 2u * q + rank;
    if ((int)t >= m_tiles) { w0 = 0; h0 = 0; n0 = p.tiles_n * p.Tn; return; }
    q = fast_div(t, p.mg_tw); w0 = (int)(t - q * p.tiles_w) * p.Tw; t = q;
    q = fast_div(t, p.mg_th); h0 = (int)(t - q * p.tiles_h) * p.Th; t = q;
    n0 = (int)t * p.Tn;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t full0 = mapa_u32(full_bar, 0);  // the leader's full barriers, as a cluster address
    uint32_t stage = 0, phase = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
      int nt, w0, h0, n0;
      decode(pt, nt, w0, h0, n0);
      for (int tap = 0; tap < p.taps; ++tap) {
        int mi = 0, cw, ch;
        if (p.ksize == 1) {
          cw = w0; ch = h0;
        } else if (p.stride == 1) {
          cw = w0 + tap % 3 - 1; ch = h0 + tap / 3 - 1;
        } else {
          const int kh = tap / 3, kw = tap % 3;
          mi = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1);
          cw = w0 - (kw == 0); ch = h0 - (kh == 0);
        }
        for (int c = 0; c < p.chunks_per_tap; ++c) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 111);
          if (elect_one()) {
            const uint32_t a_dst = ring_base + stage * stage_bytes;
            if (rank == 0) mbar_expect_tx(full_bar + 8 * stage, 2u * p.tx_bytes);  // this CTA's and the peer's bytes
            tma_load_4d_2sm(a_dst, &maps.a[mi], full0 + 8 * stage, c * p.Cc, cw, ch, n0);
            tma_load_2d_2sm(a_dst + p.a_slot, &maps.b, full0 + 8 * stage, tap * p.cin + c * p.Cc, nt * p.BN + (int)rank * (p.BN >> 1));
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16_m256(p.BN);
      const int kk_n = p.Cc / 16;
      const uint32_t n_stages = (uint32_t)p.stages, stage_units = stage_bytes >> 4, b_off = p.a_slot >> 4;
      const uint32_t sw_hi = (p.sbo >> 4) | (1u << 14) | (p.layout_type << 29);
      const uint32_t ring_lo0 = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t stage = 0, phase = 0;
      int ti = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += n_clusters, ++ti) {
        const int as = ti & 1;
        mbar_wait(acce_bar + 8 * as, ((ti >> 1) & 1) ^ 1, p.err_flag, 114);  // both epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * acc_stride;
        if (elect_one()) {
          uint32_t st = stage, ph = phase;
          uint32_t a_lo = ring_lo0 + st * stage_units;
          for (int k = 0; k < k_iters; ++k) {
            mbar_wait(full_bar + 8 * st, ph, p.err_flag, 112);
            tc_fence_after();
            if (kk_n == 4) pair_issue<4>(tmem_acc, a_lo, a_lo + b_off, sw_hi, idesc, k != 0);
            else if (kk_n == 2) pair_issue<2>(tmem_acc, a_lo, a_lo + b_off, sw_hi, idesc, k != 0);
            else pair_issue<1>(tmem_acc, a_lo, a_lo + b_off, sw_hi, idesc, k != 0);
            umma2_commit_both(empty_bar + 8 * st);  // frees the stage in BOTH CTAs once these MMAs retire
            if (++st == n_stages) { st = 0; ph ^= 1; a_lo = ring_lo0; }
            else a_lo += stage_units;
          }
          umma2_commit_both(accf_bar + 8 * as);  // accumulator complete, in both CTAs
        }
        __syncwarp();
        uint32_t adv = stage + (uint32_t)k_iters;
        while (adv >= n_stages) { adv -= n_stages; phase ^= 1; }
        stage = adv;
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue (warps 2..9, both CTAs): as the CTA-wide
    // epilogue of conv_tc_body (incl. its fat form), on this CTA's 128 rows
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int tw = r % p.Tw, th = (r / p.Tw) % p.Th, tn = r / (p.Tw * p.Th);
    const bool leader = ew == 0 && lane == 0;
    const uint32_t esz = p.out_f32 ? 4u : 2u;
    const int halves = p.cw / kCpw;
    const uint32_t pitch = (uint32_t)p.cw * esz;
    const uint32_t swz = (r / (128u / pitch)) & (pitch / 16u - 1u);
    const uint32_t stg_bytes = 128u * pitch;
    const uint32_t row_addr = r * pitch;
    const int n_chunks = (p.BN + p.cw - 1) / p.cw;
    const bool active = half < halves;
    const uint32_t acce0 = mapa_u32(acce_bar, 0);  // the leader's accumulator-empty barriers
    int ti = 0;
    uint32_t sb = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += n_clusters, ++ti) {
      int nt, w0, h0, n0;
      decode(pt, nt, w0, h0, n0);
      const int ow = w0 + tw, oh = h0 + th, on = n0 + tn;
      const bool valid = (r < p.Tw * p.Th * p.Tn) && ow < p.Wout && oh < p.Hout && on < p.B;
      const size_t pix = (static_cast<size_t>(on) * p.Hout + oh) * p.Wout + ow;
      const size_t rpix = p.res_pre ? (static_cast<size_t>(on) * (p.Hout >> 1) + (oh >> 1)) * (p.Wout >> 1) + (ow >> 1) : pix;
      const __nv_bfloat16* res_row = static_cast<const __nv_bfloat16*>(p.res) + rpix * p.res_ct + p.res_co + nt * p.BN;
      const int as = ti & 1;
      mbar_wait(accf_bar + 8 * as, (ti >> 1) & 1, p.err_flag, 113);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * acc_stride + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c = 0; c < n_chunks; ++c) {
        const int col = c * p.cw + half * kCpw;
        const uint32_t dst = staging_base + sb * stg_bytes + row_addr;
        if (active && col < p.BN) {
          const bool two = kCpw == 32 && col + 16 < p.BN;
          uint4 ra0 = make_uint4(0, 0, 0, 0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
          if (p.res && valid) {
            ra0 = *reinterpret_cast<const uint4*>(res_row + col);
            ra1 = *reinterpret_cast<const uint4*>(res_row + col + 8);
            if (two) {
              rb0 = *reinterpret_cast<const uint4*>(res_row + col + 16);
              rb1 = *reinterpret_cast<const uint4*>(res_row + col + 24);
            }
          }
          uint32_t va[16], vb[16];
          tmem_ld16(taddr + col, va);
          if (two) tmem_ld16(taddr + col + 16, vb);
          tmem_ld_wait();
          const int n = nt * p.BN + col;
          epi16<false>(p, va, ra0, ra1, n, dst, (uint32_t)(half * kCpw) / 8u, swz);
          if (two) epi16<false>(p, vb, rb0, rb1, n + 16, dst, (uint32_t)(half * kCpw + 16) / 8u, swz);
          fence_async_smem();
        }
        if (leader) bulk_wait_read0();
        named_bar_sync(1, kEpiWarps * 32);
        if (leader) {
          tma_store_4d(&maps.out, staging_base + sb * stg_bytes, nt * p.BN + c * p.cw, w0, h0, n0);
          bulk_commit();
        }
        sb ^= 1u;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acce0 + 8 * as);
    }
    if (leader) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees its TMEM) while the other may still signal it or use the pair's accumulators
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_pair(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcParams p) {
  conv_tc_pair_body<16>(maps, p);
}
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_pair_fat(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcParams p) {
  conv_tc_pair_body<32>(maps, p);
}

__global__ void __launch_bounds__(kThreads, 3)
conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcParams p) {
  conv_tc_body<16, false>(maps, p);
}
// e4m3 variant (operands and / or output in fp8): same body with the quantised paths compiled in; 2 CTAs per SM
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_q(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcParams p) {
  conv_tc_body<16, true>(maps, p);
}
// "fat" epilogue variant: up to 102 registers per thread, 2 CTAs per SM
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_fat(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcParams p) {
  conv_tc_body<32, false>(maps, p);
}

int encode_map(y11_engine* eng, CUtensorMap* m, CUtensorMapDataType dt, int rank, void* base, const cuuint64_t* gdim,
               const cuuint64_t* gstr, const cuuint32_t* box, CUtensorMapSwizzle swz) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = eng->encode_tiled(m, dt, rank, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    y11_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
                  (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)(rank > 2 ? gdim[2] : 0),
                  (unsigned long long)(rank > 3 ? gdim[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return -3;
  }
  return 0;
}

}  // namespace

// Choose the Tw x Th x Tn pixel box (<= 128 GEMM rows) that wastes the fewest MMA rows over the whole layer.
static void pick_tile(int W, int H, int B, int* tw, int* th, int* tn) {
  double best = -1.0;
  for (int w = 1; w <= std::min(W, 128); ++w) {
    for (int h = 1; h <= std::min(H, 128 / w); ++h) {
      const int n = std::min(B, 128 / (w * h));
      if (n < 1) continue;
      const double tiles = double(y11_ceil_div(W, w)) * y11_ceil_div(H, h) * y11_ceil_div(B, n);
      double eff = double(W) * H * B / (tiles * 128.0);
      eff += 1e-6 * w;  // tie-break: longer contiguous runs in w
      if (eff > best) { best = eff; *tw = w; *th = h; *tn = n; }
    }
  }
}

int conv_tc_prepare(y11_engine* eng, const y11_conv_desc* d, ConvTcLaunch* L, const ConvTcTune* tune_in) {
  const ConvTcTune tune = tune_in ? *tune_in : conv_tc_default_tune();
  Y11_REQUIRE(eng && eng->encode_tiled, "conv_tc: engine has no cuTensorMapEncodeTiled entry point");
  Y11_REQUIRE((d->k == 1 && d->stride == 1) || (d->k == 2 && d->stride == 1) || (d->k == 3 && (d->stride == 1 || d->stride == 2)),
              "conv_tc: unsupported k=%d stride=%d", d->k, d->stride);
  const int cin = d->in.c, cout = d->out.c;
  const int in_esz = d->in_fp8 ? 1 : 2, out_esz = d->out_f32 ? 4 : d->out_fp8 ? 1 : 2;
  Y11_REQUIRE(cin % 16 == 0 && cout % 16 == 0, "conv_tc: cin=%d cout=%d must be multiples of 16", cin, cout);
  Y11_REQUIRE(!(d->out_f32 && d->out_fp8), "conv_tc: out_f32 and out_fp8 exclude each other");
  Y11_REQUIRE(!d->in_fp8 || (cin % 32 == 0 && d->k != 2), "conv_tc: e4m3 input needs cin %% 32 == 0 (cin=%d) and k in {1,3}", cin);
  Y11_REQUIRE(!d->out_fp8 || cout % 32 == 0, "conv_tc: e4m3 output needs cout %% 32 == 0 (cout=%d)", cout);
  Y11_REQUIRE((d->in.c_total * in_esz) % 16 == 0 && (d->in.c_off * in_esz) % 16 == 0, "conv_tc: input view must be 16-byte aligned");
  Y11_REQUIRE((d->out.c_off * out_esz) % 16 == 0 && (d->out.c_total * out_esz) % 16 == 0,
              "conv_tc: output view must be 16-byte aligned");
  Y11_REQUIRE(!d->res.ptr || (d->res.c_off % 8 == 0 && d->res.c_total % 8 == 0), "conv_tc: residual view alignment");
  if (d->stride == 1) Y11_REQUIRE(d->Hout == d->Hin && d->Wout == d->Win, "conv_tc: stride-1 shape mismatch");
  if (d->stride == 2) Y11_REQUIRE(d->Hout == (d->Hin + 1) / 2 && d->Wout == (d->Win + 1) / 2, "conv_tc: stride-2 shape mismatch");

  std::memset(L, 0, sizeof(*L));
  ConvTcParams& p = L->p;
  // LSU-producer mode.  The TMA unit sustains only ~0.1 box rows (<= 128 B each) per cycle per SM when the rows miss to
  // DRAM (measured with tools/trace_conv.py: 1x1 layers with 32-, 64- and 128-byte pixel rows all take time proportional
  // to their ROW count, 1 or 3 CTAs per SM alike), i.e. 3-13 B/clk/SM, and a 3x3 layer issues nine such rows per pixel.
  // Layers whose whole K extent fits in shared memory therefore fetch the activation tile with cp.async (coalesced
  // 16-byte copies, weights resident for the CTA's lifetime):
  //   1x1         : the 128 tile pixels, any tile shape;
  //   3x3 stride 1: the (Tw+2)x(Th+2) halo ONCE (1.4 instead of 9 rows per pixel); the nine taps are nine descriptor
  //                 start offsets into the un-swizzled core-matrix layout, which needs Tw = 8.
  {
    const char* e = getenv("Y11_HALO");
    const int mode = tune.lsu >= 0 ? (tune.lsu ? 3 : 0) : (e ? atoi(e) : 3);  // bit 0: 3x3 halo tiles, bit 1: 1x1
    const size_t wbytes = (size_t)d->k * d->k * cin * cout * 2;
    const bool bf16_io = !d->in_fp8 && !d->out_fp8;   // the cp.async producer / un-swizzled layouts are bf16 only
    const bool k3e_any = bf16_io && d->k == 3 && d->stride == 1 && d->Hout >= 32 && d->Wout >= 32;
    const bool k3e = k3e_any && cin <= 64;
    const bool k1e = bf16_io && d->k == 1 && (cin <= 32 || (cin % 32 != 0 && cin <= 112));  // TMA rows would be 32-64 B
    // resident-weight budget: 40 KB keeps 2-3 CTAs per SM; up to Y11_LSU_WMAX KB (default 80: the 3x3 64->64 layers, 72 KB) the
    // mode is still OFFERED to the autotuner (1 CTA per SM, 4 halo stages) - those layers are bound by the L2->SM operand
    // traffic of the tap-by-tap TMA mode (9 activation + 9 weight tiles per output tile, ~45 B/clk/SM against the ~43 B/clk/SM
    // the L2 delivers chip-wide), and the halo mode moves a tenth of it; the heuristic default stays TMA above 40 KB
    static const size_t wmax = [] { const char* e = getenv("Y11_LSU_WMAX"); return (size_t)(e ? atoi(e) : 80) * 1024; }();
    const bool big = wbytes > 40 * 1024;
    const bool fits = cout <= 128 && wbytes <= (tune.lsu >= 1 ? wmax : (size_t)40 * 1024);
    const bool k3 = k3e && (mode & 1), k1 = k1e && (mode & 2);
    L->lsu_eligible = (k3e || k1e) && cout <= 128 && wbytes <= wmax;
    (void)big;
    p.halo = (k1 || k3) && fits;
    p.pad = (p.halo && k3) ? 1 : 0;
    // TMA-halo mode (tune.lsu == 2): 3x3 stride-1 layers with exactly one 128-byte channel chunk (cin = 64) and resident weights.
    // Isolated, L2 flushed, batch 64 (tools/halo_sw_probe.py): 64 -> 64 on 80x80 46 us against 52-56 us for the cp.async halo, the
    // tap-by-tap TMA mode and the CTA pair; 64 -> 32 35 against 43; 64 -> 64 on 40x40 21.5 against 23.5.  It is the heuristic's choice
    // where it applies (Y11_HALO_TMA=0 removes it) and an autotuner candidate like the others (same K order: bit-identical results).
    static const bool halo_tma_on = [] { const char* e = getenv("Y11_HALO_TMA"); return e ? atoi(e) != 0 : true; }();
    const bool k3t = halo_tma_on && k3e && cin == 64 && cout <= 128 && wbytes <= wmax;
    L->halo_tma_eligible = k3t;
    if (tune.lsu == 2) Y11_REQUIRE(k3t, "conv_tc: the TMA-halo mode needs a 3x3 stride-1 layer with cin = 64 and <= %d KB of weights", (int)(wmax >> 10));
    if (k3t && (tune.lsu == 2 || (tune.lsu < 0 && (mode & 1)))) { p.halo = 1; p.pad = 1; p.halo_tma = 1; }
    // Halo-stream mode (tune.lsu == 3): the same swizzled TMA halo for cin = 64 or 128 (one or two 64-channel chunks per tile, two
    // tile buffers) with the weights STREAMED per (tap, chunk) through the stage ring instead of resident - for the layers whose
    // weights do not fit (64 -> 128, 128 -> 64, 128 -> 128: 147-295 KB).  Against the tap-by-tap TMA mode the activation tile crosses
    // the L2 -> shared-memory path 1.4 instead of 9 times per output tile.
    static const bool hstream_on = [] { const char* e = getenv("Y11_HALO_STREAM"); return e ? atoi(e) != 0 : true; }();
    const bool k3s = hstream_on && k3e_any && (cin == 64 || cin == 128);
    L->hstream_eligible = k3s;
    if (tune.lsu == 3) Y11_REQUIRE(k3s, "conv_tc: the halo-stream mode needs a 3x3 stride-1 layer with cin = 64 or 128 on a map >= 32x32");
    static const bool hstream_default = [] { const char* e = getenv("Y11_HALO_STREAM_DEFAULT"); return e ? atoi(e) != 0 : false; }();
    if (k3s && !p.halo_tma && (tune.lsu == 3 || (tune.lsu < 0 && hstream_default))) { p.halo = 0; p.pad = 0; p.hstream = 1; }
  }
  if ((p.halo && p.pad) || p.hstream) { p.Tw = 8; p.Th = 16; p.Tn = 1; }
  else pick_tile(d->Wout, d->Hout, d->B, &p.Tw, &p.Th, &p.Tn);
  p.tiles_w = y11_ceil_div(d->Wout, p.Tw);
  p.tiles_h = y11_ceil_div(d->Hout, p.Th);
  p.tiles_n = y11_ceil_div(d->B, p.Tn);
  // N tile: largest multiple of 16 that divides cout and is <= 128
  int bn = 16;
  const int bn_cap = tune.bn_max > 0 ? tune.bn_max : 256;
  for (int c = 16; c <= std::min(cout, std::min(128, bn_cap)); c += 16)
    if (cout % c == 0) bn = c;
  // Long-K, wide layers are bound by L2->smem operand traffic at 128x128 tiles (64 FLOP/B): a 128x256 tile re-reads the
  // activation tile half as often (85 FLOP/B).  It needs all 512 TMEM columns (2 accumulator stages), i.e. 1 CTA per SM.
  {
    const char* e = getenv("Y11_BN256");
    const int mode = e ? atoi(e) : 1;
    const long long k_total = (long long)cin * d->k * d->k;
    // an explicit bn_max >= 256 (autotuner candidate / cached variant) asks for the wide tile on shorter K as well: the 1x1
    // layers with 256-512 output channels re-read their activation tile once per 128-column N tile otherwise
    const bool forced = tune.bn_max >= 256;
    if (mode && bn_cap >= 256 && cout % 256 == 0 && (cin * in_esz) % 128 == 0 && (k_total >= 1024 || forced)) bn = 256;
  }
  if (p.halo) bn = cout;
  if (p.hstream && bn > 128) bn = 128;
  p.BN = bn;
  p.n_tiles = cout / bn;
  // K chunk per pipeline stage = one swizzle row: 128 / 64 / 32 bytes of channels
  if (d->in_fp8) p.Cc = (cin % 128 == 0) ? 128 : (cin % 64 == 0) ? 64 : 32;
  else p.Cc = (cin % 64 == 0) ? 64 : (cin % 32 == 0) ? 32 : 16;
  if (p.hstream) p.Cc = 64;
  p.cin = cin;
  p.in_fp8 = d->in_fp8 ? 1 : 0;
  p.out_esz = out_esz;
  p.cscale = d->cscale;
  p.oscale = d->out_fp8 ? d->out_scale : 1.0f;
  p.chunks_per_tap = cin / p.Cc;
  p.taps = d->k * d->k;
  if (d->s2d_block) {
    const int c = d->s2d_block;
    Y11_REQUIRE(d->k == 2 && (c == 32 || c == 64) && cin == 4 * c && !d->in_fp8, "conv_tc: s2d_block=%d needs k = 2 and cin = 4 * block (cin=%d)", c, cin);
    p.Cc = 64;
    p.chunks_per_tap = cin / 64;
    // permuted block order [(1,0), (1,1), (0,1), (0,0)]: block range a tap (ty, tx) can touch
    static const int lo[4] = {1, 0, 1, 0}, hi[4] = {2, 2, 3, 4};
    for (int tap = 0; tap < 4; ++tap) {
      const int c0 = lo[tap] * c, n = ((hi[tap] - lo[tap]) * c + 63) / 64;
      for (int j = 0; j < n; ++j) p.kt[p.n_kt++] = (uint16_t)(tap | (((c0 + 64 * j) / 16) << 2));
    }
  }
  p.ksize = d->k;
  p.stride = d->stride;
  const uint32_t swz_bytes = p.Cc * in_esz;
  p.sbo = 8 * swz_bytes;
  p.layout_type = (swz_bytes == 128) ? 2u : (swz_bytes == 64) ? 4u : 6u;  // SWIZZLE_128B / 64B / 32B
  const CUtensorMapSwizzle swz = (swz_bytes == 128)  ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (swz_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
  p.a_slot = 128u * swz_bytes;
  p.b_slot = ((uint32_t)bn * swz_bytes + 1023u) & ~1023u;
  // CTA-pair variant (tune.epi_warp bit 3): each CTA of the pair holds half of the weight tile
  // heuristic (no explicit variant): the long-K 3x3 layers on 128 x 256 tiles, which the autotuner moved to the pair kernel on
  // every scale measured (they are shared-memory-bandwidth bound as single CTAs)
  static const bool pair_default = [] { const char* e = getenv("Y11_PAIR"); return e ? atoi(e) != 0 : true; }();
  const bool want_pair = tune.epi_warp >= 0 ? (tune.epi_warp & 8) != 0 : (pair_default && d->k == 3 && bn == 256);
  p.pair = want_pair && !p.halo && !p.hstream && !d->in_fp8 && !d->out_fp8 && !d->cscale && d->k != 2 &&
           bn >= 64 && bn % 32 == 0 && p.tiles_w * p.tiles_h * p.tiles_n >= 2;
  if (p.pair) p.b_slot = ((uint32_t)(bn / 2) * swz_bytes + 1023u) & ~1023u;
  // Resident weights in TMA mode: a layer with ONE N tile whose whole weight matrix is small keeps it in shared memory for
  // the CTA's lifetime instead of re-fetching a B tile with every K stage of every tile (model.1: 8 KB of the 24 KB per stage;
  // the store-heavy 1x1 layers on the large maps: a third to a half of their L2 -> shared-memory traffic).
  {
    // Measured (A/B per layer, YOLO11s B=64): it pays where a CTA walks many tiles and the weights are small enough to leave the
    // CTAs-per-SM count and the epilogue staging alone (model.2.cv2 166 -> 144 us); on maps with one or two tiles per CTA the
    // up-front load is exposed latency, and 32-40 KB of resident weights cost a CTA slot or the fat epilogue (model.1, model.4.cv1).
    static const int bres_kb = [] { const char* e = getenv("Y11_BRES_KB"); return e ? atoi(e) : 24; }();
    const int nk = p.n_kt ? p.n_kt : p.taps * p.chunks_per_tap;
    const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    // explicit variant (autotuner candidate / cached choice): bit 2 of epi_warp says it; heuristic otherwise
    const bool fits = !p.halo && !p.hstream && !p.pair && !d->in_fp8 && cout == bn;
    if (tune.epi_warp >= 0) p.bres = fits && (tune.epi_warp & 4) && (size_t)nk * p.b_slot <= (size_t)64 * 1024;
    else p.bres = fits && (size_t)nk * p.b_slot <= (size_t)bres_kb * 1024 && m_tiles >= 8ll * 3 * eng->num_sms;
    if (p.bres) p.b_res_bytes = (uint32_t)nk * p.b_slot;
  }
  if (p.halo) {
    p.n_pos = (p.Tw + 2 * p.pad) * (p.Th + 2 * p.pad) * p.Tn;
    const uint32_t plane = (uint32_t)(p.pad ? p.n_pos : 128) * 16u;  // one 8-channel plane: 16 B per position
    p.a_slot = (plane * (uint32_t)(cin / 8) + 1023u) & ~1023u;
    p.a_lbo = plane;
    p.a_sbo = p.pad ? (uint32_t)(p.Tw + 2) * 16u : 128u;  // next 8 GEMM rows: next halo row (Tw = 8) / next 8 pixels
    if (const char* e = getenv("Y11_HALO_SWAP")) if (atoi(e)) std::swap(p.a_lbo, p.a_sbo);
    p.b_res_bytes = (uint32_t)(p.taps * p.chunks_per_tap) * p.b_slot;
    if (p.halo_tma) {  // one 128-byte row per halo pixel; 8-row groups of the GEMM are one halo row (Tw + 2 pixels) apart
      p.a_slot = ((uint32_t)p.n_pos * 128u + 1023u) & ~1023u;
      p.a_lbo = 16u;
      p.a_sbo = (uint32_t)(p.Tw + 2) * 128u;
    }
    p.in = static_cast<const __nv_bfloat16*>(d->in.ptr) + d->in.c_off;
    p.in_ct = d->in.c_total; p.Hin = d->Hin; p.Win = d->Win;
    p.mg_ncg = ((1ull << 42) + (cin / 8) - 1) / (cin / 8);
  }
  if (p.hstream) {  // [2 tile buffers x cin/64 halo chunks of (Tw+2)(Th+2) 128-byte rows] in front of the ring of weight tiles
    p.n_pos = (p.Tw + 2) * (p.Th + 2);
    p.a_slot = ((uint32_t)p.n_pos * 128u + 1023u) & ~1023u;
    p.a_lbo = 16u;
    p.a_sbo = (uint32_t)(p.Tw + 2) * 128u;
    p.hbufs = 2;
    p.b_res_bytes = 2u * (uint32_t)p.chunks_per_tap * p.a_slot;
  }
  p.tx_bytes = (uint32_t)(p.Tw * p.Th * p.Tn) * swz_bytes + (p.bres ? 0u : (uint32_t)(p.pair ? bn / 2 : bn) * swz_bytes);
  if (p.hstream) p.tx_bytes = (uint32_t)bn * swz_bytes;
  const int k_iters = p.taps * p.chunks_per_tap;
  const uint32_t stage = p.hstream ? p.b_slot : (p.halo || p.bres) ? p.a_slot : p.a_slot + p.b_slot;
  // Warp-independent epilogue: possible when the tile has exactly 128 rows and every 32-row quarter (one TMEM lane
  // quadrant) is itself a (qbw x qbh x qbn) box of pixels, so that each warp can TMA-store its own rows.
  int qbw = 0, qbh = 0, qbn = 0;
  {
    const int Tw = p.Tw, Th = p.Th, Tn = p.Tn;
    bool ok = Tw * Th * Tn == 128;
    if (ok) {
      if (Tw >= 32) { ok = Tw % 32 == 0; qbw = 32; qbh = 1; qbn = 1; }
      else if (32 % Tw != 0) ok = false;
      else {
        const int rows_h = 32 / Tw;
        if (Th >= rows_h) { ok = Th % rows_h == 0; qbw = Tw; qbh = rows_h; qbn = 1; }
        else { ok = rows_h % Th == 0 && Tn % (rows_h / Th) == 0; qbw = Tw; qbh = Th; qbn = rows_h / Th; }
      }
    }
    const char* e = getenv("Y11_EPI_WARP");
    L->epi_warp_possible = ok;
    // tune.epi_warp: bit 0 = warp-independent epilogue, bit 1 = "fat" epilogue (32 columns per warp step, 64-channel chunks)
    const int mode = tune.epi_warp >= 0 ? tune.epi_warp : (e ? atoi(e) : 0);
    p.epi_warp = ok && (mode & 1);
    // fat: bf16 outputs only (64 fp32 channels would be 256-byte staging rows), and the chunk grid must tile BN
    p.fat = (mode & 2) && !d->out_f32 && bn > 32 && (bn % 64 == 0 || p.n_tiles == 1);
    p.quant = (d->in_fp8 || d->out_fp8 || d->cscale) ? 1 : 0;
    if (p.quant) { p.epi_warp = 0; p.fat = 0; }
    if (p.pair) p.epi_warp = 0;  // the pair kernel has the CTA-wide epilogue (plain or fat)  // e4m3 kernel: CTA-wide epilogue (32-byte staging rows for e4m3 stores)
  }
  // epilogue chunk width: 32 output channels per TMA store when the tile allows it, else 16 (fp32 rows: 16 in warp mode);
  // 64 in fat mode
  int cw = (bn % 32 == 0) ? 32 : 16;
  if (d->out_fp8) Y11_REQUIRE(bn % 32 == 0, "conv_tc: e4m3 output needs an N tile that is a multiple of 32 (BN=%d)", bn);
  if (p.epi_warp && d->out_f32) cw = 16;
  if (p.fat) cw = 64;
  if (p.n_tiles > 1) Y11_REQUIRE(bn % cw == 0, "conv_tc: BN=%d not a multiple of the chunk width", bn);
  p.cw = cw;
  // staging for the TMA-store epilogue: CTA-wide mode = ring of 2 whole-tile buffers (3 and 4 measured no faster and cost
  // load stages); warp mode = 2 private 32-row buffers per epilogue warp
  p.nstg = 2;
  const uint32_t opitch_b = (uint32_t)cw * (uint32_t)out_esz;
  // CTA-wide mode: nstg whole-tile (128-row) buffers; warp mode: nstg private 32-row buffers for each of the 8 warps
  uint32_t staging = (uint32_t)p.nstg * 128u * opitch_b;
  if (p.epi_warp) {
    staging = (uint32_t)kEpiWarps * p.nstg * 32u * opitch_b;
    {
      // double-buffer the per-warp staging only when that still leaves >= 3 load stages at 3 CTAs per SM
      const uint32_t avail = (220u * 1024u) / 3 - 2048u, need = kHeaderBytes + 1024u + staging + p.b_res_bytes;
      if (avail < need || (avail - need) / stage < 3) {
        p.nstg = 1;
        staging = (uint32_t)kEpiWarps * 32u * opitch_b;
      }
    }
  }
  // persistent CTAs per SM: TMEM (512 columns) and shared memory are split between them
  // (measured on B200, YOLO11s batch 64: conv time 5.25 / 3.78 / 3.65 / 3.88 ms for 1 / 2 / 3 / 4 CTAs per SM - several
  //  independent TMA->MMA->epilogue chains per SM hide the per-tile latencies better than one deep pipeline)
  int cps = 3;
  if (const char* e = getenv("Y11_CTAS_PER_SM")) cps = std::max(1, std::min(4, atoi(e)));
  if (tune.cps > 0) cps = std::max(1, std::min(4, tune.cps));
  if (p.fat || p.quant || p.pair) cps = std::min(cps, 2);  // conv_tc_kernel_fat / _q are compiled for 2 CTAs per SM (up to 102 registers)
  int cols = 64;  // >= 2 accumulator stages of max(BN, 32) columns (a partial last chunk may read up to 16 spare columns)
  while (cols < 2 * bn) cols *= 2;
  while (cps > 1 && cps * cols > 512) --cps;
  auto budget_of = [](int c) { return c == 1 ? kSmemBudget : (220u * 1024u) / c - 2048u; };
  const uint32_t fixed = kHeaderBytes + 1024u + staging + p.b_res_bytes;
  if (p.halo) {  // resident weights + at least two halo tiles must fit
    while (cps > 1 && budget_of(cps) < fixed + 2 * stage) --cps;
    Y11_REQUIRE(budget_of(cps) >= fixed + 2 * stage, "conv_tc: halo tile does not fit in shared memory");
  }
  uint32_t fixed_v = fixed;
  if (p.hstream) {  // the halo tile buffers + at least four weight stages must fit; ONE tile buffer where two would cost a CTA slot
    if (cps >= 2 && budget_of(2) < fixed + 4 * stage && budget_of(2) >= fixed - p.b_res_bytes / 2 + 4 * stage) {
      p.hbufs = 1;
      p.b_res_bytes /= 2;
      fixed_v = fixed - p.b_res_bytes;
    }
    while (cps > 1 && budget_of(cps) < fixed_v + 4 * stage) --cps;
    Y11_REQUIRE(budget_of(cps) >= fixed_v + 3 * stage, "conv_tc: halo-stream buffers do not fit in shared memory");
  }
  if (p.bres) {  // resident weights + at least three activation stages must fit
    while (cps > 1 && budget_of(cps) < fixed + 3 * stage) --cps;
    Y11_REQUIRE(budget_of(cps) >= fixed + 2 * stage, "conv_tc: resident weights do not fit in shared memory");
  }
  const uint32_t budget = budget_of(cps);
  int stages = budget > fixed_v ? (int)((budget - fixed_v) / stage) : 0;
  stages = std::max(2, std::min(stages, p.halo ? 6 : kMaxStages));
  p.stages = stages;
  p.tmem_cols = cols;
  p.B = d->B; p.Hout = d->Hout; p.Wout = d->Wout; p.cout = cout;
  p.out = d->out.ptr; p.out_ct = d->out.c_total; p.out_co = d->out.c_off; p.out_f32 = d->out_f32;
  p.res = d->res.ptr; p.res_ct = d->res.c_total; p.res_co = d->res.c_off;
  p.res_pre = (d->res.ptr && d->res_mode == Y11_RES_PRE_UP2) ? 1 : 0;
  if (p.res_pre) Y11_REQUIRE(d->Hout % 2 == 0 && d->Wout % 2 == 0, "conv_tc: RES_PRE_UP2 needs an even output size");
  p.bias = d->bias; p.act = d->act;
  p.err_flag = eng->dev_error_flag;
#ifdef Y11_TRACE
  p.trace = g_trace_buf;
#endif
  {
    const long long tt = (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
    Y11_REQUIRE(tt < (1ll << 21), "conv_tc: %lld tiles exceed the fast-division range", tt);
    auto magic = [](uint32_t dv) { return ((1ull << 42) + dv - 1) / dv; };
    p.mg_ntiles = magic(p.n_tiles); p.mg_tw = magic(p.tiles_w); p.mg_th = magic(p.tiles_h);
  }

  // activation tensor maps
  const size_t ct = d->in.c_total;
  const size_t ie = (size_t)in_esz;
  char* in_base = static_cast<char*>(d->in.ptr) + (size_t)d->in.c_off * ie;
  const bool halo_box = p.halo_tma || p.hstream;
  const cuuint32_t box[4] = {(cuuint32_t)p.Cc, (cuuint32_t)(halo_box ? p.Tw + 2 : p.Tw), (cuuint32_t)(halo_box ? p.Th + 2 : p.Th),
                             (cuuint32_t)p.Tn};
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapDataType it = d->in_fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : bf;
  if (d->stride == 1) {
    const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
    const cuuint64_t gstr[3] = {ct * ie, ct * ie * d->Win, ct * ie * d->Win * d->Hin};
    if (int e = encode_map(eng, &L->maps.a[0], it, 4, in_base, gdim, gstr, box, swz)) return e;
    L->maps.a[1] = L->maps.a[2] = L->maps.a[3] = L->maps.a[0];
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)((d->Win - pw + 1) / 2), (cuuint64_t)((d->Hin - ph + 1) / 2),
                                    (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {ct * ie * 2, ct * ie * d->Win * 2, ct * ie * d->Win * d->Hin};
        char* base = in_base + ((size_t)ph * d->Win + pw) * ct * ie;
        if (int e = encode_map(eng, &L->maps.a[ph * 2 + pw], it, 4, base, gdim, gstr, box, swz)) return e;
      }
  }
  {
    const cuuint64_t K = p.n_kt ? (cuuint64_t)p.n_kt * p.Cc : (cuuint64_t)p.taps * cin;
    const cuuint64_t gdim[2] = {K, (cuuint64_t)cout};
    const cuuint64_t gstr[1] = {K * ie};
    const cuuint32_t bbox[2] = {(cuuint32_t)p.Cc, (cuuint32_t)(p.pair ? bn / 2 : bn)};
    if (int e = encode_map(eng, &L->maps.b, it, 2, const_cast<void*>(d->w), gdim, gstr, bbox, swz)) return e;
  }
  {
    // output view: 16-channel boxes (32 B bf16 / 64 B fp32 per pixel), swizzled staging, clipped at the tensor edge
    const size_t esz = (size_t)out_esz;
    const CUtensorMapDataType ot = d->out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : d->out_fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : bf;
    const size_t oct = d->out.c_total;
    char* obase = static_cast<char*>(d->out.ptr) + (size_t)d->out.c_off * esz;
    const cuuint64_t gdim[4] = {(cuuint64_t)cout, (cuuint64_t)d->Wout, (cuuint64_t)d->Hout, (cuuint64_t)d->B};
    const cuuint64_t gstr[3] = {oct * esz, oct * esz * d->Wout, oct * esz * d->Wout * d->Hout};
    const cuuint32_t obox[4] = {(cuuint32_t)p.cw, (cuuint32_t)p.Tw, (cuuint32_t)p.Th, (cuuint32_t)p.Tn};
    const uint32_t opitch = (uint32_t)p.cw * (uint32_t)esz;
    const CUtensorMapSwizzle oswz = opitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : opitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    if (int e = encode_map(eng, &L->maps.out, ot, 4, obase, gdim, gstr, obox, oswz))
      return e;
    L->maps.outq = L->maps.out;
    if (p.epi_warp) {
      const cuuint32_t qbox[4] = {(cuuint32_t)p.cw, (cuuint32_t)qbw, (cuuint32_t)qbh, (cuuint32_t)qbn};
      if (int e = encode_map(eng, &L->maps.outq, ot, 4, obase, gdim, gstr, qbox, oswz))
        return e;
    }
  }
  if (d->k == 2 && !p.halo) {
    // zero-block scan of the weights (one-off, at plan-build time): K step (k_iter, kk) = 16 consecutive K elements
    const int k_total = p.n_kt ? p.n_kt * p.Cc : p.taps * cin, steps = k_total / 16;
    if (steps <= 64) {
      std::vector<__nv_bfloat16> hw((size_t)cout * k_total);
      Y11_CHECK_CUDA(cudaMemcpy(hw.data(), d->w, hw.size() * sizeof(__nv_bfloat16), cudaMemcpyDeviceToHost));
      unsigned long long mask = 0ull;
      for (int st = 0; st < steps; ++st) {
        bool nz = false;
        for (int co = 0; co < cout && !nz; ++co)
          for (int e = 0; e < 16; ++e)
            if (__bfloat162float(hw[(size_t)co * k_total + st * 16 + e]) != 0.0f) { nz = true; break; }
        if (nz) mask |= 1ull << st;
      }
      if (mask != 0ull && mask != (steps == 64 ? ~0ull : ((1ull << steps) - 1ull))) { p.kmask = mask; p.kmask_on = 1; }
    }
  }
  const unsigned total_tiles = (unsigned)(p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles);
  L->grid = std::min(total_tiles, (unsigned)(eng->num_sms * cps));
  if (p.pair) {
    const unsigned m_tiles = (unsigned)(p.tiles_w * p.tiles_h * p.tiles_n);
    const unsigned total_pairs = ((m_tiles + 1) / 2) * (unsigned)p.n_tiles;
    L->grid = 2u * std::min(total_pairs, (unsigned)((eng->num_sms / 2) * cps));
  }
  L->variant = ConvTcTune{p.hstream ? 3 : p.halo ? (p.halo_tma ? 2 : 1) : 0, p.epi_warp | (p.fat << 1) | (p.bres << 2) | (p.pair << 3), cps, bn};
  L->smem_bytes = kHeaderBytes + 1024u + p.b_res_bytes + (unsigned)stages * stage + staging;
  L->flops = 2.0 * d->B * d->Hout * d->Wout * (double)cout * cin * p.taps;
  (void)k_iters;
  Y11_OPT_IN_SMEM(conv_tc_kernel, 220 * 1024);
  Y11_OPT_IN_SMEM(conv_tc_kernel_fat, 220 * 1024);
  Y11_OPT_IN_SMEM(conv_tc_kernel_q, 220 * 1024);
  Y11_OPT_IN_SMEM(conv_tc_kernel_pair, 220 * 1024);
  Y11_OPT_IN_SMEM(conv_tc_kernel_pair_fat, 220 * 1024);
  return 0;
}

int conv_tc_launch(const ConvTcLaunch* L, cudaStream_t s) {
  if (L->p.pair) {
    Y11_CHECK_CUDA(y11_launch_pdl_pair(L->p.fat ? conv_tc_kernel_pair_fat : conv_tc_kernel_pair, dim3(L->grid), dim3(kThreads),
                                       L->smem_bytes, s, L->maps, L->p));
    return 0;
  }
  auto* kern = L->p.quant ? conv_tc_kernel_q : L->p.fat ? conv_tc_kernel_fat : conv_tc_kernel;
  Y11_CHECK_CUDA(y11_launch_pdl(kern, dim3(L->grid), dim3(kThreads), L->smem_bytes, s, L->maps, L->p));
  return 0;
}
