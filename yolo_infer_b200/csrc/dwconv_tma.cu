// Depthwise 3x3 (stride 1, pad 1) for the large feature maps: persistent CTAs, TMA-fed shared-memory ring.
//
// Replaces ultralytics DWConv.forward_fuse (Detect class tower) and Attention.pe.  The
// register-only kernel in conv_simt.cu is latency bound: the bytes a thread can keep in flight are limited by its
// registers (18 loads x 8 B at 95 registers/thread = 74 KB per SM, ~2.7 us per warp batch -> 1.7 TB/s).  Here the bytes in
// flight are whole halo tiles owned by the TMA engine:
//   tile    = 8 rows x 16 pixels x CC channels of one image (CC <= 128, a multiple of 8); its (8+2) x (16+2) halo is ONE
//             4-D TMA box load (rows of CC*2 >= 128 bytes; outside the image the TMA unit writes zeros == conv padding);
//   ring    = 3 stages (up to 46 KB each), one producer warp runs two tiles ahead;
//   compute = 512 threads, thread = 4 channels x 2 columns: the 9x4 weights live in registers as packed fp32 pairs, the halo
//             rows slide through three row accumulators (4 LDS.64 + 36 FFMA2 per halo row), bias/SiLU/residual fused,
//             one warp writes the 32 x 8 B = 256 contiguous bytes of a pixel.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ops.h"

namespace {

using namespace y11;

constexpr int kTH = 8, kTW = 16;
constexpr int kHaloH = kTH + 2, kHaloW = kTW + 2;
constexpr int kComputeThreads = 512;
constexpr int kThreadsDw = kComputeThreads + 32;
constexpr int kMaxStagesDw = 6;  // ring depth is per launch: as many halo tiles as fit in ~190 KB, at most 6 (v1: fixed 3)

// One thread's share of a tile: 4 channels x 2 adjacent columns x RPS output rows starting at tile row r0 (halo rows
// r0 .. r0+RPS+1, halo columns 2*colpair .. 2*colpair+3).
// v3 (round 2): the v2 kernel was ISSUE bound, not HBM bound - ncu: 72 % of the issue slots busy, 30 thread instructions per
// output element (9 FFMA, bf16 unpacking of every input once per neighbour that uses it, SiLU, 64-bit address arithmetic per
// output row), DRAM at 11 %.  Now the two channel pairs of a thread are packed fp32 pairs: 9 taps = 4.5 FFMA2 per element, SiLU =
// FMUL2 + 2 MUFU + FFMA2 per pair; a thread owns two columns, so each halo row costs four 8-byte loads + 8 unpackings for 8
// outputs per row (was 3 + 12 for 4); output / residual addresses advance by a row pitch.  Same per-element operation order
// (taps in (kh, kw) order onto the bias, fma.rn each) -> bit-identical results.
template <int RPS, bool kRes>
__device__ __forceinline__ void dw_rows(uint32_t base, uint32_t row_pitch, uint32_t px_pitch, const f32x2 (&w)[9][2], const f32x2 (&b)[2],
                                        __nv_bfloat16* optr, const __nv_bfloat16* rptr, uint32_t o_row, uint32_t r_row, uint32_t out_ct,
                                        uint32_t res_ct, int rows_ok, bool x_ok0, bool x_ok1, bool act) {
  f32x2 acc[3][2][2];  // [output row in flight: row o lives in acc[o % 3]][column][channel pair]
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) { acc[j][c][0] = b[0]; acc[j][c][1] = b[1]; }
#pragma unroll
  for (int r = 0; r < RPS + 2; ++r) {
    // halo row r of this thread's strip holds input row (first output row) + r - 1: tap row kh of local output row r - kh
    f32x2 f[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t lo, hi;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(base + (uint32_t)r * row_pitch + (uint32_t)j * px_pitch));
      f[j][0] = f2_from_bf16x2(lo);
      f[j][1] = f2_from_bf16x2(hi);
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int orow = r - kh;
      if (orow < 0 || orow >= RPS) continue;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          acc[orow % 3][c][0] = f2_fma(f[c + kw][0], w[kh * 3 + kw][0], acc[orow % 3][c][0]);
          acc[orow % 3][c][1] = f2_fma(f[c + kw][1], w[kh * 3 + kw][1], acc[orow % 3][c][1]);
        }
    }
    const int done = r - 2;  // local output row `done` has now received its three tap rows
    if (done >= 0) {
      const bool row_ok = done < rows_ok;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        f32x2 o0 = acc[done % 3][c][0], o1 = acc[done % 3][c][1];
        const bool ok = row_ok && (c ? x_ok1 : x_ok0);
        if (act) { o0 = silu2(o0); o1 = silu2(o1); }
        if (kRes) {
          uint2 rr = make_uint2(0u, 0u);
          if (ok) rr = *reinterpret_cast<const uint2*>(rptr + (uint32_t)c * res_ct);
          o0 = f2_add(o0, f2_from_bf16x2(rr.x));
          o1 = f2_add(o1, f2_from_bf16x2(rr.y));
        }
        if (ok) *reinterpret_cast<uint2*>(optr + (uint32_t)c * out_ct) = make_uint2(f2_to_bf16x2(o0), f2_to_bf16x2(o1));
        acc[done % 3][c][0] = b[0];
        acc[done % 3][c][1] = b[1];
      }
      optr += o_row;
      if (kRes) rptr += r_row;
    }
  }
}

template <int RPS, bool kRes>
__global__ void __launch_bounds__(kThreadsDw, 1)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ DwTmaParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t full_bar = smem_base, empty_bar = smem_base + 8 * kMaxStagesDw;
  const uint32_t n_stages = (uint32_t)p.stages;
  const uint32_t tiles_base = smem_base + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, kComputeThreads / 32);
    }
    mbar_fence_init();
    prefetch_tmap(&tmap);
  }
  // thread -> (4-channel group, pair of pixel columns, row group): the threads left over after one share per (group, column
  // pair) split the 8 rows between them (128 channels: 2 row groups of 4 rows)
  const int cg = threadIdx.x % p.cg4;
  const int col = ((threadIdx.x / p.cg4) % (kTW / 2)) * 2;
  const int sub = threadIdx.x / (p.cg4 * (kTW / 2));
  const bool worker = warp < kComputeThreads / 32 && sub < p.nsub;
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const uint32_t total = (uint32_t)(p.tiles_w * p.tiles_h * p.B * p.chunks);

  if (warp == kComputeThreads / 32) {
    // ------------------------------------------------------------------ producer
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
      uint32_t t = tile, q;  // tile -> (chunk, tile column, tile row, image) by magic-number division
      q = fast_div(t, p.mg_chunks); const int ck = (int)(t - q * (uint32_t)p.chunks); t = q;
      q = fast_div(t, p.mg_tw); const int tw = (int)(t - q * (uint32_t)p.tiles_w); t = q;
      q = fast_div(t, p.mg_th); const int th = (int)(t - q * (uint32_t)p.tiles_h); t = q;
      mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 201);
#ifdef Y11_DW_PROBE
      if (p.dbg & 2) {
        if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar + 8 * stage) : "memory");
      } else
#endif
      if (elect_one()) {
        mbar_expect_tx(full_bar + 8 * stage, p.stage_tx);
        tma_load_4d(tiles_base + stage * p.stage_bytes, &tmap, full_bar + 8 * stage, ck * p.CC, tw * kTW - 1, th * kTH - 1, (int)t);
      }
      __syncwarp();
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  uint32_t stage = 0, phase = 0;
  const uint32_t row_pitch = (uint32_t)kHaloW * p.CC * 2u, px_pitch = (uint32_t)p.CC * 2u;
  // per-thread constants of the tile loop: strip offset inside a halo tile, row pitches of the output / residual views (elements)
  const int r0 = sub * RPS;
  const uint32_t strip_off = (uint32_t)r0 * row_pitch + (uint32_t)col * px_pitch + (uint32_t)cg * 8u;
  const uint32_t o_row = (uint32_t)p.W * (uint32_t)p.out_ct, r_row = (uint32_t)p.W * (uint32_t)p.res_ct;
  const bool act = p.act == Y11_ACT_SILU;
  f32x2 w[9][2], b[2];  // this thread's 9 x 4 weights and its bias as fp32 pairs
  int w_ck = -1;
  for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
    uint32_t t = tile, q;
    q = fast_div(t, p.mg_chunks); const int ck = (int)(t - q * (uint32_t)p.chunks); t = q;
    q = fast_div(t, p.mg_tw); const int tw = (int)(t - q * (uint32_t)p.tiles_w); t = q;
    q = fast_div(t, p.mg_th); const int th = (int)(t - q * (uint32_t)p.tiles_h); t = q;
    const int c0 = ck * p.CC + cg * 4;
    if (worker && ck != w_ck) {  // this thread's 9x4 weights and bias (constant while the channel chunk does not change)
      const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(p.w) + c0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(wp + (size_t)tap * p.C));
        w[tap][0] = f2_from_bf16x2(u.x);
        w[tap][1] = f2_from_bf16x2(u.y);
      }
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
      b[0] = f2_pack(bb.x, bb.y);
      b[1] = f2_pack(bb.z, bb.w);
      w_ck = ck;
    }
    // first output pixel of this thread's strip and its addresses (one 64-bit multiply-add per view and tile)
    const int x = tw * kTW + col, y = th * kTH + r0;
    const uint32_t pix = (t * (uint32_t)p.H + (uint32_t)y) * (uint32_t)p.W + (uint32_t)x;  // < 2^31 pixels per tensor
    __nv_bfloat16* optr = static_cast<__nv_bfloat16*>(p.out) + (size_t)pix * (uint32_t)p.out_ct + (uint32_t)(p.out_co + c0);
    const __nv_bfloat16* rptr = nullptr;
    if (kRes) rptr = static_cast<const __nv_bfloat16*>(p.res) + (size_t)pix * (uint32_t)p.res_ct + (uint32_t)(p.res_co + c0);
    mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 202);
#ifdef Y11_DW_PROBE
    if (worker && !(p.dbg & 1))
#else
    if (worker)
#endif
      dw_rows<RPS, kRes>(tiles_base + stage * p.stage_bytes + strip_off, row_pitch, px_pitch, w, b, optr, rptr, o_row, r_row,
                         (uint32_t)p.out_ct, (uint32_t)p.res_ct, p.H - y, x < p.W, x + 1 < p.W, act);
    __syncwarp();
    if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_bar + 8 * stage) : "memory");
    }
    if (++stage == n_stages) { stage = 0; phase ^= 1; }
  }
}

}  // namespace

// Eligible: maps large enough to be worth it and channel counts that split into equal chunks of <= 128.
bool dwconv_tma_eligible(const y11_dwconv_desc* d) {
  if (const char* e = getenv("Y11_DW_TMA")) if (!atoi(e)) return false;
  const int C = d->in.c;
  static const int min_w = [] { const char* e = getenv("Y11_DW_TMA_MINW"); return e ? atoi(e) : 18; }();
  if (d->W < min_w || d->W < kHaloW || d->H < (min_w < 32 ? kTH : 2 * kTH)) return false;  // ragged right/bottom tiles are fine: TMA zero-fills, stores are guarded
  // maps narrower than 32 pixels (ragged 16x8 tiles): only with >= 4 channel chunks - 512 channels on 20x20 29.8 -> 23.4 us, but 128
  // channels equal (7.6 vs 7.8 us) and the 256-channel Attention.pe (strided view, residual) slower in the network (34 vs 29 us)
  if (d->W < 32 && C < 512) return false;
  if (C % 8 != 0 || (C > 128 && C % 128 != 0)) return false;
  if (d->in.c_total % 8 != 0 || d->in.c_off % 8 != 0) return false;  // 16-byte aligned TMA base / strides
  return true;
}

int dwconv_tma_prepare(y11_engine* eng, const y11_dwconv_desc* d, DwTmaLaunch* L) {
  Y11_REQUIRE(eng && eng->encode_tiled, "dwconv_tma: engine has no cuTensorMapEncodeTiled entry point");
  std::memset(L, 0, sizeof(*L));
  DwTmaParams& p = L->p;
  const int C = d->in.c;
  p.C = C;
  p.CC = C <= 128 ? C : 128;
  p.chunks = C / p.CC;
  p.cg4 = p.CC / 4;
  Y11_REQUIRE(p.cg4 * (kTW / 2) <= kComputeThreads, "dwconv_tma: chunk of %d channels needs more than %d threads", p.CC, kComputeThreads);
  p.nsub = kComputeThreads / (p.cg4 * (kTW / 2));  // row groups: 1 (8 rows per thread), 2 (4 rows) or 4 (2 rows)
  p.nsub = p.nsub >= 4 ? 4 : p.nsub >= 2 ? 2 : 1;
  p.B = d->B; p.H = d->H; p.W = d->W;
  p.tiles_w = y11_ceil_div(d->W, kTW);
  p.tiles_h = y11_ceil_div(d->H, kTH);
  p.stage_tx = (uint32_t)kHaloH * kHaloW * p.CC * 2u;
  p.stage_bytes = (p.stage_tx + 127u) & ~127u;
  p.w = d->w; p.bias = d->bias; p.act = d->act;
  p.out = d->out.ptr; p.out_ct = d->out.c_total; p.out_co = d->out.c_off;
  p.res = d->res.ptr; p.res_ct = d->res.c_total; p.res_co = d->res.c_off;
  p.err_flag = eng->dev_error_flag;
  {
    auto magic = [](uint32_t dv) { return ((1ull << 42) + dv - 1) / dv; };
    p.mg_chunks = magic((uint32_t)p.chunks); p.mg_tw = magic((uint32_t)p.tiles_w); p.mg_th = magic((uint32_t)p.tiles_h);
    Y11_REQUIRE((long long)p.tiles_w * p.tiles_h * p.B * p.chunks < (1ll << 21), "dwconv_tma: too many tiles for the fast-division range");
    Y11_REQUIRE((long long)d->B * d->H * d->W < (1ll << 31), "dwconv_tma: more than 2^31 pixels");
  }
  const size_t ct = d->in.c_total;
  __nv_bfloat16* base = static_cast<__nv_bfloat16*>(d->in.ptr) + d->in.c_off;
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
  const cuuint64_t gstr[3] = {ct * 2, ct * 2 * d->W, ct * 2 * d->W * d->H};
  const cuuint32_t box[4] = {(cuuint32_t)p.CC, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = eng->encode_tiled(&L->tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  Y11_REQUIRE(r == CUDA_SUCCESS, "dwconv_tma: cuTensorMapEncodeTiled failed (%d) for C=%d %dx%d", (int)r, C, d->H, d->W);
  const int total = p.tiles_w * p.tiles_h * p.B * p.chunks;
  L->grid = (unsigned)std::min(total, eng->num_sms);
  // the compute warps' top stall site was the wait for the next halo tile (ncu source view, 3 stages): run the producer as
  // far ahead as shared memory allows
  p.stages = (int)std::max(3u, std::min((uint32_t)kMaxStagesDw, (190u * 1024u) / p.stage_bytes));
#ifdef Y11_DW_PROBE
  if (const char* e = getenv("Y11_DW_DBG")) p.dbg = atoi(e);
  if (const char* e = getenv("Y11_DW_STAGES")) p.stages = std::max(2, std::min(p.stages, atoi(e)));
#endif
  L->smem_bytes = 128u + (unsigned)p.stages * p.stage_bytes;
  Y11_OPT_IN_SMEM((dwconv_tma_kernel<8, false>), 200 * 1024);
  Y11_OPT_IN_SMEM((dwconv_tma_kernel<8, true>), 200 * 1024);
  Y11_OPT_IN_SMEM((dwconv_tma_kernel<4, false>), 200 * 1024);
  Y11_OPT_IN_SMEM((dwconv_tma_kernel<4, true>), 200 * 1024);
  Y11_OPT_IN_SMEM((dwconv_tma_kernel<2, false>), 200 * 1024);
  Y11_OPT_IN_SMEM((dwconv_tma_kernel<2, true>), 200 * 1024);
  return 0;
}

int dwconv_tma_launch(const DwTmaLaunch* L, cudaStream_t s) {
  const bool res = L->p.res != nullptr;
  auto* kern = L->p.nsub == 1 ? (res ? dwconv_tma_kernel<8, true> : dwconv_tma_kernel<8, false>)
               : L->p.nsub == 2 ? (res ? dwconv_tma_kernel<4, true> : dwconv_tma_kernel<4, false>)
                                : (res ? dwconv_tma_kernel<2, true> : dwconv_tma_kernel<2, false>);
  Y11_CHECK_CUDA(y11_launch_pdl(kern, dim3(L->grid), dim3(kThreadsDw), L->smem_bytes, s, L->tmap, L->p));
  return 0;
}
