// Depthwise 3x3 (stride 1, pad 1) for the large feature maps: persistent CTAs, TMA-fed shared-memory ring.
//
// Replaces ultralytics DWConv.forward_fuse (Detect class tower) and Attention.pe.  The
// register-only kernel in conv_simt.cu is latency bound: the bytes a thread can keep in flight are limited by its
// registers (18 loads x 8 B at 95 registers/thread = 74 KB per SM, ~2.7 us per warp batch -> 1.7 TB/s).  Here the bytes in
// flight are whole halo tiles owned by the TMA engine:
//   tile    = 8 rows x 16 pixels x CC channels of one image (CC <= 128, a multiple of 8); its (8+2) x (16+2) halo is ONE
//             4-D TMA box load (rows of CC*2 >= 128 bytes; outside the image the TMA unit writes zeros == conv padding);
//   ring    = 3 stages (up to 46 KB each), one producer warp runs two tiles ahead;
//   compute = 512 threads, thread = 4 channels x 1 column: the 9x4 weights live in registers as fp32, the ten halo rows
//             slide through three row accumulators (3 LDS.64 + 36 FMA per halo row), bias/SiLU/residual fused,
//             one warp writes the 32 x 8 B = 256 contiguous bytes of a pixel.
#include <algorithm>
#include <cstring>

#include "ops.h"

namespace {

using namespace y11;

constexpr int kTH = 8, kTW = 16;
constexpr int kHaloH = kTH + 2, kHaloW = kTW + 2;
constexpr int kComputeThreads = 512;
constexpr int kThreadsDw = kComputeThreads + 32;
constexpr int kMaxStagesDw = 6;  // ring depth is per launch: as many halo tiles as fit in ~190 KB, at most 6 (v1: fixed 3)

// One thread's share of a tile: 4 channels x 1 column x RPS output rows starting at tile row r0 (halo rows r0 .. r0+RPS+1).
template <int RPS>
__device__ __forceinline__ void dw_rows(const DwTmaParams& p, uint32_t base, uint32_t row_pitch, uint32_t px_pitch, int r0,
                                        const float (&w)[9][4], const float (&b)[4], int n, int x, int y0, int c0) {
  float acc[3][4];  // three output rows are in flight: row o lives in acc[o % 3]
#pragma unroll
  for (int j = 0; j < 3; ++j) { acc[j][0] = b[0]; acc[j][1] = b[1]; acc[j][2] = b[2]; acc[j][3] = b[3]; }
#pragma unroll
  for (int r = 0; r < RPS + 2; ++r) {
    // halo row r0+r holds input row y0+r0+r-1: tap row kh of local output row r - kh
    float f[3][4];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      uint32_t lo, hi;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                   : "=r"(lo), "=r"(hi)
                   : "r"(base + (uint32_t)(r0 + r) * row_pitch + (uint32_t)kw * px_pitch));
      f[kw][0] = bf16_lo(lo); f[kw][1] = bf16_hi(lo); f[kw][2] = bf16_lo(hi); f[kw][3] = bf16_hi(hi);
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int orow = r - kh;
      if (orow < 0 || orow >= RPS) continue;
      float* a = acc[orow % 3];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = fmaf(f[kw][i], w[kh * 3 + kw][i], a[i]);
      }
    }
    const int done = r - 2;  // local output row `done` has now received its three tap rows
    if (done >= 0) {
      float* a = acc[done % 3];
      const int y = y0 + r0 + done;
      if (y < p.H && x < p.W) {
        float o[4] = {a[0], a[1], a[2], a[3]};
        if (p.act == Y11_ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = silu(o[i]);
        }
        const size_t pix = ((size_t)n * p.H + y) * p.W + x;
        if (p.res) {
          const uint2 rr = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.res) + pix * p.res_ct + p.res_co + c0);
          o[0] += bf16_lo(rr.x); o[1] += bf16_hi(rr.x); o[2] += bf16_lo(rr.y); o[3] += bf16_hi(rr.y);
        }
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.out_ct + p.out_co + c0) =
            make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      }
      a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; a[3] = b[3];
    }
  }
}

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
  // thread -> (4-channel group, pixel column, row group): with CC < 128 the spare threads split the 8 rows between them
  const int cg = threadIdx.x % p.cg4;
  const int col = (threadIdx.x / p.cg4) % kTW;
  const int sub = threadIdx.x / (p.cg4 * kTW);
  const bool worker = warp < kComputeThreads / 32 && sub < p.nsub;
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const int total = p.tiles_w * p.tiles_h * p.B * p.chunks;

  if (warp == kComputeThreads / 32) {
    // ------------------------------------------------------------------ producer
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int t = tile;
      const int ck = t % p.chunks; t /= p.chunks;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      mbar_wait(empty_bar + 8 * stage, phase ^ 1, p.err_flag, 201);
      if (elect_one()) {
        mbar_expect_tx(full_bar + 8 * stage, p.stage_tx);
        tma_load_4d(tiles_base + stage * p.stage_bytes, &tmap, full_bar + 8 * stage, ck * p.CC, tw * kTW - 1, th * kTH - 1, t);
      }
      __syncwarp();
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  uint32_t stage = 0, phase = 0;
  const uint32_t row_pitch = (uint32_t)kHaloW * p.CC * 2u, px_pitch = (uint32_t)p.CC * 2u;
  float w[9][4], b[4];
  int w_ck = -1;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int t = tile;
    const int ck = t % p.chunks; t /= p.chunks;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int n = t;
    const int c0 = ck * p.CC + cg * 4;
    if (worker && ck != w_ck) {  // this thread's 9x4 weights and bias (constant while the channel chunk does not change)
      const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(p.w) + c0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(wp + (size_t)tap * p.C));
        w[tap][0] = bf16_lo(u.x); w[tap][1] = bf16_hi(u.x); w[tap][2] = bf16_lo(u.y); w[tap][3] = bf16_hi(u.y);
      }
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
      b[0] = bb.x; b[1] = bb.y; b[2] = bb.z; b[3] = bb.w;
      w_ck = ck;
    }
    mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 202);
    if (worker) {
      const uint32_t base = tiles_base + stage * p.stage_bytes + (uint32_t)col * px_pitch + (uint32_t)cg * 8u;
      const int x = tw * kTW + col, y0 = th * kTH;
      if (p.nsub == 1) dw_rows<8>(p, base, row_pitch, px_pitch, 0, w, b, n, x, y0, c0);
      else if (p.nsub == 2) dw_rows<4>(p, base, row_pitch, px_pitch, sub * 4, w, b, n, x, y0, c0);
      else dw_rows<2>(p, base, row_pitch, px_pitch, sub * 2, w, b, n, x, y0, c0);
    }
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
  if (d->W < 32 || d->H < 16) return false;  // ragged right/bottom tiles are fine: TMA zero-fills, stores are guarded
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
  Y11_REQUIRE(p.cg4 * kTW <= kComputeThreads, "dwconv_tma: chunk of %d channels needs more than %d threads", p.CC, kComputeThreads);
  p.nsub = kComputeThreads / (p.cg4 * kTW);  // row groups: 1 (8 rows per thread), 2 (4 rows) or 4 (2 rows)
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
  L->smem_bytes = 128u + (unsigned)p.stages * p.stage_bytes;
  Y11_OPT_IN_SMEM(dwconv_tma_kernel, 200 * 1024);
  return 0;
}

int dwconv_tma_launch(const DwTmaLaunch* L, cudaStream_t s) {
  Y11_CHECK_CUDA(y11_launch_pdl(dwconv_tma_kernel, dim3(L->grid), dim3(kThreadsDw), L->smem_bytes, s, L->tmap, L->p));
  return 0;
}
