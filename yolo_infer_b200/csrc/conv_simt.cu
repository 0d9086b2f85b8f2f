// CUDA-core kernels of the YOLO11 network that are HBM/L2-bandwidth bound and do not belong on tensor cores:
//   stem conv (cin=3, K=27), depthwise 3x3 (Detect cls tower, PSA positional encoding), SPPF max-pools,
//   nearest 2x upsample into a concat slice; plus a naive direct conv used ONLY as a bring-up cross-check of
//   the tcgen05 kernel (Y11_IMPL_SIMT_DEBUG).  Reference ops replaced: SURVEY.md section 8a rows a7, a9, a11.
#include <algorithm>
#include <cstdlib>

#include "ops.h"

using namespace y11;

// ------------------------------------------------------------------------------------------------
// naive direct conv (debug cross-check): one thread per (pixel, output channel)
// ------------------------------------------------------------------------------------------------
__global__ void conv_simt_kernel(y11_conv_desc d) {
  const int cout = d.out.c, cin = d.in.c;
  const size_t total = (size_t)d.B * d.Hout * d.Wout * cout;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = idx % cout;
  const size_t pix = idx / cout;
  const int ow = pix % d.Wout, oh = (pix / d.Wout) % d.Hout, n = pix / ((size_t)d.Wout * d.Hout);
  const int pad = d.k / 2;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w) + (size_t)co * d.k * d.k * cin;
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(d.in.ptr);
  float acc = 0.f;
  for (int kh = 0; kh < d.k; ++kh) {
    const int ih = oh * d.stride + kh - pad;
    if (ih < 0 || ih >= d.Hin) continue;
    for (int kw = 0; kw < d.k; ++kw) {
      const int iw = ow * d.stride + kw - pad;
      if (iw < 0 || iw >= d.Win) continue;
      const __nv_bfloat16* ip = in + (((size_t)n * d.Hin + ih) * d.Win + iw) * d.in.c_total + d.in.c_off;
      const __nv_bfloat16* wp = w + (kh * d.k + kw) * cin;
      for (int c = 0; c < cin; ++c) acc = fmaf(__bfloat162float(ip[c]), __bfloat162float(wp[c]), acc);
    }
  }
  acc += d.bias[co];
  const bool pre = d.res.ptr && d.res_mode == Y11_RES_PRE_UP2;
  if (pre) {
    const size_t rpix = ((size_t)n * (d.Hout >> 1) + (oh >> 1)) * (d.Wout >> 1) + (ow >> 1);
    acc += __bfloat162float(static_cast<const __nv_bfloat16*>(d.res.ptr)[rpix * d.res.c_total + d.res.c_off + co]);
  }
  if (d.act == Y11_ACT_SILU) acc = silu(acc);
  if (d.res.ptr && !pre) acc += __bfloat162float(static_cast<const __nv_bfloat16*>(d.res.ptr)[pix * d.res.c_total + d.res.c_off + co]);
  if (d.out_f32)
    static_cast<float*>(d.out.ptr)[pix * d.out.c_total + d.out.c_off + co] = acc;
  else
    static_cast<__nv_bfloat16*>(d.out.ptr)[pix * d.out.c_total + d.out.c_off + co] = __float2bfloat16_rn(acc);
}

int conv_simt_launch(const y11_conv_desc* d, cudaStream_t s) {
  const size_t total = (size_t)d->B * d->Hout * d->Wout * d->out.c;
  conv_simt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*d);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// stem: 3 -> COUT, 3x3 stride 2 pad 1, +bias, SiLU.  K = 27 is far too thin for a 128-row tcgen05 tile and the layer is
// HBM bound (reads the bf16 image once, writes COUT channels per pixel), so it runs on the warp-level tensor-core path.
// v4: one CTA = ROWS output rows x PXB pixels.  The 2*ROWS+1 input rows of the tile are staged ONCE with 16-byte copies
// (global -> shared, verbatim); there is no im2col pass: K is ordered k = kh*10 + j with j = 0 a zero-weight dummy and
// j = 1..9 the (kw, c) taps, so that the A-fragment of output pixel p is the 10 consecutive bf16 starting at the EVEN
// element 6p-4 of input row kh - every mma.sync A register is one aligned 32-bit shared load.  Weights sit in registers as
// B fragments; each warp owns 32 pixels of a row, stages its bf16 result in a private shared buffer and stores 16 bytes
// per lane.  (v1: CUDA cores, 0.70 ms for YOLO11s batch 64; v3: one row per CTA + im2col pass through smem, 0.38 ms.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816_s(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// U8 = true: the input rows are read straight from the uint8 BGR source frames (y11_stem_desc.images) and converted while
// they are staged - BGR -> RGB, x * (1/255), round to bf16: bit for bit what letterbox_kernel writes for a frame that needs
// neither resizing nor padding - so that for frames already at network resolution the letterbox launch and its 2.4 MB/image
// bf16 round trip through HBM disappear.
template <int COUT, bool U8>
__global__ void __launch_bounds__(160) stem_kernel(y11_stem_desc d, int PXB, int tiles_w, int ROWS) {
  constexpr int OP = COUT + 8;   // output staging pitch in bf16 (conflict-free 16-byte reads)
  constexpr int WP = 40;         // weight staging pitch
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int row_p = 6 * PXB + 16;                    // bf16 per staged input row: [8 left pad | 6*PXB | 8 slack]
  const int n_rows = 2 * ROWS + 1;
  // (U8: 16 bytes of front slack - the first 4-pixel group of row 0 starts 4 elements before the row)
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(s_raw) + (U8 ? 8 : 0);   // [n_rows][row_p]
  __nv_bfloat16* s_o = s_in + n_rows * row_p;                            // [PXB][OP] (also weight staging at start)
  const int row_blocks = d.Hout / ROWS;
  const int tile = blockIdx.x % tiles_w;
  const int oh0 = ((blockIdx.x / tiles_w) % row_blocks) * ROWS;
  const int n = blockIdx.x / (tiles_w * row_blocks);
  const int ow0 = tile * PXB;
  const int nt = blockDim.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  // weights [COUT][27] -> s_w[COUT][32] in the k = kh*10 + j order (j = 0 and k >= 30: zero) -> B fragments in registers
  {
    __nv_bfloat16* s_w = s_o;
    const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w);
    for (int i = tid; i < COUT * 32; i += nt) {
      const int co = i >> 5, k = i & 31, kh = k / 10, j = k - kh * 10;
      s_w[co * WP + k] = (kh < 3 && j >= 1) ? w[co * 27 + kh * 9 + j - 1] : zero;
    }
  }
  __syncthreads();
  uint32_t bw[COUT / 8][2][2];
  float bias[COUT / 8][2];
#pragma unroll
  for (int nb = 0; nb < COUT / 8; ++nb) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __nv_bfloat16* wp = s_o + (nb * 8 + g) * WP + ks * 16 + 2 * t;
      bw[nb][ks][0] = *reinterpret_cast<const uint32_t*>(wp);
      bw[nb][ks][1] = *reinterpret_cast<const uint32_t*>(wp + 8);
    }
    bias[nb][0] = __ldg(d.bias + nb * 8 + 2 * t);
    bias[nb][1] = __ldg(d.bias + nb * 8 + 2 * t + 1);
  }
  pdl_wait();   // weights/bias above are constants; the image below is the previous kernel's output
  pdl_trigger();
  __syncthreads();  // s_w (aliasing s_o) fully consumed
  // stage input rows 2*oh0-1 .. 2*oh0+2*ROWS-1: smem element 8 + e  <-  row element (6*ow0 + e), e in [-8, 6*PXB)
  if (U8) {
    // groups of 4 pixels = 12 source bytes (three aligned 32-bit loads) -> 12 bf16 (three 8-byte shared stores); group g of a
    // row holds row elements [6*ow0 - 12 + 12g, +12) and lands at smem element 12g - 4 of the row (the 4 elements in front of
    // a row are unused slack of the previous row / the front slack)
    const y11_image im = d.images[n];
    const int gpr = PXB / 2 + 1;                     // groups per row
    const int total = n_rows * gpr;
    const uint32_t gmagic = 0xffffffffu / (uint32_t)gpr + 1u;  // i / gpr == umulhi(i, gmagic) for the few thousand i of a tile
    const float r255 = 1.0f / 255.0f;
    const float c255 = -8388608.0f * r255;  // exact (a power-of-two multiple of r255)
    constexpr int kFly = 4;
    for (int i0 = tid; i0 < total; i0 += kFly * nt) {
      uint32_t w[kFly][3];
      bool ok[kFly];
#pragma unroll
      for (int q = 0; q < kFly; ++q) {
        const int i = min(i0 + q * nt, total - 1);
        const int r = (int)__umulhi((uint32_t)i, gmagic), gi = i - r * gpr;
        const int ih = 2 * oh0 - 1 + r;
        const int x = 2 * ow0 - 4 + 4 * gi;          // first pixel of the group (a multiple of 4: all in or all out)
        ok[q] = ih >= 0 && ih < d.Hin && x >= 0 && x < d.Win;
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(im.src + (size_t)min(max(ih, 0), d.Hin - 1) * im.pitch +
                                                               (size_t)min(max(x, 0), d.Win - 4) * 3);
        w[q][0] = __ldg(sp); w[q][1] = __ldg(sp + 1); w[q][2] = __ldg(sp + 2);
      }
#pragma unroll
      for (int q = 0; q < kFly; ++q) {
        const int i = i0 + q * nt;
        if (i < total) {
          const int r = (int)__umulhi((uint32_t)i, gmagic), gi = i - r * gpr;
          uint2 o[3] = {make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u)};
          if (ok[q]) {
            // source bytes B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3  ->  elements R0 G0 B0 R1 G1 B1 R2 G2 B2 R3 G3 B3
            // byte -> x/255 without the conversion unit (round 2: I2F.U8 shares the quarter-rate pipe with the SiLU tanh):
            // PRMT builds the float 2^23 + byte (byte in the low mantissa bits of 0x4B000000), one FFMA2 per pair computes
            // (2^23 + b) * r - 2^23 * r with a single rounding = RN(b * r), the same bits as __fmul_rn((float)b, r)
            f32x2 f2[6];
#pragma unroll
            for (int e = 0; e < 12; e += 2) {
              const int sb0 = 3 * (e / 3) + 2 - (e % 3), sb1 = 3 * ((e + 1) / 3) + 2 - ((e + 1) % 3);  // source bytes (compile-time)
              const uint32_t v0 = __byte_perm(w[q][sb0 >> 2], 0x4B000000u, 0x7540u | (uint32_t)(sb0 & 3));
              const uint32_t v1 = __byte_perm(w[q][sb1 >> 2], 0x4B000000u, 0x7540u | (uint32_t)(sb1 & 3));
              f2[e / 2] = f2_fma(f2_pack(__uint_as_float(v0), __uint_as_float(v1)), f2_pack(r255, r255), f2_pack(c255, c255));
            }
#pragma unroll
            for (int v = 0; v < 3; ++v) o[v] = make_uint2(f2_to_bf16x2(f2[2 * v]), f2_to_bf16x2(f2[2 * v + 1]));
          }
          uint2* dst = reinterpret_cast<uint2*>(s_in + r * row_p + 12 * gi - 4);
          dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
        }
      }
    }
  } else {
    const int vec_per_row = (6 * PXB) / 8 + 1;       // one leading vector (elements -8..-1) + the segment itself
    const int total = n_rows * vec_per_row;
    const uint32_t vmagic = 0xffffffffu / (uint32_t)vec_per_row + 1u;
    const int row_e = 3 * d.Win;                     // bf16 per image row (multiple of 8)
    const __nv_bfloat16* img = static_cast<const __nv_bfloat16*>(d.in) + (size_t)n * d.Hin * row_e;
    // unconditional loads from clamped addresses, eight in flight per thread, zeroed afterwards (see dwconv_kernel); with four
    // the first use of the loaded vectors was the kernel's top stall site (16 % of the samples, ncu source view)
    constexpr int kFly = 8;
    for (int i0 = tid; i0 < total; i0 += kFly * nt) {
      uint4 u[kFly];
      bool ok[kFly];
#pragma unroll
      for (int q = 0; q < kFly; ++q) {
        const int i = min(i0 + q * nt, total - 1);
        const int r = (int)__umulhi((uint32_t)i, vmagic), v = i - r * vec_per_row;
        const int ih = 2 * oh0 - 1 + r;
        const int e0 = 6 * ow0 - 8 + 8 * v;          // first row element of this vector (a multiple of 8)
        ok[q] = ih >= 0 && ih < d.Hin && e0 >= 0 && e0 < row_e;  // a vector is entirely inside or outside the row
        u[q] = ldg_nc_v4(img + (size_t)min(max(ih, 0), d.Hin - 1) * row_e + min(max(e0, 0), row_e - 8));
      }
#pragma unroll
      for (int q = 0; q < kFly; ++q) {
        const int i = i0 + q * nt;
        if (i < total) {
          const int r = (int)__umulhi((uint32_t)i, vmagic), v = i - r * vec_per_row;
          *reinterpret_cast<uint4*>(s_in + r * row_p + 8 * v) = ok[q] ? u[q] : make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  __syncthreads();
  if (warp * 32 >= PXB) return;
  // per-lane offsets (bf16 elements, relative to pixel p's window start = smem element 4 + 6p of row kh=0) of the four
  // k-pairs this lane feeds into the two K=16 steps; k >= 30 has zero weights and re-reads k = 28 (finite data)
  int koff[2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      int k = ks * 16 + hh * 8 + 2 * t;
      if (k >= 30) k = 28;
      const int kh = k / 10;
      koff[ks][hh] = kh * row_p + (k - kh * 10) + 4;
    }
  constexpr int VPP = COUT / 8;  // 16-byte vectors per pixel
  __nv_bfloat16* so = s_o + warp * 32 * OP;
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(d.out.ptr) + d.out.c_off;
  // Write-out plan of this lane, constant over the rows of the tile (round 2: the per-store index arithmetic - pixel / vector
  // split, block selection, 64-bit multiplies - was 17 instructions per 16-byte store).  Vector j of the lane is 16-byte unit
  // v = lane % VPP of pixel lane / VPP + j * (32 / VPP) of the warp's 32 (VPP divides 32 for 16 / 32 / 64 output channels; 96
  // keeps the generic loop): its shared and global offsets are those of vector 0 plus j times a constant step, and the global
  // offset is row_base(oy) + (oy odd ? goff1 : goff0) + j * gstep.
  // space-to-depth: pixel (oy/2, ox/2) of the half-size map, channel block dy*2 + dx, or (s2d == 2) the permuted order
  // [(1,0), (1,1), (0,1), (0,0)]: dy ? dx : 3 - dx; the pixel step 32 / VPP is even, so dx is the same for all j
  constexpr bool kPlan = 32 % VPP == 0;
  constexpr int PSTEP = kPlan ? 32 / VPP : 1;
  const int px0 = lane / VPP, v0 = lane % VPP, ox0 = ow0 + warp * 32 + px0, dx0 = ox0 & 1;
  const uint32_t soff0 = (uint32_t)(px0 * OP + v0 * 8);
  int goff0, goff1, gstep;
  if (d.s2d) {
    const int b0 = d.s2d == 2 ? 3 - dx0 : dx0, b1 = d.s2d == 2 ? dx0 : 2 + dx0;
    goff0 = (ox0 >> 1) * d.out.c_total + b0 * COUT + v0 * 8;
    goff1 = (ox0 >> 1) * d.out.c_total + b1 * COUT + v0 * 8;
    gstep = (PSTEP / 2) * d.out.c_total;
  } else {
    goff0 = goff1 = ox0 * d.out.c_total + v0 * 8;
    gstep = PSTEP * d.out.c_total;
  }
  for (int r = 0; r < ROWS; ++r) {
    const __nv_bfloat16* base = s_in + 2 * r * row_p;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int p0 = warp * 32 + mt * 16 + g;
      uint32_t af[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        af[ks][0] = *reinterpret_cast<const uint32_t*>(base + 6 * p0 + koff[ks][0]);
        af[ks][1] = *reinterpret_cast<const uint32_t*>(base + 6 * (p0 + 8) + koff[ks][0]);
        af[ks][2] = *reinterpret_cast<const uint32_t*>(base + 6 * p0 + koff[ks][1]);
        af[ks][3] = *reinterpret_cast<const uint32_t*>(base + 6 * (p0 + 8) + koff[ks][1]);
      }
#pragma unroll
      for (int nb = 0; nb < COUT / 8; ++nb) {
        float c[4] = {bias[nb][0], bias[nb][1], bias[nb][0], bias[nb][1]};
        mma_bf16_16816_s(c, af[0], bw[nb][0][0], bw[nb][0][1]);
        mma_bf16_16816_s(c, af[1], bw[nb][1][0], bw[nb][1][1]);
        *reinterpret_cast<uint32_t*>(so + (mt * 16 + g) * OP + nb * 8 + 2 * t) = f2_to_bf16x2(silu2(f2_pack(c[0], c[1])));
        *reinterpret_cast<uint32_t*>(so + (mt * 16 + g + 8) * OP + nb * 8 + 2 * t) = f2_to_bf16x2(silu2(f2_pack(c[2], c[3])));
      }
    }
    __syncwarp();
    // coalesced write-out of this warp's 32 pixels: 16 bytes per lane, consecutive lanes -> consecutive bytes of a pixel row
    const int oy = oh0 + r;
    const size_t row0 = d.s2d ? ((size_t)n * (d.Hout >> 1) + (oy >> 1)) * (d.Wout >> 1) : ((size_t)n * d.Hout + oy) * d.Wout;
    __nv_bfloat16* orow = ob + row0 * d.out.c_total;
    const int dy = oy & 1;
    if (kPlan) {
      __nv_bfloat16* og = orow + (dy ? goff1 : goff0);
#pragma unroll
      for (int j = 0; j < VPP; ++j)
        if (ox0 + j * PSTEP < d.Wout) *reinterpret_cast<uint4*>(og + j * gstep) = *reinterpret_cast<const uint4*>(so + soff0 + j * (PSTEP * OP));
    } else {
#pragma unroll
      for (int i = lane; i < 32 * VPP; i += 32) {
        const int px = i / VPP, v = i % VPP;
        const int ox = ow0 + warp * 32 + px;
        if (ox < d.Wout) {
          const int dx = ox & 1;
          const int blk = d.s2d == 2 ? (dy ? dx : 3 - dx) : dy * 2 + dx;
          const size_t off = d.s2d ? (size_t)(ox >> 1) * d.out.c_total + blk * COUT : (size_t)ox * d.out.c_total;
          *reinterpret_cast<uint4*>(orow + off + v * 8) = *reinterpret_cast<const uint4*>(so + px * OP + v * 8);
        }
      }
    }
    __syncwarp();
  }
}

template <int COUT, bool U8>
static int stem_launch_t(const y11_stem_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->Win % 8 == 0 && d->Hin % 2 == 0, "stem: input must be even-sized with width a multiple of 8 (got %dx%d)", d->Hin, d->Win);
  Y11_REQUIRE(d->Wout * 2 == d->Win && d->Hout * 2 == d->Hin, "stem: output must be half the input size");
  Y11_REQUIRE(!d->s2d || (d->Hout % 2 == 0 && d->Wout % 2 == 0 && d->out.c == 4 * COUT),
              "stem: space-to-depth output needs even Hout/Wout and a 4*cout-channel view");
  int PXB = 32, best_waste = 1 << 30;
  for (int c : {160, 128, 96, 64, 32}) {  // fewest wasted pixels in the last tile, then the widest tile
    const int waste = y11_ceil_div(d->Wout, c) * c - d->Wout;
    if (waste < best_waste) { best_waste = waste; PXB = c; }
  }
  int ROWS = 1, max_rows = 8;
  if (const char* e = getenv("Y11_STEM_ROWS")) max_rows = std::max(1, atoi(e));
  for (int r : {8, 4, 2})
    if (r <= max_rows && d->Hout % r == 0) { ROWS = r; break; }
  const int tiles_w = y11_ceil_div(d->Wout, PXB);
  const int row_p = 6 * PXB + 16;
  const size_t smem = (size_t)((2 * ROWS + 1) * row_p + std::max(PXB * (COUT + 8), COUT * 40)) * 2 + (U8 ? 16 : 0);
  Y11_OPT_IN_SMEM((stem_kernel<COUT, U8>), 96 * 1024);
  Y11_CHECK_CUDA(y11_launch_pdl(stem_kernel<COUT, U8>, dim3((unsigned)(tiles_w * (d->Hout / ROWS) * d->B)), dim3(PXB), smem, s, *d,
                                PXB, tiles_w, ROWS));
  return 0;
}

int stem_launch(const y11_stem_desc* d, cudaStream_t s) {
  const int cout = d->s2d ? d->out.c / 4 : d->out.c;
  if (d->u8_src) {
    Y11_REQUIRE(d->images, "stem: u8_src without image descriptors");
    switch (cout) {
      case 16: return stem_launch_t<16, true>(d, s);
      case 32: return stem_launch_t<32, true>(d, s);
      case 64: return stem_launch_t<64, true>(d, s);
      case 96: return stem_launch_t<96, true>(d, s);
    }
  } else {
    switch (cout) {
      case 16: return stem_launch_t<16, false>(d, s);
      case 32: return stem_launch_t<32, false>(d, s);
      case 64: return stem_launch_t<64, false>(d, s);
      case 96: return stem_launch_t<96, false>(d, s);
    }
  }
  y11_set_error("stem: unsupported cout %d", d->out.c);
  return -1;
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 stride 1 pad 1.  v2: one thread = 4 channels x one column x a vertical strip of R output rows.
// The 9x4 weights live in registers as fp32; every input row of the strip is loaded once (3 x 8-byte loads: x-1, x, x+1)
// and feeds up to three output rows, so the load-store unit sees 3(R+2)/R + 1 accesses per output instead of the
// 9 inputs + 9 weights + 1 store of v1 (one thread per pixel x 8 channels, 1.4 TB/s: L1-wavefront bound).
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) dwconv_kernel(y11_dwconv_desc d, int strips) {
  const int groups = d.in.c / 4;
  const size_t total = (size_t)d.B * strips * d.W * groups;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  size_t rest = idx / groups;
  const int x = rest % d.W; rest /= d.W;
  const int y0 = (int)(rest % strips) * R;
  const size_t n = rest / strips;
  // constants first (they do not depend on the previous kernel): weights [9][c] tap-major, bias
  // (round 2: the channel pairs are packed fp32 pairs - FFMA2 / FMUL2 / FADD2, half the arithmetic instructions, same bits)
  f32x2 w[9][2], acc[R][2];
  {
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(d.w) + g * 4;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(wp + (size_t)tap * d.in.c));
      w[tap][0] = f2_from_bf16x2(u.x);
      w[tap][1] = f2_from_bf16x2(u.y);
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(d.bias + g * 4));
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][0] = f2_pack(b.x, b.y); acc[r][1] = f2_pack(b.z, b.w); }
  }
  pdl_wait();
  pdl_trigger();
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(d.in.ptr) + d.in.c_off + g * 4;
  const size_t img = n * d.H;
  // All 3(R+2) loads are issued unconditionally from CLAMPED addresses and zeroed afterwards: guarded loads would be
  // serialised row by row (the compiler cannot hoist a load above the branch that protects it), which made v1 and the
  // first strip version latency bound at 1.4 TB/s.
  const int xl = x > 0 ? -1 : 0, xr = x + 1 < d.W ? 1 : 0;
  uint2 v[R + 2][3];
#pragma unroll
  for (int r = 0; r < R + 2; ++r) {
    const int iy = min(max(y0 + r - 1, 0), d.H - 1);
    const __nv_bfloat16* rp = in + ((img + iy) * d.W + x) * d.in.c_total;
    v[r][0] = *reinterpret_cast<const uint2*>(rp + xl * d.in.c_total);
    v[r][1] = *reinterpret_cast<const uint2*>(rp);
    v[r][2] = *reinterpret_cast<const uint2*>(rp + xr * d.in.c_total);
  }
#pragma unroll
  for (int r = 0; r < R + 2; ++r) {
    const int iy = y0 + r - 1;
    const bool row_ok = iy >= 0 && iy < d.H;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const bool ok = row_ok && (kw == 1 || (kw == 0 ? xl != 0 : xr != 0));
      const f32x2 f01 = f2_from_bf16x2(ok ? v[r][kw].x : 0u), f23 = f2_from_bf16x2(ok ? v[r][kw].y : 0u);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int orow = r - kh;  // input row y0+r-1 is tap kh of output row y0+r-kh
        if (orow >= 0 && orow < R) {
          acc[orow][0] = f2_fma(f01, w[kh * 3 + kw][0], acc[orow][0]);
          acc[orow][1] = f2_fma(f23, w[kh * 3 + kw][1], acc[orow][1]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int y = y0 + r;
    if (y >= d.H) break;
    const size_t pix = (img + y) * d.W + x;
    f32x2 o0 = acc[r][0], o1 = acc[r][1];
    if (d.act == Y11_ACT_SILU) { o0 = silu2(o0); o1 = silu2(o1); }
    if (d.res.ptr) {
      const uint2 rr = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(d.res.ptr) + pix * d.res.c_total + d.res.c_off + g * 4);
      o0 = f2_add(o0, f2_from_bf16x2(rr.x));
      o1 = f2_add(o1, f2_from_bf16x2(rr.y));
    }
    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(d.out.ptr) + pix * d.out.c_total + d.out.c_off + g * 4) =
        make_uint2(f2_to_bf16x2(o0), f2_to_bf16x2(o1));
  }
}

int dwconv_launch(const y11_dwconv_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->in.c % 8 == 0 && d->in.c_off % 8 == 0 && d->in.c_total % 8 == 0 && d->out.c_off % 8 == 0 &&
                  d->out.c_total % 8 == 0,
              "dwconv: views must be 16-byte aligned (c=%d)", d->in.c);
  Y11_REQUIRE(!d->res.ptr || (d->res.c_off % 4 == 0 && d->res.c_total % 4 == 0), "dwconv: residual view alignment");
  constexpr int R = 4;
  const int strips = y11_ceil_div(d->H, R);
  const size_t total = (size_t)d->B * strips * d->W * (d->in.c / 4);
  Y11_CHECK_CUDA(y11_launch_pdl(dwconv_kernel<R>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, *d, strips));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SPPF: y1 = mp5(y0), y2 = mp5(y1), y3 = mp5(y2) in ONE kernel.  One CTA = one image x 8 channels; the
// H x W x 8 plane lives in shared memory and the three chained 5x5 pools run as separable row/column max.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* rp = reinterpret_cast<__nv_bfloat162*>(&r);
  const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) rp[i] = __hmax2(ap[i], bp[i]);
  return r;
}

__global__ void __launch_bounds__(256) sppf_kernel(y11_sppf_desc d) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint4 s_plane[];  // [2][H*W]
  const int hw = d.H * d.W;
  uint4* cur = s_plane;
  uint4* tmp = s_plane + hw;
  const int groups = d.c / 8;
  const int g = blockIdx.x % groups, n = blockIdx.x / groups;
  __nv_bfloat16* base = static_cast<__nv_bfloat16*>(d.io.ptr) + (size_t)n * hw * d.io.c_total + d.io.c_off + g * 8;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) cur[i] = *reinterpret_cast<const uint4*>(base + (size_t)i * d.io.c_total);
  __syncthreads();
  for (int stage = 1; stage <= 3; ++stage) {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {  // row pass
      const int x = i % d.W, row = i - x;
      uint4 m = cur[i];
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (dx != 0 && xx >= 0 && xx < d.W) m = bf16x8_max(m, cur[row + xx]);
      }
      tmp[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {  // column pass
      const int x = i % d.W, y = i / d.W;
      uint4 m = tmp[i];
#pragma unroll
      for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (dy != 0 && yy >= 0 && yy < d.H) m = bf16x8_max(m, tmp[yy * d.W + x]);
      }
      cur[i] = m;
      *reinterpret_cast<uint4*>(base + (size_t)i * d.io.c_total + stage * d.c) = m;
    }
    __syncthreads();
  }
}

int sppf_launch(const y11_sppf_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->c % 8 == 0 && d->io.c_off % 8 == 0 && d->io.c_total % 8 == 0, "sppf: alignment");
  const size_t smem = (size_t)2 * d->H * d->W * sizeof(uint4);
  Y11_REQUIRE(smem <= 200 * 1024, "sppf: plane %dx%d too large for shared memory", d->H, d->W);
  Y11_OPT_IN_SMEM(sppf_kernel, 200 * 1024);
  Y11_CHECK_CUDA(y11_launch_pdl(sppf_kernel, dim3((unsigned)(d->B * (d->c / 8))), dim3(256), smem, s, *d));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// nearest 2x upsample into a channel slice of the concat buffer (16-byte vectors)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(y11_upsample_desc d) {
  pdl_wait();
  pdl_trigger();
  const int groups = d.in.c / 8;
  const int Ho = 2 * d.H, Wo = 2 * d.W;
  const size_t total = (size_t)d.B * Ho * Wo * groups;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  const size_t opix = idx / groups;
  const int ox = opix % Wo, oy = (opix / Wo) % Ho;
  const size_t n = opix / ((size_t)Wo * Ho);
  const size_t ipix = (n * d.H + oy / 2) * d.W + ox / 2;
  const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(d.in.ptr) + ipix * d.in.c_total + d.in.c_off + g * 8);
  *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(d.out.ptr) + opix * d.out.c_total + d.out.c_off + g * 8) = v;
}

int upsample_launch(const y11_upsample_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->in.c % 8 == 0 && d->in.c_off % 8 == 0 && d->in.c_total % 8 == 0 && d->out.c_off % 8 == 0 &&
                  d->out.c_total % 8 == 0, "upsample: alignment");
  const size_t total = (size_t)d->B * 4 * d->H * d->W * (d->in.c / 8);
  Y11_CHECK_CUDA(y11_launch_pdl(upsample_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, *d));
  return 0;
}
