// CUDA-core kernels of the YOLO11 network that are HBM/L2-bandwidth bound and do not belong on tensor cores:
//   stem conv (cin=3, K=27), depthwise 3x3 (Detect cls tower, PSA positional encoding), SPPF max-pools,
//   nearest 2x upsample into a concat slice; plus a naive direct conv used ONLY as a bring-up cross-check of
//   the tcgen05 kernel (Y11_IMPL_SIMT_DEBUG).  Reference ops replaced: SURVEY.md section 8a rows a7, a9, a11.
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
  if (d.act == Y11_ACT_SILU) acc = silu(acc);
  if (d.res.ptr) acc += __bfloat162float(static_cast<const __nv_bfloat16*>(d.res.ptr)[pix * d.res.c_total + d.res.c_off + co]);
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
// stem: 3 -> COUT, 3x3 stride 2 pad 1, +bias, SiLU.  K = 27 (padded to 32) is far too thin for a 128-row tcgen05 tile
// and the layer is HBM bound (reads the bf16 image once, writes COUT channels per pixel), so it runs on the warp-level
// tensor-core path: per CTA, three input rows are staged with 16-byte loads, each thread builds the im2col row of one
// output pixel in shared memory ([pixel][32] bf16, 80-byte pitch = conflict-free fragment loads), every warp does
// 32 pixels x COUT with mma.sync.m16n8k16 (weights held in registers as B fragments), and the bf16 result goes back
// through shared memory so that each lane stores 16 contiguous bytes.
// (v1: CUDA cores, 1 pixel/thread, LDS-issue bound, 0.70 ms for YOLO11s batch 64; v3: register-blocked, 0.91 ms.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816_s(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int COUT>
__global__ void __launch_bounds__(160) stem_kernel(y11_stem_desc d, int PXB, int tiles_w) {
  constexpr int AP = 40;         // im2col / weight row pitch in bf16 (32 + 8 pad)
  constexpr int OP = COUT + 8;   // output staging pitch in bf16
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int row_e = (2 * PXB + 1) * 3;               // bf16 elements per staged input row
  const int row_p = (row_e + 7) & ~7;
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(s_raw);          // [3][row_p]
  __nv_bfloat16* s_a = s_in + 3 * row_p;                                 // [PXB][AP]
  __nv_bfloat16* s_w = s_a + PXB * AP;                                   // [COUT][AP]
  __nv_bfloat16* s_o = s_w + COUT * AP;                                  // [PXB][OP]
  float* s_b = reinterpret_cast<float*>(s_o + PXB * OP);                 // [COUT]
  const int tile = blockIdx.x % tiles_w;
  const int oh = (blockIdx.x / tiles_w) % d.Hout;
  const int n = blockIdx.x / (tiles_w * d.Hout);
  const int ow0 = tile * PXB;
  const int nt = blockDim.x, tid = threadIdx.x;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  // weights [COUT][27] -> [COUT][AP] (k >= 27 zero), bias
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w);
  for (int i = tid; i < COUT * AP; i += nt) {
    const int co = i / AP, k = i % AP;
    s_w[i] = k < 27 ? w[co * 27 + k] : zero;
  }
  for (int i = tid; i < COUT; i += nt) s_b[i] = d.bias[i];
  for (int i = tid; i < 3 * row_p; i += nt) s_in[i] = zero;
  pdl_wait();   // weights/bias above are constants; the image below is the previous kernel's output
  pdl_trigger();
  __syncthreads();
  // stage the three input rows (16-byte global loads; rows are 16-byte aligned because Win % 8 == 0)
  const int iw0 = 2 * ow0 - 1;
  const int e_lo = max(iw0, 0) * 3, e_hi = min(iw0 + 2 * PXB + 1, d.Win) * 3;
  const int v_lo = e_lo / 8, v_hi = (e_hi + 7) / 8;
  for (int r = 0; r < 3; ++r) {
    const int ih = 2 * oh + r - 1;
    if (ih < 0 || ih >= d.Hin) continue;
    const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(d.in) + ((size_t)n * d.Hin + ih) * d.Win * 3);
    __nv_bfloat16* dst = s_in + r * row_p - iw0 * 3;
    for (int v = v_lo + tid; v < v_hi; v += nt) {
      const uint4 u = __ldg(rp + v);
      const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int e = v * 8 + t;
        if (e >= e_lo && e < e_hi) dst[e] = h[t];
      }
    }
  }
  __syncthreads();
  // im2col: thread p builds A[p][0..31], k = kh*9 + kw*3 + c  <-  in[kh][(2p+kw)*3 + c] = in[kh][6p + (k - 9kh)]
  if (tid < PXB) {
    __nv_bfloat16* ar = s_a + tid * AP;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const __nv_bfloat16* src = s_in + kh * row_p + 6 * tid;
#pragma unroll
      for (int j = 0; j < 9; ++j) ar[kh * 9 + j] = src[j];
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) ar[k] = zero;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  if (warp * 32 >= PXB) return;
  // B fragments (weights) in registers: B[k][n] = W[n][k]
  uint32_t bw[COUT / 8][2][2];
#pragma unroll
  for (int nb = 0; nb < COUT / 8; ++nb)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __nv_bfloat16* wp = s_w + (nb * 8 + g) * AP + ks * 16 + 2 * t;
      bw[nb][ks][0] = *reinterpret_cast<const uint32_t*>(wp);
      bw[nb][ks][1] = *reinterpret_cast<const uint32_t*>(wp + 8);
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int p0 = warp * 32 + mt * 16;
    uint32_t af[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __nv_bfloat16* ap = s_a + (p0 + g) * AP + ks * 16 + 2 * t;
      af[ks][0] = *reinterpret_cast<const uint32_t*>(ap);
      af[ks][1] = *reinterpret_cast<const uint32_t*>(ap + 8 * AP);
      af[ks][2] = *reinterpret_cast<const uint32_t*>(ap + 8);
      af[ks][3] = *reinterpret_cast<const uint32_t*>(ap + 8 * AP + 8);
    }
#pragma unroll
    for (int nb = 0; nb < COUT / 8; ++nb) {
      float c[4];
      const float b0 = s_b[nb * 8 + 2 * t], b1 = s_b[nb * 8 + 2 * t + 1];
      c[0] = b0; c[1] = b1; c[2] = b0; c[3] = b1;
      mma_bf16_16816_s(c, af[0], bw[nb][0][0], bw[nb][0][1]);
      mma_bf16_16816_s(c, af[1], bw[nb][1][0], bw[nb][1][1]);
      *reinterpret_cast<uint32_t*>(s_o + (p0 + g) * OP + nb * 8 + 2 * t) = pack_bf16x2(silu(c[0]), silu(c[1]));
      *reinterpret_cast<uint32_t*>(s_o + (p0 + g + 8) * OP + nb * 8 + 2 * t) = pack_bf16x2(silu(c[2]), silu(c[3]));
    }
  }
  __syncwarp();
  // coalesced write-out of this warp's 32 pixels: 16 bytes per lane, consecutive lanes -> consecutive bytes of a pixel row
  constexpr int VPP = COUT / 8;  // 16-byte vectors per pixel
  const size_t pix0 = ((size_t)n * d.Hout + oh) * d.Wout + ow0 + warp * 32;
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(d.out.ptr) + d.out.c_off;
  for (int i = lane; i < 32 * VPP; i += 32) {
    const int px = i / VPP, v = i % VPP;
    if (ow0 + warp * 32 + px < d.Wout)
      *reinterpret_cast<uint4*>(ob + (pix0 + px) * d.out.c_total + v * 8) = *reinterpret_cast<const uint4*>(s_o + (warp * 32 + px) * OP + v * 8);
  }
}

template <int COUT>
static int stem_launch_t(const y11_stem_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->Win % 8 == 0, "stem: input width must be a multiple of 8 (got %d)", d->Win);
  int PXB = 128;
  for (int c : {160, 128, 96, 64, 32})
    if (d->Wout % c == 0) { PXB = c; break; }
  const int tiles_w = y11_ceil_div(d->Wout, PXB);
  const int row_p = (((2 * PXB + 1) * 3) + 7) & ~7;
  const size_t smem = (size_t)(3 * row_p + PXB * 40 + COUT * 40 + PXB * (COUT + 8)) * 2 + COUT * 4;
  static bool attr_set = false;
  if (!attr_set) {
    Y11_CHECK_CUDA(cudaFuncSetAttribute(stem_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  Y11_CHECK_CUDA(y11_launch_pdl(stem_kernel<COUT>, dim3((unsigned)(tiles_w * d->Hout * d->B)), dim3(PXB), smem, s, *d, PXB, tiles_w));
  return 0;
}

int stem_launch(const y11_stem_desc* d, cudaStream_t s) {
  switch (d->out.c) {
    case 16: return stem_launch_t<16>(d, s);
    case 32: return stem_launch_t<32>(d, s);
    case 64: return stem_launch_t<64>(d, s);
    case 96: return stem_launch_t<96>(d, s);
    default: y11_set_error("stem: unsupported cout %d", d->out.c); return -1;
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 stride 1 pad 1: one thread = one pixel x 8 channels (16-byte vectors)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv_kernel(y11_dwconv_desc d) {
  pdl_wait();
  pdl_trigger();
  const int groups = d.in.c / 8;
  const size_t total = (size_t)d.B * d.H * d.W * groups;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  const size_t pix = idx / groups;
  const int x = pix % d.W, y = (pix / d.W) % d.H;
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(d.in.ptr) + d.in.c_off + g * 8;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w) + g * 8;
  float acc[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(d.bias + g * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(d.bias + g * 8) + 1);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int iy = y + kh - 1;
    if (iy < 0 || iy >= d.H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ix = x + kw - 1;
      if (ix < 0 || ix >= d.W) continue;
      const size_t ipix = pix + (size_t)(kh - 1) * d.W + (kw - 1);
      const uint4 v = *reinterpret_cast<const uint4*>(in + ipix * d.in.c_total);
      const uint4 ww = __ldg(reinterpret_cast<const uint4*>(w + (size_t)(kh * 3 + kw) * d.in.c));
      const uint32_t vv[4] = {v.x, v.y, v.z, v.w}, wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fmaf(bf16_lo(vv[i]), bf16_lo(wv[i]), acc[2 * i]);
        acc[2 * i + 1] = fmaf(bf16_hi(vv[i]), bf16_hi(wv[i]), acc[2 * i + 1]);
      }
    }
  }
  if (d.act == Y11_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = silu(acc[i]);
  }
  if (d.res.ptr) {
    const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(d.res.ptr) + pix * d.res.c_total + d.res.c_off + g * 8);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] += bf16_lo(rr[i]);
      acc[2 * i + 1] += bf16_hi(rr[i]);
    }
  }
  *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(d.out.ptr) + pix * d.out.c_total + d.out.c_off + g * 8) =
      make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

int dwconv_launch(const y11_dwconv_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->in.c % 8 == 0 && d->in.c_off % 8 == 0 && d->in.c_total % 8 == 0 && d->out.c_off % 8 == 0 &&
                  d->out.c_total % 8 == 0,
              "dwconv: views must be 16-byte aligned (c=%d)", d->in.c);
  const size_t total = (size_t)d->B * d->H * d->W * (d->in.c / 8);
  Y11_CHECK_CUDA(y11_launch_pdl(dwconv_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, *d));
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
  static bool attr_set = false;
  if (!attr_set) {
    Y11_CHECK_CUDA(cudaFuncSetAttribute(sppf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
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
