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
// stem: 3 -> COUT, 3x3 stride 2 pad 1, +bias, SiLU.  One thread = one output pixel x all COUT channels
// (accumulators in registers), input rows staged in shared memory as fp32, weights broadcast from smem.
// ------------------------------------------------------------------------------------------------
constexpr int kStemPix = 128;  // output pixels (along W) per CTA

template <int COUT>
__global__ void __launch_bounds__(kStemPix) stem_kernel(y11_stem_desc d) {
  __shared__ float s_in[3][(2 * kStemPix + 1) * 3];
  __shared__ float s_w[27][COUT];
  __shared__ float s_b[COUT];
  const int tiles_w = (d.Wout + kStemPix - 1) / kStemPix;
  const int tile = blockIdx.x % tiles_w;
  const int oh = (blockIdx.x / tiles_w) % d.Hout;
  const int n = blockIdx.x / (tiles_w * d.Hout);
  const int ow0 = tile * kStemPix;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w);
  for (int i = threadIdx.x; i < 27 * COUT; i += kStemPix) s_w[i % 27][i / 27] = __bfloat162float(w[i]);  // w is [COUT][27]
  for (int i = threadIdx.x; i < COUT; i += kStemPix) s_b[i] = d.bias[i];
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(d.in);
  const int iw0 = 2 * ow0 - 1;
  constexpr int kRow = (2 * kStemPix + 1) * 3;
  for (int r = 0; r < 3; ++r) {
    const int ih = 2 * oh + r - 1;
    const bool row_ok = ih >= 0 && ih < d.Hin;
    const __nv_bfloat16* rp = in + ((size_t)n * d.Hin + (row_ok ? ih : 0)) * d.Win * 3;
    for (int i = threadIdx.x; i < kRow; i += kStemPix) {
      const int iw = iw0 + i / 3;
      float v = 0.f;
      if (row_ok && iw >= 0 && iw < d.Win) v = __bfloat162float(rp[(size_t)iw * 3 + i % 3]);
      s_in[r][i] = v;
    }
  }
  __syncthreads();
  const int ow = ow0 + threadIdx.x;
  if (ow >= d.Wout) return;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = s_b[c];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {  // j = kw*3 + c
      const float x = s_in[kh][threadIdx.x * 6 + j];
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[c] = fmaf(x, s_w[kh * 9 + j][c], acc[c]);
    }
  }
  const size_t pix = ((size_t)n * d.Hout + oh) * d.Wout + ow;
  uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(d.out.ptr) + pix * d.out.c_total + d.out.c_off);
#pragma unroll
  for (int c = 0; c < COUT; c += 8) {
    op[c / 8] = make_uint4(pack_bf16x2(silu(acc[c]), silu(acc[c + 1])), pack_bf16x2(silu(acc[c + 2]), silu(acc[c + 3])),
                           pack_bf16x2(silu(acc[c + 4]), silu(acc[c + 5])), pack_bf16x2(silu(acc[c + 6]), silu(acc[c + 7])));
  }
}

int stem_launch(const y11_stem_desc* d, cudaStream_t s) {
  const int tiles_w = (d->Wout + kStemPix - 1) / kStemPix;
  const unsigned grid = (unsigned)(tiles_w * d->Hout * d->B);
  switch (d->out.c) {
    case 16: stem_kernel<16><<<grid, kStemPix, 0, s>>>(*d); break;
    case 32: stem_kernel<32><<<grid, kStemPix, 0, s>>>(*d); break;
    case 64: stem_kernel<64><<<grid, kStemPix, 0, s>>>(*d); break;
    case 96: stem_kernel<96><<<grid, kStemPix, 0, s>>>(*d); break;
    default: y11_set_error("stem: unsupported cout %d", d->out.c); return -1;
  }
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 stride 1 pad 1: one thread = one pixel x 8 channels (16-byte vectors)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv_kernel(y11_dwconv_desc d) {
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
  dwconv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*d);
  Y11_CHECK_CUDA(cudaGetLastError());
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
  sppf_kernel<<<(unsigned)(d->B * (d->c / 8)), 256, smem, s>>>(*d);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// nearest 2x upsample into a channel slice of the concat buffer (16-byte vectors)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(y11_upsample_desc d) {
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
  upsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*d);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}
