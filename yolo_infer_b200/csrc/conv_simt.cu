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
// stem: 3 -> COUT, 3x3 stride 2 pad 1, +bias, SiLU (CUDA cores: K = 27 is too thin for an MMA tile).
// Register blocked: one thread owns P output pixels x all COUT channels, so every broadcast weight load from shared
// memory feeds 4*P FMAs (v1, 1 pixel per thread, was LDS-issue bound: 0.70 ms for YOLO11s batch 64).  A CTA covers
// TX*P consecutive pixels of one output row; thread tx owns pixels tx, tx+TX, ... (adjacent lanes -> adjacent pixels).
// The three input rows are staged with 16-byte loads (v2 staged 2 bytes at a time and was 3x slower than v1).
// ------------------------------------------------------------------------------------------------
template <int COUT, int P>
__global__ void __launch_bounds__(128) stem_kernel(y11_stem_desc d, int TX, int tiles_w) {
  extern __shared__ float s_dyn[];
  float* s_w = s_dyn;                  // [27][COUT]
  float* s_b = s_w + 27 * COUT;        // [COUT]
  float* s_in = s_b + COUT;            // [3][(2*TX*P+1)*3]
  const int span = TX * P;
  const int row_f = (2 * span + 1) * 3;
  const int tile = blockIdx.x % tiles_w;
  const int oh = (blockIdx.x / tiles_w) % d.Hout;
  const int n = blockIdx.x / (tiles_w * d.Hout);
  const int ow0 = tile * span;
  const int nt = blockDim.x;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w);
  for (int i = threadIdx.x; i < 27 * COUT; i += nt) s_w[(i % 27) * COUT + i / 27] = __bfloat162float(w[i]);  // w is [COUT][27]
  for (int i = threadIdx.x; i < COUT; i += nt) s_b[i] = d.bias[i];
  for (int i = threadIdx.x; i < 3 * row_f; i += nt) s_in[i] = 0.f;
  __syncthreads();
  const int iw0 = 2 * ow0 - 1;
  const int e_lo = max(iw0, 0) * 3, e_hi = min(iw0 + 2 * span + 1, d.Win) * 3;  // bf16 elements of the input row we need
  const int v_lo = e_lo / 8, v_hi = (e_hi + 7) / 8;                             // 16-byte chunks (rows are 16-byte aligned)
  for (int r = 0; r < 3; ++r) {
    const int ih = 2 * oh + r - 1;
    if (ih < 0 || ih >= d.Hin) continue;
    const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(d.in) + ((size_t)n * d.Hin + ih) * d.Win * 3);
    float* dst = s_in + r * row_f - iw0 * 3;
    for (int v = v_lo + threadIdx.x; v < v_hi; v += nt) {
      const uint4 u = __ldg(rp + v);
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int e = v * 8 + 2 * t;
        if (e >= e_lo && e < e_hi) dst[e] = bf16_lo(uu[t]);
        if (e + 1 >= e_lo && e + 1 < e_hi) dst[e + 1] = bf16_hi(uu[t]);
      }
    }
  }
  __syncthreads();
  const int tx = threadIdx.x;
  if (tx >= TX) return;
  float acc[P][COUT];
#pragma unroll
  for (int q = 0; q < P; ++q)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[q][c] = s_b[c];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {  // j = kw*3 + c
      float x[P];
#pragma unroll
      for (int q = 0; q < P; ++q) x[q] = s_in[kh * row_f + (tx + q * TX) * 6 + j];
      const float4* wr = reinterpret_cast<const float4*>(s_w + (kh * 9 + j) * COUT);
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        const float4 ww = wr[c4];
#pragma unroll
        for (int q = 0; q < P; ++q) {
          acc[q][4 * c4 + 0] = fmaf(x[q], ww.x, acc[q][4 * c4 + 0]);
          acc[q][4 * c4 + 1] = fmaf(x[q], ww.y, acc[q][4 * c4 + 1]);
          acc[q][4 * c4 + 2] = fmaf(x[q], ww.z, acc[q][4 * c4 + 2]);
          acc[q][4 * c4 + 3] = fmaf(x[q], ww.w, acc[q][4 * c4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const int ow = ow0 + tx + q * TX;
    if (ow >= d.Wout) continue;
    const size_t pix = ((size_t)n * d.Hout + oh) * d.Wout + ow;
    uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(d.out.ptr) + pix * d.out.c_total + d.out.c_off);
#pragma unroll
    for (int c = 0; c < COUT; c += 8) {
      op[c / 8] = make_uint4(pack_bf16x2(silu(acc[q][c]), silu(acc[q][c + 1])), pack_bf16x2(silu(acc[q][c + 2]), silu(acc[q][c + 3])),
                             pack_bf16x2(silu(acc[q][c + 4]), silu(acc[q][c + 5])), pack_bf16x2(silu(acc[q][c + 6]), silu(acc[q][c + 7])));
    }
  }
}

template <int COUT, int P>
static int stem_launch_t(const y11_stem_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->Win % 8 == 0, "stem: input width must be a multiple of 8 (got %d)", d->Win);
  const int nblk = y11_ceil_div(d->Wout, P * 128);
  const int TX = y11_ceil_div(d->Wout, P * nblk);
  const int threads = (TX + 31) / 32 * 32;
  const size_t smem = (size_t)(27 * COUT + COUT + 3 * (2 * TX * P + 1) * 3) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    Y11_CHECK_CUDA(cudaFuncSetAttribute(stem_kernel<COUT, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  Y11_REQUIRE(smem <= 64 * 1024, "stem: row too wide for shared memory");
  stem_kernel<COUT, P><<<(unsigned)(nblk * d->Hout * d->B), threads, smem, s>>>(*d, TX, nblk);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int stem_launch(const y11_stem_desc* d, cudaStream_t s) {
  switch (d->out.c) {
    case 16: return stem_launch_t<16, 4>(d, s);
    case 32: return stem_launch_t<32, 4>(d, s);
    case 64: return stem_launch_t<64, 2>(d, s);
    case 96: return stem_launch_t<96, 1>(d, s);
    default: y11_set_error("stem: unsupported cout %d", d->out.c); return -1;
  }
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
