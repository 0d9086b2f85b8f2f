// Implicit-GEMM convolution on Blackwell 5th-gen tensor cores (tcgen05.mma, TMEM accumulator, TMA operands).
//
// Replaces, for every dense conv of the fused YOLO11 network, what the reference executes as
// ultralytics Conv.forward_fuse -> torch conv2d (+SiLU) (+ Bottleneck/PSABlock residual add) (+ Concat copy)
// (SURVEY.md section 8a rows a6-a8, a11).
//
// GEMM view (NHWC bf16 activations, fp32 accumulation):
//   M = output pixels   : one CTA owns a 128-row tile = a Tw x Th x Tn box of (w, h, image)
//   N = output channels : BN <= 128 columns per CTA, accumulator = 128 lanes x BN fp32 columns of TMEM
//   K = (kh, kw, cin)   : one pipeline stage = one filter tap x Cc (16/32/64) input channels
// The A operand is never materialised (no im2col buffer): for each tap the TMA engine loads the SHIFTED
// Tw x Th x Tn x Cc box of the input straight into the canonical K-major swizzled smem layout UMMA reads;
// rows/cols outside the image are zero-filled by TMA, which is exactly the conv zero padding.  Stride-2
// convs address four parity sub-grids of the input (one tensor map each) so they are shifted boxes too.
// Epilogue (4 warps, one TMEM lane quadrant each): tcgen05.ld -> +bias -> SiLU -> +residual -> bf16/fp32
// -> written at a channel offset of the destination buffer (this is how Concat/chunk cost nothing).
//
// Warp roles: warp0 = TMA producer (1 elected lane), warp1 = MMA issuer (1 elected lane),
//             warps2-5 = epilogue; warp2 also owns TMEM alloc/dealloc.
#include <algorithm>
#include <cstring>

#include "ops.h"

namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr uint32_t kHeaderBytes = 256;  // mbarriers + tmem pointer

__global__ void __launch_bounds__(kThreads)
conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcParams p) {
  using namespace y11;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t tiles_base = (smem_base + kHeaderBytes + 1023u) & ~1023u;
  const uint32_t full_bar = smem_base;                    // kMaxStages x 8 B
  const uint32_t empty_bar = smem_base + 8 * kMaxStages;  // kMaxStages x 8 B
  const uint32_t acc_bar = smem_base + 16 * kMaxStages;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_raw + 16 * kMaxStages + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates: Cout tile fastest so CTAs sharing an activation tile are launched together (L2 reuse)
  int t = blockIdx.x;
  const int nt = t % p.n_tiles; t /= p.n_tiles;
  const int tw_i = t % p.tiles_w; t /= p.tiles_w;
  const int th_i = t % p.tiles_h; t /= p.tiles_h;
  const int tn_i = t;
  const int w0 = tw_i * p.Tw, h0 = th_i * p.Th, n0 = tn_i * p.Tn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_ptr_smem;

  const int k_iters = p.taps * p.chunks_per_tap;
  const uint32_t stage_bytes = p.a_slot + p.b_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int it = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        int mi = 0, cw, ch;
        if (p.ksize == 1) {
          cw = w0; ch = h0;
        } else if (p.stride == 1) {
          cw = w0 + tap % 3 - 1; ch = h0 + tap / 3 - 1;
        } else {
          const int kh = tap / 3, kw = tap % 3;
          mi = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1);  // input row 2*oy+kh-1 has parity (kh != 1)
          cw = w0 - (kw == 0); ch = h0 - (kh == 0);
        }
        for (int c = 0; c < p.chunks_per_tap; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(empty_bar + 8 * s, ph ^ 1, p.err_flag, 101);
          mbar_expect_tx(full_bar + 8 * s, p.tx_bytes);
          const uint32_t a_dst = tiles_base + s * stage_bytes;
          tma_load_4d(a_dst, &maps.a[mi], full_bar + 8 * s, c * p.Cc, cw, ch, n0);
          tma_load_2d(a_dst + p.a_slot, &maps.b, full_bar + 8 * s, tap * p.cin + c * p.Cc, nt * p.BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16_m128(p.BN);
      const int kk_n = p.Cc / 16;
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(full_bar + 8 * s, ph, p.err_flag, 102);
        tc_fence_after();
        const uint32_t a_src = tiles_base + s * stage_bytes;
        const uint64_t ad = make_umma_desc(a_src, p.sbo, p.layout_type);
        const uint64_t bd = make_umma_desc(a_src + p.a_slot, p.sbo, p.layout_type);
        for (int kk = 0; kk < kk_n; ++kk)  // +32 B (= 16 bf16 of K) inside the swizzle row per UMMA
          umma_bf16(tmem_acc, ad + 2 * kk, bd + 2 * kk, idesc, (it | kk) != 0);
        umma_commit(empty_bar + 8 * s);  // frees the smem slot once these MMAs retire
      }
      umma_commit(acc_bar);  // accumulator complete
    }
  } else {
    // -------------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    const int tw = r % p.Tw, th = (r / p.Tw) % p.Th, tn = r / (p.Tw * p.Th);
    const int ow = w0 + tw, oh = h0 + th, on = n0 + tn;
    const bool valid = (r < p.Tw * p.Th * p.Tn) && ow < p.Wout && oh < p.Hout && on < p.B;
    const size_t pix = (static_cast<size_t>(on) * p.Hout + oh) * p.Wout + ow;
    mbar_wait(acc_bar, 0, p.err_flag, 103);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      const int n = nt * p.BN + c0;
      if (valid && n < p.cout) {
        float f[16];
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = __ldg(b4 + i);
          f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x;
          f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
          f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z;
          f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
        }
        if (p.act == Y11_ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = silu(f[i]);
        }
        if (p.res) {
          const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.res) + pix * p.res_ct + p.res_co + n);
          const uint4 r0 = rp[0], r1 = rp[1];
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[2 * i] += bf16_lo(rr[i]);
            f[2 * i + 1] += bf16_hi(rr[i]);
          }
        }
        if (p.out_f32) {
          float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix * p.out_ct + p.out_co + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else {
          uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.out_ct + p.out_co + n);
          op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_acc, p.tmem_cols);
}

int encode_map(y11_engine* eng, CUtensorMap* m, int rank, void* base, const cuuint64_t* gdim, const cuuint64_t* gstr,
               const cuuint32_t* box, CUtensorMapSwizzle swz) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = eng->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

int conv_tc_prepare(y11_engine* eng, const y11_conv_desc* d, ConvTcLaunch* L) {
  Y11_REQUIRE(eng && eng->encode_tiled, "conv_tc: engine has no cuTensorMapEncodeTiled entry point");
  Y11_REQUIRE((d->k == 1 && d->stride == 1) || (d->k == 3 && (d->stride == 1 || d->stride == 2)),
              "conv_tc: unsupported k=%d stride=%d", d->k, d->stride);
  const int cin = d->in.c, cout = d->out.c;
  Y11_REQUIRE(cin % 16 == 0 && cout % 16 == 0, "conv_tc: cin=%d cout=%d must be multiples of 16", cin, cout);
  Y11_REQUIRE(d->in.c_total % 8 == 0 && d->in.c_off % 8 == 0, "conv_tc: input view must be 16-byte aligned");
  Y11_REQUIRE(d->out.c_off % (d->out_f32 ? 4 : 8) == 0 && d->out.c_total % (d->out_f32 ? 4 : 8) == 0,
              "conv_tc: output view must be 16-byte aligned");
  Y11_REQUIRE(!d->res.ptr || (d->res.c_off % 8 == 0 && d->res.c_total % 8 == 0), "conv_tc: residual view alignment");
  if (d->stride == 1) Y11_REQUIRE(d->Hout == d->Hin && d->Wout == d->Win, "conv_tc: stride-1 shape mismatch");
  if (d->stride == 2) Y11_REQUIRE(d->Hout == (d->Hin + 1) / 2 && d->Wout == (d->Win + 1) / 2, "conv_tc: stride-2 shape mismatch");

  std::memset(L, 0, sizeof(*L));
  ConvTcParams& p = L->p;
  pick_tile(d->Wout, d->Hout, d->B, &p.Tw, &p.Th, &p.Tn);
  p.tiles_w = y11_ceil_div(d->Wout, p.Tw);
  p.tiles_h = y11_ceil_div(d->Hout, p.Th);
  p.tiles_n = y11_ceil_div(d->B, p.Tn);
  // N tile: largest multiple of 16 that divides cout and is <= 128
  int bn = 16;
  for (int c = 16; c <= std::min(cout, 128); c += 16)
    if (cout % c == 0) bn = c;
  p.BN = bn;
  p.n_tiles = cout / bn;
  p.Cc = (cin % 64 == 0) ? 64 : (cin % 32 == 0) ? 32 : 16;
  p.cin = cin;
  p.chunks_per_tap = cin / p.Cc;
  p.taps = d->k * d->k;
  p.ksize = d->k;
  p.stride = d->stride;
  const uint32_t swz_bytes = p.Cc * 2;
  p.sbo = 8 * swz_bytes;
  p.layout_type = (p.Cc == 64) ? 2u : (p.Cc == 32) ? 4u : 6u;  // SWIZZLE_128B / 64B / 32B
  const CUtensorMapSwizzle swz = (p.Cc == 64)   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (p.Cc == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  p.a_slot = 128u * swz_bytes;
  p.b_slot = ((uint32_t)bn * swz_bytes + 1023u) & ~1023u;
  p.tx_bytes = (uint32_t)(p.Tw * p.Th * p.Tn) * swz_bytes + (uint32_t)bn * swz_bytes;
  const int k_iters = p.taps * p.chunks_per_tap;
  const uint32_t stage = p.a_slot + p.b_slot;
  int stages = (int)((96u * 1024u) / stage);
  stages = std::max(2, std::min(std::min(stages, kMaxStages), k_iters));
  if (stages < 1) stages = 1;
  p.stages = stages;
  int cols = 32;
  while (cols < bn) cols *= 2;
  p.tmem_cols = cols;
  p.B = d->B; p.Hout = d->Hout; p.Wout = d->Wout; p.cout = cout;
  p.out = d->out.ptr; p.out_ct = d->out.c_total; p.out_co = d->out.c_off; p.out_f32 = d->out_f32;
  p.res = d->res.ptr; p.res_ct = d->res.c_total; p.res_co = d->res.c_off;
  p.bias = d->bias; p.act = d->act;
  p.err_flag = eng->dev_error_flag;

  // activation tensor maps
  const size_t ct = d->in.c_total;
  __nv_bfloat16* in_base = static_cast<__nv_bfloat16*>(d->in.ptr) + d->in.c_off;
  const cuuint32_t box[4] = {(cuuint32_t)p.Cc, (cuuint32_t)p.Tw, (cuuint32_t)p.Th, (cuuint32_t)p.Tn};
  if (d->stride == 1) {
    const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
    const cuuint64_t gstr[3] = {ct * 2, ct * 2 * d->Win, ct * 2 * d->Win * d->Hin};
    if (int e = encode_map(eng, &L->maps.a[0], 4, in_base, gdim, gstr, box, swz)) return e;
    L->maps.a[1] = L->maps.a[2] = L->maps.a[3] = L->maps.a[0];
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)((d->Win - pw + 1) / 2), (cuuint64_t)((d->Hin - ph + 1) / 2),
                                    (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {ct * 2 * 2, ct * 2 * d->Win * 2, ct * 2 * d->Win * d->Hin};
        __nv_bfloat16* base = in_base + ((size_t)ph * d->Win + pw) * ct;
        if (int e = encode_map(eng, &L->maps.a[ph * 2 + pw], 4, base, gdim, gstr, box, swz)) return e;
      }
  }
  {
    const cuuint64_t K = (cuuint64_t)p.taps * cin;
    const cuuint64_t gdim[2] = {K, (cuuint64_t)cout};
    const cuuint64_t gstr[1] = {K * 2};
    const cuuint32_t bbox[2] = {(cuuint32_t)p.Cc, (cuuint32_t)bn};
    if (int e = encode_map(eng, &L->maps.b, 2, const_cast<void*>(d->w), gdim, gstr, bbox, swz)) return e;
  }
  L->grid = (unsigned)(p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles);
  L->smem_bytes = kHeaderBytes + 1024u + (unsigned)stages * stage;
  L->flops = 2.0 * d->B * d->Hout * d->Wout * (double)cout * cin * p.taps;
  static bool attr_set = false;
  if (!attr_set) {
    Y11_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  return 0;
}

int conv_tc_launch(const ConvTcLaunch* L, cudaStream_t s) {
  conv_tc_kernel<<<L->grid, kThreads, L->smem_bytes, s>>>(L->maps, L->p);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}
