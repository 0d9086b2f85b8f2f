// Fused PSA attention core (YOLO11 C2PSA / PSABlock / Attention):
//   out[b, n, h*HD + :] = sum_m softmax_m(scale * <q[b,n,h,:], k[b,m,h,:]>) * v[b,m,h,:]
// Replaces the reference's view/split + (q^T k) matmul + softmax + (v attn^T) matmul (4 ATen kernels and an
// N x N score tensor in HBM) with one kernel: the score matrix never leaves registers (online softmax).
// N = H*W/1024 tokens (400 @640^2, 1600 @1280^2), key_dim 32, head_dim 64 for every YOLO11 scale.
//
// One thread owns one query row (q and the 64-wide output accumulator live in registers, fp32); a CTA of 128
// queries streams K/V of its (image, head) through shared memory in 64-key tiles; every smem read is a
// warp-wide broadcast of one key/value row.
#include "ops.h"

using namespace y11;

namespace {
constexpr int kQ = 128;  // queries per CTA
constexpr int kT = 64;   // keys per smem tile
constexpr int kJ = 8;    // keys per online-softmax update

template <int KD, int HD>
__global__ void __launch_bounds__(kQ) attn_kernel(y11_attn_desc d) {
  __shared__ uint4 s_k[kT * KD / 8];
  __shared__ uint4 s_v[kT * HD / 8];
  const int head = blockIdx.y, b = blockIdx.z;
  const int n = blockIdx.x * kQ + threadIdx.x;
  const bool qvalid = n < d.N;
  const int ct = d.qkv.c_total;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(d.qkv.ptr) + (size_t)b * d.N * ct + d.qkv.c_off;
  const __nv_bfloat16* qb = base + head * KD;
  const __nv_bfloat16* kb = base + d.heads * KD + head * KD;
  const __nv_bfloat16* vb = base + 2 * d.heads * KD + head * HD;

  float q[KD];
  {
    const float sc = d.scale * 1.4426950408889634f;
    const uint4* qp = reinterpret_cast<const uint4*>(qb + (size_t)(qvalid ? n : 0) * ct);
#pragma unroll
    for (int i = 0; i < KD / 8; ++i) {
      const uint4 u = qp[i];
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q[i * 8 + 2 * j] = bf16_lo(uu[j]) * sc;
        q[i * 8 + 2 * j + 1] = bf16_hi(uu[j]) * sc;
      }
    }
  }
  float acc[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;

  for (int t0 = 0; t0 < d.N; t0 += kT) {
    __syncthreads();
    for (int i = threadIdx.x; i < kT * (KD / 8); i += kQ) {
      const int row = i / (KD / 8), c = i % (KD / 8);
      const int key = t0 + row;
      s_k[i] = key < d.N ? *(reinterpret_cast<const uint4*>(kb + (size_t)key * ct) + c) : make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < kT * (HD / 8); i += kQ) {
      const int row = i / (HD / 8), c = i % (HD / 8);
      const int key = t0 + row;
      s_v[i] = key < d.N ? *(reinterpret_cast<const uint4*>(vb + (size_t)key * ct) + c) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (!qvalid) continue;
    for (int j0 = 0; j0 < kT && t0 + j0 < d.N; j0 += kJ) {
      float s[kJ];
      float mx = m;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < KD / 8; ++i) {
          const uint4 u = s_k[(j0 + j) * (KD / 8) + i];
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            a = fmaf(q[i * 8 + 2 * e], bf16_lo(uu[e]), a);
            a = fmaf(q[i * 8 + 2 * e + 1], bf16_hi(uu[e]), a);
          }
        }
        s[j] = (t0 + j0 + j < d.N) ? a : -INFINITY;
        mx = fmaxf(mx, s[j]);
      }
      const float corr = exp2f(m - mx);  // m = -inf on the first block -> 0
      m = mx;
      l *= corr;
#pragma unroll
      for (int i = 0; i < HD; ++i) acc[i] *= corr;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const float pj = exp2f(s[j] - mx);
        l += pj;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          const uint4 u = s_v[(j0 + j) * (HD / 8) + i];
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[i * 8 + 2 * e] = fmaf(pj, bf16_lo(uu[e]), acc[i * 8 + 2 * e]);
            acc[i * 8 + 2 * e + 1] = fmaf(pj, bf16_hi(uu[e]), acc[i * 8 + 2 * e + 1]);
          }
        }
      }
    }
  }
  if (!qvalid) return;
  const float inv = 1.0f / l;
  uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(d.out.ptr) + ((size_t)b * d.N + n) * d.out.c_total + d.out.c_off + head * HD);
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
    op[i] = make_uint4(pack_bf16x2(acc[8 * i] * inv, acc[8 * i + 1] * inv), pack_bf16x2(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv),
                       pack_bf16x2(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv), pack_bf16x2(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv));
}
}  // namespace

int attention_launch(const y11_attn_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->kd == 32 && d->hd == 64, "attention: only key_dim 32 / head_dim 64 (every YOLO11 scale), got %d/%d", d->kd, d->hd);
  Y11_REQUIRE(d->qkv.c_total % 8 == 0 && d->qkv.c_off % 8 == 0 && d->out.c_total % 8 == 0 && d->out.c_off % 8 == 0,
              "attention: views must be 16-byte aligned");
  dim3 grid((unsigned)((d->N + kQ - 1) / kQ), (unsigned)d->heads, (unsigned)d->B);
  attn_kernel<32, 64><<<grid, kQ, 0, s>>>(*d);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}
