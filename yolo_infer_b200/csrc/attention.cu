// Fused PSA attention core (YOLO11 C2PSA / PSABlock / Attention):
//   out[b, n, h*HD + :] = sum_m softmax_m(scale * <q[b,n,h,:], k[b,m,h,:]>) * v[b,m,h,:]
// Replaces the reference's view/split + (q^T k) matmul + softmax + (v attn^T) matmul (4 ATen kernels and an
// N x N score tensor in HBM) with one flash-style kernel: the score matrix never leaves registers (online softmax).
// N = H*W/1024 tokens (400 @640^2, 1600 @1280^2), key_dim 32, head_dim 64 for every YOLO11 scale.
//
// v2: both matmuls on tensor cores (mma.sync.m16n8k16 bf16 -> fp32; these are 16 x N x 32 and 16 x 64 x N problems per
// warp - far too small for a 128-row tcgen05 tile, and < 1 % of the network's FLOPs).  One CTA = 64 queries of one
// (image, head) = 4 warps x 16 queries; K/V stream through shared memory in 64-key tiles (rows padded by 16 B so the
// fragment loads / ldmatrix are bank-conflict free); P is re-used straight from the S accumulators as the A operand
// of the P.V product.  v1 (CUDA cores, one query per thread) took 0.49 ms for YOLO11s batch 64.
#include "ops.h"

using namespace y11;

namespace {
constexpr int kQ = 64;  // queries per CTA (4 warps x 16)
constexpr int kT = 64;  // keys per smem tile

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <int KD, int HD>
__global__ void __launch_bounds__(128) attn_kernel(y11_attn_desc d) {
  constexpr int KP = KD + 8, VP = HD + 8;  // padded rows (bf16 elements): 80 B / 144 B pitches
  // two K/V tile buffers: tile i+1 streams in (cp.async) while tile i is multiplied.  (v2 loaded each tile with plain
  // loads + shared stores between two barriers: 53 % of the kernel's stall samples sat on those stores, i.e. on the global
  // loads feeding them.)
  __shared__ __align__(16) __nv_bfloat16 s_kb[2][kT * KP];
  __shared__ __align__(16) __nv_bfloat16 s_vb[2][kT * VP];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * kQ + warp * 16;
  const int ct = d.qkv.c_total;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(d.qkv.ptr) + (size_t)b * d.N * ct + d.qkv.c_off;
  const __nv_bfloat16* qb = base + head * KD;
  const __nv_bfloat16* kb = base + d.heads * KD + head * KD;
  const __nv_bfloat16* vb = base + 2 * d.heads * KD + head * HD;

  // Q fragments (A operand, row-major 16 x KD): rows g / g+8, column pairs 2t / 2t+8 of every 16-wide k-step
  uint32_t qa[KD / 16][4];
  {
    const int r0 = min(q0 + g, d.N - 1), r1 = min(q0 + g + 8, d.N - 1);
#pragma unroll
    for (int ks = 0; ks < KD / 16; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * ct + ks * 16 + 2 * t);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * ct + ks * 16 + 2 * t);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * ct + ks * 16 + 2 * t + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * ct + ks * 16 + 2 * t + 8);
    }
  }
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = d.scale * 1.4426950408889634f;
  // async copy of the 64-key K and V tiles starting at key t0 into buffer `buf` (keys beyond N: zero-filled)
  auto load_tile = [&](int buf, int t0) {
    const uint32_t kdst = smem_u32(s_kb[buf]), vdst = smem_u32(s_vb[buf]);
    for (int i = threadIdx.x; i < kT * (KD / 8); i += 128) {
      const int row = i / (KD / 8), c = i % (KD / 8);
      const int key = t0 + row;
      cp_async_16(kdst + (uint32_t)(row * KP + c * 8) * 2u, kb + (size_t)min(key, d.N - 1) * ct + c * 8, key < d.N ? 16u : 0u);
    }
    for (int i = threadIdx.x; i < kT * (HD / 8); i += 128) {
      const int row = i / (HD / 8), c = i % (HD / 8);
      const int key = t0 + row;
      cp_async_16(vdst + (uint32_t)(row * VP + c * 8) * 2u, vb + (size_t)min(key, d.N - 1) * ct + c * 8, key < d.N ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_tile(0, 0);

  for (int t0 = 0, it = 0; t0 < d.N; t0 += kT, ++it) {
    const int buf = it & 1;
    const bool more = t0 + kT < d.N;
    if (more) load_tile(buf ^ 1, t0 + kT);  // the other buffer was released by the barrier that ended the previous iteration
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const __nv_bfloat16* s_k = s_kb[buf];
    const uint32_t sv_base = smem_u32(s_vb[buf]);

    // S = Q K^T for 16 queries x 64 keys: 8 n-blocks of 8 keys
    float s[kT / 8][4];
#pragma unroll
    for (int nb = 0; nb < kT / 8; ++nb) {
      s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KD / 16; ++ks) {
        const __nv_bfloat16* kp = s_k + (nb * 8 + g) * KP + ks * 16 + 2 * t;  // B fragment: B[k=d][n=key] = K[key][d]
        mma_bf16_16816(s[nb], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
      }
    }
    if (t0 + kT > d.N) {
#pragma unroll
      for (int nb = 0; nb < kT / 8; ++nb) {
        const int key = t0 + nb * 8 + 2 * t;
        if (key >= d.N) s[nb][0] = s[nb][2] = -INFINITY;
        if (key + 1 >= d.N) s[nb][1] = s[nb][3] = -INFINITY;
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nb = 0; nb < kT / 8; ++nb) {
      mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float corr0 = exp2f((m0 - mx0) * sc), corr1 = exp2f((m1 - mx1) * sc);  // first tile: m = -inf -> 0
    m0 = mx0; m1 = mx1;
    l0 *= corr0; l1 *= corr1;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
    // P = exp2((S - max) * scale*log2e), packed as the A operand of the P.V product (16 queries x 16 keys per k-step)
    uint32_t pa[kT / 16][4];
#pragma unroll
    for (int nb = 0; nb < kT / 8; ++nb) {
      const float p0 = exp2f((s[nb][0] - mx0) * sc), p1 = exp2f((s[nb][1] - mx0) * sc);
      const float p2 = exp2f((s[nb][2] - mx1) * sc), p3 = exp2f((s[nb][3] - mx1) * sc);
      l0 += p0 + p1; l1 += p2 + p3;
      pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    // O += P V : B[k=key][n=dim] = V[key][dim] -> transposing ldmatrix on the row-major V tile
#pragma unroll
    for (int kk = 0; kk < kT / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < HD / 8; db += 2) {
        uint32_t r0, r1, r2, r3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = db * 8 + (lane >> 4) * 8;
        ldmatrix_x4_trans(r0, r1, r2, r3, sv_base + (uint32_t)(row * VP + col) * 2u);
        mma_bf16_16816(o[db], pa[kk], r0, r1);
        mma_bf16_16816(o[db + 1], pa[kk], r2, r3);
      }
    }
    __syncthreads();  // everyone is done with this buffer before the next iteration's prefetch overwrites it
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(d.out.ptr) + (size_t)b * d.N * d.out.c_total + d.out.c_off + head * HD;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int db = 0; db < HD / 8; ++db) {
    if (r0 < d.N) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * d.out.c_total + db * 8 + 2 * t) = pack_bf16x2(o[db][0] * i0, o[db][1] * i0);
    if (r1 < d.N) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * d.out.c_total + db * 8 + 2 * t) = pack_bf16x2(o[db][2] * i1, o[db][3] * i1);
  }
}
}  // namespace

int attention_launch(const y11_attn_desc* d, cudaStream_t s) {
  Y11_REQUIRE(d->kd == 32 && d->hd == 64, "attention: only key_dim 32 / head_dim 64 (every YOLO11 scale), got %d/%d", d->kd, d->hd);
  Y11_REQUIRE(d->qkv.c_total % 8 == 0 && d->qkv.c_off % 8 == 0 && d->out.c_total % 8 == 0 && d->out.c_off % 8 == 0,
              "attention: views must be 16-byte aligned");
  dim3 grid((unsigned)((d->N + kQ - 1) / kQ), (unsigned)d->heads, (unsigned)d->B);
  Y11_CHECK_CUDA(y11_launch_pdl(attn_kernel<32, 64>, grid, dim3(128), 0, s, *d));
  return 0;
}
