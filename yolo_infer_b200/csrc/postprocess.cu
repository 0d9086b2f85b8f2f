// Detect-head post-processing on device:
//   decode  : DFL softmax-expectation -> dist2bbox against anchor points -> x stride; class sigmoid
//             (ultralytics Detect._inference / DFL / make_anchors / dist2bbox, ~12 ATen kernels)
//   compact : `score > conf` candidates in ANCHOR ORDER via warp ballots + prefix sums (two passes, no staging)
//             (ultralytics non_max_suppression's boolean-mask indexing, which syncs the host per image)
//   sort    : stable descending by score (unique 64-bit keys (~score, index), bitonic)
//   nms     : class-aware greedy IoU NMS, bit-exact with torchvision.ops.nms CPU semantics: 256x256 bitmask
//             tiles + a warp-level suppression sweep over the bit rows; a candidate is tested only against the
//             boxes kept so far, so work is K*kept instead of K^2 and stops at max_det
//   output  : [:max_det], scale_boxes + clip_boxes fused into the write.
// All boxes/score arithmetic that feeds a comparison uses explicit round-to-nearest intrinsics so that nvcc
// cannot contract it into FMAs (the reference computes every step as a separate fp32 op).
#include <math.h>

#include <algorithm>

#include "ops.h"

using namespace y11;

namespace {

constexpr int kChunk = 64;         // anchors per CTA in decode/compaction (4 lanes per anchor, 256 threads)
constexpr int kNmsThreads = 1024;  // sort + nms CTA
constexpr int kTile = 256;         // candidates per NMS tile
constexpr int kSmemKeys = 16384;   // sort in shared memory up to this many keys (128 KB)
constexpr int kKeptSmem = 1024;    // kept-box list in shared memory up to this max_det (20 KB)

struct HeadParams {
  const float* head[3];
  int hl[3], wl[3];
  int off[4];  // anchor offset per level
  float stride[3];
  int nl, B, nc, A, no;
  int cls_per_lane;  // classes scanned by each of the 4 lanes of an anchor: multiple of 4, 4*cls_per_lane >= nc
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// ---- decode building blocks: FOUR lanes cooperate on one anchor ------------------------------------------------
// lane `sub` (0..3) owns DFL side `sub` (l,t,r,b: 16 logits = 64 contiguous bytes) and a contiguous quarter of the
// class logits.  Every helper below is used by the dense kernel (tests) AND the compaction kernels, so both produce
// bit-identical boxes and scores.
struct AnchorRef {
  const float* row;
  float ax, ay, stride;
};

__device__ __forceinline__ AnchorRef anchor_ref(const HeadParams& hp, int b, int a) {
  // level selection with constant indices only: a runtime index into the by-value parameter arrays makes the compiler copy
  // them to local memory at kernel entry (8 STL per thread in the ncu source view)
  const bool l1 = hp.nl > 1 && a >= hp.off[1], l2 = hp.nl > 2 && a >= hp.off[2];
  const int off = l2 ? hp.off[2] : l1 ? hp.off[1] : hp.off[0];
  const int wl = l2 ? hp.wl[2] : l1 ? hp.wl[1] : hp.wl[0];
  const int hl = l2 ? hp.hl[2] : l1 ? hp.hl[1] : hp.hl[0];
  const float* head = l2 ? hp.head[2] : l1 ? hp.head[1] : hp.head[0];
  const int i = a - off;
  AnchorRef r;
  r.row = head + ((size_t)b * hl * wl + i) * hp.no;
  r.ax = (float)(i % wl) + 0.5f;
  r.ay = (float)(i / wl) + 0.5f;
  r.stride = l2 ? hp.stride[2] : l1 ? hp.stride[1] : hp.stride[0];
  return r;
}

// softmax expectation over the 16 bins of one side (ultralytics DFL: softmax then conv with arange(16))
__device__ __forceinline__ float dfl_side(const float* p) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  float mx = v[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) mx = fmaxf(mx, v[i]);
  float den = 0.f, num = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float e = expf(v[i] - mx);
    den += e;
    num = fmaf(e, (float)i, num);
  }
  return __fdiv_rn(num, den);
}

// maximum of this lane's `q` class logits [c0, c0+q) (q a multiple of 4, <= 4*kMaxClsVec), all loads issued up front
constexpr int kMaxClsVec = 8;  // nc <= 128 -> at most 32 classes = 8 float4 per lane
__device__ __forceinline__ float lane_class_max(const float* cls_row, int q, int c0, int nc) {
  float4 t[kMaxClsVec];
#pragma unroll
  for (int j = 0; j < kMaxClsVec; ++j)
    t[j] = (4 * j < q && c0 + 4 * j < nc) ? __ldg(reinterpret_cast<const float4*>(cls_row + c0) + j) : make_float4(0, 0, 0, 0);
  float best = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxClsVec; ++j) {
    const float v[4] = {t[j].x, t[j].y, t[j].z, t[j].w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4 * j + k < q && c0 + 4 * j + k < nc) best = fmaxf(best, v[k]);
  }
  return best;
}

// all 4 lanes of the anchor's group call this (gmask = the group's lanes); returns cx,cy,w,h in every lane
__device__ __forceinline__ float4 decode_box(const AnchorRef& ar, int sub, int lane, unsigned gmask) {
  const float d = dfl_side(ar.row + 16 * sub);
  const int g0 = lane & ~3;
  const float dl = __shfl_sync(gmask, d, g0), dt = __shfl_sync(gmask, d, g0 + 1);
  const float dr = __shfl_sync(gmask, d, g0 + 2), db = __shfl_sync(gmask, d, g0 + 3);
  const float x1 = __fsub_rn(ar.ax, dl), y1 = __fsub_rn(ar.ay, dt), x2 = __fadd_rn(ar.ax, dr), y2 = __fadd_rn(ar.ay, db);
  return make_float4(__fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), ar.stride), __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), ar.stride),
                     __fmul_rn(__fsub_rn(x2, x1), ar.stride), __fmul_rn(__fsub_rn(y2, y1), ar.stride));
}

// xywh2xyxy exactly as the reference: wh/2 first, then xy -/+ it
__device__ __forceinline__ float4 xywh2xyxy_rn(float4 b) {
  const float hw = __fdiv_rn(b.z, 2.0f), hh = __fdiv_rn(b.w, 2.0f);
  return make_float4(__fsub_rn(b.x, hw), __fsub_rn(b.y, hh), __fadd_rn(b.x, hw), __fadd_rn(b.y, hh));
}

__global__ void __launch_bounds__(256) decode_dense_kernel(HeadParams hp, float* __restrict__ y) {
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int a = blockIdx.x * kChunk + (threadIdx.x >> 2);
  if (a >= hp.A) return;  // whole 4-lane groups leave together
  const unsigned gmask = 0xFu << (lane & ~3);
  const AnchorRef ar = anchor_ref(hp, b, a);
  const float4 box = decode_box(ar, sub, lane, gmask);
  float* yb = y + (size_t)b * (4 + hp.nc) * hp.A;
  const float bx[4] = {box.x, box.y, box.z, box.w};
  yb[(size_t)sub * hp.A + a] = bx[sub];
  for (int c = sub; c < hp.nc; c += 4) yb[(size_t)(4 + c) * hp.A + a] = sigmoidf_acc(__ldg(ar.row + 64 + c));
}

// MODE 0: count candidates per 64-anchor chunk.  MODE 1: write candidates at chunk_off + in-chunk prefix (anchor order).
// (Single-label with A < 65536 uses decode_onepass_kernel below instead of these two passes.)
// Only the class logits are read to decide candidacy (sigmoid is monotone: max score = sigmoid(max logit)); the DFL
// softmaxes (64 expf per anchor) run for candidates only, in MODE 1.
template <int MODE>
__global__ void __launch_bounds__(256)
decode_compact_kernel(HeadParams hp, float conf, float logit_lo, int multi_label, int cap, int nchunks, int* __restrict__ chunk_cnt,
                      const int* __restrict__ chunk_off, float4* __restrict__ cbox, float* __restrict__ cscore, float* __restrict__ ccls,
                      int* __restrict__ canchor, int* __restrict__ ncand, int* __restrict__ ncand_raw) {
  __shared__ int s_warp[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int a = blockIdx.x * kChunk + (threadIdx.x >> 2);
  const bool in_range = a < hp.A;
  const unsigned gmask = 0xFu << (lane & ~3);
  AnchorRef ar;
  ar.row = nullptr; ar.ax = ar.ay = ar.stride = 0.f;
  if (in_range) ar = anchor_ref(hp, b, a);
  // ---- class scan: this lane's quarter of the classes
  const int q = hp.cls_per_lane;  // multiple of 4, 4*q >= nc
  const int c0 = sub * q;
  float best = -INFINITY;
  unsigned mask = 0u;  // multi-label: bit j set <=> class c0+j passes
  if (in_range) {
    for (int j = 0; j < q; j += 4) {
      if (c0 + j >= hp.nc) break;
      const float4 t = __ldg(reinterpret_cast<const float4*>(ar.row + 64 + c0 + j));
      const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c0 + j + k < hp.nc) {
          best = fmaxf(best, v[k]);
          if (multi_label && v[k] > logit_lo && sigmoidf_acc(v[k]) > conf) mask |= 1u << (j + k);
        }
      }
    }
  }
  best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
  best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
  const float score = in_range ? sigmoidf_acc(best) : -1.f;
  int mycnt;  // per lane: multi-label -> my passing classes; single-label -> leader lane carries the anchor's flag
  if (multi_label) mycnt = __popc(mask);
  else mycnt = (sub == 0 && score > conf) ? 1 : 0;
  // ---- in-warp inclusive scan, warp totals, chunk base
  int incl = mycnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < 8; ++i) t += s_warp[i];
      chunk_cnt[b * nchunks + blockIdx.x] = t;
    }
    return;
  }
  int pos = chunk_off[b * nchunks + blockIdx.x] + incl - mycnt;
  for (int i = 0; i < warp; ++i) pos += s_warp[i];
  const size_t ob = (size_t)b * cap;
  const bool cand = multi_label ? (__ballot_sync(0xffffffffu, mask != 0u) & gmask) != 0u : (score > conf);
  if (!cand) return;  // whole 4-lane groups leave together
  const float4 box = xywh2xyxy_rn(decode_box(ar, sub, lane, gmask));
  if (multi_label) {
    // rows in (anchor, class) order: my classes come after those of the lanes with a smaller `sub` (pos is already
    // the exclusive prefix over lanes, which ARE ordered (anchor, sub))
    int p = pos;
    for (unsigned m = mask; m; m &= m - 1u) {
      const int j = __ffs(m) - 1;
      if (p < cap) {
        cbox[ob + p] = box;
        cscore[ob + p] = sigmoidf_acc(__ldg(ar.row + 64 + c0 + j));
        ccls[ob + p] = (float)(c0 + j);
      }
      ++p;
    }
    return;
  }
  // single label: class = FIRST index whose score equals the maximum score (the reference takes max over sigmoid
  // values, where distinct logits can tie after rounding / saturation); only logits near the maximum can tie
  const float margin = best > 15.f ? INFINITY : 1e-2f;
  int cls = 0x7fffffff;
  for (int j = 0; j < q && c0 + j < hp.nc; ++j) {
    const float v = __ldg(ar.row + 64 + c0 + j);
    if (v >= best - margin && sigmoidf_acc(v) == score) { cls = c0 + j; break; }
  }
  cls = min(cls, __shfl_xor_sync(gmask, cls, 1));
  cls = min(cls, __shfl_xor_sync(gmask, cls, 2));
  const int ppos = __shfl_sync(gmask, pos, lane & ~3);
  if (sub == 0 && ppos < cap) {
    cbox[ob + ppos] = box;
    cscore[ob + ppos] = score;
    ccls[ob + ppos] = (float)cls;
  }
}

// Single-label, A < 65536 (every real configuration): ONE pass.  Candidates are written in arbitrary order - each warp
// reserves its slots with one atomicAdd on the image's counter - and carry their anchor index (canchor); the sort key breaks
// score ties by ANCHOR, which reproduces the stable sort of the anchor-ordered list exactly.
// Two phases per 64-anchor chunk, because candidates are sparse (~20 % of the anchors at conf 0.25) but scattered: with the
// per-candidate work (accurate sigmoid, four 16-bin softmaxes, class search) done in place, nearly every warp executed it for
// one or two of its 8 anchors (90 us for 64 x 8400 anchors, instruction bound).  Phase 1 scans the class logits of all 64
// anchors and lists those whose maximum logit passes a conservative threshold; phase 2 runs the per-candidate work on the
// DENSE list (4 lanes per entry), so only ceil(n/8) warps execute it.
__global__ void __launch_bounds__(256)
decode_onepass_kernel(HeadParams hp, float conf, float logit_lo, int cap, float4* __restrict__ cbox, float* __restrict__ cscore,
                      float* __restrict__ ccls, int* __restrict__ canchor, int* __restrict__ ncand, int* __restrict__ ncand_raw) {
  __shared__ int s_list[kChunk];
  __shared__ float s_best[kChunk];
  __shared__ int s_n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int grp = threadIdx.x >> 2;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  {
    // ---- phase 1: this lane's quarter of the classes of anchor `grp` of the chunk
    const int a = blockIdx.x * kChunk + grp;
    const bool in_range = a < hp.A;
    float best = -INFINITY;
    if (in_range) {
      const AnchorRef ar = anchor_ref(hp, b, a);
      best = lane_class_max(ar.row + 64, hp.cls_per_lane, sub * hp.cls_per_lane, hp.nc);
    }
    best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
    best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
    // sigmoid(x) > conf  =>  x > logit(conf) - slack: everything that can be a candidate gets listed
    const bool pre = in_range && best > logit_lo;
    const unsigned bal = __ballot_sync(0xffffffffu, pre && sub == 0);
    if (bal) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pre && sub == 0) {
        const int slot = base + __popc(bal & ((1u << lane) - 1u));
        s_list[slot] = grp;
        s_best[slot] = best;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: list entry `grp`, 4 lanes per entry; whole warps beyond the list leave
  const int n = s_n;
  if (warp * 8 >= n) return;
  const bool have = grp < n;
  const int e = have ? grp : 0;
  const int a = blockIdx.x * kChunk + s_list[e];
  const float best = s_best[e];
  const AnchorRef ar = anchor_ref(hp, b, a);
  const float score = sigmoidf_acc(best);
  const bool cand = have && score > conf;
  const unsigned cb = __ballot_sync(0xffffffffu, cand && sub == 0);
  int wbase = 0;
  if (cb) {
    if (lane == 0) {
      wbase = atomicAdd(&ncand[b], __popc(cb));
      if (ncand_raw) atomicAdd(&ncand_raw[b], __popc(cb));
    }
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
  }
  if (!cand) return;  // whole 4-lane groups leave together
  const unsigned gmask = 0xFu << (lane & ~3);
  const int pos = wbase + __popc(cb & ((1u << (lane & ~3)) - 1u));
  const float4 box = xywh2xyxy_rn(decode_box(ar, sub, lane, gmask));
  // class = FIRST index whose score equals the maximum score (see decode_compact_kernel)
  const int q = hp.cls_per_lane, c0 = sub * q;
  const float margin = best > 15.f ? INFINITY : 1e-2f;
  int cls = 0x7fffffff;
  {
    // this lane's quarter of the class logits, all loads issued before the first use (v1 walked them one dependent load
    // at a time: the hottest loop of the kernel in the ncu source view)
    float4 t[kMaxClsVec];
#pragma unroll
    for (int j = 0; j < kMaxClsVec; ++j)
      t[j] = (4 * j < q && c0 + 4 * j < hp.nc) ? __ldg(reinterpret_cast<const float4*>(ar.row + 64 + c0) + j) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int j = kMaxClsVec - 1; j >= 0; --j) {   // descending, so that the FIRST matching index wins
      const float v[4] = {t[j].x, t[j].y, t[j].z, t[j].w};
#pragma unroll
      for (int k = 3; k >= 0; --k) {
        const int c = c0 + 4 * j + k;
        if (4 * j + k < q && c < hp.nc && v[k] >= best - margin && sigmoidf_acc(v[k]) == score) cls = c;
      }
    }
  }
  cls = min(cls, __shfl_xor_sync(gmask, cls, 1));
  cls = min(cls, __shfl_xor_sync(gmask, cls, 2));
  if (sub == 0 && pos < cap) {
    const size_t ob = (size_t)b * cap;
    cbox[ob + pos] = box;
    cscore[ob + pos] = score;
    ccls[ob + pos] = (float)cls;
    canchor[ob + pos] = a;
  }
}

// Single-label with class-emit convs (y11_plan_set_cls_emit): the conv epilogues already reduced every anchor's class logits
// to (maximum logit, class) and listed the anchors above the conservative threshold.  This is phase 2 of decode_onepass_kernel
// on that list: accurate sigmoid, conf test, DFL box of the candidates (4 lanes per entry), one atomicAdd per warp for the
// output slots.  Same arithmetic, same helpers -> bit-identical candidates (the list order is free: the sort key breaks
// score ties by anchor).
__global__ void __launch_bounds__(256)
decode_list_kernel(HeadParams hp, float conf, int cap, const int4* __restrict__ list, const int* __restrict__ list_count, int list_cap,
                   float4* __restrict__ cbox, float* __restrict__ cscore, float* __restrict__ ccls, int* __restrict__ canchor,
                   int* __restrict__ ncand, int* __restrict__ ncand_raw) {
  const int b = blockIdx.y;
  const int n = min(list_count[b], list_cap);
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 3;
  const int e0 = blockIdx.x * kChunk + (threadIdx.x >> 2);
  if (blockIdx.x * kChunk + (threadIdx.x >> 5) * 8 >= n) return;  // whole warps beyond the list leave
  const bool have = e0 < n;
  const int4 ent = list[(size_t)b * list_cap + (have ? e0 : 0)];
  const int a = ent.x;
  const float best = __int_as_float(ent.z);
  const AnchorRef ar = anchor_ref(hp, b, a);
  const float score = sigmoidf_acc(best);
  const bool cand = have && score > conf;
  const unsigned cb = __ballot_sync(0xffffffffu, cand && sub == 0);
  int wbase = 0;
  if (cb) {
    if (lane == 0) {
      wbase = atomicAdd(&ncand[b], __popc(cb));
      if (ncand_raw) atomicAdd(&ncand_raw[b], __popc(cb));
    }
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
  }
  if (!cand) return;  // whole 4-lane groups leave together
  const unsigned gmask = 0xFu << (lane & ~3);
  const int pos = wbase + __popc(cb & ((1u << (lane & ~3)) - 1u));
  const float4 box = xywh2xyxy_rn(decode_box(ar, sub, lane, gmask));
  if (sub == 0 && pos < cap) {
    const size_t ob = (size_t)b * cap;
    cbox[ob + pos] = box;
    cscore[ob + pos] = score;
    ccls[ob + pos] = (float)ent.y;
    canchor[ob + pos] = a;
  }
}

// exclusive scan of chunk counts, one CTA per image
__global__ void __launch_bounds__(1024) scan_chunks_kernel(const int* __restrict__ chunk_cnt, int* __restrict__ chunk_off, int nchunks,
                                                           int cap, int* __restrict__ ncand, int* __restrict__ ncand_raw) {
  __shared__ int s[1024];
  const int b = blockIdx.x;
  int carry = 0;
  for (int c0 = 0; c0 < nchunks; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    const int v = i < nchunks ? chunk_cnt[b * nchunks + i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nchunks) chunk_off[b * nchunks + i] = carry + s[threadIdx.x] - v;
    carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ncand[b] = min(carry, cap);
    if (ncand_raw) ncand_raw[b] = carry;
  }
}

// ------------------------------------------------------------------------------------------------ sort + NMS
__device__ __forceinline__ float iou_rn(const float4 a, float aarea, const float4 b, float barea) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
}
// `iou_rn(a, b) > thr` with the division skipped for disjoint boxes.  Exact: w <= 0 or h <= 0 makes inter = +0, so the
// quotient is 0 (or NaN for two empty boxes) and `> thr` is false for every thr >= 0 - the result the full formula gives.
// With class-aware NMS (boxes offset by cls * 7680) every cross-class pair is disjoint, i.e. ~79/80 of all pairs take the
// short path (the IEEE division alone is ~30 instructions; the NMS kernel was bound by it: ~64 K IoUs per 256-box tile).
__device__ __forceinline__ bool iou_gt(const float4 a, float aarea, const float4 b, float barea, float thr) {
  if (thr < 0.0f) return iou_rn(a, aarea, b, barea) > thr;  // degenerate threshold: 0 > thr is true, no short cut
  const float xx1 = fmaxf(a.x, b.x), xx2 = fminf(a.z, b.z);
  if (!(xx2 > xx1)) return false;
  const float yy1 = fmaxf(a.y, b.y), yy2 = fminf(a.w, b.w);
  if (!(yy2 > yy1)) return false;
  const float inter = __fmul_rn(__fsub_rn(xx2, xx1), __fsub_rn(yy2, yy1));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter)) > thr;
}

struct NmsArgs {
  const float4* cbox;    // [B, cap] xyxy without class offset
  const float* cscore;   // [B, cap]
  const float* ccls;     // [B, cap]
  const int* ncand;      // [B]
  const int* canchor;    // [B, cap] anchor index of each candidate (single-pass compaction) or nullptr (list in anchor order)
  unsigned long long* keys;  // [B, keys_stride] scratch (used when n > kSmemKeys)
  float4* kbox;          // [B, max_det] kept boxes (with class offset)
  float* karea;          // [B, max_det]
  int cap, keys_stride;
  float iou_thr;         // largest float <= (double) iou: ovr_f32 > thr  <=>  (double)ovr_f32 > iou
  float class_offset;    // max_wh, or 0 when agnostic
  int max_det, max_nms;
  const float* scale;    // [B,5] gain,pad_x,pad_y,w0,h0 or nullptr
  float* out_det;        // [B, max_det, 6] or nullptr  (may point into a PEER device's memory: result push)
  int* out_keep;         // [B, max_det] or nullptr
  int* out_count;        // [B]
  int* done_counter;     // result push: device-local counter of finished CTAs (zero between launches) or nullptr
  unsigned* signal;      // result push: system-scope counter bumped once per launch after every result write is visible
  // progressive top-K (see select_* kernels): the candidate list handed in is the score-ordered PREFIX of a longer list
  const int* skip;       // [B] or nullptr: image already final (an earlier stage kept max_det boxes) -> leave its results alone
  int* done_out;         // [B] or nullptr: 1 = this stage's result is final for the image
  const int* total;      // [B] or nullptr: length of the full candidate list the prefix was selected from
};

__global__ void __launch_bounds__(kNmsThreads) sort_nms_kernel(NmsArgs g) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ float4 t_box[kTile];
  __shared__ float t_area[kTile];
  __shared__ int t_idx[kTile];
  __shared__ unsigned long long t_mask[kTile][kTile / 64];
  __shared__ unsigned int t_dead[kTile / 32];
  __shared__ unsigned int t_nz[kTile / 32];  // bit c: candidate c of the tile suppresses at least one later candidate
  __shared__ int t_keep[kTile];
  __shared__ int s_nk;
  // kept boxes (with class offset) and their areas: every later tile tests its candidates against them, 4 threads per
  // candidate walking the list serially - from shared memory when max_det allows (it is 300), else from the global scratch
  __shared__ float4 s_kbox[kKeptSmem];
  __shared__ float s_karea[kKeptSmem];

  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t ob = (size_t)b * g.cap;
  const bool skipped = g.skip && g.skip[b];   // CTA-uniform: an earlier stage already produced this image's final result
  const bool overflow = g.ncand[b] < 0;        // the selection did not fit its buffer (huge score tie): a later stage takes it
  const int n = (skipped || overflow) ? 0 : min(g.ncand[b], g.cap);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  unsigned long long* keys = (np2 <= kSmemKeys) ? s_keys : g.keys + (size_t)b * g.keys_stride;

  // 1. keys: descending score, ascending index on ties == stable descending sort of the anchor-ordered list
  for (int i = tid; i < np2; i += kNmsThreads) {
    unsigned long long k = ~0ull;
    if (i < n) {
      unsigned u = __float_as_uint(g.cscore[ob + i]);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone map float -> unsigned
      // ties: list position when the list is in anchor order, else (anchor, slot) with the slot in the low 16 bits
      k = ((unsigned long long)(~u) << 32) | (g.canchor ? (((unsigned)g.canchor[ob + i] << 16) | (unsigned)i) : (unsigned)i);
    }
    keys[i] = k;
  }
  __syncthreads();
  // 2. bitonic sort, ascending keys
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1, lj = 31 - __clz(k >> 1); j > 0; j >>= 1, --lj) {
      for (int t = tid; t < (np2 >> 1); t += kNmsThreads) {
        const int i = ((t >> lj) << (lj + 1)) + (t & (j - 1));  // 2*j*(t/j) + t%j, j = 2^lj (the division was 20 instructions)
        const int l = i + j;
        const unsigned long long x = keys[i], y = keys[l];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { keys[i] = y; keys[l] = x; }
      }
      __syncthreads();
    }
  }
  // 3. greedy NMS over score order, 256 candidates per tile
  const int K = min(n, g.max_nms);
  float4* kbox = g.max_det <= kKeptSmem ? s_kbox : g.kbox + (size_t)b * g.max_det;
  float* karea = g.max_det <= kKeptSmem ? s_karea : g.karea + (size_t)b * g.max_det;
  if (tid == 0) s_nk = 0;
  __syncthreads();
  for (int c0 = 0; c0 < K; c0 += kTile) {
    const int nk0 = s_nk;
    if (nk0 >= g.max_det) break;
    const int cnt = min(kTile, K - c0);
    if (tid < kTile) {
      if (tid < cnt) {
        const int idx = (int)(keys[c0 + tid] & (g.canchor ? 0xffffu : 0xffffffffu));
        float4 bx = g.cbox[ob + idx];
        const float off = __fmul_rn(g.ccls[ob + idx], g.class_offset);
        bx.x = __fadd_rn(bx.x, off); bx.y = __fadd_rn(bx.y, off); bx.z = __fadd_rn(bx.z, off); bx.w = __fadd_rn(bx.w, off);
        t_box[tid] = bx;
        t_area[tid] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        t_idx[tid] = idx;
      }
      if (tid < kTile / 32) { t_dead[tid] = 0u; t_nz[tid] = 0u; }
    }
    __syncthreads();
    {
      const int c = tid & (kTile - 1), part = tid >> 8;  // 4 threads per candidate
      if (c < cnt) {
        // (a) against boxes kept by earlier tiles
        const float4 bx = t_box[c];
        const float ar = t_area[c];
        bool dead = false;
        for (int k = part; k < nk0 && !dead; k += kNmsThreads / kTile) dead = iou_gt(kbox[k], karea[k], bx, ar, g.iou_thr);
        if (dead) atomicOr(&t_dead[c >> 5], 1u << (c & 31));
        // (b) bit row of this tile: later candidates j this one would suppress
        unsigned long long bits = 0ull;
        const int j0 = part * 64;
        for (int j = max(j0, c + 1); j < min(j0 + 64, cnt); ++j)
          if (iou_gt(bx, ar, t_box[j], t_area[j], g.iou_thr)) bits |= 1ull << (j - j0);
        t_mask[c][part] = bits;
        if (bits) atomicOr(&t_nz[c >> 5], 1u << (c & 31));  // rare: most candidates suppress nobody
      }
    }
    __syncthreads();
    // (c) suppression sweep over the tile's candidates in score order.  v1 was a warp-level loop (ballot / shuffle / ffs /
    //     shared-memory load on one dependent chain, ~300 cycles per kept box: 55 % of the kernel's stall samples were the
    //     other 31 warps waiting for it).  Now ONE thread walks the four 64-bit words in registers; a kept candidate whose
    //     bit row is empty (it suppresses nobody - the common case) costs a find-first-set and a store, only the others load
    //     their row.  Rows only contain LATER candidates, so finishing word w before word w+1 is exact.
    if (tid == 0) {
      constexpr int kW = kTile / 64;
      unsigned long long removed[kW], todo[kW], nz[kW];
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        const unsigned long long dead = (unsigned long long)t_dead[2 * w] | ((unsigned long long)t_dead[2 * w + 1] << 32);
        const int lo = w * 64;
        const unsigned long long valid = cnt >= lo + 64 ? ~0ull : (cnt > lo ? ((1ull << (cnt - lo)) - 1ull) : 0ull);
        removed[w] = dead | ~valid;
        todo[w] = valid;
        nz[w] = (unsigned long long)t_nz[2 * w] | ((unsigned long long)t_nz[2 * w + 1] << 32);
      }
      int nk = nk0;
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        unsigned long long avail = ~removed[w] & todo[w];
        while (avail != 0ull && nk < g.max_det) {
          const int bit = __ffsll((long long)avail) - 1;
          const int c = w * 64 + bit;
          t_keep[nk - nk0] = c;
          ++nk;
          avail &= avail - 1ull;
          if ((nz[w] >> bit) & 1ull) {
            avail &= ~t_mask[c][w];
#pragma unroll
            for (int ww = w + 1; ww < kW; ++ww) removed[ww] |= t_mask[c][ww];
          }
        }
      }
      s_nk = nk;
    }
    __syncthreads();
    // (d) publish the boxes kept in this tile: one thread per kept box.  (Doing this inside the sweep put three dependent
    //     global loads and eight global stores on lane 0's critical path for every one of the up to max_det kept boxes.)
    for (int k = nk0 + tid; k < s_nk; k += kNmsThreads) {
      const int c = t_keep[k - nk0];
      kbox[k] = t_box[c];
      karea[k] = t_area[c];
      const int idx = t_idx[c];
      if (g.out_keep) g.out_keep[(size_t)b * g.max_det + k] = idx;
      if (g.out_det) {
        float4 bx = g.cbox[ob + idx];
        if (g.scale) {
          const float* sc = g.scale + (size_t)b * 5;
          const float gain = sc[0], px = sc[1], py = sc[2], w0 = sc[3], h0 = sc[4];
          bx.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, px), gain), 0.f), w0);
          bx.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, py), gain), 0.f), h0);
          bx.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, px), gain), 0.f), w0);
          bx.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, py), gain), 0.f), h0);
        }
        float* o = g.out_det + ((size_t)b * g.max_det + k) * 6;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        o[4] = g.cscore[ob + idx];
        o[5] = g.ccls[ob + idx];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (!skipped && !overflow) g.out_count[b] = s_nk;
    // final iff max_det boxes were kept, or the prefix was the whole list, or it reached the max_nms ranks NMS may look at
    if (g.done_out)
      g.done_out[b] = skipped ? 1 : overflow ? 0 : (s_nk >= g.max_det || !g.total || n >= min(g.total[b], g.max_nms)) ? 1 : 0;
  }
  if (g.signal) {
    // Result push: out_det / out_count live in another GPU's memory (NVLink peer mapping).  Every CTA makes its writes
    // visible system-wide, the LAST one to finish bumps the consumer's signal - the consumer never sees the counter move
    // before all B images of this launch have landed (threadFenceReduction pattern; fences are cumulative over the barrier).
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      if (atomicAdd(g.done_counter, 1) == (int)gridDim.x - 1) {
        *g.done_counter = 0;  // re-armed for the next launch (launches of one pipeline are stream-ordered)
        __threadfence_system();
        atomicAdd_system(g.signal, 1u);
      }
    }
  }
}

// Consumer side of the result push: lane i parks until signals[i] has reached `target` (wrap-safe), bounded.
__global__ void wait_signals_kernel(const volatile unsigned* signals, int n, unsigned target, int* err_flag) {
  const int i = threadIdx.x;
  if (i < n) {
    const long long t0 = clock64();
    while ((int)(signals[i] - target) < 0) {
      __nanosleep(256);
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer died
        if (err_flag) *reinterpret_cast<volatile int*>(err_flag) = 201;
        break;
      }
    }
  }
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------------ progressive top-K
// Validation runs NMS at conf 0.001 with multi-label candidates: up to nc * A = 2.7 M (anchor, class) rows per 1280x1280 image.
// Sorting all of them in one CTA took 300 ms per batch (profiles/r02_post_scale_m1280.json); greedy NMS in score order only
// ever needs the score-ordered PREFIX it walks before max_det boxes are kept (and never more than max_nms rows).  So:
//   stage 1: exact top-K1 (K1 = 8192) by a multi-CTA radix select on the monotone score keys (three histogram passes of
//            11 + 11 + 10 bits), ORDER-PRESERVING compaction of every row with key >= threshold (ties included, so the prefix
//            is exactly the first `count` rows of the stable descending sort), then the usual sort + NMS CTA on <= 16 K rows;
//   stage 2: only for images whose stage 1 kept fewer than max_det boxes although rows were left: the same with K = max_nms.
// Results are bit-identical to sorting everything (same keys, same tie order, same NMS arithmetic).
constexpr int kSelBits0 = 11, kSelBits1 = 11, kSelBits2 = 10;
constexpr int kSelBins = 2048;
constexpr int kSelChunk = 1024;   // candidates per compaction chunk

__device__ __forceinline__ unsigned score_key(float s) {
  const unsigned u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone map float -> unsigned (same as the sort key's high word)
}

struct SelState {   // per image
  unsigned prefix;  // key bits decided so far (left aligned at the current pass)
  int remaining;    // rows still to take among the keys that match the prefix
  int all;          // 1: the list has at most K rows - take everything
};

__global__ void select_init_kernel(const int* __restrict__ ncand, const int* __restrict__ skip, int K, int cap, SelState* __restrict__ st,
                                   int* __restrict__ hist) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hist[b * kSelBins + i] = 0;
  if (threadIdx.x == 0) {
    const int n = min(ncand[b], cap);
    st[b].prefix = 0u;
    st[b].remaining = min(K, n);
    st[b].all = (n <= K || (skip && skip[b])) ? 1 : 0;
  }
}

// pass p: histogram of the next bit field over the rows whose higher bits equal the prefix
__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ cscore, const int* __restrict__ ncand,
                                                          const int* __restrict__ skip, const SelState* __restrict__ st, int cap,
                                                          int shift, int bits, int hi_shift, int* __restrict__ hist) {
  __shared__ int s_h[kSelBins];
  const int b = blockIdx.y;
  if ((skip && skip[b]) || st[b].all) return;
  const int n = min(ncand[b], cap);
  const int i0 = blockIdx.x * 256 * 16;
  if (i0 >= n) return;
  for (int i = threadIdx.x; i < kSelBins; i += 256) s_h[i] = 0;
  __syncthreads();
  const unsigned prefix = st[b].prefix, mask = (1u << bits) - 1u;
  const float* sc = cscore + (size_t)b * cap;
  for (int i = i0 + threadIdx.x; i < min(n, i0 + 256 * 16); i += 256) {
    const unsigned u = score_key(sc[i]);
    if (hi_shift >= 32 || (u >> hi_shift) == prefix) atomicAdd(&s_h[(u >> shift) & mask], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (1 << bits); i += 256)
    if (s_h[i]) atomicAdd(&hist[b * kSelBins + i], s_h[i]);
}

// pick the bin in which the `remaining`-th largest key lies; one CTA per image; clears the histogram for the next pass
__global__ void __launch_bounds__(1024) select_find_kernel(SelState* __restrict__ st, int bits, int* __restrict__ hist) {
  __shared__ int s_c[kSelBins];
  const int b = blockIdx.x, nb = 1 << bits;
  if (st[b].all) return;
  for (int i = threadIdx.x; i < nb; i += 1024) s_c[i] = hist[b * kSelBins + i];
  __syncthreads();
  if (threadIdx.x == 0) {   // 2048 bins: a serial walk from the top costs ~2 us, once per pass per image
    int need = st[b].remaining, bin = nb - 1, above = 0;
    for (; bin > 0; --bin) {
      if (above + s_c[bin] >= need) break;
      above += s_c[bin];
    }
    st[b].prefix = (st[b].prefix << bits) | (unsigned)bin;
    st[b].remaining = need - above;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += 1024) hist[b * kSelBins + i] = 0;
}

// MODE 0: rows with key >= threshold per chunk.  MODE 1: copy them, in list order, behind the chunk's offset.
template <int MODE>
__global__ void __launch_bounds__(256) select_compact_kernel(const float4* __restrict__ cbox, const float* __restrict__ cscore,
                                                             const float* __restrict__ ccls, const int* __restrict__ ncand,
                                                             const int* __restrict__ skip, const SelState* __restrict__ st, int cap,
                                                             int nchunks, int cap2, int* __restrict__ chunk_cnt,
                                                             const int* __restrict__ chunk_off, float4* __restrict__ sbox,
                                                             float* __restrict__ sscore, float* __restrict__ scls) {
  __shared__ int s_warp[8];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (skip && skip[b]) {
    if (MODE == 0 && threadIdx.x == 0) chunk_cnt[b * nchunks + blockIdx.x] = 0;
    return;
  }
  const int n = min(ncand[b], cap);
  const unsigned thr = st[b].all ? 0u : st[b].prefix;
  const size_t ob = (size_t)b * cap;
  int run = MODE == 1 ? chunk_off[b * nchunks + blockIdx.x] : 0;
  int total = 0;
  for (int r = 0; r < kSelChunk / 256; ++r) {   // 4 rounds of 256 rows, in order
    const int i = blockIdx.x * kSelChunk + r * 256 + threadIdx.x;
    const bool take = i < n && score_key(cscore[ob + i]) >= thr;
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, round_total = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += s_warp[w];
      round_total += s_warp[w];
    }
    if (MODE == 1 && take) {
      const int pos = run + before + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap2) {
        const size_t o2 = (size_t)b * cap2 + pos;
        sbox[o2] = cbox[ob + i];
        sscore[o2] = cscore[ob + i];
        scls[o2] = ccls[ob + i];
      }
    }
    run += round_total;
    total += round_total;
    __syncthreads();
  }
  if (MODE == 0 && threadIdx.x == 0) chunk_cnt[b * nchunks + blockIdx.x] = total;
}

// after the scan: images whose selection overflowed the compact buffer (a huge score tie) go to the next stage / full path
__global__ void select_finish_kernel(int* __restrict__ sel_n, const int* __restrict__ sel_raw, int cap2, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && sel_raw[b] > cap2) sel_n[b] = -1;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

int fill_head(const y11_head_desc* hd, HeadParams* hp) {
  Y11_REQUIRE(hd->nl >= 1 && hd->nl <= 3, "postprocess: nl=%d", hd->nl);
  Y11_REQUIRE(hd->nc >= 1 && hd->nc <= 128, "postprocess: nc=%d unsupported (max 128)", hd->nc);
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    hp->head[l] = l < hd->nl ? hd->head[l] : nullptr;
    hp->hl[l] = l < hd->nl ? hd->hl[l] : 0;
    hp->wl[l] = l < hd->nl ? hd->wl[l] : 1;
    hp->stride[l] = l < hd->nl ? hd->stride[l] : 0.f;
    hp->off[l] = off;
    off += hp->hl[l] * (l < hd->nl ? hd->wl[l] : 0);
  }
  hp->off[3] = off;
  hp->nl = hd->nl; hp->B = hd->B; hp->nc = hd->nc; hp->A = off; hp->no = hd->row_stride;
  hp->cls_per_lane = 4 * ((hd->nc + 15) / 16);
  Y11_REQUIRE(hd->row_stride >= 64 + 16 * ((hd->nc + 15) / 16) && hd->row_stride % 4 == 0,
              "postprocess: row_stride %d must be >= 64 + nc rounded up to 16 and a multiple of 4", hd->row_stride);
  return 0;
}

float thr_round_down(double iou) {
  float f = (float)iou;
  if ((double)f > iou) f = nextafterf(f, -INFINITY);
  return f;
}

struct Workspace {
  float4* cbox; float* cscore; float* ccls; int* canchor; int* chunk_cnt; int* chunk_off; int* ncand;
  unsigned long long* keys; float4* kbox; float* karea;
  int cap, keys_stride, nchunks;
  // progressive top-K (lists longer than kSelMinCap rows)
  float4* sbox; float* sscore; float* scls; int* sel_n; int* sel_raw; int* sel_cnt; int* sel_off; int* hist; int* done_a; int* done_b;
  SelState* sel_state;
  int sel_chunks, cap2_max;
  size_t total;
};
constexpr int kSelMinCap = 32768;     // shorter candidate lists are sorted directly
constexpr int kSelK1 = 8192;          // stage-1 prefix length
constexpr int kSelCap2A = 16384;      // stage-1 compact buffer (room for ties; sorted in shared memory)
constexpr int kSelCap2B = 65536;      // stage-2 compact buffer (K = max_nms = 30000 by default)

// Candidate capacity per image.  Multi-label (val mode, conf 0.001): room for EVERY (anchor, class) pair up to 4 M, so
// that "keep the max_nms best by score" (ultralytics non_max_suppression) is exact - the list is sorted by score before
// it is cut.  (The first version capped the list at 128 K entries in anchor order, which silently dropped the
// high-resolution levels' successors whenever nearly every pair passed the threshold.)
int candidate_cap(int A, int nc, int multi_label) {
  if (!multi_label) return A;
  const long long full = (long long)A * nc;
  return (int)(full < (1ll << 22) ? full : (1ll << 22));
}
int kept_cap_of(int cap) { return cap < (1 << 15) ? cap : (1 << 15); }

void carve(Workspace* w, void* base, int B, int cap, int nchunks, int kept_cap) {
  size_t o = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) { char* r = p ? p + o : nullptr; o += align_up(bytes); return r; };
  w->cap = cap; w->nchunks = nchunks; w->keys_stride = next_pow2(cap);
  w->cbox = (float4*)take((size_t)B * cap * 16);
  w->cscore = (float*)take((size_t)B * cap * 4);
  w->ccls = (float*)take((size_t)B * cap * 4);
  w->canchor = (int*)take((size_t)B * cap * 4);
  w->chunk_cnt = (int*)take((size_t)B * nchunks * 4);
  w->chunk_off = (int*)take((size_t)B * nchunks * 4);
  w->ncand = (int*)take((size_t)B * 4);
  w->keys = (unsigned long long*)take(w->keys_stride > kSmemKeys ? (size_t)B * w->keys_stride * 8 : 0);
  w->kbox = (float4*)take((size_t)B * kept_cap * 16);
  w->karea = (float*)take((size_t)B * kept_cap * 4);
  const bool sel = cap > kSelMinCap;
  w->cap2_max = sel ? kSelCap2B : 0;
  w->sel_chunks = sel ? y11_ceil_div(cap, kSelChunk) : 0;
  w->sbox = (float4*)take((size_t)B * w->cap2_max * 16);
  w->sscore = (float*)take((size_t)B * w->cap2_max * 4);
  w->scls = (float*)take((size_t)B * w->cap2_max * 4);
  w->sel_cnt = (int*)take((size_t)B * w->sel_chunks * 4);
  w->sel_off = (int*)take((size_t)B * w->sel_chunks * 4);
  w->hist = (int*)take(sel ? (size_t)B * kSelBins * 4 : 0);
  w->sel_state = (SelState*)take(sel ? (size_t)B * sizeof(SelState) : 0);
  w->sel_n = (int*)take((size_t)B * 4);
  w->sel_raw = (int*)take((size_t)B * 4);
  w->done_a = (int*)take((size_t)B * 4);
  w->done_b = (int*)take((size_t)B * 4);
  w->total = o;
}

// One stage of the progressive top-K: exact selection of the K best rows (ties at the threshold included) of every image not
// yet final, order-preserving compaction into the compact buffers (stride cap2), then sort + NMS on that prefix.
int launch_sort_nms(const NmsArgs& a, int B, cudaStream_t s);
int topk_stage(const Workspace& w, const NmsArgs& full, int B, int K, int cap2, const int* skip, int* done_out, cudaStream_t s) {
  select_init_kernel<<<B, 256, 0, s>>>(w.ncand, skip, K, w.cap, w.sel_state, w.hist);
  const dim3 hgrid((unsigned)y11_ceil_div(w.cap, 256 * 16), (unsigned)B);
  const int bits[3] = {kSelBits0, kSelBits1, kSelBits2};
  int shift = 32;
  for (int pass = 0; pass < 3; ++pass) {
    const int hi_shift = shift;            // bits above this pass's field must equal the prefix (32 = no constraint)
    shift -= bits[pass];
    select_hist_kernel<<<hgrid, 256, 0, s>>>(w.cscore, w.ncand, skip, w.sel_state, w.cap, shift, bits[pass], hi_shift, w.hist);
    select_find_kernel<<<B, 1024, 0, s>>>(w.sel_state, bits[pass], w.hist);
  }
  const dim3 cgrid((unsigned)w.sel_chunks, (unsigned)B);
  select_compact_kernel<0><<<cgrid, 256, 0, s>>>(w.cbox, w.cscore, w.ccls, w.ncand, skip, w.sel_state, w.cap, w.sel_chunks, cap2, w.sel_cnt,
                                                 w.sel_off, w.sbox, w.sscore, w.scls);
  scan_chunks_kernel<<<B, 1024, 0, s>>>(w.sel_cnt, w.sel_off, w.sel_chunks, cap2, w.sel_n, w.sel_raw);
  select_compact_kernel<1><<<cgrid, 256, 0, s>>>(w.cbox, w.cscore, w.ccls, w.ncand, skip, w.sel_state, w.cap, w.sel_chunks, cap2, w.sel_cnt,
                                                 w.sel_off, w.sbox, w.sscore, w.scls);
  select_finish_kernel<<<y11_ceil_div(B, 256), 256, 0, s>>>(w.sel_n, w.sel_raw, cap2, B);
  Y11_CHECK_CUDA(cudaGetLastError());
  NmsArgs a = full;
  a.cbox = w.sbox; a.cscore = w.sscore; a.ccls = w.scls; a.ncand = w.sel_n; a.canchor = nullptr;
  a.cap = cap2; a.keys_stride = next_pow2(cap2);
  a.skip = skip; a.done_out = done_out; a.total = w.ncand;
  a.done_counter = nullptr; a.signal = nullptr;   // the result push is signalled by the LAST launch of the call only
  return launch_sort_nms(a, B, s);
}

int launch_sort_nms(const NmsArgs& a, int B, cudaStream_t s) {
  Y11_OPT_IN_SMEM(sort_nms_kernel, kSmemKeys * 8);
  const int np2 = next_pow2(a.cap);
  // the kernel sorts in shared memory whenever THIS image's count fits, whatever the capacity is
  const size_t smem = (size_t)(np2 <= kSmemKeys ? np2 : kSmemKeys) * 8;
  sort_nms_kernel<<<B, kNmsThreads, smem, s>>>(a);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int y11_decode_dense(y11_handle, const y11_head_desc* hd, float* y, y11_stream s) {
  HeadParams hp;
  if (int e = fill_head(hd, &hp)) return e;
  dim3 grid((unsigned)y11_ceil_div(hp.A, kChunk), (unsigned)hp.B);
  decode_dense_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(s)>>>(hp, y);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t y11_postprocess_workspace(int B, int A, int nc, int multi_label, int max_nms) {
  (void)max_nms;
  Workspace w;
  const int cap = candidate_cap(A, nc, multi_label);
  carve(&w, nullptr, B, cap, y11_ceil_div(A, kChunk), kept_cap_of(cap));
  return w.total;
}

struct EmitList {  // pre-candidate lists of class-emit convs (y11_detect_postprocess_list)
  const int4* list;
  const int* count;
  int cap;
};

static int postprocess_impl(const y11_head_desc* hd, const y11_nms_params* p, const float* scale, float* out_det, int32_t* out_count,
                            int32_t* out_ncand, void* workspace, size_t workspace_bytes, const y11_push* push, float* ms2,
                            cudaStream_t s, const EmitList* el = nullptr) {
  HeadParams hp;
  if (int e = fill_head(hd, &hp)) return e;
  Y11_REQUIRE(p->max_det >= 1, "postprocess: max_det=%d", p->max_det);
  const int cap = candidate_cap(hp.A, hp.nc, p->multi_label);
  Y11_REQUIRE(p->max_det <= kept_cap_of(cap), "postprocess: max_det=%d exceeds capacity %d", p->max_det, kept_cap_of(cap));
  const int nchunks = y11_ceil_div(hp.A, kChunk);
  Workspace w;
  carve(&w, workspace, hp.B, cap, nchunks, kept_cap_of(cap));
  Y11_REQUIRE(workspace && workspace_bytes >= w.total, "postprocess: workspace %zu < required %zu", workspace_bytes, w.total);
  // conservative logit pre-filter for the multi-label sigmoid test: sigmoid(x) > conf  =>  x > logit(conf) - slack
  const double cc = std::min(std::max((double)p->conf, 1e-30), 1.0 - 1e-9);
  const float logit_lo = (float)(log(cc / (1.0 - cc)) - 1e-2);
  dim3 grid((unsigned)nchunks, (unsigned)hp.B);
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (ms2) {
    for (auto& e : ev) Y11_CHECK_CUDA(cudaEventCreate(&e));
    Y11_CHECK_CUDA(cudaEventRecord(ev[0], s));
  }
  // single-label with A < 65536 (every real configuration): one pass with atomic slot reservation; the candidate count
  // cannot exceed the capacity (one candidate per anchor at most), so nothing depends on the arrival order
  const bool one_pass = !p->multi_label && hp.A < 65536 && cap >= hp.A;
  if (el) {
    Y11_REQUIRE(one_pass, "postprocess_list: single-label with fewer than 65536 anchors only");
    Y11_REQUIRE(el->list && el->count && el->cap > 0, "postprocess_list: null list");
    Y11_CHECK_CUDA(cudaMemsetAsync(w.ncand, 0, (size_t)hp.B * sizeof(int), s));
    if (out_ncand) Y11_CHECK_CUDA(cudaMemsetAsync(out_ncand, 0, (size_t)hp.B * sizeof(int), s));
    dim3 lgrid((unsigned)y11_ceil_div(std::min(el->cap, hp.A), kChunk), (unsigned)hp.B);
    decode_list_kernel<<<lgrid, 256, 0, s>>>(hp, p->conf, cap, el->list, el->count, el->cap, w.cbox, w.cscore, w.ccls, w.canchor,
                                             w.ncand, out_ncand);
  } else if (one_pass) {
    Y11_CHECK_CUDA(cudaMemsetAsync(w.ncand, 0, (size_t)hp.B * sizeof(int), s));
    if (out_ncand) Y11_CHECK_CUDA(cudaMemsetAsync(out_ncand, 0, (size_t)hp.B * sizeof(int), s));
    decode_onepass_kernel<<<grid, 256, 0, s>>>(hp, p->conf, logit_lo, cap, w.cbox, w.cscore, w.ccls, w.canchor, w.ncand, out_ncand);
  } else {
    decode_compact_kernel<0><<<grid, 256, 0, s>>>(hp, p->conf, logit_lo, p->multi_label, cap, nchunks, w.chunk_cnt, w.chunk_off, w.cbox,
                                                  w.cscore, w.ccls, nullptr, nullptr, nullptr);
    scan_chunks_kernel<<<hp.B, 1024, 0, s>>>(w.chunk_cnt, w.chunk_off, nchunks, cap, w.ncand, out_ncand);
    decode_compact_kernel<1><<<grid, 256, 0, s>>>(hp, p->conf, logit_lo, p->multi_label, cap, nchunks, w.chunk_cnt, w.chunk_off, w.cbox,
                                                  w.cscore, w.ccls, nullptr, nullptr, nullptr);
  }
  Y11_CHECK_CUDA(cudaGetLastError());
  if (ms2) Y11_CHECK_CUDA(cudaEventRecord(ev[1], s));
  NmsArgs a;
  a.cbox = w.cbox; a.cscore = w.cscore; a.ccls = w.ccls; a.ncand = w.ncand; a.keys = w.keys; a.kbox = w.kbox; a.karea = w.karea;
  a.canchor = one_pass ? w.canchor : nullptr;
  a.cap = cap; a.keys_stride = w.keys_stride;
  a.iou_thr = thr_round_down(p->iou);
  a.class_offset = p->agnostic ? 0.0f : (float)p->max_wh;
  a.max_det = p->max_det; a.max_nms = p->max_nms;
  a.scale = scale; a.out_det = out_det; a.out_keep = nullptr; a.out_count = out_count;
  a.skip = nullptr; a.done_out = nullptr; a.total = nullptr;
  a.done_counter = nullptr; a.signal = nullptr;
  if (!one_pass && cap > kSelMinCap) {
    // long lists (multi-label validation, A >= 65536): sort + NMS on score-ordered prefixes instead of on everything
    const int k1 = std::min(kSelK1, p->max_nms);
    if (int e = topk_stage(w, a, hp.B, k1, kSelCap2A, nullptr, w.done_a, s)) return e;
    if (int e = topk_stage(w, a, hp.B, std::min(p->max_nms, kSelCap2B / 2), kSelCap2B, w.done_a, w.done_b, s)) return e;
    a.skip = w.done_b;   // what is still open after both stages (a score tie wider than the buffers): the full list, as before
  }
  a.done_counter = push ? push->done_counter : nullptr;
  a.signal = push ? push->signal : nullptr;
  if (int e = launch_sort_nms(a, hp.B, s)) return e;
  if (ms2) {
    Y11_CHECK_CUDA(cudaEventRecord(ev[2], s));
    Y11_CHECK_CUDA(cudaEventSynchronize(ev[2]));
    Y11_CHECK_CUDA(cudaEventElapsedTime(&ms2[0], ev[0], ev[1]));
    Y11_CHECK_CUDA(cudaEventElapsedTime(&ms2[1], ev[1], ev[2]));
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return 0;
}

extern "C" int y11_detect_postprocess(y11_handle, const y11_head_desc* hd, const y11_nms_params* p, const float* scale, float* out_det,
                                      int32_t* out_count, int32_t* out_ncand, void* workspace, size_t workspace_bytes, y11_stream s) {
  return postprocess_impl(hd, p, scale, out_det, out_count, out_ncand, workspace, workspace_bytes, nullptr, nullptr,
                          static_cast<cudaStream_t>(s));
}

extern "C" int y11_detect_postprocess_push(y11_handle, const y11_head_desc* hd, const y11_nms_params* p, const float* scale,
                                           float* out_det, int32_t* out_count, int32_t* out_ncand, void* workspace,
                                           size_t workspace_bytes, const y11_push* push, y11_stream s) {
  Y11_REQUIRE(!push || (push->done_counter && push->signal), "postprocess_push: null counter/signal");
  return postprocess_impl(hd, p, scale, out_det, out_count, out_ncand, workspace, workspace_bytes, push, nullptr,
                          static_cast<cudaStream_t>(s));
}

extern "C" int y11_detect_postprocess_list(y11_handle, const y11_head_desc* hd, const y11_nms_params* p, const void* list,
                                           const int32_t* list_count, int32_t list_cap, const float* scale, float* out_det,
                                           int32_t* out_count, int32_t* out_ncand, void* workspace, size_t workspace_bytes,
                                           const y11_push* push, y11_stream s) {
  Y11_REQUIRE(hd && p && out_det && out_count, "y11_detect_postprocess_list: null argument");
  Y11_REQUIRE(!p->multi_label, "y11_detect_postprocess_list: single-label only");
  Y11_REQUIRE(!push || (push->done_counter && push->signal), "y11_detect_postprocess_list: null push field");
  const EmitList el{static_cast<const int4*>(list), list_count, list_cap};
  return postprocess_impl(hd, p, scale, out_det, out_count, out_ncand, workspace, workspace_bytes, push, nullptr,
                          static_cast<cudaStream_t>(s), &el);
}

extern "C" int y11_detect_postprocess_timed(y11_handle, const y11_head_desc* hd, const y11_nms_params* p, const float* scale,
                                            float* out_det, int32_t* out_count, int32_t* out_ncand, void* workspace,
                                            size_t workspace_bytes, float* ms_decode_nms, y11_stream s) {
  Y11_REQUIRE(ms_decode_nms, "postprocess_timed: null ms");
  return postprocess_impl(hd, p, scale, out_det, out_count, out_ncand, workspace, workspace_bytes, nullptr, ms_decode_nms,
                          static_cast<cudaStream_t>(s));
}

extern "C" int y11_detect_postprocess_list_timed(y11_handle, const y11_head_desc* hd, const y11_nms_params* p, const void* list,
                                                 const int32_t* list_count, int32_t list_cap, const float* scale, float* out_det,
                                                 int32_t* out_count, int32_t* out_ncand, void* workspace, size_t workspace_bytes,
                                                 float* ms_decode_nms, y11_stream s) {
  Y11_REQUIRE(hd && p && out_det && out_count && ms_decode_nms, "y11_detect_postprocess_list_timed: null argument");
  Y11_REQUIRE(!p->multi_label, "y11_detect_postprocess_list_timed: single-label only");
  const EmitList el{static_cast<const int4*>(list), list_count, list_cap};
  return postprocess_impl(hd, p, scale, out_det, out_count, out_ncand, workspace, workspace_bytes, nullptr, ms_decode_nms,
                          static_cast<cudaStream_t>(s), &el);
}

extern "C" int y11_wait_signals(y11_handle h, const uint32_t* signals, int n, uint32_t target, y11_stream s) {
  Y11_REQUIRE(h && signals && n >= 1 && n <= 32, "wait_signals: need 1..32 signals");
  wait_signals_kernel<<<1, 32, 0, static_cast<cudaStream_t>(s)>>>(signals, n, target, h->dev_error_flag);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t y11_nms_workspace(int B, int K) {
  const int np2 = next_pow2(K);
  return align_up(np2 > kSmemKeys ? (size_t)B * np2 * 8 : 0) + align_up((size_t)B * K * 16) + align_up((size_t)B * K * 4) + 1024;
}

extern "C" int y11_nms_batched(y11_handle, const float* boxes, const float* scores, const float* cls, const int32_t* n, int B, int K,
                               const y11_nms_params* p, int32_t* keep, int32_t* keep_count, void* workspace, size_t workspace_bytes,
                               y11_stream s_) {
  Y11_REQUIRE(K >= 1 && p->max_det >= 1 && p->max_det <= K, "nms: need 1 <= max_det (%d) <= K (%d)", p->max_det, K);
  Y11_REQUIRE(workspace && workspace_bytes >= y11_nms_workspace(B, K), "nms: workspace too small");
  const int np2 = next_pow2(K);
  char* wp = static_cast<char*>(workspace);
  NmsArgs a;
  a.keys = reinterpret_cast<unsigned long long*>(wp);
  wp += align_up(np2 > kSmemKeys ? (size_t)B * np2 * 8 : 0);
  a.kbox = reinterpret_cast<float4*>(wp);
  wp += align_up((size_t)B * K * 16);
  a.karea = reinterpret_cast<float*>(wp);
  a.cbox = reinterpret_cast<const float4*>(boxes); a.cscore = scores; a.ccls = cls; a.ncand = n;
  a.canchor = nullptr;
  a.cap = K; a.keys_stride = np2;
  a.iou_thr = thr_round_down(p->iou);
  a.class_offset = p->agnostic ? 0.0f : (float)p->max_wh;
  a.max_det = p->max_det; a.max_nms = p->max_nms;
  a.scale = nullptr; a.out_det = nullptr; a.out_keep = keep; a.out_count = keep_count;
  a.done_counter = nullptr; a.signal = nullptr;
  a.skip = nullptr; a.done_out = nullptr; a.total = nullptr;
  return launch_sort_nms(a, B, static_cast<cudaStream_t>(s_));
}
