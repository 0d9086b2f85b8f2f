// C ABI plumbing of liby11_b200: engine lifecycle, error string, and the op-list "plan" executor that replays
// the fused YOLO11 network (built once per (model, B, H, W) by the Python host from the reference topology,
// see yolo_infer_b200/network.py) as a fixed sequence of kernel launches on the caller's stream.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ops.h"

static thread_local char g_err[1024] = "";

void y11_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* y11_last_error(void) { return g_err; }
extern "C" int y11_abi_version(void) { return Y11_ABI_VERSION; }

extern "C" int y11_create(y11_handle* out, int device) {
  Y11_REQUIRE(out, "y11_create: null out");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  Y11_REQUIRE(e == cudaSuccess && count > 0, "y11_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  Y11_REQUIRE(device >= 0 && device < count, "y11_create: device %d out of range (%d devices)", device, count);
  Y11_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  Y11_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  Y11_REQUIRE(prop.major == 10, "y11_create: device %d is sm_%d%d; liby11_b200 contains sm_100a code only", device, prop.major, prop.minor);
  y11_engine* eng = new y11_engine();
  eng->device = device;
  eng->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete eng;
    y11_set_error("y11_create: cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return -2;
  }
  eng->encode_tiled = reinterpret_cast<y11_encode_tiled_fn>(fn);
  // pipeline-timeout code of the kernels' bounded mbarrier waits: mapped pinned memory, so that the host can still read it
  // after a trapped kernel has poisoned the context (y11_engine_error_code)
  e = cudaHostAlloc(reinterpret_cast<void**>(&eng->host_error_flag), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *eng->host_error_flag = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&eng->dev_error_flag), eng->host_error_flag, 0);
  }
  if (e != cudaSuccess) {
    if (eng->host_error_flag) cudaFreeHost(eng->host_error_flag);
    delete eng;
    y11_set_error("y11_create: error-flag allocation failed (%s)", cudaGetErrorString(e));
    return -2;
  }
  *out = eng;
  return 0;
}

extern "C" int y11_engine_error_code(y11_handle h) { return (h && h->host_error_flag) ? *h->host_error_flag : 0; }

extern "C" void y11_destroy(y11_handle h) {
  if (!h) return;
  if (h->host_error_flag) cudaFreeHost(h->host_error_flag);
  delete h;
}

// ------------------------------------------------------------------------------------------------ plan
enum OpKind { OP_CONV_TC, OP_CONV_SIMT, OP_STEM, OP_DWCONV, OP_SPPF, OP_UPSAMPLE, OP_ATTN };

struct PlanOp {
  OpKind kind;
  ConvTcLaunch tc;  // OP_CONV_TC
  DwTmaLaunch dwt;  // OP_DWCONV on large maps
  bool use_dwt;
  union {
    y11_conv_desc conv;
    y11_stem_desc stem;
    y11_dwconv_desc dw;
    y11_sppf_desc sppf;
    y11_upsample_desc up;
    y11_attn_desc attn;
  } d;
  double flops;
};

// Schedule entry: a kernel launch on a lane, or a fork/join edge between lane 0 (the caller's stream) and a side lane.
// Lanes let independent branches of the network (the six Detect towers) run concurrently: the towers of the 80x80 level
// start as soon as the P3 feature map exists and overlap the small, GPU-underfilling layers of the rest of the neck.
enum SchedKind { S_OP, S_FORK, S_JOIN };
struct SchedItem { SchedKind kind; int op; int lane; int pos; };  // pos = number of ops added before this item
constexpr int kMaxLanes = 8;

struct y11_plan_s {
  y11_engine* eng;
  std::vector<PlanOp*> ops;
  std::vector<SchedItem> sched;
  int cur_lane = 0;
  cudaStream_t lanes[kMaxLanes] = {};     // [0] unused (caller's stream); side lanes are created lazily
  cudaEvent_t fork_ev[kMaxLanes] = {}, join_ev[kMaxLanes] = {};
  cudaEvent_t* events = nullptr;
  int n_events = 0;
  int* counters = nullptr;  // 2 ints per op: dynamic-tile-scheduler state of the tcgen05 convs (conv_tc.cu)
  int* emit_count = nullptr;  // class-emit mode (y11_plan_set_cls_emit): per-image list counters, zeroed when op 0 is launched
  int emit_B = 0;
};
constexpr int kMaxPlanOps = 1024;

// Dynamic tile scheduler of conv_tc_kernel (global atomic tile counter instead of the static `tile += gridDim.x` walk).
// Off by default: measured on B200 it is neutral (YOLO11s 20.0 k vs 19.9 k img/s) to slightly negative (YOLO11n 32.9 k vs
// 33.6 k) - CTAs of a persistent grid become resident together, so there is no late-CTA tail for it to remove.
static bool dyn_tiles_enabled() {
  const char* e = getenv("Y11_DYN_TILES");
  return e ? atoi(e) != 0 : false;
}

extern "C" int y11_plan_create(y11_handle h, y11_plan* out) {
  Y11_REQUIRE(h && out, "y11_plan_create: null argument");
  y11_plan_s* p = new y11_plan_s();
  p->eng = h;
  if (dyn_tiles_enabled()) {
    cudaError_t e = cudaMalloc(&p->counters, 2 * kMaxPlanOps * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(p->counters, 0, 2 * kMaxPlanOps * sizeof(int));
    if (e != cudaSuccess) {
      if (p->counters) cudaFree(p->counters);
      delete p;
      y11_set_error("y11_plan_create: tile-counter allocation failed (%s)", cudaGetErrorString(e));
      return -2;
    }
  }
  *out = p;
  return 0;
}

extern "C" void y11_plan_destroy(y11_plan p) {
  if (!p) return;
  for (PlanOp* op : p->ops) delete op;
  for (int l = 1; l < kMaxLanes; ++l) {
    if (p->lanes[l]) cudaStreamDestroy(p->lanes[l]);
    if (p->fork_ev[l]) cudaEventDestroy(p->fork_ev[l]);
    if (p->join_ev[l]) cudaEventDestroy(p->join_ev[l]);
  }
  for (int i = 0; i < p->n_events; ++i) cudaEventDestroy(p->events[i]);
  delete[] p->events;
  if (p->counters) cudaFree(p->counters);
  delete p;
}

static PlanOp* new_op(OpKind k) {
  PlanOp* op = new PlanOp();
  std::memset(static_cast<void*>(op), 0, sizeof(PlanOp));
  op->kind = k;
  return op;
}

extern "C" int y11_plan_add_conv(y11_plan p, const y11_conv_desc* d) {
  Y11_REQUIRE(p && d, "plan_add_conv: null argument");
  Y11_REQUIRE(d->in.ptr && d->out.ptr && d->w && d->bias, "plan_add_conv: null tensor");
  PlanOp* op = new_op(d->impl == Y11_IMPL_SIMT_DEBUG ? OP_CONV_SIMT : OP_CONV_TC);
  op->d.conv = *d;
  op->flops = 2.0 * d->B * d->Hout * d->Wout * (double)d->out.c * d->in.c * d->k * d->k;
  if (op->kind == OP_CONV_TC) {
    if (int e = conv_tc_prepare(p->eng, d, &op->tc)) { delete op; return e; }
    if (p->counters && p->ops.size() < (size_t)kMaxPlanOps) op->tc.p.tile_counter = p->counters + 2 * p->ops.size();
  }
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_add_conv_tuned(y11_plan p, const y11_conv_desc* d, int lsu, int epi_warp, int ctas_per_sm, int bn_max) {
  Y11_REQUIRE(p && d, "plan_add_conv_tuned: null argument");
  Y11_REQUIRE(d->in.ptr && d->out.ptr && d->w && d->bias, "plan_add_conv_tuned: null tensor");
  Y11_REQUIRE(d->impl == Y11_IMPL_TCGEN05, "plan_add_conv_tuned: tcgen05 convs only");
  PlanOp* op = new_op(OP_CONV_TC);
  op->d.conv = *d;
  op->flops = 2.0 * d->B * d->Hout * d->Wout * (double)d->out.c * d->in.c * d->k * d->k;
  const ConvTcTune t{lsu, epi_warp, ctas_per_sm, bn_max};
  if (int e = conv_tc_prepare(p->eng, d, &op->tc, &t)) { delete op; return e; }
  if (p->counters && p->ops.size() < (size_t)kMaxPlanOps) op->tc.p.tile_counter = p->counters + 2 * p->ops.size();
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_op_variant(y11_plan p, int i, int32_t* out4) {
  Y11_REQUIRE(p && out4 && i >= 0 && i < (int)p->ops.size(), "plan_op_variant: bad argument");
  const PlanOp* op = p->ops[i];
  if (op->kind != OP_CONV_TC) { out4[0] = out4[1] = out4[2] = out4[3] = -1; return 0; }
  out4[0] = op->tc.variant.lsu; out4[1] = op->tc.variant.epi_warp; out4[2] = op->tc.variant.cps; out4[3] = op->tc.variant.bn_max;
  return 0;
}

// Plan autotuner: every tcgen05 conv is timed on its real buffers in each feasible launch variant (producer: TMA | cp.async,
// epilogue: CTA-wide | warp-independent, 2 | 3 persistent CTAs per SM, narrower N tiles for layers with few M tiles) and the
// fastest is kept.  All variants compute bit-identical results (same per-element K order, same epilogue arithmetic), so
// the choice affects time only.  Runs `reps` back-to-back launches per variant after one warm-up; synchronises.
extern "C" int y11_plan_autotune(y11_plan p, y11_stream s_, int reps) {
  Y11_REQUIRE(p, "plan_autotune: null plan");
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  reps = std::max(1, std::min(reps, 50));
  cudaEvent_t e0, e1;
  Y11_CHECK_CUDA(cudaEventCreate(&e0));
  Y11_CHECK_CUDA(cudaEventCreate(&e1));
  int rc = 0;
  for (PlanOp* op : p->ops) {
    if (op->kind != OP_CONV_TC) continue;
    const y11_conv_desc* d = &op->d.conv;
    const ConvTcLaunch base = op->tc;
    std::vector<ConvTcTune> cands;
    const int bn0 = base.variant.bn_max;
    const long long tiles0 = (long long)base.p.tiles_w * base.p.tiles_h * base.p.tiles_n * base.p.n_tiles;
    for (int lsu = base.halo_tma_eligible ? 2 : base.lsu_eligible ? 1 : 0; lsu >= 0; --lsu)
      for (int ew = 0; ew <= 3; ++ew)  // bit 0: per-warp epilogue, bit 1: fat epilogue
        for (int cps = 3; cps >= 2; --cps) {
          if (lsu == 1 && !base.lsu_eligible) continue;
          cands.push_back(ConvTcTune{lsu, ew, cps, -1});
          // few tiles (less than two waves of persistent CTAs): narrower N tiles spread the layer over more SMs
          if (lsu == 0 && bn0 >= 64 && tiles0 < 2ll * p->eng->num_sms * 3) {
            cands.push_back(ConvTcTune{lsu, ew, cps, bn0 / 2});
            if (bn0 == 256) cands.push_back(ConvTcTune{lsu, ew, cps, 64});
          } else if (bn0 == 256) {
            cands.push_back(ConvTcTune{lsu, ew, cps, 128});  // 128x256 tiles need all of TMEM (1 CTA/SM): not always the best trade
          }
          // ... and the other way round: 128x256 tiles for shorter-K layers whose heuristic tile is 128 wide (half the
          // activation re-reads from L2; the layers on 20x20 / 40x40 maps run at the L2 -> SM bandwidth cap)
          if (bn0 == 128 && d->out.c % 256 == 0 && d->in.c % 64 == 0 && cps == 2 && lsu == 0)
            cands.push_back(ConvTcTune{lsu, ew, 1, 256});
        }
    // halo-stream mode (lsu = 3): swizzled TMA halo + streamed weights for 3x3 stride-1 layers with 64 / 128 input channels
    // (measured: 128 -> 64 on 80x80 73.7 us against 79.9 us for the CTA pair and 92 us tap by tap - with ONE halo tile buffer and
    //  2 CTAs per SM; at 1 CTA per SM, which the fat epilogue's staging forces, it loses: 104 us; equal to the others on 40x40 maps)
    if (base.hstream_eligible)
      for (int ew = 0; ew <= 1; ++ew) cands.push_back(ConvTcTune{3, ew, 2, -1});
    // resident weights (epi_warp bit 2) for single-N-tile TMA layers that walk several tiles per CTA
    if (!base.p.halo && d->out.c == bn0 && tiles0 >= 4ll * p->eng->num_sms) {
      for (int ew = 4; ew <= 7; ++ew)
        for (int cps = 3; cps >= 2; --cps) cands.push_back(ConvTcTune{0, ew, cps, -1});
    }
    // CTA-pair variant (tcgen05.mma.cta_group::2, conv_tc_kernel_pair): half the weight-tile traffic per SM, 256-row tiles;
    // prepare declines it where it does not apply (halo / e4m3 / k = 2 layers) and the duplicate filter below drops those
    {
      static const bool pair_on = [] { const char* e = getenv("Y11_PAIR"); return e ? atoi(e) != 0 : true; }();
      if (pair_on && !(base.p.halo) && bn0 >= 64) {
        for (int cps = 2; cps >= 1; --cps)
          for (int ew = 8; ew <= 10; ew += 2) {  // plain and fat epilogue
            cands.push_back(ConvTcTune{0, ew, cps, -1});
            if (d->out.c % 256 == 0 && bn0 != 256) cands.push_back(ConvTcTune{0, ew, cps, 256});
            if (bn0 == 256) cands.push_back(ConvTcTune{0, ew, cps, 128});
          }
      }
    }
    // time one variant: best of three trials of `reps` back-to-back launches (after one warm-up launch)
    auto time_variant = [&](const ConvTcLaunch& L, float* out_ms) -> int {
      if (int e = conv_tc_launch(&L, s)) return e;
      float best_trial = 1e30f;
      for (int trial = 0; trial < 3; ++trial) {
        Y11_CHECK_CUDA(cudaEventRecord(e0, s));
        for (int r = 0; r < reps; ++r)
          if (int e = conv_tc_launch(&L, s)) return e;
        Y11_CHECK_CUDA(cudaEventRecord(e1, s));
        Y11_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        Y11_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best_trial = std::min(best_trial, ms);
      }
      *out_ms = best_trial;
      return 0;
    };
    float base_ms = 0.f;
    if ((rc = time_variant(base, &base_ms))) break;
    // a variant must beat the heuristic's choice by 4 % to replace it (timing noise must not flip layers back and forth)
    float best = base_ms * 0.96f;
    ConvTcLaunch best_l = base;
    std::vector<ConvTcTune> seen{base.variant};
    for (const ConvTcTune& c : cands) {
      ConvTcLaunch L;
      if (conv_tc_prepare(p->eng, d, &L, &c)) continue;  // variant not feasible for this layer
      L.p.tile_counter = base.p.tile_counter;
      L.p.emit = base.p.emit; L.p.emit_nc = base.p.emit_nc; L.p.emit_aoff = base.p.emit_aoff; L.p.emit_cap = base.p.emit_cap;
      L.p.emit_thr = base.p.emit_thr; L.p.emit_list = base.p.emit_list; L.p.emit_count = base.p.emit_count;
      bool dup = false;
      for (const ConvTcTune& v : seen)
        dup |= v.lsu == L.variant.lsu && v.epi_warp == L.variant.epi_warp && v.cps == L.variant.cps && v.bn_max == L.variant.bn_max;
      if (dup) continue;
      seen.push_back(L.variant);
      float ms = 0.f;
      if ((rc = time_variant(L, &ms))) break;
      if (ms < best) { best = ms; best_l = L; }
    }
    if (rc) break;
    op->tc = best_l;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  y11_set_error("");  // infeasible candidates leave their message behind
  return rc;
}

extern "C" int y11_plan_add_stem(y11_plan p, const y11_stem_desc* d) {
  Y11_REQUIRE(p && d && d->in && d->out.ptr && d->w && d->bias, "plan_add_stem: null argument");
  Y11_REQUIRE(d->out.c_off % 8 == 0 && d->out.c_total % 8 == 0, "plan_add_stem: output alignment");
  PlanOp* op = new_op(OP_STEM);
  op->d.stem = *d;
  op->flops = 2.0 * d->B * d->Hout * d->Wout * (double)(d->s2d ? d->out.c / 4 : d->out.c) * 27;
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_set_stem_source(y11_plan p, const y11_image* images) {
  Y11_REQUIRE(p, "plan_set_stem_source: null plan");
  size_t first = 0;
  for (PlanOp* op : p->ops) {
    if (op->kind != OP_STEM) continue;
    op->d.stem.images = images ? images + first : nullptr;
    op->d.stem.u8_src = images ? 1 : 0;
    first += (size_t)op->d.stem.B;
  }
  return 0;
}

// Class-emit mode of a Detect class-logit conv (cv3.l.2): see y11.h.  Like y11_plan_set_stem_source this edits the launch
// parameters of an existing op; a CUDA-graph capture records whatever is set at capture time.
extern "C" int y11_plan_set_cls_emit(y11_plan p, int op_index, const y11_cls_emit* e) {
  Y11_REQUIRE(p && op_index >= 0 && op_index < (int)p->ops.size(), "plan_set_cls_emit: bad op index %d", op_index);
  PlanOp* op = p->ops[op_index];
  Y11_REQUIRE(op->kind == OP_CONV_TC, "plan_set_cls_emit: op %d is not a tcgen05 conv", op_index);
  ConvTcParams& q = op->tc.p;   // (op->tc is re-assigned below: q stays a reference to the same member)
  if (!e || !e->list) {
    q.emit = 0; q.emit_list = nullptr; q.emit_count = nullptr;
    bool any = false;
    for (PlanOp* o : p->ops) any |= o->kind == OP_CONV_TC && o->tc.p.emit;
    if (!any) { p->emit_count = nullptr; p->emit_B = 0; }
    return 0;
  }
  Y11_REQUIRE(e->count && e->cap > 0 && e->nc > 0 && e->nc <= op->d.conv.out.c, "plan_set_cls_emit: bad list / class count");
  Y11_REQUIRE(q.out_f32 && q.act == Y11_ACT_NONE && !q.res && !q.quant && op->d.conv.out.c <= 128,
              "plan_set_cls_emit: op %d is not an fp32 logit conv of <= 128 channels", op_index);
  if (!q.emit && (q.n_tiles != 1 || !q.epi_warp)) {
    // an epilogue thread must see all columns of its row (one N tile), and the warp-independent epilogue keeps all eight
    // epilogue warps busy in emit mode (the CTA-wide one scans with four): re-prepare the op in that variant where the tile allows
    ConvTcTune t = op->tc.variant;
    t.bn_max = -1;
    t.epi_warp = op->tc.epi_warp_possible ? 1 : 0;
    ConvTcLaunch L;
    if (int rc = conv_tc_prepare(p->eng, &op->d.conv, &L, &t)) return rc;
    L.p.tile_counter = q.tile_counter;
    Y11_REQUIRE(L.p.n_tiles == 1, "plan_set_cls_emit: op %d does not fit one N tile", op_index);
    op->tc = L;
  }
  Y11_REQUIRE(!p->emit_count || p->emit_count == e->count, "plan_set_cls_emit: one counter array per plan");
  q.emit = 1; q.emit_nc = e->nc; q.emit_aoff = e->anchor_offset; q.emit_cap = e->cap; q.emit_thr = e->logit_threshold;
  q.emit_list = static_cast<int4*>(e->list); q.emit_count = e->count;
  p->emit_count = e->count;
  p->emit_B = q.B;
  return 0;
}

extern "C" int y11_plan_add_dwconv(y11_plan p, const y11_dwconv_desc* d) {
  Y11_REQUIRE(p && d && d->in.ptr && d->out.ptr && d->w && d->bias, "plan_add_dwconv: null argument");
  PlanOp* op = new_op(OP_DWCONV);
  op->d.dw = *d;
  op->flops = 2.0 * d->B * d->H * d->W * (double)d->in.c * 9;
  op->use_dwt = dwconv_tma_eligible(d);
  if (op->use_dwt) {
    if (int e = dwconv_tma_prepare(p->eng, d, &op->dwt)) { delete op; return e; }
  }
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_add_sppf(y11_plan p, const y11_sppf_desc* d) {
  Y11_REQUIRE(p && d && d->io.ptr, "plan_add_sppf: null argument");
  PlanOp* op = new_op(OP_SPPF);
  op->d.sppf = *d;
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_add_upsample(y11_plan p, const y11_upsample_desc* d) {
  Y11_REQUIRE(p && d && d->in.ptr && d->out.ptr, "plan_add_upsample: null argument");
  PlanOp* op = new_op(OP_UPSAMPLE);
  op->d.up = *d;
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_add_attention(y11_plan p, const y11_attn_desc* d) {
  Y11_REQUIRE(p && d && d->qkv.ptr && d->out.ptr, "plan_add_attention: null argument");
  PlanOp* op = new_op(OP_ATTN);
  op->d.attn = *d;
  op->flops = 2.0 * d->B * d->heads * (double)d->N * d->N * (d->kd + d->hd);
  p->ops.push_back(op);
  p->sched.push_back({S_OP, (int)p->ops.size() - 1, p->cur_lane, (int)p->ops.size() - 1});
  return 0;
}

extern "C" int y11_plan_num_ops(y11_plan p) { return p ? (int)p->ops.size() : 0; }
extern "C" int y11_plan_num_launches(y11_plan p) { return p ? (int)p->ops.size() : 0; }  // one kernel per op
extern "C" double y11_plan_op_flops(y11_plan p, int i) { return (p && i >= 0 && i < (int)p->ops.size()) ? p->ops[i]->flops : 0.0; }

static int run_op(const PlanOp* op, cudaStream_t s) {
  switch (op->kind) {
    case OP_CONV_TC: return conv_tc_launch(&op->tc, s);
    case OP_CONV_SIMT: return conv_simt_launch(&op->d.conv, s);
    case OP_STEM: return stem_launch(&op->d.stem, s);
    case OP_DWCONV: return op->use_dwt ? dwconv_tma_launch(&op->dwt, s) : dwconv_launch(&op->d.dw, s);
    case OP_SPPF: return sppf_launch(&op->d.sppf, s);
    case OP_UPSAMPLE: return upsample_launch(&op->d.up, s);
    case OP_ATTN: return attention_launch(&op->d.attn, s);
  }
  return -1;
}

extern "C" int y11_plan_fork(y11_plan p, int lane) {
  Y11_REQUIRE(p && lane >= 1 && lane < kMaxLanes, "plan_fork: lane %d out of range [1, %d)", lane, kMaxLanes);
  Y11_REQUIRE(p->cur_lane == 0, "plan_fork: forks are issued from lane 0");
  if (!p->lanes[lane]) {
    Y11_CHECK_CUDA(cudaStreamCreateWithFlags(&p->lanes[lane], cudaStreamNonBlocking));
    Y11_CHECK_CUDA(cudaEventCreateWithFlags(&p->fork_ev[lane], cudaEventDisableTiming));
    Y11_CHECK_CUDA(cudaEventCreateWithFlags(&p->join_ev[lane], cudaEventDisableTiming));
  }
  p->sched.push_back({S_FORK, -1, lane, (int)p->ops.size()});
  return 0;
}

extern "C" int y11_plan_set_lane(y11_plan p, int lane) {
  Y11_REQUIRE(p && lane >= 0 && lane < kMaxLanes && (lane == 0 || p->lanes[lane]), "plan_set_lane: lane %d not forked", lane);
  p->cur_lane = lane;
  return 0;
}

extern "C" int y11_plan_join(y11_plan p, int lane) {
  Y11_REQUIRE(p && lane >= 1 && lane < kMaxLanes && p->lanes[lane], "plan_join: lane %d not forked", lane);
  p->sched.push_back({S_JOIN, -1, lane, (int)p->ops.size()});
  return 0;
}

// Serial execution of ops [first, last) on one stream (tests, per-op timing, weight conditioning): lanes are ignored,
// the op order is a valid topological order.
extern "C" int y11_plan_run_range(y11_plan p, int first, int last, y11_stream s) {
  Y11_REQUIRE(p && first >= 0 && last <= (int)p->ops.size() && first <= last, "plan_run_range: bad range");
  for (int i = first; i < last; ++i)
    if (int e = run_op(p->ops[i], static_cast<cudaStream_t>(s))) return e;
  return 0;
}

// Ops [first, last) with their lanes: side lanes fork from / join into the caller's stream through events, which is also
// how a stream capture turns them into parallel branches of the CUDA graph.  A fork belongs to the range of the op that
// follows it, a join to the range of the op that precedes it.
extern "C" int y11_plan_run_ops(y11_plan p, int first, int last, y11_stream s_) {
  Y11_REQUIRE(p && first >= 0 && last <= (int)p->ops.size() && first <= last, "plan_run_ops: bad range");
  cudaStream_t s0 = static_cast<cudaStream_t>(s_);
  if (first == 0 && p->emit_count) Y11_CHECK_CUDA(cudaMemsetAsync(p->emit_count, 0, (size_t)p->emit_B * sizeof(int), s0));
  for (const SchedItem& it : p->sched) {
    switch (it.kind) {
      case S_OP:
        if (it.pos < first || it.pos >= last) break;
        if (int e = run_op(p->ops[it.op], it.lane == 0 ? s0 : p->lanes[it.lane])) return e;
        break;
      case S_FORK:
        if (it.pos < first || it.pos >= last) break;
        Y11_CHECK_CUDA(cudaEventRecord(p->fork_ev[it.lane], s0));
        Y11_CHECK_CUDA(cudaStreamWaitEvent(p->lanes[it.lane], p->fork_ev[it.lane], 0));
        break;
      case S_JOIN:
        if (it.pos <= first || it.pos > last) break;
        Y11_CHECK_CUDA(cudaEventRecord(p->join_ev[it.lane], p->lanes[it.lane]));
        Y11_CHECK_CUDA(cudaStreamWaitEvent(s0, p->join_ev[it.lane], 0));
        break;
    }
  }
  return 0;
}

extern "C" int y11_plan_run(y11_plan p, y11_stream s) {
  Y11_REQUIRE(p, "plan_run: null plan");
  return y11_plan_run_ops(p, 0, (int)p->ops.size(), s);
}

extern "C" int y11_plan_run_timed(y11_plan p, y11_stream s_, float* ms_per_op) {
  Y11_REQUIRE(p && ms_per_op, "plan_run_timed: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  const int n = (int)p->ops.size();
  if (p->n_events < n + 1) {
    for (int i = 0; i < p->n_events; ++i) cudaEventDestroy(p->events[i]);
    delete[] p->events;
    p->events = new cudaEvent_t[n + 1];
    p->n_events = n + 1;
    for (int i = 0; i <= n; ++i) Y11_CHECK_CUDA(cudaEventCreate(&p->events[i]));
  }
  Y11_CHECK_CUDA(cudaEventRecord(p->events[0], s));
  for (int i = 0; i < n; ++i) {
    if (int e = run_op(p->ops[i], s)) return e;
    Y11_CHECK_CUDA(cudaEventRecord(p->events[i + 1], s));
  }
  Y11_CHECK_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) Y11_CHECK_CUDA(cudaEventElapsedTime(&ms_per_op[i], p->events[i], p->events[i + 1]));
  return 0;
}
