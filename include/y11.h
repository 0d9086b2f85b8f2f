/*
 * y11.h - C ABI of liby11_b200.so: the B200 (sm_100a) YOLO11 detection hot path.
 *
 * The reference (t0saki/YOLO-Infer) is pure Python and has no FFI of its own: its boundary for this
 * path is the Python class core.model.YOLO11Model (reference core/model.py:29-295) which forwards
 * `predict` to the un-vendored ultralytics engine (core/model.py:118-133).  The entry points below
 * are what a ctypes binding behind that class calls; each cites the reference /
 * upstream-ultralytics step it replaces (SURVEY.md section 8a row numbers in brackets).
 *
 * Conventions: plain C, no torch types.  Every tensor pointer is a DEVICE pointer owned by the
 * caller (PyTorch on the Python side); every call takes the CUDA stream to launch on; return value
 * 0 = ok, negative = error (text via y11_last_error()); nothing throws across the ABI.  Handles are
 * not thread-safe (the Python side holds the same per-model lock the reference's predictor has).
 *
 * Activation layout everywhere: NHWC bf16, a "view" = (base pointer, pixel stride in channels
 * `c_total`, first channel `c_off`, channel count) so that concat / chunk / split of the reference
 * network (ultralytics Concat, C3k2.chunk, C2PSA.split) are pure address arithmetic.
 */
#ifndef Y11_H_
#define Y11_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y11_ABI_VERSION 6

typedef struct y11_engine* y11_handle;
typedef struct y11_plan_s* y11_plan;
typedef void* y11_stream; /* cudaStream_t */

/* ---- lifecycle ------------------------------------------------------------------------------- */
int y11_abi_version(void);
const char* y11_last_error(void);
/* One engine per device.  Replaces the device placement of reference core/model.py:110-112. */
int y11_create(y11_handle* out, int device);
void y11_destroy(y11_handle h);
/* Non-zero after a kernel of this engine gave up on a pipeline wait (bounded mbarrier wait, then __trap): the code names
 * the wait (101-107, see conv_tc.cu).  Readable even when the trapped kernel has left the CUDA context unusable. */
int y11_engine_error_code(y11_handle h);

/* ---- (1) letterbox preprocess  [a4: ultralytics LetterBox + BasePredictor.preprocess] -------- */
typedef struct {
  const uint8_t* src; /* device, BGR uint8 HWC */
  int32_t h0, w0;     /* source size */
  int32_t pitch;      /* source row pitch in bytes */
  int32_t new_h, new_w; /* resized (unpadded) size; == h0,w0 -> pure copy */
  int32_t top, left;  /* placement on the HxW canvas; border value 114 */
} y11_image;
/* images: DEVICE array of B descriptors.  out: bf16 NHWC [B,H,W,3], RGB, value/255.
 * Bilinear arithmetic is bit-identical to cv2.resize(INTER_LINEAR) on uint8 (11-bit fixed point,
 * 2x-downscale area fast path), then exact fp32 /255 and round-to-nearest-even to bf16. */
int y11_letterbox(y11_handle h, const y11_image* images, int B, int H, int W, void* out_bf16_nhwc, y11_stream s);
/* Same resize/pad, uint8 BGR HWC out (parity tests against cv2 bit for bit). */
int y11_letterbox_u8(y11_handle h, const y11_image* images, int B, int H, int W, uint8_t* out_u8_hwc, y11_stream s);
/* Tensor sources [a3: LoadTensor]: fp32 NCHW [B,3,H,W] / divisor -> bf16 NHWC (divisor = 255 when max>1, else 1). */
int y11_nchw_f32_to_nhwc_bf16(y11_handle h, const float* in, int B, int H, int W, float divisor, void* out, y11_stream s);
/* Same with LoadTensor's divisor rule evaluated ON THE DEVICE (no host sync, capturable in a CUDA graph): a max-reduction of
 * the whole tensor into *scratch_max (4 bytes of device memory), then the conversion divides by 255 iff max > 1 + eps.
 * This is the path the reference's own benchmark harness drives (torch.randn batches, benchmarks/speed_benchmark.py:100-102). */
int y11_nchw_f32_to_nhwc_bf16_auto(y11_handle h, const float* in, int B, int H, int W, uint32_t* scratch_max, void* out,
                                   y11_stream s);

/* ---- (2) network plan  [a6-a12: DetectionModel._predict_once over fused Conv/C3k2/SPPF/C2PSA/Detect] */
typedef struct {
  void* ptr;        /* base of the NHWC buffer */
  int32_t c_total;  /* pixel stride in channels */
  int32_t c_off;    /* first channel of the view */
  int32_t c;        /* channels in the view (multiple of 16 for tensor-core convs) */
} y11_view;

enum { Y11_ACT_NONE = 0, Y11_ACT_SILU = 1 };
enum { Y11_IMPL_TCGEN05 = 0, Y11_IMPL_SIMT_DEBUG = 1 };
/* How y11_conv_desc.res enters the conv:
 *   Y11_RES_POST      : out = act(conv + bias) + res            (Bottleneck / PSABlock shortcut; res has the output's size)
 *   Y11_RES_PRE_UP2   : out = act(conv + bias + up2(res))       res is [B, Hout/2, Wout/2] and is read with nearest-neighbour
 *                       2x upsampling.  This is how `Concat([Upsample(2)(p), skip]) -> 1x1 conv` [a11] runs without ever
 *                       materialising the upsampled tensor: a 1x1 conv commutes with nearest upsampling, so the host splits
 *                       the weights by input channel, runs W_up . p (+bias, no activation) at LOW resolution and hands the
 *                       result to the conv over `skip` as this pre-activation term. */
enum { Y11_RES_POST = 0, Y11_RES_PRE_UP2 = 1 };

/* Dense conv k in {1,3}, stride in {1,2}, pad k/2, BN folded  [a7 Conv.forward_fuse]:
 *   out = act(conv(in, w) + bias) (+ res)          (res_mode selects where the residual enters, see above)
 * k = 2 (stride 1, taps at offsets {-1,0} x {-1,0}, i.e. zero padding on the top/left only, Hout = Hin) is the form a
 * 3x3 stride-2 conv takes on a space-to-depth input (see y11_stem_desc.s2d): the host repacks the 3x3 weights.
 * w: bf16 [cout][k*k*cin] with K ordered (kh, kw, cin); bias: fp32 [cout].
 * out may be bf16 (default) or fp32 (out_f32 != 0, used for the Detect logits). */
typedef struct {
  y11_view in, out, res; /* res.ptr == NULL -> no residual; res may alias out (in-place residual) */
  const void* w;
  const float* bias;
  int32_t B, Hin, Win, Hout, Wout;
  int32_t k, stride, act, out_f32;
  int32_t impl; /* Y11_IMPL_TCGEN05 (product) | Y11_IMPL_SIMT_DEBUG (bring-up cross-check only) */
  int32_t res_mode; /* Y11_RES_POST | Y11_RES_PRE_UP2 (ignored when res.ptr == NULL) */
  /* FP8 (e4m3) operands - the Blackwell counterpart of the reference's int8 post-training quantizers
   * (optimization/quantization/quantizers.py:24-310, reached through create_quantizer(...).optimize()):
   *   in_fp8  != 0 : `in` is an e4m3 tensor (1 byte per channel, views counted in channels = bytes) and `w` holds e4m3 weights
   *                  [cout][k*k*cin]; the MMA is tcgen05.mma.kind::f8f6f4 (K = 32 per instruction, fp32 accumulate in TMEM);
   *                  cin % 32 == 0;
   *   cscale       : fp32 [cout] multiplier of the accumulator BEFORE the bias (activation scale x per-channel weight scale);
   *                  NULL = 1;
   *   out_fp8 != 0 : the result (after activation / residual) is multiplied by out_scale (= 1 / activation scale of the
   *                  consumer) and stored as e4m3 with saturation (cvt.rn.satfinite); cout % 32 == 0; excludes out_f32.
   * All three default to 0 / NULL = the bf16 path. */
  int32_t in_fp8, out_fp8;
  const float* cscale;
  float out_scale;
  /* k == 2 only.  s2d_block = c != 0 (c in {32, 64}): `in` is a space-to-depth tensor of 4 blocks of c channels in the PERMUTED
   * block order [(dy,dx)] = [(1,0), (1,1), (0,1), (0,0)] (y11_stem_desc.s2d == 2), and only the channel blocks a tap can touch are
   * loaded - the 3x3 stride-2 conv this stands for reads 1, 2, 2 and 4 of the 4 blocks at its four block taps, and the permuted
   * order makes each of those sets ONE contiguous channel range: tap (-1,-1) -> block 1, (-1,0) -> blocks 0-1, (0,-1) -> blocks
   * 1-2, (0,0) -> blocks 0-3.  `w` is then packed per K stage of 64 channels: bf16 [cout][n_stages * 64], stages ordered
   * (tap = ty*2+tx ascending; within a tap, 64-channel chunks from channel lo*c upward), zero where a channel's (dy,dx) does
   * not belong to the tap.  9/16 of the activation bytes of the plain k = 2 form (10/16 for c = 32). */
  int32_t s2d_block;
} y11_conv_desc;

/* Stem conv: 3 -> cout, 3x3 stride 2, input is the dense 3-channel bf16 NHWC letterbox output. */
typedef struct {
  const void* in; /* bf16 [B,Hin,Win,3] */
  y11_view out;
  const void* w;  /* bf16 [cout][27], K ordered (kh, kw, c) */
  const float* bias;
  int32_t B, Hin, Win, Hout, Wout;
  /* s2d != 0: write the Hout x Wout x cout result in SPACE-TO-DEPTH form, i.e. as a [Hout/2, Wout/2, 4*cout] tensor whose
   * channel block (dy*2 + dx) of pixel (y, x) is output pixel (2y+dy, 2x+dx); `out` then describes that tensor
   * (out.c = 4*cout).  The following 3x3 stride-2 layer becomes a 2x2 stride-1 conv over 4*cout channels, whose
   * 128-byte-or-longer pixel rows the TMA unit can stream (32-byte rows through four parity maps could not). */
  int32_t s2d;   /* 1: block order (dy*2 + dx);  2: permuted block order [(1,0), (1,1), (0,1), (0,0)] (see y11_conv_desc.s2d_block) */
  /* u8_src != 0: read the frames directly (`in` is ignored): images = DEVICE array of B y11_image descriptors whose frames
   * are already at network resolution (h0 == new_h == Hin, w0 == new_w == Win, top == left == 0, src 4-byte aligned, pitch a
   * multiple of 4).  The conversion (BGR->RGB, x*(1/255), bf16) is bit-identical to y11_letterbox's for such frames, so the
   * letterbox launch and its bf16 round trip through HBM are skipped.  See y11_plan_set_stem_source. */
  const y11_image* images;
  int32_t u8_src;
} y11_stem_desc;

/* Depthwise 3x3 stride 1 pad 1  [a7 DWConv, Attention.pe]: out = act(dw(in)+bias) (+ res).
 * w: bf16 [9][c] tap-major. */
typedef struct {
  y11_view in, out, res;
  const void* w;
  const float* bias;
  int32_t B, H, W, act;
} y11_dwconv_desc;

/* SPPF pools  [a9]: io view holds 4*c channels; [0,c) is cv1's output, the op writes
 * maxpool5, maxpool5^2, maxpool5^3 (stride 1, pad 2) into [c,2c), [2c,3c), [3c,4c). */
typedef struct {
  y11_view io;
  int32_t B, H, W, c;
} y11_sppf_desc;

/* nn.Upsample(2,'nearest') into a channel slice  [a11]. in: [B,H,W], out: [B,2H,2W]. */
typedef struct {
  y11_view in, out;
  int32_t B, H, W;
} y11_upsample_desc;

/* PSA attention core  [a10 Attention.forward between qkv and pe/proj]:
 * qkv view channels = [Q: heads*kd | K: heads*kd | V: heads*hd] (the qkv conv's output channels are
 * permuted into this order at weight-packing time); out[b, n, h*hd + d] =
 * sum_m softmax_m(scale * q[b,n,h,:] . k[b,m,h,:]) * v[b,m,h,d]. */
typedef struct {
  y11_view qkv, out;
  int32_t B, N, heads, kd, hd;
  float scale;
} y11_attn_desc;

int y11_plan_create(y11_handle h, y11_plan* out);
void y11_plan_destroy(y11_plan p);
int y11_plan_add_conv(y11_plan p, const y11_conv_desc* d);
/* Same, with an explicit launch variant of the tcgen05 kernel instead of the built-in per-layer heuristic (-1 = heuristic):
 * lsu: activation tiles fetched by cp.async (1) or TMA (0) where both are possible; epi_warp: warp-independent epilogue;
 * ctas_per_sm: persistent CTAs per SM; bn_max: largest N tile.  All variants produce bit-identical results. */
int y11_plan_add_conv_tuned(y11_plan p, const y11_conv_desc* d, int lsu, int epi_warp, int ctas_per_sm, int bn_max);
int y11_plan_add_stem(y11_plan p, const y11_stem_desc* d);
/* Re-point the stem op(s) of the plan: images != NULL -> uint8 frames (u8_src mode; chunk-major plans hand each of their
 * stem ops its slice of the descriptor array), NULL -> the bf16 letterbox output given at y11_plan_add_stem.  Takes effect
 * for launches (and graph captures) made after the call. */
int y11_plan_set_stem_source(y11_plan p, const y11_image* images);
/* Class-emit mode of a Detect class-logit conv (cv3.<l>.2, fp32 out, no activation) [a12 + the conf filter of a13, fused
 * into the conv epilogue].  In single-label prediction the reference reduces the nc class scores of an anchor to
 * (max score, first class attaining it) and keeps the anchor iff max score > conf (ultralytics non_max_suppression).  With
 * emit set, op `op_index` does exactly that on the accumulator rows it already holds in TMEM: it does NOT write its
 * [B, H*W, nc] fp32 logits; every row whose maximum logit exceeds `logit_threshold` (a conservative bound: sigmoid(x) > conf
 * implies x > logit_threshold) is appended to its image's list as {anchor index, class, bits of the max logit, 0}.
 * y11_detect_postprocess_list consumes the lists.  `count` is zeroed whenever op 0 of the plan is launched.  e == NULL or
 * e->list == NULL: back to storing logits.  Takes effect for launches (and graph captures) made after the call. */
typedef struct {
  void* list;             /* device, int32x4 [B][cap] */
  int32_t* count;         /* device, int32 [B]; the same array for every emitting op of a plan */
  int32_t cap;            /* entries per image (>= anchors per image) */
  int32_t nc;             /* classes = valid output columns */
  int32_t anchor_offset;  /* anchor index of this level's pixel (0, 0) */
  float logit_threshold;
} y11_cls_emit;
int y11_plan_set_cls_emit(y11_plan p, int op_index, const y11_cls_emit* e);
int y11_plan_add_dwconv(y11_plan p, const y11_dwconv_desc* d);
int y11_plan_add_sppf(y11_plan p, const y11_sppf_desc* d);
int y11_plan_add_upsample(y11_plan p, const y11_upsample_desc* d);
int y11_plan_add_attention(y11_plan p, const y11_attn_desc* d);
/* Lanes: ops added after y11_plan_set_lane(p, k) run on side lane k (1..7), concurrently with lane 0 (the caller's
 * stream).  y11_plan_fork(p, k): lane k starts after everything added to lane 0 so far; y11_plan_join(p, k): lane 0
 * continues only after everything added to lane k so far.  Used for the six independent Detect towers [a12]. */
int y11_plan_fork(y11_plan p, int lane);
int y11_plan_set_lane(y11_plan p, int lane);
int y11_plan_join(y11_plan p, int lane);
int y11_plan_num_ops(y11_plan p);
/* kernels launched by one y11_plan_run (for bench.py's gpu_launches). */
int y11_plan_num_launches(y11_plan p);
/* Enqueue every op: lane 0 on `s`, side lanes on plan-owned streams forked from / joined into `s` (capturable in a
 * CUDA graph, where the lanes become parallel branches). */
int y11_plan_run(y11_plan p, y11_stream s);
/* Ops [first, last) with their lanes (a fork/join pair must lie inside one range). */
int y11_plan_run_ops(y11_plan p, int first, int last, y11_stream s);
/* Run ops [first, last) only, serially on `s` (lanes ignored; op order is a valid topological order). */
int y11_plan_run_range(y11_plan p, int first, int last, y11_stream s);
/* Run with a CUDA-event pair around every op; ms_per_op has y11_plan_num_ops entries. Synchronises. */
int y11_plan_run_timed(y11_plan p, y11_stream s, float* ms_per_op);
/* Time every tcgen05 conv of the plan in each feasible launch variant on its real buffers (`reps` launches each) and keep
 * the fastest.  One-off, at plan-build time; synchronises `s`; overwrites activation buffers (not weights). */
int y11_plan_autotune(y11_plan p, y11_stream s, int reps);
/* Launch variant of op i: out4 = {lsu, epi_warp, ctas_per_sm, bn}; all -1 for ops that are not tcgen05 convs. */
int y11_plan_op_variant(y11_plan p, int i, int32_t* out4);
/* FLOPs (2*MAC) of op i as launched; 0 for non-conv ops. */
double y11_plan_op_flops(y11_plan p, int i);

/* ---- (3) Detect decode + NMS  [a12 Detect._inference/DFL/dist2bbox, a13 non_max_suppression,
 *          a14 torchvision.ops.nms, a15 scale_boxes+clip_boxes] -------------------------------- */
typedef struct {
  const float* head[3]; /* per level fp32 [B, H_l*W_l, row_stride]: 64 DFL logits then nc class logits */
  int32_t hl[3], wl[3];
  float stride[3];
  int32_t nl, B, nc;
  int32_t row_stride; /* floats per anchor row, >= 64+nc (64 + nc rounded up to 16 when nc % 16 != 0) */
} y11_head_desc;

typedef struct {
  float conf;        /* candidate iff score > (float)conf */
  double iou;        /* suppress iff (double)iou_f32 > iou   (torchvision CPU semantics) */
  int32_t max_det;   /* 300 */
  int32_t max_nms;   /* 30000 */
  int32_t max_wh;    /* 7680: class offset for class-aware NMS */
  int32_t agnostic;
  int32_t multi_label;
} y11_nms_params;

/* Dense decode only: y fp32 [B, 4+nc, A] exactly as Detect._inference returns it (parity tests). */
int y11_decode_dense(y11_handle h, const y11_head_desc* hd, float* y, y11_stream s);

/* Workspace bytes y11_detect_postprocess needs for (B, A, nc, multi_label). */
size_t y11_postprocess_workspace(int B, int A, int nc, int multi_label, int max_nms);

/* Fused path: decode -> conf threshold -> ordered compaction (warp-ballot prefix sums) -> stable
 * score sort -> class-aware IoU NMS (bitmask tiles + warp sweep) -> max_det cut -> scale_boxes+clip.
 * scale: per image [gain, pad_x, pad_y, w0, h0] fp32 (DEVICE, [B,5]); NULL = no rescale/clip.
 * out_det: fp32 [B, max_det, 6] = x1,y1,x2,y2,conf,cls; out_count: int32 [B];
 * out_ncand (optional, may be NULL): int32 [B] candidates above conf before NMS. */
int y11_detect_postprocess(y11_handle h, const y11_head_desc* hd, const y11_nms_params* p, const float* scale,
                           float* out_det, int32_t* out_count, int32_t* out_ncand, void* workspace,
                           size_t workspace_bytes, y11_stream s);

/* Same, for the multi-GPU result gather WITHOUT a collective ("result push"): out_det / out_count may point into
 * ANOTHER GPU's memory (an NVLink peer mapping: cudaDeviceEnablePeerAccess in one process, CUDA IPC / symmetric memory
 * across processes) - the NMS kernel's publish loop then writes the kept boxes straight into the gathering rank's buffer.
 * push->signal (system-scope, normally in the consumer's memory) is incremented ONCE per call, after every result write of
 * the call is visible system-wide; push->done_counter is a device-local int, zero when the call starts, used to find the
 * last CTA.  push == NULL: identical to y11_detect_postprocess. */
typedef struct {
  int32_t* done_counter;
  uint32_t* signal;
} y11_push;
int y11_detect_postprocess_push(y11_handle h, const y11_head_desc* hd, const y11_nms_params* p, const float* scale,
                                float* out_det, int32_t* out_count, int32_t* out_ncand, void* workspace,
                                size_t workspace_bytes, const y11_push* push, y11_stream s);

/* Same post-processing, fed by the pre-candidate lists of class-emit convs (y11_plan_set_cls_emit) instead of a scan of the
 * class logits: per list entry the accurate score, the conf test, DFL decode of the anchor's box logits (hd->head[l][..0:64],
 * which the box towers still write) - then the identical sort + NMS.  Single-label only (p->multi_label must be 0).  Results
 * are bit-identical to y11_detect_postprocess on the logits the convs would have stored.  push may be NULL. */
int y11_detect_postprocess_list(y11_handle h, const y11_head_desc* hd, const y11_nms_params* p, const void* list,
                                const int32_t* list_count, int32_t list_cap, const float* scale, float* out_det,
                                int32_t* out_count, int32_t* out_ncand, void* workspace, size_t workspace_bytes,
                                const y11_push* push, y11_stream s);

/* Consumer side: enqueue a one-warp kernel that parks until signals[i] >= target (wrap-safe) for all i < n (n <= 32). */
int y11_wait_signals(y11_handle h, const uint32_t* signals, int n, uint32_t target, y11_stream s);
/* Measurement only: same launches with CUDA events between them; ms_decode_nms[0] = decode + compaction, [1] = sort + NMS.
 * Synchronises the stream. */
int y11_detect_postprocess_timed(y11_handle h, const y11_head_desc* hd, const y11_nms_params* p, const float* scale,
                                 float* out_det, int32_t* out_count, int32_t* out_ncand, void* workspace,
                                 size_t workspace_bytes, float* ms_decode_nms, y11_stream s);

/* Measurement only: y11_detect_postprocess_list with CUDA events between its launches ([0] = list decode, [1] = sort + NMS). */
int y11_detect_postprocess_list_timed(y11_handle h, const y11_head_desc* hd, const y11_nms_params* p, const void* list,
                                      const int32_t* list_count, int32_t list_cap, const float* scale, float* out_det,
                                      int32_t* out_count, int32_t* out_ncand, void* workspace, size_t workspace_bytes,
                                      float* ms_decode_nms, y11_stream s);

/* NMS only, batched, on caller-provided candidates (the bit-exact test against torchvision.ops.nms):
 * boxes fp32 [B, K, 4] xyxy (class offset NOT yet applied), scores fp32 [B,K], cls fp32 [B,K],
 * n int32 [B] valid candidates per image.  keep: int32 [B, max_det] candidate indices in score order. */
int y11_nms_batched(y11_handle h, const float* boxes, const float* scores, const float* cls, const int32_t* n,
                    int B, int K, const y11_nms_params* p, int32_t* keep, int32_t* keep_count, void* workspace,
                    size_t workspace_bytes, y11_stream s);
size_t y11_nms_workspace(int B, int K);

/* ---- (4) result rasteriser  [a17 consumers / section 8f row 3: utils/visualization.py:18-106 draw_detections] ----------------
 * Draws every image's detections (box outline, filled label background, label text `{name}: {conf:.2f}`, palette of
 * visualization.py:get_color, painter's order) straight into uint8 BGR frames in device memory - bit-identical to the
 * reference's cv2.rectangle / cv2.getTextSize / cv2.putText loop for line_thickness 1 or 2, FONT_HERSHEY_SIMPLEX at scale 0.5,
 * thickness 1 (the defaults every call site uses); glyphs cut by the image border may differ in a few pixels. */
typedef struct {
  uint8_t* img;          /* device, BGR uint8 HWC, modified in place */
  int32_t h, w, pitch;   /* pitch in bytes */
  const float* det;      /* device, [max_det][6] = x1,y1,x2,y2,conf,cls rows of THIS image (y11_detect_postprocess layout) */
  const int32_t* count;  /* device pointer to the number of valid rows, or NULL -> n */
  int32_t n, max_det;
} y11_draw_item;
typedef struct {
  const uint32_t* glyph_bits; /* device, [n_chars][2 pen phases][cell_h] row masks, bit x = column x of the cell */
  const int32_t* advance;     /* device, [n_chars] pen advance in HALF pixels */
  int32_t first_char, n_chars, cell_h, cell_w, base_y, pad_x;
  int32_t text_h;             /* cv2.getTextSize(...)[0][1] of the font (12) */
  const char* names;          /* device, [nc][name_stride] zero-padded class names */
  int32_t nc, name_stride;
} y11_font;
/* items: DEVICE array of n_items descriptors; max_h / max_w: largest image extent among them (grid size). */
int y11_draw_detections(y11_handle h, const y11_draw_item* items, int n_items, int max_h, int max_w, const y11_font* font,
                        int line_thickness, y11_stream s);

/* ---- (5) GPU JPEG decode  [section 8f row 2: cv2.imread of utils/data_loader.py:42 / ultralytics LoadImagesAndVideos] -------------
 * Compressed bytes in HOST memory -> BGR uint8 HWC frame in DEVICE memory (what y11_image.src points at), through nvJPEG
 * (resolved with dlopen at first use; absent library -> these calls fail with a message, the rest of the ABI is unaffected).
 * Not bit-identical to cv2.imread (different IDCT / chroma upsampling): tolerance stated in tests/test_gpu_decode.py.
 * One decoder object per thread / stream of decodes. */
typedef struct y11_jpeg_s* y11_jpeg;
int y11_jpeg_create(y11_handle h, y11_jpeg* out);
void y11_jpeg_destroy(y11_jpeg j);
int y11_jpeg_info(y11_jpeg j, const uint8_t* data, size_t nbytes, int32_t* h, int32_t* w, int32_t* components);
/* out_bgr: device [h][pitch] bytes; h, w must equal y11_jpeg_info's.  Enqueued on `s` (the Huffman stage runs on the host). */
int y11_jpeg_decode(y11_jpeg j, const uint8_t* data, size_t nbytes, uint8_t* out_bgr, int32_t pitch, int32_t h, int32_t w,
                    y11_stream s);

#ifdef __cplusplus
}
#endif
#endif /* Y11_H_ */
