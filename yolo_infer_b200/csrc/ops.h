// Internal (non-ABI) declarations shared by the .cu translation units of liby11_b200.
#pragma once
#include <utility>

#include "common.cuh"

// Launch with programmatic stream serialization (PDL): see pdl_wait()/pdl_trigger() in common.cuh.
template <typename... KArgs, typename... Args>
static inline cudaError_t y11_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Same, as clusters of two CTAs (the cta_group::2 conv variant): gridDim.x must be even.
template <typename... KArgs, typename... Args>
static inline cudaError_t y11_launch_pdl_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 2;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- tcgen05 implicit-GEMM conv (conv_tc.cu) ---------------------------------------------------
struct ConvTcMaps {
  CUtensorMap a[4];  // activation views: [0] for stride 1; [ph*2+pw] parity sub-grids for stride 2
  CUtensorMap b;     // packed weights [cout][k*k*cin]
  CUtensorMap out;   // output view, cw-channel x whole-tile boxes (TMA store, CTA-wide epilogue)
  CUtensorMap outq;  // output view, cw-channel x quarter-tile (32 pixel) boxes (warp-independent epilogue)
};

struct ConvTcParams {
  int32_t Tw, Th, Tn;                 // spatial tile of output pixels, Tw*Th*Tn <= 128 GEMM rows
  int32_t tiles_w, tiles_h, tiles_n;  // spatial tile grid
  int32_t n_tiles;                    // Cout tiles of BN columns
  int32_t BN, Cc, chunks_per_tap, taps, ksize, stride, stages, tmem_cols;
  int32_t epi_warp;  // warp-independent epilogue: each warp stores its own 32-row sub-box (tile must decompose)
  int32_t bres;      // TMA mode with RESIDENT weights: all K stages of the (single) N tile are loaded once per CTA, stages carry activations only
  int32_t pair;      // conv_tc_kernel_pair: a CTA pair computes a 256-row tile with tcgen05.mma.cta_group::2 (b_slot / tx_bytes are per CTA)
  int32_t fat;       // conv_tc_kernel_fat: 32 accumulator columns per epilogue warp step, 64-channel store chunks, 2 CTAs/SM
  int32_t nstg;  // staging buffers per warp of the warp-independent epilogue (1 or 2); the CTA-wide epilogue uses 2
  int32_t cw;  // epilogue chunk width in output channels (16 or 32) = inner box of the output tensor map
  uint32_t a_slot, b_slot, tx_bytes, sbo, layout_type;
  int32_t B, Hout, Wout, cout, cin;
  void* out;
  int32_t out_ct, out_co, out_f32;
  const void* res;
  int32_t res_ct, res_co;
  int32_t res_pre;  // 1: residual is a half-resolution map added BEFORE the activation (Y11_RES_PRE_UP2)
  const float* bias;
  int32_t act;
  int32_t quant;         // conv_tc_kernel_q: any of the e4m3 paths below is in use
  int32_t in_fp8;        // A and B operands are e4m3 (tcgen05.mma.kind::f8f6f4, K = 32 per instruction)
  int32_t out_esz;       // bytes per output element: 4 (fp32 logits), 2 (bf16), 1 (e4m3)
  const float* cscale;   // per-output-channel multiplier of the accumulator before the bias (fp8 dequantisation) or nullptr
  float oscale;          // multiplier before the e4m3 conversion of the output
  // compact k = 2 form (y11_conv_desc.s2d_block): K stage k loads channels [16*(kt[k]>>2), +Cc) at block tap (kt[k] & 3);
  // n_kt = number of K stages (0: the regular tap x chunk walk)
  int32_t n_kt;
  uint16_t kt[16];
  uint64_t kmask;       // bit (k_iter * Cc/16 + kk): that 16-element K step has non-zero weights (TMA path, k == 2)
  int32_t kmask_on;
  int* err_flag;
  int* tile_counter;  // [0] next tile request, [1] CTAs done (dynamic tile scheduler); nullptr = static `tile += gridDim.x` walk
  long long* trace;  // debug timeline buffer (Y11_TRACE builds only)
  uint64_t mg_ntiles, mg_tw, mg_th;  // fast_div magics for n_tiles, tiles_w, tiles_h
  // halo mode (small-channel 3x3 stride-1 layers): the (Th+2)x(Tw+2) input halo of a tile is copied ONCE by cp.async into
  // the un-swizzled core-matrix layout [cin/8][Th+2][Tw+2][8 ch]; the nine taps are nine descriptor start offsets into it
  int32_t halo;         // LSU-producer mode (1x1: pad 0; 3x3 stride 1: pad 1, Tw = 8)
  int32_t pad, n_pos;   // positions per tile = (Tw+2pad)*(Th+2pad)*Tn
  uint64_t mg_ncg;      // fast_div magic for cin/8
  const void* in;       // input view base (channel offset applied)
  int32_t in_ct, Hin, Win;
  uint32_t b_res_bytes; // resident weight region (all taps), loaded once per CTA
  uint32_t a_lbo, a_sbo;
  int32_t hbufs;        // halo-stream mode: halo tile buffers (2, or 1 where that keeps two CTAs per SM)
  int32_t hstream;      // halo-stream mode: swizzled TMA halo tiles (cin = 64 / 128, two tile buffers) + weights streamed per (tap, chunk)
  int32_t halo_tma;     // 3x3 halo tile in the 128-byte-swizzled layout (cin = 64), loaded by ONE TMA box per tile instead of cp.async
  // class-emit epilogue (Detect cv3.l.2 in single-label prediction, y11_plan_set_cls_emit): instead of storing the fp32 class
  // logits, every output row (= anchor) keeps the maximum logit and its class, and rows above the threshold are appended to the
  // image's pre-candidate list {anchor, class, logit bits, 0}
  int32_t emit;            // 1: emit mode (the output tensor is not written)
  int32_t emit_nc;         // valid columns (classes)
  int32_t emit_aoff;       // anchor index of this level's first pixel
  int32_t emit_cap;        // list entries per image
  float emit_thr;          // conservative logit threshold: sigmoid(x) > conf  =>  x > emit_thr
  int4* emit_list;         // [B][emit_cap]
  int* emit_count;         // [B]
};

// Per-layer launch variant.  -1 = the built-in heuristic (which an environment knob may override globally); the plan
// autotuner (y11_plan_autotune) times the feasible combinations of a layer on its real buffers and keeps the fastest.
struct ConvTcTune {
  int32_t lsu;       // 0: tap-by-tap TMA producer; 1: cp.async (LSU) producer where eligible; 2: TMA-halo mode (3x3, cin = 64, resident weights); 3: halo-stream mode (3x3, cin = 64 / 128, streamed weights); -1: heuristic
  int32_t epi_warp;  // bit 0: warp-independent epilogue, bit 1: fat epilogue (conv_tc_kernel_fat), bit 3: CTA-pair kernel (cta_group::2)
  int32_t cps;       // persistent CTAs per SM (1..4)
  int32_t bn_max;    // largest N tile to consider (16..256)
};
static inline ConvTcTune conv_tc_default_tune() { return ConvTcTune{-1, -1, -1, -1}; }

struct ConvTcLaunch {
  ConvTcMaps maps;
  ConvTcParams p;
  unsigned grid;
  unsigned smem_bytes;
  double flops;
  ConvTcTune variant;  // what was actually used (resolved values)
  int32_t lsu_eligible, epi_warp_possible, halo_tma_eligible, hstream_eligible;
};

int conv_tc_prepare(y11_engine* eng, const y11_conv_desc* d, ConvTcLaunch* out, const ConvTcTune* tune = nullptr);
int conv_tc_launch(const ConvTcLaunch* l, cudaStream_t s);

// ---- CUDA-core kernels (conv_simt.cu) ------------------------------------------------------------
int conv_simt_launch(const y11_conv_desc* d, cudaStream_t s);  // bring-up cross-check only
int stem_launch(const y11_stem_desc* d, cudaStream_t s);
int dwconv_launch(const y11_dwconv_desc* d, cudaStream_t s);
int sppf_launch(const y11_sppf_desc* d, cudaStream_t s);
int upsample_launch(const y11_upsample_desc* d, cudaStream_t s);

// ---- TMA-ring depthwise conv for large maps (dwconv_tma.cu) ---------------------------------------
struct DwTmaParams {
  int32_t C, CC, chunks, cg4, nsub;   // channels, channels per chunk (<= 128), chunks, CC/4, row groups per column
  int32_t B, H, W, tiles_w, tiles_h;
  uint32_t stage_tx, stage_bytes;
  int32_t stages;  // ring depth (3..6 halo tiles)
  const void* w;
  const float* bias;
  int32_t act;
  void* out;
  int32_t out_ct, out_co;
  const void* res;
  int32_t res_ct, res_co;
  int* err_flag;
  uint64_t mg_chunks, mg_tw, mg_th;  // fast_div magics of the tile decomposition
  int32_t dbg;  // probe builds (-DY11_DW_PROBE) only: bit 0 = skip the arithmetic, bit 1 = skip the loads
};
struct DwTmaLaunch {
  CUtensorMap tmap;
  DwTmaParams p;
  unsigned grid, smem_bytes;
};
bool dwconv_tma_eligible(const y11_dwconv_desc* d);
int dwconv_tma_prepare(y11_engine* eng, const y11_dwconv_desc* d, DwTmaLaunch* out);
int dwconv_tma_launch(const DwTmaLaunch* l, cudaStream_t s);

// ---- attention.cu --------------------------------------------------------------------------------
int attention_launch(const y11_attn_desc* d, cudaStream_t s);
