// Detection rasteriser on device: boxes, label backgrounds and label text drawn straight into
// uint8 BGR frames that are already in HBM, replacing the reference's per-box Python loop
//   for i in range(len(boxes)): boxes.xyxy[i].cpu()... cv2.rectangle(...); cv2.getTextSize(...); cv2.rectangle(..., -1); cv2.putText(...)
// (/root/reference/utils/visualization.py:18-106: three device->host syncs and four cv2 calls PER BOX).
//
// Output is BIT-IDENTICAL to that loop for the defaults every call site of the reference uses (line_thickness 2 - also 1 -,
// FONT_HERSHEY_SIMPLEX at scale 0.5, thickness 1; demos/detection_demo.py:100,202,292):
//   * cv2.rectangle(thickness 2) = four thick lines = a 3-pixel band per edge plus a radius-1 disc at every vertex (OpenCV
//     ThickLine: FillConvexPoly of the half-width quad + Circle caps) - the closed form is in outline_hit();
//   * the filled label background is the inclusive rectangle [x1, x1+w+10] x [y1-h-10, y1], (w, h) = cv2.getTextSize(label);
//   * the text is blitted from a glyph atlas that tools/make_glyph_atlas.py rendered WITH cv2.putText for both sub-pixel
//     phases of the pen (OpenCV advances the pen by (right-left) * round(0.5 * 65536) in 16.16 fixed point, i.e. in half
//     pixels), composed with the same pen arithmetic;
//   * the label string f'{name}: {conf:.2f}' is formatted on the device with exact round-half-even of the float's binary value
//     (what Python's format does);
//   * painter's order: detection i+1 is drawn over detection i; inside a detection box outline, then background, then text.
// One exception, stated in tests/test_gpu_draw.py: glyphs cut by the image border may differ from cv2 in a few pixels (cv2
// clips each stroke segment before rasterising it, which moves the segment by a sub-pixel amount).
//
// Work decomposition: one CTA per 64x16-pixel tile of one image.  The CTA first lists, IN ORDER, the detections whose drawn
// extent touches the tile (ballot-ordered compaction), then every thread resolves its 4 pixels against that list.
#include "ops.h"

namespace {

constexpr int kTileW = 64, kTileH = 16, kDrawThreads = 256;
constexpr int kMaxList = 320;  // detections that may touch one tile (max_det is 300)
constexpr int kMaxLabel = 48;

struct DetGeom {
  int x1, y1, x2, y2;     // truncated box
  int cls, conf100;       // class id, round-half-even(conf * 100)
  int n_chars, w_half;    // label length, label advance in half pixels
  int bg_w;               // cv2.getTextSize(label).width
};

__constant__ unsigned char kPalette[10][3] = {  // BGR tuples of utils/visualization.py:get_color
    {255, 0, 0}, {0, 255, 0}, {0, 0, 255}, {255, 255, 0}, {255, 0, 255}, {0, 255, 255}, {128, 0, 128}, {255, 165, 0}, {0, 128, 255}, {128, 255, 0}};

// round-half-even(v * 100) for a float v in [0, 1] computed on its exact binary value (Python: f'{v:.2f}')
__device__ __forceinline__ int conf_times_100(float v) {
  const unsigned u = __float_as_uint(v);
  const int e = (int)((u >> 23) & 0xff);
  unsigned long long m = (u & 0x7fffffu) | (e ? 0x800000u : 0u);
  const int sh = 150 - (e ? e : 1);  // v = m * 2^-sh
  m *= 100ull;
  if (sh <= 0) return (int)(m << (-sh));
  if (sh >= 64) return 0;
  const unsigned long long q = m >> sh, rem = m & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
  return (int)(q + ((rem > half || (rem == half && (q & 1ull))) ? 1ull : 0ull));
}

// label character k of a detection: name, then ": ", then d.dd
__device__ __forceinline__ int label_char(const y11_font& f, const DetGeom& g, int name_len, int k) {
  if (k < name_len) return (unsigned char)f.names[(size_t)g.cls * f.name_stride + k];
  switch (k - name_len) {
    case 0: return ':';
    case 1: return ' ';
    case 2: return '0' + (g.conf100 / 100) % 10;
    case 3: return '.';
    case 4: return '0' + (g.conf100 / 10) % 10;
    default: return '0' + g.conf100 % 10;
  }
}

__device__ __forceinline__ int name_length(const y11_font& f, int cls) {
  if (cls < 0 || cls >= f.nc) return 0;
  int n = 0;
  while (n < f.name_stride && n < kMaxLabel - 6 && f.names[(size_t)cls * f.name_stride + n] != 0) ++n;
  return n;
}

__device__ __forceinline__ int glyph_index(const y11_font& f, int ch) {
  const int i = ch - f.first_char;
  return (i >= 0 && i < f.n_chars) ? i : ('?' - f.first_char);
}

// cv2.rectangle(img, (x1,y1), (x2,y2), color, t) for t in {1, 2}: is pixel (x, y) painted?
__device__ __forceinline__ bool outline_hit(int x, int y, int x1, int y1, int x2, int y2, int t) {
  if (t == 1) return ((y == y1 || y == y2) && x >= x1 && x <= x2) || ((x == x1 || x == x2) && y >= y1 && y <= y2);
  const bool in_x = x >= x1 && x <= x2, in_y = y >= y1 && y <= y2;
  const bool h_band = in_x && (abs(y - y1) <= 1 || abs(y - y2) <= 1);
  const bool v_band = in_y && (abs(x - x1) <= 1 || abs(x - x2) <= 1);
  const bool h_cap = (y == y1 || y == y2) && (x == x1 - 1 || x == x2 + 1);
  const bool v_cap = (x == x1 || x == x2) && (y == y1 - 1 || y == y2 + 1);
  return h_band || v_band || h_cap || v_cap;
}

__global__ void __launch_bounds__(kDrawThreads) draw_kernel(const y11_draw_item* __restrict__ items, y11_font f, int thickness) {
  __shared__ DetGeom s_det[kMaxList];
  __shared__ int s_n;
  __shared__ int s_warp[kDrawThreads / 32];
  const y11_draw_item it = items[blockIdx.z];
  const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
  if (tx0 >= it.w || ty0 >= it.h) return;
  const int tx1 = min(tx0 + kTileW, it.w) - 1, ty1 = min(ty0 + kTileH, it.h) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int n_det = min(it.count ? *it.count : it.n, it.max_det);
  // ---- 1. ordered list of the detections that touch this tile
  for (int base = 0; base < n_det; base += kDrawThreads) {
    const int i = base + threadIdx.x;
    bool hit = false;
    DetGeom g;
    if (i < n_det) {
      const float* d = it.det + (size_t)i * 6;
      g.x1 = (int)d[0]; g.y1 = (int)d[1]; g.x2 = (int)d[2]; g.y2 = (int)d[3];   // .astype(int): truncation
      g.conf100 = conf_times_100(d[4]);
      g.cls = (int)d[5];
      const int nl = name_length(f, g.cls);
      g.n_chars = nl + 6;
      int wh = 0;
      for (int k = 0; k < g.n_chars; ++k) wh += f.advance[glyph_index(f, label_char(f, g, nl, k))];
      g.w_half = wh;
      // cv2.getTextSize: width = cvRound(sum * fontScale + thickness) = round-half-even(wh / 2 + 1)
      g.bg_w = (wh >> 1) + 1 + ((wh & 1) ? (((wh >> 1) + 1) & 1) : 0);
      const int t = thickness == 2 ? 1 : 0;
      const int ex1 = g.x1 - t, ey1 = min(g.y1 - t, g.y1 - 5 - f.base_y);
      const int ex2 = max(g.x2 + t, g.x1 + 5 + (wh >> 1) + f.cell_w), ey2 = max(g.y2 + t, g.y1 - 5 - f.base_y + f.cell_h);
      hit = ex1 <= tx1 && ex2 >= tx0 && ey1 <= ty1 && ey2 >= ty0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_n;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (hit) {
      const int slot = off + __popc(bal & ((1u << lane) - 1u));
      if (slot < kMaxList) s_det[slot] = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = s_n;
      for (int w = 0; w < kDrawThreads / 32; ++w) t += s_warp[w];
      s_n = min(t, kMaxList);
    }
    __syncthreads();
  }
  const int n_list = s_n;
  if (n_list == 0) return;
  // ---- 2. every thread resolves 4 pixels of the tile (one row segment) against the list, in painter's order
  const int py = ty0 + (threadIdx.x >> 4), px0 = tx0 + (threadIdx.x & 15) * 4;
  if (py > ty1) return;
  unsigned char* row = it.img + (size_t)py * it.pitch;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int px = px0 + j;
    if (px > tx1) break;
    int color = -1;  // -1: untouched, 0..9 palette, 10: white
    for (int l = 0; l < n_list; ++l) {
      const DetGeom& g = s_det[l];
      if (outline_hit(px, py, g.x1, g.y1, g.x2, g.y2, thickness)) color = ((g.cls % 10) + 10) % 10;
      const int by1 = g.y1 - f.text_h - 10;
      if (px >= g.x1 && px <= g.x1 + g.bg_w + 10 && py >= by1 && py <= g.y1) color = ((g.cls % 10) + 10) % 10;
      // text: baseline origin (x1 + 5, y1 - 5); glyph cells start base_y rows above the baseline and pad_x columns left of the pen
      const int gy = py - (g.y1 - 5 - f.base_y);
      const int tx = px - (g.x1 + 5 - f.pad_x);
      if (gy >= 0 && gy < f.cell_h && tx >= 0 && tx < (g.w_half >> 1) + f.cell_w) {
        const int nl = g.n_chars - 6;
        int pen = 0;
        for (int k = 0; k < g.n_chars; ++k) {
          const int gi = glyph_index(f, label_char(f, g, nl, k));
          const int gx = tx - (pen >> 1);
          if (gx >= 0 && gx < f.cell_w && ((f.glyph_bits[((size_t)gi * 2 + (pen & 1)) * f.cell_h + gy] >> gx) & 1u)) {
            color = 10;
            break;
          }
          pen += f.advance[gi];
        }
      }
    }
    if (color >= 0) {
      unsigned char* p = row + (size_t)px * 3;
      if (color == 10) { p[0] = 255; p[1] = 255; p[2] = 255; }
      else { p[0] = kPalette[color][0]; p[1] = kPalette[color][1]; p[2] = kPalette[color][2]; }
    }
  }
}

}  // namespace

extern "C" int y11_draw_detections(y11_handle h, const y11_draw_item* items, int n_items, int max_h, int max_w, const y11_font* font,
                                   int line_thickness, y11_stream s) {
  Y11_REQUIRE(h && items && font && n_items >= 1, "draw_detections: null argument");
  Y11_REQUIRE(line_thickness == 1 || line_thickness == 2, "draw_detections: line_thickness %d unsupported (1 or 2: the closed forms of "
              "cv2's thick-line rasterisation that are implemented)", line_thickness);
  Y11_REQUIRE(font->glyph_bits && font->advance && font->names && font->cell_w <= 32 && font->n_chars > ('?' - font->first_char),
              "draw_detections: bad font");
  Y11_REQUIRE(max_h >= 1 && max_w >= 1 && n_items <= 65535, "draw_detections: bad extent");
  dim3 grid((unsigned)y11_ceil_div(max_w, kTileW), (unsigned)y11_ceil_div(max_h, kTileH), (unsigned)n_items);
  draw_kernel<<<grid, kDrawThreads, 0, static_cast<cudaStream_t>(s)>>>(items, *font, line_thickness);
  Y11_CHECK_CUDA(cudaGetLastError());
  return 0;
}
