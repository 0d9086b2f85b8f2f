// GPU JPEG decode feeding the letterbox kernel: replaces the cv2.imread of the reference's loaders
// (/root/reference/utils/data_loader.py:42, ultralytics LoadImagesAndVideos) for file sources when the caller opts in
// (`predict(..., decode="nvjpeg")`): the compressed bytes cross PCIe (10-20x fewer than the decoded frame) and the BGR frame is
// produced in HBM, where y11_letterbox / the stem read it.  nvJPEG is the CUDA toolkit's decoder (library code, like cuBLAS
// would be for a plain GEMM); it is resolved with dlopen at first use so that liby11_b200.so itself has no link dependency on it -
// a box without libnvjpeg only loses this entry point (with an error), nothing else.
//
// Parity: nvJPEG's IDCT / chroma upsampling are not libjpeg-turbo's, so frames are NOT bit-identical to cv2.imread; the stated
// tolerance (tests/test_gpu_decode.py) is measured per pixel against cv2.imdecode on the same bytes.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "ops.h"

namespace {

struct NvjpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  bool ok = false;
};

NvjpegApi& api() {
  static NvjpegApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
    auto sym = [&](const char* n) { return dlsym(a.lib, n); };
    a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(sym("nvjpegCreateSimple"));
    a.Destroy = reinterpret_cast<decltype(a.Destroy)>(sym("nvjpegDestroy"));
    a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
    a.JpegStateDestroy = reinterpret_cast<decltype(a.JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
    a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(sym("nvjpegGetImageInfo"));
    a.Decode = reinterpret_cast<decltype(a.Decode)>(sym("nvjpegDecode"));
    a.ok = a.CreateSimple && a.Destroy && a.JpegStateCreate && a.JpegStateDestroy && a.GetImageInfo && a.Decode;
  });
  return a;
}

}  // namespace

struct y11_jpeg_s {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
};

extern "C" int y11_jpeg_create(y11_handle h, y11_jpeg* out) {
  Y11_REQUIRE(h && out, "jpeg_create: null argument");
  NvjpegApi& a = api();
  Y11_REQUIRE(a.ok, "jpeg_create: libnvjpeg.so.12 not found (or incomplete): GPU JPEG decode is unavailable on this box");
  y11_jpeg_s* j = new y11_jpeg_s();
  nvjpegStatus_t st = a.CreateSimple(&j->handle);
  if (st == NVJPEG_STATUS_SUCCESS) st = a.JpegStateCreate(j->handle, &j->state);
  if (st != NVJPEG_STATUS_SUCCESS) {
    if (j->handle) a.Destroy(j->handle);
    delete j;
    y11_set_error("jpeg_create: nvjpeg status %d", (int)st);
    return -2;
  }
  *out = j;
  return 0;
}

extern "C" void y11_jpeg_destroy(y11_jpeg j) {
  if (!j) return;
  NvjpegApi& a = api();
  if (j->state) a.JpegStateDestroy(j->state);
  if (j->handle) a.Destroy(j->handle);
  delete j;
}

extern "C" int y11_jpeg_info(y11_jpeg j, const uint8_t* data, size_t nbytes, int32_t* h, int32_t* w, int32_t* components) {
  Y11_REQUIRE(j && data && h && w, "jpeg_info: null argument");
  int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t ss;
  const nvjpegStatus_t st = api().GetImageInfo(j->handle, data, nbytes, &nc, &ss, ws, hs);
  Y11_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "jpeg_info: not a decodable JPEG (nvjpeg status %d)", (int)st);
  *h = hs[0];
  *w = ws[0];
  if (components) *components = nc;
  return 0;
}

extern "C" int y11_jpeg_decode(y11_jpeg j, const uint8_t* data, size_t nbytes, uint8_t* out_bgr, int32_t pitch, int32_t h, int32_t w,
                               y11_stream s) {
  Y11_REQUIRE(j && data && out_bgr, "jpeg_decode: null argument");
  Y11_REQUIRE(pitch >= 3 * w, "jpeg_decode: pitch %d < 3 * width %d", pitch, w);
  int32_t ih = 0, iw = 0;
  if (int e = y11_jpeg_info(j, data, nbytes, &ih, &iw, nullptr)) return e;
  Y11_REQUIRE(ih == h && iw == w, "jpeg_decode: output is %dx%d but the image is %dx%d", h, w, ih, iw);
  nvjpegImage_t img;
  for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { img.channel[c] = nullptr; img.pitch[c] = 0; }
  img.channel[0] = out_bgr;
  img.pitch[0] = (size_t)pitch;
  // BGR interleaved = the layout cv2.imread produces and y11_image / the letterbox kernel expect (grayscale JPEGs are expanded)
  const nvjpegStatus_t st = api().Decode(j->handle, j->state, data, nbytes, NVJPEG_OUTPUT_BGRI, &img, static_cast<cudaStream_t>(s));
  Y11_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "jpeg_decode: nvjpeg status %d", (int)st);
  return 0;
}
