// augment.cu -- SCR's second view: RandomResizedCrop -> RandomHorizontalFlip -> ColorJitter ->
// RandomGrayscale (reference agents/scr.py:18-24, kornia 0.4.1 modules) as ONE kernel.
//
// kornia is not in the image and is not vendored by the reference, so the exact arithmetic of its
// four modules is parity-unpinned; this kernel implements their published
// definitions: bilinear crop-and-resize with corner-aligned box mapping, flip, additive
// brightness / multiplicative contrast / HSV saturation / HSV hue in a per-batch random order
// (each clamped to [0,1]), ITU-R 601 grayscale.  Per-sample random parameters are drawn on the
// host (augment.py) and passed in; the kernel is deterministic given them.
#include <math.h>

#include "common.cuh"

namespace b200ocl {
namespace {

constexpr int AUG_NPARAM = 12;
// params per sample: 0 x0, 1 y0, 2 crop_w, 3 crop_h, 4 flip, 5 jitter_on, 6 brightness delta,
//                    7 contrast factor, 8 saturation factor, 9 hue shift (turns), 10 order code, 11 gray

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
  const float d = mx - mn;
  v = mx;
  s = (mx > 0.f) ? d / mx : 0.f;
  if (d <= 0.f) { h = 0.f; return; }
  float hh;
  if (mx == r) hh = (g - b) / d;
  else if (mx == g) hh = 2.f + (b - r) / d;
  else hh = 4.f + (r - g) / d;
  hh /= 6.f;
  h = hh - floorf(hh);
}

__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float h6 = h * 6.f;
  const float i = floorf(h6);
  const float f = h6 - i;
  const float p = v * (1.f - s), q = v * (1.f - s * f), t = v * (1.f - s * (1.f - f));
  switch (((int)i) % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__global__ void __launch_bounds__(256) scr_augment_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                          const float* __restrict__ params, int N, int H, int W) {
  const int hw = H * W;
  const int total = N * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / hw, rem = i - n * hw;
    const int oy = rem / W;
    int ox = rem - oy * W;
    const float* p = params + (size_t)n * AUG_NPARAM;
    if (p[4] > 0.5f) ox = W - 1 - ox;  // flip after the crop == read the mirrored output column
    const float sx = p[0] + (W > 1 ? (float)ox * (p[2] - 1.f) / (float)(W - 1) : 0.f);
    const float sy = p[1] + (H > 1 ? (float)oy * (p[3] - 1.f) / (float)(H - 1) : 0.f);
    const float fx = fminf(fmaxf(sx, 0.f), (float)(W - 1)), fy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
    const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float ax = fx - (float)x0, ay = fy - (float)y0;
    float c[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* im = x + ((size_t)n * 3 + ch) * hw;
      const float v00 = im[y0 * W + x0], v01 = im[y0 * W + x1], v10 = im[y1 * W + x0], v11 = im[y1 * W + x1];
      c[ch] = (1.f - ay) * ((1.f - ax) * v00 + ax * v01) + ay * ((1.f - ax) * v10 + ax * v11);
    }
    if (p[5] > 0.5f) {
      int code = (int)p[10];  // four 2-bit op ids, first op in the low bits
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        const int op = code & 3;
        code >>= 2;
        if (op == 0) {
          c[0] = clamp01(c[0] + p[6]); c[1] = clamp01(c[1] + p[6]); c[2] = clamp01(c[2] + p[6]);
        } else if (op == 1) {
          c[0] = clamp01(c[0] * p[7]); c[1] = clamp01(c[1] * p[7]); c[2] = clamp01(c[2] * p[7]);
        } else {
          float h, s, v;
          rgb_to_hsv(c[0], c[1], c[2], h, s, v);
          if (op == 2) s = clamp01(s * p[8]);
          else { h = h + p[9]; h = h - floorf(h); }
          hsv_to_rgb(h, s, v, c[0], c[1], c[2]);
        }
      }
    }
    if (p[11] > 0.5f) {
      const float g = 0.299f * c[0] + 0.587f * c[1] + 0.114f * c[2];
      c[0] = c[1] = c[2] = g;
    }
    // write at the un-mirrored output position
    const int wx = rem - oy * W;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out[((size_t)n * 3 + ch) * hw + oy * W + wx] = c[ch];
  }
}

}  // namespace
}  // namespace b200ocl

extern "C" int b200ocl_scr_augment(const float* x, float* out, const float* params, int N, int H, int W, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(N >= 0 && H >= 1 && W >= 1, "need N >= 0, H,W >= 1");
  if (N == 0) return B200OCL_OK;
  B200OCL_CHECK_ARG(x && out && params && x != out, "null or aliased pointers");
  const int total = N * H * W;
  int blocks = (total + 255) / 256;
  if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
  B200OCL_PROF("scr_augment", 24.0 * total, stream);
  scr_augment_kernel<<<blocks, 256, 0, stream>>>(x, out, params, N, H, W);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}
