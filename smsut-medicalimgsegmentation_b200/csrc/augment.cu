// Input pipeline on the GPU: the reference augments every slice on the host with PIL /
// torchvision / elasticdeform in 6 DataLoader workers (data_loader/externalTransforms.py:46-101 applied by
// data_loader/balanceLoader.py:66-76 with the transforms of data_loader/baseLoader.py:87-112) -- a few hundred slices
// per second, two orders of magnitude below what the training step consumes on a B200.  Here the whole u8 dataset is
// resident in HBM and ONE launch per batch gathers, augments and normalises it.
//
// One CTA owns one plane (image or label) of one slice and keeps it in shared memory for all three joint transforms,
// which the reference applies one after the other with a u8 image in between:
//     gather -> [ping]  JointRotate  -> [pong]  JointElasticDeform -> [ping]  JointRandomResizedCrop -> registers
//            -> (gamma LUT) -> ToTensor + Normalize(0.5, 0.5) / MaskToTensor -> fp32 image / int64 label in HBM
// HBM traffic per slice: 2 x H*W bytes in, H*W * (4 + 8) bytes out -- the algorithmic minimum for this output format.
//
// Sampling rules (every stage maps an OUTPUT pixel centre to a source coordinate, "pull" form):
//   image: bilinear, 4 taps, u8 result = floor(v + 0.5);   label: nearest, round-half-to-even (torch grid_sample's rule)
//   rotate: taps outside the image are 0 (PIL's fill);  resized crop: taps clamp to the crop box (crop, then resize);
//   elastic: nearest for BOTH planes (the reference passes order=[0, 0]), outside = 0 (elasticdeform mode 'constant').
// The elastic displacement field is a cubic B-spline through `points` x `points` control displacements ~ N(0, sigma)
// (elasticdeform.deform_random_grid; that package is not installed offline -- its published algorithm is restated:
// control points span the image corner to corner, per-pixel displacement by cubic B-spline interpolation with mirrored
// ends; the host pre-filters the control values into B-spline coefficients, externalTransforms.py).
#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

constexpr int kAugThreads = 1024;
constexpr int kAugMaxPoints = 5;

// per-slice parameters as laid out by data_loader/externalTransforms.py::pack_params (SMSUT_AUG_PARAM_FLOATS floats)
struct AugParams {
  float rot_on, m00, m01, m02, m10, m11, m12;   // source = M * (x + .5, y + .5, 1) - .5   (inverse rotation)
  float ela_on;                                 // 1: apply the elastic deformation
  float crop_on, ci, cj, ch, cw;                // crop box top, left, height, width (pixels)
  float gamma_on, gamma;                        // RandomGammaCorrection (image only)
  float points;                                 // control points per axis (<= kAugMaxPoints)
  float coef[2 * kAugMaxPoints * kAugMaxPoints];  // B-spline coefficients of the (dy, dx) control displacements
};
static_assert(sizeof(AugParams) == SMSUT_AUG_PARAM_FLOATS * sizeof(float), "AugParams layout");

__device__ __forceinline__ int mirror_idx(int i, int n) {
  // scipy 'mirror' extension: ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ...
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}

__device__ __forceinline__ void bspline3(float t, float* w) {
  const float t2 = t * t, t3 = t2 * t;
  w[0] = (1.f - 3.f * t + 3.f * t2 - t3) * (1.f / 6.f);
  w[1] = (4.f - 6.f * t2 + 3.f * t3) * (1.f / 6.f);
  w[2] = (1.f + 3.f * t + 3.f * t2 - 3.f * t3) * (1.f / 6.f);
  w[3] = t3 * (1.f / 6.f);
}

template <bool CLAMP>
__device__ __forceinline__ uint8_t sample_bilinear(const uint8_t* src, int h, int w, float sy, float sx, int y0c,
                                                   int x0c, int y1c, int x1c) {
  // CLAMP: taps clamp to the box [y0c, y1c] x [x0c, x1c]; else taps outside the image contribute 0
  const float fy = floorf(sy), fx = floorf(sx);
  const int y0 = (int)fy, x0 = (int)fx;
  const float ay = sy - fy, ax = sx - fx;
  float v = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      int yy = y0 + dy, xx = x0 + dx;
      const float wgt = (dy ? ay : 1.f - ay) * (dx ? ax : 1.f - ax);
      if (CLAMP) {
        yy = min(max(yy, y0c), y1c);
        xx = min(max(xx, x0c), x1c);
        v = fmaf(wgt, (float)src[yy * w + xx], v);
      } else if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        v = fmaf(wgt, (float)src[yy * w + xx], v);
      }
    }
  return (uint8_t)min(255.f, floorf(v + 0.5f));
}

template <bool CLAMP>
__device__ __forceinline__ uint8_t sample_nearest(const uint8_t* src, int h, int w, float sy, float sx, int y0c,
                                                  int x0c, int y1c, int x1c) {
  int yy = __float2int_rn(sy), xx = __float2int_rn(sx);
  if (CLAMP) {
    yy = min(max(yy, y0c), y1c);
    xx = min(max(xx, x0c), x1c);
    return src[yy * w + xx];
  }
  return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[yy * w + xx] : (uint8_t)0;
}

// grid = (n, 2): blockIdx.y == 0 the image plane, 1 the label plane
__global__ void __launch_bounds__(kAugThreads, 1)
augment_kernel(const uint8_t* __restrict__ images, const uint8_t* __restrict__ labels,
               const long long* __restrict__ index, const AugParams* __restrict__ params, float* __restrict__ x_out,
               long long* __restrict__ y_out, int h, int w) {
  pdl_prologue();
  extern __shared__ uint8_t smem[];
  const int hw = h * w;
  uint8_t* ping = smem;
  uint8_t* pong = smem + ((hw + 15) & ~15);
  const int n = blockIdx.x;
  const bool is_label = blockIdx.y == 1;
  const uint8_t* src = (is_label ? labels : images) + (size_t)index[n] * hw;
  __shared__ AugParams p;
  for (int i = threadIdx.x; i < (int)(sizeof(AugParams) / 4); i += blockDim.x)
    reinterpret_cast<float*>(&p)[i] = reinterpret_cast<const float*>(params + n)[i];
  // gather: 16-byte loads when the plane allows it
  if ((hw & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = threadIdx.x; i < hw / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(ping)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) ping[i] = src[i];
  }
  __syncthreads();
  uint8_t* cur = ping;
  uint8_t* nxt = pong;
  // ---- JointRotate (externalTransforms.py:58-68): image bilinear, label nearest, outside = 0
  if (p.rot_on != 0.f) {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      const float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
      const float sx = fmaf(p.m00, cx, fmaf(p.m01, cy, p.m02)) - 0.5f;
      const float sy = fmaf(p.m10, cx, fmaf(p.m11, cy, p.m12)) - 0.5f;
      nxt[i] = is_label ? sample_nearest<false>(cur, h, w, sy, sx, 0, 0, 0, 0)
                        : sample_bilinear<false>(cur, h, w, sy, sx, 0, 0, 0, 0);
    }
    __syncthreads();
    uint8_t* t = cur; cur = nxt; nxt = t;
  }
  // ---- JointElasticDeform (externalTransforms.py:71-94): nearest for both planes, outside = 0
  if (p.ela_on != 0.f) {
    const int np = (int)p.points;
    const float uy_scale = h > 1 ? (float)(np - 1) / (float)(h - 1) : 0.f;
    const float ux_scale = w > 1 ? (float)(np - 1) / (float)(w - 1) : 0.f;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      const float uy = (float)y * uy_scale, ux = (float)x * ux_scale;
      const float fy = floorf(uy), fx = floorf(ux);
      float wy[4], wx[4];
      bspline3(uy - fy, wy);
      bspline3(ux - fx, wx);
      float dy = 0.f, dx = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int iy = mirror_idx((int)fy - 1 + a, np);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int ix = mirror_idx((int)fx - 1 + b, np);
          const float wgt = wy[a] * wx[b];
          dy = fmaf(wgt, p.coef[iy * np + ix], dy);
          dx = fmaf(wgt, p.coef[np * np + iy * np + ix], dx);
        }
      }
      nxt[i] = sample_nearest<false>(cur, h, w, (float)y + dy, (float)x + dx, 0, 0, 0, 0);
    }
    __syncthreads();
    uint8_t* t = cur; cur = nxt; nxt = t;
  }
  // ---- JointRandomResizedCrop (externalTransforms.py:46-55) + gamma + ToTensor / Normalize / MaskToTensor
  const bool crop = p.crop_on != 0.f;
  const int y0c = crop ? (int)p.ci : 0, x0c = crop ? (int)p.cj : 0;
  const int y1c = crop ? (int)(p.ci + p.ch) - 1 : h - 1, x1c = crop ? (int)(p.cj + p.cw) - 1 : w - 1;
  const float sy_scale = crop ? p.ch / (float)h : 1.f, sx_scale = crop ? p.cw / (float)w : 1.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    uint8_t v;
    if (crop) {
      const float sy = p.ci + ((float)y + 0.5f) * sy_scale - 0.5f;
      const float sx = p.cj + ((float)x + 0.5f) * sx_scale - 0.5f;
      v = is_label ? sample_nearest<true>(cur, h, w, sy, sx, y0c, x0c, y1c, x1c)
                   : sample_bilinear<true>(cur, h, w, sy, sx, y0c, x0c, y1c, x1c);
    } else {
      v = cur[i];
    }
    if (is_label) {
      y_out[(size_t)n * hw + i] = (long long)v;
    } else {
      float f = (float)v;
      if (p.gamma_on != 0.f)      // torchvision adjust_gamma on a PIL image: int((255 + 1 - 1e-3) * (v / 255) ** gamma)
        f = floorf((255.f + 1.f - 1e-3f) * powf(f / 255.f, p.gamma));
      // ToTensor (IEEE division by 255, as torch's byte -> float .div(255)), Normalize(mean 0.5, std 0.5): bit-exact
      x_out[(size_t)n * hw + i] = (f / 255.f - 0.5f) / 0.5f;
    }
  }
}

}  // namespace smsut

using namespace smsut;

extern "C" int smsut_augment_batch(const uint8_t* images, const uint8_t* labels, const int64_t* index,
                                   const float* params, float* x_out, int64_t* y_out, int32_t n, int32_t h, int32_t w,
                                   smsut_stream_t st) {
  SMSUT_CHECK(images && labels && index && params && x_out && y_out, -1, "augment_batch: null pointer");
  SMSUT_CHECK(n > 0 && h > 0 && w > 0, -1, "augment_batch: empty batch");
  const size_t plane = ((size_t)h * w + 15) & ~(size_t)15;
  SMSUT_CHECK(2 * plane <= 200u * 1024u, -1,
              "augment_batch: a %dx%d u8 plane (x2) does not fit the CTA's shared memory (limit 320x320)", h, w);
  static bool attr_set = false;
  if (!attr_set) {
    SMSUT_CUDA_OK(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  launch_pdl(augment_kernel, dim3((unsigned)n, 2), kAugThreads, 2 * plane, (cudaStream_t)st, images, labels,
             (const long long*)index, reinterpret_cast<const AugParams*>(params), x_out, (long long*)y_out, (int)h, (int)w);
  count_launch();
  return launch_status("augment_kernel");
}
