// Pooling / resampling / layout kernels on NHWC bf16 (8 channels = one 128-bit access per thread).
// Reference semantics: nn.MaxPool2d(2,2) (network/ugan.py:31-37, network/blocks.py:128-134),
// F.avg_pool2d(x, 2) (network/blocks.py:101-112), nn.Upsample(x2, bilinear, align_corners=False)
// (network/blocks.py:44), the generator's input assembly (network/ugan.py:154-159).
#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

__device__ __forceinline__ uint4 ld8(const void* b, size_t off) {
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(b) + off);
}
__device__ __forceinline__ void st8(void* b, size_t off, const uint4& v) {
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(b) + off) = v;
}

// one thread per (pooled pixel, 8-channel group)
__global__ void maxpool2_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int n, int h, int w, int c) {
  pdl_prologue();
  const int cg = c >> 3, ho = h >> 1, wo = w >> 1;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int ox = (int)(t % wo); t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    const size_t base = (((size_t)b * h + 2 * oy) * w + 2 * ox) * c + g * 8;
    float v0[8], v1[8], v2[8], v3[8], m[8];
    unpack8(ld8(x, base), v0);
    unpack8(ld8(x, base + c), v1);
    unpack8(ld8(x, base + (size_t)w * c), v2);
    unpack8(ld8(x, base + (size_t)w * c + c), v3);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(fmaxf(v0[j], v1[j]), fmaxf(v2[j], v3[j]));
    st8(y, (((size_t)b * ho + oy) * wo + ox) * c + g * 8, pack8(m));
  }
}

__global__ void maxpool2_bwd_kernel(const void* __restrict__ x, const void* __restrict__ dy,
                                    const void* __restrict__ add, void* __restrict__ dx, int n, int h, int w, int c) {
  pdl_prologue();
  const int cg = c >> 3, ho = h >> 1, wo = w >> 1;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int ox = (int)(t % wo); t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    const size_t base = (((size_t)b * h + 2 * oy) * w + 2 * ox) * c + g * 8;
    const size_t offs[4] = {base, base + (size_t)c, base + (size_t)w * c, base + (size_t)w * c + c};
    float v[4][8], gy[8], o[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack8(ld8(x, offs[k]), v[k]);
    unpack8(ld8(dy, (((size_t)b * ho + oy) * wo + ox) * c + g * 8), gy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // first maximum in scan order (strict > keeps the earliest), as PyTorch's max_pool2d
      int best = 0;
      float bv = v[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][j] > bv) { bv = v[k][j]; best = k; }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? gy[j] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (add != nullptr) {
        float a[8];
        unpack8(ld8(add, offs[k]), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[k][j] += a[j];
      }
      st8(dx, offs[k], pack8(o[k]));
    }
  }
}

__global__ void avgpool2_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int n, int h, int w, int c) {
  pdl_prologue();
  const int cg = c >> 3, ho = h >> 1, wo = w >> 1;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int ox = (int)(t % wo); t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    const size_t base = (((size_t)b * h + 2 * oy) * w + 2 * ox) * c + g * 8;
    float v0[8], v1[8], v2[8], v3[8], m[8];
    unpack8(ld8(x, base), v0);
    unpack8(ld8(x, base + c), v1);
    unpack8(ld8(x, base + (size_t)w * c), v2);
    unpack8(ld8(x, base + (size_t)w * c + c), v3);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = 0.25f * ((v0[j] + v1[j]) + (v2[j] + v3[j]));
    st8(y, (((size_t)b * ho + oy) * wo + ox) * c + g * 8, pack8(m));
  }
}

// (n, h, w, c) is the FULL-resolution shape of dx; dy is (n, h/2, w/2, c)
__global__ void avgpool2_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ add, void* __restrict__ dx,
                                    int n, int h, int w, int c) {
  pdl_prologue();
  const int cg = c >> 3, ho = h >> 1, wo = w >> 1;
  const long long total = (long long)n * h * w * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int x_ = (int)(t % w); t /= w;
    const int y_ = (int)(t % h);
    const int b = (int)(t / h);
    float v[8];
    unpack8(ld8(dy, (((size_t)b * ho + (y_ >> 1)) * wo + (x_ >> 1)) * c + g * 8), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
    const size_t off = (((size_t)b * h + y_) * w + x_) * c + g * 8;
    if (add != nullptr) {
      float a[8];
      unpack8(ld8(add, off), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    st8(dx, off, pack8(v));
  }
}

// bilinear x2, align_corners=False: out[2k] = .25*in[k-1] + .75*in[k], out[2k+1] = .75*in[k] + .25*in[k+1], clamped
__device__ __forceinline__ void bil_taps(int o, int len, int& i0, int& i1, float& w0, float& w1) {
  const int k = o >> 1;
  if (o & 1) { i0 = k; i1 = min(k + 1, len - 1); w0 = 0.75f; w1 = 0.25f; }
  else       { i0 = max(k - 1, 0); i1 = k; w0 = 0.25f; w1 = 0.75f; }
}

__global__ void bilinear2_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int n, int h, int w, int c) {
  pdl_prologue();
  const int cg = c >> 3, ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int ox = (int)(t % wo); t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
    bil_taps(oy, h, y0, y1, wy0, wy1);
    bil_taps(ox, w, x0, x1, wx0, wx1);
    const size_t rb = (size_t)b * h;
    float a[8], bb[8], cc[8], d[8], o[8];
    unpack8(ld8(x, ((rb + y0) * w + x0) * c + g * 8), a);
    unpack8(ld8(x, ((rb + y0) * w + x1) * c + g * 8), bb);
    unpack8(ld8(x, ((rb + y1) * w + x0) * c + g * 8), cc);
    unpack8(ld8(x, ((rb + y1) * w + x1) * c + g * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = wy0 * (wx0 * a[j] + wx1 * bb[j]) + wy1 * (wx0 * cc[j] + wx1 * d[j]);
    st8(y, (((size_t)b * ho + oy) * wo + ox) * c + g * 8, pack8(o));
  }
}

// adjoint: (n,h,w,c) is the LOW-resolution shape of dx; dy is (n, 2h, 2w, c)
__device__ __forceinline__ void bil_adj(int k, int len, int* o, float* wt) {
  // output rows touching input k: 2k-1 (.25), 2k (.75), 2k+1 (.75), 2k+2 (.25); clamping folds the border taps
  o[0] = 2 * k - 1; wt[0] = k > 0 ? 0.25f : 0.f;
  o[1] = 2 * k;     wt[1] = 0.75f + (k == 0 ? 0.25f : 0.f);
  o[2] = 2 * k + 1; wt[2] = 0.75f + (k == len - 1 ? 0.25f : 0.f);
  o[3] = 2 * k + 2; wt[3] = k < len - 1 ? 0.25f : 0.f;
}
__global__ void bilinear2_bwd_kernel(const void* __restrict__ dy, void* __restrict__ dx, int n, int h, int w, int c) {
  pdl_prologue();
  const int cg = c >> 3, ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * h * w * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int x_ = (int)(t % w); t /= w;
    const int y_ = (int)(t % h);
    const int b = (int)(t / h);
    int oy[4], ox[4]; float wy[4], wx[4];
    bil_adj(y_, h, oy, wy);
    bil_adj(x_, w, ox, wx);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.f) continue;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        if (wx[bb] == 0.f) continue;
        float v[8];
        unpack8(ld8(dy, (((size_t)b * ho + oy[a]) * wo + ox[bb]) * c + g * 8), v);
        const float wgt = wy[a] * wx[bb];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
      }
    }
    st8(dx, (((size_t)b * h + y_) * w + x_) * c + g * 8, pack8(acc));
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int c, int hw,
                                    int c_pad) {
  pdl_prologue();
  const long long total = (long long)n * hw * c_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c_pad);
    const long long p = i / c_pad;
    const int b = (int)(p / hw);
    const int q = (int)(p % hw);
    y[i] = f2bf(ch < c ? x[((size_t)b * c + ch) * hw + q] : 0.f);
  }
}
// c_pad == 16 (the network inputs: 1 .. 16 channels): one thread assembles a pixel's 16 channels and writes them with
// two 128-bit stores (the element-per-thread kernel above issues sixteen 2-byte stores per pixel)
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int c, int hw) {
  pdl_prologue();
  const long long total = (long long)n * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int q = (int)(i - (long long)b * hw);
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = k < c ? x[((size_t)b * c + k) * hw + q] : 0.f;
    uint4* dst = reinterpret_cast<uint4*>(y + (size_t)i * 16);
    dst[0] = pack8(v);
    dst[1] = pack8(v + 8);
  }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int n, int c, int hw,
                                    int x_ld) {
  pdl_prologue();
  const long long total = (long long)n * c * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % hw);
    const long long t = i / hw;
    const int ch = (int)(t % c);
    const int b = (int)(t / c);
    y[i] = bf2f(x[((size_t)b * hw + q) * x_ld + ch]);
  }
}
__global__ void build_tsl_input_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                       __nv_bfloat16* __restrict__ y, int n, int hw, int n_modal, int c_pad) {
  pdl_prologue();
  const long long total = (long long)n * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    __nv_bfloat16* dst = y + (size_t)i * c_pad;
    if (c_pad == 16 && n_modal < 16) {
      float v[16];
      v[0] = x[i];
#pragma unroll
      for (int k = 1; k < 16; ++k) v[k] = (k <= n_modal && m) ? m[b * n_modal + k - 1] : 0.f;
      reinterpret_cast<uint4*>(dst)[0] = pack8(v);
      reinterpret_cast<uint4*>(dst)[1] = pack8(v + 8);
      continue;
    }
    dst[0] = f2bf(x[i]);
    for (int k = 0; k < n_modal; ++k) dst[1 + k] = f2bf(m ? m[b * n_modal + k] : 0.f);
    for (int k = 1 + n_modal; k < c_pad; ++k) dst[k] = f2bf(0.f);
  }
}

static inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 16LL * device_sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}
static int check_pool(int n, int h, int w, int c, bool even) {
  SMSUT_CHECK(n > 0 && h > 0 && w > 0 && c >= 8 && (c & 7) == 0, -1, "bad NHWC shape (%d,%d,%d,%d)", n, h, w, c);
  if (even) SMSUT_CHECK((h & 1) == 0 && (w & 1) == 0, -1, "2x2 pooling needs even spatial dims (%dx%d)", h, w);
  return 0;
}

}  // namespace smsut

using namespace smsut;

extern "C" int smsut_maxpool2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t st) {
  int rc = check_pool(n, h, w, c, true);
  if (rc) return rc;
  launch_pdl(maxpool2_fwd_kernel, grid_for((long long)n * (h / 2) * (w / 2) * (c / 8)), 256, 0, (cudaStream_t)st, x, y, n, h, w, c);
  count_launch();
  return launch_status("maxpool2_fwd_kernel");
}
extern "C" int smsut_maxpool2_bwd(const void* x, const void* dy, const void* add, void* dx, int32_t n, int32_t h,
                                  int32_t w, int32_t c, smsut_stream_t st) {
  int rc = check_pool(n, h, w, c, true);
  if (rc) return rc;
  launch_pdl(maxpool2_bwd_kernel, grid_for((long long)n * (h / 2) * (w / 2) * (c / 8)), 256, 0, (cudaStream_t)st, x, dy, add, dx, n, h, w, c);
  count_launch();
  return launch_status("maxpool2_bwd_kernel");
}
extern "C" int smsut_avgpool2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t st) {
  int rc = check_pool(n, h, w, c, true);
  if (rc) return rc;
  launch_pdl(avgpool2_fwd_kernel, grid_for((long long)n * (h / 2) * (w / 2) * (c / 8)), 256, 0, (cudaStream_t)st, x, y, n, h, w, c);
  count_launch();
  return launch_status("avgpool2_fwd_kernel");
}
extern "C" int smsut_avgpool2_bwd(const void* dy, const void* add, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                                  smsut_stream_t st) {
  int rc = check_pool(n, h, w, c, true);
  if (rc) return rc;
  launch_pdl(avgpool2_bwd_kernel, grid_for((long long)n * h * w * (c / 8)), 256, 0, (cudaStream_t)st, dy, add, dx, n, h, w, c);
  count_launch();
  return launch_status("avgpool2_bwd_kernel");
}
extern "C" int smsut_bilinear2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t st) {
  int rc = check_pool(n, h, w, c, false);
  if (rc) return rc;
  launch_pdl(bilinear2_fwd_kernel, grid_for((long long)n * 4 * h * w * (c / 8)), 256, 0, (cudaStream_t)st, x, y, n, h, w, c);
  count_launch();
  return launch_status("bilinear2_fwd_kernel");
}
extern "C" int smsut_bilinear2_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, smsut_stream_t st) {
  int rc = check_pool(n, h, w, c, false);
  if (rc) return rc;
  launch_pdl(bilinear2_bwd_kernel, grid_for((long long)n * h * w * (c / 8)), 256, 0, (cudaStream_t)st, dy, dx, n, h, w, c);
  count_launch();
  return launch_status("bilinear2_bwd_kernel");
}
extern "C" int smsut_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w,
                                           int32_t c_pad, smsut_stream_t st) {
  SMSUT_CHECK(c_pad >= c && n > 0 && c > 0, -1, "bad shape");
  if (c_pad == 16 && c <= 16)
    launch_pdl(nchw_to_nhwc16_kernel, grid_for((long long)n * h * w), 256, 0, (cudaStream_t)st, x, (__nv_bfloat16*)y, n, c, h * w);
  else
    launch_pdl(nchw_to_nhwc_kernel, grid_for((long long)n * h * w * c_pad), 256, 0, (cudaStream_t)st, x, (__nv_bfloat16*)y, n, c, h * w, c_pad);
  count_launch();
  return launch_status("nchw_to_nhwc_kernel");
}
extern "C" int smsut_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w,
                                           int32_t x_ld, smsut_stream_t st) {
  SMSUT_CHECK(x_ld >= c && n > 0 && c > 0, -1, "bad shape");
  launch_pdl(nhwc_to_nchw_kernel, grid_for((long long)n * h * w * c), 256, 0, (cudaStream_t)st, (const __nv_bfloat16*)x, y, n, c, h * w, x_ld);
  count_launch();
  return launch_status("nhwc_to_nchw_kernel");
}
extern "C" int smsut_build_tsl_input(const float* x, const float* m, void* y, int32_t n, int32_t hw, int32_t n_modal,
                                     int32_t c_pad, smsut_stream_t st) {
  SMSUT_CHECK(c_pad >= 1 + n_modal, -1, "c_pad too small");
  launch_pdl(build_tsl_input_kernel, grid_for((long long)n * hw), 256, 0, (cudaStream_t)st, x, m, (__nv_bfloat16*)y, n, hw, n_modal, c_pad);
  count_launch();
  return launch_status("build_tsl_input_kernel");
}
