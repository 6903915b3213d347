// Loss kernels of the coraNet trainer (reference: trainer/coraNetTrainer.py:44-58 its own DiceAndCrossEntropyLoss
// with class weights / reduction='none', :137-149 softmax_mse_loss, :277-347 the three-head slicing of the
// (1 + 3 * n_label)-channel U-Net output and the masked certain / uncertain terms).
//
//   heads_split     (npix, 1 + H*L) logits -> H tensors (npix, 1 + L): head h = [background, channels 1+hL .. (h+1)L]
//                   (`torch.cat([out_back, out_h], dim=1)`, :279-286); backward sums the H background gradients
//   wce             sum_p m_p * w[y_p] * nll_p, sum_p w[y_p], sum_p m_p   -> nn.CrossEntropyLoss(weight) in 'mean' form
//                   (num / sum w[y]) and the masked form `(CE_none * mask).sum() / (mask.sum() + 1e-16)` (:301-303)
//   softmax_mse_masked   sum_p m_p * sum_c (softmax(zs) - softmax(zt))^2, sum_p m_p  (:331-337), m = mask or 1 - mask
//
// All fp32, one thread per pixel, HBM-bound streaming passes like csrc/loss.cu.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

namespace {

constexpr int kMaxC = 8;

__device__ __forceinline__ float cn_block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;   // valid in thread 0
}

template <int C>
__device__ __forceinline__ void cn_softmax(const float* z, float* p, float& lse) {
  float m = z[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { p[c] = __expf(z[c] - m); s += p[c]; }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] *= inv;
  lse = m + __logf(s);
}

__global__ void heads_split_fwd_kernel(const float* __restrict__ z, float* __restrict__ heads, long long npix, int nlab,
                                       int nheads) {
  pdl_prologue();
  const int cin = 1 + nheads * nlab, ch = 1 + nlab;
  const long long total = npix * nheads * ch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % ch);
    const long long r = i / ch;
    const long long pix = r % npix;
    const int h = (int)(r / npix);
    heads[i] = z[pix * cin + (k == 0 ? 0 : 1 + h * nlab + (k - 1))];
  }
}

__global__ void heads_split_bwd_kernel(const float* __restrict__ dheads, float* __restrict__ dz, long long npix, int nlab,
                                       int nheads) {
  pdl_prologue();
  const int cin = 1 + nheads * nlab, ch = 1 + nlab;
  const long long total = npix * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % cin);
    const long long pix = i / cin;
    float v = 0.f;
    if (k == 0) {
      for (int h = 0; h < nheads; ++h) v += dheads[((long long)h * npix + pix) * ch];
    } else {
      const int h = (k - 1) / nlab, j = (k - 1) - h * nlab;
      v = dheads[((long long)h * npix + pix) * ch + 1 + j];
    }
    dz[i] = v;
  }
}

// BWD = false: acc[0] += sum m w[y] nll, acc[1] += sum w[y], acc[2] += sum m
// BWD = true : dz = g * m * w[y] * (softmax - onehot) / D,  D = mask_den ? acc[2] + 1e-16 : acc[1]
template <int C, bool BWD>
__global__ void __launch_bounds__(256)
wce_kernel(const float* __restrict__ z, const long long* __restrict__ y, const float* __restrict__ cw,
           const float* __restrict__ mask, float* __restrict__ acc, long long* acc_q, const float* __restrict__ gscale,
           int mask_den, float* __restrict__ dz, long long npix) {
  pdl_prologue();
  __shared__ float sh[32];
  float w[C];
#pragma unroll
  for (int c = 0; c < C; ++c) w[c] = cw != nullptr ? cw[c] : 1.f;
  float num = 0.f, sw = 0.f, sm = 0.f;
  float gs = 0.f;
  if (BWD) {
    const float den = mask_den ? acc[2] + 1e-16f : acc[1];
    gs = (gscale != nullptr ? gscale[0] : 1.f) / den;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    float a[C], p[C], lse;
#pragma unroll
    for (int c = 0; c < C; ++c) a[c] = z[i * C + c];
    cn_softmax<C>(a, p, lse);
    const long long yy = y[i];
    const float m = mask != nullptr ? mask[i] : 1.f;
    const bool ok = yy >= 0 && yy < C;
    float wy = 0.f, zy = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (ok && c == (int)yy) { wy = w[c]; zy = a[c]; }
    if (BWD) {
      const float f = gs * m * wy;
#pragma unroll
      for (int c = 0; c < C; ++c) dz[i * C + c] = f * (p[c] - ((ok && c == (int)yy) ? 1.f : 0.f));
    } else {
      num = fmaf(m * wy, lse - zy, num);
      sw += wy;
      sm += m;
    }
  }
  if (!BWD) {
    float r = cn_block_sum(num, sh);
    if (threadIdx.x == 0) acc_add_at(acc, acc_q, 0, r);
    r = cn_block_sum(sw, sh);
    if (threadIdx.x == 0) acc_add_at(acc, acc_q, 1, r);
    r = cn_block_sum(sm, sh);
    if (threadIdx.x == 0) acc_add_at(acc, acc_q, 2, r);
  }
}

// BWD = false: acc[0] += sum_p m_p sum_c (ps - pt)^2, acc[1] += sum_p m_p;   m = invert ? 1 - mask : mask
// BWD = true : dzs = g * m * 2 ps_c ((ps_c - pt_c) - sum_k ps_k (ps_k - pt_k)) / (acc[1] + 1e-16)
template <int C, bool BWD>
__global__ void __launch_bounds__(256)
mse_masked_kernel(const float* __restrict__ zs, const float* __restrict__ zt, const float* __restrict__ mask, int invert,
                  float* __restrict__ acc, long long* acc_q, const float* __restrict__ gscale, float* __restrict__ dzs,
                  long long npix) {
  pdl_prologue();
  __shared__ float sh[32];
  float num = 0.f, sm = 0.f;
  const float gs = BWD ? 2.f * (gscale != nullptr ? gscale[0] : 1.f) / (acc[1] + 1e-16f) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    float a[C], b[C], p[C], t[C], lse;
#pragma unroll
    for (int c = 0; c < C; ++c) { a[c] = zs[i * C + c]; b[c] = zt[i * C + c]; }
    cn_softmax<C>(a, p, lse);
    cn_softmax<C>(b, t, lse);
    const float m = invert ? 1.f - mask[i] : mask[i];
    float sq = 0.f, dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { const float d = p[c] - t[c]; sq = fmaf(d, d, sq); dot = fmaf(d, p[c], dot); }
    if (BWD) {
#pragma unroll
      for (int c = 0; c < C; ++c) dzs[i * C + c] = gs * m * p[c] * ((p[c] - t[c]) - dot);
    } else {
      num = fmaf(m, sq, num);
      sm += m;
    }
  }
  if (!BWD) {
    float r = cn_block_sum(num, sh);
    if (threadIdx.x == 0) acc_add_at(acc, acc_q, 0, r);
    r = cn_block_sum(sm, sh);
    if (threadIdx.x == 0) acc_add_at(acc, acc_q, 1, r);
  }
}

int cn_grid(long long n, int cap = 2368) {
  long long b = (n + 255) / 256;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace smsut

using namespace smsut;

#define CN_DISPATCH(c, expr)                                                        \
  switch (c) {                                                                      \
    case 2: { constexpr int C_ = 2; expr; } break;                                  \
    case 3: { constexpr int C_ = 3; expr; } break;                                  \
    case 4: { constexpr int C_ = 4; expr; } break;                                  \
    case 5: { constexpr int C_ = 5; expr; } break;                                  \
    case 6: { constexpr int C_ = 6; expr; } break;                                  \
    case 7: { constexpr int C_ = 7; expr; } break;                                  \
    case 8: { constexpr int C_ = 8; expr; } break;                                  \
    default: SMSUT_CHECK(false, -1, "class count %d not in 2..%d", (int)(c), kMaxC); \
  }

extern "C" int smsut_heads_split_fwd(const float* z, float* heads, int64_t npix, int32_t nlab, int32_t nheads,
                                     smsut_stream_t st) {
  SMSUT_CHECK(z && heads && npix > 0 && nlab > 0 && nheads > 0, -1, "bad heads_split args");
  launch_pdl(heads_split_fwd_kernel, cn_grid(npix * nheads * (1 + nlab)), 256, 0, (cudaStream_t)st, z, heads,
             (long long)npix, (int)nlab, (int)nheads);
  count_launch();
  return launch_status("heads_split_fwd_kernel");
}

extern "C" int smsut_heads_split_bwd(const float* dheads, float* dz, int64_t npix, int32_t nlab, int32_t nheads,
                                     smsut_stream_t st) {
  SMSUT_CHECK(dheads && dz && npix > 0 && nlab > 0 && nheads > 0, -1, "bad heads_split args");
  launch_pdl(heads_split_bwd_kernel, cn_grid(npix * (1 + nheads * nlab)), 256, 0, (cudaStream_t)st, dheads, dz,
             (long long)npix, (int)nlab, (int)nheads);
  count_launch();
  return launch_status("heads_split_bwd_kernel");
}

extern "C" int smsut_wce_fwd(const float* z, const int64_t* y, const float* cw, const float* mask, float* acc, int64_t npix,
                             int32_t c, smsut_stream_t st) {
  SMSUT_CHECK(z && y && acc && npix > 0, -1, "bad wce args");
  CN_DISPATCH(c, (launch_pdl(wce_kernel<C_, false>, cn_grid(npix, 1184), 256, 0, (cudaStream_t)st, z, (const long long*)y, cw,
                             mask, acc, det_shadow(acc), (const float*)nullptr, 0, (float*)nullptr, (long long)npix)));
  count_launch();
  return launch_status("wce_fwd_kernel");
}

extern "C" int smsut_wce_bwd(const float* z, const int64_t* y, const float* cw, const float* mask, const float* acc,
                             const float* gscale, int32_t mask_den, float* dz, int64_t npix, int32_t c, smsut_stream_t st) {
  SMSUT_CHECK(z && y && acc && dz && npix > 0, -1, "bad wce args");
  CN_DISPATCH(c, (launch_pdl(wce_kernel<C_, true>, cn_grid(npix), 256, 0, (cudaStream_t)st, z, (const long long*)y, cw, mask,
                             const_cast<float*>(acc), (long long*)nullptr, gscale, (int)mask_den, dz, (long long)npix)));
  count_launch();
  return launch_status("wce_bwd_kernel");
}

extern "C" int smsut_softmax_mse_masked_fwd(const float* zs, const float* zt, const float* mask, int32_t invert, float* acc,
                                            int64_t npix, int32_t c, smsut_stream_t st) {
  SMSUT_CHECK(zs && zt && mask && acc && npix > 0, -1, "bad softmax_mse_masked args");
  CN_DISPATCH(c, (launch_pdl(mse_masked_kernel<C_, false>, cn_grid(npix, 1184), 256, 0, (cudaStream_t)st, zs, zt, mask,
                             (int)invert, acc, det_shadow(acc), (const float*)nullptr, (float*)nullptr, (long long)npix)));
  count_launch();
  return launch_status("softmax_mse_masked_fwd_kernel");
}

extern "C" int smsut_softmax_mse_masked_bwd(const float* zs, const float* zt, const float* mask, int32_t invert,
                                            const float* acc, const float* gscale, float* dzs, int64_t npix, int32_t c,
                                            smsut_stream_t st) {
  SMSUT_CHECK(zs && zt && mask && acc && dzs && npix > 0, -1, "bad softmax_mse_masked args");
  CN_DISPATCH(c, (launch_pdl(mse_masked_kernel<C_, true>, cn_grid(npix), 256, 0, (cudaStream_t)st, zs, zt, mask, (int)invert,
                             const_cast<float*>(acc), (long long*)nullptr, gscale, dzs, (long long)npix)));
  count_launch();
  return launch_status("softmax_mse_masked_bwd_kernel");
}
