// Loss kernels of the SMSUT training step (fp32, HBM-bound, 128-bit accesses where the layout allows).
//
// Reference semantics:
//   DiceAndCrossEntropyLoss / SoftDiceLoss / get_tp_fp_fn_tn   misc/loss.py:8-63 (batch_dice=True, bg dropped)
//   consistency_loss (argmax pseudo labels)                    trainer/uganConsisTrainer.py:45-53
//   L1 cycle loss, adversarial means, modality CE              trainer/uganConsisTrainer.py:129-177
//   gradient_penalty                                           trainer/uganShp0Trainer.py:127-134
//   Normalize(2), PatchNCELoss                                 network/networks.py:234-243, network/patchnce.py:13-51
#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

constexpr int kMaxC = 8;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  // sh: >= 32 floats
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (wid == 0) r = warp_sum(r);
  return r;  // valid in warp 0
}

// softmax over C (<= 8) values held in registers
template <int C>
__device__ __forceinline__ void softmax_c(const float* z, float* p, float& lse) {
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
template <int C>
__device__ __forceinline__ int argmax_first(const float* z) {
  int b = 0; float bv = z[0];
#pragma unroll
  for (int c = 1; c < C; ++c) if (z[c] > bv) { bv = z[c]; b = c; }
  return b;
}

// Each thread handles 4 consecutive pixels = 4*C floats = C float4 loads (C*16 bytes, 16-byte aligned).
template <int C>
__global__ void __launch_bounds__(256)
dice_ce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                   const float* __restrict__ label_logits, float* __restrict__ acc, long long* acc_q, long long npix) {
  pdl_prologue();
  __shared__ float sh[32];
  float tp[C], fp[C], fn[C], ce = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) tp[c] = fp[c] = fn[c] = 0.f;
  const long long nquad = npix >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += (long long)gridDim.x * blockDim.x) {
    float z[4 * C];
    const float4* src = reinterpret_cast<const float4*>(logits + q * 4 * C);
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const float4 v = src[i];
      z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
    }
    int y[4];
    if (labels != nullptr) {
      const longlong2* lp = reinterpret_cast<const longlong2*>(labels + q * 4);
      const longlong2 a = lp[0], b = lp[1];
      y[0] = (int)a.x; y[1] = (int)a.y; y[2] = (int)b.x; y[3] = (int)b.y;
    } else {
      float t[4 * C];
      const float4* ls = reinterpret_cast<const float4*>(label_logits + q * 4 * C);
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float4 v = ls[i];
        t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = argmax_first<C>(t + k * C);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float p[C], lse;
      softmax_c<C>(z + k * C, p, lse);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const bool hit = (y[k] == c);
        tp[c] += hit ? p[c] : 0.f;
        fp[c] += hit ? 0.f : p[c];
        fn[c] += hit ? (1.f - p[c]) : 0.f;
        ce += hit ? (lse - z[k * C + c]) : 0.f;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float r;
    r = block_sum(tp[c], sh); if (threadIdx.x == 0) acc_add_at(acc, acc_q, c, r);
    r = block_sum(fp[c], sh); if (threadIdx.x == 0) acc_add_at(acc, acc_q, C + c, r);
    r = block_sum(fn[c], sh); if (threadIdx.x == 0) acc_add_at(acc, acc_q, 2 * C + c, r);
  }
  const float r = block_sum(ce, sh);
  if (threadIdx.x == 0) acc_add_at(acc, acc_q, 3 * C, r);
}

__global__ void dice_ce_finish_kernel(const float* __restrict__ acc, float* __restrict__ loss, float inv_npix, int c,
                                      float w_dc, float w_ce) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float smooth = 1e-5f;
  float dsum = 0.f;
  for (int k = 1; k < c; ++k) {
    const float tp = acc[k], fp = acc[c + k], fn = acc[2 * c + k];
    dsum += (2.f * tp + smooth) / (2.f * tp + fp + fn + smooth + 1e-8f);
  }
  const float dc_loss = 1.f - dsum / (float)(c - 1);
  loss[0] = w_dc * dc_loss + w_ce * acc[3 * c] * inv_npix;
}

template <int C>
__global__ void __launch_bounds__(256)
dice_ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                   const float* __restrict__ label_logits, const float* __restrict__ acc,
                   const float* __restrict__ gscale, float scale, float* __restrict__ dlogits, long long npix,
                   float inv_npix, float w_dc, float w_ce) {
  pdl_prologue();
  // per-class dice derivative coefficients: d(dc_c)/dp = (2[y=c]*U - I)/U^2
  float I[C], U[C];
  const float smooth = 1e-5f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float tp = acc[c], fp = acc[C + c], fn = acc[2 * C + c];
    I[c] = 2.f * tp + smooth;
    U[c] = 2.f * tp + fp + fn + smooth + 1e-8f;
  }
  const float gs = (gscale ? gscale[0] : 1.f) * scale;
  const float kd = -w_dc / (float)(C - 1);
  const long long nquad = npix >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += (long long)gridDim.x * blockDim.x) {
    float z[4 * C], o[4 * C];
    const float4* src = reinterpret_cast<const float4*>(logits + q * 4 * C);
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const float4 v = src[i];
      z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
    }
    int y[4];
    if (labels != nullptr) {
      const longlong2* lp = reinterpret_cast<const longlong2*>(labels + q * 4);
      const longlong2 a = lp[0], b = lp[1];
      y[0] = (int)a.x; y[1] = (int)a.y; y[2] = (int)b.x; y[3] = (int)b.y;
    } else {
      float t[4 * C];
      const float4* ls = reinterpret_cast<const float4*>(label_logits + q * 4 * C);
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float4 v = ls[i];
        t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = argmax_first<C>(t + k * C);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float p[C], lse, G[C], dot = 0.f;
      softmax_c<C>(z + k * C, p, lse);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float hit = (y[k] == c) ? 1.f : 0.f;
        G[c] = (c == 0) ? 0.f : kd * (2.f * hit * U[c] - I[c]) / (U[c] * U[c]);
        dot += G[c] * p[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float hit = (y[k] == c) ? 1.f : 0.f;
        o[k * C + c] = gs * (p[c] * (G[c] - dot) + w_ce * (p[c] - hit) * inv_npix);
      }
    }
    float4* dst = reinterpret_cast<float4*>(dlogits + q * 4 * C);
#pragma unroll
    for (int i = 0; i < C; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
}

// mean over (pixels, classes) of (softmax(zs) - softmax(zt))^2 : mean-teacher consistency
// (trainer/meanTeacherTrainer.py:124-130).  The teacher logits are constants.
template <int C, bool BWD>
__global__ void __launch_bounds__(256)
softmax_mse_kernel(const float* __restrict__ zs, const float* __restrict__ zt, float* __restrict__ out, long long* out_q,
                   const float* __restrict__ gscale, float scale, float* __restrict__ dzs, long long npix) {
  pdl_prologue();
  __shared__ float sh[32];
  float acc = 0.f;
  const float gs = BWD ? (gscale ? gscale[0] : 1.f) * scale * 2.f : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    float a[C], b[C], p[C], t[C], lse;
#pragma unroll
    for (int c = 0; c < C; ++c) { a[c] = zs[i * C + c]; b[c] = zt[i * C + c]; }
    softmax_c<C>(a, p, lse);
    softmax_c<C>(b, t, lse);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { const float d = p[c] - t[c]; acc = fmaf(d, d, acc); dot = fmaf(d, p[c], dot); }
    if (BWD) {
#pragma unroll
      for (int c = 0; c < C; ++c) dzs[i * C + c] = gs * p[c] * ((p[c] - t[c]) - dot);
    }
  }
  if (!BWD) {
    const float r = block_sum(acc, sh);
    if (threadIdx.x == 0) acc_add(out, out_q, r * scale);
  }
}

template <int C>
__global__ void argmax_kernel(const float* __restrict__ logits, long long* __restrict__ out, long long npix) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    float z[C];
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = logits[i * C + c];
    out[i] = argmax_first<C>(z);
  }
}

// Validation metric path (trainer/baseTrainer.py:207-252, misc/utils.py:180-203): confusion counts of
// argmax(logits) against the labels, conf[label][prediction] += 1, accumulated per block in shared memory.
// Dice_k = 2 conf[k][k] / (row_k + col_k): identical integers to the reference's per-organ numpy / medpy counts.
template <int C>
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ logits,
                                                        const long long* __restrict__ labels,
                                                        unsigned long long* __restrict__ conf, long long npix) {
  pdl_prologue();
  __shared__ unsigned int sh[C * C];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    float z[C];
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = logits[i * C + c];
    const int pred = argmax_first<C>(z);
    const long long y = labels[i];
    if (y >= 0 && y < C) atomicAdd(&sh[(int)y * C + pred], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (sh[i]) atomicAdd(conf + i, (unsigned long long)sh[i]);
}

__global__ void __launch_bounds__(256) l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ out, long long* out_q, long long count,
                                                     float scale) {
  pdl_prologue();
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    s += fabsf(a[i] - b[i]);
  const float r = block_sum(s, sh);
  if (threadIdx.x == 0) acc_add(out, out_q, r * scale);
}
__global__ void l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                              const float* __restrict__ gscale, float scale, float* __restrict__ da, long long count) {
  pdl_prologue();
  const float gs = (gscale ? gscale[0] : 1.f) * scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    da[i] = d > 0.f ? gs : (d < 0.f ? -gs : 0.f);
  }
}
__global__ void __launch_bounds__(256) sum_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                  long long* out_q, long long count, float scale) {
  pdl_prologue();
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    s += x[i];
  const float r = block_sum(s, sh);
  if (threadIdx.x == 0) acc_add(out, out_q, r * scale);
}
__global__ void fill_kernel(float* __restrict__ x, long long count, float v) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}

// dx = dy * (1 - y^2)  (backward of the tanh fused into the translation head, network/ugan.py:73,82)
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                long long count) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * (1.f - y[i] * y[i]);
}
__global__ void fill_scaled_kernel(float* __restrict__ x, long long count, const float* __restrict__ gscale,
                                   float scale) {
  pdl_prologue();
  const float v = (gscale ? gscale[0] : 1.f) * scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
// out = a*x + (1-a)*y with one coefficient per sample (x_hat of the gradient penalty, uganConsisTrainer.py:139)
__global__ void lerp_rows_kernel(const float* __restrict__ alpha, const float* __restrict__ x,
                                 const float* __restrict__ y, float* __restrict__ out, long long per, long long total) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float a = alpha[i / per];
    out[i] = a * x[i] + (1.f - a) * y[i];
  }
}

// small-row cross entropy (rows <= a few hundred, c <= 8): one block
__global__ void ce_rows_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                   float* __restrict__ out, long long* out_q, int rows, int c, float scale) {
  pdl_prologue();
  __shared__ float sh[32];
  float s = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* z = logits + (size_t)r * c;
    float m = z[0];
    for (int k = 1; k < c; ++k) m = fmaxf(m, z[k]);
    float e = 0.f;
    for (int k = 0; k < c; ++k) e += expf(z[k] - m);
    s += m + logf(e) - z[target[r]];
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) acc_add(out, out_q, t * scale / (float)rows);
}
__global__ void ce_rows_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                   const float* __restrict__ gscale, float scale, float* __restrict__ dlogits,
                                   int rows, int c) {
  pdl_prologue();
  const float gs = (gscale ? gscale[0] : 1.f) * scale / (float)rows;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const float* z = logits + (size_t)r * c;
    float m = z[0];
    for (int k = 1; k < c; ++k) m = fmaxf(m, z[k]);
    float e = 0.f;
    for (int k = 0; k < c; ++k) e += expf(z[k] - m);
    const float inv = 1.f / e;
    for (int k = 0; k < c; ++k)
      dlogits[(size_t)r * c + k] = gs * (expf(z[k] - m) * inv - (k == (int)target[r] ? 1.f : 0.f));
  }
}

// gradient penalty: the squared norms are summed by (sample, split) blocks (one block per sample left 16 blocks walking
// 65 536 elements each: 57 us on the critical path of the discriminator phase), the loss kernel takes the roots
__global__ void __launch_bounds__(256) gp_norm_kernel(const float* __restrict__ g, float* norm2, long long* norm2_q,
                                                      long long per, int splits) {
  pdl_prologue();
  __shared__ float sh[32];
  const float* p = g + (size_t)blockIdx.x * per;
  const long long chunk = (per + splits - 1) / splits;
  const long long i0 = (long long)blockIdx.y * chunk;
  long long i1 = i0 + chunk;
  if (i1 > per) i1 = per;
  float s = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) s = fmaf(p[i], p[i], s);
  const float r = block_sum(s, sh);
  if (threadIdx.x == 0) acc_add_at(norm2, norm2_q, blockIdx.x, r);
}
// norm[i] = sqrt(norm2[i]); out[0] += scale * mean_i (norm_i - 1)^2
__global__ void gp_loss_kernel(const float* __restrict__ norm2, float* __restrict__ norm, float* __restrict__ out,
                               long long* out_q, int b, float scale) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float s = 0.f;
  for (int i = 0; i < b; ++i) {
    const float nr = sqrtf(norm2[i]);
    norm[i] = nr;
    s += (nr - 1.f) * (nr - 1.f);
  }
  acc_add(out, out_q, scale * s / (float)b);
}
__global__ void gp_bwd_kernel(const float* __restrict__ g, const float* __restrict__ norm,
                              const float* __restrict__ gscale, float scale, float* __restrict__ u, int b,
                              long long per) {
  pdl_prologue();
  const long long total = (long long)b * per;
  const float gs = (gscale ? gscale[0] : 1.f) * scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i / per);
    const float nr = norm[s];
    const float coef = gs * 2.f * (nr - 1.f) / ((float)b * fmaxf(nr, 1e-30f));
    u[i] = coef * g[i];
  }
}

// PatchNCE sampling
__global__ void gather_rows_kernel(const uint4* __restrict__ feat, const long long* __restrict__ ids,
                                   uint4* __restrict__ out, int n, int hw, int cvec, int nids) {
  pdl_prologue();
  const long long total = (long long)n * nids * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    const long long r = i / cvec;
    const int k = (int)(r % nids);
    const int b = (int)(r / nids);
    out[i] = feat[((size_t)b * hw + ids[k]) * cvec + v];
  }
}
__global__ void scatter_rows_add_kernel(const uint4* __restrict__ dout, const long long* __restrict__ ids,
                                        uint4* __restrict__ dfeat, int n, int hw, int cvec, int nids) {
  pdl_prologue();
  const long long total = (long long)n * nids * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    const long long r = i / cvec;
    const int k = (int)(r % nids);
    const int b = (int)(r / nids);
    uint4* dst = dfeat + ((size_t)b * hw + ids[k]) * cvec + v;
    float a[8], d[8];
    unpack8(*dst, a);
    unpack8(dout[i], d);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += d[j];
    *dst = pack8(a);
  }
}

// L2 normalisation, one warp per row
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norm,
                                  int rows, int c) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = x + (size_t)row * c;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s = fmaf(p[i], p[i], s);
  s = warp_sum(s);
  const float r = sqrtf(s);
  const float inv = 1.f / (r + 1e-7f);
  for (int i = lane; i < c; i += 32) y[(size_t)row * c + i] = p[i] * inv;
  if (lane == 0) norm[row] = r;
}
// dx = (dy - y * <dy,y> * (r+eps)/r) / (r+eps); written as bf16 (feeds the tensor-core dgrad/wgrad)
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ norm, __nv_bfloat16* __restrict__ dx, int rows, int c) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* g = dy + (size_t)row * c;
  const float* yy = y + (size_t)row * c;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s = fmaf(g[i], yy[i], s);
  s = warp_sum(s);
  const float r = norm[row];
  const float inv = 1.f / (r + 1e-7f);
  const float k = s * (r + 1e-7f) / fmaxf(r, 1e-30f);
  for (int i = lane; i < c; i += 32) dx[(size_t)row * c + i] = f2bf((g[i] - yy[i] * k) * inv);
}

// PatchNCE: one 4-warp block per query row, each warp owns a quarter of the negatives; c must be a multiple of 32
// and <= 512.  (One warp per row left 7 warps per SM walking 128 negatives one after the other: 130-175 us for a
// 34 MFLOP problem.)  kJ negatives per step: their dot products and butterfly reductions are independent streams.
template <bool BWD>
__global__ void __launch_bounds__(128)
patchnce_kernel(const float* __restrict__ q, const float* __restrict__ k, float* __restrict__ loss_rows,
                float* __restrict__ out, long long* out_q, const float* __restrict__ gscale, float scale,
                float* __restrict__ dq, int groups, int np, int c, float inv_t) {
  pdl_prologue();
  __shared__ float sh_m[4], sh_s[4];
  __shared__ float sh_acc[BWD ? 4 * 512 : 4];
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = groups * np;
  const int grp = row / np, self = row - grp * np;
  constexpr int kMaxPer = 16;
  constexpr int kJ = 8;
  const int per = c >> 5;
  const int jq = (np + 3) / 4;
  const int jb = warp * jq, je = (jb + jq < np) ? jb + jq : np;
  float qv[kMaxPer];
  for (int i = 0; i < per; ++i) qv[i] = q[(size_t)row * c + lane + 32 * i];
  // positive logit (every warp: 2 KB of L1-resident loads)
  float pos = 0.f;
  for (int i = 0; i < per; ++i) pos = fmaf(qv[i], k[(size_t)row * c + lane + 32 * i], pos);
  pos = warp_sum(pos) * inv_t;
  // pass 1: online log-sum-exp over this warp's negatives (warp 0 also counts the positive)
  float m = warp == 0 ? pos : -3.0e38f, s = warp == 0 ? 1.f : 0.f;
  for (int j0 = jb; j0 < je; j0 += kJ) {
    float d[kJ];
#pragma unroll
    for (int u = 0; u < kJ; ++u) {
      const int j = j0 + u < je ? j0 + u : je - 1;
      const float* kr = k + ((size_t)grp * np + j) * c;
      float a = 0.f;
      for (int i = 0; i < per; ++i) a = fmaf(qv[i], kr[lane + 32 * i], a);
      d[u] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < kJ; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
#pragma unroll
    for (int u = 0; u < kJ; ++u) {
      const int j = j0 + u;
      if (j >= je) break;
      const float dj = j == self ? -10.f * inv_t : d[u] * inv_t;
      const float mn = fmaxf(m, dj);
      s = s * __expf(m - mn) + __expf(dj - mn);
      m = mn;
    }
  }
  if (lane == 0) { sh_m[warp] = m; sh_s[warp] = s; }
  __syncthreads();
  float mg = fmaxf(fmaxf(sh_m[0], sh_m[1]), fmaxf(sh_m[2], sh_m[3]));
  float sg = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) sg += sh_s[w] * __expf(sh_m[w] - mg);
  const float lse = mg + __logf(sg);
  if (!BWD) {
    if (threadIdx.x == 0) {
      loss_rows[row] = lse - pos;
      acc_add(out, out_q, scale * (lse - pos) / (float)rows);
    }
    return;
  }
  // pass 2: dq = coef * inv_t * ( sum_{j != self} p_j k_j + (p_pos - 1) k_row ), each warp its own negatives
  const float coef = (gscale ? gscale[0] : 1.f) * scale / (float)rows * inv_t;
  float acc[kMaxPer];
  const float ppos = __expf(pos - lse);
  for (int i = 0; i < per; ++i) acc[i] = warp == 0 ? (ppos - 1.f) * k[(size_t)row * c + lane + 32 * i] : 0.f;
  constexpr int kJ2 = 4;
  for (int j0 = jb; j0 < je; j0 += kJ2) {
    float d[kJ2];
#pragma unroll
    for (int u = 0; u < kJ2; ++u) {
      const int j = j0 + u < je ? j0 + u : je - 1;
      const float* kr = k + ((size_t)grp * np + j) * c;
      float a = 0.f;
      for (int i = 0; i < per; ++i) a = fmaf(qv[i], kr[lane + 32 * i], a);
      d[u] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < kJ2; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
#pragma unroll
    for (int u = 0; u < kJ2; ++u) {
      const int j = j0 + u;
      if (j >= je || j == self) continue;
      const float pj = __expf(d[u] * inv_t - lse);
      const float* kr = k + ((size_t)grp * np + j) * c;
      for (int i = 0; i < per; ++i) acc[i] = fmaf(pj, kr[lane + 32 * i], acc[i]);
    }
  }
  for (int i = 0; i < per; ++i) sh_acc[warp * 512 + lane + 32 * i] = acc[i];
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 128)
    dq[(size_t)row * c + ch] = coef * (sh_acc[ch] + sh_acc[512 + ch] + sh_acc[1024 + ch] + sh_acc[1536 + ch]);
}

static inline int grid_for(long long total, int per_block = 256) {
  long long b = (total + per_block - 1) / per_block;
  const long long cap = 8LL * device_sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace smsut

using namespace smsut;

#define DISPATCH_C(c, CALL)                                                      \
  switch (c) {                                                                   \
    case 2: { constexpr int C_ = 2; CALL; } break;                               \
    case 3: { constexpr int C_ = 3; CALL; } break;                               \
    case 4: { constexpr int C_ = 4; CALL; } break;                               \
    case 5: { constexpr int C_ = 5; CALL; } break;                               \
    case 6: { constexpr int C_ = 6; CALL; } break;                               \
    case 8: { constexpr int C_ = 8; CALL; } break;                               \
    default: SMSUT_CHECK(false, -1, "unsupported class count %d (2..6, 8)", c);  \
  }

extern "C" int smsut_dice_ce_fwd(const float* logits, const int64_t* labels, const float* label_logits, float* acc,
                                 int64_t npix, int32_t c, smsut_stream_t st) {
  SMSUT_CHECK(npix % 4 == 0 && npix > 0, -1, "pixel count must be a positive multiple of 4");
  SMSUT_CHECK((labels != nullptr) != (label_logits != nullptr), -1, "exactly one of labels / label_logits");
  const int grid = grid_for(npix / 4);
  DISPATCH_C(c, (launch_pdl(dice_ce_fwd_kernel<C_>, grid, 256, 0, (cudaStream_t)st, logits, (const long long*)labels,
                                                                           label_logits, acc, det_shadow(acc), npix)));
  count_launch();
  return launch_status("dice_ce_fwd_kernel");
}
extern "C" int smsut_dice_ce_finish(const float* acc, float* loss, int64_t npix_total, int32_t c, float w_dc,
                                    float w_ce, smsut_stream_t st) {
  launch_pdl(dice_ce_finish_kernel, 1, 32, 0, (cudaStream_t)st, acc, loss, 1.f / (float)npix_total, c, w_dc, w_ce);
  count_launch();
  return launch_status("dice_ce_finish_kernel");
}
extern "C" int smsut_dice_ce_bwd(const float* logits, const int64_t* labels, const float* label_logits,
                                 const float* acc, const float* gscale, float scale, float* dlogits, int64_t npix,
                                 int64_t npix_total, int32_t c, float w_dc, float w_ce, smsut_stream_t st) {
  SMSUT_CHECK(npix % 4 == 0 && npix > 0, -1, "pixel count must be a positive multiple of 4");
  SMSUT_CHECK((labels != nullptr) != (label_logits != nullptr), -1, "exactly one of labels / label_logits");
  const int grid = grid_for(npix / 4);
  DISPATCH_C(c, (launch_pdl(dice_ce_bwd_kernel<C_>, grid, 256, 0, (cudaStream_t)st, 
                    logits, (const long long*)labels, label_logits, acc, gscale, scale, dlogits, npix,
                    1.f / (float)npix_total, w_dc, w_ce)));
  count_launch();
  return launch_status("dice_ce_bwd_kernel");
}
extern "C" int smsut_softmax_mse_fwd(const float* zs, const float* zt, float* out, int64_t npix, int32_t c,
                                     smsut_stream_t st) {
  const int grid = grid_for(npix);
  const float scale = 1.f / ((float)npix * (float)c);
  DISPATCH_C(c, (launch_pdl(softmax_mse_kernel<C_, false>, grid, 256, 0, (cudaStream_t)st, zs, zt, out, det_shadow(out), nullptr, scale,
                                                                                 nullptr, npix)));
  count_launch();
  return launch_status("softmax_mse_fwd_kernel");
}
extern "C" int smsut_softmax_mse_bwd(const float* zs, const float* zt, const float* gscale, float* dzs, int64_t npix,
                                     int32_t c, smsut_stream_t st) {
  const int grid = grid_for(npix);
  const float scale = 1.f / ((float)npix * (float)c);
  DISPATCH_C(c, (launch_pdl(softmax_mse_kernel<C_, true>, grid, 256, 0, (cudaStream_t)st, zs, zt, nullptr, nullptr, gscale, scale, dzs,
                                                                                npix)));
  count_launch();
  return launch_status("softmax_mse_bwd_kernel");
}
extern "C" int smsut_argmax_c(const float* logits, int64_t* out, int64_t npix, int32_t c, smsut_stream_t st) {
  const int grid = grid_for(npix);
  DISPATCH_C(c, (launch_pdl(argmax_kernel<C_>, grid, 256, 0, (cudaStream_t)st, logits, (long long*)out, npix)));
  count_launch();
  return launch_status("argmax_kernel");
}
extern "C" int smsut_confusion_counts(const float* logits, const int64_t* labels, uint64_t* conf, int64_t npix, int32_t c,
                                      smsut_stream_t st) {
  SMSUT_CHECK(logits && labels && conf && npix > 0, -1, "bad confusion args");
  const int grid = grid_for(npix, 2048);
  DISPATCH_C(c, (launch_pdl(confusion_kernel<C_>, grid, 256, 0, (cudaStream_t)st, logits, (const long long*)labels,
                            (unsigned long long*)conf, npix)));
  count_launch();
  return launch_status("confusion_kernel");
}
extern "C" int smsut_l1_fwd(const float* a, const float* b, float* out, int64_t count, float scale, smsut_stream_t st) {
  launch_pdl(l1_fwd_kernel, grid_for(count, 1024), 256, 0, (cudaStream_t)st, a, b, out, det_shadow(out), count, scale);
  count_launch();
  return launch_status("l1_fwd_kernel");
}
extern "C" int smsut_l1_bwd(const float* a, const float* b, const float* gscale, float scale, float* da, int64_t count,
                            smsut_stream_t st) {
  launch_pdl(l1_bwd_kernel, grid_for(count), 256, 0, (cudaStream_t)st, a, b, gscale, scale, da, count);
  count_launch();
  return launch_status("l1_bwd_kernel");
}
extern "C" int smsut_sum_f32(const float* x, float* out, int64_t count, float scale, smsut_stream_t st) {
  launch_pdl(sum_kernel, grid_for(count, 1024), 256, 0, (cudaStream_t)st, x, out, det_shadow(out), count, scale);
  count_launch();
  return launch_status("sum_kernel");
}
extern "C" int smsut_fill_f32(float* x, int64_t count, float value, smsut_stream_t st) {
  launch_pdl(fill_kernel, grid_for(count), 256, 0, (cudaStream_t)st, x, count, value);
  count_launch();
  return launch_status("fill_kernel");
}
extern "C" int smsut_tanh_bwd(const float* dy, const float* y, float* dx, int64_t count, smsut_stream_t st) {
  launch_pdl(tanh_bwd_kernel, grid_for(count), 256, 0, (cudaStream_t)st, dy, y, dx, count);
  count_launch();
  return launch_status("tanh_bwd_kernel");
}
extern "C" int smsut_fill_scaled_f32(float* x, int64_t count, const float* gscale, float scale, smsut_stream_t st) {
  launch_pdl(fill_scaled_kernel, grid_for(count), 256, 0, (cudaStream_t)st, x, count, gscale, scale);
  count_launch();
  return launch_status("fill_scaled_kernel");
}
extern "C" int smsut_lerp_rows_f32(const float* alpha, const float* x, const float* y, float* out, int32_t rows,
                                   int64_t per, smsut_stream_t st) {
  launch_pdl(lerp_rows_kernel, grid_for((long long)rows * per), 256, 0, (cudaStream_t)st, alpha, x, y, out, per,
                                                                                 (long long)rows * per);
  count_launch();
  return launch_status("lerp_rows_kernel");
}
extern "C" int smsut_ce_rows_fwd(const float* logits, const int64_t* target, float* out, int32_t rows, int32_t c,
                                 float scale, smsut_stream_t st) {
  launch_pdl(ce_rows_fwd_kernel, 1, 256, 0, (cudaStream_t)st, logits, (const long long*)target, out, det_shadow(out), rows, c, scale);
  count_launch();
  return launch_status("ce_rows_fwd_kernel");
}
extern "C" int smsut_ce_rows_bwd(const float* logits, const int64_t* target, const float* gscale, float scale,
                                 float* dlogits, int32_t rows, int32_t c, smsut_stream_t st) {
  launch_pdl(ce_rows_bwd_kernel, grid_for(rows), 256, 0, (cudaStream_t)st, logits, (const long long*)target, gscale, scale,
                                                                   dlogits, rows, c);
  count_launch();
  return launch_status("ce_rows_bwd_kernel");
}
extern "C" int smsut_gp_fwd(const float* g, float* norm, float* norm2, float* out, int32_t b, int64_t per, float scale,
                            smsut_stream_t st) {
  SMSUT_CHECK(g && norm && norm2 && out && b > 0 && per > 0, -1, "gp_fwd: bad arguments");
  long long splits = per / 2048;
  if (splits < 1) splits = 1;
  if (splits > 32) splits = 32;
  long long* q = det_shadow(norm2);
  launch_pdl(gp_norm_kernel, dim3((unsigned)b, (unsigned)splits), 256, 0, (cudaStream_t)st, g, norm2, q, (long long)per,
             (int)splits);
  count_launch();
  if (q != nullptr) {
    const int rc = smsut_det_resolve(norm2, b, st);
    if (rc) return rc;
  }
  launch_pdl(gp_loss_kernel, 1, 32, 0, (cudaStream_t)st, (const float*)norm2, norm, out, det_shadow(out), b, scale);
  count_launch();
  return launch_status("gp_fwd kernels");
}
extern "C" int smsut_gp_bwd(const float* g, const float* norm, const float* gscale, float scale, float* u, int32_t b,
                            int64_t per, smsut_stream_t st) {
  launch_pdl(gp_bwd_kernel, grid_for((long long)b * per), 256, 0, (cudaStream_t)st, g, norm, gscale, scale, u, b, per);
  count_launch();
  return launch_status("gp_bwd_kernel");
}
extern "C" int smsut_gather_rows(const void* feat, const int64_t* ids, void* out, int32_t n, int32_t hw, int32_t c,
                                 int32_t nids, smsut_stream_t st) {
  SMSUT_CHECK(c % 8 == 0, -1, "c must be a multiple of 8");
  launch_pdl(gather_rows_kernel, grid_for((long long)n * nids * (c / 8)), 256, 0, (cudaStream_t)st, 
      (const uint4*)feat, (const long long*)ids, (uint4*)out, n, hw, c / 8, nids);
  count_launch();
  return launch_status("gather_rows_kernel");
}
extern "C" int smsut_scatter_rows_add(const void* dout, const int64_t* ids, void* dfeat, int32_t n, int32_t hw,
                                      int32_t c, int32_t nids, smsut_stream_t st) {
  SMSUT_CHECK(c % 8 == 0, -1, "c must be a multiple of 8");
  launch_pdl(scatter_rows_add_kernel, grid_for((long long)n * nids * (c / 8)), 256, 0, (cudaStream_t)st, 
      (const uint4*)dout, (const long long*)ids, (uint4*)dfeat, n, hw, c / 8, nids);
  count_launch();
  return launch_status("scatter_rows_add_kernel");
}
extern "C" int smsut_l2norm_fwd(const float* x, float* y, float* norm, int32_t rows, int32_t c, smsut_stream_t st) {
  launch_pdl(l2norm_fwd_kernel, (rows + 3) / 4, 128, 0, (cudaStream_t)st, x, y, norm, rows, c);
  count_launch();
  return launch_status("l2norm_fwd_kernel");
}
extern "C" int smsut_l2norm_bwd(const float* dy, const float* y, const float* norm, void* dx, int32_t rows, int32_t c,
                                smsut_stream_t st) {
  launch_pdl(l2norm_bwd_kernel, (rows + 3) / 4, 128, 0, (cudaStream_t)st, dy, y, norm, (__nv_bfloat16*)dx, rows, c);
  count_launch();
  return launch_status("l2norm_bwd_kernel");
}
extern "C" int smsut_patchnce_fwd(const float* q, const float* k, float* loss_rows, float* out, int32_t groups,
                                  int32_t np, int32_t c, float inv_t, float scale, smsut_stream_t st) {
  SMSUT_CHECK(c % 32 == 0 && c <= 512, -1, "PatchNCE feature dim must be a multiple of 32 and <= 512");
  const int rows = groups * np;
  launch_pdl(patchnce_kernel<false>, rows, 128, 0, (cudaStream_t)st, q, k, loss_rows, out, det_shadow(out), nullptr, scale, nullptr,
                                                                      groups, np, c, inv_t);
  count_launch();
  return launch_status("patchnce_fwd_kernel");
}
extern "C" int smsut_patchnce_bwd(const float* q, const float* k, const float* gscale, float scale, float* dq,
                                  int32_t groups, int32_t np, int32_t c, float inv_t, smsut_stream_t st) {
  SMSUT_CHECK(c % 32 == 0 && c <= 512, -1, "PatchNCE feature dim must be a multiple of 32 and <= 512");
  const int rows = groups * np;
  launch_pdl(patchnce_kernel<true>, rows, 128, 0, (cudaStream_t)st, q, k, nullptr, nullptr, nullptr, gscale, scale, dq,
                                                                     groups, np, c, inv_t);
  count_launch();
  return launch_status("patchnce_bwd_kernel");
}
