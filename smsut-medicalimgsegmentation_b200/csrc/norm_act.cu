// InstanceNorm2d(affine) + LeakyReLU + residual, forward / backward / double backward, on NHWC bf16.
// All kernels are HBM-bound streaming passes with 128-bit accesses: one thread owns 8 consecutive channels
// and strides over the pixels of one sample; per-(n,c) reductions go warp-shuffle -> shared -> one fp32
// atomic per block and channel.
//
// Reference semantics: nn.InstanceNorm2d(C, affine=True) (network/blocks.py:22-23; eps 1e-5, biased variance,
// no running stats), nn.LeakyReLU(0.01) (network/blocks.py:28-34), `x += identity` (network/blocks.py:78,115)
// and, for the double backward, torch.autograd.grad(create_graph=True) through them
// (trainer/uganShp0Trainer.py:127-134).
#include <stdlib.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

constexpr int kNT = 256;
constexpr float kEps = 1e-5f;

struct Strip {
  int cg;       // channel groups of 8
  int g;        // this thread's group
  int lane0;    // first pixel handled by this thread (relative to the strip)
  int nlanes;   // pixel stride
  long long p0, p1;  // pixel range of this block inside the sample
};

__device__ __forceinline__ Strip make_strip(int c, int hw, int splits) {
  Strip s;
  s.cg = c >> 3;
  s.g = threadIdx.x % s.cg;
  s.lane0 = threadIdx.x / s.cg;
  s.nlanes = kNT / s.cg;
  const long long per = ((long long)hw + splits - 1) / splits;
  s.p0 = (long long)blockIdx.y * per;
  s.p1 = s.p0 + per;
  if (s.p1 > hw) s.p1 = hw;
  return s;
}

__device__ __forceinline__ uint4 ldg16(const void* base, size_t elem_off) {
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off);
}
__device__ __forceinline__ void stg16(void* base, size_t elem_off, const uint4& v) {
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = v;
}

// mean / rstd of 8 channels from stats[n][2][c] = {sum, sumsq}
__device__ __forceinline__ void load_mean_rstd(const float* stats, int n, int c, int ch0, float inv_hw, float* mean,
                                               float* rstd) {
  const float* s0 = stats + (size_t)n * 2 * c + ch0;
  const float* s1 = s0 + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = s0[j] * inv_hw;
    float var = s1[j] * inv_hw - m * m;
    var = fmaxf(var, 0.f);
    mean[j] = m;
    rstd[j] = rsqrtf(var + kEps);
  }
}

// shift of the folded affine form  y = x * (gamma * rstd) + in_shift(beta, mean, gamma * rstd).  One definition for
// the forward and the backward kernels: the backward recomputes the pre-activation's sign from x instead of reading
// the activation's output, and must land on the same side of zero as the forward did.
__device__ __forceinline__ float in_shift(float beta, float mean, float scale) { return fmaf(-mean, scale, beta); }

__device__ __forceinline__ float act_grad(float out, int act, float slope) {
  if (act == SMSUT_ACT_LRELU) return out > 0.f ? 1.f : slope;
  if (act == SMSUT_ACT_RELU) return out > 0.f ? 1.f : 0.f;
  return 1.f;
}

struct Strip32 {
  int cg, g, lane0, nl;     // channel groups of 8, this thread's group, its first pixel in the strip, pixel stride
  int p0, count;            // first pixel of the block's strip, number of pixels this THREAD visits
};

__device__ __forceinline__ Strip32 make_strip32(int c, int hw, int splits, int split_idx) {
  Strip32 s;
  s.cg = c >> 3;
  s.g = threadIdx.x % s.cg;
  s.lane0 = threadIdx.x / s.cg;
  s.nl = kNT / s.cg;
  const int per = (hw + splits - 1) / splits;
  s.p0 = split_idx * per;
  int p1 = s.p0 + per;
  if (p1 > hw) p1 = hw;
  const int span = p1 - s.p0 - s.lane0;
  s.count = span > 0 ? (span + s.nl - 1) / s.nl : 0;
  return s;
}

// Sum NV x 8 per-thread values over the threads of the block that share a channel group; adds to dst[v*vstride + ch].
// sh: [kNT/32][NV][min(c, 256)] floats (c <= 256: warps of a block cover the same channel groups) or unused.
// Deterministic inside the block: per-warp slots summed in warp order (no shared-memory float atomics); the one global
// add per block and value goes through acc_add (fixed-point shadow `dst_q` in deterministic mode, else nullptr).
template <int NV, typename StripT>
__device__ __forceinline__ void block_reduce_add32(float (&acc)[NV][8], float* sh, const StripT& s, float* dst,
                                                   long long* dst_q, int c, int vstride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (s.cg < 32) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = acc[v][j];
        for (int off = 16; off >= s.cg; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        acc[v][j] = x;
      }
  }
  if (c <= 256) {
    // lanes < min(cg, 32) of every warp hold that warp's totals of channel group (warp's first group + lane)
    const int ngl = s.cg < 32 ? s.cg : 32;
    if (lane < ngl) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[(warp * NV + v) * c + s.g * 8 + j] = acc[v][j];
    }
    __syncthreads();
    // warps covering the same groups: cg <= 32 -> all 8 warps; cg == 32 (c = 256) -> each warp is one pixel, all groups
    for (int i = threadIdx.x; i < NV * c; i += kNT) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kNT / 32; ++w) t += sh[w * NV * c + i];
      const int v = i / c, ch = i - v * c;
      acc_add_at(dst, dst_q, (size_t)v * vstride + ch, t);
    }
  } else {
    // wide layers (c > 256, not on the SMSUT path): one atomic per thread and value
    if (s.cg >= 32 || lane < s.cg) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_add_at(dst, dst_q, (size_t)v * vstride + s.g * 8 + j, acc[v][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// All streaming loops below keep U independent 128-bit loads per tensor in flight per thread (the kernels are
// HBM-bound: ncu showed long-scoreboard stalls at 2 blocks/SM) and stay under 64 registers (4 blocks/SM).
constexpr int kUS = 4;  // pixels per thread per iteration, statistics / apply
constexpr int kUB = 2;  // backward kernels (3-4 tensors per pixel)

__device__ __forceinline__ uint4 zero4() { return make_uint4(0u, 0u, 0u, 0u); }

__global__ void __launch_bounds__(kNT, 4) in_stats_kernel(const void* __restrict__ x, float* __restrict__ stats,
                                                          long long* stats_q, int hw, int c, int splits) {
  pdl_prologue();
  extern __shared__ float sh[];
  const Strip s = make_strip(c, hw, splits);
  const int n = blockIdx.x;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const size_t base = (size_t)n * hw * c + s.g * 8;
  for (long long p = s.p0 + s.lane0; p < s.p1; p += (long long)kUS * s.nlanes) {
    uint4 q[kUS];
#pragma unroll
    for (int u = 0; u < kUS; ++u) {
      const long long pp = p + (long long)u * s.nlanes;
      q[u] = pp < s.p1 ? ldg16(x, base + (size_t)pp * c) : zero4();
    }
#pragma unroll
    for (int u = 0; u < kUS; ++u) {
      float v[8];
      unpack8(q[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += v[j];
        acc[1][j] = fmaf(v[j], v[j], acc[1][j]);
      }
    }
  }
  block_reduce_add32<2>(acc, sh, s, stats + (size_t)n * 2 * c, stats_q ? stats_q + (size_t)n * 2 * c : nullptr, c, c);
}

template <bool HAS_B, bool HAS_RES>
__global__ void __launch_bounds__(kNT, 4)
in_apply_kernel(const void* __restrict__ xa, const float* __restrict__ stats_a, const float* __restrict__ gamma_a,
                const float* __restrict__ beta_a, const void* __restrict__ xb, const float* __restrict__ stats_b,
                const float* __restrict__ gamma_b, const float* __restrict__ beta_b, const void* __restrict__ res,
                void* __restrict__ out, int hw, int c, int cp, int splits, int act, float slope) {
  pdl_prologue();
  const Strip s = make_strip(c, hw, splits);
  const int n = blockIdx.x;
  const int ch0 = s.g * 8;
  const float inv_hw = 1.f / (float)hw;
  float sa[8], ta[8], sb[8];
  {
    float m[8], r[8];
    load_mean_rstd(stats_a, n, c, ch0, inv_hw, m, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = ch0 + j < cp;
      const float g = ok ? gamma_a[ch0 + j] : 0.f, b = ok ? beta_a[ch0 + j] : 0.f;
      sa[j] = g * r[j];
      ta[j] = in_shift(b, m[j], sa[j]);
    }
    if (HAS_B) {
      load_mean_rstd(stats_b, n, c, ch0, inv_hw, m, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = ch0 + j < cp;
        const float g = ok ? gamma_b[ch0 + j] : 0.f, b = ok ? beta_b[ch0 + j] : 0.f;
        sb[j] = g * r[j];
        ta[j] += in_shift(b, m[j], sb[j]);      // both shifts folded into one
      }
    }
  }
  // 32-bit strip walk (see Strip32): full steps of U pixels without bounds checks, then a tail
  const Strip32 t = make_strip32(c, hw, splits, blockIdx.y);
  const uint4* pa = reinterpret_cast<const uint4*>(xa);
  const uint4* pb = reinterpret_cast<const uint4*>(xb);
  const uint4* pr = reinterpret_cast<const uint4*>(res);
  uint4* po = reinterpret_cast<uint4*>(out);
  size_t idx = ((size_t)n * hw + t.p0 + t.lane0) * t.cg + t.g;
  const int step = t.nl * t.cg;
  constexpr int U = (HAS_B || HAS_RES) ? 2 : kUS;
  auto body = [&](const uint4& qa, const uint4& qb, const uint4& qr, size_t at) {
    float v[8], o[8];
    unpack8(qa, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(v[j], sa[j], ta[j]);
    if (HAS_B) {
      unpack8(qb, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(v[j], sb[j], o[j]);
    }
    if (HAS_RES) {
      unpack8(qr, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += v[j];
    }
    if (act == SMSUT_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = lrelu(o[j], slope);
    } else if (act == SMSUT_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    po[at] = pack8(o);
  };
  int it = 0;
  for (; it + U <= t.count; it += U, idx += (size_t)U * step) {
    uint4 qa[U], qb[U], qr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qa[u] = pa[idx + (size_t)u * step];
      if (HAS_B) qb[u] = pb[idx + (size_t)u * step];
      if (HAS_RES) qr[u] = pr[idx + (size_t)u * step];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(qa[u], qb[u], qr[u], idx + (size_t)u * step);
  }
  for (; it < t.count; ++it, idx += step) {
    uint4 qb = zero4(), qr = zero4();
    const uint4 qa = pa[idx];
    if (HAS_B) qb = pb[idx];
    if (HAS_RES) qr = pr[idx];
    body(qa, qb, qr, idx);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// pass 1: red[n] = { sum g, sum g*xhat_a, sum g*xhat_b } with g = dout * act'(out).
// In the loop only sum g*(x - mean) is accumulated; rstd is applied once at the end.
//
// ncu (profiles/r1_in_bwd_ncu.md) showed the first version of these two kernels ISSUE-bound at 2.5-3.3 TB/s: 48 % of
// the issue slots busy at 30 % DRAM throughput, a third of the instructions being 64-bit index arithmetic and
// per-load bounds predicates, plus a 12 k-instruction unrolled block reduction.  This version walks the strip with
// 32-bit counters and pointer increments, keeps the bounds check out of the main loop (full steps + a tail), tests
// the activation sign on the packed bf16 bits, and reduces a block through shuffles + one shared-memory pass.
// g = dout * act'(out) for 8 packed bf16: the sign test works on the raw bits (out > 0  <=>  bits in (0, 0x8000))
template <bool HAS_ACT>
__device__ __forceinline__ void load_g(const uint4& qd, const uint4& qo, float neg, float* g) {
  unpack8(qd, g);
  if (HAS_ACT) {
    const uint32_t w[4] = {qo.x, qo.y, qo.z, qo.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lo = w[k] << 16, hi = w[k] & 0xffff0000u;
      g[2 * k] *= ((int)lo > 0) ? 1.f : neg;
      g[2 * k + 1] *= ((int)hi > 0) ? 1.f : neg;
    }
  }
}

// the same from the recomputed pre-activation  pre = xa*sa + ta [+ xb*sb]  (exactly the forward's expression)
template <bool HAS_B>
__device__ __forceinline__ void load_g_recomputed(const uint4& qd, const float* va, const float* vb, const float* sa,
                                                  const float* ta, const float* sb, float neg, float* g) {
  unpack8(qd, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float pre = fmaf(va[j], sa[j], ta[j]);
    if (HAS_B) pre = fmaf(vb[j], sb[j], pre);
    g[j] *= pre > 0.f ? 1.f : neg;
  }
}

// RECOMP: `out` is not read; the activation's sign comes from the pre-activation recomputed from xa / xb (possible
// whenever the forward had no residual input), which removes one of the 3 (4) streamed tensors.
// `n` = sample, `split` = which strip of the sample this CTA streams (the stand-alone kernel takes them from
// blockIdx.x / .y, the fused kernel below from blockIdx.y / .x)
template <bool HAS_B, bool HAS_ACT, bool RECOMP, int MINB>
__device__ __forceinline__ void
in_bwd_reduce_body(const uint4* __restrict__ dout, const uint4* __restrict__ out, const uint4* __restrict__ xa,
                   const float* __restrict__ stats_a, const float* __restrict__ gamma_a,
                   const float* __restrict__ beta_a, const uint4* __restrict__ xb,
                   const float* __restrict__ stats_b, const float* __restrict__ gamma_b,
                   const float* __restrict__ beta_b, float* __restrict__ red, long long* red_q, int hw, int c, int cp,
                   int splits, float neg, int n, int split, float* sh) {
  const Strip32 s = make_strip32(c, hw, splits, split);
  const int ch0 = s.g * 8;
  const float inv_hw = 1.f / (float)hw;
  // !RECOMP: ma / mb = the means (centred accumulation).  RECOMP: ma = sa, mb = sb, ta = the folded affine form of
  // the forward; the sums are accumulated raw and centred once at the end (sum g (x - m) = sum g x - m sum g).
  float ma[8], mb[8], ta[8];
  if (RECOMP) {
    float m[8], r[8];
    load_mean_rstd(stats_a, n, c, ch0, inv_hw, m, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = ch0 + j < cp;
      const float g = ok ? gamma_a[ch0 + j] : 0.f, b = ok ? beta_a[ch0 + j] : 0.f;
      ma[j] = g * r[j];
      ta[j] = in_shift(b, m[j], ma[j]);
    }
    if (HAS_B) {
      load_mean_rstd(stats_b, n, c, ch0, inv_hw, m, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = ch0 + j < cp;
        const float g = ok ? gamma_b[ch0 + j] : 0.f, b = ok ? beta_b[ch0 + j] : 0.f;
        mb[j] = g * r[j];
        ta[j] += in_shift(b, m[j], mb[j]);
      }
    }
  } else {
    const float* s0 = stats_a + (size_t)n * 2 * c + ch0;
#pragma unroll
    for (int j = 0; j < 8; ++j) ma[j] = s0[j] * inv_hw;
    if (HAS_B) {
      const float* t0 = stats_b + (size_t)n * 2 * c + ch0;
#pragma unroll
      for (int j = 0; j < 8; ++j) mb[j] = t0[j] * inv_hw;
    }
  }
  float acc[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
  // uint4 index of (pixel p, group g) = p * cg + g
  size_t idx = ((size_t)n * hw + s.p0 + s.lane0) * s.cg + s.g;
  const int step = s.nl * s.cg;
  constexpr int U = MINB == 2 ? 4 : 2;      // 2-resident variants have the registers for deeper load batches
  int it = 0;
  auto body = [&](const uint4& qd, const uint4& qo, const uint4& qa, const uint4& qb) {
    float g[8], v[8], w[8];
    unpack8(qa, v);
    if (HAS_B) unpack8(qb, w);
    if (RECOMP) {
      load_g_recomputed<HAS_B>(qd, v, w, ma, ta, mb, neg, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += g[j];
        acc[1][j] = fmaf(g[j], v[j], acc[1][j]);
        if (HAS_B) acc[2][j] = fmaf(g[j], w[j], acc[2][j]);
      }
    } else {
      load_g<HAS_ACT>(qd, qo, neg, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += g[j];
        acc[1][j] = fmaf(g[j], v[j] - ma[j], acc[1][j]);
        if (HAS_B) acc[2][j] = fmaf(g[j], w[j] - mb[j], acc[2][j]);
      }
    }
  };
  for (; it + U <= s.count; it += U, idx += (size_t)U * step) {
    uint4 qd[U], qo[U], qa[U], qb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qd[u] = dout[idx + (size_t)u * step];
      if (HAS_ACT && !RECOMP) qo[u] = out[idx + (size_t)u * step];
      qa[u] = xa[idx + (size_t)u * step];
      if (HAS_B) qb[u] = xb[idx + (size_t)u * step];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(qd[u], qo[u], qa[u], qb[u]);
  }
  for (; it < s.count; ++it, idx += step) {
    uint4 qo = zero4(), qb = zero4();
    const uint4 qd = dout[idx];
    if (HAS_ACT && !RECOMP) qo = out[idx];
    const uint4 qa = xa[idx];
    if (HAS_B) qb = xb[idx];
    body(qd, qo, qa, qb);
  }
  {
    float m[8], r[8];
    load_mean_rstd(stats_a, n, c, ch0, inv_hw, m, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[1][j] = (RECOMP ? fmaf(-m[j], acc[0][j], acc[1][j]) : acc[1][j]) * r[j];
    if (HAS_B) {
      load_mean_rstd(stats_b, n, c, ch0, inv_hw, m, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[2][j] = (RECOMP ? fmaf(-m[j], acc[0][j], acc[2][j]) : acc[2][j]) * r[j];
    }
  }
  block_reduce_add32<3>(acc, sh, s, red + (size_t)n * 3 * c, red_q ? red_q + (size_t)n * 3 * c : nullptr, c, c);
}

template <bool HAS_B, bool HAS_ACT, bool RECOMP, int MINB>
__global__ void __launch_bounds__(kNT, MINB)
in_bwd_reduce_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ out, const uint4* __restrict__ xa,
                     const float* __restrict__ stats_a, const float* __restrict__ gamma_a,
                     const float* __restrict__ beta_a, const uint4* __restrict__ xb,
                     const float* __restrict__ stats_b, const float* __restrict__ gamma_b,
                     const float* __restrict__ beta_b, float* __restrict__ red, long long* red_q, int hw, int c, int cp,
                     int splits, float neg) {
  pdl_prologue();
  extern __shared__ float sh[];
  in_bwd_reduce_body<HAS_B, HAS_ACT, RECOMP, MINB>(dout, out, xa, stats_a, gamma_a, beta_a, xb, stats_b, gamma_b, beta_b,
                                                   red, red_q, hw, c, cp, splits, neg, blockIdx.x, blockIdx.y, sh);
}

// pass 2: dx = A*g + B*x + C per channel with A = gamma*rstd, B = -A*rstd*mean(g xhat), C = -A*mean(g) - B*mean
// RECOMP as in the reduce kernel: A = gamma * rstd is the forward's scale, so only the folded shift `ta` is extra.
// FRESH: `red` was accumulated by other CTAs of THIS launch (fused kernel): read it past L1 (ld.global.cg), or from
// the fixed-point shadow `red_q` in deterministic mode (the fp32 copy is only written by smsut_det_resolve)
template <bool HAS_B, bool HAS_RES, bool HAS_ACT, bool RECOMP, int MINB, bool FRESH>
__device__ __forceinline__ void
in_bwd_apply_body(const uint4* __restrict__ dout, const uint4* __restrict__ out, const uint4* __restrict__ xa,
                  const float* __restrict__ stats_a, const float* __restrict__ gamma_a,
                  const float* __restrict__ beta_a, uint4* __restrict__ dxa,
                  float* __restrict__ dgamma_a, float* __restrict__ dbeta_a, const uint4* __restrict__ xb,
                  const float* __restrict__ stats_b, const float* __restrict__ gamma_b,
                  const float* __restrict__ beta_b, uint4* __restrict__ dxb,
                  float* __restrict__ dgamma_b, float* __restrict__ dbeta_b, uint4* __restrict__ dres,
                  const float* red, const long long* red_q, long long* dga_q, long long* dba_q, long long* dgb_q,
                  long long* dbb_q, int hw, int c, int cp, int splits, float neg, int n, int split) {
  const Strip32 s = make_strip32(c, hw, splits, split);
  const int ch0 = s.g * 8;
  const float inv_hw = 1.f / (float)hw;
  float Aa[8], Ba[8], Ca[8], Ab[8], Bb[8], Cb[8], ta[8];
  float r0[3 * 8];      // {sum g, sum g xhat_a, sum g xhat_b} of this thread's 8 channels
  {
    const size_t base = (size_t)n * 3 * c + ch0;
#pragma unroll
    for (int v = 0; v < (HAS_B ? 3 : 2); ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (FRESH && red_q != nullptr)
          r0[v * 8 + j] = (float)((double)__ldcg(red_q + base + (size_t)v * c + j) * (1.0 / 4294967296.0));
        else if (FRESH)
          r0[v * 8 + j] = __ldcg(red + base + (size_t)v * c + j);
        else
          r0[v * 8 + j] = red[base + (size_t)v * c + j];
      }
  }
  {
    float m[8], r[8];
    load_mean_rstd(stats_a, n, c, ch0, inv_hw, m, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = ch0 + j < cp;
      const float A = (ok ? gamma_a[ch0 + j] : 0.f) * r[j];
      const float B = -A * r[j] * (r0[8 + j] * inv_hw);
      Aa[j] = A; Ba[j] = B; Ca[j] = -A * (r0[j] * inv_hw) - B * m[j];
      if (RECOMP) ta[j] = in_shift(ok ? beta_a[ch0 + j] : 0.f, m[j], A);
    }
    if (HAS_B) {
      load_mean_rstd(stats_b, n, c, ch0, inv_hw, m, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = ch0 + j < cp;
        const float A = (ok ? gamma_b[ch0 + j] : 0.f) * r[j];
        const float B = -A * r[j] * (r0[16 + j] * inv_hw);
        Ab[j] = A; Bb[j] = B; Cb[j] = -A * (r0[j] * inv_hw) - B * m[j];
        if (RECOMP) ta[j] += in_shift(ok ? beta_b[ch0 + j] : 0.f, m[j], A);
      }
    }
  }
  // parameter gradients: one block per sample adds its (n, c) sums
  if (split == 0 && s.lane0 == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ch0 + j >= cp) continue;
      if (dgamma_a) acc_add_at(dgamma_a, dga_q, ch0 + j, r0[8 + j]);
      if (dbeta_a) acc_add_at(dbeta_a, dba_q, ch0 + j, r0[j]);
      if (HAS_B) {
        if (dgamma_b) acc_add_at(dgamma_b, dgb_q, ch0 + j, r0[16 + j]);
        if (dbeta_b) acc_add_at(dbeta_b, dbb_q, ch0 + j, r0[j]);
      }
    }
  }
  size_t idx = ((size_t)n * hw + s.p0 + s.lane0) * s.cg + s.g;
  const int step = s.nl * s.cg;
  constexpr int U = MINB == 2 ? 4 : 2;      // 2-resident variants have the registers for deeper load batches
  int it = 0;
  auto body = [&](const uint4& qd, const uint4& qo, const uint4& qa, const uint4& qb, size_t at) {
    float g[8], v[8], w[8], o[8];
    unpack8(qa, v);
    if (HAS_B) unpack8(qb, w);
    if (RECOMP) load_g_recomputed<HAS_B>(qd, v, w, Aa, ta, Ab, neg, g);
    else load_g<HAS_ACT>(qd, qo, neg, g);
    if (HAS_RES) dres[at] = pack8(g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(Aa[j], g[j], fmaf(Ba[j], v[j], Ca[j]));
    dxa[at] = pack8(o);
    if (HAS_B) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(Ab[j], g[j], fmaf(Bb[j], w[j], Cb[j]));
      dxb[at] = pack8(o);
    }
  };
  for (; it + U <= s.count; it += U, idx += (size_t)U * step) {
    uint4 qd[U], qo[U], qa[U], qb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qd[u] = dout[idx + (size_t)u * step];
      if (HAS_ACT && !RECOMP) qo[u] = out[idx + (size_t)u * step];
      qa[u] = xa[idx + (size_t)u * step];
      if (HAS_B) qb[u] = xb[idx + (size_t)u * step];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(qd[u], qo[u], qa[u], qb[u], idx + (size_t)u * step);
  }
  for (; it < s.count; ++it, idx += step) {
    uint4 qo = zero4(), qb = zero4();
    const uint4 qd = dout[idx];
    if (HAS_ACT && !RECOMP) qo = out[idx];
    const uint4 qa = xa[idx];
    if (HAS_B) qb = xb[idx];
    body(qd, qo, qa, qb, idx);
  }
}

template <bool HAS_B, bool HAS_RES, bool HAS_ACT, bool RECOMP, int MINB>
__global__ void __launch_bounds__(kNT, MINB)
in_bwd_apply_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ out, const uint4* __restrict__ xa,
                    const float* __restrict__ stats_a, const float* __restrict__ gamma_a,
                    const float* __restrict__ beta_a, uint4* __restrict__ dxa,
                    float* __restrict__ dgamma_a, float* __restrict__ dbeta_a, const uint4* __restrict__ xb,
                    const float* __restrict__ stats_b, const float* __restrict__ gamma_b,
                    const float* __restrict__ beta_b, uint4* __restrict__ dxb,
                    float* __restrict__ dgamma_b, float* __restrict__ dbeta_b, uint4* __restrict__ dres,
                    const float* __restrict__ red, long long* dga_q, long long* dba_q, long long* dgb_q,
                    long long* dbb_q, int hw, int c, int cp, int splits, float neg) {
  pdl_prologue();
  in_bwd_apply_body<HAS_B, HAS_RES, HAS_ACT, RECOMP, MINB, false>(
      dout, out, xa, stats_a, gamma_a, beta_a, dxa, dgamma_a, dbeta_a, xb, stats_b, gamma_b, beta_b, dxb, dgamma_b, dbeta_b,
      dres, red, nullptr, dga_q, dba_q, dgb_q, dbb_q, hw, c, cp, splits, neg, blockIdx.x, blockIdx.y);
}

// ---------------------------------------------------------------------------------------------
// Fused backward: ONE launch instead of the reduce / apply pair.  grid = (splits, n): the CTAs of a sample are adjacent
// in launch order.  Each CTA streams its strip once for the reductions (pass 1), the CTAs of the sample meet at a
// counter barrier, then each CTA streams the SAME strip again for dx (pass 2) -- the strip is 50-200 KB per tensor and
// was read microseconds earlier, so pass 2 is served from L2 instead of HBM.  The launcher keeps the whole grid
// co-resident (<= occupancy x SMs), which is what makes the spin barrier safe; other streams' kernels only delay it.
// Not for BatchNorm: its reductions are pooled over the samples between the passes (two-kernel path).
// ---------------------------------------------------------------------------------------------
template <bool HAS_B, bool HAS_RES, bool HAS_ACT, bool RECOMP, int MINB>
__global__ void __launch_bounds__(kNT, MINB)
in_bwd_fused_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ out, const uint4* __restrict__ xa,
                    const float* __restrict__ stats_a, const float* __restrict__ gamma_a,
                    const float* __restrict__ beta_a, uint4* __restrict__ dxa,
                    float* __restrict__ dgamma_a, float* __restrict__ dbeta_a, const uint4* __restrict__ xb,
                    const float* __restrict__ stats_b, const float* __restrict__ gamma_b,
                    const float* __restrict__ beta_b, uint4* __restrict__ dxb,
                    float* __restrict__ dgamma_b, float* __restrict__ dbeta_b, uint4* __restrict__ dres,
                    float* red, long long* red_q, unsigned int* counters, long long* dga_q, long long* dba_q,
                    long long* dgb_q, long long* dbb_q, int hw, int c, int cp, int splits, float neg) {
  pdl_prologue();
  extern __shared__ float sh[];
  const int n = blockIdx.y, split = blockIdx.x;
  in_bwd_reduce_body<HAS_B, HAS_ACT, RECOMP, MINB>(dout, out, xa, stats_a, gamma_a, beta_a, xb, stats_b, gamma_b, beta_b,
                                                   red, red_q, hw, c, cp, splits, neg, n, split, sh);
  // barrier over the `splits` CTAs of sample n: the sums are complete once every CTA has added its share
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counters + n, 1u);
    unsigned int spins = 0;
    while (atomicAdd(counters + n, 0u) < (unsigned int)splits) {
      __nanosleep(40);
      if (++spins > (1u << 26)) {
        printf("smsut: in_bwd_fused barrier timed out (sample %d split %d of %d)\n", n, split, splits);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  in_bwd_apply_body<HAS_B, HAS_RES, HAS_ACT, RECOMP, MINB, true>(
      dout, out, xa, stats_a, gamma_a, beta_a, dxa, dgamma_a, dbeta_a, xb, stats_b, gamma_b, beta_b, dxb, dgamma_b, dbeta_b,
      dres, red, red_q, dga_q, dba_q, dgb_q, dbb_q, hw, c, cp, splits, neg, n, split);
}

// ---------------------------------------------------------------------------------------------
// double backward of InstanceNorm (see header for the formulas)
//   dx = gamma*r*(dy - mean(dy) - xhat*mean(dy*xhat));  L = <u, dx>
//   g_dy = gamma*r*(u - mean(u) - xhat*mean(u*xhat))
//   g_x  = -gamma*r^2*[ xhat*(mean(u*dy) - mean(u)mean(dy) - 3*b*cu) + cu*(dy - mean(dy)) + b*(u - mean(u)) ]
//          with b = mean(dy*xhat), cu = mean(u*xhat)
//   dgamma += r*HW*(mean(u*dy) - mean(u)mean(dy) - b*cu)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT)
in_bwd2_reduce_kernel(const void* __restrict__ u, const void* __restrict__ dy, const void* __restrict__ x,
                      const float* __restrict__ stats, float* __restrict__ red2, long long* red2_q, int hw, int c,
                      int splits) {
  pdl_prologue();
  extern __shared__ float sh[];
  const Strip s = make_strip(c, hw, splits);
  const int n = blockIdx.x;
  const int ch0 = s.g * 8;
  const float inv_hw = 1.f / (float)hw;
  float m[8], r[8];
  load_mean_rstd(stats, n, c, ch0, inv_hw, m, r);
  float acc[5][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = acc[3][j] = acc[4][j] = 0.f;
  const size_t base = (size_t)n * hw * c + ch0;
  for (long long p = s.p0 + s.lane0; p < s.p1; p += s.nlanes) {
    const size_t off = base + (size_t)p * c;
    float uu[8], dd[8], xx[8];
    unpack8(ldg16(u, off), uu);
    unpack8(ldg16(dy, off), dd);
    unpack8(ldg16(x, off), xx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xx[j] - m[j]) * r[j];
      acc[0][j] += uu[j];
      acc[1][j] += dd[j];
      acc[2][j] = fmaf(uu[j], xh, acc[2][j]);
      acc[3][j] = fmaf(dd[j], xh, acc[3][j]);
      acc[4][j] = fmaf(uu[j], dd[j], acc[4][j]);
    }
  }
  block_reduce_add32<5>(acc, sh, s, red2 + (size_t)n * 5 * c, red2_q ? red2_q + (size_t)n * 5 * c : nullptr, c, c);
}

__global__ void __launch_bounds__(kNT)
in_bwd2_apply_kernel(const void* __restrict__ u, const void* __restrict__ dy, const void* __restrict__ x,
                     const float* __restrict__ stats, const float* __restrict__ gamma,
                     const float* __restrict__ red2, void* __restrict__ g_dy, void* __restrict__ g_x,
                     float* __restrict__ dgamma, long long* dgamma_q, int hw, int c, int splits) {
  pdl_prologue();
  const Strip s = make_strip(c, hw, splits);
  const int n = blockIdx.x;
  const int ch0 = s.g * 8;
  const float inv_hw = 1.f / (float)hw;
  float m[8], r[8], mu[8], md[8], cu[8], b[8], e[8], gr[8];
  load_mean_rstd(stats, n, c, ch0, inv_hw, m, r);
  const float* r0 = red2 + (size_t)n * 5 * c + ch0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = r0[j] * inv_hw;
    md[j] = r0[c + j] * inv_hw;
    cu[j] = r0[2 * c + j] * inv_hw;
    b[j] = r0[3 * c + j] * inv_hw;
    e[j] = r0[4 * c + j] * inv_hw;
    gr[j] = gamma[ch0 + j] * r[j];
  }
  if (dgamma != nullptr && blockIdx.y == 0 && s.lane0 == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      acc_add_at(dgamma, dgamma_q, ch0 + j, r[j] * (float)hw * (e[j] - mu[j] * md[j] - b[j] * cu[j]));
  }
  const size_t base = (size_t)n * hw * c + ch0;
  for (long long p = s.p0 + s.lane0; p < s.p1; p += s.nlanes) {
    const size_t off = base + (size_t)p * c;
    float uu[8], dd[8], xx[8], o1[8], o2[8];
    unpack8(ldg16(u, off), uu);
    unpack8(ldg16(dy, off), dd);
    unpack8(ldg16(x, off), xx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xx[j] - m[j]) * r[j];
      o1[j] = gr[j] * (uu[j] - mu[j] - xh * cu[j]);
      o2[j] = -gr[j] * r[j] *
              (xh * (e[j] - mu[j] * md[j] - 3.f * b[j] * cu[j]) + cu[j] * (dd[j] - md[j]) + b[j] * (uu[j] - mu[j]));
    }
    stg16(g_dy, off, pack8(o1));
    stg16(g_x, off, pack8(o2));
  }
}

// Both passes of the double backward in ONE launch (grid = (splits, n), counter barrier over the CTAs of a sample, the
// sums read back past L1 or from the fixed-point shadow): the gradient penalty's double-backward chain is 140 dependent
// kernels of which 28 were these pairs, and on that chain every link costs ~25 us inside the captured iteration.
__global__ void __launch_bounds__(kNT)
in_bwd2_fused_kernel(const void* __restrict__ u, const void* __restrict__ dy, const void* __restrict__ x,
                     const float* __restrict__ stats, const float* __restrict__ gamma, float* red2, long long* red2_q,
                     unsigned int* counters, void* __restrict__ g_dy, void* __restrict__ g_x, float* dgamma,
                     long long* dgamma_q, int hw, int c, int splits) {
  pdl_prologue();
  extern __shared__ float sh[];
  const int n = blockIdx.y, split = blockIdx.x;
  Strip s;
  s.cg = c >> 3;
  s.g = threadIdx.x % s.cg;
  s.lane0 = threadIdx.x / s.cg;
  s.nlanes = kNT / s.cg;
  {
    const long long per = ((long long)hw + splits - 1) / splits;
    s.p0 = (long long)split * per;
    s.p1 = s.p0 + per;
    if (s.p1 > hw) s.p1 = hw;
  }
  const int ch0 = s.g * 8;
  const float inv_hw = 1.f / (float)hw;
  float m[8], r[8];
  load_mean_rstd(stats, n, c, ch0, inv_hw, m, r);
  const size_t base = (size_t)n * hw * c + ch0;
  {
    float acc[5][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = acc[3][j] = acc[4][j] = 0.f;
    for (long long p = s.p0 + s.lane0; p < s.p1; p += s.nlanes) {
      const size_t off = base + (size_t)p * c;
      float uu[8], dd[8], xx[8];
      unpack8(ldg16(u, off), uu);
      unpack8(ldg16(dy, off), dd);
      unpack8(ldg16(x, off), xx);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xx[j] - m[j]) * r[j];
        acc[0][j] += uu[j];
        acc[1][j] += dd[j];
        acc[2][j] = fmaf(uu[j], xh, acc[2][j]);
        acc[3][j] = fmaf(dd[j], xh, acc[3][j]);
        acc[4][j] = fmaf(uu[j], dd[j], acc[4][j]);
      }
    }
    block_reduce_add32<5>(acc, sh, s, red2 + (size_t)n * 5 * c, red2_q ? red2_q + (size_t)n * 5 * c : nullptr, c, c);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counters + n, 1u);
    unsigned int spins = 0;
    while (atomicAdd(counters + n, 0u) < (unsigned int)splits) {
      __nanosleep(40);
      if (++spins > (1u << 26)) {
        printf("smsut: in_bwd2_fused barrier timed out (sample %d split %d of %d)\n", n, split, splits);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  float mu[8], md[8], cu[8], b[8], e[8], gr[8];
  {
    const size_t rb = (size_t)n * 5 * c + ch0;
    float t[5];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int v = 0; v < 5; ++v)
        t[v] = red2_q != nullptr ? (float)((double)__ldcg(red2_q + rb + (size_t)v * c + j) * (1.0 / 4294967296.0))
                                 : __ldcg(red2 + rb + (size_t)v * c + j);
      mu[j] = t[0] * inv_hw; md[j] = t[1] * inv_hw; cu[j] = t[2] * inv_hw; b[j] = t[3] * inv_hw; e[j] = t[4] * inv_hw;
      gr[j] = gamma[ch0 + j] * r[j];
    }
  }
  if (dgamma != nullptr && split == 0 && s.lane0 == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      acc_add_at(dgamma, dgamma_q, ch0 + j, r[j] * (float)hw * (e[j] - mu[j] * md[j] - b[j] * cu[j]));
  }
  for (long long p = s.p0 + s.lane0; p < s.p1; p += s.nlanes) {
    const size_t off = base + (size_t)p * c;
    float uu[8], dd[8], xx[8], o1[8], o2[8];
    unpack8(ldg16(u, off), uu);
    unpack8(ldg16(dy, off), dd);
    unpack8(ldg16(x, off), xx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xx[j] - m[j]) * r[j];
      o1[j] = gr[j] * (uu[j] - mu[j] - xh * cu[j]);
      o2[j] = -gr[j] * r[j] *
              (xh * (e[j] - mu[j] * md[j] - 3.f * b[j] * cu[j]) + cu[j] * (dd[j] - md[j]) + b[j] * (uu[j] - mu[j]));
    }
    stg16(g_dy, off, pack8(o1));
    stg16(g_x, off, pack8(o2));
  }
}

// ---------------------------------------------------------------------------------------------
// elementwise
// ---------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec, int act,
                               float slope) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(x[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act == SMSUT_ACT_LRELU ? lrelu(v[j], slope) : (act == SMSUT_ACT_RELU ? fmaxf(v[j], 0.f) : v[j]);
    y[i] = pack8(v);
  }
}
__global__ void act_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ ref,
                               const uint4* __restrict__ add, uint4* __restrict__ dx, long long nvec, int act,
                               float slope) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float g[8], r[8];
    unpack8(dy[i], g);
    if (act != SMSUT_ACT_NONE) {
      unpack8(ref[i], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= act_grad(r[j], act, slope);
    }
    if (add != nullptr) {
      unpack8(add[i], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += r[j];
    }
    dx[i] = pack8(g);
  }
}
__global__ void add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                           long long nvec) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(a[i], x);
    unpack8(b[i], y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    out[i] = pack8(x);
  }
}

// column sums of a (rows, c) bf16 matrix (bias gradients of the netF Linear layers)
__global__ void __launch_bounds__(kNT) colsum_kernel(const void* __restrict__ x, float* __restrict__ out,
                                                     long long* out_q, int rows, int c, int splits) {
  pdl_prologue();
  extern __shared__ float sh[];
  const Strip s = make_strip(c, rows, splits);
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  for (long long p = s.p0 + s.lane0; p < s.p1; p += s.nlanes) {
    float v[8];
    unpack8(ldg16(x, (size_t)p * c + s.g * 8), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += v[j];
  }
  block_reduce_add32<1>(acc, sh, s, out, out_q, c, c);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static int pick_splits(int n, int hw, int c, int blocks_per_sm = 8) {
  // `blocks_per_sm` blocks per SM over the whole grid = two waves of the kernel's resident blocks (8 for the 4-resident
  // forward kernels; measured with scripts/in_bench.py: 4 for the statistics kernel and for the 2-resident two-branch
  // recomputed-sign backward kernels, 93 -> 81 us per pair at 256x256x16), each block streaming >= 16 KB
  const long long bytes = (long long)hw * c * 2;
  long long s = bytes / (16 * 1024);
  static int knob = -1;
  if (knob < 0) {
    const char* e = getenv("SMSUT_IN_BPS");      // development knob: overrides every kernel's choice
    knob = e && atoi(e) > 0 ? atoi(e) : 0;
  }
  const int bps = knob > 0 ? knob : blocks_per_sm;
  const long long want = ((long long)bps * device_sm_count() + n - 1) / n;
  if (s > want) s = want;
  if (s < 1) s = 1;
  if (s > hw) s = hw;
  return (int)s;
}

// shared memory of block_reduce_add32: per-warp slots for c <= 256
static size_t red_smem(int nv, int c) { return c <= 256 ? (size_t)(kNT / 32) * nv * c * sizeof(float) : 0; }

static int check_nc(int n, int hw, int c) {
  SMSUT_CHECK(n > 0 && hw > 0 && c >= 8 && (c & 7) == 0 && kNT % (c >> 3) == 0, -1,
              "instance-norm kernels need C in {8,16,...,2048} dividing 2048 (got n=%d hw=%d c=%d)", n, hw, c);
  return 0;
}

static inline int grid_for(long long nvec) {
  long long b = (nvec + 255) / 256;
  const long long cap = 8LL * device_sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace smsut

using namespace smsut;

extern "C" int smsut_in_stats(const void* x, int32_t n, int32_t hw, int32_t c, float* stats, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  const int splits = pick_splits(n, hw, c, 4);
  launch_pdl(in_stats_kernel, dim3(n, splits), kNT, red_smem(2, c), (cudaStream_t)st, x, stats, det_shadow(stats), hw, c,
             splits);
  count_launch();
  return launch_status("in_stats_kernel");
}

extern "C" int smsut_in_apply(const void* xa, const float* stats_a, const float* gamma_a, const float* beta_a,
                              const void* xb, const float* stats_b, const float* gamma_b, const float* beta_b,
                              const void* res, void* out, int32_t n, int32_t hw, int32_t c, int32_t cp, int32_t act,
                              float slope, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  const int splits = pick_splits(n, hw, c);
#define IN_APPLY(HB, HR)                                                                                          \
  launch_pdl(in_apply_kernel<HB, HR>, dim3(n, splits), kNT, 0, (cudaStream_t)st, xa, stats_a, gamma_a, beta_a, xb, stats_b, \
                                                                         gamma_b, beta_b, res, out, hw, c, cp, splits, act, slope)
  if (xb != nullptr) { if (res != nullptr) IN_APPLY(true, true); else IN_APPLY(true, false); }
  else { if (res != nullptr) IN_APPLY(false, true); else IN_APPLY(false, false); }
#undef IN_APPLY
  count_launch();
  return launch_status("in_apply_kernel");
}

// two-branch recomputed-sign kernels: 2 blocks / SM without register spills, or 3 with ~250 B of spills
// (SMSUT_IN_RECOMP_MINB=2|3; measured on B200: see DESIGN.md)
static bool recomp_two_blocks() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SMSUT_IN_RECOMP_MINB");
    v = e ? atoi(e) : 2;
  }
  return v == 2;
}

extern "C" int smsut_in_bwd_reduce(const void* dout, const void* out, const void* xa, const float* stats_a,
                                   const float* gamma_a, const float* beta_a, const void* xb, const float* stats_b,
                                   const float* gamma_b, const float* beta_b, float* red, int32_t n, int32_t hw,
                                   int32_t c, int32_t cp, int32_t act, float slope, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  SMSUT_CHECK(act == SMSUT_ACT_NONE || act == SMSUT_ACT_LRELU || act == SMSUT_ACT_RELU, -1, "in_bwd: unsupported activation");
  // out == NULL with an activation: the sign is recomputed from xa / xb, which needs the forward's affine parameters
  const bool recomp = act != SMSUT_ACT_NONE && out == nullptr;
  const int splits = pick_splits(n, hw, c, recomp && xb != nullptr && recomp_two_blocks() ? 4 : 8);
  SMSUT_CHECK(!recomp || (gamma_a && beta_a && (xb == nullptr || (gamma_b && beta_b))), -1,
              "in_bwd_reduce: out == NULL needs gamma / beta of every branch");
  const float neg = act == SMSUT_ACT_LRELU ? slope : 0.f;
  const size_t shm = red_smem(3, c);
  long long* red_q = det_shadow(red);
#define IN_BWD_RED(HB, HA, RC, MB)                                                                                 \
  launch_pdl(in_bwd_reduce_kernel<HB, HA, RC, MB>, dim3(n, splits), kNT, shm, (cudaStream_t)st, (const uint4*)dout, \
             (const uint4*)out, (const uint4*)xa, stats_a, gamma_a, beta_a, (const uint4*)xb, stats_b, gamma_b,    \
             beta_b, red, red_q, hw, c, cp, splits, neg)
  if (xb != nullptr) {
    if (recomp) { if (recomp_two_blocks()) IN_BWD_RED(true, true, true, 2); else IN_BWD_RED(true, true, true, 3); }
    else if (act != SMSUT_ACT_NONE) IN_BWD_RED(true, true, false, 3);
    else IN_BWD_RED(true, false, false, 3);
  } else {
    if (recomp) IN_BWD_RED(false, true, true, 3);
    else if (act != SMSUT_ACT_NONE) IN_BWD_RED(false, true, false, 3);
    else IN_BWD_RED(false, false, false, 3);
  }
#undef IN_BWD_RED
  count_launch();
  return launch_status("in_bwd_reduce_kernel");
}

extern "C" int smsut_in_bwd_apply(const void* dout, const void* out, const void* xa, const float* stats_a,
                                  const float* gamma_a, const float* beta_a, void* dxa, float* dgamma_a,
                                  float* dbeta_a, const void* xb, const float* stats_b, const float* gamma_b,
                                  const float* beta_b, void* dxb, float* dgamma_b, float* dbeta_b, void* dres,
                                  const float* red, int32_t n, int32_t hw, int32_t c, int32_t cp, int32_t act,
                                  float slope, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  SMSUT_CHECK(act == SMSUT_ACT_NONE || act == SMSUT_ACT_LRELU || act == SMSUT_ACT_RELU, -1, "in_bwd: unsupported activation");
  const bool recomp = act != SMSUT_ACT_NONE && out == nullptr;
  const int splits = pick_splits(n, hw, c, recomp && xb != nullptr && recomp_two_blocks() ? 4 : 8);
  SMSUT_CHECK(!recomp || (beta_a && (xb == nullptr || beta_b)), -1, "in_bwd_apply: out == NULL needs beta of every branch");
  const float neg = act == SMSUT_ACT_LRELU ? slope : 0.f;
#define IN_BWD_APPLY3(HB, HR, HA, RC)                                                                              \
  do { if (HB && RC && recomp_two_blocks()) IN_BWD_APPLY4(HB, HR, HA, RC, 2); else IN_BWD_APPLY4(HB, HR, HA, RC, 3); } while (0)
#define IN_BWD_APPLY4(HB, HR, HA, RC, MB)                                                                          \
  launch_pdl(in_bwd_apply_kernel<HB, HR, HA, RC, MB>, dim3(n, splits), kNT, 0, (cudaStream_t)st, (const uint4*)dout, \
             (const uint4*)out, (const uint4*)xa, stats_a, gamma_a, beta_a, (uint4*)dxa, dgamma_a, dbeta_a,        \
             (const uint4*)xb, stats_b, gamma_b, beta_b, (uint4*)dxb, dgamma_b, dbeta_b, (uint4*)dres, red,        \
             det_shadow(dgamma_a), det_shadow(dbeta_a), det_shadow(dgamma_b), det_shadow(dbeta_b), hw, c, cp, splits, neg)
#define IN_BWD_APPLY(HB, HR)                                        \
  do {                                                              \
    if (recomp) IN_BWD_APPLY3(HB, HR, true, true);                  \
    else if (act != SMSUT_ACT_NONE) IN_BWD_APPLY3(HB, HR, true, false); \
    else IN_BWD_APPLY3(HB, HR, false, false);                       \
  } while (0)
  if (xb != nullptr) { if (dres != nullptr) IN_BWD_APPLY(true, true); else IN_BWD_APPLY(true, false); }
  else { if (dres != nullptr) IN_BWD_APPLY(false, true); else IN_BWD_APPLY(false, false); }
#undef IN_BWD_APPLY4
#undef IN_BWD_APPLY3
#undef IN_BWD_APPLY
  count_launch();
  return launch_status("in_bwd_apply_kernel");
}

// InstanceNorm backward in ONE call: the fused kernel when the whole grid can be co-resident (see
// in_bwd_fused_kernel), else the reduce / apply pair.  `counters`: n zeroed 32-bit words (the per-sample barrier).
extern "C" int smsut_in_bwd_fused(const void* dout, const void* out, const void* xa, const float* stats_a,
                                  const float* gamma_a, const float* beta_a, void* dxa, float* dgamma_a,
                                  float* dbeta_a, const void* xb, const float* stats_b, const float* gamma_b,
                                  const float* beta_b, void* dxb, float* dgamma_b, float* dbeta_b, void* dres,
                                  float* red, void* counters, int32_t n, int32_t hw, int32_t c, int32_t cp, int32_t act,
                                  float slope, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  SMSUT_CHECK(act == SMSUT_ACT_NONE || act == SMSUT_ACT_LRELU || act == SMSUT_ACT_RELU, -1, "in_bwd: unsupported activation");
  SMSUT_CHECK(red != nullptr && counters != nullptr, -1, "in_bwd_fused: red / counters missing");
  const bool recomp = act != SMSUT_ACT_NONE && out == nullptr;
  SMSUT_CHECK(!recomp || (gamma_a && beta_a && (xb == nullptr || (gamma_b && beta_b))), -1,
              "in_bwd_fused: out == NULL needs gamma / beta of every branch");
  static int knob = -1;
  if (knob < 0) {
    const char* e = getenv("SMSUT_IN_FUSED");
    knob = e ? atoi(e) : 1;
  }
  const float neg = act == SMSUT_ACT_LRELU ? slope : 0.f;
  const size_t shm = red_smem(3, c);
  long long* red_q = det_shadow(red);
  const bool two = recomp && xb != nullptr && recomp_two_blocks();      // the 2-resident register budget (see above)
  int launched = 0;
#define IN_FUSED(HB, HR, HA, RC, MB)                                                                                  \
  do {                                                                                                                \
    static int per_sm = -1;                                                                                           \
    if (per_sm < 0) {                                                                                                 \
      int nb = 0;                                                                                                     \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, in_bwd_fused_kernel<HB, HR, HA, RC, MB>, kNT, 24 * 1024) != \
          cudaSuccess)                                                                                                \
        nb = 0;                                                                                                       \
      per_sm = nb;                                                                                                    \
    }                                                                                                                 \
    const long long cap = (long long)per_sm * device_sm_count();                                                      \
    const long long sample_bytes = (long long)hw * c * 2;                                                            \
    const bool pick = knob == 1 || (knob == 2 && sample_bytes <= 256 * 1024) || (knob == 3 && sample_bytes > 256 * 1024); \
    if (pick && shm <= 24 * 1024 && cap >= n) {                                                                       \
      int splits = pick_splits(n, hw, c, MB == 2 ? 4 : 8);                                                            \
      if ((long long)splits * n > cap) splits = (int)(cap / n);                                                       \
      launch_pdl(in_bwd_fused_kernel<HB, HR, HA, RC, MB>, dim3(splits, n), kNT, shm, (cudaStream_t)st,                \
                 (const uint4*)dout, (const uint4*)out, (const uint4*)xa, stats_a, gamma_a, beta_a, (uint4*)dxa,      \
                 dgamma_a, dbeta_a, (const uint4*)xb, stats_b, gamma_b, beta_b, (uint4*)dxb, dgamma_b, dbeta_b,       \
                 (uint4*)dres, red, red_q, (unsigned int*)counters, det_shadow(dgamma_a), det_shadow(dbeta_a),        \
                 det_shadow(dgamma_b), det_shadow(dbeta_b), hw, c, cp, splits, neg);                                  \
      launched = 1;                                                                                                   \
    }                                                                                                                 \
  } while (0)
#define IN_FUSED_A(HB, HR)                                                                  \
  do {                                                                                      \
    if (recomp) { if (two) IN_FUSED(HB, HR, true, true, 2); else IN_FUSED(HB, HR, true, true, 3); } \
    else if (act != SMSUT_ACT_NONE) IN_FUSED(HB, HR, true, false, 3);                       \
    else IN_FUSED(HB, HR, false, false, 3);                                                 \
  } while (0)
  if (xb != nullptr) { if (dres != nullptr) IN_FUSED_A(true, true); else IN_FUSED_A(true, false); }
  else { if (dres != nullptr) IN_FUSED_A(false, true); else IN_FUSED_A(false, false); }
#undef IN_FUSED_A
#undef IN_FUSED
  if (launched) {
    count_launch();
    return launch_status("in_bwd_fused_kernel");
  }
  // not co-residable (more samples than resident blocks) or switched off: the two-kernel path
  rc = smsut_in_bwd_reduce(dout, out, xa, stats_a, gamma_a, beta_a, xb, stats_b, gamma_b, beta_b, red, n, hw, c, cp, act,
                           slope, st);
  if (rc) return rc;
  if (red_q != nullptr) {
    rc = smsut_det_resolve(red, (int64_t)n * 3 * c, st);
    if (rc) return rc;
  }
  return smsut_in_bwd_apply(dout, out, xa, stats_a, gamma_a, beta_a, dxa, dgamma_a, dbeta_a, xb, stats_b, gamma_b, beta_b,
                            dxb, dgamma_b, dbeta_b, dres, red, n, hw, c, cp, act, slope, st);
}

extern "C" int smsut_in_bwd2_reduce(const void* u, const void* dy, const void* x, const float* stats, float* red2,
                                    int32_t n, int32_t hw, int32_t c, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  const int splits = pick_splits(n, hw, c);
  launch_pdl(in_bwd2_reduce_kernel, dim3(n, splits), kNT, red_smem(5, c), (cudaStream_t)st, u, dy, x, stats, red2,
             det_shadow(red2), hw, c, splits);
  count_launch();
  return launch_status("in_bwd2_reduce_kernel");
}

extern "C" int smsut_in_bwd2_apply(const void* u, const void* dy, const void* x, const float* stats,
                                   const float* gamma, const float* red2, void* g_dy, void* g_x, float* dgamma,
                                   int32_t n, int32_t hw, int32_t c, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  const int splits = pick_splits(n, hw, c);
  launch_pdl(in_bwd2_apply_kernel, dim3(n, splits), kNT, 0, (cudaStream_t)st, u, dy, x, stats, gamma, red2, g_dy, g_x, dgamma,
             det_shadow(dgamma), hw, c, splits);
  count_launch();
  return launch_status("in_bwd2_apply_kernel");
}

// double backward in one call: fused kernel when the grid can be co-resident, else the pair.  counters: n zeroed words
extern "C" int smsut_in_bwd2_fused(const void* u, const void* dy, const void* x, const float* stats, const float* gamma,
                                   float* red2, void* counters, void* g_dy, void* g_x, float* dgamma, int32_t n, int32_t hw,
                                   int32_t c, smsut_stream_t st) {
  int rc = check_nc(n, hw, c);
  if (rc) return rc;
  SMSUT_CHECK(red2 != nullptr && counters != nullptr, -1, "in_bwd2_fused: red2 / counters missing");
  static int knob = -1, per_sm = -1;
  if (knob < 0) {
    const char* e = getenv("SMSUT_IN_FUSED");
    knob = e ? atoi(e) : 1;
  }
  const size_t shm = red_smem(5, c);
  if (per_sm < 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, in_bwd2_fused_kernel, kNT, 40 * 1024) != cudaSuccess) nb = 0;
    per_sm = nb;
  }
  const long long cap = (long long)per_sm * device_sm_count();
  long long* red2_q = det_shadow(red2);
  const long long sample_bytes = (long long)hw * c * 2;
  const bool pick = knob == 1 || (knob == 2 && sample_bytes <= 256 * 1024) || (knob == 3 && sample_bytes > 256 * 1024);
  if (pick && shm <= 40 * 1024 && cap >= n) {
    int splits = pick_splits(n, hw, c);
    if ((long long)splits * n > cap) splits = (int)(cap / n);
    launch_pdl(in_bwd2_fused_kernel, dim3(splits, n), kNT, shm, (cudaStream_t)st, u, dy, x, stats, gamma, red2, red2_q,
               (unsigned int*)counters, g_dy, g_x, dgamma, det_shadow(dgamma), hw, c, splits);
    count_launch();
    return launch_status("in_bwd2_fused_kernel");
  }
  rc = smsut_in_bwd2_reduce(u, dy, x, stats, red2, n, hw, c, st);
  if (rc) return rc;
  if (red2_q != nullptr) {
    rc = smsut_det_resolve(red2, (int64_t)n * 5 * c, st);
    if (rc) return rc;
  }
  return smsut_in_bwd2_apply(u, dy, x, stats, gamma, red2, g_dy, g_x, dgamma, n, hw, c, st);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm2d on top of the InstanceNorm kernels (network/blocks.py:19-26 get_norm('batch'), the default of
// network/unet.py:14): batch statistics are the per-sample sums averaged over the samples, so the same apply /
// backward kernels run on a (n, k, c) table whose n rows all hold the batch mean of the per-sample rows.
// ---------------------------------------------------------------------------------------------
__global__ void bn_pool_kernel(const float* rows, float* out, int n, int kc) {  // out may alias rows (column-private)
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kc) return;
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += rows[(size_t)j * kc + i];
  s /= (float)n;
  for (int j = 0; j < n; ++j) out[(size_t)j * kc + i] = s;
}

// running_mean / running_var momentum update from pooled statistics (row 0 of the table): PyTorch semantics,
// unbiased variance in the running estimate (count = n*hw).
__global__ void bn_running_kernel(const float* __restrict__ pooled, int c, int cp, float inv_hw, float unbias,
                                  float momentum, float* __restrict__ rmean, float* __restrict__ rvar) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cp) return;
  const float m = pooled[i] * inv_hw;
  const float var = fmaxf(pooled[c + i] * inv_hw - m * m, 0.f);
  rmean[i] = (1.f - momentum) * rmean[i] + momentum * m;
  rvar[i] = (1.f - momentum) * rvar[i] + momentum * var * unbias;
}

// eval mode: a statistics table that makes the apply kernel normalise with the running estimates
__global__ void bn_eval_stats_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar,
                                     float* __restrict__ stats, int n, int c, int cp, float hw) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int j = i / c, ch = i - j * c;
  const float m = ch < cp ? rmean[ch] : 0.f, v = ch < cp ? rvar[ch] : 0.f;
  stats[(size_t)j * 2 * c + ch] = m * hw;
  stats[(size_t)j * 2 * c + c + ch] = (v + m * m) * hw;
}

extern "C" int smsut_bn_pool(const float* rows, float* out, int32_t n, int32_t k, int32_t c, smsut_stream_t st) {
  SMSUT_CHECK(n > 0 && k > 0 && c > 0, -1, "bn_pool: empty table");
  const int kc = k * c;
  launch_pdl(bn_pool_kernel, dim3((kc + 127) / 128), 128, 0, (cudaStream_t)st, rows, out, n, kc);
  count_launch();
  return launch_status("bn_pool_kernel");
}

extern "C" int smsut_bn_running_update(const float* pooled, int32_t n, int32_t hw, int32_t c, int32_t cp,
                                       float momentum, float* running_mean, float* running_var, smsut_stream_t st) {
  SMSUT_CHECK(n > 0 && hw > 0 && cp > 0 && cp <= c, -1, "bn_running_update: bad sizes");
  const double count = (double)n * hw;
  const float unbias = count > 1.0 ? (float)(count / (count - 1.0)) : 1.f;
  launch_pdl(bn_running_kernel, dim3((cp + 127) / 128), 128, 0, (cudaStream_t)st, pooled, c, cp, 1.f / (float)hw, unbias,
             momentum, running_mean, running_var);
  count_launch();
  return launch_status("bn_running_kernel");
}

extern "C" int smsut_bn_eval_stats(const float* running_mean, const float* running_var, float* stats, int32_t n,
                                   int32_t hw, int32_t c, int32_t cp, smsut_stream_t st) {
  SMSUT_CHECK(n > 0 && hw > 0 && cp > 0 && cp <= c, -1, "bn_eval_stats: bad sizes");
  launch_pdl(bn_eval_stats_kernel, dim3((n * c + 127) / 128), 128, 0, (cudaStream_t)st, running_mean, running_var, stats, n,
             c, cp, (float)hw);
  count_launch();
  return launch_status("bn_eval_stats_kernel");
}

extern "C" int smsut_colsum_bf16(const void* x, int32_t rows, int32_t c, float* out, smsut_stream_t st) {
  int rc = check_nc(1, rows, c);
  if (rc) return rc;
  const int splits = pick_splits(1, rows, c);
  launch_pdl(colsum_kernel, dim3(1, splits), kNT, red_smem(1, c), (cudaStream_t)st, x, out, det_shadow(out), rows, c, splits);
  count_launch();
  return launch_status("colsum_kernel");
}

extern "C" int smsut_act_fwd(const void* x, void* y, int64_t count, int32_t act, float slope, smsut_stream_t st) {
  SMSUT_CHECK(count % 8 == 0, -1, "element count must be a multiple of 8");
  launch_pdl(act_fwd_kernel, grid_for(count / 8), 256, 0, (cudaStream_t)st, (const uint4*)x, (uint4*)y, count / 8, act, slope);
  count_launch();
  return launch_status("act_fwd_kernel");
}
extern "C" int smsut_act_bwd(const void* dy, const void* ref, const void* add, void* dx, int64_t count, int32_t act,
                             float slope, smsut_stream_t st) {
  SMSUT_CHECK(count % 8 == 0, -1, "element count must be a multiple of 8");
  launch_pdl(act_bwd_kernel, grid_for(count / 8), 256, 0, (cudaStream_t)st, (const uint4*)dy, (const uint4*)ref,
                                                                    (const uint4*)add, (uint4*)dx, count / 8, act, slope);
  count_launch();
  return launch_status("act_bwd_kernel");
}
extern "C" int smsut_add_bf16(const void* a, const void* b, void* out, int64_t count, smsut_stream_t st) {
  SMSUT_CHECK(count % 8 == 0, -1, "element count must be a multiple of 8");
  launch_pdl(add_kernel, grid_for(count / 8), 256, 0, (cudaStream_t)st, (const uint4*)a, (const uint4*)b, (uint4*)out,
                                                                count / 8);
  count_launch();
  return launch_status("add_kernel");
}
