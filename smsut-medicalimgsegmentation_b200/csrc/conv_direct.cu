// Direct (CUDA-core) NHWC convolutions for the tiny-K, HBM-bound layers and as the generic cross-check
// path: the 5x5 stems (network/ugan.py:26), the discriminator's 4x4 s2 stem with bias (network/ugan.py:202),
// the 1x1 heads with bias/tanh (network/ugan.py:70-83), conv_src / conv_cls (network/ugan.py:213-215).
// Weights are the fp32 OIHW masters; accumulation is fp32.
#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

struct DirectParams {
  int n, h, w, cin, cout, kh, kw, stride, pad, ho, wo;
  const void* x; int x_ld, x_f32;
  const float* wt; const float* bias;
  void* y; int y_ld, y_f32;
  int act; float slope; int accumulate;
  int cch;  // channel chunk staged in smem
};

__device__ __forceinline__ float load_act(const void* p, size_t idx, int f32) {
  return f32 ? reinterpret_cast<const float*>(p)[idx] : bf2f(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}
__device__ __forceinline__ void store_act(void* p, size_t idx, int f32, float v, int accumulate) {
  if (f32) {
    float* q = reinterpret_cast<float*>(p) + idx;
    *q = accumulate ? *q + v : v;
  } else {
    __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(p) + idx;
    *q = f2bf(accumulate ? bf2f(*q) + v : v);
  }
}
__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == SMSUT_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SMSUT_ACT_LRELU) return lrelu(v, slope);
  if (act == 3) return tanhf(v);
  return v;
}

// ---------------------------------------------------------------------------------------------
// fprop: one thread = one output pixel x CT output channels
// ---------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(128) direct_fprop_kernel(const DirectParams p) {
  pdl_prologue();
  extern __shared__ float ws[];  // [taps][cch][CT]
  const int taps = p.kh * p.kw;
  const int co0 = blockIdx.y * CT;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npix = (long long)p.n * p.ho * p.wo;
  const bool active = pix < npix;
  int wo = 0, ho = 0, n = 0;
  if (active) {
    long long t = pix;
    wo = (int)(t % p.wo); t /= p.wo;
    ho = (int)(t % p.ho); n = (int)(t / p.ho);
  }
  float acc[CT];
#pragma unroll
  for (int o = 0; o < CT; ++o) acc[o] = 0.f;

  for (int c0 = 0; c0 < p.cin; c0 += p.cch) {
    const int cn = min(p.cch, p.cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < taps * cn * CT; i += blockDim.x) {
      const int o = i % CT;
      const int c = (i / CT) % cn;
      const int t = i / (CT * cn);
      const int co = co0 + o;
      ws[i] = co < p.cout ? p.wt[((size_t)co * p.cin + (c0 + c)) * taps + t] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int iy = ho * p.stride - p.pad + ky;
      if (iy < 0 || iy >= p.h) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int ix = wo * p.stride - p.pad + kx;
        if (ix < 0 || ix >= p.w) continue;
        const size_t base = (((size_t)n * p.h + iy) * p.w + ix) * p.x_ld + c0;
        const float* wrow = ws + (size_t)(ky * p.kw + kx) * cn * CT;
        if (!p.x_f32 && (cn & 7) == 0 && (p.x_ld & 7) == 0 && (c0 & 7) == 0) {
          const uint4* xp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + base);
          for (int c8 = 0; c8 < cn; c8 += 8) {
            float xv[8];
            unpack8(xp[c8 >> 3], xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float* wv = wrow + (size_t)(c8 + j) * CT;
#pragma unroll
              for (int o = 0; o < CT; ++o) acc[o] = fmaf(xv[j], wv[o], acc[o]);
            }
          }
        } else {
          for (int c = 0; c < cn; ++c) {
            const float xv = load_act(p.x, base + c, p.x_f32);
            const float* wv = wrow + (size_t)c * CT;
#pragma unroll
            for (int o = 0; o < CT; ++o) acc[o] = fmaf(xv, wv[o], acc[o]);
          }
        }
      }
    }
  }
  if (!active) return;
  const size_t obase = (size_t)pix * p.y_ld;
  if (CT >= 8 && !p.y_f32 && !p.accumulate && co0 + CT <= p.y_ld && ((obase + co0) & 7) == 0) {
    // bf16 output, whole CT-channel run inside the pixel: 128-bit stores instead of CT two-byte stores
    float v[CT];
#pragma unroll
    for (int o = 0; o < CT; ++o) {
      const int co = co0 + o;
      float t = acc[o];
      if (co < p.cout) {
        if (p.bias) t += p.bias[co];
        t = apply_act(t, p.act, p.slope);
      } else {
        t = 0.f;
      }
      v[o] = t;
    }
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + obase + co0);
#pragma unroll
    for (int o = 0; o < CT; o += 8) dst[o >> 3] = pack8(v + o);
    return;
  }
#pragma unroll
  for (int o = 0; o < CT; ++o) {
    const int co = co0 + o;
    if (co >= p.y_ld) break;
    float v = acc[o];
    if (co < p.cout) {
      if (p.bias) v += p.bias[co];
      v = apply_act(v, p.act, p.slope);
    } else {
      v = 0.f;  // channel padding
    }
    store_act(p.y, obase + co, p.y_f32, v, p.accumulate);
  }
}

// fprop for very few output pixels (conv_cls: 16 outputs of K = 4096; conv_src: 256 outputs of K = 2304):
// one warp per (output pixel, output channel), lanes split the input channels (coalesced NHWC reads).
__global__ void __launch_bounds__(256) direct_fprop_small_kernel(const DirectParams p) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nout = (long long)p.n * p.ho * p.wo * p.cout;
  if (wid >= nout) return;
  const int co = (int)(wid % p.cout);
  long long t = wid / p.cout;
  const int wo = (int)(t % p.wo); t /= p.wo;
  const int ho = (int)(t % p.ho);
  const int n = (int)(t / p.ho);
  const int taps = p.kh * p.kw;
  float acc = 0.f;
  for (int ky = 0; ky < p.kh; ++ky) {
    const int iy = ho * p.stride - p.pad + ky;
    if (iy < 0 || iy >= p.h) continue;
    for (int kx = 0; kx < p.kw; ++kx) {
      const int ix = wo * p.stride - p.pad + kx;
      if (ix < 0 || ix >= p.w) continue;
      const size_t base = (((size_t)n * p.h + iy) * p.w + ix) * p.x_ld;
      const float* wr = p.wt + (size_t)co * p.cin * taps + ky * p.kw + kx;
      for (int c = lane; c < p.cin; c += 32) acc = fmaf(load_act(p.x, base + c, p.x_f32), wr[(size_t)c * taps], acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (p.bias) acc += p.bias[co];
    acc = apply_act(acc, p.act, p.slope);
    store_act(p.y, (size_t)(wid / p.cout) * p.y_ld + co, p.y_f32, acc, p.accumulate);
  }
}

// ---------------------------------------------------------------------------------------------
// dgrad: one thread = one input pixel x CT input channels (gather form)
//   dx[n,iy,ix,ci] = sum_{co,ky,kx : oy*stride - pad + ky == iy} dy[n,oy,ox,co] * w[co,ci,ky,kx]
// here p.x / x_ld / x_f32 describe dx (written) and p.y / y_ld / y_f32 describe dy (read).
// ---------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(128) direct_dgrad_kernel(const DirectParams p) {
  pdl_prologue();
  extern __shared__ float ws[];  // [taps][cch(co)][CT(ci)]
  const int taps = p.kh * p.kw;
  const int ci0 = blockIdx.y * CT;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npix = (long long)p.n * p.h * p.w;
  const bool active = pix < npix;
  int ix = 0, iy = 0, n = 0;
  if (active) {
    long long t = pix;
    ix = (int)(t % p.w); t /= p.w;
    iy = (int)(t % p.h); n = (int)(t / p.h);
  }
  float acc[CT];
#pragma unroll
  for (int o = 0; o < CT; ++o) acc[o] = 0.f;

  for (int c0 = 0; c0 < p.cout; c0 += p.cch) {
    const int cn = min(p.cch, p.cout - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < taps * cn * CT; i += blockDim.x) {
      const int o = i % CT;
      const int c = (i / CT) % cn;
      const int t = i / (CT * cn);
      const int ci = ci0 + o;
      ws[i] = ci < p.cin ? p.wt[((size_t)(c0 + c) * p.cin + ci) * taps + t] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int ty = iy + p.pad - ky;
      if (ty < 0 || ty % p.stride != 0) continue;
      const int oy = ty / p.stride;
      if (oy >= p.ho) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int tx = ix + p.pad - kx;
        if (tx < 0 || tx % p.stride != 0) continue;
        const int ox = tx / p.stride;
        if (ox >= p.wo) continue;
        const size_t base = (((size_t)n * p.ho + oy) * p.wo + ox) * p.y_ld + c0;
        const float* wrow = ws + (size_t)(ky * p.kw + kx) * cn * CT;
        int c = 0;
        if (!p.y_f32 && ((base | (size_t)p.y_ld) & 7) == 0) {
          // bf16 gradient with 16-byte aligned channel runs: 8 channels per 128-bit load (the scalar loop below issued
          // 64 two-byte loads per pixel on D's stem: 85 us for 67 MFLOP)
          for (; c + 8 <= cn; c += 8) {
            const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.y) + base + c);
            float gv[8];
            unpack8(q, gv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float* wv = wrow + (size_t)(c + k) * CT;
#pragma unroll
              for (int o = 0; o < CT; ++o) acc[o] = fmaf(gv[k], wv[o], acc[o]);
            }
          }
        }
        for (; c < cn; ++c) {
          const float gv = load_act(p.y, base + c, p.y_f32);
          const float* wv = wrow + (size_t)c * CT;
#pragma unroll
          for (int o = 0; o < CT; ++o) acc[o] = fmaf(gv, wv[o], acc[o]);
        }
      }
    }
  }
  if (!active) return;
  const size_t obase = (size_t)pix * p.x_ld;
#pragma unroll
  for (int o = 0; o < CT; ++o) {
    const int ci = ci0 + o;
    if (ci >= p.x_ld) break;
    store_act(const_cast<void*>(p.x), obase + ci, p.x_f32, ci < p.cin ? acc[o] : 0.f, p.accumulate);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: block = a strip of output rows; thread owns up to OPT (co,ci,tap) outputs (grid.y chunks)
// ---------------------------------------------------------------------------------------------
constexpr int kWgOPT = 4;
__global__ void __launch_bounds__(256) direct_wgrad_kernel(const DirectParams p, float* dw, float* dbias,
                                                           long long* dw_q, long long* dbias_q,
                                                           int rows_per_block) {
  pdl_prologue();
  const int taps = p.kh * p.kw;
  const int nout = p.cout * p.cin * taps;
  const long long total_rows = (long long)p.n * p.ho;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  long long row1 = row0 + rows_per_block;
  if (row1 > total_rows) row1 = total_rows;

  int oidx[kWgOPT], co[kWgOPT], ci[kWgOPT], ky[kWgOPT], kx[kWgOPT];
  float acc[kWgOPT];
#pragma unroll
  for (int j = 0; j < kWgOPT; ++j) {
    oidx[j] = (blockIdx.y * kWgOPT + j) * blockDim.x + threadIdx.x;
    acc[j] = 0.f;
    int t = oidx[j] < nout ? oidx[j] : 0;
    kx[j] = t % p.kw; t /= p.kw;
    ky[j] = t % p.kh; t /= p.kh;
    ci[j] = t % p.cin; co[j] = t / p.cin;
  }
  for (long long r = row0; r < row1; ++r) {
    const int n = (int)(r / p.ho), oy = (int)(r % p.ho);
    const size_t dybase = ((size_t)n * p.ho + oy) * p.wo * p.y_ld;
#pragma unroll
    for (int j = 0; j < kWgOPT; ++j) {
      if (oidx[j] >= nout) continue;
      const int iy = oy * p.stride - p.pad + ky[j];
      if (iy < 0 || iy >= p.h) continue;
      const size_t xbase = ((size_t)n * p.h + iy) * p.w * p.x_ld;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      // valid ox range for this tap, then 4 independent accumulators (the loads are L1 hits: latency-bound)
      int ox0 = 0, ox1 = p.wo;
      while (ox0 < ox1 && ox0 * p.stride - p.pad + kx[j] < 0) ++ox0;
      while (ox1 > ox0 && (ox1 - 1) * p.stride - p.pad + kx[j] >= p.w) --ox1;
      const size_t dyo = dybase + co[j];
      const long long xo = (long long)xbase + ci[j] + (long long)(kx[j] - p.pad) * p.x_ld;
      const size_t xs = (size_t)p.stride * p.x_ld;
      int ox = ox0;
      for (; ox + 3 < ox1; ox += 4) {
        a0 = fmaf(load_act(p.y, dyo + (size_t)ox * p.y_ld, p.y_f32), load_act(p.x, xo + ox * xs, p.x_f32), a0);
        a1 = fmaf(load_act(p.y, dyo + (size_t)(ox + 1) * p.y_ld, p.y_f32), load_act(p.x, xo + (ox + 1) * xs, p.x_f32), a1);
        a2 = fmaf(load_act(p.y, dyo + (size_t)(ox + 2) * p.y_ld, p.y_f32), load_act(p.x, xo + (ox + 2) * xs, p.x_f32), a2);
        a3 = fmaf(load_act(p.y, dyo + (size_t)(ox + 3) * p.y_ld, p.y_f32), load_act(p.x, xo + (ox + 3) * xs, p.x_f32), a3);
      }
      for (; ox < ox1; ++ox)
        a0 = fmaf(load_act(p.y, dyo + (size_t)ox * p.y_ld, p.y_f32), load_act(p.x, xo + ox * xs, p.x_f32), a0);
      acc[j] += (a0 + a1) + (a2 + a3);
    }
  }
#pragma unroll
  for (int j = 0; j < kWgOPT; ++j)
    if (oidx[j] < nout && acc[j] != 0.f) acc_add_at(dw, dw_q, oidx[j], acc[j]);

  if (dbias != nullptr && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < p.cout; c += blockDim.x) {
      float s = 0.f;
      for (long long r = row0; r < row1; ++r) {
        const size_t dybase = (size_t)r * p.wo * p.y_ld;
        for (int ox = 0; ox < p.wo; ++ox) s += load_act(p.y, dybase + (size_t)ox * p.y_ld + c, p.y_f32);
      }
      acc_add_at(dbias, dbias_q, c, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused backward of the 1x1 heads (network/ugan.py:70-83, network/blocks.py:166): one pass over the pixels gives
// dx (bf16), dW and dbias (fp32 atomics); optional tanh' from the saved output.  x: (npix, 16) bf16.
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256)
head1x1_bwd_kernel(const uint4* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                   const float* __restrict__ w, uint4* __restrict__ dx, float* __restrict__ dw,
                   float* __restrict__ db, long long* dw_q, long long* db_q, long long npix) {
  pdl_prologue();
  __shared__ float sh[8][COUT * 17];      // per-warp slots, summed in warp order (no shared-memory float atomics)
  float wr[COUT][16], aw[COUT][16], ab[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    ab[o] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { wr[o][c] = w[o * 16 + c]; aw[o][c] = 0.f; }
  }
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    float xv[16], g[COUT], d[16];
    unpack8(x[2 * p], xv);
    unpack8(x[2 * p + 1], xv + 8);
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      g[o] = dy[p * COUT + o];
      if (y != nullptr) { const float t = y[p * COUT + o]; g[o] *= 1.f - t * t; }
      ab[o] += g[o];
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float a = 0.f;
#pragma unroll
      for (int o = 0; o < COUT; ++o) { a = fmaf(g[o], wr[o][c], a); aw[o][c] = fmaf(g[o], xv[c], aw[o][c]); }
      d[c] = a;
    }
    if (dx != nullptr) { dx[2 * p] = pack8(d); dx[2 * p + 1] = pack8(d + 8); }
  }
  __syncthreads();
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float v = warp_sum(aw[o][c]);
      if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][o * 17 + c] = v;
    }
    const float v = warp_sum(ab[o]);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][o * 17 + 16] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * 17; i += blockDim.x) {
    const int o = i / 17, c = i - o * 17;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][i];
    if (c < 16) { if (dw) acc_add_at(dw, dw_q, o * 16 + c, t); }
    else if (db) acc_add_at(db, db_q, o, t);
  }
}

// ---------------------------------------------------------------------------------------------
// The discriminator's stem (network/ugan.py:202: Conv2d(1, 16, 4, stride 2, padding 1) + bias) has its own backward
// kernels: the generic ones above ran 60-110 us on it (17 MB of traffic), three to eight times per iteration, and the
// timeline of the captured step showed them ON the critical path of the discriminator phase (the weight gradient
// held D's Adam step back by 0.8 ms).  x / dx: (N, H, W, 1) fp32; dy: (N, H/2, W/2, 16) bf16.
// ---------------------------------------------------------------------------------------------
constexpr int kStemRows = 8;      // output rows per CTA

// dW[co][0][ky][kx] += sum_p dy[p][co] x[2 oy - 1 + ky][2 ox - 1 + kx];  dbias[co] += sum_p dy[p][co]
// 256 threads = 16 (co) x 16 (tap); one output row at a time is staged in shared memory (dy row as fp32, the four
// input rows it touches with a zero halo) and every thread walks the row with two conflict-free LDS per FMA.
__global__ void __launch_bounds__(256) stem4x4_wgrad_kernel(const float* __restrict__ x, const uint4* __restrict__ dy,
                                                            float* dw, float* dbias, long long* dw_q, long long* db_q,
                                                            int n_rows_total, int h, int w, int ho, int wo) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int xs = w + 4;                 // row pitch: == 4 (mod 32) for w % 32 == 0 -> the 16 taps hit 16 banks
  float* x_s = sm;                      // [4][xs], column c holds input column c - 1
  float* dy_s = sm + 4 * xs;            // [wo][16]
  const int t = threadIdx.x, co = t >> 4, tap = t & 15, ky = tap >> 2, kx = tap & 3;
  float acc = 0.f, accb = 0.f;
  const int row0 = blockIdx.x * kStemRows;
  for (int r = row0; r < row0 + kStemRows && r < n_rows_total; ++r) {
    const int n = r / ho, oy = r - n * ho;
    __syncthreads();
    for (int i = t; i < 4 * xs; i += 256) {
      const int rr = i / xs, c = i - rr * xs;
      const int iy = 2 * oy - 1 + rr, ix = c - 1;
      x_s[i] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? x[((size_t)n * h + iy) * w + ix] : 0.f;
    }
    const uint4* drow = dy + ((size_t)n * ho + oy) * wo * 2;
    for (int i = t; i < wo * 2; i += 256) {
      float v[8];
      unpack8(drow[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy_s[i * 8 + j] = v[j];
    }
    __syncthreads();
    const float* xr = x_s + ky * xs + kx;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f;
    int ox = 0;
    for (; ox + 1 < wo; ox += 2) {
      const float d0 = dy_s[ox * 16 + co], d1 = dy_s[ox * 16 + 16 + co];
      a0 = fmaf(d0, xr[2 * ox], a0);
      a1 = fmaf(d1, xr[2 * ox + 2], a1);
      b0 += d0 + d1;
    }
    for (; ox < wo; ++ox) {
      const float d0 = dy_s[ox * 16 + co];
      a0 = fmaf(d0, xr[2 * ox], a0);
      b0 += d0;
    }
    acc += a0 + a1;
    accb += b0;
  }
  acc_add_at(dw, dw_q, (size_t)co * 16 + tap, acc);
  if (dbias != nullptr && tap == 0) acc_add_at(dbias, db_q, (size_t)co, accb);
}

// dx[n][iy][ix] = sum_{co, ky, kx} dy[n][(iy + 1 - ky) / 2][(ix + 1 - kx) / 2][co] w[co][ky][kx]  (valid, even terms):
// one thread per input pixel, 2 x 2 output pixels x 16 channels = 64 FMAs; weights in shared memory.
__global__ void __launch_bounds__(256) stem4x4_dgrad_kernel(const uint4* __restrict__ dy, const float* __restrict__ wt,
                                                            float* __restrict__ dx, long long npix, int h, int w, int ho,
                                                            int wo) {
  pdl_prologue();
  __shared__ float w_s[16 * 16];        // [ky*4 + kx][co]
  for (int i = threadIdx.x; i < 256; i += 256) {
    const int co = i >> 4, tap = i & 15;
    w_s[tap * 16 + co] = wt[i];
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= npix) return;
  const int ix = (int)(pix % w);
  const long long t = pix / w;
  const int iy = (int)(t % h), n = (int)(t / h);
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ky = ((iy + 1) & 1) + 2 * a;
    const int oy = (iy + 1 - ky) >> 1;
    if (iy + 1 - ky < 0 || oy >= ho) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kx = ((ix + 1) & 1) + 2 * b;
      const int ox = (ix + 1 - kx) >> 1;
      if (ix + 1 - kx < 0 || ox >= wo) continue;
      const uint4* g = dy + (((size_t)n * ho + oy) * wo + ox) * 2;
      float v[16];
      unpack8(g[0], v);
      unpack8(g[1], v + 8);
      const float* wr = w_s + (ky * 4 + kx) * 16;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc = fmaf(v[c], wr[c], acc);
    }
  }
  dx[pix] = acc;
}

// y[n][oy][ox][co] = act( bias[co] + sum_{ky,kx} x[n][2 oy - 1 + ky][2 ox - 1 + kx] w[co][ky][kx] ): one thread per
// output pixel, its 4 x 4 patch in registers (fp32 image), 16 x 16 weights in shared memory, two 128-bit stores.
__global__ void __launch_bounds__(256) stem4x4_fprop_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                            const float* __restrict__ bias, uint4* __restrict__ y,
                                                            long long npix, int h, int w, int ho, int wo, int act,
                                                            float slope) {
  pdl_prologue();
  __shared__ float w_s[16 * 16];        // [tap][co]
  __shared__ float b_s[16];
  {
    const int i = threadIdx.x, co = i >> 4, tap = i & 15;
    w_s[tap * 16 + co] = wt[i];
    if (i < 16) b_s[i] = bias != nullptr ? bias[i] : 0.f;
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= npix) return;
  const int ox = (int)(pix % wo);
  const long long t = pix / wo;
  const int oy = (int)(t % ho), n = (int)(t / ho);
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = b_s[c];
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= h) continue;
    const float* xr = x + ((size_t)n * h + iy) * w;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int ix = 2 * ox - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      const float v = xr[ix];
      const float* wr = w_s + (ky * 4 + kx) * 16;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = apply_act(acc[c], act, slope);
  y[2 * pix] = pack8(acc);
  y[2 * pix + 1] = pack8(acc + 8);
}

static bool is_disc_stem(const DirectParams& p) {
  return p.kh == 4 && p.kw == 4 && p.stride == 2 && p.pad == 1 && p.cin == 1 && p.cout == 16 && p.x_ld == 1 && p.x_f32 &&
         p.y_ld == 16 && !p.y_f32 && !p.accumulate && (p.w & 31) == 0 && p.ho * 2 == p.h && p.wo * 2 == p.w &&
         (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;
}

static int fill_params(const smsut_conv_direct_args* a, DirectParams* p) {
  SMSUT_CHECK(a != nullptr, -1, "null args");
  SMSUT_CHECK(a->n > 0 && a->h > 0 && a->w > 0 && a->cin > 0 && a->cout > 0 && a->stride > 0, -1, "bad conv dims");
  SMSUT_CHECK(a->ho == (a->h + 2 * a->pad - a->kh) / a->stride + 1 && a->wo == (a->w + 2 * a->pad - a->kw) / a->stride + 1,
              -1, "output dims %dx%d inconsistent with input %dx%d k%d s%d p%d", a->ho, a->wo, a->h, a->w, a->kh,
              a->stride, a->pad);
  p->n = a->n; p->h = a->h; p->w = a->w; p->cin = a->cin; p->cout = a->cout; p->kh = a->kh; p->kw = a->kw;
  p->stride = a->stride; p->pad = a->pad; p->ho = a->ho; p->wo = a->wo;
  p->x = a->x; p->x_ld = a->x_ld; p->x_f32 = a->x_f32; p->wt = a->wt; p->bias = a->bias;
  p->y = a->y; p->y_ld = a->y_ld; p->y_f32 = a->y_f32; p->act = a->act; p->slope = a->slope;
  p->accumulate = a->accumulate;
  return 0;
}

static int pick_cch(int taps, int ct, int cdim) {
  int cch = 11264 / (taps * ct);  // <= 44 KB of fp32 weights
  if (cch >= cdim) return cdim;
  cch &= ~7;
  return cch < 1 ? 1 : cch;
}

static int direct_fprop(const smsut_conv_direct_args* a, cudaStream_t stream) {
  DirectParams p;
  int rc = fill_params(a, &p);
  if (rc) return rc;
  const long long npix = (long long)p.n * p.ho * p.wo;
  const int taps = p.kh * p.kw;
  const int cw = p.cout > p.y_ld ? p.cout : p.y_ld;  // channels to write (incl. zero padding)
  if (is_disc_stem(p)) {
    launch_pdl(stem4x4_fprop_kernel, dim3((unsigned)((npix + 255) / 256)), 256, 0, stream, (const float*)p.x, p.wt, p.bias,
               (uint4*)p.y, npix, p.h, p.w, p.ho, p.wo, p.act, p.slope);
    count_launch();
    return launch_status("stem4x4_fprop_kernel");
  }
  if (npix * p.cout <= 4096 && p.y_ld == p.cout && p.cin >= 32) {
    const long long warps = npix * p.cout;
    launch_pdl(direct_fprop_small_kernel, (unsigned)((warps * 32 + 255) / 256), 256, 0, stream, p);
  } else if (cw <= 1) {
    p.cch = pick_cch(taps, 1, p.cin);
    dim3 grid((unsigned)((npix + 127) / 128), 1);
    launch_pdl(direct_fprop_kernel<1>, grid, 128, (size_t)taps * p.cch * 1 * 4, stream, p);
  } else if (cw <= 8) {
    p.cch = pick_cch(taps, 8, p.cin);
    dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((cw + 7) / 8));
    launch_pdl(direct_fprop_kernel<8>, grid, 128, (size_t)taps * p.cch * 8 * 4, stream, p);
  } else {
    p.cch = pick_cch(taps, 16, p.cin);
    dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((cw + 15) / 16));
    launch_pdl(direct_fprop_kernel<16>, grid, 128, (size_t)taps * p.cch * 16 * 4, stream, p);
  }
  count_launch();
  return launch_status("direct_fprop_kernel");
}

static int direct_dgrad(const smsut_conv_direct_args* a, cudaStream_t stream) {
  DirectParams p;
  int rc = fill_params(a, &p);
  if (rc) return rc;
  const long long npix = (long long)p.n * p.h * p.w;
  const int taps = p.kh * p.kw;
  const int cw = p.cin > p.x_ld ? p.cin : p.x_ld;
  if (is_disc_stem(p)) {
    launch_pdl(stem4x4_dgrad_kernel, dim3((unsigned)((npix + 255) / 256)), 256, 0, stream, (const uint4*)p.y, p.wt,
               (float*)const_cast<void*>(p.x), npix, p.h, p.w, p.ho, p.wo);
    count_launch();
    return launch_status("stem4x4_dgrad_kernel");
  }
  if (cw <= 1) {
    p.cch = pick_cch(taps, 1, p.cout);
    dim3 grid((unsigned)((npix + 127) / 128), 1);
    launch_pdl(direct_dgrad_kernel<1>, grid, 128, (size_t)taps * p.cch * 1 * 4, stream, p);
  } else if (cw <= 8) {
    p.cch = pick_cch(taps, 8, p.cout);
    dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((cw + 7) / 8));
    launch_pdl(direct_dgrad_kernel<8>, grid, 128, (size_t)taps * p.cch * 8 * 4, stream, p);
  } else {
    p.cch = pick_cch(taps, 16, p.cout);
    dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((cw + 15) / 16));
    launch_pdl(direct_dgrad_kernel<16>, grid, 128, (size_t)taps * p.cch * 16 * 4, stream, p);
  }
  count_launch();
  return launch_status("direct_dgrad_kernel");
}

static int direct_wgrad(const smsut_conv_direct_args* a, float* dw, float* dbias, cudaStream_t stream) {
  DirectParams p;
  int rc = fill_params(a, &p);
  if (rc) return rc;
  SMSUT_CHECK(dw != nullptr, -1, "null dw");
  if (is_disc_stem(p)) {
    const int rows = p.n * p.ho;
    const size_t smem = (size_t)(4 * (p.w + 4) + p.wo * 16) * sizeof(float);
    if (smem <= 48 * 1024) {
      launch_pdl(stem4x4_wgrad_kernel, dim3((unsigned)((rows + kStemRows - 1) / kStemRows)), 256, smem, stream,
                 (const float*)p.x, (const uint4*)p.y, dw, dbias, det_shadow(dw), det_shadow(dbias), rows, p.h, p.w, p.ho,
                 p.wo);
      count_launch();
      return launch_status("stem4x4_wgrad_kernel");
    }
  }
  const int nout = p.cout * p.cin * p.kh * p.kw;
  const long long total_rows = (long long)p.n * p.ho;
  const int chunks = (nout + 256 * kWgOPT - 1) / (256 * kWgOPT);
  long long want_blocks = 4LL * device_sm_count() / chunks;
  if (want_blocks < 1) want_blocks = 1;
  int rows_per_block = (int)((total_rows + want_blocks - 1) / want_blocks);
  if (rows_per_block < 1) rows_per_block = 1;
  dim3 grid((unsigned)((total_rows + rows_per_block - 1) / rows_per_block), (unsigned)chunks);
  launch_pdl(direct_wgrad_kernel, grid, 256, 0, stream, p, dw, dbias, det_shadow(dw), det_shadow(dbias), rows_per_block);
  count_launch();
  return launch_status("direct_wgrad_kernel");
}

}  // namespace smsut

extern "C" int smsut_head1x1_bwd(const void* x, const float* dy, const float* y, const float* w, void* dx, float* dw,
                                 float* db, int64_t npix, int32_t cin, int32_t cout, smsut_stream_t s) {
  using namespace smsut;
  SMSUT_CHECK(cin == 16 && cout >= 1 && cout <= 8 && npix > 0, -1, "head1x1_bwd supports Cin = 16, Cout <= 8");
  long long blocks = (npix + 255) / 256;
  const long long cap = 4LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
#define HEAD_CASE(C) case C: launch_pdl(head1x1_bwd_kernel<C>, (unsigned)blocks, 256, 0, st, (const uint4*)x, dy, y, w, (uint4*)dx, dw, db, det_shadow(dw), det_shadow(db), npix); break;
  switch (cout) { HEAD_CASE(1) HEAD_CASE(2) HEAD_CASE(3) HEAD_CASE(4) HEAD_CASE(5) HEAD_CASE(6) HEAD_CASE(7) HEAD_CASE(8) }
#undef HEAD_CASE
  count_launch();
  return launch_status("head1x1_bwd_kernel");
}
extern "C" int smsut_conv_direct_fprop(const smsut_conv_direct_args* a, smsut_stream_t s) {
  return smsut::direct_fprop(a, reinterpret_cast<cudaStream_t>(s));
}
extern "C" int smsut_conv_direct_dgrad(const smsut_conv_direct_args* a, smsut_stream_t s) {
  return smsut::direct_dgrad(a, reinterpret_cast<cudaStream_t>(s));
}
extern "C" int smsut_conv_direct_wgrad(const smsut_conv_direct_args* a, float* dw, float* dbias, smsut_stream_t s) {
  return smsut::direct_wgrad(a, dw, dbias, reinterpret_cast<cudaStream_t>(s));
}
