// Optimisers over flat fp32 parameter buffers, bf16 weight packing for the tensor-core convs, and the
// error / bookkeeping plumbing shared by all translation units.
//
// Reference semantics: SGD(momentum 0.9, weight_decay 1e-3) and Adam((0.9, 0.999), weight_decay 1e-3, L2 form)
// (trainer/uganShp0Trainer.py:72-74), poly LR set after each step (trainer/uganConsisTrainer.py:198-203),
// EMA teacher update (trainer/meanTeacherTrainer.py:63-69).
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include <stdlib.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SMSUT_PDL");
    on = (e && e[0] == '1') ? 1 : 0;   // measured on B200: 14.9 ms/step with, 14.6 without (graph launches already pipeline)
  }
  return on == 1;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom,
                           long long count, const float* __restrict__ lr, float momentum, float wd, float gscale) {
  pdl_prologue();
  const float step = lr[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float d = fmaf(wd, w, g[i] * gscale);
    const float b = fmaf(momentum, mom[i], d);
    mom[i] = b;
    p[i] = w - step * b;
  }
}

__global__ void tick_kernel(float* state) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) state[0] += 1.f;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long count, const float* __restrict__ lr, float b1, float b2,
                            float eps, float wd, const float* __restrict__ state, float gscale) {
  pdl_prologue();
  const float t = state[0];
  const float bc1 = 1.f - powf(b1, t);
  const float bc2 = 1.f - powf(b2, t);
  const float step = lr[0] / bc1;
  const float rs2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float d = fmaf(wd, w, g[i] * gscale);
    const float mm = fmaf(b1, m[i], (1.f - b1) * d);
    const float vv = fmaf(b2, v[i], (1.f - b2) * d * d);
    m[i] = mm;
    v[i] = vv;
    p[i] = w - step * mm / (sqrtf(vv) * rs2 + eps);
  }
}

__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ p, long long count,
                           const float* __restrict__ alpha) {
  pdl_prologue();
  const float a = alpha[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    ema[i] = fmaf(a, ema[i], (1.f - a) * p[i]);
}

// lr_out = base * (1 - max(iter-1, 0)/max_iter)^power, then iter += 1.
// (the reference sets the LR *after* step k from iter = k, so step k+1 runs with the LR of iter k.)
__global__ void poly_lr_kernel(float* iter_state, float* lr_out, float base, float max_iter, float power) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float it = iter_state[0];
  const float prev = fmaxf(it - 1.f, 0.f);
  lr_out[0] = base * powf(fmaxf(1.f - prev / max_iter, 0.f), power);
  iter_state[0] = it + 1.f;
}

// one (entry, slice) per block; see smsut_pack_entry in the header
__global__ void pack_weights_kernel(const smsut_pack_entry* __restrict__ table) {
  pdl_prologue();
  const smsut_pack_entry e = table[blockIdx.x];
  const int taps = e.kh * e.kw;
  __nv_bfloat16* f = reinterpret_cast<__nv_bfloat16*>(e.fprop);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(e.dgrad);
  // every weight has < 2^31 elements: 32-bit index arithmetic (the 64-bit divisions of the first version made this
  // launch 29 us for 25 MB of traffic); blockIdx.y strides over output rows, threads over the row
  if (!e.transposed) {
    // fprop: [cout_pad][taps][cin_pad]
    const int rowf = taps * e.cin_pad;
    for (int co = blockIdx.y; co < e.cout_pad; co += gridDim.y) {
      const float* wr = e.w + (size_t)co * e.cin * taps;
      for (int i = threadIdx.x; i < rowf; i += blockDim.x) {
        const int t = i / e.cin_pad, ci = i - t * e.cin_pad;
        f[(size_t)co * rowf + i] = f2bf((co < e.cout && ci < e.cin) ? wr[ci * taps + t] : 0.f);
      }
    }
    if (d != nullptr) {
      // dgrad: [cin_pad][taps (flipped)][cout_pad]
      const int rowd = taps * e.cout_pad;
      for (int ci = blockIdx.y; ci < e.cin_pad; ci += gridDim.y) {
        for (int i = threadIdx.x; i < rowd; i += blockDim.x) {
          const int t = i / e.cout_pad, co = i - t * e.cout_pad;
          d[(size_t)ci * rowd + i] =
              f2bf((co < e.cout && ci < e.cin) ? e.w[((size_t)co * e.cin + ci) * taps + (taps - 1 - t)] : 0.f);
        }
      }
    }
  } else {
    // ConvTranspose2d weight (cin, cout, kh, kw): fprop rows (t, co) x cin ; dgrad rows ci x (t, co)
    const int rows = taps * e.cout;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
      const int t = r / e.cout, co = r - t * e.cout;
      for (int ci = threadIdx.x; ci < e.cin; ci += blockDim.x)
        f[(size_t)r * e.cin + ci] = f2bf(e.w[((size_t)ci * e.cout + co) * taps + t]);
    }
    if (d != nullptr) {
      for (int ci = blockIdx.y; ci < e.cin; ci += gridDim.y) {
        for (int i = threadIdx.x; i < rows; i += blockDim.x) {
          const int t = i / e.cout, co = i - t * e.cout;
          d[(size_t)ci * rows + i] = f2bf(e.w[((size_t)ci * e.cout + co) * taps + t]);
        }
      }
    }
  }
}

static inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 8LL * device_sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace smsut

using namespace smsut;

extern "C" const char* smsut_last_error(void) { return get_last_error(); }
extern "C" int smsut_abi_version(void) { return 1; }
extern "C" int64_t smsut_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int smsut_sgd_step(float* p, const float* g, float* mom, int64_t count, const float* lr, float momentum,
                              float weight_decay, float grad_scale, smsut_stream_t st) {
  SMSUT_CHECK(p && g && mom && lr && count > 0, -1, "bad sgd args");
  launch_pdl(sgd_kernel, grid_for(count), 256, 0, (cudaStream_t)st, p, g, mom, count, lr, momentum, weight_decay, grad_scale);
  count_launch();
  return launch_status("sgd_kernel");
}
extern "C" int smsut_adam_step(float* p, const float* g, float* m, float* v, int64_t count, const float* lr,
                               float beta1, float beta2, float eps, float weight_decay, float* state,
                               float grad_scale, smsut_stream_t st) {
  SMSUT_CHECK(p && g && m && v && lr && state && count > 0, -1, "bad adam args");
  launch_pdl(tick_kernel, 1, 32, 0, (cudaStream_t)st, state);
  launch_pdl(adam_kernel, grid_for(count), 256, 0, (cudaStream_t)st, p, g, m, v, count, lr, beta1, beta2, eps, weight_decay,
                                                             state, grad_scale);
  count_launch(); count_launch();
  return launch_status("adam_kernel");
}
extern "C" int smsut_ema_update(float* ema, const float* p, int64_t count, const float* alpha, smsut_stream_t st) {
  SMSUT_CHECK(ema && p && alpha && count > 0, -1, "bad ema args");
  launch_pdl(ema_kernel, grid_for(count), 256, 0, (cudaStream_t)st, ema, p, count, alpha);
  count_launch();
  return launch_status("ema_kernel");
}
extern "C" int smsut_poly_lr_tick(float* iter_state, float* lr_out, float base_lr, float max_iter, float power,
                                  smsut_stream_t st) {
  SMSUT_CHECK(iter_state && lr_out, -1, "bad poly lr args");
  launch_pdl(poly_lr_kernel, 1, 32, 0, (cudaStream_t)st, iter_state, lr_out, base_lr, max_iter, power);
  count_launch();
  return launch_status("poly_lr_kernel");
}
// one entry per blockIdx.x; see smsut_unpack_entry in the header.  Reads are coalesced over (m, c) for a fixed tap,
// the taps-strided writes of a warp cover taps * 128 contiguous bytes (both sides stay in L2: <= 2.4 MB per weight).
__global__ void unpack_wgrads_kernel(const smsut_unpack_entry* __restrict__ table) {
  pdl_prologue();
  const smsut_unpack_entry e = table[blockIdx.x];
  const int mc = e.rows * e.cols, taps = e.taps;
  const int nth = gridDim.y * blockDim.x;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < mc; i += nth) {
    float* dst = e.grad + (size_t)i * taps;
    const float* src = e.scratch + i;
    if (taps == 9) {
      float v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = src[(size_t)t * mc];
#pragma unroll
      for (int t = 0; t < 9; ++t) dst[t] += v[t];
    } else {
      for (int t = 0; t < taps; ++t) dst[t] += src[(size_t)t * mc];
    }
  }
}

extern "C" int smsut_unpack_wgrads(const smsut_unpack_entry* table, int32_t n, smsut_stream_t st) {
  SMSUT_CHECK(table && n > 0, -1, "bad unpack args");
  launch_pdl(unpack_wgrads_kernel, dim3(n, 48), 256, 0, (cudaStream_t)st, table);
  count_launch();
  return launch_status("unpack_wgrads_kernel");
}

extern "C" int smsut_pack_weights(const smsut_pack_entry* table, int32_t n, smsut_stream_t st) {
  SMSUT_CHECK(table && n > 0, -1, "bad pack args");
  launch_pdl(pack_weights_kernel, dim3(n, 96), 256, 0, (cudaStream_t)st, table);
  count_launch();
  return launch_status("pack_weights_kernel");
}
