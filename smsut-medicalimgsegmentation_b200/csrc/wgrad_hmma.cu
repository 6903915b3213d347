// Weight gradient of the WIDE, narrow-channel convolutions (W % 128 == 0, 16 / 32 input channels: the 256x256 and
// 128x128 levels of the generator, the first blocks of the discriminator and of the U-Net) on warp-level tensor-core
// MMAs:
//   dW[co][ci][ky][kx] = sum_{n,h,w} dy[n,h,w,co] * x[n, h+ky-r, w+kx-r, ci]
// The contraction runs over ~1 M pixels and produces a 16..64 x (taps x 16..32) matrix: HBM-bound (x and dy are each
// read once: 67 MB for 16->16 @ 256x256, 10 us at the copy peak) and as skinny as a GEMM gets.  The tcgen05 version
// (wgrad_band.cu) has to feed that shape through M = 128 / N >= 16 UMMAs that are 37 % occupied and all issued by one
// thread -- 24 of them per 128-pixel row -- and measured 36-69 us per launch; skipping it altogether shortened the
// captured iteration from 10.2 to 8.9 ms, which makes it the most expensive kernel family of the step.  Here every
// warp of the CTA issues its own m16n8k16 bf16 MMAs (fp32 accumulate), so the issue rate scales with the warps.
// RESULT: correct on every shape (tests/test_kernels_gpu.py::test_wgrad_warp_mma_kernel) but SLOWER than the tcgen05
// kernel -- the warp-level MMA path of sm_100 is too slow for 4.8 GFLOP per launch (see wgrad_hmma_try) -- so it is
// opt-in and kept as a measured negative result.
//
//   CTA   = one strip: 128 pixels x R rows of one image (the grid is one wave of two CTAs per SM)
//   warp  = one "job" (16 output channels, one vertical tap ky) x one pixel group (chunks c = g, g + PG, ...)
//   row   : the dy row and the 2r+1 x rows it meets sit in shared-memory rings, staged once each by cp.async into
//           rows with a 16-byte pad per pixel (conflict-free ldmatrix); zero padding is written, not fetched
//   chunk : 16 pixels (the MMA's K).  A = dy^T (co x px) and B = x (px x ci) are both stored pixel-major, so both
//           fragments come from ldmatrix.trans; a horizontal tap kx is the same x row read kx pixels further right
//   end   : the pixel groups' accumulators are summed in shared memory in a fixed order, then one add per weight
//           element into the tap-major scratch / the OIHW gradient (fixed-point shadow in deterministic mode)
#include <stdlib.h>
#include <string.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

constexpr int kHmMaxWarps = 16;
constexpr int kHmMaxDepth = 6;

struct WgradHmmaParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  float* dw;
  long long* dw_q;
  int n, h, w, ks, r;
  int x_c, x_ld, dy_c, dy_ld;
  int wtiles, segs, rows_per_seg;
  int jobs, pg, nwarps;          // jobs = (dy_c / 16) * ks; pixel groups; warps = jobs * pg
  int xs_slots, ds_slots;        // ring depths
  int depth;                     // rows staged ahead of the one being multiplied (1..kHmMaxDepth)
  int x_pitch, d_pitch;          // bytes per pixel in shared memory (channels * 2 + 16)
  int x_row_bytes, d_row_bytes;  // bytes per ring slot
  int tap_major, cout_total, cout, cin_total, ci_off, c_valid, taps;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// wait until at most `n` of the most recently committed groups are still in flight (n is a small runtime value)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    default: cp_async_wait<6>(); break;
  }
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// KS = kernel size, XT = x_c / 16 (1 or 2)
template <int KS, int XT>
__global__ void __launch_bounds__(kHmMaxWarps * 32, 1) wgrad_hmma_kernel(const WgradHmmaParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int R = KS / 2;
  constexpr int NT = KS * XT * 2;        // n-tiles (8 input channels each) per job: kx x ci
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nthreads = blockDim.x;
  // strip
  int b = blockIdx.x;
  const int seg = b % p.segs; b /= p.segs;
  const int wt = b % p.wtiles; b /= p.wtiles;
  const int n = b;
  const int w0 = wt * 128;
  const int h0 = seg * p.rows_per_seg;
  int h1 = h0 + p.rows_per_seg;
  if (h1 > p.h) h1 = p.h;
  const int nrows = h1 - h0;
  // shared memory: x ring | dy ring | reduction buffer (aliases the rings after the last row)
  uint8_t* x_ring = smem;
  uint8_t* d_ring = smem + (size_t)p.xs_slots * p.x_row_bytes;
  const uint32_t x_ring_s = smem_u32(x_ring), d_ring_s = smem_u32(d_ring);
  const int xpx = 128 + 2 * R;                      // pixels per staged x row (smem pixel 0 = image column w0 - R)
  const int xparts = p.x_c >> 3, dparts = p.dy_c >> 3;   // 16-byte parts per pixel

  // stage x row `hin` (image row, may be outside -> zeros) into slot `slot`; dy row likewise
  auto stage_x = [&](int hin, int slot) {
    uint8_t* dst = x_ring + (size_t)slot * p.x_row_bytes;
    const bool row_ok = hin >= 0 && hin < p.h;
    const __nv_bfloat16* src_row = p.x + ((size_t)n * p.h + (row_ok ? hin : 0)) * p.w * p.x_ld;
    for (int i = threadIdx.x; i < xpx * xparts; i += nthreads) {
      const int px = i / xparts, part = i - px * xparts;
      const int col = w0 - R + px;
      uint8_t* d = dst + (size_t)px * p.x_pitch + part * 16;
      if (row_ok && col >= 0 && col < p.w)
        cp_async16(smem_u32(d), src_row + (size_t)col * p.x_ld + part * 8);
      else
        *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto stage_d = [&](int hrow, int slot) {
    uint8_t* dst = d_ring + (size_t)slot * p.d_row_bytes;
    const __nv_bfloat16* src_row = p.dy + (((size_t)n * p.h + hrow) * p.w + w0) * p.dy_ld;
    for (int i = threadIdx.x; i < 128 * dparts; i += nthreads) {
      const int px = i / dparts, part = i - px * dparts;
      cp_async16(smem_u32(dst + (size_t)px * p.d_pitch + part * 16), src_row + (size_t)px * p.dy_ld + part * 8);
    }
  };

  // this warp's job
  const bool active = warp < p.nwarps;
  const int job = active ? warp % p.jobs : 0, grp = active ? warp / p.jobs : 0;
  const int cot = job / KS, ky = job - cot * KS;
  float acc[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  // per-lane ldmatrix row offsets (bytes): matrix mi = lane / 8, row r8 = lane % 8
  const int mi = lane >> 3, r8 = lane & 7;
  //   A (dy^T): m0 (co 0-7, px 0-7) m1 (co 8-15, px 0-7) m2 (co 0-7, px 8-15) m3 (co 8-15, px 8-15)
  const uint32_t a_lane = (uint32_t)(((mi >> 1) * 8 + r8) * p.d_pitch + (cot * 16 + (mi & 1) * 8) * 2);
  //   B (x):    m0 (px 0-7, ci 0-7) m1 (px 8-15, ci 0-7) m2 (px 0-7, ci 8-15) m3 (px 8-15, ci 8-15)
  const uint32_t b_lane = (uint32_t)(((mi & 1) * 8 + r8) * p.x_pitch + ((mi >> 1) * 8) * 2);

  // One cp.async group per output row: group g carries what row g adds to the rings (x row h0 + g + R, dy row h0 + g;
  // group 0 also the first window's x rows h0-R .. h0+R-1).  `depth` groups stay in flight ahead of the row being
  // multiplied: one row of lookahead left the loop waiting a full DRAM round trip per row (65 us per launch).
  const int D = p.depth;
  for (int j = 0; j < 2 * R; ++j) stage_x(h0 - R + j, j % p.xs_slots);
  for (int g = 0; g < D; ++g) {
    if (g < nrows) {
      stage_x(h0 + g + R, (g + 2 * R) % p.xs_slots);
      stage_d(h0 + g, g % p.ds_slots);
    }
    cp_async_commit();
  }
  for (int i = 0; i < nrows; ++i) {
    cp_async_wait_dyn(D - 1);      // groups 0 .. i have landed (D - 1 newer ones may still be in flight)
    __syncthreads();
    if (active) {
      const uint32_t d_row = d_ring_s + (uint32_t)((i % p.ds_slots) * p.d_row_bytes) + a_lane;
      const uint32_t x_row = x_ring_s + (uint32_t)(((i + ky) % p.xs_slots) * p.x_row_bytes) + b_lane;
      for (int c = grp; c < 8; c += p.pg) {
        uint32_t a[4];
        ldmatrix_x4_trans(d_row + (uint32_t)(c * 16 * p.d_pitch), a);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
          for (int xt = 0; xt < XT; ++xt) {
            uint32_t bb[4];
            ldmatrix_x4_trans(x_row + (uint32_t)((c * 16 + kx) * p.x_pitch + xt * 32), bb);
            mma_bf16_16816(acc[(kx * XT + xt) * 2], a, bb[0], bb[1]);
            mma_bf16_16816(acc[(kx * XT + xt) * 2 + 1], a, bb[2], bb[3]);
          }
        }
      }
    }
    __syncthreads();               // every warp is done with row i: its oldest slots may be refilled
    if (i + D < nrows) {
      stage_x(h0 + i + D + R, (i + D + 2 * R) % p.xs_slots);
      stage_d(h0 + i + D, (i + D) % p.ds_slots);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncthreads();

  // reduction over the pixel groups in a fixed order, through shared memory (the rings are free now):
  // red[job][row 16][col NT * 8] fp32
  float* red = reinterpret_cast<float*>(smem);
  const int cols = NT * 8;
  const int row_a = lane >> 2, col_a = (lane & 3) * 2;
  for (int g = 0; g < p.pg; ++g) {
    if (active && grp == g) {
      float* base = red + (size_t)job * 16 * cols;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float* q0 = base + (size_t)row_a * cols + t * 8 + col_a;
        float* q1 = base + (size_t)(row_a + 8) * cols + t * 8 + col_a;
        if (g == 0) {
          q0[0] = acc[t][0]; q0[1] = acc[t][1]; q1[0] = acc[t][2]; q1[1] = acc[t][3];
        } else {
          q0[0] += acc[t][0]; q0[1] += acc[t][1]; q1[0] += acc[t][2]; q1[1] += acc[t][3];
        }
      }
    }
    __syncthreads();
  }
  // one add per weight element: job = (cot, ky), row = co within the tile, col = (kx, ci)
  const int total = p.jobs * 16 * cols;
  for (int i = threadIdx.x; i < total; i += nthreads) {
    const int j = i / (16 * cols), rem = i - j * (16 * cols);
    const int row = rem / cols, col = rem - row * cols;
    const int jcot = j / KS, jky = j - jcot * KS;
    const int kx = col / (XT * 16), ci = col - kx * (XT * 16);
    const int co = jcot * 16 + row;
    if (co >= p.cout || ci >= p.c_valid) continue;
    const int tap = jky * KS + kx;
    const size_t idx = p.tap_major ? ((size_t)tap * p.cout_total + co) * p.cin_total + p.ci_off + ci
                                   : ((size_t)co * p.cin_total + p.ci_off + ci) * p.taps + tap;
    acc_add_at(p.dw, p.dw_q, idx, red[i]);
  }
}

// returns 1 if handled, 0 if not eligible, < 0 on error
int wgrad_hmma_try(const smsut_wgrad_tc_args* a, cudaStream_t stream) {
  if (a->kind != SMSUT_TC_CONV) return 0;
  if (!(a->ksize == 1 || a->ksize == 3 || a->ksize == 5)) return 0;
  if (a->w % 128 != 0) return 0;
  if (!(a->x_c == 16 || a->x_c == 32)) return 0;
  if (!(a->dy_c == 16 || a->dy_c == 32 || a->dy_c == 64)) return 0;
  if (a->x_ld % 8 != 0 || a->dy_ld % 8 != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) != 0 || (reinterpret_cast<uintptr_t>(a->dy) & 15) != 0) return 0;
  {
    // OPT-IN (SMSUT_WGRAD_HMMA=1; read per call so that the parity test can switch it).  Measured on B200 and NOT the
    // default: 64.9 us against 58.0 us of the tcgen05 band kernel for 16->16 @ 256x256 x 16 slices (88 vs 56 us for
    // 32->32 @ 128x128), 10.8 vs 10.2 ms per captured iteration.  Deeper cp.async lookahead changed nothing: the
    // kernel is bound by the legacy warp-level MMA path itself, which on sm_100 sustains only ~70-150 TFLOP/s --
    // 4.8 GFLOP per launch cannot finish under ~35 us there, while HBM would allow 10 us.
    const char* e = getenv("SMSUT_WGRAD_HMMA");
    if (!(e && e[0] == '1')) return 0;
  }
  WgradHmmaParams p;
  memset(&p, 0, sizeof(p));
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  p.dy = reinterpret_cast<const __nv_bfloat16*>(a->dy);
  p.dw = a->dw;
  p.dw_q = det_shadow(a->dw);
  p.n = a->n; p.h = a->h; p.w = a->w; p.ks = a->ksize; p.r = a->ksize / 2;
  p.x_c = a->x_c; p.x_ld = a->x_ld; p.dy_c = a->dy_c; p.dy_ld = a->dy_ld;
  p.jobs = (a->dy_c / 16) * a->ksize;
  if (p.jobs > kHmMaxWarps) return 0;
  p.pg = 8 / p.jobs;
  if (p.pg < 1) p.pg = 1;
  if (p.pg > 8) p.pg = 8;
  p.nwarps = p.jobs * p.pg;
  p.x_pitch = a->x_c * 2 + 16;
  p.d_pitch = a->dy_c * 2 + 16;
  p.x_row_bytes = ((128 + 2 * p.r) * p.x_pitch + 127) & ~127;
  p.d_row_bytes = (128 * p.d_pitch + 127) & ~127;
  // lookahead: as deep as ~100 KB of rings allow (two CTAs per SM)
  p.depth = kHmMaxDepth;
  while (p.depth > 1 && (size_t)(2 * p.r + p.depth) * p.x_row_bytes + (size_t)p.depth * p.d_row_bytes > 100u * 1024u) --p.depth;
  p.xs_slots = 2 * p.r + p.depth;      // rows i .. i + 2r in use + depth - 1 staged ahead, refilled after the row's barrier
  p.ds_slots = p.depth;
  p.tap_major = a->dw_layout == 1 ? 1 : 0;
  p.cout_total = a->cout_total;
  p.cout = a->dy_c < a->cout_total ? a->dy_c : a->cout_total;
  p.cin_total = a->cin_total; p.ci_off = a->ci_off;
  p.c_valid = a->c_valid > 0 ? a->c_valid : a->x_c;
  p.taps = a->ksize * a->ksize;
  SMSUT_CHECK(a->dw != nullptr, -1, "null dw");
  const int xt = a->x_c / 16;
  const size_t ring = (size_t)p.xs_slots * p.x_row_bytes + (size_t)p.ds_slots * p.d_row_bytes;
  const size_t redb = (size_t)p.jobs * 16 * (a->ksize * xt * 16) * sizeof(float);
  const size_t smem = ring > redb ? ring : redb;
  if (smem > 100u * 1024u) return 0;
  // strips: ONE wave of two CTAs per SM, at least 8 rows each (every strip re-reads 2r halo rows and ends with one
  // add per weight element)
  p.wtiles = a->w / 128;
  const int sms = device_sm_count();
  const int base = a->n * p.wtiles;
  int segs = (2 * sms) / base;
  if (segs < 1) segs = 1;
  int rows = (a->h + segs - 1) / segs;
  if (rows < 8) rows = a->h < 8 ? a->h : 8;
  p.rows_per_seg = rows;
  p.segs = (a->h + rows - 1) / rows;
  const unsigned grid = (unsigned)(a->n * p.wtiles * p.segs);
  const unsigned threads = (unsigned)(((p.nwarps + 3) / 4) * 4 * 32);     // whole warps, a multiple of 128 threads
  bool launched = false;
#define HM_CASE(KS_, XT_)                                                                                           \
  if (!launched && a->ksize == KS_ && xt == XT_) {                                                                  \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      SMSUT_CUDA_OK(cudaFuncSetAttribute(wgrad_hmma_kernel<KS_, XT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         100 * 1024));                                                              \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    launch_pdl(wgrad_hmma_kernel<KS_, XT_>, dim3(grid), dim3(threads), smem, stream, p);                            \
    launched = true;                                                                                                \
  }
  HM_CASE(1, 1) HM_CASE(1, 2) HM_CASE(3, 1) HM_CASE(3, 2) HM_CASE(5, 1) HM_CASE(5, 2)
#undef HM_CASE
  if (!launched) return 0;
  count_launch();
  int st = launch_status("wgrad_hmma_kernel");
  return st ? st : 1;
}

}  // namespace smsut
