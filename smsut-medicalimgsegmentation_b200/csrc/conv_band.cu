// "Band" implicit-GEMM convolution for the wide, narrow-channel layers (W % 128 == 0, C in {16, 32, 64} per
// source): the layers whose tap-by-tap TMA re-reads made conv_tc_kernel L2-bound (ncu: 340 MB of L2->SM traffic
// for a 33 MB input, 43 % LTS throughput, 6 % DRAM).
//
// One CTA walks down a strip of 128 pixels x R output rows of one image.  Every input row of the strip (with its
// +-r pixel halo, r = ksize/2) is fetched by TMA ONCE into a ring of row slots; the (2r+1)^2 taps are formed from
// the ring purely by shared-memory descriptor arithmetic: a vertical tap picks another slot, a horizontal tap
// starts the K-major A descriptor (dx + r) rows later (the swizzle phase follows the absolute smem address, so a
// row-shifted start reads the bytes TMA wrote -- verified on hardware with scripts/rowshift_probe.py).  All
// weights of the layer stay resident in shared memory for the CTA's lifetime.  Accumulators are double-buffered in
// TMEM so the epilogue of row h overlaps the MMAs of row h+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
// L2->SM traffic per output row: (128 + 2r) * C * 2 bytes instead of (2r+1)^2 * 128 * C * 2.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_act_map(CUtensorMap* m, const void* ptr, int c, int w, int h, int n, int64_t sw, int64_t sh, int64_t sn,
                 int cc, int tw, int th, int tn);
int make_mat_map(CUtensorMap* m, const void* ptr, int64_t kdim, int64_t rows, int64_t ld, int cc, int rows_box);

constexpr int kBandThreads = 192;
constexpr int kBandMaxSlots = 12;

struct BandParams {
  int n, h, w;
  int ks, r;               // kernel size, radius
  int nsrc, cc;            // sources, channels per source (16 / 32 / 64)
  int ctot;                // nsrc * cc
  int ncols, ncols_pad;    // valid / padded GEMM columns (<= 256)
  int wtiles, segs, rows_per_seg;
  int nslots;
  uint32_t slot_bytes, src_bytes;     // bytes of one ring slot / of one source inside it (1024-aligned)
  uint32_t wtile_bytes;               // bytes of one (tap, source) weight tile
  uint32_t w_bytes;                   // all weights
  uint32_t layout_type, sbo, pitch;
  uint32_t tmem_cols, acc_cols;
  // epilogue (same contract as conv_tc_kernel, mode 0)
  void* out0; int ld0, coff0;
  void* out1; int ld1, coff1, split;
  const float* bias;
  int act; float slope;
  int accumulate, out_f32;
  float* stats;            // fused InstanceNorm statistics [n][2][ncols_pad] (optional)
  long long* stats_q;      // its fixed-point shadow in deterministic mode (common.cuh), else nullptr
  int fast;                // epilogue fast path (see band_epilogue_fast)
  long long* trace;        // development: per-row clock64 stamps of CTA 0 (SMSUT_BAND_TRACE=1)
};

__device__ __forceinline__ void band_store16(void* base, size_t off, const float* v, int nvalid, bool f32,
                                             bool accumulate) {
  if (f32) {
    float* dst = reinterpret_cast<float*>(base) + off;
    for (int i = 0; i < nvalid; ++i) dst[i] = accumulate ? dst[i] + v[i] : v[i];
    return;
  }
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(base) + off;
  if (nvalid == 16) {
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = v[i];
    uint4* p = reinterpret_cast<uint4*>(dst);
    if (accumulate) {
      float o[8];
      unpack8(p[0], o);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += o[i];
      unpack8(p[1], o);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[8 + i] += o[i];
    }
    p[0] = pack8(f);
    p[1] = pack8(f + 8);
  } else {
    for (int i = 0; i < nvalid; ++i) dst[i] = f2bf(accumulate ? bf2f(dst[i]) + v[i] : v[i]);
  }
}


// reduce-scatter over the 32 lanes of a warp: on return vals[0] of lane l holds the warp total of element l
__device__ __forceinline__ float warp_reduce_scatter32(float (&vals)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int k = 0; k < step; ++k) {
      const float send = upper ? vals[k] : vals[k + step];
      const float keep = upper ? vals[k + step] : vals[k];
      vals[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return vals[0];
}

// Epilogue of the common case (bf16 output, one destination, no bias / activation / accumulate, every column valid):
// per row and 16-column chunk one tcgen05.ld, 8 packs and two 128-bit stores per thread.  The InstanceNorm
// statistics (of the values as stored) are carried in registers down the strip -- one FADD + one FFMA per value --
// and leave the CTA through ONE shuffle reduce-scatter + atomics at the end (the per-row shuffle network of the
// generic path made the epilogue warps the SM-level issue bottleneck: ~330 instructions per row and warp).
template <int NCH, bool STATS>
__device__ __forceinline__ void band_epilogue_fast(const BandParams& p, uint64_t* tmem_full, uint64_t* tmem_empty,
                                                   uint32_t tmem_base, int q, int lane, int n, int h_begin, int w,
                                                   int nrows_out) {
  float s1[STATS ? NCH : 1][16], s2[STATS ? NCH : 1][16];
  if (STATS) {
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int k = 0; k < 16; ++k) s1[j][k] = s2[j][k] = 0.f;
  }
  // destination of chunk j: out0 below the split column, out1 above (dgrad of a concatenated input); both bf16
  __nv_bfloat16* out0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + p.coff0;
  __nv_bfloat16* out1 = p.split > 0 ? reinterpret_cast<__nv_bfloat16*>(p.out1) + p.coff1 - p.split : out0;
  const int ld0 = p.ld0, ld1 = p.split > 0 ? p.ld1 : p.ld0;
  const int split_chunk = p.split > 0 ? (p.split >> 4) : NCH;
  const bool acc = p.accumulate != 0;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  size_t pix = ((size_t)n * p.h + h_begin) * p.w + w;
  for (int i = 0; i < nrows_out; ++i, pix += p.w) {
    const int buf = i & 1;
    mbar_wait(&tmem_full[buf], ((uint32_t)i >> 1) & 1u);
    tc_fence_after();
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      uint32_t raw[16];
      tmem_ld16(lane_addr + (uint32_t)buf * p.acc_cols + (uint32_t)(j * 16), raw);
      tmem_ld_wait();
      if (j == NCH - 1) {
        // the accumulator buffer is free as soon as its values sit in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
      uint4* dst = j < split_chunk ? reinterpret_cast<uint4*>(out0 + pix * (size_t)ld0 + j * 16)
                                   : reinterpret_cast<uint4*>(out1 + pix * (size_t)ld1 + j * 16);
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
      if (acc) {
        float o[8];
        const uint4 q0 = dst[0], q1 = dst[1];
        unpack8(q0, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += o[k];
        unpack8(q1, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[8 + k] += o[k];
      }
      uint32_t w32[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w32[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
      dst[0] = make_uint4(w32[0], w32[1], w32[2], w32[3]);
      dst[1] = make_uint4(w32[4], w32[5], w32[6], w32[7]);
      if (STATS) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float lo = __uint_as_float(w32[k] << 16), hi = __uint_as_float(w32[k] & 0xffff0000u);
          s1[j][2 * k] += lo;
          s1[j][2 * k + 1] += hi;
          s2[j][2 * k] = fmaf(lo, lo, s2[j][2 * k]);
          s2[j][2 * k + 1] = fmaf(hi, hi, s2[j][2 * k + 1]);
        }
      }
    }
  }
  if (STATS) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float vals[32];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        vals[k] = s1[j][k];
        vals[16 + k] = s2[j][k];
      }
      const float tot = warp_reduce_scatter32(vals, lane);
      acc_add_at(p.stats, p.stats_q, ((size_t)n * 2 + (lane >> 4)) * p.ncols_pad + j * 16 + (lane & 15), tot);
    }
  }
}

// KS = kernel size, NSRC = sources, KK = channels per source / 16: compile-time so that the single MMA-issuing
// thread runs a fully unrolled stream of descriptor adds + tcgen05.mma (ncu: with runtime loops, modulo slot
// arithmetic and descriptor rebuilds that one thread took ~2.7 us per output row and every other warp waited on it)
// NCH = ncols_pad / 16 for the fast epilogue (1, 2 or 4), 0 = generic epilogue
template <int KS, int NSRC, int KK, int NCH>
__global__ void __launch_bounds__(kBandThreads, 1)
conv_band_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ BandParams p) {
  pdl_trigger();   // the next kernel may be scheduled; this one waits for its predecessor after its own set-up
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t slot_full[kBandMaxSlots];
  __shared__ __align__(8) uint64_t slot_empty[kBandMaxSlots];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring_base = smem_base + ((p.w_bytes + 1023u) & ~1023u);
  uint8_t* ring_ptr = smem_al + ((p.w_bytes + 1023u) & ~1023u);

  // strip coordinates
  int b = blockIdx.x;
  const int seg = b % p.segs; b /= p.segs;
  const int wt = b % p.wtiles; b /= p.wtiles;
  const int n = b;
  const int w0 = wt * 128;
  const int h_begin = seg * p.rows_per_seg;
  int h_end = h_begin + p.rows_per_seg;
  if (h_end > p.h) h_end = p.h;
  const int nrows_out = h_end - h_begin;
  const int nrows_in = nrows_out + 2 * p.r;
  const int ntaps = p.ks * p.ks;

  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    p.trace[512 + blockIdx.x * 4 + 0] = (long long)t;
    p.trace[512 + blockIdx.x * 4 + 2] = (long long)smid;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(&slot_full[s], 1);
      mbar_init(&slot_empty[s], 1);
    }
    mbar_init(&w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();      // barriers, TMEM and descriptors are ready: now the predecessor's results are needed

  if (warp == 0) {
    // ===================== TMA producer (warp-wide loop, elected issue) =====================
    {
      const uint32_t el = elect_one_u32();
      // resident weights: one tile per (tap, source), K coordinate = tap * ctot + source * cc
      mbar_arrive_expect_tx_e(&w_full, p.w_bytes, el);
      for (int t = 0; t < ntaps; ++t)
        for (int s = 0; s < p.nsrc; ++s)
          tma_load_2d_e(smem_al + (size_t)(t * p.nsrc + s) * p.wtile_bytes, &map_w, &w_full, t * p.ctot + s * p.cc, 0, el);
      // input rows h_begin - r .. h_end - 1 + r, each fetched once (out-of-image rows / halo pixels: TMA zero fill)
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t row_bytes = (uint32_t)p.nsrc * (uint32_t)(128 + 2 * p.r) * p.pitch;
      for (int j = 0; j < nrows_in; ++j) {
        mbar_wait(&slot_empty[slot], phase ^ 1u);
        mbar_arrive_expect_tx_e(&slot_full[slot], row_bytes, el);
        uint8_t* dst = ring_ptr + (size_t)slot * p.slot_bytes;
        const int hin = h_begin - p.r + j;
        tma_load_4d_e(dst, &map_a0, &slot_full[slot], 0, w0 - p.r, hin, n, el);
        if (NSRC == 2) tma_load_4d_e(dst + p.src_bytes, &map_a1, &slot_full[slot], 0, w0 - p.r, hin, n, el);
        if (++slot == p.nslots) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.ncols_pad >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t el = elect_one_u32();
    mbar_wait(&w_full, 0);
    // descriptors: hi word constant; lo word = (start >> 4) | (LBO = 1) << 16, advanced by plain adds (16-byte units)
    const uint64_t desc_hi = make_smem_desc(0, 16, p.sbo, p.layout_type) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_flags = 1u << 16;
    const uint32_t ring_lo = (ring_base >> 4) | lo_flags;
    const uint32_t w_lo = (smem_base >> 4) | lo_flags;
    const uint32_t slot_u = p.slot_bytes >> 4, src_u = p.src_bytes >> 4, pitch_u = p.pitch >> 4, wtile_u = p.wtile_bytes >> 4;
    const int nslots = p.nslots;
    int rows_ready = 0;       // input rows whose slot_full barrier has been consumed
    int ready_slot = 0;       // rows_ready % nslots, and its wrap count
    uint32_t ready_phase = 0;
    int base_slot = 0;        // i % nslots
    for (int i = 0; i < nrows_out; ++i) {
      const int buf = i & 1;
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && i < 64;
      if (tr) p.trace[i * 8 + 0] = clock64();
      mbar_wait(&tmem_empty[buf], (((uint32_t)i >> 1) & 1u) ^ 1u);
      if (tr) p.trace[i * 8 + 1] = clock64();
      // output row i needs input rows i .. i + 2r (indices relative to the strip's first input row)
      while (rows_ready <= i + KS - 1) {
        mbar_wait(&slot_full[ready_slot], ready_phase);
        ++rows_ready;
        if (++ready_slot == nslots) { ready_slot = 0; ready_phase ^= 1u; }
      }
      if (tr) p.trace[i * 8 + 2] = clock64();
      tc_fence_after();
      {
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * p.acc_cols;
#pragma unroll
        for (int ty = 0; ty < KS; ++ty) {
          int slot = base_slot + ty;
          if (slot >= nslots) slot -= nslots;
          const uint32_t row_lo = ring_lo + (uint32_t)slot * slot_u;
#pragma unroll
          for (int tx = 0; tx < KS; ++tx) {
#pragma unroll
            for (int s = 0; s < NSRC; ++s) {
              const uint32_t a_lo = row_lo + (uint32_t)s * src_u + (uint32_t)tx * pitch_u;
              const uint32_t b_lo = w_lo + (uint32_t)((ty * KS + tx) * NSRC + s) * wtile_u;
#pragma unroll
              for (int k = 0; k < KK; ++k) {
                umma_bf16_e(d_tmem, desc_hi | (uint64_t)(a_lo + 2u * k), desc_hi | (uint64_t)(b_lo + 2u * k), idesc,
                            (ty | tx | s | k) != 0 ? 1u : 0u, el);
              }
            }
          }
        }
        umma_commit_e(&tmem_full[buf], el);
        umma_commit_e(&slot_empty[base_slot], el);   // the oldest input row of this window is no longer needed
      }
      if (tr) p.trace[i * 8 + 3] = clock64();
      __syncwarp();
      if (++base_slot == nslots) base_slot = 0;
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;          // pixel inside the strip row
    const int w = w0 + r;
    const int nchunks = p.ncols_pad >> 4;
    float st_acc[4] = {0.f, 0.f, 0.f, 0.f};   // lane l: l < 16 -> sum of channel l, else sum of squares of l - 16
    const bool do_stats = p.stats != nullptr;
    constexpr int kNchStats = (NCH == 1 || NCH == 2) ? NCH : 1;
    constexpr int kNchPlain = (NCH >= 1) ? NCH : 1;
    if (NCH > 0) {
      if (NCH <= 2 && do_stats)
        band_epilogue_fast<kNchStats, true>(p, tmem_full, tmem_empty, tmem_base, q, lane, n, h_begin, w, nrows_out);
      else
        band_epilogue_fast<kNchPlain, false>(p, tmem_full, tmem_empty, tmem_base, q, lane, n, h_begin, w, nrows_out);
    } else
    for (int i = 0; i < nrows_out; ++i) {
      const int buf = i & 1;
      const uint32_t use = (uint32_t)(i >> 1);
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64 && i < 64;
      if (tr) p.trace[i * 8 + 4] = clock64();
      mbar_wait(&tmem_full[buf], use & 1u);
      if (tr) p.trace[i * 8 + 5] = clock64();
      tc_fence_after();
      const int h = h_begin + i;
      const size_t pix = ((size_t)n * p.h + h) * p.w + w;
      for (int j = 0; j < nchunks; ++j) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + (uint32_t)buf * p.acc_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
        tmem_ld_wait();
        const int col = j * 16;
        if (col >= p.ncols) continue;
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
        int nvalid = p.ncols - col;
        if (nvalid > 16) nvalid = 16;
        if (p.bias != nullptr) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < nvalid) v[k] += p.bias[col + k];
        }
        if (p.act == SMSUT_ACT_RELU) {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
        } else if (p.act == SMSUT_ACT_LRELU) {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = lrelu(v[k], p.slope);
        }
        void* base;
        int ld, coff, c;
        if (p.split > 0 && col >= p.split) {
          base = p.out1; ld = p.ld1; coff = p.coff1; c = col - p.split;
        } else {
          base = p.out0; ld = p.ld0; coff = p.coff0; c = col;
          if (p.split > 0 && col + nvalid > p.split) nvalid = p.split - col;
        }
        band_store16(base, pix * (size_t)ld + coff + c, v, nvalid, p.out_f32 != 0, p.accumulate != 0);
        if (do_stats && j < 4) {
          // reduce-scatter over the 32 pixels of this warp: 31 shuffles leave the total of value `lane` in vals[0]
          float vals[32];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float r = bf2f(f2bf(v[k]));      // statistics of the values as stored
            vals[k] = r;
            vals[16 + k] = r * r;
          }
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const bool upper = (lane & step) != 0;
#pragma unroll
            for (int k = 0; k < step; ++k) {
              const float send = upper ? vals[k] : vals[k + step];
              const float keep = upper ? vals[k + step] : vals[k];
              vals[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
            }
          }
          st_acc[j] += vals[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (tr) p.trace[i * 8 + 6] = clock64();
    }
    if (do_stats) {
      for (int j = 0; j < nchunks && j < 4; ++j) {
        const int col = j * 16 + (lane & 15);
        if (col < p.ncols)
          acc_add_at(p.stats, p.stats_q, ((size_t)n * 2 + (lane >> 4)) * p.ncols_pad + col, st_acc[j]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[512 + blockIdx.x * 4 + 1] = (long long)t;
  }
}

bool conv_band_eligible(const smsut_conv_tc_args* a) {
  if (a->kind != SMSUT_TC_CONV) return false;
  if (!(a->ksize == 1 || a->ksize == 3 || a->ksize == 5)) return false;
  if (a->ksize == 5 && a->nsrc != 1) return false;
  if (a->w % 128 != 0 || a->bn != 0) return false;
  const int cc = a->src_c[0];
  if (!(cc == 16 || cc == 32 || cc == 64)) return false;
  if (a->nsrc == 2 && a->src_c[1] != cc) return false;
  if (a->ncols_pad > 256 || a->ncols_pad % 16 != 0) return false;
  const uint32_t w_bytes = (uint32_t)(a->ksize * a->ksize) * a->nsrc * a->ncols_pad * cc * 2u;
  if (w_bytes > 96u * 1024u) return false;
  const char* e = getenv("SMSUT_NO_BAND");
  if (e && e[0] == '1') return false;
  return true;
}

bool conv_band_eligible_c(const smsut_conv_tc_args* a) { return conv_band_eligible(a); }

// statistics are fused when the band kernel takes the launch and every column goes to one bf16 destination
bool conv_band_fuses_stats(const smsut_conv_tc_args* a) {
  return conv_band_eligible(a) && a->ncols_pad <= 64 && a->out1 == nullptr && !a->out_f32 && !a->accumulate &&
         a->act == SMSUT_ACT_NONE && a->bias == nullptr;
}

// returns 1 if the launch was handled by the band kernel, 0 if the shape is not eligible, < 0 on error
int conv_band_try(const smsut_conv_tc_args* a, cudaStream_t stream) {
  if (!conv_band_eligible(a)) return 0;
  const int cc = a->src_c[0];
  BandParams p;
  memset(&p, 0, sizeof(p));
  p.n = a->n; p.h = a->h; p.w = a->w;
  p.ks = a->ksize; p.r = a->ksize / 2;
  p.nsrc = a->nsrc; p.cc = cc; p.ctot = cc * a->nsrc;
  p.ncols = a->ncols; p.ncols_pad = a->ncols_pad;
  p.pitch = (uint32_t)cc * 2u;
  p.layout_type = cc == 64 ? 2u : (cc == 32 ? 4u : 6u);
  p.sbo = 8u * p.pitch;
  const int taps = p.ks * p.ks;
  p.wtile_bytes = (uint32_t)a->ncols_pad * p.pitch;
  p.w_bytes = (uint32_t)taps * a->nsrc * p.wtile_bytes;
  if (p.w_bytes > 96u * 1024u) return 0;
  p.src_bytes = (((uint32_t)(128 + 2 * p.r) * p.pitch) + 1023u) & ~1023u;
  p.slot_bytes = p.src_bytes * (uint32_t)a->nsrc;
  int nslots = (int)((200u * 1024u - ((p.w_bytes + 1023u) & ~1023u)) / p.slot_bytes);
  if (nslots > kBandMaxSlots) nslots = kBandMaxSlots;
  if (nslots > 2 * p.r + 1 + 5) nslots = 2 * p.r + 1 + 5;
  if (nslots < 2 * p.r + 2) return 0;
  p.nslots = nslots;
  p.acc_cols = (uint32_t)a->ncols_pad;
  uint32_t tc = 32;
  while (tc < 2u * p.acc_cols) tc <<= 1;
  p.tmem_cols = tc;

  // strip decomposition: ONE balanced wave of co-resident CTAs (as many per SM as registers / shared memory allow,
  // at most 3), at least 4 output rows per strip.  (148 * 3 = 444 slots: 448 CTAs meant a second wave or 4-deep SMs.)
  p.wtiles = a->w / 128;
  const int sms = device_sm_count();
  const size_t smem_need = ((p.w_bytes + 1023u) & ~1023u) + (size_t)p.nslots * p.slot_bytes + 2048;
  int per_sm = (int)((227u * 1024u) / smem_need);
  int reg_cap = (a->ncols_pad >> 4) == 2 ? 2 : 3;
  {
    static int knob = -1;
    if (knob < 0) {
      const char* e = getenv("SMSUT_BAND_PER_SM");
      knob = e && atoi(e) > 0 ? atoi(e) : 0;
    }
    if (knob > 0 && reg_cap > knob) reg_cap = knob;
  }
  if (per_sm > reg_cap) per_sm = reg_cap;
  if (per_sm < 1) per_sm = 1;
  const int strips = a->n * p.wtiles;
  int segs = (per_sm * sms) / strips;
  if (segs < 1) segs = 1;
  int rows = (a->h + segs - 1) / segs;
  if (rows < 4) rows = a->h < 4 ? a->h : 4;
  {
    const char* e = getenv("SMSUT_BAND_ROWS");      // tuning knobs (development)
    if (e && atoi(e) > 0) rows = atoi(e);
    e = getenv("SMSUT_BAND_SLOTS");
    if (e && atoi(e) >= 2 * p.r + 2 && atoi(e) <= kBandMaxSlots) p.nslots = atoi(e);
  }
  p.rows_per_seg = rows;
  p.segs = (a->h + rows - 1) / rows;

  p.out0 = a->out0; p.ld0 = a->out0_ld; p.coff0 = a->out0_coff;
  p.out1 = a->out1; p.ld1 = a->out1_ld; p.coff1 = a->out1_coff; p.split = a->out1 ? a->split : 0;
  p.bias = a->bias; p.act = a->act; p.slope = a->slope;
  p.accumulate = a->accumulate; p.out_f32 = a->out_f32;
  p.stats = nullptr;
  if (a->stats != nullptr) {
    SMSUT_CHECK(conv_band_fuses_stats(a), -1, "stats requested for a shape that does not fuse them");
    p.stats = a->stats;
    p.stats_q = det_shadow(a->stats);
  }
  SMSUT_CHECK(a->out0 != nullptr, -1, "null output");
  if (p.split > 0) SMSUT_CHECK(p.split % 16 == 0, -1, "split must be a multiple of 16");
  {
    const int nch = a->ncols_pad >> 4;
    // plain epilogue (see band_epilogue_fast): bf16, every column valid, no bias / activation; accumulate and a
    // 16-aligned split into two destinations are handled there as well
    const bool split_ok = p.split == 0 || (p.split % 16 == 0 && a->out1_ld % 8 == 0 && a->out1_coff % 8 == 0);
    p.fast = (!a->out_f32 && a->bias == nullptr && a->act == SMSUT_ACT_NONE && split_ok &&
              a->ncols == a->ncols_pad && a->out0_ld % 8 == 0 && a->out0_coff % 8 == 0 &&
              (nch == 1 || nch == 2 || (nch == 4 && p.stats == nullptr)))
                 ? 1 : 0;
    const char* e = getenv("SMSUT_BAND_NOFAST");
    if (e && e[0] == '1') p.fast = 0;
  }

  CUtensorMap maps[2], map_w;
  memset(maps, 0, sizeof(maps));
  for (int s = 0; s < a->nsrc; ++s) {
    int rc = make_act_map(&maps[s], a->src[s], a->src_c[s], a->w, a->h, a->n, a->src_ld[s], (int64_t)a->src_ld[s] * a->w,
                          (int64_t)a->src_ld[s] * a->w * a->h, cc, 128 + 2 * p.r, 1, 1);
    if (rc) return rc;
  }
  const int64_t ktot = (int64_t)taps * p.ctot;
  int rc = make_mat_map(&map_w, a->wpack, ktot, a->ncols_pad, ktot, cc, a->ncols_pad);
  if (rc) return rc;

  const size_t smem = ((p.w_bytes + 1023u) & ~1023u) + (size_t)p.nslots * p.slot_bytes + 1024;
  const unsigned grid = (unsigned)(a->n * p.wtiles * p.segs);
  static long long* trace_dev = nullptr;
  const bool tracing = getenv("SMSUT_BAND_TRACE") != nullptr;
  if (tracing) {
    if (!trace_dev) cudaMalloc(&trace_dev, (512 + 4096) * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, (512 + 4096) * sizeof(long long), stream);
    p.trace = trace_dev;
  }
  bool launched = false;
  const int nch = p.fast ? (a->ncols_pad >> 4) : 0;
#define BAND_CASE1(KS_, NS_, KK_, NCH_)                                                                           \
  if (!launched && p.ks == KS_ && a->nsrc == NS_ && (cc >> 4) == KK_ && nch == NCH_) {                             \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      SMSUT_CUDA_OK(cudaFuncSetAttribute(conv_band_kernel<KS_, NS_, KK_, NCH_>,                                    \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    launch_pdl(conv_band_kernel<KS_, NS_, KK_, NCH_>, grid, kBandThreads, smem, stream, maps[0], maps[1], map_w, p);       \
    launched = true;                                                                                               \
  }
#define BAND_CASE(KS_, NS_, KK_) \
  BAND_CASE1(KS_, NS_, KK_, 0) BAND_CASE1(KS_, NS_, KK_, 1) BAND_CASE1(KS_, NS_, KK_, 2) BAND_CASE1(KS_, NS_, KK_, 4)
  BAND_CASE(1, 1, 1) BAND_CASE(1, 1, 2) BAND_CASE(1, 1, 4) BAND_CASE(1, 2, 1) BAND_CASE(1, 2, 2) BAND_CASE(1, 2, 4)
  BAND_CASE(3, 1, 1) BAND_CASE(3, 1, 2) BAND_CASE(3, 1, 4) BAND_CASE(3, 2, 1) BAND_CASE(3, 2, 2) BAND_CASE(3, 2, 4)
  BAND_CASE(5, 1, 1) BAND_CASE(5, 1, 2) BAND_CASE(5, 1, 4)
#undef BAND_CASE1
#undef BAND_CASE
  if (!launched) return 0;
  count_launch();
  if (tracing) {
    static long long host[512 + 4096];
    cudaStreamSynchronize(stream);
    cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
    const long long t0 = host[0];
    fprintf(stderr, "band trace ks=%d nsrc=%d cc=%d ncols=%d w=%d rows=%d slots=%d grid=%u smem=%zu\n", p.ks, p.nsrc, cc,
            p.ncols_pad, p.w, p.rows_per_seg, p.nslots, grid, smem);
    fprintf(stderr, " row | mma: top  acc_free  rows_in  issued | epi: top  acc_full  done\n");
    for (int i = 0; i < 64 && i < p.rows_per_seg; ++i)
      fprintf(stderr, " %3d | %8lld %8lld %8lld %8lld | %8lld %8lld %8lld\n", i, host[i * 8] - t0, host[i * 8 + 1] - t0,
              host[i * 8 + 2] - t0, host[i * 8 + 3] - t0, host[i * 8 + 4] - t0, host[i * 8 + 5] - t0, host[i * 8 + 6] - t0);
    // per-CTA wall clock (globaltimer, ns): start / end relative to the earliest start, and the SM it ran on
    long long g0 = host[512];
    const unsigned nc = grid < 1024 ? grid : 1024;
    for (unsigned c = 0; c < nc; ++c) if (host[512 + c * 4] < g0) g0 = host[512 + c * 4];
    long long smax = 0, emax = 0, dsum = 0, dmax = 0;
    int per_sm[256] = {0};
    for (unsigned c = 0; c < nc; ++c) {
      const long long st = host[512 + c * 4] - g0, en = host[512 + c * 4 + 1] - g0;
      if (st > smax) smax = st;
      if (en > emax) emax = en;
      dsum += en - st;
      if (en - st > dmax) dmax = en - st;
      per_sm[host[512 + c * 4 + 2] & 255]++;
    }
    int hist[8] = {0};
    for (int i = 0; i < 256; ++i) hist[per_sm[i] < 7 ? per_sm[i] : 7]++;
    fprintf(stderr, "CTAs %u: last start %lld ns, last end %lld ns, mean duration %lld ns, max %lld ns; SMs with k CTAs:", nc,
            smax, emax, dsum / nc, dmax);
    for (int k = 1; k < 8; ++k) fprintf(stderr, " %d:%d", k, hist[k]);
    fprintf(stderr, "\n");
  }
  int st = launch_status("conv_band_kernel");
  return st ? st : 1;
}

}  // namespace smsut
