// "Band" weight gradient of the stride-1 'same' convolutions (every 1x1 / 3x3 / 5x5 layer with W % 16 == 0):
//   dW[co][ci][ty][tx] = sum_{n,h,w} dy[n,h,w,co] * x[n, h+ty-r, w+tx-r, ci]
// wgrad_tc_kernel re-reads a shifted x tile per tap (and the dy tile per column group): ncu showed it bound by
// L2->SM operand traffic.  Here one CTA owns (image, row segment, <=128-pixel column tile, <=64-channel chunk of x,
// <=64-channel chunk of dy) and walks down the rows.  Each x row (with halo) and dy row is fetched by TMA ONCE into a
// ring of row slots.  Per row, 16-pixel K chunk and vertical tap ty, ONE UMMA (two when the x chunk has 64
// channels) accumulates D_ty[(tx, ci)][co]: the horizontal taps are the M-blocks of the MN-major A operand whose
// leading-dimension byte offset is ONE PIXEL ROW (LBO = pitch), i.e. block tx is the same smem band started tx
// pixels later (the swizzle phase follows the absolute smem address: scripts/rowshift_probe.py).  A 16-pixel K chunk
// never crosses an image row, so any W % 16 == 0 works.  Accumulators stay in TMEM for the whole segment; the
// epilogue adds them to the fp32 OIHW gradient with atomics.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();
int make_act_map(CUtensorMap* m, const void* ptr, int c, int w, int h, int n, int64_t sw, int64_t sh, int64_t sn,
                 int cc, int tw, int th, int tn);

constexpr int kWbThreads = 192;
constexpr int kWbMaxSlots = 12;

struct WgradBandParams {
  int n, h, w, ks, r;
  int tw, kchunks;              // pixels per row tile (16..128), tw / 16
  int xcc, dcc;                 // channels of the x / dy chunk (16, 32 or 64)
  int xchunks, dchunks;         // chunks along Cin / Cout (grid dimensions)
  int groups, tx_per_group;     // UMMAs per (row, K chunk, ty): ceil(ks * xcc / 128); 128 / xcc
  int wtiles, segs, rows_per_seg;
  int nslots;
  uint32_t x_slot_bytes, d_slot_bytes, d_base_off;
  uint32_t x_pitch, d_pitch, x_layout, d_layout;
  uint32_t tmem_cols;
  int dbg_noepi;   // development: skip the atomics (SMSUT_WGRAD_NOEPI=1)
  int stack;       // N-stacked issue: one UMMA covers all vertical taps of an x row (see the MMA issuer)
  int multi;       // one MMA-issuing warp per vertical tap (see the kernel)
  int csize;       // > 1: (csize, 1, 1) thread-block cluster whose CTAs sum their accumulators through DSMEM before the atomics
  int m64;         // M = 64 UMMAs (16-channel x chunks, KS <= 3): 4 horizontal-tap blocks instead of 8 -> half the A bytes
  int tap_major;   // dw is the tap-major scratch [tap][cout_total][cin_total]: lanes = contiguous channels
  int cout_total;
  float* dw;
  long long* dw_q;         // fixed-point shadow of dw in deterministic mode (common.cuh), else nullptr
  int cout, cin_total, ci_off, c_valid, taps;
  long long* trace;        // development: clock stamps (SMSUT_WGRAD_TRACE=1), see wgrad_band_try
};

template <int KS>
__global__ void __launch_bounds__(kWbThreads, 1)
wgrad_band_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                  const __grid_constant__ WgradBandParams p) {
  pdl_trigger();   // the next kernel may be scheduled; this one waits for its predecessor after its own set-up
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full[kWbMaxSlots];
  __shared__ __align__(8) uint64_t x_empty[kWbMaxSlots];
  __shared__ __align__(8) uint64_t d_full[kWbMaxSlots];
  __shared__ __align__(8) uint64_t d_empty[kWbMaxSlots];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

  int b = blockIdx.x;
  const int seg = b % p.segs; b /= p.segs;
  const int wt = b % p.wtiles; b /= p.wtiles;
  const int n = b;
  const int xch = blockIdx.y, dch = blockIdx.z;
  const int w0 = wt * p.tw;
  const int h_begin = seg * p.rows_per_seg;
  int h_end = h_begin + p.rows_per_seg;
  if (h_end > p.h) h_end = p.h;
  const int nrows = h_end - h_begin;
  const int nrows_in = nrows + 2 * p.r;
  const int nslots = p.nslots;
  // MMA-issuing warps: one per vertical tap (warps 1 .. KS; warps 2.. are the epilogue warps, idle until the strip ends).
  // A UMMA of this kernel is tiny (N = 16..64 columns, K = 16 pixels) and one thread issues ~1 per 75 cycles
  // (descriptor arithmetic + the issue itself): with a single issuer the 8 * KS UMMAs of a 128-pixel row take longer
  // than the row's HBM time.  The vertical taps own disjoint accumulators, so each gets its own in-order issuer.
  const int ni = p.multi ? KS : 1;
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 1024 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[512 + blockIdx.x * 4 + 0] = (long long)t;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], (uint32_t)ni);      // every issuing warp releases every row slot once per use
      mbar_init(&d_full[s], 1);
      mbar_init(&d_empty[s], (uint32_t)ni);
    }
    mbar_init(&acc_full, (uint32_t)ni);
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
    // ===================== TMA producer: x rows (with halo) and dy rows, each exactly once =====================
    {
      const uint32_t el = elect_one_u32();   // warp-wide loop, elected issue (see common.cuh)
      int xs = 0, ds = 0;
      uint32_t xph = 0, dph = 0;
      for (int j = 0; j < nrows_in; ++j) {
        mbar_wait(&x_empty[xs], xph ^ 1u);
        if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && j < 64)
          p.trace[j * 8 + 0] = clock64();
        mbar_arrive_expect_tx_e(&x_full[xs], (uint32_t)(p.tw + 2 * p.r) * p.x_pitch, el);
        tma_load_4d_e(smem_al + (size_t)xs * p.x_slot_bytes, &map_x, &x_full[xs], xch * p.xcc, w0 - p.r, h_begin - p.r + j,
                      n, el);
        if (++xs == nslots) { xs = 0; xph ^= 1u; }
        if (j < nrows) {
          mbar_wait(&d_empty[ds], dph ^ 1u);
          mbar_arrive_expect_tx_e(&d_full[ds], (uint32_t)p.tw * p.d_pitch, el);
          tma_load_4d_e(smem_al + p.d_base_off + (size_t)ds * p.d_slot_bytes, &map_dy, &d_full[ds], dch * p.dcc, w0,
                        h_begin + j, n, el);
          if (++ds == nslots) { ds = 0; dph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || (p.multi && warp <= KS)) {
    // ===================== MMA issuer(s) =====================
    // A = x slot, MN-major, M = (tx, ci): M-blocks of xcc channels one pixel row apart (LBO = pitch);
    // B = dy slot, MN-major, N = dcc;  K = pixels (16 per UMMA)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(p.dcc >> 3) << 17) | (((p.m64 ? 64u : 128u) >> 4) << 24);
    const uint64_t a_hi = make_smem_desc(0, p.x_pitch, 8u * p.x_pitch, p.x_layout) & 0xFFFFFFFFFFFF0000ull;
    const uint64_t b_hi = make_smem_desc(0, 0, 8u * p.d_pitch, p.d_layout) & 0xFFFFFFFFFFFF0000ull;
    const uint32_t x_base = smem_base >> 4, d_base = (smem_base + p.d_base_off) >> 4;
    const uint32_t xslot_u = p.x_slot_bytes >> 4, dslot_u = p.d_slot_bytes >> 4;
    const uint32_t xk_u = (16u * p.x_pitch) >> 4, dk_u = (16u * p.d_pitch) >> 4;
    const uint32_t grp_u = ((uint32_t)p.tx_per_group * p.x_pitch) >> 4;   // next group of horizontal taps
    const int kchunks = p.kchunks, groups = p.groups;
    const uint32_t el = elect_one_u32();
    if (p.stack) {
      // N-stacked issue.  ncu / the UMMA model: one UMMA per (row, 16-pixel K chunk, vertical tap) with N = dcc = 16
      // reads its 4 KB A tile for 16 columns of work, and two co-resident CTAs then saturate the tensor pipe's smem
      // read port long before HBM.  Walk the X rows instead: x row j meets the dy rows j-KS+1 .. j, which sit in
      // consecutive slots of the dy ring, so ONE UMMA with B = those slots side by side (N-blocks LBO = slot stride)
      // accumulates every vertical tap at once: N = KS * dcc, a third of the instructions and A reads.  D column
      // block b holds vertical tap ty = KS-1-b.  The block an x row touches for the first time (rows 0..KS-1) gets its
      // own UMMA with accumulate = 0; a window that wraps around the ring is issued in two pieces.
      const uint64_t b_hi_st = make_smem_desc(0, p.d_slot_bytes, 8u * p.d_pitch, p.d_layout) & 0xFFFFFFFFFFFF0000ull;
      const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
      const int dcc = p.dcc;
      int xslot = 0, dslot_w = 0;
      uint32_t xphase = 0, dphase_w = 0;
      for (int j = 0; j < nrows_in; ++j) {
        mbar_wait(&x_full[xslot], xphase);
        if (j < nrows) {
          mbar_wait(&d_full[dslot_w], dphase_w);
          if (++dslot_w == nslots) { dslot_w = 0; dphase_w ^= 1u; }
        }
        tc_fence_after();
        const int blk_lo = KS - 1 - j > 0 ? KS - 1 - j : 0;
        const int blk_hi = nrows + KS - 2 - j < KS - 1 ? nrows + KS - 2 - j : KS - 1;
        if (blk_lo <= blk_hi) {
          const int i_lo = j - KS + 1 + blk_lo;                 // dy row of block blk_lo
          const bool fresh = j <= KS - 1;                     // block blk_lo (dy row 0) is touched for the first time
          const uint32_t a0 = x_base + (uint32_t)xslot * xslot_u;
          for (int k = 0; k < kchunks; ++k) {
            const uint64_t adesc = a_hi | (uint64_t)((a0 + k * xk_u) & 0x3FFFu);
            int b = blk_lo, i = i_lo;
            if (fresh) {
              const uint32_t bs = d_base + (uint32_t)(i % nslots) * dslot_u + k * dk_u;
              umma_bf16_e(tmem_base + (uint32_t)(b * dcc), adesc, b_hi | (uint64_t)(bs & 0x3FFFu),
                          idesc0 | ((uint32_t)(dcc >> 3) << 17), k != 0 ? 1u : 0u, el);
              ++b; ++i;
            }
            while (b <= blk_hi) {
              const int s0 = i % nslots;
              int nb = blk_hi - b + 1;
              if (nb > nslots - s0) nb = nslots - s0;          // up to the end of the ring
              const uint32_t bs = d_base + (uint32_t)s0 * dslot_u + k * dk_u;
              umma_bf16_e(tmem_base + (uint32_t)(b * dcc), adesc, b_hi_st | (uint64_t)(bs & 0x3FFFu),
                          idesc0 | ((uint32_t)((nb * dcc) >> 3) << 17), 1u, el);
              b += nb; i += nb;
            }
          }
        }
        umma_commit_e(&x_empty[xslot], el);
        const int i_rel = j - KS + 1;                        // dy row no later x row needs
        if (i_rel >= 0 && i_rel < nrows) umma_commit_e(&d_empty[i_rel % nslots], el);
        if (j == nrows_in - 1) umma_commit_e(&acc_full, el);
        if (++xslot == nslots) { xslot = 0; xphase ^= 1u; }
      }
    } else if (p.multi) {
      // issuer of vertical tap ty: dy row i meets x row i + ty (ring slot (i + ty) % nslots); it waits for exactly those
      // two rows, accumulates into its own TMEM columns and releases both slots.  x rows j < ty are never read by this
      // issuer: it arrives for them up front so that every slot sees `ni` arrivals per use.
      const int ty = warp - 1;
      if (el) for (int j = 0; j < ty; ++j) mbar_arrive(&x_empty[j]);
      int xslot = ty, dslot = 0;
      uint32_t xphase = 0, dphase = 0;
      for (int i = 0; i < nrows; ++i) {
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && i < 64 &&
                        (ty == 0 || ty == KS - 1);
        if (tr) p.trace[i * 8 + (ty == 0 ? 1 : 4)] = clock64();
        mbar_wait(&x_full[xslot], xphase);
        mbar_wait(&d_full[dslot], dphase);
        tc_fence_after();
        if (tr) p.trace[i * 8 + (ty == 0 ? 2 : 5)] = clock64();
        const uint32_t b0 = d_base + (uint32_t)dslot * dslot_u;
        const uint32_t a0 = x_base + (uint32_t)xslot * xslot_u;
        for (int g = 0; g < groups; ++g) {
          const uint32_t d_tmem = tmem_base + (uint32_t)((ty * groups + g) * p.dcc);
          const uint32_t ag = a0 + (uint32_t)g * grp_u;
#pragma unroll 4
          for (int k = 0; k < kchunks; ++k)
            umma_bf16_e(d_tmem, a_hi | (uint64_t)((ag + k * xk_u) & 0x3FFFu),
                        b_hi | (uint64_t)((b0 + k * dk_u) & 0x3FFFu), idesc, (i | k) != 0 ? 1u : 0u, el);
        }
        umma_commit_e(&x_empty[xslot], el);
        umma_commit_e(&d_empty[dslot], el);
        if (i == nrows - 1) umma_commit_e(&acc_full, el);
        if (tr) p.trace[i * 8 + (ty == 0 ? 3 : 6)] = clock64();
        if (++xslot == nslots) { xslot = 0; xphase ^= 1u; }
        if (++dslot == nslots) { dslot = 0; dphase ^= 1u; }
      }
    } else {
    int rows_ready = 0, ready_slot = 0, base_slot = 0, dslot = 0;
    uint32_t ready_phase = 0, dphase = 0;
    for (int i = 0; i < nrows; ++i) {
      while (rows_ready <= i + KS - 1) {
        mbar_wait(&x_full[ready_slot], ready_phase);
        ++rows_ready;
        if (++ready_slot == nslots) { ready_slot = 0; ready_phase ^= 1u; }
      }
      mbar_wait(&d_full[dslot], dphase);
      tc_fence_after();
      {
        const uint32_t b0 = d_base + (uint32_t)dslot * dslot_u;
#pragma unroll
        for (int ty = 0; ty < KS; ++ty) {
          int slot = base_slot + ty;
          if (slot >= nslots) slot -= nslots;
          const uint32_t a0 = x_base + (uint32_t)slot * xslot_u;
          for (int g = 0; g < groups; ++g) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((ty * groups + g) * p.dcc);
            const uint32_t ag = a0 + (uint32_t)g * grp_u;
            for (int k = 0; k < kchunks; ++k)
              umma_bf16_e(d_tmem, a_hi | (uint64_t)((ag + k * xk_u) & 0x3FFFu),
                          b_hi | (uint64_t)((b0 + k * dk_u) & 0x3FFFu), idesc, (i | k) != 0 ? 1u : 0u, el);
          }
        }
        umma_commit_e(&x_empty[base_slot], el);   // oldest x row of the window
        umma_commit_e(&d_empty[dslot], el);
        if (i == nrows - 1) umma_commit_e(&acc_full, el);
      }
      if (++base_slot == nslots) base_slot = 0;
      if (++dslot == nslots) { dslot = 0; dphase ^= 1u; }
    }
    }
  }
  if (warp >= 2) {
    // ===================== epilogue: D_{ty,g}[(tx, ci)][co] -> atomics into dW[co][ci_off + ci][ty][tx] ==========
    // M = 128: accumulator row (tx, ci) = TMEM lane.  M = 64 (cta_group::1): row m sits in lane (m & 15) + 32 * (m >> 4),
    // i.e. lanes 0..15 of warp quarter q hold horizontal tap q's 16 channels and lanes 16..31 hold nothing
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int txl = p.m64 ? (lane < 16 ? q : p.tx_per_group) : row / p.xcc;
    const int cil = p.m64 ? (lane & 15) : row - txl * p.xcc;
    const int ci = xch * p.xcc + cil;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    if (p.trace != nullptr && threadIdx.x == 64 && blockIdx.x < 1024 && blockIdx.y == 0 && blockIdx.z == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      p.trace[512 + blockIdx.x * 4 + 1] = (long long)t;       // main loop done, epilogue starts
    }
    const int nch = (KS * p.groups * p.dcc) >> 4;
    if (p.csize > 1) {
      // cluster reduction, phase A: this CTA's accumulators -> its own shared memory as S[column][128 rows] fp32 (the
      // row rings are dead: every MMA that read them completed before acc_full fired)
      float* S = reinterpret_cast<float*>(smem_al);
      for (int j = 0; j < nch; ++j) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) S[(size_t)(j * 16 + k) * 128 + row] = __uint_as_float(raw[k]);
      }
    } else
    for (int j = 0; j < nch; ++j) {
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
      tmem_ld_wait();
      const int col = j * 16;
      const int region = col / p.dcc, co0 = dch * p.dcc + (col - region * p.dcc);
      const int ty = p.stack ? KS - 1 - region : region / p.groups, g = p.stack ? 0 : region - (region / p.groups) * p.groups;
      const int tx = g * p.tx_per_group + txl;
      if (txl >= p.tx_per_group || tx >= KS || ci >= p.c_valid) continue;
      const int tap = ty * KS + tx;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int co = co0 + k;
        if (co < p.cout && !p.dbg_noepi) {
          float* dst = p.tap_major ? p.dw + ((size_t)tap * p.cout_total + co) * p.cin_total + p.ci_off + ci
                                   : p.dw + ((size_t)co * p.cin_total + p.ci_off + ci) * p.taps + tap;
          acc_add(dst, p.dw_q != nullptr ? p.dw_q + (dst - p.dw) : nullptr, __uint_as_float(raw[k]));
        }
      }
    }
  }

  if (p.csize > 1) {
    // Phase B.  Every CTA of the wave adds its accumulators to the SAME KS x KS x Cin x Cout addresses: ~300 CTAs x 2 304
    // fp32 atomics on 2 304 addresses for a 16 -> 16 layer, and the kernel is not complete until they have drained
    // through the L2 atomic units -- measured 16-28 us of a 50-100 us launch (SMSUT_WGRAD_NOEPI=1).  The CTAs of a
    // cluster therefore sum their partial tiles first: CTA `rank` finishes every csize-th column (its own and the
    // peers' copies, read with ld.shared::cluster) and issues the atomics for it -- csize times fewer of them.
    cluster_sync_all();
    if (warp >= 2) {
      const int ew = warp - 2;
      const int rank = (int)cluster_ctarank();
      const int ncols = KS * p.groups * p.dcc;
      for (int col = rank + p.csize * ew; col < ncols; col += p.csize * 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t local = smem_base + (uint32_t)((col * 128 + lane * 4) * 4);
        for (int c = 0; c < p.csize; ++c) {
          const float4 t = dsmem_ld_f4(dsmem_addr(local, (uint32_t)c));
          v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
        }
        const int region = col / p.dcc, co = blockIdx.z * p.dcc + (col - region * p.dcc);
        const int ty = region / p.groups, g = region - ty * p.groups;
        if (co >= p.cout || p.dbg_noepi) continue;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const int rw = lane * 4 + r4;
          const int txl = p.m64 ? ((rw & 31) < 16 ? (rw >> 5) : p.tx_per_group) : rw / p.xcc;
          const int cil = p.m64 ? (rw & 15) : rw - txl * p.xcc;
          const int ci = blockIdx.y * p.xcc + cil, tx = g * p.tx_per_group + txl;
          if (txl >= p.tx_per_group || tx >= KS || ci >= p.c_valid) continue;
          const int tap = ty * KS + tx;
          float* dst = p.tap_major ? p.dw + ((size_t)tap * p.cout_total + co) * p.cin_total + p.ci_off + ci
                                   : p.dw + ((size_t)co * p.cin_total + p.ci_off + ci) * p.taps + tap;
          acc_add(dst, p.dw_q != nullptr ? p.dw_q + (dst - p.dw) : nullptr, v[r4]);
        }
      }
    }
    cluster_sync_all();                 // no CTA leaves (or frees its shared memory) while a peer still reads it
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 1024 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[512 + blockIdx.x * 4 + 2] = (long long)t;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static uint32_t layout_for(int c) { return c == 64 ? 2u : (c == 32 ? 4u : 6u); }
static int chunk_of(int c) { int k = 64; while (c % k) k >>= 1; return k; }

// returns 1 if handled, 0 if not eligible, < 0 on error
int wgrad_band_try(const smsut_wgrad_tc_args* a, cudaStream_t stream) {
  if (a->kind != SMSUT_TC_CONV) return 0;
  if (!(a->ksize == 1 || a->ksize == 3 || a->ksize == 5)) return 0;
  if (a->w % 16 != 0 || a->x_c % 16 != 0 || a->dy_c % 16 != 0) return 0;
  {
    const char* e = getenv("SMSUT_NO_BAND");
    if (e && e[0] == '1') return 0;
    // The kernel handles every W % 16 == 0 layer, but measured on B200 it only beats wgrad_tc_kernel on the wide,
    // narrow-channel layers (launch list r1e: 110 us vs 42-64 us on the 32x32 / 64x64 levels): keep it to those
    // unless SMSUT_WGRAD_BAND_ALL=1 (development).
    e = getenv("SMSUT_WGRAD_BAND_ALL");
    if (!(e && e[0] == '1')) {
      if (a->w % 128 != 0) return 0;
      if (!(a->x_c == 16 || a->x_c == 32)) return 0;
      if (!(a->dy_c == 16 || a->dy_c == 32 || a->dy_c == 64)) return 0;
      if (a->ksize * a->x_c > 128) return 0;
    }
  }
  {
    // development (timing bound only, WRONG results): how much of the iteration is this kernel?  scripts/gpu_ab.sh
    const char* e = getenv("SMSUT_DBG_SKIP_WGRAD_BAND");
    if (e && e[0] == '1') return 1;
  }
  WgradBandParams p;
  memset(&p, 0, sizeof(p));
  { const char* e = getenv("SMSUT_WGRAD_NOEPI"); p.dbg_noepi = (e && e[0] == '1') ? 1 : 0; }
  p.n = a->n; p.h = a->h; p.w = a->w; p.ks = a->ksize; p.r = a->ksize / 2;
  // column tile: the largest multiple of 16 (<= 128) dividing W
  int tw = 128;
  while (a->w % tw) tw -= 16;
  p.tw = tw; p.kchunks = tw / 16; p.wtiles = a->w / tw;
  p.xcc = chunk_of(a->x_c); p.dcc = chunk_of(a->dy_c);
  p.xchunks = a->x_c / p.xcc; p.dchunks = a->dy_c / p.dcc;
  {
    // M = 64 for the 16-channel chunks of the 1x1 / 3x3 layers: the MN-major A operand of a UMMA is M x 16 pixels, and
    // with M = 128 = 8 horizontal-tap blocks only KS of them are taps -- the tensor pipe's shared-memory reads of A
    // (4 KB per UMMA, 24 UMMAs per 128-pixel row) bound the kernel.  M = 64 halves them.  SMSUT_WGRAD_M64=0: off.
    const char* e = getenv("SMSUT_WGRAD_M64");      // read per call: the parity tests run both settings in one process
    // Measured (B200, us per launch alone, M = 128 -> 64): 16->16 @256^2 53.5 -> 51.4, but 16->32 @128^2 35.7 -> 41.5
    // (a 1-SM UMMA with M = 64 exposes the shared-memory read latency of A): only with 16 dy channels, or =2 everywhere.
    p.m64 = (p.xcc == 16 && a->ksize * p.xcc <= 64 && !(e && e[0] == '0') && (p.dcc == 16 || (e && e[0] == '2'))) ? 1 : 0;
  }
  p.tx_per_group = (p.m64 ? 64 : 128) / p.xcc;
  p.groups = (a->ksize + p.tx_per_group - 1) / p.tx_per_group;
  p.x_pitch = (uint32_t)p.xcc * 2u; p.d_pitch = (uint32_t)p.dcc * 2u;
  p.x_layout = layout_for(p.xcc); p.d_layout = layout_for(p.dcc);
  // A reads up to (groups * tx_per_group - 1) pixel rows past the last K chunk: pad the slot accordingly
  p.x_slot_bytes = (((uint32_t)(tw + 2 * p.r + p.groups * p.tx_per_group) * p.x_pitch) + 1023u) & ~1023u;
  p.d_slot_bytes = (((uint32_t)tw * p.d_pitch) + 1023u) & ~1023u;
  p.nslots = 2 * p.r + 1 + 4;
  if (p.nslots > kWbMaxSlots) return 0;
  p.d_base_off = (uint32_t)p.nslots * p.x_slot_bytes;
  const size_t smem = (size_t)p.nslots * (p.x_slot_bytes + p.d_slot_bytes) + 2048;
  if (smem > 200u * 1024u) return 0;
  const int cols = a->ksize * p.groups * p.dcc;
  if (cols > 512) return 0;
  uint32_t tc = 32;
  while ((int)tc < cols) tc <<= 1;
  p.tmem_cols = tc;
  {
    const char* e = getenv("SMSUT_WGRAD_STACK");
    // opt-in: measured SLOWER on B200 (16->16 @ 256x256: 88 us stacked vs 57 us; whole step 13.3 vs 12.1 ms) -- an
    // MN-major B operand gathered from three row slots 4 KB apart evidently costs more than the saved instructions
    const char* em = getenv("SMSUT_WGRAD_MULTI");      // read per call (tests run both)
    p.multi = (a->ksize > 1 && !(em && em[0] == '0')) ? 1 : 0;
    p.stack = (p.groups == 1 && !p.m64 && !p.multi && a->ksize * p.dcc <= 256 && e && e[0] == '1') ? 1 : 0;
  }
  p.dw = a->dw;
  p.dw_q = det_shadow(a->dw);
  p.tap_major = a->dw_layout == 1 ? 1 : 0;
  p.cout_total = a->cout_total;
  p.cout = a->dy_c < a->cout_total ? a->dy_c : a->cout_total;
  p.cin_total = a->cin_total; p.ci_off = a->ci_off;
  p.c_valid = a->c_valid > 0 ? a->c_valid : a->x_c;
  p.taps = a->ksize * a->ksize;
  SMSUT_CHECK(a->dw != nullptr, -1, "null dw");

  // row segments: ONE balanced wave of co-resident CTAs (up to 2 per SM), at least 8 rows each (every segment
  // re-reads 2r halo rows and pays a full-accumulator atomic epilogue)
  const int sms = device_sm_count();
  const int base = a->n * p.wtiles * p.xchunks * p.dchunks;
  int per_sm = (int)((227u * 1024u) / (smem + 1024));
  static int per_sm_knob = -1;
  if (per_sm_knob < 0) {
    const char* e = getenv("SMSUT_WGRAD_BAND_PER_SM");
    // side-stream kernel: SM-time over latency.  Single issuer: 3 -> 12.36, 2 -> 12.14 ms / step.  With one issuing warp
    // per vertical tap the tensor pipe's shared-memory reads bound the main loop, so a second CTA per SM adds nothing to
    // it but doubles the set-up, the atomics of the epilogue and the shared memory held: 2 -> 10.02, 1 -> 9.85 ms / step
    per_sm_knob = e && atoi(e) > 0 ? atoi(e) : 0;
  }
  // default: one CTA per SM.  The single-issuer 1x1 layers are slower ALONE that way (ncu launch list: 19.6 -> 30.5 us per
  // launch), but two CTAs per SM for them cost the step 9.92 ms against 9.85 ms: what the iteration pays for a side-stream
  // kernel is the SM time and shared memory it takes from the other chains, not its own duration.
  const int per_sm_cap = per_sm_knob > 0 ? per_sm_knob : 1;
  if (per_sm > per_sm_cap) per_sm = per_sm_cap;
  if (per_sm < 1) per_sm = 1;
  int segs = (per_sm * sms) / base;
  if (segs < 1) segs = 1;
  int rows = (a->h + segs - 1) / segs;
  const int min_rows = a->ksize > 1 ? 8 : 4;
  if (rows < min_rows) rows = a->h < min_rows ? a->h : min_rows;
  {
    const char* e = getenv("SMSUT_BAND_ROWS");
    if (e && atoi(e) > 0) rows = atoi(e);
  }
  p.rows_per_seg = rows;
  p.segs = (a->h + rows - 1) / rows;

  CUtensorMap map_x, map_dy;
  int rc = make_act_map(&map_x, a->x, a->x_c, a->w, a->h, a->n, a->x_ld, (int64_t)a->x_ld * a->w,
                        (int64_t)a->x_ld * a->w * a->h, p.xcc, tw + 2 * p.r, 1, 1);
  if (rc) return rc;
  rc = make_act_map(&map_dy, a->dy, a->dy_c, a->w, a->h, a->n, a->dy_ld, (int64_t)a->dy_ld * a->w,
                    (int64_t)a->dy_ld * a->w * a->h, p.dcc, tw, 1, 1);
  if (rc) return rc;

  dim3 grid((unsigned)(a->n * p.wtiles * p.segs), (unsigned)p.xchunks, (unsigned)p.dchunks);
  // cluster reduction of the accumulators before the atomics (see the kernel): clusters of 4 (2) CTAs along x when the
  // staging tile fits the ring memory and the whole grid of clusters is still co-resident.  Opt-in
  // (SMSUT_WGRAD_CLUSTER=4 | 2): correct (tests) but MEASURED SLOWER on the step -- 9.845 ms off, 9.999 ms with
  // clusters of 2, 9.934 ms with 4; per launch 32->32 @128^2 gains (52.7 -> 39.2 us) while 16->32 loses (30.8 -> 39.0)
  // and 16->16 @256^2 is unchanged (53.2 -> 52.9): with one CTA per SM the atomics' drain is no longer what bounds the
  // launch, and cluster launches lose the freedom to place CTAs on any free SM beside the other streams' kernels.
  p.csize = 1;
  {
    static bool attrs_set = false;      // before the occupancy query: it must see the opt-in shared-memory size
    if (!attrs_set) {
      SMSUT_CUDA_OK(cudaFuncSetAttribute(wgrad_band_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      SMSUT_CUDA_OK(cudaFuncSetAttribute(wgrad_band_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      SMSUT_CUDA_OK(cudaFuncSetAttribute(wgrad_band_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attrs_set = true;
    }
    const char* e = getenv("SMSUT_WGRAD_CLUSTER");      // read per call (tests run both)
    const int want = e ? atoi(e) : 1;
    const size_t stage = (size_t)a->ksize * p.groups * p.dcc * 128 * sizeof(float);
    for (int c = want > 4 ? 4 : want; c > 1 && p.csize == 1; c >>= 1) {
      if (grid.x % (unsigned)c != 0 || stage + 1024 > smem) continue;
      int fit = 0;
      if (a->ksize == 1) fit = max_active_clusters_x(wgrad_band_kernel<1>, dim3(kWbThreads), smem, (unsigned)c);
      if (a->ksize == 3) fit = max_active_clusters_x(wgrad_band_kernel<3>, dim3(kWbThreads), smem, (unsigned)c);
      if (a->ksize == 5) fit = max_active_clusters_x(wgrad_band_kernel<5>, dim3(kWbThreads), smem, (unsigned)c);
      if ((long long)fit * c >= (long long)grid.x * grid.y * grid.z) p.csize = c;
    }
  }
  static long long* trace_dev = nullptr;
  const bool tracing = getenv("SMSUT_WGRAD_TRACE") != nullptr;
  if (tracing) {
    if (!trace_dev) cudaMalloc(&trace_dev, (512 + 4096) * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, (512 + 4096) * sizeof(long long), stream);
    p.trace = trace_dev;
  }
  bool launched = false;
#define WB_CASE(KS_)                                                                                               \
  if (!launched && a->ksize == KS_) {                                                                              \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      SMSUT_CUDA_OK(cudaFuncSetAttribute(wgrad_band_kernel<KS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    if (p.csize > 1)                                                                                               \
      launch_cluster_x(wgrad_band_kernel<KS_>, grid, kWbThreads, smem, stream, (unsigned)p.csize, map_x, map_dy, p);   \
    else                                                                                                           \
      launch_pdl(wgrad_band_kernel<KS_>, grid, kWbThreads, smem, stream, map_x, map_dy, p);                          \
    launched = true;                                                                                               \
  }
  WB_CASE(1) WB_CASE(3) WB_CASE(5)
#undef WB_CASE
  count_launch();
  if (tracing) {
    // development: per-row clock64 stamps of CTA 0 (producer; first and last issuer) and per-CTA globaltimer stamps
    static long long host[512 + 4096];
    cudaStreamSynchronize(stream);
    cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
    const long long t0 = host[0];
    fprintf(stderr, "wgrad band trace ks=%d xcc=%d dcc=%d w=%d rows=%d slots=%d grid=%u,%u,%u smem=%zu m64=%d multi=%d\n", p.ks,
            p.xcc, p.dcc, p.w, p.rows_per_seg, p.nslots, grid.x, grid.y, grid.z, smem, p.m64, p.multi);
    fprintf(stderr, " row | tma slot free | ty0: top  ready  issued | tyL: top  ready  issued\n");
    for (int i = 0; i < 64 && i < p.rows_per_seg + 2 * p.r; ++i)
      fprintf(stderr, " %3d | %8lld | %8lld %8lld %8lld | %8lld %8lld %8lld\n", i, host[i * 8] - t0, host[i * 8 + 1] - t0,
              host[i * 8 + 2] - t0, host[i * 8 + 3] - t0, host[i * 8 + 4] - t0, host[i * 8 + 5] - t0, host[i * 8 + 6] - t0);
    const unsigned nc = grid.x < 1024 ? grid.x : 1024;
    long long g0 = host[512];
    for (unsigned c = 0; c < nc; ++c) if (host[512 + c * 4] < g0) g0 = host[512 + c * 4];
    long long smax = 0, emax = 0, main_sum = 0, epi_sum = 0, epi_max = 0;
    for (unsigned c = 0; c < nc; ++c) {
      const long long st0 = host[512 + c * 4] - g0, ep = host[512 + c * 4 + 1] - g0, en = host[512 + c * 4 + 2] - g0;
      if (st0 > smax) smax = st0;
      if (en > emax) emax = en;
      main_sum += ep - st0;
      epi_sum += en - ep;
      if (en - ep > epi_max) epi_max = en - ep;
    }
    fprintf(stderr, "CTAs %u: last start %lld ns, last end %lld ns; mean set-up + main loop %lld ns, mean epilogue %lld ns (max %lld)\n",
            nc, smax, emax, main_sum / nc, epi_sum / nc, epi_max);
  }
  int st = launch_status("wgrad_band_kernel");
  return st ? st : 1;
}

}  // namespace smsut
