// Implicit-GEMM convolution on tcgen05 tensor cores for sm_100a.
//
//   D[m, col] = sum_k A[m, k] * B[col, k]
//     m   = output pixel (n, h, w)            -> 128-row tiles (tn x th x tw pixels)
//     col = output channel (or tap*Cout+co for the transposed conv)
//     k   = (tap, source, channel)            -> one pipeline step per (tap, source, <=64-channel chunk)
//
// A tiles are fetched by TMA straight from the NHWC bf16 activation tensor(s): a 4-D box
// (chunk, tw, th, tn) whose start coordinate is shifted by the tap offset, so zero padding is the
// TMA out-of-bounds fill and no im2col buffer ever exists.  B tiles come from the packed bf16 weight
// matrix.  Both land in the canonical K-major swizzled layout (swizzle = chunk bytes: 32/64/128) that
// tcgen05.mma consumes through shared-memory descriptors; accumulation is fp32 in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias/activation/accumulate -> NHWC store or 2x2 scatter).
//
// Reference semantics: network/blocks.py:10-16 (conv3x3/conv1x1), :41 (ConvTranspose2d k2 s2),
// :50 (torch.cat eliminated by the two-source K loop), network/ugan.py:295 (Linear).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

constexpr int kMaxSteps = 80;
constexpr int kTileM = 128;
constexpr int kThreads = 192;

struct KStep {
  int8_t map;   // which A tensor map
  int8_t dy, dx;
  int8_t pad;
  int16_t c0;   // channel coordinate inside that map
  int16_t pad2;
  int32_t k;    // K coordinate in the packed weight matrix
};

struct ConvTcParams {
  int n, h, w;             // GEMM-row space
  int tn, th, tw;          // tile (tn*th*tw == 128)
  int tiles_h, tiles_w;
  int cc;                  // channels per K step (16/32/64)
  int nsteps;
  int bn;                  // N tile
  int stages;
  uint32_t tmem_cols;
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t layout_type, sbo;
  // epilogue
  int mode;                // 0 plain, 1 transposed-conv scatter
  void* out0; int ld0, coff0;
  void* out1; int ld1, coff1, split;
  int cout_t;              // Cout of the transposed conv (mode 1)
  int ncols;               // valid columns
  const float* bias;
  int act; float slope;
  int accumulate, out_f32;
  int dbg_rowshift;        // experiment: load the A tile one pixel to the left and start the descriptor one row later
  float* stats;            // fused InstanceNorm statistics [n][2][ncols_pad] of the stored values (optional)
  long long* stats_q;      // its fixed-point shadow in deterministic mode (common.cuh), else nullptr
  int fast;                // plain epilogue: bf16, one destination, every column valid, no bias / act / accumulate
  int ksplit;              // > 1: the K steps are split over a (1,1,ksplit) cluster, partial tiles reduced through DSMEM
  int mc;                  // > 1: (mc,1,1) cluster of M tiles sharing one weight tile: CTA r loads rows [r, r+1) * bn / mc of
                           // every B tile and multicasts them to all mc CTAs (1 / mc of the weight traffic from L2)
  // vertical-tap sharing (3x3, tile inside one image): a K step is one (dx, source, chunk); its A box holds th + 2 rows
  // fetched ONCE, the three vertical taps are UMMA descriptors started vt_row16 (16-byte units) = one image row apart,
  // and the stage carries the three taps' weight tiles (K coordinates vt_kstride apart)
  int vt;
  uint32_t vt_row16;
  int vt_kstride;
  KStep steps[kMaxSteps];
};

template <typename T>
__device__ __forceinline__ void store_chunk(T* dst, const float* v, int nvalid, bool accumulate);

template <>
__device__ __forceinline__ void store_chunk<__nv_bfloat16>(__nv_bfloat16* dst, const float* v, int nvalid,
                                                            bool accumulate) {
  if (nvalid == 16) {
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = v[i];
    uint4* p = reinterpret_cast<uint4*>(dst);
    if (accumulate) {
      float o[8];
      uint4 q0 = p[0], q1 = p[1];
      unpack8(q0, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += o[i];
      unpack8(q1, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[8 + i] += o[i];
    }
    p[0] = pack8(f);
    p[1] = pack8(f + 8);
  } else {
    for (int i = 0; i < nvalid; ++i) {
      float x = v[i];
      if (accumulate) x += bf2f(dst[i]);
      dst[i] = f2bf(x);
    }
  }
}
template <>
__device__ __forceinline__ void store_chunk<float>(float* dst, const float* v, int nvalid, bool accumulate) {
  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 q = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (accumulate) {
        float4 o = p[i];
        q.x += o.x; q.y += o.y; q.z += o.z; q.w += o.w;
      }
      p[i] = q;
    }
  } else {
    for (int i = 0; i < nvalid; ++i) dst[i] = accumulate ? dst[i] + v[i] : v[i];
  }
}


// Plain-epilogue body for one 16-column chunk held in v[] (fp32): optional accumulate into / split of the bf16
// destination(s), two 128-bit stores, and the InstanceNorm statistics of the stored values (one shuffle
// reduce-scatter per chunk; a warp's 32 pixels belong to one image: host-checked).
__device__ __forceinline__ void tc_fast_chunk(const ConvTcParams& p, float (&v)[16], int col, size_t pix, bool row_ok,
                                              int lane, int n_w) {
  __nv_bfloat16* out0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + p.coff0;
  __nv_bfloat16* out1 = p.split > 0 ? reinterpret_cast<__nv_bfloat16*>(p.out1) + p.coff1 - p.split : out0;
  const int ld1 = p.split > 0 ? p.ld1 : p.ld0;
  uint4* dst = (p.split > 0 && col >= p.split) ? reinterpret_cast<uint4*>(out1 + pix * (size_t)ld1 + col)
                                               : reinterpret_cast<uint4*>(out0 + pix * (size_t)p.ld0 + col);
  if (p.accumulate && row_ok) {
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
  if (row_ok) {
    dst[0] = make_uint4(w32[0], w32[1], w32[2], w32[3]);
    dst[1] = make_uint4(w32[4], w32[5], w32[6], w32[7]);
  }
  if (p.stats != nullptr) {
    float vals[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float lo = row_ok ? __uint_as_float(w32[k] << 16) : 0.f;
      const float hi = row_ok ? __uint_as_float(w32[k] & 0xffff0000u) : 0.f;
      vals[2 * k] = lo; vals[2 * k + 1] = hi;
      vals[16 + 2 * k] = lo * lo; vals[16 + 2 * k + 1] = hi * hi;
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
    // lane l: l < 16 -> sum of channel l, else sum of squares of channel l - 16 (image of the warp's first row)
    if (n_w < p.n) acc_add_at(p.stats, p.stats_q, ((size_t)n_w * 2 + (lane >> 4)) * p.ncols + col + (lane & 15), vals[0]);
  }
}

// KSPLIT: compiled with the cluster split-K path (opt-in, see conv_tc_impl); the default instantiation carries none of it
template <bool KSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_ws,
               const __grid_constant__ ConvTcParams p) {
  pdl_trigger();   // the next kernel may be scheduled; this one waits for its predecessor after its own set-up
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // 1024-byte aligned stage ring (required by the 128B swizzle atoms)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // tile coordinates
  int tile = blockIdx.x;
  const int tw_i = tile % p.tiles_w; tile /= p.tiles_w;
  const int th_i = tile % p.tiles_h; tile /= p.tiles_h;
  const int n0 = tile * p.tn;
  const int h0 = th_i * p.th;
  const int w0 = tw_i * p.tw;
  const int col0 = blockIdx.y * p.bn;
  // split-K over the cluster: CTA `kr` of the (1, 1, ksplit) cluster owns K steps [ks_begin, ks_end)
  const bool split = KSPLIT && p.ksplit > 1;
  const bool mcast = KSPLIT && p.mc > 1;
  const uint32_t mc_rank = mcast ? cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << (mcast ? p.mc : 1)) - 1u);
  const int kr = split ? (int)blockIdx.z : 0;
  const int ks_begin = split ? (kr * p.nsteps) / p.ksplit : 0;
  const int ks_end = split ? ((kr + 1) * p.nsteps) / p.ksplit : p.nsteps;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], mcast ? (uint32_t)p.mc : 1u);      // multicast: every CTA of the cluster releases every stage
    }
    mbar_init(&tmem_full_bar, 1);
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
  if (mcast) cluster_sync_all();      // the peers' barriers exist before anything is multicast to / arrives on them
  pdl_wait();      // barriers, TMEM and descriptors are ready: now the predecessor's results are needed

  if (warp == 0) {
    // ===================== TMA producer (warp-wide loop, elected issue) =====================
    {
      const uint32_t el = elect_one_u32();
      int stage = 0;
      uint32_t phase = 0;
      uint8_t* ring = smem_raw + (smem_base - smem_u32(smem_raw));
      for (int i = ks_begin; i < ks_end; ++i) {
        const KStep st = p.steps[i];
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx_e(&full_bar[stage], p.a_bytes + (p.vt ? 3u : 1u) * p.b_bytes, el);
        void* a_dst = ring + (size_t)stage * p.stage_bytes;
        uint8_t* b_dst = (uint8_t*)a_dst + p.a_bytes;
        const CUtensorMap* m = st.map == 0 ? &map_a0 : (st.map == 1 ? &map_a1 : (st.map == 2 ? &map_a2 : &map_a3));
        tma_load_4d_e(a_dst, m, &full_bar[stage], st.c0, w0 + st.dx - (p.dbg_rowshift ? 1 : 0), h0 + st.dy, n0, el);
        if (mcast) {
          // this CTA's slice of the weight tile(s) -> the same place in every CTA of the cluster; the peers' slices
          // arrive the same way and complete this stage's barrier (expect_tx counts the WHOLE tile)
          const uint32_t sl_rows = (uint32_t)p.bn / (uint32_t)p.mc;
          const uint32_t sl_off = mc_rank * (p.b_bytes / (uint32_t)p.mc);
          const int row0 = col0 + (int)(mc_rank * sl_rows);
          tma_load_2d_mc_e(b_dst + sl_off, &map_ws, &full_bar[stage], st.k, row0, mc_mask, el);
          if (p.vt) {
            tma_load_2d_mc_e(b_dst + p.b_bytes + sl_off, &map_ws, &full_bar[stage], st.k + p.vt_kstride, row0, mc_mask, el);
            tma_load_2d_mc_e(b_dst + 2 * p.b_bytes + sl_off, &map_ws, &full_bar[stage], st.k + 2 * p.vt_kstride, row0,
                             mc_mask, el);
          }
        } else {
        tma_load_2d_e(b_dst, &map_w, &full_bar[stage], st.k, col0, el);
        if (p.vt) {
          tma_load_2d_e(b_dst + p.b_bytes, &map_w, &full_bar[stage], st.k + p.vt_kstride, col0, el);
          tma_load_2d_e(b_dst + 2 * p.b_bytes, &map_w, &full_bar[stage], st.k + 2 * p.vt_kstride, col0, el);
        }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: fp32 accum, bf16 x bf16, both K-major, N = bn, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((128u >> 4) << 24);
    // descriptors: constant hi word, lo word advanced by adds (16-byte units); the issuing thread's instruction
    // stream is the critical path of the deep layers (36-72 K steps per tile), so keep it minimal
    const uint64_t desc_hi = make_smem_desc(0, 16, p.sbo, p.layout_type) & 0xFFFFFFFF00000000ull;
    const uint32_t base_lo = (smem_base >> 4) | (1u << 16);
    const uint32_t stage_u = p.stage_bytes >> 4, a_u = p.a_bytes >> 4;
    const int kk = p.cc >> 4;
    const uint32_t el = elect_one_u32();
    int stage = 0;
    uint32_t phase = 0;
    for (int i = ks_begin; i < ks_end; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      uint32_t a_lo = base_lo + (uint32_t)stage * stage_u;
      if (p.dbg_rowshift) a_lo += (uint32_t)(p.cc * 2) >> 4;   // experiment: start one row later
      const uint32_t b_lo = base_lo + (uint32_t)stage * stage_u + a_u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < kk)
          umma_bf16_e(tmem_base, desc_hi | (uint64_t)(a_lo + 2u * k), desc_hi | (uint64_t)(b_lo + 2u * k), idesc,
                      (i != ks_begin || k != 0) ? 1u : 0u, el);
      }
      if (p.vt) {
        // the other two vertical taps of this (dx, chunk): same stage, A one / two image rows further down
        const uint32_t b_u = p.b_bytes >> 4;
#pragma unroll
        for (int j = 1; j < 3; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < kk)
              umma_bf16_e(tmem_base, desc_hi | (uint64_t)(a_lo + (uint32_t)j * p.vt_row16 + 2u * k),
                          desc_hi | (uint64_t)(b_lo + (uint32_t)j * b_u + 2u * k), idesc, 1u, el);
          }
        }
      }
      if (mcast) umma_commit_mc_e(&empty_bar[stage], mc_mask, el);      // the stage is free in EVERY CTA only when all released it
      else umma_commit_e(&empty_bar[stage], el);
      if (i == ks_end - 1) umma_commit_e(&tmem_full_bar, el);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int wl = r % p.tw;
    const int hl = (r / p.tw) % p.th;
    const int nl = r / (p.tw * p.th);
    const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
    const bool row_ok = (n < p.n) && (h < p.h) && (w < p.w);

    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();

    const int nchunks = p.bn >> 4;
    if (split) {
      // split-K, phase A: this CTA's partial tile -> its own shared memory as [chunk][row][16] fp32 (the stage ring
      // is free: every MMA that read it has completed before tmem_full fired)
      float4* stg = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)));
      for (int j = 0; j < nchunks; ++j) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          stg[(size_t)(j * 128 + r) * 4 + k] =
              make_float4(__uint_as_float(raw[4 * k]), __uint_as_float(raw[4 * k + 1]), __uint_as_float(raw[4 * k + 2]),
                          __uint_as_float(raw[4 * k + 3]));
      }
    } else if (p.fast) {
      const size_t pix = ((size_t)n * p.h + h) * p.w + w;
      const int n_w = n0 + (q * 32) / (p.tw * p.th);
      for (int j = 0; j < nchunks; ++j) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
        tc_fast_chunk(p, v, col0 + j * 16, pix, row_ok, lane, n_w);
      }
    } else
    for (int j = 0; j < nchunks; ++j) {
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
      tmem_ld_wait();
      const int col = col0 + j * 16;
      if (!row_ok || col >= p.ncols) continue;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
      int nvalid = p.ncols - col;
      if (nvalid > 16) nvalid = 16;
      if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < nvalid) v[i] += p.bias[col + i];
      }
      if (p.act == SMSUT_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (p.act == SMSUT_ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = lrelu(v[i], p.slope);
      }
      // destination
      void* base;
      int ld, coff, c;
      size_t pix;
      if (p.mode == 0) {
        pix = ((size_t)n * p.h + h) * p.w + w;
        if (p.split > 0 && col >= p.split) {
          base = p.out1; ld = p.ld1; coff = p.coff1; c = col - p.split;
        } else {
          base = p.out0; ld = p.ld0; coff = p.coff0; c = col;
          if (p.split > 0 && col + nvalid > p.split) nvalid = p.split - col;
        }
      } else {
        const int t = col / p.cout_t;
        c = col - t * p.cout_t;
        const int ty = t >> 1, tx = t & 1;
        pix = ((size_t)n * (2 * p.h) + (2 * h + ty)) * (size_t)(2 * p.w) + (2 * w + tx);
        base = p.out0; ld = p.ld0; coff = p.coff0;
      }
      const size_t off = pix * (size_t)ld + coff + c;
      if (p.out_f32)
        store_chunk<float>(reinterpret_cast<float*>(base) + off, v, nvalid, p.accumulate != 0);
      else
        store_chunk<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(base) + off, v, nvalid, p.accumulate != 0);
    }
  }

  if (split) {
    // split-K, phase B: reduce-scatter of the partial tiles through distributed shared memory.  CTA kr of the cluster
    // finishes the chunks j = kr, kr + ksplit, ...: it sums the ksplit partials of those 16 columns (its own and the
    // peers', read with ld.shared::cluster) and runs the plain epilogue on them.
    cluster_sync_all();                 // every CTA's partial tile is in its shared memory
    if (warp >= 2) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int wl = r % p.tw, hl = (r / p.tw) % p.th, nl = r / (p.tw * p.th);
      const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
      const bool row_ok = (n < p.n) && (h < p.h) && (w < p.w);
      const size_t pix = ((size_t)n * p.h + h) * p.w + w;
      const int n_w = n0 + (q * 32) / (p.tw * p.th);
      const int nchunks = p.bn >> 4;
      for (int j = kr; j < nchunks; j += p.ksplit) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0.f;
        const uint32_t local = smem_base + (uint32_t)((j * 128 + r) * 64);
        for (int c = 0; c < p.ksplit; ++c) {
          const uint32_t ra = dsmem_addr(local, (uint32_t)c);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 t = dsmem_ld_f4(ra + 16u * k);
            v[4 * k] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
          }
        }
        tc_fast_chunk(p, v, col0 + j * 16, pix, row_ok, lane, n_w);
      }
    }
    cluster_sync_all();                 // no CTA leaves (or frees its shared memory) while a peer still reads it
  }
  if (mcast) cluster_sync_all();       // no CTA leaves while a peer's commit may still arrive on its barriers
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(sym);
  return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 4-D activation map: dims (C, W, H, N) with element strides (1, sw, sh, sn); box (cc, tw, th, tn).
int make_act_map(CUtensorMap* m, const void* ptr, int c, int w, int h, int n, int64_t sw, int64_t sh, int64_t sn,
                 int cc, int tw, int th, int tn) {
  PFN_encodeTiled enc = get_encode_tiled();
  SMSUT_CHECK(enc != nullptr, -2, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  SMSUT_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, -3, "activation pointer must be 16-byte aligned");
  SMSUT_CHECK((sw * 2) % 16 == 0 && (sh * 2) % 16 == 0 && (sn * 2) % 16 == 0, -3,
              "activation strides must be multiples of 16 bytes (sw=%lld)", (long long)sw);
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)cc, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(cc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SMSUT_CHECK(r == CUDA_SUCCESS, -4,
              "cuTensorMapEncodeTiled(activation) failed: %d (c=%d w=%d h=%d n=%d box=%d,%d,%d,%d)", (int)r, c, w, h,
              n, cc, tw, th, tn);
  return 0;
}

// 2-D matrix map: rows x kdim (kdim contiguous); box (cc, rows_box).
int make_mat_map(CUtensorMap* m, const void* ptr, int64_t kdim, int64_t rows, int64_t ld, int cc, int rows_box) {
  PFN_encodeTiled enc = get_encode_tiled();
  SMSUT_CHECK(enc != nullptr, -2, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  SMSUT_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, -3, "matrix pointer must be 16-byte aligned");
  SMSUT_CHECK((ld * 2) % 16 == 0, -3, "matrix row pitch must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)cc, (cuuint32_t)rows_box};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(cc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SMSUT_CHECK(r == CUDA_SUCCESS, -4, "cuTensorMapEncodeTiled(matrix) failed: %d (k=%lld rows=%lld box=%d,%d)", (int)r,
              (long long)kdim, (long long)rows, cc, rows_box);
  return 0;
}

static int pow2_floor_dividing(int x, int cap) {
  int t = 1;
  while (t * 2 <= cap && x % (t * 2) == 0) t *= 2;
  return t;
}

// Choose the 128-pixel tile (tn, th, tw) for an (n, h, w) GEMM-row space.
int choose_tile(int n, int h, int w, int* tn, int* th, int* tw) {
  int a = pow2_floor_dividing(w, kTileM);
  int b = pow2_floor_dividing(h, kTileM / a);
  int c = kTileM / (a * b);
  (void)n;
  *tw = a; *th = b; *tn = c;
  SMSUT_CHECK(c <= 256, -5, "unsupported spatial shape %dx%d for 128-pixel tiling", h, w);
  return 0;
}

void count_launch();
int conv_band_try(const smsut_conv_tc_args* a, cudaStream_t stream);
bool conv_band_fuses_stats(const smsut_conv_tc_args* a);
bool conv_band_eligible_c(const smsut_conv_tc_args* a);

// plain epilogue: bf16 output(s), every column valid, no bias / activation (accumulate and a 16-aligned split allowed)
static bool conv_tc_plain_epilogue(const smsut_conv_tc_args* a) {
  const bool split_ok = a->out1 == nullptr || (a->split % 16 == 0 && a->out1_ld % 8 == 0 && a->out1_coff % 8 == 0);
  return (a->kind == SMSUT_TC_CONV || a->kind == SMSUT_TC_CONVT_DGRAD) && !a->out_f32 && a->bias == nullptr &&
         a->act == SMSUT_ACT_NONE && split_ok && a->ncols == a->ncols_pad && a->ncols_pad % 16 == 0 &&
         a->out0_ld % 8 == 0 && a->out0_coff % 8 == 0;
}

// conv_tc_kernel fuses the statistics when the epilogue is the plain one and the 32 pixels of an epilogue warp belong
// to one image (tile = part of one image, or whole images of a multiple of 32 pixels)
static bool conv_tc_fuses_stats(const smsut_conv_tc_args* a) {
  if (!conv_tc_plain_epilogue(a) || a->accumulate || a->out1 != nullptr || a->kind != SMSUT_TC_CONV) return false;
  int tn, th, tw;
  if (choose_tile(a->n, a->h, a->w, &tn, &th, &tw)) return false;
  if (a->h % th != 0 || a->w % tw != 0) return false;
  return tn == 1 || (th * tw) % 32 == 0;
}

// N tile: the widest of 256..16 dividing the column count that still gives ~a wave of CTAs (or <= 64)
static int default_bn(int m_tiles, int ncols_pad) {
  const int cands[5] = {256, 128, 64, 32, 16};
  int bn = 16;
  for (int i = 0; i < 5; ++i) {
    const int c = cands[i];
    if (c > ncols_pad || ncols_pad % c != 0) continue;
    bn = c;
    if ((int64_t)m_tiles * (ncols_pad / c) >= 120 || c <= 64) break;
  }
  return bn;
}

static int conv_tc_impl(const smsut_conv_tc_args* a, cudaStream_t stream) {
  SMSUT_CHECK(a != nullptr, -1, "null args");
  SMSUT_CHECK(a->nsrc == 1 || a->nsrc == 2, -1, "nsrc must be 1 or 2");
  SMSUT_CHECK(a->n > 0 && a->h > 0 && a->w > 0, -1, "bad dims");
  const int ctot = a->src_c[0] + (a->nsrc == 2 ? a->src_c[1] : 0);
  for (int s = 0; s < a->nsrc; ++s)
    SMSUT_CHECK(a->src_c[s] % 16 == 0 && a->src_c[s] > 0 && a->src_ld[s] >= a->src_c[s] && a->src_ld[s] % 8 == 0, -1,
                "source %d channels (%d, ld %d) must be a positive multiple of 16", s, a->src_c[s], a->src_ld[s]);
  SMSUT_CHECK(a->ncols_pad % 16 == 0 && a->ncols <= a->ncols_pad && a->ncols > 0, -1, "bad ncols/ncols_pad");

  if (a->stats != nullptr)
    SMSUT_CHECK(conv_band_fuses_stats(a) || conv_tc_fuses_stats(a), -1,
                "conv_tc: stats requested but smsut_conv_tc_fuses_stats() is 0 for this shape");
  // wide, narrow-channel layers: the band kernel (each input row fetched once, taps by descriptor arithmetic)
  if (a->bn == 0 && a->ncols_pad % 16 == 0 && a->src_c[0] % 16 == 0) {
    const int rb = conv_band_try(a, stream);
    if (rb != 0) return rb < 0 ? rb : 0;
  }

  static bool attr_set = false;
  if (!attr_set) {
    SMSUT_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    SMSUT_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.n = a->n; p.h = a->h; p.w = a->w;
  int rc = choose_tile(a->n, a->h, a->w, &p.tn, &p.th, &p.tw);
  if (rc) return rc;
  p.tiles_h = a->h / p.th;
  p.tiles_w = a->w / p.tw;
  SMSUT_CHECK(p.tiles_h * p.th == a->h && p.tiles_w * p.tw == a->w, -5, "spatial dims %dx%d not tileable", a->h, a->w);
  const int tiles_n = (a->n + p.tn - 1) / p.tn;
  const int m_tiles = tiles_n * p.tiles_h * p.tiles_w;

  // channel chunk = largest of 64/32/16 dividing every source's channel count
  int cc = 64;
  {
    const char* e = getenv("SMSUT_TC_MAXCC");      // development: 32 / 16 = shorter K steps, more pipeline stages
    if (e && (atoi(e) == 32 || atoi(e) == 16) && (ctot / atoi(e)) * 9 <= kMaxSteps) cc = atoi(e);
  }
  for (int s = 0; s < a->nsrc; ++s)
    while (a->src_c[s] % cc != 0) cc >>= 1;
  p.cc = cc;
  p.layout_type = cc == 64 ? 2u : (cc == 32 ? 4u : 6u);
  p.sbo = 8u * (uint32_t)cc * 2u;

  // K steps and A tensor maps
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  int ns = 0;
  if (a->kind == SMSUT_TC_CONV || a->kind == SMSUT_TC_CONVT_FWD) {
    const int ks = a->kind == SMSUT_TC_CONV ? a->ksize : 1;
    SMSUT_CHECK(ks == 1 || ks == 3 || ks == 5, -1, "ksize must be 1, 3 or 5");
    // vertical-tap sharing: 3x3, the 128-pixel tile lies inside one image and its rows are whole 8-pixel groups, so
    // that "one image row further down" is a whole number of swizzle atoms (SMSUT_TC_VT=0 switches it off)
    // Measured on B200 (scripts/conv_classes.py, us per launch at 16 slices, off -> on): 256->256 @16x16 19.0 -> 15.0,
    // 128->256 13.5 -> 11.4, 64->128 @32x32 13.1 -> 11.6, 256->128 21.3 -> 19.6; but 64->64 @64x64 19.5 -> 25.5 (th = 2:
    // the halo doubles the rows fetched and the 56 KB stage leaves no room for the second CTA per SM) and a 256-wide
    // N tile leaves a single stage.  Hence: th >= 4 and at least two stages.  SMSUT_TC_VT=0 off, =2 wherever legal.
    const char* vt_env = getenv("SMSUT_TC_VT");      // read per call: the parity tests run both settings in one process
    const int vt_knob = vt_env ? atoi(vt_env) : 1;
    p.vt = (vt_knob && a->kind == SMSUT_TC_CONV && ks == 3 && p.tn == 1 && p.tw % 8 == 0 && p.th + 2 <= 256 &&
            p.th + 2 <= a->h) ? 1 : 0;      // the halo box never exceeds the image height
    if (p.vt && vt_knob != 2) {
      const int bn0 = a->bn > 0 ? a->bn : default_bn(m_tiles, a->ncols_pad);
      const size_t stage = (size_t)(p.th + 2) * p.tw * cc * 2 + 3u * (size_t)bn0 * cc * 2;
      if (p.th < 4 || 2 * stage > 200u * 1024u) p.vt = 0;
    }
    for (int s = 0; s < a->nsrc; ++s) {
      rc = make_act_map(&maps[s], a->src[s], a->src_c[s], a->w, a->h, a->n, a->src_ld[s],
                        (int64_t)a->src_ld[s] * a->w, (int64_t)a->src_ld[s] * a->w * a->h, cc, p.tw,
                        p.vt ? p.th + 2 : p.th, p.tn);
      if (rc) return rc;
    }
    const int r = ks / 2;
    int t = 0;
    if (p.vt) {
      p.vt_row16 = (uint32_t)(p.tw * cc * 2) >> 4;
      p.vt_kstride = 3 * ctot;
      for (int dx = -1; dx <= 1; ++dx) {
        int coff = 0;
        for (int s = 0; s < a->nsrc; ++s) {
          for (int c0 = 0; c0 < a->src_c[s]; c0 += cc) {
            SMSUT_CHECK(ns < kMaxSteps, -6, "too many K steps");
            KStep& st = p.steps[ns++];
            st.map = (int8_t)s; st.dy = -1; st.dx = (int8_t)dx; st.c0 = (int16_t)c0;
            st.k = (dx + 1) * ctot + coff + c0;      // tap (dy = -1, dx); taps (0, dx), (+1, dx) are vt_kstride apart
          }
          coff += a->src_c[s];
        }
      }
    } else
    for (int dy = -r; dy <= r; ++dy)
      for (int dx = -r; dx <= r; ++dx, ++t) {
        int coff = 0;
        for (int s = 0; s < a->nsrc; ++s) {
          for (int c0 = 0; c0 < a->src_c[s]; c0 += cc) {
            SMSUT_CHECK(ns < kMaxSteps, -6, "too many K steps");
            KStep& st = p.steps[ns++];
            st.map = (int8_t)s; st.dy = (int8_t)dy; st.dx = (int8_t)dx; st.c0 = (int16_t)c0;
            st.k = t * ctot + coff + c0;
          }
          coff += a->src_c[s];
        }
      }
  } else if (a->kind == SMSUT_TC_CONVT_DGRAD) {
    SMSUT_CHECK(a->nsrc == 1, -1, "convT dgrad takes one source");
    // four strided planes dy[:, ty::2, tx::2, :]
    const int64_t ld = a->src_ld[0];
    const int64_t W2 = 2 * (int64_t)a->w, H2 = 2 * (int64_t)a->h;
    for (int t = 0; t < 4; ++t) {
      const int ty = t >> 1, tx = t & 1;
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a->src[0]) + ((int64_t)ty * W2 + tx) * ld;
      rc = make_act_map(&maps[t], base, a->src_c[0], a->w, a->h, a->n, 2 * ld, 2 * W2 * ld, H2 * W2 * ld, cc, p.tw,
                        p.th, p.tn);
      if (rc) return rc;
      for (int c0 = 0; c0 < a->src_c[0]; c0 += cc) {
        SMSUT_CHECK(ns < kMaxSteps, -6, "too many K steps");
        KStep& st = p.steps[ns++];
        st.map = (int8_t)t; st.dy = 0; st.dx = 0; st.c0 = (int16_t)c0;
        st.k = t * ctot + c0;
      }
    }
  } else {
    SMSUT_CHECK(false, -1, "unknown kind %d", a->kind);
  }
  p.nsteps = ns;
  const int64_t ktot = (a->kind == SMSUT_TC_CONV ? (int64_t)a->ksize * a->ksize : (a->kind == SMSUT_TC_CONVT_DGRAD ? 4 : 1)) * ctot;

  // N tile
  int bn = a->bn;
  if (bn <= 0) bn = default_bn(m_tiles, a->ncols_pad);
  // Deep layers: a handful of output tiles (32 at 16x16) each with a long K loop (36-72 steps of a narrow N tile).
  // Optional (SMSUT_TC_KSPLIT=2|4): take the widest N tile instead and split K over a (1, 1, ksplit) thread-block
  // cluster; the partial tiles meet in distributed shared memory (reduce-scatter: CTA k finishes every ksplit-th
  // 16-column chunk).  Only with the plain epilogue, and only when the whole cluster grid still fits one wave.
  // MEASURED on B200 and OFF by default: 256->256 @ 16x16 takes 24.6 us split 4-way against 18.4 us un-split
  // (128->256: 22.6 vs 13.0): the 128 KB reduce-scatter through DSMEM (~20 B/cycle/SM) costs more than the shorter
  // K loop saves.
  int ksplit = 1;
  if (a->bn <= 0 && conv_tc_plain_epilogue(a)) {
    static int knob = -1;
    if (knob < 0) {
      const char* e = getenv("SMSUT_TC_KSPLIT");
      knob = e ? atoi(e) : 0;
    }
    const int sms = device_sm_count();
    if (knob > 1 && (int64_t)m_tiles * (a->ncols_pad / bn) <= sms) {
      const int cands[4] = {256, 128, 64, 32};
      for (int i = 0; i < 4 && ksplit == 1; ++i) {
        const int c = cands[i];
        if (c > a->ncols_pad || a->ncols_pad % c != 0) continue;
        const int ctas = m_tiles * (a->ncols_pad / c);
        for (int k = knob > 4 ? 4 : knob; k > 1; k >>= 1) {
          if (ctas * k <= sms && ns >= 3 * k && (c >> 4) >= k) {
            bn = c;
            ksplit = k;
            break;
          }
        }
      }
    }
  }
  SMSUT_CHECK(bn % 16 == 0 && bn >= 16 && bn <= 256 && a->ncols_pad % bn == 0, -1, "bad N tile %d for %d columns", bn,
              a->ncols_pad);
  p.bn = bn;
  uint32_t tc = 32;
  while ((int)tc < bn) tc <<= 1;
  p.tmem_cols = tc;

  CUtensorMap map_w;
  rc = make_mat_map(&map_w, a->wpack, ktot, a->ncols_pad, ktot, cc, bn);
  if (rc) return rc;

  p.a_bytes = (uint32_t)(p.vt ? (p.th + 2) * p.tw : kTileM) * cc * 2;
  p.b_bytes = (uint32_t)bn * cc * 2;
  p.stage_bytes = (p.a_bytes + (p.vt ? 3u : 1u) * p.b_bytes + 1023u) & ~1023u;
  int stages = (int)((200u * 1024u) / p.stage_bytes);
  if (stages > 6) stages = 6;
  {
    // more tiles than SMs: leave room for a second (third) CTA per SM, whose main loop then covers this one's
    // prologue and epilogue, as long as >= 3 stages remain
    static int per_sm_knob = -1;
    if (per_sm_knob < 0) {
      const char* e = getenv("SMSUT_TC_CTAS_PER_SM");
      per_sm_knob = e ? atoi(e) : 2;
    }
    const long long ctas = (long long)m_tiles * (a->ncols_pad / bn);
    if (per_sm_knob > 1 && ctas > device_sm_count()) {
      int s2 = (int)((220u * 1024u / per_sm_knob - 2048u) / p.stage_bytes);
      if (s2 >= 3 && s2 < stages) stages = s2;
    }
  }
  if (stages > ns) stages = ns;
  if (stages < 1) stages = 1;
  p.stages = stages;
  size_t smem = (size_t)stages * p.stage_bytes;
  if (ksplit > 1 && smem < (size_t)bn * 512u) smem = (size_t)bn * 512u;      // split-K staging [chunk][128][16] fp32
  smem += 1024;

  p.mode = a->kind == SMSUT_TC_CONVT_FWD ? 1 : 0;
  p.out0 = a->out0; p.ld0 = a->out0_ld; p.coff0 = a->out0_coff;
  p.out1 = a->out1; p.ld1 = a->out1_ld; p.coff1 = a->out1_coff; p.split = a->out1 ? a->split : 0;
  p.cout_t = a->kind == SMSUT_TC_CONVT_FWD ? a->ncols / 4 : 0;
  if (p.mode == 1) SMSUT_CHECK(p.cout_t % 16 == 0, -1, "convT Cout must be a multiple of 16");
  if (p.split > 0) SMSUT_CHECK(p.split % 16 == 0, -1, "split must be a multiple of 16");
  p.ncols = a->ncols;
  p.bias = a->bias; p.act = a->act; p.slope = a->slope;
  p.accumulate = a->accumulate; p.out_f32 = a->out_f32;
  {
    const char* e = getenv("SMSUT_DEBUG_ROWSHIFT");
    p.dbg_rowshift = e ? atoi(e) : 0;
  }
  SMSUT_CHECK(a->out0 != nullptr, -1, "null output");
  p.fast = conv_tc_plain_epilogue(a) ? 1 : 0;
  p.stats = nullptr;
  if (a->stats != nullptr) {
    SMSUT_CHECK(conv_tc_fuses_stats(a), -1, "stats requested for a shape conv_tc_kernel does not fuse them for");
    p.stats = a->stats;
    p.stats_q = det_shadow(a->stats);
  }
  if (!a->out_f32)
    SMSUT_CHECK(a->out0_ld % 8 == 0 && a->out0_coff % 8 == 0 || a->ncols < 16, -1, "bf16 output pitch/offset must be multiples of 8");

  p.ksplit = ksplit;
  // Weight-tile multicast (SMSUT_TC_MCAST=2|4, opt-in).  Every M tile of a layer reads the SAME weight tile from L2: at
  // 32x32 / 16x16 it is 55-66 % of a CTA's operand bytes (ncu: 113 MB L2->SM for a 9 MB input at 256->128 @32x32).  A
  // (mc,1,1) cluster of M tiles loads it once: CTA r fetches rows [r, r+1) * bn / mc and multicasts them.
  // MEASURED on B200 and OFF by default: bit-identical results, but every class is 1-4 us SLOWER per launch (64->64 @64x64
  // 19.5 -> 23.6 us, 256->128 @32x32 19.4 -> 20.5 / 21.3 us, 256->256 @16x16 15.4 -> 16.2 / 16.4 us with clusters of
  // 2 / 4) and the step 10.34 / 10.42 ms against 9.78 ms: at 32-512 tiles of 10-20 us these launches are bound by their
  // fixed latencies (launch, set-up, first TMA round trip, epilogue), not by L2->SM bandwidth, and a cluster adds two
  // cluster barriers, lock-step stage release across its CTAs and gang scheduling.
  p.mc = 1;
  CUtensorMap map_ws = map_w;
  {
    const char* e = getenv("SMSUT_TC_MCAST");      // read per call: the parity tests run both settings in one process
    const int want = e ? atoi(e) : 0;
    for (int c = want > 4 ? 4 : want; c > 1 && p.mc == 1 && ksplit == 1; c >>= 1) {
      if (m_tiles % c != 0 || bn % (8 * c) != 0 || ns < 2) continue;
      const int fit = max_active_clusters_x(conv_tc_kernel<true>, dim3(kThreads), smem, (unsigned)c);
      if (fit < 1) continue;
      rc = make_mat_map(&map_ws, a->wpack, ktot, a->ncols_pad, ktot, cc, bn / c);
      if (rc) return rc;
      p.mc = c;
    }
  }
  dim3 grid((unsigned)m_tiles, (unsigned)(a->ncols_pad / bn), (unsigned)ksplit);
  if (ksplit > 1)
    launch_cluster_z(conv_tc_kernel<true>, grid, kThreads, smem, stream, (unsigned)ksplit, maps[0], maps[1], maps[2], maps[3],
                     map_w, map_ws, p);
  else if (p.mc > 1)
    launch_cluster_x(conv_tc_kernel<true>, grid, kThreads, smem, stream, (unsigned)p.mc, maps[0], maps[1], maps[2], maps[3],
                     map_w, map_ws, p);
  else
    launch_pdl(conv_tc_kernel<false>, grid, kThreads, smem, stream, maps[0], maps[1], maps[2], maps[3], map_w, map_ws, p);
  count_launch();
  return launch_status("conv_tc_kernel");
}

}  // namespace smsut

extern "C" int smsut_conv_tc_fuses_stats(const smsut_conv_tc_args* a) {
  if (a == nullptr || a->ncols_pad % 16 != 0 || a->src_c[0] % 16 != 0) return 0;
  if (a->bn == 0 && smsut::conv_band_eligible_c(a)) return smsut::conv_band_fuses_stats(a) ? 1 : 0;
  return smsut::conv_tc_fuses_stats(a) ? 1 : 0;
}
extern "C" int smsut_conv_tc(const smsut_conv_tc_args* a, smsut_stream_t stream) {
  return smsut::conv_tc_impl(a, reinterpret_cast<cudaStream_t>(stream));
}
