// Weight gradient as a split-K GEMM on tcgen05 tensor cores (sm_100a).
//
//   conv  : dW[co][ci][tap] = sum_p dy[p][co] * x[p + tap][ci]      A = dy (M = co),  B = shifted x tiles
//   convT : dW[ci][co][tap] = sum_p x[p][ci]  * dy[2p + tap][co]    A = x  (M = ci),  B = strided dy planes
//
// The reduction dimension is the pixel index, which is the OUTER dimension of the NHWC tensors, so both
// operands are "MN-major": the very same TMA boxes the forward kernel uses ([128 pixels][chunk channels],
// 32/64/128-byte swizzle) are consumed with a_major = b_major = MN.  Several (tap, channel-chunk) B tiles
// are laid side by side in shared memory (LBO = tile bytes) so one UMMA covers up to 256 weight columns;
// narrow A operands (< 128 channels) replicate their single block (LBO = 0) and the duplicate accumulator
// rows are simply not written back.  Each CTA walks a contiguous range of pixel tiles accumulating in TMEM,
// then adds its partial result to the fp32 OIHW gradient with atomics (split-K across CTAs).
//
// Reference semantics: aten::convolution_backward(weight) behind every nn.Conv2d / nn.ConvTranspose2d of
// network/blocks.py:10-16,41.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

int make_act_map(CUtensorMap* m, const void* ptr, int c, int w, int h, int n, int64_t sw, int64_t sh, int64_t sn,
                 int cc, int tw, int th, int tn);
int choose_tile(int n, int h, int w, int* tn, int* th, int* tw);
void count_launch();

constexpr int kWgMaxBlocks = 48;
constexpr int kWgThreads = 192;

struct BBlock {
  int8_t map, dy, dx, tap;
  int16_t c0;    // channel coordinate in the tensor map
  int16_t pad;
};

struct WgradParams {
  int n, h, w;
  int tn, th, tw, tiles_h, tiles_w, tiles_total, tiles_per_cta;
  // A operand
  int a_real_blocks, a_chunk;
  uint32_t a_block_bytes, a_lbo, a_layout, a_sbo, a_kadv;
  // B operand
  int b_chunk, nblk_total, blk_per_group;
  uint32_t b_block_bytes, b_layout, b_sbo, b_kadv;
  int stages;
  uint32_t stage_bytes, tmem_cols;
  // output mapping: dw[((m_off + m) * nc + c_off + c) * taps + tap]
  int dbg_noepi;   // development: skip the atomics (SMSUT_WGRAD_NOEPI=1) to time the mainloop alone
  int tap_major;   // dw is the tap-major scratch [tap][m_full][nc]
  int m_full;      // rows of the full weight matrix (tap-major addressing)
  float* dw;
  long long* dw_q;         // fixed-point shadow of dw in deterministic mode (common.cuh), else nullptr
  int m_total, m_off, nc, c_off, taps, c_valid;
  BBlock blocks[kWgMaxBlocks];
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b0,
                const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                const __grid_constant__ CUtensorMap map_b3, const __grid_constant__ WgradParams p) {
  pdl_trigger();   // the next kernel may be scheduled; this one waits for its predecessor after its own set-up
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[4];
  __shared__ __align__(8) uint64_t empty_bar[4];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

  const int tile_begin = blockIdx.x * p.tiles_per_cta;
  int tile_end = tile_begin + p.tiles_per_cta;
  if (tile_end > p.tiles_total) tile_end = p.tiles_total;
  const int ntiles = tile_end - tile_begin;  // host guarantees >= 1
  const int m0 = blockIdx.y * 128;
  const int blk0 = blockIdx.z * p.blk_per_group;
  int nblk = p.nblk_total - blk0;
  if (nblk > p.blk_per_group) nblk = p.blk_per_group;
  const int ncols = nblk * p.b_chunk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b0);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
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
  pdl_wait();      // barriers, TMEM and descriptors are ready: now the predecessor's results are needed

  const uint32_t a_stage_bytes = (uint32_t)p.a_real_blocks * p.a_block_bytes;

  if (warp == 0) {
    {
      const uint32_t el = elect_one_u32();   // warp-wide loop, elected issue (see common.cuh)
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile_begin; t < tile_end; ++t) {
        int tile = t;
        const int tw_i = tile % p.tiles_w; tile /= p.tiles_w;
        const int th_i = tile % p.tiles_h; tile /= p.tiles_h;
        const int n0 = tile * p.tn, h0 = th_i * p.th, w0 = tw_i * p.tw;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx_e(&full_bar[stage], a_stage_bytes + (uint32_t)nblk * p.b_block_bytes, el);
        uint8_t* base = smem_al + (size_t)stage * p.stage_bytes;
        for (int b = 0; b < p.a_real_blocks; ++b)
          tma_load_4d_e(base + (size_t)b * p.a_block_bytes, &map_a, &full_bar[stage], m0 + b * p.a_chunk, w0, h0, n0, el);
        uint8_t* bb = base + a_stage_bytes;
        for (int j = 0; j < nblk; ++j) {
          const BBlock blk = p.blocks[blk0 + j];
          const CUtensorMap* m =
              blk.map == 0 ? &map_b0 : (blk.map == 1 ? &map_b1 : (blk.map == 2 ? &map_b2 : &map_b3));
          tma_load_4d_e(bb + (size_t)j * p.b_block_bytes, m, &full_bar[stage], blk.c0, w0 + blk.dx, h0 + blk.dy, n0, el);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // fp32 accum, bf16 x bf16, A and B MN-major, N = ncols, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(ncols >> 3) << 17) | ((128u >> 4) << 24);
    // constant descriptor words outside the loop; per UMMA only the start address advances
    const uint64_t a_hi = make_smem_desc(0, p.a_lbo, p.a_sbo, p.a_layout) & 0xFFFFFFFFFFFF0000ull;
    const uint64_t b_hi = make_smem_desc(0, p.b_block_bytes, p.b_sbo, p.b_layout) & 0xFFFFFFFFFFFF0000ull;
    const uint32_t a_step = p.a_kadv >> 4, b_step = p.b_kadv >> 4, stage_u = p.stage_bytes >> 4;
    const uint32_t a_off = a_stage_bytes >> 4, base_u = smem_base >> 4;
    const uint32_t el = elect_one_u32();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      {
        const uint32_t a0 = base_u + (uint32_t)stage * stage_u;
        const uint32_t b0 = a0 + a_off;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 128 pixels / 16 per UMMA
          umma_bf16_e(tmem_base, a_hi | (uint64_t)((a0 + k * a_step) & 0x3FFFu),
                      b_hi | (uint64_t)((b0 + k * b_step) & 0x3FFFu), idesc, (t | k) != 0 ? 1u : 0u, el);
        }
        umma_commit_e(&empty_bar[stage], el);
        if (t == ntiles - 1) umma_commit_e(&tmem_full_bar, el);
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;  // accumulator row = channel of the A operand
    // rows beyond the real A channels are duplicates (LBO = 0) or another layer's garbage: skip them
    const bool row_ok = (q * 32 + lane) < p.a_real_blocks * p.a_chunk && m < p.m_total;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int nch = ncols >> 4;
    for (int j = 0; j < nch; ++j) {
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), raw);
      tmem_ld_wait();
      if (!row_ok) continue;
      const int col = j * 16;
      const int jb = col / p.b_chunk;
      const BBlock blk = p.blocks[blk0 + jb];
      const int c = blk.c0 + (col - jb * p.b_chunk);
      if (p.dbg_noepi) continue;
      if (p.tap_major) {
        // 16 contiguous floats of row m in the tap-major scratch: four 128-bit reductions
        float* dst = p.dw + ((size_t)blk.tap * p.m_full + (p.m_off + m)) * p.nc + p.c_off + c;
        long long* dq = p.dw_q != nullptr ? p.dw_q + (dst - p.dw) : nullptr;     // deterministic mode: fixed-point shadow
        if (dq == nullptr && c + 16 <= p.c_valid && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(raw[i])),
                         "f"(__uint_as_float(raw[i + 1])), "f"(__uint_as_float(raw[i + 2])),
                         "f"(__uint_as_float(raw[i + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c + i < p.c_valid) acc_add(dst + i, dq != nullptr ? dq + i : nullptr, __uint_as_float(raw[i]));
        }
        continue;
      }
      float* dst = p.dw + ((size_t)(p.m_off + m) * p.nc + p.c_off + c) * p.taps + blk.tap;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < p.c_valid)
          acc_add(dst + (size_t)i * p.taps, p.dw_q != nullptr ? p.dw_q + (dst - p.dw) + (size_t)i * p.taps : nullptr,
                  __uint_as_float(raw[i]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static uint32_t layout_for_chunk(int cc) { return cc == 64 ? 2u : (cc == 32 ? 4u : 6u); }

int wgrad_band_try(const smsut_wgrad_tc_args* a, cudaStream_t stream);
int wgrad_hmma_try(const smsut_wgrad_tc_args* a, cudaStream_t stream);

static int wgrad_tc_impl(const smsut_wgrad_tc_args* a, cudaStream_t stream) {
  SMSUT_CHECK(a != nullptr, -1, "null args");
  if (a->dw != nullptr && a->x_c % 16 == 0 && a->dy_c % 16 == 0) {
    // wide, narrow-channel layers: warp-level MMAs over shared-memory row rings (wgrad_hmma.cu); SMSUT_WGRAD_HMMA=0
    // falls through to the tcgen05 band kernel (rows fetched once by TMA, taps by descriptor arithmetic)
    int rb = wgrad_hmma_try(a, stream);
    if (rb != 0) return rb < 0 ? rb : 0;
    rb = wgrad_band_try(a, stream);
    if (rb != 0) return rb < 0 ? rb : 0;
  }
  SMSUT_CHECK(a->kind == SMSUT_TC_CONV || a->kind == SMSUT_TC_CONVT_FWD, -1, "wgrad kind must be CONV or CONVT_FWD");
  SMSUT_CHECK(a->x_c % 16 == 0 && a->dy_c % 16 == 0 && a->x_c > 0 && a->dy_c > 0, -1,
              "wgrad channel counts must be multiples of 16 (x %d, dy %d)", a->x_c, a->dy_c);
  static bool attr_set = false;
  if (!attr_set) {
    SMSUT_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  { const char* e = getenv("SMSUT_WGRAD_NOEPI"); p.dbg_noepi = (e && e[0] == '1') ? 1 : 0; }
  p.n = a->n; p.h = a->h; p.w = a->w;
  int rc = choose_tile(a->n, a->h, a->w, &p.tn, &p.th, &p.tw);
  if (rc) return rc;
  p.tiles_h = a->h / p.th; p.tiles_w = a->w / p.tw;
  SMSUT_CHECK(p.tiles_h * p.th == a->h && p.tiles_w * p.tw == a->w, -5, "spatial dims %dx%d not tileable", a->h, a->w);
  p.tiles_total = ((a->n + p.tn - 1) / p.tn) * p.tiles_h * p.tiles_w;

  const bool convt = a->kind == SMSUT_TC_CONVT_FWD;
  // A = tensor indexed by the un-shifted pixel; B = tap-dependent tensor
  const void* a_ptr = convt ? a->x : a->dy;
  const int a_c = convt ? a->x_c : a->dy_c;
  const int a_ld = convt ? a->x_ld : a->dy_ld;
  const int b_c = convt ? a->dy_c : a->x_c;
  const int b_ld = convt ? a->dy_ld : a->x_ld;

  int ac = 64; while (a_c % ac) ac >>= 1;
  int bc = 64; while (b_c % bc) bc >>= 1;
  p.a_chunk = ac; p.b_chunk = bc;
  p.a_block_bytes = 128u * ac * 2; p.b_block_bytes = 128u * bc * 2;
  p.a_layout = layout_for_chunk(ac); p.b_layout = layout_for_chunk(bc);
  p.a_sbo = 8u * ac * 2; p.b_sbo = 8u * bc * 2;
  p.a_kadv = 16u * ac * 2; p.b_kadv = 16u * bc * 2;
  const int a_blocks_needed = 128 / ac;                       // blocks covering M = 128
  const int a_real = (a_c >= 128) ? a_blocks_needed : 1;      // channels >= 128: all real, else replicate block 0
  // (a_c of 64 with ac = 64 -> 1 real block of 2; a_c of 16/32 -> 1 real block of 8/4)
  p.a_real_blocks = a_real;
  p.a_lbo = (a_real == a_blocks_needed) ? p.a_block_bytes : 0u;
  if (a_c < 128) SMSUT_CHECK(a_c == ac, -1, "unsupported A channel count %d", a_c);

  CUtensorMap map_a, map_b[4];
  memset(map_b, 0, sizeof(map_b));
  rc = make_act_map(&map_a, a_ptr, a_c, a->w, a->h, a->n, a_ld, (int64_t)a_ld * a->w, (int64_t)a_ld * a->w * a->h, ac,
                    p.tw, p.th, p.tn);
  if (rc) return rc;

  int nb = 0;
  if (!convt) {
    SMSUT_CHECK(a->ksize == 1 || a->ksize == 3 || a->ksize == 5, -1, "ksize must be 1, 3 or 5");
    rc = make_act_map(&map_b[0], a->x, a->x_c, a->w, a->h, a->n, b_ld, (int64_t)b_ld * a->w,
                      (int64_t)b_ld * a->w * a->h, bc, p.tw, p.th, p.tn);
    if (rc) return rc;
    const int r = a->ksize / 2;
    int t = 0;
    for (int dy = -r; dy <= r; ++dy)
      for (int dx = -r; dx <= r; ++dx, ++t)
        for (int c0 = 0; c0 < b_c; c0 += bc) {
          SMSUT_CHECK(nb < kWgMaxBlocks, -6, "too many B blocks");
          BBlock& b = p.blocks[nb++];
          b.map = 0; b.dy = (int8_t)dy; b.dx = (int8_t)dx; b.tap = (int8_t)t; b.c0 = (int16_t)c0;
        }
    p.taps = a->ksize * a->ksize;
    p.m_total = a->dy_c < a->cout_total ? a->dy_c : a->cout_total;  // dy may carry zero padding channels
    p.m_off = 0; p.nc = a->cin_total; p.c_off = a->ci_off;
    p.m_full = a->cout_total;
    p.c_valid = a->c_valid > 0 ? a->c_valid : a->x_c;
  } else {
    const int64_t W2 = 2 * (int64_t)a->w, H2 = 2 * (int64_t)a->h;
    for (int t = 0; t < 4; ++t) {
      const int ty = t >> 1, tx = t & 1;
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a->dy) + ((int64_t)ty * W2 + tx) * b_ld;
      rc = make_act_map(&map_b[t], base, b_c, a->w, a->h, a->n, 2 * (int64_t)b_ld, 2 * W2 * b_ld, H2 * W2 * b_ld, bc,
                        p.tw, p.th, p.tn);
      if (rc) return rc;
      for (int c0 = 0; c0 < b_c; c0 += bc) {
        SMSUT_CHECK(nb < kWgMaxBlocks, -6, "too many B blocks");
        BBlock& b = p.blocks[nb++];
        b.map = (int8_t)t; b.dy = 0; b.dx = 0; b.tap = (int8_t)t; b.c0 = (int16_t)c0;
      }
    }
    p.taps = 4;
    p.m_total = a->x_c; p.m_off = a->ci_off; p.nc = a->cout_total; p.c_off = 0;
    p.m_full = a->cin_total;
    p.c_valid = a->dy_c;
  }
  p.nblk_total = nb;
  p.blk_per_group = 256 / bc;
  if (p.blk_per_group > nb) p.blk_per_group = nb;
  int ngroups = (nb + p.blk_per_group - 1) / p.blk_per_group;
  p.blk_per_group = (nb + ngroups - 1) / ngroups;   // balance the column groups (e.g. 9 blocks -> 5 + 4, not 8 + 1)
  ngroups = (nb + p.blk_per_group - 1) / p.blk_per_group;
  const int mblocks = (a_c + 127) / 128;

  uint32_t tc = 32;
  while ((int)tc < p.blk_per_group * bc) tc <<= 1;
  p.tmem_cols = tc;

  p.stage_bytes = (uint32_t)a_real * p.a_block_bytes + (uint32_t)p.blk_per_group * p.b_block_bytes;
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  int stages = (int)((200u * 1024u) / p.stage_bytes);
  if (stages > 3) stages = 3;
  if (stages < 1) stages = 1;
  p.stages = stages;

  // split-K over pixel tiles
  // CTAs over the whole launch.  These kernels run on side streams beside the dgrad chain, so SM-time matters more than
  // latency: every CTA pays a full-accumulator reduction epilogue, fewer and longer CTAs leave the SMs to the others.
  static int target_knob = -1;
  if (target_knob < 0) {
    const char* e = getenv("SMSUT_WGRAD_TARGET");
    target_knob = e && atoi(e) > 0 ? atoi(e) : (device_sm_count() + 2) / 3;   // measured: 296 -> 13.0 ms/step, 48 -> 12.3
  }
  int target = target_knob;
  int splits = target / (ngroups * mblocks);
  if (splits < 1) splits = 1;
  if (splits > p.tiles_total) splits = p.tiles_total;
  p.tiles_per_cta = (p.tiles_total + splits - 1) / splits;
  splits = (p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta;

  p.dw = a->dw;
  p.dw_q = det_shadow(a->dw);
  p.tap_major = a->dw_layout == 1 ? 1 : 0;
  SMSUT_CHECK(a->dw != nullptr, -1, "null dw");
  const size_t smem = (size_t)stages * p.stage_bytes + 1024;
  dim3 grid((unsigned)splits, (unsigned)mblocks, (unsigned)ngroups);
  launch_pdl(wgrad_tc_kernel, grid, kWgThreads, smem, stream, map_a, map_b[0], map_b[1], map_b[2], map_b[3], p);
  count_launch();
  return launch_status("wgrad_tc_kernel");
}

}  // namespace smsut

extern "C" int smsut_wgrad_tc(const smsut_wgrad_tc_args* a, smsut_stream_t stream) {
  return smsut::wgrad_tc_impl(a, reinterpret_cast<cudaStream_t>(stream));
}
