// tubedetr_b200 -- fused time-aligned cross-attention of the space-time decoder (sm_100a, tcgen05 + TMEM + TMA).
//
// Reference: models/transformer.py:724-745 (cross_attn_image): per frame f ONE query attends to that frame's S memory
// tokens; K = (memory + pos) Wk^T + bk, V = memory Wv^T + bv.  The K/V projections are 92 % of the decoder FLOPs
// (SURVEY.md 7.H1).  This kernel computes, per 128-row tile of the flat [frames*S][256] memory matrix,
//     K tile = mempb[128x256] x Wk^T  -> TMEM columns [0,256)      (tcgen05.mma, bf16 x bf16 -> fp32)
//     V tile = memb [128x256] x Wv^T  -> TMEM columns [256,512)
// and consumes them straight out of TMEM: scores s[j,h] = scale * q[f(j),h,:] . K[j,h,:] (+ key padding mask), a
// segment-local softmax over the rows of each frame inside the tile, and the partial context sum_j p[j,h] V[j,h,:].
// K and V never go to HBM.  Frames straddle tiles (S = 141 does not divide 128), so every (tile, frame) segment emits
// flash-style partials (m, l, o[256]) that tdb_xattn_merge combines (<= 3 segments per frame), also normalising the
// probabilities that the backward pass and the `ca_weights` output need.
// The K bias bk only shifts all scores of a frame by q.bk, which the softmax cancels, so it is not applied; bv is added
// after normalisation (sum_j p = 1).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = consumers (one TMEM lane = one memory token).
#include <stdlib.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);
int tdb_init_once();
int tdb_make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
int tdb_num_sms();

namespace tdb {

constexpr int XD = 256;            // model width
constexpr int XH = 8;              // heads (head dim 32)
constexpr int XBM = 128;           // memory rows per tile
constexpr int XMAXF = 4;           // max frames touched by one tile (needs S >= 43)
constexpr int XSTAGES = 4;
constexpr int XSTAGE_BYTES = XBM * 64 * 2 + XD * 64 * 2;   // A 16 KB + B 32 KB per 64-wide k-block
constexpr int XTHREADS = 256;

struct XattnSmem {
  // after the stage ring
  uint64_t full[8], empty[8], kfull, vfull;      // 8 = ring capacity of either kernel variant
  uint32_t tmem_slot, pad;
  float qs[XMAXF][XD];             // scale * q of the frames in this tile
  float owarp[4][XMAXF][XD];       // per-consumer-warp partial contexts
  float wmax[4][XMAXF][XH];
  float wsum[4][XMAXF][XH];
};
constexpr int XSMEM_BYTES = XSTAGES * XSTAGE_BYTES + 1024 + (int)sizeof(XattnSmem);
// pair variant (cta_group::2): per k-block a CTA stages its own 128 A rows (16 KB) and HALF of the weight rows (16 KB)
constexpr int X2STAGES = 6;
constexpr int X2STAGE_BYTES = XBM * 64 * 2 + (XD / 2) * 64 * 2;
constexpr int X2SMEM_BYTES = X2STAGES * X2STAGE_BYTES + 1024 + (int)sizeof(XattnSmem);

struct XattnParams {
  const bf16* q;        // [F][256] projected queries (bias included), unscaled
  const uint8_t* kpm;   // [F][S] nonzero = padded key
  const uint8_t* keep;  // [F][8][S] attention-dropout keep mask (train mode) or null
  float keep_scale;     // 1 / (1 - p_drop)
  float* p;             // [F][8][S]  out: exp(s - m_tile) (normalised later by the merge kernel)
  float* part_m;        // [tiles][XMAXF][8]
  float* part_l;        // [tiles][XMAXF][8]
  float* part_o;        // [tiles][XMAXF][256]
  int R, S, F;
  float scale;
  long long* timing;    // optional [tiles][4] SM clock stamps: kernel entry, first MMA issue, last MMA complete, CTA exit;
                        // followed by [tiles][2] %globaltimer (ns) at CTA entry / exit (CTA launch skew across the grid)
};

__device__ __forceinline__ long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// consumers (warps 4-7 of a CTA): thread = memory token = TMEM lane of this CTA's 128-row tile; K accumulator in TMEM columns
// [0,256), V in [256,512); shared by the 1-CTA and the pair kernel
__device__ __forceinline__ void xattn_consume(XattnSmem& sh, const XattnParams& p, const int tile, const int row0,
                                              const uint32_t tmem_base, const int warp, const int lane) {
    // ------------------------------------------------------------------ consumers: thread = memory token (TMEM lane)
    const int wq = warp & 3;
    const int tid = threadIdx.x - 128;                    // 0..127
    const int j = row0 + wq * 32 + lane;                  // flat memory row
    const bool valid = j < p.R;
    const int f0 = row0 / p.S;                            // first frame of the tile
    const int f = valid ? j / p.S : 0x3fffffff;
    const int tok = valid ? j - f * p.S : 0;
    const int fl = valid ? f - f0 : XMAXF;                // local frame index, XMAXF = "none"
    const int last_row = (row0 + XBM - 1 < p.R - 1) ? row0 + XBM - 1 : p.R - 1;
    const int nfr = last_row / p.S - f0 + 1;              // frames in this tile (<= XMAXF, checked on the host)
    // stage scale*q of the tile's frames, clear accumulators
    for (int e = tid; e < XMAXF * XD; e += 128) {
      int ff = e / XD, c = e - ff * XD;
      float v = 0.f;
      if (ff < nfr) v = __bfloat162float(p.q[(long long)(f0 + ff) * XD + c]) * p.scale;
      sh.qs[ff][c] = v;
    }
    for (int e = tid; e < 4 * XMAXF * XD; e += 128) (&sh.owarp[0][0][0])[e] = 0.f;
    named_bar_sync(1, 128);
    const bool masked = valid ? (p.kpm != nullptr && p.kpm[(long long)f * p.S + tok] != 0) : true;
    // warp-uniform range of local frames present in this warp
    const int fl_lo = __shfl_sync(0xffffffffu, fl, 0);
    int fl_hi_l = valid ? fl : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fl_hi_l = max(fl_hi_l, __shfl_xor_sync(0xffffffffu, fl_hi_l, o));
    const int fl_hi = fl_hi_l;

    // ---- scores from the K accumulator
    mbar_wait(&sh.kfull, 0, 13);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
    float s[XH];
    const float* qrow = sh.qs[fl < XMAXF ? fl : 0];
#pragma unroll
    for (int h = 0; h < XH; ++h) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + h * 32, r);
      tmem_ld_wait();
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) acc += __uint_as_float(r[d]) * qrow[h * 32 + d];
      s[h] = (masked || !valid) ? -INFINITY : acc;
    }
    // ---- segment-local softmax statistics: per (local frame, head) max and sum over this tile's rows
#pragma unroll
    for (int h = 0; h < XH; ++h) {
      for (int ff = fl_lo; ff <= fl_hi; ++ff) {
        float m = warp_max(fl == ff ? s[h] : -INFINITY);
        if (lane == 0) sh.wmax[wq][ff][h] = m;
      }
    }
    if (lane < XH) {
      for (int ff = 0; ff < XMAXF; ++ff)
        if (ff < fl_lo || ff > fl_hi) sh.wmax[wq][ff][lane] = -INFINITY;
    }
    named_bar_sync(1, 128);
    float pt[XH];
#pragma unroll
    for (int h = 0; h < XH; ++h) {
      float m = -INFINITY;
      if (fl < XMAXF) m = fmaxf(fmaxf(sh.wmax[0][fl][h], sh.wmax[1][fl][h]), fmaxf(sh.wmax[2][fl][h], sh.wmax[3][fl][h]));
      pt[h] = (s[h] == -INFINITY) ? 0.f : __expf(s[h] - m);
    }
#pragma unroll
    for (int h = 0; h < XH; ++h) {
      for (int ff = fl_lo; ff <= fl_hi; ++ff) {
        float l = warp_sum(fl == ff ? pt[h] : 0.f);
        if (lane == 0) sh.wsum[wq][ff][h] = l;
      }
    }
    if (lane < XH) {
      for (int ff = 0; ff < XMAXF; ++ff)
        if (ff < fl_lo || ff > fl_hi) sh.wsum[wq][ff][lane] = 0.f;
    }
    if (valid) {
#pragma unroll
      for (int h = 0; h < XH; ++h) p.p[((long long)f * XH + h) * p.S + tok] = pt[h];
      if (p.keep != nullptr) {      // dropout acts on the probabilities that weight V; the softmax sum keeps all of them
#pragma unroll
        for (int h = 0; h < XH; ++h) pt[h] = p.keep[((long long)f * XH + h) * p.S + tok] ? pt[h] * p.keep_scale : 0.f;
      }
    }
    // ---- partial context from the V accumulator: o[ff][h*32+d] += sum_j pt[j,h] V[j,h,d]
    mbar_wait(&sh.vfull, 0, 14);
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < XH; ++h) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + XD + h * 32, r);
      tmem_ld_wait();
      for (int ff = fl_lo; ff <= fl_hi; ++ff) {
        float v[32];
        const float wgt = (fl == ff) ? pt[h] : 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) v[d] = wgt * __uint_as_float(r[d]);
        // butterfly reduce-scatter: afterwards lane d holds sum over the 32 rows of column d
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            float send = up ? v[i] : v[i + off];
            float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        sh.owarp[wq][ff][h * 32 + lane] = v[0];
      }
    }
    named_bar_sync(1, 128);
    // ---- emit the (tile, frame) partials
    for (int e = tid; e < XMAXF * XD; e += 128) {
      int ff = e / XD, c = e - ff * XD;
      float o = sh.owarp[0][ff][c] + sh.owarp[1][ff][c] + sh.owarp[2][ff][c] + sh.owarp[3][ff][c];
      p.part_o[((long long)tile * XMAXF + ff) * XD + c] = o;
    }
    if (tid < XMAXF * XH) {
      int ff = tid / XH, h = tid - ff * XH;
      float m = fmaxf(fmaxf(sh.wmax[0][ff][h], sh.wmax[1][ff][h]), fmaxf(sh.wmax[2][ff][h], sh.wmax[3][ff][h]));
      float l = sh.wsum[0][ff][h] + sh.wsum[1][ff][h] + sh.wsum[2][ff][h] + sh.wsum[3][ff][h];
      p.part_m[((long long)tile * XMAXF + ff) * XH + h] = m;
      p.part_l[((long long)tile * XMAXF + ff) * XH + h] = l;
    }
}

__global__ void __launch_bounds__(XTHREADS, 1)
xattn_fused_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ XattnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET on the __shared__ pointer: an integer round trip would turn every access below into a
  // generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  XattnSmem& sh = *reinterpret_cast<XattnSmem*>(smem + XSTAGES * XSTAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int row0 = tile * XBM;
  if (p.timing && threadIdx.x == 32) p.timing[tile * 4 + 0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < XSTAGES; ++s) {
      mbar_init(&sh.full[s], 1);
      mbar_init(&sh.empty[s], 1);
    }
    mbar_init(&sh.kfull, 1);
    mbar_init(&sh.vfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&sh.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  pdl_wait();
  pdl_trigger();
  // %globaltimer reads take microseconds: a lane of the idle warp 3 does them, after the prologue barrier, so that they sit
  // neither in the MMA thread's path nor in front of a CTA-wide barrier
  if (p.timing && threadIdx.x == 96) p.timing[(long long)gridDim.x * 4 + tile * 2 + 0] = globaltimer_ns();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < 8; ++it) {        // 4 k-blocks for K, then 4 for V
        const int kk = it & 3;
        mbar_wait(&sh.empty[stage], phase ^ 1, 11);
        mbar_expect_tx(&sh.full[stage], XSTAGE_BYTES);
        uint8_t* sA = smem + stage * XSTAGE_BYTES;
        uint8_t* sB = sA + XBM * 64 * 2;
        tma_load_2d(sA, it < 4 ? &tmK : &tmV, &sh.full[stage], kk * 64, row0);
        tma_load_2d(sB, &tmW, &sh.full[stage], kk * 64, it < 4 ? 0 : XD);   // Wk rows [0,256), Wv rows [256,512)
        if (++stage == XSTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(XBM, XD, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < 8; ++it) {
        mbar_wait(&sh.full[stage], phase, 12);
        tc_fence_after();
        if (p.timing && it == 0) p.timing[tile * 4 + 1] = clock64();
        const uint32_t a_base = smem_u32(smem + stage * XSTAGE_BYTES);
        const uint32_t b_base = a_base + XBM * 64 * 2;
        const uint32_t d_tmem = tmem_base + (it < 4 ? 0 : XD);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_bf16(d_tmem, umma_smem_desc(a_base + s * 32, 16, 1024), umma_smem_desc(b_base + s * 32, 16, 1024), idesc,
                    ((it & 3) > 0 || s > 0) ? 1u : 0u);
        umma_commit(&sh.empty[stage]);
        if (it == 3) umma_commit(&sh.kfull);
        if (it == 7) umma_commit(&sh.vfull);
        if (++stage == XSTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (p.timing) {             // MMA phase = first issue .. completion of the last tcgen05.mma (observed on the V barrier)
        mbar_wait(&sh.vfull, 0, 15);
        p.timing[tile * 4 + 2] = clock64();
      }
    }
  } else if (warp >= 4) {
    xattn_consume(sh, p, tile, row0, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (p.timing && threadIdx.x == 32) p.timing[tile * 4 + 3] = clock64();
  if (p.timing && threadIdx.x == 96) p.timing[(long long)gridDim.x * 4 + tile * 2 + 1] = globaltimer_ns();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Pair variant: a CTA pair (cluster of 2, tcgen05 cta_group::2) projects 256 memory rows per k-step with ONE MMA (M = 256,
// N = 256).  Each CTA stages its own 128 A rows and only HALF of the weight rows (both tensor cores read both halves), so a
// CTA pulls 8 x 32 KB = 256 KB through the L2 -> SM path instead of 8 x 48 KB = 384 KB for the same 4096 tensor-pipe
// cycles: the 1-CTA kernel's MMA phase is bound by exactly that feed (tools/xattn_phase.py).  Accumulators land in each
// CTA's own TMEM (rows of rank r in CTA r), so the consumer code is unchanged.  Barrier protocol = tdb_gemm2_kernel:
// every TMA load signals the LEADER's full barrier, the leader's commits are multicast to both CTAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XTHREADS, 1)
xattn_fused2_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmW, const __grid_constant__ XattnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  XattnSmem& sh = *reinterpret_cast<XattnSmem*>(smem + X2STAGES * X2STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x;                 // 128-row tile of this CTA; the pair owns tiles (2i, 2i+1)
  const int row0 = tile * XBM;
  const int ntiles = (p.R + XBM - 1) / XBM;
  if (p.timing && threadIdx.x == 32 && tile < ntiles) p.timing[tile * 4 + 0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < X2STAGES; ++s) {
      mbar_init(&sh.full[s], 1);
      mbar_init(&sh.empty[s], 1);
    }
    mbar_init(&sh.kfull, 1);
    mbar_init(&sh.vfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(&sh.tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {                          // both CTAs: own A rows + own half of the weight rows, leader's barrier
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < 8; ++it) {        // 4 k-blocks for K, then 4 for V
        const int kk = it & 3;
        mbar_wait(&sh.empty[stage], phase ^ 1, 31);
        if (rank == 0) mbar_expect_tx(&sh.full[stage], 2 * X2STAGE_BYTES);
        uint8_t* sA = smem + stage * X2STAGE_BYTES;
        uint8_t* sB = sA + XBM * 64 * 2;
        tma_load_2d_2sm(sA, it < 4 ? &tmK : &tmV, &sh.full[stage], kk * 64, row0);
        tma_load_2d_2sm(sB, &tmW, &sh.full[stage], kk * 64, (it < 4 ? 0 : XD) + (int)rank * (XD / 2));
        if (++stage == X2STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {             // leader only
      const uint32_t idesc = umma_idesc_bf16(2 * XBM, XD, 0, 0);
      const uint64_t hi = umma_smem_desc(0, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < 8; ++it) {
        mbar_wait(&sh.full[stage], phase, 32);
        tc_fence_after();
        if (p.timing && it == 0) {
          const long long t = clock64();
          p.timing[tile * 4 + 1] = t;
          if (tile + 1 < ntiles) p.timing[(tile + 1) * 4 + 1] = t;
        }
        const uint32_t a_base = smem_u32(smem + stage * X2STAGE_BYTES);
        const uint64_t ad = umma_desc_at(hi, a_base);
        const uint64_t bd = umma_desc_at(hi, a_base + XBM * 64 * 2);
        const uint32_t d_tmem = tmem_base + (it < 4 ? 0 : XD);
#pragma unroll
        for (int s = 0; s < 4; ++s) umma2_bf16(d_tmem, ad + s * 2u, bd + s * 2u, idesc, ((it & 3) > 0 || s > 0) ? 1u : 0u);
        umma2_commit_mc(&sh.empty[stage]);
        if (it == 3) umma2_commit_mc(&sh.kfull);
        if (it == 7) umma2_commit_mc(&sh.vfull);
        if (++stage == X2STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (p.timing) {
        mbar_wait(&sh.vfull, 0, 35);
        const long long t = clock64();
        p.timing[tile * 4 + 2] = t;
        if (tile + 1 < ntiles) p.timing[(tile + 1) * 4 + 2] = t;
      }
    }
  } else if (warp >= 4) {
    if (tile < ntiles) xattn_consume(sh, p, tile, row0, tmem_base, warp, lane);     // an odd tile count leaves the last peer idle
  }

  tc_fence_before();
  __syncthreads();
  if (p.timing && threadIdx.x == 32 && tile < ntiles) p.timing[tile * 4 + 3] = clock64();
  cluster_sync_all();       // both CTAs are done with each other's shared memory, barriers and TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// one CTA per frame: combine the <= 3 tile segments, normalise probabilities, add bv, emit head-mean weights
__global__ void __launch_bounds__(256) xattn_merge_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                          const float* __restrict__ part_o, const float* __restrict__ bv,
                                                          float* __restrict__ p, float* __restrict__ pbar, bf16* __restrict__ o,
                                                          const uint8_t* __restrict__ keep, float keep_scale, int S, int F) {
  pdl_wait();
  pdl_trigger();
  __shared__ float Ms[XH], Ls[XH];
  __shared__ float fac[4][XH];       // exp(m_t - M) per segment
  const int f = blockIdx.x;
  const int t0 = (f * S) / XBM, t1 = (f * S + S - 1) / XBM;
  const int h = threadIdx.x >> 5, d = threadIdx.x & 31;
  if (d == 0) {
    float M = -INFINITY;
    for (int t = t0; t <= t1; ++t) {
      int ff = f - (t * XBM) / S;
      M = fmaxf(M, part_m[((long long)t * XMAXF + ff) * XH + h]);
    }
    float L = 0.f;
    for (int t = t0; t <= t1; ++t) {
      int ff = f - (t * XBM) / S;
      float m = part_m[((long long)t * XMAXF + ff) * XH + h];
      float e = (m == -INFINITY) ? 0.f : __expf(m - M);
      fac[t - t0][h] = e;
      L += part_l[((long long)t * XMAXF + ff) * XH + h] * e;
    }
    Ms[h] = M;
    Ls[h] = L;
  }
  __syncthreads();
  float acc = 0.f;
  for (int t = t0; t <= t1; ++t) {
    int ff = f - (t * XBM) / S;
    acc += part_o[((long long)t * XMAXF + ff) * XD + h * 32 + d] * fac[t - t0][h];
  }
  const float invL = 1.f / Ls[h];
  float sd = 0.f;      // sum over tokens of the (dropped) probabilities of this head: 1 without dropout
  for (int tok = d; tok < S; tok += 32) {
    int t = (f * S + tok) / XBM;
    long long idx = ((long long)f * XH + h) * S + tok;
    float pn = p[idx] * fac[t - t0][h] * invL;
    p[idx] = pn;
    if (keep) sd += keep[idx] ? pn * keep_scale : 0.f;
  }
  sd = keep ? warp_sum(sd) : 1.f;
  // context: sum_j p_j (V_j + bv) = (sum_j p_j V_j) + bv * sum_j p_j
  o[(long long)f * XD + h * 32 + d] = __float2bfloat16(acc * invL + (bv ? bv[h * 32 + d] * sd : 0.f));
  __syncthreads();
  if (pbar) {
    for (int tok = threadIdx.x; tok < S; tok += blockDim.x) {
      float sacc = 0.f;
#pragma unroll
      for (int hh = 0; hh < XH; ++hh) {
        const long long idx = ((long long)f * XH + hh) * S + tok;
        float pv = p[idx];
        if (keep) pv = keep[idx] ? pv * keep_scale : 0.f;      // the returned weights are the dropped probabilities
        sacc += pv;
      }
      pbar[(long long)f * S + tok] = sacc * (1.f / XH);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the attention core for ONE query per frame (Lq = 1): a streaming kernel, one CTA per frame.
//   pass 1 (V rows):  dPd[j,h] = dO_h . V[j,h,:] + dPbar[j]/8,  dP = dropout'(dPd),  rs[h] = sum_j P dP,  dV[j] = Pd[j,h] dO_h
//   pass 2 (K rows):  dS[j,h] = P (dP - rs[h]),  dK[j] = scale dS q_h,  dQ_h = scale sum_j dS[j,h] K[j,h,:]
// Warp = one memory row at a time (512 B, lane = 8 channels, head = lane / 4); every K / V row is read once and every dK / dV
// row written once with 16-byte coalesced accesses -- the generic (Lq x Lk) kernels spent a CTA per (frame, head) staging
// K and V in shared memory for a single query row (profiles/r01_kernel_table.txt: mha_bwd_row/col 58 + 43 us per launch).
// Reductions across warps go through shared memory in fixed order (deterministic).
struct XattnBwdParams {
  const bf16 *q, *k, *v, *dout;   // q, dout [F][256]; k, v [F*S][256]
  const float* p;                 // [F][8][S] normalised probabilities (before dropout)
  const uint8_t* keep;            // [F][8][S] or null
  float keep_scale;
  const float* dpbar;             // [F][S] or null
  bf16 *dq, *dk, *dv;
  int F, S;
  float scale;
  long long ldk, ldv, lddk, lddv; // row strides (elements, multiples of 8) of k, v, dk, dv
};

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

constexpr int XR = 10;      // rows in flight per warp in the streaming passes (S = 141: 18 rows per warp = 2 iterations)

__global__ void __launch_bounds__(256) xattn_bwd_kernel(const XattnBwdParams a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float xs[];
  const int S = a.S;
  float* sp = xs;                 // [8][S]  P
  float* spd = sp + 8 * S;        // [8][S]  post-dropout P
  float* sdp = spd + 8 * S;       // [8][S]  dP, then dS
  float* srs = sdp + 8 * S;       // [8 warps][8 heads]
  float* sdq = srs + 64;          // [8 warps][256]
  const int f = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = lane >> 2;
  const float* pg = a.p + (long long)f * 8 * S;
  const uint8_t* kg = a.keep ? a.keep + (long long)f * 8 * S : nullptr;
  for (int e = threadIdx.x; e < 8 * S; e += 256) {
    const float pv = pg[e];
    sp[e] = pv;
    spd[e] = kg ? (kg[e] ? pv * a.keep_scale : 0.f) : pv;
  }
  float dO[8], qv[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.dout + (long long)f * XD) + lane), dO);
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.q + (long long)f * XD) + lane), qv);
  __syncthreads();
  const uint4* vrow = reinterpret_cast<const uint4*>(a.v + (long long)f * S * a.ldv) + lane;
  const uint4* krow = reinterpret_cast<const uint4*>(a.k + (long long)f * S * a.ldk) + lane;
  uint4* dvrow = reinterpret_cast<uint4*>(a.dv + (long long)f * S * a.lddv) + lane;
  uint4* dkrow = reinterpret_cast<uint4*>(a.dk + (long long)f * S * a.lddk) + lane;
  const long long sv8 = a.ldv / 8, sk8 = a.ldk / 8, sdv8 = a.lddv / 8, sdk8 = a.lddk / 8;
  const float* dpb = a.dpbar ? a.dpbar + (long long)f * S : nullptr;
  // ---- pass 1
  float rs = 0.f;
  for (int j0 = warp; j0 < S; j0 += 8 * XR) {        // XR rows of this warp per iteration: XR independent 16-byte loads in flight
    uint4 raw[XR];
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      raw[u] = j < S ? __ldg(vrow + (long long)j * sv8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      if (j < S) {                                   // warp-uniform
        float vv[8];
        unpack8(raw[u], vv);
        float d = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) d += dO[t] * vv[t];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if (dpb) d += dpb[j] * 0.125f;
        if (kg) d = kg[h * S + j] ? d * a.keep_scale : 0.f;
        const float pj = sp[h * S + j];
        if ((lane & 3) == 0) {
          sdp[h * S + j] = d;
          rs += pj * d;
        }
        const float pdj = spd[h * S + j];
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = pdj * dO[t];
        dvrow[(long long)j * sdv8] = pack8(o);
      }
    }
  }
  if ((lane & 3) == 0) srs[warp * 8 + h] = rs;
  // the key rows of pass 2's first iteration are requested BEFORE the barrier: their latency overlaps the row-sum exchange
  uint4 kpre[XR];
#pragma unroll
  for (int u = 0; u < XR; ++u) {
    const int j = warp + 8 * u;
    kpre[u] = j < S ? __ldg(krow + (long long)j * sk8) : make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  float rsum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) rsum += srs[w * 8 + h];
  // ---- pass 2
  float dqa[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) dqa[t] = 0.f;
  for (int j0 = warp; j0 < S; j0 += 8 * XR) {
    uint4 raw[XR];
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      raw[u] = j0 == warp ? kpre[u] : (j < S ? __ldg(krow + (long long)j * sk8) : make_uint4(0, 0, 0, 0));
    }
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      if (j < S) {
        float kk[8];
        unpack8(raw[u], kk);
        const float pj = sp[h * S + j];
        const float ds = pj > 0.f ? pj * (sdp[h * S + j] - rsum) : 0.f;   // masked keys have P = 0 exactly
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          dqa[t] += ds * kk[t];
          o[t] = a.scale * ds * qv[t];
        }
        dkrow[(long long)j * sdk8] = pack8(o);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) sdq[warp * XD + lane * 8 + t] = dqa[t];
  __syncthreads();
  {
    const int c = threadIdx.x;            // 256 threads = 256 channels
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sdq[w * XD + c];
    a.dq[(long long)f * XD + c] = __float2bfloat16(t * a.scale);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Forward of the one-query-per-frame attention core on PROJECTED keys / values (the decoder's hoisted-K/V path: the K / V
// projections of all decoder layers are two tdb_gemm launches over the layer-invariant memory, models/transformer.py:567-579,
// 734-740).  One CTA per frame, same streaming scheme as the backward: a warp takes one 512-byte key row at a time (lane = 8
// channels, 4 lanes = one head), scores by a 2-step butterfly, softmax by one warp per head, context by per-lane accumulation
// over the value rows and a fixed-order cross-warp reduction.
struct XattnCoreParams {
  const bf16 *q, *k, *v;          // q [F][256] projected, unscaled; k, v rows [F*S] with strides ldk / ldv
  const uint8_t* kpm;             // [F][S] nonzero = padded key, or null
  const uint8_t* keep;            // [F][8][S] or null (attention dropout)
  float keep_scale;
  bf16* o;                        // [F][256]
  float* p;                       // [F][8][S] probabilities before dropout
  float* pbar;                    // [F][S] head mean of the post-dropout probabilities, or null
  int F, S;
  float scale;
  long long ldk, ldv;
};

__global__ void __launch_bounds__(256) xattn_core_fwd_kernel(const XattnCoreParams a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float xs[];
  const int S = a.S;
  float* sp = xs;                 // [8][S] scores, then P
  float* spd = sp + 8 * S;        // [8][S] post-dropout P
  float* so = spd + 8 * S;        // [8 warps][256]
  const int f = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = lane >> 2;
  float qv[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.q + (long long)f * XD) + lane), qv);
#pragma unroll
  for (int t = 0; t < 8; ++t) qv[t] *= a.scale;
  const uint4* krow = reinterpret_cast<const uint4*>(a.k + (long long)f * S * a.ldk) + lane;
  const uint4* vrow = reinterpret_cast<const uint4*>(a.v + (long long)f * S * a.ldv) + lane;
  const long long sk8 = a.ldk / 8, sv8 = a.ldv / 8;
  const uint8_t* mk = a.kpm ? a.kpm + (long long)f * S : nullptr;
  // ---- scores (XR rows of this warp per iteration = XR independent 16-byte loads in flight: the kernel is a latency chain)
  for (int j0 = warp; j0 < S; j0 += 8 * XR) {
    uint4 raw[XR];
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      raw[u] = j < S ? __ldg(krow + (long long)j * sk8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      if (j < S) {
        float kk[8];
        unpack8(raw[u], kk);
        float d = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) d += qv[t] * kk[t];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if ((lane & 3) == 0) sp[h * S + j] = (mk && mk[j]) ? -INFINITY : d;
      }
    }
  }
  // the value rows of the context pass's first iteration are requested now: their latency overlaps the softmax
  uint4 vpre[XR];
#pragma unroll
  for (int u = 0; u < XR; ++u) {
    const int j = warp + 8 * u;
    vpre[u] = j < S ? __ldg(vrow + (long long)j * sv8) : make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  // ---- softmax: warp = head
  {
    float* row = sp + warp * S;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) sum += (mx == -INFINITY) ? 0.f : __expf(row[j] - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;          // all keys masked -> NaN, like the reference softmax
    float* pg = a.p + ((long long)f * 8 + warp) * S;
    const uint8_t* kg = a.keep ? a.keep + ((long long)f * 8 + warp) * S : nullptr;
    for (int j = lane; j < S; j += 32) {
      const float pv = ((mx == -INFINITY) ? 0.f : __expf(row[j] - mx)) * inv;
      row[j] = pv;
      pg[j] = pv;
      spd[warp * S + j] = kg ? (kg[j] ? pv * a.keep_scale : 0.f) : pv;
    }
  }
  __syncthreads();
  if (a.pbar)
    for (int j = threadIdx.x; j < S; j += 256) {
      float t = 0.f;
#pragma unroll
      for (int hh = 0; hh < 8; ++hh) t += spd[hh * S + j];
      a.pbar[(long long)f * S + j] = t * 0.125f;
    }
  // ---- context
  float oa[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) oa[t] = 0.f;
  for (int j0 = warp; j0 < S; j0 += 8 * XR) {
    uint4 raw[XR];
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      raw[u] = j0 == warp ? vpre[u] : (j < S ? __ldg(vrow + (long long)j * sv8) : make_uint4(0, 0, 0, 0));
    }
#pragma unroll
    for (int u = 0; u < XR; ++u) {
      const int j = j0 + 8 * u;
      if (j < S) {
        float vv[8];
        unpack8(raw[u], vv);
        const float pdj = spd[h * S + j];
#pragma unroll
        for (int t = 0; t < 8; ++t) oa[t] += pdj * vv[t];
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) so[warp * XD + lane * 8 + t] = oa[t];
  __syncthreads();
  {
    const int c = threadIdx.x;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += so[w * XD + c];
    a.o[(long long)f * XD + c] = __float2bfloat16(t);
  }
}

}  // namespace tdb

using namespace tdb;

extern "C" int tdb_xattn_core_fwd(const void* q, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* kpm,
                                  const uint8_t* keep, float keep_scale, void* o, float* p, float* pbar, int F, int S, float scale,
                                  void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(q && k && v && o && p && F > 0 && S > 0, "tdb_xattn_core_fwd: null argument");
  TDB_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldk >= XD && ldv >= XD,
              "tdb_xattn_core_fwd: bf16 buffers must be 16-byte aligned, row strides multiples of 8 and >= 256");
  const size_t smem = sizeof(float) * ((size_t)16 * S + 8 * XD);
  TDB_REQUIRE(smem <= 200 * 1024, "tdb_xattn_core_fwd: S=%d too long", S);
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(xattn_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  XattnCoreParams a{(const bf16*)q, (const bf16*)k, (const bf16*)v, kpm, keep, keep_scale, (bf16*)o, p, pbar, F, S, scale, (long long)ldk, (long long)ldv};
  TDB_CHECK_CUDA(tdb_launch(xattn_core_fwd_kernel, dim3(F), dim3(256), smem, (cudaStream_t)stream_, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}

extern "C" int tdb_xattn_core_bwd(const void* q, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout, const float* p,
                                  const uint8_t* keep, float keep_scale, const float* dpbar, void* dq, void* dk, int64_t lddk, void* dv,
                                  int64_t lddv, int F, int S, float scale, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(q && k && v && dout && p && dq && dk && dv && F > 0 && S > 0, "tdb_xattn_core_bwd: null argument");
  TDB_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0,
              "tdb_xattn_core_bwd: bf16 buffers must be 16-byte aligned");
  TDB_REQUIRE(ldk % 8 == 0 && ldv % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && ldk >= XD && ldv >= XD && lddk >= XD && lddv >= XD,
              "tdb_xattn_core_bwd: row strides must be multiples of 8 and >= 256");
  const size_t smem = sizeof(float) * ((size_t)24 * S + 64 + 8 * XD);
  TDB_REQUIRE(smem <= 200 * 1024, "tdb_xattn_core_bwd: S=%d too long", S);
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  XattnBwdParams a{(const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, p, keep, keep_scale, dpbar,
                   (bf16*)dq, (bf16*)dk, (bf16*)dv, F, S, scale, (long long)ldk, (long long)ldv, (long long)lddk, (long long)lddv};
  TDB_CHECK_CUDA(tdb_launch(xattn_bwd_kernel, dim3(F), dim3(256), smem, (cudaStream_t)stream_, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}

extern "C" int tdb_xattn_bwd(const void* q, const void* k, const void* v, const void* dout, const float* p, const uint8_t* keep,
                             float keep_scale, const float* dpbar, void* dq, void* dk, void* dv, int F, int S, float scale,
                             void* stream_) {
  return tdb_xattn_core_bwd(q, k, XD, v, XD, dout, p, keep, keep_scale, dpbar, dq, dk, XD, dv, XD, F, S, scale, stream_);
}

// measurement hook (tools/xattn_phase.py): when set, every fused forward launch writes 4 SM clock stamps per tile
static long long* g_xattn_timing = nullptr;
extern "C" int tdb_xattn_set_timing_buffer(void* buf) {
  g_xattn_timing = (long long*)buf;
  return TDB_OK;
}

// kernel variant: 0 = one CTA per 128-row tile, 1 = CTA pair per 256 rows (cta_group::2); default from env TDB_XATTN_PAIR
static int g_xattn_pair = -1;
extern "C" int tdb_xattn_set_pair(int on) {
  g_xattn_pair = on ? 1 : 0;
  return TDB_OK;
}

extern "C" int64_t tdb_xattn_workspace_bytes(int F, int S) {
  int64_t tiles = ((int64_t)F * S + XBM - 1) / XBM;
  return tiles * XMAXF * (XH * 2 + XD) * (int64_t)sizeof(float);
}

extern "C" int tdb_xattn_fused_fwd(const void* q, const void* mempb, const void* memb, const void* wkv, const float* bv,
                                   const uint8_t* kpm, const uint8_t* keep, float keep_scale, void* o, float* p, float* pbar,
                                   void* workspace, int64_t ws_bytes, int F, int S, float scale, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(q && mempb && memb && wkv && o && p && workspace && F > 0, "tdb_xattn_fused_fwd: null argument");
  TDB_REQUIRE(S >= 43 && S <= 4096, "tdb_xattn_fused_fwd: S=%d unsupported (a 128-row tile may touch at most %d frames)", S, XMAXF);
  TDB_REQUIRE(ws_bytes >= tdb_xattn_workspace_bytes(F, S), "tdb_xattn_fused_fwd: workspace too small");
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(xattn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XSMEM_BYTES));
    attr = true;
  }
  const int64_t R = (int64_t)F * S;
  const int tiles = (int)((R + XBM - 1) / XBM);
  CUtensorMap tmK, tmV, tmW;
  if ((rc = tdb_make_tmap_bf16(&tmK, mempb, R, XD, XD, XBM))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmV, memb, R, XD, XD, XBM))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmW, wkv, 2 * XD, XD, XD, XD))) return rc;
  XattnParams prm;
  prm.q = (const bf16*)q;
  prm.kpm = kpm;
  prm.keep = keep;
  prm.keep_scale = keep_scale;
  prm.p = p;
  float* ws = (float*)workspace;
  prm.part_m = ws;
  prm.part_l = ws + (int64_t)tiles * XMAXF * XH;
  prm.part_o = ws + (int64_t)tiles * XMAXF * XH * 2;
  prm.R = (int)R;
  prm.S = S;
  prm.F = F;
  prm.scale = scale;
  prm.timing = g_xattn_timing;
  cudaStream_t st = (cudaStream_t)stream_;
  if (g_xattn_pair < 0) {
    const char* e = getenv("TDB_XATTN_PAIR");
    g_xattn_pair = e ? atoi(e) : 0;
  }
  if (g_xattn_pair) {
    static bool attr2 = false;
    if (!attr2) {
      TDB_CHECK_CUDA(cudaFuncSetAttribute(xattn_fused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X2SMEM_BYTES));
      attr2 = true;
    }
    CUtensorMap tmWh;      // weight box = half of the 256 rows of Wk (or Wv)
    if ((rc = tdb_make_tmap_bf16(&tmWh, wkv, 2 * XD, XD, XD, XD / 2))) return rc;
    const int ctas = ((tiles + 1) / 2) * 2;
    TDB_CHECK_CUDA(tdb_launch(xattn_fused2_kernel, dim3(ctas), dim3(XTHREADS), X2SMEM_BYTES, st, tmK, tmV, tmWh, prm));
  } else {
    TDB_CHECK_CUDA(tdb_launch(xattn_fused_kernel, dim3(tiles), dim3(XTHREADS), XSMEM_BYTES, st, tmK, tmV, tmW, prm));
  }
  TDB_CHECK_CUDA(cudaGetLastError());
  TDB_CHECK_CUDA(tdb_launch(xattn_merge_kernel, dim3(F), dim3(256), 0, st, prm.part_m, prm.part_l, prm.part_o, bv, p, pbar, (bf16*)o, keep, keep_scale, S, F));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(2);
  return TDB_OK;
}
