// tubedetr_b200 -- the tcgen05 / TMEM / TMA GEMM behind every convolution and linear layer (sm_100a).
//
// Persistent, warp-specialised, cta_group::1, tile 128 x BN x 64 (BN in {64,128,256}):
//   warp 0   : TMA producer (warp-uniform loop, one elected lane issues)   global -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1   : MMA issuer   (same scheme)         tcgen05.mma kind::f16 bf16 x bf16 -> fp32 in TMEM, tcgen05.commit
//   warp 2   : TMEM allocator / deallocator      2 x BN columns = double-buffered accumulator
//   warps 4-7: epilogue                          tcgen05.ld -> scale/bias/residual/ReLU/mask -> row-remapped global stores
// The accumulator double buffer lets the epilogue of tile i overlap the main loop of tile i+1.
// Implicit 3x3 convolution = 9 taps inside the reduction loop, each tap a constant row offset into the zero-haloed
// activation matrix (DESIGN.md "padded-grid implicit GEMM"); TMA zero-fills out-of-range rows.
// Replaces reference call sites K2-K5, K7, K9-K11, K13: cuDNN convs + FrozenBatchNorm2d
// (models/backbone.py:60-70,118-122), input_proj (models/tubedetr.py:131,134), all nn.Linear of models/transformer.py.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

namespace tdb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 256;
constexpr int kChunkBytes = 64 * 128;  // one [64 rows][128 B] swizzled chunk

struct GemmKParams {
  int M, N;
  int kb_per_tap;
  int ntaps;
  int a_major, b_major;
  int a_off0[TDB_MAX_TAPS], a_off1[TDB_MAX_TAPS], b_off0[TDB_MAX_TAPS], b_off1[TDB_MAX_TAPS];
  int nz, splits, kb_per_split;
  int z_b_off1[TDB_MAX_TAPS], z_out_col[TDB_MAX_TAPS];
  int m_tiles, n_tiles, total_work;
  int m_tile0;  // first M tile of this launch (tail launches start past the full waves)
  const float* scale;
  const float* bias;
  const bf16* residual;
  long long ldr;
  const bf16* mask;
  long long ldmask;
  int relu;
  void* out;
  int out_f32;
  long long ldo;
  int remap, img_h, img_w;
  int debug_flags;
  int clc, clc_unit;     // 1: one CTA per `clc_unit` tiles + cluster launch control (work stealing, tdb_common.cuh) instead of the static persistent schedule
  int epi_mode;
  // halo mode (see tdb_gemm2.cu): one row-haloed A tile per k-block serves all taps; B either streams through a ring of
  // BN x 64 stages or, when every (tap, k-block) slice fits next to the A tiles, is loaded once and stays resident
  int halo, a_min_off, a_box_rows, a_nbox, a_tile_bytes, na_stages, nb_stages, b_resident;
};

template <int BN, int EPI = 0>
struct GemmCfg {
  static constexpr int kStageBytes = BM * BK * 2 + BN * BK * 2;
  // mode 4 (TMA-fed residual, BN = 128 only) trades two pipeline stages for a double-buffered 128 x 128 residual tile
  // mode 5 (TMA residual in, TMA store out, BN = 128): three 128 x 128 tiles that hold the residual, then the output in place
  static constexpr int kStages = (EPI == 4 || EPI == 5) ? 4 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kOperandSlots = (EPI == 4) ? 2 : (EPI == 5 ? 3 : 0);
  static constexpr int kOperandBytes = kOperandSlots * BM * BN * 2;
  static constexpr int kStagingBytes = (EPI == 5) ? 1024 : 4 * 32 * 64 * 4;  // per epilogue warp 32 x 64 fp32 (mode 5: shared scale/bias)
  static constexpr int kSmemBytes = kStages * kStageBytes + kOperandBytes + 1024 /*align slack*/ + 512 /*barriers*/ + kStagingBytes + 256 /*ClcShared*/;
};

struct WorkItem {
  int m0, n0, z, split, kb_begin, iters;
};

__device__ __forceinline__ WorkItem decode_work(const GemmKParams& p, int w, int BN) {
  WorkItem it;
  int nt = w % p.n_tiles;
  int r = w / p.n_tiles;
  int mt = r % p.m_tiles;
  r /= p.m_tiles;
  it.split = r % p.splits;
  it.z = r / p.splits;
  it.m0 = (p.m_tile0 + mt) * BM;
  it.n0 = nt * BN;
  if (p.splits == 1) {
    it.kb_begin = 0;
    it.iters = p.ntaps * p.kb_per_tap;
  } else {
    it.kb_begin = it.split * p.kb_per_split;
    int e = it.kb_begin + p.kb_per_split;
    if (e > p.kb_per_tap) e = p.kb_per_tap;
    it.iters = e - it.kb_begin;
  }
  return it;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
tdb_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN, EPI>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET on the __shared__ pointer: an integer round trip would turn every access below into a
  // generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* opnd = smem + Cfg::kStages * Cfg::kStageBytes;     // [2][BN/64][128 rows][128 B] residual tiles (mode 4), 1024-aligned
  uint8_t* after = opnd + Cfg::kOperandBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint64_t* ofull_bar = tempty_bar + 2;             // [3] residual tile landed (modes 4, 5)
  uint64_t* oempty_bar = ofull_bar + 3;             // [3] residual / output tile free again
  uint64_t* afull_bar = oempty_bar + 3;             // [4] halo mode: row-haloed A tile landed
  uint64_t* aempty_bar = afull_bar + 4;             // [4] ... and consumed by all taps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 4);
  ClcShared* clcq = reinterpret_cast<ClcShared*>(after + 512 + Cfg::kStagingBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(&ofull_bar[s], 1);
      mbar_init(&oempty_bar[s], EPI == 5 ? 1 : 4);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&afull_bar[s], 1);
      mbar_init(&aempty_bar[s], 1);
    }
    for (int s = 0; s < kClcStages; ++s) {     // consumers: producer + MMA + 4 epilogue warps
      mbar_init(&clcq->full[s], 1);
      mbar_init(&clcq->empty[s], 6);
    }
    if (EPI == 5) tma_prefetch_desc(&tmO);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; from here on we read what it wrote
  pdl_trigger();
  TileCursor cur;
  cur.init(clcq, p.clc, (int)gridDim.x, 0, p.clc_unit, p.total_work);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the warp-uniform loop (coordinates and barrier addresses stay in uniform registers); one elected
    // lane arms the mbarrier and issues the bulk-tensor copies.  No per-stage division: taps and k-blocks are nested loops.
    {
      int stage = 0;
      uint32_t phase = 0;
      int oslot = 0;
      uint32_t ophase = 0;
      if ((EPI == 3 || EPI == 1) && p.halo) {
        uint8_t* bring = smem + p.na_stages * p.a_tile_bytes;
        constexpr int kBStage = BN * 128;
        int sa = 0;
        uint32_t pa = 0;
        bool first = true;
        for (int w = cur.first((int)blockIdx.x); w < p.total_work; w = cur.next(w)) {
          cur.request(1);
          const WorkItem wi = decode_work(p, w, BN);
          if (p.b_resident && first) {       // all (k-block, tap) weight slices once; they stay for every tile of this CTA
            first = false;
            if (elect_one()) {
              mbar_expect_tx(&full_bar[0], p.kb_per_tap * p.ntaps * kBStage);
              for (int kk = 0; kk < p.kb_per_tap; ++kk)
                for (int tap = 0; tap < p.ntaps; ++tap) {
                  uint8_t* sB = bring + (kk * p.ntaps + tap) * kBStage;
                  if (p.b_major == 0) {
                    tma_load_2d(sB, &tmB, &full_bar[0], kk * BK + p.b_off0[tap], wi.n0 + p.b_off1[tap]);
                  } else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                      tma_load_2d(sB + j * kChunkBytes, &tmB, &full_bar[0], wi.n0 + j * 64 + p.b_off0[tap], kk * BK + p.b_off1[tap]);
                  }
                }
            }
            __syncwarp();
          }
          for (int kk = 0; kk < p.kb_per_tap; ++kk) {
            mbar_wait(&aempty_bar[sa], pa ^ 1, 9);
            if (elect_one()) {
              mbar_expect_tx(&afull_bar[sa], p.a_tile_bytes);
              uint8_t* sA = smem + sa * p.a_tile_bytes;
              for (int j = 0; j < p.a_nbox; ++j)
                tma_load_2d(sA + j * p.a_box_rows * 128, &tmA, &afull_bar[sa], kk * BK + p.a_off0[0],
                            wi.m0 + p.a_min_off + j * p.a_box_rows);
            }
            __syncwarp();
            if (++sa == p.na_stages) {
              sa = 0;
              pa ^= 1;
            }
            if (!p.b_resident) {
              for (int tap = 0; tap < p.ntaps; ++tap) {
                mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                if (elect_one()) {
                  mbar_expect_tx(&full_bar[stage], kBStage);
                  uint8_t* sB = bring + stage * kBStage;
                  if (p.b_major == 0) {
                    tma_load_2d(sB, &tmB, &full_bar[stage], kk * BK + p.b_off0[tap], wi.n0 + p.b_off1[tap]);
                  } else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                      tma_load_2d(sB + j * kChunkBytes, &tmB, &full_bar[stage], wi.n0 + j * 64 + p.b_off0[tap],
                                  kk * BK + p.b_off1[tap]);
                  }
                }
                __syncwarp();
                if (++stage == p.nb_stages) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
        }
      } else
      for (int w = cur.first((int)blockIdx.x); w < p.total_work; w = cur.next(w)) {
        cur.request(1);
        const WorkItem wi = decode_work(p, w, BN);
        if constexpr (EPI == 5) {
          if (p.residual != nullptr) {      // residual tile, two tiles ahead of the epilogue (3 slots)
            mbar_wait(&oempty_bar[oslot], ophase ^ 1, 7);
            if (elect_one()) {
              mbar_expect_tx(&ofull_bar[oslot], BM * BN * 2);
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(opnd + (oslot * (BN / 64) + j) * (BM * 128), &tmR, &ofull_bar[oslot], wi.n0 + j * 64, wi.m0);
            }
            __syncwarp();
          }
          if (++oslot == 3) {
            oslot = 0;
            ophase ^= 1;
          }
        }
        if constexpr (EPI == 4) {
          // whole residual tile of this work item (128 rows x BN columns) by TMA, one tile ahead of the epilogue: its HBM
          // latency overlaps the previous tile's epilogue and this tile's main loop, with 2 x 32 KB in flight per SM
          mbar_wait(&oempty_bar[oslot], ophase ^ 1, 5);
          if (elect_one()) {
            mbar_expect_tx(&ofull_bar[oslot], BM * BN * 2);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(opnd + (oslot * (BN / 64) + j) * (BM * 128), &tmR, &ofull_bar[oslot], wi.n0 + j * 64, wi.m0);
          }
          __syncwarp();
          if (++oslot == 2) {
            oslot = 0;
            ophase ^= 1;
          }
        }
        const int ntap = p.splits == 1 ? p.ntaps : 1;
        const int kb0 = p.splits == 1 ? 0 : wi.kb_begin;
        const int nkb = p.splits == 1 ? p.kb_per_tap : wi.iters;
        const int zb = p.z_b_off1[wi.z];
        for (int tap = 0; tap < ntap; ++tap) {
          const int a0 = p.a_off0[tap], a1 = p.a_off1[tap], b0 = p.b_off0[tap], b1 = p.b_off1[tap];
          for (int kk = kb0; kk < kb0 + nkb; ++kk) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 1);
            if (elect_one()) {
              mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              uint8_t* sA = smem + stage * Cfg::kStageBytes;
              uint8_t* sB = sA + BM * BK * 2;
              if (p.a_major == 0) {
                tma_load_2d(sA, &tmA, &full_bar[stage], kk * BK + a0, wi.m0 + a1);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * kChunkBytes, &tmA, &full_bar[stage], wi.m0 + j * 64 + a0, kk * BK + a1);
              }
              if (p.b_major == 0) {
                tma_load_2d(sB, &tmB, &full_bar[stage], kk * BK + b0, wi.n0 + b1);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_2d(sB + j * kChunkBytes, &tmB, &full_bar[stage], wi.n0 + j * 64 + b0, kk * BK + b1 + zb);
              }
            }
            __syncwarp();
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the (warp-uniform) loop; one elected lane issues tcgen05.mma / tcgen05.commit.
    {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_major, p.b_major);
      const uint32_t a_k16 = (p.a_major ? 2048u : 32u) >> 4;      // descriptor address units (16 B) per K step of 16
      const uint32_t b_k16 = (p.b_major ? 2048u : 32u) >> 4;
      const uint32_t a_lbo = p.a_major ? (uint32_t)kChunkBytes : 16u;
      const uint32_t b_lbo = p.b_major ? (uint32_t)kChunkBytes : 16u;
      const uint64_t a_hi = (p.a_major && (p.debug_flags & 1)) ? umma_smem_desc(0, 1024, a_lbo) : umma_smem_desc(0, a_lbo, 1024);
      const uint64_t b_hi = (p.b_major && (p.debug_flags & 1)) ? umma_smem_desc(0, 1024, b_lbo) : umma_smem_desc(0, b_lbo, 1024);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if ((EPI == 3 || EPI == 1) && p.halo) {
        const uint32_t bring = smem_base + p.na_stages * p.a_tile_bytes;
        constexpr uint32_t kBStage = BN * 128;
        int sa = 0;
        uint32_t pa = 0;
        bool first = true;
        for (int w = cur.first((int)blockIdx.x); w < p.total_work; w = cur.next(w)) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
          const uint32_t d_tmem = tmem_base + acc * BN;
          if (p.b_resident && first) {
            first = false;
            mbar_wait(&full_bar[0], 0, 3);
          }
          for (int kk = 0; kk < p.kb_per_tap; ++kk) {
            mbar_wait(&afull_bar[sa], pa, 10);
            const uint32_t a_tile = smem_base + sa * p.a_tile_bytes;
            for (int tap = 0; tap < p.ntaps; ++tap) {
              uint32_t b_base;
              if (p.b_resident) {
                b_base = bring + (uint32_t)(kk * p.ntaps + tap) * kBStage;
              } else {
                mbar_wait(&full_bar[stage], phase, 3);
                b_base = bring + stage * kBStage;
              }
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ad = umma_desc_at(a_hi, a_tile + (uint32_t)(p.a_off1[tap] - p.a_min_off) * 128u);
                const uint64_t bd = umma_desc_at(b_hi, b_base);
#pragma unroll
                for (int s = 0; s < BK / 16; ++s)
                  umma_bf16(d_tmem, ad + s * 2u, bd + s * b_k16, idesc, (kk > 0 || tap > 0 || s > 0) ? 1u : 0u);
                if (!p.b_resident) umma_commit(&empty_bar[stage]);
              }
              __syncwarp();
              if (!p.b_resident && ++stage == p.nb_stages) {
                stage = 0;
                phase ^= 1;
              }
            }
            if (elect_one()) umma_commit(&aempty_bar[sa]);
            __syncwarp();
            if (++sa == p.na_stages) {
              sa = 0;
              pa ^= 1;
            }
          }
          if (elect_one()) umma_commit(&tfull_bar[acc]);
          __syncwarp();
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      } else
      for (int w = cur.first((int)blockIdx.x); w < p.total_work; w = cur.next(w)) {
        const WorkItem wi = decode_work(p, w, BN);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < wi.iters; ++it) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_base = smem_base + stage * Cfg::kStageBytes;
            const uint64_t ad = umma_desc_at(a_hi, a_base);
            const uint64_t bd = umma_desc_at(b_hi, a_base + BM * BK * 2);
#pragma unroll
            for (int s = 0; s < BK / 16; ++s) umma_bf16(d_tmem, ad + s * a_k16, bd + s * b_k16, idesc, (it > 0 || s > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          }
          __syncwarp();
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int wq = warp & 3;
    float* stage = reinterpret_cast<float*>(after + 512) + (EPI == 5 ? 0 : wq * (32 * 64));
    int oslot = 0;
    uint32_t ophase = 0;
    int tiles_done = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int Hp = p.img_h + 2, Wp = p.img_w + 2;
    for (int w = cur.first((int)blockIdx.x); w < p.total_work; w = cur.next(w)) {
      const WorkItem wi = decode_work(p, w, BN);
      const int row_t = wi.m0 + wq * 32 + lane;
      bool valid = row_t < p.M;
      long long out_row = row_t;
      int s2d_col = 0;               // per-row column offset of the space-to-depth remap (parity plane)
      if (p.remap == TDB_REMAP_COMPACT_TO_PADDED) {
        int hw = p.img_h * p.img_w;
        int n = row_t / hw;
        int rem = row_t - n * hw;
        int h = rem / p.img_w;
        int x = rem - h * p.img_w;
        out_row = ((long long)n * Hp + h + 1) * Wp + x + 1;
      } else if (p.remap == TDB_REMAP_PADDED_TO_COMPACT) {
        int hw = Hp * Wp;
        int n = row_t / hw;
        int rem = row_t - n * hw;
        int h = rem / Wp;
        int x = rem - h * Wp;
        valid = valid && h >= 1 && h <= p.img_h && x >= 1 && x <= p.img_w;
        out_row = ((long long)n * p.img_h + (h - 1)) * p.img_w + (x - 1);
      } else if (p.remap == TDB_REMAP_COMPACT_TO_S2D) {
        const int hw = p.img_h * p.img_w, ohp = ((p.img_h + 1) >> 1) + 1, owp = ((p.img_w + 1) >> 1) + 1;
        const int n = row_t / hw, rem = row_t - n * hw;
        const int h = rem / p.img_w, x = rem - h * p.img_w;
        out_row = ((long long)n * ohp + (h >> 1) + 1) * owp + (x >> 1) + 1;
        s2d_col = ((h & 1) * 2 + (x & 1)) * p.N;
      } else if (p.remap == TDB_REMAP_COMPACT_TO_PADDED1) {
        const int hw = p.img_h * p.img_w;
        const int n = row_t / hw, rem = row_t - n * hw;
        const int h = rem / p.img_w, x = rem - h * p.img_w;
        out_row = ((long long)n * (p.img_h + 1) + h + 1) * (p.img_w + 1) + x + 1;
      } else if (p.remap == TDB_REMAP_S2D_TO_COMPACT) {
        const int ohp = p.img_h + 1, owp = p.img_w + 1;
        const int n = row_t / (ohp * owp), rem = row_t - n * (ohp * owp);
        const int i = rem / owp, j = rem - i * owp;
        valid = valid && i >= 1 && j >= 1;
        out_row = ((long long)n * p.img_h + (i - 1)) * p.img_w + (j - 1);
      }
      int out_col0 = wi.n0 + p.z_out_col[wi.z] + s2d_col;
      if (p.splits > 1) out_row += (long long)wi.split * p.M;

      // mode 2: software-pipelined register prefetch of the row-wise epilogue operands (residual, ReLU mask).  The first
      // chunks are requested BEFORE the accumulator is ready, so their HBM latency hides behind this tile's main loop.
      constexpr int NCH = BN / 32;
      constexpr int PFR = NCH < 4 ? NCH : 4;   // residual prefetch depth (chunks of 32 columns = 4 x 16 B per thread)
      constexpr int PFM = NCH < 2 ? NCH : 2;   // mask prefetch depth
      uint4 rbuf[PFR][4], mbuf[PFM][4];
      const bf16* res_row = p.residual ? p.residual + out_row * p.ldr + wi.n0 : nullptr;
      const bf16* msk_row = p.mask ? p.mask + (long long)row_t * p.ldmask + wi.n0 : nullptr;
      float* ssc = stage;            // mode 3: per-warp private copy of this tile's scale / bias columns (L1 has no capacity
      float* sbi = stage + BN;       // next to ~226 KB of shared memory, so per-chunk __ldg would each pay an L2 round trip)
      if constexpr (EPI == 5) {
        // scale / bias of this tile, one copy shared by the 4 epilogue warps (named barrier 1 = the 128 epilogue threads)
        asm volatile("bar.sync 1, 128;" ::: "memory");      // previous tile's readers are done with the copy
        const int t = threadIdx.x - 128;
        if (t < BN) {
          ssc[t] = p.scale ? __ldg(p.scale + wi.n0 + t) : 1.f;
          sbi[t] = p.bias ? __ldg(p.bias + wi.n0 + t) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if constexpr (EPI == 3 || EPI == 4) {
        for (int i = lane; i < BN; i += 32) {
          ssc[i] = p.scale ? __ldg(p.scale + wi.n0 + i) : 1.f;
          sbi[i] = p.bias ? __ldg(p.bias + wi.n0 + i) : 0.f;
        }
        __syncwarp();
      }
      if constexpr (EPI == 2 || EPI == 3 || EPI == 4 || EPI == 5) {
        if (EPI != 4 && EPI != 5 && valid && res_row) {
#pragma unroll
          for (int ci = 0; ci < PFR; ++ci)
#pragma unroll
            for (int i = 0; i < 4; ++i) rbuf[ci][i] = __ldg(reinterpret_cast<const uint4*>(res_row + ci * 32) + i);
        }
        if (valid && msk_row) {
#pragma unroll
          for (int ci = 0; ci < PFM; ++ci)
#pragma unroll
            for (int i = 0; i < 4; ++i) mbuf[ci][i] = __ldg(reinterpret_cast<const uint4*>(msk_row + ci * 32) + i);
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(wq * 32) << 16);
      if constexpr (EPI == 5) {
        // TMA in, TMA out: thread r owns row r of the 128 x 128 tile in shared memory.  It reads the residual there (if any),
        // overwrites it IN PLACE with the bf16 output, and one thread hands the finished tile to the TMA store engine, so no
        // epilogue thread issues a global load or store.  The slot is recycled once the store has finished reading it.
        if (p.residual != nullptr) mbar_wait(&ofull_bar[oslot], ophase, 8);
        const int rloc = wq * 32 + lane;
        uint8_t* obase = opnd + oslot * (BN / 64) * (BM * 128) + rloc * 128;
        uint32_t r[2][32];
        tmem_ld_32x32(taddr, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = ci * 32;
          tmem_ld_wait();
          if (ci + 1 < NCH) tmem_ld_32x32(taddr + c + 32, r[(ci + 1) & 1]);
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 s4 = *reinterpret_cast<const float4*>(ssc + c + 4 * i);
            const float4 b4 = *reinterpret_cast<const float4*>(sbi + c + 4 * i);
            v[4 * i] = fmaf(__uint_as_float(r[ci & 1][4 * i]), s4.x, b4.x);
            v[4 * i + 1] = fmaf(__uint_as_float(r[ci & 1][4 * i + 1]), s4.y, b4.y);
            v[4 * i + 2] = fmaf(__uint_as_float(r[ci & 1][4 * i + 2]), s4.z, b4.z);
            v[4 * i + 3] = fmaf(__uint_as_float(r[ci & 1][4 * i + 3]), s4.w, b4.w);
          }
          uint8_t* box = obase + (ci >> 1) * (BM * 128);
          if (p.residual != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j16 = (ci & 1) * 4 + i;
              uint4 u = *reinterpret_cast<const uint4*>(box + ((j16 ^ (rloc & 7)) << 4));
              float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              v[8 * i] += f0.x; v[8 * i + 1] += f0.y; v[8 * i + 2] += f1.x; v[8 * i + 3] += f1.y;
              v[8 * i + 4] += f2.x; v[8 * i + 5] += f2.y; v[8 * i + 6] += f3.x; v[8 * i + 7] += f3.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (msk_row != nullptr && valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u = mbuf[ci % PFM][i];
              float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              v[8 * i] = f0.x > 0.f ? v[8 * i] : 0.f;         v[8 * i + 1] = f0.y > 0.f ? v[8 * i + 1] : 0.f;
              v[8 * i + 2] = f1.x > 0.f ? v[8 * i + 2] : 0.f; v[8 * i + 3] = f1.y > 0.f ? v[8 * i + 3] : 0.f;
              v[8 * i + 4] = f2.x > 0.f ? v[8 * i + 4] : 0.f; v[8 * i + 5] = f2.y > 0.f ? v[8 * i + 5] : 0.f;
              v[8 * i + 6] = f3.x > 0.f ? v[8 * i + 6] : 0.f; v[8 * i + 7] = f3.y > 0.f ? v[8 * i + 7] : 0.f;
            }
            if (ci + PFM < NCH) {
#pragma unroll
              for (int i = 0; i < 4; ++i) mbuf[ci % PFM][i] = __ldg(reinterpret_cast<const uint4*>(msk_row + (ci + PFM) * 32) + i);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j16 = (ci & 1) * 4 + i;
            *reinterpret_cast<uint4*>(box + ((j16 ^ (rloc & 7)) << 4)) =
                make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
        }
        // accumulator fully read: hand TMEM back to the MMA warp before the store hand-off
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        const bool issuer = (warp == 4 && lane == 0);
        if (issuer) {
          // the store issued one tile ago has finished reading its slot -> recycle that slot for the producer
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (tiles_done > 0) mbar_arrive(&oempty_bar[(oslot + 2) % 3]);
        }
        fence_proxy_async();                                   // my smem writes -> visible to the TMA engine
        asm volatile("bar.sync 1, 128;" ::: "memory");         // whole tile written (and the previous store drained)
        if (issuer) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmO),
                         "r"(smem_u32(opnd + (oslot * (BN / 64) + j) * (BM * 128))), "r"(out_col0 + j * 64), "r"(wi.m0)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++tiles_done;
        if (++oslot == 3) {
          oslot = 0;
          ophase ^= 1;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      } else if constexpr (EPI == 4) {
        // residual tile arrives through TMA (128-byte swizzled boxes of 64 columns); thread r reads its own row: the swizzle
        // makes 8 consecutive rows hit 8 different 16-byte bank groups, so the row-per-thread reads are conflict free
        mbar_wait(&ofull_bar[oslot], ophase, 6);
        const int rloc = wq * 32 + lane;
        const uint8_t* obase = opnd + oslot * (BN / 64) * (BM * 128) + rloc * 128;
        uint32_t r[2][32];
        tmem_ld_32x32(taddr, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = ci * 32;
          tmem_ld_wait();
          if (ci + 1 < NCH) tmem_ld_32x32(taddr + c + 32, r[(ci + 1) & 1]);
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 s4 = *reinterpret_cast<const float4*>(ssc + c + 4 * i);
              const float4 b4 = *reinterpret_cast<const float4*>(sbi + c + 4 * i);
              v[4 * i] = fmaf(__uint_as_float(r[ci & 1][4 * i]), s4.x, b4.x);
              v[4 * i + 1] = fmaf(__uint_as_float(r[ci & 1][4 * i + 1]), s4.y, b4.y);
              v[4 * i + 2] = fmaf(__uint_as_float(r[ci & 1][4 * i + 2]), s4.z, b4.z);
              v[4 * i + 3] = fmaf(__uint_as_float(r[ci & 1][4 * i + 3]), s4.w, b4.w);
            }
            const uint8_t* box = obase + (ci >> 1) * (BM * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j16 = (ci & 1) * 4 + i;                   // 16-byte chunk within the 128-byte row
              uint4 u = *reinterpret_cast<const uint4*>(box + ((j16 ^ (rloc & 7)) << 4));
              float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              v[8 * i] += f0.x; v[8 * i + 1] += f0.y; v[8 * i + 2] += f1.x; v[8 * i + 3] += f1.y;
              v[8 * i + 4] += f2.x; v[8 * i + 5] += f2.y; v[8 * i + 6] += f3.x; v[8 * i + 7] += f3.y;
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (msk_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = mbuf[ci % PFM][i];
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] = f0.x > 0.f ? v[8 * i] : 0.f;         v[8 * i + 1] = f0.y > 0.f ? v[8 * i + 1] : 0.f;
                v[8 * i + 2] = f1.x > 0.f ? v[8 * i + 2] : 0.f; v[8 * i + 3] = f1.y > 0.f ? v[8 * i + 3] : 0.f;
                v[8 * i + 4] = f2.x > 0.f ? v[8 * i + 4] : 0.f; v[8 * i + 5] = f2.y > 0.f ? v[8 * i + 5] : 0.f;
                v[8 * i + 6] = f3.x > 0.f ? v[8 * i + 6] : 0.f; v[8 * i + 7] = f3.y > 0.f ? v[8 * i + 7] : 0.f;
              }
              if (ci + PFM < NCH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mbuf[ci % PFM][i] = __ldg(reinterpret_cast<const uint4*>(msk_row + (ci + PFM) * 32) + i);
              }
            }
            if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&oempty_bar[oslot]);
        if (++oslot == 2) {
          oslot = 0;
          ophase ^= 1;
        }
      } else if constexpr (EPI == 3) {
        // software pipeline: the TMEM load of chunk ci+1 is in flight while chunk ci is processed
        uint32_t r[2][32];
        const bool dbg_nold = (p.debug_flags & 32) != 0, dbg_nost = (p.debug_flags & 16) != 0;
        if (!dbg_nold) tmem_ld_32x32(taddr, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = ci * 32;
          tmem_ld_wait();
          if (ci + 1 < NCH && !dbg_nold) tmem_ld_32x32(taddr + c + 32, r[(ci + 1) & 1]);
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 s4 = *reinterpret_cast<const float4*>(ssc + c + 4 * i);
              const float4 b4 = *reinterpret_cast<const float4*>(sbi + c + 4 * i);
              v[4 * i] = fmaf(__uint_as_float(r[ci & 1][4 * i]), s4.x, b4.x);
              v[4 * i + 1] = fmaf(__uint_as_float(r[ci & 1][4 * i + 1]), s4.y, b4.y);
              v[4 * i + 2] = fmaf(__uint_as_float(r[ci & 1][4 * i + 2]), s4.z, b4.z);
              v[4 * i + 3] = fmaf(__uint_as_float(r[ci & 1][4 * i + 3]), s4.w, b4.w);
            }
            if (res_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = rbuf[ci % PFR][i];
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] += f0.x; v[8 * i + 1] += f0.y; v[8 * i + 2] += f1.x; v[8 * i + 3] += f1.y;
                v[8 * i + 4] += f2.x; v[8 * i + 5] += f2.y; v[8 * i + 6] += f3.x; v[8 * i + 7] += f3.y;
              }
              if (ci + PFR < NCH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rbuf[ci % PFR][i] = __ldg(reinterpret_cast<const uint4*>(res_row + (ci + PFR) * 32) + i);
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (msk_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = mbuf[ci % PFM][i];
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] = f0.x > 0.f ? v[8 * i] : 0.f;         v[8 * i + 1] = f0.y > 0.f ? v[8 * i + 1] : 0.f;
                v[8 * i + 2] = f1.x > 0.f ? v[8 * i + 2] : 0.f; v[8 * i + 3] = f1.y > 0.f ? v[8 * i + 3] : 0.f;
                v[8 * i + 4] = f2.x > 0.f ? v[8 * i + 4] : 0.f; v[8 * i + 5] = f2.y > 0.f ? v[8 * i + 5] : 0.f;
                v[8 * i + 6] = f3.x > 0.f ? v[8 * i + 6] : 0.f; v[8 * i + 7] = f3.y > 0.f ? v[8 * i + 7] : 0.f;
              }
              if (ci + PFM < NCH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mbuf[ci % PFM][i] = __ldg(reinterpret_cast<const uint4*>(msk_row + (ci + PFM) * 32) + i);
              }
            }
            if (dbg_nost) {
              if (v[0] == 123456.f) reinterpret_cast<float*>(p.out)[0] = v[1];   // keep the math alive
            } else if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            }
          }
        }
      } else if constexpr (EPI == 2) {
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = ci * 32;
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            const int col = wi.n0 + c;
            if (p.scale != nullptr) {
              const float4* sp = reinterpret_cast<const float4*>(p.scale + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 s = __ldg(sp + i);
                v[4 * i] *= s.x; v[4 * i + 1] *= s.y; v[4 * i + 2] *= s.z; v[4 * i + 3] *= s.w;
              }
            }
            if (p.bias != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 s = __ldg(bp + i);
                v[4 * i] += s.x; v[4 * i + 1] += s.y; v[4 * i + 2] += s.z; v[4 * i + 3] += s.w;
              }
            }
            if (res_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = rbuf[ci % PFR][i];
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] += f0.x; v[8 * i + 1] += f0.y; v[8 * i + 2] += f1.x; v[8 * i + 3] += f1.y;
                v[8 * i + 4] += f2.x; v[8 * i + 5] += f2.y; v[8 * i + 6] += f3.x; v[8 * i + 7] += f3.y;
              }
              if (ci + PFR < NCH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rbuf[ci % PFR][i] = __ldg(reinterpret_cast<const uint4*>(res_row + (ci + PFR) * 32) + i);
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (msk_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = mbuf[ci % PFM][i];
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] = f0.x > 0.f ? v[8 * i] : 0.f;         v[8 * i + 1] = f0.y > 0.f ? v[8 * i + 1] : 0.f;
                v[8 * i + 2] = f1.x > 0.f ? v[8 * i + 2] : 0.f; v[8 * i + 3] = f1.y > 0.f ? v[8 * i + 3] : 0.f;
                v[8 * i + 4] = f2.x > 0.f ? v[8 * i + 4] : 0.f; v[8 * i + 5] = f2.y > 0.f ? v[8 * i + 5] : 0.f;
                v[8 * i + 6] = f3.x > 0.f ? v[8 * i + 6] : 0.f; v[8 * i + 7] = f3.y > 0.f ? v[8 * i + 7] : 0.f;
              }
              if (ci + PFM < NCH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mbuf[ci % PFM][i] = __ldg(reinterpret_cast<const uint4*>(msk_row + (ci + PFM) * 32) + i);
              }
            }
            if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            }
          }
        }
      } else if constexpr (EPI == 0) {
        // ---- mode 0: thread-per-row, registers -> global (16-byte stores, row-strided)
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            const int col = wi.n0 + c;  // column in scale/bias/mask/residual space
            if (p.scale != nullptr) {
              const float4* sp = reinterpret_cast<const float4*>(p.scale + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 s = __ldg(sp + i);
                v[4 * i] *= s.x; v[4 * i + 1] *= s.y; v[4 * i + 2] *= s.z; v[4 * i + 3] *= s.w;
              }
            }
            if (p.bias != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 s = __ldg(bp + i);
                v[4 * i] += s.x; v[4 * i + 1] += s.y; v[4 * i + 2] += s.z; v[4 * i + 3] += s.w;
              }
            }
            if (p.residual != nullptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + out_row * p.ldr + col);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = __ldg(rp + i);
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] += f0.x; v[8 * i + 1] += f0.y; v[8 * i + 2] += f1.x; v[8 * i + 3] += f1.y;
                v[8 * i + 4] += f2.x; v[8 * i + 5] += f2.y; v[8 * i + 6] += f3.x; v[8 * i + 7] += f3.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (p.mask != nullptr) {
              const uint4* mp = reinterpret_cast<const uint4*>(p.mask + (long long)row_t * p.ldmask + col);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = __ldg(mp + i);
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[8 * i] = f0.x > 0.f ? v[8 * i] : 0.f;         v[8 * i + 1] = f0.y > 0.f ? v[8 * i + 1] : 0.f;
                v[8 * i + 2] = f1.x > 0.f ? v[8 * i + 2] : 0.f; v[8 * i + 3] = f1.y > 0.f ? v[8 * i + 3] : 0.f;
                v[8 * i + 4] = f2.x > 0.f ? v[8 * i + 4] : 0.f; v[8 * i + 5] = f2.y > 0.f ? v[8 * i + 5] : 0.f;
                v[8 * i + 6] = f3.x > 0.f ? v[8 * i + 6] : 0.f; v[8 * i + 7] = f3.y > 0.f ? v[8 * i + 7] : 0.f;
              }
            }
            if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + out_row * p.ldo + out_col0 + c);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            }
          }
        }
      } else {
        // ---- mode 1: each warp transposes its own 32 rows x 64 columns through a private XOR-swizzled smem tile so that
        // global traffic is row-contiguous: phase 1 = thread-per-row TMEM -> smem (float4, conflict-free), phase 2 =
        // quarter-warp-per-row: 8 lanes x 8 columns = 128 B (bf16) / 256 B (fp32) contiguous per row, 4 rows per instruction.
        // The residual / mask operands of a chunk are prefetched (8 independent 16-byte loads per lane) before phase 1.
        float4* st4 = reinterpret_cast<float4*>(stage);
        const int hl = lane & 7;         // column group (8 columns) within the 64-column chunk
        const int g = lane >> 3;         // row within each group of 4 rows
        int orow_k[8];                   // output rows (32-bit: row counts stay far below 2^31)
        uint32_t vmask = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int tr = wi.m0 + wq * 32 + 4 * k + g;
          bool ok = tr < p.M;
          long long orow = tr;
          if (p.remap == TDB_REMAP_COMPACT_TO_PADDED) {
            int hw = p.img_h * p.img_w;
            int n = tr / hw;
            int rem = tr - n * hw;
            int h = rem / p.img_w;
            int x = rem - h * p.img_w;
            orow = ((long long)n * Hp + h + 1) * Wp + x + 1;
          } else if (p.remap == TDB_REMAP_PADDED_TO_COMPACT) {
            int hw = Hp * Wp;
            int n = tr / hw;
            int rem = tr - n * hw;
            int h = rem / Wp;
            int x = rem - h * Wp;
            ok = ok && h >= 1 && h <= p.img_h && x >= 1 && x <= p.img_w;
            orow = ((long long)n * p.img_h + (h - 1)) * p.img_w + (x - 1);
          }
          if (p.splits > 1) orow += (long long)wi.split * p.M;
          orow_k[k] = (int)orow;
          vmask |= (ok ? 1u : 0u) << k;
        }
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          const int col = wi.n0 + c + hl * 8;       // column in scale/bias/mask/residual space
          uint4 res[8], msk[8];
          if (p.residual != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              res[k] = ((vmask >> k) & 1) ? __ldg(reinterpret_cast<const uint4*>(p.residual + (long long)orow_k[k] * p.ldr + col))
                                          : make_uint4(0, 0, 0, 0);
          }
          if (p.mask != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              msk[k] = ((vmask >> k) & 1)
                           ? __ldg(reinterpret_cast<const uint4*>(p.mask + (long long)(wi.m0 + wq * 32 + 4 * k + g) * p.ldmask + col))
                           : make_uint4(0, 0, 0, 0);
          }
          float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, bi0 = make_float4(0.f, 0.f, 0.f, 0.f), bi1 = bi0;
          if (p.scale != nullptr) {
            sc0 = __ldg(reinterpret_cast<const float4*>(p.scale + col));
            sc1 = __ldg(reinterpret_cast<const float4*>(p.scale + col + 4));
          }
          if (p.bias != nullptr) {
            bi0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            bi1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st4[lane * 16 + ((half * 8 + q) ^ (lane & 15))] =
                  make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                              __uint_as_float(r[4 * q + 3]));
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if ((vmask >> k) & 1) {
              const int myr = 4 * k + g;
              float4 v0 = st4[myr * 16 + ((2 * hl) ^ (myr & 15))];
              float4 v1 = st4[myr * 16 + ((2 * hl + 1) ^ (myr & 15))];
              v0.x = v0.x * sc0.x + bi0.x; v0.y = v0.y * sc0.y + bi0.y; v0.z = v0.z * sc0.z + bi0.z; v0.w = v0.w * sc0.w + bi0.w;
              v1.x = v1.x * sc1.x + bi1.x; v1.y = v1.y * sc1.y + bi1.y; v1.z = v1.z * sc1.z + bi1.z; v1.w = v1.w * sc1.w + bi1.w;
              if (p.residual != nullptr) {
                float2 f0 = unpack_bf16x2(res[k].x), f1 = unpack_bf16x2(res[k].y), f2 = unpack_bf16x2(res[k].z), f3 = unpack_bf16x2(res[k].w);
                v0.x += f0.x; v0.y += f0.y; v0.z += f1.x; v0.w += f1.y; v1.x += f2.x; v1.y += f2.y; v1.z += f3.x; v1.w += f3.y;
              }
              if (p.relu) {
                v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v0.z = fmaxf(v0.z, 0.f); v0.w = fmaxf(v0.w, 0.f);
                v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f); v1.z = fmaxf(v1.z, 0.f); v1.w = fmaxf(v1.w, 0.f);
              }
              if (p.mask != nullptr) {
                float2 f0 = unpack_bf16x2(msk[k].x), f1 = unpack_bf16x2(msk[k].y), f2 = unpack_bf16x2(msk[k].z), f3 = unpack_bf16x2(msk[k].w);
                v0.x = f0.x > 0.f ? v0.x : 0.f; v0.y = f0.y > 0.f ? v0.y : 0.f; v0.z = f1.x > 0.f ? v0.z : 0.f; v0.w = f1.y > 0.f ? v0.w : 0.f;
                v1.x = f2.x > 0.f ? v1.x : 0.f; v1.y = f2.y > 0.f ? v1.y : 0.f; v1.z = f3.x > 0.f ? v1.z : 0.f; v1.w = f3.y > 0.f ? v1.w : 0.f;
              }
              const long long o = (long long)orow_k[k] * p.ldo + out_col0 + c + hl * 8;
              if (p.out_f32) {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
                op[0] = v0;
                op[1] = v1;
              } else {
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + o) =
                    make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
              }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  if constexpr (EPI == 5) {
    if (warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // last output tile has left shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ split-K reduction (fixed order => deterministic)
// generic scalar form (any N / taps)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N,
                                     const float* __restrict__ rowscale, float* __restrict__ out, int taps,
                                     int accumulate) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)M * N;
  if (idx >= total) return;
  int m = (int)(idx / N);
  int n = (int)(idx - (long long)m * N);
  float s = 0.f;
  for (int i = 0; i < splits; ++i) s += part[(long long)i * total + idx];
  if (rowscale) s *= rowscale[m];
  long long o = idx;
  if (taps > 1) {  // [Cout][tap][Cin] -> torch [Cout][Cin][kh][kw]
    int cin = N / taps;
    int tap = n / cin;
    int c = n - tap * cin;
    o = ((long long)m * cin + c) * taps + tap;
  }
  out[o] = accumulate ? out[o] + s : s;
}

// float4 form (N % 4 == 0 and (N / taps) % 4 == 0): the partials were just written by the wgrad GEMM and mostly sit in L2, so
// this kernel is bound by load latency, not bandwidth -- every thread keeps 4 independent 16-byte loads in flight (the scalar
// form above serialises one 4-byte load per split: measured 10.9 us per launch at 1.2 TB/s, profiles/r01_kernel_table.txt).
// The additions keep the split order 0,1,2,... so results are bit-identical to the scalar form.
__device__ __forceinline__ void add4(float4& s, const float4 v) {
  s.x += v.x;
  s.y += v.y;
  s.z += v.z;
  s.w += v.w;
}
__global__ void __launch_bounds__(256) splitk_reduce4_kernel(const float4* __restrict__ part, int splits, long long total4, int N,
                                                             const float* __restrict__ rowscale, float* __restrict__ out, int taps,
                                                             int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const long long e = idx << 2;
  const int m = (int)(e / N);
  const int n = (int)(e - (long long)m * N);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* p = part + idx;
  int i = 0;
  for (; i + 4 <= splits; i += 4) {
    const float4 a = __ldcs(p + (long long)i * total4);
    const float4 b = __ldcs(p + (long long)(i + 1) * total4);
    const float4 c = __ldcs(p + (long long)(i + 2) * total4);
    const float4 d = __ldcs(p + (long long)(i + 3) * total4);
    add4(s, a);
    add4(s, b);
    add4(s, c);
    add4(s, d);
  }
  for (; i < splits; ++i) add4(s, __ldcs(p + (long long)i * total4));
  if (rowscale) {
    const float r = rowscale[m];
    s.x *= r;
    s.y *= r;
    s.z *= r;
    s.w *= r;
  }
  if (taps > 1) {  // [Cout][tap][Cin] -> torch [Cout][Cin][kh][kw]: 4 consecutive cin of one tap -> stride `taps`
    const int cin = N / taps;
    const int tap = n / cin;
    const int c = n - tap * cin;
    float* o = out + ((long long)m * cin + c) * taps + tap;
    if (accumulate) {
      s.x += o[0];
      s.y += o[taps];
      s.z += o[2 * taps];
      s.w += o[3 * taps];
    }
    o[0] = s.x;
    o[taps] = s.y;
    o[2 * taps] = s.z;
    o[3 * taps] = s.w;
  } else {
    float4* o = reinterpret_cast<float4*>(out + e);
    if (accumulate) add4(s, *o);
    *o = s;
  }
}

}  // namespace tdb

// ====================================================================== host side
static thread_local char g_err[512] = "";
void tdb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void tdb_count_launch(int n) { g_launches += n; }

int tdb_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TDB_PDL"); v = e ? atoi(e) : 1; }
  return v;
}
extern "C" const char* tdb_last_error_string(void) { return g_err; }
extern "C" int tdb_version(void) { return TDB_ABI_VERSION; }
extern "C" int64_t tdb_launch_count(void) { return g_launches.load(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static std::mutex g_init_mu;

int tdb_num_sms() { return g_num_sms; }
int tdb_gemm2_try(const tdb_gemm_desc* d, void* stream_);

int tdb_init_once() {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_encode) return TDB_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  TDB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || fn == nullptr) {
    tdb_set_error("cuTensorMapEncodeTiled not available from the driver");
    return TDB_ERR_DRIVER;
  }
  int dev = 0;
  TDB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TDB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    tdb_set_error("tubedetr_b200 needs an sm_100a device (found sm_%d%d); there is no fallback path", prop.major, prop.minor);
    return TDB_ERR_DRIVER;
  }
  g_num_sms = prop.multiProcessorCount;
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<64>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<64>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<64>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<64>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<128>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<128>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<128>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<128>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<128, 4>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<128, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<128, 5>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<256>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<256>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<256>::kSmemBytes));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::tdb_gemm_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdb::GemmCfg<256>::kSmemBytes));
  g_encode = (EncodeTiledFn)fn;
  return TDB_OK;
}

// row-major bf16 matrix [rows][cols] (leading dim ld elements) -> 2D tensor map with a [box_rows][64] 128B-swizzled box
int tdb_make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0) {
    tdb_set_error("tensor map: base %p / ld %lld not 16-byte aligned", base, (long long)ld);
    return TDB_ERR_ARG;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tdb_set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r, (long long)rows,
                  (long long)cols, (long long)ld, box_rows);
    return TDB_ERR_DRIVER;
  }
  return TDB_OK;
}

static int pick_block_n(int N, long long work_per_bn1 /* m_tiles*nz*splits */, int forced) {
  if (forced == 64 || forced == 128 || forced == 256) return (N % forced == 0) ? forced : 0;
  // Wide tiles amortise the A-operand feed (shared-memory bandwidth bounds the main loop): measured on B200 a 128x256 tile
  // costs ~1.3x a 128x128 tile, not 2x.  So take the widest tile that still occupies at least half of the SMs.
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    int bn = cands[i];
    if (N % bn) continue;
    if (work_per_bn1 * (N / bn) * 2 >= g_num_sms) return bn;
  }
  for (int i = 2; i >= 0; --i)
    if (N % cands[i] == 0) return cands[i];
  return 0;
}

extern "C" int tdb_gemm(const tdb_gemm_desc* d, void* stream_) {
  using namespace tdb;
  int rc = tdb_init_once();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  TDB_REQUIRE(d && d->A && d->B && d->out, "tdb_gemm: null operand");
  TDB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "tdb_gemm: bad shape M=%d N=%d K=%d", d->M, d->N, d->K);
  TDB_REQUIRE(d->N % 64 == 0, "tdb_gemm: N=%d must be a multiple of 64", d->N);
  TDB_REQUIRE(d->ntaps >= 1 && d->ntaps <= TDB_MAX_TAPS, "tdb_gemm: ntaps=%d", d->ntaps);
  const int nz = d->nz < 1 ? 1 : d->nz;
  TDB_REQUIRE(nz <= TDB_MAX_TAPS, "tdb_gemm: nz=%d", nz);
  if (d->a_major == 0 || d->b_major == 0)
    TDB_REQUIRE(d->K % 64 == 0, "tdb_gemm: K=%d must be a multiple of 64 for K-major operands", d->K);
  int splits = d->splits < 1 ? 1 : d->splits;
  const int kb = (d->K + BK - 1) / BK;
  int kb_per_split = kb;
  if (splits > 1) {
    TDB_REQUIRE(d->ntaps == 1 && d->out_dtype == TDB_OUT_F32, "tdb_gemm: split reduction needs ntaps==1 and fp32 partials");
    TDB_REQUIRE(!d->scale && !d->bias && !d->residual && !d->mask && !d->relu && d->remap == 0, "tdb_gemm: no epilogue with splits");
    if (splits > kb) splits = kb;
    kb_per_split = (kb + splits - 1) / splits;
    splits = (kb + kb_per_split - 1) / kb_per_split;
  }
  const int m_tiles = (d->M + BM - 1) / BM;
  int epi_mode;
  {
    static int env_mode = -2;
    if (env_mode == -2) { const char* e = getenv("TDB_EPI_MODE"); env_mode = e ? atoi(e) : -1; }
    int m = (d->debug_flags >> 1) & 7;   // 0 = auto, else mode + 1: 1 direct, 2 smem-staged, 3 +reg prefetch, 4 pipelined, 5 TMA residual
    epi_mode = m ? m - 1 : (env_mode >= 0 ? env_mode : 5);
    // modes 4/5 (residual tile by TMA; 5 also stores the output tile by TMA) need un-remapped rows and 128-wide tiles;
    // otherwise the pipelined register path (3)
    const bool tma_res_ok = d->residual && d->remap == TDB_REMAP_NONE && splits == 1 && d->N % 128 == 0 &&
                            (d->block_n == 0 || d->block_n == 128);
    static int out_all = -1;
    if (out_all < 0) { const char* e = getenv("TDB_TMA_OUT_ALL"); out_all = e ? atoi(e) : 1; }
    // TMA-store epilogue also without a residual (short reductions only: it trades pipeline depth for the output tile)
    const bool plain_ok = out_all && !d->residual && d->remap == TDB_REMAP_NONE && splits == 1 && d->N % 128 == 0 &&
                          (d->block_n == 0 || d->block_n == 128) && (long long)d->K * d->ntaps <= 512 && d->M >= 4096;
    const bool tma_out_ok = (tma_res_ok || plain_ok) && d->out_dtype == TDB_OUT_BF16 && nz == 1;
    if (epi_mode == 5 && !tma_out_ok) epi_mode = 4;
    if (epi_mode == 4 && !tma_res_ok) epi_mode = 3;
  }
  const int bn = (epi_mode == 4 || epi_mode == 5) ? 128 : pick_block_n(d->N, (long long)m_tiles * nz * splits, d->block_n);
  TDB_REQUIRE(bn != 0, "tdb_gemm: no tile width for N=%d (block_n=%d)", d->N, d->block_n);
  if (d->remap != TDB_REMAP_NONE) TDB_REQUIRE(d->img_h > 0 && d->img_w > 0, "tdb_gemm: remap needs img_h/img_w");
  TDB_REQUIRE(d->remap < TDB_REMAP_COMPACT_TO_S2D || (epi_mode != 1 && splits == 1 && !d->residual), "tdb_gemm: space-to-depth remaps: no residual / split / epilogue mode 1");

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.N = d->N; p.kb_per_tap = kb; p.ntaps = d->ntaps;
  p.a_major = d->a_major ? 1 : 0; p.b_major = d->b_major ? 1 : 0;
  for (int i = 0; i < TDB_MAX_TAPS; ++i) {
    p.a_off0[i] = d->a_off0[i]; p.a_off1[i] = d->a_off1[i];
    p.b_off0[i] = d->b_off0[i]; p.b_off1[i] = d->b_off1[i];
    p.z_b_off1[i] = d->nz >= 1 ? d->z_b_off1[i] : 0;
    p.z_out_col[i] = d->nz >= 1 ? d->z_out_col[i] : 0;
  }
  p.nz = nz; p.splits = splits; p.kb_per_split = kb_per_split;
  p.m_tiles = m_tiles; p.n_tiles = d->N / bn;
  long long total = (long long)p.m_tiles * p.n_tiles * nz * splits;
  TDB_REQUIRE(total < (1ll << 30), "tdb_gemm: too many tiles");
  p.total_work = (int)total;
  p.scale = d->scale; p.bias = d->bias;
  p.residual = (const bf16*)d->residual; p.ldr = d->ldr;
  p.mask = (const bf16*)d->mask; p.ldmask = d->ldmask;
  p.relu = d->relu; p.out = d->out; p.out_f32 = d->out_dtype == TDB_OUT_F32; p.ldo = d->ldo;
  p.remap = d->remap; p.img_h = d->img_h; p.img_w = d->img_w;
  p.debug_flags = d->debug_flags;
  p.epi_mode = epi_mode;
  {
    // cluster launch control (see tdb_gemm2.cu for the policy): units of tiles sized to outlast the request's round trip
    static int clc_on = -1;
    if (clc_on < 0) { const char* e = getenv("TDB_CLC"); clc_on = e ? atoi(e) : 1; }
    const bool want_clc = clc_on == 2 || (clc_on == 1 && ((d->debug_flags >> 11) & 1));
    p.clc = 0;
    p.clc_unit = 1;
    if (want_clc && !((d->debug_flags >> 10) & 1) && total > g_num_sms) {
      const double tile_us = (bn == 256 ? 0.27 : (bn == 128 ? 0.14 : 0.09)) * (double)(splits > 1 ? kb_per_split : kb * d->ntaps) + (bn >= 128 ? 1.5 : 1.0);
      int unit = (int)(6.0 / tile_us + 0.999);
      const int cap = (int)(total / (4 * g_num_sms));
      if (unit > cap) unit = cap;
      if (unit < 1) unit = 1;
      p.clc = 1;
      p.clc_unit = unit;
    }
  }
  TDB_REQUIRE(p.ldo % 8 == 0 && (!p.residual || p.ldr % 8 == 0) && (!p.mask || p.ldmask % 8 == 0), "tdb_gemm: leading dims must be multiples of 8");
  {
    int r2 = tdb_gemm2_try(d, stream_);   // 2-CTA (cta_group::2) path for the feed-bound deep-K shapes
    if (r2 < 0) return r2;
    if (r2 == 1) return TDB_OK;
  }

  // halo mode of the implicit 3x3 convolution (K-major A, taps = row shifts): pipelined-register epilogue (3) only
  int a_box = BM;
  {
    static int halo_ok = -1;
    if (halo_ok < 0) { const char* e = getenv("TDB_GEMM_HALO"); halo_ok = e ? atoi(e) : 1; }
    if (halo_ok && !((d->debug_flags >> 7) & 1) && (p.epi_mode == 3 || p.epi_mode == 1) && p.a_major == 0 && d->ntaps > 1 && splits == 1 && nz == 1) {
      int lo = d->a_off1[0], hi = d->a_off1[0];
      bool same_cols = true;
      for (int i = 1; i < d->ntaps; ++i) {
        lo = d->a_off1[i] < lo ? d->a_off1[i] : lo;
        hi = d->a_off1[i] > hi ? d->a_off1[i] : hi;
        same_cols = same_cols && d->a_off0[i] == d->a_off0[0];
      }
      const int rows = BM + hi - lo;
      const int nbox = (rows + 255) / 256;
      const int box_rows = (((rows + nbox - 1) / nbox) + 7) & ~7;
      const int tile_bytes = nbox * box_rows * 128;
      const int ring = (bn == 256 ? 4 * 49152 : (bn == 128 ? 6 * 32768 : 8 * 24576));   // GemmCfg<bn,3>::kStages * kStageBytes
      const int bstage = bn * 128;
      const int max_stages = bn == 256 ? 4 : (bn == 128 ? 6 : 8);
      // the weights stay resident when every (k-block, tap) slice fits next to two A tiles; the A ring then takes the rest.
      // Otherwise: as many A tiles in flight as leave >= 4 weight stages (A tiles are what hides the load latency: the next
      // tile's rows can only be requested once a buffer is free)
      static int na_env = -1;
      if (na_env < 0) { const char* e = getenv("TDB_HALO_A_STAGES"); na_env = e ? atoi(e) : 0; }
      const int slices = kb * d->ntaps;
      const bool resident = p.n_tiles == 1 && slices * bstage + 2 * tile_bytes <= ring && slices * bstage < (1 << 20);
      int na, nb;
      if (resident) {
        nb = slices;
        na = (ring - slices * bstage) / tile_bytes;
      } else {
        na = (ring - 4 * bstage) / tile_bytes;
        if (na > 3) na = 3;
        if (na < 2) na = 2;
        nb = (ring - na * tile_bytes) / bstage;
      }
      if (na > 4) na = 4;
      if (na_env >= 2 && na_env <= na) na = na_env;
      if (!resident) nb = (ring - na * tile_bytes) / bstage;
      // measured (tools/halo_diag.py): with streamed weights the 1-CTA kernel gains nothing from the halo tile (the weight
      // stream, 9 x BN x 128 B per k-block, is the feed either way); with resident weights it does
      static int halo_ring = -1;
      if (halo_ring < 0) { const char* e = getenv("TDB_HALO_RING"); halo_ring = e ? atoi(e) : 0; }
      if (same_cols && nb >= 3 && na >= 2 && (resident || halo_ring)) {
        p.halo = 1; p.a_min_off = lo; p.a_box_rows = box_rows; p.a_nbox = nbox; p.a_tile_bytes = tile_bytes;
        p.na_stages = na;
        p.nb_stages = nb > max_stages ? max_stages : nb;
        p.b_resident = resident ? 1 : 0;
        a_box = box_rows;
      }
    }
  }
  CUtensorMap tmA, tmB;
  rc = tdb_make_tmap_bf16(&tmA, d->A, d->a_rows, d->a_cols, d->lda, p.a_major ? 64 : a_box);
  if (rc) return rc;
  rc = tdb_make_tmap_bf16(&tmB, d->B, d->b_rows, d->b_cols, d->ldb, p.b_major ? 64 : bn);
  if (rc) return rc;
  CUtensorMap tmR = tmA, tmO = tmA;   // only dereferenced in epilogue modes 4 / 5
  if ((p.epi_mode == 4 || p.epi_mode == 5) && d->residual) {
    rc = tdb_make_tmap_bf16(&tmR, d->residual, d->M, d->N, d->ldr, BM);
    if (rc) return rc;
  }
  if (p.epi_mode == 5) {
    rc = tdb_make_tmap_bf16(&tmO, d->out, d->M, d->N, d->ldo, BM);
    if (rc) return rc;
  }

  // Wave quantisation: a persistent grid of g CTAs finishes in ceil(tiles/g) tile-times.  When the last wave would be less
  // than half full, the launch is split: full waves with the chosen tile width, and the remaining M tiles with narrower
  // (faster) tiles spread over more CTAs.  (e.g. layer3 3x3 conv on 100 frames: 450 tiles on 148 SMs = 3.04 -> 3 + 0.25 waves.)
  auto launch = [&](int bn_, const CUtensorMap& tmB_, const GemmKParams& q) -> int {
    int grid = q.total_work < g_num_sms ? q.total_work : g_num_sms;
    if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
    if (q.clc) grid = (q.total_work + q.clc_unit - 1) / q.clc_unit;    // one CTA per unit; resident CTAs steal the units of CTAs that have not started
#define TDB_LAUNCH(BN_, EPI_) TDB_CHECK_CUDA(tdb_launch(tdb_gemm_kernel<BN_, EPI_>, dim3(grid), dim3(kGemmThreads), GemmCfg<BN_, EPI_>::kSmemBytes, stream, tmA, tmB_, tmR, tmO, q))
    if (q.epi_mode == 0) {
      switch (bn_) {
        case 64: TDB_LAUNCH(64, 0); break;
        case 128: TDB_LAUNCH(128, 0); break;
        default: TDB_LAUNCH(256, 0); break;
      }
    } else if (q.epi_mode == 1) {
      switch (bn_) {
        case 64: TDB_LAUNCH(64, 1); break;
        case 128: TDB_LAUNCH(128, 1); break;
        default: TDB_LAUNCH(256, 1); break;
      }
    } else if (q.epi_mode == 2) {
      switch (bn_) {
        case 64: TDB_LAUNCH(64, 2); break;
        case 128: TDB_LAUNCH(128, 2); break;
        default: TDB_LAUNCH(256, 2); break;
      }
    } else if (q.epi_mode == 5 && bn_ == 128) {
      TDB_LAUNCH(128, 5);
    } else if (q.epi_mode == 4 && bn_ == 128) {
      TDB_LAUNCH(128, 4);
    } else {
      switch (bn_) {
        case 64: TDB_LAUNCH(64, 3); break;
        case 128: TDB_LAUNCH(128, 3); break;
        default: TDB_LAUNCH(256, 3); break;
      }
    }
#undef TDB_LAUNCH
    tdb_count_launch(1);
    return grid;
  };
  int grid = 0, tail_m = 0, tail_bn = 0;
  const int g = (d->max_ctas > 0 && d->max_ctas < g_num_sms) ? d->max_ctas : g_num_sms;
  static int tail_split = -1;
  if (tail_split < 0) { const char* e = getenv("TDB_TAIL_SPLIT"); tail_split = e ? atoi(e) : 0; }  // measured: no gain (r01), off
  if (tail_split && !p.halo && splits == 1 && nz == 1 && bn > 64 && p.total_work > g && d->block_n == 0) {
    const int full = (p.total_work / g) * g;
    const int rem = p.total_work - full;
    const int m_main = full / p.n_tiles;                 // whole M tiles covered by the full waves
    if (rem > 0 && rem * 2 <= g && m_main > 0 && m_main < p.m_tiles) {
      tail_m = p.m_tiles - m_main;
      tail_bn = bn / 4 < 64 ? 64 : bn / 4;
      GemmKParams pm = p;
      pm.m_tiles = m_main;
      pm.total_work = m_main * p.n_tiles;
      grid = launch(bn, tmB, pm);
      GemmKParams pt = p;
      pt.m_tile0 = m_main;
      pt.m_tiles = tail_m;
      pt.n_tiles = d->N / tail_bn;
      pt.total_work = pt.m_tiles * pt.n_tiles;
      CUtensorMap tmBt;
      rc = tdb_make_tmap_bf16(&tmBt, d->B, d->b_rows, d->b_cols, d->ldb, p.b_major ? 64 : tail_bn);
      if (rc) return rc;
      launch(tail_bn, tmBt, pt);
    }
  }
  if (grid == 0) grid = launch(bn, tmB, p);
  TDB_CHECK_CUDA(cudaGetLastError());
  static FILE* logf = nullptr;
  static bool log_checked = false;
  if (!log_checked) {  // TDB_GEMM_LOG=<path>: one line per launch (profiling aid, joins with the ncu launch list)
    log_checked = true;
    const char* lp = getenv("TDB_GEMM_LOG");
    if (lp) logf = fopen(lp, "w");
  }
  if (logf) {
    fprintf(logf, "M=%d N=%d K=%d taps=%d nz=%d splits=%d bn=%d amaj=%d bmaj=%d grid=%d epi=%d%d%d%d remap=%d f32=%d tail=%dx%d\n", d->M, d->N,
            d->K, d->ntaps, nz, splits, bn, p.a_major, p.b_major, grid, d->scale != nullptr, d->residual != nullptr, d->relu,
            d->mask != nullptr, d->remap, p.out_f32, tail_m, tail_bn);
    if (tail_m) fprintf(logf, "M=%d N=%d K=%d taps=%d nz=%d splits=%d bn=%d amaj=%d bmaj=%d grid=%d epi=%d%d%d%d remap=%d f32=%d tail=-1x0\n", tail_m * 128, d->N,
            d->K, d->ntaps, nz, splits, tail_bn, p.a_major, p.b_major, grid, d->scale != nullptr, d->residual != nullptr, d->relu,
            d->mask != nullptr, d->remap, p.out_f32);
    fflush(logf);
  }
  return TDB_OK;
}

extern "C" int tdb_gemm_effective_splits(int K, int splits) {
  const int kb = (K + tdb::BK - 1) / tdb::BK;
  if (splits < 1) splits = 1;
  if (splits > kb) splits = kb;
  const int per = (kb + splits - 1) / splits;
  return (kb + per - 1) / per;
}

extern "C" int tdb_splitk_reduce(const float* part, int splits, int M, int N, const float* rowscale, float* out,
                                 int taps, int accumulate, void* stream_) {
  TDB_REQUIRE(part && out && splits >= 1 && M > 0 && N > 0 && taps >= 1 && N % taps == 0, "tdb_splitk_reduce: bad args");
  long long total = (long long)M * N;
  int threads = 256;
  const bool vec = N % 4 == 0 && (N / taps) % 4 == 0 && (((uintptr_t)part | (uintptr_t)out) & 15) == 0;
  if (vec) {
    long long total4 = total >> 2;
    long long blocks = (total4 + threads - 1) / threads;
    TDB_CHECK_CUDA(tdb_launch(tdb::splitk_reduce4_kernel, dim3((unsigned)blocks), dim3(threads), 0, (cudaStream_t)stream_, (const float4*)part, splits, total4, N, rowscale, out, taps, accumulate));
  } else {
    long long blocks = (total + threads - 1) / threads;
    TDB_CHECK_CUDA(tdb_launch(tdb::splitk_reduce_kernel, dim3((unsigned)blocks), dim3(threads), 0, (cudaStream_t)stream_, part, splits, M, N, rowscale, out, taps, accumulate));
  }
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}
