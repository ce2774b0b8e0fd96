// tubedetr_b200 -- 2-CTA (cta_group::2) variant of the tcgen05 GEMM for the compute-bound convolutions.
//
// Why: with one CTA per tile the 128 x 256 x 64 k-block needs 48 KB from L2 for 4 x 128 MMA cycles; measured on B200 the
// main loop then runs at ~52 % tensor-pipe activity (profiles/r01_ncu_gemm.txt) because the L2 -> SM feed, not the MMA, is the
// limit.  A CTA pair computes a 256 x 256 tile with ONE tcgen05.mma.cta_group::2 per k-step: each CTA loads its own 128 rows
// of A and only HALF of B (128 of the 256 weight rows); the tensor cores of both SMs read both halves.  Per SM that is 32 KB
// per k-block instead of 48 KB for the same MMA time.
// Roles per CTA: warp 0 = TMA producer (both CTAs; all loads signal the LEADER's full barrier), warp 1 = MMA issuer (leader
// CTA only; commits are multicast to both CTAs), warp 2 = TMEM allocator, warps 4-7 = epilogue of this CTA's 128 rows.
// Supports K- and MN-major operands (fprop / dgrad / wgrad), implicit-conv taps, z-batched wgrad taps, split reductions,
// row remaps and the scale/bias/residual/ReLU/mask epilogue.
#include <stdlib.h>
#include <string.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

namespace tdb {

constexpr int G2_BN = 256;
constexpr int G2_STAGE_BYTES = 128 * 64 * 2 * 2;   // A 16 KB + half of B 16 KB
constexpr int G2_STAGES = 6;
constexpr int G2_THREADS = 256;
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + 1024 + 256 + 4 * 2048 + 256;
// halo mode (implicit 3x3 convolution with a K-major A operand): the 9 taps of one k-block are row-shifted views of the SAME
// pixel rows, so A is loaded ONCE per k-block as rows [m0 + min_off, m0 + 127 + max_off] (128-byte swizzled, 128 B per row) and
// each tap's MMA descriptor starts `a_off1[tap] - min_off` rows into that tile.  The 128-byte swizzle is a function of the
// absolute shared-memory address (measured: a start address that is not 1024-byte aligned needs NO descriptor base offset --
// setting one gives wrong products; tools/halo_diag.py), so a row-shifted start address reads exactly the shifted rows.  Per k-block a CTA then pulls ~1.4 x 16 KB of A instead of 9 x 16 KB through the L2 -> SM crossbar,
// which is what bounds this kernel (profiles/r01_ncu_gemm_final.txt).  A ring: 2 tiles; B ring: up to 8 x 16 KB.
constexpr int G2_HALO_A_STAGES = 2;
constexpr int G2_HALO_B_MAX = 8;
constexpr int G2_RING_BYTES = G2_STAGES * G2_STAGE_BYTES;
// TMA-epilogue variant (tdb_gemm2_kernel<1>): bf16 outputs without a row remap leave -- and residual tiles enter -- through shared
// memory boxes of 128 rows x 64 columns (16 KB, 128-byte swizzled) moved by the TMA engine, so no epilogue thread issues a global
// load or store: a row-per-thread store writes 32 half-used sectors per instruction, which is what bounded the short-K shapes (1x1
// conv + FrozenBN + residual, K/V projection of the decoder) on this kernel.  13 x 16 KB of data: `nstages` operand stages of 32 KB
// followed by `nslots` boxes (residual: 3 + 7, the residual producer runs up to 7 boxes = 1.75 tiles ahead; plain: 5 + 3).
constexpr int G2T_DATA_BYTES = 13 * 16384;
constexpr int G2T_BOX_BYTES = 128 * 64 * 2;
constexpr int G2T_MAX_SLOTS = 8;
constexpr int G2T_SMEM = G2T_DATA_BYTES + 1024 + 512 + 2 * G2_BN * 4 + 128 * 4 + 256;

struct Gemm2Params {
  int M, N;
  int kb_per_tap, ntaps;
  int a_major, b_major;
  int nz, splits, kb_per_split;
  int z_b_off1[TDB_MAX_TAPS], z_out_col[TDB_MAX_TAPS];
  int a_off0[TDB_MAX_TAPS], a_off1[TDB_MAX_TAPS], b_off0[TDB_MAX_TAPS], b_off1[TDB_MAX_TAPS];
  int m_tiles, n_tiles, total_work;   // tiles of 256 x 256
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
  // halo mode
  int halo, a_min_off, a_box_rows, a_nbox, a_tile_bytes, nb_stages, use_base_offset;
  // TMA epilogue
  int nstages, nslots;
  int clc, clc_unit;     // 1: one cluster per `clc_unit` tiles + cluster launch control (work stealing) instead of the static persistent schedule
};

struct Work2 {
  int mt, nt, z, split, kb_begin, iters;
};
__device__ __forceinline__ Work2 decode2(const Gemm2Params& p, int w) {
  Work2 it;
  it.nt = w % p.n_tiles;
  int r = w / p.n_tiles;
  it.mt = r % p.m_tiles;
  r /= p.m_tiles;
  it.split = r % p.splits;
  it.z = r / p.splits;
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

template <int TEPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
tdb_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET on the __shared__ pointer: an integer round trip would turn every access below into a
  // generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* after = smem + (TEPI ? G2T_DATA_BYTES : G2_STAGES * G2_STAGE_BYTES);
  const int nst = TEPI ? p.nstages : G2_STAGES;      // operand ring length
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);          // [8] (halo mode: B ring)
  uint64_t* empty_bar = full_bar + G2_HALO_B_MAX;
  uint64_t* tfull_bar = empty_bar + G2_HALO_B_MAX;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* afull_bar = tempty_bar + 2;                              // [2] halo mode: A tile ring
  uint64_t* aempty_bar = afull_bar + G2_HALO_A_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + G2_HALO_A_STAGES);
  uint64_t* ofull_bar = aempty_bar + G2_HALO_A_STAGES + 1;          // [8] TMA epilogue: residual box landed
  uint64_t* oempty_bar = ofull_bar + G2T_MAX_SLOTS;                  // [8] ... box free again (its output store has left shared memory)
  ClcShared* clcq = reinterpret_cast<ClcShared*>(after + (TEPI ? 512 + 2 * G2_BN * 4 + 128 * 4 : 256 + 4 * 2048));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G2_HALO_B_MAX; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < G2_HALO_A_STAGES; ++s) {
      mbar_init(&afull_bar[s], 1);
      mbar_init(&aempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);     // 4 epilogue warps of each CTA arrive on the leader's barrier
    }
    for (int s = 0; s < kClcStages; ++s) {     // consumers: 2 producers + MMA + 2 x 4 epilogue warps (+ 2 residual producers)
      mbar_init(&clcq->full[s], 1);
      mbar_init(&clcq->empty[s], 11 + ((TEPI && p.residual != nullptr) ? 2 : 0));
    }
    if (TEPI) {
      for (int s = 0; s < G2T_MAX_SLOTS; ++s) {
        mbar_init(&ofull_bar[s], 1);
        mbar_init(&oempty_bar[s], 1);
      }
      tma_prefetch_desc(&tmO);
      if (p.residual) tma_prefetch_desc(&tmR);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; from here on we read what it wrote
  pdl_trigger();
  TileCursor cur;
  cur.init(clcq, p.clc, npairs, 1, p.clc_unit, p.total_work);

  if (warp == 0) {
    // both CTAs: the whole warp runs the warp-uniform loop, one elected lane arms the barrier and issues the copies
    {
      int stage = 0;
      uint32_t phase = 0;
      if (p.halo) {
        uint8_t* bring = smem + G2_HALO_A_STAGES * p.a_tile_bytes;
        int sa = 0;
        uint32_t pa = 0;
        for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
          if (rank == 0) cur.request(2);
          const Work2 wi = decode2(p, w);
          const int m0 = wi.mt * 256 + (int)rank * 128;
          const int nb = wi.nt * G2_BN + (int)rank * 128;
          for (int kk = 0; kk < p.kb_per_tap; ++kk) {
            mbar_wait(&aempty_bar[sa], pa ^ 1, 25);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&afull_bar[sa], 2 * p.a_tile_bytes);
              uint8_t* sA = smem + sa * p.a_tile_bytes;
              for (int j = 0; j < p.a_nbox; ++j)
                tma_load_2d_2sm(sA + j * p.a_box_rows * 128, &tmA, &afull_bar[sa], kk * 64 + p.a_off0[0],
                                m0 + p.a_min_off + j * p.a_box_rows);
            }
            __syncwarp();
            if (++sa == G2_HALO_A_STAGES) {
              sa = 0;
              pa ^= 1;
            }
            for (int tap = 0; tap < p.ntaps; ++tap) {
              mbar_wait(&empty_bar[stage], phase ^ 1, 21);
              if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * 128 * 64 * 2);
                uint8_t* sB = bring + stage * (128 * 64 * 2);
                if (p.b_major == 0) {
                  tma_load_2d_2sm(sB, &tmB, &full_bar[stage], kk * 64 + p.b_off0[tap], nb + p.b_off1[tap]);
                } else {
#pragma unroll
                  for (int j = 0; j < 2; ++j)
                    tma_load_2d_2sm(sB + j * 8192, &tmB, &full_bar[stage], nb + j * 64 + p.b_off0[tap], kk * 64 + p.b_off1[tap]);
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
      } else
      for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
        if (rank == 0) cur.request(2);
        const Work2 wi = decode2(p, w);
        const int m0 = wi.mt * 256 + (int)rank * 128;
        const int nb = wi.nt * G2_BN + (int)rank * 128;     // this CTA's half of the B tile
        const int ntap = p.splits == 1 ? p.ntaps : 1;
        const int kb0 = p.splits == 1 ? 0 : wi.kb_begin;
        const int nkb = p.splits == 1 ? p.kb_per_tap : wi.iters;
        const int zb = p.z_b_off1[wi.z];
        for (int tap = 0; tap < ntap; ++tap) {
          const int a0 = p.a_off0[tap], a1 = p.a_off1[tap], b0 = p.b_off0[tap], b1 = p.b_off1[tap];
          for (int kk = kb0; kk < kb0 + nkb; ++kk) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 21);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
              uint8_t* sA = smem + stage * G2_STAGE_BYTES;
              uint8_t* sB = sA + 128 * 64 * 2;
              if (p.a_major == 0) {
                tma_load_2d_2sm(sA, &tmA, &full_bar[stage], kk * 64 + a0, m0 + a1);
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sA + j * 8192, &tmA, &full_bar[stage], m0 + j * 64 + a0, kk * 64 + a1);
              }
              if (p.b_major == 0) {
                tma_load_2d_2sm(sB, &tmB, &full_bar[stage], kk * 64 + b0, nb + b1);
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sB + j * 8192, &tmB, &full_bar[stage], nb + j * 64 + b0, kk * 64 + b1 + zb);
              }
            }
            __syncwarp();
            if (++stage == nst) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (TEPI && warp == 3) {
    // residual producer (both CTAs, own 128 rows): one 128 x 64 box per quarter tile, as far ahead as boxes are free
    if (p.residual != nullptr) {
      uint8_t* slots = smem + G2T_DATA_BYTES - p.nslots * G2T_BOX_BYTES;
      int os = 0;
      uint32_t oph = 0;
      for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
        const Work2 wi = decode2(p, w);
        const int m0 = wi.mt * 256 + (int)rank * 128;
        const int n0 = wi.nt * G2_BN;
        for (int q = 0; q < 4; ++q) {
          mbar_wait(&oempty_bar[os], oph ^ 1, 27);
          if (elect_one()) {
            mbar_expect_tx(&ofull_bar[os], G2T_BOX_BYTES);
            tma_load_2d(slots + os * G2T_BOX_BYTES, &tmR, &ofull_bar[os], n0 + q * 64, m0);
          }
          __syncwarp();
          if (++os == p.nslots) {
            os = 0;
            oph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // leader CTA only; the whole warp runs the warp-uniform loop, one elected lane issues tcgen05.mma / tcgen05.commit
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, G2_BN, p.a_major, p.b_major);
      const uint32_t a_k16 = (p.a_major ? 2048u : 32u) >> 4;
      const uint32_t b_k16 = (p.b_major ? 2048u : 32u) >> 4;
      const uint64_t a_hi = umma_smem_desc(0, p.a_major ? 8192u : 16u, 1024);
      const uint64_t b_hi = umma_smem_desc(0, p.b_major ? 8192u : 16u, 1024);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.halo) {
        const uint32_t bring = smem_base + G2_HALO_A_STAGES * p.a_tile_bytes;
        int sa = 0;
        uint32_t pa = 0;
        for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 22);
          const uint32_t d_tmem = tmem_base + acc * G2_BN;
          for (int kk = 0; kk < p.kb_per_tap; ++kk) {
            mbar_wait(&afull_bar[sa], pa, 26);
            const uint32_t a_tile = smem_base + sa * p.a_tile_bytes;
            for (int tap = 0; tap < p.ntaps; ++tap) {
              mbar_wait(&full_bar[stage], phase, 23);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_base = a_tile + (uint32_t)(p.a_off1[tap] - p.a_min_off) * 128u;
                const uint64_t boff = p.use_base_offset ? ((uint64_t)((a_base >> 7) & 7u) << 49) : 0ull;   // bring-up probe only
                const uint64_t ad = umma_desc_at(a_hi, a_base) | boff;
                const uint64_t bd = umma_desc_at(b_hi, bring + stage * (128 * 64 * 2));
#pragma unroll
                for (int s = 0; s < 4; ++s)
                  umma2_bf16(d_tmem, ad + s * 2u, bd + s * b_k16, idesc, (kk > 0 || tap > 0 || s > 0) ? 1u : 0u);
                umma2_commit_mc(&empty_bar[stage]);
              }
              __syncwarp();
              if (++stage == p.nb_stages) {
                stage = 0;
                phase ^= 1;
              }
            }
            if (elect_one()) umma2_commit_mc(&aempty_bar[sa]);
            __syncwarp();
            if (++sa == G2_HALO_A_STAGES) {
              sa = 0;
              pa ^= 1;
            }
          }
          if (elect_one()) umma2_commit_mc(&tfull_bar[acc]);
          __syncwarp();
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      } else
      for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
        const int iters = decode2(p, w).iters;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 22);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase, 23);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_base = smem_base + stage * G2_STAGE_BYTES;
            const uint64_t ad = umma_desc_at(a_hi, a_base);
            const uint64_t bd = umma_desc_at(b_hi, a_base + 128 * 64 * 2);
#pragma unroll
            for (int s = 0; s < 4; ++s) umma2_bf16(d_tmem, ad + s * a_k16, bd + s * b_k16, idesc, (it > 0 || s > 0) ? 1u : 0u);
            umma2_commit_mc(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_mc(&tfull_bar[acc]);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    if constexpr (TEPI) {
      // ---- TMA epilogue: thread r owns row r of this CTA's 128 x 256 half tile.  Per quarter (64 columns): wait for the residual
      // box (if any), fold scale / bias / residual / ReLU / mask into the bf16 output IN PLACE, hand the box to the TMA store.
      float* ssc = reinterpret_cast<float*>(after + 512);        // one copy of this tile's scale / bias for the 4 epilogue warps
      float* sbi = ssc + G2_BN;
      uint8_t* slots = smem + G2T_DATA_BYTES - p.nslots * G2T_BOX_BYTES;     // boxes sit at the end of the data region (any ring layout)
      int* srow = reinterpret_cast<int*>(sbi + G2_BN);                      // remapped outputs: destination row of each tile row, -1 = none
      const int Hp = p.img_h + 2, Wp = p.img_w + 2;
      const int rloc = wq * 32 + lane;
      const bool issuer = (warp == 4 && lane == 0);
      const int nslots = p.nslots;
      constexpr int NCH = G2_BN / 32, PFM = 2;
      int os = 0, qdone = 0;
      uint32_t oph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
        const Work2 wi = decode2(p, w);
        const int n0 = wi.nt * G2_BN;
        const int out_col0 = n0 + p.z_out_col[wi.z];
        const int m0c = wi.mt * 256 + (int)rank * 128;
        const int row_t = m0c + rloc;
        bool valid = row_t < p.M;
        int out_row = row_t;
        if (p.remap == TDB_REMAP_COMPACT_TO_PADDED) {
          const int hw = p.img_h * p.img_w;
          const int n = row_t / hw, rem = row_t - n * hw;
          const int h = rem / p.img_w, x = rem - h * p.img_w;
          out_row = (n * Hp + h + 1) * Wp + x + 1;
        } else if (p.remap == TDB_REMAP_PADDED_TO_COMPACT) {
          const int hw = Hp * Wp;
          const int n = row_t / hw, rem = row_t - n * hw;
          const int h = rem / Wp, x = rem - h * Wp;
          valid = valid && h >= 1 && h <= p.img_h && x >= 1 && x <= p.img_w;
          out_row = (n * p.img_h + (h - 1)) * p.img_w + (x - 1);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");      // the previous tile's readers are done with the scale / bias / row copies
        for (int i = (int)threadIdx.x - 128; i < G2_BN; i += 128) {
          ssc[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.f;
          sbi[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
        }
        srow[rloc] = valid ? out_row : -1;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint4 mbuf[PFM][4];
        const bf16* msk_row = p.mask ? p.mask + (long long)row_t * p.ldmask + n0 : nullptr;
        if (valid && msk_row) {
#pragma unroll
          for (int ci = 0; ci < PFM; ++ci)
#pragma unroll
            for (int i = 0; i < 4; ++i) mbuf[ci][i] = __ldg(reinterpret_cast<const uint4*>(msk_row + ci * 32) + i);
        }
        mbar_wait(&tfull_bar[acc], acc_phase, 24);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * G2_BN + ((uint32_t)(wq * 32) << 16);
        uint32_t r[2][32];
        tmem_ld_32x32(taddr, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = ci * 32;
          tmem_ld_wait();
          if (ci + 1 < NCH) {
            tmem_ld_32x32(taddr + c + 32, r[(ci + 1) & 1]);
          } else {
            // the whole accumulator is in registers: hand TMEM back to the MMA warp before the last box is finished
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(&tempty_bar[acc]);
              else mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
            }
          }
          if ((ci & 1) == 0 && p.residual != nullptr) mbar_wait(&ofull_bar[os], oph, 28);
          uint8_t* box = slots + os * G2T_BOX_BYTES + rloc * 128;
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
          if (p.residual != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j16 = (ci & 1) * 4 + i;
              const uint4 u = *reinterpret_cast<const uint4*>(box + ((j16 ^ (rloc & 7)) << 4));
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
              const uint4 u = mbuf[ci % PFM][i];
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
          if ((ci & 1) && p.remap != TDB_REMAP_NONE) {
            // remapped rows (zero-haloed pixel grids) cannot leave through one tensor-map box: the 128 threads copy the box out, 8
            // lanes per row, so that every store instruction writes four full 128-byte row segments instead of 32 half-used
            // sectors.  Two boxes alternate; the barrier of the next quarter orders this read before the box is written again.
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = (int)threadIdx.x - 128, j = t & 7;
            bf16* ocol = reinterpret_cast<bf16*>(p.out) + out_col0 + (ci >> 1) * 64 + j * 8;
            const uint8_t* bx = slots + os * G2T_BOX_BYTES;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int row = it * 16 + (t >> 3);
              const int orow = srow[row];
              if (orow >= 0)
                *reinterpret_cast<uint4*>(ocol + (long long)orow * p.ldo) = *reinterpret_cast<const uint4*>(bx + row * 128 + ((j ^ (row & 7)) << 4));
            }
            if (++os == nslots) {
              os = 0;
              oph ^= 1;
            }
          } else if (ci & 1) {
            // box complete.  Invariant: after this barrier every store but the one issued a quarter ago has finished READING its
            // box, so the box written next (nslots >= 3) is free, and the box of two quarters ago goes back to the residual producer.
            if (issuer) {
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              if (p.residual != nullptr && qdone >= 2) mbar_arrive(&oempty_bar[(os + nslots - 2) % nslots]);
            }
            fence_proxy_async();                                   // my shared-memory writes -> visible to the TMA engine
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (issuer) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmO),
                           "r"(smem_u32(slots + os * G2T_BOX_BYTES)), "r"(out_col0 + (ci >> 1) * 64), "r"(m0c)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++qdone;
            if (++os == nslots) {
              os = 0;
              oph ^= 1;
            }
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last boxes have left shared memory
    } else {
    float* ssc = reinterpret_cast<float*>(after + 256) + wq * 512;
    float* sbi = ssc + G2_BN;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int Hp = p.img_h + 2, Wp = p.img_w + 2;
    constexpr int NCH = G2_BN / 32, PFR = 4, PFM = 2;
    for (int w = cur.first(pair); w < p.total_work; w = cur.next(w)) {
      const Work2 wi = decode2(p, w);
      const int n0 = wi.nt * G2_BN;
      int out_col0 = n0 + p.z_out_col[wi.z];
      const int row_t = wi.mt * 256 + (int)rank * 128 + wq * 32 + lane;
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
      out_col0 += s2d_col;
      if (p.splits > 1) out_row += (long long)wi.split * p.M;
      for (int i = lane; i < G2_BN; i += 32) {
        ssc[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.f;
        sbi[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
      }
      __syncwarp();
      uint4 rbuf[PFR][4], mbuf[PFM][4];
      const bf16* res_row = p.residual ? p.residual + out_row * p.ldr + n0 : nullptr;
      const bf16* msk_row = p.mask ? p.mask + (long long)row_t * p.ldmask + n0 : nullptr;
      if (valid && res_row) {
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
      mbar_wait(&tfull_bar[acc], acc_phase, 24);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * G2_BN + ((uint32_t)(wq * 32) << 16);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // both CTAs are done with each other's shared memory, barriers and TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace tdb

int tdb_init_once();
int tdb_make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
int tdb_num_sms();
void tdb_count_launch(int n);

// returns 1 if the descriptor was handled by the 2-CTA kernel, 0 if it does not qualify, <0 on error
int tdb_gemm2_try(const tdb_gemm_desc* d, void* stream_) {
  using namespace tdb;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("TDB_GEMM2");
    enabled = e ? atoi(e) : 1;   // on by default (r01: +15..23 % on deep-K shapes); TDB_GEMM2=0 disables
  }
  const bool forced = (d->debug_flags >> 6) & 1;
  if (!enabled && !forced) return 0;
  static int wgrad_ok = -1;
  if (wgrad_ok < 0) { const char* e = getenv("TDB_GEMM2_WGRAD"); wgrad_ok = e ? atoi(e) : 1; }
  if (!forced && !wgrad_ok && d->a_major) return 0;
  int splits = d->splits < 1 ? 1 : d->splits;
  const int nz = d->nz < 1 ? 1 : d->nz;
  if (d->N % 256 != 0 || (d->block_n && d->block_n != 256)) return 0;
  if ((d->a_major == 0 || d->b_major == 0) && d->K % 64 != 0) return 0;
  const int kb = (d->K + 63) / 64;
  int kb_per_split = kb;
  if (splits > 1) {       // same split arithmetic as tdb_gemm / tdb_gemm_effective_splits
    if (d->ntaps != 1 || d->out_dtype != TDB_OUT_F32) return 0;
    if (splits > kb) splits = kb;
    kb_per_split = (kb + splits - 1) / splits;
    splits = (kb + kb_per_split - 1) / kb_per_split;
  }
  const int m_tiles = (d->M + 255) / 256;
  const int n_tiles = d->N / 256;
  const long long total = (long long)m_tiles * n_tiles * nz * splits;
  const int pairs = tdb_num_sms() / 2;
  // worthwhile only for deep reductions (the feed-bound regime) with enough tiles to fill the pairs, and when the
  // 256-row pair tile does not waste more than a quarter of its rows
  const long long kdepth = (splits > 1 ? (long long)kb_per_split * 64 : (long long)d->K * d->ntaps);
  static int min_k = -1;      // experiment switch: reduction depth from which the pair kernel is taken (default 512)
  if (min_k < 0) { const char* e = getenv("TDB_GEMM2_MIN_K"); min_k = e ? atoi(e) : 512; }
  // Shared-memory-box epilogue (tdb_gemm2_kernel<1>) for bf16 outputs: TMA store (+ TMA residual) when the rows are not remapped,
  // coalesced copy-out when they are.  Measured (tools/conv1x1_bench.py, profiles/r02_conv1x1_bench.txt): it beats the row-per-thread
  // epilogue on every shape; against the 1-CTA kernel with its own TMA epilogue the pair kernel wins from a reduction depth of 512.
  static int tepi_on = -1, tepi_min_k = -1, tepi_min_m = -1, copy_on = -1, copy_halo = -1;
  if (tepi_on < 0) { const char* e = getenv("TDB_GEMM2_TMA"); tepi_on = e ? atoi(e) : 1; }
  if (tepi_min_k < 0) { const char* e = getenv("TDB_GEMM2_TMA_MIN_K"); tepi_min_k = e ? atoi(e) : 512; }
  if (tepi_min_m < 0) { const char* e = getenv("TDB_GEMM2_TMA_MIN_M"); tepi_min_m = e ? atoi(e) : 2048; }
  if (copy_on < 0) { const char* e = getenv("TDB_GEMM2_COPYOUT"); copy_on = e ? atoi(e) : 0; }
  if (copy_halo < 0) { const char* e = getenv("TDB_GEMM2_COPYOUT_HALO"); copy_halo = e ? atoi(e) : 0; }
  const bool box_fmt = !((d->debug_flags >> 9) & 1) && d->out_dtype == TDB_OUT_BF16 && splits == 1 && nz == 1 && d->ldo % 8 == 0 &&
                       (forced || d->M >= tepi_min_m);
  const bool tma_ok = box_fmt && tepi_on && d->remap == TDB_REMAP_NONE && (!d->residual || d->ldr % 8 == 0);
  const bool copy_ok = box_fmt && copy_on && d->remap != TDB_REMAP_NONE && d->remap <= TDB_REMAP_PADDED_TO_COMPACT && !d->residual && d->img_h > 0 && d->img_w > 0 &&
                       (long long)d->M * 2 < (1ll << 31);
  bool tepi = tma_ok || copy_ok;
  if (!forced && (kdepth < (tma_ok ? tepi_min_k : min_k) || total < pairs / 2 || (long long)m_tiles * 256 * 3 > (long long)d->M * 4)) return 0;
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb_gemm2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb_gemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2T_SMEM));
    attr = true;
  }
  Gemm2Params p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.N = d->N; p.kb_per_tap = kb; p.ntaps = d->ntaps;
  p.a_major = d->a_major ? 1 : 0; p.b_major = d->b_major ? 1 : 0;
  p.nz = nz; p.splits = splits; p.kb_per_split = kb_per_split;
  for (int i = 0; i < TDB_MAX_TAPS; ++i) {
    p.a_off0[i] = d->a_off0[i]; p.a_off1[i] = d->a_off1[i];
    p.b_off0[i] = d->b_off0[i]; p.b_off1[i] = d->b_off1[i];
    p.z_b_off1[i] = d->nz >= 1 ? d->z_b_off1[i] : 0;
    p.z_out_col[i] = d->nz >= 1 ? d->z_out_col[i] : 0;
  }
  p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.total_work = (int)total;
  p.scale = d->scale; p.bias = d->bias;
  p.residual = (const bf16*)d->residual; p.ldr = d->ldr;
  p.mask = (const bf16*)d->mask; p.ldmask = d->ldmask;
  p.relu = d->relu; p.out = d->out; p.out_f32 = d->out_dtype == TDB_OUT_F32; p.ldo = d->ldo;
  p.remap = d->remap; p.img_h = d->img_h; p.img_w = d->img_w;
  // halo mode: K-major A whose taps differ only by a row shift (implicit 3x3 convolution over the zero-haloed pixel grid)
  static int halo_ok = -1;
  if (halo_ok < 0) { const char* e = getenv("TDB_GEMM2_HALO"); halo_ok = e ? atoi(e) : 1; }
  int a_box = 128;
  if (halo_ok && !((d->debug_flags >> 7) & 1) && p.a_major == 0 && d->ntaps > 1 && splits == 1 && nz == 1) {
    int lo = d->a_off1[0], hi = d->a_off1[0];
    bool same_cols = true;
    for (int i = 1; i < d->ntaps; ++i) {
      lo = d->a_off1[i] < lo ? d->a_off1[i] : lo;
      hi = d->a_off1[i] > hi ? d->a_off1[i] : hi;
      same_cols = same_cols && d->a_off0[i] == d->a_off0[0];
    }
    const int rows = 128 + hi - lo;
    const int nbox = (rows + 255) / 256;
    const int box_rows = (((rows + nbox - 1) / nbox) + 7) & ~7;
    const int tile_bytes = nbox * box_rows * 128;
    int nb = (G2_RING_BYTES - G2_HALO_A_STAGES * tile_bytes) / (128 * 64 * 2);
    if (tepi) {           // the boxes take the end of the 13 x 16 KB data region: 3 boxes, the weight ring gets what is left (at most 8)
      const int nbt = (G2T_DATA_BYTES - 3 * G2T_BOX_BYTES - G2_HALO_A_STAGES * tile_bytes) / (128 * 64 * 2);
      if (nbt >= 4 && !d->residual && (hi == lo || copy_halo)) nb = nbt;
      else tepi = false;
    }
    if (same_cols && nb >= 4) {
      p.halo = 1; p.a_min_off = lo; p.a_box_rows = box_rows; p.a_nbox = nbox; p.a_tile_bytes = tile_bytes;
      p.nb_stages = nb > G2_HALO_B_MAX ? G2_HALO_B_MAX : nb;
      p.use_base_offset = (d->debug_flags >> 8) & 1;
      a_box = box_rows;
    }
  }
  CUtensorMap tmA, tmB;
  int rc = tdb_make_tmap_bf16(&tmA, d->A, d->a_rows, d->a_cols, d->lda, p.a_major ? 64 : a_box);
  if (rc) return rc;
  rc = tdb_make_tmap_bf16(&tmB, d->B, d->b_rows, d->b_cols, d->ldb, p.b_major ? 64 : 128);
  if (rc) return rc;
  int np = total < pairs ? (int)total : pairs;
  if (d->max_ctas > 1 && np * 2 > d->max_ctas) np = d->max_ctas / 2;
  // cluster launch control: one cluster per unit of tiles, resident clusters steal the units of clusters that have not started
  // (tdb_common.cuh).  TDB_CLC: 0 never, 1 (default) when the caller flags the launch as running next to other streams' kernels
  // (TDB_GEMM_FLAG_DYNAMIC_TILES: measured +5 % without contention, -25..45 % with 8-32 SMs held by another kernel), 2 always.
  static int clc_on = -1;
  if (clc_on < 0) { const char* e = getenv("TDB_CLC"); clc_on = e ? atoi(e) : 1; }
  const bool want_clc = clc_on == 2 || (clc_on == 1 && ((d->debug_flags >> 11) & 1));
  if (want_clc && !((d->debug_flags >> 10) & 1) && total > pairs) {
    // a unit should outlast the request's round trip (~2 us): pair tile = ~0.27 us per k-block + ~1 us
    const double tile_us = 0.27 * (double)(splits > 1 ? kb_per_split : kb * d->ntaps) + 1.0;
    int unit = (int)(6.0 / tile_us + 0.999);
    const int cap = (int)(total / (4 * pairs));           // keep at least ~4 units per resident pair
    if (unit > cap) unit = cap;
    if (unit < 1) unit = 1;
    p.clc = 1;
    p.clc_unit = unit;
    np = (int)((total + unit - 1) / unit);
  }
  if (tepi) {
    CUtensorMap tmR = tmA, tmO = tmA;
    if (d->remap == TDB_REMAP_NONE) {
      rc = tdb_make_tmap_bf16(&tmO, d->out, d->M, d->N + (d->nz >= 1 ? d->z_out_col[0] : 0), d->ldo, 128);
      if (rc) return rc;
    }
    if (d->residual) {
      rc = tdb_make_tmap_bf16(&tmR, d->residual, d->M, d->N, d->ldr, 128);
      if (rc) return rc;
    }
    static int st_res = -1, st_plain = -1;     // operand stages (32 KB) of the 13 x 16 KB; the rest are output / residual boxes
    if (st_res < 0) { const char* e = getenv("TDB_GEMM2_TMA_STAGES_RES"); st_res = e ? atoi(e) : 3; }
    if (st_plain < 0) { const char* e = getenv("TDB_GEMM2_TMA_STAGES"); st_plain = e ? atoi(e) : 5; }
    int nstg = d->residual ? st_res : st_plain;
    if (nstg < 2) nstg = 2;
    if (nstg > 5) nstg = 5;
    p.nstages = nstg;
    p.nslots = 13 - 2 * nstg;
    if (p.nslots > G2T_MAX_SLOTS) p.nslots = G2T_MAX_SLOTS;
    if (p.halo) p.nslots = 3;
    TDB_CHECK_CUDA(tdb_launch(tdb_gemm2_kernel<1>, dim3(np * 2), dim3(G2_THREADS), G2T_SMEM, (cudaStream_t)stream_, tmA, tmB, tmR, tmO, p));
    TDB_CHECK_CUDA(cudaGetLastError());
    tdb_count_launch(1);
    return 1;
  }
  TDB_CHECK_CUDA(tdb_launch(tdb_gemm2_kernel<0>, dim3(np * 2), dim3(G2_THREADS), G2_SMEM, (cudaStream_t)stream_, tmA, tmB, tmA, tmA, p));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return 1;
}
