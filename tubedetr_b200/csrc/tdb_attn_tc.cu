// tubedetr_b200 -- self-attention core on tcgen05 (encoder spatial attention, decoder temporal self-attention; head_dim 32).
//
// Reference: models/transformer.py:637-640 and 698-722 (nn.MultiheadAttention, need_weights path): S = (q hd^-1/2) k^T + mask,
// P = softmax(S), O = dropout(P) V.  Both contractions of the forward and all four of the backward run on the 5th-generation
// tensor cores, accumulators in TMEM, operands by TMA:
//   forward    S[128 x LKP] = Q_h[128 x 32] K_h[LKP x 32]^T      TMEM columns [0, LKP)                 (2 tcgen05.mma, K = 16 each)
//              O[128 x 64]  = P[128 x LKP] V_pair[LKP x 64]      TMEM columns [o_col, o_col + 64)      (LKP / 16 tcgen05.mma)
//   backward   dPd = dO V^T, dQ = dS K, dK = dS^T Q, dV = Pd^T dO (see mha_tc_bwd_kernel)
// One CTA = (sequence b, head h).  A head is only 32 channels = 64 bytes wide, half a 128-byte swizzle row, so Q, K and V are
// fetched by TMA as 64-column boxes covering the head PAIR g = h / 2; head e = h % 2 of the pair is addressed by starting the
// operand descriptors e * 64 bytes into the swizzled rows (the same mechanism as the +32-byte K steps of the GEMM main loop).
// V is used as an MN-major B operand with N = 64 (both heads' channels); only the 32 output columns of head e are kept.
// Softmax runs on 128 threads (thread = query row = TMEM lane): three light passes over the S accumulator (max, sum, normalise).
// Probabilities go to global memory once (fp32: the backward and the guided-attention loss read them) THROUGH a per-warp
// shared-memory transpose, so every global store / load instruction of a warp covers 128 contiguous bytes of one row (a
// thread-per-row access pattern costs 8x the L2 sector operations), and as bf16 in the 128-byte-swizzled K-major layout to shared
// memory as the A operand of the second contraction.  Attention dropout is generated in the kernel from the library's
// counter-based hash stream (tdb_common.cuh: same bits as tdb_dropout_mask(seed, site) at the flat [B][H][Lq][Lk] index): no mask
// tensor exists, the backward regenerates the bits.
// Warp roles: warp 0 = TMA + MMA issue (one elected lane), warp 2 = TMEM allocator, warps 4-7 = softmax / epilogue.
// Footprint (forward, S = 141): 105 KB of shared memory and 256 TMEM columns -> two CTAs per SM, so one CTA's softmax overlaps the
// other's MMAs and loads.
#include <stdlib.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);
int tdb_init_once();
int tdb_make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

namespace tdb {

constexpr int TC_THREADS = 256;
constexpr int TC_MAXL = 256;                 // queries and keys per sequence

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// keep bits of the 32 consecutive elements [e0, e0 + 32) of the dropout stream `base` (bit t = element e0 + t kept).
// Element e uses 16-bit lane (e & 3) of hash number (e >> 2): the 32 elements touch 8 or 9 consecutive hashes depending on e0 & 3,
// which differs from thread to thread (rows of 141 floats) -- so all 9 are evaluated unconditionally and the 36 lane bits are
// shifted into place: no divergent branch inside the warp (a per-element "new hash?" test serialises into 32 masked hash evaluations).
__device__ __forceinline__ uint32_t keep_bits32(unsigned long long base, long long e0, uint32_t thr) {
  const unsigned long long h0 = base + (unsigned long long)(e0 >> 2);
  unsigned long long m = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const unsigned long long r = splitmix64(h0 + (unsigned long long)k);
    const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
    const uint32_t b4 = (uint32_t)((lo & 0xFFFFu) >= thr) | ((uint32_t)((lo >> 16) >= thr) << 1) | ((uint32_t)((hi & 0xFFFFu) >= thr) << 2) |
                        ((uint32_t)((hi >> 16) >= thr) << 3);
    m |= (unsigned long long)b4 << (4 * k);
  }
  return (uint32_t)(m >> (int)(e0 & 3));
}

// ---- per-warp 32 x 32 fp32 transpose through shared memory (4 KB, XOR-swizzled: conflict-free both ways) -----------------------------
// thread `lane` owns row (row0 + lane) of a [nrows x ld] fp32 matrix and wants / has columns [col0, col0 + 32) of it in registers;
// the global side is accessed one ROW per warp instruction (lane = column): 128 contiguous bytes.
__device__ __forceinline__ void warp_rows_store(float* stage, int lane, float* g, long long ld, int row0, int nrows, int col0, int ncols,
                                                const float (&x)[32]) {
#pragma unroll
  for (int t = 0; t < 32; ++t) stage[lane * 32 + (t ^ lane)] = x[t];
  __syncwarp();
  const int cols = ncols - col0;                       // valid columns of this chunk
  const int rows = nrows - row0;
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr) {
    const float v = stage[rr * 32 + (lane ^ rr)];
    if (rr < rows && lane < cols) g[(long long)(row0 + rr) * ld + col0 + lane] = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_rows_load(float* stage, int lane, const float* g, long long ld, int row0, int nrows, int col0,
                                               int ncols, float (&x)[32]) {
  const int cols = ncols - col0;
  const int rows = nrows - row0;
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr) {
    const float v = (rr < rows && lane < cols) ? __ldg(g + (long long)(row0 + rr) * ld + col0 + lane) : 0.f;
    stage[rr * 32 + (lane ^ rr)] = v;
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 32; ++t) x[t] = stage[lane * 32 + (t ^ lane)];
  __syncwarp();
}

// the same load in two halves, so that the global loads of the NEXT chunk are in flight while the current one is processed:
// issue = 32 coalesced row-segment loads into registers (lane = column), commit = transpose through the staging buffer
__device__ __forceinline__ void warp_rows_issue(int lane, const float* g, long long ld, int row0, int nrows, int col0, int ncols,
                                                float (&raw)[32]) {
  const int cols = ncols - col0;
  const int rows = nrows - row0;
  const float* base = g + (long long)row0 * ld + col0 + lane;
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) raw[rr] = (rr < rows && lane < cols) ? __ldg(base + (long long)rr * ld) : 0.f;
}
__device__ __forceinline__ void warp_rows_commit(float* stage, int lane, const float (&raw)[32], float (&x)[32]) {
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) stage[rr * 32 + (lane ^ rr)] = raw[rr];
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 32; ++t) x[t] = stage[lane * 32 + (t ^ lane)];
  __syncwarp();
}

__device__ __forceinline__ void store_swizzled_row32(uint8_t* tile, int r, int c, const float (&x)[32]) {
  // columns [32c, 32c + 32) of row r of a K-major 128-byte-swizzled tile made of 64-column blocks [128 x 128 B]
  uint8_t* blk = tile + (c >> 1) * 16384 + r * 128;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const int chunk = ((c & 1) * 4 + q4) ^ (r & 7);
    uint4 w;
    w.x = pack_bf16x2(x[q4 * 8 + 0], x[q4 * 8 + 1]);
    w.y = pack_bf16x2(x[q4 * 8 + 2], x[q4 * 8 + 3]);
    w.z = pack_bf16x2(x[q4 * 8 + 4], x[q4 * 8 + 5]);
    w.w = pack_bf16x2(x[q4 * 8 + 6], x[q4 * 8 + 7]);
    *reinterpret_cast<uint4*>(blk + chunk * 16) = w;
  }
}
__device__ __forceinline__ void store_row32_bf16(bf16* dst, const uint32_t (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    d4[q4] = make_uint4(pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]), __uint_as_float(v[q4 * 8 + 1])),
                        pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]), __uint_as_float(v[q4 * 8 + 3])),
                        pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]), __uint_as_float(v[q4 * 8 + 5])),
                        pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]), __uint_as_float(v[q4 * 8 + 7])));
}

struct TcSmem {
  uint64_t kv_full, q_full, s_full, p_full, o_full, o_free;
  uint32_t tmem_slot, pad;
  uint32_t dead[TC_MAXL / 4];   // one byte per key: 1 = padded key or beyond Lk (staged once per CTA)
};

struct TcParams {
  const uint8_t* kpm;   // [B][Lk] nonzero = masked key (may be null)
  bf16* o;              // [B*Lq][ldo]
  long long ldo;
  float* p;             // [B][H][Lq][Lk] normalised probabilities before dropout
  float* pdrop;         // same shape, after dropout (null unless requested)
  const long long* drop_seed;   // device seed of the dropout stream (null = no dropout)
  unsigned long long drop_site;
  uint32_t drop_thr;
  float keep_scale;
  int B, H, Lq, Lk, LKP, MT, o_col, tmem_cols;
  float scale_log2;     // softmax scale * log2(e)
};

__global__ void __launch_bounds__(TC_THREADS, 2)
mha_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ TcParams a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: Q tile [128 x 128 B] (doubles as the 4 x 4 KB transpose staging of the softmax warps once S = Q K^T has completed)
  //         | K [LKP x 128 B] | V [LKP x 128 B] | P blocks [LKP/64 rounded up][128 x 128 B] | barriers
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 16384;
  uint8_t* sV = sK + a.LKP * 128;
  uint8_t* sP = sV + a.LKP * 128;            // LKP is a multiple of 32, so every region stays 1024-byte aligned (LKP*128 % 4096 == 0)
  const int pblocks = (a.LKP + 63) >> 6;
  TcSmem& sh = *reinterpret_cast<TcSmem*>(sP + pblocks * 16384);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // one CTA per (sequence b, head h, 128-row query tile mt)
  const int mt = blockIdx.x % a.MT;
  const int h = (blockIdx.x / a.MT) % a.H, b = blockIdx.x / (a.MT * a.H);
  const int g = h >> 1, e = h & 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&sh.kv_full, 1);
    mbar_init(&sh.q_full, 1);
    mbar_init(&sh.s_full, 1);
    mbar_init(&sh.p_full, 128);
    mbar_init(&sh.o_full, 1);
    mbar_init(&sh.o_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&sh.tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int iters = 1;                        // (kept as a loop: the barrier protocol supports several tiles per CTA)
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&sh.kv_full, (uint32_t)(2 * a.LKP * 128));
      tma_load_2d(sK, &tmK, &sh.kv_full, g * 64, b * a.Lk);
      tma_load_2d(sV, &tmV, &sh.kv_full, g * 64, b * a.Lk);
      const uint32_t idesc_s = umma_idesc_bf16(128, a.LKP, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);          // B = V, MN-major (N contiguous)
      const uint64_t k_hi = umma_smem_desc(0, 16, 1024);                // K-major, 128-byte swizzle
      const uint64_t mn_hi = umma_smem_desc(0, 8192, 1024);             // MN-major, 128-byte swizzle
      for (int it = 0; it < iters; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        // the Q buffer is free: for it > 0 the softmax warps signalled p_full(it - 1) after their last use of it as staging
        mbar_expect_tx(&sh.q_full, 16384);
        tma_load_2d(sQ, &tmQ, &sh.q_full, g * 64, b * a.Lq + (mt + it) * 128);
        if (it == 0) mbar_wait(&sh.kv_full, 0, 41);
        mbar_wait(&sh.q_full, ph, 40);
        tc_fence_after();
        // ---- S = Q_e K_e^T  (the S accumulator is free: the softmax threads signalled p_full of the previous iteration)
        const uint32_t qa = smem_u32(sQ) + e * 64;
        const uint32_t ka = smem_u32(sK) + e * 64;
        umma_bf16(tmem_base, umma_desc_at(k_hi, qa), umma_desc_at(k_hi, ka), idesc_s, 0u);
        umma_bf16(tmem_base, umma_desc_at(k_hi, qa + 32), umma_desc_at(k_hi, ka + 32), idesc_s, 1u);
        umma_commit(&sh.s_full);
        // ---- O = P V_pair once the probabilities are in shared memory; the previous O must have been read
        mbar_wait(&sh.p_full, ph, 42);
        if (it > 0) mbar_wait(&sh.o_free, ph ^ 1, 43);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
        for (int ks = 0; ks < a.LKP / 16; ++ks) {
          const uint32_t a_addr = pa + (ks >> 2) * 16384 + (ks & 3) * 32;   // 64-column blocks of [128 x 128 B], +32 B per k16
          const uint32_t b_addr = va + ks * 2048;                            // 16 key rows of 128 B
          umma_bf16(tmem_base + a.o_col, umma_desc_at(k_hi, a_addr), umma_desc_at(mn_hi, b_addr), idesc_o, ks > 0 ? 1u : 0u);
        }
        umma_commit(&sh.o_full);
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    const int r = wq * 32 + lane;                                   // row of the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint8_t* mk = a.kpm ? a.kpm + (long long)b * a.Lk : nullptr;
    const int nch = a.LKP >> 5;
    float* stage = reinterpret_cast<float*>(sQ + wq * 4096);
    {
      uint8_t* db = reinterpret_cast<uint8_t*>(sh.dead);
      for (int j = threadIdx.x - 128; j < a.LKP; j += 128) db[j] = (j >= a.Lk || (mk && mk[j])) ? 1 : 0;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const unsigned long long dbase = a.drop_seed ? drop_base(a.drop_seed, a.drop_site) : 0ull;
    const long long bh = (long long)b * a.H + h;
    float* pmat = a.p + bh * a.Lq * a.Lk;
    float* pdmat = a.pdrop ? a.pdrop + bh * a.Lq * a.Lk : nullptr;
    for (int it = 0; it < iters; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int row0 = (mt + it) * 128 + wq * 32;     // first row of this warp
      const int i = row0 + lane;
      const bool valid = i < a.Lq;
      mbar_wait(&sh.s_full, ph, 44);
      tc_fence_after();
      if (row0 >= a.Lq) {        // no row of this warp is inside the sequence: its P rows feed output rows that are never stored
        if (it > 0) mbar_wait(&sh.o_full, ph ^ 1, 45);
        tc_fence_before();
        mbar_arrive(&sh.p_full);
        mbar_wait(&sh.o_full, ph, 46);
        tc_fence_before();
        mbar_arrive(&sh.o_free);
        continue;
      }
      // pass 1: row maximum of the masked, scaled scores (log2 domain)
      float mx = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const bool dead = ((sh.dead[c * 8 + (t >> 2)] >> (8 * (t & 3))) & 0xffu) != 0;
          mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(v[t]) * a.scale_log2);
        }
      }
      // pass 2: sum of exponentials
      float sum = 0.f;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const bool dead = ((sh.dead[c * 8 + (t >> 2)] >> (8 * (t & 3))) & 0xffu) != 0;
          sum += (dead || mx == -INFINITY) ? 0.f : ex2(__uint_as_float(v[t]) * a.scale_log2 - mx);
        }
      }
      const float inv = 1.f / sum;            // all keys masked -> NaN, exactly like the reference softmax
      // the previous iteration's O = P V must have finished reading the P tile before it is overwritten
      if (it > 0) mbar_wait(&sh.o_full, ph ^ 1, 45);
      // pass 3: normalised probabilities -> global (fp32, coalesced through the staging transpose) and, after dropout, bf16 into
      // the swizzled K-major A tile
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + c * 32, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int j = c * 32 + t;
          const bool dead = ((sh.dead[c * 8 + (t >> 2)] >> (8 * (t & 3))) & 0xffu) != 0;
          const float pn = (dead || mx == -INFINITY) ? 0.f : ex2(__uint_as_float(v[t]) * a.scale_log2 - mx);
          pv[t] = (valid && j < a.Lk) ? pn * inv : 0.f;
        }
        if (c * 32 < a.Lk) warp_rows_store(stage, lane, pmat, a.Lk, row0, a.Lq, c * 32, a.Lk, pv);
        if (a.drop_seed) {
          const uint32_t bits = keep_bits32(dbase, (bh * a.Lq + i) * a.Lk + c * 32, a.drop_thr);
#pragma unroll
          for (int t = 0; t < 32; ++t) pv[t] = ((bits >> t) & 1u) ? pv[t] * a.keep_scale : 0.f;
          if (pdmat && c * 32 < a.Lk) warp_rows_store(stage, lane, pdmat, a.Lk, row0, a.Lq, c * 32, a.Lk, pv);
        }
        store_swizzled_row32(sP, r, c, pv);
      }
      fence_proxy_async();                    // generic-proxy writes (P tile, staging over the Q tile) -> visible to the async proxy
      tc_fence_before();                      // our tcgen05.ld of S are done before the next S MMA may overwrite it
      mbar_arrive(&sh.p_full);
      // ---- O epilogue: the 32 channels of head e
      mbar_wait(&sh.o_full, ph, 46);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + a.o_col + e * 32, v);
        tmem_ld_wait();
        if (valid) store_row32_bf16(a.o + ((long long)b * a.Lq + i) * a.ldo + h * 32, v);
      }
      tc_fence_before();
      mbar_arrive(&sh.o_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward on the same scheme.  Per (sequence b, head h = 2 g + e) and 128-row query tile mt:
//   MMA1   dPd[128 x LKP]  = dO_e V_e^T                                  -> TMEM [0, LKP)
//   rows   dP = dropout'(dPd + dPbar / H),  rs = sum_j P dP,  dS = scale P (dP - rs)        (thread = query row)
//          dS and Pd (post-dropout P) as bf16 into two 128-byte-swizzled K-major tiles [128 x LKP] in shared memory;
//          P (and dPbar) rows come in through the per-warp staging transpose (coalesced 128-byte row segments)
//   MMA2a  dQ[128 x 64]    = dS K_pair        (A = dS K-major,  B = K MN-major)              -> TMEM [0, 64)  (dPd is consumed)
//   MMA2b  dK[keys x 64]  += dS^T Q_pair      (A = the SAME dS tile read MN-major, B = Q MN-major) -> TMEM [256 + 64 kt, +64)
//   MMA2c  dV[keys x 64]  += Pd^T dO_pair     (A = Pd tile read MN-major, B = dO MN-major)   -> TMEM [384 + 64 kt, +64)
// dK / dV accumulate over the query tiles and are written once; every result keeps the 32 columns of head e of the 64-wide
// pair product.  The softmax scale is folded into dS, so dQ and dK come out scaled.
struct TcBwdSmem {
  uint64_t kvfull, qfull, s_full, p_full, mma2_done, dq_free;
  uint32_t tmem_slot, pad;
};

struct TcBwdParams {
  const float* p;       // [B][H][Lq][Lk] normalised probabilities before dropout
  const long long* drop_seed;
  unsigned long long drop_site;
  uint32_t drop_thr;
  float keep_scale;
  const float* dpbar;   // [B][Lq][Lk] or null
  bf16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  int B, H, Lq, Lk, LKP, MT, KT, nblk, v_reload;
  float scale;
  long long* stamps;    // measurement hook (null = off): CTA 0 stores SM clock stamps at its phase boundaries, [16] int64
};

__global__ void __launch_bounds__(TC_THREADS, 1)
mha_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                  const __grid_constant__ TcBwdParams a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: Q tile [128 x 128 B] | dO tile | K [LKP x 128 B] | V [LKP x 128 B] | dS blocks [nblk][128 x 128 B] | Pd blocks [nblk]
  //         | tail 16 KB (transpose staging 4 x 4 KB; also what the MN-major reads of a missing 4th block land in) | barriers
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + 16384;
  uint8_t* sK = sdO + 16384;
  uint8_t* sV = sK + a.LKP * 128;
  uint8_t* sDS = sV + a.LKP * 128;
  uint8_t* sPD = sDS + a.nblk * 16384;
  uint8_t* sTail = sPD + a.nblk * 16384;
  // more than 192 keys (4 blocks per tile): no room for the tail -- the staging buffers alias the V tile, which is then re-fetched
  // for every query tile (V is only read by MMA1, before the row passes start)
  TcBwdSmem& sh = *reinterpret_cast<TcBwdSmem*>(sTail + (a.v_reload ? 0 : 16384));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;      // one CTA per (sequence, head); g = head pair, e = head of the pair
  const int g = h >> 1, e = h & 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    mbar_init(&sh.kvfull, 1);
    mbar_init(&sh.qfull, 1);
    mbar_init(&sh.s_full, 1);
    mbar_init(&sh.p_full, 128);
    mbar_init(&sh.mma2_done, 1);
    mbar_init(&sh.dq_free, 128);
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

  constexpr uint32_t DK_COL = 256, DV_COL = 384;
  const int iters = a.MT;
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&sh.kvfull, (uint32_t)((a.v_reload ? 1 : 2) * a.LKP * 128));
      tma_load_2d(sK, &tmK, &sh.kvfull, g * 64, b * a.Lk);
      if (!a.v_reload) tma_load_2d(sV, &tmV, &sh.kvfull, g * 64, b * a.Lk);
      const uint32_t idesc_p = umma_idesc_bf16(128, a.LKP, 0, 0);      // dPd: A = dO (K-major), B = V (K-major)
      const uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);         // dQ:  A = dS (K-major), B = K (MN-major)
      const uint32_t idesc_kv = umma_idesc_bf16(128, 64, 1, 1);        // dK / dV: A = dS / Pd read MN-major, B = Q / dO (MN-major)
      const uint64_t k_hi = umma_smem_desc(0, 16, 1024);
      const uint64_t mn_hi = umma_smem_desc(0, 8192, 1024);            // one 64-wide MN chunk
      const uint64_t amn_hi = umma_smem_desc(0, 16384, 1024);          // A read MN-major: 64-column blocks are 16 KB apart
      for (int it = 0; it < iters; ++it) {
        const int mt = it;
        const uint32_t ph = (uint32_t)(it & 1);
        if (it > 0) mbar_wait(&sh.dq_free, ph ^ 1, 51);    // previous tile fully consumed: Q / dO buffers, dS / Pd tiles, TMEM [0, 256)
        mbar_expect_tx(&sh.qfull, (uint32_t)(2 * 16384 + (a.v_reload ? a.LKP * 128 : 0)));
        tma_load_2d(sQ, &tmQ, &sh.qfull, g * 64, b * a.Lq + mt * 128);
        tma_load_2d(sdO, &tmdO, &sh.qfull, g * 64, b * a.Lq + mt * 128);
        if (a.v_reload) tma_load_2d(sV, &tmV, &sh.qfull, g * 64, b * a.Lk);
        if (it == 0) mbar_wait(&sh.kvfull, 0, 52);
        mbar_wait(&sh.qfull, ph, 53);
        tc_fence_after();
        // ---- MMA1: dPd = dO_e V_e^T
        const uint32_t da = smem_u32(sdO) + e * 64, va = smem_u32(sV) + e * 64;
        umma_bf16(tmem_base, umma_desc_at(k_hi, da), umma_desc_at(k_hi, va), idesc_p, 0u);
        umma_bf16(tmem_base, umma_desc_at(k_hi, da + 32), umma_desc_at(k_hi, va + 32), idesc_p, 1u);
        umma_commit(&sh.s_full);
        // ---- MMA2 once dS / Pd are in shared memory (and dPd has been read)
        mbar_wait(&sh.p_full, ph, 54);
        tc_fence_after();
        const uint32_t dsa = smem_u32(sDS), pda = smem_u32(sPD), ka = smem_u32(sK), qa = smem_u32(sQ), doa = smem_u32(sdO);
        for (int ks = 0; ks < a.LKP / 16; ++ks)             // dQ: reduction over keys
          umma_bf16(tmem_base, umma_desc_at(k_hi, dsa + (ks >> 2) * 16384 + (ks & 3) * 32), umma_desc_at(mn_hi, ka + ks * 2048), idesc_q,
                    ks > 0 ? 1u : 0u);
        for (int kt = 0; kt < a.KT; ++kt) {                 // dK, dV: reduction over the 128 query rows of this tile
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t acc = (mt > 0 || ks > 0) ? 1u : 0u;
            umma_bf16(tmem_base + DK_COL + kt * 64, umma_desc_at(amn_hi, dsa + (2 * kt) * 16384 + ks * 2048),
                      umma_desc_at(mn_hi, qa + ks * 2048), idesc_kv, acc);
            umma_bf16(tmem_base + DV_COL + kt * 64, umma_desc_at(amn_hi, pda + (2 * kt) * 16384 + ks * 2048),
                      umma_desc_at(mn_hi, doa + ks * 2048), idesc_kv, acc);
          }
        }
        umma_commit(&sh.mma2_done);
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const int nch = a.LKP >> 5;
    const float invH = 1.f / a.H;
    float* stage = reinterpret_cast<float*>((a.v_reload ? sV : sTail) + wq * 4096);
    const unsigned long long dbase = a.drop_seed ? drop_base(a.drop_seed, a.drop_site) : 0ull;
    const long long bh = (long long)b * a.H + h;
    const float* pmat = a.p + bh * a.Lq * a.Lk;
    const float* dpmat = a.dpbar ? a.dpbar + (long long)b * a.Lq * a.Lk : nullptr;
    // the probabilities (and dPbar) do not depend on this kernel's MMAs: the first chunk of every query tile is requested before the
    // wait for dPd, each further chunk while its predecessor is processed
    float praw[32], draw[32];
    if (wq * 32 < a.Lq) {
      warp_rows_issue(lane, pmat, a.Lk, wq * 32, a.Lq, 0, a.Lk, praw);
      if (dpmat) warp_rows_issue(lane, dpmat, a.Lk, wq * 32, a.Lq, 0, a.Lk, draw);
    }
    for (int it = 0; it < iters; ++it) {
      const int mt = it;
      const uint32_t ph = (uint32_t)(it & 1);
      const int i = mt * 128 + r;
      const int row0 = mt * 128 + wq * 32;
      const bool valid = i < a.Lq;
      const long long prow = (bh * a.Lq + i) * a.Lk;
      const bool stamp = a.stamps && blockIdx.x == 0 && r == 0;
      if (stamp && it == 0) a.stamps[0] = clock64();
      mbar_wait(&sh.s_full, ph, 56);
      tc_fence_after();
      if (stamp) a.stamps[1 + 6 * it] = clock64();          // dPd ready (loads + MMA1)
      if (row0 >= a.Lq) {
        // no row of this warp is inside the sequence: its dS / Pd rows must be ZERO (they are summed over by the dK / dV MMAs)
        float z[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) z[t] = 0.f;
        for (int c = 0; c < nch; ++c) {
          store_swizzled_row32(sPD, r, c, z);
          store_swizzled_row32(sDS, r, c, z);
        }
      } else {
        // pass A: d = dropout'(dPd + dPbar / H) written back to TMEM, rs = sum_j P d; the post-dropout probabilities (A operand
        // of dV) go to the Pd tile and the bf16 probabilities are parked in the dS tile for pass B (no second trip to global
        // memory, no second hash evaluation)
        float rs = 0.f;
        for (int c = 0; c < nch; ++c) {
          float pj[32], db[32];
          warp_rows_commit(stage, lane, praw, pj);
          if (dpmat) warp_rows_commit(stage, lane, draw, db);
          {   // next chunk of this tile, or the first chunk of the next tile
            const bool more = c + 1 < nch;
            const int nrow0 = more ? row0 : row0 + 128, ncol0 = more ? (c + 1) * 32 : 0;
            if ((more || it + 1 < iters) && nrow0 < a.Lq && ncol0 < a.Lk) {
              warp_rows_issue(lane, pmat, a.Lk, nrow0, a.Lq, ncol0, a.Lk, praw);
              if (dpmat) warp_rows_issue(lane, dpmat, a.Lk, nrow0, a.Lq, ncol0, a.Lk, draw);
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t) praw[t] = draw[t] = 0.f;
            }
          }
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + c * 32, v);
          tmem_ld_wait();
          if (dpmat) {
#pragma unroll
            for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) + db[t] * invH);
          }
          const uint32_t bits = a.drop_seed ? keep_bits32(dbase, prow + c * 32, a.drop_thr) : 0xffffffffu;
          float pd[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const bool kept = (bits >> t) & 1u;
            const float d = kept ? __uint_as_float(v[t]) * a.keep_scale : 0.f;     // keep_scale = 1 without dropout
            v[t] = __float_as_uint(d);
            rs += pj[t] * d;                                                          // pj = 0 outside the sequence
            pd[t] = kept ? pj[t] * a.keep_scale : 0.f;
          }
          tmem_st_32x32(lane_addr + c * 32, v);
          store_swizzled_row32(sPD, r, c, pd);
          store_swizzled_row32(sDS, r, c, pj);
        }
        tmem_st_wait();
        if (stamp) a.stamps[2 + 6 * it] = clock64();        // pass A done
        // pass B: dS = scale P (d - rs) as bf16 over the parked probabilities (zero rows / columns outside the sequence: P = 0)
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + c * 32, v);
          tmem_ld_wait();
          uint8_t* blk = sDS + (c >> 1) * 16384 + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4* cell = reinterpret_cast<uint4*>(blk + ((((c & 1) * 4 + q4) ^ (r & 7)) * 16));
            const uint4 pw = *cell;
            const uint32_t pin[4] = {pw.x, pw.y, pw.z, pw.w};
            uint32_t o4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 p2 = unpack_bf16x2(pin[u]);
              const float d0 = __uint_as_float(v[q4 * 8 + 2 * u]), d1 = __uint_as_float(v[q4 * 8 + 2 * u + 1]);
              o4[u] = pack_bf16x2(a.scale * p2.x * (d0 - rs), a.scale * p2.y * (d1 - rs));
            }
            *cell = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
      }
      if (stamp) a.stamps[3 + 6 * it] = clock64();          // pass B done
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&sh.p_full);
      // ---- dQ epilogue
      mbar_wait(&sh.mma2_done, ph, 57);
      tc_fence_after();
      if (stamp) a.stamps[4 + 6 * it] = clock64();          // MMA2 (dQ, dK, dV) done
      {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + e * 32, v);
        tmem_ld_wait();
        if (valid) store_row32_bf16(a.dq + ((long long)b * a.Lq + i) * a.lddq + h * 32, v);
      }
      // ---- dK / dV epilogue after the last query tile (thread = key row of each 128-key tile)
      if (mt == a.MT - 1) {
        for (int kt = 0; kt < a.KT; ++kt) {
          const int j = kt * 128 + r;
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + DK_COL + kt * 64 + e * 32, v);
          tmem_ld_wait();
          if (j < a.Lk) store_row32_bf16(a.dk + ((long long)b * a.Lk + j) * a.lddk + h * 32, v);
          tmem_ld_32x32(lane_addr + DV_COL + kt * 64 + e * 32, v);
          tmem_ld_wait();
          if (j < a.Lk) store_row32_bf16(a.dv + ((long long)b * a.Lk + j) * a.lddv + h * 32, v);
        }
      }
      if (stamp) a.stamps[5 + 6 * it] = clock64();          // epilogues done
      tc_fence_before();
      mbar_arrive(&sh.dq_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tdb

using namespace tdb;

static long long* g_mha_tc_stamps = nullptr;
extern "C" int tdb_mha_tc_set_timing_buffer(void* buf) {      // measurement hook: device int64 [16] (NULL = off), see TcBwdParams::stamps
  g_mha_tc_stamps = (long long*)buf;
  return TDB_OK;
}
static int g_mha_tc = -1;
extern "C" int tdb_mha_set_tc(int level) {      // 0 = CUDA-core kernels, 1 = tcgen05 forward, 2 = tcgen05 forward + backward
  g_mha_tc = level < 0 ? 0 : (level > 2 ? 2 : level);
  return TDB_OK;
}
extern "C" int tdb_mha_tc_enabled(void) {
  if (g_mha_tc < 0) {
    const char* e = getenv("TDB_MHA_TC");
    g_mha_tc = e ? atoi(e) : 2;         // default: tcgen05 forward and backward
  }
  return g_mha_tc;
}

// Shapes the tcgen05 kernels take.  One-query sequences (unfused cross-attention of tiny memories) stay on the CUDA-core row kernel
// (a 128-row MMA tile per query row would be > 99 % padding).
extern "C" int tdb_mha_tc_supported(int H, int Lq, int Lk) { return (H % 2 == 0 && Lq >= 2 && Lq <= TC_MAXL && Lk >= 1 && Lk <= TC_MAXL) ? 1 : 0; }

extern "C" int tdb_mha_tc_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const uint8_t* kpm, void* o, int64_t ldo, float* p, float* pdrop, const int64_t* drop_seed,
                              int64_t drop_site, float drop_p, int B, int H, int Lq, int Lk, float scale, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(q && k && v && o && p && B > 0, "tdb_mha_tc_fwd: null argument");
  TDB_REQUIRE(tdb_mha_tc_supported(H, Lq, Lk), "tdb_mha_tc_fwd: unsupported shape H=%d Lq=%d Lk=%d (even H, 2..%d queries, <= %d keys)", H, Lq, Lk, TC_MAXL, TC_MAXL);
  TDB_REQUIRE(ldo % 8 == 0 && ((uintptr_t)o & 15) == 0, "tdb_mha_tc_fwd: o must be 16-byte aligned with ldo %% 8 == 0");
  TDB_REQUIRE(!drop_seed || (drop_p > 0.f && drop_p < 1.f), "tdb_mha_tc_fwd: dropout needs 0 < p < 1");
  TcParams a;
  a.kpm = kpm;
  a.o = (bf16*)o;
  a.ldo = ldo;
  a.p = p;
  a.pdrop = drop_seed ? pdrop : nullptr;
  a.drop_seed = (const long long*)drop_seed;
  a.drop_site = (unsigned long long)drop_site;
  a.drop_thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  a.keep_scale = drop_seed ? 1.f / (1.f - drop_p) : 1.f;
  a.B = B;
  a.H = H;
  a.Lq = Lq;
  a.Lk = Lk;
  a.LKP = (Lk + 31) / 32 * 32;
  a.MT = (Lq + 127) / 128;
  a.o_col = a.LKP <= 192 ? a.LKP : 256;
  a.tmem_cols = a.LKP <= 192 ? 256 : 512;
  a.scale_log2 = scale * 1.4426950408889634f;
  CUtensorMap tmQ, tmK, tmV;
  if ((rc = tdb_make_tmap_bf16(&tmQ, q, (int64_t)B * Lq, (int64_t)H * 32, ldq, 128))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmK, k, (int64_t)B * Lk, (int64_t)H * 32, ldk, a.LKP))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmV, v, (int64_t)B * Lk, (int64_t)H * 32, ldv, a.LKP))) return rc;
  const int pblocks = (a.LKP + 63) / 64;
  const size_t smem = 1024 + 16384 + 2 * (size_t)a.LKP * 128 + (size_t)pblocks * 16384 + sizeof(TcSmem);
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(mha_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  TDB_REQUIRE(smem <= 160 * 1024, "tdb_mha_tc_fwd: shared memory %zu", smem);
  TDB_CHECK_CUDA(tdb_launch(mha_tc_fwd_kernel, dim3(B * H * a.MT), dim3(TC_THREADS), smem, (cudaStream_t)stream_, tmQ, tmK, tmV, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}

extern "C" int tdb_mha_tc_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout,
                              int64_t lddo, const float* p, const int64_t* drop_seed, int64_t drop_site, float drop_p,
                              const float* dpbar, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B,
                              int H, int Lq, int Lk, float scale, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(q && k && v && dout && p && dq && dk && dv && B > 0, "tdb_mha_tc_bwd: null argument");
  TDB_REQUIRE(tdb_mha_tc_supported(H, Lq, Lk), "tdb_mha_tc_bwd: unsupported shape H=%d Lq=%d Lk=%d", H, Lq, Lk);
  TDB_REQUIRE((((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
              "tdb_mha_tc_bwd: gradient buffers must be 16-byte aligned with row strides %% 8 == 0");
  TDB_REQUIRE(!drop_seed || (drop_p > 0.f && drop_p < 1.f), "tdb_mha_tc_bwd: dropout needs 0 < p < 1");
  TcBwdParams a;
  a.p = p;
  a.drop_seed = (const long long*)drop_seed;
  a.drop_site = (unsigned long long)drop_site;
  a.drop_thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  a.keep_scale = drop_seed ? 1.f / (1.f - drop_p) : 1.f;
  a.dpbar = dpbar;
  a.dq = (bf16*)dq;
  a.dk = (bf16*)dk;
  a.dv = (bf16*)dv;
  a.lddq = lddq;
  a.lddk = lddk;
  a.lddv = lddv;
  a.B = B;
  a.H = H;
  a.Lq = Lq;
  a.Lk = Lk;
  a.LKP = (Lk + 31) / 32 * 32;
  a.MT = (Lq + 127) / 128;
  a.KT = (a.LKP + 127) / 128;
  a.nblk = (a.LKP + 63) / 64;
  a.scale = scale;
  a.stamps = g_mha_tc_stamps;
  CUtensorMap tmQ, tmK, tmV, tmdO;
  if ((rc = tdb_make_tmap_bf16(&tmQ, q, (int64_t)B * Lq, (int64_t)H * 32, ldq, 128))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmdO, dout, (int64_t)B * Lq, (int64_t)H * 32, lddo, 128))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmK, k, (int64_t)B * Lk, (int64_t)H * 32, ldk, a.LKP))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmV, v, (int64_t)B * Lk, (int64_t)H * 32, ldv, a.LKP))) return rc;
  // 16 KB tail after the Pd blocks: transpose staging; with an odd block count it is also where the MN-major reads of the missing
  // block land (their output rows are keys beyond the sequence, never stored)
  a.v_reload = a.nblk == 4 ? 1 : 0;      // > 192 keys: staging aliases the V tile (re-fetched per query tile), no tail
  const size_t smem = 1024 + 2 * 16384 + 2 * (size_t)a.LKP * 128 + 2 * (size_t)a.nblk * 16384 + (a.v_reload ? 0 : 16384) + sizeof(TcBwdSmem);
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(mha_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  TDB_REQUIRE(smem <= 227 * 1024, "tdb_mha_tc_bwd: shared memory %zu", smem);
  TDB_CHECK_CUDA(tdb_launch(mha_tc_bwd_kernel, dim3(B * H), dim3(TC_THREADS), smem, (cudaStream_t)stream_, tmQ, tmK, tmV, tmdO, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}
