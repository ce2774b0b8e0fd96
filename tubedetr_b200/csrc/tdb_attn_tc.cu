// tubedetr_b200 -- self-attention core on tcgen05 (encoder spatial attention, decoder temporal self-attention; head_dim 32).
//
// Reference: models/transformer.py:637-640 and 698-722 (nn.MultiheadAttention, need_weights path): S = (q hd^-1/2) k^T + mask,
// P = softmax(S), O = dropout(P) V.  Both contractions run on the 5th-generation tensor cores:
//   S[128 x LKP] = Q_h[128 x 32] K_h[LKP x 32]^T      accumulator in TMEM columns [0, LKP)        (2 tcgen05.mma, K = 16 each)
//   O[128 x 64]  = P[128 x LKP] V_pair[LKP x 64]      accumulator in TMEM columns [256, 320)      (LKP / 16 tcgen05.mma)
// One CTA = (sequence b, head PAIR g).  A head is only 32 channels = 64 bytes wide, half a 128-byte swizzle row, so Q, K and V
// are fetched by TMA as 64-column boxes covering the two heads of a pair; head e of the pair is addressed by starting the
// operand descriptors e * 64 bytes into the swizzled rows (the same mechanism as the +32-byte K steps of the GEMM main loop).
// V is used as an MN-major B operand with N = 64 (both heads' channels); only the 32 output columns of head e are kept.
// Softmax runs on 128 threads (thread = query row = TMEM lane): three light passes over the S accumulator (max, sum, normalise),
// probabilities go to global memory once (fp32, the backward kernels and the guided-attention loss read them) and, as bf16
// in the 128-byte-swizzled K-major layout, to shared memory as the A operand of the second contraction.
// Warp roles: warp 0 lane 0 = TMA + MMA issue, warp 2 = TMEM allocator, warps 4-7 = softmax / epilogue.
// Status: opt-in (TDB_MHA_TC=1 or tdb_mha_set_tc(1)); the CUDA-core kernels of tdb_attn.cu stay the default until this one is
// profiled in the step (DESIGN.md open items).  Backward keeps using the stored probabilities (tdb_attn.cu).
#include <stdlib.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);
int tdb_init_once();
int tdb_make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

namespace tdb {

constexpr int TC_THREADS = 256;
constexpr int TC_MAXL = 256;                 // queries and keys per sequence
constexpr int TC_O_COL = 256;                // TMEM column of the O accumulator

struct TcSmem {
  uint64_t full, s_full, p_full, o_full, o_free;
  uint32_t tmem_slot, pad;
  uint32_t dead[TC_MAXL / 4];   // one byte per key: 1 = padded key or beyond Lk (staged once per CTA)
};

struct TcParams {
  const uint8_t* kpm;   // [B][Lk] nonzero = masked key (may be null)
  bf16* o;              // [B*Lq][ldo]
  long long ldo;
  float* p;             // [B][H][Lq][Lk] normalised probabilities before dropout
  float* pdrop;         // same shape, after dropout (null unless requested)
  const uint8_t* keep;  // [B][H][Lq][Lk] or null
  float keep_scale;
  int B, H, Lq, Lk, LKP, MT;
  float scale;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
mha_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ TcParams a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: Q tiles [MT][128 x 128 B] | K [LKP x 128 B] | V [LKP x 128 B] | P blocks [LKP/64 rounded up][128 x 128 B] | barriers
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + a.MT * 16384;
  uint8_t* sV = sK + a.LKP * 128;
  uint8_t* sP = sV + a.LKP * 128;            // LKP is a multiple of 32, so every region stays 1024-byte aligned (LKP*128 % 4096 == 0)
  const int pblocks = (a.LKP + 63) >> 6;
  TcSmem& sh = *reinterpret_cast<TcSmem*>(sP + pblocks * 16384);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % (a.H >> 1), b = blockIdx.x / (a.H >> 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&sh.full, 1);
    mbar_init(&sh.s_full, 1);
    mbar_init(&sh.p_full, 128);
    mbar_init(&sh.o_full, 1);
    mbar_init(&sh.o_free, 128);
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

  const int iters = 2 * a.MT;                 // (head of the pair) x (128-row query tile)
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&sh.full, (uint32_t)(a.MT * 16384 + 2 * a.LKP * 128));
      for (int mt = 0; mt < a.MT; ++mt) tma_load_2d(sQ + mt * 16384, &tmQ, &sh.full, g * 64, b * a.Lq + mt * 128);
      tma_load_2d(sK, &tmK, &sh.full, g * 64, b * a.Lk);
      tma_load_2d(sV, &tmV, &sh.full, g * 64, b * a.Lk);
      mbar_wait(&sh.full, 0, 41);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_bf16(128, a.LKP, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);          // B = V, MN-major (N contiguous)
      const uint64_t k_hi = umma_smem_desc(0, 16, 1024);                // K-major, 128-byte swizzle
      const uint64_t mn_hi = umma_smem_desc(0, 8192, 1024);             // MN-major, 128-byte swizzle
      for (int it = 0; it < iters; ++it) {
        const int e = it / a.MT, mt = it - e * a.MT;
        const uint32_t ph = (uint32_t)(it & 1);
        // ---- S = Q_e K_e^T  (the S accumulator is free: the softmax threads signalled p_full of the previous iteration)
        const uint32_t qa = smem_u32(sQ + mt * 16384) + e * 64;
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
          umma_bf16(tmem_base + TC_O_COL, umma_desc_at(k_hi, a_addr), umma_desc_at(mn_hi, b_addr), idesc_o, ks > 0 ? 1u : 0u);
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
    {
      uint8_t* db = reinterpret_cast<uint8_t*>(sh.dead);
      for (int j = threadIdx.x - 128; j < a.LKP; j += 128) db[j] = (j >= a.Lk || (mk && mk[j])) ? 1 : 0;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    for (int it = 0; it < iters; ++it) {
      const int e = it / a.MT, mt = it - e * a.MT;
      const uint32_t ph = (uint32_t)(it & 1);
      const int h = 2 * g + e;
      const int i = mt * 128 + r;
      const bool valid = i < a.Lq;
      mbar_wait(&sh.s_full, ph, 44);
      tc_fence_after();
      // pass 1: row maximum of the masked, scaled scores
      float mx = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const bool dead = ((sh.dead[c * 8 + (t >> 2)] >> (8 * (t & 3))) & 0xffu) != 0;
          mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(v[t]) * a.scale);
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
          sum += (dead || mx == -INFINITY) ? 0.f : __expf(__uint_as_float(v[t]) * a.scale - mx);
        }
      }
      const float inv = 1.f / sum;            // all keys masked -> NaN, exactly like the reference softmax
      // the previous iteration's O = P V must have finished reading the P tile before it is overwritten
      if (it > 0) mbar_wait(&sh.o_full, ph ^ 1, 45);
      // pass 3: normalised probabilities -> global (fp32) and, after dropout, bf16 into the swizzled K-major A tile
      const long long prow = (((long long)b * a.H + h) * a.Lq + i) * a.Lk;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + c * 32, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int j = c * 32 + t;
          const bool dead = ((sh.dead[c * 8 + (t >> 2)] >> (8 * (t & 3))) & 0xffu) != 0;
          float pn = (dead || mx == -INFINITY) ? 0.f : __expf(__uint_as_float(v[t]) * a.scale - mx);
          pn = (j < a.Lk) ? pn * inv : 0.f;
          if (valid && j < a.Lk) {
            a.p[prow + j] = pn;
            if (a.keep) {
              pn = a.keep[prow + j] ? pn * a.keep_scale : 0.f;
              if (a.pdrop) a.pdrop[prow + j] = pn;
            }
          }
          pv[t] = valid ? pn : 0.f;
        }
        // 32 columns = 4 chunks of 16 bytes in block (c >> 1), chunk index (c & 1) * 4 + q, XOR-swizzled with the row
        uint8_t* blk = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = ((c & 1) * 4 + q4) ^ (r & 7);
          uint4 w;
          w.x = pack_bf16x2(pv[q4 * 8 + 0], pv[q4 * 8 + 1]);
          w.y = pack_bf16x2(pv[q4 * 8 + 2], pv[q4 * 8 + 3]);
          w.z = pack_bf16x2(pv[q4 * 8 + 4], pv[q4 * 8 + 5]);
          w.w = pack_bf16x2(pv[q4 * 8 + 6], pv[q4 * 8 + 7]);
          *reinterpret_cast<uint4*>(blk + chunk * 16) = w;
        }
      }
      fence_proxy_async();                    // generic-proxy writes of the P tile -> visible to the tensor core (async proxy)
      tc_fence_before();                      // our tcgen05.ld of S are done before the next S MMA may overwrite it
      mbar_arrive(&sh.p_full);
      // ---- O epilogue: the 32 channels of head e
      mbar_wait(&sh.o_full, ph, 46);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + TC_O_COL + e * 32, v);
        tmem_ld_wait();
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(a.o + ((long long)b * a.Lq + i) * a.ldo + h * 32);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            dst[q4] = make_uint4(pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]), __uint_as_float(v[q4 * 8 + 1])),
                                 pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]), __uint_as_float(v[q4 * 8 + 3])),
                                 pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]), __uint_as_float(v[q4 * 8 + 5])),
                                 pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]), __uint_as_float(v[q4 * 8 + 7])));
        }
      }
      tc_fence_before();
      mbar_arrive(&sh.o_free);
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

static int g_mha_tc = -1;
extern "C" int tdb_mha_set_tc(int on) {
  g_mha_tc = on ? 1 : 0;
  return TDB_OK;
}
extern "C" int tdb_mha_tc_enabled(void) {
  if (g_mha_tc < 0) {
    const char* e = getenv("TDB_MHA_TC");
    g_mha_tc = e ? atoi(e) : 0;
  }
  return g_mha_tc;
}

extern "C" int tdb_mha_tc_supported(int H, int Lq, int Lk) { return (H % 2 == 0 && Lq >= 1 && Lq <= TC_MAXL && Lk >= 1 && Lk <= TC_MAXL) ? 1 : 0; }

extern "C" int tdb_mha_tc_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const uint8_t* kpm, void* o, int64_t ldo, float* p, const uint8_t* keep, float* pdrop,
                              float keep_scale, int B, int H, int Lq, int Lk, float scale, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(q && k && v && o && p && B > 0, "tdb_mha_tc_fwd: null argument");
  TDB_REQUIRE(tdb_mha_tc_supported(H, Lq, Lk), "tdb_mha_tc_fwd: unsupported shape H=%d Lq=%d Lk=%d (even H, <= %d queries / keys)", H, Lq, Lk, TC_MAXL);
  TDB_REQUIRE(ldo % 8 == 0 && ((uintptr_t)o & 15) == 0, "tdb_mha_tc_fwd: o must be 16-byte aligned with ldo %% 8 == 0");
  TcParams a;
  a.kpm = kpm;
  a.o = (bf16*)o;
  a.ldo = ldo;
  a.p = p;
  a.pdrop = pdrop;
  a.keep = keep;
  a.keep_scale = keep_scale;
  a.B = B;
  a.H = H;
  a.Lq = Lq;
  a.Lk = Lk;
  a.LKP = (Lk + 31) / 32 * 32;
  a.MT = (Lq + 127) / 128;
  a.scale = scale;
  CUtensorMap tmQ, tmK, tmV;
  if ((rc = tdb_make_tmap_bf16(&tmQ, q, (int64_t)B * Lq, (int64_t)H * 32, ldq, 128))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmK, k, (int64_t)B * Lk, (int64_t)H * 32, ldk, a.LKP))) return rc;
  if ((rc = tdb_make_tmap_bf16(&tmV, v, (int64_t)B * Lk, (int64_t)H * 32, ldv, a.LKP))) return rc;
  const int pblocks = (a.LKP + 63) / 64;
  const size_t smem = 1024 + (size_t)a.MT * 16384 + 2 * (size_t)a.LKP * 128 + (size_t)pblocks * 16384 + sizeof(TcSmem);
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(mha_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  TDB_REQUIRE(smem <= 200 * 1024, "tdb_mha_tc_fwd: shared memory %zu", smem);
  TDB_CHECK_CUDA(tdb_launch(mha_tc_fwd_kernel, dim3(B * (H / 2)), dim3(TC_THREADS), smem, (cudaStream_t)stream_, tmQ, tmK, tmV, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}
