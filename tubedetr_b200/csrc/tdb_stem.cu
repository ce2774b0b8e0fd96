// tubedetr_b200 -- fused ResNet stem: conv 7x7 / 2 / pad 3 (3 -> 64) + FrozenBatchNorm + ReLU + max-pool 3x3 / 2 / pad 1, straight from
// fp32 NCHW frames to bf16 NHWC pixel rows [N * H2 * W2][64].
//
// Reference: models/backbone.py:97-105 -> torchvision resnet101 conv1 / bn1 (FrozenBatchNorm2d, backbone.py:60-70) / relu / maxpool.
// The unfused path (tdb_stem_im2col + tdb_gemm + tdb_maxpool3x3s2) materialises a K = 192 im2col matrix (1.5 GB at 125 frames of
// 352 x 352) and the 176 x 176 x 64 conv output in HBM: 4.1 GB of traffic for 0.31 GB of algorithmic bytes.  Here nothing but the
// input frames and the pooled output touches HBM:
//   * one CTA = one tile of 8 x 22 pooled pixels = 17 x 45 conv pixels (765, six 128-row MMA tiles) = a 39 x 95 x 3 input patch;
//   * the patch is staged once in shared memory as bf16; the implicit-im2col A tile of each 128 conv pixels is assembled IN SHARED
//     MEMORY in the tensor core's 128-byte-swizzled K-major layout (K order (c, kh, kw padded to 8): one 16-byte chunk per (c, kh),
//     the pad column meets a zero weight) -- 21 chunk copies per conv pixel, no global traffic;
//   * tcgen05.mma (M = 128, N = 64, 11 K-steps) accumulates in TMEM (two accumulators: the epilogue of tile m overlaps the MMAs of
//     tile m + 1); the epilogue applies FrozenBN scale / shift + ReLU and parks the bf16 conv pixels in a shared-memory tile;
//   * the 3 x 3 / 2 max-pool reads that tile (conv pixels outside the image count as 0 = the identity of max over ReLU outputs)
//     and writes 128-byte pooled pixel rows.
// Warp roles: warps 0-7 = patch loader / A-tile builder / epilogue / pool (two threads per conv pixel row), warp 8 = TMEM allocator
// and single-thread MMA issuer.  Persistent grid (one CTA per SM, 195 KB of shared memory).
#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"
#include <type_traits>

void tdb_count_launch(int n);
int tdb_init_once();
int tdb_num_sms();

namespace tdb {

constexpr int ST_PH = 8, ST_PW = 22;                  // pooled pixels per tile
constexpr int ST_CR = 2 * ST_PH + 1, ST_CC = 2 * ST_PW + 1;   // conv pixels per tile: 17 x 45
constexpr int ST_NCP = ST_CR * ST_CC;                 // 765
constexpr int ST_MT = (ST_NCP + 127) / 128;           // 6 MMA tiles
constexpr int ST_IR = 2 * ST_CR + 5, ST_IC = 2 * ST_CC + 5;   // input patch: 39 x 95
constexpr int ST_IP = 96;                             // patch row pitch (bf16 elements)
constexpr int ST_THREADS = 288;
constexpr int ST_SA = 3 * 16384, ST_SW = 3 * 8192, ST_SCONV = ST_MT * 128 * 128, ST_SPATCH = 3 * ST_IR * ST_IP * 2;
constexpr int ST_SMEM = 1024 + ST_SA + ST_SW + ST_SCONV + ST_SPATCH + 512 + 64;

struct StemParams {
  const float* x;        // [N][3][H][W] fp32
  const bf16* wk;        // [64][192] bf16, K order (c, kh, kw8), zero beyond the 7 real kw and beyond k = 168
  const float* scale;    // [64] FrozenBN scale, shift
  const float* shift;
  bf16* out;             // [N * H2 * W2][64], row pitch ldo elements
  long long ldo;
  int N, H, W, H1, W1, H2, W2, tiles_x, tiles_y, total;
};

struct StemBars {
  uint64_t a_full, mma_done[2];
  uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(ST_THREADS, 1) stem_fused_kernel(const __grid_constant__ StemParams a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW = sA + ST_SA;
  uint8_t* sConv = sW + ST_SW;
  bf16* sPatch = reinterpret_cast<bf16*>(sConv + ST_SCONV);
  float* sScale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sPatch) + ST_SPATCH);
  float* sShift = sScale + 64;
  StemBars& sh = *reinterpret_cast<StemBars*>(sShift + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&sh.a_full, 256);
    mbar_init(&sh.mma_done[0], 1);
    mbar_init(&sh.mma_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(&sh.tmem_slot, 128);
    tmem_relinquish();
  }
  // constants (weights, FrozenBN vectors are frozen buffers: not written by the previous kernel): staged before the grid dependency wait
  for (int i = tid; i < 64 * 24; i += ST_THREADS) {
    const int n = i / 24, j = i % 24;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(a.wk + n * 192 + j * 8));
    *reinterpret_cast<uint4*>(sW + (j >> 3) * 8192 + n * 128 + (((j & 7) ^ (n & 7)) * 16)) = w;
  }
  for (int i = tid; i < ST_SA / 16; i += ST_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);   // chunks 21..23 stay zero
  if (tid < 64) {
    sScale[tid] = __ldg(a.scale + tid);
    sShift[tid] = __ldg(a.shift + tid);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int my_tiles = (a.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      const uint64_t k_hi = umma_smem_desc(0, 16, 1024);
      const uint32_t aa = smem_u32(sA), wa = smem_u32(sW);
      for (int g = 0; g < my_tiles * ST_MT; ++g) {
        mbar_wait(&sh.a_full, (uint32_t)(g & 1), 61);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)((g & 1) * 64);
#pragma unroll
        for (int ks = 0; ks < 11; ++ks)
          umma_bf16(acc, umma_desc_at(k_hi, aa + (ks >> 2) * 16384 + (ks & 3) * 32), umma_desc_at(k_hi, wa + (ks >> 2) * 8192 + (ks & 3) * 32),
                    idesc, ks > 0 ? 1u : 0u);
        umma_commit(&sh.mma_done[g & 1]);
      }
    }
  } else {
    const int p = tid & 127, half = tid >> 7;            // conv pixel row of the MMA tile, which half of the work of that row
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32);
    // FrozenBN scale / shift of this thread's 32 channels stay in registers for the whole kernel (64 shared-memory loads per
    // epilogue otherwise: the kernel is instruction-issue bound, profiles/r02_ncu_kernels.txt)
    float rsc[32], rsh[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      rsc[i] = sScale[half * 32 + i];
      rsh[i] = sShift[half * 32 + i];
    }
    int g = 0;
    for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x) {
      const int n = tile / (a.tiles_x * a.tiles_y);
      const int trem = tile - n * (a.tiles_x * a.tiles_y);
      const int ty = trem / a.tiles_x, tx = trem - ty * a.tiles_x;
      const int py0 = ty * ST_PH, px0 = tx * ST_PW;
      const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;     // first conv pixel of the tile (pool pad 1)
      const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;     // first input pixel of the patch (conv pad 3)
      // ---- input patch: fp32 NCHW -> bf16 [3][39][96] (zero outside the frame), coalesced row segments
      const float* xn = a.x + (long long)n * 3 * a.H * a.W;
      // every warp owns rows warp, warp + 8, ...: all of its global loads are issued before the first conversion / store so that
      // ~45 loads per lane are in flight together (a row-at-a-time loop pays one DRAM latency per row)
      constexpr int ROWS_PER_WARP = (3 * ST_IR + 7) / 8;      // 15
      float pv[ROWS_PER_WARP][3];
#pragma unroll
      for (int k = 0; k < ROWS_PER_WARP; ++k) {
        const int row = warp + 8 * k;
        const int c = row / ST_IR, r = row - c * ST_IR;
        const int iy = iy0 + r;
        const bool rowok = row < 3 * ST_IR && iy >= 0 && iy < a.H;
        const float* src = xn + ((long long)c * a.H + iy) * a.W;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int q = lane + 32 * u, ix = ix0 + q;
          pv[k][u] = (rowok && q < ST_IC && ix >= 0 && ix < a.W) ? __ldg(src + ix) : 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < ROWS_PER_WARP; ++k) {
        const int row = warp + 8 * k;
        if (row < 3 * ST_IR) {
#pragma unroll
          for (int u = 0; u < 3; ++u) sPatch[row * ST_IP + lane + 32 * u] = __float2bfloat16_rn(pv[k][u]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      auto epilogue = [&](int m, int gg) {
        // conv pixels [128 m, 128 m + 128): FrozenBN + ReLU -> bf16 rows of the conv tile (16-byte chunks XOR-swizzled with the row)
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + (uint32_t)((gg & 1) * 64), v);
        tmem_ld_wait();
        const int cp = m * 128 + p;
        const int ry = cp / ST_CC, rx = cp - ry * ST_CC;
        const int cy = cy0 + ry, cx = cx0 + rx;
        const bool ok = cp < ST_NCP && cy >= 0 && cy < a.H1 && cx >= 0 && cx < a.W1;
        uint8_t* rowp = sConv + cp * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ch = q4 * 8 + 2 * u;
            float f0 = fmaxf(fmaf(__uint_as_float(v[ch]), rsc[ch], rsh[ch]), 0.f);
            float f1 = fmaxf(fmaf(__uint_as_float(v[ch + 1]), rsc[ch + 1], rsh[ch + 1]), 0.f);
            w[u] = ok ? pack_bf16x2(f0, f1) : 0u;
          }
          *reinterpret_cast<uint4*>(rowp + (((half * 4 + q4) ^ (cp & 7)) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      };
      for (int m = 0; m < ST_MT; ++m, ++g) {
        if (g > 0) mbar_wait(&sh.mma_done[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1), 62);     // A buffer free, accumulator g - 1 complete
        // ---- A tile of conv pixels [128 m, 128 m + 128): chunk j = (c, kh) holds patch[c][2 ry + kh][2 rx .. 2 rx + 7]
        const int cp = m * 128 + p;
        if (cp < ST_NCP) {
          const int ry = cp / ST_CC, rx = cp - ry * ST_CC;
          // fully unrolled over the chunk index (two instantiations, one per half): (c, kh), the k-block and the chunk slot are
          // compile-time constants, so a chunk costs 4 LDS + 1 STS + the XOR with the row's swizzle phase
          const uint32_t* src0 = reinterpret_cast<const uint32_t*>(sPatch + (2 * ry) * ST_IP + 2 * rx);
          uint8_t* dst0 = sA + p * 128;
          const int p7 = p & 7;
          auto build = [&](auto H) {
            constexpr int HALF = decltype(H)::value;
#pragma unroll
            for (int jj = 0; jj < (HALF ? 10 : 11); ++jj) {
              constexpr int dummy = 0;
              (void)dummy;
              const int j = HALF * 11 + jj;
              const int c = j / 7, kh = j - c * 7;
              const uint32_t* src = src0 + ((c * ST_IR + kh) * ST_IP) / 2;
              const uint4 w = make_uint4(src[0], src[1], src[2], src[3]);
              *reinterpret_cast<uint4*>(dst0 + (j >> 3) * 16384 + (((j & 7) ^ p7) * 16)) = w;
            }
          };
          if (half) build(std::integral_constant<int, 1>{});
          else build(std::integral_constant<int, 0>{});
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&sh.a_full);
        if (m > 0) {
          tc_fence_after();
          epilogue(m - 1, g - 1);
        }
      }
      mbar_wait(&sh.mma_done[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1), 63);
      tc_fence_after();
      epilogue(ST_MT - 1, g - 1);
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- 3 x 3 / 2 max-pool over the conv tile -> pooled pixel rows (item = pooled pixel x 8-channel chunk)
      for (int item = tid; item < ST_PH * ST_PW * 8; item += 256) {
        const int pp = item >> 3, ch8 = item & 7;
        const int py = pp / ST_PW, px = pp - py * ST_PW;
        const int oy = py0 + py, ox = px0 + px;
        if (oy >= a.H2 || ox >= a.W2) continue;
        __nv_bfloat162 mx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int cp = (2 * py + dy) * ST_CC + 2 * px + dx;
            const uint4 w = *reinterpret_cast<const uint4*>(sConv + cp * 128 + ((ch8 ^ (cp & 7)) * 16));
            mx[0] = __hmax2(mx[0], *reinterpret_cast<const __nv_bfloat162*>(&w.x));
            mx[1] = __hmax2(mx[1], *reinterpret_cast<const __nv_bfloat162*>(&w.y));
            mx[2] = __hmax2(mx[2], *reinterpret_cast<const __nv_bfloat162*>(&w.z));
            mx[3] = __hmax2(mx[3], *reinterpret_cast<const __nv_bfloat162*>(&w.w));
          }
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&mx[0]);
        o.y = *reinterpret_cast<uint32_t*>(&mx[1]);
        o.z = *reinterpret_cast<uint32_t*>(&mx[2]);
        o.w = *reinterpret_cast<uint32_t*>(&mx[3]);
        *reinterpret_cast<uint4*>(a.out + (((long long)n * a.H2 + oy) * a.W2 + ox) * a.ldo + ch8 * 8) = o;
      }
      // the next tile's patch load only touches sPatch (last read by the A builds above); its epilogues write sConv after the
      // bar.sync that follows the patch load, i.e. after every thread has left this pool loop
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace tdb

using namespace tdb;

extern "C" int tdb_stem_fused_ld(const float* x, const void* wk, const float* scale, const float* shift, void* out, int64_t ldo, int N, int H,
                                 int W, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(x && wk && scale && shift && out && N > 0 && H >= 7 && W >= 7 && ldo >= 64 && ldo % 8 == 0, "tdb_stem_fused: bad args");
  TDB_REQUIRE((((uintptr_t)wk | (uintptr_t)out) & 15) == 0, "tdb_stem_fused: wk / out must be 16-byte aligned");
  StemParams a;
  a.x = x;
  a.wk = (const bf16*)wk;
  a.scale = scale;
  a.shift = shift;
  a.out = (bf16*)out;
  a.ldo = ldo;
  a.N = N;
  a.H = H;
  a.W = W;
  a.H1 = (H + 6 - 7) / 2 + 1;
  a.W1 = (W + 6 - 7) / 2 + 1;
  a.H2 = (a.H1 + 2 - 3) / 2 + 1;
  a.W2 = (a.W1 + 2 - 3) / 2 + 1;
  a.tiles_y = (a.H2 + ST_PH - 1) / ST_PH;
  a.tiles_x = (a.W2 + ST_PW - 1) / ST_PW;
  const long long total = (long long)N * a.tiles_x * a.tiles_y;
  TDB_REQUIRE(total < (1ll << 30), "tdb_stem_fused: too many tiles");
  a.total = (int)total;
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    attr = true;
  }
  const int sms = tdb_num_sms();
  const int grid = a.total < sms ? a.total : sms;
  TDB_CHECK_CUDA(tdb_launch(stem_fused_kernel, dim3(grid), dim3(ST_THREADS), (size_t)ST_SMEM, (cudaStream_t)stream_, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}

extern "C" int tdb_stem_fused(const float* x, const void* wk, const float* scale, const float* shift, void* out, int N, int H, int W,
                              void* stream_) {
  return tdb_stem_fused_ld(x, wk, scale, shift, out, 64, N, H, W, stream_);
}
