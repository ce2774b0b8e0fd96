// tubedetr_b200 -- kernels of the text encoder (RoBERTa-base: 12 layers, d = 768, 12 heads of 64, a few dozen tokens per caption).
//
// Reference: models/transformer.py:130-135, 250-263 (HF RobertaModel as a library call).  At 20 tokens every GEMM has
// M = 20 rows: the linear layers run on tdb_gemm (one 128-row tile, weights streamed once) and their weight gradients -- rank-20
// outer-product sums -- on the CUDA-core kernel below (one launch writes dW and db; a split-K tensor-core GEMM + reduce + column sums
// would be four).  The attention core (L <= 128 tokens, head_dim 64) and the erf-GELU are small dedicated kernels.
#include <math.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);

namespace tdb {

// ------------------------------------------------------------------ erf GELU (HF hidden_act = "gelu"), bf16 in / out
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n2) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const float2 v = unpack_bf16x2(reinterpret_cast<const uint32_t*>(x)[i]);
  const float a = 0.5f * v.x * (1.f + erff(v.x * 0.70710678118654752f)), b = 0.5f * v.y * (1.f + erff(v.y * 0.70710678118654752f));
  reinterpret_cast<uint32_t*>(y)[i] = pack_bf16x2(a, b);
}
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx, long long n2) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const float2 v = unpack_bf16x2(reinterpret_cast<const uint32_t*>(x)[i]);
  const float2 g = unpack_bf16x2(reinterpret_cast<const uint32_t*>(dy)[i]);
  reinterpret_cast<uint32_t*>(dx)[i] = pack_bf16x2(g.x * gelu_grad(v.x), g.y * gelu_grad(v.y));
}

// ------------------------------------------------------------------ weight / bias gradient of a linear layer with FEW rows
// dW[n][k] = sum_r dy[r][n] x[r][k], db[n] = sum_r dy[r][n]; dy bf16 [R][N] (row stride ldy), x bf16 [R][K] (row stride ldx).
// Block = 4 output rows n x all K columns (thread = 2 adjacent k); rows summed in order (deterministic).
constexpr int SW_ROWS = 4;
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const bf16* __restrict__ dy, long long ldy, const bf16* __restrict__ x, long long ldx,
                                                           float* __restrict__ dW, float* __restrict__ db, int R, int N, int K) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sdy[];            // [R][SW_ROWS]
  const int n0 = blockIdx.x * SW_ROWS;
  for (int e = threadIdx.x; e < R * SW_ROWS; e += blockDim.x) {
    const int r = e / SW_ROWS, j = e - r * SW_ROWS;
    sdy[e] = (n0 + j < N) ? __bfloat162float(dy[(long long)r * ldy + n0 + j]) : 0.f;
  }
  __syncthreads();
  for (int k2 = threadIdx.x; k2 * 2 < K; k2 += blockDim.x) {
    float a0[SW_ROWS], a1[SW_ROWS];
#pragma unroll
    for (int j = 0; j < SW_ROWS; ++j) a0[j] = a1[j] = 0.f;
    for (int r = 0; r < R; ++r) {
      const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (long long)r * ldx + 2 * k2));
#pragma unroll
      for (int j = 0; j < SW_ROWS; ++j) {
        a0[j] += sdy[r * SW_ROWS + j] * xv.x;
        a1[j] += sdy[r * SW_ROWS + j] * xv.y;
      }
    }
#pragma unroll
    for (int j = 0; j < SW_ROWS; ++j)
      if (n0 + j < N) *reinterpret_cast<float2*>(dW + (long long)(n0 + j) * K + 2 * k2) = make_float2(a0[j], a1[j]);
  }
  if (db && threadIdx.x < SW_ROWS && n0 + threadIdx.x < N) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += sdy[r * SW_ROWS + threadIdx.x];
    db[n0 + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------ attention core, head_dim 64, L <= 128 (one CTA per (sequence, head))
// q, k, v, o: bf16 rows [B*L][>= H*64] with row strides; kpm [B][L] nonzero = padded key; p [B][H][L][L] fp32 probabilities before dropout
// (kept for the backward); attention dropout from the hash stream at the flat [B][H][L][L] index.
struct TextAttnParams {
  const bf16 *q, *k, *v, *dout;
  long long ldq, ldk, ldv, ldo, lddq, lddk, lddv;
  const uint8_t* kpm;
  bf16* o;
  float* p;
  bf16 *dq, *dk, *dv;
  const long long* drop_seed;
  unsigned long long drop_site;
  uint32_t drop_thr;
  float keep_scale;
  int B, H, L;
  float scale;
};

constexpr int TA_HD = 64;
constexpr int TA_PITCH = TA_HD + 1;

__device__ __forceinline__ void ta_load_tile(float* dst, const bf16* src, long long ld, int L) {
  // [L][64] bf16 rows -> fp32 [L][65]
  for (int e = threadIdx.x; e < L * (TA_HD / 2); e += blockDim.x) {
    const int r = e / (TA_HD / 2), c2 = e - r * (TA_HD / 2);
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + (long long)r * ld + 2 * c2));
    dst[r * TA_PITCH + 2 * c2] = v.x;
    dst[r * TA_PITCH + 2 * c2 + 1] = v.y;
  }
}

__global__ void __launch_bounds__(128) text_attn_fwd_kernel(const TextAttnParams a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float ts[];
  const int L = a.L, LP = L + 1;
  float* sq = ts;                       // [L][65]
  float* sk = sq + L * TA_PITCH;
  float* sv = sk + L * TA_PITCH;
  float* sp = sv + L * TA_PITCH;        // [L][L + 1] post-dropout probabilities
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ta_load_tile(sq, a.q + (long long)b * L * a.ldq + h * TA_HD, a.ldq, L);
  ta_load_tile(sk, a.k + (long long)b * L * a.ldk + h * TA_HD, a.ldk, L);
  ta_load_tile(sv, a.v + (long long)b * L * a.ldv + h * TA_HD, a.ldv, L);
  __syncthreads();
  const uint8_t* mk = a.kpm ? a.kpm + (long long)b * L : nullptr;
  const unsigned long long dbase = a.drop_seed ? drop_base(a.drop_seed, a.drop_site) : 0ull;
  for (int i = warp; i < L; i += 4) {
    float s[4], mx = -INFINITY;                       // keys lane, lane + 32, lane + 64, lane + 96
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      float d = -INFINITY;
      if (j < L && !(mk && mk[j])) {
        d = 0.f;
#pragma unroll 16
        for (int c = 0; c < TA_HD; ++c) d += sq[i * TA_PITCH + c] * sk[j * TA_PITCH + c];
        d *= a.scale;
      }
      s[u] = d;
      mx = fmaxf(mx, d);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = (s[u] == -INFINITY || mx == -INFINITY) ? 0.f : __expf(s[u] - mx);
      sum += s[u];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const long long prow = (((long long)b * a.H + h) * L + i) * L;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      if (j < L) {
        float pv = s[u] * inv;
        a.p[prow + j] = pv;
        if (a.drop_seed) pv = drop_keep(dbase, prow + j, a.drop_thr) ? pv * a.keep_scale : 0.f;
        sp[i * LP + j] = pv;
      }
    }
  }
  __syncthreads();
  for (int i = warp; i < L; i += 4) {
    float o0 = 0.f, o1 = 0.f;                          // channels lane, lane + 32
    for (int j = 0; j < L; ++j) {
      const float pv = sp[i * LP + j];
      o0 += pv * sv[j * TA_PITCH + lane];
      o1 += pv * sv[j * TA_PITCH + lane + 32];
    }
    bf16* orow = a.o + ((long long)b * L + i) * a.ldo + h * TA_HD;
    orow[lane] = __float2bfloat16(o0);
    orow[lane + 32] = __float2bfloat16(o1);
  }
}

__global__ void __launch_bounds__(128) text_attn_bwd_kernel(const TextAttnParams a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float ts[];
  const int L = a.L, LP = L + 1;
  float* sq = ts;                       // [L][65]
  float* sk = sq + L * TA_PITCH;
  float* sv = sk + L * TA_PITCH;
  float* sdo = sv + L * TA_PITCH;
  float* sds = sdo + L * TA_PITCH;      // [L][L + 1] dS (scaled)
  float* spd = sds + L * LP;            // [L][L + 1] post-dropout P
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ta_load_tile(sq, a.q + (long long)b * L * a.ldq + h * TA_HD, a.ldq, L);
  ta_load_tile(sk, a.k + (long long)b * L * a.ldk + h * TA_HD, a.ldk, L);
  ta_load_tile(sv, a.v + (long long)b * L * a.ldv + h * TA_HD, a.ldv, L);
  ta_load_tile(sdo, a.dout + (long long)b * L * a.ldo + h * TA_HD, a.ldo, L);
  __syncthreads();
  const unsigned long long dbase = a.drop_seed ? drop_base(a.drop_seed, a.drop_site) : 0ull;
  // rows: dP = dO V^T, dropout', rs, dS
  for (int i = warp; i < L; i += 4) {
    const long long prow = (((long long)b * a.H + h) * L + i) * L;
    float d[4], pj[4], rs = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      d[u] = 0.f;
      pj[u] = 0.f;
      if (j < L) {
        float t = 0.f;
#pragma unroll 16
        for (int c = 0; c < TA_HD; ++c) t += sdo[i * TA_PITCH + c] * sv[j * TA_PITCH + c];
        pj[u] = a.p[prow + j];
        bool kept = true;
        if (a.drop_seed) kept = drop_keep(dbase, prow + j, a.drop_thr);
        d[u] = kept ? t * a.keep_scale : 0.f;
        spd[i * LP + j] = kept ? pj[u] * a.keep_scale : 0.f;
        rs += pj[u] * d[u];
      }
    }
    rs = warp_sum(rs);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      if (j < L) sds[i * LP + j] = a.scale * pj[u] * (d[u] - rs);
    }
  }
  __syncthreads();
  // dQ[i] = sum_j dS[i][j] K[j];  dK[j] = sum_i dS[i][j] Q[i];  dV[j] = sum_i Pd[i][j] dO[i]   (thread = 2 channels of one row)
  for (int i = warp; i < L; i += 4) {
    float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float ds = sds[i * LP + j];
      q0 += ds * sk[j * TA_PITCH + lane];
      q1 += ds * sk[j * TA_PITCH + lane + 32];
      const float dst = sds[j * LP + i], pdt = spd[j * LP + i];       // transposed access: row j, column i
      k0 += dst * sq[j * TA_PITCH + lane];
      k1 += dst * sq[j * TA_PITCH + lane + 32];
      v0 += pdt * sdo[j * TA_PITCH + lane];
      v1 += pdt * sdo[j * TA_PITCH + lane + 32];
    }
    bf16* r0 = a.dq + ((long long)b * L + i) * a.lddq + h * TA_HD;
    bf16* r1 = a.dk + ((long long)b * L + i) * a.lddk + h * TA_HD;
    bf16* r2 = a.dv + ((long long)b * L + i) * a.lddv + h * TA_HD;
    r0[lane] = __float2bfloat16(q0);
    r0[lane + 32] = __float2bfloat16(q1);
    r1[lane] = __float2bfloat16(k0);
    r1[lane + 32] = __float2bfloat16(k1);
    r2[lane] = __float2bfloat16(v0);
    r2[lane + 32] = __float2bfloat16(v1);
  }
}

}  // namespace tdb

using namespace tdb;

#define TSTREAM ((cudaStream_t)stream_)
#define TLAUNCH_OK()                     \
  TDB_CHECK_CUDA(cudaGetLastError());    \
  tdb_count_launch(1);                   \
  return TDB_OK

extern "C" int tdb_gelu_fwd(const void* x, void* y, int64_t n, void* stream_) {
  TDB_REQUIRE(x && y && n > 0 && n % 2 == 0, "tdb_gelu_fwd: n must be even");
  TDB_CHECK_CUDA(tdb_launch(gelu_fwd_kernel, dim3((unsigned)((n / 2 + 255) / 256)), dim3(256), 0, TSTREAM, (const bf16*)x, (bf16*)y, (long long)(n / 2)));
  TLAUNCH_OK();
}
extern "C" int tdb_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream_) {
  TDB_REQUIRE(dy && x && dx && n > 0 && n % 2 == 0, "tdb_gelu_bwd: n must be even");
  TDB_CHECK_CUDA(tdb_launch(gelu_bwd_kernel, dim3((unsigned)((n / 2 + 255) / 256)), dim3(256), 0, TSTREAM, (const bf16*)dy, (const bf16*)x, (bf16*)dx,
                            (long long)(n / 2)));
  TLAUNCH_OK();
}

extern "C" int tdb_skinny_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* dW, float* db, int R, int N, int K, void* stream_) {
  TDB_REQUIRE(dy && x && dW && R > 0 && R <= 4096 && N > 0 && K > 0 && K % 2 == 0 && ldx % 2 == 0, "tdb_skinny_wgrad: bad args");
  const size_t smem = (size_t)R * SW_ROWS * sizeof(float);
  TDB_REQUIRE(smem <= 48 * 1024, "tdb_skinny_wgrad: too many rows (%d)", R);
  TDB_CHECK_CUDA(tdb_launch(skinny_wgrad_kernel, dim3((N + SW_ROWS - 1) / SW_ROWS), dim3(256), smem, TSTREAM, (const bf16*)dy, (long long)ldy,
                            (const bf16*)x, (long long)ldx, dW, db, R, N, K));
  TLAUNCH_OK();
}

static int text_attn_fill(TextAttnParams& a, int B, int H, int L, float scale, const int64_t* drop_seed, int64_t drop_site, float drop_p) {
  TDB_REQUIRE(B > 0 && H > 0 && L >= 1 && L <= 128, "tdb_text_attn: L=%d unsupported (1..128 tokens)", L);
  TDB_REQUIRE(!drop_seed || (drop_p > 0.f && drop_p < 1.f), "tdb_text_attn: dropout needs 0 < p < 1");
  a.B = B;
  a.H = H;
  a.L = L;
  a.scale = scale;
  a.drop_seed = (const long long*)drop_seed;
  a.drop_site = (unsigned long long)drop_site;
  a.drop_thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  a.keep_scale = drop_seed ? 1.f / (1.f - drop_p) : 1.f;
  return TDB_OK;
}

extern "C" int tdb_text_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* kpm, void* o,
                                 int64_t ldo, float* p, const int64_t* drop_seed, int64_t drop_site, float drop_p, int B, int H, int L,
                                 float scale, void* stream_) {
  TDB_REQUIRE(q && k && v && o && p, "tdb_text_attn_fwd: null argument");
  TDB_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0, "tdb_text_attn_fwd: row strides must be even");
  TextAttnParams a;
  memset(&a, 0, sizeof(a));
  int rc = text_attn_fill(a, B, H, L, scale, drop_seed, drop_site, drop_p);
  if (rc) return rc;
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.kpm = kpm; a.o = (bf16*)o; a.p = p;
  const size_t smem = sizeof(float) * ((size_t)3 * L * TA_PITCH + (size_t)L * (L + 1));
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(text_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  TDB_CHECK_CUDA(tdb_launch(text_attn_fwd_kernel, dim3(B * H), dim3(128), smem, TSTREAM, a));
  TLAUNCH_OK();
}

extern "C" int tdb_text_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout, int64_t lddo,
                                 const float* p, const int64_t* drop_seed, int64_t drop_site, float drop_p, void* dq, int64_t lddq, void* dk,
                                 int64_t lddk, void* dv, int64_t lddv, int B, int H, int L, float scale, void* stream_) {
  TDB_REQUIRE(q && k && v && dout && p && dq && dk && dv, "tdb_text_attn_bwd: null argument");
  TDB_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0 && lddo % 2 == 0, "tdb_text_attn_bwd: row strides must be even");
  TextAttnParams a;
  memset(&a, 0, sizeof(a));
  int rc = text_attn_fill(a, B, H, L, scale, drop_seed, drop_site, drop_p);
  if (rc) return rc;
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.dout = (const bf16*)dout;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = lddo;
  a.p = const_cast<float*>(p);
  a.dq = (bf16*)dq; a.dk = (bf16*)dk; a.dv = (bf16*)dv;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  const size_t smem = sizeof(float) * ((size_t)4 * L * TA_PITCH + (size_t)2 * L * (L + 1));
  TDB_REQUIRE(smem <= 200 * 1024, "tdb_text_attn_bwd: L=%d too long", L);
  static bool attr = false;
  if (!attr) {
    TDB_CHECK_CUDA(cudaFuncSetAttribute(text_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  TDB_CHECK_CUDA(tdb_launch(text_attn_bwd_kernel, dim3(B * H), dim3(128), smem, TSTREAM, a));
  TLAUNCH_OK();
}
