// tubedetr_b200 -- multi-head attention core for TubeDETR's three attention sites (head_dim = 32):
//   encoder spatial multi-modal self-attention  (reference models/transformer.py:637-640; S = HW + L <= 256 keys)
//   decoder temporal self-attention (TSA)       (models/transformer.py:698-722;  T <= 200 keys, head-mean P returned WITH grad)
//   decoder time-aligned cross-attention        (models/transformer.py:724-745;  1 query per frame, S keys)
// Math restated from torch F.multi_head_attention_forward (need_weights path): S = (q*hd^-1/2) k^T + (-inf on padded keys),
// P = softmax(S), O = P V, Pbar = mean_h P.  One CTA per (batch, head): K^T and V of that head live in shared memory
// (fp32, <= 64 KB), one warp per query row, probabilities written once (needed by backward and by the guided-attention loss).
// Round-1 implementation on CUDA cores in fp32 (the attention core is <1% of the step FLOPs); the projections around it
// run on tcgen05 through tdb_gemm.  DESIGN.md lists the tcgen05 fused KV-projection variant as the next step.
#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);

namespace tdb {

constexpr int HD = 32;
constexpr int kAttnThreads = 256;

// q/k/v: bf16 with row stride ld* (elements) and batch stride = L*ld*; head h occupies columns [h*32, h*32+32)
struct AttnParams {
  const bf16 *q, *k, *v;
  long long ldq, ldk, ldv;
  const uint8_t* kpm;  // [B][Lk] nonzero = masked key, may be null
  bf16* o;             // [B][Lq][H*32]
  long long ldo;
  float* p;            // [B][H][Lq][Lk]  softmax probabilities BEFORE dropout (backward needs them)
  float* pdrop;        // [B][H][Lq][Lk]  probabilities after dropout (= what torch returns / averages), null when no dropout
  const uint8_t* keep; // [B][H][Lq][Lk]  1 = kept, null = no dropout (attention dropout, reference models/transformer.py:613)
  float keep_scale;    // 1 / (1 - p_drop)
  int B, H, Lq, Lk;
  float scale;
};

// Shared-memory layout of the row kernels (fp32): [32][Lkp] transposed operand (odd pitch: conflict-free column reads),
// [Lk4][32] row-major operand (rows >= Lk zero), [warps][Lk4] per-warp probability row, [warps][32] per-warp query row.
constexpr int kMaxChunks = 16;           // Lk <= 512 keys in registers (32 lanes x 16 chunks); kernels are instantiated for
                                         // NC = 4, 5, 6, 8, 12, 16 chunks so that no predicated-off chunk costs issue slots

// stage head h of two bf16 [L][ld] matrices: T -> transposed [32][Lkp], R -> row-major [Lk4][32]; 16-byte global loads
// (thread = one 8-channel group of one row: a row of a head is 64 B = 4 groups)
__device__ __forceinline__ void stage_head(const bf16* __restrict__ tsrc, long long ldt, const bf16* __restrict__ rsrc, long long ldr,
                                           int Lk, int Lkp, int Lk4, float* __restrict__ tdst, float* __restrict__ rdst) {
  for (int i = threadIdx.x; i < Lk4 * 4; i += blockDim.x) {
    const int j = i >> 2, g = i & 3;
    float tv[8], rv[8];
    if (j < Lk) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(tsrc + (long long)j * ldt) + g);
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(rsrc + (long long)j * ldr) + g);
      float2 x;
      x = unpack_bf16x2(a.x); tv[0] = x.x; tv[1] = x.y;
      x = unpack_bf16x2(a.y); tv[2] = x.x; tv[3] = x.y;
      x = unpack_bf16x2(a.z); tv[4] = x.x; tv[5] = x.y;
      x = unpack_bf16x2(a.w); tv[6] = x.x; tv[7] = x.y;
      x = unpack_bf16x2(b.x); rv[0] = x.x; rv[1] = x.y;
      x = unpack_bf16x2(b.y); rv[2] = x.x; rv[3] = x.y;
      x = unpack_bf16x2(b.z); rv[4] = x.x; rv[5] = x.y;
      x = unpack_bf16x2(b.w); rv[6] = x.x; rv[7] = x.y;
#pragma unroll
      for (int t = 0; t < 8; ++t) tdst[(g * 8 + t) * Lkp + j] = tv[t];
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) rv[t] = 0.f;
    }
    float4* r4 = reinterpret_cast<float4*>(rdst + j * HD + g * 8);
    r4[0] = make_float4(rv[0], rv[1], rv[2], rv[3]);
    r4[1] = make_float4(rv[4], rv[5], rv[6], rv[7]);
  }
}

// s[c] = sum_d qrow[d] * T[d][c*32 + lane]  for the chunks c < nc (register blocked: one broadcast read of q per d)
// (chunks beyond Lk read in-bounds garbage of the next row / the following buffer; the callers overwrite those lanes)
template <int NC>
__device__ __forceinline__ void row_scores(const float* __restrict__ qrow, const float* __restrict__ tmat, int Lkp, int lane,
                                           float (&s)[NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) s[c] = 0.f;
#pragma unroll 8
  for (int d = 0; d < HD; ++d) {
    const float qv = qrow[d];
    const float* tr = tmat + d * Lkp + lane;
#pragma unroll
    for (int c = 0; c < NC; ++c) s[c] = fmaf(qv, tr[c * 32], s[c]);
  }
}

// acc(lane = d) = sum_j w[j] * R[j][d], four independent accumulators, w read as broadcast float4
__device__ __forceinline__ float row_combine(const float* __restrict__ w, const float* __restrict__ rmat, int Lk4, int lane) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int j = 0; j < Lk4; j += 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(w + j);
    const float* r = rmat + j * HD + lane;
    a0 = fmaf(w4.x, r[0], a0);
    a1 = fmaf(w4.y, r[HD], a1);
    a2 = fmaf(w4.z, r[2 * HD], a2);
    a3 = fmaf(w4.w, r[3 * HD], a3);
  }
  return (a0 + a1) + (a2 + a3);
}

template <int NC>
__global__ void __launch_bounds__(kAttnThreads) mha_fwd_kernel(const AttnParams a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int Lkp = a.Lk | 1;                 // odd pitch: conflict-free transposed K
  const int Lk4 = (a.Lk + 3) & ~3;
  const int nwarps = blockDim.x >> 5;
  float* kt = sm;                           // [32][Lkp]
  float* vs = kt + HD * Lkp + (4 - ((HD * Lkp) & 3)) % 4;   // [Lk4][32], 16-byte aligned
  float* prow = vs + Lk4 * HD;              // [warps][Lk4]
  float* qsm = prow + nwarps * Lk4;         // [warps][32]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_head(a.k + (long long)b * a.Lk * a.ldk + h * HD, a.ldk, a.v + (long long)b * a.Lk * a.ldv + h * HD, a.ldv, a.Lk, Lkp, Lk4, kt, vs);
  __syncthreads();
  const uint8_t* mk = a.kpm ? a.kpm + (long long)b * a.Lk : nullptr;
  float* pw = prow + warp * Lk4;
  float* qw = qsm + warp * HD;
  if (lane < Lk4 - a.Lk) pw[a.Lk + lane] = 0.f;          // zero tail of the probability row (read by row_combine)
  for (int i = blockIdx.y * nwarps + warp; i < a.Lq; i += gridDim.y * nwarps) {
    const bf16* qr = a.q + ((long long)b * a.Lq + i) * a.ldq + h * HD;
    qw[lane] = __bfloat162float(qr[lane]) * a.scale;
    __syncwarp();
    float s[NC];
    row_scores<NC>(qw, kt, Lkp, lane, s);
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = c * 32 + lane;
      if (j >= a.Lk || (mk && mk[j])) s[c] = -INFINITY;
      mx = fmaxf(mx, s[c]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      s[c] = (mx == -INFINITY || s[c] == -INFINITY) ? 0.f : __expf(s[c] - mx);
      sum += s[c];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;  // all keys masked -> NaN, exactly like the reference softmax
    const long long prow_off = (((long long)b * a.H + h) * a.Lq + i) * a.Lk;
    float* pg = a.p + prow_off;
    const uint8_t* kp = a.keep ? a.keep + prow_off : nullptr;
    float* pd = (a.keep && a.pdrop) ? a.pdrop + prow_off : nullptr;   // only needed when the head-mean weights are requested
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = c * 32 + lane;
      if (j < a.Lk) {
        float e = s[c];
        pg[j] = e * inv;
        if (kp) {                   // P <- P * keep / (1 - p); the context uses the dropped probabilities
          e = kp[j] ? e * a.keep_scale : 0.f;
          if (pd) pd[j] = e * inv;
        }
        pw[j] = e;
      }
    }
    __syncwarp();
    const float acc = row_combine(pw, vs, Lk4, lane);
    a.o[((long long)b * a.Lq + i) * a.ldo + h * HD + lane] = __float2bfloat16(acc * inv);
    __syncwarp();
  }
}

// Pbar[b][i][j] = mean_h P[b][h][i][j]
__global__ void head_mean_kernel(const float* __restrict__ p, float* __restrict__ pbar, int B, int H, long long LL) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * LL) return;
  long long b = idx / LL, r = idx - b * LL;
  float s = 0.f;
  for (int h = 0; h < H; ++h) s += p[(b * H + h) * LL + r];
  pbar[idx] = s / H;
}

struct AttnBwdParams {
  const bf16 *q, *k, *v, *dout;
  long long ldq, ldk, ldv, lddo;
  const float* p;      // [B][H][Lq][Lk] pre-dropout probabilities
  const uint8_t* keep; // dropout keep mask or null
  float keep_scale;
  float* pd_scratch;   // [B][H][Lq][Lk] dropped probabilities for the dV pass (only with dropout)
  const float* dpbar;  // [B][Lq][Lk] gradient of the head-mean (post-dropout) probabilities, may be null
  float* ds;           // [B][H][Lq][Lk] scratch: dS = P o (dP - rowsum(P o dP))
  bf16 *dq, *dk, *dv;  // same layout as q/k/v (strides lddq..)
  long long lddq, lddk, lddv;
  int B, H, Lq, Lk;
  float scale;
};

// row pass: one warp per query row: dP = dO V^T (+ dPbar/H), dS, dQ = scale * dS K
template <int NC>
__global__ void __launch_bounds__(kAttnThreads) mha_bwd_row_kernel(const AttnBwdParams a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int Lkp = a.Lk | 1;
  const int Lk4 = (a.Lk + 3) & ~3;
  const int nwarps = blockDim.x >> 5;
  float* vt = sm;                    // [32][Lkp]   V transposed
  float* ks = vt + HD * Lkp + (4 - ((HD * Lkp) & 3)) % 4;   // [Lk4][32]
  float* drow = ks + Lk4 * HD;       // [warps][Lk4]
  float* dosm = drow + nwarps * Lk4; // [warps][32]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_head(a.v + (long long)b * a.Lk * a.ldv + h * HD, a.ldv, a.k + (long long)b * a.Lk * a.ldk + h * HD, a.ldk, a.Lk, Lkp, Lk4, vt, ks);
  __syncthreads();
  float* dw = drow + warp * Lk4;
  float* dov = dosm + warp * HD;
  const float invH = 1.f / a.H;
  if (lane < Lk4 - a.Lk) dw[a.Lk + lane] = 0.f;
  for (int i = blockIdx.y * nwarps + warp; i < a.Lq; i += gridDim.y * nwarps) {
    dov[lane] = __bfloat162float(a.dout[((long long)b * a.Lq + i) * a.lddo + h * HD + lane]);
    __syncwarp();
    const long long roff = (((long long)b * a.H + h) * a.Lq + i) * a.Lk;
    const float* pr = a.p + roff;
    const float* dpb = a.dpbar ? a.dpbar + ((long long)b * a.Lq + i) * a.Lk : nullptr;
    const uint8_t* kr = a.keep ? a.keep + roff : nullptr;
    float dp[NC], pj[NC];
    row_scores<NC>(dov, vt, Lkp, lane, dp);
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = c * 32 + lane;
      pj[c] = 0.f;
      if (j < a.Lk) {
        float d = dp[c];
        if (dpb) d += dpb[j] * invH;
        if (kr) d = kr[j] ? d * a.keep_scale : 0.f;      // d(pre-dropout P) = d(post) * keep / (1 - p)
        pj[c] = pr[j];
        dp[c] = d;
        rs += pj[c] * d;
      }
    }
    rs = warp_sum(rs);
    float* dsg = a.ds + roff;
    float* pdg = kr ? a.pd_scratch + roff : nullptr;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = c * 32 + lane;
      if (j < a.Lk) {
        const float d = pj[c] > 0.f ? pj[c] * (dp[c] - rs) : 0.f;  // masked keys have P = 0 exactly
        dw[j] = d;
        dsg[j] = d;
        if (pdg) pdg[j] = kr[j] ? pj[c] * a.keep_scale : 0.f;
      }
    }
    __syncwarp();
    const float acc = row_combine(dw, ks, Lk4, lane);
    a.dq[((long long)b * a.Lq + i) * a.lddq + h * HD + lane] = __float2bfloat16(acc * a.scale);
    __syncwarp();
  }
}

// column pass: dV[j] = sum_i P[i][j] dO[i];  dK[j] = scale * sum_i dS[i][j] Q[i].
// CTA = (sequence, head, 64-key slab).  Query rows stream through shared memory in chunks of 32: the P / dS sub-tiles
// [32][64] are loaded with coalesced 256-byte row segments, dO / Q chunks [32][32] as fp32; warp w owns keys 8w..8w+7 of the
// slab and lane d owns head-dim d, so every shared-memory read is a broadcast (P, dS) or conflict-free (dO, Q).
constexpr int kColKeys = 64;
constexpr int kColRows = 32;
__global__ void __launch_bounds__(kAttnThreads) mha_bwd_col_kernel(const AttnBwdParams a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ps[kColRows][kColKeys];
  __shared__ float dss[kColRows][kColKeys];
  __shared__ float dos[kColRows][HD];
  __shared__ float qs[kColRows][HD];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int j0 = blockIdx.y * kColKeys;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* pb = (a.keep ? a.pd_scratch : a.p) + ((long long)b * a.H + h) * a.Lq * a.Lk;
  const float* dsb = a.ds + ((long long)b * a.H + h) * a.Lq * a.Lk;
  float av[8], ak[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) av[t] = ak[t] = 0.f;
  for (int i0 = 0; i0 < a.Lq; i0 += kColRows) {
    for (int e = threadIdx.x; e < kColRows * kColKeys; e += blockDim.x) {
      int r = e / kColKeys, c = e % kColKeys;
      int i = i0 + r, j = j0 + c;
      bool ok = i < a.Lq && j < a.Lk;
      ps[r][c] = ok ? __ldg(pb + (long long)i * a.Lk + j) : 0.f;
      dss[r][c] = ok ? __ldg(dsb + (long long)i * a.Lk + j) : 0.f;
    }
    for (int e = threadIdx.x; e < kColRows * HD; e += blockDim.x) {
      int r = e / HD, d = e % HD;
      int i = i0 + r;
      bool ok = i < a.Lq;
      dos[r][d] = ok ? __bfloat162float(a.dout[((long long)b * a.Lq + i) * a.lddo + h * HD + d]) : 0.f;
      qs[r][d] = ok ? __bfloat162float(a.q[((long long)b * a.Lq + i) * a.ldq + h * HD + d]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kColRows; ++r) {
      float dod = dos[r][lane], qd = qs[r][lane];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        av[t] += ps[r][warp * 8 + t] * dod;
        ak[t] += dss[r][warp * 8 + t] * qd;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    int j = j0 + warp * 8 + t;
    if (j < a.Lk) {
      a.dv[((long long)b * a.Lk + j) * a.lddv + h * HD + lane] = __float2bfloat16(av[t]);
      a.dk[((long long)b * a.Lk + j) * a.lddk + h * HD + lane] = __float2bfloat16(ak[t] * a.scale);
    }
  }
}

}  // namespace tdb

using namespace tdb;

static size_t attn_smem_bytes(int Lk, int threads) {
  int Lkp = Lk | 1, Lk4 = (Lk + 3) & ~3;
  return sizeof(float) * ((size_t)HD * Lkp + 4 + (size_t)Lk4 * HD + (size_t)(threads / 32) * (Lk4 + HD));
}
static int attn_grid_y(int BH, int Lq) {
  int per = (Lq + (kAttnThreads / 32) - 1) / (kAttnThreads / 32);
  int want = (4 * 148 + BH - 1) / BH;      // ~4 CTAs per SM (42 KB of shared memory each at Lk = 141): latency hiding across CTAs
  if (want < 1) want = 1;
  return want < per ? want : per;
}

template <int NC>
static int attn_init_nc() {
  TDB_CHECK_CUDA(cudaFuncSetAttribute(mha_fwd_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  TDB_CHECK_CUDA(cudaFuncSetAttribute(mha_bwd_row_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return TDB_OK;
}
static int attn_init() {
  static bool attr = false;
  if (!attr) {
    int rc;
    if ((rc = attn_init_nc<4>()) || (rc = attn_init_nc<5>()) || (rc = attn_init_nc<6>()) || (rc = attn_init_nc<8>()) ||
        (rc = attn_init_nc<12>()) || (rc = attn_init_nc<16>()))
      return rc;
    attr = true;
  }
  return TDB_OK;
}
// smallest instantiated chunk count covering Lk keys
#define TDB_ATTN_DISPATCH(Lk, KERNEL, ...)                                              \
  do {                                                                                  \
    const int nc_ = ((Lk) + 31) / 32;                                                   \
    if (nc_ <= 4) TDB_CHECK_CUDA(tdb_launch(KERNEL<4>, __VA_ARGS__));                   \
    else if (nc_ <= 5) TDB_CHECK_CUDA(tdb_launch(KERNEL<5>, __VA_ARGS__));              \
    else if (nc_ <= 6) TDB_CHECK_CUDA(tdb_launch(KERNEL<6>, __VA_ARGS__));              \
    else if (nc_ <= 8) TDB_CHECK_CUDA(tdb_launch(KERNEL<8>, __VA_ARGS__));              \
    else if (nc_ <= 12) TDB_CHECK_CUDA(tdb_launch(KERNEL<12>, __VA_ARGS__));            \
    else TDB_CHECK_CUDA(tdb_launch(KERNEL<16>, __VA_ARGS__));                           \
  } while (0)

extern "C" int tdb_mha_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const uint8_t* kpm, void* o, int64_t ldo, float* p, float* pbar, const uint8_t* keep, float* pdrop,
                           float keep_scale, int B, int H, int Lq, int Lk, float scale, void* stream_) {
  TDB_REQUIRE(!(keep && pbar) || pdrop, "tdb_mha_fwd: head-mean weights under dropout need the pdrop output");
  TDB_REQUIRE(q && k && v && o && p && B > 0 && H > 0 && Lq > 0 && Lk > 0, "tdb_mha_fwd: bad args");
  size_t smem = attn_smem_bytes(Lk, kAttnThreads);
  TDB_REQUIRE(Lk <= 32 * tdb::kMaxChunks && smem <= 200 * 1024, "tdb_mha_fwd: Lk=%d > %d keys", Lk, 32 * tdb::kMaxChunks);
  TDB_REQUIRE((((uintptr_t)k | (uintptr_t)v) & 15) == 0 && ldk % 8 == 0 && ldv % 8 == 0, "tdb_mha_fwd: k/v need 16-byte aligned head slices");
  int rc = attn_init();
  if (rc) return rc;
  AttnParams a{(const bf16*)q, (const bf16*)k, (const bf16*)v, ldq, ldk, ldv, kpm, (bf16*)o, ldo, p, pdrop, keep, keep_scale, B, H, Lq, Lk, scale};
  dim3 grid(B * H, attn_grid_y(B * H, Lq));
  TDB_ATTN_DISPATCH(Lk, mha_fwd_kernel, dim3(grid), dim3(kAttnThreads), smem, (cudaStream_t)stream_, a);
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  if (pbar) {
    long long LL = (long long)Lq * Lk;
    TDB_CHECK_CUDA(tdb_launch(head_mean_kernel, dim3((unsigned)(((long long)B * LL + 255) / 256)), dim3(256), 0, (cudaStream_t)stream_, keep ? pdrop : p, pbar, B, H, LL));
    TDB_CHECK_CUDA(cudaGetLastError());
    tdb_count_launch(1);
  }
  return TDB_OK;
}

extern "C" int tdb_head_mean(const float* p, float* pbar, int B, int H, int Lq, int Lk, void* stream_) {
  TDB_REQUIRE(p && pbar && B > 0 && H > 0 && Lq > 0 && Lk > 0, "tdb_head_mean: bad args");
  long long LL = (long long)Lq * Lk;
  TDB_CHECK_CUDA(tdb_launch(head_mean_kernel, dim3((unsigned)(((long long)B * LL + 255) / 256)), dim3(256), 0, (cudaStream_t)stream_, p, pbar, B, H, LL));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}

extern "C" int tdb_mha_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const void* dout, int64_t lddo, const float* p, const uint8_t* keep, float keep_scale,
                           float* pd_scratch, const float* dpbar, float* ds_scratch, void* dq,
                           int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk,
                           float scale, void* stream_) {
  TDB_REQUIRE(q && k && v && dout && p && ds_scratch && dq && dk && dv, "tdb_mha_bwd: null pointer");
  size_t smem = attn_smem_bytes(Lk, kAttnThreads);
  TDB_REQUIRE(Lk <= 32 * tdb::kMaxChunks && smem <= 200 * 1024, "tdb_mha_bwd: Lk=%d > %d keys", Lk, 32 * tdb::kMaxChunks);
  TDB_REQUIRE((((uintptr_t)k | (uintptr_t)v) & 15) == 0 && ldk % 8 == 0 && ldv % 8 == 0, "tdb_mha_bwd: k/v need 16-byte aligned head slices");
  int rc = attn_init();
  if (rc) return rc;
  TDB_REQUIRE(!keep || pd_scratch, "tdb_mha_bwd: dropout needs pd_scratch");
  AttnBwdParams a{(const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, ldq, ldk, ldv, lddo, p, keep, keep_scale,
                  pd_scratch, dpbar, ds_scratch, (bf16*)dq, (bf16*)dk, (bf16*)dv, lddq, lddk, lddv, B, H, Lq, Lk, scale};
  dim3 g1(B * H, attn_grid_y(B * H, Lq));
  TDB_ATTN_DISPATCH(Lk, mha_bwd_row_kernel, dim3(g1), dim3(kAttnThreads), smem, (cudaStream_t)stream_, a);
  dim3 g2(B * H, (Lk + kColKeys - 1) / kColKeys);
  TDB_CHECK_CUDA(tdb_launch(mha_bwd_col_kernel, dim3(g2), dim3(kAttnThreads), 0, (cudaStream_t)stream_, a));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(2);
  return TDB_OK;
}
