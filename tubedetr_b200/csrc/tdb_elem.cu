// tubedetr_b200 -- HBM-bound helper kernels around the GEMM: layout transforms for the ResNet stem / stride-2 convs,
// weight preparation, LayerNorm (+residual) forward/backward, column sums.  All coalesced, 16-byte vectorised where the
// layout allows; grid sized from the element count.  Reference call sites: SURVEY.md section 2.2 K1, K3 (stride-2), K4, K9, K14.
#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);

namespace tdb {

static inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ------------------------------------------------------------------ stem: fp32 NCHW frames -> bf16 im2col rows [N*Ho*Wo][192]
// column = c*49 + kh*7 + kw (torch weight flatten order), zero padded 147 -> 192; conv 7x7 / stride 2 / pad 3.
// One CTA per (frame, output row): the 7 input rows x 3 channels it needs are staged in shared memory with coalesced
// loads (zero padded by 3 on both sides), then every output pixel's 192-column row is written as 24 x 16-byte stores.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, bf16* __restrict__ col, int N, int H, int W,
                                                          int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float srow[];           // [3][7][W + 6]
  const int Wp = W + 6;
  const int n = blockIdx.x / Ho, ho = blockIdx.x % Ho;
  for (int e = threadIdx.x; e < 21 * Wp; e += blockDim.x) {
    int ck = e / Wp, wi = e % Wp - 3;
    int c = ck / 7, kh = ck % 7;
    int hi = ho * 2 - 3 + kh;
    float v = 0.f;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + (((long long)n * 3 + c) * H + hi) * W + wi);
    srow[e] = v;
  }
  __syncthreads();
  bf16* out = col + ((long long)n * Ho + ho) * Wo * 192;
  // thread -> fixed 16-byte group g of the 192-column row (its 8 shared-memory offsets are loop invariants), strided over the
  // output pixels: per 16-byte store 8 LDS + 4 packs, no index arithmetic; 24 consecutive threads write one 384-byte row
  if (threadIdx.x >= 240) return;
  const int g = threadIdx.x % 24, wo0 = threadIdx.x / 24;
  int off[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int cidx = g * 8 + i;
    int c = cidx / 49;
    int r = cidx - c * 49;
    int kh = r / 7, kw = r - kh * 7;
    off[i] = cidx < 147 ? (c * 7 + kh) * Wp + kw : -1;
  }
  for (int wo = wo0; wo < Wo; wo += 10) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = off[i] >= 0 ? srow[off[i] + wo * 2] : 0.f;
    *reinterpret_cast<uint4*>(out + (long long)wo * 192 + g * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// ------------------------------------------------------------------ 3x3 / stride 2 / pad 1 max pool, NHWC bf16, 8 channels per thread
__global__ void maxpool3x3s2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int cg = C / 8;
  long long total = (long long)N * Ho * Wo * cg;
  if (idx >= total) return;
  int c8 = (int)(idx % cg);
  long long t = idx / cg;
  int wo = (int)(t % Wo);
  t /= Wo;
  int ho = (int)(t % Ho);
  int n = (int)(t / Ho);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int kh = 0; kh < 3; ++kh) {
    int hi = ho * 2 - 1 + kh;
    if (hi < 0 || hi >= H) continue;
    for (int kw = 0; kw < 3; ++kw) {
      int wi = wo * 2 - 1 + kw;
      if (wi < 0 || wi >= W) continue;
      uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hi) * W + wi) * C + c8 * 8));
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], b.x); m[3] = fmaxf(m[3], b.y);
      m[4] = fmaxf(m[4], c.x); m[5] = fmaxf(m[5], c.y); m[6] = fmaxf(m[6], d.x); m[7] = fmaxf(m[7], d.y);
    }
  }
  *reinterpret_cast<uint4*>(y + (((long long)n * Ho + ho) * Wo + wo) * C + c8 * 8) =
      make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
}

// ------------------------------------------------------------------ 3x3 / stride 2 / pad 1 im2col, NHWC bf16 -> [N*Ho*Wo][9*C] (tap major)
__global__ void im2col3x3s2_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int N, int H, int W, int C, int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int cg = C / 8;
  long long total = (long long)N * Ho * Wo * 9 * cg;
  if (idx >= total) return;
  int c8 = (int)(idx % cg);
  long long t = idx / cg;
  int tap = (int)(t % 9);
  long long row = t / 9;
  int wo = (int)(row % Wo);
  long long t2 = row / Wo;
  int ho = (int)(t2 % Ho);
  int n = (int)(t2 / Ho);
  int kh = tap / 3, kw = tap - kh * 3;
  int hi = ho * 2 - 1 + kh, wi = wo * 2 - 1 + kw;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (hi >= 0 && hi < H && wi >= 0 && wi < W)
    u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hi) * W + wi) * C + c8 * 8));
  *reinterpret_cast<uint4*>(col + row * (9ll * C) + tap * C + c8 * 8) = u;
}

// transpose of the above as a gather (deterministic): dx[n,h,w,:] = sum over (ho,wo,tap) hitting (h,w); then ReLU mask by y>0
__global__ void col2im3x3s2_mask_kernel(const bf16* __restrict__ dcol, const bf16* __restrict__ ymask, bf16* __restrict__ dx,
                                        int N, int H, int W, int C, int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int cg = C / 8;
  long long total = (long long)N * H * W * cg;
  if (idx >= total) return;
  int c8 = (int)(idx % cg);
  long long pix = idx / cg;
  int w = (int)(pix % W);
  long long t = pix / W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int kh = 0; kh < 3; ++kh) {
    int hh = h + 1 - kh;  // = 2*ho
    if (hh < 0 || (hh & 1)) continue;
    int ho = hh >> 1;
    if (ho >= Ho) continue;
    for (int kw = 0; kw < 3; ++kw) {
      int ww = w + 1 - kw;
      if (ww < 0 || (ww & 1)) continue;
      int wo = ww >> 1;
      if (wo >= Wo) continue;
      long long row = ((long long)n * Ho + ho) * Wo + wo;
      uint4 u = __ldg(reinterpret_cast<const uint4*>(dcol + row * (9ll * C) + (kh * 3 + kw) * C + c8 * 8));
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
  if (ymask) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(ymask + pix * C + c8 * 8));
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    float mk[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = mk[i] > 0.f ? acc[i] : 0.f;
  }
  *reinterpret_cast<uint4*>(dx + pix * C + c8 * 8) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                                pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

// ------------------------------------------------------------------ stride-2 pixel subsample (1x1 stride-2 downsample conv input) and its transpose
__global__ void subsample2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long ldy, int N, int H, int W, int C, int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int cg = C / 8;
  long long total = (long long)N * Ho * Wo * cg;
  if (idx >= total) return;
  int c8 = (int)(idx % cg);
  long long t = idx / cg;
  int wo = (int)(t % Wo);
  t /= Wo;
  int ho = (int)(t % Ho);
  int n = (int)(t / Ho);
  *reinterpret_cast<uint4*>(y + (((long long)n * Ho + ho) * Wo + wo) * ldy + c8 * 8) =
      __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + 2 * ho) * W + 2 * wo) * C + c8 * 8));
}
__global__ void upsample2_zero_kernel(const bf16* __restrict__ y, bf16* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int cg = C / 8;
  long long total = (long long)N * H * W * cg;
  if (idx >= total) return;
  int c8 = (int)(idx % cg);
  long long pix = idx / cg;
  int w = (int)(pix % W);
  long long t = pix / W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  uint4 u = make_uint4(0, 0, 0, 0);
  if (!(h & 1) && !(w & 1) && (h >> 1) < Ho && (w >> 1) < Wo)
    u = __ldg(reinterpret_cast<const uint4*>(y + (((long long)n * Ho + (h >> 1)) * Wo + (w >> 1)) * C + c8 * 8));
  *reinterpret_cast<uint4*>(x + pix * C + c8 * 8) = u;
}

// ------------------------------------------------------------------ weights: fp32 torch [Cout][Cin][kh][kw] -> bf16 [Cout][Kpad] tap-major
// (column = tap*Cin + c), optional second copy pre-multiplied by rowscale[cout] (FrozenBN scale folded for dgrad).
// taps==49 (stem) keeps torch order c*49+tap.  Columns >= taps*Cin are zero.
__global__ void prep_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, bf16* __restrict__ out_scaled,
                                   const float* __restrict__ rowscale, int Cout, int Cin, int taps, int Kpad) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Cout * Kpad;
  if (idx >= total) return;
  int col = (int)(idx % Kpad);
  int co = (int)(idx / Kpad);
  float v = 0.f;
  if (col < taps * Cin) {
    if (taps == 49) {
      v = w[(long long)co * taps * Cin + col];
    } else {
      int tap = col / Cin, c = col - tap * Cin;
      v = w[((long long)co * Cin + c) * taps + tap];
    }
  }
  out[idx] = __float2bfloat16(v);
  if (out_scaled) out_scaled[idx] = __float2bfloat16(v * rowscale[co]);
}

// ------------------------------------------------------------------ fp32 -> bf16 cast with optional add (x + pos), 4 elements per thread
__global__ void cast_add_bf16_kernel(const float* __restrict__ x, const float* __restrict__ add, bf16* __restrict__ y, long long n4) {
  pdl_wait();
  pdl_trigger();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + idx);
  if (add) {
    float4 a = __ldg(reinterpret_cast<const float4*>(add) + idx);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  reinterpret_cast<uint2*>(y)[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// fp32 [rows][K] -> bf16 [rows][2K] = [hi | lo] with hi = bf16(x), lo = bf16(x - hi): hi + lo carries ~16 mantissa bits.
// Used for the WEIGHTS of the transformer's forward GEMMs (tdb_gemm with two taps: A x hi + A x lo): weight rounding is a
// systematic perturbation shared by every token and frame (it does not average out like activation rounding) and was measured to be
// 3x the activation-rounding share of the bf16 path's output error.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n4, int K4) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  const long long r = idx / K4;
  const int c4 = (int)(idx - r * K4);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + idx);
  const float h0 = __bfloat162float(__float2bfloat16_rn(v.x)), h1 = __bfloat162float(__float2bfloat16_rn(v.y));
  const float h2 = __bfloat162float(__float2bfloat16_rn(v.z)), h3 = __bfloat162float(__float2bfloat16_rn(v.w));
  uint2* row = reinterpret_cast<uint2*>(y + r * (long long)(8 * K4));
  row[c4] = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
  row[K4 + c4] = make_uint2(pack_bf16x2(v.x - h0, v.y - h1), pack_bf16x2(v.z - h2, v.w - h3));
}

// ------------------------------------------------------------------ LayerNorm over d=256 with fused residual add.  One warp per row.
// z = x + r (r optional, fp32);  y = (z - mean) * rstd * gamma + beta.  Writes y fp32, optional bf16 copy of y and of (y + pos).
template <int D>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ pos, float* __restrict__ y,
                                     bf16* __restrict__ y_bf, bf16* __restrict__ ypos_bf, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, int rows, float eps,
                                     const long long* __restrict__ drop_seed, unsigned long long drop_site, uint32_t drop_thr,
                                     float drop_scale) {
  pdl_wait();
  pdl_trigger();
  constexpr int PER = D / 32;
  int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * D;
  const unsigned long long dbase = drop_seed ? drop_base(drop_seed, drop_site) : 0ull;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = i * 32 + lane;
    float rv = r ? r[(long long)warp * D + c] : 0.f;
    if (drop_seed) rv = drop_keep(dbase, (long long)warp * D + c, drop_thr) ? rv * drop_scale : 0.f;   // residual dropout on r
    v[i] = xr[c] + rv;
    s += v[i];
  }
  float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    float dlt = v[i] - mean;
    q += dlt * dlt;
  }
  float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = i * 32 + lane;
    float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
    y[(long long)warp * D + c] = o;
    if (y_bf) y_bf[(long long)warp * D + c] = __float2bfloat16(o);
    if (ypos_bf) ypos_bf[(long long)warp * D + c] = __float2bfloat16(o + pos[(long long)warp * D + c]);
  }
  if (lane == 0) {
    if (mean_out) mean_out[warp] = mean;
    if (rstd_out) rstd_out[warp] = rstd;
  }
}

// backward: given dy (fp32), z recomputed from (x, r), mean, rstd: dz = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma
// writes dz (fp32; it is the gradient of BOTH x and r) and per-block partial dgamma/dbeta [blocks][2][D] for a fixed-order reduce.
template <int D, int WPB>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const bf16* __restrict__ dy2, const bf16* __restrict__ dy3,
                                     const float* __restrict__ x, const float* __restrict__ r,
                                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ dz, bf16* __restrict__ dz_bf, float* __restrict__ partial, int rows,
                                     const long long* __restrict__ drop_seed, unsigned long long drop_site, uint32_t drop_thr,
                                     float drop_scale, float* __restrict__ dr, bf16* __restrict__ dr_bf) {
  pdl_wait();
  pdl_trigger();
  constexpr int PER = D / 32;
  const unsigned long long dbase = drop_seed ? drop_base(drop_seed, drop_site) : 0ull;
  __shared__ float sg[WPB][D];      // WPB warps per block: 3 x WPB x D floats must stay under the 48 KB static limit (D = 768 -> 4 warps)
  __shared__ float sb[WPB][D];
  __shared__ float sr[WPB][D];
  int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[PER], ab[PER], ar[PER];     // column sums of dy * xhat (dgamma), dy (dbeta), and the gradient of r (bias gradient of the
#pragma unroll                       // linear layer that produced r: its separate column-sum launches disappear)
  for (int i = 0; i < PER; ++i) ag[i] = ab[i] = ar[i] = 0.f;
  for (int row = blockIdx.x * WPB + wib; row < rows; row += gridDim.x * WPB) {
    float m = mean[row], rs = rstd[row];
    float g[PER], xh[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int c = i * 32 + lane;
      long long o = (long long)row * D + c;
      float rv = r ? r[o] : 0.f;
      if (drop_seed) rv = drop_keep(dbase, o, drop_thr) ? rv * drop_scale : 0.f;
      float z = x[o] + rv;
      xh[i] = (z - m) * rs;
      float d = dy ? dy[o] : 0.f;
      if (dy2) d += __bfloat162float(dy2[o]);
      if (dy3) d += __bfloat162float(dy3[o]);
      g[i] = d * gamma[c];
      ag[i] += d * xh[i];
      ab[i] += d;
      s1 += g[i];
      s2 += g[i] * xh[i];
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      float o = rs * (g[i] - s1 - xh[i] * s2);
      const long long e = (long long)row * D + i * 32 + lane;
      dz[e] = o;
      if (drop_seed) {     // gradient of the dropped residual branch: d r = d z * keep / (1 - p); dz stays the gradient of x
        const float od = drop_keep(dbase, e, drop_thr) ? o * drop_scale : 0.f;
        if (dr) dr[e] = od;
        if (dr_bf) dr_bf[e] = __float2bfloat16(od);
        if (dz_bf) dz_bf[e] = __float2bfloat16(o);
        ar[i] += od;
      } else {
        if (dz_bf) dz_bf[e] = __float2bfloat16(o);
        ar[i] += o;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    sg[wib][i * 32 + lane] = ag[i];
    sb[wib][i * 32 + lane] = ab[i];
    sr[wib][i * 32 + lane] = ar[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int w = 0; w < WPB; ++w) {
      a += sg[w][c];
      b += sb[w][c];
      cc += sr[w][c];
    }
    partial[((long long)blockIdx.x * 3 + 0) * D + c] = a;
    partial[((long long)blockIdx.x * 3 + 1) * D + c] = b;
    partial[((long long)blockIdx.x * 3 + 2) * D + c] = cc;
  }
}
// out[c] (+)= sum_b partial[b][c] in fixed order; used for dgamma/dbeta (stride 2*D) and generic column sums.
// block = 32 columns x 8 part-groups: part-group y sums parts y, y+8, ... then the 8 groups are combined through shared memory.
__global__ void colsum_partials_kernel(const float* __restrict__ partial, int nparts, long long part_stride, int D,
                                       float* __restrict__ out, int accumulate) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < D)
    for (int b = threadIdx.y; b < nparts; b += 8) s += partial[b * part_stride + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// column sums of a bf16 [rows][N] matrix (bias gradients): stage 1 -> partial [nb][N] fp32.  Thread = 2 adjacent columns.
__global__ void colsum_bf16_kernel(const bf16* __restrict__ x, long long ld, int rows, int N, float* __restrict__ partial, int rows_per_block) {
  pdl_wait();
  pdl_trigger();
  int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (c >= N) return;
  int r0 = blockIdx.y * rows_per_block;
  int r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s0 = 0.f, s1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (long long)r * ld + c));
    s0 += v.x;
    s1 += v.y;
  }
  partial[(long long)blockIdx.y * N + c] = s0;
  partial[(long long)blockIdx.y * N + c + 1] = s1;
}

// 16-byte form of stage 1 (N % 8 == 0, ld % 8 == 0, 16-byte aligned base): block = 32 column-lanes (8 columns each = one
// 256-column slab) x 8 row-lanes; every thread keeps 4 independent 16-byte loads in flight, the 8 row-lanes are combined in
// shared memory in fixed order.  (The 4-byte form above ran at 0.3 TB/s, latency bound: profiles/r01_kernel_table.txt.)
__global__ void __launch_bounds__(256) colsum_bf16_v8_kernel(const bf16* __restrict__ x, long long ld, int rows, int N,
                                                             float* __restrict__ partial, int rows_per_block) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][32][9];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  if (c < N) {
    const bf16* base = x + c;
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (long long)(r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float2 a = unpack_bf16x2(v[u].x), b = unpack_bf16x2(v[u].y), cc = unpack_bf16x2(v[u].z), d = unpack_bf16x2(v[u].w);
        acc[0] += a.x;
        acc[1] += a.y;
        acc[2] += b.x;
        acc[3] += b.y;
        acc[4] += cc.x;
        acc[5] += cc.y;
        acc[6] += d.x;
        acc[7] += d.y;
      }
    }
    for (; r < r1; r += 8) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * ld));
      float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), cc = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
      acc[0] += a.x;
      acc[1] += a.y;
      acc[2] += b.x;
      acc[3] += b.y;
      acc[4] += cc.x;
      acc[5] += cc.y;
      acc[6] += d.x;
      acc[7] += d.y;
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[threadIdx.y][threadIdx.x][t] = acc[t];
  __syncthreads();
  // 256 threads -> 256 columns of the slab
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int col = blockIdx.x * 256 + tid;
  if (col < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][tid >> 3][tid & 7];
    partial[(long long)blockIdx.y * N + col] = t;
  }
}

// ------------------------------------------------------------------ dropout keep masks from a counter-based hash
// keep[i] = 1 with probability 1 - p.  Randomness = splitmix64(seed_dev[0], site, i / 4): the seed lives in DEVICE memory and
// is bumped once per training step by the host module (inside the captured CUDA graph), `site` distinguishes the call sites
// of one step, so a graph replay draws fresh masks without any host involvement.  One launch replaces torch's
// rand -> compare -> cast chain (3 launches and 6 x the bytes) at every attention-dropout site (reference transformer.py:613).
__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ keep, long long n, const long long* __restrict__ seed,
                                                           unsigned long long site, uint32_t thr) {
  pdl_wait();
  pdl_trigger();
  const unsigned long long base = drop_base(seed, site);
  const long long n16 = n >> 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint4 o;
    o.x = keep4(splitmix64(base + 4ull * i), thr);
    o.y = keep4(splitmix64(base + 4ull * i + 1), thr);
    o.z = keep4(splitmix64(base + 4ull * i + 2), thr);
    o.w = keep4(splitmix64(base + 4ull * i + 3), thr);
    reinterpret_cast<uint4*>(keep)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {      // tail: fewer than 16 bytes
    const long long e = (n16 << 4) + threadIdx.x;
    const unsigned long long r = splitmix64(base + (unsigned long long)(e >> 2));
    keep[e] = (uint8_t)((((uint32_t)(r >> (16 * (e & 3)))) & 0xFFFFu) >= thr);
  }
}

// y = keep ? x / (1 - p) : 0 on bf16 (FFN hidden dropout, reference transformer.py:644,749 `self.dropout(self.activation(...))`);
// keep bits = the stream of tdb_dropout_mask(seed, site).  8 elements (16 bytes) per thread.
__global__ void __launch_bounds__(256) dropout_bf16_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n,
                                                           const long long* __restrict__ seed, unsigned long long site, uint32_t thr,
                                                           float scale) {
  pdl_wait();
  pdl_trigger();
  const unsigned long long base = drop_base(seed, site);
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const unsigned long long r0 = splitmix64(base + 2ull * i), r1 = splitmix64(base + 2ull * i + 1);
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t out[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const unsigned long long r = t < 2 ? r0 : r1;
      const int sh = 32 * (t & 1);
      const bool k0 = (((uint32_t)(r >> sh)) & 0xFFFFu) >= thr, k1 = (((uint32_t)(r >> (sh + 16))) & 0xFFFFu) >= thr;
      const float2 f = unpack_bf16x2(in[t]);
      out[t] = pack_bf16x2(k0 ? f.x * scale : 0.f, k1 ? f.y * scale : 0.f);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(out[0], out[1], out[2], out[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long e = (n8 << 3) + threadIdx.x;
    y[e] = drop_keep(base, e, thr) ? __float2bfloat16(__bfloat162float(x[e]) * scale) : __float2bfloat16(0.f);
  }
}

}  // namespace tdb

using namespace tdb;
#define STREAM ((cudaStream_t)stream_)
#define LAUNCH_OK()                      \
  do {                                   \
    TDB_CHECK_CUDA(cudaGetLastError());  \
    tdb_count_launch(1);                 \
    return TDB_OK;                       \
  } while (0)

extern "C" int tdb_stem_im2col(const float* x, void* col, int N, int H, int W, void* stream_) {
  TDB_REQUIRE(x && col && N > 0 && H > 0 && W > 0, "tdb_stem_im2col: bad args");
  int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  size_t smem = (size_t)21 * (W + 6) * sizeof(float);
  TDB_REQUIRE(smem <= 48 * 1024, "tdb_stem_im2col: frame width %d too large", W);
  TDB_CHECK_CUDA(tdb_launch(stem_im2col_kernel, dim3((unsigned)(N * Ho)), dim3(256), smem, STREAM, x, (bf16*)col, N, H, W, Ho, Wo));
  LAUNCH_OK();
}
extern "C" int tdb_maxpool3x3s2(const void* x, void* y, int N, int H, int W, int C, void* stream_) {
  TDB_REQUIRE(x && y && C % 8 == 0, "tdb_maxpool3x3s2: bad args");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / 8);
  TDB_CHECK_CUDA(tdb_launch(maxpool3x3s2_kernel, dim3(nblocks(total, 256)), dim3(256), 0, STREAM, (const bf16*)x, (bf16*)y, N, H, W, C, Ho, Wo));
  LAUNCH_OK();
}
extern "C" int tdb_im2col3x3s2(const void* x, void* col, int N, int H, int W, int C, void* stream_) {
  TDB_REQUIRE(x && col && C % 8 == 0, "tdb_im2col3x3s2: bad args");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * 9 * (C / 8);
  TDB_CHECK_CUDA(tdb_launch(im2col3x3s2_kernel, dim3(nblocks(total, 256)), dim3(256), 0, STREAM, (const bf16*)x, (bf16*)col, N, H, W, C, Ho, Wo));
  LAUNCH_OK();
}
extern "C" int tdb_col2im3x3s2_mask(const void* dcol, const void* ymask, void* dx, int N, int H, int W, int C, void* stream_) {
  TDB_REQUIRE(dcol && dx && C % 8 == 0, "tdb_col2im3x3s2_mask: bad args");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * H * W * (C / 8);
  TDB_CHECK_CUDA(tdb_launch(col2im3x3s2_mask_kernel, dim3(nblocks(total, 256)), dim3(256), 0, STREAM, (const bf16*)dcol, (const bf16*)ymask, (bf16*)dx, N, H, W, C, Ho, Wo));
  LAUNCH_OK();
}
extern "C" int tdb_subsample2_ld(const void* x, void* y, int64_t ldy, int N, int H, int W, int C, void* stream_) {
  TDB_REQUIRE(x && y && C % 8 == 0 && ldy >= C && ldy % 8 == 0, "tdb_subsample2: bad args");
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / 8);
  TDB_CHECK_CUDA(tdb_launch(subsample2_kernel, dim3(nblocks(total, 256)), dim3(256), 0, STREAM, (const bf16*)x, (bf16*)y, (long long)ldy, N, H, W, C, Ho, Wo));
  LAUNCH_OK();
}
extern "C" int tdb_subsample2(const void* x, void* y, int N, int H, int W, int C, void* stream_) {
  return tdb_subsample2_ld(x, y, C, N, H, W, C, stream_);
}
extern "C" int tdb_upsample2_zero(const void* y, void* x, int N, int H, int W, int C, void* stream_) {
  TDB_REQUIRE(x && y && C % 8 == 0, "tdb_upsample2_zero: bad args");
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  long long total = (long long)N * H * W * (C / 8);
  TDB_CHECK_CUDA(tdb_launch(upsample2_zero_kernel, dim3(nblocks(total, 256)), dim3(256), 0, STREAM, (const bf16*)y, (bf16*)x, N, H, W, C, Ho, Wo));
  LAUNCH_OK();
}
extern "C" int tdb_prep_weight(const float* w, void* out, void* out_scaled, const float* rowscale, int Cout, int Cin,
                               int taps, int Kpad, void* stream_) {
  TDB_REQUIRE(w && out && Cout > 0 && Cin > 0 && (taps == 1 || taps == 9 || taps == 49) && Kpad >= taps * Cin, "tdb_prep_weight: bad args");
  TDB_REQUIRE(!out_scaled || rowscale, "tdb_prep_weight: scaled copy needs rowscale");
  long long total = (long long)Cout * Kpad;
  TDB_CHECK_CUDA(tdb_launch(prep_weight_kernel, dim3(nblocks(total, 256)), dim3(256), 0, STREAM, w, (bf16*)out, (bf16*)out_scaled, rowscale, Cout, Cin, taps, Kpad));
  LAUNCH_OK();
}
extern "C" int tdb_cast_add_bf16(const float* x, const float* add, void* y, int64_t n, void* stream_) {
  TDB_REQUIRE(x && y && n % 4 == 0, "tdb_cast_add_bf16: n must be a multiple of 4");
  TDB_CHECK_CUDA(tdb_launch(cast_add_bf16_kernel, dim3(nblocks(n / 4, 256)), dim3(256), 0, STREAM, x, add, (bf16*)y, n / 4));
  LAUNCH_OK();
}
extern "C" int tdb_split_bf16(const float* x, void* y, int64_t rows, int K, void* stream_) {
  TDB_REQUIRE(x && y && rows > 0 && K > 0 && K % 4 == 0, "tdb_split_bf16: K must be a multiple of 4");
  TDB_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "tdb_split_bf16: buffers must be 16-byte aligned");
  const long long n4 = rows * (K / 4);
  TDB_CHECK_CUDA(tdb_launch(split_bf16_kernel, dim3(nblocks(n4, 256)), dim3(256), 0, STREAM, x, (bf16*)y, n4, K / 4));
  LAUNCH_OK();
}
extern "C" int tdb_layernorm_fwd(const float* x, const float* r, const float* gamma, const float* beta, const float* pos,
                                 float* y, void* y_bf, void* ypos_bf, float* mean, float* rstd, int rows, int D, float eps,
                                 const int64_t* drop_seed, int64_t drop_site, float drop_p, void* stream_) {
  TDB_REQUIRE(x && gamma && beta && y && rows > 0 && (D == 256 || D == 768), "tdb_layernorm_fwd: D must be 256 or 768 (got %d)", D);
  TDB_REQUIRE(!ypos_bf || pos, "tdb_layernorm_fwd: ypos needs pos");
  TDB_REQUIRE(!drop_seed || (r && drop_p > 0.f && drop_p < 1.f), "tdb_layernorm_fwd: residual dropout needs r and 0 < p < 1");
  const uint32_t thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  const float dscale = drop_seed ? 1.f / (1.f - drop_p) : 1.f;
  if (D == 256)
    TDB_CHECK_CUDA(tdb_launch(layernorm_fwd_kernel<256>, dim3(nblocks((long long)rows * 32, 256)), dim3(256), 0, STREAM, x, r, gamma, beta, pos, y, (bf16*)y_bf,
                              (bf16*)ypos_bf, mean, rstd, rows, eps, (const long long*)drop_seed, (unsigned long long)drop_site, thr, dscale));
  else       // text encoder (RoBERTa, d = 768)
    TDB_CHECK_CUDA(tdb_launch(layernorm_fwd_kernel<768>, dim3(nblocks((long long)rows * 32, 256)), dim3(256), 0, STREAM, x, r, gamma, beta, pos, y, (bf16*)y_bf,
                              (bf16*)ypos_bf, mean, rstd, rows, eps, (const long long*)drop_seed, (unsigned long long)drop_site, thr, dscale));
  LAUNCH_OK();
}
extern "C" int tdb_layernorm_bwd_blocks(int rows) {
  int b = (rows + 7) / 8;
  return b > 296 ? 296 : b;
}
extern "C" int tdb_layernorm_bwd(const float* dy, const void* dy2_bf, const void* dy3_bf, const float* x, const float* r,
                                 const float* gamma, const float* mean, const float* rstd, float* dz, void* dz_bf,
                                 float* dgamma, float* dbeta, float* partial, int rows, int D, int accumulate,
                                 const int64_t* drop_seed, int64_t drop_site, float drop_p, float* dr, void* dr_bf, float* dbias,
                                 void* stream_) {
  TDB_REQUIRE((dy || dy2_bf || dy3_bf) && x && gamma && mean && rstd && dz && partial && rows > 0 && (D == 256 || D == 768), "tdb_layernorm_bwd: bad args");
  TDB_REQUIRE(!drop_seed || (r && drop_p > 0.f && drop_p < 1.f && (dr || dr_bf)), "tdb_layernorm_bwd: residual dropout needs r, 0 < p < 1 and a dr output");
  const uint32_t thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  const float dscale = drop_seed ? 1.f / (1.f - drop_p) : 1.f;
  int blocks = tdb_layernorm_bwd_blocks(rows);
  if (D == 256)
    TDB_CHECK_CUDA(tdb_launch(layernorm_bwd_kernel<256, 8>, dim3(blocks), dim3(256), 0, STREAM, dy, (const bf16*)dy2_bf, (const bf16*)dy3_bf, x, r, gamma, mean, rstd, dz, (bf16*)dz_bf, partial, rows,
                              (const long long*)drop_seed, (unsigned long long)drop_site, thr, dscale, dr, (bf16*)dr_bf));
  else
    TDB_CHECK_CUDA(tdb_launch(layernorm_bwd_kernel<768, 4>, dim3(blocks), dim3(128), 0, STREAM, dy, (const bf16*)dy2_bf, (const bf16*)dy3_bf, x, r, gamma, mean, rstd, dz, (bf16*)dz_bf, partial, rows,
                              (const long long*)drop_seed, (unsigned long long)drop_site, thr, dscale, dr, (bf16*)dr_bf));
  TDB_CHECK_CUDA(cudaGetLastError());
  if (dgamma && dbeta == dgamma + D) {   // [dgamma | dbeta (| dbias)] contiguous: the partial rows reduce in ONE launch
    const int cols = (dbias == dgamma + 2 * D) ? 3 * D : 2 * D;
    TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((cols + 31) / 32), dim3(32, 8), 0, STREAM, partial, blocks, 3 * D, cols, dgamma, accumulate));
    if (dbias && cols == 2 * D)
      TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((D + 31) / 32), dim3(32, 8), 0, STREAM, partial + 2 * D, blocks, 3 * D, D, dbias, accumulate));
    TDB_CHECK_CUDA(cudaGetLastError());
    tdb_count_launch(2);
    return TDB_OK;
  }
  if (dgamma) TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((D + 31) / 32), dim3(32, 8), 0, STREAM, partial, blocks, 3 * D, D, dgamma, accumulate));
  if (dbeta) TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((D + 31) / 32), dim3(32, 8), 0, STREAM, partial + D, blocks, 3 * D, D, dbeta, accumulate));
  if (dbias) TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((D + 31) / 32), dim3(32, 8), 0, STREAM, partial + 2 * D, blocks, 3 * D, D, dbias, accumulate));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(3);
  return TDB_OK;
}
extern "C" int tdb_colsum_bf16(const void* x, int64_t ld, int rows, int N, float* partial, int nparts, float* out,
                               int accumulate, void* stream_) {
  TDB_REQUIRE(x && partial && out && rows > 0 && N > 0 && nparts > 0, "tdb_colsum_bf16: bad args");
  TDB_REQUIRE(N % 2 == 0 && ld % 2 == 0, "tdb_colsum_bf16: N and ld must be even");
  int rpb = (rows + nparts - 1) / nparts;
  if (N % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    // fewer, fatter parts (>= 64 rows each, enough CTAs for two waves): a shorter second stage; nparts only bounds the workspace
    const int slabs = (N + 255) / 256;
    int want = (2 * 148 + slabs - 1) / slabs;
    int np = want < nparts ? want : nparts;
    const int by_rows = (rows + 63) / 64;
    if (np > by_rows) np = by_rows;
    rpb = ((rows + np - 1) / np + 7) / 8 * 8;
    np = (rows + rpb - 1) / rpb;
    TDB_CHECK_CUDA(tdb_launch(colsum_bf16_v8_kernel, dim3(slabs, np), dim3(32, 8), 0, STREAM, (const bf16*)x, ld, rows, N, partial, rpb));
    TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((N + 31) / 32), dim3(32, 8), 0, STREAM, partial, np, N, N, out, accumulate));
    TDB_CHECK_CUDA(cudaGetLastError());
    tdb_count_launch(2);
    return TDB_OK;
  }
  dim3 grid((N / 2 + 127) / 128, nparts);
  TDB_CHECK_CUDA(tdb_launch(colsum_bf16_kernel, dim3(grid), dim3(128), 0, STREAM, (const bf16*)x, ld, rows, N, partial, rpb));
  TDB_CHECK_CUDA(tdb_launch(colsum_partials_kernel, dim3((N + 31) / 32), dim3(32, 8), 0, STREAM, partial, nparts, N, N, out, accumulate));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(2);
  return TDB_OK;
}

extern "C" int tdb_dropout_mask(uint8_t* keep, int64_t n, const int64_t* seed, int64_t site, float p, void* stream_) {
  TDB_REQUIRE(keep && seed && n > 0 && p >= 0.f && p < 1.f, "tdb_dropout_mask: bad args");
  TDB_REQUIRE(((uintptr_t)keep & 15) == 0, "tdb_dropout_mask: keep must be 16-byte aligned");
  const uint32_t thr = (uint32_t)(p * 65536.0f + 0.5f);
  long long want = ((n >> 4) + 255) / 256;
  int nb = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  TDB_CHECK_CUDA(tdb_launch(dropout_mask_kernel, dim3(nb), dim3(256), 0, STREAM, keep, (long long)n, (const long long*)seed,
                            (unsigned long long)site, thr));
  LAUNCH_OK();
}

extern "C" int tdb_dropout_bf16(const void* x, void* y, int64_t n, const int64_t* seed, int64_t site, float p, void* stream_) {
  TDB_REQUIRE(x && y && seed && n > 0 && p >= 0.f && p < 1.f, "tdb_dropout_bf16: bad args");
  TDB_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "tdb_dropout_bf16: buffers must be 16-byte aligned");
  const uint32_t thr = (uint32_t)(p * 65536.0f + 0.5f);
  long long want = ((n >> 3) + 255) / 256;
  int nb = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  TDB_CHECK_CUDA(tdb_launch(dropout_bf16_kernel, dim3(nb), dim3(256), 0, STREAM, (const bf16*)x, (bf16*)y, (long long)n,
                            (const long long*)seed, (unsigned long long)site, thr, 1.f / (1.f - p)));
  LAUNCH_OK();
}
