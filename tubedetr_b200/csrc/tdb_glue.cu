// tubedetr_b200 -- fused data-movement kernels around the video-text encoder (d_model = 256, token rows are frame-major).
//
//   tdb_pos_sine          PositionEmbeddingSine in ONE kernel (reference models/position_encoding.py:71-94)
//   tdb_enc_assemble_*    encoder input: [image tokens | per-clip repeated text tokens] + position rows, fp32 residual stream and the
//                         bf16 operands of the first in-projection (reference models/transformer.py:269-331: text repeat, concat)
//   tdb_fast_mix_*        z = bf16(enc[clip(t)] + fast_encoder(fast_src))            (transformer.py:373-391, input of fast_residual)
//   tdb_aggregate_*       temporal replication + fast-branch aggregation + the decoder's bf16 memory operands in one pass:
//                         mem[t] = enc[clip(t)] (+ fast_residual update on the image rows), memb = bf16(mem),
//                         mempb = bf16(mem + pos[clip(t)]), mem_pos[t] = pos[clip(t)]    (transformer.py:393-446)
// clip(b, t) = b * n_clips + t / k: frame t of video b reads clip t / k (temporal stride k).  Backward kernels sum the <= k frames of a
// clip in a fixed order (deterministic).  Every thread handles 4 consecutive channels (16 / 8-byte accesses).
#include <math.h>

#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);

namespace tdb {

constexpr int GD = 256;          // d_model
constexpr int GQ = GD / 4;       // float4 groups per row

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldb4(const bf16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stb4(bf16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ------------------------------------------------------------------ sine position embedding (normalize = True, 128 features per axis)
__global__ void __launch_bounds__(256) pos_sine_kernel(const uint8_t* __restrict__ mask, float* __restrict__ out, int N, int h, int w) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (n, pixel, channel pair)
  const long long total = (long long)N * h * w * 128;
  if (idx >= total) return;
  const int pr = (int)(idx % 128);            // pair index: channels 2 pr, 2 pr + 1 of the 256
  const long long pix = idx / 128;
  const int x = (int)(pix % w), y = (int)((pix / w) % h);
  const int n = (int)(pix / ((long long)w * h));
  const uint8_t* m = mask + (long long)n * h * w;
  const bool ypart = pr < 64;
  float cum = 0.f, tot = 0.f;
  if (ypart) {
    for (int r = 0; r < h; ++r) {
      const float v = m[r * w + x] ? 0.f : 1.f;
      tot += v;
      if (r <= y) cum += v;
    }
  } else {
    for (int c = 0; c < w; ++c) {
      const float v = m[y * w + c] ? 0.f : 1.f;
      tot += v;
      if (c <= x) cum += v;
    }
  }
  const float e = cum / (tot + 1e-6f) * 6.283185307179586f;
  const int i = (pr & 63) * 2;                                   // feature index of the pair inside its axis: (i, i + 1) share dim_t
  const float dim_t = powf(10000.f, (float)i / 128.f);
  const float a = e / dim_t;
  reinterpret_cast<float2*>(out)[idx] = make_float2(sinf(a), cosf(a));
}

// ------------------------------------------------------------------ encoder input assembly
__global__ void __launch_bounds__(256) enc_assemble_fwd_kernel(const float* __restrict__ src, const float* __restrict__ txt,
                                                               const float* __restrict__ pos, float* __restrict__ x32, bf16* __restrict__ xb,
                                                               bf16* __restrict__ xpb, float* __restrict__ pe, int n, int HW, int L, int n_clips) {
  pdl_wait();
  pdl_trigger();
  const int S = HW + L;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * S * GQ) return;
  const int q = (int)(idx % GQ);
  const long long row = idx / GQ;
  const int s = (int)(row % S), f = (int)(row / S);
  float4 v, p;
  if (s < HW) {
    v = ld4(src + ((long long)f * HW + s) * GD + q * 4);
    p = ld4(pos + ((long long)f * HW + s) * GD + q * 4);
  } else {
    v = ld4(txt + ((long long)(f / n_clips) * L + (s - HW)) * GD + q * 4);
    p = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  reinterpret_cast<float4*>(x32)[idx] = v;
  reinterpret_cast<float4*>(pe)[idx] = p;
  stb4(xb + idx * 4, v);
  stb4(xpb + idx * 4, add4(v, p));
}

// dsrc[f][s] = g32 + gb + gpb (image rows);  dtxt[b][l] = sum over the clips of video b of the same sum at the text rows
__global__ void __launch_bounds__(256) enc_assemble_bwd_kernel(const float* __restrict__ g32, const bf16* __restrict__ gb, const bf16* __restrict__ gpb,
                                                               float* __restrict__ dsrc, float* __restrict__ dtxt, int n, int HW, int L,
                                                               int n_clips) {
  pdl_wait();
  pdl_trigger();
  const int S = HW + L, B = n / n_clips;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_img = (long long)n * HW * GQ, n_txt = (long long)B * L * GQ;
  if (idx >= n_img + n_txt) return;
  auto grad_at = [&](long long row, int q) {
    const long long e = (row * GQ + q) * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g32) g = ld4(g32 + e);
    if (gb) g = add4(g, ldb4(gb + e));
    if (gpb) g = add4(g, ldb4(gpb + e));
    return g;
  };
  if (idx < n_img) {
    const int q = (int)(idx % GQ);
    const long long r = idx / GQ;
    const int s = (int)(r % HW), f = (int)(r / HW);
    reinterpret_cast<float4*>(dsrc)[idx] = grad_at((long long)f * S + s, q);
  } else {
    const long long j = idx - n_img;
    const int q = (int)(j % GQ);
    const long long r = j / GQ;
    const int l = (int)(r % L), b = (int)(r / L);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < n_clips; ++c) acc = add4(acc, grad_at((long long)(b * n_clips + c) * S + HW + l, q));
    reinterpret_cast<float4*>(dtxt)[j] = acc;
  }
}

// ------------------------------------------------------------------ fast branch input: z[bt][hw] = bf16(enc[clip(bt)][hw] + fm[bt][hw])
__global__ void __launch_bounds__(256) fast_mix_fwd_kernel(const float* __restrict__ enc, const bf16* __restrict__ fm, bf16* __restrict__ z, int BT,
                                                           int T, int k, int n_clips, int HW, int S) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)BT * HW * GQ) return;
  const int q = (int)(idx % GQ);
  const long long r = idx / GQ;
  const int s = (int)(r % HW), bt = (int)(r / HW);
  const int c = (bt / T) * n_clips + (bt % T) / k;
  stb4(z + idx * 4, add4(ld4(enc + ((long long)c * S + s) * GD + q * 4), ldb4(fm + idx * 4)));
}

// sum over the frames of a clip of a bf16 [BT][rows_per_frame][256] gradient -> fp32 [n][S][256] rows [0, rows_per_frame) (other rows 0)
__device__ __forceinline__ void clip_range(int c, int T, int k, int n_clips, int& bt0, int& cnt) {
  const int b = c / n_clips, j = c - b * n_clips;
  bt0 = b * T + j * k;
  cnt = min(k, T - j * k);
}
__global__ void __launch_bounds__(256) fast_mix_bwd_kernel(const bf16* __restrict__ dz, float* __restrict__ denc, int n, int T, int k, int n_clips,
                                                           int HW, int S) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * S * GQ) return;
  const int q = (int)(idx % GQ);
  const long long r = idx / GQ;
  const int s = (int)(r % S), c = (int)(r / S);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s < HW) {
    int bt0, cnt;
    clip_range(c, T, k, n_clips, bt0, cnt);
    for (int t = 0; t < cnt; ++t) acc = add4(acc, ldb4(dz + (((long long)(bt0 + t) * HW + s) * GQ + q) * 4));
  }
  reinterpret_cast<float4*>(denc)[idx] = acc;
}

// ------------------------------------------------------------------ replication + aggregation + decoder operands
__global__ void __launch_bounds__(256) aggregate_fwd_kernel(const float* __restrict__ enc, const float* __restrict__ pe, const float* __restrict__ upd,
                                                            float* __restrict__ mem, float* __restrict__ mem_pos, bf16* __restrict__ memb,
                                                            bf16* __restrict__ mempb, int BT, int T, int k, int n_clips, int HW, int S) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)BT * S * GQ) return;
  const int q = (int)(idx % GQ);
  const long long r = idx / GQ;
  const int s = (int)(r % S), bt = (int)(r / S);
  const int c = (bt / T) * n_clips + (bt % T) / k;
  const long long src = ((long long)c * S + s) * GD + q * 4;
  float4 v = ld4(enc + src);
  if (upd && s < HW) v = add4(v, ld4(upd + (((long long)bt * HW + s) * GQ + q) * 4));
  const float4 p = ld4(pe + src);
  reinterpret_cast<float4*>(mem)[idx] = v;
  reinterpret_cast<float4*>(mem_pos)[idx] = p;
  stb4(memb + idx * 4, v);
  stb4(mempb + idx * 4, add4(v, p));
}

// g = gmem (fp32) + gmemb + gmempb (bf16), any may be null.  dupd[bt][hw] = g (fp32 + bf16 copy); denc[c][s] = sum over the clip's frames
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(const float* __restrict__ gmem, const bf16* __restrict__ gmemb, const bf16* __restrict__ gmempb,
                                                            float* __restrict__ denc, float* __restrict__ dupd, bf16* __restrict__ dupd_b, int n, int T,
                                                            int k, int n_clips, int HW, int S) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * S * GQ) return;
  const int q = (int)(idx % GQ);
  const long long r = idx / GQ;
  const int s = (int)(r % S), c = (int)(r / S);
  int bt0, cnt;
  clip_range(c, T, k, n_clips, bt0, cnt);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < cnt; ++t) {
    const int bt = bt0 + t;
    const long long e = (((long long)bt * S + s) * GQ + q) * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gmem) g = ld4(gmem + e);
    if (gmemb) g = add4(g, ldb4(gmemb + e));
    if (gmempb) g = add4(g, ldb4(gmempb + e));
    if (dupd && s < HW) {
      const long long u = (((long long)bt * HW + s) * GQ + q) * 4;
      *reinterpret_cast<float4*>(dupd + u) = g;
      stb4(dupd_b + u, g);
    }
    acc = add4(acc, g);
  }
  reinterpret_cast<float4*>(denc)[idx] = acc;
}

// ------------------------------------------------------------------ prediction heads: the final projection of an MLP head (256 -> J <= 8)
// Reference models/tubedetr.py:23-42, 226-252: bbox_embed = MLP(256, 256, 4, 3) + sigmoid, sted_embed = MLP(256, 256, 2, 2, dropout 0.5).
// The 256 -> 256 layers run on the tcgen05 GEMM (ReLU epilogue); this kernel is the last layer fused with its activation
// (sigmoid) and, for the start / end head in train(), the dropout on the logits (hash stream).  One warp per row.
__global__ void __launch_bounds__(256) head_out_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                                           float* __restrict__ y, int R, int J, int act, const long long* __restrict__ drop_seed,
                                                           unsigned long long drop_site, uint32_t drop_thr, float drop_scale) {
  pdl_wait();
  pdl_trigger();
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + (long long)row * GD) + lane);
  const float2 x0 = unpack_bf16x2(raw.x), x1 = unpack_bf16x2(raw.y), x2 = unpack_bf16x2(raw.z), x3 = unpack_bf16x2(raw.w);
  const unsigned long long dbase = drop_seed ? drop_base(drop_seed, drop_site) : 0ull;
  for (int j = 0; j < J; ++j) {
    const float4 w0 = ld4(W + (long long)j * GD + lane * 8), w1 = ld4(W + (long long)j * GD + lane * 8 + 4);
    float d = x0.x * w0.x + x0.y * w0.y + x1.x * w0.z + x1.y * w0.w + x2.x * w1.x + x2.y * w1.y + x3.x * w1.z + x3.y * w1.w;
    d = warp_sum(d);
    if (lane == 0) {
      float v = d + b[j];
      if (act == 1) v = 1.f / (1.f + __expf(-v));
      if (drop_seed) v = drop_keep(dbase, (long long)row * J + j, drop_thr) ? v * drop_scale : 0.f;
      y[(long long)row * J + j] = v;
    }
  }
}

// dpre = dy * act'(y) * dropout'  -> dpre [R][J] (kept for the weight-gradient kernel), dx [R][256] bf16 (optionally masked by x > 0 and
// scaled: the ReLU / hidden-dropout backward of the producing layer)
__global__ void __launch_bounds__(256) head_out_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const bf16* __restrict__ x,
                                                           const float* __restrict__ W, float* __restrict__ dpre, bf16* __restrict__ dx, int R, int J,
                                                           int act, int mask_dx, float dx_scale, const long long* __restrict__ drop_seed,
                                                           unsigned long long drop_site, uint32_t drop_thr, float drop_scale) {
  pdl_wait();
  pdl_trigger();
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const unsigned long long dbase = drop_seed ? drop_base(drop_seed, drop_site) : 0ull;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  for (int j = 0; j < J; ++j) {
    float g = dy[(long long)row * J + j];
    float yv = y[(long long)row * J + j];
    if (drop_seed) {
      const bool kept = drop_keep(dbase, (long long)row * J + j, drop_thr);
      g = kept ? g * drop_scale : 0.f;
      yv = kept ? yv / drop_scale : 0.f;       // undo the dropout scaling to get the activation value (only used by act == 1)
    }
    if (act == 1) g *= yv * (1.f - yv);
    if (lane == 0) dpre[(long long)row * J + j] = g;
    const float4 w0 = ld4(W + (long long)j * GD + lane * 8), w1 = ld4(W + (long long)j * GD + lane * 8 + 4);
    acc[0] += g * w0.x; acc[1] += g * w0.y; acc[2] += g * w0.z; acc[3] += g * w0.w;
    acc[4] += g * w1.x; acc[5] += g * w1.y; acc[6] += g * w1.z; acc[7] += g * w1.w;
  }
  if (mask_dx) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + (long long)row * GD) + lane);
    const float2 x0 = unpack_bf16x2(raw.x), x1 = unpack_bf16x2(raw.y), x2 = unpack_bf16x2(raw.z), x3 = unpack_bf16x2(raw.w);
    const float xs[8] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y};
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = xs[t] > 0.f ? acc[t] * dx_scale : 0.f;
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(dx + (long long)row * GD)[lane] = o;
}

// dW[j][c] = sum_r dpre[r][j] x[r][c], db[j] = sum_r dpre[r][j]: one block per output row j; the rows are split over 4 groups of 256
// threads (thread = column c), 8 rows of loads in flight per thread, partial sums combined in a fixed order (deterministic).  The
// kernel sits at the head of the backward's serial chain: the one-thread-per-column loop over all 600 rows it replaces took 77 us.
__global__ void __launch_bounds__(1024) head_out_wgrad_kernel(const float* __restrict__ dpre, const bf16* __restrict__ x, float* __restrict__ dW,
                                                              float* __restrict__ db, int R, int J) {
  __shared__ float part[3][GD];
  __shared__ float bpart[4];
  pdl_wait();
  pdl_trigger();
  const int j = blockIdx.x, c = threadIdx.x & (GD - 1), grp = threadIdx.x >> 8;
  const int r0 = (int)((long long)R * grp / 4), r1 = (int)((long long)R * (grp + 1) / 4);
  float acc = 0.f, bsum = 0.f;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    float g[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] = __ldg(dpre + (long long)(r + i) * J + j);
      v[i] = __bfloat162float(x[(long long)(r + i) * GD + c]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc += g[i] * v[i];
      bsum += g[i];
    }
  }
  for (; r < r1; ++r) {
    const float g = __ldg(dpre + (long long)r * J + j);
    acc += g * __bfloat162float(x[(long long)r * GD + c]);
    bsum += g;
  }
  if (grp) {
    part[grp - 1][c] = acc;
    if (c == 0) bpart[grp] = bsum;
  }
  __syncthreads();
  if (!grp) {
    acc += part[0][c];
    acc += part[1][c];
    acc += part[2][c];
    dW[(long long)j * GD + c] = acc;
    if (c == 0) db[j] = ((bsum + bpart[1]) + bpart[2]) + bpart[3];
  }
}

}  // namespace tdb

using namespace tdb;

#define GSTREAM ((cudaStream_t)stream_)
static inline unsigned gblocks(long long n) { return (unsigned)((n + 255) / 256); }
#define GLAUNCH_OK()                     \
  TDB_CHECK_CUDA(cudaGetLastError());    \
  tdb_count_launch(1);                   \
  return TDB_OK

extern "C" int tdb_pos_sine(const uint8_t* mask, float* out, int N, int h, int w, void* stream_) {
  TDB_REQUIRE(mask && out && N > 0 && h > 0 && w > 0, "tdb_pos_sine: bad args");
  TDB_CHECK_CUDA(tdb_launch(pos_sine_kernel, dim3(gblocks((long long)N * h * w * 128)), dim3(256), 0, GSTREAM, mask, out, N, h, w));
  GLAUNCH_OK();
}

extern "C" int tdb_enc_assemble_fwd(const float* src, const float* txt, const float* pos, float* x32, void* xb, void* xpb, float* pe, int n, int HW,
                                    int L, int n_clips, void* stream_) {
  TDB_REQUIRE(src && txt && pos && x32 && xb && xpb && pe && n > 0 && HW > 0 && L >= 0 && n_clips > 0 && n % n_clips == 0, "tdb_enc_assemble_fwd: bad args");
  TDB_CHECK_CUDA(tdb_launch(enc_assemble_fwd_kernel, dim3(gblocks((long long)n * (HW + L) * GQ)), dim3(256), 0, GSTREAM, src, txt, pos, x32, (bf16*)xb,
                            (bf16*)xpb, pe, n, HW, L, n_clips));
  GLAUNCH_OK();
}

extern "C" int tdb_enc_assemble_bwd(const float* g32, const void* gb, const void* gpb, float* dsrc, float* dtxt, int n, int HW, int L, int n_clips,
                                    void* stream_) {
  TDB_REQUIRE(dsrc && dtxt && n > 0 && n_clips > 0 && n % n_clips == 0, "tdb_enc_assemble_bwd: bad args");
  const long long total = (long long)n * HW * GQ + (long long)(n / n_clips) * L * GQ;
  TDB_CHECK_CUDA(tdb_launch(enc_assemble_bwd_kernel, dim3(gblocks(total)), dim3(256), 0, GSTREAM, g32, (const bf16*)gb, (const bf16*)gpb, dsrc, dtxt, n, HW,
                            L, n_clips));
  GLAUNCH_OK();
}

extern "C" int tdb_fast_mix_fwd(const float* enc, const void* fm, void* z, int B, int T, int k, int HW, int S, void* stream_) {
  TDB_REQUIRE(enc && fm && z && B > 0 && T > 0 && k > 0 && HW > 0 && S >= HW, "tdb_fast_mix_fwd: bad args");
  const int n_clips = (T + k - 1) / k;
  TDB_CHECK_CUDA(tdb_launch(fast_mix_fwd_kernel, dim3(gblocks((long long)B * T * HW * GQ)), dim3(256), 0, GSTREAM, enc, (const bf16*)fm, (bf16*)z, B * T, T,
                            k, n_clips, HW, S));
  GLAUNCH_OK();
}

extern "C" int tdb_fast_mix_bwd(const void* dz, float* denc, int B, int T, int k, int HW, int S, void* stream_) {
  TDB_REQUIRE(dz && denc && B > 0 && T > 0 && k > 0, "tdb_fast_mix_bwd: bad args");
  const int n_clips = (T + k - 1) / k;
  TDB_CHECK_CUDA(tdb_launch(fast_mix_bwd_kernel, dim3(gblocks((long long)B * n_clips * S * GQ)), dim3(256), 0, GSTREAM, (const bf16*)dz, denc, B * n_clips,
                            T, k, n_clips, HW, S));
  GLAUNCH_OK();
}

extern "C" int tdb_aggregate_fwd(const float* enc, const float* pe, const float* upd, float* mem, float* mem_pos, void* memb, void* mempb, int B, int T,
                                 int k, int HW, int S, void* stream_) {
  TDB_REQUIRE(enc && pe && mem && mem_pos && memb && mempb && B > 0 && T > 0 && k > 0, "tdb_aggregate_fwd: bad args");
  const int n_clips = (T + k - 1) / k;
  TDB_CHECK_CUDA(tdb_launch(aggregate_fwd_kernel, dim3(gblocks((long long)B * T * S * GQ)), dim3(256), 0, GSTREAM, enc, pe, upd, mem, mem_pos, (bf16*)memb,
                            (bf16*)mempb, B * T, T, k, n_clips, HW, S));
  GLAUNCH_OK();
}

extern "C" int tdb_aggregate_bwd(const float* gmem, const void* gmemb, const void* gmempb, float* denc, float* dupd, void* dupd_b, int B, int T, int k,
                                 int HW, int S, void* stream_) {
  TDB_REQUIRE(denc && B > 0 && T > 0 && k > 0 && (!dupd || dupd_b), "tdb_aggregate_bwd: bad args");
  const int n_clips = (T + k - 1) / k;
  TDB_CHECK_CUDA(tdb_launch(aggregate_bwd_kernel, dim3(gblocks((long long)B * n_clips * S * GQ)), dim3(256), 0, GSTREAM, gmem, (const bf16*)gmemb,
                            (const bf16*)gmempb, denc, dupd, (bf16*)dupd_b, B * n_clips, T, k, n_clips, HW, S));
  GLAUNCH_OK();
}

extern "C" int tdb_head_out_fwd(const void* x, const float* W, const float* b, float* y, int R, int J, int act, const int64_t* drop_seed,
                                int64_t drop_site, float drop_p, void* stream_) {
  TDB_REQUIRE(x && W && b && y && R > 0 && J >= 1 && J <= 8 && (act == 0 || act == 1), "tdb_head_out_fwd: bad args");
  TDB_REQUIRE(!drop_seed || (drop_p > 0.f && drop_p < 1.f), "tdb_head_out_fwd: dropout needs 0 < p < 1");
  const uint32_t thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  TDB_CHECK_CUDA(tdb_launch(head_out_fwd_kernel, dim3(gblocks((long long)R * 32)), dim3(256), 0, GSTREAM, (const bf16*)x, W, b, y, R, J, act,
                            (const long long*)drop_seed, (unsigned long long)drop_site, thr, drop_seed ? 1.f / (1.f - drop_p) : 1.f));
  GLAUNCH_OK();
}

extern "C" int tdb_head_out_bwd(const float* dy, const float* y, const void* x, const float* W, float* dpre, void* dx, float* dW, float* db, int R,
                                int J, int act, int mask_dx, float dx_scale, const int64_t* drop_seed, int64_t drop_site, float drop_p,
                                void* stream_) {
  TDB_REQUIRE(dy && y && x && W && dpre && dx && dW && db && R > 0 && J >= 1 && J <= 8, "tdb_head_out_bwd: bad args");
  const uint32_t thr = drop_seed ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  TDB_CHECK_CUDA(tdb_launch(head_out_bwd_kernel, dim3(gblocks((long long)R * 32)), dim3(256), 0, GSTREAM, dy, y, (const bf16*)x, W, dpre, (bf16*)dx, R, J,
                            act, mask_dx, dx_scale, (const long long*)drop_seed, (unsigned long long)drop_site, thr,
                            drop_seed ? 1.f / (1.f - drop_p) : 1.f));
  TDB_CHECK_CUDA(cudaGetLastError());
  TDB_CHECK_CUDA(tdb_launch(head_out_wgrad_kernel, dim3(J), dim3(1024), 0, GSTREAM, (const float*)dpre, (const bf16*)x, dW, db, R, J));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(2);
  return TDB_OK;
}

// ------------------------------------------------------------------ input pipeline: decoded frames -> model input
// Reference datasets/vidstg.py:104-116 + datasets/video_transforms.py (resize -> ToTensor (/255) -> Normalize(mean, std)) + util/misc.py:
// 142-172 (pad-and-pack into a NestedTensor) run on the CPU per frame.  Here ONE kernel takes the decoder's packed rgb24 frames
// [T][H0][W0][3] uint8 and writes the resized (bilinear, half-pixel centres = cv2.INTER_LINEAR / F.interpolate(align_corners=False)),
// scaled and normalised fp32 frames straight into their slot [T][3][Hp][Wp] of the padded batch tensor, zero padding and pad mask included.
namespace tdb {
__global__ void __launch_bounds__(256) frames_preprocess_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, uint8_t* __restrict__ mask,
                                                                int T, int H0, int W0, int H, int W, int Hp, int Wp, float m0, float m1, float m2,
                                                                float is0, float is1, float is2) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * Hp * Wp) return;
  const int x = (int)(idx % Wp), y = (int)((idx / Wp) % Hp), t = (int)(idx / ((long long)Wp * Hp));
  float o[3] = {0.f, 0.f, 0.f};
  const bool inside = y < H && x < W;
  if (inside) {
    const float sy = ((float)y + 0.5f) * ((float)H0 / (float)H) - 0.5f, sx = ((float)x + 0.5f) * ((float)W0 / (float)W) - 0.5f;
    const float fy = fmaxf(sy, 0.f), fx = fmaxf(sx, 0.f);
    const int y0 = min((int)fy, H0 - 1), x0 = min((int)fx, W0 - 1);
    const int y1 = min(y0 + 1, H0 - 1), x1 = min(x0 + 1, W0 - 1);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    const uint8_t* f = src + (long long)t * H0 * W0 * 3;
    const float mean[3] = {m0, m1, m2}, istd[3] = {is0, is1, is2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = f[((long long)y0 * W0 + x0) * 3 + c], b = f[((long long)y0 * W0 + x1) * 3 + c];
      const float cc = f[((long long)y1 * W0 + x0) * 3 + c], d = f[((long long)y1 * W0 + x1) * 3 + c];
      const float v = (a * (1.f - wx) + b * wx) * (1.f - wy) + (cc * (1.f - wx) + d * wx) * wy;
      o[c] = (v * (1.f / 255.f) - mean[c]) * istd[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[(((long long)t * 3 + c) * Hp + y) * Wp + x] = o[c];
  if (mask) mask[idx] = inside ? 0 : 1;
}
}  // namespace tdb

extern "C" int tdb_frames_preprocess(const uint8_t* src, float* dst, uint8_t* mask, int T, int H0, int W0, int H, int W, int Hp, int Wp,
                                     const float* mean3, const float* std3, void* stream_) {
  TDB_REQUIRE(src && dst && T > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W && mean3 && std3, "tdb_frames_preprocess: bad args");
  TDB_CHECK_CUDA(tdb_launch(tdb::frames_preprocess_kernel, dim3(gblocks((long long)T * Hp * Wp)), dim3(256), 0, GSTREAM, src, dst, mask, T, H0, W0, H, W,
                            Hp, Wp, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1], 1.f / std3[2]));
  GLAUNCH_OK();
}

// ------------------------------------------------------------------ measurement aid: occupy `ctas` SMs for `cycles` SM clocks
// (tools/clc_hog.py: how a GEMM behaves while another stream's kernel -- an NCCL all-reduce, the text encoder -- holds some SMs)
namespace tdb {
__global__ void __launch_bounds__(256) debug_spin_kernel(long long cycles, int* sink) {
  extern __shared__ uint8_t spin_smem[];
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
  if (sink && cycles < 0) sink[0] = spin_smem[threadIdx.x];
}
}  // namespace tdb

extern "C" int tdb_debug_spin(int ctas, int smem_bytes, long long cycles, void* stream_) {
  TDB_REQUIRE(ctas > 0 && smem_bytes >= 0 && smem_bytes <= 200 * 1024 && cycles >= 0 && cycles < 4000000000ll, "tdb_debug_spin: bad args");
  TDB_CHECK_CUDA(cudaFuncSetAttribute(tdb::debug_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tdb::debug_spin_kernel<<<ctas, 256, smem_bytes, (cudaStream_t)stream_>>>(cycles, nullptr);
  TDB_CHECK_CUDA(cudaGetLastError());
  return TDB_OK;
}
