// tubedetr_b200 -- optimizer-side step of the training loop as three HBM-streaming kernels.
//
// Reference: engine.py:147-161   clip_grad_norm_(model.parameters(), max_norm)  ->  optimizer.step()  ->  update_ema(...)
//            main.py:410-414     torch.optim.AdamW over three LR groups (default / "backbone" / "text_encoder")
//            util/optim.py:8-25  w_ema = w_ema * decay + (1 - decay) * w  over the whole state_dict
// In the reference this is ~5 GB of HBM traffic spread over several thousand tiny launches (a foreach norm, the multi-tensor
// AdamW passes, three elementwise kernels per state_dict entry for the EMA).  Here parameters, gradients, both Adam moments
// and the EMA copy live in FLAT fp32 buffers with identical element order (tubedetr_b200/optim.py), so the whole step is:
//   tdb_grad_sqnorm   two deterministic passes: per-block sum of squares (fp32 pairwise inside a block, one double per block),
//                     then ONE block adds the partials in fixed order (double) -> total L2 norm on the device (no host sync)
//   tdb_adamw_ema_step  one pass: g *= clip;  AdamW (decoupled weight decay, bias-corrected);  EMA;  bf16 operand copy of the new
//                     weights (what the next forward's GEMMs read: replaces one cast kernel per linear layer)
// Bytes per parameter: read 5 x 4 (p, g, m, v, ema), write 4 x 4 + 2 (p, m, v, ema, bf16) = 38 B -> HBM roofline.
#include "../../include/tubedetr_b200.h"
#include "tdb_common.cuh"

void tdb_count_launch(int n);
int tdb_init_once();
int tdb_num_sms();

namespace tdb {

constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_GROUPS = TDB_OPTIM_MAX_GROUPS;

__global__ void __launch_bounds__(OPT_THREADS) sqnorm_partials_kernel(const float* __restrict__ g, long long n, double* __restrict__ partials) {
  pdl_wait();
  pdl_trigger();
  // each block owns a contiguous chunk (fixed by n and gridDim only -> run-to-run deterministic)
  const long long n4 = n >> 2;
  const long long per = (n4 + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per, hi = min(n4, lo + per);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += OPT_THREADS) {
    const float4 v = __ldcs(g4 + i);
    a0 = fmaf(v.x, v.x, a0);
    a1 = fmaf(v.y, v.y, a1);
    a2 = fmaf(v.z, v.z, a2);
    a3 = fmaf(v.w, v.w, a3);
  }
  double acc = (double)a0 + (double)a1 + (double)a2 + (double)a3;
  if (blockIdx.x == gridDim.x - 1) {                       // tail (n % 4 elements)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) acc += (double)g[i] * (double)g[i];
  }
  __shared__ double red[OPT_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) sqnorm_final_kernel(const double* __restrict__ partials, int nparts, float* __restrict__ norm) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[OPT_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += OPT_THREADS) s += partials[i];   // fixed assignment -> fixed order
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[0] = (float)sqrt(red[0]);
}

struct OptimParams {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;            // nullable
  bf16* p_bf16;          // nullable
  long long n;
  long long g_begin[OPT_MAX_GROUPS], g_end[OPT_MAX_GROUPS];
  float g_lr[OPT_MAX_GROUPS], g_wd[OPT_MAX_GROUPS];
  int ngroups;
  float beta1, beta2, eps;
  float bias_c1, sqrt_bias_c2;      // 1 - beta1^t,  sqrt(1 - beta2^t)
  const float* grad_norm;           // device scalar (nullable = no clipping)
  float max_norm;
  float ema_decay;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr, float wd, const OptimParams& q) {
  // torch.optim.AdamW single-tensor form (torch/optim/adam.py _single_tensor_adam with decoupled weight decay):
  //   p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2) g^2;  p -= (lr / bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
  p = p * (1.f - lr * wd);
  m = m + (g - m) * (1.f - q.beta1);
  v = v * q.beta2 + (1.f - q.beta2) * g * g;
  const float denom = sqrtf(v) / q.sqrt_bias_c2 + q.eps;
  p = p - (lr / q.bias_c1) * (m / denom);
}

__global__ void __launch_bounds__(OPT_THREADS) adamw_ema_kernel(const __grid_constant__ OptimParams q) {
  pdl_wait();
  pdl_trigger();
  float clip = 1.f;
  if (q.grad_norm != nullptr) {      // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to <= 1
    clip = fminf(q.max_norm / (__ldg(q.grad_norm) + 1e-6f), 1.f);
  }
  const float om_decay = 1.f - q.ema_decay;
  const long long n4 = q.n >> 2;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += stride) {
    const long long e = i << 2;
    float4 p = reinterpret_cast<float4*>(q.p)[i];
    const float4 g = __ldcs(reinterpret_cast<const float4*>(q.g) + i);
    float4 m = reinterpret_cast<float4*>(q.m)[i];
    float4 v = reinterpret_cast<float4*>(q.v)[i];
    float lr[4], wd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {    // groups are few (3) and long: the branch is uniform except at two boundaries
      lr[j] = 0.f;
      wd[j] = 0.f;
#pragma unroll 1
      for (int k = 0; k < q.ngroups; ++k)
        if (e + j >= q.g_begin[k] && e + j < q.g_end[k]) {
          lr[j] = q.g_lr[k];
          wd[j] = q.g_wd[k];
        }
    }
    adamw_one(p.x, g.x * clip, m.x, v.x, lr[0], wd[0], q);
    adamw_one(p.y, g.y * clip, m.y, v.y, lr[1], wd[1], q);
    adamw_one(p.z, g.z * clip, m.z, v.z, lr[2], wd[2], q);
    adamw_one(p.w, g.w * clip, m.w, v.w, lr[3], wd[3], q);
    reinterpret_cast<float4*>(q.p)[i] = p;
    __stcs(reinterpret_cast<float4*>(q.m) + i, m);
    __stcs(reinterpret_cast<float4*>(q.v) + i, v);
    if (q.ema != nullptr) {
      float4 a = __ldcs(reinterpret_cast<const float4*>(q.ema) + i);
      a.x = a.x * q.ema_decay + om_decay * p.x;
      a.y = a.y * q.ema_decay + om_decay * p.y;
      a.z = a.z * q.ema_decay + om_decay * p.z;
      a.w = a.w * q.ema_decay + om_decay * p.w;
      __stcs(reinterpret_cast<float4*>(q.ema) + i, a);
    }
    if (q.p_bf16 != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(p.x, p.y);
      o.y = pack_bf16x2(p.z, p.w);
      reinterpret_cast<uint2*>(q.p_bf16)[i] = o;
    }
  }
  // tail (n % 4), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long e = n4 << 2; e < q.n; ++e) {
      float lr = 0.f, wd = 0.f;
      for (int k = 0; k < q.ngroups; ++k)
        if (e >= q.g_begin[k] && e < q.g_end[k]) {
          lr = q.g_lr[k];
          wd = q.g_wd[k];
        }
      float p = q.p[e], m = q.m[e], v = q.v[e];
      adamw_one(p, q.g[e] * clip, m, v, lr, wd, q);
      q.p[e] = p;
      q.m[e] = m;
      q.v[e] = v;
      if (q.ema != nullptr) q.ema[e] = q.ema[e] * q.ema_decay + om_decay * p;
      if (q.p_bf16 != nullptr) q.p_bf16[e] = __float2bfloat16(p);
    }
  }
}

}  // namespace tdb

using namespace tdb;

constexpr int OPT_MAX_BLOCKS = 4096;          // workspace size must not depend on a device query (callable before any launch)
static int optim_blocks() {
  int nb = tdb_num_sms() * 8;
  return nb < 1 ? 1 : (nb > OPT_MAX_BLOCKS ? OPT_MAX_BLOCKS : nb);
}

extern "C" int64_t tdb_optim_workspace_bytes(void) { return (int64_t)OPT_MAX_BLOCKS * (int64_t)sizeof(double); }

extern "C" int tdb_grad_sqnorm(const float* grad, int64_t n, void* workspace, int64_t ws_bytes, float* norm_out, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(grad && workspace && norm_out && n > 0, "tdb_grad_sqnorm: null argument");
  TDB_REQUIRE(((uintptr_t)grad & 15) == 0, "tdb_grad_sqnorm: grad must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream_;
  const int nb = optim_blocks();
  TDB_REQUIRE(ws_bytes >= (int64_t)nb * (int64_t)sizeof(double), "tdb_grad_sqnorm: workspace too small");
  TDB_CHECK_CUDA(tdb_launch(sqnorm_partials_kernel, dim3(nb), dim3(OPT_THREADS), 0, st, grad, (long long)n, (double*)workspace));
  TDB_CHECK_CUDA(tdb_launch(sqnorm_final_kernel, dim3(1), dim3(OPT_THREADS), 0, st, (const double*)workspace, nb, norm_out));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(2);
  return TDB_OK;
}

extern "C" int tdb_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_bf16,
                                  int64_t n, const tdb_optim_group* groups, int ngroups, float beta1, float beta2, float eps,
                                  int64_t step, const float* grad_norm, float max_norm, float ema_decay, void* stream_) {
  int rc = tdb_init_once();
  if (rc) return rc;
  TDB_REQUIRE(param && grad && exp_avg && exp_avg_sq && groups && n > 0, "tdb_adamw_ema_step: null argument");
  TDB_REQUIRE(ngroups >= 1 && ngroups <= OPT_MAX_GROUPS, "tdb_adamw_ema_step: %d groups (max %d)", ngroups, OPT_MAX_GROUPS);
  TDB_REQUIRE(step >= 1, "tdb_adamw_ema_step: step counts from 1");
  TDB_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)ema) & 15) == 0 &&
                  ((uintptr_t)param_bf16 & 7) == 0,
              "tdb_adamw_ema_step: buffers must be 16-byte aligned");
  OptimParams q;
  q.p = param;
  q.g = grad;
  q.m = exp_avg;
  q.v = exp_avg_sq;
  q.ema = ema;
  q.p_bf16 = (bf16*)param_bf16;
  q.n = n;
  q.ngroups = ngroups;
  for (int k = 0; k < OPT_MAX_GROUPS; ++k) {
    q.g_begin[k] = q.g_end[k] = 0;
    q.g_lr[k] = q.g_wd[k] = 0.f;
  }
  for (int k = 0; k < ngroups; ++k) {
    TDB_REQUIRE(groups[k].begin >= 0 && groups[k].end <= n && groups[k].begin <= groups[k].end, "tdb_adamw_ema_step: bad group range");
    q.g_begin[k] = groups[k].begin;
    q.g_end[k] = groups[k].end;
    q.g_lr[k] = groups[k].lr;
    q.g_wd[k] = groups[k].weight_decay;
  }
  q.beta1 = beta1;
  q.beta2 = beta2;
  q.eps = eps;
  q.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
  q.sqrt_bias_c2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  q.grad_norm = grad_norm;
  q.max_norm = max_norm;
  q.ema_decay = ema_decay;
  cudaStream_t st = (cudaStream_t)stream_;
  const long long n4 = n >> 2;
  long long want = (n4 + OPT_THREADS - 1) / OPT_THREADS;
  const int cap = tdb_num_sms() * 8;          // 8 resident CTAs of 256 threads per SM, grid-stride beyond that
  const int nb = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  TDB_CHECK_CUDA(tdb_launch(adamw_ema_kernel, dim3(nb), dim3(OPT_THREADS), 0, st, q));
  TDB_CHECK_CUDA(cudaGetLastError());
  tdb_count_launch(1);
  return TDB_OK;
}
