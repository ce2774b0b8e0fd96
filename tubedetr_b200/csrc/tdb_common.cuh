// tubedetr_b200 -- shared device helpers: raw PTX for mbarrier / TMA / tcgen05 (sm_100a only).
// No CUTLASS: every instruction the kernels rely on is spelled out here.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef TDB_WATCHDOG_CYCLES
#define TDB_WATCHDOG_CYCLES 4000000000ll  // ~2 s at 1.9 GHz: a legitimate wait never lasts this long; trap instead of hanging the box
#endif

namespace tdb {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TDB_WATCHDOG_CYCLES) {
      printf("tdb watchdog: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ TMA (cp.async.bulk.tensor), 2D tiled
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// c0 = innermost (contiguous) coordinate, c1 = row coordinate; out-of-bounds elements (also negative) are zero-filled
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (lane_base+i), r[j] = column (col+j)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same lane / column mapping as tmem_ld_32x32
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA descriptors (see DESIGN.md "descriptor cheat sheet")
// Shared-memory matrix descriptor, 128-byte swizzle.  Both operand flavours use 8-row x 128-byte swizzle atoms (1024 B):
//   K-major  tile [rows][64 bf16]: row r at r*128 B; SBO = 1024 (next 8 rows); K step of 16 elements = +32 B on the start address
//   MN-major tile chunks of [64 k-rows][64 bf16 of MN]: k-row at k*128 B; SBO = 1024 (next 8 k-rows);
//            LBO = bytes between successive 64-element MN chunks; K step of 16 = +2048 B on the start address
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// The start-address field is the only part of a descriptor that changes inside a main loop: build the constant high part
// once (umma_smem_desc(0, lbo, sbo)) and add (byte address >> 4) per MMA -- one integer add instead of ~10 ALU ops on the
// single issuing thread, whose instruction rate, not the tensor pipe, bounds narrow tiles (profiles/r01_ncu_narrow_conv.txt).
__device__ __forceinline__ uint64_t umma_desc_at(uint64_t desc_hi, uint32_t saddr) { return desc_hi | (uint64_t)((saddr >> 4) & 0x3FFF); }
// one lane of a fully active warp (the same lane every time); lets the surrounding loop stay warp-uniform so that descriptors
// and barrier addresses live in uniform registers instead of being broadcast from a divergent lane before every MMA
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Programmatic dependent launch: every kernel of this library is launched with programmatic stream serialisation (tdb_launch
// below), so its CTAs may become resident -- and run their prologue (barrier init, TMEM allocation, descriptor prefetch,
// shared-memory staging of constants) -- while the previous kernel in the stream is still draining.  pdl_wait() blocks until
// that previous grid has completed and its memory is visible; nothing before it may touch global memory that another kernel
// writes.  pdl_trigger() lets the NEXT kernel in the stream start its own pre-wait part.  Both are no-ops without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Instruction descriptor for kind::f16 with bf16 A/B, fp32 D.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                         // D format f32
         | (1u << 7)                       // A format bf16
         | (1u << 10)                      // B format bf16
         | ((uint32_t)a_mn_major << 15)    // A major: 0 = K, 1 = MN
         | ((uint32_t)b_mn_major << 16)    // B major
         | ((uint32_t)(N >> 3) << 17)      // N / 8
         | ((uint32_t)(M >> 4) << 24);     // M / 16
}

// ------------------------------------------------------------------ thread-block pairs: cluster + tcgen05 cta_group::2
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's mbarrier (address bit 24 = CTA rank, masked off)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// ------------------------------------------------------------------ cluster launch control: hardware work stealing (sm_100)
// A grid is launched with ONE cluster per tile.  A resident cluster finishes its own tile, then asks the launch hardware to CANCEL a
// cluster that has not started yet and takes over that cluster's tile (its ctaid comes back in a 16-byte response, delivered through
// the async proxy with mbarrier complete_tx -- with .multicast::cluster::all to the same shared-memory offset of every CTA of the
// cluster).  The first failed request means no work is left.  Unlike a static `tile += gridDim` schedule, SMs that are busy with
// another stream's kernels (NCCL, the text encoder) when the GEMM starts simply take fewer tiles instead of starting late and
// running their full share after everybody else has finished.
__device__ __forceinline__ void clc_try_cancel(void* resp, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(smem_u32(resp)),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void clc_try_cancel_mc(void* resp, uint64_t* bar) {
  asm volatile(
      "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];" ::"r"(
          smem_u32(resp)),
      "r"(smem_u32(bar))
      : "memory");
}
// -> ctaid.x of the cancelled cluster's first CTA, or -1 when the request failed (nothing left to steal)
__device__ __forceinline__ int clc_query(const void* resp) {
  uint32_t valid, cx, cy, cz;
  asm volatile(
      "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
      "ld.shared.b128 r, [%4];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %3, 1, 0, p1;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n\t}"
      : "=r"(cx), "=r"(cy), "=r"(cz), "=r"(valid)
      : "r"(smem_u32(resp))
      : "memory");
  (void)cy;
  (void)cz;
  return valid ? (int)cx : -1;
}
__device__ __forceinline__ void mbar_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
constexpr int kClcStages = 8;
struct ClcShared {                 // 256 bytes, 16-byte aligned
  uint4 resp[kClcStages];
  uint64_t full[kClcStages];       // response landed (every CTA of the cluster has its own)
  uint64_t empty[kClcStages];      // response read by every consumer warp of the cluster (the leader CTA's copy is the one used)
};
// Per-warp cursor over the tile sequence of its cluster.  Static mode: tile += stride.  CLC mode: the next tile is whatever the
// scheduler warp's cancellation request returned; every consumer warp of the cluster walks the same response ring.
struct TileCursor {
  ClcShared* q;
  int clc, stride, id_shift, s, unit, j, total;     // unit = tiles per claimed cluster id (consecutive tile indices)
  uint32_t ph, leader_empty0;      // cluster address of the leader CTA's empty[0]
  __device__ __forceinline__ void init(ClcShared* q_, int clc_, int stride_, int shift, int unit_, int total_) {
    q = q_; clc = clc_; stride = stride_; id_shift = shift; unit = unit_; total = total_; s = 0; ph = 0; j = 0;
    uint32_t a = smem_u32(&q_->empty[0]), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0u));
    leader_empty0 = r;
  }
  // first tile of this CTA / cluster (id = blockIdx.x >> id_shift)
  __device__ __forceinline__ int first(int id) const { return clc ? id * unit : id; }
  // Leader CTA's producer warp only, at the top of every tile: at the first tile of a unit, claim ONE unit ahead (the response is
  // read by next() when this unit is exhausted), so that a cluster never hoards more than the unit its producer will prefetch next.
  // A unit is sized (host side) to outlast the ~2 us round trip of the request.  The ring is deeper (kClcStages) only because the
  // slower consumer warps (epilogue) read their copy of a response several tiles after the producer did.  Never called again
  // after a failed request (the tile loop ends on it).
  __device__ __forceinline__ void request(int cluster_size) {
    if (!clc || j != 0) return;
    mbar_wait(&q->empty[s], ph ^ 1, 41);
    const uint32_t lane = threadIdx.x & 31;
    if ((int)lane < cluster_size) {
      uint32_t a = smem_u32(&q->full[s]), r;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(lane));
      mbar_expect_tx_remote(r, 16);
    }
    __syncwarp();
    if (elect_one()) {
      if (cluster_size > 1) clc_try_cancel_mc(&q->resp[s], &q->full[s]);
      else clc_try_cancel(&q->resp[s], &q->full[s]);
    }
    __syncwarp();
  }
  // warp-collective (all 32 lanes call it); returns the next tile or INT_MAX
  __device__ __forceinline__ int next(int w) {
    if (!clc) return w + stride;
    if (++j < unit && w + 1 < total) return w + 1;
    j = 0;
    mbar_wait(&q->full[s], ph, 40);
    const int x = clc_query(&q->resp[s]);
    fence_proxy_async();           // my generic-proxy read of the response is ordered before the async-proxy write that recycles it
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_empty0 + (uint32_t)s * 8u) : "memory");
    if (++s == kClcStages) { s = 0; ph ^= 1; }
    return x < 0 ? 0x7fffffff : (x >> id_shift) * unit;
  }
};

// ------------------------------------------------------------------ small numeric helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ counter-based dropout stream (shared by every fused dropout)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t keep4(unsigned long long r, uint32_t thr) {   // four 16-bit lanes -> four 0/1 bytes
  uint32_t o = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) o |= (uint32_t)(((uint32_t)(r >> (16 * t)) & 0xFFFFu) >= thr) << (8 * t);
  return o;
}
// per-element view of the same stream: element e uses 16-bit lane (e & 3) of hash number (e >> 2) -- identical to what
// dropout_mask_kernel writes at keep[e], so a fused consumer and an explicit mask of the same (seed, site) agree bit for bit
__device__ __forceinline__ unsigned long long drop_base(const long long* seed, unsigned long long site) {
  return splitmix64((unsigned long long)seed[0] * 0xD1342543DE82EF95ull + site);
}
__device__ __forceinline__ bool drop_keep(unsigned long long base, long long e, uint32_t thr) {
  const unsigned long long r = splitmix64(base + (unsigned long long)(e >> 2));
  return (((uint32_t)(r >> (16 * (e & 3)))) & 0xFFFFu) >= thr;
}

}  // namespace tdb

// ------------------------------------------------------------------ host side error plumbing (C-ABI never throws)
#define TDB_OK 0
#define TDB_ERR_ARG -1
#define TDB_ERR_CUDA -2
#define TDB_ERR_DRIVER -3
#define TDB_ERR_WORKSPACE -4
void tdb_set_error(const char* fmt, ...);
#define TDB_CHECK_CUDA(x)                                                                           \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) {                                                                        \
      tdb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));             \
      return TDB_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)
#define TDB_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      tdb_set_error(__VA_ARGS__);   \
      return TDB_ERR_ARG;           \
    }                               \
  } while (0)

// ------------------------------------------------------------------ host: launch with programmatic stream serialisation
int tdb_pdl_enabled();   // env TDB_PDL (default 1), defined in tdb_gemm.cu
template <typename... KArgs, typename... Args>
static inline cudaError_t tdb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = tdb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
