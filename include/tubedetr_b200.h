/* tubedetr_b200 -- C ABI of the B200-native TubeDETR hot path (libtdb.so).
 *
 * The reference (antoyang/TubeDETR) has no FFI/plugin layer: its boundary is the Python nn.Module
 * API (models/tubedetr.py:93-101 TubeDETR.forward, :397 SetCriterion.forward).  This C ABI sits
 * directly underneath the drop-in Python modules of tubedetr_b200/ and replaces the vendor-library
 * call sites listed in SURVEY.md section 2.2 (K1..K16).  Every entry point
 *   - takes plain device pointers, sizes and a cudaStream_t (passed as void*), no torch types;
 *   - only enqueues work on the given stream, never synchronises, never allocates;
 *   - returns 0 on success or a negative TDB_ERR_* code; tdb_last_error_string() explains it.
 * The caller (PyTorch, or any host) owns every buffer including workspaces.
 */
#ifndef TUBEDETR_B200_H
#define TUBEDETR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDB_ABI_VERSION 1

int tdb_version(void);
const char* tdb_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t tdb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * tdb_gemm: the tcgen05/TMEM/TMA GEMM behind every convolution and linear layer.
 *   D[m,n] = epilogue( sum_{tap} sum_{k} A[m (+tap row offset), k] * B[n, k (+tap offset)] )
 * Replaces: torchvision Bottleneck 1x1/3x3 convs + FrozenBatchNorm2d (reference models/backbone.py:60-70,
 * 118-122), input_proj (models/tubedetr.py:80,131,134), every nn.Linear / in_proj / out_proj of
 * models/transformer.py:608-751 -- forward, dgrad and wgrad.
 * Operands are bf16 row-major matrices; "major" says which dimension is contiguous:
 *   K-major  (0): matrix is [MN rows][K cols]   (activations [pixels][C], weights [Cout][taps*Cin])
 *   MN-major (1): matrix is [K rows][MN cols]   (dgrad: weights read as [Cout(K)][Cin(N)];
 *                                                wgrad: dY [pixels(K)][Cout(M)], X [pixels(K)][Cin(N)])
 * ------------------------------------------------------------------------------------------------ */
#define TDB_MAX_TAPS 9
enum { TDB_OUT_BF16 = 0, TDB_OUT_F32 = 1 };
enum { TDB_REMAP_NONE = 0, TDB_REMAP_COMPACT_TO_PADDED = 1, TDB_REMAP_PADDED_TO_COMPACT = 2 };

typedef struct tdb_gemm_desc {
  /* operands */
  const void* A; int64_t a_rows, a_cols, lda; int32_t a_major;   /* bf16, [a_rows][a_cols], leading dim lda (elements) */
  const void* B; int64_t b_rows, b_cols, ldb; int32_t b_major;
  int32_t M, N;            /* output tile space: rows m in [0,M), cols n in [0,N); N % 64 == 0 */
  int32_t K;               /* reduction length per tap; K-major operands need K % 64 == 0 (MN-major tails are zero-filled) */
  /* implicit convolution: taps iterated inside the reduction (1 for plain GEMM) */
  int32_t ntaps;
  int32_t a_off0[TDB_MAX_TAPS]; /* per tap, added to A's contiguous coordinate */
  int32_t a_off1[TDB_MAX_TAPS]; /* per tap, added to A's row coordinate       */
  int32_t b_off0[TDB_MAX_TAPS];
  int32_t b_off1[TDB_MAX_TAPS];
  /* z-batch (wgrad over filter taps): nz independent outputs sharing A; B rows shifted, output columns shifted */
  int32_t nz;
  int32_t z_b_off1[TDB_MAX_TAPS];
  int32_t z_out_col[TDB_MAX_TAPS];
  /* split of the reduction across CTAs: splits>1 writes fp32 partials [split][M][ldo] to out, no epilogue */
  int32_t splits;
  /* epilogue: v = acc*scale[n] + bias[n]; v += residual[row,n]; v = relu(v); v = mask[row,n] > 0 ? v : 0 */
  const float* scale; const float* bias;
  const void* residual; int64_t ldr;   /* bf16, indexed by OUTPUT row */
  const void* mask; int64_t ldmask;    /* bf16, indexed by TILE row (same space as A rows) */
  int32_t relu;
  void* out; int32_t out_dtype; int64_t ldo;
  /* row remap between compact NHWC rows and zero-haloed (H+2)x(W+2) rows (implicit 3x3 convolution) */
  int32_t remap, img_h, img_w;
  /* tuning: 0 = auto */
  int32_t block_n; int32_t max_ctas;
  /* bring-up only: bit0 swaps LBO/SBO of MN-major operand descriptors */
  int32_t debug_flags;
} tdb_gemm_desc;

int tdb_gemm(const tdb_gemm_desc* d, void* stream);
/* number of split partials tdb_gemm will actually write for reduction length K and a requested split count */
int tdb_gemm_effective_splits(int K, int splits);
/* sum split-K partials in fixed order: out[m, n] = rowscale[m] * sum_s part[s][m][n]; out fp32 with optional
 * 3x3 re-layout [Cout][9][Cin] -> torch [Cout][Cin][3][3] (taps>1), accumulate!=0 adds into out */
int tdb_splitk_reduce(const float* part, int splits, int M, int N, const float* rowscale, float* out, int taps,
                      int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TUBEDETR_B200_H */
