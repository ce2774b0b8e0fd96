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
 * Process-wide state: the cached driver entry point / SM count (mutex-guarded), per-kernel function attributes set on first use,
 * the launch counter, and two measurement / A-B switches (tdb_xattn_set_timing_buffer, tdb_xattn_set_pair) that are meant to be
 * flipped from a single host thread between launches.  Everything else is re-entrant; the error string is thread-local.
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
#define TDB_GEMM_FLAG_DYNAMIC_TILES (1 << 11)
enum { TDB_OUT_BF16 = 0, TDB_OUT_F32 = 1 };
enum { TDB_REMAP_NONE = 0, TDB_REMAP_COMPACT_TO_PADDED = 1, TDB_REMAP_PADDED_TO_COMPACT = 2,
       /* stride-2 3x3 convolution without an im2col matrix: the producing 1x1 conv writes its NHWC rows (img_h x img_w) space-to-depth,
        * out[(n, h/2 + 1, w/2 + 1)][plane(h&1, w&1) * N + c] with a one-pixel zero halo on the top / left of every parity plane
        * (grid (ceil(H/2) + 1) x (ceil(W/2) + 1), ldo = 4 N); the 9 taps of the convolution are then constant (row shift, plane)
        * offsets into that matrix.  S2D_TO_COMPACT drops the halo positions of such a grid (img_h x img_w = OUTPUT size). */
       TDB_REMAP_COMPACT_TO_S2D = 3, TDB_REMAP_S2D_TO_COMPACT = 4,
       /* zero-haloed grid with ONE shared halo row / column: (img_h + 1) x (img_w + 1) positions per image, data at (h + 1, w + 1).
        * The cell right of the last column is the next row's halo column, the row below the last row is the next image's halo row
        * (beyond the matrix: TMA zero fill), so a 3x3 / pad 1 convolution sees the same zeros as in the (H + 2) x (W + 2) grid of
        * COMPACT_TO_PADDED with 8 % (22 x 22) to 15 % (11 x 11) fewer rows.  The way back is S2D_TO_COMPACT (same formula). */
       TDB_REMAP_COMPACT_TO_PADDED1 = 5 };

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
  /* bring-up only: bit0 swaps LBO/SBO of MN-major operand descriptors; bits1-3 epilogue variant + 1; bit4 no stores;
     bit6 force the 2-CTA kernel; bit7 disable its halo mode; bit8 set a descriptor base offset in halo mode (wrong on purpose);
     bit9 row-per-thread epilogue instead of the shared-memory-box epilogue of the 2-CTA kernel; bit10 never schedule dynamically.
     Scheduling hint (not bring-up): TDB_GEMM_FLAG_DYNAMIC_TILES -- this launch runs next to other streams' kernels (collectives,
     the text encoder): hand out tiles by cluster launch control (work stealing) instead of the static persistent schedule */
  int32_t debug_flags;
} tdb_gemm_desc;

int tdb_gemm(const tdb_gemm_desc* d, void* stream);
/* number of split partials tdb_gemm will actually write for reduction length K and a requested split count */
int tdb_gemm_effective_splits(int K, int splits);
/* sum split-K partials in fixed order: out[m, n] = rowscale[m] * sum_s part[s][m][n]; out fp32 with optional
 * 3x3 re-layout [Cout][9][Cin] -> torch [Cout][Cin][3][3] (taps>1), accumulate!=0 adds into out */
int tdb_splitk_reduce(const float* part, int splits, int M, int N, const float* rowscale, float* out, int taps,
                      int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout transforms around the GEMM (all NHWC bf16 "pixel rows" unless noted).
 * ------------------------------------------------------------------------------------------------ */
/* ResNet stem (reference models/backbone.py:98 -> torchvision conv1 7x7/2/p3): fp32 NCHW frames -> bf16 im2col rows
 * [N*Ho*Wo][192] (147 real columns, order c*49+kh*7+kw); the conv itself + FrozenBN + ReLU is one tdb_gemm */
int tdb_stem_im2col(const float* x, void* col, int N, int H, int W, void* stream);
/* The whole stem in ONE kernel (tdb_stem.cu): conv 7x7/2/p3 + FrozenBatchNorm2d + ReLU + maxpool 3x3/2/p1 (reference
 * models/backbone.py:97-105 -> torchvision conv1 / bn1 / relu / maxpool) from fp32 NCHW frames to bf16 NHWC rows [N*H2*W2][64]:
 * input patch and implicit-im2col A tiles in shared memory, tcgen05.mma into TMEM, pooled epilogue; only the frames and the pooled
 * output touch HBM.  wk = conv1 weight as bf16 [64][192] in K order (c, kh, kw padded to 8), zero at kw = 7 and beyond k = 168. */
int tdb_stem_fused(const float* x, const void* wk, const float* scale, const float* shift, void* out, int N, int H, int W,
                   void* stream);
/* same, pooled rows ldo elements apart (>= 64) */
int tdb_stem_fused_ld(const float* x, const void* wk, const float* scale, const float* shift, void* out, int64_t ldo, int N, int H,
                      int W, void* stream);
/* torchvision maxpool 3x3/2/p1 after the stem (unfused path) */
int tdb_maxpool3x3s2(const void* x, void* y, int N, int H, int W, int C, void* stream);
/* stride-2 3x3 convs (first block of layer2/3/4, torchvision v1.5): explicit im2col [N*Ho*Wo][9*C] (tap major) ... */
int tdb_im2col3x3s2(const void* x, void* col, int N, int H, int W, int C, void* stream);
/* ... and its transpose as a deterministic gather, fused with the ReLU mask (ymask > 0) of the conv input */
int tdb_col2im3x3s2_mask(const void* dcol, const void* ymask, void* dx, int N, int H, int W, int C, void* stream);
/* stride-2 1x1 `downsample` conv input (pixel subsample) and its transpose (zero upsample) */
int tdb_subsample2(const void* x, void* y, int N, int H, int W, int C, void* stream);
/* same, output rows ldy elements apart (the gathered pixels land in a column slice of a wider matrix) */
int tdb_subsample2_ld(const void* x, void* y, int64_t ldy, int N, int H, int W, int C, void* stream);
int tdb_upsample2_zero(const void* y, void* x, int N, int H, int W, int C, void* stream);
/* fp32 torch conv weight [Cout][Cin][kh][kw] -> bf16 GEMM layout [Cout][Kpad] (tap major; taps==49 keeps torch order),
 * optional second copy scaled per output channel (FrozenBN scale folded for dgrad, reference models/backbone.py:60-70) */
int tdb_prep_weight(const float* w, void* out, void* out_scaled, const float* rowscale, int Cout, int Cin, int taps,
                    int Kpad, void* stream);
/* y(bf16) = x (+ add), n % 4 == 0 */
int tdb_cast_add_bf16(const float* x, const float* add, void* y, int64_t n, void* stream);
/* fp32 [rows][K] -> bf16 [rows][2K] = [hi | lo], hi = bf16(x), lo = bf16(x - hi).  The forward GEMMs of the transformer read their
 * weights this way (tdb_gemm with ntaps = 2, a_off0 = {0, 0}, b_off0 = {0, K}: the same A tile against the hi and the lo half), which
 * removes the weight-rounding error of the bf16 path (reference weights are fp32: models/transformer.py:608-676). K % 4 == 0 */
int tdb_split_bf16(const float* x, void* y, int64_t rows, int K, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm with fused residual add (post-norm blocks, reference models/transformer.py:641-645, 721-750, 581;
 * FeatureResizer LN eps 1e-12 :765; also the text encoder's LayerNorms).  D must be 256 or 768.  z = x + r; y = LN(z).  Optional bf16 copies of y and y + pos
 * (the operand of the next projection GEMM).  Backward returns dz (gradient of both x and r) and dgamma/dbeta.
 * Residual dropout (reference transformer.py:641-645 `src + self.dropout1(src2)` etc.): with drop_seed != NULL the kernel
 * computes z = x + keep * r / (1 - p), the keep bits coming from the same counter-based hash as tdb_dropout_mask(seed, site)
 * at element index row * D + col; backward regenerates them and writes dr = keep * dz / (1 - p) (fp32 and/or bf16) next to
 * dz (which then is the gradient of x only).  No mask tensor exists.
 * ------------------------------------------------------------------------------------------------ */
int tdb_layernorm_fwd(const float* x, const float* r, const float* gamma, const float* beta, const float* pos, float* y,
                      void* y_bf, void* ypos_bf, float* mean, float* rstd, int rows, int D, float eps,
                      const int64_t* drop_seed, int64_t drop_site, float drop_p, void* stream);
int tdb_layernorm_bwd_blocks(int rows); /* partial workspace = blocks * 3 * D floats */
/* incoming gradient = dy (fp32, may be NULL) + dy2_bf + dy3_bf (bf16, may be NULL): grads of y, bf16(y), bf16(y+pos) */
int tdb_layernorm_bwd(const float* dy, const void* dy2_bf, const void* dy3_bf, const float* x, const float* r,
                      const float* gamma, const float* mean, const float* rstd, float* dz, void* dz_bf /* optional bf16 copy */,
                      float* dgamma, float* dbeta, float* partial, int rows, int D, int accumulate,
                      const int64_t* drop_seed, int64_t drop_site, float drop_p, float* dr, void* dr_bf,
                      float* dbias /* optional [D]: column sums of the gradient of r = bias gradient of the layer that produced r;
                                      one reduction launch when [dgamma | dbeta | dbias] are contiguous */,
                      void* stream);
/* column sums of a bf16 [rows][N] matrix (bias gradients), two-stage fixed order; partial = nparts * N floats */
int tdb_colsum_bf16(const void* x, int64_t ld, int rows, int N, float* partial, int nparts, float* out, int accumulate,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention core, head_dim 32 (torch F.multi_head_attention_forward need_weights path; reference call sites
 * models/transformer.py:637-640 encoder, 698-722 temporal self-attention, 724-745 time-aligned cross-attention).
 * q/k/v/o: bf16 rows [B*L][>= H*32] with row strides ld*; kpm [B][Lk] nonzero = padded key.
 * p [B][H][Lq][Lk] probabilities (kept for backward); pbar [B][Lq][Lk] = mean over heads (guided-attention loss) or NULL.
 * Backward takes dO and optionally dPbar and returns dq/dk/dv in the layout of q/k/v.
 * Attention dropout (train mode, reference nn.MultiheadAttention(dropout=0.1)): keep [B][H][Lq][Lk] (1 = kept) and
 * keep_scale = 1/(1-p); p then holds the pre-dropout probabilities, pdrop the dropped ones (which pbar averages; pdrop may
 * be NULL when pbar is NULL: the encoder never reads its attention weights).  keep == NULL disables it.
 */
int tdb_mha_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* kpm,
                void* o, int64_t ldo, float* p, float* pbar, const uint8_t* keep, float* pdrop, float keep_scale, int B, int H,
                int Lq, int Lk, float scale, void* stream);
int tdb_mha_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout,
                int64_t lddo, const float* p, const uint8_t* keep, float keep_scale, float* pd_scratch, const float* dpbar,
                float* ds_scratch, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq,
                int Lk, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused time-aligned cross-attention of the space-time decoder (reference models/transformer.py:724-745):
 * K/V projection of the frame memories on tcgen05 (accumulators in TMEM) + scores + softmax + context in ONE kernel, K and V
 * never written to HBM; a small merge kernel combines the per-tile segments of each frame.
 *   q      [F][256] bf16   projected (bias included), unscaled queries, one per frame
 *   mempb  [F*S][256] bf16 memory + position embedding (key input);  memb [F*S][256] bf16 memory (value input)
 *   wkv    [512][256] bf16 rows 256..767 of in_proj_weight (Wk then Wv);  bv [256] fp32 value bias (bk cancels in softmax)
 *   kpm    [F][S] nonzero = padded key
 *   o [F][256] bf16 context (before out_proj);  p [F][8][S] fp32 probabilities;  pbar [F][S] head mean (may be NULL)
 * S must be >= 43 (a 128-row tile may touch at most 4 frames).
 * ------------------------------------------------------------------------------------------------ */
int64_t tdb_xattn_workspace_bytes(int F, int S);
/* measurement hook: buf = device int64 [tiles][4] (NULL = off).  Subsequent tdb_xattn_fused_fwd launches store per 128-row tile
 * the SM clock at kernel entry, first tcgen05.mma issue, completion of the last tcgen05.mma, CTA exit (tensor
 * pipe utilisation of the fused kernel over its MMA phase = 4096 MMA cycles / (stamp[2] - stamp[1])). */
int tdb_xattn_set_timing_buffer(void* buf);
/* kernel variant of tdb_xattn_fused_fwd: 0 = one CTA per 128-row tile, 1 = CTA pair (cluster of 2, tcgen05 cta_group::2) per
 * 256 rows, each CTA staging half of the weight rows (default: env TDB_XATTN_PAIR, else 0).  Results are bit-identical. */
int tdb_xattn_set_pair(int on);
/* keep [F][8][S] (1 = kept) + keep_scale = 1/(1-p): attention dropout (train mode); p stays pre-dropout, pbar averages
 * the dropped probabilities.  keep == NULL disables it. */
int tdb_xattn_fused_fwd(const void* q, const void* mempb, const void* memb, const void* wkv, const float* bv,
                        const uint8_t* kpm, const uint8_t* keep, float keep_scale, void* o, float* p, float* pbar,
                        void* workspace, int64_t ws_bytes, int F, int S, float scale, void* stream);

/* Self-attention core with every contraction on tcgen05 (accumulators in TMEM, Q/K/V/dO by TMA, softmax on 128 threads;
 * tdb_attn_tc.cu): the DEFAULT for the encoder's spatial attention and the decoder's temporal self-attention (reference
 * models/transformer.py:637-640, 698-722).  Same results as tdb_mha_fwd / tdb_mha_bwd; needs an even H, 2 <= Lq <= 256, Lk <= 256
 * (tdb_mha_tc_supported); other shapes use the CUDA-core kernels above.  The caller averages p / pdrop over heads (tdb_head_mean).
 * Attention dropout comes from the library's counter-based stream instead of a mask tensor: drop_seed (device int64, NULL = no
 * dropout), drop_site, drop_p select the same keep bits tdb_dropout_mask(seed, site, p) would write at the flat [B][H][Lq][Lk]
 * index; the backward regenerates them.  p = probabilities before dropout, pdrop (optional) after.
 * tdb_mha_set_tc(level) / env TDB_MHA_TC: 0 = CUDA-core kernels, 1 = tcgen05 forward only, 2 = forward + backward (default). */
int tdb_mha_tc_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* kpm,
                   void* o, int64_t ldo, float* p, float* pdrop, const int64_t* drop_seed, int64_t drop_site, float drop_p,
                   int B, int H, int Lq, int Lk, float scale, void* stream);
/* backward: dPd = dO V^T, dQ = dS K, dK = dS^T Q, dV = Pd^T dO on tcgen05 (the bf16 dS / Pd tiles in shared memory serve as
 * K-major AND MN-major A operands; dK / dV accumulate in TMEM over the query tiles).  dpbar [B][Lq][Lk] = gradient of the
 * head-averaged post-dropout probabilities (guided-attention loss) or NULL.  No scratch buffers. */
int tdb_mha_tc_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout,
                   int64_t lddo, const float* p, const int64_t* drop_seed, int64_t drop_site, float drop_p, const float* dpbar,
                   void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk,
                   float scale, void* stream);
int tdb_mha_tc_supported(int H, int Lq, int Lk);
/* measurement hook: buf = device int64 [16] (NULL = off); CTA 0 of subsequent tdb_mha_tc_bwd launches stores SM clock stamps at its phase
 * boundaries (kernel entry; per query tile: dPd ready, pass A done, pass B done, MMA2 done, epilogues done) */
int tdb_mha_tc_set_timing_buffer(void* buf);
/* pbar[b][i][j] = mean over heads of p[b][h][i][j] (the weights nn.MultiheadAttention returns) */
int tdb_head_mean(const float* p, float* pbar, int B, int H, int Lq, int Lk, void* stream);
int tdb_mha_set_tc(int on);
int tdb_mha_tc_enabled(void);

/* Dropout keep mask (1 = keep, probability 1 - p) from a counter-based hash of (seed[0] in DEVICE memory, site, index):
 * replaces torch.rand -> compare -> cast at the attention-dropout sites (reference models/transformer.py:613, dropout=0.1).
 * The host bumps seed[0] once per training step (a captured op), `site` numbers the call sites within a step. */
int tdb_dropout_mask(uint8_t* keep, int64_t n, const int64_t* seed, int64_t site, float p, void* stream);

/* y = keep ? x / (1 - p) : 0 on a bf16 buffer, keep bits = the stream of tdb_dropout_mask(seed, site) (FFN hidden dropout,
 * reference models/transformer.py:644, 749).  Its backward is done by the consuming GEMM (mask = y > 0, scale 1 / (1 - p)). */
int tdb_dropout_bf16(const void* x, void* y, int64_t n, const int64_t* seed, int64_t site, float p, void* stream);

/* Backward of the one-query-per-frame attention core (the cross-attention above; K/V are re-projected by tdb_gemm first):
 *   q, dout [F][256] bf16;  k, v [F*S][256] bf16 projected keys / values;  p [F][8][S] fp32 probabilities (before dropout);
 *   keep/keep_scale as in forward;  dpbar [F][S] fp32 gradient of the head-mean (post-dropout) probabilities or NULL.
 *   Outputs dq [F][256], dk, dv [F*S][256] bf16.  One streaming pass over V and one over K per frame; deterministic. */
int tdb_xattn_bwd(const void* q, const void* k, const void* v, const void* dout, const float* p, const uint8_t* keep,
                  float keep_scale, const float* dpbar, void* dq, void* dk, void* dv, int F, int S, float scale, void* stream);
/* The same attention core on projected keys / values with ROW STRIDES (elements): the decoder's default path projects K and V of
 * ALL decoder layers with two tdb_gemm launches over the layer-invariant memory (reference models/transformer.py:567-579: every
 * layer receives the same memory / pos) into [F*S][layers*256] buffers; layer l reads / writes its 256-column slice.
 * tdb_xattn_core_fwd: o [F][256] bf16 context, p [F][8][S] fp32 probabilities (before dropout), pbar [F][S] head mean (after
 * dropout) or NULL.  tdb_xattn_core_bwd: tdb_xattn_bwd with strides. */
int tdb_xattn_core_fwd(const void* q, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* kpm,
                       const uint8_t* keep, float keep_scale, void* o, float* p, float* pbar, int F, int S, float scale,
                       void* stream);
int tdb_xattn_core_bwd(const void* q, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout, const float* p,
                       const uint8_t* keep, float keep_scale, const float* dpbar, void* dq, void* dk, int64_t lddk, void* dv,
                       int64_t lddv, int F, int S, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer-side step on FLAT fp32 buffers (parameters, gradients, Adam moments, EMA copy share one element order).
 * Replaces reference engine.py:147-161: torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step (main.py:410-414, three
 * LR groups) + util/optim.py:8-25 update_ema.  SURVEY.md section 8(f).2.
 *   tdb_grad_sqnorm      norm_out[0] = || grad ||_2 (device scalar, deterministic two-pass reduction, double accumulation)
 *   tdb_adamw_ema_step   g *= min(max_norm / (norm + 1e-6), 1)  (grad_norm == NULL: no clipping);
 *                        AdamW with decoupled weight decay and bias correction for step `step` (1-based);
 *                        ema = ema * ema_decay + (1 - ema_decay) * p_new   (ema == NULL: skipped);
 *                        param_bf16[i] = bf16(p_new)                      (NULL: skipped; GEMM operand copy of the new weights)
 *   groups: element ranges [begin, end) of the flat buffers with their own lr / weight_decay (elements outside every
 *           range keep their value: lr = wd = 0, moments still updated).
 * ------------------------------------------------------------------------------------------------ */
#define TDB_OPTIM_MAX_GROUPS 8
typedef struct tdb_optim_group {
  int64_t begin, end;
  float lr, weight_decay;
} tdb_optim_group;
int64_t tdb_optim_workspace_bytes(void);
int tdb_grad_sqnorm(const float* grad, int64_t n, void* workspace, int64_t ws_bytes, float* norm_out, void* stream);
int tdb_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_bf16,
                       int64_t n, const tdb_optim_group* groups, int ngroups, float beta1, float beta2, float eps,
                       int64_t step, const float* grad_norm, float max_norm, float ema_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused data movement around the video-text encoder (tdb_glue.cu; d_model = 256; token rows frame-major; clip(b, t) = b * ceil(T/k) + t / k).
 * ------------------------------------------------------------------------------------------------ */
/* PositionEmbeddingSine(128, normalize=True) in one kernel (reference models/position_encoding.py:71-94): mask [N][h][w] (nonzero = pad)
 * -> out fp32 [N][h*w][256] (channel-last: y features 0..127, x features 128..255, sin / cos interleaved) */
int tdb_pos_sine(const uint8_t* mask, float* out, int N, int h, int w, void* stream);
/* encoder input (reference models/transformer.py:269-331): x32 [n][HW+L][256] = [src rows | text rows of the clip's video], pe = [pos | 0],
 * xb = bf16(x32), xpb = bf16(x32 + pe).  Backward: dsrc = g32 + gb + gpb on the image rows, dtxt = the same summed over a video's clips. */
int tdb_enc_assemble_fwd(const float* src, const float* txt, const float* pos, float* x32, void* xb, void* xpb, float* pe, int n, int HW, int L,
                         int n_clips, void* stream);
int tdb_enc_assemble_bwd(const float* g32, const void* gb, const void* gpb, float* dsrc, float* dtxt, int n, int HW, int L, int n_clips,
                         void* stream);
/* fast branch (reference models/transformer.py:373-391): z [B*T][HW][256] bf16 = enc[clip][hw] + fm; backward: denc [n][S][256] = sum of dz over
 * the clip's frames on the image rows, 0 on the text rows */
int tdb_fast_mix_fwd(const float* enc, const void* fm, void* z, int B, int T, int k, int HW, int S, void* stream);
int tdb_fast_mix_bwd(const void* dz, float* denc, int B, int T, int k, int HW, int S, void* stream);
/* temporal replication + aggregation + decoder operands (reference models/transformer.py:393-446): mem [B*T][S][256] = enc[clip] (+ upd on the
 * image rows, upd may be NULL), mem_pos = pe[clip], memb = bf16(mem), mempb = bf16(mem + mem_pos).  Backward: g = gmem + gmemb + gmempb (any
 * NULL), dupd (fp32 + bf16 copy, may be NULL) = g on the image rows, denc = g summed over the clip's frames. */
int tdb_aggregate_fwd(const float* enc, const float* pe, const float* upd, float* mem, float* mem_pos, void* memb, void* mempb, int B, int T, int k,
                      int HW, int S, void* stream);
int tdb_aggregate_bwd(const float* gmem, const void* gmemb, const void* gmempb, float* denc, float* dupd, void* dupd_b, int B, int T, int k, int HW,
                      int S, void* stream);

/* Last layer of a prediction head (reference models/tubedetr.py:23-42, 226-252: bbox_embed -> sigmoid, sted_embed -> dropout 0.5): x bf16
 * [R][256] (output of the previous layer's GEMM), W fp32 [J][256], b [J], J <= 8; y fp32 [R][J] = dropout(act(x W^T + b)), act 0 none /
 * 1 sigmoid, dropout from the hash stream (drop_seed NULL = off).  Backward: dx bf16 [R][256] (mask_dx: zero where x <= 0 and scaled by
 * dx_scale = ReLU / hidden-dropout backward of the producing layer), dW [J][256], db [J]; dpre [R][J] scratch. */
int tdb_head_out_fwd(const void* x, const float* W, const float* b, float* y, int R, int J, int act, const int64_t* drop_seed,
                     int64_t drop_site, float drop_p, void* stream);
int tdb_head_out_bwd(const float* dy, const float* y, const void* x, const float* W, float* dpre, void* dx, float* dW, float* db, int R, int J,
                     int act, int mask_dx, float dx_scale, const int64_t* drop_seed, int64_t drop_site, float drop_p, void* stream);

/* Input pipeline on the GPU (reference datasets/vidstg.py:104-116, datasets/video_transforms.py, util/misc.py:142-172):
 * decoded rgb24 frames src [T][H0][W0][3] uint8 -> bilinear resize to H x W (half-pixel centres), / 255, (x - mean) / std, written as fp32
 * [T][3][Hp][Wp] into a padded batch slot (zeros outside H x W); mask [T][Hp][Wp] (1 = padding) may be NULL.  mean3 / std3: HOST floats. */
int tdb_frames_preprocess(const uint8_t* src, float* dst, uint8_t* mask, int T, int H0, int W0, int H, int W, int Hp, int Wp,
                          const float* mean3, const float* std3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Text encoder (RoBERTa-base; reference models/transformer.py:130-135, 250-263 calls HF RobertaModel): linear layers on
 * tdb_gemm, LayerNorm + residual (+ dropout) on tdb_layernorm_* (D = 768), and the small kernels below (tdb_text.cu).
 * ------------------------------------------------------------------------------------------------ */
/* erf GELU on bf16 buffers (n even); backward takes the PRE-activation x */
int tdb_gelu_fwd(const void* x, void* y, int64_t n, void* stream);
int tdb_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream);
/* weight / bias gradient of a linear layer with few rows (R tokens): dW [N][K] = dy^T x, db [N] = column sums of dy (db may be NULL);
 * dy bf16 [R][N] (row stride ldy), x bf16 [R][K] (row stride ldx); one launch, rows summed in order */
int tdb_skinny_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* dW, float* db, int R, int N, int K, void* stream);
/* attention core with head_dim 64 for short sequences (L <= 128 forward, <= 100 backward): q, k, v, o bf16 rows [B*L][>= H*64] with row
 * strides, kpm [B][L] nonzero = padded key, p [B][H][L][L] fp32 probabilities before dropout, dropout from the hash stream */
int tdb_text_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* kpm, void* o,
                      int64_t ldo, float* p, const int64_t* drop_seed, int64_t drop_site, float drop_p, int B, int H, int L, float scale,
                      void* stream);
int tdb_text_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout, int64_t lddo,
                      const float* p, const int64_t* drop_seed, int64_t drop_site, float drop_p, void* dq, int64_t lddq, void* dk,
                      int64_t lddk, void* dv, int64_t lddv, int B, int H, int L, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SetCriterion in two launches (reference models/tubedetr.py:270-372, 437-458; util/box_ops.py:65-115):
 * L1 + GIoU on the kept boxes, KL of the start / end distributions, guided-attention loss, for the main output and the auxiliary
 * decoder layers at once.  All tensors fp32, contiguous, per-layer pointers (NULL = that loss family is off).
 *   tdb_criterion_fwd  losses [4][nlayers] = {loss_bbox, loss_giou, loss_sted, loss_guided_attn} x layer
 *   tdb_criterion_bwd  grad_losses [4][nlayers] (upstream gradient of every scalar) -> d_boxes [nlayers][K][4],
 *                      d_sted [nlayers][B][T][2], d_weights [nlayers][B][T][T]
 * ------------------------------------------------------------------------------------------------ */
#define TDB_LOSS_MAX_LAYERS 8
typedef struct tdb_loss_desc {
  int32_t nlayers, K, B, T;
  const float* pred_boxes[TDB_LOSS_MAX_LAYERS];  /* [K][4] cxcywh, already keep-indexed (engine.py:98-102) */
  const float* pred_sted[TDB_LOSS_MAX_LAYERS];   /* [B][T][2] start / end logits */
  const float* weights[TDB_LOSS_MAX_LAYERS];     /* [B][T][T] head-averaged temporal self-attention */
  const float* tgt_boxes;                        /* [K][4] */
  const float* num_boxes;                        /* device scalar (already clamped / averaged over ranks) */
  const float* gauss;                            /* [B][T][2] target distributions (tubedetr.py:318-331) */
  const uint8_t* time_mask;                      /* [B][T] 1 = valid frame */
  const uint8_t* neg;                            /* [B][T] 1 = row excluded from the guided-attention loss (inside the moment / padding) */
  const float* nneg;                             /* [B] number of contributing rows + 1e-6 */
} tdb_loss_desc;
int tdb_criterion_fwd(const tdb_loss_desc* d, float* losses, void* stream);
int tdb_criterion_bwd(const tdb_loss_desc* d, const float* grad_losses, float* d_boxes, float* d_sted, float* d_weights, void* stream);

/* Measurement aid (no reference counterpart): a kernel of `ctas` CTAs with `smem_bytes` of dynamic shared memory each that spins for
 * `cycles` SM clocks -- stands in for a foreign kernel (NCCL all-reduce, text encoder) holding SMs while a GEMM runs (tools/clc_hog.py). */
int tdb_debug_spin(int ctas, int smem_bytes, long long cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TUBEDETR_B200_H */
