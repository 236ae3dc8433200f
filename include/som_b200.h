/*
 * som_b200.h — C ABI of the B200-native SOM-layer hot path.
 *
 * The reference (aluo7/ViT-SOM) has no FFI of its own: its hot path is the Python class
 * `SOMLayer` (models/som_layer.py:8-152) calling ATen.  This header is the boundary a
 * maintainer binds instead of those ATen calls.  Every entry point
 *   - takes plain device pointers, sizes and a `cudaStream_t` passed as `void*`,
 *   - never allocates, never synchronises the device, never owns memory,
 *   - returns 0 on success and a negative code on failure; `som_last_error()` returns a
 *     thread-local, human readable description of the last failure.
 * All matrices are row-major fp32 with an explicit leading dimension in ELEMENTS.
 *
 * Reference call sites replaced (file:line in /root/reference):
 *   som_prep_rows            F.normalize / the operand staging of cdist   models/som_layer.py:118-121
 *   som_fwd_distances        torch.cdist(p=2) | 1 - mm(x̂, Ŵᵀ), argmin    models/som_layer.py:87-88,117-122
 *   som_bmu_decode           torch.argmin result as int64                 models/som_layer.py:88
 *   som_neighbourhood        compute_weights                              models/som_layer.py:144-152
 *   som_weighted_loss        som_loss                                     models/som_layer.py:137-142
 *   som_weighted_loss_grad   autograd MeanBackward0/MulBackward0          (autograd of :141-142)
 *   som_bwd_coeffs           EuclideanDistBackward0 / Mm+NormalizeBackward (autograd of :118-122)
 *   som_bwd_dx, som_bwd_dw   the two gradient GEMMs of the same backward
 *   som_forward              SOMLayer.forward in one call                 models/som_layer.py:83-89
 *   som_loss_fused           compute_weights + som_loss (+ backward staging) models/som_layer.py:137-152
 *   som_backward_dw/_dx      backward of the fused loss into prototypes / latents
 *   som_adamw_step           torch.optim.AdamW over som_layer.parameters()   models/vit_som.py:140-151
 *
 * Distance modes: 0 = euclidean (non-squared, ATen `_euclidean_dist` formula
 * sqrt(clamp_min(|x|^2 - 2 x.w + |w|^2, 0))), 1 = cosine (1 - x̂.ŵ with F.normalize eps 1e-12).
 * Manhattan (models/som_layer.py:115-116) is not a contraction and is out of scope.
 */
#ifndef SOM_B200_H_
#define SOM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOM_MODE_EUCLIDEAN 0
#define SOM_MODE_COSINE    1
/*
 * Operand precision, OR-ed into every `mode` argument (all calls of one forward / loss / backward must agree):
 *   0 (default)      3xTF32: staged operands are exact tf32 values in fp32 containers (hi, lo), tcgen05.mma.kind::tf32
 *   SOM_PREC_FP16X3  3xFP16: staged operands are fp16 hi/lo pairs of the ROW-SCALED matrix (each row times the power of
 *                    two that lifts its largest magnitude into [2^14, 2^15)), tcgen05.mma.kind::f16 - the same three
 *                    products hi.hi + hi.lo + lo.hi, the same 22-bit operand split and fp32 accumulation, at twice the
 *                    tensor-core rate and half the operand bytes.  The scales are exact powers of two; the epilogues take
 *                    them out again.  In this mode
 *                      - hi / lo / r_hi / r_lo point to __half matrices (passed through the float pointers), their
 *                        leading dimensions count halves and must be multiples of 8;
 *                      - an aux vector holds 3 * rows + 4 floats: aux | 2^-e | 2^e | 4 statistics words;
 *                      - row_part / col_part need 4 more floats each (1 / S of the staged R, see som_loss_fused);
 *                      - som_forward / som_bmu_decode_scaled also produce the statistic the loss kernel scales R with.
 *                    The general-path entry points (som_bwd_coeffs, som_bwd_dx, som_bwd_dw) take tf32 stagings only.
 */
#define SOM_PREC_FP16X3    16

#define SOM_OK             0
#define SOM_ERR_ARG       -1   /* bad argument (null pointer, misalignment, unsupported size)   */
#define SOM_ERR_CUDA      -2   /* a CUDA runtime / driver call failed                            */
#define SOM_ERR_DEVICE    -3   /* device is not sm_100 (Blackwell B200) - there is no fallback   */

/*
 * GEMM workspace.  Every entry point that launches the tensor-core GEMM takes `ws` / `ws_floats`: an optional,
 * 16-byte aligned device scratch buffer owned by the caller (one per stream in flight).  With it the CTA-pair
 * kernel may schedule stream-K when whole tiles would leave SMs idle: pairs that hold the middle or tail of a tile
 * leave their partial accumulators in the workspace, the pair that holds its head adds them in k order and runs the
 * epilogue (deterministic, no extra launch).  The first 16 KiB hold the hand-over flags: zero them once after
 * allocation (the kernel restores the zero state).  ws == NULL keeps the classic one-tile-at-a-time schedule.
 * som_gemm_workspace_floats() floats always suffice.
 */
int64_t som_gemm_workspace_floats(void);
/* Programmatic dependent launch (default on): the kernels of a step are launched so that each may be scheduled while
 * its predecessor in the stream drains (all of them wait for the predecessor's results before touching memory). */
void som_set_pdl(int on);
/* Stream-K policy: -1 never, 0 cost model (default), 1 whenever a workspace is available and the shape allows. */
void som_set_streamk(int mode);

/* ABI version, bumped on any signature change. */
int som_b200_abi_version(void);

/* Thread-local description of the last error returned on this thread ("" if none). */
const char* som_last_error(void);

/* Number of kernels this library has launched since load (or since the last reset). */
int64_t som_launch_count(void);
void    som_launch_count_reset(void);

/*
 * Operand staging.  For every row r < rows of src[rows, dim] (leading dimension ld_src):
 *   mode 0: aux[r] = |src_r|^2,  v = src_r
 *   mode 1: aux[r] = 1 / max(|src_r|, 1e-12),  v = src_r * aux[r]   (F.normalize)
 * and writes the exact tf32 split hi = rn_tf32(v), lo = rn_tf32(v - hi) to hi/lo[rows, ld_out]
 * (columns dim..ld_out-1 are zero filled).  ld_out must be a multiple of 4, hi/lo 16-byte aligned.
 */
int som_prep_rows(const float* src, int64_t rows, int64_t dim, int64_t ld_src, int mode,
                  float* hi, float* lo, int64_t ld_out, float* aux, void* stream);

/* packed[b] = INT64_MAX for b < B: must precede the first som_fwd_distances of a batch. */
int som_bmu_init(long long* packed, int64_t B, void* stream);

/*
 * Pairwise distances + best-matching-unit search (fused tcgen05 3xTF32 GEMM + epilogue).
 *   x_hi/x_lo [B, ldx], w_hi/w_lo [K, ldw] : staged operands from som_prep_rows (same mode)
 *   x_aux [B], w_aux [K]                   : aux vectors from som_prep_rows (unused for cosine)
 *   dist [B, ldd]                          : out, distances (may be NULL: BMU-only fast path)
 *   packed [B]                             : in/out, per-row min of (ordered(key) << 32 | k + idx_offset),
 *                                            key = squared distance (euclidean) or distance (cosine);
 *                                            combining shards = elementwise signed-int64 min.
 * idx_offset lets a rank that owns prototypes [idx_offset, idx_offset + K) emit global indices.
 */
int som_fwd_distances(const float* x_hi, const float* x_lo, int64_t ldx, const float* x_aux,
                      const float* w_hi, const float* w_lo, int64_t ldw, const float* w_aux,
                      int64_t B, int64_t K, int64_t D, int mode, int64_t idx_offset,
                      float* dist, int64_t ldd, long long* packed,
                      float* ws, int64_t ws_floats, void* stream);

/* bmu[b] = low 32 bits of packed[b] (clamped to [0, K_total)), optional min_key[b] = winning key. */
int som_bmu_decode(const long long* packed, int64_t B, int64_t K_total,
                   int64_t* bmu, float* min_key, void* stream);

/*
 * SOM_PREC_FP16X3 only: som_bmu_decode plus the batch statistic of the backward staging.  The loss kernel stages
 *   R^[b,k] = R_unit[b,k] * 2^-e_b * 2^-g_k * S      (2^e_b, 2^g_k: the row scales of the staged latents / prototypes)
 * with ONE power of two S per call, chosen from  stat = max_b(2^-e_b / d_bmu(b)) * max_k 2^-g_k  (cosine: without the
 * 1 / d) so that R^ stays below 2^14; stat is left in x_aux[3 B] (zeroed by the staging kernel of the same forward).
 *   x_aux [3 B + 4], w_aux [3 K + 4]: aux vectors of this forward's stagings; K = prototypes staged on this rank.
 * Prototype shards call it after the (key, index) minima have been reduced across the ranks.  bmu / min_key may be NULL.
 */
int som_bmu_decode_scaled(const long long* packed, int64_t B, int64_t K_total, int64_t* bmu, float* min_key,
                          float* x_aux, const float* w_aux, int64_t K, int mode, void* stream);

/*
 * Gaussian neighbourhood weights w[b,k] = exp(-|p_k - p_bmu(b)|^2 / (2 T^2)), materialised.
 *   grid_pos [K_total, 2] fp32 (models/som_layer.py:60-81); rows k_offset..k_offset+K are produced.
 *   T_dev : device pointer to the current temperature (fp32).
 */
int som_neighbourhood(const int64_t* bmu, const float* grid_pos, int64_t B, int64_t K,
                      int64_t k_offset, const float* T_dev, float* w, int64_t ldw, void* stream);

/*
 * loss = inv_count * sum_{b,k} w[b,k] * dist[b,k] with w recomputed on the fly (never stored).
 *   partials : device scratch of at least som_loss_scratch_floats(B, K) floats; word 0 is a completion
 *              counter that must be zero on entry (zero the buffer once; the kernel restores the zero
 *              state itself), so one buffer can serve calls of any size on the same stream.
 *   inv_count: 1 / (B_global * K_global) (mean over the full matrix, models/som_layer.py:142).
 */
int64_t som_loss_scratch_floats(int64_t B, int64_t K);
int som_weighted_loss(const float* dist, int64_t ldd, const int64_t* bmu, const float* grid_pos,
                      int64_t B, int64_t K, int64_t k_offset, const float* T_dev, float inv_count,
                      float* partials, float* loss_out, void* stream);

/* G[b,k] = g_out * inv_count * w[b,k]  (upstream gradient of the distances). g_out_dev: device scalar. */
int som_weighted_loss_grad(const int64_t* bmu, const float* grid_pos, int64_t B, int64_t K,
                           int64_t k_offset, const float* T_dev, const float* g_out_dev,
                           float inv_count, float* G, int64_t ldg, void* stream);

/*
 * Backward staging: from the upstream gradient G[B,K] of the distances build the GEMM operand
 * R (tf32 hi/lo split, [B, ldr]) and the rank-1 coefficients of the closed-form backward
 *   euclidean: R = G / dist (0 where dist == 0);  ax[b] = sum_k R, bx = 1;  aw[k] = sum_b R, bw = 1
 *   cosine   : R = G;  c_b = sum_k G (1 - dist), ax = x_aux^2 c_b, bx = x_aux;  same for w.
 * so that  dx = ax * x - bx * (R  . W~)   and   dw = aw * w - bw * (R^T . x~)
 * with W~/x~ the staged (normalised for cosine) operands.  ax, aw must be zero on entry.
 */
int som_bwd_coeffs(const float* G, int64_t ldg, const float* dist, int64_t ldd,
                   int64_t B, int64_t K, int mode, const float* x_aux, const float* w_aux,
                   float* r_hi, float* r_lo, int64_t ldr,
                   float* ax, float* bx, float* aw, float* bw, void* stream);

/* dx[B,D] = ax[b] * x[b,:] - bx[b] * sum_k R[b,k] W~[k,:]   (W~ read MN-major, no transpose copy). */
int som_bwd_dx(const float* r_hi, const float* r_lo, int64_t ldr,
               const float* w_hi, const float* w_lo, int64_t ldw,
               const float* x, int64_t ldx, const float* ax, const float* bx,
               int64_t B, int64_t K, int64_t D, float* dx, int64_t lddx,
               float* ws, int64_t ws_floats, void* stream);

/* dw[K,D] = aw[k] * w[k,:] - bw[k] * sum_b R[b,k] x~[b,:]   (R and x~ read MN-major). */
int som_bwd_dw(const float* r_hi, const float* r_lo, int64_t ldr,
               const float* x_hi, const float* x_lo, int64_t ldx,
               const float* w, int64_t ldw, const float* aw, const float* bw,
               int64_t B, int64_t K, int64_t D, float* dw, int64_t lddw,
               float* ws, int64_t ws_floats, void* stream);

/*
 * ---- Fused protocol entry points: one call per stage of the reference's call sequence ----------------
 * (models/vit_som.py:82-86: forward -> update_temperature -> compute_weights -> som_loss, then backward).
 *
 * som_forward  = SOMLayer.forward (models/som_layer.py:83-89) in three launches:
 *   1. operand staging of the latents (and of the prototypes when stage_w != 0: they only change when the
 *      optimizer steps) + packed[] reset, 2. the tcgen05 distance GEMM with the fused norm / clamp / sqrt /
 *      argmin epilogue, 3. packed -> int64 BMU (skipped when bmu == NULL: a prototype shard decodes after the
 *      cross-rank min reduction).
 *   x [B, ldx], W [K, ldw] raw fp32;  x_hi/x_lo/w_hi/w_lo [rows, ld_stage] staging (ld_stage % 4 == 0, >= D),
 *   x_aux [B], w_aux [K]; dist [B, ldd] may be NULL (argmin-only inference); K_total = map size for decode.
 */
int som_forward(const float* x, int64_t ldx, const float* W, int64_t ldw,
                int64_t B, int64_t K, int64_t D, int mode, int stage_w, int64_t idx_offset,
                float* x_hi, float* x_lo, float* x_aux, float* w_hi, float* w_lo, float* w_aux,
                int64_t ld_stage, float* dist, int64_t ldd, long long* packed, int64_t* bmu,
                int64_t K_total, float* ws, int64_t ws_floats, void* stream);

/*
 * som_loss_fused = compute_weights + som_loss (models/som_layer.py:137-152) and, when r_hi != NULL, the
 * staging of their backward in the same pass over dist[B,K]:
 *   loss_out   = inv_count * sum_{b,k} w[b,k] dist[b,k]              (w recomputed in registers)
 *   R_unit     = inv_count * w / dist (0 where dist == 0) | inv_count * w (cosine), tf32 hi/lo split
 *   row_part [B, n_row_parts]  : partial sums over 128-column slabs of  R_unit | R_unit (1 - dist)
 *   col_part [n_col_parts, K]  : partial sums over row blocks of the same terms
 * (n_row_parts / n_col_parts from som_loss_fused_parts).  The partial sums are plain stores - no atomics, no
 * zero-initialisation - and the gradient GEMMs add the parts of a row in index order, so gradients are run-to-run
 * bit-identical; the loss is reduced in a fixed order in fp64.
 * grid_rows / grid_cols > 0 declare that grid_pos is the canonical square grid (cell k at (k / cols, k % cols),
 * models/som_layer.py:61-67); the weight is then evaluated in its factorised form e[|dr|] * e[|dc|] from a small
 * table (a few ulp from the reference's exp(-(sqrt(dr^2+dc^2))^2 / 2T^2)).  Pass 0, 0 for any other grid (hexa).
 * The upstream gradient of the loss is applied later, in the epilogue of the gradient GEMMs.
 * scratch: at least som_loss_fused_scratch_floats(B, K) floats, word 0 zero on entry (restored on exit).
 * x_aux / w_aux: aux vectors of the forward's stagings; read in SOM_PREC_FP16X3 mode only (NULL otherwise), where R is
 * staged as the fp16 hi/lo split of R_unit * 2^-e_b * 2^-g_k * S (see som_bmu_decode_scaled; values that would leave
 * the fp16 range saturate at 60000) and 1 / S is stored at row_part[B * n_row_parts] and col_part[n_col_parts * K].
 */
int64_t som_loss_fused_scratch_floats(int64_t B, int64_t K);
int som_loss_fused_parts(int64_t B, int64_t K, int64_t* n_row_parts, int64_t* n_col_parts);
int som_loss_fused(const float* dist, int64_t ldd, const int64_t* bmu, const float* grid_pos,
                   int grid_rows, int grid_cols,
                   int64_t B, int64_t K, int64_t k_offset, const float* T_dev, float inv_count, int mode,
                   float* r_hi, float* r_lo, int64_t ldr, float* row_part, float* col_part,
                   float* scratch, float* loss_out, const float* x_aux, const float* w_aux, void* stream);

/*
 * The two gradient GEMMs of the closed-form backward (ATen _euclidean_dist_backward | MmBackward0 +
 * normalize backward), with g = *g_dev the upstream gradient of the loss:
 *   dW[k,:] (+)= g * (c_k W[k,:] - s_k sum_b R_unit[b,k] x~[b,:])     c, s from col_part / w_aux
 *   dx[b,:] (+)= g * (c_b x[b,:] - s_b sum_k R_unit[b,k] W~[k,:])     c, s from row_part / x_aux
 * euclidean: c = sum of the row's parts, s = 1;  cosine: c = aux^2 * sum, s = aux.  accumulate != 0 adds into the
 * output (row-chunked batches accumulate dW across chunks).  sm_limit > 0: the launch occupies at most that many SMs
 * (the caller runs a collective kernel beside it); 0 = all.
 * SOM_PREC_FP16X3: r_hi / r_lo hold the scaled R^ of som_loss_fused and the staged operands their row-scaled fp16 split;
 * the epilogue multiplies s of output row m by 2^e_m (x_aux[2 B + m] for dx, w_aux[2 K + m] for dW) and by 1 / S
 * (row_part[B * n_row_parts] for dx, col_part[n_col_parts * K] for dW), so dW / dx come out unscaled; x_aux / w_aux are
 * then required for both distance modes.
 */
int som_backward_dw(const float* r_hi, const float* r_lo, int64_t ldr,
                    const float* x_hi, const float* x_lo, int64_t ld_stage,
                    const float* W, int64_t ldw, const float* col_part, int64_t n_col_parts, const float* w_aux,
                    const float* g_dev, int64_t B, int64_t K, int64_t D, int mode,
                    float* dW, int64_t lddw, int accumulate, int sm_limit,
                    float* ws, int64_t ws_floats, void* stream);
int som_backward_dx(const float* r_hi, const float* r_lo, int64_t ldr,
                    const float* w_hi, const float* w_lo, int64_t ld_stage,
                    const float* x, int64_t ldx, const float* row_part, int64_t n_row_parts, const float* x_aux,
                    const float* g_dev, int64_t B, int64_t K, int64_t D, int mode,
                    float* dx, int64_t lddx, int accumulate, int sm_limit,
                    float* ws, int64_t ws_floats, void* stream);

/*
 * Both gradient GEMMs above in ONE persistent CTA-pair launch (their tiles share one stream-K work list, dW tiles
 * first): same results as som_backward_dw followed by som_backward_dx, one prologue / tail instead of two.  Needs the
 * GEMM workspace; without it (or for tiny shapes) it issues the two launches.
 * Counted launches (done != NULL): a two-phase schedule - every CTA pair first works off its share of the COUNTED
 * GEMM's tiles (dW when count_dx == 0: data parallelism; dx when count_dx != 0: prototype shards), then its share of
 * the other's.  Every finished 32-row slab of the counted result adds 1 to *done (a zero-initialised device word) and
 * *done_expected (host) receives the final count, so the exchange of that result can start from another stream behind
 * som_stream_wait_value(done, expected) while the other GEMM's tiles are still running; sm_limit > 0 then bounds the
 * SMs of the SECOND phase only (the pairs beyond it exit after phase 1 and free their SMs for the exchange kernel).
 * *done_expected = -1: not counted (fallback launches: order the exchange after this call).
 */
int som_backward_fused(const float* r_hi, const float* r_lo, int64_t ldr,
                       const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo, int64_t ld_stage,
                       const float* x, int64_t ldx, const float* W, int64_t ldw,
                       const float* row_part, int64_t n_row_parts, const float* col_part, int64_t n_col_parts,
                       const float* x_aux, const float* w_aux,
                       const float* g_dev, int64_t B, int64_t K, int64_t D, int mode,
                       float* dW, int64_t lddw, int accumulate_dw, float* dx, int64_t lddx,
                       int sm_limit, int count_dx, unsigned int* done, int64_t* done_expected,
                       float* ws, int64_t ws_floats, void* stream);

/* Stream-ordered memory operations (cuStreamWaitValue32 GEQ / cuStreamWriteValue32): `stream` proceeds once
 * *addr >= value; executed by the GPU front end (no SM is occupied while waiting), capturable in CUDA graphs. */
int som_stream_wait_value(unsigned int* addr, unsigned int value, void* stream);
int som_stream_write_value(unsigned int* addr, unsigned int value, void* stream);

/*
 * Prototype optimizer step fused with the operand staging of the next forward (the reference optimises
 * som_layer.parameters() with torch.optim.AdamW, models/vit_som.py:140-151): one pass that reads W, dW, m, v and
 * writes W, m, v AND, when w_hi != NULL, the staging of the new prototypes (tf32 hi/lo split + |w|^2 or 1/max(|w|,
 * 1e-12), exactly what som_prep_rows produces), so the next som_forward runs with stage_w = 0.
 *   hp_dev: device float[3] = {lr, t, grad_scale}: learning rate, 1-based step count, factor applied to dW
 *   beta1, beta2, eps, weight_decay: as torch.optim.AdamW (decoupled decay, bias-corrected moments, no amsgrad)
 */
int som_adamw_step(float* W, int64_t ldw, const float* dW, int64_t lddw, float* m, float* v, int64_t ldm,
                   int64_t K, int64_t D, const float* hp_dev, double beta1, double beta2, double eps,
                   double weight_decay, int mode, float* w_hi, float* w_lo, int64_t ld_stage, float* w_aux,
                   void* stream);

/*
 * Exchange step of batch-sharded data parallelism (the reference gets it from Lightning DDP's bucketed NCCL
 * all-reduce, experiments/benchmarking/train_vit_som.py:45,86-87): in-place MEAN over the ranks of the prototype
 * gradient, as a two-shot NVLS (NVLink SHARP) all-reduce - multimem.ld_reduce of this rank's slice (summed in the
 * NVSwitch), multimem.st of the mean to every replica - in one kernel with its own cross-GPU barriers.
 *   mc_ptr    multicast address of the buffer (it lives at the same offset of a symmetric allocation on every rank)
 *   flag_ptrs device array of `world` pointers to the ranks' zero-initialised flag buffers (peer mapped), each of at
 *             least som_nvls_flag_words(world) 32-bit words; the kernel leaves them zero again
 *   n_floats  multiple of 4.   Every rank of the group must make the call.
 *   blocks    grid size (512 threads each, two per SM; <= 64, the same on every rank): the kernel is bound by the bytes
 *             it keeps in flight, so give it two blocks per SM that is free - 2 x (SMs - sm_limit) beside a GEMM.
 */
int som_allreduce_mean_nvls(float* mc_ptr, void* flag_ptrs, int64_t n_floats, int rank, int world, int blocks, void* stream);
/* The same kernel with an explicit factor (1.0f = SUM): the partial dx[B,D] of prototype shards. */
int som_allreduce_nvls(float* mc_ptr, void* flag_ptrs, int64_t n_floats, int rank, int world, float scale, int blocks,
                       void* stream);
int64_t som_nvls_flag_words(int world);

/*
 * Diagnostic entry point (used by the tests to validate the tensor-core mainloop in isolation):
 * C[M,N] = A . B^T in 3xTF32 with A = a_hi + a_lo, B = b_hi + b_lo.
 *   a_mn = 0: A stored [M, Kred] (K-major)    a_mn = 1: A stored [Kred, M] (MN-major)
 *   b_mn = 0: B stored [N, Kred] (K-major)    b_mn = 1: B stored [Kred, N] (MN-major)
 *   bn: tile width (16, 32, 64, 96 or 128; 0 = auto), kchunk: k-blocks (of 32) per tensor-core
 *   accumulation chunk (0 = default), passes: 3 = 3xTF32, 1 = hi*hi only.
 */
int som_debug_gemm(const float* a_hi, const float* a_lo, int64_t lda, int a_mn,
                   const float* b_hi, const float* b_lo, int64_t ldb, int b_mn,
                   int64_t M, int64_t N, int64_t Kred, int bn, int kchunk, int passes,
                   float* C, int64_t ldc, float* ws, int64_t ws_floats, void* stream);

/* The same mainloop in 3xFP16: hi / lo are __half matrices (no row scaling here), lda / ldb in halves (multiples of 8),
 * k-blocks of 64; bn must be a multiple of 128 when an MN-major B is read by CTA pairs. */
int som_debug_gemm_f16(const void* a_hi, const void* a_lo, int64_t lda, int a_mn,
                       const void* b_hi, const void* b_lo, int64_t ldb, int b_mn,
                       int64_t M, int64_t N, int64_t Kred, int bn, int kchunk, int passes,
                       float* C, int64_t ldc, float* ws, int64_t ws_floats, void* stream);

/* Diagnostics, host only: the work decomposition of the CTA-pair kernel for `workers` CTA pairs over the k-block units
 * of one or two GEMMs (tiles x k-blocks each; tiles1 = 0 for one GEMM).  split > 0: tile-aligned split-K, 0: even
 * ranges (stream-K).  bounds_out[0..workers]: first unit of every worker, then the total.  split < 0: the two-phase
 * schedule of the data-parallel backward (every worker: its share of GEMM 0, then its share of GEMM 1);
 * bounds_out[0..workers] = phase 0, bounds_out[workers+1 .. 2 workers+1] = phase 1. */
int som_debug_schedule(int64_t tiles0, int64_t nkb0, int64_t tiles1, int64_t nkb1, int workers, int split,
                       int64_t* bounds_out);

/* Tuning knobs (process-wide): tile width override (0 = auto) and k-blocks per accumulation chunk. */
void som_set_tuning(int bn_override, int kchunk);
/* Kernel selection: 0 = cost model (default), 1 = single-CTA 128 x bn tiles, 2 = CTA-pair (cta_group::2) 256 x bn tiles. */
void som_set_cta_group(int cg);
/* Diagnostics: bit 0 = no TMA loads after the first ring pass, bit 1 = no tensor-core instructions, bit 2 / 3 = gradient
 * epilogue without src loads / without stores (results are garbage with any of these); bit 4 = one 2-D TMA operation per
 * MN-major panel instead of one 3-D operation per tile (results unchanged); bit 5 = always the generic fused loss kernel
 * instead of the straight-line fast path for aligned square maps (results equal to rounding). */
void som_set_debug(int bits);
/* Diagnostics: device buffer of 16 uint64 in which CTA 0 of the pair kernel stamps %globaltimer (ns) at its phase
 * boundaries (start, setup done, producer done, issuer done, last accumulators ready, epilogue done, pair synced,
 * TMEM freed); NULL switches it off. */
void som_set_debug_times(unsigned long long* dev_buf);

#ifdef __cplusplus
}
#endif
#endif  /* SOM_B200_H_ */
