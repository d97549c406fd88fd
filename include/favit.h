/*
 * favit.h — C ABI of the B200-native MHLA + SPPP hot path (libfavit_b200.so).
 *
 * The reference (zser092/Focused-Attention-ViT) is pure Python/PyTorch and has no FFI; the "interface"
 * each entry point replaces is therefore a span of eager PyTorch code inside a reference method, cited
 * per function below as <file>:<lines> under /root/reference.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add to call these.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer on the current device;
 *   - the caller allocates every output / workspace and passes the stream to launch on;
 *   - return 0 on success, a favit_status otherwise; favit_last_error() returns a thread-local
 *     message; nothing throws, allocates device memory or synchronises the stream;
 *   - kernels are stateless and re-entrant; the library is sm_100a only.
 */
#ifndef FAVIT_H_
#define FAVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FAVIT_F32 = 0, FAVIT_BF16 = 1 } favit_dtype;

typedef enum {
  FAVIT_OK = 0,
  FAVIT_ERR_ARG = 1,          /* bad argument (null pointer, negative size, misalignment) */
  FAVIT_ERR_UNSUPPORTED = 2,  /* valid request this build does not implement             */
  FAVIT_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed                      */
  FAVIT_ERR_WORKSPACE = 4     /* workspace too small                                      */
} favit_status;

typedef enum {
  FAVIT_EPI_NONE = 0,       /* C = A.B (+bias)                                   */
  FAVIT_EPI_GELU = 1,       /* C = gelu(A.B + bias)  (erf form, nn.GELU default)   */
  FAVIT_EPI_DGELU_MUL = 2   /* C = (A.B) * gelu'(aux)   aux has C's shape/ld      */
} favit_epilogue;

typedef void* favit_stream;   /* cudaStream_t */

int favit_version(void);
/* compute capability of the current device, major*10+minor (100 on B200); < 0 on error */
int favit_device_cc(void);
const char* favit_last_error(void);
/* Thread-local description of the kernel variant the last dispatching entry point chose on this thread, e.g.
 * "gemm_bf16_tcgen05_2cta_kernel<AUX=0> act=1 ..." or "sppp_pool_fwd_tma_kernel grid=592 ...": lets a test assert that
 * the code path it means to check is the one that ran. */
const char* favit_last_kernel(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t favit_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * MHLA windowed attention core — replaces models/mhla.py:109-154 (index table, both gathers, scaled
 * scores, mask, softmax, PV) and never materialises the [B,H,N,W,hd] windows.
 *
 *   out[b,i,h,:] = sum_j m(i,j) exp(s_ij - lse) v[b,j,h,:],  s_ij = scale * q[b,i,h,:].k[b,j,h,:]
 *   m(i,j) = window multiplicity of mhla.py:63-81 (band + duplicated edge index).
 *
 * q,k,v: element (b,i,h,d) at ptr + b*stride_b + i*stride_n + h*stride_h + d   (strides in elements;
 *        the packed [B,N,3,H,hd] output of the qkv GEMM is read in place).
 * mask:  NULL or uint8 [B,N,N], 0 = masked (mhla.py:136-143).
 * out:   [B,N,H,hd] contiguous (== [B,N,D], the layout proj consumes, mhla.py:157); lse: fp32 [B,H,N].
 * window may be even only when N <= window (the reference raises otherwise, mhla.py:83).
 * dropout_p > 0 applies attention-probability dropout (mhla.py:147) with a counter-based generator
 * keyed by (seed, b, h, i, slot); backward regenerates the same keep-mask.
 * ---------------------------------------------------------------------------------------------- */
int favit_mhla_attn_fwd(const void* q, const void* k, const void* v, const uint8_t* mask,
                        void* out, float* lse,
                        int B, int H, int N, int hd, int window, float scale,
                        int64_t stride_b, int64_t stride_n, int64_t stride_h,
                        favit_dtype dtype, float dropout_p, uint64_t seed, favit_stream stream);

/* Backward of the above (the reference relies on autograd: gather -> scatter_add_, bmm, softmax).
 * dout: [B,N,H,hd] contiguous.  dq,dk,dv use the same strides as q,k,v (written in place into a packed
 * dqkv buffer).  delta: fp32 workspace [B,H,N].  dqkv_colsum (may be NULL): [3*H*hd] fp32, ACCUMULATED with the column
 * sums of dq | dk | dv over all B*N rows — the bias gradient of the qkv projection (mhla.py:100), produced while the
 * gradient tiles are still in shared memory. */
int favit_mhla_attn_bwd(const void* q, const void* k, const void* v, const uint8_t* mask,
                        const void* out, const float* lse, const void* dout,
                        void* dq, void* dk, void* dv, float* delta, float* dqkv_colsum,
                        int B, int H, int N, int hd, int window, float scale,
                        int64_t stride_b, int64_t stride_n, int64_t stride_h,
                        favit_dtype dtype, float dropout_p, uint64_t seed, favit_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Linear layers of the block — replace nn.Linear forward/backward at models/mhla.py:100 (qkv) and
 * :158 (proj) (the latent projection :105-106 is folded into their weights by the host side) and,
 * for the "next" row, the MLP (models/vit.py:107-139, models/mhla.py:197-203).
 * dtype FAVIT_BF16: tcgen05.mma tiles, TMEM accumulators, TMA-fed (operands bf16, fp32 accumulate).
 * dtype FAVIT_F32 : SIMT FFMA tiles (parity path for the 1e-4 bar; everything fp32).
 * Row-major, leading dimensions in elements.  bias / db / dw are always fp32.
 *
 * favit_linear_fwd : Y[M,N] = act(X[M,K] . W[N,K]^T + bias[N]) + residual[M,N]
 *      epilogue FAVIT_EPI_NONE or FAVIT_EPI_GELU; with GELU and preact_out != NULL the pre-activation
 *      is also written (dtype = `dtype`, leading dimension ldy) for the backward pass.
 *      bias, residual, preact_out may be NULL.  y_dtype/res_dtype select fp32 or bf16 for Y / residual
 *      (the fp32 residual stream of a bf16 block); they must be FAVIT_F32 when dtype is FAVIT_F32.
 * favit_linear_dgrad: dX[M,K] = dY[M,N] . W[N,K]   (epilogue FAVIT_EPI_DGELU_MUL: * gelu'(preact[M,K]),
 *      preact has dX's leading dimension).  dx_colsum (may be NULL): [K] fp32, ACCUMULATED with the column sums of dX
 *      as stored — dX is the output gradient of the layer below, so this is that layer's bias gradient, produced in
 *      the GEMM epilogue instead of a second pass over dX.
 * favit_linear_wgrad: dW[N,K] (+)= dY[M,N]^T . X[M,K]  and, when db != NULL, db[N] (+)= column sums of dY.
 *      accumulate == 0 overwrites (the call zero-fills first), != 0 adds to the existing values.
 *      Split-K partial sums are combined with fp32 atomics, so the last bits may vary between runs.
 * ---------------------------------------------------------------------------------------------- */
int favit_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                     void* preact_out, int M, int N, int K, int64_t ldx, int64_t ldw, int64_t ldy,
                     int64_t ldres, favit_dtype dtype, favit_dtype y_dtype, favit_dtype res_dtype,
                     int epilogue, favit_stream stream);

int favit_linear_dgrad(const void* dy, const void* w, const void* preact, void* dx, float* dx_colsum, int M, int N,
                       int K, int64_t lddy, int64_t ldw, int64_t lddx, favit_dtype dtype, favit_dtype dx_dtype,
                       int epilogue, favit_stream stream);

int favit_linear_wgrad(const void* dy, const void* x, float* dw, float* db, int M, int N, int K,
                       int64_t lddy, int64_t ldx, int64_t lddw, favit_dtype dtype, int accumulate,
                       favit_stream stream);

/* The same two entry points with the MLP's nn.Dropout (models/vit.py:122, 131-138; main.py:106 `--dropout 0.1`) fused into
 * the epilogue, bf16 path only: after the activation and before the residual the value is multiplied by
 * keep(row, col) / (1 - p).  The keep-mask is counter based and is REGENERATED in backward, never stored: element
 * (row, col) of an [M, N] result belongs to group g = row * ceil(N / 4) + col / 4; splitmix64(*drop_seed + drop_offset +
 * g * 0x9E3779B97F4A7C15) yields four 16-bit uniforms, lane col % 4 is kept iff it is >= round(p * 65536).  drop_seed is a
 * DEVICE pointer (a captured CUDA graph draws a fresh mask per replay when the caller bumps it in-stream), drop_offset
 * tells layers / sites apart.  The reference draws its mask from torch's Philox stream, so only the distribution is
 * the reference's; oracle/mhla_oracle.py:mlp_dropout_keep_mask reproduces this generator bit for bit.
 *   fwd  : Y = dropout(act(X.W^T + bias)) + residual   (GELU: the saved pre-activation is NOT masked)
 *   dgrad: dX = (dY.W) * gelu'(preact) * keep / (1 - p)   — the mask of the dropout that followed the activation */
int favit_linear_fwd_dropout(const void* x, const void* w, const float* bias, const void* residual, void* y,
                             void* preact_out, int M, int N, int K, int64_t ldx, int64_t ldw, int64_t ldy,
                             int64_t ldres, favit_dtype dtype, favit_dtype y_dtype, favit_dtype res_dtype,
                             int epilogue, float drop_p, const uint64_t* drop_seed, uint64_t drop_offset,
                             favit_stream stream);
int favit_linear_dgrad_dropout(const void* dy, const void* w, const void* preact, void* dx, float* dx_colsum, int M,
                               int N, int K, int64_t lddy, int64_t ldw, int64_t lddx, favit_dtype dtype,
                               favit_dtype dx_dtype, int epilogue, float drop_p, const uint64_t* drop_seed,
                               uint64_t drop_offset, favit_stream stream);
/* out[M,N] (fp32 or bf16, contiguous) = keep / (1 - p) * g[M,N] (fp32) with the mask above, and colsum[N] (may be NULL,
 * ACCUMULATED) += column sums of out: the gradient that enters fc2 behind its dropout (vit.py:138) as a GEMM operand,
 * and fc2's bias gradient.  N % 8 == 0. */
int favit_dropout_cast(const float* g, void* out, favit_dtype out_dtype, float* colsum, int M, int N, float drop_p,
                       const uint64_t* drop_seed, uint64_t drop_offset, favit_stream stream);

/* out[n] += sum over the M rows of x[m,n] (x row-major with leading dimension ld): a bias gradient. */
int favit_colsum(const void* x, favit_dtype dtype, float* out, int M, int N, int64_t ld, favit_stream stream);

/* fp32 -> bf16 copies of `count` tensors in one launch per 32 tensors: the per-step cast of the fp32 master weights of
 * the MLP (models/vit.py:107-139 fc1 / fc2 of every block) into tcgen05 GEMM operands.  src / dst / numel are HOST
 * arrays of device pointers / element counts. */
int favit_cast_bf16_batched(int count, const void* const* src, void* const* dst, const int64_t* numel,
                            favit_stream stream);
/* The same with any of the four fp32 / bf16 conversions (fp32 -> fp32 is a plain multi-tensor copy): also packs the small
 * gradients of a model into one flat buffer, and unpacks them, around the data-parallel all-reduce. */
int favit_copy_batched(int count, const void* const* src, void* const* dst, const int64_t* numel,
                       favit_dtype src_dtype, favit_dtype dst_dtype, favit_stream stream);

/* Test / tuning hook: C[M,N] = A.B^T through the tcgen05 kernel with explicit operand storage
 * (a_mn / b_mn: 0 = reduction dimension contiguous, 1 = M/N dimension contiguous), tile width
 * (0 = auto, 64, 128, 256) and split-K factor (0 = auto).  C fp32 or bf16. */
int favit_gemm_bf16_raw(const void* a, int a_mn, int64_t lda, const void* b, int b_mn, int64_t ldb, void* c,
                        int64_t ldc, favit_dtype c_dtype, int M, int N, int K, int bn, int splits,
                        favit_stream stream);

/* How the persistent CTA-pair GEMM (every large Linear of models/mhla.py:100,158 and models/vit.py:107-139) hands its
 * 256 x 256 tiles to the 74 CTA pairs.  0 (default) = static striding: fastest when the GPU runs nothing else.
 * 1 = work stealing: a CTA pair that becomes resident late, because another kernel (the NCCL all-reduce of
 * data-parallel training, which overlaps the backward GEMMs) holds some SMs, finds its tiles taken by the others instead
 * of running them as a second wave; costs ~1 % without contention.  Takes effect for launches (and graph captures) made
 * after the call; process-wide.  Any other value only queries.  Returns the mode in force.  No reference counterpart:
 * the reference is single-process.
 * Limits of mode 1: every launch takes the next of 64 sets of scheduler words (round robin) and re-arms it when it
 * ends, so launches that RUN CONCURRENTLY must be fewer than 64 launches apart; a launch captured into a CUDA graph
 * keeps its set, so such a graph must not be replayed concurrently with itself or beside other mode-1 launches on
 * another stream (stream-ordered use, the training step, is always safe). */
int favit_set_gemm_tile_scheduler(int mode);

/* ------------------------------------------------------------------------------------------------
 * Latent-projection fold — replaces the two per-token Linear(hd, hd) applications of models/mhla.py:105-106 by a
 * per-step transformation of the qkv / proj WEIGHTS (exact): fwd writes the folded weights in the GEMM
 * operand dtype plus fp32 folded biases; bwd maps the gradients of the folded weights back to qkv.weight, qkv.bias,
 * proj.weight (in place) and produces latent_proj.weight / latent_proj.bias gradients.  All inputs fp32; hd <= 64.
 * ---------------------------------------------------------------------------------------------- */
int favit_latent_fold_fwd(const float* qkv_w, const float* qkv_b, const float* proj_w, const float* proj_b,
                          const float* lat_w, const float* lat_b, void* wqkv_c, float* bqkv, void* wproj_c,
                          float* bproj, int H, int hd, favit_dtype out_dtype, favit_stream stream);

int favit_latent_fold_bwd(const float* qkv_w, const float* qkv_b, const float* proj_w, const float* lat_w,
                          const float* lat_b, float* dwqkv, float* dbqkv, float* dwproj, const float* dbproj,
                          float* dlat_w, float* dlat_b, int H, int hd, favit_stream stream);

/* The same two folds for L layers of one model in one call (the work is weight-only and tiny: a launch per layer
 * costs more than the arithmetic).  ptrs is a HOST array of L rows of device pointers, in the argument order of the
 * single-layer functions:
 *   fwd, 10 per layer: qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, wqkv_c, bqkv, wproj_c, bproj
 *   bwd, 11 per layer: qkv_w, qkv_b, proj_w, lat_w, lat_b, dwqkv, dbqkv, dwproj, dbproj, dlat_w, dlat_b
 * All layers share H, hd and (fwd) out_dtype. */
int favit_latent_fold_fwd_batched(int L, const void* const* ptrs, int H, int hd, favit_dtype out_dtype,
                                  favit_stream stream);

int favit_latent_fold_bwd_batched(int L, const void* const* ptrs, int H, int hd, favit_stream stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm around the attention / MLP — "next" row: nn.LayerNorm at models/vit_mhla.py:88,107
 * (norm1 / norm2 of the block) and its autograd.  fp32 statistics; D % 4 == 0, D <= 1024.
 *
 * favit_layernorm_fwd: y[M,D] = (x - mean) * rstd * gamma + beta; mean / rstd [M] fp32 are saved for backward.
 *      With delta != NULL (dtype = y_dtype) the residual add in front of the norm is fused: xsum = x + delta is
 *      written (x's dtype) and normalised — the attention output joins the residual stream here instead of in
 *      the GEMM epilogue (models/vit_mhla.py:104-107).
 * favit_layernorm_bwd: dx[M,D] (fp32) = LN'(dy) + dres   (dres: fp32 gradient of the residual branch, may be NULL);
 *      dx_bf16 (may be NULL) receives a bf16 copy of dx (the operand of the next dgrad / wgrad GEMM);
 *      dgamma / dbeta [D] fp32 are ACCUMULATED into (zero them first); both NULL to skip.
 *      dxsum [D] (may be NULL) accumulates the column sums of dx: dx is the output gradient of the linear layer
 *      that fed this norm's residual input, so dxsum is that layer's bias gradient (no separate reduction pass).
 * ---------------------------------------------------------------------------------------------- */
int favit_layernorm_fwd(const void* x, favit_dtype x_dtype, const void* delta, void* xsum, const float* gamma,
                        const float* beta, void* y, favit_dtype y_dtype, float* mean, float* rstd, int M, int D,
                        float eps, favit_stream stream);

int favit_layernorm_bwd(const void* dy, favit_dtype dy_dtype, const void* x, favit_dtype x_dtype,
                        const float* mean, const float* rstd, const float* gamma, const float* dres, float* dx,
                        void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int M, int D,
                        favit_stream stream);

/* ------------------------------------------------------------------------------------------------
 * SPPP patch -> superpixel assignment — replaces PatchToSuperpixelMapper.map_patches,
 * models/sppp.py:91-128, for a whole batch in one call (the reference loops per image,
 * models/sppp_mhla.py:286-297).
 *
 * labels: int64 [B, img_h, img_w] (row stride img_w).  grid = img_size / patch (trailing pixels are
 * ignored like sppp.py:102).  P = grid*grid.
 * dom[B,P]      int64  dominant label of each patch (most frequent, ties -> smallest id, sppp.py:117-120)
 * slot[B,P]     int32  pooled-row index = rank of the label's first appearance in raster patch order
 *                      (dict insertion order, sppp.py:123-126)
 * num_slots[B]  int32  R of each image (may exceed r_cap: rows >= r_cap are then not written)
 * counts[B,r_cap] int32, slot_label[B,r_cap] int64, offsets[B,r_cap+1] int32, order[B,P] int32:
 *                      CSR of patches per slot, ascending patch id inside a slot (== the dict's lists)
 * All integer outputs are bit-exact with the reference.
 * ---------------------------------------------------------------------------------------------- */
int favit_sppp_assign(const int64_t* labels, int B, int img_h, int img_w, int patch, int grid,
                      int64_t* dom, int32_t* slot, int32_t* num_slots, int32_t* counts,
                      int64_t* slot_label, int32_t* offsets, int32_t* order, int r_cap,
                      favit_stream stream);

/* Superpixel centroids for the dynamic positional encoding — replaces models/sppp_mhla.py:226-262 (a Python loop over
 * images and labels with a device sync per label).  labels int64 [B,H,W]; centroids fp32 [B,K,2] = (mean x / W,
 * mean y / H) of the pixels carrying label k in [0, K), (0.5, 0.5) for a label without pixels; labels outside [0, K)
 * are ignored.  acc: workspace of B*3*K unsigned 64-bit integers (zeroed by the call): the sums are exact integers. */
int favit_sppp_centroids(const int64_t* labels, int B, int img_h, int img_w, int K, unsigned long long* acc,
                         float* centroids, favit_stream stream);

/* favit_sppp_assign and favit_sppp_centroids in ONE pass over the label map (the largest tensor of the SPPP front end:
 * 103 MB per 256-image batch at 224 px): the dominant-label kernel also accumulates the per-label pixel counts and
 * coordinate sums.  Used when the patches tile the whole image (img_h == img_w == grid * patch, patch 8 / 16 / 32);
 * otherwise the two kernels run one after the other.  Outputs as in the two separate calls. */
int favit_sppp_assign_centroids(const int64_t* labels, int B, int img_h, int img_w, int patch, int grid,
                                int64_t* dom, int32_t* slot, int32_t* num_slots, int32_t* counts,
                                int64_t* slot_label, int32_t* offsets, int32_t* order, int r_cap, int K,
                                unsigned long long* acc, float* centroids, favit_stream stream);

/* SuperpixelPooling.pool('mean') — replaces models/sppp.py:192-223 + the per-image loop/stack at
 * models/sppp_mhla.py:286-300.  x[B,P,D] (x_dtype) -> out[B,R,D] (out_dtype; the reference always
 * produces fp32, sppp.py:198).  Rows r >= num_slots[b] are zero-filled. */
int favit_sppp_pool_fwd(const void* x, favit_dtype x_dtype, const int32_t* order, const int32_t* offsets,
                        const int32_t* num_slots, void* out, favit_dtype out_dtype,
                        int B, int P, int R, int D, int r_cap, favit_stream stream);

/* Backward: dx[b,p,:] = dout[b,slot[b,p],:] / counts[b,slot[b,p]]. */
int favit_sppp_pool_bwd(const void* dout, favit_dtype dout_dtype, const int32_t* slot, const int32_t* counts,
                        void* dx, favit_dtype dx_dtype, int B, int P, int R, int D, int r_cap,
                        favit_stream stream);

/* Patch embedding + 'mean' pooling fused algebraically — replaces the P-row projection of
 * models/vit.py:36-41 followed by the per-superpixel mean of models/sppp.py:209-210 (models/sppp_mhla.py:281-300):
 *   out[b, r, (p1 p2 c)] = mean over the patches p of slot r of image[b, c, i(p)*patch + p1, j(p)*patch + p2]
 * so that  pooled[b, r, :] = out[b, r, :] . W^T + bias  (mean of a linear map = linear map of the mean): the projection
 * runs on R rows per image instead of P, and neither the [B,P,D] embeddings nor their gradients exist.
 * image: fp32 [B, C, img_h, img_w] contiguous (no gradient flows to it); order / offsets / num_slots: the CSR written by
 * favit_sppp_assign; out: [B, R, patch*patch*C] fp32 or bf16, rows of slots >= num_slots[b] are zero. */
int favit_sppp_pool_pixels(const float* image, int B, int C, int img_h, int img_w, int patch, int grid,
                           const int32_t* order, const int32_t* offsets, const int32_t* num_slots, void* out,
                           favit_dtype out_dtype, int R, int r_cap, favit_stream stream);

/* Class-token concat + centroid-based dynamic positional encoding in one pass — replaces models/sppp_mhla.py:302-310 and
 * models/sppp.py:271-299 (cat, a (0.5, 0.5) centroid row for the class token, pe = cat(sin(cx f), cos(cy f)),
 * f_i = exp(-i ln(10000) / (D/2)), add).  pooled fp32 [B,R,D], cls_token fp32 [D], centroids fp32 [B,R,2] (x, y);
 * out fp32 [B,R+1,D].  Gradients are slices of the output gradient (host side). */
int favit_sppp_embed_tokens(const float* pooled, const float* cls_token, const float* centroids, float* out, int B, int R,
                            int D, favit_stream stream);

/* The patch rearrangement of PatchEmbedding (models/vit.py:38-39: einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)') fused
 * with the cast to the GEMM operand type: image fp32 [B,C,S,S] -> out [B*(S/patch)^2, patch*patch*C] fp32 or bf16, the A
 * operand of the projection (favit_linear_fwd).  Square images, S % patch == 0, patch % 4 == 0. */
int favit_patchify(const float* image, void* out, favit_dtype out_dtype, int B, int C, int S, int patch,
                   favit_stream stream);

/* ------------------------------------------------------------------------------------------------
 * GPU superpixel segmentation — replaces the per-image skimage.segmentation.slic call of
 * models/sppp.py:26-74 (device -> host copy, CPU SLIC, host -> device copy) for the whole batch on the device:
 * Gaussian pre-smoothing (sigma), gy x gx centres on a regular grid (favit_slic_grid: gx = round(sqrt(n_segments W / H)),
 * gy = round(n_segments / gx)), `iters` assignment / update rounds with the SLIC distance
 *   sum_c (f_c - mu_c)^2 / compactness^2 + ((y - cy)^2 + (x - cx)^2) / step^2
 * over the 3 x 3 grid neighbourhood, centre sums in 64-bit fixed point (deterministic).  labels: int64 [B,H,W] in [0, gy gx).
 * No Lab conversion and no connectivity enforcement (scikit-image is not installed here: nothing to pin against);
 * oracle/slic_oracle.py reproduces this function bit for bit.
 * Workspaces (caller-allocated): feat, tmp fp32 [B,C,H,W]; centres fp32 [B, gy gx, 2 + C]; sums int64 [B, gy gx, 3 + C].
 * ---------------------------------------------------------------------------------------------- */
int favit_slic_grid(int H, int W, int n_segments, int* gy, int* gx);
int favit_slic_segment(const float* image, int B, int C, int H, int W, int n_segments, float compactness, float sigma,
                       int iters, int64_t* labels, float* feat, float* tmp, float* centres, long long* sums,
                       favit_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-tensor AdamW — the optimizer.step() of the reference's training loop,
 * experiments/mhla_pretrained.py:320-327,367 (three parameter groups, latent_proj at 5x lr) / main.py:129-132.
 * `count` fp32 tensors, HOST arrays of device pointers / sizes / per-tensor lr and weight decay (a parameter group is
 * just a run of tensors with the same lr / wd); `step` is a DEVICE int64 holding the 1-based step count of this update
 * (the caller advances it in-stream, so a captured CUDA graph steps correctly); grads are multiplied by grad_scale
 * first (1/world when the gradient all-reduce sums).  Same update rule as torch.optim.AdamW (decoupled decay).
 * ---------------------------------------------------------------------------------------------- */
int favit_adamw_multi(int count, void* const* params, const void* const* grads, void* const* exp_avg,
                      void* const* exp_avg_sq, const int64_t* numel, const float* lr, const float* weight_decay,
                      const int64_t* step, double beta1, double beta2, float eps, float grad_scale, favit_stream stream);

/* The non-default SuperpixelPooling variants (models/sppp.py:178-184, 211-216), same CSR inputs as the mean pool.
 * 'max'      : out[b,r,c] = max over the slot's patches; argmax int32 [B,R,D] (patch id, -1 for an empty row) is saved
 *              for the backward, which routes dout[b,r,c] to dx[b,argmax,c] (dx is zeroed by the call).
 * 'attention': weights fp32 [B,P] = softmax over the slot's patches of sum_c x[b,p,c] (0 for patches outside every
 *              kept slot); out[b,r,:] = sum_p weights[b,p] x[b,p,:].  Backward (dx has x's dtype, zeroed by the call):
 *              dx[p,:] = w_p dout_r + w_p (dout_r . x_p - dout_r . out_r).
 * out / dout are fp32 (the reference's `torch.zeros` default dtype, sppp.py:198). */
int favit_sppp_pool_max_fwd(const void* x, favit_dtype x_dtype, const int32_t* order, const int32_t* offsets,
                            const int32_t* num_slots, float* out, int32_t* argmax, int B, int P, int R, int D,
                            int r_cap, favit_stream stream);
int favit_sppp_pool_max_bwd(const float* dout, const int32_t* argmax, void* dx, favit_dtype dx_dtype, int B, int P,
                            int R, int D, favit_stream stream);
int favit_sppp_pool_attn_fwd(const void* x, favit_dtype x_dtype, const int32_t* order, const int32_t* offsets,
                             const int32_t* num_slots, float* out, float* weights, int B, int P, int R, int D,
                             int r_cap, favit_stream stream);
int favit_sppp_pool_attn_bwd(const void* x, favit_dtype x_dtype, const float* dout, const float* out,
                             const float* weights, const int32_t* order, const int32_t* offsets,
                             const int32_t* num_slots, void* dx, int B, int P, int R, int D, int r_cap,
                             favit_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* FAVIT_H_ */
