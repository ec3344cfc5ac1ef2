/* C ABI of libunetzoo_b200.so -- the B200 (sm_100a) hot path of gigantenbein/UNet-Zoo.
 *
 * The reference is pure Python/PyTorch: its "FFI" for this path is the set of ATen/cuDNN operator calls made by
 * torchlayers.py, models/{unet,probabilistic_unet,phiseg}.py and utils.py.  Every entry
 * point below replaces one (or a fused group) of those call sites; the citation after each prototype names them.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked [host]; nothing is retained or freed.
 *   - activations: bf16, NHWC, channel count a multiple of 16 (zero padded), pixel stride `ld*` in elements
 *     (multiple of 8) so channel slices of concat buffers can be read / written in place.
 *   - everything the reference's callers can observe (mu, sigma, z, logits, losses, parameter gradients): fp32 in the
 *     reference's NCHW / OIHW layouts.
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it.
 *   - return 0 on success; non-zero = error, text via uz_last_error().  Never throws, never synchronises.
 */
#ifndef UNETZOO_B200_H_
#define UNETZOO_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define UZ_ABI_VERSION 1

const char* uz_last_error(void);
int uz_abi_version(void);
/* storage type of activations / packed weights this library was built for: 0 = bf16 (libunetzoo_b200.so), 1 = IEEE half
 * (libunetzoo_b200_fp16.so, -DUZ_ACT_FP16: 10-bit mantissa like TF32, the tolerance-matched parity mode) */
int uz_storage_dtype(void);
int uz_device_sm_count(void);
/* number of kernel launches issued by this library so far in this process (bench.py's gpu_launches) */
long long uz_launch_count(void);
/* profiling knobs for uz_conv_fwd (results become invalid): 1 = skip epilogue body, 2 = skip MMA issue,
 * 4 = skip activation TMA loads, 8 = skip weight TMA loads, 32 = force the generic (non-persistent) kernel,
 * 512 / 1024 / 2048 / 4096 = dispatch and elision knobs of tools/family_times.py, 8192 = one CTA per SM also for narrow
 * layers, 16384 = keep output chunks wider than 128 (A/B switches of two plan decisions, see conv_tc2.cu),
 * 128 = uz_conv_fwd returns without launching, 256 = uz_conv_wgrad returns without launching (bench.py times the step
 * with and without a kernel family to get that family's in-situ time).  0 restores normal operation. */
int uz_set_debug_flags(int flags);
/* profiling build only (libunetzoo_b200_prof.so, -DUZ_PROFILE_KNOBS): device buffer of 16 uint64 that CTA 0 of the
 * small-shape tensor-core kernels fills with %globaltimer phase timestamps (tools/phase_trace.py); NULL switches it off */
int uz_set_trace_buffer(void* device_ptr);
/* Programmatic dependent launch: 0 off, 1 every kernel, 2 only the light (elementwise / reduction) kernels, 3 only the
 * tensor-core kernels (default; environment UZ_PDL): their barrier / TMEM / descriptor prologue then overlaps the tail of
 * the preceding kernel.  Measured on the PHiSeg-7/5 step: off 4.32 ms, 1: 4.27, 2: 4.48, 3: 4.23 (profiles/README.md). */
int uz_set_pdl(int enabled);
int uz_get_pdl(void);
/* Preferred shared-memory carveout (percent of the unified L1 / shared-memory array; -1 = driver default) applied to
 * every kernel of the library at its next launch.  Environment: UZ_CARVEOUT. */
int uz_set_smem_carveout(int percent);

/* ---- convolutions on tcgen05 tensor cores (conv_tc.cu, wgrad_tc.cu) ------------------------------------------------ */

/* Tile geometry the conv kernel will use for an [N,H,W] pixel grid: a TN x TH x TW box of 128 pixels per CTA.
 * num_tiles sizes the `stats_partial` buffer of uz_conv_fwd.  [host out-params] */
int uz_conv_tile_geometry(int N, int H, int W, int* TW, int* TH, int* TN, int* num_tiles);

/* 1 if uz_conv_fwd will use the persistent 16x16-tile kernel for this shape, 0 for the generic kernel. */
int uz_conv_uses_persistent_kernel(int N, int H, int W, int Cin, int Cout, int taps);

/* y[n,h,w,co] = act( scale[co] * sum_{tap,ci} x[n,h+dy,w+dx,ci] * w_packed[tap][co][ci] + shift[co] ), zero padding.
 * taps = 9 (3x3, pad 1) or 1 (1x1).  scale/shift may be NULL (1 / 0).  relu != 0 applies max(.,0).
 * stats_partial (optional): [2][Cout] fp32 accumulators, ZERO on entry; the kernel atomically adds the per-channel sum
 * and sum of squares of the STORED bf16 outputs (training-mode BatchNorm statistics; consumed by uz_bn_apply_train or
 * uz_bn_finalize with tiles = 1).  Summation order across CTAs is not fixed (fp32 atomics).
 * w_packed taps are dx-major (t = kw*3 + kh) as produced by uz_pack_conv_weight.
 * Replaces: nn.Conv2d forward (torchlayers.py:18; models/unet.py:25-29; models/phiseg.py:28,32,57-58,91), fused with
 * the eval-mode BatchNorm + ReLU of torchlayers.py:20-21 or the bias + ReLU of models/unet.py:25-30; called with
 * dgrad-packed weights it is conv2d's input-gradient (autograd of the same call sites). */
int uz_conv_fwd(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps,
                void* y, int ldy, const float* scale, const float* shift, int relu, float* stats_partial,
                void* stream);

/* Optional epilogue extensions of uz_conv_fwd_ex (all fields zero = plain uz_conv_fwd).
 *  stats_rows   0: stats_partial is [2][Cout], zero on entry, fp32 atomics (summation order across CTAs not fixed);
 *               1: stats_partial is [uz_conv_stats_rows(...)][2][Cout]; every row is written (not added to) by exactly one
 *                  CTA and uz_bn_finalize reduces the rows in fixed order -> bit-reproducible training statistics.
 *  bn_*         fused BatchNorm/ReLU backward of the layer that PRODUCED this conv's forward input (only meaningful when
 *               the call is an input-gradient, i.e. w_packed is the dgrad packing): with y = that layer's stored
 *               pre-normalisation output [pixels][Cout] (bn_ldy), a = relu(y*bn_scale + bn_shift) its activation, the
 *               epilogue stores g = out * [a > 0] (bn_relu) and accumulates sum(g), sum(g*y) per channel into
 *               bn_sums [2][Cout] (zero on entry) -- the reductions uz_bn_bwd_reduce_sums would compute in a separate
 *               pass over g and y (reference: autograd of nn.BatchNorm2d + nn.ReLU, torchlayers.py:20-21).
 *  residual     out = residual + res_sign * value (res_sign = +1 / -1): the additive coupling of a reversible block and
 *               its inverse, and the gradient accumulation dx1 = dy1 + dF/dx (torchlayers.py:67-82 via revtorch). */
typedef struct UzConvExtra {
  int stats_rows;
  const void* bn_y;
  int bn_ldy;
  const float* bn_scale;
  const float* bn_shift;
  int bn_relu;
  float* bn_sums;
  const void* residual;
  int ld_res;
  int res_sign;
} UzConvExtra;
int uz_conv_fwd_ex(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps,
                   void* y, int ldy, const float* scale, const float* shift, int relu, float* stats_partial,
                   const UzConvExtra* extra, void* stream);
/* number of statistics rows uz_conv_fwd_ex writes for this shape when extra->stats_rows = 1 */
int uz_conv_stats_rows(int N, int H, int W, int Cin, int Cout, int taps);

/* Share of the SMs (percent, 5..100, default 100) a 2-D uz_conv_wgrad launch is planned for: callers that overlap
 * weight gradients with other work (auxiliary streams) ask for fewer pixel splits -> fewer partial slabs to reduce.
 * Affects uz_wgrad_workspace_floats too: set it before sizing the workspace. */
int uz_set_wgrad_sm_percent(int percent);
/* Workspace size (floats) for uz_conv_wgrad, or -1 if the shape is unsupported. */
long long uz_wgrad_workspace_floats(int N, int H, int W, int Cin, int Cout, int taps);

/* dw[co][ci][tap] = sum_pixels dy[p][co] * x[p + offset(tap)][ci]  (fp32, PyTorch OIHW layout, logical channel
 * counts; Cin/Cout are the stored, padded counts).  Replaces conv2d's weight-gradient (autograd of torchlayers.py:18). */
int uz_conv_wgrad(const void* x, int ldx, const void* dy, int lddy, int N, int H, int W, int Cin, int Cout, int taps,
                  int Cin_logical, int Cout_logical, float* workspace, float* dw, void* stream);

/* ---- memory-bound NHWC kernels (elementwise.cu) -------------------------------------------------------------------- */

/* fp32 OIHW weights -> bf16 [taps][CoutP][CinP] (forward) and, if w_dgrad != NULL, bf16 [taps][CinP2][CoutP2] with
 * taps flipped and channels transposed (input-gradient).  Padding is zero filled.  (fp32 parameters stay
 * owned by the caller's stock Adam, train_model.py:49.) */
int uz_pack_conv_weight(const float* w, int Cout, int Cin, int taps, void* w_fwd, int CoutP, int CinP, void* w_dgrad,
                        int CinP2, int CoutP2, void* stream);

/* The same packing for every conv layer of a model in one launch.  descs_device: device array of n UzPackDesc (dgrad
 * copy, if requested, has dims [taps][CinP][CoutP]).  grid = (blocks_per_layer, n). */
typedef struct UzPackDesc {
  const void* w;      /* fp32 OIHW */
  void* w_fwd;        /* bf16 [taps][CoutP][CinP] */
  void* w_dgrad;      /* bf16 [taps][CinP][CoutP] or NULL */
  int Cout, Cin, taps, CoutP, CinP, reserved;
} UzPackDesc;
int uz_pack_conv_weights_batched(const void* descs_device, int n, int blocks_per_layer, void* stream);

/* Eval-mode BatchNorm fold (uz_bn_eval_fold) for every layer of a model in ONE launch: scale = gamma / sqrt(rv + eps),
 * shift = beta + (conv_bias - rm) * scale.  descs_device: device array of n UzFoldDesc. */
typedef struct UzFoldDesc {
  const float* conv_bias;
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float* scale;
  float* shift;
  long long C;
} UzFoldDesc;
int uz_bn_eval_fold_batched(const void* descs_device, int n, int max_channels, float eps, void* stream);

/* torch.optim.Adam(lr, betas, eps, weight_decay) -- the reference's optimizer (train_model.py:49: lr 1e-3,
 * weight_decay 1e-5 added to the gradient) -- for ALL parameters in one launch.  descs_device: device array of UzAdamDesc;
 * chunk_table_device: int [nchunks][2] = (tensor index, chunk index), one block per uz_adam_chunk_elems() elements;
 * every parameter has its own fp32 step counter on the device (torch's capturable layout), incremented before the update
 * (graph capturable; parameters without a gradient are simply not in the table and keep their count). */
typedef struct UzAdamDesc {
  void* p;            /* fp32 parameter, updated in place */
  const void* g;      /* fp32 gradient */
  void* m;            /* exp_avg */
  void* v;            /* exp_avg_sq */
  float* step;        /* this parameter's step counter (device scalar) */
  long long n;        /* elements */
} UzAdamDesc;
int uz_adam_chunk_elems(void);
int uz_adam_step_batched(const void* descs_device, int ntensors, const int* chunk_table_device, int nchunks, double lr,
                         double beta1, double beta2, double eps, double weight_decay, void* stream);

/* Adam and the bf16 weight packing in ONE pass for the conv weights: the update of a 32 x 32 x 9-tap tile stays in shared
 * memory and both packed copies (uz_pack_conv_weight layouts) are written from there -- the separate packing pass at the
 * head of the next step (100 MB read + 100 MB written, alone on the GPU) disappears.  packs_device: UzAdamPackDesc rows
 * (tensor = row of descs_device); item_table_device: int [nitems][2] = (pack row, tile), uz_adam_pack_items tiles per
 * layer.  Tensors in the chunk table get the plain update; the step counters of all ntensors rows are incremented. */
typedef struct UzAdamPackDesc {
  void* w_fwd;        /* bf16 [taps][CoutP][CinP] */
  void* w_dgrad;      /* bf16 [taps][CinP][CoutP] or NULL */
  int tensor;         /* row of the UzAdamDesc table */
  int Cout, Cin, taps, CoutP, CinP;
  const float* slab;  /* NULL: gradient = the g of the UzAdamDesc row; else the split-K slabs [splits][taps][CoutP][CinP] of
                         uz_conv_wgrad_partial, summed over the splits in order while the tile is loaded (no separate
                         uz_wgrad_reduce_batched pass, no OIHW gradient tensor) */
  int splits, reserved;
} UzAdamPackDesc;
int uz_adam_pack_items(int CoutP, int CinP, int taps);
int uz_adam_pack_step(const void* descs_device, int ntensors, const int* chunk_table_device, int nchunks,
                      const void* packs_device, const int* item_table_device, int nitems, double lr, double beta1,
                      double beta2, double eps, double weight_decay, void* stream);


/* Training-mode BatchNorm statistics: reduce the conv's per-tile partials, emit scale = gamma*invstd and
 * shift = beta - mean*scale, save mean / invstd, update running stats (momentum, unbiased variance).
 * Replaces nn.BatchNorm2d(eps=1e-3, momentum=0.01) statistics, torchlayers.py:20. */
int uz_bn_finalize(const float* partial, int tiles, int C, float count, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                   float* mean_out, float* invstd_out, void* stream);

/* Training-mode BatchNorm normalise (+ReLU) straight from the conv's [2][C] accumulators: derives scale/shift per block,
 * publishes scale/shift/mean/invstd (saved for backward) and updates the running statistics -- finalize + apply in one
 * launch.  Replaces nn.BatchNorm2d(eps=1e-3, momentum=0.01) + nn.ReLU forward, torchlayers.py:20-21. */
int uz_bn_apply_train(const void* y, int ldy, const float* sums, float count, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, float* scale_out,
                      float* shift_out, float* mean_out, float* invstd_out, int relu, void* out, int ldo,
                      long long npix, int C, void* stream);
/* uz_bn_apply_train with (a) an additive-coupling residual: out = residual + res_sign * act(...)  (reversible blocks,
 * torchlayers.py:67-75 via revtorch: y1 = x1 + F(x2) written straight into the block output) and (b) stat_updates
 * momentum updates of the running statistics in one go (2 for reversible blocks, whose F and G run a second time in
 * backward on the same batch: SURVEY.md quirk Q7). */
int uz_bn_apply_train_ex(const void* y, int ldy, const float* sums, float count, const float* gamma, const float* beta,
                         float eps, float momentum, float* running_mean, float* running_var, float* scale_out,
                         float* shift_out, float* mean_out, float* invstd_out, int relu, void* out, int ldo,
                         long long npix, int C, const void* residual, int ld_res, int res_sign, int stat_updates,
                         void* stream);
/* BatchNorm(+ReLU) backward in two launches: accumulate sum(g), sum(g*y) into [2][C] (zero on entry, fp32 atomics), then
 * dy = A*g + B*y + Cc with the coefficients derived per block; dgamma / dbeta are written by the second kernel. */
int uz_bn_bwd_reduce_sums(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift,
                          int relu, long long npix, int C, float* sums, void* stream);
/* uz_bn_bwd_reduce_sums that also inverts a reversible block's coupling in the same pass over the recomputed y:
 * inv_out = inv_in - act(y*scale + shift)   (x2 = y2 - G(y1), x1 = y1 - F(x2); revtorch backward_pass). */
int uz_bn_bwd_reduce_sums_ex(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift,
                             int relu, long long npix, int C, float* sums, const void* inv_in, int ld_inv_in,
                             void* inv_out, int ld_inv_out, void* stream);
int uz_bn_bwd_apply_train(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift,
                          int relu, const float* sums, float count, const float* gamma, const float* mean,
                          const float* invstd, float* dgamma, float* dbeta, void* dy, int lddy, long long npix, int C,
                          void* stream);
/* The same backward (reference torchlayers.py:18-21 through nn.BatchNorm2d / nn.ReLU) as ONE launch on thread-block
 * clusters: a cluster owns 16 channels, its CTAs split the pixels, stage (g, y) in shared memory, exchange the partial
 * sums through distributed shared memory in a fixed order (deterministic, no atomics) and write dy from the staged copy.
 * uz_bn_bwd_fused_supported: 1 if this size has a plan (npix <= 49152, C a multiple of 16). */
int uz_bn_bwd_fused_supported(long long npix, int C);
int uz_bn_bwd_fused(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift, int relu,
                    float count, const float* gamma, const float* mean, const float* invstd, float* dgamma,
                    float* dbeta, void* dy, int lddy, long long npix, int C, void* stream);

/* Conv2D of the reference in training mode as ONE launch (torchlayers.py:18-21: conv + bias -> BatchNorm2d with batch
 * statistics -> ReLU) for maps whose pixel tiles fit one thread-block cluster (N*H*W <= 1024: the 2x2 ... 8x8 levels at
 * batch 12): y = conv(x) + bias (bf16, kept for backward); the per-channel sums of the stored y are exchanged between the
 * tile CTAs through distributed shared memory in a fixed order (deterministic, no atomics, no second launch);
 * a = act(y*scale + shift) is written from the accumulators still in TMEM.  scale / shift / mean / invstd [Cout]: the
 * coefficients backward needs; running statistics get stat_updates momentum updates (unbiased variance). */
int uz_conv_bn_fused_supported(int N, int H, int W, int Cin, int Cout, int taps);
int uz_conv_bn_act_fused(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps,
                         const float* bias, const float* gamma, const float* beta, float eps, float momentum,
                         float* running_mean, float* running_var, int stat_updates, int relu, void* y, int ldy, void* a,
                         int lda, float* scale_out, float* shift_out, float* mean_out, float* invstd_out, void* stream);

/* The same backward for LARGE maps as one COOPERATIVE launch: a single-wave grid accumulates the sums (fp32 atomics into
 * sums[2][C], zero on entry), crosses a grid-wide barrier and writes dy, re-reading dout / y from L2. */
int uz_bn_bwd_coop(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift, int relu,
                   float* sums, float count, const float* gamma, const float* mean, const float* invstd, float* dgamma,
                   float* dbeta, void* dy, int lddy, long long npix, int C, void* stream);

/* Eval-mode fold of conv bias + BatchNorm running stats into the conv epilogue's scale / shift (train_model.py:139). */
int uz_bn_eval_fold(const float* conv_bias, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, float eps, int C, float* scale, float* shift, void* stream);

/* out = act(y * scale[c] + shift[c]) -- BatchNorm normalise + ReLU (torchlayers.py:20-21). */
int uz_affine_act(const void* y, int ldy, const float* scale, const float* shift, int relu, void* out, int ldo,
                  long long npix, int C, void* stream);

/* BatchNorm(+ReLU) backward in two passes over (dout, y): per-block partial sums -> coefficients -> dy.
 * partial: [uz_bn_bwd_num_blocks][2][C].  dy = A*g + B*y + Cc with g = dout * [y*scale+shift > 0].
 * Replaces autograd of torchlayers.py:20-21. */
int uz_bn_bwd_num_blocks(long long npix, int C);
int uz_bn_bwd_reduce(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift, int relu,
                     long long npix, int C, float* partial, void* stream);
int uz_bn_bwd_finalize(const float* partial, int nblocks, int C, float count, const float* gamma, const float* mean,
                       const float* invstd, float* coefA, float* coefB, float* coefC, float* dgamma, float* dbeta,
                       void* stream);
int uz_bn_bwd_apply(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift, int relu,
                    const float* coefA, const float* coefB, const float* coefC, void* dy, int lddy, long long npix,
                    int C, void* stream);

/* AvgPool2d(kernel 2, stride 2, ceil_mode) on even sizes (models/phiseg.py:23, models/unet.py:22,
 * models/probabilistic_unet.py:56) and its gradient (optionally accumulated into dx). */
int uz_avgpool2_fwd(const void* x, int ldx, void* out, int ldo, int N, int Ho, int Wo, int C, void* stream);
int uz_avgpool2_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int Ho, int Wo, int C, int accumulate,
                    void* stream);

/* Bilinear x2 upsampling, align_corners != 0 (models/phiseg.py:66,213-216,305-309) or == 0 (models/unet.py:67), written
 * at pixel stride ldo so the result can land inside a concat buffer (torch.cat of models/phiseg.py:71,315); gradient. */
int uz_upsample2x_fwd(const void* x, int ldx, void* out, int ldo, int N, int h, int w, int C, int align_corners,
                      void* stream);
int uz_upsample2x_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int h, int w, int C, int align_corners,
                      void* stream);

/* Strided channel-slice copy (accumulate 0), dst += src (1) or dst -= src (2): torch.cat along channels and its slice
 * gradients (models/phiseg.py:71,183,315; models/unet.py:72) and the residual add / inverse of the reversible blocks
 * (revtorch ReversibleBlock: y1 = x1 + F(x2), x2 = y2 - G(y1); torchlayers.py:71-78). */
int uz_copy_channels(const void* src, int lds, void* dst, int ldd, long long npix, int C, int accumulate, void* stream);
/* dst[p] = src[p % src_npix]: one image's channel slice replicated over a batch of copies (the skip connections of the
 * N-sample evaluation, whose encoders run once: train_model.py:177-179 feeds N identical copies). */
int uz_copy_channels_bcast(const void* src, int lds, long long src_npix, void* dst, int ldd, long long npix, int C,
                           void* stream);
/* out = a + b (sign >= 0) or a - b (sign < 0) on channel slices in one pass: the coupling y1 = x1 + F(x2) and its inverse
 * x2 = y2 - G(y1) of the reversible blocks (torchlayers.py:71-78); out may alias a or b. */
int uz_add_channels(const void* a, int lda, const void* b, int ldb, void* out, int ldo, long long npix, int C, int sign,
                    void* stream);

/* Global spatial mean [B,hw,C] -> [B,C] (bf16, fp32 accumulation) and its gradient: torch.mean over H then W in front of
 * the ProbUNet Gaussian head (models/probabilistic_unet.py:114-115). */
int uz_global_mean_fwd(const void* x, int ldx, int B, int hw, int C, void* out, int ldo, void* stream);
int uz_global_mean_bwd(const void* dout, int ldd, int B, int hw, int C, void* dx, int ldx, void* stream);

/* Network input: fp32 NCHW patch (+ optional mask of integer labels as float, [B,1,H,W]) -> bf16 NHWC [B,H,W,CP] with
 * channels [image | (mask==k)-0.5, k<nlabels | 0...].  Replaces utils.convert_batch_to_onehot (utils.py:289-311, a
 * host loop + H2D in the reference) and the cat of models/phiseg.py:176-183 / models/probabilistic_unet.py:103-109. */
int uz_input_pack(const float* patch, const float* mask, int B, int Cimg, int H, int W, int nlabels, void* out, int CP,
                  void* stream);

/* Module-boundary layout conversions fp32 NCHW <-> bf16 NHWC (ld >= C, padding zero filled). */
int uz_nchw_to_nhwc(const float* src, int B, int C, long long hw, void* dst, int ld, void* stream);
int uz_nhwc_to_nchw(const void* src, int ld, int B, int C, long long hw, float* dst, void* stream);

/* ---- latent heads and losses (heads_loss.cu) ------------------------------------------------------------------------ */

/* SampleZBlock head (models/phiseg.py:95-106): mu = 1x1 conv, sigma = softplus(1x1 conv), z = mu + sigma*eps.
 * feat bf16 NHWC [B*hw][C]; weights fp32 [zdim][C]; eps/mu/sigma/z fp32 NCHW [B,zdim,hw]. */
int uz_head_fwd(const void* feat, int ld, int C, const float* wmu, const float* bmu, const float* wsig,
                const float* bsig, const float* eps, int B, int hw, int zdim, float* mu, float* sigma, float* z,
                void* stream);
int uz_head_bwd_num_blocks(int B, int hw);
/* Gradient of the head.  dmu/dsigma/dz may be NULL.  wpartial: [blocks][2*zdim][C], bpartial: [blocks][2*zdim];
 * dw: [2*zdim][C] (mu rows then sigma rows), db: [2*zdim]; dfeat bf16 NHWC. */
int uz_head_bwd(const void* feat, int ld, int C, const float* wmu, const float* wsig, const float* eps,
                const float* sigma, const float* dmu, const float* dsigma, const float* dz, int B, int hw, int zdim,
                void* dfeat, int ldd, float* wpartial, float* bpartial, float* dw, float* db, void* stream);

/* One level of KL_two_gauss_with_diag_cov (models/phiseg.py:436-453, sigma1*sigma0 quirk) times `weight`
 * (4^level, models/phiseg.py:463): out[0] = weight * mean_b 0.5 * sum(...).  Inputs fp32 [batch][per_sample]. */
int uz_kl_num_blocks(int batch, int per_sample);
int uz_kl_fwd(const float* mu0, const float* s0, const float* mu1, const float* s1, int batch, int per_sample,
              float weight, float* out, double* partial /* [uz_kl_num_blocks] */, void* stream);
int uz_kl_bwd(const float* mu0, const float* s0, const float* mu1, const float* s1, int batch, int per_sample,
              float weight, const float* upstream, float* dmu0, float* ds0, float* dmu1, float* ds1, void* stream);

/* The whole hierarchical KL term of PHiSeg (models/phiseg.py:463-472: one KL_two_gauss_with_diag_cov per latent level and
 * the running `loss_tot += kl_weight * level`) in two launches instead of four per level: levels_out[l] = level_weight[l] *
 * KL_l, total_out[0] = sum over l = L-1 ... 0 of total_weight * levels_out[l] (fp32, the reference's order).  Arrays of L
 * host pointers / element counts; partial: L * uz_kl_hierarchy_num_blocks(max numel) doubles.  Backward: 4 * L gradient
 * pointers ([l*4 + 0..3] = d mu0, d sigma0, d mu1, d sigma1), upstream = d loss / d total_out (device scalar). */
int uz_kl_hierarchy_num_blocks(long long max_numel);
int uz_kl_hierarchy_fwd(const float* const* mu0, const float* const* s0, const float* const* mu1, const float* const* s1,
                        const long long* numel, const float* level_weight, int L, int batch, float total_weight,
                        double* partial, float* levels_out, float* total_out, void* stream);
int uz_kl_hierarchy_bwd(const float* const* mu0, const float* const* s0, const float* const* mu1, const float* const* s1,
                        const long long* numel, const float* level_weight, int L, int batch, float total_weight,
                        const float* upstream, float* const* grads, void* stream);


/* Likelihood.s_layer (1x1 conv to n_classes, no norm / activation) fused with the nearest upsample to full resolution
 * (models/phiseg.py:283-284,319-321): out fp32 NCHW [B,ncls,h*factor,wd*factor]. */
int uz_slayer_fwd(const void* feat, int ld, int C, const float* w, const float* bias, int ncls, int B, int h, int wd,
                  int factor, float* out, void* stream);
int uz_slayer_bwd_num_blocks(int B, int h, int wd);
int uz_slayer_bwd(const float* dout, const void* feat, int ld, int C, const float* w, int ncls, int B, int h, int wd,
                  int factor, void* dfeat, int ldd, float* wpartial, float* bpartial, float* dw, float* db, void* stream);

/* residual_multinoulli_loss (models/phiseg.py:481-513): for l = L-1..0, acc_l = sum_{k>=l} s_k and
 * ce_levels[l] = mean_b sum_pixels CE(acc_l, target).  s / ds: [host] arrays of L device pointers (fp32 NCHW);
 * ds[l] (optional) receives upstream[0] * d(sum_l ce_levels[l]) / d s_l (upstream: device scalar, NULL = 1).
 * partial: [uz_residual_ce_num_blocks][L]. */
int uz_residual_ce_num_blocks(int B, int hw);
int uz_residual_ce(const float* const* s, float* const* ds, const float* upstream, int L, int ncls,
                   const float* target, int B, int hw, float* partial, float* ce_levels, void* stream);

/* PHISeg.accumulate_output (models/phiseg.py:428-434): out = s[L-1] + s[0] + ... + s[L-2] (+ softmax over classes);
 * out may alias s[L-1] (the reference accumulates in place).  s: [host] array of L device pointers. */
int uz_accumulate_output(const float* const* s, int L, int ncls, int B, int hw, int use_softmax, float* out,
                         void* stream);

/* ---- sample evaluation (eval_metrics.cu) ---------------------------------------------------------------------------- */

/* Bit-pack (labels == label_values[l]) masks: labels [count][hw] of dtype 0 = int64, 1 = float32, 2 = uint8;
 * bits uint32 [count][nlabels][ceil(hw/32)], counts int32 [count][nlabels].  label_values is a [host] array.
 * Replaces the (m == lbl)*1 / torch.sum tests of utils.py:158-165. */
int uz_ged_pack_masks(const void* labels, int dtype, int count, int hw, const int* label_values, int nlabels,
                      unsigned int* bits, int* counts, void* stream);
/* generalised_energy_distance (utils.py:148-200): pair_d double [N*M + N*N + M*M] (the d_sy, d_ss, d_yy lists in the
 * reference's order); out double [4] = {GED, sum d_sy, sum d_ss, sum d_yy}, summed sequentially like Python. */
int uz_ged_pairwise(const unsigned int* bits_s, const int* cnt_s, int N, const unsigned int* bits_y, const int* cnt_y,
                    int M, int nlabels, int hw, double* pair_d, double* out, void* stream);
/* torch.argmax(dim=1) of fp32 NCHW [N,C,hw] -> uint8 [N,hw] (train_model.py:195). */
int uz_argmax_classes(const float* x, int N, int C, int hw, unsigned char* out, void* stream);
/* variance_ncc_dist (utils.py:202-247 with ncc :130-145): probs fp32 [N,C,hw]; gt one-hot [M,C,hw] of gt_dtype
 * (0 = int64, 1 = float32, 2 = uint8); work double [(1+M)*hw + M]; out double [1]. */
int uz_variance_ncc(const float* probs, const void* gt, int gt_dtype, int N, int C, int hw, int M, double* work,
                    double* out, void* stream);

/* Fused N-sample evaluation tail (train_model.py:185-222 after the network), for I images of n samples each (batch index
 * b = sample * I + image): ONE pass over the L per-level class logits levels[l] = fp32 [n*I][C][(H/f_l)*(W/f_l)] (low
 * resolution, nearest upsampling by factors[l] is index arithmetic; f = 1 for full-resolution logits) computes
 * accumulate_output (phiseg.py:428-434, same fp32 order), softmax, argmax -> bit-packed label masks
 * bits uint32 [I][n][nlabels][ceil(HW/32)] + counts int32 [I][n][nlabels], and sums fp32 [I][2][C][HW] =
 * (sum_i p_ic, sum_i log(p_ic + 1e-8)) -- everything variance_ncc_dist (utils.py:202-247) needs, additive over samples
 * and therefore over GPUs (one all-reduce).  part: workspace fp32 [I][uz_eval_sample_groups][2][C][HW].
 * levels / factors / label_values are [host] arrays. */
int uz_eval_sample_groups(int n, int I, int hw);
int uz_eval_sample_stats(const float* const* levels, const int* factors, int L, int n, int I, int C, int H, int W,
                         const int* label_values, int nlabels, unsigned int* bits, int* counts, float* part,
                         float* sums, void* stream);
/* variance_ncc_dist of ONE image from its (all-reduced) sums [2][C][hw] over N samples and the M annotator label maps
 * gt [M][hw] (dtype 0 = int64, 1 = float32, 2 = uint8; class indices) -> out[0]; with dice_counts != NULL (int32 [3*C]
 * workspace) also the per-class Dice of argmax_c(mean probs) against annotator `dice_annotator` -> out[1..C]
 * (train_model.py:207-222 incl. the empty-set conventions; medpy dc).  work: double [(1+M)*hw + M]; out double [1+C]. */
int uz_ncc_dice_from_sums(const float* sums, const void* gt, int gt_dtype, int N, int C, int hw, int M,
                          int dice_annotator, double* work, int* dice_counts, double* out, void* stream);

/* Synthetic LIDC-shaped batch generated on the device (the data plug-in's device path; replaces the
 * per-image host work of data/batch_provider.py:43-67,131-137 for synthetic runs): patch fp32 [B,1,S,S], labels uint8
 * [B,S,S,annotators], mask fp32 [B,1,S,S] (one random annotator per image).  A pure function of `seed` (counter-based
 * hash RNG); unet-zoo_b200/b200/data.py holds the numpy restatement the tests compare with. */
int uz_synth_lidc_batch(unsigned int seed, int B, int size, int annotators, float* patch, unsigned char* labels,
                        float* mask, void* stream);

/* ---- volumes: the 3-D clones of models/phiseg3D.py (NDHWC bf16 activations) ------------------------------------------ */
/* BatchNorm3d, 1x1x1 heads, KL, residual cross-entropy, channel copies and layout conversion are the flat
 * [pixels][channels] entry points above with pixels = N*D*H*W (hw = D*H*W for the fp32 NCDHW sides). */

/* nn.Conv3d 3x3x3 pad 1 / 1x1x1 forward (models/phiseg3D.py:24) with the same epilogue contract as uz_conv_fwd
 * (scale/shift/relu fold, or BatchNorm statistics accumulators); with dgrad-packed weights it is the input gradient.
 * taps = 27: w_packed [(kd*3 + kw)*3 + kh][Cout][Cin] as produced by uz_pack_conv_weight(taps = 27).
 * Same persistent tcgen05 kernel as the 2-D path, the z taps are one more factor of the K loop; any D/H/W. */
int uz_conv3d_fwd(const void* x, int N, int D, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps,
                  void* y, int ldy, const float* scale, const float* shift, int relu, float* stats_partial,
                  void* stream);
/* uz_conv3d_fwd with the epilogue extensions of uz_conv_fwd_ex (fused BatchNorm-backward sums, residual; no statistics
 * rows) */
int uz_conv3d_fwd_ex(const void* x, int N, int D, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps,
                     void* y, int ldy, const float* scale, const float* shift, int relu, float* stats_partial,
                     const UzConvExtra* extra, void* stream);
/* conv3d weight gradient: dw fp32 [Cout_logical][Cin_logical][27] (OIDHW).  Workspace from uz_wgrad3d_workspace_floats. */
long long uz_wgrad3d_workspace_floats(int N, int D, int H, int W, int Cin, int Cout);
int uz_conv3d_wgrad(const void* x, int ldx, const void* dy, int lddy, int N, int D, int H, int W, int Cin, int Cout,
                    int Cin_logical, int Cout_logical, float* workspace, float* dw, void* stream);

/* Deferred split-K reduction: uz_conv_wgrad_partial runs the tensor-core kernel only (partial slabs
 * [splits][taps][Cout][Cin] fp32 stay in `workspace`, *splits reports their number; D == 0 images, D > 0 volumes with 27
 * taps), and ONE uz_wgrad_reduce_batched launch later sums the splits of many layers in fixed order and writes their
 * OIHW gradients dw[Cout][Cin][taps] (the per-layer reduction of uz_conv_wgrad costs a launch per layer and writes with a
 * stride of `taps` floats).  accumulate = 1: the split-K CTAs ADD their blocks into ONE slab [taps][Cout][Cin] (workspace
 * zero on entry, taps*Cout*Cin floats; bulk reduce-add stores, the reduction happens in L2; *splits = 1; fp32 summation
 * order then depends on scheduling -- accumulate = 0 keeps the fixed-order slabs).  Rows of the (host) table, at most UZ_WGRAD_REDUCE_MAX_ROWS per launch -- they travel as launch
 * parameters, so a captured CUDA graph holds them by value: */
#define UZ_WGRAD_REDUCE_MAX_ROWS 64
typedef struct UzWgradReduceDesc {
  const float* partial; /* workspace of the layer */
  float* dw;            /* fp32 [Cout][Cin][taps] */
  int splits, taps;
  int CoutP, CinP;      /* stored (padded) channel counts of the slabs */
  int Cout, Cin;        /* logical channel counts of dw */
  int unit_begin;       /* filled in by the library */
  int reserved;
} UzWgradReduceDesc;
int uz_conv_wgrad_partial(const void* x, int ldx, const void* dy, int lddy, int N, int D, int H, int W, int Cin, int Cout,
                          int taps, float* workspace, int accumulate, int* splits, void* stream);
int uz_wgrad_reduce_units(int Cout_logical, int Cin_logical, int taps);
int uz_wgrad_reduce_batched(const UzWgradReduceDesc* descs, int n, void* stream);

/* nn.AvgPool3d(2, 2, ceil_mode=True) on even sizes (models/phiseg3D.py:100) and its gradient. */
int uz_avgpool3_fwd(const void* x, int ldx, void* out, int ldo, int N, int Do, int Ho, int Wo, int C, void* stream);
int uz_avgpool3_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int Do, int Ho, int Wo, int C, void* stream);
/* F.interpolate / nn.Upsample(mode='trilinear', scale_factor=2, align_corners=True) (models/phiseg3D.py:143,291-294,
 * 382-386) into a channel slice (ldo) and its gradient; (d, h, w) is the LOW resolution. */
int uz_upsample3d_fwd(const void* x, int ldx, void* out, int ldo, int N, int d, int h, int w, int C, void* stream);
int uz_upsample3d_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int d, int h, int w, int C, void* stream);
/* Likelihood.s_layer 1x1x1 + nearest upsample to full resolution (models/phiseg3D.py:366-370,396-398 with the size fix
 * of SURVEY.md 8c): out fp32 NCDHW [B,ncls,d*f,h*f,wd*f].  Backward workspace rows: uz_slayer_bwd_num_blocks(B*d,h,wd). */
int uz_slayer3d_fwd(const void* feat, int ld, int C, const float* w, const float* bias, int ncls, int B, int d, int h,
                    int wd, int factor, float* out, void* stream);
int uz_slayer3d_bwd(const float* dout, const void* feat, int ld, int C, const float* w, int ncls, int B, int d, int h,
                    int wd, int factor, void* dfeat, int ldd, float* wpartial, float* bpartial, float* dw, float* db,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETZOO_B200_H_ */
