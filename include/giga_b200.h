/*
 * giga_b200.h -- C ABI of the B200-native (sm_100a) GIGA dense-inference hot path.
 *
 * The reference (UT-Austin-RPL/GIGA) has no native FFI on this path: its boundary is
 * the Python nn.Module `ConvolutionalOccupancyNetwork` plus the library ops underneath
 * it.  Every entry point below therefore names the reference *Python* interface (or
 * library op) it replaces, file:line relative to /root/reference/src/vgn.  The Python
 * host layer in giga_b200/ (ctypes) mirrors the reference classes
 * on top of exactly these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success, a negative GIGA_E* code on failure, and
 *     leaves a message retrievable with giga_last_error() (thread local);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all
 *     device work of a call is enqueued on it, nothing synchronises unless stated;
 *   - Streams: a giga_ctx owns ONE set of device workspaces (activations, tile-dependency
 *     counters, planner volumes), so its calls execute in SUBMISSION order: an entry point
 *     called on a different stream than the previous one (including the internal streams of
 *     giga_forward_host_submit) first makes its stream wait for the previous call's work
 *     (event).  Calls from several host threads into one ctx must be serialised by the
 *     caller; use one ctx per thread for concurrency.  giga_ctx_set_param*() and
 *     giga_ctx_commit_params() drain the whole device before touching the parameters;
 *   - device tensors are dense fp32, layouts stated per argument.  There is NO CPU
 *     fallback: without a CUDA device every compute entry point fails with GIGA_ENODEV.
 *
 * Tensor layouts in HBM
 *   tsdf    [B][40][40][40]            x[b][ix][iy][iz]          (voxels.py:89)
 *   planes  [3][B][40][40][32]         plane k in (xz, xy, yz), channels-last:
 *                                      planes[k][b][row][col][c]; row/col per
 *                                      common.py:314 (xz: row=z col=x, xy: row=y col=x,
 *                                      yz: row=z col=y).  encode_inputs() exposes
 *                                      plane k as a (B,32,40,40) permuted view.
 *   points  [B][N][3]                  p[b][n][xyz] in the unit cube (clamped outside)
 *   qual    [B][N]   rot [B][N][4]   width [B][N]   occ [B][N]
 */
#ifndef GIGA_B200_H_
#define GIGA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define GIGA_GRID 40        /* TSDF / plane resolution (networks.py:95, detection_implicit.py:28) */
#define GIGA_CDIM 32        /* plane feature channels (networks.py:113) */
#define GIGA_HIDDEN 32      /* decoder hidden size (networks.py:109) */

#define GIGA_HEAD_QUAL  1u  /* decoder_qual  -> sigmoid            (models/__init__.py:119-120) */
#define GIGA_HEAD_ROT   2u  /* decoder_rot   -> L2 normalise        (models/__init__.py:121-122) */
#define GIGA_HEAD_WIDTH 4u  /* decoder_width -> raw                 (models/__init__.py:123)     */
#define GIGA_HEAD_TSDF  8u  /* decoder_tsdf  -> raw occupancy logit (models/__init__.py:64,71)   */
#define GIGA_HEAD_RAW  16u  /* modifier: skip the sigmoid / normalise epilogues, i.e. the bare
                               `LocalDecoder.forward` output (decoder.py:173-176) */

#define GIGA_OK        0
#define GIGA_EINVAL   -1    /* bad argument (shape, null pointer, unknown name) */
#define GIGA_ENODEV   -2    /* no CUDA device / wrong architecture */
#define GIGA_ECUDA    -3    /* CUDA runtime error (message has the detail) */
#define GIGA_ESTATE   -4    /* parameters missing or not committed */

typedef struct giga_ctx giga_ctx;   /* opaque: packed parameters + device workspaces of one model on one GPU */

/* library --------------------------------------------------------------------------- */
int         giga_version(void);
const char *giga_last_error(void);

/* context: replaces `get_network(name).to(device)` (networks.py:10-18,33) as the owner of
 * device state.  `device` is a CUDA ordinal. */
int  giga_ctx_create(giga_ctx **out, int device);
void giga_ctx_destroy(giga_ctx *ctx);

/* Parameters: replaces `net.load_state_dict(...)` (networks.py:34).  `name` is the
 * reference state_dict key (e.g. "encoder.unet.down_convs.0.conv1.weight",
 * "decoder_qual.fc_c.3.bias"; full list SURVEY.md 8b), `data` is the tensor in the
 * reference's own layout, `numel` its element count (checked), `on_device` says whether
 * `data` is a device pointer.  giga_ctx_commit_params() re-packs everything into the
 * kernel layouts (synchronous) and must be called after any parameter change. */
int giga_ctx_set_param(giga_ctx *ctx, const char *name, const float *data, long numel, int on_device);
/* all parameters in one call: `flat` holds the n tensors back to back (offsets[i], numels[i] in elements), on the device
 * or the host; one copy instead of n (a training loop re-uploads every step: optimizer.step() changes every tensor). */
int giga_ctx_set_params_flat(giga_ctx *ctx, int n, const char *const *names, const long *offsets, const long *numels,
                             const float *flat, long total, int on_device);
int giga_ctx_commit_params(giga_ctx *ctx);
/* bit mask of GIGA_HEAD_* whose parameters are committed (giga_aff lacks TSDF; giga_geo has only TSDF) */
unsigned giga_ctx_heads(const giga_ctx *ctx);

/* encoder: replaces `LocalVoxelEncoder.forward` = `net.encode_inputs(inputs)`
 * (encoder/voxels.py:89-121; models/__init__.py:74-87): Conv3d(1,32,3,p=1)+ReLU, the
 * three scatter_mean plane projections (voxels.py:57-66, common.py:238-261,303-318) and
 * the shared UNet (encoder/unet.py:225-239) on each plane. */
int giga_encode(giga_ctx *ctx, const float *tsdf, int B, float *planes, void *stream);

/* decoder: replaces `LocalDecoder.forward` for every head in `heads` evaluated at the
 * SAME points (decoder.py:133-176 incl. sample_plane_feature :117-122, layers.py:39-47)
 * plus the head epilogues of `decode` (models/__init__.py:111-124).  Output pointers of
 * heads not in `heads` may be NULL. */
int giga_decode(giga_ctx *ctx, const float *planes, int B, const float *points, int N, unsigned heads,
                float *qual, float *rot, float *width, float *occ, void *stream);

/* feature sampling only: mode 0 = concatenated 96-d feature (decoder.py:141-147),
 * mode 1 = summed 32-d `query_feature` (decoder.py:178-191).  out [B][N][96|32]. */
int giga_sample_feature(giga_ctx *ctx, const float *planes, int B, const float *points, int N, int mode,
                        float *out, void *stream);

/* final grasp-score reduction: per scene max / first arg-max of qual over its N points
 * (what `VGNImplicit.select` ultimately keeps, detection_implicit.py:146-174).  Writes
 * best_val[b], best_idx[b]; the multi-GPU path points these into the all-gather buffer. */
int giga_scene_argmax(giga_ctx *ctx, const float *qual, int B, int N, float *best_val, int *best_idx, void *stream);

/* whole forward with DEVICE buffers in ONE call: replaces `net(inputs, p, p_tsdf=...)`
 * (`ConvolutionalOccupancyNetwork.forward`, models/__init__.py:42-67): giga_encode, giga_decode of the
 * committed grasp heads at `p` (Ng points; p may be NULL/0 for giga_geo), giga_decode of the TSDF head
 * at `p_tsdf` (No points; may be NULL/0) and, when best_val/best_idx are non-NULL, giga_scene_argmax of
 * qual.  `planes` receives the plane features ([3][B][40][40][32]); NULL keeps them in a ctx-owned
 * buffer.  Nothing synchronises. */
int giga_forward(giga_ctx *ctx, const float *tsdf, int B, const float *p, int Ng, const float *p_tsdf, int No,
                 float *planes, float *qual, float *rot, float *width, float *occ, float *best_val, int *best_idx,
                 void *stream);

/* end-to-end host entry: replaces `predict()` (detection_implicit.py:99-113) /
 * `net(x, pos, p_tsdf=pos_occ)` (scripts/train_giga.py:204) with HOST buffers: H2D of
 * tsdf/points, encode, decode of the grasp heads at `p` (Ng pts) and of the TSDF head at
 * `p_tsdf` (No pts, may be NULL/0), D2H of the results; returns after the stream has
 * drained.  Host buffers should be pinned for full copy bandwidth. */
int giga_forward_host(giga_ctx *ctx, const float *tsdf, int B, const float *p, int Ng, const float *p_tsdf, int No,
                      float *qual, float *rot, float *width, float *occ, void *stream);

/* pipelined variant of giga_forward_host for serving loops: two request slots.  submit() enqueues the
 * H2D copies on a copy stream, the kernels on the ctx's compute stream and the D2H copies on a third
 * stream (event-chained) and returns at once; wait() blocks until that slot's results are in the host
 * buffers.  With two slots in flight the PCIe copies of request i+1 / i-1 overlap the kernels of
 * request i.  Host buffers must be pinned and stay valid until wait() returns. */
int giga_forward_host_submit(giga_ctx *ctx, int slot, const float *tsdf, int B, const float *p, int Ng, const float *p_tsdf, int No,
                             float *qual, float *rot, float *width, float *occ);
int giga_forward_host_wait(giga_ctx *ctx, int slot);

/* planner post-processing ------------------------------------------------
 * What detection_implicit.py does on the host with scipy.ndimage between predict() and the Grasp
 * list, on the device for B scenes at once.  Volumes are [B][40][40][40] fp32 with flat voxel index
 * v = (ix*40 + iy)*40 + iz = index of the lattice query point (VGNImplicit.__init__ :28-31). */
typedef struct giga_select_params {
  double gaussian_sigma;      /* process(gaussian_filter_sigma=1.0)                  :119,128-130 */
  float  min_width, max_width;/* process(min_width=0.033, max_width=0.233)           :120-121,141 */
  float  out_th;              /* process(out_th=0.5): surface mask from the TSDF     :122,133-138 */
  int    lim_x, lim_y, lim_z; /* bound(): int(limit / voxel_size) = 2, 2, 7          :87-97       */
  float  low_th;              /* LOW_TH = 0.5                                        :15,148      */
  float  threshold;           /* select(threshold=qual_th=0.9)                       :146,153     */
  int    force_detection;     /* keep only the best grasp when nothing passes        :149-150,169 */
  int    max_filter_size;     /* NMS window of ndimage.maximum_filter, 4 (8 to visualise) :156    */
} giga_select_params;
/* the reference's defaults for the 0.3 m workspace (voxel_size 0.3/40) */
void giga_select_params_default(giga_select_params *p);
/* scipy's fp64 gaussian kernel (exp(-x^2/(2 sigma^2)) / sum, x = -radius..radius) as the device passes use it;
 * out has 2*radius+1 entries.  Host-only helper (no GPU needed) so the weights can be checked against numpy. */
int giga_gaussian_kernel1d(double sigma, int radius, double *out);
/* process() + bound() + select() (detection_implicit.py:115-143, 87-97, 146-174) on DEVICE volumes:
 * tsdf = the `tsdf_process` grid, qual/width [B][64000], rot [B][64000][4] = the three predicted volumes.
 * Outputs (device): count[b] = number of grasps found for scene b (may exceed K); the first min(count,K)
 * entries of score/index/out_rot/out_width [B][K](x4) are the grasps sorted by descending score (ties:
 * larger voxel index first): score = smoothed quality, index = flat voxel index, out_rot = the raw predicted
 * quaternion, out_width = predicted width.  qual_vol (optional, [B][64000]) receives the processed quality
 * volume (what process()+bound() return).  Entries past min(count,K) are left untouched (giga_detect_host zeroes them).
 * rot and out_rot must be 16-byte aligned.  Nothing synchronises. */
int giga_select_grasps(giga_ctx *ctx, const float *tsdf, const float *qual, const float *rot, const float *width, int B,
                       const giga_select_params *prm, int K, int *count, float *score, int *index, float *out_rot,
                       float *out_width, float *qual_vol, void *stream);
/* the query lattice of the planner (`self.pos`, detection_implicit.py:28-31), HOST pointer [64000][3];
 * copied to the device once and broadcast over the scenes of giga_detect calls (synchronous). */
int giga_ctx_set_lattice(giga_ctx *ctx, const float *pos, int N);
/* the whole planner, DEVICE buffers: replaces VGNImplicit.__call__ (detection_implicit.py:33-58) from the
 * TSDF to the sorted grasps = giga_forward at the lattice (grasp heads only) + giga_select_grasps.
 * tsdf_process may be NULL (= tsdf).  Outputs as giga_select_grasps. */
int giga_detect(giga_ctx *ctx, const float *tsdf, const float *tsdf_process, int B, const giga_select_params *prm, int K,
                int *count, float *score, int *index, float *out_rot, float *out_width, void *stream);
/* same with HOST buffers in and out (H2D of the TSDFs, giga_detect, D2H of B*(1 + 7K) words); returns
 * after the stream has drained.  This is the call a simulation loop makes per planning step: inputs and results are
 * staged through pinned buffers owned by the ctx, and repeated calls of one configuration (B, K, parameters) replay a
 * captured CUDA graph (one launch instead of ~25 API calls). */
int giga_detect_host(giga_ctx *ctx, const float *tsdf, const float *tsdf_process, int B, const giga_select_params *prm, int K,
                     int *count, float *score, int *index, float *out_rot, float *out_width, void *stream);

/* VGN baseline network -----------------------------------------------------
 * Replaces `ConvNet.forward` (networks.py:48-63 with Encoder :172-188 and Decoder :191-212; `get_network("vgn")`): three stride-2
 * Conv3d + ReLU, three Conv3d + ReLU with nearest up-sampling, the three k=5 heads with sigmoid / normalise / raw epilogues.
 * Parameters are set by their reference state_dict keys (encoder.conv1.weight ... conv_width.bias) with giga_ctx_set_param*.
 * tsdf [B][40][40][40] -> qual [B][64000], rot [B][64000][4] (the reference's (B,4,40,40,40) with the quaternion innermost: the layout
 * giga_select_grasps takes), width [B][64000]; flat voxel index (ix*40 + iy)*40 + iz.  Device pointers, nothing synchronises. */
int giga_vgn_forward(giga_ctx *ctx, const float *tsdf, int B, float *qual, float *rot, float *width, void *stream);

/* Training-step tail ----------------------------------------------------------
 * giga_loss replaces `loss_fn` of scripts/train_giga.py:161-195 and its autograd graph: the value AND the gradient of loss.mean() with
 * respect to the four (post-`select`, :153-158) predictions in one launch.  label_pred [B] (probability), rot_pred [B][4], width_pred [B],
 * occ_pred [B][M] (probability, sigmoid already applied); targets label [B], rotations [B][2][4], width [B], occ [B][M].
 * loss_out [5] = loss_dict's means in its order: loss_qual, loss_rot, loss_width, loss_occ, loss_all.  g_* (each optional) receive
 * d loss_all / d prediction with the shapes of the predictions.  F.binary_cross_entropy semantics (log clamp -100, backward
 * (x - t) / max((1 - x) x, 1e-12)); torch.min's tie rule (half to each operand).  Device pointers, nothing synchronises. */
int giga_loss(giga_ctx *ctx, const float *label_pred, const float *rot_pred, const float *width_pred, const float *occ_pred,
              const float *label, const float *rotations, const float *width, const float *occ, int B, int M, float *loss_out,
              float *g_label, float *g_rot, float *g_width, float *g_occ, void *stream);
/* giga_adam_step replaces `torch.optim.Adam(net.parameters(), lr=args.lr).step()` (scripts/train_giga.py:67, 208-209; update rule of
 * torch/optim/adam.py, amsgrad=False, maximize=False) over ONE flat fp32 buffer of n elements holding every parameter (and the same
 * layout for grad / exp_avg / exp_avg_sq; 16-byte aligned device pointers): one launch instead of the optimizer's per-tensor lists.
 * step = the 1-based step count (bias corrections are computed on the host in double like the Python implementation). */
int giga_adam_step(giga_ctx *ctx, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long n, int step, double lr,
                   double beta1, double beta2, double eps, double weight_decay, void *stream);

/* Device-side parameter commit: giga_ctx_set_param* + giga_ctx_commit_params for tensors that already live on the device.  `values[i]`
 * = device pointer (fp32, 16-byte aligned, the reference's state-dict layout) of tensor `names[i]`; every operand layout of the
 * inference kernels (fp32 blobs, fp16 hi/lo splits of the tensor-core convolutions and decoder, head constants) is written by kernels
 * on `stream` -- bit-identical to the host packer -- and only conv_in's 3.6 KB of taps (kernel parameters of its kernel) are read back.
 * GIGA-family models only (all encoder tensors + whole heads).  Synchronises the device (the blobs may be in use) and `stream`. */
int giga_ctx_commit_device(giga_ctx *ctx, int n, const char *const *names, const float *const *values, void *stream);
/* debug: copy a packed operand blob to the host (0 = encoder blob, 1 = fp32 head blob, 2 = head constants, 3..7 = streamed decoder
 * weights per job type); returns the byte count (0 = not present) or a negative error */
long giga_debug_blob(giga_ctx *ctx, int which, void *dst, long capacity);

/* Native training step ---------------------------------------------------------------------------
 * Replaces the autograd graph of `_update` in scripts/train_giga.py:199-211 (net(x, pos, p_tsdf=pos_occ) ... loss.backward()): the
 * differentiable forward of conv_onet/models/__init__.py:42-67 and its backward (ATen/cuDNN conv3d / conv2d / conv_transpose2d
 * backward-data and backward-filter, max_pool2d_with_indices_backward, grid_sampler_2d_backward, addmm, threshold_backward) as
 * hand-written kernels.  Parameters stay ON THE DEVICE in the reference's state-dict layouts: nothing is packed on the host.
 *
 * giga_train_bind: n tensors by reference state-dict key (`names`), `values[i]` / `grads[i]` = device pointers (fp32, 16-byte
 *   aligned) of the parameter and of the buffer its gradient is ACCUMULATED into (+=, as autograd's AccumulateGrad; zero it to get
 *   the plain gradient).  All encoder tensors are required; heads are optional but must be complete.  Cheap to call every step
 *   (the device-side packing table is rebuilt only when a pointer changes).
 * giga_train_forward: same arguments / outputs as giga_forward (qual = sigmoid, rot = normalised, width, occupancy logits) computed
 *   from the bound parameters' CURRENT values on the fp32 FMA-pipe kernels; every activation is kept in the ctx.  detach_tsdf != 0 =
 *   the `giga_detach` variant (models/__init__.py:61-62: the TSDF head's features carry no gradient to the encoder).
 * giga_train_backward: g_* = gradients of the loss w.r.t. the four outputs of the LAST giga_train_forward ([B][Ng], [B][Ng][4],
 *   [B][Ng], [B][No]; NULL = that output does not contribute); adds the parameter gradients into the bound buffers.  tsdf / p /
 *   p_tsdf of the forward must still be alive.  One backward per forward.  g_p (optional, [B][Ng][3], accumulated +=) receives the
 *   gradient w.r.t. the grasp query positions (`grad_refine`, models/__init__.py:136-164: fc_p plus grid_sampler's grid gradient).  Weight-gradient sums use floating-point atomics (run-to-run
 *   differences at the 1e-6 relative level, as PyTorch's default cuDNN algorithms). */
int giga_train_bind(giga_ctx *ctx, int n, const char *const *names, const float *const *values, float *const *grads);
int giga_train_forward(giga_ctx *ctx, const float *tsdf, int B, const float *p, int Ng, const float *p_tsdf, int No, int detach_tsdf,
                       float *qual, float *rot, float *width, float *occ, void *stream);
int giga_train_backward(giga_ctx *ctx, const float *g_qual, const float *g_rot, const float *g_width, const float *g_occ, float *g_p,
                        void *stream);

/* TSDF integration / read-out -------------------------------------------------
 * Replaces `TSDFVolume.integrate` / `get_grid` of vgn/perception.py:79-115, i.e. open3d.pipelines.integration.UniformTSDFVolume
 * (open3d==0.12.0, third party) `integrate` on an RGBD image created with depth_scale / depth_trunc, and extract_voxel_grid + the
 * Python voxel loop.  tsdf / weight: device fp32 [R][R][R] (x, y, z), zero-initialised by the caller for a new volume and updated in
 * place; depth: device fp32 [n_views][height][width] (metres * depth_scale); extrinsics: HOST double [n_views][16], row-major 4x4
 * T_eye_task as `Transform.as_matrix()` gives it; size = edge length of the volume (origin 0), sdf_trunc as perception.py:72 (4 voxels).
 * Views are folded into the running average in submission order (one launch per 16 views).  giga_tsdf_grid writes the network's input
 * grid ([R][R][R]: (tsdf + 1) / 2 where the voxel was observed and -0.98 <= tsdf < 0.98, else 0) -- it stays on the device, so
 * perception -> giga_forward needs no host round trip. */
int giga_tsdf_integrate(giga_ctx *ctx, float *tsdf, float *weight, int resolution, double size, double sdf_trunc, const float *depth,
                        int n_views, int width, int height, double fx, double fy, double cx, double cy, const double *extrinsics,
                        double depth_scale, double depth_trunc, void *stream);
int giga_tsdf_grid(giga_ctx *ctx, const float *tsdf, const float *weight, int resolution, float *grid, void *stream);

/* Generator3D occupancy sweep ----------------------------------------------
 * Replaces the MISE loop of `Generator3D.generate_from_latent` (ConvONets/conv_onet/generation.py:127-143) together with the Cython
 * octree it drives (ConvONets/utils/libmise/mise.pyx): query -> eval_points/decode_occ -> update/subdivide until no grid point is
 * unknown, then to_dense.  The reference moves every query batch and every value through the host; here the bookkeeping (dense
 * restatement of the octree on the finest lattice), the point arithmetic (float64, rounded to fp32 as torch.FloatTensor does) and the
 * TSDF-head evaluation stay on the device and only the per-level point COUNT (4 bytes) is read back.
 * planes = the plane features of ONE scene ([3][1][40][40][32], as giga_encode writes them); resolution0 / upsampling_steps (>= 1) =
 * Generator3D's arguments (16 / 3 by default: a 129^3 lattice); threshold = log(th) - log(1 - th) as generation.py:110 computes it;
 * box_size = 1 + padding.  value_grid (device, fp32 [(R+1)^3], R = resolution0 << upsampling_steps) receives mesh_extractor.to_dense()
 * (the reference's float64 grid holds the same fp32 values widened).  stats (host, optional) = {iterations, points evaluated}.
 * Synchronises `stream` once per level. */
int giga_mise_sweep(giga_ctx *ctx, const float *planes, int resolution0, int upsampling_steps, double threshold, double box_size,
                    float *value_grid, int *stats, void *stream);

/* introspection ------------------------------------------------------------------- */
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
long giga_ctx_launch_count(const giga_ctx *ctx);
/* Range guard of the tensor-core decoder.  Its operands are fp16 (hi, lo) pairs: a plane feature or hidden activation beyond
 * +-65504 cannot be represented.  Such a value is never clamped silently: it becomes NaN, propagates to the head outputs of that
 * query point (qual/rot/width/occ read NaN) and is counted on the device.  Returns the number of non-finite head outputs produced
 * since the last reset (synchronises the device; giga_last_error() then holds a description), or a negative error.  The
 * reference's fp32 path has no such limit (decoder.py:165-174); GIGA's activations are O(10). */
long giga_ctx_overflow_count(giga_ctx *ctx, int reset);
/* implementation switches (A/B testing; defaults are the fastest parity-clean variants):
 *   "decoder_impl": 1 = warp-specialised tcgen05 decoder, 3xFP16 operand splitting, A operand in tensor memory (default),
 *                   0 = fp32 FMA pipe (no fp16 operand range limit: the fallback for giga_ctx_overflow_count() != 0)
 *   "encoder_impl": U-Net convolutions: 1 = tcgen05 3xFP16 with persistent CTAs (default), 0 = fp32 FMA pipe
 *   "graph":        1 = giga_detect_host replays a captured CUDA graph of the whole call from its third invocation of a
 *                   configuration on (default), 0 = always enqueue kernel by kernel
 *   "tile_deps":    1 = consecutive same-resolution U-Net layers synchronise per position group instead of per grid
 *                   (default; needs "pdl"), 0 = every layer waits for its whole predecessor
 *   "pdl":          1 = programmatic dependent launch between the fast-path kernels (default), 0 = plain stream order */
int giga_ctx_set_option(giga_ctx *ctx, const char *key, int value);
/* per-kernel device timing for the roofline report: when enabled every kernel launch is bracketed
 * by CUDA events recorded on the launch stream.  giga_ctx_timing_report() waits for the recorded
 * events, writes one text line "name launches total_ms" per distinct kernel into buf (in first-launch
 * order), clears the record and returns the string length. */
int  giga_ctx_set_timing(giga_ctx *ctx, int enabled);
long giga_ctx_timing_report(giga_ctx *ctx, char *buf, long cap);
/* copy an intermediate activation of the LAST giga_encode call into dst (device, fp32).
 * names: "pre" [3][B][32][40][40] (planes before the U-Net, NCHW), "d0c1","d0c2","p0",
 * "d1c1","d1c2","p1","d2c1","d2c2","u0","u0c1","u0c2","u1","u1c1","u1c2" (NCHW over the
 * 3B stacked planes).  Returns the element count, or a negative error. */
long giga_debug_copy(giga_ctx *ctx, const char *name, float *dst, long capacity, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GIGA_B200_H_ */
