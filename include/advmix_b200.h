/* advmix_b200 - C ABI of the B200-native AdvMix augmentation + target hot path.
 *
 * The reference (AIprogrammer/AdvMix) has no FFI of its own on this path: its
 * "operator API" is four Python call boundaries.  Each entry
 * point below is what a binding for that boundary calls; the reference file:line it
 * replaces is cited per function.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *  - plain C, no torch / Python types.  All pointers are DEVICE pointers unless the
 *    name ends in _h.  The library never owns persistent user-visible memory; the
 *    only internal state is a per-device cache of small constant tables (Gaussian
 *    patches, Poisson CDFs, resampling coefficients) built lazily and guarded by a
 *    mutex.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - return 0 on success, <0 on error; advmix_last_error() gives the thread-local
 *    message.  No exceptions cross the ABI.
 *  - sm_100a cubins only; there is no CPU path and no other-architecture fallback.
 */
#ifndef ADVMIX_B200_H_
#define ADVMIX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVMIX_ABI_VERSION 1

#define ADVMIX_OK 0
#define ADVMIX_ERR_INVALID (-1)     /* bad argument */
#define ADVMIX_ERR_CUDA (-2)        /* CUDA runtime error, see advmix_last_error() */
#define ADVMIX_ERR_UNSUPPORTED (-3) /* shape / op outside the built path */
#define ADVMIX_ERR_WORKSPACE (-4)   /* workspace too small */

#define ADVMIX_F32 0
#define ADVMIX_BF16 1

/* OR into the `op` argument of advmix_corrupt_u8c3: use float32 arithmetic where the op has such a kernel
 * (gaussian_noise, contrast).  Result within 1 LSB of the float64 path (north_star's floating-point bar);
 * without the flag every op follows the reference's float64 / float32 operation order exactly. */
#define ADVMIX_CORRUPT_FAST 0x100

typedef void* advmix_stream_t; /* cudaStream_t */

int advmix_abi_version(void);
const char* advmix_last_error(void);
/* 0 if `device` is an sm_100 part this library has code for. */
int advmix_device_check(int device);

/* ---- a1: top-down affine crop --------------------------------------------------
 * Replaces cv2.warpAffine(data_numpy, trans, (w,h), INTER_LINEAR) at
 * lib/dataset/JointsDataset.py:190-195 and :324-329 (also lib/utils/transforms.py:125-133),
 * including the negative-stride flip view of :184-188.  Bit-exact with OpenCV's
 * fixed-point path (AB_BITS 10, 1/32-pixel taps, 15-bit weights, border constant 0).
 *
 * Sample b reads the HWC uint8 image at src_base + src_off[b] (rows src_pitch[b] bytes
 * apart, src_h[b] x src_w[b] pixels), mirrored left-right first when flip_lr[b] != 0.
 * M_fwd[b] is the forward 2x3 float64 matrix exactly as get_affine_transform returns
 * it (the kernel inverts it the way cv::warpAffine does).
 * Outputs (either may be NULL): dst_u8 [B][dh][dw][3]; dst_norm [B][3][dh][dw] in
 * norm_dtype, = norm_lut[c][value] (ToTensor()+Normalize(), tools/train.py:116-126,
 * applied at JointsDataset.py:331-332).  norm_lut is float32 [3][256]. */
int advmix_warp_affine_u8c3(const uint8_t* src_base, const int64_t* src_off, const int32_t* src_h,
                            const int32_t* src_w, const int64_t* src_pitch,
                            const uint8_t* flip_lr, const double* M_fwd, uint8_t* dst_u8,
                            void* dst_norm, const float* norm_lut, int B, int dw, int dh,
                            int norm_dtype, advmix_stream_t stream);

/* Host -> device staging of only the source rows a crop reads (the reference pays cv2.imread + a full
 * K x fp32 H2D per sample, lib/core/function.py:130; here a decoded uint8 source crosses PCIe once and only
 * where it is sampled).  All array arguments are HOST arrays (suffix _h); host_base_h should be pinned.
 * For sample b, rows [row_lo_h[b], row_hi_h[b]) of the image at byte offset off_h[b] (pitch_h[b] bytes per
 * row) are copied to the same offset of dev_base, asynchronously on `stream`. */
int advmix_h2d_source_rows(const uint8_t* host_base_h, uint8_t* dev_base, const int64_t* off_h,
                           const int64_t* pitch_h, const int32_t* row_lo_h, const int32_t* row_hi_h, int B,
                           advmix_stream_t stream);

/* Same purpose, one launch: the SMs gather the bytes straight out of the pinned host buffer (zero-copy
 * reads over PCIe) into the device buffer.  host_base must be pinned memory visible to the device under UVA
 * (cudaHostAlloc / torch pin_memory()); boxes is a device-readable array (pinned host memory is fine) of B
 *   struct { int64 off, pitch; int32 row_lo, row_hi, byte_lo, byte_hi; float qx[4], qy[4]; }   (64 bytes)
 * byte_lo / byte_hi: multiples of 16 within [0, pitch].  (qx, qy): the crop's destination rectangle mapped
 * into SOURCE pixel coordinates (a convex quadrilateral, corners in order; mirrored already for flipped
 * samples).  Row r in [row_lo,row_hi) copies the quad's x-extent over y in [r-3, r+3], widened by 3 pixels
 * and to 16-byte chunks, clamped to [byte_lo,byte_hi).  max_rows >= max_b (row_hi - row_lo) sizes the grid.
 * bytes_out (device uint64, nullable) accumulates the bytes copied. */
int advmix_h2d_source_boxes(const uint8_t* host_base, uint8_t* dev_base, const void* boxes, int B,
                            int max_rows, unsigned long long* bytes_out, advmix_stream_t stream);

/* get_affine_transform (lib/utils/transforms.py:69-101) for a batch, inv=0, shift=0.
 * center: float32 [B][2]; scale: float64 [B][2]; rot_deg: float64 [B]; M_fwd out: float64 [B][2][3].
 * scale_is_f32 != 0: `scale * 200.0` and `src_w * -0.5` are evaluated in float32, as numpy does
 * when JointsDataset's `s` is still a float32 array (numpy < 2); 0: in float64 (NEP-50 numpy, where
 * `s * np.clip(np.random.randn()...)` at JointsDataset.py:177 promotes `s` to float64).
 * Same float32 point triples as the reference; the 3-point solve restates cv2.getAffineTransform's
 * float64 LU (cv::solve DECOMP_LU -> LUImpl<double>) operation by operation, so M_fwd is bit-identical to
 * the reference's matrix wherever sin / cos of the rotation agree after the float32 point rounding
 * (always for rot = 0). */
int advmix_affine_matrices(const float* center, const double* scale, int scale_is_f32,
                           const double* rot_deg, double* M_fwd, int B, int out_w, int out_h,
                           advmix_stream_t stream);

/* fliplr_joints (lib/utils/transforms.py:44-58, when flip_lr[b]) followed by the
 * per-joint affine_transform of JointsDataset.py:197-199 (only where vis[j][0] > 0).
 * joints/vis: float64 [B][J][3].  flip_perm: int32 [J], index of the mirrored joint
 * (identity if NULL).  src_w: int32 [B] image widths (for w - x - 1). */
int advmix_joints_flip_affine(const double* joints_in, const double* vis_in,
                              const uint8_t* flip_lr, const int32_t* src_w,
                              const int32_t* flip_perm, const double* M_fwd, double* joints_out,
                              double* vis_out, int B, int J, advmix_stream_t stream);
/* Same, reading the rows from a device-resident record table: joints / vis of sample b are rows rec_idx[b] of
 * rec_joints / rec_vis (float64 [N][J][3]) - used by advmix_crop_targets_step_rec. */
int advmix_joints_flip_affine_rec(const double* rec_joints, const double* rec_vis, const int32_t* rec_idx,
                                  const uint8_t* flip_lr, const int32_t* src_w, const int32_t* flip_perm,
                                  const double* M_fwd, double* joints_out, double* vis_out, int B, int J,
                                  advmix_stream_t stream);

/* Fused forms of the two calls above for one training batch: get_affine_transform
 * (transforms.py:69-101) is evaluated inside the kernels from (center, scale, rot) with the
 * same device function advmix_affine_matrices uses, so results are bit-identical to the
 * two-step path while the crop and the joints/heatmap branch no longer wait on a matrix
 * kernel (JointsDataset.py:189-199 as two independent launches).
 * advmix_crop_csr_u8c3 == advmix_affine_matrices + advmix_warp_affine_u8c3.
 * advmix_joints_csr   == advmix_affine_matrices + advmix_joints_flip_affine; M_fwd_out
 * (f64 [B][2][3], may be NULL) receives the matrices for callers that keep them. */
int advmix_crop_csr_u8c3(const uint8_t* src_base, const int64_t* src_off, const int32_t* src_h,
                         const int32_t* src_w, const int64_t* src_pitch, const uint8_t* flip_lr,
                         const float* center, const double* scale, int scale_is_f32,
                         const double* rot_deg, uint8_t* dst_u8, void* dst_norm,
                         const float* norm_lut, int B, int dw, int dh, int norm_dtype,
                         advmix_stream_t stream);
int advmix_joints_csr(const double* joints_in, const double* vis_in, const uint8_t* flip_lr,
                      const int32_t* src_w, const int32_t* flip_perm, const float* center,
                      const double* scale, int scale_is_f32, const double* rot_deg,
                      double* M_fwd_out, double* joints_out, double* vis_out, int B, int J,
                      int out_w, int out_h, advmix_stream_t stream);

/* ToTensor()+Normalize() alone: uint8 [B][H][W][3] -> norm_dtype [B][3][H][W]. */
int advmix_normalize_u8c3(const uint8_t* in, void* out, const float* norm_lut, int B, int H,
                          int W, int norm_dtype, advmix_stream_t stream);

/* ---- a4: Gaussian heatmap targets ----------------------------------------------
 * Replaces JointsDataset.generate_target (lib/dataset/JointsDataset.py:412-491, live
 * branch :454-486).  joints, vis: float64 [B][J][3].  joints_weight: float32 [J] or
 * NULL (LOSS.USE_DIFFERENT_JOINTS_WEIGHT).  Outputs: hm float32 [B][J][Hh][Wh]
 * (fully written, zeros included), mu float32 [B][J][2] (NULL ok), tw float32 [B][J].
 * gauss_tab: float32 [(6*sigma+1)^2], the un-normalised patch of JointsDataset.py:470-476;
 * the host builds it with the reference's own numpy expression so it is bit-identical. */
int advmix_heatmap_targets(const double* joints, const double* vis, const float* gauss_tab,
                           const float* joints_weight, float* hm, float* mu, float* tw, int B, int J,
                           int Hh, int Wh, int img_w, int img_h, int sigma, advmix_stream_t stream);

/* ---- a8: one call per training step of the K = 1 path ------------------------------------
 * JointsDataset.get_clean (lib/dataset/JointsDataset.py:258-364) for a whole batch:
 * advmix_affine_matrices -> { advmix_warp_affine_u8c3 on `stream` || advmix_joints_flip_affine +
 * advmix_heatmap_targets on a library-owned side stream, forked / joined with events }.  Results are those of the
 * four calls.  Every per-step input comes out of ONE packed device buffer `params` (the host fills one pinned
 * buffer and issues one copy); sections in this order, each starting on a 16-byte boundary:
 *   src_off int64[B] | src_pitch int64[B] | src_h int32[B] | src_w int32[B] | scale f64[B][2] | rot_deg f64[B] |
 *   center f32[B][2] | flip_lr u8[B] | joints f64[B][J][3] | vis f64[B][J][3]
 * advmix_step_params_bytes(B, J) is its size.  Safe to capture into a CUDA graph after one eager call. */
size_t advmix_step_params_bytes(int B, int J);
int advmix_crop_targets_step(const uint8_t* src_base, const void* params, const int32_t* flip_perm,
                             const float* norm_lut, const float* gauss_tab, const float* joints_weight,
                             double* M_fwd, void* inp_norm, int norm_dtype, double* joints_out,
                             double* vis_out, float* hm, float* mu, float* tw, int B, int J, int out_w,
                             int out_h, int Hh, int Wh, int sigma, advmix_stream_t stream);

/* Record-table form of the same step: the db records' joints_3d / joints_3d_vis (JointsDataset.py:268-269) of the whole
 * shard stay on the device (rec_joints, rec_vis: float64 [N][J][3]) and `params` carries their row indices -
 * the advmix_crop_targets_step layout with the joints and vis sections replaced by ONE section rec_idx int32[B]
 * (advmix_step_rec_params_bytes).  14 KB instead of 223 KB cross PCIe per 256-sample step. */
size_t advmix_step_rec_params_bytes(int B, int J);
int advmix_crop_targets_step_rec(const uint8_t* src_base, const void* params, const double* rec_joints,
                                 const double* rec_vis, const int32_t* flip_perm, const float* norm_lut,
                                 const float* gauss_tab, const float* joints_weight, double* M_fwd, void* inp_norm,
                                 int norm_dtype, double* joints_out, double* vis_out, float* hm, float* mu, float* tw,
                                 int B, int J, int out_w, int out_h, int Hh, int Wh, int sigma, advmix_stream_t stream);

/* K = 3 (sample_times = 3) form of the step for the fused chain + mix path (JointsDataset.get_base + get_var x3,
 * lib/dataset/JointsDataset.py:135-256, without materialising the chains): uint8 crop, per-image autoaug plans
 * (advmix_autoaug_plan_u8c3), joints, the heat-map targets of the clean / autoaug chains and - when vis_gm_out / hm_gm /
 * tw_gm are given - the gridmask chain's visibility (advaug.py:153-165) and targets.  `params` = the K = 1 layout followed by
 *   autoaug ops int32[B][2] | autoaug mags f32[B][2] | gridmask params int32[B][4]   (sections 16-byte aligned);
 * the gridmask section is what advmix_chains_emit_u8c3 / advmix_chainmix_fwd / _bwd take as `gridmask_params`.
 * plan_ws: >= B*768*4 bytes.  mu may be NULL. */
size_t advmix_chains_step_params_bytes(int B, int J);
int advmix_crop_chains_step(const uint8_t* src_base, const void* params, const int32_t* flip_perm,
                            const float* gauss_tab, const float* joints_weight, double* M_fwd,
                            uint8_t* crop_u8, void* plans, void* plan_ws, size_t plan_ws_bytes,
                            double* joints_out, double* vis_out, double* vis_gm_out, float* hm, float* mu,
                            float* tw, float* hm_gm, float* tw_gm, int B, int J, int out_w, int out_h, int Hh,
                            int Wh, int sigma, advmix_stream_t stream);

/* ---- a3: AdvMix per-pixel convex mix ---------------------------------------------
 * Replaces lib/core/function.py:138-144 (and its autograd for the G step, :158-164).
 * x: K device pointers given in a HOST array x_h[K], each [B][C][H][W] in `dtype`.
 * w_or_logits float32 [B][K][H][W]; apply_softmax != 0 fuses F.softmax(dim=1).
 * out [B][C][H][W] in `dtype`; w_out (nullable) receives the softmax weights.
 * Forward arithmetic is mul-then-add in float32 in the reference's order
 * (bit-exact for apply_softmax = 0). */
int advmix_mix_fwd(const void* const* x_h, const float* w_or_logits, int apply_softmax, void* out,
                   float* w_out, int B, int K, int C, int H, int W, int dtype,
                   advmix_stream_t stream);
/* grad wrt weights (through_softmax = 0) or logits (= 1; `w` must then be the softmax
 * weights).  grad_out in `dtype`, grad_w float32 [B][K][H][W]. */
int advmix_mix_bwd(const void* const* x_h, const float* w, const void* grad_out, float* grad_w,
                   int through_softmax, int B, int K, int C, int H, int W, int dtype,
                   advmix_stream_t stream);

/* ---- a5/a6: reference-actual chains ------------------------------------------------
 * autoaug: ImageNetPolicy sub-policy application (lib/dataset/advaug.py:10-107).  The
 * reachable PIL ops are equalize(1) posterize(2) solarize(3) invert(4) sharpness(5);
 * 0 = skipped.  ops: int32 [B][2] (op1, op2 AFTER the probability coin flips),
 * mags: float32 [B][2] (posterize bits / solarize threshold / sharpness factor).
 * in/out uint8 [B][H][W][3]; out_norm (nullable) [B][3][H][W] norm_dtype.
 * workspace: advmix_autoaug_workspace_bytes(B,H,W). */
size_t advmix_autoaug_workspace_bytes(int B, int H, int W);
int advmix_autoaug_u8c3(const uint8_t* in, uint8_t* out, void* out_norm, const float* norm_lut,
                        const int32_t* ops, const float* mags, int B, int H, int W, int norm_dtype,
                        void* workspace, size_t ws_bytes, advmix_stream_t stream);
/* gridmask: grid_aug(mode=1, rotate=1, ratio=0.5) of lib/dataset/advaug.py:111-170 on the
 * normalised tensor.  params int32 [B][4] = (apply, d, st_h, st_w).  img [B][3][H][W]
 * in `dtype` (in -> out, may alias).  joints float64 [B][J][3]; vis_in -> vis_out float64
 * [B][J][3] with [j][0:2] zeroed where the joint lands on a masked cell.  img_in == img_out == NULL: only the
 * visibility update (the fused chain + mix path never materialises the masked tensor). */
int advmix_gridmask(const void* img_in, void* img_out, const int32_t* params, const double* joints,
                    const double* vis_in, double* vis_out, int B, int H, int W, int J, int dtype,
                    advmix_stream_t stream);

/* ---- N1: fused chain + mix -----------------------------------------
 * The reference pays K x float32 round trips for the mix (lib/core/function.py:137-146: 2 949 120 B per 256x192
 * sample).  These entry points read the uint8 crop once and RECOMPUTE the K = 3 reference-actual chains
 * ['clean', 'autoaug', 'gridmask'] (lib/dataset/JointsDataset.py:124, lib/dataset/advaug.py) in registers:
 *   clean = norm_lut[c][v];  autoaug = norm_lut[c][post[sharpen?(pre[v])]] from the per-image plan;
 *   gridmask = clean * mask(params)  (advaug.py:166, multiply on the normalised tensor).
 * plans: B records of advmix_autoaug_plan_bytes(1) bytes written by advmix_autoaug_plan_u8c3 (the histogram + plan
 * half of advmix_autoaug_u8c3; workspace >= B*768*4 bytes - still checked, no longer written: the histograms stay in
 * the shared memory of a thread-block cluster); NULL = chain 1 is the clean crop.
 * gridmask_params: int32 [B][4] as in advmix_gridmask; NULL = chain 2 is the clean crop.
 *
 * advmix_chains_emit_u8c3: G_input = torch.cat(inputs, dim=1) (function.py:137) written directly,
 *   [B][9][H][W] in `dtype` (chain-major: planes 3k..3k+2 are chain k).
 * advmix_chainmix_fwd: tmp = sum_k chain_k * w[:, k]  (function.py:138-144).  w_or_logits [B][3][H][W] in w_dtype
 *   (ADVMIX_F32 | ADVMIX_BF16); apply_softmax fuses F.softmax(dim=1); out [B][3][H][W] in out_dtype; w_out (float32,
 *   nullable) receives the weights.  Same float32 mul-then-add order as advmix_mix_fwd: bit-identical results.
 *   Bytes per 256x192 sample: 147 456 + 589 824 + 589 824 (float32 logits / out), 737 280 with bfloat16 both.
 * advmix_chainmix_bwd: grad wrt the logits (through_softmax = 1: the softmax is recomputed from the logits, no saved
 *   weights) or wrt the weights (= 0); grad_out in out_dtype, grad_w float32 [B][3][H][W].  (function.py:158-164)
 * W % 4 == 0, H, W <= 1024.
 *
 * advmix_mix_u8_fwd / _bwd: the general form for chains that cannot be recomputed in registers (the target workload
 * of BASELINE configs[2]: chains drawn from the 15x5 corruption set): K <= 4 uint8 HWC chain images x_h[k]
 * [B][H][W][3] (HOST array of device pointers), normalised by norm_lut in registers. */
size_t advmix_autoaug_plan_bytes(int B);
int advmix_autoaug_plan_u8c3(const uint8_t* in, const int32_t* ops, const float* mags, void* plans_out, int B, int H,
                             int W, void* workspace, size_t ws_bytes, advmix_stream_t stream);
int advmix_chains_emit_u8c3(const uint8_t* crop, const void* plans, const int32_t* gridmask_params,
                            const float* norm_lut, void* g_input, int B, int H, int W, int dtype,
                            advmix_stream_t stream);
int advmix_chainmix_fwd(const uint8_t* crop, const void* plans, const int32_t* gridmask_params, const float* norm_lut,
                        const void* w_or_logits, int w_dtype, int apply_softmax, void* out, int out_dtype,
                        float* w_out, int B, int H, int W, advmix_stream_t stream);
int advmix_chainmix_bwd(const uint8_t* crop, const void* plans, const int32_t* gridmask_params, const float* norm_lut,
                        const void* w_or_logits, int w_dtype, int through_softmax, const void* grad_out,
                        int out_dtype, float* grad_w, int B, int H, int W, advmix_stream_t stream);
int advmix_mix_u8_fwd(const uint8_t* const* x_h, const float* norm_lut, const void* w_or_logits, int w_dtype,
                      int apply_softmax, void* out, int out_dtype, float* w_out, int B, int K, int H, int W,
                      advmix_stream_t stream);
int advmix_mix_u8_bwd(const uint8_t* const* x_h, const float* norm_lut, const void* w_or_logits, int w_dtype,
                      int through_softmax, const void* grad_out, int out_dtype, float* grad_w, int B, int K, int H,
                      int W, advmix_stream_t stream);

/* ---- a2: imagecorruptions -----------------------------------------------------------
 * Replaces imagecorruptions.corrupt(image, severity, corruption_name) as called at
 * tools/make_datasets.py:41 and lib/dataset/JointsDataset.py:286.
 * op: 0 gaussian_noise 1 shot_noise 2 impulse_noise 3 defocus_blur 4 glass_blur
 *     5 motion_blur 6 zoom_blur 7 snow 8 frost 9 fog 10 brightness 11 contrast
 *     12 elastic_transform 13 pixelate 14 jpeg_compression, and the package's
 *     'validation' set (make_datasets.py:38 builds all 19, test_corruption.py:132):
 *     15 speckle_noise 16 gaussian_blur 17 spatter 18 saturate.   severity: 1..5.
 * in/out: uint8 [*][H][W][3].  n images are processed; image i is index
 * (idx ? idx[i] : i) of both in and out (idx: int32 device array, nullable).
 *
 * Random draws.  If rand_field / rand_param are given they are consumed (parity mode);
 * if NULL the same draws are generated in-register from Philox4x32-10 keyed by
 * (seed, sample_base + image index, op) (perf mode).  advmix_corrupt_fill_rand writes
 * exactly the values perf mode would consume, in the injected layout:
 *   op  rand_field (per image, image i at i*field_bytes)         rand_param double[n][4]
 *   0   float32 [H][W][3]    N(0,1)                              -
 *   1   float32 [H][W][3]    U[0,1) on the 2^-24 grid (inverse-CDF Poisson)  -
 *   2   float32 [2][H][W][3] U[0,1)  (flip, salt)                -
 *   4   int8    [iters][H][W][2] (dx,dy) in [-delta, delta-1]    -
 *   5   -                                                        [0] = angle, U(-45,45)
 *   7   float32 [H][W]       N(0,1)                              [0] = angle, U(-135,-45)
 *   8   -                                                        [0..2] = texture idx, x_start, y_start
 *   9   float32 [M][M]       U[0,1), M = next_pow2(max(H,W))     -
 *   12  float32 [2][H][W]    U[0,1)  (dx field, dy field)        -
 *   15  float32 [H][W][3]    N(0,1)                              -
 *   17  float32 [H][W]       N(0,1)  (liquid layer)              -
 * frost_bank: uint8 [frost_n][frost_h][frost_w][3] RGB textures (host code prepares
 * them; the package's PNG/JPG assets are not redistributable here). */
size_t advmix_corrupt_workspace_bytes(int op, int severity, int n, int H, int W);
size_t advmix_corrupt_rand_field_bytes(int op, int severity, int H, int W);
int advmix_corrupt_fill_rand(int op, int severity, int n, int H, int W, uint64_t seed,
                             int64_t sample_base, const int32_t* idx, void* rand_field,
                             double* rand_param, int frost_n, int frost_h, int frost_w,
                             advmix_stream_t stream);
int advmix_corrupt_u8c3(int op, int severity, const uint8_t* in, uint8_t* out, int n,
                        const int32_t* idx, int H, int W, const void* rand_field,
                        const double* rand_param, uint64_t seed, int64_t sample_base,
                        const uint8_t* frost_bank, int frost_n, int frost_h, int frost_w,
                        void* workspace, size_t ws_bytes, advmix_stream_t stream);

/* The five severities of one corruption from ONE read of the crops: tools/make_datasets.py:38-45 runs
 * `for severity in range(5)` innermost over the same image.  outs: HOST array of 5 device pointers, outs[s] = uint8
 * [*][H][W][3] receives exactly what advmix_corrupt_u8c3(op, s + 1, ...) writes in perf mode for the same seed (the
 * Philox key does not contain the severity).  gaussian_noise, impulse_noise, frost, brightness, contrast and zoom_blur
 * have fused kernels (shared draws / colour conversion / channel sums / zoom layers; gaussian_noise, contrast and
 * zoom_blur only with ADVMIX_CORRUPT_FAST); the other ops run their five per-severity launches.  No injected draws.
 * workspace: advmix_corrupt_sweep_workspace_bytes (the maximum over the severities). */
size_t advmix_corrupt_sweep_workspace_bytes(int op, int n, int H, int W);
int advmix_corrupt_sweep_u8c3(int op, const uint8_t* in, uint8_t* const* outs, int n, const int32_t* idx, int H, int W,
                              uint64_t seed, int64_t sample_base, const uint8_t* frost_bank, int frost_n, int frost_h,
                              int frost_w, void* workspace, size_t ws_bytes, advmix_stream_t stream);

/* ---- f3: heat-map consumers (validation / inference side) -------------------------------
 * advmix_heatmap_decode replaces get_max_preds (lib/core/inference.py:22-49) and, when
 * preds != NULL, get_final_preds (:52-95): arg-max per [Hh][Wh] plane (first maximum, like
 * np.argmax), coordinates zeroed where the maximum is <= 0, optional TEST.POST_PROCESS
 * quarter-pixel step (:64-76), then transform_preds (lib/utils/transforms.py:61-66) with
 * get_affine_transform(center, scale, 0, [Wh, Hh], inv=1).
 * heatmaps float32 [B][J][Hh][Wh]; center float32 [B][2]; scale float64 [B][2] with the
 * scale_is_f32 meaning of advmix_affine_matrices (validation metas are float32).
 * Outputs: preds float32 [B][J][2] image coordinates (NULL: skip the transform), maxvals
 * float32 [B][J], coords_hm float32 [B][J][2] heat-map coordinates (NULL ok).
 *
 * advmix_flip_merge replaces the FLIP_TEST merge at lib/core/function.py:241-261:
 * flip_back (lib/utils/transforms.py:16-41: reverse x, swap matched joints; flip_perm as in
 * advmix_joints_flip_affine), TEST.SHIFT_HEATMAP (columns 1.. take the flipped map's
 * columns 0..Wh-2), then (output + output_flipped) * 0.5.  merged may alias output.  output == NULL returns
 * the flipped-back (and shifted) map alone. */
int advmix_heatmap_decode(const float* heatmaps, const float* center, const double* scale,
                          int scale_is_f32, int post_process, float* preds, float* maxvals,
                          float* coords_hm, int B, int J, int Hh, int Wh, advmix_stream_t stream);
int advmix_flip_merge(const float* output, const float* output_flipped, const int32_t* flip_perm,
                      int shift_heatmap, float* merged, int B, int J, int Hh, int Wh,
                      advmix_stream_t stream);

/* ---- f1: baseline JPEG decode of the source images ---------------------------------------
 * Replaces cv2.imread(image_file, IMREAD_COLOR | IMREAD_IGNORE_ORIENTATION) at
 * lib/dataset/JointsDataset.py:148 (and PIL Image.open at tools/make_datasets.py:37): the encoded files
 * cross PCIe, the pixels are produced in HBM.  Numeric pipeline = libjpeg(-turbo)'s integer decoder
 * (JDCT_ISLOW, fancy up-sampling), bit-identical to cv2.imdecode / PIL.
 * Handled: baseline / extended-sequential Huffman, 8 bit, one interleaved scan, grayscale or YCbCr with
 * 4:4:4, 4:2:2 (h2v1) or 4:2:0 sampling, restart intervals, custom tables.  Anything else (progressive,
 * arithmetic, CMYK, RGB-coded, other samplings) is reported per image in the plan's status field and makes
 * advmix_jpeg_plan_h return ADVMIX_ERR_UNSUPPORTED - there is no CPU fallback.
 *
 * advmix_jpeg_plan_h (HOST arrays): parses the headers of B files (file b = files_h[off_h[b] .. +len_h[b]))
 * into B plan records of advmix_jpeg_plan_stride() bytes each (Huffman look-up tables, quantisation tables,
 * geometry, output and workspace layout).  Record fields the caller reads (byte offsets): int64 out_off @16,
 * int64 out_pitch @24 (3*width rounded up to 16), int32 width @32, height @36, ncomp @40, status @156
 * (0 ok, 1 corrupt, 2 progressive, 3 unsupported), int32 plane_w[4] @224, plane_h[4] @240.
 * Totals: *out_bytes (decoded images, packed), *coef_elems (int16 coefficients), *plane_bytes.
 *
 * advmix_jpeg_decode (DEVICE pointers): files = the same bytes at the same offsets, plans = the records;
 * writes image b as uint8 HWC (RGB, or BGR like cv2 if bgr != 0) at out + out_off with out_pitch bytes per
 * row.  workspace >= align256(2*coef_elems) + align256(plane_bytes) + files_bytes + 64 (files_bytes = size
 * of the files buffer; file offsets must be multiples of 16).  max_blocks / max_pixels: the largest per-image
 * number of 8x8 blocks (all components, padded planes) and of pixels, for grid sizing.  any_restart != 0 if
 * some file has a restart interval (int32 @52 of its record): those take the sequential entropy decoder. */
size_t advmix_jpeg_plan_stride(void);
int advmix_jpeg_plan_h(const uint8_t* files_h, const int64_t* off_h, const int64_t* len_h, int B,
                       void* plans_h, int64_t* out_bytes, int64_t* coef_elems, int64_t* plane_bytes);
int advmix_jpeg_decode(const uint8_t* files, const void* plans, int B, int max_blocks, int max_pixels,
                       uint8_t* out, void* workspace, size_t ws_bytes, int64_t coef_elems,
                       int64_t plane_bytes, int64_t files_bytes, int any_restart, int bgr,
                       advmix_stream_t stream);

/* ---- f1 (other half): baseline JPEG encode of the corrupted images ----------------------------
 * Replaces Image.fromarray(corrupted).save(corrupted_path) at tools/make_datasets.py:45 (PIL -> libjpeg-turbo
 * jpeg_write_scanlines with PIL's defaults: quality 75, YCbCr 4:2:0, JDCT_ISLOW, the Annex-K Huffman tables,
 * JFIF APP0 with density 1:1 / unit 0, no restart markers).  The files are produced in HBM, BYTE-identical to
 * what PIL writes, so only the ~10x smaller encoded bytes cross PCIe on the way to the disk.
 * images: uint8 [n][H][W][3] RGB on the device.  File i is written at out + i*out_stride, its size to
 * lengths[i] (int32, device); lengths[i] = -1 if the file (623 header bytes + stuffed scan + EOI) does not fit
 * out_stride, in which case nothing of it is written.  Typical files take 0.1-0.5 bytes per pixel; the hard
 * bound is 623 + 12*Hp*Wp + 2 (Hp, Wp = H, W rounded up to 16).  Includes libjpeg's edge rules for sizes that are not multiples of 16
 * (replicated samples inside real blocks, DC-only dummy blocks outside ceil(W/8) x ceil(H/8), jccoefct.c). */
size_t advmix_jpeg_encode_workspace_bytes(int n, int H, int W);
int advmix_jpeg_encode_u8c3(const uint8_t* images, int n, int H, int W, int quality, uint8_t* out,
                            size_t out_stride, int32_t* lengths, void* workspace, size_t ws_bytes,
                            advmix_stream_t stream);

/* Packs the n encoded files (file i: lengths[i] bytes at files + i*stride; lengths[i] < 0 counts as 0) back to back into
 * `packed` (capacity >= sum of the lengths; n*stride always suffices) and writes their byte offsets to offsets[0..n]
 * (int64, device; offsets[n] = total).  The host then copies offsets and exactly offsets[n] bytes: only encoded bytes cross
 * PCIe on the way to `open(corrupted_path, 'wb')` (tools/make_datasets.py:45).  files and stride 16-byte aligned. */
int advmix_pack_files(const uint8_t* files, size_t stride, const int32_t* lengths, int n, uint8_t* packed,
                      int64_t* offsets, advmix_stream_t stream);

/* ---- f4: per-record helpers of the dataset classes, batched (one thread per record) --------
 * advmix_xywh2cs: COCODataset._xywh2cs (lib/dataset/coco.py:205-220): boxes float64 [B][4] (x, y, w, h) ->
 *   center float32 [B][2], scale float32 [B][2] (aspect-ratio fix, / pixel_std, x1.25).
 * advmix_half_body_cs: JointsDataset.half_body_transform (lib/dataset/JointsDataset.py:69-111): joints / vis
 *   float64 [B][J][3]; upper_body_mask uint8 [J] (1 for self.upper_body_ids); randn_draw float64 [B] = the
 *   np.random.randn() of :80.  Outputs center / scale float32 [B][2] and valid uint8 [B] (0 where the
 *   reference returns (None, None)).
 * advmix_select_data: JointsDataset.select_data (:366-399): keep uint8 [B] = the record passes the
 *   joints-centre vs box-centre test.  center / scale float32 [B][2] as stored in the db records. */
int advmix_xywh2cs(const double* boxes_xywh, float* center, float* scale, int B, double aspect_ratio,
                   double pixel_std, advmix_stream_t stream);
int advmix_half_body_cs(const double* joints, const double* vis, const uint8_t* upper_body_mask,
                        const double* randn_draw, float* center, float* scale, uint8_t* valid, int B,
                        int J, double aspect_ratio, double pixel_std, advmix_stream_t stream);
int advmix_select_data(const double* joints, const double* vis, const float* center, const float* scale,
                       uint8_t* keep, int B, int J, double pixel_std, advmix_stream_t stream);
/* advmix_base_cs: the centre / scale bookkeeping of get_base / get_clean (lib/dataset/JointsDataset.py:167-188,
 * :303-322) for a batch, given the random draws (the host draws them, in the reference's RNG order if it wants
 * a replay; all geometry happens here):
 *   take_half_body uint8 [B] (nullable): the sample passed `sum(vis) > NUM_JOINTS_HALF_BODY and rand() < PROB_HALF_BODY`
 *     (:168-169); its centre / scale become half_body_transform's (hb_randn float64 [B] = the randn() of :80)
 *     unless that returns (None, None) (:174-175);
 *   scale_factor float64 [B] (nullable = evaluation): s = s * clip(randn()*sf + 1, 1-sf, 1+sf) (:177-179), a float32
 *     array times a numpy float64 scalar, i.e. float64 under NEP 50;
 *   flip_lr uint8 [B] (nullable): c[0] = width - c[0] - 1 in float32 (:188), src_w int32 [B].
 * rec_center / rec_scale float32 [B][2] are the db records' values.  Outputs: center float32 [B][2], scale float64 [B][2]
 * (exactly float32-representable when scale_factor is NULL: pass scale_is_f32 = 1 downstream). */
int advmix_base_cs(const float* rec_center, const float* rec_scale, const double* joints, const double* vis,
                   const uint8_t* upper_body_mask, const uint8_t* take_half_body, const double* hb_randn,
                   const double* scale_factor, const uint8_t* flip_lr, const int32_t* src_w, float* center, double* scale,
                   int B, int J, double aspect_ratio, double pixel_std, advmix_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADVMIX_B200_H_ */
