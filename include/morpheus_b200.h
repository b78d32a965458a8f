/*
 * morpheus_b200 -- C ABI of the B200-native MorpheuS render-and-loss hot path.
 *
 * Every entry point: extern "C", plain pointers + sizes, returns 0 on success or a negative
 * MB_E* code (text via mb_last_error()).  All pointers are DEVICE pointers owned by the caller
 * unless stated otherwise; nothing is allocated or freed inside (same ownership rule as the
 * reference extension: external/encoders/gridencoder/grid.py:50,56,84,87).  `stream` is a
 * cudaStream_t passed as void* (the reference launches on the legacy default stream,
 * gridencoder.cu:386; we take the caller's stream so CUDA graphs / NCCL overlap work).
 *
 * Reference interfaces replaced (file:line relative to /root/reference):
 *   mb_grid_encode_forward/backward  <- external/encoders/gridencoder/src/gridencoder.h:12-13
 *                                       (bindings.cpp:6-7, called from grid.py:61,91)
 *   mb_sample_rays_*                 <- nerfacc OccGridEstimator.sampling        (morpheus.py:629-638)
 *   mb_composite_*                   <- nerfacc render_weight_from_density +
 *                                       accumulate_along_rays                   (morpheus.py:675-685)
 *   mb_field_forward/backward        <- scene_representation.forward/density/normal/warp
 *                                       (models/model.py:273-307,367-398,412-437,439-533)
 *   mb_occ_update                    <- nerfacc OccGridEstimator.update_every_n_steps (morpheus.py:905-913)
 *   mb_adam_step                     <- torch.optim.Adam over get_params_all()   (morpheus.py:154-155)
 *   mb_sds_grad                      <- Zero123.train_step scalar math           (models/guidance/zero123_utils.py:177-212)
 *   mb_ray_points_*                  <- xyzs = rays_o[ri] + rays_d[ri]*t         (morpheus.py:645-646)
 *   mb_pack_arena_*                  <- nn.utils.weight_norm of MLP layers       (models/decoders.py:51-52)
 *   mb_sdf_loss_*                    <- utils.get_sdf_loss                        (utils.py:91-113)
 */
#ifndef MORPHEUS_B200_H
#define MORPHEUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK 0
#define MB_EINVAL (-1)      /* bad argument (null pointer, unsupported D/C/dtype ...) */
#define MB_ECUDA (-2)       /* CUDA runtime error at launch */
#define MB_EUNSUPPORTED (-3)

#define MB_DTYPE_F32 0

typedef void* mb_stream_t;

int mb_version(void);
const char* mb_last_error(void);
/* number of SMs of the current device (grid sizing is a multiple of this) */
int mb_sm_count(void);

/* ---- (1) grid encoder: argument order follows gridencoder.h:12-13 exactly ------------------- */
/* inputs [B,D] f32 in [0,1]; embeddings [sO,C]; offsets [L+1] i32; outputs [L,B,C] (caller
 * zero-fills levels >= max_level, grid.py:53); dy_dx [B, L*D*C] or NULL.
 * Supported: D in {2,3}, C in {1,2,4,8}, dtype f32, gridtype 0 hash / 1 tiled,
 * interp 0 linear / 1 smoothstep.  Anything else -> MB_EUNSUPPORTED (the reference throws
 * std::runtime_error, gridencoder.cu:392,409). */
int mb_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                           uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                           void* dy_dx, uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                           mb_stream_t stream);
/* grad [L,B,C]; grad_embeddings [sO,C] pre-zeroed by caller (grid.py:84), accumulated with
 * red.global.add; grad_inputs [B,D] or NULL (written, not accumulated). */
int mb_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                            void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level,
                            float S, uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype,
                            int align_corners, uint32_t interp, int dtype, mb_stream_t stream);

/* ---- (2) ray sampling through a binary occupancy grid ---------------------------------------- */
/* Two passes (count, then write after an exclusive scan by the caller).  rays_o/rays_d [N,3];
 * binaries [res^3] uint8 (x-major: idx = (ix*res+iy)*res+iz, nerfacc layout); aabb[6] host floats;
 * jitter [N] in [0,1) or NULL.  counts [N] i32. */
int mb_sample_rays_count(const float* rays_o, const float* rays_d, uint32_t N, const uint8_t* binaries, uint32_t res,
                         const float* aabb_host6, float step, float near_plane, float far_plane, const float* jitter,
                         int32_t* counts, mb_stream_t stream);
/* offsets [N] i32 exclusive scan of counts; writes ray_indices [M] i64, t_starts [M], t_ends [M]. */
int mb_sample_rays_write(const float* rays_o, const float* rays_d, uint32_t N, const uint8_t* binaries, uint32_t res,
                         const float* aabb_host6, float step, float near_plane, float far_plane, const float* jitter,
                         const int32_t* offsets, int64_t* ray_indices, float* t_starts, float* t_ends,
                         mb_stream_t stream);
/* fixed-S lattice over the AABB chord (BASELINE cfg-1/2/4 synthetic sampler): writes [N*S] triples. */
int mb_sample_rays_uniform(const float* rays_o, const float* rays_d, uint32_t N, uint32_t S, const float* aabb_host6,
                           const float* jitter, int64_t* ray_indices, float* t_starts, float* t_ends,
                           mb_stream_t stream);

/* ---- (3) compositing -------------------------------------------------------------------------- */
/* Packed samples sorted by ray; seg [N+1] i32 = first sample of each ray (seg[N] = M).
 * sigmas/t_starts/t_ends [M]; rgbs [M,3] or NULL.  Outputs (any may be NULL): weights/trans/alphas [M],
 * opacity [N], depth [N] (sum w*(t0+t1)/2), rgb [N,3]. */
int mb_composite_forward(const int32_t* seg, uint32_t N, uint32_t M, const float* sigmas, const float* t_starts,
                         const float* t_ends, const float* rgbs, float* weights, float* trans, float* alphas,
                         float* opacity, float* depth, float* rgb, mb_stream_t stream);
/* Upstream grads (any may be NULL = zero): g_weights [M], g_opacity [N], g_depth [N], g_rgb [N,3].
 * Writes g_sigmas [M] and g_rgbs [M,3] (if rgbs != NULL). */
int mb_composite_backward(const int32_t* seg, uint32_t N, uint32_t M, const float* sigmas, const float* t_starts,
                          const float* t_ends, const float* rgbs, const float* g_weights, const float* g_opacity,
                          const float* g_depth, const float* g_rgb, float* g_sigmas, float* g_rgbs,
                          mb_stream_t stream);

/* ---- (4) fused scene-field query ---------------------------------------------------------------- */
#define MB_F_WARP 1u          /* deform/topo MLPs from (x, t)                       model.py:412-437 */
#define MB_F_MAIN 2u          /* sdf (+sigma) at x+deform                           model.py:273-293 */
#define MB_F_COLOR 4u         /* colour grid + colour MLP (needs MAIN)              model.py:295-302 */
#define MB_F_FD 8u            /* 6-point finite-difference normal                   model.py:367-398 */
#define MB_F_FD_WARPED 16u    /* ... evaluated at x+deform instead of x             model.py:391-395 */
#define MB_F_TOPO_IN 32u      /* take topo from topo_in instead of zeros / warp */
#define MB_F_SKIP_WARP_BWD 64u /* backward only: do not back-propagate the deform/topology nets; write d/d(deform), d/d(topo)
                                 to g_def_out / g_topo_out instead (consumed by mb_field_backward_warp_tc) */
#define MB_F_FD_DELEGATE 128u  /* backward only (mb_field_backward_sdf_tc): do not run the FD-query chains; write d/d(sdf) of the six
                                 +-eps queries of every sample to g_fd [M,6] instead (consumed by mb_field_backward_fd_tc) */
#define MB_SHADE_ALBEDO 0
#define MB_SHADE_LAMBERTIAN 1   /* albedo*(ratio+(1-ratio)*max(n.l,0)); 'albedo_normal' is ratio=1 */
#define MB_SHADE_TEXTURELESS 2
#define MB_SHADE_NORMAL 3

/* Offsets (in floats) into the packed parameter arena, produced by morpheus_b200.packing. Each
 * dense layer stores Wt[K_pad][N_pad] (k-major, forward operand), W[N_pad][K_pad] (n-major,
 * dgrad operand) and bias[N_pad]; the gradient arena uses the same offsets (dW is accumulated
 * into the Wt slot only). */
typedef struct mb_layer_desc {
    uint32_t wt_off, w_off, b_off;
    uint32_t K, N, K_pad, N_pad;
} mb_layer_desc;

typedef struct mb_field_params {
    const float* arena;             /* packed effective weights (weight_norm already applied) */
    mb_layer_desc deform[6], topo[6], sdf[3], color[3];
    const float* emb_sdf;           /* [sO,2] */
    const float* emb_col;           /* [sO,2] */
    const int32_t* offsets;         /* [17] */
    const float* code[3];           /* deform_code.volumes.i as [16, S_i] */
    uint32_t code_len[3];
    const float* beta;              /* device scalar: abs(beta_param)+1e-4 (device-side so no host sync is needed) */
    float bound;                    /* 1.01 */
    float two_bound;                /* (float)(2*bound) evaluated in double first, as Python does (grid.py:157) */
    float S; uint32_t H;            /* log2(per_level_scale), base resolution */
    uint32_t n_levels;              /* ceil(max_level*16) clamped [1,16] */
    uint32_t n_freq;                /* int(max_level*6) */
} mb_field_params;

typedef struct mb_field_io {
    uint32_t M;
    uint32_t flags; int shading; float ratio;
    const float* x;                 /* [M,3] */
    const float* t;                 /* [M]  (needed with WARP) */
    const float* light;             /* [M,3] or NULL */
    const float* topo_in;           /* [M,2] or NULL */
    /* outputs, each may be NULL */
    float* sdf; float* sigma; float* color; float* normal; float* normal_raw; float* deform; float* topo;
} mb_field_io;

int mb_field_forward(const mb_field_params* p, const mb_field_io* io, mb_stream_t stream);

/* Tensor-core engine (tcgen05 / TMEM): same contract as mb_field_forward.  `tc_weights` is produced by mb_pack_tc from
 * the fp32 arena: per layer, fp16 (hi, lo) K=16 slabs in the UMMA canonical shared-memory byte order;
 * layer_desc: n_layers x 8 u32 {w_off, K, K_pad, N_pad, kind, dst_off, K_tc, 0}; tc_off: n_layers x 3 u32
 * {dst_off, K_tc/16, N_pad} (both tables device-resident, built by morpheus_b200.packing.tc_tables). */
int mb_pack_tc(const float* arena, const uint32_t* layer_desc, int n_layers, void* out, mb_stream_t stream);
/* stash (nullable): [ceil(M/128)][10][65536] bytes; with WARP the hidden activations of the deform / topology nets are
 * stored there (fp16 hi/lo operand tiles) for mb_field_backward_warp_tc */
int mb_field_forward_tc(const mb_field_params* p, const mb_field_io* io, const void* tc_weights, const uint32_t* tc_off,
                        void* stash, mb_stream_t stream);

typedef struct mb_field_grads {
    /* upstream (NULL = zero) */
    const float* g_sdf; const float* g_sigma; const float* g_color; const float* g_normal; const float* g_normal_raw; const float* g_deform; const float* g_topo;
    /* saved from forward: deform [M,3], topo [M,2] (required with WARP); normal_raw [M,3] (required with FD) */
    const float* deform; const float* topo; const float* normal_raw;
    /* accumulated (caller zero-fills): */
    float* g_arena; float* g_emb_sdf; float* g_emb_col; float* g_code[3]; float* g_beta;
    /* written: */
    float* g_x;                     /* [M,3] or NULL */
    float* g_topo_in;               /* [M,2] or NULL */
    float* g_def_out;               /* [M,3], with MB_F_SKIP_WARP_BWD */
    float* g_topo_out;              /* [M,2], with MB_F_SKIP_WARP_BWD */
    float* g_fd;                    /* [M,6] or NULL: written with MB_F_FD_DELEGATE, read by mb_field_backward_fd_tc */
} mb_field_grads;

int mb_field_backward(const mb_field_params* p, const mb_field_io* io, const mb_field_grads* g, mb_stream_t stream);

/* Tensor-core backward of the deformation + topology networks (12 dense layers): consumes the activation stash of
 * mb_field_forward_tc, d/d(deform) [M,3] and d/d(topo) [M,2] (from mb_field_backward with MB_F_SKIP_WARP_BWD) and
 * accumulates weight/bias gradients into g_arena, code-line gradients into g_code, and ADDS d/dx into g_x [M,3].
 * tc_weights_t / tc_off_t: dgrad operands packed by mb_pack_tc in mode 1 (12 layers: deform[6], topo[6]). */
/* Tensor-core backward of the SDF + colour networks, Laplace density, shading and FD-normal queries (everything of
 * mb_field_backward except the deform/topology nets; with WARP it requires MB_F_SKIP_WARP_BWD and writes g_def_out /
 * g_topo_out for mb_field_backward_warp_tc).  tc_weights/tc_off: forward tables of mb_field_forward_tc;
 * tc_weights_t/tc_off_t: dgrad operands of sdf[3], color[3] (mb_pack_tc mode 1). */
int mb_field_backward_sdf_tc(const mb_field_params* p, const mb_field_io* io, const mb_field_grads* g, const void* tc_weights,
                             const uint32_t* tc_off, const void* tc_weights_t, const uint32_t* tc_off_t, mb_stream_t stream);
/* Tensor-core backward of the FD-normal queries only (two CTAs per SM, 96-row sub-tiles = 16 samples x 6 queries).  Upstream:
 * g->g_fd [M,6] (from mb_field_backward_sdf_tc with MB_F_FD_DELEGATE) or, if NULL, g_normal / g_normal_raw with the saved
 * normal_raw (ALBEDO shading).  accumulate != 0: ADD to g_x / g_def_out / g_topo_out / g_topo_in (which the main kernel wrote);
 * accumulate == 0: write them (stand-alone FD query, e.g. scene_representation.normal). */
int mb_field_backward_fd_tc(const mb_field_params* p, const mb_field_io* io, const mb_field_grads* g, const void* tc_weights,
                            const uint32_t* tc_off, const void* tc_weights_t, const uint32_t* tc_off_t, int accumulate, mb_stream_t stream);
/* debug: cumulative clock64 cycles per phase of mb_field_backward_fd_tc, summed over CTAs (16 host words); reset != 0 clears */
int mb_debug_fd_phases(unsigned long long* host_out16, int reset);
int mb_debug_fdr_phases(unsigned long long* host_out16, int reset);   /* same for mb_fd_regulariser_tc (build with -DMB_FDR_PHASE_TIMING=1) */
/* Fused finite-difference normal regulariser of a real-view step: loss_normal_perturb of MorpheuS.render_rays (morpheus.py:714-741) =
 * mean |n(x, topo) - n(x + noise * noise_std, topo = 0)| with n = safe_normalize(6-point FD of the SDF, models/model.py:367-398), forward
 * AND backward in one launch (csrc/field_fd_reg_tc.cu).  Valid when the colour does not depend on the normal ('albedo_normal', ratio 1).
 *   loss[0]   += gmul * sum_{samples, axes} |n - n_p|                    (caller zero-fills; gmul = 1 / (3 M) gives the mean)
 *   normal / normal_raw [M,3]  the unperturbed set's normals (nullable)
 *   g_x [M,3], g_topo [M,2]    d loss / d x, d loss / d topo (written; g_topo required iff topo != NULL)
 *   g_emb_sdf, g_arena         d loss / d (SDF hash table), d loss / d (packed arena: sdf layers) ACCUMULATED (red.global.add)
 * noise may be NULL (zeros).  tc_* as for mb_field_backward_fd_tc. */
int mb_fd_regulariser_tc(const mb_field_params* p, const float* x, const float* topo, const float* noise, float noise_std, uint32_t M,
                         float gmul, float* normal, float* normal_raw, float* loss, float* g_x, float* g_topo, float* g_emb_sdf,
                         float* g_arena, const void* tc_weights, const uint32_t* tc_off, const void* tc_weights_t, const uint32_t* tc_off_t,
                         mb_stream_t stream);

int mb_field_backward_warp_tc(const mb_field_params* p, const float* x, const float* t, uint32_t M, const float* g_def,
                              const float* g_topo, const void* stash, const void* tc_weights_t, const uint32_t* tc_off_t,
                              float* g_arena, float* const g_code[3], float* g_x, mb_stream_t stream);

/* ---- (5) occupancy refresh: occs = max(decay*occs, sigma*step) on selected cells ---------------- */
int mb_occ_update(float* occs, const int64_t* cell_idx, const float* sigma, uint32_t n, float decay, float step,
                  mb_stream_t stream);
int mb_occ_binarize(const float* occs, uint32_t n, float thre, uint8_t* binaries, mb_stream_t stream);
/* same with the threshold read from device memory (thre_dev[0]): the refresh then never synchronises with the host */
int mb_occ_binarize_dev(const float* occs, uint32_t n, const float* thre_dev, uint8_t* binaries, mb_stream_t stream);

/* ---- (6) fused Adam over a flat arena ------------------------------------------------------------ */
/* p,g,m,v [n]; lr_scale [n_groups] device floats indexed by group_id [n] u8 (per-parameter-group lr,
 * models/model.py:313-324); bias-corrected torch.optim.Adam semantics, no weight decay. */
int mb_adam_step(float* p, const float* g, float* m, float* v, const uint8_t* group_id, const float* group_lr,
                 uint64_t n, float beta1, float beta2, float eps, int step, mb_stream_t stream);
/* same, with the (1-based) step count read from device memory so the launch can be replayed inside a CUDA graph */
int mb_adam_step_dev(float* p, const float* g, float* m, float* v, const uint8_t* group_id, const float* group_lr,
                     uint64_t n, float beta1, float beta2, float eps, const int32_t* step_dev, mb_stream_t stream);

/* torch.optim.Adam's skip-if-grad-is-None semantics per parameter group: group_active [n_groups] u8 (0 = this step produced no gradient
 * for the group: its p/m/v and its step count stay untouched), group_step [n_groups] i32 device counters (incremented here for the
 * active groups; bias corrections use the group's own count, like torch's per-parameter `step`).  zero_after != 0 clears g after it
 * was read (the next step's zero_grad() folded into this launch).  n_groups <= 32.  Graph-replayable. */
int mb_adam_step_groups(float* p, float* g, float* m, float* v, const uint8_t* group_id, const float* group_lr,
                        const uint8_t* group_active, int32_t* group_step, int n_groups, uint64_t n, float beta1, float beta2,
                        float eps, int zero_after, mb_stream_t stream);

/* ---- (7) SDS scalar chain: grad = grad_scale*(1-abar_t)*(eps_u + s*(eps_c-eps_u) - eps), nan_to_num --- */
int mb_sds_grad(const float* eps_uncond, const float* eps_cond, const float* noise, float guidance_scale,
                float w_t_times_grad_scale, float* grad, uint32_t n, mb_stream_t stream);
/* latents_noisy = sqrt(abar)*z + sqrt(1-abar)*eps */
int mb_add_noise(const float* z, const float* eps, float sqrt_abar, float sqrt_one_minus_abar, float* out, uint32_t n,
                 mb_stream_t stream);

/* device-scalar variants (no host synchronisation, CUDA-graph capturable): t_dev = the diffusion step (int64 [1] on the device),
 * alphas_cumprod [1000] on the device; latents_noisy = sqrt(abar_t) z + sqrt(1 - abar_t) eps  (DDIMScheduler.add_noise, zero123_utils.py:180) */
int mb_add_noise_dev(const float* z, const float* eps, const float* alphas_cumprod, const int64_t* t_dev, float* out, uint32_t n,
                     mb_stream_t stream);
/* grad (+)= view_weight * grad_scale_dev[0] * (1 - abar_t) * (eps_u + s (eps_c - eps_u) - eps)   (zero123_utils.py:204-212; the caller applies
 * nan_to_num once after the weighted sum over the reference views) */
int mb_sds_grad_dev(const float* eps_uncond, const float* eps_cond, const float* noise, float guidance_scale, const float* grad_scale_dev,
                    const float* alphas_cumprod, const int64_t* t_dev, float view_weight, float* grad, uint32_t n, int accumulate,
                    mb_stream_t stream);

/* ---- (7b) tensor-core convolution for the frozen SDS networks (csrc/conv_tc.cu) --------------------------------------------------
 * 3x3 (stride 1, pad 1; ntaps = 9) or 1x1 (ntaps = 1) convolution as an implicit GEMM on tcgen05 with the 3-term fp16 split (fp32-grade
 * accuracy).  Replaces F.conv2d of ResBlock / ResnetBlock / proj_in / proj_out (ldm/modules/diffusionmodules/openaimodel.py:256-276,
 * model.py:82-140, ldm/modules/attention.py:239-253) in strict-fp32 mode; the VAE input-gradient backward is the same kernel on weights
 * packed with transposed != 0.
 *   mb_conv_pack_weights: w fp32 [Cout, Cin, kh, kw] -> out (Cout * Cin * ntaps * 4 bytes), n_tile in {128, 160} must divide the packed
 *                         operator's rows (Cout, or Cin when transposed), 64 its contracted channels
 *   mb_nchw_split:        x fp32 [B, C, HW] (NCHW) -> hi, lo fp16 [B, HW, C] (NHWC); act = 1 applies SiLU first
 *   mb_conv_tc:           out fp32 [B, Cout, H, W] = conv(x) + bias; nsplit > 1 splits K over gridDim.z and ACCUMULATES with
 *                         red.global.add (caller zero-fills out)
 * Scaling (fp16 range): weights are multiplied by the power of two `wscale` at pack time, activations by scale_dev[0] (device scalar,
 * nullable) in the split; the epilogue multiplies the accumulator by out_mul * out_mul_dev[0] (= 1 / wscale, 1 / activation scale). */
int mb_conv_pack_weights(const float* w, int Cout, int Cin, int ntaps, int n_tile, int transposed, float wscale, void* out, mb_stream_t stream);
int mb_nchw_split(const float* x, int B, int C, int HW, int act, const float* scale_dev, void* hi, void* lo, mb_stream_t stream);
int mb_conv_tc(const void* x_hi, const void* x_lo, const void* w_packed, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
               int ntaps, int n_tile, int nsplit, float out_mul, const float* out_mul_dev, int out_rows, mb_stream_t stream);
/* linear layers of the UNet's transformer blocks (ldm/modules/attention.py:37-64, :152-193) = 1x1 convolution over tokens: x [rows, C] fp32
 * (row-major) -> hi, lo fp16 of the same layout; mb_conv_tc with B = 1, H = rows, W = 1, ntaps = 1, out_rows = 1 writes out [rows, Cout] */
int mb_rows_split(const float* x, uint64_t n_elems, const float* scale_dev, void* hi, void* lo, mb_stream_t stream);

/* ---- (8) host-glue kernels of the step (each replaces tens to hundreds of eager launches) ------------------------- */
/* xyz[i] = rays_o[ray_indices[i]] + rays_d[ray_indices[i]] * (t_starts[i]+t_ends[i])/2        morpheus.py:645-646 */
int mb_ray_points_forward(const float* rays_o, const float* rays_d, const int64_t* ray_indices, const float* t_starts,
                          const float* t_ends, uint32_t M, float* xyz, mb_stream_t stream);
/* g_o[r] = sum_i g_xyz[i], g_d[r] = sum_i g_xyz[i]*t_mid[i] over the packed samples seg[r]..seg[r+1] of ray r (either may be NULL) */
int mb_ray_points_backward(const int32_t* seg, uint32_t N, const float* t_starts, const float* t_ends, const float* g_xyz,
                           float* g_o, float* g_d, mb_stream_t stream);
/* weight_norm + transpose + pad of the dense layers into the parameter arena (models/decoders.py:51-52 semantics).
 * layer_table: n_layers x 8 int64 {weight_v|weight ptr, weight_g ptr or 0, bias ptr, K, N, K_pad, N_pad, wt_off} (device);
 * backward writes d/d(weight_v|weight), d/d(weight_g), d/d(bias) into flat_grads at the float offsets of
 * grad_table: n_layers x 4 int64 {gv_off, gg_off, gb_off, 0}; with flat_grads == NULL the table holds absolute device
 * addresses (the parameters' .grad tensors) and the gradients are ACCUMULATED there. */
int mb_pack_arena_forward(const int64_t* layer_table, int n_layers, float* arena, mb_stream_t stream);
int mb_pack_arena_backward(const int64_t* layer_table, const int64_t* grad_table, int n_layers, const float* g_arena,
                           float* flat_grads, mb_stream_t stream);
/* utils.get_sdf_loss (utils.py:91-113) on packed samples: out2[0] += sum fs_i/n_i, out2[1] += sum |s_i-bound_i| band_i/n_i
 * (caller zero-fills out2 and divides by count_nonzero(depth)); backward: g_sdf[i] = g_out2[0]*dfs_i + g_out2[1]*dsl_i. */
int mb_sdf_loss_forward(const float* t_starts, const float* t_ends, const int64_t* ray_indices, const float* depth, const float* mask,
                        const float* sdf, uint32_t M, float truncation, float* out2, mb_stream_t stream);
int mb_sdf_loss_backward(const float* t_starts, const float* t_ends, const int64_t* ray_indices, const float* depth, const float* mask,
                         const float* sdf, uint32_t M, float truncation, const float* g_out2, float* g_sdf, mb_stream_t stream);

/* pose correction of a ray batch (models/model.py:335-346, models/pose.py:35-58): pose [F,6] = (Euler a,b,g, translation);
 * rays_o_out = rays_o + t_f, rays_d_out = R_f rays_d.  Backward ACCUMULATES into g_pose [F,6] (caller zero-fills) and writes
 * g_rays_o / g_rays_d (nullable). */
int mb_pose_rays_forward(const float* pose, const int64_t* frame_ids, const float* rays_o, const float* rays_d, uint32_t N,
                         float* rays_o_out, float* rays_d_out, mb_stream_t stream);
int mb_pose_rays_backward(const float* pose, const int64_t* frame_ids, const float* rays_d, const float* g_o_out, const float* g_d_out,
                          uint32_t N, float* g_pose, float* g_rays_o, float* g_rays_d, mb_stream_t stream);
/* per-ray loss heads of a real view (morpheus.py:946-983): out1[0] += w_rgb*mse(image,gt_rgb) + w_mask*bce(clip(opacity),gt_mask)
 * + w_depth*mse(depth*dm, gt_depth*dm); also writes the per-ray gradients of that scalar (g_image [N,3], g_opacity [N], g_depth [N]). */
int mb_ray_loss(const float* image, const float* opacity, const float* depth, const float* gt_rgb, const float* gt_depth,
                const float* gt_mask, const float* rays_o, const float* rays_d, uint32_t N, float w_rgb, float w_mask, float w_depth,
                float* out1, float* g_image, float* g_opacity, float* g_depth, mb_stream_t stream);

/* deformation-code regulariser (morpheus.py:762-771): out1[0] = mean_c (2 c(t) - c(t-1/F) - c(t+1/F))^2 over the 48 channels of the three
 * code lines code[v] [16, code_len[v]] (t_dev: device scalar).  g_out1 != NULL: backward, g_code[v] += g_out1[0] * d(out1)/d(code[v]). */
int mb_code_reg(const float* const code[3], const int code_len[3], const float* t_dev, float inv_frames, float* out1, const float* g_out1,
                float* const g_code[3], mb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MORPHEUS_B200_H */
