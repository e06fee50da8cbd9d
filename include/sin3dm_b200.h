/*
 * sin3dm_b200 — C ABI of the B200-native triplane denoising path.
 *
 * The reference (Sin3DM) has no FFI layer: its boundary for this path is two Python classes,
 *   TriplaneUNetModelSmall / ...SmallRaw   (reference src/diffusion/unet_triplane.py:315-510, 513-702)
 *   GaussianDiffusion / SpacedDiffusion     (reference src/diffusion/gaussian_diffusion.py:102-947,
 *                                            src/diffusion/respace.py:63-128)
 * The entry points below are what a ctypes binding of those classes needs (INTEGRATION.md shows the stub);
 * `sin3dm_b200/` is that binding.  Plain pointers and sizes only — no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; s3d_last_error() gives the message
 *     (thread-local).  No call synchronises the host with the device unless stated.
 *   - `*_dev` pointers are device pointers BORROWED from the caller (contiguous fp32, NCHW at the
 *     boundary: [B, C, H+D, W+D] "composed" triplane, reference src/utils/triplane_util.py:7-25).
 *     Weights are copied / re-packed into memory owned by the handle.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef SIN3DM_B200_H
#define SIN3DM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_MAX_LEVELS 8
#define S3D_NCOEF 12

typedef struct s3d_unet s3d_unet;

/* Constructor arguments of the reference UNet (unet_triplane.py:346-357) + implementation knobs. */
typedef struct {
    int in_channels;
    int model_channels;            /* must be a multiple of 64 for the tcgen05 path */
    int out_channels;
    int num_res_blocks;            /* the reference only runs with 1 */
    int n_levels;
    int channel_mult[S3D_MAX_LEVELS];
    int use_scale_shift_norm;      /* FiLM (unet_triplane.py:285-297) vs additive embedding */
    int rollout;                   /* 1: TriplaneUNetModelSmall, 0: ...SmallRaw */
    int precision;                 /* conv operand terms (v = hi + lo/2048, fp16 halves): 3: Ah*Bh + Ah*Bl + Al*Bh (fp32-grade);
                                      2: Ah*(Bh + Bl) (exact weights, fp16 activations); 4: (Ah + Al)*Bh (exact activations, fp16
                                      weights); 1: Ah*Bh; 5: mode 2 for the conv tiles with mode 3 for the rollout 1-D GEMM tiles.  DESIGN.md §3 has the measured
                                      full-chain error of each. */
    int conv_impl;                 /* 0: tcgen05 implicit GEMM; 1: CUDA-core debug kernel */
} s3d_unet_config;

int s3d_abi_version(void);
const char* s3d_last_error(void);

/* ---- UNet (replaces TriplaneUNetModelSmall[Raw].__init__/load_state_dict/forward) ---- */
int s3d_unet_create(const s3d_unet_config* cfg, int device, s3d_unet** out);
int s3d_unet_destroy(s3d_unet* u);
/* Expected checkpoint tensors, in reference state_dict() order. */
int s3d_unet_num_tensors(const s3d_unet* u);
int s3d_unet_tensor_info(const s3d_unet* u, int index, const char** name, int* ndim, int64_t shape[4]);
/* Copy one fp32 tensor (host memory, contiguous, reference layout) by its state_dict key. */
int s3d_unet_load_tensor(s3d_unet* u, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* Re-pack everything loaded so far for the kernels (fp16 hi/lo K-major conv operands, folded rollout
 * weights, concatenated FiLM projection).  Fails if a tensor is missing.  Synchronises the device. */
int s3d_unet_finalize(s3d_unet* u);
/* Row width of the per-timestep conditioning vector produced by the embedding MLPs
 * (time_embed + every block's emb_layers, unet_triplane.py:232-238, 371-375). */
int s3d_unet_film_dim(const s3d_unet* u);
/* film[n][film_dim] for n timesteps (fp32 values, already mapped through timestep_map / rescaled). */
int s3d_unet_film(s3d_unet* u, const float* t_dev, int n, float* film_dev, void* stream);
/* forward(x, timesteps, H, W, D) -> out (same shape as x with out_channels). */
int s3d_unet_forward(s3d_unet* u, const float* x_dev, const float* t_dev, float* out_dev, int B, int H, int W, int D,
                     void* stream);
/* Same, with the conditioning rows precomputed: sample b uses film_dev[row_dev ? row_dev[b] : b]. */
int s3d_unet_forward_film(s3d_unet* u, const float* x_dev, const float* film_dev, const int* row_dev, float* out_dev,
                          int B, int H, int W, int D, void* stream);
/* Kernel launches issued by the last forward (for bench accounting). */
int s3d_unet_last_launches(const s3d_unet* u);
/* Bytes of device workspace (activations, operands, statistics) owned by the current launch plan. */
int64_t s3d_unet_workspace_bytes(const s3d_unet* u);
/* Launch list of the current plan: kernel name and the dense algorithmic FLOPs the op stands for (convs). */
int s3d_unet_op_count(const s3d_unet* u);
int s3d_unet_op_info(const s3d_unet* u, int index, const char** kernel, double* dense_flops);
/* Mean device time (ms) of every op of the plan; synchronises.  The step is captured into one CUDA graph with an
 * event-record node before every launch and replayed `iters` times on `stream` (steady-state times, as the sampling
 * loop sees them); if the driver refuses that, falls back to eager launches with events around each.
 * Uses the tensors bound by the last forward / sampling loop (bench.py's roofline leg). */
int s3d_unet_profile_ops(s3d_unet* u, int iters, float* ms_out, void* stream);
/* How the last s3d_unet_profile_ops measured: 1 = graph replay, 0 = eager launches, -1 = never ran. */
int s3d_unet_profile_mode(const s3d_unet* u);

/* Opt-in phase trace (bring-up / profiling): when enabled, one thread per CTA of every kernel of the plan stamps
 * %globaltimer (ns) and clock64 at s3d_trace_slots() named points.  s3d_unet_trace_read copies the stamps of op
 * `op_index` (-1: the scheduler kernel of s3d_sample_loop) as [max_ctas][2][slots] (0 = never stamped), then clears
 * them; it synchronises.  Enabling / disabling drops the current launch plan.  tools/trace_step.py prints a timeline. */
int s3d_unet_trace_enable(s3d_unet* u, int on);
int s3d_trace_slots(void);
int s3d_unet_trace_read(s3d_unet* u, int op_index, uint64_t* host_out, int max_ctas);

/* ---- scheduler step (replaces p_sample / ddim_sample / ddim_reverse_sample element-wise math) ----
 * coef_dev is [T][S3D_NCOEF] fp32, built by the host mirror from the fp64 tables exactly as
 * _extract_into_tensor rounds them (gaussian_diffusion.py:934-947):
 *   0 sqrt_recip_alphas_cumprod   1 sqrt_recipm1_alphas_cumprod   2 posterior_mean_coef1
 *   3 posterior_mean_coef2        4 exp(0.5*model_log_variance)    5 sqrt(alpha_bar_prev)
 *   6 sqrt(1-alpha_bar_prev-sigma^2)  7 ddim sigma (eta folded)   8 sqrt(alpha_bar_next)
 *   9 sqrt(1-alpha_bar_next)     10 sqrt_alphas_cumprod           11 sqrt_one_minus_alphas_cumprod */
enum { S3D_DDPM = 0, S3D_DDIM = 1, S3D_DDIM_REVERSE = 2 };
enum { S3D_START_X = 0, S3D_EPSILON = 1 };

typedef struct {
    int kind;               /* S3D_DDPM / S3D_DDIM / S3D_DDIM_REVERSE */
    int mean_type;          /* S3D_START_X / S3D_EPSILON */
    int clip_denoised;
    int is_mask_t0;
    int B;
    int C;                  /* channels of the composed tensor */
    int64_t n_per_sample;   /* C*(H+D)*(W+D) */
    const float* model_out; /* [B, n] */
    const float* x;         /* [B, n] x_t */
    const float* noise;     /* [B, n] or NULL -> Philox4x32-10: element (c, pixel) = component c&3 of the block with counter
                               (pixel*ceil(C/4) + c/4, t_idx[b], sample_base+b), key = seed  (oracle/philox_ref.py) */
    const float* y0;        /* optional inpainting target (ddim_sample y0/mask), NULL if unused */
    const float* mask;
    float* sample;          /* [B, n] may alias x */
    float* pred_xstart;     /* [B, n] or NULL */
    const float* coef_dev;  /* [T][S3D_NCOEF] */
    const int* t_idx_dev;   /* [B] step index into coef_dev */
    uint64_t seed;
    uint32_t sample_base;
} s3d_sched_args;

int s3d_sched_step(const s3d_sched_args* a, void* stream);
/* q_sample: x_t = coef[t][10]*x0 + coef[t][11]*noise (gaussian_diffusion.py:189-207). */
int s3d_q_sample(const float* x0_dev, const float* noise_dev, float* out_dev, const float* coef_dev, const int* t_idx_dev,
                 int B, int64_t n_per_sample, void* stream);
/* ---- variational-bound terms (replaces GaussianDiffusion._vb_terms_bpd, gaussian_diffusion.py:736-769, and the two MSEs
 * calc_bpd_loop adds per step, :912-916; normal_kl / discretized_gaussian_log_likelihood are src/diffusion/losses.py:12-77).
 * One fused pass + a fixed-order fp64 reduction: out[b] = { KL(q(x_{t-1}|x_t,x_0) || p(x_{t-1}|x_t)) in bits per dimension,
 * or the discretised decoder NLL where t_idx[b] == 0; mean (pred_xstart - x_start)^2; mean (eps - noise)^2 (0 if noise == NULL) }.
 * logvar_dev is [T][2] fp32: posterior_log_variance_clipped, model log-variance (FIXED_LARGE / FIXED_SMALL table). */
typedef struct {
    int mean_type;          /* S3D_START_X / S3D_EPSILON */
    int clip_denoised;
    int B;
    int64_t n_per_sample;
    const float* x_start;   /* [B, n] */
    const float* x_t;       /* [B, n] */
    const float* model_out; /* [B, n] */
    const float* noise;     /* [B, n] the noise x_t was drawn with, or NULL */
    float* pred_xstart;     /* [B, n] or NULL */
    const float* coef_dev;  /* [T][S3D_NCOEF] */
    const float* logvar_dev;
    const int* t_idx_dev;   /* [B] */
    void* workspace;        /* s3d_vb_workspace_bytes(B, n) bytes of device memory */
    float* out;             /* [B][3] */
} s3d_vb_args;
int64_t s3d_vb_workspace_bytes(int B, int64_t n_per_sample);
int s3d_vb_terms(const s3d_vb_args* a, void* stream);
/* Per-plane MSE of the training objective (GaussianDiffusion.training_losses, gaussian_diffusion.py:822-851: mean_flat of
 * (target - output)^2 over the xy, xz and yz planes of the composed tensors [B, C, H+D, W+D]): out_dev [B][3] = mse_xy, mse_xz,
 * mse_yz.  One pass, fixed-order fp64 reduction.  workspace: s3d_vb_workspace_bytes(B, C*(H+D)*(W+D)) bytes. */
int s3d_plane_mse(const float* target_dev, const float* output_dev, int B, int C, int H, int W, int D, void* workspace, float* out_dev,
                  void* stream);
/* ---- fused optimizer step (replaces torch.optim.AdamW.step() + update_ema of TrainLoop.run_step, train_util.py:82-84, 160-167,
 * 237-239; nn.py:53-63) over ONE flat fp32 parameter buffer: decoupled weight decay, Adam moments with bias correction, parameter
 * update, then ema[k] = ema[k] * ema_rate[k] + param * (1 - ema_rate[k]) for up to four EMA copies — a single HBM pass.
 * `step` is the 1-based count of this update; lr is the (possibly annealed) rate of this step.  All buffers 16-byte aligned. */
typedef struct {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    float* ema[4];
    float ema_rate[4];
    int n_ema;
    int64_t n;
    double lr, beta1, beta2, eps, weight_decay;   /* python floats, as torch.optim receives them */
    int step;
} s3d_adamw_args;
int s3d_adamw_ema_step(const s3d_adamw_args* a, void* stream);
/* N(0,1) fill [B, C, hw] with the same counter-based generator the sampler uses (noise of sample s at step i). */
int s3d_philox_normal(float* out_dev, int B, int C, int64_t hw, uint64_t seed, uint32_t sample_base, uint32_t step,
                      void* stream);

/* ---- whole sampling loop (replaces p_sample_loop / ddim_sample_loop, gaussian_diffusion.py:442-536, 640-734) ----
 * Runs i = n_steps-1 .. 0 on the device: UNet forward + scheduler step per iteration, captured once as a
 * CUDA graph and replayed; the step index lives in device memory, so there is no per-step host work. */
typedef struct {
    int kind;               /* S3D_DDPM / S3D_DDIM */
    int mean_type;
    int clip_denoised;
    int is_mask_t0;
    int n_steps;            /* iterations to run */
    int t_start;            /* step index of the first iteration (the loop runs t_start, t_start-1, ...; a full chain passes the
                               table length - 1); needs t_start - n_steps + 1 >= 0 */
    int B, H, W, D;
    int64_t n_per_sample;   /* elements of one sample of x_dev; must equal out_channels * (H+D) * (W+D) (checked) */
    float* x_dev;           /* in: x_T, out: final sample (in place) [B, C, H+D, W+D] */
    float* pred_xstart_dev; /* optional */
    const float* coef_dev;  /* [T][S3D_NCOEF], T > t_start */
    const float* film_dev;  /* [T][film_dim] conditioning of step index i (already timestep_map'ed) */
    const float* step_noise_dev; /* NULL -> Philox; else [T][B][n] with row i used at step index i */
    const float* y0_dev;
    const float* mask_dev;
    uint64_t seed;
    uint32_t sample_base;   /* global index of sample 0 of this batch (multi-GPU sharding) */
    int use_graph;          /* 1: CUDA graph replay, 0: plain launches */
} s3d_loop_args;

/* The captured graph is cached per (sampler options, buffer pointers); seed, sample_base and t_start live in device memory, so
 * repeated calls with the same buffers replay the cached graph without re-capturing. */
int s3d_sample_loop(s3d_unet* u, const s3d_loop_args* a, void* stream);
/* CUDA graphs captured + instantiated by s3d_sample_loop since the handle was created (bench.py asserts that none is built inside
 * its timed region). */
int s3d_unet_graph_builds(const s3d_unet* u);


/* ---- training backward (replaces loss.backward() through TriplaneUNetModelSmall[Raw] in TrainLoop.forward_backward, reference
 *      src/diffusion/train_util.py:198-235; the forward being differentiated is src/diffusion/unet_triplane.py:465-510) ----
 * s3d_unet_set_training(u, 1): the next forward builds a plan that keeps every activation and statistics buffer and carries the
 * backward op list (the sampling loop is refused in this mode).  s3d_unet_backward differentiates the LAST s3d_unet_forward[_film]:
 *   grad_out_dev : dL/d(out), [B, out_channels, H+D, W+D] fp32 (the composed layout of the forward's output)
 *   grads_dev    : out, flat fp32 buffer of s3d_unet_grad_numel() floats: the gradient of checkpoint tensor i (reference layout,
 *                  s3d_unet_tensor_info order) starts at s3d_unet_grad_offset(u, i).  The time_embed.* / emb_layers.* slots are left
 *                  zero: their gradient is returned through dfilm_dev instead —
 *   dfilm_dev    : out (optional), [B][film_dim]: dL/d(conditioning row) of every sample, to be pushed through the embedding MLP
 *                  (time_embed + emb_layers: a [B, 256] GEMV chain the host mirror owns).
 * x (the forward's input) and the conditioning rows must still be alive.  No gradient w.r.t. x is produced (TrainLoop does not
 * need one).  Gradients are carried multiplied by a device-chosen power-of-two loss scale and un-scaled at the end. */
int s3d_unet_set_training(s3d_unet* u, int on);
/* Re-pack every kernel operand from checkpoint tensors that are already on the device: tensors_dev[i] = device pointer of tensor i
 * (s3d_unet_tensor_info order, n = s3d_unet_num_tensors without the "__freqs" pseudo tensor), fp32, contiguous, reference layout.
 * The device-side equivalent of s3d_unet_load_tensor x n + s3d_unet_finalize (same packed values), for the optimizer step of a
 * training loop: no host round trip, no synchronisation.  Needs one host-side load + finalize before (buffer allocation). */
int s3d_unet_refresh_dev(s3d_unet* u, const float* const* tensors_dev, int n, void* stream);
int64_t s3d_unet_grad_numel(s3d_unet* u);
int64_t s3d_unet_grad_offset(s3d_unet* u, int index);
int s3d_unet_backward(s3d_unet* u, const float* grad_out_dev, float* grads_dev, float* dfilm_dev, void* stream);
/* Launch list of the backward (kernel name, dense algorithmic FLOPs) and its per-op device times (bench.py --workload cfg4). */
int s3d_unet_bwd_op_count(const s3d_unet* u);
int s3d_unet_bwd_op_info(const s3d_unet* u, int index, const char** kernel, double* dense_flops);
int s3d_unet_profile_bwd_ops(s3d_unet* u, int iters, float* ms_out, void* stream);

/* ---- triplane decoder (replaces AutoEncoderGroupSkip.decode and ShapeAutoEncoder.decode_batch / decode_grid:
 *      reference src/encoding/networks.py:134-223, src/encoding/model.py:319-349) ----
 * Latent planes in, SDF + texture at query points out.  The two TriplaneGroupResnetBlocks (blocks.py:189-256) depend on
 * the latent only: s3d_decoder_set_planes runs them ONCE and keeps the 64(+64)-channel feature planes on the device; the
 * reference recomputes them for every 16 384-point chunk.  s3d_decoder_decode[_grid] is one fused launch: aabb
 * normalisation -> bilinear gather-sum over the three planes (F.grid_sample border / align_corners=False,
 * networks.py:182-190) -> DecoderMLPSkipConcat (blocks.py:65-91) on tcgen05 -> sigmoid on the texture head (+ the [0,1]
 * clamp of decode_batch, model.py:332). */
typedef struct s3d_decoder s3d_decoder;

/* Constructor arguments of the reference auto-encoder that matter for decode (networks.py:134-136) + implementation knobs. */
typedef struct {
    int geo_feat_channels;         /* fdim_geo (parser_util.py:22), 4 */
    int tex_feat_channels;         /* fdim_tex, 8 */
    int feat_channel_up;           /* fdim_up, must be 64 */
    int mlp_hidden_channels;       /* hidden_dim, 256 for the tensor-core kernel */
    int mlp_hidden_layers;         /* n_hidden_layers, 4 for the tensor-core kernel */
    int use_tex;                   /* data_type != "sdf" */
    int tex_channels;              /* 3 */
    int ks;                        /* kernel size of the TriplaneGroupResnetBlock convs (5, networks.py:152-153) */
    int precision;                 /* 3: fp16 hi/lo split, 3 MMAs (fp32-grade, default); 1: single fp16 MMA */
    int mlp_impl;                  /* 0: tcgen05 fused MLP; 1: CUDA-core fp32 kernel (cross-check / other MLP shapes) */
    int mlp_kind;                  /* 0: DecoderMLPSkipConcat heads (AutoEncoderGroupSkip, enc_net_type "skip", networks.py:134);
                                      1: plain DecoderMLP heads (AutoEncoderGroupV3, enc_net_type "base", networks.py:21, blocks.py:46) */
    int net_kind;                  /* 0: AutoEncoderGroupSkip / V3 layout (one ks x ks block per branch; heads geo(1), tex(tex_channels) + sigmoid);
                                      1: AutoEncoderGroupPBR (enc_net_type "pbr", networks.py:227-331): geo block ks 5, two texture blocks
                                         ks 3 (the second with input InstanceNorm + SiLU and an identity shortcut), heads geo(1), rgb(3),
                                         mr(2), normal(3) on the shared texture planes, no sigmoid; tex_channels must be 8, `ks` is unused */
} s3d_decoder_config;

int s3d_decoder_create(const s3d_decoder_config* cfg, int device, s3d_decoder** out);
int s3d_decoder_destroy(s3d_decoder* d);
/* Expected checkpoint tensors in reference state_dict() order (aabb and the encoder Conv3d weights are accepted and unused). */
int s3d_decoder_num_tensors(const s3d_decoder* d);
int s3d_decoder_tensor_info(const s3d_decoder* d, int index, const char** name, int* ndim, int64_t shape[5]);
int s3d_decoder_load_tensor(s3d_decoder* d, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* Re-pack for the kernels (per-plane conv taps, transposed fp32 MLP weights, power-of-two-scaled fp16 hi/lo MLP weights +
 * TMA descriptors).  Fails if a decode-side tensor is missing.  Synchronises the device. */
int s3d_decoder_finalize(s3d_decoder* d);
/* feat_maps of decode(): xy [1,c,H,W], xz [1,c,H,D], yz [1,c,W,D] fp32 NCHW, c = geo (+ tex) channels. */
int s3d_decoder_set_planes(s3d_decoder* d, const float* xy_dev, const float* xz_dev, const float* yz_dev, int H, int W, int D,
                           void* stream);
/* decode(x, feat_maps, aabb): pts_dev [n,3] -> out_dev [n, 1 (+ tex_channels)].  aabb is HOST memory (6 floats). */
int s3d_decoder_decode(s3d_decoder* d, const float* pts_dev, int64_t n, const float aabb[6], int clamp_tex, float* out_dev,
                       void* stream);
/* decode_grid: the points are the meshgrid (indexing 'ij') of the three coordinate vectors of sample_grid_points_aabb
 * (utils3d.py:13-25), formed inside the kernel; out_dev [nx, ny, nz, 1 (+ tex_channels)]. */
int s3d_decoder_decode_grid(s3d_decoder* d, const float* xs_dev, const float* ys_dev, const float* zs_dev, int nx, int ny, int nz,
                            const float aabb[6], int clamp_tex, float* out_dev, void* stream);
/* Encoder half (replaces AutoEncoderGroupSkip.encode, src/encoding/networks.py:164-180): vol_dev [1 (+ tex_channels), X, Y, Z]
 * fp32 -> the three latent planes xy [C, H, W], xz [C, H, D], yz [C, W, D], C = geo (+ tex) feature channels and
 * (H, W, D) = ((X-2)/2+1, (Y-2)/2+1, (Z-2)/2+1) (Conv3d k 4, stride 2, pad 1).  The strided convolutions, the three axis means
 * (64-bit fixed-point sums: results do not depend on tile order), InstanceNorm2d and tanh(x/2) run in two launches; the
 * [C, H, W, D] feature volume is never written.  Needs the geo_encoder.* (+ tex_encoder.*) tensors loaded.  The reference
 * defaults (fdim_geo 4; sdf only or fdim_tex 8 + rgb) take the specialised kernels, any other configuration (<= 32 latent
 * channels) a generic direct convolution. */
int s3d_decoder_encode(s3d_decoder* d, const float* vol_dev, int X, int Y, int Z, float* xy_dev, float* xz_dev, float* yz_dev,
                       void* stream);
/* Kernel launches issued by the last set_planes / decode / encode call (bench accounting). */
int s3d_decoder_last_launches(const s3d_decoder* d);
/* Bring-up hook: the up-convolved feature plane `plane` as [rows][cols][64 (+64)] fp32 (geo channels first). Synchronises. */
int s3d_decoder_planes_read(s3d_decoder* d, int plane, float* host_out, int64_t n_floats);

/* ---- test / bring-up hooks (not part of the drop-in surface) ----
 * After a forward, intermediate fp32 activations of the last plan can be read back by name
 * ("in_conv", "<block>.h1", "<block>.out", "down.<l>", "upcat.<j>"); plane 0/1/2 = xy/xz/yz, layout
 * [B][rows][cols][C].  Synchronises the device.  tests/ use it to localise a failing kernel. */
int s3d_unet_debug_count(const s3d_unet* u);
int s3d_unet_debug_info(const s3d_unet* u, int index, const char** name, int* channels, int rows[3], int cols[3],
                        int* batch);
int s3d_unet_debug_read(s3d_unet* u, int index, int plane, float* host_out, int64_t n_floats);

#ifdef __cplusplus
}
#endif
#endif /* SIN3DM_B200_H */
