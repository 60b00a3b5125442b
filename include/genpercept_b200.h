/* genpercept_b200 — C-ABI of the B200-native one-step perception engine.
 *
 * This is the drop-in boundary for the hot path of aim-uofa/GenPercept:
 *   GenPerceptPipeline.single_infer        /root/reference/genpercept/genpercept_pipeline.py:375-486
 *     encode_rgb  (vae.encoder, quant_conv, mean * 0.18215)                     :488-505
 *     unet(pred_latent, t=1, empty-text embed) + DDIM(beta=1) step == -v       :443-465
 *     decode_pred (/0.18215, post_quant_conv, vae.decoder, channel mean)       :507-526
 *     clip(-1,1), (x+1)/2                                                      :470-472
 *     or the DPT readout (customized_head on multi_level_feats, min-max)       :475-482
 * Everything below the boundary is hand-written sm_100a CUDA; there is no CPU fallback: every
 * entry point returns GP_ERR_CUDA when no CUDA device / kernel image is available.
 *
 * Plain C types only.  Device pointers are raw CUdeviceptr-compatible addresses (e.g.
 * torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as void*.
 */
#ifndef GENPERCEPT_B200_H
#define GENPERCEPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gp_engine gp_engine;

typedef enum {
  GP_OK = 0,
  GP_ERR_INVALID = 1,   /* bad argument / shape (reference: Python assert / ValueError) */
  GP_ERR_MISSING = 2,   /* a checkpoint tensor required by the topology was never loaded  */
  GP_ERR_NO_PLAN = 3,   /* gp_infer before gp_plan for this (B,H,W)                        */
  GP_ERR_CUDA = 4,      /* CUDA error (sticky: the engine is poisoned)                      */
  GP_ERR_STATE = 5      /* call order (e.g. gp_plan before gp_finalize)                     */
} gp_status;

typedef enum { GP_F32 = 0, GP_F16 = 1, GP_BF16 = 2, GP_U8 = 3 } gp_dtype;
typedef enum { GP_READOUT_VAE = 0, GP_READOUT_DPT = 1 } gp_readout;

typedef struct {
  int device;            /* CUDA device ordinal                                             */
  int dtype;             /* GP_F16 or GP_BF16: storage / tensor-core operand type            */
  int readout;           /* gp_readout: VAE decoder (run.py default) or DPT head (:296-301)  */
  int timestep;          /* UNet timestep; 1 for GenPercept (ddim.py + scheduler beta=1), or
                            the reference's --fix_timesteps value                           */
  int use_cuda_graph;    /* 0 eager, 1 replay one captured CUDA graph per plan, 2 auto (small plans) */
  int precision;         /* 0: 16-bit storage, fp32 accumulate (the reference's --half_precision class);
                            1: high — every activation and weight is carried as an fp16 (hi, lo) pair and every
                            contraction runs hi*hi + lo*hi + hi*lo on the tensor cores (fp32-class products,
                            fp32 accumulate): the reference's default fp32 run (run.py:273-281)              */
  int arch;              /* 0: one-step GenPercept (the scheduler's beta = 1 step folded into the UNet tail);
                            1: multi-step (run.py --archs marigold / rgb_blending): the UNet returns model_output, real
                            DDIM steps run around it (gp_infer_steps); conv_in may take 8 channels (run.py:59-78)   */
} gp_config;

/* replaces: GenPerceptPipeline.__init__/from_pretrained model assembly (run.py:314-376) */
gp_status gp_create(const gp_config* cfg, gp_engine** out);
void gp_destroy(gp_engine* e);
const char* gp_last_error(gp_engine* e);

/* replaces: load_state_dict of the diffusers-format checkpoints (run.py:336-357, :296-312).
 * `key` = "<component>.<diffusers key>", component in {unet, vae, dpt}.  Host pointer, copied. */
gp_status gp_load_tensor(gp_engine* e, const char* key, const void* host_ptr, int dtype,
                         const int64_t* shape, int ndim);
/* replaces: encode_text()'s cached self.text_embed (genpercept_pipeline.py:360-372, :425-429).
 * fp32 [n_tokens, 1024].  n_tokens == 2 (the empty prompt) takes the closed-form cross-attention; any other
 * length the general one (context projections folded into two 1x1 GEMMs around a per-head softmax).  The
 * context is a constant of the engine from gp_finalize on, like the reference's cached self.text_embed. */
gp_status gp_set_text_embed(gp_engine* e, const float* host_ptr, int n_tokens, int dim);
/* folds constants, re-packs weights K-major 16-bit, uploads. */
gp_status gp_finalize(gp_engine* e);

/* builds the static op list, activation arena and (optionally) CUDA graph for one input shape.  Any H, W >= 32: the
 * result extent (gp_tensor_shape "out") is 8*floor(H/8) x 8*floor(W/8) for the VAE readout (AutoencoderKL's
 * stride-2 stages floor) and the DPT head's pyramid extent for the DPT readout; equal to H x W for multiples of 8 / 64. */
gp_status gp_plan(gp_engine* e, int batch, int height, int width);

/* number of cached plans (the cache is bounded: least-recently-used plans are destroyed; GP_MAX_PLANS, default 4) */
int gp_plan_count(gp_engine* e);
/* Host-only introspection (no device needed): the N tile (BN) and the number of 128-pixel M tiles per CTA (MT) the planner
 * gives a stride-1 ks x ks convolution / linear layer cin -> cout over `images` maps of h x w output pixels on a GPU with
 * num_sms SMs (tokens_mode != 0: one row of images * h * w tokens).  Nothing in the reference corresponds to it (PyTorch /
 * cuDNN pick their own tiles); it pins the tile policy of DESIGN.md section 4 in the CPU tests. */
gp_status gp_tile_shape(int cout, int cin, int ks, int images, int h, int w, int tokens_mode, int num_sms, int* bn, int* mt);
/* replaces: the per-call `fix_timesteps` of single_infer (/root/reference/genpercept/genpercept_pipeline.py:405-408).
 * The timestep only enters through conv1.bias + time_emb_proj(silu(emb(t))) of the 22 UNet ResNets; those biases are
 * re-folded on the host (cached per timestep) and rewritten in place after a device synchronisation. */
gp_status gp_set_timestep(gp_engine* e, int timestep);

/* replaces: single_infer.  rgb: [B,3,H,W] NCHW, device (or pinned/pageable host if
 * rgb_on_host != 0; copied on `stream`), dtype GP_U8 (0..255, mapped x/255*2-1 as
 * genpercept_pipeline.py:245) or GP_F16/GP_F32 already in [-1,1].
 * out: fp32 [B,C,outH,outW] in [0,1] (a device `out` is written by the last kernel itself), C = out_channels (1: channel-mean modes depth/matting/dis/
 * disparity :523-525; 3: normal/seg); device, or host if out_on_host != 0.  Asynchronous on
 * `stream` unless a host buffer is involved, in which case it returns after the copy. */
gp_status gp_infer(gp_engine* e, const void* rgb, int rgb_dtype, int rgb_on_host, float* out,
                   int out_on_host, int out_channels, void* stream);

/* replaces: encode_rgb (/root/reference/genpercept/genpercept_pipeline.py:488-505).  rgb as for gp_infer; latent_dev:
 * fp32 [B,4,H/8,W/8] on the device = mean(quant_conv(encoder(rgb))) * 0.18215. */
gp_status gp_encode(gp_engine* e, const void* rgb, int rgb_dtype, int rgb_on_host, float* latent_dev, void* stream);
/* replaces: decode_pred + clip + shift (:507-526, :470-472).  latent_dev: fp32 [B,4,h,w] on the device (in the scaled
 * latent space, like the reference's pred_latent); apply_post_quant != 0 applies vae.post_quant_conv to latent / 0.18215 as
 * the reference always does (0: the latent already went through it, e.g. the engine's own "z").  out_dev: fp32
 * [B,C,8h,8w] in [0,1] (the reference clips to [-1,1] right after decode_pred, :470; map = out * 2 - 1). */
gp_status gp_decode(gp_engine* e, const float* latent_dev, int apply_post_quant, float* out_dev, int out_channels, void* stream);

/* replaces: single_infer for the multi-step archs (/root/reference/genpercept/genpercept_pipeline.py:399-472; gp_config.arch = 1):
 *   rgb_latent = encode_rgb(rgb); pred_latent = noise (marigold; fp32 [B,4,h,w], host or device) or rgb_latent (noise == NULL:
 *   rgb_blending); per step i: unet(cat([rgb_latent, pred_latent]) or pred_latent, timesteps[i]) -> DDIM step with
 *   coeffs[4 i .. 4 i + 3] = (x0 <- sample, x0 <- model_output, prev <- sample, prev <- model_output) (eta = 0; host
 *   arrays, genpercept_b200/scheduler.py); then decode_pred(pred_original_sample), clip, shift.  out as for gp_infer. */
gp_status gp_infer_steps(gp_engine* e, const void* rgb, int rgb_dtype, int rgb_on_host, const float* noise, int noise_on_host,
                         const int* timesteps, const float* coeffs, int n_steps, float* out, int out_on_host, int out_channels,
                         void* stream);

/* Stage entry points for parity tests (each runs a contiguous slice of the planned op list). */
typedef enum { GP_STAGE_PRE = 0, GP_STAGE_VAE_ENCODE = 1, GP_STAGE_UNET = 2, GP_STAGE_READOUT = 3 } gp_stage;
gp_status gp_run_stage(gp_engine* e, int stage, int out_channels, void* stream);
/* Named internal tensors kept alive by the plan: "rgb", "rgb_latent", "z" (decoder input =
 * post_quant_conv(-unet_out/0.18215)), "feat0".."feat3" (DPT taps), "out".  fp32 NCHW on host. */
gp_status gp_tensor_shape(gp_engine* e, const char* name, int64_t shape[4]);
gp_status gp_read_tensor(gp_engine* e, const char* name, float* host_out, size_t capacity_elems);
gp_status gp_write_tensor(gp_engine* e, const char* name, const float* host_in, size_t elems);

/* Introspection for bench.py */
gp_status gp_plan_info(gp_engine* e, int64_t* n_ops, int64_t* n_kernel_launches, int64_t* arena_bytes,
                       int64_t* weight_bytes, double* igemm_flops);
/* name/us of the i-th op after gp_profile_ops ran the plan once with CUDA events per op. */
gp_status gp_profile_ops(gp_engine* e, int out_channels, void* stream);
/* kind: 1 = tcgen05 implicit-GEMM launch, 2 = fused attention, 0 = other kernels.  flops = algorithmic work of the op
 * (SURVEY.md 8d); flops_exec = MMA work actually issued (differs for the upsample-fused convolutions: 4 of 9 taps). */
gp_status gp_op_info(gp_engine* e, int64_t i, char* name_buf, size_t name_cap, double* usec, double* flops,
                     double* bytes, int* kind, double* flops_exec);

/* ---- per-kernel entry points (parity tests, micro-benchmarks); all pointers are device ---- */
/* 3x3 / 1x1 convolution through the tcgen05 implicit-GEMM kernel.  x: 16-bit NHWC [N,H,W,Cin];
 * w: fp32 [Cout,Cin,ks,ks] (host); mode: 0 stride-1 pad ks/2, 1 stride-2 pad (1,1,1,1),
 * 2 stride-2 pad (0,1,0,1) (VAE encoder), 3 nearest-2x upsample then stride-1.  y: 16-bit NHWC. */
gp_status gp_conv2d(int dtype, const void* x, int N, int H, int W, int Cin, const float* w_host,
                    const float* bias_host, int Cout, int ks, int mode, const void* residual, int relu,
                    void* y, int use_direct_kernel, void* stream);
gp_status gp_groupnorm(int dtype, const void* x, int N, int H, int W, int C, int groups, const float* gamma_host,
                       const float* beta_host, float eps, int silu, void* y, void* stream);
/* GroupNorm(groups, eps)(+SiLU) -> 3x3 stride-1 convolution (+ optional 1x1 shortcut over a RAW second tensor sc_x
 * [N,H,W,Csc], + optional residual): the norm1/conv1, norm2/conv2 (+conv_shortcut) and conv_norm_out/conv_out pairs of the
 * diffusers ResnetBlock2D / VAE heads.  Where W % 128 == 0 the normalisation runs inside the convolution's operand
 * path (no normalised tensor in HBM), elsewhere as GroupNorm pass + convolution.  y: 16-bit NHWC, or (out_f32 != 0)
 * fp32 NCHW [N,Cout,H,W]. */
gp_status gp_gn_conv3x3(int dtype, const void* x, int N, int H, int W, int Cin, int groups, const float* gamma_host,
                        const float* beta_host, float eps, int silu, const float* w_host, const float* bias_host, int Cout,
                        const void* sc_x, int Csc, const float* sc_w_host, const float* sc_b_host, const void* residual,
                        void* y, int out_f32, void* stream);
gp_status gp_layernorm(int dtype, const void* x, int64_t tokens, int C, const float* gamma_host,
                       const float* beta_host, float eps, void* y, void* stream);
/* softmax(q k^T * scale) v per (batch, head); q,k,v,o: 16-bit [B,T,heads*d] */
gp_status gp_attention(int dtype, const void* q, const void* k, const void* v, int B, int T, int heads, int d,
                       float scale, void* o, void* stream);
/* replaces: the align / reduce / normalise tail of ensemble_depth (/root/reference/genpercept/util/ensemble.py:101-156,
 * 186-203): out[H,W] = median (torch.median: lower middle) or mean over the B <= 32 members of pred[b] * scale[b] + shift[b],
 * then (normalise 1) (x - min) / max(max - min, 1e-6) or (normalise 2) x / max(max, 1e-6).  pred / out: fp32 on the device. */
gp_status gp_ensemble_reduce(const float* pred_dev, int B, int H, int W, const float* scale_host, const float* shift_host,
                             int median, int normalise, float* out_dev, void* stream);
gp_status gp_bilinear_up2x(int dtype, const void* x, int N, int H, int W, int C, void* y, void* stream);
/* ---- pre/post-processing around the hot path; buffers may be host or device ------------
 * gp_resize_aa replaces torchvision.transforms.functional.resize(tensor, size, interpolation, antialias=True)
 * as called by resize_max_res (/root/reference/genpercept/util/image_util.py:75-105) and by the resize back
 * to the input resolution (/root/reference/genpercept/genpercept_pipeline.py:301-307): N planes [N,H,W] ->
 * [N,OH,OW]; src GP_U8 or GP_F32; dst GP_F32, or GP_U8 = round-half-even (+ clamp for bicubic) as torchvision
 * does for integer tensors; mode 0 bilinear, 1 bicubic.
 * The three pre/post entry points share one growing scratch buffer and the cached filter tables per device:
 * like gp_infer on one engine they are meant for one caller thread and stream-ordered use (calls on the same
 * stream, or separated by a synchronisation); they return after the copy whenever a host buffer is involved. */
gp_status gp_resize_aa(const void* src, int src_dtype, int src_on_host, int N, int H, int W, void* dst, int dst_dtype,
                       int dst_on_host, int OH, int OW, int mode, void* stream);
/* colorize_depth_maps + the uint8 cast at its call site (image_util.py:25-63, genpercept_pipeline.py:318-321):
 * pred f32 [B,H,W] -> u8 [B,H,W,3] = lut[clip(floor((x-vmin)/(vmax-vmin)*256), 0, 255)], lut = 256x3 u8 (host). */
gp_status gp_colorize(const float* pred, int pred_on_host, int B, int H, int W, float vmin, float vmax,
                      const uint8_t* lut768_host, uint8_t* out_hwc, int out_on_host, void* stream);
/* (pred*255).astype(uint8) / (pred*65535).astype(uint16) of /root/reference/run.py:449-455; bits = 8 or 16. */
gp_status gp_quantize(const float* pred, int pred_on_host, size_t n, int bits, void* out, int out_on_host, void* stream);

/* time one igemm configuration: returns average microseconds over `iters` launches */
gp_status gp_bench_conv(int dtype, int N, int H, int W, int Cin, int Cout, int ks, int mode, int iters,
                        double* usec, double* flops);
/* debug (scripts/fattn_trace.py): a device buffer of >= 1024 int64 that CTA 0 of the fused-attention launches planned
 * afterwards fills with clock64() stamps at its phase boundaries; NULL switches the stamps off again. */
void gp_debug_fattn_trace(void* dev_buf);
/* debug (scripts/patch_trace.py): a device buffer of >= 512 int64 that CTA 0 of the patch-resident kernel launches made
 * afterwards fills with clock64() stamps per K chunk (transform: wait / first row landed / done; MMA issuer: wait / ready /
 * issued); NULL switches the stamps off. */
void gp_debug_patch_trace(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif
