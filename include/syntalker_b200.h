/*
 * syntalker_b200 -- C ABI of the B200-native SynTalker sampling hot path.
 *
 * The reference (RobinWitch/SynTalker) has no FFI: its boundary for this path is four Python call
 * sites.  Each entry point below names the reference interface it replaces; the
 * Python shim in syntalker_b200/ keeps those Python signatures and forwards here through ctypes
 * (INTEGRATION.md shows the binding).  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ST_E* code; st_last_error() gives the text
 *     (thread-local).  Nothing throws across the boundary.
 *   - pointers documented "device" are CUDA device pointers on the device the handle was created on;
 *     the caller owns them; the library never frees caller memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     stream-ordered; no entry point synchronises the host except the *_host ones and *_create.
 *   - handles own their weights and a workspace that grows (cudaMalloc) only when a larger batch
 *     than ever before is seen -- never in steady state.
 *   - all floating point data is IEEE fp32 unless stated; code indices are int64 like torch.argmax.
 */
#ifndef SYNTALKER_B200_H_
#define SYNTALKER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_ABI_VERSION 1

/* error codes */
#define ST_OK 0
#define ST_EINVAL (-1)   /* bad argument (shape, null pointer, unknown tensor name, t out of range) */
#define ST_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define ST_ENOMEM (-3)   /* device allocation failed */
#define ST_ESTATE (-4)   /* call order violated (e.g. sample before cond encode) */
#define ST_EUNSUPPORTED (-5)

/* fixed geometry of the path (models/denoiser.py:19-22,36-37; diffusion_rvqvae_trainer.py:422) */
#define ST_TOKENS 32          /* latent tokens per window = 128 frames / vqvae_squeeze_scale 4 */
#define ST_FRAMES 128
#define ST_LATENT 1536        /* 3 body parts x 512 */
#define ST_DMODEL 512
#define ST_AUDIO_LEN 68224    /* 16000/30*128 samples, 2 channels (amplitude, onset) */
#define ST_SEED_FRAMES 4
#define ST_MAX_T 1000         /* original diffusion steps */
#define ST_MAX_EVALS 4

/* model variants: which MDM the weights came from */
#define ST_VARIANT_BEATX 0             /* models/denoiser.py MDM, use_motionclip=False */
#define ST_VARIANT_BEATX_MOTIONCLIP 1  /* models/denoiser.py MDM, use_motionclip=True (512-d style) */
#define ST_VARIANT_H3D 2               /* models/denoiser_h3d.py MDM (256-d prompt style, null embedding) */

/* GEMM engines (st_set_engine): SIMT = exact-fp32 FMA kernels; TC = tcgen05 split-fp16 (3 MMAs/product) */
#define ST_ENGINE_SIMT 0
#define ST_ENGINE_TC 1

typedef struct st_model st_model;
typedef struct st_vq st_vq;
typedef struct st_schedule st_schedule;

/* A named host tensor (fp32 unless the name is documented otherwise). */
typedef struct {
  const char* name;
  const void* data;   /* host pointer */
  int64_t numel;
} st_tensor;

const char* st_last_error(void);
int st_abi_version(void);
/* Number of kernels this library has launched on the calling thread's device since load (bench.py's
 * gpu_launches claim is read from here). */
int64_t st_launch_count(void);
int st_set_engine(int engine);
int st_get_engine(void);
/* The engine of ONE handle (-1 = follow the process default of st_set_engine again).  Every call runs on the engine of the handle it is
 * made on, so two handles driven from two host threads may use different engines; the st_debug_* switches stay process-wide. */
int st_model_set_engine(st_model* m, int engine);
int st_vq_set_engine(st_vq* v, int engine);
/* The sampling loop replays one captured CUDA graph per diffusion step (default on); 0 = launch every kernel eagerly. */
int st_set_graphs(int on);
/* Programmatic dependent launch between consecutive kernels (default on); 0 = plain stream order. */
int st_set_pdl(int on);
/* Debug: device buffer of >= 64 int64 receiving clock64() stamps of CTA (0,0) of every tcgen05 GEMM launch; NULL = off. */
int st_debug_timeline(long long* dev_buf);
/* Debug: only trunk-kernel launches with this N and K record the timeline (0, 0 = every launch). */
int st_debug_timeline_select(int N, int K);
int st_debug_timeline_select2(int N, int K);   /* a second shape, recorded 1024 slots behind the first (tests/boundary_probe.py) */
/* Debug: device buffer of 120000 uint64; every kernel's first thread appends (%globaltimer ns, kernel id); slot 0 = count. NULL = off. */
int st_debug_trace(unsigned long long* dev_buf);
/* debug / A-B switches.  bits 0-3: timing-only variants of the tcgen05 main loop (results are garbage); 16: trunk kernel off
 * (generic tcgen05 kernel); 32: separate attention kernel; 64: RVQ code ranking on the SIMT engine; 128: WavEncoder with fp32
 * activations; 256: body parts decode one after the other; 512: the sampling loop keeps its state in x space (no z recursion);
 * 1024: the last block's fc2 stays a layer of its own inside the z recursion; 2048: packed-fp32 FMA attention epilogue instead of
 * the tcgen05 Q K^T / P V; 4096: the two CFG evaluations as two chains on two streams (experiment); 8192: conv taps fetched per tap
 * instead of staged once; 16384: the MMA issue loop as first written; 32768: timelines record the LAST CTA of a grid instead of
 * CTA (0,0); 65536: per-CTA boundary records behind the timeline slots (tests/boundary_probe.py); 131072: the trunk layers of an
 * evaluation stack as one cluster launch (layer chains; slower, off by default); 262144: trunk layers wait for the whole predecessor
 * grid (griddepcontrol.wait) instead of the row-tile signals. */
int st_debug_probe(int flags);
/* Parity-test taps of the last st_cond_encode: atcat_out [B,128,512] = [WavEncoder output | word features]
 * per frame (models/denoiser.py:151-155, B <= 32), cst_out [B*32,512] = the hoisted conditioning constant.  Either may be NULL. */
int st_debug_cond_taps(st_model* m, float* atcat_out, float* cst_out, int B, void* stream);

/* ---- weights -------------------------------------------------------------------------------------
 * Replaces: MDM(args) construction + load_checkpoints (train.py:85-94, utils/other_tools.py:771-790).
 * `packed` are the tensors syntalker_b200/packer.py derives from the reference state dict
 * (BatchNorm folded into the WavEncoder convs, input_process/2/3 folded into one 1536->512 matrix,
 * the timestep MLP tabulated for t = 0..999).  Names are checked; a missing or
 * mis-sized tensor is ST_EINVAL.  Optional: "w_xo" [512,512] = W_x W_out and "c_xo" [512] = W_x b_out; with them the
 * deterministic DDIM loop of the tcgen05 engine carries W_x x_k between steps instead of x_k (st_sample). */
int st_model_create(const st_tensor* packed, int n, int variant, st_model** out);
void st_model_destroy(st_model* m);

/* Replaces: RVQVAE(args, D, 512, 512, 512, 2, 2, 512, 3, 3, 'relu', None) + load 'net'
 * (diffusion_rvqvae_trainer.py:106-155).  Decoder side only: 6 codebooks + conv decoder. */
int st_vq_create(const st_tensor* packed, int n, int out_dim, st_vq** out);
void st_vq_destroy(st_vq* v);
int st_vq_out_dim(const st_vq* v);

/* ---- schedule ------------------------------------------------------------------------------------
 * Replaces: create_gaussian_diffusion()/SpacedDiffusion.__init__ tables (diffusion/model_util.py:8-50,
 * respace.py:73-87, gaussian_diffusion.py:160-197) and the per-step _extract_into_tensor gathers
 * (gaussian_diffusion.py:1606-1619).  The host computes the fp64 tables and rounds the per-step
 * coefficients to fp32 in the reference's operation order (syntalker_b200/schedule.py); here they
 * are only uploaded.
 *   mode ST_MODE_DDPM: coef[k] = {coef1, coef2, sigma, 0, 0}:  x <- coef1*x0 + coef2*x + sigma*eps_k
 *   mode ST_MODE_DDIM: coef[k] = {a, b, c1, c2, sigma}:        e = (a*x - x0)/b; x <- x0*c1 + c2*e + sigma*eps_k
 * k runs S-1 .. 0;  t_model[k] is the ORIGINAL timestep fed to the denoiser (respace.py:124-129). */
#define ST_MODE_DDPM 0
#define ST_MODE_DDIM 1
#define ST_COEF_STRIDE 5
int st_schedule_create(int S, int mode, const int32_t* t_model, const float* coef, st_schedule** out);
void st_schedule_destroy(st_schedule* s);

/* ---- conditioning --------------------------------------------------------------------------------
 * The y dict of the reference (diffusion_rvqvae_trainer.py:433-443, h3d_diffusion_new_trainer.py:548-556). */
typedef struct {
  const float* audio;     /* device [B, ST_AUDIO_LEN, 2] */
  const int32_t* word;    /* device [B, 128] */
  const float* seed;      /* device [B, 4, 1536] */
  const float* style[3];  /* device [B, style_dim] or NULL.  BEATX_MOTIONCLIP: style[0] (512-d).
                             H3D: {upper, hands, lower} prompt vectors (256-d); NULL = no prompt */
} st_cond;

/* Guidance = which reference wrapper is around the model (diffusion/cfg_sampler.py). */
#define ST_CFG_NONE 0       /* bare MDM; `flags` may force uncond / uncond_audio (y['uncond'], y['uncond_audio']) */
#define ST_CFG_TEXT 1       /* ClassifierFreeSampleModel            :17-28   scale[B] */
#define ST_CFG_TWO 2        /* TwoClassifierFreeSampleModel          :38-54   scale_audio[B], scale_prompt[B]; style[0] is the prompt */
#define ST_CFG_BODYPART 3   /* TwoClassifierFreeSampleModel_Bodypart :67-117  audio_scale 1, prompt_scale 4 unless overridden */
#define ST_CFG_BODYPART1 4  /* ClassifierFreeSampleModel_Bodypart :125-167: per prompted part u + scale (ua_part - u), u = (null prompt, audio) */
#define ST_FLAG_UNCOND 1
#define ST_FLAG_UNCOND_AUDIO 2
typedef struct {
  int mode;
  int flags;                 /* ST_CFG_NONE only */
  const float* scale;        /* host [B]: TEXT scale / TWO scale_audio */
  const float* scale2;       /* host [B]: TWO scale_prompt */
  float audio_scale;         /* BODYPART (cfg_sampler.py:64)  default 1 */
  float prompt_scale;        /* BODYPART (cfg_sampler.py:65)  default 4 */
} st_guidance;

/* Encode the step-invariant conditioning once (D2, D3, D3b, D3c of SURVEY.md §8a) into the model's
 * cache.  The reference recomputes this inside every MDM.forward (denoiser.py:147-157). */
int st_cond_encode(st_model* m, const st_cond* cond, int B, void* stream);

/* ---- denoiser ------------------------------------------------------------------------------------
 * Replaces: model(x, timesteps, y=dict) -- MDM.forward (models/denoiser.py:132, denoiser_h3d.py:148)
 * optionally inside a CFG wrapper.  x, out: device [B,1536,1,32];  t: device int64 [B] original
 * timesteps (0..999).  Uses the cache of the last st_cond_encode with the same B. */
int st_denoise(st_model* m, const float* x, const int64_t* t, const st_guidance* g, float* out, int B, void* stream);

/* ---- sampler -------------------------------------------------------------------------------------
 * Replaces: diffusion.p_sample_loop / ddim_sample_loop (gaussian_diffusion.py:607,888) with
 * clip_denoised=False, no cond_fn -- the only way the trainers call it.  x_init: device [B,1536,1,32]
 * (the `noise` argument or th.randn(shape)).  noise_tape: device [S,B,1536,1,32] eps_k for k = S-1..0
 * stored in draw order (tape[0] is used at k=S-1), required when any sigma != 0, else may be NULL.
 * x_out may alias x_init.  Uses the cache of the last st_cond_encode.
 * On the tcgen05 engine with guidance NONE / TEXT / TWO the update runs in token space ("z recursion"):
 * x_{k-1} = alpha_k x0_hat + beta_k x_k + sigma_k eps_k is linear and the next step only needs W_x x_{k-1}, so between steps one
 * 512x512 GEMM replaces output GEMM + state update + input GEMM (the noise enters as W_x eps_k, formed for a chunk of steps at
 * a time); the state is formed by the last step (DDIM: x <- x0_hat, gaussian_diffusion.py:774-790 with alpha_bar_prev = 1;
 * DDPM: coef1 = 1, coef2 = 0, sigma = 0 at t = 0, :383,:556).  Same result to ~1e-5; st_debug_probe(512) keeps the x-space loop. */
int st_sample(st_model* m, const st_schedule* s, const st_guidance* g, const float* x_init, const float* noise_tape,
              int B, float* x_out, void* stream);

/* The same loop in pieces, for callers that hand the per-step noise over in chunks instead of one [S,B,1536,1,32] tape (the
 * reference draws one th.randn_like(x) per step, gaussian_diffusion.py:541,781; a 1000-step tape at B = 32 is 6.3 GB):
 *   st_sample_begin(...);  repeat { st_sample_run(m, n, noise_chunk ( [n,B,1536,1,32] in draw order, or NULL ), stream); }
 *   until S steps have run;  st_sample_end(m, x_out, stream).
 * st_sample_chunk(s) is the number of steps one captured graph holds; pieces that are multiples of it (and aligned to it) replay
 * graphs, other sizes run step by step, and a loop that needs noise on the z recursion only accepts multiples.  The schedule and
 * the guidance scales must stay alive until st_sample_end.  One loop per model handle at a time. */
int st_sample_chunk(const st_schedule* s);
int st_sample_begin(st_model* m, const st_schedule* s, const st_guidance* g, const float* x_init, int B, void* stream);
int st_sample_run(st_model* m, int n_steps, const float* noise_chunk, void* stream);
int st_sample_end(st_model* m, float* x_out, void* stream);

/* ---- RVQ-VAE decode ------------------------------------------------------------------------------
 * Replaces: vq.latent2origin(x)[0] (models/vq/model.py:102-109).  lat: device [B,T4,512] already
 * multiplied by vqvae_latent_scale; rec: device [B,4*T4,D]; idx_out: device int64 [B,T4,6] or NULL.
 * lat_stride = floats between consecutive tokens (512 when contiguous; 1536 to read one body part
 * straight out of a [B,T4,1536] sample).  If residual_out != NULL the final residual is written there
 * [B,T4,512] (the reference leaves it in its input tensor, residual_vq.py:146). */
int st_rvq_decode(st_vq* v, const float* lat, int64_t lat_stride, float lat_scale, int B, int T4, float* rec,
                  int64_t* idx_out, float* residual_out, void* stream);

/* ---- RVQ-VAE encode ---------------------------------------------------------
 * Replaces: vq.map2latent(x) (models/vq/model.py:95-100, encdec.py:5-34; callers diffusion_rvqvae_trainer.py:290-294).
 * pose: device [B,T,D] normalised pose features of one body part, T a multiple of 4; lat: device [B,T/4,512], the
 * encoder output BEFORE quantisation (the trainer divides it by vqvae_latent_scale to form latent_in / the seed).
 * Needs a handle created from a checkpoint that holds the encoder weights. */
int st_rvq_encode(st_vq* v, const float* pose, int B, int T, float* lat, void* stream);

/* ---- 330-d assembly ------------------------------------------------------------------------------
 * Replaces: diffusion_rvqvae_trainer.py:484-531.  rec_*: device decoder outputs [B,n,78|180|57];
 * mean/std: device [330]; trans_mean/std: device [3]; jaw_aa: device [B,n,3] or NULL (zeros).
 * rec_pose: device [B,n,330]; rec_trans: device [B,n,3]. */
int st_pose_assemble_330(const float* rec_upper, const float* rec_hands, const float* rec_lower, const float* mean,
                         const float* std, const float* trans_mean, const float* trans_std, const float* jaw_aa, int B,
                         int n, float* rec_pose, float* rec_trans, void* stream);
/* h3d: scatter the three decoder outputs [B,n,156|360|107] into [B,n,623]
 * (h3d_diffusion_new_trainer.py:194-221,604-607). */
int st_pose_assemble_623(const float* rec_upper, const float* rec_hands, const float* rec_lower, int B, int n,
                         float* rec_pose, void* stream);

/* ---- layout helpers ------------------------------------------------------------------------------ */
/* sample [B,1536,1,T] -> token-major [B,T,1536] * scale (trainer:457 `squeeze().permute(1,0)` batched). */
int st_sample_to_tokens(const float* sample, int B, int T, float scale, float* tokens, void* stream);

/* ---- evaluation tail (the part that needs neither SMPL-X nor the VAESKConv evaluator weights) ----------
 * st_pose_330_to_aa165: rec_pose [frames,330] (55 joints x 6d) -> poses [frames,165] axis-angle, the `poses` array of the result
 *   files (diffusion_rvqvae_trainer.py:621-622: rotation_6d_to_matrix -> matrix_to_axis_angle; np.savez format :702-710).
 * st_moments_accumulate: acc (device float64 [1 + D + D*D] = n, sum x, sum x x^T) += the moments of x [N,D]; the FID of the
 *   reference (dataloaders/data_tools.py:1615-1625: np.mean, np.cov of all latents) follows from the summed statistics, so ranks
 *   all-reduce `acc` (NCCL sum) instead of gathering latents.
 * st_l1div_accumulate: acc (device float64 [2] = sum, counter) += L1div.run(x [n,J]) (utils/metric.py:16-22), one sequence per call. */
int st_pose_330_to_aa165(const float* rec_pose, int64_t frames, float* poses_aa, void* stream);
int st_moments_accumulate(const float* x, int64_t N, int D, double* acc, void* stream);
int st_l1div_accumulate(const float* x, int n, int J, double* acc, void* stream);

/* ---- whole window, device buffers ----------------------------------------------------------------
 * cond encode -> sample -> x latent_scale -> latent2origin x3 -> 330-d, everything resident in HBM.
 * ms: device [666] = mean[330] | std[330] | trans_mean[3] | trans_std[3].  sample_out (nullable) receives
 * the final latent [B,1536,1,32]; x_init is not modified.  No host synchronisation. */
int st_generate_330(st_model* m, const st_schedule* s, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                    st_vq* vq_lower, const st_cond* cond, const float* x_init, const float* noise_tape, const float* jaw_aa,
                    const float* ms, int B, float latent_scale, float* rec_pose, float* rec_trans /* nullable */,
                    float* sample_out /* nullable */, void* stream);

/* ---- whole window, host buffers ------------------------------------------------------------------
 * One call a trainer's _g_test makes per window batch: H2D(cond, noise) -> cond encode -> sample ->
 * x latent_scale -> latent2origin x3 -> 330-d -> D2H.  All pointers are HOST (ideally pinned).
 * Synchronises `stream` before returning.  noise_tape_host may be NULL when all sigma == 0. */
typedef struct {
  const float* audio; const int32_t* word; const float* seed; const float* style[3];
  const float* x_init; const float* noise_tape; const float* jaw_aa;
  const float* mean; const float* std; const float* trans_mean; const float* trans_std;
} st_host_inputs;
int st_generate_330_host(st_model* m, const st_schedule* s, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                         st_vq* vq_lower, const st_host_inputs* in, int B, float latent_scale, float* rec_pose_host,
                         float* rec_trans_host, float* sample_host /* nullable [B,1536,1,32] */, void* stream);
/* The same call split in two, so that TWO window batches can be in flight: `slot` (0 or 1) names a staging set; _begin queues
 * the H2D copies on an internal copy stream, the computation on `stream` and the D2H copies on a second copy stream, and
 * returns; _wait blocks until the slot's results are in the host buffers given to _begin.  Inputs of batch i + 1 then cross
 * PCIe while batch i computes.  The host buffers must stay valid (and pinned, for the copies to be asynchronous) until
 * _wait returns; a slot must be waited for before it is reused (ST_ESTATE otherwise). */
int st_generate_330_host_begin(st_model* m, const st_schedule* s, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                               st_vq* vq_lower, const st_host_inputs* in, int B, float latent_scale, float* rec_pose_host,
                               float* rec_trans_host, float* sample_host /* nullable */, int slot, void* stream);
int st_generate_330_host_wait(st_model* m, int slot);

/* ---- long clip: the window loop ----------------------------------------------
 * Replaces: the `for i in range(0, roundt)` loop of _g_test with its seed hand-off and latent stitching, plus the single
 * latent2origin x3 and 330-d assembly over the whole clip (diffusion_rvqvae_trainer.py:413-531).
 * audio_long: device [B, audio_len, 2]; word_long: device int32 [B, n_words]; R windows need n_words >= 112 R + 16 and
 * audio_len >= 533 (112 R + 16).  seed0: device [B,4,1536] (window 0; later windows take the previous sample's last 4
 * tokens).  style: NULL or 3 device pointers like st_cond.  x_init: device [R,B,1536,1,32] start noise per window.
 * noise_tape: NULL or device [R,S,B,1536,1,32].  jaw_aa: NULL or device [B,n,3].  ms: device [666] = mean 330, std 330,
 * trans_mean 3, trans_std 3.  Output n = 4 (32 + 28 (R-1)) frames: rec_pose device [B,n,330], rec_trans device [B,n,3]
 * or NULL, latents_out device [B, n/4, 1536] or NULL (the stitched sample, before x latent_scale). */
int st_generate_long_330(st_model* m, const st_schedule* s, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands, st_vq* vq_lower,
                         const float* audio_long, int64_t audio_len, const int32_t* word_long, int64_t n_words, const float* seed0,
                         const float* const* style, const float* x_init, const float* noise_tape, const float* jaw_aa,
                         const float* ms, int B, int R, float latent_scale, float* rec_pose, float* rec_trans, float* latents_out,
                         void* stream);

/* ---- GEMM-engine profiling (bench.py roofline leg) -------------------------------------------------
 * Between begin and end every GEMM-engine launch is recorded.  end() replays exactly that launch sequence as
 * one captured CUDA graph bracketed by two CUDA events on its stream and returns the device time of the GEMM
 * kernels alone (ms), their summed algorithmic FLOPs (2*M*N*K per launch, K counting every tap of a conv) and
 * the launch count.  The sampler runs without its step graph while recording; never on in timed runs. */
int st_profile_begin(void);
int st_profile_end(double* ms_total, double* flops_total, int64_t* launches);

/* Steady-state device time (ms) of one GEMM launch: `reps` launches in one CUDA graph between two events. */
int st_bench_gemm(int M, int N, int K, int engine, int reps, const float* A, const float* W, const float* bias, float* out,
                  double* ms_per_launch);

/* ---- self test (no oracle involved): split-fp16 tcgen05 GEMM vs the SIMT fp32 GEMM on device ----- */
int st_selftest_gemm(int M, int N, int K, int engine, const float* A, const float* W, const float* bias, float* out,
                     void* stream);
/* The same for a stride-1 "same" Conv1d over channels-last activations (implicit GEMM): A [B,L,C], W [N, taps*C] tap-major as the
 * packer lays conv weights out (row stride rounded up to 4 floats), out [B*L, N]. */
int st_selftest_conv(int B, int L, int C, int N, int taps, int dil, int engine, const float* A, const float* W, const float* bias,
                     float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNTALKER_B200_H_ */
