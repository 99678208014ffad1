// Internal declarations shared by the translation units of libsyntalker_b200.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <map>
#include <vector>
#include <utility>

#include "../../include/syntalker_b200.h"

namespace st {

// ---- error plumbing -----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int64_t g_launches;

#define ST_CHECK_CUDA(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      (void)cudaGetLastError();                                                                  \
      st::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return ST_ECUDA;                                                                           \
    }                                                                                            \
  } while (0)

#define ST_CHECK_LAUNCH()                                                                        \
  do {                                                                                           \
    st::g_launches++;                                                                            \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      st::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));   \
      return ST_ECUDA;                                                                           \
    }                                                                                            \
  } while (0)

#define ST_TRY(expr)              \
  do {                            \
    int _r = (expr);              \
    if (_r != ST_OK) return _r;   \
  } while (0)

#define ST_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      st::set_error(__VA_ARGS__);  \
      return ST_EINVAL;            \
    }                              \
  } while (0)

// ---- launches: every kernel of the path goes through launch_k.  With PDL on (default) the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs may be scheduled while its predecessor
// drains; every kernel therefore executes pdl_wait() before it touches global memory (reads AND writes) and
// pdl_launch() right after, which keeps the chain transitively ordered (N+2 cannot pass N+1's wait).
extern bool g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// same, for kernels that run as thread-block clusters of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// "once per device" latch for per-device function attributes (cudaFuncSetAttribute): first() is true the first time it is
// called with a given current device.  A racing second thread may set the attribute twice, which is harmless.
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
#ifdef __CUDACC__
// Debug trace (st_debug_trace): CTA (0,0) thread 0 of every kernel stamps %globaltimer and a kernel id at entry.
// In constant memory: every kernel reads this pointer at entry, and as a plain __device__ variable that read was an L2
// round trip on the critical path of every launch (23 % of the trunk kernel's stall samples under ncu).
static __constant__ unsigned long long* g_trace_buf = nullptr;   // per translation unit (no -rdc); [cap][2] = (ns, id); slot 0 = counter
__device__ __forceinline__ void trace_stamp(int id, bool any_thread = false) {
  unsigned long long* t = g_trace_buf;
  if (t && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (any_thread || threadIdx.x == 0)) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    const unsigned long long slot = atomicAdd(t, 1ull) + 1;
    if (slot < 60000) { t[slot * 2] = now; t[slot * 2 + 1] = (unsigned long long)id; }
  }
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- generic implicit-GEMM descriptor ------------------------------------------------------------------
// out[m, n] = act( sum_k A(m,k) * W[n,k] + bias[n] + res_pre[m/res_div, n] ) + res_post[m/res_div, n]
// A(m,k): m -> (b = m / Lout, t = m % Lout);  k -> (j = k / C, c = k % C);  l = t*stride - pad + j*dil;
//         ups: valid iff 0 <= l < 2*Lin, then l >>= 1;  else valid iff 0 <= l < Lin;
//         value = A[b*a_batch + l*lda + c] (0 when invalid), optionally max(.,0) (a_relu).
// A plain Linear is Lout = Lin = M, C = K, stride 1, pad 0.  Conv1d over channels-last activations is the rest.
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_LRELU = 3 };
enum { RES_NONE = 0, RES_PRE = 1, RES_POST = 2 };

struct GemmP {
  const float* A = nullptr;
  const float* W = nullptr;      // [N, ldw] row-major, zero padded to ldw (multiple of 4)
  const float* bias = nullptr;   // [N] or null
  const float* res = nullptr;    // [M/res_div, ldr] or null
  float* out = nullptr;          // [M, ldo]
  int M = 0, N = 0, K = 0, ldw = 0;
  int Lout = 1, Lin = 1, C = 1, stride = 1, pad = 0, dil = 1, ups = 0;
  long long a_batch = 0;
  int lda = 0;
  int a_relu = 0, act = ACT_NONE, res_mode = RES_NONE, ldr = 0, res_div = 1, ldo = 0;
  float out_scale = 1.0f;        // applied to the accumulator before bias (used by nothing exact-critical)
  // tcgen05 engine only: operand already split into fp16 hi/lo planes [2][rows][K] (value * kActScale), and / or
  // the result wanted as such planes for the next GEMM.  `out` may then be null.
  const __half* a_planes = nullptr;
  long long a_plane_stride = 0;   // elements between the hi and the lo plane
  // trunk kernel only: the operand is the column-wise concatenation [a_planes (K - a2_K columns) | a2_planes (a2_K columns)]
  const __half* a2_planes = nullptr;
  long long a2_plane_stride = 0;
  int a2_K = 0;
  __half* o_planes = nullptr;
  long long o_plane_stride = 0;
  int o_planes_ld = 0, o_planes_relu = 0;
  // LayerNorm folded around the GEMM (tcgen05 engine, N = 512 producers / any-N consumers):
  //  consumer: out = rstd[m] * (acc - mean[m] * ln_s[n]) + ln_c[n]  with (mean, rstd) combined from ln_stats[m][16][2]
  //  producer: stats_out[m][n / 32][2] = (mean, M2) of each 32-column run of the result row (Chan-combinable)
  const float* ln_stats = nullptr;
  const float* ln_s = nullptr;
  const float* ln_c = nullptr;
  float* stats_out = nullptr;
  // Fused attention (tcgen05 engine): W is the qkv weight with rows permuted to [head][half][q 64 | k 64 | v 64], so that
  // CTA n owns 64 dims of q, k, v of one head; the epilogue runs the 32-token attention and writes o_planes [M, 512].
  int attn = 0;
  // trunk kernel only, row-tile signals: instead of waiting for the whole predecessor grid (griddepcontrol.wait) a
  // CTA of this launch waits until sig_in[its 128-row tile] has reached sig_expect (= the column CTAs of the producer layer), and when its
  // own results are in memory it adds 1 to sig_out[its row tile].  Every access of the layer must stay inside the row tile.
  const unsigned* sig_in = nullptr;
  int sig_expect = 0;
  unsigned* sig_out = nullptr;
};
int tc_fast_grid_x(const GemmP& p);       // column CTAs the trunk kernel would launch for p (0: not a trunk-kernel problem)

constexpr float kActScale = 16.0f;   // activations are stored in fp16 planes as (v * 16): |v| < 4e3 representable, lo normal for |v| > 2^-6

// state of the sampling loop that lives on the device so that ONE captured step graph serves every step
struct LoopState {
  int k;                 // current step index, S-1 .. 0
  int S;
  const float* tape;     // [S,B,1536,1,32] or null
  unsigned done;         // CTAs of the current step_update that have finished (the last one advances k)
};

GemmP linear(const float* A, int M, int K, const float* W, const float* bias, float* out, int N);

int gemm_simt(const GemmP& p, cudaStream_t s);
int gemm(const GemmP& p, cudaStream_t s);   // dispatches on the engine (TC falls back to SIMT for shapes it does not take)
bool tc_supported(const GemmP& p);
int gemm_tc(const GemmP& p, cudaStream_t s);
bool profiling();
int set_trace_kernels(unsigned long long* p);
int set_trace_tc(unsigned long long* p);
int fast_chain_begin(cudaStream_t s);   // collect the trunk launches of stream s into layer chains (gemm_tc_chain_kernel) ...
int fast_chain_end();                   // ... and launch what is pending
int fast_chain_flush();
int profile_begin();
int profile_end(double* ms, double* flops, int64_t* n);

// ---- other kernels --------------------------------------------------------------------------------------
// y (fp32) and / or planes (fp16 hi/lo of y * kActScale, [2][rows][512]) may be null
int layernorm512(const float* x, const float* gamma, const float* beta, float* y, __half* planes, int rows, cudaStream_t s);
int wav_first(const float* audio, const float* w1, const float* b1, const float* wd, const float* bd, int ldw, int cb, int Lin, int Lout,
              int stride, int pad, __half* h1_planes, long long plane_stride, float* sc, cudaStream_t s);   // block 0 of the WavEncoder: conv1 -> planes, conv shortcut -> fp32
int attention32(const float* qkv, float* out, __half* planes, int nseq, cudaStream_t s, long long plane_stride = 0);   // qkv [nseq*32,1536] -> out [nseq*32,512]
void tc_forget_weights(const float* W);
int tc_split(const float* a, int lda, int M, int K, __half* planes, cudaStream_t s);   // fp32 -> hi/lo planes of a*kActScale
// split scratch of the tcgen05 engine: one arena per stream (a GEMM whose operand is still fp32 splits it there first)
struct Arena;
size_t tc_scratch_need(const GemmP& p);
int tc_scratch_reserve(cudaStream_t s, size_t need, Arena** out = nullptr);
void tc_scratch_release(cudaStream_t s);
int advance_loop(LoopState* ls, cudaStream_t s);
int init_loop(LoopState* ls, int S, int k, const float* tape, cudaStream_t s);   // k: the step the loop stands at; tape + (S-1-k')*B*1536*32 = eps of step k'
struct TokensInP {
  const float* z;            // [B*32,512] = x_t . Wx^T
  const float* vt_table;     // [1000,512]
  const int64_t* t_dev;      // [B] or null
  int t_scalar;              // used when t_dev == null and ls == null
  const LoopState* ls;       // sampling loop: t = t_model_dev[ls->k]
  const int32_t* t_model_dev;
  const float* cst[ST_MAX_EVALS];   // per eval: [B*32,512] or [32,512] (cst_bcast) conditioning constant
  int cst_bcast[ST_MAX_EVALS];
  const float* g2;           // [B,512] seed term (always added)
  const float* sv[ST_MAX_EVALS];    // per eval: [B,512], [1,512] (sv_bcast) or null: style term
  int sv_bcast[ST_MAX_EVALS];
  const float* rope_cos;     // [32,32]
  const float* rope_sin;
  float* x;                  // [nE*B*32,512]
  __half* x_planes;          // optional: fp16 hi/lo planes of x (* kActScale), [2][nE*B*32][512]
  float* stats;              // optional: LayerNorm row statistics of x as 16 combinable partials, [nE*B*32][16][2]
  int B, nE;
  unsigned* sig_zero = nullptr;   // optional: row-tile signals of the trunk layers that follow, zeroed by block 0 ...
  int sig_n = 0;                  // ... (the launch has waited for everything before it, and everything after it waits for the launch)
};
int tokens_in(const TokensInP& p, cudaStream_t s);

// Sampling loop, deterministic DDIM on the tcgen05 engine: the loop carries z = W_x x_k instead of x_k ("z recursion",
// DESIGN.md §4).  One launch per step: z <- beta z + alpha (CFG-mix of P + c_xo) with the coefficients of the PREVIOUS
// step's update (skipped at the first step, where z comes from the input GEMM), then the tokens of this step like
// tokens_in.  The step's values are launch parameters (the captured graph holds the whole loop, one node per step).
struct TokensStepP {
  TokensInP t;               // conditioning terms, rotary tables, outputs (t.z, t.ls, t.t_dev, t.vt_table are not used)
  float* z_rw;               // [B*32,512] state, updated in place
  const float* P;            // [nE*B*32,512] = previous step's last-block rows through W_xo (W_xo2), eval-major
  const float* c_xo;         // [512] = W_x b_out
  const float* vt;           // [512] row t_model[k] of the timestep table
  float alpha, beta;         // x_k = alpha x0_hat + beta x_{k+1} + sigma eps_{k+1}
  float sigma;               // 0: deterministic step (zeps is not read)
  const float* zeps;         // [B*32,512] = W_x eps_{k+1} (DDPM / DDIM with eta > 0), or null
  int first;                 // first step of the loop: z is taken as it is
  int cfg_mode;              // ST_CFG_NONE / ST_CFG_TEXT / ST_CFG_TWO (evaluation order as in step_update)
  const float* scale;        // device [B]
  const float* scale2;       // device [B] (TWO)
};
int tokens_step(const TokensStepP& p, cudaStream_t s);

struct StepP {
  const float* o;            // [nE*B*32,1536] eval outputs, eval-major
  float* xs;                 // [B*32,1536] state, updated in place (or x0 written when mode<0)
  const float* eps;          // [B,1536,1,32] caller layout noise or null
  int B, nE;
  int cfg_mode;              // ST_CFG_*
  const float* scale;        // device [B] (TEXT: scale; TWO: scale_audio)
  const float* scale2;       // device [B] (TWO: scale_prompt)
  float part_sa[3], part_sp[3];  // BODYPART
  int part_ua[3];                // eval index of the part's prompt evaluation or -1
  int mode;                  // ST_MODE_DDPM / ST_MODE_DDIM / -1 = write combined model output only
  float c[ST_COEF_STRIDE];
  const LoopState* ls;       // when set: coefficients = coef_dev[ls->k], eps = ls->tape + (S-1-k) * B*1536*32
  const float* coef_dev;
  __half* xs_planes = nullptr;     // optional: fp16 hi/lo planes of the updated state, [2][B*32][1536]
  LoopState* ls_advance = nullptr; // optional: the last CTA decrements ls->k (replaces a separate advance launch)
};
int step_update(const StepP& p, cudaStream_t s);

int transpose_to_tokens(const float* x, float* tok, int B, int C, int T, float scale, cudaStream_t s);     // [B,C,T]->[B,T,C]
int transpose_from_tokens(const float* tok, float* x, int B, int C, int T, cudaStream_t s);                // [B,T,C]->[B,C,T]
int gather_words(const int32_t* word, const float* table, float* out, int ldo, int rows, int force_zero, int n_words, cudaStream_t s);
int avgpool4(const float* in, float* out, int rows_out, int cols, cudaStream_t s);                         // rows_out x cols, in has 4x rows
int vq_select(const float* dot, const float* cnorm, const float* codebook, float* residual, float* qsum, int64_t* idx,
              int idx_stride, int rows, int first, __half* r_planes, __half* q_planes, cudaStream_t s);   // planes: optional fp16 hi/lo copies of the new residual / of qsum
int copy_strided_scale(const float* in, long long in_stride, float scale, float* out, int rows, int cols, cudaStream_t s);
int pose330(const float* up, const float* ha, const float* lo, const float* mean, const float* std, const float* tmean,
            const float* tstd, const float* jaw, int B, int n, float* pose, float* trans, cudaStream_t s);
int pose623(const float* up, const float* ha, const float* lo, int B, int n, float* pose, cudaStream_t s);
int pose_aa165(const float* pose, long long frames, float* aa, cudaStream_t s);
int moments_accumulate(const float* x, long long N, int D, double* acc, cudaStream_t s);
int l1div_accumulate(const float* x, int n, int J, double* acc, cudaStream_t s);

// ---- device memory helpers ------------------------------------------------------------------------------
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  int reserve(size_t bytes);   // (re)allocates when cap < bytes; resets the bump pointer
  void reset() { off = 0; }
  template <typename T>
  T* take(size_t n) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + n * sizeof(T);
    return reinterpret_cast<T*>(base + a);
  }
  void release();
};

struct Weights {
  std::map<std::string, float*> dev;
  std::map<std::string, int64_t> numel;
  int upload(const st_tensor* t, int n);
  const float* get(const std::string& name, int64_t expect_numel, int* err) const;
  void release();
};

}  // namespace st
