// C-ABI entry points (include/syntalker_b200.h): handles, weight upload, conditioning cache, the denoiser
// evaluation, the sampling loop, RVQ decode, pose assembly and the host-buffer pipeline.
#include "st_internal.cuh"

#include <stdarg.h>
#include <string.h>

namespace st {

thread_local char g_err[1024] = "";
thread_local int64_t g_launches = 0;
bool g_pdl = true;
static int g_engine = ST_ENGINE_TC;   // process default: the product engine; SIMT is the exact-fp32 anchor (st_set_engine)
// The engine a call runs on belongs to the HANDLE it was made on (st_model_set_engine / st_vq_set_engine; -1 = follow the process
// default): every entry point opens an EngineScope for its handle, and the code below asks cur_engine().  A handle is driven by
// one host thread at a time, so a thread-local scope is exact, and two handles on two threads may use different engines.
static thread_local int t_engine = -1;
static int cur_engine() { return t_engine >= 0 ? t_engine : g_engine; }
struct EngineScope {
  int prev;
  explicit EngineScope(int e) : prev(t_engine) { if (e >= 0) t_engine = e; }
  ~EngineScope() { t_engine = prev; }
};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int Arena::reserve(size_t bytes) {
  off = 0;
  if (bytes <= cap) return ST_OK;
  if (base) { cudaDeviceSynchronize(); cudaFree(base); base = nullptr; cap = 0; }
  size_t want = bytes + (bytes >> 3) + (1 << 20);
  cudaError_t e = cudaMalloc(&base, want);
  if (e != cudaSuccess) {
    set_error("workspace cudaMalloc(%zu MiB) failed: %s", want >> 20, cudaGetErrorString(e));
    base = nullptr;
    return ST_ENOMEM;
  }
  cap = want;
  return ST_OK;
}
void Arena::release() {
  if (base) cudaFree(base);
  base = nullptr; cap = 0; off = 0;
}

int Weights::upload(const st_tensor* t, int n) {
  for (int i = 0; i < n; ++i) {
    ST_REQUIRE(t[i].name && t[i].data && t[i].numel > 0, "tensor %d is empty", i);
    float* d = nullptr;
    size_t bytes = (size_t)t[i].numel * sizeof(float);
    if (cudaMalloc(&d, (bytes + 255) & ~size_t(255)) != cudaSuccess) { set_error("cudaMalloc for %s failed", t[i].name); return ST_ENOMEM; }
    ST_CHECK_CUDA(cudaMemcpy(d, t[i].data, bytes, cudaMemcpyHostToDevice));
    dev[t[i].name] = d;
    numel[t[i].name] = t[i].numel;
  }
  return ST_OK;
}
const float* Weights::get(const std::string& name, int64_t expect, int* err) const {
  auto it = dev.find(name);
  if (it == dev.end()) { set_error("packed tensor '%s' is missing", name.c_str()); *err = ST_EINVAL; return nullptr; }
  if (expect > 0 && numel.at(name) != expect) {
    set_error("packed tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)numel.at(name), (long long)expect);
    *err = ST_EINVAL;
    return nullptr;
  }
  return it->second;
}
void Weights::release() {
  for (auto& kv : dev) { tc_forget_weights(kv.second); cudaFree(kv.second); }
  dev.clear(); numel.clear();
}

// ---- GEMM-engine profiling -----------------------------------------------------------------------------
// Between begin and end every GEMM-engine launch is recorded (its full descriptor; all buffers it names are
// workspace that outlives the step).  end() replays exactly that launch sequence as ONE captured CUDA graph
// between two events, so the measured time is device time of the GEMM kernels alone, not host launch latency
// (events around individual eager launches of 5 us kernels mostly measure the host).
static bool g_prof = false;
static std::vector<GemmP> g_prof_list;
static double g_prof_flops = 0.0;

static int gemm_dispatch(const GemmP& p, cudaStream_t s) {
  if (cur_engine() == ST_ENGINE_TC && tc_supported(p)) return gemm_tc(p, s);
  ST_TRY(fast_chain_flush());      // an open layer chain leaves before anything else is launched
  return gemm_simt(p, s);
}

int gemm(const GemmP& p, cudaStream_t s) {
  if (g_prof) {
    g_prof_list.push_back(p);
    g_prof_flops += 2.0 * (double)p.M * (double)p.N * (double)p.K;
  }
  return gemm_dispatch(p, s);
}
// the replay of profile_end() chains what the recorded pass chained: markers (M = -1 begin, -2 end) in the launch list; and it zeroes
// the row-tile signals where the recorded pass's token prologue did (M = -3: a one-block kernel that, like the prologue, waits for
// everything before it) -- replayed without that, every consumer would find its counter already satisfied and run ahead of its producers
static void profile_mark(int what, unsigned* sig = nullptr, int sig_n = 0) {
  if (!g_prof) return;
  GemmP m;
  m.M = what;
  m.sig_out = sig;
  m.N = sig_n;
  g_prof_list.push_back(m);
}
__global__ void zero_u32_kernel(unsigned* p, int n) {
  pdl_wait();
  pdl_launch();
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0u;
}

bool profiling() { return g_prof; }

int profile_begin() {
  g_prof_list.clear();
  g_prof_flops = 0.0;
  g_prof = true;
  return ST_OK;
}
int profile_end(double* ms, double* flops, int64_t* n) {
  g_prof = false;
  ST_CHECK_CUDA(cudaDeviceSynchronize());
  if (flops) *flops = g_prof_flops;
  if (n) { *n = 0; for (const GemmP& p : g_prof_list) *n += p.M >= 0; }
  if (ms) *ms = 0.0;
  if (g_prof_list.empty()) return ST_OK;
  cudaStream_t s = nullptr;
  ST_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  const int64_t l0 = g_launches;
  // the replay stream's split scratch is sized for the largest fp32 operand before the capture starts
  size_t need = 0;
  for (const GemmP& p : g_prof_list)
    if (p.M >= 0 && cur_engine() == ST_ENGINE_TC && tc_supported(p)) need = std::max(need, tc_scratch_need(p));
  if (need) ST_TRY(tc_scratch_reserve(s, need));
  ST_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  int r = ST_OK;
  for (const GemmP& p : g_prof_list) {
    if (p.M == -3) { launch_k(zero_u32_kernel, dim3(1), dim3(256), 0, s, p.sig_out, p.N); continue; }
    r = p.M == -1 ? fast_chain_begin(s) : p.M == -2 ? fast_chain_end() : gemm_dispatch(p, s);
    if (r != ST_OK) break;
  }
  if (r != ST_OK) fast_chain_end();
  cudaError_t ce = cudaStreamEndCapture(s, &graph);
  g_launches = l0;
  if (r != ST_OK) { if (graph) cudaGraphDestroy(graph); tc_scratch_release(s); cudaStreamDestroy(s); (void)cudaGetLastError(); return r; }
  ST_CHECK_CUDA(ce);
  ST_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaEvent_t e0, e1;
  ST_CHECK_CUDA(cudaEventCreate(&e0));
  ST_CHECK_CUDA(cudaEventCreate(&e1));
  ST_CHECK_CUDA(cudaGraphLaunch(exec, s));            // warm
  ST_CHECK_CUDA(cudaEventRecord(e0, s));
  ST_CHECK_CUDA(cudaGraphLaunch(exec, s));
  ST_CHECK_CUDA(cudaEventRecord(e1, s));
  ST_CHECK_CUDA(cudaStreamSynchronize(s));
  float t = 0.f;
  ST_CHECK_CUDA(cudaEventElapsedTime(&t, e0, e1));
  if (ms) *ms = t;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); tc_scratch_release(s); cudaStreamDestroy(s);
  g_prof_list.clear();
  return ST_OK;
}

}  // namespace st

using namespace st;

// WavEncoder geometry (models/denoiser.py:308-315): C_in, C_out, stride, pad of the first conv, conv shortcut
static const int kWav[6][5] = {{2, 64, 5, 1700, 1}, {64, 64, 6, 0, 1}, {64, 64, 1, 7, 0}, {64, 128, 6, 0, 1}, {128, 128, 1, 7, 0}, {128, 256, 3, 0, 1}};
static const int kWavLen[7] = {68224, 14322, 2385, 2385, 396, 396, 128};
static const int kCondChunk = 32;

struct ConvW { const float* w; const float* b; int ldw; };
struct BlkW {
  const float *ln1g, *ln1b, *qkv, *projw, *projb, *ln2g, *ln2b, *fc1w, *fc1b, *fc2w, *fc2b;
  const float *qkv_wg, *qkv_s, *qkv_c, *fc1_wg, *fc1_s, *fc1_c;   // LayerNorm folded into the consuming Linear (tcgen05 engine)
  const float *qkv_wgp, *qkv_sp, *qkv_cp;                         // the same with rows permuted for the fused attention epilogue
};
struct EvalSpec { int cst_null; int sv_src; };   // sv_src: -1 none, 0..2 = style k, 3 = null embedding

struct Plan {
  int nE = 1;
  EvalSpec ev[ST_MAX_EVALS];
  int cfg_mode = ST_CFG_NONE;
  float part_sa[3] = {0, 0, 0}, part_sp[3] = {0, 0, 0};
  int part_ua[3] = {-1, -1, -1};
};

struct st_model {
  int variant = 0, device = 0, style_dim = 0;
  int engine = -1;                  // -1: the process default (st_set_engine) at the time of each call
  Weights w;
  ConvW wav[6][3];   // conv1, conv2, ds
  const float *word_table = nullptr, *w_cm = nullptr, *w_seed = nullptr, *bias_all = nullptr, *w_x = nullptr, *vt_table = nullptr;
  const float *w_style = nullptr, *null_sv = nullptr, *rope_cos = nullptr, *rope_sin = nullptr, *out_w = nullptr, *out_b = nullptr;
  const float *w_xo = nullptr, *c_xo = nullptr;   // W_x W_out [512,512] and W_x b_out [512] (optional: z recursion of deterministic DDIM)
  const float *w_xo2 = nullptr, *c_xo2 = nullptr; // [W_xo | W_xo W_fc2(block 7)] [512,1536] and W_xo b_fc2 [512] (optional: the last fc2 folded into that GEMM)
  BlkW blk[8];
  // constants for audio-masked evaluations (h3d): computed lazily
  float* cst_null = nullptr;   // [32,512]
  bool null_ready = false;
  // workspace
  Arena ws, io, longws;
  struct HostSlot { Arena stage; cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_done = nullptr; bool pending = false; };
  HostSlot hslot[2];                        // staging sets of the host-buffer pipeline (two window batches in flight)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_enter = nullptr;
  int ws_B = 0;
  float *cst_real = nullptr, *g2 = nullptr, *sv[3] = {nullptr, nullptr, nullptr};
  bool have_style[3] = {false, false, false};
  int cond_B = 0;
  float *xs = nullptr, *comb = nullptr, *z = nullptr, *X = nullptr, *H = nullptr, *ATT = nullptr, *QKV = nullptr, *G = nullptr, *O = nullptr;
  float *wavbuf[4] = {nullptr, nullptr, nullptr, nullptr}, *atcat = nullptr, *pooled = nullptr;
  __half* wavp[3] = {nullptr, nullptr, nullptr};   // WavEncoder activations as fp16 hi/lo planes (tcgen05 engine), 16 slack rows each
  float *scale_dev = nullptr, *scale2_dev = nullptr;
  std::vector<float> scale_last, scale2_last;          // host copies of what scale_dev / scale2_dev hold
  unsigned long long sched_id = 0;                     // schedule whose tables t_model_dev / coef_dev hold
  int64_t* t_tmp = nullptr;
  // fp16 hi/lo operand planes for the tcgen05 engine
  __half *H_p = nullptr, *ATT_p = nullptr, *G_p = nullptr, *X_p = nullptr, *xs_p = nullptr;
  float* ln_stats = nullptr;   // [nE*B*32][16][2] partial (mean, M2) over 32 columns each of the residual rows
  unsigned* sig = nullptr;     // row-tile signals of the trunk layers of one evaluation pass: [kSigLayers][sig_rt] counters, zeroed by the token prologue
  int sig_rt = 0;
  // sampling-loop state on the device + one captured step graph per (plan, mode, engine)
  LoopState* loop = nullptr;
  int32_t* t_model_dev = nullptr;
  float* coef_dev = nullptr;
  cudaStream_t loop_stream = nullptr;       // graphs cannot be captured on the legacy default stream
  cudaStream_t chain_stream = nullptr;      // second evaluation chain of a step (experiment, st_debug_probe bit 4096)
  cudaEvent_t ev_cfork = nullptr, ev_cjoin = nullptr;
  cudaStream_t dec_stream[2] = {nullptr, nullptr};   // the three body parts decode side by side (fork / join around st_rvq_decode x3)
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // the sampling loop in flight between st_sample_begin and st_sample_end
  struct ActiveLoop {
    bool on = false;
    const st_schedule* sc = nullptr;
    Plan pl;
    StepP sp;
    int B = 0, G = 1, k = -1;      // k: next step to run, S-1 .. 0; -1 when the loop is complete
    bool zrec = false, any_sigma = false;
    std::string key;               // graph key of this loop without the chunk index
  } act;
  Arena zws;                        // z recursion with noise: W_x eps of one chunk of steps + the staging of its tape
  float *zeps = nullptr, *zcarry[2] = {nullptr, nullptr}, *ntok = nullptr;
  int zws_B = 0, zws_G = 0, ntok_steps = 0;
  std::map<std::string, cudaGraphExec_t> graphs;
  std::map<std::string, int64_t> graph_nodes;
  std::map<std::string, int> warmed;
};

static bool g_use_graphs = true;
static bool g_rank_simt = false;
static bool g_decode_streams = true;   // st_debug_probe bit 256: decode the three body parts one after the other
static bool g_wav_planes = true;   // st_debug_probe bit 128: WavEncoder with fp32 activations and a split pass per conv
static bool g_fused_attn = true;   // st_debug_probe bit 32 turns the fused qkv + attention kernel off
static bool g_attn_tc = true;      // QK^T and PV of the fused attention as tcgen05 MMAs (attn == 2); st_debug_probe bit 2048 selects the packed-fp32 FMA epilogue
static bool g_zrec = true;         // st_debug_probe bit 512 turns the z recursion of deterministic DDIM off (state kept in x space)
static bool g_dual_chain = false;  // st_debug_probe bit 4096 (experiment): the evaluations of a step as two chains on two streams
static bool g_zrec_fc2 = true;     // st_debug_probe bit 1024: the last block's fc2 stays a layer of its own inside the z recursion
static bool g_row_sig = true;      // st_debug_probe bit 262144: trunk layers wait for the whole predecessor grid instead of row-tile signals

namespace st { extern int g_tc_probe; extern bool g_tc_fast; extern bool g_tc_taps; extern bool g_tc_chain; }
extern "C" int st_debug_probe(int flags) {
  st::g_tc_taps = !(flags & 8192);
  st::g_tc_chain = (flags & 131072) != 0;
  g_row_sig = !(flags & 262144);
  st::g_tc_probe = (flags & 15) | ((flags & (16384 | 32768 | 65536)) >> 10);   // bits 16384.. reach the kernels as probe bits 16, 32, 64 (experiments)
  st::g_tc_fast = !(flags & 16);
  g_fused_attn = !(flags & 32);
  g_rank_simt = (flags & 64) != 0;
  g_decode_streams = !(flags & 256);
  g_wav_planes = !(flags & 128);
  g_zrec = !(flags & 512);
  g_attn_tc = !(flags & 2048);
  g_dual_chain = (flags & 4096) != 0;
  g_zrec_fc2 = !(flags & 1024);
  return ST_OK;
}

struct st_schedule {
  unsigned long long id = 0;     // unique per st_schedule_create (a destroyed schedule's address may be reused)
  int S = 0, mode = 0;
  std::vector<int32_t> t_model;
  std::vector<float> coef;
};

struct st_vq {
  int out_dim = 0, ldw_last = 0;
  int engine = -1;
  Weights w;
  const float *cb[6], *cnorm[6];
  ConvW c0, res1[2][3], res2[2][3], up[2], up_eo[2][2], c4, c6;
  bool has_enc = false;
  ConvW e0, edown[2], eres1[2][3], eres2[2][3], e4;    // encoder side (encdec.py:5-34)
  Arena ws;
};

// ---------------------------------------------------------------------------------------------------------
extern "C" const char* st_last_error(void) { return g_err; }
extern "C" int st_abi_version(void) { return ST_ABI_VERSION; }
extern "C" int64_t st_launch_count(void) { return g_launches; }
extern "C" int st_set_engine(int engine) {
  ST_REQUIRE(engine == ST_ENGINE_SIMT || engine == ST_ENGINE_TC, "unknown engine %d", engine);
  g_engine = engine;
  return ST_OK;
}
extern "C" int st_get_engine(void) { return g_engine; }
namespace st { extern long long* g_tc_dbg; extern int g_tc_probe; extern bool g_tc_fast; }
extern "C" int st_debug_probe(int flags);
namespace st { extern int g_tc_dbg_n, g_tc_dbg_k; }
extern "C" int st_debug_timeline(long long* dev_buf) { st::g_tc_dbg = dev_buf; return ST_OK; }
extern "C" int st_debug_timeline_select(int N, int K) { st::g_tc_dbg_n = N; st::g_tc_dbg_k = K; return ST_OK; }
namespace st { extern int g_tc_dbg_n2, g_tc_dbg_k2; }
extern "C" int st_debug_timeline_select2(int N, int K) { st::g_tc_dbg_n2 = N; st::g_tc_dbg_k2 = K; return ST_OK; }
extern "C" int st_debug_trace(unsigned long long* dev_buf) {
  ST_TRY(st::set_trace_kernels(dev_buf));
  return st::set_trace_tc(dev_buf);
}
extern "C" int st_set_pdl(int on) { st::g_pdl = on != 0; return ST_OK; }
extern "C" int st_set_graphs(int on) { g_use_graphs = on != 0; return ST_OK; }
extern "C" int st_profile_begin(void) { return st::profile_begin(); }
extern "C" int st_profile_end(double* ms_total, double* flops_total, int64_t* launches) { return st::profile_end(ms_total, flops_total, launches); }

// ---------------------------------------------------------------------------------------------------------
static int model_resolve(st_model* m) {
  int err = ST_OK;
  char nm[96];
  const char* cname[3] = {"conv1", "conv2", "ds"};
  for (int i = 0; i < 6; ++i) {
    const int cin = kWav[i][0], cout = kWav[i][1];
    for (int c = 0; c < 3; ++c) {
      m->wav[i][c] = ConvW{nullptr, nullptr, 0};
      if (c == 2 && !kWav[i][4]) continue;
      const int cc = (c == 1) ? cout : cin;
      const int K = 15 * cc, ldw = (K + 3) & ~3;
      snprintf(nm, sizeof nm, "wav.%d.%s.w", i, cname[c]);
      m->wav[i][c].w = m->w.get(nm, (int64_t)cout * ldw, &err);
      snprintf(nm, sizeof nm, "wav.%d.%s.b", i, cname[c]);
      m->wav[i][c].b = m->w.get(nm, cout, &err);
      m->wav[i][c].ldw = ldw;
    }
  }
  m->word_table = m->w.get("word_table", 0, &err);
  m->w_cm = m->w.get("w_cm", 512 * 512, &err);
  m->w_seed = m->w.get("w_seed", 512 * 6144, &err);
  m->bias_all = m->w.get("bias_all", 512, &err);
  m->w_x = m->w.get("w_x", 512 * 1536, &err);
  m->vt_table = m->w.get("vt_table", ST_MAX_T * 512, &err);
  m->rope_cos = m->w.get("rope_cos", 32 * 32, &err);
  m->rope_sin = m->w.get("rope_sin", 32 * 32, &err);
  m->out_w = m->w.get("out.w", 1536 * 512, &err);
  m->out_b = m->w.get("out.b", 1536, &err);
  if (m->w.dev.count("w_xo") && m->w.dev.count("c_xo")) {
    m->w_xo = m->w.get("w_xo", 512 * 512, &err);
    m->c_xo = m->w.get("c_xo", 512, &err);
  }
  if (m->w_xo && m->w.dev.count("w_xo2") && m->w.dev.count("c_xo2")) {
    m->w_xo2 = m->w.get("w_xo2", 512 * 1536, &err);
    m->c_xo2 = m->w.get("c_xo2", 512, &err);
  }
  m->style_dim = m->variant == ST_VARIANT_BEATX_MOTIONCLIP ? 512 : m->variant == ST_VARIANT_H3D ? 256 : 0;
  if (m->style_dim) m->w_style = m->w.get("w_style", 512 * m->style_dim, &err);
  if (m->variant == ST_VARIANT_H3D) m->null_sv = m->w.get("null_sv", 512, &err);
  for (int i = 0; i < 8; ++i) {
    BlkW& b = m->blk[i];
    auto g = [&](const char* suffix, int64_t n) { snprintf(nm, sizeof nm, "blk.%d.%s", i, suffix); return m->w.get(nm, n, &err); };
    b.ln1g = g("ln1.g", 512); b.ln1b = g("ln1.b", 512); b.qkv = g("qkv.w", 1536 * 512);
    b.projw = g("proj.w", 512 * 512); b.projb = g("proj.b", 512);
    b.ln2g = g("ln2.g", 512); b.ln2b = g("ln2.b", 512);
    b.fc1w = g("fc1.w", 1024 * 512); b.fc1b = g("fc1.b", 1024);
    b.fc2w = g("fc2.w", 512 * 1024); b.fc2b = g("fc2.b", 512);
    b.qkv_wg = g("qkv.wg", 1536 * 512); b.qkv_s = g("qkv.s", 1536); b.qkv_c = g("qkv.c", 1536);
    b.qkv_wgp = g("qkv.wgp", 1536 * 512); b.qkv_sp = g("qkv.sp", 1536); b.qkv_cp = g("qkv.cp", 1536);
    b.fc1_wg = g("fc1.wg", 1024 * 512); b.fc1_s = g("fc1.s", 1024); b.fc1_c = g("fc1.c", 1024);
  }
  return err;
}

extern "C" int st_model_create(const st_tensor* packed, int n, int variant, st_model** out) {
  ST_REQUIRE(packed && out && n > 0, "st_model_create: null argument");
  ST_REQUIRE(variant >= 0 && variant <= 2, "st_model_create: unknown variant %d", variant);
  st_model* m = new st_model();
  m->variant = variant;
  cudaGetDevice(&m->device);
  int r = m->w.upload(packed, n);
  if (r == ST_OK) r = model_resolve(m);
  if (r != ST_OK) { m->w.release(); delete m; return r; }
  *out = m;
  return ST_OK;
}

// engine of ONE handle: ST_ENGINE_SIMT / ST_ENGINE_TC, or -1 to follow the process default again
extern "C" int st_model_set_engine(st_model* m, int engine) {
  ST_REQUIRE(m && (engine == -1 || engine == ST_ENGINE_SIMT || engine == ST_ENGINE_TC), "st_model_set_engine: bad argument");
  m->engine = engine;
  return ST_OK;
}
extern "C" int st_vq_set_engine(st_vq* v, int engine) {
  ST_REQUIRE(v && (engine == -1 || engine == ST_ENGINE_SIMT || engine == ST_ENGINE_TC), "st_vq_set_engine: bad argument");
  v->engine = engine;
  return ST_OK;
}

extern "C" void st_model_destroy(st_model* m) {
  if (!m) return;
  cudaDeviceSynchronize();
  for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);
  if (m->loop_stream) { tc_scratch_release(m->loop_stream); cudaStreamDestroy(m->loop_stream); }
  if (m->dec_stream[0]) { for (int k = 0; k < 2; ++k) { tc_scratch_release(m->dec_stream[k]); cudaStreamDestroy(m->dec_stream[k]); cudaEventDestroy(m->ev_join[k]); } cudaEventDestroy(m->ev_fork); }
  if (m->chain_stream) { tc_scratch_release(m->chain_stream); cudaStreamDestroy(m->chain_stream); cudaEventDestroy(m->ev_cfork); cudaEventDestroy(m->ev_cjoin); }
  if (m->ev_in) cudaEventDestroy(m->ev_in);
  if (m->ev_out) cudaEventDestroy(m->ev_out);
  m->w.release(); m->ws.release(); m->io.release(); m->longws.release(); m->zws.release();
  for (int k = 0; k < 2; ++k) {
    m->hslot[k].stage.release();
    if (m->hslot[k].ev_h2d) { cudaEventDestroy(m->hslot[k].ev_h2d); cudaEventDestroy(m->hslot[k].ev_comp); cudaEventDestroy(m->hslot[k].ev_done); }
  }
  if (m->h2d_stream) { cudaStreamDestroy(m->h2d_stream); cudaStreamDestroy(m->d2h_stream); cudaEventDestroy(m->ev_enter); }
  if (m->cst_null) cudaFree(m->cst_null);
  delete m;
}

constexpr size_t kSigLayers = 40;     // >= trunk launches of one evaluation pass (8 blocks x 4 + the GEMM that ends it)
static int model_workspace(st_model* m, int B) {
  if (B <= m->ws_B) return ST_OK;
  const size_t rows = (size_t)B * 32, nE = ST_MAX_EVALS;
  const int cb = B < kCondChunk ? B : kCondChunk;
  size_t f = 0;   // floats
  f += rows * 512 + (size_t)B * 512 * 4;                       // cst_real, g2, sv[3]
  f += rows * 1536 * 2 + rows * 512;                           // xs, comb, z
  f += nE * rows * (512 * 3 + 1536 * 2 + 1024);                // X,H,ATT,QKV,O,G
  f += 4 * (size_t)cb * kWavLen[1] * 64 + (size_t)cb * 128 * 512 + (size_t)cb * 32 * 512;
  f += 3 * ((size_t)cb * kWavLen[1] + 16) * 64 + 3 * 64;      // wavp: 2 halves = 1 float per element
  f += 2 * (size_t)B + 64;
  f += nE * rows * (512 * 3 + 1024) + rows * 1536;             // fp16 hi+lo planes H_p, ATT_p, X_p, G_p, xs_p (2 halves = 1 float each)
  f += 1000 * (1 + ST_COEF_STRIDE) + 64 + nE * rows * 32;
  const size_t sig_rt = (nE * rows + 127) / 128;
  f += kSigLayers * sig_rt;
  size_t bytes = f * sizeof(float) + (size_t)B * sizeof(int64_t) + 64 * 256;
  for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);  // captured pointers die with the old block
  m->graphs.clear(); m->warmed.clear(); m->graph_nodes.clear();
  ST_TRY(m->ws.reserve(bytes));
  Arena& a = m->ws;
  m->cst_real = a.take<float>(rows * 512);
  m->g2 = a.take<float>((size_t)B * 512);
  for (int k = 0; k < 3; ++k) m->sv[k] = a.take<float>((size_t)B * 512);
  m->xs = a.take<float>(rows * 1536);
  m->comb = a.take<float>(rows * 1536);
  m->z = a.take<float>(rows * 512);
  m->X = a.take<float>(nE * rows * 512);
  m->H = a.take<float>(nE * rows * 512);
  m->ATT = a.take<float>(nE * rows * 512);
  m->QKV = a.take<float>(nE * rows * 1536);
  m->G = a.take<float>(nE * rows * 1024);
  m->O = a.take<float>(nE * rows * 1536);
  for (int k = 0; k < 4; ++k) m->wavbuf[k] = a.take<float>((size_t)cb * kWavLen[1] * 64);
  for (int k = 0; k < 3; ++k) m->wavp[k] = a.take<__half>(2 * ((size_t)cb * kWavLen[1] + 16) * 64);
  m->atcat = a.take<float>((size_t)cb * 128 * 512);
  m->pooled = a.take<float>((size_t)cb * 32 * 512);
  m->scale_dev = a.take<float>(B);
  m->scale2_dev = a.take<float>(B);
  m->t_tmp = a.take<int64_t>(B);
  m->H_p = a.take<__half>(2 * nE * rows * 512);
  m->ATT_p = a.take<__half>(2 * nE * rows * 512);
  m->X_p = a.take<__half>(2 * nE * rows * 512);
  m->G_p = a.take<__half>(2 * nE * rows * 1024);
  m->xs_p = a.take<__half>(2 * rows * 1536);
  m->ln_stats = a.take<float>(nE * rows * 32);
  m->sig = a.take<unsigned>(kSigLayers * sig_rt);
  m->sig_rt = (int)sig_rt;
  ST_CHECK_CUDA(cudaMemsetAsync(m->sig, 0, kSigLayers * sig_rt * sizeof(unsigned), nullptr));
  m->loop = a.take<LoopState>(1);
  m->t_model_dev = a.take<int32_t>(1000);
  m->coef_dev = a.take<float>(1000 * ST_COEF_STRIDE);
  m->ws_B = B;
  m->cond_B = 0;   // the cache lived in the old block
  m->scale_last.clear(); m->scale2_last.clear(); m->sched_id = 0;
  return ST_OK;
}

// WavEncoder + word path + mix + pool for `cb` clips -> pooled [cb*32, 512]   (denoiser.py:151-157)
// WavEncoder on the tcgen05 engine with the activations chained as fp16 hi/lo planes: every conv streams the planes its
// producer's epilogue wrote, the conv shortcut and conv1 of a block share one operand, and an fp32 copy is written only
// where it is a residual (identity shortcuts of blocks 2 and 4, the conv shortcuts' outputs) or the result.  Only the
// raw-audio convs of block 0 (C_in = 2) are one exact-fp32 direct-convolution pass (wav_first_kernel).
static int encode_wav_tc(st_model* m, const float* audio, int cb, cudaStream_t s) {
  const float* in_f = audio;          // fp32 view of the block input (null when nobody needs it)
  const __half* in_p = nullptr;       // plane view of the block input
  long long in_ps = 0;
  int pin = -1, fin = -1;             // which wavp / wavbuf the block input occupies
  for (int i = 0; i < 6; ++i) {
    const int cin = kWav[i][0], cout = kWav[i][1], stride = kWav[i][2], pad = kWav[i][3], ds = kWav[i][4];
    const int Lin = kWavLen[i], Lout = kWavLen[i + 1];
    const bool last = (i == 5);
    const bool next_identity = !last && !kWav[i + 1][4];          // the next block adds its input back: keep an fp32 copy
    int fb[3], nf = 0, pb[2], np = 0;
    for (int k = 0; k < 4 && nf < 3; ++k) if (k != fin) fb[nf++] = k;
    for (int k = 0; k < 3 && np < 2; ++k) if (k != pin) pb[np++] = k;
    float* h1_f = m->wavbuf[fb[0]];
    float* sc = m->wavbuf[fb[1]];
    float* out_f = last ? m->atcat : m->wavbuf[fb[2]];
    __half* h1_p = m->wavp[pb[0]];
    __half* out_p = m->wavp[pb[1]];
    const long long h1_ps = ((long long)cb * Lout + 16) * cout, out_ps = h1_ps;
    GemmP p;
    p.A = in_f; p.W = m->wav[i][0].w; p.bias = m->wav[i][0].b;
    p.M = cb * Lout; p.N = cout; p.K = 15 * cin; p.ldw = m->wav[i][0].ldw;
    p.Lout = Lout; p.Lin = Lin; p.C = cin; p.stride = stride; p.pad = pad; p.dil = 1;
    p.a_batch = (long long)Lin * cin; p.lda = cin; p.ldo = cout; p.act = ACT_LRELU;
    GemmP q = p;                                                   // conv shortcut: same operand, same geometry
    q.W = m->wav[i][2].w; q.bias = m->wav[i][2].b; q.ldw = m->wav[i][2].ldw; q.out = sc; q.act = ACT_NONE;
    if (i == 0) {
      // raw-audio side: conv1 and the conv shortcut in one exact-fp32 pass (30-tap dot products), conv1 straight to planes
      ST_TRY(wav_first(audio, m->wav[0][0].w, m->wav[0][0].b, m->wav[0][2].w, m->wav[0][2].b, m->wav[0][0].ldw, cb, Lin, Lout, stride, pad,
                       h1_p, h1_ps, sc, s));
    } else {
      p.a_planes = in_p; p.a_plane_stride = in_ps; q.a_planes = in_p; q.a_plane_stride = in_ps;
      p.o_planes = h1_p; p.o_plane_stride = h1_ps; p.o_planes_ld = cout;
    }
    if (i > 0) ST_TRY(gemm(p, s));
    const float* shortcut = in_f;
    if (ds) { if (i > 0) ST_TRY(gemm(q, s)); shortcut = sc; }
    GemmP c2;
    c2.A = h1_f; c2.W = m->wav[i][1].w; c2.bias = m->wav[i][1].b;
    c2.M = cb * Lout; c2.N = cout; c2.K = 15 * cout; c2.ldw = m->wav[i][1].ldw;
    c2.Lout = Lout; c2.Lin = Lout; c2.C = cout; c2.stride = 1; c2.pad = 7; c2.dil = 1;
    c2.a_batch = (long long)Lout * cout; c2.lda = cout; c2.ldo = last ? 512 : cout;
    c2.res = shortcut; c2.res_mode = RES_PRE; c2.ldr = cout; c2.res_div = 1; c2.act = ACT_LRELU;
    c2.A = nullptr; c2.a_planes = h1_p; c2.a_plane_stride = h1_ps;
    c2.out = (last || next_identity) ? out_f : nullptr;
    if (!last) { c2.o_planes = out_p; c2.o_plane_stride = out_ps; c2.o_planes_ld = cout; }
    ST_TRY(gemm(c2, s));
    in_f = (last || next_identity) ? out_f : nullptr;
    fin = (last || !next_identity) ? -1 : fb[2];
    in_p = out_p; in_ps = out_ps; pin = pb[1];
  }
  return ST_OK;
}

static int encode_audio_words(st_model* m, const float* audio, const int32_t* word, int cb, int null_inputs, cudaStream_t s) {
  if (cur_engine() == ST_ENGINE_TC && g_wav_planes) {
    ST_TRY(encode_wav_tc(m, audio, cb, s));
    ST_TRY(gather_words(word, m->word_table, m->atcat + 256, 512, cb * 128, null_inputs, (int)(m->w.numel.at("word_table") / 256), s));
    ST_TRY(avgpool4(m->atcat, m->pooled, cb * 32, 512, s));
    return ST_OK;
  }
  const float* in = audio;
  int in_buf = -1;
  for (int i = 0; i < 6; ++i) {
    const int cin = kWav[i][0], cout = kWav[i][1], stride = kWav[i][2], pad = kWav[i][3], ds = kWav[i][4];
    const int Lin = kWavLen[i], Lout = kWavLen[i + 1];
    int free_b[3], nf = 0;
    for (int k = 0; k < 4 && nf < 3; ++k) if (k != in_buf) free_b[nf++] = k;
    float* h1 = m->wavbuf[free_b[0]];
    float* sc = m->wavbuf[free_b[1]];
    const bool last = (i == 5);
    float* outb = last ? m->atcat : m->wavbuf[free_b[2]];
    GemmP p;
    p.A = in; p.W = m->wav[i][0].w; p.bias = m->wav[i][0].b; p.out = h1;
    p.M = cb * Lout; p.N = cout; p.K = 15 * cin; p.ldw = m->wav[i][0].ldw;
    p.Lout = Lout; p.Lin = Lin; p.C = cin; p.stride = stride; p.pad = pad; p.dil = 1;
    p.a_batch = (long long)Lin * cin; p.lda = cin; p.ldo = cout; p.act = ACT_LRELU;
    ST_TRY(gemm(p, s));
    const float* shortcut = in;
    if (ds) {
      GemmP q = p;
      q.W = m->wav[i][2].w; q.bias = m->wav[i][2].b; q.ldw = m->wav[i][2].ldw; q.out = sc; q.act = ACT_NONE;
      ST_TRY(gemm(q, s));
      shortcut = sc;
    }
    GemmP c2;
    c2.A = h1; c2.W = m->wav[i][1].w; c2.bias = m->wav[i][1].b; c2.out = outb;
    c2.M = cb * Lout; c2.N = cout; c2.K = 15 * cout; c2.ldw = m->wav[i][1].ldw;
    c2.Lout = Lout; c2.Lin = Lout; c2.C = cout; c2.stride = 1; c2.pad = 7; c2.dil = 1;
    c2.a_batch = (long long)Lout * cout; c2.lda = cout; c2.ldo = last ? 512 : cout;
    c2.res = shortcut; c2.res_mode = RES_PRE; c2.ldr = cout; c2.res_div = 1; c2.act = ACT_LRELU;
    ST_TRY(gemm(c2, s));
    in = outb;
    in_buf = last ? -1 : free_b[2];
  }
  ST_TRY(gather_words(word, m->word_table, m->atcat + 256, 512, cb * 128, null_inputs, (int)(m->w.numel.at("word_table") / 256), s));
  ST_TRY(avgpool4(m->atcat, m->pooled, cb * 32, 512, s));
  return ST_OK;
}

static int ensure_null_consts(st_model* m, cudaStream_t s) {
  if (m->null_ready) return ST_OK;
  // audio := 0 before the WavEncoder and word ids := 0 (denoiser_h3d.py:173-179): a per-model constant [32,512]
  if (!m->cst_null) ST_CHECK_CUDA(cudaMalloc(&m->cst_null, 32 * 512 * sizeof(float)));
  float* zeros = m->wavbuf[3];   // not touched by block 0 when in_buf == -1 (uses buffers 0,1,2)
  ST_CHECK_CUDA(cudaMemsetAsync(zeros, 0, (size_t)ST_AUDIO_LEN * 2 * sizeof(float), s));
  ST_TRY(encode_audio_words(m, zeros, nullptr, 1, 1, s));
  GemmP p = linear(m->pooled, 32, 512, m->w_cm, m->bias_all, m->cst_null, 512);
  ST_TRY(gemm(p, s));
  m->null_ready = true;
  return ST_OK;
}

extern "C" int st_cond_encode(st_model* m, const st_cond* c, int B, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && c && B > 0, "st_cond_encode: null argument or B <= 0");
  ST_REQUIRE(c->audio && c->word && c->seed, "st_cond_encode: audio, word and seed are required");
  cudaStream_t s = (cudaStream_t)stream;
  ST_TRY(model_workspace(m, B));
  if (m->variant == ST_VARIANT_H3D) ST_TRY(ensure_null_consts(m, s));
  for (int b0 = 0; b0 < B; b0 += kCondChunk) {
    const int cb = (B - b0) < kCondChunk ? (B - b0) : kCondChunk;
    ST_TRY(encode_audio_words(m, c->audio + (size_t)b0 * ST_AUDIO_LEN * 2, c->word + (size_t)b0 * 128, cb, 0, s));
    GemmP p = linear(m->pooled, cb * 32, 512, m->w_cm, m->bias_all, m->cst_real + (size_t)b0 * 32 * 512, 512);
    ST_TRY(gemm(p, s));
  }
  GemmP ps = linear(c->seed, B, 6144, m->w_seed, nullptr, m->g2, 512);
  ST_TRY(gemm(ps, s));
  for (int k = 0; k < 3; ++k) {
    m->have_style[k] = false;
    if (m->style_dim && c->style[k]) {
      GemmP pv = linear(c->style[k], B, m->style_dim, m->w_style, nullptr, m->sv[k], 512);
      ST_TRY(gemm(pv, s));
      m->have_style[k] = true;
    }
  }
  m->cond_B = B;
  return ST_OK;
}

// Stage taps of the conditioning encoder for parity tests: the last st_cond_encode's [WavEncoder output |
// word features] rows, [B,128,512] (denoiser.py:151-155, before mix_audio_text), and the hoisted conditioning constant
// W_cm pool4(.) + biases, [B*32,512].  Either pointer may be NULL.  atcat covers one encoder chunk, so B <= 32.
extern "C" int st_debug_cond_taps(st_model* m, float* atcat_out, float* cst_out, int B, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && B > 0 && m->cond_B == B, "st_debug_cond_taps: the conditioning cache does not hold B=%d", B);
  cudaStream_t s = (cudaStream_t)stream;
  if (atcat_out) {
    ST_REQUIRE(B <= kCondChunk, "st_debug_cond_taps: the encoder staging holds one chunk of %d clips", kCondChunk);
    ST_CHECK_CUDA(cudaMemcpyAsync(atcat_out, m->atcat, (size_t)B * 128 * 512 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  if (cst_out) ST_CHECK_CUDA(cudaMemcpyAsync(cst_out, m->cst_real, (size_t)B * 32 * 512 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return ST_OK;
}

// ---- guidance -> evaluation list ---------------------------------------------------------------------
static int make_plan(const st_model* m, const st_guidance* g, Plan* pl) {
  const int mode = g ? g->mode : ST_CFG_NONE;
  const int flags = g ? g->flags : 0;
  const int v = m->variant;
  Plan& p = *pl;
  p = Plan();
  auto need_style = [&](int k) -> int {
    if (!m->have_style[k]) { set_error("guidance needs style[%d] but st_cond_encode did not receive it", k); return ST_ESTATE; }
    return ST_OK;
  };
  if (mode == ST_CFG_NONE || v == ST_VARIANT_BEATX) {
    // bare model; denoiser.MDM without motionclip never reads y['uncond'] (denoiser.py:144,172-174), so any
    // CFG wrapper around it degenerates to the conditional output exactly.
    ST_REQUIRE(mode == ST_CFG_NONE || mode == ST_CFG_TEXT, "variant beatx supports CFG_NONE / CFG_TEXT only");
    p.nE = 1; p.cfg_mode = ST_CFG_NONE;
    p.ev[0].cst_null = (v == ST_VARIANT_H3D) && (flags & ST_FLAG_UNCOND_AUDIO);
    if (v == ST_VARIANT_BEATX) p.ev[0].sv_src = -1;
    else if (flags & ST_FLAG_UNCOND) p.ev[0].sv_src = (v == ST_VARIANT_H3D) ? 3 : -1;
    else { ST_TRY(need_style(0)); p.ev[0].sv_src = 0; }
    return ST_OK;
  }
  if (mode == ST_CFG_TEXT) {
    ST_REQUIRE(g->scale, "CFG_TEXT needs scale[B]");
    ST_TRY(need_style(0));
    p.nE = 2; p.cfg_mode = ST_CFG_TEXT;
    const int an = (v == ST_VARIANT_H3D);          // the wrapper sets uncond_audio on both passes (cfg_sampler.py:18,22)
    p.ev[0] = EvalSpec{an, 0};
    p.ev[1] = EvalSpec{an, v == ST_VARIANT_H3D ? 3 : -1};
    return ST_OK;
  }
  ST_REQUIRE(v == ST_VARIANT_H3D, "CFG_TWO / CFG_BODYPART / CFG_BODYPART1 are defined for the h3d model only");
  if (mode == ST_CFG_BODYPART1) {
    // ClassifierFreeSampleModel_Bodypart (cfg_sampler.py:133-167): out_uncond = (null prompt, audio kept); per prompted part one
    // evaluation (that prompt, audio masked); unprompted parts keep out_uncond; out_uncond + scale * (out - out_uncond)
    ST_REQUIRE(g->scale, "CFG_BODYPART1 needs scale[B]");
    p.cfg_mode = ST_CFG_BODYPART1;
    p.ev[0] = EvalSpec{0, 3};
    p.nE = 1;
    for (int k = 0; k < 3; ++k) {
      p.part_ua[k] = -1;
      if (m->have_style[k]) { p.ev[p.nE] = EvalSpec{1, k}; p.part_ua[k] = p.nE; p.nE++; }
    }
    return ST_OK;
  }
  p.ev[0] = EvalSpec{1, 3};    // uu: prompt null, audio null
  p.ev[1] = EvalSpec{0, 3};    // ut: prompt null, audio real
  if (mode == ST_CFG_TWO) {
    ST_REQUIRE(g->scale && g->scale2, "CFG_TWO needs scale (audio) and scale2 (prompt)");
    ST_TRY(need_style(0));
    p.nE = 3; p.cfg_mode = ST_CFG_TWO;
    p.ev[2] = EvalSpec{1, 0};
    return ST_OK;
  }
  ST_REQUIRE(mode == ST_CFG_BODYPART, "unknown guidance mode %d", mode);
  p.cfg_mode = ST_CFG_BODYPART;
  p.nE = 2;
  const float a_s = g->audio_scale, p_s = g->prompt_scale;
  for (int k = 0; k < 3; ++k) {
    if (!m->have_style[k]) { p.part_sa[k] = a_s; p.part_sp[k] = 0.f; p.part_ua[k] = -1; continue; }
    p.part_sa[k] = (k == 0) ? 1.0f : 0.0f;         // cfg_sampler.py:100-105
    p.part_sp[k] = p_s;
    if (p_s != 0.f) { p.ev[p.nE] = EvalSpec{1, k}; p.part_ua[k] = p.nE; p.nE++; }
  }
  return ST_OK;
}

// ---- one trunk pass over the state m->xs for all planned evaluations -> m->O ---------------------------
// SIMT engine: every activation stays fp32.  tcgen05 engine: LayerNorm, attention and the GELU epilogue emit the
// fp16 hi/lo planes the next GEMM streams with TMA; only the residual stream, QKV and the outputs stay fp32.
// z = x . W_x^T for the state m->xs (tcgen05 engine: from the state planes m->xs_p)
static int trunk_input(st_model* m, int B, cudaStream_t s) {
  const int rows = B * 32;
  GemmP pz = linear(m->xs, rows, 1536, m->w_x, nullptr, m->z, 512);
  if (cur_engine() == ST_ENGINE_TC) { pz.a_planes = m->xs_p; pz.a_plane_stride = (long long)rows * 1536; }
  return gemm(pz, s);
}

// One step of the z recursion (sampling loop of deterministic DDIM on the tcgen05 engine): the step starts from z
// (tokens_step applies the previous step's update) and ends in P = X . W_xo^T instead of the output GEMM unless it is the last.
// The step's values are host-side here: the captured graph holds the whole loop, one set of nodes per step.
struct ZStep {
  bool on = false, first = false, last = true;
  int t = 0;                 // original timestep fed to the denoiser (respace.py:124-129)
  float alpha = 0.f, beta = 0.f, sigma = 0.f;
  const float* zeps = nullptr;   // W_x eps of the update that produced this step's state (sigma != 0 only)
};
// x_k = alpha x0_hat + beta x_{k+1} + sigma eps_{k+1}: the update that produced x_k ran with the coefficients of step k + 1.
//   DDIM (gaussian_diffusion.py:772-790): e = (a x - x0) / b, x <- c1 x0 + c2 e + sigma eps   =>  alpha = c1 - c2 / b, beta = c2 a / b
//   DDPM (:383, :556):                    x <- coef1 x0 + coef2 x + sigma eps                 =>  alpha = coef1, beta = coef2
static ZStep zstep_of(const st_schedule* sc, int k, const float* zeps_prev = nullptr) {
  ZStep z;
  z.on = true; z.first = (k == sc->S - 1); z.last = (k == 0); z.t = sc->t_model[k];
  if (!z.first) {
    const float* cf = &sc->coef[(size_t)(k + 1) * ST_COEF_STRIDE];
    if (sc->mode == ST_MODE_DDIM) {
      const double a = cf[0], b = cf[1], c1 = cf[2], c2 = cf[3];
      z.alpha = (float)(c1 - c2 / b);
      z.beta = (float)(c2 * a / b);
      z.sigma = cf[4];
    } else {
      z.alpha = cf[0]; z.beta = cf[1]; z.sigma = cf[2];
    }
    z.zeps = z.sigma != 0.f ? zeps_prev : nullptr;
  }
  return z;
}

static int run_trunk(st_model* m, const Plan& pl, int B, const int64_t* t_dev, int t_scalar, bool loop, cudaStream_t s,
                     const ZStep& zs = ZStep(), const StepP* mix = nullptr) {
  const bool zstep = zs.on, last = zs.last;
  const int rows = B * 32, R = pl.nE * rows;
  const bool tc = (cur_engine() == ST_ENGINE_TC);
  const long long ps512 = (long long)R * 512, ps1024 = (long long)R * 1024;
  const bool fold_fc2 = zstep && !last && g_zrec_fc2 && m->w_xo2 != nullptr;
  if (!zstep) {
    // the state lives in the model's own planes (a captured step graph must not point into the engine's split
    // scratch); inside the sampling loop step_update keeps them current, a single evaluation splits here
    if (tc && !loop) ST_TRY(tc_split(m->xs, 1536, rows, 1536, m->xs_p, s));
    ST_TRY(trunk_input(m, B, s));
  }
  TokensInP tp;
  tp.z = m->z; tp.vt_table = m->vt_table; tp.t_dev = t_dev; tp.t_scalar = t_scalar; tp.g2 = m->g2;
  tp.ls = loop ? m->loop : nullptr; tp.t_model_dev = m->t_model_dev;
  tp.rope_cos = m->rope_cos; tp.rope_sin = m->rope_sin; tp.x = m->X; tp.B = B; tp.nE = pl.nE;
  tp.x_planes = tc ? m->X_p : nullptr; tp.stats = tc ? m->ln_stats : nullptr;
  // row-tile signals between the trunk layers (tcgen05 engine, fused attention, one launch per layer, one chain of evaluations): the token
  // prologue zeroes the counters -- it has waited for everything before it and every layer after it waits for it
  const bool row_sig = tc && g_row_sig && g_fused_attn && !st::g_tc_chain && !(g_dual_chain && pl.nE >= 2 && loop) && m->sig != nullptr &&
                       (R + 127) / 128 <= m->sig_rt;
  if (row_sig) {
    tp.sig_zero = m->sig; tp.sig_n = (int)(kSigLayers * (size_t)m->sig_rt);
    profile_mark(-3, tp.sig_zero, tp.sig_n);
  }
  for (int e = 0; e < ST_MAX_EVALS; ++e) { tp.cst[e] = m->cst_real; tp.cst_bcast[e] = 0; tp.sv[e] = nullptr; tp.sv_bcast[e] = 0; }
  for (int e = 0; e < pl.nE; ++e) {
    if (pl.ev[e].cst_null) { tp.cst[e] = m->cst_null; tp.cst_bcast[e] = 1; }
    const int sv = pl.ev[e].sv_src;
    if (sv >= 0 && sv <= 2) tp.sv[e] = m->sv[sv];
    else if (sv == 3) { tp.sv[e] = m->null_sv; tp.sv_bcast[e] = 1; }
  }
  if (zstep) {
    TokensStepP ts;
    ts.t = tp; ts.z_rw = m->z; ts.P = m->H; ts.c_xo = m->c_xo; ts.vt = m->vt_table + (size_t)zs.t * 512;
    ts.alpha = zs.alpha; ts.beta = zs.beta; ts.first = zs.first ? 1 : 0;
    ts.sigma = zs.zeps ? zs.sigma : 0.f; ts.zeps = zs.zeps;
    ts.cfg_mode = mix->cfg_mode; ts.scale = mix->scale; ts.scale2 = mix->scale2;
    ST_TRY(tokens_step(ts, s));
  } else {
    ST_TRY(tokens_in(tp, s));
  }
  // tcgen05 engine: the 8 blocks (and the GEMM that ends a z step) for evaluations [e0, e0 + ne) on stream cs.  4 launches per
  // block.  LayerNorm never runs as a kernel: the GEMM that consumes it reads the raw residual planes and applies (mean, 1/sigma)
  // in its epilogue from the row statistics its producer left.  Evaluations are stacked along the rows, so a subset is a row range
  // of every buffer (the plane stride stays that of the whole stack).
  auto chain_layers = [&](int e0, int ne, cudaStream_t cs) -> int {
    const int Rc = ne * rows;
    const size_t r0 = (size_t)e0 * rows;
    // layer j signals counters [j][row tile]; layer j + 1 waits for the column CTAs of layer j on its own row tile
    int sig_j = 0, sig_prev_cols = 0;
    auto gemm = [&](GemmP& p, cudaStream_t st_) -> int {
      if (row_sig && sig_j + 1 < (int)kSigLayers) {
        const int cols = tc_fast_grid_x(p);
        if (cols > 0) {
          if (sig_prev_cols > 0) { p.sig_in = m->sig + (size_t)(sig_j - 1) * m->sig_rt; p.sig_expect = sig_prev_cols; }
          p.sig_out = m->sig + (size_t)sig_j * m->sig_rt;
          ++sig_j;
        }
        sig_prev_cols = cols;            // a layer outside the trunk kernel breaks the chain: its successor waits for the grid
      }
      return st::gemm(p, st_);
    };
    float* X = m->X + r0 * 512; float* P = m->H + r0 * 512; float* stats = m->ln_stats + r0 * 32;
    __half* X_p = m->X_p + r0 * 512; __half* ATT_p = m->ATT_p + r0 * 512; __half* G_p = m->G_p + r0 * 1024;
    for (int i = 0; i < 8; ++i) {
      const BlkW& b = m->blk[i];
      if (g_fused_attn) {
        // qkv GEMM + 32-token attention in one kernel (cluster of 2 CTAs per (4 sequences, head))
        GemmP pq = linear(X, Rc, 512, b.qkv_wgp, nullptr, nullptr, 1536);
        pq.a_planes = X_p; pq.a_plane_stride = ps512; pq.ln_stats = stats; pq.ln_s = b.qkv_sp; pq.ln_c = b.qkv_cp;
        pq.attn = g_attn_tc ? 2 : 1; pq.o_planes = ATT_p; pq.o_plane_stride = ps512; pq.o_planes_ld = 512;
        ST_TRY(gemm(pq, cs));
      } else {
        GemmP pq = linear(X, Rc, 512, b.qkv_wg, nullptr, m->QKV + r0 * 1536, 1536);
        pq.a_planes = X_p; pq.a_plane_stride = ps512; pq.ln_stats = stats; pq.ln_s = b.qkv_s; pq.ln_c = b.qkv_c;
        ST_TRY(gemm(pq, cs));
        ST_TRY(fast_chain_flush());
        ST_TRY(attention32(m->QKV + r0 * 1536, nullptr, ATT_p, ne * B, cs, ps512));
      }
      GemmP pp = linear(nullptr, Rc, 512, b.projw, b.projb, X, 512);
      pp.res = X; pp.res_mode = RES_POST; pp.ldr = 512; pp.a_planes = ATT_p; pp.a_plane_stride = ps512;
      pp.o_planes = X_p; pp.o_plane_stride = ps512; pp.o_planes_ld = 512; pp.stats_out = stats;
      ST_TRY(gemm(pp, cs));
      GemmP p1 = linear(X, Rc, 512, b.fc1_wg, nullptr, nullptr, 1024);
      p1.act = ACT_GELU; p1.a_planes = X_p; p1.a_plane_stride = ps512; p1.ln_stats = stats; p1.ln_s = b.fc1_s; p1.ln_c = b.fc1_c;
      p1.o_planes = G_p; p1.o_plane_stride = ps1024; p1.o_planes_ld = 1024;
      ST_TRY(gemm(p1, cs));
      if (i == 7 && fold_fc2) continue;        // x + fc2(g) is linear in (x, g): folded into the GEMM that ends the step
      GemmP p2 = linear(nullptr, Rc, 1024, b.fc2w, b.fc2b, X, 512);
      p2.res = X; p2.res_mode = RES_POST; p2.ldr = 512; p2.a_planes = G_p; p2.a_plane_stride = ps1024;
      p2.o_planes = X_p; p2.o_plane_stride = ps512; p2.o_planes_ld = 512; p2.stats_out = stats;
      ST_TRY(gemm(p2, cs));
    }
    if (zstep && !last) {
      // the next step only needs W_x x_{k-1}: P = X (W_x W_out)^T per evaluation; tokens_step mixes and applies the update
      if (fold_fc2) {
        // P = [X_mid | G] [W_xo | W_xo W_fc2]^T + W_xo b_fc2: the operand is the residual planes followed by the GELU planes
        GemmP px = linear(X, Rc, 1536, m->w_xo2, m->c_xo2, P, 512);
        px.a_planes = X_p; px.a_plane_stride = ps512; px.a2_planes = G_p; px.a2_plane_stride = ps1024; px.a2_K = 1024;
        return gemm(px, cs);
      }
      GemmP px = linear(X, Rc, 512, m->w_xo, nullptr, P, 512);
      px.a_planes = X_p; px.a_plane_stride = ps512;
      return gemm(px, cs);
    }
    GemmP po = linear(X, Rc, 512, m->out_w, m->out_b, m->O + r0 * 1536, 1536);
    po.a_planes = X_p; po.a_plane_stride = ps512;
    return gemm(po, cs);
  };
  // every trunk layer only reads rows of its own 128-row tile, so the layers leave as chains: one cluster launch per run
  auto chain_tc = [&](int e0, int ne, cudaStream_t cs) -> int {
    ST_TRY(fast_chain_begin(cs));
    profile_mark(-1);
    const int r = chain_layers(e0, ne, cs);
    const int r2 = fast_chain_end();
    profile_mark(-2);
    return r != ST_OK ? r : r2;
  };
  if (tc) {
    if (g_dual_chain && pl.nE >= 2 && loop) {
      // EXPERIMENT (st_debug_probe bit 4096): the evaluations of a step are independent until the guidance mix, so they run as two
      // chains on two streams (fork after the token prologue, join before the next one): 2 x 64 CTAs per layer side by side
      if (!m->chain_stream) {
        ST_CHECK_CUDA(cudaStreamCreateWithFlags(&m->chain_stream, cudaStreamNonBlocking));
        ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_cfork, cudaEventDisableTiming));
        ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_cjoin, cudaEventDisableTiming));
      }
      const int na = pl.nE / 2;
      ST_CHECK_CUDA(cudaEventRecord(m->ev_cfork, s));
      ST_CHECK_CUDA(cudaStreamWaitEvent(m->chain_stream, m->ev_cfork, 0));
      ST_TRY(chain_tc(na, pl.nE - na, m->chain_stream));
      ST_TRY(chain_tc(0, na, s));
      ST_CHECK_CUDA(cudaEventRecord(m->ev_cjoin, m->chain_stream));
      ST_CHECK_CUDA(cudaStreamWaitEvent(s, m->ev_cjoin, 0));
      return ST_OK;
    }
    return chain_tc(0, pl.nE, s);
  }
  for (int i = 0; i < 8; ++i) {
    const BlkW& b = m->blk[i];
    ST_TRY(layernorm512(m->X, b.ln1g, b.ln1b, m->H, nullptr, R, s));
    GemmP pq = linear(m->H, R, 512, b.qkv, nullptr, m->QKV, 1536);
    ST_TRY(gemm(pq, s));
    ST_TRY(attention32(m->QKV, m->ATT, nullptr, pl.nE * B, s));
    GemmP pp = linear(m->ATT, R, 512, b.projw, b.projb, m->X, 512);
    pp.res = m->X; pp.res_mode = RES_POST; pp.ldr = 512;
    ST_TRY(gemm(pp, s));
    ST_TRY(layernorm512(m->X, b.ln2g, b.ln2b, m->H, nullptr, R, s));
    GemmP p1 = linear(m->H, R, 512, b.fc1w, b.fc1b, m->G, 1024);
    p1.act = ACT_GELU;
    ST_TRY(gemm(p1, s));
    GemmP p2 = linear(m->G, R, 1024, b.fc2w, b.fc2b, m->X, 512);
    p2.res = m->X; p2.res_mode = RES_POST; p2.ldr = 512;
    ST_TRY(gemm(p2, s));
  }
  GemmP po = linear(m->X, R, 512, m->out_w, m->out_b, m->O, 1536);
  ST_TRY(gemm(po, s));
  return ST_OK;
}

static int upload_scales(st_model* m, const Plan& pl, const st_guidance* g, int B, StepP* sp, cudaStream_t s) {
  sp->cfg_mode = pl.cfg_mode;
  sp->nE = pl.nE;
  sp->B = B;
  sp->o = m->O;
  sp->scale = nullptr; sp->scale2 = nullptr;
  for (int k = 0; k < 3; ++k) { sp->part_sa[k] = pl.part_sa[k]; sp->part_sp[k] = pl.part_sp[k]; sp->part_ua[k] = pl.part_ua[k]; }
  // The scales come from pageable host memory, and a pageable cudaMemcpyAsync first drains the stream on the host side: in
  // steady state (same values as the last call) the device copy is reused, so the host keeps running ahead of the GPU.
  auto upload = [&](float* dst, const float* src, std::vector<float>& last) -> int {
    if ((int)last.size() == B && memcmp(last.data(), src, B * sizeof(float)) == 0) return ST_OK;
    ST_CHECK_CUDA(cudaMemcpyAsync(dst, src, B * sizeof(float), cudaMemcpyHostToDevice, s));
    last.assign(src, src + B);
    return ST_OK;
  };
  if (pl.cfg_mode == ST_CFG_TEXT || pl.cfg_mode == ST_CFG_TWO || pl.cfg_mode == ST_CFG_BODYPART1) {
    ST_TRY(upload(m->scale_dev, g->scale, m->scale_last));
    sp->scale = m->scale_dev;
    if (pl.cfg_mode == ST_CFG_TWO) {
      ST_TRY(upload(m->scale2_dev, g->scale2, m->scale2_last));
      sp->scale2 = m->scale2_dev;
    }
  }
  return ST_OK;
}

extern "C" int st_denoise(st_model* m, const float* x, const int64_t* t, const st_guidance* g, float* out, int B, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && x && t && out && B > 0, "st_denoise: null argument or B <= 0");
  if (m->cond_B != B) { set_error("st_denoise: B=%d but the conditioning cache holds B=%d (call st_cond_encode first)", B, m->cond_B); return ST_ESTATE; }
  cudaStream_t s = (cudaStream_t)stream;
  Plan pl;
  ST_TRY(make_plan(m, g, &pl));
  ST_TRY(transpose_to_tokens(x, m->xs, B, 1536, 32, 1.0f, s));
  ST_TRY(run_trunk(m, pl, B, t, 0, false, s));
  StepP sp;
  ST_TRY(upload_scales(m, pl, g, B, &sp, s));
  sp.xs = m->comb; sp.eps = nullptr; sp.mode = -1;
  ST_TRY(step_update(sp, s));
  ST_TRY(transpose_from_tokens(m->comb, out, B, 1536, 32, s));
  return ST_OK;
}

extern "C" int st_schedule_create(int S, int mode, const int32_t* t_model, const float* coef, st_schedule** out) {
  ST_REQUIRE(S > 0 && t_model && coef && out, "st_schedule_create: null argument or S <= 0");
  ST_REQUIRE(mode == ST_MODE_DDPM || mode == ST_MODE_DDIM, "st_schedule_create: unknown mode %d", mode);
  for (int k = 0; k < S; ++k) ST_REQUIRE(t_model[k] >= 0 && t_model[k] < ST_MAX_T, "t_model[%d]=%d out of [0,%d)", k, t_model[k], ST_MAX_T);
  static unsigned long long next_id = 0;
  st_schedule* sc = new st_schedule();
  sc->id = ++next_id;
  sc->S = S; sc->mode = mode;
  sc->t_model.assign(t_model, t_model + S);
  sc->coef.assign(coef, coef + (size_t)S * ST_COEF_STRIDE);
  *out = sc;
  return ST_OK;
}
extern "C" void st_schedule_destroy(st_schedule* s) { delete s; }

static std::string plan_key(const Plan& pl, int B, int mode) {
  char buf[256];
  int n = snprintf(buf, sizeof buf, "B%d m%d e%d c%d n%d", B, mode, cur_engine(), pl.cfg_mode, pl.nE);
  for (int e = 0; e < pl.nE; ++e) n += snprintf(buf + n, sizeof buf - n, " %d:%d", pl.ev[e].cst_null, pl.ev[e].sv_src);
  for (int k = 0; k < 3; ++k) n += snprintf(buf + n, sizeof buf - n, " p%g/%g/%d", pl.part_sa[k], pl.part_sp[k], pl.part_ua[k]);
  return buf;
}

// one diffusion step: trunk for every planned evaluation, CFG mix + sampler update, k -= 1  (all step-dependent
// values are read from device memory, so the same launch sequence -- or its captured graph -- serves every step)
static int one_step(st_model* m, const Plan& pl, const StepP& sp0, int B, cudaStream_t s, const ZStep& zs = ZStep(),
                    const st_schedule* sc = nullptr) {
  ST_TRY(run_trunk(m, pl, B, nullptr, 0, true, s, zs, &sp0));
  if (zs.on && !zs.last) return ST_OK;   // the update is applied to z by the next step's tokens_step
  StepP sp = sp0;
  sp.xs = m->xs; sp.eps = nullptr;
  if (zs.on) {
    // last step of the z recursion: the one real update, with step 0's coefficients as launch parameters (alpha_bar_prev = 1:
    // x <- x0_hat whatever the stale state holds); the device-side loop state is not used by this loop
    sp.ls = nullptr;
    for (int q = 0; q < ST_COEF_STRIDE; ++q) sp.c[q] = sc->coef[q];
    ST_TRY(step_update(sp, s));
    return ST_OK;
  }
  sp.ls = m->loop; sp.coef_dev = m->coef_dev;
  sp.xs_planes = cur_engine() == ST_ENGINE_TC ? m->xs_p : nullptr;
  sp.ls_advance = m->loop;          // the update's last CTA also moves the loop to the next step
  ST_TRY(step_update(sp, s));
  return ST_OK;
}

static bool schedule_has_sigma(const st_schedule* sc) {
  for (int k = 0; k < sc->S; ++k)
    if (sc->coef[(size_t)k * ST_COEF_STRIDE + (sc->mode == ST_MODE_DDPM ? 2 : 4)] != 0.f) return true;
  return false;
}
// steps per captured graph: the largest of 50, 25, 20, 10, 5, 4, 2, 1 that divides S.  Kernels of consecutive steps inside one
// graph hand over through programmatic edges like the kernels of one step; a graph launch boundary costs ~4 us.
static int chunk_steps(const st_schedule* sc) {
  for (int c : {50, 25, 20, 10, 5, 4, 2}) if (sc->S % c == 0) return c;
  return 1;
}
extern "C" int st_sample_chunk(const st_schedule* sc) { return sc ? chunk_steps(sc) : 0; }

// W_x eps for `n` steps of the caller's tape ([n,B,1536,1,32], draw order) -> dst [n][B*32][512]: transposed to token-major in
// pieces of <= ntok_steps steps, then one GEMM per piece (the noise does not depend on the chain: this runs ahead of the steps)
static int prep_zeps(st_model* m, const float* tape, int n, int B, float* dst, cudaStream_t s) {
  const size_t per = (size_t)B * 1536 * 32;
  for (int i = 0; i < n; i += m->ntok_steps) {
    const int c = (n - i) < m->ntok_steps ? (n - i) : m->ntok_steps;
    ST_TRY(transpose_to_tokens(tape + (size_t)i * per, m->ntok, c * B, 1536, 32, 1.0f, s));
    GemmP p = linear(m->ntok, c * B * 32, 1536, m->w_x, nullptr, dst + (size_t)i * B * 32 * 512, 512);
    ST_TRY(gemm(p, s));
  }
  return ST_OK;
}

// ---- the sampling loop as begin / run / end: the caller may hand the per-step noise over in chunks (the reference draws one
// randn_like per step, gaussian_diffusion.py:541,781; a 1000-step tape at B = 32 is 6.3 GB) --------------------------------
extern "C" int st_sample_begin(st_model* m, const st_schedule* sc, const st_guidance* g, const float* x_init, int B, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && sc && x_init && B > 0, "st_sample_begin: null argument or B <= 0");
  if (m->cond_B != B) { set_error("st_sample: B=%d but the conditioning cache holds B=%d (call st_cond_encode first)", B, m->cond_B); return ST_ESTATE; }
  cudaStream_t s = (cudaStream_t)stream;
  st_model::ActiveLoop& a = m->act;
  a.on = false;
  ST_TRY(make_plan(m, g, &a.pl));
  a.any_sigma = schedule_has_sigma(sc);
  ST_TRY(upload_scales(m, a.pl, g, B, &a.sp, s));
  a.sp.mode = sc->mode;
  if (m->sched_id != sc->id) {     // the schedule's tables, once per schedule
    ST_CHECK_CUDA(cudaMemcpyAsync(m->t_model_dev, sc->t_model.data(), sc->S * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ST_CHECK_CUDA(cudaMemcpyAsync(m->coef_dev, sc->coef.data(), (size_t)sc->S * ST_COEF_STRIDE * sizeof(float), cudaMemcpyHostToDevice, s));
    m->sched_id = sc->id;
  }
  ST_TRY(init_loop(m->loop, sc->S, sc->S - 1, nullptr, s));
  ST_TRY(transpose_to_tokens(x_init, m->xs, B, 1536, 32, 1.0f, s));
  if (cur_engine() == ST_ENGINE_TC) ST_TRY(tc_split(m->xs, 1536, B * 32, 1536, m->xs_p, s));
  a.G = chunk_steps(sc);
  // The tcgen05 engine keeps the loop in token space ("z recursion"): x_{k-1} = alpha x0_hat + beta x_k + sigma eps_k is linear and
  // the next step only needs W_x x_{k-1}, so between steps ONE 512 x 512 GEMM (W_x W_out, folded by the packer) replaces the
  // 512 -> 1536 output GEMM, the state update and the 1536 -> 512 input GEMM; the noise enters as W_x eps_k, formed for a whole
  // chunk of steps ahead of them; the state itself is formed once, by the last step (alpha_bar_prev = 1: x <- x0_hat, coef2 = 0,
  // sigma = 0).  The steps differ (coefficients are launch parameters), so a captured graph holds one specific chunk of the loop.
  a.zrec = g_zrec && cur_engine() == ST_ENGINE_TC && m->w_xo && a.pl.cfg_mode != ST_CFG_BODYPART && a.pl.cfg_mode != ST_CFG_BODYPART1;
  if (a.zrec) ST_TRY(trunk_input(m, B, s));
  if (a.zrec && a.any_sigma && (m->zws_B != B || m->zws_G != a.G)) {
    for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);      // captured z-step nodes point into the old block
    m->graphs.clear(); m->warmed.clear(); m->graph_nodes.clear();
    const size_t rows = (size_t)B * 32;
    int ns = (int)(32768 / rows);
    ns = ns < 1 ? 1 : (ns > a.G ? a.G : ns);
    ST_TRY(m->zws.reserve(((size_t)(a.G + 1) * rows * 512 + (size_t)ns * rows * 1536) * sizeof(float) + 8 * 256));
    m->zeps = m->zws.take<float>((size_t)(a.G > 1 ? a.G - 1 : 1) * rows * 512);
    m->zcarry[0] = m->zws.take<float>(rows * 512);
    m->zcarry[1] = m->zws.take<float>(rows * 512);
    m->ntok = m->zws.take<float>((size_t)ns * rows * 1536);
    m->ntok_steps = ns; m->zws_B = B; m->zws_G = a.G;
  }
  a.key = plan_key(a.pl, B, sc->mode) + " G" + std::to_string(a.G) + (a.zrec ? (g_zrec_fc2 ? " zf" : " z") + std::to_string(sc->id) : std::string());
  a.sc = sc; a.B = B; a.k = sc->S - 1; a.on = true;
  return ST_OK;
}

extern "C" int st_sample_run(st_model* m, int n_steps, const float* noise_chunk, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && n_steps > 0, "st_sample_run: null argument or n_steps <= 0");
  st_model::ActiveLoop& a = m->act;
  ST_REQUIRE(a.on, "st_sample_run: no sampling loop in flight (call st_sample_begin first)");
  const st_schedule* sc = a.sc;
  const int B = a.B, S = sc->S, G = a.G;
  ST_REQUIRE(n_steps <= a.k + 1, "st_sample_run: %d steps requested, %d left", n_steps, a.k + 1);
  ST_REQUIRE(!a.any_sigma || noise_chunk, "st_sample_run: the schedule has sigma != 0 but no noise was given");
  const bool znoise = a.zrec && a.any_sigma;
  ST_REQUIRE(!znoise || ((S - 1 - a.k) % G == 0 && n_steps % G == 0),
             "st_sample_run: with noise the loop advances in whole chunks of %d steps (st_sample_chunk)", G);
  cudaStream_t s = (cudaStream_t)stream;
  const bool graphs_ok = g_use_graphs && !st::profiling();
  // Capture and replay run on the model's own stream (the caller's may be the legacy default stream, which cannot capture),
  // fenced against the caller's stream with events on both sides.
  cudaStream_t ls = s;
  if (graphs_ok) {
    if (!m->loop_stream) {
      ST_CHECK_CUDA(cudaStreamCreateWithFlags(&m->loop_stream, cudaStreamNonBlocking));
      ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming));
      ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_out, cudaEventDisableTiming));
    }
    ls = m->loop_stream;
    ST_CHECK_CUDA(cudaEventRecord(m->ev_in, s));
    ST_CHECK_CUDA(cudaStreamWaitEvent(ls, m->ev_in, 0));
  }
  const size_t per = (size_t)B * 1536 * 32, zrows = (size_t)B * 32 * 512;
  int done = 0;
  while (done < n_steps) {
    const int pos = S - 1 - a.k;                       // steps already taken
    const bool aligned = (pos % G) == 0 && (n_steps - done) >= G;
    const int g = aligned ? G : 1;
    const int chunk = pos / G;
    const float* noise = noise_chunk ? noise_chunk + (size_t)done * per : nullptr;
    const float* carry_in = nullptr;
    if (znoise) {
      // W_x eps of this chunk's steps: step k_hi - j in slot j; the last step's goes to the carry the NEXT chunk's first step reads
      carry_in = m->zcarry[chunk & 1];
      if (G > 1) ST_TRY(prep_zeps(m, noise, G - 1, B, m->zeps, ls));
      ST_TRY(prep_zeps(m, noise + (size_t)(G - 1) * per, 1, B, m->zcarry[(chunk + 1) & 1], ls));
    } else if (a.any_sigma) {
      // x-space loop: step_update reads eps of step k at tape + (S - 1 - k) * per (device-side loop state)
      ST_TRY(init_loop(m->loop, S, a.k, noise - (size_t)pos * per, ls));
    }
    auto zs_of = [&](int k) {
      const int j = (S - 1 - k) - chunk * G;            // index inside the chunk
      return zstep_of(sc, k, znoise ? (j == 0 ? carry_in : m->zeps + (size_t)(j - 1) * zrows) : nullptr);
    };
    cudaGraphExec_t exec = nullptr;
    const std::string key = a.key + (a.zrec ? " c" + std::to_string(chunk) : std::string());
    if (aligned && graphs_ok && m->warmed[key] >= 1) {
      auto it = m->graphs.find(key);
      if (it != m->graphs.end()) exec = it->second;
      else {
        if (m->graphs.size() >= 96) {
          // the z recursion keys its graphs by schedule and chunk: a caller that keeps creating schedules must not grow the cache for ever
          ST_CHECK_CUDA(cudaStreamSynchronize(ls));
          for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);
          m->graphs.clear(); m->graph_nodes.clear();
        }
        cudaGraph_t graph = nullptr;
        const int64_t l0 = g_launches;
        ST_CHECK_CUDA(cudaStreamBeginCapture(ls, cudaStreamCaptureModeRelaxed));
        int r = ST_OK;
        for (int i = 0; i < G && r == ST_OK; ++i) r = a.zrec ? one_step(m, a.pl, a.sp, B, ls, zs_of(a.k - i), sc) : one_step(m, a.pl, a.sp, B, ls);
        m->graph_nodes[key] = g_launches - l0;     // kernels per replay; capturing itself executed nothing
        g_launches = l0;
        cudaError_t ce = cudaStreamEndCapture(ls, &graph);
        if (r != ST_OK) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return r; }
        ST_CHECK_CUDA(ce);
        ST_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        m->graphs[key] = exec;
      }
    }
    if (exec) {
      ST_CHECK_CUDA(cudaGraphLaunch(exec, ls));
      g_launches += m->graph_nodes[key];
    } else {
      for (int i = 0; i < g; ++i) ST_TRY(a.zrec ? one_step(m, a.pl, a.sp, B, ls, zs_of(a.k - i), sc) : one_step(m, a.pl, a.sp, B, ls));
    }
    if (aligned) m->warmed[key] += 1;
    a.k -= g;
    done += g;
  }
  if (ls != s) {
    ST_CHECK_CUDA(cudaEventRecord(m->ev_out, ls));
    ST_CHECK_CUDA(cudaStreamWaitEvent(s, m->ev_out, 0));
  }
  return ST_OK;
}

extern "C" int st_sample_end(st_model* m, float* x_out, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && x_out, "st_sample_end: null argument");
  st_model::ActiveLoop& a = m->act;
  ST_REQUIRE(a.on, "st_sample_end: no sampling loop in flight");
  ST_REQUIRE(a.k < 0, "st_sample_end: %d steps of the loop have not run", a.k + 1);
  a.on = false;
  return transpose_from_tokens(m->xs, x_out, a.B, 1536, 32, (cudaStream_t)stream);
}

// the whole loop in one call; noise_tape = [S,B,1536,1,32] in draw order (t = S-1 .. 0) or NULL when every sigma is 0
extern "C" int st_sample(st_model* m, const st_schedule* sc, const st_guidance* g, const float* x_init, const float* noise_tape,
                         int B, float* x_out, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && sc && x_init && x_out && B > 0, "st_sample: null argument or B <= 0");
  ST_REQUIRE(!schedule_has_sigma(sc) || noise_tape, "st_sample: schedule has sigma != 0 but noise_tape is NULL");
  ST_TRY(st_sample_begin(m, sc, g, x_init, B, stream));
  ST_TRY(st_sample_run(m, sc->S, noise_tape, stream));
  return st_sample_end(m, x_out, stream);
}

// ---------------------------------------------------------------------------------------------------------
static int vq_resolve(st_vq* v) {
  int err = ST_OK;
  char nm[96];
  auto conv = [&](const char* name, int cout, int K) {
    ConvW c;
    c.ldw = (K + 3) & ~3;
    snprintf(nm, sizeof nm, "dec.%s.w", name);
    c.w = v->w.get(nm, (int64_t)cout * c.ldw, &err);
    snprintf(nm, sizeof nm, "dec.%s.b", name);
    c.b = v->w.get(nm, cout, &err);
    return c;
  };
  for (int q = 0; q < 6; ++q) {
    snprintf(nm, sizeof nm, "cb.%d", q);
    v->cb[q] = v->w.get(nm, 512 * 512, &err);
    snprintf(nm, sizeof nm, "cnorm.%d", q);
    v->cnorm[q] = v->w.get(nm, 512, &err);
  }
  v->c0 = conv("0", 512, 1536);
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 3; ++j) {
      char n2[32];
      snprintf(n2, sizeof n2, "%d.0.%d.conv1", i + 2, j);
      v->res1[i][j] = conv(n2, 512, 1536);
      snprintf(n2, sizeof n2, "%d.0.%d.conv2", i + 2, j);
      v->res2[i][j] = conv(n2, 512, 512);
    }
    char n3[32];
    snprintf(n3, sizeof n3, "%d.2", i + 2);
    v->up[i] = conv(n3, 512, 1536);
    snprintf(n3, sizeof n3, "%d.2.even", i + 2);
    v->up_eo[i][0] = conv(n3, 512, 1024);
    snprintf(n3, sizeof n3, "%d.2.odd", i + 2);
    v->up_eo[i][1] = conv(n3, 512, 1024);
  }
  v->c4 = conv("4", 512, 1536);
  v->c6 = conv("6", v->out_dim, 1536);
  v->has_enc = v->w.dev.count("enc.0.w") != 0;
  if (v->has_enc) {
    auto econv = [&](const char* name, int cout, int K) {
      ConvW c;
      c.ldw = (K + 3) & ~3;
      snprintf(nm, sizeof nm, "enc.%s.w", name);
      c.w = v->w.get(nm, (int64_t)cout * c.ldw, &err);
      snprintf(nm, sizeof nm, "enc.%s.b", name);
      c.b = v->w.get(nm, cout, &err);
      return c;
    };
    v->e0 = econv("0", 512, 3 * v->out_dim);
    for (int i = 0; i < 2; ++i) {
      char n2[32];
      snprintf(n2, sizeof n2, "%d.0", i + 2);
      v->edown[i] = econv(n2, 512, 2048);
      for (int j = 0; j < 3; ++j) {
        snprintf(n2, sizeof n2, "%d.1.%d.conv1", i + 2, j);
        v->eres1[i][j] = econv(n2, 512, 1536);
        snprintf(n2, sizeof n2, "%d.1.%d.conv2", i + 2, j);
        v->eres2[i][j] = econv(n2, 512, 512);
      }
    }
    v->e4 = econv("4", 512, 1536);
  }
  return err;
}

extern "C" int st_vq_create(const st_tensor* packed, int n, int out_dim, st_vq** out) {
  ST_REQUIRE(packed && out && n > 0 && out_dim > 0, "st_vq_create: null argument");
  st_vq* v = new st_vq();
  v->out_dim = out_dim;
  int r = v->w.upload(packed, n);
  if (r == ST_OK) r = vq_resolve(v);
  if (r != ST_OK) { v->w.release(); delete v; return r; }
  *out = v;
  return ST_OK;
}
extern "C" void st_vq_destroy(st_vq* v) {
  if (!v) return;
  cudaDeviceSynchronize();
  v->w.release(); v->ws.release();
  delete v;
}
extern "C" int st_vq_out_dim(const st_vq* v) { return v ? v->out_dim : 0; }

static GemmP conv3(const float* in, const ConvW& w, float* out, int B, int Lin, int Lout, int N, int dil, int ups) {
  GemmP p;
  p.A = in; p.W = w.w; p.bias = w.b; p.out = out;
  p.M = B * Lout; p.N = N; p.K = 1536; p.ldw = w.ldw;
  p.Lout = Lout; p.Lin = Lin; p.C = 512; p.stride = 1; p.pad = dil; p.dil = dil; p.ups = ups;
  p.a_batch = (long long)Lin * 512; p.lda = 512; p.ldo = N;
  return p;
}

// Decoder conv stack on the tcgen05 engine: every conv streams its input as fp16 hi/lo planes written by the epilogue of its
// producer (ReLU applied there when the consumer is `ReLU -> Conv`, resnet.py:48-69), so the only split pass is the one over
// the quantised latent.  The fp32 copy of an activation is kept where it is a residual or the result.
//   c0: relu(conv(x)) = h                       res block: h += conv1x1(relu(conv_k3_dil(relu(h))))
//   x2 upsample + k3 conv = two 2-tap convs on the low-resolution rows (even / odd output rows, weights pre-summed by the
//   packer): out[2u] = W0 in[u-1] + (W1+W2) in[u];  out[2u+1] = (W0+W1) in[u] + W2 in[u+1]
static int decode_convs_tc(st_vq* v, const float* qsum, const __half* qsum_planes, int B, int T4, float* rec, float* hA, float* hB, float* hC,
                           cudaStream_t s) {
  const size_t big = (size_t)B * T4 * 4 * 512;                 // elements of the largest activation [B, 4*T4, 512]
  __half* P[3];
  for (int k = 0; k < 3; ++k) P[k] = v->ws.take<__half>(2 * big);
  int T = T4;
  long long ps = (long long)B * T * 512;                        // plane stride of the current resolution
  float* h = hA; float* nxt = hC;
  __half* Ph = P[0]; __half* Pt = P[1]; __half* Pn = P[2];
  GemmP p0 = conv3(qsum, v->c0, h, B, T, T, 512, 1, 0);
  if (qsum_planes) { p0.a_planes = qsum_planes; p0.a_plane_stride = ps; }
  p0.act = ACT_RELU; p0.o_planes = Ph; p0.o_plane_stride = ps; p0.o_planes_ld = 512;
  ST_TRY(gemm(p0, s));
  const int dils[3] = {9, 3, 1};
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 3; ++j) {
      GemmP p1 = conv3(nullptr, v->res1[i][j], nullptr, B, T, T, 512, dils[j], 0);
      p1.a_planes = Ph; p1.a_plane_stride = ps;                                   // relu(h)
      p1.o_planes = Pt; p1.o_plane_stride = ps; p1.o_planes_ld = 512; p1.o_planes_relu = 1;
      ST_TRY(gemm(p1, s));
      GemmP p2 = linear(nullptr, B * T, 512, v->res2[i][j].w, v->res2[i][j].b, h, 512);
      p2.a_planes = Pt; p2.a_plane_stride = ps;
      p2.res = h; p2.res_mode = RES_POST; p2.ldr = 512;
      p2.o_planes = Ph; p2.o_plane_stride = ps; p2.o_planes_ld = 512; p2.o_planes_relu = j < 2;   // the upsampling conv takes h itself
      ST_TRY(gemm(p2, s));
    }
    const long long ps2 = (long long)B * 2 * T * 512;
    for (int par = 0; par < 2; ++par) {
      GemmP pu;
      pu.W = v->up_eo[i][par].w; pu.bias = v->up_eo[i][par].b; pu.out = nxt + par * 512;
      pu.M = B * T; pu.N = 512; pu.K = 1024; pu.ldw = 1024;
      pu.Lout = T; pu.Lin = T; pu.C = 512; pu.stride = 1; pu.pad = par == 0 ? 1 : 0; pu.dil = 1;
      pu.a_batch = (long long)T * 512; pu.lda = 512; pu.ldo = 1024;
      pu.a_planes = Ph; pu.a_plane_stride = ps;
      pu.o_planes = Pn + par * 512; pu.o_plane_stride = ps2; pu.o_planes_ld = 1024; pu.o_planes_relu = i == 0;   // next: res block / conv 4
      ST_TRY(gemm(pu, s));
    }
    T *= 2; ps = ps2;
    { float* o = h; h = nxt; nxt = o; }
    { __half* o = Ph; Ph = Pn; Pn = o; }
  }
  GemmP p4 = conv3(nullptr, v->c4, nullptr, B, T, T, 512, 1, 0);
  p4.act = ACT_RELU; p4.a_planes = Ph; p4.a_plane_stride = ps;
  p4.o_planes = Pt; p4.o_plane_stride = ps; p4.o_planes_ld = 512;
  ST_TRY(gemm(p4, s));
  GemmP p6 = conv3(nullptr, v->c6, rec, B, T, T, v->out_dim, 1, 0);
  p6.a_planes = Pt; p6.a_plane_stride = ps;
  ST_TRY(gemm(p6, s));
  (void)hB;
  return ST_OK;
}

extern "C" int st_rvq_decode(st_vq* v, const float* lat, int64_t lat_stride, float lat_scale, int B, int T4, float* rec,
                             int64_t* idx_out, float* residual_out, void* stream) {
  EngineScope es(v ? v->engine : -1);
  ST_REQUIRE(v && lat && rec && B > 0 && T4 > 0, "st_rvq_decode: null argument or empty batch");
  ST_REQUIRE(lat_stride >= 512 && (lat_stride & 3) == 0, "st_rvq_decode: lat_stride=%lld", (long long)lat_stride);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t rows = (size_t)B * T4;
  const size_t big = rows * 4 * 512;
  ST_TRY(v->ws.reserve((rows * 512 * 3 + big * 3) * sizeof(float) + (3 * 2 * big + 4 * rows * 512) * sizeof(__half) + 40 * 256));
  float* r = v->ws.take<float>(rows * 512);
  float* dot = v->ws.take<float>(rows * 512);
  float* qsum = v->ws.take<float>(rows * 512);
  float* hA = v->ws.take<float>(big);
  float* hB = v->ws.take<float>(big);
  float* hC = v->ws.take<float>(big);
  ST_TRY(copy_strided_scale(lat, lat_stride, lat_scale, r, (int)rows, 512, s));
  // residual quantisation, 6 layers (residual_vq.py:132-152).  tcgen05 engine: the residual travels as planes owned by this
  // handle (vq_select writes the next layer's operand, and the quantised sum for the decoder's first conv), so nothing
  // here touches the engine's shared split scratch and the three body parts may run on different streams.
  const bool tc = cur_engine() == ST_ENGINE_TC && !g_rank_simt;
  __half* rp = tc ? v->ws.take<__half>(2 * rows * 512) : nullptr;
  __half* qp = tc ? v->ws.take<__half>(2 * rows * 512) : nullptr;
  if (tc) ST_TRY(tc_split(r, 512, (int)rows, 512, rp, s));
  for (int q = 0; q < 6; ++q) {
    GemmP p = linear(r, (int)rows, 512, v->cb[q], nullptr, dot, 512);
    if (tc) { p.a_planes = rp; p.a_plane_stride = (long long)rows * 512; }
    ST_TRY(g_rank_simt ? gemm_simt(p, s) : gemm(p, s));   // split-fp16 products carry fp32-class error; bit 64 of st_debug_probe forces SIMT
    ST_TRY(vq_select(dot, v->cnorm[q], v->cb[q], r, qsum, idx_out ? idx_out + q : nullptr, 6, (int)rows, q == 0, q < 5 ? rp : nullptr,
                     q == 5 ? qp : nullptr, s));
  }
  if (residual_out) ST_CHECK_CUDA(cudaMemcpyAsync(residual_out, r, rows * 512 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  // decoder (encdec.py:51-68)
  if (cur_engine() == ST_ENGINE_TC) return decode_convs_tc(v, qsum, qp, B, T4, rec, hA, hB, hC, s);
  int T = T4;
  GemmP p0 = conv3(qsum, v->c0, hA, B, T, T, 512, 1, 0);
  p0.act = ACT_RELU;
  ST_TRY(gemm(p0, s));
  float* h = hA; float* tmp = hB; float* nxt = hC;
  const int dils[3] = {9, 3, 1};
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 3; ++j) {
      GemmP p1 = conv3(h, v->res1[i][j], tmp, B, T, T, 512, dils[j], 0);
      p1.a_relu = 1;
      ST_TRY(gemm(p1, s));
      GemmP p2 = linear(tmp, B * T, 512, v->res2[i][j].w, v->res2[i][j].b, h, 512);
      p2.a_relu = 1; p2.res = h; p2.res_mode = RES_POST; p2.ldr = 512;
      ST_TRY(gemm(p2, s));
    }
    GemmP pu = conv3(h, v->up[i], nxt, B, T, 2 * T, 512, 1, 1);
    ST_TRY(gemm(pu, s));
    T *= 2;
    float* o = h; h = nxt; nxt = o;
  }
  GemmP p4 = conv3(h, v->c4, tmp, B, T, T, 512, 1, 0);
  p4.act = ACT_RELU;
  ST_TRY(gemm(p4, s));
  GemmP p6 = conv3(tmp, v->c6, rec, B, T, T, v->out_dim, 1, 0);
  ST_TRY(gemm(p6, s));
  return ST_OK;
}


// ---- RVQ-VAE encode: pose features -> latent before quantisation (models/vq/model.py:95-100, encdec.py:5-34) ----
// Conv k3 + ReLU; 2 x (Conv k4 stride 2 pad 1; 3 res blocks h += conv1x1(relu(conv_k3_dil(relu(h)))), dilations 9,3,1); Conv k3.
// Channels-last implicit GEMMs through the engine dispatch: the D-channel input conv and the two strided convs run on the
// exact-fp32 engine (C_in not a multiple of 64 / padded stride), the res blocks and the last conv on tcgen05.
extern "C" int st_rvq_encode(st_vq* v, const float* pose, int B, int T, float* lat, void* stream) {
  EngineScope es(v ? v->engine : -1);
  ST_REQUIRE(v && pose && lat && B > 0 && T > 0 && (T % 4) == 0, "st_rvq_encode: null argument or T not a multiple of 4");
  ST_REQUIRE(v->has_enc, "st_rvq_encode: the checkpoint this handle was created from has no encoder weights");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t big = (size_t)B * T * 512;
  ST_TRY(v->ws.reserve(big * 3 * sizeof(float) + 16 * 256));
  float* h = v->ws.take<float>(big);
  float* tmp = v->ws.take<float>(big);
  float* nxt = v->ws.take<float>(big);
  const int D = v->out_dim;
  GemmP p0;
  p0.A = pose; p0.W = v->e0.w; p0.bias = v->e0.b; p0.out = h;
  p0.M = B * T; p0.N = 512; p0.K = 3 * D; p0.ldw = v->e0.ldw;
  p0.Lout = T; p0.Lin = T; p0.C = D; p0.stride = 1; p0.pad = 1; p0.dil = 1;
  p0.a_batch = (long long)T * D; p0.lda = D; p0.ldo = 512; p0.act = ACT_RELU;
  ST_TRY(gemm(p0, s));
  int L = T;
  const int dils[3] = {9, 3, 1};
  for (int i = 0; i < 2; ++i) {
    GemmP pd;
    pd.A = h; pd.W = v->edown[i].w; pd.bias = v->edown[i].b; pd.out = nxt;
    pd.M = B * (L / 2); pd.N = 512; pd.K = 2048; pd.ldw = v->edown[i].ldw;
    pd.Lout = L / 2; pd.Lin = L; pd.C = 512; pd.stride = 2; pd.pad = 1; pd.dil = 1;
    pd.a_batch = (long long)L * 512; pd.lda = 512; pd.ldo = 512;
    ST_TRY(gemm(pd, s));
    L /= 2;
    { float* o = h; h = nxt; nxt = o; }
    for (int j = 0; j < 3; ++j) {
      GemmP p1 = conv3(h, v->eres1[i][j], tmp, B, L, L, 512, dils[j], 0);
      p1.a_relu = 1;
      ST_TRY(gemm(p1, s));
      GemmP p2 = linear(tmp, B * L, 512, v->eres2[i][j].w, v->eres2[i][j].b, h, 512);
      p2.a_relu = 1; p2.res = h; p2.res_mode = RES_POST; p2.ldr = 512;
      ST_TRY(gemm(p2, s));
    }
  }
  GemmP p4 = conv3(h, v->e4, lat, B, L, L, 512, 1, 0);
  ST_TRY(gemm(p4, s));
  return ST_OK;
}

extern "C" int st_pose_assemble_330(const float* rec_upper, const float* rec_hands, const float* rec_lower, const float* mean,
                                    const float* std, const float* trans_mean, const float* trans_std, const float* jaw_aa, int B,
                                    int n, float* rec_pose, float* rec_trans, void* stream) {
  ST_REQUIRE(rec_upper && rec_hands && rec_lower && mean && std && rec_pose && B > 0 && n > 0, "st_pose_assemble_330: null argument");
  ST_REQUIRE(!rec_trans || (trans_mean && trans_std), "st_pose_assemble_330: rec_trans needs trans_mean/std");
  return pose330(rec_upper, rec_hands, rec_lower, mean, std, trans_mean, trans_std, jaw_aa, B, n, rec_pose, rec_trans, (cudaStream_t)stream);
}
extern "C" int st_pose_assemble_623(const float* rec_upper, const float* rec_hands, const float* rec_lower, int B, int n,
                                    float* rec_pose, void* stream) {
  ST_REQUIRE(rec_upper && rec_hands && rec_lower && rec_pose && B > 0 && n > 0, "st_pose_assemble_623: null argument");
  return pose623(rec_upper, rec_hands, rec_lower, B, n, rec_pose, (cudaStream_t)stream);
}
// ---- evaluation tail: result-file pose vector and the rank-reducible metric statistics --------------------
extern "C" int st_pose_330_to_aa165(const float* rec_pose, int64_t frames, float* poses_aa, void* stream) {
  ST_REQUIRE(rec_pose && poses_aa && frames > 0, "st_pose_330_to_aa165: null argument");
  return pose_aa165(rec_pose, frames, poses_aa, (cudaStream_t)stream);
}
extern "C" int st_moments_accumulate(const float* x, int64_t N, int D, double* acc, void* stream) {
  ST_REQUIRE(x && acc && N > 0 && D > 0 && D <= 4096, "st_moments_accumulate: null argument or bad shape");
  return moments_accumulate(x, N, D, acc, (cudaStream_t)stream);
}
extern "C" int st_l1div_accumulate(const float* x, int n, int J, double* acc, void* stream) {
  ST_REQUIRE(x && acc && n > 0 && J > 0, "st_l1div_accumulate: null argument or bad shape");
  return l1div_accumulate(x, n, J, acc, (cudaStream_t)stream);
}

extern "C" int st_sample_to_tokens(const float* sample, int B, int T, float scale, float* tokens, void* stream) {
  ST_REQUIRE(sample && tokens && B > 0 && T > 0, "st_sample_to_tokens: null argument");
  return transpose_to_tokens(sample, tokens, B, 1536, T, scale, (cudaStream_t)stream);
}


// latent2origin of the three body parts (trainer:480-482).  The parts are independent and each launch covers only part of
// the machine at window-batch sizes (64-256 CTAs), so they run on three streams: fork after `s`, join back into `s`.
static int decode_three(st_model* m, st_vq* const vq[3], const float* tok, float latent_scale, int B, int T4, float* const rec[3], cudaStream_t s) {
  const bool side = g_decode_streams && vq[0] != vq[1] && vq[1] != vq[2] && vq[0] != vq[2];
  if (!side) {
    for (int k = 0; k < 3; ++k) ST_TRY(st_rvq_decode(vq[k], tok + 512 * k, 1536, latent_scale, B, T4, rec[k], nullptr, nullptr, s));
    return ST_OK;
  }
  if (!m->dec_stream[0]) {
    for (int k = 0; k < 2; ++k) {
      ST_CHECK_CUDA(cudaStreamCreateWithFlags(&m->dec_stream[k], cudaStreamNonBlocking));
      ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_join[k], cudaEventDisableTiming));
    }
    ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  }
  ST_CHECK_CUDA(cudaEventRecord(m->ev_fork, s));
  for (int k = 0; k < 2; ++k) ST_CHECK_CUDA(cudaStreamWaitEvent(m->dec_stream[k], m->ev_fork, 0));
  ST_TRY(st_rvq_decode(vq[1], tok + 512, 1536, latent_scale, B, T4, rec[1], nullptr, nullptr, m->dec_stream[0]));   // hands: the widest decoder first
  ST_TRY(st_rvq_decode(vq[0], tok, 1536, latent_scale, B, T4, rec[0], nullptr, nullptr, s));
  ST_TRY(st_rvq_decode(vq[2], tok + 1024, 1536, latent_scale, B, T4, rec[2], nullptr, nullptr, m->dec_stream[1]));
  for (int k = 0; k < 2; ++k) {
    ST_CHECK_CUDA(cudaEventRecord(m->ev_join[k], m->dec_stream[k]));
    ST_CHECK_CUDA(cudaStreamWaitEvent(s, m->ev_join[k], 0));
  }
  return ST_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int st_generate_330(st_model* m, const st_schedule* sc, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                               st_vq* vq_lower, const st_cond* cond, const float* x_init, const float* noise_tape,
                               const float* jaw_aa, const float* ms, int B, float latent_scale, float* rec_pose, float* rec_trans,
                               float* sample_out, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && sc && vq_upper && vq_hands && vq_lower && cond && x_init && ms && rec_pose && B > 0, "st_generate_330: null argument");
  ST_REQUIRE(vq_upper->out_dim == 78 && vq_hands->out_dim == 180 && vq_lower->out_dim == 57, "st_generate_330: decoders must be 78/180/57 wide");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nx = (size_t)B * ST_LATENT * ST_TOKENS;
  ST_TRY(m->io.reserve((nx * 2 + (size_t)B * 128 * (78 + 180 + 57)) * sizeof(float) + 16 * 256));
  float* d_x = m->io.take<float>(nx);
  float* d_tok = m->io.take<float>(nx);
  float* d_up = m->io.take<float>((size_t)B * 128 * 78);
  float* d_ha = m->io.take<float>((size_t)B * 128 * 180);
  float* d_lo = m->io.take<float>((size_t)B * 128 * 57);
  ST_TRY(st_cond_encode(m, cond, B, stream));
  float* xo = sample_out ? sample_out : d_x;
  ST_TRY(st_sample(m, sc, g, x_init, noise_tape, B, xo, stream));
  ST_TRY(transpose_to_tokens(xo, d_tok, B, 1536, 32, 1.0f, s));
  st_vq* const vq3[3] = {vq_upper, vq_hands, vq_lower};
  float* const rec3[3] = {d_up, d_ha, d_lo};
  ST_TRY(decode_three(m, vq3, d_tok, latent_scale, B, 32, rec3, s));
  ST_TRY(pose330(d_up, d_ha, d_lo, ms, ms + 330, ms + 660, ms + 663, jaw_aa, B, 128, rec_pose, rec_trans, s));
  return ST_OK;
}

// Host-buffer pipeline.  `slot` (0 / 1) names one of two staging sets, so that two window batches can be in flight: the
// inputs of batch i + 1 cross PCIe on the H2D stream and the results of batch i return on the D2H stream while the
// compute stream works (the computations themselves are serialised on `stream`: they share the model's workspace).
// st_generate_330_host_begin enqueues everything and returns; st_generate_330_host_wait blocks until the slot's results
// are in the caller's host buffers.  A slot must be waited for before it is reused.
extern "C" int st_generate_330_host_begin(st_model* m, const st_schedule* sc, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                                          st_vq* vq_lower, const st_host_inputs* in, int B, float latent_scale, float* rec_pose_host,
                                          float* rec_trans_host, float* sample_host, int slot, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && sc && in && rec_pose_host && B > 0, "st_generate_330_host: null argument");
  ST_REQUIRE(slot == 0 || slot == 1, "st_generate_330_host: slot must be 0 or 1");
  ST_REQUIRE(in->audio && in->word && in->seed && in->x_init && in->mean && in->std, "st_generate_330_host: missing host input");
  st_model::HostSlot& hs = m->hslot[slot];
  if (hs.pending) { set_error("st_generate_330_host: slot %d is still in flight (call st_generate_330_host_wait first)", slot); return ST_ESTATE; }
  cudaStream_t s = (cudaStream_t)stream;
  if (!m->h2d_stream) {
    ST_CHECK_CUDA(cudaStreamCreateWithFlags(&m->h2d_stream, cudaStreamNonBlocking));
    ST_CHECK_CUDA(cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking));
    ST_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_enter, cudaEventDisableTiming));
  }
  if (!hs.ev_h2d) {
    ST_CHECK_CUDA(cudaEventCreateWithFlags(&hs.ev_h2d, cudaEventDisableTiming));
    ST_CHECK_CUDA(cudaEventCreateWithFlags(&hs.ev_comp, cudaEventDisableTiming));
    ST_CHECK_CUDA(cudaEventCreateWithFlags(&hs.ev_done, cudaEventDisableTiming));
  }
  const size_t nx = (size_t)B * ST_LATENT * ST_TOKENS;
  const int sdim = m->style_dim;
  bool any_sigma = false;
  for (int k = 0; k < sc->S; ++k) any_sigma |= sc->coef[(size_t)k * ST_COEF_STRIDE + (sc->mode == ST_MODE_DDPM ? 2 : 4)] != 0.f;
  ST_REQUIRE(!any_sigma || in->noise_tape, "st_generate_330_host: schedule has sigma != 0 but noise_tape is NULL");
  // staging lives in the slot's own arena (m->io is used by st_generate_330)
  size_t f = (size_t)B * ST_AUDIO_LEN * 2 + (size_t)B * 128 + (size_t)B * 6144 + 3 * (size_t)B * (sdim ? sdim : 1) + nx * 2 +
             (any_sigma ? nx * sc->S : 0) + (size_t)B * 128 * (3 + 330 + 3) + 666 + 64 * 16;
  ST_TRY(hs.stage.reserve(f * sizeof(float) + 32 * 256));
  Arena& a = hs.stage;
  float* d_audio = a.take<float>((size_t)B * ST_AUDIO_LEN * 2);
  int32_t* d_word = a.take<int32_t>((size_t)B * 128);
  float* d_seed = a.take<float>((size_t)B * 6144);
  float* d_style[3] = {nullptr, nullptr, nullptr};
  float* d_x = a.take<float>(nx);
  float* d_s = a.take<float>(nx);
  float* d_tape = any_sigma ? a.take<float>(nx * sc->S) : nullptr;
  float* d_jaw = in->jaw_aa ? a.take<float>((size_t)B * 128 * 3) : nullptr;
  float* d_pose = a.take<float>((size_t)B * 128 * 330);
  float* d_trans = a.take<float>((size_t)B * 128 * 3);
  float* d_ms = a.take<float>(666);
  // ---- inputs: H2D stream.  Not ordered behind `stream` (that is the point: the previous batch is still computing there);
  // the slot's staging set is free because its previous use was waited for ----
  cudaStream_t hi = m->h2d_stream, ho = m->d2h_stream;
  auto h2d = [&](void* d, const void* h, size_t bytes) { return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, hi); };
  ST_CHECK_CUDA(h2d(d_audio, in->audio, (size_t)B * ST_AUDIO_LEN * 2 * sizeof(float)));
  ST_CHECK_CUDA(h2d(d_word, in->word, (size_t)B * 128 * sizeof(int32_t)));
  ST_CHECK_CUDA(h2d(d_seed, in->seed, (size_t)B * 6144 * sizeof(float)));
  for (int k = 0; k < 3; ++k)
    if (sdim && in->style[k]) {
      d_style[k] = a.take<float>((size_t)B * sdim);
      ST_CHECK_CUDA(h2d(d_style[k], in->style[k], (size_t)B * sdim * sizeof(float)));
    }
  ST_CHECK_CUDA(h2d(d_x, in->x_init, nx * sizeof(float)));
  if (d_tape) ST_CHECK_CUDA(h2d(d_tape, in->noise_tape, nx * sc->S * sizeof(float)));
  if (d_jaw) ST_CHECK_CUDA(h2d(d_jaw, in->jaw_aa, (size_t)B * 128 * 3 * sizeof(float)));
  ST_CHECK_CUDA(cudaMemsetAsync(d_ms, 0, 666 * sizeof(float), hi));
  ST_CHECK_CUDA(h2d(d_ms, in->mean, 330 * sizeof(float)));
  ST_CHECK_CUDA(h2d(d_ms + 330, in->std, 330 * sizeof(float)));
  const bool want_trans = rec_trans_host && in->trans_mean && in->trans_std;
  if (want_trans) {
    ST_CHECK_CUDA(h2d(d_ms + 660, in->trans_mean, 3 * sizeof(float)));
    ST_CHECK_CUDA(h2d(d_ms + 663, in->trans_std, 3 * sizeof(float)));
  }
  ST_CHECK_CUDA(cudaEventRecord(hs.ev_h2d, hi));
  // ---- compute: the caller's stream ----
  ST_CHECK_CUDA(cudaStreamWaitEvent(s, hs.ev_h2d, 0));
  st_cond c;
  c.audio = d_audio; c.word = d_word; c.seed = d_seed;
  for (int k = 0; k < 3; ++k) c.style[k] = d_style[k];
  ST_TRY(st_generate_330(m, sc, g, vq_upper, vq_hands, vq_lower, &c, d_x, d_tape, d_jaw, d_ms, B, latent_scale, d_pose,
                         want_trans ? d_trans : nullptr, d_s, stream));
  ST_CHECK_CUDA(cudaEventRecord(hs.ev_comp, s));
  // ---- results: D2H stream ----
  ST_CHECK_CUDA(cudaStreamWaitEvent(ho, hs.ev_comp, 0));
  if (sample_host) ST_CHECK_CUDA(cudaMemcpyAsync(sample_host, d_s, nx * sizeof(float), cudaMemcpyDeviceToHost, ho));
  ST_CHECK_CUDA(cudaMemcpyAsync(rec_pose_host, d_pose, (size_t)B * 128 * 330 * sizeof(float), cudaMemcpyDeviceToHost, ho));
  if (want_trans) ST_CHECK_CUDA(cudaMemcpyAsync(rec_trans_host, d_trans, (size_t)B * 128 * 3 * sizeof(float), cudaMemcpyDeviceToHost, ho));
  ST_CHECK_CUDA(cudaEventRecord(hs.ev_done, ho));
  hs.pending = true;
  return ST_OK;
}

extern "C" int st_generate_330_host_wait(st_model* m, int slot) {
  ST_REQUIRE(m && (slot == 0 || slot == 1), "st_generate_330_host_wait: bad argument");
  st_model::HostSlot& hs = m->hslot[slot];
  if (!hs.pending) return ST_OK;
  hs.pending = false;
  ST_CHECK_CUDA(cudaEventSynchronize(hs.ev_done));
  return ST_OK;
}

// one window batch, blocking: begin on slot 0 + wait
extern "C" int st_generate_330_host(st_model* m, const st_schedule* sc, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                                    st_vq* vq_lower, const st_host_inputs* in, int B, float latent_scale, float* rec_pose_host,
                                    float* rec_trans_host, float* sample_host, void* stream) {
  ST_TRY(st_generate_330_host_wait(m, 0));
  ST_TRY(st_generate_330_host_begin(m, sc, g, vq_upper, vq_hands, vq_lower, in, B, latent_scale, rec_pose_host, rec_trans_host, sample_host, 0, stream));
  return st_generate_330_host_wait(m, 0);
}

// ---- long clip: the trainer's window loop on the device (diffusion_rvqvae_trainer.py:413-531) ------------------------
// R windows of 128 frames, consecutive windows share pre_frames * 4 = 16 frames: window i reads words
// [112 i, 112 i + 128) and audio samples [533 * 112 i, + 68224); its seed is the GT-derived seed0 for i = 0 and the last 4
// latent tokens of window i - 1's sample afterwards (trainer:431); window 0 contributes all 32 tokens, later windows
// their last 28 (trainer:462-469).  The concatenated latents are decoded ONCE per body part over all 32 + 28 (R - 1)
// tokens and assembled into 330-d features.  Stream-ordered, no host synchronisation, no return to the caller between
// windows.  B long clips run side by side (the reference's glue is bs = 1; rows are independent).
extern "C" int st_generate_long_330(st_model* m, const st_schedule* sc, const st_guidance* g, st_vq* vq_upper, st_vq* vq_hands,
                                    st_vq* vq_lower, const float* audio_long, int64_t audio_len, const int32_t* word_long,
                                    int64_t n_words, const float* seed0, const float* const* style, const float* x_init,
                                    const float* noise_tape, const float* jaw_aa, const float* ms, int B, int R,
                                    float latent_scale, float* rec_pose, float* rec_trans, float* latents_out, void* stream) {
  EngineScope es(m ? m->engine : -1);
  ST_REQUIRE(m && sc && vq_upper && vq_hands && vq_lower && audio_long && word_long && seed0 && x_init && ms && rec_pose && B > 0 && R > 0,
             "st_generate_long_330: null argument");
  ST_REQUIRE(vq_upper->out_dim == 78 && vq_hands->out_dim == 180 && vq_lower->out_dim == 57, "st_generate_long_330: decoders must be 78/180/57 wide");
  const int round_l = 128 - 16;                                    // pose_length - pre_frames * vqvae_squeeze_scale
  const int64_t hop_a = (int64_t)(16000 / 30) * round_l;           // audio samples between windows (trainer:422)
  ST_REQUIRE(n_words >= (int64_t)round_l * R + 16 && audio_len >= hop_a * R + (16000 / 30) * 16,
             "st_generate_long_330: %d windows need %lld words and %lld audio samples", R, (long long)round_l * R + 16,
             (long long)(hop_a * R + (16000 / 30) * 16));
  cudaStream_t s = (cudaStream_t)stream;
  const int Ttot = 32 + 28 * (R - 1), n = 4 * Ttot;
  const size_t nx = (size_t)B * ST_LATENT * ST_TOKENS;
  ST_TRY(m->longws.reserve(((size_t)B * ST_AUDIO_LEN * 2 + (size_t)B * 128 + (size_t)B * 6144 + nx + (size_t)B * Ttot * 1536 +
                            (size_t)B * n * (78 + 180 + 57)) * sizeof(float) + 16 * 256));
  Arena& a = m->longws;
  float* d_audio = a.take<float>((size_t)B * ST_AUDIO_LEN * 2);
  int32_t* d_word = a.take<int32_t>((size_t)B * 128);
  float* d_seed = a.take<float>((size_t)B * 6144);
  float* d_x = a.take<float>(nx);
  float* d_lat = latents_out ? latents_out : a.take<float>((size_t)B * Ttot * 1536);
  float* d_up = a.take<float>((size_t)B * n * 78);
  float* d_ha = a.take<float>((size_t)B * n * 180);
  float* d_lo = a.take<float>((size_t)B * n * 57);
  const size_t tok_pitch = (size_t)32 * 1536 * sizeof(float), lat_pitch = (size_t)Ttot * 1536 * sizeof(float);
  for (int i = 0; i < R; ++i) {
    ST_CHECK_CUDA(cudaMemcpy2DAsync(d_audio, (size_t)ST_AUDIO_LEN * 2 * sizeof(float), audio_long + (size_t)i * hop_a * 2,
                                    (size_t)audio_len * 2 * sizeof(float), (size_t)ST_AUDIO_LEN * 2 * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
    ST_CHECK_CUDA(cudaMemcpy2DAsync(d_word, 128 * sizeof(int32_t), word_long + (size_t)i * round_l, (size_t)n_words * sizeof(int32_t),
                                    128 * sizeof(int32_t), B, cudaMemcpyDeviceToDevice, s));
    st_cond c;
    c.audio = d_audio; c.word = d_word; c.seed = i == 0 ? seed0 : d_seed;
    for (int k = 0; k < 3; ++k) c.style[k] = style ? style[k] : nullptr;
    ST_TRY(st_cond_encode(m, &c, B, stream));
    const float* tape = noise_tape ? noise_tape + (size_t)i * sc->S * nx : nullptr;
    ST_TRY(st_sample(m, sc, g, x_init + (size_t)i * nx, tape, B, d_x, stream));
    // m->xs holds the sample token-major [B, 32, 1536]: hand the last 4 tokens over as the next seed, keep 32 / 28 tokens
    ST_CHECK_CUDA(cudaMemcpy2DAsync(d_seed, 6144 * sizeof(float), m->xs + 28 * 1536, tok_pitch, 6144 * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
    const int first = i == 0 ? 0 : 4, ntok = 32 - first, at = i == 0 ? 0 : 32 + 28 * (i - 1);
    ST_CHECK_CUDA(cudaMemcpy2DAsync(d_lat + (size_t)at * 1536, lat_pitch, m->xs + (size_t)first * 1536, tok_pitch,
                                    (size_t)ntok * 1536 * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
  }
  st_vq* const vq3[3] = {vq_upper, vq_hands, vq_lower};
  float* const rec3[3] = {d_up, d_ha, d_lo};
  ST_TRY(decode_three(m, vq3, d_lat, latent_scale, B, Ttot, rec3, s));
  ST_TRY(pose330(d_up, d_ha, d_lo, ms, ms + 330, ms + 660, ms + 663, jaw_aa, B, n, rec_pose, rec_trans, s));
  return ST_OK;
}

// Device time per launch of one GEMM in steady state: `reps` launches captured into one CUDA graph, timed by two
// events (no host launch latency in the number).  engine 0 = SIMT, 1 = tcgen05.
extern "C" int st_bench_gemm(int M, int N, int K, int engine, int reps, const float* A, const float* W, const float* bias, float* out,
                             double* ms_per_launch) {
  ST_REQUIRE(A && W && out && ms_per_launch && M > 0 && N > 0 && K > 0 && reps > 0, "st_bench_gemm: null argument");
  GemmP p = linear(A, M, K, W, bias, out, N);
  cudaStream_t s = nullptr;
  ST_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  const bool tc = engine == ST_ENGINE_TC;
  if (tc && !tc_supported(p)) { cudaStreamDestroy(s); set_error("st_bench_gemm: shape not supported by the tcgen05 engine"); return ST_EUNSUPPORTED; }
  __half* planes = nullptr;
  if (tc) {
    tc_forget_weights(W);
    ST_TRY(gemm_tc(p, s));                              // builds weight planes, sizes the scratch, splits A once
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    ST_CHECK_CUDA(cudaMalloc(&planes, (size_t)2 * M * K * sizeof(__half)));
    ST_TRY(tc_split(A, K, M, K, planes, s));            // steady state of the product: the operand arrives as planes
    p.a_planes = planes; p.a_plane_stride = (long long)M * K;
  }
  cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
  const int64_t l0 = g_launches;
  ST_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  int r = ST_OK;
  for (int i = 0; i < reps && r == ST_OK; ++i) r = tc ? gemm_tc(p, s) : gemm_simt(p, s);
  cudaError_t ce = cudaStreamEndCapture(s, &graph);
  g_launches = l0;
  if (r != ST_OK) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return r; }
  ST_CHECK_CUDA(ce);
  ST_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  ST_CHECK_CUDA(cudaGraphLaunch(exec, s));
  ST_CHECK_CUDA(cudaEventRecord(e0, s));
  ST_CHECK_CUDA(cudaGraphLaunch(exec, s));
  ST_CHECK_CUDA(cudaEventRecord(e1, s));
  ST_CHECK_CUDA(cudaStreamSynchronize(s));
  float t = 0.f;
  cudaEventElapsedTime(&t, e0, e1);
  *ms_per_launch = (double)t / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); tc_scratch_release(s); cudaStreamDestroy(s);
  if (tc) { tc_forget_weights(W); cudaFree(planes); }
  return ST_OK;
}

// stride-1 "same" Conv1d over channels-last activations as the engine runs it (implicit GEMM): A [B, L, C], W [N, taps * C]
// (tap-major, packer layout), out [B * L, N].  engine 0 = SIMT, 1 = tcgen05 (weight planes rebuilt on every call).
extern "C" int st_selftest_conv(int B, int L, int C, int N, int taps, int dil, int engine, const float* A, const float* W,
                                const float* bias, float* out, void* stream) {
  ST_REQUIRE(A && W && out && B > 0 && L > 0 && C > 0 && N > 0 && taps > 0 && dil > 0, "st_selftest_conv: null argument");
  GemmP p;
  p.A = A; p.W = W; p.bias = bias; p.out = out;
  p.M = B * L; p.N = N; p.K = taps * C; p.ldw = (p.K + 3) & ~3;
  p.Lout = L; p.Lin = L; p.C = C; p.stride = 1; p.pad = dil * (taps - 1) / 2; p.dil = dil;
  p.a_batch = (long long)L * C; p.lda = C; p.ldo = N;
  if (engine == ST_ENGINE_TC) {
    if (!tc_supported(p)) { set_error("st_selftest_conv: shape not supported by the tcgen05 engine"); return ST_EUNSUPPORTED; }
    tc_forget_weights(W);
    const int r = gemm_tc(p, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    tc_forget_weights(W);
    return r;
  }
  return gemm_simt(p, (cudaStream_t)stream);
}

extern "C" int st_selftest_gemm(int M, int N, int K, int engine, const float* A, const float* W, const float* bias, float* out,
                                void* stream) {
  ST_REQUIRE(A && W && out && M > 0 && N > 0 && K > 0, "st_selftest_gemm: null argument");
  GemmP p = linear(A, M, K, W, bias, out, N);
  if (engine == 2) {                       // tcgen05, weight planes kept from the previous call (caller keeps W unchanged)
    if (!tc_supported(p)) { set_error("st_selftest_gemm: shape not supported by the tcgen05 engine"); return ST_EUNSUPPORTED; }
    return gemm_tc(p, (cudaStream_t)stream);
  }
  if (engine == ST_ENGINE_TC) {
    if (!tc_supported(p)) { set_error("st_selftest_gemm: shape M=%d N=%d K=%d not supported by the tcgen05 engine", M, N, K); return ST_EUNSUPPORTED; }
    tc_forget_weights(W);                 // the caller may reuse the address with new contents
    const int r = gemm_tc(p, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    tc_forget_weights(W);
    return r;
  }
  return gemm_simt(p, (cudaStream_t)stream);
}
