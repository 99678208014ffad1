// Exact-fp32 kernels of the SynTalker sampling path for sm_100a: the implicit-GEMM SIMT engine (Linear and
// channels-last Conv1d in one kernel), LayerNorm, the 32-token attention core, RoPE + conditioning add,
// CFG combine + DDIM/DDPM update, the RVQ code selection and the 330-d pose assembly.
// Reference arithmetic each kernel reproduces is cited at the kernel.
#include "st_internal.cuh"

#include <math.h>

namespace st {

// =========================================================================================================
// 1. Implicit-GEMM SIMT engine.  128x64 CTA tile, BK=16, 256 threads, 8x4 register tile, double-buffered
//    shared memory with register prefetch.  A is gathered through the (b,t,j,c) map of GemmP so the same
//    kernel does nn.Linear (denoiser.py, transformer.py), the k=15 strided WavEncoder convs
//    (denoiser.py:304-322, layer.py:144-184) and the dilated / upsampled k=3 decoder convs
//    (encdec.py:37-68, resnet.py:12-69) over channels-last activations.
// =========================================================================================================
constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_GELU: return v * 0.5f * (1.0f + erff(v * 0.70710678118654752440f));   // exact-erf GELU (transformer.py:131 nn.GELU)
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_LRELU: return v > 0.0f ? v : v * 0.01f;                                 // nn.LeakyReLU default slope (layer.py:146)
    default: return v;
  }
}

// packed fp32 pair FMA (sm_100): (d0, d1) += (a0, a1) * (b0, b1)
__device__ __forceinline__ void ffma2k(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

__device__ __forceinline__ void split16(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v * kActScale, -65000.0f), 65000.0f);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
// store 4 consecutive values as fp16 hi/lo planes (8-byte stores)
__device__ __forceinline__ void store_planes4(__half* hi_ptr, long long plane_stride, float a, float b, float c, float d) {
  __half h[4], l[4];
  split16(a, h[0], l[0]); split16(b, h[1], l[1]); split16(c, h[2], l[2]); split16(d, h[3], l[3]);
  __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
  __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
  uint2 hv, lv;
  hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
  lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
  *reinterpret_cast<uint2*>(hi_ptr) = hv;
  *reinterpret_cast<uint2*>(hi_ptr + plane_stride) = lv;
}

template <bool VEC_A>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_simt_kernel(GemmP p) {
  pdl_wait();
  trace_stamp(8);
  pdl_launch();
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A gather bookkeeping: two (row, k4) slots per thread ----
  int a_row[2], a_k4[2], a_basel[2];
  const float* a_ptr[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int slot = tid + i * GEMM_THREADS;
    a_row[i] = slot >> 2;
    a_k4[i] = slot & 3;
    const int m = m0 + a_row[i];
    a_ok[i] = m < p.M;
    const int mm = a_ok[i] ? m : 0;
    const int b = mm / p.Lout;
    const int t = mm - b * p.Lout;
    a_basel[i] = t * p.stride - p.pad;
    a_ptr[i] = p.A + (long long)b * p.a_batch;
  }
  const int b_n = tid >> 2, b_k4 = tid & 3;
  const bool b_ok = (n0 + b_n) < p.N;
  const float* b_ptr = p.W + (long long)(n0 + (b_ok ? b_n : 0)) * p.ldw;
  const int lim = p.ups ? 2 * p.Lin : p.Lin;

  float4 ra[2], rb;
  auto load_tile = [&](int kt) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = kt * BK + a_k4[i] * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (VEC_A) {
        if (a_ok[i] && k < p.K) {
          const int j = k / p.C;
          const int c = k - j * p.C;
          int l = a_basel[i] + j * p.dil;
          if (l >= 0 && l < lim) {
            if (p.ups) l >>= 1;
            v = __ldg(reinterpret_cast<const float4*>(a_ptr[i] + (long long)l * p.lda + c));
          }
        }
      } else {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (a_ok[i]) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kk = k + q;
            if (kk < p.K) {
              const int j = kk / p.C;
              const int c = kk - j * p.C;
              int l = a_basel[i] + j * p.dil;
              if (l >= 0 && l < lim) {
                if (p.ups) l >>= 1;
                e[q] = __ldg(a_ptr[i] + (long long)l * p.lda + c);
              }
            }
          }
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
      if (p.a_relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      ra[i] = v;
    }
    const int kb = kt * BK + b_k4 * 4;
    rb = (b_ok && kb < p.ldw) ? __ldg(reinterpret_cast<const float4*>(b_ptr + kb)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kk = a_k4[i] * 4;
      As[buf][kk + 0][a_row[i]] = ra[i].x;
      As[buf][kk + 1][a_row[i]] = ra[i].y;
      As[buf][kk + 2][a_row[i]] = ra[i].z;
      As[buf][kk + 3][a_row[i]] = ra[i].w;
    }
    const int kk = b_k4 * 4;
    Bs[buf][kk + 0][b_n] = rb.x;
    Bs[buf][kk + 1][b_n] = rb.y;
    Bs[buf][kk + 2][b_n] = rb.z;
    Bs[buf][kk + 3][b_n] = rb.w;
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nkt = (p.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nkt) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const int nb = n0 + tx * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (nb + j < p.N) bias[j] = __ldg(p.bias + nb + j);
  }
  const bool vec_out = ((p.ldo & 3) == 0) && (nb + 3 < p.N);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    float v[4];
    const float* rrow = p.res ? p.res + (long long)(m / p.res_div) * p.ldr : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j] * p.out_scale + bias[j];
      const float r = (rrow && nb + j < p.N) ? rrow[nb + j] : 0.f;
      if (p.res_mode == RES_PRE) x += r;
      x = apply_act(x, p.act);
      if (p.res_mode == RES_POST) x += r;
      v[j] = x;
    }
    float* orow = p.out + (long long)m * p.ldo;
    if (vec_out) {
      *reinterpret_cast<float4*>(orow + nb) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (nb + j < p.N) orow[nb + j] = v[j];
    }
  }
}

GemmP linear(const float* A, int M, int K, const float* W, const float* bias, float* out, int N) {
  GemmP p;
  p.A = A; p.W = W; p.bias = bias; p.out = out;
  p.M = M; p.N = N; p.K = K; p.ldw = K;
  p.Lout = M; p.Lin = M; p.C = K; p.lda = K; p.ldo = N;
  return p;
}

int gemm_simt(const GemmP& p, cudaStream_t s) {
  ST_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  ST_REQUIRE((p.ldw & 3) == 0 && p.ldw >= p.K, "gemm: ldw=%d must be a multiple of 4 and >= K=%d", p.ldw, p.K);
  ST_REQUIRE(p.A && p.out && !p.ln_stats && !p.stats_out, "gemm_simt: fp32 operand and output required; LayerNorm folding is a tcgen05-engine feature");
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
  const bool vec = ((p.C & 3) == 0) && ((p.lda & 3) == 0) && ((p.a_batch & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
  if (vec) launch_k(gemm_simt_kernel<true>, dim3(grid), dim3(GEMM_THREADS), 0, s, p);
  else launch_k(gemm_simt_kernel<false>, dim3(grid), dim3(GEMM_THREADS), 0, s, p);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 2. LayerNorm over 512 channels, eps 1e-5, affine (transformer.py:172,185).  One warp per row.
// =========================================================================================================
__global__ void __launch_bounds__(256) layernorm512_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ b, float* __restrict__ y,
                                                           __half* __restrict__ planes, int rows) {
  pdl_wait();
  trace_stamp(7);
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * 512);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 512.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q * (1.0f / 512.0f) + 1e-5f);
  float4* yr = y ? reinterpret_cast<float4*>(y + (long long)row * 512) : nullptr;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 gg = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * gg.x + bb.x;
    o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
    o.z = (v[i].z - mean) * rstd * gg.z + bb.z;
    o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
    if (yr) yr[lane + 32 * i] = o;
    if (planes) store_planes4(planes + (long long)row * 512 + (lane + 32 * i) * 4, (long long)rows * 512, o.x, o.y, o.z, o.w);
  }
}

int layernorm512(const float* x, const float* gamma, const float* beta, float* y, __half* planes, int rows, cudaStream_t s) {
  launch_k(layernorm512_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, gamma, beta, y, planes, rows);
  ST_CHECK_LAUNCH();
  return ST_OK;
}


// =========================================================================================================
// 2b. First WavEncoder block, raw-audio side (denoiser.py:308-315 block 0, layer.py:159-163): conv1 (k15, stride 5,
//     pad 1700, BatchNorm folded, LeakyReLU) and the conv shortcut (same geometry, no activation) share their input
//     window, so one pass computes both: C_in = 2 makes this a 30-tap dot product per output, FMA / HBM-write bound,
//     not a GEMM.  A warp walks output positions; lane l owns channels 2l, 2l+1 of both convs (120 weights in
//     registers), the 30 inputs of a position are broadcast reads from a staged window.  conv1 leaves fp16 hi/lo
//     planes for the tcgen05 conv that follows, the shortcut stays fp32 (it is a residual).
// =========================================================================================================
constexpr int WF_POS = 256;                 // output positions per staged window
constexpr int WF_WIN = 4;                   // windows per CTA (the per-CTA weight fetch is amortised over 1024 positions)
__global__ void __launch_bounds__(256) wav_first_kernel(const float* __restrict__ audio, const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ wd, const float* __restrict__ bd, int ldw, int Lin, int Lout,
                                                        int stride, int pad, __half* __restrict__ h1_planes, long long plane_stride,
                                                        float* __restrict__ sc) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  __shared__ __align__(16) float xs[(WF_POS * 5 + 16) * 2];
  __shared__ __align__(16) float ws[2][30][64];                  // [conv][tap*2 + c_in][channel]: transposed so that lanes read consecutive words
  const int clip = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * 64 * 30; i += 256) {
    const int cv = i / (64 * 30), r = i - cv * 64 * 30, ch = r / 30, k = r - ch * 30;
    ws[cv][k][ch] = __ldg((cv ? wd : w1) + (long long)ch * ldw + k);
  }
  __syncthreads();
  float wa[2][30], wb[2][30];
#pragma unroll
  for (int k = 0; k < 30; ++k) {
    const float2 u = *reinterpret_cast<const float2*>(&ws[0][k][2 * lane]), v = *reinterpret_cast<const float2*>(&ws[1][k][2 * lane]);
    wa[0][k] = u.x; wa[1][k] = u.y; wb[0][k] = v.x; wb[1][k] = v.y;
  }
  const float ba0 = __ldg(b1 + 2 * lane), ba1 = __ldg(b1 + 2 * lane + 1), bb0 = __ldg(bd + 2 * lane), bb1 = __ldg(bd + 2 * lane + 1);
  const float* a = audio + (long long)clip * Lin * 2;
  const int nwin = (WF_POS - 1) * stride + 15;
  for (int win = 0; win < WF_WIN; ++win) {
    const int t0 = (blockIdx.x * WF_WIN + win) * WF_POS;
    if (t0 >= Lout) break;
    const int l0 = t0 * stride - pad;                            // first input position of the window
    __syncthreads();                                             // the previous window is consumed
    for (int i = threadIdx.x; i < nwin; i += 256) {
      const int l = l0 + i;
      float2 v = make_float2(0.f, 0.f);
      if (l >= 0 && l < Lin) v = __ldg(reinterpret_cast<const float2*>(a) + l);
      reinterpret_cast<float2*>(xs)[i] = v;
    }
    __syncthreads();
    // two output positions per iteration (8 independent packed-FMA chains per warp)
    for (int tt = warp; tt < WF_POS; tt += 16) {
      const int tA = t0 + tt, tB = tA + 8;
      if (tA >= Lout) break;
      const float2* xa = reinterpret_cast<const float2*>(xs) + tt * stride;
      const float2* xb = xa + 8 * stride;
      float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};   // (conv1 ch0, conv1 ch1, shortcut ch0, shortcut ch1)
#pragma unroll
      for (int j = 0; j < 15; ++j) {
        const float2 x = xa[j], y = xb[j];                          // (c_in 0, c_in 1) of input position j: broadcast reads
        ffma2k(accA[0], accA[1], x.x, x.x, wa[0][2 * j], wa[1][2 * j]);     ffma2k(accA[2], accA[3], x.x, x.x, wb[0][2 * j], wb[1][2 * j]);
        ffma2k(accB[0], accB[1], y.x, y.x, wa[0][2 * j], wa[1][2 * j]);     ffma2k(accB[2], accB[3], y.x, y.x, wb[0][2 * j], wb[1][2 * j]);
        ffma2k(accA[0], accA[1], x.y, x.y, wa[0][2 * j + 1], wa[1][2 * j + 1]); ffma2k(accA[2], accA[3], x.y, x.y, wb[0][2 * j + 1], wb[1][2 * j + 1]);
        ffma2k(accB[0], accB[1], y.y, y.y, wa[0][2 * j + 1], wa[1][2 * j + 1]); ffma2k(accB[2], accB[3], y.y, y.y, wb[0][2 * j + 1], wb[1][2 * j + 1]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = u ? tB : tA;
        if (t >= Lout) break;
        const float* acc = u ? accB : accA;
        float a0 = acc[0] + ba0, a1 = acc[1] + ba1;
        a0 = a0 > 0.f ? a0 : a0 * 0.01f; a1 = a1 > 0.f ? a1 : a1 * 0.01f;
        const long long row = (long long)clip * Lout + t;
        __half h0, l0h, h1, l1h;
        split16(a0, h0, l0h); split16(a1, h1, l1h);
        *reinterpret_cast<__half2*>(h1_planes + row * 64 + 2 * lane) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2*>(h1_planes + plane_stride + row * 64 + 2 * lane) = __halves2half2(l0h, l1h);
        *reinterpret_cast<float2*>(sc + row * 64 + 2 * lane) = make_float2(acc[2] + bb0, acc[3] + bb1);
      }
    }
  }
}

int wav_first(const float* audio, const float* w1, const float* b1, const float* wd, const float* bd, int ldw, int cb, int Lin, int Lout,
              int stride, int pad, __half* h1_planes, long long plane_stride, float* sc, cudaStream_t s) {
  if (stride != 5) { set_error("wav_first: built for the stride-5 first block"); return ST_EINVAL; }
  launch_k(wav_first_kernel, dim3((Lout + WF_POS * WF_WIN - 1) / (WF_POS * WF_WIN), cb), dim3(256), 0, s, audio, w1, b1, wd, bd, ldw, Lin, Lout, stride, pad,
           h1_planes, plane_stride, sc);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 3. Attention core for one (sequence, head): 32 tokens x 128 dims, softmax(q k^T / sqrt(128)) v, no mask
//    (transformer.py:83-104; qkv channel = which*512 + head*128 + d).  256 threads, fp32 throughout.
// =========================================================================================================
constexpr int ATT_LD = 132;   // padded row stride (floats) -> conflict-free float4 reads

__global__ void __launch_bounds__(256) attention32_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                          __half* __restrict__ planes, long long plane_stride) {
  pdl_wait();
  trace_stamp(2);
  pdl_launch();
  // One CTA per (sequence, head, half of the 32 query rows): 512 CTAs for a config-2 step keep ~28 warps per SM busy and
  // halve each warp's instruction stream (the kernel is latency-bound: 1800 dependent instructions per warp before).
  extern __shared__ __align__(16) float att_smem[];
  float (*q)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem);                    // 16 rows
  float (*k)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem + 16 * ATT_LD);      // 32 rows
  float (*v)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem + 48 * ATT_LD);      // 32 rows
  float (*pt)[20] = reinterpret_cast<float (*)[20]>(att_smem + 80 * ATT_LD);             // softmax, transposed: pt[j][i_local]
  const int seq = blockIdx.x, head = blockIdx.y, qh = blockIdx.z, tid = threadIdx.x;
  const float* base = qkv + (long long)seq * 32 * 1536 + head * 128;
  for (int i = tid; i < 32 * 32; i += 256) {
    const int r = i >> 5, c4 = i & 31;
    const float* src = base + (long long)r * 1536 + c4 * 4;
    *reinterpret_cast<float4*>(&k[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(src + 512));
    *reinterpret_cast<float4*>(&v[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(src + 1024));
    if (r < 16) *reinterpret_cast<float4*>(&q[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(src + (long long)qh * 16 * 1536));
  }
  __syncthreads();
  // ---- scores: thread = 2 rows x 4 cols of S over one interleaved quarter of the 128 dims ----
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane & 3, bj = lane >> 2;              // a warp holds 2 full rows of S
  float acc[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll
  for (int d4 = 0; d4 < 8; ++d4) {
    const int col = (4 * d4 + g) * 4;
    float4 qa[2], kb[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) qa[r] = *reinterpret_cast<const float4*>(&q[2 * warp + r][col]);
#pragma unroll
    for (int c = 0; c < 4; ++c) kb[c] = *reinterpret_cast<const float4*>(&k[4 * bj + c][col]);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[r][c] = fmaf(qa[r].x, kb[c].x, acc[r][c]);
        acc[r][c] = fmaf(qa[r].y, kb[c].y, acc[r][c]);
        acc[r][c] = fmaf(qa[r].z, kb[c].z, acc[r][c]);
        acc[r][c] = fmaf(qa[r].w, kb[c].w, acc[r][c]);
      }
  }
  const float scale = 0.08838834764831845f;            // 128^-0.5
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float a = acc[r][c];
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      acc[r][c] = a * scale;
      mx = fmaxf(mx, acc[r][c]);
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[r][c] = expf(acc[r][c] - mx); sum += acc[r][c]; }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] *= inv;
  }
  if (g == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<float2*>(&pt[4 * bj + c][2 * warp]) = make_float2(acc[0][c], acc[1][c]);
  }
  __syncthreads();
  // ---- out = P V: warp = 2 rows, lane = 4 columns ----
  float4 o4[2];
  o4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
  o4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const float2 p2 = *reinterpret_cast<const float2*>(&pt[j][2 * warp]);
    const float4 vv = *reinterpret_cast<const float4*>(&v[j][lane * 4]);
    o4[0].x = fmaf(p2.x, vv.x, o4[0].x); o4[0].y = fmaf(p2.x, vv.y, o4[0].y);
    o4[0].z = fmaf(p2.x, vv.z, o4[0].z); o4[0].w = fmaf(p2.x, vv.w, o4[0].w);
    o4[1].x = fmaf(p2.y, vv.x, o4[1].x); o4[1].y = fmaf(p2.y, vv.y, o4[1].y);
    o4[1].z = fmaf(p2.y, vv.z, o4[1].z); o4[1].w = fmaf(p2.y, vv.w, o4[1].w);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const long long off = ((long long)seq * 32 + qh * 16 + 2 * warp + r) * 512 + head * 128 + lane * 4;
    if (out) *reinterpret_cast<float4*>(out + off) = o4[r];
    if (planes) store_planes4(planes + off, plane_stride, o4[r].x, o4[r].y, o4[r].z, o4[r].w);
  }
}

int attention32(const float* qkv, float* out, __half* planes, int nseq, cudaStream_t s, long long plane_stride) {
  constexpr int smem = (80 * ATT_LD + 32 * 20) * sizeof(float);
  static DeviceOnce attr_set;   // the attribute is per device
  if (attr_set.first()) ST_CHECK_CUDA(cudaFuncSetAttribute(attention32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  launch_k(attention32_kernel, dim3(nseq, 4, 2), dim3(256), smem, s, qkv, out, planes, plane_stride ? plane_stride : (long long)nseq * 32 * 512);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 4. Token prologue: x_e[b,tau,:] = RoPE( z[b,tau,:] + vt[t_b,:] + cst_e[b,tau,:] + g2[b,:] + sv_e[b,:] )
//    = input_process / input_process2 / input_process3 with the step-invariant terms hoisted
//    (denoiser.py:160-174) followed by the rotary embedding on 8 groups of 64 (denoiser.py:178-186,324-343).
//    One thread per (eval, row, pair j<256): columns g*64 + j' and g*64 + j' + 32.
// =========================================================================================================
__global__ void __launch_bounds__(256) tokens_in_kernel(TokensInP p) {
  pdl_wait();
  trace_stamp(3);
  pdl_launch();
  if (p.sig_zero && blockIdx.x == 0)
    for (int i = threadIdx.x; i < p.sig_n; i += 256) p.sig_zero[i] = 0u;
  // one warp per (evaluation, row); lane l owns the rotary pairs (g*64 + l, g*64 + l + 32), g = 0..7
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long rows_e = (long long)p.B * 32;
  if (wid >= rows_e * p.nE) return;
  const int e = (int)(wid / rows_e);
  const int row = (int)(wid - e * rows_e);          // b*32 + tau
  const int b = row >> 5, tau = row & 31;
  // a timestep that reaches the device out of range is clamped into the table (the host mirror validates CPU tensors; checking a
  // device tensor would cost a synchronisation per model call)
  const int t = p.ls ? p.t_model_dev[p.ls->k] : min(max(p.t_dev ? (int)p.t_dev[b] : p.t_scalar, 0), ST_MAX_T - 1);
  const float* vt = p.vt_table + (long long)t * 512;
  const float* cst = p.cst[e] + (long long)(p.cst_bcast[e] ? tau : row) * 512;
  const float* g2 = p.g2 + (long long)b * 512;
  const float* zr = p.z + (long long)row * 512;
  const float* sv = p.sv[e] ? p.sv[e] + (p.sv_bcast[e] ? 0 : (long long)b * 512) : nullptr;
  const float cs = p.rope_cos[tau * 32 + lane], sn = p.rope_sin[tau * 32 + lane];
  const long long orow = (long long)e * rows_e + row;
  float v1[8], v2[8];
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int c1 = g * 64 + lane, c2 = c1 + 32;
    float x1 = zr[c1] + vt[c1] + cst[c1] + g2[c1];
    float x2 = zr[c2] + vt[c2] + cst[c2] + g2[c2];
    if (sv) { x1 += sv[c1]; x2 += sv[c2]; }
    v1[g] = __fadd_rn(__fmul_rn(x1, cs), __fmul_rn(-x2, sn));
    v2[g] = __fadd_rn(__fmul_rn(x2, cs), __fmul_rn(x1, sn));
    sum += v1[g] + v2[g];
  }
  float* xo = p.x + orow * 512;
#pragma unroll
  for (int g = 0; g < 8; ++g) { xo[g * 64 + lane] = v1[g]; xo[g * 64 + lane + 32] = v2[g]; }
  if (p.x_planes) {
    const long long ps = rows_e * p.nE * 512;
    __half* xp = p.x_planes + orow * 512;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      __half h, l;
      split16(v1[g], h, l); xp[g * 64 + lane] = h; xp[ps + g * 64 + lane] = l;
      split16(v2[g], h, l); xp[g * 64 + lane + 32] = h; xp[ps + g * 64 + lane + 32] = l;
    }
  }
  if (p.stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / 512.0f);
    float m2 = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) { const float d1 = v1[g] - mean, d2 = v2[g] - mean; m2 += d1 * d1 + d2 * d2; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    // sixteen identical partials (mean, M2/16) combine to exactly (mean, M2)
    if (lane < 16) *reinterpret_cast<float2*>(p.stats + (orow * 16 + lane) * 2) = make_float2(mean, m2 * 0.0625f);
  }
}

int tokens_in(const TokensInP& p, cudaStream_t s) {
  const long long n = (long long)p.nE * p.B * 32;      // warps
  launch_k(tokens_in_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, s, p);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// ---- 4b. the same for a step of the z recursion: TS_SPLIT warps per row, each over 8 / TS_SPLIT rotary groups of 64 columns ----
// Everything that depends on the step (timestep row of the embedding table, alpha, beta) is a launch parameter: the captured
// graph holds the whole loop, so every step's node carries its own values and the kernel starts without a chain of dependent
// loads through the loop state.  All loads of a warp are issued before its first store.
constexpr int TS_SPLIT = 4;                 // warps per row: each takes 8 / TS_SPLIT of the eight 64-column rotary groups
constexpr int TS_NG = 8 / TS_SPLIT, TS_NV = 2 * TS_NG;
__global__ void __launch_bounds__(256) tokens_step_kernel(TokensStepP q) {
  pdl_wait();
  trace_stamp(3);
  pdl_launch();
  const TokensInP& p = q.t;
  if (p.sig_zero && blockIdx.x == 0)
    for (int i = threadIdx.x; i < p.sig_n; i += 256) p.sig_zero[i] = 0u;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long rows_e = (long long)p.B * 32;
  if (wid >= rows_e * TS_SPLIT) return;
  const int row = (int)(wid / TS_SPLIT), part = (int)(wid % TS_SPLIT);
  const int b = row >> 5, tau = row & 31;
  float* __restrict__ zr = q.z_rw + (long long)row * 512;
  int col[TS_NV];
#pragma unroll
  for (int g = 0; g < TS_NG; ++g) { col[2 * g] = (part * TS_NG + g) * 64 + lane; col[2 * g + 1] = col[2 * g] + 32; }
  float zv[TS_NV], vtv[TS_NV], g2v[TS_NV], cv[2][TS_NV], sw[2][TS_NV];
  const float* __restrict__ vt = q.vt;
  const float* __restrict__ g2 = p.g2 + (long long)b * 512;
  auto load_eval = [&](int e, float* c, float* s_) {
    const float* __restrict__ cst = p.cst[e] + (long long)(p.cst_bcast[e] ? tau : row) * 512;
    const float* __restrict__ sv = p.sv[e] ? p.sv[e] + (p.sv_bcast[e] ? 0 : (long long)b * 512) : nullptr;
#pragma unroll
    for (int i = 0; i < TS_NV; ++i) { c[i] = __ldg(cst + col[i]); s_[i] = sv ? __ldg(sv + col[i]) : 0.f; }
    return sv != nullptr;
  };
#pragma unroll
  for (int i = 0; i < TS_NV; ++i) { zv[i] = zr[col[i]]; vtv[i] = __ldg(vt + col[i]); g2v[i] = __ldg(g2 + col[i]); }
  bool has_sv = load_eval(0, cv[0], sw[0]);
  const float cs = __ldg(p.rope_cos + tau * 32 + lane), sn = __ldg(p.rope_sin + tau * 32 + lane);
  if (!q.first) {
    // x_k = alpha x0_hat + beta x_{k+1}  (gaussian_diffusion.py:772-790, sigma = 0) applied to z = W_x x
    const float* __restrict__ P0 = q.P + (long long)row * 512;
    const long long pe = rows_e * 512;
    float pm[TS_NV], cx[TS_NV];
#pragma unroll
    for (int i = 0; i < TS_NV; ++i) cx[i] = __ldg(q.c_xo + col[i]);
    if (q.cfg_mode == ST_CFG_NONE) {
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) pm[i] = __ldg(P0 + col[i]);
    } else if (q.cfg_mode == ST_CFG_TEXT) {
      const float sc = __ldg(q.scale + b);            // eval 0 = conditional, eval 1 = unconditional (cfg_sampler.py:28)
      float c[TS_NV], u[TS_NV];
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) { c[i] = __ldg(P0 + col[i]); u[i] = __ldg(P0 + pe + col[i]); }
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) pm[i] = u[i] + sc * (c[i] - u[i]);
    } else {
      const float sa = __ldg(q.scale + b), sp = __ldg(q.scale2 + b);   // eval 0 = uu, 1 = ut, 2 = ua (cfg_sampler.py:54)
      float uu[TS_NV], ut[TS_NV], ua[TS_NV];
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) { uu[i] = __ldg(P0 + col[i]); ut[i] = __ldg(P0 + pe + col[i]); ua[i] = __ldg(P0 + 2 * pe + col[i]); }
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) pm[i] = uu[i] + sa * (ut[i] - uu[i]) + sp * (ua[i] - uu[i]);
    }
#pragma unroll
    for (int i = 0; i < TS_NV; ++i) zv[i] = q.beta * zv[i] + q.alpha * (pm[i] + cx[i]);
    if (q.sigma != 0.f) {
      // stochastic step (DDPM gaussian_diffusion.py:556, DDIM with eta > 0 :781-790): + sigma W_x eps, the noise already in token space
      const float* __restrict__ ze = q.zeps + (long long)row * 512;
      float nz[TS_NV];
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) nz[i] = __ldg(ze + col[i]);
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) zv[i] = fmaf(q.sigma, nz[i], zv[i]);
    }
  }
  float base[TS_NV];
#pragma unroll
  for (int i = 0; i < TS_NV; ++i) base[i] = zv[i] + vtv[i];
  for (int e = 0; e < p.nE; ++e) {
    const int cur = e & 1, nxt = cur ^ 1;
    const bool sv_cur = has_sv;
    if (e + 1 < p.nE) has_sv = load_eval(e + 1, cv[nxt], sw[nxt]);     // the next evaluation's terms are in flight during this one's stores
    const long long orow = (long long)e * rows_e + row;
    float v[TS_NV];
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < TS_NG; ++g) {
      float x1 = base[2 * g] + cv[cur][2 * g] + g2v[2 * g];
      float x2 = base[2 * g + 1] + cv[cur][2 * g + 1] + g2v[2 * g + 1];
      if (sv_cur) { x1 += sw[cur][2 * g]; x2 += sw[cur][2 * g + 1]; }
      v[2 * g] = __fadd_rn(__fmul_rn(x1, cs), __fmul_rn(-x2, sn));
      v[2 * g + 1] = __fadd_rn(__fmul_rn(x2, cs), __fmul_rn(x1, sn));
      sum += v[2 * g] + v[2 * g + 1];
    }
    float* xo = p.x + orow * 512;
#pragma unroll
    for (int i = 0; i < TS_NV; ++i) xo[col[i]] = v[i];
    if (p.x_planes) {
      const long long ps = rows_e * p.nE * 512;
      __half* xp = p.x_planes + orow * 512;
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) { __half h, l; split16(v[i], h, l); xp[col[i]] = h; xp[ps + col[i]] = l; }
    }
    if (p.stats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.0f / (64.0f * TS_NG));
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < TS_NV; ++i) { const float d = v[i] - mean; m2 += d * d; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      // 2 TS_NG identical partials (mean, M2 / (2 TS_NG)) of this part of the row combine with the other parts' to exactly (mean, M2)
      if (lane < 2 * TS_NG) *reinterpret_cast<float2*>(p.stats + (orow * 16 + part * 2 * TS_NG + lane) * 2) = make_float2(mean, m2 * (0.5f / TS_NG));
    }
  }
  if (!q.first) {
#pragma unroll
    for (int i = 0; i < TS_NV; ++i) zr[col[i]] = zv[i];
  }
}

int tokens_step(const TokensStepP& p, cudaStream_t s) {
  const long long n = (long long)TS_SPLIT * p.t.B * 32;      // warps
  launch_k(tokens_step_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, s, p);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 5. CFG combine + sampler update on the token-major state [B*32,1536].
//    CFG: cfg_sampler.py:28 / :54 / :67-117.   DDIM: gaussian_diffusion.py:772-790 (eps re-derived from x0).
//    DDPM: gaussian_diffusion.py:383 (posterior mean) + :556.  Intrinsics keep the reference's
//    separate fp32 roundings (no FMA contraction), so given equal inputs the update is bit-identical.
// =========================================================================================================
__global__ void __launch_bounds__(256) step_update_kernel(StepP p) {
  pdl_wait();
  trace_stamp(4);
  pdl_launch();
  const long long n4 = (long long)p.B * 32 * 1536 / 4;
  const long long i4 = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i4 >= n4) return;
  const long long e0 = i4 * 4;
  const int row = (int)(e0 / 1536);
  const int col = (int)(e0 - (long long)row * 1536);
  const int b = row >> 5, tau = row & 31;
  const long long per_eval = (long long)p.B * 32 * 1536;
  auto ld = [&](int e) { return *reinterpret_cast<const float4*>(p.o + e * per_eval + e0); };
  float x0[4];
  if (p.cfg_mode == ST_CFG_NONE) {
    const float4 a = ld(0);
    x0[0] = a.x; x0[1] = a.y; x0[2] = a.z; x0[3] = a.w;
  } else if (p.cfg_mode == ST_CFG_TEXT) {
    const float4 c = ld(0), u = ld(1);                  // eval 0 = conditional, eval 1 = unconditional
    const float sc = p.scale[b];
    const float cv[4] = {c.x, c.y, c.z, c.w}, uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) x0[q] = __fadd_rn(uv[q], __fmul_rn(sc, __fsub_rn(cv[q], uv[q])));
  } else if (p.cfg_mode == ST_CFG_BODYPART1) {
    // eval 0 = out_uncond (no prompt, audio), eval >= 1 = (prompt of a part, no audio); cfg_sampler.py:149-167
    const float4 u = ld(0);
    const int ua_e = p.part_ua[col / 512];
    const float sc = p.scale[b];
    const float uv[4] = {u.x, u.y, u.z, u.w};
    float cv[4] = {uv[0], uv[1], uv[2], uv[3]};       // a part without a prompt keeps out_uncond: u + scale * (u - u)
    if (ua_e >= 0) { const float4 c = ld(ua_e); cv[0] = c.x; cv[1] = c.y; cv[2] = c.z; cv[3] = c.w; }
#pragma unroll
    for (int q = 0; q < 4; ++q) x0[q] = __fadd_rn(uv[q], __fmul_rn(sc, __fsub_rn(cv[q], uv[q])));
  } else {
    // eval 0 = uu (no prompt, no audio), eval 1 = ut (no prompt, audio), eval >= 2 = ua (prompt, no audio)
    const float4 uu = ld(0), ut = ld(1);
    float sa, sp;
    int ua_e;
    if (p.cfg_mode == ST_CFG_TWO) { sa = p.scale[b]; sp = p.scale2[b]; ua_e = 2; }
    else { const int part = col / 512; sa = p.part_sa[part]; sp = p.part_sp[part]; ua_e = p.part_ua[part]; }
    const float uuv[4] = {uu.x, uu.y, uu.z, uu.w}, utv[4] = {ut.x, ut.y, ut.z, ut.w};
    float uav[4] = {0.f, 0.f, 0.f, 0.f};
    if (ua_e >= 0) { const float4 ua = ld(ua_e); uav[0] = ua.x; uav[1] = ua.y; uav[2] = ua.z; uav[3] = ua.w; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float acc = __fadd_rn(uuv[q], __fmul_rn(sa, __fsub_rn(utv[q], uuv[q])));
      if (ua_e >= 0) acc = __fadd_rn(acc, __fmul_rn(sp, __fsub_rn(uav[q], uuv[q])));
      x0[q] = acc;
    }
  }
  float4* xs4 = reinterpret_cast<float4*>(p.xs + e0);
  if (p.mode < 0) { *xs4 = make_float4(x0[0], x0[1], x0[2], x0[3]); return; }
  float cf[ST_COEF_STRIDE];
  const float* eps_ptr = p.eps;
  if (p.ls) {
    const int k = p.ls->k;
#pragma unroll
    for (int q = 0; q < ST_COEF_STRIDE; ++q) cf[q] = p.coef_dev[k * ST_COEF_STRIDE + q];
    eps_ptr = p.ls->tape ? p.ls->tape + (long long)(p.ls->S - 1 - k) * p.B * 1536 * 32 : nullptr;
  } else {
#pragma unroll
    for (int q = 0; q < ST_COEF_STRIDE; ++q) cf[q] = p.c[q];
  }
  const float4 xv4 = *xs4;
  const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
  float nz[4] = {0.f, 0.f, 0.f, 0.f};
  const float sigma = p.mode == ST_MODE_DDPM ? cf[2] : cf[4];
  if (eps_ptr && sigma != 0.f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) nz[q] = eps_ptr[((long long)b * 1536 + col + q) * 32 + tau];   // caller layout [B,1536,1,32]
  }
  float o[4];
  if (p.mode == ST_MODE_DDPM) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float mean = __fadd_rn(__fmul_rn(cf[0], x0[q]), __fmul_rn(cf[1], xv[q]));
      o[q] = sigma != 0.f ? __fadd_rn(mean, __fmul_rn(sigma, nz[q])) : mean;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(cf[0], xv[q]), x0[q]), cf[1]);
      const float mean = __fadd_rn(__fmul_rn(x0[q], cf[2]), __fmul_rn(cf[3], eps));
      o[q] = sigma != 0.f ? __fadd_rn(mean, __fmul_rn(sigma, nz[q])) : mean;
    }
  }
  *xs4 = make_float4(o[0], o[1], o[2], o[3]);
  // the next step's first GEMM streams the state as fp16 hi/lo planes: emit them here instead of a split pass
  if (p.xs_planes) store_planes4(p.xs_planes + e0, per_eval, o[0], o[1], o[2], o[3]);
  if (p.ls_advance) {
    // last CTA to finish moves the loop to the next step (every thread of every CTA has read ls->k by then)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned t = atomicAdd(&p.ls_advance->done, 1u);
      if (t == gridDim.x - 1) { p.ls_advance->done = 0; p.ls_advance->k -= 1; }
    }
  }
}

__global__ void advance_loop_kernel(LoopState* ls) {
  pdl_wait();
  trace_stamp(5);
  pdl_launch();
  ls->k -= 1;
}
__global__ void init_loop_kernel(LoopState* ls, int S, int k, const float* tape) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  ls->k = k; ls->S = S; ls->tape = tape; ls->done = 0;
}
int advance_loop(LoopState* ls, cudaStream_t s) {
  launch_k(advance_loop_kernel, dim3(1), dim3(1), 0, s, ls);
  ST_CHECK_LAUNCH();
  return ST_OK;
}
int init_loop(LoopState* ls, int S, int k, const float* tape, cudaStream_t s) {
  launch_k(init_loop_kernel, dim3(1), dim3(1), 0, s, ls, S, k, tape);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int step_update(const StepP& p, cudaStream_t s) {
  const long long n4 = (long long)p.B * 32 * 1536 / 4;
  launch_k(step_update_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, p);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 6. Layout: [B,C,T] <-> [B,T,C] tile transposes (trainer:457 squeeze().permute(1,0); OutputProcess permute).
// =========================================================================================================
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc,
                                                        float scale) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  // in: [B, R, Cc] -> out: [B, Cc, R]
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* ib = in + (long long)b * R * Cc;
  float* ob = out + (long long)b * R * Cc;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < R && c < Cc) tile[i][tx] = ib[(long long)r * Cc + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < Cc) ob[(long long)c * R + r] = tile[tx][i] * scale;
  }
}

int transpose_to_tokens(const float* x, float* tok, int B, int C, int T, float scale, cudaStream_t s) {
  launch_k(transpose_kernel, dim3(dim3((T + 31) / 32, (C + 31) / 32, B)), dim3(256), 0, s, x, tok, C, T, scale);
  ST_CHECK_LAUNCH();
  return ST_OK;
}
int transpose_from_tokens(const float* tok, float* x, int B, int C, int T, cudaStream_t s) {
  launch_k(transpose_kernel, dim3(dim3((C + 31) / 32, (T + 31) / 32, B)), dim3(256), 0, s, tok, x, T, C, 1.0f);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 7. Conditioning helpers: word-embedding gather (Embedding + Linear folded into one table,
//    denoiser.py:152-153) and the 4-frame average pool (denoiser.py:157).
// =========================================================================================================
__global__ void __launch_bounds__(256) gather_words_kernel(const int32_t* __restrict__ word, const float* __restrict__ table,
                                                           float* __restrict__ out, int ldo, int rows, int force_zero, int n_words) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;    // rows x 64 float4
  if (gid >= (long long)rows * 64) return;
  const int r = (int)(gid >> 6), c4 = (int)(gid & 63);
  // nn.Embedding raises on an id outside the table (denoiser.py:152); the host mirror checks CPU tensors, and an id that
  // reaches the device anyway is clamped to <unk>-like row n_words - 1 / row 0 rather than read out of bounds
  const int w = force_zero ? 0 : min(max(word[r], 0), n_words - 1);
  const float4 v = __ldg(reinterpret_cast<const float4*>(table + (long long)w * 256) + c4);
  *reinterpret_cast<float4*>(out + (long long)r * ldo + c4 * 4) = v;
}
int gather_words(const int32_t* word, const float* table, float* out, int ldo, int rows, int force_zero, int n_words, cudaStream_t s) {
  const long long n = (long long)rows * 64;
  launch_k(gather_words_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, word, table, out, ldo, rows, force_zero, n_words);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

__global__ void __launch_bounds__(256) avgpool4_kernel(const float* __restrict__ in, float* __restrict__ out, long long n4,
                                                       int cols4) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= n4) return;
  const long long r = gid / cols4;
  const int c = (int)(gid - r * cols4);
  const float4* src = reinterpret_cast<const float4*>(in) + (r * 4) * cols4 + c;
  const float4 a = src[0], b = src[cols4], d = src[2 * (long long)cols4], e = src[3 * (long long)cols4];
  float4 o;
  o.x = (((a.x + b.x) + d.x) + e.x) * 0.25f;
  o.y = (((a.y + b.y) + d.y) + e.y) * 0.25f;
  o.z = (((a.z + b.z) + d.z) + e.z) * 0.25f;
  o.w = (((a.w + b.w) + d.w) + e.w) * 0.25f;
  reinterpret_cast<float4*>(out)[gid] = o;
}
int avgpool4(const float* in, float* out, int rows_out, int cols, cudaStream_t s) {
  const long long n4 = (long long)rows_out * cols / 4;
  launch_k(avgpool4_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, in, out, n4, cols / 4);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

__global__ void __launch_bounds__(256) copy_strided_scale_kernel(const float* __restrict__ in, long long in_stride, float scale,
                                                                 float* __restrict__ out, long long n4, int cols4) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= n4) return;
  const long long r = gid / cols4;
  const int c = (int)(gid - r * cols4);
  float4 v = *(reinterpret_cast<const float4*>(in + r * in_stride) + c);
  v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
  reinterpret_cast<float4*>(out)[gid] = v;
}
int copy_strided_scale(const float* in, long long in_stride, float scale, float* out, int rows, int cols, cudaStream_t s) {
  const long long n4 = (long long)rows * cols / 4;
  launch_k(copy_strided_scale_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, in, in_stride, scale, out, n4, cols / 4);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 8. RVQ code selection for one quantizer layer (quantizer.py:67-80,132-158; residual_vq.py:143-147).
//    dot[r,j] = residual_r . c_j comes from the GEMM engine.  dist = (|r|^2 - 2 dot) + |c_j|^2 in fp32 like
//    the reference, argmin with first-index ties (= argmax(-dist)), then x_q = r + (c - r),
//    residual -= x_q, qsum += x_q.  One warp per row.
// =========================================================================================================
__global__ void __launch_bounds__(256) vq_select_kernel(const float* __restrict__ dot, const float* __restrict__ cnorm,
                                                        const float* __restrict__ codebook, float* __restrict__ residual,
                                                        float* __restrict__ qsum, int64_t* __restrict__ idx, int idx_stride,
                                                        int rows, int first, __half* __restrict__ r_planes, __half* __restrict__ q_planes) {
  pdl_wait();
  trace_stamp(9);
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* rr = reinterpret_cast<float4*>(residual + (long long)row * 512);
  float4 r[4];
  float x2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r[i] = rr[lane + 32 * i];
    x2 += (r[i].x * r[i].x + r[i].y * r[i].y) + (r[i].z * r[i].z + r[i].w * r[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x2 += __shfl_xor_sync(0xffffffffu, x2, o);
  const float* dr = dot + (long long)row * 512;
  float best = INFINITY;
  int bi = 0x7fffffff;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int j = lane + 32 * i;
    const float d = __fadd_rn(__fsub_rn(x2, __fmul_rn(2.0f, dr[j])), __ldg(cnorm + j));
    if (d < best) { best = d; bi = j; }        // ascending j per lane: strict < keeps the first index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0 && idx) idx[(long long)row * idx_stride] = bi;
  const float4* cb = reinterpret_cast<const float4*>(codebook + (long long)bi * 512);
  float4* qs = reinterpret_cast<float4*>(qsum + (long long)row * 512);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 c = __ldg(cb + lane + 32 * i);
    float4 xq, nr, acc;
    xq.x = __fadd_rn(r[i].x, __fsub_rn(c.x, r[i].x));
    xq.y = __fadd_rn(r[i].y, __fsub_rn(c.y, r[i].y));
    xq.z = __fadd_rn(r[i].z, __fsub_rn(c.z, r[i].z));
    xq.w = __fadd_rn(r[i].w, __fsub_rn(c.w, r[i].w));
    nr.x = __fsub_rn(r[i].x, xq.x); nr.y = __fsub_rn(r[i].y, xq.y);
    nr.z = __fsub_rn(r[i].z, xq.z); nr.w = __fsub_rn(r[i].w, xq.w);
    rr[lane + 32 * i] = nr;
    // tcgen05 engine: the next layer's ranking GEMM streams the new residual as fp16 hi/lo planes -- written here, no split pass
    if (r_planes) store_planes4(r_planes + (long long)row * 512 + (lane + 32 * i) * 4, (long long)rows * 512, nr.x, nr.y, nr.z, nr.w);
    if (first) acc = make_float4(0.f + xq.x, 0.f + xq.y, 0.f + xq.z, 0.f + xq.w);
    else {
      const float4 old = qs[lane + 32 * i];
      acc = make_float4(__fadd_rn(old.x, xq.x), __fadd_rn(old.y, xq.y), __fadd_rn(old.z, xq.z), __fadd_rn(old.w, xq.w));
    }
    qs[lane + 32 * i] = acc;
    if (q_planes) store_planes4(q_planes + (long long)row * 512 + (lane + 32 * i) * 4, (long long)rows * 512, acc.x, acc.y, acc.z, acc.w);
  }
}

int vq_select(const float* dot, const float* cnorm, const float* codebook, float* residual, float* qsum, int64_t* idx,
              int idx_stride, int rows, int first, __half* r_planes, __half* q_planes, cudaStream_t s) {
  launch_k(vq_select_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, dot, cnorm, codebook, residual, qsum, idx, idx_stride, rows, first,
           r_planes, q_planes);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// =========================================================================================================
// 9. 330-d assembly (diffusion_rvqvae_trainer.py:484-531; utils/rotation_conversions.py:39-64,96-118,
//    432-508,511-550).  One thread per (b, frame, joint): de-normalise the 6d, Gram-Schmidt, matrix ->
//    quaternion -> axis-angle -> quaternion -> matrix, keep the first two rows.  Jaw (joint 22) comes from
//    the caller, eyes (23, 24) are zero rotations.  Root translation: v*std+mean, cumsum over frames for
//    x/z, y kept as is.
// =========================================================================================================
__constant__ int c_upper_j[13] = {3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21};
__constant__ int c_lower_j[9] = {0, 1, 2, 4, 5, 7, 8, 10, 11};

__device__ __forceinline__ float sqrt_pos(float x) { return x > 0.f ? sqrtf(x) : 0.f; }
__device__ __forceinline__ float copysign_ref(float a, float b) { return ((a < 0.f) != (b < 0.f)) ? -a : a; }
__device__ __forceinline__ float half_sinc(float angle, float half) {
  return fabsf(angle) < 1e-6f ? 0.5f - angle * angle / 48.0f : sinf(half) / angle;
}

__device__ void aa_to_6d(const float aa[3], float out[6]) {
  const float angle = sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  const float half = 0.5f * angle;
  const float s = half_sinc(angle, half);
  const float r = cosf(half), i = aa[0] * s, j = aa[1] * s, k = aa[2] * s;
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  out[0] = 1.0f - two_s * (j * j + k * k);
  out[1] = two_s * (i * j - k * r);
  out[2] = two_s * (i * k + j * r);
  out[3] = two_s * (i * j + k * r);
  out[4] = 1.0f - two_s * (i * i + k * k);
  out[5] = two_s * (j * k - i * r);
}

__device__ void sixd_to_aa(const float d6[6], float aa[3]);
__device__ void sixd_roundtrip(const float d6[6], float out[6]) {
  float aa[3];
  sixd_to_aa(d6, aa);
  aa_to_6d(aa, out);
}
// rotation_6d_to_matrix -> matrix_to_quaternion -> quaternion_to_axis_angle (utils/rotation_conversions.py:511-532, 96-118, 470-508)
__device__ void sixd_to_aa(const float d6[6], float aa[3]) {
  // rotation_6d_to_matrix
  const float n1 = fmaxf(sqrtf(d6[0] * d6[0] + d6[1] * d6[1] + d6[2] * d6[2]), 1e-12f);
  const float b1[3] = {d6[0] / n1, d6[1] / n1, d6[2] / n1};
  const float dp = b1[0] * d6[3] + b1[1] * d6[4] + b1[2] * d6[5];
  float b2[3] = {d6[3] - dp * b1[0], d6[4] - dp * b1[1], d6[5] - dp * b1[2]};
  const float n2 = fmaxf(sqrtf(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]), 1e-12f);
  b2[0] /= n2; b2[1] /= n2; b2[2] /= n2;
  const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
  // matrix_to_quaternion (rows b1,b2,b3)
  const float m00 = b1[0], m11 = b2[1], m22 = b3[2];
  const float o0 = 0.5f * sqrt_pos(1.f + m00 + m11 + m22);
  const float qx = 0.5f * sqrt_pos(1.f + m00 - m11 - m22);
  const float qy = 0.5f * sqrt_pos(1.f - m00 + m11 - m22);
  const float qz = 0.5f * sqrt_pos(1.f - m00 - m11 + m22);
  const float o1 = copysign_ref(qx, b3[1] - b2[2]);
  const float o2 = copysign_ref(qy, b1[2] - b3[0]);
  const float o3 = copysign_ref(qz, b2[0] - b1[1]);
  // quaternion_to_axis_angle
  const float nrm = sqrtf(o1 * o1 + o2 * o2 + o3 * o3);
  const float half = atan2f(nrm, o0);
  const float angle = 2.0f * half;
  const float s = half_sinc(angle, half);
  aa[0] = o1 / s; aa[1] = o2 / s; aa[2] = o3 / s;
}

__global__ void __launch_bounds__(256) pose330_kernel(const float* __restrict__ up, const float* __restrict__ ha,
                                                      const float* __restrict__ lo, const float* __restrict__ mean,
                                                      const float* __restrict__ std, const float* __restrict__ jaw, int BN_,
                                                      float* __restrict__ pose) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= (long long)BN_ * 55) return;
  const int f = (int)(gid / 55), j = (int)(gid - (long long)f * 55);
  float out[6];
  const float* src = nullptr;
  if (j >= 25) src = ha + (long long)f * 180 + (j - 25) * 6;
  else if (j == 22) {
    float aa[3] = {0.f, 0.f, 0.f};
    if (jaw) { aa[0] = jaw[(long long)f * 3]; aa[1] = jaw[(long long)f * 3 + 1]; aa[2] = jaw[(long long)f * 3 + 2]; }
    aa_to_6d(aa, out);
  } else if (j == 23 || j == 24) {
    const float aa[3] = {0.f, 0.f, 0.f};
    aa_to_6d(aa, out);
  } else {
    for (int q = 0; q < 13; ++q) if (c_upper_j[q] == j) src = up + (long long)f * 78 + q * 6;
    for (int q = 0; q < 9; ++q) if (c_lower_j[q] == j) src = lo + (long long)f * 57 + q * 6;
  }
  if (src) {
    float d6[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) d6[q] = __fadd_rn(__fmul_rn(src[q], std[j * 6 + q]), mean[j * 6 + q]);
    sixd_roundtrip(d6, out);
  }
  float* o = pose + (long long)f * 330 + j * 6;
#pragma unroll
  for (int q = 0; q < 6; ++q) o[q] = out[q];
}

__global__ void trans_kernel(const float* __restrict__ lo, const float* __restrict__ tmean, const float* __restrict__ tstd,
                             int B, int n, float* __restrict__ trans) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * 3) return;
  const int b = gid / 3, c = gid - b * 3;
  float run = 0.f;
  for (int f = 0; f < n; ++f) {
    const float v = __fadd_rn(__fmul_rn(lo[((long long)b * n + f) * 57 + 54 + c], tstd[c]), tmean[c]);
    run = (f == 0) ? v : __fadd_rn(run, v);
    trans[((long long)b * n + f) * 3 + c] = (c == 1) ? v : run;
  }
}

int pose330(const float* up, const float* ha, const float* lo, const float* mean, const float* std, const float* tmean,
            const float* tstd, const float* jaw, int B, int n, float* pose, float* trans, cudaStream_t s) {
  const long long tot = (long long)B * n * 55;
  launch_k(pose330_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, s, up, ha, lo, mean, std, jaw, B * n, pose);
  ST_CHECK_LAUNCH();
  if (trans) {
    launch_k(trans_kernel, dim3((B * 3 + 127) / 128), dim3(128), 0, s, lo, tmean, tstd, B, n, trans);
    ST_CHECK_LAUNCH();
  }
  return ST_OK;
}

// =========================================================================================================
// 9b. Evaluation tail (the part that needs no SMPL-X): the 330-d features as the 165-d axis-angle vector of the
//     result files (diffusion_rvqvae_trainer.py:621-622, 702-710), and the sufficient statistics of the two metrics that reduce over
//     ranks -- first / second moments of the FID latents (dataloaders/data_tools.py:1615-1625: np.mean, np.cov) and the L1 diversity
//     sums (utils/metric.py:12-27) -- accumulated in float64 so that per-rank partial sums add up exactly enough for an all-reduce.
// =========================================================================================================
__global__ void __launch_bounds__(256) pose_aa165_kernel(const float* __restrict__ pose, long long nj, float* __restrict__ aa) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;     // (frame, joint)
  if (gid >= nj) return;
  float d6[6], a[3];
#pragma unroll
  for (int q = 0; q < 6; ++q) d6[q] = pose[gid * 6 + q];
  sixd_to_aa(d6, a);
  aa[gid * 3] = a[0]; aa[gid * 3 + 1] = a[1]; aa[gid * 3 + 2] = a[2];
}
int pose_aa165(const float* pose, long long frames, float* aa, cudaStream_t s) {
  const long long nj = frames * 55;
  launch_k(pose_aa165_kernel, dim3((unsigned)((nj + 255) / 256)), dim3(256), 0, s, pose, nj, aa);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// acc = [n | s1[D] | s2[D][D]] (float64): n += N, s1[j] += sum_i x[i][j], s2[j][k] += sum_i x[i][j] x[i][k].  Block (j, k-tile of 64):
// every (j, k) has exactly one writer and a fixed summation order, so the result is deterministic.
__global__ void __launch_bounds__(64) moments_kernel(const float* __restrict__ x, long long N, int D, double* __restrict__ acc) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  const int j = blockIdx.x, k = blockIdx.y * 64 + threadIdx.x;
  if (k >= D) return;
  double s2 = 0.0, s1 = 0.0;
  for (long long i = 0; i < N; ++i) {
    const double xj = (double)x[i * D + j], xk = (double)x[i * D + k];
    s2 = fma(xj, xk, s2);
    s1 += xk;
  }
  acc[1 + D + (long long)j * D + k] += s2;
  if (j == 0) acc[1 + k] += s1;
  if (j == 0 && k == 0) acc[0] += (double)N;
}
int moments_accumulate(const float* x, long long N, int D, double* acc, cudaStream_t s) {
  launch_k(moments_kernel, dim3(D, (D + 63) / 64), dim3(64), 0, s, x, N, D, acc);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// L1div.run(results [n,J]) (utils/metric.py:16-22): counter += n; sum += sum_ij |x_ij - mean_j|.  acc = [sum, counter] (float64).
__global__ void __launch_bounds__(256) l1div_kernel(const float* __restrict__ x, int n, int J, double* __restrict__ acc) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  __shared__ double part[256];
  double tot = 0.0;
  for (int j = threadIdx.x; j < J; j += 256) {
    double m = 0.0;
    for (int i = 0; i < n; ++i) m += (double)x[(long long)i * J + j];
    m /= (double)n;
    const float mf = (float)m;                                     // the reference's mean is a float32 array (np.mean of float32)
    for (int i = 0; i < n; ++i) tot += (double)fabsf(x[(long long)i * J + j] - mf);
  }
  part[threadIdx.x] = tot;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { acc[0] += part[0]; acc[1] += (double)n; }
}
int l1div_accumulate(const float* x, int n, int J, double* acc, cudaStream_t s) {
  launch_k(l1div_kernel, dim3(1), dim3(256), 0, s, x, n, J, acc);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// h3d 623-d scatter (h3d_diffusion_new_trainer.py:194-221,604-607): column c of the output belongs to one body
// part at a fixed position; enumerate the masks exactly like the reference builds them.
__device__ int h3d_lookup(int c, int* part) {
  // returns position inside the part's decoder output, sets *part (0 upper, 1 hands, 2 lower), or -1 (stays zero)
  const int uj[13] = {3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21};
  const int lj[9] = {0, 1, 2, 4, 5, 7, 8, 10, 11};
  int pos = 0;
  for (int q = 0; q < 13; ++q) {
    const int i = uj[q];
    for (int k = 0; k < 3; ++k, ++pos) if (c == 4 + (i - 1) * 3 + k) { *part = 0; return pos; }
    for (int k = 0; k < 6; ++k, ++pos) if (c == 4 + 51 * 3 + (i - 1) * 6 + k) { *part = 0; return pos; }
    for (int k = 0; k < 3; ++k, ++pos) if (c == 4 + 51 * 9 + i * 3 + k) { *part = 0; return pos; }
  }
  pos = 0;
  for (int i = 22; i < 52; ++i) {
    for (int k = 0; k < 3; ++k, ++pos) if (c == 4 + (i - 1) * 3 + k) { *part = 1; return pos; }
    for (int k = 0; k < 6; ++k, ++pos) if (c == 4 + 51 * 3 + (i - 1) * 6 + k) { *part = 1; return pos; }
    for (int k = 0; k < 3; ++k, ++pos) if (c == 4 + 51 * 9 + i * 3 + k) { *part = 1; return pos; }
  }
  pos = 0;
  for (int k = 0; k < 4; ++k, ++pos) if (c == k) { *part = 2; return pos; }
  for (int k = 619; k < 623; ++k, ++pos) if (c == k) { *part = 2; return pos; }
  for (int q = 0; q < 9; ++q) {
    const int i = lj[q];
    if (i > 0) {
      for (int k = 0; k < 3; ++k, ++pos) if (c == 4 + (i - 1) * 3 + k) { *part = 2; return pos; }
      for (int k = 0; k < 6; ++k, ++pos) if (c == 4 + 51 * 3 + (i - 1) * 6 + k) { *part = 2; return pos; }
    }
    for (int k = 0; k < 3; ++k, ++pos) if (c == 4 + 51 * 9 + i * 3 + k) { *part = 2; return pos; }
  }
  return -1;
}

__global__ void __launch_bounds__(256) pose623_kernel(const float* __restrict__ up, const float* __restrict__ ha,
                                                      const float* __restrict__ lo, int frames, float* __restrict__ pose) {
  pdl_wait();
  trace_stamp(10);
  pdl_launch();
  __shared__ int s_part[623], s_pos[623];
  for (int c = threadIdx.x; c < 623; c += 256) {
    int part = -1;
    s_pos[c] = h3d_lookup(c, &part);
    s_part[c] = part;
  }
  __syncthreads();
  for (int f = blockIdx.x; f < frames; f += gridDim.x) {
    for (int c = threadIdx.x; c < 623; c += 256) {
      float v = 0.f;
      const int pos = s_pos[c];
      if (pos >= 0) {
        const int part = s_part[c];
        v = part == 0 ? up[(long long)f * 156 + pos] : part == 1 ? ha[(long long)f * 360 + pos] : lo[(long long)f * 107 + pos];
      }
      pose[(long long)f * 623 + c] = v;
    }
  }
}

int pose623(const float* up, const float* ha, const float* lo, int B, int n, float* pose, cudaStream_t s) {
  const int frames = B * n;
  launch_k(pose623_kernel, dim3(frames < 1184 ? frames : 1184), dim3(256), 0, s, up, ha, lo, frames, pose);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int set_trace_kernels(unsigned long long* p) {
  ST_CHECK_CUDA(cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)));
  return ST_OK;
}

}  // namespace st
