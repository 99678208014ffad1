// tcgen05 GEMM engine for sm_100a: fp32-class accuracy from fp16 tensor cores by operand splitting.
//
//   a = a_hi + a_lo,  w = w_hi + w_lo   (fp16 pairs of the pre-scaled fp32 values, 22 significand bits)
//   a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi     (three tcgen05.mma per K step, fp32 accumulation in TMEM;
//                                                  the dropped a_lo.w_lo term is 2^-24 relative)
//
// Data path per CTA (one 128 x BN output tile): a single elected thread streams K-blocks of the four operand
// planes with TMA (cp.async.bulk.tensor, 128B swizzle, both planes of an operand in one 3-D box) into a
// STAGES-deep shared-memory ring; a second elected thread issues the MMAs from shared-memory descriptors into
// a TMEM accumulator and releases ring slots with tcgen05.commit; four epilogue warps read the accumulator
// back with tcgen05.ld (lane = output row), apply scale/bias/activation/residual and store fp32 -- and, when
// asked, the fp16 hi/lo planes of the result so that the next GEMM can TMA them directly.
#include "st_internal.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>

namespace st {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  ptxas recognises elect.sync and issues the uniform-datapath instructions of the elected
// branch (UTCHMMA, UTMALDG, UTCBAR) straight; under a plain `lane == 0` test it wraps each of them in an
// ELECT / BRA.U.ANY loop (measured: 99 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p mov.u32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must end in a trap (launch failure), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start address >> 4 | LBO (ignored for swizzled K-major, 1) << 16 | SBO = 1024 B between 8-row groups << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// A K-major tile whose first row is NOT the first row of a 1024-byte swizzle atom (a conv tap reads the staged rows shifted by
// tap * dilation) takes the same descriptor with the shifted start address and the base-offset field left 0: the hardware derives the
// swizzle phase of a row from the absolute shared-memory address, like the TMA engine that wrote it.  Measured (tests/conv_probe.py):
// integer-valued operands reproduce the per-tap-fetch result bit for bit at every shift; setting bits [49,52) to (address >> 7) & 7,
// as the PTX ISA text on the matrix descriptor suggests for unaligned starts, reads the wrong rows.
// The same for an MN-major operand (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>, canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): rows of 64 contiguous MN elements (128 B, swizzled in groups of 8 rows),
// one row per K index; LBO = bytes between 64-element groups along MN, SBO = bytes between groups of 8 K rows.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t kIdescBMajorMN = 1u << 16;      // InstrDescriptor b_major_: B operand is MN-major
// Instruction descriptor (InstrDescriptor): D = F32 (1 << 4), A = B = F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// ---------------------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM lane group)
constexpr int TC_A_PLANE = TC_BM * TC_BK * 2;   // bytes of one A plane tile (128 rows x 128 B)
constexpr int TAPS_ROWS = 152;                  // staged rows of a conv tile: 128 outputs + (taps - 1) * dilation <= 24 of halo
constexpr int TAPS_SLOT = 2 * TAPS_ROWS * 128;  // hi + lo plane of one channel block, 38 912 B (a multiple of the 1024 B swizzle atom per plane)

// ---- MMA issue loop of both tcgen05 kernels -------------------------------------------------------------------
// The tensor pipe runs at most two MMAs behind the issuing thread (tests/umma_probe.cu: a delay of d cycles between two K
// blocks costs d - 30 cycles at BN = 64, d - 215 at BN = 128), so whatever that thread does between the last MMA of a K block
// and the first of the next is exposed.  As first written (blocking barrier wait, descriptors rebuilt from the stage index on
// the vector datapath and moved to uniform registers with a dozen R2URs, probe / trace tests against constant memory: ~275
// cycles) a 64-wide K block took 694 cycles against 454 of tensor time, a 128-wide one 833 against 776.  Here the loop is
// unrolled over the ring (S compile-time), so every descriptor is loop-invariant and lives in a uniform register, and the NEXT
// stage's barrier is tested (non-blocking) between the third and the fourth k-step, its answer read after the fourth: in the
// steady state the thread goes from the last MMA of a block straight to the commit and the next block's first MMA.
template <bool CAT>
__device__ __forceinline__ void tc_issue_loop(uint64_t* full_bar, uint64_t* empty_bar, const uint32_t smem_base, const int stage_bytes,
                                              const int w_plane, const int BN, const int STAGES, const uint32_t tmem_base, const int num_kb,
                                              long long* dbg) {
  constexpr uint64_t kDescHi = (64ull << 32) | (1ull << 46) | (2ull << 61);
  uint32_t idesc_w = umma_idesc_f16(TC_BM, CAT ? 2 * BN : BN), idesc_n = umma_idesc_f16(TC_BM, BN);
  uint32_t d_base = ((smem_base >> 4) & 0x3FFF) | (1u << 16);                  // low word of umma_desc_sw128(stage 0)
  uint32_t d_stage = (uint32_t)stage_bytes >> 4, d_wlo = (uint32_t)w_plane >> 4;
  uint32_t tm_main = tmem_base, tm_corr = tmem_base + BN;
  uint32_t fb0 = smem_u32(full_bar), eb0 = smem_u32(empty_bar);
  int stages = STAGES, nkb = num_kb;
  // everything the loop needs is formed (and pinned: the empty asm statements are definition points) BEFORE the wait for the first
  // operands, i.e. while the thread has nothing else to do; kernel parameters would otherwise be re-read from constant memory
  // inside the loop
  asm volatile("" : "+r"(idesc_w), "+r"(idesc_n), "+r"(d_base), "+r"(d_stage), "+r"(d_wlo), "+r"(tm_main), "+r"(tm_corr), "+r"(fb0), "+r"(eb0),
               "+r"(stages), "+r"(nkb));
  mbar_wait(full_bar, 0);
  tc_fence_after();
  trace_stamp(21, true);                             // first operands landed
  uint32_t lo = d_base, ph = 0, eb = eb0;
  int s = 0;
  for (int kb = 0; kb < nkb; ++kb) {
    if (dbg && kb < 16) dbg[24 + kb] = clock64();
    // the k-step advance (+2 per 32 bytes) is applied to the low word: the high word stays a compile-time constant
    const uint32_t ah = lo, al = ah + (TC_A_PLANE >> 4), wh = ah + (2 * TC_A_PLANE >> 4), wl = wh + d_wlo;
    const uint32_t acc0 = kb != 0;
    auto kstep = [&](int k) {
      const uint64_t dah = kDescHi | (uint64_t)(ah + 2 * k), dal = kDescHi | (uint64_t)(al + 2 * k);
      const uint64_t dwh = kDescHi | (uint64_t)(wh + 2 * k), dwl = kDescHi | (uint64_t)(wl + 2 * k);
      if (CAT) {
        // columns [0, BN) = a_hi.w_hi (main), [BN, 2BN) = a_hi.w_lo; then a_lo.w_hi joins the second half
        umma_f16(tm_main, dah, dwh, idesc_w, k ? 1u : acc0);
        umma_f16(tm_corr, dal, dwh, idesc_n, 1);
      } else {
        // hi.hi goes to the main accumulator; the two 2^-11-sized cross terms go to their own accumulator so that the
        // tensor core's truncating fp32 adds see a 3x shorter chain on the large sum (measured: error / 3)
        umma_f16(tm_main, dah, dwh, idesc_n, k ? 1u : acc0);
        umma_f16(tm_corr, dah, dwl, idesc_n, k ? 1u : acc0);
        umma_f16(tm_corr, dal, dwh, idesc_n, 1);
      }
    };
    kstep(0);
    kstep(1);
    // the next block's stage, barrier addresses and descriptor base, formed while this block's MMAs are queued
    const bool wrap = s + 1 == stages;
    int sn = wrap ? 0 : s + 1;
    uint32_t phn = wrap ? ph ^ 1 : ph;
    uint32_t lo_n = wrap ? d_base : lo + d_stage;
    uint32_t fbn = fb0 + 8 * sn, ebn = eb0 + 8 * sn;
    asm volatile("" : "+r"(sn), "+r"(phn), "+r"(lo_n), "+r"(fbn), "+r"(ebn));
    const bool more = kb + 1 < nkb;
    kstep(2);
    if (more) asm volatile("mbarrier.test_wait.parity.shared::cta.b64 st_pfull, [%0], %1;" ::"r"(fbn), "r"(phn) : "memory");
    kstep(3);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(eb) : "memory");
    if (more) {
      uint32_t ok;
      asm volatile("selp.u32 %0, 1, 0, st_pfull;" : "=r"(ok));
      if (!ok) mbar_wait(full_bar + sn, phn);
      tc_fence_after();
    }
    s = sn; ph = phn; lo = lo_n; eb = ebn;
  }
}

struct TcEpi {
  float* out;            // [M, ldo] fp32 or null
  const float* bias;     // [N] or null
  const float* res;      // [M/res_div, ldr] or null
  __half* planes;        // optional fp16 hi/lo planes of the result: [2][M][ld_planes]
  long long plane_stride;
  int ld_planes;
  float planes_scale;    // result * planes_scale is what gets split
  int planes_relu;
  int ldo, ldr, res_mode, res_div, act;
  float scale;           // 2^-(sa+sw): undo the operand pre-scaling
  int M, N;
  // operand addressing of A:
  //  mode 0  plain rows            3-D map {K, rows, 2}                coords (kb*64, m0, 0)
  //  mode 1  conv, clips packed    4-D map {C, T, clips, 2}, T | 128   coords (cb*64, j*dil - pad, m0/T, 0)
  //  mode 2  conv, long clips      4-D map {C, Lin, B, 2}              coords (cb*64, t0 + j*dil - pad, b, 0)      tile -> (b, t0)
  //  mode 3  strided conv, pad 0   3-D map {s*C, ceil(B*Lin/s), 2}     coords ((G%s)*C + cb*64, G/s, 0), G = b*Lin + t0*s + j
  //  mode 4  conv, long clips, rows staged ONCE per channel block (4-D map {C, Lin, B, 2}, box {64, 152, 1, 2} at t0 - pad): the
  //          MMAs of tap j read them through a descriptor shifted by j*dil rows; only the weight tiles stream per (tap, block)
  int mode, T, kb_per_tap, dil, pad, stride, C, Lin, Lout, tpc;
  int a_slots;           // mode 4: staged row slots (1 or 2)
  const float* ln_stats; // consumer: [rows][8][2] partial (mean, M2) of each 512-wide input row, or null
  const float* ln_s;     //           [N] column sums of the gamma-scaled weight
  const float* ln_c;     //           [N] W beta + bias
  float* stats_out;      // producer (N = 512, BN = 64): [rows][8][2], this CTA writes slot blockIdx.x
  long long* dbg;        // optional timeline of CTA (0,0): clock64 stamps (debug / profiling only)
  int probe;             // debug: 1 skip MMAs, 2 skip TMA, 4 MMAs grouped by accumulator, 8 single accumulator (timing only)
};

__device__ __forceinline__ float tc_act(float v, int act) {
  switch (act) {
    case ACT_GELU: return v * 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_LRELU: return v > 0.0f ? v : v * 0.01f;
    default: return v;
  }
}

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v, -65000.0f), 65000.0f);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// (a, b) -> packed fp16 pairs hi = rn(a, b) (saturating to the largest finite fp16), lo = rn((a, b) - hi): six instructions per
// pair (F2FP.SATFINITE.PACK_AB, 2 HADD2.F32, 2 FADD, F2FP.PACK_AB) where the clamp + scalar form took ten.
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));       // low half <- a, high half <- b
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - f.y), "f"(a - f.x));
}
// eight consecutive values (times s) -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void split8_f16(const float* x, float s, uint4& h, uint4& l) {
  split2_f16(x[0] * s, x[1] * s, h.x, l.x);
  split2_f16(x[2] * s, x[3] * s, h.y, l.y);
  split2_f16(x[4] * s, x[5] * s, h.z, l.z);
  split2_f16(x[6] * s, x[7] * s, h.w, l.w);
}

// One binary for every tile width (BN = 64 / 128 / 192 and the ring depth are run-time values): the sampling loop
// alternates GEMMs of different widths back to back, and separate template instantiations evicted each other from
// the instruction cache at every launch (the in-chain cost of a GEMM was ~4 us above its same-kernel-chain cost).
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcEpi ep, const int num_kb,
               const int BN, const int STAGES) {
  const int W_PLANE = BN * TC_BK * 2;
  const int STAGE_BYTES = 2 * TC_A_PLANE + 2 * W_PLANE;
  const int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;   // main + correction accumulators
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // mode 4 (staged taps): [2 row slots of 2 planes x 152 rows][STAGES weight slots]; otherwise [STAGES slots of A | W]
  const bool taps_mode = ep.mode == 4;
  const int NSLOT = ep.a_slots;                  // 1 when the conv has a single channel block (C = 64), else 2
  const int ring_bytes = taps_mode ? NSLOT * TAPS_SLOT + STAGES * 2 * W_PLANE : STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ring_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint64_t* afull_bar = acc_bar + 1;            // mode 4: row slot filled / free
  uint64_t* aempty_bar = afull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;                    // n fastest: CTAs that share an A tile run together
  // output rows of this tile: [row_base, row_base + rows_valid)
  int row_base, rows_valid, clip_b = 0, t0 = 0;
  if (ep.mode >= 2) {
    clip_b = blockIdx.y / ep.tpc;
    t0 = (blockIdx.y - clip_b * ep.tpc) * TC_BM;
    row_base = clip_b * ep.Lout + t0;
    rows_valid = min(TC_BM, ep.Lout - t0);
  } else {
    row_base = blockIdx.y * TC_BM;
    rows_valid = min(TC_BM, ep.M - row_base);
  }
  const int m0 = row_base;
  long long* dbg = (ep.dbg && blockIdx.x == 0 && blockIdx.y == 0) ? ep.dbg : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&afull_bar[i], 1); mbar_init(&aempty_bar[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // REDUX leaves its result in a uniform register: ptxas then issues the UTCHMMAs back to back instead of wrapping each
  // one in an ELECT / R2UR / BRA.U.ANY waterfall (measured: 99 cycles per MMA, 3x the N = 64 tensor floor)
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
  // barriers, TMEM and descriptor prefetch above overlap the predecessor's tail; nothing before this line touches
  // global memory
  pdl_wait();
  pdl_launch();
  trace_stamp(1);
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0 && taps_mode) {
    if (elect_one()) {
      // ===== TMA producer, staged taps: the rows [t0 - pad, t0 - pad + 152) of a channel block once, then one weight tile per tap =====
      const int taps = num_kb / ep.kb_per_tap;
      int n = 0;
      for (int cb = 0; cb < ep.kb_per_tap; ++cb) {
        const int as = cb % NSLOT;
        mbar_wait(&aempty_bar[as], ((cb / NSLOT) & 1) ^ 1);
        mbar_expect_tx(&afull_bar[as], TAPS_SLOT);
        tma_load_4d(smem + as * TAPS_SLOT, &tmA, &afull_bar[as], cb * TC_BK, t0 - ep.pad, clip_b, 0);
        for (int j = 0; j < taps; ++j, ++n) {
          const int s = n % STAGES;
          mbar_wait(&empty_bar[s], ((n / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], 2 * W_PLANE);
          tma_load_3d(smem + NSLOT * TAPS_SLOT + s * 2 * W_PLANE, &tmW, &full_bar[s], (j * ep.kb_per_tap + cb) * TC_BK, n0, 0);
        }
      }
    }
  } else if (warp == 1 && taps_mode) {
    if (elect_one()) {
      // ===== MMA issuer, staged taps =====
      const int taps = num_kb / ep.kb_per_tap;
      const uint32_t idesc = umma_idesc_f16(TC_BM, BN), idesc_w = umma_idesc_f16(TC_BM, 2 * BN);
      int n = 0;
      for (int cb = 0; cb < ep.kb_per_tap; ++cb) {
        const int as = cb % NSLOT;
        mbar_wait(&afull_bar[as], (cb / NSLOT) & 1);
        tc_fence_after();
        const uint32_t rows_hi = smem_u32(smem + as * TAPS_SLOT), rows_lo = rows_hi + TAPS_SLOT / 2;
        for (int j = 0; j < taps; ++j, ++n) {
          const int s = n % STAGES;
          mbar_wait(&full_bar[s], (n / STAGES) & 1);
          tc_fence_after();
          const uint32_t shift = (uint32_t)(j * ep.dil) * 128;                    // tap j: the same rows, j*dil positions further on
          const uint32_t w_hi = smem_u32(smem + NSLOT * TAPS_SLOT + s * 2 * W_PLANE), w_lo = w_hi + W_PLANE;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t dah = umma_desc_sw128(rows_hi + shift) + 2 * k, dal = umma_desc_sw128(rows_lo + shift) + 2 * k;
            const uint64_t dwh = umma_desc_sw128(w_hi) + 2 * k, dwl = umma_desc_sw128(w_lo) + 2 * k;
            const uint32_t acc = (n | k) != 0;
            if (BN <= 128) {
              umma_f16(tmem_base, dah, dwh, idesc_w, acc);                        // a_hi.[w_hi ; w_lo]
              umma_f16(tmem_base + BN, dal, dwh, idesc, 1);                        // + a_lo.w_hi
            } else {
              umma_f16(tmem_base, dah, dwh, idesc, acc);
              umma_f16(tmem_base + BN, dah, dwl, idesc, acc);
              umma_f16(tmem_base + BN, dal, dwh, idesc, 1);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&aempty_bar[as]);
      }
      umma_commit(acc_bar);
    }
  } else if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        if (ep.probe & 2) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
          continue;
        }
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        if (ep.mode == 0) {
          tma_load_3d(st, &tmA, &full_bar[s], kb * TC_BK, m0, 0);
        } else {
          // implicit conv: K block kb = (tap j, channel block cb); out-of-range positions are zero-filled by TMA,
          // which is exactly the conv's zero padding
          const int j = kb / ep.kb_per_tap, cb = kb - j * ep.kb_per_tap;
          if (ep.mode == 1) tma_load_4d(st, &tmA, &full_bar[s], cb * TC_BK, j * ep.dil - ep.pad, m0 / ep.T, 0);
          else if (ep.mode == 2) tma_load_4d(st, &tmA, &full_bar[s], cb * TC_BK, t0 + j * ep.dil - ep.pad, clip_b, 0);
          else {
            const long long G = (long long)clip_b * ep.Lin + (long long)t0 * ep.stride + j;
            const int gr = (int)(G / ep.stride), gp = (int)(G - (long long)gr * ep.stride);
            tma_load_3d(st, &tmA, &full_bar[s], gp * ep.C + cb * TC_BK, gr, 0);
          }
        }
        tma_load_3d(st + 2 * TC_A_PLANE, &tmW, &full_bar[s], kb * TC_BK, n0, 0);
        if (dbg && kb < 16) dbg[8 + kb] = clock64();
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      if (!(ep.probe & 13)) {
        asm volatile(".reg .pred st_pfull;");
        if (BN <= 128) tc_issue_loop<true>(full_bar, empty_bar, smem_u32(smem), STAGE_BYTES, W_PLANE, BN, STAGES, tmem_base, num_kb, dbg);
        else tc_issue_loop<false>(full_bar, empty_bar, smem_u32(smem), STAGE_BYTES, W_PLANE, BN, STAGES, tmem_base, num_kb, dbg);
      }
      const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
      for (int kb = 0; kb < ((ep.probe & 13) ? num_kb : 0); ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (dbg && kb < 16) dbg[24 + kb] = clock64();
        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t a_lo = a_hi + TC_A_PLANE;
        const uint32_t w_hi = a_hi + 2 * TC_A_PLANE;
        const uint32_t w_lo = w_hi + W_PLANE;
        if (ep.probe & 13) {
          if (!(ep.probe & 1)) {
            const uint32_t corr = (ep.probe & 8) ? tmem_base : tmem_base + BN;
            if (ep.probe & 4) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                umma_f16(tmem_base, umma_desc_sw128(a_hi) + 2 * k, umma_desc_sw128(w_hi) + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                umma_f16(corr, umma_desc_sw128(a_hi) + 2 * k, umma_desc_sw128(w_lo) + 2 * k, idesc, (kb | k) != 0);
                umma_f16(corr, umma_desc_sw128(a_lo) + 2 * k, umma_desc_sw128(w_hi) + 2 * k, idesc, 1);
              }
            } else {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                umma_f16(tmem_base, umma_desc_sw128(a_hi) + 2 * k, umma_desc_sw128(w_hi) + 2 * k, idesc, (kb | k) != 0);
                umma_f16(corr, umma_desc_sw128(a_hi) + 2 * k, umma_desc_sw128(w_lo) + 2 * k, idesc, 1);
                umma_f16(corr, umma_desc_sw128(a_lo) + 2 * k, umma_desc_sw128(w_hi) + 2 * k, idesc, 1);
              }
            }
          }
          umma_commit(&empty_bar[s]);
          continue;
        }
        if (BN <= 128) {
          // the lo weight plane follows the hi plane in the stage: one MMA of width 2 BN forms a_hi.[w_hi ; w_lo] (main
          // accumulator + first cross term), a second of width BN adds a_lo.w_hi (see the trunk kernel)
          const uint32_t idesc_w = umma_idesc_f16(TC_BM, 2 * BN);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t dah = umma_desc_sw128(a_hi) + 2 * k, dal = umma_desc_sw128(a_lo) + 2 * k;
            const uint64_t dwh = umma_desc_sw128(w_hi) + 2 * k;
            umma_f16(tmem_base, dah, dwh, idesc_w, (kb | k) != 0);
            umma_f16(tmem_base + BN, dal, dwh, idesc, 1);
          }
        } else {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t dah = umma_desc_sw128(a_hi) + 2 * k, dal = umma_desc_sw128(a_lo) + 2 * k;
            const uint64_t dwh = umma_desc_sw128(w_hi) + 2 * k, dwl = umma_desc_sw128(w_lo) + 2 * k;
            // hi.hi goes to the main accumulator; the two 2^-11-sized cross terms go to their own accumulator so that
            // the tensor core's truncating fp32 adds see a 3x shorter chain on the large sum (measured: error / 3)
            umma_f16(tmem_base, dah, dwh, idesc, (kb | k) != 0);
            umma_f16(tmem_base + BN, dah, dwl, idesc, (kb | k) != 0);
            umma_f16(tmem_base + BN, dal, dwh, idesc, 1);
          }
        }
        umma_commit(&empty_bar[s]);     // slot reusable once these MMAs have read it
      }
      umma_commit(acc_bar);             // accumulator complete
      if (dbg) dbg[2] = clock64();
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane group = warp % 4, two warps per group (column halves, then row halves) =====
    // Phase 1: accumulators (main + correction) -> registers -> a padded fp32 tile in the now idle stage memory
    // (lane = row, so direct global stores would hit 32 different lines per instruction: measured 7000 cycles per
    // 32 columns).  Phase 2: lanes run along the columns, so bias / residual loads and fp32 / fp16-plane stores are
    // full-line coalesced; four rows are in flight per iteration to hide latency with one warp per scheduler.
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;             // 0: warps 2..5, 1: warps 6..9
    // While the main loop runs these warps are idle: fetch everything phase 2 needs from global memory now (bias,
    // the residual rows of this warp, kernel parameters), so that after the accumulator is ready only shared-memory
    // reads, arithmetic and stores remain.
    const float e_scale = ep.scale, e_pscale = ep.planes_scale;
    const int e_N = ep.N, e_ldo = ep.ldo, e_ldr = ep.ldr, e_ldp = ep.ld_planes, e_act = ep.act, e_prelu = ep.planes_relu;
    const int e_resmode = ep.res ? ep.res_mode : RES_NONE, e_resdiv = ep.res_div;
    const long long e_pstride = ep.plane_stride;
    float* const e_out = ep.out;
    const float* const e_res = ep.res;
    __half* const e_planes = ep.planes;
    const int NPIECE = BN == 192 ? 2 : 1;
    const int LPR0 = BN >= 128 ? 32 : 16;         // piece 0 geometry: lanes per row
    const int cl0 = (lane % LPR0) * 4, n_0 = n0 + cl0;
    const bool vec0 = (n_0 + 3 < e_N) && ((e_ldo & 3) == 0) && (!e_res || (e_ldr & 3) == 0);
    const float* const e_bias = ep.ln_stats ? ep.ln_c : ep.bias;      // folded LayerNorm: c = W beta + b replaces the bias
    const float* const e_lns = ep.ln_stats ? ep.ln_s : nullptr;
    float* const e_stats_out = ep.stats_out;
    float bias0[4] = {0.f, 0.f, 0.f, 0.f}, lns0[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n_0 + q < e_N) {
        if (e_bias) bias0[q] = __ldg(e_bias + n_0 + q);
        if (e_lns) lns0[q] = __ldg(e_lns + n_0 + q);
      }
    // folded LayerNorm, consumer side: this thread's phase-1 row needs (mean, 1/sigma) of its input row, combined
    // from the 8 partials (mean_j, M2_j) its producer left (Chan et al.: M2 = sum M2_j + n_j sum (mean_j - mean)^2)
    float ln_rstd = 1.0f, ln_mu = 0.0f;
    if (ep.ln_stats && (lg * 32 + lane) < rows_valid) {
      const float4* st4 = reinterpret_cast<const float4*>(ep.ln_stats + (long long)(m0 + lg * 32 + lane) * 16);
      const float4 a = st4[0], b4 = st4[1], c4 = st4[2], d4 = st4[3];
      const float mj[8] = {a.x, a.z, b4.x, b4.z, c4.x, c4.z, d4.x, d4.z};
      const float qj[8] = {a.y, a.w, b4.y, b4.w, c4.y, c4.w, d4.y, d4.w};
      float mu = 0.f, m2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) mu += mj[j];
      mu *= 0.125f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float dm = mj[j] - mu; m2 += qj[j] + 64.0f * dm * dm; }
      ln_mu = mu;
      ln_rstd = 1.0f / sqrtf(m2 * (1.0f / 512.0f) + 1e-5f);
    }
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    if (dbg && threadIdx.x == 64) dbg[3] = clock64();
    const int LDT = BN + 4;
    const uint32_t tile = smem_u32(smem) + (uint32_t)(lg * 32) * LDT * 4;      // shared-space byte address of this lane group's rows
    const int CH = BN / 32;                       // 32-column chunks; each warp of the pair takes CH/2 of them
#pragma unroll 1
    for (int cc = 0; cc < CH / 2; ++cc) {
      const int c = half * (CH / 2) + cc;
      uint32_t v[32], vc[32];
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + c * 32, v);
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + BN + c * 32, vc);
      tmem_ld_wait();
      const uint32_t trow = tile + (uint32_t)(lane * LDT + c * 32) * 4;
      const float sc = e_scale * ln_rstd;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 t;
        t.x = (__uint_as_float(v[j]) + __uint_as_float(vc[j])) * sc;
        t.y = (__uint_as_float(v[j + 1]) + __uint_as_float(vc[j + 1])) * sc;
        t.z = (__uint_as_float(v[j + 2]) + __uint_as_float(vc[j + 2])) * sc;
        t.w = (__uint_as_float(v[j + 3]) + __uint_as_float(vc[j + 3])) * sc;
        sts128(trow + j * 4, t);
      }
    }
    const uint32_t rowu = smem_u32(smem) + (uint32_t)(TC_BM * LDT) * 4;       // [128] rstd*mean per tile row, after the tile
    if (e_lns && half == 0) sts32(rowu + (uint32_t)(lg * 32 + lane) * 4, ln_rstd * ln_mu);
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
    asm volatile("bar.sync %0, 64;" ::"r"(1 + lg) : "memory");     // the two warps of this lane group
    if (dbg && threadIdx.x == 64) dbg[7] = clock64();
    // pieces of the tile row handled with 32 lanes x float4 (128 columns) or 16 lanes x float4 (64 columns).
    // The row loop is deliberately NOT unrolled: this code runs once per CTA, and an unrolled epilogue (40 KB of
    // SASS) spent ~2000 cycles per row-instruction in instruction-cache misses.  Loads run two rows ahead instead.
#pragma unroll 1
    for (int piece = 0; piece < NPIECE; ++piece) {
      const int pc0 = piece * 128;                                  // first column of the piece
      const int pw = (BN - pc0) >= 128 ? 128 : 64;                  // its width
      const int LPR = pw / 4;                                       // lanes per row
      const int RPI = 32 / LPR;                                     // rows per warp instruction
      const int iters = 16 / RPI;                                   // row-instructions for this warp's 16 rows
      const int cl = pc0 + (lane % LPR) * 4;
      const int n = n0 + cl;
      const int rsub = lane / LPR;
      const bool col_ok = n < e_N;
      const bool vec = (n + 3 < e_N) && ((e_ldo & 3) == 0) && (!e_res || (e_ldr & 3) == 0);
      const bool pvec = (n + 3 < e_N) && ((e_ldp & 3) == 0);
      float bias4[4] = {bias0[0], bias0[1], bias0[2], bias0[3]};
      float lns4[4] = {lns0[0], lns0[1], lns0[2], lns0[3]};
      if (piece > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          bias4[q] = (e_bias && n + q < e_N) ? __ldg(e_bias + n + q) : 0.f;
          lns4[q] = (e_lns && n + q < e_N) ? __ldg(e_lns + n + q) : 0.f;
        }
      }
      auto load_row = [&](int it, float4& t4, float4& r4) {
        const int r = half * 16 + it * RPI + rsub;
        t4 = lds128(tile + (uint32_t)(r * LDT + cl) * 4);
        r4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e_res && col_ok && (lg * 32 + r) < rows_valid) {
          const int grow = m0 + lg * 32 + r;
          const float* rrow = e_res + (long long)(e_resdiv == 1 ? grow : grow / e_resdiv) * e_ldr + n;
          if (vec) r4 = *reinterpret_cast<const float4*>(rrow);
          else {
            r4.x = rrow[0];
            if (n + 1 < e_N) r4.y = rrow[1];
            if (n + 2 < e_N) r4.z = rrow[2];
            if (n + 3 < e_N) r4.w = rrow[3];
          }
        }
      };
      float4 tA, rA, tB, rB, tC, rC;
      load_row(0, tA, rA);
      load_row(iters > 1 ? 1 : 0, tB, rB);
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {
        if (it + 2 < iters) load_row(it + 2, tC, rC);
        const int r = half * 16 + it * RPI + rsub;
        const bool valid = col_ok && (lg * 32 + r) < rows_valid;
        const int grow = m0 + lg * 32 + r;
        float x[4] = {tA.x + bias4[0], tA.y + bias4[1], tA.z + bias4[2], tA.w + bias4[3]};
        if (e_lns) {                                   // folded LayerNorm: tile holds rstd*acc, rowu holds rstd*mean
          const float u = lds32(rowu + (uint32_t)(lg * 32 + r) * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] -= u * lns4[q];
        }
        const float rs[4] = {rA.x, rA.y, rA.z, rA.w};
        if (e_resmode == RES_PRE) { x[0] += rs[0]; x[1] += rs[1]; x[2] += rs[2]; x[3] += rs[3]; }
        if (e_act == ACT_GELU) {
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] = x[q] * 0.5f * (1.0f + erff(x[q] * 0.70710678118654752440f));
        } else if (e_act == ACT_RELU) {
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] = fmaxf(x[q], 0.0f);
        } else if (e_act == ACT_LRELU) {
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] = x[q] > 0.0f ? x[q] : x[q] * 0.01f;
        }
        if (e_resmode == RES_POST) { x[0] += rs[0]; x[1] += rs[1]; x[2] += rs[2]; x[3] += rs[3]; }
        if (e_stats_out) {
          // producer of a folded LayerNorm (BN = 64: the 16 lanes of a row hold this CTA's 64 columns): partial mean and
          // M2 of the new residual row, slot blockIdx.x of 8.  All lanes take part in the shuffles.
          float sm = (x[0] + x[1]) + (x[2] + x[3]);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
          const float mj = sm * (1.0f / 64.0f);
          float qd = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float dd = x[q] - mj; qd += dd * dd; }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) qd += __shfl_xor_sync(0xffffffffu, qd, o);
          if (valid && (lane % LPR) == 0) {
            float* so = e_stats_out + ((long long)grow * 8 + blockIdx.x) * 2;
            so[0] = mj; so[1] = qd;
          }
        }
        if (valid) {
          if (e_out) {
            float* orow = e_out + (long long)grow * e_ldo + n;
            if (vec) *reinterpret_cast<float4*>(orow) = make_float4(x[0], x[1], x[2], x[3]);
            else { for (int q = 0; q < 4; ++q) if (n + q < e_N) orow[q] = x[q]; }
          }
          if (e_planes) {
            __half h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a = x[q] * e_pscale;
              if (e_prelu) a = fmaxf(a, 0.f);
              split_f16(a, h[q], l[q]);
            }
            __half* prow = e_planes + (long long)grow * e_ldp + n;
            if (pvec) {
              __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
              __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
              uint2 hv, lv;
              hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
              lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
              *reinterpret_cast<uint2*>(prow) = hv;
              *reinterpret_cast<uint2*>(prow + e_pstride) = lv;
            } else {
              for (int q = 0; q < 4; ++q) if (n + q < e_N) { prow[q] = h[q]; prow[e_pstride + q] = l[q]; }
            }
          }
        }
        tA = tB; rA = rB; tB = tC; rB = rC;
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 64) dbg[4] = clock64();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (dbg && threadIdx.x == 0) dbg[5] = clock64();
}


// ---------------------------------------------------------------------------------------------------------
// Trunk kernel: the Linear layers of the sampling loop (plain row-major operands, N a multiple of 64).
//
// Same TMA / mbarrier / TMEM pipeline as gemm_tc_kernel, with the two parts the timeline showed to dominate rebuilt:
//  * main loop: for tiles up to 128 columns the hi and lo weight planes sit back to back in a stage, so ONE MMA of
//    width 2*BN computes a_hi.[w_hi ; w_lo] -- the main accumulator and the first cross term -- and a second one of
//    width BN adds a_lo.w_hi.  Eight MMAs per K block instead of twelve, and the A tile is read from shared memory
//    twice instead of three times (the SS-mode operand path, ~90 B/cycle, is what bounds 128 x 64 tiles).
//  * epilogue: thread = accumulator row.  A thread reads 32 columns of both accumulators with tcgen05.ld and finishes
//    them in registers (scale, folded LayerNorm, bias, residual from a TMA-prefetched swizzled tile, exact-erf GELU,
//    row statistics, fp16 hi/lo split), writes the results into 128B-swizzled staging tiles (conflict-free for
//    lane = row) and one elected thread hands them to TMA stores.  No per-row loop, no global address arithmetic.
// ---------------------------------------------------------------------------------------------------------
struct FastEpi {
  const float* bias;       // [N] or null; folded LayerNorm: c = W beta + b
  const float* ln_s;       // [N] column sums of the gamma-scaled weight, or null (no folded LayerNorm)
  const float* ln_stats;   // consumer: [rows][16][2] partial (mean, M2) over 32 columns each of the 512-wide input row
  float* stats_out;        // producer (N = 512): [rows][16][2]
  float scale;             // undoes the operand pre-scaling
  int act, has_res, has_out, has_planes;
  int M, N;
  int attn;                // 1: the tile is [q 64 | k 64 | v 64] of one head half; the epilogue runs the 32-token attention
  // operand addressing: mode 0 plain rows (3-D map); mode 1 stride-1 "same" Conv1d over clips packed into the 128-row tile
  // (4-D map {C, T, clips, 2}, T | 128): K block kb = (tap j, channel block cb), TMA zero-fills the padding
  int mode, T, kb_per_tap, dil, pad;
  int planes_relu;         // the planes carry max(result, 0)
  int kb_split;            // mode 0: K blocks >= kb_split come from the second operand tensor (tmA2), counted from its column 0
  const unsigned* sig_in;  // row-tile signals (GemmP): wait for sig_in[blockIdx.y] >= sig_expect instead of the grid dependency ...
  int sig_expect;
  unsigned* sig_out;       // ... and add 1 to sig_out[blockIdx.y] when this CTA's results are in memory
  long long* dbg;
  int probe;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// the CTA only has to keep its staging memory alive until the TMA engine has read it; the writes themselves are
// performed before the grid counts as complete, which is what the dependent kernel's griddepcontrol.wait observes
__device__ __forceinline__ void tma_store_commit_wait(bool complete = false) {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  if (complete) {                                  // the writes themselves: a consumer that is not ordered by a kernel boundary follows
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.global;" ::: "memory");
  } else {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
// Dependency of a trunk layer: the whole predecessor grid (griddepcontrol.wait), or -- row-tile signals -- the column CTAs of the producer
// layer that wrote this CTA's 128 input rows.  The producers belong to an EARLIER launch, all of whose CTAs were resident before this launch
// became eligible (programmatic launch: every CTA of the primary has passed launch_dependents), so the wait cannot starve them.  Bounded
// like mbar_wait: a protocol bug ends in a trap.
__device__ __forceinline__ void tc_wait_dep(const unsigned* sig, int expect, int row_tile) {
  if (!sig) { pdl_wait(); return; }
  const unsigned* p = sig + row_tile;
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if ((int)v >= expect) return;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (clock64() - t0 > 4000000000LL) __trap();
  } while ((int)v < expect);
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  __half2 h = __halves2half2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// packed fp32 pair FMA (sm_100): (d0, d1) += (a0, a1) * (b0, b1)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}
// GELU of the trunk epilogue: x/2 (1 + erf(x / sqrt 2)) with a single-branch erf(z) = 1 - 2^(-z q(z)), q of degree 10 fitted in
// float64 to -log2(erfc z) / z on [0, 4.2] (tests/fit_erf.py; erf(4.2) rounds to 1 in fp32).  Max abs error of the fp32
// evaluation against float64 6.5e-7 -- the reference's own fp32 nn.GELU is at 1.3e-6 -- in 18 instructions per element
// instead of erff's 30 (two-range coefficient selects); the epilogue of the GELU layer is instruction-issue bound.
__device__ __constant__ float kErfQ[11] = {1.62790707e+00f, 9.18446271e-01f, 1.48284197e-01f, -2.76306103e-02f, -2.98958275e-04f, 2.56001420e-03f, -1.07292213e-03f, 2.56761395e-04f, -3.83041166e-05f, 3.31221616e-06f, -1.26982216e-07f};
__device__ __forceinline__ float gelu_sb(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.2f);
  float q = -1.26982216e-07f;
  q = fmaf(q, z, 3.31221616e-06f); q = fmaf(q, z, -3.83041166e-05f); q = fmaf(q, z, 2.56761395e-04f);
  q = fmaf(q, z, -1.07292213e-03f); q = fmaf(q, z, 2.56001420e-03f); q = fmaf(q, z, -2.98958275e-04f);
  q = fmaf(q, z, -2.76306103e-02f); q = fmaf(q, z, 1.48284197e-01f); q = fmaf(q, z, 9.18446271e-01f);
  q = fmaf(q, z, 1.62790707e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(z * q)));
  const float er = copysignf(1.0f - e, x);
  const float h = 0.5f * x;
  return fmaf(h, er, h);
}

// the same for two values at once: the degree-10 Horner chain runs as ten packed fma.rn.f32x2 (same roundings, so the results are
// bit-identical to gelu_sb); the GELU layer's epilogue is instruction-issue bound and the polynomial is more than half of it
__device__ __forceinline__ unsigned long long pack_f2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma_f2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void gelu_sb2(float& x0, float& x1) {
  const float z0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.2f), z1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.2f);
  const unsigned long long z = pack_f2(z0, z1);
  unsigned long long q = pack_f2(-1.26982216e-07f, -1.26982216e-07f);
  q = fma_f2(q, z, pack_f2(3.31221616e-06f, 3.31221616e-06f)); q = fma_f2(q, z, pack_f2(-3.83041166e-05f, -3.83041166e-05f));
  q = fma_f2(q, z, pack_f2(2.56761395e-04f, 2.56761395e-04f)); q = fma_f2(q, z, pack_f2(-1.07292213e-03f, -1.07292213e-03f));
  q = fma_f2(q, z, pack_f2(2.56001420e-03f, 2.56001420e-03f)); q = fma_f2(q, z, pack_f2(-2.98958275e-04f, -2.98958275e-04f));
  q = fma_f2(q, z, pack_f2(-2.76306103e-02f, -2.76306103e-02f)); q = fma_f2(q, z, pack_f2(1.48284197e-01f, 1.48284197e-01f));
  q = fma_f2(q, z, pack_f2(9.18446271e-01f, 9.18446271e-01f)); q = fma_f2(q, z, pack_f2(1.62790707e+00f, 1.62790707e+00f));
  float q0, q1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-(z0 * q0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-(z1 * q1)));
  const float h0 = 0.5f * x0, h1 = 0.5f * x1;
  x0 = fmaf(h0, copysignf(1.0f - e0, x0), h0);
  x1 = fmaf(h1, copysignf(1.0f - e1, x1), h1);
}

constexpr int ATT_LDK = 68;                      // K / V staging row stride (floats): conflict-free float4 stores for lane = row
constexpr int ATT_LDS = 36;                      // score row stride (floats)
constexpr int ATT_SX_BYTES = 128 * ATT_LDS * 4;  // partial scores written by the peer CTA of the cluster
constexpr int ATT_SX_TX = 128 * 32 * 4;          // bytes of them (32 scores per row; the row stride has 4 spare floats)
constexpr int ATT_KV_BYTES = 128 * ATT_LDK * 4;
constexpr int ATT_PL_OFF = 124928;               // plane staging box behind Q, K, V and P (3 x 34816 + 18432 = 122880 -> 1024-aligned)
// tensor-core attention (attn == 2): operand tiles in the (dead) ring memory, offsets from the ring base, all 128B-swizzled
// tiles with 128-byte rows.  Q / K / V: [128 tokens x 64 dims] hi, lo -- thread = token row writes its own row, no transposition:
// K_hi and K_lo back to back are one K-major N = 256 operand of Q K^T, and V is the MN-major B operand of P V (rows = keys = K index,
// the 64 dims of a row = N; V_hi ; V_lo one N = 128 operand with LBO = plane stride); P per 64-key block: [128 queries x 64 keys]
// hi, lo, K-major; the result's plane staging box reuses the Q tiles.
constexpr int ATT2_Q = 0, ATT2_K = 32768, ATT2_V = 65536, ATT2_P = 98304, ATT2_O = 0;
constexpr float kPScale = 1024.0f;               // softmax probabilities are split as (p * 1024): lo stays a normal fp16
constexpr int FAST_BOX_F32 = 128 * 32 * 4;       // one fp32 staging box: 128 rows x 32 columns, 128B rows
constexpr int FAST_BOX_PL = 2 * TC_A_PLANE;      // one plane staging box: 2 planes x 128 rows x 64 halfs

// LEGACY = true keeps what the product no longer runs -- the packed-fp32 FMA attention epilogue (attn == 1) and the MMA issue loop as
// first written -- as a second instantiation for A/B measurements and the test that holds the tensor-core attention against it
// (st_debug_probe bits 2048 / 16384); the product's instantiation is a third shorter, which its instruction fetch feels (-0.4 %).
//
// CHAIN = true is the body of gemm_tc_chain_kernel: the CTA runs layer after layer of a list (li of nl) as one CTA of a cluster
// of 8 = the 8 column tiles of one 128-row tile.  Every dependency between consecutive trunk layers stays inside a row tile (a
// Linear layer and the 32-token attention only read their own rows), so a cluster barrier after the layer's stores have completed
// replaces the kernel boundary: no CTA launch (0.9 us after the previous CTA's exit), no set-up, no grid-completion latency
// (tests/boundary_probe.py: the dependent's wait returns 1.3 - 2.1 us after the last CTA has exited).  TMEM is allocated once; the
// barriers of layer li + 1 are initialised (second barrier set) while layer li runs.
struct FastLayer {
  CUtensorMap tmA, tmW, tmO, tmP, tmA2, tmS;
  FastEpi ep;
  int num_kb, BN, STAGES, pad_;
};
constexpr int FAST_TAIL = 8192;                  // CHAIN: bias / column sums / barriers of two consecutive layers + the statistics staging at the end of shared memory
constexpr int FAST_STATS = 4096;                 // staging of the row statistics a producer of a folded LayerNorm leaves: 128 rows x <= 4 chunks x (mean, M2)
template <bool LEGACY, bool CHAIN>
__device__ __forceinline__ void fast_layer(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const CUtensorMap& tmP,
                                           const FastEpi& ep, const int num_kb, const int BN, const int STAGES, const CUtensorMap& tmA2, const CUtensorMap& tmS,
                                           const int li, const int nl, uint32_t& tmem_keep, const FastLayer* next, const int prev_stages) {
  const int W_PLANE = BN * TC_BK * 2;
  const int STAGE_BYTES = 2 * TC_A_PLANE + 2 * W_PLANE;
  const int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  const int NCH = BN >> 5;                            // 32-column chunks of the tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res_tile = smem + STAGES * STAGE_BYTES;                              // NCH boxes (only when has_res)
  uint32_t dyn_smem;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
  // CHAIN: the small per-layer state sits at a fixed place (two sets, layer parity), so that the next layer's set can be prepared
  // while this layer runs; otherwise it follows the ring
  uint8_t* tail_set = smem_raw + dyn_smem - FAST_TAIL;
  float* bias_s = CHAIN ? reinterpret_cast<float*>(tail_set + (li & 1) * 2048)
                        : reinterpret_cast<float*>(res_tile + (ep.attn ? ATT_SX_BYTES : ep.has_res ? NCH * FAST_BOX_F32 : 0));
  float* lns_s = bias_s + 192;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(lns_s + 192);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint64_t* res_bar = acc_bar + 1;
  // tensor-core attention (attn == 2): operand tiles ready for Q K^T (256 arrivals), scores complete (commit), P and V^T
  // ready (256 arrivals), P V complete (commit)
  // [4]: the peer CTA's partial scores have landed in this CTA's score buffer (st.async complete_tx, 16 KB)
  uint64_t* att_bar = res_bar + 1;
  uint32_t* tmem_slot = CHAIN ? reinterpret_cast<uint32_t*>(tail_set + 2048 - 16) : reinterpret_cast<uint32_t*>(att_bar + 5);
  // row statistics leave through a TMA store like every other result (no ordinary global store is left in the epilogue, so the
  // row-tile signal needs no gpu-scope fence): [128 rows][NCH] (mean, M2) pairs
  uint8_t* stats_s = CHAIN ? tail_set + 4096 : reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot) + 16 + 127) & ~uintptr_t(127));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;
  // timeline of one CTA: (0,0), which the launch puts on an idle SM ahead of time, or (probe bit 32) the LAST one, which like most
  // CTAs of a layer starts when its SM's CTA of the previous layer has exited
  const bool dbg_cta = (ep.probe & 32) ? (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1) : (blockIdx.x == 0 && blockIdx.y == 0);
  long long* dbg = (ep.dbg && dbg_cta) ? ep.dbg : nullptr;
  const long long t_entry = dbg ? clock64() : 0;     // every thread's own copy: a stamp that is relative to the entry does not read another thread's store
  if (dbg && threadIdx.x == 0) dbg[0] = t_entry;
  // boundary records (probe bit 64): every CTA of the selected launch leaves {globaltimer at entry, SM id, globaltimer after the
  // dependency wait, globaltimer at exit} behind the 64 timeline slots (tests/boundary_probe.py)
  long long* brec = (ep.dbg && (ep.probe & 64)) ? ep.dbg + 64 + 4 * (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (brec && threadIdx.x == 0) {
    unsigned long long gt; uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    brec[0] = (long long)gt; brec[1] = smid;
  }
  trace_stamp(20);                                   // CTA entry

  if (warp == 0 && lane == 0 && (!CHAIN || li == 0)) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (ep.has_out || ep.has_res) tma_prefetch_desc(&tmO);
    if (ep.has_planes) tma_prefetch_desc(&tmP);
    if (ep.kb_split < num_kb) tma_prefetch_desc(&tmA2);
  }
  // The producer warp initialises the barriers itself and goes straight to its loads: what lies between a CTA's entry and the request
  // for its first activation tile is on the critical path of every layer (the CTA that replaces the last-finishing CTA of the previous
  // layer enters 0.9 us after that exit), and the CTA barrier of the set-up -- TMEM allocation, bias loads -- was 0.45 us of it.  The
  // other nine warps meet at a named barrier the producer warp only ARRIVES at.
  if (warp == 0 && lane == 0) {
    auto init_set = [](uint64_t* fb, int stages, int attn, int old_stages) {
      uint64_t* eb = fb + stages;
      uint64_t* ab = eb + stages;                    // acc_bar, res_bar, att_bar[0..4]
      // the set held the barriers of the layer before this one's predecessor (all phases complete, nobody waiting): an mbarrier
      // object is invalidated before its memory becomes another one
      for (int i = 0; i < 2 * old_stages + 7 && old_stages > 0; ++i)
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(fb + i)) : "memory");
      for (int i = 0; i < stages; ++i) { mbar_init(&fb[i], 1); mbar_init(&eb[i], 1); }
      mbar_init(ab, 1);
      mbar_init(ab + 1, 1);
      mbar_init(ab + 2, 256); mbar_init(ab + 3, 1); mbar_init(ab + 4, 256); mbar_init(ab + 5, 1);
      mbar_init(ab + 6, 1);
      fence_barrier_init();
      if (attn == 2) mbar_expect_tx(ab + 6, ATT_SX_TX);            // the one arrival; the phase completes when the peer's 16 KB are in
    };
    if (!CHAIN || li == 0) init_set(full_bar, STAGES, ep.attn, 0);
    if (CHAIN && next)
      init_set(reinterpret_cast<uint64_t*>(tail_set + ((li + 1) & 1) * 2048 + 2 * 192 * 4), next->STAGES, next->ep.attn, li > 0 ? prev_stages : 0);
  }
  if (warp == 2 && (!CHAIN || li == 0)) {
    tmem_alloc(tmem_slot, CHAIN ? 512 : TMEM_COLS);
    tmem_relinquish();
  }
  // bias and LayerNorm column sums are weights: safe to read before the predecessor has finished.  The loads are issued
  // here and land in shared memory after the CTA barrier, off the set-up critical path (an L2 round trip).
  float bias_r = 0.f, lns_r = 0.f;
  const int et = threadIdx.x - 64;
  if (warp >= 2 && et < BN && n0 + et < ep.N) {
    if (ep.bias) bias_r = __ldg(ep.bias + n0 + et);
    if (ep.ln_s) lns_r = __ldg(ep.ln_s + n0 + et);
  }
  if (!CHAIN || li == 0) {
    if (warp == 0) {
      __syncwarp();
      asm volatile("bar.arrive 8, 320;" ::: "memory");       // barriers initialised (lane 0 above); does not wait for the others
    } else {
      tc_fence_before();
      asm volatile("bar.sync 8, 320;" ::: "memory");
      tc_fence_after();
      tmem_keep = *tmem_slot;
    }
  }
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, tmem_keep);    // REDUX result lives in a uniform register
  if (dbg && threadIdx.x == 0) dbg[32] = clock64();                       // set-up done
  if (!CHAIN || li == 0) pdl_launch();     // dependents may be scheduled; they still wait for this grid's completion before touching memory
  if (ep.attn && (!CHAIN || li == 0)) {
    // the attention epilogue writes into the peer CTA's shared memory: both CTAs of the cluster must be running before
    // either does (co-scheduling guarantees residency, not that the peer has started).  Everybody arrives here; the epilogue warps
    // wait before their first work, the producer and MMA warps after theirs (nothing of theirs touches the peer)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (warp >= 2) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      // The weight halves of the first ring pass do not depend on the predecessor kernel: they are issued before the wait.
      // What runs between the wait's return and the first activation tile's request is on the critical path of every layer
      // (timeline of a CTA that starts late, tests/layer_timeline.py with ST_PROBE=32768: 950 cycles when the generic loop -- a
      // division for the conv modes, probe / trace tests against constant memory, cold instruction-cache lines -- came first),
      // so everything is formed and pinned before the wait and the first pass of A tiles is one short loop behind it.
      const int pre = num_kb < STAGES ? num_kb : STAGES;
      const bool no_tma = (ep.probe & 2) != 0;
      // inside a chain the cluster barrier that ended the previous layer ordered the peers' stores before this thread; the proxy
      // fence orders them before its TMA loads.  It comes FIRST: behind the weight loads it waited for them to land (timeline:
      // 3 157 cycles from the layer's start to the first activation tile's request, 6 087 with the 128-wide weight tiles)
      if (dbg) dbg[33] = clock64();
      if (CHAIN && li > 0) asm volatile("fence.proxy.async;" ::: "memory");
      if (dbg) dbg[34] = clock64();
      if (!no_tma) {
        for (int kb = 0; kb < pre; ++kb) {
          mbar_expect_tx(&full_bar[kb], STAGE_BYTES);
          tma_load_3d(smem + kb * STAGE_BYTES + 2 * TC_A_PLANE, &tmW, &full_bar[kb], kb * TC_BK, n0, 0);
        }
      }
      if (dbg) dbg[35] = clock64();
      int mode = ep.mode, kb_split = ep.kb_split, kpt = ep.kb_per_tap, dil = ep.dil, pad = ep.pad;
      int clip = mode ? m0 / ep.T : 0;
      uint32_t a_dst = smem_u32(smem), a_bar = smem_u32(full_bar);
      int row0 = m0, stage_b = STAGE_BYTES, npre = no_tma ? 0 : pre;
      asm volatile("" : "+r"(mode), "+r"(kb_split), "+r"(kpt), "+r"(dil), "+r"(pad), "+r"(clip), "+r"(a_dst), "+r"(a_bar), "+r"(row0), "+r"(stage_b),
                   "+r"(npre));
      auto load_a = [&](int kb, uint32_t dst, uint32_t bar) {
        if (mode == 0) {
          if (kb < kb_split)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(bar), "r"(kb * TC_BK), "r"(row0), "r"(0) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmA2)), "r"(bar), "r"((kb - kb_split) * TC_BK), "r"(row0), "r"(0) : "memory");
        } else {
          const int j = kb / kpt, cb = kb - j * kpt;
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(bar), "r"(cb * TC_BK), "r"(j * dil - pad), "r"(clip), "r"(0) : "memory");
        }
      };
      if (!CHAIN || li == 0) tc_wait_dep(ep.sig_in, ep.sig_expect, blockIdx.y);
      // (the activation tiles were written through the async proxy, completed and fenced by their writers, and are read through the async
      // proxy by requests initiated after the acquire: no proxy fence on this side.  Probe bit 8 adds one: +400 cycles per layer, it waits
      // for the weight tiles in flight)
      if (ep.probe & 8) asm volatile("fence.proxy.async.global;" ::: "memory");
      for (int kb = 0; kb < npre; ++kb) load_a(kb, a_dst + kb * stage_b, a_bar + 8 * kb);
      if (CHAIN && next) {            // the next layer's descriptors, off its critical path
        tma_prefetch_desc(&next->tmA); tma_prefetch_desc(&next->tmW); tma_prefetch_desc(&next->tmO); tma_prefetch_desc(&next->tmP);
        tma_prefetch_desc(&next->tmA2);
      }
      trace_stamp(1, true);
      if (dbg) { const long long now = clock64(); dbg[1] = now; dbg[8] = now; }
      if (brec) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); brec[2] = (long long)gt; }
      // the residual tile is only needed by the epilogue: it is requested behind the first ring pass of operands, not in front
      // of it (all 128 CTAs start at once, and the first K block's arrival is what the tensor pipe waits for)
      if (ep.has_res) {
        mbar_expect_tx(res_bar, NCH * FAST_BOX_F32);
        for (int c = 0; c < NCH; ++c) tma_load_2d(res_tile + c * FAST_BOX_F32, &tmO, res_bar, n0 + c * 32, m0);
      }
      if (no_tma)
        for (int kb = 0; kb < pre; ++kb) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[kb])) : "memory");
      int s = pre == STAGES ? 0 : pre;
      uint32_t ph = pre == STAGES ? 1 : 0;                       // ring pass parity
      for (int kb = pre; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (no_tma) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
        } else {
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_3d(smem + s * STAGE_BYTES + 2 * TC_A_PLANE, &tmW, &full_bar[s], kb * TC_BK, n0, 0);
          load_a(kb, a_dst + s * stage_b, a_bar + 8 * s);
          if (dbg && kb < 16) dbg[8 + kb] = clock64();
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      asm volatile(".reg .pred st_pfull;");          // answer of the early barrier test (a predicate cannot cross asm statements otherwise)
      const uint32_t sb = smem_u32(smem);
      if (ep.probe & 1) {                            // timing probe: no MMAs, the ring is only drained
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == 0) trace_stamp(21, true);
          if (dbg && kb < 16) dbg[24 + kb] = clock64();
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      } else if (LEGACY && (ep.probe & 16)) {
        // the loop as first written (A/B reference, st_debug_probe bit 16384): blocking wait, descriptors rebuilt per K block
        const bool cat = BN <= 128;
        const uint32_t idesc_w = umma_idesc_f16(TC_BM, cat ? 2 * BN : BN), idesc_n = umma_idesc_f16(TC_BM, BN);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == 0) trace_stamp(21, true);
          if (dbg && kb < 16) dbg[24 + kb] = clock64();
          const uint32_t a_hi = sb + s * STAGE_BYTES, a_lo = a_hi + TC_A_PLANE, w_hi = a_hi + 2 * TC_A_PLANE, w_lo = w_hi + W_PLANE;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t dah = umma_desc_sw128(a_hi) + 2 * k, dal = umma_desc_sw128(a_lo) + 2 * k;
            const uint64_t dwh = umma_desc_sw128(w_hi) + 2 * k, dwl = umma_desc_sw128(w_lo) + 2 * k;
            if (cat) {
              umma_f16(tmem_base, dah, dwh, idesc_w, (kb | k) != 0);
              umma_f16(tmem_base + BN, dal, dwh, idesc_n, 1);
            } else {
              umma_f16(tmem_base, dah, dwh, idesc_n, (kb | k) != 0);
              umma_f16(tmem_base + BN, dah, dwl, idesc_n, (kb | k) != 0);
              umma_f16(tmem_base + BN, dal, dwh, idesc_n, 1);
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      } else if (BN <= 128) tc_issue_loop<true>(full_bar, empty_bar, sb, STAGE_BYTES, W_PLANE, BN, STAGES, tmem_base, num_kb, dbg);
      else tc_issue_loop<false>(full_bar, empty_bar, sb, STAGE_BYTES, W_PLANE, BN, STAGES, tmem_base, num_kb, dbg);
      umma_commit(acc_bar);
      if (dbg) dbg[2] = clock64();
    }
    if (ep.attn == 2) {
      // ===== tensor-core attention, MMA side: S = Q K^T over this CTA's 64 dims, then O = P V over the 128 keys of the tile
      // (block diagonal P: a query row only carries the 32 keys of its own sequence) =====
      const uint32_t s0 = smem_u32(smem);
      __syncwarp();
      if (elect_one()) {
        mbar_wait(&att_bar[0], 0);
        tc_fence_after();
        const uint32_t id_n = umma_idesc_f16(TC_BM, 128);
        const uint64_t dqh = umma_desc_sw128(s0 + ATT2_Q), dql = umma_desc_sw128(s0 + ATT2_Q + TC_A_PLANE);
        const uint64_t dkh = umma_desc_sw128(s0 + ATT2_K), dkl = umma_desc_sw128(s0 + ATT2_K + TC_A_PLANE);
        // the v columns of the qkv accumulators ([128,192) and [320,384)) are still being read while these run: the scores go
        // to [0,128) (q, k main: dead) and [384,512) (free)
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          umma_f16(tmem_base, dqh + 2 * k, dkh + 2 * k, id_n, k != 0);         // [0,128)   q_hi.k_hi
          umma_f16(tmem_base + 384, dqh + 2 * k, dkl + 2 * k, id_n, k != 0);   // [384,512) q_hi.k_lo
          umma_f16(tmem_base + 384, dql + 2 * k, dkh + 2 * k, id_n, 1);        //         + q_lo.k_hi
        }
        umma_commit(&att_bar[1]);
      }
      __syncwarp();
      if (elect_one()) {
        mbar_wait(&att_bar[2], 0);
        tc_fence_after();
        const uint32_t id_cat = umma_idesc_f16(TC_BM, 128) | kIdescBMajorMN, id_n = umma_idesc_f16(TC_BM, 64) | kIdescBMajorMN;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dph = umma_desc_sw128(s0 + ATT2_P + kb * 32768), dpl = umma_desc_sw128(s0 + ATT2_P + kb * 32768 + TC_A_PLANE);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // keys 64 kb + 16 k .. + 16 of V (128 bytes per key row); N = [64 dims of V_hi ; 64 dims of V_lo]
            const uint64_t dv = umma_desc_mn_sw128(s0 + ATT2_V + kb * 8192 + k * 2048, TC_A_PLANE, 1024);
            umma_f16(tmem_base + 128, dph + 2 * k, dv, id_cat, (kb | k) != 0);   // [128,192) p_hi.v_hi, [192,256) p_hi.v_lo (all of qkv is dead)
            umma_f16(tmem_base + 192, dpl + 2 * k, dv, id_n, 1);                 // + p_lo.v_hi
          }
        }
        umma_commit(&att_bar[3]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane group = warp % 4 (rows), the two warps of a group alternate 32-column chunks =====
    if (!CHAIN || li == 0) tc_wait_dep(ep.sig_in, ep.sig_expect, blockIdx.y);
    const int lg = warp & 3, cpart = (warp - 2) >> 2;
    const int r = lg * 32 + lane, grow = m0 + r;
    const bool row_ok = grow < ep.M;
    const uint32_t sw = (uint32_t)(r & 7);
    float rstd = 1.0f, u = 0.0f;
    const bool ln = ep.ln_s != nullptr;
    if (ln && row_ok) {
      // folded LayerNorm, consumer side: (mean, 1/sigma) of this thread's input row from its 16 partials over 32
      // columns each (Chan et al.: M2 = sum M2_j + n_j sum (mean_j - mean)^2)
      const float4* st4 = reinterpret_cast<const float4*>(ep.ln_stats + (long long)grow * 32);
      float mj[16], qj[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldcg(st4 + i);            // written by other SMs a layer ago: never from this SM's L1
        mj[2 * i] = t.x; qj[2 * i] = t.y; mj[2 * i + 1] = t.z; qj[2 * i + 1] = t.w;
      }
      float mu = 0.f, m2 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) mu += mj[j];
      mu *= 0.0625f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { const float dm = mj[j] - mu; m2 += qj[j] + 32.0f * dm * dm; }
      rstd = 1.0f / sqrtf(m2 * (1.0f / 512.0f) + 1e-5f);
      u = rstd * mu;
    }
    const float sc = ep.scale * rstd;
    if (et < BN) { bias_s[et] = bias_r; lns_s[et] = lns_r; }
    asm volatile("bar.sync 1, 256;" ::: "memory");                  // bias_s / lns_s visible to the eight epilogue warps
    if (ep.has_res) mbar_wait(res_bar, 0);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(22, true);    // accumulator complete
    if (dbg && threadIdx.x == 64) dbg[3] = clock64();
    const uint32_t stage0 = smem_u32(smem);
    const uint32_t out_base = stage0;                                                  // NCH fp32 boxes
    const uint32_t pl_base = stage0 + (ep.has_out ? NCH * FAST_BOX_F32 : 0);           // BN / 64 plane boxes
    const uint32_t res_base = smem_u32(res_tile);
    if (ep.attn == 2) {
      // ---- fused attention on the tensor cores (transformer.py:83-104).  The tile: 4 sequences x 32 tokens (TMEM lane group =
      // sequence, lane = token), 64 of the 128 dims of q, k, v of one head; the cluster peer holds the other 64 dims.
      //  1. thread = token row: q, k, v -> fp16 hi/lo operand tiles (row-major, 128B swizzle; the two warps of a lane group take
      //     the column halves [0,32) / [32,64) of each of q, k, v), plus the zeros of its row of the block-diagonal P
      //  2. MMA warp: S = Q K^T as three split MMAs into TMEM columns [0,256) (the qkv accumulators are dead by then)
      //  3. thread (row, key half): 16 partial scores of its row's own sequence -> the peer CTA (distributed shared memory),
      //     cluster barrier, add the peer's partial sums over the other 64 dims
      //  4. softmax over the row's 32 keys: each warp reduces its 16 keys, the pair merges (max, sum) through shared memory;
      //     P (times 1024) -> hi/lo tiles
      //  5. MMA warp: O = P V as three split MMAs (V is the MN-major B operand) into columns [384,512)
      //  6. thread (row, dim half): 32 dims of O -> plane staging (over the dead Q tiles) -> one TMA store
      const uint32_t sx = smem_u32(res_tile);
      uint32_t sx_peer;
      uint32_t peer_rank;                            // the other dim half of this head: the neighbour in the cluster (of 2, or of 8 in a chain)
      asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(peer_rank));
      peer_rank ^= 1u;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(sx_peer) : "r"(sx), "r"(peer_rank));
      uint32_t sxbar_peer;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(sxbar_peer) : "r"(smem_u32(&att_bar[4])), "r"(peer_rank));
      const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < 3; ++cc) {
        if (cc == 2) {
          // q and k are in place: the score MMAs run while v is converted
          tc_fence_before();
          fence_proxy_async();
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&att_bar[0])) : "memory");
        }
        const int c = 2 * cc + cpart;                    // chunks [q0 q1 k0 k1 v0 v1]: this warp's half of q, k, v
        uint32_t v[32], vc[32];
        tmem_ld32(trow + c * 32, v);
        tmem_ld32(trow + BN + c * 32, vc);
        tmem_ld_wait();
        const uint32_t row = stage0 + (uint32_t)(cc * 32768 + r * 128);                 // ATT2_Q / ATT2_K / ATT2_V
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x[8];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + q * 8 + t * 4);
            const float4 s4 = *reinterpret_cast<const float4*>(lns_s + c * 32 + q * 8 + t * 4);
            const int j = 8 * q + 4 * t;
            x[4 * t + 0] = (__uint_as_float(v[j + 0]) + __uint_as_float(vc[j + 0])) * sc + b4.x - u * s4.x;
            x[4 * t + 1] = (__uint_as_float(v[j + 1]) + __uint_as_float(vc[j + 1])) * sc + b4.y - u * s4.y;
            x[4 * t + 2] = (__uint_as_float(v[j + 2]) + __uint_as_float(vc[j + 2])) * sc + b4.z - u * s4.z;
            x[4 * t + 3] = (__uint_as_float(v[j + 3]) + __uint_as_float(vc[j + 3])) * sc + b4.w - u * s4.w;
          }
          uint4 h, l;
          split8_f16(x, kActScale, h, l);
          const uint32_t off = (((uint32_t)(cpart * 4 + q)) ^ sw) << 4;
          sts128u(row + off, h);
          sts128u(row + TC_A_PLANE + off, l);
        }
      }
      {
        // zeros of row r of plane `cpart` of P: the other sequence's 32 keys in the row's own 64-key block, all 64 keys of the other
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t own = stage0 + (uint32_t)(ATT2_P + (lg >> 1) * 32768 + cpart * TC_A_PLANE + r * 128);
        const uint32_t oth = stage0 + (uint32_t)(ATT2_P + ((lg >> 1) ^ 1) * 32768 + cpart * TC_A_PLANE + r * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) sts128u(own + ((((uint32_t)(((lg & 1) ^ 1) * 4 + q)) ^ sw) << 4), z4);
#pragma unroll
        for (int q = 0; q < 8; ++q) sts128u(oth + (((uint32_t)q ^ sw) << 4), z4);
      }
      tc_fence_before();                                 // this thread's reads of the v accumulators precede the P V MMAs (att_bar[2])
      if (dbg && threadIdx.x == 64) { dbg[40] = t_entry; dbg[41] = dbg[3]; dbg[42] = clock64(); }
      // partial scores of this row over the 64 local dims: keys [16 cpart, +16) of the sequence's diagonal block
      float sv[16];
      const uint32_t srow = (uint32_t)(r * ATT_LDS + cpart * 16) * 4;
      mbar_wait(&att_bar[1], 0);
      tc_fence_after();
      {
        uint32_t v[16], vc[16];
        tmem_ld16(trow + lg * 32 + cpart * 16, v);
        tmem_ld16(trow + 384 + lg * 32 + cpart * 16, vc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sv[j] = __uint_as_float(v[j]) + __uint_as_float(vc[j]);
        // to the peer's score buffer; every store also counts its 16 bytes on the peer's barrier, so the receiver waits for data,
        // not for a cluster-wide barrier
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                       ::"r"(sx_peer + srow + 16 * q), "f"(sv[4 * q]), "f"(sv[4 * q + 1]), "f"(sv[4 * q + 2]), "f"(sv[4 * q + 3]), "r"(sxbar_peer) : "memory");
      }
      if (dbg && threadIdx.x == 64) dbg[43] = clock64();
      mbar_wait(&att_bar[4], 0);
      if (dbg && threadIdx.x == 64) dbg[44] = clock64();
      {
        // softmax over the 32 keys of the row; the operands were scaled by kActScale each, exp(x) = 2^(x log2 e)
        const float ssc = 0.08838834764831845f * 1.4426950408889634f / (kActScale * kActScale);
        float mw = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 o4 = lds128(sx + srow + 16 * q);
          sv[4 * q] = (sv[4 * q] + o4.x) * ssc; sv[4 * q + 1] = (sv[4 * q + 1] + o4.y) * ssc;
          sv[4 * q + 2] = (sv[4 * q + 2] + o4.z) * ssc; sv[4 * q + 3] = (sv[4 * q + 3] + o4.w) * ssc;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) mw = fmaxf(mw, sv[j]);
        float lw = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { sv[j] = exp2f(sv[j] - mw); lw += sv[j]; }
        // merge with the partner warp's 16 keys: (max, sum) pairs meet in the four spare floats of the row's score slot
        const uint32_t mrow = sx + (uint32_t)(r * ATT_LDS + 32) * 4;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(mrow + 8 * cpart), "f"(mw), "f"(lw) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(2 + lg) : "memory");
        float mo, lo_;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(mo), "=f"(lo_) : "r"(mrow + 8 * (cpart ^ 1)) : "memory");
        const float m = fmaxf(mw, mo);
        const float fw = exp2f(mw - m), fo = exp2f(mo - m);
        const float inv = fw * kPScale / (lw * fw + lo_ * fo);
        const uint32_t prow = stage0 + (uint32_t)(ATT2_P + (lg >> 1) * 32768 + r * 128);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 h, l;
          split8_f16(sv + 8 * q, inv, h, l);
          const uint32_t off = (((uint32_t)((lg & 1) * 4 + cpart * 2 + q)) ^ sw) << 4;
          sts128u(prow + off, h);
          sts128u(prow + TC_A_PLANE + off, l);
        }
        tc_fence_before();
        fence_proxy_async();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&att_bar[2])) : "memory");
      }
      if (dbg && threadIdx.x == 64) dbg[45] = clock64();
      // O: this thread's row, 32 of the CTA's 64 dims; accumulators carry (p * 1024) (v * 16), the planes want o * 16
      mbar_wait(&att_bar[3], 0);
      tc_fence_after();
      if (dbg && threadIdx.x == 64) dbg[46] = clock64();
      {
        uint32_t v[32], vc[32];
        tmem_ld32(trow + 128 + cpart * 32, v);
        tmem_ld32(trow + 192 + cpart * 32, vc);
        tmem_ld_wait();
        const uint32_t orow = stage0 + (uint32_t)(ATT2_O + r * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[8 * q + e]) + __uint_as_float(vc[8 * q + e]);
          uint4 h, l;
          split8_f16(x, 1.0f / kPScale, h, l);
          const uint32_t off = (((uint32_t)(cpart * 4 + q)) ^ sw) << 4;
          sts128u(orow + off, h);
          sts128u(orow + TC_A_PLANE + off, l);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && elect_one()) {
        tma_store_3d(&tmP, stage0 + ATT2_O, blockIdx.x * 64, m0, 0);
        tma_store_commit_wait(CHAIN || ep.sig_out != nullptr);
      }
      if (dbg && threadIdx.x == 64) { dbg[47] = clock64(); dbg[4] = dbg[47]; }
    } else if (LEGACY && ep.attn) {
      // ---- fused attention (transformer.py:83-104).  This CTA holds, for 4 sequences x 32 tokens (TMEM lane group = one
      // sequence, lane = token), 64 of the 128 dims of q, k and v of one head; its cluster peer holds the other 64.
      // q, k, v go to shared memory (fp32); the two warps of a lane group then take 16 query rows each and work in
      // 4 x 4 register tiles (packed fp32x2 FMAs): partial scores over the 64 local dims, exchanged with the peer
      // through distributed shared memory, softmax across the 8 lanes of a row, and this CTA's 64 dims of P V.
      const uint32_t qs = stage0, ks = stage0 + ATT_KV_BYTES, vs = stage0 + 2 * ATT_KV_BYTES, pp = stage0 + 3 * ATT_KV_BYTES;
      const uint32_t sx = smem_u32(res_tile);
      uint32_t sx_peer;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(sx_peer) : "r"(sx), "r"((uint32_t)((blockIdx.x & 1) ^ 1)));
      // phase A: the 6 chunks [q0 q1 k0 k1 v0 v1] of this lane group's rows -> shared memory, 3 chunks per warp
#pragma unroll 1
      for (int cc = 0; cc < 3; ++cc) {
        const int c = cpart * 3 + cc;
        uint32_t v[32], vc[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + c * 32, v);
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + BN + c * 32, vc);
        tmem_ld_wait();
        const uint32_t dst = stage0 + (uint32_t)((c >> 1) * ATT_KV_BYTES + (r * ATT_LDK + (c & 1) * 32) * 4);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + t * 4);
          const float4 s4 = *reinterpret_cast<const float4*>(lns_s + c * 32 + t * 4);
          float4 x;
          x.x = (__uint_as_float(v[4 * t + 0]) + __uint_as_float(vc[4 * t + 0])) * sc + b4.x - u * s4.x;
          x.y = (__uint_as_float(v[4 * t + 1]) + __uint_as_float(vc[4 * t + 1])) * sc + b4.y - u * s4.y;
          x.z = (__uint_as_float(v[4 * t + 2]) + __uint_as_float(vc[4 * t + 2])) * sc + b4.z - u * s4.z;
          x.w = (__uint_as_float(v[4 * t + 3]) + __uint_as_float(vc[4 * t + 3])) * sc + b4.w - u * s4.w;
          sts128(dst + t * 16, x);
        }
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 64;" ::"r"(2 + lg) : "memory");     // q, k, v of this sequence are in place
      if (dbg && threadIdx.x == 64) { dbg[40] = t_entry; dbg[41] = dbg[3]; dbg[42] = clock64(); }
      // phase B: thread (rq, kq) of warp (lg, cpart): query rows 16*cpart + rq + 4a, keys kq + 8b  (a, b = 0..3)
      const int rq = lane >> 3, kq = lane & 7;
      const int row0 = lg * 32 + cpart * 16 + rq;                     // tile row of a = 0
      float acc[4][4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.f;
      {
        const uint32_t qrow = qs + (uint32_t)(row0 * ATT_LDK) * 4, krow = ks + (uint32_t)((lg * 32 + kq) * ATT_LDK) * 4;
        // software pipeline: the operands of step d4 + 1 are fetched while step d4's FMAs issue
        float4 qa[2][4], kb[2][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) qa[0][a] = lds128(qrow + (uint32_t)(4 * a * ATT_LDK) * 4);
#pragma unroll
        for (int b = 0; b < 4; ++b) kb[0][b] = lds128(krow + (uint32_t)(8 * b * ATT_LDK) * 4);
#pragma unroll
        for (int d4 = 0; d4 < 16; ++d4) {
          const int cur = d4 & 1, nxt = cur ^ 1;
          if (d4 + 1 < 16) {
#pragma unroll
            for (int a = 0; a < 4; ++a) qa[nxt][a] = lds128(qrow + (uint32_t)(4 * a * ATT_LDK + 4 * (d4 + 1)) * 4);
#pragma unroll
            for (int b = 0; b < 4; ++b) kb[nxt][b] = lds128(krow + (uint32_t)(8 * b * ATT_LDK + 4 * (d4 + 1)) * 4);
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              ffma2(acc[a][b][0], acc[a][b][1], qa[cur][a].x, qa[cur][a].y, kb[cur][b].x, kb[cur][b].y);
              ffma2(acc[a][b][0], acc[a][b][1], qa[cur][a].z, qa[cur][a].w, kb[cur][b].z, kb[cur][b].w);
            }
        }
      }
      float sv[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          sv[a][b] = acc[a][b][0] + acc[a][b][1];
          asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(sx_peer + (uint32_t)((row0 + 4 * a) * ATT_LDS + kq + 8 * b) * 4), "f"(sv[a][b]) : "memory");
        }
      if (dbg && threadIdx.x == 64) dbg[43] = clock64();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      if (dbg && threadIdx.x == 64) dbg[44] = clock64();
      // phase C: add the peer's half, scale by 128^-0.5, softmax over the 32 keys of each row (8 lanes x 4 keys).  The four
      // rows of a thread advance in lock step so that their shuffle / exp chains overlap.
      float mx[4], sum[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        mx[a] = -INFINITY;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          sv[a][b] = (sv[a][b] + lds32(sx + (uint32_t)((row0 + 4 * a) * ATT_LDS + kq + 8 * b) * 4)) * 0.08838834764831845f;
          mx[a] = fmaxf(mx[a], sv[a][b]);
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1)
#pragma unroll
        for (int a = 0; a < 4; ++a) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        sum[a] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) { sv[a][b] = expf(sv[a][b] - mx[a]); sum[a] += sv[a][b]; }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1)
#pragma unroll
        for (int a = 0; a < 4; ++a) sum[a] += __shfl_xor_sync(0xffffffffu, sum[a], o);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float inv = 1.0f / sum[a];
#pragma unroll
        for (int b = 0; b < 4; ++b) sts32(pp + (uint32_t)((row0 + 4 * a) * ATT_LDS + kq + 8 * b) * 4, sv[a][b] * inv);
      }
      __syncwarp();                                                  // a warp consumes only the P rows it produced
      if (dbg && threadIdx.x == 64) dbg[45] = clock64();
      // phase D: thread (rq, dq): rows as above, dims 4*dq..+3 and 32 + 4*dq..+3 of this CTA's 64
      float o[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 8; ++d) o[a][d] = 0.f;
      {
        const uint32_t prow = pp + (uint32_t)(row0 * ATT_LDS) * 4, vrow = vs + (uint32_t)(lg * 32 * ATT_LDK + 4 * kq) * 4;
        // per key j: p[a][j] (4 rows) times v[j][8 dims]; the v rows of key j + 1 are fetched while key j's FMAs issue
        float4 v0[2], v1[2];
        v0[0] = lds128(vrow);
        v1[0] = lds128(vrow + 32 * 4);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 pa[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) pa[a] = lds128(prow + (uint32_t)(4 * a * ATT_LDS + 4 * j4) * 4);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * j4 + jj, cur = j & 1, nxt = cur ^ 1;
            if (j + 1 < 32) {
              v0[nxt] = lds128(vrow + (uint32_t)((j + 1) * ATT_LDK) * 4);
              v1[nxt] = lds128(vrow + (uint32_t)((j + 1) * ATT_LDK + 32) * 4);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const float pj = jj == 0 ? pa[a].x : jj == 1 ? pa[a].y : jj == 2 ? pa[a].z : pa[a].w;
              ffma2(o[a][0], o[a][1], pj, pj, v0[cur].x, v0[cur].y); ffma2(o[a][2], o[a][3], pj, pj, v0[cur].z, v0[cur].w);
              ffma2(o[a][4], o[a][5], pj, pj, v1[cur].x, v1[cur].y); ffma2(o[a][6], o[a][7], pj, pj, v1[cur].z, v1[cur].w);
            }
          }
        }
      }
      // fp16 hi/lo planes of the result into the swizzled staging box: row, 16-byte chunk (dq >> 1) and (dq >> 1) + 4
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int rr = row0 + 4 * a;
        const uint32_t base = stage0 + ATT_PL_OFF + (uint32_t)(rr * 128) + (uint32_t)((kq & 1) * 8);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          __half h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_f16(o[a][4 * g + e] * kActScale, h[e], l[e]);
          const uint32_t off = (((uint32_t)((kq >> 1) + 4 * g)) ^ (uint32_t)(rr & 7)) << 4;
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + off), "r"(pack_h2(h[0], h[1])), "r"(pack_h2(h[2], h[3])) : "memory");
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + TC_A_PLANE + off), "r"(pack_h2(l[0], l[1])), "r"(pack_h2(l[2], l[3])) : "memory");
        }
      }
      if (dbg && threadIdx.x == 64) dbg[46] = clock64();
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && elect_one()) {
        tma_store_3d(&tmP, stage0 + ATT_PL_OFF, blockIdx.x * 64, m0, 0);
        tma_store_commit_wait(CHAIN || ep.sig_out != nullptr);
      }
      if (dbg && threadIdx.x == 64) dbg[47] = clock64();
      if (dbg && threadIdx.x == 64) dbg[4] = clock64();
    } else {
#pragma unroll 1
    for (int c = cpart; c < NCH; c += 2) {
      uint32_t v[32], vc[32];
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + c * 32, v);
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + BN + c * 32, vc);
      tmem_ld_wait();
      float x[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + q * 4);
        x[4 * q + 0] = (__uint_as_float(v[4 * q + 0]) + __uint_as_float(vc[4 * q + 0])) * sc + b4.x;
        x[4 * q + 1] = (__uint_as_float(v[4 * q + 1]) + __uint_as_float(vc[4 * q + 1])) * sc + b4.y;
        x[4 * q + 2] = (__uint_as_float(v[4 * q + 2]) + __uint_as_float(vc[4 * q + 2])) * sc + b4.z;
        x[4 * q + 3] = (__uint_as_float(v[4 * q + 3]) + __uint_as_float(vc[4 * q + 3])) * sc + b4.w;
      }
      if (ln) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(lns_s + c * 32 + q * 4);
          x[4 * q + 0] -= u * s4.x; x[4 * q + 1] -= u * s4.y; x[4 * q + 2] -= u * s4.z; x[4 * q + 3] -= u * s4.w;
        }
      }
      if (ep.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) gelu_sb2(x[j], x[j + 1]);
      } else if (ep.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
      }
      if (ep.has_res) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 r4 = lds128(res_base + (uint32_t)(c * FAST_BOX_F32 + r * 128) + (((uint32_t)q ^ sw) << 4));
          x[4 * q + 0] += r4.x; x[4 * q + 1] += r4.y; x[4 * q + 2] += r4.z; x[4 * q + 3] += r4.w;
        }
      }
      if (ep.stats_out) {
        // producer of a folded LayerNorm: (mean, M2) of this thread's 32 columns of the new residual row
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) sm += x[j];
        const float mj = sm * (1.0f / 32.0f);
        float qd = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float dd = x[j] - mj; qd += dd * dd; }
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(smem_u32(stats_s) + (uint32_t)((r * NCH + c) * 8)), "f"(mj), "f"(qd) : "memory");
      }
      // Results leave as soon as a staging box is complete, not at the end of the epilogue: the TMA store of a box drains
      // (all 128 CTAs write 4-12 MB at once) while the planes of the same chunk / the next chunk are still being computed.
      if (ep.has_out) {
        const uint32_t orow = out_base + (uint32_t)(c * FAST_BOX_F32 + r * 128);
#pragma unroll
        for (int q = 0; q < 8; ++q) sts128(orow + (((uint32_t)q ^ sw) << 4), make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]));
        fence_proxy_async();                                          // staging writes -> visible to the TMA engine
        asm volatile("bar.sync %0, 128;" ::"r"(6 + cpart) : "memory");   // the four warps that own fp32 box c
        if (lg == 0 && lane == 0) {
          tma_store_2d(&tmO, out_base + c * FAST_BOX_F32, n0 + c * 32, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (ep.has_planes) {
        const uint32_t prow = pl_base + (uint32_t)((c >> 1) * FAST_BOX_PL + r * 128);
        if (ep.planes_relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);      // only the planes carry the ReLU; x is not used after them
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 h, l;
          split8_f16(x + 8 * q, kActScale, h, l);
          const uint32_t off = (((uint32_t)((c & 1) * 4 + q)) ^ sw) << 4;
          sts128u(prow + off, h);
          sts128u(prow + TC_A_PLANE + off, l);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");                // both column halves of plane box c / 2 (all eight warps)
        if (warp == 4 && lane == 0) {
          tma_store_3d(&tmP, pl_base + (c >> 1) * FAST_BOX_PL, n0 + (c >> 1) * 64, m0, 0);
          // the statistics of all chunks are staged once the last plane box is (the barrier above had all eight warps)
          if (ep.stats_out && c + 2 >= NCH) tma_store_2d(&tmS, smem_u32(stats_s), blockIdx.x * NCH * 2, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
    // the staging memory must outlive the TMA engine's reads: each issuing thread waits for its own bulk groups
    if (lg == 0 && lane == 0) {
      if (CHAIN || ep.sig_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (ep.sig_out) asm volatile("fence.proxy.async.global;" ::: "memory");   // the completed async-proxy writes, before the generic-proxy count (free here: nothing in flight)
      if (dbg && warp == 4) dbg[36] = clock64();
    }

    if (dbg && threadIdx.x == 64) dbg[4] = clock64();
    }
  }
  if (ep.attn && (!CHAIN || li == 0) && warp < 2) {
    __syncwarp();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // pairs with the arrive of the set-up
  }
  if (LEGACY && ep.attn == 1 && warp < 2) {
    // the producer and MMA warps take part in the cluster barrier of the FMA attention epilogue (the tensor-core one exchanges
    // its scores through st.async + an mbarrier instead)
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[37] = clock64();
  if (ep.sig_out && threadIdx.x == 0 && (!CHAIN || li + 1 == nl)) {
    // Every result of the layer -- the row statistics too -- left through TMA stores that their issuers waited for
    // (cp.async.bulk.wait_group 0: the writes are performed) before the CTA barrier above, so the count needs no fence: a releasing
    // count (gpu-scope fence) kept the CTA alive ~1 300 cycles longer (timeline), and the next layer's CTA on this SM starts that much later.
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ep.sig_out + blockIdx.y) : "memory");
  }
  trace_stamp(23);                                   // epilogue and stores done
  if (CHAIN && li + 1 < nl) {
    // layer boundary inside the chain: this CTA's stores are complete (the issuing threads waited for the writes, the row
    // statistics are ordinary stores), release them to the 7 peers of the row tile and acquire theirs
    // (one warp releases for the CTA -- the CTA barrier above made the other warps' writes its own, release is cumulative --
    // the rest arrive relaxed: 320 releasing threads per CTA took 3 200 cycles from the last store to the next layer's start)
    if (warp == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    } else {
      asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
  } else if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, CHAIN ? 512 : TMEM_COLS);
  }
  if (dbg && threadIdx.x == 0) dbg[5] = clock64();
  if (brec && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); brec[3] = (long long)gt; }
}

template <bool LEGACY>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_fast_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP, const FastEpi ep,
                    const int num_kb, const int BN, const int STAGES, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmS) {
  uint32_t keep = 0;
  fast_layer<LEGACY, false>(tmA, tmW, tmO, tmP, ep, num_kb, BN, STAGES, tmA2, tmS, 0, 1, keep, nullptr, 0);
}

// The trunk layers of one evaluation stack as ONE launch: clusters of 8 CTAs (grid.x = 8 column tiles of every layer), one cluster
// per 128-row tile, every CTA walking the layer list (see fast_layer<., true>).
constexpr int FAST_CHAIN_MAX = 32;
struct FastChain {
  int nl, pad_[15];
  FastLayer ly[FAST_CHAIN_MAX];
};
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_chain_kernel(const __grid_constant__ FastChain P) {
  uint32_t keep = 0;
  const int nl = P.nl;
#pragma unroll 1
  for (int li = 0; li < nl; ++li) {
    const FastLayer& L = P.ly[li];
    fast_layer<false, true>(L.tmA, L.tmW, L.tmO, L.tmP, L.ep, L.num_kb, L.BN, L.STAGES, L.tmA2, L.tmS, li, nl, keep, li + 1 < nl ? &P.ly[li + 1] : nullptr,
                            li > 0 ? P.ly[li - 1].STAGES : 0);
  }
}

// fp32 [M,K] (row stride lda) -> fp16 planes [2][M][Kp], value * scale split into hi + lo
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ a, int lda, int M, int K, int Kp, float scale, int relu,
                                                           __half* __restrict__ planes, long long plane_stride) {
  pdl_wait();
  trace_stamp(6);
  pdl_launch();
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;     // one thread per 4 elements
  const int kq = Kp >> 2;
  if (gid >= (long long)M * kq) return;
  const int m = (int)(gid / kq), k = (int)(gid - (long long)m * kq) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (k + 3 < K && (lda & 3) == 0) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(a + (long long)m * lda + k));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    for (int q = 0; q < 4; ++q)
      if (k + q < K) v[q] = a[(long long)m * lda + k + q];
  }
  __half h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float t = v[q] * scale;
    if (relu) t = fmaxf(t, 0.f);
    split_f16(t, h[q], l[q]);
  }
  __half* ph = planes + (long long)m * Kp + k;
  __half* pl = ph + plane_stride;
  *reinterpret_cast<__half2*>(ph) = __halves2half2(h[0], h[1]);
  *reinterpret_cast<__half2*>(ph + 2) = __halves2half2(h[2], h[3]);
  *reinterpret_cast<__half2*>(pl) = __halves2half2(l[0], l[1]);
  *reinterpret_cast<__half2*>(pl + 2) = __halves2half2(l[2], l[3]);
}

__global__ void absmax_kernel(const float* __restrict__ a, long long n, float* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(a[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // non-negative floats order like ints
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapKey {
  const void* base; long long plane_stride; int d0, d1, d2, box;
  bool operator<(const MapKey& o) const {
    return std::tie(base, plane_stride, d0, d1, d2, box) < std::tie(o.base, o.plane_stride, o.d0, o.d1, o.d2, o.box);
  }
};
static std::map<MapKey, CUtensorMap> g_maps;    // a tensor map encodes address and geometry only, never data: safe to reuse
static std::mutex g_tc_mu;

// planes [2][rows][Kp] fp16 (planes `plane_stride` elements apart) -> 3-D map {Kp, rows, 2}, box {64, box_rows, 2},
// 128B swizzle, zero fill outside the tensor
static int get_map_3d(const CUtensorMap** out, const __half* base, long long plane_stride, int rows, int Kp, int box_rows, int ld = 0) {
  if (ld == 0) ld = Kp;                 // row stride in elements (a column window of a wider tensor has ld > Kp)
  MapKey key{base, plane_stride, Kp, rows, ld == Kp ? 0 : ld, box_rows};
  std::lock_guard<std::mutex> lk(g_tc_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = &it->second; return ST_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ST_ECUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {TC_BK, (cuuint32_t)box_rows, 2};
  cuuint32_t est[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3d rows=%d Kp=%d box=%d) failed: %d", rows, Kp, box_rows, (int)r); return ST_ECUDA; }
  *out = &(g_maps[key] = m);
  return ST_OK;
}
// planes [2][clips][T][C] fp16 -> 4-D map {C, T, clips, 2}, box {64, T, 128/T, 2}
static int get_map_4d(const CUtensorMap** out, const __half* base, long long plane_stride, int clips, int T, int C) {
  MapKey key{base, plane_stride, C, T, clips, -1};
  std::lock_guard<std::mutex> lk(g_tc_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = &it->second; return ST_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ST_ECUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)clips, 2};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[4] = {TC_BK, (cuuint32_t)T, (cuuint32_t)(TC_BM / T), 2};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(4d clips=%d T=%d C=%d) failed: %d", clips, T, C, (int)r); return ST_ECUDA; }
  *out = &(g_maps[key] = m);
  return ST_OK;
}

struct WPlanes {
  __half* planes = nullptr;
  int N = 0, Kp = 0;
  float inv_scale = 1.f;   // 2^-sw
};
static std::unordered_map<const float*, WPlanes> g_wplanes;
static std::mutex g_w_mu;
long long* g_tc_dbg = nullptr;  // set by st_debug_timeline
int g_tc_probe = 0;             // set by st_debug_probe
int g_tc_dbg_n2 = -1, g_tc_dbg_k2 = -1;   // st_debug_timeline_select2: a second shape, recorded 1024 slots further on
int g_tc_dbg_n = 0, g_tc_dbg_k = 0;   // st_debug_timeline_select: only launches with this N, K record the timeline (0 = all)
bool g_tc_fast = true;          // trunk kernel for the shapes it takes (st_debug_probe bit 16 turns it off)
bool g_tc_taps = true;          // staged-taps conv mode of the generic kernel (st_debug_probe bit 8192 turns it off)
// Activation planes of a GEMM whose operand is still fp32: one arena PER STREAM.  Stream order serialises the reuse on one
// stream; two handles driven on two streams at once (the library is re-entrant per handle) must not share it.
static std::map<cudaStream_t, Arena> g_scratch;
static std::mutex g_scratch_mu;

static int split_launch(const float* a, int lda, int M, int K, int Kp, float scale, int relu, __half* planes, cudaStream_t s) {
  const long long n = (long long)M * (Kp >> 2);
launch_k(split_planes_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, a, lda, M, K, Kp, scale, relu, planes, (long long)M * Kp);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// Weight planes are made once per weight matrix (first use): per-tensor power-of-two scale so that max|w| lands in
// [2^12, 2^13), then the fp16 hi/lo split.  This synchronises; it happens during warm-up, never in steady state.
static int get_wplanes(const GemmP& p, cudaStream_t s, WPlanes** out) {
  std::lock_guard<std::mutex> lk(g_w_mu);
  auto it = g_wplanes.find(p.W);
  if (it != g_wplanes.end() && it->second.N == p.N) { *out = &it->second; return ST_OK; }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);
  if (cap != cudaStreamCaptureStatusNone) { set_error("weight planes must exist before graph capture"); return ST_ESTATE; }
  WPlanes w;
  w.N = p.N;
  w.Kp = (p.K + TC_BK - 1) / TC_BK * TC_BK;
  float* d_max = nullptr;
  ST_CHECK_CUDA(cudaMalloc(&d_max, sizeof(float)));
  ST_CHECK_CUDA(cudaMemsetAsync(d_max, 0, sizeof(float), s));
launch_k(absmax_kernel, dim3(148), dim3(256), 0, s, p.W, (long long)p.N * p.ldw, d_max);
  ST_CHECK_LAUNCH();
  float h_max = 0.f;
  ST_CHECK_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(float), cudaMemcpyDeviceToHost, s));
  ST_CHECK_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_max);
  int e = 0;
  if (h_max > 0.f) {
    frexpf(h_max, &e);            // h_max = f * 2^e, f in [0.5, 1)
    e = 13 - e;                   // h_max * 2^e in [2^12, 2^13)
  }
  const float scale = ldexpf(1.0f, e);
  w.inv_scale = ldexpf(1.0f, -e);
  const size_t bytes = (size_t)2 * p.N * w.Kp * sizeof(__half);
  if (cudaMalloc(&w.planes, bytes) != cudaSuccess) { set_error("cudaMalloc(weight planes %zu B) failed", bytes); return ST_ENOMEM; }
  ST_TRY(split_launch(p.W, p.ldw, p.N, p.K, w.Kp, scale, 0, w.planes, s));
  ST_CHECK_CUDA(cudaStreamSynchronize(s));
  g_wplanes[p.W] = w;
  *out = &g_wplanes[p.W];
  return ST_OK;
}

int tc_split(const float* a, int lda, int M, int K, __half* planes, cudaStream_t s) { return split_launch(a, lda, M, K, K, kActScale, 0, planes, s); }

void tc_forget_weights(const float* W) {
  std::lock_guard<std::mutex> lk(g_w_mu);
  auto it = g_wplanes.find(W);
  if (it == g_wplanes.end()) return;
  {
    std::lock_guard<std::mutex> lk2(g_tc_mu);
    for (auto m = g_maps.begin(); m != g_maps.end();) m = (m->first.base == it->second.planes) ? g_maps.erase(m) : std::next(m);
  }
  cudaFree(it->second.planes);
  g_wplanes.erase(it);
}

static int conv_mode(const GemmP& p) {
  // 0 plain, 1 packed clips, 2 long clips stride 1, 3 strided pad-0, -1 unsupported
  if (p.ups) return -1;
  if (p.Lout == p.M && p.Lin == p.M && p.C == p.K && p.pad == 0 && p.stride == 1) return 0;
  const int taps = p.K / p.C;
  if (taps * p.C != p.K || (p.C % TC_BK) != 0 || p.lda != p.C || p.a_batch != (long long)p.Lin * p.C || (p.M % p.Lout) != 0) return -1;
  if (p.stride == 1) {
    if (p.Lout == p.Lin && (TC_BM % p.Lout) == 0) return 1;
    // long clips: one tile = 128 consecutive positions of one clip; TMA zero-fills positions outside [0, Lin), so any
    // left pad works (the 2-tap even / odd upsampling convs pad on one side only)
    if (p.Lout == p.Lin + 2 * p.pad - p.dil * (taps - 1) || p.Lout == p.Lin) return 2;
    return -1;
  }
  if (p.pad == 0 && p.dil == 1 && (long long)(p.Lout - 1) * p.stride + taps <= p.Lin) return 3;
  return -1;
}

static bool tc_fast_supported(const GemmP& p);
bool tc_supported(const GemmP& p) {
  if (p.attn) return tc_fast_supported(p);
  if ((!p.out && !p.o_planes) || p.out_scale != 1.0f || p.M < 1 || p.N < 16) return false;   // rows beyond M: TMA zero fill + masked stores
  const int mode = conv_mode(p);
  if (mode < 0) return false;
  if (mode == 0) return (p.K % TC_BK) == 0;
  // the flat strided view reads up to stride - 1 rows past the tensor: caller-owned planes must carry that slack
  if (mode == 3 && p.a_planes && p.a_plane_stride < ((long long)(p.M / p.Lout) * p.Lin + 16) * p.C) return false;
  return true;
}

static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, TcEpi ep, int num_kb, int mtiles, int BN, cudaStream_t s) {
  // ring depth: as many stages as fit beside the barriers in 227 KB (4 / 3 / 2 for BN = 64 / 128 / 192).  A tall grid of 64-wide
  // tiles (the WavEncoder convs: thousands of tiles, K = 15 taps) instead runs TWO CTAs per SM on half the ring: set-up and epilogue
  // of one tile then overlap the main loop of the other (the kernel has one tile per CTA and no other overlap between tiles).
  const int w_slot = 2 * BN * TC_BK * 2;
  const int stage_bytes = 2 * TC_A_PLANE + w_slot;
  const int ntiles = (ep.N + BN - 1) / BN;
  const bool two_per_sm = BN == 64 && (long long)mtiles * ntiles > 2 * 148;
  const int budget = two_per_sm ? 112 * 1024 : 227 * 1024 - 2048;
  int stages, ring;
  if (ep.mode == 4) {                     // staged taps: one or two row slots + a ring of weight tiles only
    ep.a_slots = ep.kb_per_tap == 1 ? 1 : 2;
    stages = (budget - ep.a_slots * TAPS_SLOT) / w_slot;
    stages = stages > 4 ? 4 : stages;
    ring = ep.a_slots * TAPS_SLOT + stages * w_slot;
  } else {
    ep.a_slots = 0;
    stages = budget / stage_bytes;
    stages = stages > 4 ? 4 : stages;
    ring = stages * stage_bytes;
  }
  if (stages < 2) { set_error("launch_tc: no room for a two-stage ring (BN=%d mode=%d)", BN, ep.mode); return ST_EINVAL; }
  const int epi_tile = (TC_BM * (BN + 4) + TC_BM) * 4;       // the epilogue's fp32 tile reuses the ring
  if (ring < epi_tile) { set_error("launch_tc: ring of %d B cannot hold the epilogue tile (%d B)", ring, epi_tile); return ST_EINVAL; }
  const int smem = ring + (2 * stages + 5) * 8 + 16 + 1024;
  static DeviceOnce attr;       // the attribute is per device (one process may drive models on several GPUs)
  if (attr.first()) ST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  dim3 grid(ntiles, mtiles);
  launch_k(gemm_tc_kernel, grid, dim3(TC_THREADS), (size_t)smem, s, tmA, tmW, ep, num_kb, BN, stages);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

// 3-D flat view for strided convs: planes [2][rows(+slack)][C] seen as {s*C, ceil(rows/s), 2}
static int get_map_strided(const CUtensorMap** out, const __half* base, long long plane_stride, long long rows, int C, int stride) {
  const int rows_v = (int)((rows + stride - 1) / stride);
  MapKey key{base, plane_stride, stride * C, rows_v, -7, TC_BM};
  std::lock_guard<std::mutex> lk(g_tc_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = &it->second; return ST_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ST_ECUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)stride * C, (cuuint64_t)rows_v, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)stride * C * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {TC_BK, TC_BM, 2};
  cuuint32_t est[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(strided rows=%lld C=%d s=%d) failed: %d", rows, C, stride, (int)r); return ST_ECUDA; }
  *out = &(g_maps[key] = m);
  return ST_OK;
}
// 4-D long-clip view {C, Lin, B, 2}, box {64, 128, 1, 2}
static int get_map_long(const CUtensorMap** out, const __half* base, long long plane_stride, int B, int Lin, int C, int box_rows = TC_BM) {
  MapKey key{base, plane_stride, C, Lin, B, -2 - box_rows};
  std::lock_guard<std::mutex> lk(g_tc_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = &it->second; return ST_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ST_ECUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Lin, (cuuint64_t)B, 2};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Lin * C * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[4] = {TC_BK, (cuuint32_t)box_rows, 1, 2};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(long B=%d Lin=%d C=%d) failed: %d", B, Lin, C, (int)r); return ST_ECUDA; }
  *out = &(g_maps[key] = m);
  return ST_OK;
}

// fp32 [rows, ld] row-major -> 2-D map {cols, rows}, box {32, 128}, 128B swizzle (residual prefetch and result store of
// the trunk kernel)
// row statistics [rows][16][2] fp32 (128 B per row): box = the 2 * nch floats a CTA with nch 32-column chunks owns x 128 rows, no swizzle
static int get_map_stats(const CUtensorMap** out, const float* base, int rows, int nch) {
  MapKey key{base, (long long)32, nch, rows, -10, 32};
  std::lock_guard<std::mutex> lk(g_tc_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = &it->second; return ST_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ST_ECUDA; }
  cuuint64_t gdim[2] = {32, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {(cuuint32_t)(2 * nch), TC_BM};
  cuuint32_t est[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(row statistics rows=%d nch=%d) failed: %d", rows, nch, (int)r); return ST_ECUDA; }
  *out = &(g_maps[key] = m);
  return ST_OK;
}

static int get_map_2d_f32(const CUtensorMap** out, const float* base, int rows, int cols, int ld) {
  MapKey key{base, (long long)ld, cols, rows, -9, 32};
  std::lock_guard<std::mutex> lk(g_tc_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = &it->second; return ST_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ST_ECUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, TC_BM};
  cuuint32_t est[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d f32 rows=%d cols=%d ld=%d) failed: %d", rows, cols, ld, (int)r); return ST_ECUDA; }
  *out = &(g_maps[key] = m);
  return ST_OK;
}

// Tile width of the trunk kernel: the candidate with the smallest modelled time  waves * (K blocks * cadence + epilogue),
// cadence and epilogue in cycles as measured per width (tests/probe_mainloop.py).  At M = 2048 this gives 64 / 128 / 192
// for N = 512 / 1024 / 1536 (one wave of 128 CTAs each); at M = 1024 it gives 64 for N = 1024 (128 CTAs instead of 64).
static int fast_bn(const GemmP& p) {
  if (p.attn) return 192;
  const int mt = (p.M + TC_BM - 1) / TC_BM;
  const int nkb = (p.K + TC_BK - 1) / TC_BK;
  const int cand[3] = {64, 128, 192};
  const double cadence[3] = {638.0, 776.0, 1150.0}, epi[3] = {2600.0, 4300.0, 6000.0};
  int best = 64;
  double best_t = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (p.N % bn != 0 && !(bn == 64)) continue;                 // 64 is always admissible (N is a multiple of 64)
    const long long ctas = (long long)mt * ((p.N + bn - 1) / bn);
    const double waves = (double)((ctas + 147) / 148);
    const double t = waves * (nkb * cadence[i] + epi[i] * (p.act == ACT_GELU ? 2.0 : 1.0) + 4000.0);
    if (t < best_t) { best_t = t; best = bn; }
  }
  return best;
}

// The trunk kernel takes plain Linear layers whose operand / result layouts TMA can describe.
static bool tc_fast_supported(const GemmP& p) {
  if (p.attn)
    return conv_mode(p) == 0 && p.K == 512 && p.N == 1536 && (p.M % 32) == 0 && !p.out && p.o_planes && p.o_planes_ld == 512 && p.ln_stats &&
           p.ln_s && p.ln_c && !p.res && p.act == ACT_NONE && !p.stats_out;
  const int cmode = conv_mode(p);
  if ((cmode != 0 && cmode != 1) || (p.K % TC_BK) != 0 || (p.N % 64) != 0 || p.out_scale != 1.0f) return false;
  if (p.a_relu && p.a_planes) return false;          // a ReLU on the input is applied when the operand is split, or by the producer
  if (p.a2_planes && (cmode != 0 || !p.a_planes || p.a2_K <= 0 || p.a2_K >= p.K || (p.a2_K % TC_BK) != 0)) return false;
  if (!p.out && !p.o_planes) return false;
  if (p.act != ACT_NONE && p.act != ACT_GELU && p.act != ACT_RELU) return false;
  if (p.res && (p.res != p.out || p.res_div != 1 || p.ldr != p.ldo || (p.res_mode == RES_PRE && p.act != ACT_NONE))) return false;
  if (p.out && ((p.ldo & 3) || (reinterpret_cast<uintptr_t>(p.out) & 15))) return false;
  if (p.o_planes && ((p.o_planes_ld & 7) || p.o_planes_ld < p.N || (reinterpret_cast<uintptr_t>(p.o_planes) & 15) || (p.o_plane_stride & 7))) return false;
  if (p.stats_out && p.N != 512) return false;
  if (p.ln_stats && (!p.ln_s || !p.ln_c)) return false;
  const int BN = fast_bn(p);
  const int stage = 2 * TC_A_PLANE + 2 * BN * TC_BK * 2;
  const int staging = (p.out ? (BN / 32) * FAST_BOX_F32 : 0) + (p.o_planes ? ((BN + 63) / 64) * FAST_BOX_PL : 0);
  const int res = p.res ? (BN / 32) * FAST_BOX_F32 : 0;
  int stages = (232448 - 4096 - res) / stage;
  if (stages > 4) stages = 4;
  return stages >= 2 && staging <= stages * stage && (BN % 64 == 0 || !p.o_planes);
}

// ---- layer chains (gemm_tc_chain_kernel).  Between fast_chain_begin and fast_chain_end the trunk launches of one stream are collected
// instead of launched; a run of layers that all have 8 column tiles and the same number of row tiles leaves as one cluster launch.
// Anything else that is launched in between flushes the run first, so the stream order of the caller's calls is kept.
struct ChainState {
  bool open = false;
  cudaStream_t s = nullptr;
  unsigned grid_y = 0;
  int seen = 0;                 // trunk layers since fast_chain_begin (ST_CHAIN_SKIP: the first ones stay single launches; for bisecting)
  std::vector<FastLayer> ly;
  std::vector<int> smem;        // single-launch shared memory of each pending layer (used when a run of one is flushed)
};
static thread_local ChainState g_chain;
bool g_tc_chain = false;       // st_debug_probe bit 131072 turns layer chaining ON (measured slower than one launch per layer: DESIGN.md)

static int launch_fast_single(const FastLayer& L, unsigned grid_y, int smem, bool legacy, cudaStream_t s) {
  static DeviceOnce attr;
  if (attr.first()) {
    ST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    ST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  dim3 grid((L.ep.N + L.BN - 1) / L.BN, grid_y);
  if (L.ep.attn) launch_k_cluster(legacy ? gemm_tc_fast_kernel<true> : gemm_tc_fast_kernel<false>, grid, dim3(TC_THREADS), (size_t)smem, s, 2, L.tmA, L.tmW, L.tmO, L.tmP, L.ep, L.num_kb, L.BN, L.STAGES, L.tmA2, L.tmS);
  else launch_k(legacy ? gemm_tc_fast_kernel<true> : gemm_tc_fast_kernel<false>, grid, dim3(TC_THREADS), (size_t)smem, s, L.tmA, L.tmW, L.tmO, L.tmP, L.ep, L.num_kb, L.BN, L.STAGES, L.tmA2, L.tmS);
  ST_CHECK_LAUNCH();
  return ST_OK;
}

int fast_chain_flush() {
  ChainState& c = g_chain;
  if (c.ly.empty()) return ST_OK;
  int r = ST_OK;
  if (c.ly.size() == 1) {
    r = launch_fast_single(c.ly[0], c.grid_y, c.smem[0], false, c.s);
  } else {
    static DeviceOnce attr;
    if (attr.first()) ST_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    if (getenv("ST_CHAIN_DEBUG")) {
      for (int cs = 2; cs <= 8; cs *= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8, c.grid_y); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = 232448;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, gemm_tc_chain_kernel, &cfg);
        fprintf(stderr, "chain kernel: cluster size %d -> max active clusters %d (%s)\n", cs, n, cudaGetErrorString(e));
      }
    }
    static thread_local FastChain P;               // 25 KB: passed by value, kept off the stack
    P.nl = (int)c.ly.size();
    for (int i = 0; i < P.nl; ++i) P.ly[i] = c.ly[i];
    launch_k_cluster(gemm_tc_chain_kernel, dim3(8, c.grid_y), dim3(TC_THREADS), (size_t)232448, c.s, 8, P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("gemm_tc_chain_kernel launch (%d layers): %s", P.nl, cudaGetErrorString(e));
      r = ST_ECUDA;
    } else {
      st::g_launches++;
    }
  }
  c.ly.clear();
  c.smem.clear();
  return r;
}
int fast_chain_begin(cudaStream_t s) {
  ST_TRY(fast_chain_flush());
  g_chain.open = g_tc_chain;
  g_chain.s = s;
  g_chain.seen = 0;
  return ST_OK;
}
int fast_chain_end() {
  const int r = fast_chain_flush();
  g_chain.open = false;
  return r;
}

int tc_fast_grid_x(const GemmP& p) {
  if (!(g_tc_fast || p.attn) || !tc_supported(p) || !tc_fast_supported(p)) return 0;
  const int BN = fast_bn(p);
  return (p.N + BN - 1) / BN;
}

static int gemm_tc_fast(const GemmP& p, WPlanes* w, const __half* planes, long long pstride, cudaStream_t s) {
  const int BN = fast_bn(p);
  const int stage = 2 * TC_A_PLANE + 2 * BN * TC_BK * 2;
  const int res = p.res ? (BN / 32) * FAST_BOX_F32 : 0;
  int stages = (232448 - 4096 - res) / stage;
  if (stages > 4) stages = 4;
  if (p.attn) { stages = 2; }
  const int smem = stages * stage + (p.attn ? ATT_SX_BYTES : res) + 2 * 192 * 4 + (2 * stages + 2 + 5) * 8 + 16 + 1024 + (p.stats_out ? FAST_STATS + 128 : 0);
  if (smem > 232448) { set_error("gemm_tc_fast: %d bytes of shared memory (BN=%d stages=%d)", smem, BN, stages); return ST_EINVAL; }
  const CUtensorMap *tmA = nullptr, *tmW = nullptr, *tmO = nullptr, *tmP = nullptr;
  const int cmode = conv_mode(p);
  const CUtensorMap* tmA2 = nullptr;
  if (cmode == 0) ST_TRY(get_map_3d(&tmA, planes, pstride, p.M, p.K - p.a2_K, TC_BM));
  else ST_TRY(get_map_4d(&tmA, planes, pstride, p.M / p.Lout, p.Lin, p.C));
  if (p.a2_planes) ST_TRY(get_map_3d(&tmA2, p.a2_planes, p.a2_plane_stride, p.M, p.a2_K, TC_BM));
  else tmA2 = tmA;
  ST_TRY(get_map_3d(&tmW, w->planes, (long long)p.N * w->Kp, p.N, w->Kp, BN));
  tmO = tmA; tmP = tmA;
  if (p.out) ST_TRY(get_map_2d_f32(&tmO, p.out, p.M, p.N, p.ldo));
  if (p.o_planes) ST_TRY(get_map_3d(&tmP, p.o_planes, p.o_plane_stride, p.M, p.attn ? 512 : p.N, TC_BM, p.o_planes_ld));
  FastEpi ep;
  ep.bias = p.ln_stats ? p.ln_c : p.bias;
  ep.ln_s = p.ln_stats ? p.ln_s : nullptr;
  ep.ln_stats = p.ln_stats; ep.stats_out = p.stats_out;
  ep.scale = w->inv_scale / kActScale;
  ep.act = p.act; ep.has_res = p.res ? 1 : 0; ep.has_out = p.out ? 1 : 0; ep.has_planes = p.o_planes ? 1 : 0;
  ep.M = p.M; ep.N = p.N; ep.dbg = (g_tc_dbg_n == 0 || (g_tc_dbg_n == p.N && g_tc_dbg_k == p.K)) ? g_tc_dbg : (g_tc_dbg && g_tc_dbg_n2 == p.N && g_tc_dbg_k2 == p.K) ? g_tc_dbg + 1024 : nullptr; ep.probe = g_tc_probe; ep.attn = p.attn;
  ep.mode = cmode; ep.T = p.Lout; ep.kb_per_tap = cmode == 0 ? 1 : p.C / TC_BK; ep.dil = p.dil; ep.pad = p.pad;
  ep.planes_relu = p.o_planes_relu;
  ep.kb_split = p.a2_planes ? (p.K - p.a2_K) / TC_BK : 0x7fffffff;
  ep.sig_in = p.sig_in; ep.sig_expect = p.sig_expect; ep.sig_out = p.sig_out;
  const bool legacy = p.attn == 1 || (ep.probe & 16);
  dim3 grid((p.N + BN - 1) / BN, (p.M + TC_BM - 1) / TC_BM);
  FastLayer L;
  const CUtensorMap* tmS = tmA;
  if (p.stats_out) ST_TRY(get_map_stats(&tmS, p.stats_out, p.M, BN / 32));
  L.tmA = *tmA; L.tmW = *tmW; L.tmO = *tmO; L.tmP = *tmP; L.tmA2 = *tmA2; L.tmS = *tmS;
  L.ep = ep; L.num_kb = w->Kp / TC_BK; L.BN = BN; L.STAGES = stages; L.pad_ = 0;
  ChainState& c = g_chain;
  // a chain layer: 8 column tiles (= the cluster), full row tiles, plain rows, and room for the fixed tail block behind ring + residual
  static const int chain_skip = [] { const char* e = getenv("ST_CHAIN_SKIP"); return e ? atoi(e) : 0; }();
  const bool chainable = c.open && c.seen++ >= chain_skip && s == c.s && !legacy && grid.x == 8 && cmode == 0 && (p.M % TC_BM) == 0 &&
                         stages * stage + (p.attn ? ATT_SX_BYTES : res) + 1024 + FAST_TAIL <= 232448;
  if (chainable && (c.ly.empty() || c.grid_y == grid.y)) {
    c.grid_y = grid.y;
    c.ly.push_back(L);
    c.smem.push_back(smem);
    static const int chain_max = [] { const char* e = getenv("ST_CHAIN_MAX"); const int v = e ? atoi(e) : FAST_CHAIN_MAX; return v < 1 ? 1 : v > FAST_CHAIN_MAX ? FAST_CHAIN_MAX : v; }();
    if ((int)c.ly.size() >= chain_max) return fast_chain_flush();   // (ST_CHAIN_MAX: shorter chains, for bisecting)
    return ST_OK;
  }
  ST_TRY(fast_chain_flush());
  if (chainable) {                                   // a run with a different number of row tiles starts
    c.grid_y = grid.y;
    c.ly.push_back(L);
    c.smem.push_back(smem);
    return ST_OK;
  }
  return launch_fast_single(L, grid.y, smem, legacy, s);
}

// bytes of split scratch a GEMM needs (0: its operand already arrives as planes)
size_t tc_scratch_need(const GemmP& p) {
  if (p.a_planes) return 0;
  const int mode = conv_mode(p);
  const int Ka = mode == 0 ? p.K : p.C;
  const long long rows = mode == 0 ? p.M : (long long)(p.M / p.Lout) * p.Lin;
  return (size_t)2 * (rows + 16) * Ka * sizeof(__half) + 1024;
}
// the stream's arena, grown to `need` (never during graph capture: captured launches keep the pointer)
int tc_scratch_reserve(cudaStream_t s, size_t need, Arena** out) {
  Arena* sc = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    sc = &g_scratch[s];                                   // std::map nodes are stable: the pointer survives later insertions
  }
  if (need > sc->cap) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (cap != cudaStreamCaptureStatusNone) { set_error("split scratch must be sized before graph capture"); return ST_ESTATE; }
    ST_TRY(sc->reserve(need + need / 2));
    ST_CHECK_CUDA(cudaMemsetAsync(sc->base, 0, sc->cap, s));
  }
  if (out) *out = sc;
  return ST_OK;
}
// a stream that is about to be destroyed gives its arena back (stream handles are reused by the driver)
void tc_scratch_release(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  auto it = g_scratch.find(s);
  if (it != g_scratch.end()) { it->second.release(); g_scratch.erase(it); }
}

int gemm_tc(const GemmP& p, cudaStream_t s) {
  if (!tc_supported(p)) { set_error("gemm_tc: unsupported problem"); return ST_EUNSUPPORTED; }
  WPlanes* w = nullptr;
  ST_TRY(get_wplanes(p, s, &w));
  const int mode = conv_mode(p);
  const int Ka = mode == 0 ? p.K : p.C;                   // columns of the activation planes
  const int nclips = mode == 0 ? 1 : p.M / p.Lout;
  const long long rows = mode == 0 ? p.M : (long long)nclips * p.Lin;   // rows of the activation tensor
  const __half* planes = p.a_planes;
  long long pstride = p.a_plane_stride;
  if (!planes) {
    // operand still fp32: split it into the scratch planes first (stream order serialises reuse of the scratch)
    pstride = (rows + 16) * Ka;                           // slack rows keep the strided flat view inside the buffer
    Arena* sc = nullptr;
    ST_TRY(tc_scratch_reserve(s, tc_scratch_need(p), &sc));
    __half* sp = reinterpret_cast<__half*>(sc->base);
    const long long n4 = rows * (Ka >> 2);
    ST_TRY(fast_chain_flush());
    launch_k(split_planes_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, p.A, p.lda, (int)rows, Ka, Ka, kActScale, p.a_relu, sp, pstride);
    ST_CHECK_LAUNCH();
    planes = sp;
  }
  if ((g_tc_fast || p.attn) && tc_fast_supported(p)) return gemm_tc_fast(p, w, planes, pstride, s);
  ST_TRY(fast_chain_flush());
  if (p.a2_planes) { set_error("gemm_tc: a two-tensor operand is served by the trunk kernel only"); return ST_EUNSUPPORTED; }
  if (p.attn) { set_error("gemm_tc: fused attention needs the 1536 x 512 qkv layout"); return ST_EUNSUPPORTED; }
  if (p.ln_stats || p.stats_out) { set_error("gemm_tc: folded LayerNorm is served by the trunk kernel only"); return ST_EUNSUPPORTED; }
  const CUtensorMap* tmA = nullptr;
  const CUtensorMap* tmW = nullptr;
  TcEpi ep;
  // a stride-1 conv over long clips whose taps reach at most 24 rows beyond the tile stages its rows once per channel block and
  // shifts the operand descriptor per tap (mode 4) instead of fetching the same rows once per tap (st_debug_probe bit 8192: off)
  const int taps = mode == 0 ? 1 : p.K / p.C;
  const bool staged = g_tc_taps && mode == 2 && taps > 1 && (taps - 1) * p.dil + TC_BM <= TAPS_ROWS;
  ep.mode = staged ? 4 : mode; ep.T = p.Lout; ep.kb_per_tap = mode == 0 ? 1 : p.C / TC_BK; ep.dil = p.dil; ep.pad = p.pad; ep.stride = p.stride;
  ep.C = p.C; ep.Lin = p.Lin; ep.Lout = p.Lout; ep.tpc = (p.Lout + TC_BM - 1) / TC_BM;
  if (mode == 0) ST_TRY(get_map_3d(&tmA, planes, pstride, (int)rows, Ka, TC_BM));
  else if (mode == 1) ST_TRY(get_map_4d(&tmA, planes, pstride, nclips, p.Lin, p.C));
  else if (mode == 2) ST_TRY(get_map_long(&tmA, planes, pstride, nclips, p.Lin, p.C, staged ? TAPS_ROWS : TC_BM));
  else ST_TRY(get_map_strided(&tmA, planes, pstride, rows, p.C, p.stride));
  // tile width: keep the grid at one wave of <= 148 CTAs (1 CTA per SM) whenever the shape allows it
  const int mt = mode >= 2 ? nclips * ((p.Lout + TC_BM - 1) / TC_BM) : (p.M + TC_BM - 1) / TC_BM;
  int BN = p.N <= 512 ? 64 : 128;
  if (p.N % 192 == 0 && mt * (p.N / 128) > 148 && mt * (p.N / 192) <= 148) BN = 192;
  ST_TRY(get_map_3d(&tmW, w->planes, (long long)p.N * w->Kp, p.N, w->Kp, BN));
  ep.out = p.out; ep.bias = p.bias; ep.res = p.res;
  ep.planes = p.o_planes; ep.plane_stride = p.o_plane_stride; ep.ld_planes = p.o_planes_ld; ep.planes_scale = kActScale;
  ep.planes_relu = p.o_planes_relu; ep.ldo = p.ldo; ep.ldr = p.ldr; ep.res_mode = p.res ? p.res_mode : RES_NONE; ep.res_div = p.res_div; ep.act = p.act;
  ep.scale = w->inv_scale / kActScale;
  ep.M = p.M; ep.N = p.N;
  ep.ln_stats = p.ln_stats; ep.ln_s = p.ln_s; ep.ln_c = p.ln_c; ep.stats_out = p.stats_out;
  if (p.stats_out && (p.N != 512 || BN != 64)) { set_error("gemm_tc: stats_out needs N = 512 (8 tiles of 64 columns)"); return ST_EINVAL; }
  if (p.ln_stats && (!p.ln_s || !p.ln_c)) { set_error("gemm_tc: ln_stats needs ln_s and ln_c"); return ST_EINVAL; }
  ep.dbg = g_tc_dbg;
  ep.probe = g_tc_probe;
  const int num_kb = w->Kp / TC_BK;
  const int mtiles = mode >= 2 ? nclips * ep.tpc : (p.M + TC_BM - 1) / TC_BM;
  return launch_tc(*tmA, *tmW, ep, num_kb, mtiles, BN, s);
}

int set_trace_tc(unsigned long long* p) {
  ST_CHECK_CUDA(cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)));
  return ST_OK;
}

}  // namespace st
