// GPU probe (not a test, not product code): issue cadence of the split-fp16 MMA patterns of the trunk kernel, one CTA per SM
// (cta_group::1) against CTA pairs (cta_group::2, M = 256, each CTA holding half of the B operand).  No TMA: the operand
// tiles are filled once by the threads, the MMA thread issues NKB K blocks back to back.  Also checks the accumulators of
// the pair mode against the closed form, so that the operand split across the pair is pinned before the product uses it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/umma_probe.bin tests/umma_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.u32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
template <int G>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (G == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) if (clock64() - t0 > 2000000000LL) __trap();
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
      : "r"(taddr) : "memory");
}

// pattern: 0  a_hi.[w_hi;w_lo] (N = 2 BN) + a_lo.w_hi (N = BN)     (the product's concatenated form)
//          1  only the wide MMA      2  only the narrow MMA     3  three MMAs of width BN (the BN = 192 form)
// G = 2: the pair's B operand of the wide MMA is [w_hi (rank 0) ; w_lo (rank 1)]; the narrow MMA reads a second region that
// holds rows [0, BN/2) of w_hi in rank 0 and rows [BN/2, BN) in rank 1.
constexpr int STAGES = 2;
template <int G, int pattern>
__global__ void __launch_bounds__(192, 1) probe(int BN, int nkb, int M, long long* cycles, int* errs, int delay) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar, ring_bar[4], ready_bar;
  __shared__ uint32_t tmem_slot;
  uint32_t rank = 0;
  if (G == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5;
  const int A_PLANE = 128 * 128;                 // 128 rows x 64 halfs
  const int wide_rows = (pattern == 3 ? BN : 2 * BN) / G;  // (patterns 4..9 are timing only)     // B rows this CTA holds for the wide MMA
  const int narrow_rows = BN / G;
  const int STAGE = 2 * A_PLANE + 3 * BN * 128;  // A_hi, A_lo, wide region (<= 2 BN rows), narrow region (<= BN rows)
  // fill: A rows hold (1 + rank); B row n of the wide operand holds ((n % 7) + 1), of the narrow operand ((n % 5) + 1)
  for (int s = 0; s < STAGES; ++s) {
    __half* a = reinterpret_cast<__half*>(smem + s * STAGE);
    for (int i = threadIdx.x; i < 2 * 128 * 64; i += blockDim.x) a[i] = __float2half(1.0f + rank);
    __half* bw = reinterpret_cast<__half*>(smem + s * STAGE + 2 * A_PLANE);
    for (int i = threadIdx.x; i < wide_rows * 64; i += blockDim.x) bw[i] = __float2half((float)(((int)rank * wide_rows + i / 64) % 7 + 1));
    __half* bn = reinterpret_cast<__half*>(smem + s * STAGE + 2 * A_PLANE + 2 * BN * 128);
    for (int i = threadIdx.x; i < narrow_rows * 64; i += blockDim.x) bn[i] = __float2half((float)(((int)rank * narrow_rows + i / 64) % 5 + 1));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring_bar[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ready_bar)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar)) : "memory");    // phase 0 complete for good
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (G == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (G == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __reduce_or_sync(0xffffffffu, tmem_slot);
  long long t0 = 0, t1 = 0;
  if (warp == 1 && rank == 0) {
    if (elect_one()) {
      const uint32_t id_w = idesc_f16(M * G, pattern == 3 ? BN : 2 * BN), id_n = idesc_f16(M * G, BN);
      t0 = clock64();
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t base = smem_u32(smem + (kb % STAGES) * STAGE);
        const uint32_t a_hi = base, a_lo = base + A_PLANE, w_w = base + 2 * A_PLANE, w_n = w_w + 2 * BN * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = desc_sw128(a_hi) + 2 * k, dal = desc_sw128(a_lo) + 2 * k;
          const uint64_t dww = desc_sw128(w_w) + 2 * k, dwn = desc_sw128(w_n) + 2 * k;
          const uint32_t acc = (kb | k) != 0;
          if (pattern == 0) { umma<G>(tmem, dah, dww, id_w, acc); umma<G>(tmem + 2 * BN, dal, dwn, id_n, acc); }
          else if (pattern == 1) umma<G>(tmem, dah, dww, id_w, acc);
          else if (pattern == 2) umma<G>(tmem + 2 * BN, dal, dwn, id_n, acc);
          else if (pattern == 3) { umma<G>(tmem, dah, dww, id_w, acc); umma<G>(tmem + BN, dah, dwn, id_n, acc); umma<G>(tmem + BN, dal, dww, id_w, 1); }
          else if (pattern == 4) umma<G>(tmem, dal, dwn, id_n, acc);                 // narrow, D at column 0
          else if (pattern == 5) umma<G>(tmem + 2 * BN, dah, dwn, id_n, acc);        // narrow, A = a_hi
          else if (pattern == 6) umma<G>(tmem + 2 * BN, dal, dww, id_n, acc);        // narrow, B = start of the wide region
          else if (pattern >= 10) { umma<G>(tmem, dah, dww, id_w, acc); umma<G>(tmem + BN, dal, dww, id_n, 1); }
          else if (pattern == 7) { umma<G>(tmem, dah, dww, id_w, acc); umma<G>(tmem + BN, dal, dww, id_n, 1); }   // the product's form (G = 1)
          else if (pattern == 8) umma<G>(tmem + ((kb & 1) ? 2 * BN : 0), dah, dww, id_n, (kb >> 1 | k) != 0);    // narrow, two accumulators alternating per K block
          else if (pattern == 9) umma<G>(tmem + ((k & 1) ? 2 * BN : 0), dah, dww, id_n, (kb | (k >> 1)) != 0);   // narrow, alternating per MMA
        }
        if (delay > 0) { const long long c0 = clock64(); while (clock64() - c0 < delay) {} }
        if (pattern == 10 || pattern == 12) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ring_bar[kb & 3])) : "memory");
        if (pattern == 11 || pattern == 12) {
          if (!mbar_try(&ready_bar, 0)) __trap();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
      if (G == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
      else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                        ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
      mbar_wait(&done_bar, 0);
      t1 = clock64();
      if (blockIdx.x == 0) { cycles[0] = t1 - t0; }
    }
    __syncwarp();
  }
  // check (pattern 0, M = 128 only): D_wide[m][n] = (1 + rank) ((n % 7) + 1) 64 nkb ; D_narrow[m][n] = (1 + rank) ((n % 5) + 1) 64 nkb
  if (warp >= 2) {
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (pattern == 0 && M == 128) {
      const int lg = warp & 3;
      int bad = 0;
      for (int c = 0; c < 3 * BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) {
          const int n = c + j;
          const float want = (1.0f + rank) * 64.0f * nkb * (n < 2 * BN ? (float)(n % 7 + 1) : (float)((n - 2 * BN) % 5 + 1));
          if (__uint_as_float(v[j]) != want) ++bad;
        }
      }
      if (bad) atomicAdd(errs, bad);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (G == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
  long long* d_cyc; int* d_err;
  CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_err, 4));
  const int smem = STAGES * (2 * 128 * 128 + 3 * 128 * 128) + 2048;     // BN <= 128
#define ROW(G) {(void*)probe<G, 0>, (void*)probe<G, 1>, (void*)probe<G, 2>, (void*)probe<G, 3>, (void*)probe<G, 4>, (void*)probe<G, 5>, (void*)probe<G, 6>, (void*)probe<G, 7>, (void*)probe<G, 8>, (void*)probe<G, 9>, (void*)probe<G, 10>, (void*)probe<G, 11>, (void*)probe<G, 12>}
  void* table[2][13] = {ROW(1), ROW(2)};
  for (int g = 0; g < 2; ++g) for (int q = 0; q < 13; ++q) CK(cudaFuncSetAttribute(table[g][q], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int nkb = 64;
  const char* pname[13] = {"wide + narrow", "wide only", "narrow only", "3 x BN", "narrow D@0", "narrow a_hi", "narrow B=wide", "product form", "narrow 2acc/kb", "narrow 2acc/mma", "product+commit", "product+trywait", "product+both"};
  for (int M : {128})
    for (int BN : {64, 128})
      for (int pattern = 0; pattern < 13; ++pattern)
        for (int G = 1; G <= 2; ++G) for (int delay = 0; delay <= 500; delay += 100) {
          if (delay > 0 && (pattern != 7 || G != 1)) continue;
          if (M == 64 && (G == 2 || pattern > 3)) continue;
          if (G == 2 && pattern > 3) continue;
          if (pattern != 3 && 3 * BN > 512) continue;
          if (pattern >= 4 && pattern <= 9 && pattern != 7) continue;
          if (G == 2 && BN / 2 < 16) continue;
          CK(cudaMemset(d_err, 0, 4));
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(128); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          for (int rep = 0; rep < 2; ++rep) {
            void* fn = table[G - 1][pattern];
            void* args[] = {&BN, (void*)&nkb, &M, &d_cyc, &d_err, &delay};
            CK(cudaLaunchKernelExC(&cfg, fn, args));
            CK(cudaDeviceSynchronize());
          }
          long long cyc; int err;
          CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
          printf("M=%3d BN=%3d %-14s cta_group::%d  delay %3d  %7.1f cycles per K block (64 deep)  mismatches %d\n", M * G, BN, pname[pattern], G,
                 delay, (double)cyc / nkb, err);
        }
  return 0;
}
