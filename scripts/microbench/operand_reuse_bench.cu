// Micro-benchmark: what does the ~85-cycle floor of a tcgen05.mma (M=128, K=16, cta_group::1) depend on?
// tma_mma_bench.cu measured 84 cycles per MMA when every instruction names new operands and 42-64 when the same
// descriptors are repeated.  This one separates the cases: same A / new B, new A / same B, the N-concatenated
// split-B scheme the conv kernels use (A_hi x [B_hi|B_lo] then A_lo x B_hi), the plain three-pass orders, and
// grouping the passes across the K steps.  Output: cycles per K step (all MMAs of one 16-channel step).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../osvos_pytorch_b200/csrc -o operand_reuse_bench operand_reuse_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace osvos;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum Pattern { kDiffADiffB, kSameADiffB, kDiffASameB, kSameSame, kConcat, kThreePass, kThreePassAGrouped, kSingle,
               kConcatKGrouped, kConcatLoFirst, kNumPatterns };
static const char* kNames[kNumPatterns] = {
    "2 MMAs: (A0,B0) (A1,B1)              new A, new B",
    "2 MMAs: (A0,B0) (A0,B1)              same A, new B",
    "2 MMAs: (A0,B0) (A1,B0)              new A, same B",
    "2 MMAs: (A0,B0) (A0,B0)              same A, same B",
    "concat: (A0,[B0|B1],2N) (A1,B0,N)    conv kernels today",
    "3-pass: (A1,B0) (A0,B1) (A0,B0)      first version",
    "3-pass: (A0,B0) (A0,B1) (A1,B0)      A-grouped",
    "1 MMA : (A0,B0)",
    "concat, passes grouped over the 4 K steps: 4 x (A0,[B0|B1],2N) then 4 x (A1,B0,N)",
    "concat, lo first: (A1,B0,N) (A0,[B0|B1],2N)",
};

template <int N, int P>
__global__ void __launch_bounds__(64, 1) pattern_kernel(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_smem();
  const uint32_t tm = tmem_slot;
  long long t0 = clock64();
  if (threadIdx.x < 32) {
    constexpr uint32_t idN = make_idesc_f16(128, N, true);
    constexpr uint32_t id2N = make_idesc_f16(128, 2 * N <= 256 ? 2 * N : 256, true);
    const uint32_t a = smem_u32(smem);
    const uint64_t A0 = make_smem_desc(a, 16, 1024, kLayoutSW128);
    const uint64_t A1 = make_smem_desc(a + 16384, 16, 1024, kLayoutSW128);
    const uint64_t B0 = make_smem_desc(a + 32768, 16, 1024, kLayoutSW128);            // [B0 | B1] contiguous: 2N rows
    const uint64_t B1 = make_smem_desc(a + 32768 + N * 128, 16, 1024, kLayoutSW128);
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
        if (P == kConcatKGrouped) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tm, A0 + 2 * k, B0 + 2 * k, id2N, 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tm, A1 + 2 * k, B0 + 2 * k, idN, 1);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a0 = A0 + 2 * k, a1 = A1 + 2 * k, b0 = B0 + 2 * k, b1 = B1 + 2 * k;
            if (P == kDiffADiffB) { umma_f16(tm, a0, b0, idN, 1); umma_f16(tm, a1, b1, idN, 1); }
            if (P == kSameADiffB) { umma_f16(tm, a0, b0, idN, 1); umma_f16(tm, a0, b1, idN, 1); }
            if (P == kDiffASameB) { umma_f16(tm, a0, b0, idN, 1); umma_f16(tm, a1, b0, idN, 1); }
            if (P == kSameSame) { umma_f16(tm, a0, b0, idN, 1); umma_f16(tm, a0, b0, idN, 1); }
            if (P == kConcat) { umma_f16(tm, a0, b0, id2N, 1); umma_f16(tm, a1, b0, idN, 1); }
            if (P == kConcatLoFirst) { umma_f16(tm, a1, b0, idN, 1); umma_f16(tm, a0, b0, id2N, 1); }
            if (P == kThreePass) { umma_f16(tm, a1, b0, idN, 1); umma_f16(tm, a0, b1, idN, 1); umma_f16(tm, a0, b0, idN, 1); }
            if (P == kThreePassAGrouped) { umma_f16(tm, a0, b0, idN, 1); umma_f16(tm, a0, b1, idN, 1); umma_f16(tm, a1, b0, idN, 1); }
            if (P == kSingle) { umma_f16(tm, a0, b0, idN, 1); }
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int P>
static void run(long long* cyc, std::vector<long long>& h) {
  const int smem = 100 * 1024;
  auto kern = pattern_kernel<N, P>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) { kern<<<148, 64, smem>>>(iters, cyc); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N %3d  %7.1f cycles per K step   %s\n", N, (double)mx / (iters * 4.0), kNames[P]);
}

// Two issuing warps, each streaming its own MMAs (own operands, own accumulator columns): is the ~85-cycle floor per
// ISSUER (then two issuers double the rate) or per SM?  mode 0: both warps 1 MMA per K step (A0,B0 | A1,B1);
// mode 1: warp 0 issues the N-concatenated pass (A0,[B0|B1],2N), warp 1 the lo pass (A1,B0,N) into other columns.
template <int N>
__global__ void __launch_bounds__(64, 1) dual_issuer_kernel(int iters, int mode, int issuers, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_smem();
  const uint32_t tm = tmem_slot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp < issuers) {
    constexpr uint32_t idN = make_idesc_f16(128, N, true);
    constexpr uint32_t id2N = make_idesc_f16(128, 2 * N <= 256 ? 2 * N : 256, true);
    const uint32_t a = smem_u32(smem);
    const uint64_t A = make_smem_desc(a + warp * 16384, 16, 1024, kLayoutSW128);
    const uint64_t B = make_smem_desc(a + 32768 + (mode == 0 ? warp * N * 128 : 0), 16, 1024, kLayoutSW128);
    const uint32_t idesc = (mode == 1 && warp == 0) ? id2N : idN;
    const uint32_t d = tm + warp * 256;
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, A + 2 * k, B + 2 * k, idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N>
static void run_dual(long long* cyc, std::vector<long long>& h) {
  const int smem = 100 * 1024;
  auto kern = dual_issuer_kernel<N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int issuers = 1; issuers <= 2; ++issuers) {
      for (int rep = 0; rep < 2; ++rep) { kern<<<148, 64, smem>>>(iters, mode, issuers, cyc); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("N %3d  %7.1f cycles per K step   %d issuing warp(s), %s\n", N, (double)mx / (iters * 4.0), issuers,
             mode == 0 ? "each 1 MMA of N per K step (own A, B, accumulator)"
                       : "warp 0: (A0,[B0|B1],2N), warp 1: (A1,B0,N) into separate accumulator columns");
    }
}

template <int N>
static void run_all(long long* cyc, std::vector<long long>& h) {
  run<N, kSingle>(cyc, h);
  run<N, kDiffADiffB>(cyc, h);
  run<N, kSameADiffB>(cyc, h);
  run<N, kDiffASameB>(cyc, h);
  run<N, kSameSame>(cyc, h);
  run<N, kConcat>(cyc, h);
  run<N, kConcatLoFirst>(cyc, h);
  run<N, kConcatKGrouped>(cyc, h);
  run<N, kThreePass>(cyc, h);
  run<N, kThreePassAGrouped>(cyc, h);
}

int main() {
  long long* cyc; CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
  std::vector<long long> h(1024);
  printf("== tcgen05.mma operand-reuse patterns, M=128 K=16 bf16, one issuing thread, 148 CTAs (max over CTAs)\n");
  printf("   ideal tensor time per K step: N/2 cycles per N-wide MMA (8192 dense bf16 flop/clk/SM)\n");
  run_all<64>(cyc, h);
  run_all<128>(cyc, h);
  printf("== one vs two issuing warps per CTA\n");
  run_dual<64>(cyc, h);
  run_dual<128>(cyc, h);
  return 0;
}
