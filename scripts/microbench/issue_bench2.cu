// Micro-benchmark 2: per-step overheads of the single-thread MMA-issuer and TMA-producer loops.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace osvos;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// variant bits: 1 = wait on an already-completed barrier each step, 2 = commit each step, 4 = fence_after each step
template <int N>
__global__ void __launch_bounds__(64, 1) mma_step_kernel(int iters, int mmas_per_step, int variant, long long* cycles, int nacc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar_done, bar_ready, bar_sink;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_ready, 1); mbar_init(&bar_sink, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_smem();
  const uint32_t tm = tmem_slot;
  long long t0 = clock64();
  if (threadIdx.x < 32) { if (elect_one()) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, true);
    const uint32_t a = smem_u32(smem), b = a + 16384;
    for (int i = 0; i < iters; ++i) {
      if (variant & 1) mbar_wait(&bar_ready, 1);          // parity of the "previous" phase: completes immediately
      if (variant & 4) tc_fence_after();
      const uint64_t da = make_smem_desc(a, 16, 1024, kLayoutSW128);
      const uint64_t db = make_smem_desc(b, 16, 1024, kLayoutSW128);
      for (int k = 0; k < mmas_per_step; ++k) umma_f16(tm + (k % nacc) * N, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
      if (variant & 2) umma_commit(&bar_sink);
    }
    umma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
  } __syncwarp(); }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// producer-only: S-stage ring, consumer thread arrives immediately. tmas_per_step boxes per expect_tx.
// variant bit 1: second producer thread (warp 1... thread 64) handles odd steps.
__global__ void __launch_bounds__(128, 1)
tma_step_kernel(const __grid_constant__ CUtensorMap map, int stages, int box_bytes, int tmas_per_step, int iters,
                int rows_total, int box_rows, int producers, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = box_bytes * tmas_per_step;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  long long t0 = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0 && warp < producers) {
    int row = (blockIdx.x * 977) % (rows_total - box_rows);
    for (int i = warp; i < iters; i += producers) {
      const int st = i % stages; const uint32_t ph = (i / stages) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&full[st], stage_bytes);
      for (int t = 0; t < tmas_per_step; ++t) {
        tma_load_3d(&map, &full[st], smem + st * stage_bytes + t * box_bytes, 0, row, 0);
        row += box_rows; if (row + box_rows > rows_total) row = 0;
      }
    }
  } else if (lane == 0 && warp == 3) {
    for (int i = 0; i < iters; ++i) {
      const int st = i % stages; const uint32_t ph = (i / stages) & 1;
      mbar_wait(&full[st], ph);
      mbar_arrive(&empty[st]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  long long* cyc; CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
  std::vector<long long> h(1024);
  const int smem = 64 * 1024;
  printf("== MMA issuer step cost (M=128 N=128 K=16), cycles per step; variants: w=wait(ready) c=commit f=fence\n");
  CK(cudaFuncSetAttribute(mma_step_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(mma_step_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(mma_step_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int n : {128, 64, 16}) for (int nacc : {1, 2, 4}) for (int variant : {0, 3}) {
    const int iters = 1000, mm = 12;
    if (n * nacc > 512) continue;
    for (int rep = 0; rep < 2; ++rep) {
      if (n == 128) mma_step_kernel<128><<<148, 64, smem>>>(iters, mm, variant, cyc, nacc);
      else if (n == 64) mma_step_kernel<64><<<148, 64, smem>>>(iters, mm, variant, cyc, nacc);
      else mma_step_kernel<16><<<148, 64, smem>>>(iters, mm, variant, cyc, nacc);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N %3d nacc %d variant %s%s : %7.1f cycles/step (%.1f per MMA)\n", n, nacc, (variant & 1) ? "w" : "-",
           (variant & 2) ? "c" : "-", (double)mx / iters, (double)mx / iters / mm);
  }
  return 0;
  printf("== TMA producer step cost: cycles per step and B/clk/SM\n");
  const int rows_total = 1 << 14;
  void* buf; CK(cudaMalloc(&buf, (size_t)rows_total * 128)); CK(cudaMemset(buf, 0, (size_t)rows_total * 128));
  for (int box_rows : {32, 128}) for (int tmas : {1, 2, 4}) for (int producers : {1, 2}) for (int stages : {3, 6}) {
    const int box_bytes = box_rows * 128;
    if (stages * box_bytes * tmas > 200 * 1024) continue;
    CUtensorMap m;
    cuuint64_t dims[3] = {64, (cuuint64_t)rows_total, 1}; cuuint64_t strides[2] = {128, (cuuint64_t)rows_total * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1}; cuuint32_t es[3] = {1, 1, 1};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    const int iters = 2000;
    const int sm = stages * box_bytes * tmas + 1024 + 256;
    CK(cudaFuncSetAttribute(tma_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    for (int rep = 0; rep < 2; ++rep) { tma_step_kernel<<<148, 128, sm>>>(m, stages, box_bytes, tmas, iters, rows_total, box_rows, producers, cyc); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("box %5d B x%d per step, producers %d, stages %d : %7.1f cycles/step, %6.1f B/clk/SM\n", box_bytes, tmas, producers,
           stages, (double)mx / iters, (double)iters * box_bytes * tmas / mx);
  }
  return 0;
}
