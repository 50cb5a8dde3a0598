// Micro-benchmarks that bound the conv kernel's design space on B200:
//  (1) per-SM TMA ingest (global/L2 -> smem) as a function of ring depth and box size;
//  (2) tcgen05.mma issue rate for M=128, N in {16,64,128,256}, K=16 (no loads).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../osvos_pytorch_b200/csrc -o tma_mma_bench tma_mma_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace osvos;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ring of `stages` buffers of `box_bytes`; thread 0 produces, thread 32 consumes (arrives immediately)
__global__ void __launch_bounds__(64, 1)
tma_ingest_kernel(const __grid_constant__ CUtensorMap map, int stages, int box_bytes, int iters, int rows_total,
                  int box_rows, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * box_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    int row = (blockIdx.x * 977) % (rows_total - box_rows);
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&full[st], box_bytes);
      tma_load_3d(&map, &full[st], smem + st * box_bytes, 0, row, 0);
      row += box_rows; if (row + box_rows > rows_total) row = 0;
      if (++st == stages) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[st], ph);
      mbar_arrive(&empty[st]);
      if (++st == stages) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int N>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int iters, int passes, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_smem();
  const uint32_t tm = tmem_slot;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, true);
    const uint32_t a = smem_u32(smem), b = a + 16384;
    for (int i = 0; i < iters; ++i) {
      const uint64_t da = make_smem_desc(a, 16, 1024, kLayoutSW128);
      const uint64_t db = make_smem_desc(b, 16, 1024, kLayoutSW128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        for (int p = 0; p < passes; ++p) umma_f16(tm, da + 2 * k, db + 2 * k, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  const int rows_total = 1 << 20;  // 1M rows x 128 B = 128 MiB (> L2 when streamed by all SMs) ; small window => L2 hits
  void* buf; CK(cudaMalloc(&buf, (size_t)rows_total * 128)); CK(cudaMemset(buf, 0, (size_t)rows_total * 128));
  long long* cyc; CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
  std::vector<long long> h(1024);
  printf("== TMA ingest (box = rows x 128 B, SW128), bytes/clk/SM (min over CTAs = slowest)\n");
  for (int grid : {148, 16}) {
    for (int window_rows : {1 << 20, 1 << 14}) {   // 128 MiB window (DRAM) vs 2 MiB window (L2 hits)
      for (int box_rows : {32, 128, 256}) {
        for (int stages : {2, 4, 8, 12}) {
          const int box_bytes = box_rows * 128;
          if (stages * box_bytes > 200 * 1024) continue;
          CUtensorMap m;
          cuuint64_t dims[3] = {64, (cuuint64_t)window_rows, 1}; cuuint64_t strides[2] = {128, (cuuint64_t)window_rows * 128};
          cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1}; cuuint32_t es[3] = {1, 1, 1};
          if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
          const int iters = (8 << 20) / box_bytes;  // 8 MiB per CTA
          const int smem = stages * box_bytes + 1024 + 256;
          CK(cudaFuncSetAttribute(tma_ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          for (int rep = 0; rep < 2; ++rep) {
            tma_ingest_kernel<<<grid, 64, smem>>>(m, stages, box_bytes, iters, window_rows, box_rows, cyc);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
          long long mx = 0; double avg = 0; for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; } avg /= grid;
          printf("grid %3d window %4d MiB box %6d B stages %2d : %6.1f B/clk/SM (avg %6.1f)\n", grid, window_rows >> 13, box_bytes,
                 stages, (double)iters * box_bytes / mx, (double)iters * box_bytes / avg);
        }
      }
    }
  }
  printf("== tcgen05.mma rate, M=128 K=16, cycles per MMA (issue loop of 4 K-steps x passes)\n");
  const int smem = 64 * 1024;
  auto run = [&](auto kern, int n, int passes) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) { kern<<<148, 64, smem>>>(iters, passes, cyc); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N %3d passes %d : %.1f cycles/MMA\n", n, passes, (double)mx / (iters * 4.0 * passes));
  };
  for (int passes : {1, 3}) { run(mma_rate_kernel<16>, 16, passes); run(mma_rate_kernel<64>, 64, passes);
                              run(mma_rate_kernel<128>, 128, passes); run(mma_rate_kernel<256>, 256, passes); }
  return 0;
}
