// Backward of the side branch in FOLDED (rank-2) form - training path.
//
// side_prep[i] has no ReLU (reference networks/vgg_osvos.py:67), so the whole side branch of a scale,
//   feat = side_prep(x);  p = score_dsn(feat);  q = fuse_slice . feat            (:41,44,54 run at :67,69,72)
// is ONE linear 3x3 convolution C -> 2 with weights W'[o][c][t] = sum_f proj[o][f] * W_side[f][c][t] (the inference
// path already runs it that way, osvos_fold_side_weights).  Its backward therefore only ever sees the TWO gradient
// channels dpq = (dL/dp, dL/dq) - not the 16 feature gradients autograd materialises:
//
//   G[t][o][c] = sum_px dpq[px - t][o] * x[px][c]        "folded weight gradient": 18 numbers per channel     (1)
//   S[o]       = sum_px dpq[px][o]
//   dX[px][c]  = sum_{t,o} W'[o][c][t] * dpq[px - t][o]   gradient w.r.t. the stage output (before its ReLU)   (2)
//
// and every parameter gradient of the branch is algebra on G and S (side_grads_finish_kernel):
//   d side_prep.weight[f][c][t] = proj[0][f] G[t][0][c] + proj[1][f] G[t][1][c]
//   d side_prep.bias[f]         = proj[0][f] S[0]       + proj[1][f] S[1]
//   d score_dsn.weight[f]       = <W_side[f], G[.][0][.]> + b_side[f] S[0],     d score_dsn.bias = S[0]
//   d fuse.weight[16 i + f]     = <W_side[f], G[.][1][.]> + b_side[f] S[1]
// (1) reads the stage output ONCE on CUDA cores (18 FMAs per element, fp32 accumulate over hi + lo) instead of nine
// shifted passes of a 64-wide tensor-core wgrad whose N is 3/4 zero padding; (2) is 18 FMAs per element inside the
// max-unpool / ReLU-mask kernel that consumes it (bwd_kernels.cu), instead of a 3x3 dgrad convolution 16 -> C that
// wrote an fp32 map of the stage's size only to be read back once.  The 16 side features, their gradient and the padded
// 64-channel operand copies are never formed.  Replaces the autograd of networks/vgg_osvos.py:67,69,72 triggered at
// train_online.py:141 / train_parent.py:164.
#include "common.cuh"
#include "ptx.cuh"

namespace osvos {

constexpr int kSwThreads = 256;     // 8 warps; a warp owns a 128-channel slab (4 channels per lane) of a row segment
constexpr int kSwSlab = 128;

// G[t][o][c] (+ S[2] behind it), t = 3 r + s.  Work item = (image, row, segment of `seg` pixels); a block's eight warps
// share one 128-channel slab (blockIdx % slabs) and walk items together, so that ONE shared-memory reduction per block
// precedes the global atomics (18 x 128 floats per block).  Per item the 3 x (seg + 2) window of dpq goes to shared memory
// once (coalesced), and the lane's four channels of the next FOUR pixels are always in flight: the first version loaded a
// dpq column and one pixel per iteration and waited for both (54 us for the 52 MB stage-2 map, 1 TB/s).
constexpr int kSwMaxSeg = 32;
__global__ void __launch_bounds__(kSwThreads, 2)
side_folded_wgrad_kernel(const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                         const float* __restrict__ dpq, float* __restrict__ g, int n, int h, int w, int c, int seg) {
  __shared__ __align__(16) float red[18 * kSwSlab];
  __shared__ float red_s[2];
  __shared__ float2 win_all[kSwThreads / 32][3][kSwMaxSeg + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 (*win)[kSwMaxSeg + 2] = win_all[warp];
  const int slabs = c / kSwSlab;
  const int slab = static_cast<int>(blockIdx.x) % slabs;
  const int blk = static_cast<int>(blockIdx.x) / slabs, nblk = (static_cast<int>(gridDim.x) + slabs - 1 - slab) / slabs;
  const int c0 = slab * kSwSlab + lane * 4;
  for (int i = threadIdx.x; i < 18 * kSwSlab; i += kSwThreads) red[i] = 0.f;
  if (threadIdx.x < 2) red_s[threadIdx.x] = 0.f;
  __syncthreads();
  pdl_wait();               // dpq / x are outputs of earlier kernels of the stream (ptx.cuh)
  pdl_launch_dependents();

  float acc[9][2][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][o][j] = 0.f;
  float s0 = 0.f, s1 = 0.f;

  const int segs_x = (w + seg - 1) / seg;
  const int items = n * h * segs_x;
  const bool has_lo = x_lo != nullptr;
  for (int item = blk * 8 + warp; item < items; item += nblk * 8) {
    const int sx = item % segs_x, row = item / segs_x;
    const int y = row % h, img = row / h;
    const int x0 = sx * seg;
    const int x1 = min(x0 + seg, w);
    // window: win[r][k] = dpq[(y + 1 - r, x0 - 1 + k)], zero outside the image (= dpq[px - t] for t = (r - 1, s - 1) at
    // k = (x - x0) + 2 - s)
    __syncwarp();
    for (int idx = lane; idx < 3 * (seg + 2); idx += 32) {
      const int r = idx / (seg + 2), k = idx - r * (seg + 2);
      const int yy = y + 1 - r, xx = x0 - 1 + k;
      float2 v = make_float2(0.f, 0.f);
      if (yy >= 0 && yy < h && xx >= 0 && xx < w)
        v = __ldg(reinterpret_cast<const float2*>(dpq) + (static_cast<size_t>(img) * h + yy) * w + xx);
      win[r][k] = v;
    }
    const size_t rowbase = (static_cast<size_t>(img) * h + y) * w;
    uint2 rh[4], rl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      rh[u] = rl[u] = make_uint2(0, 0);
      if (x0 + u < x1) {
        rh[u] = __ldg(reinterpret_cast<const uint2*>(x_hi + (rowbase + x0 + u) * c + c0));
        if (has_lo) rl[u] = __ldg(reinterpret_cast<const uint2*>(x_lo + (rowbase + x0 + u) * c + c0));
      }
    }
    __syncwarp();
    for (int xb = x0; xb < x1; xb += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int xx = xb + u;
        if (xx < x1) {
          float v[4];
          v[0] = bf16_lo_to_float(rh[u].x) + bf16_lo_to_float(rl[u].x);
          v[1] = bf16_hi_to_float(rh[u].x) + bf16_hi_to_float(rl[u].x);
          v[2] = bf16_lo_to_float(rh[u].y) + bf16_lo_to_float(rl[u].y);
          v[3] = bf16_hi_to_float(rh[u].y) + bf16_hi_to_float(rl[u].y);
          if (xx + 4 < x1) {          // this slot's next tenant: four pixels ahead
            rh[u] = __ldg(reinterpret_cast<const uint2*>(x_hi + (rowbase + xx + 4) * c + c0));
            if (has_lo) rl[u] = __ldg(reinterpret_cast<const uint2*>(x_lo + (rowbase + xx + 4) * c + c0));
          }
          const int k0 = xx - x0 + 2;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const float2 d = win[r][k0 - s];
              if (r == 1 && s == 1) {
                s0 += d.x;
                s1 += d.y;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc[3 * r + s][0][j] = fmaf(d.x, v[j], acc[3 * r + s][0][j]);
                acc[3 * r + s][1][j] = fmaf(d.y, v[j], acc[3 * r + s][1][j]);
              }
            }
        }
      }
    }
  }
  // block reduction in shared memory, then one vector atomic per (tap, o, 4 channels)
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&red[(2 * t + o) * kSwSlab + lane * 4 + j], acc[t][o][j]);
  if (slab == 0 && lane == 0) {   // every lane of a warp saw the same dpq: one lane counts
    atomicAdd(&red_s[0], s0);
    atomicAdd(&red_s[1], s1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 18 * (kSwSlab / 4); i += kSwThreads) {
    const int to = i / (kSwSlab / 4), c4 = i % (kSwSlab / 4);
    const float4 val = *reinterpret_cast<const float4*>(&red[to * kSwSlab + c4 * 4]);
    atomicAdd(reinterpret_cast<float4*>(g + static_cast<size_t>(to) * c + slab * kSwSlab + c4 * 4), val);
  }
  if (slab == 0 && threadIdx.x < 2) atomicAdd(g + static_cast<size_t>(18) * c + threadIdx.x, red_s[threadIdx.x]);
}

// Parameter gradients of the side branch of up to four scales from G / S: one block per (scale, feature f).
struct SideGradScale {
  const float* g;        // [18][c] + S[2]
  const float* side_w;   // [16][c][9]
  const float* side_b;   // [16] or null
  const float* proj;     // [32]
  float* d_side_w;       // [16][c][9]
  float* d_side_b;       // [16]
  float* d_score_w;      // [16] or null
  float* d_score_b;      // [1] or null
  float* d_fuse_w;       // [16] or null
  int c;
  int accumulate;
};
struct SideGradTable {
  SideGradScale s[4];
  int count;
};

constexpr int kFinThreads = 1024;
__global__ void __launch_bounds__(kFinThreads) side_grads_finish_kernel(const __grid_constant__ SideGradTable t) {
  extern __shared__ float gs[];            // G of this block's scale: 18 rows of c floats at pitch c + 1 (bank spread), S[2]
  const SideGradScale& L = t.s[blockIdx.x / 16];
  const int f = blockIdx.x % 16;
  const int c = L.c, pitch = c + 1;
  pdl_wait();
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 18 * c; i += kFinThreads) gs[(i / c) * pitch + i % c] = __ldcg(L.g + i);
  if (threadIdx.x < 2) gs[18 * pitch + threadIdx.x] = __ldcg(L.g + 18 * c + threadIdx.x);
  __syncthreads();
  const float ps = __ldg(L.proj + f), pf = __ldg(L.proj + 16 + f);
  const float S0 = gs[18 * pitch], S1 = gs[18 * pitch + 1];
  float dot0 = 0.f, dot1 = 0.f;
  const float* __restrict__ wrow = L.side_w + static_cast<size_t>(f) * c * 9;
  float* __restrict__ drow = L.d_side_w + static_cast<size_t>(f) * c * 9;
  const bool accumulate = L.accumulate != 0;
#pragma unroll 2
  for (int i = threadIdx.x; i < c * 9; i += kFinThreads) {
    const int ci = i / 9, tap = i - ci * 9;
    const float g0 = gs[2 * tap * pitch + ci];
    const float g1 = gs[(2 * tap + 1) * pitch + ci];
    const float wv = __ldg(wrow + i);
    dot0 = fmaf(wv, g0, dot0);
    dot1 = fmaf(wv, g1, dot1);
    const float dv = fmaf(ps, g0, pf * g1);
    drow[i] = accumulate ? drow[i] + dv : dv;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, off);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, off);
  }
  __shared__ float r0[kFinThreads / 32], r1[kFinThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    r0[threadIdx.x >> 5] = dot0;
    r1[threadIdx.x >> 5] = dot1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < kFinThreads / 32; ++i) {
      a += r0[i];
      b += r1[i];
    }
    const float bs = L.side_b ? __ldg(L.side_b + f) : 0.f;
    auto put = [&](float* dst, float v) {
      if (dst) *dst = accumulate ? *dst + v : v;
    };
    put(L.d_side_b + f, fmaf(ps, S0, pf * S1));
    if (L.d_score_w) put(L.d_score_w + f, fmaf(bs, S0, a));
    if (L.d_fuse_w) put(L.d_fuse_w + f, fmaf(bs, S1, b));
    if (L.d_score_b && f == 0) put(L.d_score_b, S0);
  }
}

}  // namespace osvos

using namespace osvos;

extern "C" size_t osvos_side_folded_wgrad_floats(int c) { return static_cast<size_t>(18) * c + 2; }

extern "C" int osvos_side_folded_wgrad(const void* x_hi, const void* x_lo, const float* dpq, float* g, int n, int h,
                                       int w, int c, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(x_hi != nullptr && dpq != nullptr && g != nullptr && n > 0 && h > 0 && w > 0);
  OSVOS_CHECK_ARG(c >= kSwSlab && c % kSwSlab == 0);
  OSVOS_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  const int slabs = c / kSwSlab;
  const long rows = static_cast<long>(n) * h;
  const long blocks_cap = static_cast<long>(device_sm_count()) * 2 / slabs > 0 ? static_cast<long>(device_sm_count()) * 2 / slabs : 1;
  // segment length: the longest of 32 / 16 / 8 pixels that still leaves ~2.5 items per warp of the slab (balance of the
  // static item assignment against the per-item window load)
  int seg = kSwMaxSeg;
  while (seg > 8 && rows * ((w + seg - 1) / seg) * 2 < blocks_cap * 8 * 5) seg >>= 1;
  const long items = rows * ((w + seg - 1) / seg);
  OSVOS_CHECK_ARG(items < (1l << 30));
  long blocks_per_slab = (items + 7) / 8;
  if (blocks_per_slab > blocks_cap) blocks_per_slab = blocks_cap;
  const unsigned grid = static_cast<unsigned>(blocks_per_slab * slabs);
  OSVOS_CHECK_CUDA(launch_pdl(side_folded_wgrad_kernel, dim3(grid), dim3(kSwThreads), 0, static_cast<cudaStream_t>(stream_),
                              static_cast<const __nv_bfloat16*>(x_hi), static_cast<const __nv_bfloat16*>(x_lo), dpq, g, n,
                              h, w, c, seg));
  return OSVOS_OK;
}

extern "C" int osvos_side_grads_finish(const osvos_side_grads_item* items, int count, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(items != nullptr && count > 0 && count <= 4);
  SideGradTable t;
  t.count = count;
  for (int i = 0; i < count; ++i) {
    const osvos_side_grads_item& it = items[i];
    OSVOS_CHECK_ARG(it.g != nullptr && it.side_w != nullptr && it.proj_w != nullptr && it.d_side_w != nullptr &&
                    it.d_side_b != nullptr && it.c > 0);
    SideGradScale& L = t.s[i];
    L.g = it.g;
    L.side_w = it.side_w;
    L.side_b = it.side_b;
    L.proj = it.proj_w;
    L.d_side_w = it.d_side_w;
    L.d_side_b = it.d_side_b;
    L.d_score_w = it.d_score_w;
    L.d_score_b = it.d_score_b;
    L.d_fuse_w = it.d_fuse_w;
    L.c = it.c;
    L.accumulate = it.accumulate ? 1 : 0;
  }
  int cmax = 0;
  for (int i = 0; i < count; ++i) cmax = items[i].c > cmax ? items[i].c : cmax;
  OSVOS_CHECK_ARG(cmax <= 2048);
  const size_t smem = (static_cast<size_t>(18) * (cmax + 1) + 2) * sizeof(float);
  static uint64_t attr_done = 0;
  if (smem > 48 * 1024)
    OSVOS_CHECK_CUDA(ensure_dynamic_smem(side_grads_finish_kernel, (18 * 2049 + 2) * static_cast<int>(sizeof(float)), &attr_done));
  OSVOS_CHECK_CUDA(launch_pdl(side_grads_finish_kernel, dim3(16 * count), dim3(kFinThreads), smem, static_cast<cudaStream_t>(stream_), t));
  return OSVOS_OK;
}
