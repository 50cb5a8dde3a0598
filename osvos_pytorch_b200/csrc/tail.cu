// Side-branch tail, forward: zero-padded bilinear "deconvolution" of the four
// low-resolution score / fuse-slice maps, centre crop, fusion sum and (optional)
// the per-pixel class-balanced BCE terms with their spatial reductions - one
// bandwidth-bound pass.
//
// Reference ops replaced (see include/osvos_b200.h): ConvTranspose2d with
// interp_surgery weights (networks/vgg_osvos.py:45-46,68-69;
// layers/osvos_layers.py:59-85), center_crop (layers/osvos_layers.py:51-56),
// cat + fuse (networks/vgg_osvos.py:71-72), loss terms
// (layers/osvos_layers.py:28-41).
//
// Math: the deconvolution with kernel 2s / stride s and taps
// f[t] = 1 - |t - (s - .5)| / s touches at most two source rows and columns per
// output pixel: with o = y + crop_top, a = o / s, r = o % s the rows are
// a (weight (r + .5)/s, if a < h_k) and a - 1 (weight 1 - (r + .5)/s, if a >= 1);
// rows outside the source contribute zero (zero padding => attenuated border).
#include "common.cuh"
#include "ptx.cuh"

namespace osvos {

struct TailScale {
  const float* pq;  // [n, hk, wk, 2]
  int hk, wk, s, log2s, top, left;
};
struct TailParams {
  TailScale sc[4];
  const float* fuse_bias;
  float* out[5];
  const float* label;
  double* sums;
  int n, h, w;
  int vec_mask;  // bit k: out[k] is 16-byte aligned; bit 5: label is
};

constexpr int kTailThreads = 256;

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x))); }

__global__ void __launch_bounds__(kTailThreads) tail_fwd_kernel(const TailParams p) {
  pdl_wait();               // side maps, biases and the accumulators all come from earlier kernels (ptx.cuh)
  pdl_launch_dependents();
  const size_t hw = static_cast<size_t>(p.h) * p.w;
  const size_t total = static_cast<size_t>(p.n) * hw;
  const size_t nvec = (total + 3) / 4;
  const float fb = p.fuse_bias ? __ldg(p.fuse_bias) : 0.f;

  float s_pos[5] = {0, 0, 0, 0, 0}, s_neg[5] = {0, 0, 0, 0, 0};
  float cnt_pos = 0.f;

  for (size_t v = blockIdx.x * static_cast<size_t>(kTailThreads) + threadIdx.x; v < nvec;
       v += static_cast<size_t>(gridDim.x) * kTailThreads) {
    const size_t e0 = v * 4;
    float o[5][4];
    float lab[4] = {0, 0, 0, 0};
    const bool full = (e0 + 3 < total);
    if (p.label) {
      if (full && (p.vec_mask & 32)) {
        const float4 l4 = __ldg(reinterpret_cast<const float4*>(p.label + e0));
        lab[0] = l4.x, lab[1] = l4.y, lab[2] = l4.z, lab[3] = l4.w;
      } else {
        for (int j = 0; j < 4; ++j)
          if (e0 + j < total) lab[j] = __ldg(p.label + e0 + j);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t e = e0 + j;
      float fused = fb;
      if (e < total) {
        const int img = static_cast<int>(e / hw);
        const int rem = static_cast<int>(e - static_cast<size_t>(img) * hw);
        const int y = rem / p.w, x = rem - y * p.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const TailScale& sc = p.sc[k];
          const int oy = y + sc.top, ox = x + sc.left;
          const int ay = oy >> sc.log2s, ax = ox >> sc.log2s;
          const float inv = 1.f / static_cast<float>(sc.s);
          const float fy1 = (static_cast<float>(oy & (sc.s - 1)) + 0.5f) * inv;  // weight of row ay
          const float fx1 = (static_cast<float>(ox & (sc.s - 1)) + 0.5f) * inv;  // weight of col ax
          const float wy[2] = {ay >= 1 ? 1.f - fy1 : 0.f, ay < sc.hk ? fy1 : 0.f};
          const float wx[2] = {ax >= 1 ? 1.f - fx1 : 0.f, ax < sc.wk ? fx1 : 0.f};
          const float2* base = reinterpret_cast<const float2*>(sc.pq) + static_cast<size_t>(img) * sc.hk * sc.wk;
          float sp = 0.f, sq = 0.f;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const int iy = ay - 1 + dy;
            if (wy[dy] == 0.f) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int ix = ax - 1 + dx;
              if (wx[dx] == 0.f) continue;
              const float2 t = __ldg(base + static_cast<size_t>(iy) * sc.wk + ix);
              const float wgt = wy[dy] * wx[dx];
              sp = fmaf(wgt, t.x, sp);
              sq = fmaf(wgt, t.y, sq);
            }
          }
          o[k][j] = sp;
          fused += sq;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][j] = 0.f;
      }
      o[4][j] = fused;
      if (p.label && e < total) {
        const bool pos = lab[j] >= 0.5f;
        cnt_pos += pos ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float xk = o[k][j];
          const float sp = softplus_f(xk);
          if (pos) s_pos[k] += sp - xk;
          else s_neg[k] += sp;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (!p.out[k]) continue;
      if (full && (p.vec_mask & (1 << k))) {
        *reinterpret_cast<float4*>(p.out[k] + e0) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
      } else {
        for (int j = 0; j < 4; ++j)
          if (e0 + j < total) p.out[k][e0 + j] = o[k][j];
      }
    }
  }

  if (p.label && p.sums) {
    __shared__ float red[kTailThreads / 32][11];
    float vals[11];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      vals[2 * k] = s_pos[k];
      vals[2 * k + 1] = s_neg[k];
    }
    vals[10] = cnt_pos;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 11; ++i) red[warp][i] = vals[i];
    }
    __syncthreads();
    if (threadIdx.x < 11) {
      double acc = 0.0;
      for (int wv = 0; wv < kTailThreads / 32; ++wv) acc += static_cast<double>(red[wv][threadIdx.x]);
      atomicAdd(p.sums + threadIdx.x, acc);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.sums[11] = static_cast<double>(total);
  }
}

}  // namespace osvos

using namespace osvos;

extern "C" int osvos_tail_fwd(const osvos_tail_fwd_args* a, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(a != nullptr && a->n > 0 && a->h > 0 && a->w > 0);
  OSVOS_CHECK_ARG(a->label == nullptr || a->sums != nullptr);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TailParams p;
  int hk = a->h, wk = a->w;
  for (int k = 0; k < 4; ++k) {
    OSVOS_CHECK_ARG(a->pq[k] != nullptr);
    hk = (hk + 1) / 2;
    wk = (wk + 1) / 2;
    const int s = 2 << k;
    p.sc[k].pq = a->pq[k];
    p.sc[k].hk = hk;
    p.sc[k].wk = wk;
    p.sc[k].s = s;
    p.sc[k].log2s = k + 1;
    p.sc[k].top = ((hk + 1) * s - a->h) / 2;   // layers/osvos_layers.py:52-56: floor(d/2) rows cropped on top
    p.sc[k].left = ((wk + 1) * s - a->w) / 2;
  }
  p.vec_mask = 0;
  for (int k = 0; k < 5; ++k) {
    p.out[k] = a->out[k];
    OSVOS_CHECK_ARG((reinterpret_cast<uintptr_t>(a->out[k]) & 3) == 0);
    if ((reinterpret_cast<uintptr_t>(a->out[k]) & 15) == 0) p.vec_mask |= 1 << k;
  }
  if ((reinterpret_cast<uintptr_t>(a->label) & 15) == 0) p.vec_mask |= 32;
  p.fuse_bias = a->fuse_bias;
  p.label = a->label;
  p.sums = a->sums;
  p.n = a->n;
  p.h = a->h;
  p.w = a->w;
  if (a->sums) OSVOS_CHECK_CUDA(cudaMemsetAsync(a->sums, 0, 12 * sizeof(double), stream));
  const size_t nvec = (static_cast<size_t>(a->n) * a->h * a->w + 3) / 4;
  size_t blocks = (nvec + kTailThreads - 1) / kTailThreads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  // (with a loss, the memset above is this kernel's stream predecessor: plain launch)
  if (a->sums) tail_fwd_kernel<<<static_cast<int>(blocks), kTailThreads, 0, stream>>>(p);
  else OSVOS_CHECK_CUDA(launch_pdl(tail_fwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kTailThreads), 0, stream, p));
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}
