// The callers either side of the forward/backward path:
//   * test-time output (train_online.py:181-187): fused logits -> 8-bit probability / bytescaled PNG payload /
//     binary mask, on the device, so H*W bytes instead of 4*H*W cross PCIe;
//   * the optimizer step (train_online.py:79-88,147; train_parent.py:87-103,170): momentum SGD with per-tensor
//     lr / weight decay over all ~50 trainable tensors in one launch, fused with gradient zeroing and with the
//     re-emission of the tensor-core operand layouts of the 3x3 conv weights.
// Both are HBM-bound streaming kernels: 128-bit accesses where alignment allows, grids sized from the SM count.
#include "common.cuh"

namespace osvos {

// ------------------------------------------------------------------ logits -> u8
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int kOutThreads = 256;

// ws[2f] = max over frame f of ordered(x), ws[2f+1] = max of ~ordered(x) (i.e. the minimum); both start at 0.
__global__ void __launch_bounds__(kOutThreads)
frame_minmax_kernel(const float* __restrict__ x, size_t per_frame, uint32_t* __restrict__ ws) {
  const int f = blockIdx.y;
  const float* xf = x + static_cast<size_t>(f) * per_frame;
  uint32_t kmax = 0u, kmin = 0xFFFFFFFFu;
  for (size_t i = blockIdx.x * static_cast<size_t>(kOutThreads) + threadIdx.x; i < per_frame;
       i += static_cast<size_t>(gridDim.x) * kOutThreads) {
    const uint32_t k = float_to_ordered(__ldg(xf + i));
    kmax = max(kmax, k);
    kmin = min(kmin, k);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, off));
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(ws + 2 * f, kmax);
    atomicMax(ws + 2 * f + 1, ~kmin);
  }
}

__global__ void __launch_bounds__(kOutThreads)
logits_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, const uint32_t* __restrict__ ws,
                    size_t per_frame, int mode, int vec_ok) {
  const int f = blockIdx.y;
  const float* xf = x + static_cast<size_t>(f) * per_frame;
  uint8_t* of = out + static_cast<size_t>(f) * per_frame;
  float pmin = 0.f, scale = 255.f;
  if (mode == OSVOS_U8_BYTESCALE) {
    // sigmoid is monotone: the extrema of p are the sigmoids of the extrema of x
    pmin = sigmoid_f(ordered_to_float(~ws[2 * f + 1]));
    const float pmax = sigmoid_f(ordered_to_float(ws[2 * f]));
    float cscale = pmax - pmin;
    if (cscale == 0.f) cscale = 1.f;
    scale = 255.f / cscale;
  }
  auto conv = [&](float v) -> uint32_t {
    if (mode == OSVOS_U8_MASK) return v > 0.f ? 255u : 0u;
    const float b = (sigmoid_f(v) - pmin) * scale;
    return static_cast<uint32_t>(fminf(fmaxf(b, 0.f), 255.f) + 0.5f);
  };
  const size_t nvec = vec_ok ? per_frame / 4 : 0;
  for (size_t v = blockIdx.x * static_cast<size_t>(kOutThreads) + threadIdx.x; v < nvec;
       v += static_cast<size_t>(gridDim.x) * kOutThreads) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(xf) + v);
    const uint32_t packed = conv(xv.x) | (conv(xv.y) << 8) | (conv(xv.z) << 16) | (conv(xv.w) << 24);
    reinterpret_cast<uint32_t*>(of)[v] = packed;
  }
  for (size_t e = nvec * 4 + blockIdx.x * static_cast<size_t>(kOutThreads) + threadIdx.x; e < per_frame;
       e += static_cast<size_t>(gridDim.x) * kOutThreads)
    of[e] = static_cast<uint8_t>(conv(xf[e]));
}

// ------------------------------------------------------------------ fused SGD step
// Work decomposition: a plain tensor is cut into items of kSgdChunk elements; a packed 3x3 conv weight into tiles
// of kTileCo output channels x kTileCi input channels x 9 taps (kSgdChunk elements again), so that both packed
// layouts receive whole 32-byte sectors: forward rows are [tap][co][ci..ci+31] (64 B), flipped rows
// [8-tap][ci][co..co+15] (32 B).
constexpr int kSgdThreads = 256;
constexpr int kTileCo = 16, kTileCi = 32;
constexpr int kSgdChunk = kTileCo * kTileCi * 9;  // 4608

__device__ __forceinline__ void sgd_update(float& p, float g, float& m, float lr, float wd, float mu) {
  const float gp = fmaf(wd, p, g);
  m = __fadd_rn(__fmul_rn(mu, m), gp);
  p = fmaf(-lr, m, p);
}

__global__ void __launch_bounds__(kSgdThreads)
sgd_step_kernel(const osvos_sgd_segment* __restrict__ segs, int count, int zero_grad) {
  __shared__ float tile[kTileCo][kTileCi * 9 + 1];
  __shared__ int s_seg;
  __shared__ uint32_t s_item;
  if (threadIdx.x == 0) {
    uint32_t item = blockIdx.x;
    int sidx = 0;
    while (sidx < count && item >= segs[sidx].work_items) {
      item -= segs[sidx].work_items;
      ++sidx;
    }
    s_seg = sidx;
    s_item = item;
  }
  __syncthreads();
  if (s_seg >= count) return;
  const osvos_sgd_segment sg = segs[s_seg];
  const uint32_t item = s_item;
  const float lr = sg.lr, wd = sg.weight_decay, mu = sg.momentum_coef;
  if (sg.packed_fwd == nullptr && sg.packed_flip == nullptr) {
    const uint64_t begin = static_cast<uint64_t>(item) * kSgdChunk;
    const uint64_t end = min(begin + static_cast<uint64_t>(kSgdChunk), sg.numel);
    const bool vec = ((reinterpret_cast<uintptr_t>(sg.param) | reinterpret_cast<uintptr_t>(sg.grad) |
                       reinterpret_cast<uintptr_t>(sg.momentum)) & 15) == 0;
    uint64_t e = begin;
    if (vec) {  // begin is a multiple of 4 elements
      const uint64_t nvec = (end - begin) / 4;
      for (uint64_t v = threadIdx.x; v < nvec; v += kSgdThreads) {
        float4 p = reinterpret_cast<float4*>(sg.param + begin)[v];
        const float4 g = reinterpret_cast<const float4*>(sg.grad + begin)[v];
        float4 m = reinterpret_cast<float4*>(sg.momentum + begin)[v];
        sgd_update(p.x, g.x, m.x, lr, wd, mu);
        sgd_update(p.y, g.y, m.y, lr, wd, mu);
        sgd_update(p.z, g.z, m.z, lr, wd, mu);
        sgd_update(p.w, g.w, m.w, lr, wd, mu);
        reinterpret_cast<float4*>(sg.param + begin)[v] = p;
        reinterpret_cast<float4*>(sg.momentum + begin)[v] = m;
        if (zero_grad) reinterpret_cast<float4*>(sg.grad + begin)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      e = begin + nvec * 4;
    }
    for (uint64_t i = e + threadIdx.x; i < end; i += kSgdThreads) {
      float p = sg.param[i], m = sg.momentum[i];
      sgd_update(p, sg.grad[i], m, lr, wd, mu);
      sg.param[i] = p;
      sg.momentum[i] = m;
      if (zero_grad) sg.grad[i] = 0.f;
    }
    return;
  }
  // ---- 3x3 conv weight tile: co0..co0+15, ci0..ci0+31, all 9 taps (OIHW rows of 288 contiguous floats)
  const int ci_tiles = sg.cin / kTileCi;
  const int co0 = static_cast<int>(item / ci_tiles) * kTileCo;
  const int ci0 = static_cast<int>(item % ci_tiles) * kTileCi;
  constexpr int kRow = kTileCi * 9;
  for (int i = threadIdx.x; i < kTileCo * kRow; i += kSgdThreads) {
    const int co = i / kRow, r = i % kRow;
    const size_t e = (static_cast<size_t>(co0 + co) * sg.cin + ci0) * 9 + r;
    float p = sg.param[e], m = sg.momentum[e];
    sgd_update(p, sg.grad[e], m, lr, wd, mu);
    sg.param[e] = p;
    sg.momentum[e] = m;
    if (zero_grad) sg.grad[e] = 0.f;
    tile[co][r] = p;
  }
  __syncthreads();
  if (sg.packed_fwd != nullptr) {  // [plane][tap][co][colp_fwd]
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(sg.packed_fwd);
    const size_t plane = static_cast<size_t>(9) * sg.cout * sg.colp_fwd;
    for (int i = threadIdx.x; i < 9 * kTileCo * (kTileCi / 2); i += kSgdThreads) {
      const int cp = i % (kTileCi / 2);
      const int co = (i / (kTileCi / 2)) % kTileCo;
      const int t = i / (kTileCi / 2 * kTileCo);
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(tile[co][(2 * cp) * 9 + t], h0, l0);
      split_bf16(tile[co][(2 * cp + 1) * 9 + t], h1, l1);
      const size_t o = (static_cast<size_t>(t) * sg.cout + co0 + co) * sg.colp_fwd + ci0 + 2 * cp;
      *reinterpret_cast<uint32_t*>(out + o) = pack_bf16x2(h0, h1);
      *reinterpret_cast<uint32_t*>(out + plane + o) = pack_bf16x2(l0, l1);
    }
  }
  if (sg.packed_flip != nullptr) {  // [plane][8-tap][ci][colp_flip], columns = co
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(sg.packed_flip);
    const size_t plane = static_cast<size_t>(9) * sg.cin * sg.colp_flip;
    for (int i = threadIdx.x; i < 9 * kTileCi * (kTileCo / 2); i += kSgdThreads) {
      const int cp = i % (kTileCo / 2);
      const int ci = (i / (kTileCo / 2)) % kTileCi;
      const int t = i / (kTileCo / 2 * kTileCi);
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(tile[2 * cp][ci * 9 + t], h0, l0);
      split_bf16(tile[2 * cp + 1][ci * 9 + t], h1, l1);
      const size_t o = (static_cast<size_t>(8 - t) * sg.cin + ci0 + ci) * sg.colp_flip + co0 + 2 * cp;
      *reinterpret_cast<uint32_t*>(out + o) = pack_bf16x2(h0, h1);
      *reinterpret_cast<uint32_t*>(out + plane + o) = pack_bf16x2(l0, l1);
    }
  }
}

}  // namespace osvos

using namespace osvos;

extern "C" int osvos_logits_to_u8(const float* logits, uint8_t* out, uint32_t* minmax_ws, int frames, size_t per_frame,
                                  int mode, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(logits != nullptr && out != nullptr && frames > 0 && per_frame > 0);
  OSVOS_CHECK_ARG(mode == OSVOS_U8_PROB || mode == OSVOS_U8_BYTESCALE || mode == OSVOS_U8_MASK);
  OSVOS_CHECK_ARG(mode != OSVOS_U8_BYTESCALE || minmax_ws != nullptr);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  size_t bx = (per_frame / 4 + kOutThreads - 1) / kOutThreads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 8 / static_cast<size_t>(frames) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(frames));
  if (mode == OSVOS_U8_BYTESCALE) {
    OSVOS_CHECK_CUDA(cudaMemsetAsync(minmax_ws, 0, sizeof(uint32_t) * 2 * frames, stream));
    frame_minmax_kernel<<<grid, kOutThreads, 0, stream>>>(logits, per_frame, minmax_ws);
  }
  const int vec_ok = ((reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 &&
                      per_frame % 4 == 0) ? 1 : 0;
  logits_to_u8_kernel<<<grid, kOutThreads, 0, stream>>>(logits, out, minmax_ws, per_frame, mode, vec_ok);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

extern "C" uint32_t osvos_sgd_work_items(uint64_t numel, int cout, int cin) {
  if (cout > 0 && cin > 0) {
    if (cout % kTileCo != 0 || cin % kTileCi != 0 || numel != static_cast<uint64_t>(cout) * cin * 9) return 0;
    return static_cast<uint32_t>((cout / kTileCo) * (cin / kTileCi));
  }
  return static_cast<uint32_t>((numel + kSgdChunk - 1) / kSgdChunk);
}

extern "C" int osvos_sgd_step(const osvos_sgd_segment* segments, int count, uint32_t total_work_items, int zero_grad,
                              osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(segments != nullptr && count > 0 && count <= OSVOS_SGD_MAX_SEGMENTS && total_work_items > 0);
  sgd_step_kernel<<<total_work_items, kSgdThreads, 0, static_cast<cudaStream_t>(stream_)>>>(segments, count, zero_grad);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}
