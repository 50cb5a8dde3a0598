// Bandwidth-bound backward kernels around the tensor-core dgrad / wgrad GEMMs:
// adjoint of the bilinear tail, side-branch 1x1 backward, max-unpool + ReLU
// mask, per-channel bias-gradient sums and the conv1_1 (Cin = 3) backward.
// They replace the autograd graph PyTorch builds for reference
// networks/vgg_osvos.py:59-74 (triggered at train_online.py:141, train_parent.py:164).
#include "common.cuh"

namespace osvos {

// ------------------------------------------------------------------ tail bwd
// dpq[k][img, iy, ix] = { sum f f g_k , sum f f g_4 } over the (2s)^2 footprint of
// the low-res pixel in the cropped full-resolution maps (adjoint of tail_fwd).
struct TailBwdParams {
  const float* gk;   // [n,1,h,w] gradient of side output k (may be NULL)
  const float* g4;   // gradient of the fused output (may be NULL)
  float* dpq;        // [n,hk,wk,2]
  int n, h, w, hk, wk, s, top, left;
};

// LANES threads cooperate on one low-res pixel (2 / 8 / 32 / 32 for s = 2 / 4 / 8 / 16): 8 .. 32 taps per lane,
// sub-warp shuffle reduction.
template <int LANES>
__global__ void __launch_bounds__(256) tail_bwd_kernel(const TailBwdParams p) {
  const int sub = threadIdx.x % LANES;
  const size_t grp_global = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) / LANES;
  const size_t ngroups = (static_cast<size_t>(gridDim.x) * blockDim.x) / LANES;
  const size_t total = static_cast<size_t>(p.n) * p.hk * p.wk;
  const int fs = 2 * p.s;
  const float inv = 1.f / static_cast<float>(p.s);
  const size_t rounds = (total + ngroups - 1) / ngroups;   // uniform trip count: every lane takes part in the shuffles
  for (size_t rd = 0; rd < rounds; ++rd) {
    const size_t i = grp_global + rd * ngroups;
    const bool live = i < total;
    float dp = 0.f, dq = 0.f;
    if (live) {
      const int ix = static_cast<int>(i % p.wk);
      const int iy = static_cast<int>((i / p.wk) % p.hk);
      const int img = static_cast<int>(i / (static_cast<size_t>(p.wk) * p.hk));
      const size_t base = static_cast<size_t>(img) * p.h * p.w;
      for (int t = sub; t < fs * fs; t += LANES) {
        const int ty = t / fs, tx = t - ty * fs;
        const int y = iy * p.s + ty - p.top, x = ix * p.s + tx - p.left;
        if (y < 0 || y >= p.h || x < 0 || x >= p.w) continue;
        const float fy = 1.f - fabsf(static_cast<float>(ty) - (static_cast<float>(p.s) - 0.5f)) * inv;
        const float fx = 1.f - fabsf(static_cast<float>(tx) - (static_cast<float>(p.s) - 0.5f)) * inv;
        const float wgt = fy * fx;
        const size_t o = base + static_cast<size_t>(y) * p.w + x;
        if (p.gk) dp = fmaf(wgt, __ldg(p.gk + o), dp);
        if (p.g4) dq = fmaf(wgt, __ldg(p.g4 + o), dq);
      }
    }
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
      dp += __shfl_xor_sync(0xffffffffu, dp, off);
      dq += __shfl_xor_sync(0xffffffffu, dq, off);
    }
    if (live && sub == 0) *reinterpret_cast<float2*>(p.dpq + i * 2) = make_float2(dp, dq);
  }
}

// ---------------------------------------------------------------- generic sum
__global__ void __launch_bounds__(256) sum_f32_kernel(const float* __restrict__ x, size_t n, double* __restrict__ out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    acc += __ldg(x + i);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += static_cast<double>(red[i]);
    atomicAdd(out, t);
  }
}
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}
// side_bwd finalize: out[0:34] = sums; out[34+c] = d side_prep.bias[c] = w_score[c]*sum(dp) + w_fuse[c]*sum(dq)
__global__ void side_finalize_kernel(const double* __restrict__ in, const float* __restrict__ pw, float* __restrict__ out) {
  const int i = threadIdx.x;
  if (i < 34) out[i] = static_cast<float>(in[i]);
  else if (i < 50) out[i] = static_cast<float>(static_cast<double>(pw[i - 34]) * in[16] + static_cast<double>(pw[i - 18]) * in[33]);
}

// ------------------------------------------------------------------ side bwd
// feat [npix][16] fp32, dpq [npix][2], pw[32] = {score_dsn w, fuse slice}:
//   dfeat[px][c] = dp*pw[c] + dq*pw[16+c]  -> act with 64 channels (16..63 zero)
//   acc[0:16] += dp*feat, acc[16] += dp, acc[17:33] += dq*feat, acc[33] += dq   (fp64 atomics)
__global__ void __launch_bounds__(256)
side_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ dpq, const float* __restrict__ pw,
                __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, double* __restrict__ acc,
                size_t npix) {
  float a[34];
#pragma unroll
  for (int j = 0; j < 34; ++j) a[j] = 0.f;
  float w[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) w[j] = __ldg(pw + j);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float2 g = __ldg(reinterpret_cast<const float2*>(dpq) + i);
    float f[16];
    if (feat) {
      const float4* fp = reinterpret_cast<const float4*>(feat + i * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = __ldg(fp + j);
        f[4 * j] = v.x, f[4 * j + 1] = v.y, f[4 * j + 2] = v.z, f[4 * j + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        a[j] = fmaf(g.x, f[j], a[j]);
        a[17 + j] = fmaf(g.y, f[j], a[17 + j]);
      }
    }
    a[16] += g.x;
    a[33] += g.y;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v0 = fmaf(g.x, w[2 * j], g.y * w[16 + 2 * j]);
      const float v1 = fmaf(g.x, w[2 * j + 1], g.y * w[16 + 2 * j + 1]);
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v0, h0, l0);
      split_bf16(v1, h1, l1);
      hi[j] = pack_bf16x2(h0, h1);
      lo[j] = pack_bf16x2(l0, l1);
    }
    uint4* dh = reinterpret_cast<uint4*>(d_hi + i * 64);
    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 2; j < 8; ++j) dh[j] = z;
    if (d_lo) {
      uint4* dl = reinterpret_cast<uint4*>(d_lo + i * 64);
      dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
#pragma unroll
      for (int j = 2; j < 8; ++j) dl[j] = z;
    }
  }
  __shared__ float red[8][34];
#pragma unroll
  for (int j = 0; j < 34; ++j) {
    float v = a[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 34) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += static_cast<double>(red[i][threadIdx.x]);
    atomicAdd(acc + threadIdx.x, t);
  }
}

// ------------------------------------------------- max-unpool + add + ReLU mask
// dz[n,h,w,c] = (x > 0) * (dside + (pixel is the argmax of its 2x2 window ? dpool : 0))
// One thread per (pooled pixel, 8 channels).  Pass 1 reads the four activations and keeps only the argmax index and
// the sign bits (a few registers, so that many threads / loads are in flight); pass 2 streams dside / dpool and writes.
__device__ __forceinline__ void load_pair8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&v)[8]) {
  const uint4 vh = __ldg(reinterpret_cast<const uint4*>(hi + off));
  uint4 vl = make_uint4(0, 0, 0, 0);
  if (lo) vl = __ldg(reinterpret_cast<const uint4*>(lo + off));
  const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w};
  const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    v[2 * t] = bf16_lo_to_float(hw[t]) + bf16_lo_to_float(lw[t]);
    v[2 * t + 1] = bf16_hi_to_float(hw[t]) + bf16_hi_to_float(lw[t]);
  }
}

__global__ void __launch_bounds__(256)
unpool_add_mask_kernel(const __nv_bfloat16* __restrict__ dp_hi, const __nv_bfloat16* __restrict__ dp_lo,
                       const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                       const float* __restrict__ dside, __nv_bfloat16* __restrict__ dz_hi,
                       __nv_bfloat16* __restrict__ dz_lo, float* __restrict__ colsum, int n, int h, int w, int c,
                       int oh, int ow) {
  extern __shared__ float cs[];  // [c] block-local channel sums (fused bias gradient)
  if (colsum) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) cs[i] = 0.f;
    __syncthreads();
  }
  const int groups = c / 8;
  const size_t total = static_cast<size_t>(n) * oh * ow * groups;
  // blockDim (256) is a multiple of `groups`, so a thread keeps the same channel group over the whole loop
  const int g = static_cast<int>(threadIdx.x % groups);
  float csum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i / groups;
    const int ox = static_cast<int>(r % ow);
    r /= ow;
    const int oy = static_cast<int>(r % oh);
    const int nn = static_cast<int>(r / oh);
    // pass 1: argmax (first maximum in (dy, dx) scan order) and positivity of the four window elements
    float best[8];
    uint32_t arg = 0, pos = 0;  // arg: 2 bits per channel; pos: bit (q * 8 + j) = x[q][j] > 0
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int iy = 2 * oy + (q >> 1), ix = 2 * ox + (q & 1);
      if (iy >= h || ix >= w) continue;
      float v[8];
      load_pair8(x_hi, x_lo, ((static_cast<size_t>(nn) * h + iy) * w + ix) * c + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (q == 0 || v[j] > best[j]) {
          best[j] = v[j];
          arg = (arg & ~(3u << (2 * j))) | (static_cast<uint32_t>(q) << (2 * j));
        }
        if (v[j] > 0.f) pos |= 1u << (q * 8 + j);
      }
    }
    float dpv[8];
    load_pair8(dp_hi, dp_lo, ((static_cast<size_t>(nn) * oh + oy) * ow + ox) * c + g * 8, dpv);
    // pass 2: gradients of the four positions
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int iy = 2 * oy + (q >> 1), ix = 2 * ox + (q & 1);
      if (iy >= h || ix >= w) continue;
      const size_t dst = ((static_cast<size_t>(nn) * h + iy) * w + ix) * c + g * 8;
      float ds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (dside) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(dside + dst));
        const float4 b = __ldg(reinterpret_cast<const float4*>(dside + dst) + 1);
        ds[0] = a.x, ds[1] = a.y, ds[2] = a.z, ds[3] = a.w, ds[4] = b.x, ds[5] = b.y, ds[6] = b.z, ds[7] = b.w;
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float v0 = ds[2 * t] + (((arg >> (4 * t)) & 3u) == static_cast<uint32_t>(q) ? dpv[2 * t] : 0.f);
        float v1 = ds[2 * t + 1] + (((arg >> (4 * t + 2)) & 3u) == static_cast<uint32_t>(q) ? dpv[2 * t + 1] : 0.f);
        if (!((pos >> (q * 8 + 2 * t)) & 1u)) v0 = 0.f;
        if (!((pos >> (q * 8 + 2 * t + 1)) & 1u)) v1 = 0.f;
        csum[2 * t] += v0;
        csum[2 * t + 1] += v1;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v0, h0, l0);
        split_bf16(v1, h1, l1);
        hi[t] = pack_bf16x2(h0, h1);
        lo[t] = pack_bf16x2(l0, l1);
      }
      *reinterpret_cast<uint4*>(dz_hi + dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (dz_lo) *reinterpret_cast<uint4*>(dz_lo + dst) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  if (colsum) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&cs[g * 8 + j], csum[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(colsum + i, cs[i]);
  }
}

// ------------------------------------------------- per-channel sums of an act
// out[c] += sum_px (hi + lo)[px][c]   (bias gradient; out zeroed by the caller)
__global__ void __launch_bounds__(256)
channel_sum_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, size_t npix, int c,
                   float* __restrict__ out) {
  const int groups = c / 8;               // threads across channels (8 channels each)
  const int rows = 256 / groups;          // pixel rows handled concurrently by the block
  const int g = threadIdx.x % groups, ry = threadIdx.x / groups;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ry < rows) {
    for (size_t px = blockIdx.x * static_cast<size_t>(rows) + ry; px < npix; px += static_cast<size_t>(gridDim.x) * rows) {
      const uint4 vh = __ldg(reinterpret_cast<const uint4*>(hi + px * c + g * 8));
      uint4 vl = make_uint4(0, 0, 0, 0);
      if (lo) vl = __ldg(reinterpret_cast<const uint4*>(lo + px * c + g * 8));
      const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w};
      const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc[2 * t] += bf16_lo_to_float(hw[t]) + bf16_lo_to_float(lw[t]);
        acc[2 * t + 1] += bf16_hi_to_float(hw[t]) + bf16_hi_to_float(lw[t]);
      }
    }
  }
  extern __shared__ float sm[];  // [256][8]
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    const int gg = ch / 8, j = ch % 8;
    float t = 0.f;
    for (int r = 0; r < rows; ++r) t += sm[(r * groups + gg) * 8 + j];
    atomicAdd(out + ch, t);
  }
}

// -------------------------------------------------------------- conv1_1 bwd
// dW[co][ci][r][s] = sum_px dz[px][co] * x[ci][px + (r-1, s-1)]   (64 x 27 outputs, reduction over all pixels)
// Register-tiled: a thread owns a 4 (co) x 7 (k) tile of dW; 64 threads cover the 64 x 28 tile, the block's four
// 64-thread units take every fourth pixel of a staged 64-pixel chunk.  Per pixel a thread does one float4 + 7 scalar
// shared-memory loads for 28 FMAs.
constexpr int kFwPix = 64;
__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dz_hi,
                        const __nv_bfloat16* __restrict__ dz_lo, float* __restrict__ dw, int n, int h, int w) {
  __shared__ __align__(16) float dzs[kFwPix][68];
  __shared__ float xs[kFwPix][28];
  __shared__ float red[64 * 28];
  const int unit = threadIdx.x >> 6;         // 0..3: which pixels of the chunk
  const int t = threadIdx.x & 63;
  const int cg = t & 15, kg = t >> 4;        // co = 4*cg .. 4*cg+3 ; k = 7*kg .. 7*kg+6
  float acc[4][7];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
  const size_t npix = static_cast<size_t>(n) * h * w;
  for (size_t base = static_cast<size_t>(blockIdx.x) * kFwPix; base < npix; base += static_cast<size_t>(gridDim.x) * kFwPix) {
    // stage dz: 64 px x 64 co, 8 channels (16 B per plane) per thread-iteration
    for (int i = threadIdx.x; i < kFwPix * 8; i += 256) {
      const int pp = i >> 3, g = i & 7;
      const size_t px = base + pp;
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (px < npix) {
        const uint4 vh = __ldg(reinterpret_cast<const uint4*>(dz_hi + px * 64 + g * 8));
        uint4 vl = make_uint4(0, 0, 0, 0);
        if (dz_lo) vl = __ldg(reinterpret_cast<const uint4*>(dz_lo + px * 64 + g * 8));
        const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w};
        const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[2 * q] = bf16_lo_to_float(hw[q]) + bf16_lo_to_float(lw[q]);
          v[2 * q + 1] = bf16_hi_to_float(hw[q]) + bf16_hi_to_float(lw[q]);
        }
      }
      *reinterpret_cast<float4*>(&dzs[pp][g * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&dzs[pp][g * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int i = threadIdx.x; i < kFwPix * 28; i += 256) {
      const int pp = i / 28, k = i - pp * 28;
      const size_t px = base + pp;
      float v = 0.f;
      if (px < npix && k < 27) {
        const int xx = static_cast<int>(px % w);
        const int yy = static_cast<int>((px / w) % h);
        const int nn = static_cast<int>(px / (static_cast<size_t>(w) * h));
        const int ci = k / 9, r = (k % 9) / 3, sft = k % 3;
        const int iy = yy + r - 1, ix = xx + sft - 1;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = __ldg(x + ((static_cast<size_t>(nn) * 3 + ci) * h + iy) * w + ix);
      }
      xs[pp][k] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = unit; pp < kFwPix; pp += 4) {
      const float4 d = *reinterpret_cast<const float4*>(&dzs[pp][cg * 4]);
      float xv[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) xv[j] = xs[pp][kg * 7 + j];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        acc[0][j] = fmaf(d.x, xv[j], acc[0][j]);
        acc[1][j] = fmaf(d.y, xv[j], acc[1][j]);
        acc[2][j] = fmaf(d.z, xv[j], acc[2][j]);
        acc[3][j] = fmaf(d.w, xv[j], acc[3][j]);
      }
    }
    __syncthreads();
  }
  // reduce the four units in shared memory, then one global atomic per output and block
  for (int i = threadIdx.x; i < 64 * 28; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) atomicAdd(&red[(cg * 4 + i) * 28 + kg * 7 + j], acc[i][j]);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 27; i += 256) {
    const int co = i / 27, k = i - co * 27;
    atomicAdd(dw + i, red[co * 28 + k]);
  }
}

// dx[ci][y][x] = sum_{r,s,co} dz[y - (r-1)][x - (s-1)][co] * w[co][ci][r][s]
__global__ void __launch_bounds__(128)
conv_first_dgrad_kernel(const __nv_bfloat16* __restrict__ dz_hi, const __nv_bfloat16* __restrict__ dz_lo,
                        const float* __restrict__ wgt, float* __restrict__ dx, int n, int h, int w) {
  __shared__ float ws[27 * 64];  // [k = ci*9 + r*3 + s][co]
  for (int i = threadIdx.x; i < 27 * 64; i += 128) {
    const int co = i & 63, k = i >> 6;
    ws[i] = wgt[co * 27 + k];
  }
  __syncthreads();
  const int xx = blockIdx.x * 128 + threadIdx.x, yy = blockIdx.y, nn = blockIdx.z;
  if (xx >= w) return;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int r = 0; r < 3; ++r) {
    const int iy = yy - (r - 1);
    if (iy < 0 || iy >= h) continue;
    for (int s = 0; s < 3; ++s) {
      const int ix = xx - (s - 1);
      if (ix < 0 || ix >= w) continue;
      const size_t src = ((static_cast<size_t>(nn) * h + iy) * w + ix) * 64;
      const uint4* ph = reinterpret_cast<const uint4*>(dz_hi + src);
      const uint4* pl = dz_lo ? reinterpret_cast<const uint4*>(dz_lo + src) : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 vh = __ldg(ph + j);
        uint4 vl = make_uint4(0, 0, 0, 0);
        if (pl) vl = __ldg(pl + j);
        const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w};
        const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float d0 = bf16_lo_to_float(hw[t]) + bf16_lo_to_float(lw[t]);
          const float d1 = bf16_hi_to_float(hw[t]) + bf16_hi_to_float(lw[t]);
          const int co = 8 * j + 2 * t;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            const int k = ci * 9 + r * 3 + s;
            acc[ci] = fmaf(d0, ws[k * 64 + co], acc[ci]);
            acc[ci] = fmaf(d1, ws[k * 64 + co + 1], acc[ci]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) dx[((static_cast<size_t>(nn) * 3 + ci) * h + yy) * w + xx] = acc[ci];
}

static inline int grid_cap(size_t blocks, int per_sm) {
  const size_t cap = static_cast<size_t>(device_sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace osvos

using namespace osvos;

extern "C" int osvos_tail_bwd(const osvos_tail_bwd_args* a, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(a != nullptr && a->n > 0 && a->h > 0 && a->w > 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int hk = a->h, wk = a->w;
  for (int k = 0; k < 4; ++k) {
    hk = (hk + 1) / 2;
    wk = (wk + 1) / 2;
    OSVOS_CHECK_ARG(a->dpq[k] != nullptr);
    TailBwdParams p;
    p.gk = a->grad_out[k];
    p.g4 = a->grad_out[4];
    p.dpq = a->dpq[k];
    p.n = a->n;
    p.h = a->h;
    p.w = a->w;
    p.hk = hk;
    p.wk = wk;
    p.s = 2 << k;
    p.top = ((hk + 1) * p.s - a->h) / 2;
    p.left = ((wk + 1) * p.s - a->w) / 2;
    const size_t pixels = static_cast<size_t>(a->n) * hk * wk;
    if (k == 0) tail_bwd_kernel<2><<<grid_cap((pixels * 2 + 255) / 256, 16), 256, 0, stream>>>(p);
    else if (k == 1) tail_bwd_kernel<8><<<grid_cap((pixels * 8 + 255) / 256, 16), 256, 0, stream>>>(p);
    else tail_bwd_kernel<32><<<grid_cap((pixels * 32 + 255) / 256, 16), 256, 0, stream>>>(p);
  }
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

extern "C" int osvos_sum_f32(const float* x, size_t n, double* scratch, float* out, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(x != nullptr && scratch != nullptr && out != nullptr && n > 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OSVOS_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), stream));
  sum_f32_kernel<<<grid_cap((n + 255) / 256, 4), 256, 0, stream>>>(x, n, scratch);
  f64_to_f32_kernel<<<1, 32, 0, stream>>>(scratch, out, 1);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

extern "C" int osvos_side_bwd(const float* feat, const float* dpq, const float* proj_w, void* dfeat_hi, void* dfeat_lo,
                              double* scratch, float* param_grads, int n, int h, int w, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(dpq != nullptr && proj_w != nullptr && dfeat_hi != nullptr && scratch != nullptr &&
                  param_grads != nullptr && n > 0 && h > 0 && w > 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t npix = static_cast<size_t>(n) * h * w;
  OSVOS_CHECK_CUDA(cudaMemsetAsync(scratch, 0, 34 * sizeof(double), stream));
  side_bwd_kernel<<<grid_cap((npix + 255) / 256, 4), 256, 0, stream>>>(
      feat, dpq, proj_w, static_cast<__nv_bfloat16*>(dfeat_hi), static_cast<__nv_bfloat16*>(dfeat_lo), scratch, npix);
  side_finalize_kernel<<<1, 64, 0, stream>>>(scratch, proj_w, param_grads);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

extern "C" int osvos_unpool_add_mask(const void* dpool_hi, const void* dpool_lo, const void* x_hi, const void* x_lo,
                                     const float* dside, void* dz_hi, void* dz_lo, float* colsum, int n, int h, int w,
                                     int c, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(dpool_hi != nullptr && x_hi != nullptr && dz_hi != nullptr && n > 0 && h > 0 && w > 0 && c % 8 == 0);
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  const size_t total = static_cast<size_t>(n) * oh * ow * (c / 8);
  unpool_add_mask_kernel<<<grid_cap((total + 255) / 256, 8), 256, c * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(dpool_hi), static_cast<const __nv_bfloat16*>(dpool_lo),
      static_cast<const __nv_bfloat16*>(x_hi), static_cast<const __nv_bfloat16*>(x_lo), dside,
      static_cast<__nv_bfloat16*>(dz_hi), static_cast<__nv_bfloat16*>(dz_lo), colsum, n, h, w, c, oh, ow);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

extern "C" int osvos_channel_sum(const void* act_hi, const void* act_lo, float* out, size_t npix, int c,
                                 osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(act_hi != nullptr && out != nullptr && npix > 0 && c % 8 == 0 && c >= 8 && c <= 2048 &&
                  256 % (c / 8) == 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OSVOS_CHECK_CUDA(cudaMemsetAsync(out, 0, c * sizeof(float), stream));
  const int rows = 256 / (c / 8);
  const size_t blocks = (npix + rows - 1) / rows;
  channel_sum_kernel<<<grid_cap(blocks, 4), 256, 256 * 8 * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(act_hi), static_cast<const __nv_bfloat16*>(act_lo), npix, c, out);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

extern "C" int osvos_conv_first_bwd(const float* x_nchw, const void* dz_hi, const void* dz_lo, const float* w_oihw,
                                    float* dw, float* dx_nchw, int n, int h, int w, osvos_stream_t stream_) {
  OSVOS_CHECK_ARG(x_nchw != nullptr && dz_hi != nullptr && dw != nullptr && n > 0 && h > 0 && w > 0);
  OSVOS_CHECK_ARG(dx_nchw == nullptr || w_oihw != nullptr);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OSVOS_CHECK_CUDA(cudaMemsetAsync(dw, 0, 64 * 27 * sizeof(float), stream));
  const size_t npix = static_cast<size_t>(n) * h * w;
  conv_first_wgrad_kernel<<<grid_cap((npix + kFwPix - 1) / kFwPix, 4), 256, 0, stream>>>(
      x_nchw, static_cast<const __nv_bfloat16*>(dz_hi), static_cast<const __nv_bfloat16*>(dz_lo), dw, n, h, w);
  if (dx_nchw) {
    dim3 grid((w + 127) / 128, h, n);
    conv_first_dgrad_kernel<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(dz_hi),
                                                       static_cast<const __nv_bfloat16*>(dz_lo), w_oihw, dx_nchw, n, h,
                                                       w);
  }
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}
