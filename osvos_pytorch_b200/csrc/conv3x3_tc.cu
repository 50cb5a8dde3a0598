// 3x3 / pad 1 / stride 1 convolution as a tcgen05 implicit GEMM (sm_100a).
//
//   D[pixel, co] = sum_{tap, ci} X[pixel + tap, ci] * Wt[tap][co][ci]
//
// M = pixels, tiled as patches of 16 rows x 8 px of one image (128 GEMM rows);
// N = BLOCK_N output channels; K = 9 taps x cin, consumed in blocks of 64
// channels of one tap.  The A operand of a K block is the input patch shifted by
// the tap: one 4-D TMA box {64 ch, 8 px, 16 rows, 1 image} whose out-of-image
// part (the conv zero padding, and ragged right/bottom tiles) is zero-filled by
// the TMA unit, so im2col never exists in memory.  The box lands in shared
// memory as 128 rows x 128 B with the 128-byte swizzle, which is exactly the
// canonical K-major SWIZZLE_128B UMMA operand.  B is the packed weight slab
// [tap][co][ci] (box {64, BLOCK_N, 1}).
//
// Precision: activations/weights are split bf16 (v ~= hi + lo).  Exact mode
// (PLANES == 2) issues hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator
// (~2^-17 relative operand error, i.e. fp32-class results); fast mode
// (PLANES == 1) issues hi*hi only.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
// alloc/dealloc), warps 2..5 = epilogue (TMEM -> registers -> bias / ReLU /
// mask / split -> global).  Persistent CTAs walk tiles round-robin; two TMEM
// accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Replaces nn.Conv2d(k=3, p=1)[+ReLU] of the reference
// (networks/vgg_osvos.py:41,142-143) and, with flipped weights, its dgrad.
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"

namespace osvos {

template <int BLOCK_N, int PLANES>
struct ConvCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = PLANES * (kABytes + kBBytes);
  static constexpr int kStagesRaw = (212 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N) < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "pipeline too shallow");
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment");
};

template <int BLOCK_N, int PLANES>
__global__ void __launch_bounds__(64 + EpiCfg<BLOCK_N>::kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                  const ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N, PLANES>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x_hi);
    tma_prefetch_desc(&map_w_hi);
    if (PLANES == 2) {
      tma_prefetch_desc(&map_x_lo);
      tma_prefetch_desc(&map_w_lo);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EpiCfg<BLOCK_N>::kThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = 9 * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Warp-uniform control flow; one elected lane issues.  (Issuing from a divergent `if (lane == 0)` branch
    // makes the compiler wrap every UTMALDG / UTCHMMA in an ELECT/R2UR/BRA.U.ANY loop: ~350 cycles per step.)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nb, tx, ty, img;
        decode_tile(p, tile, nb, tx, ty, img);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - 3 * r;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* st = smem + stage * Cfg::kStageBytes;
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tma_load_4d(&map_x_hi, &full_bar[stage], st, kc * kBlockK, tx * kTileW + s - 1, ty * kTileH + r - 1, img);
              tma_load_3d(&map_w_hi, &full_bar[stage], st + PLANES * kABytes, kc * kBlockK, nb * BLOCK_N, tap);
              if (PLANES == 2) {
                tma_load_4d(&map_x_lo, &full_bar[stage], st + kABytes, kc * kBlockK, tx * kTileW + s - 1,
                            ty * kTileH + r - 1, img);
                tma_load_3d(&map_w_lo, &full_bar[stage], st + 2 * kABytes + Cfg::kBBytes, kc * kBlockK,
                            nb * BLOCK_N, tap);
              }
            }
            __syncwarp();
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N, /*bf16=*/true);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t b_hi = a_hi + PLANES * kABytes;
            const uint64_t da_hi = make_smem_desc(a_hi, 16, 1024, kLayoutSW128);
            const uint64_t db_hi = make_smem_desc(b_hi, 16, 1024, kLayoutSW128);
            const uint64_t da_lo = make_smem_desc(a_hi + kABytes, 16, 1024, kLayoutSW128);
            const uint64_t db_lo = make_smem_desc(b_hi + Cfg::kBBytes, 16, 1024, kLayoutSW128);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t adv = static_cast<uint64_t>(k * 2);  // 32 bytes >> 4
              if (PLANES == 2) {
                umma_f16(tmem_d, da_lo + adv, db_hi + adv, idesc, (kb | k) != 0);
                umma_f16(tmem_d, da_hi + adv, db_lo + adv, idesc, 1);
                umma_f16(tmem_d, da_hi + adv, db_hi + adv, idesc, 1);
              } else {
                umma_f16(tmem_d, da_hi + adv, db_hi + adv, idesc, (kb | k) != 0);
              }
            }
            umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    conv_epilogue_loop<BLOCK_N>(p, tmem_base, tfull_bar, tempty_bar, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// --------------------------------------------------------------------- host
template <int BLOCK_N, int PLANES>
static int launch_conv(const osvos_conv3x3_args* a, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, PLANES>;
  ConvParams p;
  fill_conv_params(p, a, BLOCK_N);

  CUtensorMap mx_hi, mx_lo, mw_hi, mw_lo;
  {
    const uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
    const uint64_t strides[3] = {(uint64_t)a->cin * 2, (uint64_t)a->w * a->cin * 2,
                                 (uint64_t)a->h * a->w * a->cin * 2};
    const uint32_t box[4] = {kBlockK, kTileW, kTileH, 1};
    int rc = encode_tensor_map(&mx_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 4, a->x_hi, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tensor_map(&mx_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 4, PLANES == 2 ? a->x_lo : a->x_hi, dims,
                           strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    int rc = encode_weight_maps(&mw_hi, &mw_lo, a, BLOCK_N);
    if (rc) return rc;
  }

  auto kern = conv3x3_tc_kernel<BLOCK_N, PLANES>;
  static uint64_t attr_done = 0;   // per instantiation: bit d = device d has the shared-memory opt-in
  OSVOS_CHECK_CUDA(ensure_dynamic_smem(kern, Cfg::kSmemBytes, &attr_done));
  const int sms = device_sm_count();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  kern<<<grid, 64 + EpiCfg<BLOCK_N>::kThreads, Cfg::kSmemBytes, stream>>>(mx_hi, mx_lo, mw_hi, mw_lo, p);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

static int check_conv_args(const osvos_conv3x3_args* a) {
  OSVOS_CHECK_ARG(a != nullptr);
  OSVOS_CHECK_ARG(a->x_hi != nullptr && a->w_packed != nullptr);
  OSVOS_CHECK_ARG(a->n > 0 && a->h > 0 && a->w > 0);
  OSVOS_CHECK_ARG(a->cin > 0 && a->cin % 64 == 0);
  OSVOS_CHECK_ARG(a->cout == 16 || a->cout % 64 == 0);
  OSVOS_CHECK_ARG((a->flags & OSVOS_FLAG_FAST) || a->x_lo != nullptr);
  OSVOS_CHECK_ARG(a->y_hi != nullptr || a->y_f32 != nullptr || a->pq != nullptr || a->pool_hi != nullptr);
  OSVOS_CHECK_ARG(!(a->flags & OSVOS_FLAG_RELU_MASK) || a->mask_hi != nullptr);
  OSVOS_CHECK_ARG(a->pq == nullptr || (a->cout == 16 && a->proj_w != nullptr));
  OSVOS_CHECK_ARG((a->pool_hi == nullptr && a->colsum == nullptr) || a->cout >= 64);
  OSVOS_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x_hi) & 15) == 0);
  OSVOS_CHECK_ARG((reinterpret_cast<uintptr_t>(a->w_packed) & 15) == 0);
  OSVOS_CHECK_ARG((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0);
  OSVOS_CHECK_ARG(a->k_valid >= 0 && a->k_valid <= 64 && a->k_valid % 16 == 0);
  return OSVOS_OK;
}

}  // namespace osvos

using namespace osvos;

extern "C" size_t osvos_conv3x3_splitk_workspace_bytes(int n, int h, int w, int cin, int cout) {
  if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return 0;
  const char* impl = getenv("OSVOS_CONV_IMPL");
  if (impl != nullptr && (strcmp(impl, "tap") == 0 || strcmp(impl, "halo2") == 0)) return 0;
  return conv3x3_splitk_workspace_bytes(n, h, w, cin, cout);
}

extern "C" int osvos_conv3x3(const osvos_conv3x3_args* a, osvos_stream_t stream_) {
  int rc = check_conv_args(a);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool fast = (a->flags & OSVOS_FLAG_FAST) != 0;
  // Default: halo-reuse kernel (conv3x3_halo.cu) with packed rows (pitch 10) - validated on B200: the UMMA
  // swizzle is a function of the absolute smem address, so shifted descriptor starts need no base offset.
  // OSVOS_CONV_IMPL=tap selects the one-box-per-tap kernel below (kept as a cross-check).
  const char* impl = getenv("OSVOS_CONV_IMPL");
  // side_prep shape (16 outputs, fp32 features / projections only): nine-taps-along-N kernel (side_conv.cu)
  if (a->cout == 16 && a->y_hi == nullptr && !(a->flags & OSVOS_FLAG_RELU_MASK) && a->colsum == nullptr && impl == nullptr) {
    const char* side = getenv("OSVOS_SIDE_IMPL");
    if (side == nullptr || strcmp(side, "generic") != 0) return side_conv_dispatch(a, stream);
  }
  if (impl != nullptr && strcmp(impl, "halo2") == 0) return conv3x3_halo2_dispatch(a, stream);
  if (impl == nullptr || strcmp(impl, "tap") != 0) {
    const char* ps = getenv("OSVOS_HALO_PITCH");
    const char* bo = getenv("OSVOS_HALO_BO");
    const int pitch = (ps && atoi(ps) == 16) ? 16 : 10;
    return conv3x3_halo_dispatch(a, stream, pitch, bo ? atoi(bo) : 0);
  }
  if (a->cout == 16) return fast ? launch_conv<16, 1>(a, stream) : launch_conv<16, 2>(a, stream);
  if (a->cout == 64) return fast ? launch_conv<64, 1>(a, stream) : launch_conv<64, 2>(a, stream);
  return fast ? launch_conv<128, 1>(a, stream) : launch_conv<128, 2>(a, stream);
}
