// 3x3 convolution, halo-reuse implicit GEMM on CTA PAIRS (tcgen05 cta_group::2, sm_100a).
//
// Same algorithm as conv3x3_halo.cu, but two CTAs of a cluster cooperate on one UMMA of M = 256: each CTA owns
// one 16 x 8 pixel tile (its 128 accumulator rows live in its own TMEM), loads its own halo patch, and holds
// HALF of the weight slab (BLOCK_N / 2 rows); the leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2, which reads A and B halves from both CTAs' shared memory.  Motivation (measured on
// B200, scripts/microbench): a cta_group::1 tcgen05.mma of M = 128 costs ~85 cycles however small N is, so 1-CTA
// tiles with N <= 128 cannot exceed 75 % (N = 128) / 37 % (N = 64) of the tensor pipe; pairing doubles the work
// per instruction and halves each CTA's weight traffic, and allows N = 256 (128 rows of B per CTA).
//
// Barrier protocol (all mbarriers at identical smem offsets in both CTAs):
//   a_full / b_full : waited on by the leader's MMA warp only.  Both producers signal the LEADER's barrier: TMA
//                     loads use the .cta_group::2 form with the leader's barrier address (mapa), and each producer
//                     arrives once with its own expect_tx byte count (barrier count 2).
//   a_empty / b_empty / tmem_full : tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11 arrives on
//                     the local copy in both CTAs.
//   tmem_empty      : lives in the leader; the epilogue threads of both CTAs arrive on it (remote arrive from the
//                     peer), count = 2 x epilogue threads.
// TMEM is allocated / freed with the cta_group::2 forms by one warp of each CTA; cluster barriers bracket setup
// and teardown so that neither CTA touches the other's shared memory outside its lifetime.
#include "conv_common.cuh"

namespace osvos {

constexpr int kHalo2Rows = kTileH + 2;  // 18
constexpr int kHalo2Pitch = 10;         // packed rows: the UMMA swizzle is a function of the absolute smem address

template <int BLOCK_N, int PLANES>
struct Halo2Cfg {
  static constexpr int kHalfN = BLOCK_N / 2;                                     // B rows held by each CTA
  static constexpr int kABoxBytes = kHalo2Rows * kHalo2Pitch * 128;
  static constexpr int kAPlaneBytes = (kABoxBytes + 1023) / 1024 * 1024;
  static constexpr int kAStageBytes = PLANES * kAPlaneBytes;
  static constexpr int kAStages = 2;
  static constexpr int kBPlaneBytes = (kHalfN * 128 + 1023) / 1024 * 1024;
  static constexpr int kBBoxBytes = kHalfN * 128;
  static constexpr int kBStageBytes = PLANES * kBPlaneBytes;
  static constexpr int kBudget = 214 * 1024 - kAStages * kAStageBytes;
  static constexpr int kBStagesRaw = kBudget / kBStageBytes;
  static constexpr int kBStages = kBStagesRaw > 9 ? 9 : kBStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N) < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStageBytes + 1024 + 512;
  static_assert(kBStages >= 2, "weight ring too shallow");
  static_assert(kTmemCols <= 512, "TMEM overflow");
};

template <int BLOCK_N, int PLANES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + EpiCfg<BLOCK_N>::kThreads, 1)
conv3x3_halo2_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                     const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                     const ConvParams p) {
  using Cfg = Halo2Cfg<BLOCK_N, PLANES>;
  constexpr int SA = Cfg::kAStages, SB = Cfg::kBStages;
  constexpr int PITCH = kHalo2Pitch;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + SA * Cfg::kAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + SB * Cfg::kBStageBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + SA;
  uint64_t* b_full = bars + 2 * SA;
  uint64_t* b_empty = bars + 2 * SA + SB;
  uint64_t* tfull_bar = bars + 2 * SA + 2 * SB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int w_first = blockIdx.x >> 1, w_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x_hi);
    tma_prefetch_desc(&map_w_hi);
    if (PLANES == 2) {
      tma_prefetch_desc(&map_x_lo);
      tma_prefetch_desc(&map_w_lo);
    }
    for (int i = 0; i < SA; ++i) {
      mbar_init(&a_full[i], 2);   // one arrive.expect_tx per CTA of the pair (used in the leader only)
      mbar_init(&a_empty[i], 1);  // multicast commit
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(&b_full[i], 2);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EpiCfg<BLOCK_N>::kThreads);  // epilogue threads of both CTAs (leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one per CTA)
    // full barriers of the LEADER, addressed through the cluster window
    const uint32_t a_full_leader = mapa_shared(smem_u32(a_full), 0);
    const uint32_t b_full_leader = mapa_shared(smem_u32(b_full), 0);
    int a_stage = 0, b_stage = 0;
    uint32_t a_phase = 0, b_phase = 0;
    auto issue_a = [&](int w, int kc) {
      int nb, tx, ty, img;
      decode_pair(p, w, rank, nb, tx, ty, img);
      mbar_wait(&a_empty[a_stage], a_phase ^ 1);
      if (elect_one()) {
        uint8_t* st = smem_a + a_stage * Cfg::kAStageBytes;
        const uint32_t bar = a_full_leader + a_stage * 8;
        mbar_arrive_expect_tx_cluster(bar, PLANES * Cfg::kABoxBytes);
        tma_load_4d_2sm(&map_x_hi, bar, st, kc * kBlockK, tx * kTileW - 1, ty * kTileH - 1, img);
        if (PLANES == 2)
          tma_load_4d_2sm(&map_x_lo, bar, st + Cfg::kAPlaneBytes, kc * kBlockK, tx * kTileW - 1, ty * kTileH - 1, img);
      }
      __syncwarp();
      if (++a_stage == SA) {
        a_stage = 0;
        a_phase ^= 1;
      }
    };
    if (w_first < p.total_pairs) issue_a(w_first, 0);
    for (int w = w_first; w < p.total_pairs; w += w_stride) {
      const int nb = w % p.n_blocks;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        for (int tap = 0; tap < 9; ++tap) {
          if (tap == 3) {
            if (kc + 1 < p.k_chunks) issue_a(w, kc + 1);
            else if (w + w_stride < p.total_pairs) issue_a(w + w_stride, 0);
          }
          mbar_wait(&b_empty[b_stage], b_phase ^ 1);
          if (elect_one()) {
            uint8_t* st = smem_b + b_stage * Cfg::kBStageBytes;
            const uint32_t bar = b_full_leader + b_stage * 8;
            mbar_arrive_expect_tx_cluster(bar, PLANES * Cfg::kBBoxBytes);
            const int row0 = nb * BLOCK_N + rank * Cfg::kHalfN;  // this CTA's half of the weight slab
            tma_load_3d_2sm(&map_w_hi, bar, st, kc * kBlockK, row0, tap);
            if (PLANES == 2) tma_load_3d_2sm(&map_w_lo, bar, st + Cfg::kBPlaneBytes, kc * kBlockK, row0, tap);
          }
          __syncwarp();
          if (++b_stage == SB) {
            b_stage = 0;
            b_phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, BLOCK_N, /*bf16=*/true);
      int a_stage = 0, b_stage = 0;
      uint32_t a_phase = 0, b_phase = 0;
      int it = 0;
      for (int w = w_first; w < p.total_pairs; w += w_stride, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&a_full[a_stage], a_phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem_a + a_stage * Cfg::kAStageBytes);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - 3 * r;
            mbar_wait(&b_full[b_stage], b_phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_hi = a_base + (r * PITCH + s) * 128;
              const uint32_t a_lo = a_hi + Cfg::kAPlaneBytes;
              const uint32_t b_hi = smem_u32(smem_b + b_stage * Cfg::kBStageBytes);
              const uint64_t da_hi = make_smem_desc(a_hi, 16, PITCH * 128, kLayoutSW128);
              const uint64_t da_lo = make_smem_desc(a_lo, 16, PITCH * 128, kLayoutSW128);
              const uint64_t db_hi = make_smem_desc(b_hi, 16, 1024, kLayoutSW128);
              const uint64_t db_lo = make_smem_desc(b_hi + Cfg::kBPlaneBytes, 16, 1024, kLayoutSW128);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint64_t adv = static_cast<uint64_t>(k * 2);
                const uint32_t first = (kc | tap | k) != 0;
                if (PLANES == 2) {
                  umma_f16_2cta(tmem_d, da_lo + adv, db_hi + adv, idesc, first);
                  umma_f16_2cta(tmem_d, da_hi + adv, db_lo + adv, idesc, 1);
                  umma_f16_2cta(tmem_d, da_hi + adv, db_hi + adv, idesc, 1);
                } else {
                  umma_f16_2cta(tmem_d, da_hi + adv, db_hi + adv, idesc, first);
                }
              }
              umma_commit_2cta(&b_empty[b_stage], 3);
              if (tap == 8) {
                umma_commit_2cta(&a_empty[a_stage], 3);
                if (kc == p.k_chunks - 1) umma_commit_2cta(&tfull_bar[as], 3);
              }
            }
            __syncwarp();
            if (++b_stage == SB) {
              b_stage = 0;
              b_phase ^= 1;
            }
          }
          if (++a_stage == SA) {
            a_stage = 0;
            a_phase ^= 1;
          }
        }
      }
    }
  } else {
    conv_epilogue_loop<BLOCK_N, true>(p, tmem_base, tfull_bar, tempty_bar, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int PLANES>
static int launch_halo2(const osvos_conv3x3_args* a, cudaStream_t stream) {
  using Cfg = Halo2Cfg<BLOCK_N, PLANES>;
  ConvParams p;
  fill_conv_params(p, a, BLOCK_N);
  CUtensorMap mx_hi, mx_lo, mw_hi, mw_lo;
  {
    const uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
    const uint64_t strides[3] = {(uint64_t)a->cin * 2, (uint64_t)a->w * a->cin * 2,
                                 (uint64_t)a->h * a->w * a->cin * 2};
    const uint32_t box[4] = {kBlockK, kHalo2Pitch, kHalo2Rows, 1};
    int rc = encode_tensor_map(&mx_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 4, a->x_hi, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tensor_map(&mx_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 4, PLANES == 2 ? a->x_lo : a->x_hi, dims,
                           strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  int rc = encode_weight_maps(&mw_hi, &mw_lo, a, Cfg::kHalfN);  // box = half of the N block per CTA
  if (rc) return rc;
  auto kern = conv3x3_halo2_kernel<BLOCK_N, PLANES>;
  static uint64_t attr_done = 0;   // per instantiation: bit d = device d has the shared-memory opt-in
  OSVOS_CHECK_CUDA(ensure_dynamic_smem(kern, Cfg::kSmemBytes, &attr_done));
  const int sms = device_sm_count();
  int clusters = sms / 2;
  if (clusters > p.total_pairs) clusters = p.total_pairs;
  kern<<<2 * clusters, 64 + EpiCfg<BLOCK_N>::kThreads, Cfg::kSmemBytes, stream>>>(mx_hi, mx_lo, mw_hi, mw_lo, p);
  OSVOS_CHECK_CUDA(cudaGetLastError());
  return OSVOS_OK;
}

int conv3x3_halo2_dispatch(const osvos_conv3x3_args* a, cudaStream_t stream) {
  const bool fast = (a->flags & OSVOS_FLAG_FAST) != 0;
  if (a->cout == 16) return fast ? launch_halo2<16, 1>(a, stream) : launch_halo2<16, 2>(a, stream);
  if (a->cout == 64) return fast ? launch_halo2<64, 1>(a, stream) : launch_halo2<64, 2>(a, stream);
  if (a->cout % 256 == 0 && a->h * a->w * a->n >= 4096)   // enough pixel tiles to fill the chip with N = 256 pairs
    return fast ? launch_halo2<256, 1>(a, stream) : launch_halo2<256, 2>(a, stream);
  return fast ? launch_halo2<128, 1>(a, stream) : launch_halo2<128, 2>(a, stream);
}

}  // namespace osvos
