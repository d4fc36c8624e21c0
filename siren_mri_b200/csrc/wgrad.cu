// Weight-gradient kernel on tcgen05:  dW_l[out, in] = sum over coordinates (and jet streams) of
// adj_l[n, out] * act_{l-1}[n, in]   -- the dW = dz^T h term of the autograd backward that
// training.py:91 triggers through modules.py:25.
//
// The contraction runs over coordinates, so both operands are "MN-major" for the tensor core:
// TMA drops [KC coordinates x 64 features] boxes (128-byte swizzle) straight from the row-major
// planes and the UMMA descriptors read them transposed; no transposed copies exist in HBM.
// One CTA owns a whole 256x256 fp32 accumulator (2 x 256 TMEM columns) for a slice of the
// coordinates and flushes it once with vector red.add.
#include "common.cuh"
#include "ptx.cuh"

namespace siren {

namespace {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kStages = 3;

template <bool SPLIT>
struct WgCfg {
  static constexpr int KC = SPLIT ? 32 : 64;            // coordinates per pipeline stage
  static constexpr int OPER = KC * H * 2;               // bytes of one operand block (all 256 features)
  static constexpr int STAGE = OPER * (SPLIT ? 4 : 2);  // 64 KB either way
  static constexpr int SMEM = kStages * STAGE + 1024 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

struct Item {
  int layer, row0, row1, task;
};

__device__ __forceinline__ Item decode_item(const WgradParams& p, int idx) {
  Item it;
  it.layer = idx % p.n_layers;
  int rest = idx / p.n_layers;
  const int grp = rest / p.slices;
  const int sl = rest % p.slices;
  const int rows_group = p.per_task ? p.rows_per_task : p.R;
  const int tiles = rows_group / TILE_M;
  const int base = tiles / p.slices, rem = tiles % p.slices;
  const int t0 = sl * base + (sl < rem ? sl : rem);
  const int n = base + (sl < rem ? 1 : 0);
  it.row0 = grp * rows_group + t0 * TILE_M;
  it.row1 = it.row0 + n * TILE_M;
  it.task = p.per_task ? grp : 0;
  return it;
}

template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgCfg<SPLIT>;
  constexpr int KC = Cfg::KC;
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(TILE_M, 256, 1, 1);
  constexpr uint32_t LBO = KC * 128;     // bytes between 64-feature blocks
  constexpr uint32_t SBO = 1024;         // bytes between groups of 8 coordinates

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = bars + 2 * kStages + 1;
  uint64_t* ready = bars + 2 * kStages + 2;         // [kStages] phase_b: the B block has been turned into sines
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = p.per_task ? p.tasks : 1;
  const int n_items = p.n_layers * groups * p.slices;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
      ptx::mbar_init(&ready[i], kEpiWarps);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
        const Item it = decode_item(p, idx);
        const CUtensorMap* mA_hi = &p.tmA_hi[it.layer];
        const CUtensorMap* mB_hi = &p.tmB_hi[it.layer];
        const CUtensorMap* mA_lo = &p.tmA_lo[it.layer];
        const CUtensorMap* mB_lo = &p.tmB_lo[it.layer];
        for (int r = it.row0; r < it.row1; r += KC)
          for (int s = 0; s < p.S; ++s) {
            ptx::mbar_wait(&empty[stage], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE);
            uint8_t* st = smem + stage * Cfg::STAGE;
            const int y = s * p.R + r;
#pragma unroll
            for (int fb = 0; fb < 4; ++fb) {
              ptx::tma_load_2d(st + fb * LBO, mA_hi, &full[stage], fb * 64, y);
              ptx::tma_load_2d(st + Cfg::OPER + fb * LBO, mB_hi, &full[stage], fb * 64, y);
              if (SPLIT) {
                ptx::tma_load_2d(st + 2 * Cfg::OPER + fb * LBO, mA_lo, &full[stage], fb * 64, y);
                ptx::tma_load_2d(st + 3 * Cfg::OPER + fb * LBO, mB_lo, &full[stage], fb * 64, y);
              }
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++local) {
      const Item it = decode_item(p, idx);
      ptx::mbar_wait(acc_empty, (uint32_t(local) & 1u) ^ 1u);
      ptx::tc_fence_after();
      bool first = true;
      for (int r = it.row0; r < it.row1; r += KC)
        for (int s = 0; s < p.S; ++s) {
          ptx::mbar_wait(p.phase_b ? &ready[stage] : &full[stage], phase);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t base = ptx::smem_u32(smem + stage * Cfg::STAGE);
#pragma unroll
            for (int mh = 0; mh < 2; ++mh) {
              const uint32_t d_tmem = tmem_base + uint32_t(mh * 256);
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint32_t a_hi = base + 2 * mh * LBO + ks * 2048;
                const uint32_t b_hi = base + Cfg::OPER + ks * 2048;
                const uint32_t acc0 = (first && ks == 0) ? 0u : 1u;
                ptx::umma_bf16(d_tmem, ptx::umma_smem_desc(a_hi, LBO, SBO), ptx::umma_smem_desc(b_hi, LBO, SBO),
                               IDESC, acc0);
                if (SPLIT) {
                  const uint32_t a_lo = base + 2 * Cfg::OPER + 2 * mh * LBO + ks * 2048;
                  const uint32_t b_lo = base + 3 * Cfg::OPER + ks * 2048;
                  ptx::umma_bf16(d_tmem, ptx::umma_smem_desc(a_hi, LBO, SBO),
                                 ptx::umma_smem_desc(b_lo, LBO, SBO), IDESC, 1u);
                  ptx::umma_bf16(d_tmem, ptx::umma_smem_desc(a_lo, LBO, SBO),
                                 ptx::umma_smem_desc(b_hi, LBO, SBO), IDESC, 1u);
                }
              }
            }
            ptx::umma_commit(&empty[stage]);
          }
          __syncwarp();
          first = false;
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      if (lane == 0) ptx::umma_commit(acc_full);
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    const int e = warp - kEpiWarp0;
    const int q = warp & 3;
    const int chalf = e >> 2;
    int local = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++local) {
      const Item it = decode_item(p, idx);
      if (!SPLIT && p.phase_b) {
        // The B planes carry the layer input's phase theta (fp16): the operand the MMA needs is sin(theta) in
        // bf16 -- same element size, same place.  The eight warps convert each staged 32 KB block in shared
        // memory (flat, 16 bytes per thread and step) and hand the stage to the MMA warp.
        const int tid = e * 32 + lane;
        // While a stage is in their hands the same warps also take the column sums of its adjoint block (the A
        // operand, [4 feature blocks][64 coordinates][64 features], 128-byte swizzle): the bias gradient
        // db_l = sum over coordinates of zbar_l.  Thread = one pair of adjacent columns and one half of the rows.
        float* dbp = p.db[it.layer];
        const int cp = tid & 127, rhalf = tid >> 7;               // column pair 0..127, row half 0..1
        const uint32_t db_off = uint32_t(cp >> 5) * LBO + (uint32_t(cp & 3) << 2);   // feature block, word in its unit
        const uint32_t db_unit = uint32_t((cp & 31) >> 2);        // 16-byte unit inside the 128-byte row
        float bs0 = 0.f, bs1 = 0.f;
        for (int r = it.row0; r < it.row1; r += KC) {
          ptx::mbar_wait(&full[stage], phase);
          const uint32_t blk = ptx::smem_u32(smem + stage * Cfg::STAGE + Cfg::OPER);
          if (dbp) {
            const uint32_t ablk = ptx::smem_u32(smem + stage * Cfg::STAGE) + db_off;
#pragma unroll
            for (int rb = 0; rb < KC / 2; rb += 8) {       // eight loads in flight before the first is consumed
              uint32_t u[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = rhalf * (KC / 2) + rb + i;
                asm volatile("ld.shared.b32 %0, [%1];"
                             : "=r"(u[i])
                             : "r"(ablk + uint32_t(row) * 128u + ((db_unit ^ uint32_t(row & 7)) << 4)));
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                bs0 += bf16_lo_f(u[i]);
                bs1 += bf16_hi_f(u[i]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < Cfg::OPER / (kEpiWarps * 32 * 16); ++i) {
            const uint32_t a = blk + uint32_t(i * kEpiWarps * 32 + tid) * 16u;
            uint32_t w[4];
            ptx::ld_shared_v4(a, w[0], w[1], w[2], w[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 th = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
              w[j] = pack_bf16(__sinf(th.x), __sinf(th.y));
            }
            ptx::st_shared_v4(a, w[0], w[1], w[2], w[3]);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&ready[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (dbp && it.row1 > it.row0) {
          float* dst = dbp + size_t(it.task) * H + 2 * cp;
          atomicAdd(dst, bs0);
          atomicAdd(dst + 1, bs1);
        }
      }
      ptx::mbar_wait(acc_full, uint32_t(local) & 1u);
      ptx::tc_fence_after();
      float* dW = p.dW[it.layer] + size_t(it.task) * H * H;
      const bool has_rows = it.row1 > it.row0;
#pragma unroll
      for (int mh = 0; mh < 2; ++mh) {
        const int orow = mh * 128 + q * 32 + lane;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int col = chalf * 128 + cc * 32;
          float v[32];
          ptx::tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(mh * 256 + col),
                         reinterpret_cast<uint32_t*>(v));
          ptx::tmem_wait_ld();
          if (has_rows) {
            float* dst = dW + size_t(orow) * H + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) red_add_v4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int wgrad_kc(bool split) { return split ? 32 : 64; }

cudaError_t launch_wgrad(const WgradParams& p, bool split, int num_sms, cudaStream_t stream) {
  const int groups = p.per_task ? p.tasks : 1;
  const int n_items = p.n_layers * groups * p.slices;
  int grid = n_items < num_sms ? n_items : num_sms;
  if (grid < 1) return cudaSuccess;
  if (split) {
    static bool set = false;
    if (!set) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           WgCfg<true>::SMEM);
      if (e != cudaSuccess) return e;
      set = true;
    }
    wgrad_kernel<true><<<grid, kThreads, WgCfg<true>::SMEM, stream>>>(p);
  } else {
    static bool set = false;
    if (!set) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           WgCfg<false>::SMEM);
      if (e != cudaSuccess) return e;
      set = true;
    }
    wgrad_kernel<false><<<grid, kThreads, WgCfg<false>::SMEM, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace siren
