// Fast-path row-GEMM kernels (bf16 mode, value stream only): same tcgen05 / TMA / TMEM main loop
// as gemm_rows.cu, with an epilogue that never touches global memory with row-strided accesses:
//   * results are packed to bf16, written to a 128-byte-swizzled shared-memory tile and leave
//     the SM as TMA bulk stores (fully coalesced 128-byte rows);
//   * the cosine stash the backward needs arrives the same way (TMA load, double buffered);
//   * reductions that would otherwise need another pass over HBM are taken from the staged tile
//     while the TMA store drains it: the bias gradient (column sums of zbar) and, for the first
//     layer, dW0 = zbar0^T x;  the forward of the top hidden layer also evaluates the outermost
//     linear layer (a 256-long dot product per coordinate) from the sine values in registers.
//
//   forward :  h' = sin(w0 (h W^T + b)), c' = cos(.)        (modules.py:25-26, 38)
//              [top layer]  y = h' W_L^T + b_L               (modules.py:25-26, outermost linear)
//   backward:  zbar_{l-1} = w0 c_{l-1} * (zbar_l W_l);  db_{l-1} = sum_n zbar_{l-1}
//              [l-1 = 0]    dW0 = zbar_0^T x                 (autograd of the above, training.py:91)
#include "common.cuh"
#include "ptx.cuh"
#include "simt.h"

namespace siren {

namespace {

// threads = 4 control warps + NSUB epilogue warps per TMEM lane quadrant (NSUB = 2 or 4)
constexpr int kEpiWarp0 = 4;
constexpr int BN = 256;
constexpr int NST = 3;
constexpr int A_STAGE = TILE_M * 128;          // 16 KB
constexpr int B_BYTES = 4 * BN * 128;          // 128 KB
constexpr int STG = TILE_M * 128;              // one staged [128 x 64] bf16 tile
constexpr int QSLICE = 32 * 128;               // a quadrant's 32 rows of it
constexpr int MISC = 2048;                     // barriers + the tile's coordinates (backward, layer 0)
constexpr int SMEM_FAST = B_BYTES + NST * A_STAGE + 3 * STG + MISC + 1024;
static_assert(SMEM_FAST <= 232448, "shared memory budget");

struct TileRange {
  int t0, t1;
};
__device__ __forceinline__ TileRange cta_tiles(int tiles_m, int g, int G) {
  int base = tiles_m / G, rem = tiles_m % G;
  int t0 = g * base + (g < rem ? g : rem);
  int n = base + (g < rem ? 1 : 0);
  return {t0, t0 + n};
}

__device__ __forceinline__ void red_add(float* dst, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
}

// write CW floats (this thread's row, columns sub*CW..) as bf16 into a swizzled [128][128 B] tile
template <int CW>
__device__ __forceinline__ void stage_row(uint32_t tile_addr, int row, int sub, const float* v) {
  const uint32_t row_addr = tile_addr + uint32_t(row) * 128u;
#pragma unroll
  for (int jj = 0; jj < CW / 8; ++jj) {
    const uint32_t chunk = uint32_t(sub * (CW / 8) + jj) ^ uint32_t(row & 7);
    ptx::st_shared_v4(row_addr + (chunk << 4), pack_bf16(v[8 * jj + 0], v[8 * jj + 1]),
                      pack_bf16(v[8 * jj + 2], v[8 * jj + 3]), pack_bf16(v[8 * jj + 4], v[8 * jj + 5]),
                      pack_bf16(v[8 * jj + 6], v[8 * jj + 7]));
  }
}
template <int CW>
__device__ __forceinline__ void unstage_row(uint32_t tile_addr, int row, int sub, float* v) {
  const uint32_t row_addr = tile_addr + uint32_t(row) * 128u;
#pragma unroll
  for (int jj = 0; jj < CW / 8; ++jj) {
    const uint32_t chunk = uint32_t(sub * (CW / 8) + jj) ^ uint32_t(row & 7);
    uint32_t a, b, c, d;
    ptx::ld_shared_v4(row_addr + (chunk << 4), a, b, c, d);
    v[8 * jj + 0] = bf16_lo_f(a); v[8 * jj + 1] = bf16_hi_f(a);
    v[8 * jj + 2] = bf16_lo_f(b); v[8 * jj + 3] = bf16_hi_f(b);
    v[8 * jj + 4] = bf16_lo_f(c); v[8 * jj + 5] = bf16_hi_f(c);
    v[8 * jj + 6] = bf16_lo_f(d); v[8 * jj + 7] = bf16_hi_f(d);
  }
}

template <int MODE, int NSUB>   // MODE 0 forward, 1 backward; NSUB column slices (warps) per quadrant
__global__ void __launch_bounds__(128 + NSUB * 128, 1) rows_fast_kernel(const __grid_constant__ RowsFastParams p) {
  constexpr int CW = 64 / NSUB;            // columns per thread per 64-column chunk
  constexpr int QTHREADS = NSUB * 32;      // threads per lane quadrant
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(TILE_M, BN, 0, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = smem + B_BYTES;
  uint8_t* sStg = sA + NST * A_STAGE;                    // 3 staged tiles
  uint8_t* sMisc = sStg + 3 * STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMisc);
  uint64_t* full = bars;                                  // [NST]
  uint64_t* empty = bars + NST;                           // [NST]
  uint64_t* b_full = bars + 2 * NST;
  uint64_t* b_empty = bars + 2 * NST + 1;
  uint64_t* acc_full = bars + 2 * NST + 2;                // [2]
  uint64_t* acc_empty = bars + 2 * NST + 4;               // [2]
  uint64_t* c_full = bars + 2 * NST + 6;                  // [4 quadrants][2]  (backward: cosine slices)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 14);
  float* sX = reinterpret_cast<float*>(sMisc + 256);      // [128][3]  coordinates of the tile (backward, layer 0)
  float* sY = reinterpret_cast<float*>(sStg + 2 * STG);   // [128][NSUB][2] partial last-layer dots (forward only: 3rd tile unused)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = p.R / TILE_M;
  const TileRange tr = cta_tiles(tiles_m, blockIdx.x, gridDim.x);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
    ptx::prefetch_tmap(&p.tmO0);
    ptx::prefetch_tmap(&p.tmO1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(b_empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 4 * NSUB);
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&c_full[i], 1);
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
    // ===================== TMA producer (A tiles, weight block) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int cur_task = -1;
      uint32_t b_gen = 0;
      for (int t = tr.t0; t < tr.t1; ++t) {
        const int row0 = t * TILE_M;
        const int task = p.per_task ? row0 / p.rows_per_task : 0;
        if (task != cur_task) {
          if (b_gen > 0) ptx::mbar_wait(b_empty, (b_gen - 1) & 1u);
          ++b_gen;
          ptx::mbar_arrive_expect_tx(b_full, B_BYTES);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) ptx::tma_load_2d(sB + kc * BN * 128, &p.tmB, b_full, kc * KCHUNK, task * H);
          cur_task = task;
        }
        for (int kc = 0; kc < 4; ++kc) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[stage], A_STAGE);
          ptx::tma_load_2d(sA + stage * A_STAGE, &p.tmA, &full[stage], kc * KCHUNK, row0);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int cur_task = -1;
    uint32_t b_loads = 0;
    int local = 0;
    for (int t = tr.t0; t < tr.t1; ++t, ++local) {
      const int row0 = t * TILE_M;
      const int task = p.per_task ? row0 / p.rows_per_task : 0;
      const int a = local & 1;
      ptx::mbar_wait(&acc_empty[a], ((uint32_t(local >> 1)) & 1u) ^ 1u);
      if (task != cur_task) {
        ptx::mbar_wait(b_full, b_loads & 1u);
        ++b_loads;
        cur_task = task;
      }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(a * BN);
      for (int kc = 0; kc < 4; ++kc) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = ptx::smem_u32(sA + stage * A_STAGE);
          const uint32_t b_addr = ptx::smem_u32(sB + kc * BN * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16(d_tmem, ptx::umma_smem_desc(a_addr + ks * 32, 16, 1024),
                           ptx::umma_smem_desc(b_addr + ks * 32, 16, 1024), IDESC, (kc | ks) ? 1u : 0u);
          ptx::umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
      const int next_task = (t + 1 < tr.t1) ? (p.per_task ? (row0 + TILE_M) / p.rows_per_task : 0) : -1;
      if (lane == 0) {
        ptx::umma_commit(&acc_full[a]);
        if (next_task != task) ptx::umma_commit(b_empty);
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (4 lane quadrants x NSUB column slices) ===========
    // Each quadrant (NSUB warps on the same SM sub-partition) owns rows 32q..32q+31 of the tile:
    // its own slice of the staging tiles, its own TMA boxes (64 cols x 32 rows) and its own named
    // barrier, so the four quadrants never wait for one another.
    const int e = warp - kEpiWarp0;
    const int q = warp & 3;
    const int sub = e >> 2;                                // CW-column slice of each 64-column chunk
    const int tid_q = sub * 32 + lane;                     // thread index inside the quadrant
    const int row_t = q * 32 + lane;                       // row inside the tile
    const bool dma = (tid_q == 0);
    const int bar_id = 1 + q;
    const uint32_t stg0 = ptx::smem_u32(sStg), stg1 = stg0 + STG, stg2 = stg0 + 2 * STG;
    uint8_t* const q_stg = sStg + q * QSLICE;              // this quadrant's slice of staged tile 0
    uint64_t* const cq_full = c_full + 2 * q;              // [2] cosine-slice barriers of this quadrant
    const float w0 = p.w0, w0_rev = p.w0 * 0.15915494309189535f;

    // backward: column-sum ownership -- 2 adjacent columns of each 64-column chunk, 32/NSUB rows
    const int cpair = tid_q & 31, rgrp = tid_q >> 5;
    float cs[4][8];            // db partials  [chunk][col of this thread's 8-column group]
    float cw[4][2][3];         // dW0 partials [chunk][col][i]
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 8; ++b) cs[a][b] = 0.f;
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < 3; ++i) cw[a][b][i] = 0.f;
    }
    int acc_task = -1;
    auto flush_sums = [&](int task) {
      if (MODE != 1 || task < 0) return;
      const int wt = p.per_task ? task : 0;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (p.db) {
          // lanes l, l^8, l^16, l^24 own the same 8 columns (different rows): combine them first --
          // same-address atomics serialise in L2, so every partial removed here is time saved
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            float v = cs[cc][b];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 8) red_add(p.db + size_t(wt) * H + cc * 64 + lane * 8 + b, v);
            cs[cc][b] = 0.f;
          }
        }
        if (p.dW0) {
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int i = 0; i < 3; ++i)
              if (i < p.d) {
                red_add(p.dW0 + (size_t(wt) * H + cc * 64 + cpair * 2 + b) * p.d + i, cw[cc][b][i]);
                cw[cc][b][i] = 0.f;
              }
        }
      }
    };

    uint32_t g = 0;            // backward: chunk counter (cosine slice double buffering)
    if (MODE == 1 && dma && tr.t0 < tr.t1) {
      ptx::mbar_arrive_expect_tx(&cq_full[0], QSLICE);
      ptx::tma_load_2d(q_stg, &p.tmO1, &cq_full[0], 0, tr.t0 * TILE_M + q * 32);
    }

    int local = 0;
    for (int t = tr.t0; t < tr.t1; ++t, ++local) {
      const int row0 = t * TILE_M;
      const int task = row0 / p.rows_per_task;          // true task of the rows (coordinates, y); weights use per_task ? task : 0
      const int a = local & 1;
      const int n_row = row0 + row_t - task * p.rows_per_task;      // coordinate index inside the task
      if (MODE == 1) {
        const int wtask = p.per_task ? task : 0;      // the sums are per weight set
        if (wtask != acc_task) {
          flush_sums(acc_task);
          acc_task = wtask;
        }
        if (p.dW0) {
          ptx::named_bar_sync(bar_id, QTHREADS);   // the previous tile's readers of this quadrant's sX are done
          if (sub == 0) {                       // stage the coordinates of this quadrant's rows (zero for pad rows)
            float* xs = sX + row_t * 3;
            for (int i = 0; i < 3; ++i)
              xs[i] = (i < p.d && n_row < p.n) ? __ldg(p.x + (size_t(task) * p.n + n_row) * p.d + i) : 0.f;
          }
        }
      }
      ptx::mbar_wait(&acc_full[a], (uint32_t(local >> 1)) & 1u);
      ptx::tc_fence_after();
      float ydot[2] = {0.f, 0.f};

#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int colb = cc * 64 + sub * CW;
        float v[CW];
        ptx::tmem_ld<CW>(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(a * BN + colb), reinterpret_cast<uint32_t*>(v));
        ptx::tmem_wait_ld();
        if (cc == 3) {            // accumulator fully read: hand it back to the MMA warp early
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[a]);
        }
        if (MODE == 0) {
          float cosv[CW];
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + (p.per_task ? task * H : 0) + colb);
          if (p.no_stash) {
#pragma unroll
            for (int j4 = 0; j4 < CW / 4; ++j4) {
              const float4 bb = __ldg(bias4 + j4);
              v[4 * j4 + 0] = sin_rev((v[4 * j4 + 0] + bb.x) * w0_rev);
              v[4 * j4 + 1] = sin_rev((v[4 * j4 + 1] + bb.y) * w0_rev);
              v[4 * j4 + 2] = sin_rev((v[4 * j4 + 2] + bb.z) * w0_rev);
              v[4 * j4 + 3] = sin_rev((v[4 * j4 + 3] + bb.w) * w0_rev);
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < CW / 4; ++j4) {
              const float4 bb = __ldg(bias4 + j4);
              sincos_rev((v[4 * j4 + 0] + bb.x) * w0_rev, &v[4 * j4 + 0], &cosv[4 * j4 + 0]);
              sincos_rev((v[4 * j4 + 1] + bb.y) * w0_rev, &v[4 * j4 + 1], &cosv[4 * j4 + 1]);
              sincos_rev((v[4 * j4 + 2] + bb.z) * w0_rev, &v[4 * j4 + 2], &cosv[4 * j4 + 2]);
              sincos_rev((v[4 * j4 + 3] + bb.w) * w0_rev, &v[4 * j4 + 3], &cosv[4 * j4 + 3]);
            }
          }
          if (p.fuse_last) {
            for (int oi = 0; oi < p.o; ++oi) {
              const float4* wl = reinterpret_cast<const float4*>(p.WL + (size_t(p.per_task ? task : 0) * p.o + oi) * H + colb);
              float acc = 0.f;
#pragma unroll
              for (int j4 = 0; j4 < CW / 4; ++j4) {
                const float4 ww = __ldg(wl + j4);
                acc = fmaf(v[4 * j4 + 0], ww.x, acc); acc = fmaf(v[4 * j4 + 1], ww.y, acc);
                acc = fmaf(v[4 * j4 + 2], ww.z, acc); acc = fmaf(v[4 * j4 + 3], ww.w, acc);
              }
              ydot[oi] += acc;
            }
          }
          // inference of the top layer with the outermost linear fused: nothing leaves but y
          const bool store_sine = !(p.no_stash && p.fuse_last);
          if (store_sine) {
            if (dma) ptx::bulk_wait_read_all();
            ptx::named_bar_sync(bar_id, QTHREADS);
            stage_row<CW>(stg0, row_t, sub, v);
            if (!p.no_stash) stage_row<CW>(stg1, row_t, sub, cosv);
            ptx::fence_proxy_async();
            ptx::named_bar_sync(bar_id, QTHREADS);
            if (dma) {
              ptx::tma_store_2d(&p.tmO0, q_stg, cc * 64, row0 + q * 32);
              if (!p.no_stash) ptx::tma_store_2d(&p.tmO1, q_stg + STG, cc * 64, row0 + q * 32);
              ptx::bulk_commit();
            }
          }
        } else {
          // prefetch the next cosine slice, then consume this one
          if (dma) {
            const bool more = (cc < 3) || (t + 1 < tr.t1);
            if (more) {
              const int ncc = (cc + 1) & 3;
              const int nrow0 = (cc < 3) ? row0 : row0 + TILE_M;
              const uint32_t nb = (g + 1) & 1u;
              ptx::mbar_arrive_expect_tx(&cq_full[nb], QSLICE);
              ptx::tma_load_2d(q_stg + nb * STG, &p.tmO1, &cq_full[nb], ncc * 64, nrow0 + q * 32);
            }
          }
          ptx::mbar_wait(&cq_full[g & 1u], (g >> 1) & 1u);
          float cosv[CW];
          unstage_row<CW>((g & 1u) ? stg1 : stg0, row_t, sub, cosv);
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = w0 * cosv[j] * v[j];
          if (dma) ptx::bulk_wait_read_all();
          ptx::named_bar_sync(bar_id, QTHREADS);
          stage_row<CW>(stg2, row_t, sub, v);
          ptx::fence_proxy_async();
          ptx::named_bar_sync(bar_id, QTHREADS);
          if (dma) {
            ptx::tma_store_2d(&p.tmO0, q_stg + 2 * STG, cc * 64, row0 + q * 32);
            ptx::bulk_commit();
          }
          // column sums of the staged (bf16-rounded) slice while the TMA store drains it.
          // db: each thread owns 8 adjacent columns (one 16-byte vector per row) of RPG rows
          if (p.db) {
            constexpr int RPG = 32 / (QTHREADS / 8);          // rows per thread
            const int cg = tid_q & 7, rg = tid_q >> 3;
#pragma unroll
            for (int rr = 0; rr < RPG; ++rr) {
              const int r = q * 32 + rg * RPG + rr;
              uint32_t a0, a1, a2, a3;
              ptx::ld_shared_v4(stg2 + uint32_t(r) * 128u + ((uint32_t(cg) ^ uint32_t(r & 7)) << 4), a0, a1, a2, a3);
              cs[cc][0] += bf16_lo_f(a0); cs[cc][1] += bf16_hi_f(a0);
              cs[cc][2] += bf16_lo_f(a1); cs[cc][3] += bf16_hi_f(a1);
              cs[cc][4] += bf16_lo_f(a2); cs[cc][5] += bf16_hi_f(a2);
              cs[cc][6] += bf16_lo_f(a3); cs[cc][7] += bf16_hi_f(a3);
            }
          }
          // dW0 = zbar0^T x (layer 0 only): each thread owns 2 adjacent columns of 32/NSUB rows
          if (p.dW0) {
#pragma unroll 4
            for (int rr = 0; rr < 32 / NSUB; ++rr) {
              const int r = q * 32 + rgrp * (32 / NSUB) + rr;
              const uint32_t addr = stg2 + uint32_t(r) * 128u + ((uint32_t(cpair >> 2) ^ uint32_t(r & 7)) << 4) +
                                    uint32_t(cpair & 3) * 4u;
              const uint32_t u = ptx::ld_shared_u32(addr);
              const float z0 = bf16_lo_f(u), z1 = bf16_hi_f(u);
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float xi = sX[r * 3 + i];
                cw[cc][0][i] = fmaf(z0, xi, cw[cc][0][i]);
                cw[cc][1][i] = fmaf(z1, xi, cw[cc][1][i]);
              }
            }
          }
          ++g;
        }
      }

      if (MODE == 0 && p.fuse_last) {
        // combine the four column slices of each row (fixed order: deterministic) and write y
        if (sub != 0) {
          sY[(row_t * NSUB + sub) * 2 + 0] = ydot[0];
          sY[(row_t * NSUB + sub) * 2 + 1] = ydot[1];
        }
        ptx::named_bar_sync(bar_id, QTHREADS);
        if (sub == 0 && n_row < p.n) {
          const int wt = p.per_task ? task : 0;
          for (int oi = 0; oi < p.o; ++oi) {
            float tot = ydot[oi];
#pragma unroll
            for (int u = 1; u < NSUB; ++u) tot += sY[(row_t * NSUB + u) * 2 + oi];
            p.y[(size_t(task) * p.n + n_row) * p.o + oi] = tot + __ldg(p.bL + size_t(wt) * p.o + oi);
          }
        }
      }
    }
    if (dma) ptx::bulk_wait_all();
    if (MODE == 1 && (p.db || p.dW0) && acc_task >= 0) {
      // final flush: reduce the partials of all epilogue warps in shared memory (the staging tiles are
      // free now), then ONE atomic per output element and CTA.  Same-address atomics serialise in L2.
      constexpr int EPI_THREADS = 4 * QTHREADS;
      const int wt = p.per_task ? acc_task : 0;
      const int tid_e = threadIdx.x - kEpiWarp0 * 32;
      float* red = reinterpret_cast<float*>(sStg);            // [4*NSUB warps][256]  (<= 16 KB)
      ptx::named_bar_sync(15, EPI_THREADS);                   // every quadrant is done with the staging tiles
      if (p.db) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            float v = cs[cc][b];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 8) red[e * H + cc * 64 + lane * 8 + b] = v;
          }
        ptx::named_bar_sync(15, EPI_THREADS);
        for (int col = tid_e; col < H; col += EPI_THREADS) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < 4 * NSUB; ++w) v += red[w * H + col];
          red_add(p.db + size_t(wt) * H + col, v);
        }
        ptx::named_bar_sync(15, EPI_THREADS);
      }
      if (p.dW0) {
        for (int i = 0; i < p.d; ++i) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int b = 0; b < 2; ++b) red[e * H + cc * 64 + cpair * 2 + b] = (i == 0) ? cw[cc][b][0] : (i == 1) ? cw[cc][b][1] : cw[cc][b][2];
          ptx::named_bar_sync(15, EPI_THREADS);
          for (int col = tid_e; col < H; col += EPI_THREADS) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 4 * NSUB; ++w) v += red[w * H + col];
            red_add(p.dW0 + (size_t(wt) * H + col) * p.d + i, v);
          }
          ptx::named_bar_sync(15, EPI_THREADS);
        }
      }
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

// forward: 4 warps per quadrant hide the SFU latency of the sine/cosine epilogue best;
// backward: 2 warps per quadrant (measured: more warps only add contention on the staged tile)
constexpr int NSUB_FWD = 4;
constexpr int NSUB_BWD = 2;

cudaError_t launch_rows_fast(const RowsFastParams& p, int mode, int num_sms, cudaStream_t stream) {
  const int tiles_m = p.R / TILE_M;
  int G = num_sms < tiles_m ? num_sms : tiles_m;
  if (G < 1) G = 1;
  if (mode == 0) {
    auto kern = rows_fast_kernel<0, NSUB_FWD>;
    SIREN_ENSURE_SMEM(kern, SMEM_FAST);
    kern<<<G, 128 + NSUB_FWD * 128, SMEM_FAST, stream>>>(p);
  } else {
    auto kern = rows_fast_kernel<1, NSUB_BWD>;
    SIREN_ENSURE_SMEM(kern, SMEM_FAST);
    kern<<<G, 128 + NSUB_BWD * 128, SMEM_FAST, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace siren
