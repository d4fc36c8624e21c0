// Whole-MLP fused forward (bf16 mode, value stream, d_in <= 4, <= 4 hidden layers): a CTA carries a
// PAIR of 128-row tiles through every layer without the activations leaving the SM.
//
//   layer 0        sin(w0 (x W0^T + b0)) computed by the epilogue warps straight into the A-operand
//                  tiles in shared memory (K-major, 128-byte swizzle: the layout TMA would have written)
//   layers 1..NH   tcgen05.mma with A = those tiles, B = W_l streamed from L2 in [128 x 64] chunks
//                  (one chunk feeds both tiles of the pair), accumulators in TMEM (2 x 256 columns);
//                  the epilogue writes sin() back IN PLACE as the next layer's A operand
//   last layer     the outermost linear (d_out <= 2) is a dot product in the top layer's epilogue
//
// Epilogue work split: warp (q, sub) owns TMEM lanes / tile rows [32q, 32q+32) and the 64 columns of
// K-chunk `sub`, i.e. one contiguous 4 KB slice of the A tile -- no warp ever waits for another one
// inside a layer.  Columns are processed in pieces of 16 with the next TMEM load in flight.
//
// STASH = true (training): every layer's sine slice is TMA-stored from where it sits in the A tile and
// the cosine goes out through a warp-private double-buffered staging slot -- the stash the backward
// kernels expect (act[l], c[l]).  STASH = false (inference): nothing but y is written.
//
// The sine argument is formed as fma(acc, w0, w0*b) and handed to the SFU without the explicit
// one-revolution reduction of the per-layer kernels: the SFU's own 1/(2 pi) scaling keeps the absolute
// error below |arg| * 2^-23, two orders under the bf16 rounding of this precision mode.
//
// Reference semantics: modules.py:25-26 (BatchLinear), :38 (Sine), :92-97 (FCBlock chain).
#include "common.cuh"
#include "ptx.cuh"
#include "simt.h"

namespace siren {

namespace {

constexpr int MAX_FUSED_LAYERS = MAX_FUSED_HIDDEN_SMEM;   // hidden layers whose biases fit the shared-memory budget
constexpr int NSUB = 4;                         // epilogue warps per TMEM lane quadrant (= K chunks of a tile)
constexpr int kThreads = 128 + NSUB * 128;      // 4 control warps + 16 epilogue warps
constexpr int EPI_WARPS = 4 * NSUB;
constexpr int PW = 16;                          // columns per piece
constexpr int NPIECE = 64 / PW;
constexpr int A_TILE = 4 * TILE_M * 128;        // 64 KB: [4 k-chunks][128 rows][128 B]
constexpr int BN_CH = 128;                      // weight chunk: [128 out rows][64 k] = 16 KB
constexpr int B_STAGE = BN_CH * 128;
constexpr int NSTB = 3;
constexpr int C_SLOT = 32 * PW * 2;             // 1 KB: [32 rows][16 bf16], 32-byte swizzle
constexpr int C_STG = EPI_WARPS * 2 * C_SLOT;   // 32 KB: two slots per warp
constexpr int Y_BYTES = 2 * TILE_M * (NSUB - 1) * 2 * 4;   // partial last-layer dots [2 tiles][128][3][2]
constexpr int W0_BYTES = H * 4 * 4;             // (w0 * W0 | w0 * b0) as one float4 per column (d <= 3) ...
constexpr int B0_BYTES = H * 4;                 // ... and w0 * b0 separately for d == 4
constexpr int BIAS_BYTES = MAX_FUSED_LAYERS * H * 4;
constexpr int MISC = 1024;
constexpr int SMEM_FUSED = 2 * A_TILE + NSTB * B_STAGE + C_STG + Y_BYTES + W0_BYTES + B0_BYTES + BIAS_BYTES + MISC + 1024;
static_assert(SMEM_FUSED <= 232448, "shared memory budget");

struct PairInfo {
  int task, row0x, row0y;   // row0y < 0: the pair has a single tile
};
// pairs are formed inside a task so that both tiles share the weights
__device__ __forceinline__ PairInfo pair_info(const MlpFwdParams& p, int pair) {
  const int tiles_task = p.rows_per_task / TILE_M;
  const int pairs_task = (tiles_task + 1) / 2;
  PairInfo pi;
  pi.task = pair / pairs_task;
  const int lp = pair - pi.task * pairs_task;
  pi.row0x = pi.task * p.rows_per_task + (2 * lp) * TILE_M;
  pi.row0y = (2 * lp + 1 < tiles_task) ? pi.row0x + TILE_M : -1;
  return pi;
}

// Per-warp output state of the epilogue.
struct EpiOut {
  uint32_t a_row;      // shared address of this thread's 128-byte row inside its A slice (tile 0)
  uint32_t c_base;     // shared address of this warp's two cosine slots
  uint32_t c_row;      // byte offset of this thread's 32-byte row inside a slot
  int cur;             // cosine slot in use
  int row7, swz32;     // swizzle terms of this thread's row
  int lane;
};

// sincos of 16 arguments; sine -> A slice (bf16, in place), cosine -> staging slot + TMA store.
template <bool STASH>
__device__ __forceinline__ void piece_out(EpiOut& eo, int tl, int pc, const float* t, float* s, bool write_a,
                                          bool drain, const CUtensorMap* tmC, int gx, int gy) {
  float c[PW];
#pragma unroll
  for (int j = 0; j < PW; ++j) {
    s[j] = __sinf(t[j]);
    if (STASH) c[j] = __cosf(t[j]);
  }
  if (STASH) {
    if (eo.lane == 0) {
      // the slot written two pieces ago has been read.  In a single-tile pair the store that may still be
      // in flight is the sine slice this piece is about to overwrite: drain it too.
      if (drain) ptx::bulk_wait_read<0>();
      else ptx::bulk_wait_read<1>();
    }
    __syncwarp();
  }
  if (write_a) {
    const uint32_t arow = eo.a_row + uint32_t(tl) * A_TILE;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      ptx::st_shared_v4(arow + (uint32_t((2 * pc + h) ^ eo.row7) << 4), pack_bf16(s[8 * h], s[8 * h + 1]),
                        pack_bf16(s[8 * h + 2], s[8 * h + 3]), pack_bf16(s[8 * h + 4], s[8 * h + 5]),
                        pack_bf16(s[8 * h + 6], s[8 * h + 7]));
  }
  if (STASH) {
    const uint32_t slot = eo.c_base + uint32_t(eo.cur) * C_SLOT;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      ptx::st_shared_v4(slot + eo.c_row + (uint32_t(h ^ eo.swz32) << 4), pack_bf16(c[8 * h], c[8 * h + 1]),
                        pack_bf16(c[8 * h + 2], c[8 * h + 3]), pack_bf16(c[8 * h + 4], c[8 * h + 5]),
                        pack_bf16(c[8 * h + 6], c[8 * h + 7]));
    ptx::fence_proxy_async();
    __syncwarp();
    if (eo.lane == 0) {
      ptx::tma_store_2d(tmC, reinterpret_cast<const void*>(__cvta_shared_to_generic(slot)), gx, gy);
      ptx::bulk_commit();
    }
    eo.cur ^= 1;
  }
}

template <bool STASH>
__global__ void __launch_bounds__(kThreads, 1) mlp_fused_fwd_kernel(const __grid_constant__ MlpFwdParams p) {
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(TILE_M, BN_CH, 0, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                               // [2 tiles][4 chunks][128][128 B]
  uint8_t* sB = sA + 2 * A_TILE;                    // [NSTB][128][128 B]
  uint8_t* sC = sB + NSTB * B_STAGE;                // cosine staging
  float* sY = reinterpret_cast<float*>(sC + C_STG); // [2][128][NSUB-1][2]
  float4* sW0 = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(sY) + Y_BYTES);   // [256]
  float* sB0 = reinterpret_cast<float*>(sW0 + H);   // [256]
  float* sBias = sB0 + H;                           // [MAX_FUSED_LAYERS][256], times w0
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + MAX_FUSED_LAYERS * H);
  uint64_t* b_full = bars;                          // [NSTB]
  uint64_t* b_empty = bars + NSTB;                  // [NSTB]
  uint64_t* acc_full = bars + 2 * NSTB;             // [2]
  uint64_t* a_ready = bars + 2 * NSTB + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTB + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NH = p.n_hidden;
  // contiguous range of pairs for this CTA
  const int tiles_task = p.rows_per_task / TILE_M;
  const int n_pairs = ((tiles_task + 1) / 2) * p.tasks;
  const int base = n_pairs / gridDim.x, rem = n_pairs % gridDim.x;
  const int pr0 = blockIdx.x * base + (int(blockIdx.x) < rem ? int(blockIdx.x) : rem);
  const int pr1 = pr0 + base + (int(blockIdx.x) < rem ? 1 : 0);

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < NH; ++l) ptx::prefetch_tmap(&p.tmW[l]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTB; ++i) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&a_ready[i], EPI_WARPS);
    }
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
  const bool trace = p.dbg != nullptr && blockIdx.x == 0;
#define TRACE(pr_, l_, k_) do { if (trace && lane == 0 && (pr_) - pr0 < 3) p.dbg[(((pr_) - pr0) * 8 + (l_)) * 8 + (k_)] = clock64(); } while (0)
  if (trace && threadIdx.x == 0) p.dbg[0] = clock64();

  if (warp == 0) {
    // ===================== weight-chunk producer =====================
    if (lane == 0) {
      uint32_t seq = 0;
      for (int pr = pr0; pr < pr1; ++pr) {
        const PairInfo pi = pair_info(p, pr);
        const int wrow = (p.per_task ? pi.task : 0) * H;
        for (int l = 0; l < NH; ++l)
          for (int kc = 0; kc < 4; ++kc)
            for (int nh = 0; nh < 2; ++nh, ++seq) {
              const uint32_t st = seq % NSTB, ph = (seq / NSTB) & 1u;
              ptx::mbar_wait(&b_empty[st], ph ^ 1u);
              ptx::mbar_arrive_expect_tx(&b_full[st], B_STAGE);
              ptx::tma_load_2d(sB + st * B_STAGE, &p.tmW[l], &b_full[st], kc * KCHUNK, wrow + nh * BN_CH);
            }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t seq = 0;
    uint32_t rnd = 0u;               // bit tl: phase of a_ready[tl] ((pair, layer) rounds seen, mod 2)
    for (int pr = pr0; pr < pr1; ++pr) {
      const PairInfo pi = pair_info(p, pr);
      const int ntile = pi.row0y >= 0 ? 2 : 1;
      for (int l = 0; l < NH; ++l) {
        TRACE(pr, l + 1, 0);
        for (int kc = 0; kc < 4; ++kc)
          for (int nh = 0; nh < 2; ++nh, ++seq) {
            const uint32_t st = seq % NSTB, ph = (seq / NSTB) & 1u;
            ptx::mbar_wait(&b_full[st], ph);
            for (int tl = 0; tl < ntile; ++tl) {
              if (kc == 0 && nh == 0) {       // A tile written (and its accumulator drained)
                ptx::mbar_wait(&a_ready[tl], (rnd >> tl) & 1u);
                rnd ^= 1u << tl;
                TRACE(pr, l + 1, 1 + tl);
              }
              ptx::tc_fence_after();
              if (lane == 0) {
                const uint32_t a_addr = ptx::smem_u32(sA + tl * A_TILE + kc * (TILE_M * 128));
                const uint32_t b_addr = ptx::smem_u32(sB + st * B_STAGE);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  ptx::umma_bf16(tmem_base + uint32_t(tl * 256 + nh * BN_CH), ptx::umma_smem_desc(a_addr + ks * 32, 16, 1024),
                                 ptx::umma_smem_desc(b_addr + ks * 32, 16, 1024), IDESC, (kc | ks) ? 1u : 0u);
              }
              __syncwarp();
            }
            if (lane == 0) ptx::umma_commit(&b_empty[st]);
            __syncwarp();
          }
        if (lane == 0) {
          ptx::umma_commit(&acc_full[0]);
          if (ntile == 2) ptx::umma_commit(&acc_full[1]);
        }
        TRACE(pr, l + 1, 3);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue / layer-0 warps =====================
    const int e = warp - 4;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int sub = e >> 2;                 // K chunk (64 columns) this warp owns
    const int tid_e = threadIdx.x - 128;
    const int row_t = q * 32 + lane;
    const float w0 = p.w0;
    EpiOut eo;
    eo.a_row = ptx::smem_u32(sA) + uint32_t(sub) * (TILE_M * 128) + uint32_t(row_t) * 128u;
    eo.c_base = ptx::smem_u32(sC) + uint32_t(e) * (2 * C_SLOT);
    eo.c_row = uint32_t(lane) * 32u;
    eo.cur = 0;
    eo.row7 = row_t & 7;
    eo.swz32 = (lane >> 2) & 1;
    eo.lane = lane;
    uint32_t accph = 0u;                    // bit tl: phase of acc_full[tl]
    int cur_task = -1;
    const int colw = sub * 64;              // first column of this warp

    for (int pr = pr0; pr < pr1; ++pr) {
      const PairInfo pi = pair_info(p, pr);
      const int ntile = pi.row0y >= 0 ? 2 : 1;
      const int wt = p.per_task ? pi.task : 0;
      if (wt != cur_task) {                  // (re)load the first-layer weights and the biases of this task
        ptx::named_bar_sync(15, EPI_WARPS * 32);
        for (int col = tid_e; col < H; col += EPI_WARPS * 32) {
          const float* wr = p.W0 + (size_t(wt) * H + col) * p.d;
          const float bb = w0 * __ldg(p.b0 + size_t(wt) * H + col);
          float4 w;
          w.x = w0 * __ldg(wr);
          w.y = p.d > 1 ? w0 * __ldg(wr + 1) : 0.f;
          w.z = p.d > 2 ? w0 * __ldg(wr + 2) : 0.f;
          w.w = p.d > 3 ? w0 * __ldg(wr + 3) : bb;
          sW0[col] = w;
          sB0[col] = bb;
          for (int l = 0; l < NH; ++l) sBias[l * H + col] = w0 * __ldg(p.bias[l] + size_t(wt) * H + col);
        }
        ptx::named_bar_sync(15, EPI_WARPS * 32);
        cur_task = wt;
      }
      // ---------------- layer 0: straight into the A tiles ----------------
      for (int tl = 0; tl < ntile; ++tl) {
        const int row0 = tl ? pi.row0y : pi.row0x;
        const int n_row = row0 + row_t - pi.task * p.rows_per_task;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
        if (n_row < p.n) {
          const float* xp = p.x + (size_t(pi.task) * p.n + n_row) * p.d;
          x0 = __ldg(xp);
          if (p.d > 1) x1 = __ldg(xp + 1);
          if (p.d > 2) x2 = __ldg(xp + 2);
          if (p.d > 3) x3 = __ldg(xp + 3);
        }
        if (e == 0) TRACE(pr, 0, 4 + 2 * tl);
        const bool d4 = p.d > 3;
        if (!d4) x3 = 1.f;                   // .w of the packed column holds w0 * b0
#pragma unroll
        for (int pc = 0; pc < NPIECE; ++pc) {
          float t[PW], s[PW];
#pragma unroll
          for (int j = 0; j < PW; ++j) {
            const float4 w = sW0[colw + pc * PW + j];
            float z = x0 * w.x;
            z = fmaf(x1, w.y, z);
            z = fmaf(x2, w.z, z);
            z = fmaf(x3, w.w, z);
            if (d4) z += sB0[colw + pc * PW + j];
            t[j] = z;
          }
          piece_out<STASH>(eo, tl, pc, t, s, true, ntile == 1 && pc == 0, &p.tmCos[0], colw + pc * PW, row0 + q * 32);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (STASH) {
            ptx::tma_store_2d(&p.tmAct[0], sA + tl * A_TILE + sub * (TILE_M * 128) + q * (32 * 128), colw, row0 + q * 32);
            ptx::bulk_commit();
          }
          ptx::mbar_arrive(&a_ready[tl]);
        }
        if (e == 0) TRACE(pr, 0, 5 + 2 * tl);
      }
      // ---------------- hidden layers ----------------
      for (int l = 1; l <= NH; ++l) {
        const bool top = (l == NH);
        const float4* bias4 = reinterpret_cast<const float4*>(sBias + (l - 1) * H + colw);
        for (int tl = 0; tl < ntile; ++tl) {
          const int row0 = tl ? pi.row0y : pi.row0x;
          const int n_row = row0 + row_t - pi.task * p.rows_per_task;
          const bool write_a = STASH || !top;      // the sine slice is the next layer's operand and/or the stash
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(tl * 256 + colw);
          float ydot0 = 0.f, ydot1 = 0.f;
          float va[PW], vb[PW];
          ptx::mbar_wait(&acc_full[tl], (accph >> tl) & 1u);
          accph ^= 1u << tl;
          ptx::tc_fence_after();
          if (e == 0) TRACE(pr, l, 4 + 2 * tl);
          ptx::tmem_ld<PW>(taddr, reinterpret_cast<uint32_t*>(va));
#pragma unroll
          for (int pc = 0; pc < NPIECE; ++pc) {
            float* v = (pc & 1) ? vb : va;
            ptx::tmem_wait_ld();
            if (pc + 1 < NPIECE)
              ptx::tmem_ld<PW>(taddr + uint32_t((pc + 1) * PW), reinterpret_cast<uint32_t*>((pc & 1) ? va : vb));
            float t[PW], s[PW];
#pragma unroll
            for (int j4 = 0; j4 < PW / 4; ++j4) {
              const float4 bb = bias4[pc * (PW / 4) + j4];
              t[4 * j4 + 0] = fmaf(v[4 * j4 + 0], w0, bb.x);
              t[4 * j4 + 1] = fmaf(v[4 * j4 + 1], w0, bb.y);
              t[4 * j4 + 2] = fmaf(v[4 * j4 + 2], w0, bb.z);
              t[4 * j4 + 3] = fmaf(v[4 * j4 + 3], w0, bb.w);
            }
            piece_out<STASH>(eo, tl, pc, t, s, write_a, ntile == 1 && pc == 0, &p.tmCos[l], colw + pc * PW, row0 + q * 32);
            if (top && p.fuse_last) {
              const float4* wl0 = reinterpret_cast<const float4*>(p.WL + (size_t(wt) * p.o) * H + colw + pc * PW);
#pragma unroll
              for (int j4 = 0; j4 < PW / 4; ++j4) {
                const float4 ww = __ldg(wl0 + j4);
                ydot0 = fmaf(s[4 * j4 + 0], ww.x, ydot0); ydot0 = fmaf(s[4 * j4 + 1], ww.y, ydot0);
                ydot0 = fmaf(s[4 * j4 + 2], ww.z, ydot0); ydot0 = fmaf(s[4 * j4 + 3], ww.w, ydot0);
              }
              if (p.o > 1) {
                const float4* wl1 = wl0 + H / 4;
#pragma unroll
                for (int j4 = 0; j4 < PW / 4; ++j4) {
                  const float4 ww = __ldg(wl1 + j4);
                  ydot1 = fmaf(s[4 * j4 + 0], ww.x, ydot1); ydot1 = fmaf(s[4 * j4 + 1], ww.y, ydot1);
                  ydot1 = fmaf(s[4 * j4 + 2], ww.z, ydot1); ydot1 = fmaf(s[4 * j4 + 3], ww.w, ydot1);
                }
              }
            }
          }
          ptx::tc_fence_before();
          if (write_a) {
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (STASH) {
                ptx::tma_store_2d(&p.tmAct[l], sA + tl * A_TILE + sub * (TILE_M * 128) + q * (32 * 128), colw, row0 + q * 32);
                ptx::bulk_commit();
              }
              if (!top) ptx::mbar_arrive(&a_ready[tl]);
            }
          }
          if (e == 0) TRACE(pr, l, 5 + 2 * tl);
          if (top && p.fuse_last) {
            float* sy = sY + tl * (TILE_M * (NSUB - 1) * 2);
            if (sub != 0) {
              sy[(row_t * (NSUB - 1) + sub - 1) * 2 + 0] = ydot0;
              sy[(row_t * (NSUB - 1) + sub - 1) * 2 + 1] = ydot1;
            }
            ptx::named_bar_sync(1 + q, NSUB * 32);
            if (sub == 0 && n_row < p.n) {
#pragma unroll
              for (int u = 0; u < NSUB - 1; ++u) {
                ydot0 += sy[(row_t * (NSUB - 1) + u) * 2 + 0];
                ydot1 += sy[(row_t * (NSUB - 1) + u) * 2 + 1];
              }
              float* yp = p.y + (size_t(pi.task) * p.n + n_row) * p.o;
              yp[0] = ydot0 + __ldg(p.bL + size_t(wt) * p.o);
              if (p.o > 1) yp[1] = ydot1 + __ldg(p.bL + size_t(wt) * p.o + 1);
            }
          }
        }
      }
    }
    if (lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

cudaError_t launch_mlp_fused_fwd(const MlpFwdParams& p, bool stash, int num_sms, cudaStream_t stream) {
  static bool set0 = false, set1 = false;
  const int tiles_task = p.rows_per_task / TILE_M;
  const int n_pairs = ((tiles_task + 1) / 2) * p.tasks;
  int G = num_sms < n_pairs ? num_sms : n_pairs;
  if (G < 1) G = 1;
  if (stash) {
    if (!set1) {
      cudaError_t e = cudaFuncSetAttribute(mlp_fused_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FUSED);
      if (e != cudaSuccess) return e;
      set1 = true;
    }
    mlp_fused_fwd_kernel<true><<<G, kThreads, SMEM_FUSED, stream>>>(p);
  } else {
    if (!set0) {
      cudaError_t e = cudaFuncSetAttribute(mlp_fused_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FUSED);
      if (e != cudaSuccess) return e;
      set0 = true;
    }
    mlp_fused_fwd_kernel<false><<<G, kThreads, SMEM_FUSED, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace siren
