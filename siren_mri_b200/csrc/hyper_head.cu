// Hypernetwork head in the consumer's layout.
//
// Replaces, for one HIDDEN weight matrix of the hypo-network, the last linear of its HyperNetwork head
// (meta_modules.py:32-35, 50-54: FCBlock(..., outermost_linear=True, 'relu').net[-1] = BatchLinear(k_h -> 256 * 256),
// reshaped to [B, 256, 256]) AND the per-call conversion of that fp32 tensor into the tensor-core operands
// (simt.cu prep_weights_kernel) AND the sum of squares loss_functions.hypo_weight_loss (loss_functions.py:279-287) takes:
//
//   W[b][o][i] = sum_k h[b][k] Wlast[o * 256 + i][k] + blast[o * 256 + i]                  fp32, the hypo_params entry
//   wk16[b][o][i] = fp16(W[b][o][i])                    operand of the fused forward   (mlp_fused_pair.cu)
//   wt16[b][i][o] = bf16(w0 * W[b][o][i])               operand of the dgrad chain     (mlp_fused_bwd.cu)
//
// The contraction has M = B tasks (8 per GPU in the sharded MRI configuration), N = 65,536, K = k_h <= 512: a skinny
// product bound by the one pass over Wlast (67 MB at k_h = 256), so it runs on the CUDA cores -- a warp reads a row of
// Wlast coalesced and keeps BT task accumulators per lane; the tensor cores would idle behind the same stream.
// A block owns a 16 x 16 tile of the hypo weight (256 blocks per pass) so that all three layouts leave as full 32-byte
// sectors (the transposed copy through the shared-memory tile).  More than BT tasks take ceil(B / BT) passes over
// Wlast (grid.z).
#include <cuda_fp16.h>

#include "common.cuh"
#include "simt.h"

namespace siren {

namespace {

constexpr int TS = 16;      // tile side

template <int BT>
__global__ void __launch_bounds__(256) hyper_head_kernel(HyperHeadParams p) {
  extern __shared__ __align__(16) float hh_smem[];
  float* sh = hh_smem;                                   // [BT][k_h]
  float* tile = hh_smem + BT * p.k_h;                    // [BT][TS][TS + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * TS, o0 = blockIdx.y * TS, b0 = blockIdx.z * BT;
  const int K = p.k_h;
  for (int idx = threadIdx.x; idx < BT * K; idx += 256) {
    const int b = idx / K, k = idx - b * K;
    sh[idx] = (b0 + b < p.tasks) ? p.h[size_t(b0 + b) * K + k] : 0.f;
  }
  __syncthreads();

  // which task's total this lane ends up with after the halving butterfly below
  const int my_b = (BT == 8) ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))
                             : (((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1));
  const bool writer = (BT == 8) ? ((lane & 3) == 0) : ((lane & 1) == 0);

  for (int seg = 0; seg < TS / 8; ++seg) {
    const int r = warp * (TS / 8) + seg;                 // out row inside the tile
    const float* wrow0 = p.Wlast + (size_t(o0 + r) * H + i0) * K;
#pragma unroll 4
    for (int c = 0; c < TS; ++c) {
      const float* wrow = wrow0 + size_t(c) * K;
      float acc[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[b] = 0.f;
      for (int k = lane * 4; k < K; k += 128) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float4 hv = *reinterpret_cast<const float4*>(sh + b * K + k);
          acc[b] = fmaf(w.x, hv.x, acc[b]);
          acc[b] = fmaf(w.y, hv.y, acc[b]);
          acc[b] = fmaf(w.z, hv.z, acc[b]);
          acc[b] = fmaf(w.w, hv.w, acc[b]);
        }
      }
      // halving butterfly: each round sends the half of the partials the partner keeps (BT -> 1 values per lane),
      // then the remaining lanes are summed: BT - 1 + log2(32 / BT) shuffles instead of 5 BT
#pragma unroll
      for (int half = BT / 2, m = 16; half >= 1; half >>= 1, m >>= 1) {
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float send = upper ? acc[i] : acc[i + half];
          const float recv = __shfl_xor_sync(0xffffffffu, send, m);
          acc[i] = (upper ? acc[i + half] : acc[i]) + recv;
        }
      }
      float v = acc[0];
      if (BT == 8) {
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
      } else {
        v += __shfl_xor_sync(0xffffffffu, v, 1);
      }
      if (writer) tile[(my_b * TS + r) * (TS + 1) + c] = v + __ldg(p.blast + size_t(o0 + r) * H + i0 + c);
    }
  }
  __syncthreads();

  // the three layouts, two 16-wide row segments per warp instruction
  float ss = 0.f;
  const int cl = lane & (TS - 1);
  for (int b = 0; b < BT && b0 + b < p.tasks; ++b) {
    const size_t base = size_t(b0 + b) * H * H;
    for (int r = warp * 2 + (lane >> 4); r < TS; r += 16) {
      const float v = tile[(b * TS + r) * (TS + 1) + cl];                // W[b][o0 + r][i0 + cl]
      p.W_out[base + size_t(o0 + r) * H + i0 + cl] = v;
      if (p.wk16) p.wk16[base + size_t(o0 + r) * H + i0 + cl] = __float2half_rn(v);
      ss = fmaf(v, v, ss);
      if (p.wt16) {
        const float t = tile[(b * TS + cl) * (TS + 1) + r] * p.w0;       // W[b][o0 + cl][i0 + r]  ->  wt16[b][i0 + r][o0 + cl]
        p.wt16[base + size_t(i0 + r) * H + o0 + cl] = __float2bfloat16_rn(t);
      }
    }
  }
  if (p.sumsq) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    __shared__ float part[8];
    if (lane == 0) part[warp] = ss;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(p.sumsq, part[0] + part[1] + part[2] + part[3] + part[4] + part[5] + part[6] + part[7]);
  }
}

}  // namespace

cudaError_t launch_hyper_head(const HyperHeadParams& p, cudaStream_t stream) {
  if (p.tasks > 8) {
    constexpr int BT = 16;
    const size_t smem = size_t(BT) * (p.k_h + TS * (TS + 1)) * sizeof(float);
    SIREN_ENSURE_SMEM(hyper_head_kernel<BT>, int(size_t(BT) * (512 + TS * (TS + 1)) * sizeof(float)));
    hyper_head_kernel<BT><<<dim3(H / TS, H / TS, (p.tasks + BT - 1) / BT), 256, smem, stream>>>(p);
  } else {
    constexpr int BT = 8;
    const size_t smem = size_t(BT) * (p.k_h + TS * (TS + 1)) * sizeof(float);
    SIREN_ENSURE_SMEM(hyper_head_kernel<BT>, int(size_t(BT) * (512 + TS * (TS + 1)) * sizeof(float)));
    hyper_head_kernel<BT><<<dim3(H / TS, H / TS, 1), 256, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace siren
