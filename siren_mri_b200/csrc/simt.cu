// CUDA-core kernels around the tensor-core layers: weight conversion, the skinny first and last
// layers (K = d_in and N = d_out are far too small for an MMA), bias gradients, fused
// clip + Adam and the MSE loss gradient used by the fast training step.
#include "common.cuh"
#include "simt.h"

namespace siren {

namespace {

constexpr int MAXD = 16;   // max input features served natively
constexpr int MAXO = 8;    // max output features served natively

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// -------------------------------------------------------------------------------------------
// fp32 weights [T][256][256] -> bf16 (hi, lo) in both orientations
// -------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ W, bf16* __restrict__ k_hi, bf16* __restrict__ k_lo,
                                    bf16* __restrict__ t_hi, bf16* __restrict__ t_lo, int split) {
  __shared__ float tile[32][33];
  const size_t base = size_t(blockIdx.z) * H * H;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const float v = W[base + size_t(by + i) * H + bx + threadIdx.x];
    tile[i][threadIdx.x] = v;
    const bf16 h = __float2bfloat16_rn(v);
    k_hi[base + size_t(by + i) * H + bx + threadIdx.x] = h;
    if (split) k_lo[base + size_t(by + i) * H + bx + threadIdx.x] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const float v = tile[threadIdx.x][i];
    const bf16 h = __float2bfloat16_rn(v);
    t_hi[base + size_t(bx + i) * H + by + threadIdx.x] = h;
    if (split) t_lo[base + size_t(bx + i) * H + by + threadIdx.x] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

template <bool SPLIT>
__device__ __forceinline__ void put_operand(bf16* hi, bf16* lo, size_t off, float v) {
  const bf16 h = __float2bfloat16_rn(v);
  hi[off] = h;
  if (SPLIT) lo[off] = __float2bfloat16_rn(v - __bfloat162float(h));
}
template <bool SPLIT>
__device__ __forceinline__ float get_operand(const bf16* hi, const bf16* lo, size_t off) {
  float v = __bfloat162float(hi[off]);
  if (SPLIT) v += __bfloat162float(lo[off]);
  return v;
}
template <bool F32>
__device__ __forceinline__ void put_stash(void* base, size_t off, float v) {
  if (F32) reinterpret_cast<float*>(base)[off] = v;
  else reinterpret_cast<bf16*>(base)[off] = __float2bfloat16_rn(v);
}
template <bool F32>
__device__ __forceinline__ float get_stash(const void* base, size_t off) {
  if (F32) return reinterpret_cast<const float*>(base)[off];
  return __bfloat162float(reinterpret_cast<const bf16*>(base)[off]);
}

// -------------------------------------------------------------------------------------------
// first layer forward: z0 = x W0^T + b0, h0 = sin(w0 z0), c0 = cos(w0 z0); input jets are unit
// vectors, so Jz_k = W0[:, k] and Dz_k = 0.   One thread per feature column, rows looped.
// -------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) first_fwd_kernel(FirstParams p) {
  const int col = threadIdx.x;
  const int rows_per_block = p.rows_per_block;
  const int row_begin = blockIdx.x * rows_per_block;
  const size_t plane = size_t(p.R) * H;
  const float w0 = p.w0, w0_rev = p.w0 * 0.15915494309189535f;
  int cur_task = -1;
  float w[MAXD];
  float b = 0.f;
  for (int r = row_begin; r < row_begin + rows_per_block && r < p.R; ++r) {
    const int task = r / p.n_pad, n = r - task * p.n_pad;
    const int wt = p.per_task ? task : 0;
    if (wt != cur_task) {
#pragma unroll
      for (int i = 0; i < MAXD; ++i) w[i] = (i < p.d) ? p.W[(size_t(wt) * H + col) * p.d + i] : 0.f;
      b = p.b[size_t(wt) * H + col];
      cur_task = wt;
    }
    float z = b;
    if (n < p.n) {
      const float* x = p.x + (size_t(task) * p.n + n) * p.d;
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (i < p.d) z = fmaf(__ldg(x + i), w[i], z);
    }
    float s, c;
    sincos_w0<SPLIT>(z, w0, w0_rev, &s, &c);
    const size_t off = size_t(r) * H + col;
    put_operand<SPLIT>(p.act_hi, p.act_lo, off, s);
    put_stash<SPLIT>(p.c, off, c);
    if (p.order >= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < p.d) {
          put_operand<SPLIT>(p.act_hi, p.act_lo, size_t(1 + k) * plane + off, w0 * c * w[k]);
          if (p.order == 2)
            put_operand<SPLIT>(p.act_hi, p.act_lo, size_t(1 + p.d + k) * plane + off, -(w0 * w0) * s * w[k] * w[k]);
        }
    }
  }
}

// -------------------------------------------------------------------------------------------
// last layer forward: out[s][n, o] = plane_s[n, :] . W_L[o, :] (+ b_L for the value stream)
// One warp per row.
// -------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) last_fwd_kernel(LastParams p) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const size_t plane = size_t(p.R) * H;
  const int S = 1 + p.order * p.d;
  for (int r = warp_global; r < p.R; r += nwarps) {
    const int task = r / p.n_pad, n = r - task * p.n_pad;
    if (n >= p.n) continue;
    const float* W = p.W + size_t(p.per_task ? task : 0) * p.o * H;
    const size_t orow = size_t(task) * p.n + n;
    for (int s = 0; s < S; ++s) {
      float h[8];
      load_operand_chunk<8, SPLIT>(p.act_hi, p.act_lo, size_t(s) * plane + size_t(r) * H + lane * 8, h);
      for (int oi = 0; oi < p.o; ++oi) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(W + oi * H + lane * 8));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(W + oi * H + lane * 8 + 4));
        float acc = h[0] * wa.x + h[1] * wa.y + h[2] * wa.z + h[3] * wa.w + h[4] * wb.x + h[5] * wb.y +
                    h[6] * wb.z + h[7] * wb.w;
        acc = warp_sum(acc);
        if (lane == 0) {
          if (s == 0) {
            p.y[orow * p.o + oi] = acc + p.b[size_t(p.per_task ? task : 0) * p.o + oi];
          } else if (s <= p.d) {
            p.J[(orow * p.o + oi) * p.d + (s - 1)] = acc;
          } else {
            p.Dd[(orow * p.o + oi) * p.d + (s - 1 - p.d)] = acc;
          }
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// last layer backward: adjoints of the top sine layer from (gy, gJ, gD), its sine reverse, and
// dW_L / db_L.  One thread per feature column, rows looped.
// -------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) last_bwd_kernel(LastParams p) {
  const int col = threadIdx.x;
  const int row_begin = blockIdx.x * p.rows_per_block;
  const size_t plane = size_t(p.R) * H;
  const float w0 = p.w0;
  const int d = p.d, o = p.o, order = p.order;
  float w[MAXO], dw[MAXO];
  float dbias = 0.f;   // only thread `col < o` uses it: db_L[col]
  int cur_task = -1;
#pragma unroll
  for (int i = 0; i < MAXO; ++i) dw[i] = 0.f;

  auto flush = [&](int wt) {
#pragma unroll
    for (int i = 0; i < MAXO; ++i)
      if (i < o) {
        if (dw[i] != 0.f) atomicAdd(p.dW + (size_t(wt) * o + i) * H + col, dw[i]);
        dw[i] = 0.f;
      }
    if (col < o && dbias != 0.f) atomicAdd(p.db + size_t(wt) * o + col, dbias);
    dbias = 0.f;
  };

  for (int r = row_begin; r < row_begin + p.rows_per_block && r < p.R; ++r) {
    const int task = r / p.n_pad, n = r - task * p.n_pad;
    const int wt = p.per_task ? task : 0;
    if (wt != cur_task) {
      if (cur_task >= 0) flush(cur_task);
#pragma unroll
      for (int i = 0; i < MAXO; ++i) w[i] = (i < o) ? p.W[(size_t(wt) * o + i) * H + col] : 0.f;
      cur_task = wt;
    }
    const size_t off = size_t(r) * H + col;
    const bool valid = n < p.n;
    const size_t orow = size_t(task) * p.n + n;
    float ab = 0.f;
    if (valid) {
      const float s = get_operand<SPLIT>(p.act_hi, p.act_lo, off);
#pragma unroll
      for (int i = 0; i < MAXO; ++i)
        if (i < o) {
          const float g = __ldg(p.gy + orow * o + i);
          ab = fmaf(g, w[i], ab);
          dw[i] = fmaf(g, s, dw[i]);
          if (col == i) dbias += g;
        }
      const float c = get_stash<SPLIT>(p.c, off);
      float zb = w0 * c * ab;
      if (order >= 1) {
        for (int k = 0; k < d; ++k) {
          float jb = 0.f, db = 0.f;
          const float jact = get_operand<SPLIT>(p.act_hi, p.act_lo, size_t(1 + k) * plane + off);
          float dact = 0.f;
          if (order == 2) dact = get_operand<SPLIT>(p.act_hi, p.act_lo, size_t(1 + d + k) * plane + off);
#pragma unroll
          for (int i = 0; i < MAXO; ++i)
            if (i < o) {
              const float gj = p.gJ ? __ldg(p.gJ + (orow * o + i) * d + k) : 0.f;
              jb = fmaf(gj, w[i], jb);
              dw[i] = fmaf(gj, jact, dw[i]);
              if (order == 2) {
                const float gd = p.gD ? __ldg(p.gD + (orow * o + i) * d + k) : 0.f;
                db = fmaf(gd, w[i], db);
                dw[i] = fmaf(gd, dact, dw[i]);
              }
            }
          const float jz = p.top_is_first ? p.w_first[(size_t(wt) * H + col) * d + k]
                                          : get_stash<SPLIT>(p.jz, size_t(k) * plane + off);
          zb -= (w0 * w0) * s * jz * jb;
          float jzb = w0 * c * jb;
          if (order == 2) {
            const float dz = p.top_is_first ? 0.f : get_stash<SPLIT>(p.jz, size_t(d + k) * plane + off);
            zb -= (w0 * w0) * s * dz * db + (w0 * w0 * w0) * c * jz * jz * db;
            jzb -= 2.f * (w0 * w0) * s * jz * db;
            put_operand<SPLIT>(p.adj_hi, p.adj_lo, size_t(1 + d + k) * plane + off, w0 * c * db);
          }
          put_operand<SPLIT>(p.adj_hi, p.adj_lo, size_t(1 + k) * plane + off, jzb);
        }
      }
      put_operand<SPLIT>(p.adj_hi, p.adj_lo, off, zb);
    } else {
      const int S = 1 + order * d;
      for (int s = 0; s < S; ++s) put_operand<SPLIT>(p.adj_hi, p.adj_lo, size_t(s) * plane + off, 0.f);
    }
  }
  if (cur_task >= 0) flush(cur_task);
}

// -------------------------------------------------------------------------------------------
// first layer backward: dW0[col, i] = sum_n zbar0[n, col] x[n, i] (+ sum_n Jzbar_i[n, col]),
// db0[col] = sum_n zbar0[n, col].
// -------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) first_bwd_kernel(FirstParams p) {
  const int col = threadIdx.x;
  const int row_begin = blockIdx.x * p.rows_per_block;
  const size_t plane = size_t(p.R) * H;
  float dw[MAXD];
  float db = 0.f;
#pragma unroll
  for (int i = 0; i < MAXD; ++i) dw[i] = 0.f;
  int cur_task = -1;
  auto flush = [&](int wt) {
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < p.d) {
        atomicAdd(p.dW + (size_t(wt) * H + col) * p.d + i, dw[i]);
        dw[i] = 0.f;
      }
    atomicAdd(p.db + size_t(wt) * H + col, db);
    db = 0.f;
  };
  for (int r = row_begin; r < row_begin + p.rows_per_block && r < p.R; ++r) {
    const int task = r / p.n_pad, n = r - task * p.n_pad;
    if (n >= p.n) continue;
    const int wt = p.per_task ? task : 0;
    if (wt != cur_task) {
      if (cur_task >= 0) flush(cur_task);
      cur_task = wt;
    }
    const size_t off = size_t(r) * H + col;
    const float zb = get_operand<SPLIT>(p.adj_hi, p.adj_lo, off);
    const float* x = p.x + (size_t(task) * p.n + n) * p.d;
    db += zb;
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < p.d) dw[i] = fmaf(zb, __ldg(x + i), dw[i]);
    if (p.order >= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < p.d) dw[k] += get_operand<SPLIT>(p.adj_hi, p.adj_lo, size_t(1 + k) * plane + off);
    }
  }
  if (cur_task >= 0) flush(cur_task);
}

// gradient reaching the coordinates through z0:  gx[n, i] = sum_col zbar0[n, col] W0[col, i]
template <bool SPLIT>
__global__ void __launch_bounds__(256) coords_grad_kernel(FirstParams p) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp_global; r < p.R; r += nwarps) {
    const int task = r / p.n_pad, n = r - task * p.n_pad;
    if (n >= p.n) continue;
    const float* W = p.W + size_t(p.per_task ? task : 0) * H * p.d;
    float zb[8];
    load_operand_chunk<8, SPLIT>(p.adj_hi, p.adj_lo, size_t(r) * H + lane * 8, zb);
    for (int i = 0; i < p.d; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(zb[j], __ldg(W + size_t(lane * 8 + j) * p.d + i), acc);
      acc = warp_sum(acc);
      if (lane == 0) p.gx[(size_t(task) * p.n + n) * p.d + i] = acc;
    }
  }
}

// -------------------------------------------------------------------------------------------
// bias gradients of the hidden layers: column sums of adjoint plane 0
// -------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo,
                                                     float* __restrict__ db, int R, int n_pad, int per_task,
                                                     int rows_per_block) {
  const int col = threadIdx.x;
  const int row_begin = blockIdx.x * rows_per_block;
  float acc = 0.f;
  int cur_task = -1;
  for (int r = row_begin; r < row_begin + rows_per_block && r < R; ++r) {
    const int wt = per_task ? r / n_pad : 0;
    if (wt != cur_task) {
      if (cur_task >= 0) atomicAdd(db + size_t(cur_task) * H + col, acc);
      acc = 0.f;
      cur_task = wt;
    }
    acc += get_operand<SPLIT>(hi, lo, size_t(r) * H + col);
  }
  if (cur_task >= 0) atomicAdd(db + size_t(cur_task) * H + col, acc);
}

// -------------------------------------------------------------------------------------------
// optimizer tail: global grad norm (for clip_grad_norm_) and Adam, over one flat buffer
// -------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// one thread: advance the step counter and refresh the bias corrections (in double, like the
// Python-side arithmetic of torch.optim.Adam).  Keeping the counter on the device makes the
// whole optimizer tail CUDA-graph capturable.
__global__ void adam_tick_kernel(AdamState* st, double b1, double b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long step = st->step + 1;
    st->step = step;
    st->bc1 = (float)(1.0 - pow(b1, (double)step));
    st->bc2_sqrt = (float)sqrt(1.0 - pow(b2, (double)step));
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                            float max_norm, float grad_scale, const AdamState* __restrict__ st) {
  float scale = grad_scale;
  if (max_norm > 0.f) {
    const float total = sqrtf(st->sumsq) * grad_scale;
    const float coef = max_norm / (total + 1e-6f);
    scale *= fminf(coef, 1.0f);
  }
  const float step_size = lr / st->bc1;
  const float bc2_sqrt = st->bc2_sqrt;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// gy = 2 w (y - gt); loss += w sum (y - gt)^2
__global__ void mse_grad_kernel(const float* __restrict__ y, const float* __restrict__ gt, float* __restrict__ gy,
                                long n, float weight, float* __restrict__ loss) {
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float dlt = y[i] - gt[i];
    gy[i] = 2.f * weight * dlt;
    acc = fmaf(dlt, dlt, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, v * weight);
  }
}

// fp32 rows [R,256] -> bf16 (hi, lo) planes (debug entry points)
__global__ void to_planes_kernel(const float* __restrict__ src, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                 long n, int split) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = src[i];
    const bf16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (split) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

int rows_per_block_for(int R, int num_sms) {
  // a few blocks per SM, at least 16 rows each
  int target_blocks = num_sms * 8;
  int rpb = (R + target_blocks - 1) / target_blocks;
  if (rpb < 16) rpb = 16;
  return rpb;
}

}  // namespace

cudaError_t launch_prep_weights(const float* W, bf16* k_hi, bf16* k_lo, bf16* t_hi, bf16* t_lo, int tasks,
                                bool split, cudaStream_t stream) {
  dim3 grid(H / 32, H / 32, tasks), block(32, 8);
  prep_weights_kernel<<<grid, block, 0, stream>>>(W, k_hi, k_lo, t_hi, t_lo, split ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_first_fwd(FirstParams p, bool split, int num_sms, cudaStream_t stream) {
  p.rows_per_block = rows_per_block_for(p.R, num_sms);
  const int grid = (p.R + p.rows_per_block - 1) / p.rows_per_block;
  if (split) first_fwd_kernel<true><<<grid, 256, 0, stream>>>(p);
  else first_fwd_kernel<false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_first_bwd(FirstParams p, bool split, int num_sms, cudaStream_t stream) {
  p.rows_per_block = rows_per_block_for(p.R, num_sms);
  const int grid = (p.R + p.rows_per_block - 1) / p.rows_per_block;
  if (split) first_bwd_kernel<true><<<grid, 256, 0, stream>>>(p);
  else first_bwd_kernel<false><<<grid, 256, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (p.gx) {
    const int g2 = num_sms * 8;
    if (split) coords_grad_kernel<true><<<g2, 256, 0, stream>>>(p);
    else coords_grad_kernel<false><<<g2, 256, 0, stream>>>(p);
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_last_fwd(LastParams p, bool split, int num_sms, cudaStream_t stream) {
  const int grid = num_sms * 8;
  if (split) last_fwd_kernel<true><<<grid, 256, 0, stream>>>(p);
  else last_fwd_kernel<false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_last_bwd(LastParams p, bool split, int num_sms, cudaStream_t stream) {
  p.rows_per_block = rows_per_block_for(p.R, num_sms);
  const int grid = (p.R + p.rows_per_block - 1) / p.rows_per_block;
  if (split) last_bwd_kernel<true><<<grid, 256, 0, stream>>>(p);
  else last_bwd_kernel<false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_colsum(const bf16* hi, const bf16* lo, float* db, int R, int n_pad, int per_task, bool split,
                          int num_sms, cudaStream_t stream) {
  const int rpb = rows_per_block_for(R, num_sms);
  const int grid = (R + rpb - 1) / rpb;
  if (split) colsum_kernel<true><<<grid, 256, 0, stream>>>(hi, lo, db, R, n_pad, per_task, rpb);
  else colsum_kernel<false><<<grid, 256, 0, stream>>>(hi, lo, db, R, n_pad, per_task, rpb);
  return cudaGetLastError();
}

cudaError_t launch_sumsq(const float* g, long n, float* out, int num_sms, cudaStream_t stream) {
  long blocks = (n + 1023) / 1024;
  if (blocks > num_sms * 4) blocks = num_sms * 4;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, stream>>>(g, n, out);
  return cudaGetLastError();
}

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long n, float lr, double b1, double b2,
                        float eps, float max_norm, float grad_scale, AdamState* st, int num_sms,
                        cudaStream_t stream) {
  adam_tick_kernel<<<1, 32, 0, stream>>>(st, b1, b2);
  long blocks = (n + 255) / 256;
  if (blocks > num_sms * 8) blocks = num_sms * 8;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, stream>>>(p, g, m, v, n, lr, (float)b1, (float)b2, eps, max_norm,
                                                grad_scale, st);
  return cudaGetLastError();
}

cudaError_t launch_mse_grad(const float* y, const float* gt, float* gy, long n, float weight, float* loss,
                            int num_sms, cudaStream_t stream) {
  long blocks = (n + 1023) / 1024;
  if (blocks > num_sms * 4) blocks = num_sms * 4;
  if (blocks < 1) blocks = 1;
  mse_grad_kernel<<<(int)blocks, 256, 0, stream>>>(y, gt, gy, n, weight, loss);
  return cudaGetLastError();
}

cudaError_t launch_to_planes(const float* src, bf16* hi, bf16* lo, long n, bool split, cudaStream_t stream) {
  to_planes_kernel<<<1024, 256, 0, stream>>>(src, hi, lo, n, split ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace siren
