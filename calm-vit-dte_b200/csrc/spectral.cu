// Batched spectral normalisation for every sn(...) layer of the model in ONE launch per direction.
// Replaces the per-layer forward-pre-hook of torch.nn.utils.spectral_norm (torch/nn/utils/spectral_norm.py:92-114):
//     v <- normalize(W^T u) ; u <- normalize(W v) ; sigma = u . (W v) ; W_eff = W / sigma
// (same update order, eps and in-place buffer semantics), and writes the bf16 operand copies the tcgen05 GEMMs read:
// W_eff (rows, cols) for forward/wgrad and its transpose (cols, rows) for dgrad. An optional LayerScale vector is folded
// into the rows of W_eff (out_proj*ls_att, mlp.3*ls_mlp: Vi_Tools_CNN_less_V2.py:300,314).
// HBM-bound: each fp32 W is read ~3x (2nd/3rd read from L2), one CTA per layer, deterministic (no atomics).
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

constexpr int SN_THREADS = 512;
constexpr int SN_WARPS = SN_THREADS / 32;

// t[c] = sum_r u[r] * W[r,c]   (u in smem, result in smem t[cols]); scratch part[SN_WARPS][32*VEC]
template <int VEC>
__device__ void wt_u(const float* __restrict__ W, const float* u_s, float* t_s, float* part, int rows, int cols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = 32 * VEC;
  for (int c0 = 0; c0 < cols; c0 += chunk) {
    const int c = c0 + lane * VEC;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    if (c < cols) {
#pragma unroll 4
      for (int r = warp; r < rows; r += SN_WARPS) {
        const float ur = u_s[r];
        if constexpr (VEC == 4) {
          const float4 w = *reinterpret_cast<const float4*>(W + (size_t)r * cols + c);
          acc[0] += ur * w.x; acc[1] += ur * w.y; acc[2] += ur * w.z; acc[3] += ur * w.w;
        } else {
          acc[0] += ur * W[(size_t)r * cols + c];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) part[warp * chunk + lane * VEC + j] = acc[j];
    __syncthreads();
    for (int i = threadIdx.x; i < chunk; i += SN_THREADS) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < SN_WARPS; ++w) s += part[w * chunk + i];
      if (c0 + i < cols) t_s[c0 + i] = s;
    }
    __syncthreads();
  }
}

// s[r] = sum_c W[r,c] * v[c]   (v in smem, result in smem s[rows])
__device__ void w_v(const float* __restrict__ W, const float* v_s, float* s_s, int rows, int cols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += SN_WARPS) {
    const float* wr = W + (size_t)r * cols;
    float acc = 0.f;
    if ((cols & 3) == 0) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 w = *reinterpret_cast<const float4*>(wr + c);
        acc += w.x * v_s[c] + w.y * v_s[c + 1] + w.z * v_s[c + 2] + w.w * v_s[c + 3];
      }
    } else {
      for (int c = lane; c < cols; c += 32) acc += wr[c] * v_s[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) s_s[r] = acc;
  }
  __syncthreads();
}

__device__ float sumsq_smem(const float* x, int n, float* red) {
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += SN_THREADS) a += x[i] * x[i];
  return block_sum(a, red);
}

__global__ void __launch_bounds__(SN_THREADS)
sn_forward_kernel(const calm_sn_layer* __restrict__ table, int training, float eps, int max_rows, int max_cols) {
  extern __shared__ float sm[];
  const calm_sn_layer L = table[blockIdx.x];
  const int rows = L.rows, cols = L.cols;
  float* u_s = sm;                       // rows  (also holds s = W v)
  float* v_s = sm + max_rows;            // cols  (also holds t = W^T u)
  float* part = sm + max_rows + max_cols;  // SN_WARPS*128 floats of reduction scratch
  __shared__ float red[32];
  __shared__ float tile[32][33];

  for (int i = threadIdx.x; i < rows; i += SN_THREADS) u_s[i] = L.u[i];
  for (int i = threadIdx.x; i < cols; i += SN_THREADS) v_s[i] = L.v[i];
  __syncthreads();

  float sigma;
  if (training) {
    // v <- normalize(W^T u)
    if ((cols & 3) == 0) wt_u<4>(L.w, u_s, v_s, part, rows, cols);
    else                 wt_u<1>(L.w, u_s, v_s, part, rows, cols);
    const float nv = sqrtf(sumsq_smem(v_s, cols, red));
    const float inv_nv = 1.0f / fmaxf(nv, eps);
    for (int i = threadIdx.x; i < cols; i += SN_THREADS) { const float x = v_s[i] * inv_nv; v_s[i] = x; L.v[i] = x; }
    __syncthreads();
    // u <- normalize(W v) ; sigma = u . (W v)
    w_v(L.w, v_s, u_s, rows, cols);
    const float ss = sumsq_smem(u_s, rows, red);
    const float ns = sqrtf(ss);
    const float inv_ns = 1.0f / fmaxf(ns, eps);
    for (int i = threadIdx.x; i < rows; i += SN_THREADS) L.u[i] = u_s[i] * inv_ns;
    sigma = ss * inv_ns;
  } else {
    // sigma = u . (W v) with the stored buffers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc_total = 0.f;
    for (int r = warp; r < rows; r += SN_WARPS) {
      const float* wr = L.w + (size_t)r * cols;
      float acc = 0.f;
      for (int c = lane; c < cols; c += 32) acc += wr[c] * v_s[c];
      acc = warp_sum(acc);
      if (lane == 0) acc_total += acc * u_s[r];
    }
    sigma = block_sum(acc_total, red);
  }
  if (threadIdx.x == 0) *L.sigma = sigma;
  const float inv_sigma = 1.0f / sigma;

  // W_eff = rowscale * W / sigma  -> bf16 (rows, cols) [+ transposed bf16 (cols, rows)] or fp32
  if (L.eff_f32) {
    float* out = reinterpret_cast<float*>(L.w_eff);
    for (int i = threadIdx.x; i < rows * cols; i += SN_THREADS) {
      const float rs = L.rowscale ? L.rowscale[i / cols] : 1.f;
      out[i] = L.w[i] * inv_sigma * rs;
    }
    return;
  }
  bf16* out = reinterpret_cast<bf16*>(L.w_eff);
  bf16* out_t = reinterpret_cast<bf16*>(L.w_eff_t);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 16
  for (int r0 = 0; r0 < rows; r0 += 32) {
    for (int c0 = 0; c0 < cols; c0 += 32) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = r0 + ty + j * 16, c = c0 + tx;
        float val = 0.f;
        if (r < rows && c < cols) {
          const float rs = L.rowscale ? L.rowscale[r] : 1.f;
          val = L.w[(size_t)r * cols + c] * inv_sigma * rs;
          out[(size_t)r * cols + c] = __float2bfloat16(val);
        }
        tile[ty + j * 16][tx] = val;
      }
      __syncthreads();
      if (out_t) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = c0 + ty + j * 16, r = r0 + tx;
          if (r < rows && c < cols) out_t[(size_t)c * L.ld_t + r] = __float2bfloat16(tile[tx][ty + j * 16]);
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward:  H = rs (.) sum_s G_s ;  dW = H/sigma - (<H,W>/sigma^2) u v^T ;  d rs[r] = sum_c G[r,c] W[r,c]/sigma
// grid (n_layers, SN_BWD_CHUNKS); kernel A writes H into grad_w and per-chunk partial dots into tmp.
// ------------------------------------------------------------------------------------------------------------
constexpr int SN_BWD_CHUNKS = 16;
constexpr int SN_BWD_THREADS = 256;

__global__ void __launch_bounds__(SN_BWD_THREADS)
sn_backward_a_kernel(const calm_sn_layer* __restrict__ table) {
  const calm_sn_layer L = table[blockIdx.x];
  const int rows = L.rows, cols = L.cols;
  const int rpc = (rows + SN_BWD_CHUNKS - 1) / SN_BWD_CHUNKS;
  const int r_begin = blockIdx.y * rpc, r_end = min(rows, r_begin + rpc);
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = SN_BWD_THREADS / 32;
  const size_t plane = (size_t)L.g_split_stride;
  float dot = 0.f;
  for (int r = r_begin + warp; r < r_end; r += nwarps) {
    const float rs = L.rowscale ? L.rowscale[r] : 1.f;
    float rowdot = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const size_t idx = (size_t)r * cols + c;
      float g = 0.f;
      for (int s = 0; s < L.g_splits; ++s) g += L.g_eff[s * plane + idx];
      const float w = L.w[idx];
      rowdot += g * w;
      L.grad_w[idx] = g * rs;
    }
    rowdot = warp_sum(rowdot);
    if (lane == 0) {
      if (L.grad_rowscale) L.grad_rowscale[r] = rowdot / (*L.sigma);
      dot += rowdot * rs;
    }
  }
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) L.tmp[blockIdx.y] = dot;
}

__global__ void __launch_bounds__(SN_BWD_THREADS)
sn_backward_b_kernel(const calm_sn_layer* __restrict__ table) {
  const calm_sn_layer L = table[blockIdx.x];
  const int rows = L.rows, cols = L.cols;
  const int rpc = (rows + SN_BWD_CHUNKS - 1) / SN_BWD_CHUNKS;
  const int r_begin = blockIdx.y * rpc, r_end = min(rows, r_begin + rpc);
  if (r_begin >= r_end) return;
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < SN_BWD_CHUNKS; ++i) dot += L.tmp[i];
  const float sigma = *L.sigma;
  const float inv_sigma = 1.0f / sigma;
  const float coef = dot * inv_sigma * inv_sigma;
  const int n = (r_end - r_begin) * cols;
  for (int i = threadIdx.x; i < n; i += SN_BWD_THREADS) {
    const int r = r_begin + i / cols, c = i % cols;
    const size_t idx = (size_t)r * cols + c;
    L.grad_w[idx] = L.grad_w[idx] * inv_sigma - coef * L.u[r] * L.v[c];
  }
}

}  // namespace

extern "C" int32_t calm_sn_forward(const calm_sn_layer* table_dev, int32_t n_layers, int32_t max_rows, int32_t max_cols,
                                   int32_t training, float eps, cudaStream_t stream) {
  CALM_CHECK_ARG(table_dev != nullptr && n_layers > 0, "calm_sn_forward: empty table");
  CALM_CHECK_ARG(max_rows > 0 && max_cols > 0, "calm_sn_forward: bad max dims");
  const int mr = (max_rows + 3) & ~3, mc = (max_cols + 3) & ~3;
  const size_t smem = (size_t)(mr + mc + SN_WARPS * 128) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(sn_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { calm_set_error("calm_sn_forward: smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured = smem;
  }
  sn_forward_kernel<<<n_layers, SN_THREADS, smem, stream>>>(table_dev, training, eps, mr, mc);
  CALM_CHECK_LAUNCH("calm_sn_forward");
  return CALM_OK;
}

extern "C" int32_t calm_sn_backward(const calm_sn_layer* table_dev, int32_t n_layers, int32_t max_rows, int32_t max_cols,
                                    cudaStream_t stream) {
  CALM_CHECK_ARG(table_dev != nullptr && n_layers > 0, "calm_sn_backward: empty table");
  (void)max_rows; (void)max_cols;
  dim3 grid(n_layers, SN_BWD_CHUNKS);
  sn_backward_a_kernel<<<grid, SN_BWD_THREADS, 0, stream>>>(table_dev);
  CALM_CHECK_LAUNCH("calm_sn_backward(a)");
  sn_backward_b_kernel<<<grid, SN_BWD_THREADS, 0, stream>>>(table_dev);
  CALM_CHECK_LAUNCH("calm_sn_backward(b)");
  return CALM_OK;
}
