// Batched spectral normalisation for every sn(...) layer of a scope (one Block: ~40 layers, 5 M weights).
// Replaces the per-layer forward-pre-hook of torch.nn.utils.spectral_norm (torch/nn/utils/spectral_norm.py:92-114):
//     v <- normalize(W^T u) ; u <- normalize(W v) ; sigma = u . (W v) ; W_eff = W / sigma
// (same update order, eps and in-place buffer semantics) and writes the bf16 operands the tcgen05 GEMMs read. An optional
// LayerScale vector is folded into the rows of W_eff (out_proj*ls_att, mlp.3*ls_mlp: Vi_Tools_CNN_less_V2.py:300,314).
//
// v2: the work is cut into (layer, row-chunk) ITEMS so that the big matrices are streamed by many CTAs; the two global
// reductions of the power iteration become two tiny per-layer kernels. Five launches per scope (three in eval mode):
//   A  item : partial t = u_chunk^T W_chunk                     (W read #1, HBM)
//   B  layer: t = sum of partials ; v = t / max(|t|, eps)
//   C  item : s_chunk = W_chunk v                               (W read #2, L2)
//   D  layer: u = s / max(|s|, eps) ; sigma = |s|^2 / max(|s|, eps)      (eval: sigma = u . s with the stored u)
//   E  item : W_eff_chunk = rowscale * W_chunk / sigma -> bf16 | fp32   (W read #3, L2)
// Backward:  H = rowscale (.) sum_splits G ;  dW = H/sigma - (<H,W>/sigma^2) u v^T ;
// d rowscale[r] = <G[r,:], W[r,:]>/sigma — item kernel (H + partial dots), then item kernel (rank-1 correction).
// All reductions are fixed-order: results are deterministic.
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

constexpr int SN_THREADS = 256;
constexpr int SN_WARPS = SN_THREADS / 32;

__device__ __forceinline__ float block_sum256(float v, float* red) { return block_sum(v, red); }

// ---- A: partial t[c] = sum_{r in chunk} u[r] W[r,c]
__global__ void __launch_bounds__(SN_THREADS)
sn_a_kernel(const calm_sn_layer* __restrict__ table, const calm_sn_item* __restrict__ items) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const calm_sn_item it = items[blockIdx.x];
  const calm_sn_layer L = table[it.layer];
  const int cols = L.cols;
  float* tp = L.tpart + (size_t)it.local_index * cols;
  if ((cols & 3) == 0) {
    for (int c4 = threadIdx.x; c4 < (cols >> 2); c4 += SN_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int r = it.row_begin; r < it.row_end; ++r) {
        const float ur = L.u[r];
        const float4 w = reinterpret_cast<const float4*>(L.w + (size_t)r * cols)[c4];
        acc.x = fmaf(ur, w.x, acc.x); acc.y = fmaf(ur, w.y, acc.y); acc.z = fmaf(ur, w.z, acc.z); acc.w = fmaf(ur, w.w, acc.w);
      }
      reinterpret_cast<float4*>(tp)[c4] = acc;
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += SN_THREADS) {
      float acc = 0.f;
      for (int r = it.row_begin; r < it.row_end; ++r) acc = fmaf(L.u[r], L.w[(size_t)r * cols + c], acc);
      tp[c] = acc;
    }
  }
}

// ---- B: v = normalize(sum of partials)
__global__ void __launch_bounds__(SN_THREADS)
sn_b_kernel(const calm_sn_layer* __restrict__ table, float eps) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const calm_sn_layer L = table[blockIdx.x];
  const int cols = L.cols;
  float ss = 0.f;
  for (int c = threadIdx.x; c < cols; c += SN_THREADS) {
    // four independent partial sums (fixed association: deterministic): the single dependent chain of item_count L2 loads per
    // column made this one-CTA-per-layer kernel ~100 us per launch on the wide layers (0.9 ms of a 224^2 step)
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    const float* tp = L.tpart + c;
    int k = 0;
#pragma unroll 2
    for (; k + 4 <= L.item_count; k += 4) {
      t0 += tp[(size_t)k * cols]; t1 += tp[(size_t)(k + 1) * cols]; t2 += tp[(size_t)(k + 2) * cols]; t3 += tp[(size_t)(k + 3) * cols];
    }
    for (; k < L.item_count; ++k) t0 += tp[(size_t)k * cols];
    const float t = (t0 + t1) + (t2 + t3);
    L.v[c] = t;  // un-normalised for the moment (only this CTA touches v)
    ss = fmaf(t, t, ss);
  }
  ss = block_sum256(ss, red);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  for (int c = threadIdx.x; c < cols; c += SN_THREADS) L.v[c] *= inv;  // same thread wrote the element
}

// ---- C: s[r] = W[r,:] . v   (warp per row)
__global__ void __launch_bounds__(SN_THREADS)
sn_c_kernel(const calm_sn_layer* __restrict__ table, const calm_sn_item* __restrict__ items) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const calm_sn_item it = items[blockIdx.x];
  const calm_sn_layer L = table[it.layer];
  const int cols = L.cols, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = it.row_begin + warp; r < it.row_end; r += SN_WARPS) {
    const float* wr = L.w + (size_t)r * cols;
    float acc = 0.f;
    if ((cols & 3) == 0) {
      for (int c4 = lane; c4 < (cols >> 2); c4 += 32) {
        const float4 w = reinterpret_cast<const float4*>(wr)[c4];
        const float4 v = reinterpret_cast<const float4*>(L.v)[c4];
        acc += w.x * v.x + w.y * v.y + w.z * v.z + w.w * v.w;
      }
    } else {
      for (int c = lane; c < cols; c += 32) acc = fmaf(wr[c], L.v[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) L.svec[r] = acc;
  }
}

// ---- D: u, sigma
__global__ void __launch_bounds__(SN_THREADS)
sn_d_kernel(const calm_sn_layer* __restrict__ table, int training, float eps) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const calm_sn_layer L = table[blockIdx.x];
  const int rows = L.rows;
  float acc = 0.f;
  if (training) {
    for (int r = threadIdx.x; r < rows; r += SN_THREADS) { const float s = L.svec[r]; acc = fmaf(s, s, acc); }
    const float ss = block_sum256(acc, red);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    for (int r = threadIdx.x; r < rows; r += SN_THREADS) L.u[r] = L.svec[r] * inv;
    if (threadIdx.x == 0) *L.sigma = ss * inv;
  } else {
    for (int r = threadIdx.x; r < rows; r += SN_THREADS) acc = fmaf(L.u[r], L.svec[r], acc);
    const float sg = block_sum256(acc, red);
    if (threadIdx.x == 0) *L.sigma = sg;
  }
}

// ---- E: W_eff = rowscale * W / sigma
__global__ void __launch_bounds__(SN_THREADS)
sn_e_kernel(const calm_sn_layer* __restrict__ table, const calm_sn_item* __restrict__ items) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const calm_sn_item it = items[blockIdx.x];
  const calm_sn_layer L = table[it.layer];
  const int cols = L.cols;
  const float inv_sigma = 1.0f / (*L.sigma);
  const size_t base = (size_t)it.row_begin * cols;
  const int n = (it.row_end - it.row_begin) * cols;
  if (L.eff_f32) {
    float* out = reinterpret_cast<float*>(L.w_eff);
    for (int i = threadIdx.x; i < n; i += SN_THREADS) {
      const float rs = L.rowscale ? L.rowscale[it.row_begin + i / cols] : 1.f;
      out[base + i] = L.w[base + i] * inv_sigma * rs;
    }
  } else if ((cols & 3) == 0) {
    bf16* out = reinterpret_cast<bf16*>(L.w_eff);
    const int c4n = cols >> 2;
    for (int i = threadIdx.x; i < (n >> 2); i += SN_THREADS) {
      const int r = it.row_begin + i / c4n;
      const float sc = inv_sigma * (L.rowscale ? L.rowscale[r] : 1.f);
      const float4 w = reinterpret_cast<const float4*>(L.w + base)[i];
      uint2 o;
      o.x = pack_bf16x2(w.x * sc, w.y * sc); o.y = pack_bf16x2(w.z * sc, w.w * sc);
      reinterpret_cast<uint2*>(out + base)[i] = o;
    }
  } else {
    bf16* out = reinterpret_cast<bf16*>(L.w_eff);
    for (int i = threadIdx.x; i < n; i += SN_THREADS) {
      const float rs = L.rowscale ? L.rowscale[it.row_begin + i / cols] : 1.f;
      out[base + i] = __float2bfloat16(L.w[base + i] * inv_sigma * rs);
    }
  }
}

// ---- backward A: H = rowscale * sum_splits G -> grad_w ; per-item partial <H, W> ; d rowscale
__global__ void __launch_bounds__(SN_THREADS)
sn_bwd_a_kernel(const calm_sn_layer* __restrict__ table, const calm_sn_item* __restrict__ items) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const calm_sn_item it = items[blockIdx.x];
  const calm_sn_layer L = table[it.layer];
  const int cols = L.cols, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)L.g_split_stride;
  const float inv_sigma = 1.0f / (*L.sigma);
  float dot = 0.f;
  for (int r = it.row_begin + warp; r < it.row_end; r += SN_WARPS) {
    const float rs = L.rowscale ? L.rowscale[r] : 1.f;
    const size_t rowoff = (size_t)r * cols;
    float rowdot = 0.f;
    if ((cols & 3) == 0) {
      for (int c4 = lane; c4 < (cols >> 2); c4 += 32) {
        // split-K partials: up to 8 planes requested before the first add (a runtime-length loop of dependent
        // load -> add steps ran one memory latency per split: 32 of them on the small layers); summed in split order
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* gp = reinterpret_cast<const float4*>(L.g_eff + rowoff) + c4;
        const size_t plane4 = plane >> 2;
        for (int s0 = 0; s0 < L.g_splits; s0 += 8) {
          float4 p[8];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (s0 + k < L.g_splits) p[k] = __ldcs(gp + (size_t)(s0 + k) * plane4);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (s0 + k < L.g_splits) { g.x += p[k].x; g.y += p[k].y; g.z += p[k].z; g.w += p[k].w; }
        }
        const float4 w = reinterpret_cast<const float4*>(L.w + rowoff)[c4];
        rowdot += g.x * w.x + g.y * w.y + g.z * w.z + g.w * w.w;
        reinterpret_cast<float4*>(L.grad_w + rowoff)[c4] = make_float4(g.x * rs, g.y * rs, g.z * rs, g.w * rs);
      }
    } else {
      for (int c = lane; c < cols; c += 32) {
        float g = 0.f;
        for (int s = 0; s < L.g_splits; ++s) g += L.g_eff[s * plane + rowoff + c];
        rowdot = fmaf(g, L.w[rowoff + c], rowdot);
        L.grad_w[rowoff + c] = g * rs;
      }
    }
    rowdot = warp_sum(rowdot);
    if (lane == 0) {
      if (L.grad_rowscale) L.grad_rowscale[r] = rowdot * inv_sigma;
      dot += rowdot * rs;
    }
  }
  dot = block_sum256(dot, red);
  if (threadIdx.x == 0) L.tpart[it.local_index] = dot;  // tpart is free again in backward: one float per item
}

// ---- backward B: grad_w = H / sigma - (<H,W> / sigma^2) u v^T
__global__ void __launch_bounds__(SN_THREADS)
sn_bwd_b_kernel(const calm_sn_layer* __restrict__ table, const calm_sn_item* __restrict__ items) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const calm_sn_item it = items[blockIdx.x];
  const calm_sn_layer L = table[it.layer];
  const int cols = L.cols;
  float dot = 0.f;
  for (int k = 0; k < L.item_count; ++k) dot += L.tpart[k];
  const float inv_sigma = 1.0f / (*L.sigma);
  const float coef = dot * inv_sigma * inv_sigma;
  const size_t base = (size_t)it.row_begin * cols;
  const int n = (it.row_end - it.row_begin) * cols;
  for (int i = threadIdx.x; i < n; i += SN_THREADS) {
    const int r = it.row_begin + i / cols, c = i % cols;
    L.grad_w[base + i] = L.grad_w[base + i] * inv_sigma - coef * L.u[r] * L.v[c];
  }
}

}  // namespace

extern "C" int32_t calm_sn_forward(const calm_sn_layer* table_dev, int32_t n_layers, const calm_sn_item* items_dev, int32_t n_items,
                                   int32_t training, float eps, cudaStream_t stream) {
  CALM_CHECK_ARG(table_dev != nullptr && n_layers > 0 && items_dev != nullptr && n_items >= n_layers, "calm_sn_forward: empty table");
  if (training) {
    CALM_LAUNCH((sn_a_kernel), n_items, SN_THREADS, 0, stream, table_dev, items_dev);
    CALM_CHECK_LAUNCH("calm_sn_forward(A)");
    CALM_LAUNCH((sn_b_kernel), n_layers, SN_THREADS, 0, stream, table_dev, eps);
    CALM_CHECK_LAUNCH("calm_sn_forward(B)");
  }
  CALM_LAUNCH((sn_c_kernel), n_items, SN_THREADS, 0, stream, table_dev, items_dev);
  CALM_CHECK_LAUNCH("calm_sn_forward(C)");
  CALM_LAUNCH((sn_d_kernel), n_layers, SN_THREADS, 0, stream, table_dev, training, eps);
  CALM_CHECK_LAUNCH("calm_sn_forward(D)");
  CALM_LAUNCH((sn_e_kernel), n_items, SN_THREADS, 0, stream, table_dev, items_dev);
  CALM_CHECK_LAUNCH("calm_sn_forward(E)");
  return CALM_OK;
}

extern "C" int32_t calm_sn_backward(const calm_sn_layer* table_dev, int32_t n_layers, const calm_sn_item* items_dev, int32_t n_items,
                                    cudaStream_t stream) {
  CALM_CHECK_ARG(table_dev != nullptr && n_layers > 0 && items_dev != nullptr && n_items >= n_layers, "calm_sn_backward: empty table");
  CALM_LAUNCH((sn_bwd_a_kernel), n_items, SN_THREADS, 0, stream, table_dev, items_dev);
  CALM_CHECK_LAUNCH("calm_sn_backward(a)");
  CALM_LAUNCH((sn_bwd_b_kernel), n_items, SN_THREADS, 0, stream, table_dev, items_dev);
  CALM_CHECK_LAUNCH("calm_sn_backward(b)");
  return CALM_OK;
}
