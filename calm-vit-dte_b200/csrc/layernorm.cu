// Weight-only LayerNorm (eps 1e-6, no bias) forward/backward: ln_q, ln_kv, ln_2, ln_final
// (Vi_Tools_CNN_less_V2.py:131-132,197,494; used :211-215,311,523).  HBM-bound warp-per-row kernels:
//   fwd : 4 B read + 2 B write per element (fp32 residual stream in, bf16 GEMM operand out)
//   bwd : dx = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*w   (+ fused residual-gradient add)
//         dw = sum_rows dy*xhat via per-CTA partials + a deterministic second-stage reduce (no atomics)
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

constexpr int LN_WARPS = 8;
constexpr int LN_THREADS = LN_WARPS * 32;

template <bool OUT_F32>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, void* __restrict__ y, float* __restrict__ mean,
              float* __restrict__ rstd, long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * LN_WARPS + warp;
  if (row >= rows) return;
  const float* xr = x + row * D;
  const int nv = D >> 2;  // D % 4 == 0
  float s = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    s += v.x + v.y + v.z + v.w;
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    const float a = v.x - mu, b = v.y - mu, c = v.z - mu, d = v.w - mu;
    q += a * a + b * b + c * c + d * d;
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  for (int i = lane; i < nv; i += 32) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    const float4 g = reinterpret_cast<const float4*>(w)[i];
    const float a = (v.x - mu) * rs * g.x, b = (v.y - mu) * rs * g.y, c = (v.z - mu) * rs * g.z, d = (v.w - mu) * rs * g.w;
    if (OUT_F32) {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * D)[i] = make_float4(a, b, c, d);
    } else {
      uint2 o; o.x = pack_bf16x2(a, b); o.y = pack_bf16x2(c, d);
      reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + row * D)[i] = o;
    }
  }
}

template <bool DY_F32>
__device__ __forceinline__ float4 load_dy4(const void* dy, long long row, int D, int i) {
  if (DY_F32) return reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * D)[i];
  const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + row * D)[i];
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// grid = nparts CTAs, each strides over rows; dynamic smem = LN_WARPS * D floats (per-warp dw accumulators)
template <bool DY_F32>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
              float* __restrict__ dx, float* __restrict__ dw_partial, long long rows, int D) {
  extern __shared__ float acc_s[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* my = acc_s + (size_t)warp * D;
  for (int i = lane; i < D; i += 32) my[i] = 0.f;
  const int nv = D >> 2;
  for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < nv; i += 32) {
      const float4 v = reinterpret_cast<const float4*>(xr)[i];
      const float4 g = reinterpret_cast<const float4*>(w)[i];
      const float4 d = load_dy4<DY_F32>(dy, row, D, i);
      const float h0 = (v.x - mu) * rs, h1 = (v.y - mu) * rs, h2 = (v.z - mu) * rs, h3 = (v.w - mu) * rs;
      const float g0 = d.x * g.x, g1 = d.y * g.y, g2 = d.z * g.z, g3 = d.w * g.w;
      s1 += g0 + g1 + g2 + g3;
      s2 += g0 * h0 + g1 * h1 + g2 * h2 + g3 * h3;
      float4 a = reinterpret_cast<float4*>(my)[i];
      a.x += d.x * h0; a.y += d.y * h1; a.z += d.z * h2; a.w += d.w * h3;
      reinterpret_cast<float4*>(my)[i] = a;
    }
    const float c2 = warp_sum(s1) / D, c1 = warp_sum(s2) / D;
    for (int i = lane; i < nv; i += 32) {
      const float4 v = reinterpret_cast<const float4*>(xr)[i];
      const float4 g = reinterpret_cast<const float4*>(w)[i];
      const float4 d = load_dy4<DY_F32>(dy, row, D, i);
      float4 o;
      o.x = rs * (d.x * g.x - c2 - (v.x - mu) * rs * c1);
      o.y = rs * (d.y * g.y - c2 - (v.y - mu) * rs * c1);
      o.z = rs * (d.z * g.z - c2 - (v.z - mu) * rs * c1);
      o.w = rs * (d.w * g.w - c2 - (v.w - mu) * rs * c1);
      if (dres) {
        const float4 r = reinterpret_cast<const float4*>(dres + row * D)[i];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dx + row * D)[i] = o;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += LN_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < LN_WARPS; ++wv) s += acc_s[(size_t)wv * D + c];
    dw_partial[(size_t)blockIdx.x * D + c] = s;
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n + c];
  out[c] = s;
}

}  // namespace

extern "C" int32_t calm_layernorm_fwd(const float* x, const float* w, void* y, int32_t y_dtype, float* mean, float* rstd,
                                      int64_t rows, int32_t D, float eps, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0, "calm_layernorm_fwd: rows=%lld D=%d (D must be a multiple of 4)", (long long)rows, D);
  const unsigned grid = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
  if (y_dtype == CALM_F32) ln_fwd_kernel<true><<<grid, LN_THREADS, 0, stream>>>(x, w, y, mean, rstd, rows, D, eps);
  else                     ln_fwd_kernel<false><<<grid, LN_THREADS, 0, stream>>>(x, w, y, mean, rstd, rows, D, eps);
  CALM_CHECK_LAUNCH("calm_layernorm_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_layernorm_bwd_parts(int64_t rows, int32_t D) {
  (void)D;
  long long need = (rows + LN_WARPS - 1) / LN_WARPS;
  const long long cap = 2LL * calm_num_sms();
  return (int32_t)(need < cap ? need : cap);
}

extern "C" int32_t calm_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* w, const float* mean,
                                      const float* rstd, const float* dres, float* dx, float* dw_partial, int32_t nparts,
                                      float* dw, int64_t rows, int32_t D, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0, "calm_layernorm_bwd: rows=%lld D=%d", (long long)rows, D);
  CALM_CHECK_ARG(nparts == calm_layernorm_bwd_parts(rows, D), "calm_layernorm_bwd: nparts=%d, expected %d", nparts, calm_layernorm_bwd_parts(rows, D));
  const size_t smem = (size_t)LN_WARPS * D * sizeof(float);
  static size_t configured[2] = {0, 0};
  const int which = dy_dtype == CALM_F32 ? 1 : 0;
  if (smem > 48 * 1024 && smem > configured[which]) {
    cudaError_t e = which ? cudaFuncSetAttribute(ln_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(ln_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { calm_set_error("calm_layernorm_bwd: smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured[which] = smem;
  }
  if (which) ln_bwd_kernel<true><<<nparts, LN_THREADS, smem, stream>>>(dy, x, w, mean, rstd, dres, dx, dw_partial, rows, D);
  else       ln_bwd_kernel<false><<<nparts, LN_THREADS, smem, stream>>>(dy, x, w, mean, rstd, dres, dx, dw_partial, rows, D);
  CALM_CHECK_LAUNCH("calm_layernorm_bwd");
  reduce_partials_kernel<<<(D + 127) / 128, 128, 0, stream>>>(dw_partial, dw, nparts, D);
  CALM_CHECK_LAUNCH("calm_layernorm_bwd(reduce)");
  return CALM_OK;
}
