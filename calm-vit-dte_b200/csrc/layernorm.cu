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

// One warp per row. The row lives in registers (VPL float4 per lane, D <= 128 VPL): one pass over global memory, mean and the
// centred second moment from the registers. VPL = 0 is the generic three-pass form (re-reads hit the L1).
template <bool OUT_F32, int VPL>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, void* __restrict__ y, float* __restrict__ mean,
              float* __restrict__ rstd, long long rows, int D, float eps) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * LN_WARPS + warp;
  if (row >= rows) return;
  const float* xr = x + row * D;
  const int nv = D >> 2;  // D % 4 == 0
  if (VPL > 0) {
    float4 v[VPL > 0 ? VPL : 1];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + 32 * k;
      v[k] = i < nv ? __ldcs(reinterpret_cast<const float4*>(xr) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    const float mu = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (lane + 32 * k < nv) {
        const float a = v[k].x - mu, b = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
        q += a * a + b * b + c * c + d * d;
      }
    }
    const float rs = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + 32 * k;
      if (i < nv) {
        const float4 g = reinterpret_cast<const float4*>(w)[i];
        const float a = (v[k].x - mu) * rs * g.x, b = (v[k].y - mu) * rs * g.y, c = (v[k].z - mu) * rs * g.z, d = (v[k].w - mu) * rs * g.w;
        if (OUT_F32) {
          reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * D)[i] = make_float4(a, b, c, d);
        } else {
          uint2 o; o.x = pack_bf16x2(a, b); o.y = pack_bf16x2(c, d);
          reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + row * D)[i] = o;
        }
      }
    }
    return;
  }
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

// First block only: the row tokenisation of the NCHW image (Vi_Tools_CNN_less_V2.py:389-391: permute(0,2,3,1) +
// reshape) fused into the block's first LayerNorm. One warp per token (b, image row): lane owns pixels lane, lane + 32, ...; the
// three channel planes are read coalesced, the fp32 token row (the residual stream's first tensor) and the bf16 LayerNorm output
// are written in the same pass — the stand-alone tokenise kernel's extra read of the batch is gone.
constexpr int LN_IMG_PPL = 16;     // pixels per lane: S <= 512
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_image_kernel(const float* __restrict__ img, const float* __restrict__ w, bf16* __restrict__ y, float* __restrict__ tokens,
                    float* __restrict__ mean, float* __restrict__ rstd, int B, int S, float eps) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * LN_WARPS + warp;      // token index b * S + image row
  if (row >= (long long)B * S) return;
  const int b = (int)(row / S), yr = (int)(row - (long long)b * S);
  const int D = 3 * S;
  const float* p0 = img + ((long long)b * 3 * S + yr) * S;              // channel c at p0 + c * S * S
  const long long plane = (long long)S * S;
  float v[LN_IMG_PPL][3];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_IMG_PPL; ++k) {
    const int xc = lane + 32 * k;
    if (xc < S) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { v[k][c] = __ldcs(p0 + c * plane + xc); s += v[k][c]; }
    } else {
      v[k][0] = v[k][1] = v[k][2] = 0.f;
    }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_IMG_PPL; ++k)
    if (lane + 32 * k < S) {
      const float a = v[k][0] - mu, bb = v[k][1] - mu, c = v[k][2] - mu;
      q += a * a + bb * bb + c * c;
    }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  float* trow = tokens + row * D;
  bf16* yrow = y + row * D;
#pragma unroll
  for (int k = 0; k < LN_IMG_PPL; ++k) {
    const int xc = lane + 32 * k;
    if (xc < S) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        trow[xc * 3 + c] = v[k][c];
        yrow[xc * 3 + c] = __float2bfloat16((v[k][c] - mu) * rs * w[xc * 3 + c]);
      }
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
              float* __restrict__ dx, bf16* __restrict__ dx16, float* __restrict__ dw_partial, long long rows, int D) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
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
      if (dx16) reinterpret_cast<uint2*>(dx16 + row * D)[i] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
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

// Register-resident variant for D <= 128 * VPL * WPR: each row is read ONCE (x, dy stay in registers between the statistics and
// the output pass), a lane owns fixed columns so the dw accumulators live in registers too (no shared-memory
// read-modify-write per element); shared memory is only used to combine the warps of a CTA at the very end.
// WPR = warps per row: 1 for D <= 768; 2 for the wide rows of the 384^2 / 512^2 configs (D up to 1536), where one warp per row
// needed VPL = 12 (245 registers, one 8-warp CTA per SM, the residual-gradient load after the row reduction): 0.24 of the HBM rate.
template <int VPL, bool DY_F32, int WPR>
__global__ void __launch_bounds__(LN_THREADS, VPL <= 3 ? 3 : VPL <= 6 ? 2 : 1)
ln_bwd_reg_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
                  float* __restrict__ dx, bf16* __restrict__ dx16, float* __restrict__ dw_partial, long long rows, int D) {
  extern __shared__ float acc_s[];  // (LN_WARPS / WPR) * D, then D floats of w
  __shared__ float pair_red[2][LN_WARPS];
  pdl_wait(); pdl_launch_small_dependent();   // the dw reduction kernel may be scheduled as CTAs of this grid retire (it waits for the whole grid)
  constexpr int RPC = LN_WARPS / WPR;      // rows per CTA iteration
  constexpr int LANES = 32 * WPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = warp % WPR, slot = warp / WPR;
  const int l = sub * 32 + lane;           // lane index within the row
  const int nv = D >> 2;
  // w is read from shared memory at every use: keeping it in registers next to x-hat, dy and the dw accumulators costs
  // 128-138 registers per thread = one 8-warp CTA per SM, too few loads in flight for the HBM (44 % of peak at D = 672)
  float* w_s = acc_s + (size_t)RPC * D;
  for (int i = threadIdx.x; i < D; i += LN_THREADS) w_s[i] = w[i];
  __syncthreads();
  const float4* wv = reinterpret_cast<const float4*>(w_s) + l;
  float4 dwa[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) dwa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float invD = 1.0f / D;
  // WPR > 1: the CTA walks the rows in lock step (the two warps of a row meet on __syncthreads), so the trip count is uniform
  for (long long base = (long long)blockIdx.x * RPC; base < rows; base += (long long)gridDim.x * RPC) {
    const long long row = base + slot;
    const bool active = row < rows;
    if (WPR == 1 && !active) break;
    const float mu = active ? mean[row] : 0.f, rs = active ? rstd[row] : 0.f;
    float4 hv[VPL], dv[VPL];
    // the residual-stream gradient is requested together with x and dy (VPL <= 6: it fits the register budget of 2 CTAs / SM):
    // loaded after the row reduction it put one full HBM latency per row on the critical path of a warp (0.45 of the copy rate)
    constexpr bool EARLY = VPL <= 6;
    float4 rv[EARLY ? VPL : 1];
    if (EARLY && dres) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int idx = l + LANES * i;
        rv[EARLY ? i : 0] = (active && idx < nv) ? __ldcs(reinterpret_cast<const float4*>(dres + row * D) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int idx = l + LANES * i;
      if (active && idx < nv) {
        const float4 xv = reinterpret_cast<const float4*>(x + row * D)[idx];
        dv[i] = load_dy4<DY_F32>(dy, row, D, idx);
        hv[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      } else {
        dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        hv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float4 wq = idx < nv ? wv[LANES * i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float g0 = dv[i].x * wq.x, g1 = dv[i].y * wq.y, g2 = dv[i].z * wq.z, g3 = dv[i].w * wq.w;
      s1 += g0 + g1 + g2 + g3;
      s2 += g0 * hv[i].x + g1 * hv[i].y + g2 * hv[i].z + g3 * hv[i].w;
      dwa[i].x = fmaf(dv[i].x, hv[i].x, dwa[i].x); dwa[i].y = fmaf(dv[i].y, hv[i].y, dwa[i].y);
      dwa[i].z = fmaf(dv[i].z, hv[i].z, dwa[i].z); dwa[i].w = fmaf(dv[i].w, hv[i].w, dwa[i].w);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (WPR > 1) {                                   // combine the warps of the row (fixed order: deterministic)
      if (lane == 0) { pair_red[0][warp] = s1; pair_red[1][warp] = s2; }
      __syncthreads();
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int k = 0; k < WPR; ++k) { s1 += pair_red[0][slot * WPR + k]; s2 += pair_red[1][slot * WPR + k]; }
      __syncthreads();                               // the slots are rewritten in the next iteration
    }
    const float c2 = s1 * invD, c1 = s2 * invD;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int idx = l + LANES * i;
      if (active && idx < nv) {
        const float4 wq = wv[LANES * i];
        float4 o;
        o.x = rs * (dv[i].x * wq.x - c2 - hv[i].x * c1);
        o.y = rs * (dv[i].y * wq.y - c2 - hv[i].y * c1);
        o.z = rs * (dv[i].z * wq.z - c2 - hv[i].z * c1);
        o.w = rs * (dv[i].w * wq.w - c2 - hv[i].w * c1);
        if (dres) {
          const float4 r = EARLY ? rv[EARLY ? i : 0] : reinterpret_cast<const float4*>(dres + row * D)[idx];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(dx + row * D)[idx] = o;
        // optional bf16 copy: the GEMMs that consume this gradient next (dgrad / wgrad of the producing Linear) take bf16
        // operands; writing it here replaces a separate 4-byte-read + 2-byte-write cast pass over the tensor
        if (dx16) reinterpret_cast<uint2*>(dx16 + row * D)[idx] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = l + LANES * i;
    if (idx < nv) reinterpret_cast<float4*>(acc_s + (size_t)slot * D)[idx] = dwa[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += LN_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int wv2 = 0; wv2 < RPC; ++wv2) s += acc_s[(size_t)wv2 * D + c];
    dw_partial[(size_t)blockIdx.x * D + c] = s;
  }
}

// out[c] = sum_p partial[p][c]: 32 columns per CTA, 32 row-groups (one warp each) reduced through shared memory (deterministic).
// With 8 row-groups a thread walked ~74 partial rows one dependent-latency at a time (12.6 us for 1.6 MB, 57 launches per step).
constexpr int RP_GROUPS = 32;
__global__ void __launch_bounds__(32 * RP_GROUPS)
reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int n) {
  __shared__ float red[RP_GROUPS][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  pdl_wait();                                   // launched while the backward kernel drains: its partial rows are complete from here on
  float s0 = 0.f, s1 = 0.f;
  if (c < n) {
    int p = ry;
    for (; p + RP_GROUPS < nparts; p += 2 * RP_GROUPS) {       // two independent loads in flight per thread
      s0 += partial[(size_t)p * n + c];
      s1 += partial[(size_t)(p + RP_GROUPS) * n + c];
    }
    if (p < nparts) s0 += partial[(size_t)p * n + c];
  }
  red[ry][cx] = s0 + s1;
  __syncthreads();
  if (ry == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < RP_GROUPS; ++k) t += red[k][cx];
    out[c] = t;
  }
}

}  // namespace

extern "C" int32_t calm_layernorm_fwd(const float* x, const float* w, void* y, int32_t y_dtype, float* mean, float* rstd,
                                      int64_t rows, int32_t D, float eps, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0, "calm_layernorm_fwd: rows=%lld D=%d (D must be a multiple of 4)", (long long)rows, D);
  const unsigned grid = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
  const int vpl = (D / 4 + 31) / 32;  // float4 per lane and row
#define LN_FWD_LAUNCH(F32, VPL) CALM_LAUNCH((ln_fwd_kernel<F32, VPL>), grid, LN_THREADS, 0, stream, x, w, y, mean, rstd, rows, D, eps)
  if (y_dtype == CALM_F32) {
    if (vpl <= 2) LN_FWD_LAUNCH(true, 2); else if (vpl <= 4) LN_FWD_LAUNCH(true, 4); else if (vpl <= 6) LN_FWD_LAUNCH(true, 6);
    else if (vpl <= 12) LN_FWD_LAUNCH(true, 12); else LN_FWD_LAUNCH(true, 0);
  } else {
    if (vpl <= 2) LN_FWD_LAUNCH(false, 2); else if (vpl <= 4) LN_FWD_LAUNCH(false, 4); else if (vpl <= 6) LN_FWD_LAUNCH(false, 6);
    else if (vpl <= 12) LN_FWD_LAUNCH(false, 12); else LN_FWD_LAUNCH(false, 0);
  }
#undef LN_FWD_LAUNCH
  CALM_CHECK_LAUNCH("calm_layernorm_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_layernorm_fwd_image(const float* img, const float* w, void* y_bf16, float* tokens, float* mean, float* rstd,
                                            int32_t B, int32_t S, float eps, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0 && S <= 32 * LN_IMG_PPL, "calm_layernorm_fwd_image: B=%d S=%d (S <= %d)", B, S, 32 * LN_IMG_PPL);
  const long long rows = (long long)B * S;
  CALM_LAUNCH((ln_fwd_image_kernel), (unsigned)((rows + LN_WARPS - 1) / LN_WARPS), LN_THREADS, 0, stream, img, w, reinterpret_cast<bf16*>(y_bf16), tokens, mean,
                                                                                              rstd, B, S, eps);
  CALM_CHECK_LAUNCH("calm_layernorm_fwd_image");
  return CALM_OK;
}

extern "C" int32_t calm_layernorm_bwd_parts(int64_t rows, int32_t D) {
  (void)D;
  long long need = (rows + LN_WARPS - 1) / LN_WARPS;
  const long long cap = 4LL * calm_num_sms();
  return (int32_t)(need < cap ? need : cap);
}

extern "C" int32_t calm_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* w, const float* mean,
                                      const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dw_partial,
                                      int32_t nparts, float* dw, int64_t rows, int32_t D, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0, "calm_layernorm_bwd: rows=%lld D=%d", (long long)rows, D);
  CALM_CHECK_ARG(nparts == calm_layernorm_bwd_parts(rows, D), "calm_layernorm_bwd: nparts=%d, expected %d", nparts, calm_layernorm_bwd_parts(rows, D));
  const size_t smem = (size_t)(LN_WARPS + 1) * D * sizeof(float);   // per-warp dw rows + w
  const bool f32 = dy_dtype == CALM_F32;
  const int vpl = (D / 4 + 31) / 32;  // float4 per lane and row
#define LN_BWD_LAUNCH(KERNEL)                                                                                         \
  do {                                                                                                                \
    if (smem > 48 * 1024) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
      if (e != cudaSuccess) { calm_set_error("calm_layernorm_bwd: smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; } \
    }                                                                                                                 \
    CALM_LAUNCH((KERNEL), nparts, LN_THREADS, smem, stream, dy, x, w, mean, rstd, dres, dx, reinterpret_cast<bf16*>(dx_bf16), dw_partial, rows, D);                  \
  } while (0)
  if (vpl <= 3) { if (f32) LN_BWD_LAUNCH((ln_bwd_reg_kernel<3, true, 1>)); else LN_BWD_LAUNCH((ln_bwd_reg_kernel<3, false, 1>)); }
  else if (vpl <= 6) { if (f32) LN_BWD_LAUNCH((ln_bwd_reg_kernel<6, true, 1>)); else LN_BWD_LAUNCH((ln_bwd_reg_kernel<6, false, 1>)); }
  else if (vpl <= 12) { if (f32) LN_BWD_LAUNCH((ln_bwd_reg_kernel<6, true, 2>)); else LN_BWD_LAUNCH((ln_bwd_reg_kernel<6, false, 2>)); }
  else { if (f32) LN_BWD_LAUNCH((ln_bwd_kernel<true>)); else LN_BWD_LAUNCH((ln_bwd_kernel<false>)); }
#undef LN_BWD_LAUNCH
  CALM_CHECK_LAUNCH("calm_layernorm_bwd");
  {
    cudaError_t e = calm_launch_pdl(reduce_partials_kernel, dim3((D + 31) / 32), dim3(32 * RP_GROUPS), 0, stream, nullptr, 0,
                                    (const float*)dw_partial, dw, nparts, D);
    if (e != cudaSuccess) { calm_set_error("calm_layernorm_bwd(reduce): launch failed: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
  }
  return CALM_OK;
}
