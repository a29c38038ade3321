// RoPE with learned inverse frequencies (Vi_Tools_CNN_less_V2.py:55-95, NeoX rotate-half) fused with the per-head
// [content | rope] concat of the decoupled-RoPE latent blocks (:278-281). Non-reduce blocks use dc = 0 (:283-285).
// cos/sin are rebuilt from inv_freq every call (the parameter is learned, :70-72,86-91); backward returns d inv_freq.
// All math fp32, storage bf16 — the reference promotes to fp32 against the fp32 cos table and SDPA/bmm re-round to bf16.
#include "common.cuh"
#include <initializer_list>
#include "../../include/calm_b200.h"

namespace {

constexpr int ROPE_BCH = 8;  // batch chunks for the d-theta partial sums

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ cs, int S, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * half) return;
  const int s = i / half, j = i % half;
  const float ang = (float)s * inv_freq[j];
  float sn, cn;
  sincosf(ang, &sn, &cn);
  cs[2 * i] = cn;
  cs[2 * i + 1] = sn;
}

// V consecutive bf16 values moved as one 2V-byte access
template <int V> struct BVec;
template <> struct BVec<1> { using T = unsigned short; };
template <> struct BVec<2> { using T = uint32_t; };
template <> struct BVec<4> { using T = uint2; };
template <> struct BVec<8> { using T = uint4; };
template <int V>
__device__ __forceinline__ void ld_bf16(const bf16* p, float* f) {
  union { typename BVec<V>::T t; bf16 e[V]; } u;
  u.t = *reinterpret_cast<const typename BVec<V>::T*>(p);
#pragma unroll
  for (int e = 0; e < V; ++e) f[e] = __bfloat162float(u.e[e]);
}
template <int V>
__device__ __forceinline__ void st_bf16(bf16* p, const float* f) {
  union { typename BVec<V>::T t; bf16 e[V]; } u;
#pragma unroll
  for (int e = 0; e < V; ++e) u.e[e] = __float2bfloat16(f[e]);
  *reinterpret_cast<typename BVec<V>::T*>(p) = u.t;
}

// grid (S, ROPE_BCH): a CTA owns one sequence position and a slice of the batch. A thread owns V consecutive elements of one
// head (a copied content chunk, or the rotation pair chunks [j, j+V) and [j+half, j+half+V)), loads its cos/sin once and walks
// the images with several independent 2V-byte loads in flight. (The first version moved single bf16 elements: 21 % of the HBM
// rate at 224^2; the index arithmetic and 2-byte accesses were the limit, not the memory system.)
constexpr int ROPE_UNROLL = 4;
template <int V>
__global__ void __launch_bounds__(256)
rope_fwd_kernel(const bf16* __restrict__ content, long long ld_content, const bf16* __restrict__ ropein, long long ld_rope,
                bf16* __restrict__ out, long long ld_out, const float* __restrict__ cs, int B, int S, int heads, int dc, int dr) {
  const int half = dr >> 1;
  const int per_head = (dc + half) / V;
  const int pos = blockIdx.x;
  const int bpc = (B + (int)gridDim.y - 1) / (int)gridDim.y;
  const int b0 = blockIdx.y * bpc, b1 = min(B, b0 + bpc);
  const float2* cs2 = reinterpret_cast<const float2*>(cs) + (long long)pos * half;
  for (int w = threadIdx.x; w < heads * per_head; w += blockDim.x) {
    const int h = w / per_head, i = (w - h * per_head) * V;
    if (i < dc) {
      const bf16* src = content + (long long)h * dc + i;
      bf16* dst = out + (long long)h * (dc + dr) + i;
      for (int b = b0; b < b1; ++b) {
        const long long t = (long long)b * S + pos;
        *reinterpret_cast<typename BVec<V>::T*>(dst + t * ld_out) = *reinterpret_cast<const typename BVec<V>::T*>(src + t * ld_content);
      }
    } else {
      const int j = i - dc;
      float c[V], sn[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { const float2 v = cs2[j + e]; c[e] = v.x; sn[e] = v.y; }
      const bf16* src = ropein + (long long)h * dr + j;
      bf16* dst = out + (long long)h * (dc + dr) + dc + j;
      for (int bb = b0; bb < b1; bb += ROPE_UNROLL) {
        float x1[ROPE_UNROLL][V], x2[ROPE_UNROLL][V];
#pragma unroll
        for (int k = 0; k < ROPE_UNROLL; ++k) {
          if (bb + k < b1) {
            const long long t = (long long)(bb + k) * S + pos;
            ld_bf16<V>(src + t * ld_rope, x1[k]);
            ld_bf16<V>(src + t * ld_rope + half, x2[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < ROPE_UNROLL; ++k) {
          if (bb + k < b1) {
            const long long t = (long long)(bb + k) * S + pos;
            float y1[V], y2[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
              y1[e] = x1[k][e] * c[e] - x2[k][e] * sn[e];
              y2[e] = x2[k][e] * c[e] + x1[k][e] * sn[e];
            }
            st_bf16<V>(dst + t * ld_out, y1);
            st_bf16<V>(dst + t * ld_out + half, y2);
          }
        }
      }
    }
  }
}

// same decomposition; d theta[pos, j] partial sums per (head, j) slot of shared memory, then over the heads (fixed order)
template <int V>
__global__ void __launch_bounds__(256)
rope_bwd_kernel(const bf16* __restrict__ dout, long long ld_dout, const bf16* __restrict__ out, long long ld_out,
                bf16* __restrict__ dcontent, long long ld_dcontent, bf16* __restrict__ dropein, long long ld_drope,
                const float* __restrict__ cs, float* __restrict__ dtheta_part, int B, int S, int heads, int dc, int dr) {
  extern __shared__ float dth_s[];  // heads * half floats (one slot per (head, j): deterministic reduce)
  const int half = dr >> 1;
  const int per_head = (dc + half) / V;
  const int pos = blockIdx.x, chunk = blockIdx.y;
  const int bpc = (B + ROPE_BCH - 1) / ROPE_BCH;
  const int b0 = chunk * bpc, b1 = min(B, b0 + bpc);
  const float2* cs2 = reinterpret_cast<const float2*>(cs) + (long long)pos * half;
  for (int w = threadIdx.x; w < heads * per_head; w += blockDim.x) {
    const int h = w / per_head, i = (w - h * per_head) * V;
    if (i < dc) {
      const bf16* src = dout + (long long)h * (dc + dr) + i;
      bf16* dst = dcontent + (long long)h * dc + i;
      for (int b = b0; b < b1; ++b) {
        const long long t = (long long)b * S + pos;
        *reinterpret_cast<typename BVec<V>::T*>(dst + t * ld_dcontent) = *reinterpret_cast<const typename BVec<V>::T*>(src + t * ld_dout);
      }
    } else {
      const int j = i - dc;
      float c[V], sn[V], acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { const float2 v = cs2[j + e]; c[e] = v.x; sn[e] = v.y; acc[e] = 0.f; }
      const bf16* dyp = dout + (long long)h * (dc + dr) + dc + j;
      const bf16* yp = out + (long long)h * (dc + dr) + dc + j;
      bf16* dxp = dropein + (long long)h * dr + j;
      for (int bb = b0; bb < b1; bb += ROPE_UNROLL) {
        float dy1[ROPE_UNROLL][V], dy2[ROPE_UNROLL][V], y1[ROPE_UNROLL][V], y2[ROPE_UNROLL][V];
#pragma unroll
        for (int k = 0; k < ROPE_UNROLL; ++k) {
          if (bb + k < b1) {
            const long long t = (long long)(bb + k) * S + pos;
            ld_bf16<V>(dyp + t * ld_dout, dy1[k]);
            ld_bf16<V>(dyp + t * ld_dout + half, dy2[k]);
            ld_bf16<V>(yp + t * ld_out, y1[k]);
            ld_bf16<V>(yp + t * ld_out + half, y2[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < ROPE_UNROLL; ++k) {
          if (bb + k < b1) {
            const long long t = (long long)(bb + k) * S + pos;
            float dx1[V], dx2[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
              dx1[e] = dy1[k][e] * c[e] + dy2[k][e] * sn[e];
              dx2[e] = dy2[k][e] * c[e] - dy1[k][e] * sn[e];
              acc[e] += y1[k][e] * dy2[k][e] - y2[k][e] * dy1[k][e];
            }
            st_bf16<V>(dxp + t * ld_drope, dx1);
            st_bf16<V>(dxp + t * ld_drope + half, dx2);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) dth_s[h * half + j + e] = acc[e];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float s = 0.f;
    for (int h = 0; h < heads; ++h) s += dth_s[h * half + j];
    dtheta_part[((size_t)chunk * S + pos) * half + j] = s;
  }
}

// widest vector (elements) all the operands of a call allow
int rope_vec(int dc, int dr, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  for (int v = 8; v > 1; v >>= 1) {
    bool ok = (dr / 2) % v == 0 && dc % v == 0;
    for (long long ld : lds) ok = ok && ld % v == 0;
    for (const void* p : ptrs) ok = ok && (reinterpret_cast<uintptr_t>(p) % (2 * v)) == 0;
    if (ok) return v;
  }
  return 1;
}
int rope_threads(int heads, int dc, int dr, int v) {
  int t = heads * ((dc + dr / 2) / v);
  t = ((t + 31) / 32) * 32;
  return t > 256 ? 256 : t;
}

// d inv_freq[j] = sum_pos pos * sum_chunk dtheta_part[chunk,pos,j]
__global__ void rope_dfreq_kernel(const float* __restrict__ dtheta_part, float* __restrict__ dinv, int S, int half) {
  const int j = blockIdx.x;
  __shared__ float red[32];
  float acc = 0.f;
  for (int pos = threadIdx.x; pos < S; pos += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < ROPE_BCH; ++c) s += dtheta_part[((size_t)c * S + pos) * half + j];
    acc += (float)pos * s;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) dinv[j] = acc;
}

}  // namespace

extern "C" int32_t calm_rope_table(const float* inv_freq, float* cos_sin, int32_t S, int32_t half, cudaStream_t stream) {
  CALM_CHECK_ARG(S > 0 && half > 0, "calm_rope_table: S=%d half=%d", S, half);
  const int n = S * half;
  rope_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(inv_freq, cos_sin, S, half);
  CALM_CHECK_LAUNCH("calm_rope_table");
  return CALM_OK;
}

extern "C" int32_t calm_rope_fwd(const void* content, int64_t ld_content, const void* ropein, int64_t ld_rope, void* out,
                                 int64_t ld_out, const float* cos_sin, int64_t tokens, int32_t S, int32_t heads, int32_t dc,
                                 int32_t dr, cudaStream_t stream) {
  CALM_CHECK_ARG(tokens > 0 && S > 0 && heads > 0 && dr > 0 && dr % 2 == 0 && dc >= 0, "calm_rope_fwd: bad dims");
  CALM_CHECK_ARG(dc == 0 || content != nullptr, "calm_rope_fwd: content missing");
  CALM_CHECK_ARG(tokens % S == 0, "calm_rope_fwd: tokens=%lld is not a multiple of S=%d", (long long)tokens, S);
  const int B = (int)(tokens / S);
  const int v = rope_vec(dc, dr, {(long long)ld_content, (long long)ld_rope, (long long)ld_out}, {content, ropein, out});
  dim3 grid(S, B < ROPE_BCH ? B : ROPE_BCH);
  const int threads = rope_threads(heads, dc, dr, v);
#define CALM_ROPE_FWD(V)                                                                                                          \
  rope_fwd_kernel<V><<<grid, threads, 0, stream>>>(reinterpret_cast<const bf16*>(content), ld_content,                            \
                                                   reinterpret_cast<const bf16*>(ropein), ld_rope, reinterpret_cast<bf16*>(out), \
                                                   ld_out, cos_sin, B, S, heads, dc, dr)
  if (v == 8) CALM_ROPE_FWD(8); else if (v == 4) CALM_ROPE_FWD(4); else if (v == 2) CALM_ROPE_FWD(2); else CALM_ROPE_FWD(1);
#undef CALM_ROPE_FWD
  CALM_CHECK_LAUNCH("calm_rope_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_rope_bwd_scratch_floats(int32_t S, int32_t dr) { return ROPE_BCH * S * (dr / 2); }

extern "C" int32_t calm_rope_bwd(const void* dout, int64_t ld_dout, const void* out, int64_t ld_out, void* dcontent,
                                 int64_t ld_dcontent, void* dropein, int64_t ld_drope, const float* cos_sin, float* dtheta_part,
                                 float* dinv_freq, int64_t tokens, int32_t S, int32_t heads, int32_t dc, int32_t dr,
                                 cudaStream_t stream) {
  CALM_CHECK_ARG(tokens > 0 && S > 0 && tokens % S == 0 && heads > 0 && dr > 0 && dr % 2 == 0 && dc >= 0, "calm_rope_bwd: bad dims");
  CALM_CHECK_ARG(dc == 0 || dcontent != nullptr, "calm_rope_bwd: dcontent missing");
  const int B = (int)(tokens / S), half = dr / 2;
  const int v = rope_vec(dc, dr, {(long long)ld_dout, (long long)ld_out, (long long)ld_dcontent, (long long)ld_drope},
                         {dout, out, dcontent, dropein});
  const int threads = rope_threads(heads, dc, dr, v);
  dim3 grid(S, ROPE_BCH);
  const size_t smem = (size_t)heads * half * sizeof(float);
#define CALM_ROPE_BWD(V)                                                                                                       \
  rope_bwd_kernel<V><<<grid, threads, smem, stream>>>(reinterpret_cast<const bf16*>(dout), ld_dout,                            \
                                                      reinterpret_cast<const bf16*>(out), ld_out, reinterpret_cast<bf16*>(dcontent), \
                                                      ld_dcontent, reinterpret_cast<bf16*>(dropein), ld_drope, cos_sin, dtheta_part, \
                                                      B, S, heads, dc, dr)
  if (v == 8) CALM_ROPE_BWD(8); else if (v == 4) CALM_ROPE_BWD(4); else if (v == 2) CALM_ROPE_BWD(2); else CALM_ROPE_BWD(1);
#undef CALM_ROPE_BWD
  CALM_CHECK_LAUNCH("calm_rope_bwd");
  rope_dfreq_kernel<<<half, 128, 0, stream>>>(dtheta_part, dinv_freq, S, half);
  CALM_CHECK_LAUNCH("calm_rope_bwd(dfreq)");
  return CALM_OK;
}
