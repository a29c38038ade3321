// RoPE with learned inverse frequencies (Vi_Tools_CNN_less_V2.py:55-95, NeoX rotate-half) fused with the per-head
// [content | rope] concat of the decoupled-RoPE latent blocks (:278-281). Non-reduce blocks use dc = 0 (:283-285).
// cos/sin are rebuilt from inv_freq inside every kernel (the parameter is learned, :70-72,86-91); backward returns d inv_freq.
// All math fp32, storage bf16 — the reference promotes to fp32 against the fp32 cos table and SDPA/bmm re-round to bf16.
#include "common.cuh"
#include <initializer_list>
#include "../../include/calm_b200.h"

namespace {

constexpr int ROPE_BCH = 8;  // batch chunks for the d-theta partial sums

// cos / sin of this CTA's sequence position for every frequency, computed from the learned inv_freq at the top of each
// kernel (<= 64 sincosf per CTA) — the stand-alone table kernel this replaces was 48 extra launches per training step.
// Same arithmetic as the reference: angle = fp32(pos) * inv_freq[j] (torch.outer), full-precision cos / sin (Vi_Tools...:86-91).
constexpr int ROPE_MAX_HALF = 64;
__device__ __forceinline__ const float2* rope_cos_sin(const float* __restrict__ inv_freq, int pos, int half) {
  __shared__ float2 cs_sm[ROPE_MAX_HALF];
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float sn, cn;
    sincosf((float)pos * inv_freq[j], &sn, &cn);
    cs_sm[j] = make_float2(cn, sn);
  }
  __syncthreads();
  return cs_sm;
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
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int half = dr >> 1;
  const int per_head = (dc + half) / V;
  const int pos = blockIdx.x;
  const int bpc = (B + (int)gridDim.y - 1) / (int)gridDim.y;
  const int b0 = blockIdx.y * bpc, b1 = min(B, b0 + bpc);
  const float2* cs2 = rope_cos_sin(cs, pos, half);
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
  pdl_wait(); pdl_launch_small_dependent();   // its dependent is the ~12-CTA d inv_freq reduction
  extern __shared__ float dth_s[];  // heads * half floats (one slot per (head, j): deterministic reduce)
  const int half = dr >> 1;
  const int per_head = (dc + half) / V;
  const int pos = blockIdx.x, chunk = blockIdx.y;
  const int bpc = (B + ROPE_BCH - 1) / ROPE_BCH;
  const int b0 = chunk * bpc, b1 = min(B, b0 + bpc);
  const float2* cs2 = rope_cos_sin(cs, pos, half);
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

// ------------------------------------------------------------------------------------------------------------------
// Row-staged kernels (v3, the default whenever every row starts on a 16-byte boundary): the head width is rarely a multiple
// of 16 bytes (hd = 56, 44, 20 -> rotation halves of 28, 22, 10 elements), which held the kernels above to 8-, 4- or 2-byte
// accesses (30-39 % of the HBM rate at 224^2 / 176^2). Here a thread owns one aligned 16-byte chunk of the ROW, whatever heads
// or halves it straddles: rows go through shared memory so that the rotation partner (half a head away) can be fetched from
// there, all global accesses are 128-bit and fully coalesced. A CTA still owns one position (cos/sin per thread in registers)
// and a slice of the batch; R = 256 / chunks rows are processed side by side, U such groups are in flight per iteration,
// and the staging buffer is double-buffered (one __syncthreads per iteration).
// Staging coordinates of a row: [content row (heads*dc) | rope row (heads*dr)]; out coordinates: h*(dc+dr) + i.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RS_NT = 256;
constexpr int RS_UF = 4;   // forward: row groups in flight
constexpr int RS_UB = 2;   // backward (two operands per row)

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
// staging layout: 4 bytes of padding after every 128 bytes, so that the 32 lanes' scalar partner reads (16 bytes apart) and the
// word-wise chunk writes fall into 32 different banks. sidx(i) = position of row element i.
__host__ __device__ __forceinline__ int sidx(int i) { return i + ((i >> 6) << 1); }
__host__ __device__ __forceinline__ int rope_pitch(int rowlen) { return rowlen + 2 * ((rowlen + 63) / 64); }
__device__ __forceinline__ void sts_chunk(bf16* row, int k, const uint4& v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(row + sidx(8 * k));
  w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
__device__ __forceinline__ uint4 lds_chunk(const bf16* row, int k) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row + sidx(8 * k));
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float lds_bf16(const bf16* p) {
  return __uint_as_float(((uint32_t)*reinterpret_cast<const unsigned short*>(p)) << 16);
}

// per-thread description of the 8 out-coordinate columns [8k, 8k+8): cos, signed sin, own / partner position (sidx applied)
template <bool BWD>
__device__ __forceinline__ void rope_row_map(int k, int heads, int dc, int dr, const float2* __restrict__ cs2, float* c, float* s,
                                             unsigned short* self_stage, unsigned short* partner) {
  const int half = dr >> 1, hd = dc + dr, HC = heads * dc;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = 8 * k + e;
    const int h = col / hd, i = col - h * hd;
    if (i < dc) {
      c[e] = 1.f; s[e] = 0.f;
      self_stage[e] = (unsigned short)sidx(h * dc + i);
      partner[e] = BWD ? (unsigned short)sidx(col) : self_stage[e];
    } else {
      const int j = i - dc;
      const bool first = j < half;
      const float2 v = cs2[first ? j : j - half];
      c[e] = v.x;
      // forward: y1 = x1 c - x2 s, y2 = x2 c + x1 s ; backward: dx1 = dy1 c + dy2 s, dx2 = dy2 c - dy1 s
      s[e] = (first != BWD) ? -v.y : v.y;
      self_stage[e] = (unsigned short)sidx(HC + h * dr + j);
      const int pofs = first ? half : -half;
      partner[e] = (unsigned short)sidx((BWD ? col : HC + h * dr + j) + pofs);   // backward stages the dout row in out coordinates
    }
  }
}

template <bool DC0, bool PAIR>
__global__ void __launch_bounds__(RS_NT, 4)
rope_rows_fwd_kernel(const bf16* __restrict__ content, long long ld_content, const bf16* __restrict__ ropein, long long ld_rope,
                     bf16* __restrict__ out, long long ld_out, const float* __restrict__ cs, int B, int S, int heads, int dc, int dr,
                     int R, int chunks) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  extern __shared__ __align__(16) unsigned char rs_smem[];
  bf16* stage = reinterpret_cast<bf16*>(rs_smem);
  const int rowlen = heads * (dc + dr), HC = heads * dc, half = dr >> 1;
  const int tid = threadIdx.x, r = tid / chunks, k = tid - r * chunks;
  const bool active = r < R;
  const int pos = blockIdx.x;
  const int bpc = (B + (int)gridDim.y - 1) / (int)gridDim.y;
  const int b0 = blockIdx.y * bpc, b1 = min(B, b0 + bpc);
  float c[8], sn[8];
  unsigned short self[8], partner[8];
  rope_row_map<false>(active ? k : 0, heads, dc, dr, rope_cos_sin(cs, pos, half), c, sn, self, partner);
  const bool from_content = 8 * k < HC;
  const bf16* src = from_content ? content + 8 * k : ropein + (8 * k - HC);
  const long long src_ld = from_content ? ld_content : ld_rope;
  const int pitch = rope_pitch(rowlen), set = RS_UF * R * pitch;
  int it = 0;
  for (int bb = b0; bb < b1; bb += RS_UF * R, ++it) {
    bf16* st = stage + (it & 1) * set;
    uint4 v[RS_UF];
    bool ok[RS_UF];
#pragma unroll
    for (int u = 0; u < RS_UF; ++u) {
      const int b = bb + u * R + r;
      ok[u] = active && b < b1;
      if (ok[u]) v[u] = __ldcs(reinterpret_cast<const uint4*>(src + ((long long)b * S + pos) * src_ld));
    }
#pragma unroll
    for (int u = 0; u < RS_UF; ++u)
      if (ok[u]) sts_chunk(st + (u * R + r) * pitch, k, v[u]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < RS_UF; ++u) {
      if (!ok[u]) continue;
      const bf16* row = st + (u * R + r) * pitch;
      float x[8], xp[8], y[8];
      if (DC0) unpack8(v[u], x);
      if (PAIR) {   // even half / content widths: the partners of (e, e+1) are one aligned 32-bit word
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(row + partner[2 * q]);
          xp[2 * q] = __uint_as_float(w << 16); xp[2 * q + 1] = __uint_as_float(w & 0xffff0000u);
          if (!DC0) {
            const uint32_t ws = *reinterpret_cast<const uint32_t*>(row + self[2 * q]);
            x[2 * q] = __uint_as_float(ws << 16); x[2 * q + 1] = __uint_as_float(ws & 0xffff0000u);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (!DC0) x[e] = lds_bf16(row + self[e]);
          xp[e] = lds_bf16(row + partner[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = fmaf(xp[e], sn[e], x[e] * c[e]);
      const int b = bb + u * R + r;
      *reinterpret_cast<uint4*>(out + ((long long)b * S + pos) * ld_out + 8 * k) = pack8(y);
    }
  }
}

template <bool DC0, bool PAIR>
__global__ void __launch_bounds__(RS_NT, 4)
rope_rows_bwd_kernel(const bf16* __restrict__ dout, long long ld_dout, const bf16* __restrict__ out, long long ld_out,
                     bf16* __restrict__ dcontent, long long ld_dcontent, bf16* __restrict__ dropein, long long ld_drope,
                     const float* __restrict__ cs, float* __restrict__ dtheta_part, int B, int S, int heads, int dc, int dr,
                     int R, int chunks) {
  pdl_wait(); pdl_launch_small_dependent();   // its dependent is the ~12-CTA d inv_freq reduction
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int rowlen = heads * (dc + dr), HC = heads * dc, half = dr >> 1, hd = dc + dr;
  const int pitch = rope_pitch(rowlen), set = RS_UB * R * pitch;
  bf16* stage = reinterpret_cast<bf16*>(rs_smem);            // 2 sets: dout rows, out coordinates
  bf16* result = stage + 2 * set;                              // 1 set (dc > 0 only): results, staging coordinates
  float* accs = reinterpret_cast<float*>(result + (DC0 ? 0 : set));   // R * rowlen floats
  const int tid = threadIdx.x, r = tid / chunks, k = tid - r * chunks;
  const bool active = r < R;
  const int pos = blockIdx.x, chunk = blockIdx.y;
  const int bpc = (B + ROPE_BCH - 1) / ROPE_BCH;
  const int b0 = chunk * bpc, b1 = min(B, b0 + bpc);
  float c[8], sn[8], acc[8];
  unsigned short dest[8], partner[8];
  rope_row_map<true>(active ? k : 0, heads, dc, dr, rope_cos_sin(cs, pos, half), c, sn, dest, partner);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const bool to_content = 8 * k < HC;
  bf16* dst = to_content ? dcontent + 8 * k : dropein + (8 * k - HC);
  const long long dst_ld = to_content ? ld_dcontent : ld_drope;
  int it = 0;
  for (int bb = b0; bb < b1; bb += RS_UB * R, ++it) {
    bf16* st = stage + (it & 1) * set;
    uint4 vdy[RS_UB], vy[RS_UB];
    bool ok[RS_UB];
#pragma unroll
    for (int u = 0; u < RS_UB; ++u) {
      const int b = bb + u * R + r;
      ok[u] = active && b < b1;
      if (ok[u]) {
        const long long t = (long long)b * S + pos;
        vdy[u] = __ldcs(reinterpret_cast<const uint4*>(dout + t * ld_dout + 8 * k));
        vy[u] = __ldcs(reinterpret_cast<const uint4*>(out + t * ld_out + 8 * k));
      }
    }
#pragma unroll
    for (int u = 0; u < RS_UB; ++u)
      if (ok[u]) sts_chunk(st + (u * R + r) * pitch, k, vdy[u]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < RS_UB; ++u) {
      if (!ok[u]) continue;
      const bf16* row = st + (u * R + r) * pitch;
      float dy[8], y[8], dx[8], dyp[8];
      unpack8(vdy[u], dy);
      unpack8(vy[u], y);
      if (PAIR) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(row + partner[2 * q]);
          dyp[2 * q] = __uint_as_float(w << 16); dyp[2 * q + 1] = __uint_as_float(w & 0xffff0000u);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) dyp[e] = lds_bf16(row + partner[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dx[e] = fmaf(dyp[e], sn[e], dy[e] * c[e]);
        acc[e] = fmaf(y[e], dyp[e], acc[e]);
      }
      if (DC0) {
        const int b = bb + u * R + r;
        *reinterpret_cast<uint4*>(dropein + ((long long)b * S + pos) * ld_drope + 8 * k) = pack8(dx);
      } else {
        bf16* res = result + (u * R + r) * pitch;
        if (PAIR) {
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<uint32_t*>(res + dest[2 * q]) = pack_bf16x2(dx[2 * q], dx[2 * q + 1]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) res[dest[e]] = __float2bfloat16(dx[e]);
        }
      }
    }
    if (!DC0) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < RS_UB; ++u) {
        if (!ok[u]) continue;
        const int b = bb + u * R + r;
        *reinterpret_cast<uint4*>(dst + ((long long)b * S + pos) * dst_ld) = lds_chunk(result + (u * R + r) * pitch, k);
      }
    }
  }
  // d theta[pos, j] = sum over rows, heads of (y1 dy2 - y2 dy1): per-thread partial sums meet in shared memory (fixed order)
  __syncthreads();
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) accs[r * rowlen + 8 * k + e] = acc[e];
  }
  __syncthreads();
  for (int j = tid; j < half; j += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < R; ++rr)
      for (int h = 0; h < heads; ++h) {
        const float* a = accs + rr * rowlen + h * hd + dc;
        s += a[j] - a[j + half];
      }
    dtheta_part[((size_t)chunk * S + pos) * half + j] = s;
  }
}

// can the row-staged kernels run? every row of every operand must be 16-byte aligned and a row must fit one CTA pass
bool rope_rows_ok(int heads, int dc, int dr, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  const int rowlen = heads * (dc + dr);
  bool ok = (heads * dc) % 8 == 0 && (heads * dr) % 8 == 0 && rowlen / 8 <= RS_NT && rope_pitch(rowlen) <= 65535;
  for (long long ld : lds) ok = ok && ld % 8 == 0;
  for (const void* p : ptrs) ok = ok && (reinterpret_cast<uintptr_t>(p) % 16) == 0;
  return ok;
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
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
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

extern "C" int32_t calm_rope_fwd(const void* content, int64_t ld_content, const void* ropein, int64_t ld_rope, void* out,
                                 int64_t ld_out, const float* cos_sin, int64_t tokens, int32_t S, int32_t heads, int32_t dc,
                                 int32_t dr, cudaStream_t stream) {
  CALM_CHECK_ARG(tokens > 0 && S > 0 && heads > 0 && dr > 0 && dr % 2 == 0 && dc >= 0, "calm_rope_fwd: bad dims");
  CALM_CHECK_ARG(dr / 2 <= ROPE_MAX_HALF, "calm_rope_fwd: rope width %d > %d", dr, 2 * ROPE_MAX_HALF);
  CALM_CHECK_ARG(dc == 0 || content != nullptr, "calm_rope_fwd: content missing");
  CALM_CHECK_ARG(tokens % S == 0, "calm_rope_fwd: tokens=%lld is not a multiple of S=%d", (long long)tokens, S);
  const int B = (int)(tokens / S);
  dim3 grid(S, B < ROPE_BCH ? B : ROPE_BCH);
  if (rope_rows_ok(heads, dc, dr, {(long long)(dc ? ld_content : 8), (long long)ld_rope, (long long)ld_out}, {content, ropein, out})) {
    const int rowlen = heads * (dc + dr), chunks = rowlen / 8, R = RS_NT / chunks;
    const int nthreads = ((R * chunks + 31) / 32) * 32;
    const size_t smem = (size_t)2 * RS_UF * R * rope_pitch(rowlen) * sizeof(bf16);
    if (smem + sizeof(float2) * ROPE_MAX_HALF <= 48 * 1024) {
      const bool pair = (dr / 2) % 2 == 0 && dc % 2 == 0;
#define CALM_ROPE_ROWS_FWD(DC0, PAIR)                                                                                            \
  do {                                                                                                                           \
    static CalmDeviceOnce carve;                                                                                                   \
    if (carve.pending()) { cudaFuncSetAttribute(rope_rows_fwd_kernel<DC0, PAIR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); carve.done(); } \
    CALM_LAUNCH((rope_rows_fwd_kernel<DC0, PAIR>), grid, nthreads, smem, stream, reinterpret_cast<const bf16*>(content), ld_content,        \
        reinterpret_cast<const bf16*>(ropein), ld_rope, reinterpret_cast<bf16*>(out), ld_out, cos_sin, B, S, heads, dc, dr, R, chunks); \
  } while (0)
      if (dc == 0) { if (pair) CALM_ROPE_ROWS_FWD(true, true); else CALM_ROPE_ROWS_FWD(true, false); }
      else { if (pair) CALM_ROPE_ROWS_FWD(false, true); else CALM_ROPE_ROWS_FWD(false, false); }
#undef CALM_ROPE_ROWS_FWD
      CALM_CHECK_LAUNCH("calm_rope_fwd(rows)");
      return CALM_OK;
    }
  }
  const int v = rope_vec(dc, dr, {(long long)ld_content, (long long)ld_rope, (long long)ld_out}, {content, ropein, out});
  const int threads = rope_threads(heads, dc, dr, v);
#define CALM_ROPE_FWD(V)                                                                                                          \
  CALM_LAUNCH((rope_fwd_kernel<V>), grid, threads, 0, stream, reinterpret_cast<const bf16*>(content), ld_content,                            \
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
  CALM_CHECK_ARG(dr / 2 <= ROPE_MAX_HALF, "calm_rope_bwd: rope width %d > %d", dr, 2 * ROPE_MAX_HALF);
  CALM_CHECK_ARG(dc == 0 || dcontent != nullptr, "calm_rope_bwd: dcontent missing");
  const int B = (int)(tokens / S), half = dr / 2;
  dim3 grid(S, ROPE_BCH);
  if (rope_rows_ok(heads, dc, dr, {(long long)ld_dout, (long long)ld_out, (long long)(dc ? ld_dcontent : 8), (long long)ld_drope},
                   {dout, out, dcontent, dropein})) {
    const int rowlen = heads * (dc + dr), chunks = rowlen / 8, R = RS_NT / chunks;
    const int nthreads = ((R * chunks + 31) / 32) * 32;
    const size_t rs_smem = (size_t)(dc ? 3 : 2) * RS_UB * R * rope_pitch(rowlen) * sizeof(bf16) + (size_t)R * rowlen * sizeof(float);
    if (rs_smem + sizeof(float2) * ROPE_MAX_HALF <= 48 * 1024) {
      const bool pair = half % 2 == 0 && dc % 2 == 0;
#define CALM_ROPE_ROWS_BWD(DC0, PAIR)                                                                                            \
  do {                                                                                                                           \
    static CalmDeviceOnce carve;                                                                                                   \
    if (carve.pending()) { cudaFuncSetAttribute(rope_rows_bwd_kernel<DC0, PAIR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); carve.done(); } \
    CALM_LAUNCH((rope_rows_bwd_kernel<DC0, PAIR>), grid, nthreads, rs_smem, stream,                                                         \
        reinterpret_cast<const bf16*>(dout), ld_dout, reinterpret_cast<const bf16*>(out), ld_out, reinterpret_cast<bf16*>(dcontent), \
        ld_dcontent, reinterpret_cast<bf16*>(dropein), ld_drope, cos_sin, dtheta_part, B, S, heads, dc, dr, R, chunks);          \
  } while (0)
      if (dc == 0) { if (pair) CALM_ROPE_ROWS_BWD(true, true); else CALM_ROPE_ROWS_BWD(true, false); }
      else { if (pair) CALM_ROPE_ROWS_BWD(false, true); else CALM_ROPE_ROWS_BWD(false, false); }
#undef CALM_ROPE_ROWS_BWD
      CALM_CHECK_LAUNCH("calm_rope_bwd(rows)");
      CALM_LAUNCH((rope_dfreq_kernel), half, 128, 0, stream, dtheta_part, dinv_freq, S, half);
      CALM_CHECK_LAUNCH("calm_rope_bwd(dfreq)");
      return CALM_OK;
    }
  }
  const int v = rope_vec(dc, dr, {(long long)ld_dout, (long long)ld_out, (long long)ld_dcontent, (long long)ld_drope},
                         {dout, out, dcontent, dropein});
  const int threads = rope_threads(heads, dc, dr, v);
  const size_t smem = (size_t)heads * half * sizeof(float);
#define CALM_ROPE_BWD(V)                                                                                                       \
  CALM_LAUNCH((rope_bwd_kernel<V>), grid, threads, smem, stream, reinterpret_cast<const bf16*>(dout), ld_dout,                            \
                                                      reinterpret_cast<const bf16*>(out), ld_out, reinterpret_cast<bf16*>(dcontent), \
                                                      ld_dcontent, reinterpret_cast<bf16*>(dropein), ld_drope, cos_sin, dtheta_part, \
                                                      B, S, heads, dc, dr)
  if (v == 8) CALM_ROPE_BWD(8); else if (v == 4) CALM_ROPE_BWD(4); else if (v == 2) CALM_ROPE_BWD(2); else CALM_ROPE_BWD(1);
#undef CALM_ROPE_BWD
  CALM_CHECK_LAUNCH("calm_rope_bwd");
  CALM_LAUNCH((rope_dfreq_kernel), half, 128, 0, stream, dtheta_part, dinv_freq, S, half);
  CALM_CHECK_LAUNCH("calm_rope_bwd(dfreq)");
  return CALM_OK;
}
