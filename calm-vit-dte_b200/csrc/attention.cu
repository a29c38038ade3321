// Fused axial attention with a learned additive bias shared by all heads (forward + backward).
//   O = softmax(Q K^T / sqrt(hd) + bias[b]) V          replaces F.scaled_dot_product_attention(q,k,v,attn_mask=...)
//   (Vi_Tools_CNN_less_V2.py:293-298), the bias being linear_mask(q_mask @ k_mask^T).unsqueeze(1) (:288-291).
// Each image is one sequence of S row- (or column-) tokens per head (S = 80..224 at 224^2, <= 512 at 512^2); scores never
// touch HBM. Backward: dV = P^T dO, dP = dO V^T, dS = P (.) (dP - rowsum(dO (.) O)),
// dQ = dS K / sqrt(hd), dK = dS^T Q / sqrt(hd), dbias[b] = sum_h dS_h  — split into a query-major kernel (dQ) and a
// key-major kernel (dK, dV, dbias accumulated over heads in shared memory) so that nothing needs atomics and the result
// is deterministic.
// Round-1 implementation: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate), 64x64 tiles, online softmax in
// registers with warp shuffles. Q/K/V are read straight from the token-major GEMM outputs (B*S, heads*hd) — head dims
// that are not multiples of 16 (56, 44, 20, ...) are zero-padded in shared memory only.
#include "common.cuh"
#include "../../include/calm_b200.h"
#include "attention_tc.h"
#include <math.h>

namespace {

constexpr int TILE = 64;            // query rows / keys per tile
constexpr int ATT_THREADS = 128;    // 4 warps x 16 rows
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Load a TILE x hd slab (rows row0.. of a token-major matrix, head offset already applied) into smem [TILE][P],
// zero-filling rows >= nvalid and columns >= hd. 8-byte (4 x bf16) accesses: hd % 4 == 0 and ld % 4 == 0 are required.
template <int HDP>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* __restrict__ src, long long ld, int nvalid, int hd) {
  constexpr int P = HDP + 8;
  constexpr int V = HDP / 4;  // uint2 per padded row
  const int hv = hd >> 2;
#pragma unroll 4
  for (int it = 0; it < (TILE * V + ATT_THREADS - 1) / ATT_THREADS; ++it) {
    const int idx = threadIdx.x + it * ATT_THREADS;
    if (idx < TILE * V) {
      const int r = idx / V, c = idx - r * V;
      uint2 val = make_uint2(0u, 0u);
      if (r < nvalid && c < hv) val = *reinterpret_cast<const uint2*>(src + (long long)r * ld + 4 * c);
      *reinterpret_cast<uint2*>(dst + r * P + 4 * c) = val;
    }
  }
}

// Load the 64 x 64 block (rows q0.., columns k0..) of one image's (S, S) bias matrix into smem [TILE][BP], zero-filled.
constexpr int BP = TILE + 8;
__device__ __forceinline__ void load_bias_tile(bf16* dst, const bf16* __restrict__ bias_b, int S, int q0, int k0) {
#pragma unroll 4
  for (int it = 0; it < (TILE * (TILE / 4)) / ATT_THREADS; ++it) {
    const int idx = threadIdx.x + it * ATT_THREADS;
    const int r = idx >> 4, c = idx & 15;
    uint2 val = make_uint2(0u, 0u);
    if (q0 + r < S && k0 + 4 * c < S) val = *reinterpret_cast<const uint2*>(bias_b + (long long)(q0 + r) * S + k0 + 4 * c);
    *reinterpret_cast<uint2*>(dst + r * BP + 4 * c) = val;
  }
}

// A-operand fragments (16 rows x HDP) of this warp's 16-row slab
template <int HDP>
__device__ __forceinline__ void load_a_frags(uint32_t (*f)[4], const bf16* tile, int row0, int lane) {
  constexpr int P = HDP + 8;
#pragma unroll
  for (int ks = 0; ks < HDP / 16; ++ks)
    ldsm_x4(sm_addr(tile + (row0 + (lane & 15)) * P + ks * 16 + (lane >> 4) * 8), f[ks][0], f[ks][1], f[ks][2], f[ks][3]);
}

// C[16 x 64] += A(16 x HDP) . T^T where T is a [64][P] smem tile (rows = output columns), i.e. "A . T^T"
template <int HDP>
__device__ __forceinline__ void mma_a_tT(float (*c)[4], const uint32_t (*af)[4], const bf16* tile, int lane) {
  constexpr int P = HDP + 8;
#pragma unroll
  for (int ks = 0; ks < HDP / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sm_addr(tile + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * P + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
      mma16816(c[2 * np], af[ks], b0, b1);
      mma16816(c[2 * np + 1], af[ks], b2, b3);
    }
  }
}

// C[16 x HDP] += Pm(16 x 64, from accumulator registers) . T where T is a [64][P] smem tile (rows = contraction index)
template <int HDP>
__device__ __forceinline__ void mma_p_t(float (*c)[4], const float (*pm)[4], const bf16* tile, int lane) {
  constexpr int P = HDP + 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(pm[2 * kk][0], pm[2 * kk][1]);
    a[1] = pack_bf16x2(pm[2 * kk][2], pm[2 * kk][3]);
    a[2] = pack_bf16x2(pm[2 * kk + 1][0], pm[2 * kk + 1][1]);
    a[3] = pack_bf16x2(pm[2 * kk + 1][2], pm[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < HDP / 16; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(sm_addr(tile + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * P + np * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
      mma16816(c[2 * np], a, b0, b1);
      mma16816(c[2 * np + 1], a, b2, b3);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: grid (ceil(S/64), heads, B)
// ---------------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const bf16* __restrict__ bias,
                bf16* __restrict__ o, float* __restrict__ lse, long long ld_q, long long ld_k, long long ld_v, long long ld_o,
                int S, int heads, int hd, float scale_log2) {
  constexpr int P = HDP + 8, NT = HDP / 8;
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* Qs = reinterpret_cast<bf16*>(att_smem);
  bf16* Ks = Qs + TILE * P;
  bf16* Vs = Ks + TILE * P;
  bf16* Bs = Vs + TILE * P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const long long tok0 = (long long)b * S;

  load_tile<HDP>(Qs, q + (tok0 + q0) * ld_q + (long long)h * hd, ld_q, S - q0, hd);
  __syncthreads();
  uint32_t qf[HDP / 16][4];
  load_a_frags<HDP>(qf, Qs, warp * 16, lane);

  float oacc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f; }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  const int row_lo = q0 + warp * 16 + (lane >> 2);  // this thread's rows: row_lo and row_lo + 8
  const bf16* bias_b = bias + (long long)b * S * S;

  for (int kb0 = 0; kb0 < S; kb0 += TILE) {
    __syncthreads();
    load_tile<HDP>(Ks, k + (tok0 + kb0) * ld_k + (long long)h * hd, ld_k, S - kb0, hd);
    load_tile<HDP>(Vs, v + (tok0 + kb0) * ld_v + (long long)h * hd, ld_v, S - kb0, hd);
    load_bias_tile(Bs, bias_b, S, q0, kb0);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    mma_a_tT<HDP>(s, qf, Ks, lane);
    // scale, bias, key mask
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kb0 + nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float2 bb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Bs + (warp * 16 + (lane >> 2) + hh * 8) * BP + nt * 8 + (lane & 3) * 2));
        const float b0 = bb.x, b1 = bb.y;
        float v0 = s[nt][2 * hh] * scale_log2 + b0 * LOG2E;
        float v1 = s[nt][2 * hh + 1] * scale_log2 + b1 * LOG2E;
        if (key >= S) v0 = -INFINITY;
        if (key + 1 >= S) v1 = -INFINITY;
        s[nt][2 * hh] = v0; s[nt][2 * hh + 1] = v1;
        mx[hh] = fmaxf(mx[hh], fmaxf(v0, v1));
      }
    }
    float alpha[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      const float mnew = fmaxf(m[hh], mx[hh]);
      alpha[hh] = exp2f(m[hh] - mnew);
      m[hh] = mnew;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float p0 = exp2f(s[nt][2 * hh] - m[hh]), p1 = exp2f(s[nt][2 * hh + 1] - m[hh]);
        s[nt][2 * hh] = p0; s[nt][2 * hh + 1] = p1;
        rs[hh] += p0 + p1;
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 1);
      rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 2);
      l[hh] = l[hh] * alpha[hh] + rs[hh];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      oacc[i][0] *= alpha[0]; oacc[i][1] *= alpha[0]; oacc[i][2] *= alpha[1]; oacc[i][3] *= alpha[1];
    }
    mma_p_t<HDP>(oacc, s, Vs, lane);
  }

#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int row = row_lo + hh * 8;
    if (row < S) {
      const float inv = 1.0f / l[hh];
      bf16* op = o + (tok0 + row) * ld_o + (long long)h * hd;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + (lane & 3) * 2;
        if (c < hd) *reinterpret_cast<uint32_t*>(op + c) = pack_bf16x2(oacc[nt][2 * hh] * inv, oacc[nt][2 * hh + 1] * inv);
      }
      if ((lane & 3) == 0) lse[((long long)b * heads + h) * S + row] = (m[hh] + log2f(l[hh])) * LN2;
    }
  }
}

// delta[b,h,s] = sum_c dO[t, h*hd + c] * O[t, h*hd + c]   (one thread per (token, head): adjacent threads read adjacent
// 2*hd-byte segments of the same token row, so the warp's accesses are contiguous)
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta, long long ld_o,
                                  long long ld_do, long long tokens, int S, int heads, int hd) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= tokens * heads) return;
  const long long t = w / heads;
  const int h = (int)(w - t * heads);
  const uint2* op = reinterpret_cast<const uint2*>(o + t * ld_o + (long long)h * hd);
  const uint2* dp = reinterpret_cast<const uint2*>(d_o + t * ld_do + (long long)h * hd);
  float acc = 0.f;
  for (int c = 0; c < (hd >> 2); ++c) {
    const uint2 a = op[c], g = dp[c];
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y);
    acc += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y;
  }
  delta[((t / S) * heads + h) * S + (t % S)] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// backward, query-major: dQ.  grid (ceil(S/64), heads, B)
// ---------------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS, 3)
attn_bwd_dq_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const bf16* __restrict__ bias,
                   const bf16* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dq,
                   bf16* __restrict__ ds_out /* optional (B, heads, S, S) bf16: this head's dS, summed over heads by dbias_reduce */,
                   long long ld_q, long long ld_k, long long ld_v, long long ld_do, long long ld_dq, int S, int heads, int hd,
                   float scale, float scale_log2) {
  constexpr int P = HDP + 8, NT = HDP / 8;
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* Qs = reinterpret_cast<bf16*>(att_smem);
  bf16* dOs = Qs + TILE * P;
  bf16* Ks = dOs + TILE * P;
  bf16* Vs = Ks + TILE * P;
  bf16* Bs = Vs + TILE * P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const long long tok0 = (long long)b * S;

  load_tile<HDP>(Qs, q + (tok0 + q0) * ld_q + (long long)h * hd, ld_q, S - q0, hd);
  load_tile<HDP>(dOs, d_o + (tok0 + q0) * ld_do + (long long)h * hd, ld_do, S - q0, hd);
  __syncthreads();
  uint32_t qf[HDP / 16][4], dof[HDP / 16][4];
  load_a_frags<HDP>(qf, Qs, warp * 16, lane);
  load_a_frags<HDP>(dof, dOs, warp * 16, lane);

  const int row_lo = q0 + warp * 16 + (lane >> 2);
  float lse2[2], dlt[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int row = row_lo + hh * 8;
    const long long idx = ((long long)b * heads + h) * S + row;
    lse2[hh] = row < S ? lse[idx] * LOG2E : 0.f;
    dlt[hh] = row < S ? delta[idx] : 0.f;
  }
  float dqacc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { dqacc[i][0] = dqacc[i][1] = dqacc[i][2] = dqacc[i][3] = 0.f; }
  const bf16* bias_b = bias + (long long)b * S * S;

  for (int kb0 = 0; kb0 < S; kb0 += TILE) {
    __syncthreads();
    load_tile<HDP>(Ks, k + (tok0 + kb0) * ld_k + (long long)h * hd, ld_k, S - kb0, hd);
    load_tile<HDP>(Vs, v + (tok0 + kb0) * ld_v + (long long)h * hd, ld_v, S - kb0, hd);
    load_bias_tile(Bs, bias_b, S, q0, kb0);
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_a_tT<HDP>(s, qf, Ks, lane);
    mma_a_tT<HDP>(dp, dof, Vs, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kb0 + nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = row_lo + hh * 8;
        const bool ok = row < S && key < S;
        const float2 bb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Bs + (warp * 16 + (lane >> 2) + hh * 8) * BP + nt * 8 + (lane & 3) * 2));
        const float b0 = bb.x, b1 = bb.y;
        const float p0 = ok ? exp2f(s[nt][2 * hh] * scale_log2 + b0 * LOG2E - lse2[hh]) : 0.f;
        const float p1 = (ok && key + 1 < S) ? exp2f(s[nt][2 * hh + 1] * scale_log2 + b1 * LOG2E - lse2[hh]) : 0.f;
        s[nt][2 * hh] = p0 * (dp[nt][2 * hh] - dlt[hh]);          // dS
        s[nt][2 * hh + 1] = p1 * (dp[nt][2 * hh + 1] - dlt[hh]);
      }
    }
    if (ds_out) {
      // this warp's 16 x 64 dS block goes through its own rows of the (consumed) bias tile and leaves as 16-byte row pieces
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          *reinterpret_cast<uint32_t*>(Bs + (warp * 16 + (lane >> 2) + hh * 8) * BP + nt * 8 + (lane & 3) * 2) = pack_bf16x2(s[nt][2 * hh], s[nt][2 * hh + 1]);
      __syncwarp();
      bf16* dsb = ds_out + ((long long)b * heads + h) * S * S;
#pragma unroll
      for (int i = lane; i < 128; i += 32) {
        const int rr = i >> 3, cc = i & 7;
        const int row = q0 + warp * 16 + rr, key = kb0 + 8 * cc;
        if (row < S && key < S)
          *reinterpret_cast<uint4*>(dsb + (long long)row * S + key) = *reinterpret_cast<const uint4*>(Bs + (warp * 16 + rr) * BP + 8 * cc);
      }
    }
    mma_p_t<HDP>(dqacc, s, Ks, lane);
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int row = row_lo + hh * 8;
    if (row < S) {
      bf16* op = dq + (tok0 + row) * ld_dq + (long long)h * hd;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + (lane & 3) * 2;
        if (c < hd) *reinterpret_cast<uint32_t*>(op + c) = pack_bf16x2(dqacc[nt][2 * hh] * scale, dqacc[nt][2 * hh + 1] * scale);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, key-major: dK, dV and dbias = sum over heads of dS.  grid (ceil(S/64), B); loops heads x query tiles.
// Works on transposed tiles (keys x queries) so that P^T / dS^T feed the next MMA straight from registers.
// ---------------------------------------------------------------------------------------------------------------
// PER_HEAD: grid (key tiles, heads, B), no dbias (the dq kernel wrote every head's dS to the scratch, dbias_reduce sums them) —
// 12x the parallelism of the all-heads-in-one-CTA form, whose S x 65 fp32 dbias tile also held it to one 4-warp CTA per SM
// (measured at 384^2 / 512^2: 69 / 117 ms per step in this kernel, 1.4 % of the tensor peak).
template <int HDP, bool PER_HEAD>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_bwd_dkv_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const bf16* __restrict__ bias,
                    const bf16* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dk,
                    bf16* __restrict__ dv, bf16* __restrict__ dbias, long long ld_q, long long ld_k, long long ld_v, long long ld_do,
                    long long ld_dk, long long ld_dv, int S, int heads, int hd, float scale, float scale_log2, int s_pad) {
  constexpr int P = HDP + 8, NT = HDP / 8;
  constexpr int DBP = TILE + 1;  // dbias smem pitch (floats)
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* Ks = reinterpret_cast<bf16*>(att_smem);
  bf16* Vs = Ks + TILE * P;
  bf16* Qs = Vs + TILE * P;
  bf16* dOs = Qs + TILE * P;
  bf16* Bs = dOs + TILE * P;                                 // TILE x BP bias block [query][key]
  float* lse_s = reinterpret_cast<float*>(Bs + TILE * BP);   // TILE
  float* dlt_s = lse_s + TILE;                               // TILE
  float* db_s = dlt_s + TILE;                                // s_pad x DBP   [query][key]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k0 = blockIdx.x * TILE, b = PER_HEAD ? blockIdx.z : blockIdx.y;
  const long long tok0 = (long long)b * S;
  const bf16* bias_b = bias + (long long)b * S * S;
  if (!PER_HEAD)
    for (int i = threadIdx.x; i < s_pad * DBP; i += ATT_THREADS) db_s[i] = 0.f;
  const int key_lo = k0 + warp * 16 + (lane >> 2);  // this thread's keys: key_lo, key_lo + 8
  const int h_begin = PER_HEAD ? blockIdx.y : 0, h_end = PER_HEAD ? blockIdx.y + 1 : heads;

  for (int h = h_begin; h < h_end; ++h) {
    __syncthreads();
    load_tile<HDP>(Ks, k + (tok0 + k0) * ld_k + (long long)h * hd, ld_k, S - k0, hd);
    load_tile<HDP>(Vs, v + (tok0 + k0) * ld_v + (long long)h * hd, ld_v, S - k0, hd);
    __syncthreads();
    uint32_t kf[HDP / 16][4], vf[HDP / 16][4];
    load_a_frags<HDP>(kf, Ks, warp * 16, lane);
    load_a_frags<HDP>(vf, Vs, warp * 16, lane);
    float dkacc[NT][4], dvacc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      dkacc[i][0] = dkacc[i][1] = dkacc[i][2] = dkacc[i][3] = 0.f;
      dvacc[i][0] = dvacc[i][1] = dvacc[i][2] = dvacc[i][3] = 0.f;
    }
    for (int q0 = 0; q0 < S; q0 += TILE) {
      __syncthreads();
      load_tile<HDP>(Qs, q + (tok0 + q0) * ld_q + (long long)h * hd, ld_q, S - q0, hd);
      load_tile<HDP>(dOs, d_o + (tok0 + q0) * ld_do + (long long)h * hd, ld_do, S - q0, hd);
      load_bias_tile(Bs, bias_b, S, q0, k0);
      if (threadIdx.x < TILE) {
        const int row = q0 + threadIdx.x;
        const long long idx = ((long long)b * heads + h) * S + row;
        lse_s[threadIdx.x] = row < S ? lse[idx] * LOG2E : 0.f;
        dlt_s[threadIdx.x] = row < S ? delta[idx] : 0.f;
      }
      __syncthreads();
      float st[8][4], dpt[8][4];  // [keys(16) x queries(64)] transposed tiles
#pragma unroll
      for (int i = 0; i < 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
      mma_a_tT<HDP>(st, kf, Qs, lane);    // S^T = K Q^T
      mma_a_tT<HDP>(dpt, vf, dOs, lane);  // dP^T = V dO^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key_lo + (e >> 1) * 8;
          const int ql = nt * 8 + (lane & 3) * 2 + (e & 1);
          const int qrow = q0 + ql;
          float p = 0.f;
          if (key < S && qrow < S) {
            const float bb = __bfloat162float(Bs[ql * BP + (key - k0)]);
            p = exp2f(st[nt][e] * scale_log2 + bb * LOG2E - lse_s[ql]);
          }
          st[nt][e] = p;                                   // P^T
          const float ds = p * (dpt[nt][e] - dlt_s[ql]);
          dpt[nt][e] = ds;                                 // dS^T
          if (!PER_HEAD && key < S && qrow < S) db_s[qrow * DBP + (key - k0)] += ds;  // each (q,key) owned by exactly one thread
        }
      }
      mma_p_t<HDP>(dvacc, st, dOs, lane);  // dV += P^T dO
      mma_p_t<HDP>(dkacc, dpt, Qs, lane);  // dK += dS^T Q
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int key = key_lo + hh * 8;
      if (key < S) {
        bf16* dkp = dk + (tok0 + key) * ld_dk + (long long)h * hd;
        bf16* dvp = dv + (tok0 + key) * ld_dv + (long long)h * hd;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = nt * 8 + (lane & 3) * 2;
          if (c < hd) {
            *reinterpret_cast<uint32_t*>(dkp + c) = pack_bf16x2(dkacc[nt][2 * hh] * scale, dkacc[nt][2 * hh + 1] * scale);
            *reinterpret_cast<uint32_t*>(dvp + c) = pack_bf16x2(dvacc[nt][2 * hh], dvacc[nt][2 * hh + 1]);
          }
        }
      }
    }
  }
  if (PER_HEAD) return;
  __syncthreads();
  bf16* db_b = dbias + (long long)b * S * S;
  const int nk = min(TILE, S - k0);
  for (int i = threadIdx.x; i < S * TILE; i += ATT_THREADS) {
    const int qrow = i / TILE, kk = i - qrow * TILE;
    if (kk < nk) db_b[(long long)qrow * S + k0 + kk] = __float2bfloat16(db_s[qrow * DBP + kk]);
  }
}

template <int HDP> size_t fwd_smem() { return (size_t)(3 * TILE * (HDP + 8) + TILE * BP) * sizeof(bf16); }
template <int HDP> size_t dq_smem() { return (size_t)(4 * TILE * (HDP + 8) + TILE * BP) * sizeof(bf16); }
template <int HDP> size_t dkv_smem(int s_pad) {
  return (size_t)(4 * TILE * (HDP + 8) + TILE * BP) * sizeof(bf16) + 2 * TILE * sizeof(float) + (size_t)s_pad * (TILE + 1) * sizeof(float);
}

template <typename K>
int ensure_smem(K kernel, size_t smem, const char* name) {
  if (smem <= 48 * 1024) return CALM_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { calm_set_error("%s: smem %zu: %s", name, smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
  return CALM_OK;
}

template <int HDP>
int launch_fwd(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
               int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream) {
  const size_t smem = fwd_smem<HDP>();
  int rc = ensure_smem(attn_fwd_kernel<HDP>, smem, "calm_attention_fwd");
  if (rc) return rc;
  const float scale = 1.0f / sqrtf((float)hd);
  dim3 grid((S + TILE - 1) / TILE, heads, B);
  attn_fwd_kernel<HDP><<<grid, ATT_THREADS, smem, stream>>>(
      reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v),
      reinterpret_cast<const bf16*>(bias), reinterpret_cast<bf16*>(o), lse, ld_q, ld_k, ld_v, ld_o, S, heads, hd, scale * LOG2E);
  CALM_CHECK_LAUNCH("calm_attention_fwd");
  return CALM_OK;
}

template <int HDP>
int launch_bwd(const void* q, const void* k, const void* v, const void* bias, const void* d_o, const float* lse, const float* delta,
               void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q, int64_t ld_k, int64_t ld_v, int64_t ld_do,
               int64_t ld_dq, int64_t ld_dk, int64_t ld_dv, int B, int S, int heads, int hd, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)hd);
  // with a scratch (and 16-byte rows: S % 8 == 0) every head writes its dS once and dK / dV run one CTA per (key tile, head, image)
  const bool per_head = ds_scratch != nullptr && S % 8 == 0;
  {
    const size_t smem = dq_smem<HDP>();
    int rc = ensure_smem(attn_bwd_dq_kernel<HDP>, smem, "calm_attention_bwd(dq)");
    if (rc) return rc;
    dim3 grid((S + TILE - 1) / TILE, heads, B);
    attn_bwd_dq_kernel<HDP><<<grid, ATT_THREADS, smem, stream>>>(
        reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v),
        reinterpret_cast<const bf16*>(bias), reinterpret_cast<const bf16*>(d_o), lse, delta, reinterpret_cast<bf16*>(dq),
        per_head ? reinterpret_cast<bf16*>(ds_scratch) : nullptr, ld_q, ld_k, ld_v, ld_do, ld_dq, S, heads, hd, scale, scale * LOG2E);
    CALM_CHECK_LAUNCH("calm_attention_bwd(dq)");
  }
  if (per_head) {
    const size_t smem = dkv_smem<HDP>(0);
    int rc = ensure_smem(attn_bwd_dkv_kernel<HDP, true>, smem, "calm_attention_bwd(dkv)");
    if (rc) return rc;
    dim3 grid((S + TILE - 1) / TILE, heads, B);
    attn_bwd_dkv_kernel<HDP, true><<<grid, ATT_THREADS, smem, stream>>>(
        reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v),
        reinterpret_cast<const bf16*>(bias), reinterpret_cast<const bf16*>(d_o), lse, delta, reinterpret_cast<bf16*>(dk),
        reinterpret_cast<bf16*>(dv), nullptr, ld_q, ld_k, ld_v, ld_do, ld_dk, ld_dv, S, heads, hd, scale, scale * LOG2E, 0);
    CALM_CHECK_LAUNCH("calm_attention_bwd(dkv)");
    return calm_attention_dbias_reduce(ds_scratch, dbias, B, S, heads, stream);
  }
  {
    const int s_pad = ((S + TILE - 1) / TILE) * TILE;
    const size_t smem = dkv_smem<HDP>(s_pad);
    if (smem > 227 * 1024) { calm_set_error("calm_attention_bwd: S=%d hd=%d needs %zu B smem", S, hd, smem); return CALM_ERR_UNSUPPORTED; }
    int rc = ensure_smem(attn_bwd_dkv_kernel<HDP, false>, smem, "calm_attention_bwd(dkv)");
    if (rc) return rc;
    dim3 grid((S + TILE - 1) / TILE, B);
    attn_bwd_dkv_kernel<HDP, false><<<grid, ATT_THREADS, smem, stream>>>(
        reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v),
        reinterpret_cast<const bf16*>(bias), reinterpret_cast<const bf16*>(d_o), lse, delta, reinterpret_cast<bf16*>(dk),
        reinterpret_cast<bf16*>(dv), reinterpret_cast<bf16*>(dbias), ld_q, ld_k, ld_v, ld_do, ld_dk, ld_dv, S, heads, hd, scale,
        scale * LOG2E, s_pad);
    CALM_CHECK_LAUNCH("calm_attention_bwd(dkv)");
  }
  return CALM_OK;
}

int check_common(const char* name, int B, int S, int heads, int hd) {
  CALM_CHECK_ARG(B > 0 && S > 0 && heads > 0 && hd > 0, "%s: empty problem", name);
  CALM_CHECK_ARG(hd % 4 == 0 && hd <= 128, "%s: head_dim=%d must be a multiple of 4 and <= 128", name, hd);
  CALM_CHECK_ARG(S % 4 == 0, "%s: S=%d must be a multiple of 4", name, S);
  return CALM_OK;
}

}  // namespace

#define DISPATCH_HDP(hd, CALL)                         \
  do {                                                 \
    if ((hd) <= 16) { constexpr int HDP = 16; CALL; }  \
    else if ((hd) <= 32) { constexpr int HDP = 32; CALL; }  \
    else if ((hd) <= 48) { constexpr int HDP = 48; CALL; }  \
    else if ((hd) <= 64) { constexpr int HDP = 64; CALL; }  \
    else if ((hd) <= 96) { constexpr int HDP = 96; CALL; }  \
    else { constexpr int HDP = 128; CALL; }            \
  } while (0)

extern "C" int32_t calm_attention_fwd(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse,
                                      int64_t ld_q, int64_t ld_k, int64_t ld_v, int64_t ld_o, int32_t B, int32_t S,
                                      int32_t heads, int32_t hd, cudaStream_t stream) {
  int rc = check_common("calm_attention_fwd", B, S, heads, hd);
  if (rc) return rc;
  CALM_CHECK_ARG(ld_q % 2 == 0 && ld_k % 2 == 0 && ld_v % 2 == 0 && ld_o % 2 == 0, "calm_attention_fwd: leading dims must be even");
  CALM_CHECK_ARG(ld_q % 4 == 0 && ld_k % 4 == 0 && ld_v % 4 == 0, "calm_attention_fwd: q/k/v leading dims must be multiples of 4");
  {
    const int64_t lds[3] = {ld_q, ld_k, ld_v};
    const void* ptrs[4] = {q, k, v, bias};
    {
      const void* sp[5] = {q, k, v, bias, o};
      if (calm_attention_small_eligible(B, S, heads, hd, lds, 3, sp, 5))     // short rows, narrow heads: a warp per item, scores in registers
        return calm_attention_fwd_small(q, k, v, bias, o, lse, ld_q, ld_k, ld_v, ld_o, B, S, heads, hd, stream);
    }
#ifndef CALM_FORCE_LONG_ATTENTION     // tuning builds only: route every eligible shape through the chunked kernels
    if (calm_attention_tc_eligible(B, S, heads, hd, lds, 3, ptrs, 4))
#else
    if (false)
#endif
      return calm_attention_fwd_tc(q, k, v, bias, o, lse, ld_q, ld_k, ld_v, ld_o, B, S, heads, hd, stream);
    if (ld_o % 4 == 0 && calm_attention_long_eligible(B, S, heads, hd, lds, 3, ptrs, 4))
      return calm_attention_fwd_long(q, k, v, bias, o, lse, ld_q, ld_k, ld_v, ld_o, B, S, heads, hd, stream);
  }
  DISPATCH_HDP(hd, return launch_fwd<HDP>(q, k, v, bias, o, lse, ld_q, ld_k, ld_v, ld_o, B, S, heads, hd, stream));
  return CALM_OK;
}

extern "C" int64_t calm_attention_bwd_scratch_bytes(int32_t B, int32_t S, int32_t heads, int32_t hd) {
  (void)hd;
  return (int64_t)calm_attention_bwd_tc_scratch_bytes(B, S, heads);
}

extern "C" int32_t calm_attention_bwd(const void* q, const void* k, const void* v, const void* bias, const void* o, const void* d_o,
                                      const float* lse, float* delta, void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q,
                                      int64_t ld_k, int64_t ld_v, int64_t ld_o, int64_t ld_do, int64_t ld_dq, int64_t ld_dk,
                                      int64_t ld_dv, int32_t B, int32_t S, int32_t heads, int32_t hd, cudaStream_t stream) {
  int rc = check_common("calm_attention_bwd", B, S, heads, hd);
  if (rc) return rc;
  CALM_CHECK_ARG(ld_q % 2 == 0 && ld_k % 2 == 0 && ld_v % 2 == 0 && ld_o % 2 == 0 && ld_do % 2 == 0 && ld_dq % 2 == 0 &&
                 ld_dk % 2 == 0 && ld_dv % 2 == 0, "calm_attention_bwd: leading dims must be even");
  CALM_CHECK_ARG(ld_q % 4 == 0 && ld_k % 4 == 0 && ld_v % 4 == 0 && ld_o % 4 == 0 && ld_do % 4 == 0,
                 "calm_attention_bwd: q/k/v/o/dO leading dims must be multiples of 4");
  const long long tokens = (long long)B * S;
  const long long nthr = tokens * heads;
  CALM_LAUNCH((attn_delta_kernel), (unsigned)((nthr + 255) / 256), 256, 0, stream,
              reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), delta, ld_o, ld_do, tokens, S, heads, hd);
  CALM_CHECK_LAUNCH("calm_attention_bwd(delta)");
  {
    const int64_t lds[4] = {ld_q, ld_k, ld_v, ld_do};
    const void* ptrs[6] = {q, k, v, d_o, bias, dbias};
#ifndef CALM_FORCE_LONG_ATTENTION
    if (ds_scratch && calm_attention_tc_eligible(B, S, heads, hd, lds, 4, ptrs, 6))
#else
    if (false)
#endif
      return calm_attention_bwd_tc(q, k, v, bias, d_o, lse, delta, dq, dk, dv, dbias, ds_scratch, ld_q, ld_k, ld_v, ld_do, ld_dq, ld_dk,
                                   ld_dv, B, S, heads, hd, stream);
    if (ds_scratch && ld_dq % 4 == 0 && ld_dk % 4 == 0 && ld_dv % 4 == 0 && (reinterpret_cast<uintptr_t>(ds_scratch) & 15) == 0 &&
        calm_attention_long_eligible(B, S, heads, hd, lds, 4, ptrs, 6))
      return calm_attention_bwd_long(q, k, v, bias, d_o, lse, delta, dq, dk, dv, dbias, ds_scratch, ld_q, ld_k, ld_v, ld_do, ld_dq, ld_dk,
                                     ld_dv, B, S, heads, hd, stream);
  }
  DISPATCH_HDP(hd, return launch_bwd<HDP>(q, k, v, bias, d_o, lse, delta, dq, dk, dv, dbias, ds_scratch, ld_q, ld_k, ld_v, ld_do, ld_dq, ld_dk,
                                          ld_dv, B, S, heads, hd, stream));
  return CALM_OK;
}
