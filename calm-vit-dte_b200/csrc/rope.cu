// RoPE with learned inverse frequencies (Vi_Tools_CNN_less_V2.py:55-95, NeoX rotate-half) fused with the per-head
// [content | rope] concat of the decoupled-RoPE latent blocks (:278-281). Non-reduce blocks use dc = 0 (:283-285).
// cos/sin are rebuilt from inv_freq every call (the parameter is learned, :70-72,86-91); backward returns d inv_freq.
// All math fp32, storage bf16 — the reference promotes to fp32 against the fp32 cos table and SDPA/bmm re-round to bf16.
#include "common.cuh"
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

// One CTA per ROPE_TPC consecutive tokens; a thread resolves its (head, element) once and then walks the CTA's tokens, so the
// index arithmetic is amortised and ROPE_TPC independent load chains are in flight per thread.
constexpr int ROPE_TPC = 8;
__global__ void __launch_bounds__(256)
rope_fwd_kernel(const bf16* __restrict__ content, long long ld_content, const bf16* __restrict__ ropein, long long ld_rope,
                bf16* __restrict__ out, long long ld_out, const float* __restrict__ cs, long long tokens, int S, int heads,
                int dc, int dr) {
  const int half = dr >> 1;
  const int per_head = dc + half;  // work items per (token, head): dc copies + half rotations
  const int per_tok = heads * per_head;
  const long long t0 = (long long)blockIdx.x * ROPE_TPC;
  const int nt = (int)min((long long)ROPE_TPC, tokens - t0);
  const int p0 = (int)(t0 % S);
  const float2* cs2 = reinterpret_cast<const float2*>(cs);
  for (int w = threadIdx.x; w < per_tok; w += blockDim.x) {
    const int h = w / per_head, i = w - h * per_head;
    if (i < dc) {
#pragma unroll
      for (int k = 0; k < ROPE_TPC; ++k)
        if (k < nt) out[(t0 + k) * ld_out + (long long)h * (dc + dr) + i] = content[(t0 + k) * ld_content + (long long)h * dc + i];
    } else {
      const int j = i - dc;
      float x1[ROPE_TPC], x2[ROPE_TPC];
      float2 csv[ROPE_TPC];
#pragma unroll
      for (int k = 0; k < ROPE_TPC; ++k) {
        if (k < nt) {
          int pos = p0 + k;
          if (pos >= S) pos -= S;
          const bf16* r = ropein + (t0 + k) * ld_rope + (long long)h * dr;
          x1[k] = __bfloat162float(r[j]);
          x2[k] = __bfloat162float(r[j + half]);
          csv[k] = cs2[pos * half + j];
        }
      }
#pragma unroll
      for (int k = 0; k < ROPE_TPC; ++k) {
        if (k < nt) {
          bf16* o = out + (t0 + k) * ld_out + (long long)h * (dc + dr) + dc;
          o[j] = __float2bfloat16(x1[k] * csv[k].x - x2[k] * csv[k].y);
          o[j + half] = __float2bfloat16(x2[k] * csv[k].x + x1[k] * csv[k].y);
        }
      }
    }
  }
}

// grid (S, ROPE_BCH); each CTA owns one position and a slice of the batch, threads own (head, j) pairs.
__global__ void rope_bwd_kernel(const bf16* __restrict__ dout, long long ld_dout, const bf16* __restrict__ out, long long ld_out,
                                bf16* __restrict__ dcontent, long long ld_dcontent, bf16* __restrict__ dropein, long long ld_drope,
                                const float* __restrict__ cs, float* __restrict__ dtheta_part, int B, int S, int heads, int dc, int dr) {
  extern __shared__ float dth_s[];  // heads * half floats (one slot per (head, j): deterministic reduce)
  const int half = dr >> 1;
  const int pos = blockIdx.x, chunk = blockIdx.y;
  const int bpc = (B + ROPE_BCH - 1) / ROPE_BCH;
  const int b0 = chunk * bpc, b1 = min(B, b0 + bpc);
  const int per_head = dc + half;
  for (int w = threadIdx.x; w < heads * per_head; w += blockDim.x) {
    const int h = w / per_head, i = w - h * per_head;
    if (i < dc) {
      for (int b = b0; b < b1; ++b) {
        const long long t = (long long)b * S + pos;
        dcontent[t * ld_dcontent + (long long)h * dc + i] = dout[t * ld_dout + (long long)h * (dc + dr) + i];
      }
    } else {
      const int j = i - dc;
      const float c = cs[2 * (pos * half + j)], s = cs[2 * (pos * half + j) + 1];
      float acc = 0.f;
      for (int b = b0; b < b1; ++b) {
        const long long t = (long long)b * S + pos;
        const bf16* dyp = dout + t * ld_dout + (long long)h * (dc + dr) + dc;
        const bf16* yp = out + t * ld_out + (long long)h * (dc + dr) + dc;
        const float dy1 = __bfloat162float(dyp[j]), dy2 = __bfloat162float(dyp[j + half]);
        const float y1 = __bfloat162float(yp[j]), y2 = __bfloat162float(yp[j + half]);
        bf16* dxp = dropein + t * ld_drope + (long long)h * dr;
        dxp[j] = __float2bfloat16(dy1 * c + dy2 * s);
        dxp[j + half] = __float2bfloat16(dy2 * c - dy1 * s);
        acc += y1 * dy2 - y2 * dy1;
      }
      dth_s[h * half + j] = acc;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float s = 0.f;
    for (int h = 0; h < heads; ++h) s += dth_s[h * half + j];
    dtheta_part[((size_t)chunk * S + pos) * half + j] = s;
  }
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
  CALM_CHECK_ARG(S >= ROPE_TPC, "calm_rope_fwd: S=%d too short", S);
  const long long blocks = (tokens + ROPE_TPC - 1) / ROPE_TPC;
  rope_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(content), ld_content,
                                                         reinterpret_cast<const bf16*>(ropein), ld_rope,
                                                         reinterpret_cast<bf16*>(out), ld_out, cos_sin, tokens, S, heads, dc, dr);
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
  int threads = heads * (dc + half);
  threads = ((threads + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  dim3 grid(S, ROPE_BCH);
  rope_bwd_kernel<<<grid, threads, (size_t)heads * half * sizeof(float), stream>>>(
      reinterpret_cast<const bf16*>(dout), ld_dout, reinterpret_cast<const bf16*>(out), ld_out, reinterpret_cast<bf16*>(dcontent),
      ld_dcontent, reinterpret_cast<bf16*>(dropein), ld_drope, cos_sin, dtheta_part, B, S, heads, dc, dr);
  CALM_CHECK_LAUNCH("calm_rope_bwd");
  rope_dfreq_kernel<<<half, 128, 0, stream>>>(dtheta_part, dinv_freq, S, half);
  CALM_CHECK_LAUNCH("calm_rope_bwd(dfreq)");
  return CALM_OK;
}
