// Small HBM-bound helpers of the CALM-ViT hot path: token re-tokenisation transposes, bias-gradient column sums,
// skip-connection adds, casts and the classifier's sequence mean.
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

// out[b, j, i, :] = in[b, i, j, :] on (B, S, S, 3) fp32: the row<->column token swap of Block.forward
// (Vi_Tools_CNN_less_V2.py:394-395,397-398). 32x32-pixel tiles staged through shared memory, 12 B pixels.
// Global accesses are 16-byte (a row of 32 pixels = 96 floats = 24 float4; S * 3 floats per image row is a multiple of 4 whenever S is),
// the transposition itself happens on scalar shared-memory reads (pitch 97: conflict-free both ways).
template <bool VEC>
__global__ void __launch_bounds__(256)
token_transpose_kernel(const float* __restrict__ in, const float* __restrict__ addend, float* __restrict__ out,
                       bf16* __restrict__ out16, int S) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float tile[32][32 * 3 + 1];
  const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const float* src = in + (long long)b * S * S * 3;
  float* dst = out + (long long)b * S * S * 3;
  bf16* dst16 = out16 ? out16 + (long long)b * S * S * 3 : nullptr;   // optional bf16 copy (the next GEMM's operand form)
  const float* add = addend ? addend + (long long)b * S * S * 3 : nullptr;
  if (VEC) {
    const int row_f = S * 3;
    for (int idx = threadIdx.x; idx < 32 * 24; idx += 256) {
      const int r = idx / 24, q4 = (idx - r * 24) * 4;
      const int i = i0 + r, col = j0 * 3 + q4;
      if (i < S && col < row_f) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (long long)i * row_f + col));
        tile[r][q4] = v.x; tile[r][q4 + 1] = v.y; tile[r][q4 + 2] = v.z; tile[r][q4 + 3] = v.w;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 24; idx += 256) {
      const int r = idx / 24, q4 = (idx - r * 24) * 4;   // output row j0 + r, output floats [q4, q4 + 4) of the segment starting at pixel i0
      const int j = j0 + r, col = i0 * 3 + q4;
      if (j < S && col < row_f) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int e = q4 + k; v[k] = tile[e / 3][r * 3 + e % 3]; }
        const long long o = (long long)j * row_f + col;
        if (add) {
          const float4 a = __ldcs(reinterpret_cast<const float4*>(add + o));
          v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
        }
        *reinterpret_cast<float4*>(dst + o) = make_float4(v[0], v[1], v[2], v[3]);
        if (dst16) *reinterpret_cast<uint2*>(dst16 + o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
      const int r = idx / 96, cc = idx - r * 96;
      const int i = i0 + r, j = j0 + cc / 3;
      if (i < S && j < S) tile[r][cc] = src[((long long)i * S + j0) * 3 + cc];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
      const int r = idx / 96, cc = idx - r * 96;  // output row j0 + r, output pixel column i0 + cc/3
      const int j = j0 + r, i = i0 + cc / 3;
      if (i < S && j < S) {
        const long long o = ((long long)j * S + i0) * 3 + cc;
        float val = tile[cc / 3][r * 3 + cc % 3];
        if (add) val += add[o];
        dst[o] = val;
        if (dst16) dst16[o] = __float2bfloat16(val);
      }
    }
  }
}

// (B,3,S,S) NCHW -> (B,S,S,3) tokens  (x.permute(0,2,3,1).reshape(B,S,3S), Vi_Tools_CNN_less_V2.py:389-391)
__global__ void nchw_to_tokens_kernel(const float* __restrict__ in, float* __restrict__ out, long long npix_total, long long plane) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*S*S*3 output elements
  if (i >= npix_total * 3) return;
  const long long pix = i / 3;
  const int ch = (int)(i - pix * 3);
  const long long b = pix / plane, p = pix - b * plane;
  out[i] = in[(b * 3 + ch) * plane + p];
}

// per-CTA column sums of a bf16 (rows, N) matrix: a thread owns 2 adjacent columns (one 4-byte load per row), rows are
// strided over the CTAs and unrolled 4x for memory-level parallelism
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ partial, long long rows, int N) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  for (int c = 2 * threadIdx.x; c < N; c += 512) {
    float s0 = 0.f, s1 = 0.f;
    long long r = blockIdx.x;
    const long long step = gridDim.x;
    for (; r + 3 * step < rows; r += 4 * step) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + c));
      const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (r + step) * ld + c));
      const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (r + 2 * step) * ld + c));
      const float2 e = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (r + 3 * step) * ld + c));
      s0 += (a.x + b.x) + (d.x + e.x);
      s1 += (a.y + b.y) + (d.y + e.y);
    }
    for (; r < rows; r += step) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + c));
      s0 += a.x; s1 += a.y;
    }
    partial[(size_t)blockIdx.x * N + c] = s0;
    partial[(size_t)blockIdx.x * N + c + 1] = s1;
  }
}

// 128-bit variant (N, ld multiples of 8, 16-byte aligned base, N <= 2048): a thread owns 8 adjacent columns of one of the
// R = 256 / (N/8) rows the CTA reads side by side, 8 independent 16-byte loads in flight per thread (the 4-byte form above
// kept ~7 KB per SM in flight: 2.4 TB/s); the R row slots meet in shared memory.
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ partial, long long rows, int N) {
  __shared__ float red[2048];
  pdl_wait(); pdl_launch_small_dependent();
  const int chunks = N >> 3, R = 256 / chunks;
  const int r = threadIdx.x / chunks, k = threadIdx.x - r * chunks;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  if (r < R) {
    const long long step = (long long)gridDim.x * R;
    long long row = (long long)blockIdx.x * R + r;
    const bf16* p = x + 8 * k;
    for (; row + 7 * step < rows; row += 8 * step) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(p + (row + u * step) * ld));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float2 a = unpack_bf16x2(v[u].x), b = unpack_bf16x2(v[u].y), c = unpack_bf16x2(v[u].z), d = unpack_bf16x2(v[u].w);
        s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
      }
    }
    for (; row < rows; row += step) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p + row * ld));
      const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
      s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[r * N + 8 * k + e] = s[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += 256) {
    float t = 0.f;
    for (int rr = 0; rr < R; ++rr) t += red[rr * N + c];
    partial[(size_t)blockIdx.x * N + c] = t;
  }
}

// out[c] = sum_p partial[p][c]: 32 columns per CTA, 32 row-groups (one warp each, two loads in flight per thread) combined through
// shared memory (deterministic)
constexpr int RC_GROUPS = 32;
__global__ void __launch_bounds__(32 * RC_GROUPS)
reduce_cols_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int n) {
  __shared__ float red[RC_GROUPS][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  pdl_wait();                                   // launched while the column-sum kernel drains
  float s0 = 0.f, s1 = 0.f;
  if (c < n) {
    int p = ry;
    for (; p + RC_GROUPS < nparts; p += 2 * RC_GROUPS) {
      s0 += partial[(size_t)p * n + c];
      s1 += partial[(size_t)(p + RC_GROUPS) * n + c];
    }
    if (p < nparts) s0 += partial[(size_t)p * n + c];
  }
  red[ry][cx] = s0 + s1;
  __syncthreads();
  if (ry == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < RC_GROUPS; ++k) t += red[k][cx];
    out[c] = t;
  }
}

__global__ void add3_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ c,
                            float4* __restrict__ out, long long n4) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = a[i];
    const float4 w = b[i];
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    if (c) { const float4 u = c[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
    out[i] = v;
  }
}

__global__ void cast_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    uint2 o; o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
    out[i] = o;
  }
}

__global__ void cast_f32_kernel(const uint2* __restrict__ in, float4* __restrict__ out, long long n4) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint2 v = in[i];
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
    out[i] = make_float4(a.x, a.y, b.x, b.y);
  }
}

// out[b, d] = mean_s x[b, s, d]
__global__ void seq_mean_fwd_kernel(const float* __restrict__ x, bf16* __restrict__ out, int S, int D) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int b = blockIdx.y, d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float* p = x + (long long)b * S * D + d;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += p[(long long)i * D];
  out[(long long)b * D + d] = __float2bfloat16(s / S);
}
__global__ void seq_mean_bwd_kernel(const bf16* __restrict__ dout, float* __restrict__ dx, int S, int D) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int b = blockIdx.z, s = blockIdx.y, d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  dx[((long long)b * S + s) * D + d] = __bfloat162float(dout[(long long)b * D + d]) / S;
}

// 4 columns per thread, 16 sequence positions per CTA: dout is read once per CTA, dx leaves as float4 (the scalar form above
// launches B*S*D/128 CTAs that each write 512 bytes: 180 us for the 154 MB of the 224^2 classifier)
constexpr int SMB_ROWS = 16;
__global__ void __launch_bounds__(256)
seq_mean_bwd4_kernel(const bf16* __restrict__ dout, float* __restrict__ dx, int S, int D) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int b = blockIdx.y, s0 = blockIdx.x * SMB_ROWS;
  for (int q = threadIdx.x; q < (D >> 2); q += 256) {
    const uint2 v = *reinterpret_cast<const uint2*>(dout + (long long)b * D + 4 * q);
    const float2 lo = unpack_bf16x2(v.x), hi = unpack_bf16x2(v.y);
    const float4 o = make_float4(lo.x / S, lo.y / S, hi.x / S, hi.y / S);
#pragma unroll 4
    for (int s = s0; s < min(S, s0 + SMB_ROWS); ++s) reinterpret_cast<float4*>(dx + ((long long)b * S + s) * D)[q] = o;
  }
}

}  // namespace

extern "C" int32_t calm_token_transpose(const float* in, const float* addend, float* out, void* out_bf16, int32_t B, int32_t S,
                                        cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0 && in != out, "calm_token_transpose: B=%d S=%d (out of place only)", B, S);
  dim3 grid((S + 31) / 32, (S + 31) / 32, B);
  const bool vec = S % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(addend) |
                                   reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0;
  if (vec) CALM_LAUNCH((token_transpose_kernel<true>), grid, 256, 0, stream, in, addend, out, reinterpret_cast<bf16*>(out_bf16), S);
  else CALM_LAUNCH((token_transpose_kernel<false>), grid, 256, 0, stream, in, addend, out, reinterpret_cast<bf16*>(out_bf16), S);
  CALM_CHECK_LAUNCH("calm_token_transpose");
  return CALM_OK;
}

extern "C" int32_t calm_nchw_to_tokens(const float* in, float* out, int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_nchw_to_tokens: B=%d S=%d", B, S);
  const long long plane = (long long)S * S, npix = plane * B;
  CALM_LAUNCH((nchw_to_tokens_kernel), (unsigned)((npix * 3 + 255) / 256), 256, 0, stream, in, out, npix, plane);
  CALM_CHECK_LAUNCH("calm_nchw_to_tokens");
  return CALM_OK;
}

extern "C" int32_t calm_colsum_parts(int64_t rows, int32_t N) {
  (void)N;
  const long long cap = 2LL * calm_num_sms();   // 2 CTAs x 256 threads x 8 loads of 16 bytes = 64 KB in flight per SM
  return (int32_t)(rows < cap ? rows : cap);
}

extern "C" int32_t calm_colsum(const void* x, int64_t ld, float* partial, int32_t nparts, float* out, int64_t rows, int32_t N,
                               cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && N > 0 && N % 2 == 0 && ld % 2 == 0, "calm_colsum: rows=%lld N=%d ld=%lld (N, ld must be even)", (long long)rows, N, (long long)ld);
  CALM_CHECK_ARG(nparts == calm_colsum_parts(rows, N), "calm_colsum: nparts=%d expected %d", nparts, calm_colsum_parts(rows, N));
  if (N % 8 == 0 && ld % 8 == 0 && N <= 2048 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    CALM_LAUNCH((colsum_vec_kernel), nparts, 256, 0, stream, reinterpret_cast<const bf16*>(x), ld, partial, rows, N);
  else
    CALM_LAUNCH((colsum_kernel), nparts, 256, 0, stream, reinterpret_cast<const bf16*>(x), ld, partial, rows, N);
  CALM_CHECK_LAUNCH("calm_colsum");
  {
    cudaError_t e = calm_launch_pdl(reduce_cols_kernel, dim3((N + 31) / 32), dim3(32 * RC_GROUPS), 0, stream, nullptr, 0,
                                    (const float*)partial, out, nparts, N);
    if (e != cudaSuccess) { calm_set_error("calm_colsum(reduce): launch failed: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
  }
  return CALM_OK;
}

extern "C" int32_t calm_add3(const float* a, const float* b, const float* c, float* out, int64_t n, cudaStream_t stream) {
  CALM_CHECK_ARG(n > 0 && n % 4 == 0, "calm_add3: n=%lld must be a positive multiple of 4", (long long)n);
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = 16LL * calm_num_sms();
  if (blocks > cap) blocks = cap;
  CALM_LAUNCH((add3_kernel), (unsigned)blocks, 256, 0, stream, reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                    reinterpret_cast<const float4*>(c), reinterpret_cast<float4*>(out), n4);
  CALM_CHECK_LAUNCH("calm_add3");
  return CALM_OK;
}

extern "C" int32_t calm_cast_bf16(const float* in, void* out, int64_t n, cudaStream_t stream) {
  CALM_CHECK_ARG(n > 0 && n % 4 == 0, "calm_cast_bf16: n=%lld must be a positive multiple of 4", (long long)n);
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = 16LL * calm_num_sms();
  if (blocks > cap) blocks = cap;
  CALM_LAUNCH((cast_bf16_kernel), (unsigned)blocks, 256, 0, stream, reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out), n4);
  CALM_CHECK_LAUNCH("calm_cast_bf16");
  return CALM_OK;
}

extern "C" int32_t calm_cast_f32(const void* in, float* out, int64_t n, cudaStream_t stream) {
  CALM_CHECK_ARG(n > 0 && n % 4 == 0, "calm_cast_f32: n=%lld must be a positive multiple of 4", (long long)n);
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = 16LL * calm_num_sms();
  if (blocks > cap) blocks = cap;
  CALM_LAUNCH((cast_f32_kernel), (unsigned)blocks, 256, 0, stream, reinterpret_cast<const uint2*>(in), reinterpret_cast<float4*>(out), n4);
  CALM_CHECK_LAUNCH("calm_cast_f32");
  return CALM_OK;
}

extern "C" int32_t calm_seq_mean_fwd(const float* x, void* out_bf16, int32_t B, int32_t S, int32_t D, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0 && D > 0, "calm_seq_mean_fwd: bad dims");
  dim3 grid((D + 127) / 128, B);
  CALM_LAUNCH((seq_mean_fwd_kernel), grid, 128, 0, stream, x, reinterpret_cast<bf16*>(out_bf16), S, D);
  CALM_CHECK_LAUNCH("calm_seq_mean_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_seq_mean_bwd(const void* dout_bf16, float* dx, int32_t B, int32_t S, int32_t D, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0 && D > 0, "calm_seq_mean_bwd: bad dims");
  if (D % 4 == 0 && ((reinterpret_cast<uintptr_t>(dout_bf16) & 7) == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0)) {
    dim3 grid4((S + SMB_ROWS - 1) / SMB_ROWS, B);
    CALM_LAUNCH((seq_mean_bwd4_kernel), grid4, 256, 0, stream, reinterpret_cast<const bf16*>(dout_bf16), dx, S, D);
  } else {
    dim3 grid((D + 127) / 128, S, B);
    CALM_LAUNCH((seq_mean_bwd_kernel), grid, 128, 0, stream, reinterpret_cast<const bf16*>(dout_bf16), dx, S, D);
  }
  CALM_CHECK_LAUNCH("calm_seq_mean_bwd");
  return CALM_OK;
}
