// Latent-bottleneck sampling of the stage-changing cross blocks (Vi_Tools_CNN_less_V2.py:232-244) fused with the
// ResidualStateManager running sum and KL terms (:23-30, mode "sum"):
//   [mu | rho] = encoder_{q,kv}(x)  (bf16 GEMM output, (rows, 2M));  sigma = softplus(rho) + 1e-6
//   z = mu + eps * sigma   (eps ~ N(0,1) drawn by the host with torch.randn_like so the RNG stream matches the reference)
//   zsum = zsum_prev + z ;  KL = -0.5 * mean(1 + 2 log sigma - mu^2 - sigma^2)   (per-CTA partial sums returned)
// HBM-bound elementwise kernels: 4 B (mu,rho) + 4 B eps + 4 B zsum_prev read, 4 + 2 B written per element.
#include "common.cuh"
#include <initializer_list>
#include "../../include/calm_b200.h"

namespace {

constexpr int LAT_THREADS = 256;

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_ref(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(LAT_THREADS)
latent_fwd_kernel(const bf16* __restrict__ mv, const float* __restrict__ eps, const float* __restrict__ zsum_prev,
                  float* __restrict__ zsum, bf16* __restrict__ zsum_bf16, float* __restrict__ kl_partial, long long rows, int Mh) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const long long total = rows * Mh;
  float kl = 0.f;
  for (long long i = (long long)blockIdx.x * LAT_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * LAT_THREADS) {
    const long long r = i / Mh;
    const int c = (int)(i - r * Mh);
    const float mu = __bfloat162float(mv[r * 2 * Mh + c]);
    const float rho = __bfloat162float(mv[r * 2 * Mh + Mh + c]);
    const float sg = softplus_ref(rho) + 1e-6f;
    float z = mu;
    if (eps) z += eps[i] * sg;
    if (zsum_prev) z += zsum_prev[i];
    zsum[i] = z;
    if (zsum_bf16) zsum_bf16[i] = __float2bfloat16(z);
    kl += 1.f + 2.f * logf(sg) - mu * mu - sg * sg;
  }
  kl = block_sum(kl, red);
  if (threadIdx.x == 0) kl_partial[blockIdx.x] = kl;
}

__global__ void __launch_bounds__(LAT_THREADS)
latent_bwd_kernel(const bf16* __restrict__ mv, const float* __restrict__ eps, const float* __restrict__ dz,
                  const bf16* __restrict__ dz2, float kl_scale, const float* __restrict__ dkl, bf16* __restrict__ dmv,
                  float* __restrict__ dz_total, long long rows, int Mh) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const long long total = rows * Mh;
  // KL = kl_scale * sum(1 + 2 log sg - mu^2 - sg^2)  =>  dKL/dmu = -2 kl_scale mu ; dKL/dsg = -2 kl_scale (sg - 1/sg)
  const float gk = dkl ? -2.f * (*dkl) * kl_scale : 0.f;
  for (long long i = (long long)blockIdx.x * LAT_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * LAT_THREADS) {
    const long long r = i / Mh;
    const int c = (int)(i - r * Mh);
    const float mu = __bfloat162float(mv[r * 2 * Mh + c]);
    const float rho = __bfloat162float(mv[r * 2 * Mh + Mh + c]);
    const float sg = softplus_ref(rho) + 1e-6f;
    float g = dz ? dz[i] : 0.f;
    if (dz2) g += __bfloat162float(dz2[i]);   // gradient through the bf16 copy of the running sum (seq-axis GEMM operand)
    if (dz_total) dz_total[i] = g;            // = gradient wrt the previous running sum (zsum = zsum_prev + z)
    const float dmu = g + gk * mu;
    float dsg = gk * (sg - 1.f / sg);
    if (eps) dsg += g * eps[i];
    dmv[r * 2 * Mh + c] = __float2bfloat16(dmu);
    dmv[r * 2 * Mh + Mh + c] = __float2bfloat16(dsg * sigmoid_ref(rho));
  }
}

// 4-column variants (Mh % 4 == 0, 16-byte aligned operands): one thread handles 4 adjacent latent columns per step with 8 / 16
// byte accesses (the scalar kernels above moved 2- and 4-byte words with one 64-bit division per element: 46-51 us per call for
// 88-108 MB, 30 % of the HBM rate). Same arithmetic, same per-element operation order.
__device__ __forceinline__ void ld4_bf16(const bf16* p, float* f) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void st4_bf16(bf16* p, const float* f) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
}

__global__ void __launch_bounds__(LAT_THREADS)
latent_fwd4_kernel(const bf16* __restrict__ mv, const float* __restrict__ eps, const float* __restrict__ zsum_prev,
                   float* __restrict__ zsum, bf16* __restrict__ zsum_bf16, float* __restrict__ kl_partial, long long rows, int Mh) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const int q = Mh >> 2;
  const long long total4 = rows * q;
  float kl = 0.f;
  for (long long i4 = (long long)blockIdx.x * LAT_THREADS + threadIdx.x; i4 < total4; i4 += (long long)gridDim.x * LAT_THREADS) {
    const long long r = i4 / q;
    const int c = (int)(i4 - r * q) << 2;
    const long long i = r * Mh + c;
    float mu[4], rho[4], z[4];
    ld4_bf16(mv + r * 2 * Mh + c, mu);
    ld4_bf16(mv + r * 2 * Mh + Mh + c, rho);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f), pv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (eps) e = __ldcs(reinterpret_cast<const float4*>(eps + i));
    if (zsum_prev) pv = *reinterpret_cast<const float4*>(zsum_prev + i);
    const float ev[4] = {e.x, e.y, e.z, e.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sg = softplus_ref(rho[k]) + 1e-6f;
      float zz = mu[k];
      if (eps) zz += ev[k] * sg;
      if (zsum_prev) zz += pp[k];
      z[k] = zz;
      kl += 1.f + 2.f * logf(sg) - mu[k] * mu[k] - sg * sg;
    }
    *reinterpret_cast<float4*>(zsum + i) = make_float4(z[0], z[1], z[2], z[3]);
    if (zsum_bf16) st4_bf16(zsum_bf16 + i, z);
  }
  kl = block_sum(kl, red);
  if (threadIdx.x == 0) kl_partial[blockIdx.x] = kl;
}

__global__ void __launch_bounds__(LAT_THREADS)
latent_bwd4_kernel(const bf16* __restrict__ mv, const float* __restrict__ eps, const float* __restrict__ dz,
                   const bf16* __restrict__ dz2, float kl_scale, const float* __restrict__ dkl, bf16* __restrict__ dmv,
                   float* __restrict__ dz_total, long long rows, int Mh) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int q = Mh >> 2;
  const long long total4 = rows * q;
  const float gk = dkl ? -2.f * (*dkl) * kl_scale : 0.f;
  for (long long i4 = (long long)blockIdx.x * LAT_THREADS + threadIdx.x; i4 < total4; i4 += (long long)gridDim.x * LAT_THREADS) {
    const long long r = i4 / q;
    const int c = (int)(i4 - r * q) << 2;
    const long long i = r * Mh + c;
    float mu[4], rho[4], g[4] = {0.f, 0.f, 0.f, 0.f}, dmu[4], drho[4];
    ld4_bf16(mv + r * 2 * Mh + c, mu);
    ld4_bf16(mv + r * 2 * Mh + Mh + c, rho);
    if (dz) { const float4 t = __ldcs(reinterpret_cast<const float4*>(dz + i)); g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w; }
    if (dz2) { float t[4]; ld4_bf16(dz2 + i, t); g[0] += t[0]; g[1] += t[1]; g[2] += t[2]; g[3] += t[3]; }
    if (dz_total) *reinterpret_cast<float4*>(dz_total + i) = make_float4(g[0], g[1], g[2], g[3]);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (eps) e = __ldcs(reinterpret_cast<const float4*>(eps + i));
    const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sg = softplus_ref(rho[k]) + 1e-6f;
      dmu[k] = g[k] + gk * mu[k];
      float dsg = gk * (sg - 1.f / sg);
      if (eps) dsg += g[k] * ev[k];
      drho[k] = dsg * sigmoid_ref(rho[k]);
    }
    st4_bf16(dmv + r * 2 * Mh + c, dmu);
    st4_bf16(dmv + r * 2 * Mh + Mh + c, drho);
  }
}

__host__ inline bool latent_vec_ok(int Mh, std::initializer_list<const void*> ptrs) {
  bool ok = Mh % 4 == 0;
  for (const void* p : ptrs) ok = ok && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  return ok;
}

// kl_out = kl_prev + scale * (sum part_q + sum part_kv): running KL total of ResidualStateManager (Vi_Tools…:24-26)
__global__ void latent_kl_kernel(const float* __restrict__ part_q, const float* __restrict__ part_kv, int nblocks,
                                 const float* __restrict__ kl_prev, float* __restrict__ kl_out, float scale) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += part_q[i] + part_kv[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) kl_out[0] = (kl_prev ? kl_prev[0] : 0.f) + scale * s;
}

}  // namespace

extern "C" int32_t calm_latent_kl(const float* part_q, const float* part_kv, int32_t nblocks, const float* kl_prev,
                                  float* kl_out, float scale, cudaStream_t stream) {
  CALM_CHECK_ARG(part_q && part_kv && kl_out && nblocks > 0, "calm_latent_kl: bad args");
  CALM_LAUNCH((latent_kl_kernel), 1, 256, 0, stream, part_q, part_kv, nblocks, kl_prev, kl_out, scale);
  CALM_CHECK_LAUNCH("calm_latent_kl");
  return CALM_OK;
}

extern "C" int32_t calm_latent_blocks(int64_t rows, int32_t Mh) {
  long long need = (rows * Mh + LAT_THREADS - 1) / LAT_THREADS;
  const long long cap = 8LL * calm_num_sms();
  return (int32_t)(need < cap ? need : cap);
}

extern "C" int32_t calm_latent_fwd(const void* mv, const float* eps, const float* zsum_prev, float* zsum, void* zsum_bf16,
                                   float* kl_partial, int32_t nblocks, int64_t rows, int32_t Mh, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && Mh > 0, "calm_latent_fwd: rows=%lld Mh=%d", (long long)rows, Mh);
  CALM_CHECK_ARG(nblocks == calm_latent_blocks(rows, Mh), "calm_latent_fwd: nblocks=%d expected %d", nblocks, calm_latent_blocks(rows, Mh));
  if (latent_vec_ok(Mh, {mv, eps, zsum_prev, zsum, zsum_bf16}))
    CALM_LAUNCH((latent_fwd4_kernel), nblocks, LAT_THREADS, 0, stream, reinterpret_cast<const bf16*>(mv), eps, zsum_prev, zsum,
                                                            reinterpret_cast<bf16*>(zsum_bf16), kl_partial, rows, Mh);
  else
    CALM_LAUNCH((latent_fwd_kernel), nblocks, LAT_THREADS, 0, stream, reinterpret_cast<const bf16*>(mv), eps, zsum_prev, zsum,
                                                           reinterpret_cast<bf16*>(zsum_bf16), kl_partial, rows, Mh);
  CALM_CHECK_LAUNCH("calm_latent_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_latent_bwd(const void* mv, const float* eps, const float* dz, const void* dz_bf16, float kl_scale,
                                   const float* dkl, void* dmv, float* dz_total, int64_t rows, int32_t Mh, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && Mh > 0, "calm_latent_bwd: rows=%lld Mh=%d", (long long)rows, Mh);
  if (latent_vec_ok(Mh, {mv, eps, dz, dz_bf16, dmv, dz_total}))
    CALM_LAUNCH((latent_bwd4_kernel), calm_latent_blocks(rows, Mh), LAT_THREADS, 0, stream, reinterpret_cast<const bf16*>(mv), eps, dz,
                                                                                 reinterpret_cast<const bf16*>(dz_bf16), kl_scale, dkl,
                                                                                 reinterpret_cast<bf16*>(dmv), dz_total, rows, Mh);
  else
    CALM_LAUNCH((latent_bwd_kernel), calm_latent_blocks(rows, Mh), LAT_THREADS, 0, stream, reinterpret_cast<const bf16*>(mv), eps, dz,
                                                                                reinterpret_cast<const bf16*>(dz_bf16), kl_scale, dkl,
                                                                                reinterpret_cast<bf16*>(dmv), dz_total, rows, Mh);
  CALM_CHECK_LAUNCH("calm_latent_bwd");
  return CALM_OK;
}
