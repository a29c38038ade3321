// Latent-bottleneck sampling of the stage-changing cross blocks (Vi_Tools_CNN_less_V2.py:232-244) fused with the
// ResidualStateManager running sum and KL terms (:23-30, mode "sum"):
//   [mu | rho] = encoder_{q,kv}(x)  (bf16 GEMM output, (rows, 2M));  sigma = softplus(rho) + 1e-6
//   z = mu + eps * sigma   (eps ~ N(0,1) drawn by the host with torch.randn_like so the RNG stream matches the reference)
//   zsum = zsum_prev + z ;  KL = -0.5 * mean(1 + 2 log sigma - mu^2 - sigma^2)   (per-CTA partial sums returned)
// HBM-bound elementwise kernels: 4 B (mu,rho) + 4 B eps + 4 B zsum_prev read, 4 + 2 B written per element.
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

constexpr int LAT_THREADS = 256;

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_ref(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(LAT_THREADS)
latent_fwd_kernel(const bf16* __restrict__ mv, const float* __restrict__ eps, const float* __restrict__ zsum_prev,
                  float* __restrict__ zsum, bf16* __restrict__ zsum_bf16, float* __restrict__ kl_partial, long long rows, int Mh) {
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

// kl_out = kl_prev + scale * (sum part_q + sum part_kv): running KL total of ResidualStateManager (Vi_Tools…:24-26)
__global__ void latent_kl_kernel(const float* __restrict__ part_q, const float* __restrict__ part_kv, int nblocks,
                                 const float* __restrict__ kl_prev, float* __restrict__ kl_out, float scale) {
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
  latent_kl_kernel<<<1, 256, 0, stream>>>(part_q, part_kv, nblocks, kl_prev, kl_out, scale);
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
  latent_fwd_kernel<<<nblocks, LAT_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(mv), eps, zsum_prev, zsum,
                                                         reinterpret_cast<bf16*>(zsum_bf16), kl_partial, rows, Mh);
  CALM_CHECK_LAUNCH("calm_latent_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_latent_bwd(const void* mv, const float* eps, const float* dz, const void* dz_bf16, float kl_scale,
                                   const float* dkl, void* dmv, float* dz_total, int64_t rows, int32_t Mh, cudaStream_t stream) {
  CALM_CHECK_ARG(rows > 0 && Mh > 0, "calm_latent_bwd: rows=%lld Mh=%d", (long long)rows, Mh);
  latent_bwd_kernel<<<calm_latent_blocks(rows, Mh), LAT_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(mv), eps, dz,
                                                                              reinterpret_cast<const bf16*>(dz_bf16), kl_scale, dkl,
                                                                              reinterpret_cast<bf16*>(dmv), dz_total, rows, Mh);
  CALM_CHECK_LAUNCH("calm_latent_bwd");
  return CALM_OK;
}
