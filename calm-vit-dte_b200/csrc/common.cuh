// Shared device/host helpers for the calm_b200 sm_100a kernel library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

#define CALM_OK 0
#define CALM_ERR_ARG (-1)
#define CALM_ERR_CUDA (-2)
#define CALM_ERR_UNSUPPORTED (-3)

// Bound of the mbarrier waits of the tcgen05 kernels: a protocol bug becomes a trap (a sticky CUDA error with the barrier id in
// the error-flag buffer) instead of a hung GPU. The default (2^37 cycles, about 70 s) is far above anything a legitimate wait
// can take, also under compute-sanitizer, cuda-gdb or a time-sliced GPU; -DCALM_MBAR_TIMEOUT_CYCLES=... tightens it for bring-up.
#ifndef CALM_MBAR_TIMEOUT_CYCLES
#define CALM_MBAR_TIMEOUT_CYCLES (1LL << 37)
#endif

void calm_set_error(const char* fmt, ...);
extern int* g_calm_err_flag;

#define CALM_CHECK_ARG(cond, ...)                    \
  do {                                               \
    if (!(cond)) {                                   \
      calm_set_error(__VA_ARGS__);                   \
      return CALM_ERR_ARG;                           \
    }                                                \
  } while (0)

#define CALM_CHECK_LAUNCH(name)                                               \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      calm_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return CALM_ERR_CUDA;                                                   \
    }                                                                         \
  } while (0)

static inline int calm_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
static inline int calm_num_sms() {
  static std::atomic<int> sms[64];
  const int dev = calm_current_device() & 63;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// cudaFuncSetAttribute is a per-device setting: a call site keeps one "configured" bit per device ordinal.
//   static CalmDeviceOnce once;  if (once.pending()) { cudaFuncSetAttribute(...); once.done(); }
struct CalmDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool pending() const { return !((mask.load(std::memory_order_acquire) >> (calm_current_device() & 63)) & 1ull); }
  void done() { mask.fetch_or(1ull << (calm_current_device() & 63), std::memory_order_release); }
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------------
// A training step is ~1,500 short kernels in one stream; between two of them the GPU idles for the grid-completion -> flush ->
// next-launch latency. Kernels launched through calm_launch_pdl / CALM_LAUNCH carry cudaLaunchAttributeProgrammaticStreamSerialization:
// their CTAs may be scheduled as soon as the predecessor's CTAs have exited (or earlier, if the predecessor triggers), and block in
// pdl_wait() until the predecessor has completed and its writes are visible.
// RULE: a kernel launched this way executes pdl_wait() before its first global-memory access (read OR write).
// Since every such kernel waits for its predecessor before finishing, completion stays transitive along the stream.
// CALM_PDL=0 in the environment (read once at load) launches the same kernels without the attribute (A/B and fault isolation).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Early trigger: lets the dependents' CTAs become resident while this grid still runs. Measured (r02f build, all small kernels converted):
// -0.2 ms/step at 224^2 but +2.7 ms at 512^2 — CTAs of several future kernels pile up resident and waiting, and a 227 KB GEMM /
// attention CTA of a forked stream no longer finds an empty SM. So the trigger is compiled out (-DCALM_PDL_EARLY brings it back):
// dependents launch when this grid's CTAs have exited, which still hides the completion -> flush -> launch latency.
// The exception: a kernel whose dependent is its own tiny reduction (LayerNorm backward -> dw partial sums, column sums -> their
// reduction): ~20 CTAs that follow immediately and have no PDL dependents that could pile up behind them.
__device__ __forceinline__ void pdl_launch_small_dependent() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
#ifdef CALM_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
inline bool calm_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CALM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
// extra = further launch attributes (e.g. the cluster dimension); n_extra <= 3
template <typename... KArgs, typename... Args>
inline cudaError_t calm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   const cudaLaunchAttribute* extra, int n_extra, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[4];
  int n = 0;
  for (int i = 0; i < n_extra && i < 3; ++i) attr[n++] = extra[i];
  if (calm_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute; the kernel must start with pdl_wait() (see above). Errors
// surface through the CALM_CHECK_LAUNCH that follows (cudaLaunchKernelEx records them as the last error).
#define CALM_LAUNCH(KERNEL, GRID, BLOCK, SMEM, STREAM, ...) \
  (void)calm_launch_pdl(KERNEL, dim3(GRID), dim3(BLOCK), (size_t)(SMEM), STREAM, nullptr, 0, __VA_ARGS__)

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` must hold >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// exact-erf GELU and its derivative in fp32: Phi(x) = 1 - 0.5 erfc(x/sqrt2) with erfc from Abramowitz-Stegun 7.1.26
// (|abs err| < 1.5e-7; one MUFU.RCP + one MUFU.EX2 — ~2.5x cheaper than erff, and the derivative reuses the exponential)
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void gelu_pair(float x, float& g, float& dg) {
  // 15 FP32/ALU + 2 MUFU instructions: the single-instruction rcp/ex2 forms (no denormal / range fix-up sequences: the
  // argument of rcp is in [1, inf), ex2 flushing a denormal result to 0 changes Phi by < 1e-38), 0.5 folded into the polynomial
  const float t = rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float e = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));   // exp(-x^2/2)
  const float q = p * t * e;          // Phi(-|x|)
  const float phi = 0.5f + copysignf(0.5f - q, x);
  g = x * phi;
  dg = fmaf(x * 0.3989422804014327f, e, phi);
}
// value only: x Phi(x) = max(x, 0) - |x| Phi(-|x|) saves the sign reconstruction of Phi (12 FP32 + 2 MUFU)
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float e = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));
  return fmaf(-fabsf(x), p * t * e, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float dgelu_erf(float x) { float g, dg; gelu_pair(x, g, dg); return dg; }

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  bf162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  bf162 h = *reinterpret_cast<bf162*>(&v);
  return __bfloat1622float2(h);
}

// ---- packed fp16 arithmetic (explicit PTX: the half2 intrinsics of cuda_fp16.h have no tanh / ex2 forms) ----------------
// GELU(erf) on channel / column PAIRS: Phi(x) = 0.5 + 0.5 tanh(x (a + b x^2)), a = 0.79880144, b = 0.03528205 (minimax fit of
// atanh(erf(x / sqrt2)) on [0, 4], |error| < 2.9e-4 = below the fp16 resolution of Phi; both coefficients positive, so the
// polynomial saturates monotonically and needs no clamp), gelu'(x) = Phi + x exp(-x^2/2) / sqrt(2 pi). Simulated in fp16
// arithmetic the value is 3.9e-4 rms from the exact function for x ~ N(0, 1.5): 6x finer than the bf16 rounding the
// reference's autocast path applies to the same activations (2.4e-3 rms).
__device__ __forceinline__ uint32_t h2fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t h2mul(uint32_t a, uint32_t b) { uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2tanh(uint32_t a) { uint32_t d; asm("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ uint32_t h2ex2(uint32_t a) { uint32_t d; asm("ex2.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ uint32_t h2pack(float lo, float hi) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ float2 h2unpack(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ float h2low(uint32_t v) { return __low2float(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ uint32_t bf2pack(float lo, float hi) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }

constexpr uint32_t H2_HALF = 0x38003800u, H2_ONE = 0x3c003c00u;
constexpr uint32_t H2_GA = 0x3a643a64u;   // 0.79880144
constexpr uint32_t H2_GB = 0x28842884u;   // 0.03528205
constexpr uint32_t H2_GK = 0xb9c5b9c5u;   // -0.5 log2(e)
constexpr uint32_t H2_GC = 0x36623662u;   // 1/sqrt(2 pi)

__device__ __forceinline__ uint32_t gelu_h2(uint32_t x) {
  const uint32_t x2 = h2mul(x, x);
  const uint32_t u = h2mul(h2fma(x2, H2_GB, H2_GA), x);
  const uint32_t phi = h2fma(h2tanh(u), H2_HALF, H2_HALF);
  return h2mul(x, phi);
}
__device__ __forceinline__ void gelu_pair_h2(uint32_t x, uint32_t& g, uint32_t& dg) {
  const uint32_t x2 = h2mul(x, x);
  const uint32_t u = h2mul(h2fma(x2, H2_GB, H2_GA), x);
  const uint32_t phi = h2fma(h2tanh(u), H2_HALF, H2_HALF);
  g = h2mul(x, phi);
  const uint32_t e = h2ex2(h2mul(x2, H2_GK));        // exp(-x^2/2)
  dg = h2fma(h2mul(x, H2_GC), e, phi);
}

