// Axial attention forward for SHORT sequences and narrow heads (S <= 96, head_dim <= 32: the S = 80 / hd = 20 stage of the 224^2
// configs), one WARP per (image, head) item, everything in registers:
//   O = softmax(Q K^T / sqrt(hd) + bias[b]) V          (Vi_Tools_CNN_less_V2.py:293-298)
// Why not the tcgen05 kernels of attention_sm100.cu here: an item is 80 x 80 (or 128 x 128) scores — a few hundred kFLOP — and the
// 128-row tile, the TMEM round trip and the mbarrier hand-offs of that design cost ~3.8 us per item with ONE item in flight per SM
// (67 us per launch for 43 MB of HBM traffic; this kernel: 33 us). With a warp per item, 16 items are in flight per SM, nothing is shared between warps
// (no block barrier, no mbarrier) and the score tile never leaves the register file: per 16-query block the warp holds S (16 x S) as
// mma.sync.m16n8k16 accumulator fragments, takes the row maximum / sum with two quad shuffles, and re-uses the exponentiated
// fragments directly as the A operand of P.V (the accumulator layout of an m16n8 tile IS half an A fragment).
// K and V of the item are staged once in warp-private shared memory ([S][32] bf16, pitch 80 bytes: conflict-free ldmatrix), the
// head dim is zero-padded to 32 there; Q rows and bias rows are read straight from global memory in fragment layout.
// Measured limit: at S = 112 / 128 the score fragments need 136 - 148 registers and 72 - 82 KB of shared memory per CTA (8 warps per
// SM) and the kernel is no faster than the tcgen05 one (80 vs 75 us at S = 128), so those shapes stay there.
#include "common.cuh"
#include "../../include/calm_b200.h"
#include "attention_tc.h"
#include <math.h>

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int SM_WARPS = 4;          // warps (items) per CTA; several CTAs per SM
constexpr int PITCH = 40;            // bf16 elements per staged row: 32 dims + 8 pad (80 bytes)

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
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2f(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// rows [0, S) x hd of a token-major matrix (head offset applied) -> dst[S][PITCH], columns hd..31 zero-filled. Asynchronous 8-byte
// copies: all of a warp's K and V units are in flight at once (a load / store loop exposed one global latency per iteration: 40 per item).
__device__ __forceinline__ void stage_rows_async(bf16* dst, const bf16* __restrict__ src, long long ld, int S, int hd, int lane) {
  const int upr = hd >> 2;                       // 8-byte units per row
  for (int u = lane; u < S * 8; u += 32) {
    const int row = u >> 3, c = u & 7;
    const bf16* g = src + (long long)row * ld + (c < upr ? 4 * c : 0);
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + row * PITCH + 4 * c);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(g), "r"(c < upr ? 8 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

struct SmallParams {
  const bf16 *q, *k, *v, *bias;
  bf16* o; float* lse;
  long long ld_q, ld_k, ld_v, ld_o;
  int B, S, heads, hd;
  float scale_log2;
};

// S16 = S / 16 (1..6): the score fragments of a 16-query block are c[2 * S16][4] registers
template <int S16>
__global__ void __launch_bounds__(32 * SM_WARPS, S16 <= 5 ? 4 : 2)
attn_fwd_small_kernel(const SmallParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int S = 16 * S16, NT = 2 * S16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  bf16* sK = reinterpret_cast<bf16*>(smem_raw) + (size_t)warp * 2 * S * PITCH;
  bf16* sV = sK + S * PITCH;
  const int item = blockIdx.x * SM_WARPS + warp;
  if (item >= p.B * p.heads) return;             // warps are independent: no block-level barrier below
  const int b = item / p.heads, h = item - b * p.heads;
  const int hd = p.hd;
  const long long row0 = (long long)b * S;
  const bf16* gq = p.q + row0 * p.ld_q + (long long)h * hd;
  stage_rows_async(sK, p.k + row0 * p.ld_k + (long long)h * hd, p.ld_k, S, hd, lane);
  stage_rows_async(sV, p.v + row0 * p.ld_v + (long long)h * hd, p.ld_v, S, hd, lane);
  cp_async_wait_all();
  __syncwarp();
  const bf16* gbias = p.bias + row0 * S;
  // ldmatrix row addresses: lane l supplies row (l & 7) of matrix (l >> 3)
  const int lm = lane >> 3, lr = lane & 7;
  // Q block as A fragments (2 k-steps over the zero-padded head dim), straight from global memory
  auto load_q = [&](uint32_t (*a)[4], int q0) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {     // dims 16 ks + 8 half + 2t, +1
        const int d = 16 * ks + 8 * half + 2 * t;
        uint32_t lo = 0u, hi = 0u;
        if (d < hd) {
          lo = *reinterpret_cast<const uint32_t*>(gq + (long long)(q0 + g) * p.ld_q + d);
          hi = *reinterpret_cast<const uint32_t*>(gq + (long long)(q0 + g + 8) * p.ld_q + d);
        }
        a[ks][2 * half] = lo; a[ks][2 * half + 1] = hi;
      }
    }
  };
  uint32_t aq[2][4], aq_next[2][4];
  load_q(aq, 0);
  for (int qb = 0; qb < S16; ++qb) {
    const int q0 = qb * 16;
    // the loads of this block's bias rows and of the next block's Q go out before the MMAs: their latency runs under the arithmetic
    if (qb + 1 < S16) load_q(aq_next, q0 + 16);
    const bf16* br0 = gbias + (long long)(q0 + g) * S + 2 * t;
    const bf16* br1 = br0 + 8 * S;
    uint32_t bia[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      bia[n][0] = *reinterpret_cast<const uint32_t*>(br0 + 8 * n);
      bia[n][1] = *reinterpret_cast<const uint32_t*>(br1 + 8 * n);
    }
    // ---- S = Q K^T
    float c[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int n = 0; n < NT; n += 2) {
        // matrices: (keys 8n.., dims 16ks), (keys 8n.., dims 16ks+8), (keys 8(n+1).., dims 16ks), (keys 8(n+1).., dims 16ks+8)
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sm_addr(sK + (8 * (n + (lm >> 1)) + lr) * PITCH + 16 * ks + 8 * (lm & 1)), b0, b1, b2, b3);
        mma16816(c[n], aq[ks], b0, b1);
        mma16816(c[n + 1], aq[ks], b2, b3);
      }
    }
    // ---- x = s * scale*log2e + bias*log2e ; row maxima (rows g and g + 8 of the block)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bia[n][0]));
      const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bia[n][1]));
      c[n][0] = fmaf(c[n][0], p.scale_log2, f0.x * LOG2E); c[n][1] = fmaf(c[n][1], p.scale_log2, f0.y * LOG2E);
      c[n][2] = fmaf(c[n][2], p.scale_log2, f1.x * LOG2E); c[n][3] = fmaf(c[n][3], p.scale_log2, f1.y * LOG2E);
      m0 = fmaxf(m0, fmaxf(c[n][0], c[n][1]));
      m1 = fmaxf(m1, fmaxf(c[n][2], c[n][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    // ---- P = exp2(x - m) (rounded to bf16 as the MMA operand; the row sum is taken over the same rounded values' fp32 originals)
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      c[n][0] = ex2f(c[n][0] - m0); c[n][1] = ex2f(c[n][1] - m0);
      c[n][2] = ex2f(c[n][2] - m1); c[n][3] = ex2f(c[n][3] - m1);
      l0 += c[n][0] + c[n][1];
      l1 += c[n][2] + c[n][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V (4 n-tiles over the padded head dim)
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < S16; ++kk) {
      uint32_t ap[4];
      ap[0] = pack2(c[2 * kk][0], c[2 * kk][1]);         ap[1] = pack2(c[2 * kk][2], c[2 * kk][3]);
      ap[2] = pack2(c[2 * kk + 1][0], c[2 * kk + 1][1]); ap[3] = pack2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
#pragma unroll
      for (int n = 0; n < 4; n += 2) {
        // transposed matrices: (keys 16kk.., dims 8n), (keys 16kk+8.., dims 8n), (keys 16kk.., dims 8(n+1)), (keys 16kk+8.., dims 8(n+1))
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sm_addr(sV + (16 * kk + 8 * (lm & 1) + lr) * PITCH + 8 * (n + (lm >> 1))), b0, b1, b2, b3);
        mma16816(o[n], ap, b0, b1);
        mma16816(o[n + 1], ap, b2, b3);
      }
    }
    // ---- store O / l (bf16 pairs) and lse
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    bf16* o0 = p.o + (row0 + q0 + g) * p.ld_o + (long long)h * hd + 2 * t;
    bf16* o1 = o0 + 8 * p.ld_o;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      if (8 * n + 2 * t < hd) {
        *reinterpret_cast<uint32_t*>(o0 + 8 * n) = pack2(o[n][0] * i0, o[n][1] * i0);
        *reinterpret_cast<uint32_t*>(o1 + 8 * n) = pack2(o[n][2] * i1, o[n][3] * i1);
      }
    }
    if (t == 0) {
      float* ls = p.lse + ((long long)b * p.heads + h) * S + q0 + g;
      ls[0] = (m0 + log2f(l0)) * LN2;
      ls[8] = (m1 + log2f(l1)) * LN2;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int e = 0; e < 4; ++e) aq[ks][e] = aq_next[ks][e];
  }
}

template <int S16>
int launch_small_fwd(const SmallParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)SM_WARPS * 2 * (16 * S16) * PITCH * sizeof(bf16);
  static CalmDeviceOnce configured;
  if (smem > 48 * 1024 && configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_small_kernel<S16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { calm_set_error("calm_attention_fwd(small): smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  const int items = p.B * p.heads;
  attn_fwd_small_kernel<S16><<<(items + SM_WARPS - 1) / SM_WARPS, 32 * SM_WARPS, smem, stream>>>(p);
  CALM_CHECK_LAUNCH("calm_attention_fwd(small)");
  return CALM_OK;
}

}  // namespace

bool calm_attention_small_eligible(int B, int S, int heads, int hd, const int64_t* lds, int nlds, const void* const* ptrs, int nptrs) {
  if (B <= 0 || heads <= 0 || S < 16 || S > 96 || (S & 15) || hd < 4 || hd > 32 || (hd & 3)) return false;
  for (int i = 0; i < nlds; ++i)
    if (lds[i] % 4) return false;                       // 8-byte row units
  for (int i = 0; i < nptrs; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 7) return false;
  return true;
}

int calm_attention_fwd_small(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
                             int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream) {
  SmallParams p;
  p.q = reinterpret_cast<const bf16*>(q); p.k = reinterpret_cast<const bf16*>(k); p.v = reinterpret_cast<const bf16*>(v);
  p.bias = reinterpret_cast<const bf16*>(bias); p.o = reinterpret_cast<bf16*>(o); p.lse = lse;
  p.ld_q = ld_q; p.ld_k = ld_k; p.ld_v = ld_v; p.ld_o = ld_o;
  p.B = B; p.S = S; p.heads = heads; p.hd = hd;
  p.scale_log2 = LOG2E / sqrtf((float)hd);
  switch (S / 16) {
    case 1: return launch_small_fwd<1>(p, stream);
    case 2: return launch_small_fwd<2>(p, stream);
    case 3: return launch_small_fwd<3>(p, stream);
    case 4: return launch_small_fwd<4>(p, stream);
    case 5: return launch_small_fwd<5>(p, stream);
    default: return launch_small_fwd<6>(p, stream);
  }
}
