// Fused per-Block CNN residual on the (B, S, S, 3) token image (channels-last = the row-token layout itself):
//   y = x + conv1x1_{32->3}( gelu( dwconv3x3( gelu( conv1x1_{3->32}(x) ) ) ) )
// replaces the Sequential(sn(Conv2d 3->32,k1), GELU, sn(Conv2d 32->32,k3,groups=32,pad=1), GELU, sn(Conv2d 32->3,k1)) and the
// permutes around it (Vi_Tools_CNN_less_V2.py:378-385,400-403; CALM_ViT_V2.py:60-67,80-83). The reference materialises
// four (B,32,S,S) tensors per call; here the 32-channel intermediates only ever live in shared memory, so HBM traffic is
// the 3-channel read + 3-channel write (forward) — the backward recomputes them from x with a 2-pixel halo.
//
// v3 (round 2). Round 1's kernels issued ~110 (forward) / ~240 (backward) instructions per pixel-channel for ~45 / ~90 of fp32
// arithmetic (ncu: issue-bound, 2 % of the HBM roofline). This version changes the arithmetic, not just the schedule:
//   * CHANNEL PAIRS IN fp16x2. The hidden activations (pre1, h1 = gelu(pre1), pre2, h2 = gelu(pre2) and the two GELU
//     derivatives) are O(1) quantities that the reference's autocast path stores in bf16 (8-bit mantissa); here they are
//     computed and held as packed half2 (11-bit mantissa) with one channel pair per register: one HFMA2 does the work of two
//     FFMAs and one MUFU op serves two channels. GELU(erf) in half2: Phi(x) = 0.5 + 0.5 tanh(x (a + b x^2)), a = 0.79880144,
//     b = 0.03528205 (minimax fit of atanh(erf(x/sqrt2)), |error| < 2.9e-4 — below the fp16 resolution of Phi; both
//     coefficients positive, so the polynomial saturates monotonically and needs no clamp), gelu' = Phi + x exp(-x^2/2)/sqrt(2 pi):
//     5 HFMA2-class + 1 MUFU (value) or 8 + 2 MUFU (value and derivative) per channel PAIR.
//     Everything on the gradient side (dy, dp2, dp1, all parameter-gradient sums, dx) stays fp32 / bf16-with-fp32-accumulate:
//     the GradScaler-scaled gradients do not fit fp16's range.
//   * A WARP OWNS ONE CHANNEL PAIR for the life of the CTA (16 warps = 32 channels). Its planes (h1, g1/dp1, dp2 of the tile)
//     are private, so the three phases of a tile only need __syncwarp between them; its weights live in registers; its 34
//     parameter-gradient sums are per-lane register accumulators across ALL tiles of the persistent CTA (no per-tile / per-channel
//     shuffle reductions: one butterfly per kernel). A lane owns one pixel column of the 32-wide tile and walks down the rows
//     with a rotating 3-row register window (3 new taps per pixel instead of 9).
//   * THE TWO CONTRACTIONS OVER CHANNELS RUN ON TENSOR CORES. y = W3 h2 (forward) and dx = W1^T dp1 (backward) contract over
//     the 32 channels that live in 16 different warps: the per-warp planes [channel pair][pixel] ARE the A-fragment layout of
//     mma.sync.m16n8k16 (row = pixel, k = channel; one LDS.32 per fragment register, conflict-free because the warp stride is
//     8 mod 32 words), so after a block barrier each warp finishes 16-pixel segments with 2 MMAs (fp16 x fp16 forward,
//     bf16 x bf16 backward, fp32 accumulate), adds the fp32 residual / dy in shared memory and writes 192-byte coalesced rows.
//   * Tiles are prefetched with cp.async (zero-fill outside the image) one tile ahead, so the global latency of the next
//     tile's x / dy is hidden behind the current tile's arithmetic.
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

constexpr int NWARP = 16;
constexpr int NT = NWARP * 32;
constexpr int TW = 32;            // tile width = one lane per pixel column
constexpr int RAW_COLS = 36;      // raw tile rows hold image columns tx0-2 .. tx0+33 (even start: 8-byte cp.async chunks)
constexpr int RAW_ROW = RAW_COLS * 3;
constexpr int NPARAM = CALM_CNN_NPARAM;   // w1[96] b1[32] w2[288] b2[32] w3[96] b3[3]
constexpr int P_W1 = 0, P_B1 = 96, P_W2 = 128, P_B2 = 416, P_W3 = 448, P_B3 = 544;
constexpr int WSM_FLOATS = 552;

struct CnnW { const float *w1, *b1, *w2, *b2, *w3, *b3; };

__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Asynchronous copy of image rows [gy0, gy0 + nrows) x columns [gx0, gx0 + 36) (3 floats per pixel) into dst[nrows][108];
// everything outside the S x S image is zero-filled. gx0 is even; with S even a 2-float chunk never straddles the image border.
// The (row, chunk) pairs a thread copies do not depend on the tile: they are decoded once per kernel into `slots` (row << 8 | chunk),
// so a tile's prefetch costs ~10 instructions per 8-byte chunk instead of ~40 (the tile-top bookkeeping was 10-25 % of all
// instructions of the first version of these kernels).
constexpr int PF_SLOTS = 3;
struct Prefetch { int slot[PF_SLOTS]; };
__device__ __forceinline__ Prefetch make_prefetch(int nrows, bool even) {
  Prefetch pf;
  const int per_row = even ? RAW_ROW / 2 : RAW_ROW;
#pragma unroll
  for (int k = 0; k < PF_SLOTS; ++k) {
    const int i = threadIdx.x + k * NT;
    const int row = i / per_row;
    pf.slot[k] = row < nrows ? (row << 8 | (i - row * per_row)) : -1;
  }
  return pf;
}
__device__ __forceinline__ void prefetch_rows(const Prefetch& pf, float* dst, const float* __restrict__ img, int S, int gy0, int gx0, int nrows) {
  const int lim = S * 3, f0 = gx0 * 3;
  if (!(S & 1)) {
#pragma unroll
    for (int k = 0; k < PF_SLOTS; ++k) {
      if (pf.slot[k] >= 0) {
        const int row = pf.slot[k] >> 8, ch = pf.slot[k] & 255;
        const int gy = gy0 + row, f = f0 + 2 * ch;
        const bool ok = (unsigned)gy < (unsigned)S && (unsigned)f < (unsigned)lim;
        cp_async8(dst + row * RAW_ROW + 2 * ch, ok ? img + (long long)gy * lim + f : img, ok ? 8 : 0);
      }
    }
  } else {   // odd S (not a trainer shape): 4-byte elements, plain loop
    for (int i = threadIdx.x; i < nrows * RAW_ROW; i += NT) {
      const int row = i / RAW_ROW, e = i - row * RAW_ROW;
      const int gy = gy0 + row, f = f0 + e;
      const bool ok = gy >= 0 && gy < S && f >= 0 && f < lim;
      cp_async4(dst + row * RAW_ROW + e, ok ? img + (long long)gy * lim + f : img, ok ? 4 : 0);
    }
  }
}

__device__ __forceinline__ void load_weights(float* wsm, const CnnW& W) {
  for (int i = threadIdx.x; i < NPARAM; i += NT) {
    float v;
    if (i < P_B1) v = W.w1[i];
    else if (i < P_W2) v = W.b1[i - P_B1];
    else if (i < P_B2) v = W.w2[i - P_W2];
    else if (i < P_W3) v = W.b2[i - P_B2];
    else if (i < P_B3) v = W.w3[i - P_W3];
    else v = W.b3[i - P_B3];
    wsm[i] = v;
  }
}

// One 16-pixel row segment leaves shared memory (48 floats at seg, 8-byte aligned) for global memory: 24 lanes x float2 when the
// segment lies inside the image and rows are 8-byte aligned (S even), element-wise guarded otherwise.
__device__ __forceinline__ void store_segment(const float* seg, float* __restrict__ out, bf16* __restrict__ out16, long long row_base /* pixel index of column 0 of this image row */,
                                              int gx0, int S, int lane) {
  if (lane >= 24) return;
  const long long o = (row_base + gx0) * 3 + 2 * lane;
  if (!(S & 1) && gx0 + 16 <= S) {
    const float2 v = *reinterpret_cast<const float2*>(seg + 2 * lane);
    *reinterpret_cast<float2*>(out + o) = v;
    if (out16) *reinterpret_cast<bf162*>(out16 + o) = __floats2bfloat162_rn(v.x, v.y);
  } else {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int f = 2 * lane + e;
      if (gx0 + f / 3 < S) {
        out[o + e] = seg[f];
        if (out16) out16[o + e] = __float2bfloat16(seg[f]);
      }
    }
  }
}

// ======================================================================================================================
// forward: 32 x th tiles (th <= 16), 2 CTAs / SM
// ======================================================================================================================
constexpr int THF = 16;
constexpr int F_XIN_COLS = 34;                                 // image columns tx0-1 .. tx0+32
constexpr int F_XIN_BYTES = (THF + 2) * F_XIN_COLS * 16;       // {x0,x0 | x1,x1 | x2,x2 | mask} as half2 per pixel
constexpr int F_RAW_BYTES = (THF + 2) * RAW_ROW * 4;           // fp32 x rows ty0-1 .. ty0+th, double-buffered
constexpr int F_H1 = 0, F_H2 = (THF + 2) * F_XIN_COLS;         // per-warp plane offsets (words)
constexpr int F_WARP_WORDS = 1128;                             // 612 + 512 = 1124, padded to 8 mod 32
constexpr int FWD_SMEM = F_XIN_BYTES + 2 * F_RAW_BYTES + NWARP * F_WARP_WORDS * 4 + WSM_FLOATS * 4;
static_assert(F_WARP_WORDS % 32 == 8 && F_WARP_WORDS >= F_H2 + THF * TW, "forward plane stride");

__global__ void __launch_bounds__(NT, 2)
cnn_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, CnnW W, int B, int S, int tiles_x, int tiles_y, int th) {
  extern __shared__ __align__(16) unsigned char smraw[];
  uint4* xin = reinterpret_cast<uint4*>(smraw);
  float* raw0 = reinterpret_cast<float*>(smraw + F_XIN_BYTES);
  uint32_t* planes = reinterpret_cast<uint32_t*>(smraw + F_XIN_BYTES + 2 * F_RAW_BYTES);
  float* wsm = reinterpret_cast<float*>(planes + NWARP * F_WARP_WORDS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  load_weights(wsm, W);
  uint32_t* h1 = planes + warp * F_WARP_WORDS + F_H1;
  uint32_t* h2p = planes + warp * F_WARP_WORDS + F_H2;
  const int tiles_per_img = tiles_x * tiles_y;
  const int total = B * tiles_per_img;          // < 2^31 (checked by the host)
  int tile = blockIdx.x;
  const Prefetch pf = make_prefetch(th + 2, !(S & 1));
  auto issue = [&](int tl, int buf) {
    const int b = tl / tiles_per_img, tr = tl - b * tiles_per_img;
    const int trow = tr / tiles_x;
    prefetch_rows(pf, raw0 + buf * (F_RAW_BYTES / 4), x + (long long)b * S * S * 3, S, trow * th - 1, (tr - trow * tiles_x) * TW - 2, th + 2);
  };
  if (tile < total) issue(tile, 0);
  cp_async_commit();
  __syncthreads();
  const int ca = 2 * warp, cb = ca + 1;
  // this warp's weights (registers for the life of the CTA)
  uint32_t w1h[3], w2h[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) w1h[k] = h2pack(wsm[P_W1 + ca * 3 + k], wsm[P_W1 + cb * 3 + k]);
#pragma unroll
  for (int q = 0; q < 9; ++q) w2h[q] = h2pack(wsm[P_W2 + ca * 9 + q], wsm[P_W2 + cb * 9 + q]);
  const uint32_t b1h = h2pack(wsm[P_B1 + ca], wsm[P_B1 + cb]), b2h = h2pack(wsm[P_B2 + ca], wsm[P_B2 + cb]);
  // B fragments of y = W3 h2: B[k = channel][n = output channel] = w3[n][channel], zero for n >= 3
  uint32_t bw[2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    bw[ks][0] = g < 3 ? h2pack(wsm[P_W3 + g * 32 + 16 * ks + 2 * t], wsm[P_W3 + g * 32 + 16 * ks + 2 * t + 1]) : 0u;
    bw[ks][1] = g < 3 ? h2pack(wsm[P_W3 + g * 32 + 16 * ks + 2 * t + 8], wsm[P_W3 + g * 32 + 16 * ks + 2 * t + 9]) : 0u;
  }
  const float b3a = wsm[P_B3 + 0], b3b = wsm[P_B3 + 1], b3c = wsm[P_B3 + 2];

  for (int it = 0; tile < total; tile += gridDim.x, ++it) {
    const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
    const int trow = tr / tiles_x;
    const int ty0 = trow * th, tx0 = (tr - trow * tiles_x) * TW;
    float* raw = raw0 + (it & 1) * (F_RAW_BYTES / 4);
    cp_async_wait_all();
    __syncthreads();                         // this tile's rows have landed; the previous tile is fully consumed
    if (tile + gridDim.x < total) issue(tile + gridDim.x, (it + 1) & 1);
    cp_async_commit();
    for (int i = tid; i < (th + 2) * F_XIN_COLS; i += NT) {
      const int row = i / F_XIN_COLS, col = i - row * F_XIN_COLS;
      const int gy = ty0 - 1 + row, gx = tx0 - 1 + col;
      const bool inside = gy >= 0 && gy < S && gx >= 0 && gx < S;
      const float* p = raw + row * RAW_ROW + (col + 1) * 3;
      uint4 e;
      e.x = h2pack(p[0], p[0]); e.y = h2pack(p[1], p[1]); e.z = h2pack(p[2], p[2]); e.w = inside ? H2_ONE : 0u;
      xin[i] = e;
    }
    __syncthreads();
    // ---- phase B: h1 = gelu(conv1x1(x)) on the 1-pixel halo region, zero outside the image (the dwconv's zero padding)
    for (int rr = 0; rr < th + 2; ++rr) {
      const uint4 e = xin[rr * F_XIN_COLS + lane];
      const uint32_t pre = h2fma(w1h[0], e.x, h2fma(w1h[1], e.y, h2fma(w1h[2], e.z, b1h)));
      h1[rr * F_XIN_COLS + lane] = h2mul(gelu_h2(pre), e.w);
    }
    for (int rr = lane >> 1; rr < th + 2; rr += 16) {
      const int cc = 32 + (lane & 1);
      const uint4 e = xin[rr * F_XIN_COLS + cc];
      const uint32_t pre = h2fma(w1h[0], e.x, h2fma(w1h[1], e.y, h2fma(w1h[2], e.z, b1h)));
      h1[rr * F_XIN_COLS + cc] = h2mul(gelu_h2(pre), e.w);
    }
    __syncwarp();
    // ---- phase C: h2 = gelu(dwconv3x3(h1)) for pixel column `lane`, rotating 3-row window
    {
      uint32_t r0[3], r1[3], r2[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) { r0[j] = h1[lane + j]; r1[j] = h1[F_XIN_COLS + lane + j]; }
      auto step = [&](const uint32_t (&a)[3], const uint32_t (&bq)[3], uint32_t (&c)[3], int r) {
#pragma unroll
        for (int j = 0; j < 3; ++j) c[j] = h1[(r + 2) * F_XIN_COLS + lane + j];
        uint32_t s0 = h2fma(w2h[0], a[0], b2h), s1 = h2mul(w2h[3], bq[0]), s2 = h2mul(w2h[6], c[0]);
        s0 = h2fma(w2h[1], a[1], s0); s1 = h2fma(w2h[4], bq[1], s1); s2 = h2fma(w2h[7], c[1], s2);
        s0 = h2fma(w2h[2], a[2], s0); s1 = h2fma(w2h[5], bq[2], s1); s2 = h2fma(w2h[8], c[2], s2);
        const uint32_t pre = h2fma(H2_ONE, s0, h2fma(H2_ONE, s1, s2));
        h2p[r * TW + lane] = gelu_h2(pre);
      };
      for (int r = 0; r < th; r += 3) {
        step(r0, r1, r2, r);
        if (r + 1 < th) step(r1, r2, r0, r + 1);
        if (r + 2 < th) step(r2, r0, r1, r + 2);
      }
    }
    __syncthreads();
    // ---- y = x + b3 + W3 h2 : 16-pixel segments on the tensor cores, finished in the raw x tile and stored coalesced
    for (int mt = warp; mt < th * 2; mt += NWARP) {
      const int r = mt >> 1, c0 = (mt & 1) * 16;
      const int gy = ty0 + r;
      if (gy >= S || tx0 + c0 >= S) continue;
      const uint32_t* pb = planes + F_H2 + r * TW + c0 + g;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t* q = pb + (8 * ks + t) * F_WARP_WORDS;
        mma_f16(d, q[0], q[8], q[4 * F_WARP_WORDS], q[4 * F_WARP_WORDS + 8], bw[ks][0], bw[ks][1]);
      }
      float* seg = raw + (r + 1) * RAW_ROW + (c0 + 2) * 3;
      if (t == 0) {
        seg[g * 3 + 0] += d[0] + b3a; seg[g * 3 + 1] += d[1] + b3b;
        seg[(g + 8) * 3 + 0] += d[2] + b3a; seg[(g + 8) * 3 + 1] += d[3] + b3b;
      } else if (t == 1) {
        seg[g * 3 + 2] += d[0] + b3c; seg[(g + 8) * 3 + 2] += d[2] + b3c;
      }
      __syncwarp();
      store_segment(seg, y, nullptr, ((long long)b * S + gy) * S, tx0 + c0, S, lane);
    }
  }
  cp_async_wait_all();
}

// ======================================================================================================================
// backward: 32 x th tiles (th <= 24), 1 CTA / SM
//   phase B: h1 = gelu(pre1), g1 = gelu'(pre1) on the 2-pixel halo            (recomputed from x, fp16x2)
//   phase C: pre2 -> h2, d2 ; dp2 = (W3^T dy) * d2 on the 1-pixel halo ; dW3, db2, dW2 from the owned pixels
//   phase D: dh1 = dwconv^T(dp2), dp1 = dh1 * g1 ; dW1, db1 ; dp1 -> plane (bf16x2, over g1)
//   final  : dx = dy + W1^T dp1 on the tensor cores ; db3
// ======================================================================================================================
constexpr int THB = 24;
constexpr int B_XIN_COLS = 36;                                  // image columns tx0-2 .. tx0+33
constexpr int B_XIN_BYTES = (THB + 4) * B_XIN_COLS * 16;
constexpr int B_XRAW_BYTES = (THB + 4) * RAW_ROW * 4;           // staging of the next tile's x rows ty0-2 .. ty0+th+1
constexpr int B_DY_BYTES = (THB + 2) * RAW_ROW * 4;             // dy rows ty0-1 .. ty0+th (36 columns), double-buffered
constexpr int B_H1 = 0, B_G1 = (THB + 4) * B_XIN_COLS, B_DP2 = B_G1 + THB * TW;
constexpr int DP2_COLS = 34;                                    // image columns tx0-1 .. tx0+32
constexpr int B_WARP_WORDS = 2664;                              // 1008 + 768 + 884 = 2660, padded to 8 mod 32
constexpr int BWD_SMEM = B_XIN_BYTES + B_XRAW_BYTES + 2 * B_DY_BYTES + NWARP * B_WARP_WORDS * 4 + WSM_FLOATS * 4 + NWARP * 4 * 4;
static_assert(B_WARP_WORDS % 32 == 8 && B_WARP_WORDS >= B_DP2 + (THB + 2) * DP2_COLS, "backward plane stride");
static_assert(BWD_SMEM <= 227 * 1024, "backward shared memory");

struct RowC { uint32_t h[3]; float lo[3], hi[3]; };   // 3 taps of one h1 row: packed pair + unpacked channels a / b
struct RowD { float lo[3], hi[3]; };                  // 3 taps of one dp2 row

__global__ void __launch_bounds__(NT, 1)
cnn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, bf16* __restrict__ dx16, CnnW W,
               float* __restrict__ gpartial, int B, int S, int tiles_x, int tiles_y, int th) {
  extern __shared__ __align__(16) unsigned char smraw[];
  uint4* xin = reinterpret_cast<uint4*>(smraw);
  float* xraw = reinterpret_cast<float*>(smraw + B_XIN_BYTES);
  float* dy0 = reinterpret_cast<float*>(smraw + B_XIN_BYTES + B_XRAW_BYTES);
  uint32_t* planes = reinterpret_cast<uint32_t*>(smraw + B_XIN_BYTES + B_XRAW_BYTES + 2 * B_DY_BYTES);
  float* wsm = reinterpret_cast<float*>(planes + NWARP * B_WARP_WORDS);
  float* wred = wsm + WSM_FLOATS;   // [NWARP][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  load_weights(wsm, W);
  uint32_t* h1 = planes + warp * B_WARP_WORDS + B_H1;
  uint32_t* g1 = planes + warp * B_WARP_WORDS + B_G1;
  uint32_t* dp2 = planes + warp * B_WARP_WORDS + B_DP2;
  const int tiles_per_img = tiles_x * tiles_y;
  const int total = B * tiles_per_img;          // < 2^31 (checked by the host)
  int tile = blockIdx.x;
  const Prefetch pfx = make_prefetch(th + 4, !(S & 1)), pfd = make_prefetch(th + 2, !(S & 1));
  auto issue = [&](int tl, int buf) {
    const int b = tl / tiles_per_img, tr = tl - b * tiles_per_img;
    const int trow = tr / tiles_x;
    const int ty0 = trow * th, tx0 = (tr - trow * tiles_x) * TW;
    const long long img = (long long)b * S * S * 3;
    prefetch_rows(pfx, xraw, x + img, S, ty0 - 2, tx0 - 2, th + 4);
    prefetch_rows(pfd, dy0 + buf * (B_DY_BYTES / 4), dy + img, S, ty0 - 1, tx0 - 2, th + 2);
  };
  if (tile < total) issue(tile, 0);
  cp_async_commit();
  __syncthreads();
  const int ca = 2 * warp, cb = ca + 1;
  // per-lane parameter-gradient sums of this warp's two channels, accumulated over every tile of this CTA
  float gw1a[3] = {0.f, 0.f, 0.f}, gw1b[3] = {0.f, 0.f, 0.f}, gb1a = 0.f, gb1b = 0.f;
  float gw2a[9], gw2b[9], gb2a = 0.f, gb2b = 0.f;
  float gw3a[3] = {0.f, 0.f, 0.f}, gw3b[3] = {0.f, 0.f, 0.f};
  float gb3x = 0.f, gb3y = 0.f;     // t == 0 lanes: output channels 0 / 1 ; t == 1 lanes: gb3x = output channel 2
#pragma unroll
  for (int q = 0; q < 9; ++q) gw2a[q] = gw2b[q] = 0.f;
  // B fragments of dx = W1^T dp1: B[k = channel][n = input channel] = w1[channel][n], zero for n >= 3
  uint32_t bw[2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    bw[ks][0] = g < 3 ? bf2pack(wsm[P_W1 + (16 * ks + 2 * t) * 3 + g], wsm[P_W1 + (16 * ks + 2 * t + 1) * 3 + g]) : 0u;
    bw[ks][1] = g < 3 ? bf2pack(wsm[P_W1 + (16 * ks + 2 * t + 8) * 3 + g], wsm[P_W1 + (16 * ks + 2 * t + 9) * 3 + g]) : 0u;
  }

  for (int it = 0; tile < total; tile += gridDim.x, ++it) {
    const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
    const int trow = tr / tiles_x;
    const int ty0 = trow * th, tx0 = (tr - trow * tiles_x) * TW;
    float* dys = dy0 + (it & 1) * (B_DY_BYTES / 4);
    cp_async_wait_all();
    __syncthreads();                         // x / dy of this tile have landed; the previous tile is fully consumed
    for (int i = tid; i < (th + 4) * B_XIN_COLS; i += NT) {
      const int row = i / B_XIN_COLS, col = i - row * B_XIN_COLS;
      const int gy = ty0 - 2 + row, gx = tx0 - 2 + col;
      const bool inside = gy >= 0 && gy < S && gx >= 0 && gx < S;
      const float* p = xraw + row * RAW_ROW + col * 3;
      uint4 e;
      e.x = h2pack(p[0], p[0]); e.y = h2pack(p[1], p[1]); e.z = h2pack(p[2], p[2]); e.w = inside ? H2_ONE : 0u;
      xin[i] = e;
    }
    __syncthreads();                         // xin complete, the x staging buffer is free again
    if (tile + gridDim.x < total) issue(tile + gridDim.x, (it + 1) & 1);
    cp_async_commit();

    // ---------------- phase B
    {
      uint32_t w1h[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) w1h[k] = h2pack(wsm[P_W1 + ca * 3 + k], wsm[P_W1 + cb * 3 + k]);
      const uint32_t b1h = h2pack(wsm[P_B1 + ca], wsm[P_B1 + cb]);
      for (int rr = 0; rr < th + 4; ++rr) {
        const uint4 e = xin[rr * B_XIN_COLS + lane];
        const uint32_t pre = h2fma(w1h[0], e.x, h2fma(w1h[1], e.y, h2fma(w1h[2], e.z, b1h)));
        uint32_t hv, gv;
        gelu_pair_h2(pre, hv, gv);
        h1[rr * B_XIN_COLS + lane] = h2mul(hv, e.w);
        if (rr >= 2 && rr < th + 2 && lane >= 2) g1[(rr - 2) * TW + lane - 2] = h2mul(gv, e.w);
      }
      for (int rr = lane >> 2; rr < th + 4; rr += 8) {
        const int cc = 32 + (lane & 3);
        const uint4 e = xin[rr * B_XIN_COLS + cc];
        const uint32_t pre = h2fma(w1h[0], e.x, h2fma(w1h[1], e.y, h2fma(w1h[2], e.z, b1h)));
        uint32_t hv, gv;
        gelu_pair_h2(pre, hv, gv);
        h1[rr * B_XIN_COLS + cc] = h2mul(hv, e.w);
        if (rr >= 2 && rr < th + 2 && cc < 34) g1[(rr - 2) * TW + cc - 2] = h2mul(gv, e.w);
      }
    }
    __syncwarp();
    // ---------------- phase C : tile rows rloc = -1 .. th, pixel column `lane`
    {
      uint32_t w2h[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) w2h[q] = h2pack(wsm[P_W2 + ca * 9 + q], wsm[P_W2 + cb * 9 + q]);
      const uint32_t b2h = h2pack(wsm[P_B2 + ca], wsm[P_B2 + cb]);
      float w3a[3], w3b[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { w3a[k] = wsm[P_W3 + k * 32 + ca]; w3b[k] = wsm[P_W3 + k * 32 + cb]; }
      auto pre2 = [&](const uint32_t (&a)[3], const uint32_t (&bq)[3], const uint32_t (&c)[3]) {
        uint32_t s0 = h2fma(w2h[0], a[0], b2h), s1 = h2mul(w2h[3], bq[0]), s2 = h2mul(w2h[6], c[0]);
        s0 = h2fma(w2h[1], a[1], s0); s1 = h2fma(w2h[4], bq[1], s1); s2 = h2fma(w2h[7], c[1], s2);
        s0 = h2fma(w2h[2], a[2], s0); s1 = h2fma(w2h[5], bq[2], s1); s2 = h2fma(w2h[8], c[2], s2);
        return h2fma(H2_ONE, s0, h2fma(H2_ONE, s1, s2));
      };
      auto load_row = [&](RowC& R, int prow) {     // h1 plane row prow, image columns lane-1 .. lane+1
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          R.h[j] = h1[prow * B_XIN_COLS + lane + 1 + j];
          const float2 f = h2unpack(R.h[j]);
          R.lo[j] = f.x; R.hi[j] = f.y;
        }
      };
      RowC ra, rb, rc;
      load_row(ra, 0);
      load_row(rb, 1);
      auto step = [&](const RowC& A, const RowC& Bq, RowC& C, int rloc) {
        load_row(C, rloc + 3);
        uint32_t h2v, d2v;
        gelu_pair_h2(pre2(A.h, Bq.h, C.h), h2v, d2v);
        const float* dp = dys + ((rloc + 1) * RAW_COLS + lane + 2) * 3;
        const float d0 = dp[0], d1 = dp[1], d2y = dp[2];
        const float2 d2f = h2unpack(d2v);
        const float pa = (w3a[0] * d0 + w3a[1] * d1 + w3a[2] * d2y) * d2f.x;
        const float pb = (w3b[0] * d0 + w3b[1] * d1 + w3b[2] * d2y) * d2f.y;
        dp2[(rloc + 1) * DP2_COLS + lane + 1] = bf2pack(pa, pb);
        if (rloc >= 0 && rloc < th) {          // owned rows (warp-uniform); pixels outside the image have dy = 0
          const float2 hf = h2unpack(h2v);
          gb2a += pa; gb2b += pb;
          gw3a[0] = fmaf(d0, hf.x, gw3a[0]); gw3a[1] = fmaf(d1, hf.x, gw3a[1]); gw3a[2] = fmaf(d2y, hf.x, gw3a[2]);
          gw3b[0] = fmaf(d0, hf.y, gw3b[0]); gw3b[1] = fmaf(d1, hf.y, gw3b[1]); gw3b[2] = fmaf(d2y, hf.y, gw3b[2]);
          // dW2 from the UNROUNDED fp32 dp2 of this pixel. (Accumulating it in phase D from the bf16 dp2 plane — the taps are
          // unpacked there anyway — saved ~3 instructions per pixel-channel, but the rounding noise of dp2 survives in this
          // cancellation-dominated sum: the model-level distance of these 288-element gradients to the fp32 truth doubled.)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            gw2a[j] = fmaf(pa, A.lo[j], gw2a[j]); gw2a[3 + j] = fmaf(pa, Bq.lo[j], gw2a[3 + j]); gw2a[6 + j] = fmaf(pa, C.lo[j], gw2a[6 + j]);
            gw2b[j] = fmaf(pb, A.hi[j], gw2b[j]); gw2b[3 + j] = fmaf(pb, Bq.hi[j], gw2b[3 + j]); gw2b[6 + j] = fmaf(pb, C.hi[j], gw2b[6 + j]);
          }
        }
      };
      for (int rloc = -1; rloc <= th; rloc += 3) {
        step(ra, rb, rc, rloc);
        if (rloc + 1 <= th) step(rb, rc, ra, rloc + 1);
        if (rloc + 2 <= th) step(rc, ra, rb, rloc + 2);
      }
      // the two halo columns (image columns tx0-1 and tx0+32), one row per lane: only dp2 is needed there
      if (lane < th + 2) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int pc = side ? 33 : 0;          // h1 plane columns pc .. pc+2 ; dp2 column side ? 33 : 0 ; dy column index pc+1
          uint32_t a[3], bq[3], c[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            a[j] = h1[lane * B_XIN_COLS + pc + j]; bq[j] = h1[(lane + 1) * B_XIN_COLS + pc + j]; c[j] = h1[(lane + 2) * B_XIN_COLS + pc + j];
          }
          uint32_t h2v, d2v;
          gelu_pair_h2(pre2(a, bq, c), h2v, d2v);
          const float* dp = dys + (lane * RAW_COLS + pc + 1) * 3;
          const float d0 = dp[0], d1 = dp[1], d2y = dp[2];
          const float2 d2f = h2unpack(d2v);
          dp2[lane * DP2_COLS + pc] = bf2pack((w3a[0] * d0 + w3a[1] * d1 + w3a[2] * d2y) * d2f.x, (w3b[0] * d0 + w3b[1] * d1 + w3b[2] * d2y) * d2f.y);
        }
      }
    }
    __syncwarp();
    // ---------------- phase D : tile rows 0 .. th-1, pixel column `lane`
    {
      float w2a[9], w2b[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) { w2a[q] = wsm[P_W2 + ca * 9 + q]; w2b[q] = wsm[P_W2 + cb * 9 + q]; }
      auto load_row = [&](RowD& R, int prow) {     // dp2 plane row prow (image row ty0 + prow - 1), image columns lane-1 .. lane+1
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint32_t v = dp2[prow * DP2_COLS + lane + j];
          R.lo[j] = __uint_as_float(v << 16); R.hi[j] = __uint_as_float(v & 0xffff0000u);
        }
      };
      RowD ra, rb, rc;
      load_row(ra, 0);
      load_row(rb, 1);
      auto step = [&](const RowD& A, const RowD& Bq, RowD& C, int r) {   // A = row r-1, Bq = row r, C = row r+1
        load_row(C, r + 2);
        // h1(p) fed out(p - (ky-1, kx-1)) with weight w2[ky][kx]: row r+1-ky, column tap 2-kx
        float da = 0.f, db = 0.f;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          da = fmaf(w2a[kx], C.lo[2 - kx], da); da = fmaf(w2a[3 + kx], Bq.lo[2 - kx], da); da = fmaf(w2a[6 + kx], A.lo[2 - kx], da);
          db = fmaf(w2b[kx], C.hi[2 - kx], db); db = fmaf(w2b[3 + kx], Bq.hi[2 - kx], db); db = fmaf(w2b[6 + kx], A.hi[2 - kx], db);
        }
        const float2 gf = h2unpack(g1[r * TW + lane]);       // gelu'(pre1), zero outside the image
        const float pa = da * gf.x, pb = db * gf.y;
        const uint4 e = xin[(r + 2) * B_XIN_COLS + lane + 2];
        const float x0 = h2low(e.x), x1 = h2low(e.y), x2 = h2low(e.z);
        gw1a[0] = fmaf(pa, x0, gw1a[0]); gw1a[1] = fmaf(pa, x1, gw1a[1]); gw1a[2] = fmaf(pa, x2, gw1a[2]);
        gw1b[0] = fmaf(pb, x0, gw1b[0]); gw1b[1] = fmaf(pb, x1, gw1b[1]); gw1b[2] = fmaf(pb, x2, gw1b[2]);
        gb1a += pa; gb1b += pb;
        g1[r * TW + lane] = bf2pack(pa, pb);                 // dp1 replaces g1 (same lane, same word): the MMA operand
      };
      for (int r = 0; r < th; r += 3) {
        step(ra, rb, rc, r);
        if (r + 1 < th) step(rb, rc, ra, r + 1);
        if (r + 2 < th) step(rc, ra, rb, r + 2);
      }
    }
    __syncthreads();
    // ---------------- dx = dy + W1^T dp1 (tensor cores), db3 = sum dy
    for (int mt = warp; mt < th * 2; mt += NWARP) {
      const int r = mt >> 1, c0 = (mt & 1) * 16;
      const int gy = ty0 + r;
      if (gy >= S || tx0 + c0 >= S) continue;
      const uint32_t* pbase = planes + B_G1 + r * TW + c0 + g;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t* q = pbase + (8 * ks + t) * B_WARP_WORDS;
        mma_bf16(d, q[0], q[8], q[4 * B_WARP_WORDS], q[4 * B_WARP_WORDS + 8], bw[ks][0], bw[ks][1]);
      }
      float* seg = dys + ((r + 1) * RAW_COLS + c0 + 2) * 3;
      if (t == 0) {
        const float v0 = seg[g * 3], v1 = seg[g * 3 + 1], v2 = seg[(g + 8) * 3], v3 = seg[(g + 8) * 3 + 1];
        gb3x += v0 + v2; gb3y += v1 + v3;
        seg[g * 3] = v0 + d[0]; seg[g * 3 + 1] = v1 + d[1]; seg[(g + 8) * 3] = v2 + d[2]; seg[(g + 8) * 3 + 1] = v3 + d[3];
      } else if (t == 1) {
        const float v0 = seg[g * 3 + 2], v2 = seg[(g + 8) * 3 + 2];
        gb3x += v0 + v2;
        seg[g * 3 + 2] = v0 + d[0]; seg[(g + 8) * 3 + 2] = v2 + d[2];
      }
      __syncwarp();
      store_segment(seg, dx, dx16, ((long long)b * S + gy) * S, tx0 + c0, S, lane);
    }
  }
  cp_async_wait_all();
  // ---------------- one partial row per CTA, in the public parameter order
  float* gp = gpartial + (size_t)blockIdx.x * NPARAM;
  auto put = [&](float v, int dst) {
    v = warp_sum(v);
    if (lane == 0) gp[dst] = v;
  };
#pragma unroll
  for (int k = 0; k < 3; ++k) { put(gw1a[k], P_W1 + ca * 3 + k); put(gw1b[k], P_W1 + cb * 3 + k); }
  put(gb1a, P_B1 + ca); put(gb1b, P_B1 + cb);
#pragma unroll
  for (int q = 0; q < 9; ++q) { put(gw2a[q], P_W2 + ca * 9 + q); put(gw2b[q], P_W2 + cb * 9 + q); }
  put(gb2a, P_B2 + ca); put(gb2b, P_B2 + cb);
#pragma unroll
  for (int k = 0; k < 3; ++k) { put(gw3a[k], P_W3 + k * 32 + ca); put(gw3b[k], P_W3 + k * 32 + cb); }
  // db3: lanes with t == 0 hold output channels 0 / 1, lanes with t == 1 hold channel 2 (other lanes hold zeros)
  const float s0 = warp_sum(t == 0 ? gb3x : 0.f), s1 = warp_sum(t == 0 ? gb3y : 0.f), s2 = warp_sum(t == 1 ? gb3x : 0.f);
  if (lane == 0) { wred[warp * 4 + 0] = s0; wred[warp * 4 + 1] = s1; wred[warp * 4 + 2] = s2; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) s += wred[w * 4 + tid];
    gp[P_B3 + tid] = s;
  }
}

// out[c] = sum_p partial[p][c]: 32 parameters per CTA, 8 groups of partial rows combined through shared memory in a fixed order
__global__ void __launch_bounds__(256)
cnn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int n) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < n)
    for (int p = ry; p < nparts; p += 8) s += partial[(size_t)p * n + c];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < n) {
    float tsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tsum += red[k][cx];
    out[c] = tsum;
  }
}

int tile_height(int S, int cap) {
  const int ny = (S + cap - 1) / cap;
  return (S + ny - 1) / ny;
}

}  // namespace

extern "C" int32_t calm_cnn_fwd(const float* x, float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* w3, const float* b3, int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_cnn_fwd: B=%d S=%d", B, S);
  CALM_CHECK_ARG(((uintptr_t)x | (uintptr_t)y) % 16 == 0, "calm_cnn_fwd: x / y must be 16-byte aligned");
  const int th = tile_height(S, THF);
  const int tiles_x = (S + TW - 1) / TW, tiles_y = (S + th - 1) / th;
  const long long total = (long long)B * tiles_x * tiles_y;
  CALM_CHECK_ARG(total < (1LL << 31) - 4096, "calm_cnn_fwd: too many tiles");
  const long long cap = 2LL * calm_num_sms();
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(cnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) { calm_set_error("calm_cnn_fwd: smem %d: %s", (int)FWD_SMEM, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  CnnW W{w1, b1, w2, b2, w3, b3};
  cnn_fwd_kernel<<<grid, NT, FWD_SMEM, stream>>>(x, y, W, B, S, tiles_x, tiles_y, th);
  CALM_CHECK_LAUNCH("calm_cnn_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_cnn_bwd_blocks(int32_t B, int32_t S) {
  const int th = tile_height(S, THB);
  const long long total = (long long)B * ((S + TW - 1) / TW) * ((S + th - 1) / th);
  const long long cap = calm_num_sms();
  return (int32_t)(total < cap ? total : cap);
}

extern "C" int32_t calm_cnn_bwd(const float* x, const float* dy, float* dx, void* dx_bf16, const float* w1, const float* b1, const float* w2,
                                const float* b2, const float* w3, const float* b3, float* gpartial, int32_t nblocks, float* gparams,
                                int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_cnn_bwd: B=%d S=%d", B, S);
  CALM_CHECK_ARG(nblocks == calm_cnn_bwd_blocks(B, S), "calm_cnn_bwd: nblocks=%d expected %d", nblocks, calm_cnn_bwd_blocks(B, S));
  CALM_CHECK_ARG(((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dx_bf16) % 16 == 0, "calm_cnn_bwd: x / dy / dx must be 16-byte aligned");
  const int th = tile_height(S, THB);
  const int tiles_x = (S + TW - 1) / TW, tiles_y = (S + th - 1) / th;
  CALM_CHECK_ARG((long long)B * tiles_x * tiles_y < (1LL << 31) - 4096, "calm_cnn_bwd: too many tiles");
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(cnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
    if (e != cudaSuccess) { calm_set_error("calm_cnn_bwd: smem %d: %s", (int)BWD_SMEM, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  CnnW W{w1, b1, w2, b2, w3, b3};
  cnn_bwd_kernel<<<nblocks, NT, BWD_SMEM, stream>>>(x, dy, dx, reinterpret_cast<bf16*>(dx_bf16), W, gpartial, B, S, tiles_x, tiles_y, th);
  CALM_CHECK_LAUNCH("calm_cnn_bwd");
  cnn_reduce_kernel<<<(NPARAM + 31) / 32, 256, 0, stream>>>(gpartial, gparams, nblocks, NPARAM);
  CALM_CHECK_LAUNCH("calm_cnn_bwd(reduce)");
  return CALM_OK;
}
