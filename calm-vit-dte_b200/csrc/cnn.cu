// Fused per-Block CNN residual on the (B, S, S, 3) token image (channels-last = the row-token layout itself):
//   y = x + conv1x1_{32->3}( gelu( dwconv3x3( gelu( conv1x1_{3->32}(x) ) ) ) )
// replaces the Sequential(sn(Conv2d 3->32,k1), GELU, sn(Conv2d 32->32,k3,groups=32,pad=1), GELU, sn(Conv2d 32->3,k1)) and the
// permutes around it (Vi_Tools_CNN_less_V2.py:378-385,400-403; CALM_ViT_V2.py:60-67,80-83). The reference materialises
// four (B,32,S,S) tensors per call; here the 32-channel intermediates only ever live in shared memory, so HBM traffic is
// the 3-channel read + 3-channel write (forward) — the backward recomputes them from x with a 2-pixel halo.
//
// Mapping (v2): the depthwise structure makes channels independent, so the CHANNEL loop is the outer loop and a thread owns
// PIXELS: per chunk of channels one shared-memory plane per channel holds the hidden map of the tile (+halo); a thread
// then produces 4 horizontally adjacent pixels from 3 x (LDS.128 + LDS.64) per channel and accumulates the 32->3
// projection in registers — no cross-lane reductions in the data path, weights are warp-uniform shared-memory loads.
// The kernel is bound by the 2 GELU evaluations per pixel-channel (FP32 + MUFU pipes), not by HBM: GELU(erf) uses the
// Abramowitz-Stegun 7.1.26 erfc form (|abs err| < 1.5e-7; one MUFU.RCP + one MUFU.EX2), and its derivative reuses the
// exponential. Parameter gradients: per-warp shuffle reductions per channel into per-warp shared accumulators (no atomics,
// deterministic), one partial row per CTA, then a second-stage reduce.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/calm_b200.h"

namespace {

constexpr int CH = 32;
constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int WSM = 644;  // w1b[32][4] | w2p[32][12] | w3t[32][4] | b3[4]

struct CnnW { const float *w1, *b1, *w2, *b2, *w3, *b3; };

__device__ __forceinline__ float gelu_fast(float x) { return gelu_erf(x); }


// Sum N register values over the 32 lanes with N - 1 + log2(32/N)... shuffles instead of 5 N: at every butterfly step a lane
// keeps one half of its values and ships the other half to its partner. On return v[0] of lane L holds the all-lane sum of
// value (L >> (5 - log2 N)): the lanes of a group all hold the same total, in the same (fixed) summation order.
__device__ __forceinline__ void warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step, half = 8 >> step;
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = hi ? v[i] : v[i + half];
      const float keep = hi ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ void warp_sum8(float (&v)[8], int lane) {
#pragma unroll
  for (int step = 0; step < 3; ++step) {
    const int off = 16 >> step, half = 4 >> step;
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = hi ? v[i] : v[i + half];
      const float keep = hi ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ void load_weights(float* wsm, const CnnW& W) {
  for (int i = threadIdx.x; i < CH; i += NT) {
    wsm[i * 4 + 0] = W.w1[i * 3 + 0]; wsm[i * 4 + 1] = W.w1[i * 3 + 1]; wsm[i * 4 + 2] = W.w1[i * 3 + 2]; wsm[i * 4 + 3] = W.b1[i];
    float* w2p = wsm + 128 + i * 12;
#pragma unroll
    for (int k = 0; k < 9; ++k) w2p[k] = W.w2[i * 9 + k];
    w2p[9] = W.b2[i]; w2p[10] = 0.f; w2p[11] = 0.f;
    float* w3t = wsm + 512 + i * 4;
    w3t[0] = W.w3[i]; w3t[1] = W.w3[CH + i]; w3t[2] = W.w3[2 * CH + i]; w3t[3] = 0.f;
  }
  if (threadIdx.x < 3) wsm[640 + threadIdx.x] = W.b3[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------------------------
// forward: 32 x 32 pixel tiles, 4 channels per chunk, double-buffered planes (one __syncthreads per chunk)
// ------------------------------------------------------------------------------------------------------------------
constexpr int FT = 32;                      // tile side
constexpr int FCC = 4;                      // channels per chunk
constexpr int FPW = 36;                     // plane pitch (34 columns used)
constexpr int FPLANE = (FT + 2) * FPW;      // 1224
constexpr int FXS = (FT + 2) * (FT + 2) * 3;  // 3468
constexpr int FWD_SMEM_FLOATS = FXS + WSM + 2 * FCC * FPLANE;

__global__ void __launch_bounds__(NT, 3)
cnn_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, CnnW W, int B, int S, int tiles_side) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* wsm = xs + FXS;
  float* planes = wsm + WSM;
  load_weights(wsm, W);
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const int tiles_per_img = tiles_side * tiles_side;
  const long long total_tiles = (long long)B * tiles_per_img;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile - (long long)b * tiles_per_img);
    const int ty0 = (tr / tiles_side) * FT, tx0 = (tr % tiles_side) * FT;
    const float* xb = x + (long long)b * S * S * 3;
    __syncthreads();
    for (int i = tid; i < FXS; i += NT) {
      const int row = i / ((FT + 2) * 3), rem = i - row * ((FT + 2) * 3);
      const int gy = ty0 + row - 1, gx3 = (tx0 - 1) * 3 + rem;
      xs[i] = (gy >= 0 && gy < S && gx3 >= 0 && gx3 < S * 3) ? xb[(long long)gy * S * 3 + gx3] : 0.f;
    }
    __syncthreads();
    float out[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j][0] = out[j][1] = out[j][2] = 0.f;

    for (int chunk = 0; chunk < CH / FCC; ++chunk) {
      float* buf = planes + (chunk & 1) * FCC * FPLANE;
      // phase B: hidden map h1 = gelu(conv1x1(x)) on the 34 x 34 halo region (zero outside the image: conv zero padding)
      for (int idx = tid; idx < (FT + 2) * (FT + 2); idx += NT) {
        const int py = idx / (FT + 2), px = idx - py * (FT + 2);
        const int gy = ty0 + py - 1, gx = tx0 + px - 1;
        const bool inside = gy >= 0 && gy < S && gx >= 0 && gx < S;
        const float x0 = xs[idx * 3], x1 = xs[idx * 3 + 1], x2 = xs[idx * 3 + 2];
#pragma unroll
        for (int k = 0; k < FCC; ++k) {
          const float4 w = *reinterpret_cast<const float4*>(wsm + (chunk * FCC + k) * 4);
          const float pre = fmaf(w.x, x0, fmaf(w.y, x1, fmaf(w.z, x2, w.w)));
          buf[k * FPLANE + py * FPW + px] = inside ? gelu_fast(pre) : 0.f;
        }
      }
      __syncthreads();
      // phase C: depthwise 3x3 + GELU + 32->3 projection for this thread's 4 pixels
#pragma unroll
      for (int k = 0; k < FCC; ++k) {
        const int c = chunk * FCC + k;
        const float* pl = buf + k * FPLANE + ty * FPW + 4 * tx;
        float r[3][6];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(pl + q * FPW);
          const float2 bb = *reinterpret_cast<const float2*>(pl + q * FPW + 4);
          r[q][0] = a.x; r[q][1] = a.y; r[q][2] = a.z; r[q][3] = a.w; r[q][4] = bb.x; r[q][5] = bb.y;
        }
        const float4 wa = *reinterpret_cast<const float4*>(wsm + 128 + c * 12);
        const float4 wb = *reinterpret_cast<const float4*>(wsm + 128 + c * 12 + 4);
        const float4 wc = *reinterpret_cast<const float4*>(wsm + 128 + c * 12 + 8);
        const float4 w3 = *reinterpret_cast<const float4*>(wsm + 512 + c * 4);
        const float w2[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float pre = wc.y;  // b2
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) pre = fmaf(w2[ky * 3 + kx], r[ky][j + kx], pre);
          const float h2 = gelu_fast(pre);
          out[j][0] = fmaf(w3.x, h2, out[j][0]);
          out[j][1] = fmaf(w3.y, h2, out[j][1]);
          out[j][2] = fmaf(w3.z, h2, out[j][2]);
        }
      }
    }
    // y = x + cnn(x) + b3
    const int gy = ty0 + ty;
    if (gy < S) {
      const float b30 = wsm[640], b31 = wsm[641], b32 = wsm[642];
      float* yrow = y + ((long long)b * S * S + (long long)gy * S) * 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gx = tx0 + 4 * tx + j;
        if (gx < S) {
          const float* xc = xs + ((ty + 1) * (FT + 2) + (4 * tx + j + 1)) * 3;
          yrow[gx * 3 + 0] = xc[0] + out[j][0] + b30;
          yrow[gx * 3 + 1] = xc[1] + out[j][1] + b31;
          yrow[gx * 3 + 2] = xc[2] + out[j][2] + b32;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: 32 x th pixel tiles (th <= 24), 2 channels per chunk.
//   phase B: h1 = gelu(pre1), g1 = gelu'(pre1) on the 2-pixel halo        (recomputed from x)
//   phase C: pre2 -> h2, dp2 = (W3^T dy) * gelu'(pre2) on the 1-pixel halo; dW3, db2, dW2 from the owned pixels
//   phase D: dh1 = dwconv^T(dp2), dp1 = dh1 * g1 ; dx += W1^T dp1 ; dW1, db1
// parameter-gradient layout (CALM_CNN_NPARAM = 547): w1[96] b1[32] w2[288] b2[32] w3[96] b3[3]
// ------------------------------------------------------------------------------------------------------------------
constexpr int BT = 32;          // tile width
constexpr int BTH = 24;         // max tile height
constexpr int BCC = 2;
constexpr int H1P = 40, DPP = 36, G1P = 32;
constexpr int B_XS = (BTH + 4) * (BT + 4) * 3;     // 3024
constexpr int B_DYS = (BTH + 2) * (BT + 2) * 3;    // 2652
constexpr int B_H1 = BCC * (BTH + 4) * H1P;        // 2240
constexpr int B_G1 = 2 * BCC * BTH * G1P;          // 3072 (double-buffered)
constexpr int B_DP2 = BCC * (BTH + 2) * DPP;       // 1872
constexpr int WACC = 548;                          // per-warp accumulators: 32 x 17 + 3 (+1 pad)
constexpr int BWD_SMEM_FLOATS = B_XS + B_DYS + B_H1 + B_G1 + B_DP2 + WSM + NW * WACC;

__device__ __forceinline__ void
cnn_bwd_body(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, bf16* __restrict__ dx16, const CnnW& W,
             float* __restrict__ gpartial, int B, int S, int tiles_x, int tiles_y, int th) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* dys = xs + B_XS;
  float* h1s = dys + B_DYS;
  float* g1s = h1s + B_H1;
  float* dp2s = g1s + B_G1;
  float* wsm = dp2s + B_DP2;
  float* wacc = wsm + WSM;
  load_weights(wsm, W);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NW * WACC; i += NT) wacc[i] = 0.f;
  float* myacc = wacc + warp * WACC;
  const int tiles_per_img = tiles_x * tiles_y;
  const long long total_tiles = (long long)B * tiles_per_img;
  const int h1_plane = (BTH + 4) * H1P, g1_plane = BTH * G1P, dp_plane = (BTH + 2) * DPP;
  // phase C task: row rr (0..th+1) of the 1-halo region, 4-pixel group gc (0..8)
  const bool c_active = tid < 9 * (th + 2);
  // row and 4-pixel group packed into one opaque register: ptxas otherwise re-derives tid / 9 inside the channel loop
  unsigned c_pack = (unsigned)(tid / 9) | (unsigned)(tid % 9) << 8;
  asm volatile("" : "+r"(c_pack));
#define c_rr ((int)(c_pack & 255u))
#define c_g ((int)(c_pack >> 8))
  // phase D task: row dr (0..th-1), group dg (0..7)
  const bool d_active = tid < 8 * th;
  const int d_r = tid >> 3, d_g = tid & 7;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile - (long long)b * tiles_per_img);
    const int ty0 = (tr / tiles_x) * th, tx0 = (tr % tiles_x) * BT;
    const float* xb = x + (long long)b * S * S * 3;
    const float* dyb = dy + (long long)b * S * S * 3;
    __syncthreads();
    for (int i = tid; i < (th + 4) * (BT + 4) * 3; i += NT) {
      const int row = i / ((BT + 4) * 3), rem = i - row * ((BT + 4) * 3);
      const int gy = ty0 + row - 2, gx3 = (tx0 - 2) * 3 + rem;
      xs[i] = (gy >= 0 && gy < S && gx3 >= 0 && gx3 < S * 3) ? xb[(long long)gy * S * 3 + gx3] : 0.f;
    }
    for (int i = tid; i < (th + 2) * (BT + 2) * 3; i += NT) {
      const int row = i / ((BT + 2) * 3), rem = i - row * ((BT + 2) * 3);
      const int gy = ty0 + row - 1, gx3 = (tx0 - 1) * 3 + rem;
      dys[i] = (gy >= 0 && gy < S && gx3 >= 0 && gx3 < S * 3) ? dyb[(long long)gy * S * 3 + gx3] : 0.f;
    }
    __syncthreads();
    float dxa[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) dxa[j][0] = dxa[j][1] = dxa[j][2] = 0.f;
    // phase C pixel classes of this thread are the same for all 32 channels: bit j = inside the image (and the 1-pixel halo
    // columns of the tile), bit 4 + j = owned by this tile
    unsigned cmask = 0;
    {
      const int rloc = c_rr - 1, gy = ty0 + rloc;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cloc = 4 * c_g + j - 1, gx = tx0 + cloc;
        const bool valid = cloc <= BT && gy >= 0 && gy < S && gx >= 0 && gx < S;
        const bool owned = valid && rloc >= 0 && rloc < th && cloc >= 0 && cloc < BT;
        cmask |= (valid ? 1u : 0u) << j | (owned ? 16u : 0u) << j;
      }
    }
    asm volatile("" : "+r"(cmask));   // opaque from here on: ptxas otherwise re-derives the eight comparisons inside the channel loop

    for (int chunk = 0; chunk < CH / BCC; ++chunk) {
      float* g1b = g1s + (chunk & 1) * BCC * g1_plane;
      // ---------------- phase B
      for (int idx = tid; idx < (th + 4) * (BT + 4); idx += NT) {
        const int rr = idx / (BT + 4), cc = idx - rr * (BT + 4);
        const int r = rr - 2, c = cc - 2;
        const int gy = ty0 + r, gx = tx0 + c;
        const bool inside = gy >= 0 && gy < S && gx >= 0 && gx < S;
        const bool interior = r >= 0 && r < th && c >= 0 && c < BT;
        const float x0 = xs[idx * 3], x1 = xs[idx * 3 + 1], x2 = xs[idx * 3 + 2];
#pragma unroll
        for (int k = 0; k < BCC; ++k) {
          const float4 w = *reinterpret_cast<const float4*>(wsm + (chunk * BCC + k) * 4);
          const float pre = fmaf(w.x, x0, fmaf(w.y, x1, fmaf(w.z, x2, w.w)));
          float h, g;
          gelu_pair(pre, h, g);
          h1s[k * h1_plane + rr * H1P + cc] = inside ? h : 0.f;
          if (interior) g1b[k * g1_plane + r * G1P + c] = inside ? g : 0.f;
        }
      }
      __syncthreads();
      // ---------------- phase C
#pragma unroll
      for (int k = 0; k < BCC; ++k) {
        const int ch = chunk * BCC + k;
        float gw3[3] = {0.f, 0.f, 0.f}, gb2 = 0.f, gw2[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) gw2[q] = 0.f;
        if (c_active) {
          const float* pl = h1s + k * h1_plane + c_rr * H1P + 4 * c_g;
          float r[3][6];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(pl + q * H1P);
            const float2 bb = *reinterpret_cast<const float2*>(pl + q * H1P + 4);
            r[q][0] = a.x; r[q][1] = a.y; r[q][2] = a.z; r[q][3] = a.w; r[q][4] = bb.x; r[q][5] = bb.y;
          }
          const float4 wa = *reinterpret_cast<const float4*>(wsm + 128 + ch * 12);
          const float4 wb = *reinterpret_cast<const float4*>(wsm + 128 + ch * 12 + 4);
          const float4 wc = *reinterpret_cast<const float4*>(wsm + 128 + ch * 12 + 8);
          const float4 w3 = *reinterpret_cast<const float4*>(wsm + 512 + ch * 4);
          const float w2[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
          float dpv[4];
          const float* dyp = dys + (c_rr * (BT + 2) + 4 * c_g) * 3;   // pixel cloc = 4 c_g + j - 1 sits at column cloc + 1
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool valid = (cmask >> j) & 1u;
            float pre = wc.y;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) pre = fmaf(w2[ky * 3 + kx], r[ky][j + kx], pre);
            float h2, d2;
            gelu_pair(pre, h2, d2);
            // dy is zero outside the image; the two right-most pixels of the last group lie past the halo and read whatever
            // follows in shared memory (planes, other rows): their product is discarded, never a NaN in dp2
            const float d0 = dyp[3 * j], d1 = dyp[3 * j + 1], d2y = dyp[3 * j + 2];
            const float dpre = valid ? (w3.x * d0 + w3.y * d1 + w3.z * d2y) * d2 : 0.f;
            dpv[j] = dpre;
            if ((cmask >> (4 + j)) & 1u) {
              gw3[0] = fmaf(d0, h2, gw3[0]); gw3[1] = fmaf(d1, h2, gw3[1]); gw3[2] = fmaf(d2y, h2, gw3[2]);
              gb2 += dpre;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) gw2[ky * 3 + kx] = fmaf(dpre, r[ky][j + kx], gw2[ky * 3 + kx]);
            }
          }
          *reinterpret_cast<float4*>(dp2s + k * dp_plane + c_rr * DPP + 4 * c_g) = make_float4(dpv[0], dpv[1], dpv[2], dpv[3]);
        }
        // per-warp reduction of the 13 per-channel sums (all lanes take part; inactive lanes hold zeros): lane L ends up
        // with the total of value L >> 1, the even lanes of the first 13 pairs add it to the warp's accumulator row
        float v[16] = {gw2[0], gw2[1], gw2[2], gw2[3], gw2[4], gw2[5], gw2[6], gw2[7], gw2[8], gb2, gw3[0], gw3[1], gw3[2], 0.f, 0.f, 0.f};
        warp_sum16(v, lane);
        if (!(lane & 1) && lane < 26) myacc[ch * 17 + 4 + (lane >> 1)] += v[0];
      }
      __syncthreads();
      // ---------------- phase D
      float dsum[4 * BCC];
#pragma unroll
      for (int k = 0; k < BCC; ++k) {
        const int ch = chunk * BCC + k;
        float gw1[3] = {0.f, 0.f, 0.f}, gb1 = 0.f;
        if (d_active) {
          const float* pl = dp2s + k * dp_plane + d_r * DPP + 4 * d_g;
          float r[3][6];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(pl + q * DPP);
            const float2 bb = *reinterpret_cast<const float2*>(pl + q * DPP + 4);
            r[q][0] = a.x; r[q][1] = a.y; r[q][2] = a.z; r[q][3] = a.w; r[q][4] = bb.x; r[q][5] = bb.y;
          }
          const float4 wa = *reinterpret_cast<const float4*>(wsm + 128 + ch * 12);
          const float4 wb = *reinterpret_cast<const float4*>(wsm + 128 + ch * 12 + 4);
          const float4 wc = *reinterpret_cast<const float4*>(wsm + 128 + ch * 12 + 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wsm + ch * 4);
          const float w2[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
          const float4 gv = *reinterpret_cast<const float4*>(g1b + k * g1_plane + d_r * G1P + 4 * d_g);
          const float g1v[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float dh = 0.f;  // out(q) used h1(q + (ky-1,kx-1))  =>  h1(p) fed out(p - (ky-1,kx-1))
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) dh = fmaf(w2[ky * 3 + kx], r[2 - ky][j + 2 - kx], dh);
            const float dp1 = dh * g1v[j];  // g1 = 0 outside the image
            const float* xc = xs + ((d_r + 2) * (BT + 4) + (4 * d_g + j + 2)) * 3;
            dxa[j][0] = fmaf(w1.x, dp1, dxa[j][0]);
            dxa[j][1] = fmaf(w1.y, dp1, dxa[j][1]);
            dxa[j][2] = fmaf(w1.z, dp1, dxa[j][2]);
            gw1[0] = fmaf(dp1, xc[0], gw1[0]); gw1[1] = fmaf(dp1, xc[1], gw1[1]); gw1[2] = fmaf(dp1, xc[2], gw1[2]);
            gb1 += dp1;
          }
        }
        dsum[4 * k] = gw1[0]; dsum[4 * k + 1] = gw1[1]; dsum[4 * k + 2] = gw1[2]; dsum[4 * k + 3] = gb1;
      }
      static_assert(BCC == 2, "phase D reduces the 2 x 4 sums of a chunk in one 8-value butterfly");
      warp_sum8(dsum, lane);   // lane L: total of value L >> 2 = (channel (L >> 4), slot (L >> 2) & 3)
      if (!(lane & 3)) myacc[(chunk * BCC + (lane >> 4)) * 17 + ((lane >> 2) & 3)] += dsum[0];
      // no barrier here: the next phase B writes h1 (last read before the barrier above) and the OTHER g1 buffer;
      // dp2 is rewritten only after the next barrier
    }
    // dx = dy + W1^T dp1 ; db3 += dy over the owned pixels
    float gb3[3] = {0.f, 0.f, 0.f};
    if (d_active) {
      const int gy = ty0 + d_r;
      if (gy < S) {
        float* dxrow = dx + ((long long)b * S * S + (long long)gy * S) * 3;
        bf16* dxrow16 = dx16 ? dx16 + ((long long)b * S * S + (long long)gy * S) * 3 : nullptr;   // optional bf16 copy of dx
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int gx = tx0 + 4 * d_g + j;
          if (gx < S) {
            const float* dc = dys + ((d_r + 1) * (BT + 2) + (4 * d_g + j + 1)) * 3;
            dxrow[gx * 3 + 0] = dxa[j][0] + dc[0];
            dxrow[gx * 3 + 1] = dxa[j][1] + dc[1];
            dxrow[gx * 3 + 2] = dxa[j][2] + dc[2];
            if (dxrow16) {
              dxrow16[gx * 3 + 0] = __float2bfloat16(dxa[j][0] + dc[0]);
              dxrow16[gx * 3 + 1] = __float2bfloat16(dxa[j][1] + dc[1]);
              dxrow16[gx * 3 + 2] = __float2bfloat16(dxa[j][2] + dc[2]);
            }
            gb3[0] += dc[0]; gb3[1] += dc[1]; gb3[2] += dc[2];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) gb3[q] = warp_sum(gb3[q]);
    if (lane == 0) { myacc[544] += gb3[0]; myacc[545] += gb3[1]; myacc[546] += gb3[2]; }
  }
  // one partial row per CTA, in the public parameter order
  __syncthreads();
  float* gp = gpartial + (size_t)blockIdx.x * CALM_CNN_NPARAM;
  for (int i = tid; i < 547; i += NT) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += wacc[w * WACC + i];
    int dst;
    if (i >= 544) dst = i;
    else {
      const int ch = i / 17, q = i - ch * 17;
      if (q < 3) dst = ch * 3 + q;                    // w1[c][k]
      else if (q == 3) dst = 96 + ch;                 // b1[c]
      else if (q < 13) dst = 128 + ch * 9 + (q - 4);  // w2[c][tap]
      else if (q == 13) dst = 416 + ch;               // b2[c]
      else dst = 448 + (q - 14) * CH + ch;            // w3[o][c]
    }
    gp[dst] = s;
  }
}

#undef c_rr
#undef c_g

// two register budgets of the same body: 3 CTAs/SM (80 registers, a few spills) or 2 CTAs/SM (no spills)
__global__ void __launch_bounds__(NT, 3)
cnn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, bf16* __restrict__ dx16, CnnW W,
               float* __restrict__ gpartial, int B, int S, int tiles_x, int tiles_y, int th) {
  cnn_bwd_body(x, dy, dx, dx16, W, gpartial, B, S, tiles_x, tiles_y, th);
}
__global__ void __launch_bounds__(NT, 2)
cnn_bwd_kernel_occ2(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, bf16* __restrict__ dx16, CnnW W,
                    float* __restrict__ gpartial, int B, int S, int tiles_x, int tiles_y, int th) {
  cnn_bwd_body(x, dy, dx, dx16, W, gpartial, B, S, tiles_x, tiles_y, th);
}

// out[c] = sum_p partial[p][c]: 32 parameters per CTA, 8 groups of partial rows combined through shared memory in a fixed
// order (one thread walking all ~444 rows of a column took 30 us of pure latency per call)
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
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cx];
    out[c] = t;
  }
}

int bwd_tile_height(int S) {
  const int ny = (S + BTH - 1) / BTH;
  return (S + ny - 1) / ny;
}

}  // namespace

extern "C" int32_t calm_cnn_fwd(const float* x, float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* w3, const float* b3, int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_cnn_fwd: B=%d S=%d", B, S);
  const int tiles_side = (S + FT - 1) / FT;
  const long long total = (long long)B * tiles_side * tiles_side;
  const long long cap = 3LL * calm_num_sms();
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  const size_t smem = (size_t)FWD_SMEM_FLOATS * sizeof(float);
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(cnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { calm_set_error("calm_cnn_fwd: smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  CnnW W{w1, b1, w2, b2, w3, b3};
  cnn_fwd_kernel<<<grid, NT, smem, stream>>>(x, y, W, B, S, tiles_side);
  CALM_CHECK_LAUNCH("calm_cnn_fwd");
  return CALM_OK;
}

namespace {
int bwd_ctas_per_sm() {
  static int v = 0;
  if (!v) { const char* e = getenv("CALM_CNN_BWD_OCC"); v = (e && e[0] == '2') ? 2 : 3; }
  return v;
}
}  // namespace

extern "C" int32_t calm_cnn_bwd_blocks(int32_t B, int32_t S) {
  const int th = bwd_tile_height(S);
  const long long total = (long long)B * ((S + BT - 1) / BT) * ((S + th - 1) / th);
  const long long cap = (long long)bwd_ctas_per_sm() * calm_num_sms();
  return (int32_t)(total < cap ? total : cap);
}

extern "C" int32_t calm_cnn_bwd(const float* x, const float* dy, float* dx, void* dx_bf16, const float* w1, const float* b1, const float* w2,
                                const float* b2, const float* w3, const float* b3, float* gpartial, int32_t nblocks, float* gparams,
                                int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_cnn_bwd: B=%d S=%d", B, S);
  CALM_CHECK_ARG(nblocks == calm_cnn_bwd_blocks(B, S), "calm_cnn_bwd: nblocks=%d expected %d", nblocks, calm_cnn_bwd_blocks(B, S));
  const int th = bwd_tile_height(S);
  const int tiles_x = (S + BT - 1) / BT, tiles_y = (S + th - 1) / th;
  const size_t smem = (size_t)BWD_SMEM_FLOATS * sizeof(float);
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(cnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cnn_bwd_kernel_occ2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { calm_set_error("calm_cnn_bwd: smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  CnnW W{w1, b1, w2, b2, w3, b3};
  bf16* dx16 = reinterpret_cast<bf16*>(dx_bf16);
  if (bwd_ctas_per_sm() == 2) cnn_bwd_kernel_occ2<<<nblocks, NT, smem, stream>>>(x, dy, dx, dx16, W, gpartial, B, S, tiles_x, tiles_y, th);
  else cnn_bwd_kernel<<<nblocks, NT, smem, stream>>>(x, dy, dx, dx16, W, gpartial, B, S, tiles_x, tiles_y, th);
  CALM_CHECK_LAUNCH("calm_cnn_bwd");
  cnn_reduce_kernel<<<(CALM_CNN_NPARAM + 31) / 32, 256, 0, stream>>>(gpartial, gparams, nblocks, CALM_CNN_NPARAM);
  CALM_CHECK_LAUNCH("calm_cnn_bwd(reduce)");
  return CALM_OK;
}
