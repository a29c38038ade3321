// Fused per-Block CNN residual on the (B, S, S, 3) token image (channels-last = the row-token layout itself):
//   y = x + conv1x1_{32->3}( gelu( dwconv3x3( gelu( conv1x1_{3->32}(x) ) ) ) )
// replaces the Sequential(sn(Conv2d 3->32,k1), GELU, sn(Conv2d 32->32,k3,groups=32,pad=1), GELU, sn(Conv2d 32->3,k1)) and the
// permutes around it (Vi_Tools_CNN_less_V2.py:378-385,400-403; CALM_ViT_V2.py:60-67,80-83). The reference materialises
// four (B,32,S,S) tensors per call; here the 32-channel intermediates only ever live in shared memory, so HBM traffic is
// the 3-channel read + 3-channel write (forward) — the backward recomputes them from x with a 2-pixel halo.
// Thread mapping: lane = hidden channel (32), warp = pixel group; 32->3 reductions use warp shuffles.
#include "common.cuh"
#include "../../include/calm_b200.h"

namespace {

constexpr int CH = 32;
constexpr int TS = 16;                 // output tile side
constexpr int CNN_THREADS = 256;
constexpr int CNN_WARPS = CNN_THREADS / 32;

struct CnnW { const float *w1, *b1, *w2, *b2, *w3, *b3; };

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(CNN_THREADS)
cnn_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, CnnW W, int B, int S, int tiles_side) {
  constexpr int HS = TS + 2;  // halo side
  __shared__ float x_s[HS * HS * 3];
  __shared__ float h1_s[HS * HS * CH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane;
  const float w10 = W.w1[c * 3], w11 = W.w1[c * 3 + 1], w12 = W.w1[c * 3 + 2], b1 = W.b1[c];
  float w2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w2[i] = W.w2[c * 9 + i];
  const float b2 = W.b2[c];
  const float w30 = W.w3[c], w31 = W.w3[CH + c], w32 = W.w3[2 * CH + c];
  const float b3 = lane < 3 ? W.b3[lane] : 0.f;
  const int tiles_per_img = tiles_side * tiles_side;
  const long long total_tiles = (long long)B * tiles_per_img;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile % tiles_per_img);
    const int ty0 = (tr / tiles_side) * TS, tx0 = (tr % tiles_side) * TS;
    const float* xb = x + (long long)b * S * S * 3;
    __syncthreads();
    for (int i = threadIdx.x; i < HS * HS * 3; i += CNN_THREADS) {
      const int p = i / 3, ch = i - p * 3;
      const int yy = ty0 + p / HS - 1, xx = tx0 + p % HS - 1;
      x_s[i] = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? xb[((long long)yy * S + xx) * 3 + ch] : 0.f;
    }
    __syncthreads();
    for (int p = warp; p < HS * HS; p += CNN_WARPS) {
      const int yy = ty0 + p / HS - 1, xx = tx0 + p % HS - 1;
      float hv = 0.f;  // zero padding applies to the hidden feature map
      if (yy >= 0 && yy < S && xx >= 0 && xx < S)
        hv = gelu_erf(w10 * x_s[p * 3] + w11 * x_s[p * 3 + 1] + w12 * x_s[p * 3 + 2] + b1);
      h1_s[p * CH + c] = hv;
    }
    __syncthreads();
    for (int p = warp; p < TS * TS; p += CNN_WARPS) {
      const int py = p / TS, px = p % TS;
      const int yy = ty0 + py, xx = tx0 + px;
      if (yy >= S || xx >= S) continue;  // warp-uniform
      float pre = b2;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) pre += w2[ky * 3 + kx] * h1_s[((py + ky) * HS + (px + kx)) * CH + c];
      const float h2 = gelu_erf(pre);
      const float o0 = warp_sum(w30 * h2), o1 = warp_sum(w31 * h2), o2 = warp_sum(w32 * h2);
      if (lane < 3) {
        const float ov = lane == 0 ? o0 : (lane == 1 ? o1 : o2);
        const int hp = (py + 1) * HS + (px + 1);
        y[((long long)b * S * S + (long long)yy * S + xx) * 3 + lane] = x_s[hp * 3 + lane] + ov + b3;
      }
    }
  }
}

// ------------------------------------------------------------------ backward
// parameter-gradient layout (CALM_CNN_NPARAM = 547): w1[96] b1[32] w2[288] b2[32] w3[96] b3[3]
__global__ void __launch_bounds__(CNN_THREADS)
cnn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, CnnW W, float* __restrict__ gpartial,
               int B, int S, int tiles_side) {
  constexpr int H2 = TS + 4;  // x / h1 halo side (2 pixels)
  constexpr int H1 = TS + 2;  // dy / dpre2 halo side (1 pixel)
  extern __shared__ float sm[];
  float* x_s = sm;                         // H2*H2*3
  float* dy_s = x_s + H2 * H2 * 3;         // H1*H1*3
  float* h1_s = dy_s + H1 * H1 * 3;        // H2*H2*CH
  float* dp2_s = h1_s + H2 * H2 * CH;      // H1*H1*CH
  float* red_s = dp2_s + H1 * H1 * CH;     // CNN_WARPS * 17 * CH  (+ 3 for b3)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane;
  const float w10 = W.w1[c * 3], w11 = W.w1[c * 3 + 1], w12 = W.w1[c * 3 + 2], b1 = W.b1[c];
  float w2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w2[i] = W.w2[c * 9 + i];
  const float b2 = W.b2[c];
  const float w30 = W.w3[c], w31 = W.w3[CH + c], w32 = W.w3[2 * CH + c];
  const int tiles_per_img = tiles_side * tiles_side;
  const long long total_tiles = (long long)B * tiles_per_img;

  float g_w1[3] = {0.f, 0.f, 0.f}, g_b1 = 0.f, g_w2[9], g_b2 = 0.f, g_w3[3] = {0.f, 0.f, 0.f}, g_b3 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) g_w2[i] = 0.f;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile % tiles_per_img);
    const int ty0 = (tr / tiles_side) * TS, tx0 = (tr % tiles_side) * TS;
    const float* xb = x + (long long)b * S * S * 3;
    const float* dyb = dy + (long long)b * S * S * 3;
    __syncthreads();
    for (int i = threadIdx.x; i < H2 * H2 * 3; i += CNN_THREADS) {
      const int p = i / 3, ch = i - p * 3;
      const int yy = ty0 + p / H2 - 2, xx = tx0 + p % H2 - 2;
      x_s[i] = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? xb[((long long)yy * S + xx) * 3 + ch] : 0.f;
    }
    for (int i = threadIdx.x; i < H1 * H1 * 3; i += CNN_THREADS) {
      const int p = i / 3, ch = i - p * 3;
      const int yy = ty0 + p / H1 - 1, xx = tx0 + p % H1 - 1;
      dy_s[i] = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? dyb[((long long)yy * S + xx) * 3 + ch] : 0.f;
    }
    __syncthreads();
    // hidden map h1 on the 2-pixel halo (0 outside the image)
    for (int p = warp; p < H2 * H2; p += CNN_WARPS) {
      const int yy = ty0 + p / H2 - 2, xx = tx0 + p % H2 - 2;
      float hv = 0.f;
      if (yy >= 0 && yy < S && xx >= 0 && xx < S)
        hv = gelu_erf(w10 * x_s[p * 3] + w11 * x_s[p * 3 + 1] + w12 * x_s[p * 3 + 2] + b1);
      h1_s[p * CH + c] = hv;
    }
    __syncthreads();
    // d pre2 on the 1-pixel halo; weight gradients only from the pixels this tile owns
    for (int p = warp; p < H1 * H1; p += CNN_WARPS) {
      const int py = p / H1, px = p % H1;
      const int yy = ty0 + py - 1, xx = tx0 + px - 1;
      float dp2 = 0.f;
      if (yy >= 0 && yy < S && xx >= 0 && xx < S) {
        float pre = b2;
        float nb[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            nb[ky * 3 + kx] = h1_s[((py + ky) * H2 + (px + kx)) * CH + c];
            pre += w2[ky * 3 + kx] * nb[ky * 3 + kx];
          }
        const float d0 = dy_s[p * 3], d1 = dy_s[p * 3 + 1], d2 = dy_s[p * 3 + 2];
        dp2 = (w30 * d0 + w31 * d1 + w32 * d2) * dgelu_erf(pre);
        const bool owned = py >= 1 && py <= TS && px >= 1 && px <= TS;
        if (owned) {
          const float h2 = gelu_erf(pre);
          g_w3[0] += d0 * h2; g_w3[1] += d1 * h2; g_w3[2] += d2 * h2;
          g_b2 += dp2;
#pragma unroll
          for (int i = 0; i < 9; ++i) g_w2[i] += dp2 * nb[i];
          if (lane < 3) g_b3 += (lane == 0 ? d0 : (lane == 1 ? d1 : d2));
        }
      }
      dp2_s[p * CH + c] = dp2;
    }
    __syncthreads();
    // d h1 -> d pre1 -> dx on the owned pixels
    for (int p = warp; p < TS * TS; p += CNN_WARPS) {
      const int py = p / TS, px = p % TS;
      const int yy = ty0 + py, xx = tx0 + px;
      if (yy >= S || xx >= S) continue;
      float dh1 = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)  // out(q) used h1(q + off) with off = (ky-1,kx-1)  =>  h1(p) fed out(p - off)
          dh1 += w2[ky * 3 + kx] * dp2_s[((py + 1 - (ky - 1)) * H1 + (px + 1 - (kx - 1))) * CH + c];
      const int xp = (py + 2) * H2 + (px + 2);
      const float x0 = x_s[xp * 3], x1 = x_s[xp * 3 + 1], x2 = x_s[xp * 3 + 2];
      const float dp1 = dh1 * dgelu_erf(w10 * x0 + w11 * x1 + w12 * x2 + b1);
      g_w1[0] += dp1 * x0; g_w1[1] += dp1 * x1; g_w1[2] += dp1 * x2; g_b1 += dp1;
      const float r0 = warp_sum(w10 * dp1), r1 = warp_sum(w11 * dp1), r2 = warp_sum(w12 * dp1);
      if (lane < 3) {
        const float rv = lane == 0 ? r0 : (lane == 1 ? r1 : r2);
        const int dp = (py + 1) * H1 + (px + 1);
        dx[((long long)b * S * S + (long long)yy * S + xx) * 3 + lane] = rv + dy_s[dp * 3 + lane];
      }
    }
  }
  // cross-warp reduction of the per-channel accumulators, then one partial row per CTA
  __syncthreads();
  float* mine = red_s + (size_t)warp * 17 * CH;
  mine[0 * CH + c] = g_w1[0]; mine[1 * CH + c] = g_w1[1]; mine[2 * CH + c] = g_w1[2];
  mine[3 * CH + c] = g_b1;
#pragma unroll
  for (int i = 0; i < 9; ++i) mine[(4 + i) * CH + c] = g_w2[i];
  mine[13 * CH + c] = g_b2;
  mine[14 * CH + c] = g_w3[0]; mine[15 * CH + c] = g_w3[1]; mine[16 * CH + c] = g_w3[2];
  float* b3_s = red_s + (size_t)CNN_WARPS * 17 * CH;
  if (lane < 3) b3_s[warp * 3 + lane] = g_b3;
  __syncthreads();
  float* gp = gpartial + (size_t)blockIdx.x * CALM_CNN_NPARAM;
  for (int i = threadIdx.x; i < 17 * CH; i += CNN_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < CNN_WARPS; ++w) s += red_s[(size_t)w * 17 * CH + i];
    const int k = i / CH, cc = i % CH;
    int dst;
    if (k < 3) dst = cc * 3 + k;                       // w1[c][k]
    else if (k == 3) dst = 96 + cc;                    // b1[c]
    else if (k < 13) dst = 128 + cc * 9 + (k - 4);     // w2[c][tap]
    else if (k == 13) dst = 416 + cc;                  // b2[c]
    else dst = 448 + (k - 14) * CH + cc;               // w3[o][c]
    gp[dst] = s;
  }
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < CNN_WARPS; ++w) s += b3_s[w * 3 + threadIdx.x];
    gp[544 + threadIdx.x] = s;
  }
}

__global__ void cnn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n + c];
  out[c] = s;
}

size_t cnn_bwd_smem() {
  return ((size_t)(TS + 4) * (TS + 4) * 3 + (size_t)(TS + 2) * (TS + 2) * 3 + (size_t)(TS + 4) * (TS + 4) * CH +
          (size_t)(TS + 2) * (TS + 2) * CH + (size_t)CNN_WARPS * 17 * CH + 32) * sizeof(float);
}

}  // namespace

extern "C" int32_t calm_cnn_fwd(const float* x, float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* w3, const float* b3, int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_cnn_fwd: B=%d S=%d", B, S);
  const int tiles_side = (S + TS - 1) / TS;
  const long long total = (long long)B * tiles_side * tiles_side;
  const long long cap = 8LL * calm_num_sms();
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  CnnW W{w1, b1, w2, b2, w3, b3};
  cnn_fwd_kernel<<<grid, CNN_THREADS, 0, stream>>>(x, y, W, B, S, tiles_side);
  CALM_CHECK_LAUNCH("calm_cnn_fwd");
  return CALM_OK;
}

extern "C" int32_t calm_cnn_bwd_blocks(int32_t B, int32_t S) {
  const int tiles_side = (S + TS - 1) / TS;
  const long long total = (long long)B * tiles_side * tiles_side;
  const long long cap = 2LL * calm_num_sms();
  return (int32_t)(total < cap ? total : cap);
}

extern "C" int32_t calm_cnn_bwd(const float* x, const float* dy, float* dx, const float* w1, const float* b1, const float* w2,
                                const float* b2, const float* w3, const float* b3, float* gpartial, int32_t nblocks, float* gparams,
                                int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(B > 0 && S > 0, "calm_cnn_bwd: B=%d S=%d", B, S);
  CALM_CHECK_ARG(nblocks == calm_cnn_bwd_blocks(B, S), "calm_cnn_bwd: nblocks=%d expected %d", nblocks, calm_cnn_bwd_blocks(B, S));
  const int tiles_side = (S + TS - 1) / TS;
  const size_t smem = cnn_bwd_smem();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(cnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { calm_set_error("calm_cnn_bwd: smem %zu: %s", smem, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured = true;
  }
  CnnW W{w1, b1, w2, b2, w3, b3};
  cnn_bwd_kernel<<<nblocks, CNN_THREADS, smem, stream>>>(x, dy, dx, W, gpartial, B, S, tiles_side);
  CALM_CHECK_LAUNCH("calm_cnn_bwd");
  cnn_reduce_kernel<<<(CALM_CNN_NPARAM + 127) / 128, 128, 0, stream>>>(gpartial, gparams, nblocks, CALM_CNN_NPARAM);
  CALM_CHECK_LAUNCH("calm_cnn_bwd(reduce)");
  return CALM_OK;
}
