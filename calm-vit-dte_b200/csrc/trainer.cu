// Training-step glue of the reference's per-rank loop, on device and without host round trips:
//   * loss heads: soft-target cross entropy (+ the dominant-class accuracy the loop prints) and Huber + 0.1 kl
//     (distributed_trainer_cls.py:63,86,97-100; distributed_trainer_reg.py:76-88);
//   * GradScaler.unscale_ + clip_grad_norm_(1.0) + GradScaler.step(AdamW) + GradScaler.update as three multi-tensor
//     launches over the ~520 parameter tensors (distributed_trainer_cls.py:88-96) instead of ~60 foreach launches.
// Everything here is HBM-bound: the gradient-norm pass reads 4 B per parameter, the AdamW pass reads 16 B and writes 12 B.
#include "common.cuh"
#include <math.h>
#include "../../include/calm_b200.h"

namespace {

constexpr int OPT_NT = 256;
constexpr int OPT_CHUNK = CALM_OPT_CHUNK;   // elements per CTA

// ------------------------------------------------------------------------------------------------------------------
// soft-target cross entropy: one CTA per row
// row_stats (B,4): loss_b = lse * tsum - sum_c t_c x_c ; lse ; tsum ; 1 if argmax x == argmax t
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float target_at(const float* __restrict__ trow, long long label, int c) {
  return trow ? trow[c] : (c == label ? 1.0f : 0.0f);
}

__global__ void __launch_bounds__(128)
soft_ce_row_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target, long long ld_t,
                   const long long* __restrict__ labels, float* __restrict__ row_stats, int C) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  __shared__ int redi[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* xr = logits + (long long)b * ld;
  const float* tr = target ? target + (long long)b * ld_t : nullptr;
  const long long label = labels ? labels[b] : -1;
  // pass 1: row max of x (with its first index) and of t (with its first index)
  float mx = -INFINITY, mt = -INFINITY;
  int ix = 0x7fffffff, it = 0x7fffffff;
  for (int c = tid; c < C; c += 128) {
    const float x = xr[c], t = target_at(tr, label, c);
    if (x > mx) { mx = x; ix = c; }
    if (t > mt) { mt = t; it = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ox = __shfl_xor_sync(0xffffffffu, mx, o); const int oix = __shfl_xor_sync(0xffffffffu, ix, o);
    if (ox > mx || (ox == mx && oix < ix)) { mx = ox; ix = oix; }
    const float ot = __shfl_xor_sync(0xffffffffu, mt, o); const int oit = __shfl_xor_sync(0xffffffffu, it, o);
    if (ot > mt || (ot == mt && oit < it)) { mt = ot; it = oit; }
  }
  if (lane == 0) { red[wid] = mx; redi[wid] = ix; red[4 + wid] = mt; redi[4 + wid] = it; }
  __syncthreads();
  mx = red[0]; ix = redi[0]; mt = red[4]; it = redi[4];
#pragma unroll
  for (int w = 1; w < 4; ++w) {
    if (red[w] > mx || (red[w] == mx && redi[w] < ix)) { mx = red[w]; ix = redi[w]; }
    if (red[4 + w] > mt || (red[4 + w] == mt && redi[4 + w] < it)) { mt = red[4 + w]; it = redi[4 + w]; }
  }
  // pass 2: sum exp(x - max), sum t, sum t x
  float se = 0.f, st = 0.f, stx = 0.f;
  for (int c = tid; c < C; c += 128) {
    const float x = xr[c], t = target_at(tr, label, c);
    se += __expf(x - mx);
    st += t;
    stx = fmaf(t, x, stx);
  }
  se = block_sum(se, red);
  st = block_sum(st, red);
  stx = block_sum(stx, red);
  if (tid == 0) {
    const float lse = mx + logf(se);
    float* rs = row_stats + (long long)b * 4;
    rs[0] = fmaf(lse, st, -stx);
    rs[1] = lse;
    rs[2] = st;
    rs[3] = (ix == it) ? 1.0f : 0.0f;
  }
}

// loss_out[0] = mean_b loss_b ; loss_out[1] = mean_b correct_b  (fixed summation order)
__global__ void __launch_bounds__(256)
soft_ce_mean_kernel(const float* __restrict__ row_stats, float* __restrict__ loss_out, int B) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  float l = 0.f, a = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) { l += row_stats[(long long)b * 4]; a += row_stats[(long long)b * 4 + 3]; }
  l = block_sum(l, red);
  a = block_sum(a, red);
  if (threadIdx.x == 0) { loss_out[0] = l / (float)B; loss_out[1] = a / (float)B; }
}

// dlogits[b,c] = (softmax(x_b)_c * tsum_b - t[b,c]) * dloss / B
__global__ void __launch_bounds__(128)
soft_ce_bwd_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target, long long ld_t,
                   const long long* __restrict__ labels, const float* __restrict__ row_stats, const float* __restrict__ dloss,
                   float* __restrict__ dlogits, long long ld_d, int B, int C) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int b = blockIdx.x;
  const float* xr = logits + (long long)b * ld;
  const float* tr = target ? target + (long long)b * ld_t : nullptr;
  const long long label = labels ? labels[b] : -1;
  const float lse = row_stats[(long long)b * 4 + 1], tsum = row_stats[(long long)b * 4 + 2];
  const float g = (dloss ? dloss[0] : 1.0f) / (float)B;
  float* dr = dlogits + (long long)b * ld_d;
  for (int c = threadIdx.x; c < C; c += 128)
    dr[c] = (__expf(xr[c] - lse) * tsum - target_at(tr, label, c)) * g;
}

// ------------------------------------------------------------------------------------------------------------------
// Huber(delta) between the generated token image (B,S,S,3) and the NCHW target (B,3,S,S); a thread owns one pixel
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float huber(float d, float delta) {
  const float a = fabsf(d);
  return a < delta ? 0.5f * d * d : delta * (a - 0.5f * delta);
}

__global__ void __launch_bounds__(256)
huber_fwd_kernel(const float* __restrict__ tokens, const float* __restrict__ target, float delta, float* __restrict__ partial,
                 long long npix, int S) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const long long plane = (long long)S * S;
  float acc = 0.f;
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < npix; p += (long long)gridDim.x * 256) {
    const long long b = p / plane, ij = p - b * plane;
    const float* t = tokens + p * 3;
    const float* y = target + b * plane * 3 + ij;
    acc += huber(t[0] - __ldg(y), delta) + huber(t[1] - __ldg(y + plane), delta) + huber(t[2] - __ldg(y + 2 * plane), delta);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// loss_out[1] = huber mean, loss_out[0] = huber mean + kl_weight * kl
__global__ void __launch_bounds__(256)
huber_final_kernel(const float* __restrict__ partial, int nparts, const float* __restrict__ kl, float kl_weight, float inv_n,
                   float* __restrict__ loss_out) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) acc += partial[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float h = acc * inv_n;
    loss_out[1] = h;
    loss_out[0] = kl ? fmaf(kl_weight, kl[0], h) : h;
  }
}

__global__ void __launch_bounds__(256)
huber_bwd_kernel(const float* __restrict__ tokens, const float* __restrict__ target, const float* __restrict__ dloss, float delta,
                 float inv_n, float kl_weight, float* __restrict__ dtokens, float* __restrict__ dkl, long long npix, int S) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const long long plane = (long long)S * S;
  const float gl = dloss ? dloss[0] : 1.0f;
  const float g = gl * inv_n;
  if (dkl && blockIdx.x == 0 && threadIdx.x == 0) dkl[0] = kl_weight * gl;
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < npix; p += (long long)gridDim.x * 256) {
    const long long b = p / plane, ij = p - b * plane;
    const float* t = tokens + p * 3;
    const float* y = target + b * plane * 3 + ij;
    float* d = dtokens + p * 3;
    d[0] = fminf(fmaxf(t[0] - __ldg(y), -delta), delta) * g;
    d[1] = fminf(fmaxf(t[1] - __ldg(y + plane), -delta), delta) * g;
    d[2] = fminf(fmaxf(t[2] - __ldg(y + 2 * plane), -delta), delta) * g;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// multi-tensor optimizer step. A CTA owns one chunk of OPT_CHUNK consecutive elements of one tensor
// (chunk_tensor[c], chunk_start[c] precomputed by the host once); moments live in two flat buffers indexed by elem_off.
// ------------------------------------------------------------------------------------------------------------------
struct OptArgs {
  float* const* params;
  const float* const* grads;
  const long long* elem_off;
  const int* chunk_tensor;
  const int* chunk_start;     // in units of OPT_CHUNK
  float* exp_avg;
  float* exp_avg_sq;
  float* partial;
  float* state;
  int nchunks;
  float beta1, beta2, eps, weight_decay, max_norm, growth_factor, backoff_factor;
  int growth_interval, use_scaler;
};

__global__ void __launch_bounds__(OPT_NT)
opt_gradsq_kernel(OptArgs a) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  const int c = blockIdx.x, t = a.chunk_tensor[c];
  const long long n = a.elem_off[t + 1] - a.elem_off[t];
  const long long begin = (long long)a.chunk_start[c] * OPT_CHUNK;
  const int len = (int)min((long long)OPT_CHUNK, n - begin);
  const float* g = a.grads[t] + begin;
  const float inv = a.use_scaler ? 1.0f / a.state[CALM_OPT_SCALE] : 1.0f;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_NT) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
      const float x = v.x * inv, y = v.y * inv, z = v.z * inv, w = v.w * inv;
      acc += x * x + y * y + z * z + w * w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_NT) { const float x = g[i] * inv; acc = fmaf(x, x, acc); }
  } else {
    for (int i = threadIdx.x; i < len; i += OPT_NT) { const float x = g[i] * inv; acc = fmaf(x, x, acc); }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) a.partial[c] = acc;
}

// one CTA: total norm, non-finite check, clip coefficient, step / bias corrections, GradScaler.update
__global__ void __launch_bounds__(1024)
opt_finalize_kernel(OptArgs a) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < a.nchunks; i += 1024) acc += a.partial[i];
  acc = block_sum(acc, red);
  if (threadIdx.x != 0) return;
  float* st = a.state;
  const float scale = st[CALM_OPT_SCALE];
  const float inv = a.use_scaler ? 1.0f / scale : 1.0f;
  const float norm = sqrtf(acc);
  const bool bad = !isfinite(acc);
  const float clip = fminf(a.max_norm / (norm + 1e-6f), 1.0f);     // clip_grad_norm_: coef clamped to 1
  st[CALM_OPT_GRAD_NORM] = norm;
  st[CALM_OPT_FOUND_INF] = (bad && a.use_scaler) ? 1.0f : 0.0f;   // without a scaler the step is never skipped (torch semantics)
  st[CALM_OPT_MULT] = a.max_norm > 0.f ? inv * clip : inv;
  if (!(bad && a.use_scaler)) {
    const float step = st[CALM_OPT_STEP] + 1.0f;
    st[CALM_OPT_STEP] = step;
    st[CALM_OPT_BIAS1] = 1.0f - powf(a.beta1, step);
    st[CALM_OPT_BIAS2_SQRT] = sqrtf(1.0f - powf(a.beta2, step));
  }
  if (a.use_scaler) {                                             // torch _amp_update_scale_
    if (bad) {
      st[CALM_OPT_SCALE] = scale * a.backoff_factor;
      st[CALM_OPT_GROWTH_TRACKER] = 0.f;
    } else {
      const float ok = st[CALM_OPT_GROWTH_TRACKER] + 1.0f;
      if ((int)ok >= a.growth_interval) {
        const float grown = scale * a.growth_factor;
        if (isfinite(grown)) st[CALM_OPT_SCALE] = grown;
        st[CALM_OPT_GROWTH_TRACKER] = 0.f;
      } else {
        st[CALM_OPT_GROWTH_TRACKER] = ok;
      }
    }
  }
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float mult, float lr, float wd, float b1, float b2,
                                          float eps, float step_size, float bias2_sqrt) {
  g *= mult;
  p -= lr * wd * p;
  m = fmaf(1.0f - b1, g - m, m);
  v = fmaf(b2, v, (1.0f - b2) * g * g);
  const float denom = sqrtf(v) / bias2_sqrt + eps;
  p -= step_size * m / denom;
}

__global__ void __launch_bounds__(OPT_NT)
opt_adamw_kernel(OptArgs a) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const float* st = a.state;
  if (st[CALM_OPT_FOUND_INF] != 0.f) return;      // GradScaler.step: skip the update when a gradient was inf / nan
  const int c = blockIdx.x, t = a.chunk_tensor[c];
  const long long off = a.elem_off[t], n = a.elem_off[t + 1] - off;
  const long long begin = (long long)a.chunk_start[c] * OPT_CHUNK;
  const int len = (int)min((long long)OPT_CHUNK, n - begin);
  float* p = a.params[t] + begin;
  const float* g = a.grads[t] + begin;
  float* m = a.exp_avg + off + begin;
  float* v = a.exp_avg_sq + off + begin;
  const float mult = st[CALM_OPT_MULT], lr = st[CALM_OPT_LR];
  const float step_size = lr / st[CALM_OPT_BIAS1], b2s = st[CALM_OPT_BIAS2_SQRT];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int done = 0;
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_NT) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
      float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adamw_one(pp.x, gg.x, mm.x, vv.x, mult, lr, a.weight_decay, a.beta1, a.beta2, a.eps, step_size, b2s);
      adamw_one(pp.y, gg.y, mm.y, vv.y, mult, lr, a.weight_decay, a.beta1, a.beta2, a.eps, step_size, b2s);
      adamw_one(pp.z, gg.z, mm.z, vv.z, mult, lr, a.weight_decay, a.beta1, a.beta2, a.eps, step_size, b2s);
      adamw_one(pp.w, gg.w, mm.w, vv.w, mult, lr, a.weight_decay, a.beta1, a.beta2, a.eps, step_size, b2s);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < len; i += OPT_NT) {
    float pp = p[i], mm = m[i], vv = v[i];
    adamw_one(pp, g[i], mm, vv, mult, lr, a.weight_decay, a.beta1, a.beta2, a.eps, step_size, b2s);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// input side: MixUp / CutMix of a device-resident batch + the soft labels they produce
// ------------------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
mix_batch_kernel(const float* __restrict__ x, float* __restrict__ out, long long per_image, int H, int W, int B, int mode, float lam,
                 float m, int bx1, int by1, int bx2, int by2) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  constexpr int V = VEC ? 4 : 1;
  const long long n = (long long)B * per_image / V;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const long long e = i * V;
    const long long b = e / per_image, r = e - b * per_image;
    const long long prev = (b == 0 ? (long long)B - 1 : b - 1) * per_image + r;     // images.roll(1, 0)
    if (mode == 0) {
      // torchvision: inpt.roll(1, 0).mul_(1.0 - lam).add_(inpt.mul(lam)) — two rounded products, one rounded sum (no FMA);
      // m = 1.0 - lam comes from the host, evaluated in double before it becomes the fp32 scalar (like the Python expression)
      if (VEC) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + e)), p = __ldg(reinterpret_cast<const float4*>(x + prev));
        *reinterpret_cast<float4*>(out + e) = make_float4(__fadd_rn(__fmul_rn(p.x, m), __fmul_rn(a.x, lam)), __fadd_rn(__fmul_rn(p.y, m), __fmul_rn(a.y, lam)),
                                                          __fadd_rn(__fmul_rn(p.z, m), __fmul_rn(a.z, lam)), __fadd_rn(__fmul_rn(p.w, m), __fmul_rn(a.w, lam)));
      } else {
        out[e] = __fadd_rn(__fmul_rn(x[prev], m), __fmul_rn(x[e], lam));
      }
    } else {
      const int pix = (int)(r % ((long long)H * W));
      const int yy = pix / W, xx = pix - yy * W;
      const bool row_in = yy >= by1 && yy < by2;
      if (VEC) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + e));
        float4 o = a;
        if (row_in && xx + 3 >= bx1 && xx < bx2) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(x + prev));
          if (xx >= bx1 && xx < bx2) o.x = p.x;
          if (xx + 1 >= bx1 && xx + 1 < bx2) o.y = p.y;
          if (xx + 2 >= bx1 && xx + 2 < bx2) o.z = p.z;
          if (xx + 3 >= bx1 && xx + 3 < bx2) o.w = p.w;
        }
        *reinterpret_cast<float4*>(out + e) = o;
      } else {
        out[e] = (row_in && xx >= bx1 && xx < bx2) ? x[prev] : x[e];
      }
    }
  }
}

// soft[b, c] = lam [c == label_b] + (1 - lam) [c == label_{b-1}]  (one-hot labels mixed like the images)
__global__ void __launch_bounds__(256)
mix_labels_kernel(const long long* __restrict__ labels, float* __restrict__ soft, int B, int num_classes, float lam, float m) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization
  const int b = blockIdx.x;
  const long long cur = labels[b], prev = labels[b == 0 ? B - 1 : b - 1];
  for (int c = threadIdx.x; c < num_classes; c += 256)
    soft[(long long)b * num_classes + c] = (c == cur ? lam : 0.0f) + (c == prev ? m : 0.0f);
}

}  // namespace

extern "C" {

int32_t calm_mix_batch(const float* x, const int64_t* labels, float* out, float* soft, int32_t B, int32_t channels, int32_t H,
                       int32_t W, int32_t num_classes, int32_t mode, float lam, float one_minus_lam, int32_t x1, int32_t y1, int32_t x2,
                       int32_t y2, float lam_labels, float one_minus_lam_labels, cudaStream_t stream) {
  CALM_CHECK_ARG(x && out && x != out && B > 0 && channels > 0 && H > 0 && W > 0, "calm_mix_batch: bad arguments (out must not alias x)");
  CALM_CHECK_ARG(mode == 0 || mode == 1, "calm_mix_batch: mode %d (0 = MixUp, 1 = CutMix)", mode);
  CALM_CHECK_ARG((labels == nullptr) == (soft == nullptr) && (!labels || num_classes > 0), "calm_mix_batch: labels and soft go together");
  const long long per_image = (long long)channels * H * W;
  const bool vec = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const long long n = (long long)B * per_image / (vec ? 4 : 1);
  long long blocks = (n + 255) / 256;
  const long long cap = 16LL * calm_num_sms();
  if (blocks > cap) blocks = cap;
  if (vec) CALM_LAUNCH((mix_batch_kernel<true>), (unsigned)blocks, 256, 0, stream, x, out, per_image, H, W, B, mode, lam, one_minus_lam, x1, y1, x2, y2);
  else CALM_LAUNCH((mix_batch_kernel<false>), (unsigned)blocks, 256, 0, stream, x, out, per_image, H, W, B, mode, lam, one_minus_lam, x1, y1, x2, y2);
  CALM_CHECK_LAUNCH("mix_batch_kernel");
  if (labels) {
    CALM_LAUNCH((mix_labels_kernel), B, 256, 0, stream, reinterpret_cast<const long long*>(labels), soft, B, num_classes, lam_labels, one_minus_lam_labels);
    CALM_CHECK_LAUNCH("mix_labels_kernel");
  }
  return CALM_OK;
}

int32_t calm_soft_ce_fwd(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* labels,
                         float* row_stats, float* loss_out, int32_t B, int32_t C, cudaStream_t stream) {
  CALM_CHECK_ARG(logits && row_stats && loss_out && B > 0 && C > 0, "calm_soft_ce_fwd: bad arguments");
  CALM_CHECK_ARG((target != nullptr) != (labels != nullptr), "calm_soft_ce_fwd: exactly one of target / labels");
  CALM_CHECK_ARG(ld >= C && (!target || ld_t >= C), "calm_soft_ce_fwd: leading dimension < C");
  CALM_LAUNCH((soft_ce_row_kernel), B, 128, 0, stream, logits, ld, target, ld_t, reinterpret_cast<const long long*>(labels), row_stats, C);
  CALM_CHECK_LAUNCH("soft_ce_row_kernel");
  CALM_LAUNCH((soft_ce_mean_kernel), 1, 256, 0, stream, row_stats, loss_out, B);
  CALM_CHECK_LAUNCH("soft_ce_mean_kernel");
  return CALM_OK;
}

int32_t calm_soft_ce_bwd(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* labels,
                         const float* row_stats, const float* dloss, float* dlogits, int64_t ld_d, int32_t B, int32_t C,
                         cudaStream_t stream) {
  CALM_CHECK_ARG(logits && row_stats && dlogits && B > 0 && C > 0, "calm_soft_ce_bwd: bad arguments");
  CALM_CHECK_ARG((target != nullptr) != (labels != nullptr), "calm_soft_ce_bwd: exactly one of target / labels");
  CALM_LAUNCH((soft_ce_bwd_kernel), B, 128, 0, stream, logits, ld, target, ld_t, reinterpret_cast<const long long*>(labels), row_stats, dloss,
                                            dlogits, ld_d, B, C);
  CALM_CHECK_LAUNCH("soft_ce_bwd_kernel");
  return CALM_OK;
}

int32_t calm_huber_parts(int32_t B, int32_t S) {
  const long long npix = (long long)B * S * S;
  const long long want = (npix + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)calm_num_sms() * 8;
  return (int32_t)(want < 1 ? 1 : (want > cap ? cap : want));
}

int32_t calm_huber_tokens_fwd(const float* tokens, const float* target_nchw, const float* kl, float kl_weight, float delta,
                              float* partial, int32_t nparts, float* loss_out, int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(tokens && target_nchw && partial && loss_out && B > 0 && S > 0 && nparts > 0, "calm_huber_tokens_fwd: bad arguments");
  const long long npix = (long long)B * S * S;
  CALM_LAUNCH((huber_fwd_kernel), nparts, 256, 0, stream, tokens, target_nchw, delta, partial, npix, S);
  CALM_CHECK_LAUNCH("huber_fwd_kernel");
  CALM_LAUNCH((huber_final_kernel), 1, 256, 0, stream, partial, nparts, kl, kl_weight, 1.0f / (float)(npix * 3), loss_out);
  CALM_CHECK_LAUNCH("huber_final_kernel");
  return CALM_OK;
}

int32_t calm_huber_tokens_bwd(const float* tokens, const float* target_nchw, const float* dloss, float kl_weight, float delta,
                              float* dtokens, float* dkl, int32_t B, int32_t S, cudaStream_t stream) {
  CALM_CHECK_ARG(tokens && target_nchw && dtokens && B > 0 && S > 0, "calm_huber_tokens_bwd: bad arguments");
  const long long npix = (long long)B * S * S;
  CALM_LAUNCH((huber_bwd_kernel), calm_huber_parts(B, S), 256, 0, stream, tokens, target_nchw, dloss, delta, 1.0f / (float)(npix * 3), kl_weight,
                                                             dtokens, dkl, npix, S);
  CALM_CHECK_LAUNCH("huber_bwd_kernel");
  return CALM_OK;
}

int32_t calm_trainer_step(const calm_trainer_step_args* s, cudaStream_t stream) {
  CALM_CHECK_ARG(s && s->params && s->grads && s->elem_off && s->chunk_tensor && s->chunk_start && s->exp_avg && s->exp_avg_sq &&
                 s->partial && s->state, "calm_trainer_step: null pointer");
  CALM_CHECK_ARG(s->n_chunks > 0 && s->n_tensors > 0, "calm_trainer_step: empty table");
  CALM_CHECK_ARG(s->beta1 >= 0.f && s->beta1 < 1.f && s->beta2 >= 0.f && s->beta2 < 1.f && s->eps > 0.f, "calm_trainer_step: bad hyper-parameters");
  if (s->grads_host) {
    cudaError_t e = cudaMemcpyAsync(const_cast<void*>(s->grads), s->grads_host, sizeof(void*) * (size_t)s->n_tensors,
                                    cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { calm_set_error("calm_trainer_step: pointer-table upload failed: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
  }
  OptArgs a;
  a.params = reinterpret_cast<float* const*>(s->params);
  a.grads = reinterpret_cast<const float* const*>(s->grads);
  a.elem_off = reinterpret_cast<const long long*>(s->elem_off);
  a.chunk_tensor = s->chunk_tensor;
  a.chunk_start = s->chunk_start;
  a.exp_avg = s->exp_avg; a.exp_avg_sq = s->exp_avg_sq; a.partial = s->partial; a.state = s->state;
  a.nchunks = s->n_chunks;
  a.beta1 = s->beta1; a.beta2 = s->beta2; a.eps = s->eps; a.weight_decay = s->weight_decay; a.max_norm = s->max_norm;
  a.growth_factor = s->growth_factor; a.backoff_factor = s->backoff_factor;
  a.growth_interval = s->growth_interval; a.use_scaler = s->use_scaler;
  CALM_LAUNCH((opt_gradsq_kernel), s->n_chunks, OPT_NT, 0, stream, a);
  CALM_CHECK_LAUNCH("opt_gradsq_kernel");
  CALM_LAUNCH((opt_finalize_kernel), 1, 1024, 0, stream, a);
  CALM_CHECK_LAUNCH("opt_finalize_kernel");
  CALM_LAUNCH((opt_adamw_kernel), s->n_chunks, OPT_NT, 0, stream, a);
  CALM_CHECK_LAUNCH("opt_adamw_kernel");
  return CALM_OK;
}

}  // extern "C"
