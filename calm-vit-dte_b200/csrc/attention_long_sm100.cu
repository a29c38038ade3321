// Axial attention forward on tcgen05 / TMEM / TMA for LONG rows and WIDE heads: S up to 1024 keys, head_dim up to 128
// (the 384^2 / 512^2 configs of BASELINE.json: S = 240..512, head_dim = 60..128, where attention_sm100.cu's "whole key axis in one
// TMEM accumulator" form does not fit).          O = softmax(Q K^T / sqrt(hd) + bias[b]) V        (Vi_Tools_CNN_less_V2.py:293-298)
//
// Work item = (image, head, 128-query tile). The key axis is walked in chunks of 128 keys, TWICE:
//   pass A: S_c = Q K_c^T (tensor core, TMEM) -> row maximum of x = S_c * scale + bias                (no exponentials, no V)
//   pass B: S_c again                          -> P_c = exp2(x - max) (bf16, shared memory) -> O += P_c V_c (one TMEM accumulator)
// Recomputing Q K^T costs tensor-core time only (+50 % of a pipe that idles under the softmax arithmetic anyway) and removes the
// online-softmax rescale of the O accumulator: no TMEM read-modify-write, no correction warps, the maximum P is scaled by is final.
// Per chunk a shared-memory STAGE holds {K chunk | V chunk | bias chunk}, 2 stages in flight; the bias chunk arrives by TMA with K
// (row-owner threads loading their own bias rows cost one L1 line per lane and instruction) and P OVERWRITES it in place — the
// thread that read bias(row, 16 keys) writes P(row, the same 16 keys) — so P needs no buffer of its own: 32 + 2 x 96 = 224 KB.
// TMEM: S chunk double-buffered (2 x 128 columns: the MMA of chunk c+1 runs under the softmax of chunk c) + O (<= 128) = 384.
// Roles: 16 worker warps (a query row = one TMEM lane; the 4 warps of a lane quadrant split the 16-column pieces of a chunk,
// per-row max / sum partials meet once per pass through the consumed K region) + 1 controller warp (TMA, MMA issue).
// Head dims that are not multiples of 16 / heads that start 8 bytes into a 16-byte TMA granule: as in attention_sm100.cu the boxes
// bring whole 64-column atoms starting at the aligned-down column; the columns of the neighbouring heads are zeroed once per item
// in Q (they then multiply K's garbage by zero), V's extra columns only reach O columns nobody stores.
#include "tcgen05.cuh"
#include "attention_tc.h"

namespace {

using namespace tc;

#define leader (elect_one() != 0u)

constexpr int LW = 512;                  // worker threads: 4 column groups x 4 TMEM lane quadrants x 32 lanes
constexpr int LGROUPS = 4;
constexpr int LTHREADS = LW + 32;        // dQ / dK-dV kernels: workers + one controller warp
constexpr int LCTRL = LW / 32;           // controller warp
constexpr int LFWD_THREADS = LW + 96;    // forward: workers + loader, S-MMA and PV-MMA warps
constexpr int LW_LOAD = LW / 32, LW_SMMA = LW_LOAD + 1, LW_PV = LW_LOAD + 2;
constexpr int KC = 128;                  // keys per chunk
constexpr int ATOM = 16384;              // [128 rows][64 bf16], 128-byte swizzle
constexpr int ST_K = 0, ST_V = 2 * ATOM, ST_B = 4 * ATOM, STAGE_BYTES = 6 * ATOM;
constexpr int Q_BYTES = 2 * ATOM;
constexpr int L_TMEM_COLS = 512;         // power of two >= 2 * 128 (S) + 128 (O)
constexpr uint32_t T_O = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr size_t LONG_FWD_SMEM = Q_BYTES + 2 * STAGE_BYTES + 256 + 1024;

struct LongParams {
  int B, S, heads, hd;
  float scale_log2;
  bf16* o; long long ld_o;
  float* lse;
  int* err_flag;
  CalmTrace trace;
};

struct HeadCols { int col0, shift, hdp; };
__device__ __forceinline__ HeadCols head_cols(int h, int hd) {
  HeadCols hc;
  const int c = h * hd;
  hc.col0 = c & ~7;
  hc.shift = c - hc.col0;
  hc.hdp = (hc.shift + hd + 15) & ~15;
  return hc;
}
__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float* f) {
  float2 t;
  t = unpack_bf16x2(a.x); f[0] = t.x; f[1] = t.y;   t = unpack_bf16x2(a.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(a.z); f[4] = t.x; f[5] = t.y;   t = unpack_bf16x2(a.w); f[6] = t.x; f[7] = t.y;
  t = unpack_bf16x2(b.x); f[8] = t.x; f[9] = t.y;   t = unpack_bf16x2(b.y); f[10] = t.x; f[11] = t.y;
  t = unpack_bf16x2(b.z); f[12] = t.x; f[13] = t.y; t = unpack_bf16x2(b.w); f[14] = t.x; f[15] = t.y;
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

__global__ void __launch_bounds__(LFWD_THREADS, 1)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                     const __grid_constant__ CUtensorMap mB, const LongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sStage = sQ + Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 2 * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  // barriers: 0 q | 1 qz | 2,3 full (V + bias) | 4,5 sready | 6,7 sfree | 8,9 pready | 10,11 pvdone | 12 ofinal | 13 ofree | 14,15 kfull (K)
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, hd = p.hd, heads = p.heads;
  CalmTraceCursor tcur = calm_trace_begin(p.trace, warp == LW_SMMA ? 0 : warp == LW_PV ? 2 : warp == LW_LOAD ? 3 : 1);
  (void)tcur;

  if (threadIdx.x == LW) {
    prefetch_tensormap(&mQ); prefetch_tensormap(&mK); prefetch_tensormap(&mV); prefetch_tensormap(&mB);
    mbar_init(BAR(0), 1); mbar_init(BAR(1), LW);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(2 + s), 1); mbar_init(BAR(4 + s), 1); mbar_init(BAR(6 + s), LW); mbar_init(BAR(8 + s), LW); mbar_init(BAR(10 + s), 1);
    }
    mbar_init(BAR(12), 1); mbar_init(BAR(13), LW); mbar_init(BAR(14), 1); mbar_init(BAR(15), 1);
    fence_barrier_init();
  }
  if (warp == LW_LOAD) tmem_alloc(smem_u32(tmem_slot), L_TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntq = (S + 127) >> 7;            // query tiles per (image, head)
  const int nch = (S + KC - 1) / KC;         // key chunks
  const int T = 2 * nch;                     // steps per item: pass A chunks, then pass B chunks
  const int items = p.B * heads * ntq;

  // Three single-purpose control warps. One warp doing all of it made the CONTROL FLOW the critical path: its blocking waits
  // (pready -> issue P V, pvdone -> issue loads, kfull / sfree -> issue S MMA) ran one after the other, so the next S MMA was
  // only issued ~3.6 us after the previous one (globaltimer trace) while the tensor core needs 0.9 us per step.
  //   loader : TMA loads. K of step t + 2 as soon as the S MMA of step t has retired (K is only read by the S MMA);
  //            V + bias of step t + 2 when step t's stage is done (pass A: the workers read the bias; pass B: its P V retired)
  //   S-MMA  : S_t = Q K_t^T into TMEM buffer t & 1 as soon as K_t landed and the workers drained the buffer (step t - 2)
  //   PV-MMA : O += P_t V_t as soon as the workers wrote P_t
  // A barrier used once per step on a stage completes its j-th phase in the stage's j-th use: parity j & 1.
  if (warp == LW_LOAD) {
    uint32_t nk[2] = {0, 0};                 // K loads issued per stage == uses of the stage started
    uint32_t nvb[2] = {0, 0};                // V / bias loads issued per stage
    uint32_t npv[2] = {0, 0};                // pass-B uses per stage (pvdone phases)
    int pend[2] = {0, 0};                    // the stage's previous use: 0 none, 1 pass A (wait pready), 2 pass B (wait pvdone)
    uint32_t ph_ofree = 0;
    bool first_item = true;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int i = it % ntq, bh = it / ntq;
      const int h = bh % heads, b = bh / heads;
      const HeadCols hc = head_cols(h, hd);
      const int na = (hc.shift + hd + 63) >> 6;                  // 64-column atoms of the head
      auto load_k = [&](int t) {
        const int s = t & 1, c = t % nch;
        if (nk[s]) mbar_wait(BAR(4 + s), (nk[s] - 1) & 1, p.err_flag, 60);          // the S MMA of the stage's previous use retired
        uint8_t* st = sStage + s * STAGE_BYTES;
        if (leader) {
          mbar_expect_tx(BAR(14 + s), (uint32_t)na * ATOM);
          for (int a = 0; a < na; ++a) tma_load_2d(smem_u32(st + ST_K + a * ATOM), &mK, BAR(14 + s), hc.col0 + 64 * a, b * S + c * KC);
        }
        ++nk[s];
      };
      auto load_vb = [&](int t) {                                // V (pass B) and the bias chunk
        const int s = t & 1, c = t % nch;
        const bool pass_b = t >= nch;
        if (pend[s] == 1) mbar_wait(BAR(8 + s), (nvb[s] - 1) & 1, p.err_flag, 61);  // workers done with the previous use's bias
        else if (pend[s] == 2) mbar_wait(BAR(10 + s), (npv[s] - 1) & 1, p.err_flag, 62);   // its P V retired (P and V free)
        uint8_t* st = sStage + s * STAGE_BYTES;
        if (leader) {
          mbar_expect_tx(BAR(2 + s), (uint32_t)((pass_b ? na : 0) + 2) * ATOM);
          if (pass_b)
            for (int a = 0; a < na; ++a) tma_load_2d(smem_u32(st + ST_V + a * ATOM), &mV, BAR(2 + s), hc.col0 + 64 * a, b * S + c * KC);
          for (int a = 0; a < 2; ++a) tma_load_2d(smem_u32(st + ST_B + a * ATOM), &mB, BAR(2 + s), c * KC + 64 * a, b * S + i * 128);
        }
        ++nvb[s];
        pend[s] = pass_b ? 2 : 1;
        if (pass_b) ++npv[s];
      };
      // the loads of the first two steps do not depend on Q / O: they go out before the previous item's epilogue is awaited
      load_k(0); load_k(1);
      load_vb(0); load_vb(1);
      if (!first_item) { mbar_wait(BAR(13), ph_ofree, p.err_flag, 64); ph_ofree ^= 1; }   // O read out, Q (the epilogue's staging) free
      first_item = false;
      if (leader) {
        mbar_expect_tx(BAR(0), (uint32_t)na * ATOM);
        for (int a = 0; a < na; ++a) tma_load_2d(smem_u32(sQ + a * ATOM), &mQ, BAR(0), hc.col0 + 64 * a, b * S + i * 128);
      }
      for (int t = 0; t + 2 < T; ++t) { load_vb(t + 2); load_k(t + 2); }
    }
  } else if (warp == LW_SMMA) {
    uint32_t ph_q = 0, ph_qz = 0, nuse[2] = {0, 0};
    const uint64_t dQ = smem_desc(smem_u32(sQ), 16, 1024);
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int h = (it / ntq) % heads;
      const HeadCols hc = head_cols(h, hd);
      const int nks = hc.hdp >> 4;                               // MMA k-steps over the head dim
      mbar_wait(BAR(0), ph_q, p.err_flag, 65); ph_q ^= 1;
      mbar_wait(BAR(1), ph_qz, p.err_flag, 66); ph_qz ^= 1;        // the workers zeroed the neighbouring heads' columns of Q
      for (int t = 0; t < T; ++t) {
        const int s = t & 1, c = t % nch;
        const int nk = min(KC, S - c * KC);
        mbar_wait(BAR(14 + s), nuse[s] & 1, p.err_flag, 67);                         // K_t landed
        if (nuse[s]) mbar_wait(BAR(6 + s), (nuse[s] - 1) & 1, p.err_flag, 68);       // the workers read S of the buffer's previous use
        ++nuse[s];
        fence_after();
        const uint64_t dK = smem_desc(smem_u32(sStage + s * STAGE_BYTES + ST_K), 16, 1024);
        const uint32_t id_s = idesc_bf16(128, nk, 0, 0);
        if (leader) {
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t off = (uint32_t)((ks >> 2) * (ATOM >> 4) + 2 * (ks & 3));
            mma_bf16(tmem + 128u * s, dQ + off, dK + off, id_s, ks > 0);
          }
          commit(BAR(4 + s));
        }
        if (leader) calm_trace(tcur, 1200 + t);                 // S MMA queued
      }
    }
  } else if (warp == LW_PV) {
    uint32_t nuse[2] = {0, 0};               // uses of the stage seen
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int h = (it / ntq) % heads;
      const HeadCols hc = head_cols(h, hd);
      const uint32_t id_o = idesc_bf16(128, hc.hdp, 0, 1);       // O = P V : A K-major, B (V: keys x hd) MN-major
      for (int t = 0; t < T; ++t) {
        const int s = t & 1, c = t % nch;
        const uint32_t j = nuse[s]++;
        // every phase is awaited in order, pass A included: a parity wait only tells the current phase from the one before it, a
        // warp that skipped ahead by two phases would read "complete" from the wrong one
        mbar_wait(BAR(8 + s), j & 1, p.err_flag, 63);                                // the workers finished step t (pass B: wrote P_t)
        if (t < nch) continue;
        const int nk16 = min(KC, S - c * KC) >> 4;
        mbar_wait(BAR(2 + s), j & 1, p.err_flag, 69);                                // V_t landed (the workers waited for the same phase)
        fence_after();
        const uint64_t dP = smem_desc(smem_u32(sStage + s * STAGE_BYTES + ST_B), 16, 1024);
        const uint64_t dVmn = smem_desc(smem_u32(sStage + s * STAGE_BYTES + ST_V), ATOM, 1024);
        if (leader) {
          for (int kk = 0; kk < nk16; ++kk)
            mma_bf16(tmem + T_O, dP + (uint32_t)((kk >> 2) * (ATOM >> 4) + 2 * (kk & 3)), dVmn + (uint32_t)(kk * 128), id_o, (c | kk) != 0);
          commit(BAR(10 + s));
        }
        if (leader) calm_trace(tcur, 1300 + t);                 // P V queued
      }
      if (leader) commit(BAR(12));                                // every P V of the item has retired: O is final
    }
  } else {
    // ============================ workers ============================
    const int grp = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;                               // query row of the tile == TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph_q = 0, ph_ofinal = 0, ph_full[2] = {0, 0}, ph_sready[2] = {0, 0};
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int i = it % ntq, bh = it / ntq;
      const int h = bh % heads, b = bh / heads;
      const HeadCols hc = head_cols(h, hd);
      const int q = i * 128 + r;
      mbar_wait(BAR(0), ph_q, p.err_flag, 71); ph_q ^= 1;
      if (grp == 0) {                                             // zero [0, shift) and [shift + hd, hdp) of this Q row (8-byte units)
        if (hc.shift) *reinterpret_cast<uint2*>(sQ + swz128(r, 0)) = make_uint2(0u, 0u);
        for (int c = hc.shift + hd; c < hc.hdp; c += 4)
          *reinterpret_cast<uint2*>(sQ + (c >> 6) * ATOM + swz128(r, (c & 63) >> 3) + ((c & 4) << 1)) = make_uint2(0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(BAR(1));
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < T; ++t) {
        const int s = t & 1, c = t % nch;
        const bool pass_b = t >= nch;
        const int nk = min(KC, S - c * KC);
        uint8_t* st = sStage + s * STAGE_BYTES;
        if (threadIdx.x == 0) calm_trace(tcur, 2000 + t);
        mbar_wait(BAR(2 + s), ph_full[s], p.err_flag, 72); ph_full[s] ^= 1;      // the bias chunk (and V) landed
        if (threadIdx.x == 0) calm_trace(tcur, 2100 + t);
        mbar_wait(BAR(4 + s), ph_sready[s], p.err_flag, 73); ph_sready[s] ^= 1;
        fence_after();
        if (threadIdx.x == 0) calm_trace(tcur, 2200 + t);
        // both 16-key pieces of this thread are requested from TMEM before the first wait (one exposed TMEM latency per step, not two)
        uint32_t sr[2][16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int kl = (LGROUPS * j + grp) * 16;
          if (kl < nk) tmem_ld16(trow + 128u * s + kl, sr[j]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int kl = (LGROUPS * j + grp) * 16;
          if (kl < nk) {
            uint8_t* ba = st + ST_B + (kl >> 6) * ATOM;
            const int g8 = (kl & 63) >> 3;
            uint4* p0 = reinterpret_cast<uint4*>(ba + swz128(r, g8));
            uint4* p1 = reinterpret_cast<uint4*>(ba + swz128(r, g8 + 1));
            float bf[16];
            unpack16(*p0, *p1, bf);
            if (!pass_b) {
#pragma unroll
              for (int e = 0; e < 16; ++e) m = fmaxf(m, fmaf(__uint_as_float(sr[j][e]), p.scale_log2, bf[e] * LOG2E));
            } else {
              float pv[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                pv[e] = ex2_approx(fmaf(__uint_as_float(sr[j][e]), p.scale_log2, fmaf(bf[e], LOG2E, -m)));
                l += pv[e];
              }
              *p0 = pack8f(pv);                                   // P overwrites the bias it was computed from
              *p1 = pack8f(pv + 8);
            }
          }
        }
        if (c == nch - 1) {
          // the four column groups of a row combine their partial max (pass A) / sum (pass B) through shared memory.
          // pass A: the V region of the stage (no V is loaded into it before the pready arrivals below); pass B (the item's last
          // step): the Q tile — every S MMA of the item has retired (this step's sready) and the next Q is only loaded after the
          // epilogue below. (Not the K region: the loader refills it as soon as the step's S MMA has retired.)
          float* xchg = reinterpret_cast<float*>(pass_b ? sQ : st + ST_V);
          xchg[grp * 128 + r] = pass_b ? l : m;
          asm volatile("bar.sync 1, %0;" ::"n"(LW) : "memory");
          if (!pass_b) {
#pragma unroll
            for (int gq = 0; gq < LGROUPS; ++gq) m = fmaxf(m, xchg[gq * 128 + r]);
          } else {
            l = 0.f;
#pragma unroll
            for (int gq = 0; gq < LGROUPS; ++gq) l += xchg[gq * 128 + r];     // same order in every thread of the row
          }
          asm volatile("bar.sync 1, %0;" ::"n"(LW) : "memory");               // nobody overwrites the slots before everyone has read
        }
        if (threadIdx.x == 0) calm_trace(tcur, 2300 + t);
        if (pass_b) fence_proxy_async();
        fence_before();
        mbar_arrive(BAR(6 + s));
        mbar_arrive(BAR(8 + s));
      }
      // ---- epilogue: O / l -> bf16 rows through the (consumed) Q atoms, lanes along the rows
      if (threadIdx.x == 0) calm_trace(tcur, 2500);
      mbar_wait(BAR(12), ph_ofinal, p.err_flag, 74); ph_ofinal ^= 1;
      fence_after();
      if (threadIdx.x == 0) calm_trace(tcur, 2600);
      const float inv = 1.0f / l;
      for (int c0 = grp * 16; c0 < hc.hdp; c0 += 16 * LGROUPS) {
        uint32_t orr[16];
        tmem_ld16(trow + T_O + c0, orr);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(orr[e]) * inv;
        uint8_t* atom = sQ + (c0 >> 6) * ATOM;
        const int g8 = (c0 & 63) >> 3;
        *reinterpret_cast<uint4*>(atom + swz128(r, g8)) = pack8f(f);
        *reinterpret_cast<uint4*>(atom + swz128(r, g8 + 1)) = pack8f(f + 8);
      }
      if (q < S && grp == 0) p.lse[((long long)b * heads + h) * S + q] = (m + log2f(l)) * LN2;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + quad) : "memory");
      {
        // warp grp stores rows quad*32 + grp*8 .. +8 of the tile: lane u owns the 8-byte unit [4u, 4u + 4) of the head's columns
        const int tc = hc.shift + 4 * lane;                       // tile column of the unit
        bf16* obase = p.o + (long long)h * hd + 4 * lane;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = quad * 32 + grp * 8 + k;
          const int qq = i * 128 + row;
          if (4 * lane < hd && qq < S) {
            const uint2 u = *reinterpret_cast<const uint2*>(sQ + (tc >> 6) * ATOM + swz128(row, (tc & 63) >> 3) + ((tc & 4) << 1));
            *reinterpret_cast<uint2*>(obase + ((long long)b * S + qq) * p.ld_o) = u;
          }
        }
      }
      fence_before();
      mbar_arrive(BAR(13));
    }
  }
  fence_before();
  __syncthreads();
  if (warp == LW_LOAD) {
    fence_after();
    tmem_dealloc(tmem, L_TMEM_COLS);
  }
}

}  // namespace

bool calm_attention_long_eligible(int B, int S, int heads, int hd, const int64_t* lds, int nlds, const void* const* ptrs, int nptrs) {
  if (B <= 0 || heads <= 0 || S < 16 || S > 4096 || (S & 15) || hd < 4 || hd > 128 || (hd & 3)) return false;
  if ((hd & 7) && hd > 124) return false;               // a head shifted by 4 columns must still fit two 64-column atoms
  for (int i = 0; i < nlds; ++i)
    if (lds[i] % 8) return false;                       // TMA: row pitch multiple of 16 bytes
  for (int i = 0; i < nptrs; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return false;
  return true;
}

int calm_attention_fwd_long(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
                            int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream) {
  LongParams p;
  p.B = B; p.S = S; p.heads = heads; p.hd = hd;
  p.scale_log2 = LOG2E / sqrtf((float)hd);
  p.o = reinterpret_cast<bf16*>(o); p.ld_o = ld_o; p.lse = lse;
  p.err_flag = g_calm_err_flag;
  p.trace = calm_trace_target();
  CUtensorMap mQ, mK, mV, mB;
  int rc;
  const uint64_t rows = (uint64_t)B * S, cols = (uint64_t)heads * hd;
  if ((rc = tc::make_map_2d(&mQ, q, cols, rows, ld_q, 128))) return rc;
  if ((rc = tc::make_map_2d(&mK, k, cols, rows, ld_k, KC))) return rc;
  if ((rc = tc::make_map_2d(&mV, v, cols, rows, ld_v, KC))) return rc;
  if ((rc = tc::make_map_2d(&mB, bias, (uint64_t)S, rows, S, 128))) return rc;   // bias viewed as (B*S rows, S columns)
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LONG_FWD_SMEM);
    if (e != cudaSuccess) { calm_set_error("calm_attention_fwd(long): smem %zu: %s", (size_t)LONG_FWD_SMEM, cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  const int items = B * heads * ((S + 127) / 128);
  const int grid = items < calm_num_sms() ? items : calm_num_sms();
  attn_fwd_long_kernel<<<grid, LFWD_THREADS, LONG_FWD_SMEM, stream>>>(mQ, mK, mV, mB, p);
  CALM_CHECK_LAUNCH("calm_attention_fwd(long)");
  return CALM_OK;
}

// ====================================================================================================================
// backward for long rows / wide heads: two kernels, each recomputes S = Q K^T and dP = dO V^T on the tensor core
//   P = exp2(S * scale + bias - lse), dS = P (dP - delta)
//   kernel 1 (query-major, items (image, head, 128-query tile), 64-key chunks, double-buffered): dQ = sum_c dS_c K_c
//   kernel 2 (key-major, items (image, head, 128-key tile), 128-query steps): dK = sum_i dS_i^T Q_i, dV = sum_i P_i^T dO_i, and every
//            dS tile is TMA-stored (bf16) into the per-head scratch that dbias_reduce sums over the heads
// Recomputing S / dP in both costs tensor-core time only; in exchange every accumulator (dQ; dK, dV) lives in TMEM for the whole
// item: no atomics, no partial-sum buffers, deterministic.
// ====================================================================================================================
namespace {

constexpr int KC2 = 64;                       // keys per step of the dQ kernel
constexpr int ATOM_H = 8192;                  // [64 rows][64 bf16]
constexpr int DQ_ST_K = 0, DQ_ST_V = 2 * ATOM_H, DQ_ST_B = 4 * ATOM_H, DQ_STAGE = 4 * ATOM_H + ATOM;   // K | V | bias -> dS
constexpr int DQ_NS = 3;                      // shared-memory stages of the dQ kernel (the S | dP buffers in TMEM stay two)
constexpr int LDQ_THREADS = LW + 96;          // workers + loader, S / dP MMA and dQ MMA warps
constexpr size_t LONG_DQ_SMEM = 2 * Q_BYTES + DQ_NS * DQ_STAGE + 256 + 1024;

struct LongBwdParams {
  int B, S, heads, hd;
  float scale, scale_log2;
  const float* lse; const float* delta;
  bf16* dq; bf16* dk; bf16* dv; long long ld_dq, ld_dk, ld_dv;
  int* err_flag;
};

// rows [row0, row0 + 8) of a quadrant, staged as bf16 in 128-byte-swizzled atoms `stg` -> global rows (lane u = 8-byte unit [4u, 4u+4))
__device__ __forceinline__ void store_head_rows(const uint8_t* stg, bf16* base /* + head offset */, long long ld, long long g0 /* global row of tile row 0 */,
                                                long long g_end, int row0, const HeadCols& hc, int hd, int lane) {
  const int tc = hc.shift + 4 * lane;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = row0 + k;
    const long long g = g0 + row;
    if (4 * lane < hd && g < g_end) {
      const uint2 u = *reinterpret_cast<const uint2*>(stg + (tc >> 6) * ATOM + swz128(row, (tc & 63) >> 3) + ((tc & 4) << 1));
      *reinterpret_cast<uint2*>(base + g * ld + 4 * lane) = u;
    }
  }
}
// this thread's TMEM row: 16-column pieces c0 = grp*16, grp*16 + 64, ... < hdp, scaled, as bf16 into the staging atoms
__device__ __forceinline__ void stage_tmem_rows(uint8_t* stg, uint32_t taddr, int r, int grp, int hdp, float mul) {
  for (int c0 = grp * 16; c0 < hdp; c0 += 16 * LGROUPS) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    tmem_ld_wait();
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) * mul;
    uint8_t* atom = stg + (c0 >> 6) * ATOM;
    const int g8 = (c0 & 63) >> 3;
    *reinterpret_cast<uint4*>(atom + swz128(r, g8)) = pack8f(f);
    *reinterpret_cast<uint4*>(atom + swz128(r, g8 + 1)) = pack8f(f + 8);
  }
}

__global__ void __launch_bounds__(LDQ_THREADS, 1)
attn_bwd_long_dq_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                        const __grid_constant__ CUtensorMap mDO, const __grid_constant__ CUtensorMap mB, const LongBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + Q_BYTES;
  uint8_t* sStage = sDO + Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + DQ_NS * DQ_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  // Steps are numbered g = 0, 1, 2, ... over the whole life of the CTA (all its items): the shared-memory stage (K | V | bias -> dS)
  // of a step is g % DQ_NS, its TMEM buffer (S | dP) is g & 1. A barrier completes one phase per use of its stage / buffer, so the
  // phase of step g has parity (g / DQ_NS) & 1 resp. (g >> 1) & 1; every role awaits every phase of the barriers it uses, in order.
  // barriers: 0 q (Q + dO) | 1 qz | 2.. full[NS] | 5,6 sready | 7,8 sfree | 9.. dsready[NS] | 12.. dqdone[NS] | 15 final | 16 ofree
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_FULL = 2, B_SREADY = 5, B_SFREE = 7, B_DS = 9, B_DQ = 12, B_FINAL = 15, B_OFREE = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, hd = p.hd, heads = p.heads;
  if (threadIdx.x == LW) {
    prefetch_tensormap(&mQ); prefetch_tensormap(&mK); prefetch_tensormap(&mV); prefetch_tensormap(&mDO); prefetch_tensormap(&mB);
    mbar_init(BAR(0), 1); mbar_init(BAR(1), LW);
    for (int s = 0; s < DQ_NS; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_DS + s), LW); mbar_init(BAR(B_DQ + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_SREADY + s), 1); mbar_init(BAR(B_SFREE + s), LW); }
    mbar_init(BAR(B_FINAL), 1); mbar_init(BAR(B_OFREE), LW);
    fence_barrier_init();
  }
  if (warp == LW_LOAD) tmem_alloc(smem_u32(tmem_slot), L_TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t T_DQ = 256;               // S at 128 buf, dP at 128 buf + 64
  const int ntq = (S + 127) >> 7;
  const int nch = (S + KC2 - 1) / KC2;
  const int items = p.B * heads * ntq;

  // Three single-purpose control warps (see the forward kernel): the loader runs up to DQ_NS steps ahead of the S / dP MMA warp, the
  // dQ MMA warp trails the workers.
  if (warp == LW_LOAD) {
    uint32_t g0 = 0, ph_ofree = 0;
    bool first_item = true;
    for (int it = blockIdx.x; it < items; it += gridDim.x, g0 += (uint32_t)nch) {
      const int i = it % ntq, bh = it / ntq;
      const int h = bh % heads, b = bh / heads;
      const HeadCols hc = head_cols(h, hd);
      const int na = (hc.shift + hd + 63) >> 6;
      auto load_step = [&](int c) {
        const uint32_t g = g0 + (uint32_t)c, st_i = g % DQ_NS, u = g / DQ_NS;
        if (u) mbar_wait(BAR(B_DQ + st_i), (u - 1) & 1, p.err_flag, 81);            // the dQ MMA of the stage's previous use retired
        uint8_t* st = sStage + st_i * DQ_STAGE;
        if (leader) {
          mbar_expect_tx(BAR(B_FULL + st_i), (uint32_t)(2 * na) * ATOM_H + ATOM);
          for (int a = 0; a < na; ++a) {
            tma_load_2d(smem_u32(st + DQ_ST_K + a * ATOM_H), &mK, BAR(B_FULL + st_i), hc.col0 + 64 * a, b * S + c * KC2);
            tma_load_2d(smem_u32(st + DQ_ST_V + a * ATOM_H), &mV, BAR(B_FULL + st_i), hc.col0 + 64 * a, b * S + c * KC2);
          }
          tma_load_2d(smem_u32(st + DQ_ST_B), &mB, BAR(B_FULL + st_i), c * KC2, b * S + i * 128);
        }
      };
      const int npre = nch < DQ_NS ? nch : DQ_NS;
      for (int c = 0; c < npre; ++c) load_step(c);               // these do not depend on Q / dO: out before the previous epilogue is awaited
      if (!first_item) { mbar_wait(BAR(B_OFREE), ph_ofree, p.err_flag, 83); ph_ofree ^= 1; }
      first_item = false;
      if (leader) {
        mbar_expect_tx(BAR(0), (uint32_t)(2 * na) * ATOM);
        for (int a = 0; a < na; ++a) {
          tma_load_2d(smem_u32(sQ + a * ATOM), &mQ, BAR(0), hc.col0 + 64 * a, b * S + i * 128);
          tma_load_2d(smem_u32(sDO + a * ATOM), &mDO, BAR(0), hc.col0 + 64 * a, b * S + i * 128);
        }
      }
      for (int c = npre; c < nch; ++c) load_step(c);
    }
  } else if (warp == LW_SMMA) {
    uint32_t g0 = 0, ph_q = 0, ph_qz = 0;
    const uint64_t dQd = smem_desc(smem_u32(sQ), 16, 1024), dDOd = smem_desc(smem_u32(sDO), 16, 1024);
    for (int it = blockIdx.x; it < items; it += gridDim.x, g0 += (uint32_t)nch) {
      const int h = (it / ntq) % heads;
      const HeadCols hc = head_cols(h, hd);
      const int nks = hc.hdp >> 4;
      mbar_wait(BAR(0), ph_q, p.err_flag, 84); ph_q ^= 1;
      mbar_wait(BAR(1), ph_qz, p.err_flag, 87); ph_qz ^= 1;      // the workers zeroed the neighbouring heads' columns of Q and dO
      for (int c = 0; c < nch; ++c) {
        const uint32_t g = g0 + (uint32_t)c, st_i = g % DQ_NS, u = g / DQ_NS, tb = g & 1, ub = g >> 1;
        const int nk = min(KC2, S - c * KC2);
        mbar_wait(BAR(B_FULL + st_i), u & 1, p.err_flag, 85);
        if (ub) mbar_wait(BAR(B_SFREE + tb), (ub - 1) & 1, p.err_flag, 86);          // the workers drained the buffer's previous use
        fence_after();
        const uint64_t dK = smem_desc(smem_u32(sStage + st_i * DQ_STAGE + DQ_ST_K), 16, 1024);
        const uint64_t dV = smem_desc(smem_u32(sStage + st_i * DQ_STAGE + DQ_ST_V), 16, 1024);
        const uint32_t id_s = idesc_bf16(128, nk, 0, 0);
        if (leader) {
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t oa = (uint32_t)((ks >> 2) * (ATOM >> 4) + 2 * (ks & 3)), ob = (uint32_t)((ks >> 2) * (ATOM_H >> 4) + 2 * (ks & 3));
            mma_bf16(tmem + 128u * tb, dQd + oa, dK + ob, id_s, ks > 0);
          }
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t oa = (uint32_t)((ks >> 2) * (ATOM >> 4) + 2 * (ks & 3)), ob = (uint32_t)((ks >> 2) * (ATOM_H >> 4) + 2 * (ks & 3));
            mma_bf16(tmem + 128u * tb + 64, dDOd + oa, dV + ob, id_s, ks > 0);
          }
          commit(BAR(B_SREADY + tb));
        }
      }
    }
  } else if (warp == LW_PV) {
    // dQ (+)= dS_c K_c over the valid keys of chunk c
    uint32_t g0 = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, g0 += (uint32_t)nch) {
      const int h = (it / ntq) % heads;
      const HeadCols hc = head_cols(h, hd);
      const uint32_t id_dq = idesc_bf16(128, hc.hdp, 0, 1);      // dQ = dS K : A K-major, B (K: keys x hd) MN-major
      for (int c = 0; c < nch; ++c) {
        const uint32_t g = g0 + (uint32_t)c, st_i = g % DQ_NS, u = g / DQ_NS;
        const int nk16 = min(KC2, S - c * KC2) >> 4;
        mbar_wait(BAR(B_DS + st_i), u & 1, p.err_flag, 82);                            // the workers wrote dS_c
        mbar_wait(BAR(B_FULL + st_i), u & 1, p.err_flag, 88);                          // K_c landed (the workers waited for the same phase)
        fence_after();
        const uint64_t dDS = smem_desc(smem_u32(sStage + st_i * DQ_STAGE + DQ_ST_B), 16, 1024);
        const uint64_t dKmn = smem_desc(smem_u32(sStage + st_i * DQ_STAGE + DQ_ST_K), ATOM_H, 1024);
        if (leader) {
          for (int kk = 0; kk < nk16; ++kk) mma_bf16(tmem + T_DQ, dDS + (uint32_t)(2 * kk), dKmn + (uint32_t)(kk * 128), id_dq, (c | kk) != 0);
          commit(BAR(B_DQ + st_i));
        }
      }
      if (leader) commit(BAR(B_FINAL));
    }
  } else {
    const int grp = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t g0 = 0, ph_q = 0, ph_final = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, g0 += (uint32_t)nch) {
      const int i = it % ntq, bh = it / ntq;
      const int h = bh % heads, b = bh / heads;
      const HeadCols hc = head_cols(h, hd);
      const int q = i * 128 + r;
      const bool valid = q < S;
      const long long stat = ((long long)b * heads + h) * S + (valid ? q : 0);
      const float lse2 = p.lse[stat] * LOG2E, dl = p.delta[stat];
      // Neighbouring heads' columns: the K / V chunks carry them un-zeroed, so they are zeroed in Q and dO (once per item): S = Q K^T
      // and dP = dO V^T then ignore them, and in dQ = dS K they only reach columns nobody stores.
      mbar_wait(BAR(0), ph_q, p.err_flag, 94); ph_q ^= 1;
      if (grp < 2) {
        uint8_t* tile = grp == 0 ? sQ : sDO;
        if (hc.shift) *reinterpret_cast<uint2*>(tile + swz128(r, 0)) = make_uint2(0u, 0u);
        for (int c = hc.shift + hd; c < hc.hdp; c += 4)
          *reinterpret_cast<uint2*>(tile + (c >> 6) * ATOM + swz128(r, (c & 63) >> 3) + ((c & 4) << 1)) = make_uint2(0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(BAR(1));
      for (int c = 0; c < nch; ++c) {
        const uint32_t g = g0 + (uint32_t)c, st_i = g % DQ_NS, u = g / DQ_NS, tb = g & 1, ub = g >> 1;
        const int nk = min(KC2, S - c * KC2);
        uint8_t* st = sStage + st_i * DQ_STAGE;
        mbar_wait(BAR(B_FULL + st_i), u & 1, p.err_flag, 91);
        mbar_wait(BAR(B_SREADY + tb), ub & 1, p.err_flag, 92);
        fence_after();
        const int kl = grp * 16;
        if (kl < nk) {
          uint32_t sr[16], dr[16];
          tmem_ld16(trow + 128u * tb + kl, sr);
          tmem_ld16(trow + 128u * tb + 64 + kl, dr);
          uint4* p0 = reinterpret_cast<uint4*>(st + DQ_ST_B + swz128(r, kl >> 3));
          uint4* p1 = reinterpret_cast<uint4*>(st + DQ_ST_B + swz128(r, (kl >> 3) + 1));
          float bf[16], ds[16];
          unpack16(*p0, *p1, bf);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float pj = valid ? ex2_approx(fmaf(__uint_as_float(sr[e]), p.scale_log2, fmaf(bf[e], LOG2E, -lse2))) : 0.f;
            ds[e] = pj * (__uint_as_float(dr[e]) - dl);
          }
          *p0 = pack8f(ds);
          *p1 = pack8f(ds + 8);
        }
        fence_proxy_async();
        fence_before();
        mbar_arrive(BAR(B_SFREE + tb));
        mbar_arrive(BAR(B_DS + st_i));
      }
      mbar_wait(BAR(B_FINAL), ph_final, p.err_flag, 93); ph_final ^= 1;
      fence_after();
      stage_tmem_rows(sQ, trow + T_DQ, r, grp, hc.hdp, p.scale);
      asm volatile("bar.sync %0, 128;" ::"r"(2 + quad) : "memory");
      store_head_rows(sQ, p.dq + (long long)h * hd, p.ld_dq, (long long)b * S + i * 128, (long long)b * S + S, quad * 32 + grp * 8, hc, hd, lane);
      fence_before();
      mbar_arrive(BAR(B_OFREE));
    }
  }
  fence_before();
  __syncthreads();
  if (warp == LW_LOAD) {
    fence_after();
    tmem_dealloc(tmem, L_TMEM_COLS);
  }
}


// bulk tensor store of a swizzled [128][64] bf16 atom (shared -> global; rows / columns beyond the tensor are clipped)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr size_t LONG_DKV_SMEM = 6 * Q_BYTES + 256 + 1024;     // K_j | V_j | Q_i | dO_i | bias -> P | dS

__global__ void __launch_bounds__(LTHREADS, 1)
attn_bwd_long_dkv_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                         const __grid_constant__ CUtensorMap mDO, const __grid_constant__ CUtensorMap mB, const __grid_constant__ CUtensorMap mDS,
                         const LongBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + Q_BYTES;
  uint8_t* sQ = sV + Q_BYTES;
  uint8_t* sDO = sQ + Q_BYTES;
  uint8_t* sP = sDO + Q_BYTES;
  uint8_t* sDS = sP + Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + Q_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  // barriers: 0 kv | 1 kvz | 2 full (Q_i, dO_i, bias tile) | 3 sready | 4 dsready | 5 kvdone | 6 final | 7 ofree
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, hd = p.hd, heads = p.heads;
  if (threadIdx.x == LW) {
    prefetch_tensormap(&mQ); prefetch_tensormap(&mK); prefetch_tensormap(&mV); prefetch_tensormap(&mDO); prefetch_tensormap(&mB); prefetch_tensormap(&mDS);
    mbar_init(BAR(0), 1); mbar_init(BAR(1), LW); mbar_init(BAR(2), 1); mbar_init(BAR(3), 1); mbar_init(BAR(4), LW);
    mbar_init(BAR(5), 1); mbar_init(BAR(6), 1); mbar_init(BAR(7), LW);
    fence_barrier_init();
  }
  if (warp == LCTRL) tmem_alloc(smem_u32(tmem_slot), L_TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t T_S = 0, T_DP = 128, T_DK = 256, T_DV = 384;
  const int nt = (S + 127) >> 7;               // query tiles == key tiles
  const int items = p.B * heads * nt;

  if (warp == LCTRL) {
    uint32_t ph_kv = 0, ph_kvz = 0, ph_full = 0, ph_ds = 0, ph_kvdone = 0, ph_ofree = 0;
    bool first_item = true;
    const uint64_t dQk = smem_desc(smem_u32(sQ), 16, 1024), dKk = smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t dDOk = smem_desc(smem_u32(sDO), 16, 1024), dVk = smem_desc(smem_u32(sV), 16, 1024);
    const uint64_t dPmn = smem_desc(smem_u32(sP), ATOM, 1024), dDSmn = smem_desc(smem_u32(sDS), ATOM, 1024);
    const uint64_t dQmn = smem_desc(smem_u32(sQ), ATOM, 1024), dDOmn = smem_desc(smem_u32(sDO), ATOM, 1024);
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int j = it % nt, bh = it / nt;
      const int h = bh % heads, b = bh / heads;
      const HeadCols hc = head_cols(h, hd);
      const int na = (hc.shift + hd + 63) >> 6;
      const int nks = hc.hdp >> 4;
      const int nkj = min(128, S - j * 128);
      const uint32_t id_s = idesc_bf16(128, nkj, 0, 0);
      const uint32_t id_kv = idesc_bf16(128, hc.hdp, 1, 1);      // dK = dS^T Q, dV = P^T dO : A and B MN-major
      // K_j / V_j: their last readers (the previous item's S / dP MMAs) retired before its last kvdone, which was awaited below
      if (leader) {
        mbar_expect_tx(BAR(0), (uint32_t)(2 * na) * ATOM);
        for (int a = 0; a < na; ++a) {
          tma_load_2d(smem_u32(sK + a * ATOM), &mK, BAR(0), hc.col0 + 64 * a, b * S + j * 128);
          tma_load_2d(smem_u32(sV + a * ATOM), &mV, BAR(0), hc.col0 + 64 * a, b * S + j * 128);
        }
      }
      if (!first_item) { mbar_wait(BAR(7), ph_ofree, p.err_flag, 101); ph_ofree ^= 1; }   // dK / dV read out, the Q / dO staging free
      first_item = false;
      for (int i = 0; i < nt; ++i) {
        const int nq16 = (min(128, S - i * 128) + 15) >> 4;
        if (leader) {
          mbar_expect_tx(BAR(2), (uint32_t)(2 * na + 2) * ATOM);
          for (int a = 0; a < na; ++a) {
            tma_load_2d(smem_u32(sQ + a * ATOM), &mQ, BAR(2), hc.col0 + 64 * a, b * S + i * 128);
            tma_load_2d(smem_u32(sDO + a * ATOM), &mDO, BAR(2), hc.col0 + 64 * a, b * S + i * 128);
          }
          for (int a = 0; a < 2; ++a) tma_load_2d(smem_u32(sP + a * ATOM), &mB, BAR(2), j * 128 + 64 * a, b * S + i * 128);
        }
        if (i == 0) {
          mbar_wait(BAR(0), ph_kv, p.err_flag, 102); ph_kv ^= 1;
          mbar_wait(BAR(1), ph_kvz, p.err_flag, 103); ph_kvz ^= 1;   // the workers zeroed the neighbouring heads' columns of K_j / V_j
        }
        mbar_wait(BAR(2), ph_full, p.err_flag, 104); ph_full ^= 1;
        if (lane == 0) bulk_wait_read0();                         // the previous step's dS store has left sDS
        __syncwarp();
        fence_after();
        if (leader) {
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t off = (uint32_t)((ks >> 2) * (ATOM >> 4) + 2 * (ks & 3));
            mma_bf16(tmem + T_S, dQk + off, dKk + off, id_s, ks > 0);
          }
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t off = (uint32_t)((ks >> 2) * (ATOM >> 4) + 2 * (ks & 3));
            mma_bf16(tmem + T_DP, dDOk + off, dVk + off, id_s, ks > 0);
          }
          commit(BAR(3));
        }
        mbar_wait(BAR(4), ph_ds, p.err_flag, 105); ph_ds ^= 1;    // P and dS of this (query tile, key tile) block are in shared memory
        fence_after();
        if (lane == 0) {                                          // dS_h block -> scratch[b, h, i * 128 .., j * 128 ..] (summed over heads later)
          for (int a = 0; a < 2; ++a)
            if (j * 128 + 64 * a < S) tma_store_3d(&mDS, smem_u32(sDS) + a * ATOM, j * 128 + 64 * a, i * 128, b * heads + h);
          bulk_commit();
        }
        __syncwarp();
        if (leader) {
          for (int kq = 0; kq < nq16; ++kq) {
            const uint32_t acc = (i > 0 || kq > 0) ? 1u : 0u;
            mma_bf16(tmem + T_DV, dPmn + (uint32_t)(128 * kq), dDOmn + (uint32_t)(128 * kq), id_kv, acc);
            mma_bf16(tmem + T_DK, dDSmn + (uint32_t)(128 * kq), dQmn + (uint32_t)(128 * kq), id_kv, acc);
          }
          commit(BAR(5));
          if (i == nt - 1) commit(BAR(6));
        }
        mbar_wait(BAR(5), ph_kvdone, p.err_flag, 106); ph_kvdone ^= 1;   // Q_i / dO_i / P / dS consumed: the next step may load
      }
    }
    if (lane == 0) bulk_wait0();
  } else {
    const int grp = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph_kv = 0, ph_full = 0, ph_sready = 0, ph_final = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int j = it % nt, bh = it / nt;
      const int h = bh % heads, b = bh / heads;
      const HeadCols hc = head_cols(h, hd);
      const int nkj = min(128, S - j * 128);
      mbar_wait(BAR(0), ph_kv, p.err_flag, 111); ph_kv ^= 1;
      if (grp < 2) {
        uint8_t* tile = grp == 0 ? sK : sV;
        if (hc.shift) *reinterpret_cast<uint2*>(tile + swz128(r, 0)) = make_uint2(0u, 0u);
        for (int c = hc.shift + hd; c < hc.hdp; c += 4)
          *reinterpret_cast<uint2*>(tile + (c >> 6) * ATOM + swz128(r, (c & 63) >> 3) + ((c & 4) << 1)) = make_uint2(0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(BAR(1));
      for (int i = 0; i < nt; ++i) {
        const int q = i * 128 + r;
        const bool valid = q < S;
        const long long stat = ((long long)b * heads + h) * S + (valid ? q : 0);
        const float lse2 = p.lse[stat] * LOG2E, dl = p.delta[stat];
        mbar_wait(BAR(2), ph_full, p.err_flag, 112); ph_full ^= 1;
        mbar_wait(BAR(3), ph_sready, p.err_flag, 113); ph_sready ^= 1;
        fence_after();
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int kl = (LGROUPS * jj + grp) * 16;
          if (kl < nkj) {
            uint32_t sr[16], dr[16];
            tmem_ld16(trow + T_S + kl, sr);
            tmem_ld16(trow + T_DP + kl, dr);
            const uint32_t o0 = (uint32_t)((kl >> 6) * ATOM) + swz128(r, (kl & 63) >> 3), o1 = (uint32_t)((kl >> 6) * ATOM) + swz128(r, ((kl & 63) >> 3) + 1);
            float bf[16], pv[16], ds[16];
            unpack16(*reinterpret_cast<const uint4*>(sP + o0), *reinterpret_cast<const uint4*>(sP + o1), bf);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float pj = valid ? ex2_approx(fmaf(__uint_as_float(sr[e]), p.scale_log2, fmaf(bf[e], LOG2E, -lse2))) : 0.f;
              pv[e] = pj;
              ds[e] = pj * (__uint_as_float(dr[e]) - dl);     // rows beyond S contribute exact zeros to dK / dV
            }
            *reinterpret_cast<uint4*>(sP + o0) = pack8f(pv);  *reinterpret_cast<uint4*>(sP + o1) = pack8f(pv + 8);
            *reinterpret_cast<uint4*>(sDS + o0) = pack8f(ds); *reinterpret_cast<uint4*>(sDS + o1) = pack8f(ds + 8);
          }
        }
        fence_proxy_async();
        fence_before();
        mbar_arrive(BAR(4));
      }
      mbar_wait(BAR(6), ph_final, p.err_flag, 114); ph_final ^= 1;
      fence_after();
      stage_tmem_rows(sQ, trow + T_DK, r, grp, hc.hdp, p.scale);
      stage_tmem_rows(sDO, trow + T_DV, r, grp, hc.hdp, 1.0f);
      asm volatile("bar.sync %0, 128;" ::"r"(2 + quad) : "memory");
      const long long g0 = (long long)b * S + j * 128, g_end = (long long)b * S + S;
      store_head_rows(sQ, p.dk + (long long)h * hd, p.ld_dk, g0, g_end, quad * 32 + grp * 8, hc, hd, lane);
      store_head_rows(sDO, p.dv + (long long)h * hd, p.ld_dv, g0, g_end, quad * 32 + grp * 8, hc, hd, lane);
      fence_before();
      mbar_arrive(BAR(7));
    }
  }
  fence_before();
  __syncthreads();
  if (warp == LCTRL) {
    fence_after();
    tmem_dealloc(tmem, L_TMEM_COLS);
  }
}

// 3-D bf16 map {S cols, S rows, B * heads} over the per-head dS scratch, box {64, 128, 1}, 128B swizzle
int make_map_ds_long(CUtensorMap* map, void* base, uint64_t S, uint64_t bh) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { calm_set_error("cuTensorMapEncodeTiled entry point not found"); return CALM_ERR_CUDA; }
  cuuint64_t dims[3] = {S, S, bh};
  cuuint64_t strides[2] = {S * 2, S * S * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { calm_set_error("cuTensorMapEncodeTiled(dS scratch) failed (%d)", (int)r); return CALM_ERR_CUDA; }
  return CALM_OK;
}

}  // namespace

int calm_attention_bwd_long(const void* q, const void* k, const void* v, const void* bias, const void* d_o, const float* lse,
                            const float* delta, void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q, int64_t ld_k,
                            int64_t ld_v, int64_t ld_do, int64_t ld_dq, int64_t ld_dk, int64_t ld_dv, int B, int S, int heads, int hd,
                            cudaStream_t stream) {
  LongBwdParams p;
  p.B = B; p.S = S; p.heads = heads; p.hd = hd;
  p.scale = 1.0f / sqrtf((float)hd);
  p.scale_log2 = p.scale * LOG2E;
  p.lse = lse; p.delta = delta;
  p.dq = reinterpret_cast<bf16*>(dq); p.dk = reinterpret_cast<bf16*>(dk); p.dv = reinterpret_cast<bf16*>(dv);
  p.ld_dq = ld_dq; p.ld_dk = ld_dk; p.ld_dv = ld_dv;
  p.err_flag = g_calm_err_flag;
  CUtensorMap mQ, mK, mV, mDO, mB, mK64, mV64, mDS;
  int rc;
  const uint64_t rows = (uint64_t)B * S, cols = (uint64_t)heads * hd;
  if ((rc = tc::make_map_2d(&mQ, q, cols, rows, ld_q, 128))) return rc;
  if ((rc = tc::make_map_2d(&mDO, d_o, cols, rows, ld_do, 128))) return rc;
  if ((rc = tc::make_map_2d(&mK, k, cols, rows, ld_k, 128))) return rc;
  if ((rc = tc::make_map_2d(&mV, v, cols, rows, ld_v, 128))) return rc;
  if ((rc = tc::make_map_2d(&mK64, k, cols, rows, ld_k, KC2))) return rc;
  if ((rc = tc::make_map_2d(&mV64, v, cols, rows, ld_v, KC2))) return rc;
  if ((rc = tc::make_map_2d(&mB, bias, (uint64_t)S, rows, S, 128))) return rc;
  if ((rc = make_map_ds_long(&mDS, ds_scratch, (uint64_t)S, (uint64_t)B * heads))) return rc;
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_long_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LONG_DQ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_long_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LONG_DKV_SMEM);
    if (e != cudaSuccess) { calm_set_error("calm_attention_bwd(long): smem: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  const int items = B * heads * ((S + 127) / 128);
  const int grid = items < calm_num_sms() ? items : calm_num_sms();
  attn_bwd_long_dq_kernel<<<grid, LDQ_THREADS, LONG_DQ_SMEM, stream>>>(mQ, mK64, mV64, mDO, mB, p);
  CALM_CHECK_LAUNCH("calm_attention_bwd(long dq)");
  attn_bwd_long_dkv_kernel<<<grid, LTHREADS, LONG_DKV_SMEM, stream>>>(mQ, mK, mV, mDO, mB, mDS, p);
  CALM_CHECK_LAUNCH("calm_attention_bwd(long dkv)");
  return calm_attention_dbias_reduce(ds_scratch, dbias, B, S, heads, stream);
}
