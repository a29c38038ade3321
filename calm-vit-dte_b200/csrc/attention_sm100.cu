// Axial attention with a shared additive bias on the Blackwell tensor cores (tcgen05 + TMEM + TMA), forward and backward.
//   O = softmax(Q K^T / sqrt(hd) + bias[b]) V        (Vi_Tools_CNN_less_V2.py:293-298; bias = linear_mask(...), :288-291)
//
// The sequences of this model are short (S = image side <= 256 at 224^2: 224/176/128/80), so a whole key axis fits one
// TMEM accumulator: S = Q K^T for a 128-query tile is ONE tcgen05.mma chain into <= 256 fp32 columns, the softmax needs no
// online rescaling, and P (bf16) goes back to shared memory as the A operand of the P.V MMA. One thread owns one query row
// (= one TMEM lane), so row max / row sum are plain register reductions — no shuffles.
//
// Roles per CTA (one CTA per SM, persistent over (image, head) items): 16 row-worker warps (softmax, epilogues) — warp w reads TMEM
// lanes 32 (w % 4) .. +31, the four warps of a lane quadrant take every 4th 16-column chunk of a row (row max / row sum of the forward
// are combined through four floats of shared memory per row) — plus control warps whose elected lane issues the TMA loads and the MMAs:
// one controller warp in the backward and in the serial forward, three single-purpose warps (loader, S-MMA, PV-MMA) in the pipelined
// forward (attn_fwd_pipe_kernel, S <= 224). Hand-offs are mbarriers: TMA transaction barriers for loads, tcgen05.commit for finished
// MMAs, worker-count arrivals for "P / dS written" and "accumulator read".
//
// Layouts: Q/K/V/dO tiles are TMA boxes {64 columns, rows} with the 128-byte swizzle, i.e. directly the K-major UMMA
// operand layout (rows of 128 B). Head dims 56/44/20 are not multiples of the 16-wide MMA K step: the box still brings 64
// columns (the tail belongs to the next head), the workers zero the tail of the Q (and dO) rows in shared memory, and the
// MMA runs over hd rounded up to 16 — the garbage tail of K/V then multiplies zeros.
// V (keys x hd) serves as the MN-major B operand of P.V; in backward the same P / dS tiles serve K-major (dQ = dS K) and
// MN-major (dK = dS^T Q, dV = P^T dO) without a second copy.
//
// Backward: (image, head) work items; per item and 128-query tile
//   S = Q K^T and dP = dO V^T in key parts of <= 64 columns, two parts in flight  ->  P = exp2(S c + bias - lse), dS = P (dP - delta)
//   dQ = dS K (complete per tile), dK += dS^T Q, dV += P^T dO (accumulated in TMEM over the query tiles)
// TMEM budget (512 columns): 2 x (S part 64 | dP part 64), dQ reuses the first S part | dK 2 x 64 | dV 2 x 64.
// dbias[b] = sum over heads of dS: the bf16 dS tile the MMAs consume is also TMA-stored, per head, into a (B, H, S, S) scratch and
// dbias_reduce_kernel sums the heads afterwards in head order (deterministic; see the note above attn_bwd_tc_kernel).
#include "tcgen05.cuh"
#include "attention_tc.h"

namespace {

using namespace tc;

// `if (leader)` inside the (warp-uniform) controller loops: elect.sync is evaluated at every use, so ptxas knows the branch runs on
// exactly one lane and needs no ELECT / R2UR.BROADCAST loop to feed the uniform-register operands of tcgen05.mma / TMA.
#define leader (elect_one() != 0u)

// backward: 16 worker warps, 4 per TMEM lane quadrant (warps w, w+4, w+8, w+12 share quadrant w%4): each takes one 16-column chunk
// of every 64-key part. With 8 warps (2 per scheduler) the per-part chain TMEM load -> exp2 -> pack -> st.shared -> fence ran with
// no other warp to hide its latencies: 2.2 us per part for 0.8 us of arithmetic (globaltimer trace, tools/attn_trace.py).
constexpr int NWORKERS = 512;
constexpr int NGROUPS = NWORKERS / 128;
constexpr int NTHREADS = NWORKERS + 32;  // + controller warp
constexpr int CTRL_WARP = NWORKERS / 32;
// forward: 16 worker warps (4 per TMEM lane quadrant, each takes every 4th 16-column chunk): the two softmax passes are a serial
// chain per thread, twice as many threads halve it
constexpr int FWD_WORKERS = 512;
constexpr int FWD_GROUPS = FWD_WORKERS / 128;
constexpr int FWD_THREADS = FWD_WORKERS + 32;
constexpr int FWD_CTRL_WARP = FWD_WORKERS / 32;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int KV_BYTES = 256 * 128;     // up to 256 keys x 64 columns bf16
constexpr int QT_BYTES = 128 * 128;     // 128 query rows x 64 columns
constexpr int P_BYTES = 4 * 16384;      // 128 rows x 256 key columns bf16 = 4 swizzle atoms of [128][64]
constexpr int TMEM_COLS = 512;

struct Common {
  int B, S, heads, hd;
  float scale, scale_log2;
  const bf16* bias;          // (B, S, S)
  int* err_flag;
};

// A TMA box must start on a 16-byte boundary of the row, but a head starts at column h*hd (8-byte granularity when
// hd % 8 == 4: 44, 20, ...). The box therefore starts at the aligned-down column; the head's columns sit at
// [shift, shift + hd) of the 64-column tile, shift in {0, 4}, and the MMAs run over hdp = roundup16(shift + hd) columns.
struct HeadCols { int col0, shift, hdp; };
__device__ __forceinline__ HeadCols head_cols(int h, int hd) {
  HeadCols hc;
  const int c = h * hd;
  hc.col0 = c & ~7;
  hc.shift = c - hc.col0;
  hc.hdp = (hc.shift + hd + 15) & ~15;
  return hc;
}
// zero the columns of row r that do not belong to the head: [0, shift) and [shift + hd, hdp)   (8-byte units)
__device__ __forceinline__ void zero_outside(uint8_t* tile, int r, const HeadCols& hc, int hd) {
  if (hc.shift) *reinterpret_cast<uint2*>(tile + swz128(r, 0)) = make_uint2(0u, 0u);
  for (int c = hc.shift + hd; c < hc.hdp; c += 4)
    *reinterpret_cast<uint2*>(tile + swz128(r, c >> 3) + ((c & 4) << 1)) = make_uint2(0u, 0u);
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
// write 16 consecutive key columns [kc, kc+16) of row r into a [128][256] bf16 tile made of 4 swizzled atoms
__device__ __forceinline__ void store_row16(uint8_t* tile, int r, int kc, const float* f) {
  uint8_t* atom = tile + (kc >> 6) * 16384;
  const int g = (kc & 63) >> 3;
  *reinterpret_cast<uint4*>(atom + swz128(r, g)) = pack8f(f);
  *reinterpret_cast<uint4*>(atom + swz128(r, g + 1)) = pack8f(f + 8);
}


// Epilogue store of a [128][hdp] fp32 TMEM tile as bf16 rows of `hd` columns: the row-owner thread drops its 16-column chunks
// into a swizzled [128][64] bf16 staging atom, the two warps of the lane quadrant meet on a named barrier, then each warp
// writes 16 of the quadrant's rows with lanes walking ALONG the rows (8 bytes per lane, two rows per instruction): 2-4 cache
// lines per store instruction instead of 32 (one per lane) when every thread writes its own row.
__device__ __forceinline__ void stage_cols16(uint8_t* atom, int r, const uint32_t* v, int c0, float mul) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * mul;
  *reinterpret_cast<uint4*>(atom + swz128(r, c0 >> 3)) = pack8f(f);
  *reinterpret_cast<uint4*>(atom + swz128(r, (c0 >> 3) + 1)) = pack8f(f + 8);
}
__device__ __forceinline__ void quad_sync128(int quad) { asm volatile("bar.sync %0, 128;" ::"r"(2 + quad) : "memory"); }
// rows [row_begin, row_begin + 16) of the tile (tile-relative), global row g = g0 + row (valid while g < g_end)
template <int NROWS = 16>
__device__ __forceinline__ void store_rows16(const uint8_t* atom, bf16* base, long long ld, int row_begin, long long g0, long long g_end,
                                             const HeadCols& hc, int hd, int lane) {
  const int unit = lane & 15, sub = lane >> 4;    // 8-byte unit along the row, row of the pair
  const int c = hc.shift + 4 * unit;              // tile column of the unit
#pragma unroll
  for (int k = 0; k < NROWS / 2; ++k) {
    const int row = row_begin + 2 * k + sub;
    const long long g = g0 + row;
    if (4 * unit < hd && g < g_end) {
      const uint2 u = *reinterpret_cast<const uint2*>(atom + swz128(row, c >> 3) + ((c & 4) << 1));
      *reinterpret_cast<uint2*>(base + g * ld + 4 * unit) = u;
    }
  }
}

// ====================================================================================================================
// forward
// ====================================================================================================================
struct FwdParams {
  Common c;
  bf16* o; long long ld_o;
  float* lse;  // (B, heads, S), natural log
};

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                   const __grid_constant__ CUtensorMap mB, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;
  uint8_t* sP = sQ + QT_BYTES;
  uint8_t* sB = sP + P_BYTES;                                    // bias rows of the query tile: [128][256] bf16 in 4 swizzled atoms
  float* xchg = reinterpret_cast<float*>(sB + P_BYTES);          // [2][FWD_GROUPS][128]: row max / row sum parts of the column groups
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * FWD_GROUPS * 128);      // kv, q, mma, work
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar_kv = smem_u32(&bars[0]), bar_q = smem_u32(&bars[1]), bar_mma = smem_u32(&bars[2]), bar_work = smem_u32(&bars[3]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Common& c = p.c;
  const int S = c.S, hd = c.hd;

  if (threadIdx.x == FWD_WORKERS) {
    prefetch_tensormap(&mQ); prefetch_tensormap(&mK); prefetch_tensormap(&mV); prefetch_tensormap(&mB);
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_mma, 1); mbar_init(bar_work, FWD_WORKERS);
    fence_barrier_init();
  }
  if (warp == FWD_CTRL_WARP) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (S + 127) >> 7;
  const int natoms = (S + 63) >> 6;
  const int items = c.B * c.heads;

  if (warp == FWD_CTRL_WARP) {
    // ============================ controller: TMA + MMA issue ============================
    // The whole warp walks the control flow (loop counters, descriptors and barrier addresses stay warp-uniform, so they live in
    // uniform registers); the instructions with side effects are issued under `if (leader)` (elect.sync).
    uint32_t ph_kv = 0, ph_q = 0, ph_work = 0;
    const uint32_t id_s = idesc_bf16(128, S, 0, 0);      // S = Q K^T : both K-major
    // descriptors of the (fixed) tiles, built once: per MMA only the low word advances by (byte offset >> 4)
    const uint64_t dQk = smem_desc(smem_u32(sQ), 16, 1024), dKk = smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t dPk = smem_desc(smem_u32(sP), 16, 1024), dVmn = smem_desc(smem_u32(sV), 8192, 1024);
    // Q tile + the tile's bias rows (one transaction barrier); K / V with the item's first tile
    auto issue_q = [&](int it, int i) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols lc = head_cols(h, hd);
      if (leader) {
        mbar_expect_tx(bar_q, QT_BYTES + natoms * 16384u);
        tma_load_2d(smem_u32(sQ), &mQ, bar_q, lc.col0, b * S + i * 128);
        for (int a = 0; a < natoms; ++a) tma_load_2d(smem_u32(sB) + a * 16384, &mB, bar_q, a * 64, b * S + i * 128);
      }
    };
    auto issue_kv = [&](int it) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols lc = head_cols(h, hd);
      if (leader) {
        mbar_expect_tx(bar_kv, 2u * S * 128u);
        tma_load_2d(smem_u32(sK), &mK, bar_kv, lc.col0, b * S);
        tma_load_2d(smem_u32(sV), &mV, bar_kv, lc.col0, b * S);
      }
    };
    if ((int)blockIdx.x < items) { issue_kv(blockIdx.x); issue_q(blockIdx.x, 0); }
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int h = it % c.heads;
      const HeadCols hc = head_cols(h, hd);
      const int hdp = hc.hdp;
      const uint32_t id_o = idesc_bf16(128, hdp, 0, 1);  // O = P V   : A K-major, B (V: keys x hd) MN-major
      for (int i = 0; i < ntiles; ++i) {
        if (i == 0) {
          mbar_wait(bar_kv, ph_kv, c.err_flag, 11); ph_kv ^= 1;
          mbar_wait(bar_work, ph_work, c.err_flag, 13); ph_work ^= 1;   // A: the workers zeroed the K column tail of this item
        }
        mbar_wait(bar_q, ph_q, c.err_flag, 12); ph_q ^= 1;
        fence_after();
        if (leader) {   // one branch per batch, immediates for the operand steps: ~4 instructions per MMA on the issuing lane
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) if (ks < hdp / 16) mma_bf16(tmem, dQk + 2 * ks, dKk + 2 * ks, id_s, ks > 0);
          commit(bar_mma);
        }
        mbar_wait(bar_work, ph_work, c.err_flag, 14); ph_work ^= 1;     // B: P written; S, Q and the bias tile consumed
        fence_after();
        if (leader) {
          const int nk16 = S / 16;
          for (int a = 0; 4 * a < nk16; ++a) {    // 64 keys = one swizzle atom of P, 8 KB of MN-major V
            const uint64_t da = dPk + (uint32_t)(a * 1024), db = dVmn + (uint32_t)(a * 512);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (4 * a + j < nk16) mma_bf16(tmem + 256, da + 2 * j, db + 128 * j, id_o, (a | j) != 0);
          }
          commit(bar_mma);
        }
        // the next tile's Q and bias load under the P.V MMA and the store epilogue; a new item's K / V once P.V has retired
        if (i + 1 < ntiles) {
          issue_q(it, i + 1);
        } else if (it + (int)gridDim.x < items) {
          mbar_wait(bar_mma, 1, c.err_flag, 16);   // second completion of the tile (phases alternate S-ready / O-ready)
          issue_kv(it + gridDim.x);
          issue_q(it + gridDim.x, 0);
        }
        mbar_wait(bar_work, ph_work, c.err_flag, 15); ph_work ^= 1;     // C: epilogue done, P / TMEM free
      }
    }
  } else {
    // ============================ workers: a query row is shared by FWD_GROUPS threads (every 4th 16-column chunk each) ============================
    const int grp = warp >> 2;                    // 0..3: chunks grp, grp + 4, ...
    const int quad = warp & 3;
    const int r = quad * 32 + lane;               // query row within the tile == TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph_kv = 0, ph_q = 0, ph_mma = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols hc = head_cols(h, hd);
      const int hdp = hc.hdp;
      for (int i = 0; i < ntiles; ++i) {
        const int q = i * 128 + r;
        const bool valid = q < S;
        if (i == 0) {
          // the 64-column boxes also bring the neighbouring heads' columns: zeroing them in K (once per item) keeps them out of
          // S = Q K^T; the tails of V only reach O columns nobody stores
          mbar_wait(bar_kv, ph_kv, c.err_flag, 21); ph_kv ^= 1;
          if (grp == 0) zero_outside(sK, r, hc, hd);
          else if (grp == 1 && r + 128 < S) zero_outside(sK, r + 128, hc, hd);
          fence_proxy_async();
          mbar_arrive(bar_work);                                          // A
        }
        mbar_wait(bar_q, ph_q, c.err_flag, 22); ph_q ^= 1;               // the bias tile landed (with Q)
        mbar_wait(bar_mma, ph_mma, c.err_flag, 23); ph_mma ^= 1;
        fence_after();
        // pass 1: maximum of x = s * scale*log2e + bias*log2e over this thread's chunks, then over the row
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kc = (FWD_GROUPS * j + grp) * 16;
          if (kc < S) {
            uint32_t sr[16];
            tmem_ld16(trow + kc, sr);
            const uint8_t* ba = sB + (kc >> 6) * 16384;
            const int g = (kc & 63) >> 3;
            const uint4 b0 = *reinterpret_cast<const uint4*>(ba + swz128(r, g)), b1 = *reinterpret_cast<const uint4*>(ba + swz128(r, g + 1));
            float bf[16];
            unpack16(b0, b1, bf);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) mx = fmaxf(mx, fmaf(__uint_as_float(sr[jj]), c.scale_log2, bf[jj] * LOG2E));
          }
        }
        xchg[grp * 128 + r] = mx;
        asm volatile("bar.sync 1, %0;" ::"n"(FWD_WORKERS) : "memory");
#pragma unroll
        for (int gq = 0; gq < FWD_GROUPS; ++gq) mx = fmaxf(mx, xchg[gq * 128 + r]);
        // pass 2: P = exp2(x - max), partial row sum, bf16 P into the swizzled A-operand tile
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kc = (FWD_GROUPS * j + grp) * 16;
          if (kc < S) {
            uint32_t sr[16];
            tmem_ld16(trow + kc, sr);
            const uint8_t* ba = sB + (kc >> 6) * 16384;
            const int g = (kc & 63) >> 3;
            const uint4 b0 = *reinterpret_cast<const uint4*>(ba + swz128(r, g)), b1 = *reinterpret_cast<const uint4*>(ba + swz128(r, g + 1));
            float bf[16], pv[16];
            unpack16(b0, b1, bf);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              pv[jj] = ex2_approx(fmaf(__uint_as_float(sr[jj]), c.scale_log2, fmaf(bf[jj], LOG2E, -mx)));
              l += pv[jj];
            }
            store_row16(sP, r, kc, pv);
          }
        }
        xchg[(FWD_GROUPS + grp) * 128 + r] = l;
        fence_proxy_async();
        fence_before();
        mbar_arrive(bar_work);                                            // B
        mbar_wait(bar_mma, ph_mma, c.err_flag, 24); ph_mma ^= 1;           // (all 256 workers arrived at B before the MMA ran)
        fence_after();
        l = 0.f;
#pragma unroll
        for (int gq = 0; gq < FWD_GROUPS; ++gq) l += xchg[(FWD_GROUPS + gq) * 128 + r];   // same order in every thread of the row
        const float inv = 1.0f / l;
        // O rows: TMEM -> staging atom 0 of sP (P has been consumed) -> rows written with the lanes along the row
        for (int c0 = grp * 16; c0 < hdp; c0 += 16 * FWD_GROUPS) {
          uint32_t orr[16];
          tmem_ld16(trow + 256 + c0, orr);
          tmem_ld_wait();
          stage_cols16(sP, r, orr, c0, inv);
        }
        if (valid && grp == 0) p.lse[((long long)b * c.heads + h) * S + q] = (mx + log2f(l)) * LN2;
        quad_sync128(quad);
        store_rows16<8>(sP, p.o + (long long)h * hd, p.ld_o, quad * 32 + grp * 8, (long long)b * S + i * 128, (long long)b * S + S, hc, hd, lane);
        quad_sync128(quad);   // the partner warps rewrite these staging rows with the next tile's P
        fence_before();
        mbar_arrive(bar_work);                                            // C
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == FWD_CTRL_WARP) {
    fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ====================================================================================================================
// forward, pipelined variant (S <= 224: two S accumulators of S columns + O fit the 512 TMEM columns)
// ====================================================================================================================
// The kernel above runs the phases of a tile one after the other — Q / bias load -> S MMA -> softmax -> P.V MMA -> epilogue — and the next
// tile starts when the last one ended (5.6 us per 128-query tile at S = 224, of which the workers compute ~3). Here S is
// double-buffered in TMEM and the control flow is split over three single-purpose warps, so that the load of Q_{t+1}, the S MMA of
// tile t+1 and the load of the next item's K run under the softmax of tile t:
//   loader : per tile  sready_t -> Q_{t+1} (and the next item's K: zeroes its neighbouring-head columns itself -> k_ready),
//                      wdone_t  -> bias_{t+1},   oready_t -> the next item's V
//   S-MMA  : q_full_t, k_ready, sfree (the workers left S buffer t & 1, tile t-2) -> S_t = Q_t K^T -> sready_t
//   PV-MMA : wdone_t (P_t written; it follows the workers' epilogue of tile t-1, so O is free), v_full -> O = P_t V -> oready_t
//   workers: b_full_t, sready_t -> softmax (as above) -> wdone_t -> oready_t -> epilogue
// Tiles are numbered t = 0, 1, ... over the life of the CTA; a barrier completes once per tile (or per item: k_*, v_full): the phase
// of tile t has parity t & 1 (S buffers: (t >> 1) & 1). Every role awaits every phase of its barriers in order.
constexpr int F2_THREADS = FWD_WORKERS + 96;
constexpr int F2_LOAD = FWD_WORKERS / 32, F2_SMMA = F2_LOAD + 1, F2_PV = F2_LOAD + 2;
constexpr uint32_t F2_SCOLS = 224, F2_TO = 448;     // S buffer b at TMEM column 224 b, O at 448

__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd_pipe_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                     const __grid_constant__ CUtensorMap mB, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;
  uint8_t* sP = sQ + QT_BYTES;
  uint8_t* sB = sP + P_BYTES;
  float* xchg = reinterpret_cast<float*>(sB + P_BYTES);          // [2][FWD_GROUPS][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * FWD_GROUPS * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  // barriers: 0 k_full | 1 k_ready | 2 v_full | 3 q_full | 4 b_full | 5,6 sready | 7 wdone | 8 oready | 9,10 sfree
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_KFULL = 0, B_KREADY = 1, B_VFULL = 2, B_QFULL = 3, B_BFULL = 4, B_SREADY = 5, B_WDONE = 7, B_OREADY = 8, B_SFREE = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Common& c = p.c;
  const int S = c.S, hd = c.hd;

  if (threadIdx.x == FWD_WORKERS) {
    prefetch_tensormap(&mQ); prefetch_tensormap(&mK); prefetch_tensormap(&mV); prefetch_tensormap(&mB);
    mbar_init(BAR(B_KFULL), 1); mbar_init(BAR(B_KREADY), 1); mbar_init(BAR(B_VFULL), 1); mbar_init(BAR(B_QFULL), 1); mbar_init(BAR(B_BFULL), 1);
    mbar_init(BAR(B_SREADY), 1); mbar_init(BAR(B_SREADY + 1), 1); mbar_init(BAR(B_WDONE), FWD_WORKERS); mbar_init(BAR(B_OREADY), 1);
    mbar_init(BAR(B_SFREE), FWD_WORKERS); mbar_init(BAR(B_SFREE + 1), FWD_WORKERS);
    fence_barrier_init();
  }
  if (warp == F2_LOAD) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (S + 127) >> 7;
  const int natoms = (S + 63) >> 6;
  const int items = c.B * c.heads;

  if (warp == F2_LOAD) {
    // ============================ loader ============================
    auto load_q = [&](int it, int i) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols lc = head_cols(h, hd);
      if (leader) {
        mbar_expect_tx(BAR(B_QFULL), QT_BYTES);
        tma_load_2d(smem_u32(sQ), &mQ, BAR(B_QFULL), lc.col0, b * S + i * 128);
      }
    };
    auto load_bias = [&](int it, int i) {
      const int b = it / c.heads;
      if (leader) {
        mbar_expect_tx(BAR(B_BFULL), natoms * 16384u);
        for (int a = 0; a < natoms; ++a) tma_load_2d(smem_u32(sB) + a * 16384, &mB, BAR(B_BFULL), a * 64, b * S + i * 128);
      }
    };
    auto load_k = [&](int it) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols lc = head_cols(h, hd);
      if (leader) {
        mbar_expect_tx(BAR(B_KFULL), (uint32_t)S * 128u);
        tma_load_2d(smem_u32(sK), &mK, BAR(B_KFULL), lc.col0, b * S);
      }
    };
    auto load_v = [&](int it) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols lc = head_cols(h, hd);
      if (leader) {
        mbar_expect_tx(BAR(B_VFULL), (uint32_t)S * 128u);
        tma_load_2d(smem_u32(sV), &mV, BAR(B_VFULL), lc.col0, b * S);
      }
    };
    // the 64-column boxes also bring the neighbouring heads' columns: zeroing them in K (once per item) keeps them out of
    // S = Q K^T; the tails of V only reach O columns nobody stores
    auto finish_k = [&](int it, uint32_t n_item) {
      const int h = it % c.heads;
      const HeadCols hc = head_cols(h, hd);
      mbar_wait(BAR(B_KFULL), n_item & 1, c.err_flag, 17);
      for (int row = lane; row < S; row += 32) zero_outside(sK, row, hc, hd);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_KREADY));
      __syncwarp();
    };
    uint32_t t = 0, n_item = 0;
    if ((int)blockIdx.x < items) {
      load_k(blockIdx.x); load_v(blockIdx.x); load_q(blockIdx.x, 0); load_bias(blockIdx.x, 0);
      finish_k(blockIdx.x, 0);
    }
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_item) {
      const int nit = it + (int)gridDim.x;
      for (int i = 0; i < ntiles; ++i, ++t) {
        const bool last = i == ntiles - 1;
        mbar_wait(BAR(B_SREADY + (t & 1)), (t >> 1) & 1, c.err_flag, 11);           // S MMA of tile t retired: Q (and, after the item's last tile, K) free
        if (!last) load_q(it, i + 1);
        else if (nit < items) { load_k(nit); load_q(nit, 0); finish_k(nit, n_item + 1); }
        mbar_wait(BAR(B_WDONE), t & 1, c.err_flag, 12);                             // the workers consumed the bias tile
        if (!last) load_bias(it, i + 1);
        else if (nit < items) load_bias(nit, 0);
        mbar_wait(BAR(B_OREADY), t & 1, c.err_flag, 13);                            // P.V of tile t retired: after the item's last tile V is free
        if (last && nit < items) load_v(nit);
      }
    }
  } else if (warp == F2_SMMA) {
    // ============================ S = Q K^T ============================
    const uint32_t id_s = idesc_bf16(128, S, 0, 0);
    const uint64_t dQk = smem_desc(smem_u32(sQ), 16, 1024), dKk = smem_desc(smem_u32(sK), 16, 1024);
    uint32_t t = 0, n_item = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_item) {
      const int h = it % c.heads;
      const int hdp = head_cols(h, hd).hdp;
      mbar_wait(BAR(B_KREADY), n_item & 1, c.err_flag, 14);
      for (int i = 0; i < ntiles; ++i, ++t) {
        mbar_wait(BAR(B_QFULL), t & 1, c.err_flag, 15);
        // The workers left this S buffer (tile t - 2). A barrier of its own per buffer: a parity wait is only right while the waiter is
        // in step with the barrier, and wdone may already be a phase further (tile t - 1 done) when this warp comes back from a K wait.
        if (t >= 2) mbar_wait(BAR(B_SFREE + (t & 1)), ((t >> 1) - 1) & 1, c.err_flag, 16);
        fence_after();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) if (ks < hdp / 16) mma_bf16(tmem + F2_SCOLS * (t & 1), dQk + 2 * ks, dKk + 2 * ks, id_s, ks > 0);
          commit(BAR(B_SREADY + (t & 1)));
        }
      }
    }
  } else if (warp == F2_PV) {
    // ============================ O = P V ============================
    const uint64_t dPk = smem_desc(smem_u32(sP), 16, 1024), dVmn = smem_desc(smem_u32(sV), 8192, 1024);
    uint32_t t = 0, n_item = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_item) {
      const int h = it % c.heads;
      const uint32_t id_o = idesc_bf16(128, head_cols(h, hd).hdp, 0, 1);   // A K-major, B (V: keys x hd) MN-major
      mbar_wait(BAR(B_VFULL), n_item & 1, c.err_flag, 18);
      for (int i = 0; i < ntiles; ++i, ++t) {
        mbar_wait(BAR(B_WDONE), t & 1, c.err_flag, 19);                             // P_t written (and O of tile t - 1 read out before that)
        fence_after();
        if (leader) {
          const int nk16 = S / 16;
          for (int a = 0; 4 * a < nk16; ++a) {
            const uint64_t da = dPk + (uint32_t)(a * 1024), db = dVmn + (uint32_t)(a * 512);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (4 * a + j < nk16) mma_bf16(tmem + F2_TO, da + 2 * j, db + 128 * j, id_o, (a | j) != 0);
          }
          commit(BAR(B_OREADY));
        }
      }
    }
  } else {
    // ============================ workers (softmax as in attn_fwd_tc_kernel) ============================
    const int grp = warp >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t t = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = it / c.heads, h = it - b * c.heads;
      const HeadCols hc = head_cols(h, hd);
      const int hdp = hc.hdp;
      for (int i = 0; i < ntiles; ++i, ++t) {
        const int q = i * 128 + r;
        const bool valid = q < S;
        const uint32_t ts = trow + F2_SCOLS * (t & 1);
        mbar_wait(BAR(B_BFULL), t & 1, c.err_flag, 22);
        mbar_wait(BAR(B_SREADY + (t & 1)), (t >> 1) & 1, c.err_flag, 23);
        fence_after();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kc = (FWD_GROUPS * j + grp) * 16;
          if (kc < S) {
            uint32_t sr[16];
            tmem_ld16(ts + kc, sr);
            const uint8_t* ba = sB + (kc >> 6) * 16384;
            const int g = (kc & 63) >> 3;
            const uint4 b0 = *reinterpret_cast<const uint4*>(ba + swz128(r, g)), b1 = *reinterpret_cast<const uint4*>(ba + swz128(r, g + 1));
            float bf[16];
            unpack16(b0, b1, bf);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) mx = fmaxf(mx, fmaf(__uint_as_float(sr[jj]), c.scale_log2, bf[jj] * LOG2E));
          }
        }
        xchg[grp * 128 + r] = mx;
        asm volatile("bar.sync 1, %0;" ::"n"(FWD_WORKERS) : "memory");
#pragma unroll
        for (int gq = 0; gq < FWD_GROUPS; ++gq) mx = fmaxf(mx, xchg[gq * 128 + r]);
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kc = (FWD_GROUPS * j + grp) * 16;
          if (kc < S) {
            uint32_t sr[16];
            tmem_ld16(ts + kc, sr);
            const uint8_t* ba = sB + (kc >> 6) * 16384;
            const int g = (kc & 63) >> 3;
            const uint4 b0 = *reinterpret_cast<const uint4*>(ba + swz128(r, g)), b1 = *reinterpret_cast<const uint4*>(ba + swz128(r, g + 1));
            float bf[16], pv[16];
            unpack16(b0, b1, bf);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              pv[jj] = ex2_approx(fmaf(__uint_as_float(sr[jj]), c.scale_log2, fmaf(bf[jj], LOG2E, -mx)));
              l += pv[jj];
            }
            store_row16(sP, r, kc, pv);
          }
        }
        xchg[(FWD_GROUPS + grp) * 128 + r] = l;
        fence_proxy_async();
        fence_before();
        mbar_arrive(BAR(B_SFREE + (t & 1)));
        mbar_arrive(BAR(B_WDONE));
        mbar_wait(BAR(B_OREADY), t & 1, c.err_flag, 24);                // (every worker arrived at wdone before the MMA ran)
        fence_after();
        l = 0.f;
#pragma unroll
        for (int gq = 0; gq < FWD_GROUPS; ++gq) l += xchg[(FWD_GROUPS + gq) * 128 + r];   // same order in every thread of the row
        const float inv = 1.0f / l;
        for (int c0 = grp * 16; c0 < hdp; c0 += 16 * FWD_GROUPS) {
          uint32_t orr[16];
          tmem_ld16(trow + F2_TO + c0, orr);
          tmem_ld_wait();
          stage_cols16(sP, r, orr, c0, inv);
        }
        if (valid && grp == 0) p.lse[((long long)b * c.heads + h) * S + q] = (mx + log2f(l)) * LN2;
        quad_sync128(quad);
        store_rows16<8>(sP, p.o + (long long)h * hd, p.ld_o, quad * 32 + grp * 8, (long long)b * S + i * 128, (long long)b * S + S, hc, hd, lane);
        quad_sync128(quad);   // the partner warps rewrite these staging rows with the next tile's P
        // the next pass-1 exchange overwrites xchg[0..]: every thread has passed the bar.sync above AFTER reading the maxima, and the
        // sums xchg[FWD_GROUPS..] were read after oready, i.e. before any thread of the next tile writes them (bar.sync 1 in between)
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == F2_LOAD) {
    fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ====================================================================================================================
// backward
// ====================================================================================================================
struct BwdParams {
  Common c;
  const float* lse; const float* delta;  // (B, heads, S)
  bf16* dq; bf16* dk; bf16* dv; long long ld_dq, ld_dk, ld_dv;
  CalmTrace trace;                           // bring-up: CTA 0 time stamps (calm_debug_set_trace_buffer), null in production
};

// bring-up time stamps: id = role * 1000 + point * 10 + part (see attention_tc.h; `tcur` is the kernel's per-thread cursor)
#define trace_evt(p, id) calm_trace(tcur, id)

constexpr int KPART = 64;  // key columns per S / dP part; two parts in flight (TMEM: 2 x (64 + 64) | dK 2 x 64 | dV 2 x 64)

// bulk tensor store of a swizzled [128][64] bf16 atom (shared -> global, rows / columns beyond the tensor are clipped)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}

// Pipeline of one (image, head, 128-query tile):
//   controller: S/dP MMAs of parts 0 and 1 at once (two TMEM buffers); the MMAs of part p + 2 as soon as the workers have
//               drained buffer p & 1 (bar_free) -> the tensor core and the mbarrier round trip run under the workers'
//               exp2 / dS arithmetic of the other buffer instead of in series with it.
//   workers   : per part wait bar_sd[buf] -> P, dS -> bf16 tiles in shared memory, dbias accumulation, arrive bar_free[buf].
//   tile end  : dQ = dS K (into the drained buffer 0), dK += dS^T Q, dV += P^T dO; commit bar_fin -> workers read dQ (and dK / dV
//               after the last query tile) out of TMEM -> bar_acc (the next tile's MMAs may start) -> staged, coalesced global stores.
// dbias = sum over heads of dS: the bf16 dS tile the MMAs consume is also TMA-stored, per head, into a (B, H, S, S) scratch;
// dbias_reduce_kernel sums the heads afterwards. (Accumulating in the row-owner threads cost one 128-byte line per lane and
// instruction: 8 warps x 8 scattered 16-byte RMWs per part kept the L1 busy for ~2 us of every part; the TMA store moves
// whole swizzle atoms and costs the workers nothing.)
// The bias row segment of the NEXT part (next tile / head / image at a tile's last part) is requested one part ahead.
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                   const __grid_constant__ CUtensorMap mDO, const __grid_constant__ CUtensorMap mDS, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;
  uint8_t* sDO = sQ + QT_BYTES;
  uint8_t* sP = sDO + QT_BYTES;
  uint8_t* sDS = sP + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + P_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const uint32_t bar_kv = smem_u32(&bars[0]), bar_q = smem_u32(&bars[1]), bar_fin = smem_u32(&bars[2]), bar_tile = smem_u32(&bars[3]);
  const uint32_t bar_acc = smem_u32(&bars[8]);   // the workers have read this tile's accumulators (dQ; dK / dV after the last tile) out of TMEM
  const uint32_t bar_sd0 = smem_u32(&bars[4]), bar_free0 = smem_u32(&bars[6]);   // [buf] at +8 bytes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Common& c = p.c;
  const int S = c.S, hd = c.hd, heads = c.heads;
  CalmTraceCursor tcur = calm_trace_begin(p.trace, warp == CTRL_WARP ? 0 : 1);
  (void)tcur;

  if (threadIdx.x == NWORKERS) {
    prefetch_tensormap(&mQ); prefetch_tensormap(&mK); prefetch_tensormap(&mV); prefetch_tensormap(&mDO); prefetch_tensormap(&mDS);
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_fin, 1); mbar_init(bar_tile, NWORKERS); mbar_init(bar_acc, NWORKERS);
    mbar_init(bar_sd0, 1); mbar_init(bar_sd0 + 8, 1); mbar_init(bar_free0, NWORKERS); mbar_init(bar_free0 + 8, NWORKERS);
    fence_barrier_init();
  }
  if (warp == CTRL_WARP) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (S + 127) >> 7;          // query tiles == key M-tiles
  const int nparts = (S + KPART - 1) / KPART;
  constexpr uint32_t T_SD = 0 /* buffer b: S at 128 b, dP at 128 b + 64 */, T_DQ = 0 /* aliases buffer 0 */, T_DK = 256, T_DV = 384;

  if (warp == CTRL_WARP) {
    {
      // The whole warp walks the control flow (loop counters, descriptors and barrier addresses stay warp-uniform, so they live
      // in uniform registers); only the instructions with side effects are issued by lane 0. Under `if (lane == 0) { loops }`
      // ptxas moves every tcgen05.mma operand through an ELECT / R2UR.BROADCAST loop: ~20 instructions per MMA.
      uint32_t ph_kv = 0, ph_q = 0, ph_tile = 0, ph_fin = 0, ph_acc = 0, ph_free[2] = {0, 0};
      // The tiles never move: their UMMA descriptors are built once, an MMA's operand is base + (byte offset >> 4) in the
      // low word (the 14-bit address field cannot carry: every tile lies below 256 KB). One thread issues ~80 MMAs per
      // query tile; rebuilding two descriptors per MMA made that thread, not the tensor core, the critical path.
      const uint64_t dQk = smem_desc(smem_u32(sQ), 16, 1024), dKk = smem_desc(smem_u32(sK), 16, 1024);        // K-major (S = Q K^T)
      const uint64_t dDOk = smem_desc(smem_u32(sDO), 16, 1024), dVk = smem_desc(smem_u32(sV), 16, 1024);      // K-major (dP = dO V^T)
      const uint64_t dDSk = smem_desc(smem_u32(sDS), 16, 1024);                                                // K-major A of dQ = dS K
      const uint64_t dKmn = smem_desc(smem_u32(sK), 8192, 1024);                                               // MN-major B of dQ
      const uint64_t dDSmn = smem_desc(smem_u32(sDS), 16384, 1024), dPmn = smem_desc(smem_u32(sP), 16384, 1024);  // MN-major A of dK / dV
      const uint64_t dQmn = smem_desc(smem_u32(sQ), 8192, 1024), dDOmn = smem_desc(smem_u32(sDO), 8192, 1024);    // MN-major B of dK / dV
      // loads of one (image, head, tile): Q and dO always, K and V with the head's first tile; the following tile goes to the L2
      auto issue_loads = [&](int bb, int hh, int ii) {
        const HeadCols lc = head_cols(hh, hd);
        if (ii == 0) {
          if (leader) mbar_expect_tx(bar_kv, 2u * S * 128u);
          if (leader) tma_load_2d(smem_u32(sK), &mK, bar_kv, lc.col0, bb * S);
          if (leader) tma_load_2d(smem_u32(sV), &mV, bar_kv, lc.col0, bb * S);
        }
        if (leader) mbar_expect_tx(bar_q, 2 * QT_BYTES);
        if (leader) tma_load_2d(smem_u32(sQ), &mQ, bar_q, lc.col0, bb * S + ii * 128);
        if (leader) tma_load_2d(smem_u32(sDO), &mDO, bar_q, lc.col0, bb * S + ii * 128);
        if (ii + 1 < ntiles) {
          if (leader) tma_prefetch_2d(&mQ, lc.col0, bb * S + (ii + 1) * 128);
          if (leader) tma_prefetch_2d(&mDO, lc.col0, bb * S + (ii + 1) * 128);
        } else {
          const int nit = bb * heads + hh + (int)gridDim.x;   // this CTA's next (image, head) item
          const int nb = nit / heads, nh = nit - nb * heads;
          if (nb < c.B) {
            const HeadCols nc = head_cols(nh, hd);
            if (leader) tma_prefetch_2d(&mK, nc.col0, nb * S);
            if (leader) tma_prefetch_2d(&mV, nc.col0, nb * S);
            if (leader) tma_prefetch_2d(&mQ, nc.col0, nb * S);
            if (leader) tma_prefetch_2d(&mDO, nc.col0, nb * S);
          }
        }
      };
      // work items are (image, head) pairs, head fastest: the heads of an image run on neighbouring CTAs at the same time (its bias
      // rows are shared through the L2) and B * heads items spread over the SMs without the 2-wave tail B = 256 images left
      const int items = c.B * heads;
      if ((int)blockIdx.x < items) issue_loads((int)blockIdx.x / heads, (int)blockIdx.x % heads, 0);
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        {
          const int b = it / heads, h = it - b * heads;
          const HeadCols hc = head_cols(h, hd);
          const int hdp = hc.hdp;
          const int nks = hdp / 16;
          const uint32_t id_dq = idesc_bf16(128, hdp, 0, 1);   // dQ = dS K     : A K-major, B (K: keys x hd) MN-major
          const uint32_t id_dkv = idesc_bf16(128, hdp, 1, 1);  // dK = dS^T Q   : A MN-major, B (Q: queries x hd) MN-major
          const uint32_t id_s_full = idesc_bf16(128, KPART, 0, 0), id_s_tail = idesc_bf16(128, S - (nparts - 1) * KPART, 0, 0);
          for (int i = 0; i < ntiles; ++i) {
            if (i == 0) {
              mbar_wait(bar_kv, ph_kv, c.err_flag, 31); ph_kv ^= 1;
              mbar_wait(bar_tile, ph_tile, c.err_flag, 33); ph_tile ^= 1;       // the workers zeroed the K / V column tails of this head
            }
            mbar_wait(bar_q, ph_q, c.err_flag, 32); ph_q ^= 1;
            fence_after();
            if (leader) trace_evt(p, 1010);
            if (lane == 0) bulk_wait_read0();                                   // the previous tile's dS stores have left sDS
            for (int part = 0; part < nparts; ++part) {
              const int buf = part & 1;
              if (part >= 2) {                                                  // workers drained this buffer (part - 2)
                mbar_wait(bar_free0 + 8 * buf, ph_free[buf], c.err_flag, 34); ph_free[buf] ^= 1;
                fence_after();
              }
              if (leader) trace_evt(p, 1020 + part);
              const uint32_t id_s = part == nparts - 1 ? id_s_tail : id_s_full;
              const uint32_t ts = tmem + T_SD + 128u * buf;
              const uint32_t koff = (uint32_t)(part * KPART * 128) >> 4;      // key rows of this part inside the K / V tiles
              if (leader) {
                const uint64_t bk = dKk + koff, bv = dVk + koff;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) if (ks < nks) mma_bf16(ts, dQk + 2 * ks, bk + 2 * ks, id_s, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) if (ks < nks) mma_bf16(ts + 64, dDOk + 2 * ks, bv + 2 * ks, id_s, ks > 0);
              }
              if (leader) commit(bar_sd0 + 8 * buf);
            }
            for (int part = max(0, nparts - 2); part < nparts; ++part) {       // the last parts: P and dS of the tile complete
              const int buf = part & 1;
              mbar_wait(bar_free0 + 8 * buf, ph_free[buf], c.err_flag, 35); ph_free[buf] ^= 1;
            }
            fence_after();
            for (int a = 0; a < nparts; ++a)                                    // dS_h of this tile -> scratch[b, h, i * 128 .., a * 64 ..]
              if (lane == 0) tma_store_3d(&mDS, smem_u32(sDS) + a * 16384, a * 64, i * 128, b * heads + h);
            if (lane == 0) bulk_commit();
            if (leader) trace_evt(p, 1030);
            // dQ_i = dS_i K (contraction over the keys): 16 keys = 32 B of a K-major dS row, 2048 B of MN-major K. One branch for the
            // whole batch and immediates for the operand steps: the issuing lane runs ~4 instructions per MMA instead of ~25.
            if (leader) {
              const int nk16 = S / 16;
              for (int a = 0; a < nparts; ++a) {
                const uint64_t da = dDSk + (uint32_t)(a * 1024), db = dKmn + (uint32_t)(a * 512);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (4 * a + j < nk16) mma_bf16(tmem + T_DQ, da + 2 * j, db + 128 * j, id_dq, (a | j) != 0);
              }
              for (int t = 0; t < ntiles; ++t) {    // dK_t += dS_i^T Q_i ; dV_t += P_i^T dO_i   (contraction over this tile's 128 queries)
                const uint64_t ak = dDSmn + (uint32_t)(2 * t * 1024), av = dPmn + (uint32_t)(2 * t * 1024);
#pragma unroll
                for (int kq = 0; kq < 8; ++kq) {
                  const uint32_t acc = (i > 0 || kq > 0) ? 1u : 0u;
                  mma_bf16(tmem + T_DK + 64 * t, ak + 128 * kq, dQmn + 128 * kq, id_dkv, acc);
                  mma_bf16(tmem + T_DV + 64 * t, av + 128 * kq, dDOmn + 128 * kq, id_dkv, acc);
                }
              }
            }
            if (leader) commit(bar_fin);
            if (leader) trace_evt(p, 1040);
            // Q / dO (and, after a head's last tile, K / V) are free as soon as these MMAs have retired: the next loads run under
            // the workers' store epilogue
            mbar_wait(bar_fin, ph_fin, c.err_flag, 37); ph_fin ^= 1;
            {
              int nit = it, ni = i + 1;
              if (ni == ntiles) { ni = 0; nit = it + (int)gridDim.x; }
              if (nit < items) issue_loads(nit / heads, nit % heads, ni);
            }
            // TMEM is free again once the workers have READ the accumulators: the next tile's S / dP MMAs run under their staging and
            // global stores (the P / dS tiles are only rewritten by the workers themselves, after their own epilogue)
            mbar_wait(bar_acc, ph_acc, c.err_flag, 36); ph_acc ^= 1;
            if (leader) trace_evt(p, 1050);
          }
        }
      }
      if (lane == 0) bulk_wait0();
    }
  } else {
    const int grp = warp >> 2;                    // 16-column chunk of every 64-key part this warp works on
    const int r = (warp & 3) * 32 + lane;         // row within the tile == TMEM lane
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_kv = 0, ph_fin = 0, ph_sd[2] = {0, 0};
    // bias row segment of a part: this thread's chunk [grp * 16, grp * 16 + 16) of key columns [part * 64, part * 64 + 64)
    uint4 bcur[2], bnxt[2];
    auto load_bias = [&](uint4* dst, int bb, int ii, int part) {
      const int qq = ii * 128 + r;
      const bf16* brow = c.bias + ((long long)bb * S + (qq < S ? qq : 0)) * S + part * KPART;
      const int w = min(KPART, S - part * KPART);
      const int cc = grp * 16;
      if (cc < w) {
        dst[0] = *reinterpret_cast<const uint4*>(brow + cc);
        dst[1] = *reinterpret_cast<const uint4*>(brow + cc + 8);
      }
    };
    const int items = c.B * heads;
    if ((int)blockIdx.x < items) load_bias(bcur, (int)blockIdx.x / heads, 0, 0);
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      {
        const int b = it / heads, h = it - b * heads;
        const HeadCols hc = head_cols(h, hd);
        const int hdp = hc.hdp;
        for (int i = 0; i < ntiles; ++i) {
          const int q = i * 128 + r;
          const bool valid = q < S;
          if (i == 0) {
            // The 64-column boxes also bring the neighbouring heads' columns. Zeroing them in K and V (once per head) is enough:
            // S = Q K^T and dP = dO V^T then ignore the tails of Q and dO, and the tails only reach dQ / dK / dV columns nobody stores.
            mbar_wait(bar_kv, ph_kv, c.err_flag, 41); ph_kv ^= 1;
            if (threadIdx.x == 0) trace_evt(p, 2000);
            if (grp < 2) {
              uint8_t* tile = grp == 0 ? sK : sV;
              for (int row = r; row < S; row += 128) zero_outside(tile, row, hc, hd);
            }
            fence_proxy_async();
            mbar_arrive(bar_tile);
            if (threadIdx.x == 0) trace_evt(p, 2010);
          }
          const long long stat = ((long long)b * heads + h) * S + (valid ? q : 0);
          const float lse2 = p.lse[stat] * LOG2E, dl = p.delta[stat];
          for (int part = 0; part < nparts; ++part) {
            const int buf = part & 1;
            const int k0 = part * KPART, w = min(KPART, S - k0);
            // next part's bias (next tile / head / image after the tile's last part)
            {
              int nb = b, ni = i, np = part + 1;   // (the division by `heads` only when the item changes)
              if (np == nparts) {
                np = 0;
                if (++ni == ntiles) {
                  ni = 0;
                  const int nit = it + (int)gridDim.x;
                  nb = nit < items ? nit / heads : -1;
                }
              }
              if (nb >= 0) load_bias(bnxt, nb, ni, np);
            }
            if (threadIdx.x == 0) trace_evt(p, 2020 + part);
            mbar_wait(bar_sd0 + 8 * buf, ph_sd[buf], c.err_flag, 43); ph_sd[buf] ^= 1;
            fence_after();
            if (threadIdx.x == 0) trace_evt(p, 2030 + part);
            {
              const int cc = grp * 16;
              if (cc < w) {
                const int kc = k0 + cc;
                uint32_t sr[16], dr[16];
                tmem_ld16(trow + T_SD + 128u * buf + cc, sr);
                tmem_ld16(trow + T_SD + 128u * buf + 64 + cc, dr);
                tmem_ld_wait();
                float bf[16], pv[16], ds[16];
                unpack16(bcur[0], bcur[1], bf);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float pj = valid ? ex2_approx(fmaf(__uint_as_float(sr[e]), c.scale_log2, fmaf(bf[e], LOG2E, -lse2))) : 0.f;
                  pv[e] = pj;
                  ds[e] = pj * (__uint_as_float(dr[e]) - dl);   // rows beyond S contribute exact zeros to dK / dV
                }
                store_row16(sP, r, kc, pv);
                store_row16(sDS, r, kc, ds);
              }
            }
            if (threadIdx.x == 0) trace_evt(p, 2040 + part);
            fence_proxy_async();
            fence_before();
            mbar_arrive(bar_free0 + 8 * buf);
            if (threadIdx.x == 0) trace_evt(p, 2050 + part);
            bcur[0] = bnxt[0]; bcur[1] = bnxt[1];
          }
          mbar_wait(bar_fin, ph_fin, c.err_flag, 44); ph_fin ^= 1;
          fence_after();
          if (threadIdx.x == 0) trace_evt(p, 2060);
          // dQ rows of this tile: TMEM -> staging atom 0 of sP (free: the dV MMAs have retired; sDS is still being read by the
          // dS store) -> coalesced rows. After the head's last tile the same for dK_t (atoms t) and dV_t (atoms 2 + t).
          {
            const int quad = warp & 3;
            const int row_begin = quad * 32 + grp * (32 / NGROUPS);    // each warp of the quadrant stores 32 / NGROUPS of its rows
            const long long g_end = (long long)b * S + S;
            for (int c0 = grp * 16; c0 < hdp; c0 += 16 * NGROUPS) {
              uint32_t v[16];
              tmem_ld16(trow + T_DQ + c0, v);
              tmem_ld_wait();
              stage_cols16(sP, r, v, c0, c.scale);
            }
            if (i != ntiles - 1) { fence_before(); mbar_arrive(bar_acc); }
            quad_sync128(quad);
            store_rows16<32 / NGROUPS>(sP, p.dq + (long long)h * hd, p.ld_dq, row_begin, (long long)b * S + i * 128, g_end, hc, hd, lane);
            if (i == ntiles - 1) {
              quad_sync128(quad);                                // atom 0 is reused
              for (int t = 0; t < ntiles; ++t) {                 // TMEM lanes are key rows of M-tile t
                for (int c0 = grp * 16; c0 < hdp; c0 += 16 * NGROUPS) {
                  uint32_t kk[16], vv[16];
                  tmem_ld16(trow + T_DK + 64 * t + c0, kk);
                  tmem_ld16(trow + T_DV + 64 * t + c0, vv);
                  tmem_ld_wait();
                  stage_cols16(sP + t * 16384, r, kk, c0, c.scale);
                  stage_cols16(sP + (2 + t) * 16384, r, vv, c0, 1.0f);
                }
              }
              fence_before();
              mbar_arrive(bar_acc);
              quad_sync128(quad);
              for (int t = 0; t < ntiles; ++t) {
                store_rows16<32 / NGROUPS>(sP + t * 16384, p.dk + (long long)h * hd, p.ld_dk, row_begin, (long long)b * S + t * 128, g_end, hc, hd, lane);
                store_rows16<32 / NGROUPS>(sP + (2 + t) * 16384, p.dv + (long long)h * hd, p.ld_dv, row_begin, (long long)b * S + t * 128, g_end, hc, hd, lane);
              }
            }
            quad_sync128(quad);   // the staging rows are rewritten (P of the next tile) by the partner warps
          }
          if (threadIdx.x == 0) trace_evt(p, 2070);
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == CTRL_WARP) {
    fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// dbias[b, q, k] = sum_h dS[b, h, q, k]  (bf16 in, fp32 sum in head order, bf16 out); 8 elements (16 bytes) per thread
__global__ void dbias_reduce_kernel(const bf16* __restrict__ ds, bf16* __restrict__ dbias, int heads, long long ss8 /* S * S / 8 */,
                                    long long total8 /* B * S * S / 8 */) {
  pdl_wait(); pdl_launch_dependents();   // see common.cuh: launched with programmatic stream serialization (after the backward kernel)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const long long b = i / ss8, r = i - b * ss8;
  const uint4* src = reinterpret_cast<const uint4*>(ds) + b * heads * ss8 + r;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int h0 = 0; h0 < heads; h0 += 4) {      // 4 independent 16-byte loads in flight
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = h0 + j < heads ? __ldcs(src + (long long)(h0 + j) * ss8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t;
      t = unpack_bf16x2(u[j].x); acc[0] += t.x; acc[1] += t.y;
      t = unpack_bf16x2(u[j].y); acc[2] += t.x; acc[3] += t.y;
      t = unpack_bf16x2(u[j].z); acc[4] += t.x; acc[5] += t.y;
      t = unpack_bf16x2(u[j].w); acc[6] += t.x; acc[7] += t.y;
    }
  }
  reinterpret_cast<uint4*>(dbias)[i] = pack8f(acc);
}

// 3-D bf16 map {S cols, S rows, B * heads} over the per-head dS scratch, box {64, 128, 1}, 128B swizzle
int make_map_ds(CUtensorMap* map, void* base, uint64_t S, uint64_t bh) {
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

constexpr size_t FWD_SMEM = 2 * KV_BYTES + QT_BYTES + 2 * P_BYTES + 2 * FWD_GROUPS * 128 * 4 + 256 + 1024;
constexpr size_t BWD_SMEM = 2 * KV_BYTES + 2 * QT_BYTES + 2 * P_BYTES + 256 + 1024;

int fill_common(Common& c, const void* bias, int B, int S, int heads, int hd) {
  c.B = B; c.S = S; c.heads = heads; c.hd = hd;
  c.scale = 1.0f / sqrtf((float)hd);
  c.scale_log2 = c.scale * LOG2E;
  c.bias = reinterpret_cast<const bf16*>(bias);
  c.err_flag = g_calm_err_flag;
  return CALM_OK;
}

}  // namespace

bool calm_attention_tc_eligible(int B, int S, int heads, int hd, const int64_t* lds, int nlds, const void* const* ptrs, int nptrs) {
  if (B <= 0 || heads <= 0 || S < 16 || S > 256 || (S & 15) || hd < 4 || hd > 64 || (hd & 3)) return false;
  if ((hd & 7) && hd > 60) return false;                // a head shifted by 4 columns must still fit the 64-column box
  for (int i = 0; i < nlds; ++i)
    if (lds[i] % 8) return false;                       // TMA: row pitch multiple of 16 bytes
  for (int i = 0; i < nptrs; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return false;
  return true;
}

int calm_attention_fwd_tc(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
                          int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream) {
  FwdParams p;
  fill_common(p.c, bias, B, S, heads, hd);
  p.o = reinterpret_cast<bf16*>(o); p.ld_o = ld_o; p.lse = lse;
  CUtensorMap mQ, mK, mV, mB;
  int rc;
  const uint64_t rows = (uint64_t)B * S, cols = (uint64_t)heads * hd;
  if ((rc = tc::make_map_2d(&mQ, q, cols, rows, ld_q, 128))) return rc;
  if ((rc = tc::make_map_2d(&mK, k, cols, rows, ld_k, S))) return rc;
  if ((rc = tc::make_map_2d(&mV, v, cols, rows, ld_v, S))) return rc;
  if ((rc = tc::make_map_2d(&mB, bias, (uint64_t)S, rows, S, 128))) return rc;   // bias viewed as (B*S rows, S columns)
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) { calm_set_error("calm_attention_fwd(tcgen05): smem: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  const int items = B * heads;
  const int grid = items < calm_num_sms() ? items : calm_num_sms();
  if (S <= (int)F2_SCOLS) {      // two S accumulators + O fit TMEM: the pipelined kernel
    static CalmDeviceOnce configured2;
    if (configured2.pending()) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
      if (e != cudaSuccess) { calm_set_error("calm_attention_fwd(tcgen05, pipelined): smem: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
      configured2.done();
    }
    attn_fwd_pipe_kernel<<<grid, F2_THREADS, FWD_SMEM, stream>>>(mQ, mK, mV, mB, p);
    CALM_CHECK_LAUNCH("calm_attention_fwd(tcgen05, pipelined)");
    return CALM_OK;
  }
  attn_fwd_tc_kernel<<<grid, FWD_THREADS, FWD_SMEM, stream>>>(mQ, mK, mV, mB, p);
  CALM_CHECK_LAUNCH("calm_attention_fwd(tcgen05)");
  return CALM_OK;
}

// bring-up only (-DCALM_BRINGUP): CTA 0 of the backward kernel writes [count, (event id, globaltimer ns)...] into this buffer
static unsigned long long* g_trace_buf = nullptr;
static int g_trace_cap = 0;
#ifdef CALM_BRINGUP
extern "C" void calm_debug_set_trace_buffer(void* device_u64, int32_t capacity_events) {
  g_trace_buf = reinterpret_cast<unsigned long long*>(device_u64);
  g_trace_cap = capacity_events;
}
#endif
CalmTrace calm_trace_target() { return CalmTrace{g_trace_buf, g_trace_cap}; }

size_t calm_attention_bwd_tc_scratch_bytes(int B, int S, int heads) { return (size_t)2 * B * heads * S * S; }

int calm_attention_bwd_tc(const void* q, const void* k, const void* v, const void* bias, const void* d_o, const float* lse,
                          const float* delta, void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q, int64_t ld_k,
                          int64_t ld_v, int64_t ld_do, int64_t ld_dq, int64_t ld_dk, int64_t ld_dv, int B, int S, int heads, int hd,
                          cudaStream_t stream) {
  BwdParams p;
  fill_common(p.c, bias, B, S, heads, hd);
  p.lse = lse; p.delta = delta;
  p.dq = reinterpret_cast<bf16*>(dq); p.dk = reinterpret_cast<bf16*>(dk); p.dv = reinterpret_cast<bf16*>(dv);
  p.ld_dq = ld_dq; p.ld_dk = ld_dk; p.ld_dv = ld_dv;
  p.trace = calm_trace_target();
  CUtensorMap mQ, mK, mV, mDO, mDS;
  int rc;
  const uint64_t rows = (uint64_t)B * S, cols = (uint64_t)heads * hd;
  if ((rc = tc::make_map_2d(&mQ, q, cols, rows, ld_q, 128))) return rc;
  if ((rc = tc::make_map_2d(&mK, k, cols, rows, ld_k, S))) return rc;
  if ((rc = tc::make_map_2d(&mV, v, cols, rows, ld_v, S))) return rc;
  if ((rc = tc::make_map_2d(&mDO, d_o, cols, rows, ld_do, 128))) return rc;
  if ((rc = make_map_ds(&mDS, ds_scratch, (uint64_t)S, (uint64_t)B * heads))) return rc;
  static CalmDeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
    if (e != cudaSuccess) { calm_set_error("calm_attention_bwd(tcgen05): smem: %s", cudaGetErrorString(e)); return CALM_ERR_CUDA; }
    configured.done();
  }
  const int bwd_items = B * heads;
  const int grid = bwd_items < calm_num_sms() ? bwd_items : calm_num_sms();
  attn_bwd_tc_kernel<<<grid, NTHREADS, BWD_SMEM, stream>>>(mQ, mK, mV, mDO, mDS, p);
  CALM_CHECK_LAUNCH("calm_attention_bwd(tcgen05)");
  return calm_attention_dbias_reduce(ds_scratch, dbias, B, S, heads, stream);
}

int calm_attention_dbias_reduce(const void* ds_scratch, void* dbias, int B, int S, int heads, cudaStream_t stream) {
  const long long ss8 = (long long)S * S / 8, total8 = ss8 * B;
  CALM_LAUNCH((dbias_reduce_kernel), (unsigned)((total8 + 255) / 256), 256, 0, stream, reinterpret_cast<const bf16*>(ds_scratch),
              reinterpret_cast<bf16*>(dbias), heads, ss8, total8);
  CALM_CHECK_LAUNCH("calm_attention_bwd(dbias reduce)");
  return CALM_OK;
}
