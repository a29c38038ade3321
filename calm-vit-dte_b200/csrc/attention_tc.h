// Internal interface between attention.cu (C-ABI entry points, legacy mma.sync kernels for long sequences / wide heads)
// and attention_sm100.cu (tcgen05 kernels for S <= 256, head_dim <= 64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

bool calm_attention_tc_eligible(int B, int S, int heads, int hd, const int64_t* lds, int nlds, const void* const* ptrs, int nptrs);
int calm_attention_fwd_tc(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
                          int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream);
int calm_attention_bwd_tc(const void* q, const void* k, const void* v, const void* bias, const void* d_o, const float* lse,
                          const float* delta, void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q, int64_t ld_k,
                          int64_t ld_v, int64_t ld_do, int64_t ld_dq, int64_t ld_dk, int64_t ld_dv, int B, int S, int heads, int hd,
                          cudaStream_t stream);
size_t calm_attention_bwd_tc_scratch_bytes(int B, int S, int heads);
// dbias[b] = sum over heads of the per-head dS scratch (B, heads, S, S) bf16 -> (B, S, S) bf16 (fp32 sum in head order); S * S % 8 == 0
int calm_attention_dbias_reduce(const void* ds_scratch, void* dbias, int B, int S, int heads, cudaStream_t stream);
// attention_long_sm100.cu: tcgen05 forward for long rows / wide heads (key axis in 128-key chunks, two passes; S % 16 == 0, hd <= 128)
bool calm_attention_long_eligible(int B, int S, int heads, int hd, const int64_t* lds, int nlds, const void* const* ptrs, int nptrs);
int calm_attention_fwd_long(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
                            int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream);
int calm_attention_bwd_long(const void* q, const void* k, const void* v, const void* bias, const void* d_o, const float* lse,
                            const float* delta, void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q, int64_t ld_k,
                            int64_t ld_v, int64_t ld_do, int64_t ld_dq, int64_t ld_dk, int64_t ld_dv, int B, int S, int heads, int hd,
                            cudaStream_t stream);
// attention_small.cu: forward with one warp per (image, head) item, scores in registers (mma.sync) — S <= 96, S % 16 == 0, head_dim <= 32
bool calm_attention_small_eligible(int B, int S, int heads, int hd, const int64_t* lds, int nlds, const void* const* ptrs, int nptrs);
int calm_attention_fwd_small(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse, int64_t ld_q, int64_t ld_k,
                             int64_t ld_v, int64_t ld_o, int B, int S, int heads, int hd, cudaStream_t stream);

// Bring-up time stamps (-DCALM_BRINGUP builds only). Up to four threads of CTA 0 (region 0: controller / S-MMA lane, 1: worker thread 0,
// 2: PV-MMA lane, 3: loader lane) write (event id, globaltimer ns) pairs into their own region of `cap` pairs of the device buffer
// handed to calm_debug_set_trace_buffer, with the write position in a register: no atomics, no loads, two fire-and-forget stores per event (an atomic
// counter cost ~0.4 us per event, as much as the phases being measured). tools/attn_trace.py reads them back (a pair with id 0
// ends a region). Production builds compile all of it to nothing.
struct CalmTrace { unsigned long long* buf; int cap; };       // cap = pairs per region
CalmTrace calm_trace_target();
#ifdef __CUDACC__
struct CalmTraceCursor {
#ifdef CALM_BRINGUP
  unsigned long long* at; unsigned long long* end;
#endif
};
__device__ __forceinline__ CalmTraceCursor calm_trace_begin(const CalmTrace& t, int half) {
  CalmTraceCursor c;
#ifdef CALM_BRINGUP
  c.at = c.end = nullptr;
  if (t.buf != nullptr && blockIdx.x == 0) { c.at = t.buf + 2 * (size_t)half * t.cap; c.end = c.at + 2 * (size_t)t.cap; }
#else
  (void)t; (void)half;
#endif
  return c;
}
__device__ __forceinline__ void calm_trace(CalmTraceCursor& c, int id) {
#ifdef CALM_BRINGUP
  if (c.at != c.end) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    c.at[0] = (unsigned long long)id; c.at[1] = now;
    c.at += 2;
  }
#else
  (void)c; (void)id;
#endif
}
#endif
