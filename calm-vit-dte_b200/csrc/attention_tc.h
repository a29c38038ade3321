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

// Bring-up time stamps (-DCALM_BRINGUP builds only): CTA 0 appends (event id, globaltimer ns) pairs to the device buffer handed to
// calm_debug_set_trace_buffer (tools/attn_trace.py reads them back); production builds compile calm_trace() to nothing.
struct CalmTrace { unsigned long long* buf; int cap; };
CalmTrace calm_trace_target();
#ifdef __CUDACC__
__device__ __forceinline__ void calm_trace(const CalmTrace& t, int id) {
#ifdef CALM_BRINGUP
  if (t.buf != nullptr && blockIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    const unsigned long long i = atomicAdd(t.buf, 1ULL);
    if ((int)i < t.cap) { t.buf[1 + 2 * i] = (unsigned long long)id; t.buf[2 + 2 * i] = now; }
  }
#else
  (void)t; (void)id;
#endif
}
#endif
