/* twotower_debug.h -- measurement and test hooks of libtwotower.so.  NOT part of the product C-ABI
 * (include/twotower.h): nothing a drop-in replacement of the TFRS path needs is declared here.  Used by
 * bench.py (per-kernel event timing, in-graph timeline), tools/ and the descriptor-layout tests. */
#ifndef TWOTOWER_DEBUG_H_
#define TWOTOWER_DEBUG_H_

#include "twotower.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-kernel timing for bench.py: while enabled, every kernel the library launches outside a
 * CUDA-graph capture is bracketed by a cudaEvent pair on its own stream.  tt_profile_collect
 * synchronises and writes "kernel_name launches total_ms\n" lines into host_buf; returns the
 * buffer size needed.  tt_profile_enable(0/1) also clears the records. */
int tt_profile_enable(int32_t on);
int64_t tt_profile_collect(char* host_buf, int64_t buf_len);

/* Test hook: D[M,N] (fp32) = A * B with bf16 operands in either storage order
 * (a_mn = 0: A is [M,K]; 1: A is [K,M].  b_mn = 0: B is [N,K]; 1: B is [K,N]). */
int tt_debug_gemm_bf16(const void* A, int32_t a_mn, const void* B, int32_t b_mn, int64_t M,
                       int64_t N, int64_t K, float* out, void* stream);

/* Tuning hook: device buffer of 3 * (16 * 64 + 16) + 3 * 4 * 256 int64 that receives clock64() stamps of
 * the software pipeline of CTA (0,0) of the bf16 loss forward / dQ / dC kernels and per-CTA
 * {entry, setup done, exit, smid} globaltimer records (grids up to 256 CTAs); NULL = off. */
int tt_debug_trace_buffer(long long* device_buf);
/* Same for the fused tower kernels: 2 * 16 * 256 int64, per-CTA phase stamps (globaltimer ns). */
int tt_debug_tower_trace(long long* device_buf);
/* In-stream timeline of one training step: 32 int64 (16 kernel ids), {earliest CTA entry, latest CTA entry/exit} in globaltimer ns
 * for kernel ids 0 tower fwd, 1 loss fwd, 2 dQ, 3 dC, 4 tower bwd, 5 optimizer step, 6 sparse prepare.  The pointer
 * is read at run time (works on captured graphs).  Caller presets even slots to INT64_MAX, odd slots to 0. */
int tt_debug_timeline(long long* device_buf);

/* Test hook for the bf16 top-k: 1 = threshold scan where the shape allows it (default; also TT_TOPK_SCAN unset),
 * 0 = always the list-keeping kernel, 2 = threshold scan with 64-entry survivor buffers, so every row overflows and
 * the device-side fallback to the list-keeping kernel runs.  Returns the previous mode. */
int tt_debug_topk_scan_mode(int32_t mode);
/* Tuning hook: device buffer of 3 * 48 * 4 int64 receiving clock64() stamps of CTA (0,0) of the second-phase scan kernel
 * (roles 0 / 1: first thread of the selection warpgroup of query tile 0 / 1 -- wait begins, accumulator ready, accumulator
 * read and released, tile done; role 2: the MMA thread -- wait begins, candidate tile landed, first accumulator buffer
 * free, both units issued); NULL = off.  tools/trace_topk_scan.py */
int tt_debug_topk_scan_trace(long long* device_buf);

#ifdef __cplusplus
}
#endif

#endif /* TWOTOWER_DEBUG_H_ */
