/* libttt_b200_selftest.so -- development probes of the B200 TTT kernels (NOT part of the production ABI in ttt_b200.h).
 * The production library exports none of these.  The phase-timing build (lib/libttt_b200_dbg.so, -DTTT_PHASE_TIMING)
 * additionally exports `int ttt_b200_debug_set_timing_buffer(void* dev_buf_512_bytes)`. */
#ifndef TTT_B200_DEBUG_H
#define TTT_B200_DEBUG_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* ttt_b200_debug_last_error(void);

/* Debug/self-test: D[128][N] = A[128][K] . Bm[K][N] through one tcgen05 CTA (see csrc/umma_selftest.cu). */
int ttt_b200_debug_umma(int mode, const void* A_bf16, const void* B_bf16, float* D, int N, int K, void* stream);

/* Interference experiments (scripts/gpu_probe.py): `blocks` CTAs spinning for `cycles` SM cycles; mode 0 = FMA chains,
 * mode 1 = nanosleep, mode 2 / 3 = streaming stores / loads over sink[0 .. sink_floats); smem_bytes of dynamic shared
 * memory (to pin one CTA per SM). */
int ttt_b200_debug_spin(int blocks, int threads, long long cycles, int mode, int smem_bytes, float* sink,
                        long long sink_floats, void* stream);

/* Distributed-shared-memory microbenchmark (csrc/dsmem_probe.cu): out[0] (device float) = cycles of ONE one-way transfer of
 * `bytes` between the two CTAs of a cluster including the completion signal (ping-pong / 2).  mode 0 = one
 * cp.async.bulk.shared::cluster copy, 1 = st.shared::cluster.v4 by 256 threads, 2 = four bulk copies back to back. */
int ttt_b200_debug_dsmem(int mode, int bytes, int iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
