// extern "C" surface of libttt_b200_selftest.so (declared in include/ttt_b200_debug.h): tcgen05 descriptor self-test
// and the interference spin kernel.  Kept out of the production library (libttt_b200.so exports only include/ttt_b200.h).
#include "../../include/ttt_b200_debug.h"
#include "capi_util.h"

namespace tb { thread_local const char* g_where = ""; unsigned* g_timing_buf = nullptr; }

extern "C" {

const char* ttt_b200_debug_last_error(void) { return g_err; }

int ttt_b200_debug_umma(int mode, const void* A, const void* Bm, float* D, int N, int K, void* stream) {
  TB_BIND_DEVICE(A);
  return cuda_ret(tb::launch_umma_selftest(mode, A, Bm, D, N, K, (cudaStream_t)stream), "ttt_b200_debug_umma");
}

// Occupancy / interference experiments: `blocks` CTAs that spin for `cycles` SM cycles with a tiny code footprint and no
// memory traffic.  mode 0: dependent FMA chains (ALU busy), mode 1: nanosleep (SM occupied but idle), mode 2 / 3:
// streaming stores / loads over sink[0 .. sink_floats).  smem_bytes of dynamic shared memory pins one CTA per SM when
// set close to the maximum.
int ttt_b200_debug_spin(int blocks, int threads, long long cycles, int mode, int smem_bytes, float* sink,
                        long long sink_floats, void* stream) {
  if (!sink) return fail(-1, "ttt_b200_debug_spin: null pointer argument");
  TB_BIND_DEVICE(sink);
  return cuda_ret(tb::launch_debug_spin(blocks, threads, cycles, mode, smem_bytes, sink, sink_floats, (cudaStream_t)stream),
                  "ttt_b200_debug_spin");
}

// Distributed-shared-memory microbenchmark (csrc/dsmem_probe.cu): out[0] = cycles of one one-way transfer of `bytes` between the
// CTAs of a 2-CTA cluster including the completion signal.  mode 0 = bulk copy, 1 = st.shared::cluster, 2 = 4 bulk copies.
int ttt_b200_debug_dsmem(int mode, int bytes, int iters, float* out, void* stream) {
  if (!out) return fail(-1, "ttt_b200_debug_dsmem: null pointer argument");
  TB_BIND_DEVICE(out);
  return cuda_ret(tb::launch_dsmem_probe(mode, bytes, iters, out, (cudaStream_t)stream), "ttt_b200_debug_dsmem");
}

}  // extern "C"
