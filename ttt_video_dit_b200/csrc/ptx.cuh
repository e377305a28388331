// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (TMEM alloc / mma / ld / st /
// commit / fences) and the shared-memory matrix + instruction descriptors used by every kernel in this directory.
// Nothing here comes from CUTLASS/ThunderKittens; bit layouts follow the PTX ISA "tcgen05 matrix descriptor" and
// "instruction descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp field positions).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// Phase timing (debug builds only: -DTTT_PHASE_TIMING -> lib/libttt_b200_dbg.so).  TICK(i) adds the cycles since the
// previous TICK to slot i for two observer threads of block 0; the kernel epilogue dumps the slots to tb::g_timing_buf.
#ifdef TTT_PHASE_TIMING
#define TICK_DECL(NSLOT, OBS_B)                                                                         \
  const int tick_obs = (blockIdx.x == 0 && threadIdx.x == 0) ? 0 : ((blockIdx.x == 0 && threadIdx.x == (OBS_B)) ? 1 : -1); \
  unsigned tick_acc[NSLOT];                                                                             \
  _Pragma("unroll") for (int i_ = 0; i_ < (NSLOT); ++i_) tick_acc[i_] = 0;                             \
  unsigned tick_last = clock();
#define TICK(i)                          \
  do {                                   \
    const unsigned now_ = clock();       \
    tick_acc[i] += now_ - tick_last;     \
    tick_last = now_;                    \
  } while (0)
#define TICK_DUMP(NSLOT, buf)                                                                              \
  do {                                                                                                     \
    if (tick_obs >= 0 && (buf) != nullptr)                                                                 \
      _Pragma("unroll") for (int i_ = 0; i_ < (NSLOT); ++i_)(buf)[tick_obs * 64 + i_] = tick_acc[i_];   \
  } while (0)
#else
#define TICK_DECL(NSLOT, OBS_B)
#define TICK(i)
#define TICK_DUMP(NSLOT, buf)
#endif

namespace tb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ----------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as a launch failure (trap), never as a hung GPU.
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar_addr, uint32_t parity) {
  printf("ttt_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
         (int)threadIdx.x, bar_addr, parity);
  __trap();
}
// slow path out of line: the kernels are instruction-cache sensitive and have dozens of wait sites
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity) {
  // try_wait itself may sleep for a hardware-defined interval, so the clock is checked on every poll
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((clock64() - t0) > 4000000000LL) mbar_timeout_trap(bar_addr, parity);
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(smem_u32(bar), parity);
}

// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tiled load: box lands at dst (swizzle mode comes from the tensor map), completes tx bytes on bar
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(smem_u32(src))
               : "memory");
}
// 1-D bulk copies (no tensor map): contiguous bytes, size multiple of 16
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all previously issued tcgen05.mma of this thread arrive (once) on bar when complete
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc];   kind::f16 (bf16/fp16 operands, fp32 accumulate)
// One elected lane of a fully converged warp.  Single-thread tcgen05 / TMA issue must sit under a branch ptxas can prove
// to be "one lane of a warp-uniform region": `if (warp_uniform == W && elect_one())`.  Under a plain `if (tid == 0)` every
// uniform-datapath instruction (UTCHMMA, UTCBAR, UTMALDG, UBLKCP) is wrapped in its own 5-instruction election loop and
// the descriptors are re-materialised per instruction (scripts/sass_experiments/elect_issue.cu: 12 issue slots per MMA
// instead of 1).
#define TB_HAS_ELECT_ONE 1
#ifndef TB_NO_ELECT
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp index as a value the compiler knows to be warp-uniform
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
#else  // A/B build (lib/libttt_b200_noelect.so): the round-1 `if (tid == 0)` issue pattern, for scripts/r02_gpu_check.sh
__device__ __forceinline__ bool elect_one() { return (threadIdx.x & 31) == 0; }
__device__ __forceinline__ int uniform_warp_id() { return (int)(threadIdx.x >> 5); }
#endif

__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM ("TS" form): A[128 lanes][16 k] sits in 8 consecutive 32-bit TMEM columns of the lanes (two bf16 of
// consecutive k per column, low half = even k), always K-major; B through a shared-memory descriptor as above.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// tcgen05.ld 32x32b: thread i of the warp reads TMEM lane (lane_base+i), N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// one 32-bit cell per lane: scalar exchange between the two warps that share a TMEM lane quadrant
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
  return r;
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ----------------------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset>>4 bits [46,48) version (1)   bits [61,64) layout (2 = SWIZZLE_128B)
// Tile storage convention in this repo ("SW128 row tile"): rows of 64 bf16 (128 B); row r lives at r*128 B; the
// 16-byte chunk c of row r is stored at chunk position (c ^ (r & 7)); tile base is 1024-B aligned.
//   K-major view  (rows = M/N index, 64 contiguous = K):  SBO = 1024 (8 rows), K-step of 16 elements = +32 B.
//   MN-major view (rows = K index, 64 contiguous = M/N):  SBO = 1024 (8 k-rows), LBO = byte distance between
//                  64-wide M/N blocks, K-step of 16 rows = +2048 B.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// advance the start-address field by `bytes` (multiple of 16)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// Instruction descriptor, kind::f16: D=f32, A=B=bf16.
//   [4,6) c_format (1=f32) [7,10) a_format (1=bf16) [10,13) b_format [13] negA [14] negB [15] a_major (1 = MN)
//   [16] b_major [17,23) N>>3 [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn, bool b_mn, bool neg_a = false,
                                                       bool a_f16 = false, bool b_f16 = false) {
  return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((neg_a ? 1u : 0u) << 13) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------- packed fp32 pairs
// FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add .f32x2): two fp32 lanes per register pair and per instruction, each lane rounded
// exactly like the scalar instruction.  The fma pipe accepts one warp instruction every 2 cycles per scheduler
// (B300_MICROARCH.md "Pipe rates"), so the scalar form caps an SM at 64 FMA/clk; the element-wise phases of the scan
// kernels are bound by that pipe, not by issue slots or latency (round 2: doubling the warps per CTA changed nothing).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pk2(float v) { return pk2(v, v); }
__device__ __forceinline__ void up2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; up2(v, a, b); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; up2(v, a, b); return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {  // two fp32 bit patterns (e.g. straight out of tcgen05.ld)
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void up2u(f32x2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
// a packed bf16 pair -> packed fp32 pair (low half first)
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t v) { return pk2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u)); }
__device__ __forceinline__ uint32_t pack_bf16(f32x2 v) {
  float a, b;
  up2(v, a, b);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// ----------------------------------------------------------------------------------------------- SW128 helpers
// byte offset of 16-B chunk c (0..7) of row r in a SW128 row tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
// fp16 operand tiles (forward-type GEMMs of the TTT-MLP scans: 11-bit mantissa instead of bf16's 8 -- the first mini-batch
// of a sequence has a LayerNorm std of ~3e-3 and amplifies operand rounding by 1/std, DESIGN.md "operand format")
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f16_lo(uint32_t v) {
  float r;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tcvt.f32.f16 %0, lo;\n\t}" : "=f"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float f16_hi(uint32_t v) {
  float r;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tcvt.f32.f16 %0, hi;\n\t}" : "=f"(r) : "r"(v));
  return r;
}
// a bf16 pair -> the same two values as an fp16 pair (exact for |x| in [6.1e-5, 65504]; smaller magnitudes keep 2^-24 steps)
__device__ __forceinline__ uint32_t bf16x2_to_f16x2(uint32_t v) { return pack_f16(bf16_lo(v), bf16_hi(v)); }
// operand-format switch of the TTT-MLP forward-type kernels
template <bool kF16> __device__ __forceinline__ uint32_t pack_op(float lo, float hi) { return kF16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
template <bool kF16> __device__ __forceinline__ uint32_t pack_op(f32x2 v) { float a, b; up2(v, a, b); return pack_op<kF16>(a, b); }
template <bool kF16> __device__ __forceinline__ f32x2 op_to_f32x2(uint32_t v) { return kF16 ? pk2(f16_lo(v), f16_hi(v)) : pk2(bf16_lo(v), bf16_hi(v)); }
template <bool kF16> __device__ __forceinline__ float op_lo(uint32_t v) { return kF16 ? f16_lo(v) : bf16_lo(v); }
template <bool kF16> __device__ __forceinline__ float op_hi(uint32_t v) { return kF16 ? f16_hi(v) : bf16_hi(v); }

__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t saddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr) : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace tb
