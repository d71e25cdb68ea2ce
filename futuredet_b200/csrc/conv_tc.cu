// tcgen05 / TMEM gather -> implicit-GEMM convolution for sm_100a (FD_PREC_BF16X3, FD_PREC_BF16).
//
//   out[o, :] = act( (sum_k in[nbr(o,k), :] @ W[k]) * scale + shift (+ residual[o, :]) )
//
// GEMM view: M = output rows (active sites / pixels), N = Cout, K = (kernel offsets x Cin) flattened into
// 64-element "K stages".  The kernel is L2->SM bandwidth bound (measured: same time with 1/3 of the MMAs), so
// the schedule is organised around bytes moved:
//   * a work unit is a SUPER-TILE of T consecutive 128-row M tiles x one NT-column N tile; for every K stage the
//     weight tile B(ks) is fetched ONCE (one cp.async.bulk of a pre-swizzled block) and reused by the T M tiles,
//     whose fp32 accumulators all live in TMEM (2 buffers x T x NT columns <= 512);
//   * activations travel between layers as FD_FMT_SPLIT_BF16 rows (bf16 hi | bf16 lo), so the A operand is
//     gathered with plain 16-byte cp.async copies (zero-filled where the rulebook has no neighbour) straight into
//     the UMMA canonical K-major SWIZZLE_128B layout -- no register staging, no conversion work, and the
//     hardware arrives on the stage's mbarrier when the bytes land (cp.async.mbarrier.arrive.noinc);
//   * kernel offsets that are empty for the whole super-tile are skipped (per-tile masks from the rulebook).
//
//   * dense 3x3 stride-1 convolutions (neck, head) run as "tall stages" (TALL instantiation): an M tile is an 8 x 16-pixel
//     patch, an A stage the TMA box {64 ch, 8 px, 18 lines} of one horizontal tap; the three vertical taps read the same
//     bytes through descriptors 0 / 1024 / 2048 bytes into the stage (a patch line = 8 rows = one swizzle atom), so a
//     pixel crosses L2 -> shared memory 3.4 times instead of 9.
//
// Warp roles (416 threads):
//   warps 0-7   producers, split into 4 groups that fill different A stages concurrently (A ring of 4-5 x 32 KB);
//               the group that opens a K stage also issues the B bulk copy (B ring of 2 slots);
//   warp  8     MMA issuer: one lane issues tcgen05.mma (M=128, N=NT, K=16, kind::f16, bf16 x bf16 -> fp32):
//               D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  (3-term split, ~2^-16 relative error: fp32-class results
//               on the bf16 tensor pipe); tcgen05.commit releases A / B slots and publishes the accumulators;
//   warps 9-12  epilogue: tcgen05.ld, fused BN(eval)/bias, residual, ReLU, split once into bf16 hi/lo, store.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "conv_common.cuh"

namespace fd {

constexpr int TC_BM = 128;
// K elements (bf16) per pipeline stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B; twice the
// stages in flight in the same shared memory).  Both are parity-green; measured on the 4-scene forward, 64 is faster
// (14.6 ms vs 18.5 ms): the per-stage hand-shake cost outweighs the deeper ring.
#define FD_TC_BK 64
#ifndef FD_TC_SB2_MIN_NT
#define FD_TC_SB2_MIN_NT 128     // N tiles at least this wide use a 2-slot weight ring (the rest of shared memory is A ring)
#endif
constexpr int TC_BK = FD_TC_BK;
constexpr int TC_ROWB = TC_BK * 2;        // bytes of one stage row (one swizzle atom row)
constexpr int TC_CHUNKS = TC_ROWB / 16;   // 16-byte chunks per row
constexpr int TC_SWZ_SHIFT = TC_BK == 64 ? 0 : 1;   // chunk ^= (row >> shift) & (CHUNKS - 1): SWIZZLE_128B / SWIZZLE_64B
constexpr int TC_MAXK = 32;               // max kernel offsets (27 for 3x3x3)
constexpr int TC_DENSE_MAXK = 9;          // dense 2-D mode: up to 3x3 taps (row indices cached in shared memory)
#ifndef FD_TC_SB64
#define FD_TC_SB64 4                 // weight-ring slots of the 64-column tiles
#endif
#ifndef FD_TC_GROUPS
#define FD_TC_GROUPS 4
#endif
constexpr int TC_GROUPS = FD_TC_GROUPS;   // producer groups (each fills whole A stages on its own)
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_PRODUCERS = TC_PRODUCER_WARPS * 32;
constexpr int TC_MMA_WARP = TC_PRODUCER_WARPS;         // warp 8
constexpr int TC_B_WARP = TC_PRODUCER_WARPS + 5;       // warp 13 (warps 9-12: epilogue, TMEM lane quarters 1,2,3,0)
constexpr int TC_THREADS = TC_PRODUCERS + 32 + 128 + 32;
constexpr int TC_SS_MAX = 512;                         // folded BN scale / shift cached in shared memory up to this Cout
constexpr int TC_A_PLANE = TC_BM * TC_ROWB;   // bytes of one A plane (hi or lo) per stage
constexpr int TC_EPI_WARP_BYTES = 2048;       // per epilogue warp: 32 rows x 64 bytes of staging for coalesced row I/O
constexpr int TC_EPI_BYTES = 4 * TC_EPI_WARP_BYTES;
// dense 3x3 stride-1 "tall" stages (tma == 3): an A stage is an 8-pixel-wide, 16 + 2-line patch of one 64-channel chunk
// at one horizontal tap; the three vertical taps are the same shared-memory bytes read 8 rows (= 1024 bytes, one swizzle
// atom) further down, so a pixel is loaded 3 x 18/16 times instead of 9 times
constexpr int TC_TALL_BW = 8, TC_TALL_BH = 16;
constexpr int TC_TALL_PLANE = (TC_TALL_BH + 2) * TC_TALL_BW * TC_ROWB;   // 18 KB per plane
constexpr int TC_TALL_BYTES = 2 * TC_TALL_PLANE;                          // hi + lo
constexpr int TC_TALL_SLOTS = 4;                                          // two tiles of the current tap column + two of the next
constexpr int TC_TALL_T = 2;


// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must never hang the GPU (a hung box is a lost box).  After ~1 s of spinning the
// waiter reports which barrier starved, raises g_tc_abort and gives up; every other waiter then falls through
// too, the kernel drains (its output is garbage) and conv_forward_tc() turns the flag into an error.
__device__ int g_tc_abort = 0;
__device__ long long g_tc_timeout = 2000000000LL;      // clock cycles; raised for runs under compute-sanitizer (10-100x slower)
// FD_TC_DEBUG & 32: block 0 records clock64() timestamps of its pipeline events (perf triage only)
constexpr int TC_TRACE_N = 8192;
__device__ long long g_tc_trace[4][TC_TRACE_N];
#define TC_TRACE(role, i, v) do { if ((t.dbg & 32) && blockIdx.x == 0 && (i) < TC_TRACE_N) g_tc_trace[role][i] = (v); } while (0)
// `backoff_ns` > 0: the waiter sleeps between polls.  Twelve of the thirteen warps of a CTA spend most of their
// life waiting; polling in a tight loop they would steal the issue slots of the one warp that feeds the tensor core
// (measured: 2.3 M + 3.6 M spin iterations per launch and an MMA warp running at a third of its speed).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0, uint32_t backoff_ns = 0,
                                          uint32_t hint_ns = 0x989680u) {
  uint32_t done;
  long long t0 = 0;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");   // suspend-time hint (ns)
    if (!done) {
      if (backoff_ns) __nanosleep(backoff_ns);
      if ((++spins & 0x3ff) == 0) {
        if (*(volatile int*)&g_tc_abort) return;
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > g_tc_timeout) {
          if (atomicAdd(&g_tc_abort, 1) < 8)
            printf("futuredet_b200: mbarrier wait timed out (tag %d, block %d, thread %d, bar 0x%x, parity %u)\n", tag,
                   (int)blockIdx.x, (int)threadIdx.x, bar, parity);
          return;
        }
      }
    }
  } while (!done);
}
// Non-blocking probe (mbarrier.test_wait): issued one stage ahead so that its ~350-cycle latency overlaps the MMA
// issue of the current stage; the result is only consumed an iteration later.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
// One A stage of the MMA role as a single asm block, so that the instruction order is exactly
//   [probe of the NEXT stage's full barrier] [all tcgen05.mma of THIS stage] [tcgen05.commit] [read probe result]
// and the ~350-cycle latency of the barrier probe is hidden behind the MMA issue instead of heading every stage.
// MODE 0: single pass (A_hi*B_hi); 1: fused  A_hi*[B_hi|B_lo] (width 2N, idesc2) + A_lo*B_hi; 2: three products.
// Every lane executes it with identical (warp-uniform) operands; the elected lane issues mma / commit.
#define FD_UMMA_STAGE_OPERANDS                                                                                      \
  : "=r"(ready)                                                                                                     \
  : "r"(tmem_d), "l"(dA_hi), "l"(dA_lo), "l"(dB_hi), "l"(dB_lo), "r"(idesc), "r"(idesc2), "r"(acc0), "r"(next_bar), \
    "r"(next_parity), "r"(commit_bar)                                                                               \
  : "memory"
template <int MODE>
__device__ __forceinline__ uint32_t umma_stage(uint32_t tmem_d, uint64_t dA_hi, uint64_t dA_lo, uint64_t dB_hi,
                                               uint64_t dB_lo, uint32_t idesc, uint32_t idesc2, uint32_t acc0,
                                               uint32_t next_bar, uint32_t next_parity, uint32_t commit_bar) {
  uint32_t ready;
  if constexpr (MODE == 0) {
    asm volatile(
        "{\n\t"
        ".reg .pred pn, pa, q, one;\n\t"
        ".reg .b64 ah, al, bh, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pn, [%9], %10;\n\t"
        "setp.ne.b32 pa, %8, 0;\n\t"
        "setp.eq.b32 one, 0, 0;\n\t"
        "add.u64 ah, %2, 0; add.u64 al, %3, 0; add.u64 bh, %4, 0; add.u64 bl, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, pa;\n\t"
        "add.u64 ah, %2, 2; add.u64 al, %3, 2; add.u64 bh, %4, 2; add.u64 bl, %5, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, one;\n\t"
#if FD_TC_BK == 64
        "add.u64 ah, %2, 4; add.u64 al, %3, 4; add.u64 bh, %4, 4; add.u64 bl, %5, 4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, one;\n\t"
        "add.u64 ah, %2, 6; add.u64 al, %3, 6; add.u64 bh, %4, 6; add.u64 bl, %5, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, one;\n\t"
#endif
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "selp.u32 %0, 1, 0, pn;\n\t"
        "}"
        FD_UMMA_STAGE_OPERANDS);
  } else if constexpr (MODE == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred pn, pa, q, one;\n\t"
        ".reg .b64 ah, al, bh, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pn, [%9], %10;\n\t"
        "setp.ne.b32 pa, %8, 0;\n\t"
        "setp.eq.b32 one, 0, 0;\n\t"
        "add.u64 ah, %2, 0; add.u64 al, %3, 0; add.u64 bh, %4, 0; add.u64 bl, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %7, pa;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
        "add.u64 ah, %2, 2; add.u64 al, %3, 2; add.u64 bh, %4, 2; add.u64 bl, %5, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %7, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
#if FD_TC_BK == 64
        "add.u64 ah, %2, 4; add.u64 al, %3, 4; add.u64 bh, %4, 4; add.u64 bl, %5, 4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %7, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
        "add.u64 ah, %2, 6; add.u64 al, %3, 6; add.u64 bh, %4, 6; add.u64 bl, %5, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %7, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
#endif
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "selp.u32 %0, 1, 0, pn;\n\t"
        "}"
        FD_UMMA_STAGE_OPERANDS);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred pn, pa, q, one;\n\t"
        ".reg .b64 ah, al, bh, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pn, [%9], %10;\n\t"
        "setp.ne.b32 pa, %8, 0;\n\t"
        "setp.eq.b32 one, 0, 0;\n\t"
        "add.u64 ah, %2, 0; add.u64 al, %3, 0; add.u64 bh, %4, 0; add.u64 bl, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, pa;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bl, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
        "add.u64 ah, %2, 2; add.u64 al, %3, 2; add.u64 bh, %4, 2; add.u64 bl, %5, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bl, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
#if FD_TC_BK == 64
        "add.u64 ah, %2, 4; add.u64 al, %3, 4; add.u64 bh, %4, 4; add.u64 bl, %5, 4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bl, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
        "add.u64 ah, %2, 6; add.u64 al, %3, 6; add.u64 bh, %4, 6; add.u64 bl, %5, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bl, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], al, bh, %6, one;\n\t"
#endif
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "selp.u32 %0, 1, 0, pn;\n\t"
        "}"
        FD_UMMA_STAGE_OPERANDS);
  }
  return ready;
}
// Spin on the non-suspending probe.  try_wait parks the thread and its wake-up after the phase flips was measured at
// several hundred cycles; on the per-stage hand-shakes of the pipeline (MMA warp, producers, weight loader) that wake-up
// latency lands on the critical path every time a role is even slightly ahead of its partner.
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity, int tag = 0) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_test(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)&g_tc_abort) return;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > g_tc_timeout) {
        if (atomicAdd(&g_tc_abort, 1) < 8)
          printf("futuredet_b200: mbarrier spin timed out (tag %d, block %d, thread %d, bar 0x%x, parity %u)\n", tag,
                 (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        return;
      }
    }
  }
}
// Warp-level wait: ONE lane polls, the rest of the warp parks at __syncwarp().  32 lanes polling the same mbarrier
// word serialise in the shared-memory pipe (measured: ~450 cycles per already-completed wait vs ~100).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int tag = 0, uint32_t backoff_ns = 0) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, tag, backoff_ns);
  __syncwarp();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;     // src-size 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16_sz(uint32_t dst, const void* src, uint32_t sz) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// TMA gather: rows r0..r3 of the 2-D tensor (64 bf16 = 128 bytes each, starting at column `col`) -> 4 consecutive
// 128-byte shared-memory rows at `dst`, swizzled by the hardware (SWIZZLE_128B); a row index outside the tensor
// (the rulebook's -1) is filled with zeros and reads nothing; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
// TMA tile load of a 4-D box {64 channels, bw, bh, 1} at (c, x, y, b); coordinates may lie outside the tensor
// (negative / >= extent): those elements are zero-filled, which is exactly the convolution's zero padding
__device__ __forceinline__ void tma_tile4d(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int b, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c), "r"(x), "r"(y), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has landed (no wait in the thread)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}


// UMMA shared-memory descriptor, K-major, one swizzle atom per row (SWIZZLE_128B for 64-element stages, SWIZZLE_64B
// for 32): 8-row groups 8*TC_ROWB bytes apart (SBO), version 1, layout type 2 (128B) / 4 (64B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  constexpr uint64_t sbo = (8 * TC_ROWB) >> 4, layout = TC_BK == 64 ? 2 : 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// Warp-uniform variants: every lane of the (converged) MMA warp executes them with identical operands so the
// compiler keeps descriptors in uniform registers; one elected lane issues the instruction.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- weight packing ------------------------------------------------------------------------------
// W fp32 [K, Cin, Cout] -> for every (N tile tn, K stage ks) one contiguous block that is a byte-exact image of
// the shared-memory B stage: [2 planes (hi, lo)][NT rows][TC_BK bf16, 16-byte chunks XOR-swizzled by the row].
// A stage's weights are then ONE cp.async.bulk (TMA bulk copy) instead of per-thread gathers.
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w, int K, int cin, int cout, int NT, int n_tiles_n, int n_kstages,
                    __nv_bfloat16* __restrict__ out) {
  const long long per_block = 2LL * NT * TC_BK;
  const long long total = (long long)n_tiles_n * n_kstages * NT * TC_BK;     // elements of one plane
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(e % TC_BK);
    long long r = e / TC_BK;
    const int nrow = (int)(r % NT); r /= NT;
    const int ks = (int)(r % n_kstages);
    const int tn = (int)(r / n_kstages);
    const int n = tn * NT + nrow, kf = ks * TC_BK + kc;
    float v = 0.f;
    if (n < cout && kf < K * cin) {
      int k = kf / cin, ci = kf - k * cin;
      v = w[((size_t)k * cin + ci) * cout + n];
    }
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const long long blk = ((long long)tn * n_kstages + ks) * per_block;
    const int off = nrow * TC_BK + ((((kc >> 3) ^ ((nrow >> TC_SWZ_SHIFT) & (TC_CHUNKS - 1))) << 3) | (kc & 7));
    out[blk + off] = hi;
    out[blk + (long long)NT * TC_BK + off] = lo;
  }
}

struct TcArgs {
  // TMA descriptor of the input activation matrix viewed as a 2-D bf16 tensor [rows, hi plane | lo plane] with a
  // {64 channels, 1 row} box and SWIZZLE_128B: the operand of cp.async.bulk.tensor...tile::gather4 (valid when tma)
  alignas(64) CUtensorMap tmap;
  ConvArgs c;
  const __nv_bfloat16* wp;   // packed weights
  const uint32_t* tile_mask; // per 128-row tile: active kernel offsets (NULL: all)
  int cout_pad, ktot_pad;
  int n_tiles_n;
  int T;                     // M tiles per super-tile (weight reuse factor)
  int split;                 // 1: bf16x3, 0: single-pass bf16
  int dbg;                   // FD_TC_DEBUG ablation bits (perf triage only): 1 no A gather, 2 no MMA, 4 no B copy, 8 no stores
  int tma;                   // 0: per-thread cp.async gather; 1: TMA gather4 (wide split-bf16 inputs);
                             // 2: dense 2-D stride-1 convs, TMA TILE loads: an M tile is a bw x bh pixel patch and
                             //    every (tap, 64-channel chunk) stage is ONE 4-D box per plane, shifted by the tap,
                             //    out-of-image pixels zero-filled by the copy engine (tmap is 4-D [B,H,W,C] then)
  int bw, bh, tiles_x, tiles_y;   // tma == 2: patch shape and patches per image
  int nbr_vec;               // rulebook rows may be read as int4 (16-byte aligned table)
};

template <int NT> struct TcCfg {
  static constexpr int A_BYTES = 2 * TC_A_PLANE;
  static constexpr int B_BYTES = 2 * NT * TC_ROWB;
  static constexpr int SB = (TC_BK == 64 && NT >= FD_TC_SB2_MIN_NT) ? 2 : (NT == 64 ? FD_TC_SB64 : 4);   // B ring slots (weights are requested SB-1 K stages ahead)
  // NT <= 64: the split products A_hi*B_hi and A_hi*B_lo are issued as ONE MMA of width 2*NT against the adjacent
  // [B_hi | B_lo] planes (each MMA re-reads its whole A tile from shared memory, which is what bounds narrow tiles),
  // so a tile owns two accumulator column blocks that the epilogue adds.
  static constexpr bool FUSE_N = NT <= 64;
  static constexpr int ACC_COLS = FUSE_N ? 2 * NT : NT;
  static constexpr int TMAX = 256 / ACC_COLS > 4 ? 4 : 256 / ACC_COLS;   // M tiles sharing one weight fetch
  static constexpr int TMEM_COLS = 2 * TMAX * ACC_COLS < 32 ? 32 : 2 * TMAX * ACC_COLS;
  static constexpr int IDX_BYTES = TMAX * TC_DENSE_MAXK * TC_BM * 4;   // dense-mode row index cache
  // the A ring takes every byte that is left of the 227 KB a CTA may own
  static constexpr int FIXED = 1024 + SB * B_BYTES + IDX_BYTES + 2 * TC_SS_MAX * 4 + TC_EPI_BYTES + (2 * 16 + 2 * SB + 4) * 8 + 64;
  static constexpr int SA_FIT = (232448 - FIXED) / A_BYTES;
  static constexpr int SA = SA_FIT > 12 ? 12 : SA_FIT;            // A ring slots
  static_assert(SA >= TC_GROUPS, "every producer group needs a slot of its own");
  static constexpr int NBAR = 2 * SA + 2 * SB + 4;
  static constexpr size_t SMEM = 1024 + (size_t)SA * A_BYTES + (size_t)SB * B_BYTES + IDX_BYTES + 2 * TC_SS_MAX * 4 + TC_EPI_BYTES + NBAR * 8 + 64;
  static_assert(SMEM <= 232448, "shared memory budget");
  // tall-stage layout: A ring of TC_TALL_SLOTS x 36 KB, no dense index cache
  static constexpr size_t SMEM_TALL = 1024 + (size_t)TC_TALL_SLOTS * TC_TALL_BYTES + (size_t)SB * B_BYTES + 2 * TC_SS_MAX * 4 + TC_EPI_BYTES + NBAR * 8 + 64;
  static_assert(SMEM_TALL <= 232448 && SA >= TC_TALL_SLOTS, "shared memory budget (tall stages)");
};

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* v) {
  if constexpr (N == 16) {
    tmem_ld16(taddr, v);
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  }
}

// (128 registers is the hard limit of 14 warps: an SM sub-partition holds 4 of them in its 16 K registers; __maxnreg__(144)
// compiles without spills but cannot launch)
// TALL: the tall-stage dense 3x3 path (t.tma == 3) is its own instantiation -- the kernel sits at the 128-register ceiling, and
// any code added to the common one changes what ptxas spills in the gather producers' loop (measured: sparse family +8 %)
template <int NT, bool TALL>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcArgs t) {
  using Cfg = TcCfg<NT>;
  constexpr int SA = Cfg::SA, SB = Cfg::SB;
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, NT);
  constexpr uint32_t IDESC2 = umma_idesc_bf16(TC_BM, Cfg::FUSE_N ? 2 * NT : NT);
  constexpr int ACC = Cfg::ACC_COLS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms want 1024-B (128B) / 512-B (64B) alignment
  constexpr bool tall = TALL;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + (tall ? TC_TALL_SLOTS * TC_TALL_BYTES : SA * Cfg::A_BYTES);
  int* s_idx = (int*)(b_ring + SB * Cfg::B_BYTES);                // [TMAX][TC_DENSE_MAXK][128] (dense mode)
  float* s_scale = (float*)((uint8_t*)s_idx + (tall ? 0 : Cfg::IDX_BYTES));    // [TC_SS_MAX] folded BN scale, then shift
  float* s_shift = s_scale + TC_SS_MAX;
  uint8_t* s_epi = (uint8_t*)(s_shift + TC_SS_MAX);               // [4 epilogue warps][TC_EPI_WARP_BYTES]
  uint64_t* bars = (uint64_t*)(s_epi + TC_EPI_BYTES);
  uint64_t* a_full = bars;                  // [SA]
  uint64_t* a_empty = a_full + SA;          // [SA]
  uint64_t* b_full = a_empty + SA;          // [SB]
  uint64_t* b_empty = b_full + SB;          // [SB]
  uint64_t* t_full = b_empty + SB;          // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* s_tmem = (uint32_t*)(t_empty + 2);

  const ConvArgs& a = t.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.d_n ? min(*a.d_n, a.n_cap) : a.n_cap;
  const int T = t.T;
  const bool tiled = t.tma >= 2;
  const int tiles_img = t.tiles_x * t.tiles_y;
  const int n_tiles_m = tiled ? (n / (a.Hout * a.Wout)) * tiles_img : (n + TC_BM - 1) / TC_BM;
  // tiled mode: output pixel of accumulator row r of M tile tm (or -1: outside the patch / the image)
  auto tile_row_to_o = [&](int tm, int r) -> int {
    const int b = tm / tiles_img, rem = tm - b * tiles_img;
    const int ty = rem / t.tiles_x, tx = rem - ty * t.tiles_x;
    const int h = TALL ? r / TC_TALL_BW : r / t.bw, w = r - h * (TALL ? TC_TALL_BW : t.bw);
    const int oy = ty * t.bh + h, ox = tx * t.bw + w;
    if (h >= t.bh || oy >= a.Hout || ox >= a.Wout) return -1;
    return (b * a.Hout + oy) * a.Wout + ox;
  };
  const int n_super = (n_tiles_m + T - 1) / T;
  const int n_units = n_super * t.n_tiles_n;
  const int ktot = a.K * a.cin;
  const int n_kstages = (ktot + TC_BK - 1) / TC_BK;
  const bool wide = a.cin >= TC_BK;
  const int spo = wide ? a.cin / TC_BK : 1;      // stages per kernel offset   (wide:   Cin multiple of TC_BK)
  const int opk = wide ? 1 : TC_BK / a.cin;      // kernel offsets per stage   (narrow: Cin divides TC_BK)
  // K stages come in groups: a kernel offset's `spo` stages (wide) or one stage covering `opk` offsets (narrow)
  const int n_groups = wide ? a.K : n_kstages;   // <= 32
  const int spg = spo;                           // stages per group (narrow: 1)
  constexpr int G = TC_GROUPS, GT = TC_PRODUCERS / G;
  const bool ss_smem = t.cout_pad <= TC_SS_MAX;

  if (threadIdx.x == 0) {
    // A stage complete: every thread of the filling group arrives (cp.async path) / one expect_tx arrival + the
    // stage's bytes (TMA path)
    // (tall stages: a slot is free when the commits of all three vertical taps have arrived)
    for (int s = 0; s < SA; ++s) { mbar_init(smem_u32(&a_full[s]), t.tma ? 1 : GT); mbar_init(smem_u32(&a_empty[s]), tall ? 3 : 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&t_full[s]), 1); mbar_init(smem_u32(&t_empty[s]), 128); }
    fence_barrier_init();
  }
  if (ss_smem)
    for (int c = threadIdx.x; c < t.cout_pad; c += TC_THREADS) {
      s_scale[c] = (a.scale && c < a.cout) ? a.scale[c] : 1.f;
      s_shift[c] = (a.shift && c < a.cout) ? a.shift[c] : 0.f;
    }
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // Groups of a super-tile that touch an active kernel offset, as a bit mask (group 0 is forced for an all-empty
  // unit so that the accumulator is always defined).  Identical in every role; the unit's stage list is the stages
  // of its active groups in ascending order.  Without a rulebook mask (dense 2-D convolutions) every group is active.
  const uint32_t all_groups = n_groups >= 32 ? 0xffffffffu : ((1u << n_groups) - 1u);
  auto unit_gmask = [&](int st) -> uint32_t {
    if (t.tile_mask == nullptr) return all_groups;
    uint32_t m = 0;
    for (int i = 0; i < T; ++i) {
      const int tm = st * T + i;
      if (tm < n_tiles_m) m |= __ldg(t.tile_mask + tm);
    }
    uint32_t gm = m;
    if (!wide) {
      gm = 0;
      for (int ks = 0; ks < n_kstages; ++ks)
        if (m & (((1u << opk) - 1u) << (ks * opk))) gm |= 1u << ks;
    }
    gm &= all_groups;
    return gm ? gm : 1u;
  };

  if (warp < TC_PRODUCER_WARPS && tall) {
    // ===================================== PRODUCER (dense 3x3, tall TMA stages) =====================================
    // one elected thread: per (horizontal tap, 64-channel chunk, M tile) two box loads {64 ch, 8 px, 18 lines}
    if (warp == 0 && lane == 0) {
      const uint32_t a_ring_u32 = smem_u32(a_ring);
      const CUtensorMap* tmap = &t.tmap;
      uint32_t slot = 0, phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int st = unit / t.n_tiles_n;
        const int live = min(T, n_tiles_m - st * T);
        for (int kx = 0; kx < 3; ++kx) {
          for (int sub = 0; sub < spo; ++sub) {
            for (int ti = 0; ti < live; ++ti) {
              const int tm = st * T + ti;
              const int b = tm / tiles_img, rem = tm - b * tiles_img;
              const int ty = rem / t.tiles_x, tx = rem - ty * t.tiles_x;
              mbar_wait(smem_u32(&a_empty[slot]), phase ^ 1, 2);
              const uint32_t fbar = smem_u32(&a_full[slot]);
              const uint32_t dst = a_ring_u32 + slot * TC_TALL_BYTES;
              mbar_arrive_expect_tx(fbar, TC_TALL_BYTES);
              const int x0 = tx * TC_TALL_BW - a.pw + kx, y0 = ty * TC_TALL_BH - a.ph;
              tma_tile4d(dst, tmap, sub * TC_BK, x0, y0, b, fbar);
              tma_tile4d(dst + TC_TALL_PLANE, tmap, a.in_ctot + sub * TC_BK, x0, y0, b, fbar);
              if (++slot == (uint32_t)TC_TALL_SLOTS) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp < TC_PRODUCER_WARPS && tiled) {
    // ===================================== PRODUCER (dense 2-D, TMA tile loads) =====================================
    // one elected thread: per A stage (M tile x tap x 64-channel chunk) two box loads (hi plane, lo plane)
    if (warp == 0 && lane == 0) {
      const uint32_t a_ring_u32 = smem_u32(a_ring);
      const CUtensorMap* tmap = &t.tmap;
      const uint32_t stage_bytes = (uint32_t)(2 * t.bw * t.bh * TC_ROWB);
      uint32_t slot = 0, phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int st = unit / t.n_tiles_n;
        const int live = min(T, n_tiles_m - st * T);
        for (int k = 0; k < a.K; ++k) {
          const int ky = k / a.kw, kx = k - ky * a.kw;
          for (int sub = 0; sub < spo; ++sub) {
            for (int ti = 0; ti < live; ++ti) {
              const int tm = st * T + ti;
              const int b = tm / tiles_img, rem = tm - b * tiles_img;
              const int ty = rem / t.tiles_x, tx = rem - ty * t.tiles_x;
              mbar_wait(smem_u32(&a_empty[slot]), phase ^ 1, 2);
              const uint32_t fbar = smem_u32(&a_full[slot]);
              const uint32_t dst = a_ring_u32 + slot * Cfg::A_BYTES;
              if (t.dbg & 1) {
                mbar_arrive(fbar);
              } else {
                mbar_arrive_expect_tx(fbar, stage_bytes);
                const int x0 = tx * t.bw - a.pw + kx, y0 = ty * t.bh - a.ph + ky;
                tma_tile4d(dst, tmap, sub * TC_BK, x0, y0, b, fbar);
                tma_tile4d(dst + TC_A_PLANE, tmap, a.in_ctot + sub * TC_BK, x0, y0, b, fbar);
              }
              if (++slot == (uint32_t)SA) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp < TC_PRODUCER_WARPS && t.tma) {
    // ===================================== PRODUCERS (A gather by TMA) =====================================
    // One warp per group (warps 0..G-1); group g fills emitted A stages g, g+G, ... : lane l owns rows 4l..4l+3 of the
    // 128-row tile, reads their four rulebook entries as one int4 (prefetched one owned stage ahead) and issues two
    // gather4 copies (hi plane, lo plane).  The copy engine does the address generation, the swizzled shared-memory
    // writes and the byte-counted arrival on the stage's mbarrier; the SM spends 2 instructions per stage and lane.
    if (warp < G) {
      const int grp = warp;
      const uint32_t a_ring_u32 = smem_u32(a_ring), idx_u32 = smem_u32(s_idx);
      const bool table = a.mode == FD_GATHER_TABLE;
      const CUtensorMap* tmap = &t.tmap;
      uint32_t c_base = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int st = unit / t.n_tiles_n;
        const int live = min(T, n_tiles_m - st * T);
        if (!table) {
          named_bar_sync(1, G * 32);                       // every group is done reading the previous unit's indices
          for (int e = threadIdx.x; e < live * a.K * TC_BM; e += G * 32) {
            const int ti = e / (a.K * TC_BM), rem = e - ti * a.K * TC_BM;
            const int k = rem >> 7, r = rem & (TC_BM - 1);
            const int o = (st * T + ti) * TC_BM + r;
            sts_u32(idx_u32 + (uint32_t)((ti * TC_DENSE_MAXK + k) * TC_BM + r) * 4, (uint32_t)(o < n ? gather_row(a, o, k) : -1));
          }
          named_bar_sync(1, G * 32);
        }
        const uint32_t gmask = unit_gmask(st);
        const int n_emit = __popc(gmask) * spg * live;
        int p = (int)((grp + G - (c_base % G)) % G);
        uint32_t slot = (c_base + p) % SA, phase = ((c_base + p) / SA) & 1;
        uint32_t rem = gmask;
        int act_idx = 0;
        int4 nxt = make_int4(-1, -1, -1, -1);
        int col_n = 0;
        auto fetch = [&](int pos) {
          const int want = pos / live;
          const int ti = pos - want * live;
          const int gi = want / spg, sub = want - gi * spg;
          while (act_idx < gi) { rem &= rem - 1; ++act_idx; }
          const int kk = __ffs((int)rem) - 1;              // wide layers: one kernel offset per group of stages
          col_n = sub * TC_BK;
          const int row0 = (st * T + ti) * TC_BM + 4 * lane;
          if (table) {
            const int32_t* nrow = a.nbr + (size_t)kk * a.nbr_stride + row0;
            if (row0 + 3 < n && t.nbr_vec) {
              nxt = __ldg(reinterpret_cast<const int4*>(nrow));
            } else {
              nxt.x = row0 < n ? __ldg(nrow) : -1;
              nxt.y = row0 + 1 < n ? __ldg(nrow + 1) : -1;
              nxt.z = row0 + 2 < n ? __ldg(nrow + 2) : -1;
              nxt.w = row0 + 3 < n ? __ldg(nrow + 3) : -1;
            }
          } else {
            const uint32_t ib = idx_u32 + (uint32_t)((ti * TC_DENSE_MAXK + kk) * TC_BM + 4 * lane) * 4;
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(nxt.x), "=r"(nxt.y), "=r"(nxt.z), "=r"(nxt.w) : "r"(ib) : "memory");
          }
        };
        if (p < n_emit) fetch(p);
        for (; p < n_emit; p += G) {
          const int4 cur = nxt;
          const int col = col_n;
          if (p + G < n_emit) fetch(p + G);
          mbar_wait_warp(smem_u32(&a_empty[slot]), phase ^ 1, 2);
          const uint32_t fbar = smem_u32(&a_full[slot]);
          const uint32_t dst = a_ring_u32 + slot * Cfg::A_BYTES + lane * (4 * TC_ROWB);
          slot += G;
          while (slot >= (uint32_t)SA) { slot -= SA; phase ^= 1; }
          if (t.dbg & 1) {                                   // triage: no data movement
            if (lane == 0) mbar_arrive(fbar);
            continue;
          }
          if (lane == 0) mbar_arrive_expect_tx(fbar, Cfg::A_BYTES);
          __syncwarp();
          tma_gather4(dst, tmap, col, cur.x, cur.y, cur.z, cur.w, fbar);
          tma_gather4(dst + TC_A_PLANE, tmap, a.in_ctot + col, cur.x, cur.y, cur.z, cur.w, fbar);
        }
        c_base += (uint32_t)n_emit;
      }
    }
  } else if (warp < TC_PRODUCER_WARPS) {
    // ===================================== PRODUCERS (A gather) =====================================
    // 4 groups of 64 threads; group g fills emitted A stages g, g+4, ... on its own, so several stages are being
    // issued concurrently, and the rulebook indices of a group's next stage are prefetched while it waits for a slot.
    constexpr int ROWS_PER_PASS = GT / TC_CHUNKS;
    constexpr int PASSES = TC_BM / ROWS_PER_PASS;
    static_assert(ROWS_PER_PASS % 8 == 0 && PASSES % 4 == 0, "swizzle phase must not change between passes");
    const int tid = threadIdx.x;                 // 0..TC_PRODUCERS-1
    const int grp = tid / GT, gt = tid % GT;
    const int j = gt % TC_CHUNKS, rbase = gt / TC_CHUNKS;   // 16-byte chunk column, first row (rows rbase + p*ROWS_PER_PASS)
    const uint32_t a_ring_u32 = smem_u32(a_ring), idx_u32 = smem_u32(s_idx);
    const uint32_t a_off0 = rbase * TC_ROWB + ((j ^ ((rbase >> TC_SWZ_SHIFT) & (TC_CHUNKS - 1))) << 4);   // swizzle phase is the same for every pass
    const int jk = wide ? 0 : (j * 8) / a.cin;   // which of the stage's offsets this thread's chunk belongs to
    const int jc = wide ? j * 8 : (j * 8) % a.cin;
    const uint32_t row_bytes = (uint32_t)a.in_stride * 4;
    const char* in_b = reinterpret_cast<const char*>(a.in);
    const char* in_lo = in_b + (size_t)a.in_ctot * 2;
    const bool table = a.mode == FD_GATHER_TABLE;
    uint32_t c_base = 0;                         // emitted A stages before this unit (ring position bookkeeping)
    int ptrace_i = 0;

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int st = unit / t.n_tiles_n;
      const int live = min(T, n_tiles_m - st * T);
      if (!table) {
        // dense 2-D: row indices are arithmetic; computed once per unit into shared memory
        named_bar_sync(1, TC_PRODUCERS);                   // every group is done reading the previous unit's indices
        for (int e = tid; e < live * a.K * TC_BM; e += TC_PRODUCERS) {
          const int ti = e / (a.K * TC_BM), rem = e - ti * a.K * TC_BM;
          const int k = rem >> 7, r = rem & (TC_BM - 1);
          const int o = (st * T + ti) * TC_BM + r;
          sts_u32(idx_u32 + (uint32_t)((ti * TC_DENSE_MAXK + k) * TC_BM + r) * 4, (uint32_t)(o < n ? gather_row(a, o, k) : -1));
        }
        named_bar_sync(1, TC_PRODUCERS);
      }
      const uint32_t gmask = unit_gmask(st);
      const int n_emit = __popc(gmask) * spg * live;
      // this group's positions in the unit's emitted-stage list: p = p0, p0 + G, ...
      int p = (int)((grp + G - (c_base % G)) % G);
      uint32_t slot = (c_base + p) % SA, phase = ((c_base + p) / SA) & 1;
      uint32_t rem = gmask;                         // active groups not yet passed
      int act_idx = 0;                              // index (among active groups) of the lowest bit of `rem`
      int src[PASSES], kk_n = 0, ch_n = 0, ti_n = 0;
      // (ks, ti) of position p, this thread's kernel offset / channel, and the 16 gathered row indices
      auto fetch = [&](int pos) {
        const int want = pos / live;
        ti_n = pos - want * live;
        const int gi = want / spg, sub = want - gi * spg;
        while (act_idx < gi) { rem &= rem - 1; ++act_idx; }
        const int g = __ffs((int)rem) - 1;
        kk_n = wide ? g : g * opk + jk;
        ch_n = wide ? sub * TC_BK + jc : jc;
        const bool kvalid = kk_n < a.K;
        const int row0 = (st * T + ti_n) * TC_BM + rbase;
        if (table) {
          const int32_t* nrow = a.nbr + (size_t)(kvalid ? kk_n : 0) * a.nbr_stride + row0;
#pragma unroll
          for (int q = 0; q < PASSES; ++q)
            src[q] = (kvalid && row0 + q * ROWS_PER_PASS < n) ? __ldg(nrow + q * ROWS_PER_PASS) : -1;
        } else {
          const uint32_t ib = idx_u32 + (uint32_t)((ti_n * TC_DENSE_MAXK + (kvalid ? kk_n : 0)) * TC_BM + rbase) * 4;
#pragma unroll
          for (int q = 0; q < PASSES; ++q) src[q] = kvalid ? (int)lds_u32(ib + q * ROWS_PER_PASS * 4) : -1;
        }
      };
      if (p < n_emit) fetch(p);
      for (; p < n_emit; p += G) {
        int cur[PASSES];
#pragma unroll
        for (int q = 0; q < PASSES; ++q) cur[q] = src[q];
        const int ch = ch_n;
        if (p + G < n_emit) fetch(p + G);                  // prefetch the next owned stage's indices (in flight during the wait)
        if (tid == 0) TC_TRACE(1, 4 * ptrace_i, clock64());
        // every producer thread parks in try_wait.  Measured alternatives, all slower: every thread polling test_wait, one
        // polling lane per warp + __syncwarp (256 / 8 threads hammering the mbarrier words), and a manager warp that watches
        // the barrier and releases the group through a named barrier (c64 3.6 vs 3.4 ms).  The group observes a flip
        // ~1000 cycles late either way: its wait queues behind the 64 cp.async the warp has just issued (LSU, in order).
        mbar_wait(smem_u32(&a_empty[slot]), phase ^ 1, 2);
        if (tid == 0) { TC_TRACE(1, 4 * ptrace_i + 1, clock64()); TC_TRACE(1, 4 * ptrace_i + 3, (long long)(c_base + p)); }
        const uint32_t fbar = smem_u32(&a_full[slot]);
        const uint32_t dst0 = a_ring_u32 + slot * Cfg::A_BYTES + a_off0;
        slot += G;
        while (slot >= SA) { slot -= SA; phase ^= 1; }
        // ---- A: gather 128 rows x TC_BK channels (this thread: chunk j of rows rbase + ROWS_PER_PASS q)
        if (t.dbg & 1) {
          mbar_arrive(fbar);
        } else if (a.in_fmt == FD_FMT_SPLIT_BF16) {
          // pre-split bf16 hi/lo rows: pure 16-byte cp.async copies (zero-filled where the rulebook has no
          // neighbour); nothing is waited for -- the hardware arrives on the full barrier when they land
          const char* gh = in_b + ch * 2;
          const char* gl = in_lo + ch * 2;
#pragma unroll
#pragma unroll
          for (int q = 0; q < PASSES; ++q) {
            const uint32_t sz = cur[q] >= 0 ? 16u : 0u;
            const size_t goff = (size_t)(uint32_t)max(cur[q], 0) * row_bytes;
            cp_async16_sz(dst0 + q * ROWS_PER_PASS * TC_ROWB, gh + goff, sz);
            cp_async16_sz(dst0 + q * ROWS_PER_PASS * TC_ROWB + TC_A_PLANE, gl + goff, sz);
          }
          cp_async_mbar_arrive_noinc(fbar);
        } else {
          // fp32 rows (the stem reading voxel features): split into bf16 hi/lo in registers, 4 rows at a time
#pragma unroll
          for (int p0 = 0; p0 < PASSES; p0 += 4) {
            float4 v[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (cur[p0 + q] >= 0) {
                const float4* gp = reinterpret_cast<const float4*>(a.in + (size_t)cur[p0 + q] * a.in_stride + ch);
                v[q][0] = __ldg(gp);
                v[q][1] = __ldg(gp + 1);
              } else {
                v[q][0] = v[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float f[8] = {v[q][0].x, v[q][0].y, v[q][0].z, v[q][0].w, v[q][1].x, v[q][1].y, v[q][1].z, v[q][1].w};
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                float2 hf = __bfloat1622float2(h);
                __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
                hi[e] = *reinterpret_cast<uint32_t*>(&h);
                lo[e] = *reinterpret_cast<uint32_t*>(&l);
              }
              sts_u128(dst0 + (p0 + q) * ROWS_PER_PASS * TC_ROWB, hi[0], hi[1], hi[2], hi[3]);
              sts_u128(dst0 + (p0 + q) * ROWS_PER_PASS * TC_ROWB + TC_A_PLANE, lo[0], lo[1], lo[2], lo[3]);
            }
          }
          fence_proxy_async();                             // generic-proxy smem writes -> visible to the tensor core
          mbar_arrive(fbar);
        }
        if (tid == 0) TC_TRACE(1, 4 * ptrace_i + 2, clock64());
        ++ptrace_i;
      }
      c_base += (uint32_t)n_emit;
    }
    cp_async_wait_all();
    (void)ptrace_i;
  } else if (warp == TC_MMA_WARP) {
    // ===================================== MMA ISSUER =====================================
    // the whole warp walks the loop converged with warp-uniform operands; an elected lane issues mma / commit
    uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
    const uint32_t a_ring_u32 = smem_u32(a_ring), b_ring_u32 = smem_u32(b_ring);
    uint32_t a_ready = 0, b_ready = 0;          // probe results for the upcoming A / B slot (issued one step ahead)
    int it = 0, trace_i = 0;
    if (tall) {
      // tall stages: for every (horizontal tap, chunk) the T tiles' A stages are acquired once and read by the three vertical
      // taps at 0 / 1024 / 2048 bytes; every tap commits to the slot's empty barrier, which expects three arrivals.  `a_slot` is the next slot to be
      // acquired; every stage's asm block probes its barrier while the MMAs issue.
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int st = unit / t.n_tiles_n;
        const int live = min(T, n_tiles_m - st * T);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(smem_u32(&t_empty[acc]), acc_phase ^ 1, 3);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int kx = 0; kx < 3; ++kx) {
          for (int sub = 0; sub < spo; ++sub) {
            const uint32_t g_slot = a_slot;                  // first slot of this tap column's tiles
            for (int ky = 0; ky < 3; ++ky) {
              if (!b_ready) mbar_spin(smem_u32(&b_full[b_slot]), b_phase, 4);
              const uint32_t sB_hi = b_ring_u32 + b_slot * Cfg::B_BYTES, sB_lo = sB_hi + NT * TC_ROWB;
              const uint64_t dB_hi = umma_desc_sw128(sB_hi), dB_lo = umma_desc_sw128(sB_lo);
              const uint32_t nb_slot = b_slot + 1 == SB ? 0 : b_slot + 1, nb_phase = b_slot + 1 == SB ? b_phase ^ 1 : b_phase;
              b_ready = mbar_test(smem_u32(&b_full[nb_slot]), nb_phase);
              uint32_t sl = g_slot;
              for (int ti = 0; ti < live; ++ti) {
                if (ky == 0) {
                  if (!a_ready) mbar_spin(smem_u32(&a_full[a_slot]), a_phase, 5);
                  a_ready = 0;
                  if (++a_slot == (uint32_t)TC_TALL_SLOTS) { a_slot = 0; a_phase ^= 1; }
                }
                tc_fence_after();
                const uint32_t sA_hi = a_ring_u32 + sl * TC_TALL_BYTES + (uint32_t)ky * (TC_TALL_BW * TC_ROWB);
                const uint64_t dA_hi = umma_desc_sw128(sA_hi), dA_lo = umma_desc_sw128(sA_hi + TC_TALL_PLANE);
                const uint32_t tmem_d = tmem_base + (uint32_t)((acc * T + ti) * ACC);
                const uint32_t nbar = smem_u32(&a_full[a_slot]);
                const uint32_t cbar = smem_u32(&a_empty[sl]);       // third arrival (ky == 2) completes the phase
                uint32_t r;
                if (!t.split) r = umma_stage<0>(tmem_d, dA_hi, dA_lo, dB_hi, dB_lo, IDESC, IDESC2, accumulate, nbar, a_phase, cbar);
                else if (Cfg::FUSE_N) r = umma_stage<1>(tmem_d, dA_hi, dA_lo, dB_hi, dB_lo, IDESC, IDESC2, accumulate, nbar, a_phase, cbar);
                else r = umma_stage<2>(tmem_d, dA_hi, dA_lo, dB_hi, dB_lo, IDESC, IDESC2, accumulate, nbar, a_phase, cbar);
                a_ready |= r;
                if (++sl == (uint32_t)TC_TALL_SLOTS) sl = 0;
              }
              umma_commit_elect(smem_u32(&b_empty[b_slot]));
              b_slot = nb_slot; b_phase = nb_phase;
              accumulate = 1;
            }
          }
        }
        umma_commit_elect(smem_u32(&t_full[acc]));
      }
    } else
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int st = unit / t.n_tiles_n;
      const int live = min(T, n_tiles_m - st * T);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_act = __popc(unit_gmask(st)) * spg;
      mbar_wait(smem_u32(&t_empty[acc]), acc_phase ^ 1, 3);        // epilogue has drained this accumulator buffer
      tc_fence_after();
      uint32_t accumulate = 0;
      for (int ia = 0; ia < n_act; ++ia) {
        // the single-thread roles (MMA issuer, weight loader) poll with the non-suspending test_wait: a parked try_wait wakes
        // up ~1000 cycles after the flip, which lands on the critical path of every stage (FD_TC_DEBUG & 128: park instead)
        if (!b_ready) { if (!(t.dbg & 128)) mbar_spin(smem_u32(&b_full[b_slot]), b_phase, 4); else mbar_wait(smem_u32(&b_full[b_slot]), b_phase, 4); }
        const uint32_t sB_hi = b_ring_u32 + b_slot * Cfg::B_BYTES, sB_lo = sB_hi + NT * TC_ROWB;
        const uint64_t dB_hi = umma_desc_sw128(sB_hi), dB_lo = umma_desc_sw128(sB_lo);
        const uint32_t nb_slot = b_slot + 1 == SB ? 0 : b_slot + 1, nb_phase = b_slot + 1 == SB ? b_phase ^ 1 : b_phase;
        b_ready = mbar_test(smem_u32(&b_full[nb_slot]), nb_phase);   // consumed at the next K stage
        for (int ti = 0; ti < live; ++ti) {
          if (lane == 0) TC_TRACE(3, 4 * trace_i + 2, clock64());
          if (!a_ready) { if (!(t.dbg & 128)) mbar_spin(smem_u32(&a_full[a_slot]), a_phase, 5); else mbar_wait(smem_u32(&a_full[a_slot]), a_phase, 5); }
          if (lane == 0) TC_TRACE(0, 2 * trace_i, clock64());
          tc_fence_after();
          if (lane == 0) TC_TRACE(3, 4 * trace_i + 3, clock64());
          const uint32_t na_slot = a_slot + 1 == SA ? 0 : a_slot + 1, na_phase = a_slot + 1 == SA ? a_phase ^ 1 : a_phase;
          const uint32_t sA_hi = a_ring_u32 + a_slot * Cfg::A_BYTES, sA_lo = sA_hi + TC_A_PLANE;
          const uint64_t dA_hi = umma_desc_sw128(sA_hi), dA_lo = umma_desc_sw128(sA_lo);
          const uint32_t tmem_d = tmem_base + (uint32_t)((acc * T + ti) * ACC);
          const uint32_t nbar = smem_u32(&a_full[na_slot]), cbar = smem_u32(&a_empty[a_slot]);
          if (t.dbg & 2) {
            if (t.dbg & 256) { if (lane == 0) mbar_arrive(cbar); __syncwarp(); }   // triage: plain arrive instead of tcgen05.commit
            else umma_commit_elect(cbar);
            a_ready = 0;
          } else if (!t.split) {
            a_ready = umma_stage<0>(tmem_d, dA_hi, dA_lo, dB_hi, dB_lo, IDESC, IDESC2, accumulate, nbar, na_phase, cbar);
          } else if (Cfg::FUSE_N) {
            a_ready = umma_stage<1>(tmem_d, dA_hi, dA_lo, dB_hi, dB_lo, IDESC, IDESC2, accumulate, nbar, na_phase, cbar);
          } else {
            a_ready = umma_stage<2>(tmem_d, dA_hi, dA_lo, dB_hi, dB_lo, IDESC, IDESC2, accumulate, nbar, na_phase, cbar);
          }
          if (lane == 0) TC_TRACE(0, 2 * trace_i + 1, clock64());
          if (t.dbg & 512) {                                   // triage: how long after its issue does the commit arrive?
            mbar_spin(cbar, a_phase, 9);
            if (lane == 0) TC_TRACE(3, 4 * trace_i, clock64());
          }
          ++trace_i;
          a_slot = na_slot; a_phase = na_phase;
        }
        umma_commit_elect(smem_u32(&b_empty[b_slot]));     // weight slot free after the unit's last tile
        b_slot = nb_slot; b_phase = nb_phase;
        accumulate = 1;
      }
      umma_commit_elect(smem_u32(&t_full[acc]));
    }
    __syncwarp();
    if ((t.dbg & 1024) && lane == 0) *(volatile uint32_t*)s_tmem = 0xffffffffu;      // (tmem_base was read by everyone long ago)
  } else if (warp == TC_B_WARP) {
    // ===================================== WEIGHT LOADER =====================================
    // one lane: per (unit, active K stage) one TMA bulk copy of the pre-swizzled weight stage, shared by the T tiles
    if (lane == 1 && (t.dbg & 1024) && blockIdx.x == 0) {
      // triage monitor: when does the a_empty barrier of emitted stage g actually flip? (trace role 3, word 4 g + 1)
      for (int g = 0; g < TC_TRACE_N / 4; ++g) {
        bool stop = false;
        while (!mbar_test(smem_u32(&a_empty[g % SA]), (uint32_t)((g / SA) & 1)))
          if (*(volatile uint32_t*)s_tmem == 0xffffffffu) { stop = true; break; }      // the MMA warp is done
        if (stop) break;
        g_tc_trace[3][4 * g + 1] = clock64();
      }
    }
    if (lane == 0) {
      uint32_t b_slot = 0, b_phase = 0;
      const uint32_t b_ring_u32 = smem_u32(b_ring);
      const char* wblocks = reinterpret_cast<const char*>(t.wp);
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int st = unit / t.n_tiles_n, tn = unit - st * t.n_tiles_n;
        const char* wtile = wblocks + (size_t)tn * n_kstages * Cfg::B_BYTES;
        uint32_t rem = unit_gmask(st);
        const int n_act = __popc(rem) * spg;
        for (int ia = 0, sub = 0; ia < n_act; ++ia) {
          int ks;
          if (tall) {        // order of the tall-stage MMA loop: horizontal tap, chunk, vertical tap
            const int kx = ia / (3 * spo), r3 = ia - kx * 3 * spo, sb = r3 / 3, ky = r3 - sb * 3;
            ks = (ky * 3 + kx) * spo + sb;
          } else {
            ks = (__ffs((int)rem) - 1) * spg + sub;
            if (++sub == spg) { sub = 0; rem &= rem - 1; }
          }
          const uint32_t bbar = smem_u32(&b_full[b_slot]);
          if (!(t.dbg & 128)) mbar_spin(smem_u32(&b_empty[b_slot]), b_phase ^ 1, 1); else mbar_wait(smem_u32(&b_empty[b_slot]), b_phase ^ 1, 1);
          if (!(t.dbg & 4)) {
            mbar_expect_tx(bbar, Cfg::B_BYTES);
            bulk_g2s(b_ring_u32 + b_slot * Cfg::B_BYTES, wtile + (size_t)ks * Cfg::B_BYTES, Cfg::B_BYTES, bbar);
          }
          mbar_arrive(bbar);
          if (++b_slot == SB) { b_slot = 0; b_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================== EPILOGUE =====================================
    constexpr int CH = NT >= 32 ? 32 : 16;                 // accumulator columns per TMEM load
    const int q = warp & 3;                                 // TMEM lane quarter this warp may access
    int it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int st = unit / t.n_tiles_n, tn = unit - st * t.n_tiles_n;
      const int live = min(T, n_tiles_m - st * T);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int col0 = tn * NT;
      // sorted rulebook (fd_rulebook_sort_rows): tile position -> output row.  The lookup is a dependent global load in
      // front of everything a tile's epilogue does, so it is issued one tile ahead (the first one before the wait for the
      // accumulators): sorted narrow layers are epilogue bound.
      auto perm_row = [&](int ti) -> int {
        const int idx = (st * T + ti) * TC_BM + q * 32 + lane;
        return idx < n ? __ldg(a.row_perm + idx) : -1;
      };
      int o_next = a.row_perm ? perm_row(0) : 0;
      mbar_wait(smem_u32(&t_full[acc]), acc_phase, 6, 64);
      if (threadIdx.x == (TC_MMA_WARP + 1) * 32) TC_TRACE(2, 2 * it, clock64());
      tc_fence_after();
      for (int ti = 0; ti < live; ++ti) {
        int o = tiled ? tile_row_to_o(st * T + ti, q * 32 + lane) : (st * T + ti) * TC_BM + q * 32 + lane;
        if (a.row_perm) {
          o = o_next;
          if (ti + 1 < live) o_next = perm_row(ti + 1);
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * T + ti) * ACC);
        const bool live_row = o >= 0 && o < n;
        OutRow orow{nullptr, 0, 1};
        if (live_row) orow = map_out_row(a, o);
        const float* res = (a.residual && live_row) ? a.residual + (size_t)o * a.res_stride : nullptr;
        // coalesced row I/O: output row this lane serves in store / load instruction k (rows k * RPI + lane / NCH of the warp)
        constexpr int NCHq = CH / 8, RPIq = 32 / NCHq;
        int orr_k[NCHq];
#pragma unroll
        for (int k = 0; k < NCHq; ++k) {
          const int rr = k * RPIq + lane / NCHq;
          // (the common instantiation keeps its round-2 arithmetic: at the 128-register ceiling any change to this code moves
          // ptxas's spills into the gather producers -- taking the row from lane rr everywhere cost the sparse family 8 %)
          const int v = (TALL || a.row_perm) ? __shfl_sync(0xffffffffu, o, rr)
                        : tiled    ? tile_row_to_o(st * T + ti, q * 32 + rr) : (st * T + ti) * TC_BM + q * 32 + rr;
          orr_k[k] = (v >= 0 && v < n) ? v : -1;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += CH) {
          const int cbase = col0 + c0;
          const bool work = live_row && !(t.dbg & 8) && cbase < a.cout;
          const bool fullc = cbase + CH <= a.cout;
          // identity (residual) values first: they do not depend on the accumulator, so their latency overlaps
          // the TMEM load
          float y[CH];
          // (the residual rows are read one row per lane: a coalesced, shared-memory staged version of these loads measured
          // slower -- 32 more registers in the hottest part of the epilogue)
          if (work && res) {
            if (a.res_fmt == FD_FMT_SPLIT_BF16 && fullc) {
              const uint4* rh = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(res) + cbase);
              const uint4* rl = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(res) + a.res_ctot + cbase);
              uint4 hv[CH / 8], lv[CH / 8];
#pragma unroll
              for (int h = 0; h < CH / 8; ++h) { hv[h] = __ldg(rh + h); lv[h] = __ldg(rl + h); }
#pragma unroll
              for (int h = 0; h < CH / 8; ++h) {
                const uint32_t hw[4] = {hv[h].x, hv[h].y, hv[h].z, hv[h].w}, lw[4] = {lv[h].x, lv[h].y, lv[h].z, lv[h].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  y[h * 8 + 2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
                  y[h * 8 + 2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < CH; ++i) y[i] = (cbase + i < a.cout) ? load_residual(a, o, cbase + i) : 0.f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) y[i] = 0.f;
          }
          uint32_t r[CH];
          tmem_ld<CH>(taddr + c0, r);
          if (Cfg::FUSE_N && t.split) {                      // second accumulator block: the A_hi*B_lo products
            uint32_t r2[CH];
            tmem_ld<CH>(taddr + NT + c0, r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < CH; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
          }
          tmem_ld_wait();
          if (ti == live - 1 && c0 + CH >= NT) {              // all TMEM reads of this unit are done
            tc_fence_before();
            mbar_arrive(smem_u32(&t_empty[acc]));
          }
          // Coalesced row output (warp-uniform condition): identity rows in split-bf16 format.  With one row per lane a
          // warp-level 16-byte store touches 32 different rows = 32 L2 requests; staged through 2 KB of shared memory per
          // warp, 32 / NCH rows x NCH chunks go out per instruction (8 requests for a 32-column chunk) -- the epilogue's
          // requests were ~a third of all LSU traffic of a tile and removing its stores gained 16 % (FD_TC_DEBUG & 8).
          const bool co_out = !(t.dbg & (8 | 8192)) && cbase < a.cout && fullc && a.out_map == FD_OUTMAP_IDENTITY &&
                              a.out_fmt == FD_FMT_SPLIT_BF16 && (a.out_ctot & 7) == 0 && (a.out_stride & 3) == 0 &&
                              ((((uintptr_t)a.out) + 2 * (size_t)cbase) & 15) == 0;
          if (!work && !co_out) continue;
          if (ss_smem) {
            // folded scale / shift from shared memory, four channels per load (the epilogue's instruction stream competes with
            // the producer and MMA warps for issue slots: FD_TC_DEBUG & 8 -- no epilogue math -- is worth 14 %)
#pragma unroll
            for (int i4 = 0; i4 < CH / 4; ++i4) {
              const float4 sc = *reinterpret_cast<const float4*>(s_scale + cbase + 4 * i4);
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + cbase + 4 * i4);
              const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float v = fmaf(__uint_as_float(r[4 * i4 + j]), scv[j], shv[j]) + y[4 * i4 + j];
                if (a.relu) v = fmaxf(v, 0.f);
                y[4 * i4 + j] = v;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int c = cbase + i;
              const float sc = (a.scale && c < a.cout) ? __ldg(a.scale + c) : 1.f, sh = (a.shift && c < a.cout) ? __ldg(a.shift + c) : 0.f;
              float v = fmaf(__uint_as_float(r[i]), sc, sh) + y[i];
              if (a.relu) v = fmaxf(v, 0.f);
              y[i] = v;
            }
          }
          if (co_out) {
            constexpr int NCH = CH / 8;                        // 16-byte chunks of a row's CH bf16 columns (per plane)
            constexpr int RPI = 32 / NCH;                      // rows per store instruction
            const uint32_t sbuf = smem_u32(s_epi) + (uint32_t)q * TC_EPI_WARP_BYTES;
#pragma unroll
            for (int plane = 0; plane < 2; ++plane) {
#pragma unroll
              for (int g8 = 0; g8 < NCH; ++g8) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float f0 = y[g8 * 8 + 2 * i], f1 = y[g8 * 8 + 2 * i + 1];
                  __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
                  if (plane == 0) {
                    w[i] = *reinterpret_cast<uint32_t*>(&h);
                  } else {
                    float2 hf = __bfloat1622float2(h);
                    __nv_bfloat162 l = __floats2bfloat162_rn(f0 - hf.x, f1 - hf.y);
                    w[i] = *reinterpret_cast<uint32_t*>(&l);
                  }
                }
                sts_u128(sbuf + (uint32_t)lane * (NCH * 16) + (uint32_t)((g8 ^ (lane & (NCH - 1))) << 4), w[0], w[1], w[2], w[3]);
              }
              __syncwarp();
#pragma unroll
              for (int k = 0; k < NCH; ++k) {
                const int rr = k * RPI + lane / NCH, jj = lane % NCH;
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                             : "r"(sbuf + (uint32_t)rr * (NCH * 16) + (uint32_t)((jj ^ (rr & (NCH - 1))) << 4)) : "memory");
                const int orr = orr_k[k];
                if (orr >= 0) {
                  unsigned short* ob = reinterpret_cast<unsigned short*>(a.out + (size_t)orr * a.out_stride) + cbase +
                                       (plane ? a.out_ctot : 0) + jj * 8;
                  *reinterpret_cast<uint4*>(ob) = make_uint4(w0, w1, w2, w3);
                }
              }
              __syncwarp();
            }
          } else if (orow.cstride == 1 && fullc && a.out_fmt == FD_FMT_SPLIT_BF16 &&
              ((((uintptr_t)orow.base) + 2 * (size_t)(orow.coff + cbase)) & 15) == 0 && (a.out_ctot & 7) == 0) {
            // split once here so that every consumer layer gathers ready-made bf16 hi/lo planes
            unsigned short* ob = reinterpret_cast<unsigned short*>(orow.base) + orow.coff + cbase;
            uint4* dh = reinterpret_cast<uint4*>(ob);
            uint4* dl = reinterpret_cast<uint4*>(ob + a.out_ctot);
#pragma unroll
            for (int g8 = 0; g8 < CH / 8; ++g8) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(y[g8 * 8 + 2 * i], y[g8 * 8 + 2 * i + 1]);
                float2 hf = __bfloat1622float2(h);
                __nv_bfloat162 l = __floats2bfloat162_rn(y[g8 * 8 + 2 * i] - hf.x, y[g8 * 8 + 2 * i + 1] - hf.y);
                hi[i] = *reinterpret_cast<uint32_t*>(&h);
                lo[i] = *reinterpret_cast<uint32_t*>(&l);
              }
              dh[g8] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              dl[g8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          } else if (orow.cstride == 1 && fullc && a.out_fmt == FD_FMT_FP32 &&
                     ((((uintptr_t)(orow.base + orow.coff + cbase)) & 15) == 0)) {
            float4* dst = reinterpret_cast<float4*>(orow.base + orow.coff + cbase);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (cbase + i < a.cout) store_out(a, orow, cbase + i, y[i]);
          }
        }
      }
      if (threadIdx.x == (TC_MMA_WARP + 1) * 32) TC_TRACE(2, 2 * it + 1, clock64());
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library must load (and export its symbols)
// on a machine without libcuda.so.1 -- the CPU-only build / ABI checks
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int pick_nt(int cout) { return cout >= 128 ? 128 : cout >= 64 ? 64 : cout >= 32 ? 32 : 16; }
static int pad_to(int v, int m) { return (v + m - 1) / m * m; }

// perf-triage knobs, settable at run time through fd_debug_set_tc (not part of the documented ABI)
static int g_dbg = -1;          // FD_TC_DEBUG bits
static int g_tma = 0;              // TMA gather4 producer where the layer allows it (measured slower than the cp.async
                                   // gather: ~6 cycles per 128-byte row in the copy engine vs ~4 through the LSU)
static int g_tma_dense = 1;        // TMA tile loads for dense stride-1 2-D convolutions
static int g_tall = 1;             // ... as tall stages (one load per horizontal tap feeds the three vertical taps) for 3x3 kernels

template <int NT>
static int launch_tc(TcArgs& t, cudaStream_t stream) {
  using Cfg = TcCfg<NT>;
  static bool configured = false;
  const size_t smem = t.tma == 3 && Cfg::SMEM_TALL > Cfg::SMEM ? Cfg::SMEM_TALL : Cfg::SMEM;
  if (!configured) {
    FD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    FD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(Cfg::SMEM_TALL > Cfg::SMEM ? Cfg::SMEM_TALL : Cfg::SMEM)));
    configured = true;
  }
  // weight reuse factor: as many M tiles per weight fetch as TMEM allows while keeping >= one unit per SM (a cost model
  // that also counted the wave tail picked smaller T and measured slower: the extra weight traffic outweighs the tail)
  const int tiles_m = t.tma >= 2 ? (t.c.n_cap / (t.c.Hout * t.c.Wout)) * t.tiles_x * t.tiles_y : ceil_div(t.c.n_cap, TC_BM);
  int T = t.tma == 3 && Cfg::TMAX > TC_TALL_T ? TC_TALL_T : Cfg::TMAX;
  while (T > 1 && (int64_t)ceil_div(tiles_m, T) * t.n_tiles_n < kNumSMs) T >>= 1;
  t.T = T;
  const int64_t units = (int64_t)ceil_div(tiles_m, T) * t.n_tiles_n;
  const int grid = units < kNumSMs ? (int)units : kNumSMs;   // persistent: one CTA per SM
  if (t.tma == 3) conv_tc_kernel<NT, true><<<grid, TC_THREADS, smem, stream>>>(t);
  else conv_tc_kernel<NT, false><<<grid, TC_THREADS, smem, stream>>>(t);
  FD_LAUNCHED();
  return 0;
}

int conv_forward_tc(const ConvArgs& a, int precision, cudaStream_t stream) {
  if (a.n_cap <= 0) return 0;
  FD_REQUIRE(a.wp != nullptr, "fd_conv_forward: tensor-core precision needs d_w_packed (fd_conv_pack_weights)");
  FD_REQUIRE(a.cin % 8 == 0 && (a.cin % TC_BK == 0 || TC_BK % a.cin == 0),
             "fd_conv_forward: tensor-core arm needs Cin in {8,16} or a multiple of %d (got %d)", TC_BK, a.cin);
  FD_REQUIRE(a.K <= TC_MAXK, "fd_conv_forward: tensor-core arm supports at most %d kernel offsets", TC_MAXK);
  FD_REQUIRE(a.mode == FD_GATHER_TABLE || a.K <= TC_DENSE_MAXK,
             "fd_conv_forward: tensor-core arm supports dense 2-D kernels of at most %d taps", TC_DENSE_MAXK);
  FD_REQUIRE(a.in_stride % 4 == 0 && (((uintptr_t)a.in) & 15) == 0 && (a.in_fmt != FD_FMT_SPLIT_BF16 || a.in_ctot % 8 == 0),
             "fd_conv_forward: tensor-core arm needs 16-byte aligned input rows / planes");
  FD_REQUIRE((((uintptr_t)a.wp) & 15) == 0, "fd_conv_forward: d_w_packed must be 16-byte aligned");
  TcArgs t{};
  t.c = a;
  t.wp = (const __nv_bfloat16*)a.wp;
  t.tile_mask = a.mode == FD_GATHER_TABLE ? a.tile_mask : nullptr;
  const int NT = pick_nt(a.cout);
  t.cout_pad = pad_to(a.cout, NT);
  t.ktot_pad = pad_to(a.K * a.cin, TC_BK);
  t.n_tiles_n = t.cout_pad / NT;
  t.split = precision == FD_PREC_BF16X3;
  if (g_dbg < 0) {
    g_dbg = getenv("FD_TC_DEBUG") ? atoi(getenv("FD_TC_DEBUG")) : 0;
    if (getenv("FD_TC_TALL")) g_tall = atoi(getenv("FD_TC_TALL"));
  }
  t.dbg = g_dbg;
  // TMA gather for the A operand: wide layers (Cin a multiple of the 64-channel stage) reading split-bf16 rows
  t.tma = 0;
  t.nbr_vec = a.mode == FD_GATHER_TABLE && a.nbr_stride % 4 == 0 && (((uintptr_t)a.nbr) & 15) == 0;
  if (g_tma && a.in_fmt == FD_FMT_SPLIT_BF16 && a.cin % TC_BK == 0 && TC_BK == 64 &&
      (a.mode == FD_GATHER_TABLE || a.mode == FD_GATHER_CONV2D || a.mode == FD_GATHER_CONV2D_DGRAD) &&
      ((size_t)a.in_stride * 4) % 16 == 0) {
    // 2-D view: dim0 = bf16 columns of one row (hi plane, then the lo plane in_ctot further), dim1 = rows.  The row
    // count only bounds the zero-filled "no neighbour" index -1: valid indices always point at real rows.
    const cuuint64_t dims[2] = {(cuuint64_t)a.in_ctot + (cuuint64_t)a.cin, (cuuint64_t)1 << 30};
    const cuuint64_t strides[1] = {(cuuint64_t)a.in_stride * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, 1};
    const cuuint32_t estr[2] = {1, 1};
    EncodeTiledFn enc = encode_tiled_fn();
    FD_REQUIRE(enc != nullptr, "fd_conv_forward: the CUDA driver does not export cuTensorMapEncodeTiled");
    CUresult cr = enc(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)a.in, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS)
      return set_error((int)cr, "fd_conv_forward: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    t.tma = 1;
  }
  // dense stride-1 convolutions: TMA tile loads over the [B,H,W,C] activation tensor
  if (g_tma_dense && a.in_fmt == FD_FMT_SPLIT_BF16 && a.cin % TC_BK == 0 && TC_BK == 64 && a.mode == FD_GATHER_CONV2D &&
      a.sh == 1 && a.sw == 1 && ((size_t)a.in_stride * 4) % 16 == 0 && a.n_cap % (a.Hout * a.Wout) == 0) {
    // patch shape bw x bh <= 128 rows with the fewest wasted accumulator rows
    int bw = 1, bh = 1;
    double best_util = -1.0;
    for (int w = 1; w <= TC_BM && w <= a.Wout; ++w) {
      int h = TC_BM / w;
      if (h > a.Hout) h = a.Hout;
      if (h < 1 || h > 256) continue;
      const double util = (double)a.Hout * a.Wout / ((double)ceil_div(a.Wout, w) * ceil_div(a.Hout, h) * TC_BM);
      if (util > best_util + 1e-9) { best_util = util; bw = w; bh = h; }
    }
    const int nb = a.n_cap / (a.Hout * a.Wout);
    const cuuint64_t dims[4] = {(cuuint64_t)a.in_ctot + (cuuint64_t)a.cin, (cuuint64_t)a.Win, (cuuint64_t)a.Hin, (cuuint64_t)nb};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_stride * 4, (cuuint64_t)a.in_stride * 4 * a.Win,
                                   (cuuint64_t)a.in_stride * 4 * a.Win * a.Hin};
    const cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    EncodeTiledFn enc = encode_tiled_fn();
    FD_REQUIRE(enc != nullptr, "fd_conv_forward: the CUDA driver does not export cuTensorMapEncodeTiled");
    CUresult cr = enc(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)a.in, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS)
      return set_error((int)cr, "fd_conv_forward: cuTensorMapEncodeTiled (4-D) failed (%d)", (int)cr);
    t.tma = 2;
    t.bw = bw; t.bh = bh;
    t.tiles_x = ceil_div(a.Wout, bw); t.tiles_y = ceil_div(a.Hout, bh);
    if (g_tall && a.kh == 3 && a.kw == 3 && a.Wout >= TC_TALL_BW && a.Hout >= TC_TALL_BH) {
      // tall stages: 8 x 16-pixel tiles, box {64 ch, 8 px, 18 lines}
      const cuuint32_t tbox[4] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_TALL_BW, (cuuint32_t)(TC_TALL_BH + 2), 1};
      cr = enc(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)a.in, dims, strides, tbox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS)
        return set_error((int)cr, "fd_conv_forward: cuTensorMapEncodeTiled (tall box) failed (%d)", (int)cr);
      t.tma = 3;
      t.bw = TC_TALL_BW; t.bh = TC_TALL_BH;
      t.tiles_x = ceil_div(a.Wout, t.bw); t.tiles_y = ceil_div(a.Hout, t.bh);
    }
  }
  switch (NT) {
    case 128: return launch_tc<128>(t, stream);
    case 64: return launch_tc<64>(t, stream);
    case 32: return launch_tc<32>(t, stream);
    default: return launch_tc<16>(t, stream);
  }
}

}  // namespace fd

extern "C" {

/* perf-triage helper (not part of the documented ABI): copy the FD_TC_DEBUG&32 trace of block 0 to the host */
int fd_debug_read_tc_trace(long long* out, int role) {
  if (role < 0 || role >= 4) return -1;
  return (int)cudaMemcpyFromSymbol(out, fd::g_tc_trace, sizeof(long long) * fd::TC_TRACE_N,
                                   sizeof(long long) * fd::TC_TRACE_N * role, cudaMemcpyDeviceToHost);
}

/* perf-triage helper (not part of the documented ABI): key 0 = FD_TC_DEBUG bits, 5 = TMA gather4 producer on/off, 6 = TMA tile loads for dense convs on/off, 7 = mbarrier watchdog, 8 = tall stages for dense 3x3 convs on/off
 * (the ring-size and L1-gather knobs of the round-2 triage were removed again: run-time ring sizes cost the narrow
 * layers 10 %, gathers through L1 gained nothing) */
int fd_debug_set_tc(int key, int value) {
  switch (key) {
    case 0: fd::g_dbg = value; return 0;
    case 5: fd::g_tma = value; return 0;
    case 6: fd::g_tma_dense = value; return 0;
    case 8: fd::g_tall = value; return 0;
    case 7: {                                     // watchdog of the mbarrier waits, in units of 2^30 cycles (0: ~never)
      const long long v = value > 0 ? (long long)value << 30 : (1LL << 62);
      return (int)cudaMemcpyToSymbol(fd::g_tc_timeout, &v, sizeof(v));
    }
  }
  return -1;
}

size_t fd_conv_packed_bytes(int K, int cin, int cout) {
  if (K < 1 || cin < 1 || cout < 1) return 0;
  const int NT = fd::pick_nt(cout);
  return (size_t)2 * fd::pad_to(cout, NT) * fd::pad_to(K * cin, fd::TC_BK) * sizeof(__nv_bfloat16);
}

int fd_conv_pack_weights(const float* d_w, int K, int cin, int cout, void* d_packed, void* stream) {
  using namespace fd;
  FD_REQUIRE(d_w && d_packed && K >= 1 && cin >= 1 && cout >= 1, "fd_conv_pack_weights: bad argument");
  const int NT = pick_nt(cout);
  const int cout_pad = pad_to(cout, NT), ktot_pad = pad_to(K * cin, TC_BK);
  pack_weights_kernel<<<persistent_grid(ceil_div((int64_t)cout_pad * ktot_pad, 256), 8), 256, 0, (cudaStream_t)stream>>>(
      d_w, K, cin, cout, NT, cout_pad / NT, ktot_pad / TC_BK, (__nv_bfloat16*)d_packed);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
