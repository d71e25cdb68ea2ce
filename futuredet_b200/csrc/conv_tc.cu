// tcgen05 / TMEM gather -> implicit-GEMM convolution for sm_100a (FD_PREC_BF16X3, FD_PREC_BF16).
//
//   out[o, :] = act( (sum_k in[nbr(o,k), :] @ W[k]) * scale + shift (+ residual[o, :]) )
//
// GEMM view: M = output rows (active sites / pixels), N = Cout, K = (kernel offsets x Cin) flattened.
// One CTA owns a 128-row x NT-column output tile; the fp32 accumulator lives in TMEM (double buffered),
// operands are staged in shared memory in the UMMA canonical K-major SWIZZLE_128B layout
// (64 bf16 = 128 B per row, 16-byte chunk index XOR (row & 7)), 3-4 stage mbarrier ring.
//
// Warp roles (288 threads):
//   warps 0-3  producers : gather 128 input rows per stage through the rulebook (or the dense 2-D index
//                          arithmetic), split every fp32 into bf16 hi + bf16 lo on the fly and store both
//                          planes swizzled; cp.async the pre-packed bf16 hi/lo weight tile; kernel offsets
//                          with no neighbour in the whole tile are skipped.
//   warp  4    MMA issuer: one lane issues tcgen05.mma (M=128, N=NT, K=16, kind::f16, bf16 x bf16 -> fp32):
//                          D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi   (3-term split, ~2^-16 relative error,
//                          i.e. fp32-class results on the bf16 tensor pipe); tcgen05.commit releases stages.
//   warps 5-8  epilogue  : tcgen05.ld the accumulator, fused BN(eval)/bias, residual, ReLU, mapped store.
//
// fp32 activations stay fp32 in HBM (the reference's interface); precision is a property of the kernel.
#include <cuda_bf16.h>

#include "conv_common.cuh"

namespace fd {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                 // bf16 elements per stage row (128 bytes, one swizzle atom)
constexpr int TC_MAXK = 32;               // max kernel offsets (27 for 3x3x3)
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_PRODUCERS = TC_PRODUCER_WARPS * 32;
constexpr int TC_THREADS = TC_PRODUCERS + 32 + 128;   // producer warps + 1 MMA warp + 4 epilogue warps
constexpr int TC_A_PLANE = TC_BM * 128;   // bytes of one A plane (hi or lo) per stage

__host__ __device__ constexpr int tc_stage_bytes(int NT) { return 2 * TC_A_PLANE + 2 * NT * 128; }
__host__ __device__ constexpr int tc_num_stages(int NT) { return NT >= 128 ? 3 : 4; }

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
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

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
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
// the shared-memory B stage: [2 planes (hi, lo)][NT rows][64 bf16, 16-byte chunks XOR-swizzled by (row & 7)].
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
    const int off = nrow * TC_BK + ((((kc >> 3) ^ (nrow & 7)) << 3) | (kc & 7));
    out[blk + off] = hi;
    out[blk + (long long)NT * TC_BK + off] = lo;
  }
}

struct TcArgs {
  ConvArgs c;
  const __nv_bfloat16* wp;   // packed weights
  int cout_pad, ktot_pad;
  int n_tiles_n;
  int split;                 // 1: bf16x3, 0: single-pass bf16
};

// ---- the kernel -----------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const TcArgs t) {
  constexpr int STAGES = tc_num_stages(NT);
  constexpr int STAGE_BYTES = tc_stage_bytes(NT);
  constexpr int TMEM_COLS = (2 * NT) < 32 ? 32 : 2 * NT;
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, NT);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B wants 1024-B alignment
  uint8_t* stage_base = smem;
  int* s_nbr = (int*)(smem + STAGES * STAGE_BYTES);                 // [2][TC_MAXK][128]
  uint64_t* bars = (uint64_t*)(s_nbr + 2 * TC_MAXK * TC_BM);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;         // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;    // [2]
  uint32_t* s_meta = (uint32_t*)(bars + 2 * STAGES + 4);   // [STAGES] bit0 first, bit1 last
  uint32_t* s_tmem = s_meta + STAGES;                      // [1]
  uint32_t* s_active = s_tmem + 1;                         // [2]

  const ConvArgs& a = t.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.d_n ? min(*a.d_n, a.n_cap) : a.n_cap;
  const int n_tiles_m = (n + TC_BM - 1) / TC_BM;
  const int n_tiles = n_tiles_m * t.n_tiles_n;
  const int ktot = a.K * a.cin;
  const int n_kstages = (ktot + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), TC_PRODUCERS / (NT >= 128 ? 2 : 4));   // one producer group per stage
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 128);
    }
    fence_barrier_init();
  }
  if (warp == TC_PRODUCER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < TC_PRODUCER_WARPS) {
    // ===================================== PRODUCERS =====================================
    // The producer warps are split into G groups; group g fills every G-th emitted stage on its own, so G
    // stages are being issued concurrently (a single warp's issue chain per stage is several hundred cycles).
    constexpr int G = NT >= 128 ? 2 : 4;
    constexpr int GT = TC_PRODUCERS / G;         // threads per group
    constexpr int ROWS_PER_PASS = GT / 8;
    constexpr int PASSES = TC_BM / ROWS_PER_PASS;
    constexpr uint32_t B_BYTES = 2 * NT * 128;   // one pre-swizzled weight stage (hi + lo planes)
    const int tid = threadIdx.x;                 // 0..TC_PRODUCERS-1
    const int grp = tid / GT, gt = tid % GT;
    const int j = gt & 7, rbase = gt >> 3;       // 16-byte chunk column, first row (rows rbase + p*ROWS_PER_PASS)
    uint32_t slot = 0, phase = 0, turn = 0;      // ring slot / phase / owning group of the next emitted stage
    const uint32_t stage_u32 = smem_u32(stage_base);
    const uint32_t nbr_u32 = smem_u32(s_nbr);
    // per-thread constant of the swizzled A stores ((row & 7) is the same for every pass of a thread)
    const uint32_t a_off0 = rbase * 128 + ((j ^ (rbase & 7)) << 4);
    // flattened-K bookkeeping without per-stage divisions: Cin is a multiple of 64, or divides 64
    const bool wide = a.cin >= TC_BK;
    const int spo = wide ? a.cin / TC_BK : 1;    // stages per kernel offset   (wide)
    const int opk = wide ? 1 : TC_BK / a.cin;    // kernel offsets per stage   (narrow)
    const int jk = wide ? 0 : (j * 8) / a.cin;   // which of the stage's offsets this thread's chunk belongs to
    const int jc = wide ? j * 8 : (j * 8) % a.cin;
    const uint32_t row_bytes = (uint32_t)a.in_stride * 4;
    const char* in_b = reinterpret_cast<const char*>(a.in);
    const char* in_lo = in_b + (size_t)a.in_ctot * 2;
    const char* wblocks = reinterpret_cast<const char*>(t.wp);
    // rulebook rows of one tile -> s_nbr[buf] (cp.async for the table mode, arithmetic for dense 2-D)
    auto load_nbr = [&](int tile, int buf) {
      const int tm = tile / t.n_tiles_n;
      for (int e = tid; e < a.K * TC_BM; e += TC_PRODUCERS) {
        const int k = e >> 7, r = e & (TC_BM - 1);
        const int o = tm * TC_BM + r;
        const uint32_t dst = nbr_u32 + (uint32_t)(buf * TC_MAXK * TC_BM + e) * 4;
        if (a.mode == FD_GATHER_TABLE && o < n) cp_async4(dst, a.nbr + (size_t)k * a.nbr_stride + o);
        else sts_u32(dst, (uint32_t)(o < n ? gather_row(a, o, k) : -1));
      }
    };
    int it = 0;
    if ((int)blockIdx.x < n_tiles) load_nbr(blockIdx.x, 0);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int tn = tile % t.n_tiles_n;
      const uint32_t nb_u32 = nbr_u32 + (uint32_t)(buf * TC_MAXK * TC_BM) * 4;
      cp_async_wait_group<0>();                            // rulebook rows of this tile (committed a whole tile ago;
                                                           // stage gathers are never committed, so not waited for)
      if (tid == 0) s_active[buf] = 0;
      named_bar_sync(1, TC_PRODUCERS);                     // s_nbr[buf] complete, s_active cleared
      {
        uint32_t m = 0;
        for (int e = tid; e < a.K * TC_BM; e += TC_PRODUCERS)       // k is warp-uniform (32 consecutive rows)
          if (__any_sync(0xffffffffu, (int)lds_u32(nb_u32 + (uint32_t)e * 4) >= 0)) m |= 1u << (e >> 7);
        if (lane == 0 && m) atomicOr(&s_active[buf], m);
      }
      if (tile + (int)gridDim.x < n_tiles) load_nbr(tile + gridDim.x, buf ^ 1);   // prefetch next tile's rows
      cp_async_commit();
      named_bar_sync(1, TC_PRODUCERS);
      const uint32_t active = s_active[buf];
      // stage list: flattened-K stages that touch at least one active kernel offset
      auto stage_mask = [&](int ks) -> uint32_t {
        return wide ? (1u << (ks / spo)) : ((((1u << opk) - 1u) << (ks * opk)));
      };
      int first_ks = -1, last_ks = -1;
      for (int ks = 0; ks < n_kstages; ++ks)
        if (active & stage_mask(ks)) { if (first_ks < 0) first_ks = ks; last_ks = ks; }
      if (first_ks < 0) first_ks = last_ks = 0;            // degenerate tile: one all-zero stage
      const char* wtile = wblocks + (size_t)tn * n_kstages * B_BYTES;
      int kk_run = wide ? first_ks / spo : 0, c_idx = wide ? first_ks - kk_run * spo : 0;
      for (int ks = first_ks; ks <= last_ks; ++ks) {
        const uint32_t smask = wide ? (1u << kk_run) : (((1u << opk) - 1u) << (ks * opk));
        const int kk = wide ? kk_run : ks * opk + jk;
        const int ch = wide ? c_idx * TC_BK + jc : jc;
        if (wide && ++c_idx == spo) { c_idx = 0; ++kk_run; }
        if (!(active & smask) && ks != first_ks && ks != last_ks) continue;
        // every group walks the same emitted-stage sequence; only the owner of this stage fills it
        const uint32_t my_slot = slot, my_phase = phase;
        const bool mine = turn == (uint32_t)grp;
        if (++slot == STAGES) { slot = 0; phase ^= 1; }
        if (++turn == G) turn = 0;
        if (!mine) continue;
        mbar_wait(smem_u32(&empty_bar[my_slot]), my_phase ^ 1);
        const uint32_t sbase = stage_u32 + my_slot * STAGE_BYTES;
        const uint32_t fbar = smem_u32(&full_bar[my_slot]);
        if (gt == 0) {
          s_meta[my_slot] = (ks == first_ks ? 1u : 0u) | (ks == last_ks ? 2u : 0u);
          // ---- B: one TMA bulk copy of the pre-swizzled weight stage, completes (in bytes) on the full barrier
          mbar_expect_tx(fbar, B_BYTES);
          bulk_g2s(sbase + 2 * TC_A_PLANE, wtile + (size_t)ks * B_BYTES, B_BYTES, fbar);
        }
        // ---- A: gather 128 rows x 64 channels (this thread: chunk j of rows rbase + ROWS_PER_PASS p)
        const bool kvalid = kk < a.K;
        const uint32_t nb_row = nb_u32 + (uint32_t)((kvalid ? kk : 0) * TC_BM + rbase) * 4;
        const uint32_t dst0 = sbase + a_off0;
        if (a.in_fmt == FD_FMT_SPLIT_BF16) {
          // pre-split bf16 hi/lo rows: pure 16-byte cp.async copies (zero-filled where the rulebook has no
          // neighbour); nothing is waited for -- the hardware arrives on the full barrier when they land
          const char* gh = in_b + ch * 2;
          const char* gl = in_lo + ch * 2;
#pragma unroll 8
          for (int p = 0; p < PASSES; ++p) {
            const int src = kvalid ? (int)lds_u32(nb_row + p * ROWS_PER_PASS * 4) : -1;
            const uint32_t sz = src >= 0 ? 16u : 0u;
            const size_t goff = (size_t)(uint32_t)max(src, 0) * row_bytes;
            cp_async16_sz(dst0 + p * ROWS_PER_PASS * 128, gh + goff, sz);
            cp_async16_sz(dst0 + p * ROWS_PER_PASS * 128 + TC_A_PLANE, gl + goff, sz);
          }
          cp_async_mbar_arrive_noinc(fbar);
        } else {
          // fp32 rows (the stem reading voxel features): split into bf16 hi/lo in registers, 4 rows at a time
#pragma unroll 1
          for (int p0 = 0; p0 < PASSES; p0 += 4) {
            float4 v[4][2];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const int src = kvalid ? (int)lds_u32(nb_row + (p0 + p) * ROWS_PER_PASS * 4) : -1;
              if (src >= 0) {
                const float4* gp = reinterpret_cast<const float4*>(a.in + (size_t)src * a.in_stride + ch);
                v[p][0] = __ldg(gp);
                v[p][1] = __ldg(gp + 1);
              } else {
                v[p][0] = v[p][1] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float f[8] = {v[p][0].x, v[p][0].y, v[p][0].z, v[p][0].w, v[p][1].x, v[p][1].y, v[p][1].z, v[p][1].w};
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
                float2 hf = __bfloat1622float2(h);
                __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * q] - hf.x, f[2 * q + 1] - hf.y);
                hi[q] = *reinterpret_cast<uint32_t*>(&h);
                lo[q] = *reinterpret_cast<uint32_t*>(&l);
              }
              sts_u128(dst0 + (p0 + p) * ROWS_PER_PASS * 128, hi[0], hi[1], hi[2], hi[3]);
              sts_u128(dst0 + (p0 + p) * ROWS_PER_PASS * 128 + TC_A_PLANE, lo[0], lo[1], lo[2], lo[3]);
            }
          }
          fence_proxy_async();                             // generic-proxy smem writes -> visible to the tensor core
          mbar_arrive(fbar);
        }
      }
    }
    cp_async_wait_all();
  } else if (warp == TC_PRODUCER_WARPS) {
    // ===================================== MMA ISSUER =====================================
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);       // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * NT;
      while (true) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        fence_proxy_async();
        const uint32_t meta = s_meta[stage];
        if (lane == 0) {
          const uint32_t sA_hi = smem_u32(stage_base + stage * STAGE_BYTES);
          const uint32_t sA_lo = sA_hi + TC_A_PLANE;
          const uint32_t sB_hi = sA_lo + TC_A_PLANE;
          const uint32_t sB_lo = sB_hi + NT * 128;
          const uint64_t dA_hi = umma_desc_sw128(sA_hi), dA_lo = umma_desc_sw128(sA_lo);
          const uint64_t dB_hi = umma_desc_sw128(sB_hi), dB_lo = umma_desc_sw128(sB_lo);
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
            const uint64_t adv = (uint64_t)(k4 * 2);       // 16 bf16 = 32 bytes = 2 x 16-byte units
            umma_bf16(tmem_d, dA_hi + adv, dB_hi + adv, IDESC, ((meta & 1u) && k4 == 0) ? 0u : 1u);
            if (t.split) {
              umma_bf16(tmem_d, dA_hi + adv, dB_lo + adv, IDESC, 1u);
              umma_bf16(tmem_d, dA_lo + adv, dB_hi + adv, IDESC, 1u);
            }
          }
          umma_commit(smem_u32(&empty_bar[stage]));        // smem stage reusable once these MMAs retire
          if (meta & 2u) umma_commit(smem_u32(&tfull_bar[acc]));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (meta & 2u) break;
      }
    }
  } else {
    // ===================================== EPILOGUE =====================================
    const int q = warp & 3;                                 // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int tm = tile / t.n_tiles_n, tn = tile - tm * t.n_tiles_n;
      const int o = tm * TC_BM + q * 32 + lane;
      const int col0 = tn * NT;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NT;
      const bool live = o < n;
      OutRow orow{nullptr, 0, 1};
      if (live) orow = map_out_row(a, o);
      const float* res = (a.residual && live) ? a.residual + (size_t)o * a.res_stride : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (c0 + 16 >= NT) {                                // all TMEM reads of this tile are done
          tc_fence_before();
          mbar_arrive(smem_u32(&tempty_bar[acc]));
        }
        if (!live) continue;
        const int cbase = col0 + c0;
        if (cbase >= a.cout) continue;
        const bool full16 = cbase + 16 <= a.cout;
        float y[16];
        // residual (identity) values
        if (res) {
          if (a.res_fmt == FD_FMT_SPLIT_BF16 && full16) {
            const uint4* rh = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(res) + cbase);
            const uint4* rl = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(res) + a.res_ctot + cbase);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint4 a4 = __ldg(rh + h), b4 = __ldg(rl + h);
              const uint32_t hw[4] = {a4.x, a4.y, a4.z, a4.w}, lw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                y[h * 8 + 2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
                y[h * 8 + 2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = (cbase + i < a.cout) ? load_residual(a, o, cbase + i) : 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = cbase + i;
          float v = __uint_as_float(r[i]);
          if (c < a.cout) {
            const float sc = a.scale ? __ldg(a.scale + c) : 1.f;
            const float sh = a.shift ? __ldg(a.shift + c) : 0.f;
            v = fmaf(v, sc, sh) + y[i];
            if (a.relu) v = fmaxf(v, 0.f);
          }
          y[i] = v;
        }
        if (orow.cstride == 1 && full16 && a.out_fmt == FD_FMT_SPLIT_BF16 &&
            ((((uintptr_t)orow.base) + 2 * (size_t)(orow.coff + cbase)) & 15) == 0 && (a.out_ctot & 7) == 0) {
          // split once here so that every consumer layer gathers ready-made bf16 hi/lo planes
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
            float2 hf = __bfloat1622float2(h);
            __nv_bfloat162 l = __floats2bfloat162_rn(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
            hi[i] = *reinterpret_cast<uint32_t*>(&h);
            lo[i] = *reinterpret_cast<uint32_t*>(&l);
          }
          unsigned short* ob = reinterpret_cast<unsigned short*>(orow.base) + orow.coff + cbase;
          uint4* dh = reinterpret_cast<uint4*>(ob);
          uint4* dl = reinterpret_cast<uint4*>(ob + a.out_ctot);
          dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        } else if (orow.cstride == 1 && full16 && a.out_fmt == FD_FMT_FP32 &&
                   ((((uintptr_t)(orow.base + orow.coff + cbase)) & 15) == 0)) {
          float4* dst = reinterpret_cast<float4*>(orow.base + orow.coff + cbase);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cbase + i < a.cout) store_out(a, orow, cbase + i, y[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

static int pick_nt(int cout) { return cout >= 128 ? 128 : cout >= 64 ? 64 : cout >= 32 ? 32 : 16; }
static int pad_to(int v, int m) { return (v + m - 1) / m * m; }

template <int NT>
static int launch_tc(const TcArgs& t, cudaStream_t stream) {
  constexpr int STAGES = tc_num_stages(NT);
  const size_t smem = 1024 + (size_t)STAGES * tc_stage_bytes(NT) + 2 * TC_MAXK * TC_BM * sizeof(int) +
                      (2 * STAGES + 4) * sizeof(uint64_t) + (STAGES + 4) * sizeof(uint32_t);
  static bool configured = false;
  if (!configured) {
    FD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int tiles = ceil_div(t.c.n_cap, TC_BM) * t.n_tiles_n;
  int grid = tiles < kNumSMs ? tiles : kNumSMs;           // persistent: one CTA per SM
  conv_tc_kernel<NT><<<grid, TC_THREADS, smem, stream>>>(t);
  FD_LAUNCHED();
  return 0;
}

int conv_forward_tc(const ConvArgs& a, int precision, cudaStream_t stream) {
  if (a.n_cap <= 0) return 0;
  FD_REQUIRE(a.wp != nullptr, "fd_conv_forward: tensor-core precision needs d_w_packed (fd_conv_pack_weights)");
  FD_REQUIRE(a.cin % 8 == 0 && (a.cin % TC_BK == 0 || TC_BK % a.cin == 0),
             "fd_conv_forward: tensor-core arm needs Cin in {8,16,32} or a multiple of 64 (got %d)", a.cin);
  FD_REQUIRE(a.K <= TC_MAXK, "fd_conv_forward: tensor-core arm supports at most %d kernel offsets", TC_MAXK);
  FD_REQUIRE(a.in_stride % 4 == 0 && (((uintptr_t)a.in) & 15) == 0 && (a.in_fmt != FD_FMT_SPLIT_BF16 || a.in_ctot % 8 == 0),
             "fd_conv_forward: tensor-core arm needs 16-byte aligned input rows / planes");
  TcArgs t{};
  t.c = a;
  t.wp = (const __nv_bfloat16*)a.wp;
  const int NT = pick_nt(a.cout);
  t.cout_pad = pad_to(a.cout, NT);
  t.ktot_pad = pad_to(a.K * a.cin, TC_BK);
  t.n_tiles_n = t.cout_pad / NT;
  FD_REQUIRE((((uintptr_t)a.wp) & 15) == 0, "fd_conv_forward: d_w_packed must be 16-byte aligned");
  t.split = precision == FD_PREC_BF16X3;
  switch (NT) {
    case 128: return launch_tc<128>(t, stream);
    case 64: return launch_tc<64>(t, stream);
    case 32: return launch_tc<32>(t, stream);
    default: return launch_tc<16>(t, stream);
  }
}

}  // namespace fd

extern "C" {

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
