// Convolution weight gradient on tcgen05 / TMEM for sm_100a (FD_PREC_BF16X3 arm of fd_conv_wgrad).
//
//   dW[k][ci][co] += sum over pairs (i, o) of offset k:  x[i][ci] * dy[o][co]
//
// GEMM view per kernel offset: D[M = ci][N = co] = X_k^T [ci x pairs] . dY_k [pairs x co], i.e. the reduction (K)
// dimension is the list of rulebook pairs and BOTH operands arrive "MN-major": a gathered row of x (or dy) is one K
// index holding its channels contiguously.  That is exactly the UMMA canonical MN-major SWIZZLE_128B layout
// (64 channels = one 128-byte row, 8 rows = one swizzle atom), so the gathered rows are stored as they come -- no
// transposition anywhere -- and the instruction descriptor marks A and B as MN-major.
//
// One CTA = one kernel offset x one (<=128 ci) x (<=128 co) tile x one chunk of output rows.  Active pairs are compacted
// on the fly (ballot + ring, as in the CUDA-core arm), 64 pairs form a stage: every thread gathers fp32 rows, splits
// them into bf16 hi/lo planes in registers and stores them swizzled; one thread issues 12 tcgen05.mma per stage
// (K = 16 pairs each; A_hi*B_hi + A_hi*B_lo + A_lo*B_hi: ~2^-16 relative, fp32-class) into a 128-lane x NT-column fp32
// accumulator in TMEM; two stages alternate so the gather of one overlaps the MMAs of the other.  The epilogue reads
// TMEM (lane = ci, column = co).  Several row chunks share a tile: with a `partial` buffer every chunk stores its tile
// into its own slot and wgrad_reduce_kernel adds the slots in ascending order (bit-reproducible gradients); without
// one the chunks add into dW with fp32 atomics (order varies run to run).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "conv_common.cuh"

namespace fd {

constexpr int WOS_TRACE_N = 4096;
__device__ long long g_wos_trace[4][WOS_TRACE_N];       // FD_WG_DBG & 32: clock64 timeline of block (0, 0) (perf triage only)
#define WOS_TRACE(role, i, v) do { if ((pl.dbg & 32) && blockIdx.x == 0 && blockIdx.y == 0 && (i) < WOS_TRACE_N) g_wos_trace[role][i] = (v); } while (0)
__device__ int g_wg_abort = 0;      // raised by the first mbarrier wait that times out: every later wait gives up at once

namespace wg {

constexpr int THREADS = 512;                // 16 warps: the stage is latency bound (gather -> split -> store), not MMA bound
constexpr int KP = 64;                     // pairs per stage
constexpr int ROWB = 128;                  // bytes of one operand row (64 bf16 channels)
constexpr int BLOCK_BYTES = KP * ROWB;     // one 64-channel column block of one plane: 8 KB
constexpr int PLANE_BYTES = 2 * BLOCK_BYTES;   // up to 128 channels
constexpr int OPER_BYTES = 2 * PLANE_BYTES;    // hi + lo
constexpr int STAGE_BYTES = 2 * OPER_BYTES;    // A + B: 64 KB
constexpr int QN = 1024;                   // pair ring: < KP leftovers + one batch of THREADS rows
constexpr size_t SMEM = 1024 + 2 * STAGE_BYTES + 2 * QN * 4 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// bounded wait: a protocol bug must not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; spins < (1u << 24); ++spins) {
    if ((spins & 1023) == 1023 && *(volatile int*)&g_wg_abort) return false;
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
// same, polling with the non-suspending test_wait (a parked try_wait wakes up hundreds of cycles after the phase flips)
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
__device__ __forceinline__ bool mbar_spin(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; spins < (1u << 26); ++spins) {
    if ((spins & 4095) == 4095 && *(volatile int*)&g_wg_abort) return false;
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B: LBO = distance between 64-channel column blocks,
// SBO = distance between 8-row (8 K index) groups = 1024 B, version 1, layout type 2
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
  constexpr uint64_t lbo = BLOCK_BYTES >> 4, sbo = (8 * ROWB) >> 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32, A = B = bf16, A and B MN-major (bits 15, 16), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
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

// 8 fp32 channels -> bf16 hi / lo (4 words each), stored at 16-byte chunk `c16` of row `r` of a plane pair
__device__ __forceinline__ void store_split8(uint32_t plane_hi, int r, int c16, const float4& v0, const float4& v1) {
  const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    const float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    hi[e] = *reinterpret_cast<uint32_t*>(&h);
    lo[e] = *reinterpret_cast<uint32_t*>(&l);
  }
  // column block = c16 / 8, chunk inside the 128-byte row XOR-swizzled by the row's position in its 8-row atom
  const uint32_t off = (uint32_t)(c16 >> 3) * BLOCK_BYTES + (uint32_t)r * ROWB + (uint32_t)(((c16 & 7) ^ (r & 7)) << 4);
  sts_u128(plane_hi + off, hi[0], hi[1], hi[2], hi[3]);
  sts_u128(plane_hi + PLANE_BYTES + off, lo[0], lo[1], lo[2], lo[3]);
}

// same with the swizzled offset precomputed by the caller
__device__ __forceinline__ void store_split8_at(uint32_t addr_hi, const float4& v0, const float4& v1) {
  const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    const float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    hi[e] = *reinterpret_cast<uint32_t*>(&h);
    lo[e] = *reinterpret_cast<uint32_t*>(&l);
  }
  sts_u128(addr_hi, hi[0], hi[1], hi[2], hi[3]);
  sts_u128(addr_hi + PLANE_BYTES, lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace wg

__global__ void __launch_bounds__(wg::THREADS, 1)
conv_wgrad_tc_kernel(const ConvArgs a, float* __restrict__ dw, float* __restrict__ partial, int rows_per_cta, int tiles_ci,
                     int tiles_co) {
  using namespace wg;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_u32 = smem_u32(base);
  int* q_in = reinterpret_cast<int*>(base + 2 * STAGE_BYTES);
  int* q_out = q_in + QN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_out + QN);           // [0], [1]: MMAs of stage s done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
  int* s_wcnt = reinterpret_cast<int*>(s_tmem + 1);                     // THREADS / 32 ints

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.d_n ? min(*a.d_n, a.n_cap) : a.n_cap;
  int t = blockIdx.y;
  const int to = t % tiles_co; t /= tiles_co;
  const int ti = t % tiles_ci;
  const int k = t / tiles_ci;
  const int ci0 = ti * 128, co0 = to * 128;
  const int cin_t = min(128, a.cin - ci0), cout_t = min(128, a.cout - co0);
  const int NT = (cout_t + 15) & ~15;
  // the row chunks are cut from the DEVICE-side row count: capacities of strided levels are several times the number
  // of active sites, and chunks cut from the capacity would leave most CTAs without rows
  rows_per_cta = ((n + (int)gridDim.x - 1) / (int)gridDim.x + 63) & ~63;
  const long long rb = (long long)blockIdx.x * rows_per_cta;
  // deterministic mode: this chunk's slot of the partial buffer, laid out like dW
  float* pslot = partial ? partial + (size_t)blockIdx.x * a.K * a.cin * a.cout : nullptr;
  if (rb >= n) {
    if (pslot)                                        // a chunk without rows still owns (and zeroes) its tile
      for (int e = tid; e < cin_t * cout_t; e += THREADS)
        pslot[((size_t)k * a.cin + ci0 + e / cout_t) * a.cout + co0 + e % cout_t] = 0.f;
    return;
  }
  const int row_begin = (int)rb, row_end = (int)min((long long)n, rb + rows_per_cta);

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t idesc = idesc_mn(128, NT);

  const int a_chunks = cin_t >> 3, b_chunks = cout_t >> 3;   // 16-byte (8-channel) chunks per gathered row
  // item -> (row, chunk) of this thread, fixed for the whole launch (power-of-two chunk counts: shifts, no division in
  // the stage loop -- the loop is instruction-issue bound: ~500 instructions per thread and stage before this)
  constexpr int ITEMS_PRE = KP * 16 / THREADS;
  int a_r[ITEMS_PRE], a_c[ITEMS_PRE], b_r[ITEMS_PRE], b_c[ITEMS_PRE];
  uint32_t a_so[ITEMS_PRE], b_so[ITEMS_PRE];                  // swizzled shared-memory offset of the item inside a plane
#pragma unroll
  for (int i = 0; i < ITEMS_PRE; ++i) {
    const int e = tid + i * THREADS;
    a_r[i] = e / a_chunks; a_c[i] = e - a_r[i] * a_chunks;
    b_r[i] = e / b_chunks; b_c[i] = e - b_r[i] * b_chunks;
    a_so[i] = (uint32_t)(a_c[i] >> 3) * BLOCK_BYTES + (uint32_t)a_r[i] * ROWB + (uint32_t)(((a_c[i] & 7) ^ (a_r[i] & 7)) << 4);
    b_so[i] = (uint32_t)(b_c[i] >> 3) * BLOCK_BYTES + (uint32_t)b_r[i] * ROWB + (uint32_t)(((b_c[i] & 7) ^ (b_r[i] & 7)) << 4);
  }
  int q_head = 0, q_cnt = 0, it = 0;
  bool ok = true;
  for (int rbase = row_begin; rbase < row_end || q_cnt > 0; rbase += THREADS) {
    if (rbase < row_end) {
      const int o = rbase + tid;
      const int src = o < row_end ? gather_row(a, o, k) : -1;
      const unsigned ballot = __ballot_sync(0xffffffffu, src >= 0);
      if (lane == 0) s_wcnt[warp] = __popc(ballot);
      __syncthreads();
      int woff = 0, total = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) {
        const int c = s_wcnt[w];
        if (w < warp) woff += c;
        total += c;
      }
      if (src >= 0) {
        const int pos = (q_head + q_cnt + woff + __popc(ballot & ((1u << lane) - 1u))) & (QN - 1);
        q_in[pos] = src;
        q_out[pos] = o;
      }
      q_cnt += total;
      __syncthreads();
    }
    const bool last = rbase + THREADS >= row_end;
    while (q_cnt >= KP || (last && q_cnt > 0)) {
      const int take = min(q_cnt, KP);
      const int s = it & 1;
      if (it >= 2) ok = mbar_wait(smem_u32(&bars[s]), (uint32_t)(((it >> 1) - 1) & 1)) && ok;   // MMAs that read this buffer are done
      const uint32_t sA = stage_u32 + (uint32_t)s * STAGE_BYTES, sB = sA + OPER_BYTES;
      // ---- gather + split: A rows = x[in], B rows = dy[out]; rows past `take` are zeroed (0 * garbage could be NaN).
      // Every load of the stage is issued before the first conversion so that a thread keeps up to 16 x 16 bytes in
      // flight (the stage is latency bound otherwise: one L2 round trip per item).
      constexpr int ITEMS = KP * 16 / THREADS;               // 4 items of 8 channels per thread and operand at most
      float4 va[ITEMS][2], vb[ITEMS][2];
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int e = tid + i * THREADS;
        const int r = a_r[i], c16 = a_c[i];
        va[i][0] = va[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < KP * a_chunks && r < take) {
          const float4* p = reinterpret_cast<const float4*>(a.in + (size_t)q_in[(q_head + r) & (QN - 1)] * a.in_stride + ci0 + c16 * 8);
          va[i][0] = __ldg(p); va[i][1] = __ldg(p + 1);
        }
      }
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int e = tid + i * THREADS;
        const int r = b_r[i], c16 = b_c[i];
        vb[i][0] = vb[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < KP * b_chunks && r < take) {
          const float4* p = reinterpret_cast<const float4*>(a.out + (size_t)q_out[(q_head + r) & (QN - 1)] * a.out_stride + co0 + c16 * 8);
          vb[i][0] = __ldg(p); vb[i][1] = __ldg(p + 1);
        }
      }
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int e = tid + i * THREADS;
        if (e < KP * a_chunks) store_split8_at(sA + a_so[i], va[i][0], va[i][1]);
      }
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int e = tid + i * THREADS;
        if (e < KP * b_chunks) store_split8_at(sB + b_so[i], vb[i][0], vb[i][1]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dAh = desc_mn_sw128(sA), dAl = desc_mn_sw128(sA + PLANE_BYTES);
        const uint64_t dBh = desc_mn_sw128(sB), dBl = desc_mn_sw128(sB + PLANE_BYTES);
#pragma unroll
        for (int ks = 0; ks < KP / 16; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 16 * ROWB) >> 4);       // 16 K indices = two 8-row atoms further
          umma(tmem_base, dAh + adv, dBh + adv, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          umma(tmem_base, dAh + adv, dBl + adv, idesc, 1u);
          umma(tmem_base, dAl + adv, dBh + adv, idesc, 1u);
        }
        umma_commit(smem_u32(&bars[s]));
      }
      ++it;
      q_head = (q_head + take) & (QN - 1);
      q_cnt -= take;
    }
  }
  // ---- drain: the last commit of each buffer covers every earlier MMA
  if (it > 0) {
    const int last_it = it - 1;
    ok = mbar_wait(smem_u32(&bars[last_it & 1]), (uint32_t)((last_it >> 1) & 1)) && ok;
    if (it > 1) {
      const int prev = it - 2;
      ok = mbar_wait(smem_u32(&bars[prev & 1]), (uint32_t)((prev >> 1) & 1)) && ok;
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && tid == 0) atomicAdd(&g_wg_abort, 1);
  // ---- epilogue: TMEM lane = ci, column = co; warps 0-3 own the four lane quarters
  if (pslot && it == 0) {                              // no pair in this chunk: the slot still has to be defined
    for (int e = tid; e < cin_t * cout_t; e += THREADS)
      pslot[((size_t)k * a.cin + ci0 + e / cout_t) * a.cout + co0 + e % cout_t] = 0.f;
  }
  // all 16 warps: warp w may read TMEM lanes (w % 4) * 32 .. + 31 (its ci rows); the four warps of a lane quarter split
  // the columns
  if (it > 0 && ok) {
    const int lq = warp & 3;
    const int ci = lq * 32 + lane;
    const int cols_per = (((NT + 15) >> 4) + 3) / 4 * 16;             // 16-column groups per column quarter
    const int cbeg = (warp >> 2) * cols_per, cend = min(NT, cbeg + cols_per);
    float* dwk = (pslot ? pslot : dw) + ((size_t)k * a.cin + ci0 + ci) * a.cout + co0;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (ci < cin_t) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float x = __uint_as_float(v[j]);
          if (c0 + j < cout_t) {
            if (pslot) dwk[c0 + j] = x;
            else if (x != 0.f) atomicAdd(dwk + c0 + j, x);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}

// =====================================================================================================================
// Output-stationary weight gradient ("wos", round 2).  The kernel above gives every kernel offset its own CTAs, so dY
// rows are gathered once per offset (27x), every stage pays gather -> split -> store -> barrier serially, and a layer
// with Cin < 128 uses a fraction of the 128 accumulator lanes.  Here a CTA owns a chunk of consecutive OUTPUT rows and
// a "pass" of M groups: per 64-row K stage the dY tile is loaded ONCE (contiguous rows, no gather) and reused by every
// group of the pass; a group's A operand is 128 accumulator lanes = 128 / Cin kernel offsets side by side (Cin < 128:
// the gathered rows of 2 / 4 / 8 offsets share the 128-byte MN-major rows, so narrow layers fill the tensor core's M)
// or one 128-channel slice of one offset; every group has its own NT-column accumulator in TMEM (<= 512 columns per
// pass: 27 offsets of a 64 -> 64 layer are 4 passes of 3-4 groups with the fused split products below, of a 32 -> 32 layer
// one pass).  Operands are split-bf16 hi / lo rows -- the copies the training step's fd_affine_act / fd_bn_backward
// already wrote, else split_rows_kernel makes them (one streaming pass over x and dy) -- so the 8 gather warps (4 groups
// filling 4 items concurrently) are pure cp.async (missing neighbours zero-filled) into a 4-5 slot A ring; a loader
// warp owns the dY ring (2 slots) and a 4-deep ring of [K][64] rulebook slices; one thread issues the 8-12 tcgen05.mma
// of a (group, stage) item, probing the next item's barrier first, and commits the slot back; the accumulators are
// read once at the end of the chunk (lane = (offset, ci), column = co) into the chunk's slot of the partial buffer
// (fixed-tree ordered reduce afterwards: bit-reproducible).  Cout > 128 runs as 128-column slices (grid.z); a
// ConvTranspose2d(k == s) phase maps its output rows to the phase's pixels of dL/dy.
namespace wos {

constexpr int KP = 64, ROWB = 128, BLOCK_BYTES = KP * ROWB;        // one 64-channel column block of one plane: 8 KB
constexpr int A_PLANE = 2 * BLOCK_BYTES, A_BYTES = 2 * A_PLANE;     // 128 lanes x 64 rows, hi + lo: 32 KB
constexpr int SB = 2;
constexpr int PRODUCERS = 256, PGROUPS = 4, THREADS = PRODUCERS + 64;      // + MMA warp + loader warp (dY stages, rulebook slices)
constexpr int IDX_K = 32;                                           // kernel offsets of a staged rulebook slice
constexpr int IDX_SLOTS = 4;                                        // staged rulebook slices (stages) in flight
constexpr int IDX_BYTES = IDX_SLOTS * IDX_K * KP * 4;               // [K][64] input-row indices per stage: 32 KB
// A ring slots: what the dY ring (2 x 32 KB for Cout = 128, 2 x 16 KB below) and the index slices leave of the CTA's
// shared memory (5 / 4 slots)
constexpr size_t smem_bytes(int sa, int b_bytes) { return 1024 + (size_t)sa * A_BYTES + (size_t)SB * b_bytes + IDX_BYTES + 256; }

__device__ __forceinline__ void cp_async16_sz(uint32_t dst, const void* src, uint32_t sz) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ uint32_t wg_lds(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void wg_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// Warp-level wait.  Default: every lane waits on the barrier itself (try_wait parks the lanes; this is also the form
// compute-sanitizer's racecheck can follow -- with one polling lane + a shuffle the other lanes never touch the
// barrier and the tool reports the ring hand-over as hazards).  `one_lane`: lane 0 polls with test_wait, the warp parks
// at the shuffle (triage knob FD_WG_DBG & 8).
__device__ __forceinline__ bool wait_warp(uint32_t bar, uint32_t parity, bool one_lane = false) {
  bool ok = true;
  if (one_lane) {
    if ((threadIdx.x & 31) == 0) ok = wg::mbar_spin(bar, parity);
    return __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
  }
  ok = wg::mbar_wait(bar, parity);
  return __all_sync(0xffffffffu, ok) != 0;
}

struct Plan {
  int opg;          // kernel offsets per group (Cin < 128) or 1
  int ci_tiles;     // 128-channel slices per offset (Cin >= 128) or 1
  int n_groups;     // groups of the whole layer
  int gpp;          // groups per pass
  int passes;
  int NT;           // accumulator columns per group
  int tmem_cols;
  int fuse;         // Cout <= 64: A_hi * [B_hi | B_lo] as ONE MMA of width 2 NT (each MMA re-reads its whole 4 KB A slice from
                    // shared memory -- the port is what bounds an item), two accumulator column blocks per group
  int gcols;        // accumulator columns per group (NT or 2 NT)
  int chunks;
  int co_tiles;     // 128-column slices of Cout (grid.z)
  int b_bytes;      // bytes of one dY ring slot
  int dbg;          // FD_WG_DBG triage bits: 1 no gathers, 2 no MMAs, 4 no proxy fence, 8 one polling lane per gather warp,
                    // 32 clock64 timeline of block (0,0) (tools/wgrad_trace.py)
};

}  // namespace wos

// fp32 rows [n, C] (row stride in floats) -> FD_FMT_SPLIT_BF16 rows [n][C hi | C lo], dense
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, int stride, int C, const int32_t* __restrict__ d_n, int n_cap,
                  unsigned short* __restrict__ out) {
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int c8 = C >> 3;
  const long long total = (long long)n * c8;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / c8), c = (int)(e - (long long)r * c8) * 8;
    const float4* p = reinterpret_cast<const float4*>(x + (size_t)r * stride + c);
    const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
    const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      const float2 hf = __bfloat1622float2(h);
      __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
      hi[i] = *reinterpret_cast<uint32_t*>(&h);
      lo[i] = *reinterpret_cast<uint32_t*>(&l);
    }
    unsigned short* o = out + (size_t)r * 2 * C + c;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + C) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// a.in / a.out: split-bf16 copies of x / dy (dense rows of 2*cin / 2*cout bf16); everything else as the forward conv
template <int SA>
__global__ void __launch_bounds__(wos::THREADS, 1)
conv_wgrad_os_kernel(const ConvArgs a, float* __restrict__ dw, float* __restrict__ partial, const wos::Plan pl) {
  using namespace wos;
  const int B_BYTES = pl.b_bytes;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_ring = wg::smem_u32(base), b_ring = a_ring + SA * A_BYTES;
  const uint32_t idx_ring = b_ring + SB * B_BYTES;                     // [IDX_SLOTS][IDX_K][KP] int32 (rulebook-table layers)
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + SA * A_BYTES + SB * B_BYTES + IDX_BYTES);
  uint64_t* a_full = bars, *a_empty = bars + SA, *b_full = bars + 2 * SA, *b_empty = b_full + SB, *done = b_empty + SB;
  uint64_t* i_full = done + 1, *i_empty = i_full + IDX_SLOTS;          // [IDX_SLOTS] each
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(i_empty + IDX_SLOTS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.d_n ? min(*a.d_n, a.n_cap) : a.n_cap;
  const int pass = blockIdx.y;
  // this pass's groups: the layer's groups spread evenly over the passes (14 groups in 4 passes: 3, 4, 3, 4)
  const int g0 = (pass * pl.n_groups) / pl.passes, G = ((pass + 1) * pl.n_groups) / pl.passes - g0;
  const int rows_per_cta = ((n + (int)gridDim.x - 1) / (int)gridDim.x + KP - 1) & ~(KP - 1);
  const long long rb = (long long)blockIdx.x * rows_per_cta;
  const int row_begin = (int)min((long long)n, rb), row_end = (int)min((long long)n, rb + rows_per_cta);
  const int n_stages = (row_end - row_begin + KP - 1) / KP;
  const int cin = a.cin, cout = a.cout, NT = pl.NT;
  const int co0 = blockIdx.z * 128;                      // this CTA's slice of the output channels (NT columns)
  float* pslot = partial ? partial + (size_t)blockIdx.x * a.K * cin * cout : nullptr;
  // dY row (of the split copy) of output row o: identity, or the phase pixel of a ConvTranspose2d(k == s)
  auto dy_row = [&](int o) -> size_t {
    if (a.out_map == FD_OUTMAP_IDENTITY) return (size_t)o;
    const int hw = a.Hin * a.Win, b = o / hw, r = o - b * hw, y = r / a.Win, x = r - y * a.Win;
    return ((size_t)b * a.Hin * a.up_s + (size_t)y * a.up_s + a.up_dy) * (a.Win * a.up_s) + (size_t)x * a.up_s + a.up_dx;
  };
  // (offset, first channel) of lane-slot `ls` of group `grp`
  auto group_k = [&](int grp, int ls, int& ch) -> int {
    if (cin >= 128) { ch = (grp % pl.ci_tiles) * 128 + ls; return grp / pl.ci_tiles; }
    ch = ls % cin;
    return grp * pl.opg + ls / cin;
  };

  if (n_stages == 0) {                                   // a chunk without rows still owns (and zeroes) its tiles
    if (pslot)
      for (int g = 0; g < G; ++g)
        for (int e = tid; e < 128 * NT; e += THREADS) {
          int ch; const int k = group_k(g0 + g, e / NT, ch);
          if (k < a.K) pslot[((size_t)k * cin + ch) * cout + co0 + e % NT] = 0.f;
        }
    return;
  }

  if (tid == 0) {
    for (int s = 0; s < SA; ++s) { wg::mbar_init(wg::smem_u32(&a_full[s]), PRODUCERS / PGROUPS); wg::mbar_init(wg::smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < SB; ++s) { wg::mbar_init(wg::smem_u32(&b_full[s]), 32); wg::mbar_init(wg::smem_u32(&b_empty[s]), 1); }
    wg::mbar_init(wg::smem_u32(done), 1);
    for (int s = 0; s < IDX_SLOTS; ++s) { wg::mbar_init(wg::smem_u32(&i_full[s]), 32); wg::mbar_init(wg::smem_u32(&i_empty[s]), PRODUCERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PRODUCERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wg::smem_u32(s_tmem)), "r"(pl.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  bool ok = true;

  if (warp < PRODUCERS / 32) {
    // ===================================== PRODUCERS =====================================
    // 4 groups of 64 threads; group p fills items p, p + 4, ... ((stage, group) pairs in MMA order) on its own, so four
    // items are being issued concurrently: a single gather warp's dependent instruction stream (index reads, address
    // arithmetic, 32 cp.async, barrier hand-shake) takes ~1500 cycles per item even with nothing to copy (clock64
    // timeline, tools/wgrad_trace.py).  Thread: 16-byte chunk c of the 128-lane A row, rows r0 + 4 q.
    // Rulebook-table layers read their source rows from the [K][64] slice of the neighbour table that the loader warp
    // stages in shared memory; dense 2-D layers compute them arithmetically.  The gather warps issue NOTHING but the A
    // gathers: cp.async completes in order per thread, so a dY / index load in the same thread's stream would put its
    // DRAM round trip in front of the next A stage's arrival.
    const int pgrp = tid >> 6, t64 = tid & 63;
    const int c = t64 & 15, r0 = t64 >> 4;               // rows r0 + 4 q, q = 0..15
    // swizzled offsets of (row r0, chunk c) and (row r0 + 4, chunk c): rows 8 apart keep the swizzle phase
    const uint32_t a_off0 = (uint32_t)(c >> 3) * BLOCK_BYTES + (uint32_t)r0 * ROWB + (uint32_t)(((c & 7) ^ (r0 & 7)) << 4);
    const uint32_t a_off1 = (uint32_t)(c >> 3) * BLOCK_BYTES + (uint32_t)(r0 + 4) * ROWB + (uint32_t)(((c & 7) ^ ((r0 + 4) & 7)) << 4);
    const char* xs = reinterpret_cast<const char*>(a.in);
    const size_t x_row = (size_t)cin * 4;
    const bool table = a.mode == FD_GATHER_TABLE;
    const int koff = cin >= 128 ? 0 : (c * 8) / cin;     // this thread's kernel offset inside a group (Cin < 128)
    const int ch_n = cin >= 128 ? c * 8 : (c * 8) % cin;
    const int n_items = n_stages * G;
    int cur_st = 0;                                      // stage whose index slice this warp is reading
    bool have_idx = false;
    for (int item = pgrp; item < n_items; item += PGROUPS) {
      const int st = item / G, g = item - st * G;
      const uint32_t as = (uint32_t)item % SA, aph = ((uint32_t)item / SA) & 1;
      int k, ch;
      if (cin >= 128) { const int grp = g0 + g; k = grp / pl.ci_tiles; ch = (grp - k * pl.ci_tiles) * 128 + ch_n; }
      else { k = (g0 + g) * pl.opg + koff; ch = ch_n; }
      const bool kv = k < a.K;
      if (table) {
        if (st != cur_st) {                                 // done with the slices of stages cur_st .. st - 1
          __syncwarp();
          if (lane == 0) for (int s2 = cur_st; s2 < st; ++s2) wg_mbar_arrive(wg::smem_u32(&i_empty[s2 % IDX_SLOTS]));
          cur_st = st;
          have_idx = false;
        }
        if (!have_idx) { ok = wait_warp(wg::smem_u32(&i_full[st % IDX_SLOTS]), (uint32_t)((st / IDX_SLOTS) & 1), pl.dbg & 8) && ok; have_idx = true; }
      }
      const int o0 = row_begin + st * KP + r0;
      const uint32_t ibase = idx_ring + (uint32_t)(st % IDX_SLOTS) * (IDX_K * KP * 4) + (uint32_t)(k * KP + r0) * 4;
      int src[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int o = o0 + 4 * q;
        if (!kv || o >= row_end) src[q] = -1;
        else if (table) src[q] = (int)wg_lds(ibase + (uint32_t)(4 * q) * 4);
        else src[q] = gather_row(a, o, k);
      }
      if (t64 == 0) WOS_TRACE(0, item, clock64());
      ok = wait_warp(wg::smem_u32(&a_empty[as]), aph ^ 1, pl.dbg & 8) && ok;
      if (t64 == 0) WOS_TRACE(1, item, clock64());
      const uint32_t adst = a_ring + as * A_BYTES;
      const char* xh = xs + ch * 2;
      if (!(pl.dbg & 1)) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const uint32_t sz = src[q] >= 0 ? 16u : 0u;
          const char* sp = xh + (size_t)(uint32_t)max(src[q], 0) * x_row;
          const uint32_t d = adst + ((q & 1) ? a_off1 : a_off0) + (uint32_t)(q >> 1) * (8 * ROWB);
          cp_async16_sz(d, sp, sz);
          cp_async16_sz(d + A_PLANE, sp + cin * 2, sz);
        }
      }
      cp_async_arrive_noinc(wg::smem_u32(&a_full[as]));
    }
    if (table) {                                           // release the slices of the remaining stages
      __syncwarp();
      if (lane == 0) for (int s2 = cur_st; s2 < n_stages; ++s2) wg_mbar_arrive(wg::smem_u32(&i_empty[s2 % IDX_SLOTS]));
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == PRODUCERS / 32 + 1) {
    // ===================================== LOADER =====================================
    // per stage: the dY tile (64 consecutive output rows, loaded once for all groups of the pass) and, for rulebook
    // layers, the [K][64] slice of the neighbour table -- up to two stages ahead of the gather warps
    const char* ys = reinterpret_cast<const char*>(a.out);
    const size_t y_row = (size_t)cout * 4;
    const int bchunks = NT >> 3, b_items = KP * bchunks;
    const uint32_t b_plane = (uint32_t)((NT + 63) >> 6) * BLOCK_BYTES;
    const bool table = a.mode == FD_GATHER_TABLE;
    const bool idx_vec = table && (a.nbr_stride & 3) == 0 && (((uintptr_t)a.nbr) & 15) == 0 && (row_begin & 3) == 0;
    // index slice of stage si into ring slot si % IDX_SLOTS (after every gather warp has released its previous occupant)
    auto load_indices = [&](int si) {
      const int buf = si % IDX_SLOTS, o0 = row_begin + si * KP;
      if (si >= IDX_SLOTS) ok = wait_warp(wg::smem_u32(&i_empty[buf]), (uint32_t)((si / IDX_SLOTS - 1) & 1)) && ok;
      const uint32_t dst = idx_ring + (uint32_t)buf * (IDX_K * KP * 4);
      if (idx_vec) {
        // 16-byte copies (4 rows); rows past the chunk's end are zero-filled and never read
        for (int e = lane; e < a.K * (KP / 4); e += 32) {
          const int k = e >> 4, r = (e & 15) * 4;
          const int valid = min(max(row_end - (o0 + r), 0), 4);
          const int32_t* src = a.nbr + (size_t)k * a.nbr_stride + min(o0 + r, (row_end - 1) & ~3);
          cp_async16_sz(dst + (uint32_t)(k * KP + r) * 4, src, (uint32_t)valid * 4);
        }
      } else {
        for (int e = lane; e < a.K * KP; e += 32) {
          const int k = e >> 6, r = e & (KP - 1);
          const int o = min(o0 + r, row_end - 1);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)e * 4), "l"(a.nbr + (size_t)k * a.nbr_stride + o) : "memory");
        }
      }
      cp_async_arrive_noinc(wg::smem_u32(&i_full[buf]));
    };
    // the slices run IDX_SLOTS - 2 stages ahead of the dY tiles (which are gated by the MMAs of two stages ago): the
    // gather warps are up to a ring of items ahead of the MMA and must never wait for a slice's DRAM round trip
    constexpr int AHEAD = IDX_SLOTS - 2;
    if (table) for (int si = 0; si < AHEAD && si < n_stages; ++si) load_indices(si);
    uint32_t bs = 0, bph = 0;
    for (int st = 0; st < n_stages; ++st) {
      const int o0 = row_begin + st * KP;
      if (table && st + AHEAD < n_stages) load_indices(st + AHEAD);
      ok = wait_warp(wg::smem_u32(&b_empty[bs]), bph ^ 1) && ok;
      const uint32_t bdst = b_ring + bs * B_BYTES;
      for (int e = lane; e < b_items; e += 32) {
        const int r = e / bchunks, cc = e - r * bchunks;
        const uint32_t sz = o0 + r < row_end ? 16u : 0u;
        const char* src = ys + dy_row(min(o0 + r, row_end - 1)) * y_row + co0 * 2 + cc * 16;
        const uint32_t d = bdst + (uint32_t)(cc >> 3) * BLOCK_BYTES + (uint32_t)r * ROWB + (uint32_t)(((cc & 7) ^ (r & 7)) << 4);
        cp_async16_sz(d, src, sz);
        // lo plane: the next 64-channel block, or (fused narrow tiles) the channels right after the hi ones in the same rows
        const uint32_t dlo = (pl.fuse && NT < 64) ? bdst + (uint32_t)r * ROWB + (uint32_t)((((cc + bchunks) & 7) ^ (r & 7)) << 4) : d + b_plane;
        cp_async16_sz(dlo, src + cout * 2, sz);
      }
      cp_async_arrive_noinc(wg::smem_u32(&b_full[bs]));
      if (++bs == SB) { bs = 0; bph ^= 1; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (lane == 0) {
    // ===================================== MMA ISSUER =====================================
    const uint32_t idesc = wg::idesc_mn(128, NT), idesc2 = wg::idesc_mn(128, 2 * NT);
    const uint32_t b_plane = (uint32_t)((NT + 63) >> 6) * BLOCK_BYTES;
    // The barrier of the NEXT item (and, at a stage's last group, of the next dY stage) is probed with a non-blocking
    // test_wait issued BEFORE the current item's MMAs, so that its latency overlaps their issue (which blocks on
    // execution); a blocking wait only follows when the probe said "not yet".
    uint32_t as = 0, aph = 0, bs = 0, bph = 0;
    uint32_t a_ready = 0, b_ready = 0;
    for (int st = 0; st < n_stages && ok; ++st) {
      if (!b_ready) ok = wg::mbar_wait(wg::smem_u32(&b_full[bs]), bph) && ok;
      const uint32_t sB = b_ring + bs * B_BYTES;
      const uint64_t dBh = wg::desc_mn_sw128(sB), dBl = wg::desc_mn_sw128(sB + b_plane);
      const uint32_t nbs = bs + 1 == SB ? 0 : bs + 1, nbph = bs + 1 == SB ? bph ^ 1 : bph;
      for (int g = 0; g < G && ok; ++g) {
        if (!a_ready) ok = wg::mbar_wait(wg::smem_u32(&a_full[as]), aph) && ok;
        WOS_TRACE(2, st * G + g, clock64());
        if (!(pl.dbg & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t nas = as + 1 == SA ? 0 : as + 1, naph = as + 1 == SA ? aph ^ 1 : aph;
        a_ready = wg::mbar_test(wg::smem_u32(&a_full[nas]), naph);
        if (g == G - 1) b_ready = wg::mbar_test(wg::smem_u32(&b_full[nbs]), nbph);
        const uint32_t sA = a_ring + as * A_BYTES;
        const uint64_t dAh = wg::desc_mn_sw128(sA), dAl = wg::desc_mn_sw128(sA + A_PLANE);
        const uint32_t td = tmem_base + (uint32_t)(g * pl.gcols);
        if (!(pl.dbg & 2)) {
#pragma unroll
          for (int ks = 0; ks < KP / 16; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 16 * ROWB) >> 4);
            const uint32_t acc = (st > 0 || ks > 0) ? 1u : 0u;
            if (pl.fuse) {                                   // columns [0, NT): A_hi B_hi + A_lo B_hi; [NT, 2 NT): A_hi B_lo
              wg::umma(td, dAh + adv, dBh + adv, idesc2, acc);
              wg::umma(td, dAl + adv, dBh + adv, idesc, 1u);
            } else {
              wg::umma(td, dAh + adv, dBh + adv, idesc, acc);
              wg::umma(td, dAh + adv, dBl + adv, idesc, 1u);
              wg::umma(td, dAl + adv, dBh + adv, idesc, 1u);
            }
          }
        }
        wg::umma_commit(wg::smem_u32(&a_empty[as]));
        WOS_TRACE(3, st * G + g, clock64());
        as = nas; aph = naph;
      }
      wg::umma_commit(wg::smem_u32(&b_empty[bs]));
      bs = nbs; bph = nbph;
    }
    wg::umma_commit(wg::smem_u32(done));
  }
  // ===================================== EPILOGUE =====================================
  __syncwarp();
  if (warp < PRODUCERS / 32) {
    ok = wait_warp(wg::smem_u32(done), 0) && ok;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
      const int lq = warp & 3, half = warp >> 2;           // TMEM lane quarter; the two warps of a quarter alternate groups
      const int ls = lq * 32 + lane;
      for (int g = half; g < G; g += 2) {
        int ch; const int k = group_k(g0 + g, ls, ch);
        float* dst = (pslot ? pslot : dw) + ((size_t)k * cin + ch) * cout + co0;
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          wg::tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(g * pl.gcols + c0), v);
          if (pl.fuse) {
            uint32_t v2[16];
            wg::tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(g * pl.gcols + NT + c0), v2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (k < a.K) {
            if (pslot) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (v[j] != 0u) atomicAdd(dst + c0 + j, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  if (!ok && lane == 0) atomicAdd(&g_wg_abort, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == PRODUCERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols));
}

static bool wos_plan(const ConvArgs& a, wos::Plan* out) {
  const int cin = a.cin, cout = a.cout;
  const bool cin_ok = cin == 16 || cin == 32 || cin == 64 || (cin >= 128 && cin % 128 == 0);
  const bool cout_ok = cout == 16 || cout == 32 || cout == 64 || (cout >= 128 && cout % 128 == 0);
  if (!cin_ok || !cout_ok || (a.out_map != FD_OUTMAP_IDENTITY && a.out_map != OUTMAP_UPSAMPLE)) return false;
  wos::Plan p{};
  p.opg = cin >= 128 ? 1 : 128 / cin;
  p.ci_tiles = cin >= 128 ? cin / 128 : 1;
  p.n_groups = cin >= 128 ? a.K * p.ci_tiles : ceil_div(a.K, p.opg);
  p.NT = cout < 128 ? cout : 128;
  p.co_tiles = cout < 128 ? 1 : cout / 128;
  static const int nofuse = getenv("FD_WG_NOFUSE") ? atoi(getenv("FD_WG_NOFUSE")) : 0;
  p.fuse = (cout <= 64 && !nofuse) ? 1 : 0;
  p.gcols = p.fuse ? 2 * p.NT : p.NT;
  const int max_gpp = 512 / p.gcols;
  p.passes = ceil_div(p.n_groups, max_gpp);
  p.gpp = ceil_div(p.n_groups, p.passes);
  p.passes = ceil_div(p.n_groups, p.gpp);
  int cols = 32;
  while (cols < p.gpp * p.gcols) cols <<= 1;
  p.tmem_cols = cols;
  int chunks = kNumSMs / (p.passes * p.co_tiles);
  const int max_chunks = ceil_div(a.n_cap, 4 * wos::KP);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  p.chunks = chunks;
  p.b_bytes = 2 * ((p.NT + 63) / 64) * wos::BLOCK_BYTES;
  static const int dbg = getenv("FD_WG_DBG") ? atoi(getenv("FD_WG_DBG")) : 0;
  p.dbg = dbg;
  *out = p;
  return true;
}

// eligibility of the output-stationary kernel (needs the row capacity of d_in to pre-split it, and a workspace)
bool wgrad_os_supported(const ConvArgs& a) {
  wos::Plan p;
  static const int off = getenv("FD_WG_OS_OFF") ? atoi(getenv("FD_WG_OS_OFF")) : 0;
  return !off && a.n_in_cap > 0 && a.in_fmt == FD_FMT_FP32 && a.out_fmt == FD_FMT_FP32 && a.in_stride % 4 == 0 &&
         a.out_stride % 4 == 0 && ((((uintptr_t)a.in) | ((uintptr_t)a.out)) & 15) == 0 && a.K <= wos::IDX_K && wos_plan(a, &p);
}
int conv_wgrad_os_chunks(const ConvArgs& a) {
  wos::Plan p;
  return wos_plan(a, &p) ? p.chunks : 0;
}
static size_t wos_align(size_t x) { return (x + 255) & ~(size_t)255; }
// bytes of the split copies of x and dy that follow the partial slots in the workspace
static size_t wos_dy_rows(const ConvArgs& a) {      // rows of dL/dy: a ConvTranspose2d phase reads every s-th pixel of the full map
  return a.out_map == OUTMAP_UPSAMPLE ? (size_t)a.n_cap * a.up_s * a.up_s : (size_t)a.n_cap;
}
size_t conv_wgrad_os_extra_bytes(const ConvArgs& a) {
  return wos_align((size_t)a.n_in_cap * a.cin * 4) + wos_align(wos_dy_rows(a) * a.cout * 4) + 512;
}
// `partial`: chunks x [K, Cin, Cout] slots (ordered reduce by the caller); `extra`: conv_wgrad_os_extra_bytes(a) bytes
int conv_wgrad_os(const ConvArgs& a, float* partial, void* extra, cudaStream_t stream) {
  if (a.n_cap <= 0) return 0;
  wos::Plan p;
  FD_REQUIRE(wos_plan(a, &p), "fd_conv_wgrad: unsupported shape for the output-stationary kernel");
  static bool configured = false;
  if (!configured) {
    FD_CUDA(cudaFuncSetAttribute(conv_wgrad_os_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wos::smem_bytes(4, 32768)));
    FD_CUDA(cudaFuncSetAttribute(conv_wgrad_os_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wos::smem_bytes(5, 16384)));
    configured = true;
  }
  char* e = (char*)(((uintptr_t)extra + 255) & ~(uintptr_t)255);
  unsigned short* xs = (unsigned short*)e;
  unsigned short* ys = (unsigned short*)(e + wos_align((size_t)a.n_in_cap * a.cin * 4));
  if (a.in_split) {
    xs = (unsigned short*)a.in_split;
  } else {
    split_rows_kernel<<<persistent_grid(ceil_div((int64_t)a.n_in_cap * (a.cin / 8), 256), 8), 256, 0, stream>>>(
        a.in, a.in_stride, a.cin, nullptr, a.n_in_cap, xs);
    FD_LAUNCHED();
  }
  if (a.out_split) {
    ys = (unsigned short*)a.out_split;
  } else {
    const int64_t dy_rows = (int64_t)wos_dy_rows(a);
    split_rows_kernel<<<persistent_grid(ceil_div(dy_rows * (a.cout / 8), 256), 8), 256, 0, stream>>>(
        a.out, a.out_stride, a.cout, a.out_map == FD_OUTMAP_IDENTITY ? a.d_n : nullptr, dy_rows, ys);
    FD_LAUNCHED();
  }
  ConvArgs s = a;
  s.in = reinterpret_cast<const float*>(xs); s.in_stride = a.cin; s.in_ctot = a.cin; s.in_fmt = FD_FMT_SPLIT_BF16;
  s.out = reinterpret_cast<float*>(ys); s.out_stride = a.cout; s.out_ctot = a.cout; s.out_fmt = FD_FMT_SPLIT_BF16;
  const dim3 grid(p.chunks, p.passes, p.co_tiles);
  if (p.b_bytes <= 16384) conv_wgrad_os_kernel<5><<<grid, wos::THREADS, wos::smem_bytes(5, p.b_bytes), stream>>>(s, nullptr, partial, p);
  else conv_wgrad_os_kernel<4><<<grid, wos::THREADS, wos::smem_bytes(4, p.b_bytes), stream>>>(s, nullptr, partial, p);
  FD_LAUNCHED();
  return 0;
}

// eligibility of the tensor-core arm: plain fp32 rows with 16-byte aligned 8-channel groups on both operands
bool wgrad_tc_supported(const ConvArgs& a) {
  return a.out_map == FD_OUTMAP_IDENTITY && a.cin % 8 == 0 && a.cout % 8 == 0 && a.in_stride % 4 == 0 &&
         a.out_stride % 4 == 0 && ((((uintptr_t)a.in) | ((uintptr_t)a.out)) & 15) == 0;
}

// row chunks (= partial slots) of a launch
int conv_wgrad_tc_chunks(const ConvArgs& a) {
  const int tiles = a.K * ceil_div(a.cin, 128) * ceil_div(a.cout, 128);
  static const int waves = getenv("FD_WG_WAVES") ? atoi(getenv("FD_WG_WAVES")) : 4;
  int chunks = ceil_div((int64_t)kNumSMs * waves, tiles);
  const int max_chunks = ceil_div(a.n_cap, 1024);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int rows_per_cta = ceil_div(a.n_cap, chunks);
  rows_per_cta = ceil_div(rows_per_cta, 64) * 64;
  return ceil_div(a.n_cap, rows_per_cta);
}

int conv_wgrad_tc(const ConvArgs& a, float* dw, float* partial, cudaStream_t stream) {
  if (a.n_cap <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    FD_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::SMEM));
    configured = true;
  }
  const int tiles_ci = ceil_div(a.cin, 128), tiles_co = ceil_div(a.cout, 128);
  const int tiles = a.K * tiles_ci * tiles_co;
  FD_REQUIRE(tiles <= 65535, "fd_conv_wgrad: K*tiles = %d exceeds the grid limit", tiles);
  const int chunks = conv_wgrad_tc_chunks(a);
  const int rows_per_cta = ceil_div(ceil_div(a.n_cap, chunks), 64) * 64;
  conv_wgrad_tc_kernel<<<dim3(chunks, tiles), wg::THREADS, wg::SMEM, stream>>>(a, dw, partial, rows_per_cta, tiles_ci, tiles_co);
  FD_LAUNCHED();
  return 0;
}

}  // namespace fd

extern "C" {
/* perf-triage helper (not part of the documented ABI): copy the FD_WG_DBG&32 timeline of block (0,0) to the host */
int fd_debug_read_wgrad_trace(long long* out, int role) {
  if (role < 0 || role >= 4) return -1;
  return (int)cudaMemcpyFromSymbol(out, fd::g_wos_trace, sizeof(long long) * fd::WOS_TRACE_N,
                                   sizeof(long long) * fd::WOS_TRACE_N * role, cudaMemcpyDeviceToHost);
}
}
