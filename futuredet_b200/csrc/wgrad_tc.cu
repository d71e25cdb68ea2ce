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

__device__ int g_wg_abort = 0;

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
