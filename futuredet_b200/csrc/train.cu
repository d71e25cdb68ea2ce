// Backward (training) kernels of the hot path for sm_100a: rulebook transpose, convolution weight gradient,
// train-mode BatchNorm forward statistics / backward, bias gradient, gradient accumulation, BEV scatter/gather.
//
// Replaces what torch autograd runs under the reference's `loss.backward()` (det3d/torchie/trainer/trainer.py:317-344):
// spconv 1.x `indice_conv_backward` (call sites det3d/models/backbones/scn.py:11-34,98-146), ATen batch_norm
// forward(training)/backward (scn.py:51-52,64-80; det3d/models/necks/rpn.py:124-142; center_head.py:129-152),
// cuDNN Conv2d/ConvTranspose2d wgrad, and SparseConvTensor.dense() backward (scn.py:165-168).
// Data gradients reuse the forward implicit-GEMM kernels (see include/futuredet_b200.h).
// All reductions except the weight-gradient accumulation are deterministic (two-stage, fixed order, fp64).
#include "conv_common.cuh"

namespace fd {

// ---------------------------------------------------------------------------------------------- rulebook
__global__ void __launch_bounds__(256)
nbr_transpose_kernel(const int* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ d_n, int n_cap, int K,
                     int* __restrict__ nbr_t, int nbr_t_stride, int n_in_cap) {
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const long long total = (long long)K * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / n), o = (int)(e - (long long)k * n);
    const int i = nbr[(size_t)k * nbr_stride + o];
    // (input row, offset) determines the output row uniquely (out*s - p + k = in), so there are no write conflicts
    if (i >= 0 && i < n_in_cap) nbr_t[(size_t)k * nbr_t_stride + i] = o;
  }
}

// ---------------------------------------------------------------------------------------------- wgrad
// dW[k][ci][co] += sum_o in[gather(o,k)][ci] * dy[o][co].  One CTA = one kernel offset x one TI x TO (ci,co) tile x one
// chunk of output rows.  The 256 threads form G groups; every group owns the whole tile with an RI x RO register block
// per thread (8 x 8 for tile sides >= 64, else 4 x 4) and takes every G-th row of a 32-row batch staged in shared
// memory (128-bit gathers when the rows allow it), so narrow layers keep all threads busy.
constexpr int WG_R = 32;
constexpr int WG_THREADS = 256;
constexpr int WG_Q = 512;          // pair ring: holds < WG_R leftovers + one batch of WG_THREADS rows

// output-channel owned by register j of thread column tx: groups of 4 consecutive channels, interleaved across the 16
// thread columns so that the 128-bit shared-memory reads of a warp are conflict free
constexpr int wg_rows(int width) {      // rows per shared-memory batch: 32 (larger batches measured slower: occupancy)
  int r = 32;
  while (r * width > 8192) r >>= 1;
  return r;
}

template <int RO, int TXN>
__device__ __forceinline__ int wg_col(int tx, int j) {
  return (j >> 2) * (TXN * 4) + tx * 4 + (j & 3);
}

template <int TI, int TO>
__global__ void __launch_bounds__(WG_THREADS)
conv_wgrad_kernel(const ConvArgs a, float* __restrict__ dw, float* __restrict__ partial, int rows_per_cta, int tiles_ci,
                  int tiles_co) {
  constexpr int RI = TI >= 64 ? 8 : 4, RO = TO >= 64 ? 8 : 4;
  constexpr int TXN = TO / RO, TYN = TI / RI, GS = TXN * TYN, G = WG_THREADS / GS;
  // rows per shared-memory batch: 32 KB of operands in flight per CTA whatever the tile width
  constexpr int R = wg_rows(TI + TO);
  static_assert(GS <= WG_THREADS && WG_THREADS % GS == 0 && R % G == 0 && R - 1 + WG_THREADS <= WG_Q, "bad wgrad tiling");
  __shared__ __align__(16) float As[R][TI];
  __shared__ __align__(16) float Bs[R][TO];
  __shared__ int q_in[WG_Q], q_out[WG_Q];
  __shared__ int s_wcnt[WG_THREADS / 32];
  const int tid = threadIdx.x;
  const int n = a.d_n ? min(*a.d_n, a.n_cap) : a.n_cap;
  int t = blockIdx.y;
  const int to = t % tiles_co; t /= tiles_co;
  const int ti = t % tiles_ci;
  const int k = t / tiles_ci;
  const int ci0 = ti * TI, co0 = to * TO;
  // the row chunks are cut from the DEVICE-side row count: capacities of strided levels are several times the number
  // of active sites, and chunks cut from the capacity would leave most CTAs without rows
  rows_per_cta = ((n + (int)gridDim.x - 1) / (int)gridDim.x + 63) & ~63;
  const long long rb = (long long)blockIdx.x * rows_per_cta;
  const int grp = tid / GS, tx = (tid % GS) % TXN, ty = (tid % GS) / TXN;   // tx -> RO output, ty -> RI input channels
  // deterministic mode: every (row chunk, thread group) owns a slot of the partial buffer laid out like dW; the slots
  // are added in ascending order by wgrad_reduce_kernel (no float atomics -> bit-reproducible gradients)
  float* pslot = partial ? partial + ((size_t)blockIdx.x * G + grp) * a.K * a.cin * a.cout : nullptr;
  if (rb >= n && !pslot) return;
  const int row_begin = (int)min((long long)n, rb);
  const int row_end = (int)min((long long)n, rb + rows_per_cta);
  const bool vec_a = (a.cin % 4 == 0) && (a.in_stride % 4 == 0) && ((((uintptr_t)a.in) & 15) == 0);
  const bool vec_b = a.out_map == FD_OUTMAP_IDENTITY && (a.cout % 4 == 0) && (a.out_stride % 4 == 0) &&
                     ((((uintptr_t)a.out) & 15) == 0);
  float acc[RI][RO];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < RO; ++j) acc[i][j] = 0.f;

  // (input row, output row) pairs of this offset are compacted on the fly into a ring and consumed 32 at a time, so the
  // GEMM only ever sees rows that have a neighbour (about 20 % of the rows per offset on the first sparse levels)
  const int lane = tid & 31, warp = tid >> 5;
  int q_head = 0, q_cnt = 0;                  // uniform across the CTA
  for (int base = row_begin; base < row_end || q_cnt > 0; base += WG_THREADS) {
    if (base < row_end) {
      const int o = base + tid;
      const int src = o < row_end ? gather_row(a, o, k) : -1;
      const unsigned ballot = __ballot_sync(0xffffffffu, src >= 0);
      if (lane == 0) s_wcnt[warp] = __popc(ballot);
      __syncthreads();
      int woff = 0, total = 0;
#pragma unroll
      for (int w = 0; w < WG_THREADS / 32; ++w) {
        const int c = s_wcnt[w];
        if (w < warp) woff += c;
        total += c;
      }
      if (src >= 0) {
        const int pos = (q_head + q_cnt + woff + __popc(ballot & ((1u << lane) - 1u))) & (WG_Q - 1);
        q_in[pos] = src;
        q_out[pos] = o;
      }
      q_cnt += total;
      __syncthreads();
    }
    const bool last = base + WG_THREADS >= row_end;
    while (q_cnt >= R || (last && q_cnt > 0)) {
      const int take = min(q_cnt, R);
      // ---- A: gathered input rows
      if (vec_a) {
        for (int e = tid; e < R * TI / 4; e += WG_THREADS) {
          const int r = e / (TI / 4), c = (e % (TI / 4)) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < take && ci0 + c < a.cin)
            v = __ldg(reinterpret_cast<const float4*>(a.in + (size_t)q_in[(q_head + r) & (WG_Q - 1)] * a.in_stride + ci0 + c));
          *reinterpret_cast<float4*>(&As[r][c]) = v;
        }
      } else {
        for (int e = tid; e < R * TI; e += WG_THREADS) {
          const int r = e / TI, c = e % TI;
          As[r][c] = (r < take && ci0 + c < a.cin)
                         ? __ldg(a.in + (size_t)q_in[(q_head + r) & (WG_Q - 1)] * a.in_stride + ci0 + c) : 0.f;
        }
      }
      // ---- B: dL/dy rows of the same pairs
      if (vec_b) {
        for (int e = tid; e < R * TO / 4; e += WG_THREADS) {
          const int r = e / (TO / 4), c = (e % (TO / 4)) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < take && co0 + c < a.cout)
            v = *reinterpret_cast<const float4*>(a.out + (size_t)q_out[(q_head + r) & (WG_Q - 1)] * a.out_stride + co0 + c);
          *reinterpret_cast<float4*>(&Bs[r][c]) = v;
        }
      } else {
        for (int e = tid; e < R * TO; e += WG_THREADS) {
          const int r = e / TO, c = e % TO;
          float v = 0.f;
          if (r < take && co0 + c < a.cout) {
            const OutRow orow = map_out_row(a, q_out[(q_head + r) & (WG_Q - 1)]);
            v = orow.base[orow.coff + (co0 + c) * orow.cstride];
          }
          Bs[r][c] = v;
        }
      }
      __syncthreads();
#pragma unroll 2
      for (int r = grp; r < R; r += G) {
        float av[RI], bv[RO];
#pragma unroll
        for (int i = 0; i < RI; ++i) av[i] = As[r][ty * RI + i];
#pragma unroll
        for (int j = 0; j < RO; ++j) bv[j] = Bs[r][wg_col<RO, TXN>(tx, j)];
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
          for (int j = 0; j < RO; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
      q_head = (q_head + take) & (WG_Q - 1);
      q_cnt -= take;
    }
  }
  float* dwk = (pslot ? pslot : dw) + (size_t)k * a.cin * a.cout;
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int ci = ci0 + ty * RI + i;
    if (ci >= a.cin) continue;
#pragma unroll
    for (int j = 0; j < RO; ++j) {
      const int co = co0 + wg_col<RO, TXN>(tx, j);
      if (co < a.cout) {
        if (pslot) dwk[(size_t)ci * a.cout + co] = acc[i][j];
        else if (acc[i][j] != 0.f) atomicAdd(dwk + (size_t)ci * a.cout + co, acc[i][j]);
      }
    }
  }
}

// dW[e] += sum over slots of partial[slot][e]: the ordered reduce of the deterministic weight gradient.  A block is 32
// consecutive elements x 8 slot lanes (warp w adds slots w, w + 8, ... in ascending order, 4 loads in flight), then
// the 8 lane sums are added in lane order: a fixed summation tree, so the result is bit-identical run to run.  (One
// thread per element walking all slots took ~100 us on the small head layers: hundreds of dependent round trips.)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int slots, long long elems, float* __restrict__ dw) {
  __shared__ float sh[8][32];
  const int el = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (long long e0 = (long long)blockIdx.x * 32; e0 < elems; e0 += (long long)gridDim.x * 32) {
    const long long e = e0 + el;
    float s = 0.f;
    if (e < elems) {
      int c = sl;
      for (; c + 24 < slots; c += 32) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldcs(partial + (size_t)(c + 8 * j) * elems + e);
#pragma unroll
        for (int j = 0; j < 4; ++j) s = __fadd_rn(s, v[j]);
      }
      for (; c < slots; c += 8) s = __fadd_rn(s, __ldcs(partial + (size_t)c * elems + e));
    }
    sh[sl][el] = s;
    __syncthreads();
    if (sl == 0 && e < elems) {
      float t = sh[0][el];
#pragma unroll
      for (int j = 1; j < 8; ++j) t = __fadd_rn(t, sh[j][el]);
      dw[e] = __fadd_rn(dw[e], t);
    }
    __syncthreads();
  }
}

static int wg_tile(int c) { return c >= 96 ? 128 : c >= 48 ? 64 : c >= 24 ? 32 : 16; }

constexpr int wg_groups(int TI, int TO) {
  return WG_THREADS / ((TO / (TO >= 64 ? 8 : 4)) * (TI / (TI >= 64 ? 8 : 4)));
}
static int wg_simt_chunks(const ConvArgs& a, int TI, int TO) {
  const int tiles = a.K * ceil_div(a.cin, TI) * ceil_div(a.cout, TO);
  int chunks = ceil_div((int64_t)kNumSMs * 8, tiles);
  const int max_chunks = ceil_div(a.n_cap, 8 * WG_R);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int rows_per_cta = ceil_div(a.n_cap, chunks);
  rows_per_cta = ceil_div(rows_per_cta, WG_R) * WG_R;
  return ceil_div(a.n_cap, rows_per_cta);
}

// partial == nullptr: accumulate with atomics; otherwise *slots receives the number of partial slots written
template <int TI, int TO>
static int launch_wgrad_t(const ConvArgs& a, float* dw, float* partial, int* slots, cudaStream_t stream) {
  const int tiles_ci = ceil_div(a.cin, TI), tiles_co = ceil_div(a.cout, TO);
  const int tiles = a.K * tiles_ci * tiles_co;
  FD_REQUIRE(tiles <= 65535, "fd_conv_wgrad: K*tiles = %d exceeds the grid limit", tiles);
  const int chunks = wg_simt_chunks(a, TI, TO);
  const int rows_per_cta = ceil_div(ceil_div(a.n_cap, chunks), WG_R) * WG_R;
  if (slots) *slots = chunks * wg_groups(TI, TO);
  if (!stream && !dw) return 0;                                   // planning call
  conv_wgrad_kernel<TI, TO><<<dim3(chunks, tiles), WG_THREADS, 0, stream>>>(a, dw, partial, rows_per_cta, tiles_ci, tiles_co);
  FD_LAUNCHED();
  return 0;
}

template <int TI>
static int launch_wgrad_i(const ConvArgs& a, float* dw, float* partial, int* slots, cudaStream_t stream) {
  switch (wg_tile(a.cout)) {
    case 128: return launch_wgrad_t<TI, 128>(a, dw, partial, slots, stream);
    case 64: return launch_wgrad_t<TI, 64>(a, dw, partial, slots, stream);
    case 32: return launch_wgrad_t<TI, 32>(a, dw, partial, slots, stream);
    default: return launch_wgrad_t<TI, 16>(a, dw, partial, slots, stream);
  }
}

// Number of partial slots the deterministic path of this launch writes (planning only, nothing is launched).
static int wgrad_slots(const ConvArgs& a, int precision) {
  if (a.n_cap <= 0) return 0;
  if (precision == FD_PREC_BF16X3 && wgrad_os_supported(a)) return conv_wgrad_os_chunks(a);
  if (precision == FD_PREC_BF16X3 && wgrad_tc_supported(a)) return conv_wgrad_tc_chunks(a);
  int slots = 0;
  switch (wg_tile(a.cin)) {
    case 128: launch_wgrad_i<128>(a, nullptr, nullptr, &slots, nullptr); break;
    case 64: launch_wgrad_i<64>(a, nullptr, nullptr, &slots, nullptr); break;
    case 32: launch_wgrad_i<32>(a, nullptr, nullptr, &slots, nullptr); break;
    default: launch_wgrad_i<16>(a, nullptr, nullptr, &slots, nullptr); break;
  }
  return slots;
}

// workspace of the deterministic path: the partial slots (+ the split-bf16 copies of x and dy for the output-stationary
// tcgen05 kernel)
static size_t wgrad_ws_bytes(const ConvArgs& a, int precision) {
  size_t b = ((size_t)wgrad_slots(a, precision) * a.K * a.cin * a.cout * sizeof(float) + 255) & ~(size_t)255;
  if (a.n_cap > 0 && precision == FD_PREC_BF16X3 && wgrad_os_supported(a)) b += conv_wgrad_os_extra_bytes(a);
  return b;
}

// ws == nullptr: fp32 atomics into dw; otherwise per-chunk partial tiles in ws + an ordered reduce into dw
static int launch_wgrad(const ConvArgs& a, float* dw, cudaStream_t stream, int precision = FD_PREC_FP32, float* ws = nullptr,
                        size_t ws_bytes = 0) {
  if (a.n_cap <= 0) return 0;
  const long long elems = (long long)a.K * a.cin * a.cout;
  int slots = 0;
  if (ws) {
    slots = wgrad_slots(a, precision);
    FD_REQUIRE(wgrad_ws_bytes(a, precision) <= ws_bytes, "fd_conv_wgrad_det: workspace %zu < required %zu", ws_bytes,
               wgrad_ws_bytes(a, precision));
  }
  int rc;
  if (ws && precision == FD_PREC_BF16X3 && wgrad_os_supported(a)) {
    rc = conv_wgrad_os(a, ws, (char*)ws + (((size_t)slots * elems * sizeof(float) + 255) & ~(size_t)255), stream);
  } else if (precision == FD_PREC_BF16X3 && wgrad_tc_supported(a)) {
    rc = conv_wgrad_tc(a, dw, ws, stream);
  } else {
    switch (wg_tile(a.cin)) {
      case 128: rc = launch_wgrad_i<128>(a, dw, ws, nullptr, stream); break;
      case 64: rc = launch_wgrad_i<64>(a, dw, ws, nullptr, stream); break;
      case 32: rc = launch_wgrad_i<32>(a, dw, ws, nullptr, stream); break;
      default: rc = launch_wgrad_i<16>(a, dw, ws, nullptr, stream); break;
    }
  }
  if (rc || !ws) return rc;
  wgrad_reduce_kernel<<<persistent_grid(ceil_div(elems, 32), 8), 256, 0, stream>>>(ws, slots, elems, dw);
  FD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------- reductions
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = kNumSMs * 4;
enum { RED_STATS = 0, RED_BN_BWD = 1, RED_COLSUM = 2 };

struct RedArgs {
  int mode;
  const float* x; int x_stride;
  const float* dy; int dy_stride;
  const float* y; int y_stride; int relu;
  const float* mean; const float* invstd;
  int C; int cp;                    // cp: power of two <= 256, channels handled side by side
  const int32_t* d_n; long long n_cap;
  double* partial;                  // [gridDim.x][2][C]
};

__global__ void __launch_bounds__(RED_THREADS)
red_partial_kernel(const RedArgs a) {
  __shared__ double sh[2][RED_THREADS];
  const long long n = a.d_n ? min((long long)*a.d_n, a.n_cap) : a.n_cap;
  const int tid = threadIdx.x;
  const int ch = tid & (a.cp - 1), rl = tid / a.cp, RL = RED_THREADS / a.cp;
  const long long chunk = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * chunk;
  const long long r1 = min(n, r0 + chunk);
  for (int c0 = 0; c0 < a.C; c0 += a.cp) {
    const int c = c0 + ch;
    double s1 = 0.0, s2 = 0.0;
    if (c < a.C) {
      float mu = 0.f, is = 0.f;
      if (a.mode == RED_BN_BWD) { mu = a.mean[c]; is = a.invstd[c]; }
      for (long long r = r0 + rl; r < r1; r += RL) {
        if (a.mode == RED_STATS) {
          const float v = a.x[(size_t)r * a.x_stride + c];
          s1 += (double)v;
          s2 += (double)v * (double)v;
        } else if (a.mode == RED_BN_BWD) {
          float dz = a.dy[(size_t)r * a.dy_stride + c];
          if (a.relu && !(a.y[(size_t)r * a.y_stride + c] > 0.f)) dz = 0.f;
          const float xh = (a.x[(size_t)r * a.x_stride + c] - mu) * is;
          s1 += (double)dz;
          s2 += (double)dz * (double)xh;
        } else {
          s1 += (double)a.x[(size_t)r * a.x_stride + c];
        }
      }
    }
    sh[0][tid] = s1;
    sh[1][tid] = s2;
    __syncthreads();
    if (rl == 0 && c < a.C) {
      for (int j = 1; j < RL; ++j) { s1 += sh[0][j * a.cp + ch]; s2 += sh[1][j * a.cp + ch]; }
      a.partial[((size_t)blockIdx.x * 2 + 0) * a.C + c] = s1;
      a.partial[((size_t)blockIdx.x * 2 + 1) * a.C + c] = s2;
    }
    __syncthreads();
  }
}

// 4 channels per thread (128-bit loads): the scalar kernel above moved ~3 TB/s on the large levels.  Same partial layout,
// same fixed summation order within a block (row lanes added in ascending order), so results are deterministic.
__global__ void __launch_bounds__(RED_THREADS)
red_partial_vec4_kernel(const RedArgs a) {
  __shared__ double sh[8][RED_THREADS];
  const long long n = a.d_n ? min((long long)*a.d_n, a.n_cap) : a.n_cap;
  const int tid = threadIdx.x;
  const int cq = a.cp >> 2;                          // thread columns (4 channels each), power of two <= 64
  const int chq = tid & (cq - 1), rl = tid / cq, RL = RED_THREADS / cq;
  const long long chunk = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * chunk;
  const long long r1 = min(n, r0 + chunk);
  for (int c0 = 0; c0 < a.C; c0 += a.cp) {
    const int c = c0 + 4 * chq;
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    if (c < a.C) {
      float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu;
      if (a.mode == RED_BN_BWD) { mu = *reinterpret_cast<const float4*>(a.mean + c); is = *reinterpret_cast<const float4*>(a.invstd + c); }
      const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, i4[4] = {is.x, is.y, is.z, is.w};
      for (long long r = r0 + rl; r < r1; r += RL) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)r * a.x_stride + c));
        const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
        if (a.mode == RED_STATS) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { s1[j] += (double)x4[j]; s2[j] += (double)x4[j] * (double)x4[j]; }
        } else if (a.mode == RED_BN_BWD) {
          const float4 dv = __ldg(reinterpret_cast<const float4*>(a.dy + (size_t)r * a.dy_stride + c));
          float d4[4] = {dv.x, dv.y, dv.z, dv.w};
          if (a.relu) {
            const float4 yv = __ldg(reinterpret_cast<const float4*>(a.y + (size_t)r * a.y_stride + c));
            const float y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) if (!(y4[j] > 0.f)) d4[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xh = (x4[j] - m4[j]) * i4[j];
            s1[j] += (double)d4[j];
            s2[j] += (double)d4[j] * (double)xh;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) s1[j] += (double)x4[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { sh[j][tid] = s1[j]; sh[4 + j][tid] = s2[j]; }
    __syncthreads();
    if (rl == 0 && c < a.C) {
      for (int q = 1; q < RL; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1[j] += sh[j][q * cq + chq]; s2[j] += sh[4 + j][q * cq + chq]; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a.partial[((size_t)blockIdx.x * 2 + 0) * a.C + c + j] = s1[j];
        a.partial[((size_t)blockIdx.x * 2 + 1) * a.C + c + j] = s2[j];
      }
    }
    __syncthreads();
  }
}

static bool red_vec4_ok(const RedArgs& a) {
  auto al = [](const void* p, int stride) { return p == nullptr || ((((uintptr_t)p) & 15) == 0 && stride % 4 == 0); };
  return a.C % 4 == 0 && a.cp >= 4 && al(a.x, a.x_stride) && al(a.dy, a.dy_stride) && (!a.relu || al(a.y, a.y_stride)) &&
         al(a.mean, 4) && al(a.invstd, 4);
}
static void launch_red_partial(const RedArgs& r, int G, cudaStream_t stream) {
  if (red_vec4_ok(r)) red_partial_vec4_kernel<<<G, RED_THREADS, 0, stream>>>(r);
  else red_partial_kernel<<<G, RED_THREADS, 0, stream>>>(r);
}

struct FinArgs {
  int mode; int C; int G;
  const double* partial;
  const int32_t* d_n; long long n_cap;
  float eps, momentum;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float* mean; float* invstd; float* scale; float* shift;      // RED_STATS outputs
  float* dgamma; float* dbeta; float* c1; float* c2;            // RED_BN_BWD outputs (c1 = dbeta/n, c2 = dgamma/n)
  float* out;                                                    // RED_COLSUM output
};

// One warp per channel: lane l sums partials l, l+32, ... in order, then a fixed-order butterfly -> deterministic.
__global__ void __launch_bounds__(RED_THREADS)
red_finalize_kernel(const FinArgs a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= a.C) return;
  const long long n = a.d_n ? min((long long)*a.d_n, a.n_cap) : a.n_cap;
  double s1 = 0.0, s2 = 0.0;
  for (int g = lane; g < a.G; g += 32) {
    s1 += a.partial[((size_t)g * 2 + 0) * a.C + c];
    s2 += a.partial[((size_t)g * 2 + 1) * a.C + c];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane != 0) return;
  const double dn = n > 0 ? (double)n : 1.0;
  if (a.mode == RED_STATS) {
    const double mu = s1 / dn;
    double var = s2 / dn - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)a.eps));
    const float g = a.gamma ? a.gamma[c] : 1.f;
    const float b = a.beta ? a.beta[c] : 0.f;
    const float sc = g * is;
    a.mean[c] = (float)mu;
    a.invstd[c] = is;
    a.scale[c] = sc;
    a.shift[c] = b - (float)mu * sc;
    if (a.running_mean) a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (float)mu;
    if (a.running_var) {
      const double unbiased = n > 1 ? var * dn / (dn - 1.0) : var;
      a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
    }
  } else if (a.mode == RED_BN_BWD) {
    if (a.dbeta) a.dbeta[c] = (float)s1;
    if (a.dgamma) a.dgamma[c] = (float)s2;
    a.c1[c] = (float)(s1 / dn);
    a.c2[c] = (float)(s2 / dn);
  } else {
    a.out[c] = (float)s1;
  }
}

static int red_cp(int C) {
  int cp = 1;
  while (cp < C && cp < RED_THREADS) cp <<= 1;
  return cp;
}
// blocks of a reduction over [n, C]: by rows for the long sparse levels, by elements for the short, wide BEV maps
// (8100 x 256 got 32 blocks = a fifth of the GPU and took as long as a 40x larger sparse level)
static int red_blocks(long long n_cap, int C = 1) {
  long long g = (n_cap + 255) / 256;
  const long long ge = (n_cap * C + 8191) / 8192;
  if (ge > g) g = ge;
  if (g > n_cap) g = n_cap;
  if (g < 1) g = 1;
  if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
  return (int)g;
}
static size_t red_partial_bytes(int C) { return sizeof(double) * (size_t)RED_MAX_BLOCKS * 2 * C; }

// ---------------------------------------------------------------------------------------------- elementwise
__global__ void __launch_bounds__(256)
affine_act_kernel(const float* __restrict__ x, int xs, int C, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ res, int rs, int relu,
                  float* __restrict__ y, int ys, unsigned short* __restrict__ ysplit, int split_ctot,
                  const int32_t* __restrict__ d_n, long long n_cap) {
  const long long n = d_n ? min((long long)*d_n, n_cap) : n_cap;
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    float v = x[(size_t)r * xs + c];
    v = fmaf(v, scale ? scale[c] : 1.f, shift ? shift[c] : 0.f);
    if (res) v += res[(size_t)r * rs + c];
    if (relu) v = fmaxf(v, 0.f);
    y[(size_t)r * ys + c] = v;
    if (ysplit) {          // FD_FMT_SPLIT_BF16 copy for the tensor-core consumers (next conv, its weight gradient)
      const unsigned short hi = bf16_bits_rn(v);
      unsigned short* o = ysplit + (size_t)r * 2 * split_ctot + c;
      o[0] = hi;
      o[split_ctot] = bf16_bits_rn(v - __uint_as_float((unsigned)hi << 16));
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, int dys, const float* __restrict__ y, int ys, int relu,
                    const float* __restrict__ x, int xs, int C, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ c1,
                    const float* __restrict__ c2, float* __restrict__ dx, int dxs, float* __restrict__ dres, int drs,
                    unsigned short* __restrict__ dxsplit, const int32_t* __restrict__ d_n, long long n_cap) {
  const long long n = d_n ? min((long long)*d_n, n_cap) : n_cap;
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    float dz = dy[(size_t)r * dys + c];
    if (relu && !(y[(size_t)r * ys + c] > 0.f)) dz = 0.f;
    const float is = invstd[c];
    const float xh = (x[(size_t)r * xs + c] - mean[c]) * is;
    const float g = gamma ? gamma[c] : 1.f;
    const float d = g * is * (dz - c1[c] - xh * c2[c]);
    dx[(size_t)r * dxs + c] = d;
    if (dxsplit) {         // dense FD_FMT_SPLIT_BF16 copy [rows][C hi | C lo] for the data- / weight-gradient convolutions
      const unsigned short hi = bf16_bits_rn(d);
      unsigned short* o = dxsplit + (size_t)r * 2 * C + c;
      o[0] = hi;
      o[C] = bf16_bits_rn(d - __uint_as_float((unsigned)hi << 16));
    }
    if (dres) dres[(size_t)r * drs + c] = dz;
  }
}

// ---- 4 channels per thread (128-bit loads / stores) for 16-byte aligned rows with C % 4 == 0
__device__ __forceinline__ void split4_store(unsigned short* o, int ctot, const float v[4]) {
  unsigned short hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = bf16_bits_rn(v[j]);
    lo[j] = bf16_bits_rn(v[j] - __uint_as_float((unsigned)hi[j] << 16));
  }
  *reinterpret_cast<uint2*>(o) = make_uint2((unsigned)hi[0] | ((unsigned)hi[1] << 16), (unsigned)hi[2] | ((unsigned)hi[3] << 16));
  *reinterpret_cast<uint2*>(o + ctot) = make_uint2((unsigned)lo[0] | ((unsigned)lo[1] << 16), (unsigned)lo[2] | ((unsigned)lo[3] << 16));
}

__global__ void __launch_bounds__(256)
affine_act_vec4_kernel(const float* __restrict__ x, int xs, int C, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ res, int rs, int relu,
                       float* __restrict__ y, int ys, unsigned short* __restrict__ ysplit, int split_ctot,
                       const int32_t* __restrict__ d_n, long long n_cap) {
  const long long n = d_n ? min((long long)*d_n, n_cap) : n_cap;
  const int cq = C >> 2;
  const long long total = n * cq;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cq;
    const int c = (int)(e - r * cq) * 4;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * xs + c));
    const float4 sc = scale ? __ldg(reinterpret_cast<const float4*>(scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sh = shift ? __ldg(reinterpret_cast<const float4*>(shift + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float v[4] = {fmaf(xv.x, sc.x, sh.x), fmaf(xv.y, sc.y, sh.y), fmaf(xv.z, sc.z, sh.z), fmaf(xv.w, sc.w, sh.w)};
    if (res) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(res + (size_t)r * rs + c));
      v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    *reinterpret_cast<float4*>(y + (size_t)r * ys + c) = make_float4(v[0], v[1], v[2], v[3]);
    if (ysplit) split4_store(ysplit + (size_t)r * 2 * split_ctot + c, split_ctot, v);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_vec4_kernel(const float* __restrict__ dy, int dys, const float* __restrict__ y, int ys, int relu,
                         const float* __restrict__ x, int xs, int C, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ c1,
                         const float* __restrict__ c2, float* __restrict__ dx, int dxs, float* __restrict__ dres, int drs,
                         unsigned short* __restrict__ dxsplit, const int32_t* __restrict__ d_n, long long n_cap) {
  const long long n = d_n ? min((long long)*d_n, n_cap) : n_cap;
  const int cq = C >> 2;
  const long long total = n * cq;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cq;
    const int c = (int)(e - r * cq) * 4;
    const float4 dv = __ldg(reinterpret_cast<const float4*>(dy + (size_t)r * dys + c));
    float dz[4] = {dv.x, dv.y, dv.z, dv.w};
    if (relu) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y + (size_t)r * ys + c));
      const float y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) if (!(y4[j] > 0.f)) dz[j] = 0.f;
    }
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * xs + c));
    const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float is = invstd[c + j];
      const float xh = (x4[j] - mean[c + j]) * is;
      const float g = gamma ? gamma[c + j] : 1.f;
      d[j] = g * is * (dz[j] - c1[c + j] - xh * c2[c + j]);
    }
    *reinterpret_cast<float4*>(dx + (size_t)r * dxs + c) = make_float4(d[0], d[1], d[2], d[3]);
    if (dxsplit) split4_store(dxsplit + (size_t)r * 2 * C + c, C, d);
    if (dres) *reinterpret_cast<float4*>(dres + (size_t)r * drs + c) = make_float4(dz[0], dz[1], dz[2], dz[3]);
  }
}

static bool rows_vec4(const void* p, int stride) { return p == nullptr || ((((uintptr_t)p) & 15) == 0 && stride % 4 == 0); }

__global__ void __launch_bounds__(256)
add_rows_kernel(float* __restrict__ dst, int ds, const float* __restrict__ src, int ss, int C,
                const int32_t* __restrict__ d_n, long long n_cap) {
  const long long n = d_n ? min((long long)*d_n, n_cap) : n_cap;
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    dst[(size_t)r * ds + c] += src[(size_t)r * ss + c];
  }
}

// rows [n, C] <-> channels-last BEV [B, H, W, C*D], channel = c*D + z
template <bool TO_BEV>
__global__ void __launch_bounds__(256)
bev_rows_kernel(float* __restrict__ rows, int row_stride, int C, const int4* __restrict__ coords,
                const int32_t* __restrict__ d_n, int n_cap, int D, int H, int W, float* __restrict__ bev) {
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const long long total = (long long)n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / C), c = (int)(e - (long long)r * C);
    const int4 q = coords[r];   // (b, z, y, x)
    const size_t at = (((size_t)q.x * H + q.z) * W + q.w) * ((size_t)C * D) + (size_t)c * D + q.y;
    if (TO_BEV) bev[at] = rows[(size_t)r * row_stride + c];
    else rows[(size_t)r * row_stride + c] = bev[at];
  }
}

// ---------------------------------------------------------------------------------------------- loss backward
struct LossBwdArgs {
  const float* hm; float* ghm; long long hm_sb, hm_sc, hm_ssp;
  const float* gt; int B, C, HW;
  const long long* ind; const unsigned char* mask; const long long* cat; int M, T, NC;
  const float* const* pred_ptr; float* const* gpred_ptr; const long long* pred_sb; const long long* pred_ssp;
  const float* const* tgt_ptr; int tgt_dim; const int* tgt_sel;
  const float* code_w; const float* code_w_forecast; float weight;
  const float* gscale;
};

__device__ __forceinline__ float block_npos(const unsigned char* mask, int BM, int* sh) {
  int v = 0;
  for (int i = threadIdx.x; i < BM; i += blockDim.x) v += mask[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  __syncthreads();
  return (float)t;
}

__device__ __forceinline__ float dsigmoid_clamped(float p) {
  // p = clamp(sigmoid(x), 1e-4, 1-1e-4): d p / d x = p (1 - p) inside the clamp, 0 on it
  return (p > 1e-4f && p < 1.f - 1e-4f) ? p * (1.f - p) : 0.f;
}

// dense part: d/dx of  -(sum log(1-p) p^2 (1-gt)^4) / num_pos
__global__ void __launch_bounds__(256)
focal_grad_kernel(const LossBwdArgs a) {
  __shared__ int sh[8];
  const float npos = block_npos(a.mask, a.B * a.M, sh);
  const float gs = a.gscale ? *a.gscale : 1.f;
  const float coef = npos == 0.f ? -gs : -gs / npos;
  const long long total = (long long)a.B * a.C * a.HW;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e % a.HW);
    const long long bc = e / a.HW;
    const int c = (int)(bc % a.C), b = (int)(bc / a.C);
    const long long at = b * a.hm_sb + c * a.hm_sc + s * a.hm_ssp;
    const float p = a.hm[at];
    const float g = 1.f - a.gt[e];
    const float g2 = g * g;
    const float dneg = (g2 * g2) * (2.f * p * logf(1.f - p) - p * p / (1.f - p));
    a.ghm[at] = coef * dneg * dsigmoid_clamped(p);
  }
}

// object part: positive focal term + masked L1 of every timestep, one thread per (b, m)
__global__ void __launch_bounds__(256)
object_grad_kernel(const LossBwdArgs a) {
  __shared__ int sh[8];
  const int BM = a.B * a.M;
  const float npos = block_npos(a.mask, BM, sh);
  const float gs = a.gscale ? *a.gscale : 1.f;
  const float denom = npos + 1e-4f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BM; i += gridDim.x * blockDim.x) {
    if (!a.mask[i]) continue;
    const int b = i / a.M;
    const long long s = a.ind[i];
    if (npos > 0.f) {
      const long long at = b * a.hm_sb + a.cat[i] * a.hm_sc + s * a.hm_ssp;
      const float p = a.hm[at];
      const float om = 1.f - p;
      const float dpos = om * om / p - 2.f * om * logf(p);
      atomicAdd(a.ghm + at, (-gs / npos) * dpos * dsigmoid_clamped(p));
    }
    for (int t = 0; t < a.T; ++t) {
      const float* tg = a.tgt_ptr[t] + (size_t)i * a.tgt_dim;
      for (int c = 0; c < a.NC; ++c) {
        const float cw = t == 0 ? a.code_w[c] : a.code_w_forecast[c];
        if (cw == 0.f) continue;
        const int q = t * a.NC + c;
        const long long at = b * a.pred_sb[q] + s * a.pred_ssp[q];
        const float d = a.pred_ptr[q][at] - tg[a.tgt_sel[c]];
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        if (sg != 0.f) atomicAdd(a.gpred_ptr[q] + at, gs * a.weight * cw * sg / denom);
      }
    }
  }
}

}  // namespace fd

extern "C" {

int fd_rulebook_transpose(const int32_t* d_nbr, int nbr_stride, const int32_t* d_n_out, int n_out_cap, int K,
                          int32_t* d_nbr_t, int nbr_t_stride, int n_in_cap, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_nbr && d_nbr_t && K >= 1 && n_out_cap >= 0 && n_in_cap >= 0 && nbr_stride >= n_out_cap &&
                 nbr_t_stride >= n_in_cap, "fd_rulebook_transpose: bad argument");
  if (n_in_cap > 0) FD_CUDA(cudaMemsetAsync(d_nbr_t, 0xff, sizeof(int32_t) * (size_t)K * nbr_t_stride, stream));
  if (n_out_cap == 0 || n_in_cap == 0) return 0;
  nbr_transpose_kernel<<<persistent_grid(ceil_div((int64_t)K * n_out_cap, 256), 8), 256, 0, stream>>>(
      d_nbr, nbr_stride, d_n_out, n_out_cap, K, d_nbr_t, nbr_t_stride, n_in_cap);
  FD_LAUNCHED();
  return 0;
}

static int wgrad_entry(const fd_conv_desc* d, float* d_dw, float* ws, size_t ws_bytes, size_t* need, void* stream_);

int fd_conv_wgrad(const fd_conv_desc* d, float* d_dw, void* stream_) {
  return wgrad_entry(d, d_dw, nullptr, 0, nullptr, stream_);
}

size_t fd_conv_wgrad_workspace_bytes(const fd_conv_desc* d) {
  size_t need = 0;
  if (wgrad_entry(d, nullptr, nullptr, 0, &need, nullptr) != 0) return 0;
  return need;
}

int fd_conv_wgrad_det(const fd_conv_desc* d, float* d_dw, void* d_workspace, size_t workspace_bytes, void* stream_) {
  FD_REQUIRE(d_workspace != nullptr, "fd_conv_wgrad_det: null workspace");
  return wgrad_entry(d, d_dw, (float*)d_workspace, workspace_bytes, nullptr, stream_);
}

// need != nullptr: planning call, *need = workspace bytes of the deterministic path, nothing is launched
static int wgrad_entry(const fd_conv_desc* d, float* d_dw, float* ws, size_t ws_bytes, size_t* need, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d != nullptr && (d_dw != nullptr || need), "fd_conv_wgrad: null argument");
  FD_REQUIRE(d->d_in && d->d_out, "fd_conv_wgrad: null in / dy pointer");
  FD_REQUIRE(d->cin >= 1 && d->cout >= 1 && d->K >= 1, "fd_conv_wgrad: bad cin/cout/K");
  FD_REQUIRE(d->in_format == FD_FMT_FP32 && d->out_format == FD_FMT_FP32, "fd_conv_wgrad: fp32 rows only");
  FD_REQUIRE(d->in_stride >= d->cin && d->n_out_cap >= 0, "fd_conv_wgrad: bad stride / row count");
  FD_REQUIRE(d->d_row_perm == nullptr, "fd_conv_wgrad: d_row_perm (sorted tables) is a forward-only option; pass the unsorted table");
  ConvArgs a{};
  a.in = (const float*)d->d_in; a.in_stride = d->in_stride; a.cin = d->cin;
  a.in_fmt = FD_FMT_FP32; a.in_ctot = d->cin; a.out_fmt = FD_FMT_FP32; a.out_ctot = d->cout;
  a.cout = d->cout; a.K = d->K;
  a.out = (float*)d->d_out; a.out_stride = d->out_stride;
  a.d_n = d->d_n_out; a.n_cap = d->n_out_cap;
  a.mode = d->mode; a.nbr = d->d_nbr; a.nbr_stride = d->nbr_stride;
  a.Hin = d->Hin; a.Win = d->Win; a.Hout = d->Hout; a.Wout = d->Wout;
  a.kh = d->kh; a.kw = d->kw; a.sh = d->sh; a.sw = d->sw; a.ph = d->ph; a.pw = d->pw;
  a.out_map = d->out_map;
  a.out_coords = (const int4*)d->d_out_coords4; a.bevD = d->bevD; a.bevH = d->bevH; a.bevW = d->bevW;
  switch (d->mode) {
    case FD_GATHER_TABLE:
      FD_REQUIRE(d->d_nbr && d->nbr_stride >= d->n_out_cap, "fd_conv_wgrad: bad neighbour table");
      FD_REQUIRE(d->out_map == FD_OUTMAP_IDENTITY && d->out_stride >= d->cout, "fd_conv_wgrad: identity out rows only");
      a.n_in_cap = d->n_in_cap;
      a.in_split = d->d_in_split; a.out_split = d->d_out_split;
      if (need) { *need = wgrad_ws_bytes(a, d->precision); return 0; }
      return launch_wgrad(a, d_dw, stream, d->precision, ws, ws_bytes);
    case FD_GATHER_CONV2D:
      FD_REQUIRE(d->K == d->kh * d->kw && d->sh >= 1 && d->sw >= 1, "fd_conv_wgrad: bad conv2d geometry");
      FD_REQUIRE(d->n_out_cap == d->B * d->Hout * d->Wout && !d->d_n_out, "fd_conv_wgrad: conv2d rows must be B*Hout*Wout");
      FD_REQUIRE(d->out_map == FD_OUTMAP_IDENTITY && d->out_stride >= d->cout, "fd_conv_wgrad: identity out rows only");
      a.n_in_cap = d->B * d->Hin * d->Win;
      a.in_split = d->d_in_split; a.out_split = d->d_out_split;
      if (need) { *need = wgrad_ws_bytes(a, d->precision); return 0; }
      return launch_wgrad(a, d_dw, stream, d->precision, ws, ws_bytes);
    case FD_GATHER_CONVT2D: {
      FD_REQUIRE(d->kh == d->sh && d->kw == d->sw && d->kh == d->kw && d->ph == 0 && d->pw == 0 && d->K == d->kh * d->kw,
                 "fd_conv_wgrad: convT2d supports kernel == stride, pad 0 only");
      FD_REQUIRE(d->n_out_cap == d->B * d->Hin * d->Win && !d->d_n_out && d->out_stride >= d->cout,
                 "fd_conv_wgrad: convT2d rows must be B*Hin*Win (input pixels)");
      for (int k = 0; k < d->K; ++k) {
        ConvArgs p = a;
        p.mode = FD_GATHER_CONV2D;
        p.K = 1; p.kh = p.kw = 1; p.sh = p.sw = 1; p.ph = p.pw = 0;
        p.Hout = d->Hin; p.Wout = d->Win;
        p.out_map = OUTMAP_UPSAMPLE;
        p.up_s = d->sh; p.up_dy = k / d->kw; p.up_dx = k % d->kw;
        // the tcgen05 arm handles a phase when the output-stationary kernel does (dY rows = the phase's pixels)
        p.n_in_cap = d->B * d->Hin * d->Win;
        p.in_split = d->d_in_split; p.out_split = d->d_out_split;
        const int pp = (d->precision == FD_PREC_BF16X3 && wgrad_os_supported(p)) ? FD_PREC_BF16X3 : FD_PREC_FP32;
        if (need) { *need = wgrad_ws_bytes(p, pp); return 0; }
        int rc = launch_wgrad(p, d_dw + (size_t)k * d->cin * d->cout, stream, pp, ws, ws_bytes);
        if (rc) return rc;
      }
      return 0;
    }
    default:
      return set_error(-1, "fd_conv_wgrad: unsupported gather mode %d", d->mode);
  }
}

size_t fd_bn_workspace_bytes(int C) {
  if (C < 1) return 0;
  return fd::red_partial_bytes(C) + sizeof(float) * 2 * (size_t)C + 256;
}

int fd_bn_train_stats(const float* d_x, int x_stride, int C, const int32_t* d_n, int64_t n_cap, float eps,
                      float momentum, const float* d_gamma, const float* d_beta, float* d_running_mean,
                      float* d_running_var, float* d_mean, float* d_invstd, float* d_scale, float* d_shift,
                      void* d_workspace, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_x && d_mean && d_invstd && d_scale && d_shift && d_workspace && C >= 1 && x_stride >= C && n_cap >= 1,
             "fd_bn_train_stats: bad argument");
  RedArgs r{};
  r.mode = RED_STATS; r.x = d_x; r.x_stride = x_stride; r.C = C; r.cp = red_cp(C);
  r.d_n = d_n; r.n_cap = n_cap; r.partial = (double*)d_workspace;
  const int G = red_blocks(n_cap, C);
  launch_red_partial(r, G, stream);
  FD_LAUNCHED();
  FinArgs f{};
  f.mode = RED_STATS; f.C = C; f.G = G; f.partial = r.partial; f.d_n = d_n; f.n_cap = n_cap;
  f.eps = eps; f.momentum = momentum; f.gamma = d_gamma; f.beta = d_beta;
  f.running_mean = d_running_mean; f.running_var = d_running_var;
  f.mean = d_mean; f.invstd = d_invstd; f.scale = d_scale; f.shift = d_shift;
  red_finalize_kernel<<<ceil_div((int64_t)C * 32, RED_THREADS), RED_THREADS, 0, stream>>>(f);
  FD_LAUNCHED();
  return 0;
}

int fd_affine_act(const float* d_x, int x_stride, int C, const float* d_scale, const float* d_shift,
                  const float* d_res, int res_stride, int relu, float* d_y, int y_stride, void* d_y_split,
                  int split_ctot, const int32_t* d_n, int64_t n_cap, void* stream) {
  using namespace fd;
  FD_REQUIRE(!d_y_split || split_ctot >= C, "fd_affine_act: split_ctot < C");
  FD_REQUIRE(d_x && d_y && C >= 1 && x_stride >= C && y_stride >= C && (!d_res || res_stride >= C) && n_cap >= 0,
             "fd_affine_act: bad argument");
  if (n_cap == 0) return 0;
  if (C % 4 == 0 && rows_vec4(d_x, x_stride) && rows_vec4(d_y, y_stride) && rows_vec4(d_res, res_stride) &&
      rows_vec4(d_scale, 4) && rows_vec4(d_shift, 4) && (!d_y_split || ((((uintptr_t)d_y_split) & 7) == 0 && split_ctot % 4 == 0)))
    affine_act_vec4_kernel<<<persistent_grid(ceil_div(n_cap * (C / 4), 256), 8), 256, 0, (cudaStream_t)stream>>>(
        d_x, x_stride, C, d_scale, d_shift, d_res, res_stride, relu, d_y, y_stride, (unsigned short*)d_y_split, split_ctot,
        d_n, n_cap);
  else
    affine_act_kernel<<<persistent_grid(ceil_div(n_cap * C, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        d_x, x_stride, C, d_scale, d_shift, d_res, res_stride, relu, d_y, y_stride, (unsigned short*)d_y_split, split_ctot,
        d_n, n_cap);
  FD_LAUNCHED();
  return 0;
}

int fd_bn_backward(const float* d_dy, int dy_stride, const float* d_y, int y_stride, int relu, const float* d_x,
                   int x_stride, int C, const int32_t* d_n, int64_t n_cap, const float* d_mean,
                   const float* d_invstd, const float* d_gamma, float* d_dx, int dx_stride, void* d_dx_split,
                   float* d_dres, int dres_stride, float* d_dgamma, float* d_dbeta, void* d_workspace, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_dy && d_x && d_mean && d_invstd && d_dx && d_workspace && C >= 1 && n_cap >= 1,
             "fd_bn_backward: bad argument");
  FD_REQUIRE(!relu || d_y, "fd_bn_backward: relu needs the forward output y");
  FD_REQUIRE(dy_stride >= C && x_stride >= C && dx_stride >= C && (!relu || y_stride >= C) &&
                 (!d_dres || dres_stride >= C), "fd_bn_backward: bad stride");
  RedArgs r{};
  r.mode = RED_BN_BWD; r.x = d_x; r.x_stride = x_stride; r.dy = d_dy; r.dy_stride = dy_stride;
  r.y = d_y; r.y_stride = y_stride; r.relu = relu; r.mean = d_mean; r.invstd = d_invstd;
  r.C = C; r.cp = red_cp(C); r.d_n = d_n; r.n_cap = n_cap; r.partial = (double*)d_workspace;
  const int G = red_blocks(n_cap, C);
  launch_red_partial(r, G, stream);
  FD_LAUNCHED();
  float* c1 = (float*)((char*)d_workspace + red_partial_bytes(C));
  float* c2 = c1 + C;
  FinArgs f{};
  f.mode = RED_BN_BWD; f.C = C; f.G = G; f.partial = r.partial; f.d_n = d_n; f.n_cap = n_cap;
  f.dgamma = d_dgamma; f.dbeta = d_dbeta; f.c1 = c1; f.c2 = c2;
  red_finalize_kernel<<<ceil_div((int64_t)C * 32, RED_THREADS), RED_THREADS, 0, stream>>>(f);
  FD_LAUNCHED();
  if (C % 4 == 0 && rows_vec4(d_dy, dy_stride) && (!relu || rows_vec4(d_y, y_stride)) && rows_vec4(d_x, x_stride) &&
      rows_vec4(d_dx, dx_stride) && rows_vec4(d_dres, dres_stride) && (((uintptr_t)d_dx_split) & 7) == 0)
    bn_bwd_apply_vec4_kernel<<<persistent_grid(ceil_div(n_cap * (C / 4), 256), 8), 256, 0, stream>>>(
        d_dy, dy_stride, d_y, y_stride, relu, d_x, x_stride, C, d_mean, d_invstd, d_gamma, c1, c2, d_dx, dx_stride,
        d_dres, dres_stride, (unsigned short*)d_dx_split, d_n, n_cap);
  else
    bn_bwd_apply_kernel<<<persistent_grid(ceil_div(n_cap * C, 256), 8), 256, 0, stream>>>(
        d_dy, dy_stride, d_y, y_stride, relu, d_x, x_stride, C, d_mean, d_invstd, d_gamma, c1, c2, d_dx, dx_stride,
        d_dres, dres_stride, (unsigned short*)d_dx_split, d_n, n_cap);
  FD_LAUNCHED();
  return 0;
}

int fd_col_sum(const float* d_x, int x_stride, int C, const int32_t* d_n, int64_t n_cap, float* d_out,
               void* d_workspace, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_x && d_out && d_workspace && C >= 1 && x_stride >= C && n_cap >= 1, "fd_col_sum: bad argument");
  RedArgs r{};
  r.mode = RED_COLSUM; r.x = d_x; r.x_stride = x_stride; r.C = C; r.cp = red_cp(C);
  r.d_n = d_n; r.n_cap = n_cap; r.partial = (double*)d_workspace;
  const int G = red_blocks(n_cap, C);
  launch_red_partial(r, G, stream);
  FD_LAUNCHED();
  FinArgs f{};
  f.mode = RED_COLSUM; f.C = C; f.G = G; f.partial = r.partial; f.d_n = d_n; f.n_cap = n_cap; f.out = d_out;
  red_finalize_kernel<<<ceil_div((int64_t)C * 32, RED_THREADS), RED_THREADS, 0, stream>>>(f);
  FD_LAUNCHED();
  return 0;
}

int fd_add_rows(float* d_dst, int dst_stride, const float* d_src, int src_stride, int C, const int32_t* d_n,
                int64_t n_cap, void* stream) {
  using namespace fd;
  FD_REQUIRE(d_dst && d_src && C >= 1 && dst_stride >= C && src_stride >= C && n_cap >= 0, "fd_add_rows: bad argument");
  if (n_cap == 0) return 0;
  add_rows_kernel<<<persistent_grid(ceil_div(n_cap * C, 256), 8), 256, 0, (cudaStream_t)stream>>>(
      d_dst, dst_stride, d_src, src_stride, C, d_n, n_cap);
  FD_LAUNCHED();
  return 0;
}

int fd_rows_to_bev(const float* d_rows, int row_stride, int C, const int32_t* d_coords4, const int32_t* d_n,
                   int n_cap, int B, int D, int H, int W, float* d_bev, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_rows && d_coords4 && d_bev && C >= 1 && row_stride >= C && B >= 1 && D >= 1 && H >= 1 && W >= 1 &&
                 n_cap >= 0, "fd_rows_to_bev: bad argument");
  FD_CUDA(cudaMemsetAsync(d_bev, 0, sizeof(float) * (size_t)B * H * W * C * D, stream));
  if (n_cap == 0) return 0;
  bev_rows_kernel<true><<<persistent_grid(ceil_div((int64_t)n_cap * C, 256), 8), 256, 0, stream>>>(
      const_cast<float*>(d_rows), row_stride, C, (const int4*)d_coords4, d_n, n_cap, D, H, W, d_bev);
  FD_LAUNCHED();
  return 0;
}

int fd_bev_to_rows(const float* d_bev, int C, const int32_t* d_coords4, const int32_t* d_n, int n_cap, int B, int D,
                   int H, int W, float* d_rows, int row_stride, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_rows && d_coords4 && d_bev && C >= 1 && row_stride >= C && B >= 1 && D >= 1 && H >= 1 && W >= 1 &&
                 n_cap >= 0, "fd_bev_to_rows: bad argument");
  if (n_cap == 0) return 0;
  bev_rows_kernel<false><<<persistent_grid(ceil_div((int64_t)n_cap * C, 256), 8), 256, 0, stream>>>(
      d_rows, row_stride, C, (const int4*)d_coords4, d_n, n_cap, D, H, W, const_cast<float*>(d_bev));
  FD_LAUNCHED();
  return 0;
}

int fd_center_head_loss_backward(const float* d_hm, float* d_ghm, int64_t hm_sb, int64_t hm_sc, int64_t hm_ssp,
                                 const float* d_hm_target, int B, int C, int H, int W, const int64_t* d_ind,
                                 const uint8_t* d_mask, const int64_t* d_cat, int M, int T, int NC,
                                 const float* const* d_pred_ptr, float* const* d_gpred_ptr,
                                 const int64_t* d_pred_sb, const int64_t* d_pred_ssp,
                                 const float* const* d_tgt_ptr, int tgt_dim, const int32_t* d_tgt_sel,
                                 const float* d_code_w, const float* d_code_w_forecast, float weight,
                                 const float* d_gscale, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_hm && d_ghm && d_hm_target && d_ind && d_mask && d_cat && d_pred_ptr && d_gpred_ptr && d_pred_sb &&
                 d_pred_ssp && d_tgt_ptr && d_tgt_sel && d_code_w && d_code_w_forecast,
             "fd_center_head_loss_backward: null argument");
  FD_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && M >= 1 && T >= 1 && NC >= 1 && tgt_dim >= 1,
             "fd_center_head_loss_backward: bad shape");
  LossBwdArgs a{};
  a.hm = d_hm; a.ghm = d_ghm; a.hm_sb = hm_sb; a.hm_sc = hm_sc; a.hm_ssp = hm_ssp;
  a.gt = d_hm_target; a.B = B; a.C = C; a.HW = H * W;
  a.ind = (const long long*)d_ind; a.mask = d_mask; a.cat = (const long long*)d_cat; a.M = M; a.T = T; a.NC = NC;
  a.pred_ptr = d_pred_ptr; a.gpred_ptr = d_gpred_ptr;
  a.pred_sb = (const long long*)d_pred_sb; a.pred_ssp = (const long long*)d_pred_ssp;
  a.tgt_ptr = d_tgt_ptr; a.tgt_dim = tgt_dim; a.tgt_sel = d_tgt_sel;
  a.code_w = d_code_w; a.code_w_forecast = d_code_w_forecast; a.weight = weight; a.gscale = d_gscale;
  focal_grad_kernel<<<persistent_grid(ceil_div((int64_t)B * C * H * W, 256), 4), 256, 0, stream>>>(a);
  FD_LAUNCHED();
  object_grad_kernel<<<persistent_grid(ceil_div((int64_t)B * M, 256), 4), 256, 0, stream>>>(a);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
