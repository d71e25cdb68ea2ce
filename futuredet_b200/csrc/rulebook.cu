// Rulebook construction for SubMConv3d / SparseConv3d on sm_100a.
//
// spconv 1.x (external to the reference; call sites det3d/models/backbones/scn.py:98-146)
// builds `indice_pairs` with a dense B*D*H*W int32 grid plus a thrust sort/unique per
// strided layer.  Here:
//   * coordinate -> row lookups go through a small open-addressing hash (L2 resident);
//   * the active output set of a strided conv is a *bitmap* over the output grid
//     (1 bit per cell, 1.4 MB for 21x720x720) + a popcount scan, which yields the
//     ascending-linear-index order of spconv's sort+unique with no sort at all;
//   * the rulebook is stored in gather form nbr[k][o] (input row or -1), which is what
//     the output-stationary implicit-GEMM kernels consume -- no scatter atomics, so the
//     convolution is deterministic.  fd_rulebook_to_pairs() exports spconv's layout.
#include <stdlib.h>

#include "common.cuh"
#include "scan.cuh"

namespace fd {

struct Shape3 { int d, h, w; };
struct Conv3Geom { int k[3], s[3], p[3]; };

__device__ __forceinline__ long long lin_key(int b, int z, int y, int x, Shape3 sh) {
  return (((long long)b * sh.d + z) * sh.h + y) * sh.w + x;
}

// One 64-bit word per hash entry: (linear coordinate << 32) | row, 0xFFFF...F = empty.  A lookup is a single
// 8-byte load (L2 resident table), and the 27 probes of a row are issued back to back before any is consumed.
constexpr unsigned long long kEmptyEntry = ~0ull;

__global__ void __launch_bounds__(256)
coord_index_insert(const int4* __restrict__ coords, const int32_t* __restrict__ d_n, int n_cap, Shape3 sh,
                   unsigned long long* __restrict__ table, uint32_t mask) {
  int n = d_n ? min(*d_n, n_cap) : n_cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    const uint32_t key = (uint32_t)lin_key(c.x, c.y, c.z, c.w, sh);
    const unsigned long long entry = ((unsigned long long)key << 32) | (uint32_t)i;
    uint32_t h = hash64(key) & mask;
    while (true) {
      unsigned long long prev = atomicCAS(&table[h], kEmptyEntry, entry);
      if (prev == kEmptyEntry || (uint32_t)(prev >> 32) == key) break;     // duplicates: first writer wins
      h = (h + 1) & mask;
    }
  }
}

__device__ __forceinline__ int coord_index_resolve(const unsigned long long* __restrict__ table, uint32_t mask,
                                                   uint32_t key, uint32_t h, unsigned long long e) {
  while (true) {
    if ((uint32_t)(e >> 32) == key) return (int)(uint32_t)e;
    if (e == kEmptyEntry) return -1;
    h = (h + 1) & mask;
    e = table[h];
  }
}

// t = q * s exactly?  Strides are 1 or 2 in every layer of the backbone: no integer division on those paths.
__device__ __forceinline__ bool exact_div(int t, int s, int& q) {
  if (s == 1) { q = t; return true; }
  if (s == 2) { q = t >> 1; return (t & 1) == 0; }
  q = t / s;
  return q * s == t;
}

// mark every output cell reachable from an active input: out = (in + p - k) / s when divisible
__global__ void __launch_bounds__(256)
outset_mark(const int4* __restrict__ coords, const int32_t* __restrict__ d_n, int n_cap, Conv3Geom g,
            Shape3 osh, uint32_t* __restrict__ bitmap) {
  int n = min(*d_n, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    for (int kz = 0; kz < g.k[0]; ++kz) {
      int tz = c.y + g.p[0] - kz;
      int oz;
      if (tz < 0 || !exact_div(tz, g.s[0], oz)) continue;
      if (oz >= osh.d) continue;
      for (int ky = 0; ky < g.k[1]; ++ky) {
        int ty = c.z + g.p[1] - ky;
        int oy;
        if (ty < 0 || !exact_div(ty, g.s[1], oy)) continue;
        if (oy >= osh.h) continue;
        for (int kx = 0; kx < g.k[2]; ++kx) {
          int tx = c.w + g.p[2] - kx;
          int ox;
          if (tx < 0 || !exact_div(tx, g.s[2], ox)) continue;
          if (ox >= osh.w) continue;
          long long key = lin_key(c.x, oz, oy, ox, osh);
          // up to 27 inputs mark the same output cell: look before the atomic (a stale 0 only costs a redundant atomicOr)
          if (!(__ldcg(&bitmap[key >> 5]) & (1u << (key & 31)))) atomicOr(&bitmap[key >> 5], 1u << (key & 31));
        }
      }
    }
  }
}

// enumerate set bits in ascending order -> out coords
__global__ void __launch_bounds__(256)
outset_emit(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ prefix, int64_t words, Shape3 osh,
            int4* __restrict__ out_coords, int n_out_cap) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words;
       w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t bits = bitmap[w];
    if (!bits) continue;
    int row = prefix[w];
    // the word's first cell is decoded once with 32-bit divisions (B*D*H*W < 2^32); its bits are consecutive cells
    uint32_t t = (uint32_t)w * 32u;
    const int x0 = (int)(t % (uint32_t)osh.w); t /= (uint32_t)osh.w;
    const int y0 = (int)(t % (uint32_t)osh.h); t /= (uint32_t)osh.h;
    const int z0 = (int)(t % (uint32_t)osh.d);
    const int b0 = (int)(t / (uint32_t)osh.d);
    while (bits) {
      int j = __ffs(bits) - 1;
      bits &= bits - 1;
      if (row < n_out_cap) {
        int x = x0 + j, y = y0, z = z0, b = b0;
        while (x >= osh.w) { x -= osh.w; ++y; }
        while (y >= osh.h) { y -= osh.h; ++z; }
        while (z >= osh.d) { z -= osh.d; ++b; }
        out_coords[row] = make_int4(b, z, y, x);
      }
      ++row;
    }
  }
}

__global__ void clamp_count(int32_t* n, int cap) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && *n > cap) *n = cap;
}

// 9-bit signature of a row's neighbour pattern `m` (bit k: neighbour through kernel offset k): bit g = some neighbour through
// the offsets [g*gs, (g+1)*gs) -- for 3x3x3, gs = 3: the (dz, dy) line g.  Sort key of fd_rulebook_sort_rows.
__device__ __forceinline__ uint32_t pattern_key(uint32_t m, int K, int gs) {
  uint32_t key = 0;
  for (int g = 0; g * gs < K; ++g) key |= (uint32_t)(((m >> (g * gs)) & ((1u << gs) - 1u)) != 0) << g;
  return key;
}

// nbr[k][o]: one thread per output row; kernel offsets in groups of kx-rows so that the first probes of a group
// are independent loads in flight together (the lookups are pure latency otherwise)
__global__ void __launch_bounds__(256)
neighbors_kernel(const int4* __restrict__ out_coords, const int32_t* __restrict__ d_n, int n_cap,
                 const unsigned long long* __restrict__ table, uint32_t mask, Shape3 ish, Conv3Geom g,
                 int* __restrict__ nbr, int nbr_stride, int* __restrict__ pair_num, uint32_t* __restrict__ tile_mask,
                 uint16_t* __restrict__ row_key, int K, int key_gs) {
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int lane = threadIdx.x & 31;
  const int n_round = (n + 31) & ~31;  // keep warps converged for the ballots
  constexpr int MAXG = 9;              // probes in flight per thread (ky x kx up to 3x3)
  const int gsz = g.k[1] * g.k[2];
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_round; o += gridDim.x * blockDim.x) {
    const bool live = o < n;
    int4 c = live ? out_coords[o] : make_int4(0, 0, 0, 0);
    const int z0 = c.y * g.s[0] - g.p[0], y0 = c.z * g.s[1] - g.p[1], x0 = c.w * g.s[2] - g.p[2];
    uint32_t my_mask = 0, my_bits = 0;
    for (int kz = 0; kz < g.k[0]; ++kz) {
      const int z = z0 + kz;
      const bool zok = live && z >= 0 && z < ish.d;
      if (gsz <= MAXG) {
        uint32_t key[MAXG], h[MAXG];
        unsigned long long e[MAXG];
        bool ok[MAXG];
#pragma unroll
        for (int q = 0; q < MAXG; ++q) {
          ok[q] = false;
          if (q < gsz) {
            const int ky = q / g.k[2], kx = q - ky * g.k[2];
            const int y = y0 + ky, x = x0 + kx;
            ok[q] = zok && y >= 0 && y < ish.h && x >= 0 && x < ish.w;
            key[q] = (uint32_t)lin_key(c.x, z, y, x, ish);
            h[q] = hash64(key[q]) & mask;
            e[q] = ok[q] ? __ldg(&table[h[q]]) : kEmptyEntry;
          }
        }
#pragma unroll
        for (int q = 0; q < MAXG; ++q) {
          if (q < gsz) {
            const int k = kz * gsz + q;
            int r = ok[q] ? coord_index_resolve(table, mask, key[q], h[q], e[q]) : -1;
            if (live) nbr[(size_t)k * nbr_stride + o] = r;
            my_bits |= (uint32_t)(r >= 0) << (k & 31);
            if (__any_sync(0xffffffffu, r >= 0)) my_mask |= 1u << (k & 31);
          }
        }
      } else {
        for (int q = 0; q < gsz; ++q) {
          const int ky = q / g.k[2], kx = q - ky * g.k[2];
          const int y = y0 + ky, x = x0 + kx, k = kz * gsz + q;
          int r = -1;
          if (zok && y >= 0 && y < ish.h && x >= 0 && x < ish.w) {
            const uint32_t key = (uint32_t)lin_key(c.x, z, y, x, ish);
            const uint32_t hh = hash64(key) & mask;
            r = coord_index_resolve(table, mask, key, hh, table[hh]);
          }
          if (live) nbr[(size_t)k * nbr_stride + o] = r;
          my_bits |= (uint32_t)(r >= 0) << (k & 31);
          if (__any_sync(0xffffffffu, r >= 0)) my_mask |= 1u << (k & 31);
        }
      }
    }
    if (row_key && live) row_key[o] = (uint16_t)pattern_key(my_bits, K, key_gs);
    // per 128-row tile activity mask (bit k: some row of the tile has a neighbour through offset k); consumed by
    // the implicit-GEMM kernels to skip kernel offsets that are empty for a whole tile
    if (tile_mask && lane == 0 && my_mask) atomicOr(&tile_mask[o >> 7], my_mask);
  }
}

// Same neighbour search against a BITMAP index: the active set of a strided conv's output grid is already available
// as (1 bit per cell, exclusive popcount prefix per word) from fd_rulebook_out_coords, and its rows are in ascending
// linear order, so  row(cell) = prefix[word] + popc(bits below).  The three x-neighbours of a (kz,ky) pair share one
// or two words, so a row needs ~9 word loads with high locality instead of 27 random hash probes.
__global__ void __launch_bounds__(256)
neighbors_bitmap_kernel(const int4* __restrict__ out_coords, const int32_t* __restrict__ d_n, int n_cap,
                        const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ prefix, Shape3 ish, Conv3Geom g,
                        int* __restrict__ nbr, int nbr_stride, int* __restrict__ pair_num, uint32_t* __restrict__ tile_mask,
                        uint16_t* __restrict__ row_key, int K, int key_gs) {
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int lane = threadIdx.x & 31;
  const int n_round = (n + 31) & ~31;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_round; o += gridDim.x * blockDim.x) {
    const bool live = o < n;
    int4 c = live ? out_coords[o] : make_int4(0, 0, 0, 0);
    const int z0 = c.y * g.s[0] - g.p[0], y0 = c.z * g.s[1] - g.p[1], x0 = c.w * g.s[2] - g.p[2];
    uint32_t my_mask = 0, my_bits = 0;
    int k = 0;
    for (int kz = 0; kz < g.k[0]; ++kz) {
      const int z = z0 + kz;
      for (int ky = 0; ky < g.k[1]; ++ky) {
        const int y = y0 + ky;
        const bool rowok = live && z >= 0 && z < ish.d && y >= 0 && y < ish.h;
        const long long lin0 = lin_key(c.x, z, y, 0, ish);       // cell (x = 0) of this line
        long long cached_w = -1;
        uint32_t bits = 0;
        int base = 0;
        for (int kx = 0; kx < g.k[2]; ++kx, ++k) {
          const int x = x0 + kx;
          int r = -1;
          if (rowok && x >= 0 && x < ish.w) {
            const long long lin = lin0 + x;
            const long long w = lin >> 5;
            if (w != cached_w) { cached_w = w; bits = __ldg(bitmap + w); base = -1; }
            const uint32_t bit = 1u << (lin & 31);
            if (bits & bit) {
              if (base < 0) base = __ldg(prefix + w);
              r = base + __popc(bits & (bit - 1));
            }
          }
          if (live) nbr[(size_t)k * nbr_stride + o] = r;
          my_bits |= (uint32_t)(r >= 0) << (k & 31);
          if (__any_sync(0xffffffffu, r >= 0)) my_mask |= 1u << (k & 31);
        }
      }
    }
    if (row_key && live) row_key[o] = (uint16_t)pattern_key(my_bits, K, key_gs);
    if (tile_mask && lane == 0 && my_mask) atomicOr(&tile_mask[o >> 7], my_mask);
  }
}

// Strided (SparseConv3d) rulebooks, input-stationary: an active input reaches at most prod(ceil(k/s)) outputs (8 for
// k = 3, s = 2; 3.4 on average), while an output has prod(k) = 27 candidate inputs of which most are empty.  One thread
// per INPUT row enumerates its (output cell, offset) pairs, finds the output row through the bitmap + popcount prefix
// that fd_rulebook_out_coords left behind (row = rank, ascending linear order) and writes nbr[k][out] = in.  Every
// (k, out) slot has exactly one candidate input cell, so no two threads write the same slot: deterministic, same table
// as the output-stationary search with ~8x fewer lookups.  The caller pre-fills nbr with -1.
__global__ void __launch_bounds__(256)
neighbors_scatter_kernel(const int4* __restrict__ in_coords, const int32_t* __restrict__ d_n_in, int n_in_cap, Conv3Geom g,
                         Shape3 osh, const uint32_t* __restrict__ out_bitmap, const int32_t* __restrict__ out_prefix,
                         int n_out_cap, int* __restrict__ nbr, int nbr_stride, uint32_t* __restrict__ tile_mask) {
  const int n = min(*d_n_in, n_in_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 c = in_coords[i];
    for (int kz = 0; kz < g.k[0]; ++kz) {
      const int tz = c.y + g.p[0] - kz;
      int oz;
      if (tz < 0 || !exact_div(tz, g.s[0], oz)) continue;
      if (oz >= osh.d) continue;
      for (int ky = 0; ky < g.k[1]; ++ky) {
        const int ty = c.z + g.p[1] - ky;
        int oy;
        if (ty < 0 || !exact_div(ty, g.s[1], oy)) continue;
        if (oy >= osh.h) continue;
        for (int kx = 0; kx < g.k[2]; ++kx) {
          const int tx = c.w + g.p[2] - kx;
          int ox;
          if (tx < 0 || !exact_div(tx, g.s[2], ox)) continue;
          if (ox >= osh.w) continue;
          const long long key = lin_key(c.x, oz, oy, ox, osh);
          const uint32_t bits = __ldg(out_bitmap + (key >> 5));
          const int o = __ldg(out_prefix + (key >> 5)) + __popc(bits & ((1u << (key & 31)) - 1u));
          if (o >= n_out_cap) continue;                       // output set clamped to its capacity
          const int k = (kz * g.k[1] + ky) * g.k[2] + kx;
          nbr[(size_t)k * nbr_stride + o] = i;
          // tile activity bit: hundreds of pairs of a tile set the same bit -- look before the atomic
          if (tile_mask && !(__ldcg(&tile_mask[o >> 7]) & (1u << (k & 31)))) atomicOr(&tile_mask[o >> 7], 1u << (k & 31));
        }
      }
    }
  }
}

// nbr[k][o] = -1 for the live output rows only (the capacity of a strided level is several times its row count)
__global__ void __launch_bounds__(256)
nbr_fill_empty_kernel(int* __restrict__ nbr, int nbr_stride, int K, const int32_t* __restrict__ d_n_out, int n_out_cap) {
  const int n = min(*d_n_out, n_out_cap);
  const int n4 = (n + 3) >> 2;                              // rows in int4 units (table rows are 16-byte aligned)
  const long long total = (long long)K * n4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / n4), q = (int)(e - (long long)k * n4);
    reinterpret_cast<int4*>(nbr + (size_t)k * nbr_stride)[q] = make_int4(-1, -1, -1, -1);
  }
}

// pairs per kernel offset (spconv `indice_pair_num`): one block per offset, deterministic
__global__ void __launch_bounds__(256)
count_pairs_kernel(const int* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ d_n, int n_cap,
                   int* __restrict__ pair_num) {
  __shared__ int ws[8];
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int* row = nbr + (size_t)blockIdx.x * nbr_stride;
  int c = 0;
  for (int o = threadIdx.x; o < n; o += blockDim.x) c += row[o] >= 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    pair_num[blockIdx.x] = t;
  }
}

// ---- export to spconv layout ----------------------------------------------------
struct LoadValid {
  const int* nbr;
  int nbr_stride;
  const int32_t* d_n;
  int n_cap;
  __device__ __forceinline__ int operator()(int64_t i) const {
    int n = d_n ? min(*d_n, n_cap) : n_cap;
    return (int)(i % nbr_stride) < n && nbr[i] >= 0;
  }
};

__global__ void __launch_bounds__(256)
pairs_emit(const int* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ d_n, int n_cap,
           const int32_t* __restrict__ pos, int32_t* __restrict__ pairs, int pair_cap, int k) {
  int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int* row = nbr + (size_t)k * nbr_stride;
  const int32_t* prow = pos + (size_t)k * nbr_stride;
  int base = prow[0];
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    int i = row[o];
    if (i >= 0) {
      int q = prow[o] - base;
      if (q < pair_cap) {
        pairs[((size_t)k * 2 + 0) * pair_cap + q] = i;
        pairs[((size_t)k * 2 + 1) * pair_cap + q] = o;
      }
    }
  }
}


// ---- tile sorting ------------------------------------------------------------------------------
// An output-stationary tile of 128 consecutive rows pays for every kernel offset that ANY of its rows uses, and the rows of
// a LiDAR level use few and different ones (16 % / 39 % / 66 % of the 27 slots at the 16 / 32 / 64-channel levels): in
// voxelizer or raster order nearly every tile needs all 27 offsets.  Sorting the rows of a window by a 9-bit signature of
// their neighbour pattern (bit g: a neighbour through one of the offsets 3g..3g+2, i.e. on the (dz, dy) line g) puts rows
// with the same pattern into the same tiles, and the tile masks -- hence the K stages a tile gathers and multiplies --
// shrink to 0.37 / 0.65 / 0.80 of the unsorted count at those levels (0.28 / 0.42 / 0.62 on the strided layers between
// them).  A row's result does not depend on its tile (skipped stages only ever added exact zeros), so the sorted table gives
// bit-identical outputs; fd_conv_forward writes tile position j to row d_row_perm[j].
// Counting sort, three launches: per-CTA histogram + rank of each row among the rows of its CTA with the same key
// (shared-memory atomics; the order inside a bucket is arbitrary and irrelevant), one atomicAdd per (CTA, non-empty bucket)
// on the window's histogram to order the CTAs inside a bucket, scan of the window histograms, scatter of the table columns.
constexpr int RS_ROWS = 2048;          // rows per CTA (256 threads x 8)
constexpr int RS_BUCKETS = 512;

__device__ __forceinline__ uint32_t row_pattern(const int32_t* __restrict__ nbr, int nbr_stride, int K, int i) {
  uint32_t m = 0;
  for (int k = 0; k < K; ++k) m |= (uint32_t)(__ldg(nbr + (size_t)k * nbr_stride + i) >= 0) << k;
  return m;
}
__global__ void __launch_bounds__(256)
rowsort_count_kernel(const int32_t* __restrict__ nbr, int nbr_stride, int K, int gs, const int32_t* __restrict__ d_n,
                     int n_cap, int window, const uint16_t* __restrict__ keys_in, uint16_t* __restrict__ keys,
                     uint16_t* __restrict__ lrank, int32_t* __restrict__ blk_base, int32_t* __restrict__ win_hist) {
  __shared__ int hist[RS_BUCKETS];
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int row0 = blockIdx.x * RS_ROWS;
  if (row0 >= n) return;
  for (int b = threadIdx.x; b < RS_BUCKETS; b += 256) hist[b] = 0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ROWS / 256; ++j) {
    const int i = row0 + j * 256 + threadIdx.x;
    if (i < n) {
      // the neighbour kernels can hand the keys over (they have the row's pattern in registers): 2 bytes per row instead of
      // a pass over the K table columns
      const uint32_t key = keys_in ? (uint32_t)keys_in[i] : pattern_key(row_pattern(nbr, nbr_stride, K, i), K, gs);
      if (!keys_in) keys[i] = (uint16_t)key;
      lrank[i] = (uint16_t)atomicAdd(&hist[key], 1);
    }
  }
  __syncthreads();
  int32_t* wh = win_hist + (size_t)(row0 / window) * RS_BUCKETS;
  for (int b = threadIdx.x; b < RS_BUCKETS; b += 256) {
    const int c = hist[b];
    blk_base[(size_t)blockIdx.x * RS_BUCKETS + b] = c ? atomicAdd(&wh[b], c) : 0;
  }
}

// win_hist[w][b] -> first sorted position of bucket b of window w
__global__ void __launch_bounds__(RS_BUCKETS)
rowsort_scan_kernel(int32_t* __restrict__ win_hist, int window) {
  __shared__ int ws[RS_BUCKETS / 32];
  int32_t* h = win_hist + (size_t)blockIdx.x * RS_BUCKETS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = h[threadIdx.x];
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = lane < RS_BUCKETS / 32 ? ws[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane < RS_BUCKETS / 32) ws[lane] = s;
  }
  __syncthreads();
  h[threadIdx.x] = blockIdx.x * window + (warp ? ws[warp - 1] : 0) + inc - v;
}

// Scatter of the table columns.  A CTA first orders its rows by key in shared memory (local position = start of the key
// inside the CTA + the row's rank), so that consecutive local positions of one key map to consecutive sorted positions:
// the columns are then read in source order (coalesced), transposed through shared memory RS_KB columns at a time and
// written as runs of consecutive table entries instead of one 4-byte store per (row, offset) at a random place.
template <int RS_KB>                   // table columns staged per round
__global__ void __launch_bounds__(256)
rowsort_scatter_kernel(const int32_t* __restrict__ nbr, int nbr_stride, int K, const int32_t* __restrict__ d_n, int n_cap,
                       int window, const uint16_t* __restrict__ keys, const uint16_t* __restrict__ lrank,
                       const int32_t* __restrict__ blk_base, const int32_t* __restrict__ win_start,
                       int32_t* __restrict__ perm, int32_t* __restrict__ nbr_sorted, uint32_t* __restrict__ tile_mask) {
  constexpr int RPT = RS_ROWS / 256;                 // rows per thread
  __shared__ int base[RS_BUCKETS];                   // sorted position of the CTA's first row of each key
  __shared__ int lstart[RS_BUCKETS];                 // local position of the CTA's first row of each key
  __shared__ uint16_t skey[RS_ROWS];                 // key of local position lp
  extern __shared__ int stage_raw[];
  int (*stage)[RS_ROWS] = reinterpret_cast<int (*)[RS_ROWS]>(stage_raw);   // [RS_KB][RS_ROWS]
  __shared__ int wsum[8];
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const int row0 = blockIdx.x * RS_ROWS;
  if (row0 >= n) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = tid; b < RS_BUCKETS; b += 256) lstart[b] = 0;
  __syncthreads();
  int key_s[RPT], lp_s[RPT];                         // this thread's source rows: key, then local position
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int i = row0 + j * 256 + tid;
    key_s[j] = i < n ? (int)keys[i] : -1;
    if (key_s[j] >= 0) atomicAdd(&lstart[key_s[j]], 1);
  }
  __syncthreads();
  {  // exclusive scan of the 512 per-key counts (2 per thread)
    const int c0 = lstart[2 * tid], c1 = lstart[2 * tid + 1];
    int inc = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const int ex = woff + inc - c0 - c1;
    lstart[2 * tid] = ex;
    lstart[2 * tid + 1] = ex + c0;
  }
  const int32_t* ws = win_start + (size_t)(row0 / window) * RS_BUCKETS;
  for (int b = tid; b < RS_BUCKETS; b += 256) base[b] = ws[b] + blk_base[(size_t)blockIdx.x * RS_BUCKETS + b];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    lp_s[j] = -1;
    if (key_s[j] >= 0) {
      const int i = row0 + j * 256 + tid;
      lp_s[j] = lstart[key_s[j]] + (int)lrank[i];
      skey[lp_s[j]] = (uint16_t)key_s[j];
      stage[0][lp_s[j]] = i;
    }
  }
  __syncthreads();
  const int live = min(RS_ROWS, n - row0);
  int g_d[RPT];                                      // this thread's local positions tid + 256 j: sorted position
  uint32_t m_d[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int lp = j * 256 + tid;
    g_d[j] = -1;
    m_d[j] = 0;
    if (lp < live) {
      const int key = skey[lp];
      g_d[j] = base[key] + lp - lstart[key];
      perm[g_d[j]] = stage[0][lp];
    }
  }
  for (int k0 = 0; k0 < K; k0 += RS_KB) {
    __syncthreads();
    const int kn = min(RS_KB, K - k0);
    for (int kk = 0; kk < kn; ++kk) {
      const int32_t* col = nbr + (size_t)(k0 + kk) * nbr_stride + row0 + tid;
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        if (lp_s[j] >= 0) stage[kk][lp_s[j]] = __ldg(col + j * 256);
    }
    __syncthreads();
    for (int kk = 0; kk < kn; ++kk) {
      int32_t* col = nbr_sorted + (size_t)(k0 + kk) * nbr_stride;
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        if (g_d[j] >= 0) {
          const int v = stage[kk][j * 256 + tid];
          col[g_d[j]] = v;
          m_d[j] |= (uint32_t)(v >= 0) << (k0 + kk);
        }
    }
  }
  if (tile_mask) {
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (g_d[j] >= 0 && (m_d[j] & ~__ldcg(&tile_mask[g_d[j] >> 7]))) atomicOr(&tile_mask[g_d[j] >> 7], m_d[j]);
  }
}

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace fd

extern "C" {

int fd_rulebook_count_pairs(const int32_t* d_nbr, int nbr_stride, const int32_t* d_n_out, int n_out_cap, int K,
                            int32_t* d_pair_num, void* stream) {
  using namespace fd;
  FD_REQUIRE(d_nbr && d_pair_num && K >= 1 && nbr_stride >= n_out_cap, "fd_rulebook_count_pairs: bad argument");
  count_pairs_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(d_nbr, nbr_stride, d_n_out, n_out_cap, d_pair_num);
  FD_LAUNCHED();
  return 0;
}

int fd_coord_index_build(const int32_t* d_coords4, const int32_t* d_n, int n_cap, int B, const int32_t* shape3,
                         uint64_t* d_table, int64_t cap, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_coords4 && shape3 && d_table && B >= 1, "fd_coord_index_build: null argument");
  FD_REQUIRE((int64_t)B * shape3[0] * shape3[1] * shape3[2] < 0xFFFFFFFFLL,
             "fd_coord_index_build: B*D*H*W = %lld does not fit the 32-bit key of the packed index",
             (long long)B * shape3[0] * shape3[1] * shape3[2]);
  FD_REQUIRE(is_pow2(cap) && cap >= 2 * (int64_t)n_cap && cap <= (1LL << 31),
             "fd_coord_index_build: cap %lld must be a power of two >= 2*n_cap (%d)", (long long)cap, n_cap);
  FD_REQUIRE(((uintptr_t)d_coords4 & 15) == 0, "fd_coord_index_build: coords must be 16-byte aligned");
  FD_CUDA(cudaMemsetAsync(d_table, 0xff, sizeof(uint64_t) * cap, stream));
  if (n_cap <= 0) return 0;
  Shape3 sh{shape3[0], shape3[1], shape3[2]};
  coord_index_insert<<<persistent_grid(ceil_div(n_cap, 256), 8), 256, 0, stream>>>(
      (const int4*)d_coords4, d_n, n_cap, sh, (unsigned long long*)d_table, (uint32_t)(cap - 1));
  FD_LAUNCHED();
  return 0;
}

int fd_rulebook_out_coords(const int32_t* d_in_coords4, const int32_t* d_n_in, int n_in_cap, int B,
                           const int32_t* in_shape3, const int32_t* ksize3, const int32_t* stride3,
                           const int32_t* pad3, const int32_t* out_shape3, uint32_t* d_bitmap,
                           int32_t* d_wordprefix, void* d_scan_tmp, int32_t* d_out_coords4, int n_out_cap,
                           int32_t* d_n_out, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_in_coords4 && d_n_in && in_shape3 && ksize3 && stride3 && pad3 && out_shape3 && d_bitmap &&
                 d_wordprefix && d_scan_tmp && d_out_coords4 && d_n_out,
             "fd_rulebook_out_coords: null argument");
  Conv3Geom g;
  for (int j = 0; j < 3; ++j) {
    g.k[j] = ksize3[j]; g.s[j] = stride3[j]; g.p[j] = pad3[j];
    FD_REQUIRE(g.k[j] >= 1 && g.s[j] >= 1 && g.p[j] >= 0, "fd_rulebook_out_coords: bad conv geometry");
    int expect = (in_shape3[j] + 2 * g.p[j] - g.k[j]) / g.s[j] + 1;
    FD_REQUIRE(out_shape3[j] == expect, "fd_rulebook_out_coords: out_shape[%d]=%d, expected %d", j,
               out_shape3[j], expect);
  }
  Shape3 osh{out_shape3[0], out_shape3[1], out_shape3[2]};
  int64_t cells = (int64_t)B * osh.d * osh.h * osh.w;
  FD_REQUIRE(cells > 0 && cells < (1LL << 36), "fd_rulebook_out_coords: output grid too large");
  int64_t words = (cells + 31) / 32;
  FD_CUDA(cudaMemsetAsync(d_bitmap, 0, sizeof(uint32_t) * words, stream));
  if (n_in_cap > 0) {
    outset_mark<<<persistent_grid(ceil_div(n_in_cap, 256), 8), 256, 0, stream>>>(
        (const int4*)d_in_coords4, d_n_in, n_in_cap, g, osh, d_bitmap);
    FD_LAUNCHED();
  }
  int rc = exclusive_scan_popc(d_bitmap, d_wordprefix, words, d_n_out, d_scan_tmp, stream);
  if (rc) return rc;
  outset_emit<<<persistent_grid(ceil_div(words, 256), 8), 256, 0, stream>>>(
      d_bitmap, d_wordprefix, words, osh, (int4*)d_out_coords4, n_out_cap);
  FD_LAUNCHED();
  clamp_count<<<1, 32, 0, stream>>>(d_n_out, n_out_cap);
  FD_LAUNCHED();
  return 0;
}

int fd_rulebook_neighbors(const int32_t* d_out_coords4, const int32_t* d_n_out, int n_out_cap,
                          const uint64_t* d_in_table, int64_t in_cap,
                          const int32_t* in_shape3, const int32_t* ksize3, const int32_t* stride3,
                          const int32_t* pad3, int32_t* d_nbr, int nbr_stride, int32_t* d_pair_num,
                          uint32_t* d_tile_mask, uint16_t* d_row_key, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_out_coords4 && d_in_table && in_shape3 && ksize3 && stride3 && pad3 && d_nbr,
             "fd_rulebook_neighbors: null argument");
  FD_REQUIRE(is_pow2(in_cap), "fd_rulebook_neighbors: in_cap must be a power of two");
  FD_REQUIRE(nbr_stride >= n_out_cap, "fd_rulebook_neighbors: nbr_stride < n_out_cap");
  Conv3Geom g;
  for (int j = 0; j < 3; ++j) { g.k[j] = ksize3[j]; g.s[j] = stride3[j]; g.p[j] = pad3[j]; }
  const int K = g.k[0] * g.k[1] * g.k[2];
  FD_REQUIRE(K >= 1 && K <= 343, "fd_rulebook_neighbors: kernel volume %d unsupported", K);
  FD_REQUIRE(!d_tile_mask || K <= 32, "fd_rulebook_neighbors: tile masks support at most 32 kernel offsets");
  if (d_pair_num) FD_CUDA(cudaMemsetAsync(d_pair_num, 0, sizeof(int32_t) * K, stream));
  if (d_tile_mask) FD_CUDA(cudaMemsetAsync(d_tile_mask, 0, sizeof(uint32_t) * (size_t)ceil_div(n_out_cap > 0 ? n_out_cap : 1, 128), stream));
  if (n_out_cap <= 0) return 0;
  Shape3 ish{in_shape3[0], in_shape3[1], in_shape3[2]};
  neighbors_kernel<<<persistent_grid(ceil_div(n_out_cap, 256), 8), 256, 0, stream>>>(
      (const int4*)d_out_coords4, d_n_out, n_out_cap, (const unsigned long long*)d_in_table,
      (uint32_t)(in_cap - 1), ish, g, d_nbr, nbr_stride, d_pair_num, d_tile_mask, K <= 32 ? d_row_key : nullptr, K,
      ceil_div(K, 9));
  FD_LAUNCHED();
  if (d_pair_num) return fd_rulebook_count_pairs(d_nbr, nbr_stride, d_n_out, n_out_cap, K, d_pair_num, stream_);
  return 0;
}

int fd_rulebook_neighbors_bitmap(const int32_t* d_out_coords4, const int32_t* d_n_out, int n_out_cap,
                                 const uint32_t* d_in_bitmap, const int32_t* d_in_wordprefix, const int32_t* in_shape3,
                                 const int32_t* ksize3, const int32_t* stride3, const int32_t* pad3, int32_t* d_nbr,
                                 int nbr_stride, int32_t* d_pair_num, uint32_t* d_tile_mask, uint16_t* d_row_key,
                                 void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_out_coords4 && d_in_bitmap && d_in_wordprefix && in_shape3 && ksize3 && stride3 && pad3 && d_nbr,
             "fd_rulebook_neighbors_bitmap: null argument");
  FD_REQUIRE(nbr_stride >= n_out_cap, "fd_rulebook_neighbors_bitmap: nbr_stride < n_out_cap");
  Conv3Geom g;
  for (int j = 0; j < 3; ++j) { g.k[j] = ksize3[j]; g.s[j] = stride3[j]; g.p[j] = pad3[j]; }
  const int K = g.k[0] * g.k[1] * g.k[2];
  FD_REQUIRE(K >= 1 && K <= 343, "fd_rulebook_neighbors_bitmap: kernel volume %d unsupported", K);
  FD_REQUIRE(!d_tile_mask || K <= 32, "fd_rulebook_neighbors_bitmap: tile masks support at most 32 kernel offsets");
  if (d_pair_num) FD_CUDA(cudaMemsetAsync(d_pair_num, 0, sizeof(int32_t) * K, stream));
  if (d_tile_mask) FD_CUDA(cudaMemsetAsync(d_tile_mask, 0, sizeof(uint32_t) * (size_t)ceil_div(n_out_cap > 0 ? n_out_cap : 1, 128), stream));
  if (n_out_cap <= 0) return 0;
  Shape3 ish{in_shape3[0], in_shape3[1], in_shape3[2]};
  neighbors_bitmap_kernel<<<persistent_grid(ceil_div(n_out_cap, 256), 8), 256, 0, stream>>>(
      (const int4*)d_out_coords4, d_n_out, n_out_cap, d_in_bitmap, d_in_wordprefix, ish, g, d_nbr, nbr_stride,
      d_pair_num, d_tile_mask, K <= 32 ? d_row_key : nullptr, K, ceil_div(K, 9));
  FD_LAUNCHED();
  if (d_pair_num) return fd_rulebook_count_pairs(d_nbr, nbr_stride, d_n_out, n_out_cap, K, d_pair_num, stream_);
  return 0;
}

int fd_rulebook_neighbors_scatter(const int32_t* d_in_coords4, const int32_t* d_n_in, int n_in_cap,
                                  const uint32_t* d_out_bitmap, const int32_t* d_out_wordprefix, const int32_t* out_shape3,
                                  const int32_t* ksize3, const int32_t* stride3, const int32_t* pad3,
                                  const int32_t* d_n_out, int n_out_cap, int32_t* d_nbr, int nbr_stride,
                                  uint32_t* d_tile_mask, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_in_coords4 && d_n_in && d_out_bitmap && d_out_wordprefix && out_shape3 && ksize3 && stride3 && pad3 && d_nbr,
             "fd_rulebook_neighbors_scatter: null argument");
  FD_REQUIRE(nbr_stride >= n_out_cap && n_out_cap >= 0, "fd_rulebook_neighbors_scatter: nbr_stride < n_out_cap");
  Conv3Geom g;
  for (int j = 0; j < 3; ++j) { g.k[j] = ksize3[j]; g.s[j] = stride3[j]; g.p[j] = pad3[j]; }
  const int K = g.k[0] * g.k[1] * g.k[2];
  FD_REQUIRE(K >= 1 && K <= 343, "fd_rulebook_neighbors_scatter: kernel volume %d unsupported", K);
  FD_REQUIRE(!d_tile_mask || K <= 32, "fd_rulebook_neighbors_scatter: tile masks support at most 32 kernel offsets");
  FD_REQUIRE(d_n_out != nullptr, "fd_rulebook_neighbors_scatter: null d_n_out");
  FD_REQUIRE(nbr_stride % 4 == 0 && (((uintptr_t)d_nbr) & 15) == 0, "fd_rulebook_neighbors_scatter: table rows must be 16-byte aligned");
  if (n_out_cap > 0) {
    nbr_fill_empty_kernel<<<persistent_grid(ceil_div((int64_t)K * ceil_div(n_out_cap, 4), 256), 8), 256, 0, stream>>>(
        d_nbr, nbr_stride, K, d_n_out, n_out_cap);
    FD_LAUNCHED();
  }
  if (d_tile_mask) FD_CUDA(cudaMemsetAsync(d_tile_mask, 0, sizeof(uint32_t) * (size_t)ceil_div(n_out_cap > 0 ? n_out_cap : 1, 128), stream));
  if (n_in_cap <= 0 || n_out_cap <= 0) return 0;
  Shape3 osh{out_shape3[0], out_shape3[1], out_shape3[2]};
  neighbors_scatter_kernel<<<persistent_grid(ceil_div(n_in_cap, 256), 8), 256, 0, stream>>>(
      (const int4*)d_in_coords4, d_n_in, n_in_cap, g, osh, d_out_bitmap, d_out_wordprefix, n_out_cap, d_nbr, nbr_stride,
      d_tile_mask);
  FD_LAUNCHED();
  return 0;
}

static int rowsort_window(int window) {
  if (window <= 0) window = 256 * 1024;
  return (window + fd::RS_ROWS - 1) / fd::RS_ROWS * fd::RS_ROWS;
}

size_t fd_rulebook_sort_workspace_bytes(int n_cap, int window) {
  using namespace fd;
  if (n_cap <= 0) return 256;
  window = rowsort_window(window);
  const size_t blocks = (size_t)ceil_div(n_cap, RS_ROWS), wins = (size_t)ceil_div(n_cap, window);
  const size_t rows = ((size_t)n_cap * 2 + 255) & ~(size_t)255;
  return 2 * rows + (blocks + wins) * RS_BUCKETS * sizeof(int32_t) + 256;
}

int fd_rulebook_sort_rows(const int32_t* d_nbr, int nbr_stride, int K, const int32_t* d_n_out, int n_out_cap, int window,
                          const uint16_t* d_row_key, int32_t* d_row_perm, int32_t* d_nbr_sorted, uint32_t* d_tile_mask_sorted, void* d_workspace,
                          size_t workspace_bytes, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_nbr && d_row_perm && d_nbr_sorted && d_workspace, "fd_rulebook_sort_rows: null argument");
  FD_REQUIRE(K >= 1 && K <= 32, "fd_rulebook_sort_rows: supports at most 32 kernel offsets (got %d)", K);
  FD_REQUIRE(nbr_stride >= n_out_cap && n_out_cap >= 0, "fd_rulebook_sort_rows: nbr_stride < n_out_cap");
  FD_REQUIRE(d_nbr != d_nbr_sorted, "fd_rulebook_sort_rows: the sorted table cannot alias the source");
  FD_REQUIRE(workspace_bytes >= fd_rulebook_sort_workspace_bytes(n_out_cap, window), "fd_rulebook_sort_rows: workspace too small");
  if (n_out_cap <= 0) return 0;
  window = rowsort_window(window);
  const int blocks = ceil_div(n_out_cap, RS_ROWS), wins = ceil_div(n_out_cap, window);
  const size_t rows = ((size_t)n_out_cap * 2 + 255) & ~(size_t)255;
  uint16_t* keys = (uint16_t*)d_workspace;
  uint16_t* lrank = (uint16_t*)((char*)d_workspace + rows);
  int32_t* blk_base = (int32_t*)((char*)d_workspace + 2 * rows);
  int32_t* win_hist = blk_base + (size_t)blocks * RS_BUCKETS;
  const int gs = ceil_div(K, 9);                 // kernel offsets per signature bit (3 for 3x3x3: one (dz, dy) line)
  FD_CUDA(cudaMemsetAsync(win_hist, 0, sizeof(int32_t) * (size_t)wins * RS_BUCKETS, stream));
  if (d_tile_mask_sorted) FD_CUDA(cudaMemsetAsync(d_tile_mask_sorted, 0, sizeof(uint32_t) * (size_t)ceil_div(n_out_cap, 128), stream));
  rowsort_count_kernel<<<blocks, 256, 0, stream>>>(d_nbr, nbr_stride, K, gs, d_n_out, n_out_cap, window, d_row_key, keys, lrank,
                                                   blk_base, win_hist);
  FD_LAUNCHED();
  rowsort_scan_kernel<<<wins, RS_BUCKETS, 0, stream>>>(win_hist, window);
  FD_LAUNCHED();
  static int kb = 0;
  if (!kb) {
    // columns staged per round: 3 (24 KB, 7 CTAs per SM).  Measured per 16-scene forward (six tables): 9 columns (2 CTAs per SM)
    // 1.50 ms, 5: 1.22, 3: 1.09, 2: 1.06, 1: 1.06 -- the kernel is occupancy (latency) bound, not bound by the number of rounds
    kb = getenv("FD_RS_KB") ? atoi(getenv("FD_RS_KB")) : 3;
    FD_CUDA(cudaFuncSetAttribute(rowsort_scatter_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * RS_ROWS * 4));
    FD_CUDA(cudaFuncSetAttribute(rowsort_scatter_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * RS_ROWS * 4));
    FD_CUDA(cudaFuncSetAttribute(rowsort_scatter_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * RS_ROWS * 4));
    FD_CUDA(cudaFuncSetAttribute(rowsort_scatter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RS_ROWS * 4));
    FD_CUDA(cudaFuncSetAttribute(rowsort_scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 * RS_ROWS * 4));
  }
#define FD_RS_LAUNCH(KB)                                                                                              \
  rowsort_scatter_kernel<KB><<<blocks, 256, KB * RS_ROWS * 4, stream>>>(d_nbr, nbr_stride, K, d_n_out, n_out_cap, window, \
                                                     d_row_key ? d_row_key : keys, lrank, blk_base,                   \
                                                     win_hist, d_row_perm, d_nbr_sorted, d_tile_mask_sorted)
  if (kb == 1) FD_RS_LAUNCH(1); else if (kb == 2) FD_RS_LAUNCH(2); else if (kb == 3) FD_RS_LAUNCH(3); else if (kb == 5) FD_RS_LAUNCH(5); else FD_RS_LAUNCH(9);
#undef FD_RS_LAUNCH
  FD_LAUNCHED();
  return 0;
}

int fd_rulebook_to_pairs(const int32_t* d_nbr, int nbr_stride, const int32_t* d_n_out, int n_out_cap, int K,
                         int32_t* d_pairs, int pair_cap, void* d_scan_tmp, void* stream_) {
  // d_scan_tmp: fd_scan_tmp_bytes(K*nbr_stride) + K*nbr_stride*4 bytes (positions)
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_nbr && d_pairs && d_scan_tmp && K >= 1, "fd_rulebook_to_pairs: null argument");
  int64_t total = (int64_t)K * nbr_stride;
  int32_t* pos = (int32_t*)d_scan_tmp;
  void* tmp = (char*)d_scan_tmp + ((sizeof(int32_t) * total + 255) & ~(size_t)255);
  FD_CUDA(cudaMemsetAsync(d_pairs, 0xff, sizeof(int32_t) * 2 * (size_t)K * pair_cap, stream));
  int rc = scan_impl(LoadValid{d_nbr, nbr_stride, d_n_out, n_out_cap}, pos, total, nullptr, tmp, stream);
  if (rc) return rc;
  for (int k = 0; k < K; ++k) {
    pairs_emit<<<persistent_grid(ceil_div(n_out_cap, 256), 4), 256, 0, stream>>>(
        d_nbr, nbr_stride, d_n_out, n_out_cap, pos, d_pairs, pair_cap, k);
    FD_LAUNCHED();
  }
  return 0;
}

}  // extern "C"
