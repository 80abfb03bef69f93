// sort.cuh — hand-written LSD radix sort for the LBVH builder and the k-NN grid.  8-bit digits.  No CUB / Thrust.
//   one-sweep path (default): ONE histogram kernel counts every digit of every pass up front; each pass is then a
//     single kernel that reads and writes every key once: a tile (4096 keys) draws its index from a ticket counter,
//     ranks its keys with warp match_any multisplit, publishes its 256 digit counts in a status word (2 flag bits +
//     30 value bits), finds its global offsets by DECOUPLED LOOK-BACK over the tiles before it (they hold earlier
//     tickets, so they are running: no deadlock), stages the tile in shared memory in sorted order and writes every
//     digit run contiguously.  Keys are either (uint64 key, uint32 value) pairs or PACKED 64-bit words that carry the
//     value in their low bits (8 instead of 12 bytes per element and pass).
//   three-kernel-scan path (n >= 2^30, or J3DG_SORT=lsd): tile histogram -> exclusive scan of the digit-major
//     [256][tiles] table -> the same stable scatter.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace rsort {

constexpr int THREADS = 256;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;  // 4096 keys per block
constexpr int RADIX = 256;
constexpr int WARPS = THREADS / 32;
constexpr size_t SCATTER_SMEM = (size_t)TILE * 8 + (size_t)TILE * 4 + (size_t)WARPS * RADIX * 4 + RADIX * 4 * 2;


// Lanes of the warp whose 8-bit digit equals mine.  Eight ballots instead of one MATCH.ANY: the match instruction was the
// top stall of the scatter (30 % of the samples waited for its result; ncu, round 2).
static __device__ __forceinline__ uint32_t match_digit(uint32_t d) {
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool p = (d >> b) & 1u;
    const uint32_t m = __ballot_sync(0xffffffffu, p);
    peers &= p ? m : ~m;
  }
  return peers;
}

static __global__ void __launch_bounds__(THREADS) histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                             uint32_t* __restrict__ table, uint32_t ntiles) {
  __shared__ uint32_t hist[RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * TILE;
#pragma unroll 4
  for (int j = 0; j < ITEMS; ++j) {
    uint32_t i = base + j * THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  table[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// ---- exclusive scan of a uint32 array (3 small kernels) ------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;  // 2048

static __device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    uint32_t si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;  // exclusive
    if (lane == SCAN_THREADS / 32 - 1) *total = si;
  }
  __syncthreads();
  uint32_t r = incl - v + warp_sums[w];
  __syncthreads();
  return r;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_chunk_sums(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t total;
  size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (base + j < n) s += in[base + j];
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_sums_serial(uint32_t* __restrict__ sums, uint32_t nchunks) {
  // single block: scan the chunk totals in strips of SCAN_THREADS
  __shared__ uint32_t total;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nchunks; base += SCAN_THREADS) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nchunks ? sums[i] : 0;
    uint32_t e = block_exclusive_scan(v, &total);
    if (i < nchunks) sums[i] = e + carry;
    carry += total;
    __syncthreads();
  }
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_apply(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t total;
  size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? data[base + j] : 0;
    s += v[j];
  }
  uint32_t e = block_exclusive_scan(s, &total) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) data[base + j] = e;
    e += v[j];
  }
}

// In-place exclusive prefix sum of n uint32 (sums: scratch of ceil(n / SCAN_CHUNK) uint32).
static inline uint32_t scan_scratch_count(size_t n) { return (uint32_t)((n + SCAN_CHUNK - 1) / SCAN_CHUNK); }
static inline int exclusive_scan_u32(j3dg_ctx* ctx, uint32_t* data, size_t n, uint32_t* sums) {
  if (!n) return J3DG_OK;
  const uint32_t nchunks = scan_scratch_count(n);
  scan_chunk_sums<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(data, n, sums);
  KERNEL_CHECK(ctx);
  scan_sums_serial<<<1, SCAN_THREADS, 0, ctx->stream>>>(sums, nchunks);
  KERNEL_CHECK(ctx);
  scan_apply<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(data, n, sums);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}

// ---- stable scatter ------------------------------------------------------------------
#ifndef J3DG_SCATTER_MIN_BLOCKS
#define J3DG_SCATTER_MIN_BLOCKS 3   // 80 registers: three 52-KB blocks per SM (8.2 -> 7.7 ms build on 28 M triangles)
#endif
static __global__ void __launch_bounds__(THREADS, J3DG_SCATTER_MIN_BLOCKS) scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                           uint32_t n, int shift, const uint32_t* __restrict__ table, uint32_t ntiles,
                                                           int iota_vals) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* skeys = (uint64_t*)smem_raw;
  uint32_t* svals = (uint32_t*)(smem_raw + (size_t)TILE * 8);
  uint32_t* warp_hist = svals + TILE;          // [WARPS][RADIX]
  uint32_t* digit_start = warp_hist + WARPS * RADIX;  // [RADIX]
  uint32_t* gofs = digit_start + RADIX;               // [RADIX]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < WARPS * RADIX; i += THREADS) warp_hist[i] = 0;
  __syncthreads();

  const uint32_t tile_base = blockIdx.x * TILE;
  const uint32_t valid = min((uint32_t)TILE, n - tile_base);
  uint64_t key[ITEMS];
  uint32_t val[ITEMS];
  uint32_t rank[ITEMS];
  uint32_t* wh = warp_hist + warp * RADIX;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t local = warp * (32 * ITEMS) + j * 32 + lane;
    const uint32_t i = tile_base + local;
    const bool ok = local < valid;
    key[j] = ok ? keys_in[i] : ~0ull;
    val[j] = ok ? (iota_vals ? i : vals_in[i]) : 0u;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {  // ranked after ALL loads are in flight (see onesweep_kernel)
    const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
    const uint32_t peers = match_digit(d);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) {
      base = wh[d];
      wh[d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = base + __popc(peers & lanemask_lt);
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive offsets over warps, block total
  {
    const int d = threadIdx.x;
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      uint32_t c = warp_hist[w * RADIX + d];
      warp_hist[w * RADIX + d] = sum;
      sum += c;
    }
    // exclusive scan of the 256 digit totals
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __shared__ uint32_t wsum[WARPS];
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
      if (w < warp) woff += wsum[w];
    const uint32_t start = incl - sum + woff;
    digit_start[d] = start;
    gofs[d] = table[(size_t)d * ntiles + blockIdx.x] - start;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
    const uint32_t pos = digit_start[d] + wh[d] + rank[j];
    skeys[pos] = key[j];
    svals[pos] = val[j];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < valid; i += THREADS) {
    const uint64_t k = skeys[i];
    const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
    const uint32_t out = gofs[d] + i;
    keys_out[out] = k;
    vals_out[out] = svals[i];
  }
}

// ---- one-sweep: all digit histograms up front, then one kernel per pass (decoupled look-back) ------------------------
constexpr int MAX_PASSES = 8;
constexpr uint32_t ST_AGGREGATE = 1u << 30, ST_INCLUSIVE = 2u << 30, ST_VALUE = (1u << 30) - 1u;

// ghist[p][d] += number of keys whose digit p (bits [first_shift + 8 p, +8)) is d.  Persistent blocks, shared-memory
// counters, one flush of passes * 256 global atomics per block.
static __global__ void __launch_bounds__(512) digit_histograms_kernel(const uint64_t* __restrict__ keys, uint32_t n, int first_shift, int passes,
                                                                      uint32_t* __restrict__ ghist) {
  __shared__ uint32_t h[MAX_PASSES * RADIX];
  for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {  // four loads in flight per thread
    uint64_t k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) k[u] = keys[i + u * stride] >> first_shift;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      for (int p = 0; p < passes; ++p) atomicAdd(&h[p * RADIX + ((uint32_t)(k[u] >> (8 * p)) & 0xffu)], 1u);
  }
  for (; i < n; i += stride) {
    const uint64_t k = keys[i] >> first_shift;
    for (int p = 0; p < passes; ++p) atomicAdd(&h[p * RADIX + ((uint32_t)(k >> (8 * p)) & 0xffu)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x)
    if (h[i]) atomicAdd(&ghist[i], h[i]);
}

#ifndef J3DG_ONESWEEP_BLOCKS_PAIRS
#define J3DG_ONESWEEP_BLOCKS_PAIRS 3
#endif
#ifndef J3DG_ONESWEEP_BLOCKS_KEYS
#define J3DG_ONESWEEP_BLOCKS_KEYS 4   // 64 registers (a few spilled words): 5.68 -> 5.61 ms on config B
#endif
template <bool PAIRS> constexpr size_t onesweep_smem() { return (size_t)TILE * 8 + (PAIRS ? (size_t)TILE * 4 : 0) + (size_t)WARPS * RADIX * 4 + RADIX * 4 * 2; }

// first_mode: 0 plain; 1 (PAIRS) the values are generated as 0..n-1; 2 (keys only) the input are raw codes and the key is
// packed on load: ((code >> pack_rshift) << pack_lshift) | position.  `shift` always addresses the key that is written.
// status: [ntiles][256] words, zero before the launch, followed by the ticket counter (zero as well).
template <bool PAIRS>
static __global__ void __launch_bounds__(THREADS, PAIRS ? J3DG_ONESWEEP_BLOCKS_PAIRS : J3DG_ONESWEEP_BLOCKS_KEYS)
onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                uint32_t n, int shift, const uint32_t* __restrict__ ghist, uint32_t* status, uint32_t ntiles, int first_mode, int pack_rshift, int pack_lshift) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* skeys = (uint64_t*)smem_raw;
  uint32_t* svals = (uint32_t*)(smem_raw + (size_t)TILE * 8);
  uint32_t* warp_hist = svals + (PAIRS ? TILE : 0);   // [WARPS][RADIX]
  uint32_t* digit_start = warp_hist + WARPS * RADIX;  // [RADIX]
  uint32_t* gofs = digit_start + RADIX;               // [RADIX]
  __shared__ uint32_t s_tile;
  __shared__ uint32_t wsum[2][WARPS];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  if (threadIdx.x == 0) s_tile = atomicAdd(status + (size_t)ntiles * RADIX, 1u);  // tiles are handed out in launch order: a tile only ever waits for tiles that are running
  for (int i = threadIdx.x; i < WARPS * RADIX; i += THREADS) warp_hist[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_base = tile * TILE;
  const uint32_t valid = min((uint32_t)TILE, n - tile_base);
  uint64_t key[ITEMS];
  uint32_t val[PAIRS ? ITEMS : 1];
  uint32_t rank[ITEMS];
  uint32_t* wh = warp_hist + warp * RADIX;
  // all loads of the tile first: __syncwarp() in the ranking loop is a memory barrier, the compiler does not move a load across
  // it, and sixteen DRAM latencies in a row per tile were what bounded the kernel (255 us per pass whatever the bytes)
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t local = warp * (32 * ITEMS) + j * 32 + lane;
    const uint32_t i = tile_base + local;
    const bool ok = local < valid;
    uint64_t k = ok ? keys_in[i] : ~0ull;
    if (!PAIRS && first_mode == 2 && ok) k = ((k >> pack_rshift) << pack_lshift) | (uint64_t)i;
    key[j] = k;
    if (PAIRS) val[j] = ok ? (first_mode == 1 ? i : vals_in[i]) : 0u;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
    const uint32_t peers = match_digit(d);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) {
      base = wh[d];
      wh[d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = base + __popc(peers & lanemask_lt);
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // THREADS == RADIX: thread d owns digit d
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const uint32_t c = warp_hist[w * RADIX + d];
      warp_hist[w * RADIX + d] = sum;
      sum += c;
    }
    volatile uint32_t* st = status;
    st[(size_t)tile * RADIX + d] = (tile == 0 ? ST_INCLUSIVE : ST_AGGREGATE) | sum;  // published before anything else: the followers wait for it
    // exclusive scans over the digits: of this tile's counts (position in the staged tile) and of the global counts
    const uint32_t gh = ghist[d];
    uint32_t incl = sum, gincl = gh;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t0 = __shfl_up_sync(0xffffffffu, incl, o), t1 = __shfl_up_sync(0xffffffffu, gincl, o);
      if (lane >= o) { incl += t0; gincl += t1; }
    }
    if (lane == 31) { wsum[0][warp] = incl; wsum[1][warp] = gincl; }
    __syncthreads();
    uint32_t woff = 0, gwoff = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
      if (w < warp) { woff += wsum[0][w]; gwoff += wsum[1][w]; }
    const uint32_t start = incl - sum + woff;
    const uint32_t gstart = gincl - gh + gwoff;
    // look back: keys of digit d in the tiles before this one
    // Four predecessors per round trip: the walk is as long as the number of tiles that are in their look-back at the same
    // time (it does not shrink on its own: each step costs an L2 latency, in which a dozen more tiles start), so what counts
    // is the number of DEPENDENT reads (27 % of the samples sat here with one predecessor per step).
    uint32_t before = 0;
    for (uint32_t t = tile; t > 0;) {  // tiles [0, t) are still to be accounted for
      const uint32_t v0 = st[(size_t)(t - 1) * RADIX + d];
      const uint32_t v1 = t >= 2 ? st[(size_t)(t - 2) * RADIX + d] : (uint32_t)(2u << 30);
      const uint32_t v2 = t >= 3 ? st[(size_t)(t - 3) * RADIX + d] : (uint32_t)(2u << 30);
      const uint32_t v3 = t >= 4 ? st[(size_t)(t - 4) * RADIX + d] : (uint32_t)(2u << 30);
      if ((v0 >> 30) == 0u) continue;  // not published yet
      before += v0 & ST_VALUE;
      if (v0 & ST_INCLUSIVE) break;
      --t;
      if ((v1 >> 30) == 0u) continue;
      before += v1 & ST_VALUE;
      if (v1 & ST_INCLUSIVE) break;
      --t;
      if ((v2 >> 30) == 0u) continue;
      before += v2 & ST_VALUE;
      if (v2 & ST_INCLUSIVE) break;
      --t;
      if ((v3 >> 30) == 0u) continue;
      before += v3 & ST_VALUE;
      if (v3 & ST_INCLUSIVE) break;
      --t;
    }
    if (tile > 0) st[(size_t)tile * RADIX + d] = ST_INCLUSIVE | (before + sum);
    digit_start[d] = start;
    gofs[d] = gstart + before - start;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
    const uint32_t pos = digit_start[d] + wh[d] + rank[j];
    skeys[pos] = key[j];
    if (PAIRS) svals[pos] = val[j];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < valid; i += THREADS) {
    const uint64_t k = skeys[i];
    const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
    const uint32_t out = gofs[d] + i;
    keys_out[out] = k;
    if (PAIRS) vals_out[out] = svals[i];
  }
}
static_assert(THREADS == RADIX, "onesweep_kernel: one thread per digit");

static inline bool use_onesweep(uint32_t n) {
  static const bool lsd = [] { const char* e = getenv("J3DG_SORT"); return e && !strcmp(e, "lsd"); }();
  return !lsd && n < (1u << 30);
}

// Scratch requirement in bytes for n keys, excluding the ping-pong buffers: the status words of one pass (or the digit-major
// table of the scan path, same size) + the chunk sums of the scan path + ticket + the digit histograms of all passes.
static inline size_t scratch_bytes(uint32_t n) {
  const size_t ntiles = ((size_t)n + TILE - 1) / TILE;
  const size_t table = ntiles * RADIX;
  const size_t nchunks = (table + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return (table + nchunks + 64 + (size_t)MAX_PASSES * RADIX) * sizeof(uint32_t);
}

// The one-sweep driver.  PAIRS: (keys, vals) ping-pong, values generated if iota_values.  Keys only: `pack` turns the raw codes
// into packed keys ((code >> first_bit) << idx_bits) | position in the first pass; the sorted bits are then [idx_bits, idx_bits + 8 passes).
template <bool PAIRS>
static inline int onesweep_sort(j3dg_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, int first_bit, int passes,
                                uint32_t* scratch, bool* result_in_b, bool iota_values, bool pack, int idx_bits) {
  CU_CHECK(ctx, cudaFuncSetAttribute(onesweep_kernel<PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<PAIRS>()));  // per device
  const uint32_t ntiles = (n + TILE - 1) / TILE;
  const size_t status_words = (size_t)ntiles * RADIX + 1;  // + the ticket
  uint32_t* status = scratch;
  uint32_t* ghist = scratch + ((status_words + 63) & ~(size_t)63);
  CU_CHECK(ctx, cudaMemsetAsync(ghist, 0, (size_t)passes * RADIX * sizeof(uint32_t), ctx->stream));
  const int hist_blocks = (int)std::min<size_t>(((size_t)n + 511) / 512, (size_t)ctx->sm_count * 4);
  digit_histograms_kernel<<<hist_blocks, 512, 0, ctx->stream>>>(keys_a, n, first_bit, passes, ghist);
  KERNEL_CHECK(ctx);
  bool in_b = false;
  for (int pass = 0; pass < passes; ++pass) {
    const uint64_t* kin = in_b ? keys_b : keys_a;
    const uint32_t* vin = in_b ? vals_b : vals_a;
    uint64_t* kout = in_b ? keys_a : keys_b;
    uint32_t* vout = in_b ? vals_a : vals_b;
    CU_CHECK(ctx, cudaMemsetAsync(status, 0, status_words * sizeof(uint32_t), ctx->stream));
    const int shift = (pack ? idx_bits : first_bit) + 8 * pass;
    const int mode = pass == 0 ? (pack ? 2 : (iota_values ? 1 : 0)) : 0;
    onesweep_kernel<PAIRS><<<ntiles, THREADS, onesweep_smem<PAIRS>(), ctx->stream>>>(kin, vin, kout, vout, n, shift, ghist + (size_t)pass * RADIX, status, ntiles, mode,
                                                                                      first_bit, idx_bits);
    KERNEL_CHECK(ctx);
    in_b = !in_b;
  }
  *result_in_b = in_b;
  return J3DG_OK;
}

// Sorts by bits [first_bit, key_bits) of the keys (keys that agree on those bits keep their input order).  Result ends in (keys_a, vals_a) if the number of
// passes is even, else in (keys_b, vals_b); returns which through *result_in_b.
// iota_values: vals_a need not be initialised, the first pass generates 0..n-1.
static inline int sort_pairs(j3dg_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n,
                             int key_bits, uint32_t* scratch, bool* result_in_b, bool iota_values = true, int first_bit = 0) {
  const int npasses = (key_bits - first_bit + 7) / 8;
  if (use_onesweep(n) && npasses >= 1 && npasses <= MAX_PASSES)
    return onesweep_sort<true>(ctx, keys_a, vals_a, keys_b, vals_b, n, first_bit, npasses, scratch, result_in_b, iota_values, false, 0);
  // per device, not per process: a second context on another GPU needs it as well (a few microseconds per sort)
  CU_CHECK(ctx, cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM));
  const uint32_t ntiles = (n + TILE - 1) / TILE;
  const size_t table_n = (size_t)ntiles * RADIX;
  const uint32_t nchunks = (uint32_t)((table_n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  uint32_t* table = scratch;
  uint32_t* sums = scratch + table_n;
  bool in_b = false;
  int pass = 0;
  for (int shift = first_bit; shift < key_bits; shift += 8, ++pass) {
    const uint64_t* kin = in_b ? keys_b : keys_a;
    const uint32_t* vin = in_b ? vals_b : vals_a;
    uint64_t* kout = in_b ? keys_a : keys_b;
    uint32_t* vout = in_b ? vals_a : vals_b;
    histogram_kernel<<<ntiles, THREADS, 0, ctx->stream>>>(kin, n, shift, table, ntiles);
    KERNEL_CHECK(ctx);
    scan_chunk_sums<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(table, table_n, sums);
    KERNEL_CHECK(ctx);
    scan_sums_serial<<<1, SCAN_THREADS, 0, ctx->stream>>>(sums, nchunks);
    KERNEL_CHECK(ctx);
    scan_apply<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(table, table_n, sums);
    KERNEL_CHECK(ctx);
    scatter_kernel<<<ntiles, THREADS, SCATTER_SMEM, ctx->stream>>>(kin, vin, kout, vout, n, shift, table, ntiles, (pass == 0 && iota_values) ? 1 : 0);
    KERNEL_CHECK(ctx);
    in_b = !in_b;
  }
  *result_in_b = in_b;
  return J3DG_OK;
}

// Packed sort for the builder: `codes` (keys_a) hold raw 48-bit codes; on return the result buffer holds
// ((code >> first_bit) << idx_bits) | original position, ordered by the code bits [first_bit, first_bit + 8 passes), equal codes in input order.
// Only valid when the sorted bits (48 - first_bit <= 8 passes) and idx_bits fit 64 bits together and the one-sweep path applies (can_sort_packed).
static inline bool can_sort_packed(uint32_t n, int passes, int idx_bits, int sorted_bits) {
  static const bool off = [] { const char* e = getenv("J3DG_SORT"); return e && !strcmp(e, "pairs"); }();
  return !off && use_onesweep(n) && passes >= 1 && passes <= MAX_PASSES && sorted_bits <= 8 * passes && sorted_bits + idx_bits <= 64;
}
static inline int sort_packed(j3dg_ctx* ctx, uint64_t* keys_a, uint64_t* keys_b, uint32_t n, int first_bit, int passes, int idx_bits, uint32_t* scratch,
                              bool* result_in_b) {
  return onesweep_sort<false>(ctx, keys_a, nullptr, keys_b, nullptr, n, first_bit, passes, scratch, result_in_b, false, true, idx_bits);
}

}  // namespace rsort
