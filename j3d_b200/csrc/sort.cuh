// sort.cuh — hand-written LSD radix sort of (uint64 key, uint32 value) pairs for the LBVH
// builder.  8-bit digits; per pass: tile histogram -> exclusive scan of the digit-major
// [256][tiles] table -> stable scatter.  The scatter ranks keys with warp match_any
// multisplit, stages the tile in shared memory in sorted order and writes each digit run
// contiguously, so global writes are coalesced per run.  No CUB / Thrust.
#pragma once
#include "common.cuh"

namespace rsort {

constexpr int THREADS = 256;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;  // 4096 keys per block
constexpr int RADIX = 256;
constexpr int WARPS = THREADS / 32;
constexpr size_t SCATTER_SMEM = (size_t)TILE * 8 + (size_t)TILE * 4 + (size_t)WARPS * RADIX * 4 + RADIX * 4 * 2;

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
    const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
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

// Scratch requirement in bytes for n keys (table + chunk sums), excluding the ping-pong buffers.
static inline size_t scratch_bytes(uint32_t n) {
  const size_t ntiles = ((size_t)n + TILE - 1) / TILE;
  const size_t table = ntiles * RADIX;
  const size_t nchunks = (table + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return (table + nchunks + 64) * sizeof(uint32_t);
}

// Sorts by bits [first_bit, key_bits) of the keys (keys that agree on those bits keep their input order).  Result ends in (keys_a, vals_a) if the number of
// passes is even, else in (keys_b, vals_b); returns which through *result_in_b.
// iota_values: vals_a need not be initialised, the first pass generates 0..n-1.
static inline int sort_pairs(j3dg_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n,
                             int key_bits, uint32_t* scratch, bool* result_in_b, bool iota_values = true, int first_bit = 0) {
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

}  // namespace rsort
