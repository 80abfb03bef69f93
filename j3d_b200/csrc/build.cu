// build.cu — GPU BVH build.  Replaces `new qbvh(triangles, vertices)` + compute_triangle_normals
// + compute_bb of add_object (j3d/scene.cpp:8-25; jtk/qbvh.h:1679-1686, 2133-2459, 5299-5339).
//
// Not a port of the reference's top-down binned-SAH QBVH: the closest hit does not depend on
// the tree, so the tree is built the way a GPU builds one:
//   1. vertex bbox (== compute_bb)                               bbox_kernel
//   2. 48-bit Morton code of each triangle's box centre           morton_kernel
//   3. LSD radix sort of (code, triangle)                         sort.cuh
//   4. binary radix tree over the sorted codes (Karras 2012),
//      pre-gathered 48-byte triangle records, box fit             tree_fit_kernel (+ climb_kernel for the upper tree)
//   6. top-down collapse into 8-wide quantised 128-byte nodes,
//      opening the largest-area child first (SAH-greedy)          collapse_kernel (one launch per level)
// Every subtree of the radix tree owns a contiguous range of the sorted triangle records, so a
// leaf reference is just (first, count).
#include "common.cuh"
#include "sort.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int MORTON_BITS_PER_AXIS = 16;
constexpr int MORTON_BITS = 3 * MORTON_BITS_PER_AXIS;

// ---- 1. bbox ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __host__ __forceinline__ float ordered_to_float(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
#ifdef __CUDA_ARCH__
  f = __uint_as_float(b);
#else
  memcpy(&f, &b, 4);
#endif
  return f;
}

__global__ void bbox_init_kernel(uint32_t* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = 0xFFFFFFFFu;
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

// The vertex array is read as a stream of float4 (16-byte loads, a third of the load instructions; measured: 69 us on config B
// either way, the kernel waits on DRAM latency, not on issue): the component of the
// first float of vector i is i mod 3, and with a grid stride that is a multiple of 3 it stays the same for a thread, so the
// four lanes of a vector go to fixed accumulators.  Unaligned arrays and the last (3 nv mod 4) floats take the scalar path.
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ v, uint32_t nv, uint32_t* __restrict__ bb) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  const size_t total = 3 * (size_t)nv;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;  // the host makes the stride a multiple of 3
  const bool aligned = (reinterpret_cast<uintptr_t>(v) & 15u) == 0u;
  const size_t nvec = aligned ? total / 4 : 0;
  {
    float a0n = FLT_MAX, a1n = FLT_MAX, a2n = FLT_MAX, a0x = -FLT_MAX, a1x = -FLT_MAX, a2x = -FLT_MAX;
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (size_t i = gtid; i < nvec; i += stride) {
      const float4 f = v4[i];
      a0n = fminf(a0n, fminf(f.x, f.w)); a0x = fmaxf(a0x, fmaxf(f.x, f.w));
      a1n = fminf(a1n, f.y); a1x = fmaxf(a1x, f.y);
      a2n = fminf(a2n, f.z); a2x = fmaxf(a2x, f.z);
    }
    const int r = (int)(gtid % 3);  // component of f.x (and f.w); f.y is r + 1, f.z is r + 2
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int which = (j - r + 3) % 3;
      mn[j] = which == 0 ? a0n : which == 1 ? a1n : a2n;
      mx[j] = which == 0 ? a0x : which == 1 ? a1x : a2x;
    }
  }
  for (size_t i = 4 * nvec + gtid; i < total; i += stride) {  // the tail, or everything if the array is not 16-byte aligned
    const float c = v[i];
    const int j = (int)(i % 3);
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q == j) { mn[q] = fminf(mn[q], c); mx[q] = fmaxf(mx[q], c); }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j)
    for (int o = 16; o; o >>= 1) {
      mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], o));
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      atomicMin(&bb[j], float_to_ordered(mn[j]));
      atomicMax(&bb[3 + j], float_to_ordered(mx[j]));
    }
  }
}

// ---- 2. Morton codes ------------------------------------------------------------------------
// 8 bits -> every third bit of 24 (32-bit operations; the 48-bit code is assembled from a low and a high half)
__device__ __forceinline__ uint32_t spread8(uint32_t x) {
  x &= 0xffu;
  x = (x | (x << 8)) & 0x00F00Fu;
  x = (x | (x << 4)) & 0x0C30C3u;
  x = (x | (x << 2)) & 0x249249u;
  return x;
}

// The kernel is bound by instruction issue (ncu: 74 % of the issue slots, 245 instructions per triangle with 64-bit
// bit spreading, three IEEE divisions and a log2 per thread), so: the scale of the quantisation once per block, the code
// from 32-bit halves, the size statistic only in the blocks that sample it.  Same keys bit for bit.
__global__ void __launch_bounds__(256) morton_kernel(const float* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t nt,
                                                      const uint32_t* __restrict__ bb, uint64_t* __restrict__ keys, unsigned long long* __restrict__ size_acc) {
  __shared__ float s_mn[3], s_inv[3], s_ext[3];
  if (threadIdx.x < 3) {
    const float lo = ordered_to_float(bb[threadIdx.x]);
    const float ext = ordered_to_float(bb[3 + threadIdx.x]) - lo;
    s_mn[threadIdx.x] = lo;
    s_ext[threadIdx.x] = ext;
    s_inv[threadIdx.x] = ext > 0.f ? 65535.99f / ext : 0.f;
  }
  __syncthreads();
  const uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t0 < nt;
  const uint32_t t = valid ? t0 : nt - 1u;  // the tail lanes recompute the last triangle and contribute nothing
  const uint32_t i0 = idx[3 * (size_t)t], i1 = idx[3 * (size_t)t + 1], i2 = idx[3 * (size_t)t + 2];
  uint32_t q[3];
  float tri_ext = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float a = verts[3 * (size_t)i0 + j], b = verts[3 * (size_t)i1 + j], c = verts[3 * (size_t)i2 + j];
    const float lo = fminf(a, fminf(b, c)), hi = fmaxf(a, fmaxf(b, c));
    const float ctr = 0.5f * (lo + hi);
    float f = (ctr - s_mn[j]) * s_inv[j];
    f = fminf(fmaxf(f, 0.f), 65535.f);
    q[j] = (uint32_t)f;
    tri_ext = fmaxf(tri_ext, hi - lo);
  }
  if (valid) {
    const uint32_t lo24 = (spread8(q[0]) << 2) | (spread8(q[1]) << 1) | spread8(q[2]);
    const uint32_t hi24 = (spread8(q[0] >> 8) << 2) | (spread8(q[1] >> 8) << 1) | spread8(q[2] >> 8);
    keys[t] = ((uint64_t)hi24 << 24) | (uint64_t)lo24;
  }
  // how many times the scene box is larger than this triangle, in bits (sum over the mesh: the builder derives the
  // Morton resolution that is worth sorting from the mean, see j3dg_build_bvh)
  // a sample is enough (every 16th block), and it keeps the atomics on the two words rare; integer sums: same total in any order
  if ((blockIdx.x & 15u) == 0u) {
    const float box_ext = fmaxf(s_ext[0], fmaxf(s_ext[1], s_ext[2]));
    const float bits = tri_ext > 0.f ? fminf(fmaxf(log2f(box_ext / tri_ext), 0.f), 24.f) : 24.f;
    const uint32_t fixed = __reduce_add_sync(0xffffffffu, valid ? (uint32_t)(bits * 256.f) : 0u);
    const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, valid));
    if ((threadIdx.x & 31) == 0) { atomicAdd(size_acc, (unsigned long long)fixed); atomicAdd(size_acc + 1, (unsigned long long)cnt); }
  }
}

// ---- 4. binary radix tree (Karras 2012) ----------------------------------------------------
// Node numbering: inner nodes 0..n-2 (root 0), leaf k = n-1+k.
struct BinTree {
  // topo / parent / flags are written for the UPPER tree only (nodes whose range straddles a 256-leaf block, and parent[] of their children):
  // climb_kernel is their one reader.  A node fitted inside a block hands its children to the collapse in its box.
  uint4* topo;       // [n-1] {left child, right child, first sorted leaf, last sorted leaf}: one 16-byte load per node
  uint32_t* parent;  // [2n-1]
  uint32_t* flags;   // [n-1] arrival counters of the atomic walk
  float4* box;       // [2 (n-1)] inner nodes only: box[2 i] = {min, left child}, box[2 i + 1] = {max, right child} (children as bit patterns in .w): one 32-byte sector tells the collapse
                     // everything about a node; a leaf's box is recomputed from its 48-byte record, and ranges follow from the parent's range and the split
};


// `mask` selects the key bits that were sorted; keys that agree on them are told apart by their position.
// ki = keys[i] & mask, kept in a register by the caller (upper_tree_kernel; tree_fit_kernel works on the table of adjacent deltas).
__device__ __forceinline__ int delta(uint64_t ki, const uint64_t* __restrict__ keys, uint64_t mask, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint64_t b = __ldg(keys + j) & mask;
  if (ki != b) return __clzll((long long)(ki ^ b));
  return 64 + __clz(i ^ j);
}

// ---- 4 + 5. radix tree, leaf records, box fit: ONE kernel -----------------------------------
// One block per 256 consecutive sorted triangles; thread k owns leaf k and inner node k (Karras numbering: node k's range
// has k at one end).
//   * The gather chain of the leaf (key -> three indices -> three vertices: three dependent DRAM round trips) and the
//     construction of the node are INTERLEAVED in program order, so the node is built while the gathers are in flight.  As
//     separate kernels (radix tree with Karras' searches on the keys, then refit) they took 0.78 ms (issue slots 77 % busy,
//     DRAM idle) + 1.2 ms (issue slots 28 % busy).
//   * The node comes from a shared-memory table of the deltas of ADJACENT keys (the block's 256 pairs and 256 on either
//     side): range and split by binary lifting, see below.  tests/test_builder_model.py pins the equivalence with Karras'
//     searches in plain Python.
//   * A subtree owns a contiguous leaf range, so the box of every node whose range lies inside the block (all but ~1 %)
//     is a RANGE minimum / maximum over the block's leaf boxes: a sparse table is grown in shared memory level by level
//     (T_j[i] = boxes [i, i + 2^j), eight rounds of fully active threads, one barrier each, ping-pong buffers) and a node
//     of length [2^j, 2^(j+1)) takes its box from two entries of level j.  min / max are exact: bit-identical to a
//     bottom-up fit.
//   * Topology, parent and arrival flag go to global memory only for the nodes that STRADDLE the block (the upper tree);
//     a node fitted here hands its children to the collapse in the .w words of its box and nobody else asks.
//   * The block's top nodes (parent straddles the block) are LISTED; climb_kernel walks the upper tree from them with the
//     classic atomic rule (the second child to arrive fits the node and moves on).  At most 2 per level of the tree can
//     exist per block (one per boundary: a straddling node has at most one child that lies inside), CLIMB_SLOTS covers
//     80 levels; an overflow is reported, never dropped.
constexpr int REFIT_THREADS = 256;
constexpr int CLIMB_SLOTS = 160;
constexpr int DT_PAIRS = 3 * REFIT_THREADS;  // window of the delta table: the block's pairs and 256 on either side

// Box of a binary node: inner nodes from the box array, a leaf from its 48-byte record.  L2 loads: the data may have been
// written by another block of the running kernel.
__device__ __forceinline__ void node_box(const BinTree& t, const TriRec* recs, uint32_t first_leaf, uint32_t node, float4& mn, float4& mx) {
  if (node >= first_leaf) {
    const float4* rp = reinterpret_cast<const float4*>(recs + (node - first_leaf));
    const float4 a = __ldcg(rp), b = __ldcg(rp + 1), c = __ldcg(rp + 2);
    mn = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
    mx = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
  } else {
    mn = __ldcg(&t.box[2 * (size_t)(node)]);
    mx = __ldcg(&t.box[2 * (size_t)(node) + 1]);
  }
}

// The atomic walk: the second child to arrive at a node fits it and moves on.  The node's topology and parent are fetched
// while the arrival atomic is in flight.
__device__ __forceinline__ void climb(const BinTree& t, const TriRec* recs, uint32_t first_leaf, uint32_t cur_node, float4 mn, float4 mx) {
  uint32_t p = t.parent[cur_node];
  while (p != 0xFFFFFFFFu) {
    const uint4 tp = t.topo[p];
    const uint32_t pp = t.parent[p];
    if (atomicAdd(&t.flags[p], 1u) == 0u) break;  // first arrival: the sibling subtree finishes this node
    const uint32_t sibling = (tp.x == cur_node) ? tp.y : tp.x;
    float4 omn, omx;
    node_box(t, recs, first_leaf, sibling, omn, omx);
    mn = make_float4(fminf(mn.x, omn.x), fminf(mn.y, omn.y), fminf(mn.z, omn.z), __uint_as_float(tp.x));
    mx = make_float4(fmaxf(mx.x, omx.x), fmaxf(mx.y, omx.y), fmaxf(mx.z, omx.z), __uint_as_float(tp.y));
    t.box[2 * (size_t)(p)] = mn;
    t.box[2 * (size_t)(p) + 1] = mx;
    __threadfence();
    cur_node = p;
    p = pp;
  }
}

// keys: the sorted keys (packed: the triangle rides in the low bits, idx_mask selects it; pairs: sorted_tri holds it).
// key_mask selects the bits that were sorted.  status[0] is set if a block has more top nodes than CLIMB_SLOTS, status[1] counts the listed nodes, status[2] the deferred ones.
#ifndef J3DG_TREEFIT_MIN_BLOCKS
#define J3DG_TREEFIT_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(REFIT_THREADS, J3DG_TREEFIT_MIN_BLOCKS) tree_fit_kernel(const float* __restrict__ verts, const uint32_t* __restrict__ idx,
                                                                  const uint32_t* __restrict__ sorted_tri, const uint64_t* __restrict__ keys, uint64_t key_mask,
                                                                  uint32_t idx_mask, int n, BinTree t, TriRec* __restrict__ recs, uint32_t* __restrict__ climbers,
                                                                  uint32_t climb_cap, uint32_t* __restrict__ deferred_nodes, uint32_t* __restrict__ status) {
  __shared__ float s_tab[2][6][REFIT_THREADS];  // sparse table, ping-pong: [.][0..2] min, [.][3..5] max
  __shared__ uint8_t s_lpar[REFIT_THREADS], s_ipar[REFIT_THREADS];  // leaf k / inner node k has its parent fitted in this block
  __shared__ uint32_t s_nclimb, s_base, s_ndefer, s_dbase;
  __shared__ uint8_t s_dt[9][DT_PAIRS];  // delta table: [v][a] = min over the adjacent pairs [a, a + 2^v) of 1 + delta
  const int tid = threadIdx.x;
  const int s = blockIdx.x * REFIT_THREADS, e = min(s + REFIT_THREADS, n);
  const int k = s + tid;
  const uint32_t first_leaf = (uint32_t)(n - 1);
  const bool leaf_ok = k < e, node_ok = k < e && k < n - 1;
  // ---- stage 1 of the gather: the key (and with it the triangle), and the neighbour keys of the delta table ----
  // Window of adjacent pairs: local pair a <-> keys (q, q + 1) with q = s - 256 + a, a in [0, 768): the block's own 256 pairs
  // in the middle, 256 on either side.  Thread tid owns pairs tid, tid + 256 (= (k, k + 1)), tid + 512.
  const uint64_t key_raw = leaf_ok ? __ldg(keys + k) : 0ull;
  uint64_t kq[3][2];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    const long long q = (long long)s - REFIT_THREADS + tid + (long long)w * REFIT_THREADS;
    const bool ok = q >= 0 && q + 1 < (long long)n;
    kq[w][0] = ok ? (w == 1 ? key_raw : __ldg(keys + q)) & key_mask : 0ull;
    kq[w][1] = ok ? __ldg(keys + q + 1) & key_mask : 0ull;
  }
  const uint32_t tri = leaf_ok ? (sorted_tri ? sorted_tri[k] : ((uint32_t)key_raw & idx_mask)) : 0u;
  // ---- stage 2 issued: the three vertex indices ----
  uint32_t i0 = 0, i1 = 0, i2 = 0;
  if (leaf_ok) { i0 = idx[3 * (size_t)tri]; i1 = idx[3 * (size_t)tri + 1]; i2 = idx[3 * (size_t)tri + 2]; }
  // ---- delta table, level 0: 1 + delta of each adjacent pair (0 past the ends of the array) ----
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    const long long q = (long long)s - REFIT_THREADS + tid + (long long)w * REFIT_THREADS;
    const bool ok = q >= 0 && q + 1 < (long long)n;
    const uint64_t x = kq[w][0] ^ kq[w][1];
    const int dl = x ? __clzll((long long)x) : 64 + __clz((int)((uint32_t)q ^ (uint32_t)(q + 1)));
    s_dt[0][w * REFIT_THREADS + tid] = ok ? (uint8_t)(dl + 1) : (uint8_t)0;
  }
  // ---- stage 3 issued: the vertices ----
  float3 a = make_float3(0.f, 0.f, 0.f), b = a, c = a;
  if (leaf_ok) {
    a = make_float3(verts[3 * (size_t)i0], verts[3 * (size_t)i0 + 1], verts[3 * (size_t)i0 + 2]);
    b = make_float3(verts[3 * (size_t)i1], verts[3 * (size_t)i1 + 1], verts[3 * (size_t)i1 + 2]);
    c = make_float3(verts[3 * (size_t)i2], verts[3 * (size_t)i2 + 1], verts[3 * (size_t)i2 + 2]);
  }
  // ---- delta table, levels 1..8: T[v][a] = min of the pairs [a, a + 2^v) (clipped at the window's end) ----
#pragma unroll 1
  for (int v = 1; v <= 8; ++v) {
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const int p = w * REFIT_THREADS + tid;
      s_dt[v][p] = min(s_dt[v - 1][p], s_dt[v - 1][min(p + (1 << (v - 1)), DT_PAIRS - 1)]);
    }
  }
  __syncthreads();
  // ---- node k from the table.  The common prefix of two sorted keys is the minimum over the adjacent pairs between them,
  // so Karras' searches (how far does the range reach, where does it split) become binary lifting over T: nine uniform
  // steps of one byte load each instead of ~20 data-dependent evaluations of delta on 64-bit keys with every lane of the
  // warp waiting for the longest search (38 % of the fused kernel's instructions at 14 active threads).  Identical tree.
  // A range of 256 pairs or more is DEFERRED to upper_tree_kernel (it cannot be fitted here anyway, and its probes of far-away
  // keys would keep the block waiting).
  int left = 0, right = 0, lo = 0, hi = 0;
  bool mine = false, deferred = false;
  if (node_ok && k == 0) t.parent[0] = 0xFFFFFFFFu;
  if (node_ok) {
    const int pf = REFIT_THREADS + tid, pb = REFIT_THREADS + tid - 1;  // pairs (k, k + 1) and (k - 1, k)
    const uint32_t fwd = s_dt[0][pf], bwd = s_dt[0][pb];
    const int d = fwd > bwd ? 1 : -1;  // never equal: the two pairs differ in the index bits at the latest
    const uint32_t dmin = min(fwd, bwd);
    int l = 0;  // pairs in the range
    if (d > 0) {
      int pos = pf;
#pragma unroll
      for (int v = 8; v >= 0; --v)
        if (pos + (1 << v) <= DT_PAIRS && s_dt[v][pos] > dmin) pos += 1 << v;
      l = pos - pf;
    } else {
      int pos = pb;
#pragma unroll
      for (int v = 8; v >= 0; --v)
        if (pos - (1 << v) + 1 >= 0 && s_dt[v][pos - (1 << v) + 1] > dmin) pos -= 1 << v;
      l = pb - pos;
    }
    deferred = l >= REFIT_THREADS;  // the window shows at least 256 pairs either way, so l < 256 is exact
    if (!deferred) {
      const int p0 = d > 0 ? pf : pb - l + 1, p1 = p0 + l - 1;  // first and last pair of the range
      const int lv = 31 - __clz(l);
      const uint32_t dnode = min(s_dt[lv][p0], s_dt[lv][p1 - (1 << lv) + 1]);
      int pos = p0;  // the split: the one pair of the range whose delta is the minimum
#pragma unroll
      for (int v = 7; v >= 0; --v)
        if (pos + (1 << v) - 1 <= p1 && s_dt[v][pos] > dnode) pos += 1 << v;
      const int gamma = s - REFIT_THREADS + pos;
      lo = d > 0 ? k : k - l;
      hi = d > 0 ? k + l : k;
      left = (lo == gamma) ? (n - 1 + gamma) : gamma;
      right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
      mine = lo >= s && hi < e;  // fitted here iff the whole range lies in this block
      if (!mine) {  // the upper tree is the only reader of these
        t.topo[k] = make_uint4((uint32_t)left, (uint32_t)right, (uint32_t)lo, (uint32_t)hi);
        t.parent[left] = (uint32_t)k;
        t.parent[right] = (uint32_t)k;
        t.flags[k] = 0;
      }
    }
  }
  // ---- the leaf: record and box ----
  float lmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};  // identity of min / max past the end
  if (leaf_ok) {
    TriRec r;
    r.v0 = make_float4(a.x, a.y, a.z, __uint_as_float(tri));
    r.v1 = make_float4(b.x, b.y, b.z, 0.f);
    r.v2 = make_float4(c.x, c.y, c.z, 0.f);
    recs[k] = r;
    lmn[0] = fminf(a.x, fminf(b.x, c.x)); lmn[1] = fminf(a.y, fminf(b.y, c.y)); lmn[2] = fminf(a.z, fminf(b.z, c.z));
    lmx[0] = fmaxf(a.x, fmaxf(b.x, c.x)); lmx[1] = fmaxf(a.y, fmaxf(b.y, c.y)); lmx[2] = fmaxf(a.z, fmaxf(b.z, c.z));
  }
  if (n == 1) return;
#pragma unroll
  for (int v = 0; v < 3; ++v) { s_tab[0][v][tid] = lmn[v]; s_tab[0][3 + v][tid] = lmx[v]; }
  s_lpar[tid] = 0;
  s_ipar[tid] = 0;
  if (tid == 0) { s_nclimb = 0; s_ndefer = 0; }
  const int rlo = lo - s, rhi = hi - s;                          // range relative to the block (mine only)
  const int level = mine ? 31 - __clz(rhi - rlo + 1) : 0;        // 2^level <= length < 2^(level + 1), length >= 2
  __syncthreads();  // table level 0 and the cleared flags
  if (mine) {  // the children of a node fitted here lie in this block
    if ((uint32_t)left >= first_leaf) s_lpar[left - (int)first_leaf - s] = 1; else s_ipar[left - s] = 1;
    if ((uint32_t)right >= first_leaf) s_lpar[right - (int)first_leaf - s] = 1; else s_ipar[right - s] = 1;
  }
  float imn[3] = {0.f, 0.f, 0.f}, imx[3] = {0.f, 0.f, 0.f};
  int cur = 0;
#pragma unroll 1
  for (int j = 1; j <= 8; ++j) {
    const int other = min(tid + (1 << (j - 1)), REFIT_THREADS - 1);  // past the end: a repeated or an identity entry, harmless for min / max
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      s_tab[cur ^ 1][v][tid] = fminf(s_tab[cur][v][tid], s_tab[cur][v][other]);
      s_tab[cur ^ 1][3 + v][tid] = fmaxf(s_tab[cur][3 + v][tid], s_tab[cur][3 + v][other]);
    }
    __syncthreads();
    cur ^= 1;
    if (mine && level == j) {
      const int second = rhi + 1 - (1 << j);
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        imn[v] = fminf(s_tab[cur][v][rlo], s_tab[cur][v][second]);
        imx[v] = fmaxf(s_tab[cur][3 + v][rlo], s_tab[cur][3 + v][second]);
      }
      t.box[2 * (size_t)(k)] = make_float4(imn[0], imn[1], imn[2], __int_as_float(left));
      t.box[2 * (size_t)(k) + 1] = make_float4(imx[0], imx[1], imx[2], __int_as_float(right));
    }
  }
  // ---- list the block's top nodes: leaf k and / or inner node k whose parent is not fitted here ----
  // (the flags were written before the first barrier of the loop above).  One global atomic per block reserves a
  // contiguous piece of ONE dense list, so climb_kernel runs with full warps.
  uint32_t slot_leaf = 0xFFFFFFFFu, slot_node = 0xFFFFFFFFu, slot_defer = 0xFFFFFFFFu;
  if (leaf_ok && !s_lpar[tid]) slot_leaf = atomicAdd(&s_nclimb, 1u);
  if (mine && !s_ipar[tid]) slot_node = atomicAdd(&s_nclimb, 1u);
  if (deferred) slot_defer = atomicAdd(&s_ndefer, 1u);
  __syncthreads();
  if (tid == 0) {
    const uint32_t base = atomicAdd(&status[1], s_nclimb);
    if (s_nclimb > (uint32_t)CLIMB_SLOTS || base + s_nclimb > climb_cap) { status[0] = 1u; s_base = 0xFFFFFFFFu; }
    else s_base = base;
    s_dbase = s_ndefer ? atomicAdd(&status[2], s_ndefer) : 0u;
  }
  __syncthreads();
  if (slot_defer != 0xFFFFFFFFu) deferred_nodes[s_dbase + slot_defer] = (uint32_t)k;  // at most one entry per node: the list holds n
  const uint32_t base = s_base;
  if (base == 0xFFFFFFFFu) return;
  if (slot_leaf != 0xFFFFFFFFu) climbers[base + slot_leaf] = first_leaf + (uint32_t)k;
  if (slot_node != 0xFFFFFFFFu) climbers[base + slot_node] = (uint32_t)k;
}

// The nodes tree_fit_kernel deferred (range longer than a block, ~n / 256 of them): the complete Karras search, every lane
// on a long search of its own.
__global__ void __launch_bounds__(256) upper_tree_kernel(const uint32_t* __restrict__ deferred_nodes, const uint32_t* __restrict__ count, const uint64_t* __restrict__ keys,
                                                          uint64_t key_mask, int n, BinTree t) {
  const uint32_t total = *count;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    const int k = (int)deferred_nodes[w];
    const uint64_t ki = __ldg(keys + k) & key_mask;
    const int d = (delta(ki, keys, key_mask, n, k, k + 1) - delta(ki, keys, key_mask, n, k, k - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(ki, keys, key_mask, n, k, k - d);
    int lmax = 2 * REFIT_THREADS;  // tree_fit_kernel has verified delta(k, k + 256 d) > dmin
    while (delta(ki, keys, key_mask, n, k, k + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int st = lmax >> 1; st >= 1; st >>= 1)
      if (delta(ki, keys, key_mask, n, k, k + (l + st) * d) > dmin) l += st;
    const int j = k + l * d;
    const int dnode = delta(ki, keys, key_mask, n, k, j);
    int sp = 0;
    int sh = 1;
    int tstep = (l + 1) >> 1;
    while (true) {
      if (delta(ki, keys, key_mask, n, k, k + (sp + tstep) * d) > dnode) sp += tstep;
      if (tstep == 1) break;
      ++sh;
      tstep = (l + (1 << sh) - 1) >> sh;
    }
    const int gamma = k + sp * d + min(d, 0);
    const int lo = min(k, j), hi = max(k, j);
    const int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    t.topo[k] = make_uint4((uint32_t)left, (uint32_t)right, (uint32_t)lo, (uint32_t)hi);
    t.parent[left] = (uint32_t)k;
    t.parent[right] = (uint32_t)k;
    t.flags[k] = 0;
  }
}

// The upper tree: one thread per listed top node.  Boxes, topology, parents and arrival flags are complete
// (tree_fit_kernel has finished).
__global__ void __launch_bounds__(256) climb_kernel(const uint32_t* __restrict__ climbers, const uint32_t* __restrict__ count, int n, BinTree t, const TriRec* recs) {
  const uint32_t first_leaf = (uint32_t)(n - 1);
  const uint32_t total = *count;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t cur_node = climbers[i];
    float4 mn, mx;
    node_box(t, recs, first_leaf, cur_node, mn, mx);
    climb(t, recs, first_leaf, cur_node, mn, mx);
  }
}

// ---- 6. collapse to 8-wide quantised nodes ---------------------------------------------------
struct WorkItem { uint32_t bin, wide, lo, hi; };  // binary node, the wide node to fill, the node's range of sorted records

__device__ __forceinline__ uint32_t pick_exponent(float extent) {
  // smallest biased exponent e with extent <= 255 * 2^(e-127) (plus one for rounding slack)
  float s = extent / 255.f;
  uint32_t bits = __float_as_uint(s);
  uint32_t e = (bits >> 23) & 0xffu;
  if (bits & 0x7fffffu) e += 1;
  if (e < 1) e = 1;
  if (e > 230) e = 230;  // the node stores step * 2^15 and the fit loop may still raise e
  return e;
}

// A candidate child is opened further only while it owns more than J3DG_LEAF_KEEP triangles: the
// traversal kernel tests all (<= 8) triangles of a leaf in ONE lane-parallel round, so very small
// leaves only add node levels.  Measured on config B with three frames in flight (profiles/r2_cast_refill_leafkeep_knobs.log):
// 4 / 2 / 1 -> 0.771 / 0.757 / 0.756 ms per frame.
#ifndef J3DG_LEAF_KEEP
#define J3DG_LEAF_KEEP 2
#endif

__device__ __forceinline__ void mark_leaf_end(TriRec* __restrict__ recs, uint32_t last) {
  reinterpret_cast<uint32_t*>(&recs[last].v1)[3] = 1u;
}

// One thread per wide node.  The kernel is bound by the latency of DEPENDENT loads (ncu, round 2: 3.9 warps per scheduler, 0.08
// eligible, 85 % of the stall cycles on L1TEX), so the work is arranged to keep the chain short:
//   * a candidate keeps everything that was loaded with it — children, leaf range, box — in registers (the arrays are only
//     indexed with unrolled compile-time indices), so opening a candidate needs no load for ITS topology, only one round of
//     independent loads (topology + box of both children): one latency per opening, ~9 per node instead of ~60;
//   * the wide-node indices and the queue slots of all children of all 32 nodes of a warp are taken with ONE pair of atomics
//     (siblings end up adjacent in memory);
//   * the level counters rotate over three words (read, append, zero for the next level): no reset launches.
struct Cand {
  uint32_t id, l, r, lo, hi;  // binary node, its children, its range of sorted triangle records
  float area;                 // < 0: not opened any further
  float mn[3], mx[3];
};

// Children (a, b) of a node that owns records [lo, hi]: the split follows from the left child's number (inner node gamma or
// leaf first_leaf + gamma), so ranges never have to be stored.  Every load is issued before the first use, without branches:
// an inner node is one 32-byte sector {min, left | max, right}, a leaf its 48-byte record.
__device__ __forceinline__ void load_pair(const BinTree& t, const TriRec* recs, uint32_t first_leaf, uint32_t a, uint32_t b, uint32_t lo, uint32_t hi,
                                          Cand& ca, Cand& cb) {
  const bool la = a >= first_leaf, lb = b >= first_leaf;
  const uint32_t gamma = la ? a - first_leaf : a;
  const float4* pa = la ? reinterpret_cast<const float4*>(recs + lo) : t.box + 2 * (size_t)a;
  const float4* pb = lb ? reinterpret_cast<const float4*>(recs + hi) : t.box + 2 * (size_t)b;
  const float4 a0 = pa[0], a1 = pa[1], a2 = pa[la ? 2 : 1];
  const float4 b0 = pb[0], b1 = pb[1], b2 = pb[lb ? 2 : 1];
  ca.id = a; ca.lo = lo; ca.hi = gamma;
  ca.l = __float_as_uint(a0.w); ca.r = __float_as_uint(a1.w);
  ca.mn[0] = la ? fminf(a0.x, fminf(a1.x, a2.x)) : a0.x; ca.mn[1] = la ? fminf(a0.y, fminf(a1.y, a2.y)) : a0.y; ca.mn[2] = la ? fminf(a0.z, fminf(a1.z, a2.z)) : a0.z;
  ca.mx[0] = la ? fmaxf(a0.x, fmaxf(a1.x, a2.x)) : a1.x; ca.mx[1] = la ? fmaxf(a0.y, fmaxf(a1.y, a2.y)) : a1.y; ca.mx[2] = la ? fmaxf(a0.z, fmaxf(a1.z, a2.z)) : a1.z;
  {
    const float dx = ca.mx[0] - ca.mn[0], dy = ca.mx[1] - ca.mn[1], dz = ca.mx[2] - ca.mn[2];
    ca.area = (la || ca.hi - ca.lo + 1u <= (uint32_t)J3DG_LEAF_KEEP) ? -1.f : fmaxf(dx * dy + dy * dz + dz * dx, 0.f);
  }
  cb.id = b; cb.lo = gamma + 1u; cb.hi = hi;
  cb.l = __float_as_uint(b0.w); cb.r = __float_as_uint(b1.w);
  cb.mn[0] = lb ? fminf(b0.x, fminf(b1.x, b2.x)) : b0.x; cb.mn[1] = lb ? fminf(b0.y, fminf(b1.y, b2.y)) : b0.y; cb.mn[2] = lb ? fminf(b0.z, fminf(b1.z, b2.z)) : b0.z;
  cb.mx[0] = lb ? fmaxf(b0.x, fmaxf(b1.x, b2.x)) : b1.x; cb.mx[1] = lb ? fmaxf(b0.y, fmaxf(b1.y, b2.y)) : b1.y; cb.mx[2] = lb ? fmaxf(b0.z, fmaxf(b1.z, b2.z)) : b1.z;
  {
    const float dx = cb.mx[0] - cb.mn[0], dy = cb.mx[1] - cb.mn[1], dz = cb.mx[2] - cb.mn[2];
    cb.area = (lb || cb.hi - cb.lo + 1u <= (uint32_t)J3DG_LEAF_KEEP) ? -1.f : fmaxf(dx * dy + dy * dz + dz * dx, 0.f);
  }
}

// counts[0..2]: queue sizes, rotating (level L reads [L % 3], appends to [(L + 1) % 3], zeroes [(L + 2) % 3]); counts[3] node count, counts[4] overflow
#ifndef J3DG_COLLAPSE_MIN_BLOCKS
#define J3DG_COLLAPSE_MIN_BLOCKS 3   // 168 registers, three blocks per SM (4.66 -> 4.59 ms; 4 blocks: 4.68)
#endif
__global__ void __launch_bounds__(128, J3DG_COLLAPSE_MIN_BLOCKS) collapse_kernel(BinTree t, int n, const WorkItem* __restrict__ in, WorkItem* __restrict__ out, uint32_t* __restrict__ counts, int level,
                                                        WideNode* __restrict__ nodes, uint32_t node_cap, TriRec* __restrict__ recs) {
  const uint32_t count = counts[level % 3];
  uint32_t* out_count = counts + (level + 1) % 3;
  uint32_t* node_count = counts + 3;
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[(level + 2) % 3] = 0u;  // the append counter of the NEXT level; nobody reads or writes it during this one
  const uint32_t first_leaf = (uint32_t)(n - 1);
  const uint32_t lane = threadIdx.x & 31u;
  // warp-uniform trip count: all 32 lanes stay in the loop (full-mask shuffles below)
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < count; base += gridDim.x * blockDim.x) {
    const uint32_t w = base + lane;
    const bool valid = w < count;
    const WorkItem item = valid ? in[w] : WorkItem{0u, 0u, 0u, first_leaf};  // idle lanes walk the root (no stores)
    const float4 nmn = t.box[2 * (size_t)(item.bin)], nmx = t.box[2 * (size_t)(item.bin) + 1];  // {min, left child}, {max, right child}
    Cand c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i].id = 0; c[i].l = c[i].r = c[i].lo = c[i].hi = 0; c[i].area = -1.f; for (int j = 0; j < 3; ++j) { c[i].mn[j] = 0.f; c[i].mx[j] = 0.f; } }
    int nc = 2;
    load_pair(t, recs, first_leaf, __float_as_uint(nmn.w), __float_as_uint(nmx.w), item.lo, item.hi, c[0], c[1]);
    while (nc < 8) {
      int best = -1;
      float ba = -1.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nc && c[i].area > ba) { ba = c[i].area; best = i; }
      if (best < 0) break;
      uint32_t a = 0, b = 0, plo = 0, phi = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i == best) { a = c[i].l; b = c[i].r; plo = c[i].lo; phi = c[i].hi; }
      Cand ca, cb;
      load_pair(t, recs, first_leaf, a, b, plo, phi, ca, cb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i == best) c[i] = ca;
        if (i == nc) c[i] = cb;
      }
      ++nc;
    }
    WideNode node;
    node.ox = nmn.x; node.oy = nmn.y; node.oz = nmn.z;
    node.pad0 = 0u;
    uint32_t e[3] = {pick_exponent(nmx.x - nmn.x), pick_exponent(nmx.y - nmn.y), pick_exponent(nmx.z - nmn.z)};
    const float org[3] = {nmn.x, nmn.y, nmn.z};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      for (;;) {
        const float scale = __uint_as_float(e[ax] << 23);
        const float inv = 1.f / scale;
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nc && ok) {
            const float lo = c[i].mn[ax], hi = c[i].mx[ax];
            int ql = (int)floorf((lo - org[ax]) * inv);
            int qh = (int)ceilf((hi - org[ax]) * inv);
            ql = max(0, min(ql, 255));
            qh = max(0, qh);
            // conservative against the rounding of the decode fl(q*scale + origin)
            while (ql > 0 && __fmaf_rn((float)ql, scale, org[ax]) > lo) --ql;
            while (qh <= 255 && __fmaf_rn((float)qh, scale, org[ax]) < hi) ++qh;
            if (qh > 255) ok = false;
            else { node.box[i][ax] = (uint8_t)ql; node.box[i][3 + ax] = (uint8_t)qh; }
          }
        }
        if (ok) break;
        e[ax] += 1;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i >= nc) { node.box[i][ax] = 255; node.box[i][3 + ax] = 0; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { node.box[i][6] = 0x80; node.box[i][7] = 0x3F; }
    node.sx = __uint_as_float((e[0] + 15u) << 23); node.sy = __uint_as_float((e[1] + 15u) << 23); node.sz = __uint_as_float((e[2] + 15u) << 23);  // step * 2^15
    node.nchild = (uint32_t)nc;
    // children: leaves refer to their record range, the others get a wide node and a queue slot
    uint32_t need = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      node.child[i] = J3DG_EMPTY_CHILD;
      if (valid && i < nc) {
        if (c[i].hi - c[i].lo + 1u <= (uint32_t)J3DG_MAX_LEAF) {
          node.child[i] = J3DG_LEAF_BIT | c[i].lo;
          mark_leaf_end(recs, c[i].hi);
        } else {
          need |= 1u << i;
        }
      }
    }
    const uint32_t k = (uint32_t)__popc(need);
    uint32_t incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t wbase = 0, qbase = 0;
    if (lane == 0 && total) { wbase = atomicAdd(node_count, total); qbase = atomicAdd(out_count, total); }
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    qbase = __shfl_sync(0xffffffffu, qbase, 0);
    uint32_t wi = wbase + incl - k, qi = qbase + incl - k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (need & (1u << i)) {
        if (wi >= node_cap) {  // node array too small: the host retries with the hard bound
          counts[4] = 1u;
          node.box[i][0] = 255; node.box[i][3] = 0;
        } else {
          node.child[i] = wi;
          out[qi] = WorkItem{c[i].id, wi, c[i].lo, c[i].hi};
        }
        ++wi; ++qi;
      }
    }
    if (valid) nodes[item.wide] = node;
  }
}

// n == 1: a root with a single leaf child
__global__ void single_triangle_root_kernel(WideNode* nodes, TriRec* recs) {
  const float4 a = recs[0].v0, b = recs[0].v1, c = recs[0].v2;
  const float4 mn = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
  const float4 mx = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
  WideNode node;
  node.ox = mn.x; node.oy = mn.y; node.oz = mn.z;
  node.pad0 = 0u;
  const float ext[3] = {mx.x - mn.x, mx.y - mn.y, mx.z - mn.z};
  uint32_t e[3];
  for (int ax = 0; ax < 3; ++ax) {
    e[ax] = pick_exponent(ext[ax]) + 1;
    for (int i = 0; i < 8; ++i) { node.box[i][ax] = 255; node.box[i][3 + ax] = 0; }
    node.box[0][ax] = 0;
    node.box[0][3 + ax] = 255;
  }
  for (int i = 0; i < 8; ++i) { node.box[i][6] = 0x80; node.box[i][7] = 0x3F; }
  node.sx = __uint_as_float((e[0] + 15u) << 23); node.sy = __uint_as_float((e[1] + 15u) << 23); node.sz = __uint_as_float((e[2] + 15u) << 23);  // step * 2^15
  node.nchild = 1;
  for (int i = 0; i < 8; ++i) node.child[i] = J3DG_EMPTY_CHILD;
  node.child[0] = J3DG_LEAF_BIT | 0u;
  mark_leaf_end(recs, 0);
  nodes[0] = node;
}

__global__ void init_queue_kernel(WorkItem* q, uint32_t* counts, uint32_t n) {
  q[0] = WorkItem{0u, 0u, 0u, n - 1u};
  counts[0] = 1u;  // queue sizes of levels 0, 1, 2 (rotating)
  counts[1] = 0u;
  counts[2] = 0u;
  counts[3] = 1u;  // wide nodes
  counts[4] = 0u;  // overflow
}

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = (T*)(base + off);
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace

// CUDA loads kernels lazily, at their first launch (about a millisecond each): the first j3dg_mesh_create of a process
// paid ~25 ms for that.  j3dg_ctx_create calls this from its background thread instead.
void j3dg_preload_build_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, bbox_init_kernel); cudaFuncGetAttributes(&a, bbox_kernel); cudaFuncGetAttributes(&a, morton_kernel);
  cudaFuncGetAttributes(&a, rsort::histogram_kernel); cudaFuncGetAttributes(&a, rsort::scan_chunk_sums); cudaFuncGetAttributes(&a, rsort::scan_sums_serial);
  cudaFuncGetAttributes(&a, rsort::scan_apply); cudaFuncGetAttributes(&a, rsort::scatter_kernel);
  cudaFuncGetAttributes(&a, rsort::digit_histograms_kernel); cudaFuncGetAttributes(&a, rsort::onesweep_kernel<true>); cudaFuncGetAttributes(&a, rsort::onesweep_kernel<false>);
  cudaFuncGetAttributes(&a, tree_fit_kernel); cudaFuncGetAttributes(&a, upper_tree_kernel); cudaFuncGetAttributes(&a, climb_kernel); cudaFuncGetAttributes(&a, collapse_kernel);
  cudaFuncGetAttributes(&a, init_queue_kernel);
  cudaGetLastError();
}

// Builds m->d_nodes / m->d_tris from m->d_vertices / m->d_indices (device resident).
int j3dg_build_bvh(j3dg_mesh* m) {
  j3dg_ctx* ctx = m->ctx;
  const uint32_t n = m->nt;
  cudaStream_t st = ctx->stream;
  m->info.nr_of_vertices = m->nv;
  m->info.nr_of_triangles = n;
  m->info.nr_of_leaf_triangles = n;
  m->info.node_bytes = sizeof(WideNode);
  m->info.triangle_bytes = sizeof(TriRec);
  m->info.sah_cost = 0.f;

  // ---- scratch arena ----
  const size_t nn = n ? n : 1;
  size_t need = 4096;
  need += 2 * 256 + 8 * sizeof(uint32_t) + 256;                     // bbox + counters + size accumulator
  need += 2 * (256 + nn * sizeof(uint64_t)) + 2 * (256 + nn * sizeof(uint32_t));  // keys/vals ping-pong
  need += 256 + rsort::scratch_bytes(n);
  need += 256 + nn * sizeof(int2) + 256 + nn * sizeof(uint2) + 256 + 2 * nn * sizeof(uint32_t) + 256 + nn * sizeof(uint32_t);
  need += 2 * (256 + 2 * nn * sizeof(float4));  // boxes (inner nodes) + slack
  need += 2 * (256 + nn * sizeof(WorkItem));
  need += 512 + ((nn + REFIT_THREADS - 1) / REFIT_THREADS) * (CLIMB_SLOTS + 1) * sizeof(uint32_t);
  if (j3dg_reserve(ctx, &ctx->d_misc, &ctx->misc_cap, need) != J3DG_OK) return J3DG_ENOMEM;
  Arena ar;
  ar.base = (char*)ctx->d_misc;
  ar.cap = ctx->misc_cap;
  uint32_t* d_bb = ar.take<uint32_t>(8);
  uint32_t* d_counts = ar.take<uint32_t>(8);  // [0..2] queue sizes (rotating over the levels), [3] node count, [4] node overflow, [5] climb-list overflow, [6] listed top nodes, [7] deferred nodes
  unsigned long long* d_size_acc = ar.take<unsigned long long>(2);
  uint64_t* keys_a = ar.take<uint64_t>(nn);
  uint64_t* keys_b = ar.take<uint64_t>(nn);
  uint32_t* vals_a = ar.take<uint32_t>(nn);
  uint32_t* vals_b = ar.take<uint32_t>(nn);
  uint32_t* sort_scratch = ar.take<uint32_t>(rsort::scratch_bytes(n) / 4);
  BinTree bt;
  bt.topo = ar.take<uint4>(nn);
  bt.parent = ar.take<uint32_t>(2 * nn);
  bt.flags = ar.take<uint32_t>(nn);
  bt.box = ar.take<float4>(2 * nn);
  uint32_t* d_climbers = ar.take<uint32_t>(((nn + REFIT_THREADS - 1) / REFIT_THREADS) * CLIMB_SLOTS);
  WorkItem* q0 = ar.take<WorkItem>(nn);
  uint32_t* d_deferred = reinterpret_cast<uint32_t*>(q0);  // the collapse queues are idle until the upper tree is done
  WorkItem* q1 = ar.take<WorkItem>(nn);

  // ---- output arrays ----
  if (!m->d_tris && n) {
    if (cudaMalloc((void**)&m->d_tris, ((size_t)n + J3DG_TRI_PAD) * sizeof(TriRec)) != cudaSuccess) {
      j3dg_set_error(ctx, "out of device memory (triangle records)");
      return J3DG_ENOMEM;
    }
    CU_CHECK(ctx, cudaMemsetAsync(m->d_tris + n, 0, J3DG_TRI_PAD * sizeof(TriRec), st));
  }
  uint32_t cap = std::max<uint32_t>(16u, n / 3 + 1024u);
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (m->node_cap < cap) {
      if (m->d_nodes) cudaFree(m->d_nodes);
      m->d_nodes = nullptr;
      if (cudaMalloc((void**)&m->d_nodes, (size_t)cap * sizeof(WideNode)) != cudaSuccess) {
        j3dg_set_error(ctx, "out of device memory (BVH nodes)");
        return J3DG_ENOMEM;
      }
      m->node_cap = cap;
    }
    CU_CHECK(ctx, cudaEventRecord(ctx->ev[6], st));
    CU_CHECK(ctx, cudaMemsetAsync(d_counts, 0, 8 * sizeof(uint32_t), st));  // [5]: climb-list overflow, set by tree_fit_kernel
    // 1. bbox
    bbox_init_kernel<<<1, 32, 0, st>>>(d_bb);
    KERNEL_CHECK(ctx);
    if (m->nv) {
      int blocks = (int)std::min<size_t>(((size_t)m->nv + 255) / 256, (size_t)ctx->sm_count * 8);
      blocks = (blocks + 2) / 3 * 3;  // grid stride a multiple of 3 (bbox_kernel)
      bbox_kernel<<<blocks, 256, 0, st>>>(m->d_vertices, m->nv, d_bb);
      KERNEL_CHECK(ctx);
    }
    uint32_t h_counts[8] = {0, 0, 0, 1, 0, 0, 0, 0};
    if (n) {
      const uint32_t tb = (n + 255) / 256;
      CU_CHECK(ctx, cudaMemsetAsync(d_size_acc, 0, 2 * sizeof(unsigned long long), st));
      morton_kernel<<<tb, 256, 0, st>>>(m->d_vertices, m->d_indices, n, d_bb, keys_a, d_size_acc);
      KERNEL_CHECK(ctx);
      // Morton resolution worth sorting: a cell between a quarter and a half of the typical (geometric-mean) triangle
      // extent.  Finer bits do not change the order (measured on the 28 M-triangle mesh: 16 ... 12 bits per axis give the
      // identical tree, 10 bits a worse one), and every 8 bits less is one sort pass less.  Keys keep all 48 bits; the low ones are simply not sorted and masked in the radix tree.
      unsigned long long h_acc[2] = {0, 0};
      CU_CHECK(ctx, cudaMemcpyAsync(h_acc, d_size_acc, sizeof(h_acc), cudaMemcpyDeviceToHost, st));
      CU_CHECK(ctx, cudaStreamSynchronize(st));
      const double mean_bits = (double)h_acc[0] / 256.0 / (double)std::max<unsigned long long>(h_acc[1], 1);
      int axis_bits = std::min(MORTON_BITS_PER_AXIS, std::max(10, (int)std::ceil(mean_bits) + 1));
      if (const char* e = getenv("J3DG_MORTON_BITS")) axis_bits = std::min(MORTON_BITS_PER_AXIS, std::max(1, atoi(e)));  // developer knob
      const int passes = (3 * axis_bits + 7) / 8;
      int first_bit = MORTON_BITS - 8 * passes;  // sort the top 8 * passes bits
      uint64_t key_mask = ~0ull << first_bit;
      bool in_b = false;
      // When the sorted code bits and the triangle index fit one 64-bit word together (config B: 36 bits are worth sorting, five
      // passes would sort 40, 39 + 25 index bits fit: the top digit is then only partly used), the sort moves 8-byte packed keys
      // instead of 12-byte (key, value) pairs.
      // Same order (equal codes keep their input order either way), same tree: the radix tree masks the index bits.
      int idx_bits = 1;
      while (idx_bits < 32 && (1ull << idx_bits) < (unsigned long long)n) ++idx_bits;
      int sorted_bits = 8 * passes;
      if (sorted_bits + idx_bits > 64 && 3 * axis_bits + idx_bits <= 64) sorted_bits = 64 - idx_bits;  // still at least the bits worth sorting
      const bool packed = rsort::can_sort_packed(n, passes, idx_bits, sorted_bits);
      if (packed) { first_bit = MORTON_BITS - sorted_bits; key_mask = ~0ull << first_bit; }
      int rc = packed ? rsort::sort_packed(ctx, keys_a, keys_b, n, first_bit, passes, idx_bits, sort_scratch, &in_b)
                      : rsort::sort_pairs(ctx, keys_a, vals_a, keys_b, vals_b, n, MORTON_BITS, sort_scratch, &in_b, true, first_bit);
      if (rc != J3DG_OK) return rc;
      if (packed) key_mask = ~0ull << idx_bits;
      const uint64_t* keys = in_b ? keys_b : keys_a;
      const uint32_t* vals = packed ? nullptr : (in_b ? vals_b : vals_a);
      tree_fit_kernel<<<tb, REFIT_THREADS, 0, st>>>(m->d_vertices, m->d_indices, vals, keys, key_mask, (uint32_t)((1ull << idx_bits) - 1ull), (int)n, bt, m->d_tris, d_climbers,
                                                    tb * (uint32_t)CLIMB_SLOTS, d_deferred, d_counts + 5);
      KERNEL_CHECK(ctx);
      if (n > 1) {
        upper_tree_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_deferred, d_counts + 7, keys, key_mask, (int)n, bt);
        KERNEL_CHECK(ctx);
        climb_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_climbers, d_counts + 6, (int)n, bt, m->d_tris);
        KERNEL_CHECK(ctx);
      }
      if (n == 1) {
        single_triangle_root_kernel<<<1, 1, 0, st>>>(m->d_nodes, m->d_tris);
        KERNEL_CHECK(ctx);
      } else {
        init_queue_kernel<<<1, 1, 0, st>>>(q0, d_counts, n);
        KERNEL_CHECK(ctx);
        // Level-synchronous collapse without host round trips: a fixed-size grid strides over
        // the device-side queue; the loop runs until the host sees an empty level (checked after
        // 12 levels, then every 4, to keep syncs rare).
        const int grid = ctx->sm_count * 8;
        int level = 0;
        for (;;) {
          for (int k = 0, chunk = level == 0 ? 12 : 4; k < chunk; ++k, ++level) {  // 8^11 > 2^32: twelve levels cover any balanced tree, unbalanced ones go on in fours
            WorkItem* qi = (level & 1) ? q1 : q0;
            WorkItem* qo = (level & 1) ? q0 : q1;
            collapse_kernel<<<grid, 128, 0, st>>>(bt, (int)n, qi, qo, d_counts, level, m->d_nodes, m->node_cap, m->d_tris);
            KERNEL_CHECK(ctx);
          }
          CU_CHECK(ctx, cudaMemcpyAsync(h_counts, d_counts, sizeof(h_counts), cudaMemcpyDeviceToHost, st));
          CU_CHECK(ctx, cudaStreamSynchronize(st));
          if (h_counts[level % 3] == 0 || h_counts[4]) break;
          if (level > 4096) { j3dg_set_error(ctx, "BVH collapse did not terminate"); return J3DG_ECUDA; }
        }
      }
    }
    CU_CHECK(ctx, cudaEventRecord(ctx->ev[7], st));
    CU_CHECK(ctx, cudaEventSynchronize(ctx->ev[7]));
    if (h_counts[5]) { j3dg_set_error(ctx, "BVH build: more top nodes in a block than CLIMB_SLOTS"); return J3DG_ECUDA; }
    if (h_counts[4]) {  // node array too small for a degenerate tree: retry with the hard bound
      cap = std::max<uint32_t>(16u, n);
      continue;
    }
    float ms = 0.f;
    CU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
    m->info.build_ms = ms;
    m->nr_nodes = n ? h_counts[3] : 0;
    m->info.nr_of_nodes = m->nr_nodes;
    uint32_t h_bb[6];
    CU_CHECK(ctx, cudaMemcpy(h_bb, d_bb, sizeof(h_bb), cudaMemcpyDeviceToHost));
    for (int j = 0; j < 3; ++j) {
      m->info.bbox_min[j] = m->nv ? ordered_to_float(h_bb[j]) : 0.f;
      m->info.bbox_max[j] = m->nv ? ordered_to_float(h_bb[3 + j]) : 0.f;
    }
    return J3DG_OK;
  }
  j3dg_set_error(ctx, "BVH node allocation overflow");
  return J3DG_ECUDA;
}
