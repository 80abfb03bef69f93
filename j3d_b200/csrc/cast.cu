// cast.cu — per-pixel ray cast.  Replaces canvas::update_canvas (j3d/canvas.cpp:677-874):
// ray generation (773-782), qbvh_two_level_with_transformations::find_closest_triangle
// (jtk/qbvh.h:3303-3387) -> qbvh::find_closest_triangle (1701-1852) with the Woop test
// (4793-4869), hit -> pixel record (788-834) and the shadow ray (836-857).
//
// Two traversal kernels share one BVH (8-wide, 128-byte child-major nodes, leaves of <= 8 triangles):
//
//  lane_kernel  — one ray per LANE, one warp per 8x4-pixel tile, persistent warps pulling tiles from
//                 a global counter, while-while traversal (all lanes descend nodes, then all test
//                 leaves), nearest child in registers, short stack in shared memory.  This is the
//                 throughput path: fewest instructions per ray.  Measured on the 28 M-triangle mesh it
//                 finishes 98.7 % of the tiles in the first half of its run time and then waits for a
//                 few silhouette tiles whose 32 grazing rays visit 100-250 nodes each in lockstep
//                 (profiles/README.md).  So every ray gets a BUDGET of node visits; a ray that exceeds
//                 it is evicted to a "hard ray" list together with its best hit so far.
//  group_kernel — one ray per 8 LANES: lane c tests child c of the current node (or triangle c of the
//                 current leaf), the group votes, picks the nearest child with three shuffle-min steps
//                 and pushes the other hit children on ONE stack per ray in shared memory.  A traversal
//                 step is ~8x shorter than in the lane kernel, rays are handed to groups one by one, so
//                 there is no lockstep tail.  This is the latency path: it finishes the hard rays
//                 (restarted from the root, pruned by the best hit the lane kernel already found) and
//                 serves the generic find_closest query.
//
// Pipeline of one j3dg_cast: lane<PRIMARY> -> group<PRIMARY, hard list> (raw hit: t, u, v, record slot)
// -> resolve_kernel (hit -> pixel record; appends shadow rays) -> lane<SHADOW> -> group<SHADOW, hard list>.
//
// Parity rules: the ray, the Woop edge functions, t/u/v, the triangle normal and
// its two transforms are evaluated with separately rounded mul/add/sub (no FMA), in the
// reference's operation order.  Only the box tests use FMA — they are conservative and do not
// influence which triangle is the closest hit.
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace {

constexpr int GROUP = 8;                                  // group kernel: lanes per ray = children per node = max triangles per leaf
constexpr int BLOCK_THREADS = 128;
constexpr int LANE_SM_STACK = 12;                         // lane kernel: stack entries per ray in shared memory (12 * 8 B * 128 = 12 KB per block) ...
constexpr int LANE_STACK = 96;                            // ... of this many in total (the rest in local memory, rarely touched)
#ifndef J3DG_LANE_MIN_BLOCKS
#define J3DG_LANE_MIN_BLOCKS 8
#endif
constexpr int GROUPS_PER_BLOCK = BLOCK_THREADS / GROUP;   // 16 rays in flight per block
constexpr int STACK_SIZE = 96;                            // entries per ray; 96 * 8 B * 16 = 12 KB shared memory per block
constexpr int TILE_W = 8, TILE_H = 4;                     // a warp's ray pool = one 8x4 pixel tile
constexpr int SUPER_W = 4, SUPER_H = 8;                   // tiles are numbered super-tile by super-tile (32x32 pixels)
#ifndef J3DG_CAST_MIN_BLOCKS
#define J3DG_CAST_MIN_BLOCKS 8
#endif

enum Mode { PRIMARY = 0, SHADOW = 1, RAYLIST = 2 };

struct TraceParams {
  const MeshDev* meshes;
  uint32_t nm;
  ViewDev vw;
  int x0, y0, x1, y1;
  j3dg_pixel* out;               // PRIMARY: raw hits are written here; SHADOW: mark bit 0 is set here
  uint32_t stride;
  unsigned long long* stats;     // [0] node rounds [1] triangle tests [2] overflow flag [3] pool counter [4] shadow rays (accumulating) [5] shadow list length
  const float4* shadow_pos;      // SHADOW: ray origins (xyzw as the reference computes them)
  const uint32_t* shadow_pix;    // SHADOW: pixel offset (y * stride + x) of each ray
  // hard-ray hand-over between the two kernels
  uint32_t budget;               // lane kernel: node visits a ray may spend before it is evicted
  unsigned int* pool_ctr;        // pool counter of THIS launch
  unsigned int* hard_count;      // lane kernel: append position; group kernel (list mode): number of entries
  uint2* hard_id;                // {ray id, mesh of the best hit so far}
  float4* hard_best;             // {t, u, v, record slot bits} of the best hit so far (slot 0xFFFFFFFF: none)
  const float* rays;             // RAYLIST: n x 8 floats
  float* hits;                   // RAYLIST: n x 4 floats
  uint32_t* ids;                 // RAYLIST: n triangle ids
  uint32_t nrays;
};

__device__ __forceinline__ float pick(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ float safe_rcp(float d) {
  // box tests only: keep the reciprocal finite so 0 * inf never produces NaN
  const float big = 1e18f;
  if (fabsf(d) < 1e-18f) return d < 0.f ? -big : big;
  return 1.f / d;
}

__device__ __forceinline__ float4 transform_point(const float* __restrict__ m, float4 p) {  // jtk::transform, qbvh.h:5140-5151
  float4 r = mat_vec(m, p);
  if (r.w != 1.f && r.w != 0.f) { r.x = fdiv(r.x, r.w); r.y = fdiv(r.y, r.w); r.z = fdiv(r.z, r.w); r.w = 1.f; }
  return r;
}

// min over the 8 lanes of a group (xor shuffles stay inside an aligned group of 8)
// (executed by the whole warp: the four groups reduce side by side)
__device__ __forceinline__ float group_min(float v) {
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}

// byte -> float without the quarter-rate I2F (XU pipe): PRMT builds the bits of 2^23 + q from the child's
// 8-byte box record (bytes 6 and 7 hold 0x00 and 0x4B), the subtraction is exact.
__device__ __forceinline__ float plane(uint32_t lo, uint32_t hi, uint32_t sel) {
  return __fsub_rn(__uint_as_float(__byte_perm(lo, hi, sel)), 8388608.f);
}

struct WorldRay { float4 org, dir; float t_near, t_far; };

// The ray of slot `id` in world space.  PRIMARY: id = (y << 16) | x, canvas.cpp:773-782.
template <int MODE>
__device__ __forceinline__ WorldRay world_ray(const TraceParams& p, uint32_t id) {
  WorldRay r;
  if (MODE == PRIMARY) {
    const int x = (int)(id & 0xffffu), y = (int)(id >> 16);
    const float w = (float)p.vw.width, h = (float)p.vw.height;
    float4 sp;
    sp.x = fsub(fmul(2.f, fdiv(fadd((float)x, 0.5f), w)), 1.f);
    sp.y = fsub(fmul(2.f, fdiv(fadd((float)y, 0.5f), h)), 1.f);
    sp.z = p.vw.near_plane;
    sp.w = 1.f;
    float4 dir = mat_vec(p.vw.pinv, sp);
    dir.w = 0.f;
    r.dir = mat_vec(p.vw.cs, dir);
    r.org = make_float4(p.vw.origin[0], p.vw.origin[1], p.vw.origin[2], p.vw.origin[3]);
    r.t_near = fdiv(p.vw.diagonal, 100.f);
    r.t_far = FLT_MAX;
  } else if (MODE == SHADOW) {  // canvas.cpp:849-854
    const float4 pos = __ldg(p.shadow_pos + id);
    r.org = pos;
    r.dir.x = fsub(p.vw.light[0], pos.x); r.dir.y = fsub(p.vw.light[1], pos.y);
    r.dir.z = fsub(p.vw.light[2], pos.z); r.dir.w = fsub(p.vw.light[3], pos.w);
    r.t_near = 1e-3f;
    r.t_far = FLT_MAX;
  } else {
    const float* q = p.rays + 8 * (size_t)id;
    r.org = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 1.f);
    r.dir = make_float4(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5), 0.f);
    r.t_near = __ldg(q + 6);
    r.t_far = __ldg(q + 7);
  }
  return r;
}

// MODE PRIMARY: closest hit per pixel, raw result into the pixel buffer.
// MODE SHADOW : any hit (first accepted triangle ends the ray), sets mark bit 0.
// MODE RAYLIST: qbvh::find_closest_triangle semantics for arbitrary (also negative) t ranges:
//               closest = smallest |t|, bounds shrink on the side of the hit (qbvh.h:1812-1823).
// LIST: the rays are the entries of the hard-ray list the lane kernel left behind.
template <int MODE, bool LIST>
__global__ void __launch_bounds__(BLOCK_THREADS, J3DG_CAST_MIN_BLOCKS) group_kernel(const TraceParams p) {
  constexpr bool ANY_HIT = MODE == SHADOW;
  constexpr bool GENERAL = MODE == RAYLIST;
  __shared__ uint2 s_stack[STACK_SIZE * GROUPS_PER_BLOCK];
  const int lane = threadIdx.x & 31;
  const int c = lane & 7;                               // my child / triangle slot
  const int gshift = lane & 24;                         // first lane of my group
  uint2* const stk = s_stack + (threadIdx.x >> 3);      // entry i at stk[i * GROUPS_PER_BLOCK]
  const uint32_t below = (1u << c) - 1u;
  uint32_t* const overflow_flag = reinterpret_cast<uint32_t*>(p.stats + 2);

  // ---- number of ray slots ----
  uint32_t total_pools;   // a pool = 32 consecutive slots
  uint32_t supers_x = 1;
  uint32_t list_n = 0;
  if (LIST) {
    list_n = *p.hard_count;
    total_pools = (list_n + 3u) / 4u;
  } else if (MODE == PRIMARY) {
    const uint32_t tiles_x = (uint32_t)(p.x1 - p.x0 + TILE_W) / TILE_W, tiles_y = (uint32_t)(p.y1 - p.y0 + TILE_H) / TILE_H;
    supers_x = (tiles_x + SUPER_W - 1) / SUPER_W;
    total_pools = supers_x * ((tiles_y + SUPER_H - 1) / SUPER_H) * (SUPER_W * SUPER_H);
  } else {
    list_n = MODE == SHADOW ? (uint32_t)p.stats[5] : p.nrays;
    total_pools = (list_n + 31u) / 32u;
  }

  // ---- per-ray state, replicated in the 8 lanes of the group ----
  bool have_ray = false;
  uint32_t ray_id = 0;          // PRIMARY: (y << 16) | x; otherwise index into the ray list
  uint32_t mesh_k = 0;
  const WideNode* __restrict__ nodes = nullptr;
  const TriRec* __restrict__ tris = nullptr;
  float ox = 0.f, oy = 0.f, oz = 0.f, idx = 0.f, idy = 0.f, idz = 0.f;
  float Sx = 0.f, Sy = 0.f, Sz = 0.f;
  int kx = 0, ky = 1, kz = 2;
  uint32_t sel_nx = 0, sel_fx = 0, sel_ny = 0, sel_fy = 0, sel_nz = 0, sel_fz = 0;
  float t_near = 0.f, t_far = 0.f;
  float best_t = FLT_MAX, best_u = 0.f, best_v = 0.f;
  uint32_t best_slot = 0xFFFFFFFFu, best_mesh = 0;
  uint32_t cur = J3DG_EMPTY_CHILD;
  int sp = 0;

  // ---- warp-uniform pool state ----
  // A pool is what one warp fetches at a time: a 32-pixel tile, or — for the few, long hard rays — just
  // one ray per group, so that they spread over every resident warp instead of queueing in a few.
  constexpr uint32_t POOL = LIST ? 4u : 32u;
  uint32_t pool_next = POOL, pool_id = 0;
  bool exhausted = false;

  auto pop = [&]() -> uint32_t {
    while (sp > 0) {
      --sp;
      const uint2 e = stk[sp * GROUPS_PER_BLOCK];
      // entry points of popped boxes that now lie beyond the shrunk interval are skipped
      if (GENERAL || __uint_as_float(e.y) <= t_far) return e.x;
    }
    return J3DG_EMPTY_CHILD;
  };

  // object-space ray for mesh k (qbvh.h:3358-3359) + intersect_woop_precompute (qbvh.h:4793-4823)
  auto enter_mesh = [&](const WorldRay& wr, uint32_t k) {
    const MeshDev& m = p.meshes[k];
    nodes = m.nodes;
    tris = m.tris;
    const float4 d2 = mat_vec(m.cs_inv, wr.dir);
    const float4 o2 = mat_vec(m.cs_inv, wr.org);
    ox = o2.x; oy = o2.y; oz = o2.z;
    const float ax = fabsf(d2.x), ay = fabsf(d2.y), az = fabsf(d2.z);
    kz = 2;
    if (ax > ay) { if (ax > az) kz = 0; }
    else { if (ay > az) kz = 1; }
    kx = kz == 2 ? 0 : kz + 1;
    ky = kx == 2 ? 0 : kx + 1;
    const float dkz = pick(d2.x, d2.y, d2.z, kz);
    if (dkz < 0.f) { const int t = kx; kx = ky; ky = t; }
    Sz = fdiv(1.f, dkz);
    Sx = fmul(pick(d2.x, d2.y, d2.z, kx), Sz);
    Sy = fmul(pick(d2.x, d2.y, d2.z, ky), Sz);
    idx = safe_rcp(d2.x); idy = safe_rcp(d2.y); idz = safe_rcp(d2.z);
    // PRMT selectors: result bytes = {plane byte, 0x00 (byte 6), 0x00 (byte 6), 0x4B (byte 7)} = bits of 2^23 + q
    sel_nx = 0x7660u | (idx < 0.f ? 3u : 0u); sel_fx = 0x7660u | (idx < 0.f ? 0u : 3u);
    sel_ny = 0x7660u | (idy < 0.f ? 4u : 1u); sel_fy = 0x7660u | (idy < 0.f ? 1u : 4u);
    sel_nz = 0x7660u | (idz < 0.f ? 5u : 2u); sel_fz = 0x7660u | (idz < 0.f ? 2u : 5u);
    sp = 0;
    cur = m.nt ? 0u : J3DG_EMPTY_CHILD;
  };

  for (;;) {
    // =========================== (A) finish rays, move to the next mesh, refill ===========================
    if (have_ray && cur == J3DG_EMPTY_CHILD) {
      ++mesh_k;
      const bool found = best_slot != 0xFFFFFFFFu;
      if (mesh_k < p.nm && !(ANY_HIT && found)) {
        const WorldRay wr = world_ray<MODE>(p, ray_id);
        enter_mesh(wr, mesh_k);
      } else {
        // ---- write the result ----
        if (MODE == PRIMARY) {
          const int x = (int)(ray_id & 0xffffu), y = (int)(ray_id >> 16);
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)y * p.stride + x);
          if (c == 0) {  // depth; misses already carry the final record (canvas.cpp:859-866)
            dst[0] = make_uint4(0u, 0u, 0u, __float_as_uint(found ? best_t : FLT_MAX));
          } else if (c == 1) {  // raw hit: record slot, barycentrics, mesh index (resolve_kernel finishes it)
            dst[1] = found ? make_uint4(best_slot, __float_as_uint(best_u), __float_as_uint(best_v), best_mesh) : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
          }
        } else if (MODE == SHADOW) {
          if (found && c == 0) {
            uint8_t* mark = reinterpret_cast<uint8_t*>(p.out + __ldg(p.shadow_pix + ray_id));
            *mark = *mark | 1u;  // canvas.cpp:856
          }
        } else {
          if (c == 0) {
            float* h = p.hits + 4 * (size_t)ray_id;
            h[0] = found ? best_u : 0.f;
            h[1] = found ? best_v : 0.f;
            h[2] = best_t;
            h[3] = found ? 1.f : 0.f;
            p.ids[ray_id] = found ? __float_as_uint(p.meshes[best_mesh].tris[best_slot].v0.w) : 0xFFFFFFFFu;
          }
        }
        have_ray = false;
      }
    }
    // groups without a ray take the next slots of the warp's pool (warp-uniform loop)
    uint32_t need = __ballot_sync(0xffffffffu, !have_ray) & 0x01010101u;
    while (need && !exhausted) {
      if (pool_next >= POOL) {
        uint32_t id = 0;
        if (lane == 0) id = atomicAdd(p.pool_ctr, 1u);
        pool_id = __shfl_sync(0xffffffffu, id, 0);
        pool_next = 0;
        if (pool_id >= total_pools) { exhausted = true; break; }
      }
      const uint32_t rank = __popc(need & ((1u << gshift) - 1u));  // requesting groups before mine
      const uint32_t slot = pool_next + rank;
      const bool take = !have_ray && slot < POOL;
      pool_next += __popc(need);
      if (take) {
        bool ok;
        float4 seed = make_float4(FLT_MAX, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu));
        uint32_t seed_mesh = 0;
        if (LIST) {
          const uint32_t i = pool_id * POOL + slot;
          ok = i < list_n;
          if (ok) {
            const uint2 e = p.hard_id[i];
            ray_id = e.x;
            seed_mesh = e.y;
            seed = p.hard_best[i];
          }
        } else if (MODE == PRIMARY) {
          const uint32_t sup = pool_id / (SUPER_W * SUPER_H), in = pool_id % (SUPER_W * SUPER_H);
          const uint32_t tx = (sup % supers_x) * SUPER_W + (in % SUPER_W), ty = (sup / supers_x) * SUPER_H + (in / SUPER_W);
          const int x = p.x0 + (int)tx * TILE_W + (int)(slot & (TILE_W - 1));
          const int y = p.y0 + (int)ty * TILE_H + (int)(slot / TILE_W);
          ok = x <= p.x1 && y <= p.y1;
          ray_id = ((uint32_t)y << 16) | (uint32_t)x;
        } else {
          ray_id = pool_id * 32u + slot;
          ok = ray_id < list_n;
        }
        if (ok && p.nm == 0u) {  // empty scene: every ray misses
          if (MODE == PRIMARY) {
            const int x = (int)(ray_id & 0xffffu), y = (int)(ray_id >> 16);
            uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)y * p.stride + x);
            if (c == 0) dst[0] = make_uint4(0u, 0u, 0u, __float_as_uint(FLT_MAX));
            else if (c == 1) dst[1] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
          } else if (MODE == RAYLIST && c == 0) {
            float* h = p.hits + 4 * (size_t)ray_id;
            h[0] = 0.f; h[1] = 0.f; h[2] = FLT_MAX; h[3] = 0.f;
            p.ids[ray_id] = 0xFFFFFFFFu;
          }
          ok = false;
        }
        if (ok) {
          const WorldRay wr = world_ray<MODE>(p, ray_id);
          t_near = wr.t_near; t_far = wr.t_far;
          best_t = seed.x; best_u = seed.y; best_v = seed.z; best_slot = __float_as_uint(seed.w); best_mesh = seed_mesh;
          if (best_slot != 0xFFFFFFFFu) t_far = best_t;  // the lane kernel's best hit so far prunes the restart
          mesh_k = 0;
          enter_mesh(wr, 0);
          have_ray = true;
        }
      }
      need = __ballot_sync(0xffffffffu, !have_ray) & 0x01010101u;
    }
    if (exhausted && !__any_sync(0xffffffffu, have_ray)) break;

    // Phases (B) and (C) are entered by the WHOLE warp whenever any of its four groups needs them, and every
    // collective uses the full mask: the groups stay in lockstep and ballot / shuffle compile to single
    // instructions (a per-group member mask makes the compiler emit a MATCH.ANY + divergent fallback around each).
    // =========================== (B) inner node: lane c tests child c ===========================
    const bool at_node = have_ray && !(cur & J3DG_LEAF_BIT);
    if (__any_sync(0xffffffffu, at_node)) {
      uint4 h0 = make_uint4(0u, 0u, 0u, 0u), h1 = make_uint4(0u, 0u, 0u, 0u);
      uint2 q = make_uint2(0u, 0u);
      uint32_t ref = J3DG_EMPTY_CHILD;
      if (at_node) {
        const char* np = reinterpret_cast<const char*>(nodes + cur);
        h0 = __ldg(reinterpret_cast<const uint4*>(np));          // ox oy oz | nchild      (broadcast)
        h1 = __ldg(reinterpret_cast<const uint4*>(np) + 1);      // sx sy sz | pad         (broadcast)
        q = __ldg(reinterpret_cast<const uint2*>(np + 32) + c);  // my child's quantised box
        ref = __ldg(reinterpret_cast<const uint32_t*>(np + 96) + c);
      }
      const float sx = __uint_as_float(h1.x) * idx, sy = __uint_as_float(h1.y) * idy, sz = __uint_as_float(h1.z) * idz;
      const float bx = (__uint_as_float(h0.x) - ox) * idx;
      const float by = (__uint_as_float(h0.y) - oy) * idy;
      const float bz = (__uint_as_float(h0.z) - oz) * idz;
      float tmin = fmaxf(fmaxf(fmaf(plane(q.x, q.y, sel_nx), sx, bx), fmaf(plane(q.x, q.y, sel_ny), sy, by)), fmaxf(fmaf(plane(q.x, q.y, sel_nz), sz, bz), t_near));
      float tmax = fminf(fminf(fmaf(plane(q.x, q.y, sel_fx), sx, bx), fmaf(plane(q.x, q.y, sel_fy), sy, by)), fminf(fmaf(plane(q.x, q.y, sel_fz), sz, bz), t_far));
      // conservative padding against rounding of the slab arithmetic
      tmin = fmaf(-fabsf(tmin), 2e-6f, tmin);
      tmax = fmaf(fabsf(tmax), 2e-6f, tmax);
      const bool hit = at_node && tmin <= tmax;  // empty slots have inverted boxes and never pass
      const uint32_t hm = (__ballot_sync(0xffffffffu, hit) >> gshift) & 0xFFu;
      const float key = hit ? tmin : FLT_MAX;
      const float nearest = group_min(key);
      const uint32_t nm8 = (__ballot_sync(0xffffffffu, hit && key == nearest) >> gshift) & 0xFFu;
      const int near_lane = __ffs(nm8) - 1;  // -1 when nothing was hit
      const uint32_t next = __shfl_sync(0xffffffffu, ref, gshift + (near_lane & 7));
      if (at_node) {
        if (hm == 0u) {
          cur = pop();
        } else {
          const uint32_t others = hm & ~(1u << near_lane);
          if (hit && c != near_lane) {
            const int pos = sp + __popc(others & below);
            if (pos < STACK_SIZE) stk[pos * GROUPS_PER_BLOCK] = make_uint2(ref, __float_as_uint(tmin));
            else *overflow_flag = 1u;
          }
          sp = min(sp + __popc(others), STACK_SIZE);
          cur = next;
        }
      }
      __syncwarp();  // the pushes must be visible to whichever lane pops them
    }

    // =========================== (C) leaf: lane c tests triangle c ===========================
    const bool at_leaf = have_ray && (cur & J3DG_LEAF_BIT) && cur != J3DG_EMPTY_CHILD;
    if (__any_sync(0xffffffffu, at_leaf)) {
      const uint32_t first = cur & J3DG_LEAF_FIRST_MASK;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
      if (at_leaf) {
        const float4* tp = reinterpret_cast<const float4*>(tris + first + c);
        v0 = __ldg(tp); v1 = __ldg(tp + 1); v2 = __ldg(tp + 2);
      }
      const uint32_t lm = (__ballot_sync(0xffffffffu, __float_as_uint(v1.w) != 0u) >> gshift) & 0xFFu;
      const int last = lm ? __ffs(lm) - 1 : GROUP - 1;
      bool hit = at_leaf && c <= last;
      float t = 0.f, u = 0.f, v = 0.f;
      if (hit) {  // one lane of intersect_woop (qbvh.h:4825-4869); 1/det is correctly rounded instead of rcpps + NR
        const float Ax_ = fsub(v0.x, ox), Ay_ = fsub(v0.y, oy), Az_ = fsub(v0.z, oz);
        const float Bx_ = fsub(v1.x, ox), By_ = fsub(v1.y, oy), Bz_ = fsub(v1.z, oz);
        const float Cx_ = fsub(v2.x, ox), Cy_ = fsub(v2.y, oy), Cz_ = fsub(v2.z, oz);
        const float Akz = pick(Ax_, Ay_, Az_, kz), Bkz = pick(Bx_, By_, Bz_, kz), Ckz = pick(Cx_, Cy_, Cz_, kz);
        const float Ax = fsub(pick(Ax_, Ay_, Az_, kx), fmul(Sx, Akz));
        const float Ay = fsub(pick(Ax_, Ay_, Az_, ky), fmul(Sy, Akz));
        const float Bx = fsub(pick(Bx_, By_, Bz_, kx), fmul(Sx, Bkz));
        const float By = fsub(pick(Bx_, By_, Bz_, ky), fmul(Sy, Bkz));
        const float Cx = fsub(pick(Cx_, Cy_, Cz_, kx), fmul(Sx, Ckz));
        const float Cy = fsub(pick(Cx_, Cy_, Cz_, ky), fmul(Sy, Ckz));
        const float U = fsub(fmul(Cx, By), fmul(Cy, Bx));
        const float V = fsub(fmul(Ax, Cy), fmul(Ay, Cx));
        const float W = fsub(fmul(Bx, Ay), fmul(By, Ax));
        hit = ((U <= 0.f) && (V <= 0.f) && (W <= 0.f)) || ((U >= 0.f) && (V >= 0.f) && (W >= 0.f));
        const float det = fadd(fadd(U, V), W);
        hit = hit && (det != 0.f);
        if (hit) {
          const float inv_det = fdiv(1.f, det);
          const float Az = fmul(Sz, Akz), Bz = fmul(Sz, Bkz), Cz = fmul(Sz, Ckz);
          const float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
          t = fmul(T, inv_det);
          hit = (t_far > t) && (t > t_near);
          u = fmul(V, inv_det);
          v = fmul(W, inv_det);
        }
      }
      const float key = hit ? (GENERAL ? fabsf(t) : t) : FLT_MAX;
      const float nearest = group_min(key);
      const uint32_t wm = (__ballot_sync(0xffffffffu, hit && key == nearest) >> gshift) & 0xFFu;
      const int win = __ffs(wm) - 1;  // lowest slot wins ties, like the sequential strict-less update; -1: no hit
      const float wt = __shfl_sync(0xffffffffu, t, gshift + (win & 7));
      const float wu = __shfl_sync(0xffffffffu, u, gshift + (win & 7));
      const float wv = __shfl_sync(0xffffffffu, v, gshift + (win & 7));
      if (at_leaf) {
        bool done = false;
        if (win >= 0) {
          const bool closer = GENERAL ? (nearest < fabsf(best_t)) : (wt < best_t);
          if (closer) {
            best_t = wt; best_u = wu; best_v = wv; best_slot = first + (uint32_t)win; best_mesh = mesh_k;
            if (!GENERAL || wt > 0.f) t_far = wt; else t_near = wt;
            done = ANY_HIT;
          }
        }
        if (done) { cur = J3DG_EMPTY_CHILD; sp = 0; }
        else cur = pop();
      }
    }
  }
}

// =====================================================================================================
// lane kernel: one ray per lane
// =====================================================================================================

// Per-thread traversal stack: the first LANE_SM_STACK entries in shared memory (column tid of a
// [LANE_SM_STACK][BLOCK_THREADS] array: conflict-free), deeper entries in local memory.
struct LaneStack {
  uint2* sm;  // &s_stack[threadIdx.x]
  uint2 deep[LANE_STACK - LANE_SM_STACK];
  int sp;
  __device__ __forceinline__ void push(uint32_t ref, float t, uint32_t* overflow_flag) {
    const uint2 e = make_uint2(ref, __float_as_uint(t));
    if (sp < LANE_SM_STACK) sm[sp * BLOCK_THREADS] = e;
    else if (sp < LANE_STACK) deep[sp - LANE_SM_STACK] = e;
    else { *overflow_flag = 1u; return; }
    ++sp;
  }
  __device__ __forceinline__ uint2 pop() {
    --sp;
    return sp < LANE_SM_STACK ? sm[sp * BLOCK_THREADS] : deep[sp - LANE_SM_STACK];
  }
};

struct LaneBest {
  float t, u, v;
  uint32_t slot, mesh;
};

// Traverses one mesh with one lane.  Returns false when the ray ran out of budget (the caller evicts it).
// Control flow is "while-while": all lanes of the warp first descend inner nodes until each holds a leaf
// (or is done), then all test their leaves — the two phases reconverge separately.
template <bool ANY_HIT, bool STATS>
__device__ __forceinline__ bool lane_traverse(const MeshDev& m, uint32_t mesh_index, float ox, float oy, float oz, float dx, float dy, float dz,
                                              float t_near, float& t_far, LaneBest& best, uint32_t& visits, uint32_t budget,
                                              uint32_t& stat_tris, uint32_t* overflow_flag, LaneStack& stk) {
  if (m.nt == 0) return true;
  // intersect_woop_precompute, qbvh.h:4793-4823
  int kz = 2;
  {
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    if (ax > ay) { if (ax > az) kz = 0; }
    else { if (ay > az) kz = 1; }
  }
  int kx = kz == 2 ? 0 : kz + 1;
  int ky = kx == 2 ? 0 : kx + 1;
  const float dkz = pick(dx, dy, dz, kz);
  if (dkz < 0.f) { const int t = kx; kx = ky; ky = t; }
  const float Sz = fdiv(1.f, dkz);
  const float Sx = fmul(pick(dx, dy, dz, kx), Sz);
  const float Sy = fmul(pick(dx, dy, dz, ky), Sz);
  const float idx = safe_rcp(dx), idy = safe_rcp(dy), idz = safe_rcp(dz);
  const uint32_t sel_nx = 0x7660u | (idx < 0.f ? 3u : 0u), sel_fx = 0x7660u | (idx < 0.f ? 0u : 3u);
  const uint32_t sel_ny = 0x7660u | (idy < 0.f ? 4u : 1u), sel_fy = 0x7660u | (idy < 0.f ? 1u : 4u);
  const uint32_t sel_nz = 0x7660u | (idz < 0.f ? 5u : 2u), sel_fz = 0x7660u | (idz < 0.f ? 2u : 5u);

  stk.sp = 0;
  uint32_t cur = 0;  // root node; J3DG_EMPTY_CHILD (which has the leaf bit set) = nothing left
  const WideNode* __restrict__ nodes = m.nodes;
  const TriRec* __restrict__ tris = m.tris;

  auto pop = [&]() -> uint32_t {
    while (stk.sp > 0) {
      const uint2 e = stk.pop();
      // entry points of popped boxes that now lie beyond the shrunk interval are skipped
      if (__uint_as_float(e.y) <= t_far) return e.x;
    }
    return J3DG_EMPTY_CHILD;
  };

  for (;;) {
    // ---- phase 1: inner nodes, 8 quantised child boxes each ----
    while (!(cur & J3DG_LEAF_BIT)) {
      if (visits >= budget) return false;
      ++visits;
      const uint4* np = reinterpret_cast<const uint4*>(nodes + cur);
      const uint4 h0 = __ldg(np + 0);  // ox oy oz | nchild
      const uint4 h1 = __ldg(np + 1);  // sx sy sz | pad
      const uint4 b0 = __ldg(np + 2);  // boxes of children 0, 1
      const uint4 b1 = __ldg(np + 3);
      const uint4 b2 = __ldg(np + 4);
      const uint4 b3 = __ldg(np + 5);
      const uint4 c0 = __ldg(np + 6);
      const uint4 c1 = __ldg(np + 7);
      const float sx = __uint_as_float(h1.x) * idx, sy = __uint_as_float(h1.y) * idy, sz = __uint_as_float(h1.z) * idz;
      const float bx = (__uint_as_float(h0.x) - ox) * idx;
      const float by = (__uint_as_float(h0.y) - oy) * idy;
      const float bz = (__uint_as_float(h0.z) - oz) * idz;
      uint32_t near_ref = J3DG_EMPTY_CHILD;
      float near_t = FLT_MAX;
      auto test_child = [&](uint32_t lo, uint32_t hi, uint32_t ref) {
        float tmin = fmaxf(fmaxf(fmaf(plane(lo, hi, sel_nx), sx, bx), fmaf(plane(lo, hi, sel_ny), sy, by)), fmaxf(fmaf(plane(lo, hi, sel_nz), sz, bz), t_near));
        float tmax = fminf(fminf(fmaf(plane(lo, hi, sel_fx), sx, bx), fmaf(plane(lo, hi, sel_fy), sy, by)), fminf(fmaf(plane(lo, hi, sel_fz), sz, bz), t_far));
        // conservative padding against rounding of the slab arithmetic
        tmin = fmaf(-fabsf(tmin), 2e-6f, tmin);
        tmax = fmaf(fabsf(tmax), 2e-6f, tmax);
        if (tmin <= tmax) {  // empty slots have inverted boxes and never pass
          float tt = tmin;
          if (tt < near_t) {  // keep the nearest in registers, push the other one
            const uint32_t r2 = near_ref; const float t2 = near_t;
            near_ref = ref; near_t = tt;
            ref = r2; tt = t2;
          }
          if (ref != J3DG_EMPTY_CHILD) stk.push(ref, tt, overflow_flag);
        }
      };
      test_child(b0.x, b0.y, c0.x);
      test_child(b0.z, b0.w, c0.y);
      test_child(b1.x, b1.y, c0.z);
      test_child(b1.z, b1.w, c0.w);
      test_child(b2.x, b2.y, c1.x);
      test_child(b2.z, b2.w, c1.y);
      test_child(b3.x, b3.y, c1.z);
      test_child(b3.z, b3.w, c1.w);
      cur = (near_ref != J3DG_EMPTY_CHILD) ? near_ref : pop();
    }
    if (cur == J3DG_EMPTY_CHILD) return true;
    // ---- phase 2: leaves, 1..8 consecutive triangle records each, the last one flagged ----
    do {
      uint32_t slot = cur & J3DG_LEAF_FIRST_MASK;
      for (;;) {
        const float4* tp = reinterpret_cast<const float4*>(tris + slot);
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        if (STATS) ++stat_tris;
        // one lane of intersect_woop (qbvh.h:4825-4869); 1/det is correctly rounded instead of rcpps + NR
        const float Ax_ = fsub(v0.x, ox), Ay_ = fsub(v0.y, oy), Az_ = fsub(v0.z, oz);
        const float Bx_ = fsub(v1.x, ox), By_ = fsub(v1.y, oy), Bz_ = fsub(v1.z, oz);
        const float Cx_ = fsub(v2.x, ox), Cy_ = fsub(v2.y, oy), Cz_ = fsub(v2.z, oz);
        const float Akz = pick(Ax_, Ay_, Az_, kz), Bkz = pick(Bx_, By_, Bz_, kz), Ckz = pick(Cx_, Cy_, Cz_, kz);
        const float Ax = fsub(pick(Ax_, Ay_, Az_, kx), fmul(Sx, Akz));
        const float Ay = fsub(pick(Ax_, Ay_, Az_, ky), fmul(Sy, Akz));
        const float Bx = fsub(pick(Bx_, By_, Bz_, kx), fmul(Sx, Bkz));
        const float By = fsub(pick(Bx_, By_, Bz_, ky), fmul(Sy, Bkz));
        const float Cx = fsub(pick(Cx_, Cy_, Cz_, kx), fmul(Sx, Ckz));
        const float Cy = fsub(pick(Cx_, Cy_, Cz_, ky), fmul(Sy, Ckz));
        const float U = fsub(fmul(Cx, By), fmul(Cy, Bx));
        const float V = fsub(fmul(Ax, Cy), fmul(Ay, Cx));
        const float W = fsub(fmul(Bx, Ay), fmul(By, Ax));
        const bool inside = ((U <= 0.f) && (V <= 0.f) && (W <= 0.f)) || ((U >= 0.f) && (V >= 0.f) && (W >= 0.f));
        const float det = fadd(fadd(U, V), W);
        if (inside && det != 0.f) {
          const float inv_det = fdiv(1.f, det);
          const float Az = fmul(Sz, Akz), Bz = fmul(Sz, Bkz), Cz = fmul(Sz, Ckz);
          const float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
          const float t = fmul(T, inv_det);
          if ((t_far > t) && (t > t_near) && (t < best.t)) {
            best.t = t; best.u = fmul(V, inv_det); best.v = fmul(W, inv_det); best.slot = slot; best.mesh = mesh_index;
            t_far = t;
            if (ANY_HIT) return true;
          }
        }
        if (__float_as_uint(v1.w) != 0u) break;  // end of leaf
        ++slot;
      }
      cur = pop();
    } while (cur != J3DG_EMPTY_CHILD && (cur & J3DG_LEAF_BIT));
    if (cur == J3DG_EMPTY_CHILD) return true;
  }
}

// Persistent warps; lane L of a warp owns slot L of the warp's current pool (an 8x4 pixel tile, or 32
// consecutive entries of the shadow-ray list).
template <int MODE, bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS, J3DG_LANE_MIN_BLOCKS) lane_kernel(const TraceParams p) {
  constexpr bool ANY_HIT = MODE == SHADOW;
  __shared__ uint2 s_stack[LANE_SM_STACK * BLOCK_THREADS];
  LaneStack stk;
  stk.sm = s_stack + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t* const overflow_flag = reinterpret_cast<uint32_t*>(p.stats + 2);
  uint32_t total_pools, supers_x = 1, list_n = 0;
  if (MODE == PRIMARY) {
    const uint32_t tiles_x = (uint32_t)(p.x1 - p.x0 + TILE_W) / TILE_W, tiles_y = (uint32_t)(p.y1 - p.y0 + TILE_H) / TILE_H;
    supers_x = (tiles_x + SUPER_W - 1) / SUPER_W;
    total_pools = supers_x * ((tiles_y + SUPER_H - 1) / SUPER_H) * (SUPER_W * SUPER_H);
  } else {
    list_n = (uint32_t)p.stats[5];
    total_pools = (list_n + 31u) / 32u;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 4, (unsigned long long)list_n);  // shadow rays traced
  }
  uint32_t sum_nodes = 0, sum_tris = 0;
  for (;;) {
    uint32_t pool = 0;
    if (lane == 0) pool = atomicAdd(p.pool_ctr, 1u);
    pool = __shfl_sync(0xffffffffu, pool, 0);
    if (pool >= total_pools) break;
    uint32_t ray_id;
    bool ok;
    if (MODE == PRIMARY) {
      const uint32_t sup = pool / (SUPER_W * SUPER_H), in = pool % (SUPER_W * SUPER_H);
      const uint32_t tx = (sup % supers_x) * SUPER_W + (in % SUPER_W), ty = (sup / supers_x) * SUPER_H + (in / SUPER_W);
      const int x = p.x0 + (int)tx * TILE_W + (lane & (TILE_W - 1));
      const int y = p.y0 + (int)ty * TILE_H + (lane / TILE_W);
      ok = x <= p.x1 && y <= p.y1;
      ray_id = ((uint32_t)y << 16) | (uint32_t)x;
    } else {
      ray_id = pool * 32u + (uint32_t)lane;
      ok = ray_id < list_n;
    }
    if (!ok) continue;
    const WorldRay wr = world_ray<MODE>(p, ray_id);
    float t_far = wr.t_far;
    LaneBest best;
    best.t = FLT_MAX; best.u = 0.f; best.v = 0.f; best.slot = 0xFFFFFFFFu; best.mesh = 0;
    uint32_t visits = 0, ntris = 0;
    bool finished = true;
    for (uint32_t k = 0; k < p.nm; ++k) {
      const MeshDev& m = p.meshes[k];
      // qbvh.h:3358-3359: the ray is taken into object space by the inverted object matrix
      const float4 d2 = mat_vec(m.cs_inv, wr.dir);
      const float4 o2 = mat_vec(m.cs_inv, wr.org);
      finished = lane_traverse<ANY_HIT, STATS>(m, k, o2.x, o2.y, o2.z, d2.x, d2.y, d2.z, wr.t_near, t_far, best, visits, p.budget, ntris,
                                               overflow_flag, stk);
      if (!finished || (ANY_HIT && best.slot != 0xFFFFFFFFu)) break;
    }
    if (STATS) { sum_nodes += visits; sum_tris += ntris; }
    const bool found = best.slot != 0xFFFFFFFFu;
    if (!finished) {  // out of budget: the group kernel finishes this ray, starting from the best hit so far
      const uint32_t i = atomicAdd(p.hard_count, 1u);
      p.hard_id[i] = make_uint2(ray_id, best.mesh);
      p.hard_best[i] = make_float4(best.t, best.u, best.v, __uint_as_float(best.slot));
    } else if (MODE == PRIMARY) {
      const int x = (int)(ray_id & 0xffffu), y = (int)(ray_id >> 16);
      uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)y * p.stride + x);
      uint4 lo = make_uint4(0u, 0u, 0u, __float_as_uint(found ? best.t : FLT_MAX));  // misses already carry the final record (canvas.cpp:859-866)
      if (STATS) { lo.y = visits; lo.z = ntris; }  // the counting pass returns per-pixel costs in the u / v slots
      dst[0] = lo;
      // raw hit: record slot, barycentrics, mesh index (resolve_kernel finishes it)
      dst[1] = found ? make_uint4(best.slot, __float_as_uint(best.u), __float_as_uint(best.v), best.mesh) : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
    } else if (found) {
      uint8_t* mark = reinterpret_cast<uint8_t*>(p.out + __ldg(p.shadow_pix + ray_id));
      *mark = *mark | 1u;  // canvas.cpp:856
    }
  }
  if (STATS) {
    for (int o = 16; o; o >>= 1) {
      sum_nodes += __shfl_xor_sync(0xffffffffu, sum_nodes, o);
      sum_tris += __shfl_xor_sync(0xffffffffu, sum_tris, o);
    }
    if (lane == 0) {
      atomicAdd(p.stats + 0, (unsigned long long)sum_nodes);
      atomicAdd(p.stats + 1, (unsigned long long)sum_tris);
    }
  }
}

// ---- hit -> pixel record (canvas.cpp:788-834) + shadow ray generation (836-854) --------------------------
// One thread per pixel, one warp per 8x4 tile (the order the trace kernels use).  Misses already hold
// their final record.  Shadow rays of hit pixels are appended to a list with one atomic per warp.
__global__ void __launch_bounds__(256) resolve_kernel(const MeshDev* __restrict__ meshes, ViewDev vw, int x0, int y0, int x1, int y1,
                                                       j3dg_pixel* __restrict__ out, uint32_t stride, float4* __restrict__ shadow_pos,
                                                       uint32_t* __restrict__ shadow_pix, unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const uint32_t tiles_x = (uint32_t)(x1 - x0 + TILE_W) / TILE_W;
  const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int x = x0 + (int)(tile % tiles_x) * TILE_W + (lane & (TILE_W - 1));
  const int y = y0 + (int)(tile / tiles_x) * TILE_H + (lane / TILE_W);
  bool shadow_ray = false;
  float4 pos = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x <= x1 && y <= y1) {
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)y * stride + x);
    const uint4 raw = dst[1];
    if (raw.x != 0xFFFFFFFFu) {
      const float t = __uint_as_float(dst[0].w);
      const MeshDev& m = meshes[raw.w];
      const float4* tp = reinterpret_cast<const float4*>(m.tris + raw.x);
      const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
      const uint32_t tri = __float_as_uint(v0.w);
      // compute_triangle_normals (geometry.h:2561-2578; vec.h:477-487, 531-536)
      const float lx = fsub(v1.x, v0.x), ly = fsub(v1.y, v0.y), lz = fsub(v1.z, v0.z);
      const float rx = fsub(v2.x, v0.x), ry = fsub(v2.y, v0.y), rz = fsub(v2.z, v0.z);
      float nx = fsub(fmul(ly, rz), fmul(lz, ry));
      float ny = fsub(fmul(lz, rx), fmul(lx, rz));
      float nz = fsub(fmul(lx, ry), fmul(ly, rx));
      const float denom = fsqrt(fadd(fadd(fmul(nx, nx), fmul(ny, ny)), fmul(nz, nz)));
      if (denom != 0.f) { nx = fdiv(nx, denom); ny = fdiv(ny, denom); nz = fdiv(nz, denom); }
      // canvas.cpp:790-792
      float4 n = mat_vec(vw.cs_inv, make_float4(nx, ny, nz, 0.f));
      n = mat_vec(m.cs, n);
      uint32_t mark = 0, r = 0, g = 0, b = 0;
      const float bu = __uint_as_float(raw.y), bv = __uint_as_float(raw.z);
      const float k = fsub(fsub(1.f, bu), bv);
      if ((vw.flags & J3DG_TEXTURED) && m.uv != nullptr && m.texture != nullptr) {  // canvas.cpp:803-820
        const float* uvc = m.uv + 6 * (size_t)tri;
        float cx = fadd(fadd(fmul(k, uvc[0]), fmul(bu, uvc[2])), fmul(bv, uvc[4]));
        float cy = fadd(fadd(fmul(k, uvc[1]), fmul(bu, uvc[3])), fmul(bv, uvc[5]));
        cx = fmaxf(fminf(cx, 1.f), 0.f);
        cy = fmaxf(fminf(cy, 1.f), 0.f);
        const int tw = (int)m.tex_w, th = (int)m.tex_h;
        int X = __float2int_rz(fmul(cx, (float)tw)), Y = __float2int_rz(fmul(cy, (float)th));
        X = X < 0 ? 0 : X >= tw ? tw - 1 : X;
        Y = Y < 0 ? 0 : Y >= th ? th - 1 : Y;
        const uint32_t color = m.texture[(size_t)Y * m.tex_stride + X];
        r = color & 0xffu; g = (color >> 8) & 0xffu; b = (color >> 16) & 0xffu;
        mark |= 2u;
      } else if ((vw.flags & J3DG_VERTEXCOLORS) && m.vertex_colors != nullptr) {  // canvas.cpp:821-834
        const uint32_t* id = m.indices + 3 * (size_t)tri;
        const float* c0 = m.vertex_colors + 3 * (size_t)id[0];
        const float* c1 = m.vertex_colors + 3 * (size_t)id[1];
        const float* c2 = m.vertex_colors + 3 * (size_t)id[2];
        const float cr = fadd(fadd(fmul(c0[0], k), fmul(bu, c1[0])), fmul(bv, c2[0]));
        const float cg = fadd(fadd(fmul(c0[1], k), fmul(bu, c1[1])), fmul(bv, c2[1]));
        const float cb = fadd(fadd(fmul(c0[2], k), fmul(bu, c1[2])), fmul(bv, c2[2]));
        r = (uint32_t)__float2int_rz(fmul(cr, 255.f)) & 0xffu;
        g = (uint32_t)__float2int_rz(fmul(cg, 255.f)) & 0xffu;
        b = (uint32_t)__float2int_rz(fmul(cb, 255.f)) & 0xffu;
        mark |= 2u;
      }
      if (vw.flags & J3DG_SHADOW) {  // canvas.cpp:836-848
        const float4 V0 = transform_point(m.cs, make_float4(v0.x, v0.y, v0.z, 1.f));
        const float4 V1 = transform_point(m.cs, make_float4(v1.x, v1.y, v1.z, 1.f));
        const float4 V2 = transform_point(m.cs, make_float4(v2.x, v2.y, v2.z, 1.f));
        pos.x = fadd(fadd(fmul(V0.x, k), fmul(bu, V1.x)), fmul(bv, V2.x));
        pos.y = fadd(fadd(fmul(V0.y, k), fmul(bu, V1.y)), fmul(bv, V2.y));
        pos.z = fadd(fadd(fmul(V0.z, k), fmul(bu, V1.z)), fmul(bv, V2.z));
        pos.w = fadd(fadd(fmul(V0.w, k), fmul(bu, V1.w)), fmul(bv, V2.w));
        shadow_ray = true;
      }
      dst[0] = make_uint4(mark | (r << 8) | (g << 16) | (b << 24), __float_as_uint(n.x), __float_as_uint(n.y), __float_as_uint(t));
      dst[1] = make_uint4(tri, raw.y, raw.z, m.db_id);
    }
  }
  if (vw.flags & J3DG_SHADOW) {  // warp-uniform
    const uint32_t votes = __ballot_sync(0xffffffffu, shadow_ray);
    if (votes) {
      const int leader = __ffs(votes) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(stats + 5, (unsigned long long)__popc(votes));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (shadow_ray) {
        const uint32_t i = (uint32_t)base + __popc(votes & ((1u << lane) - 1u));
        shadow_pos[i] = pos;
        shadow_pix[i] = (uint32_t)((size_t)y * stride + x);
      }
    }
  }
}

void fill_mesh_dev(const j3dg_mesh* m, MeshDev& d) {
  d.nodes = m->d_nodes;
  d.tris = m->d_tris;
  d.indices = m->d_indices;
  d.vertices = m->d_vertices;
  d.vertex_colors = m->d_vcolors;
  d.uv = m->d_uv;
  d.texture = m->d_texture;
  d.tex_w = m->tex_w; d.tex_h = m->tex_h; d.tex_stride = m->tex_w;
  d.nt = m->nt;
  d.db_id = m->db_id;
  memcpy(d.cs, m->cs, sizeof(d.cs));
  memcpy(d.cs_inv, m->cs_inv, sizeof(d.cs_inv));
  for (int j = 0; j < 3; ++j) { d.root_min[j] = m->info.bbox_min[j]; d.root_max[j] = m->info.bbox_max[j]; }
}

// host copies of the reference's float helpers (single-rounded operations; x86-64 baseline has no FMA)
void host_mat_vec(const float* m, const float* v, float* out) {
  for (int r = 0; r < 4; ++r) {
    volatile float a = m[r] * v[0], b = m[4 + r] * v[1], c = m[8 + r] * v[2], d = m[12 + r] * v[3];
    volatile float s = a + b;
    s = s + c;
    s = s + d;
    out[r] = s;
  }
}

// persistent grid: fill every SM once, never more blocks than there is work
template <class K>
int persistent_grid(j3dg_ctx* ctx, K kernel, long long pools, int* grid) {
  int nb = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, BLOCK_THREADS, 0);
  if (e != cudaSuccess) return j3dg_cuda_fail(ctx, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor", __FILE__, __LINE__);
  const int warps_per_block = BLOCK_THREADS / 32;
  *grid = (int)std::max<long long>(1, std::min<long long>((long long)ctx->sm_count * std::max(nb, 1), (pools + warps_per_block - 1) / warps_per_block));
  return J3DG_OK;
}

}  // namespace

void j3dg_make_view_dev(const j3dg_view* v, ViewDev& d) {
  d.width = v->width; d.height = v->height;
  d.near_plane = v->near_plane; d.diagonal = v->diagonal;
  memcpy(d.pinv, v->projection_inv, 64);
  memcpy(d.cs, v->cs, 64);
  memcpy(d.cs_inv, v->cs_inv, 64);
  const float o4[4] = {0.f, 0.f, 0.f, 1.f};
  host_mat_vec(v->cs, o4, d.origin);  // canvas.cpp:700-701
  volatile float d3 = v->diagonal * 3.f;  // canvas.cpp:703-705
  volatile float l0 = v->pivot[0] + d3, l1 = v->pivot[1] + d3, l2 = v->pivot[2] + d3;
  const float l4[4] = {l0, l1, l2, 1.f};
  host_mat_vec(v->cs, l4, d.light);
  d.flags = v->flags;
}

int j3dg_launch_cast(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, const j3dg_view* view,
                     int x0, int y0, int x1, int y1, j3dg_pixel* d_pixels, uint32_t stride, bool stats) {
  const int w = (int)view->width, h = (int)view->height;
  if (w <= 0 || h <= 0) return J3DG_OK;
  if (w > 65535 || h > 65535) { j3dg_set_error(ctx, "j3dg_cast: canvas larger than 65535 pixels on a side"); return J3DG_EINVAL; }
  // canvas.cpp:682-698
  x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::max(x1, 0); y1 = std::max(y1, 0);
  if (x0 >= w) x0 = w - 1; if (y0 >= h) y0 = h - 1; if (x1 >= w) x1 = w - 1; if (y1 >= h) y1 = h - 1;
  if (x1 < x0 || y1 < y0) return J3DG_OK;
  std::vector<MeshDev> host(nm ? nm : 1);
  uint32_t used = 0;
  for (uint32_t i = 0; i < nm; ++i) {
    if (!meshes[i]) { j3dg_set_error(ctx, "j3dg_cast: null mesh"); return J3DG_EINVAL; }
    if (!meshes[i]->d_nodes || meshes[i]->nt == 0) continue;  // canvas.cpp:730-731: objects without a BVH are skipped
    fill_mesh_dev(meshes[i], host[used++]);
  }
  {
    void* p = ctx->d_meshes;
    int rc = j3dg_reserve(ctx, &p, &ctx->meshes_cap, sizeof(MeshDev) * std::max<uint32_t>(used, 1));
    ctx->d_meshes = (MeshDev*)p;
    if (rc != J3DG_OK) return rc;
  }
  host.resize(used);
  if (used && (ctx->meshes_uploaded.size() != used || memcmp(ctx->meshes_uploaded.data(), host.data(), sizeof(MeshDev) * used) != 0)) {
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_meshes, host.data(), sizeof(MeshDev) * used, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // pageable source
    ctx->meshes_uploaded = host;
  }
  const bool shadows = (view->flags & J3DG_SHADOW) && used && !stats;
  const int rw = x1 - x0 + 1, rh = y1 - y0 + 1;
  const size_t npx = (size_t)rw * rh;
  const size_t shadow_pix_off = (npx * sizeof(float4) + 255) & ~(size_t)255;
  if (shadows) {  // one shadow ray per hit pixel at most
    int rc = j3dg_reserve(ctx, &ctx->d_shadow, &ctx->shadow_cap, shadow_pix_off + npx * sizeof(uint32_t));
    if (rc != J3DG_OK) return rc;
  }
  const size_t hard_id_off = (npx * sizeof(float4) + 255) & ~(size_t)255;
  {  // hard-ray list: worst case every ray is evicted
    int rc = j3dg_reserve(ctx, &ctx->d_hard, &ctx->hard_cap, hard_id_off + npx * sizeof(uint2));
    if (rc != J3DG_OK) return rc;
  }
  // stats slots (u64 each): [0] node visits [1] triangle tests [2] stack overflow flag [3] pool counter lane<PRIMARY>
  // [4] shadow rays traced (accumulates until the timings are reset) [5] shadow list length [6] hard rays PRIMARY
  // [7] pool counter group<PRIMARY> [8] pool counter lane<SHADOW> [9] hard rays SHADOW [10] pool counter group<SHADOW>
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 5, 0, 6 * sizeof(unsigned long long), ctx->stream));
  TraceParams tp = {};
  tp.meshes = ctx->d_meshes;
  tp.nm = used;
  j3dg_make_view_dev(view, tp.vw);
  if (!shadows) tp.vw.flags &= ~J3DG_SHADOW;
  tp.x0 = x0; tp.y0 = y0; tp.x1 = x1; tp.y1 = y1;
  tp.out = d_pixels;
  tp.stride = stride;
  tp.stats = ctx->d_stats;
  tp.shadow_pos = (const float4*)ctx->d_shadow;
  tp.shadow_pix = (const uint32_t*)((const char*)ctx->d_shadow + shadow_pix_off);
  tp.hard_best = (float4*)ctx->d_hard;
  tp.hard_id = (uint2*)((char*)ctx->d_hard + hard_id_off);
  tp.budget = stats ? 0xFFFFFFFFu : ctx->lane_budget;
  auto ctr = [&](int slot) { return reinterpret_cast<unsigned int*>(ctx->d_stats + slot); };
  const long long ntiles = (long long)((rw + TILE_W - 1) / TILE_W) * ((rh + TILE_H - 1) / TILE_H);
  int grid = 1, rc;
  rc = j3dg_stage_begin(ctx, 0);
  if (rc != J3DG_OK) return rc;
  // ---- primary rays ----
  if (stats) {
    tp.pool_ctr = ctr(3); tp.hard_count = ctr(6);
    if ((rc = persistent_grid(ctx, lane_kernel<PRIMARY, true>, ntiles, &grid)) != J3DG_OK) return rc;
    lane_kernel<PRIMARY, true><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
  } else if (ctx->cast_algo == 1) {  // group kernel only (A/B testing)
    tp.pool_ctr = ctr(7);
    if ((rc = persistent_grid(ctx, group_kernel<PRIMARY, false>, ntiles, &grid)) != J3DG_OK) return rc;
    group_kernel<PRIMARY, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
  } else {
    tp.pool_ctr = ctr(3); tp.hard_count = ctr(6);
    if ((rc = persistent_grid(ctx, lane_kernel<PRIMARY, false>, ntiles, &grid)) != J3DG_OK) return rc;
    lane_kernel<PRIMARY, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
    tp.pool_ctr = ctr(7);
    if ((rc = persistent_grid(ctx, group_kernel<PRIMARY, true>, ((long long)npx + 3) / 4, &grid)) != J3DG_OK) return rc;
    group_kernel<PRIMARY, true><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
  }
  if (used && !stats) {
    const uint32_t warps = 256 / 32;
    resolve_kernel<<<(uint32_t)((ntiles + warps - 1) / warps), 256, 0, ctx->stream>>>(ctx->d_meshes, tp.vw, x0, y0, x1, y1, d_pixels, stride,
                                                                                     (float4*)tp.shadow_pos, (uint32_t*)tp.shadow_pix, ctx->d_stats);
    KERNEL_CHECK(ctx);
  }
  // ---- shadow rays of the hit pixels ----
  if (shadows) {
    const long long pools = ((long long)npx + 31) / 32;
    if (ctx->cast_algo == 1) {
      tp.pool_ctr = ctr(10);
      if ((rc = persistent_grid(ctx, group_kernel<SHADOW, false>, pools, &grid)) != J3DG_OK) return rc;
      group_kernel<SHADOW, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
      KERNEL_CHECK(ctx);
    } else {
      tp.pool_ctr = ctr(8); tp.hard_count = ctr(9);
      if ((rc = persistent_grid(ctx, lane_kernel<SHADOW, false>, pools, &grid)) != J3DG_OK) return rc;
      lane_kernel<SHADOW, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
      KERNEL_CHECK(ctx);
      tp.pool_ctr = ctr(10);
      if ((rc = persistent_grid(ctx, group_kernel<SHADOW, true>, ((long long)npx + 3) / 4, &grid)) != J3DG_OK) return rc;
      group_kernel<SHADOW, true><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
      KERNEL_CHECK(ctx);
    }
  }
  rc = j3dg_stage_end(ctx, 0);
  if (rc != J3DG_OK) return rc;
  ctx->rays_primary += (uint64_t)rw * rh;  // shadow rays (d_stats[4]) are added when the timings are read
  return J3DG_OK;
}

int j3dg_launch_find_closest(j3dg_mesh* m, const float* d_rays, uint32_t n, float* d_hits, uint32_t* d_ids) {
  j3dg_ctx* ctx = m->ctx;
  if (!n) return J3DG_OK;
  MeshDev d;
  fill_mesh_dev(m, d);
  {
    void* p = ctx->d_meshes;
    int rc = j3dg_reserve(ctx, &p, &ctx->meshes_cap, sizeof(MeshDev));
    ctx->d_meshes = (MeshDev*)p;
    if (rc != J3DG_OK) return rc;
  }
  CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_meshes, &d, sizeof(MeshDev), cudaMemcpyHostToDevice, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // `d` lives on this stack frame
  ctx->meshes_uploaded.clear();
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  TraceParams tp = {};
  tp.meshes = ctx->d_meshes;
  tp.nm = (d.nodes && d.nt) ? 1u : 0u;
  tp.stats = ctx->d_stats;
  tp.rays = d_rays; tp.hits = d_hits; tp.ids = d_ids; tp.nrays = n;
  tp.pool_ctr = reinterpret_cast<unsigned int*>(ctx->d_stats + 3);
  int grid = 1;
  int rc = persistent_grid(ctx, group_kernel<RAYLIST, false>, ((long long)n + 31) / 32, &grid);
  if (rc != J3DG_OK) return rc;
  group_kernel<RAYLIST, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}
