// cast.cu — per-pixel ray cast.  Replaces canvas::update_canvas (j3d/canvas.cpp:677-874):
// ray generation (773-782), qbvh_two_level_with_transformations::find_closest_triangle
// (jtk/qbvh.h:3303-3387) -> qbvh::find_closest_triangle (1701-1852) with the Woop test
// (4793-4869), hit -> pixel record (788-834) and the shadow ray (836-857).
//
// One BVH (8-wide, 128-byte child-major nodes, leaves of <= 8 triangles), two ways to walk it:
//
//  lane mode  — one ray per LANE, one warp per 8x4-pixel tile, persistent warps pulling tiles from a global counter;
//               per step the warp votes between a node step (8 quantised child boxes, nearest child in a register,
//               the others pushed on a short shared-memory stack) and a triangle step (one 48-byte record).  Fewest
//               instructions per ray — the throughput path.  Left alone it finishes 98.7 % of the tiles in the first
//               half of its run time and then waits for a few silhouette tiles whose grazing rays visit 100-250
//               nodes each (profiles/README.md), so every ray gets a BUDGET of node visits; a ray that exceeds it is
//               evicted, with its best hit so far, into a "hard ray" queue.
//  group mode — one ray per 8 LANES: lane c tests child c of the node (or triangle c of the leaf), the group votes,
//               picks the nearest child with three shuffle-min steps and pushes the rest on one stack per ray.  A step
//               is ~8x shorter and rays are claimed one by one, so there is no lockstep tail — the latency path for
//               the hard rays (restarted from the root, pruned by the hit the lane warp already found; carrying the
//               lane's stack along instead was measured and buys nothing) and for the generic find_closest query.
//
// cast_kernel is ONE persistent launch per ray type (a plain launch of at most one machine-full of blocks): every warp first traces tiles in lane mode (producing the
// queue), then — warp by warp, on its own slice of the block's shared-memory stacks, no block barrier — turns into
// four 8-lane groups that drain the queue.  Pipeline of one j3dg_cast:
//   cast_kernel<PRIMARY> (raw hit: t, u, v, record slot) -> resolve_kernel (hit -> pixel record, hit bounding
//   rectangle, shadow ray list) -> cast_kernel<SHADOW> (any hit).
//
// Parity rules: the ray, the Woop edge functions, t/u/v, the triangle normal and
// its two transforms are evaluated with separately rounded mul/add/sub (no FMA), in the
// reference's operation order.  Only the box tests use FMA — they are conservative and do not
// influence which triangle is the closest hit.
#include "common.cuh"
#include "traverse.cuh"

#include <algorithm>
#include <cstring>
#include <cstdio>

namespace {

constexpr int GROUP = 8;                                  // group kernel: lanes per ray = children per node = max triangles per leaf
constexpr int BLOCK_THREADS = 128;
#ifndef J3DG_LANE_MIN_BLOCKS
// 7 blocks of 4 warps per SM = 72 registers per thread: the node step (32 registers of node data on top of the ray) fits
// without spills.  At 8 blocks (64 registers) ptxas parks 6 registers in local memory around EVERY node step as soon as
// anything is added to the kernel (measured on config B: 1.02 instead of 0.97 ms per frame); 6 blocks lose more in
// occupancy than the registers give (1.05 ms).
#define J3DG_LANE_MIN_BLOCKS 7
#endif
#ifndef J3DG_LANE_REFILL_MIN
#define J3DG_LANE_REFILL_MIN 8                            // lane kernel: idle lanes that trigger a refill from the pool (2 / 4 / 8 / 12 / 16: 0.789 / 0.771 / 0.762 / 0.776 / 0.764 ms per frame, three frames in flight)
#endif
constexpr int GROUPS_PER_BLOCK = BLOCK_THREADS / GROUP;   // 16 rays in flight per block
constexpr int STACK_SIZE = 96;                            // entries per ray; 96 * 8 B * 16 = 12 KB shared memory per block
constexpr int TILE_W = 8, TILE_H = 4;                     // a warp's ray pool = one 8x4 pixel tile
constexpr int SUPER_W = 4, SUPER_H = 8;                   // tiles are numbered super-tile by super-tile (32x32 pixels)
#ifndef J3DG_CAST_MIN_BLOCKS
#define J3DG_CAST_MIN_BLOCKS 8
#endif

enum Mode { PRIMARY = 0, SHADOW = 1, RAYLIST = 2 };

// Enumeration of the 8x4-pixel tiles of a rectangle as "pools" (one pool = one tile = 32 rays).  Tiles are
// numbered super-tile by super-tile (32x32 pixels), super-tile rows = BANDS of J3DG_SHARD_BAND_ROWS rows.
// Screen sharding (j3dg_ctx_set_screen_shard): band b belongs to rank b mod world, so the pools
// of a rank are the tiles of its own bands, followed by HALO pools: the tile row directly above each owned
// band, of which only the bottom pixel row is traced — the edge shader reads the up neighbour
// (canvas.cpp:625-640), the right neighbour lies in the same band.  world == 1: every band, no halo.
struct TileGrid {
  uint32_t supers_x;      // super-tiles per band
  uint32_t main_pools;    // own_bands * supers_x * 32
  uint32_t total_pools;   // main_pools + halo pools
  uint32_t rank, world;
};
static_assert(SUPER_H * TILE_H == J3DG_SHARD_BAND_ROWS, "a band is one super-tile row");

TileGrid make_tile_grid(int x0, int y0, int x1, int y1, uint32_t rank, uint32_t world) {
  TileGrid g;
  const uint32_t tiles_x = (uint32_t)(x1 - x0 + TILE_W) / TILE_W, tiles_y = (uint32_t)(y1 - y0 + TILE_H) / TILE_H;
  g.supers_x = (tiles_x + SUPER_W - 1) / SUPER_W;
  const uint32_t bands = (tiles_y + SUPER_H - 1) / SUPER_H;
  g.rank = rank; g.world = world ? world : 1u;
  const uint32_t own = rank < bands ? (bands - rank + g.world - 1) / g.world : 0u;
  g.main_pools = own * g.supers_x * (SUPER_W * SUPER_H);
  g.total_pools = g.main_pools + (g.world > 1 ? own * g.supers_x * SUPER_W : 0u);
  return g;
}

// pool -> tile coordinates; false: nothing to trace.  `halo`: only the tile's bottom pixel row is wanted.
__device__ __forceinline__ bool tile_of_pool(const TileGrid& g, uint32_t pool, uint32_t& tx, uint32_t& ty, bool& halo) {
  if (pool < g.main_pools) {
    const uint32_t sup = pool / (SUPER_W * SUPER_H), in = pool % (SUPER_W * SUPER_H);
    const uint32_t band = g.rank + (sup / g.supers_x) * g.world;
    tx = (sup % g.supers_x) * SUPER_W + (in % SUPER_W);
    ty = band * SUPER_H + (in / SUPER_W);
    halo = false;
    return true;
  }
  const uint32_t h = pool - g.main_pools, per = g.supers_x * SUPER_W;
  const uint32_t band = g.rank + (h / per) * g.world;
  tx = h % per;
  ty = band * SUPER_H - 1u;
  halo = true;
  return band != 0u;
}

// Top-level tree over the objects of a scene (qbvh_two_level_with_transformations, jtk/qbvh.h:3251-3301: a QBVH over the
// transformed root boxes, rebuilt every frame, canvas.cpp:762).  8-wide, plain float boxes in WORLD space, built on the
// host whenever the object table changes (build_top_tree below).  child: bit 31 = object leaf (low bits = index into the
// mesh table), otherwise index of a top node; unused slots hold J3DG_EMPTY_CHILD.  Node 0 is the root.
struct __align__(16) TopNode {
  float lo[8][3], hi[8][3];
  uint32_t child[8];
  uint32_t pad[8];
};
static_assert(sizeof(TopNode) == 256, "TopNode is 256 bytes");
constexpr int TOP_DEPTH = 8;                 // 8^8 objects
constexpr uint32_t TOP_NONE = 0xFFFFFFFFu;

struct TraceParams {
  const MeshDev* meshes;
  uint32_t nm;
  const TopNode* top;            // TOP kernels: the tree over the nm objects
  ViewDev vw;
  int x0, y0, x1, y1;
  TileGrid grid;                 // PRIMARY: the pools of this launch
  j3dg_pixel* out;               // PRIMARY: raw hits are written here; SHADOW: mark bit 0 is set here
  uint32_t stride;
  uint2* spill;                  // pool mode: global continuation of the per-slot stacks, [warp][POOL_SPILL][PSLOTS]
  uint32_t* sticky;              // mapped host status words (common.cuh, j3dg_ctx::d_status): [1] = a traversal stack overflowed
  unsigned long long* stats;     // [0] node rounds [1] triangle tests [2] overflow flag [3] pool counter [4] shadow rays (accumulating) [5] shadow list length
  const float4* shadow_pos;      // SHADOW: ray origins (xyzw as the reference computes them)
  const uint32_t* shadow_pix;    // SHADOW: pixel offset (y * stride + x) of each ray
  // hard-ray hand-over between the two kernels
  uint32_t budget;               // lane kernel: node visits a ray may spend before it is evicted
  unsigned int* pool_ctr;        // pool counter of THIS launch
  // hard-ray queue (single launch: lane warps produce, group warps consume)
  unsigned int* hard_count;      // producers: append position
  unsigned int* hard_taken;      // consumers: next entry to claim
  unsigned int* done_blocks;     // lane WARPS that have finished producing
  unsigned int* started;         // lane WARPS that have begun to produce (bumped before a warp's first pool fetch)
  uint32_t hard_capacity;        // entries the queue arrays hold
  uint2* hard_id;                // {ray id (0xFFFFFFFF = entry not written yet), mesh of the best hit so far}
  float4* hard_best;             // {t, u, v, record slot bits} of the best hit so far (slot 0xFFFFFFFF: none)
  const float* rays;             // RAYLIST: n x 8 floats
  float* hits;                   // RAYLIST: n x 4 floats
  uint32_t* ids;                 // RAYLIST: n triangle ids
  uint32_t nrays;
};

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

// Walk of the top-level tree, one per ray: a stack of (node << 8 | children still to visit).  top_next returns the next
// object whose world box the ray meets inside [t_near, t_far] (t_far as of NOW: boxes behind the best hit so far are
// dropped when they come up), or TOP_NONE.  Only called where a ray changes objects, so it lives in local memory and is
// written for size, not speed; object order = tree order (ties between objects may fall differently than in a linear
// loop, as they do in the reference's own top-level traversal).
struct TopWalk {
  uint32_t stk[TOP_DEPTH];
  int sp;
};

__device__ __forceinline__ uint32_t top_box_mask(const TopNode& n, const float o[3], const float inv[3], float t_near, float t_far) {
  uint32_t m = 0;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    if (n.child[c] == J3DG_EMPTY_CHILD) continue;  // unused slot (the min / max slab form would take its inverted box for everything)
    float tmin = t_near, tmax = t_far;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float t0 = (n.lo[c][a] - o[a]) * inv[a], t1 = (n.hi[c][a] - o[a]) * inv[a];
      tmin = fmaxf(tmin, fminf(t0, t1));
      tmax = fminf(tmax, fmaxf(t0, t1));
    }
    tmin = fmaf(-fabsf(tmin), 4e-6f, tmin);  // conservative against the rounding of the slab arithmetic
    tmax = fmaf(fabsf(tmax), 4e-6f, tmax);
    if (tmin <= tmax) m |= 1u << c;
  }
  return m;
}

__device__ __noinline__ uint32_t top_next(const TopNode* __restrict__ top, TopWalk& w, const WorldRay& wr, float t_far, bool restart) {
  const float o[3] = {wr.org.x, wr.org.y, wr.org.z};
  const float inv[3] = {safe_rcp(wr.dir.x), safe_rcp(wr.dir.y), safe_rcp(wr.dir.z)};
  if (restart) {
    w.sp = 1;
    w.stk[0] = top_box_mask(top[0], o, inv, wr.t_near, t_far);  // node 0
  }
  while (w.sp > 0) {
    const uint32_t e = w.stk[w.sp - 1];
    const uint32_t mask = e & 0xFFu, node = e >> 8;
    if (!mask) { --w.sp; continue; }
    const int c = __ffs(mask) - 1;
    w.stk[w.sp - 1] = (node << 8) | (mask & (mask - 1u));
    const TopNode& n = top[node];
    const uint32_t ref = n.child[c];
    if (ref & J3DG_LEAF_BIT) {
      // the object's box again, against the interval as it is now
      float tmin = wr.t_near, tmax = t_far;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float t0 = (n.lo[c][a] - o[a]) * inv[a], t1 = (n.hi[c][a] - o[a]) * inv[a];
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
      }
      tmin = fmaf(-fabsf(tmin), 4e-6f, tmin);
      tmax = fmaf(fabsf(tmax), 4e-6f, tmax);
      if (tmin <= tmax) return ref & ~J3DG_LEAF_BIT;
    } else if (w.sp < TOP_DEPTH) {
      const uint32_t m = top_box_mask(top[ref], o, inv, wr.t_near, t_far);
      if (m) w.stk[w.sp++] = (ref << 8) | m;
    }
  }
  return TOP_NONE;
}

// MODE PRIMARY: closest hit per pixel, raw result into the pixel buffer.
// MODE SHADOW : any hit (first accepted triangle ends the ray), sets mark bit 0.
// MODE RAYLIST: qbvh::find_closest_triangle semantics for arbitrary (also negative) t ranges:
//               closest = smallest |t|, bounds shrink on the side of the hit (qbvh.h:1812-1823).
// SRC POOLS: rays come in pools of 32 slots (a pixel tile / 32 consecutive list entries) from p.pool_ctr.
// SRC QUEUE: rays are the entries of the hard-ray queue the lane warps of the same launch fill; an entry
//            is claimed with an atomic, polled until its ray id is written (the producer writes the seed,
//            fences, then the id; the consumer restores the empty marker), and a group retires when its
//            claimed index lies past the final length after every producer block has signed off.
//            "Every producer has signed off" counts the warps that HAVE STARTED, not the warps of the grid: a warp
//            announces itself (TraceParams::started) before its first pool fetch and signs off after its last ray,
//            and a warp turns consumer only when the pool counter has run out — so whatever a waiting group still
//            waits for is being traced by RUNNING warps.  Blocks of the launch that have not started yet (frames
//            in flight: the machine is shared with the previous frame's kernel) have nothing left to take, announce
//            and sign off in one go when they finally run, and nobody waits for them.  Producers never wait.
//            Hence launches cannot deadlock however many of them share the machine, and no block has to be
//            resident at any particular time.
enum Source { POOLS = 0, QUEUE = 1 };
constexpr uint32_t QUEUE_EMPTY = 0xFFFFFFFFu;
#ifndef J3DG_IDLE_BACKOFF_MAX_NS
#define J3DG_IDLE_BACKOFF_MAX_NS 8000u                    // longest wait of a warp that found nothing to claim
#endif

// `stk` = entry 0 of this group's stack, entry i at stk[i * GSTRIDE] (GSTRIDE groups are interleaved).
template <int MODE, int SRC, int GSTRIDE, bool TOP = false>
__device__ __forceinline__ void group_loop(const TraceParams& p, uint2* const stk) {
  constexpr bool ANY_HIT = MODE == SHADOW;
  constexpr bool GENERAL = MODE == RAYLIST;
  TopWalk walk;  // TOP only (replicated in the 8 lanes of a group)
  walk.sp = 0;
  const int lane = threadIdx.x & 31;
  const int c = lane & 7;                               // my child / triangle slot
  const int gshift = lane & 24;                         // first lane of my group
  const uint32_t below = (1u << c) - 1u;
  uint32_t* const overflow_flag = reinterpret_cast<uint32_t*>(p.stats + 2);

  // ---- number of ray slots ----
  uint32_t total_pools;   // a pool = 32 consecutive slots
  uint32_t list_n = 0;
  if (SRC == QUEUE) {
    total_pools = 0;
  } else if (MODE == PRIMARY) {
    total_pools = p.grid.total_pools;
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

  // ---- warp-uniform pool state (POOLS) / per-group queue state (QUEUE) ----
  uint32_t pool_next = 32u, pool_id = 0;
  bool exhausted = false;
  uint32_t claimed = QUEUE_EMPTY;   // QUEUE: entry this group has claimed and waits for
  bool retired = false;             // QUEUE: nothing left for this group
  uint32_t backoff = 250u;          // QUEUE: nanoseconds an idle warp sleeps before it polls again (doubles up to 4 us)
  uint32_t round_no = 0;            // QUEUE: while other groups of the warp traverse, an idle group polls only every 4th round
  bool warp_busy = false;

  auto pop = [&]() -> uint32_t {
    while (sp > 0) {
      --sp;
      const uint2 e = stk[sp * GSTRIDE];
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
    sel_nx = plane_sel(idx < 0.f ? 3u : 0u); sel_fx = plane_sel(idx < 0.f ? 0u : 3u);
    sel_ny = plane_sel(idy < 0.f ? 4u : 1u); sel_fy = plane_sel(idy < 0.f ? 1u : 4u);
    sel_nz = plane_sel(idz < 0.f ? 5u : 2u); sel_fz = plane_sel(idz < 0.f ? 2u : 5u);
    sp = 0;
    cur = m.nt ? 0u : J3DG_EMPTY_CHILD;
  };

  for (;;) {
    // =========================== (A) finish rays, move to the next mesh, refill ===========================
    if (have_ray && cur == J3DG_EMPTY_CHILD) {
      const bool found = best_slot != 0xFFFFFFFFu;
      bool more = false;
      if (TOP) {  // next object of the top-level walk (qbvh.h:3340-3385)
        if (!(ANY_HIT && found)) {
          const WorldRay wr = world_ray<MODE>(p, ray_id);
          const uint32_t k = top_next(p.top, walk, wr, t_far, false);
          if (k != TOP_NONE) { mesh_k = k; enter_mesh(wr, k); more = true; }
        }
      } else {
        ++mesh_k;
        if (mesh_k < p.nm && !(ANY_HIT && found)) {
          const WorldRay wr = world_ray<MODE>(p, ray_id);
          enter_mesh(wr, mesh_k);
          more = true;
        }
      }
      if (!more) {
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
    if (SRC == QUEUE) {
      // groups without a ray claim the next queue entry and poll it (once per round: never a spin inside a warp
      // whose other groups are traversing)
      ++round_no;
      if (!have_ray && !retired && (!warp_busy || (round_no & 3u) == 0u)) {
        if (claimed == QUEUE_EMPTY) {
          uint32_t i = 0;
          if (c == 0) i = atomicAdd(p.hard_taken, 1u);
          claimed = __shfl_sync(0xFFu << gshift, i, gshift);
        }
        uint2 e = make_uint2(QUEUE_EMPTY, 0u);
        if (claimed < p.hard_capacity)  // claims run past the end of the queue while it drains
          asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "l"(p.hard_id + claimed));
        if (e.x != QUEUE_EMPTY) {
          __threadfence();  // the seed was written before the id
          const float4 seed = __ldcg(p.hard_best + claimed);
          __syncwarp(0xFFu << gshift);
          if (c == 0) p.hard_id[claimed].x = QUEUE_EMPTY;  // leave the queue clean for the next launch
          claimed = QUEUE_EMPTY;
          ray_id = e.x;
          const WorldRay wr = world_ray<MODE>(p, ray_id);
          t_near = wr.t_near; t_far = wr.t_far;
          best_t = seed.x; best_u = seed.y; best_v = seed.z; best_slot = __float_as_uint(seed.w); best_mesh = e.y;
          if (best_slot != 0xFFFFFFFFu) t_far = best_t;  // the lane warp's best hit so far prunes the restart
          mesh_k = 0;
          if (TOP) {
            const uint32_t k = top_next(p.top, walk, wr, t_far, true);
            if (k != TOP_NONE) { mesh_k = k; enter_mesh(wr, k); }
            else { cur = J3DG_EMPTY_CHILD; sp = 0; }
          } else {
            enter_mesh(wr, 0);
          }
          have_ray = true;
        } else {
          // done | started in one load.  A group only gets here after the pool counter has run out, so no warp can take
          // a first pool any more: every warp that ever will produce has bumped `started` (it does so before its first
          // fetch), and once as many have signed off (after fencing their writes) the queue length is final.
          uint2 ds;
          asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(ds.x), "=r"(ds.y) : "l"(p.done_blocks));
          if (ds.x >= ds.y && claimed >= *reinterpret_cast<volatile unsigned int*>(p.hard_count)) retired = true;
        }
      }
      const uint32_t busy = __ballot_sync(0xffffffffu, have_ray);
      warp_busy = busy != 0u;
      if (!busy) {
        if (__all_sync(0xffffffffu, retired)) break;
        // nothing to do yet: do not hammer the queue, nor the issue slots of the working warps.  __nanosleep may return
        // long before its argument has passed (measured: 0.7 us per round with a 4 us request, which made the polling
        // 20 % of all instructions of the kernel), so the wait is held against the global timer.
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
          __nanosleep(backoff);
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (t1 - t0 < (unsigned long long)backoff);
        backoff = min(backoff * 2u, J3DG_IDLE_BACKOFF_MAX_NS);
        continue;
      }
      backoff = 250u;
    } else {
    // groups without a ray take the next slots of the warp's pool (warp-uniform loop)
    uint32_t need = __ballot_sync(0xffffffffu, !have_ray) & 0x01010101u;
    while (need && !exhausted) {
      if (pool_next >= 32u) {
        uint32_t id = 0;
        if (lane == 0) id = atomicAdd(p.pool_ctr, 1u);
        pool_id = __shfl_sync(0xffffffffu, id, 0);
        pool_next = 0;
        if (pool_id >= total_pools) { exhausted = true; break; }
      }
      const uint32_t rank = __popc(need & ((1u << gshift) - 1u));  // requesting groups before mine
      const uint32_t slot = pool_next + rank;
      const bool take = !have_ray && slot < 32u;
      pool_next += __popc(need);
      if (take) {
        bool ok;
        if (MODE == PRIMARY) {
          uint32_t tx, ty;
          bool halo;
          const bool tile_ok = tile_of_pool(p.grid, pool_id, tx, ty, halo);
          const int x = p.x0 + (int)tx * TILE_W + (int)(slot & (TILE_W - 1));
          const int y = p.y0 + (int)ty * TILE_H + (int)(slot / TILE_W);
          ok = tile_ok && x <= p.x1 && y <= p.y1 && (!halo || slot / TILE_W == TILE_H - 1);
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
          best_t = FLT_MAX; best_u = 0.f; best_v = 0.f; best_slot = 0xFFFFFFFFu; best_mesh = 0;
          mesh_k = 0;
          if (TOP) {
            const uint32_t k = top_next(p.top, walk, wr, t_far, true);
            if (k != TOP_NONE) { mesh_k = k; enter_mesh(wr, k); }
            else { cur = J3DG_EMPTY_CHILD; sp = 0; }
          } else {
            enter_mesh(wr, 0);
          }
          have_ray = true;
        }
      }
      need = __ballot_sync(0xffffffffu, !have_ray) & 0x01010101u;
    }
    if (exhausted && !__any_sync(0xffffffffu, have_ray)) break;
    }

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
        const U32x8 hh = ldg256(np);                             // ox oy oz | nchild | sx sy sz | pad   (broadcast)
        h0 = make_uint4(hh.v[0], hh.v[1], hh.v[2], hh.v[3]);
        h1 = make_uint4(hh.v[4], hh.v[5], hh.v[6], hh.v[7]);
        q = __ldg(reinterpret_cast<const uint2*>(np + 32) + c);  // my child's quantised box
        ref = __ldg(reinterpret_cast<const uint32_t*>(np + 96) + c);
      }
      const Slab X = slab(__uint_as_float(h1.x), __uint_as_float(h0.x), ox, idx);
      const Slab Y = slab(__uint_as_float(h1.y), __uint_as_float(h0.y), oy, idy);
      const Slab Z = slab(__uint_as_float(h1.z), __uint_as_float(h0.z), oz, idz);
      float tmin = fmaxf(fmaxf(fmaf(plane(q.x, q.y, sel_nx), X.S, X.Bn), fmaf(plane(q.x, q.y, sel_ny), Y.S, Y.Bn)), fmaxf(fmaf(plane(q.x, q.y, sel_nz), Z.S, Z.Bn), t_near));
      float tmax = fminf(fminf(fmaf(plane(q.x, q.y, sel_fx), X.S, X.Bf), fmaf(plane(q.x, q.y, sel_fy), Y.S, Y.Bf)), fminf(fmaf(plane(q.x, q.y, sel_fz), Z.S, Z.Bf), t_far));
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
            if (pos < STACK_SIZE) stk[pos * GSTRIDE] = make_uint2(ref, __float_as_uint(tmin));
            else { *overflow_flag = 1u; p.sticky[1] = 1u; }
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

template <int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS, J3DG_CAST_MIN_BLOCKS) group_kernel(const TraceParams p) {
  __shared__ uint2 s_stack[STACK_SIZE * GROUPS_PER_BLOCK];
  group_loop<MODE, POOLS, GROUPS_PER_BLOCK>(p, s_stack + (threadIdx.x >> 3));
}

// =====================================================================================================
// lane kernel: one ray per lane, persistent lanes
// =====================================================================================================
// Every lane runs its own ray through a small state machine (at a node / at a leaf / done); the warp
// alternates a node phase and a leaf phase ("while-while") so that lanes decoding a node never serialise
// against lanes testing triangles.  A lane whose ray is done writes its pixel and — as soon as
// LANE_REFILL_MIN lanes are idle — takes the next ray of the warp's pool, so the 32 lanes stay busy instead
// of waiting for the slowest ray of a tile (dynamic ray fetch, Aila & Laine 2009).  A pool is one 8x4
// pixel tile (or 32 shadow rays); when the warp fetches a pool its 32 lanes set up the 32 rays in
// parallel — ray generation, object-space transform, Woop shear constants, reciprocals — and park them
// in shared memory, so a refill costs a few shared-memory loads.
// The traversal stack is LANE_STACK entries per lane, all in shared memory, pushed without branches.
// A ray that spends more than `budget` node visits, or would overflow the stack, is evicted to the
// hard-ray list with its best hit so far (group_kernel finishes it).
constexpr int LANE_STACK = 12;        // 13 rows (one scratch row) * 8 B * 128 lanes = 13 KB shared memory per block
constexpr int LANE_REFILL_MIN = J3DG_LANE_REFILL_MIN;
#ifndef J3DG_LANE_TRI_WEIGHT
#define J3DG_LANE_TRI_WEIGHT 1                            // lane kernel: a triangle step is taken when tri lanes * weight >= node lanes
#endif
constexpr int LANE_TRI_WEIGHT = J3DG_LANE_TRI_WEIGHT;
#ifndef J3DG_TRI_PREFETCH
#define J3DG_TRI_PREFETCH 5                               // lane kernel: L1 prefetch distance of a triangle step in float4 units (0 = off)
#endif
constexpr int LANE_TRI_PREFETCH = J3DG_TRI_PREFETCH;
// every warp owns a contiguous (LANE_STACK + 1) x 32 slice of the stack area: when its tiles run out it turns into
// four 8-lane groups that reuse the same slice ((LANE_STACK + 1) * 32 / 4 = 104 >= STACK_SIZE entries per group),
// so no block-wide barrier separates the two phases
constexpr int LANE_STRIDE = 32;
constexpr int RAY_WORDS = 12;         // parked ray: ox oy oz | idx idy idz | Sx Sy Sz | kx,ky,kz packed | t_far | ray id

// Object-space traversal constants of a ray for one mesh (qbvh.h:3358-3359, 4793-4823).
struct LaneRay {
  float ox, oy, oz, idx, idy, idz, Sx, Sy, Sz;
  uint32_t kpack;  // kx | ky << 2 | kz << 4
};

__device__ __forceinline__ LaneRay lane_ray_setup(const MeshDev& m, const WorldRay& wr) {
  LaneRay r;
  const float4 d2 = mat_vec(m.cs_inv, wr.dir);
  const float4 o2 = mat_vec(m.cs_inv, wr.org);
  r.ox = o2.x; r.oy = o2.y; r.oz = o2.z;
  int kz = 2;
  const float ax = fabsf(d2.x), ay = fabsf(d2.y), az = fabsf(d2.z);
  if (ax > ay) { if (ax > az) kz = 0; }
  else { if (ay > az) kz = 1; }
  int kx = kz == 2 ? 0 : kz + 1;
  int ky = kx == 2 ? 0 : kx + 1;
  const float dkz = pick(d2.x, d2.y, d2.z, kz);
  if (dkz < 0.f) { const int t = kx; kx = ky; ky = t; }
  r.Sz = fdiv(1.f, dkz);
  r.Sx = fmul(pick(d2.x, d2.y, d2.z, kx), r.Sz);
  r.Sy = fmul(pick(d2.x, d2.y, d2.z, ky), r.Sz);
  r.idx = safe_rcp(d2.x); r.idy = safe_rcp(d2.y); r.idz = safe_rcp(d2.z);
  r.kpack = (uint32_t)kx | ((uint32_t)ky << 2) | ((uint32_t)kz << 4);
  return r;
}

constexpr size_t LANE_SMEM_STACK = (size_t)(LANE_STACK + 1) * BLOCK_THREADS * sizeof(uint2);
constexpr size_t LANE_SMEM_RAYS = (size_t)(BLOCK_THREADS / 32) * RAY_WORDS * 32 * sizeof(uint32_t);
constexpr size_t GROUP_SMEM = (size_t)STACK_SIZE * GROUPS_PER_BLOCK * sizeof(uint2);
constexpr size_t CAST_SMEM = LANE_SMEM_STACK + LANE_SMEM_RAYS > GROUP_SMEM ? LANE_SMEM_STACK + LANE_SMEM_RAYS : GROUP_SMEM;

template <int MODE, bool STATS, bool TOP = false>
__device__ __forceinline__ void lane_loop(const TraceParams& p, uint2* const stk, uint32_t* s_rays) {  // stk: row i of this lane at stk[i * LANE_STRIDE]
  constexpr bool ANY_HIT = MODE == SHADOW;
  TopWalk walk;  // TOP only
  walk.sp = 0;
  uint32_t* const park = s_rays + (threadIdx.x >> 5) * (RAY_WORDS * 32);              // word w of slot s at park[w * 32 + s]
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const float t_near = MODE == PRIMARY ? fdiv(p.vw.diagonal, 100.f) : 1e-3f;          // canvas.cpp:781 / 853
  uint32_t total_pools, list_n = 0;
  if (MODE == PRIMARY) {
    total_pools = p.grid.total_pools;
  } else {
    list_n = (uint32_t)p.stats[5];
    total_pools = (list_n + 31u) / 32u;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 4, (unsigned long long)list_n);  // shadow rays traced
  }

  // ---- per-lane ray state ----
  bool have = false;
  uint32_t ray_id = 0, mesh_k = 0, visits = 0, ntris = 0;
  const WideNode* __restrict__ nodes = nullptr;
  const TriRec* __restrict__ tris = nullptr;
  LaneRay r = {};
  uint32_t sel_nx = 0, sel_fx = 0, sel_ny = 0, sel_fy = 0, sel_nz = 0, sel_fz = 0;
  float t_far = 0.f;
  float best_t = FLT_MAX, best_u = 0.f, best_v = 0.f;
  uint32_t best_slot = 0xFFFFFFFFu, best_mesh = 0;
  uint32_t cur = J3DG_EMPTY_CHILD;
  int sp = 0;
  bool evict = false;
  uint32_t sum_nodes = 0, sum_tris = 0;
  // ---- warp-uniform pool state ----
  uint32_t pool_next = 0, pool_count = 0;
  bool exhausted = false;
  if (lane == 0) atomicAdd(p.started, 1u);  // before the first pool fetch (group_loop, SRC QUEUE)

  auto enter = [&](const LaneRay& lr, uint32_t k) {
    const MeshDev& m = p.meshes[k];
    nodes = m.nodes; tris = m.tris;
    r = lr;
    sel_nx = plane_sel(r.idx < 0.f ? 3u : 0u); sel_fx = plane_sel(r.idx < 0.f ? 0u : 3u);
    sel_ny = plane_sel(r.idy < 0.f ? 4u : 1u); sel_fy = plane_sel(r.idy < 0.f ? 1u : 4u);
    sel_nz = plane_sel(r.idz < 0.f ? 5u : 2u); sel_fz = plane_sel(r.idz < 0.f ? 2u : 5u);
    sp = 0;
    cur = m.nt ? 0u : J3DG_EMPTY_CHILD;
  };
  auto pop = [&]() -> uint32_t {
    while (sp > 0) {
      --sp;
      const uint2 e = stk[sp * LANE_STRIDE];
      // entry points of popped boxes that now lie beyond the shrunk interval are skipped
      if (__uint_as_float(e.y & ~7u) <= t_far) return e.x;
    }
    return J3DG_EMPTY_CHILD;
  };

  for (;;) {
    // =========================== (A) rays that are done or evicted ===========================
    if (have && (cur == J3DG_EMPTY_CHILD || evict)) {
      const bool found = best_slot != 0xFFFFFFFFu;
      bool more = false;
      if (!evict && !(ANY_HIT && found)) {  // next object (qbvh.h:3340-3385)
        if (TOP) {
          const WorldRay wr = world_ray<MODE>(p, ray_id);
          const uint32_t k = top_next(p.top, walk, wr, t_far, false);
          if (k != TOP_NONE) { mesh_k = k; enter(lane_ray_setup(p.meshes[k], wr), k); more = true; }
        } else if (mesh_k + 1 < p.nm) {
          ++mesh_k;
          enter(lane_ray_setup(p.meshes[mesh_k], world_ray<MODE>(p, ray_id)), mesh_k);
          more = true;
        }
      }
      if (!more) {
        if (evict) {  // a group of 8 lanes finishes this ray, starting from the best hit so far
          const uint32_t i = atomicAdd(p.hard_count, 1u);
          __stcg(p.hard_best + i, make_float4(best_t, best_u, best_v, __uint_as_float(best_slot)));
          __threadfence();  // the consumer reads the seed after it has seen the id
          asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" :: "l"(p.hard_id + i), "r"(ray_id), "r"(best_mesh) : "memory");
        } else if (MODE == PRIMARY) {
          const int x = (int)(ray_id & 0xffffu), y = (int)(ray_id >> 16);
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)y * p.stride + x);
          uint4 lo = make_uint4(0u, 0u, 0u, __float_as_uint(found ? best_t : FLT_MAX));  // misses already carry the final record (canvas.cpp:859-866)
          if (STATS) { lo.y = visits; lo.z = ntris; }  // the counting pass returns per-pixel costs in the u / v slots
          dst[0] = lo;
          // raw hit: record slot, barycentrics, mesh index (resolve_kernel finishes it)
          dst[1] = found ? make_uint4(best_slot, __float_as_uint(best_u), __float_as_uint(best_v), best_mesh) : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
        } else if (found) {
          uint8_t* mark = reinterpret_cast<uint8_t*>(p.out + __ldg(p.shadow_pix + ray_id));
          *mark = *mark | 1u;  // canvas.cpp:856
        }
        if (STATS) { sum_nodes += visits; sum_tris += ntris; }
        have = false;
      }
    }
    // =========================== (B) refill idle lanes from the warp's pool ===========================
    uint32_t idle = __ballot_sync(0xffffffffu, !have);
    if (idle && !exhausted && ((int)__popc(idle) >= LANE_REFILL_MIN || idle == 0xffffffffu || !__any_sync(0xffffffffu, have && cur != J3DG_EMPTY_CHILD))) {
      while (idle && !exhausted) {
        if (pool_next >= pool_count) {  // fetch the next pool; its 32 rays are set up by the 32 lanes in parallel
          uint32_t pool = 0;
          if (lane == 0) pool = atomicAdd(p.pool_ctr, 1u);
          pool = __shfl_sync(0xffffffffu, pool, 0);
          if (pool >= total_pools) { exhausted = true; break; }
          uint32_t id;
          bool ok;
          if (MODE == PRIMARY) {
            uint32_t tx, ty;
            bool halo;
            const bool tile_ok = tile_of_pool(p.grid, pool, tx, ty, halo);
            const int x = p.x0 + (int)tx * TILE_W + (lane & (TILE_W - 1));
            const int y = p.y0 + (int)ty * TILE_H + (lane / TILE_W);
            ok = tile_ok && x <= p.x1 && y <= p.y1 && (!halo || lane / TILE_W == TILE_H - 1);
            id = ((uint32_t)y << 16) | (uint32_t)x;
          } else {
            id = pool * 32u + (uint32_t)lane;
            ok = id < list_n;
          }
          if (ok && p.nm == 0u) {  // empty scene: every ray misses
            if (MODE == PRIMARY) {
              uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(id >> 16) * p.stride + (id & 0xffffu));
              dst[0] = make_uint4(0u, 0u, 0u, __float_as_uint(FLT_MAX));
              dst[1] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
            }
            ok = false;
          }
          const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
          __syncwarp();  // every lane is past its reads of the previous pool
          if (ok) {
            const WorldRay wr = world_ray<MODE>(p, id);
            const LaneRay lr = lane_ray_setup(p.meshes[0], wr);
            const uint32_t s = __popc(okmask & lt_mask);  // dense slots
            park[0 * 32 + s] = __float_as_uint(lr.ox); park[1 * 32 + s] = __float_as_uint(lr.oy); park[2 * 32 + s] = __float_as_uint(lr.oz);
            park[3 * 32 + s] = __float_as_uint(lr.idx); park[4 * 32 + s] = __float_as_uint(lr.idy); park[5 * 32 + s] = __float_as_uint(lr.idz);
            park[6 * 32 + s] = __float_as_uint(lr.Sx); park[7 * 32 + s] = __float_as_uint(lr.Sy); park[8 * 32 + s] = __float_as_uint(lr.Sz);
            park[9 * 32 + s] = lr.kpack; park[10 * 32 + s] = __float_as_uint(wr.t_far); park[11 * 32 + s] = id;
          }
          __syncwarp();
          pool_next = 0;
          pool_count = __popc(okmask);
          if (pool_count == 0) continue;
        }
        const uint32_t s = pool_next + __popc(idle & lt_mask);  // the r-th idle lane takes the r-th parked ray
        if (!have && s < pool_count) {
          LaneRay lr;
          lr.ox = __uint_as_float(park[0 * 32 + s]); lr.oy = __uint_as_float(park[1 * 32 + s]); lr.oz = __uint_as_float(park[2 * 32 + s]);
          lr.idx = __uint_as_float(park[3 * 32 + s]); lr.idy = __uint_as_float(park[4 * 32 + s]); lr.idz = __uint_as_float(park[5 * 32 + s]);
          lr.Sx = __uint_as_float(park[6 * 32 + s]); lr.Sy = __uint_as_float(park[7 * 32 + s]); lr.Sz = __uint_as_float(park[8 * 32 + s]);
          lr.kpack = park[9 * 32 + s];
          t_far = __uint_as_float(park[10 * 32 + s]);
          ray_id = park[11 * 32 + s];
          best_t = FLT_MAX; best_u = 0.f; best_v = 0.f; best_slot = 0xFFFFFFFFu; best_mesh = 0;
          mesh_k = 0; visits = 0; ntris = 0; evict = false;
          if (TOP) {  // the parked constants are those of object 0; the walk says which object comes first
            const WorldRay wr = world_ray<MODE>(p, ray_id);
            const uint32_t k = top_next(p.top, walk, wr, t_far, true);
            if (k == TOP_NONE) { sp = 0; cur = J3DG_EMPTY_CHILD; r = lr; }
            else { mesh_k = k; enter(k == 0u ? lr : lane_ray_setup(p.meshes[k], wr), k); }
          } else {
            enter(lr, 0);
          }
          have = true;
        }
        pool_next += __popc(idle);
        idle = __ballot_sync(0xffffffffu, !have);
      }
    }
    if (exhausted && !__any_sync(0xffffffffu, have)) break;

    // =========================== (C) steps ===========================
    // Every lane with a ray is either AT A NODE (cur = inner node) or AT A TRIANGLE (cur = leaf bit | record slot).
    // The warp takes one step at a time — a node step (8 quantised child boxes) or a triangle step (ONE record) —
    // and votes which: a step of one kind idles the lanes waiting for the other kind, so the warp takes the
    // kind whose idle lanes cost less (a node step is ~TRI_WEIGHT times longer than a triangle step).  Rays leave
    // a leaf after its flagged last record, so leaves of different sizes do not hold each other up.
    const int kx = (int)(r.kpack & 3u), ky = (int)((r.kpack >> 2) & 3u), kz = (int)(r.kpack >> 4);
    for (;;) {
      const bool at_node = have && !evict && !(cur & J3DG_LEAF_BIT);
      const bool at_tri = have && !evict && (cur & J3DG_LEAF_BIT) && cur != J3DG_EMPTY_CHILD;
      const int nn = __popc(__ballot_sync(0xffffffffu, at_node)), nt = __popc(__ballot_sync(0xffffffffu, at_tri));
      if (nn + nt == 0) break;
      // leave for (A) / (B) when enough lanes ran dry and there is something to refill them with
      if (32 - nn - nt >= LANE_REFILL_MIN && !(exhausted && pool_next >= pool_count)) break;
      if (nn == 0 || nt * LANE_TRI_WEIGHT >= nn) {
        // ---------------- triangle step: one record per lane ----------------
        if (at_tri) {
          const uint32_t slot = cur & J3DG_LEAF_FIRST_MASK;
          const float4* tp = reinterpret_cast<const float4*>(tris + slot);
          const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
          if (LANE_TRI_PREFETCH) asm volatile("prefetch.global.L1 [%0];" :: "l"(tp + LANE_TRI_PREFETCH));  // the line the leaf's next record ends in
          if (STATS) ++ntris;
          // one lane of intersect_woop (qbvh.h:4825-4869); 1/det is correctly rounded instead of rcpps + NR
          const float Ax_ = fsub(v0.x, r.ox), Ay_ = fsub(v0.y, r.oy), Az_ = fsub(v0.z, r.oz);
          const float Bx_ = fsub(v1.x, r.ox), By_ = fsub(v1.y, r.oy), Bz_ = fsub(v1.z, r.oz);
          const float Cx_ = fsub(v2.x, r.ox), Cy_ = fsub(v2.y, r.oy), Cz_ = fsub(v2.z, r.oz);
          const float Akz = pick(Ax_, Ay_, Az_, kz), Bkz = pick(Bx_, By_, Bz_, kz), Ckz = pick(Cx_, Cy_, Cz_, kz);
          const float Ax = fsub(pick(Ax_, Ay_, Az_, kx), fmul(r.Sx, Akz));
          const float Ay = fsub(pick(Ax_, Ay_, Az_, ky), fmul(r.Sy, Akz));
          const float Bx = fsub(pick(Bx_, By_, Bz_, kx), fmul(r.Sx, Bkz));
          const float By = fsub(pick(Bx_, By_, Bz_, ky), fmul(r.Sy, Bkz));
          const float Cx = fsub(pick(Cx_, Cy_, Cz_, kx), fmul(r.Sx, Ckz));
          const float Cy = fsub(pick(Cx_, Cy_, Cz_, ky), fmul(r.Sy, Ckz));
          const float U = fsub(fmul(Cx, By), fmul(Cy, Bx));
          const float V = fsub(fmul(Ax, Cy), fmul(Ay, Cx));
          const float W = fsub(fmul(Bx, Ay), fmul(By, Ax));
          const bool inside = ((U <= 0.f) && (V <= 0.f) && (W <= 0.f)) || ((U >= 0.f) && (V >= 0.f) && (W >= 0.f));
          const float det = fadd(fadd(U, V), W);
          if (inside && det != 0.f) {
            const float inv_det = fdiv(1.f, det);
            const float Az = fmul(r.Sz, Akz), Bz = fmul(r.Sz, Bkz), Cz = fmul(r.Sz, Ckz);
            const float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
            const float t = fmul(T, inv_det);
            if ((t_far > t) && (t > t_near) && (t < best_t)) {
              best_t = t; best_u = fmul(V, inv_det); best_v = fmul(W, inv_det); best_slot = slot; best_mesh = mesh_k;
              t_far = t;
            }
          }
          if (ANY_HIT && best_slot != 0xFFFFFFFFu) { cur = J3DG_EMPTY_CHILD; sp = 0; }
          else if (__float_as_uint(v1.w) != 0u) cur = pop();  // end of leaf
          else cur = cur + 1u;
        }
      } else if (at_node) {
        // ---------------- node step: 8 quantised child boxes ----------------
        if (visits >= p.budget) {
          evict = true;
        } else {
          ++visits;
          const char* np = reinterpret_cast<const char*>(nodes + cur);
          const U32x8 q0 = ldg256(np), q1 = ldg256(np + 32), q2 = ldg256(np + 64), q3 = ldg256(np + 96);
          const uint4 h0 = make_uint4(q0.v[0], q0.v[1], q0.v[2], q0.v[3]);  // ox oy oz | nchild
          const uint4 h1 = make_uint4(q0.v[4], q0.v[5], q0.v[6], q0.v[7]);  // sx sy sz | pad
          const uint4 b0 = make_uint4(q1.v[0], q1.v[1], q1.v[2], q1.v[3]);  // boxes of children 0, 1
          const uint4 b1 = make_uint4(q1.v[4], q1.v[5], q1.v[6], q1.v[7]);
          const uint4 b2 = make_uint4(q2.v[0], q2.v[1], q2.v[2], q2.v[3]);
          const uint4 b3 = make_uint4(q2.v[4], q2.v[5], q2.v[6], q2.v[7]);
          const uint4 c0 = make_uint4(q3.v[0], q3.v[1], q3.v[2], q3.v[3]);
          const uint4 c1 = make_uint4(q3.v[4], q3.v[5], q3.v[6], q3.v[7]);
          const Slab X = slab(__uint_as_float(h1.x), __uint_as_float(h0.x), r.ox, r.idx);
          const Slab Y = slab(__uint_as_float(h1.y), __uint_as_float(h0.y), r.oy, r.idy);
          const Slab Z = slab(__uint_as_float(h1.z), __uint_as_float(h0.z), r.oz, r.idz);
          // Entry distance of child i, low 3 bits replaced by i (distances are positive, so their bit patterns
          // order like integers); missed children get +inf.  Empty slots have inverted boxes and never pass.
          constexpr uint32_t MISS_KEY = 0x7F800000u;
          auto child_key = [&](uint32_t lo, uint32_t hi, uint32_t i) -> uint32_t {
            float tmin = fmaxf(fmaxf(fmaf(plane(lo, hi, sel_nx), X.S, X.Bn), fmaf(plane(lo, hi, sel_ny), Y.S, Y.Bn)), fmaxf(fmaf(plane(lo, hi, sel_nz), Z.S, Z.Bn), t_near));
            float tmax = fminf(fminf(fmaf(plane(lo, hi, sel_fx), X.S, X.Bf), fmaf(plane(lo, hi, sel_fy), Y.S, Y.Bf)), fminf(fmaf(plane(lo, hi, sel_fz), Z.S, Z.Bf), t_far));
            // conservative padding against rounding of the slab arithmetic
            tmin = fmaf(-fabsf(tmin), 2e-6f, tmin);
            tmax = fmaf(fabsf(tmax), 2e-6f, tmax);
            return tmin <= tmax ? ((__float_as_uint(tmin) & ~7u) | i) : (MISS_KEY | i);
          };
          const uint32_t k0 = child_key(b0.x, b0.y, 0u), k1 = child_key(b0.z, b0.w, 1u), k2 = child_key(b1.x, b1.y, 2u), k3 = child_key(b1.z, b1.w, 3u);
          const uint32_t k4 = child_key(b2.x, b2.y, 4u), k5 = child_key(b2.z, b2.w, 5u), k6 = child_key(b3.x, b3.y, 6u), k7 = child_key(b3.z, b3.w, 7u);
          const int nearest = min(min(min((int)k0, (int)k1), min((int)k2, (int)k3)), min(min((int)k4, (int)k5), min((int)k6, (int)k7)));
          const uint32_t ni = (uint32_t)nearest & 7u;
          uint32_t near_ref = c0.x;
          near_ref = ni == 1u ? c0.y : near_ref; near_ref = ni == 2u ? c0.z : near_ref; near_ref = ni == 3u ? c0.w : near_ref;
          near_ref = ni == 4u ? c1.x : near_ref; near_ref = ni == 5u ? c1.y : near_ref; near_ref = ni == 6u ? c1.z : near_ref;
          near_ref = ni == 7u ? c1.w : near_ref;
          if ((uint32_t)nearest >= MISS_KEY) near_ref = J3DG_EMPTY_CHILD;
          // branch-free pushes of the other hit children (row LANE_STACK is scratch; the low key bits are cleared at pop)
          int wanted = sp;
          auto push = [&](uint32_t key, uint32_t ref) {
            stk[sp * LANE_STRIDE] = make_uint2(ref, key);
            const int go = (key < MISS_KEY && (int)key != nearest) ? 1 : 0;
            wanted += go;
            sp = min(sp + go, LANE_STACK);
          };
          push(k0, c0.x); push(k1, c0.y); push(k2, c0.z); push(k3, c0.w);
          push(k4, c1.x); push(k5, c1.y); push(k6, c1.z); push(k7, c1.w);
          evict = wanted > LANE_STACK;  // stack full: the group kernel (96 entries) takes the ray
          cur = (near_ref != J3DG_EMPTY_CHILD) ? near_ref : pop();
        }
      }
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

// =====================================================================================================
// pool mode: the lane kernel with the rays decoupled from the lanes
// =====================================================================================================
// The lane kernel above keeps one ray per lane, so a warp step (node or triangle) only uses the lanes whose ray is in
// the matching state: measured 11.5 of 32 lanes on config B (profiles/README.md, round 2).  Here every warp owns
// PSLOTS = 64 ray SLOTS in shared memory — the whole traversal state of a ray (reciprocal direction, shear constants,
// best hit, current node, short stack) lives in its slot, not in a lane.  Every iteration the warp classifies its slots
// (free / at a node / at a triangle) with two loads and six ballots, picks the kind of step that fills more lanes,
// compacts up to 32 slots of that kind onto the lanes through a small list, and runs the step: ~30 of 32 lanes do
// useful work, and a ray that needs 200 node visits no longer idles 31 lanes — it just keeps its slot.  Free slots are
// refilled from a parked tile of 32 set-up rays, so the slots stay full until the tiles run out.  Rays that exceed the
// node budget or the short stack are still evicted to the hard-ray queue (the 8-lane groups bound the tail of the
// frame).  Same arithmetic, same result as lane_loop; only the scheduling differs.
//
// MEASURED (round 2, config B, profiles/README.md "pool mode"): node steps run with 27.5 instead of 18.9 lanes, but the
// slot bookkeeping (classification, compaction, state in shared memory, 24 instead of 32 warps per SM) costs what the
// fuller steps save: the main phase takes 540 us against 495 us, the frame 1.00 ms against 0.90 ms.  Both designs end
// with the same ~250 us drain of the longest grazing rays (up to 400 dependent steps), which is what bounds the frame.
// The lane kernel therefore stays the default; -DJ3DG_POOL_MODE=1 builds this variant (all parity tests pass on it).
#ifndef J3DG_POOL_MODE
#define J3DG_POOL_MODE 0
#endif
#ifndef J3DG_POOL_STACK
#define J3DG_POOL_STACK 8                                 // stack entries per slot (8 B each); overflow = eviction to the 96-entry group stack
#endif
#ifndef J3DG_POOL_REFILL_MIN
#define J3DG_POOL_REFILL_MIN 8                            // free slots that trigger a refill from the parked tile
#endif
#ifndef J3DG_POOL_NODE_FULL
#define J3DG_POOL_NODE_FULL 32                            // a node step runs whenever this many slots wait at a node; below it the fuller kind of step wins
#endif
#ifndef J3DG_POOL_PREFETCH
#define J3DG_POOL_PREFETCH 1                              // 1: prefetch.global.L1 of the next node / record when a step ends (0 off, 2: L2)
#endif
#ifndef J3DG_POOL_MIN_BLOCKS
#define J3DG_POOL_MIN_BLOCKS 6
#endif
#ifndef J3DG_POOL_SPILL
#define J3DG_POOL_SPILL 56                                // stack entries per slot that continue in global memory above the shared-memory rows
#endif
constexpr int PSLOTS = 64;
constexpr int POOL_SPILL = J3DG_POOL_SPILL;

template <int MODE, bool STATS>
struct PoolLayout {
  static constexpr bool ORG = MODE == SHADOW;               // per-ray origin (primary rays share the camera origin)
  static constexpr int PSTACK = STATS ? 24 : J3DG_POOL_STACK - 1;  // usable entries (+ one scratch row for the branch-free pushes); the counting pass must not lose rays to the (uncounted) group kernel
  // static words of a slot, copied from the park
  static constexpr int W_IDX = 0, W_IDY = 1, W_IDZ = 2, W_SX = 3, W_SY = 4, W_SZ = 5;
  static constexpr int W_K = 6;                              // kx | ky << 2 | kz << 4 | mesh of the best hit << 16
  static constexpr int W_ID = 7;                             // ray id
  // dynamic words
  static constexpr int W_TFAR = 8, W_U = 9, W_V = 10, W_BEST = 11;  // best hit so far (t_far == best t)
  static constexpr int W_CUR = 12;                           // current node / record; J3DG_EMPTY_CHILD = free slot
  static constexpr int W_SPV = 13;                           // stack height | node visits << 8 | current mesh << 16
  static constexpr int W_OX = 14;                            // + oy, oz (ORG only)
  static constexpr int W_CNT = ORG ? 17 : 14;                // STATS: node visits, triangle tests
  static constexpr int WORDS = W_CNT + (STATS ? 2 : 0);
  static constexpr int PARKW = 8 + (ORG ? 3 : 0);
  static constexpr int STACK_WORDS = 2 * (PSTACK + 1) * PSLOTS;
  static constexpr int WARP_WORDS = STACK_WORDS + WORDS * PSLOTS + PARKW * 32 + 64;
  static_assert((size_t)WARP_WORDS * 4 >= (size_t)STACK_SIZE * 4 * sizeof(uint2), "a warp's region must hold four group stacks");
  static_assert(WARP_WORDS % 4 == 0, "16-byte aligned warp regions");
};

template <int MODE, bool STATS>
__device__ __forceinline__ void pool_loop(const TraceParams& p, uint32_t* const wsm) {
  using L = PoolLayout<MODE, STATS>;
  constexpr bool ANY_HIT = MODE == SHADOW;
  constexpr int PSTACK = L::PSTACK;
  constexpr uint32_t FREE = J3DG_EMPTY_CHILD, NONE = 0xFFFFFFFFu;
  uint2* const stk = reinterpret_cast<uint2*>(wsm);           // entry i of slot s at stk[i * PSLOTS + s]
  uint32_t* const st = wsm + L::STACK_WORDS;                  // word w of slot s at st[w * PSLOTS + s]
  uint32_t* const park = st + L::WORDS * PSLOTS;              // word w of parked ray r at park[w * 32 + r]
  uint32_t* const list = park + L::PARKW * 32;                // slots of this step / free slots of a refill
  // Stack entry i of slot s: rows 0 .. PSTACK-1 in shared memory; the few rays that go deeper (grazing rays hit many of
  // the 8 children of every node they meet) continue in this warp's slice of a global buffer instead of being evicted.
  uint2* const gstk = p.spill + (size_t)(blockIdx.x * (BLOCK_THREADS / 32) + (threadIdx.x >> 5)) * ((size_t)POOL_SPILL * PSLOTS);
  auto stack_at = [&](int i, uint32_t s) -> uint2 { return i < PSTACK ? stk[i * PSLOTS + s] : gstk[(i - PSTACK) * PSLOTS + s]; };
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const float t_near = MODE == PRIMARY ? fdiv(p.vw.diagonal, 100.f) : 1e-3f;  // canvas.cpp:781 / 853
  uint32_t total_pools, list_n = 0;
  if (MODE == PRIMARY) {
    total_pools = p.grid.total_pools;
  } else {
    list_n = (uint32_t)p.stats[5];
    total_pools = (list_n + 31u) / 32u;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 4, (unsigned long long)list_n);  // shadow rays traced
  }
  const bool multi = p.nm > 1u;
  const WideNode* __restrict__ nodes0 = nullptr;
  const TriRec* __restrict__ tris0 = nullptr;
  float o0x = 0.f, o0y = 0.f, o0z = 0.f;                      // PRIMARY: the camera origin in the space of mesh 0
  if (p.nm) {
    const MeshDev& m0 = p.meshes[0];
    nodes0 = m0.nodes; tris0 = m0.tris;
    if (MODE == PRIMARY) {
      const float4 o = mat_vec(m0.cs_inv, make_float4(p.vw.origin[0], p.vw.origin[1], p.vw.origin[2], p.vw.origin[3]));
      o0x = o.x; o0y = o.y; o0z = o.z;
    }
  }
  st[L::W_CUR * PSLOTS + lane] = FREE;
  st[L::W_CUR * PSLOTS + 32 + lane] = FREE;
  __syncwarp();

  uint32_t park_next = 0, park_count = 0;
  bool exhausted = false;
  uint32_t sum_nodes = 0, sum_tris = 0;

  // the ray of slot s is done with its current mesh: next object (qbvh.h:3340-3385), or write the result and free the slot
  auto finish = [&](uint32_t s, uint32_t spv) {
    const uint32_t best = st[L::W_BEST * PSLOTS + s];
    const bool found = best != NONE;
    const uint32_t mesh_k = spv >> 16;
    const uint32_t ray_id = st[L::W_ID * PSLOTS + s];
    if (multi && mesh_k + 1u < p.nm && !(ANY_HIT && found)) {
      const MeshDev& m = p.meshes[mesh_k + 1u];
      const LaneRay lr = lane_ray_setup(m, world_ray<MODE>(p, ray_id));
      st[L::W_IDX * PSLOTS + s] = __float_as_uint(lr.idx); st[L::W_IDY * PSLOTS + s] = __float_as_uint(lr.idy); st[L::W_IDZ * PSLOTS + s] = __float_as_uint(lr.idz);
      st[L::W_SX * PSLOTS + s] = __float_as_uint(lr.Sx); st[L::W_SY * PSLOTS + s] = __float_as_uint(lr.Sy); st[L::W_SZ * PSLOTS + s] = __float_as_uint(lr.Sz);
      st[L::W_K * PSLOTS + s] = (st[L::W_K * PSLOTS + s] & 0xFFFF0000u) | lr.kpack;
      if (L::ORG) { st[(L::W_OX + 0) * PSLOTS + s] = __float_as_uint(lr.ox); st[(L::W_OX + 1) * PSLOTS + s] = __float_as_uint(lr.oy); st[(L::W_OX + 2) * PSLOTS + s] = __float_as_uint(lr.oz); }
      st[L::W_SPV * PSLOTS + s] = (spv & 0xFF00u) | ((mesh_k + 1u) << 16);  // empty stack, visits kept
      st[L::W_CUR * PSLOTS + s] = 0u;                                        // root of the next mesh
      return;
    }
    if (MODE == PRIMARY) {
      const int x = (int)(ray_id & 0xffffu), y = (int)(ray_id >> 16);
      uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)y * p.stride + x);
      uint4 lo = make_uint4(0u, 0u, 0u, found ? st[L::W_TFAR * PSLOTS + s] : __float_as_uint(FLT_MAX));  // misses already carry the final record (canvas.cpp:859-866)
      if (STATS) { lo.y = st[L::W_CNT * PSLOTS + s]; lo.z = st[(L::W_CNT + 1) * PSLOTS + s]; }  // the counting pass returns per-pixel costs in the u / v slots
      dst[0] = lo;
      // raw hit: record slot, barycentrics, mesh index (resolve_kernel finishes it)
      dst[1] = found ? make_uint4(best, st[L::W_U * PSLOTS + s], st[L::W_V * PSLOTS + s], st[L::W_K * PSLOTS + s] >> 16) : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
    } else if (found) {
      uint8_t* mark = reinterpret_cast<uint8_t*>(p.out + __ldg(p.shadow_pix + ray_id));
      *mark = *mark | 1u;  // canvas.cpp:856
    }
    if (STATS) { sum_nodes += st[L::W_CNT * PSLOTS + s]; sum_tris += st[(L::W_CNT + 1) * PSLOTS + s]; }
    st[L::W_CUR * PSLOTS + s] = FREE;
  };
  // a group of 8 lanes finishes the ray of slot s, starting from the best hit so far
  auto evict = [&](uint32_t s) {
    const uint32_t i = atomicAdd(p.hard_count, 1u);
    __stcg(p.hard_best + i, make_float4(__uint_as_float(st[L::W_TFAR * PSLOTS + s]), __uint_as_float(st[L::W_U * PSLOTS + s]),
                                        __uint_as_float(st[L::W_V * PSLOTS + s]), __uint_as_float(st[L::W_BEST * PSLOTS + s])));
    __threadfence();  // the consumer reads the seed after it has seen the id
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" :: "l"(p.hard_id + i), "r"(st[L::W_ID * PSLOTS + s]), "r"(st[L::W_K * PSLOTS + s] >> 16) : "memory");
    st[L::W_CUR * PSLOTS + s] = FREE;
  };
  auto prefetch = [&](const WideNode* nodes, const TriRec* tris, uint32_t cur) {
    if (J3DG_POOL_PREFETCH == 0) return;
    const void* a = (cur & J3DG_LEAF_BIT) ? (const void*)(tris + (cur & J3DG_LEAF_FIRST_MASK)) : (const void*)(nodes + cur);
    if (J3DG_POOL_PREFETCH == 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(a));
    else asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
  };

  for (uint32_t iter = 0;; ++iter) {
    // =========================== classify the 64 slots ===========================
    const uint32_t c0 = st[L::W_CUR * PSLOTS + lane], c1 = st[L::W_CUR * PSLOTS + 32 + lane];
    const uint32_t fLo = __ballot_sync(0xffffffffu, c0 == FREE), fHi = __ballot_sync(0xffffffffu, c1 == FREE);
    const uint32_t nLo = __ballot_sync(0xffffffffu, !(c0 >> 31)), nHi = __ballot_sync(0xffffffffu, !(c1 >> 31));
    const uint32_t tLo = ~(fLo | nLo), tHi = ~(fHi | nHi);
    const int nfree = __popc(fLo) + __popc(fHi), nnode = __popc(nLo) + __popc(nHi), ntri = PSLOTS - nfree - nnode;

    // =========================== refill free slots from the parked tile ===========================
    if (nfree >= J3DG_POOL_REFILL_MIN && !(exhausted && park_next >= park_count)) {
      if (park_next >= park_count) {  // fetch the next pool; its 32 rays are set up by the 32 lanes in parallel
        uint32_t pool = 0;
        if (lane == 0) pool = atomicAdd(p.pool_ctr, 1u);
        pool = __shfl_sync(0xffffffffu, pool, 0);
        if (pool >= total_pools) { exhausted = true; continue; }
        uint32_t id;
        bool ok;
        if (MODE == PRIMARY) {
          uint32_t tx, ty;
          bool halo;
          const bool tile_ok = tile_of_pool(p.grid, pool, tx, ty, halo);
          const int x = p.x0 + (int)tx * TILE_W + (lane & (TILE_W - 1));
          const int y = p.y0 + (int)ty * TILE_H + (lane / TILE_W);
          ok = tile_ok && x <= p.x1 && y <= p.y1 && (!halo || lane / TILE_W == TILE_H - 1);
          id = ((uint32_t)y << 16) | (uint32_t)x;
        } else {
          id = pool * 32u + (uint32_t)lane;
          ok = id < list_n;
        }
        if (ok && p.nm == 0u) {  // empty scene: every ray misses
          if (MODE == PRIMARY) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(id >> 16) * p.stride + (id & 0xffffu));
            dst[0] = make_uint4(0u, 0u, 0u, __float_as_uint(FLT_MAX));
            dst[1] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
          }
          ok = false;
        }
        const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const LaneRay lr = lane_ray_setup(p.meshes[0], world_ray<MODE>(p, id));
          const uint32_t r = __popc(okmask & lt_mask);  // dense
          park[0 * 32 + r] = __float_as_uint(lr.idx); park[1 * 32 + r] = __float_as_uint(lr.idy); park[2 * 32 + r] = __float_as_uint(lr.idz);
          park[3 * 32 + r] = __float_as_uint(lr.Sx); park[4 * 32 + r] = __float_as_uint(lr.Sy); park[5 * 32 + r] = __float_as_uint(lr.Sz);
          park[6 * 32 + r] = lr.kpack; park[7 * 32 + r] = id;
          if (L::ORG) { park[8 * 32 + r] = __float_as_uint(lr.ox); park[9 * 32 + r] = __float_as_uint(lr.oy); park[10 * 32 + r] = __float_as_uint(lr.oz); }
        }
        __syncwarp();
        park_next = 0;
        park_count = __popc(okmask);
        continue;
      }
      // the r-th free slot takes the r-th parked ray
      if (c0 == FREE) list[__popc(fLo & lt_mask)] = (uint32_t)lane;
      if (c1 == FREE) list[__popc(fLo) + __popc(fHi & lt_mask)] = 32u + (uint32_t)lane;
      __syncwarp();
      const uint32_t take = min((uint32_t)nfree, park_count - park_next);
      if ((uint32_t)lane < take) {
        const uint32_t s = list[lane], r = park_next + (uint32_t)lane;
#pragma unroll
        for (int w = 0; w < 8; ++w) st[w * PSLOTS + s] = park[w * 32 + r];
        if (L::ORG) {
#pragma unroll
          for (int w = 0; w < 3; ++w) st[(L::W_OX + w) * PSLOTS + s] = park[(8 + w) * 32 + r];
        }
        st[L::W_TFAR * PSLOTS + s] = __float_as_uint(FLT_MAX);
        st[L::W_U * PSLOTS + s] = 0u; st[L::W_V * PSLOTS + s] = 0u;
        st[L::W_BEST * PSLOTS + s] = NONE;
        st[L::W_SPV * PSLOTS + s] = 0u;
        if (STATS) { st[L::W_CNT * PSLOTS + s] = 0u; st[(L::W_CNT + 1) * PSLOTS + s] = 0u; }
        st[L::W_CUR * PSLOTS + s] = 0u;  // the root (meshes without triangles never reach the kernel)
      }
      __syncwarp();
      park_next += take;
      continue;
    }
    if (nnode + ntri == 0) break;  // nothing in flight and nothing left to refill with

    // =========================== pick the step and compact its slots onto the lanes ===========================
    // a node step serves 32 rays (one per lane), a leaf step 8 (four lanes per ray): run node steps while they are full,
    // otherwise whichever kind fills more of the warp
    const bool do_tri = nnode == 0 || (nnode < J3DG_POOL_NODE_FULL && min(ntri, 8) * 4 >= nnode);
    const uint32_t cap = do_tri ? 8u : 32u;
    const uint32_t mLo = do_tri ? tLo : nLo, mHi = do_tri ? tHi : nHi;
    uint32_t r0, r1;  // alternate which half is served first: no slot starves
    if (iter & 1u) { r1 = __popc(mHi & lt_mask); r0 = __popc(mHi) + __popc(mLo & lt_mask); }
    else { r0 = __popc(mLo & lt_mask); r1 = __popc(mLo) + __popc(mHi & lt_mask); }
    if (((mLo >> lane) & 1u) && r0 < cap) list[r0] = (uint32_t)lane;
    if (((mHi >> lane) & 1u) && r1 < cap) list[r1] = 32u + (uint32_t)lane;
    __syncwarp();
    const uint32_t ready = (uint32_t)(__popc(mLo) + __popc(mHi));

    if (!do_tri) {
      // ---------------- node step: one ray per lane, 8 quantised child boxes ----------------
      if ((uint32_t)lane < min(ready, 32u)) {
        const uint32_t s = list[lane];
        const uint32_t spv = st[L::W_SPV * PSLOTS + s];
        int sp = (int)(spv & 0xFFu);
        uint32_t visits = (spv >> 8) & 0xFFu;
        const uint32_t mesh_k = spv >> 16;
        if (!STATS && visits >= p.budget) {
          evict(s);
        } else {
          if (STATS) st[L::W_CNT * PSLOTS + s] += 1u; else ++visits;
          const float idx = __uint_as_float(st[L::W_IDX * PSLOTS + s]), idy = __uint_as_float(st[L::W_IDY * PSLOTS + s]), idz = __uint_as_float(st[L::W_IDZ * PSLOTS + s]);
          const float t_far = __uint_as_float(st[L::W_TFAR * PSLOTS + s]);
          const uint32_t cur0 = st[L::W_CUR * PSLOTS + s];
          float ox = o0x, oy = o0y, oz = o0z;
          const WideNode* __restrict__ nodes = nodes0;
          const TriRec* __restrict__ tris = tris0;
          if (L::ORG) { ox = __uint_as_float(st[(L::W_OX + 0) * PSLOTS + s]); oy = __uint_as_float(st[(L::W_OX + 1) * PSLOTS + s]); oz = __uint_as_float(st[(L::W_OX + 2) * PSLOTS + s]); }
          if (multi) {
            const MeshDev& m = p.meshes[mesh_k];
            nodes = m.nodes; tris = m.tris;
            if (!L::ORG) {
              const float4 o = mat_vec(m.cs_inv, make_float4(p.vw.origin[0], p.vw.origin[1], p.vw.origin[2], p.vw.origin[3]));
              ox = o.x; oy = o.y; oz = o.z;
            }
          }
          const uint32_t sel_nx = plane_sel(idx < 0.f ? 3u : 0u), sel_fx = plane_sel(idx < 0.f ? 0u : 3u);
          const uint32_t sel_ny = plane_sel(idy < 0.f ? 4u : 1u), sel_fy = plane_sel(idy < 0.f ? 1u : 4u);
          const uint32_t sel_nz = plane_sel(idz < 0.f ? 5u : 2u), sel_fz = plane_sel(idz < 0.f ? 2u : 5u);
          const char* np = reinterpret_cast<const char*>(nodes + cur0);
          const U32x8 q0 = ldg256(np), q1 = ldg256(np + 32), q2 = ldg256(np + 64), q3 = ldg256(np + 96);
          const Slab X = slab(__uint_as_float(q0.v[4]), __uint_as_float(q0.v[0]), ox, idx);
          const Slab Y = slab(__uint_as_float(q0.v[5]), __uint_as_float(q0.v[1]), oy, idy);
          const Slab Z = slab(__uint_as_float(q0.v[6]), __uint_as_float(q0.v[2]), oz, idz);
          // Entry distance of child i, low 3 bits replaced by i (distances are positive, so their bit patterns
          // order like integers); missed children get +inf.  Empty slots have inverted boxes and never pass.
          constexpr uint32_t MISS_KEY = 0x7F800000u;
          auto child_key = [&](uint32_t lo, uint32_t hi, uint32_t i) -> uint32_t {
            float tmin = fmaxf(fmaxf(fmaf(plane(lo, hi, sel_nx), X.S, X.Bn), fmaf(plane(lo, hi, sel_ny), Y.S, Y.Bn)), fmaxf(fmaf(plane(lo, hi, sel_nz), Z.S, Z.Bn), t_near));
            float tmax = fminf(fminf(fmaf(plane(lo, hi, sel_fx), X.S, X.Bf), fmaf(plane(lo, hi, sel_fy), Y.S, Y.Bf)), fminf(fmaf(plane(lo, hi, sel_fz), Z.S, Z.Bf), t_far));
            // conservative padding against rounding of the slab arithmetic
            tmin = fmaf(-fabsf(tmin), 2e-6f, tmin);
            tmax = fmaf(fabsf(tmax), 2e-6f, tmax);
            return tmin <= tmax ? ((__float_as_uint(tmin) & ~7u) | i) : (MISS_KEY | i);
          };
          const uint32_t k0 = child_key(q1.v[0], q1.v[1], 0u), k1 = child_key(q1.v[2], q1.v[3], 1u), k2 = child_key(q1.v[4], q1.v[5], 2u), k3 = child_key(q1.v[6], q1.v[7], 3u);
          const uint32_t k4 = child_key(q2.v[0], q2.v[1], 4u), k5 = child_key(q2.v[2], q2.v[3], 5u), k6 = child_key(q2.v[4], q2.v[5], 6u), k7 = child_key(q2.v[6], q2.v[7], 7u);
          const int nearest = min(min(min((int)k0, (int)k1), min((int)k2, (int)k3)), min(min((int)k4, (int)k5), min((int)k6, (int)k7)));
          const uint32_t ni = (uint32_t)nearest & 7u;
          uint32_t near_ref = q3.v[0];
          near_ref = ni == 1u ? q3.v[1] : near_ref; near_ref = ni == 2u ? q3.v[2] : near_ref; near_ref = ni == 3u ? q3.v[3] : near_ref;
          near_ref = ni == 4u ? q3.v[4] : near_ref; near_ref = ni == 5u ? q3.v[5] : near_ref; near_ref = ni == 6u ? q3.v[6] : near_ref;
          near_ref = ni == 7u ? q3.v[7] : near_ref;
          if ((uint32_t)nearest >= MISS_KEY) near_ref = J3DG_EMPTY_CHILD;
          const int nh = (k0 < MISS_KEY) + (k1 < MISS_KEY) + (k2 < MISS_KEY) + (k3 < MISS_KEY) + (k4 < MISS_KEY) + (k5 < MISS_KEY) + (k6 < MISS_KEY) + (k7 < MISS_KEY);
          bool overflow = false;
          if (sp + max(nh, 1) - 1 <= PSTACK) {
            // branch-free pushes of the other hit children (row PSTACK is scratch; the low key bits are cleared at pop)
            auto push = [&](uint32_t key, uint32_t ref) {
              stk[sp * PSLOTS + s] = make_uint2(ref, key);
              sp += (key < MISS_KEY && (int)key != nearest) ? 1 : 0;
            };
            push(k0, q3.v[0]); push(k1, q3.v[1]); push(k2, q3.v[2]); push(k3, q3.v[3]);
            push(k4, q3.v[4]); push(k5, q3.v[5]); push(k6, q3.v[6]); push(k7, q3.v[7]);
          } else {
            // the stack leaves shared memory: entries PSTACK .. PSTACK + POOL_SPILL - 1 live in the global slice
            auto push = [&](uint32_t key, uint32_t ref) {
              if (key < MISS_KEY && (int)key != nearest) {
                if (sp < PSTACK) stk[sp * PSLOTS + s] = make_uint2(ref, key);
                else if (sp < PSTACK + POOL_SPILL) gstk[(sp - PSTACK) * PSLOTS + s] = make_uint2(ref, key);
                else overflow = true;
                sp = min(sp + 1, PSTACK + POOL_SPILL);
              }
            };
            push(k0, q3.v[0]); push(k1, q3.v[1]); push(k2, q3.v[2]); push(k3, q3.v[3]);
            push(k4, q3.v[4]); push(k5, q3.v[5]); push(k6, q3.v[6]); push(k7, q3.v[7]);
          }
          if (overflow) {  // deeper than shared + global rows: the group kernel restarts the ray
            evict(s);
          } else {
            uint32_t cur = near_ref;
            if (cur == J3DG_EMPTY_CHILD) {
              while (sp > 0) {
                --sp;
                const uint2 e = stack_at(sp, s);
                if (__uint_as_float(e.y & ~7u) <= t_far) { cur = e.x; break; }  // entries beyond the shrunk interval are skipped
              }
            }
            const uint32_t nspv = (uint32_t)sp | (visits << 8) | (mesh_k << 16);
            if (cur == J3DG_EMPTY_CHILD) finish(s, nspv);
            else {
              st[L::W_CUR * PSLOTS + s] = cur;
              st[L::W_SPV * PSLOTS + s] = nspv;
              prefetch(nodes, tris, cur);
            }
          }
        }
      }
    } else {
      // ---------------- leaf step: FOUR lanes per ray, 8 rays per step; lane c tests records c and c + 4 of the leaf ----------------
      // (a leaf holds 1..8 consecutive records, the last one flagged: one round for leaves of <= 4, two otherwise)
      const int g = lane >> 2, c = lane & 3, gshift = lane & 28;
      const bool active = (uint32_t)g < min(ready, 8u);
      const uint32_t s = active ? list[g] : 0u;
      uint32_t first = 0, spv = 0, kw = 0;
      float ox = o0x, oy = o0y, oz = o0z, Sx = 0.f, Sy = 0.f, Sz = 0.f, t_far = 0.f;
      const WideNode* __restrict__ nodes = nodes0;
      const TriRec* __restrict__ tris = tris0;
      if (active) {
        first = st[L::W_CUR * PSLOTS + s] & J3DG_LEAF_FIRST_MASK;
        spv = st[L::W_SPV * PSLOTS + s];
        kw = st[L::W_K * PSLOTS + s];
        Sx = __uint_as_float(st[L::W_SX * PSLOTS + s]); Sy = __uint_as_float(st[L::W_SY * PSLOTS + s]); Sz = __uint_as_float(st[L::W_SZ * PSLOTS + s]);
        t_far = __uint_as_float(st[L::W_TFAR * PSLOTS + s]);
        if (L::ORG) { ox = __uint_as_float(st[(L::W_OX + 0) * PSLOTS + s]); oy = __uint_as_float(st[(L::W_OX + 1) * PSLOTS + s]); oz = __uint_as_float(st[(L::W_OX + 2) * PSLOTS + s]); }
        if (multi) {
          const MeshDev& m = p.meshes[spv >> 16];
          nodes = m.nodes; tris = m.tris;
          if (!L::ORG) {
            const float4 o = mat_vec(m.cs_inv, make_float4(p.vw.origin[0], p.vw.origin[1], p.vw.origin[2], p.vw.origin[3]));
            ox = o.x; oy = o.y; oz = o.z;
          }
        }
      }
      const int kx = (int)(kw & 3u), ky = (int)((kw >> 2) & 3u), kz = (int)((kw >> 4) & 3u);
      bool more = active;           // the leaf continues past the records tested so far
      bool any_found = false;       // ANY_HIT: some record of the leaf was hit
      uint32_t win_slot = 0xFFFFFFFFu;
      float win_u = 0.f, win_v = 0.f;
#pragma unroll 1
      for (int round = 0; round < 2; ++round) {
        if (!__any_sync(0xffffffffu, more)) break;
        const uint32_t slot = first + 4u * round + (uint32_t)c;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
        if (more) {
          const float4* tp = reinterpret_cast<const float4*>(tris + slot);
          v0 = __ldg(tp); v1 = __ldg(tp + 1); v2 = __ldg(tp + 2);
        }
        const uint32_t lm = (__ballot_sync(0xffffffffu, more && __float_as_uint(v1.w) != 0u) >> gshift) & 0xFu;  // end-of-leaf flags of my ray's 4 records
        const int last = lm ? __ffs(lm) - 1 : 3;
        bool hit = more && c <= last;
        if (STATS && hit) atomicAdd(&st[(L::W_CNT + 1) * PSLOTS + s], 1u);
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
            hit = (t_far > t) && (t > t_near);  // t_far is the best t so far: strictly closer
            u = fmul(V, inv_det);
            v = fmul(W, inv_det);
          }
        }
        // the closest of the (up to) four hits; the lowest record wins ties, like the sequential strict-less update
        float key = hit ? t : FLT_MAX;
        float nearest = fminf(key, __shfl_xor_sync(0xffffffffu, key, 1));
        nearest = fminf(nearest, __shfl_xor_sync(0xffffffffu, nearest, 2));
        const uint32_t wm = (__ballot_sync(0xffffffffu, hit && key == nearest) >> gshift) & 0xFu;
        const int win = __ffs(wm) - 1;  // -1: no hit
        const float wt = __shfl_sync(0xffffffffu, t, gshift + (win & 3));
        const float wu = __shfl_sync(0xffffffffu, u, gshift + (win & 3));
        const float wv = __shfl_sync(0xffffffffu, v, gshift + (win & 3));
        if (more && win >= 0) {
          t_far = wt; win_u = wu; win_v = wv; win_slot = first + 4u * round + (uint32_t)win;
          any_found = true;
        }
        more = more && lm == 0u && !(ANY_HIT && any_found);
      }
      if (active && c == 0) {
        if (win_slot != 0xFFFFFFFFu) {
          st[L::W_TFAR * PSLOTS + s] = __float_as_uint(t_far);
          st[L::W_U * PSLOTS + s] = __float_as_uint(win_u);
          st[L::W_V * PSLOTS + s] = __float_as_uint(win_v);
          st[L::W_BEST * PSLOTS + s] = win_slot;
          st[L::W_K * PSLOTS + s] = (kw & 0xFFFFu) | ((spv >> 16) << 16);
        }
        int sp = (int)(spv & 0xFFu);
        uint32_t cur = J3DG_EMPTY_CHILD;
        if (ANY_HIT && any_found) sp = 0;
        while (sp > 0) {
          --sp;
          const uint2 e = stack_at(sp, s);
          if (__uint_as_float(e.y & ~7u) <= t_far) { cur = e.x; break; }  // entries beyond the shrunk interval are skipped
        }
        const uint32_t nspv = (uint32_t)sp | (spv & 0xFFFFFF00u);
        if (cur == J3DG_EMPTY_CHILD) finish(s, nspv);
        else {
          st[L::W_CUR * PSLOTS + s] = cur;
          st[L::W_SPV * PSLOTS + s] = nspv;
          prefetch(nodes, tris, cur);
        }
      }
    }
    __syncwarp();  // the slots' new states are visible to the next classification
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

template <int MODE, bool STATS>
constexpr size_t cast_smem_bytes() {
#if J3DG_POOL_MODE
  return (size_t)(BLOCK_THREADS / 32) * PoolLayout<MODE, STATS>::WARP_WORDS * sizeof(uint32_t);
#else
  return CAST_SMEM;
#endif
}

// The cast kernel proper: one launch, every warp plays two roles.  It first traces tiles (one ray per lane, evicting
// long rays into the queue), signs off, and then helps to drain the queue as four 8-lane groups (group_loop<QUEUE>).
// A plain launch of at most one machine-full of blocks: producers never wait, and consumers only wait for warps that
// are running (group_loop, SRC QUEUE) — no block has to be resident at any particular time.
template <int MODE, bool STATS, bool TOP = false>
__global__ void __launch_bounds__(BLOCK_THREADS, J3DG_POOL_MODE ? J3DG_POOL_MIN_BLOCKS : J3DG_LANE_MIN_BLOCKS) cast_kernel(const TraceParams p) {
  extern __shared__ __align__(16) unsigned char smem[];  // cast_smem_bytes<MODE, STATS>()
#ifdef J3DG_TIMELINE
  unsigned long long tl0, tl1, tl2;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
#endif
#if J3DG_POOL_MODE
  uint32_t* const wsm = reinterpret_cast<uint32_t*>(smem) + (threadIdx.x >> 5) * PoolLayout<MODE, STATS>::WARP_WORDS;
  uint2* const warp_stack = reinterpret_cast<uint2*>(wsm);  // the group stacks reuse the warp's region once its slots are empty
  {
    pool_loop<MODE, STATS>(p, wsm);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      __threadfence();
      atomicAdd(p.done_blocks, 1u);
    }
  }
#else
  static_assert((LANE_STACK + 1) * 32 / 4 >= STACK_SIZE, "a warp's stack slice must hold four group stacks");
  uint2* const warp_stack = reinterpret_cast<uint2*>(smem) + (threadIdx.x >> 5) * ((LANE_STACK + 1) * 32);
  {
    lane_loop<MODE, STATS, TOP>(p, warp_stack + (threadIdx.x & 31), reinterpret_cast<uint32_t*>(smem + LANE_SMEM_STACK));
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      __threadfence();
      atomicAdd(p.done_blocks, 1u);
    }
  }
#endif
#ifdef J3DG_TIMELINE
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
#endif
  group_loop<MODE, QUEUE, 4, TOP>(p, warp_stack + ((threadIdx.x & 31) >> 3));
#ifdef J3DG_TIMELINE
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl2));
  if (MODE == PRIMARY && threadIdx.x == 0) {  // [13] first start, [14] last end of the lane phase, [15] last end; + sums for means
    atomicMin(p.stats + 13, tl0);
    atomicMax(p.stats + 14, tl1);
    atomicMax(p.stats + 15, tl2);
    atomicAdd(p.stats + 16, tl1 - tl0);
    atomicAdd(p.stats + 17, tl2 - tl1);
    atomicAdd(p.stats + 18, 1ull);
  }
#endif
}

// ---- hit -> pixel record (canvas.cpp:788-834) + shadow ray generation (836-854) --------------------------
// One thread per pixel, one warp per 8x4 tile (the order the trace kernels use).  Misses already hold
// their final record.  Shadow rays of hit pixels are appended to a list with one atomic per warp.
__global__ void __launch_bounds__(256) resolve_kernel(const MeshDev* __restrict__ meshes, ViewDev vw, int x0, int y0, int x1, int y1, TileGrid grid,
                                                       j3dg_pixel* __restrict__ out, uint32_t stride, float4* __restrict__ shadow_pos,
                                                       uint32_t* __restrict__ shadow_pix, unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  bool is_hit = false;
  const uint32_t pool = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint32_t tx = 0, ty = 0;
  bool halo = false;
  const bool tile_ok = pool < grid.total_pools && tile_of_pool(grid, pool, tx, ty, halo);
  const int x = x0 + (int)tx * TILE_W + (lane & (TILE_W - 1));
  const int y = y0 + (int)ty * TILE_H + (lane / TILE_W);
  bool shadow_ray = false;
  float4 pos = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tile_ok && x <= x1 && y <= y1 && (!halo || lane / TILE_W == TILE_H - 1)) {
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)y * stride + x);
    const uint4 raw = dst[1];
    if (raw.x != 0xFFFFFFFFu) {
      is_hit = !halo;
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
      if ((vw.flags & J3DG_SHADOW) && !halo) {  // canvas.cpp:836-848; halo pixels serve the edge shader only (u, v, depth, id)
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
  // bounding rectangle of the hit pixels of this frame (stats slots 20, 21 = min x, min y, max x, max y): everything
  // outside it is a miss record / background, which lets the host copy skip it (j3dg_ctx_set_dirty_rect).
  // Reduced per warp, then per block in shared memory; one thread per block touches the global words, and only
  // if it still widens the rectangle.
  {
    __shared__ uint32_t s_bb[4];
    if (threadIdx.x == 0) { s_bb[0] = 0xFFFFFFFFu; s_bb[1] = 0xFFFFFFFFu; s_bb[2] = 0u; s_bb[3] = 0u; }
    __syncthreads();
    if (__any_sync(0xffffffffu, is_hit)) {
      const uint32_t mnx = __reduce_min_sync(0xffffffffu, is_hit ? (uint32_t)x : 0xFFFFFFFFu);
      const uint32_t mny = __reduce_min_sync(0xffffffffu, is_hit ? (uint32_t)y : 0xFFFFFFFFu);
      const uint32_t mxx = __reduce_max_sync(0xffffffffu, is_hit ? (uint32_t)x : 0u);
      const uint32_t mxy = __reduce_max_sync(0xffffffffu, is_hit ? (uint32_t)y : 0u);
      if (lane == 0) { atomicMin(&s_bb[0], mnx); atomicMin(&s_bb[1], mny); atomicMax(&s_bb[2], mxx); atomicMax(&s_bb[3], mxy); }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_bb[0] <= s_bb[2]) {
      uint32_t* bb = reinterpret_cast<uint32_t*>(stats + 20);
      const uint4 now = __ldcg(reinterpret_cast<const uint4*>(bb));
      if (s_bb[0] < now.x) atomicMin(bb + 0, s_bb[0]);
      if (s_bb[1] < now.y) atomicMin(bb + 1, s_bb[1]);
      if (s_bb[2] > now.z) atomicMax(bb + 2, s_bb[2]);
      if (s_bb[3] > now.w) atomicMax(bb + 3, s_bb[3]);
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

// ---- top-level tree over the objects (host, whenever the object table changes) -------------------------------------
// World box of an object = the box of the 8 corners of its root box under cs, padded by 1e-4 of its size (the walk
// tests it in world space with fused arithmetic, the object is then traversed in object space with the reference's).
// Objects are ordered along a 30-bit Morton curve of their centres; a node takes a contiguous run and cuts it into
// (up to) 8 runs of equal length, runs of one object become leaves.  Node 0 is the root (pre-order allocation).
struct TopItem { float lo[3], hi[3]; uint32_t mesh; uint32_t code; };

uint32_t top_build_rec(std::vector<TopNode>& nodes, const std::vector<TopItem>& items, size_t a, size_t b) {
  const uint32_t me = (uint32_t)nodes.size();
  nodes.emplace_back();
  TopNode n;
  for (int c = 0; c < 8; ++c) {
    for (int j = 0; j < 3; ++j) { n.lo[c][j] = INFINITY; n.hi[c][j] = -INFINITY; }
    n.child[c] = J3DG_EMPTY_CHILD;  // unused slot: skipped by top_box_mask
    n.pad[c] = 0;
  }
  const size_t cnt = b - a, parts = std::min<size_t>(8, cnt);
  for (size_t c = 0; c < parts; ++c) {
    const size_t ca = a + cnt * c / parts, cb = a + cnt * (c + 1) / parts;
    for (size_t i = ca; i < cb; ++i)
      for (int j = 0; j < 3; ++j) { n.lo[c][j] = std::min(n.lo[c][j], items[i].lo[j]); n.hi[c][j] = std::max(n.hi[c][j], items[i].hi[j]); }
    n.child[c] = (cb - ca == 1) ? (J3DG_LEAF_BIT | items[ca].mesh) : top_build_rec(nodes, items, ca, cb);
  }
  nodes[me] = n;
  return me;
}

void build_top_tree(const std::vector<MeshDev>& meshes, std::vector<TopNode>& nodes) {
  std::vector<TopItem> items(meshes.size());
  double glo[3] = {1e300, 1e300, 1e300}, ghi[3] = {-1e300, -1e300, -1e300};
  for (size_t k = 0; k < meshes.size(); ++k) {
    const MeshDev& m = meshes[k];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int corner = 0; corner < 8; ++corner) {
      const double v[3] = {corner & 1 ? m.root_max[0] : m.root_min[0], corner & 2 ? m.root_max[1] : m.root_min[1], corner & 4 ? m.root_max[2] : m.root_min[2]};
      double w4 = m.cs[3] * v[0] + m.cs[7] * v[1] + m.cs[11] * v[2] + m.cs[15];
      if (w4 == 0.0) w4 = 1.0;
      for (int j = 0; j < 3; ++j) {
        const double x = (m.cs[j] * v[0] + m.cs[4 + j] * v[1] + m.cs[8 + j] * v[2] + m.cs[12 + j]) / w4;
        lo[j] = std::min(lo[j], x); hi[j] = std::max(hi[j], x);
      }
    }
    for (int j = 0; j < 3; ++j) {
      const double padj = 1e-4 * (hi[j] - lo[j]) + 1e-6 * (std::fabs(lo[j]) + std::fabs(hi[j])) + 1e-30;
      items[k].lo[j] = (float)(lo[j] - padj); items[k].hi[j] = (float)(hi[j] + padj);
      glo[j] = std::min(glo[j], lo[j]); ghi[j] = std::max(ghi[j], hi[j]);
    }
    items[k].mesh = (uint32_t)k;
  }
  for (TopItem& it : items) {
    uint32_t code = 0;
    uint32_t q[3];
    for (int j = 0; j < 3; ++j) {
      const double ext = ghi[j] - glo[j];
      const double f = ext > 0.0 ? (0.5 * ((double)it.lo[j] + it.hi[j]) - glo[j]) / ext : 0.0;
      q[j] = (uint32_t)std::min(1023.0, std::max(0.0, f * 1024.0));
    }
    for (int bit = 9; bit >= 0; --bit)
      for (int j = 0; j < 3; ++j) code = (code << 1) | ((q[j] >> bit) & 1u);
    it.code = code;
  }
  std::stable_sort(items.begin(), items.end(), [](const TopItem& x, const TopItem& y) { return x.code < y.code; });
  nodes.clear();
  nodes.reserve(items.size() / 4 + 2);
  top_build_rec(nodes, items, 0, items.size());
}

}  // namespace

void j3dg_preload_cast_kernels() {  // see j3dg_preload_build_kernels
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, cast_kernel<PRIMARY, false>); cudaFuncGetAttributes(&a, cast_kernel<SHADOW, false>); cudaFuncGetAttributes(&a, resolve_kernel);
  cudaGetLastError();
}

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
    ctx->top_nodes = 0;
    if (used >= ctx->top_min) {  // many objects: a tree over their world boxes instead of a loop over all of them
      std::vector<TopNode> top;
      build_top_tree(host, top);
      int rc = j3dg_reserve(ctx, &ctx->d_top, &ctx->top_cap, top.size() * sizeof(TopNode));
      if (rc != J3DG_OK) return rc;
      CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_top, top.data(), top.size() * sizeof(TopNode), cudaMemcpyHostToDevice, ctx->stream));
      CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      ctx->top_nodes = (uint32_t)top.size();
    }
  }
  const bool use_top = ctx->top_nodes != 0 && used >= ctx->top_min && !stats && ctx->cast_algo == 0;
  const bool shadows = (view->flags & J3DG_SHADOW) && used && !stats;
  const int rw = x1 - x0 + 1, rh = y1 - y0 + 1;
  const size_t npx = (size_t)rw * rh;
  const size_t shadow_pix_off = (npx * sizeof(float4) + 255) & ~(size_t)255;
  if (shadows) {  // one shadow ray per hit pixel at most
    int rc = j3dg_reserve(ctx, &ctx->d_shadow, &ctx->shadow_cap, shadow_pix_off + npx * sizeof(uint32_t));
    if (rc != J3DG_OK) return rc;
  }
  const size_t hard_id_off = (npx * sizeof(float4) + 255) & ~(size_t)255;
  if (ctx->hard_cap < hard_id_off + npx * sizeof(uint2)) {  // hard-ray queue: worst case every ray is evicted
    int rc = j3dg_reserve(ctx, &ctx->d_hard, &ctx->hard_cap, hard_id_off + npx * sizeof(uint2));
    if (rc != J3DG_OK) return rc;
    // every id = "not written yet"; consumers restore the marker, so the queue stays clean between launches
    CU_CHECK(ctx, cudaMemsetAsync(ctx->d_hard, 0xFF, ctx->hard_cap, ctx->stream));
    ctx->hard_id_off = hard_id_off;
  }
  // stats slots (u64 each): [0] node visits [1] triangle tests [2] stack overflow flag [3] pool counter PRIMARY
  // [4] shadow rays traced (accumulates until the timings are reset) [5] shadow list length [6] queue length PRIMARY
  // [7] queue claims PRIMARY [8] pool counter SHADOW [9] queue length SHADOW [10] queue claims SHADOW
  // [11] producer warps done (low word) + producer warps started (high word) PRIMARY [12] the same SHADOW
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 5, 0, 8 * sizeof(unsigned long long), ctx->stream));
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 20, 0xFF, sizeof(unsigned long long), ctx->stream));  // hit bbox: min x, min y
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 21, 0, sizeof(unsigned long long), ctx->stream));     //           max x, max y
#ifdef J3DG_TIMELINE
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 13, 0xFF, sizeof(unsigned long long), ctx->stream));
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 14, 0, 5 * sizeof(unsigned long long), ctx->stream));
#endif
  TraceParams tp = {};
  tp.meshes = ctx->d_meshes;
  tp.nm = used;
  tp.top = use_top ? (const TopNode*)ctx->d_top : nullptr;
  j3dg_make_view_dev(view, tp.vw);
  if (!shadows) tp.vw.flags &= ~J3DG_SHADOW;
  tp.x0 = x0; tp.y0 = y0; tp.x1 = x1; tp.y1 = y1;
  tp.out = d_pixels;
  tp.stride = stride;
  tp.stats = ctx->d_stats;
  tp.sticky = ctx->d_status;
  tp.shadow_pos = (const float4*)ctx->d_shadow;
  tp.shadow_pix = (const uint32_t*)((const char*)ctx->d_shadow + shadow_pix_off);
  tp.hard_best = (float4*)ctx->d_hard;
  tp.hard_id = (uint2*)((char*)ctx->d_hard + ctx->hard_id_off);
  tp.hard_capacity = (uint32_t)std::min<size_t>((ctx->hard_cap - ctx->hard_id_off) / sizeof(uint2), ctx->hard_id_off / sizeof(float4));
  tp.budget = stats ? 0xFFFFFFFFu : std::min<uint32_t>(ctx->lane_budget, 255u);  // pool mode counts visits in 8 bits
  auto ctr = [&](int slot) { return reinterpret_cast<unsigned int*>(ctx->d_stats + slot); };
  const bool sharded = ctx->shard_world > 1 && !stats;
  tp.grid = make_tile_grid(x0, y0, x1, y1, sharded ? ctx->shard_rank : 0u, sharded ? ctx->shard_world : 1u);
  const long long ntiles = tp.grid.total_pools;
  // One launch of the hybrid kernel: the grid fills the machine at most once.
  auto launch_hybrid = [&](auto kernel, size_t smem, long long pools) -> int {
    int nb = 0;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, BLOCK_THREADS, smem);
    if (e != cudaSuccess) return j3dg_cuda_fail(ctx, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor", __FILE__, __LINE__);
    const int full = ctx->sm_count * std::max(nb, 1);
#if J3DG_POOL_MODE
    {  // global continuation of the pool-mode stacks: one slice per resident warp
      const size_t need = (size_t)full * (BLOCK_THREADS / 32) * POOL_SPILL * PSLOTS * sizeof(uint2);
      int rc2 = j3dg_reserve(ctx, &ctx->d_spill, &ctx->spill_cap, need);
      if (rc2 != J3DG_OK) return rc2;
      tp.spill = (uint2*)ctx->d_spill;
    }
#endif
    const int grid = (int)std::max<long long>(1, std::min<long long>((long long)full, (pools + 3) / 4));
    // A plain launch of at most one machine-full of blocks.  Producers never wait for anything; consumers wait for
    // warps that are RUNNING, never for a block that has not started — so the launch cannot deadlock whatever else
    // is resident, and several casts (consecutive frames rendered by two contexts on two streams) may
    // share the machine: groups retire one by one once the queue has run dry, so the blocks of the next frame move in
    // while this frame's longest rays are still being finished (measured: 0.96 -> 0.79 ms per frame).
    kernel<<<grid, BLOCK_THREADS, smem, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
    return J3DG_OK;
  };
  int grid = 1, rc;
  rc = j3dg_stage_begin(ctx, 0);
  if (rc != J3DG_OK) return rc;
  // ---- primary rays ----
  tp.pool_ctr = ctr(3); tp.hard_count = ctr(6); tp.hard_taken = ctr(7); tp.done_blocks = ctr(11); tp.started = ctr(11) + 1;
  if (stats) {
    if ((rc = launch_hybrid(cast_kernel<PRIMARY, true>, cast_smem_bytes<PRIMARY, true>(), ntiles)) != J3DG_OK) return rc;
  } else if (ctx->cast_algo == 1) {  // 8-lanes-per-ray kernel only (A/B testing)
    if ((rc = persistent_grid(ctx, group_kernel<PRIMARY>, ntiles, &grid)) != J3DG_OK) return rc;
    group_kernel<PRIMARY><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
  } else {
    if (use_top) rc = launch_hybrid(cast_kernel<PRIMARY, false, true>, cast_smem_bytes<PRIMARY, false>(), ntiles);
    else rc = launch_hybrid(cast_kernel<PRIMARY, false>, cast_smem_bytes<PRIMARY, false>(), ntiles);
    if (rc != J3DG_OK) return rc;
  }
  if (used && !stats) {
    const uint32_t warps = 256 / 32;
    resolve_kernel<<<(uint32_t)((ntiles + warps - 1) / warps), 256, 0, ctx->stream>>>(ctx->d_meshes, tp.vw, x0, y0, x1, y1, tp.grid, d_pixels, stride,
                                                                                     (float4*)tp.shadow_pos, (uint32_t*)tp.shadow_pix, ctx->d_stats);
    KERNEL_CHECK(ctx);
  }
  // ---- shadow rays of the hit pixels ----
  if (shadows) {
    const long long pools = ((long long)npx + 31) / 32;
    tp.pool_ctr = ctr(8); tp.hard_count = ctr(9); tp.hard_taken = ctr(10); tp.done_blocks = ctr(12); tp.started = ctr(12) + 1;
    tp.budget = std::min<uint32_t>(ctx->shadow_budget, 255u);  // any-hit rays that start on the surface: their long ones are long from the start
    if (ctx->cast_algo == 1) {
      if ((rc = persistent_grid(ctx, group_kernel<SHADOW>, pools, &grid)) != J3DG_OK) return rc;
      group_kernel<SHADOW><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
      KERNEL_CHECK(ctx);
    } else {
      if (use_top) rc = launch_hybrid(cast_kernel<SHADOW, false, true>, cast_smem_bytes<SHADOW, false>(), pools);
      else rc = launch_hybrid(cast_kernel<SHADOW, false>, cast_smem_bytes<SHADOW, false>(), pools);
      if (rc != J3DG_OK) return rc;
    }
  }
  rc = j3dg_stage_end(ctx, 0);
  if (rc != J3DG_OK) return rc;
  if (!sharded) {
    ctx->rays_primary += (uint64_t)rw * rh;  // shadow rays (d_stats[4]) are added when the timings are read
  } else {  // own bands + one halo row above each (except the first band of the rectangle)
    const int bands = (rh + J3DG_SHARD_BAND_ROWS - 1) / J3DG_SHARD_BAND_ROWS;
    for (int b = (int)ctx->shard_rank; b < bands; b += (int)ctx->shard_world)
      ctx->rays_primary += (uint64_t)rw * (std::min(rh, (b + 1) * J3DG_SHARD_BAND_ROWS) - b * J3DG_SHARD_BAND_ROWS + (b ? 1 : 0));
  }
#ifdef J3DG_TIMELINE
  if (!stats) {
    unsigned long long t[20];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(t, ctx->d_stats, sizeof(t), cudaMemcpyDeviceToHost);
    fprintf(stderr, "timeline: lane phase ends by %.0f us, kernel ends %.0f us; per block mean lane %.0f us, mean drain %.0f us; hard rays %llu\n",
            (t[14] - t[13]) * 1e-3, (t[15] - t[13]) * 1e-3, t[16] * 1e-3 / t[18], t[17] * 1e-3 / t[18], t[6]);
  }
#endif
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
  tp.sticky = ctx->d_status;
  tp.rays = d_rays; tp.hits = d_hits; tp.ids = d_ids; tp.nrays = n;
  tp.pool_ctr = reinterpret_cast<unsigned int*>(ctx->d_stats + 3);
  int grid = 1;
  int rc = persistent_grid(ctx, group_kernel<RAYLIST>, ((long long)n + 31) / 32, &grid);
  if (rc != J3DG_OK) return rc;
  group_kernel<RAYLIST><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}
