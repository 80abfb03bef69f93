// cast.cu — per-pixel ray cast.  Replaces canvas::update_canvas (j3d/canvas.cpp:677-874):
// ray generation (773-782), qbvh_two_level_with_transformations::find_closest_triangle
// (jtk/qbvh.h:3303-3387) -> qbvh::find_closest_triangle (1701-1852) with the Woop test
// (4793-4869), hit -> pixel record (788-834) and the shadow ray (836-857).
//
// Mapping onto the GPU ("one ray per 8 lanes"): the BVH is 8 wide and a leaf holds at most 8
// triangles, so a ray is traversed by a GROUP of 8 lanes — lane c tests child c of the current
// node (or triangle c of the current leaf), the group votes (ballot), picks the nearest child
// with three shuffle-min steps and pushes the other hit children onto ONE stack per ray in
// shared memory.  A warp carries four rays.  Compared with one ray per lane this keeps every
// lane of a node / leaf test busy on the same cache line, makes one traversal step ~8x shorter
// (the kernel used to be bound by the serial latency of its slowest 32-ray tiles) and moves the
// whole stack into shared memory.  Warps are persistent: each pulls 8x4-pixel tiles from a global
// counter and its four groups pull rays from the tile one by one, so a group whose ray ends
// early (background, silhouette) immediately starts the next ray.
//
// Pipeline of one j3dg_cast: trace<PRIMARY> (raw hit: t, u, v, record slot) -> resolve_kernel
// (hit -> pixel record; appends shadow rays) -> trace<SHADOW> over the appended list (any hit).
//
// Parity rules: the ray, the Woop edge functions, t/u/v, the triangle normal and
// its two transforms are evaluated with separately rounded mul/add/sub (no FMA), in the
// reference's operation order.  Only the box tests use FMA — they are conservative and do not
// influence which triangle is the closest hit.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int GROUP = 8;                                  // lanes per ray = children per node = max triangles per leaf
constexpr int BLOCK_THREADS = 128;
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
__device__ __forceinline__ float group_min(uint32_t gmask, float v) {
  v = fminf(v, __shfl_xor_sync(gmask, v, 1));
  v = fminf(v, __shfl_xor_sync(gmask, v, 2));
  v = fminf(v, __shfl_xor_sync(gmask, v, 4));
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

// MODE PRIMARY: closest hit per pixel, raw result into the pixel buffer.
// MODE SHADOW : any hit (first accepted triangle ends the ray), sets mark bit 0.
// MODE RAYLIST: qbvh::find_closest_triangle semantics for arbitrary (also negative) t ranges:
//               closest = smallest |t|, bounds shrink on the side of the hit (qbvh.h:1812-1823).
template <int MODE, bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS, J3DG_CAST_MIN_BLOCKS) trace_kernel(const TraceParams p) {
  constexpr bool ANY_HIT = MODE == SHADOW;
  constexpr bool GENERAL = MODE == RAYLIST;
  __shared__ uint2 s_stack[STACK_SIZE * GROUPS_PER_BLOCK];
  const int lane = threadIdx.x & 31;
  const int c = lane & 7;                               // my child / triangle slot
  const int gshift = lane & 24;                         // first lane of my group
  const uint32_t gmask = 0xFFu << gshift;
  uint2* const stk = s_stack + (threadIdx.x >> 3);      // entry i at stk[i * GROUPS_PER_BLOCK]
  const uint32_t below = (1u << c) - 1u;
  uint32_t* const overflow_flag = reinterpret_cast<uint32_t*>(p.stats + 2);

  // ---- number of ray slots ----
  uint32_t total_pools;   // a pool = 32 consecutive slots
  uint32_t supers_x = 1;
  uint32_t list_n = 0;
  if (MODE == PRIMARY) {
    const uint32_t tiles_x = (uint32_t)(p.x1 - p.x0 + TILE_W) / TILE_W, tiles_y = (uint32_t)(p.y1 - p.y0 + TILE_H) / TILE_H;
    supers_x = (tiles_x + SUPER_W - 1) / SUPER_W;
    total_pools = supers_x * ((tiles_y + SUPER_H - 1) / SUPER_H) * (SUPER_W * SUPER_H);
  } else {
    list_n = MODE == SHADOW ? (uint32_t)p.stats[5] : p.nrays;
    total_pools = (list_n + 31u) / 32u;
    if (MODE == SHADOW && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 4, (unsigned long long)list_n);
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
  uint32_t pn = 0, pt = 0;

  // ---- warp-uniform pool state ----
  uint32_t pool_next = 32, pool_id = 0;
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
            uint4 lo = make_uint4(0u, 0u, 0u, __float_as_uint(found ? best_t : FLT_MAX));
            if (STATS) { lo.y = pn; lo.z = pt; }  // the counting pass returns per-pixel costs in the u / v slots
            dst[0] = lo;
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
        if (STATS && c == 0) {
          atomicAdd(p.stats + 0, (unsigned long long)pn);
          atomicAdd(p.stats + 1, (unsigned long long)pt);
        }
        have_ray = false;
      }
    }
    // groups without a ray take the next slots of the warp's pool (warp-uniform loop)
    uint32_t need = __ballot_sync(0xffffffffu, !have_ray) & 0x01010101u;
    while (need && !exhausted) {
      if (pool_next >= 32u) {
        uint32_t id = 0;
        if (lane == 0) id = atomicAdd(reinterpret_cast<unsigned int*>(p.stats + 3), 1u);
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
          best_t = FLT_MAX; best_u = 0.f; best_v = 0.f; best_slot = 0xFFFFFFFFu; best_mesh = 0;
          mesh_k = 0;
          pn = 0; pt = 0;
          enter_mesh(wr, 0);
          have_ray = true;
        }
      }
      need = __ballot_sync(0xffffffffu, !have_ray) & 0x01010101u;
    }
    if (exhausted && !__any_sync(0xffffffffu, have_ray)) break;

    // =========================== (B) inner node: lane c tests child c ===========================
    if (have_ray && !(cur & J3DG_LEAF_BIT)) {
      const char* np = reinterpret_cast<const char*>(nodes + cur);
      const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(np));          // ox oy oz | nchild      (broadcast)
      const uint4 h1 = __ldg(reinterpret_cast<const uint4*>(np) + 1);      // sx sy sz | pad         (broadcast)
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(np + 32) + c);  // my child's quantised box
      const uint32_t ref = __ldg(reinterpret_cast<const uint32_t*>(np + 96) + c);
      if (STATS) ++pn;
      const float sx = __uint_as_float(h1.x) * idx, sy = __uint_as_float(h1.y) * idy, sz = __uint_as_float(h1.z) * idz;
      const float bx = (__uint_as_float(h0.x) - ox) * idx;
      const float by = (__uint_as_float(h0.y) - oy) * idy;
      const float bz = (__uint_as_float(h0.z) - oz) * idz;
      // byte -> float without the quarter-rate I2F: PRMT builds the bits of 2^23 + q, the subtraction is exact
      const float qnx = __fsub_rn(__uint_as_float(__byte_perm(q.x, q.y, sel_nx)), 8388608.f);
      const float qfx = __fsub_rn(__uint_as_float(__byte_perm(q.x, q.y, sel_fx)), 8388608.f);
      const float qny = __fsub_rn(__uint_as_float(__byte_perm(q.x, q.y, sel_ny)), 8388608.f);
      const float qfy = __fsub_rn(__uint_as_float(__byte_perm(q.x, q.y, sel_fy)), 8388608.f);
      const float qnz = __fsub_rn(__uint_as_float(__byte_perm(q.x, q.y, sel_nz)), 8388608.f);
      const float qfz = __fsub_rn(__uint_as_float(__byte_perm(q.x, q.y, sel_fz)), 8388608.f);
      float tmin = fmaxf(fmaxf(fmaf(qnx, sx, bx), fmaf(qny, sy, by)), fmaxf(fmaf(qnz, sz, bz), t_near));
      float tmax = fminf(fminf(fmaf(qfx, sx, bx), fmaf(qfy, sy, by)), fminf(fmaf(qfz, sz, bz), t_far));
      // conservative padding against rounding of the slab arithmetic
      tmin = fmaf(-fabsf(tmin), 2e-6f, tmin);
      tmax = fmaf(fabsf(tmax), 2e-6f, tmax);
      const bool hit = tmin <= tmax;  // empty slots have inverted boxes and never pass
      const uint32_t hm = (__ballot_sync(gmask, hit) >> gshift) & 0xFFu;
      if (hm == 0u) {
        cur = pop();
      } else {
        const float key = hit ? tmin : FLT_MAX;
        const float nearest = group_min(gmask, key);
        const uint32_t nm8 = (__ballot_sync(gmask, hit && key == nearest) >> gshift) & 0xFFu;
        const int near_lane = __ffs(nm8) - 1;
        const uint32_t others = hm & ~(1u << near_lane);
        if (hit && c != near_lane) {
          const int pos = sp + __popc(others & below);
          if (pos < STACK_SIZE) stk[pos * GROUPS_PER_BLOCK] = make_uint2(ref, __float_as_uint(tmin));
          else *overflow_flag = 1u;
        }
        sp = min(sp + __popc(others), STACK_SIZE);
        cur = __shfl_sync(gmask, ref, gshift + near_lane);
        __syncwarp(gmask);  // the pushes must be visible to whichever lane pops them
      }
    }

    // =========================== (C) leaf: lane c tests triangle c ===========================
    if (have_ray && (cur & J3DG_LEAF_BIT) && cur != J3DG_EMPTY_CHILD) {
      const uint32_t first = cur & J3DG_LEAF_FIRST_MASK;
      const float4* tp = reinterpret_cast<const float4*>(tris + first + c);
      const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
      const uint32_t lm = (__ballot_sync(gmask, __float_as_uint(v1.w) != 0u) >> gshift) & 0xFFu;
      const int last = lm ? __ffs(lm) - 1 : GROUP - 1;
      bool hit = c <= last;
      if (STATS) pt += (uint32_t)(last + 1);
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
      const uint32_t anyhit = __ballot_sync(gmask, hit) & gmask;
      bool done = false;
      if (anyhit) {
        const float nearest = group_min(gmask, key);
        const uint32_t wm = (__ballot_sync(gmask, hit && key == nearest) >> gshift) & 0xFFu;
        const int win = __ffs(wm) - 1;  // lowest slot wins ties, like the sequential strict-less update
        const float wt = __shfl_sync(gmask, t, gshift + win);
        const float wu = __shfl_sync(gmask, u, gshift + win);
        const float wv = __shfl_sync(gmask, v, gshift + win);
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
  if (used) CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_meshes, host.data(), sizeof(MeshDev) * used, cudaMemcpyHostToDevice, ctx->stream));
  const bool shadows = (view->flags & J3DG_SHADOW) && used && !stats;
  const int rw = x1 - x0 + 1, rh = y1 - y0 + 1;
  const size_t shadow_pix_off = ((size_t)rw * rh * sizeof(float4) + 255) & ~(size_t)255;
  if (shadows) {  // one shadow ray per hit pixel at most
    int rc = j3dg_reserve(ctx, &ctx->d_shadow, &ctx->shadow_cap, shadow_pix_off + (size_t)rw * rh * sizeof(uint32_t));
    if (rc != J3DG_OK) return rc;
  }
  // [0..3] per-launch counters, [4] accumulates until the timings are reset, [5] shadow list length
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 5, 0, sizeof(unsigned long long), ctx->stream));
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
  const long long ntiles = (long long)((rw + TILE_W - 1) / TILE_W) * ((rh + TILE_H - 1) / TILE_H);
  int grid = 1;
  { int rc = j3dg_stage_begin(ctx, 0); if (rc != J3DG_OK) return rc; }
  if (stats) {
    int rc = persistent_grid(ctx, trace_kernel<PRIMARY, true>, ntiles, &grid);
    if (rc != J3DG_OK) return rc;
    trace_kernel<PRIMARY, true><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
  } else {
    int rc = persistent_grid(ctx, trace_kernel<PRIMARY, false>, ntiles, &grid);
    if (rc != J3DG_OK) return rc;
    trace_kernel<PRIMARY, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
  }
  KERNEL_CHECK(ctx);
  if (used && !stats) {
    const uint32_t warps = 256 / 32;
    resolve_kernel<<<(uint32_t)((ntiles + warps - 1) / warps), 256, 0, ctx->stream>>>(ctx->d_meshes, tp.vw, x0, y0, x1, y1, d_pixels, stride,
                                                                                     (float4*)tp.shadow_pos, (uint32_t*)tp.shadow_pix, ctx->d_stats);
    KERNEL_CHECK(ctx);
  }
  if (shadows) {
    CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats + 3, 0, sizeof(unsigned long long), ctx->stream));  // pool counter of the second launch
    int rc = persistent_grid(ctx, trace_kernel<SHADOW, false>, ((long long)rw * rh + 31) / 32, &grid);
    if (rc != J3DG_OK) return rc;
    trace_kernel<SHADOW, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
    KERNEL_CHECK(ctx);
  }
  { int rc = j3dg_stage_end(ctx, 0); if (rc != J3DG_OK) return rc; }
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
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  TraceParams tp = {};
  tp.meshes = ctx->d_meshes;
  tp.nm = (d.nodes && d.nt) ? 1u : 0u;
  tp.stats = ctx->d_stats;
  tp.rays = d_rays; tp.hits = d_hits; tp.ids = d_ids; tp.nrays = n;
  int grid = 1;
  int rc = persistent_grid(ctx, trace_kernel<RAYLIST, false>, ((long long)n + 31) / 32, &grid);
  if (rc != J3DG_OK) return rc;
  trace_kernel<RAYLIST, false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(tp);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}
