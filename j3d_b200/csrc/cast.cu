// cast.cu — per-pixel ray cast.  Replaces canvas::update_canvas (j3d/canvas.cpp:677-874):
// ray generation (773-782), qbvh_two_level_with_transformations::find_closest_triangle
// (jtk/qbvh.h:3303-3387) -> qbvh::find_closest_triangle (1701-1852) with the Woop test
// (4793-4869), hit -> pixel record (788-834) and the shadow ray (836-857).
//
// Parity rules: the ray, the Woop edge functions, t/u/v, the triangle normal and
// its two transforms are evaluated with separately rounded mul/add/sub (no FMA), in the
// reference's operation order.  Only the box tests use FMA — they are conservative and do not
// influence which triangle is the closest hit.
#include "common.cuh"

#include <algorithm>
#include <type_traits>

namespace {

constexpr int STACK_SIZE = 96;               // total traversal stack entries per ray
constexpr int SM_STACK = 12;                 // of which the first SM_STACK live in shared memory (rest: local memory, rarely touched)
constexpr int TILE_W = 8, TILE_H = 4;        // one warp = 8x4 pixels
constexpr int SUPER_W = 4, SUPER_H = 8;      // tiles are handed out super-tile by super-tile (32x32 pixels), row-major inside
constexpr int BLOCK_THREADS = 128;           // 4 warps; every warp fetches its own tiles (persistent threads)
#ifndef J3DG_CAST_MIN_BLOCKS
#define J3DG_CAST_MIN_BLOCKS 6
#endif

struct RayPre {  // intersect_woop_precompute, qbvh.h:4793-4823
  int kx, ky, kz;
  float Sx, Sy, Sz;
};

__device__ __forceinline__ float pick(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ RayPre woop_precompute(float dx, float dy, float dz) {
  RayPre o;
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  o.kz = 2;
  if (ax > ay) { if (ax > az) o.kz = 0; }
  else { if (ay > az) o.kz = 1; }
  o.kx = o.kz == 2 ? 0 : o.kz + 1;
  o.ky = o.kx == 2 ? 0 : o.kx + 1;
  const float dkz = pick(dx, dy, dz, o.kz);
  if (dkz < 0.f) { const int t = o.kx; o.kx = o.ky; o.ky = t; }
  o.Sz = fdiv(1.f, dkz);
  o.Sx = fmul(pick(dx, dy, dz, o.kx), o.Sz);
  o.Sy = fmul(pick(dx, dy, dz, o.ky), o.Sz);
  return o;
}

struct Best {
  float t, u, v;
  uint32_t slot;   // index of the triangle record
  uint32_t mesh;
  bool found;
};

// One triangle, one lane of intersect_woop (qbvh.h:4825-4869).  Returns true when the triangle
// is hit inside (t_near, t_far); the reference's reciprocal(det) (rcpps + 1 NR step, ~2e-7) is
// replaced by the correctly rounded 1/det.
__device__ __forceinline__ bool woop_test(const float4 v0, const float4 v1, const float4 v2, const RayPre& p,
                                          float ox, float oy, float oz, float t_near, float t_far, float& t, float& u, float& v) {
  const float Ax_ = fsub(v0.x, ox), Ay_ = fsub(v0.y, oy), Az_ = fsub(v0.z, oz);
  const float Bx_ = fsub(v1.x, ox), By_ = fsub(v1.y, oy), Bz_ = fsub(v1.z, oz);
  const float Cx_ = fsub(v2.x, ox), Cy_ = fsub(v2.y, oy), Cz_ = fsub(v2.z, oz);
  const float Akz = pick(Ax_, Ay_, Az_, p.kz), Bkz = pick(Bx_, By_, Bz_, p.kz), Ckz = pick(Cx_, Cy_, Cz_, p.kz);
  const float Ax = fsub(pick(Ax_, Ay_, Az_, p.kx), fmul(p.Sx, Akz));
  const float Ay = fsub(pick(Ax_, Ay_, Az_, p.ky), fmul(p.Sy, Akz));
  const float Bx = fsub(pick(Bx_, By_, Bz_, p.kx), fmul(p.Sx, Bkz));
  const float By = fsub(pick(Bx_, By_, Bz_, p.ky), fmul(p.Sy, Bkz));
  const float Cx = fsub(pick(Cx_, Cy_, Cz_, p.kx), fmul(p.Sx, Ckz));
  const float Cy = fsub(pick(Cx_, Cy_, Cz_, p.ky), fmul(p.Sy, Ckz));
  const float U = fsub(fmul(Cx, By), fmul(Cy, Bx));
  const float V = fsub(fmul(Ax, Cy), fmul(Ay, Cx));
  const float W = fsub(fmul(Bx, Ay), fmul(By, Ax));
  const bool inside = ((U <= 0.f) && (V <= 0.f) && (W <= 0.f)) || ((U >= 0.f) && (V >= 0.f) && (W >= 0.f));
  if (!inside) return false;
  const float det = fadd(fadd(U, V), W);
  if (!(det != 0.f)) return false;
  const float inv_det = fdiv(1.f, det);
  const float Az = fmul(p.Sz, Akz), Bz = fmul(p.Sz, Bkz), Cz = fmul(p.Sz, Ckz);
  const float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
  t = fmul(T, inv_det);
  if (!((t_far > t) && (t > t_near))) return false;
  u = fmul(V, inv_det);
  v = fmul(W, inv_det);
  return true;
}

__device__ __forceinline__ float safe_rcp(float d) {
  // box tests only: keep the reciprocal finite so 0 * inf never produces NaN
  const float big = 1e18f;
  if (fabsf(d) < 1e-18f) return d < 0.f ? -big : big;
  return 1.f / d;
}

// Byte i of w as a float without the quarter-rate I2F (XU pipe): PRMT builds the bit pattern of
// 2^23 + q, the subtraction is exact.  Both instructions run on the full-rate ALU / FMA pipes.
template <int I>
__device__ __forceinline__ float ubyte(uint32_t w) {
  return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | (uint32_t)I)), 8388608.f);
}

// Per-thread traversal stack: the first SM_STACK entries in shared memory (column tid of a
// [SM_STACK][BLOCK_THREADS] array: conflict-free), deeper entries in local memory.
struct Stack {
  uint2* sm;                          // &s_stack[threadIdx.x]
  uint2 deep[STACK_SIZE - SM_STACK];
  int sp;
  __device__ __forceinline__ void push(uint32_t ref, float t, uint32_t* overflow_flag) {
    const uint2 e = make_uint2(ref, __float_as_uint(t));
    if (sp < SM_STACK) sm[sp * BLOCK_THREADS] = e;
    else if (sp < STACK_SIZE) deep[sp - SM_STACK] = e;
    else { *overflow_flag = 1u; return; }
    ++sp;
  }
  __device__ __forceinline__ uint2 pop() {
    --sp;
    return sp < SM_STACK ? sm[sp * BLOCK_THREADS] : deep[sp - SM_STACK];
  }
};

// Traverses one mesh.  GENERAL = qbvh semantics for arbitrary (also negative) t ranges:
// closest = smallest |t|, bounds shrink on the side of the hit (qbvh.h:1812-1823).
// ANY_HIT returns at the first accepted triangle (shadow rays only need `found`).
// Control flow is "while-while": all lanes of a warp first descend inner nodes until each holds
// a leaf (or is done), then all test their leaves — the two phases reconverge separately, so a
// lane testing triangles never serialises against a lane decoding a node.
template <bool ANY_HIT, bool GENERAL, bool STATS>
__device__ __forceinline__ void traverse_mesh(const MeshDev& m, uint32_t mesh_index, float ox, float oy, float oz,
                                              float dx, float dy, float dz, float& t_near, float& t_far, Best& best,
                                              uint32_t& stat_nodes, uint32_t& stat_tris, uint32_t* overflow_flag, Stack& stk) {
  if (m.nt == 0) return;
  const RayPre pre = woop_precompute(dx, dy, dz);
  const float idx = safe_rcp(dx), idy = safe_rcp(dy), idz = safe_rcp(dz);
  const bool negx = idx < 0.f, negy = idy < 0.f, negz = idz < 0.f;

  stk.sp = 0;
  uint32_t cur = 0;  // root node; J3DG_EMPTY_CHILD (which has the leaf bit set) = nothing left
  const WideNode* __restrict__ nodes = m.nodes;
  const TriRec* __restrict__ tris = m.tris;

  auto pop = [&]() -> uint32_t {
    while (stk.sp > 0) {
      const uint2 e = stk.pop();
      // entry points of popped boxes that now lie beyond the shrunk interval are skipped
      if (GENERAL || __uint_as_float(e.y) <= t_far) return e.x;
    }
    return J3DG_EMPTY_CHILD;
  };

  for (;;) {
    // ---- phase 1: inner nodes, 8 quantised child boxes each ----
    while (!(cur & J3DG_LEAF_BIT)) {
      const uint4* np = reinterpret_cast<const uint4*>(nodes + cur);
      const uint4 h0 = __ldg(np + 0);  // ox oy oz | ex ey ez n
      const uint4 q0 = __ldg(np + 1);  // qlo x[0..7] | qlo y[0..7]
      const uint4 q1 = __ldg(np + 2);  // qlo z[0..7] | qhi x[0..7]
      const uint4 q2 = __ldg(np + 3);  // qhi y[0..7] | qhi z[0..7]
      const uint4 c0 = __ldg(np + 4);
      const uint4 c1 = __ldg(np + 5);
      if (STATS) ++stat_nodes;
      const float sx = __uint_as_float(__byte_perm(h0.w, 0u, 0x4440u) << 23) * idx;
      const float sy = __uint_as_float(__byte_perm(h0.w, 0u, 0x4441u) << 23) * idy;
      const float sz = __uint_as_float(__byte_perm(h0.w, 0u, 0x4442u) << 23) * idz;
      const float bx = (__uint_as_float(h0.x) - ox) * idx;
      const float by = (__uint_as_float(h0.y) - oy) * idy;
      const float bz = (__uint_as_float(h0.z) - oz) * idz;
      // near / far plane words per axis (two 32-bit words = 8 children)
      const uint32_t nx0 = negx ? q1.z : q0.x, nx1 = negx ? q1.w : q0.y;
      const uint32_t fx0 = negx ? q0.x : q1.z, fx1 = negx ? q0.y : q1.w;
      const uint32_t ny0 = negy ? q2.x : q0.z, ny1 = negy ? q2.y : q0.w;
      const uint32_t fy0 = negy ? q0.z : q2.x, fy1 = negy ? q0.w : q2.y;
      const uint32_t nz0 = negz ? q2.z : q1.x, nz1 = negz ? q2.w : q1.y;
      const uint32_t fz0 = negz ? q1.x : q2.z, fz1 = negz ? q1.y : q2.w;
      uint32_t near_ref = J3DG_EMPTY_CHILD;
      float near_t = FLT_MAX;
      auto test_child = [&](auto IC, uint32_t wnx, uint32_t wfx, uint32_t wny, uint32_t wfy, uint32_t wnz, uint32_t wfz, uint32_t ref) {
        constexpr int B = decltype(IC)::value;
        const float tlx = fmaf(ubyte<B>(wnx), sx, bx);
        const float thx = fmaf(ubyte<B>(wfx), sx, bx);
        const float tly = fmaf(ubyte<B>(wny), sy, by);
        const float thy = fmaf(ubyte<B>(wfy), sy, by);
        const float tlz = fmaf(ubyte<B>(wnz), sz, bz);
        const float thz = fmaf(ubyte<B>(wfz), sz, bz);
        float tmin = fmaxf(fmaxf(tlx, tly), fmaxf(tlz, t_near));
        float tmax = fminf(fminf(thx, thy), fminf(thz, t_far));
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
      test_child(std::integral_constant<int, 0>{}, nx0, fx0, ny0, fy0, nz0, fz0, c0.x);
      test_child(std::integral_constant<int, 1>{}, nx0, fx0, ny0, fy0, nz0, fz0, c0.y);
      test_child(std::integral_constant<int, 2>{}, nx0, fx0, ny0, fy0, nz0, fz0, c0.z);
      test_child(std::integral_constant<int, 3>{}, nx0, fx0, ny0, fy0, nz0, fz0, c0.w);
      test_child(std::integral_constant<int, 0>{}, nx1, fx1, ny1, fy1, nz1, fz1, c1.x);
      test_child(std::integral_constant<int, 1>{}, nx1, fx1, ny1, fy1, nz1, fz1, c1.y);
      test_child(std::integral_constant<int, 2>{}, nx1, fx1, ny1, fy1, nz1, fz1, c1.z);
      test_child(std::integral_constant<int, 3>{}, nx1, fx1, ny1, fy1, nz1, fz1, c1.w);
      cur = (near_ref != J3DG_EMPTY_CHILD) ? near_ref : pop();
    }
    if (cur == J3DG_EMPTY_CHILD) return;
    // ---- phase 2: leaves, 1..4 consecutive triangle records each ----
    do {
      const uint32_t first = cur & J3DG_LEAF_FIRST_MASK;
      const uint32_t cnt = ((cur >> 29) & 3u) + 1u;
      for (uint32_t k = 0; k < cnt; ++k) {
        const float4* tp = reinterpret_cast<const float4*>(tris + first + k);
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        if (STATS) ++stat_tris;
        float t, u, v;
        if (woop_test(v0, v1, v2, pre, ox, oy, oz, t_near, t_far, t, u, v)) {
          const bool closer = GENERAL ? (fabsf(t) < fabsf(best.t)) : (t < best.t);
          if (closer) {
            best.found = true; best.t = t; best.u = u; best.v = v; best.slot = first + k; best.mesh = mesh_index;
            if (!GENERAL || t > 0.f) t_far = t; else t_near = t;
            if (ANY_HIT) return;
          }
        }
      }
      cur = pop();
    } while (cur != J3DG_EMPTY_CHILD && (cur & J3DG_LEAF_BIT));
    if (cur == J3DG_EMPTY_CHILD) return;
  }
}

template <bool ANY_HIT, bool STATS>
__device__ __forceinline__ void trace_scene(const MeshDev* __restrict__ meshes, uint32_t nm, float4 org, float4 dir,
                                            float t_near, float t_far, Best& best, uint32_t& sn, uint32_t& st, uint32_t* ovf, Stack& stk) {
  best.found = false;
  best.t = FLT_MAX;
  for (uint32_t o = 0; o < nm; ++o) {
    const MeshDev& m = meshes[o];
    // qbvh.h:3358-3359: the ray is taken into object space by the inverted object matrix
    const float4 d2 = mat_vec(m.cs_inv, dir);
    const float4 o2 = mat_vec(m.cs_inv, org);
    traverse_mesh<ANY_HIT, false, STATS>(m, o, o2.x, o2.y, o2.z, d2.x, d2.y, d2.z, t_near, t_far, best, sn, st, ovf, stk);
    if (ANY_HIT && best.found) return;
  }
}

__device__ __forceinline__ float4 transform_point(const float* __restrict__ m, float4 p) {  // jtk::transform, qbvh.h:5140-5151
  float4 r = mat_vec(m, p);
  if (r.w != 1.f && r.w != 0.f) { r.x = fdiv(r.x, r.w); r.y = fdiv(r.y, r.w); r.z = fdiv(r.z, r.w); r.w = 1.f; }
  return r;
}

// stats layout (u64 each): [0] node visits, [1] triangle tests, [2] stack-overflow flag, [3] tile counter of the
// running launch, [4] hit pixels of shadowed frames (accumulates until the timings are reset).
//
// Persistent threads: the grid is sized to fill the machine once (SMs x resident blocks); every WARP pulls the
// next 8x4-pixel tile from a global counter, so a warp whose rays finish early (background, silhouette) never
// waits for the slowest warp of its block, and there is no tail of half-empty blocks.  Tiles are numbered
// super-tile by super-tile (32x32 pixels) so that warps running at the same time work on neighbouring pixels and
// share the upper BVH levels in L1/L2.
template <bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS, J3DG_CAST_MIN_BLOCKS) cast_kernel(const MeshDev* __restrict__ meshes, uint32_t nm, ViewDev vw,
                                                              int x0, int y0, int x1, int y1, j3dg_pixel* __restrict__ out,
                                                              uint32_t stride, unsigned long long* __restrict__ stats) {
  __shared__ uint2 s_stack[SM_STACK * BLOCK_THREADS];
  Stack stk;
  stk.sm = s_stack + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles_x = (uint32_t)(x1 - x0 + TILE_W) / TILE_W, tiles_y = (uint32_t)(y1 - y0 + TILE_H) / TILE_H;
  const uint32_t supers_x = (tiles_x + SUPER_W - 1) / SUPER_W, supers_y = (tiles_y + SUPER_H - 1) / SUPER_H;
  const uint32_t total = supers_x * supers_y * (SUPER_W * SUPER_H);
  uint32_t sn = 0, st = 0;
  for (;;) {
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(reinterpret_cast<unsigned int*>(stats + 3), 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= total) break;
    const uint32_t sup = tile / (SUPER_W * SUPER_H), in = tile % (SUPER_W * SUPER_H);
    const uint32_t tx = (sup % supers_x) * SUPER_W + (in % SUPER_W), ty = (sup / supers_x) * SUPER_H + (in / SUPER_W);
    const int x = x0 + (int)tx * TILE_W + (lane & (TILE_W - 1));
    const int y = y0 + (int)ty * TILE_H + (lane / TILE_W);
    if (x > x1 || y > y1) continue;  // warp-uniform for whole tiles outside; lanes outside just skip
    unsigned long long tile_t0 = 0;
    if (STATS) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tile_t0));
    // canvas.cpp:773-776
    const float w = (float)vw.width, h = (float)vw.height;
    float4 sp;
    sp.x = fsub(fmul(2.f, fdiv(fadd((float)x, 0.5f), w)), 1.f);
    sp.y = fsub(fmul(2.f, fdiv(fadd((float)y, 0.5f), h)), 1.f);
    sp.z = vw.near_plane;
    sp.w = 1.f;
    float4 dir = mat_vec(vw.pinv, sp);
    dir.w = 0.f;
    dir = mat_vec(vw.cs, dir);
    const float4 org = make_float4(vw.origin[0], vw.origin[1], vw.origin[2], vw.origin[3]);
    Best best;
    uint32_t pn = 0, pt = 0;  // this pixel's node visits / triangle tests (STATS only)
    trace_scene<false, STATS>(meshes, nm, org, dir, fdiv(vw.diagonal, 100.f), FLT_MAX, best, pn, pt, (uint32_t*)(stats + 2), stk);
    sn += pn; st += pt;

    uint4 lo, hi;  // the 32-byte pixel record as two 16-byte stores
    if (best.found) {
      const MeshDev& m = meshes[best.mesh];
      const float4* tp = reinterpret_cast<const float4*>(m.tris + best.slot);
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
      const float bu = best.u, bv = best.v;
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
      if (vw.flags & J3DG_SHADOW) {  // canvas.cpp:836-857
        const float4 V0 = transform_point(m.cs, make_float4(v0.x, v0.y, v0.z, 1.f));
        const float4 V1 = transform_point(m.cs, make_float4(v1.x, v1.y, v1.z, 1.f));
        const float4 V2 = transform_point(m.cs, make_float4(v2.x, v2.y, v2.z, 1.f));
        float4 pos, ld;
        pos.x = fadd(fadd(fmul(V0.x, k), fmul(bu, V1.x)), fmul(bv, V2.x));
        pos.y = fadd(fadd(fmul(V0.y, k), fmul(bu, V1.y)), fmul(bv, V2.y));
        pos.z = fadd(fadd(fmul(V0.z, k), fmul(bu, V1.z)), fmul(bv, V2.z));
        pos.w = fadd(fadd(fmul(V0.w, k), fmul(bu, V1.w)), fmul(bv, V2.w));
        ld.x = fsub(vw.light[0], pos.x); ld.y = fsub(vw.light[1], pos.y);
        ld.z = fsub(vw.light[2], pos.z); ld.w = fsub(vw.light[3], pos.w);
        Best sh;
        uint32_t dn = 0, dt = 0;
        trace_scene<true, false>(meshes, nm, pos, ld, 1e-3f, FLT_MAX, sh, dn, dt, (uint32_t*)(stats + 2), stk);
        if (sh.found) mark |= 1u;
      }
      lo.x = mark | (r << 8) | (g << 16) | (b << 24);
      lo.y = __float_as_uint(n.x);
      lo.z = __float_as_uint(n.y);
      lo.w = __float_as_uint(best.t);
      hi.x = tri;
      hi.y = __float_as_uint(bu);
      hi.z = __float_as_uint(bv);
      hi.w = m.db_id;
    } else {  // canvas.cpp:859-866
      lo = make_uint4(0u, 0u, 0u, __float_as_uint(FLT_MAX));
      hi = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
    }
    if (STATS) {  // the counting pass returns per-pixel costs in the u / v slots, tile start / end time (ns) in the barycentric slots
      __syncwarp();
      unsigned long long tile_t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tile_t1));
      lo.y = pn; lo.z = pt; hi.y = (uint32_t)tile_t0; hi.z = (uint32_t)tile_t1;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)y * stride + x);
    dst[0] = lo;
    dst[1] = hi;
  }
  if (STATS) {
    for (int o = 16; o; o >>= 1) {
      sn += __shfl_xor_sync(0xffffffffu, sn, o);
      st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    if (lane == 0) {
      atomicAdd(stats + 0, (unsigned long long)sn);
      atomicAdd(stats + 1, (unsigned long long)st);
    }
  }
}

// Counts hit pixels (rays accounting for shadow rays) — tiny reduction over the db_id field.
__global__ void __launch_bounds__(256) count_hits_kernel(const j3dg_pixel* __restrict__ px, uint32_t stride, int x0, int y0, int w, int h,
                                                          unsigned long long* __restrict__ counter) {
  uint32_t c = 0;
  const size_t total = (size_t)w * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = x0 + (int)(i % w), y = y0 + (int)(i / w);
    c += px[(size_t)y * stride + x].object_id != 0xFFFFFFFFu;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, (unsigned long long)c);
}

__global__ void __launch_bounds__(BLOCK_THREADS) find_closest_kernel(MeshDev m, const float* __restrict__ rays, uint32_t n,
                                                            float* __restrict__ hits, uint32_t* __restrict__ ids, uint32_t* overflow) {
  __shared__ uint2 s_stack[SM_STACK * BLOCK_THREADS];
  Stack stk;
  stk.sm = s_stack + threadIdx.x;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = rays + 8 * (size_t)i;
  float t_near = r[6], t_far = r[7];
  Best best;
  best.found = false;
  best.t = FLT_MAX;
  uint32_t sn = 0, st = 0;
  traverse_mesh<false, true, false>(m, 0, r[0], r[1], r[2], r[3], r[4], r[5], t_near, t_far, best, sn, st, overflow, stk);
  float* h = hits + 4 * (size_t)i;
  h[0] = best.found ? best.u : 0.f;
  h[1] = best.found ? best.v : 0.f;
  h[2] = best.t;
  h[3] = best.found ? 1.f : 0.f;
  ids[i] = best.found ? __float_as_uint(m.tris[best.slot].v0.w) : 0xFFFFFFFFu;
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
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));  // [4] accumulates until reset
  ViewDev vd;
  j3dg_make_view_dev(view, vd);
  const int rw = x1 - x0 + 1, rh = y1 - y0 + 1;
  // persistent grid: fill every SM once, never more blocks than there are tiles
  static int blocks_per_sm[2] = {0, 0};
  if (!blocks_per_sm[stats ? 1 : 0]) {
    int nb = 0;
    cudaError_t e = stats ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cast_kernel<true>, BLOCK_THREADS, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cast_kernel<false>, BLOCK_THREADS, 0);
    if (e != cudaSuccess) return j3dg_cuda_fail(ctx, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor", __FILE__, __LINE__);
    blocks_per_sm[stats ? 1 : 0] = std::max(nb, 1);
  }
  const long long ntiles = (long long)((rw + TILE_W - 1) / TILE_W) * ((rh + TILE_H - 1) / TILE_H);
  const int warps_per_block = BLOCK_THREADS / 32;
  const int grid = (int)std::max<long long>(1, std::min<long long>((long long)ctx->sm_count * blocks_per_sm[stats ? 1 : 0],
                                                                   (ntiles + warps_per_block - 1) / warps_per_block));
  { int rc = j3dg_stage_begin(ctx, 0); if (rc != J3DG_OK) return rc; }
  if (stats)
    cast_kernel<true><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(ctx->d_meshes, used, vd, x0, y0, x1, y1, d_pixels, stride, ctx->d_stats);
  else
    cast_kernel<false><<<grid, BLOCK_THREADS, 0, ctx->stream>>>(ctx->d_meshes, used, vd, x0, y0, x1, y1, d_pixels, stride, ctx->d_stats);
  KERNEL_CHECK(ctx);
  { int rc = j3dg_stage_end(ctx, 0); if (rc != J3DG_OK) return rc; }
  if (ctx->profiling && (view->flags & J3DG_SHADOW)) {  // one shadow ray per hit pixel: count them
    count_hits_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_pixels, stride, x0, y0, rw, rh, ctx->d_stats + 4);
    KERNEL_CHECK(ctx);
  }
  ctx->rays_primary += (uint64_t)rw * rh;  // shadow rays (d_stats[4]) are added when the timings are read
  return J3DG_OK;
}

int j3dg_launch_find_closest(j3dg_mesh* m, const float* d_rays, uint32_t n, float* d_hits, uint32_t* d_ids) {
  j3dg_ctx* ctx = m->ctx;
  if (!n) return J3DG_OK;
  MeshDev d;
  fill_mesh_dev(m, d);
  CU_CHECK(ctx, cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  find_closest_kernel<<<(n + BLOCK_THREADS - 1) / BLOCK_THREADS, BLOCK_THREADS, 0, ctx->stream>>>(d, d_rays, n, d_hits, d_ids, (uint32_t*)(ctx->d_stats + 2));
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}
