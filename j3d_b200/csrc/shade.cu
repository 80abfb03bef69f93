// shade.cu — fused shading pass over the pixel buffer.  Replaces canvas::canvas_to_image
// (j3d/canvas.cpp:582-670) with its four modes — plain matcap / vertex colour (410-446), edges
// (590-649, compute_convex_cos_angle 287-315, get_angle_color 346-389), wireframe (449-496),
// one-bit (498-579) — plus the background copy of canvas::render_scene (893-898, 55-78).
// One thread per pixel; the right and up neighbours come through L1/L2.  The reference runs
// this serially on one core.
#include "common.cuh"

namespace {

struct ShadeParams {
  uint32_t w, h, flags;
  float near_plane;
  float pinv[16];
  const uint32_t* matcap;
  uint32_t mw, mh, mstride, cavity;
  uint32_t shard_rank, shard_world;  // screen sharding: only rows of this rank's bands are shaded
};

struct Px {  // the fields shading reads
  uint32_t mark, r, g, b;
  float u, v, depth;
  uint32_t object_id;
};

__device__ __forceinline__ Px load_px(const j3dg_pixel* __restrict__ p) {
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p));
  Px q;
  q.mark = lo.x & 0xffu; q.r = (lo.x >> 8) & 0xffu; q.g = (lo.x >> 16) & 0xffu; q.b = lo.x >> 24;
  q.u = __uint_as_float(lo.y); q.v = __uint_as_float(lo.z); q.depth = __uint_as_float(lo.w);
  q.object_id = __ldg(reinterpret_cast<const uint32_t*>(p) + 4);
  return q;
}

// (uint32_t)std::floor(x) on x86-64: cvttss2si (64-bit) then truncation to 32 bits
__device__ __forceinline__ uint32_t floor_to_u32(float x) { return (uint32_t)(long long)floorf(x); }

__device__ __forceinline__ uint32_t get_U(float u, uint32_t mw) {  // canvas.cpp:320-323
  return floor_to_u32(fadd(0.5f, fmul(fmul(fadd(u, 1.f), (float)(mw - 1)), 0.5f)));
}
__device__ __forceinline__ uint32_t get_V(float v, uint32_t mh) {  // canvas.cpp:325-328
  return floor_to_u32(fadd(0.5f, fmul(fmul(fadd(-v, 1.f), (float)(mh - 1)), 0.5f)));
}

__device__ __forceinline__ uint32_t get_color(const ShadeParams& s, uint32_t U, uint32_t V, uint32_t shadow) {  // canvas.cpp:330-344
  U = min(U, s.mw - 1);  // the reference indexes unchecked; u,v are unit-normal components so this never clamps
  V = min(V, s.mh - 1);
  uint32_t clr = __ldg(s.matcap + (size_t)V * s.mstride + U);
  if (shadow) {
    const uint32_t r = (clr & 0xffu) >> 2, g = ((clr >> 8) & 0xffu) >> 2, b = ((clr >> 16) & 0xffu) >> 2;
    clr = 0xff000000u | (b << 16) | (g << 8) | r;
  }
  return clr;
}

__device__ __forceinline__ uint32_t trunc_u32(float x) { return (uint32_t)(long long)x; }  // (uint32_t)float on x86-64

__device__ __forceinline__ uint32_t get_angle_color(const ShadeParams& s, float angle, float u, float v, uint32_t mark) {  // canvas.cpp:346-389
  uint32_t clr = get_color(s, get_U(u, s.mw), get_V(v, s.mh), mark);
  if (fabsf(angle) <= 1.f) {
    uint32_t r = clr & 0xffu, g = (clr >> 8) & 0xffu, b = (clr >> 16) & 0xffu;
    // the reference evaluates this expression in double (unqualified acos on a float) and rounds once
    const double half_pi = (double)1.57079632679489f;
    float scale = (float)((half_pi - fabs(acos((double)angle) - half_pi)) / half_pi);
    scale = fsqrt(fsub(1.f, scale));
    const uint32_t r2 = s.cavity & 0xffu, g2 = (s.cavity >> 8) & 0xffu, b2 = (s.cavity >> 16) & 0xffu;
    const float k = angle > 0.f ? 1.2f : 0.5f;  // concave : convex
    const float om = fsub(1.f, scale);
    r = trunc_u32(fadd(fmul((float)r, om), fmul(fmul((float)r2, scale), k)));
    g = trunc_u32(fadd(fmul((float)g, om), fmul(fmul((float)g2, scale), k)));
    b = trunc_u32(fadd(fmul((float)b, om), fmul(fmul((float)b2, scale), k)));
    r = min(r, 255u); g = min(g, 255u); b = min(b, 255u);
    clr = 0xff000000u | (b << 16) | (g << 8) | r;
  }
  return clr;
}

__device__ __forceinline__ float normal_z(float u, float v) { return fsqrt(fsub(fsub(1.f, fmul(u, u)), fmul(v, v))); }

__device__ __forceinline__ float convex_cos_angle(const ShadeParams& s, float x1, float y1, float u1, float v1, float depth1,
                                                  float x2, float y2, float u2, float v2, float depth2) {  // canvas.cpp:287-315
  const float w = (float)s.w, h = (float)s.h;
  float4 sp1 = make_float4(fsub(fmul(2.f, fdiv(fadd(x1, 0.5f), w)), 1.f), fsub(fmul(2.f, fdiv(fadd(y1, 0.5f), h)), 1.f), s.near_plane, 1.f);
  float4 d1 = mat_vec(s.pinv, sp1);
  float4 sp2 = make_float4(fsub(fmul(2.f, fdiv(fadd(x2, 0.5f), w)), 1.f), fsub(fmul(2.f, fdiv(fadd(y2, 0.5f), h)), 1.f), s.near_plane, 1.f);
  float4 d2 = mat_vec(s.pinv, sp2);
  const float p1x = fmul(depth1, d1.x), p1y = fmul(depth1, d1.y), p1z = fmul(depth1, d1.z);
  const float p2x = fmul(depth2, d2.x), p2y = fmul(depth2, d2.y), p2z = fmul(depth2, d2.z);
  const float n1z = normal_z(u1, v1), n2z = normal_z(u2, v2);
  const float d = fadd(fadd(fmul(u1, u2), fmul(v1, v2)), fmul(n1z, n2z));  // _mm_dp_ps 0x7f
  if ((double)fabsf(fsub(d, 1.f)) > 0.0001) {
    float px = fsub(p2x, p1x), py = fsub(p2y, p1y), pz = fsub(p2z, p1z);
    const float l = fsqrt(fadd(fadd(fmul(px, px), fmul(py, py)), fmul(pz, pz)));
    px = fdiv(px, l); py = fdiv(py, l); pz = fdiv(pz, l);
    return fadd(fadd(fmul(px, u1), fmul(py, v1)), fmul(pz, n1z));
  }
  return 0.f;
}

__device__ __forceinline__ uint32_t plain_color(const ShadeParams& s, const Px& p) {  // canvas::_get_color, canvas.cpp:410-446
  if (p.mark & 2u) {
    if (s.flags & J3DG_SHADING) {
      const float nz = normal_z(p.u, p.v);
      const float occ = (p.mark & 1u) ? 0.3f : 1.f;
      // dot((u,v,nz,0),(0,0,1,0)) via _mm_dp_ps: (u*0 + v*0) + nz*1
      const float dt = fadd(fadd(fmul(p.u, 0.f), fmul(p.v, 0.f)), fmul(nz, 1.f));
      const float cl = dt < 0.f ? 0.f : (dt > 1.f ? 1.f : dt);
      const float dif = fmul(cl, occ);
      return 0xff000000u | (trunc_u32(fmul((float)p.b, dif)) << 16) | (trunc_u32(fmul((float)p.g, dif)) << 8) | trunc_u32(fmul((float)p.r, dif));
    }
    if (p.mark & 1u) return 0xff000000u | ((p.b >> 2) << 16) | ((p.g >> 2) << 8) | (p.r >> 2);
    return 0xff000000u | (p.b << 16) | (p.g << 8) | p.r;
  }
  return get_color(s, get_U(p.u, s.mw), get_V(p.v, s.mh), p.mark);
}

__device__ __forceinline__ bool ndiff(const Px& a, const Px& b) {
  const float thr = 0.001f;
  return (fabsf(fsub(a.u, b.u)) > thr) || (fabsf(fsub(a.v, b.v)) > thr);
}

__global__ void __launch_bounds__(256) shade_kernel(const j3dg_pixel* __restrict__ px, uint32_t pstride, ShadeParams s,
                                                     const uint32_t* __restrict__ bg, uint32_t bg_stride,
                                                     uint32_t* __restrict__ rgba, uint32_t rstride) {
  const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
  const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= s.w || y >= s.h) return;
  if (s.shard_world > 1u && (y / J3DG_SHARD_BAND_ROWS) % s.shard_world != s.shard_rank) return;  // another rank's band
  const j3dg_pixel* pp = px + (size_t)y * pstride + x;
  const Px p = load_px(pp);
  uint32_t* o = rgba + (size_t)y * rstride + x;
  if (p.object_id == 0xFFFFFFFFu) {  // miss: keep what render_scene copied there (the background)
    if (bg) *o = __ldg(bg + (size_t)y * bg_stride + x);
    return;
  }
  const bool last = (x == s.w - 1);
  uint32_t clr;
  if (s.flags & J3DG_ONE_BIT) {  // canvas.cpp:498-579
    const uint32_t c = get_color(s, get_U(p.u, s.mw), get_V(p.v, s.mh), p.mark);
    const uint32_t res = ((((c & 0xff0000u) >> 16) + ((c & 0xff00u) >> 8) + (c & 0xffu)) >> 7) + 1u;
    if (last) {
      clr = ((((s.w - 1) % res) == 0) && ((y % res) == 0)) ? 0xff000000u : 0xffffffffu;
    } else {
      float angle = 1.f;
      const Px right = load_px(pp + 1);
      const Px up = load_px(px + (size_t)(y ? y - 1 : 0) * pstride + x);
      const Px* q = nullptr;
      if (right.object_id != 0xFFFFFFFFu && ndiff(p, right)) q = &right;
      else if (up.object_id != 0xFFFFFFFFu && ndiff(p, up)) q = &up;
      if (q) {
        const float w1 = normal_z(p.u, p.v), w2 = normal_z(q->u, q->v);
        angle = fadd(fadd(fmul(p.u, q->u), fmul(p.v, q->v)), fmul(w1, w2));
      }
      bool black = ((x % res) == 0) && ((y % res) == 0);
      if (fabsf(angle) < 0.95f) black = (res == 1) ? !black : true;
      clr = black ? 0xff000000u : 0xffffffffu;
    }
  } else if (s.flags & J3DG_WIREFRAME) {  // canvas.cpp:449-496
    bool wire = false;
    if (!last) {
      const uint32_t rid = __ldg(reinterpret_cast<const uint32_t*>(pp + 1) + 4);
      const uint32_t uid = __ldg(reinterpret_cast<const uint32_t*>(px + (size_t)(y ? y - 1 : 0) * pstride + x) + 4);
      wire = (rid != 0xFFFFFFFFu && rid != p.object_id) || (uid != 0xFFFFFFFFu && uid != p.object_id);
    }
    if (wire) {
      const float scale = fmul(fadd(fmul(p.u, p.u), fmul(p.v, p.v)), 0.5f);
      const uint32_t c = (uint32_t)__float2int_rz(fmul(255.f, scale)) & 0xffu;
      clr = 0xff000000u | (c << 16) | (c << 8) | c;
    } else clr = plain_color(s, p);
  } else if (s.flags & J3DG_EDGES) {  // canvas.cpp:590-649
    clr = 0;
    bool done = false;
    if (!last) {
      const Px right = load_px(pp + 1);
      if (right.object_id != 0xFFFFFFFFu && ndiff(p, right)) {
        const float a = convex_cos_angle(s, (float)x, (float)y, p.u, p.v, p.depth, fadd((float)x, 1.f), (float)y, right.u, right.v, right.depth);
        clr = get_angle_color(s, a, p.u, p.v, p.mark);
        done = true;
      } else {
        const Px up = load_px(px + (size_t)(y ? y - 1 : 0) * pstride + x);
        if (up.object_id != 0xFFFFFFFFu && ndiff(p, up)) {
          const float a = convex_cos_angle(s, (float)x, (float)y, p.u, p.v, p.depth, (float)x, fsub((float)y, 1.f), up.u, up.v, up.depth);
          clr = get_angle_color(s, a, p.u, p.v, p.mark);
          done = true;
        }
      }
    }
    if (!done) clr = plain_color(s, p);
  } else {  // canvas.cpp:650-669
    clr = plain_color(s, p);
  }
  *o = clr;
}

__global__ void __launch_bounds__(256) background_kernel(uint32_t w, uint32_t h, uint32_t top, uint32_t bottom, uint32_t* __restrict__ out, uint32_t stride) {
  // fill_background, canvas.cpp:55-78
  const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
  const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const float scale = fdiv((float)y, (float)h);
  const float inv = fsub(1.f, scale);
  uint32_t ch[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float t = (float)((top >> (8 * k)) & 0xffu), b = (float)((bottom >> (8 * k)) & 0xffu);
    ch[k] = trunc_u32(fadd(fmul(scale, b), fmul(inv, t))) & 0xffu;
  }
  out[(size_t)y * stride + x] = 0xff000000u | (ch[2] << 16) | (ch[1] << 8) | ch[0];
}

}  // namespace

int j3dg_launch_shade(j3dg_ctx* ctx, const j3dg_pixel* d_pixels, uint32_t pstride, const j3dg_view* view,
                      const uint32_t* d_matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity,
                      const uint32_t* d_bg, uint32_t bg_stride, uint32_t* d_rgba, uint32_t rstride) {
  if (!view->width || !view->height) return J3DG_OK;
  ShadeParams s;
  s.w = view->width; s.h = view->height; s.flags = view->flags; s.near_plane = view->near_plane;
  memcpy(s.pinv, view->projection_inv, 64);
  s.matcap = d_matcap; s.mw = mw; s.mh = mh; s.mstride = mstride; s.cavity = cavity;
  s.shard_rank = ctx->shard_rank; s.shard_world = ctx->shard_world;
  dim3 grid((s.w + 31) / 32, (s.h + 7) / 8);
  { int rc = j3dg_stage_begin(ctx, 1); if (rc != J3DG_OK) return rc; }
  shade_kernel<<<grid, 256, 0, ctx->stream>>>(d_pixels, pstride, s, d_bg, bg_stride, d_rgba, rstride);
  KERNEL_CHECK(ctx);
  return j3dg_stage_end(ctx, 1);
}

int j3dg_launch_background(j3dg_ctx* ctx, uint32_t w, uint32_t h, uint32_t top, uint32_t bottom, uint32_t* d_out, uint32_t stride) {
  if (!w || !h) return J3DG_OK;
  dim3 grid((w + 31) / 32, (h + 7) / 8);
  background_kernel<<<grid, 256, 0, ctx->stream>>>(w, h, top, bottom, d_out, stride);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}
