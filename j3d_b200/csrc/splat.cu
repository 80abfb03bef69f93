// splat.cu — point-cloud depth splat.  Replaces canvas::render_pointclouds_on_image
// (j3d/canvas.cpp:952-1030) = jtk::bind / _draw / present (jtk/render.h:254-288, 307-512, 514-865).
//
// The reference projects all points in parallel and then z-tests them in ONE serial loop, four
// points per SSE packet.  For a pixel that is touched by at most one lane of any packet the
// outcome is order-free: the point with the largest 1/w wins, the lowest index wins ties, and
// it must be strictly in front of the mesh depth.  That case is handled with one 64-bit
// atomicMax per point on a packed word (float bits of 1/w) << 32 | (0xFFFFFFFE - index)
// (1/w > 0 so the bit pattern is monotone), seeded from the pixel buffer with low word
// 0xFFFFFFFF; a resolve pass then shades only the winners (colour / normal are fetched for
// ~W*H points instead of all N) and patches the pixel records.
//
// When two lanes of the SAME packet land on one pixel the reference is order-dependent
// (render.h:783-807: all four lanes test against the pre-packet depth, then write in lane
// order; a lane that fails writes the old value back; masked lanes are clamped onto some pixel
// and write it back too).  Those pixels are detected during projection ("dirty"), every point
// that touches a dirty pixel is collected, radix-sorted by (pixel, index) and replayed
// sequentially per pixel with the reference's exact packet semantics — bit-identical results,
// at a cost proportional to the number of dirty pixels (a few hundred for a random 100 M-point
// cloud at 1080p; the whole cloud in the worst case of scan-ordered input).
//
// Projection, rounding (round-to-nearest-even; truncation + clip test for the last N mod 4
// points) and the Lambert term follow the reference's operation order with unfused arithmetic.
#include "common.cuh"
#include "sort.cuh"

namespace {

struct SplatParams {
  float M[16];      // projection * (camera_position * object_system), render.h:279-280
  float light[4];   // render.h:521-527
  int w, h;
  uint32_t n, tail_start;  // tail_start = n - (n & 3)
  uint32_t db_id;
  uint32_t use_normals, use_colors;
};

__global__ void __launch_bounds__(256) seed_kernel(const j3dg_pixel* __restrict__ px, uint32_t pstride, int w, int h,
                                                    unsigned long long* __restrict__ packed, float* __restrict__ zprev) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const j3dg_pixel* p = px + (size_t)y * pstride + x;
  const uint32_t db = __ldg(reinterpret_cast<const uint32_t*>(p) + 7);
  const float depth = __ldg(reinterpret_cast<const float*>(p) + 3);
  const float z = db != 0u ? fdiv(1.f, depth) : 0.f;  // canvas.cpp:968
  packed[(size_t)y * w + x] = ((unsigned long long)__float_as_uint(z) << 32) | 0xFFFFFFFFull;
  zprev[(size_t)y * w + x] = z;
}

// _mm_cvtps_epi32 / cvttss2si semantics: NaN and out-of-range give INT_MIN
__device__ __forceinline__ int cvt_rne(float x) { return (fabsf(x) < 2147483648.f) ? __float2int_rn(x) : (int)0x80000000; }
__device__ __forceinline__ int cvt_trunc(float x) { return (fabsf(x) < 2147483648.f) ? __float2int_rz(x) : (int)0x80000000; }

constexpr uint32_t REPLAY_MASKED = 0xFFFFFFFFu;  // a NaN pattern: like a masked lane, a NaN depth never passes the test

struct Lane {
  int idx;       // clamped pixel index (render.h:779-781)
  bool masked;   // outside the canvas (render.h:729-732)
  float depth;   // 1 / VW (render.h:774)
};

// One lane of the SIMD body: render.h:419-468 (projection) + 726-732 (rounding, mask) + 779-781 (index)
__device__ __forceinline__ Lane project_simd(const SplatParams& s, float x, float y, float z) {
  const float* M = s.M;
  float VX = fadd(fadd(fadd(fmul(M[0], x), fmul(M[4], y)), fmul(M[8], z)), fmul(M[12], 1.f));
  float VY = fadd(fadd(fadd(fmul(M[1], x), fmul(M[5], y)), fmul(M[9], z)), fmul(M[13], 1.f));
  const float VW = fadd(fadd(fadd(fmul(M[3], x), fmul(M[7], y)), fmul(M[11], z)), fmul(M[15], 1.f));
  VX = fdiv(VX, VW); VY = fdiv(VY, VW);
  VX = fmul(fadd(VX, 1.f), fmul((float)s.w, 0.5f));
  VY = fmul(fadd(VY, 1.f), fmul((float)s.h, 0.5f));
  const int X = cvt_rne(VX), Y = cvt_rne(VY);
  Lane l;
  l.masked = (0 > X) || (0 > Y) || (X > s.w - 1) || (Y > s.h - 1);
  int id = (int)((uint32_t)X + (uint32_t)s.w * (uint32_t)Y);  // wraps like _mm_mullo_epi32 / _mm_add_epi32
  id = max(0, id);
  id = min(s.w * s.h - 1, id);
  l.idx = id;
  l.depth = fdiv(1.f, VW);
  return l;
}

// Scalar tail point: render.h:473-511 (projection + clip info) and 814-821.  masked = skipped.
__device__ __forceinline__ Lane project_tail(const SplatParams& s, float x, float y, float z) {
  const float* M = s.M;
  float VX = fadd(fadd(fadd(fmul(M[0], x), fmul(M[4], y)), fmul(M[8], z)), M[12]);
  float VY = fadd(fadd(fadd(fmul(M[1], x), fmul(M[5], y)), fmul(M[9], z)), M[13]);
  float VZ = fadd(fadd(fadd(fmul(M[2], x), fmul(M[6], y)), fmul(M[10], z)), M[14]);
  const float VW = fadd(fadd(fadd(fmul(M[3], x), fmul(M[7], y)), fmul(M[11], z)), M[15]);
  VX = fdiv(VX, VW); VY = fdiv(VY, VW); VZ = fdiv(VZ, VW);
  Lane l;
  l.depth = fdiv(1.f, VW);
  l.idx = 0;
  l.masked = true;
  if (VX < -1.f || VX > 1.f || VY < -1.f || VY > 1.f || VZ < -1.f || VZ > 1.f) return l;  // vertex_clip_info
  VX = fmul(fmul(fadd(VX, 1.f), (float)s.w), 0.5f);
  VY = fmul(fmul(fadd(VY, 1.f), (float)s.h), 0.5f);
  const int X = cvt_trunc(VX), Y = cvt_trunc(VY);
  if (X < 0 || Y < 0 || X > s.w - 1 || Y > s.h - 1) return l;
  l.masked = false;
  l.idx = X + s.w * Y;
  return l;
}

__device__ __forceinline__ void load_packet(const float* __restrict__ pos, uint32_t packet, float (&c)[12]) {
  const float4* p = reinterpret_cast<const float4*>(pos + 12 * (size_t)packet);  // 48-byte packets of a 256-byte aligned array
  const float4 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
  c[8] = d.x; c[9] = d.y; c[10] = d.z; c[11] = d.w;
}

// Thread t < npackets: one SIMD packet (4 points); the remaining threads: one tail point each.
// COLLECT = false: atomicMax splat + dirty-pixel detection.
// COLLECT = true : append every (pixel, point) pair that touches a dirty pixel to `list`.
template <bool COLLECT>
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ pos, SplatParams s, unsigned long long* __restrict__ packed,
                                                       uint8_t* __restrict__ dirty, uint32_t* __restrict__ counters,
                                                       unsigned long long* __restrict__ list, uint32_t* __restrict__ list_depth) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t npackets = s.tail_start >> 2;
  if (t < npackets) {
    float c[12];
    load_packet(pos, t, c);
    Lane l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) l[k] = project_simd(s, c[3 * k], c[3 * k + 1], c[3 * k + 2]);
    if (l[0].masked && l[1].masked && l[2].masked && l[3].masked) return;  // render.h:734-735
    if (!COLLECT) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b)
          if (l[a].idx == l[b].idx && !(l[a].masked && l[b].masked)) {
            dirty[l[a].idx] = 1;
            counters[0] = 1u;
          }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (l[k].masked || !(l[k].depth > 0.f)) continue;  // the z-buffer is never negative
        const unsigned long long word = ((unsigned long long)__float_as_uint(l[k].depth) << 32) | (unsigned long long)(0xFFFFFFFEu - (4u * t + k));
        unsigned long long* cell = packed + l[k].idx;
        if (word > *((volatile unsigned long long*)cell)) atomicMax(cell, word);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (dirty[l[k].idx]) {  // the replay needs the lane's depth, or that it is masked (then it only writes the old value back)
          const uint32_t e = atomicAdd(&counters[1], 1u);
          list[e] = ((unsigned long long)(uint32_t)l[k].idx << 32) | (unsigned long long)(4u * t + k);
          list_depth[e] = l[k].masked ? REPLAY_MASKED : __float_as_uint(l[k].depth);
        }
    }
  } else {
    const uint32_t i = s.tail_start + (t - npackets);
    if (i >= s.n) return;
    const Lane l = project_tail(s, __ldg(pos + 3 * (size_t)i), __ldg(pos + 3 * (size_t)i + 1), __ldg(pos + 3 * (size_t)i + 2));
    if (l.masked) return;
    if (!COLLECT) {
      if (!(l.depth > 0.f)) return;
      const unsigned long long word = ((unsigned long long)__float_as_uint(l.depth) << 32) | (unsigned long long)(0xFFFFFFFEu - i);
      unsigned long long* cell = packed + l.idx;
      if (word > *((volatile unsigned long long*)cell)) atomicMax(cell, word);
    } else if (dirty[l.idx]) {
      const uint32_t e = atomicAdd(&counters[1], 1u);
      list[e] = ((unsigned long long)(uint32_t)l.idx << 32) | (unsigned long long)i;
      list_depth[e] = __float_as_uint(l.depth);
    }
  }
}

// Colour of point i as present() computes it: vertex colour (white if none), optionally Lambert-shaded
__device__ __forceinline__ uint32_t shade_point(const SplatParams& s, const float* __restrict__ nrm, const uint32_t* __restrict__ clr, uint32_t i) {
  uint32_t color = s.use_colors ? __ldg(clr + i) : 0xffffffffu;
  if (s.use_normals) {
    const float nx = __ldg(nrm + 3 * (size_t)i), ny = __ldg(nrm + 3 * (size_t)i + 1), nz = __ldg(nrm + 3 * (size_t)i + 2);
    const float red = (float)(color & 0xffu), green = (float)((color >> 8) & 0xffu), blue = (float)((color >> 16) & 0xffu);
    if (i < s.tail_start) {  // render.h:744-772
      float d = fadd(0.5f, fadd(fadd(fmul(nx, s.light[0]), fmul(ny, s.light[1])), fmul(nz, s.light[2])));
      if (d < 0.f) d = 0.f;
      if (1.f < d) d = 1.f;
      // light colour is white: intensity * (255/255.f) == intensity
      const int r2 = min(255, cvt_rne(fmul(d, red))), g2 = min(255, cvt_rne(fmul(d, green))), b2 = min(255, cvt_rne(fmul(d, blue)));
      color = 0xff000000u + ((uint32_t)b2 << 16) + ((uint32_t)g2 << 8) + (uint32_t)r2;
    } else {  // render.h:829-846
      float d = fadd(fadd(fadd(0.5f, fmul(nx, s.light[0])), fmul(ny, s.light[1])), fmul(nz, s.light[2]));
      d = d < 0.f ? 0.f : (1.f < d ? 1.f : d);
      const float fr = fmul(red, d), fg = fmul(green, d), fb = fmul(blue, d);
      const int r2 = cvt_trunc(255.f < fr ? 255.f : fr), g2 = cvt_trunc(255.f < fg ? 255.f : fg), b2 = cvt_trunc(255.f < fb ? 255.f : fb);
      color = 0xff000000u | ((uint32_t)b2 << 16) | ((uint32_t)g2 << 8) | (uint32_t)r2;
    }
  }
  return color;
}

// Sequential replay of every dirty pixel with the reference's packet semantics.  `list` is sorted by
// (pixel, point index) and carries each entry's projected depth (REPLAY_MASKED for lanes outside the canvas),
// so the fold over a pixel's entries touches no point data; the thread at the first entry of a pixel walks
// that pixel's entries and shades the surviving point once at the end.
__global__ void __launch_bounds__(128) replay_kernel(SplatParams s, const float* __restrict__ nrm, const uint32_t* __restrict__ clr,
                                                      const unsigned long long* __restrict__ list, const uint32_t* __restrict__ list_depth, uint32_t count,
                                                      unsigned long long* __restrict__ packed, float* __restrict__ zprev,
                                                      j3dg_pixel* __restrict__ px, uint32_t pstride, uint32_t* __restrict__ rgba, uint32_t rstride) {
  const uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k0 >= count) return;
  const uint32_t q = (uint32_t)(list[k0] >> 32);
  if (k0 > 0 && (uint32_t)(list[k0 - 1] >> 32) == q) return;  // not the first entry of its pixel
  float z = zprev[q];
  constexpr uint32_t NONE = 0xFFFFFFFFu;
  uint32_t color_point = NONE;  // the point whose colour the pixel ends up with
  uint32_t id_point = NONE;     // the point the pixel record ends up naming, and the z-buffer value at that moment
  float id_z = 0.f;
  uint32_t k = k0;
  while (k < count) {
    unsigned long long e = list[k];
    if ((uint32_t)(e >> 32) != q) break;
    const uint32_t i = (uint32_t)e;
    if (i >= s.tail_start) {  // scalar tail point, render.h:847-862 (clipped ones were never collected)
      const float d = __uint_as_float(list_depth[k]);
      if (d > z) { z = d; color_point = i; id_point = i; id_z = z; }
      ++k;
      continue;
    }
    // all lanes of this packet that touch the pixel: every test sees the pre-packet depth
    const uint32_t packet = i >> 2;
    const float z_old = z;
    uint32_t last_pass_point = NONE;
    bool last_lane_pass = false;
    uint32_t last_lane_point = i;
    float last_lane_depth = 0.f;
    while (k < count) {
      e = list[k];
      const uint32_t pi = (uint32_t)e;
      if ((uint32_t)(e >> 32) != q || (pi >> 2) != packet || pi >= s.tail_start) break;
      const uint32_t bits = list_depth[k];
      const float d = __uint_as_float(bits);
      const bool pass = bits != REPLAY_MASKED && (z_old < d);  // render.h:791-792
      if (pass) last_pass_point = pi;                           // the last passing lane's callback wins (canvas.cpp:999-1026)
      last_lane_pass = pass; last_lane_point = pi; last_lane_depth = d;
      ++k;
    }
    if (last_lane_pass) {  // the highest lane's write is the one that sticks (render.h:797-805)
      z = last_lane_depth;
      color_point = last_lane_point;
    }  // else: it wrote the pre-packet depth and colour back
    if (last_pass_point != NONE) { id_point = last_pass_point; id_z = z; }  // 1 / zbuffer after the packet's writes
  }
  const int qx = (int)(q % (uint32_t)s.w), qy = (int)(q / (uint32_t)s.w);
  if (color_point != NONE) rgba[(size_t)qy * rstride + qx] = shade_point(s, nrm, clr, color_point);
  if (id_point != NONE) {
    j3dg_pixel* out_px = px + (size_t)qy * pstride + qx;
    out_px->object_id = id_point;
    out_px->depth = fdiv(1.f, id_z);
    out_px->db_id = s.db_id;
  }
  zprev[q] = z;
  packed[q] = ((unsigned long long)__float_as_uint(z) << 32) | 0xFFFFFFFFull;  // resolved: resolve_kernel skips it
}

__global__ void __launch_bounds__(256) resolve_kernel(SplatParams s, const float* __restrict__ nrm, const uint32_t* __restrict__ clr,
                                                       unsigned long long* __restrict__ packed, float* __restrict__ zprev,
                                                       j3dg_pixel* __restrict__ px, uint32_t pstride, uint32_t* __restrict__ rgba, uint32_t rstride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= s.w || y >= s.h) return;
  unsigned long long* cell = packed + (size_t)y * s.w + x;
  const unsigned long long word = *cell;
  const uint32_t low = (uint32_t)word;
  if (low == 0xFFFFFFFFu) return;  // no point in front of what was there (or already replayed)
  const uint32_t i = 0xFFFFFFFEu - low;
  const float zb = __uint_as_float((uint32_t)(word >> 32));
  rgba[(size_t)y * rstride + x] = shade_point(s, nrm, clr, i);
  j3dg_pixel* p = px + (size_t)y * pstride + x;  // canvas.cpp:997-1027
  p->object_id = i;
  p->depth = fdiv(1.f, zb);
  p->db_id = s.db_id;
  *cell = (word & 0xFFFFFFFF00000000ull) | 0xFFFFFFFFull;  // becomes the depth the next cloud has to beat
  zprev[(size_t)y * s.w + x] = zb;
}

// render.h helpers, host side, each operation rounded separately
void r_matmul(float* out, const float* left, const float* right) {  // render.h:224-233
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      volatile float a = left[i] * right[(j << 2)], b = left[i + 4] * right[(j << 2) + 1];
      volatile float c = left[i + 8] * right[(j << 2) + 2], d = left[i + 12] * right[(j << 2) + 3];
      volatile float s = a + b;
      s = s + c;
      s = s + d;
      out[i + (j << 2)] = s;
    }
}
void r_matvec(float* out, const float* m, const float* v) {  // render.h:235-241
  for (int r = 0; r < 4; ++r) {
    volatile float a = m[r] * v[0], b = m[4 + r] * v[1], c = m[8 + r] * v[2], d = m[12 + r] * v[3];
    volatile float s = a + b;
    s = s + c;
    s = s + d;
    out[r] = s;
  }
}
void r_invert_orthonormal(float* out, const float* in) {  // render.h:133-154
  out[0] = in[0]; out[1] = in[4]; out[2] = in[8]; out[4] = in[1]; out[5] = in[5]; out[6] = in[9];
  out[8] = in[2]; out[9] = in[6]; out[10] = in[10]; out[3] = 0; out[7] = 0; out[11] = 0; out[15] = 1;
  for (int k = 0; k < 3; ++k) {
    volatile float a = in[4 * k] * in[12], b = in[4 * k + 1] * in[13], c = in[4 * k + 2] * in[14];
    volatile float s = a + b;
    s = s + c;
    out[12 + k] = -s;
  }
}

int bits_for(uint64_t v) {
  int b = 0;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

}  // namespace

int j3dg_launch_splat(j3dg_ctx* ctx, j3dg_cloud* const* clouds, uint32_t nc, const j3dg_view* view,
                      const j3dg_pixel* d_px_in, j3dg_pixel* d_px_inout, uint32_t pstride, uint32_t* d_rgba, uint32_t rstride) {
  if (!nc) return J3DG_OK;  // canvas.cpp:956
  const int w = (int)view->width, h = (int)view->height;
  if (w <= 0 || h <= 0) return J3DG_OK;
  const size_t npx = (size_t)w * h;
  // packed words | previous depth | dirty flags | counters
  const size_t off_z = npx * sizeof(unsigned long long);
  const size_t off_dirty = off_z + npx * sizeof(float);
  const size_t off_cnt = (off_dirty + npx + 255) & ~(size_t)255;
  {
    void* p = ctx->d_packed;
    int rc = j3dg_reserve(ctx, &p, &ctx->packed_cap, off_cnt + 256);
    ctx->d_packed = (unsigned long long*)p;
    if (rc != J3DG_OK) return rc;
  }
  unsigned long long* packed = ctx->d_packed;
  float* zprev = (float*)((char*)ctx->d_packed + off_z);
  uint8_t* dirty = (uint8_t*)ctx->d_packed + off_dirty;
  uint32_t* counters = (uint32_t*)((char*)ctx->d_packed + off_cnt);
  dim3 pgrid((w + 31) / 32, (h + 7) / 8);
  { int rc = j3dg_stage_begin(ctx, 2); if (rc != J3DG_OK) return rc; }
  seed_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_px_in, pstride, w, h, packed, zprev);
  KERNEL_CHECK(ctx);
  for (uint32_t c = 0; c < nc; ++c) {
    const j3dg_cloud* cl = clouds[c];
    if (!cl) { j3dg_set_error(ctx, "j3dg_splat: null cloud"); return J3DG_EINVAL; }
    SplatParams s;
    float temp[16];
    r_matmul(temp, view->cs_inv, cl->cs);
    r_matmul(s.M, view->projection, temp);
    float inv[16], light[4] = {0.f, 0.f, 1.f, 0.f}, tmp[4];
    r_invert_orthonormal(inv, view->cs_inv);
    r_matvec(tmp, inv, light);
    r_invert_orthonormal(inv, cl->cs);
    r_matvec(s.light, inv, tmp);
    s.w = w; s.h = h; s.n = cl->n; s.tail_start = cl->n - (cl->n & 3u); s.db_id = cl->db_id;
    s.use_normals = ((view->flags & J3DG_SHADING) && cl->d_nrm) ? 1u : 0u;   // canvas.cpp:994
    s.use_colors = (!(view->flags & J3DG_ONE_BIT) && cl->d_clr) ? 1u : 0u;    // canvas.cpp:995
    if (cl->n) {
      const uint32_t threads = (s.tail_start >> 2) + (cl->n - s.tail_start);
      const uint32_t blocks = (threads + 255) / 256;
      CU_CHECK(ctx, cudaMemsetAsync(dirty, 0, npx, ctx->stream));
      CU_CHECK(ctx, cudaMemsetAsync(counters, 0, 16, ctx->stream));
      project_kernel<false><<<blocks, 256, 0, ctx->stream>>>(cl->d_pos, s, packed, dirty, counters, nullptr, nullptr);
      KERNEL_CHECK(ctx);
      uint32_t h_cnt[2] = {0, 0};
      CU_CHECK(ctx, cudaMemcpyAsync(h_cnt, counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
      CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      if (h_cnt[0]) {  // some packet hit one pixel with two lanes: exact sequential replay of those pixels
        const size_t nn = cl->n;
        const size_t need = 4 * 256 + 2 * nn * sizeof(uint64_t) + 2 * nn * sizeof(uint32_t) + rsort::scratch_bytes(cl->n) + 256;
        int rc = j3dg_reserve(ctx, &ctx->d_misc, &ctx->misc_cap, need);
        if (rc != J3DG_OK) return rc;
        char* base = (char*)ctx->d_misc;
        size_t off = 0;
        auto take = [&](size_t bytes) { char* p = base + off; off = (off + bytes + 255) & ~(size_t)255; return p; };
        uint64_t* keys_a = (uint64_t*)take(nn * 8);
        uint64_t* keys_b = (uint64_t*)take(nn * 8);
        uint32_t* vals_a = (uint32_t*)take(nn * 4);
        uint32_t* vals_b = (uint32_t*)take(nn * 4);
        uint32_t* scratch = (uint32_t*)take(rsort::scratch_bytes(cl->n));
        project_kernel<true><<<blocks, 256, 0, ctx->stream>>>(cl->d_pos, s, packed, dirty, counters, (unsigned long long*)keys_a, vals_a);
        KERNEL_CHECK(ctx);
        CU_CHECK(ctx, cudaMemcpyAsync(h_cnt, counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        const uint32_t count = h_cnt[1];
        if (count) {
          bool in_b = false;
          const int key_bits = 32 + bits_for(npx - 1);
          rc = rsort::sort_pairs(ctx, keys_a, vals_a, keys_b, vals_b, count, key_bits, scratch, &in_b, false);
          if (rc != J3DG_OK) return rc;
          const unsigned long long* sorted = (const unsigned long long*)(in_b ? keys_b : keys_a);
          replay_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(s, cl->d_nrm, cl->d_clr, sorted, in_b ? vals_b : vals_a, count, packed, zprev,
                                                                       d_px_inout, pstride, d_rgba, rstride);
          KERNEL_CHECK(ctx);
        }
      }
    }
    resolve_kernel<<<pgrid, 256, 0, ctx->stream>>>(s, cl->d_nrm, cl->d_clr, packed, zprev, d_px_inout, pstride, d_rgba, rstride);
    KERNEL_CHECK(ctx);
  }
  return j3dg_stage_end(ctx, 2);
}
