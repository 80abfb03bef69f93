// splat.cu — point-cloud depth splat.  Replaces canvas::render_pointclouds_on_image
// (j3d/canvas.cpp:952-1030) = jtk::bind / _draw / present (jtk/render.h:254-288, 307-512, 514-865).
//
// The reference projects all points in parallel and then z-tests them in ONE serial loop, four
// points per SSE packet: all four lanes test against the PRE-packet depth, then write in lane order; a lane
// that fails writes the old depth / colour back; masked lanes are clamped onto some pixel and write it back
// too (render.h:783-807).  Stated per pixel q, with z the pixel's depth before a packet and d_1..d_j the
// depths of the packet's lanes that land on q (lane order, masked lanes = never pass):
//     depth, colour :  z' = max(z, d_j) — only the LAST lane on q counts; an earlier lane's write is always
//                      overwritten by it (with its own values if it passes, with the old ones if not)
//     pixel record  :  the canvas callback fires for EVERY passing lane (z < d_i); the last one sticks
// Folding that over all packets in order gives an order-free statement:
//     M  = max(z0, depths of last-on-q lanes)                      -> final depth
//     W  = the lowest-index last-on-q lane with depth M (if M > z0) -> final colour
//     id = the highest-index NOT-last-on-q lane after W with depth > M, else W
// (after W's packet z equals M for good, so a later lane passes iff its depth exceeds M, which only a lane
// that is not the last one on q can do).  So ONE pass suffices: last-on-q lanes do a 64-bit atomicMax on
// (float bits of 1/w) << 32 | (0xFFFFFFFE - index) (1/w > 0, so the bit pattern is monotone; seeded from the
// pixel buffer with low word 0xFFFFFFFF, i.e. "strictly in front of the mesh", lowest index wins ties);
// the rare lanes that share a pixel with a LATER lane of their own packet go to a small list that is
// checked against the final M in a second, tiny kernel (atomicMax of index + 1 per pixel).  A resolve pass
// then shades only the winners (colour / normal are fetched for <= W*H points instead of all N) and patches
// the pixel records.  Bit-identical to the reference's serial loop, without a sort, a second projection pass
// or a host round trip (round 1 replayed the collisions sequentially: 1.3 of its 2.2 ms on 100 M points).
//
// Projection, rounding (round-to-nearest-even; truncation + clip test for the last N mod 4
// points) and the Lambert term follow the reference's operation order with unfused arithmetic.
#include "common.cuh"

#include <algorithm>

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
                                                    unsigned long long* __restrict__ packed, uint32_t* __restrict__ idcand) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const j3dg_pixel* p = px + (size_t)y * pstride + x;
  const uint32_t db = __ldg(reinterpret_cast<const uint32_t*>(p) + 7);
  const float depth = __ldg(reinterpret_cast<const float*>(p) + 3);
  const float z = db != 0u ? fdiv(1.f, depth) : 0.f;  // canvas.cpp:968
  packed[(size_t)y * w + x] = ((unsigned long long)__float_as_uint(z) << 32) | 0xFFFFFFFFull;
  idcand[(size_t)y * w + x] = 0u;
}

// _mm_cvtps_epi32 / cvttss2si semantics: NaN and out-of-range give INT_MIN
__device__ __forceinline__ int cvt_rne(float x) { return (fabsf(x) < 2147483648.f) ? __float2int_rn(x) : (int)0x80000000; }
__device__ __forceinline__ int cvt_trunc(float x) { return (fabsf(x) < 2147483648.f) ? __float2int_rz(x) : (int)0x80000000; }

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

// A lane that shares its pixel with a later lane of its own packet (see the header).
struct Anomaly { uint32_t pixel, index, depth_bits; };

__device__ __forceinline__ void splat_max(unsigned long long* __restrict__ packed, int idx, float depth, uint32_t index) {
  const unsigned long long word = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned long long)(0xFFFFFFFEu - index);
  unsigned long long* cell = packed + idx;
  if (word > *((volatile unsigned long long*)cell)) atomicMax(cell, word);  // the plain read filters the ~98 % of the points that lose
}

// Thread t < npackets: one SIMD packet (4 points); the remaining threads: one tail point each.
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ pos, SplatParams s, unsigned long long* __restrict__ packed,
                                                       uint32_t* __restrict__ anomaly_count, Anomaly* __restrict__ anomalies) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t npackets = s.tail_start >> 2;
  uint32_t nanom = 0;     // lanes of this packet that go to the anomaly list (bit k)
  Lane l[4];
  if (t < npackets) {
    float c[12];
    load_packet(pos, t, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) l[k] = project_simd(s, c[3 * k], c[3 * k + 1], c[3 * k + 2]);
    if (!(l[0].masked && l[1].masked && l[2].masked && l[3].masked)) {  // render.h:734-735
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (l[k].masked || !(l[k].depth > 0.f)) continue;  // the z-buffer is never negative: such a lane never passes
        bool last = true;                                   // no later lane (masked or not) lands on the same pixel
#pragma unroll
        for (int j = k + 1; j < 4; ++j) last = last && (l[j].idx != l[k].idx);
        if (last) splat_max(packed, l[k].idx, l[k].depth, 4u * t + k);
        else nanom |= 1u << k;
      }
    }
  } else {
    const uint32_t i = s.tail_start + (t - npackets);
    if (i < s.n) {
      const Lane tl = project_tail(s, __ldg(pos + 3 * (size_t)i), __ldg(pos + 3 * (size_t)i + 1), __ldg(pos + 3 * (size_t)i + 2));
      if (!tl.masked && tl.depth > 0.f) splat_max(packed, tl.idx, tl.depth, i);  // render.h:847-862: plain d > z
    }
  }
  // append the anomalies: one atomic per warp
  if (__any_sync(0xffffffffu, nanom != 0u)) {
    const uint32_t mine = __popc(nanom);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)(threadIdx.x & 31) >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(anomaly_count, total);
    base = __shfl_sync(0xffffffffu, base, 0) + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (nanom & (1u << k)) anomalies[base++] = Anomaly{(uint32_t)l[k].idx, 4u * t + k, __float_as_uint(l[k].depth)};
  }
}

// The lanes on the list fire the canvas callback iff their depth beats the pixel's FINAL depth and they come after
// the winner W in index order; the highest such index names the pixel record (header).  Runs after project_kernel.
__global__ void __launch_bounds__(256) anomaly_kernel(const uint32_t* __restrict__ anomaly_count, const Anomaly* __restrict__ anomalies,
                                                       const unsigned long long* __restrict__ packed, uint32_t* __restrict__ idcand) {
  const uint32_t count = *anomaly_count;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const Anomaly a = anomalies[e];
    const unsigned long long word = packed[a.pixel];
    const uint32_t low = (uint32_t)word;
    const float m = __uint_as_float((uint32_t)(word >> 32));
    if (!(m < __uint_as_float(a.depth_bits))) continue;                      // render.h:791-792 against the final depth
    if (low != 0xFFFFFFFFu && a.index < 0xFFFFFFFEu - low) continue;         // before the winner: overwritten by it
    atomicMax(idcand + a.pixel, a.index + 1u);
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

__global__ void __launch_bounds__(256) resolve_kernel(SplatParams s, const float* __restrict__ nrm, const uint32_t* __restrict__ clr,
                                                       unsigned long long* __restrict__ packed, uint32_t* __restrict__ idcand,
                                                       j3dg_pixel* __restrict__ px, uint32_t pstride, uint32_t* __restrict__ rgba, uint32_t rstride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= s.w || y >= s.h) return;
  const size_t q = (size_t)y * s.w + x;
  const unsigned long long word = packed[q];
  const uint32_t low = (uint32_t)word;
  const uint32_t cand = idcand[q];
  if (low == 0xFFFFFFFFu && cand == 0u) return;  // no point in front of what was there
  const float zb = __uint_as_float((uint32_t)(word >> 32));
  uint32_t id_point;
  if (low != 0xFFFFFFFFu) {  // a winner: depth and colour change
    const uint32_t i = 0xFFFFFFFEu - low;
    rgba[(size_t)y * rstride + x] = shade_point(s, nrm, clr, i);
    id_point = i;
    packed[q] = (word & 0xFFFFFFFF00000000ull) | 0xFFFFFFFFull;  // becomes the depth the next cloud has to beat
  }
  if (cand) {                // a later lane of a colliding packet fired the callback last
    id_point = cand - 1u;
    idcand[q] = 0u;
  }
  j3dg_pixel* p = px + (size_t)y * pstride + x;  // canvas.cpp:997-1027
  p->object_id = id_point;
  p->depth = fdiv(1.f, zb);
  p->db_id = s.db_id;
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

}  // namespace

int j3dg_launch_splat(j3dg_ctx* ctx, j3dg_cloud* const* clouds, uint32_t nc, const j3dg_view* view,
                      const j3dg_pixel* d_px_in, j3dg_pixel* d_px_inout, uint32_t pstride, uint32_t* d_rgba, uint32_t rstride) {
  if (!nc) return J3DG_OK;  // canvas.cpp:956
  const int w = (int)view->width, h = (int)view->height;
  if (w <= 0 || h <= 0) return J3DG_OK;
  const size_t npx = (size_t)w * h;
  // packed words | id candidates | anomaly counter
  const size_t off_cand = npx * sizeof(unsigned long long);
  const size_t off_cnt = (off_cand + npx * sizeof(uint32_t) + 255) & ~(size_t)255;
  {
    void* p = ctx->d_packed;
    int rc = j3dg_reserve(ctx, &p, &ctx->packed_cap, off_cnt + 256);
    ctx->d_packed = (unsigned long long*)p;
    if (rc != J3DG_OK) return rc;
  }
  // anomaly list: at most three lanes of every packet (a lane is listed only if a LATER lane shares its pixel)
  uint32_t max_n = 0;
  for (uint32_t c = 0; c < nc; ++c) {
    if (!clouds[c]) { j3dg_set_error(ctx, "j3dg_splat: null cloud"); return J3DG_EINVAL; }
    max_n = std::max(max_n, clouds[c]->n);
  }
  {
    const size_t need = ((size_t)(max_n >> 2) * 3 + 4) * sizeof(Anomaly);
    int rc = j3dg_reserve(ctx, &ctx->d_misc, &ctx->misc_cap, need);
    if (rc != J3DG_OK) return rc;
  }
  unsigned long long* packed = ctx->d_packed;
  uint32_t* idcand = (uint32_t*)((char*)ctx->d_packed + off_cand);
  uint32_t* counter = (uint32_t*)((char*)ctx->d_packed + off_cnt);
  Anomaly* anomalies = (Anomaly*)ctx->d_misc;
  dim3 pgrid((w + 31) / 32, (h + 7) / 8);
  { int rc = j3dg_stage_begin(ctx, 2); if (rc != J3DG_OK) return rc; }
  seed_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_px_in, pstride, w, h, packed, idcand);
  KERNEL_CHECK(ctx);
  for (uint32_t c = 0; c < nc; ++c) {
    const j3dg_cloud* cl = clouds[c];
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
    if (!cl->n) continue;
    const uint32_t threads = (s.tail_start >> 2) + (cl->n - s.tail_start);
    CU_CHECK(ctx, cudaMemsetAsync(counter, 0, sizeof(uint32_t), ctx->stream));
    project_kernel<<<(threads + 255) / 256, 256, 0, ctx->stream>>>(cl->d_pos, s, packed, counter, anomalies);
    KERNEL_CHECK(ctx);
    anomaly_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(counter, anomalies, packed, idcand);
    KERNEL_CHECK(ctx);
    resolve_kernel<<<pgrid, 256, 0, ctx->stream>>>(s, cl->d_nrm, cl->d_clr, packed, idcand, d_px_inout, pstride, d_rgba, rstride);
    KERNEL_CHECK(ctx);
  }
  return j3dg_stage_end(ctx, 2);
}
