// splat.cu — point-cloud depth splat.  Replaces canvas::render_pointclouds_on_image
// (j3d/canvas.cpp:952-1030) = jtk::bind / _draw / present (jtk/render.h:254-288, 307-512, 514-865).
//
// The reference projects all points in parallel and then z-tests them in one serial loop; the
// result per pixel is the point with the largest 1/w, the lowest index winning ties, strictly
// in front of the mesh depth.  Here every point does one 64-bit atomicMax on a packed
// (float bits of 1/w) << 32 | (0xFFFFFFFE - index) word (1/w > 0, so the bit pattern is
// monotone), seeded from the pixel buffer with low word 0xFFFFFFFF so that a point must be
// strictly nearer than the mesh; a resolve pass then shades only the winners (colour and
// normal are fetched for ~W*H points instead of all N) and patches the pixel records.
// Projection, rounding (round-to-nearest-even; truncation + clip test for the last N mod 4
// points) and the Lambert term follow the reference's operation order with unfused arithmetic.
#include "common.cuh"

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
                                                    unsigned long long* __restrict__ packed) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const j3dg_pixel* p = px + (size_t)y * pstride + x;
  const uint32_t db = __ldg(reinterpret_cast<const uint32_t*>(p) + 7);
  const float depth = __ldg(reinterpret_cast<const float*>(p) + 3);
  const float z = db != 0u ? fdiv(1.f, depth) : 0.f;  // canvas.cpp:968
  packed[(size_t)y * w + x] = ((unsigned long long)__float_as_uint(z) << 32) | 0xFFFFFFFFull;
}

// _mm_cvtps_epi32 / cvttss2si semantics: NaN and out-of-range give INT_MIN
__device__ __forceinline__ int cvt_rne(float x) { return (fabsf(x) < 2147483648.f) ? __float2int_rn(x) : (int)0x80000000; }
__device__ __forceinline__ int cvt_trunc(float x) { return (fabsf(x) < 2147483648.f) ? __float2int_rz(x) : (int)0x80000000; }

__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ pos, SplatParams s, unsigned long long* __restrict__ packed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  const float x = __ldg(pos + 3 * (size_t)i), y = __ldg(pos + 3 * (size_t)i + 1), z = __ldg(pos + 3 * (size_t)i + 2);
  const float* M = s.M;
  int X, Y;
  float VW;
  if (i < s.tail_start) {  // SIMD body, render.h:419-468 + 726-732
    float VX = fadd(fadd(fadd(fmul(M[0], x), fmul(M[4], y)), fmul(M[8], z)), fmul(M[12], 1.f));
    float VY = fadd(fadd(fadd(fmul(M[1], x), fmul(M[5], y)), fmul(M[9], z)), fmul(M[13], 1.f));
    VW = fadd(fadd(fadd(fmul(M[3], x), fmul(M[7], y)), fmul(M[11], z)), fmul(M[15], 1.f));
    VX = fdiv(VX, VW); VY = fdiv(VY, VW);
    VX = fmul(fadd(VX, 1.f), fmul((float)s.w, 0.5f));
    VY = fmul(fadd(VY, 1.f), fmul((float)s.h, 0.5f));
    X = cvt_rne(VX); Y = cvt_rne(VY);
  } else {  // scalar tail, render.h:473-511 + 814-821
    float VX = fadd(fadd(fadd(fmul(M[0], x), fmul(M[4], y)), fmul(M[8], z)), M[12]);
    float VY = fadd(fadd(fadd(fmul(M[1], x), fmul(M[5], y)), fmul(M[9], z)), M[13]);
    float VZ = fadd(fadd(fadd(fmul(M[2], x), fmul(M[6], y)), fmul(M[10], z)), M[14]);
    VW = fadd(fadd(fadd(fmul(M[3], x), fmul(M[7], y)), fmul(M[11], z)), M[15]);
    VX = fdiv(VX, VW); VY = fdiv(VY, VW); VZ = fdiv(VZ, VW);
    if (VX < -1.f || VX > 1.f || VY < -1.f || VY > 1.f || VZ < -1.f || VZ > 1.f) return;  // vertex_clip_info
    VX = fmul(fmul(fadd(VX, 1.f), (float)s.w), 0.5f);
    VY = fmul(fmul(fadd(VY, 1.f), (float)s.h), 0.5f);
    X = cvt_trunc(VX); Y = cvt_trunc(VY);
  }
  if (X < 0 || Y < 0 || X > s.w - 1 || Y > s.h - 1) return;
  const float depth = fdiv(1.f, VW);  // render.h:774
  if (!(depth > 0.f)) return;          // the z-buffer is never negative: a point behind the eye can not pass `prev < depth`
  const unsigned long long word = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned long long)(0xFFFFFFFEu - i);
  unsigned long long* cell = packed + (size_t)Y * s.w + X;
  if (word > *((volatile unsigned long long*)cell)) atomicMax(cell, word);
}

__global__ void __launch_bounds__(256) resolve_kernel(SplatParams s, const float* __restrict__ nrm, const uint32_t* __restrict__ clr,
                                                       unsigned long long* __restrict__ packed, j3dg_pixel* __restrict__ px, uint32_t pstride,
                                                       uint32_t* __restrict__ rgba, uint32_t rstride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= s.w || y >= s.h) return;
  unsigned long long* cell = packed + (size_t)y * s.w + x;
  const unsigned long long word = *cell;
  const uint32_t low = (uint32_t)word;
  if (low == 0xFFFFFFFFu) return;  // no point in front of what was there
  const uint32_t i = 0xFFFFFFFEu - low;
  const float zb = __uint_as_float((uint32_t)(word >> 32));
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
  rgba[(size_t)y * rstride + x] = color;
  j3dg_pixel* p = px + (size_t)y * pstride + x;  // canvas.cpp:997-1027
  p->object_id = i;
  p->depth = fdiv(1.f, zb);
  p->db_id = s.db_id;
  *cell = (word & 0xFFFFFFFF00000000ull) | 0xFFFFFFFFull;  // becomes the depth the next cloud has to beat
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
  {
    void* p = ctx->d_packed;
    int rc = j3dg_reserve(ctx, &p, &ctx->packed_cap, sizeof(unsigned long long) * (size_t)w * h);
    ctx->d_packed = (unsigned long long*)p;
    if (rc != J3DG_OK) return rc;
  }
  dim3 pgrid((w + 31) / 32, (h + 7) / 8);
  { int rc = j3dg_stage_begin(ctx, 2); if (rc != J3DG_OK) return rc; }
  seed_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_px_in, pstride, w, h, ctx->d_packed);
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
      project_kernel<<<(cl->n + 255) / 256, 256, 0, ctx->stream>>>(cl->d_pos, s, ctx->d_packed);
      KERNEL_CHECK(ctx);
    }
    resolve_kernel<<<pgrid, 256, 0, ctx->stream>>>(s, cl->d_nrm, cl->d_clr, ctx->d_packed, d_px_inout, pstride, d_rgba, rstride);
    KERNEL_CHECK(ctx);
  }
  return j3dg_stage_end(ctx, 2);
}
