// j3dg_host.cpp — pure host math of the j3d interfaces above the GPU path: default camera and
// projection (j3d/camera.cpp), scene bbox / unzoom pose (j3d/scene.cpp:50-111), orbit pose
// (the composition canvas::do_mouse applies, j3d/canvas.cpp:195-210), procedural matcaps
// (j3d/matcap.cpp:9-268) and the background gradient (j3d/canvas.cpp:55-78).  These produce the
// *inputs* of the kernels (a j3dg_view, a 512x512 lookup table); they run once per frame or
// once per session on the host, as in the reference.  Built with -ffp-contract=off so every
// operation rounds once, like the reference's SSE build.
#include "j3dg_host.h"

#include <cfloat>
#include <cmath>

#define J3DGH_API extern "C" __attribute__((visibility("default")))

namespace {

struct vec4 { float x, y, z, w; };

inline vec4 col(const float* m, int c) { return {m[4 * c], m[4 * c + 1], m[4 * c + 2], m[4 * c + 3]}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

// jtk matrix_vector_multiply (qbvh.h:4564-4568): columns scaled then added left to right
inline vec4 mul(const float* m, vec4 v) { return col(m, 0) * v.x + col(m, 1) * v.y + col(m, 2) * v.z + col(m, 3) * v.w; }

struct matcap_params {
  uint32_t cavity;
  float base[3];       // additive constants
  float k1[3], k2[3];  // first two lobes; lobes 3 and 4 are (50,50,50) and (30,30,30) for all
};

const matcap_params kMatcaps[3] = {
    /* red wax */ {0xFF7D7DFF, {32.f, 0.f, 0.f}, {200.f / 1.5f, 200.f / 4.f, 150.f / 4.f}, {30.f, 25.f, 20.f}},
    /* gray    */ {0xff505050, {32.f, 32.f, 32.f}, {200.f / 4.f, 200.f / 4.f, 200.f / 4.f}, {50.f, 50.f, 50.f}},
    /* brown   */ {0xff405060, {32.f, 20.f, 10.f}, {200.f / 4.f, 180.f / 4.f, 160.f / 4.f}, {50.f, 40.f, 30.f}},
};

inline void unit(float v[3]) {
  float l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (l) { v[0] /= l; v[1] /= l; v[2] /= l; }
}

inline unsigned char to_byte(float c) {
  if (c > 255.f) c = 255.f;
  if (c < 0.f) c = 0.f;
  return (unsigned char)c;
}

}  // namespace

J3DGH_API void j3dgh_matrix_vector_multiply(const float* m, const float* v, float* out) {
  vec4 r = mul(m, {v[0], v[1], v[2], v[3]});
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

J3DGH_API void j3dgh_matrix_multiply(const float* a, const float* b, float* out) {
  float tmp[16];
  for (int c = 0; c < 4; ++c) j3dgh_matrix_vector_multiply(a, b + 4 * c, tmp + 4 * c);
  std::memcpy(out, tmp, sizeof(tmp));
}

// jtk invert_orthonormal (qbvh.h:4460-4467): rotation transposed, translation = -(R^T t)
J3DGH_API void j3dgh_invert_orthonormal(const float* m, float* out) {
  float r[16];
  vec4 c0 = {m[0], m[4], m[8], 0.f}, c1 = {m[1], m[5], m[9], 0.f}, c2 = {m[2], m[6], m[10], 0.f};
  vec4 t = c0 * m[12] + c1 * m[13] + c2 * m[14];
  const vec4 cols[4] = {c0, c1, c2, {-t.x, -t.y, -t.z, 1.f}};
  for (int c = 0; c < 4; ++c) { r[4 * c] = cols[c].x; r[4 * c + 1] = cols[c].y; r[4 * c + 2] = cols[c].z; r[4 * c + 3] = cols[c].w; }
  r[15] = 1.f;
  std::memcpy(out, r, sizeof(r));
}

// make_default_camera + make_projection_matrix + invert_projection_matrix (camera.cpp:5-73)
J3DGH_API void j3dgh_make_projection(uint32_t w, uint32_t h, float* near_plane, float* P, float* Pinv) {
  const float focal_length = 35.f, aperture_w_inch = 1.024f, aperture_h_inch = 0.768f;
  const float near_p = 0.1f, far_p = FLT_MAX, zoom = 1.f, inch_to_mm = 25.4f;
  float top = ((aperture_h_inch * inch_to_mm / 2.f) / focal_length) * near_p;
  float right = ((aperture_w_inch * inch_to_mm / 2.f) / focal_length) * near_p;
  float xs = zoom, ys = zoom;
  const float device_ratio = (int)w / (float)(int)h, film_ratio = aperture_w_inch / aperture_h_inch;
  if (film_ratio > device_ratio) ys *= film_ratio / device_ratio;  // Overscan gate fit
  else xs *= device_ratio / film_ratio;
  right *= xs;
  top *= ys;
  const float bottom = -top, left = -right;
  for (int i = 0; i < 16; ++i) P[i] = Pinv[i] = 0.f;
  P[0] = 2.f * near_p / (right - left);
  P[5] = -2.f * near_p / (top - bottom);
  P[8] = (right + left) / (right - left);
  P[9] = -(top + bottom) / (top - bottom);
  P[10] = -(far_p + near_p) / (far_p - near_p);
  P[11] = -1.f;
  P[14] = -(2.f * far_p * near_p) / (far_p - near_p);  // -inf: 2*FLT_MAX overflows, as in the reference
  Pinv[0] = 1.f / P[0];
  Pinv[5] = 1.f / P[5];
  Pinv[11] = 1.f / P[14];
  Pinv[12] = P[8] / P[0];
  Pinv[13] = P[9] / P[5];
  Pinv[14] = -1.f;
  Pinv[15] = P[10] / P[14];
  *near_plane = near_p;
}

J3DGH_API void j3dgh_compute_bb(const float* v, uint32_t nv, float* mn, float* mx) {  // mesh.cpp:39-58
  if (!nv) return;
  for (int j = 0; j < 3; ++j) mn[j] = mx[j] = v[j];
  for (uint32_t i = 1; i < nv; ++i)
    for (int j = 0; j < 3; ++j) {
      const float c = v[3 * (size_t)i + j];
      if (c < mn[j]) mn[j] = c;
      if (c > mx[j]) mx[j] = c;
    }
}

// prepare_scene's per-object step (scene.cpp:54-56, 73-77): both corners transformed as points
J3DGH_API void j3dgh_transform_bbox(const float* cs, const float* bb_min, const float* bb_max, float* out_min, float* out_max) {
  vec4 a = mul(cs, {bb_min[0], bb_min[1], bb_min[2], 1.f});
  vec4 b = mul(cs, {bb_max[0], bb_max[1], bb_max[2], 1.f});
  const float pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z};
  for (int j = 0; j < 3; ++j) {
    out_min[j] = pa[j] < pb[j] ? pa[j] : pb[j];
    out_max[j] = pa[j] > pb[j] ? pa[j] : pb[j];
  }
}

// scene.cpp:86-88 + unzoom 91-111 (y-up branch)
J3DGH_API void j3dgh_unzoom(const float* bb_min, const float* bb_max, float* diagonal, float* pivot, float* cs, float* cs_inv) {
  float d = bb_max[0] - bb_min[0];
  d = std::max(d, bb_max[1] - bb_min[1]);
  d = std::max(d, bb_max[2] - bb_min[2]);
  *diagonal = d;
  for (int i = 0; i < 16; ++i) cs[i] = (i % 5 == 0) ? 1.f : 0.f;
  for (int j = 0; j < 3; ++j) pivot[j] = (bb_min[j] + bb_max[j]) * 0.5f;
  cs[12] = pivot[0];
  cs[13] = pivot[1];
  cs[14] = pivot[2] + d * 2.f;
  j3dgh_invert_orthonormal(cs, cs_inv);
}

// Orbit step: CSinv_k = T(c) * R_y(angle) * T(-c) * CSinv_0 with c = CSinv_0 * pivot, the same
// composition a trackball drag applies (canvas.cpp:197-210); CS_k = invert_orthonormal(CSinv_k).
J3DGH_API void j3dgh_orbit(const float* cs_inv0, const float* pivot, float angle_deg, float* cs, float* cs_inv) {
  vec4 c = mul(cs_inv0, {pivot[0], pivot[1], pivot[2], 1.f});
  float t1[16], t2[16], rot[16];
  for (int i = 0; i < 16; ++i) t1[i] = t2[i] = rot[i] = (i % 5 == 0) ? 1.f : 0.f;
  t1[12] = c.x; t1[13] = c.y; t1[14] = c.z;
  t2[12] = -c.x; t2[13] = -c.y; t2[14] = -c.z;
  const double a = (double)angle_deg * 3.14159265358979323846 / 180.0;
  const float cs_a = (float)std::cos(a), sn_a = (float)std::sin(a);
  rot[0] = cs_a; rot[8] = sn_a; rot[2] = -sn_a; rot[10] = cs_a;
  float tmp1[16], tmp2[16], res[16];
  j3dgh_matrix_multiply(t2, cs_inv0, tmp1);
  j3dgh_matrix_multiply(t1, rot, tmp2);
  j3dgh_matrix_multiply(tmp2, tmp1, res);
  std::memcpy(cs_inv, res, sizeof(res));
  j3dgh_invert_orthonormal(cs_inv, cs);
}

// matcap.cpp:9-268.  The three lit matcaps are one formula with different constants:
// c = base + k1*d1 + k2*d2^3 + 50*d3^5 + 30*d3^50 over the (slightly inflated) unit disc.
J3DGH_API void j3dgh_make_matcap(int type, uint32_t* out, uint32_t* cavity) {
  const uint32_t w = 512, h = 512;
  if (type == 3) {  // sketch: ring falloff
    *cavity = 0xFF505050;
    for (uint32_t y = 0; y < h; ++y)
      for (uint32_t x = 0; x < w; ++x) {
        const float u = (float)x / (float)(w - 1) * 2.f - 1.f, v = (float)y / (float)(h - 1) * 2.f - 1.f;
        const float val = std::fabs(1.f - u * u - v * v);
        uint32_t c = 0xffe1e1e1;
        if (val < 0.4f) {
          const uint32_t s = (uint32_t)((val / 0.4f) * 0x000000e1);
          c = 0xff000000 | (s << 16) | (s << 8) | s;
        }
        out[(size_t)(h - y - 1) * w + x] = c;
      }
    return;
  }
  const matcap_params& mp = kMatcaps[(type == 1 || type == 2) ? type : 0];
  *cavity = mp.cavity;
  float l1[3] = {0, 0.8f, 1}, l2[3] = {0, 0.4f, 1}, l3[3] = {0, 0, 1};
  unit(l1); unit(l2); unit(l3);
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      const float u = (float)x / (float)(w - 1) * 2.f - 1.f, v = (float)y / (float)(h - 1) * 2.f - 1.f;
      uint32_t c = 0xff000000;
      if (u * u + v * v <= 1.01f) {
        const float lw = std::sqrt(1.01f - u * u - v * v);
        const float d1 = u * l1[0] + v * l1[1] + lw * l1[2];
        const float d2 = std::pow(u * l2[0] + v * l2[1] + lw * l2[2], 3.f);
        const float d3b = u * l3[0] + v * l3[1] + lw * l3[2];
        const float d3 = std::pow(d3b, 5.f), d4 = std::pow(d3b, 50.f);
        float rgb[3];
        for (int k = 0; k < 3; ++k) {
          const float lobes = (mp.k1[k] * d1 + mp.k2[k] * d2 + 50.f * d3 + 30.f * d4) / 1.f;
          // red wax adds no constant to g,b at all (matcap.cpp:207-208); adding +0.f is exact
          rgb[k] = (type != 1 && type != 2 && k > 0) ? lobes : mp.base[k] + lobes;
        }
        c = 0xff000000 | ((uint32_t)to_byte(rgb[2]) << 16) | ((uint32_t)to_byte(rgb[1]) << 8) | (uint32_t)to_byte(rgb[0]);
      }
      out[(size_t)(h - y - 1) * w + x] = c;
    }
}

// canvas.cpp:55-78
J3DGH_API void j3dgh_fill_background(uint32_t w, uint32_t h, uint32_t stride, uint32_t top, uint32_t bottom, uint32_t* out) {
  const uint32_t t[3] = {top & 0xff, (top >> 8) & 0xff, (top >> 16) & 0xff};
  const uint32_t b[3] = {bottom & 0xff, (bottom >> 8) & 0xff, (bottom >> 16) & 0xff};
  for (uint32_t y = 0; y < h; ++y) {
    const float s = (float)y / (float)h;
    uint32_t ch[3];
    for (int k = 0; k < 3; ++k) ch[k] = (uint32_t)(s * b[k] + (1.f - s) * t[k]) & 0xff;
    const uint32_t clr = 0xff000000 | (ch[2] << 16) | (ch[1] << 8) | ch[0];
    for (uint32_t x = 0; x < w; ++x) out[(size_t)y * stride + x] = clr;
  }
}
