/*
 * j3d_oracle.c — CPU restatement (plain C) of j3d's render hot path.
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path, never the thing shipped or
 * measured (see j3d_oracle.h for the pinning status).  Every function cites the reference
 * file:line (relative to the j3d tree) whose arithmetic it restates.  Compiled with
 * -ffp-contract=off and without FMA so each float operation rounds once, like the
 * reference's -msse4.1 build.  The closest hit is found through an own, deliberately
 * simple binary BVH (median split): parity needs the same closest hit, not the same tree.
 */
#include "j3d_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <emmintrin.h>

/* ------------------------------------------------------------------------------------
 * small math: jtk::float4x4 is column-major, m[col*4+row]
 * ---------------------------------------------------------------------------------- */

/* qbvh.h:4564-4568  out = c0*v0 + c1*v1 + c2*v2 + c3*v3, summed left to right per lane */
static void mat_vec(const float* m, const float v[4], float out[4])
{
  for (int r = 0; r < 4; ++r) {
    float a = m[r] * v[0];
    float b = m[4 + r] * v[1];
    float c = m[8 + r] * v[2];
    float d = m[12 + r] * v[3];
    out[r] = ((a + b) + c) + d;
  }
}

/* qbvh.h:4460-4467: transpose the 3x3, translation = -(c0'*m12 + c1'*m13 + c2'*m14) */
void orc_invert_orthonormal(const float m[16], float out[16])
{
  float c0[4] = {m[0], m[4], m[8], 0.f};
  float c1[4] = {m[1], m[5], m[9], 0.f};
  float c2[4] = {m[2], m[6], m[10], 0.f};
  for (int r = 0; r < 4; ++r) {
    out[r] = c0[r]; out[4 + r] = c1[r]; out[8 + r] = c2[r];
    out[12 + r] = -((c0[r] * m[12] + c1[r] * m[13]) + c2[r] * m[14]);
  }
  out[15] = 1.f;
}

/* qbvh.h matrix_matrix_multiply == column-wise mat_vec */
void orc_matrix_multiply(const float a[16], const float b[16], float out[16])
{
  float tmp[16];
  for (int c = 0; c < 4; ++c) mat_vec(a, b + 4 * c, tmp + 4 * c);
  memcpy(out, tmp, sizeof(tmp));
}

/* qbvh.h:4007-4014 reciprocal(): rcpps estimate + one Newton-Raphson step (0 -> estimate) */
static float reciprocal(float a)
{
  float res = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(a)));
  if (a == 0.f) return res;
  float muls = a * (res * res);
  return (res + res) - muls;
}

/* ------------------------------------------------------------------------------------
 * camera.cpp:5-73
 * ---------------------------------------------------------------------------------- */
void orc_make_projection(uint32_t w, uint32_t h, float* near_plane, float* P, float* Pinv)
{
  /* make_default_camera */
  const float focal = 35.f, film_w = 1.024f, film_h = 0.768f, nearp = 0.1f, farp = FLT_MAX, zoom = 1.f;
  const float inch_to_mm = 25.4f;
  float top = ((film_h * inch_to_mm / 2.f) / focal) * nearp;
  float right = ((film_w * inch_to_mm / 2.f) / focal) * nearp;
  float xscale = zoom, yscale = zoom;
  float device_aspect = (int)w / (float)(int)h;
  float film_aspect = film_w / film_h;
  /* Overscan */
  if (film_aspect > device_aspect) yscale *= film_aspect / device_aspect;
  else xscale *= device_aspect / film_aspect;
  right *= xscale;
  top *= yscale;
  float bottom = -top, left = -right;
  /* frustum(): camera.cpp:18-25 */
  memset(P, 0, 16 * sizeof(float));
  P[0] = 2.f * nearp / (right - left);
  P[5] = -2.f * nearp / (top - bottom);
  P[8] = (right + left) / (right - left);
  P[9] = -(top + bottom) / (top - bottom);
  P[10] = -(farp + nearp) / (farp - nearp);
  P[11] = -1.f;
  P[14] = -(2.f * farp * nearp) / (farp - nearp);
  /* invert_projection_matrix(): camera.cpp:27-31 */
  memset(Pinv, 0, 16 * sizeof(float));
  Pinv[0] = 1.f / P[0];
  Pinv[5] = 1.f / P[5];
  Pinv[11] = 1.f / P[14];
  Pinv[12] = P[8] / P[0];
  Pinv[13] = P[9] / P[5];
  Pinv[14] = -1.f;
  Pinv[15] = P[10] / P[14];
  *near_plane = nearp;
}

/* scene.cpp:86-88 (diagonal = largest extent) and 91-111 (unzoom, y-up branch) */
void orc_unzoom(const float bb_min[3], const float bb_max[3], float* diagonal, float pivot[3], float cs[16], float cs_inv[16])
{
  float d = bb_max[0] - bb_min[0];
  d = fmaxf(d, bb_max[1] - bb_min[1]);
  d = fmaxf(d, bb_max[2] - bb_min[2]);
  *diagonal = d;
  memset(cs, 0, 16 * sizeof(float));
  cs[0] = cs[5] = cs[10] = cs[15] = 1.f;
  for (int j = 0; j < 3; ++j) pivot[j] = (bb_min[j] + bb_max[j]) * 0.5f;
  cs[12] = pivot[0];
  cs[13] = pivot[1];
  cs[14] = pivot[2] + d * 2.f;
  orc_invert_orthonormal(cs, cs_inv);
}

/* ------------------------------------------------------------------------------------
 * matcap.cpp:9-268
 * ---------------------------------------------------------------------------------- */
static void normalize3(float v[3])
{ /* vec.h:531-536 */
  float denom = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  if (denom) { v[0] = v[0] / denom; v[1] = v[1] / denom; v[2] = v[2] / denom; }
}

void orc_make_matcap(int type, uint32_t* out, uint32_t* cavity)
{
  const uint32_t w = 512, h = 512;
  if (type == 3) { /* make_matcap_sketch, matcap.cpp:241-268 */
    *cavity = 0xFF505050;
    const float thres = 0.4f;
    for (uint32_t y = 0; y < h; ++y) {
      uint32_t* row = out + (size_t)(h - y - 1) * w;
      for (uint32_t x = 0; x < w; ++x) {
        float u = (float)x / (float)(w - 1) * 2.f - 1.f;
        float v = (float)y / (float)(h - 1) * 2.f - 1.f;
        float val = fabsf(1.f - u * u - v * v);
        if (val < thres) {
          float scale = val / thres;
          uint32_t s = (uint32_t)(scale * 0x000000e1);
          row[x] = 0xff000000 | (s << 16) | (s << 8) | s;
        } else row[x] = 0xffe1e1e1;
      }
    }
    return;
  }
  /* gray (9-85), brown (88-164), red wax (166-239) share one loop, differing in constants */
  float r1, g1, b1, r2, g2, b2, r0, g0, b0;
  if (type == 1) { *cavity = 0xff505050; r1 = g1 = b1 = 200.f / 4.f; r2 = g2 = b2 = 50.f; r0 = g0 = b0 = 32.f; }
  else if (type == 2) { *cavity = 0xff405060; r1 = 200.f / 4.f; g1 = 180.f / 4.f; b1 = 160.f / 4.f; r2 = 50.f; g2 = 40.f; b2 = 30.f; r0 = 32.f; g0 = 20.f; b0 = 10.f; }
  else { *cavity = 0xFF7D7DFF; r1 = 200.f / 1.5f; g1 = 200.f / 4.f; b1 = 150.f / 4.f; r2 = 30.f; g2 = 25.f; b2 = 20.f; r0 = 32.f; g0 = 0.f; b0 = 0.f; }
  const float r3 = 50.f, g3 = 50.f, b3 = 50.f, r4 = 30.f, g4 = 30.f, b4 = 30.f;
  float n1[3] = {0, 0.8f, 1}, n2[3] = {0, 0.4f, 1}, n3[3] = {0, 0, 1};
  normalize3(n1); normalize3(n2); normalize3(n3);
  for (uint32_t y = 0; y < h; ++y) {
    uint32_t* row = out + (size_t)(h - y - 1) * w;
    for (uint32_t x = 0; x < w; ++x) {
      float u = (float)x / (float)(w - 1) * 2.f - 1.f;
      float v = (float)y / (float)(h - 1) * 2.f - 1.f;
      if (u * u + v * v <= 1.01f) {
        float lw = sqrtf(1.01f - u * u - v * v);
        float d1 = (u * n1[0] + v * n1[1]) + lw * n1[2];
        float d2 = (u * n2[0] + v * n2[1]) + lw * n2[2];
        float d3 = (u * n3[0] + v * n3[1]) + lw * n3[2];
        float ca1 = d1, ca2 = powf(d2, 3.f), ca3 = powf(d3, 5.f), ca4 = powf(d3, 50.f);
        float r, g, b;
        r = r0 + (r1 * ca1 + r2 * ca2 + r3 * ca3 + r4 * ca4) / 1.f;
        if (type == 0) { /* red wax has no additive constant on g,b (matcap.cpp:207-208) */
          g = (g1 * ca1 + g2 * ca2 + g3 * ca3 + g4 * ca4) / 1.f;
          b = (b1 * ca1 + b2 * ca2 + b3 * ca3 + b4 * ca4) / 1.f;
        } else {
          g = g0 + (g1 * ca1 + g2 * ca2 + g3 * ca3 + g4 * ca4) / 1.f;
          b = b0 + (b1 * ca1 + b2 * ca2 + b3 * ca3 + b4 * ca4) / 1.f;
        }
        if (r > 255.f) r = 255.f; if (r < 0.f) r = 0.f;
        if (g > 255.f) g = 255.f; if (g < 0.f) g = 0.f;
        if (b > 255.f) b = 255.f; if (b < 0.f) b = 0.f;
        unsigned char red = (unsigned char)r, green = (unsigned char)g, blue = (unsigned char)b;
        row[x] = 0xff000000 | ((uint32_t)blue << 16) | ((uint32_t)green << 8) | (uint32_t)red;
      } else row[x] = 0xff000000;
    }
  }
}

/* canvas.cpp:55-78 */
void orc_fill_background(uint32_t w, uint32_t h, uint32_t top, uint32_t bottom, uint32_t* out)
{
  uint32_t tr = top & 0xff, tg = (top >> 8) & 0xff, tb = (top >> 16) & 0xff;
  uint32_t br = bottom & 0xff, bg = (bottom >> 8) & 0xff, bb = (bottom >> 16) & 0xff;
  for (uint32_t y = 0; y < h; ++y) {
    float scale = (float)y / (float)h;
    uint32_t red = (uint32_t)(scale * br + (1.f - scale) * tr);
    uint32_t green = (uint32_t)(scale * bg + (1.f - scale) * tg);
    uint32_t blue = (uint32_t)(scale * bb + (1.f - scale) * tb);
    uint32_t clr = 0xff000000 | ((uint32_t)(unsigned char)blue << 16) | ((uint32_t)(unsigned char)green << 8) | (uint32_t)(unsigned char)red;
    for (uint32_t x = 0; x < w; ++x) out[(size_t)y * w + x] = clr;
  }
}

/* ------------------------------------------------------------------------------------
 * mesh + own binary BVH
 * ---------------------------------------------------------------------------------- */
typedef struct { float mn[3], mx[3]; uint32_t left, right; uint32_t first, count; } bnode; /* count>0 => leaf */

struct orc_mesh {
  uint32_t nv, nt;
  float* verts; uint32_t* tris; float* normals;
  float* vcolors; float* uv; uint32_t* tex; uint32_t tw, th;
  float cs[16], cs_inv[16];
  uint32_t db_id;
  float bb_min[3], bb_max[3];
  bnode* nodes; uint32_t nnodes; uint32_t* order;
};

typedef struct { float c[3]; uint32_t id; } centroid_t;
static int g_axis;
static int cmp_centroid(const void* a, const void* b)
{
  float x = ((const centroid_t*)a)->c[g_axis], y = ((const centroid_t*)b)->c[g_axis];
  return (x > y) - (x < y);
}

static void tri_bounds(const orc_mesh* m, uint32_t t, float mn[3], float mx[3])
{
  for (int j = 0; j < 3; ++j) { mn[j] = FLT_MAX; mx[j] = -FLT_MAX; }
  for (int k = 0; k < 3; ++k) {
    const float* p = m->verts + 3 * (size_t)m->tris[3 * (size_t)t + k];
    for (int j = 0; j < 3; ++j) { if (p[j] < mn[j]) mn[j] = p[j]; if (p[j] > mx[j]) mx[j] = p[j]; }
  }
}

static uint32_t build_rec(orc_mesh* m, centroid_t* c, uint32_t first, uint32_t count)
{
  uint32_t idx = m->nnodes++;
  bnode* n = &m->nodes[idx];
  for (int j = 0; j < 3; ++j) { n->mn[j] = FLT_MAX; n->mx[j] = -FLT_MAX; }
  float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i = first; i < first + count; ++i) {
    float a[3], b[3];
    tri_bounds(m, c[i].id, a, b);
    for (int j = 0; j < 3; ++j) {
      if (a[j] < n->mn[j]) n->mn[j] = a[j];
      if (b[j] > n->mx[j]) n->mx[j] = b[j];
      if (c[i].c[j] < cmn[j]) cmn[j] = c[i].c[j];
      if (c[i].c[j] > cmx[j]) cmx[j] = c[i].c[j];
    }
  }
  if (count <= 4) {
    n->first = first; n->count = count; n->left = n->right = 0;
    for (uint32_t i = 0; i < count; ++i) m->order[first + i] = c[first + i].id;
    return idx;
  }
  int axis = 0;
  if (cmx[1] - cmn[1] > cmx[axis] - cmn[axis]) axis = 1;
  if (cmx[2] - cmn[2] > cmx[axis] - cmn[axis]) axis = 2;
  g_axis = axis;
  qsort(c + first, count, sizeof(centroid_t), cmp_centroid);
  uint32_t half = count / 2;
  n->count = 0; n->first = 0;
  uint32_t l = build_rec(m, c, first, half);
  uint32_t r = build_rec(m, c, first + half, count - half);
  m->nodes[idx].left = l; m->nodes[idx].right = r;
  return idx;
}

orc_mesh* orc_mesh_create(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt,
                          const float* vcolors, const float* uv, const uint32_t* tex, uint32_t tw, uint32_t th,
                          const float* cs, uint32_t db_id)
{
  orc_mesh* m = (orc_mesh*)calloc(1, sizeof(orc_mesh));
  m->nv = nv; m->nt = nt; m->db_id = db_id;
  m->verts = (float*)malloc(sizeof(float) * 3 * (size_t)(nv ? nv : 1)); memcpy(m->verts, verts, sizeof(float) * 3 * (size_t)nv);
  m->tris = (uint32_t*)malloc(sizeof(uint32_t) * 3 * (size_t)(nt ? nt : 1)); memcpy(m->tris, tris, sizeof(uint32_t) * 3 * (size_t)nt);
  if (vcolors) { m->vcolors = (float*)malloc(sizeof(float) * 3 * (size_t)nv); memcpy(m->vcolors, vcolors, sizeof(float) * 3 * (size_t)nv); }
  if (uv && tex) {
    m->uv = (float*)malloc(sizeof(float) * 6 * (size_t)nt); memcpy(m->uv, uv, sizeof(float) * 6 * (size_t)nt);
    m->tex = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)tw * th); memcpy(m->tex, tex, sizeof(uint32_t) * (size_t)tw * th);
    m->tw = tw; m->th = th;
  }
  memset(m->cs, 0, sizeof(m->cs)); m->cs[0] = m->cs[5] = m->cs[10] = m->cs[15] = 1.f;
  if (cs) memcpy(m->cs, cs, sizeof(m->cs));
  orc_invert_orthonormal(m->cs, m->cs_inv); /* canvas.cpp:734 */
  /* compute_triangle_normals, geometry.h:2561-2578 + vec.h:477-487, 531-536 */
  m->normals = (float*)malloc(sizeof(float) * 3 * (size_t)(nt ? nt : 1));
  for (uint32_t t = 0; t < nt; ++t) {
    const float* V0 = m->verts + 3 * (size_t)tris[3 * (size_t)t];
    const float* V1 = m->verts + 3 * (size_t)tris[3 * (size_t)t + 1];
    const float* V2 = m->verts + 3 * (size_t)tris[3 * (size_t)t + 2];
    float l[3] = {V1[0] - V0[0], V1[1] - V0[1], V1[2] - V0[2]};
    float r[3] = {V2[0] - V0[0], V2[1] - V0[1], V2[2] - V0[2]};
    float n[3] = {l[1] * r[2] - l[2] * r[1], l[2] * r[0] - l[0] * r[2], l[0] * r[1] - l[1] * r[0]};
    normalize3(n);
    memcpy(m->normals + 3 * (size_t)t, n, sizeof(n));
  }
  /* compute_bb, mesh.cpp:39-58 */
  for (int j = 0; j < 3; ++j) { m->bb_min[j] = nv ? verts[j] : 0.f; m->bb_max[j] = nv ? verts[j] : 0.f; }
  for (uint32_t i = 1; i < nv; ++i)
    for (int j = 0; j < 3; ++j) {
      if (verts[3 * (size_t)i + j] < m->bb_min[j]) m->bb_min[j] = verts[3 * (size_t)i + j];
      if (verts[3 * (size_t)i + j] > m->bb_max[j]) m->bb_max[j] = verts[3 * (size_t)i + j];
    }
  if (nt) {
    centroid_t* c = (centroid_t*)malloc(sizeof(centroid_t) * (size_t)nt);
    for (uint32_t t = 0; t < nt; ++t) {
      float a[3], b[3];
      tri_bounds(m, t, a, b);
      for (int j = 0; j < 3; ++j) c[t].c[j] = 0.5f * (a[j] + b[j]);
      c[t].id = t;
    }
    m->nodes = (bnode*)malloc(sizeof(bnode) * 2 * (size_t)nt);
    m->order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nt);
    m->nnodes = 0;
    build_rec(m, c, 0, nt);
    free(c);
  }
  return m;
}

void orc_mesh_destroy(orc_mesh* m)
{
  if (!m) return;
  free(m->verts); free(m->tris); free(m->normals); free(m->vcolors); free(m->uv); free(m->tex); free(m->nodes); free(m->order);
  free(m);
}

const float* orc_mesh_normals(const orc_mesh* m) { return m->normals; }
void orc_mesh_bbox(const orc_mesh* m, float mn[3], float mx[3]) { memcpy(mn, m->bb_min, 12); memcpy(mx, m->bb_max, 12); }

/* ------------------------------------------------------------------------------------
 * Woop intersection, qbvh.h:4793-4869 (one lane of the 4-wide code)
 * ---------------------------------------------------------------------------------- */
typedef struct { int kx, ky, kz; float Sx, Sy, Sz; } woop_pre;

static woop_pre woop_precompute(const float d[3])
{
  static const int modulo[5] = {0, 1, 2, 0, 1};
  woop_pre o;
  float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
  o.kz = 2;
  if (ax > ay) { if (ax > az) o.kz = 0; }
  else { if (ay > az) o.kz = 1; }
  o.kx = modulo[o.kz + 1];
  o.ky = modulo[o.kx + 1];
  if (d[o.kz] < 0.f) { int t = o.kx; o.kx = o.ky; o.ky = t; }
  o.Sz = 1.f / d[o.kz];
  o.Sx = d[o.kx] * o.Sz;
  o.Sy = d[o.ky] * o.Sz;
  return o;
}

/* returns found; t,u,v as the reference computes them */
static int woop_intersect(const float* v0, const float* v1, const float* v2, const woop_pre* p, const float org[3],
                          float t_near, float t_far, float* t_out, float* u_out, float* v_out)
{
  float A[3] = {v0[0] - org[0], v0[1] - org[1], v0[2] - org[2]};
  float B[3] = {v1[0] - org[0], v1[1] - org[1], v1[2] - org[2]};
  float C[3] = {v2[0] - org[0], v2[1] - org[1], v2[2] - org[2]};
  const float Ax = A[p->kx] - p->Sx * A[p->kz];
  const float Ay = A[p->ky] - p->Sy * A[p->kz];
  const float Bx = B[p->kx] - p->Sx * B[p->kz];
  const float By = B[p->ky] - p->Sy * B[p->kz];
  const float Cx = C[p->kx] - p->Sx * C[p->kz];
  const float Cy = C[p->ky] - p->Sy * C[p->kz];
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  int found = ((U <= 0.f) && (V <= 0.f) && (W <= 0.f)) || ((U >= 0.f) && (V >= 0.f) && (W >= 0.f));
  if (!found) return 0;
  const float det = (U + V) + W;
  if (!(det != 0.f)) return 0;
  const float inv_det = reciprocal(det);
  const float Az = p->Sz * A[p->kz];
  const float Bz = p->Sz * B[p->kz];
  const float Cz = p->Sz * C[p->kz];
  const float T = (U * Az + V * Bz) + W * Cz;
  const float t = T * inv_det;
  if (!((t_far > t) && (t > t_near))) return 0;
  *t_out = t; *u_out = V * inv_det; *v_out = W * inv_det;
  return 1;
}

typedef struct { float u, v, distance; int found; } orc_hit;

/* qbvh::find_closest_triangle, qbvh.h:1701-1852: strict |t| < |best|, bounds shrink on accept.
 * The box test below is this file's own (conservative) slab test, not the reference's. */
static orc_hit mesh_find_closest(const orc_mesh* m, const float org[3], const float dir[3], float t_near, float t_far, uint32_t* tri_id)
{
  orc_hit h; h.found = 0; h.distance = FLT_MAX; h.u = h.v = 0.f;
  if (!m->nt) return h;
  woop_pre pre = woop_precompute(dir);
  double inv[3];
  for (int j = 0; j < 3; ++j) inv[j] = 1.0 / (double)dir[j];
  uint32_t stack[128]; int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const bnode* n = &m->nodes[stack[--sp]];
    double lo = (double)t_near, hi = (double)t_far;
    int miss = 0;
    for (int j = 0; j < 3 && !miss; ++j) {
      if (dir[j] == 0.f) { if (org[j] < n->mn[j] || org[j] > n->mx[j]) miss = 1; continue; }
      double a = ((double)n->mn[j] - (double)org[j]) * inv[j], b = ((double)n->mx[j] - (double)org[j]) * inv[j];
      if (a > b) { double t = a; a = b; b = t; }
      double pad = 1e-6 * (fabs(a) + fabs(b)) + 1e-30;
      a -= pad; b += pad;
      if (a > lo) lo = a;
      if (b < hi) hi = b;
      if (lo > hi) miss = 1;
    }
    if (miss) continue;
    if (n->count) {
      for (uint32_t i = 0; i < n->count; ++i) {
        uint32_t t = m->order[n->first + i];
        const uint32_t* tr = m->tris + 3 * (size_t)t;
        float tt, uu, vv;
        if (woop_intersect(m->verts + 3 * (size_t)tr[0], m->verts + 3 * (size_t)tr[1], m->verts + 3 * (size_t)tr[2], &pre, org, t_near, t_far, &tt, &uu, &vv)) {
          if (fabsf(tt) < fabsf(h.distance)) {
            h.found = -1; h.distance = tt; h.u = uu; h.v = vv; *tri_id = t;
            if (h.distance > 0) t_far = h.distance; else t_near = h.distance;
          }
        }
      }
    } else {
      stack[sp++] = n->left; stack[sp++] = n->right;
    }
  }
  return h;
}

void orc_find_closest(const orc_mesh* m, const float* rays, uint32_t n, float* hits, uint32_t* ids)
{
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t id = 0xffffffffu;
    orc_hit h = mesh_find_closest(m, rays + 8 * (size_t)i, rays + 8 * (size_t)i + 3, rays[8 * (size_t)i + 6], rays[8 * (size_t)i + 7], &id);
    hits[4 * (size_t)i + 0] = h.found ? h.u : 0.f;
    hits[4 * (size_t)i + 1] = h.found ? h.v : 0.f;
    hits[4 * (size_t)i + 2] = h.distance;
    hits[4 * (size_t)i + 3] = h.found ? 1.f : 0.f;
    ids[i] = h.found ? id : 0xffffffffu;
  }
}

/* qbvh_two_level_with_transformations::find_closest_triangle, qbvh.h:3303-3387: per object the
 * ray is transformed by the inverted object matrix, the closest |t| over all objects wins. */
static orc_hit scene_find_closest(orc_mesh* const* meshes, uint32_t nm, const float org[4], const float dir[4],
                                  float t_near, float t_far, uint32_t* tri_id, uint32_t* obj)
{
  orc_hit h; h.found = 0; h.distance = FLT_MAX; h.u = h.v = 0.f;
  for (uint32_t o = 0; o < nm; ++o) {
    float d2[4], o2[4];
    mat_vec(meshes[o]->cs_inv, dir, d2);
    mat_vec(meshes[o]->cs_inv, org, o2);
    uint32_t cand = 0;
    orc_hit lh = mesh_find_closest(meshes[o], o2, d2, t_near, t_far, &cand);
    if (lh.found && fabsf(lh.distance) < fabsf(h.distance)) {
      h = lh; *tri_id = cand; *obj = o;
      if (h.distance > 0) t_far = h.distance; else t_near = h.distance;
    }
  }
  return h;
}

/* jtk::transform(float4x4, float4), qbvh.h:5140-5151 */
static void transform_point(const float* m, const float p[4], float out[4])
{
  mat_vec(m, p, out);
  if (out[3] != 1.f && out[3]) { out[0] /= out[3]; out[1] /= out[3]; out[2] /= out[3]; out[3] = 1.f; }
}

/* canvas::update_canvas, canvas.cpp:677-874 */
void orc_cast(orc_mesh* const* meshes, uint32_t nm, const j3dg_view* vw, int x0, int y0, int x1, int y1,
              j3dg_pixel* out, uint32_t stride)
{
  const uint32_t w = vw->width, h = vw->height;
  if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 < 0) x1 = 0; if (y1 < 0) y1 = 0;
  if (x0 >= (int)w) x0 = w - 1; if (y0 >= (int)h) y0 = h - 1; if (x1 >= (int)w) x1 = w - 1; if (y1 >= (int)h) y1 = h - 1;
  const float o4[4] = {0.f, 0.f, 0.f, 1.f};
  float origin[4];
  mat_vec(vw->cs, o4, origin);
  float light0[4] = {vw->pivot[0] + vw->diagonal * 3.f, vw->pivot[1] + vw->diagonal * 3.f, vw->pivot[2] + vw->diagonal * 3.f, 1.f};
  float light[4];
  mat_vec(vw->cs, light0, light);
  for (int y = y0; y <= y1; ++y) {
    for (int x = x0; x <= x1; ++x) {
      j3dg_pixel* p = out + (size_t)y * stride + x;
      float sp[4] = {2.f * ((x + 0.5f) / w) - 1.f, 2.f * ((y + 0.5f) / h) - 1.f, vw->near_plane, 1.f};
      float dir[4];
      mat_vec(vw->projection_inv, sp, dir);
      dir[3] = 0.f;
      float d2[4];
      mat_vec(vw->cs, dir, d2);
      uint32_t tri = 0, obj = 0;
      orc_hit hit = nm ? scene_find_closest(meshes, nm, origin, d2, vw->diagonal / 100.f, FLT_MAX, &tri, &obj) : (orc_hit){0, 0, FLT_MAX, 0};
      if (hit.found) {
        const orc_mesh* m = meshes[obj];
        float n[4] = {m->normals[3 * (size_t)tri], m->normals[3 * (size_t)tri + 1], m->normals[3 * (size_t)tri + 2], 0.f};
        float n1[4], n2[4];
        mat_vec(vw->cs_inv, n, n1);
        mat_vec(m->cs, n1, n2);
        p->u = n2[0]; p->v = n2[1];
        p->depth = hit.distance;
        p->object_id = tri;
        p->barycentric_u = hit.u; p->barycentric_v = hit.v;
        p->db_id = m->db_id;
        p->mark = 0;
        p->r = p->g = p->b = 0; /* the reference leaves r,g,b stale here; 0 = state after resize() */
        const uint32_t* tr = m->tris + 3 * (size_t)tri;
        if ((vw->flags & J3DG_TEXTURED) && m->uv) {
          const float* uvc = m->uv + 6 * (size_t)tri;
          float k = 1.f - hit.u - hit.v;
          float cx = (k * uvc[0] + hit.u * uvc[2]) + hit.v * uvc[4];
          float cy = (k * uvc[1] + hit.u * uvc[3]) + hit.v * uvc[5];
          cx = fmaxf(fminf(cx, 1.f), 0.f); cy = fmaxf(fminf(cy, 1.f), 0.f);
          int tw = (int)m->tw, th = (int)m->th;
          int X = (int)(cx * tw), Y = (int)(cy * th);
          X = X < 0 ? 0 : X >= tw ? tw - 1 : X;
          Y = Y < 0 ? 0 : Y >= th ? th - 1 : Y;
          uint32_t color = m->tex[(size_t)Y * m->tw + X];
          p->r = color & 0xff; p->g = (color >> 8) & 0xff; p->b = (color >> 16) & 0xff;
          p->mark |= 2;
        } else if ((vw->flags & J3DG_VERTEXCOLORS) && m->vcolors) {
          const float* c0 = m->vcolors + 3 * (size_t)tr[0];
          const float* c1 = m->vcolors + 3 * (size_t)tr[1];
          const float* c2 = m->vcolors + 3 * (size_t)tr[2];
          float k = 1.f - hit.u - hit.v;
          float c[3];
          for (int j = 0; j < 3; ++j) c[j] = (c0[j] * k + hit.u * c1[j]) + hit.v * c2[j];
          p->r = (uint8_t)(c[0] * 255.f); p->g = (uint8_t)(c[1] * 255.f); p->b = (uint8_t)(c[2] * 255.f);
          p->mark |= 2;
        }
        if (vw->flags & J3DG_SHADOW) {
          float V0[4] = {m->verts[3 * (size_t)tr[0]], m->verts[3 * (size_t)tr[0] + 1], m->verts[3 * (size_t)tr[0] + 2], 1.f};
          float V1[4] = {m->verts[3 * (size_t)tr[1]], m->verts[3 * (size_t)tr[1] + 1], m->verts[3 * (size_t)tr[1] + 2], 1.f};
          float V2[4] = {m->verts[3 * (size_t)tr[2]], m->verts[3 * (size_t)tr[2] + 1], m->verts[3 * (size_t)tr[2] + 2], 1.f};
          float T0[4], T1[4], T2[4];
          transform_point(m->cs, V0, T0); transform_point(m->cs, V1, T1); transform_point(m->cs, V2, T2);
          float k = 1.f - hit.u - hit.v;
          float pos[4], ld[4];
          for (int j = 0; j < 4; ++j) pos[j] = (T0[j] * k + hit.u * T1[j]) + hit.v * T2[j];
          for (int j = 0; j < 4; ++j) ld[j] = light[j] - pos[j];
          uint32_t t2 = 0, o2 = 0;
          orc_hit h2 = scene_find_closest(meshes, nm, pos, ld, 1e-3f, FLT_MAX, &t2, &o2);
          if (h2.found) p->mark |= 1;
        }
      } else {
        /* canvas.cpp:859-866; the remaining fields keep their previous content in the
         * reference — here the post-resize() state (all zero) */
        memset(p, 0, sizeof(*p));
        p->db_id = 0; p->object_id = 0xffffffffu; p->u = 0.f; p->v = 0.f; p->depth = FLT_MAX;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * shading, canvas.cpp:287-670
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const j3dg_view* vw; const uint32_t* matcap; uint32_t mw, mh, cavity;
} shade_ctx;

static uint32_t get_U(float u, uint32_t mw) { return (uint32_t)floorf(0.5f + (u + 1.f) * (mw - 1) * 0.5f); }    /* canvas.cpp:320-323 */
static uint32_t get_V(float v, uint32_t mh) { return (uint32_t)floorf(0.5f + (-v + 1.f) * (mh - 1) * 0.5f); }   /* canvas.cpp:325-328 */

static uint32_t get_color(const shade_ctx* s, uint32_t U, uint32_t V, uint32_t shadow)
{ /* canvas.cpp:330-344 */
  uint32_t clr = s->matcap[(size_t)V * s->mw + U];
  if (shadow) {
    uint32_t r = (clr & 0xff) >> 2, g = ((clr >> 8) & 0xff) >> 2, b = ((clr >> 16) & 0xff) >> 2;
    clr = 0xff000000 | (b << 16) | (g << 8) | r;
  }
  return clr;
}

static uint32_t get_angle_color(const shade_ctx* s, float angle, float u, float v, uint32_t mark)
{ /* canvas.cpp:346-389 */
  int U = get_U(u, s->mw), V = get_V(v, s->mh);
  uint32_t clr = get_color(s, U, V, mark);
  if (fabsf(angle) <= 1.f) {
    uint32_t r = clr & 0xff, g = (clr >> 8) & 0xff, b = (clr >> 16) & 0xff;
    /* `acos` is called unqualified on a float in a TU that only sees ::acos(double), so the
     * whole expression is evaluated in double and rounded to float once (canvas.cpp:357) */
    float scale = (float)(((double)1.57079632679489f - fabs(acos((double)angle) - (double)1.57079632679489f)) / (double)1.57079632679489f);
    scale = sqrtf(1.f - scale);
    uint32_t r2 = s->cavity & 0xff, g2 = (s->cavity >> 8) & 0xff, b2 = (s->cavity >> 16) & 0xff;
    float k = angle > 0.f ? 1.2f : 0.5f;
    r = (uint32_t)(r * (1.f - scale) + r2 * scale * k);
    g = (uint32_t)(g * (1.f - scale) + g2 * scale * k);
    b = (uint32_t)(b * (1.f - scale) + b2 * scale * k);
    if (r > 255) r = 255; if (g > 255) g = 255; if (b > 255) b = 255;
    clr = 0xff000000 | (b << 16) | (g << 8) | r;
  }
  return clr;
}

static float compute_convex_cos_angle(const shade_ctx* s, float x1, float y1, float u1, float v1, float depth1,
                                      float x2, float y2, float u2, float v2, float depth2)
{ /* canvas.cpp:287-315 */
  const j3dg_view* vw = s->vw;
  float sp1[4] = {2.f * ((x1 + 0.5f) / vw->width) - 1.f, 2.f * ((y1 + 0.5f) / vw->height) - 1.f, vw->near_plane, 1.f};
  float dir1[4]; mat_vec(vw->projection_inv, sp1, dir1); dir1[3] = 0.f;
  float pt1[4]; for (int j = 0; j < 4; ++j) pt1[j] = depth1 * dir1[j];
  float sp2[4] = {2.f * ((x2 + 0.5f) / vw->width) - 1.f, 2.f * ((y2 + 0.5f) / vw->height) - 1.f, vw->near_plane, 1.f};
  float dir2[4]; mat_vec(vw->projection_inv, sp2, dir2); dir2[3] = 0.f;
  float pt2[4]; for (int j = 0; j < 4; ++j) pt2[j] = depth2 * dir2[j];
  float n1[3] = {u1, v1, sqrtf(1.f - u1 * u1 - v1 * v1)};
  float n2[3] = {u2, v2, sqrtf(1.f - u2 * u2 - v2 * v2)};
  float d = (n1[0] * n2[0] + n1[1] * n2[1]) + n1[2] * n2[2]; /* _mm_dp_ps 0x7F */
  float angle;
  if (fabsf(d - 1.f) > 0.0001) { /* double literal in the reference */
    float pt[3] = {pt2[0] - pt1[0], pt2[1] - pt1[1], pt2[2] - pt1[2]};
    float l = sqrtf((pt[0] * pt[0] + pt[1] * pt[1]) + pt[2] * pt[2]);
    pt[0] = pt[0] / l; pt[1] = pt[1] / l; pt[2] = pt[2] / l;
    angle = (pt[0] * n1[0] + pt[1] * n1[1]) + pt[2] * n1[2];
  } else angle = 0.f;
  return angle;
}

static float clampf(float x, float a, float b) { return x < a ? a : (x > b ? b : x); }

static uint32_t plain_color(const shade_ctx* s, const j3dg_pixel* p)
{ /* canvas::_get_color, canvas.cpp:410-446 */
  uint32_t clr;
  if (p->mark & 2) {
    if (s->vw->flags & J3DG_SHADING) {
      float u = p->u, v = p->v;
      float nz = sqrtf(1.f - u * u - v * v);
      float occ = (p->mark & 1) ? 0.3f : 1.f;
      float dot = (u * 0.f + v * 0.f) + nz * 1.f;
      float dif = clampf(dot, 0.f, 1.f) * occ;
      clr = 0xff000000 | ((uint32_t)(p->b * dif) << 16) | ((uint32_t)(p->g * dif) << 8) | ((uint32_t)(p->r * dif));
    } else {
      if (p->mark & 1) clr = 0xff000000 | ((uint32_t)(p->b >> 2) << 16) | ((uint32_t)(p->g >> 2) << 8) | ((uint32_t)(p->r >> 2));
      else clr = 0xff000000 | ((uint32_t)p->b << 16) | ((uint32_t)p->g << 8) | ((uint32_t)p->r);
    }
  } else {
    clr = get_color(s, get_U(p->u, s->mw), get_V(p->v, s->mh), p->mark);
  }
  return clr;
}

#define IS_HIT(p) ((p)->object_id != 0xffffffffu)
#define NDIFF(a, b, thr) ((fabsf((a)->u - (b)->u) > (thr)) || (fabsf((a)->v - (b)->v) > (thr)))

void orc_shade(const j3dg_pixel* px, uint32_t stride, const j3dg_view* vw, const uint32_t* matcap,
               uint32_t mw, uint32_t mh, uint32_t cavity, uint32_t* rgba, uint32_t rgba_stride)
{
  shade_ctx s = {vw, matcap, mw, mh, cavity};
  const uint32_t w = vw->width, h = vw->height;
  const float threshold = 0.001f;
  for (uint32_t y = 0; y < h; ++y) {
    for (uint32_t x = 0; x < w; ++x) {
      const j3dg_pixel* p = px + (size_t)y * stride + x;
      if (!IS_HIT(p)) continue;
      const j3dg_pixel* right = p + 1;
      const j3dg_pixel* up = px + (size_t)(y ? y - 1 : 0) * stride + x; /* row 0 compares with itself */
      uint32_t* o = rgba + (size_t)y * rgba_stride + x;
      const int last = (x == w - 1);
      if (vw->flags & J3DG_ONE_BIT) { /* canvas.cpp:498-579 */
        uint32_t clr = get_color(&s, get_U(p->u, mw), get_V(p->v, mh), p->mark);
        int res = (((clr & 0xff0000) >> 16) + ((clr & 0xff00) >> 8) + (clr & 0xff)) >> 7;
        ++res;
        if (last) { *o = ((((w - 1) % res) == 0) && ((y % res) == 0)) ? 0xff000000 : 0xffffffff; continue; }
        float angle = 1.f;
        const j3dg_pixel* q = 0;
        if (IS_HIT(right) && NDIFF(p, right, threshold)) q = right;
        else if (IS_HIT(up) && NDIFF(p, up, threshold)) q = up;
        if (q) {
          float w1 = sqrtf(1.f - p->u * p->u - p->v * p->v), w2 = sqrtf(1.f - q->u * q->u - q->v * q->v);
          angle = p->u * q->u + p->v * q->v + w1 * w2;
        }
        int black = ((x % res) == 0) && ((y % res) == 0);
        if (fabsf(angle) < 0.95f) { if (res == 1) black = !black; else black = 1; }
        *o = black ? 0xff000000 : 0xffffffff;
      } else if (vw->flags & J3DG_WIREFRAME) { /* canvas.cpp:449-496 */
        if (!last && ((IS_HIT(right) && right->object_id != p->object_id) || (IS_HIT(up) && up->object_id != p->object_id))) {
          const float scale = (p->u * p->u + p->v * p->v) * 0.5f;
          unsigned char c = (unsigned char)(255 * scale);
          *o = 0xff000000 | ((uint32_t)c << 16) | ((uint32_t)c << 8) | c;
        } else *o = plain_color(&s, p);
      } else if (vw->flags & J3DG_EDGES) { /* canvas.cpp:590-649 */
        if (!last && IS_HIT(right) && NDIFF(p, right, threshold)) {
          float a = compute_convex_cos_angle(&s, (float)x, (float)y, p->u, p->v, p->depth, (float)x + 1.f, (float)y, right->u, right->v, right->depth);
          *o = get_angle_color(&s, a, p->u, p->v, p->mark);
        } else if (!last && IS_HIT(up) && NDIFF(p, up, threshold)) {
          float a = compute_convex_cos_angle(&s, (float)x, (float)y, p->u, p->v, p->depth, (float)x, (float)y - 1.f, up->u, up->v, up->depth);
          *o = get_angle_color(&s, a, p->u, p->v, p->mark);
        } else *o = plain_color(&s, p);
      } else { /* canvas.cpp:650-669 */
        *o = plain_color(&s, p);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * point clouds: render.h:133-154, 224-241, 273-288, 307-865 and canvas.cpp:952-1030
 * ---------------------------------------------------------------------------------- */
static void r_invert_orthonormal(float* out, const float* in)
{ /* render.h:133-154 (out[15]=1, bottom row zero) */
  out[0] = in[0]; out[1] = in[4]; out[2] = in[8]; out[4] = in[1]; out[5] = in[5]; out[6] = in[9];
  out[8] = in[2]; out[9] = in[6]; out[10] = in[10]; out[3] = 0; out[7] = 0; out[11] = 0; out[15] = 1;
  out[12] = -(in[0] * in[12] + in[1] * in[13] + in[2] * in[14]);
  out[13] = -(in[4] * in[12] + in[5] * in[13] + in[6] * in[14]);
  out[14] = -(in[8] * in[12] + in[9] * in[13] + in[10] * in[14]);
}
static void r_matrix_multiply(float* out, const float* left, const float* right)
{ /* render.h:224-233 */
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      out[i + (j << 2)] = left[i] * right[(j << 2)] + left[i + 4] * right[(j << 2) + 1] + left[i + 8] * right[(j << 2) + 2] + left[i + 12] * right[(j << 2) + 3];
}
static void r_mat_vec(float* out, const float* m, const float* v)
{ /* render.h:235-241 */
  out[0] = m[0] * v[0] + m[4] * v[1] + m[8] * v[2] + m[12] * v[3];
  out[1] = m[1] * v[0] + m[5] * v[1] + m[9] * v[2] + m[13] * v[3];
  out[2] = m[2] * v[0] + m[6] * v[1] + m[10] * v[2] + m[14] * v[3];
  out[3] = m[3] * v[0] + m[7] * v[1] + m[11] * v[2] + m[15] * v[3];
}
static int32_t cvt_rne(float x) { return _mm_cvtss_si32(_mm_set_ss(x)); }   /* _mm_cvtps_epi32 lane */
static int32_t cvt_trunc(float x) { return _mm_cvttss_si32(_mm_set_ss(x)); } /* (int) cast on x86-64 */

void orc_splat(const float* const* pos, const float* const* nrm, const uint32_t* const* clr, const uint32_t* counts,
               const float* const* cloud_cs, const uint32_t* db_ids, uint32_t nclouds, const j3dg_view* vw,
               const j3dg_pixel* px_in, j3dg_pixel* px_inout, uint32_t stride, uint32_t* rgba, uint32_t rgba_stride)
{
  if (!nclouds) return;
  const int w = (int)vw->width, h = (int)vw->height;
  float* zbuf = (float*)malloc(sizeof(float) * (size_t)w * h);
  /* canvas.cpp:962-972 */
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const j3dg_pixel* p = px_in + (size_t)y * stride + x;
      zbuf[(size_t)y * w + x] = (p->db_id != 0) ? 1.f / p->depth : 0.f;
    }
  /* the reference indexes `im` and `_canvas` as x + y*w (canvas.cpp:975, 1001); honour strides */
#define RGBA_AT(idx) rgba[(size_t)((idx) / w) * rgba_stride + ((idx) % w)]
#define PX_AT(idx) px_inout[(size_t)((idx) / w) * stride + ((idx) % w)]
  const int one_bit = (vw->flags & J3DG_ONE_BIT) != 0, shading = (vw->flags & J3DG_SHADING) != 0;
  for (uint32_t c = 0; c < nclouds; ++c) {
    const uint32_t n = counts[c];
    const float* P = pos[c];
    const float* N = shading ? nrm[c] : 0;           /* canvas.cpp:994 */
    const uint32_t* COL = one_bit ? 0 : clr[c];      /* canvas.cpp:995 */
    float object_system[16];
    if (cloud_cs && cloud_cs[c]) memcpy(object_system, cloud_cs[c], sizeof(object_system));
    else { memset(object_system, 0, sizeof(object_system)); object_system[0] = object_system[5] = object_system[10] = object_system[15] = 1.f; }
    float temp[16], M[16];
    r_matrix_multiply(temp, vw->cs_inv, object_system);     /* render.h:279 */
    r_matrix_multiply(M, vw->projection, temp);             /* render.h:280 */
    /* present(): light, render.h:521-527 */
    float inv[16], light[4] = {0.f, 0.f, 1.f, 0.f}, tmp[4];
    r_invert_orthonormal(inv, vw->cs_inv);
    r_mat_vec(tmp, inv, light);
    r_invert_orthonormal(inv, object_system);
    r_mat_vec(light, inv, tmp);
    const float light_color[3] = {255 / 255.f, 255 / 255.f, 255 / 255.f};
    const uint32_t sz = n - (n & 3);
    const float halfw = w * 0.5f, halfh = h * 0.5f;
    for (uint32_t i = 0; i < sz; i += 4) {
      int32_t X[4], Y[4], idx[4], masked[4];
      float depth[4]; uint32_t colors[4];
      int all_masked = 1;
      for (int l = 0; l < 4; ++l) { /* _draw SIMD body, render.h:419-468 */
        const float x = P[3 * (size_t)(i + l)], y = P[3 * (size_t)(i + l) + 1], z = P[3 * (size_t)(i + l) + 2];
        float VX = ((M[0] * x + M[4] * y) + M[8] * z) + M[12] * 1.f;
        float VY = ((M[1] * x + M[5] * y) + M[9] * z) + M[13] * 1.f;
        float VW = ((M[3] * x + M[7] * y) + M[11] * z) + M[15] * 1.f;
        VX = VX / VW; VY = VY / VW;
        VX = (VX + 1.f) * halfw; VY = (VY + 1.f) * halfh;
        X[l] = cvt_rne(VX); Y[l] = cvt_rne(VY);                      /* render.h:726-727 */
        masked[l] = (0 > X[l]) || (0 > Y[l]) || (X[l] > w - 1) || (Y[l] > h - 1);
        if (!masked[l]) all_masked = 0;
        depth[l] = 1.f / VW;                                        /* render.h:774 */
        int32_t id = (int32_t)((uint32_t)X[l] + (uint32_t)w * (uint32_t)Y[l]); /* wraps like epi32 */
        if (id < 0) id = 0; if (id > w * h - 1) id = w * h - 1;     /* render.h:780-781 */
        idx[l] = id;
      }
      if (all_masked) continue;                                     /* render.h:734-735 */
      for (int l = 0; l < 4; ++l) {
        colors[l] = COL ? COL[i + l] : 0xffffffffu;
        if (N) { /* render.h:744-772 */
          const float nx = N[3 * (size_t)(i + l)], ny = N[3 * (size_t)(i + l) + 1], nz = N[3 * (size_t)(i + l) + 2];
          float diffuse = 0.5f + ((nx * light[0] + ny * light[1]) + nz * light[2]);
          if (diffuse < 0.f) diffuse = 0.f;
          if (1.f < diffuse) diffuse = 1.f;
          const float ir = diffuse * light_color[0], ig = diffuse * light_color[1], ib = diffuse * light_color[2];
          const float red = (float)(int32_t)(colors[l] & 0xff), green = (float)(int32_t)((colors[l] & 0xff00) >> 8), blue = (float)(int32_t)((colors[l] & 0xff0000) >> 16);
          int32_t r2 = cvt_rne(ir * red), g2 = cvt_rne(ig * green), b2 = cvt_rne(ib * blue);
          if (r2 > 255) r2 = 255; if (g2 > 255) g2 = 255; if (b2 > 255) b2 = 255;
          colors[l] = 0xff000000u + ((uint32_t)b2 << 16) + ((uint32_t)g2 << 8) + (uint32_t)r2;
        }
      }
      /* render.h:783-807: all four lanes read first, then written in lane order (a later lane
       * of the same packet that maps to the same pixel overwrites an earlier one, and a lane
       * that fails its test writes the pre-packet value back) */
      float prev_d[4]; uint32_t prev_c[4]; int pass[4];
      for (int l = 0; l < 4; ++l) { prev_d[l] = zbuf[idx[l]]; prev_c[l] = RGBA_AT(idx[l]); }
      for (int l = 0; l < 4; ++l) pass[l] = !masked[l] && (prev_d[l] < depth[l]);
      for (int l = 0; l < 4; ++l) RGBA_AT(idx[l]) = pass[l] ? colors[l] : prev_c[l];
      for (int l = 0; l < 4; ++l) zbuf[idx[l]] = pass[l] ? depth[l] : prev_d[l];
      for (int l = 0; l < 4; ++l)
        if (pass[l]) { /* canvas.cpp:997-1027 */
          j3dg_pixel* q = &PX_AT(idx[l]);
          q->object_id = i + l; q->depth = 1.f / zbuf[idx[l]]; q->db_id = db_ids[c];
        }
    }
    for (uint32_t i = sz; i < n; ++i) { /* scalar tail: render.h:473-511 and 814-864 */
      const float x = P[3 * (size_t)i], y = P[3 * (size_t)i + 1], z = P[3 * (size_t)i + 2];
      float VX = M[0] * x + M[4] * y + M[8] * z + M[12];
      float VY = M[1] * x + M[5] * y + M[9] * z + M[13];
      float VZ = M[2] * x + M[6] * y + M[10] * z + M[14];
      float VW = M[3] * x + M[7] * y + M[11] * z + M[15];
      uint32_t clip = 0;
      VX /= VW; VY /= VW; VZ /= VW;
      if (VX < -1.f) clip |= 1; else if (VX > 1.f) clip |= 2;
      if (VY < -1.f) clip |= 4; else if (VY > 1.f) clip |= 8;
      if (VZ < -1.f) clip |= 16; else if (VZ > 1.f) clip |= 32;
      VX = (VX + 1.f) * w * 0.5f; VY = (VY + 1.f) * h * 0.5f;
      if (clip) continue;
      int Xi = cvt_trunc(VX), Yi = cvt_trunc(VY);
      if (Xi < 0 || Yi < 0 || Xi > w - 1 || Yi > h - 1) continue;
      uint32_t color = COL ? COL[i] : 0xffffffffu;
      if (N) {
        float nx = N[3 * (size_t)i], ny = N[3 * (size_t)i + 1], nz = N[3 * (size_t)i + 2];
        float diffuse = 0.5f + nx * light[0] + ny * light[1] + nz * light[2];
        diffuse = diffuse < 0.f ? 0.f : (1.f < diffuse) ? 1.f : diffuse;
        float ir = diffuse * light_color[0], ig = diffuse * light_color[1], ib = diffuse * light_color[2];
        float fr = (float)(color & 0xff) * ir, fg = (float)((color >> 8) & 0xff) * ig, fb = (float)((color >> 16) & 0xff) * ib;
        int red = (int)(255.f < fr ? 255.f : fr), green = (int)(255.f < fg ? 255.f : fg), blue = (int)(255.f < fb ? 255.f : fb);
        color = 0xff000000 | ((uint32_t)blue << 16) | ((uint32_t)green << 8) | (uint32_t)red;
      }
      float zz = 1.f / VW;
      uint32_t id = (uint32_t)(Xi + w * Yi);
      if (zz > zbuf[id]) {
        zbuf[id] = zz; RGBA_AT(id) = color;
        j3dg_pixel* q = &PX_AT(id);
        q->object_id = i; q->depth = 1.f / zbuf[id]; q->db_id = db_ids[c];
      }
    }
  }
#undef RGBA_AT
#undef PX_AT
  free(zbuf);
}

/* ---- picking ---------------------------------------------------------------
 * canvas::get_pixel (canvas.cpp:148-153), view::get_id (view.cpp:483-492), view::get_world_position
 * (view.cpp:439-469), get_closest_vertex (pixel.cpp:6-33), pivot pick of canvas::do_mouse
 * (canvas.cpp:157-179).  Plain C without contraction: every product / sum rounds separately like
 * the reference's SSE build.  Cloud pixels: world position = cloud cs * point, closest vertex =
 * the point index (pixel.cpp:28-32). */
void orc_pick(const j3dg_pixel* px, uint32_t stride, const j3dg_view* v, orc_mesh* const* meshes, uint32_t nm,
              const float* const* cloud_pos, const uint32_t* cloud_counts, const float* const* cloud_cs,
              const uint32_t* cloud_db_ids, uint32_t nclouds, const int32_t* xy, uint32_t n, j3dg_pick_result* out)
{
  union { uint32_t u; float f; } qn; qn.u = 0x7fc00000u;
  for (uint32_t i = 0; i < n; ++i) {
    j3dg_pick_result r;
    memset(&r, 0, sizeof(r));
    for (int k = 0; k < 3; ++k) { r.world_pos[k] = qn.f; r.pivot[k] = qn.f; }
    r.closest_vertex = 0xFFFFFFFFu;
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (x >= 0 && y >= 0 && x < (int)v->width && y < (int)v->height) {
      const j3dg_pixel p = px[(size_t)y * stride + x];
      r.pixel = p;
      r.db_id = p.db_id;
      if (p.db_id) {
        /* canvas.cpp:165-177 */
        const float w = (float)v->width, h = (float)v->height;
        float sp[4], dir[4], d2[4], o4[4] = {0.f, 0.f, 0.f, 1.f}, org[4];
        sp[0] = 2.f * (((float)x + 0.5f) / w) - 1.f;
        sp[1] = 2.f * (((float)y + 0.5f) / h) - 1.f;
        sp[2] = v->near_plane; sp[3] = 1.f;
        mat_vec(v->projection_inv, sp, dir);
        dir[3] = 0.f;
        mat_vec(v->cs, dir, d2);
        mat_vec(v->cs, o4, org);
        for (int k = 0; k < 3; ++k) { float t = p.depth * d2[k]; r.pivot[k] = org[k] + t; }
        int done = 0;
        for (uint32_t k = 0; k < nm && !done; ++k) {
          const orc_mesh* m = meshes[k];
          if (m->db_id != p.db_id || p.object_id >= m->nt) continue;
          const uint32_t v0 = m->tris[3 * (size_t)p.object_id], v1 = m->tris[3 * (size_t)p.object_id + 1], v2 = m->tris[3 * (size_t)p.object_id + 2];
          const float* A = m->verts + 3 * (size_t)v0; const float* B = m->verts + 3 * (size_t)v1; const float* Cc = m->verts + 3 * (size_t)v2;
          const float kk = 1.f - p.barycentric_u - p.barycentric_v;
          float pos[4];
          for (int j = 0; j < 3; ++j) { float a = A[j] * kk, b = p.barycentric_u * B[j], c = p.barycentric_v * Cc[j]; float s = a + b; pos[j] = s + c; }
          { float a = 1.f * kk, b = p.barycentric_u * 1.f, c = p.barycentric_v * 1.f; float s = a + b; pos[3] = s + c; }
          float wp[4];
          mat_vec(m->cs, pos, wp);
          for (int j = 0; j < 3; ++j) r.world_pos[j] = wp[j];
          float dA = 0.f, dB = 0.f, dC = 0.f;
          { float e0 = pos[0] - A[0], e1 = pos[1] - A[1], e2 = pos[2] - A[2]; float s = e0 * e0; float t = e1 * e1; s = s + t; t = e2 * e2; dA = s + t; }
          { float e0 = pos[0] - B[0], e1 = pos[1] - B[1], e2 = pos[2] - B[2]; float s = e0 * e0; float t = e1 * e1; s = s + t; t = e2 * e2; dB = s + t; }
          { float e0 = pos[0] - Cc[0], e1 = pos[1] - Cc[1], e2 = pos[2] - Cc[2]; float s = e0 * e0; float t = e1 * e1; s = s + t; t = e2 * e2; dC = s + t; }
          r.closest_vertex = (dA < dB) ? ((dA < dC) ? v0 : v2) : ((dB < dC) ? v1 : v2);
          done = 1;
        }
        for (uint32_t k = 0; k < nclouds && !done; ++k) {
          if (cloud_db_ids[k] != p.db_id || p.object_id >= cloud_counts[k]) continue;
          const float* q = cloud_pos[k] + 3 * (size_t)p.object_id;
          float pos[4] = {q[0], q[1], q[2], 1.f}, wp[4];
          mat_vec(cloud_cs[k], pos, wp);
          for (int j = 0; j < 3; ++j) r.world_pos[j] = wp[j];
          r.closest_vertex = p.object_id;
          done = 1;
        }
      }
    }
    out[i] = r;
  }
}

/* ---- all hits + voxel export ------------------------------------------------
 * qbvh::find_all_triangles, qbvh.h:1854-2000: every triangle the Woop test accepts inside the
 * ORIGINAL interval (t_near, t_far); nothing shrinks.  Traversal order is this file's own. */
typedef void (*hit_sink)(void* user, uint32_t tri, float t, float u, float v);

static void mesh_find_all(const orc_mesh* m, const float org[3], const float dir[3], float t_near, float t_far, hit_sink sink, void* user)
{
  if (!m->nt) return;
  woop_pre pre = woop_precompute(dir);
  double inv[3];
  for (int j = 0; j < 3; ++j) inv[j] = 1.0 / (double)dir[j];
  uint32_t stack[128]; int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const bnode* n = &m->nodes[stack[--sp]];
    double lo = (double)t_near, hi = (double)t_far;
    int miss = 0;
    for (int j = 0; j < 3 && !miss; ++j) {
      if (dir[j] == 0.f) { if (org[j] < n->mn[j] || org[j] > n->mx[j]) miss = 1; continue; }
      double a = ((double)n->mn[j] - (double)org[j]) * inv[j], b = ((double)n->mx[j] - (double)org[j]) * inv[j];
      if (a > b) { double t = a; a = b; b = t; }
      double pad = 1e-6 * (fabs(a) + fabs(b)) + 1e-30;
      a -= pad; b += pad;
      if (a > lo) lo = a;
      if (b < hi) hi = b;
      if (lo > hi) miss = 1;
    }
    if (miss) continue;
    if (n->count) {
      for (uint32_t i = 0; i < n->count; ++i) {
        uint32_t t = m->order[n->first + i];
        const uint32_t* tr = m->tris + 3 * (size_t)t;
        float tt, uu, vv;
        if (woop_intersect(m->verts + 3 * (size_t)tr[0], m->verts + 3 * (size_t)tr[1], m->verts + 3 * (size_t)tr[2], &pre, org, t_near, t_far, &tt, &uu, &vv))
          sink(user, t, tt, uu, vv);
      }
    } else {
      stack[sp++] = n->left; stack[sp++] = n->right;
    }
  }
}

typedef struct { uint32_t total, capacity; float* hits; uint32_t* ids; } csr_sink_t;
static void csr_sink(void* user, uint32_t tri, float t, float u, float v)
{
  csr_sink_t* s = (csr_sink_t*)user;
  if (s->total < s->capacity && s->hits) {
    s->hits[4 * (size_t)s->total + 0] = u; s->hits[4 * (size_t)s->total + 1] = v;
    s->hits[4 * (size_t)s->total + 2] = t; s->hits[4 * (size_t)s->total + 3] = 0.f;
    s->ids[s->total] = tri;
  }
  s->total++;
}

uint32_t orc_find_all(const orc_mesh* m, const float* rays, uint32_t n, uint32_t* offsets, float* hits, uint32_t* ids, uint32_t capacity)
{
  csr_sink_t s; s.total = 0; s.capacity = capacity; s.hits = hits; s.ids = ids;
  for (uint32_t i = 0; i < n; ++i) {
    offsets[i] = s.total;
    mesh_find_all(m, rays + 8 * (size_t)i, rays + 8 * (size_t)i + 3, rays[8 * (size_t)i + 6], rays[8 * (size_t)i + 7], csr_sink, &s);
  }
  offsets[n] = s.total;
  return s.total;
}

/* vox.cpp:154-172 */
static uint8_t color_to_index(uint8_t r, uint8_t g, uint8_t b)
{
  if (r < 16) r = 0; else r -= 16;
  if (g < 16) g = 0; else g -= 16;
  if (b < 32) b = 0; else b -= 32;
  uint8_t ret = (uint8_t)(((r >> 5) << 5) | ((g >> 5) << 2) | (b >> 6));
  return ret == 0 ? 1 : ret;
}

/* vox.cpp:274-289 */
void orc_voxel_dims(const orc_mesh* m, uint32_t max_dim, uint32_t dims[3])
{
  int largest = 0;
  if ((m->bb_max[1] - m->bb_min[1]) > (m->bb_max[largest] - m->bb_min[largest])) largest = 1;
  if ((m->bb_max[2] - m->bb_min[2]) > (m->bb_max[largest] - m->bb_min[largest])) largest = 2;
  for (int j = 0; j < 3; ++j) {
    float a = (float)max_dim * (m->bb_max[j] - m->bb_min[j]);
    float b = a / (m->bb_max[largest] - m->bb_min[largest]);
    dims[j] = (uint32_t)b;
    if (dims[j] == 0) dims[j] = 1;
  }
}

typedef struct { const orc_mesh* m; uint32_t dim[3]; uint8_t* vmax; uint8_t* vmin; } vox_sink_t;
static void vox_sink(void* user, uint32_t tri, float t, float u, float v)
{ /* vox.cpp:336-375 */
  (void)t;
  vox_sink_t* s = (vox_sink_t*)user;
  const orc_mesh* m = s->m;
  const uint32_t* tr = m->tris + 3 * (size_t)tri;
  const float* V0 = m->verts + 3 * (size_t)tr[0]; const float* V1 = m->verts + 3 * (size_t)tr[1]; const float* V2 = m->verts + 3 * (size_t)tr[2];
  const float k = 1.f - u - v;
  float pos[3];
  for (int j = 0; j < 3; ++j) { float a = V0[j] * k, b = u * V1[j], c = v * V2[j]; float q = a + b; pos[j] = q + c; }
  float clr[3] = {1.f, 1.f, 1.f};
  if (m->uv && m->tex && m->tw > 0 && m->th > 0) {
    const float* uvc = m->uv + 6 * (size_t)tri;
    float cx, cy;
    { float a = k * uvc[0], b = u * uvc[2], c = v * uvc[4]; float q = a + b; cx = q + c; }
    { float a = k * uvc[1], b = u * uvc[3], c = v * uvc[5]; float q = a + b; cy = q + c; }
    cx = cx < 1.f ? cx : 1.f; cx = cx > 0.f ? cx : 0.f;   /* std::max(std::min(c, 1.f), 0.f) */
    cy = cy < 1.f ? cy : 1.f; cy = cy > 0.f ? cy : 0.f;
    int x = (int)(cx * (float)(m->tw - 1)), y = (int)(cy * (float)(m->th - 1));
    uint32_t color = m->tex[(size_t)y * m->tw + x];
    clr[0] = (float)(color & 255) / 255.f; clr[1] = (float)((color >> 8) & 255) / 255.f; clr[2] = (float)((color >> 16) & 255) / 255.f;
  } else if (m->vcolors) {
    const float* c0 = m->vcolors + 3 * (size_t)tr[0]; const float* c1 = m->vcolors + 3 * (size_t)tr[1]; const float* c2 = m->vcolors + 3 * (size_t)tr[2];
    for (int j = 0; j < 3; ++j) { float a = c0[j] * k, b = u * c1[j], c = v * c2[j]; float q = a + b; clr[j] = q + c; }
  }
  uint32_t X[3];
  for (int j = 0; j < 3; ++j) {
    float x = (pos[j] - m->bb_min[j]) / (m->bb_max[j] - m->bb_min[j]);
    float xs = x * (float)s->dim[j];
    X[j] = xs > 0.f ? (uint32_t)xs : 0u;   /* cvttss2si of a small negative value truncates to 0 */
    if (X[j] == s->dim[j]) X[j] = s->dim[j] - 1;
    if (X[j] >= s->dim[j]) return;
  }
  float r255 = clr[0] * 255.f, g255 = clr[1] * 255.f, b255 = clr[2] * 255.f;
  uint8_t ci = color_to_index((uint8_t)r255, (uint8_t)g255, (uint8_t)b255);
  size_t idx = (size_t)X[0] + ((size_t)X[1] + (size_t)X[2] * s->dim[1]) * s->dim[0];
  if (ci > s->vmax[idx]) s->vmax[idx] = ci;
  if (s->vmin && (s->vmin[idx] == 0 || ci < s->vmin[idx])) s->vmin[idx] = ci;
}

/* The grid-filling loop of _write_vox (vox.cpp:300-379).  vmax: largest palette index written into each voxel
 * (what the CUDA path stores); vmin (nullable): smallest — the reference's own result (last writer among its
 * threads) lies between the two, and equals both wherever only one colour fell into the voxel. */
void orc_voxelize(const orc_mesh* m, uint32_t max_dim, uint32_t dims[3], uint8_t* vmax, uint8_t* vmin)
{
  vox_sink_t s;
  s.m = m; s.vmax = vmax; s.vmin = vmin;
  orc_voxel_dims(m, max_dim, s.dim);
  for (int j = 0; j < 3; ++j) dims[j] = s.dim[j];
  if (!vmax) return;
  size_t nvox = (size_t)s.dim[0] * s.dim[1] * s.dim[2];
  memset(vmax, 0, nvox);
  if (vmin) memset(vmin, 0, nvox);
  for (int dd = 0; dd < 3; ++dd) {
    float dir[3] = {0.f, 0.f, 0.f};
    dir[dd] = dd == 2 ? 2.f : 1.f;
    const int d1i = (dd + 1) % 3, d2i = (dd + 2) % 3;
    for (uint32_t d1 = 0; d1 < s.dim[d1i]; ++d1)
      for (uint32_t d2 = 0; d2 < s.dim[d2i]; ++d2) {
        float org[3];
        org[dd] = m->bb_min[dd];
        { float a = ((float)d1 + 0.5f) / (float)s.dim[d1i]; float b = a * (m->bb_max[d1i] - m->bb_min[d1i]); org[d1i] = b + m->bb_min[d1i]; }
        { float a = ((float)d2 + 0.5f) / (float)s.dim[d2i]; float b = a * (m->bb_max[d2i] - m->bb_min[d2i]); org[d2i] = b + m->bb_min[d2i]; }
        mesh_find_all(m, org, dir, 0.f, FLT_MAX, vox_sink, &s);
      }
  }
}
