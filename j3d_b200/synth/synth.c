/*
 * synth.c — procedural benchmark / test inputs: noised geodesic
 * icospheres of frequency f (T = 20 f^2 triangles, V = 10 f^2 + 2 shared vertices) and
 * noised spherical point clouds.  Deterministic for a given (f, seed); plain C, OpenMP
 * where available.  Input generation only — no renderer logic lives here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* ---- hash / value noise -------------------------------------------------------- */
static inline uint32_t hash_u32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

static inline float lattice(int32_t ix, int32_t iy, int32_t iz, uint32_t seed)
{
  uint32_t h = hash_u32((uint32_t)ix * 0x8da6b343U ^ hash_u32((uint32_t)iy * 0xd8163841U ^ hash_u32((uint32_t)iz * 0xcb1ab31fU ^ seed)));
  return (float)(h >> 8) * (1.0f / 8388608.0f) - 1.0f; /* [-1,1) */
}

static inline float smooth(float t) { return t * t * (3.0f - 2.0f * t); }

static float value_noise(float x, float y, float z, uint32_t seed)
{
  float fx = floorf(x), fy = floorf(y), fz = floorf(z);
  int32_t ix = (int32_t)fx, iy = (int32_t)fy, iz = (int32_t)fz;
  float tx = smooth(x - fx), ty = smooth(y - fy), tz = smooth(z - fz);
  float c000 = lattice(ix, iy, iz, seed), c100 = lattice(ix + 1, iy, iz, seed);
  float c010 = lattice(ix, iy + 1, iz, seed), c110 = lattice(ix + 1, iy + 1, iz, seed);
  float c001 = lattice(ix, iy, iz + 1, seed), c101 = lattice(ix + 1, iy, iz + 1, seed);
  float c011 = lattice(ix, iy + 1, iz + 1, seed), c111 = lattice(ix + 1, iy + 1, iz + 1, seed);
  float x00 = c000 + (c100 - c000) * tx, x10 = c010 + (c110 - c010) * tx;
  float x01 = c001 + (c101 - c001) * tx, x11 = c011 + (c111 - c011) * tx;
  float y0 = x00 + (x10 - x00) * ty, y1 = x01 + (x11 - x01) * ty;
  return y0 + (y1 - y0) * tz;
}

/* 3 octaves, result in about [-1,1] */
static float fbm3(float x, float y, float z, uint32_t seed)
{
  float s = 0.f, a = 1.f, f = 3.f;
  for (int o = 0; o < 3; ++o) {
    s += a * value_noise(x * f + 17.f * (float)o, y * f - 5.f * (float)o, z * f + 3.f * (float)o, seed + (uint32_t)o * 101u);
    a *= 0.5f; f *= 2.f;
  }
  return s * (1.0f / 1.75f);
}

/* ---- icosahedron --------------------------------------------------------------- */
static const double ICO_T = 1.6180339887498948482; /* golden ratio */
static void ico_corners(double c[12][3])
{
  const double t = ICO_T;
  const double raw[12][3] = {
    {-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0},
    {0, -1, t}, {0, 1, t}, {0, -1, -t}, {0, 1, -t},
    {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
  for (int i = 0; i < 12; ++i) {
    double l = sqrt(raw[i][0] * raw[i][0] + raw[i][1] * raw[i][1] + raw[i][2] * raw[i][2]);
    for (int j = 0; j < 3; ++j) c[i][j] = raw[i][j] / l;
  }
}
/* 20 faces, counter-clockwise seen from outside */
static const int ICO_F[20][3] = {
  {0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11},
  {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
  {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9},
  {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};

EXPORT void synth_icosphere_counts(uint32_t f, uint64_t* nv, uint64_t* nt)
{
  *nv = 10ull * f * f + 2;
  *nt = 20ull * f * f;
}

static void displace(const double p[3], float amp, uint32_t seed, float* out)
{
  double l = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  float x = (float)(p[0] / l), y = (float)(p[1] / l), z = (float)(p[2] / l);
  float r = 1.0f + amp * fbm3(x, y, z, seed);
  out[0] = x * r; out[1] = y * r; out[2] = z * r;
}

typedef struct { int lo, hi; } edge_t;

/* Vertex numbering: 12 corners | 30 edges x (f-1) | 20 faces x (f-1)(f-2)/2 interior. */
EXPORT void synth_icosphere(uint32_t f, float noise_amp, uint32_t seed, float* verts, uint32_t* tris)
{
  double C[12][3];
  ico_corners(C);
  edge_t edges[30];
  int edge_of[12][12];
  int ne = 0;
  for (int a = 0; a < 12; ++a) for (int b = 0; b < 12; ++b) edge_of[a][b] = -1;
  for (int t = 0; t < 20; ++t)
    for (int k = 0; k < 3; ++k) {
      int a = ICO_F[t][k], b = ICO_F[t][(k + 1) % 3];
      int lo = a < b ? a : b, hi = a < b ? b : a;
      if (edge_of[lo][hi] < 0) { edge_of[lo][hi] = edge_of[hi][lo] = ne; edges[ne].lo = lo; edges[ne].hi = hi; ++ne; }
    }
  const uint64_t F = f;
  const uint64_t per_edge = F - 1;
  const uint64_t per_face = (F >= 2) ? (F - 1) * (F - 2) / 2 : 0;
  const uint64_t edge_base = 12, face_base = 12 + 30 * per_edge;

  for (int i = 0; i < 12; ++i) displace(C[i], noise_amp, seed, verts + 3 * i);
  for (int e = 0; e < 30; ++e) {
    const double* A = C[edges[e].lo]; const double* B = C[edges[e].hi];
#pragma omp parallel for schedule(static)
    for (int64_t k = 1; k < (int64_t)F; ++k) {
      double p[3];
      for (int j = 0; j < 3; ++j) p[j] = ((double)(F - k) * A[j] + (double)k * B[j]) / (double)F;
      displace(p, noise_amp, seed, verts + 3 * (edge_base + (uint64_t)e * per_edge + (uint64_t)(k - 1)));
    }
  }
  /* interior point (i,j): weight (F-i-j) on A, i on B, j on C, with i>=1, j>=1, i+j<=F-1.
     row j holds i = 1..F-1-j; rows are stored j = 1..F-2. offset(j) = sum_{m=1}^{j-1} (F-1-m) */
  for (int t = 0; t < 20; ++t) {
    const double* A = C[ICO_F[t][0]]; const double* B = C[ICO_F[t][1]]; const double* Cc = C[ICO_F[t][2]];
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t j = 1; j <= (int64_t)F - 2; ++j) {
      uint64_t off = (uint64_t)(j - 1) * (F - 1) - (uint64_t)(j - 1) * (uint64_t)j / 2;
      for (int64_t i = 1; i <= (int64_t)F - 1 - j; ++i) {
        double p[3];
        double wa = (double)((int64_t)F - i - j), wb = (double)i, wc = (double)j;
        for (int q = 0; q < 3; ++q) p[q] = (wa * A[q] + wb * B[q] + wc * Cc[q]) / (double)F;
        displace(p, noise_amp, seed, verts + 3 * (face_base + (uint64_t)t * per_face + off + (uint64_t)(i - 1)));
      }
    }
  }

  /* triangles: face-major, row-major (scanner-like coherent order) */
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < 20; ++t) {
    const int a = ICO_F[t][0], b = ICO_F[t][1], c = ICO_F[t][2];
    uint32_t* out = tris + 3ull * (uint64_t)t * F * F;
    /* index of grid point (i,j) of this face */
#define EDGE_IDX(u, v, k) /* k steps from u towards v, 1..F-1 */ \
  (uint32_t)(edge_base + (uint64_t)edge_of[u][v] * per_edge + (uint64_t)(((u) < (v) ? (k) : ((int64_t)F - (k))) - 1))
    for (int64_t j = 0; j < (int64_t)F; ++j) {
      for (int64_t i = 0; i < (int64_t)F - j; ++i) {
        /* up triangle (i,j),(i+1,j),(i,j+1); down triangle (i+1,j),(i+1,j+1),(i,j+1) if i+j+1<F */
        int64_t gi[4] = {i, i + 1, i, i + 1}, gj[4] = {j, j, j + 1, j + 1};
        uint32_t id[4];
        int n = (i + j + 1 < (int64_t)F) ? 4 : 3;
        for (int q = 0; q < n; ++q) {
          int64_t I = gi[q], J = gj[q], K = (int64_t)F - I - J; /* weights: K on a, I on b, J on c */
          uint32_t v;
          if (I == 0 && J == 0) v = (uint32_t)a;
          else if (J == 0 && K == 0) v = (uint32_t)b;
          else if (I == 0 && K == 0) v = (uint32_t)c;
          else if (J == 0) v = EDGE_IDX(a, b, I);       /* on edge a-b, I steps from a */
          else if (I == 0) v = EDGE_IDX(a, c, J);       /* on edge a-c, J steps from a */
          else if (K == 0) v = EDGE_IDX(b, c, J);       /* on edge b-c, J steps from b */
          else {
            uint64_t off = (uint64_t)(J - 1) * (F - 1) - (uint64_t)(J - 1) * (uint64_t)J / 2;
            v = (uint32_t)(face_base + (uint64_t)t * per_face + off + (uint64_t)(I - 1));
          }
          id[q] = v;
        }
        out[0] = id[0]; out[1] = id[1]; out[2] = id[2]; out += 3;
        if (n == 4) { out[0] = id[1]; out[1] = id[3]; out[2] = id[2]; out += 3; }
      }
    }
#undef EDGE_IDX
  }
}

static inline uint64_t xorshift64s(uint64_t* s)
{
  uint64_t x = *s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  *s = x;
  return x * 0x2545F4914F6CDD1DULL;
}

/* Fisher-Yates shuffle of whole triangles (tests the builder on incoherent input). */
EXPORT void synth_shuffle_triangles(uint32_t* tris, uint64_t nt, uint64_t seed)
{
  uint64_t s = seed ? seed : 1;
  for (uint64_t i = nt - 1; i > 0; --i) {
    uint64_t j = xorshift64s(&s) % (i + 1);
    uint32_t tmp[3];
    memcpy(tmp, tris + 3 * i, 12); memcpy(tris + 3 * i, tris + 3 * j, 12); memcpy(tris + 3 * j, tmp, 12);
  }
}

/* Point cloud: random directions (uniform cube, rejected to the unit ball), radius
 * 1 + amp*noise, normal = direction, colour = hash(i) | 0xFF000000.  Point i depends only
 * on (seed, i) so any index range can be generated independently. */
EXPORT void synth_cloud_range(uint64_t first, uint64_t n, uint64_t seed, float noise_amp, float* pos, float* nrm, uint32_t* clr)
{
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)n; ++k) {
    uint64_t i = first + (uint64_t)k;
    uint64_t s = (seed + 0x9E3779B97F4A7C15ULL * (i + 1)) | 1ull;
    xorshift64s(&s);
    float x, y, z, l2;
    do {
      x = (float)(xorshift64s(&s) >> 40) * (2.0f / 16777216.0f) - 1.0f;
      y = (float)(xorshift64s(&s) >> 40) * (2.0f / 16777216.0f) - 1.0f;
      z = (float)(xorshift64s(&s) >> 40) * (2.0f / 16777216.0f) - 1.0f;
      l2 = x * x + y * y + z * z;
    } while (l2 > 1.0f || l2 < 1e-4f);
    float inv = 1.0f / sqrtf(l2);
    x *= inv; y *= inv; z *= inv;
    float r = 1.0f + noise_amp * fbm3(x, y, z, (uint32_t)seed);
    pos[3 * k + 0] = x * r; pos[3 * k + 1] = y * r; pos[3 * k + 2] = z * r;
    if (nrm) { nrm[3 * k + 0] = x; nrm[3 * k + 1] = y; nrm[3 * k + 2] = z; }
    if (clr) clr[k] = hash_u32((uint32_t)i * 2654435761u + 12345u) | 0xFF000000u;
  }
}

EXPORT void synth_cloud(uint64_t n, uint64_t seed, float noise_amp, float* pos, float* nrm, uint32_t* clr)
{
  synth_cloud_range(0, n, seed, noise_amp, pos, nrm, clr);
}

/* Smooth per-vertex colours in [0,1] (for the vertex-colour path). */
EXPORT void synth_vertex_colors(const float* verts, uint64_t nv, uint32_t seed, float* rgb)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)nv; ++i) {
    const float* p = verts + 3 * i;
    for (int c = 0; c < 3; ++c) {
      float v = 0.5f + 0.5f * fbm3(p[0] + 7.f * (float)c, p[1] - 3.f * (float)c, p[2] + 11.f * (float)c, seed + 977u * (uint32_t)c);
      rgb[3 * i + c] = v < 0.f ? 0.f : (v > 0.999f ? 0.999f : v);
    }
  }
}
