// knn.cu — point-cloud normal estimation on the device.
// Replaces estimate_normals (j3d/pc.cpp:256-334): for every point the k nearest points (jtk::point_tree::find_k_nearest,
// jtk/point_tree.h:436-475, the point itself included), a plane through them (jtk::fit_plane, jtk/fitting.h:197-215: the
// direction of the smallest eigenvalue of the 3 x 3 scatter matrix), and a consistent orientation propagated over the
// neighbour graph (pc.cpp:284-333).
//
// B200 shape: the k-d tree becomes a uniform grid — points are keyed by cell, sorted with the renderer's radix sort, and a
// dense (first, count) table per cell makes a neighbourhood a handful of contiguous ranges of a float4 array.  One thread
// per point (in cell order, so a warp walks the same cells) grows a cube of cells ring by ring until the k-th distance is
// closer than the cube's nearest face.  Squared distances are rounded exactly like the reference's
// ((dx*dx + dy*dy) + dz*dz, no FMA) so the neighbour SETS are the reference's whenever distances are distinct.  The plane
// fit accumulates centroid and scatter matrix in float in the reference's order (neighbours by ascending distance) and
// diagonalises the SAME float matrix with a double-precision Jacobi iteration — the reference runs a float SVD on it,
// so normals agree to the accuracy of that SVD, up to sign.  The orientation pass is a priority-queue graph walk: serial
// by nature, O(n k log n) on the host over the downloaded lists, with the queue discipline of jtk::hashed_heap
// (containers.h:21-92: smallest |n_i . n_j| first, sift rules restated below).
#include "common.cuh"
#include "sort.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr int KNN_MAX_K = 64;
constexpr int KNN_THREADS = 128;

struct Grid {
  float lo[3];
  float inv_cell;   // 1 / cell
  float cell;
  float eps;        // bound on the rounding error of a cell-face position
  int dim[3];
};

__device__ __forceinline__ uint32_t float_flip(float f) {  // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float float_unflip(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ pos, uint32_t n, uint32_t* __restrict__ out) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = pos[3 * (size_t)i + a];
      lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&out[a], float_flip(lo[a])); atomicMax(&out[3 + a], float_flip(hi[a])); }
  }
}

__device__ __forceinline__ int cell_coord(float v, float lo, float inv_cell, int dim) {
  const int c = (int)floorf((v - lo) * inv_cell);
  return min(max(c, 0), dim - 1);
}

__global__ void __launch_bounds__(256) cell_key_kernel(const float* __restrict__ pos, uint32_t n, Grid g, uint64_t* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = cell_coord(pos[3 * (size_t)i], g.lo[0], g.inv_cell, g.dim[0]);
  const int cy = cell_coord(pos[3 * (size_t)i + 1], g.lo[1], g.inv_cell, g.dim[1]);
  const int cz = cell_coord(pos[3 * (size_t)i + 2], g.lo[2], g.inv_cell, g.dim[2]);
  keys[i] = (uint64_t)(((uint32_t)cz * (uint32_t)g.dim[1] + (uint32_t)cy) * (uint32_t)g.dim[0] + (uint32_t)cx);
}

// After the sort: the points in cell order as float4 (w = original index), the first index of every occupied cell, and the
// number of occupied cells.
__global__ void __launch_bounds__(256) cell_table_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ order, const float* __restrict__ pos, uint32_t n,
                                                         float4* __restrict__ sorted, uint32_t* __restrict__ first, uint32_t* __restrict__ count, uint32_t* __restrict__ occupied) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = order[i];
  sorted[i] = make_float4(pos[3 * (size_t)o], pos[3 * (size_t)o + 1], pos[3 * (size_t)o + 2], __uint_as_float(o));
  const uint64_t k = keys[i];
  if (i == 0 || keys[i - 1] != k) { first[k] = i; if (occupied) atomicAdd(occupied, 1u); }
  if (count) atomicAdd(&count[k], 1u);
}

// find_k_nearest (point_tree.h:436-475) for every point.  knn: n x k original indices by ascending (distance, index).
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const float4* __restrict__ sorted, uint32_t n, Grid g, const uint32_t* __restrict__ first,
                                                          const uint32_t* __restrict__ count, int k, uint32_t* __restrict__ knn) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float bd[KNN_MAX_K];
  uint32_t bi[KNN_MAX_K];
  const float4 q = sorted[i];
  const int cx = cell_coord(q.x, g.lo[0], g.inv_cell, g.dim[0]);
  const int cy = cell_coord(q.y, g.lo[1], g.inv_cell, g.dim[1]);
  const int cz = cell_coord(q.z, g.lo[2], g.inv_cell, g.dim[2]);
  int have = 0, worst = 0;
  float worst_d = -1.f;
  uint32_t worst_i = 0;
  const int rmax = max(g.dim[0], max(g.dim[1], g.dim[2]));
  for (int r = 0; r < rmax; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dim[2] - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dim[1] - 1);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dim[0] - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const bool shell_row = (z == cz - r || z == cz + r || y == cy - r || y == cy + r);
        // a row on the shell is walked whole (its cells are contiguous in the table: one range); an inner row only
        // contributes its two end cells
        for (int part = 0; part < 2; ++part) {
          int xa, xb;
          if (shell_row) { if (part) break; xa = x0; xb = x1; }
          else {
            if (r == 0) break;
            const int x = part ? cx + r : cx - r;
            if (x < 0 || x >= g.dim[0]) continue;
            xa = xb = x;
          }
          const size_t row = ((size_t)z * g.dim[1] + y) * g.dim[0];
          for (int x = xa; x <= xb; ++x) {
            const uint32_t c = count[row + x];
            if (!c) continue;
            const uint32_t f = first[row + x];
            for (uint32_t j = f; j < f + c; ++j) {
              const float4 p = sorted[j];
              const float dx = fsub(q.x, p.x), dy = fsub(q.y, p.y), dz = fsub(q.z, p.z);
              const float d = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));  // point_tree.h:52-57
              const uint32_t id = __float_as_uint(p.w);
              if (have < k) {
                bd[have] = d; bi[have] = id;
                if (d > worst_d || (d == worst_d && id > worst_i)) { worst_d = d; worst_i = id; worst = have; }
                ++have;
              } else if (d < worst_d || (d == worst_d && id < worst_i)) {
                bd[worst] = d; bi[worst] = id;
                worst_d = -1.f;
                for (int t = 0; t < k; ++t)
                  if (bd[t] > worst_d || (bd[t] == worst_d && bi[t] > worst_i)) { worst_d = bd[t]; worst_i = bi[t]; worst = t; }
              }
            }
          }
        }
      }
    if (have >= k) {
      // everything outside the cube of cells [c - r, c + r] is at least this far away (faces on the grid border do not count)
      float dmin = FLT_MAX;
      if (cx - r > 0) dmin = fminf(dmin, q.x - (g.lo[0] + (float)(cx - r) * g.cell));
      if (cx + r < g.dim[0] - 1) dmin = fminf(dmin, (g.lo[0] + (float)(cx + r + 1) * g.cell) - q.x);
      if (cy - r > 0) dmin = fminf(dmin, q.y - (g.lo[1] + (float)(cy - r) * g.cell));
      if (cy + r < g.dim[1] - 1) dmin = fminf(dmin, (g.lo[1] + (float)(cy + r + 1) * g.cell) - q.y);
      if (cz - r > 0) dmin = fminf(dmin, q.z - (g.lo[2] + (float)(cz - r) * g.cell));
      if (cz + r < g.dim[2] - 1) dmin = fminf(dmin, (g.lo[2] + (float)(cz + r + 1) * g.cell) - q.z);
      if (dmin == FLT_MAX) break;
      dmin = fmaxf(dmin - g.eps, 0.f);  // slack for the rounding of the face positions and of the cell assignment
      if (worst_d < dmin * dmin) break;
    }
  }
  // ascending (distance, index): selection sort, k is small
  uint32_t* out = knn + (size_t)__float_as_uint(q.w) * k;
  for (int a = 0; a < have; ++a) {
    int m = a;
    for (int b = a + 1; b < have; ++b)
      if (bd[b] < bd[m] || (bd[b] == bd[m] && bi[b] < bi[m])) m = b;
    const float td = bd[m]; const uint32_t ti = bi[m];
    bd[m] = bd[a]; bi[m] = bi[a];
    bd[a] = td; bi[a] = ti;
    out[a] = ti;
  }
  for (int a = have; a < k; ++a) out[a] = 0xFFFFFFFFu;
}

// fit_plane (fitting.h:197-215) on the neighbours of every point: centroid and scatter matrix in float in list order,
// then the eigenvector of the eigenvalue of smallest magnitude (cyclic Jacobi in double on the float matrix).
__global__ void __launch_bounds__(KNN_THREADS) fit_plane_kernel(const float* __restrict__ pos, const uint32_t* __restrict__ knn, uint32_t n, int k, float* __restrict__ normals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* nb = knn + (size_t)i * k;
  int m = 0;
  float o[3] = {0.f, 0.f, 0.f};
  for (; m < k; ++m) {  // centroid (fitting.h:179-195): the first point, then += the others, then /= count
    const uint32_t j = nb[m];
    if (j == 0xFFFFFFFFu) break;
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = m ? fadd(o[a], pos[3 * (size_t)j + a]) : pos[3 * (size_t)j + a];
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = fdiv(o[a], (float)m);
  float s[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int t = 0; t < m; ++t) {
    const uint32_t j = nb[t];
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = fsub(pos[3 * (size_t)j + a], o[a]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) s[a][b] = fadd(s[a][b], fmul(d[a], d[b]));
  }
  double A[3][3], V[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) { A[a][b] = (double)s[a][b]; V[a][b] = a == b ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300 || off <= 1e-17 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]))) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
        for (int r = 0; r < 3; ++r) {  // A <- A J
          const double arp = A[r][p], arq = A[r][q];
          A[r][p] = c * arp - sn * arq; A[r][q] = sn * arp + c * arq;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {  // A <- J^T A, V <- V J
          const double apr = A[p][r], aqr = A[q][r];
          A[p][r] = c * apr - sn * aqr; A[q][r] = sn * apr + c * aqr;
          const double vrp = V[r][p], vrq = V[r][q];
          V[r][p] = c * vrp - sn * vrq; V[r][q] = sn * vrp + c * vrq;
        }
      }
  }
  int e = 0;  // fitting.h:208-212
  if (fabs(A[1][1]) < fabs(A[e][e])) e = 1;
  if (fabs(A[2][2]) < fabs(A[e][e])) e = 2;
  const double len = sqrt(V[0][e] * V[0][e] + V[1][e] * V[1][e] + V[2][e] * V[2][e]);
  const double inv = len > 0.0 ? 1.0 / len : 0.0;
  normals[3 * (size_t)i] = (float)(V[0][e] * inv);
  normals[3 * (size_t)i + 1] = (float)(V[1][e] * inv);
  normals[3 * (size_t)i + 2] = (float)(V[2][e] * inv);
}

// ---- orientation propagation on the host (pc.cpp:284-333) -----------------------------------------
// A binary min-heap of (score, edge) with the sift rules of jtk::hashed_heap (containers.h:21-92): an element rises while
// it is strictly smaller than its parent; on the way down the smaller child is taken (the right one on a tie) while it is
// strictly smaller than the moving element.  An edge (v0 -> v1) is pushed at most once (v0 is treated once).
struct EdgeHeap {
  struct Item { float score; uint32_t v0, v1; };
  std::vector<Item> h;
  bool empty() const { return h.empty(); }
  void push(float score, uint32_t v0, uint32_t v1) {
    h.push_back(Item{score, v0, v1});
    size_t i = h.size() - 1;
    const Item val = h[i];
    while (i > 0) {
      const size_t parent = (i - 1) / 2;
      if (!(val.score < h[parent].score)) break;
      h[i] = h[parent];
      i = parent;
    }
    h[i] = val;
  }
  Item pop() {
    const Item top = h.front();
    h.front() = h.back();
    h.pop_back();
    const size_t len = h.size();
    if (len) {
      size_t i = 0;
      const Item val = h[0];
      for (;;) {
        const size_t left = 2 * i + 1, right = 2 * i + 2;
        size_t pick;
        if (right < len) pick = h[left].score < h[right].score ? left : right;
        else if (left < len) pick = left;
        else break;
        if (!(h[pick].score < val.score)) break;
        h[i] = h[pick];
        i = pick;
      }
      h[i] = val;
    }
    return top;
  }
};

inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }  // host code: plain SSE, no contraction (jtk/vec.h:483-486)

void orient_normals(std::vector<float>& nrm, const std::vector<uint32_t>& knn, uint32_t n, int k) {
  std::vector<uint8_t> treated(n, 0);
  EdgeHeap heap;
  auto push_neighbours = [&](uint32_t v) {
    const uint32_t* nb = &knn[(size_t)v * k];
    for (int j = 0; j < k; ++j) {
      const uint32_t u = nb[j];
      if (u == 0xFFFFFFFFu) break;
      if (u != v && !treated[u]) heap.push(std::fabs(dot3(&nrm[3 * (size_t)v], &nrm[3 * (size_t)u])), v, u);
    }
  };
  uint32_t v = 0;
  for (;;) {
    while (v < n && treated[v]) ++v;
    if (v == n) break;
    treated[v] = 1;
    push_neighbours(v);
    while (!heap.empty()) {
      const EdgeHeap::Item e = heap.pop();
      if (treated[e.v1]) continue;
      treated[e.v1] = 1;
      float* n1 = &nrm[3 * (size_t)e.v1];
      if (dot3(&nrm[3 * (size_t)e.v0], n1) < 0.f) { n1[0] = -n1[0]; n1[1] = -n1[1]; n1[2] = -n1[2]; }
      push_neighbours(e.v1);
    }
  }
}

struct DeviceBuf {
  void* p = nullptr;
  ~DeviceBuf() { cudaFree(p); }
  template <class T> T* as() const { return (T*)p; }
  bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4) == cudaSuccess; }
};

// The k nearest neighbours and the unoriented normals of a cloud, left on the device.
int knn_normals_device(j3dg_cloud* c, uint32_t k, DeviceBuf& d_knn, DeviceBuf& d_nrm) {
  j3dg_ctx* ctx = c->ctx;
  const uint32_t n = c->n;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int kk = (int)std::min<uint32_t>(k, n);

  DeviceBuf d_box, d_keys_a, d_keys_b, d_vals_a, d_vals_b, d_scratch, d_sorted, d_first, d_count;
  bool mem = d_box.alloc(32) && d_keys_a.alloc((size_t)n * 8) && d_keys_b.alloc((size_t)n * 8) && d_vals_a.alloc((size_t)n * 4) && d_vals_b.alloc((size_t)n * 4) &&
             d_scratch.alloc(rsort::scratch_bytes(n)) && d_sorted.alloc((size_t)n * 16) && d_knn.alloc((size_t)n * k * 4) && d_nrm.alloc((size_t)n * 12);
  if (!mem) { cudaGetLastError(); j3dg_set_error(ctx, "out of device memory (k-NN)"); return J3DG_ENOMEM; }

  const uint32_t init[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
  uint32_t box[8];
  CU_CHECK(ctx, cudaMemcpyAsync(d_box.p, init, 32, cudaMemcpyHostToDevice, st));
  bbox_kernel<<<std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(c->d_pos, n, d_box.as<uint32_t>());
  KERNEL_CHECK(ctx);
  CU_CHECK(ctx, cudaMemcpyAsync(box, d_box.p, 32, cudaMemcpyDeviceToHost, st));
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) { lo[a] = float_unflip(box[a]); hi[a] = float_unflip(box[3 + a]); }
  const float ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  const float emax = std::max(ext[0], std::max(ext[1], ext[2]));

  // Cell size: start from "n / 2 cells in the box" and refine until an occupied cell holds about k / 4 points (scanned
  // surfaces fill a thin shell of the box, so the first guess is usually too coarse).
  const size_t max_cells = (size_t)1 << 27;
  const double target = std::max(1.0, kk / 4.0);
  double cell = emax > 0.f ? std::cbrt(std::max(1e-30, (double)std::max(ext[0], emax * 1e-3f) * std::max(ext[1], emax * 1e-3f) * std::max(ext[2], emax * 1e-3f)) / std::max(1.0, n / 2.0)) : 1.0;
  Grid g;
  for (int attempt = 0; attempt < 4; ++attempt) {
    size_t cells;
    for (;;) {
      cells = 1;
      for (int a = 0; a < 3; ++a) { g.dim[a] = (int)std::min(2048.0, std::floor(ext[a] / cell) + 1.0); cells *= (size_t)g.dim[a]; }
      if (cells <= max_cells) break;
      cell *= 1.26;
    }
    for (int a = 0; a < 3; ++a) g.lo[a] = lo[a];
    g.cell = (float)cell;
    g.inv_cell = 1.f / g.cell;
    g.eps = 16.f * FLT_EPSILON * std::max(emax, std::max(std::fabs(lo[0]) + std::fabs(lo[1]) + std::fabs(lo[2]), std::fabs(hi[0]) + std::fabs(hi[1]) + std::fabs(hi[2])));
    cell_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(c->d_pos, n, g, d_keys_a.as<uint64_t>());
    KERNEL_CHECK(ctx);
    int bits = 1;
    while (((size_t)1 << bits) < cells) ++bits;
    bool in_b = false;
    int rc = rsort::sort_pairs(ctx, d_keys_a.as<uint64_t>(), d_vals_a.as<uint32_t>(), d_keys_b.as<uint64_t>(), d_vals_b.as<uint32_t>(), n, bits, d_scratch.as<uint32_t>(), &in_b);
    if (rc != J3DG_OK) return rc;
    cudaFree(d_first.p); cudaFree(d_count.p); d_first.p = d_count.p = nullptr;
    if (!d_first.alloc(cells * 4) || !d_count.alloc(cells * 4)) { cudaGetLastError(); j3dg_set_error(ctx, "out of device memory (k-NN grid)"); return J3DG_ENOMEM; }
    CU_CHECK(ctx, cudaMemsetAsync(d_count.p, 0, cells * 4, st));
    CU_CHECK(ctx, cudaMemsetAsync(d_box.p, 0, 4, st));
    cell_table_kernel<<<(n + 255) / 256, 256, 0, st>>>(in_b ? d_keys_b.as<uint64_t>() : d_keys_a.as<uint64_t>(), in_b ? d_vals_b.as<uint32_t>() : d_vals_a.as<uint32_t>(), c->d_pos, n,
                                                       d_sorted.as<float4>(), d_first.as<uint32_t>(), d_count.as<uint32_t>(), d_box.as<uint32_t>());
    KERNEL_CHECK(ctx);
    uint32_t occupied = 0;
    CU_CHECK(ctx, cudaMemcpyAsync(&occupied, d_box.p, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaStreamSynchronize(st));
    const double per_cell = (double)n / std::max(1u, occupied);
    if (attempt == 3 || per_cell <= 2.0 * target || emax <= 0.f) break;
    const double shrink = std::pow(per_cell / target, -1.0 / 2.5);
    bool can_shrink = false;
    for (int a = 0; a < 3; ++a) can_shrink = can_shrink || (std::floor(ext[a] / (cell * shrink)) + 1.0 <= 2048.0 && ext[a] > 0.f);
    if (!can_shrink) break;
    cell *= shrink;
  }
  knn_kernel<<<(n + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, st>>>(d_sorted.as<float4>(), n, g, d_first.as<uint32_t>(), d_count.as<uint32_t>(), kk, d_knn.as<uint32_t>());
  KERNEL_CHECK(ctx);
  fit_plane_kernel<<<(n + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, st>>>(c->d_pos, d_knn.as<uint32_t>(), n, kk, d_nrm.as<float>());
  KERNEL_CHECK(ctx);
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  return J3DG_OK;
}

int check_args(j3dg_cloud* c, uint32_t k, const char* who) {
  if (!c || !c->ctx) return J3DG_EINVAL;
  if (k < 1 || k > KNN_MAX_K) { j3dg_set_error(c->ctx, std::string(who) + ": k must be in 1..64"); return J3DG_EINVAL; }
  if (!c->n) { j3dg_set_error(c->ctx, std::string(who) + ": empty cloud"); return J3DG_EINVAL; }
  return J3DG_OK;
}

int adopt_normals(j3dg_cloud* c, const void* src, cudaMemcpyKind kind) {
  if (!c->d_nrm) CU_CHECK(c->ctx, cudaMalloc((void**)&c->d_nrm, (size_t)c->n * 12));
  CU_CHECK(c->ctx, cudaMemcpy(c->d_nrm, src, (size_t)c->n * 12, kind));
  return J3DG_OK;
}

}  // namespace

J3DG_API int j3dg_cloud_knn_normals(j3dg_cloud* c, uint32_t k, float* normals_out, uint32_t* knn_out) {
  int rc = check_args(c, k, "j3dg_cloud_knn_normals");
  if (rc != J3DG_OK) return rc;
  DeviceBuf d_knn, d_nrm;
  if ((rc = knn_normals_device(c, k, d_knn, d_nrm)) != J3DG_OK) return rc;
  const uint32_t kk = std::min(k, c->n);
  if (normals_out) CU_CHECK(c->ctx, cudaMemcpy(normals_out, d_nrm.p, (size_t)c->n * 12, cudaMemcpyDefault));
  if (knn_out) CU_CHECK(c->ctx, cudaMemcpy(knn_out, d_knn.p, (size_t)c->n * kk * 4, cudaMemcpyDefault));
  return adopt_normals(c, d_nrm.p, cudaMemcpyDeviceToDevice);
}

J3DG_API int j3dg_cloud_estimate_normals(j3dg_cloud* c, uint32_t k, float* normals_out) {
  int rc = check_args(c, k, "j3dg_cloud_estimate_normals");
  if (rc != J3DG_OK) return rc;
  DeviceBuf d_knn, d_nrm;
  if ((rc = knn_normals_device(c, k, d_knn, d_nrm)) != J3DG_OK) return rc;
  const uint32_t kk = std::min(k, c->n);
  std::vector<float> nrm((size_t)c->n * 3);
  std::vector<uint32_t> knn((size_t)c->n * kk);
  CU_CHECK(c->ctx, cudaMemcpy(nrm.data(), d_nrm.p, nrm.size() * 4, cudaMemcpyDeviceToHost));
  CU_CHECK(c->ctx, cudaMemcpy(knn.data(), d_knn.p, knn.size() * 4, cudaMemcpyDeviceToHost));
  orient_normals(nrm, knn, c->n, (int)kk);
  if (normals_out) CU_CHECK(c->ctx, cudaMemcpy(normals_out, nrm.data(), nrm.size() * 4, cudaMemcpyDefault));
  return adopt_normals(c, nrm.data(), cudaMemcpyHostToDevice);
}
