// allhits.cu — the second ray-query client of the device BVH:
//   j3dg_mesh_find_all   every triangle a ray crosses — qbvh::find_all_triangles (jtk/qbvh.h:1854-2000)
//   j3dg_mesh_voxelize   the voxel export built on it — _write_vox (j3d/vox.cpp:270-377): three axis-aligned
//                        ray grids, every hit colours the voxel that contains the hit point
// One ray per lane, unordered traversal (every child box the ray pierces is pushed; no hit shrinks the
// interval), stack in shared memory.  The box test is the conservative quantised-slab test of the cast
// kernel (traverse.cuh); the triangle test is the reference's Woop test evaluated with separately rounded
// operations in the reference's order (qbvh.h:4793-4869), against the ORIGINAL interval (t_near, t_far).
#include "common.cuh"
#include "traverse.cuh"
#include "sort.cuh"

#include <algorithm>
#include <cstring>

namespace {

constexpr int AH_THREADS = 64;
constexpr int AH_STACK = 64;  // entries per ray; 64 * 4 B * 64 lanes = 16 KB shared memory per block (96: 4.06 ms, 64: 3.56 ms, 48: 3.48 ms
                              // for the 512-scale voxel grid of the 28 M-triangle mesh); an overflow is reported, never silent

enum AllMode { COUNT = 0, FILL = 1, VOXEL = 2 };

struct AllParams {
  const WideNode* nodes;
  const TriRec* tris;
  uint32_t nt;
  // COUNT / FILL
  const float* rays;          // n x {ox, oy, oz, dx, dy, dz, t_near, t_far}
  uint32_t n;
  uint32_t* offsets;          // COUNT: per-ray hit count is written here; FILL: exclusive prefix sums
  float4* hits;               // FILL: {u, v, t, 0}
  uint32_t* ids;              // FILL: original triangle index
  uint32_t* overflow;
  // VOXEL (vox.cpp:283-377)
  int direction_dim;
  uint32_t dim[3];
  float mn[3], mx[3];
  uint8_t* data;
  const uint32_t* indices;
  const float* vertex_colors;  // nullable
  const float* uv;             // nullable (nt x 6)
  const uint32_t* texture;     // nullable
  uint32_t tex_w, tex_h, tex_stride;
};

__device__ __forceinline__ uint8_t color_to_index(uint32_t r, uint32_t g, uint32_t b) {  // vox.cpp:154-172
  r = r < 16u ? 0u : r - 16u;
  g = g < 16u ? 0u : g - 16u;
  b = b < 32u ? 0u : b - 32u;
  const uint32_t ret = (((r >> 5) << 5) | ((g >> 5) << 2) | (b >> 6)) & 0xFFu;
  return (uint8_t)(ret == 0u ? 1u : ret);
}

// byte-wise max through a CAS on the containing word: several rays may colour the same voxel, and the
// reference's "last writer wins" across its worker threads is not reproducible — the largest palette index wins here
__device__ __forceinline__ void voxel_max(uint8_t* data, size_t idx, uint32_t value) {
  uint32_t* word = reinterpret_cast<uint32_t*>(data + (idx & ~(size_t)3));
  const uint32_t shift = (uint32_t)(idx & 3) * 8u;
  uint32_t old = *word;
  for (;;) {
    if (((old >> shift) & 0xFFu) >= value) return;
    const uint32_t assumed = old;
    old = atomicCAS(word, assumed, (assumed & ~(0xFFu << shift)) | (value << shift));
    if (old == assumed) return;
  }
}

template <int MODE>
__global__ void __launch_bounds__(AH_THREADS) allhits_kernel(const AllParams p) {
  __shared__ uint32_t s_stack[AH_STACK * AH_THREADS];
  uint32_t* const stk = s_stack + threadIdx.x;  // entry i at stk[i * AH_THREADS]
  const uint32_t i = blockIdx.x * AH_THREADS + threadIdx.x;
  if (i >= p.n) return;
  // ---- the ray ----
  float ox, oy, oz, dx, dy, dz, t_near, t_far;
  if (MODE == VOXEL) {  // vox.cpp:311-329; direction (0, 0, 2) along z
    const int dd = p.direction_dim, d1i = (dd + 1) % 3, d2i = (dd + 2) % 3;
    const uint32_t d1 = i / p.dim[d2i], d2 = i % p.dim[d2i];
    float s[3];
    s[dd] = p.mn[dd];
    s[d1i] = fadd(fmul(fdiv(fadd((float)d1, 0.5f), (float)p.dim[d1i]), fsub(p.mx[d1i], p.mn[d1i])), p.mn[d1i]);
    s[d2i] = fadd(fmul(fdiv(fadd((float)d2, 0.5f), (float)p.dim[d2i]), fsub(p.mx[d2i], p.mn[d2i])), p.mn[d2i]);
    ox = s[0]; oy = s[1]; oz = s[2];
    dx = dd == 0 ? 1.f : 0.f; dy = dd == 1 ? 1.f : 0.f; dz = dd == 2 ? 2.f : 0.f;
    t_near = 0.f; t_far = FLT_MAX;
  } else {
    const float* q = p.rays + 8 * (size_t)i;
    ox = __ldg(q); oy = __ldg(q + 1); oz = __ldg(q + 2);
    dx = __ldg(q + 3); dy = __ldg(q + 4); dz = __ldg(q + 5);
    t_near = __ldg(q + 6); t_far = __ldg(q + 7);
  }
  // intersect_woop_precompute (qbvh.h:4793-4823)
  int kz = 2;
  {
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    if (ax > ay) { if (ax > az) kz = 0; }
    else { if (ay > az) kz = 1; }
  }
  int kx = kz == 2 ? 0 : kz + 1;
  int ky = kx == 2 ? 0 : kx + 1;
  const float dkz = pick(dx, dy, dz, kz);
  if (dkz < 0.f) { const int t = kx; kx = ky; ky = t; }
  const float Sz = fdiv(1.f, dkz);
  const float Sx = fmul(pick(dx, dy, dz, kx), Sz);
  const float Sy = fmul(pick(dx, dy, dz, ky), Sz);
  const float idx = safe_rcp(dx), idy = safe_rcp(dy), idz = safe_rcp(dz);
  const uint32_t sel_nx = plane_sel(idx < 0.f ? 3u : 0u), sel_fx = plane_sel(idx < 0.f ? 0u : 3u);
  const uint32_t sel_ny = plane_sel(idy < 0.f ? 4u : 1u), sel_fy = plane_sel(idy < 0.f ? 1u : 4u);
  const uint32_t sel_nz = plane_sel(idz < 0.f ? 5u : 2u), sel_fz = plane_sel(idz < 0.f ? 2u : 5u);

  uint32_t count = 0;
  const uint32_t out_base = MODE == FILL ? p.offsets[i] : 0u;
  int sp = 0;
  uint32_t cur = p.nt ? 0u : J3DG_EMPTY_CHILD;
  while (cur != J3DG_EMPTY_CHILD) {
    if (!(cur & J3DG_LEAF_BIT)) {
      // ---- inner node: push every child box the ray pierces ----
      const char* np = reinterpret_cast<const char*>(p.nodes + cur);
      const U32x8 q0 = ldg256(np), q1 = ldg256(np + 32), q2 = ldg256(np + 64), q3 = ldg256(np + 96);
      const uint4 h0 = make_uint4(q0.v[0], q0.v[1], q0.v[2], q0.v[3]), h1 = make_uint4(q0.v[4], q0.v[5], q0.v[6], q0.v[7]);
      const uint4 b[4] = {make_uint4(q1.v[0], q1.v[1], q1.v[2], q1.v[3]), make_uint4(q1.v[4], q1.v[5], q1.v[6], q1.v[7]),
                          make_uint4(q2.v[0], q2.v[1], q2.v[2], q2.v[3]), make_uint4(q2.v[4], q2.v[5], q2.v[6], q2.v[7])};
      const uint4 c[2] = {make_uint4(q3.v[0], q3.v[1], q3.v[2], q3.v[3]), make_uint4(q3.v[4], q3.v[5], q3.v[6], q3.v[7])};
      const Slab X = slab(__uint_as_float(h1.x), __uint_as_float(h0.x), ox, idx);
      const Slab Y = slab(__uint_as_float(h1.y), __uint_as_float(h0.y), oy, idy);
      const Slab Z = slab(__uint_as_float(h1.z), __uint_as_float(h0.z), oz, idz);
      auto test_push = [&](uint32_t lo, uint32_t hi, uint32_t ref) {
        float tmin = fmaxf(fmaxf(fmaf(plane(lo, hi, sel_nx), X.S, X.Bn), fmaf(plane(lo, hi, sel_ny), Y.S, Y.Bn)), fmaxf(fmaf(plane(lo, hi, sel_nz), Z.S, Z.Bn), t_near));
        float tmax = fminf(fminf(fmaf(plane(lo, hi, sel_fx), X.S, X.Bf), fmaf(plane(lo, hi, sel_fy), Y.S, Y.Bf)), fminf(fmaf(plane(lo, hi, sel_fz), Z.S, Z.Bf), t_far));
        tmin = fmaf(-fabsf(tmin), 2e-6f, tmin);  // conservative padding against rounding of the slab arithmetic
        tmax = fmaf(fabsf(tmax), 2e-6f, tmax);
        if (tmin <= tmax && ref != J3DG_EMPTY_CHILD) {  // empty slots have inverted boxes and never pass
          if (sp < AH_STACK) stk[sp++ * AH_THREADS] = ref;
          else *p.overflow = 1u;
        }
      };
      test_push(b[0].x, b[0].y, c[0].x); test_push(b[0].z, b[0].w, c[0].y);
      test_push(b[1].x, b[1].y, c[0].z); test_push(b[1].z, b[1].w, c[0].w);
      test_push(b[2].x, b[2].y, c[1].x); test_push(b[2].z, b[2].w, c[1].y);
      test_push(b[3].x, b[3].y, c[1].z); test_push(b[3].z, b[3].w, c[1].w);
    } else {
      // ---- leaf: 1..8 consecutive records, the last one flagged ----
      uint32_t slot = cur & J3DG_LEAF_FIRST_MASK;
      for (;;) {
        const float4* tp = reinterpret_cast<const float4*>(p.tris + slot);
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        // one lane of intersect_woop (qbvh.h:4825-4869); 1/det is correctly rounded instead of rcpps + NR
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
        const bool inside = ((U <= 0.f) && (V <= 0.f) && (W <= 0.f)) || ((U >= 0.f) && (V >= 0.f) && (W >= 0.f));
        const float det = fadd(fadd(U, V), W);
        if (inside && det != 0.f) {
          const float inv_det = fdiv(1.f, det);
          const float Az = fmul(Sz, Akz), Bz = fmul(Sz, Bkz), Cz = fmul(Sz, Ckz);
          const float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
          const float t = fmul(T, inv_det);
          if ((t_far > t) && (t > t_near)) {
            const float u = fmul(V, inv_det), v = fmul(W, inv_det);
            const uint32_t tri = __float_as_uint(v0.w);
            if (MODE == FILL) {
              p.hits[out_base + count] = make_float4(u, v, t, 0.f);
              p.ids[out_base + count] = tri;
            } else if (MODE == VOXEL) {
              // vox.cpp:336-375
              const float k = fsub(fsub(1.f, u), v);
              const float px = fadd(fadd(fmul(v0.x, k), fmul(u, v1.x)), fmul(v, v2.x));
              const float py = fadd(fadd(fmul(v0.y, k), fmul(u, v1.y)), fmul(v, v2.y));
              const float pz = fadd(fadd(fmul(v0.z, k), fmul(u, v1.z)), fmul(v, v2.z));
              uint32_t r = 255u, g = 255u, b = 255u;  // clr = (1, 1, 1)
              if (p.uv != nullptr && p.texture != nullptr && p.tex_w > 0u && p.tex_h > 0u) {
                const float* uvc = p.uv + 6 * (size_t)tri;
                float cx = fadd(fadd(fmul(k, uvc[0]), fmul(u, uvc[2])), fmul(v, uvc[4]));
                float cy = fadd(fadd(fmul(k, uvc[1]), fmul(u, uvc[3])), fmul(v, uvc[5]));
                cx = fmaxf(fminf(cx, 1.f), 0.f);
                cy = fmaxf(fminf(cy, 1.f), 0.f);
                const int X = __float2int_rz(fmul(cx, (float)(p.tex_w - 1u))), Y = __float2int_rz(fmul(cy, (float)(p.tex_h - 1u)));
                const uint32_t color = p.texture[(size_t)Y * p.tex_stride + X];
                r = (uint32_t)__float2int_rz(fmul(fdiv((float)(color & 255u), 255.f), 255.f)) & 0xFFu;
                g = (uint32_t)__float2int_rz(fmul(fdiv((float)((color >> 8) & 255u), 255.f), 255.f)) & 0xFFu;
                b = (uint32_t)__float2int_rz(fmul(fdiv((float)((color >> 16) & 255u), 255.f), 255.f)) & 0xFFu;
              } else if (p.vertex_colors != nullptr) {
                const uint32_t* id = p.indices + 3 * (size_t)tri;
                const float* c0 = p.vertex_colors + 3 * (size_t)id[0];
                const float* c1 = p.vertex_colors + 3 * (size_t)id[1];
                const float* c2 = p.vertex_colors + 3 * (size_t)id[2];
                r = (uint32_t)__float2int_rz(fmul(fadd(fadd(fmul(c0[0], k), fmul(u, c1[0])), fmul(v, c2[0])), 255.f)) & 0xFFu;
                g = (uint32_t)__float2int_rz(fmul(fadd(fadd(fmul(c0[1], k), fmul(u, c1[1])), fmul(v, c2[1])), 255.f)) & 0xFFu;
                b = (uint32_t)__float2int_rz(fmul(fadd(fadd(fmul(c0[2], k), fmul(u, c1[2])), fmul(v, c2[2])), 255.f)) & 0xFFu;
              }
              uint32_t Xv = __float2uint_rz(fmul(fdiv(fsub(px, p.mn[0]), fsub(p.mx[0], p.mn[0])), (float)p.dim[0]));
              uint32_t Yv = __float2uint_rz(fmul(fdiv(fsub(py, p.mn[1]), fsub(p.mx[1], p.mn[1])), (float)p.dim[1]));
              uint32_t Zv = __float2uint_rz(fmul(fdiv(fsub(pz, p.mn[2]), fsub(p.mx[2], p.mn[2])), (float)p.dim[2]));
              if (Xv == p.dim[0]) Xv = p.dim[0] - 1u;
              if (Yv == p.dim[1]) Yv = p.dim[1] - 1u;
              if (Zv == p.dim[2]) Zv = p.dim[2] - 1u;
              if (Xv < p.dim[0] && Yv < p.dim[1] && Zv < p.dim[2])  // the reference would write out of bounds otherwise
                voxel_max(p.data, (size_t)Xv + ((size_t)Yv + (size_t)Zv * p.dim[1]) * p.dim[0], color_to_index(r, g, b));
            }
            ++count;
          }
        }
        if (__float_as_uint(v1.w) != 0u) break;  // end of leaf
        ++slot;
      }
    }
    cur = sp > 0 ? stk[--sp * AH_THREADS] : J3DG_EMPTY_CHILD;
  }
  if (MODE == COUNT) p.offsets[i] = count;
}

int fill_common(j3dg_mesh* m, AllParams& p) {
  p.nodes = m->d_nodes;
  p.tris = m->d_tris;
  p.nt = (m->d_nodes && m->d_tris) ? m->nt : 0u;
  p.overflow = reinterpret_cast<uint32_t*>(m->ctx->d_stats + 2);
  return J3DG_OK;
}

}  // namespace

J3DG_API int j3dg_mesh_find_all(j3dg_mesh* mesh, const float* rays, uint32_t n, uint32_t* offsets, float* hits, uint32_t* triangle_ids,
                                uint32_t capacity, uint32_t* total) {
  if (!mesh || !mesh->ctx) return J3DG_EINVAL;
  j3dg_ctx* ctx = mesh->ctx;
  if ((n && !rays) || !offsets || (hits && !triangle_ids) || (!hits && triangle_ids)) { j3dg_set_error(ctx, "j3dg_mesh_find_all: bad argument"); return J3DG_EINVAL; }
  if (n >= 0x7FFFFFFFu) { j3dg_set_error(ctx, "j3dg_mesh_find_all: too many rays"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  if (total) *total = 0;
  if (!n) {  // offsets may be a device pointer
    const uint32_t zero = 0;
    CU_CHECK(ctx, cudaMemcpy(offsets, &zero, sizeof(zero), cudaMemcpyDefault));
    return J3DG_OK;
  }
  const bool rays_dev = j3dg_is_device_ptr(rays), out_dev = j3dg_is_device_ptr(offsets);
  if (hits && (j3dg_is_device_ptr(hits) != out_dev || j3dg_is_device_ptr(triangle_ids) != out_dev)) {
    j3dg_set_error(ctx, "j3dg_mesh_find_all: outputs must be all host or all device");
    return J3DG_EINVAL;
  }
  float* d_rays = nullptr;
  uint32_t* d_work = nullptr;  // offsets (n + 1) + scan sums
  const size_t work_n = (size_t)n + 1 + rsort::scan_scratch_count((size_t)n + 1);
  cudaError_t e = cudaMalloc((void**)&d_work, work_n * sizeof(uint32_t));
  if (e == cudaSuccess && !rays_dev) e = cudaMalloc((void**)&d_rays, (size_t)n * 8 * sizeof(float));
  if (e != cudaSuccess) { cudaFree(d_work); return j3dg_cuda_fail(ctx, e, "cudaMalloc", __FILE__, __LINE__); }
  int rc = J3DG_OK;
  float4* d_hits = nullptr;
  uint32_t* d_ids = nullptr;
  uint32_t sum = 0;
  auto cleanup = [&]() { cudaFree(d_work); cudaFree(d_rays); if (!out_dev) { cudaFree(d_hits); cudaFree(d_ids); } };
#define AH_CHECK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return j3dg_cuda_fail(ctx, e__, #call, __FILE__, __LINE__); } } while (0)
  if (!rays_dev) AH_CHECK(cudaMemcpyAsync(d_rays, rays, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  AH_CHECK(cudaMemsetAsync(ctx->d_stats + 2, 0, sizeof(unsigned long long), ctx->stream));
  AllParams p = {};
  fill_common(mesh, p);
  p.rays = rays_dev ? rays : d_rays;
  p.n = n;
  p.offsets = d_work;
  const uint32_t blocks = (n + AH_THREADS - 1) / AH_THREADS;
  AH_CHECK(cudaMemsetAsync(d_work + n, 0, sizeof(uint32_t), ctx->stream));  // scanning n + 1 entries leaves the total in offsets[n]
  allhits_kernel<COUNT><<<blocks, AH_THREADS, 0, ctx->stream>>>(p);
  ctx->launches++;
  AH_CHECK(cudaGetLastError());
  if ((rc = rsort::exclusive_scan_u32(ctx, d_work, (size_t)n + 1, d_work + n + 1)) != J3DG_OK) { cleanup(); return rc; }
  AH_CHECK(cudaMemcpyAsync(&sum, d_work + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AH_CHECK(cudaStreamSynchronize(ctx->stream));
  if (total) *total = sum;
  AH_CHECK(cudaMemcpyAsync(offsets, d_work, ((size_t)n + 1) * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
  if (hits) {
    if (capacity < sum) {
      AH_CHECK(cudaStreamSynchronize(ctx->stream));
      cleanup();
      j3dg_set_error(ctx, "j3dg_mesh_find_all: capacity too small (see *total); offsets are valid");
      return J3DG_EINVAL;
    }
    if (sum) {
      if (out_dev) { d_hits = (float4*)hits; d_ids = triangle_ids; }
      else {
        AH_CHECK(cudaMalloc((void**)&d_hits, (size_t)sum * sizeof(float4)));
        AH_CHECK(cudaMalloc((void**)&d_ids, (size_t)sum * sizeof(uint32_t)));
      }
      p.hits = d_hits; p.ids = d_ids;
      allhits_kernel<FILL><<<blocks, AH_THREADS, 0, ctx->stream>>>(p);
      ctx->launches++;
      AH_CHECK(cudaGetLastError());
      if (!out_dev) {
        AH_CHECK(cudaMemcpyAsync(hits, d_hits, (size_t)sum * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        AH_CHECK(cudaMemcpyAsync(triangle_ids, d_ids, (size_t)sum * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
      }
    }
  }
  uint32_t ovf = 0;
  AH_CHECK(cudaMemcpyAsync(&ovf, ctx->d_stats + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AH_CHECK(cudaStreamSynchronize(ctx->stream));
  cleanup();
#undef AH_CHECK
  if (ovf) { j3dg_set_error(ctx, "traversal stack overflow (BVH deeper than the kernel's stack)"); return J3DG_ECUDA; }
  return J3DG_OK;
}

// vox.cpp:274-289
J3DG_API int j3dg_mesh_voxel_dims(const j3dg_mesh* mesh, uint32_t max_dim, uint32_t dims_out[3]) {
  if (!mesh || !dims_out) return J3DG_EINVAL;
  const float* mn = mesh->info.bbox_min;
  const float* mx = mesh->info.bbox_max;
  volatile float ext[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
  int largest = 0;
  if (ext[1] > ext[largest]) largest = 1;
  if (ext[2] > ext[largest]) largest = 2;
  for (int j = 0; j < 3; ++j) {
    volatile float a = (float)max_dim * ext[j];
    volatile float b = a / ext[largest];
    dims_out[j] = (uint32_t)b;
    if (dims_out[j] == 0) dims_out[j] = 1;
  }
  return J3DG_OK;
}

J3DG_API int j3dg_mesh_voxelize(j3dg_mesh* mesh, uint32_t max_dim, uint32_t dims_out[3], uint8_t* data, size_t capacity) {
  if (!mesh || !mesh->ctx || !dims_out) return J3DG_EINVAL;
  j3dg_ctx* ctx = mesh->ctx;
  if (!max_dim) { j3dg_set_error(ctx, "j3dg_mesh_voxelize: max_dim = 0"); return J3DG_EINVAL; }
  if (!mesh->nv) { j3dg_set_error(ctx, "j3dg_mesh_voxelize: mesh without vertices"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  uint32_t dim[3];
  j3dg_mesh_voxel_dims(mesh, max_dim, dim);
  for (int j = 0; j < 3; ++j) dims_out[j] = dim[j];
  if (!data) return J3DG_OK;  // size query
  const size_t nvox = (size_t)dim[0] * dim[1] * dim[2];
  if (capacity < nvox) { j3dg_set_error(ctx, "j3dg_mesh_voxelize: capacity too small (see dims_out)"); return J3DG_EINVAL; }
  const bool out_dev = j3dg_is_device_ptr(data);
  uint8_t* d_data = nullptr;
  const size_t padded = (nvox + 3) & ~(size_t)3;  // voxel_max works on whole words
  if (out_dev && ((uintptr_t)data & 3u) == 0 && capacity >= padded) d_data = data;
  else CU_CHECK(ctx, cudaMalloc((void**)&d_data, padded));
  cudaError_t e = cudaMemsetAsync(d_data, 0, padded, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_stats + 2, 0, sizeof(unsigned long long), ctx->stream);
  AllParams p = {};
  fill_common(mesh, p);
  for (int j = 0; j < 3; ++j) { p.dim[j] = dim[j]; p.mn[j] = mesh->info.bbox_min[j]; p.mx[j] = mesh->info.bbox_max[j]; }
  p.data = d_data;
  p.indices = mesh->d_indices;
  p.vertex_colors = mesh->d_vcolors;
  p.uv = mesh->d_uv;
  p.texture = mesh->d_texture;
  p.tex_w = mesh->tex_w; p.tex_h = mesh->tex_h; p.tex_stride = mesh->tex_w;
  for (int dd = 0; dd < 3 && e == cudaSuccess; ++dd) {
    p.direction_dim = dd;
    const size_t nr = (size_t)dim[(dd + 1) % 3] * dim[(dd + 2) % 3];
    p.n = (uint32_t)nr;
    allhits_kernel<VOXEL><<<(uint32_t)((nr + AH_THREADS - 1) / AH_THREADS), AH_THREADS, 0, ctx->stream>>>(p);
    ctx->launches++;
    e = cudaGetLastError();
  }
  uint32_t ovf = 0;
  if (e == cudaSuccess && d_data != data) e = cudaMemcpyAsync(data, d_data, nvox, cudaMemcpyDefault, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&ovf, ctx->d_stats + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (d_data != data) cudaFree(d_data);
  if (e != cudaSuccess) return j3dg_cuda_fail(ctx, e, "j3dg_mesh_voxelize", __FILE__, __LINE__);
  if (ovf) { j3dg_set_error(ctx, "traversal stack overflow (BVH deeper than the kernel's stack)"); return J3DG_ECUDA; }
  return J3DG_OK;
}
