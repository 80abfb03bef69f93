// pick.cu — picking consumers of the pixel buffer, evaluated on the device so that an interactive
// host only downloads the RGBA image per frame and asks for the few records it needs.
// Replaces, per query pixel (x, y):
//   canvas::get_pixel                    j3d/canvas.cpp:141-153   the 32-byte record itself
//   view::get_id                         j3d/view.cpp:483-492     db_id (0 = nothing under the cursor)
//   view::get_world_position             j3d/view.cpp:439-469     barycentric point of the hit triangle, object -> world
//   view::get_index / get_closest_vertex j3d/view.cpp:471-481, j3d/pixel.cpp:6-33   nearest corner of the hit triangle
//   canvas::do_mouse pivot pick          j3d/canvas.cpp:157-179   origin + depth * dir of the pixel's primary ray
// All arithmetic is separately rounded (no FMA) in the reference's operation order.
#include "common.cuh"

#include <cstring>
#include <vector>

namespace {

struct PickObject {
  const float* vertices;     // mesh: nv x 3; cloud: n x 3 positions
  const uint32_t* indices;   // mesh: nt x 3; cloud: nullptr
  uint32_t count;            // triangles / points
  uint32_t db_id;
  float cs[16];
};

__device__ __forceinline__ float3 bary_point(const float* __restrict__ V, const uint32_t* __restrict__ I, uint32_t tri, float u, float v,
                                             uint32_t& v0, uint32_t& v1, uint32_t& v2, float3& A, float3& B, float3& C) {
  v0 = I[3 * (size_t)tri]; v1 = I[3 * (size_t)tri + 1]; v2 = I[3 * (size_t)tri + 2];
  A = make_float3(V[3 * (size_t)v0], V[3 * (size_t)v0 + 1], V[3 * (size_t)v0 + 2]);
  B = make_float3(V[3 * (size_t)v1], V[3 * (size_t)v1 + 1], V[3 * (size_t)v1 + 2]);
  C = make_float3(V[3 * (size_t)v2], V[3 * (size_t)v2 + 1], V[3 * (size_t)v2 + 2]);
  const float k = fsub(fsub(1.f, u), v);  // V0 * (1 - u - v) + u * V1 + v * V2, left to right
  float3 p;
  p.x = fadd(fadd(fmul(A.x, k), fmul(u, B.x)), fmul(v, C.x));
  p.y = fadd(fadd(fmul(A.y, k), fmul(u, B.y)), fmul(v, C.y));
  p.z = fadd(fadd(fmul(A.z, k), fmul(u, B.z)), fmul(v, C.z));
  return p;
}

__device__ __forceinline__ float dot3(float3 a) { return fadd(fadd(fmul(a.x, a.x), fmul(a.y, a.y)), fmul(a.z, a.z)); }

__global__ void pick_kernel(const j3dg_pixel* __restrict__ px, uint32_t stride, ViewDev vw, const PickObject* __restrict__ objs, uint32_t nobj,
                            const int32_t* __restrict__ xy, uint32_t n, j3dg_pick_result* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float qnan = __int_as_float(0x7fc00000);
  j3dg_pick_result r;
  memset(&r, 0, sizeof(r));
  r.world_pos[0] = r.world_pos[1] = r.world_pos[2] = qnan;
  r.pivot[0] = r.pivot[1] = r.pivot[2] = qnan;
  r.closest_vertex = 0xFFFFFFFFu;
  const int x = xy[2 * i], y = xy[2 * i + 1];
  if (x >= 0 && y >= 0 && x < (int)vw.width && y < (int)vw.height) {  // view.cpp:443, 474, 486
    const j3dg_pixel p = px[(size_t)y * stride + x];
    r.pixel = p;
    r.db_id = p.db_id;
    if (p.db_id != 0u) {
      // canvas.cpp:165-177: the pixel's primary ray evaluated at the stored depth
      const float w = (float)vw.width, h = (float)vw.height;
      float4 sp;
      sp.x = fsub(fmul(2.f, fdiv(fadd((float)x, 0.5f), w)), 1.f);
      sp.y = fsub(fmul(2.f, fdiv(fadd((float)y, 0.5f), h)), 1.f);
      sp.z = vw.near_plane;
      sp.w = 1.f;
      float4 dir = mat_vec(vw.pinv, sp);
      dir.w = 0.f;
      dir = mat_vec(vw.cs, dir);
      r.pivot[0] = fadd(vw.origin[0], fmul(p.depth, dir.x));
      r.pivot[1] = fadd(vw.origin[1], fmul(p.depth, dir.y));
      r.pivot[2] = fadd(vw.origin[2], fmul(p.depth, dir.z));
      for (uint32_t k = 0; k < nobj; ++k) {
        const PickObject& o = objs[k];
        if (o.db_id != p.db_id || p.object_id >= o.count) continue;
        if (o.indices) {  // mesh
          uint32_t v0, v1, v2;
          float3 A, B, C;
          const float3 pos = bary_point(o.vertices, o.indices, p.object_id, p.barycentric_u, p.barycentric_v, v0, v1, v2, A, B, C);
          // view.cpp:457-460: the same blend with w = 1 lanes, then m->cs * pos
          const float kk = fsub(fsub(1.f, p.barycentric_u), p.barycentric_v);
          const float pw = fadd(fadd(fmul(1.f, kk), fmul(p.barycentric_u, 1.f)), fmul(p.barycentric_v, 1.f));
          const float4 wp = mat_vec(o.cs, make_float4(pos.x, pos.y, pos.z, pw));
          r.world_pos[0] = wp.x; r.world_pos[1] = wp.y; r.world_pos[2] = wp.z;
          // pixel.cpp:15-25
          const float dA = dot3(make_float3(fsub(pos.x, A.x), fsub(pos.y, A.y), fsub(pos.z, A.z)));
          const float dB = dot3(make_float3(fsub(pos.x, B.x), fsub(pos.y, B.y), fsub(pos.z, B.z)));
          const float dC = dot3(make_float3(fsub(pos.x, C.x), fsub(pos.y, C.y), fsub(pos.z, C.z)));
          r.closest_vertex = (dA < dB) ? ((dA < dC) ? v0 : v2) : ((dB < dC) ? v1 : v2);
        } else {  // point cloud: the point itself (pixel.cpp:28-32; view.cpp:462-467 with the cloud's own cs)
          const float* q = o.vertices + 3 * (size_t)p.object_id;
          const float4 wp = mat_vec(o.cs, make_float4(q[0], q[1], q[2], 1.f));
          r.world_pos[0] = wp.x; r.world_pos[1] = wp.y; r.world_pos[2] = wp.z;
          r.closest_vertex = p.object_id;
        }
        break;
      }
    }
  }
  out[i] = r;
}

}  // namespace

void j3dg_make_view_dev(const j3dg_view* v, ViewDev& d);

J3DG_API int j3dg_pick(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, j3dg_cloud* const* clouds, uint32_t nc, const j3dg_view* view,
                       const j3dg_pixel* pixels, uint32_t pixel_stride, const int32_t* xy, uint32_t n, j3dg_pick_result* out) {
  if (!ctx || !view || !out || (n && !xy) || (nm && !meshes) || (nc && !clouds)) { j3dg_set_error(ctx, "j3dg_pick: bad argument"); return J3DG_EINVAL; }
  if (!n) return J3DG_OK;
  cudaSetDevice(ctx->device);
  const uint32_t w = view->width, h = view->height;
  const j3dg_pixel* d_px = pixels;
  if (!pixels) {  // the canvas of the last rendered frame, still resident
    if (!ctx->last_canvas || ctx->last_w != w || ctx->last_h != h) { j3dg_set_error(ctx, "j3dg_pick: no resident canvas of this size (render a frame first or pass pixels)"); return J3DG_EINVAL; }
    d_px = (const j3dg_pixel*)ctx->last_canvas;
    pixel_stride = w;
  } else if (!j3dg_is_device_ptr(pixels)) { j3dg_set_error(ctx, "j3dg_pick: pixels must be a device buffer (or NULL for the last frame)"); return J3DG_EINVAL; }
  if (!pixel_stride) pixel_stride = w;
  std::vector<PickObject> objs;
  for (uint32_t i = 0; i < nm; ++i) {
    if (!meshes[i]) { j3dg_set_error(ctx, "j3dg_pick: null mesh"); return J3DG_EINVAL; }
    PickObject o;
    o.vertices = meshes[i]->d_vertices; o.indices = meshes[i]->d_indices; o.count = meshes[i]->nt; o.db_id = meshes[i]->db_id;
    memcpy(o.cs, meshes[i]->cs, sizeof(o.cs));
    if (o.vertices && o.indices) objs.push_back(o);
  }
  for (uint32_t i = 0; i < nc; ++i) {
    if (!clouds[i]) { j3dg_set_error(ctx, "j3dg_pick: null cloud"); return J3DG_EINVAL; }
    PickObject o;
    o.vertices = clouds[i]->d_pos; o.indices = nullptr; o.count = clouds[i]->n; o.db_id = clouds[i]->db_id;
    memcpy(o.cs, clouds[i]->cs, sizeof(o.cs));
    if (o.vertices) objs.push_back(o);
  }
  const size_t obj_bytes = (sizeof(PickObject) * std::max<size_t>(objs.size(), 1) + 255) & ~(size_t)255;
  const size_t xy_bytes = ((size_t)n * 2 * sizeof(int32_t) + 255) & ~(size_t)255;
  int rc = j3dg_reserve(ctx, &ctx->d_misc, &ctx->misc_cap, obj_bytes + xy_bytes + (size_t)n * sizeof(j3dg_pick_result));
  if (rc != J3DG_OK) return rc;
  char* base = (char*)ctx->d_misc;
  if (!objs.empty()) CU_CHECK(ctx, cudaMemcpyAsync(base, objs.data(), sizeof(PickObject) * objs.size(), cudaMemcpyHostToDevice, ctx->stream));
  CU_CHECK(ctx, cudaMemcpyAsync(base + obj_bytes, xy, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
  ViewDev vd;
  j3dg_make_view_dev(view, vd);
  j3dg_pick_result* d_out = j3dg_is_device_ptr(out) ? out : (j3dg_pick_result*)(base + obj_bytes + xy_bytes);
  pick_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_px, pixel_stride, vd, (const PickObject*)base, (uint32_t)objs.size(),
                                                        (const int32_t*)(base + obj_bytes), n, d_out);
  KERNEL_CHECK(ctx);
  if (d_out != out) CU_CHECK(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(j3dg_pick_result), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // `objs` and pageable sources live on this frame
  return J3DG_OK;
}
