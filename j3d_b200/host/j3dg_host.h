// j3dg_host.h — host-side (C++) mirror of the j3d interfaces that sit directly above the
// GPU hot path: camera, scene (add_object / prepare_scene / unzoom), matcap and canvas.
// Same names, argument meaning and call order as the reference so that view::render_scene
// (j3d/view.cpp:421-430) can drive `j3dg::canvas` unchanged; all per-pixel / per-triangle /
// per-point work happens behind the C ABI in include/j3dg.h (libj3dg.so, sm_100a CUDA).
// There is no CPU rendering path here: every render entry point fails loudly (throws
// std::runtime_error carrying j3dg_last_error) when the CUDA library reports an error.
#pragma once

#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "j3dg.h"

extern "C" {
// Pure host math exported by libj3dg_host.so (also used from Python through ctypes).
void j3dgh_make_projection(uint32_t w, uint32_t h, float* near_plane, float* projection, float* projection_inv);
void j3dgh_invert_orthonormal(const float* m, float* out);
void j3dgh_matrix_multiply(const float* a, const float* b, float* out);
void j3dgh_matrix_vector_multiply(const float* m, const float* v, float* out);
void j3dgh_transform_bbox(const float* cs, const float* bb_min, const float* bb_max, float* out_min, float* out_max);
void j3dgh_unzoom(const float* bb_min, const float* bb_max, float* diagonal, float* pivot, float* cs, float* cs_inv);
void j3dgh_orbit(const float* cs_inv0, const float* pivot, float angle_deg, float* cs, float* cs_inv);
void j3dgh_make_matcap(int type, uint32_t* out_512x512, uint32_t* cavity_clr);
void j3dgh_fill_background(uint32_t w, uint32_t h, uint32_t stride, uint32_t clr_top, uint32_t clr_bottom, uint32_t* out);
void j3dgh_compute_bb(const float* vertices, uint32_t nv, float* bb_min, float* bb_max);
}

namespace j3dg {

using pixel = j3dg_pixel;  // j3d/pixel.h:11-27, byte-identical

struct float4x4 {  // column-major like jtk::float4x4
  float f[16];
  float& operator[](int i) { return f[i]; }
  const float& operator[](int i) const { return f[i]; }
  static float4x4 identity() {
    float4x4 m;
    std::memset(m.f, 0, sizeof(m.f));
    m.f[0] = m.f[5] = m.f[10] = m.f[15] = 1.f;
    return m;
  }
};

// j3d/mesh.h:25-36 (the members the renderer reads)
struct mesh {
  std::vector<float> vertices;          // 3 per vertex
  std::vector<uint32_t> triangles;      // 3 per triangle
  std::vector<float> vertex_colors;     // optional, 3 per vertex in [0,1]
  std::vector<float> uv_coordinates;    // optional, 6 per triangle
  std::vector<uint32_t> texture;        // optional, texture_w * texture_h, 0xAABBGGRR
  uint32_t texture_w = 0, texture_h = 0;
  float4x4 cs = float4x4::identity();
  double acceleration_structure_construction_time_in_s = 0.0;
};

// j3d/pc.h:23-31
struct pc {
  std::vector<float> vertices;
  std::vector<float> normals;
  std::vector<uint32_t> vertex_colors;
  float4x4 cs = float4x4::identity();
};

// j3d/scene.h:10-52.  scene_object owns the device mesh (where the reference owns the qbvh).
struct scene_object {
  uint32_t db_id = 0;
  const mesh* p_mesh = nullptr;
  float min_bb[3], max_bb[3];
  float4x4 cs;
  j3dg_mesh* bvh = nullptr;  // GPU BVH + resident geometry (replaces std::unique_ptr<jtk::qbvh>)
};
struct scene_pointcloud {
  uint32_t db_id = 0;
  const pc* p_pc = nullptr;
  float min_bb[3], max_bb[3];
  float4x4 cs;
  j3dg_cloud* cloud = nullptr;
};
struct scene {
  float4x4 coordinate_system = float4x4::identity(), coordinate_system_inv = float4x4::identity();
  float pivot[3] = {0, 0, 0};
  float min_bb[3] = {0, 0, 0}, max_bb[3] = {0, 0, 0};
  float diagonal = 0.f;
  std::list<scene_object> objects;
  std::list<scene_pointcloud> pointclouds;
};

// j3d/matcap.h
struct matcap {
  std::vector<uint32_t> im;  // 512 x 512
  uint32_t w = 0, h = 0;
  uint32_t cavity_clr = 0;
};
inline void make_matcap(matcap& m, int type) {  // 0 red wax (default), 1 gray, 2 brown, 3 sketch
  m.w = m.h = 512;
  m.im.resize(512 * 512);
  j3dgh_make_matcap(type, m.im.data(), &m.cavity_clr);
}
inline void make_matcap_red_wax(matcap& m) { make_matcap(m, 0); }

class context {  // one per process / GPU
 public:
  explicit context(int device = 0) {
    if (j3dg_ctx_create(device, &_ctx) != J3DG_OK)
      throw std::runtime_error(std::string("j3dg_ctx_create: ") + j3dg_last_error(nullptr));
  }
  ~context() { j3dg_ctx_destroy(_ctx); }
  context(const context&) = delete;
  context& operator=(const context&) = delete;
  j3dg_ctx* get() const { return _ctx; }
  void check(int rc, const char* what) const {
    if (rc != J3DG_OK) throw std::runtime_error(std::string(what) + ": " + j3dg_last_error(_ctx));
  }

 private:
  j3dg_ctx* _ctx = nullptr;
};

// add_object (j3d/scene.cpp:8-36): uploads the mesh, builds normals + bbox + BVH on the GPU.
inline void add_object(context& ctx, uint32_t db_id, scene& s, mesh& m) {
  scene_object obj;
  obj.db_id = db_id;
  obj.p_mesh = &m;
  obj.cs = m.cs;
  const uint32_t nv = (uint32_t)(m.vertices.size() / 3), nt = (uint32_t)(m.triangles.size() / 3);
  ctx.check(j3dg_mesh_create(ctx.get(), m.vertices.data(), nv, m.triangles.data(), nt,
                             m.vertex_colors.empty() ? nullptr : m.vertex_colors.data(),
                             m.uv_coordinates.empty() ? nullptr : m.uv_coordinates.data(),
                             m.texture.empty() ? nullptr : m.texture.data(), m.texture_w, m.texture_h, m.texture_w,
                             m.cs.f, db_id, &obj.bvh),
            "j3dg_mesh_create");
  j3dg_mesh_info info;
  j3dg_mesh_info_get(obj.bvh, &info);
  std::memcpy(obj.min_bb, info.bbox_min, 12);
  std::memcpy(obj.max_bb, info.bbox_max, 12);
  m.acceleration_structure_construction_time_in_s = (info.build_ms + info.upload_ms) * 1e-3;
  s.objects.emplace_back(obj);
}
inline void add_object(context& ctx, uint32_t db_id, scene& s, pc& p) {
  scene_pointcloud obj;
  obj.db_id = db_id;
  obj.p_pc = &p;
  obj.cs = p.cs;
  const uint32_t n = (uint32_t)(p.vertices.size() / 3);
  ctx.check(j3dg_cloud_create(ctx.get(), p.vertices.data(), p.normals.empty() ? nullptr : p.normals.data(),
                              p.vertex_colors.empty() ? nullptr : p.vertex_colors.data(), n, p.cs.f, db_id, &obj.cloud),
            "j3dg_cloud_create");
  j3dgh_compute_bb(p.vertices.data(), n, obj.min_bb, obj.max_bb);
  s.pointclouds.emplace_back(obj);
}
// remove_object (j3d/scene.cpp:40-48)
inline void remove_object(uint32_t id, scene& s) {
  for (auto it = s.objects.begin(); it != s.objects.end(); ++it)
    if (it->db_id == id) { j3dg_mesh_destroy(it->bvh); s.objects.erase(it); break; }
  for (auto it = s.pointclouds.begin(); it != s.pointclouds.end(); ++it)
    if (it->db_id == id) { j3dg_cloud_destroy(it->cloud); s.pointclouds.erase(it); break; }
}
// prepare_scene (j3d/scene.cpp:50-89)
inline void prepare_scene(scene& s) {
  bool first = true;
  auto merge = [&](const float4x4& cs, const float* mn, const float* mx) {
    float a[3], b[3];
    j3dgh_transform_bbox(cs.f, mn, mx, a, b);
    for (int j = 0; j < 3; ++j) {
      if (first || a[j] < s.min_bb[j]) s.min_bb[j] = a[j];
      if (first || b[j] > s.max_bb[j]) s.max_bb[j] = b[j];
    }
    first = false;
  };
  for (const auto& o : s.objects) merge(o.cs, o.min_bb, o.max_bb);
  for (const auto& o : s.pointclouds) merge(o.cs, o.min_bb, o.max_bb);
  if (first)
    for (int j = 0; j < 3; ++j) s.min_bb[j] = s.max_bb[j] = 0.f;
  s.diagonal = s.max_bb[0] - s.min_bb[0];
  if (s.max_bb[1] - s.min_bb[1] > s.diagonal) s.diagonal = s.max_bb[1] - s.min_bb[1];
  if (s.max_bb[2] - s.min_bb[2] > s.diagonal) s.diagonal = s.max_bb[2] - s.min_bb[2];
}
// unzoom (j3d/scene.cpp:91-111)
inline void unzoom(scene& s) {
  float d;
  j3dgh_unzoom(s.min_bb, s.max_bb, &d, s.pivot, s.coordinate_system.f, s.coordinate_system_inv.f);
}

// The grid-filling part of write_vox (j3d/vox.cpp:270-379) for an object that is already resident: dim and the
// x + (y + z * dim[1]) * dim[0] palette-index grid the reference hands to ogt_vox (file writing stays on the host).
inline void voxelize(context& ctx, const scene_object& obj, uint32_t max_dim, uint32_t dim[3], std::vector<uint8_t>& data) {
  ctx.check(j3dg_mesh_voxelize(obj.bvh, max_dim, dim, nullptr, 0), "j3dg_mesh_voxelize");
  data.assign((size_t)dim[0] * dim[1] * dim[2], 0);
  ctx.check(j3dg_mesh_voxelize(obj.bvh, max_dim, dim, data.data(), data.size()), "j3dg_mesh_voxelize");
}

// canvas (j3d/canvas.h:11-102): same public surface for the render path.
class canvas {
 public:
  struct canvas_settings {  // j3d/canvas.h:16-25
    bool one_bit = false, shadow = false, edges = true, wireframe = false, shading = true, textured = true, vertexcolors = true;
  };

  canvas(context& ctx, uint32_t w, uint32_t h) : _ctx(ctx) { resize(w, h); }

  void resize(uint32_t w, uint32_t h) {  // canvas.cpp:117-134
    _w = w;
    _h = h;
    _stride = (w + 3u) & ~3u;  // jtk::image<uint32_t> row padding (image.h:181-185)
    im.assign((size_t)_stride * h, 0);
    background.assign((size_t)_stride * h, 0);
    _canvas.assign((size_t)w * h, pixel{});
    for (auto& p : _canvas) { p.db_id = 0; p.object_id = (uint32_t)-1; }
    j3dgh_make_projection(w, h, &_near, projection_matrix.f, projection_matrix_inv.f);
  }
  void set_background_color(uint32_t clr_top = 0xff000000, uint32_t clr_bottom = 0xff404040) {  // canvas.cpp:136-139
    _bg_top = clr_top;
    _bg_bottom = clr_bottom;
    j3dgh_fill_background(_w, _h, _stride, clr_top, clr_bottom, background.data());
  }
  uint32_t width() const { return _w; }
  uint32_t height() const { return _h; }
  void update_settings(const canvas_settings& s) { _settings = s; }
  const float4x4& get_projection_matrix() const { return projection_matrix; }
  const float4x4& get_inverse_projection_matrix() const { return projection_matrix_inv; }
  const std::vector<pixel>& get_pixels() const { return _canvas; }
  const std::vector<uint32_t>& get_image() const { return im; }
  uint32_t image_stride() const { return _stride; }

  // canvas::update_canvas (canvas.cpp:677-874): inclusive rect, clamped by the library
  void update_canvas(std::vector<pixel>& out, int x0, int y0, int x1, int y1, const scene& s) {
    if (out.size() != (size_t)_w * _h) out.assign((size_t)_w * _h, pixel{});
    std::vector<j3dg_mesh*> meshes;
    for (const auto& o : s.objects)
      if (o.bvh) { j3dg_mesh_set_cs(o.bvh, o.cs.f); meshes.push_back(o.bvh); }
    j3dg_view v = make_view(s);
    _ctx.check(j3dg_cast(_ctx.get(), meshes.data(), (uint32_t)meshes.size(), &v, x0, y0, x1, y1, out.data(), _w), "j3dg_cast");
  }
  void update_canvas(int x0, int y0, int x1, int y1, const scene& s) { update_canvas(_canvas, x0, y0, x1, y1, s); }

  // canvas::render_scene (canvas.cpp:876-898)
  void render_scene(std::vector<pixel>& out, const scene* s) {
    if (s)
      update_canvas(out, 0, 0, (int)_w - 1, (int)_h - 1, *s);
    else {
      out.assign((size_t)_w * _h, pixel{});
      for (auto& p : out) { p.db_id = 0; p.object_id = (uint32_t)-1; }
    }
  }
  void render_scene(const scene* s) {
    im = background;
    render_scene(_canvas, s);
  }
  // canvas::canvas_to_image (canvas.cpp:582-670)
  // The shader reads the canvas size, the inverse projection, the near plane and the settings only — all members of
  // the canvas, as in the reference — so the signature is the reference's (canvas.h:61).
  void canvas_to_image(const std::vector<pixel>& cnv, const matcap& mc) {
    scene none;
    j3dg_view v = make_view(none);
    _ctx.check(j3dg_shade(_ctx.get(), cnv.data(), _w, &v, mc.im.data(), mc.w, mc.h, mc.w, mc.cavity_clr, nullptr, im.data(), _stride),
               "j3dg_shade");
  }
  // canvas::render_pointclouds_on_image (canvas.cpp:952-1030)
  void render_pointclouds_on_image(const scene* s, const std::vector<pixel>& pix) {
    if (!s || s->pointclouds.empty()) return;
    std::vector<j3dg_cloud*> clouds;
    for (const auto& o : s->pointclouds) clouds.push_back(o.cloud);
    j3dg_view v = make_view(*s);
    _ctx.check(j3dg_splat(_ctx.get(), clouds.data(), (uint32_t)clouds.size(), &v, pix.data(), _canvas.data(), _w, im.data(), _stride),
               "j3dg_splat");
  }

  // ---- picking, answered on the device from the canvas that is still resident after the last
  //      render_scene / update_canvas: canvas::get_pixel (canvas.cpp:141-153), the pivot pick of canvas::do_mouse
  //      (canvas.cpp:157-179) and, for `view`, get_world_position / get_index / get_id (view.cpp:439-492) ----
  j3dg_pick_result pick(const scene& s, int x, int y) {
    std::vector<j3dg_mesh*> meshes;
    std::vector<j3dg_cloud*> clouds;
    for (const auto& o : s.objects) if (o.bvh) meshes.push_back(o.bvh);
    for (const auto& o : s.pointclouds) if (o.cloud) clouds.push_back(o.cloud);
    j3dg_view v = make_view(s);
    const int32_t xy[2] = {x, y};
    j3dg_pick_result r;
    _ctx.check(j3dg_pick(_ctx.get(), meshes.data(), (uint32_t)meshes.size(), clouds.data(), (uint32_t)clouds.size(), &v, nullptr, 0, xy, 1, &r),
               "j3dg_pick");
    return r;
  }
  void get_pixel(pixel& p, const scene& s, float pos_x, float pos_y, float mouse_offset_x, float mouse_offset_y) {
    p = pick(s, (int)(uint32_t)(pos_x - mouse_offset_x), (int)(uint32_t)(pos_y - mouse_offset_y)).pixel;
  }
  // do_mouse, button-down part: a click on an object moves the scene pivot under the cursor
  void pick_pivot(scene& s, float mouse_x, float mouse_y, float mouse_offset_x, float mouse_offset_y) {
    const j3dg_pick_result r = pick(s, (int)(uint32_t)(mouse_x - mouse_offset_x), (int)(uint32_t)(mouse_y - mouse_offset_y));
    if (r.db_id) std::memcpy(s.pivot, r.pivot, 12);
  }
  bool get_world_position(const scene& s, int x, int y, float out[3]) {  // false: invalid_vertex (NaN)
    const j3dg_pick_result r = pick(s, x, y);
    std::memcpy(out, r.world_pos, 12);
    return r.db_id != 0 && r.closest_vertex != (uint32_t)-1;
  }
  uint32_t get_index(const scene& s, int x, int y) { return pick(s, x, y).closest_vertex; }
  uint32_t get_id(const scene& s, int x, int y) { return pick(s, x, y).db_id; }

  j3dg_view make_view(const scene& s) const {
    j3dg_view v;
    v.width = _w;
    v.height = _h;
    v.near_plane = _near;
    v.diagonal = s.diagonal;
    std::memcpy(v.projection, projection_matrix.f, 64);
    std::memcpy(v.projection_inv, projection_matrix_inv.f, 64);
    std::memcpy(v.cs, s.coordinate_system.f, 64);
    std::memcpy(v.cs_inv, s.coordinate_system_inv.f, 64);
    std::memcpy(v.pivot, s.pivot, 12);
    v.flags = (_settings.one_bit ? J3DG_ONE_BIT : 0) | (_settings.shadow ? J3DG_SHADOW : 0) | (_settings.edges ? J3DG_EDGES : 0) |
              (_settings.wireframe ? J3DG_WIREFRAME : 0) | (_settings.shading ? J3DG_SHADING : 0) |
              (_settings.textured ? J3DG_TEXTURED : 0) | (_settings.vertexcolors ? J3DG_VERTEXCOLORS : 0);
    return v;
  }

 private:
  context& _ctx;
  uint32_t _w = 0, _h = 0, _stride = 0;
  float _near = 0.1f;
  uint32_t _bg_top = 0xff000000, _bg_bottom = 0xff404040;
  std::vector<uint32_t> im, background;
  std::vector<pixel> _canvas;
  float4x4 projection_matrix, projection_matrix_inv;
  canvas_settings _settings;
};

// A sweep of frames of one scene (turntables, orbit renders; BASELINE configs[4]) with several frames in flight per GPU
// (include/j3dg.h, "Frames in flight"): `lanes` contexts of one device, each pipelining frames of its own
// (j3dg_frame_submit / j3dg_frame_wait), all rendering the SAME mesh / cloud handles.  Frames complete in submission
// order.  Not part of j3d (its view renders one frame per UI event); it is the call a batch renderer built on j3d makes.
class sweep {
 public:
  sweep(int device, const matcap& mc, int lanes = 3) {
    for (int i = 0; i < (lanes < 1 ? 1 : lanes); ++i) {
      j3dg_ctx* c = nullptr;
      if (j3dg_ctx_create(device, &c) != J3DG_OK) { release(); throw std::runtime_error(std::string("j3dg_ctx_create: ") + j3dg_last_error(nullptr)); }
      _ctx.push_back(c);
      _pending.push_back(0);
      if (j3dg_ctx_set_matcap(c, mc.im.data(), mc.w, mc.h, mc.w, mc.cavity_clr) != J3DG_OK || j3dg_ctx_set_dirty_rect(c, 1) != J3DG_OK) {
        const std::string why = j3dg_last_error(c);
        release();
        throw std::runtime_error("sweep: " + why);
      }
    }
  }
  ~sweep() { release(); }
  sweep(const sweep&) = delete;
  sweep& operator=(const sweep&) = delete;

  // Enqueues one frame.  rgba_out (width * height uint32) and pixels_out (nullable, width * height records) are HOST
  // buffers that stay untouched by the caller until the matching wait(); every lane wants its own persistent buffers
  // (the dirty-rectangle readback only rewrites what changed since the SAME buffer was last filled).
  void submit(const scene& s, const j3dg_view& v, uint32_t* rgba_out, pixel* pixels_out = nullptr, uint32_t bg_top = 0xff000000, uint32_t bg_bottom = 0xff404040) {
    const size_t lane = _next % _ctx.size();
    while (_pending[lane] >= 2) wait();  // a lane double-buffers: two frames of its own at most
    std::vector<j3dg_mesh*> meshes;
    std::vector<j3dg_cloud*> clouds;
    for (const auto& o : s.objects) if (o.bvh) meshes.push_back(o.bvh);
    for (const auto& o : s.pointclouds) if (o.cloud) clouds.push_back(o.cloud);
    if (j3dg_frame_submit(_ctx[lane], meshes.data(), (uint32_t)meshes.size(), clouds.data(), (uint32_t)clouds.size(), &v, nullptr, 0, 0, 0, 0, bg_top, bg_bottom,
                          pixels_out, rgba_out) != J3DG_OK)
      throw std::runtime_error(std::string("j3dg_frame_submit: ") + j3dg_last_error(_ctx[lane]));
    _order.push_back(lane);
    ++_pending[lane];
    ++_next;
  }
  // Completes the oldest frame in flight (its host buffers are final on return); false: nothing was in flight.
  bool wait() {
    if (_head == _order.size()) return false;
    const size_t lane = _order[_head++];
    if (j3dg_frame_wait(_ctx[lane]) != J3DG_OK) throw std::runtime_error(std::string("j3dg_frame_wait: ") + j3dg_last_error(_ctx[lane]));
    --_pending[lane];
    if (_head == _order.size()) { _order.clear(); _head = 0; }
    return true;
  }
  size_t in_flight() const { return _order.size() - _head; }
  size_t lanes() const { return _ctx.size(); }
  // the lane the NEXT submit uses (callers that keep one set of host buffers per lane index them with this)
  size_t next_lane() const { return _next % _ctx.size(); }

 private:
  void release() {
    for (j3dg_ctx* c : _ctx) j3dg_ctx_destroy(c);
    _ctx.clear();
  }
  std::vector<j3dg_ctx*> _ctx;
  std::vector<int> _pending;
  std::vector<size_t> _order;
  size_t _head = 0, _next = 0;
};

}  // namespace j3dg
