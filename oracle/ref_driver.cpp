// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" wrapper that drives the UNMODIFIED reference renderer (j3d's own
// canvas.cpp / scene.cpp / camera.cpp / matcap.cpp / pixel.cpp / db.cpp / trackball.c and
// the jtk headers) headlessly.  The reference sources are compiled where they lie under
// $J3D_REF (default /root/reference) by oracle/Makefile into oracle/_ref/libj3d_ref.so;
// no reference source is copied into this repository.  This file only *calls* the
// reference's public API in the order view::render_scene does (j3d/view.cpp:421-430).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting library.
#include "canvas.h"
#include "scene.h"
#include "db.h"
#include "mesh.h"
#include "pc.h"
#include "matcap.h"
#include "mouse.h"
#include "pixel.h"
#include "vox.h"
#include "ogt/ogt_vox.h"
#include <unistd.h>
#include <cstdio>
#include <cstdlib>

// Implementation sections of the stb-style jtk headers, in the order j3d/main.cpp:15-36
// instantiates them (they are not include-guarded, so order matters).
#define JTK_FILE_UTILS_IMPLEMENTATION
#include "jtk/file_utils.h"
#define JTK_GEOMETRY_IMPLEMENTATION
#include "jtk/geometry.h"
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#define JTK_QBVH_IMPLEMENTATION
#include "jtk/qbvh.h"
#define JTK_IMAGE_IMPLEMENTATION
#include "jtk/image.h"

#include <chrono>
#include <limits>
#include <cstring>
#include <thread>

#include "../include/j3dg.h"

using namespace jtk;

// mesh.cpp drags in every file-format reader; the renderer needs exactly one function
// from it.  Same semantics as j3d/mesh.cpp:39-58 (running min/max over the vertices).
void compute_bb(vec3<float>& mn, vec3<float>& mx, uint32_t nr_of_vertices, const vec3<float>* vertices)
  {
  if (nr_of_vertices == 0)
    return;
  mn = vertices[0];
  mx = vertices[0];
  for (uint32_t i = 1; i < nr_of_vertices; ++i)
    for (int j = 0; j < 3; ++j)
      {
      mn[j] = std::min<float>(mn[j], vertices[i][j]);
      mx[j] = std::max<float>(mx[j], vertices[i][j]);
      }
  }

namespace
  {
  struct ref_state
    {
    db database;
    scene scn;
    canvas cnv;
    matcap mc;
    canvas::canvas_settings settings;
    image<pixel> pixels; // view::_pixels
    double t_build = 0, t_cast = 0, t_shade = 0, t_splat = 0, t_copy = 0;
    ref_state(uint32_t w, uint32_t h) : cnv(w, h) {}
    };

  double now()
    {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

  void set_flags(canvas::canvas_settings& s, uint32_t flags)
    {
    s.one_bit = (flags & J3DG_ONE_BIT) != 0;
    s.shadow = (flags & J3DG_SHADOW) != 0;
    s.edges = (flags & J3DG_EDGES) != 0;
    s.wireframe = (flags & J3DG_WIREFRAME) != 0;
    s.shading = (flags & J3DG_SHADING) != 0;
    s.textured = (flags & J3DG_TEXTURED) != 0;
    s.vertexcolors = (flags & J3DG_VERTEXCOLORS) != 0;
    }
  }

extern "C" {

void* ref_create(uint32_t w, uint32_t h)
  {
  ref_state* st = new ref_state(w, h);
  st->scn.coordinate_system = get_identity();
  st->scn.coordinate_system_inv = get_identity();
  st->scn.pivot[0] = st->scn.pivot[1] = st->scn.pivot[2] = 0.f;
  st->scn.diagonal = 1.f;
  st->cnv.set_background_color(0xff000000, 0xff404040);
  set_flags(st->settings, J3DG_DEFAULT_FLAGS);
  st->cnv.update_settings(st->settings);
  make_matcap_red_wax(st->mc); // settings.cpp:19 default
  return st;
  }

void ref_destroy(void* p)
  {
  delete (ref_state*)p;
  }

int ref_hardware_concurrency()
  {
  return (int)std::thread::hardware_concurrency();
  }

// view::load_mesh_from_file minus the file read (j3d/view.cpp:206-240): fill a db mesh,
// add_object (normals + bbox + BVH), prepare_scene.  Returns the db id.
uint32_t ref_add_mesh(void* p, const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt,
  const float* vcolors, const float* uv, const uint32_t* tex, uint32_t tw, uint32_t th, const float* cs)
  {
  ref_state* st = (ref_state*)p;
  mesh* m;
  uint32_t id;
  st->database.create_mesh(m, id);
  m->vertices.resize(nv);
  std::memcpy((void*)m->vertices.data(), verts, sizeof(float) * 3 * nv);
  m->triangles.resize(nt);
  std::memcpy((void*)m->triangles.data(), tris, sizeof(uint32_t) * 3 * nt);
  if (vcolors)
    {
    m->vertex_colors.resize(nv);
    std::memcpy((void*)m->vertex_colors.data(), vcolors, sizeof(float) * 3 * nv);
    }
  if (uv && tex)
    {
    m->uv_coordinates.resize(nt);
    std::memcpy((void*)m->uv_coordinates.data(), uv, sizeof(float) * 6 * nt);
    m->texture = image<uint32_t>(tw, th);
    for (uint32_t y = 0; y < th; ++y)
      std::memcpy(m->texture.row(y), tex + (size_t)y * tw, sizeof(uint32_t) * tw);
    }
  m->cs = get_identity();
  if (cs)
    for (int i = 0; i < 16; ++i)
      m->cs[i] = cs[i];
  m->visible = true;
  double t0 = now();
  add_object(id, st->scn, st->database);
  st->t_build = now() - t0;
  // the reference passes &mesh.vertex_colors / &uv_coordinates even when they are empty
  // vectors, whose data() may be non-null garbage-free but size 0; canvas.cpp:738-740 tests
  // only the pointer.  Mirror view.cpp, which leaves them as they are: empty vector ->
  // data() == nullptr for a never-filled std::vector in libstdc++.
  prepare_scene(st->scn);
  return id;
  }

uint32_t ref_add_cloud(void* p, const float* pos, const float* nrm, const uint32_t* clr, uint32_t n, const float* cs)
  {
  ref_state* st = (ref_state*)p;
  pc* c;
  uint32_t id;
  st->database.create_pc(c, id);
  c->vertices.resize(n);
  std::memcpy((void*)c->vertices.data(), pos, sizeof(float) * 3 * n);
  if (nrm) // empty vector -> data()==nullptr -> unshaded points (canvas.cpp:994)
    {
    c->normals.resize(n);
    std::memcpy((void*)c->normals.data(), nrm, sizeof(float) * 3 * n);
    }
  if (clr) // empty vector -> white points (canvas.cpp:995, render.h:715)
    {
    c->vertex_colors.resize(n);
    std::memcpy((void*)c->vertex_colors.data(), clr, sizeof(uint32_t) * n);
    }
  c->cs = get_identity();
  if (cs)
    for (int i = 0; i < 16; ++i)
      c->cs[i] = cs[i];
  c->visible = true;
  add_object(id, st->scn, st->database);
  prepare_scene(st->scn);
  return id;
  }

void ref_unzoom(void* p)
  {
  ref_state* st = (ref_state*)p;
  prepare_scene(st->scn);
  unzoom(st->scn);
  }

void ref_get_view(void* p, j3dg_view* v)
  {
  ref_state* st = (ref_state*)p;
  v->width = st->cnv.width();
  v->height = st->cnv.height();
  v->near_plane = st->cnv.get_camera().nearClippingPlane;
  v->diagonal = st->scn.diagonal;
  for (int i = 0; i < 16; ++i)
    {
    v->projection[i] = st->cnv.get_projection_matrix()[i];
    v->projection_inv[i] = st->cnv.get_inverse_projection_matrix()[i];
    v->cs[i] = st->scn.coordinate_system[i];
    v->cs_inv[i] = st->scn.coordinate_system_inv[i];
    }
  for (int i = 0; i < 3; ++i)
    v->pivot[i] = st->scn.pivot[i];
  uint32_t f = 0;
  if (st->settings.one_bit) f |= J3DG_ONE_BIT;
  if (st->settings.shadow) f |= J3DG_SHADOW;
  if (st->settings.edges) f |= J3DG_EDGES;
  if (st->settings.wireframe) f |= J3DG_WIREFRAME;
  if (st->settings.shading) f |= J3DG_SHADING;
  if (st->settings.textured) f |= J3DG_TEXTURED;
  if (st->settings.vertexcolors) f |= J3DG_VERTEXCOLORS;
  v->flags = f;
  }

// Set the camera pose / pivot / settings from a view (projection and size are the canvas' own).
void ref_set_view(void* p, const j3dg_view* v)
  {
  ref_state* st = (ref_state*)p;
  for (int i = 0; i < 16; ++i)
    {
    st->scn.coordinate_system[i] = v->cs[i];
    st->scn.coordinate_system_inv[i] = v->cs_inv[i];
    }
  for (int i = 0; i < 3; ++i)
    st->scn.pivot[i] = v->pivot[i];
  set_flags(st->settings, v->flags);
  st->cnv.update_settings(st->settings);
  }

void ref_set_matcap(void* p, int type)
  {
  ref_state* st = (ref_state*)p;
  make_matcap(st->mc, int_to_matcap_type(type), "");
  }

void ref_get_matcap(void* p, uint32_t* out, uint32_t* w, uint32_t* h, uint32_t* cavity)
  {
  ref_state* st = (ref_state*)p;
  *w = st->mc.im.width();
  *h = st->mc.im.height();
  *cavity = st->mc.cavity_clr;
  if (out)
    for (uint32_t y = 0; y < *h; ++y)
      std::memcpy(out + (size_t)y * (*w), st->mc.im.row(y), sizeof(uint32_t) * (*w));
  }

// view::render_scene (j3d/view.cpp:421-430), stage by stage, each timed.
// stages bit0 = cast, bit1 = shade, bit2 = splat.
void ref_render(void* p, uint32_t stages)
  {
  ref_state* st = (ref_state*)p;
  st->cnv.update_settings(st->settings);
  if (stages & 1)
    {
    double t0 = now();
    st->cnv.render_scene(&st->scn);
    st->t_cast = now() - t0;
    t0 = now();
    st->pixels = st->cnv.get_pixels();
    st->t_copy = now() - t0;
    }
  if (stages & 2)
    {
    double t0 = now();
    st->cnv.canvas_to_image(st->pixels, st->mc);
    st->t_shade = now() - t0;
    }
  if (stages & 4)
    {
    double t0 = now();
    st->cnv.render_pointclouds_on_image(&st->scn, st->pixels);
    st->t_splat = now() - t0;
    }
  }

// which = 0: view::_pixels (snapshot before the splat), 1: canvas::_canvas (after the splat)
void ref_get_pixels(void* p, int which, j3dg_pixel* out)
  {
  ref_state* st = (ref_state*)p;
  const image<pixel>& im = which ? st->cnv.get_pixels() : st->pixels;
  static_assert(sizeof(pixel) == sizeof(j3dg_pixel), "pixel layout");
  for (uint32_t y = 0; y < im.height(); ++y)
    std::memcpy(out + (size_t)y * im.width(), im.row(y), sizeof(pixel) * im.width());
  }

void ref_get_image(void* p, uint32_t* out)
  {
  ref_state* st = (ref_state*)p;
  const image<uint32_t>& im = st->cnv.get_image();
  for (uint32_t y = 0; y < im.height(); ++y)
    std::memcpy(out + (size_t)y * im.width(), im.row(y), sizeof(uint32_t) * im.width());
  }

// out: build, cast, shade, splat, copy seconds
void ref_get_times(void* p, double* out)
  {
  ref_state* st = (ref_state*)p;
  out[0] = st->t_build;
  out[1] = st->t_cast;
  out[2] = st->t_shade;
  out[3] = st->t_splat;
  out[4] = st->t_copy;
  }

// Bare `new qbvh(triangles, vertices)` wall time (scene.cpp:23), seconds; also node count.
double ref_time_qbvh(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, uint32_t* nr_nodes)
  {
  (void)nv;
  std::vector<vec3<uint32_t>> triangles(nt);
  std::memcpy((void*)triangles.data(), tris, sizeof(uint32_t) * 3 * nt);
  double t0 = now();
  qbvh bvh(triangles, (const vec3<float>*)verts);
  double t = now() - t0;
  if (nr_nodes)
    *nr_nodes = (uint32_t)bvh.nodes.size();
  return t;
  }

// qbvh::find_closest_triangle on an arbitrary ray batch (the path qbvh_tests.cpp:708-749
// pins).  rays: n x {ox,oy,oz,dx,dy,dz,t_near,t_far}; hits: n x {u,v,distance,found}.
void ref_find_closest(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, uint32_t leaf_size,
  const float* rays, uint32_t n, float* hits, uint32_t* ids)
  {
  (void)nv;
  qbvh_voxel total_bb, centroid_bb;
  auto voxels = build_triangle_qbvh_voxels(total_bb, centroid_bb, (const vec3<float>*)verts, (const vec3<uint32_t>*)tris, nt);
  qbvh::properties props;
  props.leaf_size = leaf_size;
  qbvh bvh(voxels, nt, total_bb, centroid_bb, props);
  delete[] voxels;
  for (uint32_t i = 0; i < n; ++i)
    {
    ray r;
    r.orig = float4(rays[8 * i + 0], rays[8 * i + 1], rays[8 * i + 2], 1.f);
    r.dir = float4(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5], 0.f);
    r.t_near = rays[8 * i + 6];
    r.t_far = rays[8 * i + 7];
    uint32_t id = (uint32_t)-1;
    hit h = bvh.find_closest_triangle(id, r, (const vec3<uint32_t>*)tris, (const vec3<float>*)verts);
    hits[4 * i + 0] = h.found ? h.u : 0.f;
    hits[4 * i + 1] = h.found ? h.v : 0.f;
    hits[4 * i + 2] = h.distance;
    hits[4 * i + 3] = h.found ? 1.f : 0.f;
    ids[i] = h.found ? id : (uint32_t)-1;
    }
  }

// Picking through the reference's own code: canvas::get_pixel (canvas.cpp:148-153), get_closest_vertex
// (pixel.cpp:6-33), the pivot pick of canvas::do_mouse (canvas.cpp:157-179; a click without movement only
// sets scene::pivot, which is restored afterwards) and the expression of view::get_world_position
// (view.cpp:439-469; view.cpp itself needs SDL, so its four arithmetic lines are evaluated here with the
// reference's float4 operators).  Reads canvas::_canvas, i.e. the buffer after the splat.
void ref_pick(void* p, const int32_t* xy, uint32_t n, j3dg_pick_result* out)
  {
  ref_state* st = (ref_state*)p;
  static_assert(sizeof(j3dg_pick_result) == 64, "pick layout");
  const float qnan = std::numeric_limits<float>::quiet_NaN();
  for (uint32_t i = 0; i < n; ++i)
    {
    j3dg_pick_result r;
    std::memset(&r, 0, sizeof(r));
    r.world_pos[0] = r.world_pos[1] = r.world_pos[2] = qnan;
    r.pivot[0] = r.pivot[1] = r.pivot[2] = qnan;
    r.closest_vertex = (uint32_t)-1;
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (x >= 0 && y >= 0 && x < (int)st->cnv.width() && y < (int)st->cnv.height())
      {
      pixel px;
      st->cnv.get_pixel(px, (float)x, (float)y, 0.f, 0.f);
      std::memcpy(&r.pixel, &px, sizeof(px));
      r.db_id = px.db_id;
      if (px.db_id)
        {
        float saved[3] = { st->scn.pivot[0], st->scn.pivot[1], st->scn.pivot[2] };
        mouse_data md;
        std::memset(&md, 0, sizeof(md));
        md.mouse_x = md.prev_mouse_x = (float)x;
        md.mouse_y = md.prev_mouse_y = (float)y;
        md.left_button_down = true;
        bool refresh = false;
        st->cnv.do_mouse(refresh, md, st->scn, 0.f, 0.f);
        for (int k = 0; k < 3; ++k) { r.pivot[k] = st->scn.pivot[k]; st->scn.pivot[k] = saved[k]; }
        mesh* m = st->database.get_mesh(px.db_id);
        if (m)
          {
          const uint32_t v0 = m->triangles[px.object_id][0];
          const uint32_t v1 = m->triangles[px.object_id][1];
          const uint32_t v2 = m->triangles[px.object_id][2];
          const float4 V0(m->vertices[v0][0], m->vertices[v0][1], m->vertices[v0][2], 1.f);
          const float4 V1(m->vertices[v1][0], m->vertices[v1][1], m->vertices[v1][2], 1.f);
          const float4 V2(m->vertices[v2][0], m->vertices[v2][1], m->vertices[v2][2], 1.f);
          const float4 pos = V0 * (1.f - px.barycentric_u - px.barycentric_v) + px.barycentric_u * V1 + px.barycentric_v * V2;
          auto world_pos = matrix_vector_multiply(m->cs, pos);
          for (int k = 0; k < 3; ++k) r.world_pos[k] = world_pos[k];
          r.closest_vertex = get_closest_vertex(px, &m->vertices, &m->triangles);
          }
        else if (pc* ptcl = st->database.get_pc(px.db_id))
          {
          // view.cpp:462-467 dereferences the (null) mesh pointer here; the evident intent is the cloud's own cs
          const float4 pos(ptcl->vertices[px.object_id][0], ptcl->vertices[px.object_id][1], ptcl->vertices[px.object_id][2], 1.f);
          auto world_pos = matrix_vector_multiply(ptcl->cs, pos);
          for (int k = 0; k < 3; ++k) r.world_pos[k] = world_pos[k];
          r.closest_vertex = get_closest_vertex(px, &ptcl->vertices, (const std::vector<vec3<uint32_t>>*)nullptr);
          }
        }
      }
    out[i] = r;
    }
  }

// qbvh::find_all_triangles (qbvh.h:1854-2000) on a ray batch, on the tree vox.cpp:308 builds
// (`new qbvh(triangles, vertices)`).  CSR output: offsets[n + 1]; hits total x {u, v, distance, 0}; ids.
// Returns the total number of hits (nothing beyond `capacity` is written).
uint32_t ref_find_all(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* rays, uint32_t n,
  uint32_t* offsets, float* hits, uint32_t* ids, uint32_t capacity)
  {
  (void)nv;
  std::vector<vec3<uint32_t>> triangles(nt);
  std::memcpy((void*)triangles.data(), tris, sizeof(uint32_t) * 3 * nt);
  qbvh bvh(triangles, (const vec3<float>*)verts);
  uint32_t total = 0;
  for (uint32_t i = 0; i < n; ++i)
    {
    offsets[i] = total;
    ray r;
    r.orig = float4(rays[8 * i + 0], rays[8 * i + 1], rays[8 * i + 2], 1.f);
    r.dir = float4(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5], 0.f);
    r.t_near = rays[8 * i + 6];
    r.t_far = rays[8 * i + 7];
    std::vector<uint32_t> triangle_ids;
    std::vector<hit> all_hits = bvh.find_all_triangles(triangle_ids, r, triangles.data(), (const vec3<float>*)verts);
    for (size_t k = 0; k < all_hits.size(); ++k, ++total)
      {
      if (total < capacity)
        {
        hits[4 * (size_t)total + 0] = all_hits[k].u;
        hits[4 * (size_t)total + 1] = all_hits[k].v;
        hits[4 * (size_t)total + 2] = all_hits[k].distance;
        hits[4 * (size_t)total + 3] = 0.f;
        ids[total] = triangle_ids[k];
        }
      }
    }
  offsets[n] = total;
  return total;
  }

// The reference's voxel export end to end: write_vox (j3d/vox.cpp:270-440, compiled unmodified) into a
// temporary .vox file, read back with the reference's own ogt_vox reader.  Returns 0 on success; dims_out
// always receives the grid size; the grid (x + (y + z * dims[1]) * dims[0]) is copied if it fits capacity.
int ref_voxelize(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* vcolors,
  const float* uv, const uint32_t* tex, uint32_t tw, uint32_t th, uint32_t max_dim, uint32_t* dims_out, uint8_t* data, uint64_t capacity)
  {
  std::vector<vec3<float>> vertices(nv), clrs;
  std::memcpy((void*)vertices.data(), verts, sizeof(float) * 3 * nv);
  std::vector<vec3<uint32_t>> triangles(nt);
  std::memcpy((void*)triangles.data(), tris, sizeof(uint32_t) * 3 * nt);
  if (vcolors)
    {
    clrs.resize(nv);
    std::memcpy((void*)clrs.data(), vcolors, sizeof(float) * 3 * nv);
    }
  std::vector<vec3<vec2<float>>> uvs;
  image<uint32_t> texture;
  if (uv && tex && tw && th)
    {
    uvs.resize(nt);
    std::memcpy((void*)uvs.data(), uv, sizeof(float) * 6 * nt);
    texture = image<uint32_t>(tw, th);
    for (uint32_t y = 0; y < th; ++y)
      std::memcpy(texture.row(y), tex + (size_t)y * tw, sizeof(uint32_t) * tw);
    }
  char path[] = "/tmp/j3d_ref_vox_XXXXXX";
  int fd = mkstemp(path);
  if (fd < 0)
    return -1;
  close(fd);
  bool ok = write_vox(path, vertices, clrs, triangles, uvs, texture, max_dim);
  int rc = -2;
  if (ok)
    {
    FILE* fp = fopen(path, "rb");
    if (fp)
      {
      fseek(fp, 0, SEEK_END);
      uint32_t size = (uint32_t)ftell(fp);
      fseek(fp, 0, SEEK_SET);
      std::vector<uint8_t> buffer(size);
      size_t got = fread(buffer.data(), 1, size, fp);
      fclose(fp);
      const ogt_vox_scene* scene = got == size ? ogt_vox_read_scene(buffer.data(), size) : nullptr;
      if (scene && scene->num_models >= 1)
        {
        const ogt_vox_model* m = scene->models[0];
        dims_out[0] = m->size_x; dims_out[1] = m->size_y; dims_out[2] = m->size_z;
        const uint64_t nvox = (uint64_t)m->size_x * m->size_y * m->size_z;
        if (data && capacity >= nvox)
          std::memcpy(data, m->voxel_data, nvox);
        rc = 0;
        }
      if (scene)
        ogt_vox_destroy_scene(scene);
      }
    }
  remove(path);
  return rc;
  }

// Traversal statistics of the reference QBVH are not exposed by the reference; none here.

} // extern "C"
