// tests/cpp/shim_frame.cpp — drives the C++ mirror of j3d's scene / canvas (j3d_b200/host/j3dg_host.h) exactly the way
// view::load_mesh_from_file + view::render_scene do (j3d/view.cpp:206-240, 421-430), on synthetic inputs read from a
// raw file written by the Python test, and dumps the pixel buffer, the image, a few picks and the voxel grid so the
// test can compare them with the ctypes path and the oracle.  TEST CODE (built by tests/test_gpu_parity.py with g++).
//
//   shim_frame <in.bin> <out.bin>
//   in : u32 w, h, nv, nt, np, flags, max_dim | float verts[3nv] | u32 tris[3nt] | float ppos[3np] | float pnrm[3np] | u32 pclr[np]
//   out: pixel records w*h*32 | u32 image w*h (stride removed) | 3 picks (64 B each) | u32 dim[3] | u8 voxels |
//        u32 number of sweep frames (three in flight, j3dg::sweep) that differ from the synchronous frame
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "j3dg_host.h"

template <class T>
static void rd(FILE* f, std::vector<T>& v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t hdr[7];
  if (fread(hdr, 4, 7, f) != 7) return 2;
  const uint32_t w = hdr[0], h = hdr[1], nv = hdr[2], nt = hdr[3], np = hdr[4], flags = hdr[5], max_dim = hdr[6];
  try {
    j3dg::context ctx(0);
    j3dg::mesh m;
    j3dg::pc cloud;
    rd(f, m.vertices, 3 * (size_t)nv);
    rd(f, m.triangles, 3 * (size_t)nt);
    rd(f, cloud.vertices, 3 * (size_t)np);
    rd(f, cloud.normals, 3 * (size_t)np);
    rd(f, cloud.vertex_colors, (size_t)np);
    fclose(f);

    j3dg::scene s;
    j3dg::add_object(ctx, 0x20000000u, s, m);           // db.h:12-28 ids
    if (np) j3dg::add_object(ctx, 0x40000000u, s, cloud);
    j3dg::prepare_scene(s);
    j3dg::unzoom(s);

    j3dg::canvas cnv(ctx, w, h);
    cnv.set_background_color();
    j3dg::canvas::canvas_settings st;
    st.one_bit = flags & J3DG_ONE_BIT; st.shadow = flags & J3DG_SHADOW; st.edges = flags & J3DG_EDGES; st.wireframe = flags & J3DG_WIREFRAME;
    st.shading = flags & J3DG_SHADING; st.textured = flags & J3DG_TEXTURED; st.vertexcolors = flags & J3DG_VERTEXCOLORS;
    j3dg::matcap mc;
    j3dg::make_matcap_red_wax(mc);

    // view::render_scene (view.cpp:421-430)
    cnv.update_settings(st);
    cnv.render_scene(&s);
    std::vector<j3dg::pixel> pixels = cnv.get_pixels();  // _pixels = _canvas.get_pixels()
    cnv.canvas_to_image(pixels, mc);
    cnv.render_pointclouds_on_image(&s, pixels);

    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    fwrite(cnv.get_pixels().data(), sizeof(j3dg::pixel), (size_t)w * h, o);
    for (uint32_t y = 0; y < h; ++y) fwrite(cnv.get_image().data() + (size_t)y * cnv.image_stride(), 4, w, o);
    const int qx[3] = {(int)w / 2, (int)w / 3, -5}, qy[3] = {(int)h / 2, (int)h / 3, 2};
    for (int k = 0; k < 3; ++k) {
      j3dg_pick_result r = cnv.pick(s, qx[k], qy[k]);
      fwrite(&r, sizeof(r), 1, o);
    }
    uint32_t dim[3];
    std::vector<uint8_t> vox;
    j3dg::voxelize(ctx, s.objects.front(), max_dim, dim, vox);
    fwrite(dim, 4, 3, o);
    fwrite(vox.data(), 1, vox.size(), o);
    // ---- a sweep with three frames in flight (j3dg::sweep): 7 orbit poses, every frame equal to the synchronous one ----
    {
      const int nframes = 7;
      j3dg::sweep sw(0, mc, 3);
      std::vector<std::vector<uint32_t>> host(nframes, std::vector<uint32_t>((size_t)w * h));
      std::vector<j3dg_view> views;
      for (int k = 0; k < nframes; ++k) {
        j3dg_view v = cnv.make_view(s);
        j3dgh_orbit(s.coordinate_system_inv.f, s.pivot, 13.f * (float)k, v.cs, v.cs_inv);
        views.push_back(v);
        sw.submit(s, v, host[k].data());
      }
      while (sw.wait()) {}
      uint32_t bad = 0;
      std::vector<uint32_t> want((size_t)w * h);
      for (int k = 0; k < nframes; ++k) {
        std::vector<j3dg_mesh*> meshes;
        std::vector<j3dg_cloud*> clouds;
        for (const auto& ob : s.objects) meshes.push_back(ob.bvh);
        for (const auto& ob : s.pointclouds) clouds.push_back(ob.cloud);
        ctx.check(j3dg_render_frame(ctx.get(), meshes.data(), (uint32_t)meshes.size(), clouds.data(), (uint32_t)clouds.size(), &views[k], mc.im.data(), mc.w, mc.h, mc.w,
                                    mc.cavity_clr, 0xff000000u, 0xff404040u, nullptr, want.data()), "j3dg_render_frame");
        if (want != host[k]) ++bad;
      }
      fwrite(&bad, 4, 1, o);
    }
    fclose(o);
    j3dg::remove_object(0x20000000u, s);
    if (np) j3dg::remove_object(0x40000000u, s);
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_frame: %s\n", e.what());
    return 1;
  }
  return 0;
}
