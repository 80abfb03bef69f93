// oracle/ref_ingest.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Second translation unit of oracle/_ref/libj3d_ref.so: the reference's mesh / point-cloud INGEST path (SURVEY §8f
// rank 4).  It instantiates the unmodified jtk PLY reader (jtk/ply.h, the code behind j3d/io.cpp:733 read_ply) and links
// the unmodified j3d/pc.cpp for estimate_normals (pc.cpp:256-334: k nearest neighbours through jtk::point_tree, a plane
// fit per point, orientation propagation over the neighbour graph).  pc.cpp also holds the file-format dispatch, whose
// readers live in io.cpp together with codecs this path does not need (trico, stb, tinygltf); the few io.h functions it
// references are defined here: read_ply as the one-line forward io.cpp:733-736 is, the others as "not available".
#include "pc.h"
#include "io.h"

#define JTK_PLY_IMPLEMENTATION
#include "jtk/ply.h"

#include <cstring>

using namespace jtk;

bool read_ply(const char* filename, std::vector<vec3<float>>& vertices, std::vector<vec3<float>>& normals, std::vector<uint32_t>& clrs, std::vector<vec3<uint32_t>>& triangles, std::vector<vec3<vec2<float>>>& uv)
  {
  return jtk::read_ply(filename, vertices, normals, clrs, triangles, uv);  // io.cpp:733-736
  }
bool write_ply(const char* filename, const std::vector<vec3<float>>& vertices, const std::vector<vec3<float>>& normals, const std::vector<uint32_t>& clrs, const std::vector<vec3<uint32_t>>& triangles, const std::vector<vec3<vec2<float>>>& uv)
  {
  return jtk::write_ply(filename, vertices, normals, clrs, triangles, uv);  // io.cpp:738-741
  }
bool read_trc(const char*, std::vector<vec3<float>>&, std::vector<vec3<float>>&, std::vector<uint32_t>&, std::vector<vec3<uint32_t>>&, std::vector<vec3<vec2<float>>>&) { return false; }
bool write_trc(const char*, const std::vector<vec3<float>>&, const std::vector<vec3<float>>&, const std::vector<uint32_t>&, const std::vector<vec3<uint32_t>>&, const std::vector<vec3<vec2<float>>>&) { return false; }
bool read_obj(const char*, std::vector<vec3<float>>&, std::vector<vec3<float>>&, std::vector<uint32_t>&, std::vector<vec3<uint32_t>>&, std::vector<vec3<vec2<float>>>&, image<uint32_t>&) { return false; }
bool write_obj(const char*, const std::vector<vec3<float>>&, const std::vector<vec3<float>>&, const std::vector<uint32_t>&, const std::vector<vec3<uint32_t>>&, const std::vector<vec3<vec2<float>>>&, const image<uint32_t>&) { return false; }
bool read_pts(const char*, std::vector<vec3<float>>&, std::vector<int>&, std::vector<uint32_t>&) { return false; }
bool write_pts(const char*, const std::vector<vec3<float>>&, const std::vector<int>&, const std::vector<uint32_t>&) { return false; }
bool read_xyz(const char*, std::vector<vec3<float>>&) { return false; }
bool write_xyz(const char*, const std::vector<vec3<float>>&) { return false; }

extern "C" {

// jtk::read_ply_from_memory (jtk/ply.h:693): sizes first (outputs NULL), then the arrays.
// counts: {nv, nn, nc, nt, nuv}.  Returns 1 on success.
int ref_read_ply(const char* buffer, uint64_t size, uint64_t* counts, float* vertices, float* normals, uint32_t* colors, uint32_t* triangles, float* uv)
  {
  std::vector<vec3<float>> v, n;
  std::vector<uint32_t> c;
  std::vector<vec3<uint32_t>> t;
  std::vector<vec3<vec2<float>>> u;
  if (!jtk::read_ply_from_memory(buffer, size, v, n, c, t, u))
    return 0;
  counts[0] = v.size(); counts[1] = n.size(); counts[2] = c.size(); counts[3] = t.size(); counts[4] = u.size();
  if (vertices && !v.empty()) std::memcpy(vertices, v.data(), v.size() * 12);
  if (normals && !n.empty()) std::memcpy(normals, n.data(), n.size() * 12);
  if (colors && !c.empty()) std::memcpy(colors, c.data(), c.size() * 4);
  if (triangles && !t.empty()) std::memcpy(triangles, t.data(), t.size() * 12);
  if (uv && !u.empty()) std::memcpy(uv, u.data(), u.size() * 24);
  return 1;
  }

// estimate_normals (j3d/pc.cpp:256-334) on n points; out: n x 3 floats.
void ref_estimate_normals(const float* positions, uint32_t n, uint32_t k, float* out)
  {
  pc p;
  p.vertices.resize(n);
  std::memcpy((void*)p.vertices.data(), positions, (size_t)n * 12);
  std::vector<vec3<float>> nrm = estimate_normals(p, k);
  std::memcpy(out, nrm.data(), (size_t)n * 12);
  }

}
