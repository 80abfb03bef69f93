// ingest.cu — mesh / point-cloud ingest on the device: binary PLY decode.
// Replaces jtk::read_ply (jtk/ply.h:577-690, reached through j3d/io.cpp:733 from mesh.cpp:104-116 and pc.cpp:51-60) for
// the binary storage modes: the header (text) is parsed on the host, the file's bytes are uploaded ONCE, and two kernels
// turn the interleaved element records into the arrays the renderer keeps resident — vertices, normals, packed colours,
// triangle indices, per-corner texture coordinates — with the reference's conversions: every scalar goes through a
// double (rply's ply_get_argument_value) and is then cast to float / uint8 / uint32 (ply.h:517-573), colours start as
// 0xffffffff, a face contributes its first three indices, texcoord lists are cut or zero-padded to six floats.
// HBM-bound byte shuffling: a block stages its contiguous slice of records in shared memory with 16-byte loads, so
// the file is read once and coalesced whatever the record size is.  ASCII PLY is text parsing and stays on the host.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

enum PlyType { T_I8 = 0, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_NONE = -1 };
__host__ __device__ inline int type_size(int t) { return t == T_I8 || t == T_U8 ? 1 : (t == T_I16 || t == T_U16 ? 2 : (t == T_F64 ? 8 : 4)); }

int parse_type(const std::string& s) {
  static const char* names[8][2] = {{"char", "int8"}, {"uchar", "uint8"}, {"short", "int16"}, {"ushort", "uint16"}, {"int", "int32"}, {"uint", "uint32"}, {"float", "float32"}, {"double", "float64"}};
  for (int t = 0; t < 8; ++t)
    if (s == names[t][0] || s == names[t][1]) return t;
  return T_NONE;
}

struct Prop { std::string name; int type = T_NONE; int count_type = T_NONE; };  // count_type != T_NONE: a list property
struct Elem { std::string name; uint64_t count = 0; std::vector<Prop> props; };

struct Header {
  int format = 0;  // 1 binary_little_endian, 2 binary_big_endian
  uint64_t bytes = 0;
  std::vector<Elem> elems;
};

bool parse_header(const char* buf, size_t n, Header& h, std::string& err) {
  size_t pos = 0;
  auto next_line = [&](std::string& line) -> bool {
    if (pos >= n) return false;
    size_t e = pos;
    while (e < n && buf[e] != '\n') ++e;
    if (e >= n) return false;
    line.assign(buf + pos, e - pos);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    pos = e + 1;
    return true;
  };
  auto split = [](const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i;
      size_t j = i;
      while (j < s.size() && s[j] != ' ' && s[j] != '\t') ++j;
      if (j > i) out.emplace_back(s.substr(i, j - i));
      i = j;
    }
    return out;
  };
  std::string line;
  if (!next_line(line) || line != "ply") { err = "not a PLY file"; return false; }
  while (next_line(line)) {
    const std::vector<std::string> w = split(line);
    if (w.empty()) continue;
    if (w[0] == "end_header") { h.bytes = pos; return h.format != 0; }
    if (w[0] == "comment" || w[0] == "obj_info") continue;
    if (w[0] == "format" && w.size() >= 2) {
      if (w[1] == "binary_little_endian") h.format = 1;
      else if (w[1] == "binary_big_endian") h.format = 2;
      else { err = "ASCII PLY is parsed on the host, not by the device decoder"; return false; }
    } else if (w[0] == "element" && w.size() >= 3) {
      Elem e;
      e.name = w[1];
      e.count = strtoull(w[2].c_str(), nullptr, 10);
      h.elems.push_back(e);
    } else if (w[0] == "property" && !h.elems.empty()) {
      Prop p;
      if (w.size() >= 5 && w[1] == "list") { p.count_type = parse_type(w[2]); p.type = parse_type(w[3]); p.name = w[4]; if (p.count_type == T_NONE) { err = "bad list count type"; return false; } }
      else if (w.size() >= 3) { p.type = parse_type(w[1]); p.name = w[2]; }
      if (p.type == T_NONE) { err = "unknown property type in the PLY header"; return false; }
      h.elems.back().props.push_back(p);
    } else { err = "unexpected line in the PLY header: " + line; return false; }
  }
  err = "PLY header without end_header";
  return false;
}

// ---- device side ----------------------------------------------------------------------------------
struct Field { int off, type; };  // off < 0: absent

__device__ __forceinline__ double load_scalar(const uint8_t* p, int type, bool swap) {
  uint64_t raw = 0;
  const int sz = type_size(type);
  if (!swap) { for (int i = 0; i < sz; ++i) raw |= (uint64_t)p[i] << (8 * i); }
  else { for (int i = 0; i < sz; ++i) raw |= (uint64_t)p[i] << (8 * (sz - 1 - i)); }
  switch (type) {
    case T_I8: return (double)(int8_t)raw;
    case T_U8: return (double)(uint8_t)raw;
    case T_I16: return (double)(int16_t)raw;
    case T_U16: return (double)(uint16_t)raw;
    case T_I32: return (double)(int32_t)raw;
    case T_U32: return (double)(uint32_t)raw;
    case T_F32: return (double)__uint_as_float((uint32_t)raw);
    default: return __longlong_as_double((long long)raw);
  }
}

// Stages bytes [lo, hi) of `src` in shared memory (16-byte loads where aligned); returns the pointer to byte `lo`.
__device__ __forceinline__ const uint8_t* stage(const uint8_t* __restrict__ src, size_t lo, size_t hi, uint8_t* smem) {
  const size_t a0 = lo & ~(size_t)15;
  const size_t words = (hi - a0 + 15) >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + a0);  // src is 256-byte aligned (cudaMalloc)
  uint4* d4 = reinterpret_cast<uint4*>(smem);
  for (size_t i = threadIdx.x; i < words; i += blockDim.x) d4[i] = __ldg(s4 + i);
  __syncthreads();
  return smem + (lo - a0);
}

constexpr int INGEST_THREADS = 256;
constexpr int MAX_STAGED_STRIDE = 96;  // records up to this size are staged through shared memory (24.6 KB per block)

struct VertexLayout {
  Field pos[3], nrm[3], clr[4];
  uint32_t stride;
  int swap;
};

__global__ void __launch_bounds__(INGEST_THREADS) ply_vertex_kernel(const uint8_t* __restrict__ data, size_t base, uint32_t n, VertexLayout L,
                                                                      float* __restrict__ pos, float* __restrict__ nrm, uint32_t* __restrict__ clr) {
  __shared__ __align__(16) uint8_t smem[INGEST_THREADS * MAX_STAGED_STRIDE + 32];
  const uint32_t first = blockIdx.x * INGEST_THREADS;
  const uint32_t last = min(first + INGEST_THREADS, n);
  const uint32_t i = first + threadIdx.x;
  const uint8_t* rec;
  if (L.stride <= MAX_STAGED_STRIDE) rec = stage(data, base + (size_t)first * L.stride, base + (size_t)last * L.stride, smem) + (size_t)threadIdx.x * L.stride;
  else rec = data + base + (size_t)i * L.stride;
  if (i >= n) return;
  const bool swap = L.swap != 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) pos[3 * (size_t)i + j] = (float)load_scalar(rec + L.pos[j].off, L.pos[j].type, swap);  // ply.h:517-525
  if (nrm) {
#pragma unroll
    for (int j = 0; j < 3; ++j) nrm[3 * (size_t)i + j] = (float)load_scalar(rec + L.nrm[j].off, L.nrm[j].type, swap);
  }
  if (clr) {  // ply.h:527-535, 643-649: channels that the file does not carry stay 0xff
    uint32_t c = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (L.clr[j].off >= 0) {
        const uint32_t v = (uint32_t)(uint8_t)(int)load_scalar(rec + L.clr[j].off, L.clr[j].type, swap);
        c = (c & ~(0xffu << (8 * j))) | (v << (8 * j));
      }
    clr[i] = c;
  }
}

// A face record as the header declares it: up to 8 properties, each a scalar or a list.
struct FaceProp { int type, count_type; int role; };  // role 0 ignore, 1 vertex indices, 2 texcoord
struct FaceLayout {
  FaceProp props[8];
  int nprops;
  int assumed[8];      // fixed-stride path: the list lengths every face is assumed to have
  uint32_t stride;     // fixed-stride path: bytes per face under that assumption
  int swap;
};

// One face: walks the properties of its record (ply.h:537-573).  FIXED: list lengths are checked against the assumption.
template <bool FIXED>
__device__ __forceinline__ bool decode_face(const uint8_t* rec, const FaceLayout& L, uint32_t* tri, float* uv) {
  const bool swap = L.swap != 0;
  bool ok = true;
  for (int k = 0; k < L.nprops; ++k) {
    const FaceProp pr = L.props[k];
    if (pr.count_type == T_NONE) { rec += type_size(pr.type); continue; }
    const long len = (long)load_scalar(rec, pr.count_type, swap);
    rec += type_size(pr.count_type);
    if (FIXED && len != L.assumed[k]) ok = false;
    const long n = FIXED ? (long)L.assumed[k] : len;
    const int isz = type_size(pr.type);
    if (pr.role == 1) {
      if (n < 3) ok = false;
      for (int j = 0; j < 3 && j < n; ++j) tri[j] = (uint32_t)(long long)load_scalar(rec + j * isz, pr.type, swap);
    } else if (pr.role == 2 && uv) {
      for (int j = 0; j < 6; ++j) uv[j] = j < n ? (float)load_scalar(rec + j * isz, pr.type, swap) : 0.f;
    }
    rec += (size_t)n * isz;
  }
  return ok;
}

template <bool FIXED>
__global__ void __launch_bounds__(INGEST_THREADS) ply_face_kernel(const uint8_t* __restrict__ data, size_t base, const uint64_t* __restrict__ offsets, uint32_t n,
                                                                    FaceLayout L, uint32_t* __restrict__ tris, float* __restrict__ uv, uint32_t* __restrict__ mismatch) {
  __shared__ __align__(16) uint8_t smem[INGEST_THREADS * MAX_STAGED_STRIDE + 32];
  const uint32_t first = blockIdx.x * INGEST_THREADS;
  const uint32_t last = min(first + INGEST_THREADS, n);
  const uint32_t i = first + threadIdx.x;
  const uint8_t* rec;
  if (FIXED && L.stride <= MAX_STAGED_STRIDE) rec = stage(data, base + (size_t)first * L.stride, base + (size_t)last * L.stride, smem) + (size_t)threadIdx.x * L.stride;
  else if (FIXED) rec = data + base + (size_t)i * L.stride;
  else rec = data + base + (i < n ? offsets[i] : 0);
  if (i >= n) return;
  uint32_t t[3] = {0u, 0u, 0u};
  float w[6];
  const bool ok = decode_face<FIXED>(rec, L, t, uv ? w : nullptr);
  if (!ok) { *mismatch = 1u; return; }
  tris[3 * (size_t)i] = t[0]; tris[3 * (size_t)i + 1] = t[1]; tris[3 * (size_t)i + 2] = t[2];
  if (uv) {
#pragma unroll
    for (int j = 0; j < 6; ++j) uv[6 * (size_t)i + j] = w[j];
  }
}

__global__ void __launch_bounds__(256) colors_to_float_kernel(const uint32_t* __restrict__ clr, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = clr[i];  // convert_vertex_colors, j3d/mesh.cpp:235-248
  out[3 * (size_t)i] = fdiv((float)(c & 255u), 255.f);
  out[3 * (size_t)i + 1] = fdiv((float)((c >> 8) & 255u), 255.f);
  out[3 * (size_t)i + 2] = fdiv((float)((c >> 16) & 255u), 255.f);
}

double host_scalar(const uint8_t* p, int type, bool swap) {
  uint8_t b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int sz = type_size(type);
  for (int i = 0; i < sz; ++i) b[i] = swap ? p[sz - 1 - i] : p[i];
  switch (type) {
    case T_I8: return (double)(int8_t)b[0];
    case T_U8: return (double)b[0];
    case T_I16: { int16_t v; memcpy(&v, b, 2); return v; }
    case T_U16: { uint16_t v; memcpy(&v, b, 2); return v; }
    case T_I32: { int32_t v; memcpy(&v, b, 4); return v; }
    case T_U32: { uint32_t v; memcpy(&v, b, 4); return v; }
    case T_F32: { float v; memcpy(&v, b, 4); return v; }
    default: { double v; memcpy(&v, b, 8); return v; }
  }
}

}  // namespace

struct j3dg_ply {
  j3dg_ctx* ctx = nullptr;
  j3dg_ply_info info = {};
  float* d_vertices = nullptr;
  float* d_normals = nullptr;
  uint32_t* d_colors = nullptr;
  uint32_t* d_triangles = nullptr;
  float* d_uv = nullptr;
};

J3DG_API void j3dg_ply_destroy(j3dg_ply* p) {
  if (!p) return;
  if (p->ctx) { cudaSetDevice(p->ctx->device); cudaStreamSynchronize(p->ctx->stream); }
  cudaFree(p->d_vertices); cudaFree(p->d_normals); cudaFree(p->d_colors); cudaFree(p->d_triangles); cudaFree(p->d_uv);
  delete p;
}

J3DG_API int j3dg_ply_decode(j3dg_ctx* ctx, const void* file_bytes, size_t nbytes, j3dg_ply** out) {
  if (!ctx || !file_bytes || !out) { j3dg_set_error(ctx, "j3dg_ply_decode: bad argument"); return J3DG_EINVAL; }
  *out = nullptr;
  cudaSetDevice(ctx->device);
  const uint8_t* file = (const uint8_t*)file_bytes;
  Header h;
  std::string err;
  if (!parse_header((const char*)file, nbytes, h, err)) { j3dg_set_error(ctx, "j3dg_ply_decode: " + err); return J3DG_EINVAL; }
  const bool swap = h.format == 2;

  // ---- locate the vertex and face elements; every element before them must have a fixed record size ----
  const Elem* ve = nullptr; const Elem* fe = nullptr;
  size_t vbase = 0, fbase = 0, cursor = h.bytes;
  for (const Elem& e : h.elems) {
    bool has_list = false;
    size_t rec = 0;
    for (const Prop& p : e.props) { if (p.count_type != T_NONE) has_list = true; else rec += type_size(p.type); }
    if (e.name == "vertex") {
      if (has_list) { j3dg_set_error(ctx, "j3dg_ply_decode: list property in the vertex element"); return J3DG_EINVAL; }
      ve = &e; vbase = cursor; cursor += rec * e.count;
    } else if (e.name == "face") {
      fe = &e; fbase = cursor;
      break;  // whatever follows the faces is not read (the reference ignores it as well)
    } else {
      if (has_list) { j3dg_set_error(ctx, "j3dg_ply_decode: an element with list properties precedes the faces"); return J3DG_EINVAL; }
      cursor += rec * e.count;
    }
  }
  if (cursor > nbytes) { j3dg_set_error(ctx, "j3dg_ply_decode: the file is shorter than its header says"); return J3DG_EINVAL; }
  if ((ve && ve->count >= 0xFFFFFFFFull) || (fe && fe->count >= 0xFFFFFFFFull)) { j3dg_set_error(ctx, "j3dg_ply_decode: too many elements"); return J3DG_EINVAL; }

  j3dg_ply* ply = new j3dg_ply();
  ply->ctx = ctx;
  ply->info.format = (uint32_t)h.format;
  ply->info.header_bytes = h.bytes;
  ply->info.file_bytes = nbytes;
  auto fail = [&](int rc) { j3dg_ply_destroy(ply); return rc; };

  // ---- upload the element data once ----
  uint8_t* d_file = nullptr;
  cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
  cudaEventRecord(e0, ctx->stream);
  const size_t data_bytes = nbytes - h.bytes;
  if (data_bytes) {
    if (cudaMalloc((void**)&d_file, data_bytes + 64) != cudaSuccess) { cudaGetLastError(); j3dg_set_error(ctx, "out of device memory (PLY bytes)"); return fail(J3DG_ENOMEM); }
    if (j3dg_copy_to_device(ctx, d_file, file + h.bytes, data_bytes) != J3DG_OK) { cudaGetLastError(); cudaFree(d_file); return fail(J3DG_ECUDA); }
  }
  cudaEventRecord(e1, ctx->stream);
  vbase -= h.bytes; fbase -= h.bytes;
  auto release = [&]() { cudaStreamSynchronize(ctx->stream); cudaFree(d_file); };
  cudaEvent_t d0 = ctx->ev[4], d1 = ctx->ev[5];
  cudaEventRecord(d0, ctx->stream);

  // ---- vertices (ply.h:598-657) ----
  if (ve && ve->count) {
    VertexLayout L;
    for (auto& f : L.pos) f = Field{-1, 0};
    for (auto& f : L.nrm) f = Field{-1, 0};
    for (auto& f : L.clr) f = Field{-1, 0};
    L.swap = swap;
    int off = 0;
    auto find = [&](const char* name, Field& f) {
      int o = 0;
      for (const Prop& p : ve->props) { if (p.name == name) { f = Field{o, p.type}; return true; } o += type_size(p.type); }
      return false;
    };
    for (const Prop& p : ve->props) off += type_size(p.type);
    L.stride = (uint32_t)off;
    const char* pn[3] = {"x", "y", "z"}; const char* nn[3] = {"nx", "ny", "nz"};
    bool have_pos = true, have_nrm = true;
    for (int j = 0; j < 3; ++j) { have_pos = find(pn[j], L.pos[j]) && have_pos; have_nrm = find(nn[j], L.nrm[j]) && have_nrm; }
    // the colour channels are looked up as "red", then "r", then "diffuse_red" (ply.h:624-641)
    const char* cn[4][3] = {{"red", "r", "diffuse_red"}, {"green", "g", "diffuse_green"}, {"blue", "b", "diffuse_blue"}, {"alpha", "a", "diffuse_alpha"}};
    for (int j = 0; j < 4; ++j)
      for (int a = 0; a < 3; ++a)
        if (find(cn[j][a], L.clr[j])) break;
    const bool have_clr = L.clr[0].off >= 0;  // the colour array is sized by the red channel (ply.h:643-644)
    if (have_pos) {
      const uint32_t nv = (uint32_t)ve->count;
      if (vbase + (size_t)nv * L.stride > data_bytes) { release(); j3dg_set_error(ctx, "j3dg_ply_decode: truncated vertex data"); return fail(J3DG_EINVAL); }
      bool mem = cudaMalloc((void**)&ply->d_vertices, (size_t)nv * 12) == cudaSuccess;
      if (mem && have_nrm) mem = cudaMalloc((void**)&ply->d_normals, (size_t)nv * 12) == cudaSuccess;
      if (mem && have_clr) mem = cudaMalloc((void**)&ply->d_colors, (size_t)nv * 4) == cudaSuccess;
      if (!mem) { cudaGetLastError(); release(); j3dg_set_error(ctx, "out of device memory (PLY vertices)"); return fail(J3DG_ENOMEM); }
      ply_vertex_kernel<<<(nv + INGEST_THREADS - 1) / INGEST_THREADS, INGEST_THREADS, 0, ctx->stream>>>(d_file, vbase, nv, L, ply->d_vertices, ply->d_normals, ply->d_colors);
      ctx->launches++;
      ply->info.nr_of_vertices = nv;
      ply->info.has_normals = have_nrm; ply->info.has_colors = have_clr;
    }
  }
  // ---- faces (ply.h:659-679) ----
  if (fe && fe->count) {
    FaceLayout L;
    memset(&L, 0, sizeof(L));
    L.swap = swap;
    if (fe->props.size() > 8) { release(); j3dg_set_error(ctx, "j3dg_ply_decode: more than 8 face properties"); return fail(J3DG_EINVAL); }
    bool have_idx = false, have_uv = false;
    L.nprops = (int)fe->props.size();
    for (int k = 0; k < L.nprops; ++k) {
      const Prop& p = fe->props[k];
      L.props[k] = FaceProp{p.type, p.count_type, 0};
      if (p.count_type != T_NONE && !have_idx && (p.name == "vertex_indices" || p.name == "vertex_index")) { L.props[k].role = 1; have_idx = true; }
      else if (p.count_type != T_NONE && !have_uv && p.name == "texcoord") { L.props[k].role = 2; have_uv = true; }
    }
    if (have_idx) {
      const uint32_t nf = (uint32_t)fe->count;
      // the list lengths of the FIRST face are assumed for all of them (all-triangle meshes: one fixed stride)
      const uint8_t* rec = file + h.bytes + fbase;
      const uint8_t* end = file + nbytes;
      uint32_t stride = 0;
      bool first_ok = true;
      for (int k = 0; k < L.nprops && first_ok; ++k) {
        if (L.props[k].count_type == T_NONE) { stride += type_size(L.props[k].type); continue; }
        if (rec + stride + type_size(L.props[k].count_type) > end) { first_ok = false; break; }
        const long len = (long)host_scalar(rec + stride, L.props[k].count_type, swap);
        if (len < 0 || len > 255) { first_ok = false; break; }
        L.assumed[k] = (int)len;
        stride += type_size(L.props[k].count_type) + (uint32_t)len * type_size(L.props[k].type);
      }
      if (!first_ok) { release(); j3dg_set_error(ctx, "j3dg_ply_decode: truncated face data"); return fail(J3DG_EINVAL); }
      L.stride = stride;
      bool mem = cudaMalloc((void**)&ply->d_triangles, (size_t)nf * 12) == cudaSuccess;
      if (mem && have_uv) mem = cudaMalloc((void**)&ply->d_uv, (size_t)nf * 24) == cudaSuccess;
      uint32_t* d_flag = nullptr;
      if (mem) mem = cudaMalloc((void**)&d_flag, 4) == cudaSuccess;
      if (!mem) { cudaGetLastError(); release(); j3dg_set_error(ctx, "out of device memory (PLY faces)"); return fail(J3DG_ENOMEM); }
      cudaMemsetAsync(d_flag, 0, 4, ctx->stream);
      uint32_t mismatch = fbase + (size_t)nf * stride > data_bytes ? 1u : 0u;
      if (!mismatch) {
        ply_face_kernel<true><<<(nf + INGEST_THREADS - 1) / INGEST_THREADS, INGEST_THREADS, 0, ctx->stream>>>(d_file, fbase, nullptr, nf, L, ply->d_triangles, ply->d_uv, d_flag);
        ctx->launches++;
        cudaMemcpyAsync(&mismatch, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
      }
      if (mismatch) {
        // mixed polygons: the record offsets are a serial prefix over the list lengths — one pass on the host, then the
        // same decode with explicit offsets
        std::vector<uint64_t> offs(nf);
        const uint8_t* p = rec;
        bool ok = true;
        for (uint32_t i = 0; i < nf && ok; ++i) {
          offs[i] = (uint64_t)(p - rec);
          for (int k = 0; k < L.nprops; ++k) {
            if (L.props[k].count_type == T_NONE) { p += type_size(L.props[k].type); continue; }
            if (p + type_size(L.props[k].count_type) > end) { ok = false; break; }
            const long len = (long)host_scalar(p, L.props[k].count_type, swap);
            if (len < 0 || (L.props[k].role == 1 && len < 3)) { ok = false; break; }
            p += type_size(L.props[k].count_type) + (size_t)len * type_size(L.props[k].type);
          }
          if (p > end) ok = false;
        }
        uint64_t* d_offs = nullptr;
        if (ok && cudaMalloc((void**)&d_offs, (size_t)nf * 8) != cudaSuccess) { cudaGetLastError(); ok = false; }
        if (!ok) { cudaFree(d_flag); release(); j3dg_set_error(ctx, "j3dg_ply_decode: malformed or truncated face data (a face needs at least 3 vertices)"); return fail(J3DG_EINVAL); }
        cudaMemcpyAsync(d_offs, offs.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(d_flag, 0, 4, ctx->stream);
        ply_face_kernel<false><<<(nf + INGEST_THREADS - 1) / INGEST_THREADS, INGEST_THREADS, 0, ctx->stream>>>(d_file, fbase, d_offs, nf, L, ply->d_triangles, ply->d_uv, d_flag);
        ctx->launches++;
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_offs);
      }
      cudaFree(d_flag);
      ply->info.nr_of_faces = nf;
      ply->info.has_uv = have_uv;
    }
  }
  cudaEventRecord(d1, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_file);
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) { j3dg_cuda_fail(ctx, e, "PLY decode", __FILE__, __LINE__); return fail(J3DG_ECUDA); }
  cudaEventElapsedTime(&ply->info.upload_ms, e0, e1);
  cudaEventElapsedTime(&ply->info.decode_ms, d0, d1);
  *out = ply;
  return J3DG_OK;
}

J3DG_API int j3dg_ply_info_get(const j3dg_ply* ply, j3dg_ply_info* out) {
  if (!ply || !out) return J3DG_EINVAL;
  *out = ply->info;
  return J3DG_OK;
}

J3DG_API int j3dg_ply_arrays(const j3dg_ply* ply, const float** vertices, const float** normals, const uint32_t** colors, const uint32_t** triangles, const float** uv) {
  if (!ply) return J3DG_EINVAL;
  if (vertices) *vertices = ply->d_vertices;
  if (normals) *normals = ply->d_normals;
  if (colors) *colors = ply->d_colors;
  if (triangles) *triangles = ply->d_triangles;
  if (uv) *uv = ply->d_uv;
  return J3DG_OK;
}

J3DG_API int j3dg_ply_copy(const j3dg_ply* ply, int which, void* host_out, size_t capacity) {
  if (!ply || !host_out) return J3DG_EINVAL;
  const void* src = nullptr;
  size_t bytes = 0;
  switch (which) {
    case 0: src = ply->d_vertices; bytes = (size_t)ply->info.nr_of_vertices * 12; break;
    case 1: src = ply->d_normals; bytes = ply->d_normals ? (size_t)ply->info.nr_of_vertices * 12 : 0; break;
    case 2: src = ply->d_colors; bytes = ply->d_colors ? (size_t)ply->info.nr_of_vertices * 4 : 0; break;
    case 3: src = ply->d_triangles; bytes = (size_t)ply->info.nr_of_faces * 12; break;
    case 4: src = ply->d_uv; bytes = ply->d_uv ? (size_t)ply->info.nr_of_faces * 24 : 0; break;
    default: return J3DG_EINVAL;
  }
  if (capacity < bytes) { j3dg_set_error(ply->ctx, "j3dg_ply_copy: buffer too small"); return J3DG_EINVAL; }
  if (bytes) CU_CHECK(ply->ctx, cudaMemcpy(host_out, src, bytes, cudaMemcpyDeviceToHost));
  return J3DG_OK;
}

// read_from_file(mesh&) for a PLY (j3d/mesh.cpp:104-116, 179-183): packed colours become float triples, a mesh with
// texture coordinates but no texture gets the 512 x 512 checkerboard of make_dummy_texture (mesh.cpp:21-36).
J3DG_API int j3dg_mesh_create_from_ply(j3dg_ctx* ctx, const j3dg_ply* ply, const float* cs, uint32_t db_id, j3dg_mesh** out) {
  if (!ctx || !ply || !out) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  float* d_vc = nullptr;
  const uint32_t nv = ply->info.nr_of_vertices;
  if (ply->d_colors && nv) {
    CU_CHECK(ctx, cudaMalloc((void**)&d_vc, (size_t)nv * 12));
    colors_to_float_kernel<<<(nv + 255) / 256, 256, 0, ctx->stream>>>(ply->d_colors, nv, d_vc);
    ctx->launches++;
  }
  std::vector<uint32_t> tex;
  if (ply->d_uv) {
    tex.resize(512 * 512);
    for (int y = 0; y < 512; ++y)
      for (int x = 0; x < 512; ++x) {
        const bool ye = ((y / 32) & 1) == 1, xe = ((x / 32) & 1) == 1;
        tex[(size_t)y * 512 + x] = ((xe && ye) || (!xe && !ye)) ? 0xff000000u : 0xffffffffu;
      }
  }
  const int rc = j3dg_mesh_create(ctx, ply->d_vertices, nv, ply->d_triangles, ply->info.nr_of_faces, d_vc, ply->d_uv, tex.empty() ? nullptr : tex.data(),
                                  tex.empty() ? 0 : 512, tex.empty() ? 0 : 512, 512, cs, db_id, out);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_vc);
  return rc;
}

// read_from_file(pc&) for a PLY (j3d/pc.cpp:51-60): positions, normals and packed colours as the file has them.
J3DG_API int j3dg_cloud_create_from_ply(j3dg_ctx* ctx, const j3dg_ply* ply, const float* cs, uint32_t db_id, j3dg_cloud** out) {
  if (!ctx || !ply || !out) return J3DG_EINVAL;
  if (!ply->info.nr_of_vertices) { j3dg_set_error(ctx, "j3dg_cloud_create_from_ply: no vertices (pc.cpp:57-58)"); return J3DG_EINVAL; }
  return j3dg_cloud_create(ctx, ply->d_vertices, ply->d_normals, ply->d_colors, ply->info.nr_of_vertices, cs, db_id, out);
}
