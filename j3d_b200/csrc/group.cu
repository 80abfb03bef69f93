// group.cu — multi-GPU plumbing behind the C ABI (SURVEY §8b "Context / multi-GPU", §8e): one process per GPU,
// an NCCL communicator per group for what IS a collective (the one-time BVH broadcast, small control words) and
// NVLink peer memory for the per-frame result exchange (every rank's shade kernel stores its RGBA straight into a
// buffer the destination rank owns; csrc/peer.cu holds the stream-ordered flag kernels).
//
// j3d itself has no distributed code, so there is no reference interface to mirror: a C++ host that
// wants N GPUs creates one context + one group per process, broadcasts the mesh it loaded on one rank, and renders
// its share of the frames (orbit sweep) or of the screen bands (one huge frame) into a j3dg_frames exchange.
//
// NCCL is loaded with dlopen when the first group is created, so single-GPU users of libj3dg.so do not need it and
// a Python process that already carries torch's NCCL shares that copy (same soname).
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only

#include <cstdio>
#include <cstring>
#include <mutex>

namespace {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string error;
};

NcclApi g_nccl;
std::mutex g_nccl_mutex;

bool nccl_load() {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.handle) return true;
  const char* names[] = {getenv("J3DG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { g_nccl.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return false; }
  auto sym = [&](const char* name) -> void* {
    void* p = dlsym(h, name);
    if (!p) g_nccl.error = std::string("NCCL symbol missing: ") + name;
    return p;
  };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Broadcast || !g_nccl.AllReduce || !g_nccl.GetErrorString) {
    dlclose(h);
    return false;
  }
  g_nccl.handle = h;
  return true;
}

}  // namespace

struct j3dg_group {
  j3dg_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  uint32_t* d_word = nullptr;   // 64 words of device staging for control messages
};

struct j3dg_frames {
  j3dg_group* g = nullptr;
  uint32_t w = 0, h = 0;
  int dst = 0;
  bool shared = false;
  char* base = nullptr;         // the exchange buffer: [2 slots][world or 1][h*w] RGBA, then the flag words
  size_t frame_bytes = 0, flags_off = 0, nbytes = 0;
  uint32_t k = 0;               // next frame number
  uint32_t nslots = 2;          // frames that may be in flight (frame k lives in slot k mod nslots)
  j3dg_ctx* lane[J3DG_FRAMES_MAX_SLOTS] = {};  // the context (stream) that renders the frames of each slot (j3dg_frames_set_lane)
};

namespace {

int nccl_fail(j3dg_ctx* ctx, ncclResult_t r, const char* what) {
  char buf[384];
  snprintf(buf, sizeof(buf), "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
  j3dg_set_error(ctx, buf);
  return J3DG_ECUDA;
}
#define NCCL_CHECK(ctx, call)                                          \
  do {                                                                 \
    ncclResult_t r__ = (call);                                         \
    if (r__ != ncclSuccess) return nccl_fail((ctx), r__, #call);       \
  } while (0)

// Broadcast `bytes` of HOST memory from root (through the device staging words; bytes <= 256).
int bcast_small(j3dg_group* g, void* host, size_t bytes, int root) {
  j3dg_ctx* ctx = g->ctx;
  if (bytes > 256) return J3DG_EINVAL;
  if (g->rank == root) CU_CHECK(ctx, cudaMemcpyAsync(g->d_word, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(ctx, g_nccl.Broadcast(g->d_word, g->d_word, bytes, ncclUint8, root, g->comm, ctx->stream));
  CU_CHECK(ctx, cudaMemcpyAsync(host, g->d_word, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return J3DG_OK;
}

// min over the ranks of an int (collective agreement on "did every rank succeed")
int all_min(j3dg_group* g, int* value) {
  j3dg_ctx* ctx = g->ctx;
  CU_CHECK(ctx, cudaMemcpyAsync(g->d_word, value, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(ctx, g_nccl.AllReduce(g->d_word, g->d_word, 1, ncclInt32, ncclMin, g->comm, ctx->stream));
  CU_CHECK(ctx, cudaMemcpyAsync(value, g->d_word, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return J3DG_OK;
}

int bcast_device(j3dg_group* g, void* dev, size_t bytes, int root) {
  const size_t chunk = (size_t)1 << 30;
  for (size_t off = 0; off < bytes; off += chunk)
    NCCL_CHECK(g->ctx, g_nccl.Broadcast((char*)dev + off, (char*)dev + off, std::min(chunk, bytes - off), ncclUint8, root, g->comm, g->ctx->stream));
  return J3DG_OK;
}

// What a rank needs to know to receive a mesh it has not seen.
struct MeshMeta {
  uint32_t nv, nt, nr_nodes, db_id;
  uint32_t has_vertices, has_indices, has_vcolors, has_uv;
  uint32_t tex_w, tex_h;
  float cs[16];
  j3dg_mesh_info info;
};
static_assert(sizeof(MeshMeta) <= 256, "the mesh meta data travels through the 256-byte staging buffer");

}  // namespace

J3DG_API int j3dg_group_unique_id(unsigned char* id_out) {
  if (!id_out) return J3DG_EINVAL;
  static_assert(sizeof(ncclUniqueId) == J3DG_GROUP_ID_BYTES, "NCCL unique id size");
  if (!nccl_load()) { j3dg_set_error(nullptr, g_nccl.error); return J3DG_ENODEV; }
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(nullptr, r, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return J3DG_OK;
}

J3DG_API int j3dg_group_create(j3dg_ctx* ctx, int rank, int world, const unsigned char* id, j3dg_group** out) {
  if (!ctx || !out || !id || world < 1 || rank < 0 || rank >= world) { j3dg_set_error(ctx, "j3dg_group_create: bad argument"); return J3DG_EINVAL; }
  *out = nullptr;
  if (!nccl_load()) { j3dg_set_error(ctx, g_nccl.error); return J3DG_ENODEV; }
  cudaSetDevice(ctx->device);
  j3dg_group* g = new j3dg_group();
  g->ctx = ctx; g->rank = rank; g->world = world;
  ncclUniqueId nid;
  memcpy(&nid, id, sizeof(nid));
  ncclResult_t r = g_nccl.CommInitRank(&g->comm, world, nid, rank);
  if (r != ncclSuccess) { delete g; return nccl_fail(ctx, r, "ncclCommInitRank"); }
  if (cudaMalloc((void**)&g->d_word, 256) != cudaSuccess) {
    cudaGetLastError();
    g_nccl.CommDestroy(g->comm);
    delete g;
    j3dg_set_error(ctx, "out of device memory (group staging)");
    return J3DG_ENOMEM;
  }
  *out = g;
  return J3DG_OK;
}

J3DG_API void j3dg_group_destroy(j3dg_group* g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  if (g->comm) g_nccl.CommDestroy(g->comm);
  cudaFree(g->d_word);
  delete g;
}

J3DG_API int j3dg_group_rank(const j3dg_group* g) { return g ? g->rank : -1; }
J3DG_API int j3dg_group_world(const j3dg_group* g) { return g ? g->world : 0; }

J3DG_API int j3dg_group_barrier(j3dg_group* g) {
  if (!g) return J3DG_EINVAL;
  cudaSetDevice(g->ctx->device);
  int one = 1;
  int rc = all_min(g, &one);
  if (rc != J3DG_OK) return rc;
  return j3dg_check_sticky(g->ctx);
}

J3DG_API int j3dg_group_max_float(j3dg_group* g, float* values, uint32_t n) {
  if (!g || !values || !n || n > 64) return J3DG_EINVAL;
  j3dg_ctx* ctx = g->ctx;
  cudaSetDevice(ctx->device);
  CU_CHECK(ctx, cudaMemcpyAsync(g->d_word, values, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(ctx, g_nccl.AllReduce(g->d_word, g->d_word, n, ncclFloat32, ncclMax, g->comm, ctx->stream));
  CU_CHECK(ctx, cudaMemcpyAsync(values, g->d_word, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return J3DG_OK;
}

J3DG_API int j3dg_group_allreduce_max_u64(j3dg_group* g, unsigned long long* dev_words, size_t n) {
  if (!g || (n && !dev_words)) return J3DG_EINVAL;
  if (!n) return J3DG_OK;
  cudaSetDevice(g->ctx->device);
  NCCL_CHECK(g->ctx, g_nccl.AllReduce(dev_words, dev_words, n, ncclUint64, ncclMax, g->comm, g->ctx->stream));
  return J3DG_OK;
}

// The mesh a rank loaded and built (its BVH, triangle records AND the indexed geometry / colours / texture the
// resolve and pick kernels read) replicated into every other rank's HBM: ~2.3 GB for the 28 M-triangle mesh,
// one ncclBroadcast per array over NVLink.
J3DG_API int j3dg_group_broadcast_mesh(j3dg_group* g, int root, j3dg_mesh** mesh_inout) {
  if (!g || !mesh_inout || root < 0 || root >= g->world) { j3dg_set_error(g ? g->ctx : nullptr, "j3dg_group_broadcast_mesh: bad argument"); return J3DG_EINVAL; }
  j3dg_ctx* ctx = g->ctx;
  cudaSetDevice(ctx->device);
  MeshMeta meta;
  memset(&meta, 0, sizeof(meta));
  j3dg_mesh* m = *mesh_inout;
  int ok = 1;
  if (g->rank == root) {
    if (!m || !m->d_nodes) ok = 0;
    else {
      meta.nv = m->nv; meta.nt = m->nt; meta.nr_nodes = m->nr_nodes; meta.db_id = m->db_id;
      meta.has_vertices = m->d_vertices != nullptr; meta.has_indices = m->d_indices != nullptr;
      meta.has_vcolors = m->d_vcolors != nullptr; meta.has_uv = m->d_uv != nullptr;
      meta.tex_w = m->d_texture ? m->tex_w : 0; meta.tex_h = m->d_texture ? m->tex_h : 0;
      memcpy(meta.cs, m->cs, sizeof(meta.cs));
      meta.info = m->info;
    }
  } else if (m) ok = 0;  // the receiving ranks get a new mesh
  int rc = all_min(g, &ok);
  if (rc != J3DG_OK) return rc;
  if (!ok) { j3dg_set_error(ctx, "j3dg_group_broadcast_mesh: the root needs a built mesh, the other ranks a NULL mesh"); return J3DG_EINVAL; }
  if ((rc = bcast_small(g, &meta, sizeof(meta), root)) != J3DG_OK) return rc;
  if (g->rank != root) {
    if ((rc = j3dg_mesh_create_empty(ctx, meta.nv, meta.nt, meta.nr_nodes, meta.cs, meta.db_id, &m)) == J3DG_OK) {
      auto alloc = [&](void** p, size_t bytes) { return bytes == 0 || cudaMalloc(p, bytes) == cudaSuccess; };
      bool mem = true;
      if (meta.has_vertices) mem = mem && alloc((void**)&m->d_vertices, (size_t)meta.nv * 12);
      if (meta.has_indices) mem = mem && alloc((void**)&m->d_indices, (size_t)meta.nt * 12);
      if (meta.has_vcolors) mem = mem && alloc((void**)&m->d_vcolors, (size_t)meta.nv * 12);
      if (meta.has_uv) mem = mem && alloc((void**)&m->d_uv, (size_t)meta.nt * 24);
      if (meta.tex_w && meta.tex_h) mem = mem && alloc((void**)&m->d_texture, (size_t)meta.tex_w * meta.tex_h * 4);
      if (!mem) { cudaGetLastError(); j3dg_set_error(ctx, "out of device memory (received mesh)"); rc = J3DG_ENOMEM; }
      m->tex_w = meta.tex_w; m->tex_h = meta.tex_h;
      m->info = meta.info;
    }
    ok = rc == J3DG_OK;
  }
  int rc2 = all_min(g, &ok);
  if (rc2 != J3DG_OK || !ok) {
    if (g->rank != root && m) j3dg_mesh_destroy(m);
    return rc != J3DG_OK ? rc : (rc2 != J3DG_OK ? rc2 : J3DG_ENOMEM);
  }
  if ((rc = bcast_device(g, m->d_nodes, (size_t)meta.nr_nodes * sizeof(WideNode), root)) != J3DG_OK) return rc;
  if ((rc = bcast_device(g, m->d_tris, (size_t)meta.nt * sizeof(TriRec), root)) != J3DG_OK) return rc;
  if (meta.has_vertices && (rc = bcast_device(g, m->d_vertices, (size_t)meta.nv * 12, root)) != J3DG_OK) return rc;
  if (meta.has_indices && (rc = bcast_device(g, m->d_indices, (size_t)meta.nt * 12, root)) != J3DG_OK) return rc;
  if (meta.has_vcolors && (rc = bcast_device(g, m->d_vcolors, (size_t)meta.nv * 12, root)) != J3DG_OK) return rc;
  if (meta.has_uv && (rc = bcast_device(g, m->d_uv, (size_t)meta.nt * 24, root)) != J3DG_OK) return rc;
  if (meta.tex_w && meta.tex_h && (rc = bcast_device(g, m->d_texture, (size_t)meta.tex_w * meta.tex_h * 4, root)) != J3DG_OK) return rc;
  *mesh_inout = m;
  return J3DG_OK;  // enqueued on the context stream; synchronise (or render, same stream) to use it
}

// ---- frames handed to one rank through NVLink peer memory (protocol: csrc/peer.cu header) ----------------------
J3DG_API int j3dg_frames_create(j3dg_group* g, uint32_t width, uint32_t height, int dst, int shared_frame, j3dg_frames** out) {
  return j3dg_frames_create_n(g, width, height, dst, shared_frame, 2, out);
}

J3DG_API int j3dg_frames_create_n(j3dg_group* g, uint32_t width, uint32_t height, int dst, int shared_frame, uint32_t nslots, j3dg_frames** out) {
  if (!g || !out || !width || !height || dst < 0 || dst >= g->world || nslots < 2 || nslots > J3DG_FRAMES_MAX_SLOTS) { j3dg_set_error(g ? g->ctx : nullptr, "j3dg_frames_create: bad argument"); return J3DG_EINVAL; }
  *out = nullptr;
  j3dg_ctx* ctx = g->ctx;
  cudaSetDevice(ctx->device);
  j3dg_frames* f = new j3dg_frames();
  f->g = g; f->w = width; f->h = height; f->dst = dst; f->shared = shared_frame != 0;
  f->nslots = nslots;
  for (auto& l : f->lane) l = ctx;
  f->frame_bytes = (size_t)width * height * 4;
  f->flags_off = (nslots * (size_t)(f->shared ? 1 : g->world) * f->frame_bytes + 255) & ~(size_t)255;
  f->nbytes = f->flags_off + 4 * (size_t)nslots * ((size_t)g->world + 1) + 256;
  unsigned char handle[J3DG_IPC_HANDLE_BYTES];
  memset(handle, 0, sizeof(handle));
  int ok = 1;
  if (g->rank == dst) {
    void* p = nullptr;
    if (j3dg_peer_alloc(ctx, f->nbytes, &p, handle) != J3DG_OK) ok = 0;
    f->base = (char*)p;
  }
  // every rank takes part in every collective of the set-up, whatever happened before: nobody is left behind
  int rc = all_min(g, &ok);
  if (rc == J3DG_OK && ok) rc = bcast_small(g, handle, sizeof(handle), dst);
  if (rc == J3DG_OK && ok && g->rank != dst) {
    void* p = nullptr;
    if (j3dg_peer_open(ctx, handle, &p) != J3DG_OK) ok = 0;
    f->base = (char*)p;
  }
  int rc2 = rc == J3DG_OK ? all_min(g, &ok) : rc;
  if (rc2 != J3DG_OK || !ok) {
    const std::string why = ctx->error;
    if (f->base) { if (g->rank == dst) j3dg_peer_free(ctx, f->base); else j3dg_peer_close(ctx, f->base); }
    delete f;
    j3dg_set_error(ctx, "peer-memory frame exchange unavailable on at least one rank" + (why.empty() ? std::string() : ": " + why));
    return rc2 != J3DG_OK ? rc2 : J3DG_ECUDA;
  }
  *out = f;
  return J3DG_OK;
}

J3DG_API void j3dg_frames_destroy(j3dg_frames* f) {
  if (!f) return;
  j3dg_group* g = f->g;
  j3dg_ctx* ctx = g->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (j3dg_ctx* l : f->lane) if (l != ctx) cudaStreamSynchronize(l->stream);
  int one = 1;
  all_min(g, &one);  // everybody is done with the buffer
  if (g->rank != f->dst) j3dg_peer_close(ctx, f->base);
  all_min(g, &one);  // the others have unmapped it
  if (g->rank == f->dst) j3dg_peer_free(ctx, f->base);
  delete f;
}

// Flag words behind the frames: arrived[slot][rank] (nslots * world words), then released[slot] (nslots words).  One set
// PER SLOT: the slots may be driven from different streams (several frames in flight), and the frames of one slot
// follow each other on one stream, so every word only ever grows.
static uint32_t* frames_arrived(j3dg_frames* f, uint32_t slot, int r) { return reinterpret_cast<uint32_t*>(f->base + f->flags_off) + slot * (uint32_t)f->g->world + r; }
static uint32_t* frames_released(j3dg_frames* f, uint32_t slot) { return reinterpret_cast<uint32_t*>(f->base + f->flags_off) + f->nslots * (uint32_t)f->g->world + slot; }

J3DG_API int j3dg_frames_set_lane(j3dg_frames* f, int slot, j3dg_ctx* ctx) {
  if (!f || slot < 0 || slot >= (int)f->nslots || !ctx) return J3DG_EINVAL;
  if (ctx->device != f->g->ctx->device) { j3dg_set_error(f->g->ctx, "j3dg_frames_set_lane: the lane context must live on the group's device"); return J3DG_EINVAL; }
  f->lane[slot] = ctx;
  return J3DG_OK;
}

J3DG_API int j3dg_frames_begin(j3dg_frames* f, uint32_t* k_out) {
  if (!f || !k_out) return J3DG_EINVAL;
  const uint32_t k = f->k;
  *k_out = k;
  const uint32_t s = k % f->nslots;
  if (k >= f->nslots)  // frame k - nslots lived in this slot: dst must have released it (released[slot] = number of the last consumed frame + 1)
    return j3dg_stream_wait_geq(f->lane[s], frames_released(f, s), 1, k - f->nslots + 1);
  return j3dg_check_sticky(f->lane[s]);
}

J3DG_API int j3dg_frames_target(j3dg_frames* f, uint32_t k, uint32_t** rgba_out) {
  if (!f || !rgba_out) return J3DG_EINVAL;
  const size_t per = f->shared ? 1 : (size_t)f->g->world;
  *rgba_out = reinterpret_cast<uint32_t*>(f->base + ((k % f->nslots) * per + (f->shared ? 0 : (size_t)f->g->rank)) * f->frame_bytes);
  return J3DG_OK;
}

J3DG_API int j3dg_frames_arrive(j3dg_frames* f, uint32_t k) {
  if (!f || k != f->k) { j3dg_set_error(f ? f->g->ctx : nullptr, "j3dg_frames_arrive: frames arrive in order (k must be the value j3dg_frames_begin returned)"); return J3DG_EINVAL; }
  const uint32_t s = k % f->nslots;
  j3dg_ctx* ctx = f->lane[s];
  int rc = j3dg_stream_signal(ctx, frames_arrived(f, s, f->g->rank), k + 1);
  if (rc != J3DG_OK) return rc;
  if (f->g->rank == f->dst && (rc = j3dg_stream_wait_geq(ctx, frames_arrived(f, s, 0), (uint32_t)f->g->world, k + 1)) != J3DG_OK) return rc;
  f->k = k + 1;
  return J3DG_OK;
}

J3DG_API int j3dg_frames_release(j3dg_frames* f, uint32_t k) {
  if (!f) return J3DG_EINVAL;
  if (f->g->rank != f->dst) return J3DG_OK;
  return j3dg_stream_signal(f->lane[k % f->nslots], frames_released(f, k % f->nslots), k + 1);
}

J3DG_API int j3dg_frames_view(j3dg_frames* f, uint32_t k, const uint32_t** frames_out) {
  if (!f || !frames_out) return J3DG_EINVAL;
  if (f->g->rank != f->dst) { j3dg_set_error(f->g->ctx, "j3dg_frames_view: only the destination rank holds the frames"); return J3DG_EINVAL; }
  const size_t per = f->shared ? 1 : (size_t)f->g->world;
  *frames_out = reinterpret_cast<const uint32_t*>(f->base + (k % f->nslots) * per * f->frame_bytes);
  return J3DG_OK;
}
