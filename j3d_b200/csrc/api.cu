// api.cu — the C ABI (include/j3dg.h): contexts, mesh / cloud handles, host<->device staging
// and the per-frame orchestration of view::render_scene (j3d/view.cpp:421-430).
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

// Host -> device copy of a big PAGEABLE array (the std::vectors of j3d's mesh / pc structs): cudaMemcpy from pageable
// memory stages through one driver thread at ~10 GB/s.  Here eight host threads copy 16 MB chunks into a ring of four
// pinned buffers while the DMA engine drains the previous chunks, which keeps PCIe busy (config B, 504 MB of vertices and
// indices: 51 -> 17-19 ms).  Pinned and device sources, and small arrays, take the direct path.
constexpr size_t STAGE_CHUNK = 16u << 20;
constexpr int STAGE_THREADS = 8;

static void j3dg_stage_ring_start(j3dg_ctx* ctx) {
  ctx->stage_init = new std::thread([ctx]() {
    cudaSetDevice(ctx->device);
    for (int i = 0; i < 4; ++i)
      if (cudaHostAlloc(&ctx->h_stage[i], STAGE_CHUNK, cudaHostAllocDefault) != cudaSuccess) { ctx->h_stage[i] = nullptr; break; }
    __sync_synchronize();
    ctx->stage_ready = 1;
    j3dg_preload_build_kernels();  // lazy module loading: have the kernels of the first mesh_create / frame resident by then
    j3dg_preload_cast_kernels();
  });
}
static void j3dg_stage_ring_join(j3dg_ctx* ctx) {
  if (!ctx->stage_init) return;
  std::thread* t = (std::thread*)ctx->stage_init;
  t->join();
  delete t;
  ctx->stage_init = nullptr;
  cudaGetLastError();
}

int j3dg_copy_to_device(j3dg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  cudaPointerAttributes a;
  bool pageable = true;
  if (cudaPointerGetAttributes(&a, src) == cudaSuccess) pageable = a.type == cudaMemoryTypeUnregistered;
  else cudaGetLastError();
  if (!pageable || bytes < 2 * STAGE_CHUNK) {
    CU_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return J3DG_OK;
  }
  while (!ctx->stage_ready) std::this_thread::yield();  // page-locking 64 MB takes tens of milliseconds: j3dg_ctx_create started it in the background
  __sync_synchronize();
  for (int i = 0; i < 4; ++i) {
    if (!ctx->h_stage[i]) {  // no pinned memory to spare: the plain copy still works
      CU_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
      return J3DG_OK;
    }
    if (!ctx->stage_ev[i]) {
      CU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
      ctx->stage_busy[i] = false;
    }
  }
  size_t off = 0;
  for (int c = 0; off < bytes; ++c, off += STAGE_CHUNK) {
    const int slot = c & 3;
    const size_t len = std::min(STAGE_CHUNK, bytes - off);
    if (ctx->stage_busy[slot]) CU_CHECK(ctx, cudaEventSynchronize(ctx->stage_ev[slot]));  // the DMA of the chunk that used this buffer is done
    char* stage = (char*)ctx->h_stage[slot];
    const char* from = (const char*)src + off;
    const size_t part = (len / STAGE_THREADS + 63) & ~(size_t)63;
    std::thread workers[STAGE_THREADS - 1];
    for (int t = 1; t < STAGE_THREADS; ++t) {
      const size_t a0 = std::min(len, part * t), a1 = std::min(len, part * (t + 1));
      workers[t - 1] = std::thread([=]() { if (a1 > a0) memcpy(stage + a0, from + a0, a1 - a0); });
    }
    memcpy(stage, from, std::min(len, part));
    for (auto& w : workers) w.join();
    CU_CHECK(ctx, cudaMemcpyAsync((char*)dst + off, stage, len, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaEventRecord(ctx->stage_ev[slot], ctx->stream));
    ctx->stage_busy[slot] = true;
  }
  return J3DG_OK;
}


namespace {
std::mutex g_err_mutex;
std::string g_last_error = "";
}

void j3dg_set_error(j3dg_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->error = msg;
  std::lock_guard<std::mutex> lock(g_err_mutex);
  g_last_error = msg;
}

int j3dg_cuda_fail(j3dg_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  j3dg_set_error(ctx, buf);
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? J3DG_ENOMEM : J3DG_ECUDA;
}

bool j3dg_is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int j3dg_reserve(j3dg_ctx* ctx, void** ptr, size_t* cap, size_t bytes) {
  if (*cap >= bytes && *ptr) return J3DG_OK;
  if (*ptr) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->last_canvas == *ptr) ctx->last_canvas = nullptr;  // j3dg_pick must not read a freed canvas
    cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
  }
  const size_t want = std::max<size_t>(bytes, 256);
  cudaError_t e = cudaMalloc(ptr, want);
  if (e != cudaSuccess) {
    *ptr = nullptr;
    char buf[256];
    snprintf(buf, sizeof(buf), "out of device memory reserving %zu bytes: %s", want, cudaGetErrorString(e));
    j3dg_set_error(ctx, buf);
    cudaGetLastError();
    return J3DG_ENOMEM;
  }
  *cap = want;
  return J3DG_OK;
}

void j3dg_invert_orthonormal_host(const float* m, float* out) {  // jtk invert_orthonormal, qbvh.h:4460-4467
  float r[16];
  const float c0[4] = {m[0], m[4], m[8], 0.f}, c1[4] = {m[1], m[5], m[9], 0.f}, c2[4] = {m[2], m[6], m[10], 0.f};
  for (int k = 0; k < 4; ++k) {
    r[k] = c0[k]; r[4 + k] = c1[k]; r[8 + k] = c2[k];
    volatile float a = c0[k] * m[12], b = c1[k] * m[13], c = c2[k] * m[14];
    volatile float s = a + b;
    s = s + c;
    r[12 + k] = -s;
  }
  r[15] = 1.f;
  memcpy(out, r, sizeof(r));
}

namespace {

const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};

template <class T>
int upload(j3dg_ctx* ctx, T** dst, const T* src, size_t count) {
  *dst = nullptr;
  if (!src || !count) return J3DG_OK;
  cudaError_t e = cudaMalloc((void**)dst, count * sizeof(T));
  if (e != cudaSuccess) {
    *dst = nullptr;
    j3dg_set_error(ctx, std::string("out of device memory: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return J3DG_ENOMEM;
  }
  return j3dg_copy_to_device(ctx, *dst, src, count * sizeof(T));
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  return ms;
}

}  // namespace

int j3dg_check_sticky(j3dg_ctx* ctx) {
  if (!ctx->h_status) return J3DG_OK;
  if (ctx->h_status[0]) { j3dg_set_error(ctx, "a stream-ordered flag wait timed out (peer exchange): a peer did not arrive / release in time"); return J3DG_ETIMEOUT; }
  if (ctx->h_status[1]) { j3dg_set_error(ctx, "traversal stack overflow in an earlier frame (BVH deeper than the kernel's stack)"); return J3DG_ECUDA; }
  return J3DG_OK;
}

namespace {

int check_overflow(j3dg_ctx* ctx) {
  unsigned long long st[4];
  CU_CHECK(ctx, cudaMemcpyAsync(st, ctx->d_stats, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if ((uint32_t)st[2]) {
    j3dg_set_error(ctx, "traversal stack overflow (BVH deeper than the kernel's stack)");
    return J3DG_ECUDA;
  }
  return J3DG_OK;
}

}  // namespace

// ---- context ---------------------------------------------------------------------------------
J3DG_API int j3dg_ctx_create(int device, j3dg_ctx** out) {
  if (!out) return J3DG_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    j3dg_set_error(nullptr, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    return J3DG_ENODEV;
  }
  if (device < 0 || device >= count) { j3dg_set_error(nullptr, "invalid device ordinal"); return J3DG_EINVAL; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); j3dg_set_error(nullptr, "cudaGetDeviceProperties failed"); return J3DG_ENODEV; }
  if (prop.major != 10) {
    char buf[512];
    snprintf(buf, sizeof(buf), "device %d (%s) is sm_%d%d; libj3dg is built for sm_100a only and has no fallback", device, prop.name, prop.major, prop.minor);
    j3dg_set_error(nullptr, buf);
    return J3DG_ENODEV;
  }
  j3dg_ctx* ctx = new j3dg_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    j3dg_set_error(nullptr, "cannot create a CUDA stream");
    cudaGetLastError();
    delete ctx;
    return J3DG_ECUDA;
  }
  ctx->stream = ctx->own_stream;
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  if (cudaMalloc((void**)&ctx->d_stats, 24 * sizeof(unsigned long long)) != cudaSuccess) {
    j3dg_set_error(nullptr, "cannot allocate device memory");
    delete ctx;
    return J3DG_ENOMEM;
  }
  cudaMemset(ctx->d_stats, 0, 24 * sizeof(unsigned long long));
  {
    uint32_t* hs = nullptr;
    if (cudaHostAlloc((void**)&hs, 4 * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->d_status, hs, 0) != cudaSuccess) {
      cudaGetLastError();
      j3dg_set_error(nullptr, "cannot allocate the mapped status words");
      j3dg_ctx_destroy(ctx);
      return J3DG_ENOMEM;
    }
    memset(hs, 0, 4 * sizeof(uint32_t));
    ctx->h_status = hs;
  }
  if (const char* e = getenv("J3DG_LANE_BUDGET")) ctx->lane_budget = (uint32_t)std::max(1, atoi(e));  // developer tuning knobs
  if (const char* e = getenv("J3DG_SHADOW_BUDGET")) ctx->shadow_budget = (uint32_t)std::max(1, atoi(e));
  if (const char* e = getenv("J3DG_CAST_ALGO")) ctx->cast_algo = strcmp(e, "group") == 0 ? 1 : 0;
  if (const char* e = getenv("J3DG_TOP_MIN")) ctx->top_min = (uint32_t)std::max(2, atoi(e));
  j3dg_stage_ring_start(ctx);
  *out = ctx;
  return J3DG_OK;
}

J3DG_API void j3dg_ctx_destroy(j3dg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_pixels); cudaFree(ctx->d_pixels_in); cudaFree(ctx->d_rgba); cudaFree(ctx->d_bg); cudaFree(ctx->d_packed);
  cudaFree(ctx->d_matcap); cudaFree(ctx->d_meshes); cudaFree(ctx->d_top); cudaFree(ctx->d_stats); cudaFree(ctx->d_misc); cudaFree(ctx->d_shadow); cudaFree(ctx->d_hard); cudaFree(ctx->d_spill);
  for (auto& sl : ctx->slot) {
    cudaFree(sl.d_px); cudaFree(sl.d_rgba);
    if (sl.kernels_done) cudaEventDestroy(sl.kernels_done);
    if (sl.copy_done) cudaEventDestroy(sl.copy_done);
  }
  j3dg_stage_ring_join(ctx);
  for (int i = 0; i < 4; ++i) {
    if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->h_overflow) cudaFreeHost(ctx->h_overflow);
  if (ctx->h_status) cudaFreeHost((void*)ctx->h_status);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto& r : ctx->ring) { for (auto e : r.a) cudaEventDestroy(e); for (auto e : r.b) cudaEventDestroy(e); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

J3DG_API const char* j3dg_last_error(const j3dg_ctx* ctx) {
  if (ctx) return ctx->error.c_str();
  std::lock_guard<std::mutex> lock(g_err_mutex);
  static thread_local std::string copy;
  copy = g_last_error;
  return copy.c_str();
}

J3DG_API int j3dg_ctx_set_stream(j3dg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return J3DG_EINVAL;
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return J3DG_OK;
}

J3DG_API int j3dg_ctx_synchronize(j3dg_ctx* ctx) {
  if (!ctx) return J3DG_EINVAL;
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return j3dg_check_sticky(ctx);
}

J3DG_API int j3dg_ctx_status(j3dg_ctx* ctx, uint32_t* flags_out, int reset) {
  if (!ctx || !flags_out) return J3DG_EINVAL;
  *flags_out = (ctx->h_status[0] ? J3DG_STATUS_WAIT_TIMEOUT : 0u) | (ctx->h_status[1] ? J3DG_STATUS_STACK_OVERFLOW : 0u);
  if (reset && *flags_out) {
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // no kernel that could still raise a word is in flight
    ctx->h_status[0] = 0u; ctx->h_status[1] = 0u;
  }
  return J3DG_OK;
}

J3DG_API int j3dg_ctx_set_profiling(j3dg_ctx* ctx, int enabled) {
  if (!ctx) return J3DG_EINVAL;
  ctx->profiling = enabled != 0;
  return J3DG_OK;
}

J3DG_API int j3dg_ctx_set_tuning(j3dg_ctx* ctx, uint32_t lane_budget, int cast_algo) {
  if (!ctx || cast_algo < 0 || cast_algo > 1) return J3DG_EINVAL;
  ctx->lane_budget = lane_budget ? lane_budget : 28u;
  ctx->cast_algo = cast_algo;
  return J3DG_OK;
}

J3DG_API int j3dg_ctx_set_screen_shard(j3dg_ctx* ctx, uint32_t rank, uint32_t world) {
  if (!ctx || world == 0 || rank >= world) { j3dg_set_error(ctx, "j3dg_ctx_set_screen_shard: need rank < world"); return J3DG_EINVAL; }
  ctx->shard_rank = rank;
  ctx->shard_world = world;
  return J3DG_OK;
}

J3DG_API int j3dg_ctx_timings(j3dg_ctx* ctx, j3dg_timings* out, int reset) {
  if (!ctx || !out) return J3DG_EINVAL;
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  float sums[3] = {0.f, 0.f, 0.f};
  for (int s = 0; s < 3; ++s)
    for (uint32_t i = 0; i < ctx->ring[s].used; ++i) sums[s] += elapsed(ctx->ring[s].a[i], ctx->ring[s].b[i]);
  unsigned long long st[8];
  CU_CHECK(ctx, cudaMemcpy(st, ctx->d_stats, sizeof(st), cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(*out));
  out->cast_ms = sums[0]; out->shade_ms = sums[1]; out->splat_ms = sums[2];
  out->cast_count = ctx->ring[0].used; out->shade_count = ctx->ring[1].used; out->splat_count = ctx->ring[2].used;
  out->rays = ctx->rays_primary + st[4];  // one shadow ray per hit pixel when shadows were on
  out->kernel_launches = ctx->launches;
  if (reset) {
    ctx->launches = 0;
    ctx->rays_primary = 0;
    for (auto& r : ctx->ring) r.used = 0;
    CU_CHECK(ctx, cudaMemset(ctx->d_stats + 4, 0, sizeof(unsigned long long)));
  }
  return J3DG_OK;
}

// Event pairs are taken from a ring that grows on demand; when a stage was launched more
// often than the ring holds since the last reset, the oldest slot is reused (the sum then
// covers the most recent launches only).
static int ring_slot(j3dg_ctx* ctx, int stage, bool begin) {
  auto& r = ctx->ring[stage];
  const uint32_t cap = 8192;
  uint32_t i = begin ? r.used : r.used - 1;
  if (begin && r.used >= cap) { r.used = cap - 1; i = r.used; }
  if (i >= r.a.size()) {
    cudaEvent_t ea, eb;
    if (cudaEventCreate(&ea) != cudaSuccess || cudaEventCreate(&eb) != cudaSuccess) return -1;
    r.a.push_back(ea); r.b.push_back(eb);
  }
  return (int)i;
}

int j3dg_stage_begin(j3dg_ctx* ctx, int stage) {
  if (!ctx->profiling) return J3DG_OK;
  const int i = ring_slot(ctx, stage, true);
  if (i < 0) { j3dg_set_error(ctx, "cudaEventCreate failed"); return J3DG_ECUDA; }
  CU_CHECK(ctx, cudaEventRecord(ctx->ring[stage].a[i], ctx->stream));
  ctx->ring[stage].used++;
  return J3DG_OK;
}

int j3dg_stage_end(j3dg_ctx* ctx, int stage) {
  if (!ctx->profiling) return J3DG_OK;
  const int i = ring_slot(ctx, stage, false);
  if (i < 0) return J3DG_OK;
  CU_CHECK(ctx, cudaEventRecord(ctx->ring[stage].b[i], ctx->stream));
  return J3DG_OK;
}

// ---- meshes ------------------------------------------------------------------------------------
J3DG_API int j3dg_mesh_create(j3dg_ctx* ctx, const float* vertices, uint32_t nv, const uint32_t* triangles, uint32_t nt,
                              const float* vertex_colors, const float* uv, const uint32_t* texture, uint32_t tex_w, uint32_t tex_h,
                              uint32_t tex_stride, const float* cs, uint32_t db_id, j3dg_mesh** out) {
  if (!ctx || !out || (nv && !vertices) || (nt && !triangles)) { j3dg_set_error(ctx, "j3dg_mesh_create: bad argument"); return J3DG_EINVAL; }
  if (nt >= (1u << 29)) { j3dg_set_error(ctx, "j3dg_mesh_create: more than 2^29 triangles"); return J3DG_EINVAL; }
  *out = nullptr;
  cudaSetDevice(ctx->device);
  j3dg_mesh* m = new j3dg_mesh();
  m->ctx = ctx; m->nv = nv; m->nt = nt; m->db_id = db_id;
  memcpy(m->cs, cs ? cs : kIdentity, sizeof(m->cs));
  j3dg_invert_orthonormal_host(m->cs, m->cs_inv);
  int rc;
  cudaEventRecord(ctx->ev[6], ctx->stream);
  if ((rc = upload(ctx, &m->d_vertices, vertices, (size_t)nv * 3)) != J3DG_OK) { j3dg_mesh_destroy(m); return rc; }
  if ((rc = upload(ctx, &m->d_indices, triangles, (size_t)nt * 3)) != J3DG_OK) { j3dg_mesh_destroy(m); return rc; }
  if ((rc = upload(ctx, &m->d_vcolors, vertex_colors, (size_t)nv * 3)) != J3DG_OK) { j3dg_mesh_destroy(m); return rc; }
  if (uv && texture && tex_w && tex_h) {
    if ((rc = upload(ctx, &m->d_uv, uv, (size_t)nt * 6)) != J3DG_OK) { j3dg_mesh_destroy(m); return rc; }
    if (cudaMalloc((void**)&m->d_texture, (size_t)tex_w * tex_h * 4) != cudaSuccess) { cudaGetLastError(); j3dg_mesh_destroy(m); j3dg_set_error(ctx, "out of device memory (texture)"); return J3DG_ENOMEM; }
    if (cudaMemcpy2DAsync(m->d_texture, (size_t)tex_w * 4, texture, (size_t)(tex_stride ? tex_stride : tex_w) * 4, (size_t)tex_w * 4, tex_h, cudaMemcpyDefault, ctx->stream) != cudaSuccess) {
      cudaGetLastError(); j3dg_mesh_destroy(m); j3dg_set_error(ctx, "texture upload failed"); return J3DG_ECUDA;
    }
    m->tex_w = tex_w; m->tex_h = tex_h;
  }
  cudaEventRecord(ctx->ev[7], ctx->stream);
  cudaEventSynchronize(ctx->ev[7]);
  m->info.upload_ms = elapsed(ctx->ev[6], ctx->ev[7]);
  rc = j3dg_build_bvh(m);
  if (rc != J3DG_OK) { j3dg_mesh_destroy(m); return rc; }
  *out = m;
  return J3DG_OK;
}

J3DG_API int j3dg_mesh_create_empty(j3dg_ctx* ctx, uint32_t nv, uint32_t nt, uint32_t nr_of_nodes, const float* cs, uint32_t db_id, j3dg_mesh** out) {
  if (!ctx || !out) return J3DG_EINVAL;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  j3dg_mesh* m = new j3dg_mesh();
  m->ctx = ctx; m->nv = nv; m->nt = nt; m->db_id = db_id;
  memcpy(m->cs, cs ? cs : kIdentity, sizeof(m->cs));
  j3dg_invert_orthonormal_host(m->cs, m->cs_inv);
  if ((nr_of_nodes && cudaMalloc((void**)&m->d_nodes, (size_t)nr_of_nodes * sizeof(WideNode)) != cudaSuccess) ||
      (nt && cudaMalloc((void**)&m->d_tris, ((size_t)nt + J3DG_TRI_PAD) * sizeof(TriRec)) != cudaSuccess)) {
    cudaGetLastError();
    j3dg_mesh_destroy(m);
    j3dg_set_error(ctx, "out of device memory (received BVH)");
    return J3DG_ENOMEM;
  }
  if (nt) CU_CHECK(ctx, cudaMemsetAsync(m->d_tris + nt, 0, J3DG_TRI_PAD * sizeof(TriRec), ctx->stream));
  m->node_cap = m->nr_nodes = nr_of_nodes;
  m->info.nr_of_vertices = nv; m->info.nr_of_triangles = nt; m->info.nr_of_nodes = nr_of_nodes; m->info.nr_of_leaf_triangles = nt;
  m->info.node_bytes = sizeof(WideNode); m->info.triangle_bytes = sizeof(TriRec);
  *out = m;
  return J3DG_OK;
}

J3DG_API void j3dg_mesh_destroy(j3dg_mesh* m) {
  if (!m) return;
  if (m->ctx) { cudaSetDevice(m->ctx->device); cudaStreamSynchronize(m->ctx->stream); }
  cudaFree(m->d_vertices); cudaFree(m->d_indices); cudaFree(m->d_vcolors); cudaFree(m->d_uv); cudaFree(m->d_texture);
  cudaFree(m->d_nodes); cudaFree(m->d_tris);
  delete m;
}

J3DG_API int j3dg_mesh_rebuild(j3dg_mesh* m) {
  if (!m || !m->d_vertices) return J3DG_EINVAL;
  return j3dg_build_bvh(m);
}

J3DG_API int j3dg_mesh_info_get(const j3dg_mesh* m, j3dg_mesh_info* out) {
  if (!m || !out) return J3DG_EINVAL;
  *out = m->info;
  return J3DG_OK;
}

J3DG_API int j3dg_mesh_set_cs(j3dg_mesh* m, const float* cs) {
  if (!m) return J3DG_EINVAL;
  memcpy(m->cs, cs ? cs : kIdentity, sizeof(m->cs));
  j3dg_invert_orthonormal_host(m->cs, m->cs_inv);
  return J3DG_OK;
}

J3DG_API int j3dg_mesh_bvh_buffer(j3dg_mesh* m, int kind, void** dev_ptr, size_t* bytes) {
  if (!m || !dev_ptr || !bytes) return J3DG_EINVAL;
  if (kind == 0) { *dev_ptr = m->d_nodes; *bytes = (size_t)m->nr_nodes * sizeof(WideNode); }
  else if (kind == 1) { *dev_ptr = m->d_tris; *bytes = (size_t)m->nt * sizeof(TriRec); }
  else return J3DG_EINVAL;
  return J3DG_OK;
}

J3DG_API int j3dg_mesh_find_closest(j3dg_mesh* m, const float* rays, uint32_t n, float* hits, uint32_t* triangle_ids) {
  if (!m || (n && (!rays || !hits || !triangle_ids))) return J3DG_EINVAL;
  j3dg_ctx* ctx = m->ctx;
  if (!n) return J3DG_OK;
  // one allocation for the three staging arrays, released on every path
  const size_t off_hits = ((size_t)n * 32 + 255) & ~(size_t)255, off_ids = (off_hits + (size_t)n * 16 + 255) & ~(size_t)255;
  char* d_buf = nullptr;
  CU_CHECK(ctx, cudaMalloc((void**)&d_buf, off_ids + (size_t)n * 4));
  auto body = [&]() -> int {
    float* d_rays = (float*)d_buf;
    float* d_hits = (float*)(d_buf + off_hits);
    uint32_t* d_ids = (uint32_t*)(d_buf + off_ids);
    CU_CHECK(ctx, cudaMemcpyAsync(d_rays, rays, (size_t)n * 32, cudaMemcpyDefault, ctx->stream));
    int rc = j3dg_launch_find_closest(m, d_rays, n, d_hits, d_ids);
    if (rc != J3DG_OK) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(hits, d_hits, (size_t)n * 16, cudaMemcpyDefault, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(triangle_ids, d_ids, (size_t)n * 4, cudaMemcpyDefault, ctx->stream));
    return check_overflow(ctx);
  };
  const int rc = body();
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_buf);
  return rc;
}

// ---- clouds ------------------------------------------------------------------------------------
J3DG_API int j3dg_cloud_create(j3dg_ctx* ctx, const float* positions, const float* normals, const uint32_t* colors, uint32_t n,
                               const float* cs, uint32_t db_id, j3dg_cloud** out) {
  if (!ctx || !out || (n && !positions)) { j3dg_set_error(ctx, "j3dg_cloud_create: bad argument"); return J3DG_EINVAL; }
  if (n >= 0xFFFFFFFEu) { j3dg_set_error(ctx, "j3dg_cloud_create: too many points"); return J3DG_EINVAL; }
  *out = nullptr;
  cudaSetDevice(ctx->device);
  j3dg_cloud* c = new j3dg_cloud();
  c->ctx = ctx; c->n = n; c->db_id = db_id;
  memcpy(c->cs, cs ? cs : kIdentity, sizeof(c->cs));
  int rc;
  if ((rc = upload(ctx, &c->d_pos, positions, (size_t)n * 3)) != J3DG_OK || (rc = upload(ctx, &c->d_nrm, normals, (size_t)n * 3)) != J3DG_OK ||
      (rc = upload(ctx, &c->d_clr, colors, (size_t)n)) != J3DG_OK) {
    j3dg_cloud_destroy(c);
    return rc;
  }
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  *out = c;
  return J3DG_OK;
}

J3DG_API void j3dg_cloud_destroy(j3dg_cloud* c) {
  if (!c) return;
  if (c->ctx) { cudaSetDevice(c->ctx->device); cudaStreamSynchronize(c->ctx->stream); }
  cudaFree(c->d_pos); cudaFree(c->d_nrm); cudaFree(c->d_clr);
  delete c;
}

// ---- frame stages ------------------------------------------------------------------------------
namespace {

int stage_matcap(j3dg_ctx* ctx, const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity,
                 const uint32_t** d_out, uint32_t* ow, uint32_t* oh, uint32_t* ostride, uint32_t* ocav) {
  if (!matcap) {
    if (!ctx->d_matcap || !ctx->mw) { j3dg_set_error(ctx, "no matcap: pass one or call j3dg_ctx_set_matcap first"); return J3DG_EINVAL; }
    *d_out = ctx->d_matcap; *ow = ctx->mw; *oh = ctx->mh; *ostride = ctx->mstride; *ocav = ctx->cavity;
    return J3DG_OK;
  }
  if (!mw || !mh) { j3dg_set_error(ctx, "empty matcap"); return J3DG_EINVAL; }
  if (!mstride) mstride = mw;
  if (j3dg_is_device_ptr(matcap)) { *d_out = matcap; *ow = mw; *oh = mh; *ostride = mstride; *ocav = cavity; return J3DG_OK; }
  void* p = ctx->d_matcap;
  int rc = j3dg_reserve(ctx, &p, &ctx->matcap_cap, (size_t)mw * mh * 4);
  ctx->d_matcap = (uint32_t*)p;
  if (rc != J3DG_OK) return rc;
  CU_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_matcap, (size_t)mw * 4, matcap, (size_t)mstride * 4, (size_t)mw * 4, mh, cudaMemcpyDefault, ctx->stream));
  ctx->mw = mw; ctx->mh = mh; ctx->mstride = mw; ctx->cavity = cavity;
  *d_out = ctx->d_matcap; *ow = mw; *oh = mh; *ostride = mw; *ocav = cavity;
  return J3DG_OK;
}

int ensure_background(j3dg_ctx* ctx, uint32_t w, uint32_t h, uint32_t top, uint32_t bottom) {
  if (ctx->d_bg && ctx->bg_w == w && ctx->bg_h == h && ctx->bg_top == top && ctx->bg_bottom == bottom) return J3DG_OK;
  void* p = ctx->d_bg;
  int rc = j3dg_reserve(ctx, &p, &ctx->bg_cap, (size_t)w * h * 4);
  ctx->d_bg = (uint32_t*)p;
  if (rc != J3DG_OK) return rc;
  rc = j3dg_launch_background(ctx, w, h, top, bottom, ctx->d_bg, w);
  if (rc != J3DG_OK) return rc;
  ctx->bg_w = w; ctx->bg_h = h; ctx->bg_top = top; ctx->bg_bottom = bottom;
  return J3DG_OK;
}

}  // namespace

J3DG_API int j3dg_ctx_set_matcap(j3dg_ctx* ctx, const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity) {
  if (!ctx || !matcap || !mw || !mh) return J3DG_EINVAL;
  if (!mstride) mstride = mw;
  void* p = ctx->d_matcap;
  int rc = j3dg_reserve(ctx, &p, &ctx->matcap_cap, (size_t)mw * mh * 4);
  ctx->d_matcap = (uint32_t*)p;
  if (rc != J3DG_OK) return rc;
  CU_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_matcap, (size_t)mw * 4, matcap, (size_t)mstride * 4, (size_t)mw * 4, mh, cudaMemcpyDefault, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->mw = mw; ctx->mh = mh; ctx->mstride = mw; ctx->cavity = cavity;
  return J3DG_OK;
}

J3DG_API int j3dg_cast(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, const j3dg_view* view, int x0, int y0, int x1, int y1,
                       j3dg_pixel* pixels_out, uint32_t stride) {
  if (!ctx || !view || !pixels_out || (nm && !meshes)) { j3dg_set_error(ctx, "j3dg_cast: bad argument"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  { const int st = j3dg_check_sticky(ctx); if (st != J3DG_OK) return st; }
  const uint32_t w = view->width, h = view->height;
  if (!w || !h) return J3DG_OK;
  if (!stride) stride = w;
  if (ctx->shard_world > 1 && !(x0 <= 0 && y0 <= 0 && x1 >= (int)w - 1 && y1 >= (int)h - 1)) {
    j3dg_set_error(ctx, "j3dg_cast: with screen sharding on (j3dg_ctx_set_screen_shard) only whole-canvas rectangles are accepted");
    return J3DG_EINVAL;
  }
  if (j3dg_is_device_ptr(pixels_out)) return j3dg_launch_cast(ctx, meshes, nm, view, x0, y0, x1, y1, pixels_out, stride, false);
  // host destination: render into the context's device canvas, copy the updated rectangle back
  int rc = j3dg_reserve(ctx, &ctx->d_pixels, &ctx->pixels_cap, (size_t)w * h * sizeof(j3dg_pixel));
  if (rc != J3DG_OK) return rc;
  rc = j3dg_launch_cast(ctx, meshes, nm, view, x0, y0, x1, y1, (j3dg_pixel*)ctx->d_pixels, w, false);
  if (rc != J3DG_OK) return rc;
  ctx->last_canvas = ctx->d_pixels; ctx->last_w = w; ctx->last_h = h;
  int cx0 = std::min(std::max(x0, 0), (int)w - 1), cy0 = std::min(std::max(y0, 0), (int)h - 1);
  int cx1 = std::min(std::max(x1, 0), (int)w - 1), cy1 = std::min(std::max(y1, 0), (int)h - 1);
  if (cx1 >= cx0 && cy1 >= cy0) {
    CU_CHECK(ctx, cudaMemcpy2DAsync(pixels_out + (size_t)cy0 * stride + cx0, (size_t)stride * sizeof(j3dg_pixel),
                                    (j3dg_pixel*)ctx->d_pixels + (size_t)cy0 * w + cx0, (size_t)w * sizeof(j3dg_pixel),
                                    (size_t)(cx1 - cx0 + 1) * sizeof(j3dg_pixel), (size_t)(cy1 - cy0 + 1), cudaMemcpyDeviceToHost, ctx->stream));
  }
  return check_overflow(ctx);
}

J3DG_API int j3dg_shade(j3dg_ctx* ctx, const j3dg_pixel* pixels, uint32_t pixel_stride, const j3dg_view* view, const uint32_t* matcap,
                        uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr, const uint32_t* background,
                        uint32_t* rgba_inout, uint32_t rgba_stride) {
  if (!ctx || !view || !pixels || !rgba_inout) { j3dg_set_error(ctx, "j3dg_shade: bad argument"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  const uint32_t w = view->width, h = view->height;
  if (!w || !h) return J3DG_OK;
  if (!pixel_stride) pixel_stride = w;
  if (!rgba_stride) rgba_stride = w;
  const uint32_t* d_mc; uint32_t dmw, dmh, dms, dcav;
  int rc = stage_matcap(ctx, matcap, mw, mh, mstride, cavity_clr, &d_mc, &dmw, &dmh, &dms, &dcav);
  if (rc != J3DG_OK) return rc;
  const bool px_dev = j3dg_is_device_ptr(pixels), out_dev = j3dg_is_device_ptr(rgba_inout);
  const j3dg_pixel* d_px = pixels;
  uint32_t d_pstride = pixel_stride;
  if (!px_dev) {
    rc = j3dg_reserve(ctx, &ctx->d_pixels_in, &ctx->pixels_in_cap, (size_t)w * h * sizeof(j3dg_pixel));
    if (rc != J3DG_OK) return rc;
    CU_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_pixels_in, (size_t)w * sizeof(j3dg_pixel), pixels, (size_t)pixel_stride * sizeof(j3dg_pixel),
                                    (size_t)w * sizeof(j3dg_pixel), h, cudaMemcpyHostToDevice, ctx->stream));
    d_px = (const j3dg_pixel*)ctx->d_pixels_in;
    d_pstride = w;
  }
  uint32_t* d_out = rgba_inout;
  uint32_t d_rstride = rgba_stride;
  if (!out_dev) {
    void* p = ctx->d_rgba;
    rc = j3dg_reserve(ctx, &p, &ctx->rgba_cap, (size_t)w * h * 4);
    ctx->d_rgba = (uint32_t*)p;
    if (rc != J3DG_OK) return rc;
    d_out = ctx->d_rgba;
    d_rstride = w;
    if (!background)  // miss pixels keep the caller's content (canvas::im after the background copy)
      CU_CHECK(ctx, cudaMemcpy2DAsync(d_out, (size_t)w * 4, rgba_inout, (size_t)rgba_stride * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, ctx->stream));
  }
  const uint32_t* d_bg = nullptr;
  uint32_t bg_stride = w;
  if (background) {
    if (j3dg_is_device_ptr(background)) d_bg = background;
    else {
      void* p = ctx->d_bg;
      rc = j3dg_reserve(ctx, &p, &ctx->bg_cap, (size_t)w * h * 4);
      ctx->d_bg = (uint32_t*)p;
      if (rc != J3DG_OK) return rc;
      ctx->bg_w = 0;  // content is the caller's, not a generated gradient
      CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_bg, background, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
      d_bg = ctx->d_bg;
    }
  }
  rc = j3dg_launch_shade(ctx, d_px, d_pstride, view, d_mc, dmw, dmh, dms, dcav, d_bg, bg_stride, d_out, d_rstride);
  if (rc != J3DG_OK) return rc;
  if (!out_dev) {
    CU_CHECK(ctx, cudaMemcpy2DAsync(rgba_inout, (size_t)rgba_stride * 4, d_out, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return J3DG_OK;
}

J3DG_API int j3dg_splat(j3dg_ctx* ctx, j3dg_cloud* const* clouds, uint32_t nc, const j3dg_view* view, const j3dg_pixel* pixels_in,
                        j3dg_pixel* pixels_inout, uint32_t pixel_stride, uint32_t* rgba_inout, uint32_t rgba_stride) {
  if (!ctx || !view || !pixels_in || !pixels_inout || !rgba_inout || (nc && !clouds)) { j3dg_set_error(ctx, "j3dg_splat: bad argument"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  const uint32_t w = view->width, h = view->height;
  if (!w || !h || !nc) return J3DG_OK;
  if (!pixel_stride) pixel_stride = w;
  if (!rgba_stride) rgba_stride = w;
  const bool in_dev = j3dg_is_device_ptr(pixels_in), io_dev = j3dg_is_device_ptr(pixels_inout), rgba_dev = j3dg_is_device_ptr(rgba_inout);
  if (in_dev && io_dev && rgba_dev)
    return j3dg_launch_splat(ctx, clouds, nc, view, pixels_in, pixels_inout, pixel_stride, rgba_inout, rgba_stride);
  if (in_dev || io_dev || rgba_dev) { j3dg_set_error(ctx, "j3dg_splat: buffers must be all host or all device"); return J3DG_EINVAL; }
  int rc = j3dg_reserve(ctx, &ctx->d_pixels_in, &ctx->pixels_in_cap, (size_t)w * h * sizeof(j3dg_pixel));
  if (rc != J3DG_OK) return rc;
  rc = j3dg_reserve(ctx, &ctx->d_pixels, &ctx->pixels_cap, (size_t)w * h * sizeof(j3dg_pixel));
  if (rc != J3DG_OK) return rc;
  void* p = ctx->d_rgba;
  rc = j3dg_reserve(ctx, &p, &ctx->rgba_cap, (size_t)w * h * 4);
  ctx->d_rgba = (uint32_t*)p;
  if (rc != J3DG_OK) return rc;
  const size_t prow = (size_t)w * sizeof(j3dg_pixel);
  CU_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_pixels_in, prow, pixels_in, (size_t)pixel_stride * sizeof(j3dg_pixel), prow, h, cudaMemcpyHostToDevice, ctx->stream));
  CU_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_pixels, prow, pixels_inout, (size_t)pixel_stride * sizeof(j3dg_pixel), prow, h, cudaMemcpyHostToDevice, ctx->stream));
  CU_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_rgba, (size_t)w * 4, rgba_inout, (size_t)rgba_stride * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, ctx->stream));
  rc = j3dg_launch_splat(ctx, clouds, nc, view, (const j3dg_pixel*)ctx->d_pixels_in, (j3dg_pixel*)ctx->d_pixels, w, ctx->d_rgba, w);
  if (rc != J3DG_OK) return rc;
  CU_CHECK(ctx, cudaMemcpy2DAsync(pixels_inout, (size_t)pixel_stride * sizeof(j3dg_pixel), ctx->d_pixels, prow, prow, h, cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaMemcpy2DAsync(rgba_inout, (size_t)rgba_stride * 4, ctx->d_rgba, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return J3DG_OK;
}

// ---- dirty-rectangle readback (j3dg_ctx_set_dirty_rect) -------------------------------------------------------
// Outside the bounding rectangle of the hit pixels a frame holds nothing but miss records and background, which a
// host buffer that received an earlier frame of the same canvas already contains.  Only the union of the rectangle
// this host buffer received last time and the current one crosses PCIe.
namespace {

struct Rect { int x0, y0, x1, y1; bool empty() const { return x1 < x0 || y1 < y0; } };

Rect rect_union(Rect a, Rect b) {
  if (a.empty()) return b;
  if (b.empty()) return a;
  return Rect{std::min(a.x0, b.x0), std::min(a.y0, b.y0), std::max(a.x1, b.x1), std::max(a.y1, b.y1)};
}

// Copies what `host` needs of the device image `dev` (elem bytes per pixel, both w pixels per row) on `stream`.
int copy_dirty(j3dg_ctx* ctx, void* host, const void* dev, uint32_t w, uint32_t h, size_t elem, uint32_t key0, uint32_t key1, Rect cur, cudaStream_t stream) {
  j3dg_ctx::DirtyBuf* e = nullptr;
  for (auto& d : ctx->dirty_bufs)
    if (d.ptr == host) { e = &d; break; }
  Rect todo;
  if (e && e->w == w && e->h == h && e->elem == (uint32_t)elem && e->key0 == key0 && e->key1 == key1) {
    todo = rect_union(Rect{e->x0, e->y0, e->x1, e->y1}, cur);
  } else {
    todo = Rect{0, 0, (int)w - 1, (int)h - 1};
    if (!e) {
      if (ctx->dirty_bufs.size() >= 16) ctx->dirty_bufs.erase(ctx->dirty_bufs.begin());
      ctx->dirty_bufs.push_back({});
      e = &ctx->dirty_bufs.back();
    }
  }
  *e = j3dg_ctx::DirtyBuf{host, w, h, (uint32_t)elem, key0, key1, cur.x0, cur.y0, cur.x1, cur.y1};
  if (todo.empty()) return J3DG_OK;
  ctx->readback_bytes += (uint64_t)(todo.x1 - todo.x0 + 1) * (todo.y1 - todo.y0 + 1) * elem;
  const size_t pitch = (size_t)w * elem, off = ((size_t)todo.y0 * w + todo.x0) * elem;
  CU_CHECK(ctx, cudaMemcpy2DAsync((char*)host + off, pitch, (const char*)dev + off, pitch, (size_t)(todo.x1 - todo.x0 + 1) * elem,
                                  (size_t)(todo.y1 - todo.y0 + 1), cudaMemcpyDeviceToHost, stream));
  return J3DG_OK;
}

Rect bbox_rect(const uint32_t* bb, uint32_t w, uint32_t h) {  // {min x, min y, max x, max y} as the resolve kernel left them
  if (bb[0] > bb[2] || bb[1] > bb[3]) return Rect{0, 0, -1, -1};
  return Rect{(int)std::min(bb[0], w - 1), (int)std::min(bb[1], h - 1), (int)std::min(bb[2], w - 1), (int)std::min(bb[3], h - 1)};
}

// Enqueue the device->host copies of a submitted frame whose kernels have finished (pipelined path).
int flush_slot_copies(j3dg_ctx* ctx, int si) {
  j3dg_ctx::FrameSlot& sl = ctx->slot[si];
  if (!sl.copies_pending) return J3DG_OK;
  CU_CHECK(ctx, cudaEventSynchronize(sl.kernels_done));  // the hit bbox of this frame is on the host now
  const Rect cur = bbox_rect(ctx->h_overflow + 8 * si + 1, sl.w, sl.h);
  int rc;
  if (sl.host_px && (rc = copy_dirty(ctx, sl.host_px, sl.d_px, sl.w, sl.h, sizeof(j3dg_pixel), 0u, 0u, cur, ctx->copy_stream)) != J3DG_OK) return rc;
  if (sl.host_rgba && (rc = copy_dirty(ctx, sl.host_rgba, sl.d_rgba, sl.w, sl.h, 4, sl.bg_top, sl.bg_bottom, cur, ctx->copy_stream)) != J3DG_OK) return rc;
  CU_CHECK(ctx, cudaEventRecord(sl.copy_done, ctx->copy_stream));
  sl.copies_pending = false;
  return J3DG_OK;
}

}  // namespace

J3DG_API int j3dg_ctx_readback_bytes(j3dg_ctx* ctx, uint64_t* bytes, int reset) {
  if (!ctx || !bytes) return J3DG_EINVAL;
  *bytes = ctx->readback_bytes;
  if (reset) ctx->readback_bytes = 0;
  return J3DG_OK;
}

J3DG_API int j3dg_ctx_set_dirty_rect(j3dg_ctx* ctx, int enabled) {
  if (!ctx) return J3DG_EINVAL;
  ctx->dirty_rect = enabled != 0;
  ctx->dirty_bufs.clear();
  return J3DG_OK;
}

J3DG_API int j3dg_render_frame(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, j3dg_cloud* const* clouds, uint32_t nc,
                               const j3dg_view* view, const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr,
                               uint32_t bg_top, uint32_t bg_bottom, j3dg_pixel* pixels_out, uint32_t* rgba_out) {
  if (!ctx || !view || (nm && !meshes) || (nc && !clouds)) { j3dg_set_error(ctx, "j3dg_render_frame: bad argument"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  { const int st = j3dg_check_sticky(ctx); if (st != J3DG_OK) return st; }
  const uint32_t w = view->width, h = view->height;
  if (!w || !h) return J3DG_OK;
  const uint32_t* d_mc; uint32_t dmw, dmh, dms, dcav;
  int rc = stage_matcap(ctx, matcap, mw, mh, mstride, cavity_clr, &d_mc, &dmw, &dmh, &dms, &dcav);
  if (rc != J3DG_OK) return rc;
  if ((rc = ensure_background(ctx, w, h, bg_top, bg_bottom)) != J3DG_OK) return rc;
  const bool px_dev = j3dg_is_device_ptr(pixels_out), rgba_dev = j3dg_is_device_ptr(rgba_out);
  j3dg_pixel* d_px = pixels_out;
  if (!px_dev) {
    if ((rc = j3dg_reserve(ctx, &ctx->d_pixels, &ctx->pixels_cap, (size_t)w * h * sizeof(j3dg_pixel))) != J3DG_OK) return rc;
    d_px = (j3dg_pixel*)ctx->d_pixels;
  }
  uint32_t* d_rgba = rgba_out;
  if (!rgba_dev) {
    void* p = ctx->d_rgba;
    rc = j3dg_reserve(ctx, &p, &ctx->rgba_cap, (size_t)w * h * 4);
    ctx->d_rgba = (uint32_t*)p;
    if (rc != J3DG_OK) return rc;
    d_rgba = ctx->d_rgba;
  }
  // view::render_scene: cast -> shade (on the pre-splat records) -> splat
  if ((rc = j3dg_launch_cast(ctx, meshes, nm, view, 0, 0, (int)w - 1, (int)h - 1, d_px, w, false)) != J3DG_OK) return rc;
  if ((rc = j3dg_launch_shade(ctx, d_px, w, view, d_mc, dmw, dmh, dms, dcav, ctx->d_bg, w, d_rgba, w)) != J3DG_OK) return rc;
  if (nc && (rc = j3dg_launch_splat(ctx, clouds, nc, view, d_px, d_px, w, d_rgba, w)) != J3DG_OK) return rc;
  ctx->last_canvas = d_px; ctx->last_w = w; ctx->last_h = h;
  bool copied = false;
  if (ctx->dirty_rect && !nc && ctx->shard_world == 1 && ((pixels_out && !px_dev) || (rgba_out && !rgba_dev))) {
    uint32_t info[5];
    CU_CHECK(ctx, cudaMemcpyAsync(info, reinterpret_cast<uint32_t*>(ctx->d_stats + 2), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(info + 1, ctx->d_stats + 20, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    const Rect cur = bbox_rect(info + 1, w, h);
    if (pixels_out && !px_dev && (rc = copy_dirty(ctx, pixels_out, d_px, w, h, sizeof(j3dg_pixel), 0u, 0u, cur, ctx->stream)) != J3DG_OK) return rc;
    if (rgba_out && !rgba_dev && (rc = copy_dirty(ctx, rgba_out, d_rgba, w, h, 4, bg_top, bg_bottom, cur, ctx->stream)) != J3DG_OK) return rc;
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (info[0]) { j3dg_set_error(ctx, "traversal stack overflow (BVH deeper than the kernel's stack)"); return J3DG_ECUDA; }
    return J3DG_OK;
  }
  if (pixels_out && !px_dev) { CU_CHECK(ctx, cudaMemcpyAsync(pixels_out, d_px, (size_t)w * h * sizeof(j3dg_pixel), cudaMemcpyDeviceToHost, ctx->stream)); copied = true; ctx->readback_bytes += (uint64_t)w * h * sizeof(j3dg_pixel); }
  if (rgba_out && !rgba_dev) { CU_CHECK(ctx, cudaMemcpyAsync(rgba_out, d_rgba, (size_t)w * h * 4, cudaMemcpyDeviceToHost, ctx->stream)); copied = true; ctx->readback_bytes += (uint64_t)w * h * 4; }
  if (copied) return check_overflow(ctx);
  return J3DG_OK;
}

// ---- pipelined frames ---------------------------------------------------------------------------
// j3dg_render_frame with host outputs costs kernels + PCIe copy back to back.  For sweeps (orbit
// renders, turntables) j3dg_frame_submit enqueues the kernels on the context stream and the
// device->host copies on a second stream, double-buffering the device canvases, so the copy of
// frame k overlaps the kernels of frame k+1; j3dg_frame_wait returns when the oldest frame's host
// buffers are complete.
J3DG_API int j3dg_frame_submit(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, j3dg_cloud* const* clouds, uint32_t nc,
                               const j3dg_view* view, const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr,
                               uint32_t bg_top, uint32_t bg_bottom, j3dg_pixel* pixels_out, uint32_t* rgba_out) {
  if (!ctx || !view || (nm && !meshes) || (nc && !clouds)) { j3dg_set_error(ctx, "j3dg_frame_submit: bad argument"); return J3DG_EINVAL; }
  if (j3dg_is_device_ptr(pixels_out) || j3dg_is_device_ptr(rgba_out)) { j3dg_set_error(ctx, "j3dg_frame_submit: outputs must be host buffers"); return J3DG_EINVAL; }
  if (ctx->frames_submitted - ctx->frames_waited >= 2) { j3dg_set_error(ctx, "j3dg_frame_submit: two frames already in flight, call j3dg_frame_wait"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  { const int st = j3dg_check_sticky(ctx); if (st != J3DG_OK) return st; }
  const uint32_t w = view->width, h = view->height;
  if (!w || !h) { j3dg_set_error(ctx, "j3dg_frame_submit: empty canvas"); return J3DG_EINVAL; }
  if (!ctx->copy_stream) {
    CU_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CU_CHECK(ctx, cudaHostAlloc((void**)&ctx->h_overflow, 16 * sizeof(uint32_t), cudaHostAllocDefault));
    for (auto& sl : ctx->slot) {
      CU_CHECK(ctx, cudaEventCreateWithFlags(&sl.kernels_done, cudaEventDisableTiming));
      CU_CHECK(ctx, cudaEventCreateWithFlags(&sl.copy_done, cudaEventDisableTiming));
    }
  }
  const int si = (int)(ctx->frames_submitted & 1);
  j3dg_ctx::FrameSlot& sl = ctx->slot[si];
  const uint32_t* d_mc; uint32_t dmw, dmh, dms, dcav;
  int rc = stage_matcap(ctx, matcap, mw, mh, mstride, cavity_clr, &d_mc, &dmw, &dmh, &dms, &dcav);
  if (rc != J3DG_OK) return rc;
  if ((rc = ensure_background(ctx, w, h, bg_top, bg_bottom)) != J3DG_OK) return rc;
  if ((rc = j3dg_reserve(ctx, &sl.d_px, &sl.px_cap, (size_t)w * h * sizeof(j3dg_pixel))) != J3DG_OK) return rc;
  if ((rc = j3dg_reserve(ctx, &sl.d_rgba, &sl.rgba_cap, (size_t)w * h * 4)) != J3DG_OK) return rc;
  j3dg_pixel* d_px = (j3dg_pixel*)sl.d_px;
  uint32_t* d_rgba = (uint32_t*)sl.d_rgba;
  // the kernels of this frame may overwrite the slot only after the copy of the frame before last left it
  if (sl.busy) CU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, sl.copy_done, 0));
  if ((rc = j3dg_launch_cast(ctx, meshes, nm, view, 0, 0, (int)w - 1, (int)h - 1, d_px, w, false)) != J3DG_OK) return rc;
  if ((rc = j3dg_launch_shade(ctx, d_px, w, view, d_mc, dmw, dmh, dms, dcav, ctx->d_bg, w, d_rgba, w)) != J3DG_OK) return rc;
  if (nc && (rc = j3dg_launch_splat(ctx, clouds, nc, view, d_px, d_px, w, d_rgba, w)) != J3DG_OK) return rc;
  ctx->last_canvas = d_px; ctx->last_w = w; ctx->last_h = h;
  CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_overflow + 8 * si, reinterpret_cast<uint32_t*>(ctx->d_stats + 2), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_overflow + 8 * si + 1, ctx->d_stats + 20, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaEventRecord(sl.kernels_done, ctx->stream));
  if (ctx->dirty_rect && !nc && ctx->shard_world == 1 && (pixels_out || rgba_out)) {
    // the copies wait until the frame's hit bbox is known: they are enqueued by the next submit (after ITS kernels, so
    // the kernel stream never idles) or by the matching wait
    sl.copies_pending = true;
    sl.host_px = pixels_out; sl.host_rgba = rgba_out;
    sl.w = w; sl.h = h; sl.bg_top = bg_top; sl.bg_bottom = bg_bottom;
  } else {
    sl.copies_pending = false;
    CU_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, sl.kernels_done, 0));
    if (pixels_out) { CU_CHECK(ctx, cudaMemcpyAsync(pixels_out, d_px, (size_t)w * h * sizeof(j3dg_pixel), cudaMemcpyDeviceToHost, ctx->copy_stream)); ctx->readback_bytes += (uint64_t)w * h * sizeof(j3dg_pixel); }
    if (rgba_out) { CU_CHECK(ctx, cudaMemcpyAsync(rgba_out, d_rgba, (size_t)w * h * 4, cudaMemcpyDeviceToHost, ctx->copy_stream)); ctx->readback_bytes += (uint64_t)w * h * 4; }
    CU_CHECK(ctx, cudaEventRecord(sl.copy_done, ctx->copy_stream));
  }
  if ((rc = flush_slot_copies(ctx, si ^ 1)) != J3DG_OK) return rc;  // the frame before this one
  sl.busy = true;
  ctx->frames_submitted++;
  return J3DG_OK;
}

J3DG_API int j3dg_frame_wait(j3dg_ctx* ctx) {
  if (!ctx) return J3DG_EINVAL;
  if (ctx->frames_waited == ctx->frames_submitted) { j3dg_set_error(ctx, "j3dg_frame_wait: no frame in flight"); return J3DG_EINVAL; }
  const int si = (int)(ctx->frames_waited & 1);
  int rc = flush_slot_copies(ctx, si);
  if (rc != J3DG_OK) return rc;
  CU_CHECK(ctx, cudaEventSynchronize(ctx->slot[si].copy_done));
  ctx->frames_waited++;
  if (ctx->h_overflow[8 * si]) {
    j3dg_set_error(ctx, "traversal stack overflow (BVH deeper than the kernel's stack)");
    return J3DG_ECUDA;
  }
  return J3DG_OK;
}

J3DG_API int j3dg_cast_stats(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, const j3dg_view* view, double* nodes_per_ray, double* tris_per_ray) {
  if (!ctx || !view || (nm && !meshes)) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  const uint32_t w = view->width, h = view->height;
  if (!w || !h) return J3DG_EINVAL;
  int rc = j3dg_reserve(ctx, &ctx->d_pixels, &ctx->pixels_cap, (size_t)w * h * sizeof(j3dg_pixel));
  if (rc != J3DG_OK) return rc;
  j3dg_view v = *view;
  v.flags &= ~J3DG_SHADOW;  // primary rays only
  const bool prof = ctx->profiling;
  ctx->profiling = false;
  rc = j3dg_launch_cast(ctx, meshes, nm, &v, 0, 0, (int)w - 1, (int)h - 1, (j3dg_pixel*)ctx->d_pixels, w, true);
  if (ctx->last_canvas == ctx->d_pixels) ctx->last_canvas = nullptr;  // the counting pass leaves costs, not records
  ctx->profiling = prof;
  if (rc != J3DG_OK) return rc;
  unsigned long long st[4];
  CU_CHECK(ctx, cudaMemcpyAsync(st, ctx->d_stats, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const double rays = (double)w * h;
  if (nodes_per_ray) *nodes_per_ray = (double)st[0] / rays;
  if (tris_per_ray) *tris_per_ray = (double)st[1] / rays;
  return J3DG_OK;
}

// Diagnostic: the counting pass's per-pixel costs (node visits, triangle tests), w*h uint32 each, host pointers.
J3DG_API int j3dg_cast_cost_image(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, const j3dg_view* view,
                                  uint32_t* nodes_out, uint32_t* tris_out) {
  if (!ctx || !view || !nodes_out || !tris_out) return J3DG_EINVAL;
  double a, b;
  int rc = j3dg_cast_stats(ctx, meshes, nm, view, &a, &b);
  if (rc != J3DG_OK) return rc;
  const size_t n = (size_t)view->width * view->height;
  std::vector<j3dg_pixel> tmp(n);
  CU_CHECK(ctx, cudaMemcpy(tmp.data(), ctx->d_pixels, n * sizeof(j3dg_pixel), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    memcpy(&nodes_out[i], &tmp[i].u, 4);
    memcpy(&tris_out[i], &tmp[i].v, 4);
  }
  return J3DG_OK;
}
