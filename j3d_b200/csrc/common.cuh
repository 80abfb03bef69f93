// common.cuh — shared device/host definitions of the sm_100a renderer (libj3dg.so).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <string>
#include <vector>

#include "j3dg.h"

#define J3DG_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------
// Device data layout (DESIGN.md "Data layout in HBM")
// ---------------------------------------------------------------------------------------

// 8-wide BVH node, 128 bytes = exactly one L1/L2 cache line, 128-byte aligned.  The layout is
// CHILD-MAJOR because the traversal kernel maps the 8 children of a node onto 8 lanes of a warp
// (one ray per 8-lane group): lane c reads its own 8-byte quantised box and its own child
// reference, the two 16-byte headers are broadcast loads.
// Child boxes are quantised to 8 bits per plane relative to (origin, scale): the decoded plane
// fl(origin + q * scale) is guaranteed by the builder to contain the child's true box (scale is a
// power of two, so q * scale is exact).
// box[c] = { qlo.x, qlo.y, qlo.z, qhi.x, qhi.y, qhi.z, 0x80, 0x3F }: the two constant bytes let a
//          single PRMT assemble the float 0x3F80_qq_00 = 1 + q * 2^-15 from any plane byte (cast.cu, plane()).
// child[c]: bit31 = 0 -> index of an inner node; bit31 = 1 -> leaf, bits 0..30 = first triangle
//           record; a leaf is 1..8 consecutive records, the last one carries TriRec::last != 0.
// Empty slots have an inverted box (qlo = 255, qhi = 0) and child = J3DG_EMPTY_CHILD.
struct __align__(128) WideNode {
  float ox, oy, oz;      // quantisation origin = node box minimum
  uint32_t nchild;       // number of used slots (diagnostic)
  float sx, sy, sz;      // quantisation step per axis (a power of two) times 2^15
  uint32_t pad0;
  uint8_t box[8][8];     // [slot][qlo.xyz, qhi.xyz, 0x80, 0x3F]
  uint32_t child[8];
};
static_assert(sizeof(WideNode) == 128, "WideNode must be 128 bytes");

#define J3DG_LEAF_BIT 0x80000000u
#define J3DG_EMPTY_CHILD 0xFFFFFFFFu
#ifndef J3DG_MAX_LEAF
#define J3DG_MAX_LEAF 8       // triangles per leaf, at most 8 (the group kernel tests a leaf with 8 lanes)
#endif
#define J3DG_LEAF_FIRST_MASK 0x7FFFFFFFu
#define J3DG_TRI_PAD 8   // records allocated past the end: a group always loads 8 consecutive records

// Pre-gathered triangle record, 48 bytes, in Morton-sorted order so that every BVH subtree
// owns a contiguous range.  Raw vertex positions (the Woop test needs v - origin exactly as
// the reference computes it); v0.w carries the original triangle index, v1.w the end-of-leaf flag.
struct __align__(16) TriRec {
  float4 v0;  // xyz, w = __uint_as_float(original triangle index)
  float4 v1;  // xyz, w = __uint_as_float(1) on the last record of a leaf, else 0
  float4 v2;  // xyz, w unused (0)
};
static_assert(sizeof(TriRec) == 48, "TriRec must be 48 bytes");

// What a traversal kernel needs of one mesh.
struct MeshDev {
  const WideNode* nodes;
  const TriRec* tris;
  const uint32_t* indices;     // nt x 3 original indices (vertex colours / uv lookups)
  const float* vertices;       // nv x 3 (unused by traversal; kept for consumers)
  const float* vertex_colors;  // nullable
  const float* uv;             // nullable, nt x 6
  const uint32_t* texture;     // nullable
  uint32_t tex_w, tex_h, tex_stride;
  uint32_t nt;
  uint32_t db_id;
  float cs[16];                // object -> world
  float cs_inv[16];            // invert_orthonormal(cs), canvas.cpp:734
  float root_min[3], root_max[3];
};

struct ViewDev {
  uint32_t width, height;
  float near_plane, diagonal;
  float pinv[16];
  float cs[16];
  float cs_inv[16];
  float origin[4];  // cs * (0,0,0,1)
  float light[4];   // cs * (pivot + 3*diagonal, 1)
  uint32_t flags;
};

// ---------------------------------------------------------------------------------------
// Exactly-rounded arithmetic: the reference is built for plain SSE (no FMA), so every
// parity-critical product/sum must round separately.  These intrinsics are never contracted.
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
static __device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
static __device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
static __device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
static __device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
static __device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// jtk matrix_vector_multiply (qbvh.h:4564-4568): c0*v0 + c1*v1 + c2*v2 + c3*v3, left to right
static __device__ __forceinline__ float4 mat_vec(const float* __restrict__ m, float4 v) {
  float4 r;
  r.x = fadd(fadd(fadd(fmul(m[0], v.x), fmul(m[4], v.y)), fmul(m[8], v.z)), fmul(m[12], v.w));
  r.y = fadd(fadd(fadd(fmul(m[1], v.x), fmul(m[5], v.y)), fmul(m[9], v.z)), fmul(m[13], v.w));
  r.z = fadd(fadd(fadd(fmul(m[2], v.x), fmul(m[6], v.y)), fmul(m[10], v.z)), fmul(m[14], v.w));
  r.w = fadd(fadd(fadd(fmul(m[3], v.x), fmul(m[7], v.y)), fmul(m[11], v.z)), fmul(m[15], v.w));
  return r;
}
#endif

// ---------------------------------------------------------------------------------------
// Host-side plumbing
// ---------------------------------------------------------------------------------------
struct j3dg_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string error;
  bool profiling = true;
  cudaEvent_t ev[8] = {};          // [6],[7]: build / upload timing
  struct EventRing {               // begin/end event pairs of one stage, summed when timings are read
    std::vector<cudaEvent_t> a, b;
    uint32_t used = 0;
  } ring[3];                       // 0 cast, 1 shade, 2 splat
  uint64_t rays_primary = 0;
  uint32_t shadow_casts = 0;
  j3dg_timings timings = {};
  uint32_t launches = 0;
  int sm_count = 148;
  // frame scratch (grown on demand)
  void* d_pixels = nullptr; size_t pixels_cap = 0;
  void* d_pixels_in = nullptr; size_t pixels_in_cap = 0;
  uint32_t* d_rgba = nullptr; size_t rgba_cap = 0;
  uint32_t* d_bg = nullptr; size_t bg_cap = 0; uint32_t bg_w = 0, bg_h = 0, bg_top = 0, bg_bottom = 0;
  unsigned long long* d_packed = nullptr; size_t packed_cap = 0;
  uint32_t* d_matcap = nullptr; size_t matcap_cap = 0; uint32_t mw = 0, mh = 0, mstride = 0, cavity = 0;
  MeshDev* d_meshes = nullptr; size_t meshes_cap = 0;
  void* h_stage[4] = {}; cudaEvent_t stage_ev[4] = {}; bool stage_busy[4] = {};  // pinned ring of the pageable-source upload (api.cu, j3dg_copy_to_device)
  void* stage_init = nullptr;                        // std::thread* that pins the ring (and preloads the kernels) in the background from j3dg_ctx_create on
  volatile int stage_ready = 0;                      // 0: still pinning, 1: ring usable (or unavailable: h_stage[i] == nullptr)
  void* d_top = nullptr; size_t top_cap = 0; uint32_t top_nodes = 0;  // top-level tree over the objects of the uploaded mesh table (cast.cu)
  uint32_t top_min = 9;                              // scenes with at least this many objects are cast through the top-level tree (J3DG_TOP_MIN)
  unsigned long long* d_stats = nullptr;
  void* d_misc = nullptr; size_t misc_cap = 0;
  void* d_shadow = nullptr; size_t shadow_cap = 0;   // shadow ray list (origins + pixel offsets)
  // pipelined frames (j3dg_frame_submit / j3dg_frame_wait): double-buffered device canvases, a copy stream
  struct FrameSlot {
    void* d_px = nullptr; size_t px_cap = 0;
    void* d_rgba = nullptr; size_t rgba_cap = 0;
    cudaEvent_t kernels_done = nullptr, copy_done = nullptr;
    bool busy = false;
    // dirty-rectangle readback: the copies of a frame are enqueued once its hit bbox is known on the host
    bool copies_pending = false;
    void* host_px = nullptr; void* host_rgba = nullptr;
    uint32_t w = 0, h = 0, bg_top = 0, bg_bottom = 0;
  } slot[2];
  uint64_t readback_bytes = 0;                       // device->host bytes of frame outputs since the last reset (j3dg_ctx_readback_bytes)
  bool dirty_rect = false;                           // j3dg_ctx_set_dirty_rect
  struct DirtyBuf { void* ptr; uint32_t w, h, elem, key0, key1; int x0, y0, x1, y1; };  // what a host buffer holds outside-of-miss
  std::vector<DirtyBuf> dirty_bufs;
  cudaStream_t copy_stream = nullptr;
  uint64_t frames_submitted = 0, frames_waited = 0;
  uint32_t* h_overflow = nullptr;                    // pinned, 8 words per in-flight frame: [0] stack-overflow flag, [1..4] hit bbox
  std::vector<MeshDev> meshes_uploaded;              // last mesh table sent to d_meshes (re-uploaded only when it changes)
  void* d_spill = nullptr; size_t spill_cap = 0;     // pool-mode stacks that left shared memory (cast.cu)
  void* d_hard = nullptr; size_t hard_cap = 0;       // hard-ray list handed from the lane kernel to the group kernel
  size_t hard_id_off = 0;                            // offset of the id array inside d_hard
  uint32_t shadow_budget = 16;                       // the same for shadow rays (J3DG_SHADOW_BUDGET; 8 / 12 / 16 / 24 / 48: 1.64 / 1.60 / 1.58 / 1.65 / 1.70 ms for 0.73 M rays)
  uint32_t lane_budget = 28;                         // node visits per ray before the lane kernel evicts it (J3DG_LANE_BUDGET)
  int cast_algo = 0;                                 // 0 hybrid (lane + group), 1 group kernel only (J3DG_CAST_ALGO=group)
  void* last_canvas = nullptr; uint32_t last_w = 0, last_h = 0;  // device canvas of the most recent frame (j3dg_pick reads it)
  uint32_t shard_rank = 0, shard_world = 1;          // screen sharding (j3dg_ctx_set_screen_shard): band b of 32 rows belongs to rank b mod world
  // Sticky status words in mapped pinned host memory (kernels store into d_status, the host reads h_status without a
  // synchronisation): [0] a stream-ordered flag wait timed out (peer.cu), [1] a traversal stack overflowed (cast.cu).
  // Never cleared by a render call: j3dg_ctx_status(reset) clears them; j3dg_ctx_synchronize and the frame entry points
  // fail once one is set, so device-output calls (which return before the kernels ran) cannot lose the condition.
  volatile uint32_t* h_status = nullptr;
  uint32_t* d_status = nullptr;
};
int j3dg_check_sticky(j3dg_ctx* ctx);  // J3DG_OK, or the error of the first sticky status word that is set

void j3dg_set_error(j3dg_ctx* ctx, const std::string& msg);
int j3dg_stage_begin(j3dg_ctx* ctx, int stage);  // records the begin event of the next pair (no-op if !profiling)
int j3dg_stage_end(j3dg_ctx* ctx, int stage);
int j3dg_cuda_fail(j3dg_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);
bool j3dg_is_device_ptr(const void* p);
int j3dg_reserve(j3dg_ctx* ctx, void** ptr, size_t* cap, size_t bytes);
void j3dg_preload_build_kernels();
void j3dg_preload_cast_kernels();
int j3dg_copy_to_device(j3dg_ctx* ctx, void* dst, const void* src, size_t bytes);  // stream-ordered; pageable sources go through a pinned ring (api.cu)

#define CU_CHECK(ctx, call)                                                         \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) return j3dg_cuda_fail((ctx), e__, #call, __FILE__, __LINE__); \
  } while (0)
#define KERNEL_CHECK(ctx)                                                           \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) return j3dg_cuda_fail((ctx), e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)

struct j3dg_mesh {
  j3dg_ctx* ctx = nullptr;
  uint32_t nv = 0, nt = 0, db_id = 0;
  float* d_vertices = nullptr;
  uint32_t* d_indices = nullptr;
  float* d_vcolors = nullptr;
  float* d_uv = nullptr;
  uint32_t* d_texture = nullptr;
  uint32_t tex_w = 0, tex_h = 0;
  WideNode* d_nodes = nullptr; uint32_t node_cap = 0; uint32_t nr_nodes = 0;
  TriRec* d_tris = nullptr;
  float cs[16];
  float cs_inv[16];
  j3dg_mesh_info info = {};
};

struct j3dg_cloud {
  j3dg_ctx* ctx = nullptr;
  uint32_t n = 0, db_id = 0;
  float* d_pos = nullptr;
  float* d_nrm = nullptr;
  uint32_t* d_clr = nullptr;
  float cs[16];
};

// stage entry points implemented in the other translation units (device pointers only)
int j3dg_build_bvh(j3dg_mesh* m);
int j3dg_launch_cast(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nm, const j3dg_view* view,
                     int x0, int y0, int x1, int y1, j3dg_pixel* d_pixels, uint32_t stride, bool stats);
int j3dg_launch_shade(j3dg_ctx* ctx, const j3dg_pixel* d_pixels, uint32_t pstride, const j3dg_view* view,
                      const uint32_t* d_matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity,
                      const uint32_t* d_bg, uint32_t bg_stride, uint32_t* d_rgba, uint32_t rstride);
int j3dg_launch_splat(j3dg_ctx* ctx, j3dg_cloud* const* clouds, uint32_t nc, const j3dg_view* view,
                      const j3dg_pixel* d_px_in, j3dg_pixel* d_px_inout, uint32_t pstride, uint32_t* d_rgba, uint32_t rstride);
int j3dg_launch_background(j3dg_ctx* ctx, uint32_t w, uint32_t h, uint32_t top, uint32_t bottom, uint32_t* d_out, uint32_t stride);
int j3dg_launch_find_closest(j3dg_mesh* m, const float* d_rays, uint32_t n, float* d_hits, uint32_t* d_ids);
void j3dg_invert_orthonormal_host(const float* m, float* out);
