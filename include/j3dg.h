/*
 * j3dg.h — C ABI of the B200-native renderer for j3d's data-parallel hot path.
 *
 * This is the drop-in boundary: every entry point below replaces one call site of
 * the reference's std::thread/TBB renderer (j3d has no FFI/plugin layer of its own;
 * the seams are C++ call sites, cited per function as reference file:line relative
 * to the j3d source tree).  Plain pointers and sizes only; no C++/torch types.
 * All functions return 0 (J3DG_OK) on success or a negative J3DG_E* code; the text
 * of the last error is available through j3dg_last_error().  Nothing throws across
 * the boundary.  A context is single-caller (externally synchronised, exactly like
 * j3d's `canvas`, which view guards with one mutex — j3d/view.cpp:1262).
 *
 * Buffers named *_out / *_inout may be HOST or DEVICE pointers; the library detects
 * which (cudaPointerGetAttributes) and copies over PCIe only for host pointers.
 * Calls are synchronous on return for host buffers; for device buffers work is
 * enqueued on the context stream (j3dg_ctx_set_stream) and NOT synchronised.
 */
#ifndef J3DG_H
#define J3DG_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define J3DG_OK 0
#define J3DG_EINVAL (-1)   /* bad argument */
#define J3DG_ECUDA (-2)    /* CUDA runtime error (see j3dg_last_error) */
#define J3DG_ENOMEM (-3)   /* device allocation failed */
#define J3DG_ENODEV (-4)   /* no CUDA device / wrong architecture: there is NO CPU fallback */
#define J3DG_ETIMEOUT (-5) /* a stream-ordered flag wait (peer exchange) gave up: a peer is gone; sticky, see j3dg_ctx_status */

/* canvas::canvas_settings (j3d/canvas.h:16-25), one bit per bool, same order. */
#define J3DG_ONE_BIT      (1u << 0)
#define J3DG_SHADOW       (1u << 1)
#define J3DG_EDGES        (1u << 2)
#define J3DG_WIREFRAME    (1u << 3)
#define J3DG_SHADING      (1u << 4)
#define J3DG_TEXTURED     (1u << 5)
#define J3DG_VERTEXCOLORS (1u << 6)
/* settings.cpp:7-13 defaults: edges, shading, textured, vertexcolors */
#define J3DG_DEFAULT_FLAGS (J3DG_EDGES | J3DG_SHADING | J3DG_TEXTURED | J3DG_VERTEXCOLORS)

/* The 32-byte per-pixel record, byte-compatible with `struct pixel` (j3d/pixel.h:11-27).
 * NB the reference's naming: object_id is the TRIANGLE index within its mesh (or the
 * POINT index after a splat); db_id identifies the object; u,v are the camera-space
 * normal x,y; depth is the ray parameter t (== camera-space z distance). */
typedef struct j3dg_pixel {
  uint8_t mark, r, g, b; /* mark bit0 = shadowed, bit1 = r,g,b valid */
  float u, v;
  float depth;
  uint32_t object_id;
  float barycentric_u, barycentric_v;
  uint32_t db_id;
} j3dg_pixel;

/* Everything update_canvas / canvas_to_image / render_pointclouds_on_image read from
 * `canvas` and `scene` per frame (j3d/canvas.cpp:677-705, 287-300, 973-983).
 * Matrices are column-major float[16] exactly as jtk::float4x4 stores them. */
typedef struct j3dg_view {
  uint32_t width, height;     /* canvas size (im.width(), im.height()) */
  float near_plane;           /* camera::nearClippingPlane (0.1 by default) */
  float diagonal;             /* scene::diagonal (largest bbox extent, scene.cpp:86-88) */
  float projection[16];       /* canvas::projection_matrix */
  float projection_inv[16];   /* canvas::projection_matrix_inv */
  float cs[16];               /* scene::coordinate_system (camera -> world) */
  float cs_inv[16];           /* scene::coordinate_system_inv */
  float pivot[3];             /* scene::pivot (only the shadow light uses it) */
  uint32_t flags;             /* J3DG_* settings bits */
} j3dg_view;

typedef struct j3dg_ctx j3dg_ctx;
typedef struct j3dg_mesh j3dg_mesh;
typedef struct j3dg_cloud j3dg_cloud;

/* Sizes / timings of one built mesh (what j3d shows as "bvh construction (s)",
 * j3d/view.cpp:230-232, 894-898, plus the device layout numbers DESIGN.md quotes). */
typedef struct j3dg_mesh_info {
  uint32_t nr_of_vertices, nr_of_triangles;
  uint32_t nr_of_nodes;        /* 8-wide nodes */
  uint32_t nr_of_leaf_triangles; /* triangle records (== nr_of_triangles) */
  uint32_t node_bytes, triangle_bytes; /* sizeof one node / one triangle record */
  float build_ms;              /* device time of the last BVH build (CUDA events) */
  float upload_ms;             /* host->device copy time of the last create */
  float bbox_min[3], bbox_max[3];
  float sah_cost;              /* SAH cost of the wide BVH (diagnostic) */
} j3dg_mesh_info;

/* Per-stage device times, CUDA events on the ctx stream, SUMMED over every stage launch since
 * the last reset (j3dg_ctx_timings(..., reset=1)); *_count = launches summed. */
typedef struct j3dg_timings {
  float cast_ms, shade_ms, splat_ms, copy_ms;
  uint32_t cast_count, shade_count, splat_count;
  uint32_t kernel_launches;    /* kernels launched by the library since the last reset */
  uint64_t rays;               /* primary + shadow rays traced since the last reset */
} j3dg_timings;

/* ---- context ------------------------------------------------------------- */
/* One context per process per GPU (one process per GPU; the multi-GPU plumbing is the
 * j3dg_group / j3dg_frames section below).  Fails with J3DG_ENODEV if `device` is
 * not a CUDA device of compute capability 10.x.  */
int j3dg_ctx_create(int device, j3dg_ctx** out);
void j3dg_ctx_destroy(j3dg_ctx* ctx);
const char* j3dg_last_error(const j3dg_ctx* ctx); /* ctx may be NULL: global string */
/* Run all library work on this cudaStream_t (0 = the context's own stream). */
int j3dg_ctx_set_stream(j3dg_ctx* ctx, void* cuda_stream);
int j3dg_ctx_synchronize(j3dg_ctx* ctx);
int j3dg_ctx_timings(j3dg_ctx* ctx, j3dg_timings* out, int reset);
/* Sticky error conditions raised by kernels that ran AFTER the call that enqueued them returned (calls with device
 * outputs are not synchronised): a traversal stack overflow (pixels of that frame are wrong) or a timed-out flag
 * wait of the peer exchange.  Once set, j3dg_ctx_synchronize, j3dg_render_frame, j3dg_frame_submit / j3dg_frame_wait,
 * j3dg_cast and the j3dg_stream_* / j3dg_frames_* calls fail (J3DG_ECUDA / J3DG_ETIMEOUT) until the status is
 * read with reset != 0.  The host reads the words from mapped memory: no synchronisation is needed to poll. */
#define J3DG_STATUS_WAIT_TIMEOUT   (1u << 0)
#define J3DG_STATUS_STACK_OVERFLOW (1u << 1)
int j3dg_ctx_status(j3dg_ctx* ctx, uint32_t* flags_out, int reset);
/* Per-stage CUDA-event timing on/off (default on; off removes the event records). */
int j3dg_ctx_set_profiling(j3dg_ctx* ctx, int enabled);

/* Performance tuning of the ray cast; the rendered result does not depend on it (up to exact ties).
 * lane_budget: node visits a ray may spend in the one-ray-per-lane kernel before it is handed to the
 * one-ray-per-8-lanes kernel (0 = default); cast_algo: 0 = hybrid (default), 1 = 8-lane kernel only. */
int j3dg_ctx_set_tuning(j3dg_ctx* ctx, uint32_t lane_budget, int cast_algo);

/* Screen sharding for one frame rendered by several GPUs (BASELINE configs[2]): the canvas is cut
 * into bands of J3DG_SHARD_BAND_ROWS rows, band b belongs to rank b mod world.  With world > 1, j3dg_cast /
 * j3dg_render_frame trace only this rank's bands plus the single pixel row above each of them (the edge
 * shader's up neighbour, canvas.cpp:625-640 — recomputed instead of exchanged), and j3dg_shade writes only
 * this rank's rows; everything else in the output buffers is left untouched.  The result in the owned rows
 * is identical to the unsharded frame.  world = 1 (default) switches sharding off.  Point-cloud splats are not
 * sharded.  Bands are counted from row 0 of the canvas; while world > 1, j3dg_cast accepts whole-canvas
 * rectangles only (J3DG_EINVAL otherwise: cast and shade must agree on which rows a rank owns). */
#define J3DG_SHARD_BAND_ROWS 32
int j3dg_ctx_set_screen_shard(j3dg_ctx* ctx, uint32_t rank, uint32_t world);

/* ---- result exchange over NVLink peer memory (BASELINE configs[2], [4]).  The ray-cast kernel is a
 *      persistent launch that fills every SM, so a collective kernel finds no room beside it; instead every rank's
 *      shade kernel stores its RGBA straight into a buffer that rank 0 owns and the others map through CUDA IPC
 *      (pass the mapped address as rgba_out of j3dg_render_frame / j3dg_shade), and frames are handed over with
 *      stream-ordered flags.  j3dg_peer_alloc: zeroed device buffer + its 64-byte IPC handle (ship the handle to
 *      the other processes, e.g. with torch.distributed); j3dg_peer_open maps it in another process (peer access
 *      is enabled lazily); j3dg_stream_signal writes `value` to *flag after everything enqueued before it on the
 *      context stream is complete and visible system-wide; j3dg_stream_wait_geq blocks the STREAM (not the host)
 *      until flags[0..n) >= value (n <= 32; gives up after 5 s and records it: j3dg_stream_wait_status).  Flags
 *      may live in local or in peer-mapped memory.  j3d_b200/dist.py::PeerFrames is the protocol built on them. */
#define J3DG_IPC_HANDLE_BYTES 64
int j3dg_peer_alloc(j3dg_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char* handle_out /* 64 bytes, nullable */);
int j3dg_peer_free(j3dg_ctx* ctx, void* dev_ptr);
int j3dg_peer_open(j3dg_ctx* ctx, const unsigned char* handle, void** dev_ptr);
int j3dg_peer_close(j3dg_ctx* ctx, void* dev_ptr);
int j3dg_stream_signal(j3dg_ctx* ctx, uint32_t* flag, uint32_t value);
int j3dg_stream_wait_geq(j3dg_ctx* ctx, const uint32_t* flags, uint32_t n, uint32_t value);
int j3dg_stream_wait_status(j3dg_ctx* ctx, int* timed_out);

/* ---- multi-GPU: one process per GPU (SURVEY §8b "Context / multi-GPU", §8e).  j3d has no distributed code, so there
 *      is no reference interface to mirror; a C++ host that wants N GPUs does, in every process:
 *          j3dg_ctx_create(local_gpu, &ctx);
 *          j3dg_group_create(ctx, rank, world, id, &group);        // id: j3dg_group_unique_id() of one rank, shipped by any channel
 *          if (rank == 0) j3dg_mesh_create(ctx, ..., &mesh);        // load + build once ...
 *          j3dg_group_broadcast_mesh(group, 0, &mesh);              // ... replicate BVH + geometry into every GPU's HBM (NCCL over NVLink)
 *          j3dg_frames_create(group, w, h, 0, shared, &frames);     // rank 0 will hold everybody's frames
 *          per frame:  j3dg_frames_begin -> j3dg_render_frame(..., rgba_out = j3dg_frames_target) -> j3dg_frames_arrive
 *                      -> (rank 0 enqueues its consumer of j3dg_frames_view on the context stream) -> j3dg_frames_release
 *      The group owns an NCCL communicator (libnccl.so.2 is loaded with dlopen by the first j3dg_group_unique_id /
 *      j3dg_group_create; single-GPU users do not need it).  All group / frames calls are COLLECTIVE: every rank
 *      makes the same calls in the same order.  Work is enqueued on the context stream. -------------------------- */
typedef struct j3dg_group j3dg_group;
typedef struct j3dg_frames j3dg_frames;
#define J3DG_GROUP_ID_BYTES 128
int j3dg_group_unique_id(unsigned char id_out[J3DG_GROUP_ID_BYTES]);
int j3dg_group_create(j3dg_ctx* ctx, int rank, int world, const unsigned char id[J3DG_GROUP_ID_BYTES], j3dg_group** out);
void j3dg_group_destroy(j3dg_group* group);
int j3dg_group_rank(const j3dg_group* group);
int j3dg_group_world(const j3dg_group* group);
/* Every rank's context stream has drained up to this point on return (also reports sticky errors, j3dg_ctx_status). */
int j3dg_group_barrier(j3dg_group* group);
/* values[0..n) <- the maximum over the ranks (n <= 64): "time on the device as the max over ranks". */
int j3dg_group_max_float(j3dg_group* group, float* values, uint32_t n);
/* In-place per-word maximum over the ranks of n DEVICE uint64 words (packed (depth, index) splat words of a
 * point-range-sharded cloud).  Enqueued on the context stream. */
int j3dg_group_allreduce_max_u64(j3dg_group* group, unsigned long long* dev_words, size_t n);
/* Replicates a built mesh: `root` passes its mesh, every other rank passes a pointer to NULL and receives a new
 * mesh handle holding the same BVH, triangle records, vertices, indices, colours, uv and texture (destroy it with
 * j3dg_mesh_destroy).  Replaces "every process loads the file and builds its own BVH". */
int j3dg_group_broadcast_mesh(j3dg_group* group, int root, j3dg_mesh** mesh_inout);

/* Frame exchange: rank `dst` allocates [2 slots][world][height*width] RGBA (shared_frame != 0: [2 slots][1] — ONE frame
 * per slot that all ranks write disjoint rows of, for j3dg_ctx_set_screen_shard) and the other ranks map it through
 * CUDA IPC; a rank renders with rgba_out = j3dg_frames_target(k), so its shade kernel's stores travel over NVLink
 * into dst's HBM — no gather kernel has to find room beside the persistent ray-cast kernel.  Stream-ordered flags
 * hand the slots over (csrc/peer.cu): begin(k) makes the stream wait until dst has RELEASED frame k - 2 (same slot),
 * arrive(k) signals this rank's frame and, on dst, makes the stream wait for every rank's frame k; release(k)
 * (a no-op off dst) lets the peers overwrite the slot and must be enqueued AFTER the work that reads
 * j3dg_frames_view(k).  A peer that never arrives times the wait out after 5 s (J3DG_ETIMEOUT, sticky). */
int j3dg_frames_create(j3dg_group* group, uint32_t width, uint32_t height, int dst, int shared_frame, j3dg_frames** out);
/* The same with nslots (2 .. J3DG_FRAMES_MAX_SLOTS) instead of 2 slots: frame k lives in slot k mod nslots, begin(k)
 * waits for the release of frame k - nslots. */
#define J3DG_FRAMES_MAX_SLOTS 8
int j3dg_frames_create_n(j3dg_group* group, uint32_t width, uint32_t height, int dst, int shared_frame, uint32_t nslots, j3dg_frames** out);
void j3dg_frames_destroy(j3dg_frames* frames);
/* Several frames in flight per GPU: the frames of every slot may be rendered by their own CONTEXT of the same device
 * (own stream; see "frames in flight" at j3dg_render_frame): begin / arrive / release of frame k are enqueued on the
 * stream of the context registered for slot k mod nslots (default: the group's context for all).  The flag words are per
 * slot, so the streams never order each other.  Local (not collective). */
int j3dg_frames_set_lane(j3dg_frames* frames, int slot, j3dg_ctx* ctx);
int j3dg_frames_begin(j3dg_frames* frames, uint32_t* k_out);
int j3dg_frames_target(j3dg_frames* frames, uint32_t k, uint32_t** rgba_out);
int j3dg_frames_arrive(j3dg_frames* frames, uint32_t k);
int j3dg_frames_release(j3dg_frames* frames, uint32_t k);
/* dst only: device pointer to the frames of step k, [world][height*width] ([1][...] for a shared frame). */
int j3dg_frames_view(j3dg_frames* frames, uint32_t k, const uint32_t** frames_out);

/* ---- BVH build: replaces `new qbvh(triangles, vertices)` + compute_triangle_normals
 *      + compute_bb in add_object (j3d/scene.cpp:8-25; jtk/qbvh.h:1679-1686). -------
 * vertices: nv x 3 float (jtk::vec3<float>), triangles: nt x 3 uint32.
 * vertex_colors: nullable, nv x 3 float in [0,1] (mesh::vertex_colors).
 * uv: nullable, nt x 6 float (mesh::uv_coordinates), texture: nullable tex_h rows of
 * tex_stride uint32 0xAABBGGRR (jtk::image<uint32_t>).
 * cs: nullable (identity) column-major object->world matrix (mesh::cs).
 * The vertex/triangle arrays may be host or device pointers; they are copied. */
int j3dg_mesh_create(j3dg_ctx* ctx, const float* vertices, uint32_t nv,
                     const uint32_t* triangles, uint32_t nt,
                     const float* vertex_colors, const float* uv,
                     const uint32_t* texture, uint32_t tex_w, uint32_t tex_h, uint32_t tex_stride,
                     const float* cs, uint32_t db_id, j3dg_mesh** out);
void j3dg_mesh_destroy(j3dg_mesh* mesh);                 /* scene.cpp:40-48 remove_object */
int j3dg_mesh_rebuild(j3dg_mesh* mesh);                  /* rebuild the BVH from resident data */
int j3dg_mesh_info_get(const j3dg_mesh* mesh, j3dg_mesh_info* out);
int j3dg_mesh_set_cs(j3dg_mesh* mesh, const float* cs);  /* mesh::cs / scene_object::cs */
/* Device arrays of the built BVH (for NCCL broadcast above the ABI): kind 0 = wide
 * nodes, 1 = triangle records.  Returns device pointer + byte size. */
int j3dg_mesh_bvh_buffer(j3dg_mesh* mesh, int kind, void** dev_ptr, size_t* bytes);
/* Create a mesh whose BVH will be received rather than built: allocates node /
 * triangle arrays of the given sizes (then fill them via j3dg_mesh_bvh_buffer). */
int j3dg_mesh_create_empty(j3dg_ctx* ctx, uint32_t nv, uint32_t nt, uint32_t nr_of_nodes,
                           const float* cs, uint32_t db_id, j3dg_mesh** out);

/* Generic closest-hit query with qbvh::find_closest_triangle semantics
 * (jtk/qbvh.h:1701-1852): rays are n x 8 floats {ox,oy,oz, dx,dy,dz, t_near, t_far},
 * t may be negative, closest = smallest |t|, strict t_near < t < t_far.
 * hits are n x 4 floats {u, v, distance, found(0/1)}, ids n x uint32 triangle index. */
int j3dg_mesh_find_closest(j3dg_mesh* mesh, const float* rays, uint32_t n, float* hits, uint32_t* triangle_ids);

/* ---- ray cast: replaces canvas::update_canvas (j3d/canvas.cpp:677-874), i.e.
 *      qbvh_two_level_with_transformations::find_closest_triangle per pixel
 *      (jtk/qbvh.h:3303-3387) + hit->pixel + shadow ray.  Inclusive rect, clamped
 *      like the reference.  pixels_out: host or device, stride in pixels. -------- */
int j3dg_cast(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nr_of_meshes, const j3dg_view* view,
              int x0, int y0, int x1, int y1, j3dg_pixel* pixels_out, uint32_t stride);

/* ---- shading: replaces canvas::canvas_to_image (j3d/canvas.cpp:582-670) incl. the
 *      background copy of canvas::render_scene (canvas.cpp:893-898).
 *      matcap: mh rows of mstride uint32; background nullable (then rgba_inout keeps
 *      its content on miss pixels, like the reference's `im`). ------------------- */
int j3dg_shade(j3dg_ctx* ctx, const j3dg_pixel* pixels, uint32_t pixel_stride, const j3dg_view* view,
               const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr,
               const uint32_t* background, uint32_t* rgba_inout, uint32_t rgba_stride);

/* ---- point clouds: replaces canvas::render_pointclouds_on_image
 *      (j3d/canvas.cpp:952-1030) = jtk::bind/_draw/present (jtk/render.h:254-865). -- */
int j3dg_cloud_create(j3dg_ctx* ctx, const float* positions, const float* normals, const uint32_t* colors,
                      uint32_t n, const float* cs, uint32_t db_id, j3dg_cloud** out);
void j3dg_cloud_destroy(j3dg_cloud* cloud);
/* pixels_in: the buffer the z-buffer is seeded from (view::_pixels); pixels_inout: the
 * buffer the splat patches (canvas::_canvas); they may alias.  rgba_inout: canvas::im. */
int j3dg_splat(j3dg_ctx* ctx, j3dg_cloud* const* clouds, uint32_t nr_of_clouds, const j3dg_view* view,
               const j3dg_pixel* pixels_in, j3dg_pixel* pixels_inout, uint32_t pixel_stride,
               uint32_t* rgba_inout, uint32_t rgba_stride);

/* ---- whole frame: view::render_scene (j3d/view.cpp:421-430) in one call:
 *      cast -> (snapshot) -> shade -> splat, everything resident on the device; only
 *      the requested outputs cross PCIe.  pixels_out / rgba_out nullable.
 *
 *      Frames in flight.  One frame leaves most of the machine idle while its longest rays finish (the last fifth of the
 *      ray-cast kernel's run time on a 28 M-triangle mesh).  A sweep should therefore keep two or three frames in
 *      flight PER GPU, each through its OWN context of the same device (own stream, own scratch; j3dg_ctx_create
 *      may be called any number of times per device): the ray-cast kernel is a plain launch that never waits for a
 *      block that is not running, so the next frame's kernel moves into the SMs as this frame's blocks leave.  Meshes
 *      and clouds may be rendered from ANY context of the device they live on (they must not be rebuilt, moved or
 *      destroyed while a frame that uses them is in flight in another context).  Settings are per context
 *      (j3dg_ctx_set_matcap, _set_tuning, _set_dirty_rect, ...).  j3dg_ctx_synchronize / j3dg_frame_wait of the
 *      context that rendered a frame completes that frame. ---------- */
int j3dg_render_frame(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nr_of_meshes,
                      j3dg_cloud* const* clouds, uint32_t nr_of_clouds, const j3dg_view* view,
                      const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr,
                      uint32_t bg_top, uint32_t bg_bottom,
                      j3dg_pixel* pixels_out, uint32_t* rgba_out);
/* Pipelined variant for sweeps (orbit renders, turntables): submit enqueues the frame and returns;
 * the device->host copies run on a second stream, the device canvases are double-buffered, so the
 * copy of frame k overlaps the kernels of frame k+1.  pixels_out / rgba_out are HOST buffers
 * (page-locked for real overlap), nullable; each must stay untouched until the matching
 * j3dg_frame_wait returns.  At most two frames may be in flight; wait completes them in order. */
int j3dg_frame_submit(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nr_of_meshes,
                      j3dg_cloud* const* clouds, uint32_t nr_of_clouds, const j3dg_view* view,
                      const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr,
                      uint32_t bg_top, uint32_t bg_bottom,
                      j3dg_pixel* pixels_out, uint32_t* rgba_out);
int j3dg_frame_wait(j3dg_ctx* ctx);
/* Dirty-rectangle readback (off by default).  When on, the HOST output buffers of j3dg_render_frame / j3dg_frame_submit
 * are assumed to be persistent per-canvas buffers that nobody else writes between frames (like view::_pixels and
 * canvas::im in j3d): the first frame a buffer receives is copied whole, afterwards only the bounding rectangle of the
 * hit pixels of the current frame united with the rectangle that buffer received last time crosses PCIe — outside it the
 * buffer already holds the identical miss records (canvas.cpp:859-866) / background pixels.  The host buffers end up
 * byte-identical to a full copy.  Frames with point clouds or screen sharding are always copied whole.
 * With pipelined frames the copies of frame k are enqueued by j3dg_frame_submit(k + 1) or j3dg_frame_wait(k). */
int j3dg_ctx_set_dirty_rect(j3dg_ctx* ctx, int enabled);
/* Device->host bytes the frame entry points (j3dg_render_frame / j3dg_frame_submit) have copied since the last reset. */
int j3dg_ctx_readback_bytes(j3dg_ctx* ctx, uint64_t* bytes, int reset);
/* Upload a matcap once and reuse it (frames then pass matcap == NULL). */
int j3dg_ctx_set_matcap(j3dg_ctx* ctx, const uint32_t* matcap, uint32_t mw, uint32_t mh, uint32_t mstride, uint32_t cavity_clr);

/* ---- picking on the device: the consumers of the pixel buffer that j3d runs on the
 *      host copy — canvas::get_pixel (j3d/canvas.cpp:141-153), view::get_id (j3d/view.cpp:483-492),
 *      view::get_world_position (view.cpp:439-469), view::get_index -> get_closest_vertex (view.cpp:471-481,
 *      j3d/pixel.cpp:6-33) and the pivot pick of canvas::do_mouse (canvas.cpp:157-179) — answered from the
 *      canvas that is still resident in HBM, so an interactive host reads back RGBA only (4 B/px instead
 *      of 36 B/px) plus 64 bytes per query.  Outside the canvas or on a miss pixel: db_id = 0,
 *      closest_vertex = 0xFFFFFFFF, world_pos = pivot = NaN (the reference's invalid_vertex / (uint32_t)-1 / 0).
 *      For a point-cloud pixel world_pos = cloud cs * point and closest_vertex = the point index. ---------- */
typedef struct j3dg_pick_result {
  j3dg_pixel pixel;        /* the record at (x, y) */
  float world_pos[3];      /* view::get_world_position */
  uint32_t closest_vertex; /* view::get_index */
  float pivot[3];          /* scene::pivot after a click on (x, y): ray origin + depth * ray dir */
  uint32_t db_id;          /* view::get_id */
} j3dg_pick_result;        /* 64 bytes */
/* pixels: DEVICE buffer to read, or NULL = the canvas of the last j3dg_render_frame / j3dg_frame_submit /
 * host-destination j3dg_cast of this context (it must have view->width x view->height pixels).
 * xy: n x {x, y} int32 (host or device); out: n records (host or device).  Synchronous on return. */
int j3dg_pick(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nr_of_meshes,
              j3dg_cloud* const* clouds, uint32_t nr_of_clouds, const j3dg_view* view,
              const j3dg_pixel* pixels, uint32_t pixel_stride, const int32_t* xy, uint32_t n, j3dg_pick_result* out);

/* ---- all hits along a ray: qbvh::find_all_triangles (jtk/qbvh.h:1854-2000), the query
 *      j3d's voxel export is built on.  rays: n x 8 floats as in j3dg_mesh_find_closest; a triangle is
 *      reported when the reference's Woop test accepts it with t_near < t < t_far (the interval never shrinks).
 *      Results in CSR form: offsets has n + 1 entries (offsets[n] = *total = number of hits); hits are
 *      total x 4 floats {u, v, distance, 0} and triangle_ids total x uint32, grouped per ray, in traversal
 *      order within a ray (the reference's order is an artefact of its own tree as well).
 *      hits / triangle_ids NULL: only offsets and *total are produced (size query).  If capacity < *total the
 *      call fails with J3DG_EINVAL after writing offsets and *total.  Buffers all host or all device. ----- */
int j3dg_mesh_find_all(j3dg_mesh* mesh, const float* rays, uint32_t n, uint32_t* offsets, float* hits,
                       uint32_t* triangle_ids, uint32_t capacity, uint32_t* total);

/* ---- voxel export: the grid-filling loop of _write_vox (j3d/vox.cpp:270-377).  dims = the reference's
 *      (max_dim scaled by the bbox extents, vox.cpp:274-289); three axis-aligned ray grids (one ray per
 *      voxel column, through the column centre, direction +x, +y, +2z); every hit writes the palette index
 *      color_to_index (vox.cpp:154-172) of the texture / vertex colour / white at the hit point into the
 *      voxel containing it: data[x + (y + z * dims[1]) * dims[0]], 0 = empty.  Where several hits with
 *      different colours fall into one voxel the reference keeps whichever of its worker threads wrote last;
 *      here the largest palette index wins (deterministic).  data NULL: only dims_out is produced.
 *      data: host or device, capacity in bytes >= dims[0] * dims[1] * dims[2]. ------------------------------ */
int j3dg_mesh_voxel_dims(const j3dg_mesh* mesh, uint32_t max_dim, uint32_t dims_out[3]);
int j3dg_mesh_voxelize(j3dg_mesh* mesh, uint32_t max_dim, uint32_t dims_out[3], uint8_t* data, size_t capacity);

/* ---- ingest: the step before the path — what view::load_mesh_from_file / load_pc_from_file do
 *      between the file and add_object (j3d/view.cpp:206-270).
 *
 *      Binary PLY: jtk::read_ply (jtk/ply.h:577-690, reached through j3d/io.cpp:733).  file_bytes: the whole file in
 *      HOST memory; the header is parsed on the host, the element data is uploaded once and decoded by two kernels
 *      into device arrays: vertices nv x 3 float, normals nv x 3 float (if nx, ny, nz exist), colours nv x uint32
 *      0xAABBGGRR (if a red / r / diffuse_red channel exists; missing channels 0xff), triangles nf x 3 uint32 (the
 *      first three indices of every face, like the reference), uv nf x 6 float (texcoord list, cut / zero-padded).
 *      Any scalar type per property, little or big endian.  ASCII files are rejected (J3DG_EINVAL): parsing text
 *      stays on the host. ------------------------------------------------------------------------------------- */
typedef struct j3dg_ply j3dg_ply;
typedef struct j3dg_ply_info {
  uint32_t nr_of_vertices, nr_of_faces;
  uint32_t has_normals, has_colors, has_uv;
  uint32_t format;              /* 1 binary_little_endian, 2 binary_big_endian */
  uint64_t header_bytes, file_bytes;
  float upload_ms, decode_ms;   /* device times (CUDA events) of the byte upload and of the decode kernels */
} j3dg_ply_info;
int j3dg_ply_decode(j3dg_ctx* ctx, const void* file_bytes, size_t nr_of_bytes, j3dg_ply** out);
void j3dg_ply_destroy(j3dg_ply* ply);
int j3dg_ply_info_get(const j3dg_ply* ply, j3dg_ply_info* out);
/* Device pointers of the decoded arrays (NULL for what the file does not have); owned by the handle. */
int j3dg_ply_arrays(const j3dg_ply* ply, const float** vertices, const float** normals, const uint32_t** colors,
                    const uint32_t** triangles, const float** uv);
/* Copies one decoded array to the host: which = 0 vertices, 1 normals, 2 colours, 3 triangles, 4 uv. */
int j3dg_ply_copy(const j3dg_ply* ply, int which, void* host_out, size_t capacity_bytes);
/* read_from_file(mesh&, "x.ply") + add_object (j3d/mesh.cpp:104-116, 179-183; j3d/scene.cpp:8-25): colours become the
 * float triples of convert_vertex_colors, texture coordinates without a texture get make_dummy_texture's checkerboard. */
int j3dg_mesh_create_from_ply(j3dg_ctx* ctx, const j3dg_ply* ply, const float* cs, uint32_t db_id, j3dg_mesh** out);
/* read_from_file(pc&, "x.ply") + add_object (j3d/pc.cpp:51-60). */
int j3dg_cloud_create_from_ply(j3dg_ctx* ctx, const j3dg_ply* ply, const float* cs, uint32_t db_id, j3dg_cloud** out);

/* Point-cloud normal estimation: estimate_normals (j3d/pc.cpp:256-334).  For every point the k nearest points (the
 * point itself included, jtk::point_tree::find_k_nearest) are found on the device through a uniform grid, a plane is
 * fitted to them (jtk::fit_plane: the eigenvector of the smallest eigenvalue of the 3 x 3 scatter matrix) — both
 * data parallel — and the normals are then oriented consistently by the reference's propagation over the neighbour
 * graph (a priority queue on |n_i . n_j|, serial by nature: it runs on the host on the downloaded k-NN lists).
 * The cloud's normals are replaced (the splat shades with them); normals_out: nullable, n x 3 floats, host or device.
 * The sign of a whole connected component is arbitrary in the reference too (it follows its SVD). 1 <= k <= 64. */
int j3dg_cloud_estimate_normals(j3dg_cloud* cloud, uint32_t k, float* normals_out);
/* The same without the orientation pass, plus the neighbour lists: knn_out nullable, n x k uint32 in ascending
 * distance order (host or device). */
int j3dg_cloud_knn_normals(j3dg_cloud* cloud, uint32_t k, float* normals_out, uint32_t* knn_out);

/* Traversal statistics of the device BVH for the given view (a counting pass, not the
 * timed kernel): mean wide-node visits and triangle tests per primary ray.  SURVEY §8d. */
int j3dg_cast_stats(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nr_of_meshes, const j3dg_view* view,
                    double* nodes_per_ray, double* tris_per_ray);

/* Diagnostic companion of j3dg_cast_stats: per-pixel node visits / triangle tests of the
 * counting pass, width*height uint32 each (host pointers, row-major, stride = width). */
int j3dg_cast_cost_image(j3dg_ctx* ctx, j3dg_mesh* const* meshes, uint32_t nr_of_meshes, const j3dg_view* view,
                         uint32_t* nodes_out, uint32_t* tris_out);

#ifdef __cplusplus
}
#endif
#endif /* J3DG_H */
