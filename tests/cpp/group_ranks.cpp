// tests/cpp/group_ranks.cpp — the multi-GPU part of the C ABI (j3dg_group_*, j3dg_frames_*; include/j3dg.h) driven by a
// plain C++ host, one PROCESS per GPU, no Python and no torch: the parent forks `world` ranks; rank 0 creates the NCCL
// id (handed over through a file), builds a vertex-coloured mesh and broadcasts it; every rank renders its share of an
// orbit sweep straight into rank 0's frame buffer over NVLink peer memory; rank 0 consumes every frame between arrival
// and release and compares it with its own render of the same pose.  Then ONE frame is rendered band-sharded by all
// ranks into one shared frame and compared with the unsharded frame.  TEST CODE (built by the tests with g++).
//
//   group_ranks <world> [width height frames]      exit 0 = all comparisons equal, 77 = fewer GPUs than ranks
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "j3dg_host.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    const int rc__ = (call);                                                                          \
    if (rc__ != J3DG_OK) { fprintf(stderr, "rank %d: %s = %d: %s\n", rank, #call, rc__, j3dg_last_error(ctx)); return 1; } \
  } while (0)

static void bumpy_sphere(int nlat, int nlon, std::vector<float>& v, std::vector<uint32_t>& t, std::vector<float>& c) {
  for (int i = 0; i <= nlat; ++i)
    for (int j = 0; j < nlon; ++j) {
      const double th = M_PI * i / nlat, ph = 2 * M_PI * j / nlon;
      const double r = 1.0 + 0.05 * std::sin(7 * th) * std::cos(5 * ph);
      v.push_back((float)(r * std::sin(th) * std::cos(ph))); v.push_back((float)(r * std::cos(th))); v.push_back((float)(r * std::sin(th) * std::sin(ph)));
      c.push_back((float)(0.5 + 0.5 * std::sin(3 * ph))); c.push_back((float)(i / (double)nlat)); c.push_back((float)(0.5 + 0.5 * std::cos(2 * th)));
    }
  for (int i = 0; i < nlat; ++i)
    for (int j = 0; j < nlon; ++j) {
      const uint32_t a = i * nlon + j, b = i * nlon + (j + 1) % nlon, d = (i + 1) * nlon + j, e = (i + 1) * nlon + (j + 1) % nlon;
      t.push_back(a); t.push_back(d); t.push_back(b);
      t.push_back(b); t.push_back(d); t.push_back(e);
    }
}

static int run_rank(int rank, int world, uint32_t w, uint32_t h, int frames, const std::string& id_path) {
  j3dg_ctx* ctx = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < world) return 77;
  CHECK(j3dg_ctx_create(rank, &ctx));
  cudaStream_t stream;
  cudaSetDevice(rank);
  cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
  CHECK(j3dg_ctx_set_stream(ctx, stream));

  // ---- the NCCL id: rank 0 makes it, the others read it from the file ----
  unsigned char id[J3DG_GROUP_ID_BYTES];
  if (rank == 0) {
    CHECK(j3dg_group_unique_id(id));
    FILE* f = fopen((id_path + ".tmp").c_str(), "wb");
    fwrite(id, 1, sizeof(id), f);
    fclose(f);
    rename((id_path + ".tmp").c_str(), id_path.c_str());
  } else {
    FILE* f = nullptr;
    for (int tries = 0; tries < 600 && !(f = fopen(id_path.c_str(), "rb")); ++tries) usleep(100000);
    if (!f || fread(id, 1, sizeof(id), f) != sizeof(id)) { fprintf(stderr, "rank %d: no NCCL id\n", rank); return 1; }
    fclose(f);
  }
  j3dg_group* group = nullptr;
  CHECK(j3dg_group_create(ctx, rank, world, id, &group));

  // ---- rank 0 loads + builds, everybody receives ----
  std::vector<float> verts, colors;
  std::vector<uint32_t> tris;
  j3dg_mesh* mesh = nullptr;
  float bb[6] = {0, 0, 0, 0, 0, 0};
  if (rank == 0) {
    bumpy_sphere(300, 400, verts, tris, colors);
    CHECK(j3dg_mesh_create(ctx, verts.data(), (uint32_t)(verts.size() / 3), tris.data(), (uint32_t)(tris.size() / 3), colors.data(), nullptr, nullptr, 0, 0, 0,
                           nullptr, 0x20000000u, &mesh));
  }
  CHECK(j3dg_group_broadcast_mesh(group, 0, &mesh));
  j3dg_mesh_info info;
  CHECK(j3dg_mesh_info_get(mesh, &info));
  memcpy(bb, info.bbox_min, 12); memcpy(bb + 3, info.bbox_max, 12);
  if (info.nr_of_triangles != 300u * 400u * 2u || info.nr_of_nodes == 0) { fprintf(stderr, "rank %d: received mesh is wrong\n", rank); return 1; }

  j3dg_view v0;
  memset(&v0, 0, sizeof(v0));
  v0.width = w; v0.height = h; v0.flags = J3DG_DEFAULT_FLAGS;
  j3dgh_make_projection(w, h, &v0.near_plane, v0.projection, v0.projection_inv);
  j3dgh_unzoom(bb, bb + 3, &v0.diagonal, v0.pivot, v0.cs, v0.cs_inv);
  auto pose = [&](int i, uint32_t flags) { j3dg_view v = v0; v.flags = flags; j3dgh_orbit(v0.cs_inv, v0.pivot, 7.f * i, v.cs, v.cs_inv); return v; };
  j3dg::matcap mc;
  j3dg::make_matcap_red_wax(mc);
  CHECK(j3dg_ctx_set_matcap(ctx, mc.im.data(), mc.w, mc.h, mc.w, mc.cavity_clr));

  const size_t npx = (size_t)w * h;
  j3dg_pixel* d_px = nullptr;
  cudaMalloc((void**)&d_px, npx * sizeof(j3dg_pixel));
  int bad = 0;

  // ---- (1) orbit sweep: frame k * world + r on rank r, every frame lands in rank 0's HBM ----
  {
    j3dg_frames* fr = nullptr;
    CHECK(j3dg_frames_create(group, w, h, 0, 0, &fr));
    std::vector<std::vector<uint32_t>> got(frames);
    for (int k = 0; k < frames; ++k) {
      uint32_t kk = 0;
      uint32_t* target = nullptr;
      CHECK(j3dg_frames_begin(fr, &kk));
      CHECK(j3dg_frames_target(fr, kk, &target));
      const j3dg_view v = pose(k * world + rank, J3DG_DEFAULT_FLAGS);
      CHECK(j3dg_render_frame(ctx, &mesh, 1, nullptr, 0, &v, nullptr, 0, 0, 0, 0, 0xff000000u, 0xff404040u, d_px, target));
      CHECK(j3dg_frames_arrive(fr, kk));
      if (rank == 0) {  // the consumer: a copy to the host, enqueued between arrival and release
        const uint32_t* all = nullptr;
        CHECK(j3dg_frames_view(fr, kk, &all));
        got[k].resize(npx * world);
        cudaMemcpyAsync(got[k].data(), all, npx * world * 4, cudaMemcpyDeviceToHost, stream);
      }
      CHECK(j3dg_frames_release(fr, kk));
    }
    CHECK(j3dg_ctx_synchronize(ctx));
    if (rank == 0) {
      std::vector<uint32_t> want(npx);
      for (int k = 0; k < frames; ++k)
        for (int r = 0; r < world; ++r) {
          const j3dg_view v = pose(k * world + r, J3DG_DEFAULT_FLAGS);
          CHECK(j3dg_render_frame(ctx, &mesh, 1, nullptr, 0, &v, nullptr, 0, 0, 0, 0, 0xff000000u, 0xff404040u, nullptr, want.data()));
          if (memcmp(want.data(), got[k].data() + (size_t)r * npx, npx * 4) != 0) { fprintf(stderr, "frame %d of rank %d differs\n", k, r); ++bad; }
        }
    }
    j3dg_frames_destroy(fr);
  }
  // ---- (2) one frame with shadow rays, 32-row bands round-robin over the ranks, written into ONE shared frame ----
  {
    j3dg_frames* fr = nullptr;
    CHECK(j3dg_frames_create(group, w, h, 0, 1, &fr));
    CHECK(j3dg_ctx_set_screen_shard(ctx, (uint32_t)rank, (uint32_t)world));
    std::vector<uint32_t> got(npx), want(npx);
    const j3dg_view v = pose(3, J3DG_DEFAULT_FLAGS | J3DG_SHADOW);
    for (int k = 0; k < 3; ++k) {
      uint32_t kk = 0;
      uint32_t* target = nullptr;
      CHECK(j3dg_frames_begin(fr, &kk));
      CHECK(j3dg_frames_target(fr, kk, &target));
      CHECK(j3dg_render_frame(ctx, &mesh, 1, nullptr, 0, &v, nullptr, 0, 0, 0, 0, 0xff000000u, 0xff404040u, d_px, target));
      CHECK(j3dg_frames_arrive(fr, kk));
      if (rank == 0) {
        const uint32_t* all = nullptr;
        CHECK(j3dg_frames_view(fr, kk, &all));
        cudaMemcpyAsync(got.data(), all, npx * 4, cudaMemcpyDeviceToHost, stream);
      }
      CHECK(j3dg_frames_release(fr, kk));
    }
    CHECK(j3dg_ctx_synchronize(ctx));
    CHECK(j3dg_ctx_set_screen_shard(ctx, 0, 1));
    if (rank == 0) {
      CHECK(j3dg_render_frame(ctx, &mesh, 1, nullptr, 0, &v, nullptr, 0, 0, 0, 0, 0xff000000u, 0xff404040u, nullptr, want.data()));
      if (memcmp(want.data(), got.data(), npx * 4) != 0) { fprintf(stderr, "the band-sharded frame differs from the unsharded one\n"); ++bad; }
    }
    j3dg_frames_destroy(fr);
  }
  // ---- (3) the sweep again with THREE frames in flight: two more contexts of the same device (own streams, own scratch,
  //          the SAME mesh) render the frames of slots 1 and 2; the hand-over flags are per slot, so the streams never
  //          order each other ----
  {
    constexpr int LANES = 3;
    j3dg_ctx* lane_ctx[LANES] = {ctx, nullptr, nullptr};
    cudaStream_t lane_stream[LANES] = {stream, nullptr, nullptr};
    j3dg_pixel* lane_px[LANES] = {d_px, nullptr, nullptr};
    for (int l = 1; l < LANES; ++l) {
      CHECK(j3dg_ctx_create(rank, &lane_ctx[l]));
      cudaStreamCreateWithFlags(&lane_stream[l], cudaStreamNonBlocking);
      CHECK(j3dg_ctx_set_stream(lane_ctx[l], lane_stream[l]));
      CHECK(j3dg_ctx_set_matcap(lane_ctx[l], mc.im.data(), mc.w, mc.h, mc.w, mc.cavity_clr));
      cudaMalloc((void**)&lane_px[l], npx * sizeof(j3dg_pixel));
    }
    j3dg_frames* fr = nullptr;
    CHECK(j3dg_frames_create_n(group, w, h, 0, 0, LANES, &fr));
    for (int l = 1; l < LANES; ++l) CHECK(j3dg_frames_set_lane(fr, l, lane_ctx[l]));
    const int frames2 = 3 * frames;
    std::vector<std::vector<uint32_t>> got(frames2);
    for (int k = 0; k < frames2; ++k) {
      uint32_t kk = 0;
      uint32_t* target = nullptr;
      CHECK(j3dg_frames_begin(fr, &kk));
      const int ln = (int)(kk % LANES);
      CHECK(j3dg_frames_target(fr, kk, &target));
      const j3dg_view v = pose(k * world + rank, J3DG_DEFAULT_FLAGS);
      if (j3dg_render_frame(lane_ctx[ln], &mesh, 1, nullptr, 0, &v, nullptr, 0, 0, 0, 0, 0xff000000u, 0xff404040u, lane_px[ln], target) != J3DG_OK) {
        fprintf(stderr, "rank %d: lane %d render: %s\n", rank, ln, j3dg_last_error(lane_ctx[ln]));
        return 1;
      }
      CHECK(j3dg_frames_arrive(fr, kk));
      if (rank == 0) {
        const uint32_t* all = nullptr;
        CHECK(j3dg_frames_view(fr, kk, &all));
        got[k].resize(npx * world);
        cudaMemcpyAsync(got[k].data(), all, npx * world * 4, cudaMemcpyDeviceToHost, lane_stream[ln]);
      }
      CHECK(j3dg_frames_release(fr, kk));
    }
    for (int l = 0; l < LANES; ++l)
      if (j3dg_ctx_synchronize(lane_ctx[l]) != J3DG_OK) { fprintf(stderr, "rank %d: lane %d: %s\n", rank, l, j3dg_last_error(lane_ctx[l])); return 1; }
    if (rank == 0) {
      std::vector<uint32_t> want(npx);
      for (int k = 0; k < frames2; ++k)
        for (int r = 0; r < world; ++r) {
          const j3dg_view v = pose(k * world + r, J3DG_DEFAULT_FLAGS);
          CHECK(j3dg_render_frame(ctx, &mesh, 1, nullptr, 0, &v, nullptr, 0, 0, 0, 0, 0xff000000u, 0xff404040u, nullptr, want.data()));
          if (memcmp(want.data(), got[k].data() + (size_t)r * npx, npx * 4) != 0) { fprintf(stderr, "three lanes: frame %d of rank %d differs\n", k, r); ++bad; }
        }
    }
    j3dg_frames_destroy(fr);
    for (int l = 1; l < LANES; ++l) {
      uint32_t st2 = 0;
      j3dg_ctx_status(lane_ctx[l], &st2, 0);
      if (st2) ++bad;
      cudaFree(lane_px[l]);
      j3dg_ctx_destroy(lane_ctx[l]);
      cudaStreamDestroy(lane_stream[l]);
    }
  }
  float worst = (float)bad;
  CHECK(j3dg_group_max_float(group, &worst, 1));
  CHECK(j3dg_group_barrier(group));
  uint32_t status = 0;
  CHECK(j3dg_ctx_status(ctx, &status, 0));
  cudaFree(d_px);
  j3dg_mesh_destroy(mesh);
  j3dg_group_destroy(group);
  j3dg_ctx_destroy(ctx);
  cudaStreamDestroy(stream);
  if (rank == 0) {
    printf("group_ranks: world %d, %d frames %ux%u + 1 sharded frame + %d frames with three in flight: %s\n", world, frames * world, w, h, 3 * frames * world, (worst == 0.f && status == 0) ? "OK" : "FAILED");
    fflush(stdout);  // the ranks leave through _exit
  }
  return (worst == 0.f && status == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 1;
  const uint32_t w = argc > 3 ? (uint32_t)atoi(argv[2]) : 640, h = argc > 3 ? (uint32_t)atoi(argv[3]) : 360;
  const int frames = argc > 4 ? atoi(argv[4]) : 5;
  if (world < 1 || world > 16) return 2;
  char path[64];
  snprintf(path, sizeof(path), "/tmp/j3dg_group_id_%d", (int)getpid());
  unlink(path);
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {  // fork BEFORE anything touches CUDA
    const pid_t p = fork();
    if (p == 0) _exit(run_rank(r, world, w, h, frames, path));
    kids.push_back(p);
  }
  int worst = 0;
  for (pid_t p : kids) {
    int st = 0;
    waitpid(p, &st, 0);
    const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 1;
    if (code == 77 && worst == 0) worst = 77;
    else if (code != 0 && code != 77) worst = 1;
  }
  unlink(path);
  return worst;
}
