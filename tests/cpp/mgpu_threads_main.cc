// Pure C++ driver of the multi-GPU data plane (include/tsdf_b200_mgpu.h): one host thread per GPU in ONE process, NCCL
// linked directly, no Python and no torch anywhere in the process.  Reads a frame file written by
// tests/test_gpu_mgpu.py, integrates every frame into a volume sharded over `world` GPUs (rank 0 holds the host
// frames), renders the last camera exactly (peer memory over NVLink) and gathers the whole volume on rank 0.
// Output file: int64 n_voxels, n_voxels x {x, y, z, tsdf}, then rgba / normal (H x W x 4) and hit depth (H x W) of the
// last view, then the all-reduced counters of the last frame and the active-block total.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "tsdf_b200_mgpu.h"

struct Header { int n_frames, w, h, world; float voxel, trunc, max_depth, K[4], bbox[6]; };

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s frames.bin out.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("frames"); return 2; }
  Header H;
  if (fread(&H, sizeof(H), 1, f) != 1) return 2;
  const size_t n = (size_t)H.w * H.h;
  std::vector<std::vector<unsigned char>> rgb(H.n_frames, std::vector<unsigned char>(3 * n));
  std::vector<std::vector<float>> depth(H.n_frames, std::vector<float>(n)), ht = depth, lt = depth;
  std::vector<std::vector<float>> pose(H.n_frames, std::vector<float>(7));
  for (int i = 0; i < H.n_frames; ++i) {
    if (fread(pose[i].data(), 4, 7, f) != 7 || fread(rgb[i].data(), 1, 3 * n, f) != 3 * n || fread(depth[i].data(), 4, n, f) != n ||
        fread(ht[i].data(), 4, n, f) != n || fread(lt[i].data(), 4, n, f) != n) return 2;
  }
  fclose(f);
  const int world = H.world;
  unsigned char id[TSDF_MGPU_ID_BYTES];
  if (tsdf_mgpu_unique_id(id) != TSDF_OK) { fprintf(stderr, "unique id: %s\n", tsdf_mgpu_last_error()); return 1; }
  std::vector<int> status(world, 0);
  std::vector<float> gathered;
  std::vector<unsigned char> rgba(4 * n), normal(4 * n);
  std::vector<float> hit(n);
  long long n_vox = 0, n_active = 0;
  tsdf_counters last_sum{};
  auto rank_main = [&](int rank) {
#define OK(call) do { if ((call) != TSDF_OK) { fprintf(stderr, "rank %d: %s: %s\n", rank, #call, tsdf_mgpu_last_error()); status[rank] = 1; return; } } while (0)
    tsdf_config cfg;
    tsdf_default_config(&cfg);
    cfg.device = rank; cfg.pool_blocks = 1 << 16; cfg.table_slots = 1 << 19; cfg.max_image_pixels = (int)n; cfg.flags = 2;
    tsdf_mgpu_handle m = nullptr;
    OK(tsdf_mgpu_create(H.voxel, H.trunc, &cfg, rank, world, id, &m));
    std::vector<tsdf_mgpu_frame> frames(H.n_frames);
    for (int i = 0; i < H.n_frames; ++i) {
      tsdf_mgpu_frame& fr = frames[i];
      memset(&fr, 0, sizeof(fr));
      if (rank == 0) { fr.rgb = rgb[i].data(); fr.depth = depth[i].data(); fr.ht = ht[i].data(); fr.lt = lt[i].data(); }
      memcpy(fr.q_xyzw, pose[i].data(), 16); memcpy(fr.t_xyz, pose[i].data() + 4, 12);
    }
    // the whole stream in one call: Integrate + exact RayCast per frame, pageable host planes on rank 0
    OK(tsdf_mgpu_run_sequence(m, 0, 0, frames.data(), H.n_frames, 0, H.n_frames, H.w, H.h, H.max_depth, H.K, 1));
    OK(tsdf_mgpu_synchronize(m));
    if (rank == 0) OK(tsdf_mgpu_fetch_images(m, rgba.data(), normal.data(), hit.data()));
    long long total = 0, act = 0;
    tsdf_counters ls{};
    OK(tsdf_mgpu_counters(m, &ls, nullptr, (int64_t*)&act));
    OK(tsdf_mgpu_gather(m, 0, nullptr, nullptr, 0, (int64_t*)&total));
    if (rank == 0) gathered.resize((size_t)total * 4);
    OK(tsdf_mgpu_gather(m, 0, nullptr, rank == 0 ? gathered.data() : nullptr, total, (int64_t*)&total));
    if (rank == 0) { n_vox = total; n_active = act; last_sum = ls; }
    OK(tsdf_mgpu_destroy(m));
#undef OK
  };
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(rank_main, r);
  for (auto& t : th) t.join();
  for (int r = 0; r < world; ++r) if (status[r]) return 1;
  FILE* g = fopen(argv[2], "wb");
  fwrite(&n_vox, 8, 1, g);
  fwrite(gathered.data(), 16, (size_t)n_vox, g);
  fwrite(rgba.data(), 1, 4 * n, g); fwrite(normal.data(), 1, 4 * n, g); fwrite(hit.data(), 4, n, g);
  fwrite(&last_sum, sizeof(last_sum), 1, g);
  fwrite(&n_active, 8, 1, g);
  fclose(g);
  printf("mgpu threads ok: world %d, %d frames, %lld voxels gathered, %lld active blocks\n", world, H.n_frames, n_vox, n_active);
  return 0;
}
