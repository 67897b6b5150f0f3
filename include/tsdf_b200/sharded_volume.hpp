// sharded_volume.hpp -- header-only C++17 RAII wrapper over the multi-GPU data plane (include/tsdf_b200_mgpu.h,
// libtsdf_b200_mgpu.so): ONE TSDF volume whose voxel blocks are sharded over the GPUs of a node.  One ShardedVolume per
// rank -- a process per GPU, or a host thread per GPU as below.  The reference has no counterpart (single GPU,
// utils/tsdf/voxel_tsdf.cuh:103-104); the methods mirror TSDFGrid's.
//
//   unsigned char id[TSDF_MGPU_ID_BYTES];
//   tsdf_b200::ShardedVolume::UniqueId(id);                       // once, handed to every rank
//   std::vector<std::thread> ranks;
//   for (int r = 0; r < n_gpus; ++r) ranks.emplace_back([&, r] {
//     tsdf_b200::ShardedVolume vol(0.02f, 0.12f, r, n_gpus, id);  // collective
//     for (const Frame& f : stream) {                             // same calls, same cameras on every rank
//       vol.Integrate(/*root=*/0, r == 0 ? f.rgb : nullptr, r == 0 ? f.depth : nullptr, r == 0 ? f.ht : nullptr,
//                     r == 0 ? f.lt : nullptr, w, h, 4.f, K, f.q, f.t);
//       vol.RayCast(4.f, w, h, K, f.q, f.t);                      // exact; images on every rank's GPU
//     }
//     vol.Synchronize();
//   });
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../tsdf_b200_mgpu.h"

namespace tsdf_b200 {

class ShardedVolume {
 public:
  struct Failure : std::runtime_error {
    int code;
    Failure(int c, const char* msg) : std::runtime_error(std::string("tsdf_b200 mgpu: ") + msg), code(c) {}
  };
  static void UniqueId(void* id128) { ok(tsdf_mgpu_unique_id(id128)); }

  // cfg: this rank's engine (device, pool_blocks / table_slots of the shard, max_image_pixels, shard granularity in flags);
  // nullptr = defaults with device = rank
  ShardedVolume(float voxel_size, float truncation, int rank, int world, const void* id128, const tsdf_config* cfg = nullptr) {
    tsdf_config c;
    if (cfg) c = *cfg;
    else { tsdf_default_config(&c); c.device = rank; c.flags = 2; }
    ok(tsdf_mgpu_create(voxel_size, truncation, &c, rank, world, id128, &h_));
  }
  ~ShardedVolume() { tsdf_mgpu_destroy(h_); }
  ShardedVolume(const ShardedVolume&) = delete;
  ShardedVolume& operator=(const ShardedVolume&) = delete;

  // TSDFGrid::Integrate; plane pointers are read on `root` only (host memory, or device memory with planes_on_device)
  void Integrate(int root, const void* rgb, const void* depth, const void* ht, const void* lt, int w, int h, float max_depth,
                 const float K[4], const float q_xyzw[4], const float t_xyz[3], bool planes_on_device = false) {
    const int rc = tsdf_mgpu_integrate(h_, root, planes_on_device ? 1 : 0, rgb, depth, ht, lt, w, h, max_depth, K, q_xyzw, t_xyz);
    if (rc != TSDF_OK) throw Failure(rc, tsdf_mgpu_last_error());
  }
  // TSDFGrid::RayCast, exact; the assembled images stay on this rank's GPU (valid until the next RayCast)
  struct Images { const void *rgba, *normal, *hit_depth; };
  Images RayCast(float max_depth, int w, int h, const float K[4], const float q_xyzw[4], const float t_xyz[3]) {
    Images im{};
    ok(tsdf_mgpu_raycast(h_, max_depth, w, h, K, q_xyzw, t_xyz, &im.rgba, &im.normal, &im.hit_depth));
    return im;
  }
  void FetchImages(uint8_t* rgba, uint8_t* normal, float* hit_depth) { ok(tsdf_mgpu_fetch_images(h_, rgba, normal, hit_depth)); }
  // GatherValid (bbox == nullptr) / GatherVoxels over all shards: records {x, y, z, tsdf} on `root`, empty elsewhere
  std::vector<float> Gather(int root, int rank, const float* bbox = nullptr) {
    int64_t n = 0;
    ok(tsdf_mgpu_gather(h_, root, bbox, nullptr, 0, &n));
    std::vector<float> out(rank == root ? static_cast<size_t>(n) * 4 : 0);
    ok(tsdf_mgpu_gather(h_, root, bbox, out.empty() ? nullptr : out.data(), n, &n));
    return out;
  }
  int64_t NumActiveBlock() { int64_t n = 0; ok(tsdf_mgpu_counters(h_, nullptr, nullptr, &n)); return n; }
  void Synchronize() { ok(tsdf_mgpu_synchronize(h_)); }
  tsdf_mgpu_handle handle() const { return h_; }
  tsdf_handle engine() const { return tsdf_mgpu_engine(h_); }

 private:
  static void ok(int rc) { if (rc != TSDF_OK) throw Failure(rc, tsdf_mgpu_last_error()); }
  tsdf_mgpu_handle h_ = nullptr;
};

}  // namespace tsdf_b200
