// Error path and back-pressure of the B200-native TSDFSystem (include/tsdf_b200/tsdf_system.hpp), with the same
// stand-in types the drop-in driver uses.  A pool of 64 blocks cannot hold the first frame: the worker's
// TSDF_E_POOL_EXHAUSTED must surface from the next front-end call as tsdf_b200::Error, never be swallowed; a bounded
// backlog must make Integrate wait instead of queueing without limit.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "modules/tsdf_module.h"  // the shadowing header: cv::Mat, SE3, CameraIntrinsics, tsdf_b200::TSDFSystemT

struct Header { int n_frames, w, h, pad; float voxel, trunc, max_depth, K[4], bbox[6]; };
using System = tsdf_b200::TSDFSystemT<cv::Mat, CameraIntrinsics<float>, SE3<float>, VoxelSpatialTSDF>;

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s frames.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("frames"); return 2; }
  Header H;
  if (fread(&H, sizeof(H), 1, f) != 1) return 2;
  const size_t n = (size_t)H.w * H.h;
  std::vector<float> pose(7), depth(n), ht(n), lt(n);
  std::vector<unsigned char> rgb(3 * n);
  if (fread(pose.data(), 4, 7, f) != 7 || fread(rgb.data(), 1, 3 * n, f) != 3 * n || fread(depth.data(), 4, n, f) != n ||
      fread(ht.data(), 4, n, f) != n || fread(lt.data(), 4, n, f) != n) return 2;
  fclose(f);
  const CameraIntrinsics<float> K(H.K[0], H.K[1], H.K[2], H.K[3]);
  const SE3<float> T(Eigen::Quaternionf(pose[3], pose[0], pose[1], pose[2]), Eigen::Vector3f(pose[4], pose[5], pose[6]));
  const cv::Mat m_rgb(H.h, H.w, CV_8UC3, rgb.data()), m_d(H.h, H.w, CV_32FC1, depth.data()), m_ht(H.h, H.w, CV_32FC1, ht.data()),
      m_lt(H.h, H.w, CV_32FC1, lt.data());

  // 1. pool exhaustion on the worker -> rethrown by Flush()
  {
    tsdf_config cfg; tsdf_default_config(&cfg);
    cfg.pool_blocks = 64; cfg.table_slots = 128; cfg.max_image_pixels = H.w * H.h;
    System sys(H.voxel, H.trunc, H.max_depth, K, SE3<float>::Identity(), 0, &cfg);
    sys.Integrate(T, m_rgb, m_d, m_ht, m_lt);
    bool thrown = false;
    try { sys.Flush(); sys.Flush(); } catch (const tsdf_b200::Error& e) {
      thrown = e.code == TSDF_E_POOL_EXHAUSTED;
      printf("caught: %s (code %d)\n", e.what(), e.code);
    }
    if (!thrown) { fprintf(stderr, "pool exhaustion was not reported\n"); return 1; }
  }
  // 2. wrong image type -> rethrown, and the system keeps working afterwards
  {
    System sys(H.voxel, H.trunc, H.max_depth, K);
    sys.Integrate(T, m_d /* CV_32FC1 where CV_8UC3 is required */, m_d, m_ht, m_lt);
    bool thrown = false;
    try { sys.Flush(); } catch (const tsdf_b200::Error& e) { thrown = e.code == TSDF_E_INVALID; }
    if (!thrown) { fprintf(stderr, "bad image type was not reported\n"); return 1; }
    sys.Integrate(T, m_rgb, m_d, m_ht, m_lt);
    sys.Flush();
    if (sys.FramesIntegrated() != 1 || sys.Grid().NumActiveBlock() <= 0) { fprintf(stderr, "system did not recover\n"); return 1; }
  }
  // 3. bounded backlog: with max_backlog = 1 the queue never holds more than one waiting frame
  {
    System sys(H.voxel, H.trunc, H.max_depth, K, SE3<float>::Identity(), 1);
    size_t worst = 0;
    for (int i = 0; i < 8; ++i) {
      sys.Integrate(T, m_rgb, m_d, m_ht, m_lt);
      worst = std::max(worst, sys.Backlog());
    }
    sys.Flush();
    if (worst > 2 || sys.Backlog() != 0 || sys.FramesIntegrated() != 8) { fprintf(stderr, "backlog %zu frames %lld\n", worst, (long long)sys.FramesIntegrated()); return 1; }
  }
  printf("native system error paths ok\n");
  return 0;
}
