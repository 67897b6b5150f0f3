// Drop-in demonstration: the reference's TSDFSystem (modules/tsdf_module.{h,cc}, compiled UNMODIFIED from
// /root/reference by tests/cpp/build_dropin.sh) running on top of the B200 engine through
// include/tsdf_b200/compat/utils/tsdf/voxel_tsdf.cuh.  Reads a frame file written by
// tests/test_gpu_dropin.py, feeds every frame through TSDFSystem::Integrate (worker thread + queue),
// renders one view through TSDFSystem::Render and writes TSDFSystem::Query(bbox) to the output file.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "modules/tsdf_module.h"

struct Header { int n_frames, w, h, pad; float voxel, trunc, max_depth, K[4], bbox[6]; };

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
  try {
    const CameraIntrinsics<float> K(H.K[0], H.K[1], H.K[2], H.K[3]);
    TSDFSystem sys(H.voxel, H.trunc, H.max_depth, K);  // extrinsics = identity
    SE3<float> last = SE3<float>::Identity();
    for (int i = 0; i < H.n_frames; ++i) {
      const float* p = pose[i].data();
      last = SE3<float>(Eigen::Quaternionf(p[3], p[0], p[1], p[2]), Eigen::Vector3f(p[4], p[5], p[6]));
      sys.Integrate(last, cv::Mat(H.h, H.w, CV_8UC3, rgb[i].data()), cv::Mat(H.h, H.w, CV_32FC1, depth[i].data()),
                    cv::Mat(H.h, H.w, CV_32FC1, ht[i].data()), cv::Mat(H.h, H.w, CV_32FC1, lt[i].data()));
    }
    const BoundingCube<float> box = {H.bbox[0], H.bbox[1], H.bbox[2], H.bbox[3], H.bbox[4], H.bbox[5]};
    std::vector<VoxelSpatialTSDF> out, prev;
#ifdef TSDF_B200_NATIVE_SYSTEM
    // the native system: one frame without probability images exercises the cached plane of ones when the frame
    // file asks for it (pad == 1), then Flush() makes the volume current
    if (H.pad == 1) sys.Integrate(last, cv::Mat(H.h, H.w, CV_8UC3, rgb.back().data()), cv::Mat(H.h, H.w, CV_32FC1, depth.back().data()));
    sys.Flush();
    if (sys.Backlog() != 0 || sys.FramesIntegrated() != H.n_frames + (H.pad == 1 ? 1 : 0)) { fprintf(stderr, "flush left a backlog\n"); return 1; }
    out = sys.Query(box);
    const int max_tries = 0;
#else
    // the reference's TSDFSystem has no flush: wait until two consecutive queries agree after the queue had time to drain
    const int max_tries = 100;
#endif
    for (int tries = 0; tries < max_tries; ++tries) {
      std::this_thread::sleep_for(std::chrono::milliseconds(300));
      out = sys.Query(box);
      if (tries > 0 && out.size() == prev.size() && !out.empty()) break;
      prev = out;
    }
    GLImage8UC4 sink;
    sys.Render(CameraParams(K, H.h, H.w), last, &sink);
    FILE* g = fopen(argv[2], "wb");
    const long long cnt = (long long)out.size();
    fwrite(&cnt, sizeof(cnt), 1, g);
    fwrite(out.data(), sizeof(VoxelSpatialTSDF), out.size(), g);
    fclose(g);
    printf("dropin ok: %d frames, %lld voxels queried\n", H.n_frames, cnt);
  } catch (const std::exception& e) {
    fprintf(stderr, "dropin failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
