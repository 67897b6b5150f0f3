// Drop-in replacement for the reference header modules/tsdf_module.h (yuzhou42/disinfect-slam).
//
// Put  -I<this repo>/include/tsdf_b200/compat_system -I<this repo>/include/tsdf_b200/compat  BEFORE the reference
// root on the include path, link libtsdf_b200.so and drop modules/tsdf_module.cc from the build: callers of
// TSDFSystem (disinfect_slam/disinfect_slam.cc:66,113, examples/tsdf/online*.cc, modules/renderer_module.cc)
// compile unchanged and get the pipelined B200-native front end of include/tsdf_b200/tsdf_system.hpp.
// (With only .../compat on the path the reference's OWN TSDFSystem runs on the engine instead; both are built
// and run by tests/test_gpu_dropin.py.)
#pragma once
#include "tsdf_b200/tsdf_system.hpp"
#include "utils/cuda/camera.cuh"
#include "utils/cuda/lie_group.cuh"
#include "utils/gl/image.h"
#include "utils/tsdf/voxel_tsdf.cuh"

#define TSDF_B200_NATIVE_SYSTEM 1

// kept for source compatibility (tsdf_module.h:16-31); the native system queues the same five members
struct TSDFSystemInput {
  SE3<float> cam_T_world;
  cv::Mat img_rgb, img_depth, img_ht, img_lt;
  TSDFSystemInput(const SE3<float>& cam_T_world, const cv::Mat& img_rgb, const cv::Mat& img_depth, const cv::Mat& img_ht,
                  const cv::Mat& img_lt)
      : cam_T_world(cam_T_world), img_rgb(img_rgb), img_depth(img_depth), img_ht(img_ht), img_lt(img_lt) {}
};

class TSDFSystem : public tsdf_b200::TSDFSystemT<cv::Mat, CameraIntrinsics<float>, SE3<float>, VoxelSpatialTSDF> {
  using Base = tsdf_b200::TSDFSystemT<cv::Mat, CameraIntrinsics<float>, SE3<float>, VoxelSpatialTSDF>;

 public:
  TSDFSystem(float voxel_size, float truncation, float max_depth, const CameraIntrinsics<float>& intrinsics,
             const SE3<float>& extrinsics = SE3<float>::Identity())
      : Base(voxel_size, truncation, max_depth, intrinsics, extrinsics) {}
  std::vector<VoxelSpatialTSDF> Query(const BoundingCube<float>& volumn) { return Base::Query(volumn); }
  void Render(const CameraParams& virtual_cam, const SE3<float> cam_T_world, GLImage8UC4* img_normal) {
    Base::Render(virtual_cam, cam_T_world, img_normal);
  }
};
