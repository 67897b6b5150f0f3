// Drop-in replacement for the reference header utils/tsdf/voxel_tsdf.cuh (yuzhou42/disinfect-slam).
//
// Put  -I<this repo>/include/tsdf_b200/compat  BEFORE the reference root on the include path and link
// libtsdf_b200.so instead of compiling utils/tsdf/{voxel_tsdf,voxel_hash,voxel_mem}.cu: callers such
// as modules/tsdf_module.cc, examples/tsdf/offline.cc and disinfect_slam/disinfect_slam.cc compile
// UNCHANGED (tests/cpp/build_dropin.sh compiles modules/tsdf_module.cc this way and runs it on B200).
// Same global-namespace names and signatures as the reference (voxel_tsdf.cuh:12-88); the GPU state
// behind them is the B200 engine (C ABI in include/tsdf_b200.h).  Still taken from the reference tree:
// the POD voxel types (utils/tsdf/voxel_types.{cuh,cu}) and utils/cuda/{camera,lie_group}.cuh.
#pragma once
#include <cuda_runtime.h>

#include <opencv2/opencv.hpp>
#include <vector>

#include "tsdf_b200/voxel_tsdf.hpp"
#include "utils/cuda/camera.cuh"
#include "utils/cuda/lie_group.cuh"
#include "utils/gl/image.h"
#include "utils/tsdf/voxel_types.cuh"

template <typename T>
struct BoundingCube {  // member order of the reference (voxel_tsdf.cuh:12-19): aggregate-initialised by callers
  T xmin, xmax, ymin, ymax, zmin, zmax;
  template <typename Tout = T>
  BoundingCube<Tout> Scale(T s) const {
    return BoundingCube<Tout>({static_cast<Tout>(xmin * s), static_cast<Tout>(xmax * s), static_cast<Tout>(ymin * s),
                               static_cast<Tout>(ymax * s), static_cast<Tout>(zmin * s), static_cast<Tout>(zmax * s)});
  }
};

class TSDFGrid {
 public:
  TSDFGrid(float voxel_size, float truncation) : impl_(voxel_size, truncation) {}

  void Integrate(const cv::Mat& img_rgb, const cv::Mat& img_depth, const cv::Mat& img_ht, const cv::Mat& img_lt,
                 float max_depth, const CameraIntrinsics<float>& intrinsics, const SE3<float>& cam_T_world) {
    impl_.Integrate(img_rgb, img_depth, img_ht, img_lt, max_depth, intrinsics, cam_T_world);
  }
  void RayCast(float max_depth, const CameraParams& virtual_cam, const SE3<float>& cam_T_world,
               GLImage8UC4* tsdf_rgba = NULL, GLImage8UC4* tsdf_normal = NULL) {
    impl_.RayCast(max_depth, virtual_cam, cam_T_world, tsdf_rgba, tsdf_normal);
  }
  std::vector<VoxelSpatialTSDF> GatherValid() { return impl_.GatherValid<VoxelSpatialTSDF>(); }
  std::vector<VoxelSpatialTSDF> GatherVoxels(const BoundingCube<float>& volumn) {
    return impl_.GatherVoxels<VoxelSpatialTSDF>(volumn);
  }
  tsdf_b200::TSDFGrid& engine() { return impl_; }

 private:
  tsdf_b200::TSDFGrid impl_;
};
