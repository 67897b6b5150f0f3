// voxel_tsdf.hpp -- header-only C++17 host class over the C ABI of libtsdf_b200.so.
//
// tsdf_b200::TSDFGrid mirrors the reference's TSDFGrid (utils/tsdf/voxel_tsdf.cuh:32-88 of
// yuzhou42/disinfect-slam): same method names, argument order and meaning.  The image, camera and
// pose parameters are templates over the reference's own types, used only through the members
// the reference itself uses, so this header needs neither Eigen, OpenCV nor GL:
//   Mat         .data  .rows  .cols  .total()  .type()            (cv::Mat)
//   Intrinsics  .fx .fy .cx .cy                                   (CameraIntrinsics<float>, camera.cuh:12-17)
//   Pose        .GetR() -> {x(), y(), z(), w()}, .GetT() -> [i]   (SE3<float>, lie_group.cuh:29-31)
//   CamParams   .intrinsics  .img_h  .img_w                       (CameraParams, camera.cuh:54-68)
//   Image       .LoadCuda(const void* device_ptr)                 (GLImage8UC4, utils/gl/image.h:46)
//   Cube        .xmin .xmax .ymin .ymax .zmin .zmax               (BoundingCube<float>, voxel_tsdf.cuh:12-19)
// The drop-in header with the reference's exact global-namespace signatures is
// include/tsdf_b200/compat/utils/tsdf/voxel_tsdf.cuh.
//
// Error behaviour: the reference's methods return void and only assert on bad input; here a failed
// ABI call throws tsdf_b200::Error (std::runtime_error) carrying tsdf_last_error().
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../tsdf_b200.h"

namespace tsdf_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const char* msg) : std::runtime_error(std::string("tsdf_b200: ") + msg), code(c) {}
};
inline void check(int rc) { if (rc != TSDF_OK) throw Error(rc, tsdf_last_error()); }

// OpenCV type codes the reference asserts on (voxel_tsdf.cu:350-351)
constexpr int kCV_8UC3 = 16, kCV_32FC1 = 5;

struct VoxelSpatialTSDF { float position[3]; float tsdf; };  // utils/tsdf/voxel_types.cuh:48-57 (16 bytes)

class TSDFGrid {
 public:
  TSDFGrid(float voxel_size, float truncation, const tsdf_config* cfg = nullptr) { check(tsdf_create(voxel_size, truncation, cfg, &h_)); }
  ~TSDFGrid() { tsdf_destroy(h_); }
  TSDFGrid(const TSDFGrid&) = delete;
  TSDFGrid& operator=(const TSDFGrid&) = delete;

  template <class Mat, class Intrinsics, class Pose>
  void Integrate(const Mat& img_rgb, const Mat& img_depth, const Mat& img_ht, const Mat& img_lt, float max_depth,
                 const Intrinsics& intrinsics, const Pose& cam_T_world) {
    if (img_rgb.type() != kCV_8UC3 || img_depth.type() != kCV_32FC1 || img_rgb.cols != img_depth.cols || img_rgb.rows != img_depth.rows)
      throw Error(TSDF_E_INVALID, "Integrate: rgb must be CV_8UC3, depth CV_32FC1, same size (voxel_tsdf.cu:350-353)");
    if (img_ht.total() != img_depth.total() || img_lt.total() != img_depth.total())
      throw Error(TSDF_E_INVALID, "Integrate: ht / lt must have the size of depth");
    const float K[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy};
    float q[4], t[3];
    unpack(cam_T_world, q, t);
    check(tsdf_integrate(h_, reinterpret_cast<const uint8_t*>(img_rgb.data), reinterpret_cast<const float*>(img_depth.data),
                         reinterpret_cast<const float*>(img_ht.data), reinterpret_cast<const float*>(img_lt.data), img_depth.cols,
                         img_depth.rows, max_depth, K, q, t));
  }

  // The sensor's own formats: CV_16UC1 depth (metres = value / depthmap_factor) and CV_16UC1 probabilities (value / 65535)
  // -- what the reference holds before `convertTo(CV_32FC1, 1. / scale)` (examples/tsdf/offline.cc:72-83).  The conversion
  // runs on the GPU with convertTo's arithmetic; 9 bytes per pixel cross PCIe instead of 15.  ht / lt may be empty Mats
  // (TSDFSystem's default of ones, modules/tsdf_module.cc:28-33).
  template <class Mat, class Intrinsics, class Pose>
  void IntegrateU16(const Mat& img_rgb, const Mat& depth_u16, const Mat& ht_u16, const Mat& lt_u16, float depthmap_factor, float max_depth,
                    const Intrinsics& intrinsics, const Pose& cam_T_world, int flags = 0) {
    constexpr int kCV_16UC1 = 2;
    if (img_rgb.type() != kCV_8UC3 || depth_u16.type() != kCV_16UC1 || img_rgb.cols != depth_u16.cols || img_rgb.rows != depth_u16.rows)
      throw Error(TSDF_E_INVALID, "IntegrateU16: rgb must be CV_8UC3, depth CV_16UC1, same size");
    const bool probs = ht_u16.total() != 0 || lt_u16.total() != 0;
    if (probs && (ht_u16.type() != kCV_16UC1 || lt_u16.type() != kCV_16UC1 || ht_u16.total() != depth_u16.total() || lt_u16.total() != depth_u16.total()))
      throw Error(TSDF_E_INVALID, "IntegrateU16: ht / lt must be CV_16UC1 of the size of depth, or both empty");
    const float K[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy};
    float q[4], t[3];
    unpack(cam_T_world, q, t);
    check(tsdf_integrate_u16(h_, reinterpret_cast<const uint8_t*>(img_rgb.data), reinterpret_cast<const uint16_t*>(depth_u16.data),
                             probs ? reinterpret_cast<const uint16_t*>(ht_u16.data) : nullptr,
                             probs ? reinterpret_cast<const uint16_t*>(lt_u16.data) : nullptr, depth_u16.cols, depth_u16.rows, depthmap_factor,
                             max_depth, K, q, t, flags));
  }

  // Images are sinks with LoadCuda(device pointer); nullptr skips that output like the reference.
  template <class CamParams, class Pose, class ImageA = std::nullptr_t, class ImageB = std::nullptr_t>
  void RayCast(float max_depth, const CamParams& virtual_cam, const Pose& cam_T_world, ImageA* tsdf_rgba = nullptr,
               ImageB* tsdf_normal = nullptr) {
    const float K[4] = {virtual_cam.intrinsics.fx, virtual_cam.intrinsics.fy, virtual_cam.intrinsics.cx, virtual_cam.intrinsics.cy};
    float q[4], t[3];
    unpack(cam_T_world, q, t);
    const void *d_rgba = nullptr, *d_normal = nullptr;
    check(tsdf_raycast_resident(h_, max_depth, virtual_cam.img_w, virtual_cam.img_h, K, q, t, &d_rgba, &d_normal, nullptr));
    check(tsdf_synchronize(h_));  // the sinks copy on their own (default) stream
    load(tsdf_rgba, d_rgba);
    load(tsdf_normal, d_normal);
  }
  // host-memory variant (headless use): rgba / normal HxWx4 bytes, hit_depth HxW floats, each optional
  template <class CamParams, class Pose>
  void RayCastToHost(float max_depth, const CamParams& virtual_cam, const Pose& cam_T_world, uint8_t* rgba, uint8_t* normal,
                     float* hit_depth = nullptr) {
    const float K[4] = {virtual_cam.intrinsics.fx, virtual_cam.intrinsics.fy, virtual_cam.intrinsics.cx, virtual_cam.intrinsics.cy};
    float q[4], t[3];
    unpack(cam_T_world, q, t);
    check(tsdf_raycast(h_, max_depth, virtual_cam.img_w, virtual_cam.img_h, K, q, t, rgba, normal, hit_depth));
  }

  template <class Voxel = VoxelSpatialTSDF>
  std::vector<Voxel> GatherValid() { return gather<Voxel>(nullptr); }
  template <class Voxel = VoxelSpatialTSDF, class Cube>
  std::vector<Voxel> GatherVoxels(const Cube& volumn) {
    const float b[6] = {volumn.xmin, volumn.xmax, volumn.ymin, volumn.ymax, volumn.zmin, volumn.zmax};
    return gather<Voxel>(b);
  }

  // Triangles of the zero level set of the blocks GatherVoxels(volumn) would select, extracted on the GPU
  // (tsdf_extract_mesh): 9 floats per triangle.  Replaces Query + KrisLibrary ExtractMesh (ros_offline.cc:258-318).
  struct Triangle { float v[3][3]; };
  std::vector<Triangle> ExtractMesh() { return mesh(nullptr); }
  template <class Cube>
  std::vector<Triangle> ExtractMesh(const Cube& volumn) {
    const float b[6] = {volumn.xmin, volumn.xmax, volumn.ymin, volumn.ymax, volumn.zmin, volumn.zmax};
    return mesh(b);
  }

  int NumActiveBlock() { int n = 0; check(tsdf_num_active_blocks(h_, &n)); return n; }
  tsdf_counters Counters() { tsdf_counters c; check(tsdf_get_counters(h_, &c)); return c; }
  tsdf_handle handle() const { return h_; }

 private:
  template <class Pose>
  static void unpack(const Pose& T, float q[4], float t[3]) {
    const auto R = T.GetR();
    const auto tr = T.GetT();
    q[0] = R.x(); q[1] = R.y(); q[2] = R.z(); q[3] = R.w();
    t[0] = tr[0]; t[1] = tr[1]; t[2] = tr[2];
  }
  template <class Image>
  static void load(Image* img, const void* dev) { if (img) img->LoadCuda(dev); }
  static void load(std::nullptr_t*, const void*) {}
  template <class Voxel>
  std::vector<Voxel> gather(const float* bbox) {
    static_assert(sizeof(Voxel) == 16, "VoxelSpatialTSDF is {float position[3]; float tsdf;}");
    int64_t n = 0;
    check(bbox ? tsdf_gather_in_bound(h_, bbox, nullptr, 0, &n) : tsdf_gather_valid(h_, nullptr, 0, &n));
    std::vector<Voxel> ret(static_cast<size_t>(n));
    if (n) check(tsdf_gather_fetch(h_, reinterpret_cast<float*>(ret.data()), n));
    return ret;
  }
  std::vector<Triangle> mesh(const float* bbox) {
    static_assert(sizeof(Triangle) == 36, "9 floats per triangle");
    int64_t n = 0;
    check(tsdf_extract_mesh(h_, bbox, nullptr, 0, &n));
    std::vector<Triangle> ret(static_cast<size_t>(n));
    if (n) check(tsdf_mesh_fetch(h_, reinterpret_cast<float*>(ret.data()), n));
    return ret;
  }
  tsdf_handle h_ = nullptr;
};

}  // namespace tsdf_b200
