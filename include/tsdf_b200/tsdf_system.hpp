// tsdf_system.hpp -- B200-native counterpart of the reference's TSDFSystem (modules/tsdf_module.{h,cc} of
// yuzhou42/disinfect-slam): the asynchronous front end every online path uses (examples/tsdf/online*.cc,
// disinfect_slam/disinfect_slam.cc:66,113) -- SURVEY.md section 8(f) rank 1.
//
// Same public surface (constructor arguments, Integrate / Query / Render, argument order and meaning), over the C
// ABI of libtsdf_b200.so.  What is different from the reference's implementation:
//   * the worker sleeps on a condition variable (the reference's Run() spins on an empty queue,
//     tsdf_module.cc:54-75, one core at 100 %);
//   * frames go through tsdf_integrate_async: the upload of frame k+1 overlaps the kernels of frame k, and the
//     worker never waits for the GPU unless a reader (Query / Render) or Flush() asks for the result;
//   * a frame without probability images uses one cached plane of ones (the reference allocates two
//     cv::Mat::ones per frame, tsdf_module.cc:28-33);
//   * the backlog is observable (Backlog()) and can be bounded (max_backlog: Integrate then blocks instead of
//     letting the queue grow without limit -- the reference only logs a warning above 10 frames);
//   * Flush() waits until every queued frame is in the volume (the reference has no way to know);
//   * an engine error raised on the worker (pool exhausted, CUDA error) is rethrown by the next call.
// Like tsdf_b200::TSDFGrid the class is duck-typed over the reference's own types (cv::Mat, SE3<float>,
// CameraIntrinsics<float>, CameraParams, GLImage8UC4, BoundingCube<float>), so it needs no OpenCV / Eigen / GL
// itself.  include/tsdf_b200/compat_system/modules/tsdf_module.h instantiates it as the global `TSDFSystem`.
#pragma once
#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

#include "voxel_tsdf.hpp"

namespace tsdf_b200 {

template <class Mat, class Intrinsics, class Pose, class Voxel = VoxelSpatialTSDF>
class TSDFSystemT {
 public:
  // voxel_size, truncation, max_depth [m]; intrinsics of the RGB-D camera; extrinsics = cam_T_posecam
  // (tsdf_module.h:35-49).  max_backlog = 0: unbounded queue like the reference.
  TSDFSystemT(float voxel_size, float truncation, float max_depth, const Intrinsics& intrinsics,
              const Pose& extrinsics = Pose::Identity(), size_t max_backlog = 0, const tsdf_config* cfg = nullptr)
      : tsdf_(voxel_size, truncation, cfg), max_depth_(max_depth), intrinsics_(intrinsics), cam_T_posecam_(extrinsics),
        max_backlog_(max_backlog), t_(&TSDFSystemT::Run, this) {}

  // stops after the frame being integrated; frames still queued are dropped, as in the reference (tsdf_module.cc:18-24)
  ~TSDFSystemT() {
    {
      std::lock_guard<std::mutex> lock(mtx_queue_);
      terminate_ = true;
    }
    cv_work_.notify_all();
    cv_space_.notify_all();
    t_.join();
  }
  TSDFSystemT(const TSDFSystemT&) = delete;
  TSDFSystemT& operator=(const TSDFSystemT&) = delete;

  // tsdf_module.cc:26-38: queue one frame; empty ht / lt mean "probability 1" for both classes
  void Integrate(const Pose& posecam_T_world, const Mat& img_rgb, const Mat& img_depth, const Mat& img_ht = Mat(),
                 const Mat& img_lt = Mat()) {
    std::unique_lock<std::mutex> lock(mtx_queue_);
    Rethrow();
    if (max_backlog_) cv_space_.wait(lock, [&] { return inputs_.size() < max_backlog_ || terminate_; });
    const bool no_probs = img_ht.empty() || img_lt.empty();
    inputs_.push_back(Input{cam_T_posecam_ * posecam_T_world, img_rgb, img_depth, no_probs ? Mat() : img_ht, no_probs ? Mat() : img_lt});
    lock.unlock();
    cv_work_.notify_one();
  }

  // tsdf_module.cc:40-43
  template <class Cube>
  std::vector<Voxel> Query(const Cube& volumn) {
    std::lock_guard<std::mutex> lock(mtx_read_);
    return tsdf_.template GatherVoxels<Voxel>(volumn);
  }

  // tsdf_module.cc:45-49
  template <class CamParams, class Image>
  void Render(const CamParams& virtual_cam, const Pose cam_T_world, Image* img_normal) {
    std::lock_guard<std::mutex> lock(mtx_read_);
    tsdf_.RayCast(max_depth_, virtual_cam, cam_T_world, static_cast<std::nullptr_t*>(nullptr), img_normal);
  }

  // ---- additions ----
  // every frame queued before the call is in the volume when it returns
  void Flush() {
    std::unique_lock<std::mutex> lock(mtx_queue_);
    cv_idle_.wait(lock, [&] { return (inputs_.empty() && !busy_) || terminate_ || error_; });
    Rethrow();
    lock.unlock();
    std::lock_guard<std::mutex> rlock(mtx_read_);
    check(tsdf_synchronize(tsdf_.handle()));
  }
  size_t Backlog() {
    std::lock_guard<std::mutex> lock(mtx_queue_);
    return inputs_.size() + (busy_ ? 1 : 0);
  }
  int64_t FramesIntegrated() {
    std::lock_guard<std::mutex> lock(mtx_queue_);
    return n_done_;
  }
  TSDFGrid& Grid() { return tsdf_; }  // callers must hold no frame in flight (Flush() first)

 private:
  struct Input { Pose cam_T_world; Mat img_rgb, img_depth, img_ht, img_lt; };

  void Rethrow() {  // mtx_queue_ held
    if (error_) { std::exception_ptr e = error_; error_ = nullptr; std::rethrow_exception(e); }
  }

  void Run() {
    for (;;) {
      Input in;
      {
        std::unique_lock<std::mutex> lock(mtx_queue_);
        cv_work_.wait(lock, [&] { return terminate_ || !inputs_.empty(); });
        if (terminate_) return;
        in = std::move(inputs_.front());
        inputs_.pop_front();
        busy_ = true;
      }
      cv_space_.notify_one();
      std::exception_ptr err;
      try {
        std::lock_guard<std::mutex> lock(mtx_read_);
        const size_t n = (size_t)in.img_depth.rows * in.img_depth.cols;
        const float *ht, *lt;
        if (in.img_ht.empty()) {
          if (ones_.size() != n) ones_.assign(n, 1.f);
          ht = lt = ones_.data();
        } else {
          ht = reinterpret_cast<const float*>(in.img_ht.data);
          lt = reinterpret_cast<const float*>(in.img_lt.data);
        }
        if (in.img_rgb.type() != kCV_8UC3 || in.img_depth.type() != kCV_32FC1 || in.img_rgb.cols != in.img_depth.cols ||
            in.img_rgb.rows != in.img_depth.rows)
          throw Error(TSDF_E_INVALID, "Integrate: rgb must be CV_8UC3, depth CV_32FC1, same size (voxel_tsdf.cu:350-353)");
        if (!in.img_ht.empty() && (in.img_ht.total() != n || in.img_lt.total() != n))
          throw Error(TSDF_E_INVALID, "Integrate: ht / lt must have the size of depth");
        const float K[4] = {intrinsics_.fx, intrinsics_.fy, intrinsics_.cx, intrinsics_.cy};
        const auto R = in.cam_T_world.GetR();
        const auto tr = in.cam_T_world.GetT();
        const float q[4] = {R.x(), R.y(), R.z(), R.w()}, t[3] = {tr[0], tr[1], tr[2]};
        // returns once the host images have been consumed; the kernels of this frame overlap the next upload
        check(tsdf_integrate_async(tsdf_.handle(), reinterpret_cast<const uint8_t*>(in.img_rgb.data),
                                   reinterpret_cast<const float*>(in.img_depth.data), ht, lt, in.img_depth.cols, in.img_depth.rows,
                                   max_depth_, K, q, t));
      } catch (...) {
        err = std::current_exception();
      }
      {
        std::lock_guard<std::mutex> lock(mtx_queue_);
        busy_ = false;
        if (err) error_ = err; else ++n_done_;
      }
      cv_idle_.notify_all();
    }
  }

  TSDFGrid tsdf_;
  const float max_depth_;
  const Intrinsics intrinsics_;
  const Pose cam_T_posecam_;
  const size_t max_backlog_;
  std::mutex mtx_queue_, mtx_read_;
  std::condition_variable cv_work_, cv_space_, cv_idle_;
  std::deque<Input> inputs_;
  std::vector<float> ones_;
  std::exception_ptr error_ = nullptr;
  bool terminate_ = false, busy_ = false;
  int64_t n_done_ = 0;
  std::thread t_;  // last member: starts after everything above is constructed
};

}  // namespace tsdf_b200
