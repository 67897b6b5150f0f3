// Minimal cv::Mat stand-in (host image container) for the reference rebuild -- test infrastructure only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
#define CV_8UC3 16
#define CV_32FC1 5
namespace cv {
class Mat {
 public:
  unsigned char* data = nullptr;
  int rows = 0, cols = 0, type_ = 0;
  Mat() {}
  Mat(int r, int c, int t, void* d) : data((unsigned char*)d), rows(r), cols(c), type_(t) {}
  static Mat ones(int r, int c, int t) {  // CV_32FC1 only (what modules/tsdf_module.cc:31-32 asks for)
    Mat m;
    m.own_ = std::make_shared<std::vector<float>>((size_t)r * c, 1.f);
    m.data = (unsigned char*)m.own_->data(); m.rows = r; m.cols = c; m.type_ = t;
    return m;
  }
  size_t total() const { return (size_t)rows * cols; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || total() == 0; }

 private:
  std::shared_ptr<std::vector<float>> own_;
};
}  // namespace cv
