// Minimal cv::Mat stand-in (host image container) for the reference rebuild -- test infrastructure only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
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
  size_t total() const { return (size_t)rows * cols; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || total() == 0; }
};
}  // namespace cv
