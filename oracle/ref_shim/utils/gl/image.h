// Headless stand-in for the reference's utils/gl/image.h (CUDA<->OpenGL texture sink): the B200 box has
// no GL.  LoadCuda is a no-op; the harness reads the render buffers directly -- test infrastructure only.
#pragma once
#include <cuda_runtime.h>
class GLImageBase {
 public:
  void LoadCuda(const void*, cudaStream_t = nullptr) {}
};
class GLImage32FC1 : public GLImageBase {};
class GLImage8UC4 : public GLImageBase {};
