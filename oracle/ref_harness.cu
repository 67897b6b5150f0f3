// ref_harness.cu -- C entry points around the UNMODIFIED reference TSDFGrid
// (/root/reference/utils/tsdf/voxel_tsdf.{cuh,cu} and friends, compiled from where they lie by
// oracle/build_ref.sh against the stand-in headers in oracle/ref_shim).
//
// TEST INFRASTRUCTURE ONLY: used by tests/ (parity of the new engine against the reference's own
// CUDA kernels on the same inputs), by tests/golden/make_ref_golden.py and by bench.py's
// `--impl reference` arm.  Nothing here is linked into libtsdf_b200.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "utils/tsdf/voxel_tsdf.cuh"

namespace {

struct DumpHdr { int n; };

// every allocated hash entry -> key + raw voxel planes (the reference has no such export; GatherValid
// only returns positions + tsdf)
__global__ void dump_entries_kernel(const VoxelHashTable table, int cap, int* count, short* keys, float* tsdf,
                                    unsigned char* rgbw, float* prob) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NUM_ENTRY) return;
  const VoxelBlock& b = table.GetBlock(idx);
  if (b.idx < 0) return;
  const int o = atomicAdd(count, 1);
  if (o >= cap) return;
  keys[3 * o + 0] = b.position[0]; keys[3 * o + 1] = b.position[1]; keys[3 * o + 2] = b.position[2];
  if (!tsdf) return;
  for (int k = 0; k < BLOCK_VOLUME; ++k) {
    const VoxelRGBW& c = table.mem.GetVoxel<VoxelRGBW>(k, b);
    tsdf[(size_t)o * BLOCK_VOLUME + k] = table.mem.GetVoxel<VoxelTSDF>(k, b).tsdf;
    prob[(size_t)o * BLOCK_VOLUME + k] = table.mem.GetVoxel<VoxelSEGM>(k, b).probability;
    unsigned char* d = rgbw + ((size_t)o * BLOCK_VOLUME + k) * 4;
    d[0] = c.rgb[0]; d[1] = c.rgb[1]; d[2] = c.rgb[2]; d[3] = c.weight;
  }
}

class RefGrid : public TSDFGrid {
 public:
  RefGrid(float vs, float tr) : TSDFGrid(vs, tr) {}
  const uchar4* rgba() const { return img_tsdf_rgba_; }
  const uchar4* normal() const { return img_tsdf_normal_; }
  cudaStream_t stream() const { return stream_; }
  const VoxelHashTable& table() const { return hash_table_; }
  int num_active() const { return hash_table_.NumActiveBlock(); }
};

SE3<float> make_pose(const float q[4], const float t[3]) {
  return SE3<float>(Eigen::Quaternionf(q[3], q[0], q[1], q[2]), Eigen::Vector3f(t[0], t[1], t[2]));
}

}  // namespace

extern "C" {

// The reference's init_hash_table_kernel leaves VoxelBlock::offset uninitialised (voxel_hash.cu:26-29)
// although Allocate reads it (voxel_hash.cu:71).  Fresh cudaMalloc memory is zero in practice; to keep
// that true when memory is recycled inside one process, hand the allocator zeroed memory first.
void* ref_create(float voxel_size, float truncation) {
  void* scrub = nullptr;
  const size_t n = (size_t)3 << 30;
  if (cudaMalloc(&scrub, n) == cudaSuccess) { cudaMemset(scrub, 0, n); cudaDeviceSynchronize(); cudaFree(scrub); }
  else cudaGetLastError();
  return new RefGrid(voxel_size, truncation);
}
void ref_destroy(void* h) { delete static_cast<RefGrid*>(h); }

int ref_integrate(void* h, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int w, int hgt,
                  float max_depth, const float K[4], const float q[4], const float t[3]) {
  RefGrid* g = static_cast<RefGrid*>(h);
  cv::Mat m_rgb(hgt, w, CV_8UC3, (void*)rgb), m_d(hgt, w, CV_32FC1, (void*)depth), m_ht(hgt, w, CV_32FC1, (void*)ht),
      m_lt(hgt, w, CV_32FC1, (void*)lt);
  g->Integrate(m_rgb, m_d, m_ht, m_lt, max_depth, CameraIntrinsics<float>(K[0], K[1], K[2], K[3]), make_pose(q, t));
  return (int)cudaGetLastError();
}

// rgba / normal: host HxWx4 or NULL (then only the kernel runs, like the GL path of the reference)
int ref_raycast(void* h, float max_depth, int w, int hgt, const float K[4], const float q[4], const float t[3],
                uint8_t* rgba, uint8_t* normal) {
  RefGrid* g = static_cast<RefGrid*>(h);
  GLImage8UC4 sink_a, sink_b;
  g->RayCast(max_depth, CameraParams(CameraIntrinsics<float>(K[0], K[1], K[2], K[3]), hgt, w), make_pose(q, t), &sink_a, &sink_b);
  if (rgba) cudaMemcpyAsync(rgba, g->rgba(), (size_t)w * hgt * 4, cudaMemcpyDeviceToHost, g->stream());
  if (normal) cudaMemcpyAsync(normal, g->normal(), (size_t)w * hgt * 4, cudaMemcpyDeviceToHost, g->stream());
  cudaStreamSynchronize(g->stream());
  return (int)cudaGetLastError();
}

long long ref_gather_valid(void* h, float* out, long long cap) {
  std::vector<VoxelSpatialTSDF> v = static_cast<RefGrid*>(h)->GatherValid();
  cudaDeviceSynchronize();  // the reference frees the device buffer right after an async copy (voxel_tsdf.cu:418-423)
  if (out) memcpy(out, v.data(), sizeof(VoxelSpatialTSDF) * (size_t)std::min<long long>(cap, (long long)v.size()));
  return (long long)v.size();
}
long long ref_gather_voxels(void* h, const float bbox[6], float* out, long long cap) {
  const BoundingCube<float> b = {bbox[0], bbox[1], bbox[2], bbox[3], bbox[4], bbox[5]};
  std::vector<VoxelSpatialTSDF> v = static_cast<RefGrid*>(h)->GatherVoxels(b);
  cudaDeviceSynchronize();
  if (out) memcpy(out, v.data(), sizeof(VoxelSpatialTSDF) * (size_t)std::min<long long>(cap, (long long)v.size()));
  return (long long)v.size();
}
int ref_num_active(void* h) { return static_cast<RefGrid*>(h)->num_active(); }

// all allocated blocks in canonical order (ascending z, y, x block coordinate); any output may be NULL
int ref_export(void* h, int16_t* keys, float* tsdf, uint8_t* rgbw, float* prob, int cap, int* n_out) {
  RefGrid* g = static_cast<RefGrid*>(h);
  cudaDeviceSynchronize();
  int* d_count = nullptr; short* d_keys = nullptr; float *d_tsdf = nullptr, *d_prob = nullptr; unsigned char* d_rgbw = nullptr;
  const int dcap = std::max(cap, 1);
  const bool vox = tsdf || rgbw || prob;
  cudaMalloc(&d_count, sizeof(int)); cudaMemset(d_count, 0, sizeof(int));
  cudaMalloc(&d_keys, sizeof(short) * 3 * (size_t)dcap);
  if (vox) {
    cudaMalloc(&d_tsdf, sizeof(float) * BLOCK_VOLUME * (size_t)dcap); cudaMalloc(&d_prob, sizeof(float) * BLOCK_VOLUME * (size_t)dcap);
    cudaMalloc(&d_rgbw, 4 * BLOCK_VOLUME * (size_t)dcap);
  }
  dump_entries_kernel<<<NUM_ENTRY / 256, 256>>>(g->table(), cap, d_count, d_keys, vox ? d_tsdf : nullptr, d_rgbw, d_prob);
  int n = 0;
  cudaMemcpy(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost);
  if (n_out) *n_out = n;
  const int m = std::min(n, cap);
  if (m > 0 && keys) {
    std::vector<short> hk(3 * (size_t)m);
    cudaMemcpy(hk.data(), d_keys, sizeof(short) * hk.size(), cudaMemcpyDeviceToHost);
    std::vector<int> order(m);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      for (int c = 2; c >= 0; --c) if (hk[3 * a + c] != hk[3 * b + c]) return hk[3 * a + c] < hk[3 * b + c];
      return false;
    });
    std::vector<float> ht, hp; std::vector<unsigned char> hc;
    if (vox) {
      ht.resize((size_t)m * BLOCK_VOLUME); hp.resize((size_t)m * BLOCK_VOLUME); hc.resize((size_t)m * BLOCK_VOLUME * 4);
      cudaMemcpy(ht.data(), d_tsdf, sizeof(float) * ht.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(hp.data(), d_prob, sizeof(float) * hp.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(hc.data(), d_rgbw, hc.size(), cudaMemcpyDeviceToHost);
    }
    for (int i = 0; i < m; ++i) {
      const int s = order[i];
      memcpy(keys + 3 * (size_t)i, hk.data() + 3 * (size_t)s, sizeof(short) * 3);
      if (tsdf) memcpy(tsdf + (size_t)i * BLOCK_VOLUME, ht.data() + (size_t)s * BLOCK_VOLUME, sizeof(float) * BLOCK_VOLUME);
      if (prob) memcpy(prob + (size_t)i * BLOCK_VOLUME, hp.data() + (size_t)s * BLOCK_VOLUME, sizeof(float) * BLOCK_VOLUME);
      if (rgbw) memcpy(rgbw + (size_t)i * BLOCK_VOLUME * 4, hc.data() + (size_t)s * BLOCK_VOLUME * 4, (size_t)BLOCK_VOLUME * 4);
    }
  }
  cudaFree(d_count); cudaFree(d_keys); cudaFree(d_tsdf); cudaFree(d_prob); cudaFree(d_rgbw);
  return (int)cudaGetLastError();
}

}  // extern "C"
