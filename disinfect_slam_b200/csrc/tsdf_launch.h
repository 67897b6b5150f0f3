// tsdf_launch.h -- host-callable launchers of the engine's kernels (one per .cu file group).
#pragma once
#include <cuda_runtime.h>

#include "tsdf_device.cuh"

namespace tsdf {

// kernels_integrate.cu
void launch_frame_allocate(const DeviceState& S, const FrameParams& P, const FrameInput& in, Texel* tex, cudaStream_t st);
void launch_insert_candidates(const DeviceState& S, const FrameParams& P, cudaStream_t st);
void launch_select_visible(const DeviceState& S, const FrameParams& P, int* visible, int* vis_state, int num_sms,
                           cudaStream_t st);
void launch_integrate_carve(const DeviceState& S, const FrameParams& P, const int* visible, int* vis_state, const Texel* tex,
                            int num_sms, int expected_blocks, cudaStream_t st);

// kernels_raycast.cu
void launch_build_skip_map(const PeerView* shards, int n_shards, const SkipMap& M, int gen, bool lazy, int num_sms,
                           cudaStream_t st);
// pulled TSDF cache of a sharded volume (kernels_raycast.cu): local [shard][stride][512] floats + one stamp per slot,
// the list / counters of the fetch kernels; epoch = content version the stamps must carry, serial = number of this launch
struct SharedCache { float* cache; int* stamp; int* list; int* count; int stride, epoch, serial, self, pad; };
void launch_raycast_shared(const PeerView* shards, const PeerView* host_shards, int n_shards, int shard_shift, const FrameParams& P, float step_size,
                           const SkipMap& M, int row0, int rows, int tile_stride, const float* mirror, int mirror_stride, const SharedCache& C,
                           uchar4* rgba, uchar4* normal, float* hit_depth, int n_out, void* const* out_rgba, void* const* out_normal,
                           void* const* out_depth, int num_sms, cudaStream_t st);
void launch_raycast(const DeviceState& S, const FrameParams& P, float step_size, const SkipMap& M, uchar4* rgba,
                    uchar4* normal, float* hit_depth, unsigned long long* packed_keys, cudaStream_t st);

// kernels_gather.cu
struct GridBound { short xmin, xmax, ymin, ymax, zmin, zmax; };
void launch_init_state(const DeviceState& S, cudaStream_t st);
void launch_select_blocks(const DeviceState& S, bool use_bound, GridBound bound, int* selected, int num_sms,
                          cudaStream_t st);
void launch_download_voxels(const DeviceState& S, const int* selected, int n_selected, float voxel_size, float4* out,
                            cudaStream_t st);
void launch_export_blocks(const DeviceState& S, const int* selected, int n_selected, short* keys, float* tsdf,
                          unsigned* rgbw, float* prob, cudaStream_t st);
void launch_allocate_list(const DeviceState& S, const short* keys, int n, cudaStream_t st);
void launch_delete_list(const DeviceState& S, const short* keys, int n, cudaStream_t st);
void launch_retrieve_list(const DeviceState& S, const short* points, int n, float* tsdf, unsigned* rgbw, float* prob,
                          int* found, cudaStream_t st);
void launch_assign_list(const DeviceState& S, const short* points, int n, const float* tsdf, const unsigned* rgbw,
                        const float* prob, cudaStream_t st);
void launch_rehash(const DeviceState& S, int num_sms, cudaStream_t st);
void launch_publish_counters(const DeviceState& S, int* host_mapped, cudaStream_t st);

// kernels_mesh.cu
void launch_mesh_count(const DeviceState& S, const int* selected, int n_selected, float voxel_size, unsigned long long* counter,
                       cudaStream_t st);
void launch_mesh_emit(const DeviceState& S, const int* selected, int n_selected, float voxel_size, float* out, long long cap_triangles,
                      unsigned long long* counter, cudaStream_t st);

}  // namespace tsdf
