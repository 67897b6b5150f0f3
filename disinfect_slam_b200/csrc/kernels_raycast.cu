// kernels_raycast.cu -- RayCast of the B200 TSDF engine (sm_100a).
//
// Replaces ray_cast_kernel (utils/tsdf/voxel_tsdf.cu:232-307): fixed-step march
// (step = truncation / 2), nearest-voxel lookups with a per-thread block cache (including the
// reference's negative cache for absent blocks, voxel_hash.cuh:124-161), sign-change hit test,
// bisection refinement, central-difference normal, Lambert shading and the semantic overlay.
// The sample positions are accumulated exactly like the reference (pos += step, float32, no FMA)
// so the hit decisions are bit-identical.  New output: per-ray hit depth (camera z, metres) and
// a packed (depth_bits << 32 | colour) key for nearest-hit min-compositing across GPUs.
#include <math_constants.h>

#include "tsdf_device.cuh"
#include "tsdf_launch.h"

namespace tsdf {

struct BlockCache { u64 key; int idx; };

__device__ __forceinline__ void cache_lookup(const DeviceState& S, BlockCache& c, int px, int py, int pz) {
  const u64 key = pack_key(px >> 3, py >> 3, pz >> 3);
  if (key != c.key) { c.key = key; c.idx = table_find(S, key); }
}
// Retrieve<VoxelTSDF>: absent -> VoxelTSDF() == +1 (voxel_types.cu:8)
__device__ __forceinline__ float fetch_tsdf(const DeviceState& S, BlockCache& c, int px, int py, int pz) {
  cache_lookup(S, c, px, py, pz);
  if (c.idx < 0) return 1.f;
  return __ldg(block_tsdf(S, c.idx) + voxel_index(px, py, pz));
}
__device__ __forceinline__ float fetch_tsdf_f(const DeviceState& S, BlockCache& c, float3 p) {
  return fetch_tsdf(S, c, round_to_voxel(p.x), round_to_voxel(p.y), round_to_voxel(p.z));
}

__device__ __forceinline__ unsigned char f2u8(float f) { return (unsigned char)min(255, max(0, __float2int_rz(f))); }

__global__ void __launch_bounds__(256) raycast_kernel(DeviceState S, FrameParams P, float step_size,
                                                      uchar4* __restrict__ img_rgba, uchar4* __restrict__ img_normal,
                                                      float* __restrict__ img_depth, u64* __restrict__ packed) {
  // 16x16 pixel tiles: neighbouring rays walk the same blocks (L1 reuse of table slots and voxels)
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= P.w || y >= P.h) return;
  const int idx = y * P.w + x;

  // voxel_tsdf.cu:243-250
  const float3 pos_cam = kmul(P.Kinv, f3((float)x, (float)y, 1.f));
  const float sq = sqnorm3(pos_cam);
  float3 ray_dir_cam = pos_cam;
  if (sq > 0.f) { const float n = sqrtf(sq); ray_dir_cam = f3(pos_cam.x / n, pos_cam.y / n, pos_cam.z / n); }
  const float3 ray_dir_world = qrot(P.world_T_cam, ray_dir_cam);
  const float3 ray_step_grid = f3(ray_dir_world.x * step_size / P.voxel_size, ray_dir_world.y * step_size / P.voxel_size,
                                  ray_dir_world.z * step_size / P.voxel_size);
  const int max_step = __float2int_rz(ceilf(P.max_depth / step_size));
  float3 pos_grid = f3(P.world_T_cam.tx / P.voxel_size, P.world_T_cam.ty / P.voxel_size, P.world_T_cam.tz / P.voxel_size);

  BlockCache cache; cache.key = kEmpty; cache.idx = -1;
  float tsdf_prev = fetch_tsdf_f(S, cache, pos_grid);
  pos_grid = f3(pos_grid.x + ray_step_grid.x, pos_grid.y + ray_step_grid.y, pos_grid.z + ray_step_grid.z);

  uchar4 out_rgba = make_uchar4(0, 0, 0, 0), out_normal = make_uchar4(0, 0, 0, 0);
  float out_depth = CUDART_INF_F;

  for (int i = 1; i < max_step; ++i) {
    const float tsdf_curr = fetch_tsdf_f(S, cache, pos_grid);
    // ray hit front surface (voxel_tsdf.cu:260)
    if (tsdf_prev > 0 && tsdf_curr <= 0 && tsdf_prev - tsdf_curr <= 1.5f) {
      float3 pos1 = f3(pos_grid.x - ray_step_grid.x, pos_grid.y - ray_step_grid.y, pos_grid.z - ray_step_grid.z);
      float3 pos2 = pos_grid;
      float3 mid = f3((pos1.x + pos2.x) / 2.f, (pos1.y + pos2.y) / 2.f, (pos1.z + pos2.z) / 2.f);
      // binary search refinement; `> .1` compares in double in the reference, i.e. >= 0.1f for floats
      for (;;) {
        const float3 dd = f3(pos1.x - pos2.x, pos1.y - pos2.y, pos1.z - pos2.z);
        if (!(dot3(dd, dd) >= 0.1f)) break;
        const float tm = fetch_tsdf_f(S, cache, mid);
        if (tm < 0) pos2 = mid; else pos1 = mid;
        mid = f3((pos1.x + pos2.x) / 2.f, (pos1.y + pos2.y) / 2.f, (pos1.z + pos2.z) / 2.f);
      }
      const int fx = round_to_voxel(mid.x), fy = round_to_voxel(mid.y), fz = round_to_voxel(mid.z);
      cache_lookup(S, cache, fx, fy, fz);
      uint32_t rgbw = 0u;  // VoxelRGBW() / VoxelSEGM() defaults for an absent voxel (voxel_types.cu:3,11)
      float prob = 0.f;
      if (cache.idx >= 0) {
        const int k = voxel_index(fx, fy, fz);
        rgbw = __ldg(block_rgbw(S, cache.idx) + k);
        prob = logit_to_prob(__ldg(block_logit(S, cache.idx) + k));
      }
      // central differences on nearest voxels (voxel_tsdf.cu:280-291); short arithmetic wraps like the reference
      const float gxp = fetch_tsdf(S, cache, (short)(fx + 1), fy, fz), gxn = fetch_tsdf(S, cache, (short)(fx - 1), fy, fz);
      const float gyp = fetch_tsdf(S, cache, fx, (short)(fy + 1), fz), gyn = fetch_tsdf(S, cache, fx, (short)(fy - 1), fz);
      const float gzp = fetch_tsdf(S, cache, fx, fy, (short)(fz + 1)), gzn = fetch_tsdf(S, cache, fx, fy, (short)(fz - 1));
      const float3 nrm = f3(gxp - gxn, gyp - gyn, gzp - gzn);
      const float3 neg_dir = f3(-ray_dir_world.x, -ray_dir_world.y, -ray_dir_world.z);
      const float diffusivity = fmaxf(dot3(nrm, neg_dir) / sqrtf(sqnorm3(nrm)), 0);
      const float alpha = fmaxf(prob - 0.5f, 0) * 2.f;  // == fmaxf(p - .5, 0) / .5 exactly
      const float r = (float)(rgbw & 0xFF), g = (float)((rgbw >> 8) & 0xFF), b = (float)((rgbw >> 16) & 0xFF);
      out_rgba = make_uchar4(f2u8(alpha * 255 + (1 - alpha) * r), f2u8((1 - alpha) * g), f2u8((1 - alpha) * b), 255);
      out_normal = make_uchar4(f2u8(alpha * 255 + (1 - alpha) * diffusivity * 255), f2u8((1 - alpha) * diffusivity * 255),
                               f2u8((1 - alpha) * diffusivity * 255), 255);
      const float3 pc = apply(P.cam_T_world, f3(mid.x * P.voxel_size, mid.y * P.voxel_size, mid.z * P.voxel_size));
      out_depth = pc.z;
      break;
    }
    tsdf_prev = tsdf_curr;
    pos_grid = f3(pos_grid.x + ray_step_grid.x, pos_grid.y + ray_step_grid.y, pos_grid.z + ray_step_grid.z);
  }

  if (img_rgba) img_rgba[idx] = out_rgba;
  if (img_normal) img_normal[idx] = out_normal;
  if (img_depth) img_depth[idx] = out_depth;
  if (packed) {
    // positive floats order like unsigned ints; a miss (+inf) loses against every hit
    const u64 dbits = (u64)__float_as_uint(fmaxf(out_depth, 0.f)) << 32;
    packed[2 * idx + 0] = dbits | (u64)(*reinterpret_cast<const uint32_t*>(&out_rgba));
    packed[2 * idx + 1] = dbits | (u64)(*reinterpret_cast<const uint32_t*>(&out_normal));
  }
}

void launch_raycast(const DeviceState& S, const FrameParams& P, float step_size, uchar4* rgba, uchar4* normal,
                    float* hit_depth, unsigned long long* packed_keys, cudaStream_t st) {
  dim3 grid((P.w + 15) / 16, (P.h + 15) / 16);
  raycast_kernel<<<grid, 256, 0, st>>>(S, P, step_size, rgba, normal, hit_depth, packed_keys);
}

}  // namespace tsdf
