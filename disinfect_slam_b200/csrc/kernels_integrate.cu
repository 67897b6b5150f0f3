// kernels_integrate.cu -- per-frame Integrate path of the B200 TSDF engine (sm_100a).
//
// Replaces, from scratch, the reference's nine launches + five host synchronisations per frame
// (utils/tsdf/voxel_tsdf.cu:347-397,456-488) by three launches and no host round trip:
//   frame_allocate_kernel   block_allocate_kernel (voxel_tsdf.cu:104-147) + per-pixel staging
//   select_visible_kernel   check_visibility_kernel + prefix_sum + gather_visible_blocks_kernel
//                           (voxel_tsdf.cu:82-102,456-472) over the dense pool directory
//   integrate_carve_kernel  tsdf_integrate_kernel + space_carving_kernel (voxel_tsdf.cu:149-230)
//                           fused: one pass over each visible block, 16-byte voxel accesses
#include "tsdf_device.cuh"
#include "tsdf_launch.h"

namespace tsdf {

// ------------------------------------------------------------------------------------------
// frame_allocate_kernel: one thread per pixel.
//  (1) staging: TexA{depth|0, range}, TexB{log ht - log lt, w_new, rgb} -- everything the
//      integrate kernel needs per pixel in two aligned gathers, with the per-pixel work
//      (range norm, depth/max_depth division, logs) done once per pixel instead of once per voxel.
//  (2) DDA over +-truncation along the pixel ray; candidate blocks are de-duplicated across the
//      warp with match.any before the (L2-resident) table is probed; only absent blocks pay the
//      8-corner visibility test and the CAS insert.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_allocate_kernel(DeviceState S, FrameParams P,
                                                             const unsigned char* __restrict__ rgb,
                                                             const float* __restrict__ depth,
                                                             const float* __restrict__ ht,
                                                             const float* __restrict__ lt,
                                                             TexA* __restrict__ texA, TexB* __restrict__ texB) {
  const int npix = P.w * P.h;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_img = idx < npix;
  const int y = in_img ? idx / P.w : 0;
  const int x = in_img ? idx - y * P.w : 0;

  float d = 0.f;
  float3 pos_cam = f3(0.f, 0.f, 1.f);
  float range = 1.f;
  bool valid = false;
  if (in_img) {
    d = depth[idx];
    // K^-1 * (x, y, 1)  -- utils/tsdf/voxel_tsdf.cu:118-120
    pos_cam = kmul(P.Kinv, f3((float)x, (float)y, 1.f));
    range = sqrtf(sqnorm3(pos_cam));
    valid = !(d == 0 || d > P.max_depth);
    TexA a; a.depth = valid ? d : 0.f; a.range = range;
    texA[idx] = a;
    TexB b;
    b.dlogit = logf(ht[idx]) - logf(lt[idx]);
    b.w_new = (1 - d / P.max_depth) * 4;  // voxel_tsdf.cu:182
    b.rgbx = (uint32_t)rgb[3 * idx] | ((uint32_t)rgb[3 * idx + 1] << 8) | ((uint32_t)rgb[3 * idx + 2] << 16);
    b.pad = 0;
    texB[idx] = b;
  }

  // ---- ray set-up, utils/tsdf/voxel_tsdf.cu:124-139 ----
  int nsteps = 0;
  float3 pos_grid = f3(0.f, 0.f, 0.f), ray_step_grid = f3(0.f, 0.f, 0.f);
  if (valid) {
    const float3 pos_world = apply(P.world_T_cam, f3(pos_cam.x * d, pos_cam.y * d, pos_cam.z * d));
    const float3 ray_dir_cam = f3(pos_cam.x / range, pos_cam.y / range, pos_cam.z / range);
    const float3 ray_dir_world = qrot(P.world_T_cam, ray_dir_cam);
    const float3 ray_start_world = f3(pos_world.x - ray_dir_world.x * P.truncation,
                                      pos_world.y - ray_dir_world.y * P.truncation,
                                      pos_world.z - ray_dir_world.z * P.truncation);
    const float3 ray_dir_grid = f3(ray_dir_world.x / P.voxel_size, ray_dir_world.y / P.voxel_size,
                                   ray_dir_world.z / P.voxel_size);
    pos_grid = f3(ray_start_world.x / P.voxel_size, ray_start_world.y / P.voxel_size,
                  ray_start_world.z / P.voxel_size);
    const float two_t = 2 * P.truncation;
    const float3 ray_grid = f3(two_t * ray_dir_grid.x, two_t * ray_dir_grid.y, two_t * ray_dir_grid.z);
    const int step_grid =
        __float2int_rz(ceilf(fmaxf(fmaxf(fabsf(ray_grid.x), fabsf(ray_grid.y)), fabsf(ray_grid.z)) / kBlockLen));
    const float denom = fmaxf((float)step_grid, 1);
    ray_step_grid = f3(ray_grid.x / denom, ray_grid.y / denom, ray_grid.z / denom);
    nsteps = step_grid + 1;  // for (i = 0; i <= step_grid; ++i)
  }

  const int max_steps = __reduce_max_sync(0xFFFFFFFFu, nsteps);
  const unsigned lane = threadIdx.x & 31;
  u64 prev_key = kEmpty;
  int n_cand = 0, n_new = 0;
  for (int i = 0; i < max_steps; ++i) {
    u64 key = kEmpty;  // sentinel: nothing to do for this lane
    if (i < nsteps) {
      const int px = round_to_voxel(pos_grid.x), py = round_to_voxel(pos_grid.y), pz = round_to_voxel(pos_grid.z);
      key = pack_key(px >> 3, py >> 3, pz >> 3);
      pos_grid = f3(pos_grid.x + ray_step_grid.x, pos_grid.y + ray_step_grid.y, pos_grid.z + ray_step_grid.z);
      if (key == prev_key) key = kEmpty; else prev_key = key;
    }
    // warp-cooperative de-duplication: one lane per distinct block coordinate probes the table
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
    const bool leader = (key != kEmpty) && ((unsigned)(__ffs(peers) - 1) == lane);
    if (leader && (S.shard_count <= 1 || owner_of(key, S.shard_count) == (unsigned)S.shard_rank)) {
      ++n_cand;
      int bx, by, bz; unpack_key(key, bx, by, bz);
      // Allocate() is a no-op for present blocks (voxel_hash.cu:62-77), so probe before the
      // (expensive) all-corners visibility test of voxel_tsdf.cu:144
      unsigned slot = hash_key(key) & S.table_mask;
      bool present = false;
      for (unsigned n = 0; n <= S.table_mask; ++n) {
        const u64 k = ld_key_cg(S.table + slot);
        if (k == key) { present = true; break; }
        if (k == kEmpty) break;
        slot = (slot + 1) & S.table_mask;
      }
      if (!present && block_visible<true>(bx, by, bz, P)) {
        if (table_insert(S, key) == 1) ++n_new;
      }
    }
  }
  // counters: one atomic per warp
  n_cand = __reduce_add_sync(0xFFFFFFFFu, n_cand);
  n_new = __reduce_add_sync(0xFFFFFFFFu, n_new);
  if (lane == 0) {
    if (n_cand) atomicAdd(&S.ctr[C_NCAND], n_cand);
    if (n_new) atomicAdd(&S.ctr[C_NNEW], n_new);
  }
}

// ------------------------------------------------------------------------------------------
// select_visible_kernel: any-corner visibility (voxel_tsdf.cu:82-93) over the dense pool
// directory block_key[0 .. high_water) -- 8 B per pool block instead of the reference's scan of
// all 2^22 hash entries + 3-launch prefix sum + compaction + host sync.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_visible_kernel(DeviceState S, FrameParams P, int* __restrict__ visible) {
  const int hw = S.ctr[C_HIGH_WATER];
  const unsigned lane = threadIdx.x & 31;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < hw; base += gridDim.x * blockDim.x) {
    const int i = base + lane;
    bool vis = false;
    if (i < hw) {
      const u64 k = S.block_key[i];
      if (k != kEmpty) {
        int bx, by, bz; unpack_key(k, bx, by, bz);
        vis = block_visible<false>(bx, by, bz, P);
      }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, vis);
    if (m) {
      int off = 0;
      if (lane == 0) off = atomicAdd(&S.ctr[C_NVIS], __popc(m));
      off = __shfl_sync(0xFFFFFFFFu, off, 0);
      if (vis) visible[off + __popc(m & ((1u << lane) - 1))] = i;
    }
  }
}

// ------------------------------------------------------------------------------------------
// integrate_carve_kernel: persistent CTAs of 128 threads; one visible block per iteration,
// 4 consecutive-x voxels per thread so that every voxel plane access is one 16-byte
// LDG/STG and a warp covers 512 contiguous bytes.
//   per voxel: voxel_tsdf.cu:157-203 (projection, nearest pixel, SDF, truncation, weighted
//   running averages, weight clamp); per block: voxel_tsdf.cu:214-229 (min |tsdf| >= .9 -> free).
// Fusions: blocks acquired this frame are initialised in registers (no init pass, no read);
// carved blocks are never written back; the carve reduction reuses the just-computed values.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld16(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint4 ld16u(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st16u(uint32_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

__global__ void __launch_bounds__(128) integrate_carve_kernel(DeviceState S, FrameParams P,
                                                              const int* __restrict__ visible,
                                                              const TexA* __restrict__ texA,
                                                              const TexB* __restrict__ texB,
                                                              float carve_threshold) {
  __shared__ float s_min[4];
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int n_vis = S.ctr[C_NVIS];
  const int vx0 = (t & 1) * 4, vy = (t >> 1) & 7, vz = t >> 4;  // voxel t*4 .. t*4+3 of the block
  unsigned n_upd_thread = 0;
  int n_carved_thread = 0;

  for (int b = blockIdx.x; b < n_vis; b += gridDim.x) {
    const int idx = visible[b];
    const u64 bk = S.block_key[idx];
    const bool is_new = (bk & kFlagNew) != 0;
    int bx, by, bz; unpack_key(bk, bx, by, bz);
    float* p_tsdf = block_tsdf(S, idx) + t * 4;
    uint32_t* p_rgbw = block_rgbw(S, idx) + t * 4;
    float* p_logit = block_logit(S, idx) + t * 4;

    // the TSDF plane is always needed (carve test); issue the load before the projection maths
    float tsdf[4];
    if (!is_new) { const float4 v = ld16(p_tsdf); tsdf[0] = v.x; tsdf[1] = v.y; tsdf[2] = v.z; tsdf[3] = v.w; }
    else { tsdf[0] = tsdf[1] = tsdf[2] = tsdf[3] = -1.f; }  // voxel_mem.cu:49

    // ---- projection + decision (voxel_tsdf.cu:157-176) ----
    const int gy = (short)((by << 3) + vy), gz = (short)((bz << 3) + vz);
    int pix[4];
    float tsdf_new[4];
    unsigned upd = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gx = (short)((bx << 3) + vx0 + j);
      const float3 pos_world = f3((float)gx * P.voxel_size, (float)gy * P.voxel_size, (float)gz * P.voxel_size);
      const float3 pos_cam = apply(P.cam_T_world, pos_world);
      const float3 pos_img_h = kmul(P.K, pos_cam);
      const int u = __float2int_rz(roundf(pos_img_h.x / pos_img_h.z));
      const int v = __float2int_rz(roundf(pos_img_h.y / pos_img_h.z));
      pix[j] = -1;
      tsdf_new[j] = 0.f;
      if (u >= 0 && u < P.w && v >= 0 && v < P.h) {
        const int img_idx = v * P.w + u;
        const float2 araw = __ldg(reinterpret_cast<const float2*>(texA + img_idx));
        TexA a; a.depth = araw.x; a.range = araw.y;
        if (a.depth != 0.f) {  // depth == 0 || depth > max_depth folded into the staging
          const float sdf = a.range * (a.depth - pos_img_h.z);
          if (sdf > -P.truncation) {
            tsdf_new[j] = fminf(1, sdf / P.truncation);
            pix[j] = img_idx;
            upd |= 1u << j;
          }
        }
      }
    }

    // ---- colour / weight / semantic planes only when this thread updates something ----
    uint32_t rgbw[4] = {0u, 0u, 0u, 0u};          // weight 0 (voxel_mem.cu:48); rgb defined as 0
    float logit[4] = {0.f, 0.f, 0.f, 0.f};        // probability .5 (voxel_mem.cu:50)
    if (upd && !is_new) {
      const uint4 c = ld16u(p_rgbw); rgbw[0] = c.x; rgbw[1] = c.y; rgbw[2] = c.z; rgbw[3] = c.w;
      const float4 l = ld16(p_logit); logit[0] = l.x; logit[1] = l.y; logit[2] = l.z; logit[3] = l.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (upd & (1u << j)) {
        const float4 braw = __ldg(reinterpret_cast<const float4*>(texB + pix[j]));
        TexB tb; tb.dlogit = braw.x; tb.w_new = braw.y; tb.rgbx = __float_as_uint(braw.z);
        const float weight_new = tb.w_new;
        const float weight_old = (float)(rgbw[j] >> 24);
        const float weight_combined = weight_old + weight_new;
        // voxel_tsdf.cu:186-194
        const float r_old = (float)(rgbw[j] & 0xFF), g_old = (float)((rgbw[j] >> 8) & 0xFF),
                    b_old = (float)((rgbw[j] >> 16) & 0xFF);
        const float r_new = (float)(tb.rgbx & 0xFF), g_new = (float)((tb.rgbx >> 8) & 0xFF),
                    b_new = (float)((tb.rgbx >> 16) & 0xFF);
        const unsigned r = (unsigned)__float2int_rz(roundf((r_old * weight_old + r_new * weight_new) / weight_combined));
        const unsigned g = (unsigned)__float2int_rz(roundf((g_old * weight_old + g_new * weight_new) / weight_combined));
        const unsigned bb = (unsigned)__float2int_rz(roundf((b_old * weight_old + b_new * weight_new) / weight_combined));
        tsdf[j] = (tsdf[j] * weight_old + tsdf_new[j] * weight_new) / weight_combined;
        const unsigned w = (unsigned)__float2int_rz(fminf(roundf(weight_combined), 40));
        rgbw[j] = (min(r, 255u)) | (min(g, 255u) << 8) | (min(bb, 255u) << 16) | (w << 24);
        // voxel_tsdf.cu:196-202 in logit space: logit' = (w_old*logit + w_new*(log ht - log lt)) / w
        logit[j] = (logit[j] * weight_old + tb.dlogit * weight_new) / weight_combined;
      }
    }
    n_upd_thread += __popc(upd);

    // ---- space carving (voxel_tsdf.cu:214-229): min |tsdf| over the 512 voxels ----
    float m = fminf(fminf(fabsf(tsdf[0]), fabsf(tsdf[1])), fminf(fabsf(tsdf[2]), fabsf(tsdf[3])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if (lane == 0) s_min[warp] = m;
    __syncthreads();
    const float block_min = fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3]));
    __syncthreads();  // s_min is reused by the next iteration

    if (block_min >= carve_threshold) {
      if (t == 0) { table_erase(S, bk & kKeyMask); ++n_carved_thread; }
    } else {
      if (upd || is_new) {
        st16(p_tsdf, make_float4(tsdf[0], tsdf[1], tsdf[2], tsdf[3]));
        st16u(p_rgbw, make_uint4(rgbw[0], rgbw[1], rgbw[2], rgbw[3]));
        st16(p_logit, make_float4(logit[0], logit[1], logit[2], logit[3]));
      }
      if (is_new && t == 0) S.block_key[idx] = bk & kKeyMask;
    }
  }

  // ---- counters: one atomic per warp ----
  n_upd_thread = __reduce_add_sync(0xFFFFFFFFu, n_upd_thread);
  if (lane == 0 && n_upd_thread) atomicAdd(reinterpret_cast<u64*>(&S.ctr[C_NUPD_LO]), (u64)n_upd_thread);
  if (t == 0 && n_carved_thread) atomicAdd(&S.ctr[C_NCARVED], n_carved_thread);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void launch_frame_allocate(const DeviceState& S, const FrameParams& P, const unsigned char* rgb, const float* depth,
                           const float* ht, const float* lt, TexA* texA, TexB* texB, cudaStream_t st) {
  const int npix = P.w * P.h;
  frame_allocate_kernel<<<(npix + 255) / 256, 256, 0, st>>>(S, P, rgb, depth, ht, lt, texA, texB);
}
void launch_select_visible(const DeviceState& S, const FrameParams& P, int* visible, int num_sms, cudaStream_t st) {
  select_visible_kernel<<<num_sms * 4, 256, 0, st>>>(S, P, visible);
}
void launch_integrate_carve(const DeviceState& S, const FrameParams& P, const int* visible, const TexA* texA,
                            const TexB* texB, int num_sms, cudaStream_t st) {
  integrate_carve_kernel<<<num_sms * 8, 128, 0, st>>>(S, P, visible, texA, texB, .9f);
}

}  // namespace tsdf
