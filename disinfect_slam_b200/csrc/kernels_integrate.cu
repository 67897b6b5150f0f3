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

#include <atomic>

namespace tsdf {

// ------------------------------------------------------------------------------------------
// frame_allocate_kernel: one thread per pixel.
//  (1) staging: Texel{depth|0, range, log ht - log lt, rgb} -- everything the integrate kernel
//      needs per pixel in ONE aligned 16-byte gather, with the per-pixel work (range norm, logs)
//      done once per pixel instead of once per voxel.
//  (2) DDA over +-truncation along the pixel ray; candidate blocks are de-duplicated across the
//      warp with match.any before the (L2-resident) table is probed; only absent blocks pay the
//      8-corner visibility test and the CAS insert.
// ------------------------------------------------------------------------------------------
constexpr int kInlineSteps = 4;

// Candidate exchange (DeviceState::xa_on).  The candidates of a CTA (32 x 8 pixels, a few blocks) are collected in a small
// shared-memory hash set first, so that every distinct block of the tile is mailed once: one cursor bump per owner and CTA
// instead of one per warp and DDA step on a handful of hot words, and ~5x fewer keys for the owners to look up.
constexpr int kXaSet = 256;  // slots of the per-CTA set (a tile whose rays meet more distinct blocks mails the rest directly)
__device__ __forceinline__ void xa_mail_one(const DeviceState& S, u64 key, int owner, int pos) {
  if (pos < S.xa_cap) {
    u64* keys = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(S.xa_inbox[owner]) + kXaHeaderBytes);
    keys[((size_t)(S.xa_parity * S.shard_count + S.shard_rank)) * S.xa_cap + pos] = key;  // local, or a posted store over NVLink
  } else {
    atomicOr(&S.ctr[C_ERROR], ERR_EXCHANGE);  // reported by this frame; the barrier kernel clamps the count it publishes
  }
}
__device__ __forceinline__ void xa_collect(const DeviceState& S, u64* s_set, u64 key) {
  unsigned h = (unsigned)hash_key(key) & (kXaSet - 1);
  for (int probe = 0; probe < kXaSet; ++probe) {
    const u64 prev = atomicCAS(reinterpret_cast<unsigned long long*>(s_set + h), (unsigned long long)kEmpty, (unsigned long long)key);
    if (prev == kEmpty || prev == key) return;
    h = (h + 1) & (kXaSet - 1);
  }
  const int owner = (int)owner_of(key, S.shard_count, S.shard_shift);  // set full
  xa_mail_one(S, key, owner, atomicAdd(&S.xa_cursor[S.xa_parity * 8 + owner], 1));
}
// end of the CTA (all threads): mail the set
__device__ __forceinline__ void xa_flush(const DeviceState& S, const u64* s_set, int* s_cnt, int* s_base) {
  __syncthreads();
  u64 key = kEmpty;
  int owner = -1, pos = 0;
  for (int t = threadIdx.x; t < kXaSet; t += blockDim.x) {  // (one slot per thread with 256 threads)
    key = s_set[t];
    if (key != kEmpty) { owner = (int)owner_of(key, S.shard_count, S.shard_shift); pos = atomicAdd(&s_cnt[owner], 1); }
  }
  __syncthreads();
  if ((int)threadIdx.x < S.shard_count && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&S.xa_cursor[S.xa_parity * 8 + threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
  if (key != kEmpty) xa_mail_one(S, key, owner, s_base[owner] + pos);
}
static_assert(kXaSet == 256, "xa_flush keeps one slot per thread of a 256-thread CTA");

template <bool EXCHANGE>
__global__ void __launch_bounds__(256, 6) frame_allocate_kernel(DeviceState S, FrameParams P, FrameInput in,
                                                             Texel* __restrict__ tex) {
  const unsigned char* __restrict__ rgb = in.rgb;
  // CTA = 32 x 8 pixels, warp = 8 x 4 pixels: the pixels of a warp fall into one or two blocks, so the
  // match.any de-duplication below leaves about one table probe per warp and DDA step
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
  const int y = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
  const bool in_img = x < P.w && y < P.h;
  const int idx = y * P.w + x;

  float d = 0.f;
  float3 pos_cam = f3(0.f, 0.f, 1.f);
  float range = 1.f;
  bool valid = false;
  if (in_img) {
    // 16-bit sensor planes are converted here, with cv::Mat::convertTo's own arithmetic for CV_16U -> CV_32F
    // (float(pixel) * float(scale), one rounding; examples/tsdf/offline.cc:77-80) -- no separate conversion pass
    d = in.depth_u16 ? (float)static_cast<const unsigned short*>(in.depth)[idx] * in.depth_scale
                     : static_cast<const float*>(in.depth)[idx];
    // K^-1 * (x, y, 1)  -- utils/tsdf/voxel_tsdf.cu:118-120
    pos_cam = kmul(P.Kinv, f3((float)x, (float)y, 1.f));
    range = sqrtf(sqnorm3(pos_cam));
    valid = !(d == 0 || d > P.max_depth);
    float p_ht = 1.f, p_lt = 1.f;  // no probability planes: TSDFSystem's default of ones (modules/tsdf_module.cc:28-33)
    if (in.ht) {
      if (in.prob_u16) {
        p_ht = (float)static_cast<const unsigned short*>(in.ht)[idx] * in.prob_scale;
        p_lt = (float)static_cast<const unsigned short*>(in.lt)[idx] * in.prob_scale;
      } else {
        p_ht = static_cast<const float*>(in.ht)[idx];
        p_lt = static_cast<const float*>(in.lt)[idx];
      }
    }
    const float dlogit = logf(p_ht) - logf(p_lt);
    const uint32_t rgbx = (uint32_t)rgb[3 * idx] | ((uint32_t)rgb[3 * idx + 1] << 8) | ((uint32_t)rgb[3 * idx + 2] << 16);
    *reinterpret_cast<uint4*>(tex + idx) =
        make_uint4(__float_as_uint(valid ? d : 0.f), __float_as_uint(range), __float_as_uint(dlogit), rgbx);
  }
  // candidate exchange: the staging above is needed on every rank (the integrate kernel gathers from it), the ray walk
  // below only on the rank this 32 x 8 tile is dealt to
  constexpr bool exchange = EXCHANGE;
  if (exchange && (blockIdx.y * gridDim.x + blockIdx.x) % (unsigned)S.shard_count != (unsigned)S.shard_rank) return;
  __shared__ u64 s_set[EXCHANGE ? kXaSet : 1];
  __shared__ int s_cnt[8], s_base[8];
  if (exchange) {
    s_set[threadIdx.x] = kEmpty;
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
  }

  // ---- ray set-up, utils/tsdf/voxel_tsdf.cu:124-139 (divisions share one reciprocal per divisor, see div_by) ----
  int nsteps = 0;
  float3 pos_grid = f3(0.f, 0.f, 0.f), ray_step_grid = f3(0.f, 0.f, 0.f);
  if (valid) {
    const float3 pos_world = apply(P.world_T_cam, f3(pos_cam.x * d, pos_cam.y * d, pos_cam.z * d));
    float3 ray_dir_cam;
    if (div_safe(range)) { const float r = rcp_refined(range); ray_dir_cam = f3(div_by(pos_cam.x, range, r), div_by(pos_cam.y, range, r), div_by(pos_cam.z, range, r)); }
    else ray_dir_cam = f3(pos_cam.x / range, pos_cam.y / range, pos_cam.z / range);
    const float3 ray_dir_world = qrot(P.world_T_cam, ray_dir_cam);
    const float3 ray_start_world = f3(pos_world.x - ray_dir_world.x * P.truncation,
                                      pos_world.y - ray_dir_world.y * P.truncation,
                                      pos_world.z - ray_dir_world.z * P.truncation);
    float3 ray_dir_grid;
    if (div_safe(P.voxel_size)) {
      const float r = rcp_refined(P.voxel_size);
      ray_dir_grid = f3(div_by(ray_dir_world.x, P.voxel_size, r), div_by(ray_dir_world.y, P.voxel_size, r), div_by(ray_dir_world.z, P.voxel_size, r));
      pos_grid = f3(div_by(ray_start_world.x, P.voxel_size, r), div_by(ray_start_world.y, P.voxel_size, r), div_by(ray_start_world.z, P.voxel_size, r));
    } else {
      ray_dir_grid = f3(ray_dir_world.x / P.voxel_size, ray_dir_world.y / P.voxel_size, ray_dir_world.z / P.voxel_size);
      pos_grid = f3(ray_start_world.x / P.voxel_size, ray_start_world.y / P.voxel_size, ray_start_world.z / P.voxel_size);
    }
    const float two_t = 2 * P.truncation;
    const float3 ray_grid = f3(two_t * ray_dir_grid.x, two_t * ray_dir_grid.y, two_t * ray_dir_grid.z);
    const int step_grid =
        __float2int_rz(ceilf(fmaxf(fmaxf(fabsf(ray_grid.x), fabsf(ray_grid.y)), fabsf(ray_grid.z)) / kBlockLen));
    const float denom = fmaxf((float)step_grid, 1);
    if (div_safe(denom)) { const float r = rcp_refined(denom); ray_step_grid = f3(div_by(ray_grid.x, denom, r), div_by(ray_grid.y, denom, r), div_by(ray_grid.z, denom, r)); }
    else ray_step_grid = f3(ray_grid.x / denom, ray_grid.y / denom, ray_grid.z / denom);
    nsteps = step_grid + 1;  // for (i = 0; i <= step_grid; ++i)
  }

  const int max_steps = __reduce_max_sync(0xFFFFFFFFu, nsteps);
  const bool sharded = S.shard_count > 1;
  int n_cand = 0, n_new = 0;
  if (max_steps <= kInlineSteps) {
    // Common case (truncation / voxel_size = 6 gives <= 3 samples): all samples' block coordinates first, then the
    // first table look of every distinct candidate in flight together, then the (rare) follow-ups -- one round trip
    // to the table per pixel instead of one per sample.
    u64 key[kInlineSteps], first[kInlineSteps];
    bool lead[kInlineSteps];
    u64 prev_key = kEmpty;
#pragma unroll
    for (int i = 0; i < kInlineSteps; ++i) {
      key[i] = kEmpty;  // sentinel: nothing to do for this lane
      if (i < nsteps) {
        const int px = round_to_voxel(pos_grid.x), py = round_to_voxel(pos_grid.y), pz = round_to_voxel(pos_grid.z);
        key[i] = pack_key(px >> 3, py >> 3, pz >> 3);
        pos_grid = f3(pos_grid.x + ray_step_grid.x, pos_grid.y + ray_step_grid.y, pos_grid.z + ray_step_grid.z);
        if (key[i] == prev_key) key[i] = kEmpty; else prev_key = key[i];
      }
    }
#pragma unroll
    for (int i = 0; i < kInlineSteps; ++i) {
      lead[i] = false; first[i] = kEmpty;
      if (i < max_steps) {  // warp-uniform
        // warp-cooperative de-duplication: one lane per distinct block coordinate probes the table
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, key[i]);
        lead[i] = (key[i] != kEmpty) && ((unsigned)(__ffs(peers) - 1) == lane) &&
                  (exchange || !sharded || owner_of(key[i], S.shard_count, S.shard_shift) == (unsigned)S.shard_rank);
        if (exchange) {  // the owner looks the key up after the barrier
          if (lead[i]) { xa_collect(S, s_set, key[i]); ++n_cand; }
          lead[i] = false;
        }
        // L1-cached look at the home slot: a key seen here IS present (keys only appear during this kernel, and L1
        // does not outlive a kernel); anything else is settled by the coherent probe below
        if (lead[i]) first[i] = ld_key_ca(S.table + (hash_key(key[i]) & S.table_mask));
      }
    }
#pragma unroll
    for (int i = 0; i < kInlineSteps; ++i) {
      if (i < max_steps) {  // warp-uniform
        // Allocate() is a no-op for present blocks (voxel_hash.cu:62-77), so the all-corners visibility test of
        // voxel_tsdf.cu:144 runs only for absent ones -- and then as a warp: one corner per lane, one ballot,
        // instead of eight serial projections on the one lane that found the block missing
        bool absent = false;
        if (lead[i]) { ++n_cand; absent = first[i] != key[i] && !table_contains(S, key[i]); }
        for (unsigned todo = __ballot_sync(0xFFFFFFFFu, absent); todo; todo &= todo - 1) {
          const int src = __ffs(todo) - 1;
          const u64 k = __shfl_sync(0xFFFFFFFFu, key[i], src);
          int bx, by, bz; unpack_key(k, bx, by, bz);
          const int cx = (short)((short)(bx << 3) + ((lane >> 0) & 1) * (kBlockLen - 1));
          const int cy = (short)((short)(by << 3) + ((lane >> 1) & 1) * (kBlockLen - 1));
          const int cz = (short)((short)(bz << 3) + ((lane >> 2) & 1) * (kBlockLen - 1));
          const unsigned vis = __ballot_sync(0xFFFFFFFFu, voxel_visible(cx, cy, cz, P));
          if ((vis & 0xFFu) == 0xFFu && (int)lane == src && table_insert(S, k) == 1) ++n_new;
        }
      }
    }
  } else {
    u64 prev_key = kEmpty;
    for (int i = 0; i < max_steps; ++i) {
      u64 key = kEmpty;
      if (i < nsteps) {
        const int px = round_to_voxel(pos_grid.x), py = round_to_voxel(pos_grid.y), pz = round_to_voxel(pos_grid.z);
        key = pack_key(px >> 3, py >> 3, pz >> 3);
        pos_grid = f3(pos_grid.x + ray_step_grid.x, pos_grid.y + ray_step_grid.y, pos_grid.z + ray_step_grid.z);
        if (key == prev_key) key = kEmpty; else prev_key = key;
      }
      const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
      const bool leader = (key != kEmpty) && ((unsigned)(__ffs(peers) - 1) == lane);
      if (exchange) {
        if (leader) { xa_collect(S, s_set, key); ++n_cand; }
        continue;
      }
      if (leader && (!sharded || owner_of(key, S.shard_count, S.shard_shift) == (unsigned)S.shard_rank)) {
        ++n_cand;
        if (!table_contains(S, key)) {
          int bx, by, bz; unpack_key(key, bx, by, bz);
          if (block_visible<true>(bx, by, bz, P) && table_insert(S, key) == 1) ++n_new;
        }
      }
    }
  }
  if (exchange) xa_flush(S, s_set, s_cnt, s_base);
  // counters: one atomic per warp
  n_cand = __reduce_add_sync(0xFFFFFFFFu, n_cand);
  n_new = __reduce_add_sync(0xFFFFFFFFu, n_new);
  if (lane == 0) {
    if (n_cand) atomicAdd(&S.ctr[C_NCAND], n_cand);
    if (n_new) atomicAdd(&S.ctr[C_NNEW], n_new);
  }
}

// ------------------------------------------------------------------------------------------
// insert_candidates_kernel: the owner's half of the candidate exchange.  blockIdx.y = sending rank; its keys for this
// frame sit in this rank's inbox (complete and visible: a barrier over the ranks lies between the senders' allocate
// kernels and this launch).  Same rule as above: a present block is left alone, an absent one is inserted if all eight
// corners project into the image (voxel_tsdf.cu:144) -- tested by eight lanes at once.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) insert_candidates_kernel(DeviceState S, FrameParams P) {
  const unsigned lane = threadIdx.x & 31;
  const int src = blockIdx.y;
  const unsigned char* inbox = reinterpret_cast<const unsigned char*>(S.xa_inbox[S.shard_rank]);
  const int n = min(reinterpret_cast<const int*>(inbox)[S.xa_parity * 8 + src], S.xa_cap);
  const u64* keys = reinterpret_cast<const u64*>(inbox + kXaHeaderBytes) + (size_t)(S.xa_parity * S.shard_count + src) * S.xa_cap;
  int n_new = 0;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + (int)lane;
    u64 key = kEmpty;
    bool absent = false;
    if (i < n) { key = keys[i]; absent = !table_contains(S, key); }
    for (unsigned todo = __ballot_sync(0xFFFFFFFFu, absent); todo; todo &= todo - 1) {
      const int from = __ffs(todo) - 1;
      const u64 k = __shfl_sync(0xFFFFFFFFu, key, from);
      int bx, by, bz; unpack_key(k, bx, by, bz);
      const int cx = (short)((short)(bx << 3) + ((lane >> 0) & 1) * (kBlockLen - 1));
      const int cy = (short)((short)(by << 3) + ((lane >> 1) & 1) * (kBlockLen - 1));
      const int cz = (short)((short)(bz << 3) + ((lane >> 2) & 1) * (kBlockLen - 1));
      const unsigned vis = __ballot_sync(0xFFFFFFFFu, voxel_visible(cx, cy, cz, P));
      if ((vis & 0xFFu) == 0xFFu && (int)lane == from && table_insert(S, k) == 1) ++n_new;
    }
  }
  n_new = __reduce_add_sync(0xFFFFFFFFu, n_new);
  if (lane == 0 && n_new) atomicAdd(&S.ctr[C_NNEW], n_new);
  // this frame's cursors have been published (by the barrier kernel, earlier on this stream): clear them for the frame
  // after next, which uses this half again
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 8) S.xa_cursor[S.xa_parity * 8 + threadIdx.x] = 0;
}

// ------------------------------------------------------------------------------------------
// select_visible_kernel: any-corner visibility (voxel_tsdf.cu:82-93) over the dense pool
// directory block_key[0 .. high_water) -- 8 B per pool block instead of the reference's scan of
// all 2^22 hash entries + 3-launch prefix sum + compaction + host sync.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_visible_kernel(DeviceState S, FrameParams P, int* __restrict__ visible,
                                                             int* __restrict__ vis_state) {
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
      if (vis) {
        const int o = off + __popc(m & ((1u << lane) - 1));
        visible[o] = i;
        vis_state[o] = 0;  // items of the block completed so far (+ 256 x items whose voxels are all carve-eligible)
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// integrate_carve_kernel: persistent warps over a device-side queue of WORK ITEMS = (visible block, group of
// SLABS 128-voxel slabs).  A lane owns 4 consecutive-x voxels of a slab, so every voxel-plane access is one 16-byte
// LDG/STG and the warp covers 512 contiguous bytes per plane.  A slab's three planes are requested before its
// projection arithmetic, its pixel gathers are issued pair by pair (the second pair's projection covers the first
// pair's latency), and the next item's directory entry is fetched one item ahead.  The first item of every warp is
// assigned statically (no start-up burst on the queue counter).  SLABS = 4 (whole blocks) is what runs: with only
// ~2.5 visible blocks per resident warp finer items would shorten the tail, but measured slower (see the launcher).
//   per voxel: voxel_tsdf.cu:157-203 (projection, nearest pixel, SDF, truncation, weighted running averages,
//   weight clamp); per block: voxel_tsdf.cu:214-229 (min |tsdf| >= .9 -> free): every item adds
//   1 + 256 * [its voxels are all >= .9] to the block's word of `vis_state`; the item that completes the block
//   sees the verdict of all of them and erases the block or clears its "new" flag.
// Fusions: blocks acquired this frame are initialised in registers (no init pass, no read); the carve reduction
// reuses the just-computed values; the per-pixel inputs arrive in one 16-byte gather; divisions share one refined
// reciprocal per divisor (div_by, bit-identical to `/`).  Operands outside the range the shared-reciprocal sequence
// is proven for (a voxel on / behind the camera plane, a vanishing combined weight) are flagged and redone with true
// divisions in one out-of-line loop per slab, so the hot path has no per-voxel slow-path code.
// FAST = truncation and max_depth are themselves safe divisors; otherwise every voxel takes the true-division code.
// ------------------------------------------------------------------------------------------
#ifndef INTEGRATE_MIN_CTAS
#define INTEGRATE_MIN_CTAS 4  // 64 registers -> 32 resident warps per SM (measured faster than 80 registers / 24 warps)
#endif
__device__ __forceinline__ float4 ld16(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint4 ld16u(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st16u(uint32_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// (int)roundf(x) for 0 <= x < 2^23 (and NaN -> 0, like cvt.rzi of roundf(NaN)): what roundf itself does
__device__ __forceinline__ unsigned round_nonneg(float x) { return (unsigned)__float2int_rz(__fadd_rz(x, 0.5f)); }
// volatile: keeps the gather where it is written (ptxas otherwise sinks the four gathers of a slab behind the last projection)
__device__ __forceinline__ uint4 ldg_texel(const Texel* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int pixel_index(float uq, float vq, const FrameParams& P) {  // voxel_tsdf.cu:165-169
  const int u = __float2int_rz(roundf(uq)), v = __float2int_rz(roundf(vq));
  return ((unsigned)u < (unsigned)P.w && (unsigned)v < (unsigned)P.h) ? v * P.w + u : -1;
}

// voxel_tsdf.cu:157-166 with true divisions (fallback of the shared-reciprocal projection)
__device__ __forceinline__ int project_exact(const FrameParams& P, int gx, int gy, int gz, float& z_cam) {
  const float3 pc = apply(P.cam_T_world, f3((float)gx * P.voxel_size, (float)gy * P.voxel_size, (float)gz * P.voxel_size));
  const float3 ph = kmul(P.K, pc);
  z_cam = pc.z;
  return pixel_index(ph.x / ph.z, ph.y / ph.z, P);
}
// voxel_tsdf.cu:178-202 with true divisions (e.g. depth == max_depth on a fresh voxel: 0 / 0 like the reference)
__device__ __forceinline__ void update_exact(const FrameParams& P, float sdf, uint4 px, float& tsdf, uint32_t& rgbw, float& logit) {
  const float depth = __uint_as_float(px.x);
  const float tsdf_new = fminf(1, sdf / P.truncation);
  const float weight_new = (1 - depth / P.max_depth) * 4;
  const float weight_old = (float)(rgbw >> 24);
  const float weight_combined = weight_old + weight_new;
  const float r_old = (float)(rgbw & 0xFF), g_old = (float)((rgbw >> 8) & 0xFF), b_old = (float)((rgbw >> 16) & 0xFF);
  const float r_new = (float)(px.w & 0xFF), g_new = (float)((px.w >> 8) & 0xFF), b_new = (float)((px.w >> 16) & 0xFF);
  const unsigned r = (unsigned)__float2int_rz(roundf((r_old * weight_old + r_new * weight_new) / weight_combined)),
                 g = (unsigned)__float2int_rz(roundf((g_old * weight_old + g_new * weight_new) / weight_combined)),
                 b = (unsigned)__float2int_rz(roundf((b_old * weight_old + b_new * weight_new) / weight_combined));
  tsdf = (tsdf * weight_old + tsdf_new * weight_new) / weight_combined;
  logit = (logit * weight_old + __uint_as_float(px.z) * weight_new) / weight_combined;
  const unsigned w = (unsigned)__float2int_rz(fminf(roundf(weight_combined), 40));
  rgbw = min(r, 255u) | (min(g, 255u) << 8) | (min(b, 255u) << 16) | (w << 24);
}

template <int SLABS, bool FAST, bool MIRROR>
__global__ void __launch_bounds__(256, INTEGRATE_MIN_CTAS) integrate_carve_kernel(DeviceState S, FrameParams P,
                                                              const int* __restrict__ visible, int* __restrict__ vis_state,
                                                              const Texel* __restrict__ tex, float carve_threshold) {
  constexpr int kItemsPerBlock = 4 / SLABS;
  const unsigned lane = threadIdx.x & 31;
  const int n_items = S.ctr[C_NVIS] * kItemsPerBlock;
  const int total_warps = gridDim.x * (blockDim.x >> 5);
  const int vx0 = (lane & 1) * 4, vy = (lane >> 1) & 7, vz_lo = lane >> 4;  // voxel (slab * 32 + lane) * 4 .. + 3
  const float r_trunc = rcp_refined(P.truncation), r_md = rcp_refined(P.max_depth);
  const float qx = P.cam_T_world.qx, qy = P.cam_T_world.qy, qz = P.cam_T_world.qz, qw = P.cam_T_world.qw;
  unsigned n_upd_thread = 0;
  int n_carved_thread = 0;

  // first item: static; later items from the queue counter, fetched (with the directory entry) one item ahead
  int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int idx = 0; u64 bk = 0;
  if (it < n_items) { idx = visible[it / kItemsPerBlock]; bk = S.block_key[idx]; }

  while (it < n_items) {
    int it_next = 0;
    if (lane == 0) it_next = total_warps + atomicAdd(&S.ctr[C_WORK], 1);
    it_next = __shfl_sync(0xFFFFFFFFu, it_next, 0);
    int idx_next = 0; u64 bk_next = 0;
    if (it_next < n_items) { idx_next = visible[it_next / kItemsPerBlock]; bk_next = S.block_key[idx_next]; }

    const bool is_new = (bk & kFlagNew) != 0;
    int bx, by, bz; unpack_key(bk, bx, by, bz);
    const int slab0 = (it % kItemsPerBlock) * SLABS;
    float* const base_tsdf = block_tsdf(S, idx) + lane * 4;
    uint32_t* const base_rgbw = block_rgbw(S, idx) + lane * 4;
    float* const base_logit = block_logit(S, idx) + lane * 4;

    float item_min = __int_as_float(0x7FFFFFFF);  // NaN: fminf's identity; a block of nothing but NaN voxels is not carved (NaN >= .9 is false), like the reference
    const int gy = (short)((by << 3) + vy);
    const int gx0 = (bx << 3) + vx0;
    const float wy = (float)gy * P.voxel_size;

#pragma unroll 1
    for (int slab = slab0; slab < slab0 + SLABS; ++slab) {
      // the three planes of this slab: requested now, first needed after the projection + gather below
      float4 c_tsdf = make_float4(-1.f, -1.f, -1.f, -1.f);  // voxel_mem.cu:48-50: tsdf -1, weight 0 (rgb := 0), p .5
      uint4 c_rgbw = make_uint4(0u, 0u, 0u, 0u);
      float4 c_logit = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!is_new) { c_tsdf = ld16(base_tsdf + slab * 128); c_rgbw = ld16u(base_rgbw + slab * 128); c_logit = ld16(base_logit + slab * 128); }
      const int gz = (short)((bz << 3) + slab * 2 + vz_lo);
      float pcz[4];
      int pix[4];
      uint4 px[4];  // the voxels' pixels (depth 0 <=> nothing to do); each gather is issued as soon as its address is known
      if (FAST) {
        // ---- SE3::Apply in Eigen's order (qrot, tsdf_device.cuh), x-independent part hoisted ----
        const float wz = (float)gz * P.voxel_size;
        const float uvx = qy * wz - qz * wy;          // uv = q.vec x v
        const float uvx2 = uvx + uvx;                 // uv += uv
        const float qx_wz = qx * wz, qx_wy = qx * wy;
        const float qw_uvx2 = qw * uvx2, qz_uvx2 = qz * uvx2, qy_uvx2 = qy * uvx2;
        // ---- phase 1: project the 4 voxels (voxel_tsdf.cu:157-166), two at a time: the gathers of a pair are issued
        // before the next pair is projected, so their latency is covered by that arithmetic (the range test between
        // the pairs is also what keeps ptxas from regrouping the four gathers behind the last projection) ----
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
#pragma unroll
          for (int j = h; j < h + 2; ++j) {
            const int gx = (short)(gx0 + j);
            const float wx = (float)gx * P.voxel_size;
            const float uvy = qz * wx - qx_wz, uvz = qx_wy - qy * wx;
            const float uvy2 = uvy + uvy, uvz2 = uvz + uvz;
            const float cx = qy * uvz2 - qz * uvy2;      // q.vec x uv
            const float cy = qz_uvx2 - qx * uvz2;
            const float cz = qx * uvy2 - qy_uvx2;
            const float pcx = ((wx + qw_uvx2) + cx) + P.cam_T_world.tx;
            const float pcy = ((wy + qw * uvy2) + cy) + P.cam_T_world.ty;
            pcz[j] = ((wz + qw * uvz2) + cz) + P.cam_T_world.tz;
            const float hx = P.K.fx * pcx + P.K.cx * pcz[j], hy = P.K.fy * pcy + P.K.cy * pcz[j];  // kmul
            const float r = rcp_refined(pcz[j]);
            pix[j] = pixel_index(div_by(hx, pcz[j], r), div_by(hy, pcz[j], r), P);
          }
          // a voxel on / behind the camera plane (or absurdly far): the reciprocal sequence is not proven exact there,
          // so such a voxel is projected again with true divisions -- cold code, one range test per pair on the hot path
          if (!(div_safe(fminf(pcz[h], pcz[h + 1])) && div_safe(fmaxf(pcz[h], pcz[h + 1])))) {
#pragma unroll
            for (int j = h; j < h + 2; ++j)
              if (!div_safe(pcz[j])) pix[j] = project_exact(P, (short)(gx0 + j), gy, gz, pcz[j]);
          }
#pragma unroll
          for (int j = h; j < h + 2; ++j) {
            px[j] = make_uint4(0u, 0u, 0u, 0u);
            if (pix[j] >= 0) px[j] = ldg_texel(tex + pix[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pix[j] = project_exact(P, (short)(gx0 + j), gy, gz, pcz[j]);
          px[j] = make_uint4(0u, 0u, 0u, 0u);
          if (pix[j] >= 0) px[j] = ldg_texel(tex + pix[j]);
        }
      }
      // ---- phase 2: update (voxel_tsdf.cu:167-203) ----
      float tsdf[4] = {c_tsdf.x, c_tsdf.y, c_tsdf.z, c_tsdf.w};
      uint32_t rgbw[4] = {c_rgbw.x, c_rgbw.y, c_rgbw.z, c_rgbw.w};
      float logit[4] = {c_logit.x, c_logit.y, c_logit.z, c_logit.w};
      unsigned upd = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float depth = __uint_as_float(px[j].x);
        if (depth != 0.f) {  // depth == 0 || depth > max_depth folded into the staging
          const float sdf = __uint_as_float(px[j].y) * (depth - pcz[j]);
          if (sdf > -P.truncation) {
            upd |= 1u << j;
            if (FAST) {
              const float tsdf_new = fminf(1, div_by(sdf, P.truncation, r_trunc));
              const float weight_new = (1 - div_by(depth, P.max_depth, r_md)) * 4;  // :182
              const float weight_old = (float)(rgbw[j] >> 24);
              const float weight_combined = weight_old + weight_new;
              if (div_safe(weight_combined)) {
                const float r_old = (float)(rgbw[j] & 0xFF), g_old = (float)((rgbw[j] >> 8) & 0xFF), b_old = (float)((rgbw[j] >> 16) & 0xFF);
                const float r_new = (float)(px[j].w & 0xFF), g_new = (float)((px[j].w >> 8) & 0xFF), b_new = (float)((px[j].w >> 16) & 0xFF);
                // voxel_tsdf.cu:186-202 (semantic fusion in logit space)
                const float num_r = r_old * weight_old + r_new * weight_new, num_g = g_old * weight_old + g_new * weight_new,
                            num_b = b_old * weight_old + b_new * weight_new;
                const float num_t = tsdf[j] * weight_old + tsdf_new * weight_new;
                const float num_l = logit[j] * weight_old + __uint_as_float(px[j].z) * weight_new;
                const float rw = rcp_refined(weight_combined);
                const unsigned r = round_nonneg(div_by(num_r, weight_combined, rw)), g = round_nonneg(div_by(num_g, weight_combined, rw)),
                               bb = round_nonneg(div_by(num_b, weight_combined, rw));
                tsdf[j] = div_by(num_t, weight_combined, rw);
                logit[j] = div_by(num_l, weight_combined, rw);
                rgbw[j] = r | (g << 8) | (bb << 16) | (min(round_nonneg(weight_combined), 40u) << 24);
              } else {
                upd |= 16u << j;  // vanishing combined weight: redone below with true divisions
              }
            } else {
              update_exact(P, sdf, px[j], tsdf[j], rgbw[j], logit[j]);
            }
          }
        }
      }
      if (FAST && (upd >> 4)) {  // cold
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if ((upd >> (4 + j)) & 1)
            update_exact(P, __uint_as_float(px[j].y) * (__uint_as_float(px[j].x) - pcz[j]), px[j], tsdf[j], rgbw[j], logit[j]);
      }
      n_upd_thread += __popc(upd & 15u);
      item_min = fminf(item_min, fminf(fminf(fabsf(tsdf[0]), fabsf(tsdf[1])), fminf(fabsf(tsdf[2]), fabsf(tsdf[3]))));
      if ((upd & 15u) || is_new) {  // a block that turns out to be carved is released anyway: writing it is harmless
        st16(base_tsdf + slab * 128, make_float4(tsdf[0], tsdf[1], tsdf[2], tsdf[3]));
        if (MIRROR && S.n_mirror) {  // (its own instantiation: the single-GPU kernel carries none of this) sharded volume: the same 16 bytes into every rank's TSDF mirror (posted stores over NVLink)
          const size_t slot = ((size_t)S.shard_rank * S.mirror_stride + idx) * kBlockVolume + lane * 4 + slab * 128;
          for (int r = 0; r < S.n_mirror; ++r) st16(S.mirror[r] + slot, make_float4(tsdf[0], tsdf[1], tsdf[2], tsdf[3]));
        }
        st16u(base_rgbw + slab * 128, make_uint4(rgbw[0], rgbw[1], rgbw[2], rgbw[3]));
        st16(base_logit + slab * 128, make_float4(logit[0], logit[1], logit[2], logit[3]));
      }
    }

    // ---- space carving (voxel_tsdf.cu:214-229): min |tsdf| over the 512 voxels ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) item_min = fminf(item_min, __shfl_xor_sync(0xFFFFFFFFu, item_min, o));
    if (lane == 0) {
      const int far = item_min >= carve_threshold ? 1 : 0;
      int n_done = kItemsPerBlock, n_far = far;
      if (kItemsPerBlock > 1) {
        const int old = atomicAdd(&vis_state[it / kItemsPerBlock], 1 + 256 * far);
        n_done = (old & 0xFF) + 1; n_far = (old >> 8) + far;
      }
      if (n_done == kItemsPerBlock) {  // this item completed the block
        if (n_far == kItemsPerBlock) { table_erase(S, bk & kKeyMask); ++n_carved_thread; if (!is_new) mark_block_set_changed(S); }
        else if (is_new) { S.block_key[idx] = bk & kKeyMask; mark_block_set_changed(S); }
      }
    }
    it = it_next; idx = idx_next; bk = bk_next;
  }

  // ---- counters: one atomic per warp ----
  n_upd_thread = __reduce_add_sync(0xFFFFFFFFu, n_upd_thread);
  if (lane == 0) {
    if (n_upd_thread) atomicAdd(reinterpret_cast<u64*>(&S.ctr[C_NUPD_LO]), (u64)n_upd_thread);
    if (n_carved_thread) atomicAdd(&S.ctr[C_NCARVED], n_carved_thread);
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void launch_frame_allocate(const DeviceState& S, const FrameParams& P, const FrameInput& in, Texel* tex, cudaStream_t st) {
  const dim3 grid((P.w + 31) / 32, (P.h + 7) / 8);
  if (S.xa_on) frame_allocate_kernel<true><<<grid, 256, 0, st>>>(S, P, in, tex);
  else frame_allocate_kernel<false><<<grid, 256, 0, st>>>(S, P, in, tex);
}
void launch_insert_candidates(const DeviceState& S, const FrameParams& P, cudaStream_t st) {
  insert_candidates_kernel<<<dim3(32, S.shard_count), 256, 0, st>>>(S, P);
}
void launch_select_visible(const DeviceState& S, const FrameParams& P, int* visible, int* vis_state, int num_sms, cudaStream_t st) {
  select_visible_kernel<<<num_sms * 4, 256, 0, st>>>(S, P, visible, vis_state);
}
static bool div_safe_host(float b) { return b > 9.5367431640625e-07f && b < 1048576.f; }  // div_safe of tsdf_device.cuh
template <int SLABS, bool FAST, bool MIRROR>
static void launch_integrate_mirror_variant(const DeviceState& S, const FrameParams& P, const int* visible, int* vis_state, const Texel* tex,
                                     int num_sms, cudaStream_t st) {
  // persistent warps: exactly the resident CTAs, work comes from the device-side queue (engines on several host threads
  // may race to fill this in: they all compute the same value)
  static std::atomic<int> cached{0};
  int ctas_per_sm = cached.load(std::memory_order_relaxed);
  if (ctas_per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, integrate_carve_kernel<SLABS, FAST, MIRROR>, 256, 0) != cudaSuccess || ctas_per_sm < 1)
      ctas_per_sm = 2;
    cached.store(ctas_per_sm, std::memory_order_relaxed);
  }
  integrate_carve_kernel<SLABS, FAST, MIRROR><<<num_sms * ctas_per_sm, 256, 0, st>>>(S, P, visible, vis_state, tex, .9f);
}
template <int SLABS, bool FAST>
static void launch_integrate_variant(const DeviceState& S, const FrameParams& P, const int* visible, int* vis_state, const Texel* tex,
                                     int num_sms, cudaStream_t st) {
  if (S.n_mirror) launch_integrate_mirror_variant<SLABS, FAST, true>(S, P, visible, vis_state, tex, num_sms, st);
  else launch_integrate_mirror_variant<SLABS, FAST, false>(S, P, visible, vis_state, tex, num_sms, st);
}
void launch_integrate_carve(const DeviceState& S, const FrameParams& P, const int* visible, int* vis_state, const Texel* tex,
                            int num_sms, int expected_blocks, cudaStream_t st) {
  // Work-item granularity.  Measured on B200 (config 2, ~11.7 k visible blocks for 4736 resident warps): whole blocks
  // (4 slabs per item) 46.0 us, 2 slabs 47.9 us, 1 slab 54.2 us -- with more than two blocks per warp the per-item
  // set-up costs more than the shorter tail saves.  A shard of a volume spread over many GPUs sees far fewer blocks
  // than there are resident warps; then a warp's single block IS the kernel's duration and finer items shorten it.
  // expected_blocks = visible blocks of the most recent finished frame (0 = unknown); the result does not depend on it.
  const bool fast = div_safe_host(P.truncation) && div_safe_host(P.max_depth);
  const int warps = num_sms * 32;  // resident warps at 64 registers per thread
  const int slabs = expected_blocks <= 0 || expected_blocks > warps ? 4 : expected_blocks > warps / 2 ? 2 : 1;
  if (slabs == 4) {
    if (fast) launch_integrate_variant<4, true>(S, P, visible, vis_state, tex, num_sms, st);
    else launch_integrate_variant<4, false>(S, P, visible, vis_state, tex, num_sms, st);
  } else if (slabs == 2) {
    if (fast) launch_integrate_variant<2, true>(S, P, visible, vis_state, tex, num_sms, st);
    else launch_integrate_variant<2, false>(S, P, visible, vis_state, tex, num_sms, st);
  } else {
    if (fast) launch_integrate_variant<1, true>(S, P, visible, vis_state, tex, num_sms, st);
    else launch_integrate_variant<1, false>(S, P, visible, vis_state, tex, num_sms, st);
  }
}

}  // namespace tsdf
