// kernels_gather.cu -- state init, GatherValid / GatherVoxels, export and the hash/pool unit-test
// kernels of the B200 TSDF engine (sm_100a).
//
// Replaces check_valid_kernel / check_bound_kernel / download_tsdf_kernel + the 3-launch prefix
// sum + compaction + cudaMalloc/cudaFree per call (utils/tsdf/voxel_tsdf.cu:14-46,399-472) by a
// scan of the dense pool directory and a fully coalesced 16-byte-per-lane emit.
#include "tsdf_device.cuh"
#include "tsdf_launch.h"

namespace tsdf {

// init_hash_table_kernel + heap_init_kernel (voxel_hash.cu:26-29, voxel_mem.cu:6-11).  The free
// stack hands out low indices first so that the directory scan stays short (high-water mark).
__global__ void init_state_kernel(DeviceState S) {
  const size_t n_slots = (size_t)S.table_mask + 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
    Slot s; s.key = kEmpty; s.val = -1; s.pad = 0;
    S.table[i] = s;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)S.pool_blocks; i += stride) {
    S.block_key[i] = kEmpty;
    S.free_stack[i] = S.pool_blocks - 1 - (int)i;
  }
  if (blockIdx.x == 0 && threadIdx.x < C_COUNT) {
    const int c = threadIdx.x;
    S.ctr[c] = c == C_FREE ? S.pool_blocks : (c >= C_MIN_X && c <= C_MIN_Z) ? 0x7FFFFFFF : (c >= C_MAX_X && c <= C_MAX_Z) ? -0x7FFFFFFF : 0;
  }
}

// check_valid_kernel / check_bound_kernel (voxel_tsdf.cu:14-32) + compaction
__global__ void __launch_bounds__(256) select_blocks_kernel(DeviceState S, bool use_bound, GridBound g,
                                                            int* __restrict__ selected) {
  const int hw = S.ctr[C_HIGH_WATER];
  const unsigned lane = threadIdx.x & 31;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < hw; base += gridDim.x * blockDim.x) {
    const int i = base + lane;
    bool sel = false;
    if (i < hw) {
      const u64 k = S.block_key[i];
      if (k != kEmpty) {
        sel = true;
        if (use_bound) {
          int bx, by, bz; unpack_key(k, bx, by, bz);
          const int x = (short)(bx << 3), y = (short)(by << 3), z = (short)(bz << 3);
          sel = (x >= g.xmin && y >= g.ymin && z >= g.zmin && x + kBlockLen - 1 <= g.xmax &&
                 y + kBlockLen - 1 <= g.ymax && z + kBlockLen - 1 <= g.zmax);
        }
      }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
    if (m) {
      int off = 0;
      if (lane == 0) off = atomicAdd(&S.ctr[C_NSEL], __popc(m));
      off = __shfl_sync(0xFFFFFFFFu, off, 0);
      if (sel) selected[off + __popc(m & ((1u << lane) - 1))] = i;
    }
  }
}

// download_tsdf_kernel (voxel_tsdf.cu:34-46): {(grid) * voxel_size, tsdf} for all 512 voxels.
// Lane k emits voxels k, k+128, k+256, k+384: 4-byte coalesced reads, 16-byte coalesced writes.
__global__ void __launch_bounds__(128) download_voxels_kernel(DeviceState S, const int* __restrict__ selected,
                                                              float voxel_size, float4* __restrict__ out) {
  const int idx = selected[blockIdx.x];
  int bx, by, bz; unpack_key(S.block_key[idx], bx, by, bz);
  const float* tsdf = block_tsdf(S, idx);
  float4* dst = out + (size_t)blockIdx.x * kBlockVolume;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = threadIdx.x + r * 128;
    const int gx = (short)((bx << 3) + (k & 7)), gy = (short)((by << 3) + ((k >> 3) & 7)), gz = (short)((bz << 3) + (k >> 6));
    dst[k] = make_float4((float)gx * voxel_size, (float)gy * voxel_size, (float)gz * voxel_size, tsdf[k]);
  }
}

// parity export: raw voxel planes of the selected blocks, probability converted from the logit
__global__ void __launch_bounds__(128) export_blocks_kernel(DeviceState S, const int* __restrict__ selected,
                                                            short* __restrict__ keys, float* __restrict__ tsdf,
                                                            unsigned* __restrict__ rgbw, float* __restrict__ prob) {
  const int idx = selected[blockIdx.x];
  const u64 bk = S.block_key[idx];
  if (threadIdx.x == 0 && keys) {
    int bx, by, bz; unpack_key(bk, bx, by, bz);
    keys[3 * blockIdx.x] = (short)bx; keys[3 * blockIdx.x + 1] = (short)by; keys[3 * blockIdx.x + 2] = (short)bz;
  }
  const size_t o = (size_t)blockIdx.x * kBlockVolume;
  for (int k = threadIdx.x; k < kBlockVolume; k += blockDim.x) {
    if (tsdf) tsdf[o + k] = block_tsdf(S, idx)[k];
    if (rgbw) rgbw[o + k] = block_rgbw(S, idx)[k];
    if (prob) prob[o + k] = logit_to_prob(block_logit(S, idx)[k]);
  }
}

// ---- unit-test kernels mirroring utils/tests/voxel_hash_test.cu:36-55 --------------------------
__device__ __forceinline__ void init_block(const DeviceState& S, int idx) {  // voxel_mem.cu:43-51
  for (int k = 0; k < kBlockVolume; ++k) { block_tsdf(S, idx)[k] = -1.f; block_rgbw(S, idx)[k] = 0u; block_logit(S, idx)[k] = 0.f; }
}
__global__ void allocate_list_kernel(DeviceState S, const short* __restrict__ keys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 key = pack_key(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]);
  if (S.shard_count > 1 && owner_of(key, S.shard_count, S.shard_shift) != (unsigned)S.shard_rank) return;
  if (table_insert(S, key) == 1) { atomicAdd(&S.ctr[C_NNEW], 1); mark_block_set_changed(S); }
}
// blocks inserted outside Integrate are materialised immediately (the integrate kernel normally
// does that in registers): clear the NEW flag and write the acquire-time defaults
__global__ void materialise_new_kernel(DeviceState S) {
  const int hw = S.ctr[C_HIGH_WATER];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const u64 k = S.block_key[i];
    if (k != kEmpty && (k & kFlagNew)) { init_block(S, i); S.block_key[i] = k & kKeyMask; }
  }
}
__global__ void delete_list_kernel(DeviceState S, const short* __restrict__ keys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (table_erase(S, pack_key(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]))) { atomicAdd(&S.ctr[C_NCARVED], 1); mark_block_set_changed(S); }
}
__global__ void retrieve_list_kernel(DeviceState S, const short* __restrict__ pts, int n, float* tsdf, unsigned* rgbw,
                                     float* prob, int* found) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
  const int idx = table_find(S, pack_key(px >> 3, py >> 3, pz >> 3));
  const int k = voxel_index(px, py, pz);
  if (tsdf) tsdf[i] = idx >= 0 ? block_tsdf(S, idx)[k] : 1.f;
  if (rgbw) rgbw[i] = idx >= 0 ? block_rgbw(S, idx)[k] : 0u;
  if (prob) prob[i] = idx >= 0 ? logit_to_prob(block_logit(S, idx)[k]) : 0.f;
  if (found) found[i] = idx >= 0;
}
__global__ void assign_list_kernel(DeviceState S, const short* __restrict__ pts, int n, const float* tsdf,
                                   const unsigned* rgbw, const float* prob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
  const int idx = table_find(S, pack_key(px >> 3, py >> 3, pz >> 3));
  if (idx < 0) return;
  const int k = voxel_index(px, py, pz);
  if (tsdf) block_tsdf(S, idx)[k] = tsdf[i];
  if (rgbw) block_rgbw(S, idx)[k] = rgbw[i];
  if (prob) block_logit(S, idx)[k] = prob_to_logit(prob[i]);
}

// tombstone garbage collection: clear the table and re-insert every active block
__global__ void clear_table_kernel(DeviceState S) {
  const size_t n_slots = (size_t)S.table_mask + 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x) {
    Slot s; s.key = kEmpty; s.val = -1; s.pad = 0;
    S.table[i] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) S.ctr[C_NONEMPTY] = 0;
}
__global__ void reinsert_kernel(DeviceState S) {
  const int hw = S.ctr[C_HIGH_WATER];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const u64 bk = S.block_key[i];
    if (bk == kEmpty) continue;
    const u64 key = bk & kKeyMask;
    unsigned slot = hash_key(key) & S.table_mask;
    for (unsigned n = 0; n <= S.table_mask; ++n) {
      if (ld_key_cg(S.table + slot) == kEmpty &&
          atomicCAS(reinterpret_cast<u64*>(&S.table[slot].key), kEmpty, key) == kEmpty) {
        S.table[slot].val = i;
        atomicAdd(&S.ctr[C_NONEMPTY], 1);
        break;
      }
      slot = (slot + 1) & S.table_mask;
    }
  }
}

// The per-frame counters go to the host through a store into mapped pinned memory, not through a copy engine: a
// 128-byte cudaMemcpyAsync on the compute stream would queue behind the multi-megabyte image downloads of other
// streams on the device-to-host engine and hold up every kernel enqueued after it.
// The same kernel then clears the per-call counters for the next frame (a cudaMemsetAsync may also be serviced by a
// copy engine).
__global__ void publish_counters_kernel(int* __restrict__ ctr, volatile int* __restrict__ host) {
  if (threadIdx.x < C_COUNT) {
    host[threadIdx.x] = ctr[threadIdx.x];
    if (threadIdx.x >= C_PER_CALL) ctr[threadIdx.x] = 0;
  }
  __threadfence_system();
}

// ---- launchers -------------------------------------------------------------------------------
void launch_publish_counters(const DeviceState& S, int* host_mapped, cudaStream_t st) {
  publish_counters_kernel<<<1, 32, 0, st>>>(S.ctr, host_mapped);
}
void launch_init_state(const DeviceState& S, cudaStream_t st) { init_state_kernel<<<1024, 256, 0, st>>>(S); }
void launch_select_blocks(const DeviceState& S, bool use_bound, GridBound bound, int* selected, int num_sms,
                          cudaStream_t st) {
  select_blocks_kernel<<<num_sms * 4, 256, 0, st>>>(S, use_bound, bound, selected);
}
void launch_download_voxels(const DeviceState& S, const int* selected, int n_selected, float voxel_size, float4* out,
                            cudaStream_t st) {
  if (n_selected > 0) download_voxels_kernel<<<n_selected, 128, 0, st>>>(S, selected, voxel_size, out);
}
void launch_export_blocks(const DeviceState& S, const int* selected, int n_selected, short* keys, float* tsdf,
                          unsigned* rgbw, float* prob, cudaStream_t st) {
  if (n_selected > 0) export_blocks_kernel<<<n_selected, 128, 0, st>>>(S, selected, keys, tsdf, rgbw, prob);
}
void launch_allocate_list(const DeviceState& S, const short* keys, int n, cudaStream_t st) {
  if (n > 0) allocate_list_kernel<<<(n + 127) / 128, 128, 0, st>>>(S, keys, n);
  materialise_new_kernel<<<256, 256, 0, st>>>(S);
}
void launch_delete_list(const DeviceState& S, const short* keys, int n, cudaStream_t st) {
  if (n > 0) delete_list_kernel<<<(n + 127) / 128, 128, 0, st>>>(S, keys, n);
}
void launch_retrieve_list(const DeviceState& S, const short* points, int n, float* tsdf, unsigned* rgbw, float* prob,
                          int* found, cudaStream_t st) {
  if (n > 0) retrieve_list_kernel<<<(n + 127) / 128, 128, 0, st>>>(S, points, n, tsdf, rgbw, prob, found);
}
void launch_assign_list(const DeviceState& S, const short* points, int n, const float* tsdf, const unsigned* rgbw,
                        const float* prob, cudaStream_t st) {
  if (n > 0) assign_list_kernel<<<(n + 127) / 128, 128, 0, st>>>(S, points, n, tsdf, rgbw, prob);
}
void launch_rehash(const DeviceState& S, int num_sms, cudaStream_t st) {
  clear_table_kernel<<<num_sms * 4, 256, 0, st>>>(S);
  reinsert_kernel<<<num_sms * 4, 256, 0, st>>>(S);
}

}  // namespace tsdf
