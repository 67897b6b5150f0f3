// kernels_raycast.cu -- RayCast of the B200 TSDF engine (sm_100a).
//
// Replaces ray_cast_kernel (utils/tsdf/voxel_tsdf.cu:232-307): fixed-step march
// (step = truncation / 2), nearest-voxel lookups with a per-thread block cache (including the
// reference's negative cache for absent blocks, voxel_hash.cuh:124-161), sign-change hit test,
// bisection refinement, central-difference normal, Lambert shading and the semantic overlay.
//
// The sample positions are accumulated exactly like the reference (pos += step, float32, no FMA),
// so every sample the reference takes is taken at the bit-identical position here.  What is new is
// that samples which provably land in unallocated space are not looked up at all: a dense
// Chebyshev-distance map over the cells of the active-block AABB (built on the GPU whenever the
// block set changed, ~1 byte per 8^3..(8<<shift)^3 voxels, L1/L2 resident) says how many of the
// following samples cannot reach an allocated block; those samples only advance the position.
// Unallocated space reads TSDF = +1 in the reference (VoxelTSDF(), voxel_types.cu:8) and a +1
// sample can neither start nor complete a hit (voxel_tsdf.cu:260), so the result is identical.
// New output: per-ray hit depth (camera z, metres) and a packed (depth_bits << 32 | colour) key
// for nearest-hit min-compositing across GPUs.
#include <math_constants.h>

#include "tsdf_device.cuh"
#include "tsdf_launch.h"

namespace tsdf {

// ------------------------------------------------------------------------------------------
// skip-map construction (all sizes live on the device: no host round trip)
// ------------------------------------------------------------------------------------------
// Four launches: mark, then the three separable passes of the capped Chebyshev transform.
//
// Build attempts are numbered (gen) by the host.  With `lazy` set, attempt g first compares the serial of the last call
// that changed each shard's block set (ctr[C_DIRTY], written on NET changes only, see mark_block_set_changed; read
// over NVLink for foreign shards) with the serials recorded by attempt g - 1: all equal means the map is still exact
// and all four kernels return at once -- a frame that neither allocated nor carved a surviving block (a static
// camera, a converged scene), or a batch of views over a finished volume, costs four empty launches instead of a
// rebuild.  The records of consecutive attempts alternate between two halves of the header, so that no thread reads
// a word another thread of the same kernel writes; the verdict for the later kernels goes to hdr[10 + (g & 1)].
//
// No fill pass: the two byte planes swap roles at every rebuild, and the plane the NEXT rebuild will mark into is
// always entirely at the cap -- it is filled once at creation, and the last pass of every rebuild restores the cells
// it had dirtied (the plane is idle by then).  hdr[8 + (g & 1)] = number of rebuilds after attempt g; its parity
// before the attempt tells every kernel which plane plays which role.
struct Planes { unsigned char* mark; unsigned char* other; };
__device__ __forceinline__ Planes planes_of(const SkipMap& M, int gen) {
  const bool odd = M.hdr[8 + ((gen - 1) & 1)] & 1;
  Planes p; p.mark = odd ? M.scratch : M.dist; p.other = odd ? M.dist : M.scratch;
  return p;
}

// Every thread derives the same header from the AABB counters (a handful of integer operations), so that no separate
// one-thread launch is needed; thread 0 publishes it for the later kernels.
__global__ void __launch_bounds__(256) skip_mark_kernel(const PeerView* __restrict__ shards, int n_shards, SkipMap M, int gen, int lazy) {
  // The persistent counters of every shard (serial of the last block-set change, high-water mark, AABB) first, all at
  // once: for foreign shards each is a load over NVLink, and fetched one after the other inside the loops below they
  // cost one round trip per shard and loop -- ~25 us of pure latency on 8 GPUs.
  __shared__ int s_ctr[kMaxPeers][C_PER_CALL];
  if ((int)threadIdx.x < n_shards * C_PER_CALL) s_ctr[threadIdx.x / C_PER_CALL][threadIdx.x % C_PER_CALL] = shards[threadIdx.x / C_PER_CALL].ctr[threadIdx.x % C_PER_CALL];
  __syncthreads();
  {
    int* const now = M.hdr + kSkipSigBase + (gen & 1) * kSkipSigInts;
    const int* const before = M.hdr + kSkipSigBase + ((gen - 1) & 1) * kSkipSigInts;
    bool rebuild = !lazy || before[0] != n_shards;
    for (int r = 0; r < n_shards; ++r) rebuild = rebuild || before[1 + r] != s_ctr[r][C_DIRTY];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      now[0] = n_shards;
      for (int r = 0; r < n_shards; ++r) now[1 + r] = s_ctr[r][C_DIRTY];
      M.hdr[10 + (gen & 1)] = rebuild ? 1 : 0;
      M.hdr[8 + (gen & 1)] = M.hdr[8 + ((gen - 1) & 1)] + (rebuild ? 1 : 0);
      if (rebuild) M.hdr[12]++;
    }
    if (!rebuild) return;
  }
  // AABB of every shard's inserts (one shard = the engine itself; several = a volume sharded over GPUs)
  int x0 = 0x7FFFFFFF, y0 = 0x7FFFFFFF, z0 = 0x7FFFFFFF, x1 = -0x7FFFFFFF, y1 = -0x7FFFFFFF, z1 = -0x7FFFFFFF;
  for (int r = 0; r < n_shards; ++r) {
    const int* c = s_ctr[r];
    x0 = min(x0, c[C_MIN_X]); y0 = min(y0, c[C_MIN_Y]); z0 = min(z0, c[C_MIN_Z]);
    x1 = max(x1, c[C_MAX_X]); y1 = max(y1, c[C_MAX_Y]); z1 = max(z1, c[C_MAX_Z]);
  }
  int shift = 0, n = 0;
  long long nx = 0, ny = 0, nz = 0;
  if (x1 >= x0) {
    for (;; ++shift) {
      nx = ((x1 - x0) >> shift) + 1; ny = ((y1 - y0) >> shift) + 1; nz = ((z1 - z0) >> shift) + 1;
      if (nx * ny * nz <= (long long)kSkipMaxCells) break;
    }
    n = (int)(nx * ny * nz);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int* h = M.hdr;
    h[13] = h[7];  // cells the previous rebuild dirtied in the plane this one restores (the layout may have shrunk: larger shift)
    h[0] = n ? x0 : 0; h[1] = n ? y0 : 0; h[2] = n ? z0 : 0; h[3] = (int)nx; h[4] = (int)ny; h[5] = (int)nz; h[6] = shift; h[7] = n;
  }
  if (n == 0) return;
  unsigned char* const plane = planes_of(M, gen).mark;  // all cells at the cap (see above)
  // one index space over the directories of all shards, so that a thread reads about ONE entry (a remote load for a
  // foreign shard) instead of one per shard, one after the other
  int first[kMaxPeers + 1];
  first[0] = 0;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) first[r + 1] = first[r] + (r < n_shards ? s_ctr[r][C_HIGH_WATER] : 0);
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < first[kMaxPeers]; g += gridDim.x * blockDim.x) {
    int r = 0;
#pragma unroll
    for (int q = 1; q < kMaxPeers; ++q) r += g >= first[q] ? 1 : 0;
    const int i = g - first[r];
    const u64 k = shards[r].block_key[i];
    if (k == kEmpty) continue;
    int bx, by, bz; unpack_key(k, bx, by, bz);
    const int cx = (bx - x0) >> shift, cy = (by - y0) >> shift, cz = (bz - z0) >> shift;
    const size_t cell = ((size_t)cz * ny + cy) * nx + cx;
    plane[cell] = 0;
    // occupied cell: >= 0.  One cell = one block (shift 0): owner shard << kIndexShardShift | pool index (the directory
    // position IS the pool index), so the ray caster needs no table probe; coarser cells: 0, the caster probes the table
    M.cells[cell] = shift == 0 ? ((r << kIndexShardShift) | i) : 0;
  }
}

// one separable pass of the capped Chebyshev distance transform along AXIS:
//   out(c) = min_k max(in(c +- k e_axis), k),  k < cap
// All 2 * (cap - 1) neighbour reads of a cell are independent (no early exit), so they are in flight together.
// Plane roles: pass 0 reads the marked plane and writes the other, pass 1 goes back, pass 2 reads the marked plane again
// and writes the merged grid the ray caster reads with ONE load per sample -- empty cells get -distance, occupied cells
// keep the entry the mark kernel wrote -- and puts the other plane (idle by now) back to the cap for the next rebuild.
template <int AXIS>
__global__ void __launch_bounds__(256) skip_pass_kernel(SkipMap M, int gen) {
  if (!M.hdr[10 + (gen & 1)]) return;
  const Planes P = planes_of(M, gen);
  const unsigned char* __restrict__ in = AXIS == 1 ? P.other : P.mark;
  unsigned char* __restrict__ out = AXIS == 1 ? P.mark : P.other;
  const int nx = M.hdr[3], ny = M.hdr[4], nz = M.hdr[5], n = M.hdr[7];
  const int len = AXIS == 0 ? nx : AXIS == 1 ? ny : nz;
  const int stride = AXIS == 0 ? 1 : AXIS == 1 ? nx : nx * ny;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = AXIS == 0 ? i % nx : AXIS == 1 ? (i / nx) % ny : i / (nx * ny);
    int best = in[i];
    if (best > 1) {
#pragma unroll
      for (int k = 1; k < kSkipCap; ++k) {
        const int lo = c - k >= 0 ? (int)in[i - k * stride] : kSkipCap;
        const int hi = c + k < len ? (int)in[i + k * stride] : kSkipCap;
        best = min(best, max(min(lo, hi), k));
      }
    }
    if (AXIS == 2) { if (best > 0) M.cells[i] = -best; out[i] = (unsigned char)kSkipCap; }
    else out[i] = (unsigned char)best;
  }
  if (AXIS == 2)  // the plane being restored was the previous rebuild's marked plane: dirty up to THAT rebuild's cell count
    for (int i = n + blockIdx.x * blockDim.x + threadIdx.x; i < M.hdr[13]; i += gridDim.x * blockDim.x) out[i] = (unsigned char)kSkipCap;
}

void launch_build_skip_map(const PeerView* shards, int n_shards, const SkipMap& M, int gen, bool lazy, int num_sms, cudaStream_t st) {
  skip_mark_kernel<<<num_sms * 2, 256, 0, st>>>(shards, n_shards, M, gen, lazy ? 1 : 0);
  skip_pass_kernel<0><<<num_sms * 8, 256, 0, st>>>(M, gen);
  skip_pass_kernel<1><<<num_sms * 8, 256, 0, st>>>(M, gen);
  skip_pass_kernel<2><<<num_sms * 8, 256, 0, st>>>(M, gen);
}

// ------------------------------------------------------------------------------------------
// ray march
// ------------------------------------------------------------------------------------------
struct Grid { int ox, oy, oz, nx, ny, nz, shift; const int* cells; };
struct BlockCache { int bx, by, bz; int entry; };  // bx = INT_MIN: nothing cached; entry < 0: absent

// Where the voxels of a block live.  A block is named by its ENTRY: the pool index (local volume), or owner shard
// << kIndexShardShift | pool index (sharded volume) -- exactly what the map's occupied cells hold.
//   tsdf_at(entry, voxel)   a value of the block's TSDF plane, all the march and the gradient need
//   block(entry)  the block itself, for the colour / logit of the one hit voxel
// Local: this engine's table and pool.  Shared: the pool of the shard that owns the block, reached through peer-mapped
// pointers (NVLink loads inside the march) -- or, when the shards keep TSDF mirrors, a LOCAL copy of every shard's
// TSDF planes that the owners' integrate kernels keep up to date with posted NVLink stores (tsdf_device.cuh).
template <bool SHARED> struct Volume;
template <> struct Volume<false> {
  DeviceState S;
  __device__ __forceinline__ const unsigned char* block(int entry) const { return S.voxels + (size_t)entry * kBlockBytes; }
  __device__ __forceinline__ float tsdf_at(int entry, int voxel) const { return __ldg(reinterpret_cast<const float*>(block(entry)) + voxel); }
  __device__ __forceinline__ int find_entry(int bx, int by, int bz) const { return table_find(S, pack_key(bx, by, bz)); }
};
template <> struct Volume<true> {
  const PeerView* shards; int n_shards, shard_shift;
  const unsigned char* pool[kMaxPeers];    // every shard's voxel pool (peer-mapped): kernel parameters, not a load per sample
  const float* mirror; int mirror_stride;  // local TSDF mirror of all shards ([shard][pool index][512]) or null
  // Pulled TSDF cache (tsdf_shared_cache_attach): like the mirror a local [shard][pool index][512] array, but filled on
  // demand -- before the march, pull_select / pull_copy_kernel fetch the TSDF planes of the foreign blocks that can meet
  // this launch's rays (bulk NVLink reads, all in flight together) and stamp them with the current content epoch.  A
  // sample of a foreign block whose stamp is not current falls back to the owner's memory, so the result never depends
  // on how good the frustum test was.  Own blocks are read from the own pool.
  const float* cache; const int* cache_stamp; int cache_epoch, self;
  // Fused exchange of the results: when n_out > 0 every ray's pixel is stored into the image buffers of ALL ranks
  // (peer-mapped pointers, posted stores over NVLink issued as the rays finish) instead of into one local image that an
  // all-gather would have to distribute afterwards.
  int n_out; uchar4* out_rgba[kMaxPeers]; uchar4* out_normal[kMaxPeers]; float* out_depth[kMaxPeers];
  // which 8-row tiles this launch renders: tile_first, tile_first + tile_stride, ...  (stride 1 = a contiguous band;
  // stride = number of ranks interleaves the tiles of a view over the ranks, so that all of them finish together)
  int tile_stride;
  __device__ __forceinline__ const unsigned char* block(int entry) const {  // owner shard in the top bits: no table probe over NVLink
    return pool[entry >> kIndexShardShift] + (size_t)(entry & ((1 << kIndexShardShift) - 1)) * kBlockBytes;
  }
  __device__ __forceinline__ float tsdf_at(int entry, int voxel) const {
    const int shard = entry >> kIndexShardShift, idx = entry & ((1 << kIndexShardShift) - 1);
    if (mirror) return __ldg(mirror + ((size_t)shard * mirror_stride + idx) * kBlockVolume + voxel);
    const float* home = reinterpret_cast<const float*>(pool[shard] + (size_t)idx * kBlockBytes) + voxel;
    if (cache && shard != self) {
      const size_t slot = (size_t)shard * mirror_stride + idx;
      const int stamp = __ldg(cache_stamp + slot);         // both loads are local and in flight together
      const float v = __ldg(cache + slot * kBlockVolume + voxel);
      if (stamp == cache_epoch) return v;
    }
    return __ldg(home);
  }
  __device__ __forceinline__ int find_entry(int bx, int by, int bz) const {
    const u64 key = pack_key(bx, by, bz);
    const int r = (int)owner_of(key, n_shards, shard_shift);
    const int idx = table_find_in(shards[r].table, shards[r].table_mask, key);
    return idx < 0 ? -1 : ((r << kIndexShardShift) | idx);
  }
};
__device__ __forceinline__ const uint32_t* base_rgbw(const unsigned char* b) { return reinterpret_cast<const uint32_t*>(b + kPlaneBytes); }
__device__ __forceinline__ const float* base_logit(const unsigned char* b) { return reinterpret_cast<const float*>(b + 2 * kPlaneBytes); }

// What the map says about block (bx, by, bz):
//   >= 0       its cell holds an active block (the value is the dense index entry when one cell is one block)
//   -d         the nearest cell holding a block is at least d cells away (Chebyshev); every cell closer is empty
//   kEscaped   the ray can never meet a block again
// `leaving`: bit a = the ray does not increase along axis a, bit 3 + a = it does not decrease.  A sample outside the
// AABB on a side the ray is moving away from (or along) can never be followed by a sample inside it -- the accumulated
// position is monotonic per axis and no block exists outside the AABB -- so the ray is a miss.
constexpr int kEscaped = (int)0x80000000;
// DENSE: one cell = one block (shift 0), the usual case; the march kernel branches once, uniformly, into the
// instantiation that knows it (no shifts, no table probes in its code)
template <bool DENSE>
__device__ __forceinline__ int cell_value(const Grid& G, int bx, int by, int bz, unsigned leaving = 0u) {
  const int sh = DENSE ? 0 : G.shift;
  const int cx = (bx - G.ox) >> sh, cy = (by - G.oy) >> sh, cz = (bz - G.oz) >> sh;
  if ((unsigned)cx < (unsigned)G.nx && (unsigned)cy < (unsigned)G.ny && (unsigned)cz < (unsigned)G.nz) {
    return __ldg(G.cells + (unsigned)((cz * G.ny + cy) * G.nx + cx));  // at most 2^22 cells
  }
  const unsigned below = (cx < 0 ? 1u : 0u) | (cy < 0 ? 2u : 0u) | (cz < 0 ? 4u : 0u);
  const unsigned above = (cx >= G.nx ? 8u : 0u) | (cy >= G.ny ? 16u : 0u) | (cz >= G.nz ? 32u : 0u);
  if ((below | above) & leaving) return kEscaped;
  // outside the AABB: the gap to the box (in cells) is a lower bound of the distance
  const int gx = cx < 0 ? -cx : (cx >= G.nx ? cx - G.nx + 1 : 0);
  const int gy = cy < 0 ? -cy : (cy >= G.ny ? cy - G.ny + 1 : 0);
  const int gz = cz < 0 ? -cz : (cz >= G.nz ? cz - G.nz + 1 : 0);
  return -max(gx, max(gy, gz));
}

// nearest voxel of a grid position: roundf + the reference's saturating float -> short cast.  CLAMP = false is chosen on
// the host when no sample of the view can leave (-32000, 32000) voxels (camera position + max_depth), so the
// saturation can never act and its two integer min/max per coordinate are dropped.
template <bool CLAMP>
__device__ __forceinline__ int nearest_voxel(float f) { return CLAMP ? round_to_voxel(f) : __float2int_rz(roundf(f)); }

// entry of block (bx, by, bz) or -1: straight from the map cell when one cell is one block (the usual case), through the
// owner's hash table otherwise
template <class V>
__device__ __forceinline__ int block_entry(const V& vol, const Grid& G, int bx, int by, int bz) {
  if (G.shift == 0) { const int g = cell_value<true>(G, bx, by, bz); return g >= 0 ? g : -1; }
  return cell_value<false>(G, bx, by, bz) >= 0 ? vol.find_entry(bx, by, bz) : -1;
}

template <class V>
__device__ __forceinline__ void cache_lookup(const V& vol, const Grid& G, BlockCache& c, int px, int py, int pz) {
  const int bx = px >> 3, by = py >> 3, bz = pz >> 3;
  if ((bx ^ c.bx) | (by ^ c.by) | (bz ^ c.bz)) {
    c.bx = bx; c.by = by; c.bz = bz;
    c.entry = block_entry(vol, G, bx, by, bz);
  }
}
// Retrieve<VoxelTSDF>: absent -> VoxelTSDF() == +1 (voxel_types.cu:8)
template <class V>
__device__ __forceinline__ float fetch_tsdf(const V& vol, const Grid& G, BlockCache& c, int px, int py, int pz) {
  cache_lookup(vol, G, c, px, py, pz);
  if (c.entry < 0) return 1.f;
  return vol.tsdf_at(c.entry, voxel_index(px, py, pz));
}
template <bool CLAMP, class V>
__device__ __forceinline__ float fetch_tsdf_f(const V& vol, const Grid& G, BlockCache& c, float3 p) {
  return fetch_tsdf(vol, G, c, nearest_voxel<CLAMP>(p.x), nearest_voxel<CLAMP>(p.y), nearest_voxel<CLAMP>(p.z));
}

// One march sample: TSDF at the voxel nearest to p and `skip` = how many of the FOLLOWING samples are
// guaranteed to land in unallocated space.  With d = distance of p's cell and cs voxels per cell,
// every voxel within Chebyshev radius (d-1)*cs of p is unallocated; m further steps move the rounded
// voxel by at most m*smax + 1 (+1 slack for float accumulation), so m = floor(((d-1)*cs - 2) / smax).
// No per-lane block cache here: with steps of 3 voxels and blocks of 8 some lane of a warp changes block in 93 % of
// the rounds (ncu), so the warp took the look-up path anyway -- and a cache updated on divergent paths cost more
// register shuffling than the load it saved.  The map cell is loaded every sample (L1-resident, 4 bytes).
template <bool CLAMP, bool DENSE, class V>
__device__ __forceinline__ float march_sample(const V& vol, const Grid& G, float3 p, float inv_smax, unsigned leaving, int& skip) {
  const int px = nearest_voxel<CLAMP>(p.x), py = nearest_voxel<CLAMP>(p.y), pz = nearest_voxel<CLAMP>(p.z);
  const int bx = px >> 3, by = py >> 3, bz = pz >> 3;
  const int g = cell_value<DENSE>(G, bx, by, bz, leaving);
  skip = 0;
  float t = 1.f;
  if (g >= 0) {
    const int entry = DENSE ? g : vol.find_entry(bx, by, bz);
    if (DENSE || entry >= 0) t = vol.tsdf_at(entry, voxel_index(px, py, pz));
  } else if (g == kEscaped) {
    skip = kEscaped;  // the ray has left the volume for good
  } else if (g <= -2) {
    skip = __float2int_rd((float)(((-g - 1) << (DENSE ? 3 : 3 + G.shift)) - 2) * inv_smax);
  }
  return t;
}

// The march proper: from sample 0 to the first hit (returns true, position = the hit sample) or to the end of the ray.
// Sample i sits at the position accumulated by i additions.  One round = one sample that is really looked up, followed
// by the advance over it and over the `skip` samples after it that provably read +1 (they can neither start nor
// complete a hit, voxel_tsdf.cu:260, and leave tsdf_prev at +1, which it already is: a sample with skip > 0 is itself
// unallocated).  tsdf_prev starts negative so that sample 0 only initialises it, like the reference's pre-loop Retrieve.
template <bool CLAMP, bool DENSE, class V>
__device__ __forceinline__ bool march(const V& vol, const Grid& G, f32x2& pxy, float& pz, f32x2 sxy, float sz, float inv_smax, unsigned leaving,
                                      int max_step) {
  float tsdf_prev = -1.f;
  int i = 0;
  for (;;) {
    int skip;
    const float tsdf_curr = march_sample<CLAMP, DENSE>(vol, G, f3(lo2(pxy), hi2(pxy), pz), inv_smax, leaving, skip);
    // ray hit front surface (voxel_tsdf.cu:260)
    if (tsdf_prev > 0 && tsdf_curr <= 0 && tsdf_prev - tsdf_curr <= 1.5f) return true;
    tsdf_prev = tsdf_curr;
    if (skip == kEscaped) return false;         // a miss, whose position nobody reads
    const int k = min(skip, max_step - 1 - i);  // samples beyond max_step - 1 do not exist
    pxy = add2(pxy, sxy); pz += sz;             // the step over the sample just taken
    if (k > 0) {                                // ... and over the skipped ones: exactly the additions the reference performs
      if (k & 1) { pxy = add2(pxy, sxy); pz += sz; }
      if (k & 2) { pxy = add2(pxy, sxy); pz += sz; pxy = add2(pxy, sxy); pz += sz; }
      if (k & 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { pxy = add2(pxy, sxy); pz += sz; }
      }
      for (int j = k >> 3; j > 0; --j) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { pxy = add2(pxy, sxy); pz += sz; }
      }
    }
    i += k + 1;
    if (i >= max_step) return false;
  }
}

// ------------------------------------------------------------------------------------------
// pulled TSDF cache of a sharded volume (see Volume<true>)
// ------------------------------------------------------------------------------------------
// The rays of a launch lie in the pyramid { x/z in [xlo, xhi], y/z in [ylo, yhi], 0 <= z <= zfar } of the camera frame.
// A block whose (padded) box has all eight corners beyond one of the six planes cannot contain a sample of these rays;
// every other foreign block that is not in the cache at the current content epoch is listed.  The padding covers the
// nearest-voxel rounding, the +-1 voxel gradient samples and the float32 accumulation of the positions; a block the test
// wrongly drops is read from its owner sample by sample (tsdf_at), so the image cannot change (`pad` voxels; 3 covers all
// of that with room to spare, a negative value makes the test drop blocks and is what the tests use to walk the fallback).
struct PullView { float xlo, xhi, ylo, yhi, zfar; };
__global__ void __launch_bounds__(256) pull_select_kernel(const PeerView* __restrict__ shards, int n_shards, int self, FrameParams P, PullView F,
                                                          const int* __restrict__ stamp, int stride, int epoch, int pad, int* __restrict__ list,
                                                          int* __restrict__ count, int serial) {
  const unsigned lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) count[(serial + 1) & 3] = 0;  // the next launch's counter (nobody touches it now)
  int* const n_listed = count + (serial & 3);
  __shared__ int s_hw[kMaxPeers];  // every shard's high-water mark in one round trip (see skip_mark_kernel)
  if ((int)threadIdx.x < n_shards) s_hw[threadIdx.x] = shards[threadIdx.x].ctr[C_HIGH_WATER];
  __syncthreads();
  int first[kMaxPeers + 1];  // one index space over the foreign directories: about one (remote) entry per thread
  first[0] = 0;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) first[r + 1] = first[r] + (r < n_shards && r != self ? min(s_hw[r], stride) : 0);
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < first[kMaxPeers]; base += gridDim.x * blockDim.x) {
    const int g = base + (int)lane;
    bool want = false;
    int r = 0, i = 0;
    if (g < first[kMaxPeers]) {
#pragma unroll
      for (int q = 1; q < kMaxPeers; ++q) r += g >= first[q] ? 1 : 0;
      i = g - first[r];
      const u64 k = shards[r].block_key[i];
      if (k != kEmpty && stamp[(size_t)r * stride + i] != epoch) {
        int bx, by, bz; unpack_key(k, bx, by, bz);
        unsigned beyond = 0x3Fu;  // bit p: every corner so far lies beyond plane p
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float wx = (float)((bx << 3) + ((c & 1) ? kBlockLen - 1 + pad : -pad)) * P.voxel_size;
          const float wy = (float)((by << 3) + ((c & 2) ? kBlockLen - 1 + pad : -pad)) * P.voxel_size;
          const float wz = (float)((bz << 3) + ((c & 4) ? kBlockLen - 1 + pad : -pad)) * P.voxel_size;
          const float3 q = apply(P.cam_T_world, f3(wx, wy, wz));
          unsigned in = 0u;
          if (!(q.x - F.xlo * q.z < 0.f)) in |= 1u;
          if (!(q.x - F.xhi * q.z > 0.f)) in |= 2u;
          if (!(q.y - F.ylo * q.z < 0.f)) in |= 4u;
          if (!(q.y - F.yhi * q.z > 0.f)) in |= 8u;
          if (!(q.z < 0.f)) in |= 16u;
          if (!(q.z > F.zfar)) in |= 32u;
          beyond &= ~in;
        }
        want = beyond == 0u;
      }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
    if (m) {
      int off = 0;
      if (lane == (unsigned)(__ffs(m) - 1)) off = atomicAdd(n_listed, __popc(m));
      off = __shfl_sync(0xFFFFFFFFu, off, __ffs(m) - 1);
      if (want) list[off + __popc(m & ((1u << lane) - 1u))] = (r << kIndexShardShift) | i;
    }
  }
}
// one warp per listed block: its 2 KB TSDF plane in four 16-byte loads per lane, all issued before the first store
struct PullPools { const unsigned char* pool[kMaxPeers]; };
__global__ void __launch_bounds__(256) pull_copy_kernel(PullPools pools, const int* __restrict__ list, const int* __restrict__ count, int serial,
                                                        float* __restrict__ cache, int* __restrict__ stamp, int stride, int epoch) {
  const int n = count[serial & 3];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n; b += warps) {
    const int entry = list[b];
    const int shard = entry >> kIndexShardShift, idx = entry & ((1 << kIndexShardShift) - 1);
    const float4* src = reinterpret_cast<const float4*>(pools.pool[shard] + (size_t)idx * kBlockBytes);
    const size_t slot = (size_t)shard * stride + idx;
    float4* dst = reinterpret_cast<float4*>(cache + slot * kBlockVolume);
    const float4 v0 = __ldcg(src + lane), v1 = __ldcg(src + lane + 32), v2 = __ldcg(src + lane + 64), v3 = __ldcg(src + lane + 96);
    dst[lane] = v0; dst[lane + 32] = v1; dst[lane + 64] = v2; dst[lane + 96] = v3;
    if (lane == 0) stamp[slot] = epoch;
  }
}

__device__ __forceinline__ unsigned char f2u8(float f) { return (unsigned char)min(255, max(0, __float2int_rz(f))); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }

__device__ __forceinline__ int tile_stride_of(const Volume<false>&) { return 1; }
__device__ __forceinline__ int tile_stride_of(const Volume<true>& v) { return v.tile_stride; }

template <bool SHARED, bool CLAMP>
__global__ void __launch_bounds__(256) raycast_kernel(Volume<SHARED> vol, FrameParams P, float step_size, SkipMap M, int row0,
                                                      int rows, uchar4* __restrict__ img_rgba, uchar4* __restrict__ img_normal,
                                                      float* __restrict__ img_depth, u64* __restrict__ packed) {
  // CTA = 32 x 8 pixels, warp = 8 x 4 pixels: neighbouring rays walk the same cells and blocks; the launch covers
  // image rows [row0, row0 + rows)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
  const int y = row0 + (SHARED ? (int)blockIdx.y * tile_stride_of(vol) : (int)blockIdx.y) * 8 + (warp >> 2) * 4 + (lane >> 3);
  const bool in_image = x < P.w && y < P.h && y < row0 + rows;
  if (!SHARED && !in_image) return;  // (the shared variant keeps every thread: its CTAs store their tile together)
  const int idx = y * P.w + x;
  uint32_t out_rgba = 0u, out_normal = 0u;  // r | g << 8 | b << 16 | a << 24; a miss is (0, 0, 0, 0) like the reference
  float out_depth = CUDART_INF_F;
  if (in_image) {
  Grid G;
  G.ox = M.hdr[0]; G.oy = M.hdr[1]; G.oz = M.hdr[2]; G.nx = M.hdr[3]; G.ny = M.hdr[4]; G.nz = M.hdr[5]; G.shift = M.hdr[6];
  G.cells = M.cells;

  // voxel_tsdf.cu:243-250
  const float3 pos_cam = kmul(P.Kinv, f3((float)x, (float)y, 1.f));
  const float sq = sqnorm3(pos_cam);
  float3 ray_dir_cam = pos_cam;
  if (sq > 0.f) {  // normalized(): v / sqrt(squaredNorm)
    const float n = sqrtf(sq);
    if (div_safe(n)) { const float r = rcp_refined(n); ray_dir_cam = f3(div_by(pos_cam.x, n, r), div_by(pos_cam.y, n, r), div_by(pos_cam.z, n, r)); }
    else ray_dir_cam = f3(pos_cam.x / n, pos_cam.y / n, pos_cam.z / n);
  }
  const float3 ray_dir_world = qrot(P.world_T_cam, ray_dir_cam);
  float3 ray_step_grid, pos_grid;
  if (div_safe(P.voxel_size)) {  // the six divisions by voxel_size share one reciprocal (div_by: bit-identical to `/`)
    const float r = rcp_refined(P.voxel_size);
    ray_step_grid = f3(div_by(ray_dir_world.x * step_size, P.voxel_size, r), div_by(ray_dir_world.y * step_size, P.voxel_size, r),
                       div_by(ray_dir_world.z * step_size, P.voxel_size, r));
    pos_grid = f3(div_by(P.world_T_cam.tx, P.voxel_size, r), div_by(P.world_T_cam.ty, P.voxel_size, r), div_by(P.world_T_cam.tz, P.voxel_size, r));
  } else {
    ray_step_grid = f3(ray_dir_world.x * step_size / P.voxel_size, ray_dir_world.y * step_size / P.voxel_size,
                       ray_dir_world.z * step_size / P.voxel_size);
    pos_grid = f3(P.world_T_cam.tx / P.voxel_size, P.world_T_cam.ty / P.voxel_size, P.world_T_cam.tz / P.voxel_size);
  }
  const int max_step = __float2int_rz(ceilf(P.max_depth / step_size));
  // conservative 1 / (largest per-step voxel displacement); only used to size skips
  const float smax = fmaxf(fmaxf(fabsf(ray_step_grid.x), fabsf(ray_step_grid.y)), fabsf(ray_step_grid.z));
  const float inv_smax = 1.f / (smax * 1.001f + 1e-6f);

  BlockCache cache; cache.bx = cache.by = cache.bz = (int)0x80000000; cache.entry = -1;
  // sides of the block AABB this ray can only move away from (see cell_distance)
  const unsigned leaving = (ray_step_grid.x <= 0.f ? 1u : 0u) | (ray_step_grid.y <= 0.f ? 2u : 0u) | (ray_step_grid.z <= 0.f ? 4u : 0u) |
                           (ray_step_grid.x >= 0.f ? 8u : 0u) | (ray_step_grid.y >= 0.f ? 16u : 0u) | (ray_step_grid.z >= 0.f ? 32u : 0u);
  // x and y of the position live in one register pair for the whole march and advance in one FADD2 per step
  // (bit-identical to two scalar adds); reading a half of the pair is free
  f32x2 pxy = pack2(pos_grid.x, pos_grid.y);
  const f32x2 sxy = pack2(ray_step_grid.x, ray_step_grid.y);
  float pz = pos_grid.z;
  const float sz = ray_step_grid.z;

  const bool hit = G.shift == 0 ? march<CLAMP, true>(vol, G, pxy, pz, sxy, sz, inv_smax, leaving, max_step)
                                : march<CLAMP, false>(vol, G, pxy, pz, sxy, sz, inv_smax, leaving, max_step);
  pos_grid = f3(lo2(pxy), hi2(pxy), pz);

  // The refinement runs after the march loop so that the lanes of a warp execute it together (once per
  // warp) instead of once per distinct hit iteration.
  if (hit) {
    float3 pos1 = f3(pos_grid.x - ray_step_grid.x, pos_grid.y - ray_step_grid.y, pos_grid.z - ray_step_grid.z);
    float3 pos2 = pos_grid;
    float3 mid = f3((pos1.x + pos2.x) / 2.f, (pos1.y + pos2.y) / 2.f, (pos1.z + pos2.z) / 2.f);
    // binary search refinement; `> .1` compares in double in the reference, i.e. >= 0.1f for floats
    for (;;) {
      const float3 dd = f3(pos1.x - pos2.x, pos1.y - pos2.y, pos1.z - pos2.z);
      if (!(dot3(dd, dd) >= 0.1f)) break;
      const float tm = fetch_tsdf_f<CLAMP>(vol, G, cache, mid);
      if (tm < 0) pos2 = mid; else pos1 = mid;
      mid = f3((pos1.x + pos2.x) / 2.f, (pos1.y + pos2.y) / 2.f, (pos1.z + pos2.z) / 2.f);
    }
    const int fx = nearest_voxel<CLAMP>(mid.x), fy = nearest_voxel<CLAMP>(mid.y), fz = nearest_voxel<CLAMP>(mid.z);
    cache_lookup(vol, G, cache, fx, fy, fz);
    const int centry = cache.entry;
    // Block of each of the 6 neighbours first (only a neighbour across a block face needs a table lookup,
    // at most one per axis), then all 8 voxel loads of the hit -- colour, logit and the 6 gradient samples -- are
    // independent and in flight together (voxel_tsdf.cu:277-291; short arithmetic wraps like the reference).
    const int nx[6] = {(short)(fx + 1), (short)(fx - 1), fx, fx, fx, fx};
    const int ny[6] = {fy, fy, (short)(fy + 1), (short)(fy - 1), fy, fy};
    const int nz[6] = {fz, fz, fz, fz, (short)(fz + 1), (short)(fz - 1)};
    int nentry[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) {
      nentry[n] = centry;
      if (((nx[n] ^ fx) | (ny[n] ^ fy) | (nz[n] ^ fz)) >> 3) {  // different block coordinate
        const int bx = nx[n] >> 3, by = ny[n] >> 3, bz = nz[n] >> 3;
        nentry[n] = block_entry(vol, G, bx, by, bz);
      }
    }
    uint32_t rgbw = 0u;  // VoxelRGBW() / VoxelSEGM() defaults for an absent voxel (voxel_types.cu:3,11)
    float lgt = 0.f;
    if (centry >= 0) {
      const int k = voxel_index(fx, fy, fz);
      const unsigned char* const cbase = vol.block(centry);
      rgbw = __ldg(base_rgbw(cbase) + k);
      lgt = __ldg(base_logit(cbase) + k);
    }
    float gv[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) gv[n] = nentry[n] >= 0 ? vol.tsdf_at(nentry[n], voxel_index(nx[n], ny[n], nz[n])) : 1.f;
    const float prob = centry >= 0 ? logit_to_prob(lgt) : 0.f;
    const float gxp = gv[0], gxn = gv[1], gyp = gv[2], gyn = gv[3], gzp = gv[4], gzn = gv[5];
    const float3 nrm = f3(gxp - gxn, gyp - gyn, gzp - gzn);
    const float3 neg_dir = f3(-ray_dir_world.x, -ray_dir_world.y, -ray_dir_world.z);
    const float diffusivity = fmaxf(dot3(nrm, neg_dir) / sqrtf(sqnorm3(nrm)), 0);
    const float alpha = fmaxf(prob - 0.5f, 0) * 2.f;  // == fmaxf(p - .5, 0) / .5 exactly
    const float r = (float)(rgbw & 0xFF), g = (float)((rgbw >> 8) & 0xFF), b = (float)((rgbw >> 16) & 0xFF);
    out_rgba = (uint32_t)f2u8(alpha * 255 + (1 - alpha) * r) | ((uint32_t)f2u8((1 - alpha) * g) << 8) |
               ((uint32_t)f2u8((1 - alpha) * b) << 16) | 0xFF000000u;
    const uint32_t shade = f2u8((1 - alpha) * diffusivity * 255);
    out_normal = (uint32_t)f2u8(alpha * 255 + (1 - alpha) * diffusivity * 255) | (shade << 8) | (shade << 16) | 0xFF000000u;
    const float3 pc = apply(P.cam_T_world, f3(mid.x * P.voxel_size, mid.y * P.voxel_size, mid.z * P.voxel_size));
    out_depth = pc.z;
  }
  }  // in_image

  if constexpr (SHARED) {
    if (vol.n_out > 0) {
      // The CTA's 32 x 8 tile goes out together: staged in shared memory in image order, then each warp stores one row
      // of 32 pixels = 128 contiguous bytes per image and destination -- four times fewer, four times larger NVLink
      // packets than 8-pixel segments stored ray by ray.
      __shared__ uint32_t s_rgba[256], s_normal[256];
      __shared__ float s_depth[256];
      const int tx = (warp & 3) * 8 + (lane & 7), ty = (warp >> 2) * 4 + (lane >> 3);
      s_rgba[ty * 32 + tx] = out_rgba; s_normal[ty * 32 + tx] = out_normal; s_depth[ty * 32 + tx] = out_depth;
      __syncthreads();
      const int ox = blockIdx.x * 32 + lane, oy = y - ty + warp;  // warp w stores row w of the tile
      if (ox < P.w && oy < P.h && oy < row0 + rows) {
        const int o = oy * P.w + ox, t = warp * 32 + lane;
        for (int r = 0; r < vol.n_out; ++r) {
          if (vol.out_rgba[r]) reinterpret_cast<uint32_t*>(vol.out_rgba[r])[o] = s_rgba[t];
          if (vol.out_normal[r]) reinterpret_cast<uint32_t*>(vol.out_normal[r])[o] = s_normal[t];
          if (vol.out_depth[r]) vol.out_depth[r][o] = s_depth[t];
        }
      }
      return;
    }
    if (!in_image) return;
  }
  if (img_rgba) reinterpret_cast<uint32_t*>(img_rgba)[idx] = out_rgba;
  if (img_normal) reinterpret_cast<uint32_t*>(img_normal)[idx] = out_normal;
  if (img_depth) img_depth[idx] = out_depth;
  if (packed) {
    // positive floats order like unsigned ints; a miss (+inf) loses against every hit
    const u64 dbits = (u64)__float_as_uint(fmaxf(out_depth, 0.f)) << 32;
    packed[2 * idx + 0] = dbits | (u64)out_rgba;
    packed[2 * idx + 1] = dbits | (u64)out_normal;
  }
}

// true when some sample of this view could reach +-32000 voxels, i.e. when the reference's saturating short cast could act:
// |position| <= |camera centre| + (max_depth + one step) along the ray, in voxels, plus slack for the float accumulation
static bool view_needs_clamp(const FrameParams& P, float step_size) {
  const float t = fmaxf(fmaxf(fabsf(P.world_T_cam.tx), fabsf(P.world_T_cam.ty)), fabsf(P.world_T_cam.tz));
  const float reach = (t + (P.max_depth + 2 * step_size) * 1.01f) / P.voxel_size + 64.f;
  return !(reach < 32000.f);  // also true for NaN / inf
}

void launch_raycast(const DeviceState& S, const FrameParams& P, float step_size, const SkipMap& M, uchar4* rgba,
                    uchar4* normal, float* hit_depth, unsigned long long* packed_keys, cudaStream_t st) {
  dim3 grid((P.w + 31) / 32, (P.h + 7) / 8);
  const SkipMap& R = M;
  Volume<false> vol; vol.S = S;
  if (view_needs_clamp(P, step_size)) raycast_kernel<false, true><<<grid, 256, 0, st>>>(vol, P, step_size, R, 0, P.h, rgba, normal, hit_depth, packed_keys);
  else raycast_kernel<false, false><<<grid, 256, 0, st>>>(vol, P, step_size, R, 0, P.h, rgba, normal, hit_depth, packed_keys);
}

// Rows [row0, row0 + rows) of a view over a volume sharded across `n_shards` engines whose tables and pools are
// mapped in `shards` (device array): bit-identical to the single-volume render, voxels of foreign blocks are read
// from their owner over NVLink.
void launch_raycast_shared(const PeerView* shards, const PeerView* host_shards, int n_shards, int shard_shift, const FrameParams& P, float step_size,
                           const SkipMap& M, int row0, int rows, int tile_stride, const float* mirror, int mirror_stride, const SharedCache& C,
                           uchar4* rgba, uchar4* normal, float* hit_depth, int n_out, void* const* out_rgba, void* const* out_normal,
                           void* const* out_depth, int num_sms, cudaStream_t st) {
  if (rows <= 0) return;
  // tile_stride > 1: `rows` bounds the rows of the image this launch may touch ([row0, row0 + rows)), of which it
  // renders every tile_stride-th 8-row tile starting at row0
  const int n_tiles = ((rows + 7) / 8 + tile_stride - 1) / tile_stride;
  dim3 grid((P.w + 31) / 32, n_tiles);
  const SkipMap& R = M;
  Volume<true> vol; vol.shards = shards; vol.n_shards = n_shards; vol.shard_shift = shard_shift;
  vol.n_out = n_out;
  vol.tile_stride = tile_stride;
  vol.mirror = mirror; vol.mirror_stride = mirror_stride;
  vol.cache = nullptr; vol.cache_stamp = nullptr; vol.cache_epoch = 0; vol.self = C.self;
  for (int r = 0; r < kMaxPeers; ++r) vol.pool[r] = r < n_shards ? host_shards[r].voxels : nullptr;
  if (!mirror && C.cache && n_shards > 1) {
    // fetch what the rays of rows [row0, row0 + rows) can meet (one pixel of slack on every side of the band)
    const float xa = P.Kinv.fx * -1.f + P.Kinv.cx, xb = P.Kinv.fx * (float)P.w + P.Kinv.cx;
    const float ya = P.Kinv.fy * (float)(row0 - 1) + P.Kinv.cy, yb = P.Kinv.fy * (float)(row0 + rows) + P.Kinv.cy;
    PullView F;
    F.xlo = fminf(xa, xb); F.xhi = fmaxf(xa, xb); F.ylo = fminf(ya, yb); F.yhi = fmaxf(ya, yb);
    F.zfar = P.max_depth + 2.f * step_size;
    PullPools pools;
    for (int r = 0; r < kMaxPeers; ++r) pools.pool[r] = vol.pool[r];
    pull_select_kernel<<<num_sms * 2, 256, 0, st>>>(shards, n_shards, C.self, P, F, C.stamp, C.stride, C.epoch, C.pad, C.list, C.count, C.serial);
    pull_copy_kernel<<<num_sms * 8, 256, 0, st>>>(pools, C.list, C.count, C.serial, C.cache, C.stamp, C.stride, C.epoch);
    vol.cache = C.cache; vol.cache_stamp = C.stamp; vol.cache_epoch = C.epoch; vol.mirror_stride = C.stride;
  }
  for (int r = 0; r < kMaxPeers; ++r) {
    vol.out_rgba[r] = r < n_out && out_rgba ? (uchar4*)out_rgba[r] : nullptr;
    vol.out_normal[r] = r < n_out && out_normal ? (uchar4*)out_normal[r] : nullptr;
    vol.out_depth[r] = r < n_out && out_depth ? (float*)out_depth[r] : nullptr;
  }
  if (view_needs_clamp(P, step_size)) raycast_kernel<true, true><<<grid, 256, 0, st>>>(vol, P, step_size, R, row0, rows, rgba, normal, hit_depth, nullptr);
  else raycast_kernel<true, false><<<grid, 256, 0, st>>>(vol, P, step_size, R, row0, rows, rgba, normal, hit_depth, nullptr);
}

}  // namespace tsdf
