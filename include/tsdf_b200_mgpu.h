/*
 * tsdf_b200_mgpu.h -- C ABI of the multi-GPU data plane (libtsdf_b200_mgpu.so): ONE volume whose voxel blocks are
 * sharded over the GPUs of a node by block-coordinate ownership (tsdf_block_owner), driven from C/C++ with NCCL
 * linked directly -- no Python and no libtorch on the per-frame path.
 *
 * The reference is single-GPU (one TSDFGrid, two streams: utils/tsdf/voxel_tsdf.cuh:103-104), so nothing here
 * replaces a reference function; the entry points mirror the TSDFGrid members they distribute:
 *
 *   tsdf_mgpu_integrate          TSDFGrid::Integrate   utils/tsdf/voxel_tsdf.cu:347-375
 *   tsdf_mgpu_raycast            TSDFGrid::RayCast     utils/tsdf/voxel_tsdf.cu:490-506
 *   tsdf_mgpu_raycast_composite  (same, nearest-hit min-compositing as BASELINE.json describes it)
 *   tsdf_mgpu_gather             TSDFGrid::GatherValid / GatherVoxels   utils/tsdf/voxel_tsdf.cu:399-454
 *
 * One rank = one GPU = one handle.  Ranks may be processes (one per GPU: exchange the id of tsdf_mgpu_unique_id by
 * any out-of-band channel -- a file, MPI, torch.distributed) or host threads of one process (one thread per GPU).
 * Every rank calls the same sequence of tsdf_mgpu_* functions with the same camera arguments, like the ranks of any
 * NCCL program.  Per frame, on every rank, everything is enqueued on CUDA streams and nothing waits on the host:
 *
 *   integrate   the four planes (15 B/px) are broadcast from the root over NVLink (one grouped ncclBroadcast on a
 *               communication stream, three staging sets, so that the broadcast of frame k+1 overlaps the kernels of
 *               frame k); then every rank enumerates the frame but allocates and integrates only the blocks it owns,
 *               and its integrate kernel also stores every TSDF value it updates into the TSDF mirrors of all ranks
 *               (posted NVLink stores, tsdf_mirror_attach).
 *               TSDF_MGPU_ALLOC=exchange: every rank stages the whole frame but walks the pixel rays of every N-th
 *               32 x 8 tile only and mails the candidate block keys to their owners' inboxes; a peer barrier publishes
 *               the counts; the owners insert (tsdf_alloc_exchange_attach).
 *   raycast     peer barrier -- a 32-thread kernel exchanging flag words with release / acquire at system scope, no
 *               collective library on the engine stream -- (every shard's Integrate has finished); the skip map is built
 *               from every shard's pool directory; each rank marches 1/N of the view's 8-row tiles, dealt round-robin
 *               (TSDF_MGPU_TILES=band: one contiguous band), over the WHOLE volume: TSDF samples from the local mirror,
 *               the colour and probability of a hit voxel from its owner over NVLink -- bit-identical to a single-GPU
 *               render -- and stores every finished ray into the images of ALL ranks (posted NVLink stores); peer barrier.
 *               TSDF_MGPU_MIRROR=pull: no mirrors; before the march the TSDF planes of the foreign blocks the view can
 *               meet are fetched into a local cache with bulk NVLink reads (tsdf_shared_cache_attach); =0: every foreign
 *               sample is a load over NVLink, a rank's memory holds just its shard.
 *               TSDF_MGPU_EXCHANGE=nccl selects the conventional form for comparison: 4-byte ncclAllReduce, local
 *               image, grouped in-place ncclAllGather.
 *
 *   TSDF_MGPU_MODE=replicas: not sharded at all.  Every rank keeps the WHOLE volume (pool_blocks of cfg must hold it) and
 *               integrates every frame after the same broadcast; view k is rendered by rank k % world with the plain
 *               single-GPU kernels, straight into rank 0's image memory.  No barrier inside a tsdf_mgpu_run_sequence;
 *               a stand-alone tsdf_mgpu_raycast ends with one.  Images, gathers and counters are those of rank 0
 *               (every replica's are identical).  It multiplies view throughput for volumes that fit one GPU; it adds
 *               no capacity.  tsdf_mgpu_raycast_composite is not available in this mode.
 *
 * Status codes are those of tsdf_b200.h; tsdf_mgpu_last_error() describes the last failure of the calling thread.
 */
#ifndef TSDF_B200_MGPU_H_
#define TSDF_B200_MGPU_H_

#include "tsdf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tsdf_mgpu* tsdf_mgpu_handle;

#define TSDF_MGPU_ID_BYTES 128 /* sizeof(ncclUniqueId) */

const char* tsdf_mgpu_last_error(void);

/* ncclGetUniqueId: called by ONE rank, the 128 bytes are handed to every rank before tsdf_mgpu_create. */
int tsdf_mgpu_unique_id(void* id /* TSDF_MGPU_ID_BYTES */);

/* Collective over all ranks.  cfg describes THIS rank's engine (device, pool_blocks and table_slots of the shard,
 * max_image_pixels; bits 0..3 of flags = shard granularity, identical on all ranks); shard_rank / shard_count are
 * filled in from rank / world.  Creates the engine, two NCCL communicators (frame broadcasts; barrier + image
 * all-gather, so that the two can overlap on different streams), and maps every other shard's table and pool
 * (CUDA IPC across processes, peer access inside one process). */
int tsdf_mgpu_create(float voxel_size, float truncation, const tsdf_config* cfg, int rank, int world, const void* id,
                     tsdf_mgpu_handle* out);
int tsdf_mgpu_destroy(tsdf_mgpu_handle h);

/* The engine of this rank (shard-local queries: counters, tsdf_export_blocks, profiling ...). */
tsdf_handle tsdf_mgpu_engine(tsdf_mgpu_handle h);

/* TSDFGrid::Integrate over the sharded volume.  On `root` the four planes are read from host memory
 * (planes_on_device = 0; pinned memory recommended) or from device memory of the root's GPU (planes_on_device = 1);
 * on the other ranks the plane pointers are ignored (pass NULL).  Camera arguments must be identical on all ranks.
 * Returns after enqueueing; host buffers may be reused after tsdf_mgpu_synchronize or two further integrate calls. */
int tsdf_mgpu_integrate(tsdf_mgpu_handle h, int root, int planes_on_device, const void* rgb, const void* depth,
                        const void* ht, const void* lt, int width, int height, float max_depth, const float K[4],
                        const float q_xyzw[4], const float t_xyz[3]);

/* TSDFGrid::RayCast over the sharded volume, exact.  Same arguments on every rank.  The results stay on the device:
 * *d_rgba / *d_normal (uchar4 per pixel) and *d_hit_depth (float) receive the addresses of the assembled HxW images
 * on this rank's GPU, valid until the next raycast on this handle; any of the three may be NULL.  Asynchronous on
 * the engine stream (tsdf_stream(tsdf_mgpu_engine(h))). */
int tsdf_mgpu_raycast(tsdf_mgpu_handle h, float max_depth, int width, int height, const float K[4],
                      const float q_xyzw[4], const float t_xyz[3], const void** d_rgba, const void** d_normal,
                      const void** d_hit_depth);
/* Nearest-hit min-compositing (the variant BASELINE.json names): every rank marches all rays over its own shard and
 * one ncclAllReduce(min) over the packed keys float_bits(hit depth) << 32 | colour keeps the nearest hit.  Not exact
 * where a hit straddles two shards; *d_keys = 2 x uint64 per pixel (rgba key, normal key). */
int tsdf_mgpu_raycast_composite(tsdf_mgpu_handle h, float max_depth, int width, int height, const float K[4],
                                const float q_xyzw[4], const float t_xyz[3], const void** d_keys);
/* Copies the images of the last tsdf_mgpu_raycast to host memory (any pointer may be NULL) and waits for them. */
int tsdf_mgpu_fetch_images(tsdf_mgpu_handle h, uint8_t* rgba, uint8_t* normal, float* hit_depth);

/* TSDFGrid::GatherValid (bbox == NULL) / GatherVoxels over all shards: every rank selects and emits its own blocks,
 * the records travel device-to-device to `root` (sizes first, then one grouped ncclSend / ncclRecv) and are copied to
 * out_xyzt there.  *n_voxels = total over all shards (on every rank); out_xyzt / cap_voxels are used on root only;
 * out_xyzt == NULL only counts.  Synchronous. */
int tsdf_mgpu_gather(tsdf_mgpu_handle h, int root, const float* bbox /* 6 floats or NULL */, float* out_xyzt,
                     int64_t cap_voxels, int64_t* n_voxels);

/* Sums over all shards (collective, synchronous): counters of the last frame / totals since tsdf_set_profiling
 * on the engines / VoxelHashTable::NumActiveBlock of the whole volume. */
int tsdf_mgpu_counters(tsdf_mgpu_handle h, tsdf_counters* last_frame_sum, tsdf_counters* totals_sum, int64_t* n_active_blocks);

/* A whole stream of frames in one call -- the per-frame loop in C++, nothing but this library and NCCL between the
 * frames.  For i in [first, first + count): frame = frames[i % n_frames]; tsdf_mgpu_integrate(frame), then
 * raycast_mode 1: tsdf_mgpu_raycast from the frame's camera, 2: tsdf_mgpu_raycast_composite, 0: no view.
 * Every rank passes the same cameras; plane pointers are read on `root` only.  Returns after enqueueing (the host only
 * ever waits for the frame two steps back, which bounds the pipeline depth).  With TSDF_MGPU_ALLOC=exchange the barrier
 * that ends a view is supplied, inside a sequence, by the next frame's candidate-exchange barrier (nothing before it
 * touches the volume or the images): a frame with a view costs two barriers, not three. */
typedef struct tsdf_mgpu_frame {
  const void *rgb, *depth, *ht, *lt; /* root only; host or device memory according to planes_on_device */
  float q_xyzw[4];
  float t_xyz[3];
  float reserved;
} tsdf_mgpu_frame;
int tsdf_mgpu_run_sequence(tsdf_mgpu_handle h, int root, int planes_on_device, const tsdf_mgpu_frame* frames, int n_frames,
                           int first, int count, int width, int height, float max_depth, const float K[4], int raycast_mode);

/* Waits for everything this rank has enqueued (engine and communication streams). */
int tsdf_mgpu_synchronize(tsdf_mgpu_handle h);

/* Collective timing.  While enabled every NCCL call is bracketed by CUDA events on its stream; the getter (which
 * synchronises) returns device milliseconds and call counts summed since enabling:
 *   [0] frame broadcast  [1] barrier before the march (peer barrier kernel / 4-byte all-reduce)  [2] barrier after the
 *   march / image all-gather  [3] composite all-reduce  [4] shared-volume raycast kernels (skip map over all shards,
 *   TSDF fetch, march + scatter)  [5] gather send / recv  [6] candidate-exchange barrier of a frame */
int tsdf_mgpu_set_profiling(tsdf_mgpu_handle h, int enabled);
int tsdf_mgpu_get_comm_ms(tsdf_mgpu_handle h, float out_ms[8], int64_t out_count[8]);

#ifdef __cplusplus
}
#endif
#endif /* TSDF_B200_MGPU_H_ */
