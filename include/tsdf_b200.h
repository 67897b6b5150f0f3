/*
 * tsdf_b200.h -- C ABI of the B200-native voxel-hashed semantic TSDF engine (libtsdf_b200.so).
 *
 * This is the drop-in boundary for the reference's `TSDFGrid` class
 * (utils/tsdf/voxel_tsdf.cuh:32-124 of yuzhou42/disinfect-slam).  The reference has no FFI
 * layer -- its "operator API" is that C++ class -- so every entry point below names the
 * member function (file:line relative to the reference root) it replaces.  Plain pointers and
 * sizes only: no torch, Eigen or OpenCV types cross this boundary.  The C++17 shim that keeps
 * the reference's own signatures on top of this ABI is include/tsdf_b200/voxel_tsdf.hpp.
 *
 * Conventions
 *  - every function returns 0 on success, a negative TSDF_E_* code otherwise;
 *    tsdf_last_error() returns a thread-local description of the last failure.
 *  - poses are cam_T_world as unit quaternion (x,y,z,w) + translation, exactly the storage of
 *    SE3<float> (utils/cuda/lie_group.cuh:43-44); intrinsics are K = {fx, fy, cx, cy}
 *    (utils/cuda/camera.cuh:12-17).
 *  - images are row-major, continuous: rgb uint8 HxWx3 (RGB order), depth float32 metres,
 *    ht / lt float32 probabilities (utils/tsdf/voxel_tsdf.cuh:47-59).
 *  - the engine is not thread-safe (like TSDFGrid); callers serialise, but may call from any
 *    host thread (the engine sets its CUDA device on entry).
 *  - there is NO CPU fallback: every call needs the CUDA device chosen at tsdf_create.
 */
#ifndef TSDF_B200_H_
#define TSDF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDF_B200_ABI_VERSION 1

enum {
  TSDF_OK = 0,
  TSDF_E_INVALID = -1,        /* bad argument (null pointer, size mismatch, image too large) */
  TSDF_E_CUDA = -2,           /* CUDA runtime error; see tsdf_last_error() */
  TSDF_E_POOL_EXHAUSTED = -3, /* block pool ran out (reference: device assert, voxel_mem.cu:39) */
  TSDF_E_TABLE_FULL = -4,     /* hash table ran out of slots */
  TSDF_E_NO_DEVICE = -5,      /* no usable CUDA device: the engine never falls back to the CPU */
  TSDF_E_EXCHANGE_FULL = -6   /* candidate exchange of a sharded volume overflowed (tsdf_alloc_exchange_attach: cap_keys) */
};

typedef struct tsdf_engine* tsdf_handle;

/* Runtime replacements for the reference's compile-time #defines
 * (NUM_BLOCK utils/tsdf/voxel_mem.cuh:11-12, NUM_ENTRY utils/tsdf/voxel_hash.cuh:13-25,
 *  MAX_IMG_SIZE utils/tsdf/voxel_tsdf.cu:10-12). */
#define TSDF_FLAG_SHARD_SHIFT_MASK 0xF
#define TSDF_FLAG_BLOCKING_SYNC 0x10 /* host waits yield the CPU (cudaEventBlockingSync) instead of spinning */

typedef struct tsdf_config {
  int32_t struct_size;      /* = sizeof(tsdf_config) */
  int32_t device;           /* CUDA ordinal; -1 = current device */
  int32_t pool_blocks;      /* 8^3-voxel blocks in the pool (default 1 << 18) */
  int32_t table_slots;      /* open-addressing slots, power of two (default 1 << 21) */
  int32_t max_image_pixels; /* largest W*H accepted (default 1920 * 1080) */
  int32_t shard_rank;       /* multi-GPU block ownership: this engine keeps only blocks with */
  int32_t shard_count;      /*   owner(block) == shard_rank of shard_count (default 0 of 1) */
  int32_t flags;            /* bits 0..3: shard granularity shift s -- (1 << s)^3 neighbouring blocks share an owner;
                               TSDF_FLAG_BLOCKING_SYNC */
} tsdf_config;

/* Counters of the last tsdf_integrate* call (what the reference only logs through
 * spdlog::debug, voxel_tsdf.cu:367,374,470). */
typedef struct tsdf_counters {
  int64_t n_active_pre;  /* active blocks before the frame */
  int64_t n_new;         /* blocks allocated by this frame */
  int64_t n_visible;     /* blocks that passed the any-corner visibility test */
  int64_t n_updated;     /* voxel updates (voxels that passed every test in the integrate kernel) */
  int64_t n_carved;      /* blocks freed by space carving */
  int64_t n_active_post; /* active blocks after the frame */
  int64_t n_candidates;  /* de-duplicated allocation requests probed in the table */
  int64_t reserved;
} tsdf_counters;

const char* tsdf_last_error(void);
int tsdf_abi_version(void);

int tsdf_default_config(tsdf_config* cfg);

/* TSDFGrid::TSDFGrid(voxel_size, truncation)            utils/tsdf/voxel_tsdf.cu:309-326
 * + VoxelHashTable() / VoxelMemPool()                   utils/tsdf/voxel_hash.cu:37-45, voxel_mem.cu:13-27 */
int tsdf_create(float voxel_size, float truncation, const tsdf_config* cfg /* may be NULL */, tsdf_handle* out);
/* TSDFGrid::~TSDFGrid()                                 utils/tsdf/voxel_tsdf.cu:328-345 */
int tsdf_destroy(tsdf_handle h);

/* TSDFGrid::Integrate(rgb, depth, ht, lt, max_depth, intrinsics, cam_T_world)
 *                                                       utils/tsdf/voxel_tsdf.cu:347-375
 * Host buffers (pinned or pageable); synchronous like the reference: the buffers may be reused
 * on return and the volume is up to date. */
int tsdf_integrate(tsdf_handle h, const uint8_t* rgb, const float* depth, const float* ht, const float* lt,
                   int width, int height, float max_depth, const float K[4], const float q_xyzw[4],
                   const float t_xyz[3]);
/* Pipelined variant: returns as soon as the host buffers have been consumed (copied to one of
 * two device staging sets); the kernels of this frame overlap the next call's upload.  Call
 * tsdf_synchronize() before reading counters / results.  Same arithmetic as tsdf_integrate.
 * Errors of the pipelined calls: TSDF_E_POOL_EXHAUSTED / TSDF_E_TABLE_FULL describe a frame submitted two calls
 * earlier (the one whose staging set this call reuses) and are reported exactly once; the frame passed to THIS call
 * has been enqueued all the same, and later frames run normally once blocks have been freed. */
int tsdf_integrate_async(tsdf_handle h, const uint8_t* rgb, const float* depth, const float* ht, const float* lt,
                         int width, int height, float max_depth, const float K[4], const float q_xyzw[4],
                         const float t_xyz[3]);
/* tsdf_integrate_async without the wait for the upload: returns as soon as everything is enqueued (no host wait at all
 * unless the frame two calls back is still running).  The host buffers must be pinned (tsdf_host_alloc) and stay
 * untouched until two further frames have been submitted or tsdf_synchronize() has returned.  ht == lt == NULL: no
 * probability planes (= planes of ones, TSDFSystem's default, modules/tsdf_module.cc:28-33), nothing is uploaded
 * for them.  This is the call a host thread that drives several engines (streams) uses. */
int tsdf_integrate_enqueue(tsdf_handle h, const uint8_t* rgb, const float* depth, const float* ht, const float* lt,
                           int width, int height, float max_depth, const float K[4], const float q_xyzw[4],
                           const float t_xyz[3]);
/* The sensor's own formats: 16-bit depth (metres = value / depthmap_factor) and 16-bit probabilities (value / 65535),
 * as the reference reads them from a log or a camera before `convertTo(CV_32FC1, 1. / scale)`
 * (examples/tsdf/offline.cc:72-83, cameras/l515.cc:24-31).  9 bytes per pixel cross PCIe instead of 15 (5 with
 * ht == lt == NULL); the conversion runs inside the allocation kernel with convertTo's arithmetic for CV_16U
 * (float32 pixel * float32 factor), so the volume is bit-identical to converting on the host and calling
 * tsdf_integrate.  flags: 0 = synchronous like tsdf_integrate, TSDF_FRAME_ASYNC = like tsdf_integrate_async,
 * TSDF_FRAME_NOWAIT = like tsdf_integrate_enqueue. */
#define TSDF_FRAME_ASYNC 1
#define TSDF_FRAME_NOWAIT 2
int tsdf_integrate_u16(tsdf_handle h, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht, const uint16_t* lt,
                       int width, int height, float depthmap_factor, float max_depth, const float K[4],
                       const float q_xyzw[4], const float t_xyz[3], int flags);
/* Several independent engines (streams) of one GPU driven by the calling thread alone: for i in [first, first + count)
 * and every stream b, frame = frames[b * n_frames + i % n_frames] goes through tsdf_integrate_enqueue (or
 * tsdf_integrate_u16 with TSDF_FRAME_NOWAIT, per frame format) and, if raycast != 0, through tsdf_raycast_async from
 * the frame's camera into host image set 2 * b + (i & 1) (each of rgba / normal / hit_depth may be NULL: that
 * image is not downloaded; TSDFGrid::RayCast itself delivers rgba + normal); the images of step i - 1 are waited for
 * during step i, every engine is synchronised at the end.  All host memory must be pinned.  This replaces one
 * blocking host thread per stream (what a TSDFSystem worker is, modules/tsdf_module.cc:51-75) where many streams
 * share a host. */
#define TSDF_FORMAT_F32 0
#define TSDF_FORMAT_U16 1
typedef struct tsdf_host_frame {
  const void *rgb, *depth, *ht, *lt; /* pinned host memory; ht == lt == NULL: no probability planes */
  float q_xyzw[4];
  float t_xyz[3];
  int32_t format; /* TSDF_FORMAT_F32: float32 depth / ht / lt; TSDF_FORMAT_U16: uint16 planes */
} tsdf_host_frame;
int tsdf_streams_run(int n_streams, const tsdf_handle* engines, const tsdf_host_frame* frames, int n_frames, int first,
                     int count, int width, int height, float depthmap_factor, float max_depth, const float K[4],
                     int raycast, uint8_t* const* rgba, uint8_t* const* normal, float* const* hit_depth);
/* Same, with the four planes already in device memory of the engine's GPU (SURVEY.md 8f rank 4:
 * the segmentation net produces ht/lt on the GPU; multi-GPU ranks receive the broadcast frame
 * in device memory).  Enqueues on the engine stream and returns without synchronising;
 * `after_event` (cudaEvent_t or NULL) is waited on by the engine stream first. */
int tsdf_integrate_device(tsdf_handle h, const void* d_rgb, const void* d_depth, const void* d_ht, const void* d_lt,
                          int width, int height, float max_depth, const float K[4], const float q_xyzw[4],
                          const float t_xyz[3], void* after_event);

/* TSDFGrid::RayCast(max_depth, virtual_cam, cam_T_world, rgba, normal)
 *                                                       utils/tsdf/voxel_tsdf.cu:490-506
 * The reference writes two uchar4 images into GL textures; here they are copied to host memory
 * (each pointer optional).  hit_depth (new, needed for multi-GPU nearest-hit compositing) is the
 * camera-space z of the refined hit in metres, +inf where the ray misses. */
int tsdf_raycast(tsdf_handle h, float max_depth, int width, int height, const float K[4], const float q_xyzw[4],
                 const float t_xyz[3], uint8_t* rgba /* HxWx4 */, uint8_t* normal /* HxWx4 */,
                 float* hit_depth /* HxW */);
/* Pipelined variant: enqueues the render and the device->host copies (on their own stream, from one of two engine-owned
 * image sets) and returns; a following tsdf_integrate_async overlaps the copies.  tsdf_raycast_wait() blocks until the
 * images of the OLDEST outstanding call are in host memory; at most two calls are outstanding (a third waits for the
 * first).  The host buffers should be pinned (tsdf_host_alloc) and must stay untouched until waited for.
 * tsdf_synchronize() and every synchronous call wait for all of them. */
int tsdf_raycast_async(tsdf_handle h, float max_depth, int width, int height, const float K[4], const float q_xyzw[4],
                       const float t_xyz[3], uint8_t* rgba, uint8_t* normal, float* hit_depth);
int tsdf_raycast_wait(tsdf_handle h);
/* Device-output variant: results stay on the GPU (replacement for GLImage8UC4::LoadCuda,
 * utils/gl/image.cc:108-119).  Pointers are device memory, each optional; if `packed_min_keys`
 * is non-NULL it receives per ray two uint64 = (float_bits(hit_depth) << 32) | rgba / normal, the
 * format used for nearest-hit min-compositing across GPUs.  Asynchronous on the engine stream. */
int tsdf_raycast_device(tsdf_handle h, float max_depth, int width, int height, const float K[4],
                        const float q_xyzw[4], const float t_xyz[3], void* d_rgba, void* d_normal,
                        void* d_hit_depth, void* d_packed_min_keys);

/* Engine-owned output variant: renders into the engine's own device images (the reference's
 * img_tsdf_rgba_ / img_tsdf_normal_, voxel_tsdf.cuh:121-122) and returns their device addresses,
 * valid until the next raycast -- what a GLImage8UC4::LoadCuda-style sink consumes.  Asynchronous
 * on the engine stream; any output pointer may be NULL. */
int tsdf_raycast_resident(tsdf_handle h, float max_depth, int width, int height, const float K[4],
                          const float q_xyzw[4], const float t_xyz[3], const void** d_rgba, const void** d_normal,
                          const void** d_hit_depth);

/* Shared-volume RayCast for a volume sharded over several engines (no reference counterpart: the reference is
 * single-GPU).  Every engine exports its table / pool (CUDA IPC, one blob per engine), attaches the blobs of all
 * shards (index = shard rank; same shard_count and granularity everywhere), and can then render any rows of a
 * view over the WHOLE volume: blocks of other shards are read from their owner's memory over NVLink inside the
 * march kernel, so the result is bit-identical to a single-engine render.  The caller must make sure (a barrier)
 * that no shard is integrating while another one renders.  tsdf_peer_attach_local does the same for engines that
 * live in one process on one device (tests).  Output pointers are device memory holding the full HxW images;
 * only rows [row0, row0 + rows) are written.  Asynchronous on the engine stream. */
#define TSDF_IPC_BLOB_BYTES 320
int tsdf_ipc_export(tsdf_handle h, void* blob /* TSDF_IPC_BLOB_BYTES */);
int tsdf_ipc_attach(tsdf_handle h, int shard_count, const void* blobs /* shard_count x TSDF_IPC_BLOB_BYTES */);
int tsdf_peer_attach_local(tsdf_handle h, int shard_count, const tsdf_handle* shards);
int tsdf_raycast_shared(tsdf_handle h, float max_depth, int width, int height, const float K[4], const float q_xyzw[4],
                        const float t_xyz[3], int row0, int rows, void* d_rgba, void* d_normal, void* d_hit_depth);
/* Same march, with the exchange of the results fused into the kernel: every finished ray is stored into the HxW images
 * of n_dest destinations (device pointers, typically the image buffers of every rank mapped as peer memory; entries or
 * whole arrays may be NULL) -- posted stores over NVLink while the march runs, instead of a local image plus an
 * all-gather afterwards.  The launch renders tile_count (0 = all that remain) of the 8-row tiles tile_first,
 * tile_first + tile_stride, ... of the view: tile_first = rank, tile_stride = number of ranks deals the tiles of a view
 * out round-robin (every rank gets the same mix of cheap and expensive rows, but meets almost every visible block);
 * tile_stride = 1 with tile_count tiles per rank gives every rank one contiguous band of rows, whose rays meet about
 * 1 / ranks of the visible blocks (what tsdf_shared_cache_attach is made for).  peers_unchanged != 0: the caller
 * vouches that no shard has integrated since this engine's previous shared view (then not even the map-maintenance
 * kernels, which would find nothing to do, are launched).  The caller orders the destinations' readers with its own
 * barrier. */
int tsdf_raycast_shared_scatter(tsdf_handle h, float max_depth, int width, int height, const float K[4],
                                const float q_xyzw[4], const float t_xyz[3], int tile_first, int tile_stride, int tile_count,
                                int peers_unchanged, int n_dest, void* const* d_rgba, void* const* d_normal,
                                void* const* d_hit_depth);

/* TSDF mirrors for a sharded volume (optional; call on an idle engine, before the first frame).  mirrors[r] = rank r's
 * mirror buffer as this GPU addresses it (peer-mapped; mirrors[shard_rank] is this engine's own): shard_count x
 * stride_blocks x 512 floats, stride_blocks >= the largest pool of any shard.  From then on the integrate kernel stores
 * every TSDF value it writes into slot [shard_rank][pool index] of ALL mirrors too (posted stores over NVLink, 16 bytes
 * per updated voxel quad and rank), and tsdf_raycast_shared* reads every TSDF sample from the local mirror instead of
 * from the owner's memory over NVLink; only the colour and probability of the hit voxel still come from the owner.
 * It trades 2 KB per block of ANY shard on every GPU (a third of the voxel data, replicated) for a march without remote
 * loads.  world == 0 detaches.  Blocks written through tsdf_assign_voxels are not mirrored. */
int tsdf_mirror_attach(tsdf_handle h, int world, void* const* mirrors, int stride_blocks);

/* Pulled TSDF cache for a sharded volume (optional, the alternative to mirrors; call on an idle engine after the peers
 * are attached).  The engine allocates a local array of shard_count x stride_blocks x 2 KB (+ 4 B stamps);
 * stride_blocks >= the largest pool of any shard.  Before every shared-volume march, two small kernels fetch the TSDF
 * planes of the FOREIGN blocks that the rays of this launch can meet -- a conservative pyramid test of the launch's rows
 * against every shard's pool directory -- with bulk NVLink reads, and stamp them; the march then samples them locally.
 * A sample of a foreign block without a current stamp is read from its owner, so results do not depend on the test.
 * Entries stay valid across views while the caller passes peers_unchanged = 1.  What a view can meet is bounded by
 * max_depth, about a room's worth of blocks (a few MB per view and rank); a contiguous band of rows (tile_stride 1)
 * needs about 1/shard_count of that.  Mirrors move about as little (only updated voxels travel) and measured faster
 * (DESIGN.md section 7); the cache keeps the integrate kernel free of remote stores and is the starting point for a
 * bounded-size cache (today it is addressed like a mirror and as large).
 * pad_voxels = how far the test grows every block (3 covers the nearest-voxel rounding, the
 * gradient samples and the float accumulation; smaller or negative values are still exact, only slower).
 * stride_blocks == 0 detaches and frees. */
int tsdf_shared_cache_attach(tsdf_handle h, int stride_blocks, int pad_voxels);
/* how many foreign blocks (2 KB each) the most recent shared-volume view of this engine fetched; waits for the stream */
int tsdf_shared_cache_stats(tsdf_handle h, int64_t* blocks_fetched_last_view);

/* Candidate exchange for a sharded volume (optional; call on an idle engine before the first frame).  Without it every
 * rank walks the depth band of ALL pixel rays and keeps the candidate blocks it owns (block_allocate_kernel,
 * voxel_tsdf.cu:104-147, replicated).  With it, rank r stages the whole frame but walks only the rays of every
 * shard_count-th 32 x 8 pixel tile, and mails each candidate key to its owner: inboxes[o] = rank o's inbox as this GPU
 * addresses it (peer-mapped; inboxes[shard_rank] is the own one), tsdf_alloc_exchange_bytes(world, cap_keys) bytes,
 * zero-initialised.  Then the frame hook runs on the engine's stream -- it must publish d_cursor[o] (clamped to cap_keys)
 * into int word [parity * 8 + shard_rank] of inbox o and order all ranks (a barrier with release / acquire semantics at
 * system scope) -- and the owner inserts what it received.  All ranks must integrate the same frames in the same order.
 * Results are those of the unsharded volume.  A rank that would mail more than cap_keys keys to one owner in one frame
 * reports TSDF_E_EXCHANGE_FULL for that frame.  world == 0 detaches. */
typedef void (*tsdf_frame_hook)(void* user, void* cuda_stream, const int* d_cursor, int parity);
size_t tsdf_alloc_exchange_bytes(int world, int cap_keys);
int tsdf_alloc_exchange_attach(tsdf_handle h, int world, void* const* inboxes, int cap_keys, tsdf_frame_hook hook, void* user);

/* TSDFGrid::GatherValid()                               utils/tsdf/voxel_tsdf.cu:399-425
 * TSDFGrid::GatherVoxels(BoundingCube<float>)           utils/tsdf/voxel_tsdf.cu:427-454
 * out = array of VoxelSpatialTSDF {float x, y, z, tsdf} (utils/tsdf/voxel_types.cuh:48-57),
 * 512 consecutive records per selected block in x + 8y + 64z order; block order is unspecified
 * (hash-layout dependent in the reference too).  bbox = {xmin,xmax,ymin,ymax,zmin,zmax} metres,
 * member order of BoundingCube (voxel_tsdf.cuh:12-19).  *n_voxels receives the number selected;
 * at most cap_voxels are written; out == NULL only counts. */
int tsdf_gather_valid(tsdf_handle h, float* out_xyzt, int64_t cap_voxels, int64_t* n_voxels);
int tsdf_gather_in_bound(tsdf_handle h, const float bbox[6], float* out_xyzt, int64_t cap_voxels, int64_t* n_voxels);
/* The gathers always leave their full result in an engine-owned device buffer (valid until the
 * next gather): call with out_xyzt == NULL to learn *n_voxels, size the host array, then
 * tsdf_gather_fetch() copies it without selecting again; or consume it on the GPU directly. */
int tsdf_gather_fetch(tsdf_handle h, float* out_xyzt, int64_t cap_voxels);
int tsdf_gather_device_result(tsdf_handle h, const void** d_out_xyzt, int64_t* n_voxels);

/* Triangle mesh of the zero level set, extracted on the GPU from the blocks the same bbox would select in
 * tsdf_gather_in_bound (bbox == NULL: every block).  Replaces the reference's mesh path -- GatherVoxels' 16 B per
 * voxel download followed by KrisLibrary's Geometry::SparseTSDFReconstruction::ExtractMesh on one CPU core
 * (examples/ros_camera_driver/ros_offline.cc:258-318, 320-350): the meshing itself moves to the GPU (the triangle
 * soup of a full-resolution surface is about as large as the voxels it came from; keep it on the device with
 * tsdf_mesh_device_result when the consumer is on the GPU).
 * out_xyz = n_triangles x 3 vertices x (x, y, z) metres, voxel centres at (grid + 0.5) * voxel_size
 * (ros_offline.cc:281-284), normals (counter-clockwise order) towards free space; triangle order unspecified.
 * Cells with an unallocated or never-observed (weight 0) corner are not meshed.  Two-call protocol like the
 * gathers: out_xyz == NULL only counts; the full result stays in an engine-owned device buffer until the next
 * extraction (tsdf_mesh_fetch / tsdf_mesh_device_result). */
int tsdf_extract_mesh(tsdf_handle h, const float* bbox /* 6 floats or NULL */, float* out_xyz, int64_t cap_triangles,
                      int64_t* n_triangles);
int tsdf_mesh_fetch(tsdf_handle h, float* out_xyz, int64_t cap_triangles);
int tsdf_mesh_device_result(tsdf_handle h, const void** d_out_xyz, int64_t* n_triangles);

/* VoxelHashTable::NumActiveBlock()                      utils/tsdf/voxel_hash.cu:200 */
int tsdf_num_active_blocks(tsdf_handle h, int* n);
int tsdf_get_counters(tsdf_handle h, tsdf_counters* out);
/* RayCast keeps an empty-space skip map over the block set (no reference counterpart: ray_cast_kernel,
 * voxel_tsdf.cu:232-307, probes the table at every sample).  *attempts = how often a RayCast found that a mutating
 * call had run since the map was built, *rebuilds = how often the block set had really changed (frames that only
 * allocate-and-carve the same edge blocks leave it valid). */
int tsdf_get_skip_map_stats(tsdf_handle h, int64_t* attempts, int64_t* rebuilds);
int tsdf_synchronize(tsdf_handle h);
/* cudaStream_t the engine enqueues on (for event interop with callers that own device data). */
void* tsdf_stream(tsdf_handle h);

/* Hash(block_pos) & BUCKET_MASK                         utils/tsdf/voxel_hash.cu:31-35
 * (21-bit mask of the reference; the engine uses the same mix, masked to its own table size). */
uint32_t tsdf_hash(int16_t bx, int16_t by, int16_t bz);

/* Rank that owns block (bx, by, bz) when the volume is sharded over shard_count engines with
 * granularity shift shard_shift (pure host function; the same mix the kernels use). */
int tsdf_block_owner(int16_t bx, int16_t by, int16_t bz, int shard_count, int shard_shift);

/* Parity / unit-test access (what utils/tests/voxel_hash_test.cu:36-55 does with its own
 * Allocate / Retrieve / Assignment kernels, and voxel_mem_test.cu with Aquire/Release).
 * keys / points are int16 triples.  Entry points marked TSDF_TEST_API exist for the test suites and for
 * checkpoint-style export: they drain the pipeline and allocate scratch device memory on every call, so they do
 * not belong in a per-frame loop.  Duplicate keys in one list are allowed (each block is inserted / released once). */
#define TSDF_TEST_API
TSDF_TEST_API int tsdf_allocate_blocks(tsdf_handle h, const int16_t* block_keys, int n); /* VoxelHashTable::Allocate, voxel_hash.cu:58 */
TSDF_TEST_API int tsdf_delete_blocks(tsdf_handle h, const int16_t* block_keys, int n);   /* VoxelHashTable::Delete,   voxel_hash.cu:122 */
/* VoxelHashTable::Retrieve<T>, voxel_hash.cuh:104-113: absent -> tsdf 1, rgbw 0, prob 0, found 0 */
TSDF_TEST_API int tsdf_retrieve_voxels(tsdf_handle h, const int16_t* points, int n, float* tsdf, uint8_t* rgbw /* n x 4 */,
                         float* prob, int32_t* found);
/* VoxelHashTable::RetrieveMutable + store, voxel_hash.cuh:124-161; any value pointer may be NULL */
TSDF_TEST_API int tsdf_assign_voxels(tsdf_handle h, const int16_t* points, int n, const float* tsdf, const uint8_t* rgbw,
                       const float* prob);
/* All active blocks in canonical order (ascending z, y, x block coordinate): keys int16[n][3],
 * tsdf float[n][512], rgbw uint8[n][512][4] (r,g,b,weight), prob float[n][512]; any may be NULL. */
TSDF_TEST_API int tsdf_export_blocks(tsdf_handle h, int16_t* keys, float* tsdf, uint8_t* rgbw, float* prob, int cap_blocks,
                       int* n_blocks);

/* Pinned host memory helpers so callers can hand tsdf_integrate DMA-able buffers. */
int tsdf_host_alloc(void** ptr, size_t bytes);
int tsdf_host_free(void* ptr);

/* Profiling.  While enabled, every phase is bracketed by CUDA events on the stream it is launched
 * on; nothing synchronises until the getters run.  Every retired frame's counters are summed
 * (profiling on or off); tsdf_set_profiling() resets the sums.  out_ms = device milliseconds summed over all calls since
 * then: [0] upload, [1] frame staging + allocate, [2] select visible, [3] integrate + carve,
 * [4] raycast (skip-map build + march), [5] gather; out_count (optional) = number of timed launches per phase.
 * tsdf_get_totals: sums of the per-frame counters and the number of frames. */
int tsdf_set_profiling(tsdf_handle h, int enabled);
int tsdf_get_phase_ms(tsdf_handle h, float out_ms[8], int64_t out_count[8]);
int tsdf_get_totals(tsdf_handle h, tsdf_counters* sums, int64_t* n_frames);

#ifdef __cplusplus
}
#endif
#endif /* TSDF_B200_H_ */
