// engine.cu -- host side of libtsdf_b200.so: buffer ownership, streams, launch sequencing and the
// C ABI declared in include/tsdf_b200.h.  Replaces the host part of TSDFGrid
// (utils/tsdf/voxel_tsdf.cu:309-506): no per-frame host synchronisation inside Integrate except
// the final one the synchronous API requires, no cudaMalloc/cudaFree per gather, double-buffered
// frame staging so that the upload of frame k+1 overlaps the kernels of frame k.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <numeric>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <unistd.h>

#include "../../include/tsdf_b200.h"
#include "tsdf_device.cuh"
#include "tsdf_launch.h"

using namespace tsdf;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
#define CU(call)                                                                                            \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) return fail(TSDF_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                                 \
  } while (0)

namespace {
enum { PH_UPLOAD = 0, PH_ALLOC, PH_SELECT, PH_INTEGRATE, PH_RAYCAST, PH_GATHER, PH_COUNT };

struct FrameBuf {
  unsigned char* rgb = nullptr; float *depth = nullptr, *ht = nullptr, *lt = nullptr;
  unsigned char* packed = nullptr;
  Texel* tex = nullptr;
  cudaEvent_t uploaded = nullptr, done = nullptr;
  int* h_ctr = nullptr;  // pinned (mapped) copy of the device counters after this frame
  int* d_h_ctr = nullptr;  // the same memory as the device addresses it
  bool in_flight = false;
};
}  // namespace

struct tsdf_engine {
  int device = 0, num_sms = 148;
  float voxel_size = 0, truncation = 0;
  tsdf_config cfg{};
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  DeviceState S{};
  FrameBuf fb[2];
  int cur = 0;
  int last_slot = -1;  // slot of the most recent frame (its counters are the "last" ones)
  int* visible = nullptr; int* selected = nullptr;
  bool blocking_sync = false;
  cudaEvent_t ev_block = nullptr;
  SkipMap skip{};                    // RayCast empty-space skip map, rebuilt when the block set changed
  PeerView* d_self = nullptr;        // device copy of this engine's own PeerView (1 entry)
  PeerView* d_peers = nullptr;       // device array [shard_count]: every shard of a volume sharded over GPUs
  PeerView h_peers[kMaxPeers] = {};  // the same on the host (pool pointers travel as kernel parameters)
  uint64_t shared_epoch = 0;         // volume_epoch at the last shared-map build attempt (see tsdf_raycast_shared_scatter)
  int n_peers = 0;
  void* ipc_opened[kMaxPeers][4] = {};  // pointers obtained from cudaIpcOpenMemHandle (closed at destroy)
  uint64_t volume_epoch = 1, skip_epoch = 0;  // host side: a mutating call was enqueued since the map was last looked at
  int skip_gen = 0;    // number of skip-map build attempts (see skip_fill_kernel)
  int serial = 0;      // number of mutating calls (frames, allocate / delete lists); DeviceState::serial of the current one
  // pipelined RayCast (tsdf_raycast_async): two sets of output images (set 0 = the three above, set 1 allocated on first
  // use), rendered on `stream`, copied to the host on `d2h_stream` while the next frame's kernels run
  // each set is ONE block [rgba | normal | hit depth] laid out for the current image size, so that host images that are
  // adjacent in memory are downloaded with one transfer
  struct RcSet { unsigned char* block = nullptr; cudaEvent_t rendered = nullptr, copied = nullptr; bool pending = false; } rc[2];
  int rc_cur = 0;
  cudaStream_t d2h_stream = nullptr;
  float4* gather_out = nullptr; size_t gather_cap = 0; int64_t gather_n = 0;
  float* mesh_out = nullptr; size_t mesh_cap = 0; int64_t mesh_n = 0;  // triangles (9 floats each), grow-only
  unsigned long long* mesh_counter = nullptr;
  int* h_scalar = nullptr;  // pinned scratch (C_COUNT ints)
  unsigned char* bounce[2] = {nullptr, nullptr};  // pinned staging for large device -> pageable-host results (see copy_to_host)
  cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
  // pulled TSDF cache of a sharded volume (tsdf_shared_cache_attach)
  float* cache = nullptr; int* cache_stamp = nullptr; int* cache_list = nullptr; int* cache_count = nullptr;
  int cache_stride = 0, cache_epoch = 1, cache_serial = 0, cache_pad = 3;
  // candidate exchange of a sharded volume (tsdf_alloc_exchange_attach)
  int* xa_cursor = nullptr; tsdf_frame_hook xa_hook = nullptr; void* xa_user = nullptr; unsigned xa_frames = 0;
  int n_active = 0;         // host mirror after the last completed frame
  tsdf_counters last{};
  bool profiling = false;
  // phase profiling: (begin, end) event pairs recorded on the launching stream, summed lazily
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs[PH_COUNT];
  cudaEvent_t ev_open[PH_COUNT] = {};
  double phase_total_ms[PH_COUNT] = {};
  int64_t phase_count[PH_COUNT] = {};
  tsdf_counters totals{};  // sums over the frames retired since tsdf_set_profiling(1)
  int64_t total_frames = 0;
};

// ---- small host helpers (same float32 arithmetic as the reference's host code) -------------------
static Intr intr_inverse(const Intr& k) {  // utils/cuda/camera.cuh:35-39
  const float fxi = 1 / k.fx, fyi = 1 / k.fy;
  Intr r; r.fx = fxi; r.fy = fyi; r.cx = -k.cx * fxi; r.cy = -k.cy * fyi; return r;
}
static Pose pose_inverse(const Pose& T) {  // utils/cuda/lie_group.cuh:25-27 + Eigen quaternion inverse
  Pose r;
  const float n2 = (T.qx * T.qx + T.qy * T.qy) + (T.qz * T.qz + T.qw * T.qw);
  if (n2 > 0.f) { r.qx = (-T.qx) / n2; r.qy = (-T.qy) / n2; r.qz = (-T.qz) / n2; r.qw = T.qw / n2; }
  else { r.qx = r.qy = r.qz = r.qw = 0.f; }
  r.tx = r.ty = r.tz = 0.f;
  const float3 t = qrot(r, f3(-T.tx, -T.ty, -T.tz));
  r.tx = t.x; r.ty = t.y; r.tz = t.z;
  return r;
}
static FrameParams make_params(const tsdf_engine* e, int w, int h, float max_depth, const float K[4], const float q[4],
                               const float t[3]) {
  FrameParams P;
  P.cam_T_world = Pose{q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
  P.world_T_cam = pose_inverse(P.cam_T_world);
  P.K = Intr{K[0], K[1], K[2], K[3]};
  P.Kinv = intr_inverse(P.K);
  P.w = w; P.h = h; P.max_depth = max_depth; P.voxel_size = e->voxel_size; P.truncation = e->truncation; P.neg_zero = -0.0f;
  return P;
}
static cudaEvent_t ev_get(tsdf_engine* e) {
  if (!e->ev_pool.empty()) { cudaEvent_t v = e->ev_pool.back(); e->ev_pool.pop_back(); return v; }
  cudaEvent_t v = nullptr; cudaEventCreate(&v); return v;
}
static void phase_begin(tsdf_engine* e, int ph, cudaStream_t st) {
  if (!e->profiling) return;
  e->ev_open[ph] = ev_get(e);
  cudaEventRecord(e->ev_open[ph], st);
}
static void phase_end(tsdf_engine* e, int ph, cudaStream_t st) {
  if (!e->profiling || !e->ev_open[ph]) return;
  cudaEvent_t b = ev_get(e);
  cudaEventRecord(b, st);
  e->ev_pairs[ph].emplace_back(e->ev_open[ph], b);
  e->ev_open[ph] = nullptr;
}
// fold all completed event pairs into the per-phase totals (call after the streams are idle)
static void phase_collect(tsdf_engine* e) {
  for (int ph = 0; ph < PH_COUNT; ++ph) {
    for (auto& pr : e->ev_pairs[ph]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { e->phase_total_ms[ph] += ms; e->phase_count[ph]++; }
      else cudaGetLastError();
      e->ev_pool.push_back(pr.first); e->ev_pool.push_back(pr.second);
    }
    e->ev_pairs[ph].clear();
  }
}

// Host wait for a stream.  Default: cudaStreamSynchronize (spins, lowest latency).  With TSDF_FLAG_BLOCKING_SYNC
// the wait goes through an event created with cudaEventBlockingSync, which yields the CPU -- for hosts that run
// more engine threads than they have cores.
static cudaError_t wait_stream(tsdf_engine* e, cudaStream_t st) {
  if (!e->blocking_sync) return cudaStreamSynchronize(st);
  cudaError_t rc = cudaEventRecord(e->ev_block, st);
  return rc != cudaSuccess ? rc : cudaEventSynchronize(e->ev_block);
}

static int check_frame_args(tsdf_engine* e, const void* a, const void* b, const void* c, const void* d, int w, int h,
                            const float* K, const float* q, const float* t) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if (!a || !b || !c || !d || !K || !q || !t) return fail(TSDF_E_INVALID, "null image / camera pointer");
  if (w <= 0 || h <= 0 || (int64_t)w * h > e->cfg.max_image_pixels)
    return fail(TSDF_E_INVALID, "image %dx%d exceeds max_image_pixels=%d", w, h, e->cfg.max_image_pixels);
  return TSDF_OK;
}

static void next_serial(tsdf_engine* e) {
  e->serial = e->serial >= 0x7FFFFFF0 ? 1 : e->serial + 1;
  e->S.serial = e->serial;
}

// kernels of one frame, enqueued on the compute stream; no host synchronisation
static void enqueue_frame(tsdf_engine* e, const FrameParams& P, const FrameInput& in, FrameBuf& f) {
  next_serial(e);  // (the per-call counters are zero here: cleared by the previous frame's publish kernel / after_mutation)
  if (e->S.xa_on) e->S.xa_parity = (int)(e->xa_frames++ & 1u);
  phase_begin(e, PH_ALLOC, e->stream);
  launch_frame_allocate(e->S, P, in, f.tex, e->stream);
  phase_end(e, PH_ALLOC, e->stream);
  // candidate exchange: counts published + all ranks ordered by the data plane's hook, then the owner's inserts
  if (e->S.xa_on) e->xa_hook(e->xa_user, (void*)e->stream, e->xa_cursor + e->S.xa_parity * 8, e->S.xa_parity);
  phase_begin(e, PH_SELECT, e->stream);
  if (e->S.xa_on) launch_insert_candidates(e->S, P, e->stream);
  launch_select_visible(e->S, P, e->visible, e->visible + e->cfg.pool_blocks, e->num_sms, e->stream);
  phase_end(e, PH_SELECT, e->stream);
  phase_begin(e, PH_INTEGRATE, e->stream);
  launch_integrate_carve(e->S, P, e->visible, e->visible + e->cfg.pool_blocks, f.tex, e->num_sms, (int)e->last.n_visible, e->stream);
  phase_end(e, PH_INTEGRATE, e->stream);
  launch_publish_counters(e->S, f.d_h_ctr, e->stream);  // a store into mapped host memory: no copy engine on the compute stream
  cudaEventRecord(f.done, e->stream);
  f.in_flight = true;
  e->volume_epoch++;
}
static FrameInput f32_input(const void* rgb, const void* depth, const void* ht, const void* lt) {
  FrameInput in{};
  in.rgb = (const unsigned char*)rgb; in.depth = depth; in.ht = ht; in.lt = lt;
  return in;
}

// wait for the frame that used slot `s`, fold its counters into the host mirror, surface errors
static int retire_slot(tsdf_engine* e, int s) {
  FrameBuf& f = e->fb[s];
  if (!f.in_flight) return TSDF_OK;
  CU(cudaEventSynchronize(f.done));
  f.in_flight = false;
  const int* c = f.h_ctr;
  tsdf_counters k{};
  k.n_active_pre = e->n_active;  // frames retire in submission order
  k.n_new = c[C_NNEW]; k.n_visible = c[C_NVIS]; k.n_carved = c[C_NCARVED]; k.n_candidates = c[C_NCAND];
  k.n_updated = (int64_t)(((u64)(unsigned)c[C_NUPD_HI] << 32) | (unsigned)c[C_NUPD_LO]);
  k.n_active_post = e->cfg.pool_blocks - c[C_FREE];
  e->last = k;
  e->n_active = (int)k.n_active_post;
  {  // sums since the last tsdf_set_profiling() call, kept whether or not phase timing is on
    e->totals.n_new += k.n_new; e->totals.n_visible += k.n_visible; e->totals.n_updated += k.n_updated;
    e->totals.n_carved += k.n_carved; e->totals.n_candidates += k.n_candidates;
    e->totals.n_active_post += k.n_active_post; e->totals.n_active_pre += k.n_active_pre;
    e->total_frames++;
  }
  // tombstone garbage collection once live + tombstoned slots exceed half of the table
  if ((unsigned)c[C_NONEMPTY] > (e->S.table_mask + 1) / 2) launch_rehash(e->S, e->num_sms, e->stream);
  // C_ERROR holds the bits of THIS frame only (it is zeroed with the per-call counters), so an exhaustion is reported
  // exactly once: the blocks that found no room are missing from this frame, later frames run normally (and succeed
  // once carving or tsdf_delete_blocks has freed blocks)
  if (c[C_ERROR] & ERR_POOL) return fail(TSDF_E_POOL_EXHAUSTED, "voxel block pool exhausted (pool_blocks=%d): some blocks of a frame were not allocated", e->cfg.pool_blocks);
  if (c[C_ERROR] & ERR_TABLE) return fail(TSDF_E_TABLE_FULL, "hash table full (table_slots=%d): some blocks of a frame were not allocated", e->cfg.table_slots);
  if (c[C_ERROR] & ERR_EXCHANGE) return fail(TSDF_E_EXCHANGE_FULL, "candidate exchange full (cap_keys=%d): some candidate blocks of a frame were dropped", e->S.xa_cap);
  return TSDF_OK;
}
// host wait for every outstanding tsdf_raycast_async copy
static int raycast_drain(tsdf_engine* e) {
  for (int i = 0; i < 2; ++i) {
    if (e->rc[i].pending) { CU(cudaEventSynchronize(e->rc[i].copied)); e->rc[i].pending = false; }
  }
  return TSDF_OK;
}
static int drain(tsdf_engine* e) {
  int rc = raycast_drain(e);
  if (rc) return rc;
  // retire in submission order so that `last` ends up describing the newest frame
  const int first = e->last_slot < 0 ? 0 : 1 - e->last_slot;
  for (int i = 0; i < 2; ++i) { const int r = retire_slot(e, (first + i) & 1); if (r != TSDF_OK) rc = r; }
  CU(wait_stream(e, e->stream));
  return rc;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)); }
};

extern "C" {

const char* tsdf_last_error(void) { return g_err; }
int tsdf_abi_version(void) { return TSDF_B200_ABI_VERSION; }

int tsdf_default_config(tsdf_config* cfg) {
  if (!cfg) return fail(TSDF_E_INVALID, "null config");
  memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = (int32_t)sizeof(tsdf_config);
  cfg->device = -1;
  cfg->pool_blocks = 1 << 18;
  cfg->table_slots = 1 << 21;
  cfg->max_image_pixels = 1920 * 1080;
  cfg->shard_rank = 0; cfg->shard_count = 1; cfg->flags = 0;
  return TSDF_OK;
}

int tsdf_create(float voxel_size, float truncation, const tsdf_config* user_cfg, tsdf_handle* out) {
  if (!out) return fail(TSDF_E_INVALID, "null output handle");
  *out = nullptr;
  if (!(voxel_size > 0.f) || !(truncation > 0.f)) return fail(TSDF_E_INVALID, "voxel_size and truncation must be > 0");
  tsdf_config cfg; tsdf_default_config(&cfg);
  if (user_cfg) {
    if (user_cfg->struct_size != (int32_t)sizeof(tsdf_config)) return fail(TSDF_E_INVALID, "tsdf_config.struct_size mismatch");
    cfg = *user_cfg;
  }
  if (cfg.pool_blocks <= 0 || cfg.pool_blocks >= (1 << kIndexShardShift) || cfg.table_slots <= 0 || (cfg.table_slots & (cfg.table_slots - 1)))
    return fail(TSDF_E_INVALID, "pool_blocks must be > 0 and table_slots a power of two");
  if (cfg.table_slots < 2 * (int64_t)cfg.pool_blocks) return fail(TSDF_E_INVALID, "table_slots must be >= 2 * pool_blocks");
  if (cfg.shard_count < 1 || cfg.shard_rank < 0 || cfg.shard_rank >= cfg.shard_count)
    return fail(TSDF_E_INVALID, "bad shard_rank / shard_count");
  if (cfg.max_image_pixels <= 0) return fail(TSDF_E_INVALID, "max_image_pixels must be > 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TSDF_E_NO_DEVICE, "no CUDA device available (this engine has no CPU fallback)");
  }
  if (cfg.device < 0) CU(cudaGetDevice(&cfg.device));
  if (cfg.device >= ndev) return fail(TSDF_E_INVALID, "device %d out of range (%d devices)", cfg.device, ndev);
  CU(cudaSetDevice(cfg.device));
  cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, cfg.device));
  if (prop.major != 10) return fail(TSDF_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg.device, prop.major, prop.minor);

  tsdf_engine* e = new tsdf_engine();
  e->device = cfg.device; e->num_sms = prop.multiProcessorCount; e->voxel_size = voxel_size; e->truncation = truncation; e->cfg = cfg;
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int rc_ = fail(TSDF_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); tsdf_destroy(e); return rc_; } } while (0)
  CUX(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUX(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  DeviceState& S = e->S;
  S.table_mask = (unsigned)cfg.table_slots - 1; S.pool_blocks = cfg.pool_blocks;
  S.shard_rank = cfg.shard_rank; S.shard_count = cfg.shard_count; S.shard_shift = cfg.flags & TSDF_FLAG_SHARD_SHIFT_MASK;
  e->blocking_sync = (cfg.flags & TSDF_FLAG_BLOCKING_SYNC) != 0;
  CUX(cudaEventCreateWithFlags(&e->ev_block, cudaEventDisableTiming | cudaEventBlockingSync));
  CUX(cudaMalloc(&S.table, sizeof(Slot) * (size_t)cfg.table_slots));
  CUX(cudaMalloc(&S.block_key, sizeof(u64) * (size_t)cfg.pool_blocks));
  CUX(cudaMalloc(&S.voxels, (size_t)kBlockBytes * (size_t)cfg.pool_blocks));
  CUX(cudaMalloc(&S.free_stack, sizeof(int) * (size_t)cfg.pool_blocks));
  CUX(cudaMalloc(&S.ctr, sizeof(int) * C_COUNT));
  CUX(cudaMalloc(&e->visible, sizeof(int) * 2 * (size_t)cfg.pool_blocks));  // [visible pool indices | per-entry carve state]
  CUX(cudaMalloc(&e->selected, sizeof(int) * (size_t)cfg.pool_blocks));
  CUX(cudaMalloc(&e->d_self, sizeof(PeerView))); CUX(cudaMalloc(&e->d_peers, sizeof(PeerView) * kMaxPeers));
  {
    PeerView v{}; v.table = S.table; v.table_mask = S.table_mask; v.block_key = S.block_key; v.voxels = S.voxels; v.ctr = S.ctr;
    CUX(cudaMemcpyAsync(e->d_self, &v, sizeof(v), cudaMemcpyHostToDevice, e->stream));
  }
  CUX(cudaMalloc(&e->skip.dist, kSkipMaxCells)); CUX(cudaMalloc(&e->skip.scratch, kSkipMaxCells));
  CUX(cudaMalloc(&e->skip.hdr, sizeof(int) * kSkipHdrInts));
  CUX(cudaMemsetAsync(e->skip.hdr, 0, sizeof(int) * kSkipHdrInts, e->stream));
  CUX(cudaMalloc(&e->skip.cells, sizeof(int) * (size_t)kSkipMaxCells));
  // both byte planes start entirely at the cap: the rebuild kernels keep the plane they will mark next in that state
  CUX(cudaMemsetAsync(e->skip.dist, kSkipCap, kSkipMaxCells, e->stream)); CUX(cudaMemsetAsync(e->skip.scratch, kSkipCap, kSkipMaxCells, e->stream));
  const size_t npx = (size_t)cfg.max_image_pixels;
  for (int i = 0; i < 2; ++i) {
    FrameBuf& f = e->fb[i];
    CUX(cudaMalloc(&f.packed, 15 * npx + 64));  // a frame whose planes are packed back to back in host memory arrives here with ONE copy
    CUX(cudaMalloc(&f.rgb, 3 * npx)); CUX(cudaMalloc(&f.depth, 4 * npx)); CUX(cudaMalloc(&f.ht, 4 * npx)); CUX(cudaMalloc(&f.lt, 4 * npx));
    CUX(cudaMalloc(&f.tex, sizeof(Texel) * npx));
    const unsigned evf = cudaEventDisableTiming | (e->blocking_sync ? cudaEventBlockingSync : 0u);
    CUX(cudaEventCreateWithFlags(&f.uploaded, evf));
    CUX(cudaEventCreateWithFlags(&f.done, evf));
    CUX(cudaHostAlloc(&f.h_ctr, sizeof(int) * C_COUNT, cudaHostAllocMapped));
    memset(f.h_ctr, 0, sizeof(int) * C_COUNT);
    CUX(cudaHostGetDevicePointer(&f.d_h_ctr, f.h_ctr, 0));
  }
  CUX(cudaMalloc(&e->rc[0].block, 12 * npx));
  CUX(cudaMallocHost(&e->h_scalar, sizeof(int) * C_COUNT));
  launch_init_state(S, e->stream);
  CUX(cudaGetLastError());
  CUX(cudaStreamSynchronize(e->stream));
#undef CUX
  *out = e;
  return TSDF_OK;
}

int tsdf_destroy(tsdf_handle e) {
  if (!e) return TSDF_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
  if (e->d2h_stream) { cudaStreamSynchronize(e->d2h_stream); cudaStreamDestroy(e->d2h_stream); }
  cudaFree(e->rc[0].block); cudaFree(e->rc[1].block);
  for (int i = 0; i < 2; ++i) { if (e->rc[i].rendered) cudaEventDestroy(e->rc[i].rendered); if (e->rc[i].copied) cudaEventDestroy(e->rc[i].copied); }
  cudaFree(e->S.table); cudaFree(e->S.block_key); cudaFree(e->S.voxels); cudaFree(e->S.free_stack); cudaFree(e->S.ctr);
  cudaFree(e->visible); cudaFree(e->selected); cudaFree(e->mesh_out); cudaFree(e->mesh_counter);
  cudaFree(e->skip.dist); cudaFree(e->skip.scratch); cudaFree(e->skip.hdr); cudaFree(e->skip.cells);
  cudaFree(e->cache); cudaFree(e->cache_stamp); cudaFree(e->cache_list); cudaFree(e->cache_count); cudaFree(e->xa_cursor);
  for (int r = 0; r < kMaxPeers; ++r) for (int k = 0; k < 4; ++k) if (e->ipc_opened[r][k]) cudaIpcCloseMemHandle(e->ipc_opened[r][k]);
  cudaFree(e->d_self); cudaFree(e->d_peers);
  for (int i = 0; i < 2; ++i) {
    FrameBuf& f = e->fb[i];
    cudaFree(f.packed); cudaFree(f.rgb); cudaFree(f.depth); cudaFree(f.ht); cudaFree(f.lt); cudaFree(f.tex);
    if (f.uploaded) cudaEventDestroy(f.uploaded);
    if (f.done) cudaEventDestroy(f.done);
    if (f.h_ctr) cudaFreeHost(f.h_ctr);
  }
  cudaFree(e->gather_out);
  if (e->h_scalar) cudaFreeHost(e->h_scalar);
  for (int i = 0; i < 2; ++i) { if (e->bounce[i]) cudaFreeHost(e->bounce[i]); if (e->bounce_ev[i]) cudaEventDestroy(e->bounce_ev[i]); }
  if (e->ev_block) cudaEventDestroy(e->ev_block);
  phase_collect(e);
  for (cudaEvent_t v : e->ev_pool) cudaEventDestroy(v);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  delete e;
  return TSDF_OK;
}

// Host frame -> staging set -> kernels.  Planes: rgb u8x3 always; depth / ht / lt float32 (bytes_px = 4) or uint16
// (bytes_px = 2, converted inside the allocation kernel); ht == lt == nullptr: no probability planes are uploaded at all.
// wait_upload: block until the DMA engine has read the host buffers (they may then be reused: what tsdf_integrate_async
// promises); without it the call returns at once and the buffers must stay untouched until the frame has retired.
static int submit_host_frame(tsdf_engine* e, const uint8_t* rgb, const void* depth, const void* ht, const void* lt, int depth_u16,
                             int prob_u16, float depth_scale, float prob_scale, int w, int h, float max_depth, const float K[4],
                             const float q[4], const float t[3], bool wait_upload) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if (!rgb || !depth || !K || !q || !t || ((ht == nullptr) != (lt == nullptr))) return fail(TSDF_E_INVALID, "null image / camera pointer");
  if (w <= 0 || h <= 0 || (int64_t)w * h > e->cfg.max_image_pixels)
    return fail(TSDF_E_INVALID, "image %dx%d exceeds max_image_pixels=%d", w, h, e->cfg.max_image_pixels);
  CU(cudaSetDevice(e->device));
  const int s = e->cur;
  FrameBuf& f = e->fb[s];
  // the frame two calls ago: bounds the pipeline depth, frees the staging set.  An exhaustion it reports belongs to
  // THAT frame: the present frame is enqueued all the same and the status is returned afterwards.
  const int rc_old = retire_slot(e, s);
  if (rc_old == TSDF_E_CUDA) return rc_old;
  const size_t n = (size_t)w * h;
  // uploads on the copy stream (they overlap the previous frame's kernels on the compute stream)
  phase_begin(e, PH_UPLOAD, e->copy_stream);
  const size_t b_rgb = 3 * n, b_d = (depth_u16 ? 2 : 4) * n, b_p = (prob_u16 ? 2 : 4) * n;
  const unsigned char *d_rgb = f.rgb, *d_depth = (const unsigned char*)f.depth, *d_ht = (const unsigned char*)f.ht, *d_lt = (const unsigned char*)f.lt;
  // Planes packed back to back in host memory ([rgb | depth | ht | lt], e.g. one pinned block per frame filled by the
  // capture threads) travel as ONE DMA transfer: with uploads and image downloads sharing the link, a few large
  // transfers reach ~48 GB/s per direction on this platform where four 2 MB ones reach ~34 (profiles/).
  const bool packed = (b_rgb % 4 == 0) && (const unsigned char*)depth == rgb + b_rgb &&
                      (!ht || ((const unsigned char*)ht == (const unsigned char*)depth + b_d && (const unsigned char*)lt == (const unsigned char*)ht + b_p));
  bool packed_done = false;
  if (packed) {  // (adjacent planes of separate pinned allocations cannot travel in one transfer: the runtime refuses, see below)
    packed_done = cudaMemcpyAsync(f.packed, rgb, b_rgb + b_d + (ht ? 2 * b_p : 0), cudaMemcpyHostToDevice, e->copy_stream) == cudaSuccess;
    if (!packed_done) cudaGetLastError();
  }
  if (packed_done) {
    d_rgb = f.packed; d_depth = f.packed + b_rgb; d_ht = d_depth + b_d; d_lt = d_ht + b_p;
  } else {
    CU(cudaMemcpyAsync(f.rgb, rgb, b_rgb, cudaMemcpyHostToDevice, e->copy_stream));
    CU(cudaMemcpyAsync(f.depth, depth, b_d, cudaMemcpyHostToDevice, e->copy_stream));
    if (ht) {
      CU(cudaMemcpyAsync(f.ht, ht, b_p, cudaMemcpyHostToDevice, e->copy_stream));
      CU(cudaMemcpyAsync(f.lt, lt, b_p, cudaMemcpyHostToDevice, e->copy_stream));
    }
  }
  phase_end(e, PH_UPLOAD, e->copy_stream);
  CU(cudaEventRecord(f.uploaded, e->copy_stream));
  CU(cudaStreamWaitEvent(e->stream, f.uploaded, 0));
  const FrameParams P = make_params(e, w, h, max_depth, K, q, t);
  FrameInput in = f32_input(d_rgb, d_depth, ht ? d_ht : nullptr, ht ? d_lt : nullptr);
  in.depth_u16 = depth_u16; in.prob_u16 = prob_u16; in.depth_scale = depth_scale; in.prob_scale = prob_scale;
  enqueue_frame(e, P, in, f);
  CU(cudaGetLastError());
  e->last_slot = s;
  e->cur = 1 - s;
  if (wait_upload) CU(cudaEventSynchronize(f.uploaded));  // the host buffers are reusable once the copies have been issued from them
  return rc_old;
}

int tsdf_integrate_async(tsdf_handle e, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int w,
                         int h, float max_depth, const float K[4], const float q[4], const float t[3]) {
  if (!ht || !lt) return fail(TSDF_E_INVALID, "null image / camera pointer");
  return submit_host_frame(e, rgb, depth, ht, lt, 0, 0, 0.f, 0.f, w, h, max_depth, K, q, t, true);
}

int tsdf_integrate_enqueue(tsdf_handle e, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int w,
                           int h, float max_depth, const float K[4], const float q[4], const float t[3]) {
  return submit_host_frame(e, rgb, depth, ht, lt, 0, 0, 0.f, 0.f, w, h, max_depth, K, q, t, false);
}

int tsdf_integrate_u16(tsdf_handle e, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht, const uint16_t* lt, int w,
                       int h, float depthmap_factor, float max_depth, const float K[4], const float q[4], const float t[3], int flags) {
  if (!(depthmap_factor > 0.f)) return fail(TSDF_E_INVALID, "depthmap_factor must be > 0");
  // img.convertTo(CV_32FC1, 1. / depth_scale) and (..., 1. / 65535): the double factor is cast to float inside OpenCV
  const float ds = (float)(1. / (double)depthmap_factor), ps = (float)(1. / 65535);
  const int rc = submit_host_frame(e, rgb, depth, ht, lt, 1, 1, ds, ps, w, h, max_depth, K, q, t, (flags & TSDF_FRAME_NOWAIT) == 0);
  if (flags & (TSDF_FRAME_ASYNC | TSDF_FRAME_NOWAIT)) return rc;
  if (rc == TSDF_E_INVALID || rc == TSDF_E_CUDA || rc == TSDF_E_NO_DEVICE) return rc;
  const int rc2 = drain(e);
  return rc2 ? rc2 : rc;
}

int tsdf_integrate(tsdf_handle e, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int w, int h,
                   float max_depth, const float K[4], const float q[4], const float t[3]) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  const int rc = tsdf_integrate_async(e, rgb, depth, ht, lt, w, h, max_depth, K, q, t);
  if (rc == TSDF_E_INVALID || rc == TSDF_E_CUDA || rc == TSDF_E_NO_DEVICE) return rc;
  const int rc2 = drain(e);  // rc, if set, is the exhaustion of an earlier asynchronous frame: this frame ran
  return rc2 ? rc2 : rc;
}

int tsdf_integrate_device(tsdf_handle e, const void* d_rgb, const void* d_depth, const void* d_ht, const void* d_lt, int w,
                          int h, float max_depth, const float K[4], const float q[4], const float t[3], void* after_event) {
  int rc = check_frame_args(e, d_rgb, d_depth, d_ht, d_lt, w, h, K, q, t);
  if (rc) return rc;
  CU(cudaSetDevice(e->device));
  const int s = e->cur;
  FrameBuf& f = e->fb[s];
  const int rc_old = retire_slot(e, s);  // an exhaustion of the frame two calls ago: reported below, this frame still runs
  if (rc_old == TSDF_E_CUDA) return rc_old;
  if (after_event) CU(cudaStreamWaitEvent(e->stream, (cudaEvent_t)after_event, 0));
  const FrameParams P = make_params(e, w, h, max_depth, K, q, t);
  enqueue_frame(e, P, f32_input(d_rgb, d_depth, d_ht, d_lt), f);
  CU(cudaGetLastError());
  e->last_slot = s;
  e->cur = 1 - s;
  return rc_old;
}

// A few host threads, many independent engines (streams) on one GPU -- the loop a multi-camera server runs.  Per step and
// stream: enqueue the frame (no wait: pinned buffers), enqueue the view and its copies to pinned host memory, then make
// sure the PREVIOUS step's images of that stream have arrived (they were enqueued a whole step ago, so this rarely
// blocks).  Nothing in the loop waits for the GPU to finish the step it has just enqueued.  One thread issues about
// thirty CUDA calls per frame, i.e. it feeds ~4 k frames/s; the streams are therefore split over `threads` host threads
// (default 2, TSDF_STREAMS_THREADS) -- still far fewer than one blocking thread per stream.
static int streams_worker(int n_streams, const tsdf_handle* engines, const tsdf_host_frame* frames, int n_frames, int first, int count,
                          int w, int h, float depthmap_factor, float max_depth, const float K[4], int raycast, uint8_t* const* rgba,
                          uint8_t* const* normal, float* const* hit_depth, int b0, int b1, char* err, size_t err_len) {
  int deferred = TSDF_OK;
  auto bail = [&](int rc) { snprintf(err, err_len, "%s", tsdf_last_error()); return rc; };
  for (int i = first; i < first + count; ++i) {
    for (int b = b0; b < b1; ++b) {
      const tsdf_host_frame& f = frames[(size_t)b * n_frames + (i % n_frames)];
      int rc;
      if (f.format == TSDF_FORMAT_U16)
        rc = tsdf_integrate_u16(engines[b], (const uint8_t*)f.rgb, (const uint16_t*)f.depth, (const uint16_t*)f.ht, (const uint16_t*)f.lt, w, h,
                                depthmap_factor, max_depth, K, f.q_xyzw, f.t_xyz, TSDF_FRAME_NOWAIT);
      else
        rc = tsdf_integrate_enqueue(engines[b], (const uint8_t*)f.rgb, (const float*)f.depth, (const float*)f.ht, (const float*)f.lt, w, h,
                                    max_depth, K, f.q_xyzw, f.t_xyz);
      if (rc == TSDF_E_POOL_EXHAUSTED || rc == TSDF_E_TABLE_FULL) { deferred = rc; rc = TSDF_OK; }
      if (rc) return bail(rc);
      if (raycast) {
        const int o = 2 * b + (i & 1);  // two sets of host images per stream, used alternately
        rc = tsdf_raycast_async(engines[b], max_depth, w, h, K, f.q_xyzw, f.t_xyz, rgba ? rgba[o] : nullptr, normal ? normal[o] : nullptr,
                                hit_depth ? hit_depth[o] : nullptr);
        if (rc) return bail(rc);
        if (i > first) { rc = tsdf_raycast_wait(engines[b]); if (rc) return bail(rc); }
      }
    }
  }
  for (int b = b0; b < b1; ++b) {
    const int rc = tsdf_synchronize(engines[b]);
    if (rc == TSDF_E_POOL_EXHAUSTED || rc == TSDF_E_TABLE_FULL) deferred = rc;
    else if (rc) return bail(rc);
  }
  if (deferred) snprintf(err, err_len, "%s", tsdf_last_error());
  return deferred;
}

int tsdf_streams_run(int n_streams, const tsdf_handle* engines, const tsdf_host_frame* frames, int n_frames, int first, int count,
                     int w, int h, float depthmap_factor, float max_depth, const float K[4], int raycast, uint8_t* const* rgba,
                     uint8_t* const* normal, float* const* hit_depth) {
  if (n_streams <= 0 || !engines || !frames || n_frames <= 0 || first < 0 || count < 0 || !K) return fail(TSDF_E_INVALID, "bad argument");
  int n_threads = 2;
  if (const char* v = getenv("TSDF_STREAMS_THREADS")) n_threads = atoi(v);
  n_threads = std::max(1, std::min(n_threads, n_streams));
  std::vector<int> rcs(n_threads, TSDF_OK);
  std::vector<std::vector<char>> errs(n_threads, std::vector<char>(512, 0));
  auto part = [&](int t) {
    const int b0 = (int)((long long)n_streams * t / n_threads), b1 = (int)((long long)n_streams * (t + 1) / n_threads);
    rcs[t] = streams_worker(n_streams, engines, frames, n_frames, first, count, w, h, depthmap_factor, max_depth, K, raycast, rgba, normal,
                            hit_depth, b0, b1, errs[t].data(), errs[t].size());
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(part, t);
  part(0);
  for (auto& x : th) x.join();
  for (int t = 0; t < n_threads; ++t)
    if (rcs[t] != TSDF_OK && rcs[t] != TSDF_E_POOL_EXHAUSTED && rcs[t] != TSDF_E_TABLE_FULL) return fail(rcs[t], "%s", errs[t].data());
  for (int t = 0; t < n_threads; ++t)
    if (rcs[t] != TSDF_OK) return fail(rcs[t], "%s", errs[t].data());
  return TSDF_OK;
}

int tsdf_synchronize(tsdf_handle e) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  return drain(e);
}

void* tsdf_stream(tsdf_handle e) { return e ? (void*)e->stream : nullptr; }

int tsdf_raycast_device(tsdf_handle e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                        void* d_rgba, void* d_normal, void* d_hit_depth, void* d_packed) {
  if (!e || !K || !q || !t) return fail(TSDF_E_INVALID, "null argument");
  if (w <= 0 || h <= 0) return fail(TSDF_E_INVALID, "bad image size %dx%d", w, h);
  CU(cudaSetDevice(e->device));
  const FrameParams P = make_params(e, w, h, max_depth, K, q, t);
  phase_begin(e, PH_RAYCAST, e->stream);
  if (e->skip_epoch != e->volume_epoch) {  // a mutating call ran since the map was built; the kernels themselves
    // return at once if that call turned out not to change the block set (device-side serial, skip_fill_kernel)
    launch_build_skip_map(e->d_self, 1, e->skip, ++e->skip_gen, true, e->num_sms, e->stream);
    e->skip_epoch = e->volume_epoch;
  }
  launch_raycast(e->S, P, e->truncation / 2, e->skip, (uchar4*)d_rgba, (uchar4*)d_normal, (float*)d_hit_depth,
                 (unsigned long long*)d_packed, e->stream);  // step = truncation / 2, voxel_tsdf.cu:497
  phase_end(e, PH_RAYCAST, e->stream);
  CU(cudaGetLastError());
  return TSDF_OK;
}

// [rgba | normal | hit depth] of n pixels each -> host.  Host images that follow one another in memory go in one transfer.
static cudaError_t download_images(tsdf_engine* e, const unsigned char* block, size_t n, uint8_t* rgba, uint8_t* normal, float* hit_depth,
                                   cudaStream_t st) {
  (void)e;
  cudaError_t r = cudaSuccess;
  if (rgba && normal == rgba + 4 * n && (!hit_depth || (uint8_t*)hit_depth == normal + 4 * n)) {
    // adjacent addresses may still belong to separate pinned allocations, which one transfer cannot span: the runtime
    // then refuses (nothing is enqueued) and the images go one by one
    if (cudaMemcpyAsync(rgba, block, (hit_depth ? 12 : 8) * n, cudaMemcpyDeviceToHost, st) == cudaSuccess) return cudaSuccess;
    cudaGetLastError();
  }
  if (rgba) r = cudaMemcpyAsync(rgba, block, 4 * n, cudaMemcpyDeviceToHost, st);
  if (r == cudaSuccess && normal) r = cudaMemcpyAsync(normal, block + 4 * n, 4 * n, cudaMemcpyDeviceToHost, st);
  if (r == cudaSuccess && hit_depth) r = cudaMemcpyAsync(hit_depth, block + 8 * n, 4 * n, cudaMemcpyDeviceToHost, st);
  return r;
}

int tsdf_raycast_async(tsdf_handle e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                       uint8_t* rgba, uint8_t* normal, float* hit_depth) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if ((int64_t)w * h > e->cfg.max_image_pixels) return fail(TSDF_E_INVALID, "image %dx%d exceeds max_image_pixels=%d", w, h, e->cfg.max_image_pixels);
  CU(cudaSetDevice(e->device));
  if (!e->d2h_stream) {  // first use: second image set, copy stream, events
    const size_t npx = (size_t)e->cfg.max_image_pixels;
    CU(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&e->rc[1].block, 12 * npx));
    const unsigned evf = cudaEventDisableTiming | (e->blocking_sync ? cudaEventBlockingSync : 0u);
    for (int i = 0; i < 2; ++i) { CU(cudaEventCreateWithFlags(&e->rc[i].rendered, evf)); CU(cudaEventCreateWithFlags(&e->rc[i].copied, evf)); }
  }
  tsdf_engine::RcSet& o = e->rc[e->rc_cur];
  if (o.pending) { CU(cudaEventSynchronize(o.copied)); o.pending = false; }  // at most two views in flight
  const size_t n = (size_t)w * h;
  int rc = tsdf_raycast_device(e, max_depth, w, h, K, q, t, o.block, o.block + 4 * n, o.block + 8 * n, nullptr);
  if (rc) return rc;
  CU(cudaEventRecord(o.rendered, e->stream));
  CU(cudaStreamWaitEvent(e->d2h_stream, o.rendered, 0));
  CU(download_images(e, o.block, n, rgba, normal, hit_depth, e->d2h_stream));
  CU(cudaEventRecord(o.copied, e->d2h_stream));
  o.pending = true;  // the host waits for this copy before the set is rendered into again (two calls from now)
  e->rc_cur ^= 1;
  return TSDF_OK;
}
int tsdf_raycast_wait(tsdf_handle e) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  // oldest outstanding view first: after the flip rc_cur names the set used two calls ago
  for (int k = 0; k < 2; ++k) {
    tsdf_engine::RcSet& o = e->rc[(e->rc_cur + k) & 1];
    if (o.pending) { CU(cudaEventSynchronize(o.copied)); o.pending = false; return TSDF_OK; }
  }
  return TSDF_OK;
}

int tsdf_raycast(tsdf_handle e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                 uint8_t* rgba, uint8_t* normal, float* hit_depth) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if ((int64_t)w * h > e->cfg.max_image_pixels) return fail(TSDF_E_INVALID, "image %dx%d exceeds max_image_pixels=%d", w, h, e->cfg.max_image_pixels);
  { int rcd = raycast_drain(e); if (rcd) return rcd; }
  const size_t n = (size_t)w * h;
  unsigned char* const blk = e->rc[0].block;
  int rc = tsdf_raycast_device(e, max_depth, w, h, K, q, t, blk, blk + 4 * n, blk + 8 * n, nullptr);
  if (rc) return rc;
  CU(download_images(e, blk, n, rgba, normal, hit_depth, e->stream));
  CU(wait_stream(e, e->stream));
  return TSDF_OK;
}

int tsdf_raycast_resident(tsdf_handle e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                          const void** d_rgba, const void** d_normal, const void** d_hit_depth) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if ((int64_t)w * h > e->cfg.max_image_pixels) return fail(TSDF_E_INVALID, "image %dx%d exceeds max_image_pixels=%d", w, h, e->cfg.max_image_pixels);
  int rc = raycast_drain(e);
  if (rc) return rc;
  const size_t n = (size_t)w * h;
  unsigned char* const blk = e->rc[0].block;
  rc = tsdf_raycast_device(e, max_depth, w, h, K, q, t, blk, blk + 4 * n, blk + 8 * n, nullptr);
  if (rc) return rc;
  if (d_rgba) *d_rgba = blk;
  if (d_normal) *d_normal = blk + 4 * n;
  if (d_hit_depth) *d_hit_depth = blk + 8 * n;
  return TSDF_OK;
}

// ---- shared-volume RayCast: a volume sharded over several engines, each reading the others' memory ----------------
namespace {
struct IpcBlob {
  cudaIpcMemHandle_t table, block_key, voxels, ctr;
  uint32_t table_mask; int32_t pool_blocks, shard_rank, shard_count, shard_shift, device;
  int64_t pid;    // exporting process: CUDA IPC handles cannot be opened by the process that made them, so shards
  void* raw[4];   //   living in the attaching process (one host thread per GPU) are reached through these pointers
};
static_assert(sizeof(IpcBlob) <= TSDF_IPC_BLOB_BYTES, "TSDF_IPC_BLOB_BYTES too small");
}  // namespace

int tsdf_ipc_export(tsdf_handle e, void* blob) {
  if (!e || !blob) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  IpcBlob b{};
  CU(cudaIpcGetMemHandle(&b.table, e->S.table)); CU(cudaIpcGetMemHandle(&b.block_key, e->S.block_key));
  CU(cudaIpcGetMemHandle(&b.voxels, e->S.voxels)); CU(cudaIpcGetMemHandle(&b.ctr, e->S.ctr));
  b.table_mask = e->S.table_mask; b.pool_blocks = e->S.pool_blocks; b.shard_rank = e->S.shard_rank;
  b.shard_count = e->S.shard_count; b.shard_shift = e->S.shard_shift; b.device = e->device;
  b.pid = (int64_t)getpid();
  b.raw[0] = e->S.table; b.raw[1] = e->S.block_key; b.raw[2] = e->S.voxels; b.raw[3] = e->S.ctr;
  memset(blob, 0, TSDF_IPC_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return TSDF_OK;
}

static int check_peer_count(tsdf_engine* e, int world) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if (world != e->S.shard_count || world < 1 || world > kMaxPeers)
    return fail(TSDF_E_INVALID, "peer count %d must equal shard_count %d (at most %d)", world, e->S.shard_count, kMaxPeers);
  return TSDF_OK;
}
static PeerView self_view(const tsdf_engine* e) {
  PeerView v{}; v.table = e->S.table; v.table_mask = e->S.table_mask; v.block_key = e->S.block_key; v.voxels = e->S.voxels; v.ctr = e->S.ctr;
  return v;
}

int tsdf_ipc_attach(tsdf_handle e, int world, const void* blobs) {
  int rc = check_peer_count(e, world);
  if (rc) return rc;
  if (!blobs) return fail(TSDF_E_INVALID, "null blobs");
  CU(cudaSetDevice(e->device));
  PeerView views[kMaxPeers] = {};
  for (int r = 0; r < world; ++r) {
    IpcBlob b;
    memcpy(&b, (const char*)blobs + (size_t)r * TSDF_IPC_BLOB_BYTES, sizeof(b));
    if (b.shard_rank != r || b.shard_count != world || b.shard_shift != e->S.shard_shift)
      return fail(TSDF_E_INVALID, "blob %d describes shard %d of %d (shift %d)", r, b.shard_rank, b.shard_count, b.shard_shift);
    if (r == e->S.shard_rank) { views[r] = self_view(e); continue; }
    void* p[4] = {};
    const cudaIpcMemHandle_t* hs[4] = {&b.table, &b.block_key, &b.voxels, &b.ctr};
    if (b.pid == (int64_t)getpid()) {  // a shard of this very process: plain peer access, no IPC mapping
      if (b.device != e->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, e->device, b.device));
        if (!can) return fail(TSDF_E_CUDA, "device %d cannot access device %d as a peer", e->device, b.device);
        const cudaError_t pe = cudaDeviceEnablePeerAccess(b.device, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(TSDF_E_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(pe));
        cudaGetLastError();
      }
      for (int k = 0; k < 4; ++k) p[k] = b.raw[k];
    } else {
      for (int k = 0; k < 4; ++k) {
        if (e->ipc_opened[r][k]) { cudaIpcCloseMemHandle(e->ipc_opened[r][k]); e->ipc_opened[r][k] = nullptr; }
        CU(cudaIpcOpenMemHandle(&p[k], *hs[k], cudaIpcMemLazyEnablePeerAccess));
        e->ipc_opened[r][k] = p[k];
      }
    }
    views[r].table = (const Slot*)p[0]; views[r].table_mask = b.table_mask; views[r].block_key = (const u64*)p[1];
    views[r].voxels = (const unsigned char*)p[2]; views[r].ctr = (const int*)p[3];
  }
  CU(cudaMemcpy(e->d_peers, views, sizeof(PeerView) * world, cudaMemcpyHostToDevice));
  memcpy(e->h_peers, views, sizeof(PeerView) * world);
  e->n_peers = world;
  return TSDF_OK;
}

int tsdf_peer_attach_local(tsdf_handle e, int world, const tsdf_handle* shards) {
  int rc = check_peer_count(e, world);
  if (rc) return rc;
  if (!shards) return fail(TSDF_E_INVALID, "null shard list");
  CU(cudaSetDevice(e->device));
  PeerView views[kMaxPeers] = {};
  for (int r = 0; r < world; ++r) {
    const tsdf_engine* o = shards[r];
    if (!o || o->S.shard_rank != r || o->S.shard_count != world || o->S.shard_shift != e->S.shard_shift || o->device != e->device)
      return fail(TSDF_E_INVALID, "shard %d: not shard %d of %d on this device with the same granularity", r, r, world);
    views[r] = self_view(o);
  }
  CU(cudaMemcpy(e->d_peers, views, sizeof(PeerView) * world, cudaMemcpyHostToDevice));
  memcpy(e->h_peers, views, sizeof(PeerView) * world);
  e->n_peers = world;
  return TSDF_OK;
}

static int raycast_shared_impl(tsdf_engine* e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                               int row0, int rows, int tile_stride, bool peers_unchanged, void* d_rgba, void* d_normal, void* d_hit_depth,
                               int n_out, void* const* out_rgba, void* const* out_normal, void* const* out_depth) {
  if (!e || !K || !q || !t) return fail(TSDF_E_INVALID, "null argument");
  if (e->n_peers < 1) return fail(TSDF_E_INVALID, "no peers attached (tsdf_ipc_attach / tsdf_peer_attach_local)");
  if (w <= 0 || h <= 0 || row0 < 0 || rows < 0 || tile_stride < 1) return fail(TSDF_E_INVALID, "bad image size / row range");
  if (n_out < 0 || n_out > kMaxPeers) return fail(TSDF_E_INVALID, "at most %d destinations", kMaxPeers);
  CU(cudaSetDevice(e->device));
  const FrameParams P = make_params(e, w, h, max_depth, K, q, t);
  phase_begin(e, PH_RAYCAST, e->stream);
  // union of every shard's blocks.  Other shards change without this host knowing, so the attempt is always launched;
  // the kernels return at once when no shard's block set changed since the last build (device-side serials)
  // ... unless the caller vouches that no shard has integrated since this engine's last shared view (the data plane knows:
  // all ranks make the same calls) and this engine has not either: then even the four empty launches are saved
  if (!(peers_unchanged && e->shared_epoch == e->volume_epoch && e->skip_epoch == 0 && e->skip_gen > 0))
    launch_build_skip_map(e->d_peers, e->n_peers, e->skip, ++e->skip_gen, true, e->num_sms, e->stream);
  e->shared_epoch = e->volume_epoch;
  e->skip_epoch = 0;  // a local RayCast must look again: the map may describe more than this engine
  SharedCache C{};
  C.self = e->S.shard_rank;
  if (e->cache) {
    // content epoch: what was fetched stays valid until some shard integrates again (the caller knows: peers_unchanged)
    if (!peers_unchanged) {
      if (e->cache_epoch == 0x7FFFFFFF) { CU(cudaMemsetAsync(e->cache_stamp, 0, sizeof(int) * (size_t)e->n_peers * e->cache_stride, e->stream)); e->cache_epoch = 0; }
      e->cache_epoch++;
    }
    e->cache_serial = (e->cache_serial + 1) & 0xFFFF;
    C.cache = e->cache; C.stamp = e->cache_stamp; C.list = e->cache_list; C.count = e->cache_count;
    C.stride = e->cache_stride; C.epoch = e->cache_epoch; C.serial = e->cache_serial; C.pad = e->cache_pad;
  }
  launch_raycast_shared(e->d_peers, e->h_peers, e->n_peers, e->S.shard_shift, P, e->truncation / 2, e->skip, row0, std::min(rows, h - row0), tile_stride,
                        e->S.n_mirror ? e->S.mirror[e->S.shard_rank] : nullptr, e->S.mirror_stride, C, (uchar4*)d_rgba, (uchar4*)d_normal, (float*)d_hit_depth,
                        n_out, out_rgba, out_normal, out_depth, e->num_sms, e->stream);
  phase_end(e, PH_RAYCAST, e->stream);
  CU(cudaGetLastError());
  return TSDF_OK;
}
int tsdf_mirror_attach(tsdf_handle e, int world, void* const* mirrors, int stride_blocks) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  if (world == 0 || !mirrors) { e->S.n_mirror = 0; return TSDF_OK; }
  if (world != e->S.shard_count || world > kMaxPeers) return fail(TSDF_E_INVALID, "mirror count %d must equal shard_count %d", world, e->S.shard_count);
  if (stride_blocks < e->S.pool_blocks) return fail(TSDF_E_INVALID, "mirror stride %d smaller than the pool (%d blocks)", stride_blocks, e->S.pool_blocks);
  for (int r = 0; r < world; ++r) {
    if (!mirrors[r]) return fail(TSDF_E_INVALID, "null mirror pointer for rank %d", r);
    e->S.mirror[r] = (float*)mirrors[r];
  }
  e->S.n_mirror = world; e->S.mirror_stride = stride_blocks;
  return TSDF_OK;
}

static void free_shared_cache(tsdf_engine* e) {
  cudaFree(e->cache); cudaFree(e->cache_stamp); cudaFree(e->cache_list); cudaFree(e->cache_count);
  e->cache = nullptr; e->cache_stamp = nullptr; e->cache_list = nullptr; e->cache_count = nullptr; e->cache_stride = 0;
}
int tsdf_shared_cache_attach(tsdf_handle e, int stride_blocks, int pad_voxels) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  free_shared_cache(e);
  if (stride_blocks == 0) return TSDF_OK;
  if (e->n_peers < 2) return fail(TSDF_E_INVALID, "attach the peers first (tsdf_ipc_attach / tsdf_peer_attach_local), at least two shards");
  if (stride_blocks < e->S.pool_blocks || stride_blocks >= (1 << kIndexShardShift)) return fail(TSDF_E_INVALID, "cache stride %d smaller than the pool (%d blocks)", stride_blocks, e->S.pool_blocks);
  const size_t slots = (size_t)e->n_peers * stride_blocks;
  CU(cudaMalloc(&e->cache, slots * kBlockVolume * sizeof(float)));
  CU(cudaMalloc(&e->cache_stamp, slots * sizeof(int)));
  CU(cudaMalloc(&e->cache_list, slots * sizeof(int)));
  CU(cudaMalloc(&e->cache_count, 4 * sizeof(int)));
  CU(cudaMemsetAsync(e->cache_stamp, 0, slots * sizeof(int), e->stream));
  CU(cudaMemsetAsync(e->cache_count, 0, 4 * sizeof(int), e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->cache_stride = stride_blocks; e->cache_epoch = 1; e->cache_serial = 0; e->cache_pad = pad_voxels;
  return TSDF_OK;
}

int tsdf_shared_cache_stats(tsdf_handle e, int64_t* blocks_fetched_last_view) {
  if (!e || !blocks_fetched_last_view) return fail(TSDF_E_INVALID, "null argument");
  *blocks_fetched_last_view = 0;
  if (!e->cache) return TSDF_OK;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  int n = 0;
  CU(cudaMemcpy(&n, e->cache_count + (e->cache_serial & 3), sizeof(int), cudaMemcpyDeviceToHost));
  *blocks_fetched_last_view = n;
  return TSDF_OK;
}

size_t tsdf_alloc_exchange_bytes(int world, int cap_keys) {
  if (world < 1 || cap_keys < 1) return 0;
  return (size_t)kXaHeaderBytes + 2 * (size_t)world * (size_t)cap_keys * sizeof(u64);
}
int tsdf_alloc_exchange_attach(tsdf_handle e, int world, void* const* inboxes, int cap_keys, tsdf_frame_hook hook, void* user) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  if (world == 0 || !inboxes) { e->S.xa_on = 0; e->xa_hook = nullptr; return TSDF_OK; }
  if (world != e->S.shard_count || world > kMaxPeers || world < 2) return fail(TSDF_E_INVALID, "inbox count %d must equal shard_count %d (2..%d)", world, e->S.shard_count, kMaxPeers);
  if (cap_keys < 1 || !hook) return fail(TSDF_E_INVALID, "cap_keys must be > 0 and the frame hook non-null");
  for (int r = 0; r < world; ++r) if (!inboxes[r]) return fail(TSDF_E_INVALID, "null inbox pointer for rank %d", r);
  if (!e->xa_cursor) CU(cudaMalloc(&e->xa_cursor, 16 * sizeof(int)));
  CU(cudaMemsetAsync(e->xa_cursor, 0, 16 * sizeof(int), e->stream));
  CU(cudaStreamSynchronize(e->stream));
  for (int r = 0; r < world; ++r) e->S.xa_inbox[r] = (u64*)inboxes[r];
  e->S.xa_cursor = e->xa_cursor; e->S.xa_cap = cap_keys; e->S.xa_parity = 0; e->xa_frames = 0;
  e->xa_hook = hook; e->xa_user = user;
  e->S.xa_on = 1;
  return TSDF_OK;
}

int tsdf_raycast_shared(tsdf_handle e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                        int row0, int rows, void* d_rgba, void* d_normal, void* d_hit_depth) {
  return raycast_shared_impl(e, max_depth, w, h, K, q, t, row0, rows, 1, false, d_rgba, d_normal, d_hit_depth, 0, nullptr, nullptr, nullptr);
}
int tsdf_raycast_shared_scatter(tsdf_handle e, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                                int tile_first, int tile_stride, int tile_count, int peers_unchanged, int n_dest, void* const* d_rgba,
                                void* const* d_normal, void* const* d_hit_depth) {
  if (n_dest < 1) return fail(TSDF_E_INVALID, "need at least one destination");
  if (tile_first < 0 || tile_stride < 1 || tile_count < 0) return fail(TSDF_E_INVALID, "bad tile selection");
  if (tile_first * 8 >= h) return TSDF_OK;  // more ranks than tiles
  int rows = h - tile_first * 8;  // the span of rows the launch may touch
  if (tile_count > 0) rows = std::min<long long>(rows, ((long long)(tile_count - 1) * tile_stride + 1) * 8);
  return raycast_shared_impl(e, max_depth, w, h, K, q, t, tile_first * 8, rows, tile_stride, peers_unchanged != 0, nullptr, nullptr,
                             nullptr, n_dest, d_rgba, d_normal, d_hit_depth);
}

// Large results (GatherValid / GatherVoxels records, meshes) to host memory at PCIe speed.
//  * pinned destination (tsdf_host_alloc, cudaHostRegister): one asynchronous copy -- the DMA engine writes the
//    caller's memory directly;
//  * pageable destination (the std::vector the reference's signatures return, a fresh numpy array): the copy is
//    pipelined in 16 MB chunks through two engine-owned pinned buffers -- the DMA of chunk i+1 runs while four host
//    threads move chunk i into the destination.  A fresh allocation is all page faults (expensive on a virtualised
//    host), so every thread first asks the kernel to populate its part in one batch (MADV_POPULATE_WRITE; ignored
//    where unsupported).  TSDF_PAGEABLE_COPY=driver selects the plain cudaMemcpy of the reference
//    (voxel_tsdf.cu:420-422,449-451) instead.
static int copy_to_host(tsdf_engine* e, void* dst, const void* d_src, size_t bytes) {
  if (!bytes) return TSDF_OK;
  cudaPointerAttributes at{};
  const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  constexpr size_t kChunk = (size_t)16 << 20;
  static const bool use_driver = [] { const char* v = getenv("TSDF_PAGEABLE_COPY"); return v && !strcmp(v, "driver"); }();
  if (pinned || bytes <= kChunk / 4 || use_driver) {
    CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(wait_stream(e, e->stream));
    return TSDF_OK;
  }
  for (int i = 0; i < 2; ++i) {
    if (!e->bounce[i]) { CU(cudaMallocHost(&e->bounce[i], kChunk)); CU(cudaEventCreateWithFlags(&e->bounce_ev[i], cudaEventDisableTiming)); }
  }
  const int n_chunks = (int)((bytes + kChunk - 1) / kChunk);
  constexpr int kThreads = 4;
  std::atomic<int> ready{0};               // chunks [0, ready) are in their bounce buffer
  std::atomic<int> finished[2] = {{0}, {0}};  // helpers done with the chunk currently in bounce[p]
  auto share = [&](int chunk, int t) {     // thread t's part of `chunk`: a contiguous quarter, page aligned
    const size_t off = (size_t)chunk * kChunk, len = std::min(kChunk, bytes - off);
    const size_t per = ((len / kThreads) + 4095) & ~(size_t)4095;
    const size_t a = std::min(len, per * t), b = t == kThreads - 1 ? len : std::min(len, per * (t + 1));
    if (b > a) {
#ifdef MADV_POPULATE_WRITE
      const uintptr_t lo = ((uintptr_t)dst + off + a + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)dst + off + b) & ~(uintptr_t)4095;
      if (hi > lo) madvise((void*)lo, hi - lo, MADV_POPULATE_WRITE);  // batch the page faults of a fresh destination
#endif
      memcpy((char*)dst + off + a, e->bounce[chunk & 1] + a, b - a);
    }
  };
  std::vector<std::thread> helpers;
  for (int t = 1; t < kThreads; ++t)
    helpers.emplace_back([&, t]() {
      for (int c = 0; c < n_chunks; ++c) {
        while (ready.load(std::memory_order_acquire) <= c) std::this_thread::yield();
        share(c, t);
        finished[c & 1].fetch_add(1, std::memory_order_release);
      }
    });
  cudaError_t err = cudaSuccess;
  auto issue = [&](int c) {
    const size_t off = (size_t)c * kChunk, len = std::min(kChunk, bytes - off);
    cudaError_t r = cudaMemcpyAsync(e->bounce[c & 1], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, e->stream);
    if (r == cudaSuccess) r = cudaEventRecord(e->bounce_ev[c & 1], e->stream);
    if (r != cudaSuccess && err == cudaSuccess) err = r;
  };
  issue(0);
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) {  // bounce[(c + 1) & 1] held chunk c - 1: every helper must be done with it
      if (c >= 1) while (finished[(c + 1) & 1].load(std::memory_order_acquire) < (kThreads - 1) * ((c + 1) / 2)) std::this_thread::yield();
      issue(c + 1);
    }
    const cudaError_t r = cudaEventSynchronize(e->bounce_ev[c & 1]);
    if (r != cudaSuccess && err == cudaSuccess) err = r;
    ready.store(c + 1, std::memory_order_release);
    share(c, 0);
  }
  for (auto& h : helpers) h.join();
  if (err != cudaSuccess) return fail(TSDF_E_CUDA, "device -> host copy failed: %s", cudaGetErrorString(err));
  CU(wait_stream(e, e->stream));
  return TSDF_OK;
}

static int select_blocks(tsdf_engine* e, const float* bbox, int* n_sel) {
  GridBound g{};
  if (bbox) {  // BoundingCube::Scale<short>(1. / voxel_size), voxel_tsdf.cuh:21-26, voxel_tsdf.cu:429
    const float scale = (float)(1. / (double)e->voxel_size);
    auto cv = [&](float v) { float p = v * scale; if (p != p) return (short)0; p = std::max(-32768.f, std::min(32767.f, p)); return (short)(int)p; };
    g.xmin = cv(bbox[0]); g.xmax = cv(bbox[1]); g.ymin = cv(bbox[2]); g.ymax = cv(bbox[3]); g.zmin = cv(bbox[4]); g.zmax = cv(bbox[5]);
  }
  CU(cudaMemsetAsync(e->S.ctr + C_NSEL, 0, sizeof(int), e->stream));
  launch_select_blocks(e->S, bbox != nullptr, g, e->selected, e->num_sms, e->stream);
  CU(cudaMemcpyAsync(e->h_scalar, e->S.ctr + C_NSEL, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU(wait_stream(e, e->stream));
  *n_sel = e->h_scalar[0];
  return TSDF_OK;
}

static int gather_impl(tsdf_engine* e, const float* bbox, float* out, int64_t cap, int64_t* n_voxels) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  phase_begin(e, PH_GATHER, e->stream);
  int n_sel = 0;
  rc = select_blocks(e, bbox, &n_sel);
  if (rc) return rc;
  const size_t need = (size_t)n_sel * kBlockVolume;
  if (need > e->gather_cap) {  // grow-only result buffer instead of cudaMalloc/cudaFree per call (voxel_tsdf.cu:439-451)
    cudaFree(e->gather_out); e->gather_out = nullptr; e->gather_cap = 0;
    const size_t want = std::max(need, (size_t)1 << 20);
    CU(cudaMalloc(&e->gather_out, sizeof(float4) * want));
    e->gather_cap = want;
  }
  launch_download_voxels(e->S, e->selected, n_sel, e->voxel_size, e->gather_out, e->stream);
  phase_end(e, PH_GATHER, e->stream);
  CU(cudaGetLastError());
  e->gather_n = (int64_t)need;
  if (n_voxels) *n_voxels = e->gather_n;
  if (out && cap > 0) {
    const size_t m = (size_t)std::min<int64_t>(cap, e->gather_n);
    return copy_to_host(e, out, e->gather_out, sizeof(float4) * m);
  }
  CU(wait_stream(e, e->stream));
  return TSDF_OK;
}
int tsdf_gather_valid(tsdf_handle e, float* out, int64_t cap, int64_t* n) { return gather_impl(e, nullptr, out, cap, n); }
int tsdf_gather_in_bound(tsdf_handle e, const float bbox[6], float* out, int64_t cap, int64_t* n) {
  if (!bbox) return fail(TSDF_E_INVALID, "null bbox");
  return gather_impl(e, bbox, out, cap, n);
}
int tsdf_gather_fetch(tsdf_handle e, float* out, int64_t cap) {
  if (!e || !out) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  const size_t m = (size_t)std::min<int64_t>(cap, e->gather_n);
  return copy_to_host(e, out, e->gather_out, sizeof(float4) * m);
}
int tsdf_gather_device_result(tsdf_handle e, const void** d_out, int64_t* n) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if (d_out) *d_out = e->gather_out;
  if (n) *n = e->gather_n;
  return TSDF_OK;
}

// ---- mesh extraction (kernels_mesh.cu) --------------------------------------------------
int tsdf_extract_mesh(tsdf_handle e, const float* bbox, float* out, int64_t cap, int64_t* n_triangles) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  phase_begin(e, PH_GATHER, e->stream);
  int n_sel = 0;
  rc = select_blocks(e, bbox, &n_sel);
  if (rc) return rc;
  if (!e->mesh_counter) CU(cudaMalloc(&e->mesh_counter, sizeof(unsigned long long)));
  // pass 1: count (the same kernel without the writes), pass 2: emit into the grow-only result buffer
  CU(cudaMemsetAsync(e->mesh_counter, 0, sizeof(unsigned long long), e->stream));
  launch_mesh_count(e->S, e->selected, n_sel, e->voxel_size, e->mesh_counter, e->stream);
  static_assert(sizeof(unsigned long long) <= sizeof(int) * 2, "h_scalar scratch");
  CU(cudaMemcpyAsync(e->h_scalar, e->mesh_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CU(wait_stream(e, e->stream));
  unsigned long long total = 0;
  memcpy(&total, e->h_scalar, sizeof(total));
  if ((size_t)total > e->mesh_cap) {
    cudaFree(e->mesh_out); e->mesh_out = nullptr; e->mesh_cap = 0;
    const size_t want = std::max((size_t)total, (size_t)1 << 18);
    CU(cudaMalloc(&e->mesh_out, sizeof(float) * 9 * want));
    e->mesh_cap = want;
  }
  CU(cudaMemsetAsync(e->mesh_counter, 0, sizeof(unsigned long long), e->stream));
  launch_mesh_emit(e->S, e->selected, n_sel, e->voxel_size, e->mesh_out, (long long)e->mesh_cap, e->mesh_counter, e->stream);
  phase_end(e, PH_GATHER, e->stream);
  CU(cudaGetLastError());
  e->mesh_n = (int64_t)total;
  if (n_triangles) *n_triangles = e->mesh_n;
  if (out && cap > 0) {
    const size_t m = (size_t)std::min<int64_t>(cap, e->mesh_n);
    return copy_to_host(e, out, e->mesh_out, sizeof(float) * 9 * m);
  }
  CU(wait_stream(e, e->stream));
  return TSDF_OK;
}
int tsdf_mesh_fetch(tsdf_handle e, float* out, int64_t cap) {
  if (!e || !out) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  const size_t m = (size_t)std::min<int64_t>(cap, e->mesh_n);
  return copy_to_host(e, out, e->mesh_out, sizeof(float) * 9 * m);
}
int tsdf_mesh_device_result(tsdf_handle e, const void** d_out, int64_t* n) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  if (d_out) *d_out = e->mesh_out;
  if (n) *n = e->mesh_n;
  return TSDF_OK;
}

int tsdf_num_active_blocks(tsdf_handle e, int* n) {
  if (!e || !n) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  CU(cudaMemcpyAsync(e->h_scalar, e->S.ctr + C_FREE, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU(wait_stream(e, e->stream));
  e->n_active = e->cfg.pool_blocks - e->h_scalar[0];
  *n = e->n_active;
  return TSDF_OK;
}

int tsdf_get_skip_map_stats(tsdf_handle e, int64_t* attempts, int64_t* rebuilds) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  if (rc) return rc;
  CU(cudaMemcpyAsync(e->h_scalar, e->skip.hdr + 12, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU(wait_stream(e, e->stream));
  if (attempts) *attempts = e->skip_gen;
  if (rebuilds) *rebuilds = e->h_scalar[0];
  return TSDF_OK;
}

int tsdf_get_counters(tsdf_handle e, tsdf_counters* out) {
  if (!e || !out) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  *out = e->last;
  return rc;
}

int tsdf_block_owner(int16_t bx, int16_t by, int16_t bz, int shard_count, int shard_shift) {
  if (shard_count < 1 || shard_shift < 0 || shard_shift > 15) return -1;
  return (int)owner_of(pack_key(bx, by, bz), shard_count, shard_shift);
}

uint32_t tsdf_hash(int16_t bx, int16_t by, int16_t bz) { return hash_block(bx, by, bz) & ((1u << 21) - 1); }

// ---- unit-test / parity access -------------------------------------------------------------------
static int after_mutation(tsdf_engine* e) {
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(e->h_scalar, e->S.ctr, sizeof(int) * C_COUNT, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemsetAsync(e->S.ctr + C_PER_CALL, 0, sizeof(int) * (C_COUNT - C_PER_CALL), e->stream));  // frames expect them cleared
  CU(wait_stream(e, e->stream));
  e->n_active = e->cfg.pool_blocks - e->h_scalar[C_FREE];
  if (e->h_scalar[C_ERROR] & ERR_POOL) return fail(TSDF_E_POOL_EXHAUSTED, "voxel block pool exhausted (pool_blocks=%d)", e->cfg.pool_blocks);
  if (e->h_scalar[C_ERROR] & ERR_TABLE) return fail(TSDF_E_TABLE_FULL, "hash table full");
  return TSDF_OK;
}

int tsdf_allocate_blocks(tsdf_handle e, const int16_t* keys, int n) {
  if (!e || (!keys && n > 0) || n < 0) return fail(TSDF_E_INVALID, "bad argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  DevBuf<short> d; CU(d.alloc(3 * (size_t)n));
  if (n) CU(cudaMemcpyAsync(d.p, keys, sizeof(short) * 3 * n, cudaMemcpyHostToDevice, e->stream));
  next_serial(e);
  CU(cudaMemsetAsync(e->S.ctr + C_PER_CALL, 0, sizeof(int) * (C_COUNT - C_PER_CALL), e->stream));
  launch_allocate_list(e->S, d.p, n, e->stream);
  e->volume_epoch++;
  return after_mutation(e);
}
int tsdf_delete_blocks(tsdf_handle e, const int16_t* keys, int n) {
  if (!e || (!keys && n > 0) || n < 0) return fail(TSDF_E_INVALID, "bad argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  DevBuf<short> d; CU(d.alloc(3 * (size_t)n));
  if (n) CU(cudaMemcpyAsync(d.p, keys, sizeof(short) * 3 * n, cudaMemcpyHostToDevice, e->stream));
  next_serial(e);
  CU(cudaMemsetAsync(e->S.ctr + C_PER_CALL, 0, sizeof(int) * (C_COUNT - C_PER_CALL), e->stream));
  launch_delete_list(e->S, d.p, n, e->stream);
  e->volume_epoch++;
  return after_mutation(e);
}
int tsdf_retrieve_voxels(tsdf_handle e, const int16_t* pts, int n, float* tsdf_out, uint8_t* rgbw, float* prob, int32_t* found) {
  if (!e || !pts || n < 0) return fail(TSDF_E_INVALID, "bad argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  DevBuf<short> dp; DevBuf<float> dt, dpr; DevBuf<unsigned> dc; DevBuf<int> df;
  CU(dp.alloc(3 * (size_t)n)); CU(dt.alloc(n)); CU(dpr.alloc(n)); CU(dc.alloc(n)); CU(df.alloc(n));
  if (n) CU(cudaMemcpyAsync(dp.p, pts, sizeof(short) * 3 * n, cudaMemcpyHostToDevice, e->stream));
  launch_retrieve_list(e->S, dp.p, n, dt.p, dc.p, dpr.p, df.p, e->stream);
  CU(cudaGetLastError());
  if (n && tsdf_out) CU(cudaMemcpyAsync(tsdf_out, dt.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  if (n && rgbw) CU(cudaMemcpyAsync(rgbw, dc.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  if (n && prob) CU(cudaMemcpyAsync(prob, dpr.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  if (n && found) CU(cudaMemcpyAsync(found, df.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  CU(wait_stream(e, e->stream));
  return TSDF_OK;
}
int tsdf_assign_voxels(tsdf_handle e, const int16_t* pts, int n, const float* tsdf_in, const uint8_t* rgbw, const float* prob) {
  if (!e || !pts || n < 0) return fail(TSDF_E_INVALID, "bad argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  DevBuf<short> dp; DevBuf<float> dt, dpr; DevBuf<unsigned> dc;
  CU(dp.alloc(3 * (size_t)n)); CU(dt.alloc(n)); CU(dpr.alloc(n)); CU(dc.alloc(n));
  if (n) CU(cudaMemcpyAsync(dp.p, pts, sizeof(short) * 3 * n, cudaMemcpyHostToDevice, e->stream));
  if (n && tsdf_in) CU(cudaMemcpyAsync(dt.p, tsdf_in, 4 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  if (n && rgbw) CU(cudaMemcpyAsync(dc.p, rgbw, 4 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  if (n && prob) CU(cudaMemcpyAsync(dpr.p, prob, 4 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  launch_assign_list(e->S, dp.p, n, tsdf_in ? dt.p : nullptr, rgbw ? dc.p : nullptr, prob ? dpr.p : nullptr, e->stream);
  CU(cudaGetLastError());
  CU(wait_stream(e, e->stream));
  return TSDF_OK;
}

int tsdf_export_blocks(tsdf_handle e, int16_t* keys, float* tsdf_out, uint8_t* rgbw, float* prob, int cap_blocks, int* n_blocks) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  int n = 0;
  rc = select_blocks(e, nullptr, &n);
  if (rc) return rc;
  if (n_blocks) *n_blocks = n;
  const int m = std::min(n, cap_blocks);
  if (m <= 0 || (!keys && !tsdf_out && !rgbw && !prob)) return TSDF_OK;
  // download the first n blocks' keys to find the canonical order, then export in chunks
  DevBuf<short> dk; CU(dk.alloc(3 * (size_t)n));
  launch_export_blocks(e->S, e->selected, n, dk.p, nullptr, nullptr, nullptr, e->stream);
  std::vector<short> hk(3 * (size_t)n);
  CU(cudaMemcpyAsync(hk.data(), dk.p, sizeof(short) * hk.size(), cudaMemcpyDeviceToHost, e->stream));
  std::vector<int> hsel(n);
  CU(cudaMemcpyAsync(hsel.data(), e->selected, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  CU(wait_stream(e, e->stream));
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    for (int c = 2; c >= 0; --c) if (hk[3 * a + c] != hk[3 * b + c]) return hk[3 * a + c] < hk[3 * b + c];
    return false;
  });
  std::vector<int> sorted_sel(n);
  for (int i = 0; i < n; ++i) sorted_sel[i] = hsel[order[i]];
  CU(cudaMemcpyAsync(e->selected, sorted_sel.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  const int chunk = 4096;
  DevBuf<short> ck; DevBuf<float> ct, cp; DevBuf<unsigned> cc;
  CU(ck.alloc(3 * (size_t)chunk)); CU(ct.alloc((size_t)chunk * 512)); CU(cp.alloc((size_t)chunk * 512)); CU(cc.alloc((size_t)chunk * 512));
  for (int b0 = 0; b0 < m; b0 += chunk) {
    const int nb = std::min(chunk, m - b0);
    launch_export_blocks(e->S, e->selected + b0, nb, ck.p, ct.p, cc.p, cp.p, e->stream);
    CU(cudaGetLastError());
    if (keys) CU(cudaMemcpyAsync(keys + 3 * (size_t)b0, ck.p, sizeof(short) * 3 * (size_t)nb, cudaMemcpyDeviceToHost, e->stream));
    if (tsdf_out) CU(cudaMemcpyAsync(tsdf_out + (size_t)b0 * 512, ct.p, 4 * (size_t)nb * 512, cudaMemcpyDeviceToHost, e->stream));
    if (rgbw) CU(cudaMemcpyAsync(rgbw + (size_t)b0 * 2048, cc.p, 4 * (size_t)nb * 512, cudaMemcpyDeviceToHost, e->stream));
    if (prob) CU(cudaMemcpyAsync(prob + (size_t)b0 * 512, cp.p, 4 * (size_t)nb * 512, cudaMemcpyDeviceToHost, e->stream));
    CU(wait_stream(e, e->stream));
  }
  return TSDF_OK;
}

int tsdf_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaMallocHost(ptr, bytes));
  return TSDF_OK;
}
int tsdf_host_free(void* ptr) { if (ptr) CU(cudaFreeHost(ptr)); return TSDF_OK; }

int tsdf_set_profiling(tsdf_handle e, int enabled) {
  if (!e) return fail(TSDF_E_INVALID, "null engine handle");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  CU(wait_stream(e, e->copy_stream));
  phase_collect(e);
  e->profiling = enabled != 0;
  for (int p = 0; p < PH_COUNT; ++p) { e->phase_total_ms[p] = 0.0; e->phase_count[p] = 0; }
  e->totals = tsdf_counters{}; e->total_frames = 0;
  return TSDF_OK;
}
int tsdf_get_phase_ms(tsdf_handle e, float out_ms[8], int64_t out_count[8]) {
  if (!e || !out_ms) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e); if (rc) return rc;
  CU(wait_stream(e, e->copy_stream));
  phase_collect(e);
  for (int p = 0; p < 8; ++p) { out_ms[p] = 0.f; if (out_count) out_count[p] = 0; }
  for (int p = 0; p < PH_COUNT; ++p) { out_ms[p] = (float)e->phase_total_ms[p]; if (out_count) out_count[p] = e->phase_count[p]; }
  return TSDF_OK;
}
int tsdf_get_totals(tsdf_handle e, tsdf_counters* sums, int64_t* n_frames) {
  if (!e || !sums) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(e->device));
  int rc = drain(e);
  *sums = e->totals;
  if (n_frames) *n_frames = e->total_frames;
  return rc;
}

}  // extern "C"
