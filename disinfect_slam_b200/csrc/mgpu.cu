// mgpu.cu -- libtsdf_b200_mgpu.so: the multi-GPU data plane of the B200 TSDF engine (C ABI in
// include/tsdf_b200_mgpu.h).  One volume, voxel blocks sharded over the GPUs of a node by block-coordinate
// ownership; NCCL is linked directly, every per-frame step is enqueued on CUDA streams and nothing waits on the host.
//
// The reference is single-GPU (utils/tsdf/voxel_tsdf.cuh:103-104: one TSDFGrid, two streams); what is distributed
// here is TSDFGrid::Integrate (voxel_tsdf.cu:347-375), TSDFGrid::RayCast (:490-506) and GatherValid / GatherVoxels
// (:399-454).
//
//   stream `cs` (communication)   frame k+1: [root: H2D of the planes] -> grouped ncclBroadcast (comm_frame)
//   stream `es` (engine)          frame k: allocate / select / integrate of the owned blocks
//                                 view k:  peer barrier -> skip map over all shards + march of this rank's rows with
//                                          peer loads over NVLink, every finished ray stored straight into the image
//                                          buffers of ALL ranks (posted NVLink stores) -> peer barrier
// The peer barrier is one 32-thread kernel: a release-store of the epoch into a flag word of every peer and an
// acquire-spin on the own flag words -- no NCCL kernel on the engine stream at all (TSDF_MGPU_EXCHANGE=nccl selects
// the conventional form instead: 4-byte ncclAllReduce, local image, grouped in-place ncclAllGather on comm_sync).
// Two communicators, because the broadcast of the next frame runs concurrently with collectives of the current view;
// every rank issues its calls in the same order.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>
#include <unistd.h>

#include "../../include/tsdf_b200_mgpu.h"

namespace {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
#define CU(call)                                                                                               \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess) return fail(TSDF_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                                       __FILE__, __LINE__);                                                    \
  } while (0)
#define NC(call)                                                                                               \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) return fail(TSDF_E_CUDA, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r_),    \
                                       __FILE__, __LINE__);                                                    \
  } while (0)
#define TS(call)                                                                                               \
  do {                                                                                                         \
    int r_ = (call);                                                                                           \
    if (r_ != TSDF_OK) return fail(r_, "%s: %s", #call, tsdf_last_error());                                    \
  } while (0)

constexpr int kMaxRanks = 8;
constexpr size_t kFlagBytes = 256;
constexpr size_t kRowSlack = (size_t)kMaxRanks * 8192;  // pixels: the in-place all-gather pads the image to world * rows_per rows
constexpr int kStage = 3;  // frame staging sets: broadcast of frame k+1 while frame k integrates and k-1 retires
enum { T_BCAST = 0, T_BARRIER, T_ALLGATHER, T_COMPOSITE, T_RAYCAST, T_GATHER, T_XBARRIER, T_COUNT };

struct Stage {
  unsigned char* buf = nullptr;     // [depth f32 | ht f32 | lt f32 | rgb u8 x 3] x max_image_pixels
  cudaEvent_t received = nullptr;   // broadcast into buf finished (recorded on cs)
  cudaEvent_t consumed = nullptr;   // the frame kernels that read buf finished (recorded on es)
  bool used = false;
};
}  // namespace

struct tsdf_mgpu {
  int rank = 0, world = 1, device = 0;
  size_t max_px = 0;
  tsdf_handle eng = nullptr;
  cudaStream_t es = nullptr, cs = nullptr;
  ncclComm_t comm_frame = nullptr, comm_sync = nullptr;
  Stage stage[kStage];
  int cur = 0;
  int* d_flag = nullptr;
  // exchange buffer of this rank, mapped by every peer: [rgba | normal | hit depth] images of (max_px + slack) pixels
  // each, then the barrier flag words (one per peer)
  unsigned char* xbuf = nullptr; size_t img_stride = 0;
  unsigned char* xpeer[kMaxRanks] = {};     // every rank's exchange buffer as seen from this GPU ([rank] = xbuf)
  void* xopened[kMaxRanks] = {};            // cudaIpcOpenMemHandle results (closed at destroy)
  unsigned char* img[3] = {nullptr, nullptr, nullptr};
  int epoch = 0;                            // barrier count (identical on all ranks: same call sequence)
  int* d_err = nullptr;                     // set by a barrier that timed out
  bool fused = true;
  int views_since_integrate = 0;            // all ranks make the same calls: > 0 means no shard has integrated since the last view
  // TSDF mirror of this rank: every shard's TSDF planes, kept up to date by the owners' integrate kernels (tsdf_mirror_attach)
  float* mirror = nullptr;
  void* mopened[kMaxRanks] = {};
  int mirror_mode = 0;                      // 0 none (foreign voxels over NVLink sample by sample), 1 pushed mirrors, 2 pulled cache
  bool row_bands = false;                   // one contiguous band of rows per rank instead of 8-row tiles dealt round-robin
  // candidate exchange (tsdf_alloc_exchange_attach): this rank's inbox, every rank's as seen from here, keys per sender
  unsigned char* inbox = nullptr; unsigned char* ipeer[kMaxRanks] = {}; void* iopened[kMaxRanks] = {}; int xa_cap = 0;
  // TSDF_MGPU_MODE=replicas: every rank keeps the WHOLE volume (an unsharded engine) and integrates every frame; the views
  // are dealt out round-robin, view k rendered by rank k % world straight into slot [k % world] of rank 0's image memory
  bool replicas = false; long long view_no = 0; int n_slots = 1; size_t flag_off = 0;
  bool pending_barrier = false;             // run_sequence: the barrier after a view is supplied by the next frame's exchange barrier
  int last_w = 0, last_h = 0;
  unsigned long long* keys = nullptr; size_t keys_cap = 0;
  long long* d_sizes = nullptr; long long* h_sizes = nullptr;  // [world] gather sizes / counter sums
  float* gather_all = nullptr; size_t gather_cap = 0;          // root: records of every shard (float4 per voxel)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs[T_COUNT];
  double total_ms[T_COUNT] = {};
  long long total_n[T_COUNT] = {};
};

namespace {
cudaEvent_t ev_get(tsdf_mgpu* m) {
  if (!m->ev_pool.empty()) { cudaEvent_t v = m->ev_pool.back(); m->ev_pool.pop_back(); return v; }
  cudaEvent_t v = nullptr; cudaEventCreate(&v); return v;
}
struct Timed {  // brackets a stretch of one stream with two events while profiling is on
  tsdf_mgpu* m; int what; cudaStream_t st; cudaEvent_t a = nullptr;
  Timed(tsdf_mgpu* m_, int what_, cudaStream_t st_) : m(m_), what(what_), st(st_) {
    if (m->profiling) { a = ev_get(m); cudaEventRecord(a, st); }
  }
  ~Timed() {
    if (a) { cudaEvent_t b = ev_get(m); cudaEventRecord(b, st); m->ev_pairs[what].emplace_back(a, b); }
  }
};
void collect(tsdf_mgpu* m) {
  for (int t = 0; t < T_COUNT; ++t) {
    for (auto& pr : m->ev_pairs[t]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { m->total_ms[t] += ms; m->total_n[t]++; }
      else cudaGetLastError();
      m->ev_pool.push_back(pr.first); m->ev_pool.push_back(pr.second);
    }
    m->ev_pairs[t].clear();
  }
}
int rows_per_rank(const tsdf_mgpu* m, int h) { return (h + m->world - 1) / m->world; }

// Barrier over the GPUs of the volume without a collective library: lane r publishes this rank's arrival at `epoch` in
// peer r's flag word [rank] (release store at system scope, after a system fence: everything earlier kernels of this
// stream wrote -- voxels, image rows stored into peer buffers -- is visible to whoever acquires the flag), then waits
// until peer r's arrival shows up in the own flag word [r].  Epochs only grow, a fast peer may already be one ahead.
struct PeerFlags { int* flags[kMaxRanks]; };
// Optional mail run of the candidate exchange: before signalling, lane r tells peer r how many keys this rank has put into
// its inbox for the frame of this parity (count word [parity * 8 + rank] of r's inbox header).
struct PeerCounts { const int* cursor; int* header[kMaxRanks]; int slot, cap; };
__global__ void peer_barrier_kernel(PeerFlags peers, PeerCounts mail, int rank, int world, int epoch, long long timeout_cycles, int* err) {
  const int r = threadIdx.x;
  if (r >= world) return;
  if (mail.cursor) mail.header[r][mail.slot] = min(mail.cursor[r], mail.cap);
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peers.flags[r] + rank), "r"(epoch) : "memory");
  const int* mine = peers.flags[rank] + r;
  const long long t0 = clock64();
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v - epoch >= 0) break;
    if (clock64() - t0 > timeout_cycles) { *err = epoch; break; }  // a peer died: report instead of hanging the GPU
  }
}

struct XBlob { cudaIpcMemHandle_t h; cudaIpcMemHandle_t hm; cudaIpcMemHandle_t hi; long long pid; void* raw; void* raw_mirror; void* raw_inbox; int device; int pool_blocks; };
}  // namespace

extern "C" {

const char* tsdf_mgpu_last_error(void) { return g_err; }

int tsdf_mgpu_unique_id(void* id) {
  if (!id) return fail(TSDF_E_INVALID, "null id");
  static_assert(sizeof(ncclUniqueId) == TSDF_MGPU_ID_BYTES, "TSDF_MGPU_ID_BYTES must be sizeof(ncclUniqueId)");
  ncclUniqueId u;
  NC(ncclGetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return TSDF_OK;
}

int tsdf_mgpu_destroy(tsdf_mgpu_handle m) {
  if (!m) return TSDF_OK;
  cudaSetDevice(m->device);
  if (m->es) cudaStreamSynchronize(m->es);
  if (m->cs) cudaStreamSynchronize(m->cs);
  if (m->comm_frame) ncclCommDestroy(m->comm_frame);
  if (m->comm_sync) ncclCommDestroy(m->comm_sync);
  for (Stage& s : m->stage) {
    cudaFree(s.buf);
    if (s.received) cudaEventDestroy(s.received);
    if (s.consumed) cudaEventDestroy(s.consumed);
  }
  for (int r = 0; r < kMaxRanks; ++r) if (m->xopened[r]) cudaIpcCloseMemHandle(m->xopened[r]);
  for (int r = 0; r < kMaxRanks; ++r) if (m->mopened[r]) cudaIpcCloseMemHandle(m->mopened[r]);
  for (int r = 0; r < kMaxRanks; ++r) if (m->iopened[r]) cudaIpcCloseMemHandle(m->iopened[r]);
  if (m->eng) tsdf_alloc_exchange_attach(m->eng, 0, nullptr, 0, nullptr, nullptr);  // (drains the engine)
  cudaFree(m->xbuf); cudaFree(m->d_err); cudaFree(m->mirror); cudaFree(m->inbox);
  cudaFree(m->d_flag); cudaFree(m->keys); cudaFree(m->d_sizes); cudaFree(m->gather_all);
  if (m->h_sizes) cudaFreeHost(m->h_sizes);
  collect(m);
  for (cudaEvent_t v : m->ev_pool) cudaEventDestroy(v);
  if (m->cs) cudaStreamDestroy(m->cs);
  if (m->eng) tsdf_destroy(m->eng);
  delete m;
  return TSDF_OK;
}

static void xa_frame_hook(void* user, void* stream, const int* d_cursor, int parity);

int tsdf_mgpu_create(float voxel_size, float truncation, const tsdf_config* user_cfg, int rank, int world, const void* id,
                     tsdf_mgpu_handle* out) {
  if (!out || !id) return fail(TSDF_E_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(TSDF_E_INVALID, "bad rank %d / world %d (at most 8 shards)", rank, world);
  tsdf_config cfg;
  tsdf_default_config(&cfg);
  if (user_cfg) {
    if (user_cfg->struct_size != (int32_t)sizeof(tsdf_config)) return fail(TSDF_E_INVALID, "tsdf_config.struct_size mismatch");
    cfg = *user_cfg;
  }
  tsdf_mgpu* m = new tsdf_mgpu();
  m->rank = rank; m->world = world;
  {
    const char* mode = getenv("TSDF_MGPU_MODE");
    m->replicas = world > 1 && mode && !strcmp(mode, "replicas");
  }
  if (m->replicas) { cfg.shard_rank = 0; cfg.shard_count = 1; cfg.flags &= ~TSDF_FLAG_SHARD_SHIFT_MASK; }
  else { cfg.shard_rank = rank; cfg.shard_count = world; }
#define MX(call) do { int r_ = (call); if (r_ != TSDF_OK) { tsdf_mgpu_destroy(m); return r_; } } while (0)
  {
    int rc = tsdf_create(voxel_size, truncation, &cfg, &m->eng);
    if (rc != TSDF_OK) { const int r2 = fail(rc, "tsdf_create: %s", tsdf_last_error()); delete m; return r2; }
  }
  m->es = (cudaStream_t)tsdf_stream(m->eng);
  m->max_px = (size_t)cfg.max_image_pixels;
  auto body = [&]() -> int {
    CU(cudaGetDevice(&m->device));  // tsdf_create left the engine's device current
    CU(cudaStreamCreateWithFlags(&m->cs, cudaStreamNonBlocking));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    // two communicators from one id: the second id travels through the first communicator
    NC(ncclCommInitRank(&m->comm_frame, world, u, rank));
    CU(cudaMalloc(&m->d_flag, 256));
    CU(cudaMemsetAsync(m->d_flag, 0, 256, m->es));
    ncclUniqueId u2;
    if (rank == 0) NC(ncclGetUniqueId(&u2));
    unsigned char* d_id = reinterpret_cast<unsigned char*>(m->d_flag) + 128;
    if (rank == 0) CU(cudaMemcpyAsync(d_id, &u2, sizeof(u2), cudaMemcpyHostToDevice, m->es));
    NC(ncclBroadcast(d_id, d_id, sizeof(u2), ncclChar, 0, m->comm_frame, m->es));
    CU(cudaMemcpyAsync(&u2, d_id, sizeof(u2), cudaMemcpyDeviceToHost, m->es));
    CU(cudaStreamSynchronize(m->es));
    NC(ncclCommInitRank(&m->comm_sync, world, u2, rank));
    for (Stage& s : m->stage) {
      CU(cudaMalloc(&s.buf, 15 * m->max_px));
      CU(cudaEventCreateWithFlags(&s.received, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    }
    CU(cudaMalloc(&m->d_sizes, sizeof(long long) * 64));
    CU(cudaMallocHost(&m->h_sizes, sizeof(long long) * 64));
    // map every shard's table / pool: blobs all-gathered through NCCL
    if (!m->replicas) {
    unsigned char* d_blobs = nullptr;
    CU(cudaMalloc(&d_blobs, (size_t)TSDF_IPC_BLOB_BYTES * world));
    std::vector<unsigned char> blobs((size_t)TSDF_IPC_BLOB_BYTES * world);
    TS(tsdf_ipc_export(m->eng, blobs.data() + (size_t)rank * TSDF_IPC_BLOB_BYTES));
    CU(cudaMemcpyAsync(d_blobs + (size_t)rank * TSDF_IPC_BLOB_BYTES, blobs.data() + (size_t)rank * TSDF_IPC_BLOB_BYTES, TSDF_IPC_BLOB_BYTES,
                       cudaMemcpyHostToDevice, m->es));
    NC(ncclAllGather(d_blobs + (size_t)rank * TSDF_IPC_BLOB_BYTES, d_blobs, TSDF_IPC_BLOB_BYTES, ncclChar, m->comm_sync, m->es));
    CU(cudaMemcpyAsync(blobs.data(), d_blobs, blobs.size(), cudaMemcpyDeviceToHost, m->es));
    CU(cudaStreamSynchronize(m->es));
    cudaFree(d_blobs);
    TS(tsdf_ipc_attach(m->eng, world, blobs.data()));
    }
    // exchange buffer (images + barrier flags) of every rank, mapped the same way
    m->img_stride = 4 * (m->max_px + kRowSlack);
    m->n_slots = m->replicas ? world : 1;   // replicas: one image set per rendering rank (only rank 0's copy is used)
    m->flag_off = 3 * m->img_stride * (size_t)m->n_slots;  // the same layout on every rank
    CU(cudaMalloc(&m->xbuf, m->flag_off + kFlagBytes));
    CU(cudaMemsetAsync(m->xbuf, 0, m->flag_off + kFlagBytes, m->es));
    CU(cudaMalloc(&m->d_err, sizeof(int)));
    CU(cudaMemsetAsync(m->d_err, 0, sizeof(int), m->es));
    for (int i = 0; i < 3; ++i) m->img[i] = m->xbuf + (size_t)i * m->img_stride;
    {
      const char* ex = getenv("TSDF_MGPU_EXCHANGE");
      m->fused = m->replicas || !(ex && !strcmp(ex, "nccl"));
      // How the march gets at foreign TSDF samples.  Default (push): every rank holds a mirror of every shard's TSDF
      // planes, kept current by the owners' integrate kernels with posted NVLink stores -- the fastest form measured at
      // 2 and at 8 GPUs.  TSDF_MGPU_MIRROR=pull: no mirrors; before the march every rank fetches the TSDF planes of the
      // foreign blocks the view can meet (about a room's worth, a few MB) into a local cache -- no remote stores in the
      // integrate kernel, but two more kernels per view and a stamp check per sample; =0: neither, every foreign sample is a load over
      // NVLink (memory per rank = its shard only).  The NCCL-exchange variant keeps the plain form.
      const char* mi = getenv("TSDF_MGPU_MIRROR");
      m->mirror_mode = world < 2 || m->replicas ? 0 : (mi && !strcmp(mi, "0")) ? 0 : (mi && !strcmp(mi, "pull")) ? 2 : (mi && !strcmp(mi, "push")) ? 1 : (m->fused ? 1 : 0);
      const bool want_mirror = m->mirror_mode == 1;
      // Which rows a rank renders: 8-row tiles dealt round-robin (default: every rank gets the same mix of cheap and
      // expensive rows) or TSDF_MGPU_TILES=band, one contiguous band per rank (fewer foreign blocks to fetch, but the
      // ranks finish at different times)
      const char* ti = getenv("TSDF_MGPU_TILES");
      m->row_bands = ti && !strcmp(ti, "band");
      // TSDF_MGPU_ALLOC=exchange: allocation pass sharded by image tiles, candidate keys mailed to their owners (needs the
      // peer barrier, i.e. the fused plane).  Default (owner): every rank walks all pixel rays and keeps its own blocks --
      // measured, the barrier the exchange needs costs more than the 20 us of ray walking it saves at 8 GPUs.
      const char* al = getenv("TSDF_MGPU_ALLOC");
      const bool want_exchange = world > 1 && m->fused && !m->replicas && al && !strcmp(al, "exchange");
      std::vector<XBlob> xb(world);
      XBlob mine{};
      CU(cudaIpcGetMemHandle(&mine.h, m->xbuf));
      mine.pid = (long long)getpid(); mine.raw = m->xbuf; mine.device = m->device; mine.pool_blocks = cfg.pool_blocks;
      if (want_mirror) {  // all shards use the same pool size here, so every rank sizes its mirror alike
        const size_t bytes = (size_t)world * cfg.pool_blocks * 2048;
        CU(cudaMalloc(&m->mirror, bytes));
        CU(cudaMemsetAsync(m->mirror, 0, bytes, m->es));
        CU(cudaIpcGetMemHandle(&mine.hm, m->mirror));
        mine.raw_mirror = m->mirror;
      }
      if (want_exchange) {
        // a sender walks 1/world of the pixels; four keys per pixel covers every frame whose rays stay within the inline
        // DDA length, with room to spare (typical frames mail ~0.1 key per pixel after the warp-level de-duplication)
        m->xa_cap = (int)std::max<size_t>(65536, 4 * ((m->max_px + world - 1) / world));
        const size_t bytes = tsdf_alloc_exchange_bytes(world, m->xa_cap);
        CU(cudaMalloc(&m->inbox, bytes));
        CU(cudaMemsetAsync(m->inbox, 0, bytes, m->es));
        CU(cudaIpcGetMemHandle(&mine.hi, m->inbox));
        mine.raw_inbox = m->inbox;
      }
      XBlob* d_xb = nullptr;
      CU(cudaMalloc(&d_xb, sizeof(XBlob) * world));
      CU(cudaMemcpyAsync(d_xb + rank, &mine, sizeof(XBlob), cudaMemcpyHostToDevice, m->es));
      NC(ncclAllGather(d_xb + rank, d_xb, sizeof(XBlob), ncclChar, m->comm_sync, m->es));
      CU(cudaMemcpyAsync(xb.data(), d_xb, sizeof(XBlob) * world, cudaMemcpyDeviceToHost, m->es));
      CU(cudaStreamSynchronize(m->es));
      cudaFree(d_xb);
      for (int r = 0; r < world; ++r) {
        if (r == rank) { m->xpeer[r] = m->xbuf; continue; }
        if (xb[r].pid == (long long)getpid()) {  // a rank of this process (thread per GPU): plain peer access
          if (xb[r].device != m->device) {
            const cudaError_t pe = cudaDeviceEnablePeerAccess(xb[r].device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(TSDF_E_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", xb[r].device, cudaGetErrorString(pe));
            cudaGetLastError();
          }
          m->xpeer[r] = (unsigned char*)xb[r].raw;
        } else {
          void* p = nullptr;
          CU(cudaIpcOpenMemHandle(&p, xb[r].h, cudaIpcMemLazyEnablePeerAccess));
          m->xopened[r] = p; m->xpeer[r] = (unsigned char*)p;
        }
      }
      if (want_mirror) {
        void* mp[kMaxRanks] = {};
        for (int r = 0; r < world; ++r) {
          if (xb[r].pool_blocks != cfg.pool_blocks) return fail(TSDF_E_INVALID, "TSDF mirrors need the same pool_blocks on every rank (%d vs %d)", xb[r].pool_blocks, cfg.pool_blocks);
          if (r == rank) mp[r] = m->mirror;
          else if (xb[r].pid == (long long)getpid()) mp[r] = xb[r].raw_mirror;  // peer access was enabled above
          else { CU(cudaIpcOpenMemHandle(&mp[r], xb[r].hm, cudaIpcMemLazyEnablePeerAccess)); m->mopened[r] = mp[r]; }
        }
        TS(tsdf_mirror_attach(m->eng, world, mp, cfg.pool_blocks));
      }
      if (m->mirror_mode == 2) {
        for (int r = 0; r < world; ++r)
          if (xb[r].pool_blocks != cfg.pool_blocks) return fail(TSDF_E_INVALID, "the TSDF cache needs the same pool_blocks on every rank (%d vs %d)", xb[r].pool_blocks, cfg.pool_blocks);
        TS(tsdf_shared_cache_attach(m->eng, cfg.pool_blocks, 3));
      }
      if (want_exchange) {
        void* ip[kMaxRanks] = {};
        for (int r = 0; r < world; ++r) {
          if (r == rank) ip[r] = m->inbox;
          else if (xb[r].pid == (long long)getpid()) ip[r] = xb[r].raw_inbox;
          else { CU(cudaIpcOpenMemHandle(&ip[r], xb[r].hi, cudaIpcMemLazyEnablePeerAccess)); m->iopened[r] = ip[r]; }
          m->ipeer[r] = (unsigned char*)ip[r];
        }
        TS(tsdf_alloc_exchange_attach(m->eng, world, ip, m->xa_cap, xa_frame_hook, m));
      }
      // nobody may signal into a buffer that its owner is still clearing
      NC(ncclAllReduce(m->d_flag, m->d_flag, 1, ncclInt, ncclSum, m->comm_sync, m->es));
      CU(cudaStreamSynchronize(m->es));
    }
    return TSDF_OK;
  };
  MX(body());
#undef MX
  *out = m;
  return TSDF_OK;
}

tsdf_handle tsdf_mgpu_engine(tsdf_mgpu_handle m) { return m ? m->eng : nullptr; }

int tsdf_mgpu_integrate(tsdf_mgpu_handle m, int root, int on_device, const void* rgb, const void* depth, const void* ht,
                        const void* lt, int w, int h, float max_depth, const float K[4], const float q[4], const float t[3]) {
  if (!m || !K || !q || !t) return fail(TSDF_E_INVALID, "null argument");
  if (root < 0 || root >= m->world) return fail(TSDF_E_INVALID, "bad root %d", root);
  if (w <= 0 || h <= 0 || (size_t)w * h > m->max_px) return fail(TSDF_E_INVALID, "image %dx%d exceeds max_image_pixels=%zu", w, h, m->max_px);
  const bool is_root = m->rank == root;
  if (is_root && (!rgb || !depth || !ht || !lt)) return fail(TSDF_E_INVALID, "root needs the four planes");
  CU(cudaSetDevice(m->device));
  Stage& s = m->stage[m->cur];
  const size_t n = (size_t)w * h;
  unsigned char* const d_depth = s.buf, *const d_ht = s.buf + 4 * n, *const d_lt = s.buf + 8 * n, *const d_rgb = s.buf + 12 * n;
  if (s.used) CU(cudaStreamWaitEvent(m->cs, s.consumed, 0));  // the frame kernels three frames ago are done with this set
  const void* src[4] = {d_depth, d_ht, d_lt, d_rgb};
  if (is_root) {
    if (on_device) {  // broadcast straight out of the caller's planes: no staging copy on the root
      src[0] = depth; src[1] = ht; src[2] = lt; src[3] = rgb;
    } else {
      CU(cudaMemcpyAsync(d_depth, depth, 4 * n, cudaMemcpyHostToDevice, m->cs));
      CU(cudaMemcpyAsync(d_ht, ht, 4 * n, cudaMemcpyHostToDevice, m->cs));
      CU(cudaMemcpyAsync(d_lt, lt, 4 * n, cudaMemcpyHostToDevice, m->cs));
      CU(cudaMemcpyAsync(d_rgb, rgb, 3 * n, cudaMemcpyHostToDevice, m->cs));
    }
  }
  if (m->world > 1) {
    Timed tm(m, T_BCAST, m->cs);
    NC(ncclGroupStart());
    NC(ncclBroadcast(src[0], d_depth, 4 * n, ncclChar, root, m->comm_frame, m->cs));
    NC(ncclBroadcast(src[1], d_ht, 4 * n, ncclChar, root, m->comm_frame, m->cs));
    NC(ncclBroadcast(src[2], d_lt, 4 * n, ncclChar, root, m->comm_frame, m->cs));
    NC(ncclBroadcast(src[3], d_rgb, 3 * n, ncclChar, root, m->comm_frame, m->cs));
    NC(ncclGroupEnd());
  } else if (is_root && on_device) {
    CU(cudaMemcpyAsync(d_depth, depth, 4 * n, cudaMemcpyDeviceToDevice, m->cs));
    CU(cudaMemcpyAsync(d_ht, ht, 4 * n, cudaMemcpyDeviceToDevice, m->cs));
    CU(cudaMemcpyAsync(d_lt, lt, 4 * n, cudaMemcpyDeviceToDevice, m->cs));
    CU(cudaMemcpyAsync(d_rgb, rgb, 3 * n, cudaMemcpyDeviceToDevice, m->cs));
  }
  CU(cudaEventRecord(s.received, m->cs));
  // the engine stream waits for the planes on the device (after_event), the host does not
  int rc = tsdf_integrate_device(m->eng, d_rgb, d_depth, d_ht, d_lt, w, h, max_depth, K, q, t, s.received);
  if (rc == TSDF_E_INVALID || rc == TSDF_E_CUDA || rc == TSDF_E_NO_DEVICE) return fail(rc, "tsdf_integrate_device: %s", tsdf_last_error());
  CU(cudaEventRecord(s.consumed, m->es));
  s.used = true;
  m->views_since_integrate = 0;
  m->cur = (m->cur + 1) % kStage;
  if (rc != TSDF_OK) return fail(rc, "%s", tsdf_last_error());  // exhaustion of an earlier frame; this frame was enqueued
  return TSDF_OK;
}

static int ensure_images(tsdf_mgpu* m, int w, int h) {
  const size_t need = (size_t)rows_per_rank(m, h) * m->world * w;
  if ((size_t)w * h > m->max_px || need > m->max_px + kRowSlack)
    return fail(TSDF_E_INVALID, "view %dx%d exceeds max_image_pixels=%zu of this volume", w, h, m->max_px);
  return TSDF_OK;
}

static int peer_barrier(tsdf_mgpu* m, const int* d_cursor = nullptr, int parity = 0) {
  PeerFlags pf{};
  for (int r = 0; r < m->world; ++r) pf.flags[r] = reinterpret_cast<int*>(m->xpeer[r] + m->flag_off);
  PeerCounts mail{};
  if (d_cursor) {
    mail.cursor = d_cursor; mail.slot = parity * 8 + m->rank; mail.cap = m->xa_cap;
    for (int r = 0; r < m->world; ++r) mail.header[r] = reinterpret_cast<int*>(m->ipeer[r]);
  }
  m->epoch++;
  m->pending_barrier = false;  // whatever an earlier view left open, this barrier closes
  peer_barrier_kernel<<<1, 32, 0, m->es>>>(pf, mail, m->rank, m->world, m->epoch, 4000000000ll, m->d_err);  // ~2 s
  CU(cudaGetLastError());
  return TSDF_OK;
}
// The engine calls this between a frame's allocate kernel and the owner's inserts (candidate exchange): publish how many
// keys went to every peer, and order the ranks.  It also stands in for the barrier a preceding view left open.
static void xa_frame_hook(void* user, void* /*stream: the engine's, == m->es*/, const int* d_cursor, int parity) {
  tsdf_mgpu* m = static_cast<tsdf_mgpu*>(user);
  Timed tm(m, T_XBARRIER, m->es);
  peer_barrier(m, d_cursor, parity);
}
static int flush_barrier(tsdf_mgpu* m) {
  if (!m->pending_barrier) return TSDF_OK;
  Timed tm(m, T_ALLGATHER, m->es);
  return peer_barrier(m);
}

static int raycast_impl(tsdf_mgpu_handle m, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                        const void** d_rgba, const void** d_normal, const void** d_depth, bool defer_barrier);
int tsdf_mgpu_raycast(tsdf_mgpu_handle m, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                      const void** d_rgba, const void** d_normal, const void** d_depth) {
  return raycast_impl(m, max_depth, w, h, K, q, t, d_rgba, d_normal, d_depth, false);
}

static int raycast_impl(tsdf_mgpu_handle m, float max_depth, int w, int h, const float K[4], const float q[4], const float t[3],
                        const void** d_rgba, const void** d_normal, const void** d_depth, bool defer_barrier) {
  if (!m || !K || !q || !t) return fail(TSDF_E_INVALID, "null argument");
  if (w <= 0 || h <= 0) return fail(TSDF_E_INVALID, "bad image size %dx%d", w, h);
  CU(cudaSetDevice(m->device));
  int rc = ensure_images(m, w, h);
  if (rc) return rc;
  const int rows = rows_per_rank(m, h), row0 = m->rank * rows;
  if (m->replicas) {
    // whole views dealt round-robin over the replicas: no rank waits for another one.  The renderer writes the view
    // straight into its slot of rank 0's image memory (posted stores over NVLink); views of one renderer are ordered by
    // its stream, views of different renderers use different slots.
    const int renderer = (int)(m->view_no % m->world);
    m->view_no++;
    if (m->rank == renderer) {
      Timed tm(m, T_RAYCAST, m->es);
      unsigned char* slot = m->xpeer[0] + 3 * m->img_stride * (size_t)renderer;
      TS(tsdf_raycast_device(m->eng, max_depth, w, h, K, q, t, slot, slot + m->img_stride, slot + 2 * m->img_stride, nullptr));
    }
    // a barrier makes every view issued so far complete on rank 0 (inside tsdf_mgpu_run_sequence: only after the last one)
    if (!defer_barrier) { Timed tm(m, T_ALLGATHER, m->es); int rc2 = peer_barrier(m); if (rc2) return rc2; }
    unsigned char* last = m->xbuf + 3 * m->img_stride * (size_t)renderer;
    for (int i = 0; i < 3; ++i) m->img[i] = last + (size_t)i * m->img_stride;
  } else if (m->world > 1 && m->fused) {
    // fused exchange: no collective library on the engine stream.  Barrier (every shard's Integrate has finished before
    // any rank reads its voxels) -> march with the finished rays stored into every rank's images -> barrier (all rows
    // have arrived everywhere, and every peer is done reading this rank's voxels before its next Integrate)
    { Timed tm(m, T_BARRIER, m->es); int rc2 = peer_barrier(m); if (rc2) return rc2; }
    {
      Timed tm(m, T_RAYCAST, m->es);
      void* o[3][kMaxRanks];
      for (int i = 0; i < 3; ++i) for (int r = 0; r < m->world; ++r) o[i][r] = m->xpeer[r] + (size_t)i * m->img_stride;
      const int unchanged = m->views_since_integrate > 0 ? 1 : 0;
      if (m->row_bands) {
        const int tiles = (h + 7) / 8, per = (tiles + m->world - 1) / m->world;
        TS(tsdf_raycast_shared_scatter(m->eng, max_depth, w, h, K, q, t, m->rank * per, 1, per, unchanged, m->world, o[0], o[1], o[2]));
      } else {
        // 8-row tiles dealt out round-robin: every rank renders the same mix of rows
        TS(tsdf_raycast_shared_scatter(m->eng, max_depth, w, h, K, q, t, m->rank, m->world, 0, unchanged, m->world, o[0], o[1], o[2]));
      }
      m->views_since_integrate++;
    }
    // barrier after the march: all rows have arrived everywhere, every peer is done with this rank's voxels.  Inside
    // tsdf_mgpu_run_sequence the next frame's exchange barrier does that job (nothing before it touches the volume or
    // the images), so the view costs one barrier less
    if (defer_barrier && m->inbox) m->pending_barrier = true;
    else { Timed tm(m, T_ALLGATHER, m->es); int rc2 = peer_barrier(m); if (rc2) return rc2; }
  } else {
    if (m->world > 1) {  // every shard's Integrate has finished before any rank reads its voxels
      Timed tm(m, T_BARRIER, m->es);
      NC(ncclAllReduce(m->d_flag, m->d_flag, 1, ncclInt, ncclSum, m->comm_sync, m->es));
    }
    {
      Timed tm(m, T_RAYCAST, m->es);
      if (row0 < h) TS(tsdf_raycast_shared(m->eng, max_depth, w, h, K, q, t, row0, rows, m->img[0], m->img[1], m->img[2]));
    }
    if (m->world > 1) {  // in place: this rank's rows already sit at chunk `rank` of each image
      Timed tm(m, T_ALLGATHER, m->es);
      const size_t chunk = (size_t)rows * w * 4;
      NC(ncclGroupStart());
      for (int i = 0; i < 3; ++i) NC(ncclAllGather(m->img[i] + (size_t)m->rank * chunk, m->img[i], chunk, ncclChar, m->comm_sync, m->es));
      NC(ncclGroupEnd());
    }
  }
  m->last_w = w; m->last_h = h;
  if (d_rgba) *d_rgba = m->img[0];
  if (d_normal) *d_normal = m->img[1];
  if (d_depth) *d_depth = m->img[2];
  return TSDF_OK;
}

int tsdf_mgpu_raycast_composite(tsdf_mgpu_handle m, float max_depth, int w, int h, const float K[4], const float q[4],
                                const float t[3], const void** d_keys) {
  if (!m || !K || !q || !t) return fail(TSDF_E_INVALID, "null argument");
  if (w <= 0 || h <= 0) return fail(TSDF_E_INVALID, "bad image size %dx%d", w, h);
  CU(cudaSetDevice(m->device));
  if (m->replicas) return fail(TSDF_E_INVALID, "min-compositing is for a sharded volume; replicas render whole views (tsdf_mgpu_raycast)");
  { int rcb = flush_barrier(m); if (rcb) return rcb; }
  const size_t need = 2 * (size_t)w * h;
  if (need > m->keys_cap) {
    CU(cudaStreamSynchronize(m->es));
    cudaFree(m->keys); m->keys = nullptr; m->keys_cap = 0;
    CU(cudaMalloc(&m->keys, sizeof(unsigned long long) * need));
    m->keys_cap = need;
  }
  TS(tsdf_raycast_device(m->eng, max_depth, w, h, K, q, t, nullptr, nullptr, nullptr, m->keys));
  if (m->world > 1) {
    Timed tm(m, T_COMPOSITE, m->es);
    NC(ncclAllReduce(m->keys, m->keys, need, ncclUint64, ncclMin, m->comm_sync, m->es));  // the nearest hit wins
  }
  if (d_keys) *d_keys = m->keys;
  return TSDF_OK;
}

int tsdf_mgpu_fetch_images(tsdf_mgpu_handle m, uint8_t* rgba, uint8_t* normal, float* depth) {
  if (!m) return fail(TSDF_E_INVALID, "null handle");
  if (m->last_w <= 0) return fail(TSDF_E_INVALID, "no raycast yet");
  CU(cudaSetDevice(m->device));
  const size_t bytes = (size_t)m->last_w * m->last_h * 4;
  if (rgba) CU(cudaMemcpyAsync(rgba, m->img[0], bytes, cudaMemcpyDeviceToHost, m->es));
  if (normal) CU(cudaMemcpyAsync(normal, m->img[1], bytes, cudaMemcpyDeviceToHost, m->es));
  if (depth) CU(cudaMemcpyAsync(depth, m->img[2], bytes, cudaMemcpyDeviceToHost, m->es));
  CU(cudaStreamSynchronize(m->es));
  return TSDF_OK;
}

int tsdf_mgpu_gather(tsdf_mgpu_handle m, int root, const float* bbox, float* out, int64_t cap, int64_t* n_voxels) {
  if (!m) return fail(TSDF_E_INVALID, "null handle");
  if (root < 0 || root >= m->world) return fail(TSDF_E_INVALID, "bad root %d", root);
  CU(cudaSetDevice(m->device));
  CU(cudaStreamSynchronize(m->cs));
  int64_t mine = 0;
  if (bbox) TS(tsdf_gather_in_bound(m->eng, bbox, nullptr, 0, &mine));
  else TS(tsdf_gather_valid(m->eng, nullptr, 0, &mine));
  const void* d_mine = nullptr;
  TS(tsdf_gather_device_result(m->eng, &d_mine, nullptr));
  // sizes first
  m->h_sizes[m->rank] = mine;
  if (m->world > 1 && !m->replicas) {
    CU(cudaMemcpyAsync(m->d_sizes + m->rank, m->h_sizes + m->rank, sizeof(long long), cudaMemcpyHostToDevice, m->es));
    NC(ncclAllGather(m->d_sizes + m->rank, m->d_sizes, 1, ncclInt64, m->comm_sync, m->es));
    CU(cudaMemcpyAsync(m->h_sizes, m->d_sizes, sizeof(long long) * m->world, cudaMemcpyDeviceToHost, m->es));
    CU(cudaStreamSynchronize(m->es));
  }
  if (m->replicas) {  // every rank holds the whole volume: the root answers from its own
    if (n_voxels) *n_voxels = mine;
    if (m->rank == root && out && cap > 0) TS(tsdf_gather_fetch(m->eng, out, cap));
    return TSDF_OK;
  }
  int64_t total = 0;
  for (int r = 0; r < m->world; ++r) total += m->h_sizes[r];
  if (n_voxels) *n_voxels = total;
  if (m->world == 1) {
    if (out && cap > 0) TS(tsdf_gather_fetch(m->eng, out, cap));
    return TSDF_OK;
  }
  if (m->rank == root && (size_t)total > m->gather_cap) {
    cudaFree(m->gather_all); m->gather_all = nullptr; m->gather_cap = 0;
    size_t want = (size_t)total > ((size_t)1 << 20) ? (size_t)total : ((size_t)1 << 20);
    CU(cudaMalloc(&m->gather_all, sizeof(float) * 4 * want));
    m->gather_cap = want;
  }
  {  // records stay on the devices: one grouped send / recv over NVLink, shard r's records at the prefix-sum offset
    Timed tm(m, T_GATHER, m->es);
    NC(ncclGroupStart());
    if (m->rank == root) {
      size_t off = 0;
      for (int r = 0; r < m->world; ++r) {
        const size_t cnt = (size_t)m->h_sizes[r] * 4;
        if (r == root) { if (cnt) CU(cudaMemcpyAsync(m->gather_all + off, d_mine, sizeof(float) * cnt, cudaMemcpyDeviceToDevice, m->es)); }
        else if (cnt) NC(ncclRecv(m->gather_all + off, cnt, ncclFloat, r, m->comm_sync, m->es));
        off += cnt;
      }
    } else if (mine) {
      NC(ncclSend(d_mine, (size_t)mine * 4, ncclFloat, root, m->comm_sync, m->es));
    }
    NC(ncclGroupEnd());
  }
  if (m->rank == root && out && cap > 0) {
    const size_t k = (size_t)(cap < total ? cap : total);
    if (k) CU(cudaMemcpyAsync(out, m->gather_all, sizeof(float) * 4 * k, cudaMemcpyDeviceToHost, m->es));
  }
  CU(cudaStreamSynchronize(m->es));
  return TSDF_OK;
}

int tsdf_mgpu_counters(tsdf_mgpu_handle m, tsdf_counters* last_sum, tsdf_counters* totals_sum, int64_t* n_active) {
  if (!m) return fail(TSDF_E_INVALID, "null handle");
  CU(cudaSetDevice(m->device));
  tsdf_counters last{}, tot{};
  int64_t frames = 0;
  int rc = tsdf_get_counters(m->eng, &last);
  if (rc == TSDF_E_CUDA || rc == TSDF_E_INVALID) return fail(rc, "tsdf_get_counters: %s", tsdf_last_error());
  rc = tsdf_get_totals(m->eng, &tot, &frames);
  if (rc == TSDF_E_CUDA || rc == TSDF_E_INVALID) return fail(rc, "tsdf_get_totals: %s", tsdf_last_error());
  int act = 0;
  TS(tsdf_num_active_blocks(m->eng, &act));
  static_assert(sizeof(tsdf_counters) == 8 * sizeof(long long), "tsdf_counters layout");
  long long* h = m->h_sizes;
  memcpy(h, &last, sizeof(last));
  memcpy(h + 8, &tot, sizeof(tot));
  h[16] = act;
  if (m->world > 1 && !m->replicas) {  // (a replica's counters already describe the whole volume)
    CU(cudaMemcpyAsync(m->d_sizes, h, sizeof(long long) * 17, cudaMemcpyHostToDevice, m->es));
    NC(ncclAllReduce(m->d_sizes, m->d_sizes, 17, ncclInt64, ncclSum, m->comm_sync, m->es));
    CU(cudaMemcpyAsync(h, m->d_sizes, sizeof(long long) * 17, cudaMemcpyDeviceToHost, m->es));
    CU(cudaStreamSynchronize(m->es));
  }
  if (last_sum) memcpy(last_sum, h, sizeof(last));
  if (totals_sum) memcpy(totals_sum, h + 8, sizeof(tot));
  if (n_active) *n_active = h[16];
  return TSDF_OK;
}

int tsdf_mgpu_run_sequence(tsdf_mgpu_handle m, int root, int on_device, const tsdf_mgpu_frame* frames, int n_frames, int first,
                           int count, int w, int h, float max_depth, const float K[4], int raycast_mode) {
  if (!m || !frames || n_frames <= 0 || first < 0 || count < 0) return fail(TSDF_E_INVALID, "bad argument");
  int deferred = TSDF_OK;
  for (int i = first; i < first + count; ++i) {
    const tsdf_mgpu_frame& f = frames[i % n_frames];
    int rc = tsdf_mgpu_integrate(m, root, on_device, f.rgb, f.depth, f.ht, f.lt, w, h, max_depth, K, f.q_xyzw, f.t_xyz);
    if (rc == TSDF_E_POOL_EXHAUSTED || rc == TSDF_E_TABLE_FULL) { deferred = rc; rc = TSDF_OK; }  // reported at the end; the stream goes on
    if (rc) return rc;
    if (raycast_mode == 1) rc = raycast_impl(m, max_depth, w, h, K, f.q_xyzw, f.t_xyz, nullptr, nullptr, nullptr, i + 1 < first + count);
    else if (raycast_mode == 2) rc = tsdf_mgpu_raycast_composite(m, max_depth, w, h, K, f.q_xyzw, f.t_xyz, nullptr);
    if (rc) return rc;
  }
  { int rcb = flush_barrier(m); if (rcb) return rcb; }
  return deferred;
}

int tsdf_mgpu_synchronize(tsdf_mgpu_handle m) {
  if (!m) return fail(TSDF_E_INVALID, "null handle");
  CU(cudaSetDevice(m->device));
  CU(cudaStreamSynchronize(m->cs));
  const int rc = tsdf_synchronize(m->eng);
  CU(cudaStreamSynchronize(m->es));
  int berr = 0;
  CU(cudaMemcpy(&berr, m->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (berr) return fail(TSDF_E_CUDA, "peer barrier %d timed out: a rank of the volume stopped making progress", berr);
  if (rc != TSDF_OK) return fail(rc, "%s", tsdf_last_error());
  return TSDF_OK;
}

int tsdf_mgpu_set_profiling(tsdf_mgpu_handle m, int enabled) {
  if (!m) return fail(TSDF_E_INVALID, "null handle");
  CU(cudaSetDevice(m->device));
  CU(cudaStreamSynchronize(m->cs));
  CU(cudaStreamSynchronize(m->es));
  collect(m);
  m->profiling = enabled != 0;
  for (int t = 0; t < T_COUNT; ++t) { m->total_ms[t] = 0.0; m->total_n[t] = 0; }
  return TSDF_OK;
}

int tsdf_mgpu_get_comm_ms(tsdf_mgpu_handle m, float out_ms[8], int64_t out_count[8]) {
  if (!m || !out_ms) return fail(TSDF_E_INVALID, "null argument");
  CU(cudaSetDevice(m->device));
  CU(cudaStreamSynchronize(m->cs));
  CU(cudaStreamSynchronize(m->es));
  collect(m);
  for (int t = 0; t < 8; ++t) { out_ms[t] = 0.f; if (out_count) out_count[t] = 0; }
  for (int t = 0; t < T_COUNT; ++t) { out_ms[t] = (float)m->total_ms[t]; if (out_count) out_count[t] = m->total_n[t]; }
  return TSDF_OK;
}

}  // extern "C"
