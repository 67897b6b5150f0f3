/*
 * tsdf_oracle.c -- CPU ORACLE for the voxel-hashed semantic TSDF path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (disinfect_slam_b200/csrc, libtsdf_b200.so) never links, imports or calls anything here.
 *
 * It is a scalar restatement (plain C, float32, no FMA contraction: build with
 * -ffp-contract=off) of the reference algorithm in /root/reference (paths below are
 * relative to that root):
 *   utils/tsdf/voxel_tsdf.cu:14-307   all kernels
 *   utils/tsdf/voxel_tsdf.cu:347-506  host order of operations
 *   utils/tsdf/voxel_hash.cu:31-35    Hash()
 *   utils/tsdf/voxel_hash.cuh:104-161 Retrieve + defaults for absent voxels
 *   utils/tsdf/voxel_mem.cuh:29-68    coordinate maths
 *   utils/tsdf/voxel_mem.cu:43-51     block init values
 *   utils/tsdf/voxel_types.cu:3-18    voxel defaults
 *   utils/cuda/camera.cuh:35-51       intrinsics
 *   utils/cuda/lie_group.cuh:25-40    SE3
 *   utils/tsdf/voxel_tsdf.cuh:21-26   BoundingCube::Scale
 *
 * Third-party arithmetic that is NOT under /root/reference: Eigen, pinned "EXACT 3.3.9"
 * (CMakeLists.txt:66).  Its published algorithms are restated here (scalar path, because
 * Eigen defines EIGEN_DONT_VECTORIZE under nvcc):
 *   q*v        QuaternionBase::_transformVector: uv = q.vec x v; uv += uv;
 *              r = v + q.w*uv + q.vec x uv          (evaluated (v + w*uv) + cross)
 *   q^-1       conjugate().coeffs() / squaredNorm() if squaredNorm() > 0
 *   reductions fixed-size unrolled redux: 3 -> c0 + (c1 + c2); 4 -> (c0+c1) + (c2+c3)
 *   normalized v / sqrt(squaredNorm) if squaredNorm > 0
 *   hnormalized (x/z, y/z)
 *
 * Semantics: IDEAL SET SEMANTICS for allocation/deletion: every requested
 * block is allocated in the frame it is first requested, every carve-eligible visible block
 * is deleted that frame.  The reference differs from this only when >= 2 distinct new blocks
 * contend for one bucket lock in one frame (voxel_hash.cu:83-88,105-117,140,159);
 * ref_hash_model.c restates that racy table for the known-answer tests and for the
 * don't-care mask used when comparing with the reference rebuild.
 *
 * Parity pin: see oracle/README.md and DESIGN.md ("oracle pinning").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_LEN 8
#define BLOCK_VOLUME 512

/* ------------------------------------------------------------------------------------------ */
/* small float32 vector helpers (operation order == Eigen 3.3.9 scalar path)                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } quat;
typedef struct { quat q; v3 t; } se3;
typedef struct { float fx, fy, cx, cy; } intr;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_muls(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_divs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline float v3_sqnorm(v3 a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
static inline float v3_norm(v3 a) { return sqrtf(v3_sqnorm(a)); }
static inline v3 v3_cross(v3 a, v3 b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* Eigen 3.3.9 Geometry/Quaternion.h _transformVector */
static inline v3 q_rot(quat q, v3 v) {
  v3 qv = V3(q.x, q.y, q.z);
  v3 uv = v3_cross(qv, v);
  uv = v3_add(uv, uv);
  v3 c = v3_cross(qv, uv);
  return V3((v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z);
}
/* utils/cuda/lie_group.cuh:33-36 */
static inline v3 se3_apply(se3 T, v3 v) { return v3_add(q_rot(T.q, v), T.t); }
/* utils/cuda/lie_group.cuh:25-27 + Eigen quaternion inverse */
static se3 se3_inverse(se3 T) {
  se3 r;
  float n2 = (T.q.x * T.q.x + T.q.y * T.q.y) + (T.q.z * T.q.z + T.q.w * T.q.w);
  if (n2 > 0.f) {
    r.q.x = (-T.q.x) / n2; r.q.y = (-T.q.y) / n2; r.q.z = (-T.q.z) / n2; r.q.w = T.q.w / n2;
  } else {
    r.q.x = r.q.y = r.q.z = r.q.w = 0.f;
  }
  r.t = q_rot(r.q, V3(-T.t.x, -T.t.y, -T.t.z));
  return r;
}
/* utils/cuda/camera.cuh:35-39 */
static inline intr intr_inverse(intr k) {
  float fxi = 1 / k.fx, fyi = 1 / k.fy;
  intr r = {fxi, fyi, -k.cx * fxi, -k.cy * fyi};
  return r;
}
/* utils/cuda/camera.cuh:48-51 */
static inline v3 intr_mul(intr k, v3 v) { return V3(k.fx * v.x + k.cx * v.z, k.fy * v.y + k.cy * v.z, v.z); }

/* float -> int conversion with the semantics of the reference's device code
 * (cvt.rzi.s32.f32: truncate, saturate, NaN -> 0) so that out-of-range projections
 * behave identically on the CPU. */
static inline int f2i(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.f) return 2147483647;
  if (f <= -2147483648.f) return (-2147483647 - 1);
  return (int)f;
}
static inline int16_t f2s(float f) { /* cvt.rzi.s16.f32 saturating */
  if (f != f) return 0;
  if (f >= 32767.f) return 32767;
  if (f <= -32768.f) return -32768;
  return (int16_t)f;
}
static inline uint8_t f2u8(float f) { /* cvt.rzi.u8.f32 saturating */
  if (f != f) return 0;
  if (f >= 255.f) return 255;
  if (f <= 0.f) return 0;
  return (uint8_t)f;
}

/* ------------------------------------------------------------------------------------------ */
/* block store: ideal set of 8^3 voxel blocks keyed by block coordinate                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int16_t pos[3];
  int alive;
  float tsdf[BLOCK_VOLUME];
  uint8_t rgbw[BLOCK_VOLUME * 4]; /* r,g,b,weight */
  float prob[BLOCK_VOLUME];
} oblock;

typedef struct tsdf_oracle {
  float voxel_size, truncation;
  oblock** blocks; /* block storage, indexed by id */
  int n_ids, cap_ids;
  int* free_ids; int n_free;
  /* open addressing map: key -> id */
  uint64_t* mkey; int* mval; uint64_t mcap; uint64_t mcount;
  int n_alive;
  int* vis; int vis_cap; /* scratch: visible ids */
} tsdf_oracle;

#define MK_EMPTY 0xFFFFFFFFFFFFFFFFull
static inline uint64_t pack_key(int16_t x, int16_t y, int16_t z) {
  return (uint64_t)(uint16_t)x | ((uint64_t)(uint16_t)y << 16) | ((uint64_t)(uint16_t)z << 32);
}
static inline uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
static void map_rehash(tsdf_oracle* o, uint64_t ncap) {
  uint64_t* ok = o->mkey; int* ov = o->mval; uint64_t oc = o->mcap;
  o->mkey = (uint64_t*)malloc(sizeof(uint64_t) * ncap);
  o->mval = (int*)malloc(sizeof(int) * ncap);
  for (uint64_t i = 0; i < ncap; ++i) o->mkey[i] = MK_EMPTY;
  o->mcap = ncap; o->mcount = 0;
  for (uint64_t i = 0; i < oc; ++i) {
    if (ok[i] == MK_EMPTY) continue;
    uint64_t s = mix64(ok[i]) & (ncap - 1);
    while (o->mkey[s] != MK_EMPTY) s = (s + 1) & (ncap - 1);
    o->mkey[s] = ok[i]; o->mval[s] = ov[i]; o->mcount++;
  }
  free(ok); free(ov);
}
static int map_find(const tsdf_oracle* o, uint64_t key) {
  uint64_t s = mix64(key) & (o->mcap - 1);
  while (o->mkey[s] != MK_EMPTY) {
    if (o->mkey[s] == key) return o->mval[s];
    s = (s + 1) & (o->mcap - 1);
  }
  return -1;
}
static void map_insert(tsdf_oracle* o, uint64_t key, int val) {
  if ((o->mcount + 1) * 2 > o->mcap) map_rehash(o, o->mcap * 2);
  uint64_t s = mix64(key) & (o->mcap - 1);
  while (o->mkey[s] != MK_EMPTY) s = (s + 1) & (o->mcap - 1);
  o->mkey[s] = key; o->mval[s] = val; o->mcount++;
}
static void map_erase(tsdf_oracle* o, uint64_t key) { /* backward-shift deletion */
  uint64_t m = o->mcap - 1, s = mix64(key) & m;
  while (o->mkey[s] != key) { if (o->mkey[s] == MK_EMPTY) return; s = (s + 1) & m; }
  uint64_t hole = s;
  for (;;) {
    s = (s + 1) & m;
    if (o->mkey[s] == MK_EMPTY) break;
    uint64_t home = mix64(o->mkey[s]) & m;
    /* can the element at s move into hole? yes iff home is not in (hole, s] cyclically */
    int between = (hole <= s) ? (home > hole && home <= s) : (home > hole || home <= s);
    if (!between) { o->mkey[hole] = o->mkey[s]; o->mval[hole] = o->mval[s]; hole = s; }
  }
  o->mkey[hole] = MK_EMPTY; o->mcount--;
}

tsdf_oracle* oracle_create(float voxel_size, float truncation) {
  tsdf_oracle* o = (tsdf_oracle*)calloc(1, sizeof(tsdf_oracle));
  o->voxel_size = voxel_size; o->truncation = truncation;
  o->mcap = 0; o->mkey = NULL; o->mval = NULL;
  map_rehash(o, 1u << 16);
  return o;
}
void oracle_destroy(tsdf_oracle* o) {
  if (!o) return;
  for (int i = 0; i < o->n_ids; ++i) free(o->blocks[i]);
  free(o->blocks); free(o->free_ids); free(o->mkey); free(o->mval); free(o->vis); free(o);
}
int oracle_num_blocks(const tsdf_oracle* o) { return o->n_alive; }

/* utils/tsdf/voxel_mem.cu:43-51: weight=0, tsdf=-1, prob=.5 (rgb is left stale by the
 * reference, i.e. unspecified; the oracle and the new engine define it as 0). */
static int block_new(tsdf_oracle* o, int16_t bx, int16_t by, int16_t bz) {
  int id;
  if (o->n_free > 0) {
    id = o->free_ids[--o->n_free];
  } else {
    if (o->n_ids == o->cap_ids) {
      o->cap_ids = o->cap_ids ? o->cap_ids * 2 : 1024;
      o->blocks = (oblock**)realloc(o->blocks, sizeof(oblock*) * o->cap_ids);
      o->free_ids = (int*)realloc(o->free_ids, sizeof(int) * o->cap_ids);
    }
    id = o->n_ids++;
    o->blocks[id] = (oblock*)malloc(sizeof(oblock));
  }
  oblock* b = o->blocks[id];
  b->pos[0] = bx; b->pos[1] = by; b->pos[2] = bz; b->alive = 1;
  for (int i = 0; i < BLOCK_VOLUME; ++i) { b->tsdf[i] = -1.f; b->prob[i] = .5f; }
  memset(b->rgbw, 0, sizeof(b->rgbw));
  map_insert(o, pack_key(bx, by, bz), id);
  o->n_alive++;
  return id;
}
static void block_delete(tsdf_oracle* o, int id) {
  oblock* b = o->blocks[id];
  map_erase(o, pack_key(b->pos[0], b->pos[1], b->pos[2]));
  b->alive = 0;
  o->free_ids[o->n_free++] = id;
  o->n_alive--;
}

/* utils/tsdf/voxel_hash.cu:31-35 -- short is sign-extended, then cast to uint */
uint32_t oracle_hash(int16_t x, int16_t y, int16_t z) {
  return (((uint32_t)(int32_t)x * 73856093u) ^ ((uint32_t)(int32_t)y * 19349669u) ^
          ((uint32_t)(int32_t)z * 83492791u)) & ((1u << 21) - 1);
}

/* ------------------------------------------------------------------------------------------ */
/* visibility: utils/tsdf/voxel_tsdf.cu:48-80                                                  */
/* ------------------------------------------------------------------------------------------ */
static inline int voxel_visible(int16_t gx, int16_t gy, int16_t gz, se3 cam_T_world, intr K,
                                int img_w, int img_h, float voxel_size) {
  v3 pw = v3_muls(V3((float)gx, (float)gy, (float)gz), voxel_size);
  v3 pc = se3_apply(cam_T_world, pw);
  v3 ph = intr_mul(K, pc);
  float u = ph.x / ph.z, v = ph.y / ph.z;
  return (u >= 0 && u <= (float)(img_w - 1) && v >= 0 && v <= (float)(img_h - 1) && ph.z >= 0);
}
static inline int block_visible(int full, int16_t bx, int16_t by, int16_t bz, se3 cam_T_world, intr K,
                                int img_w, int img_h, float voxel_size) {
  int16_t x = (int16_t)(bx << 3), y = (int16_t)(by << 3), z = (int16_t)(bz << 3);
  int visible = full;
  for (int i = 0; i < 8; ++i) {
    int16_t cx = (int16_t)(x + ((i >> 0) & 1) * (BLOCK_LEN - 1));
    int16_t cy = (int16_t)(y + ((i >> 1) & 1) * (BLOCK_LEN - 1));
    int16_t cz = (int16_t)(z + ((i >> 2) & 1) * (BLOCK_LEN - 1));
    int vv = voxel_visible(cx, cy, cz, cam_T_world, K, img_w, img_h, voxel_size);
    if (full) visible &= vv; else visible |= vv;
  }
  return visible;
}

/* ------------------------------------------------------------------------------------------ */
/* Integrate: utils/tsdf/voxel_tsdf.cu:347-375 (Allocate -> GatherVisible -> UpdateTSDF ->     */
/* SpaceCarving)                                                                               */
/* counters[8]: 0 n_active_pre, 1 n_new, 2 n_vis, 3 n_upd, 4 n_carved, 5 n_active_post,        */
/*              6 n_requests (pixel samples that passed the visibility test), 7 reserved       */
/* ------------------------------------------------------------------------------------------ */
int oracle_integrate(tsdf_oracle* o, const uint8_t* rgb, const float* depth, const float* ht,
                     const float* lt, int img_w, int img_h, float max_depth, const float Kp[4],
                     const float q_xyzw[4], const float t_xyz[3], int64_t* counters,
                     int16_t* new_keys_out, int new_keys_cap) {
  const float voxel_size = o->voxel_size, truncation = o->truncation;
  intr K = {Kp[0], Kp[1], Kp[2], Kp[3]};
  intr Ki = intr_inverse(K);
  se3 cam_T_world = {{q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]}, {t_xyz[0], t_xyz[1], t_xyz[2]}};
  se3 world_T_cam = se3_inverse(cam_T_world);
  int64_t n_pre = o->n_alive, n_new = 0, n_req = 0;

  float* range = (float*)malloc(sizeof(float) * (size_t)img_w * img_h);

  /* ---- block_allocate_kernel (voxel_tsdf.cu:104-147) ---- */
  for (int y = 0; y < img_h; ++y) {
    for (int x = 0; x < img_w; ++x) {
      const int idx = y * img_w + x;
      const float d = depth[idx];
      v3 pos_cam = intr_mul(Ki, V3((float)x, (float)y, 1.f));
      range[idx] = v3_norm(pos_cam);
      if (d == 0 || d > max_depth) continue;
      v3 pos_world = se3_apply(world_T_cam, v3_muls(pos_cam, d));
      v3 ray_dir_cam = v3_divs(pos_cam, range[idx]);
      v3 ray_dir_world = q_rot(world_T_cam.q, ray_dir_cam);
      v3 ray_start_world = v3_sub(pos_world, v3_muls(ray_dir_world, truncation));
      v3 ray_dir_grid = v3_divs(ray_dir_world, voxel_size);
      v3 ray_start_grid = v3_divs(ray_start_world, voxel_size);
      v3 ray_grid = v3_muls(ray_dir_grid, 2 * truncation);
      const int step_grid = f2i(ceilf(fmaxf(fmaxf(fabsf(ray_grid.x), fabsf(ray_grid.y)), fabsf(ray_grid.z)) / BLOCK_LEN));
      v3 ray_step_grid = v3_divs(ray_grid, fmaxf((float)step_grid, 1));
      v3 pos_grid = ray_start_grid;
      uint64_t last_key = MK_EMPTY;
      for (int i = 0; i <= step_grid; ++i, pos_grid = v3_add(pos_grid, ray_step_grid)) {
        int16_t px = f2s(roundf(pos_grid.x)), py = f2s(roundf(pos_grid.y)), pz = f2s(roundf(pos_grid.z));
        int16_t bx = (int16_t)(px >> 3), by = (int16_t)(py >> 3), bz = (int16_t)(pz >> 3);
        uint64_t key = pack_key(bx, by, bz);
        if (key == last_key) continue; /* pure speed-up: same decision as previous sample */
        last_key = key;
        if (map_find(o, key) >= 0) continue; /* Allocate() is a no-op for present blocks */
        if (block_visible(1, bx, by, bz, cam_T_world, K, img_w, img_h, voxel_size)) {
          ++n_req;
          if (new_keys_out && n_new < new_keys_cap) {
            new_keys_out[3 * n_new + 0] = bx; new_keys_out[3 * n_new + 1] = by; new_keys_out[3 * n_new + 2] = bz;
          }
          block_new(o, bx, by, bz);
          ++n_new;
        }
      }
    }
  }

  /* ---- check_visibility_kernel (voxel_tsdf.cu:82-93): any corner visible ---- */
  if (o->vis_cap < o->n_ids) { o->vis_cap = o->n_ids; o->vis = (int*)realloc(o->vis, sizeof(int) * o->vis_cap); }
  int n_vis = 0;
  for (int id = 0; id < o->n_ids; ++id) {
    oblock* b = o->blocks[id];
    if (!b->alive) continue;
    if (block_visible(0, b->pos[0], b->pos[1], b->pos[2], cam_T_world, K, img_w, img_h, voxel_size)) o->vis[n_vis++] = id;
  }

  /* ---- tsdf_integrate_kernel (voxel_tsdf.cu:149-205) ---- */
  int64_t n_upd = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : n_upd)
  for (int vi = 0; vi < n_vis; ++vi) {
    oblock* b = o->blocks[o->vis[vi]];
    const int16_t ox = (int16_t)(b->pos[0] << 3), oy = (int16_t)(b->pos[1] << 3), oz = (int16_t)(b->pos[2] << 3);
    for (int tz = 0; tz < 8; ++tz) for (int ty = 0; ty < 8; ++ty) for (int tx = 0; tx < 8; ++tx) {
      const int16_t gx = (int16_t)(ox + tx), gy = (int16_t)(oy + ty), gz = (int16_t)(oz + tz);
      v3 pos_world = v3_muls(V3((float)gx, (float)gy, (float)gz), voxel_size);
      v3 pos_cam = se3_apply(cam_T_world, pos_world);
      v3 pos_img_h = intr_mul(K, pos_cam);
      const int u = f2i(roundf(pos_img_h.x / pos_img_h.z));
      const int v = f2i(roundf(pos_img_h.y / pos_img_h.z));
      if (!(u >= 0 && u < img_w && v >= 0 && v < img_h)) continue;
      const int img_idx = v * img_w + u;
      const float d = depth[img_idx];
      if (d == 0 || d > max_depth) continue;
      const float sdf = range[img_idx] * (d - pos_img_h.z);
      if (!(sdf > -truncation)) continue;
      const float tsdf = fminf(1, sdf / truncation);
      const int vidx = tx + ty * 8 + tz * 64;
      const float weight_new = (1 - d / max_depth) * 4;
      const float weight_old = (float)b->rgbw[vidx * 4 + 3];
      const float weight_combined = weight_old + weight_new;
      uint8_t nrgb[3];
      for (int c = 0; c < 3; ++c) {
        const float rgb_old = (float)b->rgbw[vidx * 4 + c];
        const float rgb_new = (float)rgb[img_idx * 3 + c];
        nrgb[c] = f2u8(roundf((rgb_old * weight_old + rgb_new * weight_new) / weight_combined));
      }
      b->tsdf[vidx] = (b->tsdf[vidx] * weight_old + tsdf * weight_new) / weight_combined;
      b->rgbw[vidx * 4 + 3] = f2u8(fminf(roundf(weight_combined), 40));
      b->rgbw[vidx * 4 + 0] = nrgb[0]; b->rgbw[vidx * 4 + 1] = nrgb[1]; b->rgbw[vidx * 4 + 2] = nrgb[2];
      const float p = b->prob[vidx];
      const float positive = expf((weight_old * logf(p) + weight_new * logf(ht[img_idx])) / weight_combined);
      const float negative = expf((weight_old * logf(1 - p) + weight_new * logf(lt[img_idx])) / weight_combined);
      b->prob[vidx] = positive / (positive + negative);
      ++n_upd;
    }
  }

  /* ---- space_carving_kernel (voxel_tsdf.cu:207-230), threshold .9 (voxel_tsdf.cu:485) ---- */
  int64_t n_carved = 0;
  const float min_tsdf_threshold = .9;
  for (int vi = 0; vi < n_vis; ++vi) {
    oblock* b = o->blocks[o->vis[vi]];
    float m = fabsf(b->tsdf[0]);
    for (int i = 1; i < BLOCK_VOLUME; ++i) m = fminf(m, fabsf(b->tsdf[i]));
    if (m >= min_tsdf_threshold) { block_delete(o, o->vis[vi]); ++n_carved; }
  }

  free(range);
  if (counters) {
    counters[0] = n_pre; counters[1] = n_new; counters[2] = n_vis; counters[3] = n_upd;
    counters[4] = n_carved; counters[5] = o->n_alive; counters[6] = n_req; counters[7] = 0;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* export (canonical order: ascending (z, y, x) signed block coordinate)                        */
/* ------------------------------------------------------------------------------------------ */
static tsdf_oracle* g_sort_ctx;
static int cmp_ids(const void* a, const void* b) {
  const oblock* A = g_sort_ctx->blocks[*(const int*)a];
  const oblock* B = g_sort_ctx->blocks[*(const int*)b];
  for (int c = 2; c >= 0; --c) if (A->pos[c] != B->pos[c]) return A->pos[c] < B->pos[c] ? -1 : 1;
  return 0;
}
static int* sorted_alive_ids(tsdf_oracle* o, int* n_out) {
  int* ids = (int*)malloc(sizeof(int) * (o->n_alive > 0 ? o->n_alive : 1));
  int n = 0;
  for (int id = 0; id < o->n_ids; ++id) if (o->blocks[id]->alive) ids[n++] = id;
  g_sort_ctx = o;
  qsort(ids, n, sizeof(int), cmp_ids);
  *n_out = n;
  return ids;
}
/* keys: int16[3*n]; tsdf: float[512*n]; rgbw: uint8[2048*n]; prob: float[512*n] (any may be NULL) */
int oracle_export(tsdf_oracle* o, int16_t* keys, float* tsdf, uint8_t* rgbw, float* prob, int cap_blocks) {
  int n; int* ids = sorted_alive_ids(o, &n);
  int m = n < cap_blocks ? n : cap_blocks;
  for (int i = 0; i < m; ++i) {
    const oblock* b = o->blocks[ids[i]];
    if (keys) { keys[3 * i] = b->pos[0]; keys[3 * i + 1] = b->pos[1]; keys[3 * i + 2] = b->pos[2]; }
    if (tsdf) memcpy(tsdf + (size_t)i * 512, b->tsdf, sizeof(b->tsdf));
    if (rgbw) memcpy(rgbw + (size_t)i * 2048, b->rgbw, sizeof(b->rgbw));
    if (prob) memcpy(prob + (size_t)i * 512, b->prob, sizeof(b->prob));
  }
  free(ids);
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Gather: utils/tsdf/voxel_tsdf.cu:14-46,399-454; out = {x,y,z,tsdf} float32 records,         */
/* blocks in canonical order, voxels in x + 8y + 64z order.  bbox = {xmin,xmax,ymin,ymax,      */
/* zmin,zmax} in metres, or NULL for GatherValid.  Returns the number of voxels selected.      */
/* ------------------------------------------------------------------------------------------ */
int64_t oracle_gather(tsdf_oracle* o, const float* bbox, float* out, int64_t cap_voxels) {
  int n; int* ids = sorted_alive_ids(o, &n);
  int16_t g[6] = {0, 0, 0, 0, 0, 0};
  if (bbox) { /* voxel_tsdf.cuh:21-26 with scale = (float)(1. / voxel_size) (voxel_tsdf.cu:429) */
    const float scale = (float)(1. / (double)o->voxel_size);
    for (int i = 0; i < 6; ++i) g[i] = f2s(bbox[i] * scale);
  }
  int64_t nv = 0;
  for (int i = 0; i < n; ++i) {
    const oblock* b = o->blocks[ids[i]];
    const int16_t x = (int16_t)(b->pos[0] << 3), y = (int16_t)(b->pos[1] << 3), z = (int16_t)(b->pos[2] << 3);
    if (bbox) {
      if (!(x >= g[0] && y >= g[2] && z >= g[4] && x + BLOCK_LEN - 1 <= g[1] && y + BLOCK_LEN - 1 <= g[3] &&
            z + BLOCK_LEN - 1 <= g[5])) continue;
    }
    for (int k = 0; k < BLOCK_VOLUME; ++k) {
      if (out && nv < cap_voxels) {
        const int16_t gx = (int16_t)(x + (k & 7)), gy = (int16_t)(y + ((k >> 3) & 7)), gz = (int16_t)(z + (k >> 6));
        out[nv * 4 + 0] = (float)gx * o->voxel_size;
        out[nv * 4 + 1] = (float)gy * o->voxel_size;
        out[nv * 4 + 2] = (float)gz * o->voxel_size;
        out[nv * 4 + 3] = b->tsdf[k];
      }
      ++nv;
    }
  }
  free(ids);
  return nv;
}

/* ------------------------------------------------------------------------------------------ */
/* RayCast: utils/tsdf/voxel_tsdf.cu:232-307, 490-506.  Retrieve() of an absent voxel returns   */
/* the default-constructed voxel (voxel_hash.cuh:104-113, voxel_types.cu:3-11):                */
/* tsdf=+1, rgbw=0, prob=0.                                                                    */
/* hit_depth (new output, not in the reference): camera-space z of the refined hit position in */
/* metres, +inf for a miss.  counters[4]: 0 samples, 1 block switches, 2 hits, 3 reserved.     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t key; const oblock* b; int valid; } rcache;
static inline const oblock* rc_block(const tsdf_oracle* o, rcache* c, int16_t px, int16_t py, int16_t pz, int64_t* sw) {
  uint64_t key = pack_key((int16_t)(px >> 3), (int16_t)(py >> 3), (int16_t)(pz >> 3));
  if (!c->valid || c->key != key) {
    int id = map_find(o, key);
    c->key = key; c->b = id >= 0 ? o->blocks[id] : NULL; c->valid = 1;
    ++*sw;
  }
  return c->b;
}
static inline int vox_index(int16_t px, int16_t py, int16_t pz) { return (px & 7) + (py & 7) * 8 + (pz & 7) * 64; }
static inline float rc_tsdf(const tsdf_oracle* o, rcache* c, v3 p, int64_t* samples, int64_t* sw) {
  int16_t px = f2s(roundf(p.x)), py = f2s(roundf(p.y)), pz = f2s(roundf(p.z));
  const oblock* b = rc_block(o, c, px, py, pz, sw);
  ++*samples;
  return b ? b->tsdf[vox_index(px, py, pz)] : 1.f;
}
static inline float rc_tsdf_i(const tsdf_oracle* o, rcache* c, int px, int py, int pz, int64_t* samples, int64_t* sw) {
  const oblock* b = rc_block(o, c, (int16_t)px, (int16_t)py, (int16_t)pz, sw);
  ++*samples;
  return b ? b->tsdf[vox_index((int16_t)px, (int16_t)py, (int16_t)pz)] : 1.f;
}

int oracle_raycast(tsdf_oracle* o, float max_depth, int img_w, int img_h, const float Kp[4], const float q_xyzw[4],
                   const float t_xyz[3], uint8_t* img_rgba, uint8_t* img_normal, float* hit_depth, int64_t* counters) {
  const float voxel_size = o->voxel_size;
  const float step_size = o->truncation / 2; /* voxel_tsdf.cu:497 */
  intr K = {Kp[0], Kp[1], Kp[2], Kp[3]};
  intr Ki = intr_inverse(K);
  se3 cam_T_world = {{q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]}, {t_xyz[0], t_xyz[1], t_xyz[2]}};
  se3 world_T_cam = se3_inverse(cam_T_world);
  const int max_step = f2i(ceilf(max_depth / step_size));
  int64_t tot_samples = 0, tot_sw = 0, tot_hits = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tot_samples, tot_sw, tot_hits)
  for (int y = 0; y < img_h; ++y) {
    for (int x = 0; x < img_w; ++x) {
      const int idx = y * img_w + x;
      int64_t samples = 0, sw = 0;
      v3 pos_cam = intr_mul(Ki, V3((float)x, (float)y, 1.f));
      float sq = v3_sqnorm(pos_cam);
      v3 ray_dir_cam = sq > 0.f ? v3_divs(pos_cam, sqrtf(sq)) : pos_cam;
      v3 ray_dir_world = q_rot(world_T_cam.q, ray_dir_cam);
      v3 ray_step_grid = v3_divs(v3_muls(ray_dir_world, step_size), voxel_size);
      v3 pos_grid = v3_divs(world_T_cam.t, voxel_size);
      rcache cache = {0, NULL, 0};
      float tsdf_prev = rc_tsdf(o, &cache, pos_grid, &samples, &sw);
      pos_grid = v3_add(pos_grid, ray_step_grid);
      int hit = 0;
      for (int i = 1; i < max_step; ++i, pos_grid = v3_add(pos_grid, ray_step_grid)) {
        const float tsdf_curr = rc_tsdf(o, &cache, pos_grid, &samples, &sw);
        if (tsdf_prev > 0 && tsdf_curr <= 0 && (double)(tsdf_prev - tsdf_curr) <= 1.5) {
          v3 pos1 = v3_sub(pos_grid, ray_step_grid);
          v3 pos2 = pos_grid;
          v3 mid = v3_divs(v3_add(pos1, pos2), 2.f);
          for (;;) {
            v3 dd = v3_sub(pos1, pos2);
            if (!((double)v3_dot(dd, dd) > .1)) break;
            const float tm = rc_tsdf(o, &cache, mid, &samples, &sw);
            if (tm < 0) pos2 = mid; else pos1 = mid;
            mid = v3_divs(v3_add(pos1, pos2), 2.f);
          }
          const int fx_ = f2s(roundf(mid.x)), fy_ = f2s(roundf(mid.y)), fz_ = f2s(roundf(mid.z));
          const oblock* fb = rc_block(o, &cache, (int16_t)fx_, (int16_t)fy_, (int16_t)fz_, &sw);
          uint8_t r = 0, g = 0, b_ = 0; float prob = 0.f;
          if (fb) {
            const int k = vox_index((int16_t)fx_, (int16_t)fy_, (int16_t)fz_);
            r = fb->rgbw[k * 4]; g = fb->rgbw[k * 4 + 1]; b_ = fb->rgbw[k * 4 + 2]; prob = fb->prob[k];
          }
          const float gxp = rc_tsdf_i(o, &cache, fx_ + 1, fy_, fz_, &samples, &sw);
          const float gxn = rc_tsdf_i(o, &cache, fx_ - 1, fy_, fz_, &samples, &sw);
          const float gyp = rc_tsdf_i(o, &cache, fx_, fy_ + 1, fz_, &samples, &sw);
          const float gyn = rc_tsdf_i(o, &cache, fx_, fy_ - 1, fz_, &samples, &sw);
          const float gzp = rc_tsdf_i(o, &cache, fx_, fy_, fz_ + 1, &samples, &sw);
          const float gzn = rc_tsdf_i(o, &cache, fx_, fy_, fz_ - 1, &samples, &sw);
          v3 nrm = V3(gxp - gxn, gyp - gyn, gzp - gzn);
          v3 neg_dir = V3(-ray_dir_world.x, -ray_dir_world.y, -ray_dir_world.z);
          const float diffusivity = fmaxf(v3_dot(nrm, neg_dir) / v3_norm(nrm), 0);
          const float alpha = (float)(fmaxf((float)(prob - 0.5), 0) / .5);
          if (img_rgba) {
            img_rgba[idx * 4 + 0] = f2u8(alpha * 255 + (1 - alpha) * r);
            img_rgba[idx * 4 + 1] = f2u8((1 - alpha) * g);
            img_rgba[idx * 4 + 2] = f2u8((1 - alpha) * b_);
            img_rgba[idx * 4 + 3] = 255;
          }
          if (img_normal) {
            img_normal[idx * 4 + 0] = f2u8(alpha * 255 + (1 - alpha) * diffusivity * 255);
            img_normal[idx * 4 + 1] = f2u8((1 - alpha) * diffusivity * 255);
            img_normal[idx * 4 + 2] = f2u8((1 - alpha) * diffusivity * 255);
            img_normal[idx * 4 + 3] = 255;
          }
          if (hit_depth) {
            v3 pc = se3_apply(cam_T_world, v3_muls(mid, voxel_size));
            hit_depth[idx] = pc.z;
          }
          hit = 1;
          break;
        }
        tsdf_prev = tsdf_curr;
      }
      if (!hit) {
        if (img_rgba) memset(img_rgba + idx * 4, 0, 4);
        if (img_normal) memset(img_normal + idx * 4, 0, 4);
        if (hit_depth) hit_depth[idx] = INFINITY;
      }
      tot_samples += samples; tot_sw += sw; tot_hits += hit;
    }
  }
  if (counters) { counters[0] = tot_samples; counters[1] = tot_sw; counters[2] = tot_hits; counters[3] = 0; }
  return 0;
}

/* Look up one voxel (testing aid): returns 1 if the block exists. */
int oracle_get_voxel(tsdf_oracle* o, int px, int py, int pz, float* tsdf, uint8_t* rgbw4, float* prob) {
  int id = map_find(o, pack_key((int16_t)(px >> 3), (int16_t)(py >> 3), (int16_t)(pz >> 3)));
  if (id < 0) { if (tsdf) *tsdf = 1.f; if (rgbw4) memset(rgbw4, 0, 4); if (prob) *prob = 0.f; return 0; }
  const oblock* b = o->blocks[id];
  const int k = vox_index((int16_t)px, (int16_t)py, (int16_t)pz);
  if (tsdf) *tsdf = b->tsdf[k];
  if (rgbw4) memcpy(rgbw4, b->rgbw + k * 4, 4);
  if (prob) *prob = b->prob[k];
  return 1;
}

/* Overwrite one voxel of an existing block (testing aid: hand-built volumes); NULL leaves a field. */
int oracle_set_voxel(tsdf_oracle* o, int px, int py, int pz, const float* tsdf, const uint8_t* rgbw4, const float* prob) {
  int id = map_find(o, pack_key((int16_t)(px >> 3), (int16_t)(py >> 3), (int16_t)(pz >> 3)));
  if (id < 0) return 0;
  oblock* b = o->blocks[id];
  const int k = vox_index((int16_t)px, (int16_t)py, (int16_t)pz);
  if (tsdf) b->tsdf[k] = *tsdf;
  if (rgbw4) memcpy(b->rgbw + k * 4, rgbw4, 4);
  if (prob) b->prob[k] = *prob;
  return 1;
}

/* Remove a block (used to prune the volume to the reference's block set before comparing RayCast /
 * Gather outputs with fixtures produced by the racy reference table). */
int oracle_delete_block(tsdf_oracle* o, int bx, int by, int bz) {
  int id = map_find(o, pack_key((int16_t)bx, (int16_t)by, (int16_t)bz));
  if (id < 0) return 0;
  block_delete(o, id);
  return 1;
}

/* Insert a block directly with the acquire-time defaults (hash/pool unit tests). */
int oracle_allocate_block(tsdf_oracle* o, int bx, int by, int bz) {
  if (map_find(o, pack_key((int16_t)bx, (int16_t)by, (int16_t)bz)) >= 0) return 0;
  block_new(o, (int16_t)bx, (int16_t)by, (int16_t)bz);
  return 1;
}
