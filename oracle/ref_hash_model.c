/*
 * ref_hash_model.c -- CPU ORACLE (test infrastructure only, never linked into the product).
 *
 * Sequential restatement of the reference's racy bucket/list hash table and block pool, used
 *  (i) to pin the oracle against the reference's own known-answer tests
 *      (utils/tests/voxel_hash_test.cu:56-180, utils/tests/voxel_mem_test.cu:38-90), and
 * (ii) to compute which blocks are "don't-care" when the ideal-set oracle is compared with the
 *      reference CUDA rebuild (blocks that lose a bucket try-lock in a frame).
 *
 * Follows (paths relative to /root/reference):
 *   utils/tsdf/voxel_hash.cuh:13-25   table geometry (2^21 buckets x 2 entries)
 *   utils/tsdf/voxel_hash.cu:31-35    Hash
 *   utils/tsdf/voxel_hash.cu:58-120   Allocate (try-lock, in-bucket slot, overflow list append)
 *   utils/tsdf/voxel_hash.cu:122-171  Delete
 *   utils/tsdf/voxel_hash.cuh:124-161 RetrieveMutable (lookup part)
 *   utils/tsdf/voxel_mem.cu:37-61     AquireBlock / ReleaseBlock (stack heap, pops from the end)
 *
 * Threads of one launch are emulated one after another in thread-index order; locks persist
 * until ResetLocks, exactly like the device locks persist for the rest of the kernel.  For the
 * reference's Collision test this yields the same 2,3,4 active-count sequence as the lock-step
 * GPU execution (the loser of a bucket either fails atomicExch or, when run later, finds the
 * bucket full and then fails the lock of the list tail's bucket).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int16_t pos[3]; int16_t offset; int32_t idx; } ref_entry; /* voxel_mem.cuh:73-93, 12 B */

typedef struct ref_hash {
  int bucket_bits; uint32_t num_bucket, num_entry, bucket_mask, entry_mask;
  ref_entry* table; int32_t* locks;
  int num_block; int32_t* heap; int num_free;
} ref_hash;

static uint32_t ref_hash_fn(const ref_hash* h, int16_t x, int16_t y, int16_t z) {
  return (((uint32_t)(int32_t)x * 73856093u) ^ ((uint32_t)(int32_t)y * 19349669u) ^
          ((uint32_t)(int32_t)z * 83492791u)) & h->bucket_mask;
}
uint32_t refhash_hash(const ref_hash* h, int x, int y, int z) { return ref_hash_fn(h, (int16_t)x, (int16_t)y, (int16_t)z); }

ref_hash* refhash_create(int bucket_bits, int num_block) {
  ref_hash* h = (ref_hash*)calloc(1, sizeof(ref_hash));
  h->bucket_bits = bucket_bits;
  h->num_bucket = 1u << bucket_bits; h->num_entry = h->num_bucket * 2;
  h->bucket_mask = h->num_bucket - 1; h->entry_mask = h->num_entry - 1;
  h->table = (ref_entry*)calloc(h->num_entry, sizeof(ref_entry)); /* offset = 0 */
  for (uint32_t i = 0; i < h->num_entry; ++i) h->table[i].idx = -1; /* init_hash_table_kernel */
  h->locks = (int32_t*)calloc(h->num_bucket, sizeof(int32_t));
  h->num_block = num_block; h->num_free = num_block;
  h->heap = (int32_t*)malloc(sizeof(int32_t) * num_block);
  for (int i = 0; i < num_block; ++i) h->heap[i] = i; /* heap_init_kernel */
  return h;
}
void refhash_destroy(ref_hash* h) { if (h) { free(h->table); free(h->locks); free(h->heap); free(h); } }
void refhash_reset_locks(ref_hash* h) { memset(h->locks, 0, sizeof(int32_t) * h->num_bucket); }
int refhash_num_active(const ref_hash* h) { return h->num_block - h->num_free; }

static int try_lock(ref_hash* h, uint32_t b) { int was = h->locks[b]; h->locks[b] = 1; return was == 0; }
static int32_t acquire(ref_hash* h) { int idx = h->num_free--; return h->heap[idx - 1]; }
static void release(ref_hash* h, int32_t b) { h->heap[h->num_free++] = b; }
static int pos_eq(const ref_entry* e, int16_t x, int16_t y, int16_t z) { return e->pos[0] == x && e->pos[1] == y && e->pos[2] == z; }

/* returns 1 if this call inserted the block, 0 if it already existed, -1 if it lost a lock */
int refhash_allocate(ref_hash* h, int xi, int yi, int zi) {
  const int16_t x = (int16_t)xi, y = (int16_t)yi, z = (int16_t)zi;
  const uint32_t bucket = ref_hash_fn(h, x, y, z), entry = bucket << 1;
  for (int i = 0; i < 2; ++i) { const ref_entry* e = &h->table[entry + i]; if (pos_eq(e, x, y, z) && e->idx >= 0) return 0; }
  uint32_t last = entry + 1;
  while (h->table[last].offset) {
    last = (last + (uint32_t)(int32_t)h->table[last].offset) & h->entry_mask;
    const ref_entry* e = &h->table[last];
    if (pos_eq(e, x, y, z) && e->idx >= 0) return 0;
  }
  for (int i = 0; i < 2; ++i) {
    ref_entry* e = &h->table[entry + i];
    if (e->idx < 0) {
      if (try_lock(h, bucket)) { e->pos[0] = x; e->pos[1] = y; e->pos[2] = z; e->offset = 0; e->idx = acquire(h); return 1; }
      return -1;
    }
  }
  last = entry + 1;
  while (h->table[last].offset) last = (last + (uint32_t)(int32_t)h->table[last].offset) & h->entry_mask;
  const uint32_t bucket_last = last >> 1;
  uint32_t next = last;
  for (;;) {
    next = (next + 1) & h->entry_mask;
    if ((next & 1) != 1 && h->table[next].idx < 0) {
      const uint32_t bucket_next = next >> 1;
      /* short-circuit && as in voxel_hash.cu:105-106: the second lock is only tried if the first succeeds */
      if (try_lock(h, bucket_last) && try_lock(h, bucket_next)) {
        const uint32_t wrap = next > last ? 0 : h->num_entry;
        h->table[last].offset = (int16_t)(next + wrap - last);
        ref_entry* e = &h->table[next];
        e->pos[0] = x; e->pos[1] = y; e->pos[2] = z; e->offset = 0; e->idx = acquire(h);
        return 1;
      }
      return -1;
    }
  }
}

/* returns 1 deleted, 0 not found, -1 lost a lock */
int refhash_delete(ref_hash* h, int xi, int yi, int zi) {
  const int16_t x = (int16_t)xi, y = (int16_t)yi, z = (int16_t)zi;
  const uint32_t bucket = ref_hash_fn(h, x, y, z), entry = bucket << 1;
  { ref_entry* e = &h->table[entry]; /* NUM_ENTRY_PER_BUCKET - 1 == 1 non-head slot */
    if (pos_eq(e, x, y, z) && e->idx >= 0) { release(h, e->idx); e->offset = 0; e->idx = -1; return 1; } }
  uint32_t last = entry + 1;
  ref_entry* head = &h->table[last];
  if (pos_eq(head, x, y, z) && head->idx >= 0) {
    if (try_lock(h, bucket)) {
      const uint32_t nx = (last + (uint32_t)(int32_t)head->offset) & h->entry_mask;
      ref_entry* n = &h->table[nx];
      release(h, head->idx);
      head->pos[0] = n->pos[0]; head->pos[1] = n->pos[1]; head->pos[2] = n->pos[2];
      head->offset = n->offset ? (int16_t)(head->offset + n->offset) : 0;
      head->idx = n->idx;
      n->offset = 0; n->idx = -1;
      return 1;
    }
    return -1;
  }
  while (h->table[last].offset) {
    ref_entry* bl = &h->table[last];
    const uint32_t cur = (last + (uint32_t)(int32_t)bl->offset) & h->entry_mask;
    ref_entry* bc = &h->table[cur];
    if (pos_eq(bc, x, y, z) && bc->idx >= 0) {
      if (try_lock(h, bucket)) {
        bl->offset = bc->offset ? (int16_t)(bl->offset + bc->offset) : 0;
        release(h, bc->idx); bc->offset = 0; bc->idx = -1;
        return 1;
      }
      return -1;
    }
    last = cur;
  }
  return 0;
}

/* pool block index of a block coordinate, -1 if absent (lookup part of RetrieveMutable) */
int refhash_find(const ref_hash* h, int xi, int yi, int zi) {
  const int16_t x = (int16_t)xi, y = (int16_t)yi, z = (int16_t)zi;
  const uint32_t entry = ref_hash_fn(h, x, y, z) << 1;
  for (int i = 0; i < 2; ++i) { const ref_entry* e = &h->table[entry + i]; if (pos_eq(e, x, y, z) && e->idx >= 0) return e->idx; }
  uint32_t last = entry + 1;
  while (h->table[last].offset) {
    last = (last + (uint32_t)(int32_t)h->table[last].offset) & h->entry_mask;
    const ref_entry* e = &h->table[last];
    if (pos_eq(e, x, y, z) && e->idx >= 0) return e->idx;
  }
  return -1;
}

int refhash_pool_acquire(ref_hash* h) { return acquire(h); }
void refhash_pool_release(ref_hash* h, int b) { release(h, b); }
