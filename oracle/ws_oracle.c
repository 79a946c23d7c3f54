/*
 * ws_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see ws_oracle.h).
 *
 * A plain-C restatement of the reference's CPU TSDF path.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).  Eigen's integer
 * semantics are reproduced by hand:
 *   - Vector3i arithmetic is int32 and wraps; a/b truncates toward zero;
 *   - Matrix<T,3,1>::norm() on an integer T is (T)std::sqrt((double)squaredNorm()),
 *     squaredNorm accumulated in T;
 *   - int * Matrix<long> promotes to long;
 *   - Matrix4f * 32768 then cast<int> truncates.
 * The reference's uninitialised `Point prev` (update_tsdf.cpp:448) is modelled as
 * "no previous voxel" at the start of every ray.
 */
#include "ws_oracle.h"

#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MR ORC_MATRIX_RESOLUTION
#define WR ORC_WEIGHT_RESOLUTION
#define CS ORC_CHUNK_SIZE

/* ------------------------------------------------------------------ helpers */

static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

/* x86 cvttsd2si semantics for out-of-range / NaN (what the reference binary does) */
static inline int32_t d2i(double v)
{
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int32_t)v;
}
static inline int64_t d2l(double v)
{
  if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)v;
}
static inline int32_t f2i(float v)
{
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
  return (int32_t)v;
}

/* Eigen Vector3i::norm(): squaredNorm in int32 (wrapping), sqrt in double, cast back to int */
static inline int32_t norm_i32(int32_t x, int32_t y, int32_t z)
{
  int32_t sq = wrap_add(wrap_add(wrap_mul(x, x), wrap_mul(y, y)), wrap_mul(z, z));
  return d2i(sqrt((double)sq));
}
/* Eigen Matrix<long,3,1>::norm() */
static inline int64_t norm_i64(int64_t x, int64_t y, int64_t z)
{
  int64_t sq = (int64_t)((uint64_t)x * (uint64_t)x + (uint64_t)y * (uint64_t)y + (uint64_t)z * (uint64_t)z);
  return d2l(sqrt((double)sq));
}

orc_entry orc_make_entry(int value, int weight)
{
  return (uint32_t)(uint16_t)(int16_t)value | ((uint32_t)(uint16_t)(int16_t)weight << 16);
}
int orc_entry_value(orc_entry e) { return (int16_t)(e & 0xFFFFu); }
int orc_entry_weight(orc_entry e) { return (int16_t)(e >> 16); }

/* ------------------------------------------------------ global chunk store
 * In-memory stand-in for HDF5GlobalMap (src/map/hdf5_global_map.cpp): chunks of 64^3 raw
 * entries addressed by chunk position, created default-filled on first activation
 * (:59-137).  The 64-entry LRU and the HDF5 file only decide WHERE a chunk lives, not its
 * contents, so they are not modelled. */

typedef struct {
  int32_t cx, cy, cz;
  int used;
  orc_entry *data;
} chunk_slot;

typedef struct {
  chunk_slot *slots;
  int64_t cap, count;
  orc_entry default_entry;
  int refs;
} chunk_store;

static uint64_t chunk_hash(int32_t x, int32_t y, int32_t z)
{
  uint64_t h = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull;
  h ^= (uint64_t)(uint32_t)y * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= (uint64_t)(uint32_t)z * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
  return h;
}

static chunk_store *store_create(orc_entry def)
{
  chunk_store *s = (chunk_store *)calloc(1, sizeof(*s));
  s->cap = 64;
  s->slots = (chunk_slot *)calloc((size_t)s->cap, sizeof(chunk_slot));
  s->default_entry = def;
  s->refs = 1;
  return s;
}

static void store_release(chunk_store *s)
{
  if (--s->refs > 0) return;
  for (int64_t i = 0; i < s->cap; i++)
    if (s->slots[i].used) free(s->slots[i].data);
  free(s->slots);
  free(s);
}

static chunk_slot *store_find(const chunk_store *s, int32_t x, int32_t y, int32_t z)
{
  uint64_t i = chunk_hash(x, y, z) & (uint64_t)(s->cap - 1);
  while (s->slots[i].used)
  {
    chunk_slot *c = &s->slots[i];
    if (c->cx == x && c->cy == y && c->cz == z) return c;
    i = (i + 1) & (uint64_t)(s->cap - 1);
  }
  return NULL;
}

static void store_grow(chunk_store *s)
{
  chunk_slot *old = s->slots;
  int64_t oldcap = s->cap;
  s->cap *= 2;
  s->slots = (chunk_slot *)calloc((size_t)s->cap, sizeof(chunk_slot));
  for (int64_t k = 0; k < oldcap; k++)
  {
    if (!old[k].used) continue;
    uint64_t i = chunk_hash(old[k].cx, old[k].cy, old[k].cz) & (uint64_t)(s->cap - 1);
    while (s->slots[i].used) i = (i + 1) & (uint64_t)(s->cap - 1);
    s->slots[i] = old[k];
  }
  free(old);
}

/* hdf5_global_map.cpp:59-137 (activate_chunk) */
static orc_entry *store_activate(chunk_store *s, int32_t x, int32_t y, int32_t z)
{
  chunk_slot *c = store_find(s, x, y, z);
  if (c) return c->data;
  if ((s->count + 1) * 2 > s->cap) store_grow(s);
  uint64_t i = chunk_hash(x, y, z) & (uint64_t)(s->cap - 1);
  while (s->slots[i].used) i = (i + 1) & (uint64_t)(s->cap - 1);
  c = &s->slots[i];
  c->cx = x; c->cy = y; c->cz = z; c->used = 1;
  size_t n = (size_t)CS * CS * CS;
  c->data = (orc_entry *)malloc(n * sizeof(orc_entry));
  for (size_t k = 0; k < n; k++) c->data[k] = s->default_entry;
  s->count++;
  return c->data;
}

/* include/map/util.h:5-12 (floor_divide goes through float) */
static inline int32_t floor_divide(int32_t a, int32_t b)
{
  return f2i(floorf((float)a / (float)b));
}

/* ------------------------------------------------------------- local map */

struct orc_map {
  int32_t size[3], pos[3], offset[3];
  orc_entry *data;
  chunk_store *store;
};

/* include/map/hdf5_local_map.h:4-19 */
static inline int32_t overflow_i(int32_t val, int32_t max)
{
  if (val >= 2 * max) return val - 2 * max;
  if (val >= max) return val - max;
  return val;
}

int64_t orc_map_num_voxels(const orc_map *m)
{
  return (int64_t)m->size[0] * m->size[1] * m->size[2];
}

/* src/map/hdf5_local_map.cpp:5-20 */
orc_map *orc_map_create(int sx, int sy, int sz, int default_value, int default_weight)
{
  orc_map *m = (orc_map *)calloc(1, sizeof(*m));
  m->size[0] = (sx % 2 == 1) ? sx : sx + 1;
  m->size[1] = (sy % 2 == 1) ? sy : sy + 1;
  m->size[2] = (sz % 2 == 1) ? sz : sz + 1;
  for (int a = 0; a < 3; a++) { m->pos[a] = 0; m->offset[a] = m->size[a] / 2; }
  m->store = store_create(orc_make_entry(default_value, default_weight));
  /* get_value(0,0,0) activates (creates) chunk 0_0_0 (hdf5_local_map.cpp:15) */
  orc_entry def = store_activate(m->store, 0, 0, 0)[0];
  int64_t n = orc_map_num_voxels(m);
  m->data = (orc_entry *)malloc((size_t)n * sizeof(orc_entry));
  for (int64_t i = 0; i < n; i++) m->data[i] = def;
  return m;
}

void orc_map_destroy(orc_map *m)
{
  if (!m) return;
  free(m->data);
  store_release(m->store);
  free(m);
}

/* src/map/hdf5_local_map.cpp:22-31 */
orc_map *orc_map_clone(const orc_map *src)
{
  orc_map *m = (orc_map *)calloc(1, sizeof(*m));
  memcpy(m->size, src->size, sizeof(m->size));
  memcpy(m->pos, src->pos, sizeof(m->pos));
  memcpy(m->offset, src->offset, sizeof(m->offset));
  int64_t n = orc_map_num_voxels(src);
  m->data = (orc_entry *)malloc((size_t)n * sizeof(orc_entry));
  memcpy(m->data, src->data, (size_t)n * sizeof(orc_entry));
  m->store = src->store;
  m->store->refs++;
  return m;
}

void orc_map_get_size(const orc_map *m, int out[3]) { memcpy(out, m->size, 12); }
void orc_map_get_pos(const orc_map *m, int out[3]) { memcpy(out, m->pos, 12); }
void orc_map_get_offset(const orc_map *m, int out[3]) { memcpy(out, m->offset, 12); }
orc_entry *orc_map_data(orc_map *m) { return m->data; }

void orc_map_set_state(orc_map *m, const int pos[3], const int offset[3])
{
  memcpy(m->pos, pos, 12);
  memcpy(m->offset, offset, 12);
}

/* include/map/hdf5_local_map.h:275-279 */
int orc_map_in_bounds(const orc_map *m, int x, int y, int z)
{
  int32_t dx = abs(wrap_sub(x, m->pos[0]));
  int32_t dy = abs(wrap_sub(y, m->pos[1]));
  int32_t dz = abs(wrap_sub(z, m->pos[2]));
  return dx <= m->size[0] / 2 && dy <= m->size[1] / 2 && dz <= m->size[2] / 2;
}

/* include/map/hdf5_local_map.h:140-151 -- the reference computes this in int; the oracle
 * uses int64 so that 2049^3 maps can be indexed (documented deviation). */
int64_t orc_map_index(const orc_map *m, int x, int y, int z)
{
  int64_t xo = overflow_i(x - m->pos[0] + m->offset[0] + m->size[0], m->size[0]);
  int64_t yo = overflow_i(y - m->pos[1] + m->offset[1] + m->size[1], m->size[1]);
  int64_t zo = overflow_i(z - m->pos[2] + m->offset[2] + m->size[2], m->size[2]);
  return (xo * m->size[1] + yo) * m->size[2] + zo;
}

int orc_map_get(const orc_map *m, int x, int y, int z, orc_entry *out)
{
  if (!orc_map_in_bounds(m, x, y, z)) return -1; /* std::out_of_range, hdf5_local_map.h:172-181 */
  *out = m->data[orc_map_index(m, x, y, z)];
  return 0;
}

int orc_map_set(orc_map *m, int x, int y, int z, orc_entry e)
{
  if (!orc_map_in_bounds(m, x, y, z)) return -1;
  m->data[orc_map_index(m, x, y, z)] = e;
  return 0;
}

/* src/map/hdf5_local_map.cpp:120-198 */
static void save_load_area(orc_map *m, const int32_t bottom[3], const int32_t top[3], int save)
{
  int32_t start[3], end[3], cstart[3], cend[3], sdelta[3], edelta[3];
  for (int a = 0; a < 3; a++)
  {
    start[a] = bottom[a] < top[a] ? bottom[a] : top[a];
    end[a] = bottom[a] > top[a] ? bottom[a] : top[a];
    cstart[a] = floor_divide(start[a], CS);
    cend[a] = floor_divide(end[a], CS);
    sdelta[a] = start[a] - cstart[a] * CS;
    edelta[a] = end[a] - cend[a] * CS;
  }
  for (int32_t cx = cstart[0]; cx <= cend[0]; ++cx)
    for (int32_t cy = cstart[1]; cy <= cend[1]; ++cy)
      for (int32_t cz = cstart[2]; cz <= cend[2]; ++cz)
      {
        orc_entry *chunk = store_activate(m->store, cx, cy, cz);
        int dxs = cx == cstart[0] ? sdelta[0] : 0;
        int dys = cy == cstart[1] ? sdelta[1] : 0;
        int dzs = cz == cstart[2] ? sdelta[2] : 0;
        int dxe = cx == cend[0] ? edelta[0] : CS - 1;
        int dye = cy == cend[1] ? edelta[1] : CS - 1;
        int dze = cz == cend[2] ? edelta[2] : CS - 1;
        for (int dx = dxs; dx <= dxe; ++dx)
          for (int dy = dys; dy <= dye; ++dy)
            for (int dz = dzs; dz <= dze; ++dz)
            {
              int idx = dx * CS * CS + dy * CS + dz;
              int64_t li = orc_map_index(m, cx * CS + dx, cy * CS + dy, cz * CS + dz);
              if (save) chunk[idx] = m->data[li];
              else m->data[li] = chunk[idx];
            }
      }
}

/* src/map/hdf5_local_map.cpp:53-118 */
int orc_map_shift(orc_map *m, const int new_pos[3])
{
  int32_t diff[3];
  for (int a = 0; a < 3; a++)
  {
    diff[a] = new_pos[a] - m->pos[a];
    if (abs(diff[a]) > m->size[a]) return -1; /* assert at :63-65 */
  }
  for (int axis = 0; axis < 3; axis++)
  {
    if (diff[axis] == 0) continue;
    int32_t start[3], end[3];
    for (int a = 0; a < 3; a++) { start[a] = m->pos[a] - m->size[a] / 2; end[a] = m->pos[a] + m->size[a] / 2; }
    if (diff[axis] > 0) end[axis] = start[axis] + diff[axis] - 1;
    else start[axis] = end[axis] + diff[axis] + 1;
    save_load_area(m, start, end, 1);

    m->pos[axis] += diff[axis];
    m->offset[axis] = (m->offset[axis] + diff[axis] + m->size[axis]) % m->size[axis];

    for (int a = 0; a < 3; a++) { start[a] = m->pos[a] - m->size[a] / 2; end[a] = m->pos[a] + m->size[a] / 2; }
    if (diff[axis] > 0) start[axis] = end[axis] - (diff[axis] - 1);
    else end[axis] = start[axis] - diff[axis] - 1;
    save_load_area(m, start, end, 0);
  }
  return 0;
}

/* src/map/hdf5_local_map.cpp:210-217 */
void orc_map_write_back(orc_map *m)
{
  int32_t start[3], end[3];
  for (int a = 0; a < 3; a++) { start[a] = m->pos[a] - m->size[a] / 2; end[a] = m->pos[a] + m->size[a] / 2; }
  save_load_area(m, start, end, 1);
}

int64_t orc_store_num_chunks(const orc_map *m) { return m->store->count; }

int orc_store_chunk_list(const orc_map *m, int *out_xyz, int64_t cap)
{
  int64_t k = 0;
  for (int64_t i = 0; i < m->store->cap && k < cap; i++)
  {
    if (!m->store->slots[i].used) continue;
    out_xyz[3 * k + 0] = m->store->slots[i].cx;
    out_xyz[3 * k + 1] = m->store->slots[i].cy;
    out_xyz[3 * k + 2] = m->store->slots[i].cz;
    k++;
  }
  return (int)k;
}

const orc_entry *orc_store_chunk(const orc_map *m, int cx, int cy, int cz)
{
  chunk_slot *c = store_find(m->store, cx, cy, cz);
  return c ? c->data : NULL;
}

/* src/map/hdf5_global_map.cpp:53-57,139-145 */
orc_entry orc_store_get_value(orc_map *m, int x, int y, int z)
{
  int32_t cx = floor_divide(x, CS), cy = floor_divide(y, CS), cz = floor_divide(z, CS);
  orc_entry *chunk = store_activate(m->store, cx, cy, cz);
  return chunk[(x - cx * CS) * CS * CS + (y - cy * CS) * CS + (z - cz * CS)];
}

/* ------------------------------------------------------- fixed-point helpers */

/* include/util/util.h:8-11 : (mat * MATRIX_RESOLUTION).cast<int>() */
void orc_to_int_mat(const float m[16], int32_t out[16])
{
  for (int i = 0; i < 16; i++) out[i] = f2i(m[i] * (float)MR);
}

/* include/util/util.h:13-18 : (mat * [p,1]).head<3>() / MATRIX_RESOLUTION, int32 */
orc_point orc_transform_point(orc_point p, const int32_t mat[16])
{
  int32_t v[4] = { p.x, p.y, p.z, 1 };
  int32_t r[3];
  for (int i = 0; i < 3; i++)
  {
    int32_t acc = 0;
    for (int k = 0; k < 4; k++) acc = wrap_add(acc, wrap_mul(mat[k * 4 + i], v[k]));
    r[i] = acc / MR;
  }
  orc_point o = { r[0], r[1], r[2] };
  return o;
}

void orc_transform_points(const orc_point *in, int64_t n, const int32_t mat[16], orc_point *out)
{
  for (int64_t i = 0; i < n; i++) out[i] = orc_transform_point(in[i], mat);
}

/* src/warpsense/app.cpp:118-148 (App::preprocess) without ROS: x/y/z floats in metres -> unique int32 mm
 * points in the map frame.  The reference iterates a std::unordered_set (bucket order, implementation
 * defined); this restatement emits the SAME SET in scan order (first occurrence stays), which is what the
 * device path produces.  Returns the number of points written to `out` (capacity n). */
int64_t orc_preprocess(const float *xyz, int64_t n, int stride_floats, const float pose[16], int map_resolution,
                       orc_point *out)
{
  int32_t im[16];
  orc_to_int_mat(pose, im);                                                   /* :123 */
  /* open-addressing set of indices into out[] */
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  int64_t *tab = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
  for (int64_t i = 0; i < cap; i++) tab[i] = -1;
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++)
  {
    const float x = xyz[i * stride_floats], y = xyz[i * stride_floats + 1], z = xyz[i * stride_floats + 2];
    if (x < 0.3 && y < 0.3 && z < 0.3) continue;                              /* :128-131 */
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;    /* (int)NaN is undefined behaviour in the reference: dropped, as on the device */
    const float mm[3] = { x * 1000.f, y * 1000.f, z * 1000.f };               /* :133 */
    orc_point c;
    int32_t cc[3];
    for (int a = 0; a < 3; a++)                                               /* :134-139 */
      cc[a] = f2i(floorf(mm[a] / (float)map_resolution) * (float)map_resolution + (float)(map_resolution / 2));
    c.x = cc[0]; c.y = cc[1]; c.z = cc[2];
    const orc_point p = orc_transform_point(c, im);                           /* :141 */
    uint64_t h = ((uint64_t)(uint32_t)p.x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)p.y * 0xC2B2AE3D27D4EB4Full) ^
                 ((uint64_t)(uint32_t)p.z * 0x165667B19E3779F9ull);
    h ^= h >> 29;
    int64_t s = (int64_t)(h & (uint64_t)(cap - 1));
    int dup = 0;
    while (tab[s] >= 0)
    {
      const orc_point o = out[tab[s]];
      if (o.x == p.x && o.y == p.y && o.z == p.z) { dup = 1; break; }
      s = (s + 1) & (cap - 1);
    }
    if (dup) continue;
    tab[s] = m;
    out[m++] = p;
  }
  free(tab);
  return m;
}

/* src/cpu/fastsense.cpp:143-163 (the CPU node's cloud callback): metres -> (int)(v * 1000), duplicates detected
 * by the voxel of that point (`point / MAP_RESOLUTION`, truncating), the point itself transformed and pushed in
 * scan order -- the reference's own order here. */
int64_t orc_preprocess_cpu_node(const float *xyz, int64_t n, int stride_floats, const float pose[16], int map_resolution,
                                orc_point *out)
{
  int32_t im[16];
  orc_to_int_mat(pose, im);                                                   /* :146 */
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  int64_t *tab = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
  orc_point *keys = (orc_point *)malloc((size_t)(n > 0 ? n : 1) * sizeof(orc_point));
  for (int64_t i = 0; i < cap; i++) tab[i] = -1;
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++)
  {
    const float x = xyz[i * stride_floats], y = xyz[i * stride_floats + 1], z = xyz[i * stride_floats + 2];
    if (x < 0.3 && y < 0.3 && z < 0.3) continue;                              /* :152-155 */
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;    /* (int)NaN is undefined behaviour in the reference: dropped, as on the device */
    orc_point pt = { f2i(x * 1000), f2i(y * 1000), f2i(z * 1000) };           /* :156-159 */
    const orc_point k = { pt.x / map_resolution, pt.y / map_resolution, pt.z / map_resolution };   /* :160 */
    uint64_t h = ((uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)k.y * 0xC2B2AE3D27D4EB4Full) ^
                 ((uint64_t)(uint32_t)k.z * 0x165667B19E3779F9ull);
    h ^= h >> 29;
    int64_t s = (int64_t)(h & (uint64_t)(cap - 1));
    int dup = 0;
    while (tab[s] >= 0)
    {
      const orc_point o = keys[tab[s]];
      if (o.x == k.x && o.y == k.y && o.z == k.z) { dup = 1; break; }
      s = (s + 1) & (cap - 1);
    }
    if (dup) continue;
    tab[s] = m;
    keys[m] = k;
    out[m++] = orc_transform_point(pt, im);                                   /* :162 */
  }
  free(tab); free(keys);
  return m;
}

/* include/util/util.h:52-56 */
void orc_to_map(const float pose[16], int map_resolution, int out[3])
{
  for (int a = 0; a < 3; a++) out[a] = f2i(floorf(pose[12 + a] / (float)map_resolution));
}

/* src/warpsense/tsdf_mapping.cpp:77-85 */
void orc_convert_pose(const float pose[16], int map_resolution, int pos[3], int up[3])
{
  int32_t im[16], rot[16];
  orc_to_int_mat(pose, im);
  memset(rot, 0, sizeof(rot));
  rot[0] = rot[5] = rot[10] = rot[15] = 1; /* Matrix4i::Identity() -- note: 1, not 32768 */
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) rot[c * 4 + r] = im[c * 4 + r];
  orc_point u = { 0, 0, MR };
  orc_point o = orc_transform_point(u, rot);
  up[0] = o.x; up[1] = o.y; up[2] = o.z;
  orc_to_map(pose, map_resolution, pos);
}

/* include/util/util.h:66-80 (round half away from zero, then truncate) */
void orc_transform_point_cloud(orc_point *pts, int64_t n, const float m[16])
{
  for (int64_t i = 0; i < n; i++)
  {
    float in[3] = { (float)pts[i].x, (float)pts[i].y, (float)pts[i].z };
    float t[3];
    for (int r = 0; r < 3; r++)
    {
      float acc = m[0 * 4 + r] * in[0];
      acc = acc + m[1 * 4 + r] * in[1];
      acc = acc + m[2 * 4 + r] * in[2];
      t[r] = acc + m[12 + r];
      if (t[r] < 0) t[r] -= 0.5f; else t[r] += 0.5f;
    }
    pts[i].x = f2i(t[0]); pts[i].y = f2i(t[1]); pts[i].z = f2i(t[2]);
  }
}

/* include/params/map_params.h:88-106 */
void orc_scale_params(float max_distance, int max_weight_in, const float size_m[3], int resolution,
                      int *tau, int *max_weight, int size_vox[3])
{
  *tau = (int)(max_distance * 1000.f);
  *max_weight = max_weight_in * WR;
  for (int a = 0; a < 3; a++)
  {
    int s = (int)size_m[a]; /* size.x() = size1d : float -> int (map_params.h:58-63) */
    s *= 1000;
    s /= resolution;
    size_vox[a] = s;
  }
}

/* include/warpsense/test/common.h:16-26 */
int orc_calc_weight(int value, int tau, int weight_epsilon)
{
  int weight = WR;
  if (value < -weight_epsilon) weight = WR * (tau + value) / (tau - weight_epsilon);
  return weight;
}

/* ------------------------------------------------ per-scan candidate map
 * Stand-in for std::unordered_map<Point, TSDFEntry> (update_tsdf.cpp:407).  Only in-bounds
 * voxels are ever inserted (:498-501), so the ring index is a unique key. */

typedef struct {
  int64_t *keys;   /* ring linear index, -1 = empty */
  orc_entry *vals;
  int64_t cap, count;
} scan_map;

static void scan_map_init(scan_map *s, int64_t cap)
{
  s->cap = cap; s->count = 0;
  s->keys = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
  s->vals = (orc_entry *)malloc((size_t)cap * sizeof(orc_entry));
  for (int64_t i = 0; i < cap; i++) s->keys[i] = -1;
}
static void scan_map_free(scan_map *s) { free(s->keys); free(s->vals); }

static inline uint64_t mix64(uint64_t k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

static void scan_map_grow(scan_map *s)
{
  scan_map n;
  scan_map_init(&n, s->cap * 2);
  for (int64_t i = 0; i < s->cap; i++)
  {
    if (s->keys[i] < 0) continue;
    uint64_t j = mix64((uint64_t)s->keys[i]) & (uint64_t)(n.cap - 1);
    while (n.keys[j] >= 0) j = (j + 1) & (uint64_t)(n.cap - 1);
    n.keys[j] = s->keys[i]; n.vals[j] = s->vals[i];
  }
  n.count = s->count;
  scan_map_free(s);
  *s = n;
}

/* returns slot; *inserted = 1 if the key was new (try_emplace semantics) */
static inline int64_t scan_map_emplace(scan_map *s, int64_t key, orc_entry val, int *inserted)
{
  if ((s->count + 1) * 10 > s->cap * 7) scan_map_grow(s);
  uint64_t j = mix64((uint64_t)key) & (uint64_t)(s->cap - 1);
  while (s->keys[j] >= 0)
  {
    if (s->keys[j] == key) { *inserted = 0; return (int64_t)j; }
    j = (j + 1) & (uint64_t)(s->cap - 1);
  }
  s->keys[j] = key; s->vals[j] = val; s->count++;
  *inserted = 1;
  return (int64_t)j;
}

/* ------------------------------------------------------- update_tsdf oracle */

/* One scan point: the ray march with its fan of interpolated voxels into `values`
 * (update_tsdf.cpp:418-514; the rmagine/OpenMP variant's loop body :589-669 is the same code minus the
 * scan-point bounds check :430-434, which `check_point` switches) */
typedef struct {
  int res, tau, weight_epsilon, dz_per_distance;
  int32_t pos[3];
  int64_t upl[3];
} march_consts;

static void march_point(const orc_map *m, const march_consts *k, const orc_point *pt, int check_point,
                        scan_map *values, orc_update_stats *st)
{
  const int res = k->res, tau = k->tau, weight_epsilon = k->weight_epsilon;
  const int32_t *pos = k->pos;
  const int64_t *upl = k->upl;
  const int32_t p[3] = { pt->x, pt->y, pt->z };
  int32_t d[3] = { wrap_sub(p[0], pos[0]), wrap_sub(p[1], pos[1]), wrap_sub(p[2], pos[2]) }; /* :422 */
  int32_t distance = norm_i32(d[0], d[1], d[2]);                            /* :423 */
  if (distance == 0) return;                                                /* :424-428 */

  if (check_point && !orc_map_in_bounds(m, p[0] / res, p[1] / res, p[2] / res)) return;  /* :430-434 */
  st->n_marched++;

  int64_t nd[3], c1[3], iv[3];
  for (int a = 0; a < 3; a++) nd[a] = ((int64_t)d[a] * MR) / distance;      /* :438 */
  c1[0] = nd[1] * upl[2] - nd[2] * upl[1];                                  /* :439 inner cross */
  c1[1] = nd[2] * upl[0] - nd[0] * upl[2];
  c1[2] = nd[0] * upl[1] - nd[1] * upl[0];
  for (int a = 0; a < 3; a++) c1[a] /= MR;
  iv[0] = nd[1] * c1[2] - nd[2] * c1[1];                                    /* :439 outer cross */
  iv[1] = nd[2] * c1[0] - nd[0] * c1[2];
  iv[2] = nd[0] * c1[1] - nd[1] * c1[0];
  int64_t inorm = norm_i64(iv[0], iv[1], iv[2]);                            /* :440 */
  if (inorm == 0) return;                                                   /* :441-445 */
  for (int a = 0; a < 3; a++) iv[a] = (iv[a] * MR) / inorm;                 /* :446 */

  int have_prev = 0;                                                        /* :448 (uninitialised) */
  int32_t prev[2] = { 0, 0 };

  for (int32_t len = 1; len <= distance + tau; len += res / 2)              /* :450 */
  {
    int32_t proj[3], index[3];
    for (int a = 0; a < 3; a++)
    {
      proj[a] = wrap_add(pos[a], wrap_mul(d[a], len) / distance);           /* :452 */
      index[a] = proj[a] / res;                                             /* :453 */
    }
    if (have_prev && index[0] == prev[0] && index[1] == prev[1]) continue;  /* :455-458 */
    prev[0] = index[0]; prev[1] = index[1]; have_prev = 1;                  /* :459 */
    if (!orc_map_in_bounds(m, index[0], index[1], index[2])) continue;      /* :460-463 */

    int32_t tc[3];
    for (int a = 0; a < 3; a++) tc[a] = wrap_add(wrap_mul(index[a], res), res / 2); /* :466 */
    long long value = norm_i32(wrap_sub(p[0], tc[0]), wrap_sub(p[1], tc[1]), wrap_sub(p[2], tc[2])); /* :467 */
    if (value > (long long)tau) value = tau;                                /* :468 */
    if (len > distance) value = -value;                                     /* :469-472 */

    int weight = WR;                                                        /* :475 */
    if (value < -weight_epsilon)
      weight = (int)(WR * (tau + value) / (tau - weight_epsilon));          /* :476-479 */
    if (weight == 0) continue;                                              /* :480-483 */

    int32_t delta_z = wrap_mul(k->dz_per_distance, len) / MR;               /* :485 */
    int32_t iter_steps = (delta_z * 2) / res + 1;                           /* :486 */
    int32_t mid = delta_z / res;                                            /* :487 */
    int32_t lowest[3];
    for (int a = 0; a < 3; a++)
      lowest[a] = wrap_sub(proj[a], (int32_t)(((int64_t)delta_z * iv[a]) / MR)); /* :488 */

    for (int32_t step = 0; step < iter_steps; ++step)                       /* :491 */
    {
      int32_t idx[3];
      for (int a = 0; a < 3; a++)
        idx[a] = wrap_add(lowest[a], (int32_t)(((int64_t)wrap_mul(step, res) * iv[a]) / MR)) / res; /* :493 */
      if (!orc_map_in_bounds(m, idx[0], idx[1], idx[2])) continue;          /* :495-498 */

      int w = (step != mid) ? -weight : weight;                             /* :503-506 */
      orc_entry tmp = orc_make_entry((int)value, w);
      st->n_candidates++;
      if (w < 0) st->n_neg_candidates++;

      int inserted;
      int64_t slot = scan_map_emplace(values, orc_map_index(m, idx[0], idx[1], idx[2]), tmp, &inserted); /* :508 */
      if (!inserted)
      {
        orc_entry ex = values->vals[slot];
        long long av = value < 0 ? -value : value;
        if (av < abs(orc_entry_value(ex)) || orc_entry_weight(ex) < 0)      /* :509 */
          values->vals[slot] = tmp;                                         /* :511 */
      }
    }
  }
}

/* one winner into the grid: update_tsdf.cpp:542-560 (= :708-721); returns 1 if the entry was written */
static int merge_winner(orc_entry *e, orc_entry win, int max_weight)
{
  int value = orc_entry_value(win);
  int weight = orc_entry_weight(win);
  int ev = orc_entry_value(*e), ew = orc_entry_weight(*e);
  if (weight > 0 && ew > 0)                                                 /* :546-551 */
  {
    int nv = (ev * ew + value * weight) / (ew + weight);
    int nw = (ew + weight) < max_weight ? (ew + weight) : max_weight;
    *e = orc_make_entry(nv, nw);
    return 1;
  }
  if (weight != 0 && ew <= 0)                                               /* :553-557 */
  {
    *e = orc_make_entry(value, weight);
    return 1;
  }
  return 0;
}

static void march_consts_init(march_consts *k, const int scanner_pos[3], const int up[3], int tau, int res)
{
  float angle = 45.f / 128.f;                                                 /* :400 */
  k->dz_per_distance = d2i(tan((double)(angle / 180) * M_PI) / 2.0 * MR);     /* :401 */
  k->weight_epsilon = tau / 10;                                               /* :403 */
  k->res = res; k->tau = tau;
  for (int a = 0; a < 3; a++)
  {
    k->pos[a] = wrap_mul(scanner_pos[a], res);                                /* :410 */
    k->upl[a] = up[a];
  }
}

/* src/cpu/update_tsdf.cpp:397-564 (up-vector ray march, Eigen variant, thread_count = 1) */
void orc_update_tsdf(orc_map *m, const orc_point *pts, int64_t n,
                     const int scanner_pos[3], const int up[3],
                     int tau, int max_weight, int map_resolution,
                     orc_update_stats *stats)
{
  march_consts k;
  march_consts_init(&k, scanner_pos, up, tau, map_resolution);
  orc_update_stats st;
  memset(&st, 0, sizeof(st));
  st.n_points = n;

  scan_map values;
  scan_map_init(&values, 1 << 16);
  for (int64_t pi = 0; pi < n; pi++) march_point(m, &k, &pts[pi], 1, &values, &st);

  /* merge into the grid: update_tsdf.cpp:517-561 (single thread => no cross-thread skip) */
  st.n_touched = values.count;
  for (int64_t s = 0; s < values.cap; s++)
  {
    if (values.keys[s] < 0) continue;
    st.n_written += merge_winner(&m->data[values.keys[s]], values.vals[s], max_weight);
  }
  scan_map_free(&values);
  if (stats) *stats = st;
}

static int64_t scan_map_find(const scan_map *s, int64_t key)
{
  uint64_t j = mix64((uint64_t)key) & (uint64_t)(s->cap - 1);
  while (s->keys[j] >= 0)
  {
    if (s->keys[j] == key) return (int64_t)j;
    j = (j + 1) & (uint64_t)(s->cap - 1);
  }
  return -1;
}

/* src/cpu/update_tsdf.cpp:566-724: the rmagine-point overload, the reference's only multi-threaded CPU update
 * (thread_count = omp_get_max_threads(), :575).  The scan is split statically over the threads (:586), each with
 * its own candidate map; after a barrier every thread applies those of its winners that no other thread beats
 * with a strictly smaller |value| (:674-693) -- a tie between threads is applied by both, and the preference for
 * real over interpolated winners does not cross threads, so its result differs from the sequential variant's
 * except with one thread.  No scan-point bounds check here (:589-599).  Kept for TIMING (bench.py's
 * cpu_baseline "omp_variant"); the parity oracle is orc_update_tsdf.
 * The reference applies the winners from all threads concurrently (a race on ties); here every voxel has one
 * owner thread that applies its winners in thread order, so the port is deterministic and as parallel. */
void orc_update_tsdf_omp(orc_map *m, const orc_point *pts, int64_t n,
                         const int scanner_pos[3], const int up[3],
                         int tau, int max_weight, int map_resolution, int threads,
                         orc_update_stats *stats)
{
  march_consts k;
  march_consts_init(&k, scanner_pos, up, tau, map_resolution);
  if (threads < 1) threads = omp_get_max_threads();                           /* :575 */
  scan_map *values = (scan_map *)malloc((size_t)threads * sizeof(scan_map));
  orc_update_stats *sts = (orc_update_stats *)calloc((size_t)threads, sizeof(orc_update_stats));
  char *skip_flags = NULL;
  for (int t = 0; t < threads; t++) scan_map_init(&values[t], 1 << 14);

#pragma omp parallel num_threads(threads)
  {
    const int t = omp_get_thread_num();
    /* schedule(static), :586: contiguous blocks of the scan in thread order */
    const int64_t per = n / threads, extra = n % threads;
    const int64_t lo = t * per + (t < extra ? t : extra), hi = lo + per + (t < extra ? 1 : 0);
    for (int64_t pi = lo; pi < hi; pi++) march_point(m, &k, &pts[pi], 0, &values[t], &sts[t]);
  }

  /* :674-693 the cross-thread rule, evaluated in parallel (read-only) ... */
  int64_t total_cap = 0;
  int64_t *cap_off = (int64_t *)malloc((size_t)(threads + 1) * sizeof(int64_t));
  for (int t = 0; t < threads; t++) { cap_off[t] = total_cap; total_cap += values[t].cap; }
  cap_off[threads] = total_cap;
  skip_flags = (char *)calloc((size_t)total_cap, 1);
#pragma omp parallel num_threads(threads)
  {
    const int t = omp_get_thread_num();
    for (int64_t s = 0; s < values[t].cap; s++)
    {
      if (values[t].keys[s] < 0) continue;
      const int mine = abs(orc_entry_value(values[t].vals[s]));
      for (int i = 0; i < threads; i++)
      {
        if (i == t) continue;
        int64_t o = scan_map_find(&values[i], values[t].keys[s]);
        if (o >= 0 && abs(orc_entry_value(values[i].vals[o])) < mine) { skip_flags[cap_off[t] + s] = 1; break; }
      }
    }
  }
  /* ... :695-721 applied by all threads; each owns the voxels with key % threads == its number and walks the
   * maps in thread order, which is the sequential thread-order result without the reference's race on ties */
  orc_update_stats st;
  memset(&st, 0, sizeof(st));
  st.n_points = n;
  int64_t written = 0;
#pragma omp parallel num_threads(threads) reduction(+ : written)
  {
    const int me = omp_get_thread_num();
    for (int t = 0; t < threads; t++)
      for (int64_t s = 0; s < values[t].cap; s++)
      {
        const int64_t key = values[t].keys[s];
        if (key < 0 || (int)(key % threads) != me || skip_flags[cap_off[t] + s]) continue;
        written += merge_winner(&m->data[key], values[t].vals[s], max_weight);
      }
  }
  st.n_written = written;
  for (int t = 0; t < threads; t++)
  {
    st.n_marched += sts[t].n_marched;
    st.n_candidates += sts[t].n_candidates;
    st.n_neg_candidates += sts[t].n_neg_candidates;
    st.n_touched += values[t].count;          /* per-thread maps: a voxel two threads reached counts twice */
    scan_map_free(&values[t]);
  }
  free(values); free(sts); free(skip_flags); free(cap_off);
  if (stats) *stats = st;
}

/* ----------------------------------------------------- registration oracle */

void orc_jacobi_2_h(const int64_t J[6], int64_t H[36])
{
  for (int c = 0; c < 6; c++)
    for (int r = 0; r < 6; r++) H[c * 6 + r] += J[r] * J[c];
}

static inline int lookup(const orc_map *m, int x, int y, int z, int *v, int *w)
{
  if (!orc_map_in_bounds(m, x, y, z)) return 0;
  orc_entry e = m->data[orc_map_index(m, x, y, z)];
  *v = orc_entry_value(e); *w = orc_entry_weight(e);
  return 1;
}

/* src/cpu/registration.cpp:52-118 -- one accumulation pass */
void orc_reg_step(const orc_map *m, const orc_point *pts, int64_t n,
                  const float T[16], int map_resolution,
                  int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt)
{
  int32_t center[3] = { f2i(T[12]), f2i(T[13]), f2i(T[14]) };                 /* :52 */
  int32_t M[16];
  orc_to_int_mat(T, M);                                                       /* :54 */

  int64_t Hs[36]; int64_t gs[6]; int64_t es = 0, cs = 0;
  memset(Hs, 0, sizeof(Hs)); memset(gs, 0, sizeof(gs));

#pragma omp parallel
  {
    int64_t lh[36]; int64_t lg[6]; int32_t le = 0, lc = 0;
    memset(lh, 0, sizeof(lh)); memset(lg, 0, sizeof(lg));
#pragma omp for schedule(static) nowait
    for (int64_t j = 0; j < n; j++)
    {
      orc_point q = orc_transform_point(pts[j], M);                           /* :63 */
      int bx = q.x / map_resolution, by = q.y / map_resolution, bz = q.z / map_resolution; /* :65 */
      int32_t px = wrap_sub(q.x, center[0]), py = wrap_sub(q.y, center[1]), pz = wrap_sub(q.z, center[2]); /* :66 */

      int cv, cw, v1, w1, v0, w0;
      if (!lookup(m, bx, by, bz, &cv, &cw)) continue;                         /* :70 (throws) */
      if (cw == 0) continue;                                                  /* :71-74 */
      int xn_v, xn_w, xl_v, xl_w, yn_v, yn_w, yl_v, yl_w, zn_v, zn_w, zl_v, zl_w;
      if (!lookup(m, bx + 1, by, bz, &xn_v, &xn_w)) continue;                 /* :76-81 (any throw skips) */
      if (!lookup(m, bx - 1, by, bz, &xl_v, &xl_w)) continue;
      if (!lookup(m, bx, by + 1, bz, &yn_v, &yn_w)) continue;
      if (!lookup(m, bx, by - 1, bz, &yl_v, &yl_w)) continue;
      if (!lookup(m, bx, by, bz + 1, &zn_v, &zn_w)) continue;
      if (!lookup(m, bx, by, bz - 1, &zl_v, &zl_w)) continue;

      int32_t gr[3] = { 0, 0, 0 };                                            /* :83 */
      v1 = xn_v; w1 = xn_w; v0 = xl_v; w0 = xl_w;
      if (w1 != 0 && w0 != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[0] = (v1 - v0) / 2; /* :85-88 */
      v1 = yn_v; w1 = yn_w; v0 = yl_v; w0 = yl_w;
      if (w1 != 0 && w0 != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[1] = (v1 - v0) / 2; /* :89-92 */
      v1 = zn_v; w1 = zn_w; v0 = zl_v; w0 = zl_w;
      if (w1 != 0 && w0 != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[2] = (v1 - v0) / 2; /* :93-96 */

      /* jacobi << point.cross(gradient).cast<long>(), gradient.cast<long>()  (:98): cross in int32 */
      int64_t J[6];
      J[0] = wrap_sub(wrap_mul(py, gr[2]), wrap_mul(pz, gr[1]));
      J[1] = wrap_sub(wrap_mul(pz, gr[0]), wrap_mul(px, gr[2]));
      J[2] = wrap_sub(wrap_mul(px, gr[1]), wrap_mul(py, gr[0]));
      J[3] = gr[0]; J[4] = gr[1]; J[5] = gr[2];

      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) lh[c * 6 + r] += J[r] * J[c];             /* :104 */
      for (int r = 0; r < 6; r++) lg[r] += J[r] * cv;                         /* :105 */
      le += abs(cv);                                                          /* :106 */
      lc++;                                                                   /* :107 */
    }
#pragma omp critical
    {
      for (int i = 0; i < 36; i++) Hs[i] += lh[i];                            /* :115-120 */
      for (int i = 0; i < 6; i++) gs[i] += lg[i];
      es += le; cs += lc;
    }
  }
  memcpy(H, Hs, sizeof(Hs)); memcpy(g, gs, sizeof(gs));
  *err = (int32_t)es; *cnt = (int32_t)cs;
}

/* 6x6 FP64 inverse by LU with partial pivoting, the algorithm behind Eigen's
 * Matrix<double,6,6>::inverse() (PartialPivLU).  A is column-major. */
static void inverse6(const double A[36], double inv[36])
{
  double lu[6][6];
  int perm[6];
  for (int r = 0; r < 6; r++) { perm[r] = r; for (int c = 0; c < 6; c++) lu[r][c] = A[c * 6 + r]; }
  for (int k = 0; k < 6; k++)
  {
    int piv = k; double best = fabs(lu[k][k]);
    for (int r = k + 1; r < 6; r++) if (fabs(lu[r][k]) > best) { best = fabs(lu[r][k]); piv = r; }
    if (piv != k)
    {
      for (int c = 0; c < 6; c++) { double t = lu[k][c]; lu[k][c] = lu[piv][c]; lu[piv][c] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    for (int r = k + 1; r < 6; r++) lu[r][k] /= lu[k][k];
    for (int r = k + 1; r < 6; r++)
      for (int c = k + 1; c < 6; c++) lu[r][c] -= lu[r][k] * lu[k][c];
  }
  for (int col = 0; col < 6; col++)
  {
    double y[6];
    for (int r = 0; r < 6; r++) y[r] = (perm[r] == col) ? 1.0 : 0.0;
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < r; c++) y[r] -= lu[r][c] * y[c];
    for (int r = 5; r >= 0; r--)
    {
      for (int c = r + 1; c < 6; c++) y[r] -= lu[r][c] * y[c];
      y[r] /= lu[r][r];
    }
    for (int r = 0; r < 6; r++) inv[col * 6 + r] = y[r];
  }
}

/* include/warpsense/registration/util.h:5-39 */
void orc_xi_to_transform(const double xi[6], const int center[3], float out[16])
{
  double theta = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);         /* :13 */
  double l[3] = { xi[0] / theta, xi[1] / theta, xi[2] / theta };              /* :14 */
  float L[3][3] = { { 0.f, (float)(-l[2]), (float)l[1] },                     /* :15-19 */
                    { (float)l[2], 0.f, (float)(-l[0]) },
                    { (float)(-l[1]), (float)l[0], 0.f } };
  float s = (float)sin(theta);                                                /* :21-22 */
  float c1 = (float)(1 - cos(theta));
  float A[3][3], R[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i][j] = c1 * L[i][j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      float b = A[i][0] * L[0][j];
      b = b + A[i][1] * L[1][j];
      b = b + A[i][2] * L[2][j];
      float id = (i == j) ? 1.f : 0.f;
      R[i][j] = (id + s * L[i][j]) + b;
    }
  float oc[3] = { -(float)center[0], -(float)center[1], -(float)center[2] };  /* :29 */
  float xf[3] = { (float)xi[3], (float)xi[4], (float)xi[5] };
  memset(out, 0, 16 * sizeof(float));
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) out[j * 4 + i] = R[i][j];
    float shift = R[i][0] * oc[0];                                            /* :34 */
    shift = shift + R[i][1] * oc[1];
    shift = shift + R[i][2] * oc[2];
    shift = shift + 0.f * 1.f;
    out[12 + i] = (shift + (float)center[i]) + xf[i];                         /* :36 */
  }
  out[15] = 1.f;
}

/* src/cpu/registration.cpp:128-157 */
float orc_reg_solve(const int64_t H[36], const int64_t g[6], int32_t err, int32_t cnt,
                    float alpha, float T[16], double xi_out[6])
{
  int center[3] = { f2i(T[12]), f2i(T[13]), f2i(T[14]) };                     /* :52 */
  double hf[36], gf[6], inv[36], xi[6];
  for (int i = 0; i < 36; i++) hf[i] = (double)H[i];                          /* :130 */
  for (int i = 0; i < 6; i++) gf[i] = (double)g[i];                           /* :131 */
  double damp = (double)(alpha * (float)cnt);                                 /* :134 float*int -> float */
  for (int i = 0; i < 6; i++) hf[i * 6 + i] += damp * 1.0;
  inverse6(hf, inv);                                                          /* :136 */
  for (int r = 0; r < 6; r++)
  {
    double acc = 0.0;
    for (int k = 0; k < 6; k++) acc += (-inv[k * 6 + r]) * gf[k];
    xi[r] = acc;
  }
  float X[16], N[16];
  orc_xi_to_transform(xi, center, X);                                         /* :139 */
  for (int c = 0; c < 4; c++)                                                 /* :143 total = transform*total */
    for (int r = 0; r < 4; r++)
    {
      float acc = X[0 * 4 + r] * T[c * 4 + 0];
      acc = acc + X[1 * 4 + r] * T[c * 4 + 1];
      acc = acc + X[2 * 4 + r] * T[c * 4 + 2];
      acc = acc + X[3 * 4 + r] * T[c * 4 + 3];
      N[c * 4 + r] = acc;
    }
  memcpy(T, N, sizeof(N));
  if (xi_out) memcpy(xi_out, xi, sizeof(xi));
  return (float)err / cnt;                                                    /* :145 */
}

/* src/cpu/registration.cpp:14-177 */
int orc_register_cloud(const orc_map *m, orc_point *cloud, int64_t n,
                       const float pretransform[16],
                       int max_iterations, float it_weight_gradient, float epsilon,
                       int map_resolution, float out[16],
                       int64_t *trace, int trace_cap)
{
  float T[16];
  memcpy(T, pretransform, sizeof(T));
  float alpha = 0;
  float previous_errors[4] = { 0, 0, 0, 0 };
  int finished = 0;
  int it = 0;
  for (int i = 0; i < max_iterations && !finished; i++)
  {
    int64_t H[36], g[6]; int32_t e, c;
    orc_reg_step(m, cloud, n, T, map_resolution, H, g, &e, &c);
    if (trace && i < trace_cap)
    {
      int64_t *t = trace + (int64_t)i * 29;
      int k = 0;
      for (int r = 0; r < 6; r++) for (int cc = r; cc < 6; cc++) t[k++] = H[cc * 6 + r];
      for (int r = 0; r < 6; r++) t[k++] = g[r];
      t[k++] = e; t[k++] = c;
    }
    float err = orc_reg_solve(H, g, e, c, alpha, T, NULL);
    alpha += it_weight_gradient;                                              /* :141 */
    if (fabs(err - previous_errors[2]) < epsilon && fabs(err - previous_errors[0]) < epsilon) /* :146 */
      finished = 1;
    for (int k = 1; k < 4; k++) previous_errors[k - 1] = previous_errors[k];  /* :151-154 */
    previous_errors[3] = err;
    it = i + 1;
  }
  int32_t M[16];
  orc_to_int_mat(T, M);                                                       /* :166 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) cloud[i] = orc_transform_point(cloud[i], M); /* :168-172 */
  memcpy(out, T, sizeof(T));
  return it;
}

/* ---- featsense feed (src/warpsense/tsdf_mapping.cpp:145-163) ------------------------------------------------ */
typedef struct { unsigned idx; unsigned pt; } vg_pair;
static int vg_cmp(const void *a, const void *b)
{
  const vg_pair *x = (const vg_pair *)a, *y = (const vg_pair *)b;
  if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
  return x->pt < y->pt ? -1 : (x->pt > y->pt ? 1 : 0);          /* the order a stable sort leaves */
}

int64_t orc_voxelgrid(const float *xyz, int64_t n, int stride, float leaf, float *out_xyz, orc_point *out_mm)
{
  if (n <= 0) return 0;
  const float inv = 1.f / leaf;                                  /* inverse_leaf_size_ */
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  int64_t finite = 0;
  for (int64_t i = 0; i < n; i++)                                /* getMinMax3D */
  {
    const float *p = xyz + i * stride;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    finite++;
    for (int a = 0; a < 3; a++) { if (p[a] < mn[a]) mn[a] = p[a]; if (p[a] > mx[a]) mx[a] = p[a]; }
  }
  if (finite == 0) return 0;
  int64_t d[3]; int min_b[3], div_b[3];
  for (int a = 0; a < 3; a++)
  {
    d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    min_b[a] = (int)floorf(mn[a] * inv);
    div_b[a] = (int)floorf(mx[a] * inv) - min_b[a] + 1;
  }
  if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX)                    /* leaf too small: PCL returns its input */
  {
    for (int64_t i = 0; i < n; i++)
    {
      const float *p = xyz + i * stride;
      if (out_xyz) { out_xyz[3 * i] = p[0]; out_xyz[3 * i + 1] = p[1]; out_xyz[3 * i + 2] = p[2]; }
      out_mm[i].x = (int)(p[0] * 1000.f); out_mm[i].y = (int)(p[1] * 1000.f); out_mm[i].z = (int)(p[2] * 1000.f);
    }
    return n;
  }
  vg_pair *iv = (vg_pair *)malloc((size_t)finite * sizeof(vg_pair));
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++)
  {
    const float *p = xyz + i * stride;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    const int i0 = (int)(floorf(p[0] * inv) - (float)min_b[0]);
    const int i1 = (int)(floorf(p[1] * inv) - (float)min_b[1]);
    const int i2 = (int)(floorf(p[2] * inv) - (float)min_b[2]);
    iv[m].idx = (unsigned)(i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1]);
    iv[m].pt = (unsigned)i;
    m++;
  }
  qsort(iv, (size_t)m, sizeof(vg_pair), vg_cmp);
  int64_t out = 0;
  for (int64_t a = 0; a < m;)
  {
    int64_t b = a;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    while (b < m && iv[b].idx == iv[a].idx)
    {
      const float *p = xyz + (int64_t)iv[b].pt * stride;
      sx += p[0]; sy += p[1]; sz += p[2];                          /* AccumulatorXYZ::add */
      b++;
    }
    const float fm = (float)(b - a);
    const float cx = sx / fm, cy = sy / fm, cz = sz / fm;          /* AccumulatorXYZ::get: xyz / n */
    if (out_xyz) { out_xyz[3 * out] = cx; out_xyz[3 * out + 1] = cy; out_xyz[3 * out + 2] = cz; }
    out_mm[out].x = (int)(cx * 1000.f); out_mm[out].y = (int)(cy * 1000.f); out_mm[out].z = (int)(cz * 1000.f);
    out++;
    a = b;
  }
  free(iv);
  return out;
}

void orc_mm_pose_from_isometry(const double P[16], float out[16])
{
#define M_(r, c) P[(c) * 4 + (r)]
  double q[4];   /* x y z w: Eigen QuaternionBase::operator=(MatrixBase) */
  double t = M_(0, 0) + M_(1, 1) + M_(2, 2);
  if (t > 0.0)
  {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M_(2, 1) - M_(1, 2)) * t;
    q[1] = (M_(0, 2) - M_(2, 0)) * t;
    q[2] = (M_(1, 0) - M_(0, 1)) * t;
  }
  else
  {
    int i = 0;
    if (M_(1, 1) > M_(0, 0)) i = 1;
    if (M_(2, 2) > M_(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(M_(i, i) - M_(j, j) - M_(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (M_(k, j) - M_(j, k)) * t;
    q[j] = (M_(j, i) + M_(i, j)) * t;
    q[k] = (M_(k, i) + M_(i, k)) * t;
  }
#undef M_
  /* QuaternionBase::toRotationMatrix */
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  for (int i = 0; i < 16; i++) out[i] = (i % 5 == 0) ? 1.f : 0.f;
  out[0] = (float)(1.0 - (tyy + tzz)); out[1] = (float)(txy + twz); out[2] = (float)(txz - twy);
  out[4] = (float)(txy - twz); out[5] = (float)(1.0 - (txx + tzz)); out[6] = (float)(tyz + twx);
  out[8] = (float)(txz + twy); out[9] = (float)(tyz - twx); out[10] = (float)(1.0 - (txx + tyy));
  for (int r = 0; r < 3; r++) out[12 + r] = (float)(P[12 + r] * 1000.0);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
