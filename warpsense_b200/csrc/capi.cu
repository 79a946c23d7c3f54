// capi.cu -- the C ABI of include/warpsense_b200.h over the kernels in this directory.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <unistd.h>
#include "../../include/warpsense_b200.h"
#include "ws_internal.h"
#include <nvtx3/nvToolsExt.h>      // header-only NVTX: ranges around the phases of a scan (Nsight Systems / ncu --nvtx)

static_assert(sizeof(ws_point) == sizeof(ws_pt), "point layout");
#define WS_NSUM 29
#define WS_TRACE_CAP 256
#define WS_CHUNK 64

namespace {

struct DeviceGuard
{
  int prev = -1;
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename F>
int guarded(ws_handle *h, F &&f)
{
  if (!h) return WS_ERR_INVALID;
  try
  {
    DeviceGuard dg(h->device);
    return f();
  }
  catch (const std::invalid_argument &e) { h->last_error = e.what(); return WS_ERR_INVALID; }
  catch (const std::length_error &e) { h->last_error = e.what(); return WS_ERR_CAPACITY; }
  catch (const std::logic_error &e) { h->last_error = e.what(); return WS_ERR_STATE; }
  catch (const std::exception &e) { h->last_error = e.what(); return WS_ERR_CUDA; }
}

void ensure_mailbox(ws_handle *h)
{
  if (h->d_mail) return;
  const size_t bytes = ws_reg_mailbox_bytes(WS_MAX_PEERS);
  WS_CUDA_OK(cudaMalloc(&h->d_mail, bytes));
  WS_CUDA_OK(cudaMemset(h->d_mail, 0, bytes));
  WS_CUDA_OK(cudaDeviceSynchronize());
}

void detach_peers(ws_handle *h)
{
  for (int p = 0; p < WS_MAX_PEERS; p++)
  {
    if (h->peer_ipc[p] && h->peer_mail[p]) cudaIpcCloseMemHandle(h->peer_mail[p]);
    h->peer_mail[p] = nullptr; h->peer_ipc[p] = false;
  }
  h->peers_attached = false;
}

void free_handle(ws_handle *h)
{
  if (!h) return;
  DeviceGuard dg(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &t : h->timers) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
  cudaFree(h->g.grid); cudaFree(h->g.keys); cudaFree(h->g.brick_flag); cudaFree(h->g.brick_flag2); cudaFree(h->g.vstate); cudaFree(h->g.ffree); cudaFree(h->g.brick_slot_base);
  cudaFree(h->d_points); cudaFree(h->d_rays); cudaFree(h->d_grp_info); cudaFree(h->d_item_off); cudaFree(h->d_gen_list);
  cudaFree(h->d_vg_keys); cudaFree(h->d_vg_vals); cudaFree(h->d_vg_hist); cudaFree(h->d_vg_box); cudaFree(h->d_vg_xyz);
  cudaFree(h->d_pre_tmp); cudaFree(h->d_pre_val); cudaFree(h->d_pre_slot); cudaFree(h->d_pre_table); cudaFree(h->d_pre_tiles); cudaFree(h->d_pre_xyz); cudaFree(h->d_reg_points);
  cudaFree(h->d_counters); cudaFreeHost(h->h_counters);
  cudaFree(h->d_pend_addr); cudaFree(h->d_pend_prev); cudaFree(h->d_pend_key);
  cudaFree(h->d_active[0]); cudaFree(h->d_active[1]);
  cudaFree(h->d_rec); cudaFree(h->d_list); cudaFree(h->d_chunk_fill);
  for (auto &t : h->track)
  {
    cudaFree(t.d_pts); cudaFreeHost(t.h_acc); cudaFreeHost(t.h_pose); cudaFreeHost(t.h_ctr);
    if (t.copied) cudaEventDestroy(t.copied);
    if (t.done) cudaEventDestroy(t.done);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream2)
  {
    cudaStreamDestroy(h->stream2); cudaStreamDestroy(h->stream3);
    cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); cudaEventDestroy(h->ev_merged); cudaEventDestroy(h->ev_scanned);
  }
  cudaFree(h->d_shift_buf); cudaFreeHost(h->h_shift_buf);
  if (h->spill_fp) { std::fclose(h->spill_fp); std::remove(h->spill_path.c_str()); }
  detach_peers(h); cudaFree(h->d_mail); cudaFree(h->d_pose); cudaFreeHost(h->h_pose);
  cudaFree(h->d_acc); cudaFree(h->d_trace); cudaFreeHost(h->h_acc); cudaFree(h->d_reg_partials);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

void ensure_points(ws_pt **buf, size_t *cap, size_t n)
{
  if (n <= *cap) return;
  size_t want = std::max<size_t>(n, 1 << 17);
  if (*buf) WS_CUDA_OK(cudaFree(*buf));
  *buf = nullptr; *cap = 0;
  WS_CUDA_OK(cudaMalloc(buf, want * sizeof(ws_pt)));
  *cap = want;
}

// Host-only: which ring-x brick columns rank `rank` of `world` owns and which it keeps resident.
//   stripe_cols <= 0 : contiguous slabs in ring coordinates;
//   stripe_cols  > 0 : stripes of `stripe_cols` columns dealt round-robin to the ranks -- every rank then
//                      holds a part of every ray whatever the sensor position (balanced march), at the price
//                      of one halo column on each side of every stripe.
// Resident = owned columns + the columns holding the rows just outside (the registration stencil reads x+-1).
bool shard_layout(int size_x, int rank, int world, int stripe_cols, std::vector<char> &owned, std::vector<int> &cols)
{
  const int nbx = (size_x + WS_BRICK - 1) / WS_BRICK;
  owned.assign((size_t)nbx, 0);
  if (stripe_cols > 0 && nbx / stripe_cols >= world)
  {
    for (int c = 0; c < nbx; c++) owned[(size_t)c] = ((c / stripe_cols) % world) == rank;
  }
  else
  {
    const int c_lo = (int)((i64)nbx * rank / world), c_hi = (int)((i64)nbx * (rank + 1) / world);
    if (c_hi <= c_lo) return false;
    for (int c = c_lo; c < c_hi; c++) owned[(size_t)c] = 1;
  }
  cols.clear();
  std::vector<char> res((size_t)nbx, 0);
  for (int c = 0; c < nbx; c++)
  {
    if (!owned[(size_t)c]) continue;
    res[(size_t)c] = 1;
    if (world > 1)
    {
      const int lo_row = c * WS_BRICK, hi_row = std::min((c + 1) * WS_BRICK, size_x) - 1;
      res[(size_t)(((lo_row - 1 + size_x) % size_x) / WS_BRICK)] = 1;
      res[(size_t)(((hi_row + 1) % size_x) / WS_BRICK)] = 1;
    }
  }
  for (int c = 0; c < nbx; c++) if (res[(size_t)c]) cols.push_back(c);
  return !cols.empty();
}

// contiguous slabs: the owned rows as a range (ws_slab_layout)
bool slab_layout(int size_x, int rank, int world, int *own_lo, int *own_hi, std::vector<int> &cols)
{
  std::vector<char> owned;
  if (!shard_layout(size_x, rank, world, 0, owned, cols)) return false;
  int lo = -1, hi = -1;
  for (int c = 0; c < (int)owned.size(); c++) if (owned[(size_t)c]) { if (lo < 0) lo = c; hi = c; }
  *own_lo = lo * WS_BRICK;
  *own_hi = std::min((hi + 1) * WS_BRICK, size_x);
  return true;
}

int create_impl(const int32_t size[3], int tau, int max_weight, int res, int device, int rank, int world, ws_handle **out)
{
  if (!out) return WS_ERR_INVALID;
  *out = nullptr;
  if (!size || size[0] < 1 || size[1] < 1 || size[2] < 1) return WS_ERR_INVALID;
  if (size[0] % 2 == 0 || size[1] % 2 == 0 || size[2] % 2 == 0) return WS_ERR_INVALID;  // hdf5_local_map.cpp:6-8
  if (tau < 1 || tau > 32767 || res < 2 || max_weight < 0 || max_weight > 32767) return WS_ERR_INVALID;
  if (world < 1 || rank < 0 || rank >= world) return WS_ERR_INVALID;
  {
    // the candidate order field holds 2^15 march steps and 2^6 fan steps: a ray across the whole map (the longest
    // a point inside the map can give with the sensor inside it too) must fit, or every scan would fail AFTER
    // it has changed the map.  Rays from a sensor far outside the map are still checked per scan.
    const double diag = std::sqrt((double)size[0] * size[0] + (double)size[1] * size[1] + (double)size[2] * size[2]) * res;
    const double steps = (diag + tau) / (double)(res / 2 > 0 ? res / 2 : 1);
    const double dz = std::tan((double)(45.f / 128.f / 180.f) * M_PI) / 2.0 * WS_MR * (diag + tau) / WS_MR;
    if (steps > 32768.0 || 2.0 * dz / res + 1.0 > 64.0) return WS_ERR_INVALID;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return WS_ERR_CUDA;

  ws_handle *h = new ws_handle();
  h->device = device; h->tau = tau; h->max_weight = max_weight; h->res = res;
  h->rank = rank; h->world = world;
  h->default_entry = make_entry(tau, 0);
  try
  {
    DeviceGuard dg(device);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    GridDesc &g = h->g;
    for (int a = 0; a < 3; a++)
    {
      g.size[a] = size[a]; g.half[a] = size[a] / 2; g.pos[a] = 0; g.offset[a] = size[a] / 2;
      g.nb[a] = (size[a] + WS_BRICK - 1) / WS_BRICK;
    }
    if (g.nb[0] > WS_MAX_XBRICKS) throw std::invalid_argument("map too large along x");
    // x-slab residency: this rank owns ring-x brick columns [c_lo, c_hi) plus the columns holding the
    // rows just outside (the registration stencil reads x+-1)
    std::vector<int> cols;
    std::vector<char> owned;
    int stripe_cols = 0;
    if (const char *env = std::getenv("WS_STRIPE_COLS")) stripe_cols = std::atoi(env);
    if (!shard_layout(g.size[0], rank, world, stripe_cols, owned, cols))
      throw std::invalid_argument("more ranks than ring-x brick columns");
    for (int c = 0; c < WS_MAX_XBRICKS; c++) g.xown[c] = c < (int)owned.size() ? (unsigned char)owned[(size_t)c] : 0;
    g.full = (world == 1) ? 1 : 0;
    for (int c = 0; c < WS_MAX_XBRICKS; c++) g.xslot[c] = -1;
    for (size_t s = 0; s < cols.size(); s++) g.xslot[cols[s]] = (short)s;
    g.n_bricks = (i64)cols.size() * g.nb[1] * g.nb[2];
    if (g.n_bricks >= (1ll << 31)) throw std::invalid_argument("map too large");
    const size_t n_vox = (size_t)g.n_bricks * WS_BRICK_VOX;
    g.wide = (n_vox > 0xFFFFFFFFull || std::getenv("WS_FORCE_WIDE")) ? 1 : 0;

    WS_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    WS_CUDA_OK(cudaMalloc(&g.grid, n_vox * sizeof(uint32_t)));
    WS_CUDA_OK(cudaMalloc(&g.keys, n_vox * sizeof(u64)));
    WS_CUDA_OK(cudaMalloc(&g.brick_flag, (size_t)g.n_bricks * sizeof(unsigned)));
    WS_CUDA_OK(cudaMalloc(&g.brick_slot_base, (size_t)g.n_bricks * sizeof(unsigned)));
    WS_CUDA_OK(cudaMemsetAsync(g.brick_slot_base, 0xFF, (size_t)g.n_bricks * sizeof(unsigned), h->stream));
    WS_CUDA_OK(cudaMalloc(&g.brick_flag2, (size_t)g.n_bricks * sizeof(unsigned)));
    WS_CUDA_OK(cudaMemsetAsync(g.brick_flag2, 0, (size_t)g.n_bricks * sizeof(unsigned), h->stream));
    WS_CUDA_OK(cudaMalloc(&g.vstate, n_vox / 2));
    WS_CUDA_OK(cudaMemsetAsync(g.vstate, 0, n_vox / 2, h->stream));
    WS_CUDA_OK(cudaMalloc(&g.ffree, n_vox * 2));
    WS_CUDA_OK(cudaMemsetAsync(g.ffree, 0, n_vox * 2, h->stream));
    WS_CUDA_OK(cudaMalloc(&h->d_counters, sizeof(UpdateCounters)));
    WS_CUDA_OK(cudaMallocHost(&h->h_counters, sizeof(UpdateCounters)));
    std::memset(h->h_counters, 0, sizeof(UpdateCounters));
    // parked voxels per scan: a fraction of the voxels near the surfaces; 1/8 of the resident voxels (at least
    // 16 Mi, at most 512 Mi slots of 32 bytes) unless WS_PENDING_CAP says otherwise
    size_t pcap = std::min<size_t>(std::max<size_t>(16u << 20, n_vox / 8), (size_t)512 << 20);
    if (const char *env = std::getenv("WS_PENDING_CAP")) pcap = (size_t)std::strtoull(env, nullptr, 10);
    pcap = std::min(pcap, n_vox);
    pcap = std::max<size_t>(pcap, 1);
    h->pending_cap = (unsigned)pcap;
    WS_CUDA_OK(cudaMalloc(&h->d_pend_addr, pcap * sizeof(u64)));
    WS_CUDA_OK(cudaMalloc(&h->d_pend_prev, pcap * sizeof(u64)));
    WS_CUDA_OK(cudaMalloc(&h->d_pend_key, pcap * sizeof(u64)));
    WS_CUDA_OK(cudaMalloc(&h->d_active[0], pcap * sizeof(unsigned)));
    WS_CUDA_OK(cudaMalloc(&h->d_active[1], pcap * sizeof(unsigned)));
    // far-field candidate record: starts at WS_RECORD_CAP records (default: half a record per voxel, between
    // 1 Mi and 32 Mi; 16 bytes each, and as much again for the replay list) and grows on demand up to
    // WS_RECORD_MAX records (default 2 Gi)
    size_t rcap = std::min<size_t>(32u << 20, std::max<size_t>(1u << 20, n_vox / 2)), rmax = (size_t)2 << 30;
    if (const char *env = std::getenv("WS_RECORD_CAP")) rcap = (size_t)std::strtoull(env, nullptr, 10);
    if (const char *env = std::getenv("WS_RECORD_MAX")) rmax = (size_t)std::strtoull(env, nullptr, 10);
    rmax = std::max(rmax, rcap);
    ws_update_alloc(h, std::max<size_t>(1, rcap / WS_REC_CHUNK), std::max<size_t>(1, rmax / WS_REC_CHUNK));
    WS_CUDA_OK(cudaMalloc(&h->d_acc, sizeof(RegAccum) + 16 * sizeof(float)));
    WS_CUDA_OK(cudaMemsetAsync(h->d_acc, 0, sizeof(RegAccum) + 16 * sizeof(float), h->stream));
    WS_CUDA_OK(cudaMallocHost(&h->h_acc, sizeof(RegAccum)));
    std::memset(h->h_acc, 0, sizeof(RegAccum));
    h->trace_cap = WS_TRACE_CAP;
    WS_CUDA_OK(cudaMalloc(&h->d_trace, (size_t)WS_TRACE_CAP * WS_NSUM * sizeof(u64)));
    if (const char *env = std::getenv("WS_STORE_CHUNKS")) h->store_max_chunks = std::max<size_t>(1, (size_t)std::strtoull(env, nullptr, 10));
    ws_launch_fill(h, h->default_entry);
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  catch (const std::invalid_argument &)
  {
    free_handle(h);
    return WS_ERR_INVALID;
  }
  catch (const std::exception &)
  {
    free_handle(h);
    return WS_ERR_CUDA;
  }
  *out = h;
  return WS_OK;
}

// ---- chunk store (src/map/hdf5_global_map.cpp) -------------------------------------------------
inline u64 chunk_key(int x, int y, int z)
{
  return ((u64)(uint32_t)(x + (1 << 20)) << 42) | ((u64)(uint32_t)(y + (1 << 20)) << 21) | (u64)(uint32_t)(z + (1 << 20));
}
inline void chunk_unkey(u64 k, int &x, int &y, int &z)
{
  x = (int)((k >> 42) & 0x1FFFFF) - (1 << 20);
  y = (int)((k >> 21) & 0x1FFFFF) - (1 << 20);
  z = (int)(k & 0x1FFFFF) - (1 << 20);
}
// include/map/util.h:5-12
inline int floor_divide(int a, int b) { return (int)std::floor((float)a / (float)b); }

// hdf5_global_map.cpp:59-137 (activate_chunk): default-filled on first use.  Like the reference's 64 active
// chunks (hdf5_global_map.h:78) the store keeps a bounded number of chunks in memory (WS_STORE_CHUNKS, default
// 4096 = 4 GiB); the least recently used one is written to a spill file (WS_STORE_SPILL_DIR, default /tmp) and read
// back on its next activation -- the reference writes it to its HDF5 dataset at that point (:96-107).
constexpr size_t CHUNK_ENTRIES = (size_t)WS_CHUNK * WS_CHUNK * WS_CHUNK;

void store_spill_open(ws_handle *h)
{
  if (h->spill_fp) return;
  const char *dir = std::getenv("WS_STORE_SPILL_DIR");
  char path[512];
  std::snprintf(path, sizeof(path), "%s/ws_chunk_spill_%d_%p.bin", dir ? dir : "/tmp", (int)getpid(), (void *)h);
  h->spill_path = path;
  h->spill_fp = std::fopen(path, "w+b");
  if (!h->spill_fp) throw std::runtime_error(std::string("chunk store: cannot open spill file ") + path);
}

void store_evict_one(ws_handle *h)
{
  const u64 key = h->lru.back();
  auto it = h->store.find(key);
  store_spill_open(h);
  long slot;
  auto sp = h->spilled.find(key);
  if (sp != h->spilled.end()) slot = sp->second;                 // rewrite in place
  else { slot = (long)h->spill_slots++; h->spilled[key] = slot; }
  if (std::fseek(h->spill_fp, slot * (long)(CHUNK_ENTRIES * sizeof(uint32_t)), SEEK_SET) != 0 ||
      std::fwrite(it->second.data.data(), sizeof(uint32_t), CHUNK_ENTRIES, h->spill_fp) != CHUNK_ENTRIES)
    throw std::runtime_error("chunk store: spill write failed");
  h->lru.pop_back();
  h->store.erase(it);
  h->store_evictions++;
}

uint32_t *activate_chunk(ws_handle *h, int cx, int cy, int cz)
{
  const u64 key = chunk_key(cx, cy, cz);
  auto it = h->store.find(key);
  if (it != h->store.end())
  {
    h->lru.splice(h->lru.begin(), h->lru, it->second.pos);       // most recently used
    return it->second.data.data();
  }
  while (h->store.size() >= h->store_max_chunks && !h->lru.empty()) store_evict_one(h);
  StoredChunk &c = h->store[key];
  auto sp = h->spilled.find(key);
  if (sp != h->spilled.end())
  {
    c.data.resize(CHUNK_ENTRIES);
    if (std::fseek(h->spill_fp, sp->second * (long)(CHUNK_ENTRIES * sizeof(uint32_t)), SEEK_SET) != 0 ||
        std::fread(c.data.data(), sizeof(uint32_t), CHUNK_ENTRIES, h->spill_fp) != CHUNK_ENTRIES)
      throw std::runtime_error("chunk store: spill read failed");
  }
  else c.data.assign(CHUNK_ENTRIES, h->default_entry);
  h->lru.push_front(key);
  c.pos = h->lru.begin();
  return c.data.data();
}

// every chunk the store knows (in memory or spilled), visited in turn with its data
template <typename F>
void store_for_each(ws_handle *h, F &&f)
{
  std::vector<u64> keys;
  for (const auto &kv : h->store) keys.push_back(kv.first);
  for (const auto &kv : h->spilled) if (!h->store.count(kv.first)) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  for (u64 k : keys)
  {
    int x, y, z;
    chunk_unkey(k, x, y, z);
    f(x, y, z, activate_chunk(h, x, y, z));
  }
}

bool column_resident(const GridDesc &g, int x)
{
  if (g.full) return true;
  return g.xslot[ring_coord(x, g.pos[0], g.offset[0], g.size[0]) >> 3] >= 0;
}

// hdf5_local_map.cpp:120-198 (save_load_area) on the device-resident grid, in x-slices of bounded size.
// Staging buffers (device + pinned host) live in the handle; chunk <-> slab copies go a z-run at a time.
void ensure_shift_buffers(ws_handle *h, size_t entries)
{
  if (entries <= h->shift_cap) return;
  WS_CUDA_OK(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_shift_buf); cudaFreeHost(h->h_shift_buf);
  h->d_shift_buf = nullptr; h->h_shift_buf = nullptr; h->shift_cap = 0;
  WS_CUDA_OK(cudaMalloc(&h->d_shift_buf, entries * sizeof(uint32_t)));
  if (cudaMallocHost(&h->h_shift_buf, entries * sizeof(uint32_t)) != cudaSuccess)
  {
    cudaFree(h->d_shift_buf); h->d_shift_buf = nullptr;
    throw std::runtime_error("cudaMallocHost(shift buffer)");
  }
  h->shift_cap = entries;
}

void save_load_area(ws_handle *h, const int bottom[3], const int top[3], bool save)
{
  int start[3], end[3];
  for (int a = 0; a < 3; a++) { start[a] = std::min(bottom[a], top[a]); end[a] = std::max(bottom[a], top[a]); }
  const i64 plane = (i64)(end[1] - start[1] + 1) * (end[2] - start[2] + 1);
  const int max_nx = (int)std::max<i64>(1, (i64)(32 << 20) / plane);
  ensure_shift_buffers(h, (size_t)plane * std::min(max_nx, end[0] - start[0] + 1));
  uint32_t *d_buf = h->d_shift_buf, *h_buf = h->h_shift_buf;
  for (int x0 = start[0]; x0 <= end[0]; x0 += max_nx)
  {
    const int x1 = std::min(end[0], x0 + max_nx - 1);
    const int lo[3] = { x0, start[1], start[2] };
    const int ext[3] = { x1 - x0 + 1, end[1] - start[1] + 1, end[2] - start[2] + 1 };
    const size_t n = (size_t)ext[0] * ext[1] * ext[2];
    if (save)
    {
      ws_box_transfer(h, d_buf, lo, ext, true);
      WS_CUDA_OK(cudaMemcpyAsync(h_buf, d_buf, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
      WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    // chunk-wise traversal, as the reference does (touch each chunk once)
    const int cs[3] = { floor_divide(lo[0], WS_CHUNK), floor_divide(lo[1], WS_CHUNK), floor_divide(lo[2], WS_CHUNK) };
    const int ce[3] = { floor_divide(x1, WS_CHUNK), floor_divide(end[1], WS_CHUNK), floor_divide(end[2], WS_CHUNK) };
    for (int cx = cs[0]; cx <= ce[0]; ++cx)
      for (int cy = cs[1]; cy <= ce[1]; ++cy)
        for (int cz = cs[2]; cz <= ce[2]; ++cz)
        {
          const int xs = std::max(lo[0], cx * WS_CHUNK), xe = std::min(x1, cx * WS_CHUNK + WS_CHUNK - 1);
          const int ys = std::max(lo[1], cy * WS_CHUNK), ye = std::min(end[1], cy * WS_CHUNK + WS_CHUNK - 1);
          const int zs = std::max(lo[2], cz * WS_CHUNK), ze = std::min(end[2], cz * WS_CHUNK + WS_CHUNK - 1);
          // a sharded handle only holds (and stores) the columns resident on its rank
          bool any_resident = false;
          for (int x = xs; x <= xe && !any_resident; ++x) any_resident = column_resident(h->g, x);
          if (!any_resident) continue;
          uint32_t *chunk = activate_chunk(h, cx, cy, cz);
          const size_t run = (size_t)(ze - zs + 1);
          for (int x = xs; x <= xe; ++x)
          {
            if (!column_resident(h->g, x)) continue;
            for (int y = ys; y <= ye; ++y)
            {
              uint32_t *b = h_buf + ((size_t)(x - lo[0]) * ext[1] + (size_t)(y - lo[1])) * ext[2] + (zs - lo[2]);
              uint32_t *c = chunk + ((size_t)(x - cx * WS_CHUNK) * WS_CHUNK + (size_t)(y - cy * WS_CHUNK)) * WS_CHUNK + (zs - cz * WS_CHUNK);
              if (save) std::memcpy(c, b, run * sizeof(uint32_t));
              else std::memcpy(b, c, run * sizeof(uint32_t));
            }
          }
        }
    if (!save)
    {
      WS_CUDA_OK(cudaMemcpyAsync(d_buf, h_buf, n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
      ws_box_transfer(h, d_buf, lo, ext, false);
      WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
  }
}

void read_acc(ws_handle *h)
{
  WS_CUDA_OK(cudaMemcpyAsync(h->h_acc, h->d_acc, sizeof(RegAccum), cudaMemcpyDeviceToHost, h->stream));
  WS_CUDA_OK(cudaStreamSynchronize(h->stream));
}

void sums_to_hg(const u64 sums[32], int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt)
{
  int k = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++)
    {
      const int64_t v = (int64_t)sums[k++];
      H[j * 6 + i] = v;
      H[i * 6 + j] = v;
    }
  for (int i = 0; i < 6; i++) g[i] = (int64_t)sums[21 + i];
  *err = (int32_t)(int64_t)sums[27];
  *cnt = (int32_t)(int64_t)sums[28];
}

}  // namespace

void ws_timer_begin(ws_handle *h, int kind, cudaStream_t stream)
{
  static const char *const names[] = { "ws:march", "ws:merge", "ws:register", "ws:replay" };
  nvtxRangePushA(names[kind & 3]);
  h->timer_open = false;
  if (!h->profile || (h->profile == 2 && kind != WS_TIMER_REG)) return;      // level 2: registration + whole update only
  if (h->timers_used >= h->timers.size())
  {
    if (h->timers.size() >= (1u << 16)) { h->timers_used = h->timers.size() + 1; return; }
    WsTimer t;
    if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) return;
    h->timers.push_back(t);
    h->timer_kind.push_back(kind);
  }
  h->timer_kind[h->timers_used] = kind;
  cudaEventRecord(h->timers[h->timers_used].start, stream ? stream : h->stream);
  h->timer_open = true;
}

void ws_timer_end(ws_handle *h, cudaStream_t stream)
{
  nvtxRangePop();
  if (!h->profile || !h->timer_open) return;
  h->timer_open = false;
  if (h->timers_used < h->timers.size())
  {
    cudaEventRecord(h->timers[h->timers_used].stop, stream ? stream : h->stream);
    h->timers_used++;
  }
}

// a range that spans others (the whole update_tsdf of one scan): takes its own slot, returns it for ws_span_end
long ws_span_begin(ws_handle *h, int kind)
{
  if (!h->profile) return -1;
  if (h->timers_used >= h->timers.size())
  {
    if (h->timers.size() >= (1u << 16)) return -1;
    WsTimer t;
    if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) return -1;
    h->timers.push_back(t);
    h->timer_kind.push_back(kind);
  }
  const long idx = (long)h->timers_used++;
  h->timer_kind[idx] = kind;
  cudaEventRecord(h->timers[idx].start, h->stream);
  return idx;
}

void ws_span_end(ws_handle *h, long idx)
{
  if (idx < 0 || (size_t)idx >= h->timers.size()) return;
  cudaEventRecord(h->timers[idx].stop, h->stream);
}

std::vector<unsigned long long> ws_store_keys(ws_handle *h)
{
  std::vector<unsigned long long> keys;
  for (const auto &kv : h->store) keys.push_back(kv.first);
  for (const auto &kv : h->spilled) if (!h->store.count(kv.first)) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  return keys;
}

const uint32_t *ws_store_chunk(ws_handle *h, int cx, int cy, int cz) { return activate_chunk(h, cx, cy, cz); }

extern "C" {

const char *ws_version(void) { return "warpsense_b200 0.1.0 (sm_100a)"; }

int ws_create(const int32_t size[3], int32_t tau, int32_t max_weight, int32_t map_resolution, int32_t device, ws_handle **out)
{
  return create_impl(size, tau, max_weight, map_resolution, device, 0, 1, out);
}

int ws_create_sharded(const int32_t size[3], int32_t tau, int32_t max_weight, int32_t map_resolution,
                      int32_t device, int32_t rank, int32_t world, ws_handle **out)
{
  return create_impl(size, tau, max_weight, map_resolution, device, rank, world, out);
}

void ws_destroy(ws_handle *h) { free_handle(h); }

const char *ws_last_error(const ws_handle *h) { return h ? h->last_error.c_str() : "null handle"; }

int ws_set_stream(ws_handle *h, void *cuda_stream)
{
  return guarded(h, [&]() {
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (cuda_stream == nullptr)
    {
      if (!h->own_stream)
      {
        WS_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
      }
    }
    else
    {
      if (h->own_stream) cudaStreamDestroy(h->stream);
      h->own_stream = false;
      h->stream = (cudaStream_t)cuda_stream;
    }
    return WS_OK;
  });
}

void *ws_get_stream(ws_handle *h) { return h ? (void *)h->stream : nullptr; }

int ws_sync(ws_handle *h)
{
  return guarded(h, [&]() { WS_CUDA_OK(cudaStreamSynchronize(h->stream)); return WS_OK; });
}

int ws_map_upload(ws_handle *h, const uint32_t *entries, const int32_t size[3], const int32_t offset[3], const int32_t pos[3])
{
  return guarded(h, [&]() {
    if (!entries || !size || !offset || !pos) throw std::invalid_argument("ws_map_upload: null argument");
    for (int a = 0; a < 3; a++)
      if (size[a] != h->g.size[a]) throw std::invalid_argument("ws_map_upload: size differs from the handle's map");
    for (int a = 0; a < 3; a++)
    {
      if (offset[a] < 0 || offset[a] >= size[a]) throw std::invalid_argument("ws_map_upload: offset out of range");
      h->g.offset[a] = offset[a]; h->g.pos[a] = pos[a];
    }
    ws_launch_upload(h, entries);
    return WS_OK;
  });
}

int ws_map_download(ws_handle *h, uint32_t *entries)
{
  return guarded(h, [&]() {
    if (!entries) throw std::invalid_argument("ws_map_download: null argument");
    ws_launch_download(h, entries);
    return WS_OK;
  });
}

int ws_map_set_params(ws_handle *h, const int32_t offset[3], const int32_t pos[3])
{
  return guarded(h, [&]() {
    if (!offset || !pos) throw std::invalid_argument("ws_map_set_params: null argument");
    for (int a = 0; a < 3; a++)
    {
      if (offset[a] < 0 || offset[a] >= h->g.size[a]) throw std::invalid_argument("ws_map_set_params: offset out of range");
      h->g.offset[a] = offset[a]; h->g.pos[a] = pos[a];
    }
    return WS_OK;
  });
}

int ws_map_get_params(const ws_handle *h, int32_t size[3], int32_t offset[3], int32_t pos[3])
{
  if (!h) return WS_ERR_INVALID;
  for (int a = 0; a < 3; a++)
  {
    if (size) size[a] = h->g.size[a];
    if (offset) offset[a] = h->g.offset[a];
    if (pos) pos[a] = h->g.pos[a];
  }
  return WS_OK;
}

int ws_map_fill(ws_handle *h, int32_t value, int32_t weight)
{
  return guarded(h, [&]() {
    h->default_entry = make_entry(value, weight);   // also what unseen chunks of the store hold
    ws_launch_fill(h, h->default_entry);
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    return WS_OK;
  });
}

static i64 voxel_addr(ws_handle *h, int x, int y, int z)
{
  const GridDesc &g = h->g;
  if (!grid_in_bounds(g, x, y, z)) throw std::invalid_argument("voxel out of bounds");
  const int rx = ring_coord(x, g.pos[0], g.offset[0], g.size[0]);
  const int ry = ring_coord(y, g.pos[1], g.offset[1], g.size[1]);
  const int rz = ring_coord(z, g.pos[2], g.offset[2], g.size[2]);
  const i64 b = brick_of(g, rx, ry, rz);
  if (b < 0) throw std::invalid_argument("voxel not resident on this rank");
  return b * WS_BRICK_VOX + brick_local(rx, ry, rz);
}

int ws_map_checksum(ws_handle *h, int32_t x_lo, int32_t x_hi, int32_t owned_only, uint64_t *out)
{
  return guarded(h, [&]() {
    if (!out) throw std::invalid_argument("ws_map_checksum: null argument");
    u64 *scratch = nullptr;
    WS_CUDA_OK(cudaMalloc(&scratch, sizeof(u64)));
    ws_launch_checksum(h, x_lo, x_hi, owned_only, scratch);
    u64 v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, scratch, sizeof(u64), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(scratch);
    WS_CUDA_OK(e);
    *out = v;
    return WS_OK;
  });
}

int ws_map_get_voxel(ws_handle *h, int32_t x, int32_t y, int32_t z, uint32_t *entry)
{
  return guarded(h, [&]() {
    if (!entry) throw std::invalid_argument("ws_map_get_voxel: null argument");
    const i64 addr = voxel_addr(h, x, y, z);
    WS_CUDA_OK(cudaMemcpyAsync(entry, h->g.grid + addr, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    return WS_OK;
  });
}

int ws_map_set_voxel(ws_handle *h, int32_t x, int32_t y, int32_t z, uint32_t entry)
{
  return guarded(h, [&]() {
    const i64 addr = voxel_addr(h, x, y, z);
    WS_CUDA_OK(cudaMemcpyAsync(h->g.grid + addr, &entry, sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    return WS_OK;
  });
}

static void check_update_args(int64_t n, const int32_t *scanner_pos, const int32_t *up)
{
  if (n < 0 || !scanner_pos || !up) throw std::invalid_argument("ws_update_tsdf: bad argument");
  if (n > WS_MAX_POINTS) throw std::length_error("ws_update_tsdf: too many points (reference limit semantics, update_tsdf.cu:146-150)");
}

int ws_update_tsdf(ws_handle *h, const ws_point *points, int64_t n, const int32_t scanner_pos[3], const int32_t up[3])
{
  return guarded(h, [&]() {
    check_update_args(n, scanner_pos, up);
    if (n > 0 && !points) throw std::invalid_argument("ws_update_tsdf: null points");
    ensure_points(&h->d_points, &h->points_cap, (size_t)n);
    if (n > 0)
      WS_CUDA_OK(cudaMemcpyAsync(h->d_points, points, (size_t)n * sizeof(ws_pt), cudaMemcpyHostToDevice, h->stream));
    h->last_n_points = n;
    ws_launch_update(h, h->d_points, (int)n, scanner_pos, up);
    return WS_OK;
  });
}

int ws_update_tsdf_device(ws_handle *h, const ws_point *device_points, int64_t n, const int32_t scanner_pos[3], const int32_t up[3])
{
  return guarded(h, [&]() {
    check_update_args(n, scanner_pos, up);
    if (n > 0 && !device_points) throw std::invalid_argument("ws_update_tsdf_device: null points");
    h->last_n_points = n;
    ws_launch_update(h, reinterpret_cast<const ws_pt *>(device_points), (int)n, scanner_pos, up);
    return WS_OK;
  });
}

int ws_preprocess_scan(ws_handle *h, const float *xyz, int64_t n, int32_t point_step_bytes, int32_t on_device,
                       const float pose_mm[16], int32_t map_resolution, ws_point *out_host, int64_t *n_out)
{
  const int mode = (on_device & WS_PREPROCESS_CPU_NODE) ? 1 : 0;
  on_device &= 1;
  return guarded(h, [&]() {
    if (n < 0 || (n > 0 && !xyz) || !pose_mm || map_resolution < 2 || point_step_bytes < 12 || (point_step_bytes & 3))
      throw std::invalid_argument("ws_preprocess_scan: bad argument");
    if (n > WS_MAX_POINTS) throw std::length_error("ws_preprocess_scan: too many points");
    ensure_points(&h->d_points, &h->points_cap, (size_t)std::max<int64_t>(n, 1));
    const int stride = point_step_bytes / 4;
    const float *d_xyz = xyz;
    if (!on_device && n > 0)
    {
      const size_t floats = (size_t)n * stride;
      if (floats > h->pre_xyz_cap)
      {
        if (h->d_pre_xyz) WS_CUDA_OK(cudaFree(h->d_pre_xyz));
        h->d_pre_xyz = nullptr; h->pre_xyz_cap = 0;
        WS_CUDA_OK(cudaMalloc(&h->d_pre_xyz, floats * sizeof(float)));
        h->pre_xyz_cap = floats;
      }
      WS_CUDA_OK(cudaMemcpyAsync(h->d_pre_xyz, xyz, floats * sizeof(float), cudaMemcpyHostToDevice, h->stream));
      d_xyz = h->d_pre_xyz;
    }
    h->scan_n = ws_launch_preprocess(h, d_xyz, n, stride, pose_mm, map_resolution, mode);
    if (out_host && h->scan_n > 0)
    {
      WS_CUDA_OK(cudaMemcpyAsync(out_host, h->d_points, (size_t)h->scan_n * sizeof(ws_pt), cudaMemcpyDeviceToHost, h->stream));
      WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    if (n_out) *n_out = h->scan_n;
    return WS_OK;
  });
}

// host: Eigen::Quaterniond(R).toRotationMatrix().cast<float>() (tsdf_mapping.cpp:160), R column-major 3x3 in a 4x4
static void quat_roundtrip(const double P[16], float R[9])
{
  auto m = [&](int r, int c) { return P[c * 4 + r]; };
  double q[4];   // x y z w
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0)
  {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  }
  else
  {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  const double r[9] = { 1.0 - (tyy + tzz), txy + twz, txz - twy,      // column 0
                        txy - twz, 1.0 - (txx + tzz), tyz + twx,      // column 1
                        txz + twy, tyz - twx, 1.0 - (txx + tyy) };    // column 2
  for (int i = 0; i < 9; i++) R[i] = (float)r[i];
}

// include/util/util.h:8-18,52-56 + tsdf_mapping.cpp:77-85 on the host
static void host_convert_pose(const float pose[16], int res, int pos[3], int up[3])
{
  int M[16];
  for (int i = 0; i < 16; i++) M[i] = (int)(pose[i] * (float)WS_MR);
  const unsigned z = (unsigned)WS_MR;
  for (int r = 0; r < 3; r++)
  {
    const int acc = (int)((unsigned)M[8 + r] * z);                  // R_int * (0, 0, MR), translation column 0
    up[r] = acc / WS_MR;
    pos[r] = (int)std::floor(pose[12 + r] / (float)res);
  }
}

static const float *stage_xyz(ws_handle *h, const float *xyz, int64_t n, int stride, int on_device)
{
  if (on_device || n == 0) return xyz;
  const size_t floats = (size_t)n * stride;
  if (floats > h->pre_xyz_cap)
  {
    if (h->d_pre_xyz) WS_CUDA_OK(cudaFree(h->d_pre_xyz));
    h->d_pre_xyz = nullptr; h->pre_xyz_cap = 0;
    WS_CUDA_OK(cudaMalloc(&h->d_pre_xyz, floats * sizeof(float)));
    h->pre_xyz_cap = floats;
  }
  WS_CUDA_OK(cudaMemcpyAsync(h->d_pre_xyz, xyz, floats * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  return h->d_pre_xyz;
}

int ws_voxelgrid_subsample(ws_handle *h, const float *xyz, int64_t n, int32_t point_step_bytes, int32_t on_device, float leaf_m,
                           ws_point *out_mm_host, float *out_xyz_host, int64_t *n_out)
{
  return guarded(h, [&]() {
    if (n < 0 || (n > 0 && !xyz) || !(leaf_m > 0.f) || point_step_bytes < 12 || (point_step_bytes & 3))
      throw std::invalid_argument("ws_voxelgrid_subsample: bad argument");
    if (n > WS_MAX_POINTS) throw std::length_error("ws_voxelgrid_subsample: too many points");
    ensure_points(&h->d_points, &h->points_cap, (size_t)std::max<int64_t>(n, 1));
    const int stride = point_step_bytes / 4;
    const float *d_xyz = stage_xyz(h, xyz, n, stride, on_device);
    float *d_out = nullptr;
    if (out_xyz_host && n > 0)
    {
      if ((size_t)n * 3 > h->vg_xyz_cap)
      {
        if (h->d_vg_xyz) WS_CUDA_OK(cudaFree(h->d_vg_xyz));
        h->d_vg_xyz = nullptr; h->vg_xyz_cap = 0;
        const size_t want = std::max<size_t>((size_t)n * 3, 3u << 17);
        WS_CUDA_OK(cudaMalloc(&h->d_vg_xyz, want * sizeof(float)));
        h->vg_xyz_cap = want;
      }
      d_out = h->d_vg_xyz;
    }
    h->scan_n = ws_launch_voxelgrid(h, d_xyz, n, stride, leaf_m, d_out);
    if (h->scan_n > 0)
    {
      if (out_mm_host)
        WS_CUDA_OK(cudaMemcpyAsync(out_mm_host, h->d_points, (size_t)h->scan_n * sizeof(ws_pt), cudaMemcpyDeviceToHost, h->stream));
      if (out_xyz_host)
        WS_CUDA_OK(cudaMemcpyAsync(out_xyz_host, d_out, (size_t)h->scan_n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
      WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    if (n_out) *n_out = h->scan_n;
    return WS_OK;
  });
}

int ws_update_tsdf_from_ros(ws_handle *h, const float *xyz_m, int64_t n, int32_t point_step_bytes, int32_t on_device,
                            const double pose_m[16], float out_mm_pose[16], int64_t *n_points)
{
  return guarded(h, [&]() {
    if (n < 0 || (n > 0 && !xyz_m) || !pose_m || point_step_bytes < 12 || (point_step_bytes & 3))
      throw std::invalid_argument("ws_update_tsdf_from_ros: bad argument");
    if (n > WS_MAX_POINTS) throw std::length_error("ws_update_tsdf_from_ros: too many points");
    ensure_points(&h->d_points, &h->points_cap, (size_t)std::max<int64_t>(n, 1));
    const int stride = point_step_bytes / 4;
    const float *d_xyz = stage_xyz(h, xyz_m, n, stride, on_device);
    // preprocess_from_ros (tsdf_mapping.cpp:145-163): voxel grid at the map resolution, metres -> millimetres
    h->scan_n = ws_launch_voxelgrid(h, d_xyz, n, stride, (float)h->res / 1000.f, nullptr);
    float mm[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    float R[9];
    quat_roundtrip(pose_m, R);
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < 3; r++) mm[c * 4 + r] = R[c * 3 + r];
    for (int r = 0; r < 3; r++) mm[12 + r] = (float)(pose_m[12 + r] * 1000.0);
    if (out_mm_pose) std::memcpy(out_mm_pose, mm, sizeof(mm));
    if (n_points) *n_points = h->scan_n;
    int pos[3], up[3];
    host_convert_pose(mm, h->res, pos, up);
    h->last_n_points = h->scan_n;
    ws_launch_update(h, h->d_points, (int)h->scan_n, pos, up);                 // tsdf_mapping.cpp:165-173
    return WS_OK;
  });
}

const ws_point *ws_scan_points_device(ws_handle *h, int64_t *n)
{
  if (!h) return nullptr;
  if (n) *n = h->scan_n;
  return reinterpret_cast<const ws_point *>(h->d_points);
}

int ws_get_update_counters(const ws_handle *h, ws_update_counters *out)
{
  if (!h || !out) return WS_ERR_INVALID;
  const UpdateCounters &c = h->last_counters;
  out->n_points = h->last_n_points;
  out->n_candidates = (int64_t)c.n_candidates;
  out->n_touched = (int64_t)c.n_touched;
  out->n_written = (int64_t)c.n_written;
  out->n_touched_bricks = c.n_touched_bricks;
  out->n_parked = c.n_parked;
  out->n_rounds = c.rounds;
  out->n_list = c.n_list;
  out->n_record_chunks = c.n_chunks;
  for (int i = 0; i < 3; i++) out->replay_phase_ns[i] = (int64_t)(c.t_phase[i + 1] - c.t_phase[i]);
  return WS_OK;
}

int ws_reg_prepare(ws_handle *h, const ws_point *points, int64_t n)
{
  return guarded(h, [&]() {
    if (n < 0 || (n > 0 && !points)) throw std::invalid_argument("ws_reg_prepare: bad argument");
    if (n > WS_MAX_POINTS) throw std::length_error("ws_reg_prepare: too many points");
    ensure_points(&h->d_reg_points, &h->reg_points_cap, (size_t)n);
    h->d_reg_points_last = nullptr;
    if (n > 0)
      WS_CUDA_OK(cudaMemcpyAsync(h->d_reg_points, points, (size_t)n * sizeof(ws_pt), cudaMemcpyHostToDevice, h->stream));
    h->reg_n = (int)n;
    return WS_OK;
  });
}

int ws_reg_prepare_device(ws_handle *h, const ws_point *device_points, int64_t n)
{
  return guarded(h, [&]() {
    if (n < 0 || (n > 0 && !device_points)) throw std::invalid_argument("ws_reg_prepare_device: bad argument");
    if (n > WS_MAX_POINTS) throw std::length_error("ws_reg_prepare_device: too many points");
    ensure_points(&h->d_reg_points, &h->reg_points_cap, (size_t)n);
    h->d_reg_points_last = nullptr;
    if (n > 0)
      WS_CUDA_OK(cudaMemcpyAsync(h->d_reg_points, device_points, (size_t)n * sizeof(ws_pt), cudaMemcpyDeviceToDevice, h->stream));
    h->reg_n = (int)n;
    return WS_OK;
  });
}

int ws_reg_step(ws_handle *h, const float T[16], int32_t map_resolution, int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt)
{
  return guarded(h, [&]() {
    if (!T || !H || !g || !err || !cnt || map_resolution < 1) throw std::invalid_argument("ws_reg_step: bad argument");
    ws_launch_reg_reset(h, T, 0.f);
    ws_launch_reg_iteration(h, h->reg_n, map_resolution, 0, 0.f, 0.f);
    read_acc(h);
    sums_to_hg(h->h_acc->sums, H, g, err, cnt);
    return WS_OK;
  });
}

int ws_register_cloud(ws_handle *h, ws_point *cloud, int64_t n, const float pretransform[16], int32_t max_iterations,
                      float it_weight_gradient, float epsilon, int32_t map_resolution, int32_t flags,
                      float out_transform[16], int32_t *iterations)
{
  return guarded(h, [&]() {
    if (!pretransform || !out_transform || map_resolution < 1 || max_iterations < 0)
      throw std::invalid_argument("ws_register_cloud: bad argument");
    if (cloud)
    {
      if (n < 0) throw std::invalid_argument("ws_register_cloud: bad point count");
      if (n > WS_MAX_POINTS) throw std::length_error("ws_register_cloud: too many points");
      ensure_points(&h->d_reg_points, &h->reg_points_cap, (size_t)n);
      h->d_reg_points_last = nullptr;
    h->d_reg_points_last = nullptr;
      if (n > 0)
        WS_CUDA_OK(cudaMemcpyAsync(h->d_reg_points, cloud, (size_t)n * sizeof(ws_pt), cudaMemcpyHostToDevice, h->stream));
      h->reg_n = (int)n;
    }
    const int np = h->reg_n;
    ws_launch_reg_reset(h, pretransform, 0.f);
    int it_done = 0;
    h->last_reg_host = (flags & WS_REG_HOST_SOLVE) != 0;
    h->host_trace.clear();
    if (!h->last_reg_host)
    {
      // every Gauss-Newton iteration inside one persistent cooperative kernel (registration.cu)
      if (max_iterations > 0) ws_launch_reg_loop(h, np, map_resolution, max_iterations, it_weight_gradient, epsilon);
      ws_launch_transform_cloud(h, h->d_reg_points, np);
      read_acc(h);
      it_done = (int)h->h_acc->iterations;
      std::memcpy(out_transform, h->h_acc->T, 16 * sizeof(float));
      if (h->h_acc->finished == 2u)
        throw std::logic_error("register_cloud: a peer rank did not deliver its Gauss-Newton sums in time");
    }
    else
    {
      // the reference's structure: sums to the host, FP64 solve there (tsdf_registration.cpp:55-92)
      float T[16];
      std::memcpy(T, pretransform, sizeof(T));
      float alpha = 0.f;
      float prev[4] = { 0, 0, 0, 0 };
      bool finished = false;
      for (int i = 0; i < max_iterations && !finished; i++)
      {
        ws_launch_reg_iteration(h, np, map_resolution, 0, 0.f, 0.f);
        read_acc(h);
        int64_t H[36], g[6]; int32_t e, c;
        sums_to_hg(h->h_acc->sums, H, g, &e, &c);
        for (int k = 0; k < WS_NSUM; k++) h->host_trace.push_back((i64)h->h_acc->sums[k]);
        const float errv = [&]() { double xi[6]; ws_host_solve((const i64 *)H, (const i64 *)g, e, c, alpha, T, xi); return (float)e / c; }();
        alpha += it_weight_gradient;
        if (fabs(errv - prev[2]) < epsilon && fabs(errv - prev[0]) < epsilon) finished = true;
        prev[0] = prev[1]; prev[1] = prev[2]; prev[2] = prev[3]; prev[3] = errv;
        it_done = i + 1;
        ws_launch_reg_reset(h, T, alpha);
      }
      ws_launch_transform_cloud(h, h->d_reg_points, np);
      std::memcpy(out_transform, T, sizeof(T));
    }
    if (cloud && np > 0)
      WS_CUDA_OK(cudaMemcpyAsync(cloud, h->d_reg_points, (size_t)np * sizeof(ws_pt), cudaMemcpyDeviceToHost, h->stream));
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    WS_CUDA_OK(cudaGetLastError());
    h->last_reg_iterations = it_done;
    if (iterations) *iterations = it_done;
    return WS_OK;
  });
}

// ---- peer mailboxes of the fused multi-GPU registration (registration.cu) ------------------------
int ws_peer_export(ws_handle *h, uint8_t handle[WS_IPC_HANDLE_BYTES])
{
  static_assert(sizeof(cudaIpcMemHandle_t) == WS_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  return guarded(h, [&]() {
    if (!handle) throw std::invalid_argument("ws_peer_export: null handle buffer");
    ensure_mailbox(h);
    cudaIpcMemHandle_t ipc;
    WS_CUDA_OK(cudaIpcGetMemHandle(&ipc, h->d_mail));
    std::memcpy(handle, &ipc, sizeof(ipc));
    return WS_OK;
  });
}

int ws_peer_attach_ipc(ws_handle *h, const uint8_t *handles, int32_t world)
{
  return guarded(h, [&]() {
    if (!handles || world != h->world || world > WS_MAX_PEERS) throw std::invalid_argument("ws_peer_attach_ipc: bad world");
    ensure_mailbox(h);
    detach_peers(h);
    for (int p = 0; p < world; p++)
    {
      if (p == h->rank) { h->peer_mail[p] = h->d_mail; continue; }
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, handles + (size_t)p * WS_IPC_HANDLE_BYTES, sizeof(ipc));
      void *ptr = nullptr;
      WS_CUDA_OK(cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
      h->peer_mail[p] = ptr; h->peer_ipc[p] = true;
    }
    h->peers_attached = true;
    return WS_OK;
  });
}

void *ws_peer_local_ptr(ws_handle *h)
{
  if (!h) return nullptr;
  void *out = nullptr;
  guarded(h, [&]() { ensure_mailbox(h); out = h->d_mail; return WS_OK; });
  return out;
}

int ws_peer_attach_ptrs(ws_handle *h, void *const *mailboxes, const int32_t *devices, int32_t world)
{
  return guarded(h, [&]() {
    if (!mailboxes || world != h->world || world > WS_MAX_PEERS) throw std::invalid_argument("ws_peer_attach_ptrs: bad world");
    ensure_mailbox(h);
    detach_peers(h);
    for (int p = 0; p < world; p++)
    {
      if (p == h->rank) { h->peer_mail[p] = h->d_mail; continue; }
      if (!mailboxes[p]) throw std::invalid_argument("ws_peer_attach_ptrs: null mailbox");
      if (devices && devices[p] != h->device)
      {
        int can = 0;
        WS_CUDA_OK(cudaDeviceCanAccessPeer(&can, h->device, devices[p]));
        if (!can) throw std::runtime_error("ws_peer_attach_ptrs: no peer access between the devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[p], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) WS_CUDA_OK(e);
        cudaGetLastError();
      }
      h->peer_mail[p] = mailboxes[p];
    }
    h->peers_attached = true;
    return WS_OK;
  });
}

int ws_peer_set_timeout(ws_handle *h, double seconds)
{
  if (!h || !(seconds > 0.0)) return WS_ERR_INVALID;
  h->peer_timeout_ns = (unsigned long long)(seconds * 1e9);
  return WS_OK;
}

// ---- per-scan pipeline ------------------------------------------------------------------------------
// register_cloud -> pose update -> update_tsdf as ONE stream of kernels per scan (the calls App::cloud_callback
// makes, src/warpsense/app.cpp:65-112, in the order of BASELINE configs[2]): the registration leaves its
// transform on the device, pose_kernel turns it and the prior pose into the scanner voxel and the up vector
// there, and update_tsdf's kernels read them from device memory.  ws_track_submit only ENQUEUES (the scan is
// copied from host memory on a second stream into one of two device buffers); ws_track_wait collects the
// transform, the pose and the work counters.  With two scans in flight the host never idles the GPU.
static void track_slot_init(ws_handle *h, ws_handle::TrackSlot &t)
{
  if (t.h_acc) return;
  WS_CUDA_OK(cudaMallocHost(&t.h_acc, sizeof(RegAccum)));
  WS_CUDA_OK(cudaMallocHost(&t.h_pose, sizeof(PoseDev)));
  WS_CUDA_OK(cudaMallocHost(&t.h_ctr, sizeof(UpdateCounters)));
  std::memset(t.h_acc, 0, sizeof(RegAccum)); std::memset(t.h_pose, 0, sizeof(PoseDev)); std::memset(t.h_ctr, 0, sizeof(UpdateCounters));
  WS_CUDA_OK(cudaEventCreateWithFlags(&t.copied, cudaEventDisableTiming));
  WS_CUDA_OK(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
  if (!h->copy_stream) WS_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
}

// complete a tracked scan on the host side (wait, regrow the record if needed, error checks); errors are kept
// for the scan's ws_track_wait
static void track_finish_slot(ws_handle *h, ws_handle::TrackSlot &t)
{
  try { ws_update_finish(h, t.h_ctr, t.h_pose, t.done); }
  catch (const std::exception &e) { t.error = e.what(); }
  t.finished = true;
  // room to spare: the record used less than 60 % of its chunks and did not have to grow
  const size_t used = t.h_ctr->n_chunks;
  if (t.h_ctr->rec_overflow == 0u && h->record_regrows == t.regrows_at_submit && used * 10 < h->rec_cap_chunks * 6) h->rec_headroom_ok = true;
  else if (h->record_regrows != t.regrows_at_submit || used * 10 >= h->rec_cap_chunks * 9) h->rec_headroom_ok = false;
}

int ws_track_submit(ws_handle *h, const ws_point *points, int64_t n, int32_t on_device, const float *prior_pose,
                    const float *pretransform, int32_t max_iterations, float it_weight_gradient, float epsilon,
                    int32_t map_resolution, int32_t flags, int32_t *ticket)
{
  return guarded(h, [&]() {
    if (!ticket || map_resolution < 1 || max_iterations < 0 || n < 0 || (n > 0 && !points))
      throw std::invalid_argument("ws_track_submit: bad argument");
    if (n > WS_MAX_POINTS) throw std::length_error("ws_track_submit: too many points");
    if (map_resolution != h->res) throw std::invalid_argument("ws_track_submit: map_resolution differs from the map's");
    if (!prior_pose && !h->track_has_pose) throw std::logic_error("ws_track_submit: no previous pose on the device to chain from");
    ws_handle::TrackSlot &t = h->track[h->track_next];
    if (t.busy) throw std::logic_error("ws_track_submit: two scans already in flight (call ws_track_wait)");
    // A scan whose candidate record overflows is finished by the host (regrow + regenerate); nothing may be
    // queued behind it.  Until a scan has shown that the record has room to spare, the scan in flight is
    // completed here before the next one is enqueued.
    ws_handle::TrackSlot &prev = h->track[h->track_next ^ 1];
    if (prev.busy && !prev.finished && !h->rec_headroom_ok) track_finish_slot(h, prev);
    track_slot_init(h, t);
    ensure_points(&t.d_pts, &t.cap, (size_t)std::max<int64_t>(n, 1));
    if (n > 0)
    {
      if (on_device)
        WS_CUDA_OK(cudaMemcpyAsync(t.d_pts, points, (size_t)n * sizeof(ws_pt), cudaMemcpyDeviceToDevice, h->stream));
      else
      {
        // host -> device on the copy stream: overlaps the kernels of the scan before
        WS_CUDA_OK(cudaMemcpyAsync(t.d_pts, points, (size_t)n * sizeof(ws_pt), cudaMemcpyHostToDevice, h->copy_stream));
        WS_CUDA_OK(cudaEventRecord(t.copied, h->copy_stream));
        WS_CUDA_OK(cudaStreamWaitEvent(h->stream, t.copied, 0));
      }
    }
    // the registration works on (and transforms in place) the slot's buffer
    h->d_reg_points_alias = t.d_pts;
    h->reg_n = (int)n;
    h->last_reg_host = false;
    h->host_trace.clear();
    const float I16[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    // registration -> pose -> update_tsdf, whose set-up kernel transforms the cloud by the registration result first
    // (registration.cpp:164-174); results read back at the end.  (The pose stays a kernel of its own: computed in
    // the registration kernel's epilogue it cost that kernel 12 % -- its loop is sensitive to what follows it.)
    ws_launch_reg_reset(h, pretransform ? pretransform : I16, 0.f);
    if (max_iterations > 0 && n > 0) ws_launch_reg_loop(h, (int)n, map_resolution, max_iterations, it_weight_gradient, epsilon);
    ws_launch_pose(h, h->d_acc->T, prior_pose, (flags & WS_TRACK_REFERENCE_POSE) ? 1 : 0);
    h->track_has_pose = true;
    ws_update_enqueue(h, t.d_pts, (int)n, nullptr, nullptr, true, t.h_ctr, t.h_pose, h->d_acc->T);
    WS_CUDA_OK(cudaMemcpyAsync(t.h_acc, h->d_acc, sizeof(RegAccum), cudaMemcpyDeviceToHost, h->stream));
    WS_CUDA_OK(cudaEventRecord(t.done, h->stream));
    h->d_reg_points_alias = nullptr;
    t.busy = true; t.finished = false; t.error.clear(); t.n = n; t.regrows_at_submit = h->record_regrows;
    h->track_in_flight++;
    *ticket = h->track_next;
    h->track_next ^= 1;
    return WS_OK;
  });
}

int ws_track_wait(ws_handle *h, int32_t ticket, float out_transform[16], float out_pose[16], int32_t *iterations)
{
  return guarded(h, [&]() {
    if (ticket < 0 || ticket > 1 || !h->track[ticket].busy) throw std::invalid_argument("ws_track_wait: no such scan in flight");
    ws_handle::TrackSlot &t = h->track[ticket];
    struct Release { ws_handle *h; ws_handle::TrackSlot &t; ~Release() { t.busy = false; h->track_in_flight--; } } release{ h, t };
    h->last_n_points = t.n;
    if (!t.finished) track_finish_slot(h, t);
    h->last_counters = *t.h_ctr;
    if (!t.error.empty()) throw std::runtime_error(t.error);
    if (t.h_acc->finished == 2u)
      throw std::logic_error("track_scan: a peer rank did not deliver its Gauss-Newton sums in time");
    if (out_transform) std::memcpy(out_transform, t.h_acc->T, 16 * sizeof(float));
    if (out_pose) std::memcpy(out_pose, static_cast<const PoseDev *>(t.h_pose)->pose, 16 * sizeof(float));
    h->last_reg_iterations = (int)t.h_acc->iterations;
    if (iterations) *iterations = h->last_reg_iterations;
    // the registered cloud of this scan (ws_reg_points_device)
    h->d_reg_points_last = t.d_pts;
    return WS_OK;
  });
}

int ws_track_scan_ex(ws_handle *h, const ws_point *points, int64_t n, int32_t on_device, const float prior_pose[16],
                     const float *pretransform, int32_t max_iterations, float it_weight_gradient, float epsilon,
                     int32_t map_resolution, int32_t flags, float out_transform[16], float out_pose[16], int32_t *iterations)
{
  if (!prior_pose || !out_transform) return WS_ERR_INVALID;
  int32_t ticket = -1;
  int rc = ws_track_submit(h, points, n, on_device, prior_pose, pretransform, max_iterations, it_weight_gradient, epsilon,
                           map_resolution, flags, &ticket);
  if (rc != WS_OK) return rc;
  return ws_track_wait(h, ticket, out_transform, out_pose, iterations);
}

int ws_track_scan(ws_handle *h, const ws_point *points, int64_t n, int32_t on_device, const float prior_pose[16],
                  int32_t max_iterations, float it_weight_gradient, float epsilon, int32_t map_resolution,
                  float out_transform[16], float out_pose[16], int32_t *iterations)
{
  return ws_track_scan_ex(h, points, n, on_device, prior_pose, nullptr, max_iterations, it_weight_gradient, epsilon,
                          map_resolution, 0, out_transform, out_pose, iterations);
}

int ws_reg_get_trace(ws_handle *h, int64_t *out, int32_t max_iterations)
{
  if (!h || !out) return WS_ERR_INVALID;
  int n = std::min(h->last_reg_iterations, (int)max_iterations);
  int rc = guarded(h, [&]() {
    if (h->last_reg_host)
    {
      n = std::min<int>(n, (int)(h->host_trace.size() / WS_NSUM));
      std::memcpy(out, h->host_trace.data(), (size_t)n * WS_NSUM * sizeof(int64_t));
    }
    else
    {
      n = std::min(n, h->trace_cap);
      if (n > 0)
      {
        WS_CUDA_OK(cudaMemcpyAsync(out, h->d_trace, (size_t)n * WS_NSUM * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        WS_CUDA_OK(cudaStreamSynchronize(h->stream));
      }
    }
    return WS_OK;
  });
  return rc == WS_OK ? n : rc;
}

const ws_point *ws_reg_points_device(ws_handle *h, int64_t *n)
{
  if (!h) return nullptr;
  if (n) *n = h->reg_n;
  return reinterpret_cast<const ws_point *>(h->d_reg_points_last ? h->d_reg_points_last : h->d_reg_points);
}

int ws_reg_begin(ws_handle *h, const float pretransform[16])
{
  return guarded(h, [&]() {
    if (!pretransform) throw std::invalid_argument("ws_reg_begin: null argument");
    ws_launch_reg_reset(h, pretransform, 0.f);
    h->last_reg_host = false;
    return WS_OK;
  });
}

int ws_reg_accumulate(ws_handle *h, int32_t map_resolution)
{
  return guarded(h, [&]() {
    if (map_resolution < 1) throw std::invalid_argument("ws_reg_accumulate: bad resolution");
    ws_launch_reg_iteration(h, h->reg_n, map_resolution, 0, 0.f, 0.f);
    return WS_OK;
  });
}

void *ws_reg_sums_device(ws_handle *h) { return h ? (void *)h->d_acc->sums : nullptr; }

int ws_reg_sums_get(ws_handle *h, int64_t sums[29])
{
  return guarded(h, [&]() {
    if (!sums) throw std::invalid_argument("ws_reg_sums_get: null argument");
    WS_CUDA_OK(cudaMemcpyAsync(sums, h->d_acc->sums, WS_NSUM * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    return WS_OK;
  });
}

int ws_reg_sums_set(ws_handle *h, const int64_t sums[29])
{
  return guarded(h, [&]() {
    if (!sums) throw std::invalid_argument("ws_reg_sums_set: null argument");
    WS_CUDA_OK(cudaMemcpyAsync(h->d_acc->sums, sums, WS_NSUM * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    return WS_OK;
  });
}

int ws_slab_layout(int32_t size_x, int32_t rank, int32_t world, int32_t *own_lo, int32_t *own_hi,
                   int32_t *resident_cols, int32_t cap)
{
  if (size_x < 1 || world < 1 || rank < 0 || rank >= world || !own_lo || !own_hi) return WS_ERR_INVALID;
  std::vector<int> cols;
  int lo = 0, hi = 0;
  if (!slab_layout(size_x, rank, world, &lo, &hi, cols)) return WS_ERR_INVALID;
  *own_lo = lo; *own_hi = hi;
  if (resident_cols)
    for (size_t i = 0; i < cols.size() && (int)i < cap; i++) resident_cols[i] = cols[i];
  return (int)cols.size();
}

int ws_shard_layout(int32_t size_x, int32_t rank, int32_t world, int32_t stripe_cols, uint8_t *owned, uint8_t *resident, int32_t cap)
{
  if (size_x < 1 || world < 1 || rank < 0 || rank >= world) return WS_ERR_INVALID;
  std::vector<char> own;
  std::vector<int> cols;
  if (!shard_layout(size_x, rank, world, stripe_cols, own, cols)) return WS_ERR_INVALID;
  const int nbx = (int)own.size();
  for (int c = 0; c < nbx && c < cap; c++)
  {
    if (owned) owned[c] = (uint8_t)own[(size_t)c];
    if (resident) resident[c] = 0;
  }
  if (resident)
    for (int c : cols) if (c < cap) resident[c] = 1;
  return nbx;
}

int ws_shard_columns(const ws_handle *h, uint8_t *owned, uint8_t *resident, int32_t cap)
{
  if (!h) return WS_ERR_INVALID;
  const int nbx = h->g.nb[0];
  for (int c = 0; c < nbx && c < cap; c++)
  {
    if (owned) owned[c] = h->g.xown[c];
    if (resident) resident[c] = h->g.xslot[c] >= 0;
  }
  return nbx;
}

int ws_reg_solve(ws_handle *h, float it_weight_gradient, float epsilon)
{
  return guarded(h, [&]() { ws_launch_reg_solve(h, it_weight_gradient, epsilon); return WS_OK; });
}

int ws_reg_peek(ws_handle *h, int32_t *iterations, int32_t *finished)
{
  return guarded(h, [&]() {
    read_acc(h);
    if (iterations) *iterations = (int32_t)h->h_acc->iterations;
    if (finished) *finished = (int32_t)h->h_acc->finished;
    return WS_OK;
  });
}

int ws_reg_finish(ws_handle *h, float out_transform[16], int32_t *iterations, int32_t *finished)
{
  return guarded(h, [&]() {
    ws_launch_transform_cloud(h, h->d_reg_points, h->reg_n);
    read_acc(h);
    if (out_transform) std::memcpy(out_transform, h->h_acc->T, 16 * sizeof(float));
    if (iterations) *iterations = (int32_t)h->h_acc->iterations;
    if (finished) *finished = (int32_t)h->h_acc->finished;
    h->last_reg_iterations = (int)h->h_acc->iterations;
    return WS_OK;
  });
}

int ws_test_reduce(ws_handle *h, const int64_t *jacobis6, const int32_t *values, int64_t n, int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt)
{
  return guarded(h, [&]() {
    if (n < 0 || (n > 0 && (!jacobis6 || !values)) || !H || !g || !err || !cnt) throw std::invalid_argument("ws_test_reduce: bad argument");
    i64 *d_j = nullptr; int *d_v = nullptr;
    const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    ws_launch_reg_reset(h, I, 0.f);
    WS_CUDA_OK(cudaMalloc(&d_j, std::max<size_t>(1, (size_t)n * 6 * sizeof(i64))));
    WS_CUDA_OK(cudaMalloc(&d_v, std::max<size_t>(1, (size_t)n * sizeof(int))));
    if (n > 0)
    {
      WS_CUDA_OK(cudaMemcpyAsync(d_j, jacobis6, (size_t)n * 6 * sizeof(i64), cudaMemcpyHostToDevice, h->stream));
      WS_CUDA_OK(cudaMemcpyAsync(d_v, values, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    ws_launch_test_reduce(h, d_j, d_v, (int)n);
    read_acc(h);
    cudaFree(d_j); cudaFree(d_v);
    sums_to_hg(h->h_acc->sums, H, g, err, cnt);
    return WS_OK;
  });
}

int ws_shift(ws_handle *h, const int32_t new_pos[3])
{
  return guarded(h, [&]() {
    if (!new_pos) throw std::invalid_argument("ws_shift: null argument");
    GridDesc &g = h->g;
    int diff[3];
    for (int a = 0; a < 3; a++)
    {
      diff[a] = new_pos[a] - g.pos[a];
      if (std::abs(diff[a]) > g.size[a]) throw std::invalid_argument("ws_shift: shift larger than the map (hdf5_local_map.cpp:63-65)");
    }
    for (int axis = 0; axis < 3; axis++)                                    // hdf5_local_map.cpp:68-117
    {
      if (diff[axis] == 0) continue;
      int start[3], end[3];
      for (int a = 0; a < 3; a++) { start[a] = g.pos[a] - g.half[a]; end[a] = g.pos[a] + g.half[a]; }
      if (diff[axis] > 0) end[axis] = start[axis] + diff[axis] - 1;
      else start[axis] = end[axis] + diff[axis] + 1;
      save_load_area(h, start, end, true);

      g.pos[axis] += diff[axis];
      g.offset[axis] = (g.offset[axis] + diff[axis] + g.size[axis]) % g.size[axis];

      for (int a = 0; a < 3; a++) { start[a] = g.pos[a] - g.half[a]; end[a] = g.pos[a] + g.half[a]; }
      if (diff[axis] > 0) start[axis] = end[axis] - (diff[axis] - 1);
      else end[axis] = start[axis] - diff[axis] - 1;
      save_load_area(h, start, end, false);
    }
    return WS_OK;
  });
}

int ws_write_back(ws_handle *h)
{
  return guarded(h, [&]() {
    int start[3], end[3];
    for (int a = 0; a < 3; a++) { start[a] = h->g.pos[a] - h->g.half[a]; end[a] = h->g.pos[a] + h->g.half[a]; }
    save_load_area(h, start, end, true);
    return WS_OK;
  });
}

int ws_map_reload(ws_handle *h)
{
  return guarded(h, [&]() {
    int start[3], end[3];
    for (int a = 0; a < 3; a++) { start[a] = h->g.pos[a] - h->g.half[a]; end[a] = h->g.pos[a] + h->g.half[a]; }
    save_load_area(h, start, end, false);
    return WS_OK;
  });
}

int64_t ws_store_num_chunks(const ws_handle *h)
{
  if (!h) return 0;
  int64_t n = (int64_t)h->store.size();
  for (const auto &kv : h->spilled) if (!h->store.count(kv.first)) n++;
  return n;
}

int ws_store_chunk_list(const ws_handle *h, int32_t *xyz, int64_t cap)
{
  if (!h || !xyz) return WS_ERR_INVALID;
  int64_t k = 0;
  for (unsigned long long key : ws_store_keys(const_cast<ws_handle *>(h)))
  {
    if (k >= cap) break;
    int x, y, z;
    chunk_unkey(key, x, y, z);
    xyz[3 * k] = x; xyz[3 * k + 1] = y; xyz[3 * k + 2] = z;
    k++;
  }
  return (int)k;
}

int ws_store_get_chunk(const ws_handle *h_, int32_t cx, int32_t cy, int32_t cz, uint32_t *out)
{
  ws_handle *h = const_cast<ws_handle *>(h_);
  if (!h || !out) return WS_ERR_INVALID;
  const u64 key = chunk_key(cx, cy, cz);
  if (!h->store.count(key) && !h->spilled.count(key)) return WS_ERR_INVALID;
  try { std::memcpy(out, activate_chunk(h, cx, cy, cz), CHUNK_ENTRIES * sizeof(uint32_t)); }
  catch (const std::exception &e) { h->last_error = e.what(); return WS_ERR_STATE; }
  return WS_OK;
}

int ws_store_set_chunk(ws_handle *h, int32_t cx, int32_t cy, int32_t cz, const uint32_t *data)
{
  if (!h || !data) return WS_ERR_INVALID;
  try { std::memcpy(activate_chunk(h, cx, cy, cz), data, CHUNK_ENTRIES * sizeof(uint32_t)); }
  catch (const std::exception &e) { h->last_error = e.what(); return WS_ERR_STATE; }
  return WS_OK;
}

int ws_store_configure(ws_handle *h, int64_t max_chunks_in_memory)
{
  if (!h || max_chunks_in_memory < 1) return WS_ERR_INVALID;
  h->store_max_chunks = (size_t)max_chunks_in_memory;
  try { while (h->store.size() > h->store_max_chunks && !h->lru.empty()) store_evict_one(h); }
  catch (const std::exception &e) { h->last_error = e.what(); return WS_ERR_STATE; }
  return WS_OK;
}

int64_t ws_store_evictions(const ws_handle *h) { return h ? h->store_evictions : 0; }

int64_t ws_launch_count(const ws_handle *h) { return h ? h->launches : 0; }

int ws_profile_enable(ws_handle *h, int32_t on)
{
  if (!h) return WS_ERR_INVALID;
  h->profile = on == 2 ? 2 : (on != 0 ? 1 : 0);
  return WS_OK;
}

int ws_profile_reset(ws_handle *h)
{
  return guarded(h, [&]() {
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    h->timers_used = 0;
    return WS_OK;
  });
}

int64_t ws_profile_timeline(ws_handle *h, double *out, int64_t cap_ranges)
{
  // ranges of the LAST update_tsdf span (and the registration before it), as (kind, start_ms, stop_ms) relative to the span's start
  int64_t n = 0;
  const int rc = guarded(h, [&]() {
    WS_CUDA_OK(cudaDeviceSynchronize());
    const size_t used = std::min(h->timers_used, h->timers.size());
    long span = -1, first = -1;
    for (size_t i = 0; i < used; i++)
      if (h->timer_kind[i] == WS_TIMER_UPDATE) span = (long)i;
    if (span < 0) return WS_OK;
    // ... preceded by the ranges since the end of the update before (the registration of this scan): negative offsets
    first = span;
    while (first > 0 && h->timer_kind[first - 1] == WS_TIMER_REG) first--;
    for (size_t i = (size_t)first; i < used && n < cap_ranges; i++)
    {
      float a = 0.f, b = 0.f;
      if (cudaEventElapsedTime(&a, h->timers[span].start, h->timers[i].start) != cudaSuccess) continue;
      if (cudaEventElapsedTime(&b, h->timers[span].start, h->timers[i].stop) != cudaSuccess) continue;
      out[3 * n] = (double)h->timer_kind[i]; out[3 * n + 1] = a; out[3 * n + 2] = b;
      n++;
    }
    return WS_OK;
  });
  return rc == WS_OK ? n : -1;
}

int ws_profile_get(ws_handle *h, int32_t kind, double *total_ms, int64_t *launches)
{
  return guarded(h, [&]() {
    WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    int64_t cnt = 0;
    const size_t used = std::min(h->timers_used, h->timers.size());
    for (size_t i = 0; i < used; i++)
    {
      if (h->timer_kind[i] != kind) continue;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->timers[i].start, h->timers[i].stop) == cudaSuccess) { tot += ms; cnt++; }
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = cnt;
    return WS_OK;
  });
}

}  // extern "C"
