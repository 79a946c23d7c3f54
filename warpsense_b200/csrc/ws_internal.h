// ws_internal.h -- host-side state of one map handle and the launcher prototypes (not part of the ABI).
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <cstdio>
#include <list>
#include <unordered_map>
#include <vector>
#include "ws_common.cuh"
#include "march_math.cuh"

#define WS_MAX_XIV 6         // resident x intervals of a rank == march-step ranges per ray

struct UpdateParams
{
  int tau, max_weight, res, weight_epsilon, dz_per_distance;
  int pos_mm[3];       // scanner_pos * map_resolution (update_tsdf.cpp:410)
  i64 up[3];
  FastDiv div_res;     // / map_resolution
  FastDiv div_weps;    // / (tau - weight_epsilon)
  int half_res;        // map_resolution / 2
  int n_points;
  int far_len;         // march lengths >= far_len can meet an interpolated winner: their candidates are recorded
  int far_block;       // first step block (LS_BLOCK steps) that holds a far-field step
  int near_split_block; // near-field blocks from this one on are marched beside the far field (item table 3), the ones before beside the surface phase (table 1)
  // fast path of the march (march_math.cuh)
  FastDiv div_half;    // / (map_resolution / 2)
  FastDiv32 div_res32; // / map_resolution, 32-bit magic
  int tau_sq;          // tau * tau (tau <= 32767)
  int zero_weight_max; // march steps with value <= this have weight 0 and are skipped
  int coord_lim;       // |coordinate| bound of the 32-bit magic divisions, 0 = fast path off
  int lo[3];           // lowest in-bounds voxel per axis (pos - size/2)
  int ext[3];          // size - 1
  int ringc[3];        // ring coordinate of voxel lo: (offset - size/2) mod size
  // multi-GPU slabs: x intervals (mm) outside which no march step can reach a resident column; 0 = no culling
  int n_xiv;
  int xiv_lo[WS_MAX_XIV], xiv_hi[WS_MAX_XIV];
};

// scanner pose handed from the registration to update_tsdf on the device (fused per-scan pipeline)
struct PoseDev
{
  int pos_mm[3];       // scanner voxel * map_resolution
  int coord_lim;       // UpdateParams::coord_lim after the |pos_mm| check
  long long up[3];
  int pos_vox[3];      // scanner voxel (tsdf_mapping.cpp:77-85)
  int pad;
  float pose[16];      // X * prior, column-major
};

// what the device needs to turn a registration result X into the scanner pose of update_tsdf
struct PoseArgs
{
  PoseDev *out;            // nullptr: no pose wanted
  float prior[16];         // prior pose (column-major), unless chain
  int chain;               // 1: the prior is the pose left in `out` by the scan before
  int compose_reference;   // 1: App::update_pose_estimate (R = X.R * R, t += X.t); 0: the full product X * prior
  int res;
  int coord_lim;
};

#ifdef __CUDACC__
// src/warpsense/tsdf_mapping.cpp:77-85 + include/util/util.h:52-56: new pose from the registration result X and the
// prior pose, scanner voxel = floor(t / res), up = third column of to_int_mat(pose).  One thread.
__device__ inline void pose_compute(const float *X, const PoseArgs &pa)
{
  PoseDev *out = pa.out;
  float prior[16], pose[16];
  for (int i = 0; i < 16; i++) prior[i] = pa.chain ? out->pose[i] : pa.prior[i];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
    {
      float acc = X[0 * 4 + r] * prior[c * 4 + 0];
      acc = acc + X[1 * 4 + r] * prior[c * 4 + 1];
      acc = acc + X[2 * 4 + r] * prior[c * 4 + 2];
      if (!pa.compose_reference) acc = acc + X[3 * 4 + r] * prior[c * 4 + 3];
      pose[c * 4 + r] = acc;
    }
  if (pa.compose_reference)
  {
    for (int r = 0; r < 3; r++) pose[12 + r] = prior[12 + r] + X[12 + r];
    for (int c = 0; c < 4; c++) pose[c * 4 + 3] = prior[c * 4 + 3];
  }
  bool ok = true;
  for (int a = 0; a < 3; a++)
  {
    const int vox = (int)floorf(pose[12 + a] / (float)pa.res);
    out->pos_mm[a] = (int)((unsigned)vox * (unsigned)pa.res);
    out->up[a] = (long long)(int)(pose[8 + a] * (float)WS_MR);
    const long long ap = out->pos_mm[a] < 0 ? -(long long)out->pos_mm[a] : (long long)out->pos_mm[a];
    if (ap >= (long long)pa.coord_lim) ok = false;
    out->pos_vox[a] = vox;
  }
  out->coord_lim = ok ? pa.coord_lim : 0;
  for (int i = 0; i < 16; i++) out->pose[i] = pose[i];
}
#endif

// one recorded candidate: its key and the voxel address (record) or the pending slot (replay list)
struct Rec
{
  u64 key;
  u64 ref;
};
#define WS_REC_CHUNK 64    // records per chunk; a chunk belongs to one warp of the march kernel
#define WS_REC_SPAN 16     // chunks a warp reserves per allocation

// device-side work counters / status of one update (kept in one small device buffer)
struct UpdateCounters
{
  unsigned long long n_candidates;
  unsigned long long n_touched;
  unsigned long long n_written;
  unsigned n_touched_bricks; // bricks touched by the scan (surface or free-space phase)
  unsigned n_pending;        // voxels whose winner is an interpolated candidate below tau
  unsigned n_surf_bricks;    // bricks touched by the surface phase
  unsigned pending_overflow;
  unsigned error;            // bit 0: seq field overflow (too many march/fan steps); bit 1: replay list overflow
  unsigned rounds;
  unsigned n_parked;         // voxels parked by the merge pass (before any replay round)
  unsigned n_work;
  unsigned n_items[4];       // (ray group, step block) work items of the lockstep march: surface / free space near (first part) / free space far / free space near (second part)
  unsigned n_general;        // rays outside the 32-bit fast path: marched by the literal-arithmetic kernel
  unsigned gen_counter;      // ... and its dynamic fetch counter
  unsigned pad0[12];
  unsigned item_counter[4];  // dynamic item fetch of the persistent march warps, per item table (own 128-byte line)
  unsigned pad1[28];
  unsigned n_chunks;         // record chunks handed out (own 128-byte line)
  unsigned pad2[31];
  unsigned rec_overflow;     // the record did not fit its buffer
  unsigned n_list;           // entries of the replay list (round-1 survivors + free-space hits on parked voxels)
  unsigned n_active[2];      // still-pending slots, ping-pong between replay rounds
  unsigned long long t_phase[4];   // %globaltimer at the replay kernel's phase boundaries (diagnostics)
};

static_assert(offsetof(UpdateCounters, item_counter) == 128 && offsetof(UpdateCounters, n_chunks) == 256,
              "hot counters live on their own 128-byte lines");

struct RegAccum           // 29 exact sums + bookkeeping, device resident
{
  u64 sums[32];            // [0..20] H upper triangle row-major, [21..26] g, [27] err, [28] cnt
  unsigned ticket;
  unsigned finished;
  unsigned iterations;
  float alpha;
  float prev_err[4];
  float T[16];             // current total transform, column-major
};

struct WsTimer
{
  cudaEvent_t start, stop;
};

struct StoredChunk
{
  std::vector<uint32_t> data;
  std::list<u64>::iterator pos;             // place in ws_handle::lru
};

struct ws_handle
{
  int device = 0;
  int tau = 0, max_weight = 0, res = 0;
  int rank = 0, world = 1;
  GridDesc g{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  // scan points
  ws_pt *d_points = nullptr;      // update_tsdf staging
  size_t points_cap = 0;
  void *d_pose = nullptr;         // PoseDev + staged prior pose (fused per-scan pipeline)
  void *h_pose = nullptr;         // pinned mirror
  void *d_rays = nullptr;         // RaySetup[rays_cap]: the work list written by the set-up pass of update_tsdf
  size_t rays_cap = 0;
  uint2 *d_grp_info = nullptr;    // [2][groups] per group of 32 rays and phase: first step block with work, number of blocks
  unsigned *d_item_off = nullptr; // [2][groups + 1] exclusive prefix sums of the groups' block counts (+ the total)
  size_t grp_cap = 0;             // groups the two tables hold per phase
  size_t list_cap = 0;            // entries of the replay list
  bool tab_attr_set = false;
  // context of the last enqueued update (ws_update_finish regrows the record with it)
  UpdateParams upd_P{};
  int upd_n = 0;
  bool upd_pose_on_device = false;
  // asynchronous per-scan pipeline (ws_track_submit / ws_track_wait): two scans may be in flight
  struct TrackSlot
  {
    ws_pt *d_pts = nullptr; size_t cap = 0;
    RegAccum *h_acc = nullptr; void *h_pose = nullptr; UpdateCounters *h_ctr = nullptr;   // pinned
    cudaEvent_t copied = nullptr, done = nullptr;
    bool busy = false, finished = false; int64_t n = 0, regrows_at_submit = 0;
    std::string error;
  } track[2];
  cudaStream_t copy_stream = nullptr;
  cudaStream_t stream2 = nullptr;   // the near-field free-space march runs here, beside the surface march + merge
  cudaStream_t stream3 = nullptr;   // the record pass of the replay runs here, beside the far-field free-space march
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_merged = nullptr, ev_scanned = nullptr;
  int track_next = 0, track_in_flight = 0;
  bool rec_headroom_ok = false;     // a finished scan left the candidate record well below its capacity
  bool track_has_pose = false;      // d_pose holds the pose of the previous tracked scan      // dynamic shared memory of the lockstep march raised above 48 KB
  unsigned *d_gen_list = nullptr; // rays for the literal-arithmetic march
  // scan preprocessing scratch (preprocess.cu)
  void *d_pre_tmp = nullptr;      // duplicate-detection keys, scan order
  void *d_pre_val = nullptr;      // points handed on, scan order
  unsigned *d_pre_slot = nullptr; // hash slot of every point
  unsigned *d_pre_table = nullptr;// hash set: smallest point index per distinct value
  unsigned *d_pre_tiles = nullptr;// survivors per 1024-point tile (+ the total)
  float *d_pre_xyz = nullptr;     // staging of a host PointCloud2 payload
  size_t pre_cap = 0, pre_xyz_cap = 0;
  int64_t scan_n = 0;             // points the last ws_preprocess_scan / ws_voxelgrid_subsample left in d_points
  // voxel-grid subsample scratch (voxelgrid.cu)
  unsigned *d_vg_keys = nullptr, *d_vg_vals = nullptr, *d_vg_hist = nullptr;
  void *d_vg_box = nullptr;
  float *d_vg_xyz = nullptr;      // float centroids (optional output)
  size_t vg_cap = 0, vg_xyz_cap = 0;
  ws_pt *d_reg_points = nullptr;  // registration cloud
  ws_pt *d_reg_points_alias = nullptr;   // set while a tracked scan is enqueued: the registration works on the slot's buffer
  ws_pt *d_reg_points_last = nullptr;    // registered cloud of the last ws_track_wait
  size_t reg_points_cap = 0;
  int reg_n = 0;

  // update scratch
  UpdateCounters *d_counters = nullptr;
  UpdateCounters *h_counters = nullptr; // pinned
  // pending (interpolated winner) resolution
  unsigned pending_cap = 0;
  u64 *d_pend_addr = nullptr;           // voxel address of each pending slot
  u64 *d_pend_prev = nullptr;           // key of the current (interpolated) winner
  u64 *d_pend_key = nullptr;            // atomicMin target of the next round
  unsigned *d_active[2] = {nullptr, nullptr};   // still-pending slot lists (ping-pong)
  // far-field candidate record + replay list (same capacity), 64-entry chunks
  Rec *d_rec = nullptr;
  Rec *d_list = nullptr;
  unsigned *d_chunk_fill = nullptr;
  size_t rec_cap_chunks = 0, rec_max_chunks = 0;
  int replay_blocks = 0;                // cooperative grid of replay_kernel
  int64_t record_regrows = 0;

  // registration
  RegAccum *d_acc = nullptr;
  u64 *d_trace = nullptr;               // [max_trace][29]
  int trace_cap = 0;
  RegAccum *h_acc = nullptr;            // pinned
  u64 *d_reg_partials = nullptr;        // [2][reg_loop_blocks][32] per-block sums of the persistent GN loop
  int reg_loop_blocks = 0;
  // multi-GPU registration: mailboxes for the in-kernel exchange of the Gauss-Newton sums (registration.cu)
  void *d_mail = nullptr;               // this rank's mailbox (cudaMalloc, exported by cudaIpcGetMemHandle)
  void *peer_mail[WS_MAX_PEERS] = {};   // every rank's mailbox as mapped into this process (own one included)
  bool peer_ipc[WS_MAX_PEERS] = {};     // opened with cudaIpcOpenMemHandle (to be closed)
  bool peers_attached = false;
  unsigned reg_epoch = 0;
  unsigned long long peer_timeout_ns = 5000000000ull;

  // timing of the dominant kernels (cudaEvents on `stream`)
  int profile = 0;            // 0 off, 1 every range, 2 registration loop + whole update only
  bool timer_open = false;
  std::vector<WsTimer> timers;
  std::vector<int> timer_kind;
  size_t timers_used = 0;

  UpdateCounters last_counters{};
  int64_t last_n_points = 0;
  std::string last_error;
  int sm_count = 148;
  int64_t launches = 0;   // kernels launched by this handle

  // registration trace of the last ws_register_cloud when the host solved (device mode: d_trace)
  std::vector<i64> host_trace;
  int last_reg_iterations = 0;
  bool last_reg_host = false;

  // global chunk store (src/map/hdf5_global_map.cpp): 64^3 raw entries per chunk, a bounded number of them in
  // memory (LRU), the rest in a spill file
  std::unordered_map<u64, StoredChunk> store;
  std::list<u64> lru;                       // most recently used first
  std::unordered_map<u64, long> spilled;    // chunk key -> slot in the spill file
  size_t store_max_chunks = 4096;
  size_t spill_slots = 0;
  int64_t store_evictions = 0;
  FILE *spill_fp = nullptr;
  std::string spill_path;
  uint32_t default_entry = 0;
  // shift staging (device + pinned host), kept between calls
  uint32_t *d_shift_buf = nullptr, *h_shift_buf = nullptr;
  size_t shift_cap = 0;
};

// update_tsdf.cu
void ws_launch_update(ws_handle *h, const ws_pt *d_pts, int n, const int scanner_pos[3], const int up[3],
                      bool pose_on_device = false);
// d_xform != nullptr: the scan is first transformed in place by that column-major float[16] in device memory
// (registration.cpp:164-174, the registration result left in RegAccum::T) -- inside the set-up kernel
void ws_update_enqueue(ws_handle *h, ws_pt *d_pts, int n, const int scanner_pos[3], const int up[3],
                       bool pose_on_device, UpdateCounters *h_ctr, void *h_pose_out, const float *d_xform = nullptr);
void ws_update_finish(ws_handle *h, UpdateCounters *h_ctr, void *h_pose_out, cudaEvent_t done);
void ws_launch_pose(ws_handle *h, const float *d_X, const float *prior, int compose_reference);
void ws_compose_pose_host(const float X[16], const float prior[16], float pose[16]);
void ws_update_alloc(ws_handle *h, size_t initial_chunks, size_t max_chunks);
// registration.cu
void ws_launch_reg_reset(ws_handle *h, const float T[16], float alpha0);
void ws_launch_reg_iteration(ws_handle *h, int n, int res, int fused_solve, float it_weight_gradient, float epsilon);
void ws_launch_reg_solve(ws_handle *h, float it_weight_gradient, float epsilon);
void ws_launch_reg_loop(ws_handle *h, int n, int res, int max_iterations, float it_weight_gradient, float epsilon);
// the PoseArgs of ws_launch_pose (allocates the device pose on first use)
PoseArgs ws_pose_args(ws_handle *h, const float *prior, int compose_reference);
void ws_launch_transform_cloud(ws_handle *h, ws_pt *d_pts, int n);
void ws_host_solve(const i64 H[36], const i64 g[6], int err, int cnt, float alpha, float T[16], double xi_out[6]);
// preprocess.cu
int64_t ws_launch_preprocess(ws_handle *h, const float *d_xyz, int64_t n, int stride_floats, const float pose_mm[16], int res,
                             int mode);
// voxelgrid.cu
int64_t ws_launch_voxelgrid(ws_handle *h, const float *d_xyz, int64_t n, int stride_floats, float leaf_m, float *d_out_xyz);
// capi.cu: the chunk store (bounded, LRU, spill file)
std::vector<unsigned long long> ws_store_keys(ws_handle *h);
const uint32_t *ws_store_chunk(ws_handle *h, int cx, int cy, int cz);
// map_ops.cu
void ws_launch_fill(ws_handle *h, uint32_t entry);
void ws_launch_upload(ws_handle *h, const uint32_t *d_linear);
void ws_launch_download(ws_handle *h, uint32_t *d_linear);
void ws_box_transfer(ws_handle *h, uint32_t *d_buf, const int lo[3], const int ext[3], bool pack);
void ws_launch_checksum(ws_handle *h, int x_lo, int x_hi, int owned_only, u64 *d_out);
void ws_launch_test_reduce(ws_handle *h, const i64 *d_jacobis, const int *d_values, int n);

#define WS_TIMER_MARCH 0
#define WS_TIMER_MERGE 1
#define WS_TIMER_REG 2
#define WS_TIMER_REPLAY 3
#define WS_TIMER_UPDATE 4   // the whole update_tsdf of one scan on the handle's stream (the phases above overlap on three streams)
long ws_span_begin(ws_handle *h, int kind);
void ws_span_end(ws_handle *h, long idx);
void ws_timer_begin(ws_handle *h, int kind, cudaStream_t stream = nullptr);
void ws_timer_end(ws_handle *h, cudaStream_t stream = nullptr);

size_t ws_reg_mailbox_bytes(int world);

#define WS_CUDA_OK(expr)                                                                            \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
