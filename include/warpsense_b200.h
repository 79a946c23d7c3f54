/*
 * warpsense_b200.h -- C ABI of the B200-native TSDF hot path (libwarpsense_b200.so).
 *
 * Drop-in boundary for warpsense's two data-parallel hot paths.  The reference has no FFI layer:
 * its callers use C++ classes.  Every entry point below names the reference
 * interface it replaces (file:line under /root/reference); include/warpsense_b200.hpp rebuilds those
 * classes (cuda::DeviceMap, cuda::DeviceMapMemWrapper, cuda::TSDFCuda, cuda::RegistrationCuda,
 * cuda::TSDFMapping, cuda::TSDFRegistration) as header-only shims over this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types;
 *   - every call returns WS_OK (0) or a negative error code; ws_last_error() gives the text;
 *   - points are packed {int32 x,y,z} millimetres in the map frame (rmagine::Pointi, 12 bytes);
 *   - TSDF entries are the reference's packed uint32 {int16 value | int16 weight << 16}
 *     (include/map/tsdf.h:16-35); host grids use the reference's ring layout, x slowest, z fastest
 *     (include/map/hdf5_local_map.h:140-151), index computed in 64 bit;
 *   - 4x4 transforms are column-major float[16] (Eigen::Matrix4f storage), translation in mm;
 *   - H is column-major int64[36], g int64[6] (rmagine::Matrix6x6l / Point6l);
 *   - a handle is thread-compatible: callers serialise access exactly as the reference does with
 *     TSDFMapping::mutex_ (src/warpsense/tsdf_mapping.cpp:67,73,93).
 */
#ifndef WARPSENSE_B200_H
#define WARPSENSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WS_OK 0
#define WS_ERR_INVALID (-1)   /* bad argument                                            */
#define WS_ERR_CUDA (-2)      /* CUDA runtime failure (the reference exit(1)s here)      */
#define WS_ERR_CAPACITY (-3)  /* scan too large (reference: > 1,000,000 points returns)  */
#define WS_ERR_STATE (-4)     /* call order violated (e.g. reg step before prepare)      */

#define WS_MAX_POINTS (1 << 24)

typedef struct ws_handle ws_handle;

typedef struct { int32_t x, y, z; } ws_point;   /* rmagine::Pointi */

/* work counters of the last ws_update_tsdf (SURVEY.md 8d "work counters to print") */
typedef struct {
  int64_t n_points;
  int64_t n_candidates;      /* C: candidate writes                                   */
  int64_t n_touched;         /* T: distinct voxels that received a candidate (this rank's slab) */
  int64_t n_written;         /* voxels whose stored entry was rewritten               */
  int64_t n_touched_bricks;  /* 8x8x8 bricks streamed by the merge pass               */
  int64_t n_parked;          /* voxels that needed a replay round                     */
  int64_t n_rounds;          /* replay rounds that did work                           */
  int64_t n_list;            /* recorded candidates that followed a parked winner (replay list) */
  int64_t n_record_chunks;   /* 64-entry chunks of the far-field candidate record handed out   */
  int64_t replay_phase_ns[3];/* replay kernel: record pass, first resolve, remaining rounds    */
} ws_update_counters;

/* flags of ws_register_cloud */
#define WS_REG_DEVICE_SOLVE 0   /* FP64 solve + pose update on the GPU, no per-iteration host round trip */
#define WS_REG_HOST_SOLVE 1     /* reference structure: sums to the host every iteration (tsdf_registration.cpp:55-92) */

/* ---- lifetime -------------------------------------------------------------------------------
 * ws_create: cuda::TSDFCuda::TSDFCuda(existing_map, tau, max_weight, map_resolution)
 *            (include/warpsense/cuda/update_tsdf.h:12, src/warpsense/cuda/update_tsdf.cu:130-141)
 *            + cuda::RegistrationCuda::RegistrationCuda(map) (include/warpsense/cuda/registration.h:13).
 * size[3] are the local-map side lengths as HDF5LocalMap stores them (odd; even sizes must already
 * be s+1, src/map/hdf5_local_map.cpp:6-8).  The grid starts filled with (tau, 0) at pos 0,
 * offset size/2 (hdf5_local_map.cpp:11-19).
 * ws_create_sharded: same, but this handle holds only rank's x-slab of the ring (+1 halo column each
 * side) out of `world` slabs -- one handle per GPU/process. */
int ws_create(const int32_t size[3], int32_t tau, int32_t max_weight, int32_t map_resolution,
              int32_t device, ws_handle **out);
int ws_create_sharded(const int32_t size[3], int32_t tau, int32_t max_weight, int32_t map_resolution,
                      int32_t device, int32_t rank, int32_t world, ws_handle **out);
/* cuda::TSDFCuda::~TSDFCuda / RegistrationCuda::~RegistrationCuda (no cudaDeviceReset: the process may
 * hold other handles; the reference resets in cleanup.cu:3-11) */
void ws_destroy(ws_handle *h);
const char *ws_last_error(const ws_handle *h);
/* run all work of this handle on an existing CUDA stream (cudaStream_t); NULL restores the private one.
 * The reference uses the legacy default stream with blocking copies (SURVEY.md 8b "Threading"). */
int ws_set_stream(ws_handle *h, void *cuda_stream);
void *ws_get_stream(ws_handle *h);
int ws_sync(ws_handle *h);                         /* cudaDeviceSynchronize() calls of the reference */

/* ---- map transfer ---------------------------------------------------------------------------
 * cuda::DeviceMapMemWrapper::to_device / to_host / update_params
 * (include/warpsense/cuda/device_map_wrapper.h:22-27, src/warpsense/cuda/device_map_wrapper.cu:35-92).
 * `entries` is the host ring array of size[0]*size[1]*size[2] raw entries (HDF5LocalMap::get_data()). */
int ws_map_upload(ws_handle *h, const uint32_t *entries, const int32_t size[3],
                  const int32_t offset[3], const int32_t pos[3]);
int ws_map_download(ws_handle *h, uint32_t *entries);
int ws_map_set_params(ws_handle *h, const int32_t offset[3], const int32_t pos[3]);
int ws_map_get_params(const ws_handle *h, int32_t size[3], int32_t offset[3], int32_t pos[3]);
/* reset every voxel to (value, weight) and make it the default entry of unseen chunks --
 * HDF5GlobalMap(name, initial_value, initial_weight) + HDF5LocalMap ctor fill
 * (hdf5_global_map.cpp:24-27, hdf5_local_map.cpp:15-19) */
int ws_map_fill(ws_handle *h, int32_t value, int32_t weight);
/* Order-free 64-bit checksum of the stored entries of ring-x rows [x_lo, x_hi) (storage coordinates,
 * include/map/hdf5_local_map.h:140-151), computed on the device; owned_only != 0 restricts a sharded handle to
 * the rows it owns (halo columns excluded).  Equal map contents give equal checksums whatever the sharding:
 * the multi-GPU result check of bench.py and tests (no reference counterpart: the reference is single-GPU). */
int ws_map_checksum(ws_handle *h, int32_t x_lo, int32_t x_hi, int32_t owned_only, uint64_t *out);

/* single-voxel access in map (voxel) coordinates: DeviceMap::value_unchecked + in_bounds
 * (include/warpsense/cuda/device_map.h:94-150); returns WS_ERR_INVALID when out of bounds */
int ws_map_get_voxel(ws_handle *h, int32_t x, int32_t y, int32_t z, uint32_t *entry);
int ws_map_set_voxel(ws_handle *h, int32_t x, int32_t y, int32_t z, uint32_t entry);

/* ---- TSDF update ----------------------------------------------------------------------------
 * cuda::TSDFCuda::update_tsdf(scan_points, scanner_pos, up)
 * (include/warpsense/cuda/update_tsdf.h:14, src/warpsense/cuda/update_tsdf.cu:143-166); results equal
 * the CPU path update_tsdf(...) of src/cpu/update_tsdf.cpp:397-564.  scanner_pos is in voxels, up is
 * R*(0,0,32768) (src/warpsense/tsdf_mapping.cpp:77-85).  Blocking, like the reference.
 * n > WS_MAX_POINTS returns WS_ERR_CAPACITY (reference: prints and returns, update_tsdf.cu:146-150). */
int ws_update_tsdf(ws_handle *h, const ws_point *points, int64_t n,
                   const int32_t scanner_pos[3], const int32_t up[3]);
/* same with the points already in device memory (e.g. the cloud ws_register_cloud just transformed) */
int ws_update_tsdf_device(ws_handle *h, const ws_point *device_points, int64_t n,
                          const int32_t scanner_pos[3], const int32_t up[3]);
int ws_get_update_counters(const ws_handle *h, ws_update_counters *out);

/* ---- registration ---------------------------------------------------------------------------
 * ws_reg_prepare: cuda::RegistrationCuda::prepare_registration(points)
 *                 (include/warpsense/cuda/registration.h:15, src/warpsense/cuda/registration.cu:303-308).
 * ws_reg_step:    cuda::RegistrationCuda::perform_registration(map_dev, pretransform, h, g, e, c, res)
 *                 (registration.h:16-19, registration.cu:347-368): one accumulation pass for transform T.
 *                 All points are used (the reference kernel silently stops at 65,536).
 * ws_register_cloud: cuda::TSDFRegistration::register_cloud(cloud, pretransform)
 *                 (include/warpsense/tsdf_registration.h:24, src/warpsense/tsdf_registration.cpp:29-96)
 *                 following the CPU loop register_cloud(...) of src/cpu/registration.cpp:14-177
 *                 (centre recomputed every iteration; cloud transformed in place at the end).
 *                 `cloud` may be NULL to register the points given to ws_reg_prepare and leave the
 *                 transformed cloud on the device (ws_reg_points_device). */
/* ---- scan preprocessing ------------------------------------------------------
 * ws_preprocess_scan: App::preprocess(cloud, scan_points) (src/warpsense/app.cpp:118-148): the x/y/z floats of
 *                 a PointCloud2 payload (metres; `point_step_bytes` between points, >= 12) -> drop returns
 *                 with x, y and z all below 0.3 -> millimetres -> voxel centre (float arithmetic) ->
 *                 transform_point(., to_int_mat(pose_mm)) -> duplicates removed.  The reference copies its
 *                 std::unordered_set out in bucket order; here the survivors keep scan order (first
 *                 occurrence stays): the same set, deterministic.  The result stays on the device
 *                 (ws_scan_points_device) for ws_update_tsdf_device / ws_reg_prepare_device and is copied
 *                 to `out_host` (capacity n) when that is not NULL.
 *                 `on_device`: bit 0 = xyz is a device pointer; WS_PREPROCESS_CPU_NODE selects the CPU node's
 *                 variant instead (src/cpu/fastsense.cpp:143-163: duplicates are detected by the voxel of the
 *                 untransformed millimetre point and the point itself is transformed, in scan order -- which
 *                 is exactly the order produced here). */
#define WS_PREPROCESS_CPU_NODE 2
int ws_preprocess_scan(ws_handle *h, const float *xyz, int64_t n, int32_t point_step_bytes, int32_t on_device,
                       const float pose_mm[16], int32_t map_resolution, ws_point *out_host, int64_t *n_out);
const ws_point *ws_scan_points_device(ws_handle *h, int64_t *n);

/* ---- featsense feed ---------------------------------------------------------------------------
 * ws_voxelgrid_subsample: the pcl::VoxelGrid of cuda::TSDFMapping::preprocess_from_ros
 *   (src/warpsense/tsdf_mapping.cpp:147-151; PCL 1.10 filters/impl/voxel_grid.hpp applyFilter) on the device: one
 *   float centroid per occupied leaf, output in ascending leaf index (x fastest), then metres -> int millimetres
 *   (:153-158).  `xyz`: n points `point_step_bytes` apart, host or device memory.  The millimetre points stay on
 *   the device (ws_scan_points_device) and are copied to out_mm_host / out_xyz_host (float centroids) if given.
 * ws_update_tsdf_from_ros: TSDFMapping::update_tsdf_from_ros(cloud, pose) (tsdf_mapping.cpp:165-173):
 *   preprocess_from_ros (voxel grid; pose_m = column-major 4x4 Isometry3d in metres -> mm Matrix4f through the
 *   quaternion round trip of :160-161), then update_tsdf(points, mm_pose).  The caller pushes out_mm_pose to
 *   its pose buffer (:170). */
int ws_voxelgrid_subsample(ws_handle *h, const float *xyz, int64_t n, int32_t point_step_bytes, int32_t on_device,
                           float leaf_m, ws_point *out_mm_host, float *out_xyz_host, int64_t *n_out);
int ws_update_tsdf_from_ros(ws_handle *h, const float *xyz_m, int64_t n, int32_t point_step_bytes, int32_t on_device,
                            const double pose_m[16], float out_mm_pose[16], int64_t *n_points);


int ws_reg_prepare(ws_handle *h, const ws_point *points, int64_t n);
/* same, the cloud already being in device memory (device-to-device copy on the handle's stream) */
int ws_reg_prepare_device(ws_handle *h, const ws_point *device_points, int64_t n);
int ws_reg_step(ws_handle *h, const float T[16], int32_t map_resolution,
                int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt);
int ws_register_cloud(ws_handle *h, ws_point *cloud, int64_t n, const float pretransform[16],
                      int32_t max_iterations, float it_weight_gradient, float epsilon,
                      int32_t map_resolution, int32_t flags, float out_transform[16],
                      int32_t *iterations);
/* ---- per-scan pipeline ------------------------------------------------------------------------
 * The three calls App::cloud_callback makes per scan (src/warpsense/app.cpp:65-112: update_tsdf, register_cloud,
 * update_pose_estimate) as one stream of kernels, in the order of BASELINE configs[2]: register_cloud against the
 * map of the earlier scans, pose update, update_tsdf with the registered cloud and convert_pose_to_gpu(pose)
 * (tsdf_mapping.cpp:77-85); transform, scanner voxel and up vector stay on the device between the two halves.
 * NOTE the reference's own callback updates the map with the OLD pose first and registers afterwards; callers that
 * want that order use ws_update_tsdf + ws_register_cloud.
 *   pretransform  the IMU pre-rotation handed to register_cloud (app.cpp:97-102), NULL = identity
 *   prior_pose    the pose the scan was placed with; NULL (submit only) = the pose of the previous tracked scan,
 *                 taken from device memory
 *   flags         WS_TRACK_REFERENCE_POSE: compose the new pose as App::update_pose_estimate does (app.cpp:172-176,
 *                 src/cpu/fastsense.cpp:219-221): R = X.R * R, t += X.t.  Default: the full product X * prior_pose
 *                 (float32, accumulated over k = 0..3 in order).
 * ws_track_submit enqueues and returns (host points are copied on a second stream; the buffer must stay valid until
 * the matching ws_track_wait); at most two scans are in flight.  ws_track_wait blocks until the scan is done and
 * returns the registration transform, the new pose and (ws_get_update_counters) the work counters.
 * ws_track_scan[_ex] = submit + wait. */
#define WS_TRACK_REFERENCE_POSE 1
int ws_track_submit(ws_handle *h, const ws_point *points, int64_t n, int32_t on_device, const float *prior_pose,
                    const float *pretransform, int32_t max_iterations, float it_weight_gradient, float epsilon,
                    int32_t map_resolution, int32_t flags, int32_t *ticket);
int ws_track_wait(ws_handle *h, int32_t ticket, float out_transform[16], float out_pose[16], int32_t *iterations);
int ws_track_scan_ex(ws_handle *h, const ws_point *points, int64_t n, int32_t on_device, const float prior_pose[16],
                     const float *pretransform, int32_t max_iterations, float it_weight_gradient, float epsilon,
                     int32_t map_resolution, int32_t flags, float out_transform[16], float out_pose[16],
                     int32_t *iterations);
int ws_track_scan(ws_handle *h, const ws_point *points, int64_t n, int32_t on_device, const float prior_pose[16],
                  int32_t max_iterations, float it_weight_gradient, float epsilon, int32_t map_resolution,
                  float out_transform[16], float out_pose[16], int32_t *iterations);
/* per-iteration sums of the last ws_register_cloud: 29 int64 per iteration
 * (21 upper-triangle H row-major, 6 g, err, cnt); returns the number of iterations copied */
int ws_reg_get_trace(ws_handle *h, int64_t *out, int32_t max_iterations);
const ws_point *ws_reg_points_device(ws_handle *h, int64_t *n);

/* multi-GPU registration building blocks (one handle per rank; the caller all-reduces the 29 sums,
 * e.g. ncclAllReduce(int64, sum) on ws_reg_sums_device(), between the two calls -- SURVEY.md 8e) */
int ws_reg_begin(ws_handle *h, const float pretransform[16]);
int ws_reg_accumulate(ws_handle *h, int32_t map_resolution);           /* local 29 sums, async   */
void *ws_reg_sums_device(ws_handle *h);                                /* int64[29] device ptr   */
int ws_reg_sums_get(ws_handle *h, int64_t sums[29]);                   /* host copies of the same  */
int ws_reg_sums_set(ws_handle *h, const int64_t sums[29]);             /* (MPI / test reductions)  */
int ws_reg_solve(ws_handle *h, float it_weight_gradient, float epsilon);/* solve + pose update    */
int ws_reg_peek(ws_handle *h, int32_t *iterations, int32_t *finished);  /* progress so far (synchronises) */
int ws_reg_finish(ws_handle *h, float out_transform[16], int32_t *iterations, int32_t *finished);

/* Fused multi-GPU registration: with every rank's mailbox attached, ws_register_cloud on a sharded handle runs
 * the whole Gauss-Newton loop in ONE persistent kernel per rank and exchanges the 29 sums over NVLink peer
 * memory inside it (registration.cu "PeerParams"); every rank must call ws_register_cloud with the same
 * arguments at the same time.  Replaces the per-iteration blocking copies + host solve of
 * src/warpsense/tsdf_registration.cpp:55-92 and the NCCL all-reduce SURVEY.md 8e would otherwise need.
 *   ws_peer_export      : 64-byte cudaIpcMemHandle of this rank's mailbox (send it to the other ranks)
 *   ws_peer_attach_ipc  : handles[world][64] of all ranks in rank order (own entry ignored), other processes
 *   ws_peer_local_ptr / ws_peer_attach_ptrs : the same for ranks that live in ONE process (raw device
 *                         pointers; peer access between the devices is enabled here)
 *   ws_peer_set_timeout : give up (WS_ERR_STATE) when a peer does not deliver within `seconds` (default 5) */
#define WS_IPC_HANDLE_BYTES 64
int ws_peer_export(ws_handle *h, uint8_t handle[WS_IPC_HANDLE_BYTES]);
int ws_peer_attach_ipc(ws_handle *h, const uint8_t *handles, int32_t world);
void *ws_peer_local_ptr(ws_handle *h);
int ws_peer_attach_ptrs(ws_handle *h, void *const *mailboxes, const int32_t *devices, int32_t world);
int ws_peer_set_timeout(ws_handle *h, double seconds);

/* Host-only (no GPU needed): the x-slab of rank `rank` out of `world` for a map of ring side size_x --
 * owned ring-x rows [own_lo, own_hi) and the resident 8-row brick columns (owned + one halo row each
 * side); returns the number of resident columns or WS_ERR_INVALID. */
int ws_slab_layout(int32_t size_x, int32_t rank, int32_t world, int32_t *own_lo, int32_t *own_hi,
                   int32_t *resident_cols, int32_t cap);

/* which ring-x brick columns (8 rows each) a handle owns / keeps resident: contiguous slabs by default, stripes
 * of WS_STRIPE_COLS columns dealt round-robin to the ranks when that environment variable is set (the same
 * value on every rank).  Returns the number of columns. */
int ws_shard_columns(const ws_handle *h, uint8_t *owned, uint8_t *resident, int32_t cap);
/* the same without a handle (host only): stripe_cols <= 0 gives the slabs of ws_slab_layout */
int ws_shard_layout(int32_t size_x, int32_t rank, int32_t world, int32_t stripe_cols, uint8_t *owned, uint8_t *resident,
                    int32_t cap);

/* test hook for the reduction shape (test/cuda.cpp:416-532): sums of n Jacobians with values */
int ws_test_reduce(ws_handle *h, const int64_t *jacobis6, const int32_t *values, int64_t n,
                   int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt);

/* ---- local-map shift ------------------------------------------------------------------------
 * HDF5LocalMap::shift(new_pos) on the device-resident grid (src/map/hdf5_local_map.cpp:53-118;
 * today a whole-grid D2H + H2D round trip, src/warpsense/tsdf_mapping.cpp:115-123).  Leaving slabs are
 * saved into an in-memory 64^3 chunk store (src/map/hdf5_global_map.cpp:53-137), entering slabs are
 * loaded from it or default-filled. */
int ws_shift(ws_handle *h, const int32_t new_pos[3]);
/* HDF5LocalMap::write_back (hdf5_local_map.cpp:210-217): save the whole local map into the store */
int ws_write_back(ws_handle *h);
int64_t ws_store_num_chunks(const ws_handle *h);
int ws_store_chunk_list(const ws_handle *h, int32_t *xyz, int64_t cap);
/* copy one 64^3 chunk (index x*4096 + y*64 + z, hdf5_global_map.cpp:53-57); WS_ERR_INVALID if absent */
int ws_store_get_chunk(const ws_handle *h, int32_t cx, int32_t cy, int32_t cz, uint32_t *out);
/* put a chunk into the store (what loading a map file does chunk by chunk) */
int ws_store_set_chunk(ws_handle *h, int32_t cx, int32_t cy, int32_t cz, const uint32_t *data);
/* HDF5GlobalMap keeps 64 active chunks and writes the least recently used one to its file
 * (hdf5_global_map.h:78, hdf5_global_map.cpp:59-137).  Here at most `max_chunks_in_memory` chunks (default 4096,
 * WS_STORE_CHUNKS) stay in memory; the least recently used one goes to a spill file (WS_STORE_SPILL_DIR, default
 * /tmp) and is read back on its next activation. */
int ws_store_configure(ws_handle *h, int64_t max_chunks_in_memory);
int64_t ws_store_evictions(const ws_handle *h);

/* ---- global-map file ----------------------------------------------------------------------------
 * ws_export_hdf5: what HDF5GlobalMap leaves on disk (src/map/hdf5_global_map.cpp): /map/<cx>_<cy>_<cz> = 64^3
 * uint32 raw entries per stored chunk (:46-57,:120,:170), the /map attributes of write_meta (:208-221) and
 * /poses/<n>/pose = 7 float32 (x y z qx qy qz qw, already scaled and rounded as write_pose does, :175-200).
 * Written directly in the HDF5 file format (superblock v0, object headers v1, symbol-table groups,
 * contiguous little-endian datasets) -- no libhdf5.  Call ws_write_back first to include the local map. */
typedef struct ws_map_meta
{
  int32_t tau;
  int32_t map_size[3];
  float max_distance;
  int32_t map_resolution;
  int32_t max_weight;
} ws_map_meta;
int ws_export_hdf5(ws_handle *h, const char *path, const ws_map_meta *meta, const float *poses7, int64_t n_poses);
/* the same file from caller-held chunks (host only, no handle): chunk_xyz[n][3], chunk_data[n][64^3] */
int ws_hdf5_write_chunks(const char *path, const ws_map_meta *meta, const int32_t *chunk_xyz,
                         const uint32_t *chunk_data, int64_t n_chunks, const float *poses7, int64_t n_poses);

/* ws_import_hdf5: read such a file back (the reference cannot: HDF5GlobalMap truncates its file on open,
 * hdf5_global_map.cpp:5,24 -- no resume).  Every /map/<cx>_<cy>_<cz> chunk goes into the handle's chunk store
 * (one at a time: bounded memory), *meta receives the /map attributes, poses7 the first poses_cap rows of
 * /poses/<n>/pose in ascending n.  ws_map_reload then fills the device-resident local map from the store. */
int ws_import_hdf5(ws_handle *h, const char *path, ws_map_meta *meta, float *poses7, int64_t poses_cap,
                   int64_t *n_chunks, int64_t *n_poses);
/* (re)load the whole local map window from the chunk store: what HDF5LocalMap's constructor does for a fresh map
 * (src/map/hdf5_local_map.cpp:5-20: save_load_area(..., false) over the window). */
int ws_map_reload(ws_handle *h);

/* ---- timing -------------------------------------------------------------------------------- */
/* record cudaEvents around the hot kernels (kind: 0 march, 1 brick list + merge, 2 registration loop, 3 replay,
 * 4 the whole update_tsdf of one scan on the handle's stream).  Kinds 0, 1 and 3 are ranges on three streams that
 * run side by side (surface march + merge | near-field free-space march | record pass of the replay): their sums
 * exceed kind 4, which is the elapsed time of the update.  on: 0 off, 1 every range (~20 event records per scan:
 * measured 3 % of the scan rate), 2 kinds 2 and 4 only (what a timed run can afford). */
int ws_profile_enable(ws_handle *h, int32_t on);
int ws_profile_reset(ws_handle *h);
/* sum of elapsed ms and launch count per kind since the last reset (synchronises the stream) */
int ws_profile_get(ws_handle *h, int32_t kind, double *total_ms, int64_t *launches);
/* the ranges of the last update_tsdf as (kind, start_ms, stop_ms) triples relative to its start, in launch order
 * (first the kind-4 span itself); returns the number of ranges written, -1 on error.  Shows what ran side by side. */
int64_t ws_profile_timeline(ws_handle *h, double *out, int64_t cap_ranges);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t ws_launch_count(const ws_handle *h);

const char *ws_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WARPSENSE_B200_H */
