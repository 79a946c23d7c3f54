// warpsense_b200.hpp -- the reference's hot-path classes as header-only C++17 shims over the C ABI of
// warpsense_b200.h.  Same names, argument meaning and error behaviour as the reference (citations are
// file:line under /root/reference), minus ROS / PCL / Eigen / HighFive types:
//
//   rmagine::Pointi            include/warpsense/math/vector3.h:416      packed {int x, y, z}
//   Matrix4f                   Eigen::Matrix4f storage                   column-major float[16]
//   HostLocalMap               HDF5LocalMap's in-memory part             include/map/hdf5_local_map.h
//   cuda::DeviceMap            include/warpsense/cuda/device_map.h:32-164
//   cuda::DeviceMapMemWrapper  include/warpsense/cuda/device_map_wrapper.h:10-36
//   cuda::TSDFCuda             include/warpsense/cuda/update_tsdf.h:9-34
//   cuda::RegistrationCuda     include/warpsense/cuda/registration.h:10-45
//   cuda::TSDFMapping          include/warpsense/tsdf_mapping.h:17-60
//   cuda::TSDFRegistration     include/warpsense/tsdf_registration.h:17-27
//   ConcurrentRingBuffer       include/util/concurrent_ring_buffer.h (the push_nb / pop_nb / clear subset)
//   featsense::MappingFeed     the TSDF-facing half of Mapping::thread_run, src/featsense/mapping.cpp:39-147
//
// Everything that touches voxels or points runs in libwarpsense_b200.so on the GPU; this header only
// marshals arguments.  CUDA failures throw std::runtime_error (the reference prints and exit(1)s,
// include/warpsense/cuda/common.cuh:10-21).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "warpsense_b200.h"

namespace rmagine
{
struct Pointi
{
  int x = 0, y = 0, z = 0;
  Pointi() = default;
  Pointi(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
};
static_assert(sizeof(Pointi) == sizeof(ws_point), "rmagine::Pointi must be 12 packed bytes");
}  // namespace rmagine

namespace warpsense_b200
{
constexpr int WEIGHT_RESOLUTION = 64;       // include/warpsense/consts.h:9-10
constexpr int MATRIX_RESOLUTION = 32768;    // include/warpsense/consts.h:12-13

struct Matrix4f    // column-major, like Eigen::Matrix4f
{
  float m[16];
  static Matrix4f Identity()
  {
    Matrix4f r{};
    for (int i = 0; i < 4; i++) r.m[i * 4 + i] = 1.f;
    return r;
  }
  float &operator()(int row, int col) { return m[col * 4 + row]; }
  float operator()(int row, int col) const { return m[col * 4 + row]; }
};

// TSDFEntry (include/map/tsdf.h:16-57)
struct TSDFEntry
{
  uint32_t raw = 0;
  TSDFEntry() = default;
  TSDFEntry(int value, int weight) : raw((uint32_t)(uint16_t)(int16_t)value | ((uint32_t)(uint16_t)(int16_t)weight << 16)) {}
  int value() const { return (int16_t)(raw & 0xFFFFu); }
  int weight() const { return (int16_t)(raw >> 16); }
};

// include/params/map_params.h:49-106, registration_params.h:22-30 (unit scaling included)
struct MapParams
{
  int resolution = 64;
  float max_distance = 0.6f;
  int tau = 600;           // int(max_distance * 1000)
  int max_weight = 640;    // max_weight * WEIGHT_RESOLUTION
  float shift = 3.0f;      // metres the sensor must move before the local map is shifted
  float update_distance = 0.5f;   // metres between two TSDF updates of the featsense feed (mapping.cpp:81)
  int size[3] = { 313, 313, 79 };
  static MapParams from_ros(float max_distance_m, int max_weight, const float size_m[3], int resolution_mm, float shift_m)
  {
    MapParams p;
    p.resolution = resolution_mm;
    p.max_distance = max_distance_m;
    p.tau = (int)(max_distance_m * 1000.f);
    p.max_weight = max_weight * WEIGHT_RESOLUTION;
    p.shift = shift_m;
    for (int a = 0; a < 3; a++) p.size[a] = (int)size_m[a] * 1000 / resolution_mm;
    return p;
  }
};
struct RegistrationParams
{
  int max_iterations = 200;
  float it_weight_gradient = 0.1f;
  float epsilon = 0.03f;
};
struct Params
{
  MapParams map;
  RegistrationParams registration;
};

// The in-memory part of HDF5LocalMap (src/map/hdf5_local_map.cpp:5-20): ring array + size/pos/offset.
class HostLocalMap
{
public:
  using Ptr = std::shared_ptr<HostLocalMap>;
  HostLocalMap(int sx, int sy, int sz, int default_value, int default_weight)
  {
    const int s[3] = { sx, sy, sz };
    for (int a = 0; a < 3; a++)
    {
      size_[a] = s[a] % 2 == 1 ? s[a] : s[a] + 1;        // hdf5_local_map.cpp:6-8
      pos_[a] = 0;
      offset_[a] = size_[a] / 2;
    }
    data_.assign((size_t)size_[0] * size_[1] * size_[2], TSDFEntry(default_value, default_weight));
  }
  int *get_size() { return size_; }
  int *get_pos() { return pos_; }
  int *get_offset() { return offset_; }
  TSDFEntry *get_data() { return data_.data(); }
  bool in_bounds(int x, int y, int z) const                // hdf5_local_map.h:275-279
  {
    return std::abs(x - pos_[0]) <= size_[0] / 2 && std::abs(y - pos_[1]) <= size_[1] / 2 && std::abs(z - pos_[2]) <= size_[2] / 2;
  }
  TSDFEntry &value(int x, int y, int z)                    // hdf5_local_map.h:140-181 (64-bit index)
  {
    if (!in_bounds(x, y, z)) throw std::out_of_range("Index out of bounds");
    auto ov = [](int v, int m) { return ((v % m) + m) % m; };
    const int64_t i = ((int64_t)ov(x - pos_[0] + offset_[0], size_[0]) * size_[1] + ov(y - pos_[1] + offset_[1], size_[1])) * size_[2] +
                      ov(z - pos_[2] + offset_[2], size_[2]);
    return data_[(size_t)i];
  }

private:
  int size_[3], pos_[3], offset_[3];
  std::vector<TSDFEntry> data_;
};

inline void ws_check(ws_handle *h, int rc, const char *what)
{
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + (h ? ws_last_error(h) : "no handle") + " (code " + std::to_string(rc) + ")");
}

// include/util/util.h:8-18,52-56 (host helpers used to prepare the call arguments)
inline void to_int_mat(const Matrix4f &mat, int out[16])
{
  for (int i = 0; i < 16; i++) out[i] = (int)(mat.m[i] * (float)MATRIX_RESOLUTION);
}
inline rmagine::Pointi transform_point(const rmagine::Pointi &p, const int M[16])
{
  int r[3];
  for (int i = 0; i < 3; i++)
  {
    const int32_t acc = (int32_t)((uint32_t)M[i] * (uint32_t)p.x + (uint32_t)M[4 + i] * (uint32_t)p.y + (uint32_t)M[8 + i] * (uint32_t)p.z + (uint32_t)M[12 + i]);
    r[i] = acc / MATRIX_RESOLUTION;
  }
  return rmagine::Pointi(r[0], r[1], r[2]);
}
inline rmagine::Pointi to_map(const Matrix4f &pose, int map_resolution)
{
  return rmagine::Pointi((int)std::floor(pose.m[12] / map_resolution), (int)std::floor(pose.m[13] / map_resolution),
                         (int)std::floor(pose.m[14] / map_resolution));
}
}  // namespace warpsense_b200

// include/util/concurrent_ring_buffer.h, the subset the hot path's callers use: bounded FIFO, non-blocking
// push (drops the oldest entry when full if `force`), non-blocking pop.
template <typename T>
class ConcurrentRingBuffer
{
public:
  using Ptr = std::shared_ptr<ConcurrentRingBuffer<T>>;
  explicit ConcurrentRingBuffer(size_t size) : cap_(size) {}
  bool push_nb(const T &v, bool force = false)
  {
    std::lock_guard<std::mutex> lock(m_);
    if (q_.size() >= cap_)
    {
      if (!force) return false;
      q_.pop_front();
    }
    q_.push_back(v);
    return true;
  }
  bool pop_nb(T *out)
  {
    std::lock_guard<std::mutex> lock(m_);
    if (q_.empty()) return false;
    *out = q_.front();
    q_.pop_front();
    return true;
  }
  void clear() { std::lock_guard<std::mutex> lock(m_); q_.clear(); }
  bool empty() const { std::lock_guard<std::mutex> lock(m_); return q_.empty(); }
  size_t size() const { std::lock_guard<std::mutex> lock(m_); return q_.size(); }

private:
  size_t cap_;
  mutable std::mutex m_;
  std::deque<T> q_;
};

namespace cuda
{
using warpsense_b200::HostLocalMap;
using warpsense_b200::Matrix4f;
using warpsense_b200::Params;
using warpsense_b200::TSDFEntry;
using warpsense_b200::ws_check;

// cuda::DeviceMap: a non-owning view {size, offset, data, pos} of a host ring array (device_map.h:32-66)
struct DeviceMap
{
  DeviceMap(int *size, int *offset, TSDFEntry *data, int *pos) : size_(size), offset_(offset), data_(data), pos_(pos) {}
  explicit DeviceMap(const std::shared_ptr<HostLocalMap> &map)
      : size_(map->get_size()), offset_(map->get_offset()), data_(map->get_data()), pos_(map->get_pos()) {}
  DeviceMap() = default;
  const int *get_size() const { return size_; }
  const int *get_offset() const { return offset_; }
  const int *get_pos() const { return pos_; }
  int *size_ = nullptr, *offset_ = nullptr;
  TSDFEntry *data_ = nullptr;
  int *pos_ = nullptr;
};

// device_map_wrapper.h:10-36 over a shared ws_handle
struct DeviceMapMemWrapper
{
  explicit DeviceMapMemWrapper(ws_handle *h) : h_(h) {}
  void to_device(const DeviceMap &m)                               // device_map_wrapper.cu:64-75
  {
    ws_check(h_, ws_map_upload(h_, reinterpret_cast<const uint32_t *>(m.data_), m.size_, m.offset_, m.pos_), "to_device");
  }
  void to_host(const DeviceMap &m)                                 // device_map_wrapper.cu:77-92
  {
    ws_check(h_, ws_map_download(h_, reinterpret_cast<uint32_t *>(m.data_)), "to_host");
    int32_t s[3];
    ws_check(h_, ws_map_get_params(h_, s, m.offset_, m.pos_), "to_host");
  }
  void update_params(const DeviceMap &m)                           // device_map_wrapper.cu:35-43
  {
    ws_check(h_, ws_map_set_params(h_, m.offset_, m.pos_), "update_params");
  }
  ws_handle *dev() const { return h_; }
  ws_handle *h_;
};

// update_tsdf.h:9-34
class TSDFCuda
{
public:
  explicit TSDFCuda(const DeviceMap &existing_map, int tau, int max_weight, int map_resolution, int device = 0)
  {
    const int rc = ws_create(existing_map.size_, tau, max_weight, map_resolution, device, &h_);
    if (rc != WS_OK) throw std::runtime_error("TSDFCuda: ws_create failed (bad arguments or no usable CUDA device)");
    avg_map_.reset(new DeviceMapMemWrapper(h_));
    avg_map_->to_device(existing_map);
  }
  TSDFCuda(const TSDFCuda &) = delete;
  TSDFCuda &operator=(const TSDFCuda &) = delete;
  ~TSDFCuda() { ws_destroy(h_); }

  // update_tsdf.cu:169-183; more than n_max_points_ points: prints and returns, like update_tsdf.cu:146-150
  void update_tsdf(const std::vector<rmagine::Pointi> &scan_points, const rmagine::Pointi &scanner_pos, const rmagine::Pointi &up)
  {
    const int32_t sp[3] = { scanner_pos.x, scanner_pos.y, scanner_pos.z }, u[3] = { up.x, up.y, up.z };
    const int rc = ws_update_tsdf(h_, reinterpret_cast<const ws_point *>(scan_points.data()), (int64_t)scan_points.size(), sp, u);
    if (rc == WS_ERR_CAPACITY) { std::fprintf(stderr, "Too many points in scan (%zu)\n", scan_points.size()); return; }
    ws_check(h_, rc, "update_tsdf");
  }
  void update_tsdf(DeviceMap &result, const std::vector<rmagine::Pointi> &scan_points, const rmagine::Pointi &scanner_pos, const rmagine::Pointi &up)
  {
    update_tsdf(scan_points, scanner_pos, up);                     // update_tsdf.cu:143-166
    avg_map_->to_host(result);
  }
  ws_handle *device_map() { return h_; }
  DeviceMapMemWrapper &avg_map() { return *avg_map_; }
  DeviceMapMemWrapper &new_map() { return *avg_map_; }             // scratch keys live beside the map here
  ws_update_counters counters() const
  {
    ws_update_counters c;
    ws_get_update_counters(h_, &c);
    return c;
  }
  static constexpr int n_max_points_ = WS_MAX_POINTS;

private:
  ws_handle *h_ = nullptr;
  std::unique_ptr<DeviceMapMemWrapper> avg_map_;
};

// registration.h:10-45 (shares the device map of a TSDFCuda)
class RegistrationCuda
{
public:
  explicit RegistrationCuda(ws_handle *map) : h_(map) {}
  void prepare_registration(const std::vector<rmagine::Pointi> &points)          // registration.cu:303-308
  {
    ws_check(h_, ws_reg_prepare(h_, reinterpret_cast<const ws_point *>(points.data()), (int64_t)points.size()), "prepare_registration");
  }
  // registration.cu:347-368; H column-major int64[36] (rmagine::Matrix6x6l storage), g int64[6]
  void perform_registration(const Matrix4f *pretransform, int64_t H[36], int64_t g[6], int &e, int &c, int map_resolution)
  {
    int32_t ee = 0, cc = 0;
    ws_check(h_, ws_reg_step(h_, pretransform->m, map_resolution, H, g, &ee, &cc), "perform_registration");
    e = ee; c = cc;
  }
  ws_handle *h_;
};

// tsdf_mapping.h:17-60 without ROS.  With a pose buffer the constructor starts the reference's map-shift thread
// (tsdf_mapping.cpp:13-28,97-136): poses pushed by update_tsdf_from_ros (or the caller) are popped there and, once
// the sensor has moved `map.shift` metres, the local map is shifted -- ON THE DEVICE (ws_shift) instead of the
// reference's D2H / host shift / H2D round trip -- with is_shifting() true meanwhile and shifted() true afterwards.
class TSDFMapping
{
public:
  TSDFMapping(const Params &params, ConcurrentRingBuffer<Matrix4f>::Ptr &pose_buffer, HostLocalMap::Ptr &local_map, int device = 0)
      : pose_buffer_(pose_buffer), params_(params), hdf5_local_map_(local_map), cuda_map_(local_map),
        tsdf_(new TSDFCuda(cuda_map_, params.map.tau, params.map.max_weight, params.map.resolution, device)),
        last_shift_pose_(Matrix4f::Identity())
  {
    shifted_ = false;
    is_shifting_ = false;
    map_shift_run_cond_ = true;
    map_shift_thread_ = std::thread(&TSDFMapping::map_shift_loop, this);          // tsdf_mapping.cpp:27
  }
  explicit TSDFMapping(const Params &params, HostLocalMap::Ptr &local_map, int device = 0)
      : params_(params), hdf5_local_map_(local_map), cuda_map_(local_map),
        tsdf_(new TSDFCuda(cuda_map_, params.map.tau, params.map.max_weight, params.map.resolution, device)),
        last_shift_pose_(Matrix4f::Identity())
  {
    shifted_ = false;
    is_shifting_ = false;
    map_shift_run_cond_ = true;
  }
  virtual ~TSDFMapping() { join_mapping_thread(); }                // tsdf_mapping.cpp:43-48

  void join_mapping_thread()                                        // tsdf_mapping.cpp:50-60
  {
    map_shift_run_cond_ = false;
    if (pose_buffer_ != nullptr && map_shift_thread_.joinable())
    {
      pose_buffer_->clear();
      map_shift_thread_.join();
    }
  }

  void update_tsdf(const std::vector<rmagine::Pointi> &scan_points, const rmagine::Pointi &pos_rm, const rmagine::Pointi &up_rm)
  {
    std::unique_lock<std::shared_mutex> lock(mutex_);              // tsdf_mapping.cpp:71-75
    tsdf_->update_tsdf(scan_points, pos_rm, up_rm);
  }
  void update_tsdf(const std::vector<rmagine::Pointi> &scan_points, const Matrix4f &pose)
  {
    rmagine::Pointi pos_rm, up_rm;                                 // tsdf_mapping.cpp:62-69
    convert_pose_to_gpu(pose, pos_rm, up_rm);
    std::unique_lock<std::shared_mutex> lock(mutex_);
    tsdf_->update_tsdf(scan_points, pos_rm, up_rm);
  }
  void update_tsdf(DeviceMap &result, const std::vector<rmagine::Pointi> &scan_points, const Matrix4f &pose)
  {
    rmagine::Pointi pos_rm, up_rm;                                 // tsdf_mapping.cpp:87-95
    convert_pose_to_gpu(pose, pos_rm, up_rm);
    std::unique_lock<std::shared_mutex> lock(mutex_);
    tsdf_->update_tsdf(result, scan_points, pos_rm, up_rm);
  }
  // tsdf_mapping.cpp:145-163 without PCL types: `xyz_m` = n points in metres, `point_step` bytes apart (the x/y/z
  // floats of a pcl::PointXYZI cloud: 16), `pose_m` = column-major 4x4 Isometry3d in metres.  The voxel grid at the
  // map resolution runs on the device; points_rm receives the millimetre points, mm_pose the Matrix4f pose.
  void preprocess_from_ros(const float *xyz_m, int64_t n, int point_step, const double pose_m[16],
                           std::vector<rmagine::Pointi> &points_rm, Matrix4f &mm_pose)
  {
    points_rm.resize((size_t)std::max<int64_t>(n, 1));
    int64_t m = 0;
    {
      std::unique_lock<std::shared_mutex> lock(mutex_);
      ws_check(tsdf_->device_map(),
               ws_voxelgrid_subsample(tsdf_->device_map(), xyz_m, n, point_step, 0, (float)params_.map.resolution / 1000.f,
                                      reinterpret_cast<ws_point *>(points_rm.data()), nullptr, &m), "preprocess_from_ros");
    }
    points_rm.resize((size_t)m);
    mm_pose = mm_pose_from_isometry(pose_m);
  }
  // tsdf_mapping.cpp:165-173: preprocess_from_ros, pose into the shift thread's buffer, update_tsdf -- the points
  // never leave the device (ws_update_tsdf_from_ros).  Returns the number of points the voxel grid left.
  int64_t update_tsdf_from_ros(const float *xyz_m, int64_t n, int point_step, const double pose_m[16])
  {
    Matrix4f mm_pose = mm_pose_from_isometry(pose_m);
    if (pose_buffer_) pose_buffer_->push_nb(mm_pose, true);        // :170
    int64_t m = 0;
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(),
             ws_update_tsdf_from_ros(tsdf_->device_map(), xyz_m, n, point_step, 0, pose_m, nullptr, &m), "update_tsdf_from_ros");
    return m;
  }
  // the same with a result map (tsdf_mapping.cpp:175-184)
  int64_t update_tsdf_from_ros(DeviceMap &result, const float *xyz_m, int64_t n, int point_step, const double pose_m[16])
  {
    const int64_t m = update_tsdf_from_ros(xyz_m, n, point_step, pose_m);
    std::shared_lock<std::shared_mutex> lock(mutex_);
    tsdf_->avg_map().to_host(result);
    return m;
  }
  // pcl::VoxelGrid with a cubic leaf on the device (featsense's `subsample`, mapping.cpp:70,126): float centroids
  void subsample(const float *xyz_m, int64_t n, int point_step, float leaf_m, std::vector<float> &out_xyz)
  {
    out_xyz.resize((size_t)std::max<int64_t>(n, 1) * 3);
    std::vector<ws_point> mm((size_t)std::max<int64_t>(n, 1));
    int64_t m = 0;
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(),
             ws_voxelgrid_subsample(tsdf_->device_map(), xyz_m, n, point_step, 0, leaf_m, mm.data(), out_xyz.data(), &m), "subsample");
    out_xyz.resize((size_t)m * 3);
  }
  // Eigen::Quaterniond(pose.rotation()).toRotationMatrix().cast<float>(), translation * 1000 (tsdf_mapping.cpp:160-161)
  static Matrix4f mm_pose_from_isometry(const double P[16])
  {
    auto m = [&](int r, int c) { return P[c * 4 + r]; };
    double q[4];
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0)
    {
      t = std::sqrt(t + 1.0);
      q[3] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
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
      q[3] = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
    }
    const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    Matrix4f out = Matrix4f::Identity();
    out(0, 0) = (float)(1.0 - (tyy + tzz)); out(0, 1) = (float)(txy - twz); out(0, 2) = (float)(txz + twy);
    out(1, 0) = (float)(txy + twz); out(1, 1) = (float)(1.0 - (txx + tzz)); out(1, 2) = (float)(tyz - twx);
    out(2, 0) = (float)(txz - twy); out(2, 1) = (float)(tyz + twx); out(2, 2) = (float)(1.0 - (txx + tyy));
    for (int r = 0; r < 3; r++) out(r, 3) = (float)(P[12 + r] * 1000.0);
    return out;
  }
  // tsdf_mapping.cpp:77-85
  void convert_pose_to_gpu(const Matrix4f &pose, rmagine::Pointi &pos_rm, rmagine::Pointi &up_rm) const
  {
    int M[16], R[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };   // Matrix4i::Identity() ...
    warpsense_b200::to_int_mat(pose, M);
    for (int c = 0; c < 3; c++)                                               // ... with the int rotation block
      for (int r = 0; r < 3; r++) R[c * 4 + r] = M[c * 4 + r];
    up_rm = warpsense_b200::transform_point(rmagine::Pointi(0, 0, warpsense_b200::MATRIX_RESOLUTION), R);
    pos_rm = warpsense_b200::to_map(pose, params_.map.resolution);
  }
  // one turn of the map-shift loop body (tsdf_mapping.cpp:104-125), the shift itself running on the device
  bool map_shift(const Matrix4f &current_pose)
  {
    float d2 = 0.f;
    for (int a = 0; a < 3; a++)
    {
      const float d = last_shift_pose_.m[12 + a] / 1000.f - current_pose.m[12 + a] / 1000.f;
      d2 += d * d;
    }
    if (std::sqrt(d2) < params_.map.shift) return false;
    is_shifting_ = true;
    last_shift_pose_ = current_pose;
    if (shift_hook_) shift_hook_();                                // tests: something to do while is_shifting()
    const rmagine::Pointi p = warpsense_b200::to_map(current_pose, params_.map.resolution);
    {
      std::unique_lock<std::shared_mutex> lock(mutex_);
      const int32_t np[3] = { p.x, p.y, p.z };
      ws_check(tsdf_->device_map(), ws_shift(tsdf_->device_map(), np), "map_shift");
      int32_t s[3];
      ws_map_get_params(tsdf_->device_map(), s, hdf5_local_map_->get_offset(), hdf5_local_map_->get_pos());
    }
    shifted_ = true;
    is_shifting_ = false;
    return true;
  }
  void set_shift_hook(std::function<void()> f) { shift_hook_ = std::move(f); }
  void get_tsdf_map()                                              // tsdf_mapping.cpp:138-143
  {
    std::shared_lock<std::shared_mutex> lock(mutex_);
    tsdf_->avg_map().to_host(cuda_map_);
  }
  const std::unique_ptr<TSDFCuda> &tsdf() const { return tsdf_; }
  std::unique_ptr<TSDFCuda> &tsdf() { return tsdf_; }
  std::atomic<bool> &shifted() { return shifted_; }
  std::atomic<bool> &is_shifting() { return is_shifting_; }
  const Params &params() const { return params_; }

protected:
  // tsdf_mapping.cpp:97-136: pop poses, shift when the sensor has moved far enough, sleep 2 ms otherwise
  void map_shift_loop()
  {
    Matrix4f current_pose = Matrix4f::Identity();
    while (map_shift_run_cond_)
    {
      if (pose_buffer_->pop_nb(&current_pose))
      {
        try { map_shift(current_pose); }
        catch (const std::exception &e) { std::fprintf(stderr, "map_shift: %s\n", e.what()); is_shifting_ = false; }
      }
      else std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  ConcurrentRingBuffer<Matrix4f>::Ptr pose_buffer_;
  std::thread map_shift_thread_;
  std::atomic<bool> map_shift_run_cond_;
  std::function<void()> shift_hook_;
  const Params &params_;
  HostLocalMap::Ptr &hdf5_local_map_;
  DeviceMap cuda_map_;
  std::unique_ptr<TSDFCuda> tsdf_;
  std::shared_mutex mutex_;
  std::atomic<bool> shifted_, is_shifting_;
  Matrix4f last_shift_pose_;
};

// tsdf_registration.h:17-27
struct TSDFRegistration : public TSDFMapping
{
  explicit TSDFRegistration(const Params &params, HostLocalMap::Ptr &local_map, int device = 0)
      : TSDFMapping(params, local_map, device), reg_(new RegistrationCuda(tsdf_->device_map())) {}
  TSDFRegistration(const Params &params, ConcurrentRingBuffer<Matrix4f>::Ptr &pose_buffer, HostLocalMap::Ptr &local_map, int device = 0)
      : TSDFMapping(params, pose_buffer, local_map, device), reg_(new RegistrationCuda(tsdf_->device_map())) {}

  // tsdf_registration.cpp:29-96, following the CPU loop (src/cpu/registration.cpp:14-177): the cloud is
  // transformed in place and the total transform returned.  All iterations run in one persistent kernel.
  Matrix4f register_cloud(std::vector<rmagine::Pointi> &cloud, const Matrix4f &pretransform, int *iterations = nullptr)
  {
    Matrix4f out{};
    int32_t it = 0;
    std::shared_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(),
             ws_register_cloud(tsdf_->device_map(), reinterpret_cast<ws_point *>(cloud.data()), (int64_t)cloud.size(), pretransform.m,
                               params_.registration.max_iterations, params_.registration.it_weight_gradient,
                               params_.registration.epsilon, params_.map.resolution, WS_REG_DEVICE_SOLVE, out.m, &it),
             "register_cloud");
    if (iterations) *iterations = it;
    return out;
  }

  // App::preprocess (src/warpsense/app.cpp:118-148) on the device: the x/y/z floats of a PointCloud2 payload
  // (metres, `point_step` bytes apart) -> unique map-frame points, left on the device for track_scan /
  // ws_update_tsdf_device and copied into `scan_points` as well.
  void preprocess(const float *xyz, int64_t n, int point_step, const Matrix4f &pose, std::vector<rmagine::Pointi> &scan_points)
  {
    scan_points.resize((size_t)std::max<int64_t>(n, 1));
    int64_t m = 0;
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(),
             ws_preprocess_scan(tsdf_->device_map(), xyz, n, point_step, 0, pose.m, params_.map.resolution,
                                reinterpret_cast<ws_point *>(scan_points.data()), &m), "preprocess");
    scan_points.resize((size_t)m);
  }

  // App::cloud_callback's per-scan sequence (app.cpp:65-112) as one stream of kernels and one host
  // synchronisation: register_cloud(identity) -> pose = X * prior_pose -> update_tsdf(registered cloud, pose).
  // Returns the new pose; `transform` (optional) receives X.
  Matrix4f track_scan(const std::vector<rmagine::Pointi> &cloud, const Matrix4f &prior_pose, Matrix4f *transform = nullptr,
                      int *iterations = nullptr)
  {
    Matrix4f X{}, pose{};
    int32_t it = 0;
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(),
             ws_track_scan(tsdf_->device_map(), reinterpret_cast<const ws_point *>(cloud.data()), (int64_t)cloud.size(), 0,
                           prior_pose.m, params_.registration.max_iterations, params_.registration.it_weight_gradient,
                           params_.registration.epsilon, params_.map.resolution, X.m, pose.m, &it),
             "track_scan");
    if (transform) *transform = X;
    if (iterations) *iterations = it;
    return pose;
  }

  // The same, asynchronous (ws_track_submit / ws_track_wait): at most two scans in flight; `cloud` must stay alive
  // until track_wait.  prior_pose == nullptr chains from the pose of the previous tracked scan on the device;
  // reference_pose composes the pose as App::update_pose_estimate does (app.cpp:172-176).
  int track_submit(const std::vector<rmagine::Pointi> &cloud, const Matrix4f *prior_pose, const Matrix4f *pretransform = nullptr,
                   bool reference_pose = false)
  {
    int32_t ticket = -1;
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(),
             ws_track_submit(tsdf_->device_map(), reinterpret_cast<const ws_point *>(cloud.data()), (int64_t)cloud.size(), 0,
                             prior_pose ? prior_pose->m : nullptr, pretransform ? pretransform->m : nullptr,
                             params_.registration.max_iterations, params_.registration.it_weight_gradient,
                             params_.registration.epsilon, params_.map.resolution, reference_pose ? WS_TRACK_REFERENCE_POSE : 0, &ticket),
             "track_submit");
    return ticket;
  }
  Matrix4f track_wait(int ticket, Matrix4f *transform = nullptr, int *iterations = nullptr)
  {
    Matrix4f X{}, pose{};
    int32_t it = 0;
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(), ws_track_wait(tsdf_->device_map(), ticket, X.m, pose.m, &it), "track_wait");
    if (transform) *transform = X;
    if (iterations) *iterations = it;
    return pose;
  }

  // HDF5GlobalMap::write_back + the file it leaves (hdf5_global_map.cpp:140-221): local map into the chunk
  // store, then chunks, /map attributes and poses (rows x y z qx qy qz qw) into an HDF5 file.
  void export_map(const std::string &path, const std::vector<float> &poses7 = {})
  {
    std::unique_lock<std::shared_mutex> lock(mutex_);
    ws_check(tsdf_->device_map(), ws_write_back(tsdf_->device_map()), "write_back");
    ws_map_meta meta;
    meta.tau = params_.map.tau;
    for (int a = 0; a < 3; a++) meta.map_size[a] = params_.map.size[a];
    meta.max_distance = params_.map.max_distance;
    meta.map_resolution = params_.map.resolution;
    meta.max_weight = params_.map.max_weight;
    ws_check(tsdf_->device_map(),
             ws_export_hdf5(tsdf_->device_map(), path.c_str(), &meta, poses7.data(), (int64_t)(poses7.size() / 7)), "export_map");
  }
  std::unique_ptr<RegistrationCuda> reg_;
};
}  // namespace cuda

// The TSDF-facing half of featsense's Mapping::thread_run (src/featsense/mapping.cpp:39-147) without ROS / PCL:
// F-LOAM and VGICP stay outside -- the caller hands over the cloud already placed in the map
// frame (metres) and its pose (column-major 4x4 Isometry3d, metres).
//   * the first cloud initialises the map: subsample, update_tsdf_from_ros (:66-75);
//   * later clouds are used only after the sensor moved more than map.update_distance (:79-81);
//   * while the local map is shifting (TSDFMapping's thread) the clouds are accumulated, and flushed --
//     concatenated and subsampled at the map resolution -- with the next update (:115-129);
//   * every used pose is recorded (what hdf5_global_map_->write_pose would store, :137) and reaches the shift
//     thread through update_tsdf_from_ros' pose buffer (:140 pushes it a second time: update_map_shift).
namespace featsense
{
class MappingFeed
{
public:
  explicit MappingFeed(cuda::TSDFMapping &gpu) : gpu_(gpu) {}
  // one (cloud, pose) pair; xyz_m: n points, 3 floats each.  Returns true if the map was updated.
  bool push(const float *xyz_m, int64_t n, const double pose_m[16])
  {
    const float resolution = (float)gpu_.params().map.resolution / 1000.f;
    if (!initialized_)
    {
      std::memcpy(last_pose_, pose_m, sizeof(last_pose_));
      std::vector<float> sub;
      gpu_.subsample(xyz_m, n, 12, resolution, sub);
      gpu_.update_tsdf_from_ros(sub.data(), (int64_t)(sub.size() / 3), 12, pose_m);
      initialized_ = true;
      updates_++;
      return true;
    }
    double d2 = 0.0;
    for (int a = 0; a < 3; a++) { const double d = last_pose_[12 + a] - pose_m[12 + a]; d2 += d * d; }
    if (!(std::sqrt(d2) > (double)gpu_.params().map.update_distance)) return false;
    bool updated = false;
    if (gpu_.is_shifting())
    {
      accumulated_.insert(accumulated_.end(), xyz_m, xyz_m + 3 * n);                           // :117
    }
    else
    {
      if (!accumulated_.empty())
      {
        std::vector<float> all(xyz_m, xyz_m + 3 * n), sub;
        all.insert(all.end(), accumulated_.begin(), accumulated_.end());                       // :123
        accumulated_.clear();
        gpu_.subsample(all.data(), (int64_t)(all.size() / 3), 12, resolution, sub);            // :125
        gpu_.update_tsdf_from_ros(sub.data(), (int64_t)(sub.size() / 3), 12, pose_m);
      }
      else gpu_.update_tsdf_from_ros(xyz_m, n, 12, pose_m);                                    // :127
      updates_++;
      updated = true;
    }
    std::memcpy(last_pose_, pose_m, sizeof(last_pose_));
    poses_.insert(poses_.end(), pose_m, pose_m + 16);
    return updated;
  }
  size_t accumulated_points() const { return accumulated_.size() / 3; }
  int updates() const { return updates_; }
  const std::vector<double> &poses() const { return poses_; }

private:
  cuda::TSDFMapping &gpu_;
  bool initialized_ = false;
  double last_pose_[16] = {};
  std::vector<float> accumulated_;
  std::vector<double> poses_;
  int updates_ = 0;
};
}  // namespace featsense
