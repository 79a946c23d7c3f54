// voxelgrid.cu -- the voxel-grid subsample of the featsense feed on the device.
//
// Replaces the pcl::VoxelGrid<pcl::PointXYZI> of cuda::TSDFMapping::preprocess_from_ros
// (src/warpsense/tsdf_mapping.cpp:145-159 under /root/reference: leaf = map resolution in metres, then
// `Pointi(cp.x * 1000.f, ...)`).  PCL is a third-party dependency that is not vendored in the reference tree
// (ROS noetic ships PCL 1.10); what is restated here is its published algorithm, filters/impl/voxel_grid.hpp
// `VoxelGrid<PointT>::applyFilter`:
//   * min/max of the finite points; min_b = floor(min * inverse_leaf), max_b likewise, div_b = max_b - min_b + 1,
//     inverse_leaf = 1.f / leaf (float);
//   * per point ijk = (int)(floor(x * inverse_leaf) - (float)min_b), idx = ijk0 + ijk1 * div_b0 + ijk2 * div_b0 * div_b1;
//   * the (idx, point index) pairs are sorted by idx; every run of equal idx becomes one output point, in
//     ascending idx (x fastest), = the float sum of its members divided by their number (CentroidPoint /
//     AccumulatorXYZ: Eigen::Vector3f accumulation, `xyz / n`);
//   * if (dx * dy * dz) of the bounding box exceeds INT_MAX the filter gives up and returns the input cloud.
// PCL sorts with std::sort, which leaves the order INSIDE a run unspecified; here (and in the oracle) a run is
// summed in ascending point index -- the order a stable sort gives.
//
// Kernels: min/max reduction, key computation, a stable 4 x 8-bit LSD radix sort of (idx, point index), run
// heads + ordered compaction, one thread per output point for the centroid.
#include <cfloat>
#include "ws_internal.h"

#define FULL 0xFFFFFFFFu
#define VG_TILE 256
#define VG_INVALID 0xFFFFFFFFu

namespace {

struct VgBox
{
  unsigned enc_min[3], enc_max[3];      // order-preserving integer encodings of the float extrema
  unsigned n_finite;
  unsigned n_cells;
  int min_b[3], div_b[3];
  int passthrough;                      // the leaf is too small for the extent: PCL returns its input
};

WS_D unsigned enc_f(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
WS_D float dec_f(unsigned e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e); }

__global__ void __launch_bounds__(256)
vg_minmax_kernel(const float *__restrict__ xyz, const int n, const int stride, VgBox *__restrict__ box)
{
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  unsigned cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float p[3] = { xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2] };
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;        // getMinMax3D, !is_dense
    cnt++;
#pragma unroll
    for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], p[a]); mx[a] = fmaxf(mx[a], p[a]); }
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    cnt += __shfl_down_sync(FULL, cnt, o);
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
      mn[a] = fminf(mn[a], __shfl_down_sync(FULL, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_down_sync(FULL, mx[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0 && cnt)
  {
    atomicAdd(&box->n_finite, cnt);
#pragma unroll
    for (int a = 0; a < 3; a++) { atomicMin(&box->enc_min[a], enc_f(mn[a])); atomicMax(&box->enc_max[a], enc_f(mx[a])); }
  }
}

__global__ void vg_box_kernel(VgBox *__restrict__ box, const float inv_leaf)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  box->passthrough = 0;
  if (box->n_finite == 0u) { for (int a = 0; a < 3; a++) { box->min_b[a] = 0; box->div_b[a] = 1; } return; }
  long long d[3];
  for (int a = 0; a < 3; a++)
  {
    const float mn = dec_f(box->enc_min[a]), mx = dec_f(box->enc_max[a]);
    d[a] = (long long)((mx - mn) * inv_leaf) + 1;                              // voxel_grid.hpp: dx, dy, dz
    box->min_b[a] = (int)floorf(mn * inv_leaf);
    box->div_b[a] = (int)floorf(mx * inv_leaf) - box->min_b[a] + 1;
  }
  if (d[0] * d[1] * d[2] > 0x7FFFFFFFll) box->passthrough = 1;
}

__global__ void __launch_bounds__(256)
vg_key_kernel(const float *__restrict__ xyz, const int n, const int stride, const float inv_leaf, const VgBox *__restrict__ box,
              unsigned *__restrict__ keys, unsigned *__restrict__ vals)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p[3] = { xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2] };
  unsigned key = VG_INVALID;
  if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))
  {
    if (box->passthrough) key = 0x7FFFFFFFu;     // one run holding every point is not what PCL does: handled by the caller
    else
    {
      const int i0 = (int)(floorf(p[0] * inv_leaf) - (float)box->min_b[0]);
      const int i1 = (int)(floorf(p[1] * inv_leaf) - (float)box->min_b[1]);
      const int i2 = (int)(floorf(p[2] * inv_leaf) - (float)box->min_b[2]);
      key = (unsigned)(i0 + i1 * box->div_b[0] + i2 * box->div_b[0] * box->div_b[1]);
    }
  }
  keys[i] = key;
  vals[i] = (unsigned)i;
}

// ---- stable LSD radix sort, 8 bits per pass: per-tile digit histogram, digit-major scan, ordered scatter ----
__global__ void __launch_bounds__(VG_TILE)
vg_hist_kernel(const unsigned *__restrict__ keys, const int n, const int shift, unsigned *__restrict__ hist, const int n_tiles)
{
  __shared__ unsigned s_h[256];
  s_h[threadIdx.x] = 0u;
  __syncthreads();
  const int i = blockIdx.x * VG_TILE + threadIdx.x;
  if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of `count` values in place (one CTA of 1024 threads)
__global__ void __launch_bounds__(1024)
vg_scan_kernel(unsigned *__restrict__ data, const size_t count, unsigned *__restrict__ total)
{
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0u;
  __syncthreads();
  for (size_t base = 0; base < count; base += 1024)
  {
    const size_t gi = base + (size_t)tid;
    const unsigned v = gi < count ? data[gi] : 0u;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
      unsigned w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(FULL, w, o); if (lane >= o) w += y; }
      s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned incl = s_carry + x + (warp > 0 ? s_warp[warp - 1] : 0u);
    if (gi < count) data[gi] = incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = incl;
    __syncthreads();
  }
  if (tid == 0 && total) *total = s_carry;
}

__global__ void __launch_bounds__(VG_TILE)
vg_scatter_kernel(const unsigned *__restrict__ keys_in, const unsigned *__restrict__ vals_in, const int n, const int shift,
                  const unsigned *__restrict__ hist, const int n_tiles, unsigned *__restrict__ keys_out,
                  unsigned *__restrict__ vals_out)
{
  __shared__ unsigned s_d[VG_TILE];
  const int i = blockIdx.x * VG_TILE + threadIdx.x;
  const unsigned key = i < n ? keys_in[i] : 0u;
  const unsigned d = i < n ? ((key >> shift) & 255u) : 256u;
  s_d[threadIdx.x] = d;
  __syncthreads();
  if (i >= n) return;
  unsigned rank = 0;                       // equal digits before me in the tile: keeps the sort stable
  for (int j = 0; j < (int)threadIdx.x; j++) rank += s_d[j] == d ? 1u : 0u;
  const unsigned dst = hist[(size_t)d * n_tiles + blockIdx.x] + rank;
  keys_out[dst] = key;
  vals_out[dst] = vals_in[i];
}

// run heads -> 1 / 0 (then scanned), the last element of `flags` doubles as the sentinel slot
__global__ void __launch_bounds__(256)
vg_heads_kernel(const unsigned *__restrict__ keys, const int n, unsigned *__restrict__ flags)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned k = keys[i];
  flags[i] = (k != VG_INVALID && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
vg_starts_kernel(const unsigned *__restrict__ keys, const unsigned *__restrict__ pos, const int n, unsigned *__restrict__ starts)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned k = keys[i];
  if (k != VG_INVALID && (i == 0 || keys[i - 1] != k)) starts[pos[i]] = (unsigned)i;
}

// one thread per output point: float sum of the run in ascending point index, / n, metres -> int millimetres
__global__ void __launch_bounds__(256)
vg_centroid_kernel(const float *__restrict__ xyz, const int stride, const unsigned *__restrict__ keys,
                   const unsigned *__restrict__ vals, const int n, const unsigned *__restrict__ starts,
                   const unsigned *__restrict__ n_cells_p, float *__restrict__ out_xyz, ws_pt *__restrict__ out_mm)
{
  const unsigned n_cells = *n_cells_p;
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const unsigned first = starts[c];
  const unsigned key = keys[first];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  unsigned m = 0;
  for (unsigned j = first; j < (unsigned)n && keys[j] == key; j++)
  {
    const size_t p = (size_t)vals[j] * stride;
    sx += xyz[p]; sy += xyz[p + 1]; sz += xyz[p + 2];                       // AccumulatorXYZ::add
    m++;
  }
  const float fm = (float)m;
  const float cx = sx / fm, cy = sy / fm, cz = sz / fm;                     // AccumulatorXYZ::get: xyz / n
  if (out_xyz) { out_xyz[3 * (size_t)c] = cx; out_xyz[3 * (size_t)c + 1] = cy; out_xyz[3 * (size_t)c + 2] = cz; }
  ws_pt q;
  q.x = (int)(cx * 1000.f); q.y = (int)(cy * 1000.f); q.z = (int)(cz * 1000.f);   // tsdf_mapping.cpp:156-157
  out_mm[c] = q;
}

// passthrough (leaf too small for the extent): PCL hands the input on unchanged
__global__ void __launch_bounds__(256)
vg_copy_kernel(const float *__restrict__ xyz, const int stride, const int n, float *__restrict__ out_xyz, ws_pt *__restrict__ out_mm)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
  if (out_xyz) { out_xyz[3 * (size_t)i] = x; out_xyz[3 * (size_t)i + 1] = y; out_xyz[3 * (size_t)i + 2] = z; }
  ws_pt q; q.x = (int)(x * 1000.f); q.y = (int)(y * 1000.f); q.z = (int)(z * 1000.f);
  out_mm[i] = q;
}

}  // namespace

// d_xyz: n points, `stride` floats apart (device memory).  Result: int millimetre points in h->d_points (the
// update_tsdf staging buffer) and, if d_out_xyz != nullptr, the float centroids.  Returns the number of points.
int64_t ws_launch_voxelgrid(ws_handle *h, const float *d_xyz, int64_t n, int stride, float leaf_m, float *d_out_xyz)
{
  if (n <= 0) return 0;
  cudaStream_t s = h->stream;
  const int n_tiles = (int)((n + VG_TILE - 1) / VG_TILE);
  if ((size_t)n > h->vg_cap)
  {
    WS_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(h->d_vg_keys); cudaFree(h->d_vg_vals); cudaFree(h->d_vg_hist); cudaFree(h->d_vg_box);
    h->d_vg_keys = h->d_vg_vals = h->d_vg_hist = nullptr; h->d_vg_box = nullptr; h->vg_cap = 0;
    const size_t want = std::max<size_t>((size_t)n, 1 << 17);
    WS_CUDA_OK(cudaMalloc(&h->d_vg_keys, 3 * want * sizeof(unsigned)));      // ping, pong, run starts
    WS_CUDA_OK(cudaMalloc(&h->d_vg_vals, 3 * want * sizeof(unsigned)));      // ping, pong, head flags / positions
    WS_CUDA_OK(cudaMalloc(&h->d_vg_hist, 256 * ((want + VG_TILE - 1) / VG_TILE) * sizeof(unsigned)));
    WS_CUDA_OK(cudaMalloc(&h->d_vg_box, sizeof(VgBox)));
    h->vg_cap = want;
  }
  const size_t cap = h->vg_cap;
  unsigned *keys[2] = { h->d_vg_keys, h->d_vg_keys + cap }, *starts = h->d_vg_keys + 2 * cap;
  unsigned *vals[2] = { h->d_vg_vals, h->d_vg_vals + cap }, *flags = h->d_vg_vals + 2 * cap;
  VgBox *box = static_cast<VgBox *>(h->d_vg_box);
  const float inv_leaf = 1.f / leaf_m;                                       // inverse_leaf_size_ (float)

  VgBox init{};
  for (int a = 0; a < 3; a++) { init.enc_min[a] = 0xFFFFFFFFu; init.enc_max[a] = 0u; }
  WS_CUDA_OK(cudaMemcpyAsync(box, &init, sizeof(VgBox), cudaMemcpyHostToDevice, s));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
  vg_minmax_kernel<<<grid, 256, 0, s>>>(d_xyz, (int)n, stride, box);
  vg_box_kernel<<<1, 32, 0, s>>>(box, inv_leaf);
  vg_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_xyz, (int)n, stride, inv_leaf, box, keys[0], vals[0]);
  int cur = 0;
  for (int pass = 0; pass < 4; pass++)
  {
    vg_hist_kernel<<<n_tiles, VG_TILE, 0, s>>>(keys[cur], (int)n, 8 * pass, h->d_vg_hist, n_tiles);
    vg_scan_kernel<<<1, 1024, 0, s>>>(h->d_vg_hist, (size_t)256 * n_tiles, nullptr);
    vg_scatter_kernel<<<n_tiles, VG_TILE, 0, s>>>(keys[cur], vals[cur], (int)n, 8 * pass, h->d_vg_hist, n_tiles, keys[cur ^ 1], vals[cur ^ 1]);
    cur ^= 1;
  }
  vg_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys[cur], (int)n, flags);
  vg_scan_kernel<<<1, 1024, 0, s>>>(flags, (size_t)n, &box->n_cells);
  vg_starts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys[cur], flags, (int)n, starts);
  vg_centroid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_xyz, stride, keys[cur], vals[cur], (int)n, starts, &box->n_cells,
                                                                 d_out_xyz, h->d_points);
  h->launches += 19;
  VgBox hb;
  WS_CUDA_OK(cudaMemcpyAsync(&hb, box, sizeof(VgBox), cudaMemcpyDeviceToHost, s));
  WS_CUDA_OK(cudaStreamSynchronize(s));
  WS_CUDA_OK(cudaGetLastError());
  if (hb.passthrough)
  {
    vg_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_xyz, stride, (int)n, d_out_xyz, h->d_points);
    h->launches++;
    WS_CUDA_OK(cudaStreamSynchronize(s));
    return n;
  }
  return (int64_t)hb.n_cells;
}
