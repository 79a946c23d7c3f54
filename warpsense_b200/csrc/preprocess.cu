// preprocess.cu -- scan preprocessing on the device.
//
// Replaces App::preprocess (src/warpsense/app.cpp:118-148 under /root/reference): per return of the
// PointCloud2 -- drop it if x, y and z are all below 0.3 m (:128-131), metres -> millimetres (:133), snap to
// the centre of its voxel (:134-139, float arithmetic), transform into the map frame with the fixed-point
// pose (:141, include/util/util.h:8-18) -- and the std::unordered_set that removes duplicates (:120,:141).
// The reference copies the set out in its (implementation-defined) bucket order (:144-145); here the
// survivors keep SCAN ORDER: a point stays iff it is the first occurrence of its value.  Same set,
// deterministic order (update_tsdf breaks ties between equal candidates by point index).
//
//   1. insert_kernel   transform + open-addressing hash set keyed by the point value; a slot holds the
//                      SMALLEST index that carries its value (atomicCAS to claim, atomicMin to lower)
//   2. count_kernel    a point survives iff its slot still names it; survivors per 1024-point tile
//   3. scatter_kernel  tile offsets (a few hundred tiles: summed on the fly), ordered compaction
#include "ws_internal.h"

#define FULL 0xFFFFFFFFu
#define PRE_TILE 1024
#define PRE_EMPTY 0xFFFFFFFFu

namespace {

struct PreParams
{
  int M[16];          // to_int_mat(pose), column-major
  int res;
  int n;
  int stride_floats;  // floats between consecutive points (PointCloud2 point_step / 4)
  unsigned mask;      // table size - 1
  int mode;           // 0: App::preprocess (src/warpsense/app.cpp:118-148), 1: the CPU node (src/cpu/fastsense.cpp:143-163)
};

WS_D int pre_div_mr(int x) { return (x + ((x >> 31) & (WS_MR - 1))) >> WS_MR_SHIFT; }

// one return -> (key the duplicates are detected by, point handed on).  Mode 0: both are the transformed voxel
// centre.  Mode 1 (fastsense.cpp:149-163): the key is the voxel of the untransformed millimetre point
// (`point / MAP_RESOLUTION`, truncating), the point handed on is transform_point(point) itself.
WS_D bool pre_point(const PreParams &P, const float *__restrict__ xyz, int i, ws_pt &key, ws_pt &out)
{
  const float x = xyz[(size_t)i * P.stride_floats], y = xyz[(size_t)i * P.stride_floats + 1],
              z = xyz[(size_t)i * P.stride_floats + 2];
  if (x < 0.3 && y < 0.3 && z < 0.3) return false;                       // app.cpp:128-131 / fastsense.cpp:152-155
  // NaN / inf returns (PointCloud2 with is_dense == false): the reference casts them to int -- undefined behaviour,
  // INT_MIN on x86, normally out of bounds and dropped later; here they are dropped explicitly (the oracle does the same)
  if (!isfinite(x) || !isfinite(y) || !isfinite(z)) return false;
  int c[3];
  if (P.mode == 0)
  {
    const float mm[3] = { x * 1000.f, y * 1000.f, z * 1000.f };           // app.cpp:133
#pragma unroll
    for (int a = 0; a < 3; a++)                                           // :134-139: float / int -> float
      c[a] = (int)(floorf(mm[a] / (float)P.res) * (float)P.res + (float)(P.res / 2));
  }
  else
  {
    c[0] = (int)(x * 1000); c[1] = (int)(y * 1000); c[2] = (int)(z * 1000);   // fastsense.cpp:156-159 (float * int)
  }
  int q[3];
#pragma unroll
  for (int r = 0; r < 3; r++)                                             // util.h:13-18
  {
    const unsigned acc = (unsigned)P.M[r] * (unsigned)c[0] + (unsigned)P.M[4 + r] * (unsigned)c[1] +
                         (unsigned)P.M[8 + r] * (unsigned)c[2] + (unsigned)P.M[12 + r];
    q[r] = pre_div_mr((int)acc);
  }
  out.x = q[0]; out.y = q[1]; out.z = q[2];
  if (P.mode == 0) key = out;
  else { key.x = c[0] / P.res; key.y = c[1] / P.res; key.z = c[2] / P.res; }   // fastsense.cpp:160 (Eigen int division)
  return true;
}

WS_D unsigned pre_hash(const ws_pt p)
{
  unsigned h = (unsigned)p.x * 0x9E3779B1u;
  h = (h ^ (h >> 15)) + (unsigned)p.y * 0x85EBCA77u;
  h = (h ^ (h >> 13)) + (unsigned)p.z * 0xC2B2AE3Du;
  return h ^ (h >> 16);
}

__global__ void __launch_bounds__(256)
insert_kernel(const PreParams P, const float *__restrict__ xyz, ws_pt *__restrict__ tmp, ws_pt *__restrict__ vals,
              unsigned *__restrict__ slot_of, unsigned *__restrict__ table)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  ws_pt p, v;
  if (!pre_point(P, xyz, i, p, v)) { slot_of[i] = PRE_EMPTY; return; }
  tmp[i] = p;
  vals[i] = v;
  __threadfence();                       // the point must be readable by whoever finds index i in a slot
  unsigned h = pre_hash(p) & P.mask;
  for (;;)
  {
    unsigned cur = *((volatile unsigned *)&table[h]);
    if (cur == PRE_EMPTY)
    {
      cur = atomicCAS(&table[h], PRE_EMPTY, (unsigned)i);
      if (cur == PRE_EMPTY) break;       // claimed
    }
    const volatile int *op = reinterpret_cast<const volatile int *>(&tmp[cur]);
    if (op[0] == p.x && op[1] == p.y && op[2] == p.z)
    {
      atomicMin(&table[h], (unsigned)i); // same value: the slot keeps the first occurrence
      break;
    }
    h = (h + 1u) & P.mask;
  }
  slot_of[i] = h;
}

__global__ void __launch_bounds__(256)
count_kernel(const int n, const unsigned *__restrict__ slot_of, const unsigned *__restrict__ table,
             unsigned *__restrict__ tile_count)
{
  // one block per tile of PRE_TILE points, 4 points per thread
  __shared__ unsigned s_n;
  if (threadIdx.x == 0) s_n = 0u;
  __syncthreads();
  unsigned c = 0;
  for (int k = 0; k < PRE_TILE / 256; k++)
  {
    const int i = blockIdx.x * PRE_TILE + k * 256 + threadIdx.x;
    if (i < n)
    {
      const unsigned s = slot_of[i];
      if (s != PRE_EMPTY && table[s] == (unsigned)i) c++;
    }
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULL, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_n, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = s_n;
}

__global__ void __launch_bounds__(256)
scatter_kernel(const int n, const int n_tiles, const ws_pt *__restrict__ tmp, const unsigned *__restrict__ slot_of,
               const unsigned *__restrict__ table, const unsigned *__restrict__ tile_count, ws_pt *__restrict__ out,
               unsigned *__restrict__ n_out)
{
  __shared__ unsigned s_warp[8];
  __shared__ unsigned s_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // offset of this tile = survivors of all earlier tiles
  unsigned before = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += 256) before += tile_count[t];
  for (int o = 16; o > 0; o >>= 1) before += __shfl_down_sync(FULL, before, o);
  if (threadIdx.x == 0) s_off = 0u;
  __syncthreads();
  if (lane == 0 && before) atomicAdd(&s_off, before);
  __syncthreads();
  unsigned base = s_off;
  for (int k = 0; k < PRE_TILE / 256; k++)
  {
    const int i = blockIdx.x * PRE_TILE + k * 256 + threadIdx.x;
    bool keep = false;
    if (i < n)
    {
      const unsigned s = slot_of[i];
      keep = s != PRE_EMPTY && table[s] == (unsigned)i;
    }
    const unsigned m = __ballot_sync(FULL, keep);
    if (lane == 0) s_warp[warp] = (unsigned)__popc(m);
    __syncthreads();
    unsigned wbase = base, total = 0;
    for (int w = 0; w < 8; w++) { if (w < warp) wbase += s_warp[w]; total += s_warp[w]; }
    if (keep) out[wbase + (unsigned)__popc(m & ((1u << lane) - 1u))] = tmp[i];
    base += total;
    __syncthreads();
  }
  if (blockIdx.x == n_tiles - 1 && threadIdx.x == 0) *n_out = base;
}

}  // namespace

// out: h->d_points (the update_tsdf staging buffer); returns the number of surviving points
int64_t ws_launch_preprocess(ws_handle *h, const float *d_xyz, int64_t n, int stride_floats, const float pose_mm[16], int res,
                             int mode)
{
  if (n <= 0) return 0;
  cudaStream_t s = h->stream;
  // scratch: transformed points, slot per point, tile counts, result count, hash table (>= 2n slots)
  size_t tab = 1024;
  while (tab < (size_t)n * 2) tab <<= 1;
  const int n_tiles = (int)((n + PRE_TILE - 1) / PRE_TILE);
  if ((size_t)n > h->pre_cap)
  {
    WS_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(h->d_pre_tmp); cudaFree(h->d_pre_val); cudaFree(h->d_pre_slot); cudaFree(h->d_pre_table); cudaFree(h->d_pre_tiles);
    h->d_pre_val = nullptr; h->d_pre_tmp = nullptr; h->d_pre_slot = nullptr; h->d_pre_table = nullptr; h->d_pre_tiles = nullptr; h->pre_cap = 0;
    const size_t want = std::max<size_t>((size_t)n, 1 << 17);
    size_t wtab = 1024;
    while (wtab < want * 2) wtab <<= 1;
    WS_CUDA_OK(cudaMalloc(&h->d_pre_tmp, want * sizeof(ws_pt)));
    WS_CUDA_OK(cudaMalloc(&h->d_pre_val, want * sizeof(ws_pt)));
    WS_CUDA_OK(cudaMalloc(&h->d_pre_slot, want * sizeof(unsigned)));
    WS_CUDA_OK(cudaMalloc(&h->d_pre_table, wtab * sizeof(unsigned)));
    WS_CUDA_OK(cudaMalloc(&h->d_pre_tiles, ((want + PRE_TILE - 1) / PRE_TILE + 1) * sizeof(unsigned)));
    h->pre_cap = want;
  }
  PreParams P;
  for (int i = 0; i < 16; i++) P.M[i] = (int)(pose_mm[i] * (float)WS_MR);    // util.h:8-11
  P.res = res; P.n = (int)n; P.stride_floats = stride_floats; P.mask = (unsigned)(tab - 1); P.mode = mode;
  ws_pt *tmp = static_cast<ws_pt *>(h->d_pre_tmp);
  ws_pt *vals = static_cast<ws_pt *>(h->d_pre_val);
  unsigned *n_out = h->d_pre_tiles + n_tiles;
  WS_CUDA_OK(cudaMemsetAsync(h->d_pre_table, 0xFF, tab * sizeof(unsigned), s));
  insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P, d_xyz, tmp, vals, h->d_pre_slot, h->d_pre_table);
  count_kernel<<<n_tiles, 256, 0, s>>>((int)n, h->d_pre_slot, h->d_pre_table, h->d_pre_tiles);
  scatter_kernel<<<n_tiles, 256, 0, s>>>((int)n, n_tiles, vals, h->d_pre_slot, h->d_pre_table, h->d_pre_tiles,
                                        h->d_points, n_out);
  h->launches += 3;
  unsigned host_n = 0;
  WS_CUDA_OK(cudaMemcpyAsync(&host_n, n_out, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  WS_CUDA_OK(cudaStreamSynchronize(s));
  WS_CUDA_OK(cudaGetLastError());
  return (int64_t)host_n;
}
