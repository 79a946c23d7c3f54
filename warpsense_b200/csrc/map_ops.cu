// map_ops.cu -- grid residency: fill, host ring layout <-> bricked device layout, voxel access,
// box pack/unpack for the local-map shift.
//
// Replaces cuda::DeviceMapMemWrapper (src/warpsense/cuda/device_map_wrapper.cu:6-97 under
// /root/reference), which copies the whole host ring array with one blocking cudaMemcpy.  Here the
// host array keeps the reference layout (x slowest, z fastest, ring coordinates) and is re-tiled into
// 8x8x8 bricks on the way in and out, one ring-x brick column (8 x-rows) per staged copy.
#include <algorithm>
#include <stdexcept>
#include "ws_internal.h"

namespace {

__global__ void fill_kernel(uint32_t *grid, u64 *keys, unsigned *flags, unsigned *flags2, i64 n_vox, i64 n_bricks, uint32_t entry)
{
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_vox; i += stride)
  {
    grid[i] = entry;
    keys[i] = WS_KEY_EMPTY;
  }
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_bricks; i += stride) { flags[i] = 0u; flags2[i] = 0u; }
}

// rows [x0, x0+nx) of the host ring array (already on the device in `lin`, row-major [nx][sy][sz])
// <-> bricks.  TO_BRICKS selects the direction.
template <bool TO_BRICKS>
__global__ void retile_kernel(const GridDesc g, uint32_t *lin, int x0, int nx)
{
  const i64 total = (i64)nx * g.size[1] * g.size[2];
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
  {
    const int rz = (int)(i % g.size[2]);
    const i64 t = i / g.size[2];
    const int ry = (int)(t % g.size[1]);
    const int rx = x0 + (int)(t / g.size[1]);
    const i64 b = brick_of(g, rx, ry, rz);
    if (b < 0) continue;
    const i64 addr = b * WS_BRICK_VOX + brick_local(rx, ry, rz);
    if (TO_BRICKS) g.grid[addr] = lin[i];
    else lin[i] = g.grid[addr];
  }
}

// world-voxel box [lo, lo+ext) <-> dense buffer, x slowest / z fastest inside the box.
// Voxels of non-resident columns are left untouched in `buf` (pack) / skipped (unpack).
template <bool PACK>
__global__ void box_kernel(const GridDesc g, uint32_t *buf, int lx, int ly, int lz, int ex, int ey, int ez)
{
  const i64 total = (i64)ex * ey * ez;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
  {
    const int z = lz + (int)(i % ez);
    const i64 t = i / ez;
    const int y = ly + (int)(t % ey);
    const int x = lx + (int)(t / ey);
    const int rx = ring_coord(x, g.pos[0], g.offset[0], g.size[0]);
    const int ry = ring_coord(y, g.pos[1], g.offset[1], g.size[1]);
    const int rz = ring_coord(z, g.pos[2], g.offset[2], g.size[2]);
    const i64 b = brick_of(g, rx, ry, rz);
    if (b < 0) continue;
    const i64 addr = b * WS_BRICK_VOX + brick_local(rx, ry, rz);
    if (PACK) buf[i] = g.grid[addr];
    else g.grid[addr] = buf[i];
  }
}

// order-free checksum of the entries of ring-x rows [x_lo, x_hi): sum over the voxels of
// mix(ring index) * (2 * entry + 1) mod 2^64.  Identical for any sharding of the same map contents.
__device__ __forceinline__ u64 mix64(u64 x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(256)
checksum_kernel(const GridDesc g, int x_lo, int x_hi, int owned_only, u64 *out)
{
  u64 acc = 0ull;
  const i64 per_col = (i64)g.nb[1] * g.nb[2] * WS_BRICK_VOX;
  const i64 total = (i64)g.nb[0] * per_col;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
  {
    const int col = (int)(i / per_col);
    const int slot = g.full ? col : (int)g.xslot[col];
    if (slot < 0 || (owned_only && !g.full && !g.xown[col])) continue;
    const i64 r = i - (i64)col * per_col;
    const int local = (int)(r & (WS_BRICK_VOX - 1));
    const i64 bb = r >> 9;
    const int by = (int)(bb / g.nb[2]), bz = (int)(bb % g.nb[2]);
    const int rx = col * 8 + (local >> 6), ry = by * 8 + ((local >> 3) & 7), rz = bz * 8 + (local & 7);
    if (rx < x_lo || rx >= x_hi || rx >= g.size[0] || ry >= g.size[1] || rz >= g.size[2]) continue;
    const u64 lin = ((u64)rx * (u64)g.size[1] + (u64)ry) * (u64)g.size[2] + (u64)rz;
    const u64 e = g.grid[((i64)slot * g.nb[1] * g.nb[2] + bb) * WS_BRICK_VOX + local];
    acc += mix64(lin + 0x9e3779b97f4a7c15ull) * (2ull * e + 1ull);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

}  // namespace

void ws_launch_checksum(ws_handle *h, int x_lo, int x_hi, int owned_only, u64 *d_out)
{
  WS_CUDA_OK(cudaMemsetAsync(d_out, 0, sizeof(u64), h->stream));
  checksum_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(h->g, x_lo, x_hi, owned_only, d_out);
  WS_CUDA_OK(cudaGetLastError());
}

void ws_launch_fill(ws_handle *h, uint32_t entry)
{
  const i64 n_vox = h->g.n_bricks * WS_BRICK_VOX;
  fill_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(h->g.grid, h->g.keys, h->g.brick_flag, h->g.brick_flag2, n_vox, h->g.n_bricks, entry);
  WS_CUDA_OK(cudaGetLastError());
}

static void retile(ws_handle *h, uint32_t *host, bool to_device)
{
  const GridDesc &g = h->g;
  const size_t row = (size_t)g.size[1] * g.size[2];
  uint32_t *stage = nullptr;
  WS_CUDA_OK(cudaMalloc(&stage, row * WS_BRICK * sizeof(uint32_t)));
  try
  {
    for (int bx = 0; bx < g.nb[0]; bx++)
    {
      if (!g.full && g.xslot[bx] < 0) continue;
      const int x0 = bx * WS_BRICK;
      const int nx = std::min(WS_BRICK, g.size[0] - x0);
      const size_t bytes = row * nx * sizeof(uint32_t);
      if (to_device)
      {
        WS_CUDA_OK(cudaMemcpyAsync(stage, host + row * x0, bytes, cudaMemcpyHostToDevice, h->stream));
        retile_kernel<true><<<h->sm_count * 4, 256, 0, h->stream>>>(g, stage, x0, nx);
      }
      else
      {
        retile_kernel<false><<<h->sm_count * 4, 256, 0, h->stream>>>(g, stage, x0, nx);
        WS_CUDA_OK(cudaMemcpyAsync(host + row * x0, stage, bytes, cudaMemcpyDeviceToHost, h->stream));
      }
      // pageable host memory: the copy above is staged synchronously, the buffer is reused next turn
      WS_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
  }
  catch (...)
  {
    cudaFree(stage);
    throw;
  }
  WS_CUDA_OK(cudaFree(stage));
}

void ws_launch_upload(ws_handle *h, const uint32_t *host) { retile(h, const_cast<uint32_t *>(host), true); }
void ws_launch_download(ws_handle *h, uint32_t *host) { retile(h, host, false); }

void ws_box_transfer(ws_handle *h, uint32_t *d_buf, const int lo[3], const int ext[3], bool pack)
{
  const i64 total = (i64)ext[0] * ext[1] * ext[2];
  if (total <= 0) return;
  int blocks = (int)std::min<i64>((total + 255) / 256, (i64)h->sm_count * 16);
  if (pack) box_kernel<true><<<blocks, 256, 0, h->stream>>>(h->g, d_buf, lo[0], lo[1], lo[2], ext[0], ext[1], ext[2]);
  else box_kernel<false><<<blocks, 256, 0, h->stream>>>(h->g, d_buf, lo[0], lo[1], lo[2], ext[0], ext[1], ext[2]);
  WS_CUDA_OK(cudaGetLastError());
}
