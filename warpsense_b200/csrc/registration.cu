// registration.cu -- Point-to-TSDF registration inner loop for sm_100a.
//
// Replaces RegistrationCuda::perform_registration (src/warpsense/cuda/registration.cu:15-368 under
// /root/reference: calc_jacobis_krnl + h_g_e_reduction_krnl + host finish) with ONE fused kernel per
// Gauss-Newton iteration whose sums equal the reference's CPU path (src/cpu/registration.cpp:52-118)
// bit for bit: fixed-point transform -> centre + 6 face-neighbour gather from the bricked grid ->
// gated central differences -> 6x int64 Jacobian -> warp-shuffle reduction of the 21 upper-triangle
// H terms, 6 g terms, error and count -> 29 int64 atomics per block.  The last block to finish also
// runs the damped FP64 6x6 solve and the Rodrigues update (registration.cpp:128-157,
// include/warpsense/registration/util.h:5-39), so the next iteration starts with no host round trip.
// All sums are integers, hence exact and independent of reduction order and of the GPU count.
#include <cooperative_groups.h>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <cstdlib>
#include "ws_internal.h"

namespace cg = cooperative_groups;

#define FULL 0xFFFFFFFFu
#define WS_NSUM 29

// ------------------------------------------------------------------------------------------------
// FP64 damped solve + pose update, compiled for host and device from one source so the two agree.
// Mirrors Eigen's PartialPivLU-based Matrix<double,6,6>::inverse().
WS_HD void inverse6(const double A[36], double inv[36])
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

// include/warpsense/registration/util.h:5-39
WS_HD void xi_to_transform(const double xi[6], const int center[3], float out[16])
{
  double theta = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
  double l[3] = { xi[0] / theta, xi[1] / theta, xi[2] / theta };
  float L[3][3] = { { 0.f, (float)(-l[2]), (float)l[1] },
                    { (float)l[2], 0.f, (float)(-l[0]) },
                    { (float)(-l[1]), (float)l[0], 0.f } };
  float s = (float)sin(theta);
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
  float oc[3] = { -(float)center[0], -(float)center[1], -(float)center[2] };
  float xf[3] = { (float)xi[3], (float)xi[4], (float)xi[5] };
  for (int i = 0; i < 16; i++) out[i] = 0.f;
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) out[j * 4 + i] = R[i][j];
    float shift = R[i][0] * oc[0];
    shift = shift + R[i][1] * oc[1];
    shift = shift + R[i][2] * oc[2];
    shift = shift + 0.f * 1.f;
    out[12 + i] = (shift + (float)center[i]) + xf[i];
  }
  out[15] = 1.f;
}

// registration.cpp:128-145.  H column-major 6x6.  T updated in place; returns err / cnt.
WS_HD float gn_solve(const i64 H[36], const i64 g[6], int err, int cnt, float alpha, float T[16], double xi_out[6])
{
  int center[3] = { (int)T[12], (int)T[13], (int)T[14] };
  double hf[36], gf[6], inv[36], xi[6];
  for (int i = 0; i < 36; i++) hf[i] = (double)H[i];
  for (int i = 0; i < 6; i++) gf[i] = (double)g[i];
  double damp = (double)(alpha * (float)cnt);
  for (int i = 0; i < 6; i++) hf[i * 6 + i] += damp * 1.0;
  inverse6(hf, inv);
  for (int r = 0; r < 6; r++)
  {
    double acc = 0.0;
    for (int k = 0; k < 6; k++) acc += (-inv[k * 6 + r]) * gf[k];
    xi[r] = acc;
  }
  float X[16], N[16];
  xi_to_transform(xi, center, X);
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
    {
      float acc = X[0 * 4 + r] * T[c * 4 + 0];
      acc = acc + X[1 * 4 + r] * T[c * 4 + 1];
      acc = acc + X[2 * 4 + r] * T[c * 4 + 2];
      acc = acc + X[3 * 4 + r] * T[c * 4 + 3];
      N[c * 4 + r] = acc;
    }
  for (int i = 0; i < 16; i++) T[i] = N[i];
  if (xi_out) for (int i = 0; i < 6; i++) xi_out[i] = xi[i];
  return (float)err / cnt;
}

void ws_host_solve(const i64 H[36], const i64 g[6], int err, int cnt, float alpha, float T[16], double xi_out[6])
{
  gn_solve(H, g, err, cnt, alpha, T, xi_out);
}

namespace {

// expand the 21 upper-triangle sums (row-major, i <= j) into a column-major symmetric 6x6
WS_HD void expand_h(const u64 sums[WS_NSUM], i64 H[36], i64 g[6])
{
  int k = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++)
    {
      i64 v = (i64)sums[k++];
      H[j * 6 + i] = v;
      H[i * 6 + j] = v;
    }
  for (int i = 0; i < 6; i++) g[i] = (i64)sums[21 + i];
}

// H += J J^T (upper triangle), g += J v, err += |v|, cnt += 1   (registration.cpp:104-107)
WS_D void accumulate_point(i64 sum[WS_NSUM], const i64 J[6], const int cv)
{
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = a; b < 6; b++) sum[k++] += J[a] * J[b];
#pragma unroll
  for (int a = 0; a < 6; a++) sum[21 + a] += J[a] * cv;
  sum[27] += cv < 0 ? -cv : cv;
  sum[28] += 1;
}

// warp-shuffle reduction of the 29 exact sums, one shared-memory atomic per warp, 29 global atomics per
// block; returns true in the block that arrives last (its thread 0 may then read the complete sums)
WS_D bool reduce_and_commit(const i64 sum[WS_NSUM], u64 *s_sum, bool *s_last, RegAccum *acc)
{
#pragma unroll
  for (int i = 0; i < WS_NSUM; i++)
  {
    i64 v = sum[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(&s_sum[i], (u64)v);
  }
  __syncthreads();
  if (threadIdx.x < WS_NSUM && s_sum[threadIdx.x] != 0ull) atomicAdd(&acc->sums[threadIdx.x], s_sum[threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const unsigned t = atomicAdd(&acc->ticket, 1u);
    *s_last = (t == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  return *s_last;
}

// one thread: finish a Gauss-Newton iteration from the complete sums (registration.cpp:128-157)
__device__ void finish_iteration(RegAccum *acc, u64 *trace, int trace_cap, float it_weight_gradient, float epsilon)
{
  u64 sums[WS_NSUM];
  for (int i = 0; i < WS_NSUM; i++) sums[i] = ((volatile u64 *)acc->sums)[i];
  const unsigned it = acc->iterations;
  if (trace && (int)it < trace_cap)
    for (int i = 0; i < WS_NSUM; i++) trace[(size_t)it * WS_NSUM + i] = sums[i];
  i64 H[36], g[6];
  expand_h(sums, H, g);
  const int err = (int)(i64)sums[27];
  const int cnt = (int)(i64)sums[28];
  float T[16];
  for (int i = 0; i < 16; i++) T[i] = acc->T[i];
  const float e = gn_solve(H, g, err, cnt, acc->alpha, T, nullptr);
  for (int i = 0; i < 16; i++) acc->T[i] = T[i];
  acc->alpha += it_weight_gradient;                                                    // :141
  if (fabs(e - acc->prev_err[2]) < epsilon && fabs(e - acc->prev_err[0]) < epsilon)    // :146-150
    acc->finished = 1u;
  acc->prev_err[0] = acc->prev_err[1];                                                 // :151-155
  acc->prev_err[1] = acc->prev_err[2];
  acc->prev_err[2] = acc->prev_err[3];
  acc->prev_err[3] = e;
  acc->iterations = it + 1;
  for (int i = 0; i < WS_NSUM; i++) acc->sums[i] = 0ull;
}

__global__ void __launch_bounds__(256, 2)
reg_accum_kernel(const GridDesc g, const ws_pt *__restrict__ pts, const int n, const FastDiv div_res,
                 RegAccum *__restrict__ acc, u64 *__restrict__ trace, const int trace_cap,
                 const int fused_solve, const float it_weight_gradient, const float epsilon)
{
  if (acc->finished) return;

  __shared__ u64 s_sum[WS_NSUM];
  __shared__ bool s_last;
  if (threadIdx.x < WS_NSUM) s_sum[threadIdx.x] = 0ull;
  __syncthreads();

  float T[16];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = acc->T[i];
  int M[16];
  to_int_mat(T, M);                                                                    // :54
  const int cx = (int)T[12], cy = (int)T[13], cz = (int)T[14];                         // :52

  i64 sum[WS_NSUM];
#pragma unroll
  for (int i = 0; i < WS_NSUM; i++) sum[i] = 0;

  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
  {
    const ws_pt p = pts[j];
    int q[3];
    transform_point(M, p.x, p.y, p.z, q);                                              // :63
    const int bx = fd_sdiv(q[0], div_res), by = fd_sdiv(q[1], div_res), bz = fd_sdiv(q[2], div_res);  // :65
    const int px = wsub(q[0], cx), py = wsub(q[1], cy), pz = wsub(q[2], cz);           // :66

    // centre and all six face neighbours must be inside the map (:68-81 throw otherwise)
    int dx = bx - g.pos[0], dy = by - g.pos[1], dz = bz - g.pos[2];
    dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy; dz = dz < 0 ? -dz : dz;
    if (dx > g.half[0] || dy > g.half[1] || dz > g.half[2]) continue;
    const int rx = ring_coord(bx, g.pos[0], g.offset[0], g.size[0]);
    if (!g.xown[rx >> 3]) continue;                      // another rank sums this point
    const int ry = ring_coord(by, g.pos[1], g.offset[1], g.size[1]);
    const int rz = ring_coord(bz, g.pos[2], g.offset[2], g.size[2]);
    const uint32_t cur = g.grid[brick_of(g, rx, ry, rz) * WS_BRICK_VOX + brick_local(rx, ry, rz)];
    if (entry_weight(cur) == 0) continue;                                              // :71-74
    if (dx > g.half[0] - 1 || dy > g.half[1] - 1 || dz > g.half[2] - 1) continue;

    const int rxn = rx + 1 == g.size[0] ? 0 : rx + 1, rxl = rx == 0 ? g.size[0] - 1 : rx - 1;
    const int ryn = ry + 1 == g.size[1] ? 0 : ry + 1, ryl = ry == 0 ? g.size[1] - 1 : ry - 1;
    const int rzn = rz + 1 == g.size[2] ? 0 : rz + 1, rzl = rz == 0 ? g.size[2] - 1 : rz - 1;
    const uint32_t xn = g.grid[brick_of(g, rxn, ry, rz) * WS_BRICK_VOX + brick_local(rxn, ry, rz)];
    const uint32_t xl = g.grid[brick_of(g, rxl, ry, rz) * WS_BRICK_VOX + brick_local(rxl, ry, rz)];
    const uint32_t yn = g.grid[brick_of(g, rx, ryn, rz) * WS_BRICK_VOX + brick_local(rx, ryn, rz)];
    const uint32_t yl = g.grid[brick_of(g, rx, ryl, rz) * WS_BRICK_VOX + brick_local(rx, ryl, rz)];
    const uint32_t zn = g.grid[brick_of(g, rx, ry, rzn) * WS_BRICK_VOX + brick_local(rx, ry, rzn)];
    const uint32_t zl = g.grid[brick_of(g, rx, ry, rzl) * WS_BRICK_VOX + brick_local(rx, ry, rzl)];

    int gr[3] = { 0, 0, 0 };                                                           // :83-96
    {
      int v1 = entry_value(xn), v0 = entry_value(xl);
      if (entry_weight(xn) != 0 && entry_weight(xl) != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[0] = (v1 - v0) / 2;
      v1 = entry_value(yn); v0 = entry_value(yl);
      if (entry_weight(yn) != 0 && entry_weight(yl) != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[1] = (v1 - v0) / 2;
      v1 = entry_value(zn); v0 = entry_value(zl);
      if (entry_weight(zn) != 0 && entry_weight(zl) != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[2] = (v1 - v0) / 2;
    }
    i64 J[6];                                                                          // :98 (cross in int32)
    J[0] = wsub(wmul(py, gr[2]), wmul(pz, gr[1]));
    J[1] = wsub(wmul(pz, gr[0]), wmul(px, gr[2]));
    J[2] = wsub(wmul(px, gr[1]), wmul(py, gr[0]));
    J[3] = gr[0]; J[4] = gr[1]; J[5] = gr[2];
    accumulate_point(sum, J, entry_value(cur));
  }

  const bool last = reduce_and_commit(sum, s_sum, &s_last, acc);
  if (last && threadIdx.x == 0)
  {
    acc->ticket = 0u;
    if (fused_solve) finish_iteration(acc, trace, trace_cap, it_weight_gradient, epsilon);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent Gauss-Newton loop: ALL iterations of register_cloud in one cooperative kernel, one grid
// synchronisation per iteration.
//   * every thread keeps its points in registers across iterations;
//   * per iteration each block reduces its 29 exact sums (register butterfly + shared memory) and stores
//     them as one row of `partials`; after grid.sync() EVERY block adds up all rows and runs the same
//     deterministic FP64 solve (warp 0), so each block owns the next transform without a second barrier;
//   * the solve is inverse6()/gn_solve() above with the row operations spread over six lanes and the six
//     columns of the inverse over six lanes: operation order per element unchanged, results bit-identical.
#ifndef REG_THREADS
#define REG_THREADS 256
#endif
#ifndef REG_PTS
#define REG_PTS 2
#endif
#define REG_NSLOT 32

struct RegLoopParams
{
  int n;
  FastDiv div_res;
  int max_iterations;
  float it_weight_gradient;
  float epsilon;
};

// Multi-GPU exchange of the 29 Gauss-Newton sums INSIDE the persistent loop (one all-reduce
// of 21 + 6 + 2 scalars per iteration).  Every rank owns a mailbox in its own HBM, mapped into every peer
// (NVLink P2P: cudaIpc* across processes, plain pointers inside one).  After the grid barrier, block 0 of
// rank r pushes r's totals into slot r of EVERY mailbox as 8-byte words {32 data bits, 32-bit stamp}: an
// aligned 8-byte store is a single NVLink transaction, so a reader that sees the stamp sees the data (the
// NCCL "LL" idea) -- no fence, no separate flag, one NVLink write latency per iteration.  All blocks of a
// rank then poll their LOCAL mailbox, add the slots up (int64: exact, identical on every rank) and run the
// same solve.  Stamps carry (epoch, iteration); slots alternate with the iteration parity (a rank can be at
// most one iteration ahead of a peer) and with the epoch parity (... or one register_cloud call ahead).
#define WS_MAIL_WORDS 64                      // 8-byte words per (parity, sender) slot; 58 are used
#define WS_MAIL_USED (2 * WS_NSUM)
struct PeerParams
{
  int world, rank;
  unsigned stamp_base;                        // epoch << 8; stamp of iteration it = stamp_base + it + 1
  unsigned long long timeout_ns;
  uint2 *mail[WS_MAX_PEERS];                  // mail[p]: rank p's mailbox [4][world][WS_MAIL_WORDS]
};

WS_D void st_mail(uint2 *p, uint2 v)
{
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
WS_D uint2 ld_mail(const uint2 *p)
{
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
WS_D unsigned long long reg_global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one point's contribution for int transform M / centre c (registration.cpp:63-107); all seven voxel
// loads are issued together
WS_D void accumulate_cloud_point(const GridDesc &g, const int M[16], const int cx, const int cy, const int cz,
                                 const FastDiv div_res, const ws_pt p, i64 sum[WS_NSUM])
{
  int q[3];
  transform_point(M, p.x, p.y, p.z, q);                                                  // :63
  const int bx = fd_sdiv(q[0], div_res), by = fd_sdiv(q[1], div_res), bz = fd_sdiv(q[2], div_res);  // :65
  const int px = wsub(q[0], cx), py = wsub(q[1], cy), pz = wsub(q[2], cz);               // :66
  int dx = bx - g.pos[0], dy = by - g.pos[1], dz = bz - g.pos[2];
  dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy; dz = dz < 0 ? -dz : dz;
  if (dx > g.half[0] || dy > g.half[1] || dz > g.half[2]) return;                        // :68 throws
  const int rx = ring_coord(bx, g.pos[0], g.offset[0], g.size[0]);
  if (!g.xown[rx >> 3]) return;                            // another rank sums this point
  const int ry = ring_coord(by, g.pos[1], g.offset[1], g.size[1]);
  const int rz = ring_coord(bz, g.pos[2], g.offset[2], g.size[2]);
  const bool inner = !(dx > g.half[0] - 1 || dy > g.half[1] - 1 || dz > g.half[2] - 1);  // :76-81 throw otherwise
  const int rxn = rx + 1 == g.size[0] ? 0 : rx + 1, rxl = rx == 0 ? g.size[0] - 1 : rx - 1;
  const int ryn = ry + 1 == g.size[1] ? 0 : ry + 1, ryl = ry == 0 ? g.size[1] - 1 : ry - 1;
  const int rzn = rz + 1 == g.size[2] ? 0 : rz + 1, rzl = rz == 0 ? g.size[2] - 1 : rz - 1;
  const uint32_t cur = g.grid[brick_of(g, rx, ry, rz) * WS_BRICK_VOX + brick_local(rx, ry, rz)];
  uint32_t xn = 0, xl = 0, yn = 0, yl = 0, zn = 0, zl = 0;
  if (inner)
  {
    xn = g.grid[brick_of(g, rxn, ry, rz) * WS_BRICK_VOX + brick_local(rxn, ry, rz)];
    xl = g.grid[brick_of(g, rxl, ry, rz) * WS_BRICK_VOX + brick_local(rxl, ry, rz)];
    yn = g.grid[brick_of(g, rx, ryn, rz) * WS_BRICK_VOX + brick_local(rx, ryn, rz)];
    yl = g.grid[brick_of(g, rx, ryl, rz) * WS_BRICK_VOX + brick_local(rx, ryl, rz)];
    zn = g.grid[brick_of(g, rx, ry, rzn) * WS_BRICK_VOX + brick_local(rx, ry, rzn)];
    zl = g.grid[brick_of(g, rx, ry, rzl) * WS_BRICK_VOX + brick_local(rx, ry, rzl)];
  }
  if (entry_weight(cur) == 0 || !inner) return;                                          // :71-74

  int gr[3] = { 0, 0, 0 };                                                               // :83-96
  {
    int v1 = entry_value(xn), v0 = entry_value(xl);
    if (entry_weight(xn) != 0 && entry_weight(xl) != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[0] = (v1 - v0) / 2;
    v1 = entry_value(yn); v0 = entry_value(yl);
    if (entry_weight(yn) != 0 && entry_weight(yl) != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[1] = (v1 - v0) / 2;
    v1 = entry_value(zn); v0 = entry_value(zl);
    if (entry_weight(zn) != 0 && entry_weight(zl) != 0 && !((v1 > 0 && v0 < 0) || (v1 < 0 && v0 > 0))) gr[2] = (v1 - v0) / 2;
  }
  i64 J[6];                                                                              // :98 (cross in int32)
  J[0] = wsub(wmul(py, gr[2]), wmul(pz, gr[1]));
  J[1] = wsub(wmul(pz, gr[0]), wmul(px, gr[2]));
  J[2] = wsub(wmul(px, gr[1]), wmul(py, gr[0]));
  J[3] = gr[0]; J[4] = gr[1]; J[5] = gr[2];
  accumulate_point(sum, J, entry_value(cur));
}

// warp butterfly over 32 slots: afterwards lane l holds the warp total of slot l
WS_D i64 warp_transpose_reduce(i64 v[REG_NSLOT], const int lane)
{
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1)
  {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; i++)
    {
      const i64 send = upper ? v[i] : v[i + half];
      const i64 keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, half);
    }
  }
  return v[0];
}

struct GnState
{
  float T[16];
  float alpha;
  float prev_err[4];
  unsigned finished;
  unsigned iterations;
};

// warp 0 of a block: damped solve + pose update from the complete sums (registration.cpp:128-157),
// arithmetic identical to gn_solve()/inverse6() (same operations, same order per element)
WS_D void warp_gn_solve(const u64 *s_total, GnState *st, double *s_lu, int *s_perm, double *s_inv, double *s_xi,
                        const float it_weight_gradient, const float epsilon, const int lane)
{
  const int r = lane < 6 ? lane : 5;
  const int err = (int)(i64)s_total[27];
  const int cnt = (int)(i64)s_total[28];
  const double damp = (double)(st->alpha * (float)cnt);
  // row r of hf (symmetric: H[j*6+i] = H[i*6+j] = upper(i<=j))
  double row[6];
#pragma unroll
  for (int c = 0; c < 6; c++)
  {
    const int i = r < c ? r : c, j = r < c ? c : r;
    const int k = i * 6 - (i * (i - 1)) / 2 + (j - i);          // index of (i<=j) in the row-major upper triangle
    row[c] = (double)(i64)s_total[k];
    if (c == r) row[c] += damp * 1.0;
  }
  int prow = r;
  // LU with partial pivoting, one row per lane.  Rows are published to shared memory once per step and every
  // lane reads the pivot column and the pivot row from there (independent loads) -- the same arithmetic in
  // the same order as before, without the ~40 dependent warp shuffles per step that made the solve 3.7 us
#pragma unroll
  for (int k = 0; k < 6; k++)
  {
    if (lane < 6)
    {
#pragma unroll
      for (int c = 0; c < 6; c++) s_lu[lane * 6 + c] = row[c];
      s_perm[lane] = prow;
    }
    __syncwarp();
    // partial pivoting: first row (lowest index >= k) with the largest |lu[r][k]|
    int piv = k;
    double best = fabs(s_lu[k * 6 + k]);
#pragma unroll
    for (int rr = 1; rr < 6; rr++)
    {
      if (rr > k)
      {
        const double v = fabs(s_lu[rr * 6 + k]);
        if (v > best) { best = v; piv = rr; }
      }
    }
    // swap rows k and piv; the pivot row is the pre-swap row `piv`
    const int src = lane == k ? piv : (lane == piv ? k : lane);
    double prk[6];
#pragma unroll
    for (int c = 0; c < 6; c++) prk[c] = s_lu[piv * 6 + c];
    if (lane < 6 && src != lane)
    {
#pragma unroll
      for (int c = 0; c < 6; c++) row[c] = s_lu[src * 6 + c];
      prow = s_perm[src];
    }
    __syncwarp();
    // eliminate below the pivot
    if (lane < 6 && lane > k)
    {
      row[k] /= prk[k];
#pragma unroll
      for (int c = k + 1; c < 6; c++) row[c] -= row[k] * prk[c];
    }
  }
  if (lane < 6)
  {
#pragma unroll
    for (int c = 0; c < 6; c++) s_lu[lane * 6 + c] = row[c];
    s_perm[lane] = prow;
  }
  __syncwarp();
  // column `lane` of the inverse
  if (lane < 6)
  {
    double y[6];
#pragma unroll
    for (int rr = 0; rr < 6; rr++) y[rr] = (s_perm[rr] == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int rr = 0; rr < 6; rr++)
#pragma unroll
      for (int c = 0; c < rr; c++) y[rr] -= s_lu[rr * 6 + c] * y[c];
#pragma unroll
    for (int rr = 5; rr >= 0; rr--)
    {
#pragma unroll
      for (int c = rr + 1; c < 6; c++) y[rr] -= s_lu[rr * 6 + c] * y[c];
      y[rr] /= s_lu[rr * 6 + rr];
    }
#pragma unroll
    for (int rr = 0; rr < 6; rr++) s_inv[lane * 6 + rr] = y[rr];
  }
  __syncwarp();
  if (lane < 6)
  {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) acc += (-s_inv[k * 6 + lane]) * (double)(i64)s_total[21 + k];
    s_xi[lane] = acc;
  }
  __syncwarp();
  if (lane == 0)
  {
    const int center[3] = { (int)st->T[12], (int)st->T[13], (int)st->T[14] };
    double xi[6];
#pragma unroll
    for (int i = 0; i < 6; i++) xi[i] = s_xi[i];
    float X[16], N[16];
    xi_to_transform(xi, center, X);
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int rr = 0; rr < 4; rr++)
      {
        float acc = X[0 * 4 + rr] * st->T[c * 4 + 0];
        acc = acc + X[1 * 4 + rr] * st->T[c * 4 + 1];
        acc = acc + X[2 * 4 + rr] * st->T[c * 4 + 2];
        acc = acc + X[3 * 4 + rr] * st->T[c * 4 + 3];
        N[c * 4 + rr] = acc;
      }
#pragma unroll
    for (int i = 0; i < 16; i++) st->T[i] = N[i];
    const float e = (float)err / cnt;
    st->alpha += it_weight_gradient;                                                      // :141
    if (fabs(e - st->prev_err[2]) < epsilon && fabs(e - st->prev_err[0]) < epsilon)       // :146-150
      st->finished = 1u;
    st->prev_err[0] = st->prev_err[1];                                                    // :151-155
    st->prev_err[1] = st->prev_err[2];
    st->prev_err[2] = st->prev_err[3];
    st->prev_err[3] = e;
    st->iterations += 1u;
  }
}

template <bool MULTI>
__global__ void __launch_bounds__(REG_THREADS)
reg_loop_kernel(const GridDesc g, const ws_pt *__restrict__ pts, const RegLoopParams rp,
                RegAccum *__restrict__ acc, u64 *__restrict__ partials, u64 *__restrict__ trace, const int trace_cap,
                const PeerParams pp)
{
  cg::grid_group grid = cg::this_grid();
  __shared__ u64 s_w[REG_THREADS / 32][REG_NSLOT];
  __shared__ unsigned s_half[WS_MAX_PEERS][WS_MAIL_WORDS];
  __shared__ unsigned s_timeout;
  __shared__ bool s_last;
  __shared__ u64 s_total[REG_NSLOT];
  __shared__ GnState s_st;
  __shared__ double s_lu[36], s_inv[36], s_xi[6];
  __shared__ int s_perm[6];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gthread = blockIdx.x * REG_THREADS + tid;
  const int gthreads = gridDim.x * REG_THREADS;

  if (tid < 16) s_st.T[tid] = acc->T[tid];
  if (tid < 4) s_st.prev_err[tid] = 0.f;
  if (tid == 0) { s_st.alpha = acc->alpha; s_st.finished = 0u; s_st.iterations = 0u; s_timeout = 0u; }

  ws_pt my[REG_PTS];
#pragma unroll
  for (int k = 0; k < REG_PTS; k++)
  {
    const int j = gthread + k * gthreads;
    my[k].x = 0; my[k].y = 0; my[k].z = 0;
    if (j < rp.n) my[k] = pts[j];
  }
  __syncthreads();

  for (int it = 0; it < rp.max_iterations; it++)
  {
#ifdef WS_REG_TIMING
    unsigned long long tstamp[4];
    tstamp[0] = reg_global_ns();
#endif
    float T[16];
#pragma unroll
    for (int i = 0; i < 16; i++) T[i] = s_st.T[i];
    int M[16];
    to_int_mat(T, M);                                                                    // :54
    const int cx = (int)T[12], cy = (int)T[13], cz = (int)T[14];                         // :52

    i64 sum[REG_NSLOT];
#pragma unroll
    for (int i = 0; i < REG_NSLOT; i++) sum[i] = 0;
#pragma unroll
    for (int k = 0; k < REG_PTS; k++)
      if (gthread + k * gthreads < rp.n) accumulate_cloud_point(g, M, cx, cy, cz, rp.div_res, my[k], sum);
    for (int j = gthread + REG_PTS * gthreads; j < rp.n; j += gthreads)   // clouds larger than the register cache
      accumulate_cloud_point(g, M, cx, cy, cz, rp.div_res, pts[j], sum);

    const i64 mine = warp_transpose_reduce(sum, lane);
    s_w[warp][lane] = (u64)mine;
    __syncthreads();
#ifdef WS_REG_TIMING
    tstamp[1] = reg_global_ns();
#endif
#ifdef WS_REG_GRIDSYNC
    u64 *row = partials + ((size_t)(it & 1) * gridDim.x + blockIdx.x) * REG_NSLOT;
    if (tid < REG_NSLOT)
    {
      u64 a = 0ull;
#pragma unroll
      for (int w = 0; w < REG_THREADS / 32; w++) a += s_w[w][tid];
      row[tid] = a;
    }
    __threadfence();
    grid.sync();

    // total over all blocks' rows: every block (one GPU), block 0 only (several GPUs: the others read it
    // back from the mailbox together with the peers' totals)
    if (!MULTI || blockIdx.x == 0)
    {
      const u64 *base = partials + (size_t)(it & 1) * gridDim.x * REG_NSLOT;
      u64 a = 0ull;
      for (int b = warp; b < (int)gridDim.x; b += REG_THREADS / 32) a += __ldcg(&base[(size_t)b * REG_NSLOT + lane]);
      s_w[warp][lane] = a;
      __syncthreads();
      if (tid < REG_NSLOT)
      {
        u64 a2 = 0ull;
#pragma unroll
        for (int w = 0; w < REG_THREADS / 32; w++) a2 += s_w[w][tid];
        s_total[tid] = a2;
        // (the trace store sits HERE, not after the barrier: there it changed the loop's convergence-barrier
        // nesting and every grid barrier waited ~10x longer -- 0.25 -> 0.73 ms per 20 iterations on B200)
        if (!MULTI && blockIdx.x == 0 && trace && it < trace_cap && tid < WS_NSUM) trace[(size_t)it * WS_NSUM + tid] = a2;
      }
      __syncthreads();
    }
    const bool sender = blockIdx.x == 0;
#else
    // Grid-wide sum without a grid barrier: every block adds its 29 sums into one of three rotating
    // accumulators (int64 atomics: exact, order-free) and takes a ticket; whoever holds the last ticket knows
    // the accumulator is complete.  One GPU: everybody polls the ticket counter and reads the totals.  Several
    // GPUs: the last block pushes them to the peers and everybody polls the mailbox instead.
    u64 *tot = partials + (size_t)(it % 3) * REG_NSLOT;
    unsigned *tickets = reinterpret_cast<unsigned *>(partials + 3 * REG_NSLOT);
    const unsigned target = (unsigned)(it + 1) * gridDim.x;
    if (blockIdx.x == 0 && tid < REG_NSLOT) partials[(size_t)((it + 1) % 3) * REG_NSLOT + tid] = 0ull;   // free since it-2
    if (tid < WS_NSUM)
    {
      u64 a = 0ull;
#pragma unroll
      for (int w = 0; w < REG_THREADS / 32; w++) a += s_w[w][tid];
      if (a) atomicAdd(reinterpret_cast<unsigned long long *>(&tot[tid]), (unsigned long long)a);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
      const unsigned ticket = atomicAdd(tickets, 1u);
      s_last = ticket == target - 1u;
      if (!MULTI)
      {
        while (*((volatile unsigned *)tickets) < target) { }
      }
      __threadfence();
    }
    __syncthreads();
    const bool sender = s_last;
    if (!MULTI || sender)
    {
      if (tid < REG_NSLOT)
      {
        const u64 a2 = tid < WS_NSUM ? __ldcg(&tot[tid]) : 0ull;
        s_total[tid] = a2;
        if (!MULTI && blockIdx.x == 0 && trace && it < trace_cap && tid < WS_NSUM) trace[(size_t)it * WS_NSUM + tid] = a2;
      }
      __syncthreads();
    }
#endif
    if (MULTI)
    {
      const unsigned stamp = pp.stamp_base + (unsigned)it + 1u;
      const size_t parity_off = (size_t)(((pp.stamp_base >> 8) & 1u) * 2u + (unsigned)(it & 1)) * pp.world * WS_MAIL_WORDS;
      if (sender)
      {
        // push this rank's totals into slot `rank` of every mailbox (own one included)
        for (int idx = tid; idx < pp.world * WS_MAIL_USED; idx += REG_THREADS)
        {
          const int p = idx / WS_MAIL_USED, w = idx - p * WS_MAIL_USED;
          const u64 v = s_total[w >> 1];
          uint2 word;
          word.x = (w & 1) ? (unsigned)(v >> 32) : (unsigned)v;
          word.y = stamp;
          st_mail(pp.mail[p] + parity_off + (size_t)pp.rank * WS_MAIL_WORDS + w, word);
        }
      }
      // every block: wait for all senders' words of this iteration in the LOCAL mailbox
      const uint2 *mine = pp.mail[pp.rank] + parity_off;
      const unsigned long long t0 = reg_global_ns();
      for (int idx = tid; idx < pp.world * WS_MAIL_USED; idx += REG_THREADS)
      {
        const int p = idx / WS_MAIL_USED, w = idx - p * WS_MAIL_USED;
        uint2 word = ld_mail(mine + (size_t)p * WS_MAIL_WORDS + w);
        unsigned spins = 0;
        while (word.y != stamp)
        {
          if ((++spins & 1023u) == 0u && reg_global_ns() - t0 > pp.timeout_ns) { s_timeout = 1u; break; }
          word = ld_mail(mine + (size_t)p * WS_MAIL_WORDS + w);
        }
        s_half[p][w] = word.x;
      }
      __syncthreads();
      if (tid < REG_NSLOT)
      {
        u64 a = 0ull;
        if (tid < WS_NSUM)
          for (int p = 0; p < pp.world; p++) a += (u64)s_half[p][2 * tid] | ((u64)s_half[p][2 * tid + 1] << 32);
        s_total[tid] = a;
        if (blockIdx.x == 0 && trace && it < trace_cap && tid < WS_NSUM) trace[(size_t)it * WS_NSUM + tid] = a;
      }
      __syncthreads();
      if (s_timeout)          // a peer never delivered: every block of every rank gives up in the same iteration
      {
        if (tid == 0) s_st.finished = 2u;
        __syncthreads();
        break;
      }
    }
#ifdef WS_REG_TIMING
    tstamp[2] = reg_global_ns();
#endif
    if (warp == 0) warp_gn_solve(s_total, &s_st, s_lu, s_perm, s_inv, s_xi, rp.it_weight_gradient, rp.epsilon, lane);
    __syncthreads();
#ifdef WS_REG_TIMING
    tstamp[3] = reg_global_ns();
    if (blockIdx.x == 0 && tid == 0 && trace && it < trace_cap)
      for (int q = 0; q < 4; q++) trace[(size_t)it * WS_NSUM + q] = tstamp[q];
#endif
    if (s_st.finished) break;
  }

  if (blockIdx.x == 0)
  {
    if (tid < 16) acc->T[tid] = s_st.T[tid];
    if (tid < 4) acc->prev_err[tid] = s_st.prev_err[tid];
    if (tid == 0)
    {
      acc->alpha = s_st.alpha;
      acc->finished = s_st.finished;
      acc->iterations = s_st.iterations;
    }
  }
}

// test/cuda.cpp:416-532 shape: reduce caller-supplied Jacobians/values through the same code path
__global__ void __launch_bounds__(256)
test_reduce_kernel(const i64 *__restrict__ jac, const int *__restrict__ values, const int n, RegAccum *__restrict__ acc)
{
  __shared__ u64 s_sum[WS_NSUM];
  __shared__ bool s_last;
  if (threadIdx.x < WS_NSUM) s_sum[threadIdx.x] = 0ull;
  __syncthreads();
  i64 sum[WS_NSUM];
#pragma unroll
  for (int i = 0; i < WS_NSUM; i++) sum[i] = 0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
  {
    i64 J[6];
    for (int a = 0; a < 6; a++) J[a] = jac[(size_t)j * 6 + a];
    accumulate_point(sum, J, values[j]);
  }
  const bool last = reduce_and_commit(sum, s_sum, &s_last, acc);
  if (last && threadIdx.x == 0) acc->ticket = 0u;
}

__global__ void reg_solve_kernel(RegAccum *acc, u64 *trace, int trace_cap, float it_weight_gradient, float epsilon)
{
  if (acc->finished) return;
  finish_iteration(acc, trace, trace_cap, it_weight_gradient, epsilon);
}

// (the start transform comes by value: nothing staged in host memory that a second scan in flight could overwrite;
// the rotating accumulators + ticket counter of reg_loop_kernel are cleared here too)
struct RegMat16 { float m[16]; };
__global__ void reg_reset_kernel(RegAccum *acc, const RegMat16 T0, float alpha0, u64 *partials, int n_partials)
{
  const int t = threadIdx.x;
  for (int i = t; i < n_partials; i += blockDim.x) partials[i] = 0ull;
  if (t < 32) acc->sums[t] = 0ull;
  if (t < 16) acc->T[t] = T0.m[t];
  if (t < 4) acc->prev_err[t] = 0.f;
  if (t == 0) { acc->ticket = 0u; acc->finished = 0u; acc->iterations = 0u; acc->alpha = alpha0; }
}

// registration.cpp:164-174 : apply the final transform to the cloud in place
__global__ void __launch_bounds__(256)
transform_cloud_kernel(ws_pt *pts, int n, const RegAccum *acc)
{
  float T[16];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = acc->T[i];
  int M[16];
  to_int_mat(T, M);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
  {
    const ws_pt p = pts[j];
    int q[3];
    transform_point(M, p.x, p.y, p.z, q);
    ws_pt o; o.x = q[0]; o.y = q[1]; o.z = q[2];
    pts[j] = o;
  }
}

}  // namespace

static void ensure_reg_loop_buffers(ws_handle *h)
{
  if (h->reg_loop_blocks != 0) return;
  int per_sm = 0;
  WS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reg_loop_kernel<true>, REG_THREADS, 0));
  if (per_sm < 1) throw std::runtime_error("reg_loop_kernel does not fit on an SM");
  if (per_sm > 2) per_sm = 2;
  h->reg_loop_blocks = per_sm * h->sm_count;
  // tests that run several ranks' persistent kernels side by side on ONE GPU need them co-resident
  if (const char *e = std::getenv("WS_REG_BLOCKS"))
  {
    const int v = std::atoi(e);
    if (v >= 1 && v < h->reg_loop_blocks) h->reg_loop_blocks = v;
  }
  WS_CUDA_OK(cudaMalloc(&h->d_reg_partials, ((size_t)2 * h->reg_loop_blocks + 4) * REG_NSLOT * sizeof(u64)));
}

void ws_launch_reg_reset(ws_handle *h, const float T[16], float alpha0)
{
  ensure_reg_loop_buffers(h);
  RegMat16 T0;
  std::memcpy(T0.m, T, sizeof(T0.m));
  reg_reset_kernel<<<1, 128, 0, h->stream>>>(h->d_acc, T0, alpha0, h->d_reg_partials, 4 * REG_NSLOT);
  h->launches++;
}

void ws_launch_reg_iteration(ws_handle *h, int n, int res, int fused_solve, float it_weight_gradient, float epsilon)
{
  int blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  ws_timer_begin(h, WS_TIMER_REG);
  reg_accum_kernel<<<blocks, 256, 0, h->stream>>>(h->g, h->d_reg_points_alias ? h->d_reg_points_alias : h->d_reg_points, n, make_fastdiv((unsigned)res), h->d_acc,
                                                  h->d_trace, h->trace_cap, fused_solve, it_weight_gradient, epsilon);
  ws_timer_end(h);
  h->launches++;
}

size_t ws_reg_mailbox_bytes(int world) { return (size_t)4 * world * WS_MAIL_WORDS * sizeof(uint2); }

void ws_launch_reg_loop(ws_handle *h, int n, int res, int max_iterations, float it_weight_gradient, float epsilon)
{
  ensure_reg_loop_buffers(h);
  RegLoopParams rp;
  rp.n = n;
  rp.div_res = make_fastdiv((unsigned)res);
  rp.max_iterations = max_iterations;
  rp.it_weight_gradient = it_weight_gradient;
  rp.epsilon = epsilon;
  PeerParams pp{};
  pp.world = 1; pp.rank = 0;
  if (h->world > 1)
  {
    if (!h->peers_attached) throw std::logic_error("register_cloud on a sharded map needs ws_peer_attach* first");
    if (max_iterations > 255) throw std::invalid_argument("register_cloud on a sharded map: max_iterations <= 255");
    pp.world = h->world; pp.rank = h->rank;
    h->reg_epoch = (h->reg_epoch + 1u) & 0xFFFFFFu;
    if (h->reg_epoch == 0u) h->reg_epoch = 1u;
    pp.stamp_base = h->reg_epoch << 8;
    pp.timeout_ns = h->peer_timeout_ns;
    for (int p = 0; p < h->world; p++) pp.mail[p] = static_cast<uint2 *>(h->peer_mail[p]);
  }
  ws_pt *reg_pts = h->d_reg_points_alias ? h->d_reg_points_alias : h->d_reg_points;
  void *args[] = { (void *)&h->g, (void *)&reg_pts, (void *)&rp, (void *)&h->d_acc,
                   (void *)&h->d_reg_partials, (void *)&h->d_trace, (void *)&h->trace_cap, (void *)&pp };
  // (the accumulators + tickets were cleared by reg_reset_kernel, which every caller runs first)
  ws_timer_begin(h, WS_TIMER_REG);
  const void *fn = h->world > 1 ? (const void *)reg_loop_kernel<true> : (const void *)reg_loop_kernel<false>;
  WS_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(h->reg_loop_blocks), dim3(REG_THREADS), args, 0, h->stream));
  ws_timer_end(h);
  h->launches++;
}

void ws_launch_reg_solve(ws_handle *h, float it_weight_gradient, float epsilon)
{
  reg_solve_kernel<<<1, 1, 0, h->stream>>>(h->d_acc, h->d_trace, h->trace_cap, it_weight_gradient, epsilon);
  h->launches++;
}

void ws_launch_test_reduce(ws_handle *h, const i64 *d_jacobis, const int *d_values, int n)
{
  int blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  test_reduce_kernel<<<blocks, 256, 0, h->stream>>>(d_jacobis, d_values, n, h->d_acc);
  h->launches++;
}

void ws_launch_transform_cloud(ws_handle *h, ws_pt *d_pts, int n)
{
  if (n <= 0) return;
  transform_cloud_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(d_pts, n, h->d_acc);
  h->launches++;
}
