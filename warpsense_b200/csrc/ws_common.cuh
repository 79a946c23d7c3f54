// ws_common.cuh -- shared host/device definitions of the B200 TSDF path.
//
// Data layout in HBM (DESIGN.md "Layout"):
//   * the local map keeps the reference's ring-buffer semantics (include/map/hdf5_local_map.h:140-151,
//     /root/reference) but is stored BRICKED: 8x8x8 voxels (2 KB of 4-byte TSDF entries) are contiguous,
//     so a ray or a 7-point stencil touches few DRAM pages/sectors and the per-scan merge streams
//     whole bricks;
//   * a parallel scratch array holds one 64-bit candidate key per voxel (4 KB per brick);
//   * ring-x brick columns may be only partially resident (x-slab sharding across GPUs).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define WS_MR 32768            // MATRIX_RESOLUTION  (include/warpsense/consts.h:12-13)
#define WS_MR_SHIFT 15
#define WS_WR 64               // WEIGHT_RESOLUTION  (include/warpsense/consts.h:9-10)
#define WS_BRICK 8
#define WS_BRICK_VOX 512
#define WS_MAX_PEERS 16        // ranks one map can be sharded over (one NVSwitch domain: 8)
#define WS_MAX_XBRICKS 320     // ring-x brick columns a grid may have (2049/8 = 257)

#define WS_HD __host__ __device__ __forceinline__
#define WS_D __device__ __forceinline__

typedef unsigned long long u64;
typedef long long i64;

struct ws_pt { int x, y, z; };

// ---------------------------------------------------------------------------------------------
// exact integer division by a launch-uniform divisor (Lemire: q = mulhi64(ceil(2^64/d), n), valid
// for every 32-bit n and d; d == 1 handled apart).  C++ `/` on int truncates toward zero.
struct FastDiv
{
  u64 M;
  unsigned d;
};

static inline FastDiv make_fastdiv(unsigned d)
{
  FastDiv f;
  f.d = d;
  f.M = (d <= 1) ? 0ull : (~0ull) / d + 1ull;
  return f;
}

WS_HD unsigned fd_udiv(unsigned n, const FastDiv f)
{
#ifdef __CUDA_ARCH__
  return f.d == 1 ? n : (unsigned)__umul64hi(f.M, (u64)n);
#else
  return f.d == 1 ? n : (unsigned)(((unsigned __int128)f.M * (unsigned __int128)n) >> 64);
#endif
}

WS_D int fd_sdiv(int n, const FastDiv f)   // trunc toward zero, divisor > 0
{
  unsigned a = n < 0 ? 0u - (unsigned)n : (unsigned)n;
  unsigned q = fd_udiv(a, f);
  return n < 0 ? -(int)q : (int)q;
}

// trunc-toward-zero division of a signed 64-bit value by 2^15 (x / MATRIX_RESOLUTION on a long)
WS_HD i64 div_mr64(i64 x) { return (x + ((x >> 63) & (WS_MR - 1))) >> WS_MR_SHIFT; }
WS_HD int div_mr32(int x) { return (x + ((x >> 31) & (WS_MR - 1))) >> WS_MR_SHIFT; }

WS_HD int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
WS_HD int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
WS_HD int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }

// include/util/util.h:8-11
WS_HD void to_int_mat(const float T[16], int M[16])
{
  for (int i = 0; i < 16; i++) M[i] = (int)(T[i] * (float)WS_MR);
}

// include/util/util.h:13-18 (column-major M)
WS_HD void transform_point(const int M[16], int x, int y, int z, int out[3])
{
#pragma unroll
  for (int r = 0; r < 3; r++)
  {
    int acc = wadd(wadd(wadd(wmul(M[r], x), wmul(M[4 + r], y)), wmul(M[8 + r], z)), M[12 + r]);
    out[r] = div_mr32(acc);
  }
}

// (int)std::sqrt((double)sq) for 0 <= sq < 2^31 == floor(sqrt(sq)) (Eigen Vector3i::norm()).
WS_D int isqrt31(int sq)
{
  int r = (int)sqrtf((float)sq);
  // the float estimate is within +-1 of the exact root
  if ((unsigned)r * (unsigned)r > (unsigned)sq) r--;
  else if ((unsigned)(r + 1) * (unsigned)(r + 1) <= (unsigned)sq) r++;
  return r;
}

// ---------------------------------------------------------------------------------------------
// Grid descriptor handed to every kernel by value.
struct GridDesc
{
  int size[3];      // ring side lengths (odd)
  int half[3];      // size / 2
  int pos[3];       // map centre, voxel units
  int offset[3];    // ring offset of the centre (hdf5_local_map.h:59-70)
  int nb[3];        // bricks per axis = ceil(size / 8)
  int full;         // 1: every ring-x brick column is resident and slot == column
  int wide;         // 1: voxel addresses need more than 32 bits (n_bricks * 512 > 2^32; WS_FORCE_WIDE=1 forces the 64-bit path on small maps, for tests)
  i64 n_bricks;     // resident bricks
  uint32_t *grid;   // n_bricks * 512 TSDF entries  {int16 value | int16 weight << 16}
  u64 *keys;        // n_bricks * 512 candidate keys (all-ones when idle)
  unsigned *brick_flag;  // per resident brick: touched by the surface phase of the current scan's march
  unsigned *brick_flag2; // per resident brick: touched by the current scan at all (surface or free-space phase)
  unsigned *vstate;      // 4 bits per voxel (VS_*), all zero between scans
  unsigned char *ffree;  // per brick 512 + 512 bytes: a real / an interpolated free-space candidate arrived (plain stores), zero between scans
  unsigned *brick_slot_base;   // per resident brick: first pending slot of the voxels it parked in the current scan
  short xslot[WS_MAX_XBRICKS];  // ring-x brick column -> resident slot, -1 if not resident
  unsigned char xown[WS_MAX_XBRICKS];   // 1: this rank owns the column (sums its points in the registration)
};

WS_HD bool grid_in_bounds(const GridDesc &g, int x, int y, int z)
{
  // include/map/hdf5_local_map.h:275-279 : |p - pos| <= size/2 per axis
  int dx = x - g.pos[0], dy = y - g.pos[1], dz = z - g.pos[2];
  dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy; dz = dz < 0 ? -dz : dz;
  return dx <= g.half[0] && dy <= g.half[1] && dz <= g.half[2];
}

// include/map/hdf5_local_map.h:4-19,140-151 : ring coordinate of an in-bounds voxel
WS_HD int ring_coord(int v, int pos, int off, int size)
{
  int r = v - pos + off + size;
  if (r >= 2 * size) r -= 2 * size;
  else if (r >= size) r -= size;
  return r;
}

// brick-local offset of ring coordinates
WS_HD int brick_local(int rx, int ry, int rz) { return ((rx & 7) << 6) | ((ry & 7) << 3) | (rz & 7); }

// resident brick id of ring coordinates, -1 if the x column is not resident on this rank
WS_HD i64 brick_of(const GridDesc &g, int rx, int ry, int rz)
{
  int bx = rx >> 3;
  int slot = g.full ? bx : (int)g.xslot[bx];
  if (slot < 0) return -1;
  return ((i64)slot * g.nb[1] + (ry >> 3)) * g.nb[2] + (rz >> 3);
}

// ---------------------------------------------------------------------------------------------
// Candidate key (one 64-bit atomicMin per candidate reproduces the reference's sequential
// collision rule, update_tsdf.cpp:508-512 -- derivation in DESIGN.md "Collision rule"):
//   [63:62] tag      00 candidate, 10 PENDING (parked: [53:45] rank in the brick's slots, [44:0] seq of the parked
//                    winner), 11..1 EMPTY
//   [61:47] |value|  (<= tau <= 32767)
//   [46]    1 = interpolated candidate (negative weight)
//   [45:1]  order    seq for real candidates, SEQ_MAX - seq for interpolated ones
//   [0]     1 = value negative
// seq = point[24] | march step[15] | fan step[6]  == the reference's processing order.
#define WS_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define WS_KEY_PENDING_TAG 0x8000000000000000ull
#define WS_SEQ_BITS 45
#define WS_SEQ_MAX ((1ull << WS_SEQ_BITS) - 1ull)
#define WS_SEQ_STEP_BITS 6
#define WS_SEQ_MARCH_BITS 15
#define WS_SEQ_POINT_BITS 24

WS_HD u64 make_seq(unsigned point, unsigned march, unsigned fan)
{
  return ((u64)point << (WS_SEQ_MARCH_BITS + WS_SEQ_STEP_BITS)) | ((u64)march << WS_SEQ_STEP_BITS) | (u64)fan;
}

WS_HD u64 make_key(int value, bool interpolated, u64 seq)
{
  unsigned av = value < 0 ? (unsigned)(-value) : (unsigned)value;
  u64 ord = interpolated ? (WS_SEQ_MAX - seq) : seq;
  return ((u64)av << 47) | ((u64)(interpolated ? 1 : 0) << 46) | (ord << 1) | (u64)(value < 0 ? 1 : 0);
}

// Per-voxel scan state, 4 bits per voxel, eight voxels per 32-bit word in address order:
//   VS_FREE_REAL / VS_FREE_INT  a free-space candidate (|value| == tau, real / interpolated) arrived
//   VS_CLOSED                   the voxel has a candidate with |value| < tau: free-space candidates cannot win it
//   VS_PARKED                   ... and its winner so far is an interpolated one (settled by the replay)
#define VS_FREE_REAL 1u
#define VS_FREE_INT 2u
#define VS_CLOSED 4u
#define VS_PARKED 8u
WS_HD size_t vstate_word(u64 addr) { return (size_t)(addr >> 3); }
WS_HD unsigned vstate_shift(u64 addr) { return (unsigned)(addr & 7ull) << 2; }

WS_HD bool key_is_candidate(u64 k) { return (k >> 62) == 0; }
WS_HD bool key_is_pending(u64 k) { return (k >> 62) == 2; }
WS_HD int key_abs_value(u64 k) { return (int)((k >> 47) & 0x7FFF); }
WS_HD bool key_interpolated(u64 k) { return ((k >> 46) & 1) != 0; }
WS_HD int key_value(u64 k) { int av = key_abs_value(k); return (k & 1) ? -av : av; }
WS_HD u64 key_seq(u64 k)
{
  u64 ord = (k >> 1) & WS_SEQ_MAX;
  return key_interpolated(k) ? (WS_SEQ_MAX - ord) : ord;
}

// update_tsdf.cpp:475-479 : weight as a function of the (signed, clamped) value
WS_HD int tsdf_weight(int value, int tau, int weight_epsilon)
{
  int weight = WS_WR;
  if (value < -weight_epsilon) weight = WS_WR * (tau + value) / (tau - weight_epsilon);
  return weight;
}

// the same with the per-scan magic for / (tau - weight_epsilon) (UpdateParams::div_weps): the numerator is
// WEIGHT_RESOLUTION * (tau + value) >= 0 for every clamped value, so the unsigned magic is the C++ quotient
WS_HD int tsdf_weight_fd(int value, int tau, int weight_epsilon, const FastDiv div_weps)
{
  int weight = WS_WR;
  if (value < -weight_epsilon) weight = (tau + value) >= 0 ? (int)fd_udiv((unsigned)(WS_WR * (tau + value)), div_weps)
                                                           : WS_WR * (tau + value) / (tau - weight_epsilon);
  return weight;
}

WS_HD uint32_t make_entry(int value, int weight)
{
  return (uint32_t)(uint16_t)(int16_t)value | ((uint32_t)(uint16_t)(int16_t)weight << 16);
}
WS_HD int entry_value(uint32_t e) { return (int)(int16_t)(e & 0xFFFFu); }
WS_HD int entry_weight(uint32_t e) { return (int)(int16_t)(e >> 16); }

// n / d (C++ truncation) for d > 0 from a single-precision estimate and one exact correction step: the running
// average of update_tsdf.cpp:549 divides by ew + weight, a different divisor per voxel, and the compiler's generic
// 32-bit division is ~35 instructions (a fifth of all instructions of the two merge kernels, ncu source view).
// The estimate is within 0.25 of the real quotient when the quotient is below 2^20 (I2F, RCP and FMUL each
// contribute <= 2^-23 relative error), so its truncation is off by at most one; anything larger takes `/`.
WS_HD int div_trunc_small(int n, int d)
{
  const unsigned an = n < 0 ? 0u - (unsigned)n : (unsigned)n;
  if ((an >> 20) >= (unsigned)d || d >= (1 << 24)) return n / d;
#ifdef __CUDA_ARCH__
  int q = __float2int_rz(__fdividef(__uint2float_rn(an), __int2float_rn(d)));
#else
  int q = (int)((float)an / (float)d);
#endif
  const int r = (int)(an - (unsigned)q * (unsigned)d);
  if (r < 0) q--;
  else if (r >= d) q++;
  return n < 0 ? -q : q;
}

// update_tsdf.cpp:542-560 : fold one scan's surviving candidate into the stored entry
WS_HD uint32_t merge_entry(uint32_t e, int value, int weight, int max_weight)
{
  int ev = entry_value(e), ew = entry_weight(e);
  if (weight > 0 && ew > 0)
  {
    int nv = div_trunc_small(ev * ew + value * weight, ew + weight);
    int nw = (ew + weight) < max_weight ? (ew + weight) : max_weight;
    return make_entry(nv, nw);
  }
  if (weight != 0 && ew <= 0) return make_entry(value, weight);
  return e;
}
