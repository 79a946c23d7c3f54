// march_math.cuh -- exact 32-bit arithmetic of the ray march's fast path (host + device, unit-tested on the
// host by tests/cpp/test_march_math.cu against the plain C formulas of update_tsdf.cpp:450-506).
//
// The reference marches `len = 1, 1+h, ...` and computes, per step and axis,
//     proj = pos + (d * len) / distance        (int32, truncating)            update_tsdf.cpp:452
//     idx  = proj / res                        (truncating)                   update_tsdf.cpp:453
// Done literally that is two 64-bit magic divisions per axis and step.  Here:
//   * (|d| * len) / distance advances by a DDA: a lane that handles steps i, i+32, i+64 ... keeps the
//     quotient and the remainder and adds the (warp-uniform) quotient/remainder of |d|*32*h / distance;
//   * x / res uses a 32-bit magic (one IMAD.HI) that is exact for |x| < 2^32 / res.
// Both need bounded operands; `ray_is_small` states the bounds, rays outside them take the general path
// (64-bit magics, wrapping int32 products, exactly like the oracle).
#pragma once
#include "ws_common.cuh"

struct FastDiv32
{
  unsigned M;   // floor(2^32 / d) + 1
  unsigned d;
};

static inline FastDiv32 make_fastdiv32(unsigned d)
{
  FastDiv32 f;
  f.d = d;
  f.M = d <= 1 ? 0u : (unsigned)((1ull << 32) / d) + 1u;
  return f;
}

WS_HD unsigned ws_umulhi(unsigned a, unsigned b)
{
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (unsigned)(((u64)a * (u64)b) >> 32);
#endif
}

// n / d for n * d < 2^32, d >= 2
WS_HD unsigned fd32_udiv(unsigned n, const FastDiv32 f) { return ws_umulhi(n, f.M); }

// truncating n / d for |n| * d < 2^32, d >= 2
WS_HD int fd32_sdiv(int n, const FastDiv32 f)
{
  const unsigned a = n < 0 ? 0u - (unsigned)n : (unsigned)n;
  const int q = (int)ws_umulhi(a, f.M);
  return n < 0 ? -q : q;
}

// the same with the SIGNED high product: floor(n * M / 2^32) is the floored quotient, one below the truncated one
// for every negative n (multiples of d included: M > 2^32 / d), so adding n's sign bit gives C's `/`.
// Three instructions (IMAD.HI, SHF, IADD3 -- the add folds into a following subtraction).  Needs M < 2^31: d >= 3.
WS_HD int fd32_sdiv_s(int n, const FastDiv32 f)
{
#ifdef __CUDA_ARCH__
  return __mulhi(n, (int)f.M) + (int)((unsigned)n >> 31);
#else
  return (int)(((i64)n * (i64)f.M) >> 32) + (int)((unsigned)n >> 31);
#endif
}

// floor(n / dist) and the remainder from a double-precision reciprocal: the estimate is within +-1
// for n < 2^46 (relative error of fl(n * fl(1/dist)) <= 2^-52), one correction step makes it exact.
WS_HD void divrem_rcp(unsigned n, unsigned dist, double rdist, unsigned &q, unsigned &rem)
{
  unsigned qe = (unsigned)((double)n * rdist);
  int r = (int)(n - qe * dist);            // exact in wrap-around arithmetic: |true remainder| < 2 * dist
  if (r < 0) { qe--; r += (int)dist; }
  else if (r >= (int)dist) { qe++; r -= (int)dist; }
  q = qe; rem = (unsigned)r;
}

// the same for a 64-bit numerator below 2^46 (quotient only)
WS_HD u64 div_rcp64(u64 n, unsigned dist, double rdist)
{
  u64 qe = (u64)((double)n * rdist);
  i64 r = (i64)(n - qe * (u64)dist);
  if (r < 0) qe--;
  else if (r >= (i64)dist) qe++;
  return qe;
}

// DDA state of one axis: |d| * len = q * distance + rem
struct DdaAxis
{
  unsigned q, rem;     // per lane
  unsigned dq, drem;   // per 32 steps (warp-uniform)
};

WS_HD void dda_init(DdaAxis &s, unsigned absd, unsigned len0, unsigned stride_len, unsigned dist, double rdist)
{
  divrem_rcp(absd * len0, dist, rdist, s.q, s.rem);
  divrem_rcp(absd * stride_len, dist, rdist, s.dq, s.drem);
}

WS_HD void dda_advance(DdaAxis &s, unsigned dist)
{
  s.q += s.dq;
  s.rem += s.drem;
  if (s.rem >= dist) { s.rem -= dist; s.q++; }
}

// Bounds under which the fast path is exact for a ray (all products below fit int32 without wrapping and
// every dividend of a 32-bit magic division stays below 2^32 / res):
//   max|d| * (distance + tau + 32*h + 1) < 2^31      DDA numerators incl. one stride beyond the last step
//   dz * (distance + tau)               < 2^30       delta_z < 2^15: delta_z*iv and step*res*iv fit int32
//   |p|, |pos| < coord_lim                            (host: coord_lim = 2^31 / res - tau - 70*res - 2^16)
WS_HD bool ray_is_small(const int d[3], const int p[3], int distance, int tau, int half_res, int dz, int coord_lim)
{
  unsigned m = 0, c = 0;
#pragma unroll
  for (int a = 0; a < 3; a++)
  {
    const unsigned ad = d[a] < 0 ? 0u - (unsigned)d[a] : (unsigned)d[a];
    const unsigned ap = p[a] < 0 ? 0u - (unsigned)p[a] : (unsigned)p[a];
    m = ad > m ? ad : m;
    c = ap > c ? ap : c;
  }
  const i64 L = (i64)distance + tau + 32ll * half_res + 1;
  return coord_lim > 0 && c < (unsigned)coord_lim && (i64)m * L < (1ll << 31) && (i64)dz * L < (1ll << 30);
}

// Multi-GPU culling: the march-step ranges [seg[2k], seg[2k+1]) of a ray that can put a candidate into one of
// the x intervals [xiv_lo[k], xiv_hi[k]) (mm, map frame; sorted, disjoint, already widened by the reach of the
// interpolation fan).  The ray's x coordinate pos_x + (dx * len) / distance is monotone in the step index
// (len = 1 + i * h), so every interval maps to one contiguous range; ranges are conservative by two steps on
// either side, emitted in ascending order, overlapping neighbours merged, packed from slot 0, the rest zero.
// n_xiv <= 0 means "everything resident": one range [0, n_steps).
template <int MAXSEG>
WS_HD void ray_step_ranges(int n_xiv, const int *xiv_lo, const int *xiv_hi, int pos_x, int half_res, int dx,
                           int distance, int n_steps, int seg[2 * MAXSEG])
{
#pragma unroll
  for (int k = 0; k < 2 * MAXSEG; k++) seg[k] = 0;
  if (n_xiv <= 0) { seg[1] = n_steps; return; }
  int n = 0;
  for (int kk = 0; kk < n_xiv && n < MAXSEG; kk++)
  {
    const int k = dx >= 0 ? kk : n_xiv - 1 - kk;        // walk the intervals in the ray's direction of travel
    // (one millimetre wider on either side: the march truncates (dx * len) / distance, so a step's x may sit up to
    // 1 mm nearer to the sensor than the real ray -- for a ray that runs almost along y that is many steps)
    const double x0 = (double)xiv_lo[k] - (double)pos_x - 1.0, x1 = (double)xiv_hi[k] - (double)pos_x + 1.0;
    int i0 = 0, i1 = n_steps;
    if (dx == 0)
    {
      if (!(x0 <= 0.0 && 0.0 < x1)) continue;
    }
    else
    {
      const double s = (double)distance / (double)dx;
      double l0 = x0 * s, l1 = x1 * s;                   // lengths at which the ray crosses the interval ends
      if (l0 > l1) { const double t = l0; l0 = l1; l1 = t; }
      const double h = (double)half_res;
      const double f0 = floor((l0 - 1.0) / h) - 2.0, f1 = ceil((l1 - 1.0) / h) + 3.0;
      if (f1 <= 0.0 || f0 >= (double)n_steps) continue;
      i0 = f0 < 0.0 ? 0 : (int)f0;
      i1 = f1 > (double)n_steps ? n_steps : (int)f1;
    }
    if (i0 >= i1) continue;
    if (n > 0 && i0 <= seg[2 * n - 1])                  // touches the previous range: extend it
    {
      if (i1 > seg[2 * n - 1]) seg[2 * n - 1] = i1;
      if (i0 < seg[2 * n - 2]) seg[2 * n - 2] = i0;
      continue;
    }
    seg[2 * n] = i0; seg[2 * n + 1] = i1;
    n++;
  }
}
