// update_tsdf.cu -- TSDF volume update for sm_100a.
//
// Replaces the reference's TSDFCuda::update_tsdf (src/warpsense/cuda/update_tsdf.cu:13-166 under
// /root/reference) with kernels whose RESULT equals the reference's CPU path
// (src/cpu/update_tsdf.cpp:397-564), which is the parity target.
//
// Pipeline per scan (one stream, one host synchronisation at the end to fetch the work counters):
//   0. setup_kernel         one THREAD per ray: direction, distance, interpolation vector, DDA increments,
//                           fast-path flag and -- on a sharded map -- the march-step ranges that can reach this
//                           rank's columns; 128-byte RaySetup per ray, in scan order.
//   1. march_kernel<true>   persistent warps, one ray per warp at a time (indices from a global counter, the
//                           next ray's RaySetup staged to shared memory by cp.async while this one is
//                           marched).  Lanes stride over the res/2 march steps (exact DDA, 32-bit magics:
//                           march_math.cuh); the steps that survive the reference's "same (x,y) column as
//                           the previous step" filter (:455-458) are compacted through a per-warp
//                           shared-memory queue so the expensive part (value, fan of interpolated voxels,
//                           addressing) runs with full lanes.  Every candidate (voxel, value,
//                           real|interpolated, order) is ONE 64-bit atomicMin (RED, no return) on the
//                           voxel's key -- see ws_common.cuh / DESIGN.md for why min over (|value|,
//                           interpolated, order) reproduces the sequential rule at :508-512 -- plus a plain
//                           store that flags the voxel's 8x8x8 brick.  Far-field candidates (the only ones
//                           that can meet an interpolated winner) are also appended to a record list in
//                           64-entry chunks owned by the warp.
//   2. brick_list_kernel    touched-brick flags -> compact list (and flags reset for the next scan).
//   3. merge_kernel         moves the touched bricks (4 KB keys + 2 KB entries each) through shared memory
//                           with TMA bulk copies (cp.async.bulk + mbarrier, three stages per CTA), folds final
//                           winners into the grid (:542-560), resets the keys, and parks the voxels whose
//                           winner is an interpolated candidate below tau ("pending"): their key word then
//                           carries the slot rank within the brick and the order of the parked winner.
//   4. replay_kernel        cooperative (grid-synchronised).  Round 1 streams the record list once and
//                           keeps, per pending voxel, the minimum key among the candidates that follow the
//                           parked winner in the reference's order; those candidates are also compacted
//                           into a short list that later rounds stream instead.  Rounds repeat on the
//                           device until every pending voxel has its final winner.
// If the record list overflows its buffer the host grows it and regenerates it with march_kernel<false>
// (far part of every ray, no atomics) before running replay_kernel again.
// The scanner pose arrives as kernel parameters or, in the fused per-scan pipeline (ws_track_scan), from
// device memory written by pose_kernel right after the registration.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>
#include "ws_internal.h"
#include "march_math.cuh"

namespace cg = cooperative_groups;

#define FULL 0xFFFFFFFFu
#define PEND_DONE 0xFFFFFFFFFFFFFFFFull
#ifndef MARCH_WARPS
#define MARCH_WARPS 8
#endif
#ifndef MARCH_CTAS
#define MARCH_CTAS 3          // CTAs per SM (register budget 65536 / (CTAS * THREADS))
#endif
#define MARCH_THREADS (MARCH_WARPS * 32)
#define QCAP 64
#ifndef RAY_BATCH
#define RAY_BATCH 1          // same-box A/B on B200: 1 -> 0.77 ms, 2 -> 0.79, 4 -> 0.91, 8 -> 1.07 (tail of long rays)
#endif

namespace {

WS_D unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// include/map/hdf5_local_map.h:275-279, |v - pos| <= size/2 per axis, as 0 <= v - lo <= size - 1
WS_D bool box_in_bounds(const UpdateParams &P, int x, int y, int z)
{
  return (unsigned)x - (unsigned)P.lo[0] <= (unsigned)P.ext[0] && (unsigned)y - (unsigned)P.lo[1] <= (unsigned)P.ext[1] &&
         (unsigned)z - (unsigned)P.lo[2] <= (unsigned)P.ext[2];
}

// atomicAdd by ONE lane whose result is consumed much later.  Plain atomicAdd() in divergent code is
// warp-aggregated by the compiler, which appends a SHFL of the result right behind the ATOMG and so
// waits for the full L2 round trip on the spot (18 % of all stall samples of the march kernel).
WS_D unsigned atom_add_async(unsigned *addr, unsigned v)
{
  unsigned old;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
  return old;
}

struct Ray
{
  int p[3];
  int d[3];
  int distance;
  int iv[3];
  int n_steps;
  bool small;          // march_math.cuh: the 32-bit fast path is exact for this ray
};

// What the set-up pass leaves per ray for the march warps: six 16-byte words, staged in shared memory
#define RAY_WORDS 8
#define RAY_SEGS WS_MAX_XIV
struct __align__(16) RaySetup
{
  int p[3], distance;  // word 0
  int d[3], n_steps;   // word 1
  int iv[3], flags;    // word 2: bit 0 = small (fast path), bits 1.. = ray index
  unsigned dq[3], pad0;    // word 3: DDA increments of 32 march steps per axis (fast path), quotient ...
  unsigned drem[3], pad1;  // word 4: ... and remainder of |d| * 32 * h / distance
  int seg[2 * RAY_SEGS];   // words 5-7: march-step ranges [seg[2k], seg[2k+1]) this rank has to process
};
static_assert(sizeof(RaySetup) == 16 * RAY_WORDS, "RaySetup layout");

// read one staged word; volatile asm so the compiler neither hoists it out of the march loops nor keeps the
// per-ray constants in registers across them (register pressure made it re-derive loop invariants in FP64)
WS_D int4 lds_word(const int4 *p)
{
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

// per-ray setup (update_tsdf.cpp:420-446), every lane redundantly (warp-uniform result, no shuffles).
// The truncating 64-bit divisions of the reference run as FP64 estimates with an exact integer
// correction (march_math.cuh) wherever their operands are below 2^46, which they are for any sane ray.
WS_D bool ray_setup(const UpdateParams &P, const ws_pt pt, Ray &r)
{
  r.p[0] = pt.x; r.p[1] = pt.y; r.p[2] = pt.z;
#pragma unroll
  for (int a = 0; a < 3; a++) r.d[a] = wsub(r.p[a], P.pos_mm[a]);                                // :422
  const int sq = wadd(wadd(wmul(r.d[0], r.d[0]), wmul(r.d[1], r.d[1])), wmul(r.d[2], r.d[2]));
  if (sq <= 0) return false;          // distance == 0 (:424) or int32 norm overflow (out of contract)
  r.distance = isqrt31(sq);                                                                      // :423
  if (r.distance == 0) return false;
  if (!box_in_bounds(P, fd_sdiv(r.p[0], P.div_res), fd_sdiv(r.p[1], P.div_res), fd_sdiv(r.p[2], P.div_res)))
    return false;                                                                                // :430-434
  r.small = ray_is_small(r.d, r.p, r.distance, P.tau, P.half_res, P.dz_per_distance, P.coord_lim);

  const double rdist = 1.0 / (double)r.distance;
  i64 nd[3], c1[3];
#pragma unroll
  for (int a = 0; a < 3; a++)                                                                    // :438
  {
    const u64 ad = (u64)(r.d[a] < 0 ? 0u - (unsigned)r.d[a] : (unsigned)r.d[a]);
    const i64 q = (i64)div_rcp64(ad << WS_MR_SHIFT, (unsigned)r.distance, rdist);
    nd[a] = r.d[a] < 0 ? -q : q;
  }
  c1[0] = div_mr64(nd[1] * P.up[2] - nd[2] * P.up[1]);                                           // :439
  c1[1] = div_mr64(nd[2] * P.up[0] - nd[0] * P.up[2]);
  c1[2] = div_mr64(nd[0] * P.up[1] - nd[1] * P.up[0]);
  i64 iv[3];
  iv[0] = nd[1] * c1[2] - nd[2] * c1[1];
  iv[1] = nd[2] * c1[0] - nd[0] * c1[2];
  iv[2] = nd[0] * c1[1] - nd[1] * c1[0];
  const i64 isq = (i64)((u64)iv[0] * (u64)iv[0] + (u64)iv[1] * (u64)iv[1] + (u64)iv[2] * (u64)iv[2]);
  const double fnorm = sqrt(__ll2double_rn(isq));
  const i64 inorm = __double2ll_rz(fnorm);                                                       // :440 (FP64, like Eigen)
  if (inorm <= 0) return false;                                                                  // :441-445
#pragma unroll
  for (int a = 0; a < 3; a++)                                                                    // :446
  {
    const i64 num = iv[a] * WS_MR;
    const u64 an = (u64)(num < 0 ? -num : num);
    u64 q;
    // a correctly rounded FP64 quotient of integers below 2^52 truncates to the exact integer quotient
    if (an < (1ull << 52)) q = (u64)__double2ll_rz(__ll2double_rn((i64)an) / __ll2double_rn(inorm));
    else q = an / (u64)inorm;
    r.iv[a] = (int)(num < 0 ? -(i64)q : (i64)q);
  }
  // len = 1, 1+h, ... <= distance + tau  (:450)
  r.n_steps = (int)fd_udiv((unsigned)(r.distance + P.tau - 1), P.div_half) + 1;
  return true;
}

// general (wrapping, 64-bit magic) projection of one march step
WS_D void step_index(const UpdateParams &P, const int pos_mm[3], const int d[3], const FastDiv div_dist, int len,
                     int proj[3], int idx[3])
{
#pragma unroll
  for (int a = 0; a < 3; a++)
  {
    proj[a] = wadd(pos_mm[a], fd_sdiv(wmul(d[a], len), div_dist));                               // :452
    idx[a] = fd_sdiv(proj[a], P.div_res);                                                        // :453
  }
}

// Record writer.  The record is a sequence of 64-entry chunks with a fill count each.  A warp reserves
// WS_REC_SPAN consecutive chunks at a time (one same-address atomic per 2048 records -- the allocation
// counter is a single L2 word, and same-address atomics serialise), fills them in order, and on its way
// out zeroes the fill counts of the chunks it reserved but did not use.
struct RecWriter
{
  unsigned cur;        // chunk being filled
  unsigned span_end;   // end of the span `cur` belongs to
  unsigned next_span;  // next reserved span (valid in lane 0 once the atomic has returned)
  int fill;            // entries in `cur`
};

WS_D void rec_init(RecWriter &w, UpdateCounters *ctr, int lane)
{
  unsigned b = 0;
  if (lane == 0) b = atomicAdd(&ctr->n_chunks, 2u * WS_REC_SPAN);
  b = __shfl_sync(FULL, b, 0);
  w.cur = b; w.span_end = b + WS_REC_SPAN; w.next_span = b + WS_REC_SPAN; w.fill = 0;
}

// warp-collective append of one record per lane with want == true
WS_D void rec_append(RecWriter &w, const bool want, const u64 key, const u64 addr, const int lane,
                     Rec *__restrict__ rec, unsigned *__restrict__ chunk_fill, const unsigned cap_chunks,
                     UpdateCounters *__restrict__ ctr)
{
  const unsigned m = __ballot_sync(FULL, want);
  if (m == 0u) return;
  const int n = __popc(m);
  const int slot = w.fill + __popc(m & ((1u << lane) - 1u));
  // the chunk after `cur`: the next one of the span, or the first one of the next span
  unsigned after = w.cur + 1u;
  const bool span_done = after == w.span_end;
  if (w.fill + n >= WS_REC_CHUNK && span_done) after = __shfl_sync(FULL, w.next_span, 0);
  if (want)
  {
    const unsigned chunk = slot < WS_REC_CHUNK ? w.cur : after;
    if (chunk < cap_chunks)
    {
      Rec e; e.key = key; e.ref = addr;
      rec[(size_t)chunk * WS_REC_CHUNK + (unsigned)(slot & (WS_REC_CHUNK - 1))] = e;
    }
  }
  w.fill += n;
  if (w.fill >= WS_REC_CHUNK)
  {
    if (lane == 0)
    {
      if (w.cur < cap_chunks) chunk_fill[w.cur] = WS_REC_CHUNK;
      else ctr->rec_overflow = 1u;
    }
    w.cur = after;
    w.fill -= WS_REC_CHUNK;
    if (span_done)
    {
      w.span_end = after + WS_REC_SPAN;
      if (lane == 0) w.next_span = atom_add_async(&ctr->n_chunks, (unsigned)WS_REC_SPAN);   // consumed a span later
    }
  }
}

WS_D void rec_finish(RecWriter &w, const int lane, unsigned *__restrict__ chunk_fill, const unsigned cap_chunks,
                     UpdateCounters *__restrict__ ctr)
{
  const unsigned next_span = __shfl_sync(FULL, w.next_span, 0);
  if (lane == 0)
  {
    if (w.cur < cap_chunks) chunk_fill[w.cur] = (unsigned)w.fill;
    else if (w.fill > 0) ctr->rec_overflow = 1u;
  }
  // unused remainder of the current span and the whole reserved next span
  for (unsigned c = w.cur + 1u + (unsigned)lane; c < w.span_end; c += 32u)
    if (c < cap_chunks) chunk_fill[c] = 0u;
  for (unsigned c = next_span + (unsigned)lane; c < next_span + WS_REC_SPAN; c += 32u)
    if (c < cap_chunks) chunk_fill[c] = 0u;
}

struct MarchCtx
{
  RecWriter rw;
  unsigned n_cand;
  unsigned err;
};

// One ray, one warp.  FAST: march_math.cuh (DDA projection, 32-bit magics, int32 fan arithmetic); !FAST: the
// literal wrapping arithmetic of the oracle.  Both produce identical candidates wherever FAST is allowed.
template <bool ATOMIC, bool FAST>
WS_D void march_ray(const GridDesc &g, const UpdateParams &P, const int pos_mm[3], const int4 *ray_s, const int start,
                    const int end, const int lane,
                    int4 *qa_s, int4 *qb_s, MarchCtx &cx, Rec *__restrict__ rec, unsigned *__restrict__ chunk_fill,
                    const unsigned cap_chunks, UpdateCounters *__restrict__ ctr)
{
  const unsigned lt = (1u << lane) - 1u;
  // the step loop keeps only the direction and the distance in registers; the hit point, the interpolation
  // vector and the ray index are re-read from the staged set-up where the heavy part needs them
  int dvec[3], distance;
  {
    const int4 w0 = lds_word(ray_s + 0), w1 = lds_word(ray_s + 1);
    distance = w0.w;
    dvec[0] = w1.x; dvec[1] = w1.y; dvec[2] = w1.z;
    if (w1.w > (1 << WS_SEQ_MARCH_BITS)) cx.err |= 1u;
  }

  FastDiv div_dist;
  div_dist.d = (unsigned)distance; div_dist.M = 0ull;
  DdaAxis dda[3];
  if (FAST)
  {
    const double rdist = 1.0 / (double)distance;
    const unsigned len0 = 1u + (unsigned)(start + lane) * (unsigned)P.half_res;
    const int4 w4 = lds_word(ray_s + 3), w5 = lds_word(ray_s + 4);
    const unsigned dq[3] = { (unsigned)w4.x, (unsigned)w4.y, (unsigned)w4.z };
    const unsigned dr[3] = { (unsigned)w5.x, (unsigned)w5.y, (unsigned)w5.z };
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
      divrem_rcp((dvec[a] < 0 ? 0u - (unsigned)dvec[a] : (unsigned)dvec[a]) * len0, (unsigned)distance, rdist, dda[a].q, dda[a].rem);
      dda[a].dq = dq[a]; dda[a].drem = dr[a];
    }
  }
  else if (distance > 1) div_dist.M = (~0ull) / (u64)(unsigned)distance + 1ull;

  int carry_x = 0, carry_y = 0;
  if (start > 0)                          // a range that starts inside the ray: the step before it seeds the column filter
  {
    FastDiv dd = div_dist;
    if (FAST && distance > 1) dd.M = (~0ull) / (u64)(unsigned)distance + 1ull;
    int pj[3], ix[3];
    step_index(P, pos_mm, dvec, dd, 1 + (start - 1) * P.half_res, pj, ix);
    carry_x = ix[0]; carry_y = ix[1];
  }

  int qn = 0, qh = 0;                     // queue fill / head (warp-uniform)
  for (int base = start; base < end || qn > 0; base += 32)
  {
    if (base < end)
    {
      const int i = base + lane;
      int proj[3], index[3];
      if (FAST)
      {
#pragma unroll
        for (int a = 0; a < 3; a++)
        {
          proj[a] = pos_mm[a] + (dvec[a] < 0 ? -(int)dda[a].q : (int)dda[a].q);                  // :452
          index[a] = fd32_sdiv(proj[a], P.div_res32);                                            // :453
          dda_advance(dda[a], (unsigned)distance);
        }
      }
      else step_index(P, pos_mm, dvec, div_dist, 1 + i * P.half_res, proj, index);

      int px = __shfl_up_sync(FULL, index[0], 1);
      int py = __shfl_up_sync(FULL, index[1], 1);
      if (lane == 0) { px = carry_x; py = carry_y; }
      carry_x = __shfl_sync(FULL, index[0], 31);
      carry_y = __shfl_sync(FULL, index[1], 31);

      bool active = i < end;
      if (i > 0 && index[0] == px && index[1] == py) active = false;                             // :455-458
      if (!box_in_bounds(P, index[0], index[1], index[2])) active = false;                       // :460-463
      const unsigned m = __ballot_sync(FULL, active);
      if (active)
      {
        const int pos = (qh + qn + __popc(m & lt)) & (QCAP - 1);
        qa_s[pos] = make_int4(proj[0], proj[1], proj[2], i);
        qb_s[pos] = make_int4(index[0], index[1], index[2], 0);
      }
      qn += __popc(m);
      __syncwarp();
      if (qn < 32 && base + 32 < end) continue;           // keep filling
    }
    if (qn == 0) continue;

    // ---- heavy part: up to 32 compacted march steps, one per lane ---------------------------
    const int take = qn < 32 ? qn : 32;
    bool have = lane < take;
    const int4 qa = qa_s[(qh + lane) & (QCAP - 1)];
    const int4 qb = qb_s[(qh + lane) & (QCAP - 1)];
    __syncwarp();
    qh = (qh + take) & (QCAP - 1);
    qn -= take;

    const int4 w0 = lds_word(ray_s + 0), w2 = lds_word(ray_s + 2);
    const int rp[3] = { w0.x, w0.y, w0.z };
    const int riv[3] = { w2.x, w2.y, w2.z };
    const int ray_id = w2.w >> 1;
    const int i = qa.w;
    const int len = 1 + i * P.half_res;
    // distance of the hit to the centre of the marched voxel (:466-472), clamped to tau; the square root is
    // skipped when the whole warp is farther than tau from the hit (free space, most of every ray)
    const int tcx = wadd(wmul(qb.x, P.res), P.half_res);
    const int tcy = wadd(wmul(qb.y, P.res), P.half_res);
    const int tcz = wadd(wmul(qb.z, P.res), P.half_res);
    const int ex = wsub(rp[0], tcx), ey = wsub(rp[1], tcy), ez = wsub(rp[2], tcz);
    const int vsq = wadd(wadd(wmul(ex, ex), wmul(ey, ey)), wmul(ez, ez));
    int value = P.tau;
    if (__any_sync(FULL, (unsigned)vsq < (unsigned)P.tau_sq))       // vsq < 0 (wrapped) reads as "far" too
    {
      const int root = isqrt31(vsq < 0 ? 0 : vsq);
      if ((unsigned)vsq < (unsigned)P.tau_sq) value = root < P.tau ? root : P.tau;
    }
    if (len > distance) value = -value;

    if (value <= P.zero_weight_max) have = false;       // weight == 0 (:475-483); the merge recomputes the weight

    const int delta_z = div_mr32(wmul(P.dz_per_distance, len));                                  // :485

    // ---- near field: every step of this batch has a fan of exactly one voxel (iter_steps == 1, mid == 0)
    // and lies before the recorded range, so there is no fan loop and no record
    if (FAST && ATOMIC && 1 + __shfl_sync(FULL, qa.w, take - 1) * P.half_res < P.far_len)
    {
      const int vx = fd32_sdiv(qa.x - div_mr32(delta_z * riv[0]), P.div_res32);                 // :488,:493
      const int vy = fd32_sdiv(qa.y - div_mr32(delta_z * riv[1]), P.div_res32);
      const int vz = fd32_sdiv(qa.z - div_mr32(delta_z * riv[2]), P.div_res32);
      const unsigned tx = (unsigned)vx - (unsigned)P.lo[0];
      const unsigned ty = (unsigned)vy - (unsigned)P.lo[1];
      const unsigned tz = (unsigned)vz - (unsigned)P.lo[2];
      if (tx > (unsigned)P.ext[0] || ty > (unsigned)P.ext[1] || tz > (unsigned)P.ext[2]) have = false;   // :495-498
      unsigned brick = 0xFFFFFFFFu;
      if (have)
      {
        int rx = (int)tx + P.ringc[0]; rx -= rx >= g.size[0] ? g.size[0] : 0;
        int ry = (int)ty + P.ringc[1]; ry -= ry >= g.size[1] ? g.size[1] : 0;
        int rz = (int)tz + P.ringc[2]; rz -= rz >= g.size[2] ? g.size[2] : 0;
        const int slot = g.full ? (rx >> 3) : (int)g.xslot[rx >> 3];
        if (slot >= 0)
        {
          cx.n_cand++;
          brick = (unsigned)((slot * g.nb[1] + (ry >> 3)) * g.nb[2] + (rz >> 3));
          const u64 addr = (u64)brick * WS_BRICK_VOX + (u64)brick_local(rx, ry, rz);
          const unsigned av = (unsigned)(value < 0 ? -value : value);
          const u64 key = ((u64)av << 47) | (make_seq((unsigned)ray_id, (unsigned)i, 0u) << 1) | (u64)(value < 0 ? 1 : 0);
#ifndef WHATIF_NOATOM
          atomicMin(&g.keys[addr], key);                                                         // :508-512
#endif
        }
      }
      // consecutive lanes are consecutive march steps: mostly the same brick, flag it once
      const unsigned prev_brick = __shfl_up_sync(FULL, brick, 1);
      if (brick != 0xFFFFFFFFu && (lane == 0 || brick != prev_brick)) g.brick_flag[brick] = 1u;
      continue;
    }
    int iter_steps, mid, low_x, low_y, low_z;
    if (FAST)
    {
      iter_steps = (int)fd32_udiv((unsigned)(delta_z * 2), P.div_res32) + 1;                     // :486
      mid = (int)fd32_udiv((unsigned)delta_z, P.div_res32);                                      // :487
      low_x = qa.x - div_mr32(delta_z * riv[0]);                                                // :488
      low_y = qa.y - div_mr32(delta_z * riv[1]);
      low_z = qa.z - div_mr32(delta_z * riv[2]);
    }
    else
    {
      iter_steps = fd_sdiv(delta_z * 2, P.div_res) + 1;
      mid = fd_sdiv(delta_z, P.div_res);
      low_x = wsub(qa.x, (int)div_mr64((i64)delta_z * riv[0]));
      low_y = wsub(qa.y, (int)div_mr64((i64)delta_z * riv[1]));
      low_z = wsub(qa.z, (int)div_mr64((i64)delta_z * riv[2]));
    }
    if (have && iter_steps > (1 << WS_SEQ_STEP_BITS)) cx.err |= 1u;
    const bool far = len >= P.far_len;
    const int max_steps = __reduce_max_sync(FULL, have ? iter_steps : 0);
    unsigned last_brick = 0xFFFFFFFFu;

    // keys of this march step's fan: real candidate at `mid`, interpolated ones around it (:503-506)
    const unsigned av = (unsigned)(value < 0 ? -value : value);
    const u64 seq0 = make_seq((unsigned)ray_id, (unsigned)i, 0u);
    const u64 kbase = ((u64)av << 47) | (u64)(value < 0 ? 1 : 0);
    const u64 key_real0 = kbase | (seq0 << 1);
    const u64 key_int0 = kbase | (1ull << 46) | ((WS_SEQ_MAX - seq0) << 1);

    for (int step = 0; step < max_steps; ++step)                                                 // :491
    {
      bool valid = have && step < iter_steps;
      const int sr = wmul(step, P.res);
      int vx, vy, vz;
      if (FAST)
      {
        vx = fd32_sdiv(low_x + div_mr32(sr * riv[0]), P.div_res32);                             // :493
        vy = fd32_sdiv(low_y + div_mr32(sr * riv[1]), P.div_res32);
        vz = fd32_sdiv(low_z + div_mr32(sr * riv[2]), P.div_res32);
      }
      else
      {
        vx = fd_sdiv(wadd(low_x, (int)div_mr64((i64)sr * riv[0])), P.div_res);
        vy = fd_sdiv(wadd(low_y, (int)div_mr64((i64)sr * riv[1])), P.div_res);
        vz = fd_sdiv(wadd(low_z, (int)div_mr64((i64)sr * riv[2])), P.div_res);
      }
      // in bounds (:495-498) <=> 0 <= v - lo <= size - 1 per axis
      const unsigned tx = (unsigned)vx - (unsigned)P.lo[0];
      const unsigned ty = (unsigned)vy - (unsigned)P.lo[1];
      const unsigned tz = (unsigned)vz - (unsigned)P.lo[2];
      if (tx > (unsigned)P.ext[0] || ty > (unsigned)P.ext[1] || tz > (unsigned)P.ext[2]) valid = false;

      u64 key = 0ull, addr = 0ull;
      bool resident = false;
      if (valid)
      {
        // ring coordinates (hdf5_local_map.h:140-151): (v - pos + offset) mod size
        int rx = (int)tx + P.ringc[0]; rx -= rx >= g.size[0] ? g.size[0] : 0;
        int ry = (int)ty + P.ringc[1]; ry -= ry >= g.size[1] ? g.size[1] : 0;
        int rz = (int)tz + P.ringc[2]; rz -= rz >= g.size[2] ? g.size[2] : 0;
        const int slot = g.full ? (rx >> 3) : (int)g.xslot[rx >> 3];
        if (slot >= 0)                                 // else: the column lives on another rank
        {
          resident = true;
          cx.n_cand++;
          const unsigned brick = (unsigned)((slot * g.nb[1] + (ry >> 3)) * g.nb[2] + (rz >> 3));
          addr = (u64)brick * WS_BRICK_VOX + (u64)brick_local(rx, ry, rz);
          const u64 s2 = (u64)(unsigned)step << 1;
          key = step == mid ? key_real0 + s2 : key_int0 - s2;
          if (ATOMIC)
          {
#ifndef WHATIF_NOATOM
            atomicMin(&g.keys[addr], key);                                                       // :508-512
#endif
            if (brick != last_brick)
            {
              g.brick_flag[brick] = 1u;
              last_brick = brick;
            }
          }
        }
      }
#ifndef WHATIF_NOREC
      rec_append(cx.rw, resident && far, key, addr, lane, rec, chunk_fill, cap_chunks, ctr);
#endif
    }
  }
}

// March steps of a ray whose candidates can land in a resident x column (multi-GPU slabs / stripes): see
// ray_step_ranges in march_math.cuh (host-tested).  The exact per-candidate residency test stays in the march.
WS_D void ray_segments(const UpdateParams &P, const Ray &r, int seg[2 * RAY_SEGS])
{
  ray_step_ranges<RAY_SEGS>(P.n_xiv, P.xiv_lo, P.xiv_hi, P.pos_mm[0], P.half_res, r.d[0], r.distance, r.n_steps, seg);
}

// Set-up pass: one THREAD per ray (the march needs the result warp-uniform; computing it there costs every
// lane of a warp the same ~600 instructions).  Rays without work on this rank get empty step ranges.
__global__ void __launch_bounds__(256)
setup_kernel(const UpdateParams Pin, const ws_pt *__restrict__ pts, RaySetup *__restrict__ rays,
             UpdateCounters *__restrict__ ctr, const PoseDev *__restrict__ pose)
{
  // the sensor pose comes from the host (kernel parameters) or, in the fused per-scan pipeline, from the
  // registration that ran just before on the same stream (device memory)
  UpdateParams P = Pin;
  if (pose)
  {
#pragma unroll
    for (int a = 0; a < 3; a++) { P.pos_mm[a] = pose->pos_mm[a]; P.up[a] = pose->up[a]; }
    P.coord_lim = pose->coord_lim;
  }
  const int ray_id = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  RaySetup o;
  if (ray_id < P.n_points)
  {
    Ray r;
    if (ray_setup(P, pts[ray_id], r))
    {
#pragma unroll
      for (int a = 0; a < 3; a++) { o.p[a] = r.p[a]; o.d[a] = r.d[a]; o.iv[a] = r.iv[a]; }
      o.distance = r.distance; o.n_steps = r.n_steps; o.flags = (ray_id << 1) | (r.small ? 1 : 0);
      o.pad0 = o.pad1 = 0u;
      const double rdist = 1.0 / (double)r.distance;
#pragma unroll
      for (int a = 0; a < 3; a++)
      {
        o.dq[a] = o.drem[a] = 0u;
        const unsigned ad = r.d[a] < 0 ? 0u - (unsigned)r.d[a] : (unsigned)r.d[a];
        if (r.small) divrem_rcp(ad * 32u * (unsigned)P.half_res, (unsigned)r.distance, rdist, o.dq[a], o.drem[a]);
      }
      ray_segments(P, r, o.seg);
      valid = o.seg[1] > o.seg[0];                      // ranges are packed from slot 0
    }
  }
  // Written in place, NOT compacted: the march takes rays in scan order, so the ~3,500 rays in flight at
  // any time belong to a few neighbouring beams and their voxels (keys) stay L2-resident.  A compacted list
  // filled in atomic order scrambles the beams and made the march 40 % slower (same-box A/B).
  if (ray_id < P.n_points)
  {
    if (!valid)
    {
#pragma unroll
      for (int k = 0; k < 2 * RAY_SEGS; k++) o.seg[k] = 0;
      o.flags = ray_id << 1;
    }
    int4 *dst = reinterpret_cast<int4 *>(&rays[ray_id]);
    const int4 *src = reinterpret_cast<const int4 *>(&o);
    if (valid)
    {
      // a fully resident map has one step range per ray: the last two words (ranges 2..5) are not needed
#pragma unroll
      for (int w = 0; w < RAY_WORDS; w++)
        if (w < 6 || P.n_xiv > 0) dst[w] = src[w];
    }
    else { dst[2] = src[2]; dst[5] = src[5]; }
  }
}

// ATOMIC: the scan's first pass (candidate keys + brick flags + record).  !ATOMIC: regenerate the record
// only (far part of every ray), used when the record buffer had to grow.
template <bool ATOMIC>
__global__ void __launch_bounds__(MARCH_THREADS, MARCH_CTAS)
march_kernel(const GridDesc g, const UpdateParams P, const RaySetup *__restrict__ rays,
             UpdateCounters *__restrict__ ctr, Rec *__restrict__ rec,
             unsigned *__restrict__ chunk_fill, const unsigned cap_chunks, const PoseDev *__restrict__ pose)
{
  __shared__ int4 s_qa[MARCH_WARPS][QCAP], s_qb[MARCH_WARPS][QCAP];   // per-warp queue of surviving march steps
  int pos_mm[3];
#pragma unroll
  for (int a = 0; a < 3; a++) pos_mm[a] = pose ? pose->pos_mm[a] : P.pos_mm[a];
  __shared__ int4 s_ray[MARCH_WARPS][2][RAY_WORDS];                    // the warp's current and next ray (cp.async)
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int4 *qa_s = s_qa[wib], *qb_s = s_qb[wib];
  const unsigned total_warps = gridDim.x * MARCH_WARPS;
  const unsigned n_work = (unsigned)P.n_points;

  MarchCtx cx;
  rec_init(cx.rw, ctr, lane);
  cx.n_cand = 0u;
  cx.err = 0u;

  // Work distribution: one ray per fetch from a global counter (same-box A/B of 1/2/4/8 rays per fetch: the
  // tail of long rays costs more than the atomics), software-pipelined two deep -- while ray k is marched
  // out of one shared-memory slot, the set-up of ray k+1 is in flight to the other (cp.async) and the index
  // of ray k+2 is in flight from the counter.
  unsigned k_cur = blockIdx.x * MARCH_WARPS + wib;
  unsigned k_nxt = 0;
  if (lane == 0) k_nxt = atomicAdd(&ctr->ray_counter, 1u);
  k_nxt = total_warps + __shfl_sync(FULL, k_nxt, 0);
  int buf = 0;
  if (lane < RAY_WORDS)
  {
    if (k_cur < n_work)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_ray[wib][0][lane])),
                   "l"(reinterpret_cast<const int4 *>(&rays[k_cur]) + lane) : "memory");
    if (k_nxt < n_work)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_ray[wib][1][lane])),
                   "l"(reinterpret_cast<const int4 *>(&rays[k_nxt]) + lane) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();

  while (k_cur < n_work)
  {
    unsigned fetched = 0;
    if (lane == 0) fetched = atom_add_async(&ctr->ray_counter, 1u);

    const int4 *ray_s = s_ray[wib][buf];
    const int4 w2 = lds_word(ray_s + 2);
    const bool small = (w2.w & 1) != 0;
    const int n_segs = P.n_xiv > 0 ? RAY_SEGS : 1;
#pragma unroll 1
    for (int sg = 0; sg < n_segs; sg++)
    {
      // two ranges per staged word; an empty range ends the list
      const int4 ws = lds_word(ray_s + 5 + (sg >> 1));
      int start = (sg & 1) ? ws.z : ws.x;
      const int end = (sg & 1) ? ws.w : ws.y;
      if (start >= end) break;
      if (!ATOMIC && P.far_len > 1)
      {
        const int fs = (P.far_len - 1) / P.half_res;
        start = start > fs ? start : fs;
      }
      if (start >= end) continue;
      if (small) march_ray<ATOMIC, true>(g, P, pos_mm, ray_s, start, end, lane, qa_s, qb_s, cx, rec, chunk_fill, cap_chunks, ctr);
      else march_ray<ATOMIC, false>(g, P, pos_mm, ray_s, start, end, lane, qa_s, qb_s, cx, rec, chunk_fill, cap_chunks, ctr);
    }

    // rotate the pipeline.  The empty asm ties the fetched index to a value that is only known once the
    // march is done: without it the compiler broadcasts (and waits for) the atomic's result right away.
    asm volatile("" : "+r"(fetched) : "r"(cx.n_cand), "r"(cx.rw.fill));
    asm volatile("cp.async.wait_all;" ::: "memory");     // ray k+1 has landed in the other slot
    __syncwarp();                                        // ... and every lane is done reading this one
    k_cur = k_nxt;
    k_nxt = total_warps + __shfl_sync(FULL, fetched, 0);
    if (k_nxt < n_work && lane < RAY_WORDS)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_ray[wib][buf][lane])),
                   "l"(reinterpret_cast<const int4 *>(&rays[k_nxt]) + lane) : "memory");
    buf ^= 1;
  }

  rec_finish(cx.rw, lane, chunk_fill, cap_chunks, ctr);
  if (ATOMIC)
  {
    // candidate counter (work statistics): warp shuffle, then one atomic per warp
    unsigned long long n_cand = cx.n_cand;
    for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);
    if (lane == 0 && n_cand) atomicAdd(&ctr->n_candidates, n_cand);
  }
  if (cx.err) atomicOr(&ctr->error, cx.err);
}

// touched-brick flags -> compact list; flags are reset for the next scan
__global__ void __launch_bounds__(256)
brick_list_kernel(const GridDesc g, unsigned *__restrict__ brick_list, UpdateCounters *__restrict__ ctr)
{
  const int lane = threadIdx.x & 31;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const i64 n_round = (g.n_bricks + 31) & ~31ll;
  for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < n_round; b += stride)
  {
    const bool set = b < g.n_bricks && g.brick_flag[b] != 0u;
    const unsigned m = __ballot_sync(FULL, set);
    if (m == 0u) continue;
    unsigned first = 0;
    if (lane == 0) first = atomicAdd(&ctr->n_touched_bricks, (unsigned)__popc(m));
    first = __shfl_sync(FULL, first, 0);
    if (set)
    {
      brick_list[first + (unsigned)__popc(m & ((1u << lane) - 1u))] = (unsigned)b;
      g.brick_flag[b] = 0u;
    }
  }
}

// final winner -> grid entry (update_tsdf.cpp:542-560); returns 1 if the reference writes the entry
WS_D unsigned apply_winner_to(uint32_t *slot, const uint32_t e, const UpdateParams &P, const u64 key)
{
  const int value = key_value(key);
  int weight = tsdf_weight(value, P.tau, P.weight_epsilon);
  if (key_interpolated(key)) weight = -weight;
  const uint32_t n = merge_entry(e, value, weight, P.max_weight);
  if (n != e) *slot = n;
  const int ew = entry_weight(e);
  return ((weight > 0 && ew > 0) || (weight != 0 && ew <= 0)) ? 1u : 0u;
}

WS_D bool winner_is_final(u64 key, int tau)
{
  // a real candidate wins outright; an interpolated one at |v| == tau cannot be followed by anything
  // with a larger |v| (values are clamped to tau), so it is the reference's last write as well
  return !key_interpolated(key) || key_abs_value(key) >= tau;
}

#define PEND_SEEN_FLAG 0x8000000000000000ull   // pend_addr: the stored entry has weight > 0

// ---- merge ---------------------------------------------------------------------------------------
// Folds the final winners of every touched brick into the grid (update_tsdf.cpp:542-560), resets the keys,
// writes the parked bitmap and parks the voxels whose winner is an interpolated candidate below tau.
// A touched brick (4 KB of keys + 2 KB of entries, both contiguous) travels as two bulk copies (TMA): cp.async.bulk global -> shared with an mbarrier counting the bytes, three stages per
// CTA so two bricks are always in flight behind the one being folded, and -- after the fold has rewritten
// the stage in place (new entries; keys EMPTY or PENDING|slot) -- two bulk copies shared -> global.  The SM
// issues 4 copy instructions per brick instead of ~1,000 vector loads/stores.
#define MERGE_STAGES 3
#ifndef MERGE_CTAS
#define MERGE_CTAS 6            // CTAs per SM (18.5 KB of stages each); same-box A/B: 2 -> 0.36 ms, 4 -> 0.22, 6 -> 0.19
#endif
#define MERGE_KEY_BYTES (WS_BRICK_VOX * 8)
#define MERGE_ENT_BYTES (WS_BRICK_VOX * 4)

WS_D unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
WS_D void mbar_init(u64 *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
WS_D void mbar_expect_tx(u64 *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
WS_D void mbar_wait(u64 *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
WS_D void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, u64 *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
WS_D void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256, MERGE_CTAS)
merge_kernel(const GridDesc g, const UpdateParams P, const unsigned *__restrict__ brick_list,
                 UpdateCounters *__restrict__ ctr, const unsigned pending_cap,
                 u64 *__restrict__ pend_addr, u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key)
{
  __shared__ __align__(128) u64 s_keys[MERGE_STAGES][WS_BRICK_VOX];
  __shared__ __align__(128) uint32_t s_ent[MERGE_STAGES][WS_BRICK_VOX];
  __shared__ __align__(8) u64 s_bar[MERGE_STAGES];
  __shared__ unsigned s_cnt[8][2];
  __shared__ unsigned s_base;
  const unsigned n_tb = ctr->n_touched_bricks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned touched = 0, written = 0;

  if (tid == 0)
  {
    for (int st = 0; st < MERGE_STAGES; st++) mbar_init(&s_bar[st], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // bricks of this CTA: blockIdx.x, + gridDim.x, ...; iteration n uses stage n % MERGE_STAGES
  const unsigned n_mine = blockIdx.x < n_tb ? (n_tb - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
  if (tid == 0)
    for (unsigned n = 0; n < MERGE_STAGES - 1 && n < n_mine; n++)
    {
      const size_t base = (size_t)brick_list[blockIdx.x + n * gridDim.x] * WS_BRICK_VOX;
      mbar_expect_tx(&s_bar[n], MERGE_KEY_BYTES + MERGE_ENT_BYTES);
      bulk_g2s(s_keys[n], g.keys + base, MERGE_KEY_BYTES, &s_bar[n]);
      bulk_g2s(s_ent[n], g.grid + base, MERGE_ENT_BYTES, &s_bar[n]);
    }

  for (unsigned n = 0; n < n_mine; n++)
  {
    const int st = (int)(n % MERGE_STAGES);
    const unsigned parity = (n / MERGE_STAGES) & 1u;
    const size_t base = (size_t)brick_list[blockIdx.x + n * gridDim.x] * WS_BRICK_VOX;
    if (tid == 0)
    {
      // the stage brick n-1 was stored from becomes the landing zone of brick n + STAGES - 1
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      const unsigned m = n + MERGE_STAGES - 1;
      if (m < n_mine)
      {
        const int sm = (int)(m % MERGE_STAGES);
        const size_t bm = (size_t)brick_list[blockIdx.x + m * gridDim.x] * WS_BRICK_VOX;
        mbar_expect_tx(&s_bar[sm], MERGE_KEY_BYTES + MERGE_ENT_BYTES);
        bulk_g2s(s_keys[sm], g.keys + bm, MERGE_KEY_BYTES, &s_bar[sm]);
        bulk_g2s(s_ent[sm], g.grid + bm, MERGE_ENT_BYTES, &s_bar[sm]);
      }
    }
    mbar_wait(&s_bar[st], parity);

    // fold: two voxels per thread, in place in the stage
    const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(&s_keys[st][2 * tid]);
    const uint2 ee = *reinterpret_cast<const uint2 *>(&s_ent[st][2 * tid]);
    const u64 k2[2] = { kk.x, kk.y };
    uint32_t e2[2] = { ee.x, ee.y };
    bool parked[2];
    unsigned pm[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      const u64 k = k2[j];
      const bool occupied = key_is_candidate(k);
      const bool fin = occupied && winner_is_final(k, P.tau);
      parked[j] = occupied && !fin;
      if (occupied) touched++;
      if (fin)
      {
        const int value = key_value(k);
        int weight = tsdf_weight(value, P.tau, P.weight_epsilon);
        if (key_interpolated(k)) weight = -weight;
        const int ew = entry_weight(e2[j]);
        written += ((weight > 0 && ew > 0) || (weight != 0 && ew <= 0)) ? 1u : 0u;
        e2[j] = merge_entry(e2[j], value, weight, P.max_weight);
      }
      pm[j] = __ballot_sync(FULL, parked[j]);
    }
    *reinterpret_cast<uint2 *>(&s_ent[st][2 * tid]) = make_uint2(e2[0], e2[1]);
    if (lane == 0)
      *reinterpret_cast<uint2 *>(&g.park_bits[park_word((u64)base + 2u * (unsigned)tid)]) = make_uint2(pm[0], pm[1]);
    // slots for the parked voxels -- one global atomic per brick (block-aggregated: same-address atomics
    // serialise in L2; reserving slots in per-CTA chunks instead was slower, the gaps cost the replay more)
    if (lane < 2) s_cnt[warp][lane] = (unsigned)__popc(pm[lane]);
    __syncthreads();
    if (tid == 0)
    {
      unsigned tot = 0;
      for (int w = 0; w < 8; w++)
        for (int q = 0; q < 2; q++) { const unsigned c = s_cnt[w][q]; s_cnt[w][q] = tot; tot += c; }
      s_base = tot ? atomicAdd(&ctr->n_pending, tot) : 0u;
      g.brick_slot_base[base >> 9] = s_base;
    }
    __syncthreads();
    u64 nk[2] = { WS_KEY_EMPTY, WS_KEY_EMPTY };
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      if (parked[j])
      {
        const unsigned slot = s_base + s_cnt[warp][j] + (unsigned)__popc(pm[j] & ((1u << lane) - 1u));
        if (slot < pending_cap)
        {
          pend_addr[slot] = (u64)(base + 2u * (unsigned)tid + (unsigned)j) | (entry_weight(e2[j]) > 0 ? PEND_SEEN_FLAG : 0ull);
          pend_prev[slot] = k2[j];
          pend_key[slot] = WS_KEY_EMPTY;
          // what the replay needs from a parked voxel in ONE load: its slot (per-brick base + rank) and the
          // order of the parked winner (candidates that come later in the reference's order are offered)
          nk[j] = WS_KEY_PENDING_TAG | ((u64)(slot - s_base) << WS_SEQ_BITS) | key_seq(k2[j]);
        }
        else atomicAdd(&ctr->pending_overflow, 1u);
      }
    }
    *reinterpret_cast<ulonglong2 *>(&s_keys[st][2 * tid]) = make_ulonglong2(nk[0], nk[1]);
    // generic-proxy writes to the stage -> visible to the async proxy, then one thread stores the brick
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0)
    {
      bulk_s2g(g.keys + base, s_keys[st], MERGE_KEY_BYTES);
      bulk_s2g(g.grid + base, s_ent[st], MERGE_ENT_BYTES);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

  for (int o = 16; o > 0; o >>= 1)
  {
    touched += __shfl_down_sync(FULL, touched, o);
    written += __shfl_down_sync(FULL, written, o);
  }
  if (lane == 0)
  {
    if (touched) atomicAdd(&ctr->n_touched, (unsigned long long)touched);
    if (written) atomicAdd(&ctr->n_written, (unsigned long long)written);
  }
}

// replay list writer: a warp reserves WS_LIST_SPAN entries at a time (one same-address atomic per span) and
// pads what it does not use with entries whose slot is LIST_NONE
#define WS_LIST_SPAN 128
#define LIST_NONE 0xFFFFFFFFFFFFFFFFull
struct ListWriter
{
  unsigned base, used;
};

WS_D void list_pad(ListWriter &w, const int lane, Rec *__restrict__ list)
{
  for (unsigned t = w.used + (unsigned)lane; t < WS_LIST_SPAN; t += 32u)
  {
    Rec e; e.key = 0ull; e.ref = LIST_NONE;
    list[w.base + t] = e;
  }
  w.used = WS_LIST_SPAN;
}

WS_D void list_append(ListWriter &w, const bool want, const Rec e, const int lane, Rec *__restrict__ list, unsigned *counter)
{
  const unsigned m = __ballot_sync(FULL, want);
  if (m == 0u) return;
  const unsigned n = (unsigned)__popc(m);
  if (w.used + n > WS_LIST_SPAN)
  {
    if (w.used < WS_LIST_SPAN) list_pad(w, lane, list);
    unsigned b = 0;
    if (lane == 0) b = atomicAdd(counter, (unsigned)WS_LIST_SPAN);
    w.base = __shfl_sync(FULL, b, 0);
    w.used = 0;
  }
  if (want) list[w.base + w.used + (unsigned)__popc(m & ((1u << lane) - 1u))] = e;
  w.used += n;
}

// Cooperative launch: settles every parked voxel on the device (see the file header, step 4).
__global__ void __launch_bounds__(256, 4)
replay_kernel(const GridDesc g, const UpdateParams P, UpdateCounters *__restrict__ ctr, const unsigned pending_cap,
              u64 *__restrict__ pend_addr, u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key,
              const Rec *__restrict__ rec, const unsigned *__restrict__ chunk_fill, const unsigned cap_chunks,
              Rec *__restrict__ list, unsigned *__restrict__ active0, unsigned *__restrict__ active1)
{
  cg::grid_group grid = cg::this_grid();
  unsigned n_pend = ctr->n_pending;
  if (n_pend > pending_cap) n_pend = pending_cap;
  if (n_pend == 0u || ctr->rec_overflow != 0u || ctr->pending_overflow != 0u) return;   // grid-uniform

  const int lane = threadIdx.x & 31;
  const unsigned gthread = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned gthreads = gridDim.x * blockDim.x;
  const unsigned gwarp = gthread >> 5, gwarps = gthreads >> 5;
  unsigned written = 0;
  if (gthread == 0) ctr->t_phase[0] = global_ns();

  // ---- round 1, candidates: the record, one warp per pair of 64-entry chunks.  Four independent
  // record -> parked bit -> key -> parked-winner chains per lane are in flight at a time (the pass is
  // latency bound); the parked bitmap (1 bit per voxel, L2 resident) spares the key lookup for the ~80 %
  // of the records whose voxel is not parked.
  unsigned n_chunks = ctr->n_chunks;
  if (n_chunks > cap_chunks) n_chunks = cap_chunks;
  ListWriter lw;
  lw.base = 0u; lw.used = WS_LIST_SPAN;
  for (unsigned c0 = gwarp * 2u; c0 < n_chunks; c0 += gwarps * 2u)
  {
    Rec rr[4];
    bool ok[4];
    u64 kv[4];
    unsigned pw[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      const unsigned c = c0 + (unsigned)(u >> 1);
      const unsigned j = (unsigned)(u & 1) * 32u + (unsigned)lane;
      ok[u] = c < n_chunks && j < chunk_fill[c < n_chunks ? c : c0];
      rr[u].key = 0ull; rr[u].ref = 0ull;
      if (ok[u]) rr[u] = rec[(size_t)c * WS_REC_CHUNK + j];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) pw[u] = ok[u] ? __ldcg(&g.park_bits[park_word(rr[u].ref)]) : 0u;
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      ok[u] = ok[u] && ((pw[u] >> park_bit(rr[u].ref)) & 1u);
      kv[u] = ok[u] ? __ldcg(&g.keys[rr[u].ref]) : WS_KEY_EMPTY;
    }
    unsigned sb[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      ok[u] = ok[u] && key_is_pending(kv[u]);
      sb[u] = ok[u] ? __ldcg(&g.brick_slot_base[rr[u].ref >> 9]) : 0u;        // 1 MB table, L2 resident
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      const unsigned slot = sb[u] + (unsigned)((kv[u] >> WS_SEQ_BITS) & 0x1FFull);
      const bool hit = ok[u] && key_seq(rr[u].key) > (kv[u] & WS_SEQ_MAX);
      if (hit) atomicMin(&pend_key[slot], rr[u].key);
      Rec e; e.key = rr[u].key; e.ref = (u64)slot;
      list_append(lw, hit, e, lane, list, &ctr->n_list);
    }
  }
  if (lw.used < WS_LIST_SPAN) list_pad(lw, lane, list);
  grid.sync();
  if (gthread == 0) ctr->t_phase[1] = global_ns();

  // ---- rounds: resolve, then offer the short list to what is still pending ----------------------
  for (unsigned round = 1;; round++)
  {
    unsigned *act_out = (round & 1u) ? active1 : active0;
    const unsigned *act_in = (round & 1u) ? active0 : active1;
    unsigned *n_out = &ctr->n_active[round & 1u];
    const unsigned n_in = round == 1u ? n_pend : ctr->n_active[(round & 1u) ^ 1u];
    const unsigned n_in_round = (n_in + 31u) & ~31u;
    // four slots per thread and turn, every level of the dependent chain (slot -> keys -> entry) issued for
    // all four before the first result is used: the pass is a latency-bound pointer chase
    for (unsigned t0 = gthread; t0 < n_in_round; t0 += 4u * gthreads)
    {
      unsigned sl[4];
      u64 pa[4], k2[4], pv[4], fin[4];
      bool live[4], still[4], touch[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
        const unsigned t = t0 + (unsigned)u * gthreads;
        live[u] = t < n_in;
        sl[u] = live[u] ? (round == 1u ? t : act_in[t]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
        pa[u] = live[u] ? pend_addr[sl[u]] : PEND_DONE;
        k2[u] = live[u] ? pend_key[sl[u]] : 0ull;
        pv[u] = live[u] ? pend_prev[sl[u]] : 0ull;
      }
      uint32_t ent[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
        live[u] = live[u] && pa[u] != PEND_DONE;
        still[u] = false; touch[u] = false; fin[u] = 0ull;
        if (live[u])
        {
          if (k2[u] == WS_KEY_EMPTY) fin[u] = pv[u];            // nothing later in the order: the parked winner stands
          else if (winner_is_final(k2[u], P.tau)) fin[u] = k2[u];
          else still[u] = true;                                 // still interpolated: restart after it
          // an interpolated (negative-weight) winner never changes an entry that already has weight > 0
          // (update_tsdf.cpp:546-557): no grid access at all for those
          touch[u] = !still[u] && !(key_interpolated(fin[u]) && (pa[u] & PEND_SEEN_FLAG));
        }
        ent[u] = touch[u] ? g.grid[pa[u] & ~PEND_SEEN_FLAG] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
        if (live[u])
        {
          if (still[u]) { pend_prev[sl[u]] = k2[u]; pend_key[sl[u]] = WS_KEY_EMPTY; }
          else
          {
            if (touch[u]) written += apply_winner_to(&g.grid[pa[u] & ~PEND_SEEN_FLAG], ent[u], P, fin[u]);
            // the voxel's key keeps its PENDING tag: the next scan's atomicMin / merge treat it as empty
            pend_addr[sl[u]] = PEND_DONE;
          }
        }
        const unsigned m = __ballot_sync(FULL, still[u]);
        if (m)
        {
          unsigned first = 0;
          if (lane == 0) first = atomicAdd(n_out, (unsigned)__popc(m));
          first = __shfl_sync(FULL, first, 0);
          if (still[u]) act_out[first + (unsigned)__popc(m & ((1u << lane) - 1u))] = sl[u];
        }
      }
    }
    grid.sync();
    const unsigned n_still = *((volatile unsigned *)n_out);
    if (gthread == 0)
    {
      if (round == 1u) ctr->t_phase[2] = global_ns();
      ctr->rounds = round;
      ctr->n_active[(round & 1u) ^ 1u] = 0u;      // next round's output counter (its readers are done)
    }
    if (n_still == 0u) break;
    const unsigned n_list = ctr->n_list;
    for (unsigned t = gthread; t < n_list; t += gthreads)
    {
      const Rec e = list[t];
      if (e.ref == LIST_NONE) continue;
      const unsigned slot = (unsigned)e.ref;
      if (pend_addr[slot] != PEND_DONE && key_seq(e.key) > key_seq(pend_prev[slot])) atomicMin(&pend_key[slot], e.key);
    }
    grid.sync();
  }

  for (int o = 16; o > 0; o >>= 1) written += __shfl_down_sync(FULL, written, o);
  if (lane == 0 && written) atomicAdd(&ctr->n_written, (unsigned long long)written);
  if (gthread == 0)
  {
    ctr->n_parked = n_pend;
    ctr->n_pending = 0u;
    ctr->t_phase[3] = global_ns();
  }
}

// between a record overflow and its regeneration
__global__ void rec_reset_kernel(UpdateCounters *ctr)
{
  ctr->n_chunks = 0u;
  ctr->rec_overflow = 0u;
  ctr->ray_counter = 0u;
  ctr->n_list = 0u;
}

}  // namespace

static int far_start_len(int res, int dz)
{
  // interpolated candidates need iter_steps >= 2  <=>  delta_z >= ceil(res/2)  <=>  len >= L0.
  // A voxel parked by such a candidate lies within ~5*res of the march point that produced it, so any
  // candidate that can land on it comes from len >= L0 - 10*res - 16 (DESIGN.md "Replay rounds").
  if (dz <= 0) return 1;
  long long need = (long long)((res + 1) / 2) * WS_MR;
  long long L0 = (need + dz - 1) / dz;
  long long fl = L0 - 10LL * res - 16;
  return fl < 1 ? 1 : (int)(fl > 0x7fffffff ? 0x7fffffff : fl);
}

// Multi-GPU slabs: the x intervals (mm, map frame) in which a march step can still produce a candidate for a
// resident brick column -- the resident voxel columns widened by the reach of the interpolation fan
// (|lowest - proj| <= delta_z and the fan spans 2*delta_z + res more, update_tsdf.cpp:485-493) plus two voxels
// for the truncating divisions.  Resident columns are cyclically contiguous in ring space (slab + halo), so
// there are at most two intervals for a slab; round-robin stripes give one per stripe (up to WS_MAX_XIV, beyond
// that the culling is switched off).
static void resident_x_intervals(const ws_handle *h, UpdateParams &P)
{
  P.n_xiv = 0;
  const GridDesc &g = h->g;
  if (g.full) return;
  const int size = g.size[0];
  std::vector<char> res_t((size_t)size, 0);
  for (int t = 0; t < size; t++)
  {
    int ring = t + P.ringc[0];
    if (ring >= size) ring -= size;
    res_t[(size_t)t] = g.xslot[ring >> 3] >= 0;
  }
  const double len_max = 1.7320508 * (double)std::max(g.size[0], std::max(g.size[1], g.size[2])) * h->res + P.tau + h->res;
  const long long dz_max = (long long)((double)P.dz_per_distance * len_max / WS_MR) + 1;
  const int margin = (int)((3 * dz_max + h->res) / h->res) + 3;
  struct Iv { long long lo, hi; };
  std::vector<Iv> ivs;
  for (int t = 0; t < size; )
  {
    if (!res_t[(size_t)t]) { t++; continue; }
    int e = t;
    while (e < size && res_t[(size_t)e]) e++;
    Iv v;
    v.lo = ((long long)P.lo[0] + t - margin) * h->res;
    v.hi = ((long long)P.lo[0] + e + margin) * h->res;
    if (!ivs.empty() && v.lo <= ivs.back().hi + 4LL * h->res) ivs.back().hi = v.hi;
    else ivs.push_back(v);
    t = e;
  }
  if (ivs.empty() || ivs.size() > WS_MAX_XIV) return;  // nothing resident cannot happen; too many: no culling
  for (auto &v : ivs)
    if (v.lo < -(1ll << 30) || v.hi > (1ll << 30)) return;
  P.n_xiv = (int)ivs.size();
  for (size_t k = 0; k < ivs.size(); k++) { P.xiv_lo[k] = (int)ivs[k].lo; P.xiv_hi[k] = (int)ivs[k].hi; }
}

static void ensure_record(ws_handle *h, size_t chunks)
{
  if (chunks <= h->rec_cap_chunks) return;
  WS_CUDA_OK(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_rec); cudaFree(h->d_list); cudaFree(h->d_chunk_fill);
  h->d_rec = nullptr; h->d_list = nullptr; h->d_chunk_fill = nullptr; h->rec_cap_chunks = 0;
  WS_CUDA_OK(cudaMalloc(&h->d_rec, chunks * WS_REC_CHUNK * sizeof(Rec)));
  // replay list: every record can land in it, spans are padded (< 32 of 128 entries) and every warp of the
  // replay grid (<= 4 blocks x 8 warps per SM) may leave one span partly used
  const size_t list_entries = chunks * WS_REC_CHUNK * 3 / 2 + (size_t)h->sm_count * 4 * 8 * WS_LIST_SPAN;
  WS_CUDA_OK(cudaMalloc(&h->d_list, list_entries * sizeof(Rec)));
  WS_CUDA_OK(cudaMalloc(&h->d_chunk_fill, chunks * sizeof(unsigned)));
  h->rec_cap_chunks = chunks;
}

static void launch_replay(ws_handle *h, const UpdateParams &P)
{
  if (h->replay_blocks == 0)
  {
    int per_sm = 0;
    WS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, replay_kernel, 256, 0));
    if (per_sm < 1) throw std::runtime_error("replay_kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;
    h->replay_blocks = per_sm * h->sm_count;
  }
  unsigned cap_chunks = (unsigned)h->rec_cap_chunks;
  void *args[] = { (void *)&h->g, (void *)&P, (void *)&h->d_counters, (void *)&h->pending_cap,
                   (void *)&h->d_pend_addr, (void *)&h->d_pend_prev, (void *)&h->d_pend_key,
                   (void *)&h->d_rec, (void *)&h->d_chunk_fill, (void *)&cap_chunks,
                   (void *)&h->d_list, (void *)&h->d_active[0], (void *)&h->d_active[1] };
  WS_CUDA_OK(cudaLaunchCooperativeKernel((const void *)replay_kernel, dim3(h->replay_blocks), dim3(256), args, 0, h->stream));
  h->launches++;
}

// src/warpsense/tsdf_mapping.cpp:77-85 + include/util/util.h:52-56 on the device: pose = X * prior (float32, the
// accumulation order of ws_compose_pose_host), scanner voxel = floor(t / res), up = third column of
// to_int_mat(pose) -- so update_tsdf can follow register_cloud on the stream without the host in between
__global__ void pose_kernel(const float *__restrict__ X, const float *__restrict__ prior, const int res,
                            const int coord_lim, PoseDev *__restrict__ out)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float pose[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
    {
      float acc = X[0 * 4 + r] * prior[c * 4 + 0];
      acc = acc + X[1 * 4 + r] * prior[c * 4 + 1];
      acc = acc + X[2 * 4 + r] * prior[c * 4 + 2];
      acc = acc + X[3 * 4 + r] * prior[c * 4 + 3];
      pose[c * 4 + r] = acc;
    }
  bool ok = true;
  for (int a = 0; a < 3; a++)
  {
    const int vox = (int)floorf(pose[12 + a] / (float)res);
    out->pos_mm[a] = (int)((unsigned)vox * (unsigned)res);
    out->up[a] = (long long)(int)(pose[8 + a] * (float)WS_MR);
    const long long ap = out->pos_mm[a] < 0 ? -(long long)out->pos_mm[a] : (long long)out->pos_mm[a];
    if (ap >= (long long)coord_lim) ok = false;
    out->pos_vox[a] = vox;
  }
  out->coord_lim = ok ? coord_lim : 0;
  for (int i = 0; i < 16; i++) out->pose[i] = pose[i];
}

void ws_compose_pose_host(const float X[16], const float prior[16], float pose[16])
{
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
    {
      float acc = X[0 * 4 + r] * prior[c * 4 + 0];
      acc = acc + X[1 * 4 + r] * prior[c * 4 + 1];
      acc = acc + X[2 * 4 + r] * prior[c * 4 + 2];
      acc = acc + X[3 * 4 + r] * prior[c * 4 + 3];
      pose[c * 4 + r] = acc;
    }
}

void ws_launch_pose(ws_handle *h, const float *d_X, const float prior[16])
{
  if (!h->d_pose)
  {
    WS_CUDA_OK(cudaMalloc(&h->d_pose, sizeof(PoseDev) + 16 * sizeof(float)));
    WS_CUDA_OK(cudaMallocHost(&h->h_pose, sizeof(PoseDev) + 16 * sizeof(float)));
  }
  float *d_prior = reinterpret_cast<float *>(reinterpret_cast<char *>(h->d_pose) + sizeof(PoseDev));
  float *h_prior = reinterpret_cast<float *>(reinterpret_cast<char *>(h->h_pose) + sizeof(PoseDev));
  std::memcpy(h_prior, prior, 16 * sizeof(float));
  WS_CUDA_OK(cudaMemcpyAsync(d_prior, h_prior, 16 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  long long lim = (1ll << 31) / h->res - h->tau - (1ll << 17);
  if (lim < 0 || std::getenv("WS_MARCH_GENERAL")) lim = 0;
  pose_kernel<<<1, 32, 0, h->stream>>>(d_X, d_prior, h->res, (int)lim, static_cast<PoseDev *>(h->d_pose));
  h->launches++;
}

void ws_launch_update(ws_handle *h, const ws_pt *d_pts, int n, const int scanner_pos_in[3], const int up_in[3],
                      bool pose_on_device)
{
  // pose_on_device: the scanner pose was left in h->d_pose by ws_launch_pose on this stream; the host copy
  // (needed only if the candidate record has to be regenerated) is read back with the counters
  const PoseDev *d_pose = pose_on_device ? static_cast<const PoseDev *>(h->d_pose) : nullptr;
  const int zero3[3] = { 0, 0, 0 };
  const int *scanner_pos = pose_on_device ? zero3 : scanner_pos_in;
  const int *up = pose_on_device ? zero3 : up_in;
  UpdateParams P{};
  P.tau = h->tau;
  P.max_weight = h->max_weight;
  P.res = h->res;
  P.weight_epsilon = h->tau / 10;                                      // update_tsdf.cpp:403
  {
    float angle = 45.f / 128.f;                                        // update_tsdf.cpp:400-401
    P.dz_per_distance = (int)(tan((double)(angle / 180) * M_PI) / 2.0 * WS_MR);
  }
  for (int a = 0; a < 3; a++)
  {
    P.pos_mm[a] = (int)((unsigned)scanner_pos[a] * (unsigned)h->res);  // update_tsdf.cpp:410
    P.up[a] = up[a];
  }
  P.div_res = make_fastdiv((unsigned)h->res);
  P.div_weps = make_fastdiv((unsigned)(P.tau - P.weight_epsilon > 0 ? P.tau - P.weight_epsilon : 1));
  // largest value whose weight (:475-479) truncates to 0: such march steps are skipped (:480-483)
  P.zero_weight_max = -P.tau - 1;
  for (int v = -P.tau; v < -P.weight_epsilon && tsdf_weight(v, P.tau, P.weight_epsilon) == 0; v++) P.zero_weight_max = v;
  P.half_res = h->res / 2;
  P.div_half = make_fastdiv((unsigned)P.half_res);
  P.tau_sq = P.tau * P.tau;
  P.div_res32 = make_fastdiv32((unsigned)h->res);
  // fast-path bound on |coordinate| (march_math.cuh): every dividend of a 32-bit magic stays below 2^32 / res
  {
    const long long lim = (1ll << 31) / h->res - P.tau - (1ll << 17);
    P.coord_lim = lim > 0 ? (int)lim : 0;
    for (int a = 0; a < 3; a++)
      if (std::llabs((long long)P.pos_mm[a]) >= lim) P.coord_lim = 0;
    if (std::getenv("WS_MARCH_GENERAL")) P.coord_lim = 0;   // tests: force the literal-arithmetic path
  }
  for (int a = 0; a < 3; a++)
  {
    const GridDesc &g = h->g;
    P.lo[a] = g.pos[a] - g.half[a];
    P.ext[a] = g.size[a] - 1;
    P.ringc[a] = (int)((((long long)g.offset[a] - g.half[a]) % g.size[a] + g.size[a]) % g.size[a]);
  }
  P.n_points = n;
  P.far_len = far_start_len(h->res, P.dz_per_distance);
  resident_x_intervals(h, P);

  cudaStream_t s = h->stream;
  WS_CUDA_OK(cudaMemsetAsync(h->d_counters, 0, sizeof(UpdateCounters), s));
  if (n > 0)
  {
    const int march_blocks = h->sm_count * MARCH_CTAS;
    unsigned cap_chunks = (unsigned)h->rec_cap_chunks;
    ws_timer_begin(h, WS_TIMER_MARCH);
    if ((size_t)n > h->rays_cap)
    {
      WS_CUDA_OK(cudaStreamSynchronize(s));
      cudaFree(h->d_rays);
      h->d_rays = nullptr; h->rays_cap = 0;
      const size_t want = std::max<size_t>((size_t)n, 1 << 17);
      WS_CUDA_OK(cudaMalloc(&h->d_rays, want * sizeof(RaySetup)));
      h->rays_cap = want;
    }
    RaySetup *rays = static_cast<RaySetup *>(h->d_rays);
    setup_kernel<<<(n + 255) / 256, 256, 0, s>>>(P, d_pts, rays, h->d_counters, d_pose);
    march_kernel<true><<<march_blocks, MARCH_THREADS, 0, s>>>(h->g, P, rays, h->d_counters, h->d_rec,
                                                              h->d_chunk_fill, cap_chunks, d_pose);
    h->launches++;
    ws_timer_end(h);
    ws_timer_begin(h, WS_TIMER_MERGE);
    brick_list_kernel<<<h->sm_count * 4, 256, 0, s>>>(h->g, h->d_brick_list, h->d_counters);
    merge_kernel<<<h->sm_count * MERGE_CTAS, 256, 0, s>>>(h->g, P, h->d_brick_list, h->d_counters, h->pending_cap,
                                                          h->d_pend_addr, h->d_pend_prev, h->d_pend_key);
    ws_timer_end(h);
    h->launches += 3;
    ws_timer_begin(h, WS_TIMER_REPLAY);
    launch_replay(h, P);
    ws_timer_end(h);
    WS_CUDA_OK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
    if (pose_on_device)
      WS_CUDA_OK(cudaMemcpyAsync(h->h_pose, h->d_pose, sizeof(PoseDev), cudaMemcpyDeviceToHost, s));
    WS_CUDA_OK(cudaStreamSynchronize(s));
    // the record did not fit: grow it, regenerate it from the far part of every ray, settle again
    int guard = 0;
    while (h->h_counters->rec_overflow != 0u && h->h_counters->pending_overflow == 0u)
    {
      if (++guard > 8) throw std::runtime_error("update_tsdf: candidate record keeps overflowing");
      size_t want = (size_t)h->h_counters->n_chunks + (size_t)h->h_counters->n_chunks / 4 + 1024;
      if (want > h->rec_max_chunks) want = h->rec_max_chunks;
      if (want <= h->rec_cap_chunks)
        throw std::runtime_error("update_tsdf: candidate record exceeds WS_RECORD_MAX");
      ensure_record(h, want);
      cap_chunks = (unsigned)h->rec_cap_chunks;
      rec_reset_kernel<<<1, 1, 0, s>>>(h->d_counters);
      march_kernel<false><<<march_blocks, MARCH_THREADS, 0, s>>>(h->g, P, rays, h->d_counters, h->d_rec,
                                                                 h->d_chunk_fill, cap_chunks, d_pose);
      h->launches += 2;
      launch_replay(h, P);
      WS_CUDA_OK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
      WS_CUDA_OK(cudaStreamSynchronize(s));
      h->record_regrows++;
    }
  }
  else
  {
    WS_CUDA_OK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
    if (pose_on_device)
      WS_CUDA_OK(cudaMemcpyAsync(h->h_pose, h->d_pose, sizeof(PoseDev), cudaMemcpyDeviceToHost, s));
    WS_CUDA_OK(cudaStreamSynchronize(s));
  }
  WS_CUDA_OK(cudaGetLastError());
  h->last_counters = *h->h_counters;
  if (h->last_counters.pending_overflow)
    throw std::runtime_error("update_tsdf: pending-voxel capacity exceeded (raise WS_PENDING_CAP)");
  if (h->last_counters.n_pending != 0u)
    throw std::runtime_error("update_tsdf: parked voxels were left unsettled");
  if (h->last_counters.error & 1u)
    throw std::runtime_error("update_tsdf: ray too long for the candidate order field (march steps > 32768 or fan > 64)");
}

void ws_update_alloc(ws_handle *h, size_t initial_chunks, size_t max_chunks)
{
  h->rec_max_chunks = max_chunks;
  ensure_record(h, initial_chunks < max_chunks ? initial_chunks : max_chunks);
}
