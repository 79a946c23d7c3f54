// update_tsdf.cu -- TSDF volume update for sm_100a.
//
// Replaces the reference's TSDFCuda::update_tsdf (src/warpsense/cuda/update_tsdf.cu:13-166 under
// /root/reference) with kernels whose RESULT equals the reference's CPU path
// (src/cpu/update_tsdf.cpp:397-564), which is the parity target.
//
// Pipeline per scan (three streams that fork from and join the handle's stream; nothing waits on the host):
//   0. setup_kernel             one THREAD per ray: the registered cloud's transform, direction, distance,
//                               interpolation vector, DDA increments, fast-path flag, the split between the free-space
//                               and the surface part of the ray and -- on a sharded map -- the march-step ranges that
//                               can reach this rank's columns; 128-byte RaySetup per ray, in scan order; per group of
//                               32 rays the blocks of 64 march steps that hold work, per phase.  item_scan_kernel
//                               turns the block counts into work items (prefix sums, one CTA per item table).
//   1. march_lockstep_kernel    one RAY per LANE, a work item = (group of 32 rays, block of 64 steps) from a global
//      <surface>                counter.  Step phase per lane: exact DDA, 32-bit magic division, the reference's "same
//                               (x,y) column as the previous step" filter (:455-458) and the bounds test against the
//                               lane's own previous step; survivors go through a per-warp shared-memory queue so the
//                               heavy part (value, fan of interpolated voxels, addressing through three shared-memory
//                               tables) runs on full batches.  Every surface candidate (voxel, value, real |
//                               interpolated, order) is ONE 64-bit atomicMin (RED, no return) on the voxel's key --
//                               see ws_common.cuh / DESIGN.md for why min over (|value|, interpolated, order)
//                               reproduces the sequential rule at :508-512 -- plus a plain store that flags the
//                               voxel's 8x8x8 brick; far-field candidates (the only ones that can meet an
//                               interpolated winner) are also appended to a record in 64-entry chunks owned by the warp.
//      <free space, near field> beside it on the second stream: value == tau by construction, a candidate is one plain
//                               byte store ("a real / an interpolated free-space candidate arrived"), no key, no record.
//      march_kernel             the literal wrapping arithmetic for rays outside the 32-bit fast path (none on a sane scan).
//   2. merge_kernel             every CTA compacts its share of the surface-brick flags, then moves those bricks (4 KB
//                               keys + 2 KB entries each) through shared memory with TMA bulk copies (cp.async.bulk +
//                               mbarrier, three stages per CTA), folds final winners into the grid (:542-560), resets the
//                               keys, and parks the voxels whose winner is an interpolated candidate below tau
//                               ("pending"): their key word then carries the slot rank within the brick and the order
//                               of the parked winner.
//   3. march_lockstep_kernel    <free space, far field>: as the near field, but a candidate that lands on a parked
//                               voxel is offered to the replay on the spot.  Beside it, on the third stream,
//      replay_scan_kernel       streams the record once and keeps, per pending voxel, the minimum key among the
//                               candidates that follow the parked winner in the reference's order; those candidates
//                               are also compacted into a short list that later rounds stream instead.
//   4. replay_kernel            cooperative (grid-synchronised): resolve / re-offer rounds on the device until every
//                               pending voxel has its final winner.
//   5. fmerge_kernel            every touched brick (flags compacted by the CTAs themselves): voxels that are not closed
//                               and hold a free-space candidate get (tau, +-WEIGHT_RESOLUTION) folded in; the per-scan
//                               state is cleared.
// If the record overflows its buffer the host grows it and regenerates it with the <surface, no atomics> instantiation
// (far part of every ray) before the far field, the replay and the final merge run again.
// The scanner pose arrives as kernel parameters or, in the fused per-scan pipeline (ws_track_submit), from
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
#ifndef LS_BLOCK
#define LS_BLOCK 64          // march steps per work item of the lockstep march
#endif
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
  unsigned dq[3], pad0;    // word 3: DDA increments of ONE march step per axis (fast path), quotient ...; pad0 = first step of the surface phase
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

// A record chunk (1 KB): 64 keys (512 B), the low words of the 64 voxel addresses (256 B) and -- for maps beyond 2^32
// voxels -- their high words (256 B; never touched otherwise, so a record costs 12 bytes of traffic, and the replay's
// record pass reads the keys only where a voxel turns out to be parked).
WS_D u64 *rec_keys(Rec *rec, unsigned chunk) { return reinterpret_cast<u64 *>(rec + (size_t)chunk * WS_REC_CHUNK); }
WS_D const u64 *rec_keys(const Rec *rec, unsigned chunk) { return reinterpret_cast<const u64 *>(rec + (size_t)chunk * WS_REC_CHUNK); }
WS_D unsigned *rec_refs(Rec *rec, unsigned chunk) { return reinterpret_cast<unsigned *>(rec_keys(rec, chunk) + WS_REC_CHUNK); }
WS_D const unsigned *rec_refs(const Rec *rec, unsigned chunk) { return reinterpret_cast<const unsigned *>(rec_keys(rec, chunk) + WS_REC_CHUNK); }

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
WS_D void rec_append(RecWriter &w, const bool want, const u64 key, const u64 addr, const bool wide, const int lane,
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
      const unsigned j = (unsigned)(slot & (WS_REC_CHUNK - 1));
      rec_keys(rec, chunk)[j] = key;
      rec_refs(rec, chunk)[j] = (unsigned)addr;
      if (wide) rec_refs(rec, chunk)[WS_REC_CHUNK + j] = (unsigned)(addr >> 32);    // maps beyond 2^32 voxels only
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
      rec_append(cx.rw, resident && far, key, addr, g.wide != 0, lane, rec, chunk_fill, cap_chunks, ctr);
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
setup_kernel(const UpdateParams Pin, ws_pt *__restrict__ pts, RaySetup *__restrict__ rays,
             uint2 *__restrict__ grp_info, const unsigned grp_stride, unsigned *__restrict__ gen_list,
             UpdateCounters *__restrict__ ctr, const PoseDev *__restrict__ pose, const float *__restrict__ xform)
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
    ws_pt pt = pts[ray_id];
    if (xform)
    {
      // the registered cloud (registration.cpp:164-174): transformed here instead of by a kernel of its own, and
      // written back for the callers that read it (ws_reg_points_device)
      float T[16];
#pragma unroll
      for (int i = 0; i < 16; i++) T[i] = __ldg(&xform[i]);
      int M[16], q[3];
      to_int_mat(T, M);
      transform_point(M, pt.x, pt.y, pt.z, q);
      pt.x = q[0]; pt.y = q[1]; pt.z = q[2];
      pts[ray_id] = pt;
    }
    if (ray_setup(P, pt, r))
    {
#pragma unroll
      for (int a = 0; a < 3; a++) { o.p[a] = r.p[a]; o.d[a] = r.d[a]; o.iv[a] = r.iv[a]; }
      o.distance = r.distance; o.n_steps = r.n_steps; o.flags = (ray_id << 1) | (r.small ? 1 : 0);
      o.pad1 = 0u;
      const double rdist = 1.0 / (double)r.distance;
#pragma unroll
      for (int a = 0; a < 3; a++)
      {
        o.dq[a] = o.drem[a] = 0u;
        const unsigned ad = r.d[a] < 0 ? 0u - (unsigned)r.d[a] : (unsigned)r.d[a];
        if (r.small) divrem_rcp(ad * (unsigned)P.half_res, (unsigned)r.distance, rdist, o.dq[a], o.drem[a]);
      }
      ray_segments(P, r, o.seg);
      valid = o.seg[1] > o.seg[0];                      // ranges are packed from slot 0
      // first march step of the surface phase: len <= distance - tau - 3 * res - 4 keeps the centre of the
      // marched voxel (at most 1.5 * res * sqrt(3) + sqrt(3) from the ray point, truncation included) farther
      // than tau from the hit, so min(norm, tau) == tau (update_tsdf.cpp:466-468)
      {
        const long long lim = (long long)r.distance - P.tau - 3ll * P.res - 4ll;
        o.pad0 = lim < 1 ? 0u : (unsigned)((lim - 1) / P.half_res + 1);
        if (o.pad0 > (unsigned)r.n_steps) o.pad0 = (unsigned)r.n_steps;
      }
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
  // The lockstep march's work items, per phase: the step blocks in which any of the group's (= this warp's)
  // fast-path rays has work.  A ray's steps split at `split` (word 3, pad0): before it the marched voxel is
  // farther than tau from the hit whatever the rounding (free-space phase: value == tau), from it on the
  // surface phase computes the value.  Rays outside the fast path go on a list for march_kernel.
  const bool small = valid && (o.flags & 1);
  int lo_s = 0x7fffffff, hi_s = 0, lo_f = 0x7fffffff, hi_f = 0;
  if (small)
  {
    const int split = (int)o.pad0;
#pragma unroll
    for (int k = 0; k < RAY_SEGS; k++)
    {
      const int a = o.seg[2 * k], b = o.seg[2 * k + 1];
      if (b <= a) continue;
      const int as = a > split ? a : split, bf = b < split ? b : split;
      if (b > as) { lo_s = lo_s < as ? lo_s : as; hi_s = b; }
      if (bf > a) { lo_f = lo_f < a ? lo_f : a; hi_f = bf; }
    }
  }
  lo_s = __reduce_min_sync(FULL, lo_s); hi_s = __reduce_max_sync(FULL, hi_s);
  lo_f = __reduce_min_sync(FULL, lo_f); hi_f = __reduce_max_sync(FULL, hi_f);
  if ((threadIdx.x & 31) == 0 && ray_id < P.n_points)
  {
    unsigned b_lo = hi_s > lo_s ? (unsigned)lo_s / LS_BLOCK : 0u;
    unsigned nb = hi_s > lo_s ? ((unsigned)hi_s + LS_BLOCK - 1u) / LS_BLOCK - b_lo : 0u;
    grp_info[ray_id >> 5] = make_uint2(b_lo, nb);
    // free space: the blocks before the far field need nothing from the surface phase, the others wait for
    // its merge (they look for parked voxels)
    b_lo = hi_f > lo_f ? (unsigned)lo_f / LS_BLOCK : 0u;
    const unsigned b_hi = hi_f > lo_f ? ((unsigned)hi_f + LS_BLOCK - 1u) / LS_BLOCK : 0u;
    const unsigned b_far = (unsigned)P.far_block, b_split = (unsigned)P.near_split_block;
    const unsigned near_hi = b_hi < b_far ? b_hi : b_far, far_lo = b_lo > b_far ? b_lo : b_far;
    // the near field in two parts: blocks before near_split_block run beside the surface phase, the rest beside the
    // far field (whose launch otherwise leaves 40 % of the issue slots idle)
    const unsigned a_hi = near_hi < b_split ? near_hi : b_split, b2_lo = b_lo > b_split ? b_lo : b_split;
    grp_info[grp_stride + (ray_id >> 5)] = make_uint2(b_lo, a_hi > b_lo ? a_hi - b_lo : 0u);
    grp_info[2u * grp_stride + (ray_id >> 5)] = make_uint2(far_lo, b_hi > far_lo ? b_hi - far_lo : 0u);
    grp_info[3u * grp_stride + (ray_id >> 5)] = make_uint2(b2_lo, near_hi > b2_lo ? near_hi - b2_lo : 0u);
  }
  if (valid && !small) gen_list[atomicAdd(&ctr->n_general, 1u)] = (unsigned)ray_id;
}

// The literal-arithmetic march for the rays the set-up pass found outside the bounds of the 32-bit fast path
// (gen_list; none on a sane scan -- the kernel then returns at once): one ray per warp, lanes stride over
// the march steps, survivors of the column filter compacted through a per-warp shared-memory queue.
// ATOMIC: the scan's first pass (candidate keys + brick flags + record).  !ATOMIC: regenerate the record
// only (far part of every ray), used when the record buffer had to grow.
template <bool ATOMIC>
__global__ void __launch_bounds__(MARCH_THREADS, MARCH_CTAS)
march_kernel(const GridDesc g, const UpdateParams P, const RaySetup *__restrict__ rays,
             const unsigned *__restrict__ gen_list,
             UpdateCounters *__restrict__ ctr, Rec *__restrict__ rec,
             unsigned *__restrict__ chunk_fill, const unsigned cap_chunks, const PoseDev *__restrict__ pose)
{
  const unsigned n_work = __ldcg(&ctr->n_general);
  if (n_work == 0u) return;
  __shared__ int4 s_qa[MARCH_WARPS][QCAP], s_qb[MARCH_WARPS][QCAP];   // per-warp queue of surviving march steps
  int pos_mm[3];
#pragma unroll
  for (int a = 0; a < 3; a++) pos_mm[a] = pose ? pose->pos_mm[a] : P.pos_mm[a];
  __shared__ int4 s_ray[MARCH_WARPS][RAY_WORDS];                       // the warp's current ray
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int4 *qa_s = s_qa[wib], *qb_s = s_qb[wib];
  const unsigned total_warps = gridDim.x * MARCH_WARPS;

  MarchCtx cx;
  rec_init(cx.rw, ctr, lane);
  cx.n_cand = 0u;
  cx.err = 0u;

  unsigned k_cur = blockIdx.x * MARCH_WARPS + wib;
  while (k_cur < n_work)
  {
    if (lane < RAY_WORDS) s_ray[wib][lane] = reinterpret_cast<const int4 *>(&rays[gen_list[k_cur]])[lane];
    __syncwarp();
    const int4 *ray_s = s_ray[wib];
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
      march_ray<ATOMIC, false>(g, P, pos_mm, ray_s, start, end, lane, qa_s, qb_s, cx, rec, chunk_fill, cap_chunks, ctr);
    }
    __syncwarp();
    unsigned nxt = 0;
    if (lane == 0) nxt = atomicAdd(&ctr->gen_counter, 1u);
    k_cur = total_warps + __shfl_sync(FULL, nxt, 0);
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

// replay list writer: a warp reserves WS_LIST_SPAN entries at a time (one same-address atomic per span) and
// pads what it does not use with entries whose slot is LIST_NONE
#define WS_LIST_SPAN 128
#define LIST_NONE 0xFFFFFFFFFFFFFFFFull
struct ListWriter
{
  unsigned base, used;
};

WS_D void list_pad(ListWriter &w, const int lane, Rec *__restrict__ list, const unsigned cap)
{
  for (unsigned t = w.used + (unsigned)lane; t < WS_LIST_SPAN; t += 32u)
  {
    Rec e; e.key = 0ull; e.ref = LIST_NONE;
    if (w.base + t < cap) list[w.base + t] = e;
  }
  w.used = WS_LIST_SPAN;
}

WS_D void list_append(ListWriter &w, const bool want, const Rec e, const int lane, Rec *__restrict__ list, const unsigned cap,
                      UpdateCounters *__restrict__ ctr)
{
  const unsigned m = __ballot_sync(FULL, want);
  if (m == 0u) return;
  const unsigned n = (unsigned)__popc(m);
  if (w.used + n > WS_LIST_SPAN)
  {
    if (w.used < WS_LIST_SPAN) list_pad(w, lane, list, cap);
    unsigned b = 0;
    if (lane == 0)
    {
      b = atomicAdd(&ctr->n_list, (unsigned)WS_LIST_SPAN);
      if (b + WS_LIST_SPAN > cap) atomicOr(&ctr->error, 2u);
    }
    w.base = __shfl_sync(FULL, b, 0);
    w.used = 0;
  }
  const unsigned at = w.base + w.used + (unsigned)__popc(m & ((1u << lane) - 1u));
  if (want && at < cap) list[at] = e;
  w.used += n;
}

// ---- lockstep march --------------------------------------------------------------------------------
// One RAY per LANE: the 32 lanes of a warp walk 32 consecutive rays of the scan in lockstep, march step i of
// all of them in the same loop turn.  What depends on the step alone (len, delta_z, the shape of the
// interpolation fan, near/far) is warp-uniform; the DDA, the "same (x,y) column" filter
// (update_tsdf.cpp:455-458) and the brick flag are private to the lane -- no shuffles, no queue.  A work item
// is (group of 32 rays, block of LS_BLOCK march steps), handed out by a global counter, group-major.
//
// Two phases (two launches), split per ray at RaySetup::pad0:
//   SURF   the last ~2 * tau of every ray: the value of the marched voxel is computed (:466-472), every
//          candidate is a 64-bit atomicMin on the voxel's key, far-field candidates are recorded -- the
//          order-dependent machinery (merge, parked voxels, replay) works on these candidates alone;
//   FREE   everything before: value == tau by construction, so a candidate is just "a real / an interpolated
//          free-space candidate arrived" -- one bit OR-ed into the voxel's 4-bit state (L2 resident), no key,
//          no record, no DRAM traffic.  Free-space candidates lose against any candidate of the surface phase
//          (|value| < tau); the only thing the order of processing still decides is whether one of them
//          FOLLOWS the interpolated winner of a parked voxel, so this phase runs after the surface merge has
//          parked those voxels and offers such a candidate to the replay on the spot.
// Voxel addresses come from three per-axis tables in shared memory (ring wrap, residency, bricking).
// Rays outside the bounds of the 32-bit arithmetic (march_math.cuh) are left to march_kernel.
#ifndef LS_CTAS
#define LS_CTAS 3
#endif
#define TAB_INVALID 0xFFFFFFFFu
#define LS_NO_TAB 0xFFFFFFFFu      // march_lockstep_kernel: no second item table

WS_D u64 step_bits(int a, int b)   // bits [a, b) of a 64-bit mask, 0 <= a < b <= 64
{
  const u64 hi = b >= 64 ? ~0ull : ((1ull << b) - 1ull);
  return hi & ~((1ull << a) - 1ull);
}

struct LsShared        // per CTA: address tables; per warp: ray table and queue of surviving march steps
{
  const unsigned *tx, *ty, *tz;     // voxel - lo (0 .. size-1) -> address parts, see the table set-up in the kernel
  int4 *ray;                        // [32][2] per lane of the group: {hit point, distance}, {interpolation vector, -}
  unsigned queue_s;                 // shared-memory address of the warp's queue: [QCAP] {proj x, y, z, march step << 5 | lane}
  unsigned *pd;                     // [3][32] FREE: state words of the previous batch's candidates (cp.async)
};

struct LsOut           // where the free-space phase offers candidates that land on parked voxels
{
  u64 *pend_key;
  Rec *list;
  unsigned pending_cap, list_cap;
};

#define NO_PARK 0xFFFFFFFFu     // brick_slot_base of a brick without parked voxels

struct LsWarp          // a warp's running state across work items
{
  RecWriter rw;
  ListWriter lw;
  unsigned n_cand, err;
  unsigned last_brick;
  // FREE: the lane's candidates of the previous batch (fan steps 0..2): the voxel's state word was loaded when
  // the candidate was written and is looked at one batch later, when the load has long landed
  unsigned pd_mask;                  // which of the three slots hold a candidate
  unsigned pd_info[3];               // info: state shift [4:0], interpolated [5], fan step [11:6], march step [26:12], lane of the ray [31:27]
  u64 pd_addr[3];
};

// (n * r) / MATRIX_RESOLUTION (C++ truncation, :488 and :493) for n >= 0: the product has r's sign, so the quotient is
// the shifted magnitude with r's sign -- IMAD, SHF, IMAD (the last one folds into the addition that follows) instead of
// div_mr32's IMAD, SHF, LOP3, IADD, SHF: the ALU pipe (SHF / LOP3 at half rate) is what bounds the march.
// `ar` = |r|, `sr` = +-1; n * |r| < 2^31 on the fast path (march_math.cuh ray_is_small).
#ifndef WS_FAN_SIGN
#define WS_FAN_SIGN 1
#endif
#ifndef WS_NEAR_CONST
#define WS_NEAR_CONST 1
#endif
WS_D int mr_scaled(int n, int r, int ar, int sr)
{
#if WS_FAN_SIGN
  (void)r;
  return sr * (int)((unsigned)(n * ar) >> WS_MR_SHIFT);
#else
  (void)ar; (void)sr;
  return div_mr32(n * r);
#endif
}

// FREE phase, rare: a free-space candidate (|value| == tau) landed on a parked voxel.  What the replay's first
// round does for a recorded candidate: offer it to the voxel if it follows the parked winner in the
// reference's order, and keep it for the later rounds.  Warp-collective (list_append).
WS_D void offer_parked(const GridDesc &g, const UpdateParams &P, const LsOut &out, LsWarp &W,
                       const bool chk, const u64 addr, const unsigned info, const unsigned ray_base, const int lane,
                       UpdateCounters *__restrict__ ctr)
{
  unsigned slot = 0u;
  bool hit = false;
  u64 key = 0ull;
  if (chk)
  {
    const unsigned step = (info >> 6) & 63u, i = (info >> 12) & 0x7FFFu, ray = ray_base + (info >> 27);
    const u64 seq2 = make_seq(ray, i, step) << 1;
    key = (info & 32u) ? (((u64)(unsigned)P.tau << 47) | (1ull << 46) | (((1ull << 46) - 2ull) - seq2))
                       : (((u64)(unsigned)P.tau << 47) | seq2);
    const u64 kv = __ldcg(&g.keys[addr]);
    if (key_is_pending(kv))
    {
      slot = __ldcg(&g.brick_slot_base[addr >> 9]) + (unsigned)((kv >> WS_SEQ_BITS) & 0x1FFull);
      hit = slot < out.pending_cap && key_seq(key) > (kv & WS_SEQ_MAX);
      if (hit) atomicMin(&out.pend_key[slot], key);
    }
  }
  Rec e; e.key = key; e.ref = (u64)slot;
  list_append(W.lw, hit, e, lane, out.list, out.list_cap, ctr);
}

// the candidates of the previous batch: did any land on a parked voxel?
WS_D void free_check_pending(const GridDesc &g, const UpdateParams &P, const LsShared &sh, const LsOut &out, LsWarp &W,
                             const unsigned ray_base, const int lane, UpdateCounters *__restrict__ ctr)
{
  asm volatile("cp.async.wait_all;" ::: "memory");      // issued a batch ago: long landed
#pragma unroll
  for (int sl = 0; sl < 3; sl++)
  {
    const bool chk = ((W.pd_mask >> sl) & 1u) && ((sh.pd[sl * 32 + lane] >> (W.pd_info[sl] & 31u)) & VS_PARKED) != 0u;
    if (__any_sync(FULL, chk)) offer_parked(g, P, out, W, chk, W.pd_addr[sl], W.pd_info[sl], ray_base, lane, ctr);
  }
  W.pd_mask = 0u;
}

// FREE phase, one fan step of a batch: OR the candidate's bit into the voxel's state (no return value, the
// loop never waits for memory); the state word is loaded beside it for the check one batch later
template <bool WIDE, int SLOT>
WS_D void free_candidate(const GridDesc &g, const UpdateParams &P, const LsShared &sh, const LsOut &out, LsWarp &W,
                         const bool have, const int step, const int iter_steps, const int mid, const int low[3],
                         const int riv[3], const int ariv[3], const int sriv[3], const int i, const int rl,
                         const unsigned ray_base, const int lane, UpdateCounters *__restrict__ ctr)
{
  const int nlo[3] = { -P.lo[0], -P.lo[1], -P.lo[2] };
  int fx = 0, fy = 0, fz = 0;
  if (step > 0)
  {
    const int sr = wmul(step, P.res);
    fx = mr_scaled(sr, riv[0], ariv[0], sriv[0]); fy = mr_scaled(sr, riv[1], ariv[1], sriv[1]); fz = mr_scaled(sr, riv[2], ariv[2], sriv[2]);
  }
  const unsigned tx = (unsigned)(fd32_sdiv_s(low[0] + fx, P.div_res32) + nlo[0]);          // :493
  const unsigned ty = (unsigned)(fd32_sdiv_s(low[1] + fy, P.div_res32) + nlo[1]);
  const unsigned tz = (unsigned)(fd32_sdiv_s(low[2] + fz, P.div_res32) + nlo[2]);
  bool valid = have && step < iter_steps &&
               !(tx > (unsigned)P.ext[0] || ty > (unsigned)P.ext[1] || tz > (unsigned)P.ext[2]);       // :495-498
  unsigned ax = TAB_INVALID;
  if (valid) ax = sh.tx[tx];
  valid = ax != TAB_INVALID;                       // else: out of bounds, or the column lives on another rank
  u64 addr = 0ull;
  unsigned info = 0u;
  bool chk = false;
  if (valid)
  {
    W.n_cand++;
    unsigned brick;
    if (WIDE) { addr = ((u64)ax << 6) + (u64)(sh.ty[ty] + sh.tz[tz]); brick = (unsigned)(addr >> 9); }
    else { const unsigned a32 = ax + sh.ty[ty] + sh.tz[tz]; addr = (u64)a32; brick = a32 >> 9; }
    if (brick != W.last_brick) { g.brick_flag2[brick] = 1u; W.last_brick = brick; }
    const unsigned shft = vstate_shift(addr);
    unsigned *word = g.vstate + vstate_word(addr);
    if (SLOT >= 0) info = shft | ((unsigned)(step == mid ? 0 : 1) << 5) | ((unsigned)step << 6) | ((unsigned)i << 12) | ((unsigned)rl << 27);
    if (SLOT >= 0 && SLOT < 3)
    {
      // the state word travels to shared memory by cp.async: no register, no scoreboard to wait on
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&sh.pd[(SLOT >= 0 && SLOT < 3 ? SLOT : 0) * 32 + lane])),
                   "l"(word) : "memory");
      W.pd_addr[SLOT >= 0 && SLOT < 3 ? SLOT : 0] = addr;
      W.pd_info[SLOT >= 0 && SLOT < 3 ? SLOT : 0] = info;
      W.pd_mask |= 1u << (SLOT >= 0 ? SLOT : 0);
    }
    else if (SLOT >= 3) chk = ((__ldcg(word) >> shft) & VS_PARKED) != 0u;     // long fans (delta_z >= res): checked on the spot
    // a plain byte store: idempotent, so no atomic and nothing to serialise when many rays hit one voxel
    g.ffree[2 * addr - (addr & 511ull) + (step == mid ? 0u : 512u)] = 1;     // brick * 1024 + local (+ 512: interpolated)
  }
  if (SLOT >= 3 && __any_sync(FULL, chk)) offer_parked(g, P, out, W, chk, addr, info, ray_base, lane, ctr);   // warp-collective
}

// The heavy part: `take` queued march steps, one per lane, of any of the group's rays.
// WIDE: voxel addresses need more than 32 bits (the x table then holds address / 64)
template <bool SURF, bool ATOMIC, bool WIDE>
WS_D void march_drain(const GridDesc &g, const UpdateParams &P, const LsShared &sh, const int qh, const int take,
                      const unsigned ray_base, const int lane, LsWarp &W, Rec *__restrict__ rec,
                      unsigned *__restrict__ chunk_fill, const unsigned cap_chunks, UpdateCounters *__restrict__ ctr,
                      const LsOut &out)
{
  bool have = lane < take;
  int4 q;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(sh.queue_s + (unsigned)(((qh + lane) & (QCAP - 1)) << 4)) : "memory");
  const int rl = q.w & 31, i = q.w >> 5;
  const int4 r1 = sh.ray[2 * rl + 1];
  const int riv[3] = { r1.x, r1.y, r1.z };
  const int ariv[3] = { r1.x < 0 ? -r1.x : r1.x, r1.y < 0 ? -r1.y : r1.y, r1.z < 0 ? -r1.z : r1.z };
  const int sriv[3] = { r1.x < 0 ? -1 : 1, r1.y < 0 ? -1 : 1, r1.z < 0 ? -1 : 1 };
  const int nlo[3] = { -P.lo[0], -P.lo[1], -P.lo[2] };
#define LS_VOX(x, a) ((unsigned)(fd32_sdiv_s((x), P.div_res32) + nlo[a]))
  const int len = 1 + i * P.half_res;

  u64 key_real0 = 0ull, key_int0 = 0ull;
  if (SURF)
  {
    // ---- value of the marched voxel (:466-472) -------------------------------------------------
    const int4 r0 = sh.ray[2 * rl];
    const int mx = fd32_sdiv_s(q.x, P.div_res32), my = fd32_sdiv_s(q.y, P.div_res32), mz = fd32_sdiv_s(q.z, P.div_res32);
    const int ex = wsub(r0.x, wadd(wmul(mx, P.res), P.half_res));
    const int ey = wsub(r0.y, wadd(wmul(my, P.res), P.half_res));
    const int ez = wsub(r0.z, wadd(wmul(mz, P.res), P.half_res));
    const int vsq = wadd(wadd(wmul(ex, ex), wmul(ey, ey)), wmul(ez, ez));
    int value = P.tau;
    if (__any_sync(FULL, have && (unsigned)vsq < (unsigned)P.tau_sq))
    {
      const int root = isqrt31(vsq < 0 ? 0 : vsq);
      if ((unsigned)vsq < (unsigned)P.tau_sq) value = root < P.tau ? root : P.tau;
    }
    if (len > r0.w) value = -value;
    if (value <= P.zero_weight_max) have = false;        // weight == 0 (:475-483)
    const unsigned ray = ray_base + (unsigned)rl;
    const unsigned av = (unsigned)(value < 0 ? -value : value);
    const unsigned klo = (ray << 22) | ((unsigned)i << (WS_SEQ_STEP_BITS + 1)) | (value < 0 ? 1u : 0u);
    key_real0 = ((u64)((av << 15) | (ray >> 10)) << 32) | (u64)klo;
    // interpolated: order field = SEQ_MAX - seq, i.e. (2^46 - 2) - (seq << 1) in place
    key_int0 = ((u64)av << 47) | (1ull << 46) | (((1ull << 46) - 2ull) - (((u64)(ray >> 10) << 32) | (u64)(klo & ~1u))) | (u64)(klo & 1u);
  }

  // ---- the fan of interpolated voxels around the step (:485-506) -----------------------------------
#if WS_FAN_SIGN
  const int delta_z = (int)((unsigned)(P.dz_per_distance * len) >> WS_MR_SHIFT);             // :485 (both factors > 0, product < 2^30)
#else
  const int delta_z = div_mr32(wmul(P.dz_per_distance, len));                                 // :485
#endif
  const bool far = len >= P.far_len;
#if WS_NEAR_CONST
  // near field (before far_start_len): delta_z < res / 2 there by the definition of far_len, so the fan is the marched
  // voxel's one real candidate (iter_steps == 1, mid == 0) -- no divisions (IMAD.HI issues at a quarter of the rate),
  // no warp reduction; no voxel there can be parked either: the candidate is stored and forgotten
  if (!SURF && !__any_sync(FULL, have && far))
  {
    const int low1[3] = { q.x - mr_scaled(delta_z, riv[0], ariv[0], sriv[0]), q.y - mr_scaled(delta_z, riv[1], ariv[1], sriv[1]),
                          q.z - mr_scaled(delta_z, riv[2], ariv[2], sriv[2]) };                                                // :488
    free_candidate<WIDE, -1>(g, P, sh, out, W, have, 0, 1, 0, low1, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
    return;
  }
#endif
  const int iter_steps = (int)fd32_udiv((unsigned)(delta_z * 2), P.div_res32) + 1;           // :486
  const int mid = (int)fd32_udiv((unsigned)delta_z, P.div_res32);                            // :487
  if (have && iter_steps > (1 << WS_SEQ_STEP_BITS)) W.err |= 1u;
  const int low[3] = { q.x - mr_scaled(delta_z, riv[0], ariv[0], sriv[0]), q.y - mr_scaled(delta_z, riv[1], ariv[1], sriv[1]),
                       q.z - mr_scaled(delta_z, riv[2], ariv[2], sriv[2]) };                                                   // :488
  const int max_steps = __reduce_max_sync(FULL, have ? iter_steps : 0);

  if (!SURF)
  {
    // near field (before far_start_len): no voxel there can be parked, the candidates are stored and forgotten
    if (!__any_sync(FULL, have && far))
    {
      free_candidate<WIDE, -1>(g, P, sh, out, W, have, 0, iter_steps, mid, low, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
#pragma unroll 1
      for (int step = 1; step < max_steps; ++step)
        free_candidate<WIDE, -1>(g, P, sh, out, W, have, step, iter_steps, mid, low, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
      return;
    }
    free_check_pending(g, P, sh, out, W, ray_base, lane, ctr);
    free_candidate<WIDE, 0>(g, P, sh, out, W, have, 0, iter_steps, mid, low, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
    if (max_steps > 1) free_candidate<WIDE, 1>(g, P, sh, out, W, have, 1, iter_steps, mid, low, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
    if (max_steps > 2) free_candidate<WIDE, 2>(g, P, sh, out, W, have, 2, iter_steps, mid, low, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
#pragma unroll 1
    for (int step = 3; step < max_steps; ++step)
      free_candidate<WIDE, 3>(g, P, sh, out, W, have, step, iter_steps, mid, low, riv, ariv, sriv, i, rl, ray_base, lane, ctr);
    return;
  }

#pragma unroll 1
  for (int step = 0; step < max_steps; ++step)                                               // :491
  {
    int fx = 0, fy = 0, fz = 0;
    if (step > 0)
    {
      const int sr = wmul(step, P.res);
      fx = mr_scaled(sr, riv[0], ariv[0], sriv[0]); fy = mr_scaled(sr, riv[1], ariv[1], sriv[1]); fz = mr_scaled(sr, riv[2], ariv[2], sriv[2]);
    }
    const unsigned tx = LS_VOX(low[0] + fx, 0), ty = LS_VOX(low[1] + fy, 1), tz = LS_VOX(low[2] + fz, 2);   // :493
    bool valid = have && step < iter_steps &&
                 !(tx > (unsigned)P.ext[0] || ty > (unsigned)P.ext[1] || tz > (unsigned)P.ext[2]);       // :495-498
    unsigned ax = TAB_INVALID;
    if (valid) ax = sh.tx[tx];
    valid = ax != TAB_INVALID;                       // else: out of bounds, or the column lives on another rank
    u64 key = 0ull, addr = 0ull;
    if (valid)
    {
      W.n_cand++;
      unsigned brick;
      if (WIDE) { addr = ((u64)ax << 6) + (u64)(sh.ty[ty] + sh.tz[tz]); brick = (unsigned)(addr >> 9); }
      else { const unsigned a32 = ax + sh.ty[ty] + sh.tz[tz]; addr = (u64)a32; brick = a32 >> 9; }
      const u64 s2 = (u64)(unsigned)step << 1;
      key = step == mid ? key_real0 + s2 : key_int0 - s2;                                    // :503-506
      if (ATOMIC)
      {
        atomicMin(&g.keys[addr], key);                                                       // :508-512
        if (brick != W.last_brick) { g.brick_flag[brick] = 1u; W.last_brick = brick; }
      }
    }
    rec_append(W.rw, valid && far, key, addr, WIDE, lane, rec, chunk_fill, cap_chunks, ctr);     // warp-collective
  }
#undef LS_VOX
}

// One work item: LS_BLOCK march steps of a group of 32 rays.  Step phase in lockstep (one ray per lane),
// survivors of the column filter queued, heavy part on full batches of 32.
template <bool SURF, bool ATOMIC, bool WIDE>
WS_D void march_block(const GridDesc &g, const UpdateParams &P, const int pos_mm[3], const LsShared &sh,
                      const RaySetup *__restrict__ rays, const unsigned ray_base, const int s0, const int lane, LsWarp &W,
                      Rec *__restrict__ rec, unsigned *__restrict__ chunk_fill, const unsigned cap_chunks,
                      UpdateCounters *__restrict__ ctr, const LsOut &out)
{
  // ---- this lane's ray ------------------------------------------------------------------------
  const unsigned ray = ray_base + (unsigned)lane;
  const bool in_scan = ray < (unsigned)P.n_points;
  const int4 *rw4 = reinterpret_cast<const int4 *>(rays + (in_scan ? ray : 0u));
  const int4 w2 = __ldg(rw4 + 2), w3 = __ldg(rw4 + 3);
  const bool ok = in_scan && (w2.w & 1);            // fast-path rays only
  u64 mask = 0ull;                                  // march steps of this block the lane has to process
  if (ok)
  {
    const int n_segs = P.n_xiv > 0 ? RAY_SEGS : 1;
    const int split = w3.w;
    int lo_lim = SURF ? split : 0;
    const int hi_lim = SURF ? 0x7fffffff : split;
    if (!ATOMIC && P.far_len > 1) { const int fs = (P.far_len - 1) / P.half_res; lo_lim = lo_lim > fs ? lo_lim : fs; }
    lo_lim = lo_lim > s0 ? lo_lim : s0;
    const int hi_blk = hi_lim < s0 + LS_BLOCK ? hi_lim : s0 + LS_BLOCK;
#pragma unroll 1
    for (int sg = 0; sg < n_segs; sg += 2)
    {
      const int4 ws = __ldg(rw4 + 5 + (sg >> 1));
      if (ws.x >= ws.y) break;
      int a = (ws.x > lo_lim ? ws.x : lo_lim) - s0, b = (ws.y < hi_blk ? ws.y : hi_blk) - s0;
      if (a < b) mask |= step_bits(a, b);
      if (n_segs == 1 || ws.z >= ws.w) break;
      a = (ws.z > lo_lim ? ws.z : lo_lim) - s0; b = (ws.w < hi_blk ? ws.w : hi_blk) - s0;
      if (a < b) mask |= step_bits(a, b);
    }
  }
  const int f = __reduce_min_sync(FULL, mask ? __ffsll((long long)mask) - 1 : LS_BLOCK);
  const int l = __reduce_max_sync(FULL, mask ? 63 - __clzll((long long)mask) : -1);
  if (l < 0) return;
  const int i0 = s0 + f, i1 = s0 + l;

  const int4 w0 = __ldg(rw4 + 0), w1 = __ldg(rw4 + 1), w4 = __ldg(rw4 + 4);
  const int distance = ok ? w0.w : 1;
  if (ok && w1.w > (1 << WS_SEQ_MARCH_BITS)) W.err |= 1u;
  __syncwarp();                                      // the previous item's heavy part is done with the tables
  sh.ray[2 * lane] = w0;
  sh.ray[2 * lane + 1] = w2;
  __syncwarp();

  // ---- DDA (march_math.cuh), one step per turn: proj = pos + sgn * floor(|d| * len / distance) ----------
  int proj[3], sdq[3], sgn[3];
  unsigned rem[3], drem[3];
  {
    const int ii = i0 > 0 ? i0 - 1 : 0;
    const unsigned len_i = 1u + (unsigned)ii * (unsigned)P.half_res;
    const double rdist = 1.0 / (double)distance;
    const int dv[3] = { w1.x, w1.y, w1.z };
    const unsigned dq1[3] = { (unsigned)w3.x, (unsigned)w3.y, (unsigned)w3.z };
    drem[0] = (unsigned)w4.x; drem[1] = (unsigned)w4.y; drem[2] = (unsigned)w4.z;
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
      const unsigned ad = dv[a] < 0 ? 0u - (unsigned)dv[a] : (unsigned)dv[a];
      unsigned q;
      divrem_rcp(ad * len_i, (unsigned)distance, rdist, q, rem[a]);
      sgn[a] = dv[a] < 0 ? -1 : 1;
      proj[a] = pos_mm[a] + sgn[a] * (int)q;
      sdq[a] = sgn[a] * (int)dq1[a];
    }
  }
  // Measured issue rates on B200 (tools/ubench/pipes.cu, warp instructions per clock and scheduler): IADD3 0.98, LOP3 / SHF /
  // ISETP / SEL / IMAD 0.49, IMAD.HI 0.25, POPC 0.12.  The compiler turns the `if` of the advance into two SELs per axis
  // (ALU pipe, which bounds the step phase); written with predicated adds it is one compare and four full-rate adds.
  // (The carry as a multiplication -- IMAD.HI with floor(2^32 / distance) + 1 -- was slower: 804 against 818 scans/s.)
#ifndef WS_STEP_PTX
#define WS_STEP_PTX 1
#endif

#if WS_STEP_PTX
#define LS_ADVANCE()                                                                   \
  _Pragma("unroll") for (int a = 0; a < 3; a++)                                        \
    asm("{\n\t.reg .pred p;\n\t"                                                      \
        "add.u32 %0, %0, %2;\n\t"                                                      \
        "add.s32 %1, %1, %4;\n\t"                                                      \
        "setp.ge.u32 p, %0, %3;\n\t"                                                   \
        "@p sub.u32 %0, %0, %3;\n\t"                                                   \
        "@p add.s32 %1, %1, %5;\n\t}"                                                  \
        : "+r"(rem[a]), "+r"(proj[a]) : "r"(drem[a]), "r"((unsigned)distance), "r"(sdq[a]), "r"(sgn[a]));
#else
#define LS_ADVANCE()                                                                   \
  _Pragma("unroll") for (int a = 0; a < 3; a++)                                        \
  {                                                                                    \
    proj[a] += sdq[a];                                                                 \
    rem[a] += drem[a];                                                                 \
    if (rem[a] >= (unsigned)distance) { rem[a] -= (unsigned)distance; proj[a] += sgn[a]; } \
  }
#endif
  // voxel - lo per axis (the in-bounds test is then one unsigned compare, :460-463).  (ptxas re-loads the loop's seven
  // constants from the parameter bank every turn -- six LDC; values made opaque to it are re-derived every turn instead)
  const int nlo[3] = { -P.lo[0], -P.lo[1], -P.lo[2] };
  const unsigned ext[3] = { (unsigned)P.ext[0], (unsigned)P.ext[1], (unsigned)P.ext[2] };
  const FastDiv32 dres = P.div_res32;
#define LS_VOX(x, a) ((unsigned)(fd32_sdiv_s((x), dres) + nlo[a]))
  unsigned px = 0x80000000u, py = 0x80000000u;       // no previous step (update_tsdf.cpp:448)
  if (i0 > 0)
  {
    px = LS_VOX(proj[0], 0);
    py = LS_VOX(proj[1], 1);
    LS_ADVANCE();
  }

  const unsigned lt = (1u << lane) - 1u;
  int qn = 0, qh = 0;                                // queue fill / head (warp-uniform)
  for (int i = i0; i <= i1; ++i)
  {
    const int cx = proj[0], cy = proj[1], cz = proj[2];                                         // :452
    LS_ADVANCE();
    const unsigned ix = LS_VOX(cx, 0), iy = LS_VOX(cy, 1), iz = LS_VOX(cz, 2);                  // :453 (- lo)
#if WS_STEP_PTX
    // the step's mask bit, the column filter (:455-458) and the bounds test (:460-463) as one chain of predicates and
    // the ballot on it (the compiler's version keeps `act` in a register: five SELs and a PRMT on the ALU pipe)
    unsigned m, act_r;
    {
      const unsigned mbit = (unsigned)(mask >> (i - s0));
      asm volatile("{\n\t.reg .pred p;\n\t"
                   "setp.ne.u32 p, %2, %3;\n\t"
                   "setp.ne.or.u32 p, %4, %5, p;\n\t"
                   "setp.le.and.u32 p, %2, %7, p;\n\t"
                   "setp.le.and.u32 p, %4, %8, p;\n\t"
                   "setp.le.and.u32 p, %6, %9, p;\n\t"
                   "and.b32 %1, %10, 1;\n\t"
                   "setp.ne.and.u32 p, %1, 0, p;\n\t"
                   "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                   "selp.u32 %1, 1, 0, p;\n\t}"
                   : "=r"(m), "=r"(act_r)
                   : "r"(ix), "r"(px), "r"(iy), "r"(py), "r"(iz), "r"(ext[0]), "r"(ext[1]), "r"(ext[2]), "r"(mbit));
    }
    px = ix; py = iy;
    if (m == 0u) continue;
    const bool act = act_r != 0u;
#else
    bool act = ((mask >> (i - s0)) & 1ull) != 0ull;
    if (ix == px && iy == py) act = false;                                                      // :455-458
    px = ix; py = iy;
    if (ix > ext[0] || iy > ext[1] || iz > ext[2]) act = false;                                   // :460-463
    const unsigned m = __ballot_sync(FULL, act);
    if (m == 0u) continue;
#endif
    if (act)
      asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};"
                   ::"r"(sh.queue_s + (unsigned)(((qh + qn + __popc(m & lt)) & (QCAP - 1)) << 4)), "r"(cx), "r"(cy), "r"(cz), "r"((i << 5) | lane) : "memory");
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32)
    {
      march_drain<SURF, ATOMIC, WIDE>(g, P, sh, qh, 32, ray_base, lane, W, rec, chunk_fill, cap_chunks, ctr, out);
      qh = (qh + 32) & (QCAP - 1);
      qn -= 32;
      __syncwarp();
    }
  }
  if (qn > 0) march_drain<SURF, ATOMIC, WIDE>(g, P, sh, qh, qn, ray_base, lane, W, rec, chunk_fill, cap_chunks, ctr, out);
  if (!SURF) free_check_pending(g, P, sh, out, W, ray_base, lane, ctr);     // before the ray table changes
#undef LS_ADVANCE
#undef LS_VOX
}

// (CTAs per SM of the free-space instantiations: 4 fits -- 64 registers, 20 bytes of spills -- and is slower, 794
// against 822 scans/s in a same-box A/B; so is 4 for the surface phase: the march is bound by the ALU pipe, not by
// latency that more resident warps could hide)
#ifndef LS_CTAS_FREE
#define LS_CTAS_FREE LS_CTAS
#endif
template <bool SURF, bool ATOMIC, bool WIDE>
__global__ void __launch_bounds__(MARCH_THREADS, SURF ? LS_CTAS : LS_CTAS_FREE)
march_lockstep_kernel(const GridDesc g, const UpdateParams P, const RaySetup *__restrict__ rays,
                      const uint2 *__restrict__ grp_info, const unsigned *__restrict__ item_off, const unsigned n_groups,
                      UpdateCounters *__restrict__ ctr, Rec *__restrict__ rec, unsigned *__restrict__ chunk_fill,
                      const unsigned cap_chunks, const PoseDev *__restrict__ pose, const LsOut out, const unsigned tab,
                      const unsigned tab_b, const unsigned warp_offset, const unsigned total_warps)
{
  extern __shared__ unsigned s_tab[];               // size[0] + size[1] + size[2] address parts
  __shared__ int4 s_ray[MARCH_WARPS][64];
  __shared__ __align__(16) int4 s_queue[MARCH_WARPS][QCAP];
  __shared__ unsigned s_pd[MARCH_WARPS][96];
  // item tables: 0 = surface phase; free-space phase: 1 = near field (needs nothing from the surface phase, runs beside
  // it on a second stream), 2 = far field (looks for parked voxels: launched after the surface merge)
  // ... 3 = the second part of the near field, marched by the far-field launch after its own items (tab_b)
  if (__ldcg(&ctr->n_items[tab]) == 0u && (tab_b == LS_NO_TAB || __ldcg(&ctr->n_items[tab_b]) == 0u)) return;
  if (tab == 2u && __ldcg(&ctr->rec_overflow) != 0u && __ldcg(&ctr->pending_overflow) == 0u) return;   // redone after the regrow
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  LsShared sh;
  sh.tx = s_tab; sh.ty = s_tab + g.size[0]; sh.tz = sh.ty + g.size[1];
  sh.ray = s_ray[wib]; sh.pd = s_pd[wib];
  sh.queue_s = (unsigned)__cvta_generic_to_shared(s_queue[wib]);
  asm volatile("" : "+r"(sh.queue_s));              // opaque: kept in a register instead of being re-derived every turn
  for (int t = threadIdx.x; t < g.size[0] + g.size[1] + g.size[2]; t += blockDim.x)
  {
    // ring coordinate (hdf5_local_map.h:140-151) of voxel lo + t, then its share of the bricked address:
    // address = x part (<< 6 if WIDE) + y part + z part
    unsigned v;
    if (t < g.size[0])
    {
      int r = t + P.ringc[0]; r -= r >= g.size[0] ? g.size[0] : 0;
      const int slot = g.full ? (r >> 3) : (int)g.xslot[r >> 3];
      v = slot < 0 ? TAB_INVALID : (unsigned)slot * (unsigned)(g.nb[1] * g.nb[2]) * 8u + (unsigned)(r & 7);
      if (!WIDE && slot >= 0) v <<= 6;
    }
    else if (t < g.size[0] + g.size[1])
    {
      int r = t - g.size[0] + P.ringc[1]; r -= r >= g.size[1] ? g.size[1] : 0;
      v = (unsigned)(r >> 3) * (unsigned)g.nb[2] * WS_BRICK_VOX + (unsigned)((r & 7) << 3);
    }
    else
    {
      int r = t - g.size[0] - g.size[1] + P.ringc[2]; r -= r >= g.size[2] ? g.size[2] : 0;
      v = (unsigned)(r >> 3) * WS_BRICK_VOX + (unsigned)(r & 7);
    }
    s_tab[t] = v;
  }
  __syncthreads();

  int pos_mm[3];
#pragma unroll
  for (int a = 0; a < 3; a++) pos_mm[a] = pose ? pose->pos_mm[a] : P.pos_mm[a];
  // (warp_offset / total_warps: a phase may be marched by two launches that share the item counters -- the far field:
  // two CTAs per SM at once, a third one when the replay's record pass has left the SMs)

  LsWarp W;
  if (SURF) rec_init(W.rw, ctr, lane);
  W.lw.base = 0u; W.lw.used = WS_LIST_SPAN;
  W.n_cand = 0u; W.err = 0u; W.last_brick = 0xFFFFFFFFu;
  W.pd_mask = 0u;
#pragma unroll
  for (int sl = 0; sl < 3; sl++) { W.pd_info[sl] = 0u; W.pd_addr[sl] = 0ull; }

  for (int pass = 0; pass < 2; pass++)
  {
  const unsigned tb = pass == 0 ? tab : tab_b;
  if (tb == LS_NO_TAB) break;
  const unsigned n_items = __ldcg(&ctr->n_items[tb]);
  if (n_items == 0u) continue;
  unsigned *counter = &ctr->item_counter[tb];
  const uint2 *ginfo = grp_info + (size_t)tb * n_groups;
  const unsigned *ioff = item_off + (size_t)tb * (n_groups + 1u);
  // items: the first one by warp index, the others from a global counter, fetched one item ahead
  unsigned k_cur = warp_offset + blockIdx.x * MARCH_WARPS + wib;
  unsigned k_nxt = 0;
  if (lane == 0) k_nxt = atomicAdd(counter, 1u);
  k_nxt = total_warps + __shfl_sync(FULL, k_nxt, 0);
  while (k_cur < n_items)
  {
    unsigned fetched = 0;
    if (lane == 0) fetched = atom_add_async(counter, 1u);
    // item -> (group, block): the last group whose first item is <= k_cur
    unsigned lo = 0u, hi = n_groups;
    while (hi - lo > 1u)
    {
      const unsigned m = (lo + hi) >> 1;
      if (__ldg(&ioff[m]) <= k_cur) lo = m; else hi = m;
    }
    const uint2 gi = __ldg(&ginfo[lo]);
    const int blk = (int)(gi.x + (k_cur - __ldg(&ioff[lo])));
    march_block<SURF, ATOMIC, WIDE>(g, P, pos_mm, sh, rays, lo * 32u, blk * LS_BLOCK, lane, W, rec, chunk_fill, cap_chunks, ctr, out);
    asm volatile("" : "+r"(fetched) : "r"(W.n_cand), "r"(W.lw.used));   // keep the fetch in flight behind the march
    k_cur = k_nxt;
    k_nxt = total_warps + __shfl_sync(FULL, fetched, 0);
  }
  }

  if (SURF) rec_finish(W.rw, lane, chunk_fill, cap_chunks, ctr);
  else if (W.lw.used < WS_LIST_SPAN) list_pad(W.lw, lane, out.list, out.list_cap);
  if (ATOMIC)
  {
    unsigned long long nc = W.n_cand;
    for (int o = 16; o > 0; o >>= 1) nc += __shfl_down_sync(FULL, nc, o);
    if (lane == 0 && nc) atomicAdd(&ctr->n_candidates, nc);
  }
  if (W.err) atomicOr(&ctr->error, W.err);
}

// exclusive prefix sums of the groups' block counts, per phase -> item_off[phase][0 .. n_groups], n_items[phase]
// (one CTA per table)
__global__ void __launch_bounds__(1024)
item_scan_kernel(const uint2 *__restrict__ grp_info, const unsigned n_groups, unsigned *__restrict__ item_off,
                 UpdateCounters *__restrict__ ctr)
{
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const int phase = blockIdx.x;          // one CTA per item table
    const uint2 *gi = grp_info + (size_t)phase * n_groups;
    unsigned *io = item_off + (size_t)phase * (n_groups + 1u);
    __syncthreads();
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    for (unsigned base = 0; base < n_groups; base += 1024u)
    {
      const unsigned gidx = base + (unsigned)tid;
      const unsigned v = gidx < n_groups ? gi[gidx].y : 0u;
      unsigned x = v;
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) s_warp[warp] = x;
      __syncthreads();
      if (warp == 0)
      {
        unsigned w = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1)
        {
          const unsigned y = __shfl_up_sync(FULL, w, o);
          if (lane >= o) w += y;
        }
        s_warp[lane] = w;
      }
      __syncthreads();
      const unsigned incl = s_carry + x + (warp > 0 ? s_warp[warp - 1] : 0u);
      if (gidx < n_groups) io[gidx] = incl - v;
      __syncthreads();
      if (tid == 1023) s_carry = incl;
      __syncthreads();
    }
    if (tid == 0) { io[n_groups] = s_carry; ctr->n_items[phase] = s_carry; }
  }
}

// final winner -> grid entry (update_tsdf.cpp:542-560); returns 1 if the reference writes the entry
WS_D unsigned apply_winner_to(uint32_t *slot, const uint32_t e, const UpdateParams &P, const u64 key)
{
  const int value = key_value(key);
  int weight = tsdf_weight_fd(value, P.tau, P.weight_epsilon, P.div_weps);
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
      "WAIT_LOOP_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE_%=;\n"
      "bra.uni WAIT_LOOP_%=;\n"
      "WAIT_DONE_%=:\n"
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

// The bricks a merge CTA folds: every gridDim.x-th brick of the map whose touched flag is set (the marches set the
// flags with plain stores), compacted into shared memory, BRICK_SHARE_CAP map bricks per round; the flags are cleared
// for the next scan.  This used to be a kernel of its own (flags -> global list, 8.5 us plus a launch gap in front of
// either merge); the strided assignment spreads the touched bricks evenly (they cluster in space).
#ifndef BRICK_SHARE_CAP
#define BRICK_SHARE_CAP 512
#endif
template <bool SURF>
WS_D unsigned collect_bricks(const GridDesc &g, const unsigned first, unsigned *s_list, unsigned *s_n)
{
  unsigned *flag = SURF ? g.brick_flag : g.brick_flag2;
  const int tid = threadIdx.x, lane = tid & 31;
  __syncthreads();                                  // everybody is done with the previous round's list
  if (tid == 0) *s_n = 0u;
  __syncthreads();
  for (unsigned k = (unsigned)tid; k < BRICK_SHARE_CAP; k += 256u)
  {
    const i64 b = (i64)blockIdx.x + (i64)(first + k) * (i64)gridDim.x;
    const bool set = b < (i64)g.n_bricks && flag[b] != 0u;
    const unsigned m = __ballot_sync(FULL, set);
    if (m == 0u) continue;
    unsigned at = 0u;
    if (lane == 0) at = atomicAdd(s_n, (unsigned)__popc(m));
    at = __shfl_sync(FULL, at, 0);
    if (set)
    {
      s_list[at + (unsigned)__popc(m & ((1u << lane) - 1u))] = (unsigned)b;
      flag[b] = 0u;
      if (SURF) g.brick_flag2[b] = 1u;              // a surface brick is a touched brick
    }
  }
  __syncthreads();
  return *s_n;
}

__global__ void __launch_bounds__(256, MERGE_CTAS)
merge_kernel(const GridDesc g, const UpdateParams P,
                 UpdateCounters *__restrict__ ctr, const unsigned pending_cap,
                 u64 *__restrict__ pend_addr, u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key)
{
  __shared__ __align__(128) u64 s_keys[MERGE_STAGES][WS_BRICK_VOX];
  __shared__ __align__(128) uint32_t s_ent[MERGE_STAGES][WS_BRICK_VOX];
  __shared__ __align__(8) u64 s_bar[MERGE_STAGES];
  __shared__ unsigned s_cnt[8][2];
  __shared__ unsigned s_base;
  __shared__ __align__(128) unsigned char s_state[MERGE_STAGES][WS_BRICK_VOX / 2];
  __shared__ unsigned s_list[BRICK_SHARE_CAP];
  __shared__ unsigned s_n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned touched = 0, written = 0;

  if (tid == 0)
  {
    for (int st = 0; st < MERGE_STAGES; st++) mbar_init(&s_bar[st], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // bricks of this CTA: the flagged ones among blockIdx.x, + gridDim.x, ... (collect_bricks, BRICK_SHARE_CAP map
  // bricks per round); the it-th brick the CTA folds uses stage it % MERGE_STAGES, the mbarrier phases run on across rounds
  const unsigned share = (i64)blockIdx.x < (i64)g.n_bricks ? (unsigned)(((i64)g.n_bricks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  unsigned it0 = 0u;
  for (unsigned r0 = 0u; r0 < share; r0 += BRICK_SHARE_CAP)
  {
  const unsigned n_mine = collect_bricks<true>(g, r0, s_list, &s_n);
  if (tid == 0)
  {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    for (unsigned n = 0; n < MERGE_STAGES - 1 && n < n_mine; n++)
    {
      const int sn = (int)((it0 + n) % MERGE_STAGES);
      const size_t base = (size_t)s_list[n] * WS_BRICK_VOX;
      mbar_expect_tx(&s_bar[sn], MERGE_KEY_BYTES + MERGE_ENT_BYTES);
      bulk_g2s(s_keys[sn], g.keys + base, MERGE_KEY_BYTES, &s_bar[sn]);
      bulk_g2s(s_ent[sn], g.grid + base, MERGE_ENT_BYTES, &s_bar[sn]);
    }
  }

  for (unsigned n = 0; n < n_mine; n++)
  {
    const int st = (int)((it0 + n) % MERGE_STAGES);
    const unsigned parity = ((it0 + n) / MERGE_STAGES) & 1u;
    const size_t base = (size_t)s_list[n] * WS_BRICK_VOX;
    if (tid == 0)
    {
      // the stage brick n-1 was stored from becomes the landing zone of brick n + STAGES - 1
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      const unsigned m = n + MERGE_STAGES - 1;
      if (m < n_mine)
      {
        const int sm = (int)((it0 + m) % MERGE_STAGES);
        const size_t bm = (size_t)s_list[m] * WS_BRICK_VOX;
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
    unsigned pm[2], nib[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      const u64 k = k2[j];
      const bool occupied = key_is_candidate(k);
      // a winner at |value| == tau is a free-space candidate: it goes into the voxel's state like the ones
      // of the free-space phase (a real one of that phase still beats an interpolated one) and is folded
      // by fmerge_kernel; below tau the voxel is closed: final winner folded here, interpolated one parked
      const bool open_tau = occupied && key_abs_value(k) >= P.tau;
      const bool closed = occupied && !open_tau;
      const bool fin = closed && !key_interpolated(k);
      parked[j] = closed && !fin;
      nib[j] = open_tau ? (key_interpolated(k) ? VS_FREE_INT : VS_FREE_REAL)
                        : (closed ? (VS_CLOSED | (parked[j] ? VS_PARKED : 0u)) : 0u);
      if (closed) touched++;
      if (fin)
      {
        const int value = key_value(k);
        int weight = tsdf_weight_fd(value, P.tau, P.weight_epsilon, P.div_weps);
        if (key_interpolated(k)) weight = -weight;
        const int ew = entry_weight(e2[j]);
        written += ((weight > 0 && ew > 0) || (weight != 0 && ew <= 0)) ? 1u : 0u;
        e2[j] = merge_entry(e2[j], value, weight, P.max_weight);
      }
      pm[j] = __ballot_sync(FULL, parked[j]);
    }
    // entries and keys go back sector by sector, only where they changed (a surface brick is mostly empty space:
    // the bulk store of the whole brick wrote 6 KB where a few hundred bytes had changed)
    if (e2[0] != ee.x || e2[1] != ee.y) *reinterpret_cast<uint2 *>(g.grid + base + 2 * tid) = make_uint2(e2[0], e2[1]);
    s_state[st][tid] = (unsigned char)(nib[0] | (nib[1] << 4));
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
      g.brick_slot_base[base >> 9] = tot ? s_base : NO_PARK;
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
    if (kk.x != nk[0] || kk.y != nk[1]) *reinterpret_cast<ulonglong2 *>(g.keys + base + 2 * tid) = make_ulonglong2(nk[0], nk[1]);
    // generic-proxy writes to the state stage -> visible to the async proxy, then one thread stores it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0)
    {
      bulk_s2g(reinterpret_cast<unsigned char *>(g.vstate) + base / 2, s_state[st], WS_BRICK_VOX / 2);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  it0 += n_mine;
  }
  if (tid == 0)
  {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (it0) atomicAdd(&ctr->n_surf_bricks, it0);
  }

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

// ---- free-space merge ------------------------------------------------------------------------------
// Last kernel of a scan: for every brick the scan touched, the voxels that are not closed and hold a
// free-space candidate get (tau, +-WEIGHT_RESOLUTION) folded in (update_tsdf.cpp:542-560; real beats
// interpolated at equal |value|, :508-512), and the brick's per-voxel state is cleared for the next scan.
// 256 B of state + 1 KB of free-space bytes + 2 KB of entries per brick arrive as TMA bulk copies, three stages per
// CTA; the changed entries go back one by one (see below), the state is cleared by bulk stores of zeros.
#ifndef FMERGE_CTAS
#define FMERGE_CTAS 8
#endif
__global__ void __launch_bounds__(256, FMERGE_CTAS)
fmerge_kernel(const GridDesc g, const UpdateParams P, UpdateCounters *__restrict__ ctr)
{
  __shared__ __align__(128) uint32_t s_ent[MERGE_STAGES][WS_BRICK_VOX];
  __shared__ __align__(128) unsigned char s_free[MERGE_STAGES][2 * WS_BRICK_VOX];
  __shared__ __align__(128) unsigned char s_state[MERGE_STAGES][WS_BRICK_VOX / 2];
  __shared__ __align__(128) unsigned char s_zero[2 * WS_BRICK_VOX];
  __shared__ __align__(8) u64 s_bar[MERGE_STAGES];
  __shared__ unsigned s_list[BRICK_SHARE_CAP];
  __shared__ unsigned s_n;
  if (ctr->rec_overflow != 0u && ctr->pending_overflow == 0u) return;       // redone after the regrow
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned touched = 0, written = 0;
  reinterpret_cast<unsigned *>(s_zero)[tid] = 0u;
  if (tid == 0)
  {
    for (int st = 0; st < MERGE_STAGES; st++) mbar_init(&s_bar[st], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const unsigned stage_bytes = MERGE_ENT_BYTES + 2 * WS_BRICK_VOX + WS_BRICK_VOX / 2;

  const unsigned share = (i64)blockIdx.x < (i64)g.n_bricks ? (unsigned)(((i64)g.n_bricks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  unsigned it0 = 0u;
  for (unsigned r0 = 0u; r0 < share; r0 += BRICK_SHARE_CAP)
  {
  const unsigned n_mine = collect_bricks<false>(g, r0, s_list, &s_n);
  if (tid == 0)
    for (unsigned n = 0; n < MERGE_STAGES - 1 && n < n_mine; n++)
    {
      const int sn = (int)((it0 + n) % MERGE_STAGES);
      const size_t base = (size_t)s_list[n] * WS_BRICK_VOX;
      mbar_expect_tx(&s_bar[sn], stage_bytes);
      bulk_g2s(s_ent[sn], g.grid + base, MERGE_ENT_BYTES, &s_bar[sn]);
      bulk_g2s(s_free[sn], g.ffree + base * 2, 2 * WS_BRICK_VOX, &s_bar[sn]);
      bulk_g2s(s_state[sn], reinterpret_cast<const unsigned char *>(g.vstate) + base / 2, WS_BRICK_VOX / 2, &s_bar[sn]);
    }

  for (unsigned n = 0; n < n_mine; n++)
  {
    const int st = (int)((it0 + n) % MERGE_STAGES);
    const unsigned parity = ((it0 + n) / MERGE_STAGES) & 1u;
    const size_t base = (size_t)s_list[n] * WS_BRICK_VOX;
    if (tid == 0)
    {
      const unsigned m = n + MERGE_STAGES - 1;
      if (m < n_mine)
      {
        const int sm = (int)((it0 + m) % MERGE_STAGES);
        const size_t bm = (size_t)s_list[m] * WS_BRICK_VOX;
        mbar_expect_tx(&s_bar[sm], stage_bytes);
        bulk_g2s(s_ent[sm], g.grid + bm, MERGE_ENT_BYTES, &s_bar[sm]);
        bulk_g2s(s_free[sm], g.ffree + bm * 2, 2 * WS_BRICK_VOX, &s_bar[sm]);
        bulk_g2s(s_state[sm], reinterpret_cast<const unsigned char *>(g.vstate) + bm / 2, WS_BRICK_VOX / 2, &s_bar[sm]);
      }
    }
    mbar_wait(&s_bar[st], parity);

    // two voxels per thread: state nibbles (free-space winners of the surface phase, closed) + the bytes the
    // free-space phase stored
    const unsigned sb = s_state[st][tid];
    const unsigned fr = *reinterpret_cast<const unsigned short *>(&s_free[st][2 * tid]);
    const unsigned fi = *reinterpret_cast<const unsigned short *>(&s_free[st][WS_BRICK_VOX + 2 * tid]);
    if ((sb & 0x33u) | fr | fi)
    {
      const uint2 ee = *reinterpret_cast<const uint2 *>(&s_ent[st][2 * tid]);
      uint32_t e2[2] = { ee.x, ee.y };
      unsigned changed = 0u;
#pragma unroll
      for (int j = 0; j < 2; j++)
      {
        const unsigned nb = (sb >> (4 * j)) & 15u;
        const bool real = (nb & VS_FREE_REAL) || ((fr >> (8 * j)) & 0xFFu);
        const bool intp = (nb & VS_FREE_INT) || ((fi >> (8 * j)) & 0xFFu);
        if (!(real || intp) || (nb & VS_CLOSED)) continue;
        const int weight = real ? WS_WR : -WS_WR;
        const int ew = entry_weight(e2[j]);
        touched++;
        written += ((weight > 0 && ew > 0) || ew <= 0) ? 1u : 0u;
        const uint32_t ne = merge_entry(e2[j], P.tau, weight, P.max_weight);
        if (ne != e2[j]) changed |= 1u << j;
        e2[j] = ne;
      }
      // entry by entry, never the whole brick: only the sectors that changed go back to DRAM
      uint32_t *dst = g.grid + base + 2 * tid;
      if (changed == 3u) *reinterpret_cast<uint2 *>(dst) = make_uint2(e2[0], e2[1]);
      else if (changed == 1u) dst[0] = e2[0];
      else if (changed == 2u) dst[1] = e2[1];
    }
    __syncthreads();                       // everybody is done with this stage before it is refilled
    if (tid == 0)
    {
      bulk_s2g(g.ffree + base * 2, s_zero, 2 * WS_BRICK_VOX);
      bulk_s2g(reinterpret_cast<unsigned char *>(g.vstate) + base / 2, s_zero, WS_BRICK_VOX / 2);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      g.brick_slot_base[base >> 9] = NO_PARK;
    }
  }
  it0 += n_mine;
  }
  if (tid == 0)
  {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (it0) atomicAdd(&ctr->n_touched_bricks, it0);
  }

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

// Replay, first half (plain launch, runs beside the far-field free-space march on its own stream): the surface
// record against the parked voxels.  One warp per pair of 64-entry chunks; four independent
// record -> parked bit -> key -> parked-winner chains per lane are in flight at a time (the pass streams the
// record at DRAM speed and is otherwise latency bound); the per-voxel state (4 bits per voxel, L2 resident)
// spares the key lookup for the records whose voxel is not parked.
__global__ void __launch_bounds__(256, 4)
replay_scan_kernel(const GridDesc g, UpdateCounters *__restrict__ ctr, const unsigned pending_cap,
                   u64 *__restrict__ pend_key, const Rec *__restrict__ rec, const unsigned *__restrict__ chunk_fill,
                   const unsigned cap_chunks, Rec *__restrict__ list, const unsigned list_cap)
{
  if (ctr->n_pending == 0u || ctr->rec_overflow != 0u || ctr->pending_overflow != 0u) return;   // grid-uniform
  const int lane = threadIdx.x & 31;
  const unsigned gthread = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned gthreads = gridDim.x * blockDim.x;
  const unsigned gwarp = gthread >> 5, gwarps = gthreads >> 5;
  if (gthread == 0) ctr->t_phase[0] = global_ns();
  unsigned n_chunks = ctr->n_chunks;
  if (n_chunks > cap_chunks) n_chunks = cap_chunks;
  const bool wide = g.wide != 0;
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
      if (ok[u])
      {
        rr[u].ref = (u64)rec_refs(rec, c)[j];
        if (wide) rr[u].ref |= (u64)rec_refs(rec, c)[WS_REC_CHUNK + j] << 32;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) pw[u] = ok[u] ? __ldcg(&g.vstate[vstate_word(rr[u].ref)]) : 0u;
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      ok[u] = ok[u] && ((pw[u] >> vstate_shift(rr[u].ref)) & VS_PARKED);
      kv[u] = ok[u] ? __ldcg(&g.keys[rr[u].ref]) : WS_KEY_EMPTY;
      if (ok[u]) rr[u].key = rec_keys(rec, c0 + (unsigned)(u >> 1))[(unsigned)(u & 1) * 32u + (unsigned)lane];
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
      const bool hit = ok[u] && slot < pending_cap && key_seq(rr[u].key) > (kv[u] & WS_SEQ_MAX);
      if (hit) atomicMin(&pend_key[slot], rr[u].key);
      Rec e; e.key = rr[u].key; e.ref = (u64)slot;
      list_append(lw, hit, e, lane, list, list_cap, ctr);
    }
  }
  if (lw.used < WS_LIST_SPAN) list_pad(lw, lane, list, list_cap);
}

// Replay, second half.  Cooperative launch: settles every parked voxel on the device (see the file header, step 4).
__global__ void __launch_bounds__(256, 4)
replay_kernel(const GridDesc g, const UpdateParams P, UpdateCounters *__restrict__ ctr, const unsigned pending_cap,
              u64 *__restrict__ pend_addr, u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key,
              const Rec *__restrict__ list, const unsigned list_cap, unsigned *__restrict__ active0, unsigned *__restrict__ active1)
{
  cg::grid_group grid = cg::this_grid();
  unsigned n_pend = ctr->n_pending;
  if (n_pend > pending_cap) n_pend = pending_cap;
  if (n_pend == 0u || ctr->rec_overflow != 0u || ctr->pending_overflow != 0u) return;   // grid-uniform

  const int lane = threadIdx.x & 31;
  const unsigned gthread = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned gthreads = gridDim.x * blockDim.x;
  unsigned written = 0;
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
    unsigned n_list = ctr->n_list;
    if (n_list > list_cap) n_list = list_cap;
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
  ctr->item_counter[0] = 0u;
  ctr->item_counter[2] = 0u;
  ctr->item_counter[3] = 0u;
  ctr->gen_counter = 0u;
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
  // a candidate lies within delta_z (fan offset, |iv| = MR) + res (truncating divisions) + 1 (proj rounding) of the
  // ray point: voxels beyond that margin of a resident column cannot be reached
  const int margin = (int)((dz_max + 2 * (long long)h->res + 1) / h->res) + 1;
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
  // ... and so may every warp of the free-space march, which appends its hits on parked voxels
  const size_t list_entries = std::min<size_t>(chunks * WS_REC_CHUNK * 3 / 2 + (size_t)h->sm_count * (4 + LS_CTAS) * 8 * WS_LIST_SPAN,
                                               0xFFFFFF00u);
  WS_CUDA_OK(cudaMalloc(&h->d_list, list_entries * sizeof(Rec)));
  h->list_cap = list_entries;
  WS_CUDA_OK(cudaMalloc(&h->d_chunk_fill, chunks * sizeof(unsigned)));
  h->rec_cap_chunks = chunks;
}

// CTAs per SM of one launch (tuning knob for A/B runs; the default is what the launcher was tuned to)
static int ls_ctas_env(const char *name, int dflt)
{
  const char *e = std::getenv(name);
  const int v = e ? std::atoi(e) : dflt;
  return v >= 1 && v <= 8 ? v : dflt;
}

// first half of the replay (the surface record against the parked voxels) on `stream`
static void launch_replay_scan(ws_handle *h, cudaStream_t stream)
{
  replay_scan_kernel<<<h->sm_count * 4, 256, 0, stream>>>(h->g, h->d_counters, h->pending_cap, h->d_pend_key, h->d_rec, h->d_chunk_fill,
                                                         (unsigned)h->rec_cap_chunks, h->d_list, (unsigned)h->list_cap);
  h->launches++;
}

static void launch_replay(ws_handle *h, const UpdateParams &P)
{
  if (h->replay_blocks == 0)
  {
    int per_sm = 0;
    WS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, replay_kernel, 256, 0));
    if (per_sm < 1) throw std::runtime_error("replay_kernel does not fit on an SM");
    const int want = ls_ctas_env("WS_REPLAY_CTAS", 4);
    if (per_sm > want) per_sm = want;
    h->replay_blocks = per_sm * h->sm_count;
  }
  unsigned list_cap = (unsigned)h->list_cap;
  void *args[] = { (void *)&h->g, (void *)&P, (void *)&h->d_counters, (void *)&h->pending_cap,
                   (void *)&h->d_pend_addr, (void *)&h->d_pend_prev, (void *)&h->d_pend_key,
                   (void *)&h->d_list, (void *)&list_cap, (void *)&h->d_active[0], (void *)&h->d_active[1] };
  WS_CUDA_OK(cudaLaunchCooperativeKernel((const void *)replay_kernel, dim3(h->replay_blocks), dim3(256), args, 0, h->stream));
  h->launches++;
}

// src/warpsense/tsdf_mapping.cpp:77-85 + include/util/util.h:52-56 on the device, so update_tsdf can follow
// register_cloud on the stream without the host in between: new pose from the registration result X and the
// prior pose, scanner voxel = floor(t / res), up = third column of to_int_mat(pose).
//   compose_reference == 0: pose = X * prior (float32, the accumulation order of ws_compose_pose_host);
//   compose_reference != 0: App::update_pose_estimate (src/warpsense/app.cpp:172-176, same in
//                           src/cpu/fastsense.cpp:219-221): R = X.R * prior.R, t = prior.t + X.t.
// The prior comes by value, or (chain) is the pose this kernel left in `out` for the previous scan.
__global__ void pose_kernel(const float *__restrict__ X, const PoseArgs pa)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  pose_compute(X, pa);
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

// prior == nullptr: chain from the pose the previous scan left on the device
PoseArgs ws_pose_args(ws_handle *h, const float *prior, int compose_reference)
{
  if (!h->d_pose)
  {
    WS_CUDA_OK(cudaMalloc(&h->d_pose, sizeof(PoseDev)));
    WS_CUDA_OK(cudaMemsetAsync(h->d_pose, 0, sizeof(PoseDev), h->stream));
    WS_CUDA_OK(cudaMallocHost(&h->h_pose, sizeof(PoseDev)));
  }
  PoseArgs pa;
  pa.out = static_cast<PoseDev *>(h->d_pose);
  for (int i = 0; i < 16; i++) pa.prior[i] = prior ? prior[i] : 0.f;
  pa.chain = prior ? 0 : 1;
  pa.compose_reference = compose_reference;
  pa.res = h->res;
  long long lim = (1ll << 31) / h->res - h->tau - (1ll << 17);
  if (lim < 0 || std::getenv("WS_MARCH_GENERAL") || h->res < 4) lim = 0;
  pa.coord_lim = (int)lim;
  return pa;
}

void ws_launch_pose(ws_handle *h, const float *d_X, const float *prior, int compose_reference)
{
  const PoseArgs pa = ws_pose_args(h, prior, compose_reference);
  pose_kernel<<<1, 32, 0, h->stream>>>(d_X, pa);
  h->launches++;
}

#define LS_LAUNCH_ON(SURF_, ATOMIC_, TAB_, STREAM_) LS_LAUNCH_GRID(SURF_, ATOMIC_, TAB_, STREAM_, lockstep_blocks)
#define LS_LAUNCH_GRID(SURF_, ATOMIC_, TAB_, STREAM_, GRID_) LS_LAUNCH_GRID2(SURF_, ATOMIC_, TAB_, LS_NO_TAB, STREAM_, GRID_)
#define LS_LAUNCH_GRID2(SURF_, ATOMIC_, TAB_, TABB_, STREAM_, GRID_) LS_LAUNCH_PART(SURF_, ATOMIC_, TAB_, TABB_, STREAM_, GRID_, 0, GRID_)
// one of several launches that share a phase's item counters: CTAs [FIRST_, FIRST_ + GRID_) of ALL_
#define LS_LAUNCH_PART(SURF_, ATOMIC_, TAB_, TABB_, STREAM_, GRID_, FIRST_, ALL_)                                        \
  do {                                                                                                                   \
    if (wide) march_lockstep_kernel<SURF_, ATOMIC_, true><<<GRID_, MARCH_THREADS, tab_bytes, STREAM_>>>(                 \
        h->g, P, rays, h->d_grp_info, h->d_item_off, n_groups, h->d_counters, h->d_rec, h->d_chunk_fill, cap_chunks, d_pose, out, TAB_, TABB_, \
        (unsigned)(FIRST_) * MARCH_WARPS, (unsigned)(ALL_) * MARCH_WARPS);                                               \
    else march_lockstep_kernel<SURF_, ATOMIC_, false><<<GRID_, MARCH_THREADS, tab_bytes, STREAM_>>>(                     \
        h->g, P, rays, h->d_grp_info, h->d_item_off, n_groups, h->d_counters, h->d_rec, h->d_chunk_fill, cap_chunks, d_pose, out, TAB_, TABB_, \
        (unsigned)(FIRST_) * MARCH_WARPS, (unsigned)(ALL_) * MARCH_WARPS);                                               \
  } while (0)

// Enqueues one update_tsdf on the handle's stream; the work counters (and the device-side pose) land in
// `h_ctr` / `h_pose_out` (pinned) behind it.  ws_update_finish() waits, regrows the record if it overflowed and
// reports errors.
void ws_update_enqueue(ws_handle *h, ws_pt *d_pts, int n, const int scanner_pos_in[3], const int up_in[3],
                       bool pose_on_device, UpdateCounters *h_ctr, void *h_pose_out, const float *d_xform)
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
    if (h->res < 4) P.coord_lim = 0;                        // the signed 32-bit magic needs a divisor >= 3
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
  P.far_block = (P.far_len > 1 ? (P.far_len - 2) / P.half_res + 1 : 0) / LS_BLOCK;      // block of the first far step
  {
    // share of the near-field blocks that waits for the far-field launch (WS_NEAR_SPLIT: eighths of the near field
    // that stay beside the surface phase; 8 = all of it, the default).  Same-box A/B, scans/s end to end: 8 -> 880,
    // 5 -> 873, 4 -> 867 with two far-field CTAs per SM: the far-field launch has idle issue slots (57 % issue
    // active) but more warps do not fill them -- the marches are bound by the ALU pipe
    static const int keep8 = [] { const char *e = std::getenv("WS_NEAR_SPLIT"); const int v = e ? std::atoi(e) : 8; return v < 0 ? 0 : (v > 8 ? 8 : v); }();
    P.near_split_block = (P.far_block * keep8 + 4) / 8;
  }
  resident_x_intervals(h, P);

  cudaStream_t s = h->stream;
  WS_CUDA_OK(cudaMemsetAsync(h->d_counters, 0, sizeof(UpdateCounters), s));
  if (n > 0)
  {
    const int march_blocks = h->sm_count * MARCH_CTAS;
    const int lockstep_blocks = h->sm_count * LS_CTAS;
    const size_t tab_bytes = (size_t)(h->g.size[0] + h->g.size[1] + h->g.size[2]) * sizeof(unsigned);
    if (tab_bytes > 200 * 1024) throw std::invalid_argument("update_tsdf: map sides too large for the address tables");
    if (tab_bytes > 48 * 1024 && !h->tab_attr_set)
    {
      WS_CUDA_OK(cudaFuncSetAttribute(march_lockstep_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      WS_CUDA_OK(cudaFuncSetAttribute(march_lockstep_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      WS_CUDA_OK(cudaFuncSetAttribute(march_lockstep_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      WS_CUDA_OK(cudaFuncSetAttribute(march_lockstep_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      WS_CUDA_OK(cudaFuncSetAttribute(march_lockstep_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      WS_CUDA_OK(cudaFuncSetAttribute(march_lockstep_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      h->tab_attr_set = true;
    }
    unsigned cap_chunks = (unsigned)h->rec_cap_chunks;
    const long span = ws_span_begin(h, WS_TIMER_UPDATE);
    ws_timer_begin(h, WS_TIMER_MARCH);
    if ((size_t)n > h->rays_cap)
    {
      WS_CUDA_OK(cudaStreamSynchronize(s));
      cudaFree(h->d_rays); cudaFree(h->d_grp_info); cudaFree(h->d_item_off); cudaFree(h->d_gen_list);
      h->d_rays = nullptr; h->d_grp_info = nullptr; h->d_item_off = nullptr; h->d_gen_list = nullptr; h->rays_cap = 0;
      const size_t want = std::max<size_t>((size_t)n, 1 << 17);
      h->grp_cap = want / 32 + 1;
      WS_CUDA_OK(cudaMalloc(&h->d_rays, want * sizeof(RaySetup)));
      WS_CUDA_OK(cudaMalloc(&h->d_grp_info, 4 * h->grp_cap * sizeof(uint2)));
      WS_CUDA_OK(cudaMalloc(&h->d_item_off, 4 * (h->grp_cap + 1) * sizeof(unsigned)));
      WS_CUDA_OK(cudaMalloc(&h->d_gen_list, want * sizeof(unsigned)));
      h->rays_cap = want;
    }
    RaySetup *rays = static_cast<RaySetup *>(h->d_rays);
    const unsigned n_groups = (unsigned)((n + 31) / 32);
    LsOut out;
    out.pend_key = h->d_pend_key; out.list = h->d_list; out.pending_cap = h->pending_cap; out.list_cap = (unsigned)h->list_cap;
    const bool wide = h->g.wide != 0;
    if (!h->stream2)
    {
      WS_CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
      WS_CUDA_OK(cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
      WS_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      WS_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      WS_CUDA_OK(cudaEventCreateWithFlags(&h->ev_merged, cudaEventDisableTiming));
      WS_CUDA_OK(cudaEventCreateWithFlags(&h->ev_scanned, cudaEventDisableTiming));
    }
    cudaStream_t s2 = h->stream2;
    setup_kernel<<<(n + 255) / 256, 256, 0, s>>>(P, d_pts, rays, h->d_grp_info, n_groups, h->d_gen_list, h->d_counters, d_pose, d_xform);
    item_scan_kernel<<<4, 1024, 0, s>>>(h->d_grp_info, n_groups, h->d_item_off, h->d_counters);
    ws_timer_end(h);
    // Two streams from here.  Surface phase (this stream): keys, record, merge -> parked voxels.  The near-field part
    // of the free-space phase needs nothing from it and runs beside it on the second stream (both marches are
    // persistent grids; the block scheduler fills the SMs with whatever fits, nobody waits on the device); the
    // far-field part looks for parked voxels and follows the surface merge.
    WS_CUDA_OK(cudaEventRecord(h->ev_fork, s));
    WS_CUDA_OK(cudaStreamWaitEvent(s2, h->ev_fork, 0));
    ws_timer_begin(h, WS_TIMER_MARCH);
    LS_LAUNCH_GRID(true, true, 0u, s, h->sm_count * ls_ctas_env("WS_LS_GRID_S", 2));
    // (the rays outside the fast path -- none on a sane scan, the kernel returns at once -- stay on this stream: on a
    // side stream the kernel cannot become resident before a persistent march CTA retires, and the merge behind it
    // started 45 us late: 776 instead of 821 scans/s, same-box A/B)
    march_kernel<true><<<march_blocks, MARCH_THREADS, 0, s>>>(h->g, P, rays, h->d_gen_list, h->d_counters, h->d_rec,
                                                              h->d_chunk_fill, cap_chunks, d_pose);
    ws_timer_end(h);
    ws_timer_begin(h, WS_TIMER_MARCH, s2);
    LS_LAUNCH_GRID(false, true, 1u, s2, h->sm_count * ls_ctas_env("WS_LS_GRID_N", 3));
    ws_timer_end(h, s2);
    WS_CUDA_OK(cudaEventRecord(h->ev_join, s2));
    ws_timer_begin(h, WS_TIMER_MERGE);
    merge_kernel<<<h->sm_count * MERGE_CTAS, 256, 0, s>>>(h->g, P, h->d_counters, h->pending_cap,
                                                          h->d_pend_addr, h->d_pend_prev, h->d_pend_key);
    ws_timer_end(h);
    // the record pass of the replay (streams the record, DRAM bound) beside the far-field march (compute bound)
    WS_CUDA_OK(cudaEventRecord(h->ev_merged, s));
    WS_CUDA_OK(cudaStreamWaitEvent(h->stream3, h->ev_merged, 0));
    ws_timer_begin(h, WS_TIMER_REPLAY, h->stream3);
    launch_replay_scan(h, h->stream3);
    ws_timer_end(h, h->stream3);
    WS_CUDA_OK(cudaEventRecord(h->ev_scanned, h->stream3));
    // The far field: two CTAs per SM -- they leave the record pass the registers of one CTA of its own.  Three from the
    // start keep the record pass off the SMs until they retire (it then ends after the march and the replay rounds wait
    // for it: 892 scans/s on one box, 852 on the next); WS_LS_GRID_F2 more per SM can follow on the second stream once the
    // record pass is through, drawing from the same item counters (measured: no gain, 882 either way -- off by default).
    const int far_a = h->sm_count * ls_ctas_env("WS_LS_GRID_F", 2), far_b = h->sm_count * ls_ctas_env("WS_LS_GRID_F2", 0);
    ws_timer_begin(h, WS_TIMER_MARCH);
    LS_LAUNCH_PART(false, true, 2u, 3u, s, far_a, 0, far_a + far_b);
    ws_timer_end(h);
    if (far_b > 0)
    {
      WS_CUDA_OK(cudaStreamWaitEvent(s2, h->ev_scanned, 0));
      LS_LAUNCH_PART(false, true, 2u, 3u, s2, far_b, far_a, far_a + far_b);
      WS_CUDA_OK(cudaEventRecord(h->ev_join, s2));
      h->launches++;
    }
    WS_CUDA_OK(cudaStreamWaitEvent(s, h->ev_join, 0));
    WS_CUDA_OK(cudaStreamWaitEvent(s, h->ev_scanned, 0));
    // (the replay rounds and the free-space merge write disjoint voxels and could run side by side, but the rounds
    // need every register of the SMs to hide their latency and the merge every CTA slot: measured, no gain)
    ws_timer_begin(h, WS_TIMER_REPLAY);
    launch_replay(h, P);
    ws_timer_end(h);
    ws_timer_begin(h, WS_TIMER_MERGE);
    fmerge_kernel<<<h->sm_count * FMERGE_CTAS, 256, 0, s>>>(h->g, P, h->d_counters);
    ws_timer_end(h);
    ws_span_end(h, span);
    h->launches += 9;        // + the two replay launches counted where they are made
  }
  WS_CUDA_OK(cudaMemcpyAsync(h_ctr, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
  if (pose_on_device)
    WS_CUDA_OK(cudaMemcpyAsync(h_pose_out, h->d_pose, sizeof(PoseDev), cudaMemcpyDeviceToHost, s));
  // what ws_update_finish needs should the record have to grow
  h->upd_P = P; h->upd_n = n; h->upd_pose_on_device = pose_on_device;
}

// Second half of an update: wait for the stream (or `done`), regrow + regenerate the record if it overflowed,
// publish the counters, throw on errors.
void ws_update_finish(ws_handle *h, UpdateCounters *h_ctr, void *h_pose_out, cudaEvent_t done)
{
  cudaStream_t s = h->stream;
  if (done) WS_CUDA_OK(cudaEventSynchronize(done));
  else WS_CUDA_OK(cudaStreamSynchronize(s));
  const UpdateParams &P = h->upd_P;
  const int n = h->upd_n;
  if (n > 0)
  {
    const PoseDev *d_pose = h->upd_pose_on_device ? static_cast<const PoseDev *>(h->d_pose) : nullptr;
    const int march_blocks = h->sm_count * MARCH_CTAS;
    const int lockstep_blocks = h->sm_count * LS_CTAS;
    const size_t tab_bytes = (size_t)(h->g.size[0] + h->g.size[1] + h->g.size[2]) * sizeof(unsigned);
    RaySetup *rays = static_cast<RaySetup *>(h->d_rays);
    const unsigned n_groups = (unsigned)((n + 31) / 32);
    const bool wide = h->g.wide != 0;
    unsigned cap_chunks = (unsigned)h->rec_cap_chunks;
    LsOut out;
    out.pend_key = h->d_pend_key; out.list = h->d_list; out.pending_cap = h->pending_cap; out.list_cap = (unsigned)h->list_cap;
    // the record did not fit: grow it, regenerate it from the far part of the surface phase, then run what
    // was held back (free-space phase, replay, free-space merge)
    int guard = 0;
    while (h_ctr->rec_overflow != 0u && h_ctr->pending_overflow == 0u)
    {
      if (h->track_in_flight > 1)
        throw std::logic_error("update_tsdf: candidate record overflow with another scan already enqueued (raise WS_RECORD_CAP)");
      if (++guard > 8) throw std::runtime_error("update_tsdf: candidate record keeps overflowing");
      size_t want = (size_t)h_ctr->n_chunks + (size_t)h_ctr->n_chunks / 4 + 1024;
      if (want > h->rec_max_chunks) want = h->rec_max_chunks;
      if (want <= h->rec_cap_chunks)
        throw std::runtime_error("update_tsdf: candidate record exceeds WS_RECORD_MAX");
      ensure_record(h, want);
      cap_chunks = (unsigned)h->rec_cap_chunks;
      out.list = h->d_list; out.list_cap = (unsigned)h->list_cap;
      rec_reset_kernel<<<1, 1, 0, s>>>(h->d_counters);
      LS_LAUNCH_ON(true, false, 0u, s);
      march_kernel<false><<<march_blocks, MARCH_THREADS, 0, s>>>(h->g, P, rays, h->d_gen_list, h->d_counters, h->d_rec,
                                                                 h->d_chunk_fill, cap_chunks, d_pose);
      LS_LAUNCH_GRID2(false, true, 2u, 3u, s, lockstep_blocks);
      launch_replay_scan(h, s);
      launch_replay(h, P);
      fmerge_kernel<<<h->sm_count * FMERGE_CTAS, 256, 0, s>>>(h->g, P, h->d_counters);
      h->launches += 6;
      WS_CUDA_OK(cudaMemcpyAsync(h_ctr, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
      WS_CUDA_OK(cudaStreamSynchronize(s));
      h->record_regrows++;
    }
  }
  (void)h_pose_out;
  WS_CUDA_OK(cudaGetLastError());
  h->last_counters = *h_ctr;
  if (h->last_counters.pending_overflow)
    throw std::runtime_error("update_tsdf: pending-voxel capacity exceeded (raise WS_PENDING_CAP)");
  if (h->last_counters.n_pending != 0u)
    throw std::runtime_error("update_tsdf: parked voxels were left unsettled");
  if (h->last_counters.error & 1u)
    throw std::runtime_error("update_tsdf: ray too long for the candidate order field (march steps > 32768 or fan > 64)");
  if (h->last_counters.error & 2u)
    throw std::runtime_error("update_tsdf: replay list overflow (raise WS_RECORD_CAP)");
}

void ws_launch_update(ws_handle *h, const ws_pt *d_pts, int n, const int scanner_pos[3], const int up[3], bool pose_on_device)
{
  ws_update_enqueue(h, const_cast<ws_pt *>(d_pts), n, scanner_pos, up, pose_on_device, h->h_counters, h->h_pose);   // no transform: read only
  ws_update_finish(h, h->h_counters, h->h_pose, nullptr);
}

void ws_update_alloc(ws_handle *h, size_t initial_chunks, size_t max_chunks)
{
  h->rec_max_chunks = max_chunks;
  ensure_record(h, initial_chunks < max_chunks ? initial_chunks : max_chunks);
}
