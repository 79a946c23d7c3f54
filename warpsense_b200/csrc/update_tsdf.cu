// update_tsdf.cu -- TSDF volume update for sm_100a.
//
// Replaces the reference's TSDFCuda::update_tsdf (src/warpsense/cuda/update_tsdf.cu:13-166 under
// /root/reference) with kernels whose RESULT equals the reference's CPU path
// (src/cpu/update_tsdf.cpp:397-564), which is the parity target.
//
// Pipeline per scan (all on one stream, no host round trip in the common case):
//   1. march_kernel<true>   one warp per ray, lanes stride over the res/2 march steps; every candidate
//                           (voxel, value, real|interpolated, order) becomes ONE 64-bit atomicMin on the
//                           voxel's key -- see ws_common.cuh / DESIGN.md for why min over
//                           (|value|, interpolated, order) reproduces the sequential rule at :508-512;
//                           first touch of a brick appends it to the touched-brick list.
//   2. merge_kernel         persistent blocks stream the touched bricks (4 KB keys + 2 KB entries each),
//                           fold final winners into the grid (:542-560), reset the keys, and park the
//                           rare voxels whose winner is an interpolated candidate below tau ("pending").
//   3. march_kernel<false> + resolve_kernel, repeated while pending voxels remain: replay only the far
//                           part of the rays against the pending voxels, restricted to candidates later
//                           in the reference's order than the current winner.
#include <stdexcept>
#include "ws_internal.h"

#define FULL 0xFFFFFFFFu
#define PEND_DONE 0xFFFFFFFFFFFFFFFFull

namespace {

struct Ray
{
  int p[3];
  int d[3];
  int distance;
  i64 iv[3];
  FastDiv div_dist;
  int n_steps;
  bool valid;
};

// per-ray setup: update_tsdf.cpp:420-446
WS_D Ray ray_setup(const GridDesc &g, const UpdateParams &P, const ws_pt pt)
{
  Ray r;
  r.valid = false;
  r.n_steps = 0;
  r.p[0] = pt.x; r.p[1] = pt.y; r.p[2] = pt.z;
#pragma unroll
  for (int a = 0; a < 3; a++) r.d[a] = wsub(r.p[a], P.pos_mm[a]);                               // :422
  int sq = wadd(wadd(wmul(r.d[0], r.d[0]), wmul(r.d[1], r.d[1])), wmul(r.d[2], r.d[2]));
  if (sq <= 0) return r;             // distance == 0 (:424) or int32 norm overflow (out of contract)
  r.distance = isqrt31(sq);                                                                      // :423
  if (r.distance == 0) return r;
  if (!grid_in_bounds(g, fd_sdiv(r.p[0], P.div_res), fd_sdiv(r.p[1], P.div_res), fd_sdiv(r.p[2], P.div_res)))
    return r;                                                                                    // :430-434
  i64 nd[3], c1[3];
#pragma unroll
  for (int a = 0; a < 3; a++) nd[a] = ((i64)r.d[a] * WS_MR) / r.distance;                        // :438
  c1[0] = div_mr64(nd[1] * P.up[2] - nd[2] * P.up[1]);                                           // :439
  c1[1] = div_mr64(nd[2] * P.up[0] - nd[0] * P.up[2]);
  c1[2] = div_mr64(nd[0] * P.up[1] - nd[1] * P.up[0]);
  i64 iv0 = nd[1] * c1[2] - nd[2] * c1[1];
  i64 iv1 = nd[2] * c1[0] - nd[0] * c1[2];
  i64 iv2 = nd[0] * c1[1] - nd[1] * c1[0];
  i64 isq = (i64)((u64)iv0 * (u64)iv0 + (u64)iv1 * (u64)iv1 + (u64)iv2 * (u64)iv2);
  i64 inorm = __double2ll_rz(sqrt(__ll2double_rn(isq)));                                         // :440 (FP64, like Eigen)
  if (inorm <= 0) return r;                                                                      // :441-445
  r.iv[0] = (iv0 * WS_MR) / inorm;                                                               // :446
  r.iv[1] = (iv1 * WS_MR) / inorm;
  r.iv[2] = (iv2 * WS_MR) / inorm;
  r.div_dist.d = (unsigned)r.distance;
  r.div_dist.M = r.distance <= 1 ? 0ull : (~0ull) / (unsigned)r.distance + 1ull;
  // len = 1, 1+h, ... <= distance + tau  (:450)
  r.n_steps = (r.distance + P.tau - 1) / P.half_res + 1;
  r.valid = true;
  return r;
}

WS_D void step_index(const UpdateParams &P, const Ray &r, int len, int proj[3], int idx[3])
{
#pragma unroll
  for (int a = 0; a < 3; a++)
  {
    proj[a] = wadd(P.pos_mm[a], fd_sdiv(wmul(r.d[a], len), r.div_dist));                         // :452
    idx[a] = fd_sdiv(proj[a], P.div_res);                                                        // :453
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(256)
march_kernel(const GridDesc g, const UpdateParams P, const ws_pt *__restrict__ pts,
             unsigned *__restrict__ brick_list, UpdateCounters *__restrict__ ctr,
             const u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key)
{
  if (!FIRST && ctr->n_pending == 0) return;   // nothing parked: the replay round is a no-op

  const int lane = threadIdx.x & 31;
  const int ray_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  unsigned n_cand = 0;
  unsigned err = 0;

  if (ray_id < P.n_points)
  {
    const ws_pt pt = pts[ray_id];
    const Ray r = ray_setup(g, P, pt);
    if (r.valid)
    {
      int start = 0;
      if (!FIRST && P.far_only && P.far_len > 1) start = (P.far_len - 1) / P.half_res;
      if (r.n_steps > (1 << WS_SEQ_MARCH_BITS)) err |= 1u;

      int carry_x = 0, carry_y = 0;
      if (start > 0 && start < r.n_steps)
      {
        int pj[3], ix[3];
        step_index(P, r, 1 + (start - 1) * P.half_res, pj, ix);
        carry_x = ix[0]; carry_y = ix[1];
      }

      for (int base = start; base < r.n_steps; base += 32)
      {
        const int i = base + lane;
        const int len = 1 + i * P.half_res;
        int proj[3], index[3];
        step_index(P, r, len, proj, index);

        int px = __shfl_up_sync(FULL, index[0], 1);
        int py = __shfl_up_sync(FULL, index[1], 1);
        if (lane == 0) { px = carry_x; py = carry_y; }
        carry_x = __shfl_sync(FULL, index[0], 31);
        carry_y = __shfl_sync(FULL, index[1], 31);

        bool active = i < r.n_steps;
        if (i > 0 && index[0] == px && index[1] == py) active = false;                           // :455-458
        if (active && !grid_in_bounds(g, index[0], index[1], index[2])) active = false;          // :460-463
        if (!active) continue;

        // distance of the hit to the centre of the marched voxel (:466-472)
        int tcx = wadd(wmul(index[0], P.res), P.half_res);
        int tcy = wadd(wmul(index[1], P.res), P.half_res);
        int tcz = wadd(wmul(index[2], P.res), P.half_res);
        int ex = wsub(r.p[0], tcx), ey = wsub(r.p[1], tcy), ez = wsub(r.p[2], tcz);
        int vsq = wadd(wadd(wmul(ex, ex), wmul(ey, ey)), wmul(ez, ez));
        int value = vsq < 0 ? P.tau : isqrt31(vsq);
        value = value < P.tau ? value : P.tau;
        if (len > r.distance) value = -value;

        int weight = WS_WR;                                                                      // :475-479
        if (value < -P.weight_epsilon) weight = (int)fd_udiv((unsigned)(WS_WR * (P.tau + value)), P.div_weps);
        if (weight == 0) continue;                                                               // :480-483

        const int delta_z = div_mr32(wmul(P.dz_per_distance, len));                              // :485
        const int iter_steps = fd_sdiv(delta_z * 2, P.div_res) + 1;                              // :486
        const int mid = fd_sdiv(delta_z, P.div_res);                                             // :487
        int lowest[3];
#pragma unroll
        for (int a = 0; a < 3; a++) lowest[a] = wsub(proj[a], (int)div_mr64((i64)delta_z * r.iv[a]));   // :488
        if (iter_steps > (1 << WS_SEQ_STEP_BITS)) err |= 1u;

        for (int step = 0; step < iter_steps; ++step)                                            // :491
        {
          int v[3];
          const i64 sr = (i64)wmul(step, P.res);
#pragma unroll
          for (int a = 0; a < 3; a++)
            v[a] = fd_sdiv(wadd(lowest[a], (int)div_mr64(sr * r.iv[a])), P.div_res);             // :493
          if (!grid_in_bounds(g, v[0], v[1], v[2])) continue;                                    // :495-498
          n_cand++;

          const int rx = ring_coord(v[0], g.pos[0], g.offset[0], g.size[0]);
          const int ry = ring_coord(v[1], g.pos[1], g.offset[1], g.size[1]);
          const int rz = ring_coord(v[2], g.pos[2], g.offset[2], g.size[2]);
          const i64 brick = brick_of(g, rx, ry, rz);
          if (brick < 0) continue;                       // column lives on another rank
          const i64 addr = brick * WS_BRICK_VOX + brick_local(rx, ry, rz);
          const bool interp = step != mid;                                                       // :503-506
          const u64 seq = make_seq((unsigned)ray_id, (unsigned)i, (unsigned)step);
          const u64 key = make_key(value, interp, seq);

          if (FIRST)
          {
            atomicMin(&g.keys[addr], key);                                                       // :508-512
            // first touch of a brick: publish it for the merge pass (one probe per distinct brick per warp)
            const unsigned m = __match_any_sync(__activemask(), (unsigned)brick);
            if ((__ffs(m) - 1) == lane && __ldcg(&g.brick_flag[brick]) == 0u)
            {
              if (atomicExch(&g.brick_flag[brick], 1u) == 0u)
                brick_list[atomicAdd(&ctr->n_touched_bricks, 1u)] = (unsigned)brick;
            }
          }
          else
          {
            const u64 cur = __ldcg(&g.keys[addr]);
            if (key_is_pending(cur))
            {
              const unsigned slot = (unsigned)(cur & 0xFFFFFFFFull);
              if (seq > key_seq(pend_prev[slot])) atomicMin(&pend_key[slot], key);
            }
          }
        }
      }
    }
  }

  if (FIRST)
  {
    // candidate counter (work statistics): warp shuffle, then one atomic per warp leader
    for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);
    __shared__ unsigned s_cand;
    if (threadIdx.x == 0) s_cand = 0;
    __syncthreads();
    if (lane == 0 && n_cand) atomicAdd(&s_cand, n_cand);
    __syncthreads();
    if (threadIdx.x == 0 && s_cand) atomicAdd(&ctr->n_candidates, (unsigned long long)s_cand);
  }
  if (err) atomicOr(&ctr->error, err);
}

// final winner -> grid entry (update_tsdf.cpp:542-560); returns 1 if the entry changed
WS_D unsigned apply_winner(const GridDesc &g, const UpdateParams &P, i64 addr, u64 key)
{
  const int value = key_value(key);
  int weight = tsdf_weight(value, P.tau, P.weight_epsilon);
  if (key_interpolated(key)) weight = -weight;
  const uint32_t e = g.grid[addr];
  const uint32_t n = merge_entry(e, value, weight, P.max_weight);
  if (n != e) g.grid[addr] = n;
  const int ew = entry_weight(e);
  return ((weight > 0 && ew > 0) || (weight != 0 && ew <= 0)) ? 1u : 0u;
}

WS_D bool winner_is_final(u64 key, int tau)
{
  // a real candidate wins outright; an interpolated one at |v| == tau cannot be followed by anything
  // with a larger |v| (values are clamped to tau), so it is the reference's last write as well
  return !key_interpolated(key) || key_abs_value(key) >= tau;
}

__global__ void __launch_bounds__(256)
merge_kernel(const GridDesc g, const UpdateParams P, const unsigned *__restrict__ brick_list,
             UpdateCounters *__restrict__ ctr, unsigned pending_cap,
             u64 *__restrict__ pend_addr, u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key)
{
  const unsigned n_tb = ctr->n_touched_bricks;
  unsigned touched = 0, written = 0;
  for (unsigned bi = blockIdx.x; bi < n_tb; bi += gridDim.x)
  {
    const i64 brick = brick_list[bi];
    const i64 base = brick * WS_BRICK_VOX + 2 * threadIdx.x;
    const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(&g.keys[base]);
    const u64 k2[2] = { kk.x, kk.y };
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      const u64 k = k2[j];
      const bool occupied = k != WS_KEY_EMPTY;
      const bool fin = occupied && winner_is_final(k, P.tau);
      const bool park = occupied && !fin;
      const i64 addr = base + j;
      if (occupied) touched++;
      if (fin)
      {
        written += apply_winner(g, P, addr, k);
        g.keys[addr] = WS_KEY_EMPTY;
      }
      // parked voxels: one slot counter bump per warp
      const unsigned pm = __ballot_sync(FULL, park);
      if (pm)
      {
        const int lane = threadIdx.x & 31;
        unsigned first = 0;
        if (lane == (__ffs(pm) - 1)) first = atomicAdd(&ctr->n_pending, (unsigned)__popc(pm));
        first = __shfl_sync(FULL, first, __ffs(pm) - 1);
        if (park)
        {
          const unsigned slot = first + (unsigned)__popc(pm & ((1u << lane) - 1u));
          if (slot < pending_cap)
          {
            pend_addr[slot] = (u64)addr;
            pend_prev[slot] = k;
            pend_key[slot] = WS_KEY_EMPTY;
            g.keys[addr] = WS_KEY_PENDING_TAG | (u64)slot;
          }
          else
          {
            atomicAdd(&ctr->pending_overflow, 1u);
            g.keys[addr] = WS_KEY_EMPTY;
          }
        }
      }
    }
    if (threadIdx.x == 0) g.brick_flag[brick] = 0u;
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    touched += __shfl_down_sync(FULL, touched, o);
    written += __shfl_down_sync(FULL, written, o);
  }
  if ((threadIdx.x & 31) == 0)
  {
    if (touched) atomicAdd(&ctr->n_touched, (unsigned long long)touched);
    if (written) atomicAdd(&ctr->n_written, (unsigned long long)written);
  }
}

// after a replay round: settle every parked voxel that now has its final winner
__global__ void __launch_bounds__(256)
resolve_kernel(const GridDesc g, const UpdateParams P, UpdateCounters *__restrict__ ctr, unsigned pending_cap,
               u64 *__restrict__ pend_addr, u64 *__restrict__ pend_prev, u64 *__restrict__ pend_key)
{
  unsigned n = ctr->n_pending;
  if (n > pending_cap) n = pending_cap;
  unsigned written = 0, still = 0;
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
  {
    const u64 addr = pend_addr[s];
    if (addr == PEND_DONE) continue;
    const u64 k2 = pend_key[s];
    u64 fin;
    if (k2 == WS_KEY_EMPTY) fin = pend_prev[s];          // nothing later in the order: the parked winner stands
    else if (winner_is_final(k2, P.tau)) fin = k2;
    else
    {
      pend_prev[s] = k2;                                 // still interpolated: restart after it
      pend_key[s] = WS_KEY_EMPTY;
      still++;
      continue;
    }
    written += apply_winner(g, P, (i64)addr, fin);
    g.keys[addr] = WS_KEY_EMPTY;
    pend_addr[s] = PEND_DONE;
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    written += __shfl_down_sync(FULL, written, o);
    still += __shfl_down_sync(FULL, still, o);
  }
  if ((threadIdx.x & 31) == 0)
  {
    if (written) atomicAdd(&ctr->n_written, (unsigned long long)written);
    if (still) atomicAdd(&ctr->n_pending_next, still);
  }
}

// one thread: roll the pending counters over to the next replay round
__global__ void round_advance_kernel(UpdateCounters *ctr)
{
  if (ctr->n_pending == 0) return;
  if (ctr->rounds == 0) ctr->n_parked = ctr->n_pending;
  ctr->rounds += 1;
  if (ctr->n_pending_next == 0) ctr->n_pending = 0;   // everything settled: later rounds become no-ops
  ctr->n_pending_next = 0;
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

void ws_launch_update(ws_handle *h, const ws_pt *d_pts, int n, const int scanner_pos[3], const int up[3])
{
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
  P.half_res = h->res / 2;
  P.n_points = n;
  P.far_only = 0;
  P.far_len = far_start_len(h->res, P.dz_per_distance);

  cudaStream_t s = h->stream;
  WS_CUDA_OK(cudaMemsetAsync(h->d_counters, 0, sizeof(UpdateCounters), s));
  if (n > 0)
  {
    const int rays_per_block = 8;
    const int blocks = (n + rays_per_block - 1) / rays_per_block;
    ws_timer_begin(h, WS_TIMER_MARCH);
    march_kernel<true><<<blocks, 256, 0, s>>>(h->g, P, d_pts, h->d_brick_list, h->d_counters, nullptr, nullptr);
    ws_timer_end(h);
    h->launches += 2;
    const int dev_sms = h->sm_count;
    ws_timer_begin(h, WS_TIMER_MERGE);
    merge_kernel<<<dev_sms * 8, 256, 0, s>>>(h->g, P, h->d_brick_list, h->d_counters, h->pending_cap,
                                            h->d_pend_addr, h->d_pend_prev, h->d_pend_key);
    ws_timer_end(h);
    // replay rounds for parked voxels; each is a no-op (early exit) once nothing is parked
    UpdateParams P2 = P;
    P2.far_only = 1;
    const int fixed_rounds = 3;
    for (int r = 0; r < fixed_rounds; r++)
    {
      march_kernel<false><<<blocks, 256, 0, s>>>(h->g, P2, d_pts, h->d_brick_list, h->d_counters,
                                                 h->d_pend_prev, h->d_pend_key);
      resolve_kernel<<<dev_sms * 4, 256, 0, s>>>(h->g, P2, h->d_counters, h->pending_cap,
                                                 h->d_pend_addr, h->d_pend_prev, h->d_pend_key);
      round_advance_kernel<<<1, 1, 0, s>>>(h->d_counters);
      h->launches += 3;
    }
    WS_CUDA_OK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
    WS_CUDA_OK(cudaStreamSynchronize(s));
    // extremely deep chains of interpolated winners: keep replaying until settled
    int guard = 0;
    while (h->h_counters->n_pending != 0 && h->h_counters->pending_overflow == 0)
    {
      if (++guard > 100000) throw std::runtime_error("update_tsdf: replay rounds do not converge");
      march_kernel<false><<<blocks, 256, 0, s>>>(h->g, P2, d_pts, h->d_brick_list, h->d_counters,
                                                 h->d_pend_prev, h->d_pend_key);
      resolve_kernel<<<dev_sms * 4, 256, 0, s>>>(h->g, P2, h->d_counters, h->pending_cap,
                                                 h->d_pend_addr, h->d_pend_prev, h->d_pend_key);
      round_advance_kernel<<<1, 1, 0, s>>>(h->d_counters);
      h->launches += 3;
      WS_CUDA_OK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
      WS_CUDA_OK(cudaStreamSynchronize(s));
    }
  }
  else
  {
    WS_CUDA_OK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(UpdateCounters), cudaMemcpyDeviceToHost, s));
    WS_CUDA_OK(cudaStreamSynchronize(s));
  }
  WS_CUDA_OK(cudaGetLastError());
  h->last_counters = *h->h_counters;
  if (h->last_counters.pending_overflow)
    throw std::runtime_error("update_tsdf: pending-voxel capacity exceeded (raise WS_PENDING_CAP)");
  if (h->last_counters.error & 1u)
    throw std::runtime_error("update_tsdf: ray too long for the candidate order field (march steps > 32768 or fan > 64)");
}
