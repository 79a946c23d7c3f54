// Host-side unit test of warpsense_b200/csrc/march_math.cuh: the fast 32-bit arithmetic of the ray march
// must equal the literal formulas of update_tsdf.cpp:450-506 (as restated in oracle/ws_oracle.c) for every
// ray that `ray_is_small` admits.  Compiled with nvcc, runs on the CPU (no device code is launched).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../warpsense_b200/csrc/march_math.cuh"

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (failures < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } failures++; } } while (0)

static int norm_i32(int x, int y, int z)
{
  int sq = (int)((unsigned)x * (unsigned)x + (unsigned)y * (unsigned)y + (unsigned)z * (unsigned)z);
  return (int)std::sqrt((double)sq);
}

int main()
{
  std::mt19937_64 rng(12345);
  auto uni = [&](long long lo, long long hi) { return (long long)(lo + (long long)(rng() % (unsigned long long)(hi - lo + 1))); };

  // 1. 32-bit magic division, exact for |n| * d < 2^32
  const unsigned divisors[] = { 2, 3, 5, 7, 10, 16, 20, 25, 50, 64, 100, 127, 128, 250, 999, 1000, 1023, 4096, 65535 };
  for (unsigned d : divisors)
  {
    const FastDiv32 f = make_fastdiv32(d);
    const long long lim = (1ll << 32) / d - 1;
    for (int t = 0; t < 200000; t++)
    {
      long long n = t < 1000 ? lim - t : uni(0, lim);
      if (n < 0) n = 0;
      CHECK(fd32_udiv((unsigned)n, f) == (unsigned)(n / d), "udiv %lld / %u", n, d);
      if (n < (1ll << 31))
      {
        CHECK(fd32_sdiv((int)n, f) == (int)(n / d), "sdiv %lld / %u", n, d);
        if (d >= 3) CHECK(fd32_sdiv_s((int)n, f) == (int)(n / d) && fd32_sdiv_s((int)-n, f) == (int)((-n) / (long long)d), "sdiv_s %lld / %u", n, d);
        CHECK(fd32_sdiv((int)-n, f) == (int)((-n) / (long long)d), "sdiv -%lld / %u", n, d);
      }
    }
    for (long long n = 0; n < 70000 && n <= lim; n++)
      CHECK(fd32_sdiv((int)-n, f) == (int)((-n) / (long long)d) && fd32_udiv((unsigned)n, f) == (unsigned)(n / d), "small %lld / %u", n, d);
  }

  // 1b. the merge's running average (update_tsdf.cpp:549): float estimate + one correction == C++ `/`
  for (int t = 0; t < 4000000; t++)
  {
    int n, d;
    switch (t & 3)
    {
      case 0: { const int ev = (int)uni(-32768, 32767), ew = (int)uni(1, 32767), v = (int)uni(-32767, 32767), w = (int)uni(1, 64);
                n = ev * ew + v * w; d = ew + w; break; }                                   // what merge_entry feeds it
      case 1: d = (int)uni(1, 800); n = (int)uni(-(1ll << 31) + 1, (1ll << 31) - 1); break; // quotients above 2^20: falls back
      case 2: d = (int)uni(1, (1 << 25)); n = (int)uni(-(1ll << 31) + 1, (1ll << 31) - 1); break;
      default: { d = (int)uni(1, 70000); const long long k = uni(0, ((1ll << 31) - 2) / d); n = (int)(k * d + uni(-1, 1));
                 if (t & 4) n = -n; break; }                                                // multiples and their neighbours
    }
    CHECK(div_trunc_small(n, d) == n / d, "div_trunc_small %d / %d -> %d", n, d, div_trunc_small(n, d));
  }
  // 1c. the candidate weight (update_tsdf.cpp:475-479) through the per-scan magic
  for (int tau : { 1, 2, 9, 10, 11, 100, 600, 1000, 1234, 32767 })
  {
    const int eps = tau / 10;
    const FastDiv fw = make_fastdiv((unsigned)(tau - eps > 0 ? tau - eps : 1));
    for (int v = -tau; v <= tau; v++)
      CHECK(tsdf_weight_fd(v, tau, eps, fw) == tsdf_weight(v, tau, eps), "tsdf_weight tau %d value %d", tau, v);
  }

  // 1d. (n * r) / 32768 for n >= 0 as the shifted magnitude with r's sign (update_tsdf.cu mr_scaled)
  for (int t = 0; t < 4000000; t++)
  {
    const int r = (int)uni(-WS_MR, WS_MR), n = (int)uni(0, t & 1 ? 60000 : 2000);
    if ((long long)n * (r < 0 ? -r : r) >= (1ll << 31)) continue;
    const int ar = r < 0 ? -r : r, sr = r < 0 ? -1 : 1;
    const int fast = sr * (int)((unsigned)(n * ar) >> WS_MR_SHIFT);
    CHECK(fast == div_mr32(n * r) && fast == (int)(((long long)n * r) / WS_MR), "mr_scaled %d * %d", n, r);
  }

  // 2. reciprocal-based division with remainder
  for (int t = 0; t < 2000000; t++)
  {
    const unsigned dist = (unsigned)uni(1, t & 1 ? 70000 : (1 << 30));
    const unsigned n = (unsigned)uni(0, (1ll << 31) - 1);
    unsigned q, r;
    divrem_rcp(n, dist, 1.0 / (double)dist, q, r);
    CHECK(q == n / dist && r == n % dist, "divrem %u / %u -> %u r %u", n, dist, q, r);
    const u64 n64 = (u64)uni(0, (1ll << 46) - 1);
    CHECK(div_rcp64(n64, dist, 1.0 / (double)dist) == n64 / dist, "div_rcp64 %llu / %u", n64, dist);
    // exact multiples and neighbours
    const u64 k = (u64)uni(0, (1ll << 46) / dist);
    for (int o = -1; o <= 1; o++)
    {
      const u64 m = k * dist + (u64)(o + 1);
      if (m == 0 || m >= (1ull << 46)) continue;
      CHECK(div_rcp64(m - 1, dist, 1.0 / (double)dist) == (m - 1) / dist, "div_rcp64 edge %llu / %u", m - 1, dist);
    }
  }

  // 3. FP64 quotient of integers below 2^52 truncates to the integer quotient (iv normalisation, :446)
  for (int t = 0; t < 2000000; t++)
  {
    const long long b = uni(1, t & 1 ? (1ll << 31) : (1ll << 40));
    long long a = uni(0, (1ll << 52) - 1);
    if (t % 3 == 0) a = (a / b) * b;            // exact multiples
    if (t % 3 == 1 && a / b > 0) a = (a / b) * b - 1;
    CHECK((long long)((double)a / (double)b) == a / b, "fp64 quotient %lld / %lld", a, b);
  }

  // 4. DDA projection + magic voxel index against the literal formulas, lane-strided exactly like the kernel
  long long rays = 0, steps = 0;
  for (int t = 0; t < 6000; t++)
  {
    const int res = (int[]){ 20, 50, 64, 100, 33, 7 }[t % 6];
    const int h = res / 2, tau = (int[]){ 1000, 600, 32767 }[t % 3];
    const int dz = 100;
    const long long lim = (1ll << 31) / res - tau - (1ll << 17);
    const int coord_lim = lim > 0 ? (int)lim : 0;
    const int span = t % 5 == 0 ? 60000 : 26000;
    int pos[3], p[3], d[3];
    const long long centre = t % 7 == 0 ? uni(-(lim / 2), lim / 2) : uni(-2000, 2000);
    for (int a = 0; a < 3; a++)
    {
      pos[a] = (int)(centre + uni(-span / 2, span / 2));
      p[a] = (int)(centre + uni(-span / 2, span / 2));
      if (t % 11 == 0 && a == 2) p[a] = pos[a];            // degenerate axes
      d[a] = p[a] - pos[a];
    }
    const int dist = norm_i32(d[0], d[1], d[2]);
    if (dist == 0) continue;
    bool pos_ok = true;
    for (int a = 0; a < 3; a++) if (std::llabs((long long)pos[a]) >= lim) pos_ok = false;
    if (!pos_ok || !ray_is_small(d, p, dist, tau, h, dz, coord_lim)) continue;
    rays++;
    const FastDiv32 f = make_fastdiv32((unsigned)res);
    const int n_steps = (dist + tau - 1) / h + 1;
    const double rdist = 1.0 / (double)dist;
    for (int lane = 0; lane < 32; lane += (t % 4 == 0 ? 1 : 7))
    {
      DdaAxis s[3];
      for (int a = 0; a < 3; a++)
        dda_init(s[a], (unsigned)std::abs(d[a]), 1u + (unsigned)lane * (unsigned)h, 32u * (unsigned)h, (unsigned)dist, rdist);
      for (int base = 0; base < n_steps; base += 32)
      {
        const int i = base + lane;
        const int len = 1 + i * h;
        for (int a = 0; a < 3; a++)
        {
          const int proj = pos[a] + (d[a] < 0 ? -(int)s[a].q : (int)s[a].q);
          const int idx = fd32_sdiv(proj, f);
          dda_advance(s[a], (unsigned)dist);
          if (i >= n_steps) continue;
          const int want_proj = pos[a] + (int)(((long long)d[a] * len) / dist);
          CHECK(proj == want_proj, "proj axis %d i %d: %d vs %d (d %d dist %d)", a, i, proj, want_proj, d[a], dist);
          CHECK(idx == want_proj / res, "idx axis %d i %d", a, i);
        }
        if (i < n_steps) steps++;
        // fan arithmetic (:485-493) at this step, 32-bit against 64-bit
        if (i < n_steps && (i % 13) == 0)
        {
          const int delta_z = (dz * len) / WS_MR;
          const int iter_steps = (delta_z * 2) / res + 1, mid = delta_z / res;
          CHECK((int)fd32_udiv((unsigned)(delta_z * 2), f) + 1 == iter_steps && (int)fd32_udiv((unsigned)delta_z, f) == mid, "fan counts");
          for (int a = 0; a < 3; a++)
          {
            const int iv = (int)uni(-WS_MR, WS_MR);
            const int proj = pos[a] + (int)(((long long)d[a] * len) / dist);
            const int low64 = proj - (int)(((long long)delta_z * iv) / WS_MR);
            const int low32 = proj - div_mr32(delta_z * iv);
            CHECK(low64 == low32, "lowest");
            for (int step = 0; step < iter_steps && step < 64; step++)
            {
              const int v64 = (low64 + (int)(((long long)(step * res) * iv) / WS_MR)) / res;
              const int v32 = fd32_sdiv(low32 + div_mr32(step * res * iv), f);
              CHECK(v64 == v32, "fan voxel step %d", step);
            }
          }
        }
      }
    }
  }
  // 5. multi-GPU culling: every march step whose x lies in a resident interval is inside one of the step ranges
  long long culled_rays = 0, covered_steps = 0, skipped_steps = 0;
  for (int t = 0; t < 60000; t++)
  {
    const int res = (int[]){ 20, 50, 64, 100 }[t % 4], h = res / 2, tau = 1000;
    int n_xiv = 1 + (int)uni(0, 5);
    int lo[6], hi[6];
    long long x = uni(-30000, -20000);
    for (int k = 0; k < n_xiv; k++)
    {
      lo[k] = (int)x; x += uni(1, 6000); hi[k] = (int)x; x += uni(5 * res, 9000);     // sorted, disjoint
    }
    const int pos_x = (int)uni(-26000, 26000);
    int dx = (int)uni(-26000, 26000);
    if (t % 17 == 0) dx = 0;
    else if (t % 5 == 0) dx = (int)uni(-300, 300);      // rays almost along y: one millimetre of x is many march steps
    const int dy = (int)uni(-20000, 20000);
    const int dist = norm_i32(dx, dy, 0);
    if (dist == 0) continue;
    const int n_steps = (dist + tau - 1) / h + 1;
    int seg[12];
    ray_step_ranges<6>(n_xiv, lo, hi, pos_x, h, dx, dist, n_steps, seg);
    culled_rays++;
    int prev_end = -1;
    for (int k = 0; k < 6; k++)
    {
      if (seg[2 * k + 1] <= seg[2 * k]) { for (int q = k; q < 6; q++) CHECK(seg[2 * q] == 0 && seg[2 * q + 1] == 0, "ranges not packed"); break; }
      CHECK(seg[2 * k] >= 0 && seg[2 * k + 1] <= n_steps && seg[2 * k] > prev_end, "ranges not ascending / disjoint");
      prev_end = seg[2 * k + 1];
    }
    for (int i = 0; i < n_steps; i++)
    {
      const int len = 1 + i * h;
      const long long px = pos_x + ((long long)dx * len) / dist;
      bool need = false;
      for (int k = 0; k < n_xiv; k++) if (px >= lo[k] && px < hi[k]) need = true;
      bool have = false;
      for (int k = 0; k < 6; k++) if (i >= seg[2 * k] && i < seg[2 * k + 1]) have = true;
      CHECK(!need || have, "step %d (x %lld) of a ray with dx %d is resident but not in a range", i, px, dx);
      if (have) covered_steps++; else skipped_steps++;
    }
  }
  printf("culling: %lld rays, %lld steps kept, %lld skipped\n", culled_rays, covered_steps, skipped_steps);
  CHECK(culled_rays > 10000 && skipped_steps > covered_steps / 4, "culling coverage too small");

  printf("rays %lld steps %lld\n", rays, steps);
  CHECK(rays > 1000 && steps > 1000000, "coverage too small");
  if (failures) { printf("march math: %d FAILURES\n", failures); return 1; }
  printf("march math ok\n");
  return 0;
}
