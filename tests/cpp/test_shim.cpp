// C++ host-side test of include/warpsense_b200.hpp: the reference's own known-answer test for update_tsdf
// (test/map.cpp:9-90 == test/cuda.cpp:268-347 under /root/reference, "G1") and a registration round trip,
// written against the reference's class names.  Needs a CUDA device; built and run by tests/test_cpp_shim.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include "warpsense_b200.hpp"

#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); return 1; } \
  } while (0)

// include/warpsense/test/common.h:16-26
static int calc_weight(int value, int tau, int eps)
{
  int weight = warpsense_b200::WEIGHT_RESOLUTION;
  if (value < -eps) weight = warpsense_b200::WEIGHT_RESOLUTION * (tau + value) / (tau - eps);
  return weight;
}

int main()
{
  using namespace warpsense_b200;
  // ---- G1: one point, res 1000, tau 3000, 21^3 map -------------------------------------------------
  {
    Params params;
    params.map.resolution = 1000; params.map.tau = 3000; params.map.max_weight = 10 * WEIGHT_RESOLUTION;
    auto local_map = std::make_shared<HostLocalMap>(20, 20, 20, params.map.tau, 0);
    cuda::TSDFMapping gpu(params, local_map);
    std::vector<rmagine::Pointi> scan = { rmagine::Pointi(5500, 500, 500) };
    gpu.update_tsdf(scan, Matrix4f::Identity());
    gpu.get_tsdf_map();
    const int want[9] = { 0, 3000, 3000, 2000, 1000, 0, -1000, -2000, 3000 };
    const int eps = params.map.tau / 10;
    for (int k = 1; k <= 7; k++)
    {
      const TSDFEntry e = local_map->value(k, 0, 0);
      CHECK(e.value() == want[k]);
      CHECK(e.weight() == calc_weight(want[k], params.map.tau, eps));
    }
    CHECK(local_map->value(8, 0, 0).value() == 3000 && local_map->value(8, 0, 0).weight() == 0);
    const ws_update_counters c = gpu.tsdf()->counters();
    CHECK(c.n_candidates == 8 && c.n_touched == 8);
  }
  // ---- registration: build a wall, shift the cloud by (+120, -80, 0) mm, register it back ------------------
  {
    Params params;
    params.map.resolution = 64; params.map.tau = 600; params.map.max_weight = 10 * WEIGHT_RESOLUTION;
    params.registration.max_iterations = 50; params.registration.epsilon = 0.f;
    auto local_map = std::make_shared<HostLocalMap>(128, 128, 64, params.map.tau, 0);
    cuda::TSDFRegistration gpu(params, local_map);
    std::vector<rmagine::Pointi> scan;
    for (int a = -40; a <= 40; a++)
      for (int b = -12; b <= 12; b++)
      {
        scan.emplace_back(2500, a * 40, b * 40);      // wall x = 2.5 m
        scan.emplace_back(a * 40, 2200, b * 40);      // wall y = 2.2 m
        scan.emplace_back(-2300, a * 40, b * 40);
        scan.emplace_back(a * 40, -2600, b * 40);
      }
    for (int a = -30; a <= 30; a++)
      for (int b = -30; b <= 30; b++)
      {
        scan.emplace_back(a * 60, b * 60, -700);      // floor
        scan.emplace_back(a * 60, b * 60, 800);       // ceiling
      }
    gpu.update_tsdf(scan, Matrix4f::Identity());
    std::vector<rmagine::Pointi> cloud = scan;
    for (auto &p : cloud) { p.x += 120; p.y -= 80; }
    int it = 0;
    const Matrix4f T = gpu.register_cloud(cloud, Matrix4f::Identity(), &it);
    std::printf("register_cloud: %d iterations, t = (%.1f, %.1f, %.1f) mm\n", it, T(0, 3), T(1, 3), T(2, 3));
    CHECK(it == 50);
    // the reference's damped Gauss-Newton walks back slowly (the CPU oracle gives t = (-63, 60, 32) mm here after
    // 50 iterations; exact agreement with it is asserted by the Python parity tests): direction and progress
    CHECK(T(0, 3) < -30.f && T(0, 3) > -130.f && T(1, 3) > 30.f && T(1, 3) < 90.f);
    // the cloud came back transformed in place (src/cpu/registration.cpp:168-174): closer than the 200 mm it started at
    long err = 0;
    for (size_t i = 0; i < cloud.size(); i++)
      err += std::labs(cloud[i].x - scan[i].x) + std::labs(cloud[i].y - scan[i].y) + std::labs(cloud[i].z - scan[i].z);
    CHECK(err / (long)cloud.size() < 150);
    // single accumulation step through RegistrationCuda
    gpu.reg_->prepare_registration(scan);
    int64_t H[36], g[6]; int e = 0, c = 0;
    const Matrix4f I = Matrix4f::Identity();
    gpu.reg_->perform_registration(&I, H, g, e, c, params.map.resolution);
    CHECK(c > 1000);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) CHECK(H[i * 6 + j] == H[j * 6 + i]);
    // map shift on the device keeps voxels addressable at the same world coordinates
    const TSDFEntry before = [&] { gpu.get_tsdf_map(); return local_map->value(39, 0, 0); }();
    Matrix4f pose = Matrix4f::Identity();
    pose(0, 3) = 3500.f;                                // 3.5 m > params.map.shift (3 m)
    CHECK(gpu.map_shift(pose));
    gpu.get_tsdf_map();
    CHECK(local_map->get_pos()[0] == 54);               // floor(3500 / 64)
    const TSDFEntry after = local_map->value(39, 0, 0);
    CHECK(before.raw == after.raw);
    // App::preprocess on the device: duplicates collapse, the all-below-0.3 return is dropped (app.cpp:128)
    const float xyz[] = { 1.00f, 2.00f, 0.50f,   1.01f, 2.01f, 0.51f,   0.1f, 0.2f, -0.4f,   -3.0f, 0.5f, 0.2f };
    std::vector<rmagine::Pointi> pre;
    gpu.preprocess(xyz, 4, 12, Matrix4f::Identity(), pre);
    CHECK(pre.size() == 2);
    CHECK(pre[0].x == 992 && pre[0].y == 2016 && pre[0].z == 480);      // voxel centres at 64 mm, res/2 = 32
    // the fused per-scan call gives the pose = X * prior
    std::vector<rmagine::Pointi> cloud2 = scan;
    Matrix4f X{};
    int it2 = 0;
    const Matrix4f newpose = gpu.track_scan(cloud2, pose, &X, &it2);
    CHECK(it2 == 50);
    CHECK(std::fabs(newpose(0, 3) - (X(0, 3) + pose(0, 3))) < 1.0f);
    // global-map file: HDF5 signature and a plausible size (64^3 uint32 per stored chunk)
    const char *path = "/tmp/ws_shim_map.h5";
    gpu.export_map(path, { 1.f, 2.f, 3.f, 0.f, 0.f, 0.f, 1.f });
    FILE *fp = std::fopen(path, "rb");
    CHECK(fp != nullptr);
    if (fp)
    {
      unsigned char sig[8] = { 0 };
      CHECK(std::fread(sig, 1, 8, fp) == 8);
      CHECK(sig[0] == 0x89 && sig[1] == 'H' && sig[2] == 'D' && sig[3] == 'F');
      std::fseek(fp, 0, SEEK_END);
      CHECK(std::ftell(fp) > 64L * 64 * 64 * 4);
      std::fclose(fp);
      std::remove(path);
    }
  }
  // ---- featsense feed in C++: background map-shift thread, accumulate-while-shifting, flush ---------------------
  {
    Params params;
    params.map.resolution = 100; params.map.tau = 1000; params.map.max_weight = 10 * WEIGHT_RESOLUTION;
    params.map.shift = 1.0f; params.map.update_distance = 0.2f;
    auto local_map = std::make_shared<HostLocalMap>(96, 96, 96, params.map.tau, 0);
    auto pose_buffer = std::make_shared<ConcurrentRingBuffer<Matrix4f>>(16);
    cuda::TSDFMapping gpu(params, pose_buffer, local_map);
    featsense::MappingFeed feed(gpu);
    // a box room around the sensor, in metres, map frame
    std::vector<float> cloud;
    for (int a = -30; a <= 30; a++)
      for (int b = -10; b <= 10; b++)
      {
        const float u = a * 0.1f, v = b * 0.1f;
        const float w[4][3] = { { 3.f, u, v }, { -3.f, u, v }, { u, 3.f, v }, { u, -3.f, v } };
        for (auto &p : w) { cloud.push_back(p[0]); cloud.push_back(p[1]); cloud.push_back(p[2]); cloud.push_back(p[0] + 0.01f); cloud.push_back(p[1]); cloud.push_back(p[2]); }
      }
    const int64_t n = (int64_t)(cloud.size() / 3);
    auto pose_at = [](double x, double out[16]) { for (int i = 0; i < 16; i++) out[i] = (i % 5 == 0) ? 1.0 : 0.0; out[12] = x; };
    double P[16];
    // the voxel grid halves the doubled points
    std::vector<float> sub;
    gpu.subsample(cloud.data(), n, 12, 0.1f, sub);
    CHECK(sub.size() / 3 < (size_t)n * 3 / 4 && sub.size() / 3 > (size_t)n / 4);
    pose_at(0.0, P);
    CHECK(feed.push(cloud.data(), n, P));                                       // initialises (:66-75)
    pose_at(0.1, P);
    CHECK(!feed.push(cloud.data(), n, P));                                      // below update_distance (:81)
    // while the shift thread is inside map_shift, the feed must accumulate instead of updating
    std::atomic<int> pushed_during_shift{ 0 };
    std::atomic<bool> saw_flag{ false };
    gpu.set_shift_hook([&] {
      saw_flag = gpu.is_shifting().load();
      double Q[16];
      pose_at(1.6, Q);
      if (!feed.push(cloud.data(), n, Q)) pushed_during_shift++;                // accumulated (:115-119)
    });
    pose_at(1.3, P);                                                            // 1.3 m > map.shift (1 m): pose reaches the thread
    CHECK(feed.push(cloud.data(), n, P));
    for (int spin = 0; spin < 5000 && !gpu.shifted(); spin++) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    CHECK(gpu.shifted() && !gpu.is_shifting());
    CHECK(saw_flag && pushed_during_shift == 1 && feed.accumulated_points() == (size_t)n);
    gpu.set_shift_hook(nullptr);
    CHECK(local_map->get_pos()[0] == 13);                                       // floor(1300 / 100)
    pose_at(1.9, P);
    const int before = feed.updates();
    CHECK(feed.push(cloud.data(), n, P));                                       // flush: concatenated + subsampled (:121-126)
    CHECK(feed.updates() == before + 1 && feed.accumulated_points() == 0);
    CHECK(feed.poses().size() / 16 == 3);
    gpu.join_mapping_thread();
    gpu.get_tsdf_map();
    int seen = 0;
    for (int x = 20; x <= 40; x++) seen += local_map->in_bounds(x, 0, 0) && local_map->value(x, 0, 0).weight() != 0;
    CHECK(seen > 5);                                                            // the wall at x = 3 m is in the map
  }
  // ---- asynchronous per-scan pipeline from C++ -----------------------------------------------------------------
  {
    Params params;
    params.map.resolution = 64; params.map.tau = 600; params.map.max_weight = 10 * WEIGHT_RESOLUTION;
    params.registration.max_iterations = 10; params.registration.epsilon = 0.f;
    auto local_map = std::make_shared<HostLocalMap>(128, 128, 64, params.map.tau, 0);
    cuda::TSDFRegistration gpu(params, local_map);
    std::vector<rmagine::Pointi> scan;
    for (int a = -40; a <= 40; a++)
      for (int b = -12; b <= 12; b++) { scan.emplace_back(2500, a * 40, b * 40); scan.emplace_back(a * 40, 2200, b * 40); scan.emplace_back(-2300, a * 40, b * 40); }
    gpu.update_tsdf(scan, Matrix4f::Identity());
    const Matrix4f I = Matrix4f::Identity();
    const int t0 = gpu.track_submit(scan, &I);
    const int t1 = gpu.track_submit(scan, nullptr);                             // prior chained on the device
    Matrix4f X0{}, X1{};
    int it = 0;
    const Matrix4f p0 = gpu.track_wait(t0, &X0, &it);
    CHECK(it == 10);
    const Matrix4f p1 = gpu.track_wait(t1, &X1, &it);
    CHECK(it == 10);
    CHECK(std::fabs(p0(0, 3) - X0(0, 3)) < 1e-3f);                              // pose0 = X0 * I
    CHECK(std::fabs(p1(0, 3) - (X1(0, 0) * p0(0, 3) + X1(0, 1) * p0(1, 3) + X1(0, 2) * p0(2, 3) + X1(0, 3))) < 1e-2f);
  }
  std::printf("cpp shim ok\n");
  return 0;
}
