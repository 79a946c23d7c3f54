// ref_cuda_shim.cu -- C entry points over the REFERENCE's own CUDA classes, for timing them on the B200.
//
// TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/).  This file contains no reference
// code: it includes the reference headers and is linked with the reference's .cu files where they lie
// under /root/reference (recipe: oracle/Makefile target _ref/libws_refcuda.so).  The resulting library
// is a secondary, same-hardware baseline ("the reference's Jetson-era kernels recompiled for sm_100")
// -- NOT the parity target: it differs observably from the CPU path and its registration
// kernel only looks at the first 65,536 points (src/warpsense/cuda/registration.cu:353).
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "warpsense/cuda/update_tsdf.h"
#include "warpsense/cuda/registration.h"

namespace {
struct RefCuda
{
  int size[3], offset[3], pos[3];
  std::vector<TSDFEntry> data;
  cuda::DeviceMap *view = nullptr;
  cuda::TSDFCuda *tsdf = nullptr;
  cuda::RegistrationCuda *reg = nullptr;
  std::vector<rmagine::Pointi> pts;
  int res = 0;
};
}  // namespace

extern "C" {

void *refcuda_create(int sx, int sy, int sz, int tau, int max_weight, int res)
{
  RefCuda *r = new RefCuda();
  r->size[0] = sx; r->size[1] = sy; r->size[2] = sz;
  for (int a = 0; a < 3; a++) { r->offset[a] = r->size[a] / 2; r->pos[a] = 0; }
  r->data.assign((size_t)sx * sy * sz, TSDFEntry(tau, 0));
  r->view = new cuda::DeviceMap(r->size, r->offset, r->data.data(), r->pos);
  r->tsdf = new cuda::TSDFCuda(*r->view, tau, max_weight, res);
  r->reg = new cuda::RegistrationCuda(*r->view);
  r->res = res;
  return r;
}

void refcuda_destroy(void *p)
{
  RefCuda *r = (RefCuda *)p;
  if (!r) return;
  delete r->reg; delete r->tsdf; delete r->view; delete r;
}

void refcuda_set_points(void *p, const int32_t *xyz, int64_t n)
{
  RefCuda *r = (RefCuda *)p;
  r->pts.resize((size_t)n);
  std::memcpy((void *)r->pts.data(), xyz, (size_t)n * 12);
}

// cuda::TSDFCuda::update_tsdf(points, scanner_pos, up): H2D of the scan + both kernels, blocking
void refcuda_update(void *p, const int32_t pos[3], const int32_t up[3])
{
  RefCuda *r = (RefCuda *)p;
  rmagine::Pointi sp(pos[0], pos[1], pos[2]), u(up[0], up[1], up[2]);
  r->tsdf->update_tsdf(r->pts, sp, u);
  cudaDeviceSynchronize();
}

void refcuda_reg_prepare(void *p)
{
  RefCuda *r = (RefCuda *)p;
  r->reg->prepare_registration(r->pts);
}

// cuda::RegistrationCuda::perform_registration: one Gauss-Newton accumulation (H2D 4x4, two kernels,
// D2H of 32 partials, host reduce).  T column-major; H returned column-major.
void refcuda_reg_step(void *p, const float T[16], int64_t H[36], int64_t g[6], int32_t *e, int32_t *c)
{
  RefCuda *r = (RefCuda *)p;
  rmagine::Matrix4x4f M;
  std::memcpy((void *)&M, T, sizeof(float) * 16);
  rmagine::Matrix6x6l h;
  rmagine::Point6l gg;
  int ee = 0, cc = 0;
  r->reg->perform_registration(r->tsdf->device_map(), &M, h, gg, ee, cc, r->res);
  std::memcpy(H, (void *)&h, sizeof(int64_t) * 36);
  std::memcpy(g, (void *)&gg, sizeof(int64_t) * 6);
  *e = ee; *c = cc;
}

void refcuda_download(void *p, uint32_t *out)
{
  RefCuda *r = (RefCuda *)p;
  r->tsdf->avg_map().to_host(*r->view);
  std::memcpy(out, (void *)r->data.data(), r->data.size() * sizeof(uint32_t));
}

}  // extern "C"
