/*
 * ws_oracle.h -- CPU oracle for the warpsense TSDF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free C restatement of the
 * reference's CPU path (src/cpu).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (warpsense_b200/) never links, imports or calls anything in this directory.
 *
 * Parity pinning: the restatement is checked against every known-answer test the
 * reference holds for this path (G1 test/map.cpp:9-90, G2 test/map.cpp:240-365,
 * G3 test/cuda.cpp:416-532, G4 test/cuda.cpp:760-827, G5 test/params_a.cpp:26-41,
 * J*J^T test/cuda.cpp:837-923).  register_cloud's numeric output and multi-point
 * update_tsdf results are NOT pinned by any assertion in the reference:
 * for those, parity = this restatement vs the CUDA path.
 *
 * All citations are file:line under /root/reference.
 */
#ifndef WS_ORACLE_H
#define WS_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/warpsense/consts.h:9-13 */
#define ORC_WEIGHT_RESOLUTION 64
#define ORC_MATRIX_RESOLUTION 32768
/* include/map/hdf5_global_map.h:78 */
#define ORC_CHUNK_SIZE 64

typedef struct { int32_t x, y, z; } orc_point;

/* include/map/tsdf.h:16-35 : {int16 value; int16 weight} unioned with uint32 raw.
 * raw = (uint16)value | (uint16)weight << 16 on a little-endian host. */
typedef uint32_t orc_entry;

typedef struct orc_map orc_map;

typedef struct {
  int64_t n_points;        /* points given                                         */
  int64_t n_marched;       /* points that passed the distance/own-cell/iv guards    */
  int64_t n_candidates;    /* C: try_emplace calls (update_tsdf.cpp:508)            */
  int64_t n_touched;       /* T: distinct voxels in the per-scan map                */
  int64_t n_written;       /* voxels whose grid entry was actually rewritten        */
  int64_t n_neg_candidates;/* candidates with negative (interpolated) weight        */
} orc_update_stats;

/* ---- entry helpers (include/map/tsdf.h) ---- */
orc_entry orc_make_entry(int value, int weight);
int orc_entry_value(orc_entry e);
int orc_entry_weight(orc_entry e);

/* ---- local map + in-memory global chunk store
 *      (include/map/hdf5_local_map.h, src/map/hdf5_local_map.cpp, src/map/hdf5_global_map.cpp) ---- */
orc_map *orc_map_create(int sx, int sy, int sz, int default_value, int default_weight);
void orc_map_destroy(orc_map *m);
orc_map *orc_map_clone(const orc_map *m);           /* hdf5_local_map.cpp:22-31 (shares the chunk store) */
void orc_map_get_size(const orc_map *m, int out[3]);
void orc_map_get_pos(const orc_map *m, int out[3]);
void orc_map_get_offset(const orc_map *m, int out[3]);
orc_entry *orc_map_data(orc_map *m);                /* size[0]*size[1]*size[2] raw entries, ring layout */
int64_t orc_map_num_voxels(const orc_map *m);
int orc_map_in_bounds(const orc_map *m, int x, int y, int z);          /* hdf5_local_map.h:275-279 */
int64_t orc_map_index(const orc_map *m, int x, int y, int z);          /* hdf5_local_map.h:140-151 */
int orc_map_get(const orc_map *m, int x, int y, int z, orc_entry *out);/* 0 ok, -1 out of range    */
int orc_map_set(orc_map *m, int x, int y, int z, orc_entry e);
void orc_map_set_state(orc_map *m, const int pos[3], const int offset[3]);
int orc_map_shift(orc_map *m, const int new_pos[3]);                   /* hdf5_local_map.cpp:53-118 */
void orc_map_write_back(orc_map *m);                                   /* hdf5_local_map.cpp:210-217 */
/* global chunk store inspection (hdf5_global_map.cpp:53-57,139-153) */
int64_t orc_store_num_chunks(const orc_map *m);
int orc_store_chunk_list(const orc_map *m, int *out_xyz, int64_t cap);
const orc_entry *orc_store_chunk(const orc_map *m, int cx, int cy, int cz); /* NULL if absent */
orc_entry orc_store_get_value(orc_map *m, int x, int y, int z);        /* hdf5_global_map.cpp:139-145 */

/* ---- fixed-point helpers (include/util/util.h) ---- */
void orc_to_int_mat(const float m_colmajor[16], int32_t out_colmajor[16]);             /* util.h:8-11   */
orc_point orc_transform_point(orc_point p, const int32_t mat_colmajor[16]);            /* util.h:13-18  */
int64_t orc_preprocess(const float *xyz, int64_t n, int stride_floats, const float pose_colmajor[16],
                       int map_resolution, orc_point *out);                                 /* app.cpp:118-148 */
int64_t orc_preprocess_cpu_node(const float *xyz, int64_t n, int stride_floats, const float pose_colmajor[16],
                                int map_resolution, orc_point *out);                        /* fastsense.cpp:143-163 */
void orc_transform_points(const orc_point *in, int64_t n, const int32_t mat_colmajor[16], orc_point *out);
void orc_to_map(const float pose_colmajor[16], int map_resolution, int out[3]);        /* util.h:52-56  */
void orc_convert_pose(const float pose_colmajor[16], int map_resolution,
                      int pos[3], int up[3]);                   /* tsdf_mapping.cpp:77-85 */
void orc_transform_point_cloud(orc_point *pts, int64_t n, const float m_colmajor[16]); /* util.h:66-80 */

/* ---- parameter scaling (include/params/map_params.h:88-106) ---- */
void orc_scale_params(float max_distance, int max_weight_in, const float size_m[3], int resolution,
                      int *tau, int *max_weight, int size_vox[3]);
/* include/warpsense/test/common.h:16-26 */
int orc_calc_weight(int value, int tau, int weight_epsilon);

/* ---- the update oracle: src/cpu/update_tsdf.cpp:397-564 ---- */
void orc_update_tsdf(orc_map *m, const orc_point *pts, int64_t n,
                     const int scanner_pos[3], const int up[3],
                     int tau, int max_weight, int map_resolution,
                     orc_update_stats *stats /* may be NULL */);
/* src/cpu/update_tsdf.cpp:566-724: the OpenMP overload (per-thread candidate maps + cross-thread rule); equals
 * orc_update_tsdf with threads = 1 and every scan point inside the map; for timing.  threads < 1: all cores */
void orc_update_tsdf_omp(orc_map *m, const orc_point *pts, int64_t n,
                         const int scanner_pos[3], const int up[3],
                         int tau, int max_weight, int map_resolution, int threads,
                         orc_update_stats *stats /* may be NULL */);

/* ---- the registration oracle: src/cpu/registration.cpp:14-177 ---- */
/* one Gauss-Newton accumulation pass (registration.cpp:52-118): H column-major 6x6 */
void orc_reg_step(const orc_map *m, const orc_point *pts, int64_t n,
                  const float T_colmajor[16], int map_resolution,
                  int64_t H[36], int64_t g[6], int32_t *err, int32_t *cnt);
/* FP64 damped solve + Rodrigues update (registration.cpp:128-157, registration/util.h:5-39).
 * T is updated in place; returns err/count as float. */
float orc_reg_solve(const int64_t H[36], const int64_t g[6], int32_t err, int32_t cnt,
                    float alpha, float T_colmajor[16], double xi_out[6]);
void orc_xi_to_transform(const double xi[6], const int center[3], float out_colmajor[16]);
/* full loop; cloud is transformed in place (registration.cpp:168-174); returns iterations run.
 * trace (optional): per iteration 29 int64 = 21 upper-tri H (row-major i<=j), 6 g, err, cnt. */
int orc_register_cloud(const orc_map *m, orc_point *cloud, int64_t n,
                       const float pretransform_colmajor[16],
                       int max_iterations, float it_weight_gradient, float epsilon,
                       int map_resolution, float out_colmajor[16],
                       int64_t *trace, int trace_cap);

/* test/cuda.cpp:837-923 : H += J J^T for one Jacobian */
void orc_jacobi_2_h(const int64_t J[6], int64_t H[36]);

/* ---- featsense feed: pcl::VoxelGrid<PointXYZI> + metres -> millimetres (src/warpsense/tsdf_mapping.cpp:145-159).
 * PCL is not vendored in the reference tree (ROS noetic: PCL 1.10); this restates filters/impl/voxel_grid.hpp
 * applyFilter (leaf index = floor(x * inverse_leaf) - min_b, output sorted by ijk0 + ijk1*div0 + ijk2*div0*div1,
 * float centroid sums, `xyz / n`).  PARITY UNPINNED: no PCL here to run; runs of equal leaf index are summed in
 * ascending point index (std::sort leaves that order unspecified).  Returns the number of output points. */
int64_t orc_voxelgrid(const float *xyz, int64_t n, int stride_floats, float leaf, float *out_xyz, orc_point *out_mm);
/* Eigen::Quaterniond(pose.rotation()).toRotationMatrix().cast<float>(), translation * 1000 (tsdf_mapping.cpp:160-161) */
void orc_mm_pose_from_isometry(const double pose_colmajor[16], float out_colmajor[16]);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
