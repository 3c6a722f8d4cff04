/*
 * oracle/mcl_oracle.h -- C ABI of the CPU oracle for the MCL update path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the reference's
 * particle-filter update (particle_filter_cpp, /root/reference/src/particle_filter.cpp)
 * and exists to CHECK the CUDA path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (monte_carlo_localization_b200/) never links, imports or calls anything in here.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit against the
 * unmodified reference sources compiled against interface shims (oracle/_ref,
 * built by oracle/Makefile; see tests/test_oracle_vs_reference.py) and against
 * the frozen vectors under tests/golden/.
 */
#ifndef MCL_ORACLE_H_
#define MCL_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_filter orc_filter;
typedef struct orc_rng orc_rng;

/* Mirrors the ROS parameters that reach the path (src/particle_filter.cpp:23-78). */
typedef struct orc_params {
    int max_particles;
    int num_threads;              /* omp threads for the ray batch; 0 = all */
    int use_parallel_raycasting;  /* :44, :592 */
    double squash_factor;         /* :26, :53 */
    double max_range;             /* :27 */
    double z_short, z_max, z_rand, z_hit, sigma_hit;            /* :30-34 */
    double motion_dispersion_x, motion_dispersion_y, motion_dispersion_theta; /* :35-37 */
} orc_params;

/* Six TimingStats buckets (include/particle_filter_cpp/utils.hpp:49-61) + pose. */
typedef struct orc_timing {
    double total_ms, resample_ms, motion_ms, query_ms, raycast_ms, sensor_ms, pose_ms;
    int count;
} orc_timing;

void orc_default_params(orc_params* p);

orc_filter* orc_create(const orc_params* p);
void orc_destroy(orc_filter* f);

/* get_omap() :190-213 -- data is row-major int8, row 0 = bottom; resolution is the
 * float32 of OccupancyGrid.info.resolution.  Also rebuilds the sensor table (:224). */
int orc_set_map(orc_filter* f, const int8_t* data, int width, int height,
                float resolution, double ox, double oy, double oyaw);
int orc_max_range_px(const orc_filter* f);
/* sensor_model_table_ column-major (r + d*(M+1)); out has (M+1)^2 doubles. */
int orc_get_sensor_table(const orc_filter* f, double* out);
/* lidarCB :297-313 -- downsampled beam angles (float32). */
int orc_set_beam_angles(orc_filter* f, const float* angles, int n);

/* State upload / download.  particles are column-major N x 3 (x[N], y[N], theta[N]). */
int orc_set_state(orc_filter* f, const double* particles_colmajor, const double* weights);
int orc_get_state(const orc_filter* f, double* particles_colmajor, double* weights);

/* initialize_particles_pose :382-399 with injected standard normals z[3N]
 * (order x,y,theta per particle). */
int orc_init_pose(orc_filter* f, const double pose[3], const double* z3n);
/* initialize_global :401-446 with injected free-cell ordinals and headings. */
int orc_init_global(orc_filter* f, const int32_t* cell_ordinal, const double* theta);
int orc_num_free_cells(const orc_filter* f);

/* cast_ray :611-650 (one ray) and calc_range_many :586-609 (batch, column-major Q x 3). */
float orc_cast_ray(const orc_filter* f, double x, double y, double angle);
int orc_calc_range_many(orc_filter* f, const double* queries_colmajor, int64_t n, float* out);

/* One MCL() :652-694 with injected noise: u[N] are the canonical uniforms the
 * discrete_distribution would have drawn, z[3N] the standard normals the motion
 * model would have drawn (x,y,theta per particle).  idx_out (nullable) receives
 * the N resample indices; ranges are kept for orc_get_ranges. */
int orc_update(orc_filter* f, const double action[3], const float* obs, int n_obs,
               const double* u, const double* z3n, int32_t* idx_out);
/* expected_pose :696-716 */
int orc_expected_pose(orc_filter* f, double pose_out[3]);
/* ranges_ of the last update (float metres, particle-major N*R) */
int orc_get_ranges(const orc_filter* f, float* out);
/* unnormalised weights of the last update (before :679-686), for stage diffs */
int orc_get_raw_weights(const orc_filter* f, double* out);
/* mean number of grid cells the march sampled per ray in the last update
 * (C-bar; range_idx+1 on a hit, M otherwise) */
double orc_mean_cells_per_ray(const orc_filter* f);

void orc_get_timing(const orc_filter* f, orc_timing* t);
void orc_reset_timing(orc_filter* f);

/* Stand-alone pieces used by the stage-level parity tests. */
int orc_motion_model(orc_filter* f, double* particles_colmajor, const double action[3],
                     const double* z3n);
int orc_sensor_weights(orc_filter* f, const double* particles_colmajor, const float* obs,
                       int n_obs, double* weights_out);
/* discrete_distribution ctor + N lower_bound draws (random.tcc:2657-2714) */
int orc_resample_indices(const double* weights, int n, const double* u, int n_draws,
                         int32_t* idx_out, double* cdf_out /*nullable, n*/);
double orc_normalize_angle(double a); /* src/utils.cpp:43-48 */

/* Noise source: the libstdc++ objects the reference owns (particle_filter.hpp:165-167),
 * run as a twin so the injected arrays equal what the reference would draw. */
orc_rng* orc_rng_create(uint32_t seed);
void orc_rng_destroy(orc_rng* r);
void orc_rng_canonical(orc_rng* r, int64_t n, double* out);   /* generate_canonical<double,53> */
void orc_rng_normal(orc_rng* r, int64_t n, double* out);      /* normal_distribution<double>(0,1), state kept */
void orc_rng_uniform_int(orc_rng* r, int64_t n, int32_t lo, int32_t hi, int32_t* out);
void orc_rng_uniform_real(orc_rng* r, int64_t n, double lo, double hi, double* out);
/* initialize_global draws interleaved (cell, theta) per particle :433-441 */
void orc_rng_global_init(orc_rng* r, int64_t n, int32_t n_free, int32_t* cell_out, double* theta_out);
uint32_t orc_rng_raw(orc_rng* r);

#ifdef __cplusplus
}
#endif
#endif
