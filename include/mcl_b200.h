/*
 * include/mcl_b200.h -- C ABI of the B200-native MCL update (libmcl_b200.so).
 *
 * This is the drop-in boundary for ONE path of AE-HYU/monte_carlo_localization
 * (package particle_filter_cpp): ParticleFilter::MCL + expected_pose and the setup
 * state they read.  The reference has no FFI of its own -- the path is reached through
 * private members of particle_filter_cpp::ParticleFilter
 * (include/particle_filter_cpp/particle_filter.hpp:37-75) -- so every entry point below
 * names the member function or statement range it replaces.  Host code (the C++
 * ParticleFilter mirror in monte_carlo_localization_b200/host/, a ROS 2 node, ctypes)
 * binds exactly these symbols; there are no torch or C++ types in the signatures.
 *
 * Conventions
 *   - every function returns 0 on success or a negative mcl_status; nothing throws.
 *   - a context is used from one host thread at a time (the reference serialises the path
 *     with state_lock_, src/particle_filter.cpp:756).
 *   - all pointers are HOST pointers unless the name ends in _dev.
 *   - particles are column-major N x 3 doubles (x[N], y[N], theta[N]) -- the layout of the
 *     reference's Eigen::MatrixXd particles_ (particle_filter.hpp:102).
 *   - there is no CPU fallback: without a CUDA device mcl_create fails with
 *     MCL_ERR_NO_DEVICE.
 */
#ifndef MCL_B200_H_
#define MCL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCL_B200_ABI_VERSION 3

typedef enum mcl_status {
    MCL_OK = 0,
    MCL_ERR_INVALID = -1,      /* bad argument / call order */
    MCL_ERR_NO_DEVICE = -2,    /* no usable CUDA device (no CPU fallback exists) */
    MCL_ERR_CUDA = -3,         /* CUDA runtime error; see mcl_last_error */
    MCL_ERR_NO_MAP = -4,       /* reference: cast_ray returns MAX_RANGE when !map_initialized_ (:613) */
    MCL_ERR_UNSUPPORTED = -5,  /* e.g. more than 4096 beams, or a request that does not apply to the context's mode */
    MCL_ERR_NO_FREE_SPACE = -6 /* reference: "No free space found in map!" (:423-427) */
} mcl_status;

/* The ROS parameters that reach the path, same names and defaults as
 * src/particle_filter.cpp:23-47 (config/mcl_config.yaml overrides some). */
typedef struct mcl_params {
    int32_t max_particles;            /* :24   2000 */
    int32_t max_viz_particles;        /* :25   60   */
    int32_t angle_step;               /* :23   18   (host-side scan downsampling) */
    double squash_factor;             /* :26   2.2  */
    double max_range;                 /* :27   12.0 */
    double z_short, z_max, z_rand, z_hit, sigma_hit;                 /* :30-34 */
    double motion_dispersion_x, motion_dispersion_y, motion_dispersion_theta; /* :35-37 */
    uint64_t seed;                    /* device RNG seed (reference: std::random_device, :20) */
    int32_t num_filters;              /* batch of independent filters sharing map/params; 1 = the reference */
} mcl_params;

/* Injected noise for ONE update of ONE filter, in the reference's draw order:
 * u[N]  = the uniforms std::discrete_distribution would draw (:661-665),
 * z[3N] = the standard normals motion_model would draw, x,y,theta per particle (:496-498).
 * Either pointer may be NULL (device Philox stream used for that part). */
typedef struct mcl_noise {
    const double* u_resample;
    const double* z_motion;
} mcl_noise;

/* Per-stage device time of the last update in milliseconds (CUDA events; filled only
 * when profiling is enabled with mcl_set_profiling). */
typedef struct mcl_stage_ms {
    float cdf, resample_motion, raycast_weight, normalize_pose, total;
    float ray_march;   /* the ray kernel alone (raycast_weight also covers the table product) */
    float exchange;    /* unused since ABI 3 (exchanges happen inside the kernels); kept for layout */
} mcl_stage_ms;

typedef struct mcl_ctx mcl_ctx;

void mcl_default_params(mcl_params* p);
const char* mcl_last_error(void);
const char* mcl_status_str(int status);
int mcl_abi_version(void);
int mcl_device_count(void);

/* ParticleFilter ctor :19-112 (parameter read-out and buffer allocation). */
int mcl_create(const mcl_params* p, int device, mcl_ctx** out);
int mcl_destroy(mcl_ctx* ctx);

/* get_omap :190-213 + precompute_sensor_model :233-292.  data: int8 row-major, row 0 =
 * bottom, as nav_msgs/OccupancyGrid; resolution is the message's float32. */
int mcl_set_map(mcl_ctx* ctx, const int8_t* data, int width, int height, float resolution,
                double origin_x, double origin_y, double origin_yaw);
int mcl_max_range_px(const mcl_ctx* ctx);                 /* MAX_RANGE_PX :195 */
int mcl_get_sensor_table(const mcl_ctx* ctx, double* table_colmajor); /* (M+1)^2 */
/* Override the internally built table (e.g. with the reference's own bytes). */
int mcl_set_sensor_table(mcl_ctx* ctx, const double* table_colmajor, int table_width);
/* lidarCB :297-313: the downsampled beam angles (float32). */
int mcl_set_beam_angles(mcl_ctx* ctx, const float* angles, int num_beams);

/* initialize_particles_pose :382-399 (sigma 0.5/0.5/0.4, uniform weights).
 * normals_3n NULL => device RNG.  filter = -1 applies to every filter of a batch. */
int mcl_init_pose(mcl_ctx* ctx, int filter, const double pose[3], const double* normals_3n);
/* initialize_global :401-446.  cell_ordinal/theta NULL => device RNG;
 * otherwise the injected draws (index into the row-major list of free cells, heading). */
int mcl_init_global(mcl_ctx* ctx, int filter, const int32_t* cell_ordinal, const double* theta);
int mcl_num_free_cells(const mcl_ctx* ctx);
/* Upload / download state (parity harness, visualize() :944-963). */
int mcl_set_particles(mcl_ctx* ctx, int filter, const double* particles_colmajor, const double* weights);
int mcl_get_particles(mcl_ctx* ctx, int filter, double* particles_colmajor);
int mcl_get_weights(mcl_ctx* ctx, int filter, double* weights);

/* MCL(action, observation) :652-694 followed by expected_pose() :696-716, as
 * timer_update calls them (:777-778).  action = [forward, (ignored), angular] (:456-457).
 * obs = downsampled ranges (float32 metres), num_beams of them.  pose_out = [x, y, theta].
 * For a batch, action/obs/pose_out hold num_filters consecutive records and noise (if not
 * NULL) points at num_filters mcl_noise records. */
int mcl_update(mcl_ctx* ctx, const double* action, const float* obs, int num_beams,
               const mcl_noise* noise, double* pose_out);
/* Same update with inputs already resident on the device (action_dev: 3 doubles per
 * filter, obs_dev: num_beams floats per filter) and no host synchronisation: the pose is
 * left in device memory (mcl_pose_dev) and copied out by mcl_read_pose. */
int mcl_update_dev(mcl_ctx* ctx, const double* action_dev, const float* obs_dev, int num_beams);
int mcl_read_pose(mcl_ctx* ctx, double* pose_out);
/* mcl_update_dev with injected noise already on the device (u_dev: N uniforms per filter, z_dev: 3 N
 * normals per filter; a sharded rank: of the WHOLE filter; either may be NULL).  Never graph-replayed. */
int mcl_update_dev_noise(mcl_ctx* ctx, const double* action_dev, const float* obs_dev, int num_beams,
                         const double* u_dev, const double* z_dev);
int mcl_synchronize(mcl_ctx* ctx);

/* expected_pose() alone :696-716 over the current state. */
int mcl_expected_pose(mcl_ctx* ctx, int filter, double pose_out[3]);

/* calc_range_many :586-609 / cast_ray :611-650: queries column-major n x 3
 * (x[n], y[n], angle[n]); ranges in metres as float32. */
int mcl_calc_range_many(mcl_ctx* ctx, const double* queries_colmajor, int64_t n, float* ranges_out);
int mcl_cast_ray(mcl_ctx* ctx, double x, double y, double angle, float* range_out);

/* Stage read-backs of the last update (parity tests). */
int mcl_get_resample_indices(mcl_ctx* ctx, int filter, int32_t* idx_out);      /* N */
int mcl_get_ranges(mcl_ctx* ctx, int filter, float* ranges_out);               /* N*R particle-major, metres */
int mcl_get_range_steps(mcl_ctx* ctx, int filter, uint8_t* steps_out);         /* N*R step index, M = no hit */
/* The same as 16-bit values: the only form available when MAX_RANGE_PX > 254 or there are more than 128
 * beams.  Such "wide" contexts (the reference has no limit on either, :195, :307-310) march every ray with
 * the reference's own FP64 arithmetic instead of the skip-map kernels -- same results, no skipping. */
int mcl_get_range_steps16(mcl_ctx* ctx, int filter, uint16_t* steps_out);
int mcl_get_raw_weights(mcl_ctx* ctx, int filter, double* weights_out);        /* before :679-686 */
int mcl_get_cdf(mcl_ctx* ctx, int filter, double* cdf_out);                    /* discrete_distribution _M_cp */
/* visualize() :946-958: k weighted samples of the particle set (k x 3 column-major). */
int mcl_sample_particles(mcl_ctx* ctx, int filter, int k, double* particles_out);
/* The same draw with the k canonical uniforms the reference's generator would produce injected
 * (u; NULL = device RNG) and the drawn particle indices returned (idx_out, nullable): index i is
 * lower_bound(_M_cp, u[i]) exactly as std::discrete_distribution::operator() (random.tcc:2709-2713). */
int mcl_sample_particles_u(mcl_ctx* ctx, int filter, int k, const double* u, double* particles_out, int32_t* idx_out);

/* Options / introspection. */
int mcl_set_profiling(mcl_ctx* ctx, int enabled);
int mcl_get_stage_ms(mcl_ctx* ctx, mcl_stage_ms* out);
/* Device time of every kernel of the last profiled update, in launch order (CUDA events around each
 * launch).  names_out: capacity x 48 bytes, NUL-terminated; ms_out: capacity floats; *count = kernels. */
int mcl_get_kernel_ms(mcl_ctx* ctx, char* names_out, float* ms_out, int capacity, int* count);
int mcl_set_keep_ranges(mcl_ctx* ctx, int enabled);
/* Diagnostics: SM cycle counts of the phases of one kind of exact-sum pass (csrc/exact_kernels.cuh; 0 S1,
 * 1 normalise+pose+S2, 2 S2 of stored weights, 3 cdf; < 0 off) in the updates that follow.  out (nullable, 8
 * values, read and cleared): slowest CTA's tile phase | last CTA until it knows it is last | pose fold | tile
 * scan | ordered opaque list | exchange | serial evaluation + tile starts | opaque chunks of this rank.
 * pass_kind 8: k_route of a sharded filter instead -- slowest CTA's scan + serve cycles | the same incl. the system
 * fence | the last CTA's wait for the peers. */
int mcl_debug_pass_cycles(mcl_ctx* ctx, int pass_kind, unsigned long long* out);   /* store per-ray steps for read-back */
int mcl_kernel_launches(mcl_ctx* ctx, int64_t* count); /* kernels launched so far by this ctx */
/* Use the caller's CUDA stream (cudaStream_t passed as void*) instead of the ctx's own. */
int mcl_set_stream(mcl_ctx* ctx, void* cuda_stream);

/* mcl_update and mcl_update_dev replay their steady state (no injected noise, no diagnostics, whole
 * filter on this GPU, weights untouched since the previous update) as ONE CUDA graph per
 * state-buffer parity instead of ~18 kernel launches; results are identical (mcl_update_dev first
 * copies action and scan into the context's staging buffers, device to device).  On by default;
 * graphs are dropped and re-captured whenever a setter changes a buffer, the stream or a mode. */
int mcl_set_graphs(mcl_ctx* ctx, int enabled);

/* Programmatic dependent launches between the kernels of an update (on by default): a kernel's blocks are
 * scheduled while its predecessor drains and wait on griddepcontrol.wait, which hides launch latency at every
 * kernel boundary; results are identical.  Off: plain stream order (for comparison). */
int mcl_set_pdl(mcl_ctx* ctx, int enabled);

/* Ray stage selection.  One filter of at least 1024 particles gets, besides the isotropic
 * skip-map kernel, the DIRECTIONAL stage: per-heading-sector skip maps built at mcl_set_map, rays
 * grouped by sector, each sector's window staged in shared memory.  Both compute the reference's
 * cast_ray (src/particle_filter.cpp:611-650) exactly; they differ only in how many samples they
 * can prove irrelevant.  mode 0 (default): directional whenever the context is eligible -- particles
 * inside the window box around the cloud centre march shared memory, the others (all of them right
 * after mcl_init_global) the same sector maps in L2; 1: isotropic kernel only; 2: directional
 * always (MCL_ERR_UNSUPPORTED if the context is not eligible).  A BATCH of
 * filters whose whole padded map fits one window runs the directional stage over the pool of all
 * filters' particles (2.67 against 3.28 ms per step of 1024 x 4000 particles on sibal1); other batches and
 * filters of fewer than 1024 particles use the isotropic kernel. */
int mcl_set_ray_mode(mcl_ctx* ctx, int mode);
/* directional_ready: the context is eligible and its sector maps are built; last_mode: 1 if the
 * last update ran the directional stage; box_cells: side of the window box; units: work units of
 * the last update.  Any pointer may be NULL. */
int mcl_ray_stage_info(mcl_ctx* ctx, int* directional_ready, int* last_mode, int* box_cells, int* units);

/* Diagnostics: one sector's directional skip map (padded grid, *pw x *ph bytes, see
 * csrc/dirmap.cuh for the code); sector -1: the isotropic skip codes (csrc/map_prep.h), as built on the
 * device by mcl_set_map.  out may be NULL to query the size. */
int mcl_get_dir_map(mcl_ctx* ctx, int sector, uint8_t* out, int* pw, int* ph);

/* Gather micro-benchmark (measurement aid): random single-byte reads per second from
 * an L2-resident array of array_bytes (shared = 0) or from a shared-memory window (shared = 1,
 * capped at 128 KB) -- the access pattern of the ray march without its arithmetic.  It gives the
 * ray stage a gather-rate roofline next to the HBM one. */
int mcl_microbench_gather(int device, int shared, size_t array_bytes, int iters_per_thread, double* gathers_per_second);

/* ---- particle-sharded filter: one rank per GPU, ONE global filter ---------------------------
 * The reference has no multi-process path; the coupling points of its update are the weight sum
 * (:679), the CDF of std::discrete_distribution + the source gather (:658-665) and the pose sums
 * (:702-710).  A sharded context holds ONLY its own slot range [rank * n, (rank + 1) * n) of every
 * per-particle array (n = max_particles / world; mcl_params.max_particles is the particle count of
 * the WHOLE filter) and keeps the reference's exact global multinomial resampling:
 *   - the three sequentially rounded reductions (weight sum, the distribution's own sum, the CDF) run
 *     per rank; each exchanges one < 2 KB summary per rank (step maps of csrc/exact_sum.cuh + the few
 *     chunks that cross a binade), after which every rank evaluates the same short serial chain;
 *   - resampling is sender-driven: every rank evaluates all draws, serves those that fall into its own
 *     CDF range and PUSHES the source poses to the slots' owners over NVLink (k_route);
 *   - no rank reads peer memory on the hot path; every exchange is stores + a system-scope release
 *     flag written by the kernels themselves (csrc/shard.cuh), so an update has no host call between
 *     its launches and replays as one CUDA graph.
 * With the same injected noise the ranks' slices equal the single-filter update bit for bit.
 * mcl_update / mcl_update_dev / mcl_read_pose / mcl_init_* / mcl_set_particles / mcl_get_* work on a
 * sharded context and address the rank's own slice; injected noise (mcl_noise) covers the WHOLE
 * filter and is indexed by the global slot.  Every rank must issue the same sequence of updates. */

/* Creation without NCCL (the caller moves the 512-byte blobs): mcl_shard_create on every rank ->
 * mcl_shard_export -> exchange -> mcl_shard_connect(blobs of all ranks in rank order).  Ranks that
 * live in ONE process (one host thread per GPU, or test ranks emulated on one GPU) connect with
 * mcl_shard_connect_local instead. */
#define MCL_SHARD_BLOB_BYTES 512
int mcl_shard_create(const mcl_params* p, int device, int world, int rank, mcl_ctx** out);
int mcl_shard_export(mcl_ctx* ctx, void* blob, size_t capacity);
int mcl_shard_connect(mcl_ctx* ctx, const void* blobs /* world x MCL_SHARD_BLOB_BYTES */);
int mcl_shard_connect_local(mcl_ctx* ctx, mcl_ctx* const* ranks /* world contexts, rank order */);
int mcl_shard_info(const mcl_ctx* ctx, int* world, int* rank, int64_t* n_local, int64_t* n_global);

/* How the ranks meet at an exchange.  fused = 1 (default): the publishing kernel's last block waits for
 * the peers' flags itself -- one rank per GPU only.  fused = 0: host-ordered; the library calls `hook`
 * (after synchronising its stream) wherever every rank must have published before any rank consumes --
 * for ranks emulated on one GPU, whose kernels must never spin on each other -- or, with hook == NULL on
 * a context made by mcl_create_sharded, enqueues a one-word ncclAllGather as the barrier. */
typedef int (*mcl_barrier_fn)(void* user);
int mcl_shard_set_exchange(mcl_ctx* ctx, int fused, mcl_barrier_fn hook, void* user);

/* How the draws of the global multinomial resampling (src/particle_filter.cpp:658-665) reach the rank that
 * owns their source particle.  two_hop = 1 (the default from 3 ranks on): the owner of a slot classifies its own N draws against
 * the ranks' CDF ranges and appends a 4-byte request to the source rank's inbox; after one exchange of the
 * counts the source ranks search and push the poses (work per rank independent of the number of ranks).
 * two_hop = 0 (the default for 2 ranks): every rank evaluates all world * N draws and serves the ones in its own
 * range (one exchange less, work per rank grows with the world size).  two_hop < 0 restores the default.  Both
 * draw the same particles bit for bit; every rank of a filter must use the same setting. */
int mcl_shard_set_route(mcl_ctx* ctx, int two_hop);

/* The same with the library owning the NCCL communicator (bound at run time from libnccl.so.2):
 * rank 0 obtains an id (mcl_nccl_unique_id, 128 bytes) and hands it to the other ranks by any means;
 * mcl_create_sharded = mcl_shard_create + ncclCommInitRank + ncclAllGather of the blobs on the
 * library's stream + mcl_shard_connect.  mcl_sharded_gather all-gathers the whole filter's particles
 * (NG x 3 column-major) and normalised weights to the host of every rank (either may be NULL). */
int mcl_nccl_unique_id(void* id_out, size_t capacity);
int mcl_create_sharded(const mcl_params* p, int device, int world, int rank, const void* nccl_unique_id, mcl_ctx** out);
int mcl_sharded_gather(mcl_ctx* ctx, double* particles_colmajor, double* weights);

#ifdef __cplusplus
}
#endif
#endif /* MCL_B200_H_ */
