/*
 * rbphd.h -- C ABI of librbphd.so: the B200 (sm_100a) engine behind MonoRFS's PHDNavigator.
 *
 * This is the drop-in boundary.  A C# class `GpuPHDNavigator : Navigator<...>` with
 * PHDNavigator's public surface P/Invokes these entry points; the
 * convention is the reference's own P/Invoke precedent for libisam2.so
 * (mono-rfs-lib/SLAM/Navigators/ISAM2Navigator.cs:597-622, isam2/isam2.cpp:46,145-365):
 *   - opaque handle from a `new` call, released by a `delete` call (ISAM2Navigator.cs:446-452);
 *   - inputs are pinned managed double[] / int[] passed as pointers, flat row-major packing
 *     (poses 7 doubles x,y,z,qw,qx,qy,qz; points 3; covariances 9), copied before return;
 *   - int status return: 0 ok, >0 a known failure class, -1 generic; no exception crosses
 *     the boundary (isam2.cpp:317-335); rbphd_last_error() gives the text;
 *   - outputs are library-owned host buffers returned as pointer + length, valid until the
 *     next mutating call on the same handle (isam2.cpp:339-365).
 * No torch / CUDA types appear in any signature.  Handles are independent (own device buffers
 * and stream); calls on ONE handle must come from one thread at a time, calls on different
 * handles may run concurrently (LoopyPHDNavigator.cs:525-551 creates one navigator per task).
 *
 * Each entry point cites the reference member it replaces (file:line under /root/reference,
 * PHD = mono-rfs-lib/SLAM/Navigators/PHDNavigator.cs).
 */
#ifndef RBPHD_H
#define RBPHD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBPHD_OK               0
#define RBPHD_ERR_GENERIC     (-1)
#define RBPHD_ERR_CUDA          1   /* a CUDA runtime call failed */
#define RBPHD_ERR_CAPACITY      2   /* a per-particle capacity (components, pairs, edges, blocks) was exceeded */
#define RBPHD_ERR_ARGUMENT      3   /* bad argument (index out of range, M > max_measurements, ...) */
#define RBPHD_ERR_NO_DEVICE     4   /* no CUDA device / kernels not built for it: there is NO CPU fallback */

/* Navigator parameters = the Config statics the reference's PHDNavigator reads live
 * (PHD:56-113, mono-rfs-lib/Config.cs:46-91,238-263) plus the measurer (PRM3DMeasurer.cs:70-114). */
typedef struct rbphd_config {
    int32_t model;          /* 0 = PRM3D (Pose3D + PixelRangeMeasurement); the only model on the GPU */
    int32_t max_quantity;   /* Config.MaxQuantity */
    int32_t gate_metric;    /* KD-tree radius semantics: 0 |d|^2 <= r^2, 1 |d|^2 <= r */
    int32_t nthreads;       /* Config.NParallel (ignored on the GPU) */
    double  R[9];           /* vehicle.MeasurementCovariance * MeasurementCovarianceMultiplier */
    double  Q[36];          /* vehicle.MotionCovariance * MotionCovarianceMultiplier */
    double  pd;             /* Config.NavigatorPD */
    double  clutter;        /* Config.NavigatorClutterDensity */
    double  birth_cov[9];   /* Config.BirthCovariance */
    double  birth_weight;   /* Config.BirthWeight */
    double  min_weight;     /* Config.MinWeight */
    double  merge_threshold;            /* Config.MergeThreshold */
    double  exploration_threshold;      /* Config.ExplorationThreshold */
    double  density_distance_threshold; /* Config.DensityDistanceThreshold */
    double  min_effective_particle;     /* Config.MinEffectiveParticle */
    double  visibility_ramp[3];         /* Config.VisibilityRamp */
    double  measurer[7];    /* PRM3DMeasurer.ToLinear(): focal, rangemin, rangemax, filmX, filmY, filmW, filmH */
} rbphd_config;

/* Sizing of the device buffers (not part of the reference's semantics). 0 = derive a default. */
typedef struct rbphd_limits {
    int32_t device;              /* CUDA device ordinal */
    int32_t max_particles;       /* capacity in particles (>= every particle count used later) */
    int32_t max_components;      /* per-particle map capacity; default max(max_quantity, 64) rounded up */
    int32_t max_measurements;    /* per-frame measurement capacity; default 1024 */
    int32_t max_pairs;           /* gated (component, measurement) pairs per particle; default 4*max_measurements */
    int32_t resident_frames;     /* device slots for per-frame inputs (gauss, z); default 1 */
    int32_t reserved[2];
} rbphd_limits;

typedef struct rbphd_navigator rbphd_navigator;

/* ---- lifecycle: PHDNavigator ctor / Dispose (PHD:192-208) ---- */
rbphd_navigator* rbphd_new(const rbphd_config* config, const rbphd_limits* limits);
void             rbphd_delete(rbphd_navigator* nav);
const char*      rbphd_last_error(const rbphd_navigator* nav);   /* nav may be NULL: last creation error */

/* reset / CollapseParticles / StartSlam / StartMapping (PHD:214-266): all particles <- (pose, map),
 * weights 1/P, BestParticle 0.  Map is AoS: w[n], mean[n*3], cov[n*9]. */
int rbphd_reset(rbphd_navigator* nav, int particles, const double* pose7, int n, const double* w,
                const double* mean, const double* cov);
/* ResetMapModel (PHD:271-276) */
int rbphd_clear_maps(rbphd_navigator* nav);
int rbphd_particle_count(const rbphd_navigator* nav);

/* ---- Update (PHD:295-314 -> TrackVehicle.UpdateNoisy TRK:89-102 -> Pose3D.AddOdometry POSE:314-333).
 * gauss = particles x 6 N(0,1) draws made by the host in the reference's order (Util.Gaussian,
 * UTIL:198), so the random stream stays the C# one.  perfect_still = Config.PerfectStill. */
int rbphd_update(rbphd_navigator* nav, const double* reading6, double dt, const double* gauss,
                 int perfect_still);
/* direct pose writes: OnlyMapping Update (PHD:297-300), LoopyPHDNavigator.cs:741,756, tests */
int rbphd_set_pose(rbphd_navigator* nav, int particle, const double* pose7);
int rbphd_set_poses(rbphd_navigator* nav, const double* poses);   /* particles x 7 */
int rbphd_get_poses(rbphd_navigator* nav, const double** poses, int* particles);

/* ---- SlamUpdate (PHD:323-362): PredictConditional + CorrectConditional + PruneModel + WeightAlpha
 * per particle, then weight normalisation, BestParticle, ParticleDepleted and ResampleParticles.
 * z = m x 3 (px, py, range).  u_resample = the value Util.Uniform.Next() returns at PHD:727 (drawn by
 * the host only when resampled comes back 1 in the reference; pass one per call here).
 * best / resampled may be NULL. */
int rbphd_slam_update(rbphd_navigator* nav, const double* z, int m, int only_mapping, double u_resample,
                      int* best, int* resampled);
/* SlamUpdate in two steps, for hosts that must consume their random stream exactly as the reference does:
 * PHD:355-357 calls ResampleParticles() -- and with it Util.Uniform.Next() at PHD:727 -- only when
 * ParticleDepleted() is true.  _begin runs the frame up to that decision and reports it; when *depleted comes back
 * 1 the host draws the uniform and calls _finish(u), which runs the wheel and the particle copy (PHD:724-760).
 * When *depleted is 0 the frame is complete and _finish is a no-op.  Same results as rbphd_slam_update. */
int rbphd_slam_update_begin(rbphd_navigator* nav, const double* z, int m, int only_mapping, int* best,
                            int* depleted);
int rbphd_slam_update_finish(rbphd_navigator* nav, double u_resample, int* best);
/* KinectMeasurer (BaseStructures/Measurers/KinectMeasurer.cs:123-173, KinectTrackVehicle.cs:61-74): attach the
 * current depth frame, depth_xy[x * resy + y] in metres (NaN = no reading), so that the detection probability of a
 * landmark also ramps to 0 as it moves behind the measured surface (occlusion).  The configured film rectangle must
 * already be the border-deflated one the reference gives its KinectMeasurer.  The frame is copied; pass NULL to
 * return to the plain pixel-range measurer.  Upload a new frame before each SlamUpdate (1.2 MB at 640 x 480). */
int rbphd_set_depth_frame(rbphd_navigator* nav, const float* depth_xy, int resx, int resy);
/* Leave-one-out batches (LoopyPHDNavigator.FilterMissing, LoopyPHDNavigator.cs:729-763, called for every frame
 * index by UpdateMessagesFromMap, :511-552): the T one-particle, mapping-only re-filters of a smoothing pass differ
 * only in which frame they skip, so they run as T "particles" of one navigator -- every frame rbphd_set_poses gives
 * all of them that frame's trajectory pose, rbphd_set_holdout names the one filter that skips it (its map is
 * carried over unchanged), and a mapping-only frame updates the rest.  particle < 0 clears the hold-out. */
int rbphd_set_holdout(rbphd_navigator* nav, int particle);
/* Device-resident frame loop: Update + SlamUpdate enqueued on the handle's stream with nothing copied
 * back and no host synchronisation.  The frame's inputs (gauss: particles x 6, z: m x 3) are uploaded
 * beforehand into input slot `slot` (0 <= slot < resident_frames); either pointer may be NULL to keep
 * what the slot holds.  rbphd_synchronize() waits and reports deferred capacity errors. */
int rbphd_upload_frame_inputs(rbphd_navigator* nav, int slot, const double* gauss, const double* z, int m);
int rbphd_frame_async(rbphd_navigator* nav, int slot, const double* reading6, double dt, int perfect_still,
                      int m, int only_mapping, double u_resample);
/* Update alone, enqueued without synchronisation (gauss from input slot `slot`) */
int rbphd_update_async(rbphd_navigator* nav, int slot, const double* reading6, double dt, int perfect_still);
int rbphd_synchronize(rbphd_navigator* nav);

/* ResampleParticles (PHD:724-760) / ParticleDepleted (PHD:768-777) as public methods */
int rbphd_resample(rbphd_navigator* nav, double u_resample);
int rbphd_particle_depleted(rbphd_navigator* nav, int* depleted);

/* ---- state read-back / write: VehicleWeights, BestParticle, MapModels[i] (PHD:123-161) ---- */
int rbphd_get_weights(rbphd_navigator* nav, const double** weights, int* particles);
int rbphd_set_weights(rbphd_navigator* nav, const double* weights);
int rbphd_get_alphas(rbphd_navigator* nav, const double** alphas, int* particles);   /* last WeightAlpha values */
int rbphd_get_best(rbphd_navigator* nav, int* best);
int rbphd_get_ancestors(rbphd_navigator* nav, const int** ancestors, int* particles);  /* of the last SlamUpdate */
int rbphd_get_map_counts(rbphd_navigator* nav, const int** counts, int* particles);
int rbphd_get_map(rbphd_navigator* nav, int particle, const double** w, const double** mean,
                  const double** cov, int* n);
int rbphd_set_map(rbphd_navigator* nav, int particle, int n, const double* w, const double* mean,
                  const double* cov);

/* ---- the per-particle public methods, on caller-supplied single maps, so the reference's unit
 * tests (Test/PHDNavigatorTest.cs) run against the GPU path.  Outputs are library-owned. ---- */
/* PredictConditional (PHD:793-819) */
int rbphd_stage_predict(rbphd_navigator* nav, const double* pose7, int n, const double* w, const double* mean,
                        const double* cov, const double* z, int m, const double** ow, const double** omean,
                        const double** ocov, int* on);
/* CorrectConditional (PHD:829-906); input = predicted map; gate_radius < 0 = ungated */
int rbphd_stage_correct(rbphd_navigator* nav, const double* pose7, int n, const double* w, const double* mean,
                        const double* cov, const double* z, int m, double gate_radius, const double** ow,
                        const double** omean, const double** ocov, int* on);
/* PruneModel (PHD:913-948) */
int rbphd_stage_prune(rbphd_navigator* nav, int n, const double* w, const double* mean, const double* cov,
                      const double** ow, const double** omean, const double** ocov, int* on);
/* WeightAlpha (PHD:373-393); out7 = alpha, setloglik, ploglik, cloglik, pcount, ccount, J */
int rbphd_stage_weight_alpha(rbphd_navigator* nav, const double* pose7, const double* z, int m, int np,
                             const double* pw, const double* pmean, const double* pcov, int nc,
                             const double* cw, const double* cmean, const double* ccov, double* out7);
/* SetLogLikelihood (PHD:462-515) on an explicit landmark list (the IMap of Gaussian(mean, I, 1)) */
int rbphd_stage_set_loglikelihood(rbphd_navigator* nav, const double* pose7, int j, const double* jmean,
                                  const double* z, int m, double* loglik);
/* The static likelihood functions of PHDNavigator over a landmark list (jmean: j x 3) seen from pose7:
 * SetLikelihood (PHD:402-406) = exp(SetLogLikelihood); QuasiSetLogLikelihood (PHD:526-532, 561-713 without the
 * gradient: full visibility PD_i = PD, association gate d < 12) -- what LoopyPHDNavigator evaluates hundreds of
 * times per smoothing pass (LoopyPHDNavigator.cs:810-815, 891-897, 948); SetLogLikeMatrix (PHD:415-460) as
 * (row, column, value) triplets sorted by (row, column): rows/columns 0..j-1 landmarks / measurements-then-misses
 * exactly as the reference indexes its SparseMatrix (library-owned buffers, valid until the next call). */
int rbphd_set_likelihood(rbphd_navigator* nav, const double* pose7, int j, const double* jmean, const double* z,
                         int m, double* likelihood);
int rbphd_quasi_set_loglikelihood(rbphd_navigator* nav, const double* pose7, int j, const double* jmean,
                                  const double* z, int m, double* loglik);
int rbphd_set_loglike_matrix(rbphd_navigator* nav, const double* pose7, int j, const double* jmean, const double* z,
                             int m, const int** rows, const int** cols, const double** vals, int* nnz);
/* QuasiSetLogLikelihood(measurements, map, pose, out gradient) (PHD:544-549, 561-713): the value as this overload
 * computes it and the pose gradient (6 entries, MeasurementJacobianP's parametrisation, PRM:185-209) -- what
 * LogLikeGradientAscent / LogLikeFitCovariance evaluate (LoopyPHDNavigator.cs:928-934, 989-994).  The reference's
 * TemperedAverage (Util/MatrixExtensions.cs:400-440) overwrites the shared 200-entry buffer in place and normalises
 * it with Accord's Normalize(): sum_normalised = 0 follows that literally (Euclidean norm of the whole buffer);
 * sum_normalised = 1 divides by the sum instead, the variant the reference's own LogLike2D test
 * (LoopyPHDNavigatorTest.cs:352-425) passes (oracle/README.md D10). */
int rbphd_quasi_set_loglikelihood_gradient(rbphd_navigator* nav, const double* pose7, int j, const double* jmean,
                                           const double* z, int m, int sum_normalised, double* loglik,
                                           double* gradient6);

/* ---- multi-GPU: particles sharded by rank, collectives inside the library.
 * One navigator per GPU / rank; rank g owns the block [g*P/G, (g+1)*P/G) of the P global particles.  NCCL is
 * bound at run time (libnccl.so.2), so a single-GPU host does not need it.  After rbphd_comm_init_rank the frame
 * entry points (rbphd_frame_async, rbphd_slam_update) run SlamUpdate's coupled tail (PHD:343-358) over all ranks:
 * one ncclAllGather of the un-normalised weights, the identical serial normalise / ESS / wheel on every rank (so all
 * ranks hold the same ancestors), one small read-back of the decision, and -- only on resampling frames -- an
 * allgather of the component counts plus grouped ncclSend / ncclRecv of the ancestors' (pose, map) records,
 * 8 + 13 n doubles each, every remote ancestor once per destination rank.  Particle indices reported by these
 * calls (best, ancestors) are GLOBAL.  Replaces the Parallel.For / shared-memory coupling of PHD:326-358. ---- */
/* a fresh NCCL unique id (call on one rank, hand the 128 bytes to the others by any out-of-band means) */
int rbphd_comm_unique_id(unsigned char out128[128]);
/* join the communicator; the navigator must have been reset with this rank's block of `total_particles`.
 * Sets the local weights to 1 / total_particles (PHD:245-266) and allocates the exchange buffers. */
int rbphd_comm_init_rank(rbphd_navigator* nav, const unsigned char id128[128], int rank, int world,
                         int total_particles);
int rbphd_comm_destroy(rbphd_navigator* nav);
/* best particle and resampling flag of the most recent frame (multi-GPU: already on the host; single GPU: one
 * synchronising read of the device state) */
int rbphd_frame_result(rbphd_navigator* nav, int* best, int* resampled);
/* since rbphd_comm_init_rank: resampling frames, bytes sent, bytes received, records sent */
int rbphd_comm_stats(const rbphd_navigator* nav, int64_t out4[4]);
/* test hook: the exchange plan rank `rank` of `world` derives on the device from global ancestors and counts
 * (host arrays).  local_src / rec_off: one entry per local particle; send_idx / send_off: local particles + world
 * entries; hdr: 2 * world + 3 entries (doubles to send per rank, doubles to receive per rank, records to pack,
 * records to receive, 1 if the ancestors were sorted). */
int rbphd_debug_migration_plan(int device, const int* ancestors, const int* counts, int total, int world, int rank,
                               int* local_src, int64_t* rec_off, int* send_idx, int64_t* send_off, int64_t* hdr);

/* ---- around the hot path --------------------------------------------------------------
 * SimulatedVehicle.Measure (mono-rfs-lib/SLAM/Vehicles/SimulatedVehicle.cs:243-295) with the host's random numbers
 * (the reference draws them from Util.Uniform / Accord's Gaussian, SIMV:215,257; a host passes the same draws):
 * landmark i is measured when pd_i > 0 and uniforms[i] < pd_i, as h(pose, landmark) + C gauss[3 i ..] with C C^T = R
 * (chol9 row-major lower root; NULL: computed from the configured R as Util.RandomGaussianVector does,
 * Util/Util.cs:173-202); then nc clutter points from clutter_u[3 k ..] (PRM3DMeasurer.RandomMeasure,
 * BaseStructures/Measurers/PRM3DMeasurer.cs:249-256; the Poisson count, capped at 10 lambda, is the host's).
 * z: room for (n + nc) x 3, assoc (may be NULL): n + nc entries, landmark index or INT_MIN for clutter
 * (SimulatedVehicle.DataAssociation).  With a depth frame attached the visibility is the occlusion-aware one. */
int rbphd_generate_measurements(rbphd_navigator* nav, const double* pose7, const double* landmarks, int n,
                                const double* uniforms, const double* gauss, const double* chol9,
                                const double* clutter_u, int nc, double* z, int* assoc, int* count);

/* Plot.OSPA (postanalysis/Plot.cs:531-581): OSPA distance of order p and cutoff c between two landmark sets
 * (positions, na x 3 and nb x 3; swapped when na > nb), the optimal assignment
 * (GraphCombinatorics.LinearAssignment, mono-rfs-lib/Math/GraphCombinatorics.cs:52-175) found on the device.
 * cardinality_error may be NULL.  At most 8192 landmarks per set. */
int rbphd_ospa(rbphd_navigator* nav, const double* a, int na, const double* b, int nb, double c, double p,
               double* ospa, double* cardinality_error);

/* ---- instrumentation ---- */
/* number of kernel launches issued on this handle since creation */
int64_t rbphd_kernel_launches(const rbphd_navigator* nav);
/* Per-stage device times: after rbphd_profile_enable(nav, max_frames) every rbphd_frame_async records
 * CUDA events on the handle's stream around its kernels; rbphd_profile_read() (after a synchronise)
 * returns, per recorded frame, RBPHD_STAGES durations in ms: pose, prep, particle_update,
 * normalize_resample, copy_particles. */
#define RBPHD_STAGES 5
int rbphd_profile_enable(rbphd_navigator* nav, int max_frames);
int rbphd_profile_read(rbphd_navigator* nav, double* ms, int max_frames, int* frames);
/* device-side work counters since the last reset: prior components read, pruned components written,
 * gated (component, measurement) pairs evaluated, particle-frames processed */
int rbphd_get_counters(rbphd_navigator* nav, int64_t out4[4], int reset);
/* diagnostics: SM cycles per internal phase of the fused kernel summed over CTAs (64 entries) followed
 * by 16 event counters */
int rbphd_get_phase_cycles(rbphd_navigator* nav, int64_t out80[80]);
/* launch geometry of the fused per-particle kernel on this handle: threads per CTA, resident CTAs per SM,
 * dynamic shared memory per CTA (bytes), scratch slabs (= CTAs launched), bytes per scratch slab */
int rbphd_launch_shape(const rbphd_navigator* nav, int64_t out5[5]);
/* FP64 pipe microbenchmark on `device` (needs no navigator): out6 = { DFMA TFLOP/s, unfused DMUL+DADD TFLOP/s,
 * DADD TFLOP/s, FP64 thread-instructions/s fused (1e12), the same unfused (1e12), SM count }.  The frame path is
 * compiled without contraction, so the unfused figures are its FP64 roofline. */
int rbphd_bench_fp64(int device, int outer_iterations, double out6[6]);
void* rbphd_stream(rbphd_navigator* nav);   /* cudaStream_t of the handle, for event timing by the host */

#ifdef __cplusplus
}
#endif
#endif
