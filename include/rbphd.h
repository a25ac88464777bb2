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

/* ---- multi-GPU plumbing (particles sharded by rank).  The data-path
 * collectives themselves are issued by the host runtime on these DEVICE buffers. ---- */
/* SlamUpdate split at the coupling point PHD:343: phase 1 = the Parallel.For body + w *= alpha
 * (inputs from slot `slot`, enqueued without synchronisation) */
int rbphd_slam_update_local(rbphd_navigator* nav, int slot, int m, int only_mapping);
/* device pointer to this rank's un-normalised weights (particles doubles) for the allgather */
int rbphd_device_weights(rbphd_navigator* nav, void** dev_ptr, int* particles);
/* phase 2: normalise / best / ESS / wheel over the GLOBAL weight vector (device pointer, all ranks'
 * weights in rank order); fills global ancestors; returns this rank's slice decisions */
int rbphd_resample_global(rbphd_navigator* nav, const void* dev_global_weights, int global_particles,
                          int rank_offset, double u_resample, int* best_global, int* resampled,
                          const int** ancestors_global);
/* pack / unpack particles (pose + map) to flat device records of rbphd_particle_record_bytes() each */
int64_t rbphd_particle_record_bytes(const rbphd_navigator* nav);
int rbphd_pack_particles(rbphd_navigator* nav, const int* local_indices, int count, void** dev_buf,
                         int64_t* bytes);
/* record records[j] of dev_buf becomes local particle slots[j] (a record may be used several times) */
int rbphd_unpack_particles(rbphd_navigator* nav, const void* dev_buf, const int* records, const int* slots,
                           int count);
int rbphd_commit_resample_local(rbphd_navigator* nav, const int* local_sources, int count);

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
/* diagnostics: SM cycles per internal phase of the fused kernel summed over CTAs (32 entries) followed
 * by 16 event counters */
int rbphd_get_phase_cycles(rbphd_navigator* nav, int64_t out48[48]);
/* launch geometry of the fused per-particle kernel on this handle: threads per CTA, resident CTAs per SM,
 * dynamic shared memory per CTA (bytes), scratch slabs (= CTAs launched), bytes per scratch slab */
int rbphd_launch_shape(const rbphd_navigator* nav, int64_t out5[5]);
void* rbphd_stream(rbphd_navigator* nav);   /* cudaStream_t of the handle, for event timing by the host */

#ifdef __cplusplus
}
#endif
#endif
