/*
 * rbphd_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a plain C++ restatement of the
 * MonoRFS (afalchetti/monorfs) Rao-Blackwellized PHD-SLAM per-frame update,
 * written from the reference's C# in the reference's operation order.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (monorfs_b200/, librbphd.so) never
 * links, imports or executes anything under oracle/.
 *
 * Parity status (see oracle/README.md): pinned against the reference's own
 * NUnit known-answer tests restated in tests/test_oracle_golden.py
 * (PHDNavigatorTest, GraphCombinatoricsTest, SimulationTest.resample,
 * LoopyPHDNavigatorTest pixel-range fixtures, Pose3DTest, QuaternionTest).
 * Items the reference's tests do not pin (KD-tree gate metric and order,
 * List.Sort tie order, Accord SVD pseudo-inverse thresholds, WeightAlpha
 * values) are "parity unpinned" and are DEFINED here (documented per function).
 *
 * File:line citations use the abbreviations of SURVEY.md (PHD, GAUSS, MAP,
 * PRM, POSE, QUAT, GC, SPM, MX, UTIL, SIMV, TRK, VEH, CFG).
 */
#ifndef RBPHD_ORACLE_H
#define RBPHD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field layout as rbphd_config in include/rbphd.h (kept separate on
 * purpose: the oracle does not include product headers). */
typedef struct orc_config {
    int32_t model;          /* 0 = PRM3D (Pose3D + pixel-range), 1 = Linear2D */
    int32_t max_quantity;   /* CFG:83 */
    int32_t gate_metric;    /* 0: |d|^2 <= r^2 (default); 1: |d|^2 <= r (Accord squared-metric hypothesis) */
    int32_t nthreads;       /* CFG:46 NParallel */
    double  R[9];           /* measurement covariance, row-major, stride 3 (Linear2D: top-left 2x2) */
    double  Q[36];          /* motion covariance 6x6 */
    double  pd;             /* CFG:90 NavigatorPD */
    double  clutter;        /* CFG:91 NavigatorClutterDensity */
    double  birth_cov[9];   /* CFG:77 */
    double  birth_weight;   /* CFG:80 */
    double  min_weight;     /* CFG:81 */
    double  merge_threshold;            /* CFG:84 */
    double  exploration_threshold;      /* CFG:85 */
    double  density_distance_threshold; /* CFG:74 */
    double  min_effective_particle;     /* CFG:82 */
    double  visibility_ramp[3];         /* CFG:257 */
    double  measurer[7];    /* PRM:92 focal,rangemin,rangemax,filmX,filmY,filmW,filmH ; Linear2D: range */
} orc_config;

/* ---- geometry (QUAT, POSE, PRM) ---- */
void orc_quat_mul(const double a[4], const double b[4], double out[4]);
void orc_quat_exp(const double lie[3], double out[4]);
void orc_quat_log(const double q[4], double out[3]);
void orc_quat_sqrt(const double q[4], double out[4]);
void orc_quat_from_ypr(double yaw, double pitch, double roll, double out[4]);
void orc_quat_to_matrix(const double q[4], double out[9]);
void orc_quat_vector_rotator(const double from[3], const double to[3], double out[4]);
void orc_pose_from_state(const double state[7], double out[7]);
void orc_pose_add_odometry(const double pose[7], const double delta[6], double out[7]);
void orc_pose_diff_odometry(const double pose[7], const double origin[7], double out[6]);
void orc_measure_perfect(const orc_config* c, const double* pose, const double m[3], double out[3]);
void orc_measurement_jacobian_l(const orc_config* c, const double* pose, const double m[3], double out[9]);
void orc_measure_to_map(const orc_config* c, const double* pose, const double z[3], double out[3]);
double orc_fuzzy_visible(const orc_config* c, const double z[3]);
void orc_fit_to_measurement(const orc_config* c, const double pose0[7], const double z[3],
                            const double landmark[3], double out[7]);

/* ---- gaussian (GAUSS) ---- */
double orc_gaussian_evaluate(const double m[3], const double P[9], const double x[3]);
void   orc_gaussian_merge(int n, const double* w, const double* m, const double* P,
                          double* ow, double om[3], double oP[9]);

/* ---- per-particle stages (PHD:793-959, PHD:373-515).  Maps are AoS: w[n], m[n*3], P[n*9]. ---- */
int orc_predict(const orc_config* c, const double* pose, int n, const double* w, const double* m,
                const double* P, int M, const double* z, int cap, double* ow, double* om, double* oP,
                int* nbirth);
/* gate_radius < 0 -> ungated (the form the reference's Correct test pins). */
int orc_correct(const orc_config* c, const double* pose, int n, const double* w, const double* m,
                const double* P, int M, const double* z, double gate_radius, int cap, double* ow,
                double* om, double* oP);
int orc_prune(const orc_config* c, int n, const double* w, const double* m, const double* P, int cap,
              double* ow, double* om, double* oP);
int orc_best_map_estimate(int n, const double* w, int cap, int* picks);
double orc_set_loglikelihood(const orc_config* c, const double* pose, int J, const double* jm, int M,
                             const double* z);
/* KinectMeasurer.FuzzyVisibleM (KinectMeasurer.cs:151-173): attach a depth frame depth_xy[x * resy + y] to the
 * PRM3D measurer of every following call (NULL detaches).  One global frame: test infrastructure. */
void orc_set_depth_frame(const float* depth_xy, int resx, int resy);
/* PHD:526-532, 561-713 (value only): full visibility, gate d < 12 */
double orc_quasi_set_loglikelihood(const orc_config* c, const double* pose, int J, const double* jm, int M,
                                   const double* z);
/* PHD:544-549, 561-713 with the gradient: returns the value, fills gradient[OdoSize] (6 for PRM3D, 2 for Linear2D) */
double orc_quasi_set_loglikelihood_gradient(const orc_config* c, const double* pose, int J, const double* jm, int M,
                                            const double* z, double* gradient);
/* experiment switch for D10 (oracle/README.md): 1 = TemperedAverage normalises by the sum instead of Accord's Euclidean norm */
void orc_set_tempered_norm(int sum_instead_of_euclidean);
/* MeasurementJacobianP (PRM:185-209, Linear2DMeasurer.cs:133-137): dz x OdoSize, row stride 6 */
void orc_measurement_jacobian_p(const orc_config* c, const double* pose, const double m[3], double out[18]);
/* PHD:415-460 (quasi: PHD:561-640) as triplets in insertion order; returns the number of entries */
int orc_set_loglike_matrix(const orc_config* c, const double* pose, int J, const double* jm, int M, const double* z,
                           int quasi, int cap, int* rows, int* cols, double* vals);
/* out[0]=alpha, out[1]=setloglik, out[2]=ploglik, out[3]=cloglik, out[4]=pcount, out[5]=ccount, out[6]=J */
void orc_weight_alpha(const orc_config* c, const double* pose, int M, const double* z, int np,
                      const double* pw, const double* pm, const double* pP, int nc, const double* cw,
                      const double* cm, const double* cP, double out[7]);

/* ---- particle set (PHD:343-358, 724-777) ---- */
/* weights in/out (normalised in place); returns 1 if resampled. ancestors[P] filled (identity when not). */
int orc_normalize_resample(const orc_config* c, int P, double* weights, double u, int* best,
                           int* ancestors, int force_resample);

/* ---- graph combinatorics (GC, SPM) on dense n x n matrices with a 'defined' mask ---- */
/* postanalysis/Plot.cs:531-581: OSPA distance between two landmark sets (positions, n x 3) */
double orc_ospa(int na, const double* a, int nb, const double* b, double C, double P, double* cardinality);
/* SIMV:243-295 with caller-supplied random numbers; returns the measurement count */
int orc_generate_measurements(const orc_config* c, const double* pose, int n, const double* landmarks,
                              const double* uniforms, const double* gauss, const double* chol, int nc,
                              const double* clutter_u, double* z, int* assoc);
int orc_hungarian(int n, const double* val, const uint8_t* defined, double defval, int* match);
int orc_connected_components(int h, int w, const uint8_t* defined);
int orc_lexicographical(int n, const double* val, const uint8_t* defined, double defval, int modelsize,
                        int cap, int* perms, double* values);
int orc_murty(int n, const double* val, const uint8_t* defined, double defval, int cap, int* perms,
              double* values);
/* forced/eliminated as (i,k) pairs; children written as [nforced, nelim, pairs...] records; returns count */
int orc_murty_children(int n, const int* assignment, int nf, const int* forced, int ne, const int* elim,
                       int cap, int* out);

/* ---- whole navigator (PHD:192-362) ---- */
typedef struct orc_nav orc_nav;
orc_nav* orc_nav_new(const orc_config* c, int P, const double* pose, int only_mapping);
void orc_nav_delete(orc_nav* nav);
void orc_nav_set_map(orc_nav* nav, int i, int n, const double* w, const double* m, const double* P);
int  orc_nav_get_map(orc_nav* nav, int i, int cap, double* w, double* m, double* P);
void orc_nav_set_pose(orc_nav* nav, int i, const double* pose);
void orc_nav_get_poses(orc_nav* nav, double* poses);
void orc_nav_set_weights(orc_nav* nav, const double* w);
void orc_nav_get_weights(orc_nav* nav, double* w);
void orc_nav_get_alphas(orc_nav* nav, double* a);
int  orc_nav_particle_count(orc_nav* nav);
/* gauss: P x 6 N(0,1) draws (second draw block of TRK:95); perfect_still: CFG PerfectStill */
void orc_nav_update(orc_nav* nav, const double reading[6], double dt, const double* gauss, int perfect_still);
/* [first,last) restricts the Parallel.For body to a particle range (used by bounded CPU timing). */
void orc_nav_slam_update(orc_nav* nav, int M, const double* z, double u, int* best, int* resampled,
                         int* ancestors);
void orc_nav_map_update_range(orc_nav* nav, int M, const double* z, int first, int last);

#ifdef __cplusplus
}
#endif
#endif
