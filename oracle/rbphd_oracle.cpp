/*
 * rbphd_oracle.cpp -- CPU oracle: restatement of MonoRFS's RB-PHD SLAM update.
 *
 * TEST INFRASTRUCTURE ONLY (see rbphd_oracle.h).  Build: oracle/Makefile
 * (g++ -O2 -ffp-contract=off: no FMA contraction, so +,-,*,/ and sqrt round
 * exactly as the reference's scalar C# does).
 *
 * Every function cites the reference file:line it follows (abbreviations of
 * SURVEY.md).  Things the reference leaves to un-vendored libraries are
 * DEFINED here and listed in oracle/README.md:
 *   D1  3x3 / 2x2 inverse and determinant: closed form (adjugate * (1/det))
 *       instead of Accord's SVD PseudoInverse/PseudoDeterminant (GAUSS:152-153).
 *   D2  Map enumeration order = insertion order (Accord KDTree order unpinned).
 *   D3  Radius gate: |m - c|^2 <= r^2 (gate_metric 0) or <= r (gate_metric 1).
 *   D4  List.Sort -> stable sort (weight descending, then insertion index).
 *   D5  Dictionary enumeration = insertion order (what .NET does without
 *       re-insertion after removal).
 *   D6  Accord vector/matrix helpers accumulate left to right starting at 0.
 */
#include "rbphd_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

const double kInf = std::numeric_limits<double>::infinity();
const double kPi  = 3.14159265358979323846;

/* ------------------------------------------------------------------ */
/* small dense helpers (D6)                                            */
/* ------------------------------------------------------------------ */
struct Mat3 { double a[9]; };

inline double dot3(const double* a, const double* b)
{
    double s = 0;
    for (int i = 0; i < 3; i++) s += a[i] * b[i];
    return s;
}

inline double euclid(const double* v, int n)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += v[i] * v[i];
    return std::sqrt(s);
}

/* C = A(ra x ca) * B(ca x cb), row-major with explicit strides; sums run k = 0.. in order */
inline void matmul(const double* A, int ra, int ca, int lda, const double* B, int cb, int ldb,
                   double* C, int ldc)
{
    for (int i = 0; i < ra; i++)
        for (int j = 0; j < cb; j++) {
            double s = 0;
            for (int k = 0; k < ca; k++) s += A[i * lda + k] * B[k * ldb + j];
            C[i * ldc + j] = s;
        }
}

inline void matvec(const double* A, int r, int c, int lda, const double* x, double* y)
{
    for (int i = 0; i < r; i++) {
        double s = 0;
        for (int k = 0; k < c; k++) s += A[i * lda + k] * x[k];
        y[i] = s;
    }
}

/* D1: closed-form inverse/determinant of the top-left dim x dim block (stride 3) */
inline double inverse_dim(const double* a, int dim, double* inv)
{
    for (int i = 0; i < 9; i++) inv[i] = 0;
    if (dim == 3) {
        double c00 = a[4] * a[8] - a[5] * a[7];
        double c01 = a[3] * a[8] - a[5] * a[6];
        double c02 = a[3] * a[7] - a[4] * a[6];
        double det = a[0] * c00 - a[1] * c01 + a[2] * c02;
        double id  = 1.0 / det;
        inv[0] = c00 * id;
        inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
        inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
        inv[3] = (a[5] * a[6] - a[3] * a[8]) * id;
        inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
        inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
        inv[6] = c02 * id;
        inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
        inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
        return det;
    }
    if (dim == 2) {
        double det = a[0] * a[4] - a[1] * a[3];
        double id  = 1.0 / det;
        inv[0] = a[4] * id;
        inv[1] = -a[1] * id;
        inv[3] = -a[3] * id;
        inv[4] = a[0] * id;
        return det;
    }
    inv[0] = 1.0 / a[0];
    return a[0];
}

/* ------------------------------------------------------------------ */
/* Gaussian (GAUSS:40-157)                                             */
/* ------------------------------------------------------------------ */
struct Gaussian {
    double w;
    double m[3];
    double P[9];     /* stride 3; measurement-space gaussians of dim 2 use the top-left block */
    double Pinv[9];
    double det;
    double mult;
    int    dim;
};

/* GAUSS:148-157.  Multiplier uses INTEGER division -dim/2 (quirk A9.1): -1 for dim 2 and 3, 0 for dim 1 */
Gaussian make_gaussian(const double* m, const double* P, double w, int dim = 3)
{
    Gaussian g;
    g.dim = dim;
    for (int i = 0; i < 3; i++) g.m[i] = (i < dim) ? m[i] : 0.0;
    for (int i = 0; i < 9; i++) g.P[i] = P[i];
    g.det = inverse_dim(P, dim, g.Pinv);
    g.w   = std::isnan(w) ? 0.0 : w;
    int ipow = -dim / 2;
    double twopi_pow = (ipow == 0) ? 1.0 : 1.0 / (2 * kPi);   /* Math.Pow(2 pi, -1) */
    g.mult = twopi_pow / std::sqrt(g.det);
    return g;
}

/* x^T Pinv x with Accord's order: InnerProduct(diff, Pinv.Multiply(diff)) */
inline double quadform(const Gaussian& g, const double* diff)
{
    double t[3];
    for (int i = 0; i < g.dim; i++) {
        double s = 0;
        for (int k = 0; k < g.dim; k++) s += g.Pinv[i * 3 + k] * diff[k];
        t[i] = s;
    }
    double s = 0;
    for (int i = 0; i < g.dim; i++) s += diff[i] * t[i];
    return s;
}

/* GAUSS:199-204 */
inline double evaluate(const Gaussian& g, const double* x)
{
    double diff[3];
    for (int i = 0; i < g.dim; i++) diff[i] = x[i] - g.m[i];
    return g.mult * std::exp(-0.5 * quadform(g, diff));
}

/* GAUSS:354-358 (diff = Mean - point) */
inline double mahalanobis(const Gaussian& g, const double* x)
{
    double diff[3];
    for (int i = 0; i < g.dim; i++) diff[i] = g.m[i] - x[i];
    return std::sqrt(quadform(g, diff));
}

/* GAUSS:365-369 */
inline double square_mahalanobis(const Gaussian& g, const double* x)
{
    double diff[3];
    for (int i = 0; i < g.dim; i++) diff[i] = g.m[i] - x[i];
    return quadform(g, diff);
}

/* GAUSS:243-246 */
inline bool are_close(const Gaussian& a, const Gaussian& b, double threshold)
{
    return square_mahalanobis(a, b.m) < threshold * threshold;
}

/* GAUSS:297-347 (raw-moment form, restated literally) */
Gaussian merge(const std::vector<const Gaussian*>& comps)
{
    const Gaussian& first = *comps[0];
    double weight = 0.0;
    double mean[3] = {0, 0, 0};
    double cov[9]  = {0, 0, 0, 0, 0, 0, 0, 0, 0};

    for (const Gaussian* g : comps) {
        double w = g->w;
        weight += w;
        for (int i = 0; i < 3; i++) mean[i] = mean[i] + w * g->m[i];
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 3; k++)
                cov[i * 3 + k] = cov[i * 3 + k] + w * (g->P[i * 3 + k] + g->m[i] * g->m[k]);
    }

    if (weight < 1e-15) {   /* GAUSS:339-341, UTIL:143-153 */
        double inf[9] = {1e12, 0, 0, 0, 1e12, 0, 0, 0, 1e12};
        return make_gaussian(first.m, inf, 0.0);
    }

    for (int i = 0; i < 3; i++) mean[i] = mean[i] / weight;          /* Accord vector Divide */
    double r = 1 / weight;                                            /* MX:264-267 */
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) cov[i * 3 + k] = r * cov[i * 3 + k] - mean[i] * mean[k];

    return make_gaussian(mean, cov, weight);
}

typedef std::vector<Gaussian> Map;   /* D2: insertion-ordered flat list (MAP:41-81) */

/* ------------------------------------------------------------------ */
/* Quaternion (QUAT) -- stored (w,x,y,z)                               */
/* ------------------------------------------------------------------ */
struct Quat { double w, x, y, z; };

inline Quat qmul(const Quat& a, const Quat& b)   /* QUAT:295-301 */
{
    Quat r;
    r.w = a.w * b.w - (a.x * b.x + a.y * b.y + a.z * b.z);
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
inline Quat qconj(const Quat& q) { return Quat{q.w, -q.x, -q.y, -q.z}; }   /* QUAT:155-158 */
inline Quat qscale(double a, const Quat& q) { return Quat{a * q.w, a * q.x, a * q.y, a * q.z}; }
inline Quat qnormalize(const Quat& q)   /* QUAT:240-245 */
{
    double mag = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    double alpha = 1 / mag;
    return qscale(alpha, q);
}
Quat qexp(const double* lie)   /* QUAT:185-196 */
{
    double phi = euclid(lie, 3);
    if (phi < 1e-12) return Quat{1, 0, 0, 0};
    double s = std::sin(phi);
    /* lie.Normalize(): divide by the Euclidean norm (D6) */
    return Quat{std::cos(phi), s * (lie[0] / phi), s * (lie[1] / phi), s * (lie[2] / phi)};
}
void qlog(const Quat& q0, double* out)   /* QUAT:203-217 */
{
    Quat q = qnormalize(q0);
    double phi = std::acos(q.w);
    double v[3] = {q.x, q.y, q.z};
    double mag = euclid(v, 3);
    if (mag < 1e-12) { out[0] = out[1] = out[2] = 0; return; }
    for (int i = 0; i < 3; i++) out[i] = phi * (v[i] / mag);
}
Quat qsqrt(const Quat& q)   /* QUAT:225-235 */
{
    if (std::fabs(q.w - -1.0) < 1e-8) return Quat{1, 0, 0, 0};
    double rw = std::sqrt(0.5 * (1 + q.w));
    double alpha = 1 / (2 * rw);
    return Quat{rw, alpha * q.x, alpha * q.y, alpha * q.z};
}
void qtomatrix(const Quat& q, double* m)   /* QUAT:327-342 */
{
    double xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    double xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w;
    double yz = q.y * q.z, yw = q.y * q.w, zw = q.z * q.w;
    m[0] = 1 - 2 * (yy + zz); m[1] = 2 * (xy - zw);     m[2] = 2 * (xz + yw);
    m[3] = 2 * (xy + zw);     m[4] = 1 - 2 * (xx + zz); m[5] = 2 * (yz - xw);
    m[6] = 2 * (xz - yw);     m[7] = 2 * (yz + xw);     m[8] = 1 - 2 * (xx + yy);
}

/* ------------------------------------------------------------------ */
/* Pose3D (POSE) -- state (x,y,z,qw,qx,qy,qz)                          */
/* ------------------------------------------------------------------ */
struct Pose { double t[3]; Quat q; };

inline Pose pose_load(const double* s) { return Pose{{s[0], s[1], s[2]}, Quat{s[3], s[4], s[5], s[6]}}; }
inline void pose_store(const Pose& p, double* s)
{
    s[0] = p.t[0]; s[1] = p.t[1]; s[2] = p.t[2];
    s[3] = p.q.w;  s[4] = p.q.x;  s[5] = p.q.y;  s[6] = p.q.z;
}
Pose pose_from_state(const double* s)   /* POSE:142-162 (normalises) */
{
    double w = s[3], x = s[4], y = s[5], z = s[6];
    double a = 1.0 / std::sqrt(w * w + x * x + y * y + z * z);
    return Pose{{s[0], s[1], s[2]}, Quat{a * w, a * x, a * y, a * z}};
}
Pose add_odometry(const Pose& p, const double* d)   /* POSE:314-333 */
{
    double half[3] = {0.5 * d[3], 0.5 * d[4], 0.5 * d[5]};   /* QUAT:145-149 */
    Quat dq   = qexp(half);
    Quat newq = qmul(p.q, dq);
    Quat midd = qsqrt(dq);
    Quat midr = qmul(p.q, midd);
    Quat dl   = qmul(qmul(midr, Quat{0, d[0], d[1], d[2]}), qconj(midr));
    Pose r;
    r.t[0] = p.t[0] + dl.x; r.t[1] = p.t[1] + dl.y; r.t[2] = p.t[2] + dl.z;
    r.q = qnormalize(newq);
    return r;
}
void diff_odometry(const Pose& self, const Pose& origin, double* out)   /* POSE:339-359 */
{
    Quat dq   = qmul(qconj(origin.q), self.q);
    Quat midd = qsqrt(dq);
    Quat midr = qmul(origin.q, midd);
    double dxg[3] = {self.t[0] - origin.t[0], self.t[1] - origin.t[1], self.t[2] - origin.t[2]};
    Quat dx = qmul(qmul(qconj(midr), Quat{0, dxg[0], dxg[1], dxg[2]}), midr);
    double lie[3];
    qlog(dq, lie);                       /* ToLinear = 2 * Log (QUAT:135-139) */
    out[0] = dx.x; out[1] = dx.y; out[2] = dx.z;
    out[3] = 2 * lie[0]; out[4] = 2 * lie[1]; out[5] = 2 * lie[2];
}

/* ------------------------------------------------------------------ */
/* Measurers: PRM3D (PRM) and Linear2D (Linear2DMeasurer.cs)           */
/* ------------------------------------------------------------------ */
inline int meas_dim(const orc_config* c) { return c->model == 0 ? 3 : 2; }

void measure_perfect(const orc_config* c, const double* pose, const double* m, double* out)
{
    if (c->model == 1) {   /* Linear2DMeasurer.cs:111-114 */
        out[0] = m[0] - pose[0]; out[1] = m[1] - pose[1]; out[2] = 0;
        return;
    }
    /* PRM:138-149 */
    Pose p = pose_load(pose);
    double diff[3] = {m[0] - p.t[0], m[1] - p.t[1], m[2] - p.t[2]};
    Quat local = qmul(qmul(qconj(p.q), Quat{0, diff[0], diff[1], diff[2]}), p.q);
    double sgn = (local.z > 0) ? 1.0 : ((local.z < 0) ? -1.0 : 0.0);   /* Math.Sign */
    double f = c->measurer[0];
    out[2] = sgn * euclid(diff, 3);
    out[0] = f * local.x / local.z;
    out[1] = f * local.y / local.z;
}

/* H is dz x 3, stride 3 */
void jacobian_l(const orc_config* c, const double* pose, const double* m, double* H)
{
    if (c->model == 1) {   /* Linear2DMeasurer.cs:122-126 */
        double h[9] = {1, 0, 0, 0, 1, 0, 0, 0, 0};
        std::memcpy(H, h, sizeof h);
        return;
    }
    /* PRM:157-177 */
    Pose p = pose_load(pose);
    double diff[3] = {m[0] - p.t[0], m[1] - p.t[1], m[2] - p.t[2]};
    Quat l = qmul(qmul(qconj(p.q), Quat{0, diff[0], diff[1], diff[2]}), p.q);
    double f = c->measurer[0];
    double mag = ((l.z > 0) ? 1 : -1) * std::sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
    double jp[9] = {f / l.z, 0,       -f * l.x / (l.z * l.z),
                    0,       f / l.z, -f * l.y / (l.z * l.z),
                    l.x / mag, l.y / mag, l.z / mag};
    double jr[9];
    qtomatrix(qconj(p.q), jr);
    matmul(jp, 3, 3, 3, jr, 3, 3, H, 3);
}

inline int odo_size(const orc_config* c) { return c->model == 0 ? 6 : 2; }

/* MeasurementJacobianP: dz x OdoSize, stride 6 (PRM:185-209; Linear2DMeasurer.cs:133-137) */
void jacobian_p(const orc_config* c, const double* pose, const double* m, double* Jp)
{
    for (int i = 0; i < 18; i++) Jp[i] = 0;
    if (c->model == 1) {
        Jp[0] = -1; Jp[6 + 1] = -1;
        return;
    }
    Pose p = pose_load(pose);
    double diff[3] = {m[0] - p.t[0], m[1] - p.t[1], m[2] - p.t[2]};
    Quat l = qmul(qmul(qconj(p.q), Quat{0, diff[0], diff[1], diff[2]}), p.q);
    double f = c->measurer[0];
    double mag = ((l.z > 0) ? 1 : -1) * std::sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
    double jp[9] = {f / l.z, 0,       -f * l.x / (l.z * l.z),
                    0,       f / l.z, -f * l.y / (l.z * l.z),
                    l.x / mag, l.y / mag, l.z / mag};
    double rot[9], jloc[9], cross[9], jrot[9];
    qtomatrix(qconj(p.q), rot);
    for (int i = 0; i < 9; i++) jloc[i] = -1.0 * rot[i];                 /* (-1.0).Multiply(R(q*)) */
    double cr[9] = {0, -diff[2], diff[1], diff[2], 0, -diff[0], -diff[1], diff[0], 0};   /* UTIL:107-112 */
    std::memcpy(cross, cr, sizeof cr);
    matmul(jloc, 3, 3, 3, cross, 3, 3, jrot, 3);
    double jlocal[18];
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) { jlocal[r * 6 + k] = jloc[r * 3 + k]; jlocal[r * 6 + 3 + k] = jrot[r * 3 + k]; }
    matmul(jp, 3, 3, 3, jlocal, 6, 6, Jp, 6);
}

void measure_to_map(const orc_config* c, const double* pose, const double* z, double* out)
{
    if (c->model == 1) {   /* Linear2DMeasurer.cs:198-201 */
        out[0] = pose[0] + z[0]; out[1] = pose[1] + z[1]; out[2] = 0;
        return;
    }
    /* PRM:299-312 */
    Pose p = pose_load(pose);
    double f = c->measurer[0];
    double px = z[0], py = z[1], range = z[2];
    double alpha = range / std::sqrt(f * f + px * px + py * py);
    double diff[3] = {alpha * px, alpha * py, alpha * f};
    Quat r = qmul(qmul(p.q, Quat{0, diff[0], diff[1], diff[2]}), qconj(p.q));
    out[0] = p.t[0] + r.x; out[1] = p.t[1] + r.y; out[2] = p.t[2] + r.z;
}

/* depth frame of the KinectMeasurer variant (test infrastructure: one global frame, set by orc_set_depth_frame) */
static const float* g_depth = nullptr;
static int g_resx = 0, g_resy = 0;
static std::vector<float> g_depth_store;

double fuzzy_visible(const orc_config* c, const double* z)
{
    double mind = kInf;
    if (c->model == 1) {   /* Linear2DMeasurer.cs:179-190 */
        double range = c->measurer[0];
        mind = std::fmin(mind, (z[0] - -range) / c->visibility_ramp[0]);
        mind = std::fmin(mind, (range - z[0]) / c->visibility_ramp[0]);
        mind = std::fmin(mind, (z[1] - -range) / c->visibility_ramp[1]);
        mind = std::fmin(mind, (range - z[1]) / c->visibility_ramp[1]);
        return std::fmax(0, std::fmin(1, mind));
    }
    /* PRM:277-291; film is an int Rectangle, range clip is a float Range (PRM:65,73,110-111) */
    int left = (int)c->measurer[3], top = (int)c->measurer[4];
    int right = left + (int)c->measurer[5], bottom = top + (int)c->measurer[6];
    double rmin = (double)(float)c->measurer[1], rmax = (double)(float)c->measurer[2];
    mind = std::fmin(mind, (z[0] - left) / c->visibility_ramp[0]);
    mind = std::fmin(mind, (right - z[0]) / c->visibility_ramp[0]);
    mind = std::fmin(mind, (z[1] - top) / c->visibility_ramp[1]);
    mind = std::fmin(mind, (bottom - z[1]) / c->visibility_ramp[1]);
    mind = std::fmin(mind, (z[2] - rmin) / c->visibility_ramp[2]);
    mind = std::fmin(mind, (rmax - z[2]) / c->visibility_ramp[2]);
    mind = std::fmax(0, std::fmin(1, mind));
    if (g_depth) {   /* KinectMeasurer.FuzzyVisibleM, KinectMeasurer.cs:151-173 */
        if (mind == 0) return 0;
        float resx = (float)g_resx, resy = (float)g_resy;
        int x = (int)(z[0] + resx / 2), y = (int)(z[1] + resy / 2);
        if (x < 0 || x >= g_resx || y < 0 || y >= g_resy) return 0;   /* (the C# would throw; unreachable inside the film) */
        float d = g_depth[(size_t)x * g_resy + y];
        if (std::isnan(d)) return 0;
        float range = (float)z[2], rminf = (float)c->measurer[1];
        mind = std::fmin(mind, (range - rminf) / c->visibility_ramp[2]);
        mind = std::fmin(mind, (d - range) / c->visibility_ramp[2]);
        mind = std::fmax(0, std::fmin(1, mind));
    }
    return mind;
}

/* SIMV:324-339 */
inline double detection_probability_m(const orc_config* c, const double* z)
{
    return fuzzy_visible(c, z) * c->pd;
}

/* ------------------------------------------------------------------ */
/* Map queries (MAP:170-220), D2 + D3                                  */
/* ------------------------------------------------------------------ */
inline bool in_gate(const orc_config* c, const double* a, const double* b, double radius)
{
    double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    return (c->gate_metric == 0) ? (d2 <= radius * radius) : (d2 <= radius);
}

/* MAP:210-220 */
double map_evaluate_gated(const orc_config* c, const Map& map, const double* x, double radius)
{
    double value = 0;
    for (const Gaussian& g : map)
        if (in_gate(c, g.m, x, radius)) value += g.w * evaluate(g, x);
    return value;
}

/* MAP:192-202 */
double map_evaluate(const Map& map, const double* x)
{
    double value = 0;
    for (const Gaussian& g : map) value += g.w * evaluate(g, x);
    return value;
}

/* MAP:61-71 */
double expected_size(const Map& map)
{
    double e = 0;
    for (const Gaussian& g : map) e += g.w;
    return e;
}

/* PHD:956-959 */
inline bool explored(const orc_config* c, const Map& model, const double* x)
{
    return map_evaluate_gated(c, model, x, 3 * c->density_distance_threshold) >= c->exploration_threshold;
}

/* ------------------------------------------------------------------ */
/* PredictConditional (PHD:793-819)                                    */
/* ------------------------------------------------------------------ */
Map predict_conditional(const orc_config* c, const double* pose, const Map& model, int M, const double* z,
                        int* nbirth)
{
    Map predicted(model);
    std::vector<double> unexplored;
    for (int k = 0; k < M; k++) {
        double cand[3];
        measure_to_map(c, pose, z + 3 * k, cand);
        if (!explored(c, model, cand)) unexplored.insert(unexplored.end(), cand, cand + 3);
    }
    for (size_t b = 0; b < unexplored.size() / 3; b++)
        predicted.push_back(make_gaussian(&unexplored[3 * b], c->birth_cov, c->birth_weight));
    if (nbirth) *nbirth = (int)(unexplored.size() / 3);
    return predicted;
}

/* ------------------------------------------------------------------ */
/* CorrectConditional (PHD:829-906)                                    */
/* ------------------------------------------------------------------ */
Map correct_conditional(const orc_config* c, const double* pose, const Map& model, int M, const double* z,
                        double gate_radius)
{
    const int dz = meas_dim(c);
    const int n  = (int)model.size();
    Map corrected;
    corrected.reserve(n + M);

    /* PHD:837-840 */
    for (const Gaussian& g : model) {
        double mp[3];
        measure_perfect(c, pose, g.m, mp);
        double pd = detection_probability_m(c, mp);
        Gaussian r = g;
        r.w = (1 - pd) * g.w;
        corrected.push_back(r);
    }

    /* PHD:858-870 */
    std::vector<double>   mp(3 * n), H(9 * n), PH(9 * n), PD(n);
    std::vector<Gaussian> mc(n);
    for (int i = 0; i < n; i++) {
        const Gaussian& g = model[i];
        measure_perfect(c, pose, g.m, &mp[3 * i]);
        jacobian_l(c, pose, g.m, &H[9 * i]);
        /* PH = P * H^T : (3x3)*(3xdz) via an explicit transpose (MX:276-279) */
        double Ht[9];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) Ht[a * 3 + b] = H[9 * i + b * 3 + a];
        matmul(g.P, 3, 3, 3, Ht, dz, 3, &PH[9 * i], 3);
        double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        matmul(&H[9 * i], dz, 3, 3, &PH[9 * i], dz, 3, S, 3);
        for (int a = 0; a < dz; a++)
            for (int b = 0; b < dz; b++) S[a * 3 + b] = S[a * 3 + b] + c->R[a * 3 + b];
        mc[i] = make_gaussian(&mp[3 * i], S, g.w, dz);
        PD[i] = detection_probability_m(c, &mp[3 * i]);
    }

    /* PHD:881-903 */
    std::vector<int> near;
    for (int k = 0; k < M; k++) {
        const double* zk = z + 3 * k;
        double cand[3];
        measure_to_map(c, pose, zk, cand);
        near.clear();
        for (int i = 0; i < n; i++)
            if (gate_radius < 0 || in_gate(c, model[i].m, cand, gate_radius)) near.push_back(i);

        double weightsum = 0;
        for (int i : near) weightsum += PD[i] * mc[i].w * evaluate(mc[i], zk);

        for (int i : near) {
            double gain[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            matmul(&PH[9 * i], 3, dz, 3, mc[i].Pinv, dz, 3, gain, 3);
            double innov[3] = {0, 0, 0}, kd[3];
            for (int a = 0; a < dz; a++) innov[a] = zk[a] - mp[3 * i + a];
            matvec(gain, 3, dz, 3, innov, kd);
            double mean[3];
            for (int a = 0; a < 3; a++) mean[a] = model[i].m[a] + kd[a];
            double KH[9], IKH[9], cov[9];
            matmul(gain, 3, dz, 3, &H[9 * i], 3, 3, KH, 3);
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) IKH[a * 3 + b] = ((a == b) ? 1.0 : 0.0) - KH[a * 3 + b];
            matmul(IKH, 3, 3, 3, model[i].P, 3, 3, cov, 3);
            double weight = PD[i] * mc[i].w * evaluate(mc[i], zk) / (c->clutter + weightsum);
            corrected.push_back(make_gaussian(mean, cov, weight));
        }
    }
    return corrected;
}

/* ------------------------------------------------------------------ */
/* PruneModel (PHD:913-948), D4                                        */
/* ------------------------------------------------------------------ */
Map prune_model(const orc_config* c, const Map& model)
{
    Map pruned;
    std::vector<Gaussian> landmarks(model);
    std::stable_sort(landmarks.begin(), landmarks.end(),
                     [](const Gaussian& a, const Gaussian& b) { return (b.w - a.w) < 0; });

    int weightcut = 0;
    int lim = std::min(c->max_quantity, (int)landmarks.size());
    for (weightcut = 0; weightcut < lim; weightcut++)
        if (landmarks[weightcut].w < c->min_weight) break;

    for (int i = 0; i < weightcut; i++) {
        std::vector<Gaussian> close;
        close.push_back(landmarks[i]);
        for (int k = i + 1; k < weightcut; k++) {
            if (are_close(landmarks[i], landmarks[k], c->merge_threshold)) {
                close.push_back(landmarks[k]);
                landmarks.erase(landmarks.begin() + k);
                k--;
                weightcut--;
            }
        }
        std::vector<const Gaussian*> ptrs;
        for (const Gaussian& g : close) ptrs.push_back(&g);
        pruned.push_back(merge(ptrs));
    }
    return pruned;
}

/* ------------------------------------------------------------------ */
/* BestMapEstimate (MAP:119-142), D4.  Returns the picked component    */
/* indices in emission order.  The reference re-sorts the whole list   */
/* after appending one element; with a stable sort that equals an      */
/* insertion after the last element whose weight is >= the new one.    */
/* ------------------------------------------------------------------ */
std::vector<int> best_map_estimate(int n, const double* w)
{
    struct Item { double w; int src; };
    std::vector<Item> mlist(n);
    double expected = 0;
    for (int i = 0; i < n; i++) { mlist[i] = Item{w[i], i}; expected += w[i]; }
    int size = (int)expected;
    std::stable_sort(mlist.begin(), mlist.end(),
                     [](const Item& a, const Item& b) { return (b.w - a.w) < 0; });
    std::vector<int> picks;
    for (int i = 0; i < size; i++) {
        Item comp = mlist[i];
        picks.push_back(comp.src);
        Item re{comp.w - 1, comp.src};
        size_t pos = mlist.size();
        while (pos > 0 && (re.w - mlist[pos - 1].w) > 0) pos--;   /* stable: after all >= */
        mlist.insert(mlist.begin() + pos, re);
    }
    return picks;
}

/* ------------------------------------------------------------------ */
/* Sparse matrix with .NET Dictionary semantics (SPM), D5.             */
/* ------------------------------------------------------------------ */
struct SparseRow { int key; std::vector<std::pair<int, double> > items; };
struct Sparse {
    std::vector<SparseRow> rows;   /* insertion ordered */
    double defval;
    int width, height;
    Sparse(int h = 0, int w = 0, double d = 0) : defval(d), width(w), height(h) {}
    SparseRow* find_row(int i)
    {
        for (SparseRow& r : rows) if (r.key == i) return &r;
        return nullptr;
    }
    const SparseRow* find_row(int i) const
    {
        for (const SparseRow& r : rows) if (r.key == i) return &r;
        return nullptr;
    }
    double get(int i, int k) const   /* SPM:141-148 */
    {
        const SparseRow* r = find_row(i);
        if (r) for (const auto& it : r->items) if (it.first == k) return it.second;
        return defval;
    }
    bool defines(int i, int k) const
    {
        const SparseRow* r = find_row(i);
        if (r) for (const auto& it : r->items) if (it.first == k) return true;
        return false;
    }
    void set(int i, int k, double v)   /* SPM:150-164 */
    {
        SparseRow* r = find_row(i);
        if (!r) { rows.push_back(SparseRow{i, {}}); r = &rows.back(); }
        bool found = false;
        for (auto& it : r->items) if (it.first == k) { it.second = v; found = true; break; }
        if (!found) r->items.push_back(std::make_pair(k, v));
        height = std::max(height, i + 1);
        width  = std::max(width, k + 1);
    }
    int count() const { int c = 0; for (const SparseRow& r : rows) c += (int)r.items.size(); return c; }
    void remove_row(int i)   /* SPM:385-388 */
    {
        for (size_t a = 0; a < rows.size(); a++) if (rows[a].key == i) { rows.erase(rows.begin() + a); return; }
    }
    void remove_column(int k)   /* SPM:394-412 */
    {
        for (SparseRow& r : rows)
            for (size_t a = 0; a < r.items.size(); a++)
                if (r.items[a].first == k) { r.items.erase(r.items.begin() + a); break; }
        for (size_t a = 0; a < rows.size();)
            if (rows[a].items.empty()) rows.erase(rows.begin() + a); else a++;
    }
    void remove_at(int i, int k)   /* SPM:367-379 */
    {
        SparseRow* r = find_row(i);
        if (!r) return;
        for (size_t a = 0; a < r->items.size(); a++)
            if (r->items[a].first == k) { r->items.erase(r->items.begin() + a); break; }
        if (r->items.empty()) remove_row(i);
    }
    /* SPM:592-628: rows/cols sorted ascending, re-indexed */
    Sparse compact(std::vector<int>& orows, std::vector<int>& ocols) const
    {
        orows.clear(); ocols.clear();
        for (const SparseRow& r : rows) {
            if (r.items.empty()) continue;
            orows.push_back(r.key);
            for (const auto& it : r.items) ocols.push_back(it.first);
        }
        std::sort(orows.begin(), orows.end());
        orows.erase(std::unique(orows.begin(), orows.end()), orows.end());
        std::sort(ocols.begin(), ocols.end());
        ocols.erase(std::unique(ocols.begin(), ocols.end()), ocols.end());
        Sparse out((int)orows.size(), (int)ocols.size(), defval);
        for (const SparseRow& r : rows)
            for (const auto& it : r.items) {
                int ri = (int)(std::lower_bound(orows.begin(), orows.end(), r.key) - orows.begin());
                int ci = (int)(std::lower_bound(ocols.begin(), ocols.end(), it.first) - ocols.begin());
                out.set(ri, ci, it.second);
            }
        out.height = (int)orows.size();
        out.width  = (int)ocols.size();
        return out;
    }
};

/* GC:358-425.  Component extraction order = order of complete.Any (SPM:187-213): the first
 * remaining row in insertion order (D5).  The DFS below is the reference's, stack for stack. */
std::vector<Sparse> connected_components(const Sparse& original)
{
    Sparse complete(original);
    std::vector<Sparse> components;

    while (complete.count() > 0) {
        std::vector<int> rowstack, colstack;
        std::vector<char> rowvisited(complete.height, 0), colvisited(complete.width, 0);
        Sparse component(complete.height, complete.width, original.defval);

        rowstack.push_back(complete.rows[0].key);
        colstack.push_back(complete.rows[0].items[0].first);

        while (!rowstack.empty() || !colstack.empty()) {
            while (!rowstack.empty()) {
                int nextrow = rowstack.back();
                rowstack.pop_back();
                if (rowvisited[nextrow]) continue;
                rowvisited[nextrow] = 1;
                const SparseRow* rp = complete.find_row(nextrow);
                std::vector<std::pair<int, double> > row;
                if (rp) row = rp->items;
                for (const auto& it : row) component.set(nextrow, it.first, it.second);   /* AddRow */
                complete.remove_row(nextrow);
                for (const auto& it : row)
                    if (!colvisited[it.first] &&
                        std::find(colstack.begin(), colstack.end(), it.first) == colstack.end())
                        colstack.push_back(it.first);
            }
            while (!colstack.empty()) {
                int nextcol = colstack.back();
                colstack.pop_back();
                if (colvisited[nextcol]) continue;
                colvisited[nextcol] = 1;
                std::vector<std::pair<int, double> > col;   /* SPM:231-243 */
                for (const SparseRow& r : complete.rows)
                    for (const auto& it : r.items)
                        if (it.first == nextcol) col.push_back(std::make_pair(r.key, it.second));
                for (const auto& it : col) component.set(it.first, nextcol, it.second);   /* AddColumn */
                complete.remove_column(nextcol);
                for (const auto& it : col)
                    if (!rowvisited[it.first] &&
                        std::find(rowstack.begin(), rowstack.end(), it.first) == rowstack.end())
                        rowstack.push_back(it.first);
            }
        }
        components.push_back(component);
    }
    return components;
}

/* GC:183-197 */
double assignment_value(const Sparse& profit, const std::vector<int>& matches, bool isnull)
{
    if (isnull) return -kInf;
    double total = 0;
    for (size_t i = 0; i < matches.size(); i++) total += profit.get((int)i, matches[i]);
    return total;
}

/* GC:64-175.  Returns false for "no solution" (null). */
bool hungarian(const Sparse& matrix, std::vector<int>& matchx)
{
    const int nx = matrix.width, ny = matrix.height;
    std::vector<double> labelx(nx, 0.0), labely(ny, 0.0), slack(ny);
    std::vector<int> matchy(nx, -1), parent(ny);
    std::vector<char> visitx(nx), visity(ny);
    matchx.assign(nx, -1);

    for (const SparseRow& r : matrix.rows) {   /* FoldRows(Math.Max, 0) (SPM:464-478) */
        double folded = 0;
        for (const auto& it : r.items) folded = std::fmax(folded, it.second);
        if (r.key < nx) labelx[r.key] = folded;
    }

    int root;
    while (true) {
        root = -1;
        for (int i = 0; i < nx; i++) if (matchx[i] == -1) { root = i; break; }
        if (root == -1) break;

        for (int i = 0; i < ny; i++) parent[i] = root;
        for (int i = 0; i < ny; i++) slack[i] = labelx[root] + labely[i] - matrix.get(root, i);
        std::fill(visitx.begin(), visitx.end(), 0);
        std::fill(visity.begin(), visity.end(), 0);
        visitx[root] = 1;

        int iminslack = -1;
        bool found = false;
        while (!found) {
            iminslack = -1;
            double delta = kInf;
            for (int i = 0; i < ny; i++)
                if (!visity[i] && slack[i] < delta) { iminslack = i; delta = slack[i]; }
            if (std::isinf(delta) && delta > 0) return false;

            for (int i = 0; i < nx; i++) if (visitx[i]) labelx[i] -= delta;
            for (int i = 0; i < ny; i++) { if (visity[i]) labely[i] += delta; else slack[i] -= delta; }

            visity[iminslack] = 1;
            if (matchy[iminslack] != -1) {
                int match = matchy[iminslack];
                visitx[match] = 1;
                for (int i = 0; i < ny; i++)
                    if (!visity[i]) {
                        double mdelta = labelx[match] + labely[i] - matrix.get(match, i);
                        if (mdelta < slack[i]) { slack[i] = mdelta; parent[i] = match; }
                    }
            }
            else found = true;
        }

        int px, py, ty;
        for (py = iminslack, px = parent[py]; px != root; py = ty, px = parent[py]) {
            ty = matchx[px];
            matchx[px] = py;
            matchy[py] = px;
        }
        matchx[px] = py;
        matchy[py] = px;
    }
    return true;
}

struct MKey { int i, k; };
struct MurtyNode {
    std::vector<MKey> forced, eliminated;
    std::vector<int>  assignment;
    bool null_assignment;
    MurtyNode() : null_assignment(true) {}
};

inline bool key_in(const std::vector<MKey>& v, int i, int k)
{
    for (const MKey& e : v) if (e.i == i && e.k == k) return true;
    return false;
}

/* GC:469-509 */
std::vector<MurtyNode> murty_children(const MurtyNode& node)
{
    std::vector<MurtyNode> children;
    if (node.null_assignment) return children;
    std::vector<MKey> irem;
    for (size_t i = 0; i < node.assignment.size(); i++)
        if (!key_in(node.forced, (int)i, node.assignment[i])) irem.push_back(MKey{(int)i, node.assignment[i]});
    for (int i = 0; i + 1 < (int)irem.size(); i++) {
        if (key_in(node.forced, irem[i].i, irem[i].k)) continue;
        MurtyNode child;
        child.eliminated = node.eliminated;
        child.eliminated.push_back(irem[i]);
        child.forced = node.forced;
        for (int k = 0; k < i; k++) child.forced.push_back(irem[k]);
        children.push_back(child);
    }
    return children;
}

/* GC:206-234 */
Sparse reduce_profit(const Sparse& full, const MurtyNode& reducer)
{
    Sparse reduced(full);
    reduced.defval = -kInf;
    for (const MKey& f : reducer.forced) reduced.remove_row(f.i);
    for (const MKey& f : reducer.forced) reduced.remove_column(f.k);
    int w = reduced.width, h = reduced.height;
    for (const MKey& f : reducer.forced) reduced.set(f.i, f.k, 1);
    reduced.width = std::max(w, reduced.width);
    reduced.height = std::max(h, reduced.height);
    for (const MKey& e : reducer.eliminated) reduced.remove_at(e.i, e.k);
    return reduced;
}

/* Pull-style enumerators so SetLogLikelihood can stop early exactly like the C# foreach/yield. */
struct PairingEnumerator {
    virtual ~PairingEnumerator() {}
    virtual bool next(std::vector<int>& perm, double& value) = 0;
};

/* GC:280-350 */
struct LexEnumerator : PairingEnumerator {
    const Sparse& profit;
    std::vector<int> perm;
    int measurestart;
    bool started;
    LexEnumerator(const Sparse& p, int modelsize) : profit(p), started(false)
    {
        for (const SparseRow& r : profit.rows) perm.push_back(r.key);
        std::sort(perm.begin(), perm.end());
        measurestart = (int)perm.size();
        for (size_t i = 0; i < perm.size(); i++)
            if (perm[i] >= modelsize) { measurestart = (int)i; break; }
    }
    static bool last(const std::vector<int>& p)
    {
        for (size_t i = 1; i < p.size(); i++) if (p[i - 1] < p[i]) return false;
        return true;
    }
    bool next(std::vector<int>& out, double& value)
    {
        if (!started) {
            started = true;
            std::reverse(perm.begin() + measurestart, perm.end());
            out = perm;
            value = assignment_value(profit, perm, false);
            return true;
        }
        if (last(perm)) return false;
        int a, b, n = (int)perm.size();
        for (a = n - 2; a > 0; a--) if (perm[a] < perm[a + 1]) break;
        for (b = n - 1; b > a; b--) if (perm[a] < perm[b]) break;
        std::swap(perm[a], perm[b]);
        std::reverse(perm.begin() + a + 1, perm.end());
        std::reverse(perm.begin() + measurestart, perm.end());
        out = perm;
        value = assignment_value(profit, perm, false);
        return true;
    }
};

/* GC:241-272 with the list-backed PriorityQueue of GC:595-707 (D4: stable re-sort) */
struct MurtyEnumerator : PairingEnumerator {
    const Sparse& profit;
    std::vector<std::pair<double, MurtyNode> > frontier;   /* ascending; Pop takes the back */
    std::vector<MurtyNode> pending_children;
    bool have_best;
    MurtyNode best;
    void add(double pr, const MurtyNode& n)
    {
        size_t pos = frontier.size();
        while (pos > 0 && (frontier[pos - 1].first - pr) > 0) pos--;   /* stable ascending */
        frontier.insert(frontier.begin() + pos, std::make_pair(pr, n));
    }
    MurtyEnumerator(const Sparse& p) : profit(p), have_best(false)
    {
        MurtyNode first;
        first.null_assignment = !hungarian(profit, first.assignment);
        add(assignment_value(profit, first.assignment, first.null_assignment), first);
    }
    void expand()
    {
        for (MurtyNode child : murty_children(best)) {
            Sparse red = reduce_profit(profit, child);
            child.null_assignment = !hungarian(red, child.assignment);
            if (!child.null_assignment) add(assignment_value(profit, child.assignment, false), child);
        }
    }
    bool next(std::vector<int>& out, double& value)
    {
        if (have_best) { expand(); have_best = false; }   /* code after the previous yield */
        if (frontier.empty()) return false;
        value = frontier.back().first;
        best  = frontier.back().second;
        frontier.pop_back();
        have_best = true;
        out = best.assignment;
        return true;
    }
};

/* MX:361-389 */
double log_sum_exp(const double* v, int begin, int end)
{
    double mx = -kInf, value = 0;
    for (int i = begin; i < end; i++) mx = std::fmax(mx, v[i]);
    if (std::isinf(mx) && mx < 0) return -kInf;
    for (int i = begin; i < end; i++) value += std::exp(v[i] - mx);
    return mx + std::log(value);
}

/* PHD:415-453; quasi = the matrix of QuasiSetLogLikelihood (PHD:561-640): every landmark fully visible
 * (zprobs weight 1, log PD / log(1 - PD) constants) and the association gate d < 12 instead of d < 5 */
Sparse set_loglike_matrix(const orc_config* c, const double* pose, int J, const double* jm, int M,
                          const double* z, bool quasi = false)
{
    const int dz = meas_dim(c);
    Sparse logprobs(J + M, J + M, -kInf);
    double logclutter = std::log(c->clutter);
    std::vector<Gaussian> zprobs(J);
    for (int i = 0; i < J; i++) {
        double ml[3];
        measure_perfect(c, pose, jm + 3 * i, ml);
        zprobs[i] = make_gaussian(ml, c->R, quasi ? 1.0 : detection_probability_m(c, ml), dz);
    }
    if (quasi) {
        const double logPD = std::log(c->pd), log1PD = std::log(1 - c->pd);
        for (int i = 0; i < J; i++)
            for (int k = 0; k < M; k++) {
                double d = mahalanobis(zprobs[i], z + 3 * k);
                if (d < 12) logprobs.set(i, k, logPD + std::log(zprobs[i].mult) - 0.5 * d * d);
            }
        for (int i = 0; i < J; i++) logprobs.set(i, M + i, log1PD);
    }
    else {
        for (int i = 0; i < J; i++)
            for (int k = 0; k < M; k++) {
                double d = mahalanobis(zprobs[i], z + 3 * k);
                if (d < 5) logprobs.set(i, k, std::log(zprobs[i].w) + std::log(zprobs[i].mult) - 0.5 * d * d);
            }
        for (int i = 0; i < J; i++) logprobs.set(i, M + i, std::log(1 - zprobs[i].w));
    }
    for (int i = 0; i < M; i++) logprobs.set(J + i, i, logclutter);
    return logprobs;
}

/* PHD:462-515, including the stale-buffer early exit (quirk A9.4); quasi: PHD:561-713 without the gradient */
double set_loglikelihood(const orc_config* c, const double* pose, int J, const double* jm, int M,
                         const double* z, bool quasi = false)
{
    Sparse llmatrix = set_loglike_matrix(c, pose, J, jm, M, z, quasi);
    std::vector<Sparse> connected = connected_components(llmatrix);
    double logcomp[200];
    for (int i = 0; i < 200; i++) logcomp[i] = 0;
    double total = 0;

    for (size_t ci = 0; ci < connected.size(); ci++) {
        std::vector<int> rows, cols;
        Sparse component = connected[ci].compact(rows, cols);
        for (size_t k = 0; k < rows.size(); k++)
            if (rows[k] >= J)
                for (size_t h = 0; h < cols.size(); h++)
                    if (cols[h] >= M) component.set((int)k, (int)h, 0);

        PairingEnumerator* en;
        bool enumerateall;
        if (component.rows.size() <= 5) { en = new LexEnumerator(component, J); enumerateall = true; }
        else                            { en = new MurtyEnumerator(component);  enumerateall = false; }

        int m = 0;
        std::vector<int> perm;
        double value;
        while (en->next(perm, value)) {
            if (m >= 200 || (!enumerateall && logcomp[m] - logcomp[0] < -10)) break;
            logcomp[m] = value;
            m++;
        }
        delete en;
        total += log_sum_exp(logcomp, 0, m);
    }
    return total;
}

/* MX:400-440 TemperedAverage: softmax-like average of the vectors with log-weights weights[begin, end) -- which it
 * overwrites IN PLACE with exp(w - max) (the caller's logcomp buffer), then normalises with Accord's Normalize():
 * the EUCLIDEAN norm of the whole array, stale entries beyond `end` included (the same extension method makes unit
 * vectors in QuaternionTest.cs:103 and Manipulator.cs:720).  nvec = length of the arrays (200). */
static int g_tempered_sum = 0;
void tempered_average(const std::vector<std::vector<double> >& vectors, double* weights, int nvec, int begin, int end,
                      int dim, double* value)
{
    for (int a = 0; a < dim; a++) value[a] = 0;
    double mx = -kInf;
    for (int i = begin; i < end; i++) mx = std::fmax(mx, weights[i]);
    if (std::isinf(mx) && mx < 0) return;
    for (int i = begin; i < end; i++) weights[i] = std::exp(weights[i] - mx);
    double norm = 0;
    if (g_tempered_sum) { for (int i = 0; i < nvec; i++) norm += weights[i]; }   /* experiment switch, see orc_set_tempered_norm */
    else { for (int i = 0; i < nvec; i++) norm += weights[i] * weights[i]; norm = std::sqrt(norm); }
    for (int i = begin; i < end; i++) {
        double wi = (norm == 0) ? weights[i] : weights[i] / norm;
        for (int a = 0; a < dim; a++) value[a] = value[a] + wi * vectors[i][a];
    }
}

/* PHD:561-713 with calcgradient = true: the value (its Murty early exit reads the buffer TemperedAverage mutated, so
 * it can differ from the value-only overload) and the pose gradient (OdoSize entries) */
double quasi_set_loglikelihood_gradient(const orc_config* c, const double* pose, int J, const double* jm, int M,
                                        const double* z, double* gradient)
{
    const int dz = meas_dim(c), od = odo_size(c);
    Sparse llmatrix(J + M, J + M, -kInf);
    const double logPD = std::log(c->pd), log1PD = std::log(1 - c->pd), logclutter = std::log(c->clutter);
    std::vector<Gaussian> zprobs(J);
    std::vector<double> zjac(18 * (size_t)std::max(J, 1));
    for (int i = 0; i < J; i++) {
        double ml[3];
        measure_perfect(c, pose, jm + 3 * i, ml);
        zprobs[i] = make_gaussian(ml, c->R, 1.0, dz);
    }
    for (int i = 0; i < J; i++) jacobian_p(c, pose, jm + 3 * i, &zjac[18 * (size_t)i]);
    std::vector<std::vector<double> > dlldp((size_t)J * std::max(M, 1));   /* [i * M + k], empty = zeros */
    for (int i = 0; i < J; i++)
        for (int k = 0; k < M; k++) {
            const double* m = z + 3 * k;
            double d = mahalanobis(zprobs[i], m);
            if (d < 12) {
                llmatrix.set(i, k, logPD + std::log(zprobs[i].mult) - 0.5 * d * d);
                double diff[3], row[3];
                for (int a = 0; a < dz; a++) diff[a] = m[a] - zprobs[i].m[a];
                for (int b = 0; b < dz; b++) {          /* (1 x dz) . Sigma^-1 */
                    double sum = 0;
                    for (int a = 0; a < dz; a++) sum += diff[a] * zprobs[i].Pinv[a * 3 + b];
                    row[b] = sum;
                }
                std::vector<double> g(od);
                for (int l = 0; l < od; l++) {          /* . zjacobians[i] */
                    double sum = 0;
                    for (int b = 0; b < dz; b++) sum += row[b] * zjac[18 * (size_t)i + b * 6 + l];
                    g[l] = sum;
                }
                dlldp[(size_t)i * M + k] = g;
            }
        }
    for (int i = 0; i < J; i++) llmatrix.set(i, M + i, log1PD);
    for (int i = 0; i < M; i++) llmatrix.set(J + i, i, logclutter);

    std::vector<Sparse> connected = connected_components(llmatrix);
    double logcomp[200];
    for (int i = 0; i < 200; i++) logcomp[i] = 0;
    std::vector<std::vector<double> > dlogcompdp(200);
    double total = 0;
    for (int a = 0; a < od; a++) gradient[a] = 0;

    for (size_t ci = 0; ci < connected.size(); ci++) {
        std::vector<int> rows, cols;
        Sparse component = connected[ci].compact(rows, cols);
        for (size_t k = 0; k < rows.size(); k++)
            if (rows[k] >= J)
                for (size_t h = 0; h < cols.size(); h++)
                    if (cols[h] >= M) component.set((int)k, (int)h, 0);
        PairingEnumerator* en;
        bool enumerateall;
        if (component.rows.size() <= 5) { en = new LexEnumerator(component, J); enumerateall = true; }
        else                            { en = new MurtyEnumerator(component);  enumerateall = false; }
        int m = 0;
        std::vector<int> perm;
        double value;
        while (en->next(perm, value)) {
            if (m >= 200 || (!enumerateall && logcomp[m] - logcomp[0] < -10)) break;
            logcomp[m] = value;
            dlogcompdp[m].assign(od, 0.0);
            for (size_t p = 0; p < perm.size(); p++) {
                int r = rows[p], q = cols[perm[p]];
                if (r < J && q < M && !dlldp[(size_t)r * M + q].empty())
                    for (int a = 0; a < od; a++) dlogcompdp[m][a] = dlogcompdp[m][a] + dlldp[(size_t)r * M + q][a];
            }
            m++;
        }
        delete en;
        total += log_sum_exp(logcomp, 0, m);
        double avg[6];
        for (int i = m; i < 200; i++) if (dlogcompdp[i].empty()) dlogcompdp[i].assign(od, 0.0);
        tempered_average(dlogcompdp, logcomp, 200, 0, m, od, avg);
        for (int a = 0; a < od; a++) gradient[a] = gradient[a] + avg[a];
    }
    return total;
}

/* PHD:373-393 */
void weight_alpha(const orc_config* c, const double* pose, int M, const double* z, const Map& predicted,
                  const Map& corrected, double* out)
{
    std::vector<double> cw(corrected.size());
    for (size_t i = 0; i < corrected.size(); i++) cw[i] = corrected[i].w;
    std::vector<int> picks = best_map_estimate((int)corrected.size(), cw.data());
    std::vector<double> jm(3 * picks.size() + 3);
    for (size_t j = 0; j < picks.size(); j++)
        for (int a = 0; a < 3; a++) jm[3 * j + a] = corrected[picks[j]].m[a];

    double plog = 0, clog = 0;
    for (size_t j = 0; j < picks.size(); j++) {
        plog += std::log(map_evaluate(predicted, &jm[3 * j]));
        clog += std::log(map_evaluate(corrected, &jm[3 * j]));
    }
    double pcount = expected_size(predicted);
    double ccount = expected_size(corrected);
    double setll  = set_loglikelihood(c, pose, (int)picks.size(), jm.data(), M, z);
    double ratio  = (plog - pcount) - (clog - ccount);
    out[0] = std::exp(setll + ratio);
    out[1] = setll; out[2] = plog; out[3] = clog; out[4] = pcount; out[5] = ccount;
    out[6] = (double)picks.size();
}

/* ------------------------------------------------------------------ */
/* helpers for the C interface                                         */
/* ------------------------------------------------------------------ */
Map map_from_arrays(int n, const double* w, const double* m, const double* P)
{
    Map map;
    map.reserve(n);
    for (int i = 0; i < n; i++) map.push_back(make_gaussian(m + 3 * i, P + 9 * i, w[i]));
    return map;
}
int map_to_arrays(const Map& map, int cap, double* w, double* m, double* P)
{
    int n = (int)map.size();
    for (int i = 0; i < n && i < cap; i++) {
        w[i] = map[i].w;
        std::memcpy(m + 3 * i, map[i].m, 3 * sizeof(double));
        std::memcpy(P + 9 * i, map[i].P, 9 * sizeof(double));
    }
    return n;
}
Sparse sparse_from_dense(int h, int w, const double* val, const uint8_t* defined, double defval)
{
    Sparse s(h, w, defval);
    for (int i = 0; i < h; i++)
        for (int k = 0; k < w; k++)
            if (defined[i * w + k]) s.set(i, k, val ? val[i * w + k] : 1.0);
    s.height = h; s.width = w;
    return s;
}

/* UTIL:173-202: lower Cholesky root C (C C^T = Q, diagonal floored at 1e-40), y = C g */
void cholesky6(const double* Qin, double* C)
{
    double Q[36];
    std::memcpy(Q, Qin, sizeof Q);
    for (int i = 0; i < 6; i++) if (Q[i * 6 + i] < 1e-40) Q[i * 6 + i] = 1e-40;
    for (int i = 0; i < 36; i++) C[i] = 0;
    for (int j = 0; j < 6; j++) {
        double s = Q[j * 6 + j];
        for (int k = 0; k < j; k++) s -= C[j * 6 + k] * C[j * 6 + k];
        C[j * 6 + j] = std::sqrt(s);
        for (int i = j + 1; i < 6; i++) {
            double t = Q[i * 6 + j];
            for (int k = 0; k < j; k++) t -= C[i * 6 + k] * C[j * 6 + k];
            C[i * 6 + j] = t / C[j * 6 + j];
        }
    }
}

}  // namespace

/* ------------------------------------------------------------------ */
/* whole navigator                                                     */
/* ------------------------------------------------------------------ */
struct orc_nav {
    orc_config cfg;
    int P;
    int only_mapping;
    std::vector<double> poses;     /* P x 7 */
    std::vector<double> weights;   /* P */
    std::vector<double> alphas;    /* P, last WeightAlpha values */
    std::vector<Map>    maps;
    int best;
    double chol[36];
};

namespace {

/* body of the Parallel.For of PHD:326-339 for one particle */
void particle_map_update(orc_nav* nav, int i, int M, const double* z)
{
    const orc_config* c = &nav->cfg;
    const double* pose = &nav->poses[7 * i];
    Map predicted = predict_conditional(c, pose, nav->maps[i], M, z, nullptr);
    Map corrected = correct_conditional(c, pose, predicted, M, z, c->density_distance_threshold);
    corrected = prune_model(c, corrected);
    if (!nav->only_mapping) {
        double out[7];
        weight_alpha(c, pose, M, z, predicted, corrected, out);
        nav->alphas[i] = out[0];
        nav->weights[i] *= out[0];
    }
    nav->maps[i] = corrected;
}

void parallel_particles(orc_nav* nav, int first, int last, int M, const double* z)
{
    int nt = std::max(1, nav->cfg.nthreads);
    nt = std::min(nt, std::max(1, last - first));
    if (nt == 1) {
        for (int i = first; i < last; i++) particle_map_update(nav, i, M, z);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
        pool.emplace_back([=]() { for (int i = first + t; i < last; i += nt) particle_map_update(nav, i, M, z); });
    for (std::thread& th : pool) th.join();
}

/* PHD:343-358 + PHD:724-777 on a bare weight vector */
int normalize_resample(const orc_config* c, int P, double* w, double u, int* best, int* ancestors,
                       int force)
{
    int b = *best;
    double maxweight = 0;
    for (int i = 0; i < P; i++) ancestors[i] = i;
    if (force != 2) {   /* force == 2: ResampleParticles() alone, weights taken as they are */
        double sum = 0;
        for (int i = 0; i < P; i++) sum += w[i];
        sum = (sum == 0) ? 1 : sum;
        for (int i = 0; i < P; i++) w[i] = w[i] / sum;

        for (int i = 0; i < P; i++) if (w[i] > maxweight) { maxweight = w[i]; b = i; }

        double cum = 0;
        for (int i = 0; i < P; i++) cum += w[i] * w[i];
        bool depleted = (1.0 / cum < c->min_effective_particle * P);
        if (!(depleted || force)) { *best = b; return 0; }
    }

    /* PHD:724-760 */
    double random = u / P;
    std::vector<double> old(w, w + P);
    maxweight = 0;
    for (int i = 0, k = 0; i < P; i++) {
        for (; random > 0 && k < P; k++) random -= old[k];
        int a = (k == 0) ? 0 : k - 1;   /* k == 0 only if u == 0 (the C# would throw) */
        ancestors[i] = a;
        w[i] = 1.0 / P;
        random += 1.0 / P;
        if (old[a] > maxweight) { maxweight = old[a]; b = i; }
    }
    *best = b;
    return 1;
}

}  // namespace

extern "C" {

void orc_quat_mul(const double a[4], const double b[4], double out[4])
{
    Quat r = qmul(Quat{a[0], a[1], a[2], a[3]}, Quat{b[0], b[1], b[2], b[3]});
    out[0] = r.w; out[1] = r.x; out[2] = r.y; out[3] = r.z;
}
void orc_quat_exp(const double lie[3], double out[4])
{
    Quat r = qexp(lie);
    out[0] = r.w; out[1] = r.x; out[2] = r.y; out[3] = r.z;
}
void orc_quat_log(const double q[4], double out[3]) { qlog(Quat{q[0], q[1], q[2], q[3]}, out); }
void orc_quat_sqrt(const double q[4], double out[4])
{
    Quat r = qsqrt(Quat{q[0], q[1], q[2], q[3]});
    out[0] = r.w; out[1] = r.x; out[2] = r.y; out[3] = r.z;
}
void orc_quat_from_ypr(double yaw, double pitch, double roll, double out[4])   /* QUAT:254-273 */
{
    double y2 = 0.5 * yaw, p2 = 0.5 * pitch, r2 = 0.5 * roll;
    double sy = std::sin(y2), cy = std::cos(y2), sp = std::sin(p2), cp = std::cos(p2);
    double sr = std::sin(r2), cr = std::cos(r2);
    out[0] = cy * cp * cr + sy * sp * sr;
    out[1] = cy * sp * cr + sy * cp * sr;
    out[2] = sy * cp * cr - cy * sp * sr;
    out[3] = cy * cp * sr - sy * sp * cr;
}
void orc_quat_to_matrix(const double q[4], double out[9]) { qtomatrix(Quat{q[0], q[1], q[2], q[3]}, out); }
void orc_quat_vector_rotator(const double from[3], const double to[3], double out[4])   /* QUAT:281-284 */
{
    double cx = from[1] * to[2] - from[2] * to[1];
    double cy = from[2] * to[0] - from[0] * to[2];
    double cz = from[0] * to[1] - from[1] * to[0];
    Quat r = qnormalize(Quat{1 + dot3(from, to), cx, cy, cz});
    out[0] = r.w; out[1] = r.x; out[2] = r.y; out[3] = r.z;
}
void orc_pose_from_state(const double state[7], double out[7]) { pose_store(pose_from_state(state), out); }
void orc_pose_add_odometry(const double pose[7], const double delta[6], double out[7])
{
    pose_store(add_odometry(pose_load(pose), delta), out);
}
void orc_pose_diff_odometry(const double pose[7], const double origin[7], double out[6])
{
    diff_odometry(pose_load(pose), pose_load(origin), out);
}
void orc_measure_perfect(const orc_config* c, const double* pose, const double m[3], double out[3])
{
    measure_perfect(c, pose, m, out);
}
void orc_measurement_jacobian_l(const orc_config* c, const double* pose, const double m[3], double out[9])
{
    jacobian_l(c, pose, m, out);
}
void orc_measure_to_map(const orc_config* c, const double* pose, const double z[3], double out[3])
{
    measure_to_map(c, pose, z, out);
}
double orc_fuzzy_visible(const orc_config* c, const double z[3]) { return fuzzy_visible(c, z); }

/* PRM:221-243 */
void orc_fit_to_measurement(const orc_config* c, const double pose0[7], const double z[3],
                            const double landmark[3], double out[7])
{
    Pose p = pose_load(pose0);
    double diff[3] = {landmark[0] - p.t[0], landmark[1] - p.t[1], landmark[2] - p.t[2]};
    double rot[9], ll[3], ml[3];
    qtomatrix(qconj(p.q), rot);
    matvec(rot, 3, 3, 3, diff, ll);
    double invf = 1.0 / c->measurer[0];
    ml[2] = z[2] / std::sqrt(1 + (z[0] * z[0] + z[1] * z[1]) * invf * invf);
    ml[0] = z[0] * ml[2] * invf;
    ml[1] = z[1] * ml[2] * invf;
    double nl = euclid(ll, 3), nm = euclid(ml, 3);
    double a[3] = {ll[0] / nl, ll[1] / nl, ll[2] / nl}, b[3] = {ml[0] / nm, ml[1] / nm, ml[2] / nm};
    double ar[4];
    orc_quat_vector_rotator(a, b, ar);
    Quat rotation = qmul(qconj(Quat{ar[0], ar[1], ar[2], ar[3]}), p.q);
    double rm[9], rl[3];
    qtomatrix(rotation, rm);
    matvec(rm, 3, 3, 3, ml, rl);
    Pose r{{landmark[0] - rl[0], landmark[1] - rl[1], landmark[2] - rl[2]}, rotation};
    pose_store(r, out);
}

double orc_gaussian_evaluate(const double m[3], const double P[9], const double x[3])
{
    return evaluate(make_gaussian(m, P, 1.0), x);
}
void orc_gaussian_merge(int n, const double* w, const double* m, const double* P, double* ow, double om[3],
                        double oP[9])
{
    Map comps = map_from_arrays(n, w, m, P);
    std::vector<const Gaussian*> ptrs;
    for (const Gaussian& g : comps) ptrs.push_back(&g);
    Gaussian r = merge(ptrs);
    *ow = r.w;
    std::memcpy(om, r.m, sizeof r.m);
    std::memcpy(oP, r.P, sizeof r.P);
}

int orc_predict(const orc_config* c, const double* pose, int n, const double* w, const double* m,
                const double* P, int M, const double* z, int cap, double* ow, double* om, double* oP,
                int* nbirth)
{
    return map_to_arrays(predict_conditional(c, pose, map_from_arrays(n, w, m, P), M, z, nbirth), cap, ow,
                         om, oP);
}
int orc_correct(const orc_config* c, const double* pose, int n, const double* w, const double* m,
                const double* P, int M, const double* z, double gate_radius, int cap, double* ow,
                double* om, double* oP)
{
    return map_to_arrays(correct_conditional(c, pose, map_from_arrays(n, w, m, P), M, z, gate_radius), cap,
                         ow, om, oP);
}
int orc_prune(const orc_config* c, int n, const double* w, const double* m, const double* P, int cap,
              double* ow, double* om, double* oP)
{
    return map_to_arrays(prune_model(c, map_from_arrays(n, w, m, P)), cap, ow, om, oP);
}
int orc_best_map_estimate(int n, const double* w, int cap, int* picks)
{
    std::vector<int> p = best_map_estimate(n, w);
    for (size_t i = 0; i < p.size() && (int)i < cap; i++) picks[i] = p[i];
    return (int)p.size();
}
double orc_set_loglikelihood(const orc_config* c, const double* pose, int J, const double* jm, int M,
                             const double* z)
{
    return set_loglikelihood(c, pose, J, jm, M, z);
}
void orc_set_depth_frame(const float* depth_xy, int resx, int resy)
{
    if (!depth_xy) { g_depth = nullptr; g_depth_store.clear(); return; }
    g_depth_store.assign(depth_xy, depth_xy + (size_t)resx * resy);
    g_depth = g_depth_store.data(); g_resx = resx; g_resy = resy;
}
double orc_quasi_set_loglikelihood(const orc_config* c, const double* pose, int J, const double* jm, int M,
                                   const double* z)
{
    return set_loglikelihood(c, pose, J, jm, M, z, true);
}
double orc_quasi_set_loglikelihood_gradient(const orc_config* c, const double* pose, int J, const double* jm, int M,
                                            const double* z, double* gradient)
{
    return quasi_set_loglikelihood_gradient(c, pose, J, jm, M, z, gradient);
}
void orc_set_tempered_norm(int sum_instead_of_euclidean) { g_tempered_sum = sum_instead_of_euclidean; }
void orc_measurement_jacobian_p(const orc_config* c, const double* pose, const double m[3], double out[18])
{
    jacobian_p(c, pose, m, out);
}
int orc_set_loglike_matrix(const orc_config* c, const double* pose, int J, const double* jm, int M, const double* z,
                           int quasi, int cap, int* rows, int* cols, double* vals)
{
    Sparse mx = set_loglike_matrix(c, pose, J, jm, M, z, quasi != 0);
    int n = 0;
    for (const SparseRow& r : mx.rows)
        for (const auto& it : r.items) {
            if (n < cap) { rows[n] = r.key; cols[n] = it.first; vals[n] = it.second; }
            n++;
        }
    return n;
}
void orc_weight_alpha(const orc_config* c, const double* pose, int M, const double* z, int np,
                      const double* pw, const double* pm, const double* pP, int nc, const double* cw,
                      const double* cm, const double* cP, double out[7])
{
    weight_alpha(c, pose, M, z, map_from_arrays(np, pw, pm, pP), map_from_arrays(nc, cw, cm, cP), out);
}
int orc_normalize_resample(const orc_config* c, int P, double* weights, double u, int* best,
                           int* ancestors, int force_resample)
{
    return normalize_resample(c, P, weights, u, best, ancestors, force_resample);
}

int orc_hungarian(int n, const double* val, const uint8_t* defined, double defval, int* match)
{
    Sparse s = sparse_from_dense(n, n, val, defined, defval);
    std::vector<int> mx;
    bool ok = hungarian(s, mx);
    if (!ok) return 0;
    for (int i = 0; i < n; i++) match[i] = mx[i];
    return 1;
}
/* postanalysis/Plot.cs:531-581 (OSPA) over GC:52-55, 64-175 (one global Hungarian) and SPM:526-542 (Apply also maps
 * the default value, so undefined entries cost C^P).  a, b: na x 3 and nb x 3 landmark positions. */
double orc_ospa(int na, const double* a, int nb, const double* b, double C, double P, double* cardinality)
{
    if (na > nb) { std::swap(na, nb); std::swap(a, b); }
    if (na == 0) { *cardinality = (nb == 0) ? 0.0 : C; return *cardinality; }
    Sparse transport(nb, nb, 0.0);
    const double CP = std::pow(C, P);
    for (int i = 0; i < na; i++)
        for (int k = 0; k < nb; k++) {
            double d[3] = {a[3 * i] - b[3 * k], a[3 * i + 1] - b[3 * k + 1], a[3 * i + 2] - b[3 * k + 2]};
            double distance = std::pow(std::fmin(C, euclid(d, 3)), P);   /* Plot.cs:583-586 */
            if (CP - distance > 1e-5) transport.set(i, k, CP - distance);
        }
    transport.height = nb; transport.width = nb;
    std::vector<int> best;
    if (!hungarian(transport, best)) { *cardinality = kInf; return kInf; }
    double total = 0;
    for (int i = 0; i < nb; i++) {
        total += CP - transport.get(i, best[i]);
    }
    *cardinality = C * std::pow((double)(nb - na) / nb, 1.0 / P);
    return std::pow(total / nb, 1.0 / P);
}

/* SIMV:243-295 with the caller's random numbers: uniforms[i] per landmark, gauss[3 i ..] per landmark (used only when
 * detected), chol = lower root of R (UTIL:173-202), clutter_u[3 k ..] for the nc clutter points (PRM:249-256).
 * Returns the number of measurements; assoc = landmark index or INT_MIN. */
int orc_generate_measurements(const orc_config* c, const double* pose, int n, const double* landmarks,
                              const double* uniforms, const double* gauss, const double* chol, int nc,
                              const double* clutter_u, double* z, int* assoc)
{
    int count = 0;
    for (int i = 0; i < n; i++) {
        double mp[3];
        measure_perfect(c, pose, landmarks + 3 * i, mp);
        double pd = detection_probability_m(c, mp);
        if (pd > 0 && uniforms[i] < pd) {
            for (int r = 0; r < 3; r++) {
                double sum = 0;
                for (int k = 0; k < 3; k++) sum += chol[r * 3 + k] * gauss[3 * i + k];
                z[3 * count + r] = mp[r] + (0.0 + sum);
            }
            assoc[count++] = i;
        }
    }
    const int left = (int)c->measurer[3], top = (int)c->measurer[4], width = (int)c->measurer[5],
              height = (int)c->measurer[6];
    const float rminf = (float)c->measurer[1], length = (float)c->measurer[2] - rminf;   /* AForge.Range: floats */
    for (int k = 0; k < nc; k++) {
        z[3 * count]     = clutter_u[3 * k] * width + left;
        z[3 * count + 1] = clutter_u[3 * k + 1] * height + top;
        z[3 * count + 2] = clutter_u[3 * k + 2] * (double)length + (double)rminf;
        assoc[count++] = std::numeric_limits<int>::min();
    }
    return count;
}

int orc_connected_components(int h, int w, const uint8_t* defined)
{
    Sparse s = sparse_from_dense(h, w, nullptr, defined, 0.0);
    return (int)connected_components(s).size();
}
int orc_lexicographical(int n, const double* val, const uint8_t* defined, double defval, int modelsize,
                        int cap, int* perms, double* values)
{
    Sparse s = sparse_from_dense(n, n, val, defined, defval);
    LexEnumerator en(s, modelsize);
    std::vector<int> perm;
    double value;
    int count = 0;
    while (en.next(perm, value)) {
        if (count < cap) {
            for (int i = 0; i < n; i++) perms[count * n + i] = perm[i];
            values[count] = value;
        }
        count++;
    }
    return count;
}
int orc_murty(int n, const double* val, const uint8_t* defined, double defval, int cap, int* perms,
              double* values)
{
    Sparse s = sparse_from_dense(n, n, val, defined, defval);
    MurtyEnumerator en(s);
    std::vector<int> perm;
    double value;
    int count = 0;
    while (en.next(perm, value)) {
        if (count < cap) {
            for (int i = 0; i < n; i++) perms[count * n + i] = (i < (int)perm.size()) ? perm[i] : -1;
            values[count] = value;
        }
        count++;
        if (count >= cap) break;
    }
    return count;
}
int orc_murty_children(int n, const int* assignment, int nf, const int* forced, int ne, const int* elim,
                       int cap, int* out)
{
    MurtyNode node;
    node.null_assignment = false;
    node.assignment.assign(assignment, assignment + n);
    for (int i = 0; i < nf; i++) node.forced.push_back(MKey{forced[2 * i], forced[2 * i + 1]});
    for (int i = 0; i < ne; i++) node.eliminated.push_back(MKey{elim[2 * i], elim[2 * i + 1]});
    std::vector<MurtyNode> ch = murty_children(node);
    int pos = 0;
    for (const MurtyNode& c : ch) {
        int need = 2 + 2 * (int)(c.forced.size() + c.eliminated.size());
        if (pos + need > cap) break;
        out[pos++] = (int)c.forced.size();
        out[pos++] = (int)c.eliminated.size();
        for (const MKey& k : c.forced) { out[pos++] = k.i; out[pos++] = k.k; }
        for (const MKey& k : c.eliminated) { out[pos++] = k.i; out[pos++] = k.k; }
    }
    return (int)ch.size();
}

/* PHD:192-208, 245-266 */
orc_nav* orc_nav_new(const orc_config* c, int P, const double* pose, int only_mapping)
{
    orc_nav* nav = new orc_nav();
    nav->cfg = *c;
    nav->only_mapping = only_mapping;
    if (only_mapping && P < 1) P = 1;
    nav->P = P;
    nav->poses.resize(7 * P);
    for (int i = 0; i < P; i++) std::memcpy(&nav->poses[7 * i], pose, 7 * sizeof(double));
    nav->weights.assign(P, 1.0 / P);
    nav->alphas.assign(P, 1.0);
    nav->maps.resize(P);
    nav->best = 0;
    cholesky6(c->Q, nav->chol);
    return nav;
}
void orc_nav_delete(orc_nav* nav) { delete nav; }
void orc_nav_set_map(orc_nav* nav, int i, int n, const double* w, const double* m, const double* P)
{
    nav->maps[i] = map_from_arrays(n, w, m, P);
}
int orc_nav_get_map(orc_nav* nav, int i, int cap, double* w, double* m, double* P)
{
    return map_to_arrays(nav->maps[i], cap, w, m, P);
}
void orc_nav_set_pose(orc_nav* nav, int i, const double* pose)
{
    std::memcpy(&nav->poses[7 * i], pose, 7 * sizeof(double));
}
void orc_nav_get_poses(orc_nav* nav, double* poses)
{
    std::memcpy(poses, nav->poses.data(), nav->poses.size() * sizeof(double));
}
void orc_nav_set_weights(orc_nav* nav, const double* w) { nav->weights.assign(w, w + nav->P); }
void orc_nav_get_weights(orc_nav* nav, double* w)
{
    std::memcpy(w, nav->weights.data(), nav->P * sizeof(double));
}
void orc_nav_get_alphas(orc_nav* nav, double* a) { std::memcpy(a, nav->alphas.data(), nav->P * sizeof(double)); }
int orc_nav_particle_count(orc_nav* nav) { return nav->P; }

/* PHD:295-314 -> TRK:89-102 -> SIMV:190-202 / VEH:325-336 -> POSE:314-333.
 * The first six draws of VEH:330-333 only perturb OdometryPose (never read by the filter)
 * and are not modelled; gauss holds the second block (TRK:95-97). */
void orc_nav_update(orc_nav* nav, const double reading[6], double dt, const double* gauss, int perfect_still)
{
    if (nav->only_mapping) return;   /* PHD:297-300: the caller sets particle 0's pose */
    bool zero = true;
    for (int i = 0; i < 6; i++) if (reading[i] != 0) zero = false;
    for (int p = 0; p < nav->P; p++) {
        Pose pose = add_odometry(pose_load(&nav->poses[7 * p]), reading);
        if (!(perfect_still && zero)) {
            double cg[6], noise[6];
            matvec(nav->chol, 6, 6, 6, gauss + 6 * p, cg);
            for (int i = 0; i < 6; i++) noise[i] = dt * (0.0 + cg[i]);
            pose = add_odometry(pose, noise);
        }
        pose_store(pose, &nav->poses[7 * p]);
    }
}

void orc_nav_map_update_range(orc_nav* nav, int M, const double* z, int first, int last)
{
    parallel_particles(nav, first, last, M, z);
}

/* PHD:323-362 */
void orc_nav_slam_update(orc_nav* nav, int M, const double* z, double u, int* best, int* resampled,
                         int* ancestors)
{
    parallel_particles(nav, 0, nav->P, M, z);
    std::vector<int> anc(nav->P);
    for (int i = 0; i < nav->P; i++) anc[i] = i;
    int res = 0;
    if (!nav->only_mapping) {
        res = normalize_resample(&nav->cfg, nav->P, nav->weights.data(), u, &nav->best, anc.data(), 0);
        if (res) {
            std::vector<double> poses(nav->poses.size());
            std::vector<Map> maps(nav->P);
            for (int i = 0; i < nav->P; i++) {
                std::memcpy(&poses[7 * i], &nav->poses[7 * anc[i]], 7 * sizeof(double));
                maps[i] = nav->maps[anc[i]];
            }
            nav->poses.swap(poses);
            nav->maps.swap(maps);
        }
    }
    if (best) *best = nav->best;
    if (resampled) *resampled = res;
    if (ancestors) std::memcpy(ancestors, anc.data(), nav->P * sizeof(int));
}

}  // extern "C"
