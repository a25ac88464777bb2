// rbphd_math.cuh -- FP64 device arithmetic of the pixel-range PHD update.
//
// Compiled with -fmad=false: every +,-,*,/ and sqrt rounds once, in the operation order of the
// reference's scalar C# (citations: abbreviations of SURVEY.md), so that values agree with the
// CPU oracle to the last bit except where exp/log/sin/cos differ by an ulp between libms.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace rbphd {

struct DevCfg {
    double R[9];
    double Rinv[9];
    double multR;          // (2 pi)^-1 / sqrt(det R)   (GAUSS:155 with dim 3)
    double logmultR;
    double chol[36];       // lower Cholesky root of Q (UTIL:183-191)
    double pd, clutter, logclutter;
    double birth_cov[9];
    double birth_w;
    double min_w;
    double merge_t;        // MergeThreshold (compared as d2 < t*t, GAUSS:245)
    double explore_thr;
    double gate_r2;        // correct gate (PHD:882), in the metric's units
    double gate_r;         // Euclidean radius bound of the correct gate (for culling)
    double explore_r2;     // explore gate (PHD:958)
    double explore_r;
    double min_eff;
    double ramp[3];
    double focal, left, right, top, bottom, rmin, rmax;
    int    maxq;
    int    ungated;        // stage tests: gate_radius < 0
    // KinectMeasurer (KinectMeasurer.cs:123-173): the current depth frame, depth[x * resy + y], or null
    const float* depth;
    int    resx, resy;
    double half_x, half_y; // ResX / 2, ResY / 2 (float divisions, promoted)
    float  rmin_f;         // RangeClip.Min as the float it is
};

struct Quat { double w, x, y, z; };

__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b)   // QUAT:295-301
{
    Quat r;
    r.w = a.w * b.w - (a.x * b.x + a.y * b.y + a.z * b.z);
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__device__ __forceinline__ Quat qconj(const Quat& q) { return Quat{q.w, -q.x, -q.y, -q.z}; }
__device__ __forceinline__ Quat qnormalize(const Quat& q)   // QUAT:240-245
{
    double mag = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    double a = 1 / mag;
    return Quat{a * q.w, a * q.x, a * q.y, a * q.z};
}
__device__ __forceinline__ double euclid3(double x, double y, double z)
{
    double s = 0;
    s += x * x; s += y * y; s += z * z;
    return sqrt(s);
}
__device__ inline Quat qexp(const double* lie)   // QUAT:185-196
{
    double phi = euclid3(lie[0], lie[1], lie[2]);
    if (phi < 1e-12) return Quat{1, 0, 0, 0};
    double s = sin(phi);
    return Quat{cos(phi), s * (lie[0] / phi), s * (lie[1] / phi), s * (lie[2] / phi)};
}
__device__ __forceinline__ Quat qsqrt(const Quat& q)   // QUAT:225-235
{
    if (fabs(q.w - -1.0) < 1e-8) return Quat{1, 0, 0, 0};
    double rw = sqrt(0.5 * (1 + q.w));
    double alpha = 1 / (2 * rw);
    return Quat{rw, alpha * q.x, alpha * q.y, alpha * q.z};
}
__device__ __forceinline__ void qtomatrix(const Quat& q, double* m)   // QUAT:327-342
{
    double xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    double xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w;
    double yz = q.y * q.z, yw = q.y * q.w, zw = q.z * q.w;
    m[0] = 1 - 2 * (yy + zz); m[1] = 2 * (xy - zw);     m[2] = 2 * (xz + yw);
    m[3] = 2 * (xy + zw);     m[4] = 1 - 2 * (xx + zz); m[5] = 2 * (yz - xw);
    m[6] = 2 * (xz - yw);     m[7] = 2 * (yz + xw);     m[8] = 1 - 2 * (xx + yy);
}

struct Pose { double t[3]; Quat q; };

__device__ __forceinline__ Pose pose_load(const double* s)
{
    return Pose{{s[0], s[1], s[2]}, Quat{s[3], s[4], s[5], s[6]}};
}
__device__ __forceinline__ void pose_store(const Pose& p, double* s)
{
    s[0] = p.t[0]; s[1] = p.t[1]; s[2] = p.t[2];
    s[3] = p.q.w;  s[4] = p.q.x;  s[5] = p.q.y;  s[6] = p.q.z;
}

__device__ inline Pose add_odometry(const Pose& p, const double* d)   // POSE:314-333
{
    double half[3] = {0.5 * d[3], 0.5 * d[4], 0.5 * d[5]};
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

// local = q* (0,diff) q : landmark in the camera frame (first step of PRM:138-149 / 157-177)
__device__ __forceinline__ void to_local(const Pose& p, const double* m, double* diff, Quat& local)
{
    diff[0] = m[0] - p.t[0]; diff[1] = m[1] - p.t[1]; diff[2] = m[2] - p.t[2];
    local = qmul(qmul(qconj(p.q), Quat{0, diff[0], diff[1], diff[2]}), p.q);
}

// PRM:138-149 given local/diff
__device__ __forceinline__ void measure_from_local(const DevCfg& c, const double* diff, const Quat& local,
                                                   double* mp)
{
    double sgn = (local.z > 0) ? 1.0 : ((local.z < 0) ? -1.0 : 0.0);
    mp[2] = sgn * euclid3(diff[0], diff[1], diff[2]);
    mp[0] = c.focal * local.x / local.z;
    mp[1] = c.focal * local.y / local.z;
}

// PRM:277-291 (+ SIMV:324-339: times detectionProbability); with a depth frame attached the occlusion-aware
// KinectMeasurer.FuzzyVisibleM (KinectMeasurer.cs:151-173): a landmark behind the measured surface is invisible
__device__ __forceinline__ double detection_probability(const DevCfg& c, const double* z)
{
    double mind = INFINITY;
    mind = fmin(mind, (z[0] - c.left) / c.ramp[0]);
    mind = fmin(mind, (c.right - z[0]) / c.ramp[0]);
    mind = fmin(mind, (z[1] - c.top) / c.ramp[1]);
    mind = fmin(mind, (c.bottom - z[1]) / c.ramp[1]);
    mind = fmin(mind, (z[2] - c.rmin) / c.ramp[2]);
    mind = fmin(mind, (c.rmax - z[2]) / c.ramp[2]);
    mind = fmax(0.0, fmin(1.0, mind));
    if (c.depth != nullptr) {
        if (mind == 0) return 0.0;                       // outside of the image region: no depth to look up
        const int x = (int)(z[0] + c.half_x), y = (int)(z[1] + c.half_y);
        if (x < 0 || x >= c.resx || y < 0 || y >= c.resy) return 0.0;
        const float d = c.depth[(size_t)x * c.resy + y];
        if (d != d) return 0.0;
        const float range = (float)z[2];
        mind = fmin(mind, (double)(range - c.rmin_f) / c.ramp[2]);
        mind = fmin(mind, (double)(d - range) / c.ramp[2]);
        mind = fmax(0.0, fmin(1.0, mind));
    }
    return mind * c.pd;
}

// PRM:299-312
__device__ __forceinline__ void measure_to_map(const DevCfg& c, const Pose& p, const double* z, double* out)
{
    double px = z[0], py = z[1], range = z[2];
    double alpha = range / sqrt(c.focal * c.focal + px * px + py * py);
    Quat r = qmul(qmul(p.q, Quat{0, alpha * px, alpha * py, alpha * c.focal}), qconj(p.q));
    out[0] = p.t[0] + r.x; out[1] = p.t[1] + r.y; out[2] = p.t[2] + r.z;
}

// camera-frame back-projection (particle independent part of PRM:305-306)
__device__ __forceinline__ void measure_to_camera(const DevCfg& c, const double* z, double* v)
{
    double px = z[0], py = z[1], range = z[2];
    double alpha = range / sqrt(c.focal * c.focal + px * px + py * py);
    v[0] = alpha * px; v[1] = alpha * py; v[2] = alpha * c.focal;
}

// ---- 3x3 helpers (row-major; sums run k = 0,1,2 from 0) ----
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
            s += A[i * 3 + 0] * B[0 * 3 + j];
            s += A[i * 3 + 1] * B[1 * 3 + j];
            s += A[i * 3 + 2] * B[2 * 3 + j];
            C[i * 3 + j] = s;
        }
}
// C = A * B^T
__device__ __forceinline__ void mat3_mul_bt(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
            s += A[i * 3 + 0] * B[j * 3 + 0];
            s += A[i * 3 + 1] * B[j * 3 + 1];
            s += A[i * 3 + 2] * B[j * 3 + 2];
            C[i * 3 + j] = s;
        }
}
__device__ __forceinline__ void mat3_vec(const double* A, const double* x, double* y)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double s = 0;
        s += A[i * 3 + 0] * x[0];
        s += A[i * 3 + 1] * x[1];
        s += A[i * 3 + 2] * x[2];
        y[i] = s;
    }
}
// closed-form inverse (adjugate * (1/det)); returns det.  Same formula as the oracle (D1).
__device__ __forceinline__ double mat3_inv(const double* a, double* inv)
{
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
// diff^T Ainv diff with Accord's order: InnerProduct(diff, Ainv.Multiply(diff)) (GAUSS:203)
__device__ __forceinline__ double quadform3(const double* Ainv, const double* d)
{
    double t[3];
    mat3_vec(Ainv, d, t);
    double s = 0;
    s += d[0] * t[0]; s += d[1] * t[1]; s += d[2] * t[2];
    return s;
}

#define RBPHD_INV_TWO_PI (1.0 / (2 * 3.14159265358979323846))

// Gaussian multiplier (GAUSS:155, integer division -3/2 = -1)
__device__ __forceinline__ double gauss_mult(double det) { return RBPHD_INV_TWO_PI / sqrt(det); }

// PRM:157-177: H = Jproj(local) * R(q*)
__device__ __forceinline__ void jacobian_l(const DevCfg& c, const Pose& p, const Quat& l, double* H)
{
    double mag = ((l.z > 0) ? 1 : -1) * sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
    double jp[9];
    jp[0] = c.focal / l.z; jp[1] = 0;             jp[2] = -c.focal * l.x / (l.z * l.z);
    jp[3] = 0;             jp[4] = c.focal / l.z; jp[5] = -c.focal * l.y / (l.z * l.z);
    jp[6] = l.x / mag;     jp[7] = l.y / mag;     jp[8] = l.z / mag;
    double jr[9];
    qtomatrix(qconj(p.q), jr);
    mat3_mul(jp, jr, H);
}

// order-preserving key of a non-negative double, inverted so an ascending sort gives weight-descending
__device__ __forceinline__ unsigned long long weight_desc_key(double w)
{
    if (!(w > 0)) w = 0.0;   // -0, NaN -> +0
    return ~(unsigned long long)__double_as_longlong(w);
}
// order-preserving key of any finite double (ascending)
__device__ __forceinline__ unsigned long long double_asc_key(double x)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

}  // namespace rbphd
