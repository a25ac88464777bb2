// rbphd_analysis.cu -- sm_100a kernels next to the navigator's hot path:
//   k_generate_measurements   SimulatedVehicle.Measure (SIMV:211-218, 243-295) with host-supplied random numbers
//   k_ospa                    the OSPA map metric (postanalysis/Plot.cs:531-581): optimal assignment of the
//                             thresholded landmark-distance matrix (GraphCombinatorics.LinearAssignment, GC:52-175)
// Neither is on the frame path; they let large accuracy experiments run without a C# host in the loop.
#include "rbphd_analysis.cuh"

#include <limits.h>

namespace rbphd {

__global__ void __launch_bounds__(kBlock) k_generate_measurements(DevCfg cfg, const double* pose7,
                                                                 const double* landmarks, int n,
                                                                 const double* uniforms, const double* gauss,
                                                                 const double* chol9, const double* clutter_u, int nc,
                                                                 double* z, int* assoc, int* count)
{
    __shared__ BlockShared sh;
    const Pose pose = pose_load(pose7);
    int base = 0;
    for (int start = 0; start < n; start += kBlock) {
        const int i = start + threadIdx.x;
        bool hit = false;
        double zi[3] = {0, 0, 0};
        if (i < n) {
            const double m[3] = {landmarks[3 * i], landmarks[3 * i + 1], landmarks[3 * i + 2]};
            double diff[3], mp[3];
            Quat local;
            to_local(pose, m, diff, local);
            measure_from_local(cfg, diff, local, mp);
            const double pd = detection_probability(cfg, mp);       // SIMV:324-339
            if (pd > 0 && uniforms[i] < pd) {                       // SIMV:255-258
                hit = true;
                // MeasureDetected (SIMV:211-218): h(x, m) + C g, C C^T = R (UTIL:173-202)
                const double* g = gauss + 3 * (size_t)i;
                for (int r = 0; r < 3; r++) {
                    double sum = 0;
                    for (int k = 0; k < 3; k++) sum += chol9[r * 3 + k] * g[k];
                    zi[r] = mp[r] + (0.0 + sum);
                }
            }
        }
        int total;
        const int slot = base + block_excl_scan(sh, hit ? 1 : 0, &total);
        if (hit) {
            z[3 * (size_t)slot] = zi[0]; z[3 * (size_t)slot + 1] = zi[1]; z[3 * (size_t)slot + 2] = zi[2];
            assoc[slot] = i;
        }
        base += total;
        __syncthreads();
    }
    // clutter (PRM:249-256): uniform over the film and the range clip
    for (int k = threadIdx.x; k < nc; k += kBlock) {
        const size_t o = (size_t)base + k;
        z[3 * o] = clutter_u[3 * k] * (cfg.right - cfg.left) + cfg.left;
        z[3 * o + 1] = clutter_u[3 * k + 1] * (cfg.bottom - cfg.top) + cfg.top;
        z[3 * o + 2] = clutter_u[3 * k + 2] * (double)((float)cfg.rmax - (float)cfg.rmin) + cfg.rmin;
        assoc[o] = INT_MIN;
    }
    if (threadIdx.x == 0) *count = base + nc;
}

void launch_generate_measurements(cudaStream_t s, const DevCfg& cfg, const double* pose7, const double* landmarks, int n,
                                  const double* uniforms, const double* gauss, const double* chol9,
                                  const double* clutter_u, int nc, double* z, int* assoc, int* count)
{
    k_generate_measurements<<<1, kBlock, 0, s>>>(cfg, pose7, landmarks, n, uniforms, gauss, chol9, clutter_u, nc, z, assoc,
                                                 count);
}

// ------------------------------------------------------------------------------------------------
// OSPA.  profit[i][k] = C^P - min(C, |a_i - b_k|)^P where that exceeds 1e-5, else 0 (rows i >= na: all 0), and the
// assignment that maximises the total profit (Kuhn-Munkres with slack arrays; the column loops of every step run
// across the CTA, the minimum by a block reduction).  Any optimal assignment gives the same metric.
// ------------------------------------------------------------------------------------------------
size_t ospa_workspace_doubles(int nb) { return (size_t)nb * nb + 8 * (size_t)nb + 16; }

__global__ void __launch_bounds__(kBlock) k_ospa(const double* a, int na, const double* b, int nb, double C, double P,
                                                double* work, double* out2)
{
    __shared__ double s_val[kWarps];
    __shared__ int s_idx[kWarps];
    __shared__ double s_delta;
    __shared__ int s_col, s_root, s_done;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = nb;
    double* profit = work;
    double* labelx = work + (size_t)n * n;
    double* labely = labelx + n;
    double* slack = labely + n;
    int* matchx = reinterpret_cast<int*>(slack + n);
    int* matchy = matchx + n;
    int* parent = matchy + n;
    int* visitx = parent + n;
    int* visity = visitx + n;
    const double CP = pow(C, P);
    if (na == 0) {
        if (tid == 0) { out2[1] = (nb == 0) ? 0.0 : C; out2[0] = out2[1]; }
        return;
    }
    // profit matrix and the initial feasible labelling (row maxima)
    for (int i = warp; i < n; i += kWarps) {
        double rmax = 0;
        for (int k = lane; k < n; k += 32) {
            double v = 0;
            if (i < na) {
                const double dx = a[3 * i] - b[3 * k], dy = a[3 * i + 1] - b[3 * k + 1], dz = a[3 * i + 2] - b[3 * k + 2];
                double s = 0;
                s += dx * dx; s += dy * dy; s += dz * dz;
                const double dist = pow(fmin(C, sqrt(s)), P);
                if (CP - dist > 1e-5) v = CP - dist;
            }
            profit[(size_t)i * n + k] = v;
            rmax = fmax(rmax, v);
        }
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        if (lane == 0) { labelx[i] = rmax; matchx[i] = -1; }
    }
    for (int k = tid; k < n; k += kBlock) { labely[k] = 0; matchy[k] = -1; }
    __syncthreads();

    for (int root = 0; root < n; root++) {
        if (matchx[root] != -1) continue;   // (uniform: matchx is only written between barriers)
        for (int k = tid; k < n; k += kBlock) {
            slack[k] = labelx[root] + labely[k] - profit[(size_t)root * n + k];
            parent[k] = root; visity[k] = 0; visitx[k] = 0;
        }
        __syncthreads();
        if (tid == 0) { visitx[root] = 1; s_done = 0; }
        __syncthreads();
        while (true) {
            // column with the smallest slack outside the tree (lowest index among equals)
            double bv = INFINITY;
            int bi = 0x7fffffff;
            for (int k = tid; k < n; k += kBlock)
                if (!visity[k] && (slack[k] < bv || (slack[k] == bv && k < bi))) { bv = slack[k]; bi = k; }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < kWarps; w++)
                    if (s_val[w] < bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
                s_delta = bv; s_col = bi;
            }
            __syncthreads();
            const double delta = s_delta;
            const int col = s_col;
            if (delta > 0) {   // move the labels so that the column becomes tight
                for (int k = tid; k < n; k += kBlock) {
                    if (visitx[k]) labelx[k] -= delta;
                    if (visity[k]) labely[k] += delta; else slack[k] -= delta;
                }
            }
            __syncthreads();
            if (tid == 0) {
                visity[col] = 1;
                if (matchy[col] == -1) {   // augment along the tree
                    int y = col;
                    while (true) {
                        const int x = parent[y];
                        const int prev = matchx[x];
                        matchx[x] = y; matchy[y] = x;
                        if (x == root) break;
                        y = prev;
                    }
                    s_done = 1;
                }
                else { s_root = matchy[col]; visitx[s_root] = 1; }
            }
            __syncthreads();
            if (s_done) break;
            const int x = s_root;
            for (int k = tid; k < n; k += kBlock) {
                if (visity[k]) continue;
                const double sl = labelx[x] + labely[k] - profit[(size_t)x * n + k];
                if (sl < slack[k]) { slack[k] = sl; parent[k] = x; }
            }
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) {
        double total = 0;   // AssignmentValue of the cost matrix C^P - profit (GC:183-197), rows in order
        for (int i = 0; i < n; i++) total += CP - profit[(size_t)i * n + matchx[i]];
        out2[1] = C * pow((double)(nb - na) / nb, 1.0 / P);
        out2[0] = pow(total / nb, 1.0 / P);
    }
}

void launch_ospa(cudaStream_t s, const double* a, int na, const double* b, int nb, double c, double p, double* work,
                 double* out2)
{
    k_ospa<<<1, kBlock, 0, s>>>(a, na, b, nb, c, p, work, out2);
}

}  // namespace rbphd
