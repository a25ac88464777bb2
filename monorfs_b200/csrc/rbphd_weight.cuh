// rbphd_weight.cuh -- WeightAlpha (PHD:373-393) for one particle, CTA-wide:
//   BestMapEstimate (MAP:119-142), Map.Evaluate over the predicted and corrected maps (MAP:192-202),
//   SetLogLikeMatrix / SetLogLikelihood (PHD:415-515) with connected components (GC:358-425),
//   exhaustive lexicographic pairing (GC:280-350) and log-sum-exp (MX:361-389).
// Included by rbphd_kernels.cu (needs Smem / Slab / KParams).
#pragma once

namespace rbphd {

// Terms of Map.Evaluate with Mahalanobis distance^2 above kEvalD2 are < 2e-22 of the component's peak and
// are skipped (the reference sums them; the parity bar for weights is 1e-9 relative).
constexpr double kQueryCell = 0.6;   // cell edge of the grid over the map-estimate points

struct CompSrc {
    const double* w;
    const double* mx; const double* my; const double* mz;
    const double* cov;     // field-major covariance (a * covstride + i), valid for i < ncov
    size_t covstride;
    int ncov;
    const double* defcov;  // covariance of components i >= ncov (births)
    int n;
};

__device__ __forceinline__ void comp_cov(const CompSrc& c, int i, double* P)
{
    if (i < c.ncov) {
#pragma unroll
        for (int a = 0; a < 9; a++) P[a] = c.cov[(size_t)a * c.covstride + i];
    }
    else {
#pragma unroll
        for (int a = 0; a < 9; a++) P[a] = c.defcov[a];
    }
}

// w_i N(x; m_i, P_i) with the inverse recomputed (only used by the rare all-terms fallback)
__device__ inline double eval_term_full(const CompSrc& c, int i, double x, double y, double z)
{
    double P[9], Pinv[9];
    comp_cov(c, i, P);
    double det = mat3_inv(P, Pinv);
    double d[3] = {x - c.mx[i], y - c.my[i], z - c.mz[i]};
    return c.w[i] * (gauss_mult(det) * exp(-0.5 * quadform3(Pinv, d)));
}

// v[t] = sum_i w_i N(jm_t; m_i, P_i) for the J points in s.jm (MAP:192-202), component-major: each thread
// owns components, looks up the query points inside the component's influence radius in the cell grid
// over the J points (sm.ctx.grid / sm.gstart() / s.gitems, built by the caller) and accumulates into
// vs[t] with double atomics.  Returns sum_t ln v[t].
__device__ double eval_map_at_points(const KParams& p, Smem& sm, const Slab& s, const CompSrc& c, int J,
                                     double* vs, double* erad, bool erad_ready)
{
    const int tid = threadIdx.x, capj = p.lay.cap_j;
    const double* jx = s.jm; const double* jy = s.jm + capj; const double* jz = s.jm + 2 * capj;
    const CellGrid& g = sm.ctx.grid;
    // cell-ordered single-precision copy of the points (x, y, z, index) in shared memory, made by the caller
    const bool qsm = (J * 2 <= kVsCap);
    const float4* qf = reinterpret_cast<const float4*>(sm.vs());
    for (int t = tid; t < J; t += kBlock) vs[t] = 0.0;
    __syncthreads();
    PHASE_MARK(sm, 20);
    // per-component cull radius (eval_radius2): in a frame it was computed where the covariances were in
    // registers (A2 / A5 for the predicted map, B6 for the corrected one); the stage entry point makes it here
    if (!erad_ready) {
        for (int i = tid; i < c.n; i += kBlock) {
            double P[9];
            comp_cov(c, i, P);
            erad[i] = eval_radius2(P);
        }
    }
    __syncthreads();
    // w_i N(jm_t; m_i, P_i) -> vs[t]; the inverse is formed per term (a few thousand terms per particle: cheaper
    // than keeping an inverse per component in the slab)
    auto term_into = [&](int i, int t) {
        const double d[3] = {jx[t] - c.mx[i], jy[t] - c.my[i], jz[t] - c.mz[i]};   // x - Mean (GAUSS:201)
        double P[9], Pinv[9];
        comp_cov(c, i, P);
        const double mult = gauss_mult(mat3_inv(P, Pinv));
        atomicAdd(&vs[t], c.w[i] * (mult * exp(-0.5 * quadform3(Pinv, d))));
    };
    const int fatcap = p.M;
    if (tid == 0) sm.ctx.nU = 0;
    __syncthreads();
    enumerate_then_process(
        sm, c.n, reinterpret_cast<uint2*>(sm.skey()), (int)p.smem_sort_cap, reinterpret_cast<uint2*>(s.edst),
        p.lay.cap_edges / 2,
        [&](int i, auto emit) {
            const double x = c.mx[i], y = c.my[i], z = c.mz[i];
            const double r2 = erad[i];
            bool brute = !(r2 >= 0) || isinf(r2);   // NaN covariance: never cull
            int lo[3], hi[3];
            if (!brute) {
                if (!grid_range(g, x, y, z, sqrt(r2), lo, hi)) return;
                long cells = (long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
                if (cells > 2048) brute = true;
                else if (cells > 27) {   // wide component: a whole warp walks its cells later (balance)
                    const int f = atomicAdd(&sm.ctx.nU, 1);
                    if (f < fatcap) { sm.kidx()[f] = i; return; }
                }
            }
            if (brute) {
                for (int t = 0; t < J; t++) {
                    const double dx = jx[t] - x, dy = jy[t] - y, dz = jz[t] - z;
                    if (!(r2 >= 0) || dx * dx + dy * dy + dz * dz <= r2) emit(i, t);
                }
                return;
            }
            const float xf = (float)x, yf = (float)y, zf = (float)z;
            const float r2f = (float)r2 * 1.001f + 1e-4f;   // generous: an extra far term is harmless
            for (int cz = lo[2]; cz <= hi[2]; cz++)
                for (int cy = lo[1]; cy <= hi[1]; cy++) {
                    const int rowc = (cz * g.dim[1] + cy) * g.dim[0];
                    const int b = sm.gstart()[rowc + lo[0]], e = sm.gstart()[rowc + hi[0] + 1];
                    if (qsm) {
                        for (int q = b; q < e; q++) {
                            const float4 v = qf[q];
                            const float dx = v.x - xf, dy = v.y - yf, dz = v.z - zf;
                            if (dx * dx + dy * dy + dz * dz <= r2f) emit(i, __float_as_int(v.w));
                        }
                    }
                    else {
                        for (int q = b; q < e; q++) {
                            const int t = s.gitems[q];
                            const double dx = jx[t] - x, dy = jy[t] - y, dz = jz[t] - z;
                            if (dx * dx + dy * dy + dz * dz <= r2) emit(i, t);
                        }
                    }
                }
        },
        term_into);
    {
        const int nfat = min(sm.ctx.nU, fatcap);
        if (tid == 0) sm.ctx.dbg[7] += sm.ctx.nU;
        const int lane = tid & 31, warp = tid >> 5;
        uint2* list = reinterpret_cast<uint2*>(sm.skey());
        uint2* ovf = reinterpret_cast<uint2*>(s.edst);
        const int list_cap = (int)p.smem_sort_cap, ovf_cap = p.lay.cap_edges / 2;
        for (int fbase = 0; fbase < nfat; fbase += kWarps) {
            if (tid == 0) sm.ctx.nsel2 = 0;
            __syncthreads();
            const int f = fbase + warp;
            if (f < nfat) {
                const int i = sm.kidx()[f];
                const double x = c.mx[i], y = c.my[i], z = c.mz[i];
                const double r2 = erad[i];
                int lo[3], hi[3];
                if (grid_range(g, x, y, z, sqrt(r2), lo, hi)) {
                    const int ny = hi[1] - lo[1] + 1, rows = ny * (hi[2] - lo[2] + 1);
                    for (int rw = lane; rw < rows; rw += 32) {
                        const int cz = lo[2] + rw / ny, cy = lo[1] + rw % ny;
                        const int rowc = (cz * g.dim[1] + cy) * g.dim[0];
                        const int b = sm.gstart()[rowc + lo[0]], e = sm.gstart()[rowc + hi[0] + 1];
                        for (int q = b; q < e; q++) {
                            int t;
                            bool in;
                            if (qsm) {
                                const float4 v = qf[q];
                                const float dx = v.x - (float)x, dy = v.y - (float)y, dz = v.z - (float)z;
                                in = dx * dx + dy * dy + dz * dz <= (float)r2 * 1.001f + 1e-4f;
                                t = __float_as_int(v.w);
                            }
                            else {
                                t = s.gitems[q];
                                const double dx = jx[t] - x, dy = jy[t] - y, dz = jz[t] - z;
                                in = dx * dx + dy * dy + dz * dz <= r2;
                            }
                            if (in) {
                                // one atomic per group of lanes that hit together (thousands of hits per particle
                                // on one shared-memory counter serialise otherwise)
                                const unsigned act = __activemask();
                                const int leader = __ffs(act) - 1;
                                int idx = 0;
                                if (lane == leader) idx = atomicAdd(&sm.ctx.nsel2, __popc(act));
                                idx = __shfl_sync(act, idx, leader) + __popc(act & ((1u << lane) - 1u));
                                if (idx < list_cap) list[idx] = make_uint2((unsigned)i, (unsigned)t);
                                else if (idx - list_cap < ovf_cap) ovf[idx - list_cap] = make_uint2((unsigned)i, (unsigned)t);
                                else sm.ctx.status |= ST_OVER_PAIRS;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            const int tot = sm.ctx.nsel2;
            const int cnt = min(tot, list_cap);
            for (int e = tid; e < cnt; e += kBlock) term_into((int)list[e].x, (int)list[e].y);
            const int nov = min(max(tot - list_cap, 0), ovf_cap);
            for (int e = tid; e < nov; e += kBlock) term_into((int)ovf[e].x, (int)ovf[e].y);
            __syncthreads();
        }
    }
    __syncthreads();
    PHASE_MARK(sm, 21);
    // nothing nearby at all: the reference's full sum decides between a denormal and log(0) = -inf
    double lsum = 0;
    for (int t = tid; t < J; t += kBlock) {
        double v = vs[t];
        if (!(v > 0)) {
            v = 0;
            for (int i = 0; i < c.n; i++) v += eval_term_full(c, i, jx[t], jy[t], jz[t]);
        }
        lsum += log(v);
    }
    double tot = block_sum(sm.sh, lsum);
    __syncthreads();
    PHASE_MARK(sm, 22);
    return tot;
}

// MX:361-389
__device__ __forceinline__ double log_sum_exp(const double* v, int n)
{
    double mx = -INFINITY, value = 0;
    for (int i = 0; i < n; i++) mx = fmax(mx, v[i]);
    if (isinf(mx) && mx < 0) return -INFINITY;
    for (int i = 0; i < n; i++) value += exp(v[i] - mx);
    return mx + log(value);
}


// landmarks (ascending) and measurements (ascending) of the edge group [e, end); -1 if more than maxn rows
__device__ inline int collect_block(const unsigned long long* skey, int e, int end, int maxn, int* ts, int* ks,
                                    int* pa, int* pb)
{
    int a = 0, b = 0;
    for (int q = e; q < end; q++) {
        int t = (int)((skey[q] >> 20) & 0xfffff), k = (int)(skey[q] & 0xfffff);
        bool ft = false, fk = false;
        for (int i = 0; i < a; i++) ft |= (ts[i] == t);
        for (int i = 0; i < b; i++) fk |= (ks[i] == k);
        if (!ft) { if (a + b >= maxn) return -1; ts[a++] = t; }
        if (!fk) {
            if (a + b >= maxn) return -1;
            int i = b++;
            while (i > 0 && ks[i - 1] > k) { ks[i] = ks[i - 1]; i--; }
            ks[i] = k;
        }
    }
    *pa = a; *pb = b;
    return a + b;
}

// the compacted block (SPM:592-628): rows = landmarks then clutter rows, columns = measurements then miss
// columns, detection edges from the gate, clutter x miss quadrant filled with 0 (PHD:480-488)
__device__ inline void fill_block(const DevCfg& c, const Slab& s, const unsigned long long* skey,
                                  const unsigned int* sval, int e, int end, const int* ts, const int* ks, int a,
                                  int b, double* Mx, int ld)
{
    const int n = a + b;
    for (int r = 0; r < n; r++)
        for (int cc = 0; cc < n; cc++) {
            double v;
            if (r < a) v = (cc >= b && cc - b == r) ? log(1 - s.jpd[ts[r]]) : -INFINITY;
            else       v = (cc < b) ? ((cc == r - a) ? c.logclutter : -INFINITY) : 0.0;
            Mx[r * ld + cc] = v;
        }
    for (int q = e; q < end; q++) {
        int t = (int)((skey[q] >> 20) & 0xfffff), k = (int)(skey[q] & 0xfffff);
        int r = 0, cc = 0;
        while (ts[r] != t) r++;
        while (ks[cc] != k) cc++;
        Mx[r * ld + cc] = s.llval[sval[q]];
    }
}

__device__ inline int lexicographical_values(const double* Mx, int n, int modelsize, double* vals,
                                             signed char* perms = nullptr);

// logcomp[m] as the reference's shared 200-entry buffer holds it when the block starting at edge `head`
// begins (quirk A9.4): the m-th value of the most recent earlier block that produced more than m values
// (blocks without detection edges only ever write index 0 and come last), else the initial 0.
__device__ double stale_value(const KParams& p, const Slab& s, MurtyWork& mw, const unsigned long long* skey,
                              const unsigned int* sval, int head, int nbig_done, int J, int m)
{
    int pos = head;
    while (pos > 0) {
        unsigned long long lab = skey[pos - 1] >> 40;
        int st = pos - 1;
        while (st > 0 && (skey[st - 1] >> 40) == lab) st--;
        int bj = -1;
        for (int q = 0; q < nbig_done; q++) if (mw.bighead[q] == st) bj = q;
        if (bj >= 0) {
            if (mw.bigcnt[bj] > m) return mw.bigvals[bj][m];
        }
        else {
            int ts[5], ks[5], a, b;
            int n = collect_block(skey, st, pos, 5, ts, ks, &a, &b);
            if (n > 0) {
                double Mx[25];
                fill_block(p.cfg, s, skey, sval, st, pos, ts, ks, a, b, Mx, 5);
                int cnt = lexicographical_values(Mx, n, J, mw.tmpvals);
                if (cnt > m) return mw.tmpvals[m];
            }
        }
        pos = st;
    }
    return 0.0;
}

// serial Murty lane for the blocks with more than five rows of one particle (thread 0 only)
__device__ double murty_lane(const KParams& p, Smem& sm, const Slab& s, MurtyWork& mw,
                             const unsigned long long* skey, const unsigned int* sval, int nll, int J, int nbig)
{
    if (nbig > kMurtyBig) { sm.ctx.status |= ST_OVER_MURTY; nbig = kMurtyBig; }
    for (int a = 1; a < nbig; a++) {   // component order = ascending head
        int v = mw.bighead[a], b = a - 1;
        while (b >= 0 && mw.bighead[b] > v) { mw.bighead[b + 1] = mw.bighead[b]; b--; }
        mw.bighead[b + 1] = v;
    }
    double contrib = 0;
    for (int bi = 0; bi < nbig; bi++) {
        const int e = mw.bighead[bi];
        const unsigned long long lab = skey[e] >> 40;
        int end = e + 1;
        while (end < nll && (skey[end] >> 40) == lab) end++;
        int ts[kMurtyN], ks[kMurtyN], a, b;
        const int n = collect_block(skey, e, end, kMurtyN, ts, ks, &a, &b);
        mw.bigcnt[bi] = 0;
        if (n < 0) { sm.ctx.status |= ST_OVER_BLOCK; continue; }
        fill_block(p.cfg, s, skey, sval, e, end, ts, ks, a, b, mw.profit, kMurtyN);
        murty_begin(mw, n);
        double* vals = mw.bigvals[bi];
        int m = 0;
        double v;
        bool have = murty_next(mw, n, &v);
        while (have) {
            // PHD:503 (the m = 0 test compares logcomp[0] with itself and never fires)
            if (m >= 200) break;
            if (m >= 1 && stale_value(p, s, mw, skey, sval, e, bi, J, m) - vals[0] < -10) break;
            vals[m++] = v;
            // the next assignment is only needed if the test for index m lets it through
            if (m >= 200) break;
            if (stale_value(p, s, mw, skey, sval, e, bi, J, m) - vals[0] < -10) break;
            have = murty_next(mw, n, &v);
        }
        if (mw.overflow) sm.ctx.status |= ST_OVER_MURTY;
        mw.bigcnt[bi] = m;
        contrib += log_sum_exp(vals, m);
    }
    return contrib;
}

// GC:280-350 on a dense n x n block (n <= 5), values pushed to vals (at most 200: PHD:469,503); perms (optional):
// the assignment of every value, five entries each
__device__ inline int lexicographical_values(const double* Mx, int n, int modelsize, double* vals,
                                             signed char* perms)
{
    int perm[5];
    for (int i = 0; i < n; i++) perm[i] = i;
    int ms = n;
    for (int i = 0; i < n; i++) if (perm[i] >= modelsize) { ms = i; break; }
    auto reverse = [&](int from, int to) {   // [from, to)
        for (int a = from, b = to - 1; a < b; a++, b--) { int t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
    };
    auto value = [&]() {
        double total = 0;
        for (int i = 0; i < n; i++) total += Mx[i * 5 + perm[i]];
        return total;
    };
    auto last = [&]() {
        for (int i = 1; i < n; i++) if (perm[i - 1] < perm[i]) return false;
        return true;
    };
    auto push = [&](int m) {
        vals[m] = value();
        if (perms) for (int i = 0; i < 5; i++) perms[5 * m + i] = (signed char)((i < n) ? perm[i] : -1);
    };
    int m = 0;
    reverse(ms, n);
    push(m++);
    while (!last() && m < 200) {
        int a, b;
        for (a = n - 2; a > 0; a--) if (perm[a] < perm[a + 1]) break;
        for (b = n - 1; b > a; b--) if (perm[a] < perm[b]) break;
        int t = perm[a]; perm[a] = perm[b]; perm[b] = t;
        reverse(a + 1, n);
        reverse(ms, n);
        push(m++);
    }
    return m;
}

// MeasurementJacobianP (PRM:185-209): d h(m) / d pose, 3 x 6 row-major: Jproj * [ -R(q*) | -R(q*) [diff]_x ]
__device__ inline void jacobian_p(const DevCfg& c, const Pose& p, const double* diff, const Quat& l, double* Jp)
{
    double mag = ((l.z > 0) ? 1 : -1) * sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
    double jp[9];
    jp[0] = c.focal / l.z; jp[1] = 0;             jp[2] = -c.focal * l.x / (l.z * l.z);
    jp[3] = 0;             jp[4] = c.focal / l.z; jp[5] = -c.focal * l.y / (l.z * l.z);
    jp[6] = l.x / mag;     jp[7] = l.y / mag;     jp[8] = l.z / mag;
    double rot[9], jloc[9], jrot[9];
    qtomatrix(qconj(p.q), rot);
    for (int i = 0; i < 9; i++) jloc[i] = -1.0 * rot[i];
    const double cross[9] = {0, -diff[2], diff[1], diff[2], 0, -diff[0], -diff[1], diff[0], 0};   // UTIL:107-112
    mat3_mul(jloc, cross, jrot);
    for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 6; cc++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += jp[r * 3 + k] * ((cc < 3) ? jloc[k * 3 + cc] : jrot[k * 3 + cc - 3]);
            Jp[r * 6 + cc] = s;
        }
}

// d ll[t, k] / d pose of every detection edge: (z_k - h(m_t))^T R^-1 Jp(m_t) (PHD:620-623), 6 doubles per edge
__device__ __noinline__ void edge_gradients(const KParams& p, Smem& sm, const Slab& s, int nll)
{
    const DevCfg& c = p.cfg;
    const int capj = p.lay.cap_j;
    const Pose pose = pose_load(sm.ctx.pose);
    for (int e = threadIdx.x; e < nll; e += kBlock) {
        const int t = (int)(s.llkey[e] >> 32), k = (int)(s.llkey[e] & 0xffffffffu);
        const double m[3] = {s.jm[t], s.jm[capj + t], s.jm[2 * capj + t]};
        double diff[3], mp[3], Jp[18], row[3];
        Quat local;
        to_local(pose, m, diff, local);
        measure_from_local(c, diff, local, mp);
        jacobian_p(c, pose, diff, local, Jp);
        const double dz[3] = {sm.zs()[3 * k] - mp[0], sm.zs()[3 * k + 1] - mp[1], sm.zs()[3 * k + 2] - mp[2]};
        for (int bb = 0; bb < 3; bb++) {
            double sum = 0;
            for (int aa = 0; aa < 3; aa++) sum += dz[aa] * c.Rinv[aa * 3 + bb];
            row[bb] = sum;
        }
        for (int l = 0; l < 6; l++) {
            double sum = 0;
            for (int bb = 0; bb < 3; bb++) sum += row[bb] * Jp[bb * 6 + l];
            s.llgrad[6 * (size_t)e + l] = sum;
        }
    }
    __syncthreads();
}

// QuasiSetLogLikelihood with the gradient (PHD:544-549, 561-713), one thread, blocks in the reference's order with
// its REAL shared 200-entry logcomp buffer: TemperedAverage (MX:400-440) overwrites logcomp[0, m) with exp(w - max)
// and normalises over the whole buffer (stale entries included), and the Murty lane's early exit reads that buffer.
// skey / sval: the detection edges sorted by (block label, landmark, measurement); grads: 6 doubles per edge.
__device__ __noinline__ double quasi_gradient_serial(const KParams& p, Smem& sm, const Slab& s, MurtyWork& mw,
                                        const unsigned long long* skey, const unsigned int* sval, int nll, int J,
                                        const int* deg, double* gradient)
{
    const DevCfg& c = p.cfg;
    const int M = p.M;
    double* logcomp = mw.tmpvals;          // 200 entries, persistent across the blocks
    double* dl = mw.lgrad;                 // 200 x 6: gradient sum of every assignment of the current block
    for (int i = 0; i < 200; i++) logcomp[i] = 0;
    for (int a = 0; a < 6; a++) gradient[a] = 0;
    double total = 0;
    const bool sumnorm = (p.ll_flags & LL_TEMPERED_SUM) != 0;
    auto edge_grad = [&](int e, int end, int t, int k) -> const double* {
        for (int q = e; q < end; q++)
            if ((int)((skey[q] >> 20) & 0xfffff) == t && (int)(skey[q] & 0xfffff) == k) return s.llgrad + 6 * (size_t)sval[q];
        return nullptr;
    };
    auto finish_block = [&](int m) {
        total += log_sum_exp(logcomp, m);
        double mx = -INFINITY;
        for (int i = 0; i < m; i++) mx = fmax(mx, logcomp[i]);
        if (isinf(mx) && mx < 0) return;
        for (int i = 0; i < m; i++) logcomp[i] = exp(logcomp[i] - mx);
        double norm = 0;
        if (sumnorm) { for (int i = 0; i < 200; i++) norm += logcomp[i]; }
        else { for (int i = 0; i < 200; i++) norm += logcomp[i] * logcomp[i]; norm = sqrt(norm); }
        double avg[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < m; i++) {
            const double wi = (norm == 0) ? logcomp[i] : logcomp[i] / norm;
            for (int a = 0; a < 6; a++) avg[a] = avg[a] + wi * dl[6 * i + a];
        }
        for (int a = 0; a < 6; a++) gradient[a] = gradient[a] + avg[a];
    };
    for (int e = 0; e < nll;) {
        const unsigned long long lab = skey[e] >> 40;
        int end = e + 1;
        while (end < nll && (skey[end] >> 40) == lab) end++;
        int ts[kMurtyN], ks[kMurtyN], a, b;
        const int n = collect_block(skey, e, end, kMurtyN, ts, ks, &a, &b);
        if (n < 0) { sm.ctx.status |= ST_OVER_BLOCK; e = end; continue; }
        int m = 0;
        if (n <= 5) {
            double Mx[25];
            signed char* perms = mw.lperm;
            fill_block(c, s, skey, sval, e, end, ts, ks, a, b, Mx, 5);
            m = lexicographical_values(Mx, n, J, logcomp, perms);
            for (int i = 0; i < m; i++) {
                for (int q = 0; q < 6; q++) dl[6 * i + q] = 0;
                for (int r = 0; r < a; r++) {
                    const int col = perms[5 * i + r];
                    if (col >= 0 && col < b) {
                        const double* g = edge_grad(e, end, ts[r], ks[col]);
                        if (g) for (int q = 0; q < 6; q++) dl[6 * i + q] = dl[6 * i + q] + g[q];
                    }
                }
            }
        }
        else {
            fill_block(c, s, skey, sval, e, end, ts, ks, a, b, mw.profit, kMurtyN);
            murty_begin(mw, n);
            double v;
            while (murty_next(mw, n, &v)) {
                if (m >= 200 || logcomp[m] - logcomp[0] < -10) break;   // PHD:503 on the live buffer
                logcomp[m] = v;
                for (int q = 0; q < 6; q++) dl[6 * m + q] = 0;
                const MurtyNode& nd = mw.nodes[mw.best];
                for (int r = 0; r < a; r++) {
                    const int col = nd.assign[r];
                    if (col >= 0 && col < b) {
                        const double* g = edge_grad(e, end, ts[r], ks[col]);
                        if (g) for (int q = 0; q < 6; q++) dl[6 * m + q] = dl[6 * m + q] + g[q];
                    }
                }
                m++;
            }
            if (mw.overflow) sm.ctx.status |= ST_OVER_MURTY;
        }
        finish_block(m);
        e = end;
    }
    // blocks without a detection (one value each, no gradient; they come last and only ever touch logcomp[0])
    for (int t = 0; t < J; t++) if (deg[t] == 0) total += log(1 - s.jpd[t]);
    for (int k = 0; k < M; k++) if (deg[J + k] == 0) total += c.logclutter;
    return total;
}

__device__ __forceinline__ bool grid_range3(const CellGrid& g, const double* q, const double* r, int* lo, int* hi)
{
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (!(q[a] + r[a] >= g.org[a]) || !(q[a] - r[a] <= g.hi[a])) return false;
        lo[a] = grid_coord(g, a, q[a] - r[a]);
        hi[a] = grid_coord(g, a, q[a] + r[a]);
    }
    return true;
}

// SetLogLikelihood (PHD:462-515) for the J landmarks in s.jm
__device__ double phase_set_loglikelihood(const KParams& p, Smem& sm, const Slab& s, int J)
{
    const DevCfg& c = p.cfg;
    const int tid = threadIdx.x, M = p.M, capj = p.lay.cap_j, capll = p.lay.cap_ll;
    const Pose pose = pose_load(sm.ctx.pose);
    const double* jx = s.jm; const double* jy = s.jm + capj; const double* jz = s.jm + 2 * capj;
    int* deg = s.bcnt;           // J + M
    int* label = s.uf;           // J + M
    __shared__ int s_nll, s_changed;
    if (tid == 0) s_nll = 0;
    for (int t = tid; t < J + M; t += kBlock) { deg[t] = 0; label[t] = t; }
    __syncthreads();

    // PHD:426-442: predicted measurement and detection probability per landmark, then the d < 5 gate
    const CellGrid& zg = p.zgrid->g;
    // QuasiSetLogLikelihood (PHD:561-713) is the same computation with full visibility (PD_i = PD) and a wider
    // association gate (d < 12 instead of d < 5)
    const bool quasi = (p.ll_flags & LL_QUASI) != 0;
    const bool want_grad = (p.ll_flags & LL_GRADIENT) != 0;
    const double gate = quasi ? 12.0 : 5.0;
    const double rad[3] = {gate * sqrt(c.R[0]) * (1 + 1e-9), gate * sqrt(c.R[4]) * (1 + 1e-9),
                           gate * sqrt(c.R[8]) * (1 + 1e-9)};
    for (int t = tid; t < J; t += kBlock) {
        double m[3] = {jx[t], jy[t], jz[t]}, diff[3], mp[3];
        Quat local;
        to_local(pose, m, diff, local);
        measure_from_local(c, diff, local, mp);
        double pdt = quasi ? c.pd : detection_probability(c, mp);
        s.jpd[t] = pdt;
        s.jmp[t] = mp[0]; s.jmp[capj + t] = mp[1]; s.jmp[2 * capj + t] = mp[2];
        int lo[3], hi[3];
        if (M == 0 || !(mp[0] == mp[0]) || !grid_range3(zg, mp, rad, lo, hi)) continue;
        const double lw = log(pdt);
        for (int cz = lo[2]; cz <= hi[2]; cz++)
            for (int cy = lo[1]; cy <= hi[1]; cy++) {
                int rowc = (cz * zg.dim[1] + cy) * zg.dim[0];
                int b = __ldg(&p.zgrid->start[rowc + lo[0]]), e = __ldg(&p.zgrid->start[rowc + hi[0] + 1]);
                for (int q = b; q < e; q++) {
                    int k = __ldg(&p.zitems[q]);
                    double d3[3] = {mp[0] - sm.zs()[3 * k], mp[1] - sm.zs()[3 * k + 1], mp[2] - sm.zs()[3 * k + 2]};
                    double d = sqrt(quadform3(c.Rinv, d3));
                    if (d < gate) {
                        int idx = atomicAdd(&s_nll, 1);
                        if (idx < capll) {
                            s.llkey[idx] = ((unsigned long long)t << 32) | (unsigned)k;
                            s.llval[idx] = lw + c.logmultR - 0.5 * d * d;
                        }
                        atomicAdd(&deg[t], 1);
                        atomicAdd(&deg[J + k], 1);
                    }
                }
            }
    }
    __syncthreads();
    if (tid == 0 && s_nll > capll) { s_nll = capll; sm.ctx.status |= ST_OVER_LL; }
    __syncthreads();
    const int nll = s_nll;
    if (tid == 0) { sm.ctx.dbg[9] += nll; sm.ctx.dbg[10] += J; }
    if (want_grad) edge_gradients(p, sm, s, nll);   // (kept out of the gate loop above: the frame path never needs it)
    if (p.ll_flags & LL_DUMP_MATRIX) {
        // SetLogLikeMatrix (PHD:415-460) as (row, column, value) triplets: detections, misses, clutter
        const int total = nll + J + M;
        if (3 * (long long)total <= (long long)kFields * p.dump_cap) {
            for (int e = tid; e < total; e += kBlock) {
                double r, cc, v;
                if (e < nll) { r = (double)(s.llkey[e] >> 32); cc = (double)(s.llkey[e] & 0xffffffffu); v = s.llval[e]; }
                else if (e < nll + J) { const int i = e - nll; r = i; cc = M + i; v = log(1 - s.jpd[i]); }
                else { const int k = e - nll - J; r = J + k; cc = k; v = c.logclutter; }
                p.dump[3 * (size_t)e] = r; p.dump[3 * (size_t)e + 1] = cc; p.dump[3 * (size_t)e + 2] = v;
            }
            if (tid == 0) *p.dump_count = total;
        }
        else if (tid == 0) { *p.dump_count = -1; sm.ctx.status |= ST_OVER_LL; }
    }

    PHASE_MARK(sm, 24);
    // GC:358-425: connected components by min-label propagation over the detection edges
    for (int it = 0; it < J + M + 1; it++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int e = tid; e < nll; e += kBlock) {
            int t = (int)(s.llkey[e] >> 32), k = (int)(s.llkey[e] & 0xffffffffu);
            int a = label[t], b = label[J + k];
            if (a < b) { atomicMin(&label[J + k], a); s_changed = 1; }
            else if (b < a) { atomicMin(&label[t], b); s_changed = 1; }
        }
        __syncthreads();
        int ch = s_changed;
        __syncthreads();
        if (!ch) break;
    }

    PHASE_MARK(sm, 25);
    // edges ordered by (component label, landmark, measurement)
    const int n2 = next_pow2(nll > 1 ? nll : 1);
    unsigned long long* skey = (n2 <= (int)p.smem_sort_cap) ? sm.skey() : s.skey;
    unsigned int* sval = (n2 <= (int)p.smem_sort_cap) ? sm.sval() : s.sval;
    for (int e = tid; e < n2; e += kBlock) {
        if (e < nll) {
            unsigned long long t = s.llkey[e] >> 32, k = s.llkey[e] & 0xffffffffu;
            skey[e] = ((unsigned long long)label[t] << 40) | (t << 20) | k;
            sval[e] = (unsigned)e;
        }
        else { skey[e] = ~0ull; sval[e] = ~0u; }
    }
    block_bitonic_sort(skey, sval, n2);

    PHASE_MARK(sm, 26);
    if (want_grad) {
        __shared__ double s_total;
        if (tid == 0) {
            double g[6];
            s_total = quasi_gradient_serial(p, sm, s, *reinterpret_cast<MurtyWork*>(s.mslots), skey, sval, nll, J, deg, g);
            for (int a = 0; a < 6; a++) sm.ctx.grad[a] = g[a];
        }
        __syncthreads();
        return s_total;
    }
    double contrib = 0;
    __shared__ int s_nbig;
    if (tid == 0) s_nbig = 0;
    __syncthreads();
    MurtyWork& mw = *reinterpret_cast<MurtyWork*>(s.mslots);
    // blocks with at least one detection edge: one thread per block (head = first edge of the label)
    for (int e = tid; e < nll; e += kBlock) {
        unsigned long long lab = skey[e] >> 40;
        if (e > 0 && (skey[e - 1] >> 40) == lab) continue;
        int end = e + 1;
        while (end < nll && (skey[end] >> 40) == lab) end++;
        int ts[5], ks[5], a, b;
        const int n = collect_block(skey, e, end, 5, ts, ks, &a, &b);
        if (n < 0) {   // more than five rows: Murty lane, handled serially below (PHD:496-499)
            int slot = atomicAdd(&s_nbig, 1);
            if (slot < kMurtyBig) mw.bighead[slot] = e;
            continue;
        }
        double Mx[25], vals[200];
        fill_block(c, s, skey, sval, e, end, ts, ks, a, b, Mx, 5);
        int m = lexicographical_values(Mx, n, J, vals);
        contrib += log_sum_exp(vals, m);
    }
    PHASE_MARK(sm, 40);
    // isolated landmarks (1x1 block: ln(1 - PD)) and isolated measurements (1x1 block: ln clutter)
    for (int t = tid; t < J; t += kBlock) if (deg[t] == 0) contrib += log(1 - s.jpd[t]);
    for (int k = tid; k < M; k += kBlock) if (deg[J + k] == 0) contrib += c.logclutter;
    __syncthreads();
    if (tid == 0) sm.ctx.dbg[11] += s_nbig;
    if (tid == 0 && s_nbig > 0) contrib += murty_lane(p, sm, s, mw, skey, sval, nll, J, s_nbig);
    double total = block_sum(sm.sh, contrib);
    __syncthreads();
    return total;
}

__device__ double phase_weight(const KParams& p, Smem& sm, const Slab& s, const double* predmap, int npriorcov,
                               const double* corr, int ncorr, double* parts, bool recs_ready)
{
    const int tid = threadIdx.x, capp = p.lay.cap_pred, capj = p.lay.cap_j;
    const int Npred = sm.ctx.Npred;

    // MAP:61-71 ExpectedSize of both maps
    double a = 0, b = 0;
    for (int i = tid; i < Npred; i += kBlock) a += s.pwt[i];
    for (int i = tid; i < ncorr; i += kBlock) b += mfield(corr, p.cap, 0)[i];
    const double pcount = block_sum(sm.sh, a);
    double ccount = block_sum(sm.sh, b);
    // size = (int)ccount truncates: when the tree sum lands within rounding distance of an integer, the
    // reference's in-order sum (MAP:61-71) decides which side it falls on -- replay it serially (rare)
    {
        const double nearest = floor(ccount + 0.5);
        if (fabs(ccount - nearest) <= 1e-9 * fmax(1.0, fabs(ccount))) {
            __shared__ double s_serial;
            if (tid == 0) {
                double acc = 0;
                for (int i = 0; i < ncorr; i++) acc += mfield(corr, p.cap, 0)[i];
                s_serial = acc;
            }
            __syncthreads();
            ccount = s_serial;
            __syncthreads();
        }
    }
    int size = (ccount > 0) ? ((ccount < 2.0e9) ? (int)ccount : 2000000000) : 0;
    PHASE_MARK(sm, 37);

    // MAP:119-142 BestMapEstimate: the `size` largest values of the multiset {w_i - j : j = 0,1,..},
    // ties ordered (generation, index) -- what the reference's append-and-stable-re-sort produces
    for (int i = tid; i < ncorr; i += kBlock) {
        double w = mfield(corr, p.cap, 0)[i];
        int g = 0;
        if (w > 0) { double cw = ceil(w); g = (cw < (double)size) ? (int)cw : size; }
        s.nflag[i] = g;
    }
    __syncthreads();
    int total = block_scan_array(sm.sh, s.nflag, ncorr);
    if (total > p.lay.cap_sort) { if (tid == 0) sm.ctx.status |= ST_OVER_JMAP; }
    // expanded multiset (straight into the shared-memory sort buffer when it fits: the sort then runs in place),
    // then only its `size` largest values are needed in order
    const int gcap = p.lay.cap_sort;
    const bool direct = total <= (int)p.smem_sort_cap && total <= kInPlaceRows * kBlock;
    {
        unsigned long long* ek = direct ? sm.skey() : s.skey;
        unsigned int* ev = direct ? sm.sval() : s.sval;
        for (int i = tid; i < ncorr; i += kBlock) {
            double w = mfield(corr, p.cap, 0)[i];
            int off = s.nflag[i];
            int g = ((i + 1 < ncorr) ? s.nflag[i + 1] : total) - off;
            for (int j = 0; j < g; j++) {
                if (off + j < gcap) {
                    ek[off + j] = weight_desc_key(w - (double)j);
                    ev[off + j] = (unsigned)j * (unsigned)p.cap + (unsigned)i;
                }
            }
        }
    }
    __syncthreads();
    PHASE_MARK(sm, 38);
    const int tot = min(total, gcap);
    const int wantj = min(size, tot);
    unsigned long long* skey = direct ? sm.skey() : s.skey;
    unsigned int* sval = direct ? sm.sval() : s.sval;
    int nsort = tot;
    if (wantj > 0) {
        // the whole multiset fits the shared-memory buffer: bucket sort (weights are spread out)
        bool sorted = false;
        if (direct) {
            sorted = block_bucket_sort(sm.sh, sm.skey(), sm.sval(), sm.skey(), sm.sval(), tot, reinterpret_cast<int*>(sm.vs()),
                                       reinterpret_cast<int*>(sm.vs()) + kSortBuckets + 1, s.skey, s.sval);
            if (!sorted) {   // degenerate keys: the general path below works from the slab copy
                for (int j = tid; j < tot; j += kBlock) { s.skey[j] = sm.skey()[j]; s.sval[j] = sm.sval()[j]; }
                __syncthreads();
                skey = s.skey; sval = s.sval;
            }
        }
        else if (tot <= (int)p.smem_sort_cap) {
            sorted = block_bucket_sort(sm.sh, s.skey, s.sval, sm.skey(), sm.sval(), tot, reinterpret_cast<int*>(sm.vs()),
                                       reinterpret_cast<int*>(sm.vs()) + kSortBuckets + 1, s.skey, s.sval);
            if (sorted) { skey = sm.skey(); sval = sm.sval(); }
        }
        if (!sorted) {   // too large, or degenerate keys: radix-select the part that is needed, bitonic sort
            int cnt = -1;
            if (wantj < tot)
                cnt = block_select_smallest(s.skey, s.sval, tot, wantj, sm.skey(), sm.sval(), (int)p.smem_sort_cap, sm.hist(),
                                            &sm.ctx.nsel, &sm.ctx.selkey);
            else if (tot <= (int)p.smem_sort_cap) {
                for (int j = tid; j < tot; j += kBlock) { sm.skey()[j] = s.skey[j]; sm.sval()[j] = s.sval[j]; }
                cnt = tot;
            }
            if (cnt >= 0) { skey = sm.skey(); sval = sm.sval(); nsort = cnt; }
            const int n2 = next_pow2(nsort > 1 ? nsort : 1);
            if (skey == sm.skey() || n2 <= gcap) {
                for (int j = nsort + tid; j < n2; j += kBlock) { skey[j] = ~0ull; sval[j] = ~0u; }
                block_bitonic_sort(skey, sval, n2);
            }
        }
    }
    __syncthreads();
    PHASE_MARK(sm, 39);
    int J = min(size, total);
    if (J > capj) { J = capj; if (tid == 0) sm.ctx.status |= ST_OVER_JMAP; }
    for (int t = tid; t < J; t += kBlock) {
        int i = (int)(sval[t] % (unsigned)p.cap);
        s.jidx[t] = i;
        s.jm[t] = mfield(corr, p.cap, 1)[i]; s.jm[capj + t] = mfield(corr, p.cap, 2)[i];
        s.jm[2 * capj + t] = mfield(corr, p.cap, 3)[i];
    }
    __syncthreads();

    // PHD:381-384: sum_j ln v_pred(m_j), sum_j ln v_corr(m_j)
    CompSrc pred{s.pwt, s.pm, s.pm + capp, s.pm + 2 * capp, mfield(predmap, p.cap, 4), (size_t)p.cap, npriorcov,
                 p.cfg.birth_cov, Npred};
    double* vs = s.vsum;   // global: double atomics are native there (shared-memory ones are CAS loops)
    PHASE_MARK(sm, 11);
    grid_build(sm.sh, sm.ctx.grid, sm.gstart(), s.gitems, s.jm, s.jm + capj, s.jm + 2 * capj, J, kQueryCell, kQueryCell,
               kQueryCell);
    if (J * 2 <= kVsCap) {
        float4* qf = reinterpret_cast<float4*>(sm.vs());
        for (int q = tid; q < J; q += kBlock) {
            const int t = s.gitems[q];
            qf[q] = make_float4((float)s.jm[t], (float)s.jm[capj + t], (float)s.jm[2 * capj + t], __int_as_float(t));
        }
        __syncthreads();
    }
    PHASE_MARK(sm, 23);
    const double plog = eval_map_at_points(p, sm, s, pred, J, vs, s.erad, recs_ready);
    PHASE_MARK(sm, 12);
    CompSrc cor{mfield(corr, p.cap, 0), mfield(corr, p.cap, 1), mfield(corr, p.cap, 2), mfield(corr, p.cap, 3),
                mfield(corr, p.cap, 4), (size_t)p.cap, ncorr, p.cfg.birth_cov, ncorr};
    const double clog = eval_map_at_points(p, sm, s, cor, J, vs, s.erad2, recs_ready);
    PHASE_MARK(sm, 13);

    const double setll = phase_set_loglikelihood(p, sm, s, J);
    PHASE_MARK(sm, 14);
    const double ratio = (plog - pcount) - (clog - ccount);
    const double alpha = exp(setll + ratio);
    parts[0] = alpha; parts[1] = setll; parts[2] = plog; parts[3] = clog; parts[4] = pcount; parts[5] = ccount;
    parts[6] = (double)J; parts[7] = 0;
    return alpha;
}

}  // namespace rbphd
