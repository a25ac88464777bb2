// rbphd_kernels.cu -- sm_100a kernels of the RB-PHD SLAM per-frame update.
//
//   k_predict_pose        particle prediction                       (PHD:295-314, TRK:89-102, POSE:314-333)
//   k_frame_prep          per-frame measurement grids (camera frame + measurement space), shared by all particles
//   k_particle_update     persistent, one CTA per particle at a time: PredictConditional + CorrectConditional
//                         + PruneModel + WeightAlpha fused; every intermediate (gated pairs, pre-prune list,
//                         merge graph) stays in a per-CTA scratch slab that is reused particle after particle
//                         (L2-resident for small maps; at config 4 the 148 slabs of 3.2 MB exceed the 126 MB L2 and
//                         measured DRAM traffic is 12x the algorithmic map bytes -- DESIGN.md sections 4 and 7.2)
//   k_normalize_resample  weight normalisation, best particle, ESS test, systematic wheel   (PHD:343-358, 724-777)
//   k_copy_particles      device-side copy of the ancestors' maps and poses                (PHD:740-742)
//
// FP64 CUDA-core arithmetic throughout (batched 3x3 algebra; tensor cores do not apply); compiled with
// -fmad=false so results match the CPU oracle bit for bit apart from libm transcendentals.
#include "rbphd_kernels.cuh"

#include <cstdlib>
#include <mutex>


namespace rbphd {

// ------------------------------------------------------------------------------------------------
// shared-memory context of one CTA of k_particle_update
// ------------------------------------------------------------------------------------------------
struct Ctx {
    int N, B, Npred, npairs, npairs_prior, L, ncand, W0, nedges, nout, nF, nU, nsel, nsel2, nact, work;
    int status;
    unsigned long long selkey;
    double pose[7];
    double grad[6];    // MODE_STAGE_SETLL with LL_GRADIENT: pose gradient of the quasi set log-likelihood
    CellGrid grid;
    CellGrid vg;       // copy of the frame's camera-frame measurement grid header
    long long tlast;
    unsigned long long tphase[64];
    unsigned int dbg[16];
};
#define DBG_ADD(sm, idx, v) atomicAdd(&(sm).ctx.dbg[idx], (unsigned int)(v))

// phase timer: thread 0 attributes the cycles since the previous mark to phase `ph`
// per-warp section timer (experiment builds, -DRBPHD_WARPTIME): lane 0 of every warp adds the cycles since its
// previous mark to slot `ph`; the slot then holds the sum over the 32 warps
#ifdef RBPHD_WARPTIME
#define WARP_T0() long long wt__ = clock64()
#define WARP_T(sm, ph)                                                                   \
    do {                                                                                 \
        long long n__ = clock64();                                                       \
        if ((threadIdx.x & 31) == 0) atomicAdd(&(sm).ctx.tphase[ph], (unsigned long long)(n__ - wt__)); \
        wt__ = n__;                                                                      \
    } while (0)
#else
#define WARP_T0() do {} while (0)
#define WARP_T(sm, ph) do {} while (0)
#endif

#define PHASE_MARK(sm, ph)                                                   \
    do {                                                                     \
        if (threadIdx.x == 0) {                                              \
            long long now__ = clock64();                                     \
            (sm).ctx.tphase[ph] += (unsigned long long)(now__ - (sm).ctx.tlast); \
            (sm).ctx.tlast = now__;                                          \
        }                                                                    \
    } while (0)


// Buffer sizes are chosen so that the CTA of a 500-measurement frame needs 166 864 B: the SM then runs with the
// 164 KB shared-memory carveout and 92 KB of L1 instead of 196 KB / 60 KB.  The L1 serves the scattered covariance
// gathers (B2, B3c, B6, C4b) and the spills: measured on c4s, 28 KB of L1 (carveout forced to 228 KB) costs +21 %,
// 92 KB instead of 60 KB gains 3 % (profiles/r02_experiments.md section 5).  Config 4 sorts ~6 000 candidates.
#ifndef RBPHD_SORT_CAP
#define RBPHD_SORT_CAP 6656
#endif
#ifndef RBPHD_VS_CAP
#define RBPHD_VS_CAP 4096
#endif
constexpr int kSortCap = RBPHD_SORT_CAP;   // (key,val) pairs the shared-memory sort buffer holds
constexpr int kVsCap = RBPHD_VS_CAP;

// Dynamic shared memory of k_particle_update: this header, then the fixed-size buffers at compile-time
// offsets, then the per-measurement arrays (sized by the frame's M).  The accessors go through the
// `extern __shared__` symbol, so the compiler knows the address space and emits LDS/STS (pointers kept in a
// struct in shared memory would turn every access into a generic load behind a pointer load).
extern __shared__ __align__(16) unsigned char g_smem[];

struct Smem {
    BlockShared sh;
    Ctx ctx;
    int Mc;          // measurements rounded up to even (set once per launch)

    static __host__ __device__ constexpr size_t off_skey() { return (sizeof(Smem) + 15) & ~size_t(15); }
    static __host__ __device__ constexpr size_t off_sval() { return off_skey() + sizeof(unsigned long long) * kSortCap; }
    static __host__ __device__ constexpr size_t off_vs() { return off_sval() + sizeof(unsigned int) * kSortCap; }
    static __host__ __device__ constexpr size_t off_hist() { return off_vs() + sizeof(double) * kVsCap; }
    static __host__ __device__ constexpr size_t off_gstart() { return off_hist() + sizeof(int) * 256; }
    static __host__ __device__ constexpr size_t off_var() { return (off_gstart() + sizeof(int) * (kGridMaxCells + 1) + 15) & ~size_t(15); }
    static __host__ __device__ size_t bytes(int M)
    {
        const size_t Mc = (size_t)(((M + 1) & ~1) > 0 ? ((M + 1) & ~1) : 2);
        return (off_var() + sizeof(double) * 7 * Mc + 2 * sizeof(int) * (Mc + 2) + 15) & ~size_t(15);
    }

    __device__ __forceinline__ unsigned long long* skey() const { return reinterpret_cast<unsigned long long*>(g_smem + off_skey()); }   // sort buffer (kSortCap)
    __device__ __forceinline__ unsigned int* sval() const { return reinterpret_cast<unsigned int*>(g_smem + off_sval()); }
    __device__ __forceinline__ double* vs() const { return reinterpret_cast<double*>(g_smem + off_vs()); }        // kVsCap: Map.Evaluate query copies / sort histograms
    __device__ __forceinline__ int* hist() const { return reinterpret_cast<int*>(g_smem + off_hist()); }          // 256
    __device__ __forceinline__ int* gstart() const { return reinterpret_cast<int*>(g_smem + off_gstart()); }      // kGridMaxCells + 1
    __device__ __forceinline__ double* zs() const { return reinterpret_cast<double*>(g_smem + off_var()); }       // 3*M measurements
    __device__ __forceinline__ double* cs() const { return zs() + 3 * Mc; }      // 3*M measurement points in map space for the current particle
    __device__ __forceinline__ double* dens() const { return zs() + 6 * Mc; }    // M   exploration density accumulators
    __device__ __forceinline__ int* kflag() const { return reinterpret_cast<int*>(zs() + 7 * Mc); }   // M   explored flag
    __device__ __forceinline__ int* kidx() const { return kflag() + (Mc + 2); }  // M+1 birth index / generic per-measurement ints
};

static_assert(sizeof(double) * kVsCap >= sizeof(int) * kWarps * 256, "the radix-sort histograms alias the vs buffer");
static_assert(sizeof(unsigned long long) * kSortCap >= sizeof(unsigned long long) * kWarps * (kBucketWarpMax + kBucketWarpMax / 2),
              "the bucket sort's per-warp staging aliases the shared-memory key buffer");
static_assert(sizeof(double) * kVsCap + sizeof(int) * 256 >= sizeof(int) * (kSortBuckets + 1 + kBigBuckets + 2),
              "the bucket-sort histogram and bucket list alias the vs buffer and the 256-int histogram behind it");

struct Slab {
    // predicted map = prior components followed by births
    double *pm, *pwt, *pwmd, *ppd;
    int *cact, *bidx;      // predicted components with at least one gated pair; pair slots in output order
    unsigned long long* pkey;
    double *pt, *pmean, *pwgt;
    unsigned long long* hits4;   // per predicted component: up to four gated measurements packed by the counting walk
    double *crec, *cpn;    // per gated component: measurement-space record (kRecFields doubles) and updated covariance (9)
    unsigned long long *skey, *skey2;
    unsigned int *sval, *sval2;
    double *tw, *tm, *rho;
    unsigned int* tloc;    // per ranked candidate: where its covariance lives (cov_at)
    int *edst, *nstate, *nowner, *nflag, *gitems;
    int *jidx;
    double *jm, *jmp, *jpd, *vsum, *erad, *erad2, *cnorm, *crad;
    unsigned long long* llkey;
    double *llval, *llgrad;
    int *uf;
    int* bcnt;
    unsigned char* mslots;
};

__device__ __forceinline__ Slab make_slab(unsigned char* base, const ScratchLayout& l)
{
    Slab s;
    s.pm = (double*)(base + l.pm); s.pwt = (double*)(base + l.pwt); s.pwmd = (double*)(base + l.pwmd);
    s.ppd = (double*)(base + l.ppd); s.cact = (int*)(base + l.cact);
    s.bidx = (int*)(base + l.bidx);
    s.pkey = (unsigned long long*)(base + l.pkey); s.pt = (double*)(base + l.pt);
    s.pmean = (double*)(base + l.pmean); s.pwgt = (double*)(base + l.pwgt);
    s.crec = (double*)(base + l.crec); s.cpn = (double*)(base + l.cpn);
    s.hits4 = (unsigned long long*)(base + l.hits4);
    s.skey = (unsigned long long*)(base + l.skey); s.sval = (unsigned int*)(base + l.sval);
    s.skey2 = (unsigned long long*)(base + l.skey2); s.sval2 = (unsigned int*)(base + l.sval2);
    s.tw = (double*)(base + l.tw); s.tm = (double*)(base + l.tm); s.tloc = (unsigned int*)(base + l.tloc);
    s.rho = (double*)(base + l.rho);
    s.edst = (int*)(base + l.edst); s.nstate = (int*)(base + l.nstate);
    s.nowner = (int*)(base + l.nowner); s.nflag = (int*)(base + l.nflag); s.gitems = (int*)(base + l.gitems);
    s.jidx = (int*)(base + l.jidx); s.jm = (double*)(base + l.jm); s.jmp = (double*)(base + l.jmp);
    s.jpd = (double*)(base + l.jpd); s.vsum = (double*)(base + l.vsum); s.erad = (double*)(base + l.erad);
    s.erad2 = (double*)(base + l.erad2);
    s.cnorm = (double*)(base + l.cnorm); s.crad = (double*)(base + l.crad);
    s.llkey = (unsigned long long*)(base + l.llkey); s.llval = (double*)(base + l.llval);
    s.llgrad = (double*)(base + l.llgrad);
    s.uf = (int*)(base + l.uf); s.bcnt = (int*)(base + l.bcnt);
    s.mslots = base + l.mslots;
    return s;
}

// map field accessors: slab of 13*cap doubles
__device__ __forceinline__ const double* mfield(const double* map, int cap, int f) { return map + (size_t)f * cap; }
__device__ __forceinline__ double* mfield(double* map, int cap, int f) { return map + (size_t)f * cap; }

// component i of the predicted map (prior component i < N, otherwise a birth)
__device__ __forceinline__ void load_pred(const KParams& p, const Slab& s, const double* in, int N, int i,
                                          double& w, double* m, double* P)
{
    const int capp = p.lay.cap_pred;
    w = s.pwt[i];
    m[0] = s.pm[i]; m[1] = s.pm[capp + i]; m[2] = s.pm[2 * capp + i];
    if (i < N) {
#pragma unroll
        for (int a = 0; a < 9; a++) P[a] = mfield(in, p.cap, 4 + a)[i];
    }
    else {
#pragma unroll
        for (int a = 0; a < 9; a++) P[a] = p.cfg.birth_cov[a];
    }
}


// ------------------------------------------------------------------------------------------------
// enumerate -> compact -> process.  Thread-per-item loops whose inner trip count depends on the data
// run at a few active lanes per warp; so the enumeration only tests and appends (a, b) pairs to a list in
// shared memory (the idle sort buffer) and the expensive per-pair arithmetic then runs densely, one
// thread per pair.  Items are taken kBlock at a time; pairs beyond the list capacity are processed
// inline by the enumerating thread.
// ------------------------------------------------------------------------------------------------
template <class Enum, class Proc>
__device__ __forceinline__ void enumerate_then_process(Smem& sm, int n_items, uint2* list, int list_cap, uint2* ovf,
                                                       int ovf_cap, Enum enumerate, Proc process,
                                                       int mark_enum = 29, int mark_proc = 30)
{
    // Items are handed out to the warps 32 at a time from a shared counter (walk lengths differ a lot between
    // items; a static assignment leaves most warps waiting at the barrier for the unlucky one).  A round of
    // enumeration ends when the items run out or the shared-memory list is full; the pairs the warps still
    // in flight add beyond that go to an overflow list in the slab (never processed inside the enumeration
    // loop: that would drag the heavy code and its registers into the innermost loop).
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) sm.ctx.work = 0;
    for (;;) {
        if (threadIdx.x == 0) sm.ctx.nsel2 = 0;
        __syncthreads();
        const int done = sm.ctx.work;
        __syncthreads();
        if (done >= n_items) break;
        for (;;) {
            int base = n_items;
            if (lane == 0 && *(volatile int*)&sm.ctx.nsel2 < list_cap) base = atomicAdd(&sm.ctx.work, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n_items) break;
            const int i = base + lane;
            if (i < n_items)
                enumerate(i, [&](int a, int b) {
                    // one atomic per group of lanes that emit together
                    const unsigned act = __activemask();
                    const int leader = __ffs(act) - 1;
                    int idx = 0;
                    if (lane == leader) idx = atomicAdd(&sm.ctx.nsel2, __popc(act));
                    idx = __shfl_sync(act, idx, leader) + __popc(act & ((1u << lane) - 1u));
                    if (idx < list_cap) list[idx] = make_uint2((unsigned)a, (unsigned)b);
                    else if (idx - list_cap < ovf_cap) ovf[idx - list_cap] = make_uint2((unsigned)a, (unsigned)b);
                    else sm.ctx.status |= ST_OVER_PAIRS;
                });
            __syncwarp();
        }
        __syncthreads();
        PHASE_MARK(sm, mark_enum);
        const int tot = sm.ctx.nsel2;
        if (threadIdx.x == 0 && mark_enum == 29) sm.ctx.dbg[6] += tot;
        const int cnt = min(tot, list_cap);
        for (int e = threadIdx.x; e < cnt; e += kBlock) process((int)list[e].x, (int)list[e].y);
        const int nov = min(max(tot - list_cap, 0), ovf_cap);
        for (int e = threadIdx.x; e < nov; e += kBlock) process((int)ovf[e].x, (int)ovf[e].y);
        PHASE_MARK(sm, mark_proc);
    }
}

// ------------------------------------------------------------------------------------------------
// gate lookup: f(k) for every measurement k with |m_i - c_k|^2 within the correct gate (PHD:882, MAP:170-184)
// ------------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void gate_walk(const KParams& p, const Smem& sm, const double* m, const Quat& local,
                                          bool grid_in_smem, F f)
{
    const DevCfg& c = p.cfg;
    const int M = p.M;
    if (c.ungated) {
        for (int k = 0; k < M; k++) f(k);
        return;
    }
    // grid_in_smem: the cell offsets / items of the camera-frame grid were copied to sm.gstart() / sm.kidx()
    const CellGrid& g = sm.ctx.vg;
    int lo[3], hi[3];
    if (!grid_range(g, local.x, local.y, local.z, c.gate_r + 1e-9, lo, hi)) return;
    for (int cz = lo[2]; cz <= hi[2]; cz++)
        for (int cy = lo[1]; cy <= hi[1]; cy++) {
            int rowc = (cz * g.dim[1] + cy) * g.dim[0];
            int b, e;
            if (grid_in_smem) { b = sm.gstart()[rowc + lo[0]]; e = sm.gstart()[rowc + hi[0] + 1]; }
            else { b = __ldg(&p.vgrid->start[rowc + lo[0]]); e = __ldg(&p.vgrid->start[rowc + hi[0] + 1]); }
            for (int t = b; t < e; t++) {
                int k = grid_in_smem ? sm.kidx()[t] : __ldg(&p.vitems[t]);
                double dx = m[0] - sm.cs()[3 * k], dy = m[1] - sm.cs()[3 * k + 1], dz = m[2] - sm.cs()[3 * k + 2];
                double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 <= c.gate_r2) f(k);
            }
        }
}

// The counting walk of a component also packs its first four gated measurements (15 bits each) and
// min(count, 15) into one word, so that the pass that writes the pairs re-walks the grid only for the few
// components with more than four pairs.
__device__ __forceinline__ void hits_add(unsigned long long& packed, int& cnt, int k)
{
    if (cnt < 4) packed |= (unsigned long long)(unsigned)k << (15 * cnt);
    cnt++;
}
__device__ __forceinline__ unsigned long long hits_close(unsigned long long packed, int cnt, int M)
{
    return packed | ((unsigned long long)((M <= 32767) ? min(cnt, 15) : 15) << 60);
}

// ------------------------------------------------------------------------------------------------
// CorrectConditional, the part that depends on the component only (PHD:858-866, 895-897): expected
// measurement, S^-1 and the Gaussian multiplier, Kalman gain K, updated covariance (I - K H) P.  The
// reference recomputes these for every (measurement, component) pair; they are the same numbers each
// time, so they are computed once per gated component and kept in a record the pairs read.
// record (kRecFields fields, struct of arrays over the gated components in index order):
//   0-2 h(m)  3 mult  4-12 S^-1  13-21 K  22 pd*w  23-25 m
// (Evaluating the pairs in this same thread while K and S^-1 are in registers was tried: with 64 registers
// the loop runs out of spilled values and costs twice the separate dense pass.)
// ------------------------------------------------------------------------------------------------
// field f of slot a of a struct-of-arrays record block (consecutive slots are consecutive in memory, so the
// dense per-slot passes read and write it coalesced)
template <class T>
struct RecRef {
    T* base;
    size_t stride;
    __device__ __forceinline__ T& operator[](int f) const { return base[(size_t)f * stride]; }
};

__device__ __noinline__ void comp_update(const KParams& p, Smem& sm, const Slab& s, const double* in, int a,
                                         int pair_offset, bool grid_in_smem)
{
    const DevCfg& c = p.cfg;
    const int N = sm.ctx.N;
    const int i = s.cact[a];
    const int pair_base = pair_offset + s.nflag[i];
    const RecRef<double> rec{s.crec + a, (size_t)p.lay.cap_pred};
    double w, m[3], P[9];
    load_pred(p, s, in, N, i, w, m, P);
    const Pose pose = pose_load(sm.ctx.pose);
    double H[9], PH[9], Sinv[9];
    {
        double diff[3], mp[3];
        Quat local;
        to_local(pose, m, diff, local);
        // the component's pairs, in one contiguous block of the pair list (slots counted by the A2 walk)
        int slot = pair_base;
        const unsigned long long packed = s.hits4[i];
        const int npacked = (int)(packed >> 60);
        if (npacked <= 4) {
            for (int t = 0; t < npacked; t++, slot++)
                if (slot < p.lay.cap_pairs)
                    s.pkey[slot] = ((unsigned long long)((packed >> (15 * t)) & 0x7fffull) << 32) | (unsigned)a;
        }
        else {
            gate_walk(p, sm, m, local, grid_in_smem, [&](int k) {
                if (slot < p.lay.cap_pairs) s.pkey[slot] = ((unsigned long long)k << 32) | (unsigned)a;
                slot++;
            });
        }
        measure_from_local(c, diff, local, mp);
        jacobian_l(c, pose, local, H);
        rec[0] = mp[0]; rec[1] = mp[1]; rec[2] = mp[2];
        rec[22] = s.ppd[i] * w;
        rec[23] = m[0]; rec[24] = m[1]; rec[25] = m[2];
    }
    mat3_mul_bt(P, H, PH);          // PH = P H^T
    {
        double S[9];
        mat3_mul(H, PH, S);         // S = H P H^T + R
#pragma unroll
        for (int f = 0; f < 9; f++) S[f] = S[f] + c.R[f];
        const double det = mat3_inv(S, Sinv);
        rec[3] = gauss_mult(det);
#pragma unroll
        for (int f = 0; f < 9; f++) rec[4 + f] = Sinv[f];
    }
    double K[9];
    mat3_mul(PH, Sinv, K);
#pragma unroll
    for (int f = 0; f < 9; f++) rec[13 + f] = K[f];
    double KH[9], IKH[9], Pn[9];
    mat3_mul(K, H, KH);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int b = 0; b < 3; b++) IKH[r * 3 + b] = ((r == b) ? 1.0 : 0.0) - KH[r * 3 + b];
    mat3_mul(IKH, P, Pn);
#pragma unroll
    for (int f = 0; f < 9; f++) s.cpn[(size_t)f * p.lay.cap_pred + a] = Pn[f];
}

// ------------------------------------------------------------------------------------------------
// one gated pair: un-normalised weight term and updated mean (PHD:886-902) from the component's record
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void eval_pair(const KParams& p, Smem& sm, const Slab& s, int j)
{
    const int capq = p.lay.cap_pairs;
    const unsigned long long key = s.pkey[j];
    const int a = (int)(key & 0xffffffffu), k = (int)(key >> 32);
    const RecRef<const double> rec{s.crec + a, (size_t)p.lay.cap_pred};
    const double* zk = &sm.zs()[3 * k];
    const double innov[3] = {zk[0] - rec[0], zk[1] - rec[1], zk[2] - rec[2]};
    // row by row (same operation order as mat3_vec / quadform3), so that only one row of S^-1 or K is live
    {
        double quad = 0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            double t = 0;
            t += rec[4 + 3 * r + 0] * innov[0];
            t += rec[4 + 3 * r + 1] * innov[1];
            t += rec[4 + 3 * r + 2] * innov[2];
            quad += innov[r] * t;
        }
        const double q = rec[3] * exp(-0.5 * quad);
        s.pt[j] = rec[22] * q;
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double kd = 0;
        kd += rec[13 + 3 * r + 0] * innov[0];
        kd += rec[13 + 3 * r + 1] * innov[1];
        kd += rec[13 + 3 * r + 2] * innov[2];
        s.pmean[(size_t)r * capq + j] = rec[23 + r] + kd;
    }
}

// Squared cull radius of one component for Map.Evaluate (MAP:192-202): kEvalD2 * bound with
// bound >= lambda_max(P): ||P^2||_F^(1/2) = (sum lambda^4)^(1/4), within 32 % of lambda_max (the trace is up
// to 3x larger).  Computed where the covariance is in registers anyway (A2, A5, B6).
constexpr double kEvalD2 = 100.0;   // terms of Map.Evaluate beyond this Mahalanobis distance^2 are < 2e-22 of the peak
__device__ __forceinline__ double eval_radius2(const double* P)
{
    double P2[9];
    mat3_mul(P, P, P2);
    double f = 0;
#pragma unroll
    for (int a = 0; a < 9; a++) f += P2[a] * P2[a];
    return kEvalD2 * sqrt(sqrt(f)) * (1.0 + 1e-6);
}

// ------------------------------------------------------------------------------------------------
// Phase A: PredictConditional + CorrectConditional
// ------------------------------------------------------------------------------------------------
__device__ void phase_predict_correct(const KParams& p, Smem& sm, const Slab& s, const double* in, int particle)
{
    const DevCfg& c = p.cfg;
    const int M = p.M, tid = threadIdx.x, capp = p.lay.cap_pred;
    const int N = sm.ctx.N;
    const Pose pose = pose_load(sm.ctx.pose);
    const bool do_births = (p.mode == MODE_FRAME || p.mode == MODE_STAGE_PREDICT);
    const bool do_correct = (p.mode == MODE_FRAME || p.mode == MODE_STAGE_CORRECT);
    const bool eval_recs = (p.mode == MODE_FRAME && !p.only_mapping);   // WeightAlpha follows: keep P^-1 records

    // the frame's camera-frame measurement grid: offsets and items into shared memory for the A2 walk
    // (sm.gstart() / sm.kidx() are reused by the per-particle grids later in the frame)
    const bool vgrid_smem = do_correct && !c.ungated && M > 0;
    if (vgrid_smem) {
        const int ncell = sm.ctx.vg.ncell;
        for (int a = tid; a <= ncell; a += kBlock) sm.gstart()[a] = __ldg(&p.vgrid->start[a]);
        for (int a = tid; a < M; a += kBlock) sm.kidx()[a] = __ldg(&p.vitems[a]);
    }
    // A1: measurements in map space (PRM:299-312)
    for (int k = tid; k < M; k += kBlock) {
        double ck[3];
        measure_to_map(c, pose, &sm.zs()[3 * k], ck);
        sm.cs()[3 * k] = ck[0]; sm.cs()[3 * k + 1] = ck[1]; sm.cs()[3 * k + 2] = ck[2];
        sm.kflag()[k] = do_births ? 0 : 1;
    }
    __syncthreads();

    PHASE_MARK(sm, 0);
    // A2: prior components: miss-detection weight (PHD:837-840), gate lookup, frustum flag
    WARP_T0();
    for (int i = tid; i < N; i += kBlock) {
        double w = mfield(in, p.cap, 0)[i];
        double m[3] = {mfield(in, p.cap, 1)[i], mfield(in, p.cap, 2)[i], mfield(in, p.cap, 3)[i]};
        s.pm[i] = m[0]; s.pm[capp + i] = m[1]; s.pm[2 * capp + i] = m[2];
        s.pwt[i] = w;
        double diff[3], mp[3];
        Quat local;
        to_local(pose, m, diff, local);
        measure_from_local(c, diff, local, mp);
        double pdi = detection_probability(c, mp);
        s.ppd[i] = pdi;
        s.pwmd[i] = (1 - pdi) * w;
        int cnt = 0;
        unsigned long long packed = 0;
        WARP_T(sm, 41);
        if (do_correct) {
            gate_walk(p, sm, m, local, vgrid_smem, [&](int k) { hits_add(packed, cnt, k); });
            s.nflag[i] = cnt;
            s.nstate[i] = (cnt > 0) ? 1 : 0;
            s.hits4[i] = hits_close(packed, cnt, M);
        }
        WARP_T(sm, 42);
        if (do_births) {
            // upper bound of ln(w N(x; m, P)) at distance d: ln(w mult) - d^2 / (2 trace P)  (lambda_max <= trace)
            double P[9];
#pragma unroll
            for (int a = 0; a < 9; a++) P[a] = mfield(in, p.cap, 4 + a)[i];
            const double det = P[0] * (P[4] * P[8] - P[5] * P[7]) - P[1] * (P[3] * P[8] - P[5] * P[6]) +
                               P[2] * (P[3] * P[7] - P[4] * P[6]);
            const double tr = P[0] + P[4] + P[8];
            const bool spd = (det > 0) && (tr > 0) && (w >= 0);
            s.cnorm[i] = spd ? log(w * gauss_mult(det)) : INFINITY;
            s.crad[i] = spd ? 1.0 / (2.0 * tr) : 0.0;
            // (frames that go on to WeightAlpha also need the component's cull radius for Map.Evaluate)
            if (eval_recs) s.erad[i] = eval_radius2(P);
            WARP_T(sm, 43);
            // Early exploration decisions (PHD:956-959, MAP:210-220): one term w_i N(c_k; m_i, P_i) >= threshold
            // decides measurement k (all terms are >= 0).  Only the component's first gated measurements are
            // tried; whatever stays undecided gets the exact sum in A4.
            if (do_correct && cnt > 0) {
                const int nt = min(cnt, 4);
                bool need = false;
                for (int t = 0; t < nt; t++) need |= !sm.kflag()[(int)((packed >> (15 * t)) & 0x7fffull)];
                if (need && M <= 32767) {
                    double Pinv[9];
                    const double gm = gauss_mult(mat3_inv(P, Pinv));
                    for (int t = 0; t < nt; t++) {
                        const int k = (int)((packed >> (15 * t)) & 0x7fffull);
                        if (sm.kflag()[k]) continue;
                        const double* ck = &sm.cs()[3 * k];
                        const double dc[3] = {ck[0] - m[0], ck[1] - m[1], ck[2] - m[2]};
                        const double e = w * (gm * exp(-0.5 * quadform3(Pinv, dc)));
                        if (e >= c.explore_thr) sm.kflag()[k] = 1;
                    }
                }
            }
            WARP_T(sm, 44);
        }
    }
    __syncthreads();
    WARP_T(sm, 45);
    // pair slots: every component's pairs take one contiguous block, blocks in component order, so that the
    // threads of a warp in the per-pair pass read the records of a handful of components (broadcast loads)
    // (and the gated components one slot each, in index order, for their records)
    int npairs_prior = 0, nact_prior = 0;
    if (do_correct) {
        block_scan_array2(sm.sh, s.nflag, s.nstate, N, &npairs_prior, &nact_prior);
        for (int i = tid; i < N; i += kBlock) {
            const int a = s.nstate[i];
            if (((i + 1 < N) ? s.nstate[i + 1] : nact_prior) != a) s.cact[a] = i;
        }
    }
    if (tid == 0) {
        if (npairs_prior > p.lay.cap_pairs) { sm.ctx.status |= ST_OVER_PAIRS; }
        sm.ctx.npairs_prior = sm.ctx.npairs = min(npairs_prior, p.lay.cap_pairs);
        sm.ctx.nact = nact_prior;
    }
    npairs_prior = min(npairs_prior, p.lay.cap_pairs);
    __syncthreads();

    PHASE_MARK(sm, 1);
    // A3: gated prior components (dense: one thread per component that has a pair), then their pairs.
    // A round of comp_update costs about the same whether 1024 or 100 threads take part (a long dependent
    // FP64 chain), so a short last round is put off and shares the round of the births (A7).  Its pairs then
    // miss the early exploration flags; those measurements are decided by the exact sum of A4 instead.
    // (The pairs themselves are all evaluated in one dense pass after the births' components, A7: the early
    // exploration decisions they used to feed are made in A2.)
    int a_split = nact_prior;
    if (do_births && nact_prior > kBlock && (nact_prior % kBlock) != 0 && (nact_prior % kBlock) <= kBlock / 2)
        a_split = (nact_prior / kBlock) * kBlock;
    {
        WARP_T0();
        for (int a = tid; a < a_split; a += kBlock) comp_update(p, sm, s, in, a, 0, vgrid_smem);
        WARP_T(sm, 46);
        __syncthreads();
        WARP_T(sm, 47);
    }
    PHASE_MARK(sm, 31);

    int B = 0;
    if (do_births) {
        // A4: measurements not yet known to be explored: gated density sum over the prior map (MAP:210-220).
        // All terms are >= 0, so the sum taken in any order differs from the reference's in-order sum by
        // at most n ulps; it decides unless it lands within 1e-9 of the threshold, in which case the
        // in-order sum is replayed (one warp per measurement).
        if (tid == 0) sm.ctx.nU = 0;
        __syncthreads();
        for (int k = tid; k < M; k += kBlock)
            if (!sm.kflag()[k]) { int u = atomicAdd(&sm.ctx.nU, 1); sm.kidx()[u] = k; }
        __syncthreads();
        const int nU = sm.ctx.nU;
        if (tid == 0) sm.ctx.dbg[0] += nU;
        PHASE_MARK(sm, 16);
        if (nU > 0) {
            // small cell grid over the undecided measurement points (cell = explore radius); every prior
            // component then visits only the points in its 3x3x3 neighbourhood.  Partial sums go to the
            // slab with native FP64 global atomics (shared-memory FP64 atomics are CAS loops).
            double* ux = reinterpret_cast<double*>(sm.skey());
            double* uy = ux + nU;
            double* uz = uy + nU;
            int* uitems = reinterpret_cast<int*>(uz + nU);
            for (int u = tid; u < nU; u += kBlock) {
                const int k = sm.kidx()[u];
                ux[u] = sm.cs()[3 * k]; uy[u] = sm.cs()[3 * k + 1]; uz[u] = sm.cs()[3 * k + 2];
                s.vsum[u] = 0.0;
            }
            __syncthreads();
            PHASE_MARK(sm, 27);
                        // cell = half the explore radius: most components reach much less than the radius (their own bound
            // above), so the finer grid halves the points a walk tests (c4: A4b 113 -> 92 k cycles per particle)
            grid_build(sm.sh, sm.ctx.grid, sm.gstart(), uitems, ux, uy, uz, nU, 0.5 * c.explore_r, 0.5 * c.explore_r,
                       0.5 * c.explore_r, 4096);
            PHASE_MARK(sm, 28);
            const CellGrid& g = sm.ctx.grid;
            const double logskip = log(c.explore_thr) - 32.3;   // ln(1e-14)
            // the point arrays live in the first 28 * nU bytes of the sort buffer; the pair list behind them
            const int list_off = (28 * nU + 15) / 16 * 2;   // in uint2 units, rounded to 16 bytes
            uint2* hits = reinterpret_cast<uint2*>(sm.skey()) + list_off;
            const int hits_cap = (int)p.smem_sort_cap - list_off;
            enumerate_then_process(
                sm, N, hits, hits_cap, reinterpret_cast<uint2*>(s.edst), p.lay.cap_edges / 2,
                [&](int i, auto emit) {
                    const double m[3] = {s.pm[i], s.pm[capp + i], s.pm[2 * capp + i]};
                    // terms whose upper bound is below 1e-14 of the threshold cannot move the sum out of the
                    // +-1e-9 band that triggers the exact replay (at most max_components of them are skipped);
                    // that bound also gives each component its own search radius inside the 1.5 m gate
                    const double lognorm = s.cnorm[i], itr = s.crad[i];
                    double reach = c.explore_r + 1e-9;
                    if (itr > 0) {
                        const double r2max = (lognorm - logskip) / itr;
                        if (!(r2max > 0)) return;
                        reach = fmin(reach, sqrt(r2max) * (1.0 + 1e-9) + 1e-12);
                    }
                    int lo[3], hi[3];
                    if (!grid_range(g, m[0], m[1], m[2], reach, lo, hi)) return;
                    for (int cz = lo[2]; cz <= hi[2]; cz++)
                        for (int cy = lo[1]; cy <= hi[1]; cy++) {
                            const int rowc = (cz * g.dim[1] + cy) * g.dim[0];
                            const int qb = sm.gstart()[rowc + lo[0]], qe = sm.gstart()[rowc + hi[0] + 1];
                            for (int q = qb; q < qe; q++) {
                                const int u = uitems[q];
                                const double dx = m[0] - ux[u], dy = m[1] - uy[u], dz = m[2] - uz[u];
                                const double d2 = dx * dx + dy * dy + dz * dz;
                                if (d2 <= c.explore_r2 && lognorm - d2 * itr >= logskip) emit(i, u);
                            }
                        }
                },
                [&](int ci, int u) {
                    const double m[3] = {s.pm[ci], s.pm[capp + ci], s.pm[2 * capp + ci]};
                    double P[9], Pinv[9];
#pragma unroll
                    for (int a = 0; a < 9; a++) P[a] = mfield(in, p.cap, 4 + a)[ci];
                    const double wm = gauss_mult(mat3_inv(P, Pinv));
                    const double dc[3] = {ux[u] - m[0], uy[u] - m[1], uz[u] - m[2]};
                    atomicAdd(&s.vsum[u], s.pwt[ci] * (wm * exp(-0.5 * quadform3(Pinv, dc))));
                    DBG_ADD(sm, 1, 1);
                },
                17, 17);
            __syncthreads();
            for (int u = tid; u < nU; u += kBlock) sm.dens()[u] = s.vsum[u];
            __syncthreads();
            PHASE_MARK(sm, 17);
            int ambiguous = 0;
            for (int u = tid; u < nU; u += kBlock) {
                const double d = sm.dens()[u];
                const int k = sm.kidx()[u];
                if (d >= c.explore_thr * (1.0 + 1e-9)) sm.kflag()[k] = 1;
                else if (d >= c.explore_thr * (1.0 - 1e-9)) { sm.kflag()[k] = 2; ambiguous = 1; }
            }
            if (__syncthreads_or(ambiguous)) {
                const int nF = N;
                const int lane = tid & 31, warp = tid >> 5;
                for (int k = warp; k < M; k += kWarps) {
                    if (sm.kflag()[k] != 2) continue;
                    const double ck[3] = {sm.cs()[3 * k], sm.cs()[3 * k + 1], sm.cs()[3 * k + 2]};
                    double sum = 0;
                    for (int base = 0; base < nF; base += 32) {
                        int f = base + lane;
                        bool hit = false;
                        double term = 0;
                        if (f < nF) {
                            int i = f;
                            double m[3] = {s.pm[i], s.pm[capp + i], s.pm[2 * capp + i]};
                            double dx = m[0] - ck[0], dy = m[1] - ck[1], dz = m[2] - ck[2];
                            double d2 = dx * dx + dy * dy + dz * dz;
                            if (d2 <= c.explore_r2) {
                                hit = true;
                                double P[9], Pinv[9];
#pragma unroll
                                for (int a = 0; a < 9; a++) P[a] = mfield(in, p.cap, 4 + a)[i];
                                double detp = mat3_inv(P, Pinv);
                                double dc[3] = {ck[0] - m[0], ck[1] - m[1], ck[2] - m[2]};
                                term = s.pwt[i] * (gauss_mult(detp) * exp(-0.5 * quadform3(Pinv, dc)));
                            }
                        }
                        unsigned mask = __ballot_sync(0xffffffffu, hit);
                        while (mask) {
                            int l = __ffs(mask) - 1;
                            sum += __shfl_sync(0xffffffffu, term, l);
                            mask &= mask - 1;
                        }
                    }
                    if (lane == 0) sm.kflag()[k] = (sum >= c.explore_thr) ? 1 : 0;
                }
                __syncthreads();
            }
        }
        __syncthreads();

        PHASE_MARK(sm, 18);
        // A5: births in measurement order (PHD:806-816)
        for (int k = tid; k < M; k += kBlock) sm.kidx()[k] = sm.kflag()[k] ? 0 : 1;
        __syncthreads();
        B = block_scan_array(sm.sh, sm.kidx(), M);
        if (N + B > capp) { B = capp - N; if (tid == 0) sm.ctx.status |= ST_OVER_COMPONENTS; }
        for (int k = tid; k < M; k += kBlock) {
            if (!sm.kflag()[k]) {
                int b = sm.kidx()[k];
                if (b < B) {
                    int i = N + b;
                    s.pm[i] = sm.cs()[3 * k]; s.pm[capp + i] = sm.cs()[3 * k + 1]; s.pm[2 * capp + i] = sm.cs()[3 * k + 2];
                    s.pwt[i] = c.birth_w;
                    if (eval_recs) s.erad[i] = eval_radius2(c.birth_cov);
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) { sm.ctx.B = B; sm.ctx.Npred = N + B; }
    __syncthreads();
    PHASE_MARK(sm, 3);
    const int Npred = N + B;

    // A6/A7: births as predicted components: miss-detection weight, gate lookup, pairs
    for (int i = N + tid; i < Npred; i += kBlock) {
        double m[3] = {s.pm[i], s.pm[capp + i], s.pm[2 * capp + i]};
        double diff[3], mp[3];
        Quat local;
        to_local(pose, m, diff, local);
        measure_from_local(c, diff, local, mp);
        double pdi = detection_probability(c, mp);
        s.ppd[i] = pdi;
        s.pwmd[i] = (1 - pdi) * c.birth_w;
        if (do_correct) {
            int cnt = 0;
            unsigned long long packed = 0;
            gate_walk(p, sm, m, local, false, [&](int k) { hits_add(packed, cnt, k); });
            s.nflag[i] = cnt;
            s.nstate[i] = (cnt > 0) ? 1 : 0;
            s.hits4[i] = hits_close(packed, cnt, M);
        }
    }
    __syncthreads();
    int nact = nact_prior;
    if (do_correct) {
        int nb = 0, na = 0;
        block_scan_array2(sm.sh, s.nflag + N, s.nstate + N, Npred - N, &nb, &na);
        for (int i = N + tid; i < Npred; i += kBlock) {
            const int a = s.nstate[i];
            if (((i + 1 < Npred) ? s.nstate[i + 1] : na) != a) s.cact[nact_prior + a] = i;
        }
        nact = nact_prior + na;
        const int np_all = npairs_prior + nb;
        if (tid == 0) {
            if (np_all > p.lay.cap_pairs) sm.ctx.status |= ST_OVER_PAIRS;
            sm.ctx.npairs = min(np_all, p.lay.cap_pairs);
            sm.ctx.nact = nact;
        }
        __syncthreads();
    }
    for (int a = a_split + tid; a < nact; a += kBlock)
        comp_update(p, sm, s, in, a, (a < nact_prior) ? 0 : npairs_prior, false);
    __syncthreads();
    PHASE_MARK(sm, 4);
    {
        WARP_T0();
        for (int j = tid; j < sm.ctx.npairs; j += kBlock) eval_pair(p, sm, s, j);
        WARP_T(sm, 48);
        __syncthreads();
        WARP_T(sm, 49);
    }
    PHASE_MARK(sm, 2);

    PHASE_MARK(sm, 4);
    // A8: order the pairs by (measurement, component) -- the reference's output order (PHD:881-903):
    // counting sort by measurement, then each measurement's short segment by component index
    const int np = sm.ctx.npairs;
    int* segstart = sm.kidx();   // M + 1 (free after the births)
    int* cursor = sm.kflag();    // M
    for (int k = tid; k <= M; k += kBlock) segstart[k] = 0;
    __syncthreads();
    for (int j = tid; j < np; j += kBlock) atomicAdd(&segstart[(int)(s.pkey[j] >> 32)], 1);
    __syncthreads();
    block_scan_array(sm.sh, segstart, M + 1);
    for (int k = tid; k < M; k += kBlock) cursor[k] = segstart[k];
    __syncthreads();
    // (measurement, slot) packed into 32 bits for the in-segment comparison
    const bool in_smem = np <= (int)p.smem_sort_cap && M <= 32767 && p.lay.cap_pred <= (1 << 17);
    if (in_smem) {
        // everything stays in shared memory: packed (measurement, component slot), pair slot and weight term
        // are scattered to the measurement's segment; A9 then ranks each pair inside its segment (one thread
        // per PAIR: balanced even when a few measurements gate dozens of components), sums the segment in
        // component order (weightsum, PHD:886-902) and writes detection weights and output order coalesced.
        unsigned int* scomp = reinterpret_cast<unsigned int*>(sm.skey());
        unsigned int* sorted = reinterpret_cast<unsigned int*>(sm.skey()) + p.smem_sort_cap;   // upper half of the key buffer
        unsigned int* sslot = sm.sval();
        double* sptmp = sm.vs();      // np doubles: runs on through the (idle) histogram and cell-offset buffers
        static_assert(sizeof(double) * kVsCap + sizeof(int) * 256 + sizeof(int) * kGridMaxCells >= sizeof(double) * kSortCap,
                      "weight terms of kSortCap pairs must fit vs + hist + gstart");
        double* wsum = sm.dens();     // M
        for (int j = tid; j < np; j += kBlock) {
            const unsigned long long key = s.pkey[j];
            const int k = (int)(key >> 32);
            const int pos = atomicAdd(&cursor[k], 1);
            scomp[pos] = ((unsigned)k << 17) | (unsigned)(key & 0x1ffffu);
            sslot[pos] = (unsigned)j;
            sptmp[pos] = s.pt[j];
        }
        __syncthreads();
        for (int t = tid; t < np; t += kBlock) {
            const unsigned myc = scomp[t];
            const int k = (int)(myc >> 17);
            const int b = segstart[k], e = segstart[k + 1];
            int rank = 0;
            for (int q = b; q < e; q++) rank += (scomp[q] < myc) ? 1 : 0;
            sorted[b + rank] = (unsigned)t;
        }
        __syncthreads();
        for (int k = tid; k < M; k += kBlock) {
            const int b = segstart[k], e = segstart[k + 1];
            double ws = 0;
            for (int q = b; q < e; q++) ws += sptmp[sorted[q]];
            wsum[k] = ws;
        }
        __syncthreads();
        for (int t = tid; t < np; t += kBlock) {
            const unsigned src = sorted[t];
            double wv = sptmp[src] / (c.clutter + wsum[scomp[src] >> 17]);
            if (wv != wv) wv = 0;   // GAUSS:154
            s.pwgt[t] = wv;
            s.bidx[t] = (int)sslot[src];
        }
    }
    else {
        for (int j = tid; j < np; j += kBlock) {
            int pos = atomicAdd(&cursor[(int)(s.pkey[j] >> 32)], 1);
            s.bidx[pos] = j;
        }
        __syncthreads();
        for (int k = tid; k < M; k += kBlock) {
            const int b = segstart[k], e = segstart[k + 1];
            for (int a = b + 1; a < e; a++) {
                const int v = s.bidx[a];
                const unsigned vi = (unsigned)(s.pkey[v] & 0xffffffffu);
                int q = a - 1;
                while (q >= b && (unsigned)(s.pkey[s.bidx[q]] & 0xffffffffu) > vi) { s.bidx[q + 1] = s.bidx[q]; q--; }
                s.bidx[q + 1] = v;
            }
            double ws = 0;
            for (int t = b; t < e; t++) ws += s.pt[s.bidx[t]];
            for (int t = b; t < e; t++) {
                double wv = s.pt[s.bidx[t]] / (c.clutter + ws);
                if (wv != wv) wv = 0;   // GAUSS:154
                s.pwgt[t] = wv;
            }
        }
    }
    __syncthreads();
    if (tid == 0) sm.ctx.L = Npred + np;
    __syncthreads();
    PHASE_MARK(sm, 5);
    (void)particle;
}

// entry e of the corrected (pre-prune) list: e < Npred the miss-detected copy, else a detection
__device__ __forceinline__ void load_corrected(const KParams& p, const Smem& sm, const Slab& s, const double* in,
                                               int e, double& w, double* m, double* P)
{
    const int Npred = sm.ctx.Npred;
    if (e < Npred) {
        double w0;
        load_pred(p, s, in, sm.ctx.N, e, w0, m, P);
        w = s.pwmd[e];
    }
    else {
        const int t = e - Npred, j = s.bidx[t], capq = p.lay.cap_pairs;
        w = s.pwgt[t];
        m[0] = s.pmean[j]; m[1] = s.pmean[capq + j]; m[2] = s.pmean[2 * capq + j];
        const size_t ca = (size_t)(s.pkey[j] & 0xffffffffu);   // the updated covariance belongs to the component
#pragma unroll
        for (int f = 0; f < 9; f++) P[f] = s.cpn[(size_t)f * p.lay.cap_pred + ca];
    }
}

// Where the covariance of entry e of the corrected list lives, as one word: a prior component (its index in the
// particle's map), a birth (the configured birth covariance) or a detection (slot of the gated component whose
// updated covariance it shares).  The ranked candidates keep this word instead of a copy of the covariance.
constexpr unsigned kLocBirth = 0x80000000u, kLocDet = 0x40000000u;
__device__ __forceinline__ void cov_at(const KParams& p, const Slab& s, const double* in, unsigned loc, double* P)
{
    if (loc & kLocBirth) {
#pragma unroll
        for (int f = 0; f < 9; f++) P[f] = p.cfg.birth_cov[f];
    }
    else if (loc & kLocDet) {
        const size_t ca = loc & ~kLocDet;
#pragma unroll
        for (int f = 0; f < 9; f++) P[f] = s.cpn[(size_t)f * p.lay.cap_pred + ca];
    }
    else {
#pragma unroll
        for (int f = 0; f < 9; f++) P[f] = mfield(in, p.cap, 4 + f)[loc];
    }
}

// ------------------------------------------------------------------------------------------------
// Phase B: PruneModel (PHD:913-948): stable weight-descending order, MinWeight / MaxQuantity cut,
// greedy Mahalanobis clustering, moment-matched merge (GAUSS:243-246, 297-347)
// ------------------------------------------------------------------------------------------------
__device__ void phase_prune(const KParams& p, Smem& sm, const Slab& s, const double* in, double* out)
{
    const DevCfg& c = p.cfg;
    const int tid = threadIdx.x;
    const int L = sm.ctx.L, Npred = sm.ctx.Npred;
    const int capw = p.lay.cap_top;

    // B1: entries that survive the MinWeight test, sorted by (weight desc, list position asc): an ordered
    // compaction (each warp owns a contiguous chunk of the list) followed by a stable radix sort on the weight
    auto entry_weight = [&](int e) { return (e < Npred) ? s.pwmd[e] : s.pwgt[e - Npred]; };
    {
        const int lane = tid & 31, warp = tid >> 5;
        const int C = (((L + kWarps - 1) / kWarps) + 31) & ~31;
        const int beg = min(L, warp * C), end = min(L, beg + C);
        constexpr int kRows = 4;   // rows of 32 entries loaded ahead of their use
        int cnt = 0;
        for (int base = beg; base < end; base += 32 * kRows) {
            double w[kRows];
#pragma unroll
            for (int r = 0; r < kRows; r++) { const int e = base + 32 * r + lane; w[r] = (e < end) ? entry_weight(e) : 0.0; }
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                const bool keep = (base + 32 * r + lane < end) && !(w[r] < c.min_w);
                cnt += __popc(__ballot_sync(0xffffffffu, keep));
            }
        }
        int total;
        int wbase = block_excl_scan(sm.sh, (lane == 0) ? cnt : 0, &total);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        // the candidates go straight into the shared-memory sort buffer when they fit (the sort then runs in place)
        const bool direct = total <= (int)p.smem_sort_cap && total <= kInPlaceRows * kBlock;
        unsigned long long* ck = direct ? sm.skey() : s.skey;
        unsigned int* cv = direct ? sm.sval() : s.sval;
        for (int base = beg; base < end; base += 32 * kRows) {
            double w[kRows];
#pragma unroll
            for (int r = 0; r < kRows; r++) { const int e = base + 32 * r + lane; w[r] = (e < end) ? entry_weight(e) : 0.0; }
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                const int e = base + 32 * r + lane;
                const bool keep = (e < end) && !(w[r] < c.min_w);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int idx = wbase + __popc(m & ((1u << lane) - 1u));
                    ck[idx] = weight_desc_key(w[r]);
                    cv[idx] = (unsigned)e;
                }
                wbase += __popc(m);
            }
        }
        if (tid == 0) sm.ctx.ncand = total;
        __syncthreads();
    }
    PHASE_MARK(sm, 36);
    const int nc = sm.ctx.ncand;
    unsigned int* sval = s.sval;   // list positions in sorted order (wherever the sort leaves them)
    {
        const bool direct = nc <= (int)p.smem_sort_cap && nc <= kInPlaceRows * kBlock;
        const bool in_smem = nc <= (int)p.smem_sort_cap;
        unsigned long long* k0 = direct ? sm.skey() : s.skey;
        unsigned int* v0 = direct ? sm.sval() : s.sval;
        unsigned long long* k1 = direct ? s.skey : (in_smem ? sm.skey() : s.skey2);
        unsigned int* v1 = direct ? s.sval : (in_smem ? sm.sval() : s.sval2);
        // weights are spread out: one bucket pass + tiny per-bucket sorts; the radix sort is the fallback for
        // degenerate key sets (thousands of equal weights)
        if (block_bucket_sort(sm.sh, k0, v0, direct ? k0 : k1, direct ? v0 : v1, nc, reinterpret_cast<int*>(sm.vs()),
                              reinterpret_cast<int*>(sm.vs()) + kSortBuckets + 1, direct ? s.skey : s.skey,
                              direct ? s.sval : s.sval, (direct || in_smem) ? nullptr : sm.skey())) {
            sval = direct ? v0 : v1;
        }
        else {
            sval = block_radix_sort(sm.sh, k0, v0, k1, v1, nc, reinterpret_cast<int*>(sm.vs()), sm.hist()) ? v1 : v0;
        }
    }
    PHASE_MARK(sm, 6);

    const int W0 = min(min(c.maxq, nc), capw);
    if (tid == 0) { sm.ctx.W0 = W0; if (min(c.maxq, nc) > capw) sm.ctx.status |= ST_OVER_COMPONENTS; }

    // B2: weight, mean and covariance locator of the W0 heaviest entries in rank order; merge radius bound per
    // candidate: d2 = D^T P^-1 D >= |D|^2 / trace(P), so d2 < t^2 needs |D|^2 < t^2 trace(P)
    double rsum = 0;
    {
        const int N = sm.ctx.N, capp = p.lay.cap_pred, capq = p.lay.cap_pairs;
        for (int r = tid; r < W0; r += kBlock) {
            const int e = (int)sval[r];
            double w, m[3], tr;
            unsigned loc;
            if (e < Npred) {
                w = s.pwmd[e];
                m[0] = s.pm[e]; m[1] = s.pm[capp + e]; m[2] = s.pm[2 * capp + e];
                if (e < N) {
                    loc = (unsigned)e;
                    tr = mfield(in, p.cap, 4)[e] + mfield(in, p.cap, 8)[e] + mfield(in, p.cap, 12)[e];
                }
                else {
                    loc = kLocBirth;
                    tr = c.birth_cov[0] + c.birth_cov[4] + c.birth_cov[8];
                }
            }
            else {
                const int t = e - Npred, j = s.bidx[t];
                w = s.pwgt[t];
                m[0] = s.pmean[j]; m[1] = s.pmean[capq + j]; m[2] = s.pmean[2 * capq + j];
                const size_t ca = (size_t)(s.pkey[j] & 0xffffffffu);   // the updated covariance belongs to the component
                loc = kLocDet | (unsigned)ca;
                tr = s.cpn[ca] + s.cpn[(size_t)4 * capp + ca] + s.cpn[(size_t)8 * capp + ca];
            }
            s.tw[r] = w;
            s.tm[r] = m[0]; s.tm[capw + r] = m[1]; s.tm[2 * capw + r] = m[2];
            s.tloc[r] = loc;
            double rho = c.merge_t * sqrt(tr) * (1.0 + 1e-9);
            if (!(tr > 0) || !(rho == rho)) rho = INFINITY;
            s.rho[r] = rho;
            if (!isinf(rho)) rsum += rho;
        }
    }
    double rmean = block_sum(sm.sh, rsum) / (W0 > 0 ? W0 : 1);
    __syncthreads();

    PHASE_MARK(sm, 7);
    // B3: cell grid over the W0 means; edges r -> r' (r' > r, close w.r.t. candidate r's covariance).
    // The neighbour walk prefilters on a cell-ordered single-precision copy of the means in shared memory
    // (conservative margin) and touches the FP64 data only for the few survivors; edges are rare (a few
    // per hundred components), so they are appended to one list and sorted by (r, r') afterwards.
    const double* tx = s.tm; const double* ty = s.tm + capw; const double* tz = s.tm + 2 * capw;
    double mincell = 2.0 * rmean;
    // (the cell offsets of this grid run from the idle vs buffer through the histogram into gstart -- the three
    // are contiguous -- so the merge grid gets kMergeCells cells instead of kGridMaxCells: the cells are far larger
    // than the merge radius at any affordable count, and every extra cell saves tests in the walk)
    constexpr int kMergeCells = (int)((sizeof(double) * kVsCap + sizeof(int) * 256 + sizeof(int) * (kGridMaxCells + 1)) / sizeof(int)) - 1;
    int* const mstart = reinterpret_cast<int*>(sm.vs());
    grid_build<kMergeCells>(sm.sh, sm.ctx.grid, mstart, s.gitems, tx, ty, tz, W0, mincell, mincell, mincell, kMergeCells);
    const double t2 = c.merge_t * c.merge_t;
    PHASE_MARK(sm, 19);
    float* fx = reinterpret_cast<float*>(sm.skey());        // cell-ordered copies (sort buffer is idle here)
    float* fy = fx + W0;
    float* fz = fy + W0;
    int* frank = reinterpret_cast<int*>(sm.sval());
    const bool fsm = (3 * (size_t)W0 * sizeof(float) <= sizeof(unsigned long long) * p.smem_sort_cap) &&
                     ((size_t)W0 <= p.smem_sort_cap);
    if (fsm) {
        for (int t = tid; t < W0; t += kBlock) {
            const int r = s.gitems[t];
            frank[t] = r;
            fx[t] = (float)tx[r]; fy[t] = (float)ty[r]; fz[t] = (float)tz[r];
        }
    }
    if (tid == 0) sm.ctx.nedges = 0;
    __syncthreads();
    PHASE_MARK(sm, 32);
    unsigned long long* elist = s.llkey;   // (r << 32 | r') edge keys; capacity cap_ll
    const int cape = p.lay.cap_ll;
    {
        const CellGrid& g = sm.ctx.grid;
        // exact test of one candidate pair (r, r2 > r): box, then Mahalanobis distance under r's covariance
        auto test = [&](int r, int r2) {
            const double rho = s.rho[r];
            const double x = tx[r], y = ty[r], z = tz[r];
            const double d[3] = {x - tx[r2], y - ty[r2], z - tz[r2]};   // a.Mean - b.Mean (GAUSS:367)
            if (rho == rho && !isinf(rho) && (fabs(d[0]) > rho || fabs(d[1]) > rho || fabs(d[2]) > rho)) return;
            double P[9], Pinv[9];
            cov_at(p, s, in, s.tloc[r], P);
            mat3_inv(P, Pinv);
            if (quadform3(Pinv, d) < t2) {
                int idx = atomicAdd(&sm.ctx.nedges, 1);
                if (idx < cape) elist[idx] = ((unsigned long long)r << 32) | (unsigned)r2;
            }
        };
        // the walk only prefilters (single precision, shared memory) and lists the surviving pairs; the exact
        // tests then run densely, one thread per pair (no global-memory latency inside the divergent walk)
        const int list_off = fsm ? (3 * W0 * (int)sizeof(float) + 15) / 16 * 2 : 0;   // uint2 units behind fx/fy/fz
        uint2* list = reinterpret_cast<uint2*>(sm.skey()) + list_off;
        const int list_cap = (int)p.smem_sort_cap - list_off;
        enumerate_then_process(
            sm, W0, list, list_cap, reinterpret_cast<uint2*>(s.edst), p.lay.cap_edges / 2,
            [&](int r, auto emit) {
                const double rho = s.rho[r];
                const double x = tx[r], y = ty[r], z = tz[r];
                int lo[3], hi[3];
                bool brute = !(rho == rho) || isinf(rho);
                if (!brute) {
                    if (!grid_range(g, x, y, z, rho, lo, hi)) return;
                    long cells = (long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
                    if (cells > 128) brute = true;
                }
                if (brute) {
                    for (int r2 = r + 1; r2 < W0; r2++) {
                        if (rho == rho && !isinf(rho) &&
                            (fabs(tx[r2] - x) > rho || fabs(ty[r2] - y) > rho || fabs(tz[r2] - z) > rho)) continue;
                        emit(r, r2);
                    }
                    return;
                }
                const float xf = (float)x, yf = (float)y, zf = (float)z;
                const float rf = (float)rho * 1.0001f + 1e-6f + 1e-6f * (fabsf(xf) + fabsf(yf) + fabsf(zf));
                for (int cz = lo[2]; cz <= hi[2]; cz++)
                    for (int cy = lo[1]; cy <= hi[1]; cy++) {
                        const int rowc = (cz * g.dim[1] + cy) * g.dim[0];
                        const int qb = mstart[rowc + lo[0]], qe = mstart[rowc + hi[0] + 1];
                        if (fsm) {
                            for (int q = qb; q < qe; q++)
                                if (fabsf(fx[q] - xf) <= rf && fabsf(fy[q] - yf) <= rf && fabsf(fz[q] - zf) <= rf) {
                                    const int r2 = frank[q];
                                    if (r2 > r) emit(r, r2);
                                }
                        }
                        else {   // more candidates than the shared-memory copy holds: walk the cell lists in the slab
                            for (int q = qb; q < qe; q++) {
                                const int r2 = s.gitems[q];
                                if (r2 > r && fabs(tx[r2] - x) <= rho && fabs(ty[r2] - y) <= rho && fabs(tz[r2] - z) <= rho)
                                    emit(r, r2);
                            }
                        }
                    }
            },
            test, 8, 33);
    }
    __syncthreads();
    PHASE_MARK(sm, 33);
    int ne = sm.ctx.nedges;
    if (tid == 0) { sm.ctx.dbg[2] += sm.ctx.nact; sm.ctx.dbg[3] += ne; sm.ctx.dbg[4] += W0; sm.ctx.dbg[5] += sm.ctx.ncand; if (ne > cape) sm.ctx.status |= ST_OVER_EDGES; }
    ne = min(ne, cape);
    // edges ordered by (r, r'): r' ascending = list order of the reference's inner loop (PHD:936-942)
    const int ne2 = next_pow2(ne > 1 ? ne : 1);
    unsigned long long* ekey = (ne2 <= (int)p.smem_sort_cap) ? sm.skey() : s.skey;
    unsigned int* eval_ = (ne2 <= (int)p.smem_sort_cap) ? sm.sval() : s.sval;
    __syncthreads();
    if (ne2 <= p.lay.cap_sort) {
        for (int e = tid; e < ne2; e += kBlock) { ekey[e] = (e < ne) ? elist[e] : ~0ull; eval_[e] = 0u; }
        block_bitonic_sort(ekey, eval_, ne2);
    }
    else if (tid == 0) sm.ctx.status |= ST_OVER_EDGES;

    PHASE_MARK(sm, 34);
    // B4: which candidates survive.  A component is absorbed iff some surviving earlier candidate is
    // close to it; resolve in rounds (the lowest undecided rank is always decidable).
    for (int r = tid; r < W0; r += kBlock) { s.nstate[r] = 0; s.nflag[r] = 0; s.nowner[r] = 0x7fffffff; }
    __syncthreads();
    for (int round = 0; round <= W0; round++) {
        for (int e = tid; e < ne; e += kBlock) {
            const int src = (int)(ekey[e] >> 32), dst = (int)(ekey[e] & 0xffffffffu);
            const int st = s.nstate[src];
            if (st != 2 && s.nstate[dst] == 0) atomicOr(&s.nflag[dst], (st == 1) ? 2 : 1);
        }
        __syncthreads();
        int und = 0;
        for (int r = tid; r < W0; r += kBlock) {
            if (s.nstate[r] == 0) {
                int f = s.nflag[r];
                if (f & 2) s.nstate[r] = 2;
                else if (!(f & 1)) s.nstate[r] = 1;
                else und++;
                s.nflag[r] = 0;
            }
        }
        int tot = block_sum_int(sm.sh, und);
        __syncthreads();
        if (tot == 0) break;
    }
    // B5: owner of each absorbed component = the first surviving candidate that is close to it
    for (int e = tid; e < ne; e += kBlock) {
        const int src = (int)(ekey[e] >> 32), dst = (int)(ekey[e] & 0xffffffffu);
        if (s.nstate[src] == 1) atomicMin(&s.nowner[dst], src);
    }
    __syncthreads();
    PHASE_MARK(sm, 9);
    // B6: output slot of each survivor, then the merge (GAUSS:329-346) in list order
    for (int r = tid; r < W0; r += kBlock) s.nflag[r] = (s.nstate[r] == 1) ? 1 : 0;
    __syncthreads();
    int nout = block_scan_array(sm.sh, s.nflag, W0);
    PHASE_MARK(sm, 35);
    if (nout > p.cap) { nout = p.cap; if (tid == 0) sm.ctx.status |= ST_OVER_COMPONENTS; }
    for (int r = tid; r < W0; r += kBlock) {
        if (s.nstate[r] != 1) continue;
        int o = s.nflag[r];
        if (o >= nout) continue;
        double weight = 0.0, mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int b = lower_bound_u64(ekey, ne, (unsigned long long)r << 32);
        int e = b;
        while (e < ne && (int)(ekey[e] >> 32) == r) e++;
        int a = b - 1;
        int member = r;
        while (true) {
            double w = s.tw[member];
            double m[3] = {tx[member], ty[member], tz[member]};
            double Pm[9];
            cov_at(p, s, in, s.tloc[member], Pm);
            weight += w;
#pragma unroll
            for (int i = 0; i < 3; i++) mean[i] = mean[i] + w * m[i];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    cov[i * 3 + k] = cov[i * 3 + k] + w * (Pm[i * 3 + k] + m[i] * m[k]);
            // next member owned by r
            a++;
            while (a < e && s.nowner[(int)(ekey[a] & 0xffffffffu)] != r) a++;
            if (a >= e) break;
            member = (int)(ekey[a] & 0xffffffffu);
        }
        double ow, om[3], oP[9];
        if (weight < 1e-15) {
            ow = 0.0;
            om[0] = tx[r]; om[1] = ty[r]; om[2] = tz[r];
#pragma unroll
            for (int i = 0; i < 9; i++) oP[i] = (i % 4 == 0) ? 1e12 : 0.0;
        }
        else {
#pragma unroll
            for (int i = 0; i < 3; i++) om[i] = mean[i] / weight;
            double rw = 1 / weight;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) oP[i * 3 + k] = rw * cov[i * 3 + k] - om[i] * om[k];
            ow = weight;
            if (ow != ow) ow = 0;
        }
        mfield(out, p.cap, 0)[o] = ow;
        mfield(out, p.cap, 1)[o] = om[0]; mfield(out, p.cap, 2)[o] = om[1]; mfield(out, p.cap, 3)[o] = om[2];
#pragma unroll
        for (int i = 0; i < 9; i++) mfield(out, p.cap, 4 + i)[o] = oP[i];
        if (p.mode == MODE_FRAME && !p.only_mapping) s.erad2[o] = eval_radius2(oP);
    }
    if (tid == 0) sm.ctx.nout = nout;
    __syncthreads();
    PHASE_MARK(sm, 10);
}

// implemented in rbphd_weight.cuh (included below): WeightAlpha (PHD:373-393)
__device__ double phase_weight(const KParams& p, Smem& sm, const Slab& s, const double* predcov, int npriorcov,
                               const double* corr, int ncorr, double* parts, bool recs_ready);
__device__ double phase_set_loglikelihood(const KParams& p, Smem& sm, const Slab& s, int J);

}  // namespace rbphd

#include "rbphd_murty.cuh"
#include "rbphd_weight.cuh"

namespace rbphd {

// dump helper for the stage entry points: write component (w,m,P) at position o of the dump buffer
__device__ __forceinline__ void dump_comp(const KParams& p, int o, double w, const double* m, const double* P)
{
    if (o >= p.dump_cap) return;
    double* d = p.dump;
    const int dc = p.dump_cap;
    d[o] = w;
    d[(size_t)1 * dc + o] = m[0]; d[(size_t)2 * dc + o] = m[1]; d[(size_t)3 * dc + o] = m[2];
    for (int a = 0; a < 9; a++) d[(size_t)(4 + a) * dc + o] = P[a];
}

// ------------------------------------------------------------------------------------------------
// the fused per-particle kernel (persistent: CTA b processes particles b, b+grid, ...)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, kCtasPerSm) k_particle_update(const __grid_constant__ KParams p)
{
    Smem& sm = *reinterpret_cast<Smem*>(g_smem);
    if (threadIdx.x == 0) sm.Mc = ((p.M + 1) & ~1) > 0 ? ((p.M + 1) & ~1) : 2;
    __syncthreads();
    const int tid = threadIdx.x;
    const Slab s = make_slab(p.scratch + (size_t)blockIdx.x * p.lay.bytes, p.lay);
    const int cur = p.st->cur;

    // measurements -> shared memory once per CTA, by the TMA engine (1-D bulk copy, mbarrier completion)
    {
        __shared__ __align__(8) unsigned long long bar;
        const unsigned bytes = (unsigned)(sizeof(double) * 3 * p.M);
        const bool bulk_ok = (bytes % 16 == 0) && bytes > 0 && ((reinterpret_cast<uintptr_t>(p.z) & 15) == 0);
        if (bulk_ok) {
            if (tid == 0) {
                unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
                unsigned dst_a = (unsigned)__cvta_generic_to_shared(sm.zs());
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                    "l"(p.z), "r"(bytes), "r"(bar_a)
                    : "memory");
            }
            __syncthreads();
            unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
            unsigned done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar_a), "r"(0u)
                    : "memory");
            }
        }
        else {
            for (int t = tid; t < 3 * p.M; t += kBlock) sm.zs()[t] = p.z[t];
        }
        __syncthreads();
    }

    if (tid == 0) sm.ctx.vg = p.vgrid->g;
    unsigned long long acc_in = 0, acc_out = 0, acc_pairs = 0, acc_pf = 0;   // thread 0 only
    if (tid == 0) { for (int a = 0; a < 64; a++) sm.ctx.tphase[a] = 0; for (int a = 0; a < 16; a++) sm.ctx.dbg[a] = 0; sm.ctx.tlast = clock64(); }
    for (int particle = p.first + blockIdx.x; particle < p.first + p.P; particle += gridDim.x) {
        const double* in = p.maps[cur] + (size_t)particle * kFields * p.cap;
        double* out = p.maps[1 - cur] + (size_t)particle * kFields * p.cap;
        if (p.mode == MODE_FRAME && particle == p.holdout) {
            // held-out filter of a leave-one-out batch (LoopyPHDNavigator.FilterMissing, LOOPY:729-763): this frame's
            // factor is skipped, the map moves to the other buffer unchanged
            const int n = min(p.counts[cur][particle], p.cap);
            for (int f = 0; f < kFields; f++)
                for (int t = tid; t < n; t += kBlock) out[(size_t)f * p.cap + t] = in[(size_t)f * p.cap + t];
            if (tid == 0) p.counts[1 - cur][particle] = n;
            continue;
        }
        if (tid == 0) {
            Ctx& c = sm.ctx;
            c.N = min(p.counts[cur][particle], p.cap);
            c.B = 0; c.Npred = c.N; c.npairs = 0; c.npairs_prior = 0; c.L = c.N; c.ncand = 0; c.W0 = 0;
            c.nedges = 0; c.nout = 0; c.nF = 0; c.status = 0; c.nact = 0;
            for (int a = 0; a < 7; a++) c.pose[a] = p.poses[(size_t)particle * 7 + a];
        }
        __syncthreads();

        if (p.mode == MODE_STAGE_PRUNE) {
            // the given map IS the list to prune
            const int N = sm.ctx.N, capp = p.lay.cap_pred;
            for (int i = tid; i < N; i += kBlock) {
                s.pwt[i] = mfield(in, p.cap, 0)[i];
                s.pwmd[i] = s.pwt[i];
                s.pm[i] = mfield(in, p.cap, 1)[i]; s.pm[capp + i] = mfield(in, p.cap, 2)[i];
                s.pm[2 * capp + i] = mfield(in, p.cap, 3)[i];
            }
            __syncthreads();
        }
        else if (p.mode == MODE_STAGE_WEIGHT || p.mode == MODE_STAGE_SETLL) {
            const int N = sm.ctx.N, capp = p.lay.cap_pred;
            for (int i = tid; i < N; i += kBlock) {
                s.pwt[i] = mfield(in, p.cap, 0)[i];
                s.pm[i] = mfield(in, p.cap, 1)[i]; s.pm[capp + i] = mfield(in, p.cap, 2)[i];
                s.pm[2 * capp + i] = mfield(in, p.cap, 3)[i];
            }
            __syncthreads();
        }
        else {
            phase_predict_correct(p, sm, s, in, particle);
        }

        if (p.mode == MODE_STAGE_PREDICT) {
            const int Npred = sm.ctx.Npred;
            for (int i = tid; i < Npred; i += kBlock) {
                double w, m[3], P[9];
                load_pred(p, s, in, sm.ctx.N, i, w, m, P);
                dump_comp(p, i, w, m, P);
            }
            if (tid == 0) *p.dump_count = Npred;
        }
        else if (p.mode == MODE_STAGE_CORRECT) {
            const int L = sm.ctx.L;
            for (int e = tid; e < L; e += kBlock) {
                double w, m[3], P[9];
                load_corrected(p, sm, s, in, e, w, m, P);
                dump_comp(p, e, w, m, P);
            }
            if (tid == 0) *p.dump_count = L;
        }
        else if (p.mode == MODE_FRAME || p.mode == MODE_STAGE_PRUNE) {
            phase_prune(p, sm, s, in, out);
            if (tid == 0) p.counts[1 - cur][particle] = sm.ctx.nout;
            if (p.mode == MODE_FRAME && !p.only_mapping) {
                double parts[8];
                double alpha = phase_weight(p, sm, s, in, sm.ctx.N, out, sm.ctx.nout, parts, true);
                if (tid == 0) {
                    p.alphas[particle] = alpha;
                    p.weights[particle] *= alpha;
                    if (p.alpha_parts)
                        for (int a = 0; a < 8; a++) p.alpha_parts[(size_t)particle * 8 + a] = parts[a];
                }
            }
        }
        else if (p.mode == MODE_STAGE_WEIGHT) {
            // buffer cur = predicted map, buffer 1-cur = corrected (pruned) map of the same particle slot
            double parts[8];
            int nc = min(p.counts[1 - cur][particle], p.cap);
            double alpha = phase_weight(p, sm, s, in, sm.ctx.N, out, nc, parts, false);
            if (tid == 0) {
                p.alphas[particle] = alpha;
                if (p.alpha_parts)
                    for (int a = 0; a < 8; a++) p.alpha_parts[(size_t)particle * 8 + a] = parts[a];
            }
        }
        else if (p.mode == MODE_STAGE_SETLL) {
            // landmark list = the means of the map in buffer cur
            const int J = sm.ctx.N, capj = p.lay.cap_j, capp = p.lay.cap_pred;
            for (int t = tid; t < J && t < capj; t += kBlock) {
                s.jm[t] = s.pm[t]; s.jm[capj + t] = s.pm[capp + t]; s.jm[2 * capj + t] = s.pm[2 * capp + t];
            }
            __syncthreads();
            double ll = phase_set_loglikelihood(p, sm, s, min(J, capj));
            if (tid == 0) {
                p.alphas[particle] = ll;
                if ((p.ll_flags & LL_GRADIENT) && p.alpha_parts)
                    for (int a = 0; a < 6; a++) p.alpha_parts[(size_t)particle * 8 + a] = sm.ctx.grad[a];
            }
        }
        __syncthreads();
        PHASE_MARK(sm, 15);
        if (tid == 0) {
            if (sm.ctx.status) atomicOr(&p.st->status, sm.ctx.status);
            acc_in += (unsigned long long)sm.ctx.N; acc_out += (unsigned long long)sm.ctx.nout;
            acc_pairs += (unsigned long long)sm.ctx.npairs; acc_pf += 1;
        }
        __syncthreads();
    }
    if (tid == 0 && p.mode == MODE_FRAME) {
        atomicAdd(&p.st->comps_in, acc_in); atomicAdd(&p.st->comps_out, acc_out);
        atomicAdd(&p.st->pairs, acc_pairs); atomicAdd(&p.st->particle_frames, acc_pf);
        for (int a = 0; a < 64; a++) atomicAdd(&p.st->phase_cycles[a], sm.ctx.tphase[a]);
        for (int a = 0; a < 16; a++) atomicAdd(&p.st->dbg[a], (unsigned long long)sm.ctx.dbg[a]);
    }
}

// ------------------------------------------------------------------------------------------------
// particle prediction (PHD:295-314): pose <- pose (+) reading, then (+) dt * C g   (TRK:89-102)
// ------------------------------------------------------------------------------------------------
__global__ void k_predict_pose(DevCfg cfg, int P, double* poses, Reading6 reading, double dt,
                               const double* gauss, int perfect_still)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double rd[6];
    bool zero = true;
    for (int a = 0; a < 6; a++) { rd[a] = reading.v[a]; if (rd[a] != 0) zero = false; }
    Pose pose = add_odometry(pose_load(poses + (size_t)i * 7), rd);
    if (!(perfect_still && zero)) {
        double noise[6];
        for (int a = 0; a < 6; a++) {
            double sum = 0;
            for (int k = 0; k < 6; k++) sum += cfg.chol[a * 6 + k] * gauss[(size_t)i * 6 + k];
            noise[a] = dt * (0.0 + sum);
        }
        pose = add_odometry(pose, noise);
    }
    pose_store(pose, poses + (size_t)i * 7);
}

// ------------------------------------------------------------------------------------------------
// per-frame grids over the measurements (particle independent):
//   vgrid: back-projected points alpha*(px,py,f) in the CAMERA frame (rigid image of MeasureToMap),
//          cell >= gate radius, used to find the measurements within 0.5 m of a component
//   zgrid: the measurements themselves in (px,py,range) space, cell >= 5 sigma per axis, used by the
//          d < 5 Mahalanobis gate of SetLogLikeMatrix (PHD:435-436)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_frame_prep(DevCfg cfg, const double* z, int M, FrameGrid* vg,
                                                       int* vitems, FrameGrid* zg, int* zitems, double* pts)
{
    __shared__ BlockShared sh;
    const int Mc = M > 0 ? M : 1;
    for (int k = threadIdx.x; k < M; k += kBlock) {
        double v[3];
        measure_to_camera(cfg, z + 3 * k, v);
        pts[k] = v[0]; pts[Mc + k] = v[1]; pts[2 * Mc + k] = v[2];
        pts[3 * Mc + k] = z[3 * k]; pts[4 * Mc + k] = z[3 * k + 1]; pts[5 * Mc + k] = z[3 * k + 2];
    }
    __syncthreads();
    double gr = cfg.gate_r;
    if (!(gr > 0) || isinf(gr)) gr = 0.5;
    grid_build(sh, vg->g, vg->start, vitems, pts, pts + Mc, pts + 2 * Mc, M, gr, gr, gr);
    grid_build(sh, zg->g, zg->start, zitems, pts + 3 * Mc, pts + 4 * Mc, pts + 5 * Mc, M,
               5.0 * sqrt(cfg.R[0]), 5.0 * sqrt(cfg.R[4]), 5.0 * sqrt(cfg.R[8]));
}

// ------------------------------------------------------------------------------------------------
// weight normalisation, best particle, ESS test, systematic wheel (PHD:343-358, 724-777).  One CTA.
// The reference's sums and the wheel are running floating-point recurrences.  The sums are taken here as
// fixed-order tree sums (deterministic; they differ from the in-order sums by a few ulps, far inside the
// 1e-9 bar on weights -- the weights already carry ~1e-15 of summation-order noise from Map.Evaluate).
// The wheel's DECISIONS must be exact: ancestor(i) = (number of prefix sums C_k = w_0 + .. + w_{k-1} below
// T_i = u/P + i/P) - 1 is found by a binary search over parallel prefix sums, and every decision is checked
// against a rigorous bound on what the reference's interleaved recurrence (random -= w[k]; random += 1/P)
// can differ from C_k and T_i by; if any T_i comes that close to a prefix sum, the whole wheel is replayed
// serially, operation by operation (thread 0, tiles staged through shared memory).
// ------------------------------------------------------------------------------------------------
constexpr int kTile = 4096;

__device__ __forceinline__ double block_sum_ordered(double* warp_part, double v)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < kWarps; w++) s += warp_part[w];
    return s;
}

// (value, index) maximum with the lowest index among equal values; returns the winner to every thread
__device__ __forceinline__ void block_argmax_first(double* warp_w, int* warp_i, double& bw, int& bi)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ow = __shfl_down_sync(0xffffffffu, bw, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ow > bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { warp_w[warp] = bw; warp_i[warp] = bi; }
    __syncthreads();
    bw = warp_w[0]; bi = warp_i[0];
    for (int w = 1; w < kWarps; w++)
        if (warp_w[w] > bw || (warp_w[w] == bw && warp_i[w] < bi)) { bw = warp_w[w]; bi = warp_i[w]; }
}

__global__ void __launch_bounds__(kBlock) k_normalize_resample(DevCfg cfg, int P, double* weights, double u,
                                                              int force, int* ancestors, DeviceState* st,
                                                              double* cum)
{
    __shared__ double tile[kTile];
    __shared__ double s_part[kWarps + 1];
    __shared__ double s_maxw[kWarps];
    __shared__ int s_maxi[kWarps];
    __shared__ int s_best, s_dep;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // a capacity error in this frame's map update (status bits set by k_particle_update): the frame's results
    // are not trustworthy -- leave the weights un-normalised, do not resample, do not publish the new maps;
    // rbphd_synchronize / rbphd_slam_update report the error
    if (force != 2 && st->status != 0) {
        for (int i = tid; i < P; i += kBlock) ancestors[i] = i;
        if (tid == 0) { st->resampled = 0; st->depleted = 0; }
        return;
    }
    // every thread owns a contiguous run of particles
    const int per = (P + kBlock - 1) / kBlock;
    const int i0 = min(P, tid * per), i1 = min(P, i0 + per);

    if (force != 2) {
        double part = 0;
        for (int i = i0; i < i1; i++) part += weights[i];
        double sum = block_sum_ordered(s_part, part);
        sum = (sum == 0) ? 1 : sum;
        // normalise, first-maximum argmax (strict >, starting from 0: PHD:347-354), ESS (PHD:768-777)
        double bw = 0, sq = 0;
        int bi = 0x7fffffff;
        for (int i = i0; i < i1; i++) {
            const double w = weights[i] / sum;
            weights[i] = w;
            sq += w * w;
            if (w > bw) { bw = w; bi = i; }
        }
        block_argmax_first(s_maxw, s_maxi, bw, bi);
        const double cumsq = block_sum_ordered(s_part, sq);
        if (tid == 0) {
            s_best = (bw > 0 && bi != 0x7fffffff) ? bi : st->best;
            s_dep = ((1.0 / cumsq < cfg.min_eff * P) || force == 1) ? 1 : 0;
        }
        __syncthreads();
    }
    else {
        if (tid == 0) { s_best = st->best; s_dep = 1; }
        __syncthreads();
    }
    const int dep = s_dep;
    if (!dep) {
        for (int i = tid; i < P; i += kBlock) ancestors[i] = i;
        if (tid == 0) { st->best = s_best; st->resampled = 0; st->depleted = 0; st->cur = 1 - st->cur; }
        return;
    }
    if (force == 3) {
        // decision only (rbphd_slam_update_begin): the particles are depleted, the host now draws the wheel's
        // uniform (PHD:727 draws it only in this case) and calls back with force = 2; the posterior maps stay
        // unpublished in buffer 1-cur, which is where the wheel's copy expects them
        for (int i = tid; i < P; i += kBlock) ancestors[i] = i;
        if (tid == 0) { st->best = s_best; st->resampled = 0; st->depleted = 1; }
        return;
    }

    // ---- systematic wheel (PHD:724-760) ----
    const double invP = 1.0 / P;
    int ambiguous = 0;
    // exclusive prefix sums C_k of the normalised weights: cum[k] = w_0 + .. + w_{k-1}, cum[P] = total
    double wmax = 0;
    {
        double part = 0;
        for (int i = i0; i < i1; i++) { const double w = weights[i]; part += w; wmax = fmax(wmax, fabs(w)); }
        double incl = part;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        if (tid == 0) {
            double run = 0;
            for (int w = 0; w < kWarps; w++) { const double t = s_part[w]; s_part[w] = run; run += t; }
            s_part[kWarps] = run;
        }
        __syncthreads();
        double c = s_part[warp] + (incl - part);
        for (int i = i0; i < i1; i++) { cum[i] = c; c += weights[i]; }
        if (i1 == P && i0 < P) cum[P] = c;
        const double total = s_part[kWarps];
        if (!(total == total) || isinf(total) || !(u == u)) ambiguous = 1;   // NaN / inf: the serial replay decides
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        __syncthreads();
        if (lane == 0) s_maxw[warp] = wmax;
        __syncthreads();
        for (int w = 0; w < kWarps; w++) wmax = fmax(wmax, s_maxw[w]);
    }
    __syncthreads();
    // The reference's `random` after any number of steps is (u/P + i/P - C_k) up to the rounding of at most 2P + 2
    // additions of values bounded by 1/P + wmax, and C_k / T_i here carry at most P roundings each: margin
    const double delta = 8.0 * 2.220446049250313e-16 * ((double)P + 64.0) * (invP + wmax + 1.0 * invP) + 1e-300;
    const double start = u / P;
    double bw = 0;
    int bi = 0x7fffffff;
    for (int i = tid; i < P; i += kBlock) {
        const double T = start + (double)i * invP;
        // k = number of weights consumed once particle i is served = smallest k in [0, P] with C_k >= T
        int lo = 0, hi = P + 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum[mid] >= T) hi = mid; else lo = mid + 1;
        }
        const int k = min(lo, P);
        if (k <= P && fabs(cum[k] - T) <= delta) ambiguous = 1;
        if (k > 0 && fabs(cum[k - 1] - T) <= delta) ambiguous = 1;
        const int a = (k == 0) ? 0 : k - 1;
        ancestors[i] = a;
        const double wa = weights[a];
        if (wa > bw) { bw = wa; bi = i; }   // PHD:745-748: first new particle whose ancestor has the largest weight
    }
    if (!__syncthreads_or(ambiguous)) {
        block_argmax_first(s_maxw, s_maxi, bw, bi);
        __syncthreads();
        for (int i = tid; i < P; i += kBlock) weights[i] = invP;
        if (tid == 0) {
            st->best = (bw > 0 && bi != 0x7fffffff) ? bi : s_best;
            st->resampled = 1; st->depleted = 1;
        }
        return;
    }

    // ---- serial replay (some T_i within rounding distance of a prefix sum, or non-finite weights) ----
    __shared__ double s_random, s_maxweight;
    __shared__ int s_k, s_i, s_newbest;
    if (tid == 0) { s_random = u / P; s_maxweight = 0; s_k = 0; s_i = 0; s_newbest = s_best; }
    __syncthreads();
    // walk the weight tiles; for each tile thread 0 advances the wheel as far as the tile allows
    for (int base = 0; base < P; base += kTile) {
        int n = min(kTile, P - base);
        for (int i = tid; i < n; i += kBlock) tile[i] = weights[base + i];
        __syncthreads();
        if (tid == 0) {
            double random = s_random, maxweight = s_maxweight;
            int k = s_k, i = s_i, nb = s_newbest;
            const bool last_tile = (base + n >= P);
            while (i < P) {
                // inner loop of PHD:736: consume weights while random > 0
                while (random > 0 && k < base + n) { random -= tile[k - base]; k++; }
                if (random > 0 && k < P && !last_tile) break;   // need the next tile
                int a = (k == 0) ? 0 : k - 1;
                ancestors[i] = a;
                double wa = (a >= base) ? tile[a - base] : weights[a];
                random += invP;
                if (wa > maxweight) { maxweight = wa; nb = i; }
                i++;
            }
            s_random = random; s_maxweight = maxweight; s_k = k; s_i = i; s_newbest = nb;
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < P; i += kBlock) weights[i] = 1.0 / P;
    if (tid == 0) { st->best = s_newbest; st->resampled = 1; st->depleted = 1; }
}

// ------------------------------------------------------------------------------------------------
// device-side particle copy after resampling (PHD:740-742): new particle i <- ancestor's map and pose.
// Runs every SLAM frame; exits at once when the frame did not resample.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_copy_particles(int P, int cap, double* map0, double* map1, int* cnt0,
                                                          int* cnt1, const double* poses, double* poses_tmp,
                                                          const int* ancestors, const DeviceState* st)
{
    if (!st->resampled) return;
    const int cur = st->cur;
    const double* src_maps = cur ? map0 : map1;   // posterior was written to buffer 1-cur
    double* dst_maps = cur ? map1 : map0;
    const int* src_cnt = cur ? cnt0 : cnt1;
    int* dst_cnt = cur ? cnt1 : cnt0;
    for (int i = blockIdx.x; i < P; i += gridDim.x) {
        const int a = ancestors[i];
        const int n = src_cnt[a];
        const double* src = src_maps + (size_t)a * kFields * cap;
        double* dst = dst_maps + (size_t)i * kFields * cap;
        for (int f = 0; f < kFields; f++) {
            const double2* s2 = reinterpret_cast<const double2*>(src + (size_t)f * cap);
            double2* d2 = reinterpret_cast<double2*>(dst + (size_t)f * cap);
            for (int t = threadIdx.x; t < (n + 1) / 2; t += kBlock) d2[t] = s2[t];
        }
        if (threadIdx.x == 0) dst_cnt[i] = n;
        if (threadIdx.x < 7) poses_tmp[(size_t)i * 7 + threadIdx.x] = poses[(size_t)a * 7 + threadIdx.x];
    }
}

__global__ void k_commit_poses(int P, double* poses, const double* poses_tmp, const DeviceState* st)
{
    if (!st->resampled) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P * 7) poses[i] = poses_tmp[i];
}

// ------------------------------------------------------------------------------------------------
// multi-GPU migration after a resampling decision (particles block-partitioned over the ranks).
// Every rank holds the identical global ancestor vector (the wheel gives NON-DECREASING ancestors, PHD:736)
// and the allgathered component counts of the posterior maps.  One record = [n][pose 7][13 fields x n],
// 8 + 13 n doubles: only the components that exist travel.  An ancestor goes once to each destination rank
// that needs it; both sides order the records of a (source, destination) pair by ancestor index, so no index
// lists are exchanged.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int part_lo(int r, int world, int total) { return (int)(((long long)total * r) / world); }
__device__ __forceinline__ int part_owner(int idx, int world, int total)
{
    int g = (int)(((long long)idx * world) / total);
    while (part_lo(g + 1, world, total) <= idx) g++;
    while (part_lo(g, world, total) > idx) g--;
    return g;
}

// exclusive scan of one long long per thread over the CTA (any kBlock); total in *tot
__device__ __forceinline__ long long block_excl_scan_ll(BlockShared& sh, long long v, long long* tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    long long* wsum = reinterpret_cast<long long*>(sh.warp_d);   // kWarps + 1 entries
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const long long w = (lane < kWarps) ? wsum[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kWarps) wsum[lane] = wi - w;
        if (lane == kWarps - 1) wsum[kWarps] = wi;
    }
    __syncthreads();
    const long long res = incl - v + wsum[warp];
    *tot = wsum[kWarps];
    return res;
}

// The exchange plan of rank `rank`, from the global ancestors and counts (one CTA; identical arithmetic on
// every rank, so senders and receivers agree without talking):
//   local_src[slot]  local index of the ancestor of new local particle `slot`, or -1 if it arrives by record
//   rec_off[slot]    offset (doubles) of that record in the receive buffer (records in slot order = source-major)
//   send_idx/off[t]  t-th record to pack: local particle and offset in the send buffer (destination-major)
//   hdr              [0..world) doubles to send to each rank, [world..2 world) doubles to receive from each rank,
//                    [2 world] records to pack, [2 world + 1] records to receive, [2 world + 2] 1 if the ancestors
//                    were non-decreasing (the plan is only valid then)
constexpr int kMaxRanks = 64;
__global__ void __launch_bounds__(kBlock) k_migration_plan(const int* ganc, const int* gcounts, int total, int world,
                                                          int rank, int* local_src, long long* rec_off,
                                                          int* send_idx, long long* send_off, long long* hdr)
{
    __shared__ BlockShared sh;
    __shared__ unsigned long long s_send[kMaxRanks], s_recv[kMaxRanks];
    __shared__ int s_mono;
    const int tid = threadIdx.x;
    const int lo = part_lo(rank, world, total), hi = part_lo(rank + 1, world, total);
    if (tid < kMaxRanks) { s_send[tid] = 0; s_recv[tid] = 0; }
    if (tid == 0) s_mono = 1;
    __syncthreads();
    // receiver side: my new particles lo..hi-1 in order
    long long carry = 0;
    int nrec = 0;
    for (int base = lo; base < hi; base += kBlock) {
        const int i = base + tid;
        long long size = 0;
        int a = 0, src = rank;
        bool first = false;
        if (i < hi) {
            a = ganc[i];
            src = part_owner(a, world, total);
            first = (src != rank) && (i == lo || ganc[i - 1] != a);
            if (first) size = 8 + 13ll * gcounts[a];
        }
        long long tot;
        const long long off = carry + block_excl_scan_ll(sh, size, &tot);
        carry += tot;
        nrec += __syncthreads_count(first);
        if (i < hi) {
            local_src[i - lo] = (src == rank) ? a - lo : -1;
            rec_off[i - lo] = first ? off : -1;
            if (first) atomicAdd(&s_recv[src], (unsigned long long)size);
        }
    }
    __syncthreads();
    // later copies of a remote ancestor share the record of its first occurrence in my range (binary search:
    // the ancestors are sorted)
    for (int i = lo + tid; i < hi; i += kBlock) {
        if (local_src[i - lo] >= 0 || rec_off[i - lo] >= 0) continue;
        const int a = ganc[i];
        int l = lo, h = i;
        while (l < h) { const int mid = (l + h) >> 1; if (ganc[mid] < a) l = mid + 1; else h = mid; }
        rec_off[i - lo] = rec_off[l - lo];
    }
    // sender side: every new particle of every OTHER rank whose ancestor I own, first occurrence per destination
    carry = 0;
    int nsend = 0;
    for (int base = 0; base < total; base += kBlock) {
        const int i = base + tid;
        long long size = 0;
        bool first = false;
        int a = 0, d = 0;
        if (i < total) {
            a = ganc[i];
            if (i > 0 && ganc[i - 1] > a) s_mono = 0;
            if (a >= lo && a < hi) {
                d = part_owner(i, world, total);
                first = (d != rank) && (i == part_lo(d, world, total) || ganc[i - 1] != a);
                if (first) size = 8 + 13ll * gcounts[a];
            }
        }
        long long tot;
        const long long off = carry + block_excl_scan_ll(sh, size, &tot);
        carry += tot;
        int cnt;
        const int idx = nsend + block_excl_scan(sh, first ? 1 : 0, &cnt);
        nsend += cnt;
        if (first) {
            send_idx[idx] = a - lo;
            send_off[idx] = off;
            atomicAdd(&s_send[d], (unsigned long long)size);
        }
    }
    __syncthreads();
    if (tid < world) { hdr[tid] = (long long)s_send[tid]; hdr[world + tid] = (long long)s_recv[tid]; }
    if (tid == 0) { hdr[2 * world] = nsend; hdr[2 * world + 1] = nrec; hdr[2 * world + 2] = s_mono; }
}

// record t of the send buffer <- local particle send_idx[t] of the posterior buffer (CTA per record)
__global__ void __launch_bounds__(256) k_pack_records(int cap, const double* maps, const int* counts,
                                                      const double* poses, const int* send_idx,
                                                      const long long* send_off, double* sendbuf)
{
    const int i = send_idx[blockIdx.x];
    double* r = sendbuf + send_off[blockIdx.x];
    const int n = counts[i];
    if (threadIdx.x == 0) r[0] = (double)n;
    if (threadIdx.x < 7) r[1 + threadIdx.x] = poses[(size_t)i * 7 + threadIdx.x];
    const double* src = maps + (size_t)i * kFields * cap;
    for (int f = 0; f < kFields; f++)
        for (int t = threadIdx.x; t < n; t += blockDim.x) r[8 + (size_t)f * n + t] = src[(size_t)f * cap + t];
}

// new local particle `slot` <- its local ancestor (posterior buffer) or its record (CTA per particle)
__global__ void __launch_bounds__(256) k_unpack_records(int cap, const double* src_maps, const int* src_counts,
                                                        double* dst_maps, int* dst_counts, const double* poses,
                                                        double* poses_tmp, const int* local_src,
                                                        const long long* rec_off, const double* recvbuf)
{
    const int i = blockIdx.x, a = local_src[i];
    double* dst = dst_maps + (size_t)i * kFields * cap;
    if (a >= 0) {
        const int n = src_counts[a];
        if (threadIdx.x == 0) dst_counts[i] = n;
        if (threadIdx.x < 7) poses_tmp[(size_t)i * 7 + threadIdx.x] = poses[(size_t)a * 7 + threadIdx.x];
        const double* src = src_maps + (size_t)a * kFields * cap;
        for (int f = 0; f < kFields; f++)
            for (int t = threadIdx.x; t < n; t += blockDim.x) dst[(size_t)f * cap + t] = src[(size_t)f * cap + t];
    }
    else {
        const double* r = recvbuf + rec_off[i];
        const int n = min((int)r[0], cap);
        if (threadIdx.x == 0) dst_counts[i] = n;
        if (threadIdx.x < 7) poses_tmp[(size_t)i * 7 + threadIdx.x] = r[1 + threadIdx.x];
        for (int f = 0; f < kFields; f++)
            for (int t = threadIdx.x; t < n; t += blockDim.x) dst[(size_t)f * cap + t] = r[8 + (size_t)f * n + t];
    }
}

__global__ void k_copy_doubles(size_t n, double* dst, const double* src)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

__global__ void k_fill_doubles(size_t n, double* dst, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

void launch_migration_plan(cudaStream_t s, const int* ganc, const int* gcounts, int total, int world, int rank,
                           int* local_src, long long* rec_off, int* send_idx, long long* send_off, long long* hdr)
{
    k_migration_plan<<<1, kBlock, 0, s>>>(ganc, gcounts, total, world, rank, local_src, rec_off, send_idx, send_off, hdr);
}
void launch_pack_records(cudaStream_t s, int cap, const double* maps, const int* counts, const double* poses,
                         const int* send_idx, const long long* send_off, int count, double* sendbuf)
{
    if (count > 0) k_pack_records<<<count, 256, 0, s>>>(cap, maps, counts, poses, send_idx, send_off, sendbuf);
}
void launch_unpack_records(cudaStream_t s, int P, int cap, const double* src_maps, const int* src_counts,
                           double* dst_maps, int* dst_counts, const double* poses, double* poses_tmp,
                           const int* local_src, const long long* rec_off, const double* recvbuf)
{
    if (P > 0)
        k_unpack_records<<<P, 256, 0, s>>>(cap, src_maps, src_counts, dst_maps, dst_counts, poses, poses_tmp, local_src,
                                           rec_off, recvbuf);
}
void launch_copy_doubles(cudaStream_t s, size_t n, double* dst, const double* src)
{
    if (n > 0) k_copy_doubles<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, dst, src);
}
void launch_fill_doubles(cudaStream_t s, size_t n, double* dst, double v)
{
    if (n > 0) k_fill_doubles<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, dst, v);
}

// publish the maps written to buffer 1-cur (mapping-only frames, stage calls); not after a capacity error
__global__ void k_flip(DeviceState* st)
{
    st->resampled = 0;
    if (st->status == 0) st->cur = 1 - st->cur;
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
void launch_predict_pose(cudaStream_t s, const DevCfg& cfg, int P, double* poses, Reading6 reading, double dt,
                         const double* gauss, int perfect_still)
{
    k_predict_pose<<<(P + 127) / 128, 128, 0, s>>>(cfg, P, poses, reading, dt, gauss, perfect_still);
}

void launch_frame_prep(cudaStream_t s, const DevCfg& cfg, const double* z, int M, FrameGrid* vg, int* vitems,
                       FrameGrid* zg, int* zitems, double* pts)
{
    k_frame_prep<<<1, kBlock, 0, s>>>(cfg, z, M, vg, vitems, zg, zitems, pts);
}

size_t murty_workspace_bytes() { return sizeof(MurtyWork); }

size_t particle_update_smem(int max_measurements, size_t* sort_cap)
{
    if (sort_cap) *sort_cap = kSortCap;
    return Smem::bytes(max_measurements);
}

int particle_update_max_ctas_per_sm(size_t smem)
{
    // The attribute belongs to the function on the current device, not to a handle: navigators with different
    // max_measurements coexist (and a smaller one created later must not lower the limit under an older, larger one)
    static std::mutex mu;
    static size_t granted[64] = {0};
    {
        std::lock_guard<std::mutex> lock(mu);
        int dev = 0;
        cudaGetDevice(&dev);
        size_t& g = granted[(dev >= 0 && dev < 64) ? dev : 0];
        if (smem > g) {
            if (cudaFuncSetAttribute(k_particle_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess)
                g = smem;
        }
        if (const char* co = std::getenv("RBPHD_CARVEOUT"))   // experiments: shared-memory carveout in percent
            cudaFuncSetAttribute(k_particle_update, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(co));
    }
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_particle_update, kBlock, smem);
    return n;
}

void launch_particle_update(cudaStream_t s, const KParams& prm, int grid, size_t smem)
{
    k_particle_update<<<grid, kBlock, smem, s>>>(prm);
}

void launch_normalize_resample(cudaStream_t s, const DevCfg& cfg, int P, double* weights, double u, int force,
                               int* ancestors, DeviceState* st, double* cum)
{
    k_normalize_resample<<<1, kBlock, 0, s>>>(cfg, P, weights, u, force, ancestors, st, cum);
}

void launch_copy_particles(cudaStream_t s, int P, int cap, double* const maps[2], int* const counts[2],
                           double* poses, double* poses_tmp, const int* ancestors, DeviceState* st)
{
    int grid = P < 148 * 8 ? P : 148 * 8;
    if (grid < 1) grid = 1;
    k_copy_particles<<<grid, kBlock, 0, s>>>(P, cap, maps[0], maps[1], counts[0], counts[1], poses, poses_tmp,
                                             ancestors, st);
    k_commit_poses<<<(P * 7 + 255) / 256, 256, 0, s>>>(P, poses, poses_tmp, st);
}

void launch_flip(cudaStream_t s, DeviceState* st) { k_flip<<<1, 1, 0, s>>>(st); }

}  // namespace rbphd
