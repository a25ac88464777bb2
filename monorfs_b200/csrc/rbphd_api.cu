// rbphd_api.cu -- the C ABI of librbphd.so (include/rbphd.h): handle lifecycle, host<->device
// staging, kernel sequencing on the handle's stream.  No CPU fallback exists: every entry point
// that computes fails with RBPHD_ERR_NO_DEVICE / RBPHD_ERR_CUDA when the device path is unavailable.
#include "../../include/rbphd.h"
#include "rbphd_kernels.cuh"
#include "rbphd_analysis.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace rbphd;

namespace {

thread_local std::string g_last_error;

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    PinnedBuf(PinnedBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    void* get(size_t n)
    {
        if (n > bytes) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            size_t want = std::max(n, bytes * 2);
            if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; bytes = 0; return nullptr; }
            bytes = want;
        }
        return p;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

}  // namespace

struct rbphd_navigator {
    rbphd_config cfg{};
    rbphd_limits lim{};
    DevCfg dcfg{};
    ScratchLayout lay{};
    int device = 0;
    int P = 0;              // current particle count
    int maxP = 0, cap = 0, Mcap = 0;
    cudaStream_t stream = nullptr;
    // device state
    double* maps[2] = {nullptr, nullptr};
    int* counts[2] = {nullptr, nullptr};
    double *poses = nullptr, *poses_tmp = nullptr, *weights = nullptr, *alphas = nullptr, *alpha_parts = nullptr;
    double *z = nullptr, *gauss = nullptr, *pts = nullptr;
    double* cumw = nullptr;       // maxP + 1: prefix sums of the wheel
    int* ancestors = nullptr;
    DeviceState* st = nullptr;
    FrameGrid *vgrid = nullptr, *zgrid = nullptr;
    int *vitems = nullptr, *zitems = nullptr;
    unsigned char* scratch = nullptr;
    double* dump = nullptr;
    int* dump_count = nullptr;
    int dump_cap = 0;
    // multi-GPU (rbphd_comm_init_rank): communicator, this rank's block, global work vectors, exchange buffers
    void* comm = nullptr;         // ncclComm_t
    int rank = 0, world = 1, total = 0;
    double* gweights = nullptr;   // all ranks' weights in rank order
    double* gcum = nullptr;       // total + 1: prefix sums of the wheel over the global vector
    int *ganc = nullptr, *gcounts = nullptr;
    int *local_src = nullptr, *send_idx = nullptr;
    long long *rec_off = nullptr, *send_off = nullptr, *plan_hdr = nullptr;
    double *sendbuf = nullptr, *recvbuf = nullptr;
    size_t sendcap = 0, recvcap = 0;   // doubles
    int last_best = 0, last_resampled = 0;
    int pending_wheel = 0;        // rbphd_slam_update_begin found the particles depleted; _finish runs the wheel
    int64_t comm_resamples = 0, comm_sent_bytes = 0, comm_recv_bytes = 0, comm_records = 0;
    // launch geometry
    int nslab = 0, ctas_per_sm = 0;
    size_t smem = 0, sort_cap = 0;
    int M_last = 0;
    int slots = 1;
    size_t zstride = 0, gstride = 0;   // doubles per input slot
    // profiling: 6 events per frame
    std::vector<cudaEvent_t> pev;
    int prof_frames = 0, prof_max = 0;
    // host mirrors (library-owned outputs)
    PinnedBuf h_in, h_out, h_state, h_map, h_plan;
    std::vector<PinnedBuf> slot_stage;     // per input slot: pinned staging of (gauss, z)
    std::vector<cudaEvent_t> slot_ev;      // ... and the event after its last host-to-device copy
    std::vector<double> o_w, o_m, o_P;
    std::vector<int> o_anc, o_rows, o_cols;
    int ll_flags = 0;
    int holdout = -1;             // rbphd_set_holdout
    float* depth = nullptr;       // rbphd_set_depth_frame: device copy of the Kinect depth frame
    size_t depth_cap = 0;
    double* ana = nullptr;        // scratch of rbphd_generate_measurements / rbphd_ospa
    size_t ana_cap = 0;
    int64_t launches = 0;
    std::string error;
    rbphd_navigator* stage = nullptr;   // lazily created 2-slot navigator for the stage entry points
};

namespace {

int fail(rbphd_navigator* nav, int code, const std::string& msg)
{
    if (nav) nav->error = msg;
    g_last_error = msg;
    return code;
}

void comm_free(rbphd_navigator* nav);
int comm_slam_tail(rbphd_navigator* nav, double u, cudaEvent_t* ev);

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(nav, RBPHD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

double inv3(const double* a, double* inv)
{
    double c00 = a[4] * a[8] - a[5] * a[7];
    double c01 = a[3] * a[8] - a[5] * a[6];
    double c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 - a[1] * c01 + a[2] * c02;
    double id = 1.0 / det;
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

// UTIL:173-202: lower Cholesky root, diagonal floored at 1e-40
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

void make_devcfg(const rbphd_config& c, DevCfg& d, double gate_radius_override, bool use_override)
{
    std::memset(&d, 0, sizeof d);
    std::memcpy(d.R, c.R, sizeof d.R);
    double det = inv3(c.R, d.Rinv);
    d.multR = (1.0 / (2 * 3.14159265358979323846)) / std::sqrt(det);
    d.logmultR = std::log(d.multR);
    cholesky6(c.Q, d.chol);
    d.pd = c.pd;
    d.clutter = c.clutter;
    d.logclutter = std::log(c.clutter);
    std::memcpy(d.birth_cov, c.birth_cov, sizeof d.birth_cov);
    d.birth_w = c.birth_weight;
    d.min_w = c.min_weight;
    d.merge_t = c.merge_threshold;
    d.explore_thr = c.exploration_threshold;
    double gr = use_override ? gate_radius_override : c.density_distance_threshold;
    d.ungated = (use_override && gate_radius_override < 0) ? 1 : 0;
    if (d.ungated) gr = c.density_distance_threshold;
    double er = 3 * c.density_distance_threshold;
    if (c.gate_metric == 0) {
        d.gate_r2 = gr * gr;  d.gate_r = gr;
        d.explore_r2 = er * er;  d.explore_r = er;
    }
    else {
        d.gate_r2 = gr;  d.gate_r = std::sqrt(gr > 0 ? gr : 0);
        d.explore_r2 = er;  d.explore_r = std::sqrt(er > 0 ? er : 0);
    }
    d.min_eff = c.min_effective_particle;
    for (int i = 0; i < 3; i++) d.ramp[i] = c.visibility_ramp[i];
    d.focal = c.measurer[0];
    int left = (int)c.measurer[3], top = (int)c.measurer[4];
    d.left = left;  d.top = top;
    d.right = left + (int)c.measurer[5];
    d.bottom = top + (int)c.measurer[6];
    d.rmin = (double)(float)c.measurer[1];   // AForge.Range holds floats (PRM:65,110)
    d.rmin_f = (float)c.measurer[1];
    d.rmax = (double)(float)c.measurer[2];
    d.maxq = c.max_quantity;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void make_layout(ScratchLayout& l, int cap, int Mcap, int cap_pairs, int maxq)
{
    std::memset(&l, 0, sizeof l);
    l.cap_pred = cap + Mcap;
    l.cap_pairs = cap_pairs;
    l.cap_list = l.cap_pred + cap_pairs;
    l.cap_j = 2 * cap;
    int need = std::max(l.cap_list, cap + l.cap_j);
    int ps = 1;
    while (ps < need) ps <<= 1;
    l.cap_sort = ps;
    l.cap_top = std::max(1, std::min(maxq, cap));
    l.cap_edges = 8 * l.cap_top + 64;
    l.cap_ll = std::max(8 * Mcap, 1024);
    l.cap_nodes = std::max(std::max(l.cap_pred, l.cap_top), l.cap_j) + 2;
    size_t off = 0;
    auto take = [&](size_t& field, size_t bytes) { field = off; off = align_up(off + bytes, 16); };
    const size_t D = sizeof(double), I = sizeof(int), U = sizeof(unsigned long long);
    take(l.pm, 3 * D * l.cap_pred);  take(l.pwt, D * l.cap_pred);  take(l.pwmd, D * l.cap_pred);
    take(l.ppd, D * l.cap_pred);     take(l.cact, I * (l.cap_pred + 2));
    take(l.bidx, I * (cap_pairs + 2));
    take(l.pkey, U * cap_pairs);     take(l.pt, D * cap_pairs);    take(l.pmean, 3 * D * cap_pairs);
    take(l.pwgt, D * cap_pairs);
    // Three groups of arrays with disjoint lifetimes share one region, so that the slab touches fewer distinct
    // cache lines per particle (the L2 holds ~0.85 MB per resident CTA):
    //   the Kalman records (written by comp_update, dead once the pairs are evaluated: A3 - A7),
    //   the ranked candidates of PruneModel (B2 - B6),
    //   the map-estimate points of WeightAlpha (C3 - C5).
    {
        const size_t base = off;
        take(l.crec, kRecFields * D * l.cap_pred);
        const size_t end_a = off;
        off = base;
        take(l.tw, D * l.cap_top);   take(l.tm, 3 * D * l.cap_top);
        take(l.tloc, sizeof(unsigned) * l.cap_top);
        take(l.rho, D * l.cap_top);
        const size_t end_b = off;
        off = base;
        take(l.jidx, I * l.cap_j);   take(l.jm, 3 * D * l.cap_j);  take(l.jmp, 3 * D * l.cap_j);
        take(l.jpd, D * l.cap_j);
        const size_t end_c = off;
        off = std::max(end_a, std::max(end_b, end_c));
    }
    take(l.cpn, 9 * D * l.cap_pred);
    take(l.hits4, U * l.cap_pred);
    take(l.skey, U * l.cap_sort);    take(l.sval, sizeof(unsigned) * l.cap_sort);
    take(l.skey2, U * l.cap_sort);   take(l.sval2, sizeof(unsigned) * l.cap_sort);
    take(l.edst, I * l.cap_edges);
    take(l.nstate, I * l.cap_nodes); take(l.nowner, I * l.cap_nodes); take(l.nflag, I * l.cap_nodes);
    take(l.gitems, I * l.cap_nodes);
    take(l.vsum, D * l.cap_j);
    take(l.erad, D * l.cap_pred); take(l.erad2, D * l.cap_pred); take(l.cnorm, D * l.cap_pred); take(l.crad, D * l.cap_pred);   // exploration bound per component
    take(l.llkey, U * l.cap_ll);     take(l.llval, D * l.cap_ll);   take(l.llgrad, 6 * D * l.cap_ll);
    take(l.uf, I * (l.cap_j + Mcap + 2)); take(l.bcnt, I * (l.cap_j + Mcap + 2));
    take(l.mslots, murty_workspace_bytes());
    l.bytes = align_up(off, 256);
}

void free_device(rbphd_navigator* nav)
{
    cudaSetDevice(nav->device);
    for (int b = 0; b < 2; b++) { cudaFree(nav->maps[b]); cudaFree(nav->counts[b]); }
    cudaFree(nav->poses); cudaFree(nav->poses_tmp); cudaFree(nav->weights); cudaFree(nav->alphas);
    cudaFree(nav->alpha_parts); cudaFree(nav->z); cudaFree(nav->gauss); cudaFree(nav->pts);
    cudaFree(nav->ancestors); cudaFree(nav->cumw); cudaFree(nav->st); cudaFree(nav->vgrid); cudaFree(nav->zgrid);
    cudaFree(nav->vitems); cudaFree(nav->zitems); cudaFree(nav->scratch); cudaFree(nav->dump);
    cudaFree(nav->dump_count); cudaFree(nav->depth); cudaFree(nav->ana);
    for (auto& e : nav->pev) cudaEventDestroy(e);
    for (auto& e : nav->slot_ev) if (e) cudaEventDestroy(e);
    if (nav->stream) cudaStreamDestroy(nav->stream);
}

int set_device(rbphd_navigator* nav)
{
    CK(cudaSetDevice(nav->device));
    return RBPHD_OK;
}

KParams base_params(rbphd_navigator* nav, int mode, int M, int only_mapping, int slot = 0)
{
    KParams k;
    std::memset(&k, 0, sizeof k);
    k.cfg = nav->dcfg;
    k.lay = nav->lay;
    k.P = nav->P;
    k.first = 0;
    k.M = M;
    k.cap = nav->cap;
    k.mode = mode;
    k.only_mapping = only_mapping;
    k.maps[0] = nav->maps[0];  k.maps[1] = nav->maps[1];
    k.counts[0] = nav->counts[0];  k.counts[1] = nav->counts[1];
    k.poses = nav->poses;
    k.weights = nav->weights;
    k.alphas = nav->alphas;
    k.alpha_parts = nav->alpha_parts;
    k.z = nav->z + nav->zstride * (size_t)slot;
    k.vgrid = nav->vgrid;  k.vitems = nav->vitems;
    k.zgrid = nav->zgrid;  k.zitems = nav->zitems;
    k.scratch = nav->scratch;
    k.st = nav->st;
    k.dump = nav->dump;
    k.dump_count = nav->dump_count;
    k.dump_cap = nav->dump_cap;
    k.smem_sort_cap = nav->sort_cap;
    k.ll_flags = nav->ll_flags;
    k.holdout = nav->holdout;
    return k;
}

int check_async(rbphd_navigator* nav, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nav, RBPHD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return RBPHD_OK;
}

// prep + fused per-particle kernel
int enqueue_map_update(rbphd_navigator* nav, int M, int only_mapping, int mode, int slot = 0,
                       cudaEvent_t* ev = nullptr)
{
    if (mode == MODE_FRAME && nav->pending_wheel)
        return fail(nav, RBPHD_ERR_ARGUMENT,
                    "rbphd_slam_update_begin reported depleted particles: call rbphd_slam_update_finish first");
    KParams k = base_params(nav, mode, M, only_mapping, slot);
    launch_frame_prep(nav->stream, nav->dcfg, k.z, M, nav->vgrid, nav->vitems, nav->zgrid, nav->zitems, nav->pts);
    if (ev) cudaEventRecord(ev[0], nav->stream);
    int grid = std::min(nav->nslab, std::max(1, nav->P));
    launch_particle_update(nav->stream, k, grid, nav->smem);
    if (ev) cudaEventRecord(ev[1], nav->stream);
    nav->launches += 2;
    nav->M_last = M;
    return check_async(nav, "particle update launch");
}

int read_state(rbphd_navigator* nav, DeviceState* out)
{
    DeviceState* h = (DeviceState*)nav->h_state.get(sizeof(DeviceState));
    if (!h) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
    CK(cudaMemcpyAsync(h, nav->st, sizeof(DeviceState), cudaMemcpyDeviceToHost, nav->stream));
    CK(cudaStreamSynchronize(nav->stream));
    *out = *h;
    return RBPHD_OK;
}

int status_to_error(rbphd_navigator* nav, int status)
{
    if (!status) return RBPHD_OK;
    std::string msg = "capacity exceeded:";
    if (status & ST_OVER_COMPONENTS) msg += " components(max_components)";
    if (status & ST_OVER_PAIRS) msg += " gated-pairs(max_pairs)";
    if (status & ST_OVER_EDGES) msg += " merge-edges";
    if (status & ST_OVER_JMAP) msg += " map-estimate-size";
    if (status & ST_OVER_LL) msg += " likelihood-edges";
    if (status & ST_OVER_BLOCK) msg += " association-block-rows(>24)";
    if (status & ST_OVER_MURTY) msg += " murty-node-pool";
    return fail(nav, RBPHD_ERR_CAPACITY, msg);
}

int upload(rbphd_navigator* nav, void* dst, const void* src, size_t bytes, size_t stage_off = 0)
{
    if (bytes == 0) return RBPHD_OK;
    char* h = (char*)nav->h_in.get(stage_off + bytes);
    if (!h) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
    std::memcpy(h + stage_off, src, bytes);
    CK(cudaMemcpyAsync(dst, h + stage_off, bytes, cudaMemcpyHostToDevice, nav->stream));
    return RBPHD_OK;
}

// AoS (w[n], mean[n*3], cov[n*9]) -> the particle's SoA slab of buffer b
int write_map(rbphd_navigator* nav, int buffer, int particle, int n, const double* w, const double* mean,
              const double* cov)
{
    if (n > nav->cap) return fail(nav, RBPHD_ERR_CAPACITY, "map larger than max_components");
    const int cap = nav->cap;
    size_t bytes = sizeof(double) * kFields * cap + sizeof(int);
    // a dedicated staging area per call: synchronise before reuse
    CK(cudaStreamSynchronize(nav->stream));
    double* h = (double*)nav->h_map.get(bytes);
    if (!h) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
    std::memset(h, 0, sizeof(double) * kFields * cap);
    for (int i = 0; i < n; i++) {
        h[i] = w[i];
        for (int a = 0; a < 3; a++) h[(size_t)(1 + a) * cap + i] = mean[3 * i + a];
        for (int a = 0; a < 9; a++) h[(size_t)(4 + a) * cap + i] = cov[9 * i + a];
    }
    int* hc = (int*)(h + (size_t)kFields * cap);
    *hc = n;
    CK(cudaMemcpyAsync(nav->maps[buffer] + (size_t)particle * kFields * cap, h, sizeof(double) * kFields * cap,
                       cudaMemcpyHostToDevice, nav->stream));
    CK(cudaMemcpyAsync(nav->counts[buffer] + particle, hc, sizeof(int), cudaMemcpyHostToDevice, nav->stream));
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

int soa_to_outputs(rbphd_navigator* nav, const double* h, int stride, int n, const double** w, const double** mean,
                   const double** cov, int* on)
{
    nav->o_w.resize(std::max(n, 1));
    nav->o_m.resize(std::max(3 * n, 1));
    nav->o_P.resize(std::max(9 * n, 1));
    for (int i = 0; i < n; i++) {
        nav->o_w[i] = h[i];
        for (int a = 0; a < 3; a++) nav->o_m[3 * i + a] = h[(size_t)(1 + a) * stride + i];
        for (int a = 0; a < 9; a++) nav->o_P[9 * i + a] = h[(size_t)(4 + a) * stride + i];
    }
    if (w) *w = nav->o_w.data();
    if (mean) *mean = nav->o_m.data();
    if (cov) *cov = nav->o_P.data();
    if (on) *on = n;
    return RBPHD_OK;
}

int ensure_stage(rbphd_navigator* nav)
{
    if (nav->stage) return RBPHD_OK;
    rbphd_limits lim = nav->lim;
    lim.max_particles = 1;
    lim.max_components = nav->cap + nav->Mcap;   // a predicted map holds the prior components plus the births
    nav->stage = rbphd_new(&nav->cfg, &lim);
    if (!nav->stage) return fail(nav, RBPHD_ERR_CUDA, "stage navigator: " + g_last_error);
    {
        DevCfg& d = nav->stage->dcfg;
        d.depth = nav->dcfg.depth; d.resx = nav->dcfg.resx; d.resy = nav->dcfg.resy;
        d.half_x = nav->dcfg.half_x; d.half_y = nav->dcfg.half_y;
    }
    return RBPHD_OK;
}

}  // namespace

extern "C" {

const char* rbphd_last_error(const rbphd_navigator* nav)
{
    return nav ? nav->error.c_str() : g_last_error.c_str();
}

rbphd_navigator* rbphd_new(const rbphd_config* config, const rbphd_limits* limits)
{
    if (!config) { g_last_error = "null config"; return nullptr; }
    if (config->model != 0) { g_last_error = "only the PRM3D model (0) runs on the GPU"; return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_last_error = std::string("no CUDA device (") + cudaGetErrorString(e) + "); librbphd has no CPU fallback";
        return nullptr;
    }
    rbphd_navigator* nav = new rbphd_navigator();
    nav->cfg = *config;
    rbphd_limits lim{};
    if (limits) lim = *limits;
    nav->device = lim.device;
    if (lim.max_particles <= 0) lim.max_particles = 1;
    if (lim.max_components <= 0) lim.max_components = std::max(config->max_quantity, 64);
    lim.max_components = std::max(lim.max_components, std::min(config->max_quantity, 1 << 20));
    lim.max_components = (lim.max_components + 31) / 32 * 32;
    if (lim.max_measurements <= 0) lim.max_measurements = 1024;
    lim.max_measurements = (lim.max_measurements + 1) & ~1;
    if (lim.max_pairs <= 0) lim.max_pairs = 4 * lim.max_measurements;
    if (lim.resident_frames <= 0) lim.resident_frames = 1;
    nav->lim = lim;
    nav->slots = lim.resident_frames;
    nav->maxP = lim.max_particles;
    nav->cap = lim.max_components;
    nav->Mcap = lim.max_measurements;
    make_devcfg(*config, nav->dcfg, 0, false);
    make_layout(nav->lay, nav->cap, nav->Mcap, lim.max_pairs, config->max_quantity);

    auto bail = [&](const std::string& msg) -> rbphd_navigator* {
        g_last_error = msg;
        free_device(nav);
        delete nav;
        return nullptr;
    };
#define CKN(call)                                                                       \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) return bail(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

    CKN(cudaSetDevice(nav->device));
    cudaDeviceProp prop;
    CKN(cudaGetDeviceProperties(&prop, nav->device));
    if (prop.major < 10)
        return bail("device compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor) +
                    " < 10.0: librbphd is built for sm_100a only");
    CKN(cudaStreamCreateWithFlags(&nav->stream, cudaStreamNonBlocking));
    const size_t mapbytes = sizeof(double) * kFields * (size_t)nav->cap * nav->maxP;
    for (int b = 0; b < 2; b++) {
        CKN(cudaMalloc(&nav->maps[b], mapbytes));
        CKN(cudaMalloc(&nav->counts[b], sizeof(int) * nav->maxP));
        CKN(cudaMemsetAsync(nav->counts[b], 0, sizeof(int) * nav->maxP, nav->stream));
    }
    CKN(cudaMalloc(&nav->poses, sizeof(double) * 7 * nav->maxP));
    CKN(cudaMalloc(&nav->poses_tmp, sizeof(double) * 7 * nav->maxP));
    CKN(cudaMalloc(&nav->weights, sizeof(double) * nav->maxP));
    CKN(cudaMalloc(&nav->alphas, sizeof(double) * nav->maxP));
    CKN(cudaMalloc(&nav->alpha_parts, sizeof(double) * 8 * nav->maxP));
    nav->zstride = 3 * (size_t)nav->Mcap;
    nav->gstride = 6 * (size_t)nav->maxP + (6 * (size_t)nav->maxP) % 2;
    CKN(cudaMalloc(&nav->z, sizeof(double) * nav->zstride * nav->slots + 64));
    CKN(cudaMalloc(&nav->gauss, sizeof(double) * nav->gstride * nav->slots));
    CKN(cudaMalloc(&nav->pts, sizeof(double) * 6 * nav->Mcap + 64));
    CKN(cudaMalloc(&nav->ancestors, sizeof(int) * nav->maxP));
    CKN(cudaMalloc(&nav->cumw, sizeof(double) * ((size_t)nav->maxP + 1)));
    CKN(cudaMalloc(&nav->st, sizeof(DeviceState)));
    CKN(cudaMemsetAsync(nav->st, 0, sizeof(DeviceState), nav->stream));
    CKN(cudaMemsetAsync(nav->alphas, 0, sizeof(double) * nav->maxP, nav->stream));
    CKN(cudaMemsetAsync(nav->gauss, 0, sizeof(double) * nav->gstride * nav->slots, nav->stream));
    CKN(cudaMalloc(&nav->vgrid, sizeof(FrameGrid)));
    CKN(cudaMalloc(&nav->zgrid, sizeof(FrameGrid)));
    CKN(cudaMalloc(&nav->vitems, sizeof(int) * (nav->Mcap + 2)));
    CKN(cudaMalloc(&nav->zitems, sizeof(int) * (nav->Mcap + 2)));
    nav->dump_cap = nav->lay.cap_list;
    CKN(cudaMalloc(&nav->dump, sizeof(double) * kFields * (size_t)nav->dump_cap));
    CKN(cudaMalloc(&nav->dump_count, sizeof(int)));

    nav->smem = particle_update_smem(nav->Mcap, &nav->sort_cap);
    if (nav->smem > (size_t)prop.sharedMemPerBlockOptin)
        return bail("max_measurements too large for shared memory (" + std::to_string(nav->smem) + " B needed)");
    int per_sm = particle_update_max_ctas_per_sm(nav->smem);
    if (per_sm < 1) return bail("k_particle_update cannot be resident (shared memory / registers)");
    nav->ctas_per_sm = per_sm;
    nav->nslab = std::max(1, std::min(prop.multiProcessorCount * per_sm, nav->maxP));
    if (const char* lim_ctas = std::getenv("RBPHD_MAX_CTAS"))   // experiments: fewer persistent CTAs than the device holds
        nav->nslab = std::max(1, std::min(nav->nslab, std::atoi(lim_ctas)));
    CKN(cudaMalloc(&nav->scratch, nav->lay.bytes * (size_t)nav->nslab));
    CKN(cudaStreamSynchronize(nav->stream));
#undef CKN
    nav->P = 0;
    return nav;
}

void rbphd_delete(rbphd_navigator* nav)
{
    if (!nav) return;
    if (nav->stage) rbphd_delete(nav->stage);
    cudaSetDevice(nav->device);
    if (nav->stream) cudaStreamSynchronize(nav->stream);
    comm_free(nav);
    free_device(nav);
    delete nav;
}

int rbphd_particle_count(const rbphd_navigator* nav) { return nav ? nav->P : 0; }

int rbphd_set_depth_frame(rbphd_navigator* nav, const float* depth_xy, int resx, int resy)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    if (!depth_xy) {   // back to the plain pixel-range measurer
        nav->dcfg.depth = nullptr;
        if (nav->stage) nav->stage->dcfg.depth = nullptr;
        return RBPHD_OK;
    }
    if (resx < 1 || resy < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "depth frame resolution");
    const size_t n = (size_t)resx * resy;
    CK(cudaStreamSynchronize(nav->stream));
    if (n > nav->depth_cap) {
        cudaFree(nav->depth);
        nav->depth = nullptr;
        CK(cudaMalloc(&nav->depth, sizeof(float) * n));
        nav->depth_cap = n;
    }
    if (int r = upload(nav, nav->depth, depth_xy, sizeof(float) * n)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    nav->dcfg.depth = nav->depth;
    nav->dcfg.resx = resx; nav->dcfg.resy = resy;
    nav->dcfg.half_x = (double)((float)resx / 2.0f); nav->dcfg.half_y = (double)((float)resy / 2.0f);
    if (nav->stage) {   // the stage navigator evaluates with the same measurer
        DevCfg& d = nav->stage->dcfg;
        d.depth = nav->depth; d.resx = resx; d.resy = resy; d.half_x = nav->dcfg.half_x; d.half_y = nav->dcfg.half_y;
    }
    return RBPHD_OK;
}

int rbphd_set_holdout(rbphd_navigator* nav, int particle)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (particle >= nav->P) return fail(nav, RBPHD_ERR_ARGUMENT, "hold-out particle out of range");
    nav->holdout = particle < 0 ? -1 : particle;
    return RBPHD_OK;
}

int rbphd_reset(rbphd_navigator* nav, int particles, const double* pose7, int n, const double* w,
                const double* mean, const double* cov)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (particles < 1 || particles > nav->maxP) return fail(nav, RBPHD_ERR_ARGUMENT, "particle count out of range");
    if (n < 0 || n > nav->cap) return fail(nav, RBPHD_ERR_CAPACITY, "map larger than max_components");
    if (nav->comm && particles != nav->P)
        return fail(nav, RBPHD_ERR_ARGUMENT, "particle count differs from the communicator's block: rbphd_comm_destroy first");
    if (int r = set_device(nav)) return r;
    nav->P = particles;
    const int cap = nav->cap;
    // buffer 0 becomes current; particle 0 is uploaded, the rest replicated on the device
    DeviceState st0{};
    if (int r = upload(nav, nav->st, &st0, sizeof st0)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    if (int r = write_map(nav, 0, 0, n, w, mean, cov)) return r;
    for (int filled = 1; filled < particles;) {
        int cnt = std::min(filled, particles - filled);
        CK(cudaMemcpyAsync(nav->maps[0] + (size_t)filled * kFields * cap, nav->maps[0],
                           sizeof(double) * kFields * cap * (size_t)cnt, cudaMemcpyDeviceToDevice, nav->stream));
        CK(cudaMemcpyAsync(nav->counts[0] + filled, nav->counts[0], sizeof(int) * cnt, cudaMemcpyDeviceToDevice,
                           nav->stream));
        filled += cnt;
    }
    // PHD:245-266: every particle starts at 1 / P (the GLOBAL particle count when the particles are sharded)
    std::vector<double> hp(7 * (size_t)particles),
        hw(particles, 1.0 / ((nav->comm && nav->world > 1) ? nav->total : particles));
    for (int i = 0; i < particles; i++) std::memcpy(&hp[7 * (size_t)i], pose7, 7 * sizeof(double));
    if (int r = upload(nav, nav->poses, hp.data(), hp.size() * sizeof(double))) return r;
    CK(cudaStreamSynchronize(nav->stream));
    if (int r = upload(nav, nav->weights, hw.data(), hw.size() * sizeof(double))) return r;
    CK(cudaMemsetAsync(nav->alphas, 0, sizeof(double) * particles, nav->stream));
    CK(cudaStreamSynchronize(nav->stream));
    nav->o_anc.assign(particles, 0);
    for (int i = 0; i < particles; i++) nav->o_anc[i] = i;
    return RBPHD_OK;
}

int rbphd_clear_maps(rbphd_navigator* nav)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    for (int b = 0; b < 2; b++) CK(cudaMemsetAsync(nav->counts[b], 0, sizeof(int) * nav->maxP, nav->stream));
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

int rbphd_upload_frame_inputs(rbphd_navigator* nav, int slot, const double* gauss, const double* z, int m)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (m < 0 || m > nav->Mcap) return fail(nav, RBPHD_ERR_ARGUMENT, "m > max_measurements");
    if (slot < 0 || slot >= nav->slots) return fail(nav, RBPHD_ERR_ARGUMENT, "input slot out of range");
    if (int r = set_device(nav)) return r;
    const size_t gb = gauss ? sizeof(double) * 6 * (size_t)nav->P : 0;
    const size_t zb = z ? sizeof(double) * 3 * (size_t)m : 0;
    if (gb + zb == 0) return RBPHD_OK;
    // every input slot has its own pinned staging area and an event that says when its last copy has left it: the
    // upload of frame t+1 does not wait for frame t's kernels (no stream synchronisation here)
    if ((int)nav->slot_stage.size() < nav->slots) {
        nav->slot_stage.resize(nav->slots);
        nav->slot_ev.resize(nav->slots, nullptr);
    }
    if (nav->slot_ev[slot]) CK(cudaEventSynchronize(nav->slot_ev[slot]));
    else CK(cudaEventCreateWithFlags(&nav->slot_ev[slot], cudaEventDisableTiming));
    char* h = (char*)nav->slot_stage[slot].get(sizeof(double) * (6 * (size_t)nav->maxP + 3 * (size_t)nav->Mcap) + 16);
    if (!h) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
    if (gauss) {
        std::memcpy(h, gauss, gb);
        CK(cudaMemcpyAsync(nav->gauss + nav->gstride * (size_t)slot, h, gb, cudaMemcpyHostToDevice, nav->stream));
    }
    if (z) {
        std::memcpy(h + gb, z, zb);
        CK(cudaMemcpyAsync(nav->z + nav->zstride * (size_t)slot, h + gb, zb, cudaMemcpyHostToDevice, nav->stream));
    }
    CK(cudaEventRecord(nav->slot_ev[slot], nav->stream));
    return RBPHD_OK;
}

int rbphd_update(rbphd_navigator* nav, const double* reading6, double dt, const double* gauss, int perfect_still)
{
    if (!nav || !reading6) return RBPHD_ERR_ARGUMENT;
    if (nav->P < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "navigator not reset");
    if (int r = set_device(nav)) return r;
    if (gauss) if (int r = rbphd_upload_frame_inputs(nav, 0, gauss, nullptr, 0)) return r;
    Reading6 rd;
    std::memcpy(rd.v, reading6, sizeof rd.v);
    launch_predict_pose(nav->stream, nav->dcfg, nav->P, nav->poses, rd, dt, nav->gauss, perfect_still);
    nav->launches += 1;
    if (int r = check_async(nav, "predict_pose launch")) return r;
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

int rbphd_update_async(rbphd_navigator* nav, int slot, const double* reading6, double dt, int perfect_still)
{
    if (!nav || !reading6) return RBPHD_ERR_ARGUMENT;
    if (nav->P < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "navigator not reset");
    if (slot < 0 || slot >= nav->slots) return fail(nav, RBPHD_ERR_ARGUMENT, "input slot out of range");
    if (int r = set_device(nav)) return r;
    Reading6 rd;
    std::memcpy(rd.v, reading6, sizeof rd.v);
    launch_predict_pose(nav->stream, nav->dcfg, nav->P, nav->poses, rd, dt, nav->gauss + nav->gstride * (size_t)slot,
                        perfect_still);
    nav->launches += 1;
    return check_async(nav, "predict_pose launch");
}

int rbphd_set_pose(rbphd_navigator* nav, int particle, const double* pose7)
{
    if (!nav || !pose7) return RBPHD_ERR_ARGUMENT;
    if (particle < 0 || particle >= nav->P) return fail(nav, RBPHD_ERR_ARGUMENT, "particle index out of range");
    if (int r = set_device(nav)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    if (int r = upload(nav, nav->poses + 7 * (size_t)particle, pose7, 7 * sizeof(double))) return r;
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

int rbphd_set_poses(rbphd_navigator* nav, const double* poses)
{
    if (!nav || !poses) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    if (int r = upload(nav, nav->poses, poses, 7 * sizeof(double) * (size_t)nav->P)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

static int download(rbphd_navigator* nav, const void* dev, size_t bytes, void** host)
{
    void* h = nav->h_out.get(std::max<size_t>(bytes, 16));
    if (!h) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
    CK(cudaMemcpyAsync(h, dev, bytes, cudaMemcpyDeviceToHost, nav->stream));
    CK(cudaStreamSynchronize(nav->stream));
    *host = h;
    return RBPHD_OK;
}

int rbphd_get_poses(rbphd_navigator* nav, const double** poses, int* particles)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    void* h;
    if (int r = download(nav, nav->poses, sizeof(double) * 7 * (size_t)nav->P, &h)) return r;
    if (poses) *poses = (const double*)h;
    if (particles) *particles = nav->P;
    return RBPHD_OK;
}

int rbphd_get_weights(rbphd_navigator* nav, const double** weights, int* particles)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    void* h;
    if (int r = download(nav, nav->weights, sizeof(double) * (size_t)nav->P, &h)) return r;
    if (weights) *weights = (const double*)h;
    if (particles) *particles = nav->P;
    return RBPHD_OK;
}

int rbphd_get_alphas(rbphd_navigator* nav, const double** alphas, int* particles)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    void* h;
    if (int r = download(nav, nav->alphas, sizeof(double) * (size_t)nav->P, &h)) return r;
    if (alphas) *alphas = (const double*)h;
    if (particles) *particles = nav->P;
    return RBPHD_OK;
}

int rbphd_set_weights(rbphd_navigator* nav, const double* weights)
{
    if (!nav || !weights) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    if (int r = upload(nav, nav->weights, weights, sizeof(double) * (size_t)nav->P)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

int rbphd_get_best(rbphd_navigator* nav, int* best)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    if (best) *best = st.best;
    return RBPHD_OK;
}

int rbphd_get_ancestors(rbphd_navigator* nav, const int** ancestors, int* particles)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    void* h;
    nav->o_anc.resize(std::max(nav->P, 1));
    if (int r = download(nav, nav->ancestors, sizeof(int) * (size_t)nav->P, &h)) return r;
    std::memcpy(nav->o_anc.data(), h, sizeof(int) * (size_t)nav->P);
    if (ancestors) *ancestors = nav->o_anc.data();
    if (particles) *particles = nav->P;
    return RBPHD_OK;
}

int rbphd_get_map_counts(rbphd_navigator* nav, const int** counts, int* particles)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    void* h;
    if (int r = download(nav, nav->counts[st.cur], sizeof(int) * (size_t)nav->P, &h)) return r;
    if (counts) *counts = (const int*)h;
    if (particles) *particles = nav->P;
    return RBPHD_OK;
}

int rbphd_get_map(rbphd_navigator* nav, int particle, const double** w, const double** mean, const double** cov,
                  int* n)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (particle < 0 || particle >= nav->P) return fail(nav, RBPHD_ERR_ARGUMENT, "particle index out of range");
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    void* hc;
    if (int r = download(nav, nav->counts[st.cur] + particle, sizeof(int), &hc)) return r;
    int cnt = std::min(*(int*)hc, nav->cap);
    void* h;
    if (int r = download(nav, nav->maps[st.cur] + (size_t)particle * kFields * nav->cap,
                         sizeof(double) * kFields * (size_t)nav->cap, &h))
        return r;
    return soa_to_outputs(nav, (const double*)h, nav->cap, cnt, w, mean, cov, n);
}

int rbphd_set_map(rbphd_navigator* nav, int particle, int n, const double* w, const double* mean, const double* cov)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (particle < 0 || particle >= nav->P) return fail(nav, RBPHD_ERR_ARGUMENT, "particle index out of range");
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    return write_map(nav, st.cur, particle, n, w, mean, cov);
}

int rbphd_synchronize(rbphd_navigator* nav)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    if (st.status) {
        int status = st.status;
        CK(cudaMemsetAsync(&nav->st->status, 0, sizeof(int), nav->stream));
        CK(cudaStreamSynchronize(nav->stream));
        return status_to_error(nav, status);
    }
    return RBPHD_OK;
}

static int enqueue_slam_tail(rbphd_navigator* nav, int only_mapping, double u, int force,
                             cudaEvent_t* ev = nullptr)
{
    if (only_mapping) {
        launch_flip(nav->stream, nav->st);
        nav->launches += 1;
        if (ev) { cudaEventRecord(ev[0], nav->stream); cudaEventRecord(ev[1], nav->stream); }
    }
    else {
        launch_normalize_resample(nav->stream, nav->dcfg, nav->P, nav->weights, u, force, nav->ancestors, nav->st, nav->cumw);
        if (ev) cudaEventRecord(ev[0], nav->stream);
        launch_copy_particles(nav->stream, nav->P, nav->cap, nav->maps, nav->counts, nav->poses, nav->poses_tmp,
                              nav->ancestors, nav->st);
        if (ev) cudaEventRecord(ev[1], nav->stream);
        nav->launches += 3;
    }
    return check_async(nav, "slam tail launch");
}

int rbphd_frame_async(rbphd_navigator* nav, int slot, const double* reading6, double dt, int perfect_still, int m,
                      int only_mapping, double u_resample)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (nav->P < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "navigator not reset");
    if (m < 0 || m > nav->Mcap) return fail(nav, RBPHD_ERR_ARGUMENT, "m > max_measurements");
    if (slot < 0 || slot >= nav->slots) return fail(nav, RBPHD_ERR_ARGUMENT, "input slot out of range");
    if (int r = set_device(nav)) return r;
    cudaEvent_t* ev = nullptr;
    if (nav->prof_frames < nav->prof_max) ev = &nav->pev[6 * (size_t)nav->prof_frames++];
    if (ev) cudaEventRecord(ev[0], nav->stream);
    if (reading6 && !only_mapping) {
        Reading6 rd;
        std::memcpy(rd.v, reading6, sizeof rd.v);
        launch_predict_pose(nav->stream, nav->dcfg, nav->P, nav->poses, rd, dt,
                            nav->gauss + nav->gstride * (size_t)slot, perfect_still);
        nav->launches += 1;
    }
    if (ev) cudaEventRecord(ev[1], nav->stream);
    if (int r = enqueue_map_update(nav, m, only_mapping, MODE_FRAME, slot, ev ? ev + 2 : nullptr)) return r;
    if (nav->comm && nav->world > 1 && !only_mapping) return comm_slam_tail(nav, u_resample, ev ? ev + 4 : nullptr);
    return enqueue_slam_tail(nav, only_mapping, u_resample, 0, ev ? ev + 4 : nullptr);
}

int rbphd_slam_update(rbphd_navigator* nav, const double* z, int m, int only_mapping, double u_resample, int* best,
                      int* resampled)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (nav->P < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "navigator not reset");
    if (m < 0 || m > nav->Mcap) return fail(nav, RBPHD_ERR_ARGUMENT, "m > max_measurements");
    if (m > 0 && !z) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    if (int r = rbphd_upload_frame_inputs(nav, 0, nullptr, z, m)) return r;
    if (int r = enqueue_map_update(nav, m, only_mapping, MODE_FRAME)) return r;
    if (nav->comm && nav->world > 1 && !only_mapping) { if (int r = comm_slam_tail(nav, u_resample, nullptr)) return r; }
    else if (int r = enqueue_slam_tail(nav, only_mapping, u_resample, 0)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    if (best) *best = st.best;
    if (resampled) *resampled = only_mapping ? 0 : st.resampled;
    if (st.status) {
        int status = st.status;
        CK(cudaMemsetAsync(&nav->st->status, 0, sizeof(int), nav->stream));
        CK(cudaStreamSynchronize(nav->stream));
        return status_to_error(nav, status);
    }
    return RBPHD_OK;
}

// SlamUpdate in two steps, so that the host draws the wheel's uniform exactly when the reference does
// (PHD:355-357 calls ResampleParticles, and with it Util.Uniform.Next() at PHD:727, only when depleted)
int rbphd_slam_update_begin(rbphd_navigator* nav, const double* z, int m, int only_mapping, int* best, int* depleted)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (nav->P < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "navigator not reset");
    if (m < 0 || m > nav->Mcap) return fail(nav, RBPHD_ERR_ARGUMENT, "m > max_measurements");
    if (m > 0 && !z) return RBPHD_ERR_ARGUMENT;
    if (nav->comm && nav->world > 1) return fail(nav, RBPHD_ERR_ARGUMENT, "two-step SlamUpdate is single-GPU: use rbphd_slam_update");
    if (int r = set_device(nav)) return r;
    if (int r = rbphd_upload_frame_inputs(nav, 0, nullptr, z, m)) return r;
    if (int r = enqueue_map_update(nav, m, only_mapping, MODE_FRAME)) return r;
    if (only_mapping) { launch_flip(nav->stream, nav->st); nav->launches += 1; }
    else {
        launch_normalize_resample(nav->stream, nav->dcfg, nav->P, nav->weights, 0.0, 3, nav->ancestors, nav->st, nav->cumw);
        nav->launches += 1;
    }
    if (int r = check_async(nav, "slam update (begin)")) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    if (best) *best = st.best;
    if (depleted) *depleted = only_mapping ? 0 : st.depleted;
    nav->pending_wheel = (!only_mapping && st.depleted && !st.status) ? 1 : 0;
    if (st.status) {
        int status = st.status;
        CK(cudaMemsetAsync(&nav->st->status, 0, sizeof(int), nav->stream));
        CK(cudaStreamSynchronize(nav->stream));
        return status_to_error(nav, status);
    }
    return RBPHD_OK;
}

int rbphd_slam_update_finish(rbphd_navigator* nav, double u_resample, int* best)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    if (nav->pending_wheel) {
        nav->pending_wheel = 0;
        launch_normalize_resample(nav->stream, nav->dcfg, nav->P, nav->weights, u_resample, 2, nav->ancestors, nav->st, nav->cumw);
        launch_copy_particles(nav->stream, nav->P, nav->cap, nav->maps, nav->counts, nav->poses, nav->poses_tmp,
                              nav->ancestors, nav->st);
        nav->launches += 3;
        if (int r = check_async(nav, "slam update (finish)")) return r;
    }
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    if (best) *best = st.best;
    return RBPHD_OK;
}

int rbphd_resample(rbphd_navigator* nav, double u_resample)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (nav->P < 1) return fail(nav, RBPHD_ERR_ARGUMENT, "navigator not reset");
    if (int r = set_device(nav)) return r;
    {
        DeviceState st;   // an unacknowledged capacity error leaves the maps unpublished: report it first
        if (int r = read_state(nav, &st)) return r;
        if (st.status) return status_to_error(nav, st.status);
    }
    // the copy kernel moves particles from buffer 1-cur to buffer cur: make the current maps "1-cur" first
    launch_flip(nav->stream, nav->st);
    launch_normalize_resample(nav->stream, nav->dcfg, nav->P, nav->weights, u_resample, 2, nav->ancestors, nav->st, nav->cumw);
    launch_copy_particles(nav->stream, nav->P, nav->cap, nav->maps, nav->counts, nav->poses, nav->poses_tmp,
                          nav->ancestors, nav->st);
    nav->launches += 4;
    if (int r = check_async(nav, "resample launch")) return r;
    CK(cudaStreamSynchronize(nav->stream));
    return RBPHD_OK;
}

int rbphd_particle_depleted(rbphd_navigator* nav, int* depleted)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    const double* w;
    int P;
    if (int r = rbphd_get_weights(nav, &w, &P)) return r;
    double cum = 0;   // PHD:768-777 on the host mirror (a query, not part of the frame path)
    for (int i = 0; i < P; i++) cum += w[i] * w[i];
    if (depleted) *depleted = (1.0 / cum < nav->cfg.min_effective_particle * P) ? 1 : 0;
    return RBPHD_OK;
}

// ------------------------------------------------------------------ stage entry points
static int stage_run(rbphd_navigator* nav, rbphd_navigator** sp, const double* pose7, int n, const double* w,
                     const double* mean, const double* cov, const double* z, int m)
{
    if (int r = ensure_stage(nav)) return r;
    rbphd_navigator* s = nav->stage;
    *sp = s;
    static const double ident[7] = {0, 0, 0, 1, 0, 0, 0};
    if (int r = rbphd_reset(s, 1, pose7 ? pose7 : ident, n, w, mean, cov)) return fail(nav, r, s->error);
    if (m > s->Mcap) return fail(nav, RBPHD_ERR_ARGUMENT, "m > max_measurements");
    if (int r = rbphd_upload_frame_inputs(s, 0, nullptr, z, m)) return fail(nav, r, s->error);
    return RBPHD_OK;
}

static int stage_finish_dump(rbphd_navigator* nav, rbphd_navigator* s, const double** ow, const double** om,
                             const double** oP, int* on)
{
    DeviceState st;
    if (int r = read_state(s, &st)) return fail(nav, r, s->error);
    void* hc;
    if (int r = download(s, s->dump_count, sizeof(int), &hc)) return fail(nav, r, s->error);
    int cnt = *(int*)hc;
    if (cnt > s->dump_cap) return fail(nav, RBPHD_ERR_CAPACITY, "stage output larger than the dump buffer");
    void* h;
    if (int r = download(s, s->dump, sizeof(double) * kFields * (size_t)s->dump_cap, &h)) return fail(nav, r, s->error);
    soa_to_outputs(nav, (const double*)h, s->dump_cap, cnt, ow, om, oP, on);
    if (st.status) {
        cudaMemsetAsync(&s->st->status, 0, sizeof(int), s->stream);
        cudaStreamSynchronize(s->stream);
        return status_to_error(nav, st.status);
    }
    return RBPHD_OK;
}

int rbphd_stage_predict(rbphd_navigator* nav, const double* pose7, int n, const double* w, const double* mean,
                        const double* cov, const double* z, int m, const double** ow, const double** omean,
                        const double** ocov, int* on)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* s;
    if (int r = stage_run(nav, &s, pose7, n, w, mean, cov, z, m)) return r;
    if (int r = enqueue_map_update(s, m, 1, MODE_STAGE_PREDICT)) return fail(nav, r, s->error);
    return stage_finish_dump(nav, s, ow, omean, ocov, on);
}

int rbphd_stage_correct(rbphd_navigator* nav, const double* pose7, int n, const double* w, const double* mean,
                        const double* cov, const double* z, int m, double gate_radius, const double** ow,
                        const double** omean, const double** ocov, int* on)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* s;
    if (int r = stage_run(nav, &s, pose7, n, w, mean, cov, z, m)) return r;
    DevCfg saved = s->dcfg;
    make_devcfg(s->cfg, s->dcfg, gate_radius, true);
    s->dcfg.depth = saved.depth; s->dcfg.resx = saved.resx; s->dcfg.resy = saved.resy;
    s->dcfg.half_x = saved.half_x; s->dcfg.half_y = saved.half_y;
    if ((long long)n * m + n > s->dump_cap && gate_radius < 0) {
        s->dcfg = saved;
        return fail(nav, RBPHD_ERR_CAPACITY, "ungated stage_correct output exceeds max_pairs");
    }
    int r = enqueue_map_update(s, m, 1, MODE_STAGE_CORRECT);
    s->dcfg = saved;
    if (r) return fail(nav, r, s->error);
    return stage_finish_dump(nav, s, ow, omean, ocov, on);
}

int rbphd_stage_prune(rbphd_navigator* nav, int n, const double* w, const double* mean, const double* cov,
                      const double** ow, const double** omean, const double** ocov, int* on)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* s;
    if (int r = stage_run(nav, &s, nullptr, n, w, mean, cov, nullptr, 0)) return r;
    if (int r = enqueue_map_update(s, 0, 1, MODE_STAGE_PRUNE)) return fail(nav, r, s->error);
    launch_flip(s->stream, s->st);
    s->launches += 1;
    DeviceState st;
    if (int r = read_state(s, &st)) return fail(nav, r, s->error);
    const double *tw, *tm, *tP;
    int tn;
    if (int r = rbphd_get_map(s, 0, &tw, &tm, &tP, &tn)) return fail(nav, r, s->error);
    nav->o_w.assign(tw, tw + std::max(tn, 1));
    nav->o_m.assign(tm, tm + std::max(3 * tn, 1));
    nav->o_P.assign(tP, tP + std::max(9 * tn, 1));
    if (ow) *ow = nav->o_w.data();
    if (omean) *omean = nav->o_m.data();
    if (ocov) *ocov = nav->o_P.data();
    if (on) *on = tn;
    if (st.status) {
        cudaMemsetAsync(&s->st->status, 0, sizeof(int), s->stream);
        cudaStreamSynchronize(s->stream);
        return status_to_error(nav, st.status);
    }
    return RBPHD_OK;
}

int rbphd_stage_weight_alpha(rbphd_navigator* nav, const double* pose7, const double* z, int m, int np,
                             const double* pw, const double* pmean, const double* pcov, int nc, const double* cw,
                             const double* cmean, const double* ccov, double* out7)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* s;
    if (int r = stage_run(nav, &s, pose7, np, pw, pmean, pcov, z, m)) return r;
    if (int r = write_map(s, 1, 0, nc, cw, cmean, ccov)) return fail(nav, r, s->error);
    if (int r = enqueue_map_update(s, m, 0, MODE_STAGE_WEIGHT)) return fail(nav, r, s->error);
    DeviceState st;
    if (int r = read_state(s, &st)) return fail(nav, r, s->error);
    void* h;
    if (int r = download(s, s->alpha_parts, sizeof(double) * 8, &h)) return fail(nav, r, s->error);
    if (out7) std::memcpy(out7, h, 7 * sizeof(double));
    if (st.status) {
        cudaMemsetAsync(&s->st->status, 0, sizeof(int), s->stream);
        cudaStreamSynchronize(s->stream);
        return status_to_error(nav, st.status);
    }
    return RBPHD_OK;
}

// the static likelihood functions of PHDNavigator over a landmark list (PHD:395-460, 526-559)
static int stage_setll(rbphd_navigator* nav, const double* pose7, int j, const double* jmean, const double* z, int m,
                       int flags, double* loglik, rbphd_navigator** sp)
{
    std::vector<double> w(std::max(j, 1), 1.0), P(9 * (size_t)std::max(j, 1), 0.0);
    for (int i = 0; i < j; i++) P[9 * (size_t)i] = P[9 * (size_t)i + 4] = P[9 * (size_t)i + 8] = 1.0;
    rbphd_navigator* s;
    if (int r = stage_run(nav, &s, pose7, j, w.data(), jmean, P.data(), z, m)) return r;
    if (sp) *sp = s;
    if (j > s->lay.cap_j) return fail(nav, RBPHD_ERR_CAPACITY, "landmark list larger than the map-estimate capacity");
    s->ll_flags = flags;
    int rc = enqueue_map_update(s, m, 0, MODE_STAGE_SETLL);
    s->ll_flags = 0;
    if (rc) return fail(nav, rc, s->error);
    DeviceState st;
    if (int r = read_state(s, &st)) return fail(nav, r, s->error);
    void* h;
    if (int r = download(s, s->alphas, sizeof(double), &h)) return fail(nav, r, s->error);
    if (loglik) *loglik = *(double*)h;
    if (st.status) {
        cudaMemsetAsync(&s->st->status, 0, sizeof(int), s->stream);
        cudaStreamSynchronize(s->stream);
        return status_to_error(nav, st.status);
    }
    return RBPHD_OK;
}

int rbphd_set_likelihood(rbphd_navigator* nav, const double* pose7, int j, const double* jmean, const double* z, int m,
                         double* likelihood)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    double ll = 0;
    if (int r = stage_setll(nav, pose7, j, jmean, z, m, 0, &ll, nullptr)) return r;
    if (likelihood) *likelihood = std::exp(ll);   // PHD:405
    return RBPHD_OK;
}

int rbphd_quasi_set_loglikelihood(rbphd_navigator* nav, const double* pose7, int j, const double* jmean,
                                  const double* z, int m, double* loglik)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    return stage_setll(nav, pose7, j, jmean, z, m, LL_QUASI, loglik, nullptr);
}

int rbphd_quasi_set_loglikelihood_gradient(rbphd_navigator* nav, const double* pose7, int j, const double* jmean,
                                           const double* z, int m, int sum_normalised, double* loglik, double* gradient6)
{
    if (!nav || !gradient6) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* s = nullptr;
    const int flags = LL_QUASI | LL_GRADIENT | (sum_normalised ? LL_TEMPERED_SUM : 0);
    if (int r = stage_setll(nav, pose7, j, jmean, z, m, flags, loglik, &s)) return r;
    void* h;
    if (int r = download(s, s->alpha_parts, sizeof(double) * 8, &h)) return fail(nav, r, s->error);
    std::memcpy(gradient6, h, 6 * sizeof(double));
    return RBPHD_OK;
}

int rbphd_set_loglike_matrix(rbphd_navigator* nav, const double* pose7, int j, const double* jmean, const double* z,
                             int m, const int** rows, const int** cols, const double** vals, int* nnz)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* s = nullptr;
    if (int r = stage_setll(nav, pose7, j, jmean, z, m, LL_DUMP_MATRIX, nullptr, &s)) return r;
    void* hc;
    if (int r = download(s, s->dump_count, sizeof(int), &hc)) return fail(nav, r, s->error);
    const int cnt = *(int*)hc;
    if (cnt < 0) return fail(nav, RBPHD_ERR_CAPACITY, "likelihood matrix larger than the dump buffer");
    void* h;
    if (int r = download(s, s->dump, sizeof(double) * 3 * (size_t)std::max(cnt, 1), &h)) return fail(nav, r, s->error);
    const double* t = (const double*)h;
    // detection entries come off the device in arbitrary order: sort by (row, column)
    std::vector<int> order(cnt);
    for (int e = 0; e < cnt; e++) order[e] = e;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (t[3 * (size_t)a] != t[3 * (size_t)b]) return t[3 * (size_t)a] < t[3 * (size_t)b];
        return t[3 * (size_t)a + 1] < t[3 * (size_t)b + 1];
    });
    nav->o_rows.resize(std::max(cnt, 1));
    nav->o_cols.resize(std::max(cnt, 1));
    nav->o_w.resize(std::max(cnt, 1));
    for (int e = 0; e < cnt; e++) {
        const size_t q = 3 * (size_t)order[e];
        nav->o_rows[e] = (int)t[q]; nav->o_cols[e] = (int)t[q + 1]; nav->o_w[e] = t[q + 2];
    }
    if (rows) *rows = nav->o_rows.data();
    if (cols) *cols = nav->o_cols.data();
    if (vals) *vals = nav->o_w.data();
    if (nnz) *nnz = cnt;
    return RBPHD_OK;
}

// ------------------------------------------------------------------ around the hot path
static int analysis_scratch(rbphd_navigator* nav, size_t doubles)
{
    if (doubles <= nav->ana_cap) return RBPHD_OK;
    CK(cudaStreamSynchronize(nav->stream));
    cudaFree(nav->ana);
    nav->ana = nullptr; nav->ana_cap = 0;
    CK(cudaMalloc(&nav->ana, sizeof(double) * doubles));
    nav->ana_cap = doubles;
    return RBPHD_OK;
}

int rbphd_generate_measurements(rbphd_navigator* nav, const double* pose7, const double* landmarks, int n,
                                const double* uniforms, const double* gauss, const double* chol9,
                                const double* clutter_u, int nc, double* z, int* assoc, int* count)
{
    if (!nav || !pose7 || n < 0 || nc < 0 || !z || !count) return RBPHD_ERR_ARGUMENT;
    if (n > 0 && (!landmarks || !uniforms || !gauss)) return RBPHD_ERR_ARGUMENT;
    if (nc > 0 && !clutter_u) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    double chol[9] = {0};
    if (chol9) std::memcpy(chol, chol9, sizeof chol);
    else {   // UTIL:173-202 on the 3 x 3 measurement covariance
        double R[9];
        std::memcpy(R, nav->cfg.R, sizeof R);
        for (int i = 0; i < 3; i++) if (R[i * 3 + i] < 1e-40) R[i * 3 + i] = 1e-40;
        for (int j = 0; j < 3; j++) {
            double d = R[j * 3 + j];
            for (int k = 0; k < j; k++) d -= chol[j * 3 + k] * chol[j * 3 + k];
            chol[j * 3 + j] = std::sqrt(d);
            for (int i = j + 1; i < 3; i++) {
                double t = R[i * 3 + j];
                for (int k = 0; k < j; k++) t -= chol[i * 3 + k] * chol[j * 3 + k];
                chol[i * 3 + j] = t / chol[j * 3 + j];
            }
        }
    }
    // device layout: pose 7 | chol 9 | landmarks 3n | uniforms n | gauss 3n | clutter 3nc | z 3(n+nc) | assoc, count
    const size_t N = (size_t)n, NC = (size_t)nc, T = N + NC;
    const size_t o_pose = 0, o_chol = 8, o_lm = 24, o_un = o_lm + 3 * N, o_ga = o_un + N, o_cl = o_ga + 3 * N,
                 o_z = o_cl + 3 * NC, o_as = o_z + 3 * T, total = o_as + (T + 2) / 2 + 2;
    if (int r = analysis_scratch(nav, total)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    std::vector<double> in(o_z, 0.0);
    std::memcpy(&in[o_pose], pose7, 7 * sizeof(double));
    std::memcpy(&in[o_chol], chol, sizeof chol);
    if (n) {
        std::memcpy(&in[o_lm], landmarks, 3 * N * sizeof(double));
        std::memcpy(&in[o_un], uniforms, N * sizeof(double));
        std::memcpy(&in[o_ga], gauss, 3 * N * sizeof(double));
    }
    if (nc) std::memcpy(&in[o_cl], clutter_u, 3 * NC * sizeof(double));
    if (int r = upload(nav, nav->ana, in.data(), o_z * sizeof(double))) return r;
    double* d = nav->ana;
    int* d_assoc = reinterpret_cast<int*>(d + o_as);
    int* d_count = d_assoc + T;
    launch_generate_measurements(nav->stream, nav->dcfg, d + o_pose, d + o_lm, n, d + o_un, d + o_ga, d + o_chol, d + o_cl,
                                 nc, d + o_z, d_assoc, d_count);
    CK(cudaGetLastError());
    nav->launches++;
    void* h;
    if (int r = download(nav, d + o_z, sizeof(double) * (total - o_z), &h)) return r;
    const double* hz = (const double*)h;
    const int* ha = reinterpret_cast<const int*>(hz + 3 * T);
    const int cnt = ha[T];
    std::memcpy(z, hz, sizeof(double) * 3 * (size_t)cnt);
    if (assoc) std::memcpy(assoc, ha, sizeof(int) * (size_t)cnt);
    *count = cnt;
    return RBPHD_OK;
}

int rbphd_ospa(rbphd_navigator* nav, const double* a, int na, const double* b, int nb, double c, double p,
               double* ospa, double* cardinality_error)
{
    if (!nav || na < 0 || nb < 0 || !ospa) return RBPHD_ERR_ARGUMENT;
    if ((na > 0 && !a) || (nb > 0 && !b)) return RBPHD_ERR_ARGUMENT;
    if (na > nb) { std::swap(na, nb); std::swap(a, b); }   // Plot.cs:533-537
    if (nb > 8192) return fail(nav, RBPHD_ERR_CAPACITY, "rbphd_ospa: more than 8192 landmarks");
    if (int r = set_device(nav)) return r;
    const size_t A = 3 * (size_t)na, B = 3 * (size_t)nb;
    const size_t o_b = A + (A & 1), o_work = o_b + B + (B & 1), o_out = o_work + ospa_workspace_doubles(nb);
    if (int r = analysis_scratch(nav, o_out + 2)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    std::vector<double> in(std::max<size_t>(o_work, 2), 0.0);
    if (na) std::memcpy(&in[0], a, A * sizeof(double));
    if (nb) std::memcpy(&in[o_b], b, B * sizeof(double));
    if (int r = upload(nav, nav->ana, in.data(), in.size() * sizeof(double))) return r;
    launch_ospa(nav->stream, nav->ana, na, nav->ana + o_b, nb, c, p, nav->ana + o_work, nav->ana + o_out);
    CK(cudaGetLastError());
    nav->launches++;
    void* h;
    if (int r = download(nav, nav->ana + o_out, 2 * sizeof(double), &h)) return r;
    *ospa = ((const double*)h)[0];
    if (cardinality_error) *cardinality_error = ((const double*)h)[1];
    return RBPHD_OK;
}

int rbphd_stage_set_loglikelihood(rbphd_navigator* nav, const double* pose7, int j, const double* jmean,
                                  const double* z, int m, double* loglik)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    return stage_setll(nav, pose7, j, jmean, z, m, 0, loglik, nullptr);
}

// ------------------------------------------------------------------ multi-GPU: particles sharded by rank
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process, e.g. the one a host
// runtime loaded, or the system's), so a single-GPU host needs no NCCL at all.
}  // extern "C"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char* nm) {
            void* p = dlsym(api.lib, nm);
            if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + nm;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}

#define CKNCCL(call)                                                                                     \
    do {                                                                                                 \
        ncclResult_t e__ = (call);                                                                       \
        if (e__ != ncclSuccess)                                                                          \
            return fail(nav, RBPHD_ERR_CUDA, std::string(#call) + ": " + nccl_api()->GetErrorString(e__)); \
    } while (0)

int part_lo_host(int r, int world, int total) { return (int)(((long long)total * r) / world); }

void comm_free(rbphd_navigator* nav)
{
    if (!nav->comm) return;
    cudaSetDevice(nav->device);
    if (nav->stream) cudaStreamSynchronize(nav->stream);
    nccl_api()->CommDestroy((ncclComm_t)nav->comm);
    nav->comm = nullptr;
    cudaFree(nav->gweights); cudaFree(nav->gcum); cudaFree(nav->ganc); cudaFree(nav->gcounts); cudaFree(nav->local_src);
    cudaFree(nav->rec_off); cudaFree(nav->send_idx); cudaFree(nav->send_off); cudaFree(nav->plan_hdr);
    cudaFree(nav->sendbuf); cudaFree(nav->recvbuf);
    nav->gweights = nullptr; nav->gcum = nullptr; nav->ganc = nullptr; nav->gcounts = nullptr; nav->local_src = nullptr;
    nav->rec_off = nullptr; nav->send_idx = nullptr; nav->send_off = nullptr; nav->plan_hdr = nullptr;
    nav->sendbuf = nullptr; nav->recvbuf = nullptr;
    nav->world = 1; nav->rank = 0; nav->total = 0;
}

// allgather of a per-particle array over the block partition (even partition: one ncclAllGather; otherwise one
// broadcast per rank in a group)
template <class T>
int comm_allgather(rbphd_navigator* nav, const T* local, T* global, ncclDataType_t dt)
{
    NcclApi* nc = nccl_api();
    ncclComm_t comm = (ncclComm_t)nav->comm;
    if (nav->total % nav->world == 0) {
        CKNCCL(nc->AllGather(local, global, (size_t)nav->P, dt, comm, nav->stream));
    }
    else {
        CKNCCL(nc->GroupStart());
        for (int r = 0; r < nav->world; r++) {
            const int lo = part_lo_host(r, nav->world, nav->total), hi = part_lo_host(r + 1, nav->world, nav->total);
            CKNCCL(nc->Broadcast(local, global + lo, (size_t)(hi - lo), dt, r, comm, nav->stream));
        }
        CKNCCL(nc->GroupEnd());
    }
    return RBPHD_OK;
}

// the coupled tail of SlamUpdate over all ranks (PHD:343-358): weight allgather, identical normalise / ESS /
// wheel on every rank, and -- only when resampling fired -- the exchange of the ancestors' records
int comm_slam_tail(rbphd_navigator* nav, double u, cudaEvent_t* ev)
{
    NcclApi* nc = nccl_api();
    ncclComm_t comm = (ncclComm_t)nav->comm;
    const int world = nav->world, rank = nav->rank, total = nav->total;
    if (int r = comm_allgather<double>(nav, nav->weights, nav->gweights, ncclDouble)) return r;
    launch_normalize_resample(nav->stream, nav->dcfg, total, nav->gweights, u, 0, nav->ganc, nav->st, nav->gcum);
    const int lo = part_lo_host(rank, world, total);
    CK(cudaMemcpyAsync(nav->weights, nav->gweights + lo, sizeof(double) * (size_t)nav->P, cudaMemcpyDeviceToDevice,
                       nav->stream));
    nav->launches += 1;
    if (ev) cudaEventRecord(ev[0], nav->stream);
    // the host has to know whether the exchange is needed: one small read-back per frame
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    nav->last_best = st.best;
    nav->last_resampled = st.resampled;
    if (st.resampled) {
        static const bool trace = std::getenv("RBPHD_COMM_TRACE") != nullptr;   // debugging: per-step wall times
        auto t_prev = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (!trace) return;
            cudaStreamSynchronize(nav->stream);
            auto now = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[rbphd comm rank %d] %-16s %8.3f ms\n", rank, what,
                         std::chrono::duration<double, std::milli>(now - t_prev).count());
            t_prev = now;
        };
        const int post = 1 - st.cur;   // the wheel does not publish: the posterior maps are in buffer 1-cur
        if (int r = comm_allgather<int>(nav, nav->counts[post], nav->gcounts, ncclInt32)) return r;
        lap("allgather counts");
        launch_migration_plan(nav->stream, nav->ganc, nav->gcounts, total, world, rank, nav->local_src, nav->rec_off,
                              nav->send_idx, nav->send_off, nav->plan_hdr);
        long long* hdr = (long long*)nav->h_plan.get(sizeof(long long) * (2 * kMaxCommRanks + 4));
        if (!hdr) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
        CK(cudaMemcpyAsync(hdr, nav->plan_hdr, sizeof(long long) * (2 * world + 3), cudaMemcpyDeviceToHost, nav->stream));
        CK(cudaStreamSynchronize(nav->stream));
        lap("plan");
        if (!hdr[2 * world + 2]) return fail(nav, RBPHD_ERR_GENERIC, "migration plan: ancestors are not sorted");
        long long send_total = 0, recv_total = 0;
        for (int r = 0; r < world; r++) { send_total += hdr[r]; recv_total += hdr[world + r]; }
        if ((size_t)send_total > nav->sendcap || (size_t)recv_total > nav->recvcap)
            return fail(nav, RBPHD_ERR_CAPACITY, "migration buffers too small");
        launch_pack_records(nav->stream, nav->cap, nav->maps[post], nav->counts[post], nav->poses, nav->send_idx,
                            nav->send_off, (int)hdr[2 * world], nav->sendbuf);
        lap("pack");
        CKNCCL(nc->GroupStart());
        long long soff = 0, roff = 0;
        for (int r = 0; r < world; r++) {
            if (hdr[r] > 0) CKNCCL(nc->Send(nav->sendbuf + soff, (size_t)hdr[r], ncclDouble, r, comm, nav->stream));
            if (hdr[world + r] > 0) CKNCCL(nc->Recv(nav->recvbuf + roff, (size_t)hdr[world + r], ncclDouble, r, comm, nav->stream));
            soff += hdr[r];
            roff += hdr[world + r];
        }
        CKNCCL(nc->GroupEnd());
        lap("send/recv");
        launch_unpack_records(nav->stream, nav->P, nav->cap, nav->maps[post], nav->counts[post], nav->maps[st.cur],
                              nav->counts[st.cur], nav->poses, nav->poses_tmp, nav->local_src, nav->rec_off, nav->recvbuf);
        launch_copy_doubles(nav->stream, 7 * (size_t)nav->P, nav->poses, nav->poses_tmp);
        lap("unpack");
        nav->launches += 4;
        nav->comm_resamples += 1;
        nav->comm_sent_bytes += 8 * send_total;
        nav->comm_recv_bytes += 8 * recv_total;
        nav->comm_records += hdr[2 * world];
    }
    if (ev) cudaEventRecord(ev[1], nav->stream);
    return check_async(nav, "multi-GPU slam tail");
}

}  // namespace

extern "C" {

int rbphd_comm_unique_id(unsigned char out128[128])
{
    if (!out128) return RBPHD_ERR_ARGUMENT;
    NcclApi* nc = nccl_api();
    if (!nc->error.empty()) { g_last_error = nc->error; return RBPHD_ERR_GENERIC; }
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t e = nc->GetUniqueId(&id);
    if (e != ncclSuccess) { g_last_error = std::string("ncclGetUniqueId: ") + nc->GetErrorString(e); return RBPHD_ERR_GENERIC; }
    std::memcpy(out128, &id, 128);
    return RBPHD_OK;
}

int rbphd_comm_init_rank(rbphd_navigator* nav, const unsigned char id128[128], int rank, int world, int total_particles)
{
    if (!nav || !id128) return RBPHD_ERR_ARGUMENT;
    if (world < 1 || world > kMaxCommRanks || rank < 0 || rank >= world)
        return fail(nav, RBPHD_ERR_ARGUMENT, "rank / world out of range");
    const int lo = part_lo_host(rank, world, total_particles), hi = part_lo_host(rank + 1, world, total_particles);
    if (hi - lo != nav->P || nav->P < 1)
        return fail(nav, RBPHD_ERR_ARGUMENT, "reset the navigator with this rank's block of particles first (block partition)");
    NcclApi* nc = nccl_api();
    if (!nc->error.empty()) return fail(nav, RBPHD_ERR_GENERIC, nc->error);
    if (int r = set_device(nav)) return r;
    comm_free(nav);
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm;
    CKNCCL(nc->CommInitRank(&comm, world, id, rank));
    nav->comm = comm;
    nav->rank = rank; nav->world = world; nav->total = total_particles;
    const size_t rec = 8 + (size_t)kFields * nav->cap;
    nav->sendcap = rec * ((size_t)nav->P + world);
    nav->recvcap = rec * (size_t)nav->P;
    CK(cudaMalloc(&nav->gweights, sizeof(double) * (size_t)total_particles));
    CK(cudaMalloc(&nav->gcum, sizeof(double) * ((size_t)total_particles + 1)));
    CK(cudaMalloc(&nav->ganc, sizeof(int) * (size_t)total_particles));
    CK(cudaMalloc(&nav->gcounts, sizeof(int) * (size_t)total_particles));
    CK(cudaMalloc(&nav->local_src, sizeof(int) * (size_t)nav->P));
    CK(cudaMalloc(&nav->rec_off, sizeof(long long) * (size_t)nav->P));
    CK(cudaMalloc(&nav->send_idx, sizeof(int) * ((size_t)nav->P + world)));
    CK(cudaMalloc(&nav->send_off, sizeof(long long) * ((size_t)nav->P + world)));
    CK(cudaMalloc(&nav->plan_hdr, sizeof(long long) * (2 * kMaxCommRanks + 4)));
    CK(cudaMalloc(&nav->sendbuf, sizeof(double) * nav->sendcap));
    CK(cudaMalloc(&nav->recvbuf, sizeof(double) * nav->recvcap));
    if (!nav->h_plan.get(sizeof(long long) * (2 * kMaxCommRanks + 4))) return fail(nav, RBPHD_ERR_CUDA, "pinned allocation failed");
    // the reference starts every particle at 1 / P_total (PHD:245-266), whatever the size of this rank's block
    launch_fill_doubles(nav->stream, (size_t)nav->P, nav->weights, 1.0 / total_particles);
    // NCCL connects its point-to-point channels lazily, a few hundred ms each the first time a channel to a peer
    // carries data (measured: every other exchange of the first dozen).  Pay that here, not in a resampling
    // frame: a few all-to-all rounds over the (still empty) exchange buffers, sized to use every channel.
    if (world > 1) {
        // (the same size on every rank, also when the block partition is uneven: counts of a send and its receive
        // must agree)
        size_t warm = rec * (size_t)(total_particles / world) / (size_t)world;
        warm = std::min(warm, (size_t)8 << 20);   // 64 MB per peer
        if (warm > 0) {
            CK(cudaMemsetAsync(nav->sendbuf, 0, sizeof(double) * warm * (size_t)world, nav->stream));
            for (int rep = 0; rep < 16; rep++) {
                CKNCCL(nc->GroupStart());
                for (int r = 0; r < world; r++) {
                    if (r == rank) continue;
                    CKNCCL(nc->Send(nav->sendbuf + warm * (size_t)r, warm, ncclDouble, r, comm, nav->stream));
                    CKNCCL(nc->Recv(nav->recvbuf + warm * (size_t)r, warm, ncclDouble, r, comm, nav->stream));
                }
                CKNCCL(nc->GroupEnd());
            }
        }
    }
    CK(cudaStreamSynchronize(nav->stream));
    nav->comm_resamples = nav->comm_sent_bytes = nav->comm_recv_bytes = nav->comm_records = 0;
    return RBPHD_OK;
}

int rbphd_comm_destroy(rbphd_navigator* nav)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    comm_free(nav);
    return RBPHD_OK;
}

int rbphd_comm_stats(const rbphd_navigator* nav, int64_t out4[4])
{
    if (!nav || !out4) return RBPHD_ERR_ARGUMENT;
    out4[0] = nav->comm_resamples; out4[1] = nav->comm_sent_bytes; out4[2] = nav->comm_recv_bytes;
    out4[3] = nav->comm_records;
    return RBPHD_OK;
}

int rbphd_frame_result(rbphd_navigator* nav, int* best, int* resampled)
{
    if (!nav) return RBPHD_ERR_ARGUMENT;
    if (nav->world > 1 && nav->comm) {   // known to the host since the frame's read-back
        if (best) *best = nav->last_best;
        if (resampled) *resampled = nav->last_resampled;
        return RBPHD_OK;
    }
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    if (best) *best = st.best;
    if (resampled) *resampled = st.resampled;
    return RBPHD_OK;
}

// the migration plan of one rank for a given global ancestor / count vector (host arrays in, host arrays out):
// what the frame path computes on the device, exposed for tests
int rbphd_debug_migration_plan(int device, const int* ancestors, const int* counts, int total, int world, int rank,
                               int* local_src, int64_t* rec_off, int* send_idx, int64_t* send_off, int64_t* hdr)
{
    if (!ancestors || !counts || !local_src || !rec_off || !send_idx || !send_off || !hdr) return RBPHD_ERR_ARGUMENT;
    if (world < 1 || world > kMaxCommRanks || rank < 0 || rank >= world || total < world) return RBPHD_ERR_ARGUMENT;
    rbphd_navigator* nav = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return RBPHD_ERR_NO_DEVICE;
    const int lo = part_lo_host(rank, world, total), hi = part_lo_host(rank + 1, world, total), Pl = hi - lo;
    int *d_anc = nullptr, *d_cnt = nullptr, *d_ls = nullptr, *d_si = nullptr;
    long long *d_ro = nullptr, *d_so = nullptr, *d_hdr = nullptr;
    CK(cudaMalloc(&d_anc, sizeof(int) * total));
    CK(cudaMalloc(&d_cnt, sizeof(int) * total));
    CK(cudaMalloc(&d_ls, sizeof(int) * std::max(Pl, 1)));
    CK(cudaMalloc(&d_ro, sizeof(long long) * std::max(Pl, 1)));
    CK(cudaMalloc(&d_si, sizeof(int) * (Pl + world)));
    CK(cudaMalloc(&d_so, sizeof(long long) * (Pl + world)));
    CK(cudaMalloc(&d_hdr, sizeof(long long) * (2 * world + 3)));
    CK(cudaMemcpy(d_anc, ancestors, sizeof(int) * total, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cnt, counts, sizeof(int) * total, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_si, 0xff, sizeof(int) * (Pl + world)));
    CK(cudaMemset(d_so, 0xff, sizeof(long long) * (Pl + world)));
    launch_migration_plan(nullptr, d_anc, d_cnt, total, world, rank, d_ls, d_ro, d_si, d_so, d_hdr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(local_src, d_ls, sizeof(int) * Pl, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rec_off, d_ro, sizeof(long long) * Pl, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(send_idx, d_si, sizeof(int) * (Pl + world), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(send_off, d_so, sizeof(long long) * (Pl + world), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdr, d_hdr, sizeof(long long) * (2 * world + 3), cudaMemcpyDeviceToHost));
    cudaFree(d_anc); cudaFree(d_cnt); cudaFree(d_ls); cudaFree(d_ro); cudaFree(d_si); cudaFree(d_so); cudaFree(d_hdr);
    return RBPHD_OK;
}

int rbphd_get_phase_cycles(rbphd_navigator* nav, int64_t out80[80])
{
    if (!nav || !out80) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    for (int a = 0; a < 64; a++) out80[a] = (int64_t)st.phase_cycles[a];
    for (int a = 0; a < 16; a++) out80[64 + a] = (int64_t)st.dbg[a];
    return RBPHD_OK;
}

int64_t rbphd_kernel_launches(const rbphd_navigator* nav) { return nav ? nav->launches : 0; }

int rbphd_profile_enable(rbphd_navigator* nav, int max_frames)
{
    if (!nav || max_frames < 0) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    while ((int)nav->pev.size() < 6 * max_frames) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        nav->pev.push_back(e);
    }
    nav->prof_max = max_frames;
    nav->prof_frames = 0;
    return RBPHD_OK;
}

int rbphd_profile_read(rbphd_navigator* nav, double* ms, int max_frames, int* frames)
{
    if (!nav || !ms) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    CK(cudaStreamSynchronize(nav->stream));
    int n = std::min(nav->prof_frames, max_frames);
    for (int f = 0; f < n; f++)
        for (int st = 0; st < RBPHD_STAGES; st++) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, nav->pev[6 * (size_t)f + st], nav->pev[6 * (size_t)f + st + 1]));
            ms[(size_t)f * RBPHD_STAGES + st] = t;
        }
    if (frames) *frames = n;
    nav->prof_frames = 0;
    return RBPHD_OK;
}

int rbphd_get_counters(rbphd_navigator* nav, int64_t out4[4], int reset)
{
    if (!nav || !out4) return RBPHD_ERR_ARGUMENT;
    if (int r = set_device(nav)) return r;
    DeviceState st;
    if (int r = read_state(nav, &st)) return r;
    out4[0] = (int64_t)st.comps_in; out4[1] = (int64_t)st.comps_out;
    out4[2] = (int64_t)st.pairs; out4[3] = (int64_t)st.particle_frames;
    if (reset) {
        CK(cudaMemsetAsync(&nav->st->comps_in, 0, (4 + 64 + 16) * sizeof(unsigned long long), nav->stream));
        CK(cudaStreamSynchronize(nav->stream));
    }
    return RBPHD_OK;
}

int rbphd_launch_shape(const rbphd_navigator* nav, int64_t out5[5])
{
    if (!nav || !out5) return RBPHD_ERR_ARGUMENT;
    out5[0] = kBlock; out5[1] = nav->ctas_per_sm; out5[2] = (int64_t)nav->smem; out5[3] = nav->nslab;
    out5[4] = (int64_t)nav->lay.bytes;
    return RBPHD_OK;
}

void* rbphd_stream(rbphd_navigator* nav) { return nav ? (void*)nav->stream : nullptr; }

}  // extern "C"
