// rbphd_kernels.cuh -- declarations shared by the kernels (rbphd_kernels.cu) and the C ABI (rbphd_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "rbphd_block.cuh"

namespace rbphd {

// Per-particle map record, struct-of-arrays inside one slab of 13 * cap doubles:
// field 0 = weight, 1..3 = mean, 4..12 = covariance (row-major 3x3, NOT re-symmetrised: quirk A9.9)
constexpr int kFields = 13;

enum RunMode {
    MODE_FRAME = 0,          // predict + correct + prune (+ weight)
    MODE_STAGE_PREDICT = 1,  // PredictConditional only; dump predicted map
    MODE_STAGE_CORRECT = 2,  // CorrectConditional on the given (already predicted) map; dump corrected list
    MODE_STAGE_PRUNE = 3,    // PruneModel on the given map
    MODE_STAGE_WEIGHT = 4,   // WeightAlpha on (buffer cur = predicted, buffer 1-cur = corrected)
    MODE_STAGE_SETLL = 5     // SetLogLikelihood on the landmark list in buffer cur
};

enum StatusBits {
    ST_OVER_COMPONENTS = 1, ST_OVER_PAIRS = 2, ST_OVER_EDGES = 4, ST_OVER_JMAP = 8, ST_OVER_LL = 16,
    ST_OVER_BLOCK = 32, ST_OVER_MURTY = 64
};

struct DeviceState {
    int cur;         // which of the two map buffers holds the current maps
    int best;        // BestParticle
    int resampled;   // last SlamUpdate resampled
    int status;      // OR of StatusBits
    int depleted;
    int pad[3];
    unsigned long long comps_in, comps_out, pairs, particle_frames;   // work counters
    unsigned long long phase_cycles[64];
    unsigned long long dbg[16];            // diagnostic event counts (see capi.Handle.DEBUG_COUNTERS)   // SM cycles spent per phase of k_particle_update (thread 0, all CTAs)
};

struct FrameGrid {
    CellGrid g;
    int start[kGridMaxCells + 1];
};

// byte offsets inside one CTA's scratch slab

constexpr int kRecFields = 26;       // Kalman record of a gated component (rbphd_kernels.cu: comp_update)

struct ScratchLayout {
    int cap_pred, cap_pairs, cap_list, cap_sort, cap_top, cap_edges, cap_j, cap_ll, cap_nodes;
    size_t pm, pwt, pwmd, ppd, cact, bidx;
    size_t pkey, pt, pmean, pwgt, crec, cpn, hits4;
    size_t skey, sval, skey2, sval2;
    size_t tw, tm, tloc, rho;
    size_t edst, nstate, nowner, nflag, gitems;
    // weight stage
    size_t jidx, jm, jmp, jpd, vsum, erad, erad2, cnorm, crad, llkey, llval, llgrad, uf, bcnt, mslots;
    size_t bytes;
};

struct KParams {
    DevCfg cfg;
    ScratchLayout lay;
    int P;                 // particles to process
    int first;             // first particle (stage calls use 0)
    int M;                 // measurements this frame
    int cap;               // map capacity per particle
    int mode;
    int only_mapping;
    double* maps[2];       // two buffers of P * 13 * cap doubles
    int* counts[2];
    double* poses;         // P x 7
    double* weights;       // P
    double* alphas;        // P
    double* alpha_parts;   // P x 8 diagnostic parts of WeightAlpha
    const double* z;       // M x 3 (device)
    const FrameGrid* vgrid;   // camera-frame grid over back-projected measurements
    const int* vitems;
    const FrameGrid* zgrid;   // measurement-space grid over z
    const int* zitems;
    unsigned char* scratch;   // gridDim.x slabs
    DeviceState* st;
    double* dump;          // stage dumps: 13 * dump_cap doubles (SoA) + count in dump_count
    int* dump_count;
    int dump_cap;
    size_t smem_sort_cap;  // elements of the shared-memory sort buffer
    int ll_flags;          // LikelihoodFlags of a MODE_STAGE_SETLL call
    int holdout;           // MODE_FRAME: this particle skips the frame (its map is carried over unchanged); -1 = none
};

enum LikelihoodFlags { LL_QUASI = 1, LL_DUMP_MATRIX = 2, LL_GRADIENT = 4, LL_TEMPERED_SUM = 8 };   // KParams::ll_flags (MODE_STAGE_SETLL)

struct Reading6 { double v[6]; };
void launch_predict_pose(cudaStream_t s, const DevCfg& cfg, int P, double* poses, Reading6 reading, double dt,
                         const double* gauss, int perfect_still);
void launch_frame_prep(cudaStream_t s, const DevCfg& cfg, const double* z, int M, FrameGrid* vg, int* vitems,
                       FrameGrid* zg, int* zitems, double* pts);
void launch_particle_update(cudaStream_t s, const KParams& prm, int grid, size_t smem);
// force: 0 = SlamUpdate tail (PHD:343-358); 1 = same but always resample; 2 = ResampleParticles() alone (PHD:724-760);
// 3 = SlamUpdate tail up to the depletion decision (the wheel follows later with force = 2)
// cum: P + 1 doubles of scratch (prefix sums of the wheel)
void launch_normalize_resample(cudaStream_t s, const DevCfg& cfg, int P, double* weights, double u, int force,
                               int* ancestors, DeviceState* st, double* cum);
void launch_copy_particles(cudaStream_t s, int P, int cap, double* const maps[2], int* const counts[2],
                           double* poses, double* poses_tmp, const int* ancestors, DeviceState* st);
void launch_flip(cudaStream_t s, DeviceState* st);
constexpr int kMaxCommRanks = 64;
void launch_migration_plan(cudaStream_t s, const int* ganc, const int* gcounts, int total, int world, int rank,
                           int* local_src, long long* rec_off, int* send_idx, long long* send_off, long long* hdr);
void launch_pack_records(cudaStream_t s, int cap, const double* maps, const int* counts, const double* poses,
                         const int* send_idx, const long long* send_off, int count, double* sendbuf);
void launch_unpack_records(cudaStream_t s, int P, int cap, const double* src_maps, const int* src_counts,
                           double* dst_maps, int* dst_counts, const double* poses, double* poses_tmp,
                           const int* local_src, const long long* rec_off, const double* recvbuf);
void launch_copy_doubles(cudaStream_t s, size_t n, double* dst, const double* src);
void launch_fill_doubles(cudaStream_t s, size_t n, double* dst, double v);
size_t particle_update_smem(int max_measurements, size_t* sort_cap);
size_t murty_workspace_bytes();
int particle_update_max_ctas_per_sm(size_t smem);

}  // namespace rbphd
