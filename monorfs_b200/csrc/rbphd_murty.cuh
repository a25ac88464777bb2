// rbphd_murty.cuh -- the slow lane of SetLogLikelihood (PHD:496-508): association blocks with more
// than five rows go through Murty's k-best enumeration (GC:241-272) on top of the Hungarian solver
// (GC:64-175), restated literally.  One thread runs it per block (blocks this large are rare), with
// its workspace in the CTA's scratch slab.  Included by rbphd_kernels.cu before rbphd_weight.cuh.
#pragma once

namespace rbphd {

constexpr int kMurtyN = 24;       // largest block (rows) the lane accepts; larger -> ST_OVER_BLOCK
constexpr int kMurtyNodes = 384;  // node pool of one enumeration; exhausted -> ST_OVER_MURTY
constexpr int kMurtyBig = 32;     // large blocks per particle whose value lists are kept for the stale-buffer rule

struct MurtyNode {
    int parent;                  // node this one was derived from (-1: root); eliminated set = chain to the root
    short elim_i, elim_k;        // the edge eliminated by this node (-1: none)
    short nf;                    // forced edges
    short isnull;                // Hungarian found no assignment
    double value;
    signed char fi[kMurtyN], fk[kMurtyN];
    signed char assign[kMurtyN];
};

struct MurtyWork {
    double profit[kMurtyN * kMurtyN];
    double reduced[kMurtyN * kMurtyN];
    double labelx[kMurtyN], labely[kMurtyN], slack[kMurtyN];
    int matchx[kMurtyN], matchy[kMurtyN], parent[kMurtyN];
    int visitx[kMurtyN], visity[kMurtyN];
    MurtyNode nodes[kMurtyNodes];
    int frontier[kMurtyNodes];
    int nnodes, nfront, best, have_best, overflow;
    double bigvals[kMurtyBig][200];
    int bigcnt[kMurtyBig], bighead[kMurtyBig];
    double tmpvals[200];
    double lgrad[200 * 6];       // gradient lane: pose-gradient sum of every assignment of the current block
    signed char lperm[200 * 5];  // ... and the assignments of a lexicographically enumerated block
};

// GC:64-175 on a dense n x n matrix whose undefined entries are -inf.  false = "no solution".
__device__ inline bool murty_hungarian(MurtyWork& w, const double* mat, int n, signed char* out)
{
    for (int i = 0; i < n; i++) {
        double folded = 0;   // FoldRows(Math.Max, 0)
        for (int k = 0; k < n; k++) folded = fmax(folded, mat[i * kMurtyN + k]);
        w.labelx[i] = folded;
        w.labely[i] = 0;
        w.matchx[i] = -1;
        w.matchy[i] = -1;
    }
    while (true) {
        int root = -1;
        for (int i = 0; i < n; i++) if (w.matchx[i] == -1) { root = i; break; }
        if (root == -1) break;
        for (int i = 0; i < n; i++) {
            w.parent[i] = root;
            w.slack[i] = w.labelx[root] + w.labely[i] - mat[root * kMurtyN + i];
            w.visitx[i] = 0;
            w.visity[i] = 0;
        }
        w.visitx[root] = 1;
        int iminslack = -1;
        bool found = false;
        while (!found) {
            iminslack = -1;
            double delta = INFINITY;
            for (int i = 0; i < n; i++)
                if (!w.visity[i] && w.slack[i] < delta) { iminslack = i; delta = w.slack[i]; }
            if (isinf(delta) && delta > 0) return false;
            for (int i = 0; i < n; i++) if (w.visitx[i]) w.labelx[i] -= delta;
            for (int i = 0; i < n; i++) { if (w.visity[i]) w.labely[i] += delta; else w.slack[i] -= delta; }
            w.visity[iminslack] = 1;
            if (w.matchy[iminslack] != -1) {
                int match = w.matchy[iminslack];
                w.visitx[match] = 1;
                for (int i = 0; i < n; i++)
                    if (!w.visity[i]) {
                        double mdelta = w.labelx[match] + w.labely[i] - mat[match * kMurtyN + i];
                        if (mdelta < w.slack[i]) { w.slack[i] = mdelta; w.parent[i] = match; }
                    }
            }
            else found = true;
        }
        int px, py, ty;
        for (py = iminslack, px = w.parent[py]; px != root; py = ty, px = w.parent[py]) {
            ty = w.matchx[px];
            w.matchx[px] = py;
            w.matchy[py] = px;
        }
        w.matchx[px] = py;
        w.matchy[py] = px;
    }
    for (int i = 0; i < n; i++) out[i] = (signed char)w.matchx[i];
    return true;
}

// GC:183-197
__device__ inline double murty_assignment_value(const MurtyWork& w, int n, const MurtyNode& nd)
{
    if (nd.isnull) return -INFINITY;
    double total = 0;
    for (int i = 0; i < n; i++) total += w.profit[i * kMurtyN + nd.assign[i]];
    return total;
}

// PriorityQueue.Add (GC:638-642) with a stable ascending order; Pop takes the back
__device__ inline void murty_push(MurtyWork& w, int node)
{
    double pr = w.nodes[node].value;
    int pos = w.nfront;
    while (pos > 0 && (w.nodes[w.frontier[pos - 1]].value - pr) > 0) pos--;
    for (int a = w.nfront; a > pos; a--) w.frontier[a] = w.frontier[a - 1];
    w.frontier[pos] = node;
    w.nfront++;
}

__device__ inline void murty_begin(MurtyWork& w, int n)
{
    w.nnodes = 1; w.nfront = 0; w.have_best = 0; w.best = -1; w.overflow = 0;
    MurtyNode& r = w.nodes[0];
    r.parent = -1; r.elim_i = r.elim_k = -1; r.nf = 0;
    r.isnull = murty_hungarian(w, w.profit, n, r.assign) ? 0 : 1;
    r.value = murty_assignment_value(w, n, r);
    murty_push(w, 0);
}

// children of the last yielded node (GC:469-509), each solved on its reduced profit (GC:206-234)
__device__ inline void murty_expand(MurtyWork& w, int n)
{
    const int b = w.best;
    if (w.nodes[b].isnull) return;
    signed char ri[kMurtyN], rk[kMurtyN];
    int nrem = 0;
    for (int i = 0; i < n; i++) {
        bool forced = false;
        for (int f = 0; f < w.nodes[b].nf; f++)
            if (w.nodes[b].fi[f] == i && w.nodes[b].fk[f] == w.nodes[b].assign[i]) forced = true;
        if (!forced) { ri[nrem] = (signed char)i; rk[nrem] = w.nodes[b].assign[i]; nrem++; }
    }
    for (int ci = 0; ci + 1 < nrem; ci++) {
        if (w.nnodes >= kMurtyNodes) { w.overflow = 1; return; }
        const int id = w.nnodes;
        MurtyNode& ch = w.nodes[id];
        ch.parent = b;
        ch.elim_i = ri[ci]; ch.elim_k = rk[ci];
        ch.nf = w.nodes[b].nf;
        for (int f = 0; f < ch.nf; f++) { ch.fi[f] = w.nodes[b].fi[f]; ch.fk[f] = w.nodes[b].fk[f]; }
        for (int k = 0; k < ci; k++) { ch.fi[ch.nf] = ri[k]; ch.fk[ch.nf] = rk[k]; ch.nf++; }
        // reduceprofit
        for (int a = 0; a < n * kMurtyN; a++) w.reduced[a] = w.profit[a];
        for (int f = 0; f < ch.nf; f++) {
            for (int k = 0; k < n; k++) w.reduced[ch.fi[f] * kMurtyN + k] = -INFINITY;
            for (int i = 0; i < n; i++) w.reduced[i * kMurtyN + ch.fk[f]] = -INFINITY;
        }
        for (int f = 0; f < ch.nf; f++) w.reduced[ch.fi[f] * kMurtyN + ch.fk[f]] = 1;
        for (int a = id; a >= 0; a = w.nodes[a].parent) {
            if (w.nodes[a].elim_i >= 0) w.reduced[w.nodes[a].elim_i * kMurtyN + w.nodes[a].elim_k] = -INFINITY;
            if (w.nodes[a].parent < 0) break;
        }
        ch.isnull = murty_hungarian(w, w.reduced, n, ch.assign) ? 0 : 1;
        if (!ch.isnull) {
            ch.value = murty_assignment_value(w, n, ch);
            w.nnodes++;
            murty_push(w, id);
        }
    }
}

// next assignment value in Murty's order; false when the enumeration is exhausted
__device__ inline bool murty_next(MurtyWork& w, int n, double* value)
{
    if (w.have_best) { murty_expand(w, n); w.have_best = 0; }
    if (w.nfront == 0) return false;
    w.best = w.frontier[--w.nfront];
    w.have_best = 1;
    *value = w.nodes[w.best].value;
    return true;
}

}  // namespace rbphd
