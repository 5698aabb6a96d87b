// pedigraph.cpp — Linux `pedigraph` sampler: the host-side chain logic of the reference's
// exec/pedigraph (a macOS binary shipped without source), restated from its -O0 disassembly,
// with every likelihood evaluation delegated to the engine behind the C ABI of
// include/pedfac_cuda.h. Same CLI (-d DIR -n N -r SEEDFILE [-c J] [-s] [-w] [-l]) and the same
// file contract (prior.txt, geno.txt, optional marriage.txt -> out/ped.txt, out/ll.txt,
// out/historyNode.csv, out/historyEdge.csv, out/geno.txt), so R/runPedFac.R drives it unmodified
// (R/runPedFac.R:99-108).
//
// Each routine cites the reference routine it follows as pedigraph@<VM address> <symbol>.
// Differences from the reference, all deliberate and listed in DESIGN.md §7:
//   * swap proposals of kind 4 (two spouses) are not generated: the reference's enumeration of them
//     can never fire (_EvalMoves@0x100016136-0x100016292 looks up the parents of the focal
//     individual itself, which is pruned at that point); swaps that touch a selfing marriage are
//     skipped (the reference reads an unset message there);
//   * -c 1 (throttle) runs as -c 0, -l (loopy BP) exits where the reference does (both as observed on the
//     reference binary: DESIGN.md §7);
//   * the sweep that ends _PickAndUpdate is merged with the one that follows the next
//     _PruneIndiv (nothing reads messages in between), so a Gibbs step costs one sweep;
//   * mnode kid arrays grow (the reference overflows a malloc(8) block);
//   * a chosen parent (or marriage) that the trimming of the previous parents deletes in the same step is
//     replaced by a new unobserved parent (the reference links the kid to the freed slot).
//
// The engine entry points are reached through PFN(name); the default binds libpedfac_cuda
// (pf_*). The test infrastructure builds the same file against its CPU checker by defining
// PF_BACKEND_HEADER / PF_BACKEND_PREFIX on the compiler command line.
#ifndef PF_BACKEND_HEADER
#define PF_BACKEND_HEADER "../../../include/pedfac_cuda.h"
#define PF_BACKEND_PREFIX pf_
#endif
#include PF_BACKEND_HEADER

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <sys/stat.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <ucontext.h>
#include <unistd.h>
#include <vector>
#include <cstdarg>
#include <thread>

#define PF_CAT2(a, b) a##b
#define PF_CAT(a, b) PF_CAT2(a, b)
#define PFN(name) PF_CAT(PF_BACKEND_PREFIX, name)
#ifndef PF_ENGINE_T
#define PF_ENGINE_T PF_CAT(PF_BACKEND_PREFIX, engine)
#endif

// ---------------------------------------------------------------------------------------------
// ranlib subset (netlib ranlib.c, Brown & Lovato; L'Ecuyer-Cote combined generator). Only
// generator 1 is ever current, so setall() reduces to loading the two seeds
// (_setall@0x10002a2d8, _ignlgi@0x100029fa3, _ranf@0x10002b838, _genunf@0x10002e948, _genbet@0x10002accf).
// stdout of the chains other than chain 0 (PEDFAC_CHAINS) is dropped: one program, one progress log
static thread_local bool g_quiet = false;   // per scheduler thread (PEDFAC_THREADS)
static int out_printf(const char *fmt, ...)
{
    if (g_quiet) return 0;
    va_list ap;
    va_start(ap, fmt);
    const int n = vprintf(fmt, ap);
    va_end(ap);
    return n;
}
#define printf out_printf

namespace rng {
static thread_local long s1 = 1234567890L, s2 = 123456789L;   // one stream per scheduler thread; fibers swap theirs in and out
static const long m1 = 2147483563L, m2 = 2147483399L, a1 = 40014L, a2 = 40692L;

static void setall(long iseed1, long iseed2) { s1 = iseed1; s2 = iseed2; }

static long ignlgi()
{
    long k = s1 / 53668L;
    s1 = a1 * (s1 - k * 53668L) - k * 12211L;
    if (s1 < 0) s1 += m1;
    k = s2 / 52774L;
    s2 = a2 * (s2 - k * 52774L) - k * 3791L;
    if (s2 < 0) s2 += m2;
    long z = s1 - s2;
    if (z < 1) z += (m1 - 1);
    return z;
}
// _ranf@0x10002b83d-0x10002b8b0: redrawn while the float equals 0.0 or 1.0 (E.C. Anderson's change)
static float ranf()
{
    float r;
    do { r = (float)((double)ignlgi() * 4.656613057E-10); } while (r == 1.0f || r == 0.0f);
    return r;
}
static float genunf(float low, float high) { return low + (high - low) * ranf(); }

static float genbet(float aa, float bb)
{
    const double expmax = 89.0;
    const float infnty = 1.0E38f;
    static thread_local float olda = -1.0f, oldb = -1.0f;
    static thread_local float result, a, alpha, b, beta, delta, gamma, k1, k2, r, s, t, u1, u2, v, w, y, z;
    const bool qsame = olda == aa && oldb == bb;
    if (!qsame) {
        if (aa <= 0.0 || bb <= 0.0) {
            fputs(" AA or BB <= 0 in GENBET - Abort!", stderr);
            fprintf(stderr, " AA: %16.6E BB %16.6E\n", aa, bb);
            exit(1);
        }
        olda = aa; oldb = bb;
    }
    if (std::min(aa, bb) > 1.0) {            // algorithm BB
        if (!qsame) {
            a = std::min(aa, bb); b = std::max(aa, bb);
            alpha = a + b;
            beta = (float)sqrt((alpha - 2.0) / (2.0 * a * b - alpha));
            gamma = (float)(a + 1.0 / beta);
        }
        for (;;) {
            u1 = ranf(); u2 = ranf();
            v = (float)(beta * log(u1 / (1.0 - u1)));
            w = v > expmax ? infnty : (float)(a * exp(v));
            z = (float)(pow(u1, 2.0) * u2);
            r = (float)(gamma * v - 1.3862944);
            s = a + r - w;
            if (s + 2.609438 >= 5.0 * z) break;
            t = (float)log(z);
            if (s > t) break;
            if (!(r + alpha * log(alpha / (b + w)) < t)) break;
        }
        result = aa == a ? w / (b + w) : b / (b + w);
        return result;
    }
    // algorithm BC
    if (!qsame) {
        a = std::max(aa, bb); b = std::min(aa, bb);
        alpha = a + b;
        beta = (float)(1.0 / b);
        delta = (float)(1.0 + a - b);
        k1 = (float)(delta * (1.38889E-2 + 4.16667E-2 * b) / (a * beta - 0.777778));
        k2 = (float)(0.25 + (0.5 + 0.25 / delta) * b);
    }
    for (;;) {
        u1 = ranf(); u2 = ranf();
        if (u1 < 0.5) {
            y = u1 * u2; z = u1 * y;
            if (0.25 * u2 + z - y >= k1) continue;
        } else {
            z = (float)(pow(u1, 2.0) * u2);
            if (z <= 0.25) {
                v = (float)(beta * log(u1 / (1.0 - u1)));
                w = v > expmax ? infnty : (float)(a * exp(v));
                break;
            }
            if (z >= k2) continue;
        }
        v = (float)(beta * log(u1 / (1.0 - u1)));
        w = v > expmax ? infnty : (float)(a * exp(v));
        if (alpha * (log(alpha / (b + w)) + v) - 1.3862944 < log(z)) continue;
        break;
    }
    result = a == aa ? w / (b + w) : b / (b + w);
    return result;
}
} // namespace rng

// ---------------------------------------------------------------------------------------------
// Pedigree state (`ped`, `indiv`, `mnode`, `choice`)

struct Indiv {
    int sex = 0, sexWasUnknown = 0, userID = 0, gen = 0;
    double genTime = 0;
    bool observed = false, founder = false, reductive = false, queued = false;
    int unobsDegree = 0;
    std::vector<int> M, edge;       // attached marriages and the edge index in each (0 pa, 1 ma, >=2 kid)
};
struct Marr {
    int pa = -1, ma = -1;
    std::vector<int> kids;
    bool selfing = false, mark = false;
    int nEdges() const { return 2 + (int)kids.size(); }
};
struct Choice { int mnode, pa, ma, selfing, kind; int swapKind = 0, marrZ = -1, swapZ = -1; };   // kind 3 = swap (aux: kind, marr_z, z | -1)

struct Opts {
    std::string dir, seedfile = "rand.tem";
    int nIter = 1, cyclic = 0;
    bool selfing = false, diswap = false, loopy = false;
};

struct Ped {
    int S = 0, nGen = 0, nIndivDecl = 0, nObs = 0, nMarDecl = 0, maxID = 0, maxUnobsLayer = 2;
    double maxAge = 1.0, minAge = 1.0;
    std::vector<double> afreq, eps, obsFrac;
    std::vector<int> nObsPerGen, nIndivPerGen;
    int nIndivUsed = 0, nMarrUsed = 0, capIndiv = 0, capMarr = 0;
    std::vector<Indiv> indiv;
    std::vector<Marr> marr;
    std::vector<char> indivActive, marrActive;
    std::vector<int> cc, ccDirty;
    int nCC = 0, nDirty = 0;
    std::vector<int> marked;                         // marriage slots whose `mark` is set (see the prong block of EvalMoves)
    std::vector<std::vector<int>> ccMem;             // the active individuals of every component label (kept from CheckPed on): the
                                                     // reference scans all individuals wherever this is walked
    std::vector<double> LLcc;
    double LLtot = 0;
    std::vector<int> freeIndiv, freeMarr;
    std::vector<std::vector<int>> parentCand;   // ped->ff0 (<= 15 per observed kid)
    std::vector<int> observedSorted, unobsQueue;
    int iter = 0;
    // choice list
    std::vector<Choice> ch;
    std::vector<double> chll;
    int prev[3] = {-1, -1, -1}, neu[4] = {-1, -1, -1, 0}, picked = -1, t = 0;
    int keep[3] = {-1, -1, -1};                                // chosen pa, ma, marriage: not trimmed (see PickAndUpdate)
    // candidate lists by sex (0 unknown, 1 male, 2 female)
    std::vector<std::pair<int, double>> cand[3];
    PF_ENGINE_T *eng = nullptr;
    int chain = 0;                                   // row block of this chain inside its engine (PEDFAC_CHAINS)
    int chain_id = 0;                                // its number among all chains of the run (seeds, output directory)
    FILE *fNode = nullptr, *fEdge = nullptr, *fPed = nullptr, *fGeno = nullptr, *fLL = nullptr;
    long n_proposals = 0, n_sweeps = 0, n_steps = 0, n_swaps = 0, n_swaps_picked = 0;
    bool underflow_warned = false; long underflow_steps = 0;
    bool check = false;                              // PEDFAC_CHECK: verify every picked move against the next sweep
    std::vector<double> chll_data;
    double t_prune = 0, t_moves = 0, t_pick = 0;     // seconds in _PruneIndiv, _EvalMoves (engine calls included), _AddObsFracPrior + _PickAndUpdate
    double t_sweep = 0, t_eval = 0, t_view = 0;      // seconds inside the engine calls / building views (PEDFAC_STATS)
};

static double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void die(const char *msg)
{
    fprintf(stderr, "%s: %s\n", msg, PFN(last_error)());
    exit(-1);
}

// history rows (graph-visualiser event trace)
static int group_of(const Indiv &x) { return x.sex * 10 + (x.observed ? 1 : 0); }
static void hist_node(Ped &p, int action, int i)
{
    const Indiv &x = p.indiv[i];
    fprintf(p.fNode, "%i,%i,%i,%i,%i,%i\n", p.t, action, x.userID, x.userID, x.gen << 1, group_of(x));
}
static void hist_mnode(Ped &p, int action, int m)
{
    const Marr &mm = p.marr[m];
    const int a = p.indiv[mm.pa].userID, b = p.indiv[mm.ma].userID;
    fprintf(p.fNode, "%i,%i,%i-%i,%i-%i,%i,m-nodes\n", p.t, action, a, b, a, b, (p.indiv[mm.pa].gen << 1) - 1);
}
static void hist_edge(Ped &p, int action, int m, int to)
{
    const Marr &mm = p.marr[m];
    const int a = p.indiv[mm.pa].userID, b = p.indiv[mm.ma].userID;
    fprintf(p.fEdge, "%i,%i,%i-%i-%i,%i-%i,%i\n", p.t, action, a, b, p.indiv[to].userID, a, b, p.indiv[to].userID);
}

static void mark_dirty(Ped &p, int label)
{
    if (p.ccDirty[label] == 0) p.nDirty++;
    p.ccDirty[label] = 1;
}

// _Transverse_fg@0x100006a08 + _DFS@0x100004d58: label everything connected to `start`
static void label_component(Ped &p, int start, int label)
{
    std::vector<int> stack{start};
    std::vector<int> &mem = p.ccMem[label];
    p.cc[start] = label; mem.push_back(start);
    while (!stack.empty()) {
        const int i = stack.back(); stack.pop_back();
        for (int m : p.indiv[i].M) {
            const Marr &mm = p.marr[m];
            auto visit = [&](int y) { if (p.cc[y] != label) { p.cc[y] = label; mem.push_back(y); stack.push_back(y); } };
            visit(mm.pa);
            if (!mm.selfing) visit(mm.ma);
            for (int k : mm.kids) visit(k);
        }
    }
}

// _UpdateConComp@0x100013128: the component that contained `id` may have split (or vanished)
static void UpdateConComp(Ped &p, int id)
{
    const int old = p.cc[id];
    std::vector<int> mem;
    mem.swap(p.ccMem[old]);
    for (int i : mem) p.cc[i] = -5;
    if (mem.empty()) {
        if (p.ccDirty[old] == 1) p.nDirty--;
        p.ccDirty[old] = 0;
        if (p.nCC - 1 != old) {
            p.ccMem[old].swap(p.ccMem[p.nCC - 1]);
            for (int i : p.ccMem[old]) p.cc[i] = old;
            if (p.ccDirty[p.nCC - 1] == 1) { p.ccDirty[old] = 1; p.ccDirty[p.nCC - 1] = 0; }
            p.LLcc[old] = p.LLcc[p.nCC - 1];
        }
        p.nCC--;
        return;
    }
    std::sort(mem.begin(), mem.end());                  // the reference's scan order: by slot
    int found = 0, label = old;
    for (int i : mem) {
        if (p.cc[i] != -5) continue;
        label_component(p, i, label);
        found++;
        if (found == 1) {
            mark_dirty(p, label);
            label = p.nCC;
        } else {
            p.ccDirty[p.nCC] = 1; p.nDirty++;
            p.nCC++;
            label = p.nCC;
        }
    }
}

// _AssignConComp@0x100013548: merge two labels (smaller wins; last label fills the hole)
static void AssignConComp(Ped &p, int a, int b)
{
    if (a == b) return;
    const int hi = std::max(a, b), lo = std::min(a, b);
    if (p.ccDirty[hi] == 1) {
        if (p.ccDirty[lo] == 1) p.nDirty--;
        p.ccDirty[lo] = 1; p.ccDirty[hi] = 0;
    }
    for (int i : p.ccMem[hi]) { p.cc[i] = lo; p.ccMem[lo].push_back(i); }
    p.ccMem[hi].clear();
    if (p.nCC - 1 != hi) {
        p.ccMem[hi].swap(p.ccMem[p.nCC - 1]);
        for (int i : p.ccMem[hi]) p.cc[i] = hi;
        if (p.ccDirty[p.nCC - 1] == 1) { p.ccDirty[hi] = 1; p.ccDirty[p.nCC - 1] = 0; }
        p.LLcc[hi] = p.LLcc[p.nCC - 1];
    }
    p.nCC--;
}

// _DissociateParentMnode@0x100013804: drop marriage m from x's attachment list (swap with last)
static void DissociateParentMnode(Ped &p, int x, int m)
{
    Indiv &ix = p.indiv[x];
    const int n = (int)ix.M.size();
    for (int k = 0; k < n - 1; k++)
        if (ix.M[k] == m) { ix.M[k] = ix.M[n - 1]; ix.edge[k] = ix.edge[n - 1]; break; }
    ix.M.pop_back(); ix.edge.pop_back();
}
// _DissociateSibMnode@0x100013948: drop kid x from marriage m (swap with last, renumber the moved sib)
static void DissociateSibMnode(Ped &p, int x, int m)
{
    Marr &mm = p.marr[m];
    const int nk = (int)mm.kids.size();
    for (int k = 0; k < nk - 1; k++) {
        if (mm.kids[k] != x) continue;
        mm.kids[k] = mm.kids[nk - 1];
        Indiv &sib = p.indiv[mm.kids[k]];
        for (size_t j = 0; j < sib.M.size(); j++)
            if (sib.edge[j] > 1) sib.edge[j] = k + 2;
        break;
    }
    mm.kids.pop_back();
}

// _UpdateUnobsDegree@0x100007550
static void UpdateUnobsDegree(Ped &p, int x)
{
    if (x < 0) return;
    Indiv &ix = p.indiv[x];
    if (ix.observed) { ix.unobsDegree = 0; return; }
    int best = p.maxUnobsLayer + 1;
    for (size_t k = 0; k < ix.M.size(); k++) {
        const Marr &mm = p.marr[ix.M[k]];
        const int e = ix.edge[k];
        auto look = [&](int y) -> bool {
            if (p.indiv[y].observed) { ix.unobsDegree = 1; return true; }
            best = std::min(best, p.indiv[y].unobsDegree + 1);
            return false;
        };
        if (e != 0 && look(mm.pa)) return;
        if (e != 1 && look(mm.ma)) return;
        for (size_t c = 0; c < mm.kids.size(); c++)
            if ((int)c + 2 != e && look(mm.kids[c])) return;
        if (best == 1) { ix.unobsDegree = 1; return; }
    }
    ix.unobsDegree = best;
}

// _DFS_i2m_obs@0x100007070: does individual i contribute information to marriage `via`
// (observed, a parent elsewhere, or a kid elsewhere while `via` has further kids)?
static bool informs(Ped &p, int i, int via)
{
    const Indiv &x = p.indiv[i];
    if (x.observed) return true;
    for (size_t k = 0; k < x.M.size(); k++) {
        if (x.M[k] == via) continue;
        if (x.edge[k] <= 1) return true;
        if (p.marr[via].nEdges() > 3) return true;
    }
    return false;
}
// _isReductiveParentRecursive@0x100007170: nothing informative anywhere up y's ancestry
static bool reductive_up(Ped &p, int y)
{
    const Indiv &iy = p.indiv[y];
    for (size_t k = 0; k < iy.M.size(); k++) {
        if (iy.edge[k] <= 1) continue;
        const Marr &mm = p.marr[iy.M[k]];
        if (informs(p, mm.pa, iy.M[k]) || !reductive_up(p, mm.pa)) return false;
        if (informs(p, mm.ma, iy.M[k]) || !reductive_up(p, mm.ma)) return false;
    }
    return true;
}
// _isReductiveParent@0x100007340: an unobserved individual that is nobody's parent and whose
// ancestry holds nothing informative can be dropped without changing any likelihood
static bool isReductiveParent(Ped &p, int x)
{
    if (x < 0) return false;
    const Indiv &ix = p.indiv[x];
    if (ix.observed) return false;
    for (size_t k = 0; k < ix.M.size(); k++) {
        if (ix.edge[k] <= 1) return false;
        const Marr &mm = p.marr[ix.M[k]];
        if (informs(p, mm.pa, ix.M[k]) || !reductive_up(p, mm.pa)) return false;
        if (informs(p, mm.ma, ix.M[k]) || !reductive_up(p, mm.ma)) return false;
    }
    return true;
}

// _TrimPed@0x100013b30: remove an unobserved individual that no longer links anything
static void TrimPed(Ped &p, int x)
{
    Indiv &ix = p.indiv[x];
    if (ix.observed) return;
    if (x == p.keep[0] || x == p.keep[1]) return;
    if (ix.M.size() > 1) return;
    if (ix.M.size() == 1 && ix.edge[0] < 2) return;
    p.indivActive[x] = 0;
    if (p.cc[x] >= 0) {
        std::vector<int> &mem = p.ccMem[p.cc[x]];
        for (size_t k = 0; k < mem.size(); k++) if (mem[k] == x) { mem[k] = mem.back(); mem.pop_back(); break; }
    }
    ix.queued = false;
    p.freeIndiv.push_back(x);
    p.nIndivPerGen[ix.gen]--;
    hist_node(p, 0, x);
    if (ix.M.empty()) return;
    const int m = ix.M[0];
    hist_edge(p, 0, m, x);
    Marr &mm = p.marr[m];
    if (mm.nEdges() == 3 && m != p.keep[2]) {
        p.marrActive[m] = 0;
        p.freeMarr.push_back(m);
        hist_edge(p, 0, m, mm.pa);
        if (!mm.selfing) hist_edge(p, 0, m, mm.ma);
        hist_mnode(p, 0, m);
        const int pa = mm.pa, ma = mm.ma;
        const bool self = mm.selfing;
        DissociateParentMnode(p, pa, m);
        if (!self) DissociateParentMnode(p, ma, m);
        mm.kids.clear();
        TrimPed(p, pa);
        if (!self) TrimPed(p, ma);
    } else {
        DissociateSibMnode(p, x, m);
    }
    ix.M.clear(); ix.edge.clear();
}

// _RevertSex@0x100014910
static void RevertSex(Ped &p, int x)
{
    Indiv &ix = p.indiv[x];
    if (ix.sexWasUnknown != 1) return;
    for (int e : ix.edge) if (e < 2) return;
    ix.sex = 0;
}

// ---------------------------------------------------------------------------------------------
// _Propagating_msg@0x10001ba10: build the view of the dirty components and sweep it

struct ViewBuf {
    std::vector<int> indiv, icc, marr, mpa, mma, kb{0}, kids, labels;
    std::vector<uint8_t> founder, selfing;
    pf_graph_view v;
};

// the worklists of _TurnOnSpotlight@0x10001bb90 for the components with want[label] != 0
static void build_view(Ped &p, const std::vector<char> &want, ViewBuf &b)
{
    std::vector<int> local(p.nCC, -1);
    for (int c = 0; c < p.nCC; c++)
        if (want[c]) { local[c] = (int)b.labels.size(); b.labels.push_back(c); }
    // individuals and marriages in slot order, as the reference's scans list them
    for (int c : b.labels) b.indiv.insert(b.indiv.end(), p.ccMem[c].begin(), p.ccMem[c].end());
    std::sort(b.indiv.begin(), b.indiv.end());
    for (int i : b.indiv) {
        b.icc.push_back(local[p.cc[i]]); b.founder.push_back(p.indiv[i].founder ? 1 : 0);
        for (int m : p.indiv[i].M) if (p.marr[m].pa == i && p.marrActive[m]) b.marr.push_back(m);
    }
    std::sort(b.marr.begin(), b.marr.end());
    b.marr.erase(std::unique(b.marr.begin(), b.marr.end()), b.marr.end());
    const std::vector<int> ms = b.marr;
    b.marr.clear();
    for (int m : ms) {
        const Marr &mm = p.marr[m];
        b.marr.push_back(m); b.mpa.push_back(mm.pa); b.mma.push_back(mm.ma); b.selfing.push_back(mm.selfing ? 1 : 0);
        b.kids.insert(b.kids.end(), mm.kids.begin(), mm.kids.end());
        b.kb.push_back((int)b.kids.size());
    }
    pf_graph_view &v = b.v;
    v.n_indiv = (int)b.indiv.size(); v.indiv = b.indiv.data(); v.indiv_cc = b.icc.data(); v.indiv_founder = b.founder.data();
    v.n_marr = (int)b.marr.size(); v.marr = b.marr.data(); v.marr_pa = b.mpa.data(); v.marr_ma = b.mma.data();
    v.marr_selfing = b.selfing.data(); v.marr_kid_begin = b.kb.data(); v.marr_kids = b.kids.data();
    v.n_cc = (int)b.labels.size();
}

// the dirty components as a view (all of them when everything is dirty)
static void dirty_view(Ped &p, ViewBuf &b)
{
    const bool all = p.nDirty == p.nCC;
    std::vector<char> want(p.nCC, 0);
    for (int c = 0; c < p.nCC; c++) want[c] = all || p.ccDirty[c];
    build_view(p, want, b);
}

static const double kUnderflowMeanPerLocus = -354.0;          // half of log(DBL_MIN)

// store the component log-likelihoods a sweep returned and recompute the total
// (_UpdateProb@0x10001dd0c-0x10001dd65)
static void store_sweep(Ped &p, const ViewBuf &b, const std::vector<double> &ll)
{
    p.n_sweeps++;
    for (size_t j = 0; j < b.labels.size(); j++) {
        p.LLcc[b.labels[j]] = ll[j];
        // Messages are unnormalised products (as in the reference): once a component's partition value
        // at a locus approaches the smallest normal double (log = -708), its messages lose precision and
        // then vanish. Only the sum over loci comes back, and the loci spread widely around their mean
        // (a sibship of 700 unrelated individuals: mean -449, minimum -737), hence half the range.
        if (ll[j] < kUnderflowMeanPerLocus * p.S) {
            p.underflow_steps++;
            if (!p.underflow_warned) {
                p.underflow_warned = true;
                fprintf(stderr, "warning: a component's log-likelihood is %.1f over %d loci (%.0f per locus): its unnormalised messages can leave "
                                "the double range at some loci (the limit of the reference's arithmetic as well); scores involving it are unreliable\n",
                        ll[j], p.S, ll[j] / p.S);
            }
        }
    }
    p.LLtot = 0;
    for (int c = 0; c < p.nCC; c++) p.LLtot = p.LLcc[c] + p.LLtot;
    std::fill(p.ccDirty.begin(), p.ccDirty.begin() + std::max(p.nCC, 0), 0);
    p.nDirty = 0;
}

// ---------------------------------------------------------------------------------------------
// Several chains in one process (PEDFAC_CHAINS=C; BASELINE.json's last configuration: independent
// chains batched per GPU). Every chain is the unchanged single-chain sampler running on its own fiber
// (ucontext) with its own Ped, RNG stream (seeds + c) and output directory; the only thing they share
// is the engine, whose n_chains row blocks they own one each. The likelihood calls below do not go to
// the engine directly: a fiber posts its request and yields, and when every live chain has posted,
// the scheduler submits the requests of a kind as ONE pf_sweep_batch / pf_eval_batch — one launch
// serves all chains — and hands the results back. A chain therefore computes exactly what it would
// compute alone: chain c's trace is the single-chain trace for the seeds (a + c, b + c).
namespace mc {
enum Kind { NONE = 0, SWEEP, EVAL, SWEEP_EVAL, SWAPS };
struct Req { Kind kind = NONE; const pf_graph_view *view = nullptr; double *ll = nullptr; int kid = -1, n = 0;
             const pf_proposal *props = nullptr; double *lsum = nullptr; const pf_swap *swaps = nullptr; int rc = 0; };
struct Fiber { ucontext_t ctx; void *stack = nullptr; Req req; bool done = false; long s1 = 0, s2 = 0; int chain = 0; };
// one scheduler per thread (PEDFAC_THREADS): its fibers are the chains base .. base + fibers.size() - 1
static thread_local bool active = false;
static thread_local int current = -1, base = 0;
static thread_local std::vector<Fiber> fibers;
static thread_local ucontext_t sched_ctx;

static int post(const Req &r)
{
    Fiber &f = fibers[current];
    f.req = r;
    swapcontext(&f.ctx, &sched_ctx);             // back in the scheduler; resumed when the result is in
    return f.req.rc;
}
} // namespace mc

static int eng_sweep(Ped &p, const pf_graph_view *v, double *ll)
{
    if (!mc::active) return PFN(sweep)(p.eng, p.chain, v, ll);
    mc::Req r; r.kind = mc::SWEEP; r.view = v; r.ll = ll;
    return mc::post(r);
}
static int eng_eval(Ped &p, int kid, int n, const pf_proposal *props, double *lsum)
{
    if (!mc::active) return PFN(eval)(p.eng, p.chain, kid, n, props, lsum);
    mc::Req r; r.kind = mc::EVAL; r.kid = kid; r.n = n; r.props = props; r.lsum = lsum;
    return mc::post(r);
}
static int eng_sweep_eval(Ped &p, const pf_graph_view *v, double *ll, int kid, int n, const pf_proposal *props, double *lsum)
{
    if (!mc::active) return PFN(sweep_eval)(p.eng, p.chain, v, ll, kid, n, props, lsum);
    mc::Req r; r.kind = mc::SWEEP_EVAL; r.view = v; r.ll = ll; r.kid = kid; r.n = n; r.props = props; r.lsum = lsum;
    return mc::post(r);
}
static int eng_eval_swaps(Ped &p, const pf_graph_view *v, int n, const pf_swap *sw, double *lsum)
{
    if (!mc::active) return PFN(eval_swaps)(p.eng, p.chain, v, n, sw, lsum);
    mc::Req r; r.kind = mc::SWAPS; r.view = v; r.n = n; r.swaps = sw; r.lsum = lsum;
    return mc::post(r);
}

static void Propagating_msg(Ped &p)
{
    if (p.nDirty == 0) return;     // the reference re-propagates everything; nothing would change
    const double t0 = now_s();
    ViewBuf b;
    dirty_view(p, b);
    std::vector<double> ll(b.labels.size() + 1);
    const double t1 = now_s();
    if (eng_sweep(p, &b.v, ll.data()) != PF_OK) die("sweep failed");
    p.t_view += t1 - t0; p.t_sweep += now_s() - t1;
    store_sweep(p, b, ll);
}

// ---------------------------------------------------------------------------------------------
// proposals

struct Batch { std::vector<pf_proposal> props; std::vector<int> slot; };   // slot = index into p.ch

static double compose(Ped &p, int kid, const pf_proposal &q, double lsum)
{
    const int ck = p.cc[kid];
    if (q.kind == PF_PROP_PRONG) {                             // _CalculateLikeByProngs@0x10001fc55-0x10001fce1
        const int cp = p.cc[p.marr[q.marr].pa];
        return pf_compose_ll(p.LLtot, lsum, p.LLcc[ck], cp != ck, cp != ck ? p.LLcc[cp] : 0.0, 0, 0.0);
    }
    const int cp = q.pa < 0 ? -1 : p.cc[q.pa];
    if (q.kind == PF_PROP_SELFING)                             // _CalculateLikeByStubsSelfing@0x10001f929-0x10001f9f7
        return pf_compose_ll(p.LLtot, lsum, p.LLcc[ck], cp != -1 && cp != ck, cp >= 0 ? p.LLcc[cp] : 0.0, 0, 0.0);
    const int cm = q.ma < 0 ? -1 : p.cc[q.ma];                 // _CalculateLikeByStubs@0x10001f547-0x10001f6ae
    const int use_pa = cp != -1 && cp != ck, use_ma = cm != -1 && cm != cp && cm != ck;
    return pf_compose_ll(p.LLtot, lsum, p.LLcc[ck], use_pa, use_pa ? p.LLcc[cp] : 0.0, use_ma, use_ma ? p.LLcc[cm] : 0.0);
}

static void flush(Ped &p, int kid, Batch &b)
{
    if (b.props.empty()) return;
    std::vector<double> lsum(b.props.size());
    const double t0 = now_s();
    if (eng_eval(p, kid, (int)b.props.size(), b.props.data(), lsum.data()) != PF_OK) die("eval failed");
    p.t_eval += now_s() - t0;
    p.n_proposals += (long)b.props.size();
    for (size_t k = 0; k < b.props.size(); k++) p.chll[b.slot[k]] = compose(p, kid, b.props[k], lsum[k]);
    b.props.clear(); b.slot.clear();
}

// The pending sweep (_Propagating_msg after _PruneIndiv) and the first batch of proposals of the
// same Gibbs step in one engine submission: the proposal list depends on the graph only.
struct SwapPlan;
static int submit_step(Ped &p, const pf_graph_view *v, double *ll, int kid, Batch &b, double *lsum, SwapPlan *sp);
static void sweep_and_flush(Ped &p, int kid, Batch &b, SwapPlan *sp)
{
    if (p.nDirty == 0) { flush(p, kid, b); return; }
    const double t0 = now_s();
    ViewBuf vb;
    dirty_view(p, vb);
    std::vector<double> ll(vb.labels.size() + 1), lsum(b.props.size() + 1);
    const double t1 = now_s();
    if (submit_step(p, &vb.v, ll.data(), kid, b, lsum.data(), sp) != PF_OK) die("sweep_eval failed");
    p.t_view += t1 - t0; p.t_sweep += now_s() - t1;
    store_sweep(p, vb, ll);
    p.n_proposals += (long)b.props.size();
    for (size_t k = 0; k < b.props.size(); k++) p.chll[b.slot[k]] = compose(p, kid, b.props[k], lsum[k]);
    b.props.clear(); b.slot.clear();
}

static int add_choice(Ped &p, Batch &b, int kind, int pa, int ma, int marr)
{
    pf_proposal q{kind, pa, ma, marr};
    static const bool dbg = getenv("PEDFAC_DEBUG_CHOICES") != nullptr;
    if (dbg) fprintf(stderr, "choice kind %d pa %d ma %d marr %d\n", kind, pa, ma, marr);
    Choice c;
    if (kind == PF_PROP_PRONG) c = Choice{marr, p.marr[marr].pa, p.marr[marr].ma, p.marr[marr].selfing ? 1 : 0, 0};
    else if (kind == PF_PROP_SELFING) c = Choice{-1, pa, pa, 1, 0};
    else c = Choice{-1, pa, ma, 0, 0};
    p.ch.push_back(c); p.chll.push_back(0.0);
    b.props.push_back(q); b.slot.push_back((int)p.ch.size() - 1);
    return (int)p.ch.size() - 1;
}

// _ProposeSingletMove@0x1000149d0 (enumeration part; the likelihood is evaluated in one batch)
static void ProposeSingletMove(Ped &p, const Opts &o, Batch &b, int c, int kid, std::vector<std::pair<int, int>> cand_slot[3])
{
    if (!p.indivActive[c]) return;
    const double gap = p.indiv[c].genTime - p.indiv[kid].genTime;
    if (gap > p.maxAge || p.minAge > gap) return;
    const int sex = p.indiv[c].sex;
    if (p.cc[c] == p.cc[kid]) return;          // -c 0: would close a loop (throttle modes unsupported)
    int slot;
    if (sex == 1) slot = add_choice(p, b, PF_PROP_TRIO, c, -1, -1);
    else slot = add_choice(p, b, PF_PROP_TRIO, -1, c, -1);
    cand_slot[sex == 1 ? 1 : sex == 2 ? 2 : 0].push_back({c, slot});
    if (o.selfing) add_choice(p, b, PF_PROP_SELFING, c, -1, -1);
}

// _checkMatchingCC@0x1000241b0 under -c 0
static bool checkMatchingCC(Ped &p, int pa, int ma, int kid)
{
    const int ck = p.cc[kid];
    if (!(p.cc[pa] == p.cc[ma] || ck == p.cc[pa] || ck == p.cc[ma])) return false;
    if (p.cc[pa] == p.cc[ma]) {
        const Indiv &ip = p.indiv[pa];
        for (size_t k = 0; k < ip.M.size(); k++) {
            if (ip.edge[k] > 1) continue;
            if (p.marr[ip.M[k]].ma == ma) {
                if (!p.marr[ip.M[k]].mark) p.marked.push_back(ip.M[k]);
                p.marr[ip.M[k]].mark = true;
                break;
            }
        }
    }
    return true;
}

static bool by_ll_desc(const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.second > b.second; }

// The swap block of _EvalMoves@0x100015989-0x100016651 with _ProposeSwapping@0x1000223d0: for every
// marriage of the focal individual x whose other parent z has parents itself, exchange the fathers
// (kind 1), the mothers (kind 2) or both (kind 3) between x's previous family and z's family; kinds
// 1 and 2 come in two flavours (siblings stay / siblings move too). All swaps of a step are scored
// by one engine call on the view of the three components involved. The list depends on the graph
// only, so it is drawn up (PlanSwaps) before the step's first engine submission, which queues it
// behind the sweep; the choices enter the list where the reference appends them (FinishSwaps).
struct SwapPlan {
    std::vector<pf_swap> sw;
    std::vector<Choice> ch;
    ViewBuf view;
    int cx = -1, P = -1, Q = -1;
    bool queued = false;            // submitted with the sweep: the values are the engine's deferred results
};

static void PlanSwaps(Ped &p, int x, bool redPa, bool redMa, SwapPlan &sp)
{
    const int P = p.prev[0], Q = p.prev[1], M0 = p.prev[2];
    if (P < 0) return;
    const int cx = p.cc[x];
    if (p.cc[P] == cx || p.cc[Q] == cx) return;
    const Indiv &ix = p.indiv[x];
    auto spouse = [&](int m) { return p.marr[m].pa == x ? p.marr[m].ma : p.marr[m].pa; };
    auto parents_marr = [&](int z) { const Indiv &iz = p.indiv[z]; for (size_t k = 0; k < iz.M.size(); k++) if (iz.edge[k] > 1) return iz.M[k]; return -1; };
    for (int m : ix.M) {                                       // @0x1000159fd-0x100015bec
        const int mz = parents_marr(spouse(m));
        if (mz < 0) continue;
        const int a = p.marr[mz].pa, b = p.marr[mz].ma;
        p.indiv[a].reductive = isReductiveParent(p, a);
        p.indiv[b].reductive = isReductiveParent(p, b);
    }
    auto propose = [&](int kind, int z, int mz, int mx) {      // _ProposeSwapping@0x1000223d0
        auto add = [&](int move) {
            sp.sw.push_back(pf_swap{kind, move, x, z, mz, mx, P, Q, M0, 0});
            Choice c{M0, P, Q, 0, 3};
            c.swapKind = kind; c.marrZ = mz; c.swapZ = move ? z : -1;
            sp.ch.push_back(c);
        };
        if (kind == 3) { add(1); return; }
        add(0);
        if (M0 == -1 && p.marr[mz].nEdges() == 3) return;      // @0x1000226c4-0x1000226f8
        add(1);
    };
    for (int mx : ix.M) {                                      // @0x100015bec-0x100016651
        const int z = spouse(mx);
        const int mz = parents_marr(z);
        if (mz < 0) continue;
        if (p.marr[mz].selfing || p.marr[mx].selfing || P == Q) continue;   // see the header of this file
        const int paz = p.marr[mz].pa, maz = p.marr[mz].ma;
        const bool rp = p.indiv[paz].reductive, rm = p.indiv[maz].reductive;
        if (redPa && redMa && rp && rm) continue;
        const bool k1 = !(redPa && rp), k2 = !(redMa && rm);
        if (k1) propose(1, z, mz, mx);
        if (k2) { propose(2, z, mz, mx); if (k1) propose(3, z, mz, mx); }
        // kind 4 (pairs of x's marriages, @0x1000160da-0x100016638) never fires in the reference
    }
    if (sp.sw.empty()) return;
    const double t0 = now_s();
    sp.cx = cx; sp.P = P; sp.Q = Q;
    std::vector<char> want(p.nCC, 0);
    want[cx] = want[p.cc[P]] = want[p.cc[Q]] = 1;
    build_view(p, want, sp.view);
    p.t_view += now_s() - t0;
}

static void FinishSwaps(Ped &p, SwapPlan &sp)
{
    if (sp.sw.empty()) return;
    const size_t first = p.ch.size();
    for (const Choice &c : sp.ch) { p.ch.push_back(c); p.chll.push_back(0.0); }
    std::vector<double> lsum(sp.sw.size());
    const double t1 = now_s();
    if (sp.queued) {
        if (PFN(deferred_results)(p.eng, (int)sp.sw.size(), lsum.data()) != PF_OK) die("deferred_results failed");
    } else if (eng_eval_swaps(p, &sp.view.v, (int)sp.sw.size(), sp.sw.data(), lsum.data()) != PF_OK) die("eval_swaps failed");
    p.t_eval += now_s() - t1;
    p.n_proposals += (long)sp.sw.size(); p.n_swaps += (long)sp.sw.size();
    const int cp = p.cc[sp.P], cq = p.cc[sp.Q];
    for (size_t k = 0; k < sp.sw.size(); k++)                   // _calculateBundlell@0x100022291-0x1000223a8
        p.chll[first + k] = pf_compose_swap_ll(p.LLtot, lsum[k], p.LLcc[sp.cx], p.LLcc[cp], cp != cq && cp != sp.cx, p.LLcc[cq]);
}

// the step's first submission: the pending sweep, the single-parent proposals and, queued behind them, the swaps
static int submit_step(Ped &p, const pf_graph_view *v, double *ll, int kid, Batch &b, double *lsum, SwapPlan *sp)
{
    if (mc::active || !sp || sp->sw.empty()) return eng_sweep_eval(p, v, ll, kid, (int)b.props.size(), b.props.data(), lsum);
    sp->queued = true;
    return PFN(sweep_eval_swaps)(p.eng, p.chain, v, ll, kid, (int)b.props.size(), b.props.data(), lsum, &sp->view.v, (int)sp->sw.size(), sp->sw.data());
}

// _EvalMoves@0x100015010
static void EvalMoves(Ped &p, const Opts &o, int kid)
{
    p.ch.clear(); p.chll.clear();
    Batch b;
    const double kt = p.indiv[kid].genTime;
    const int ck = p.cc[kid];
    bool redPa = false, redMa = false;
    const int ppa = p.prev[0], pma = p.prev[1];
    if (ppa >= 0) { redPa = p.indiv[ppa].reductive; redMa = p.indiv[pma].reductive; }
    if (redPa) p.indivActive[ppa] = 0;
    if (redMa) p.indivActive[pma] = 0;
    add_choice(p, b, PF_PROP_TRIO, -1, -1, -1);
    if (o.selfing) add_choice(p, b, PF_PROP_SELFING, -1, -1, -1);
    std::vector<std::pair<int, int>> cand_slot[3];
    if (kid >= p.nObs) {
        for (int c = 0; c < p.nIndivUsed; c++) ProposeSingletMove(p, o, b, c, kid, cand_slot);
    } else {
        for (int c : p.parentCand[kid]) ProposeSingletMove(p, o, b, c, kid, cand_slot);
        for (int c = p.nObs; c < p.nIndivUsed; c++) ProposeSingletMove(p, o, b, c, kid, cand_slot);
    }
    if (redPa) p.indivActive[ppa] = 1;
    if (redMa) p.indivActive[pma] = 1;
    SwapPlan sp;
    if (!o.diswap) PlanSwaps(p, kid, redPa, redMa, sp);
    sweep_and_flush(p, kid, b, &sp);                           // single-parent scores are needed for the top-10 cut
    for (int s = 0; s < 3; s++) {
        p.cand[s].clear();
        for (auto &cs : cand_slot[s]) p.cand[s].push_back({cs.first, p.chll[cs.second]});
        std::stable_sort(p.cand[s].begin(), p.cand[s].end(), by_ll_desc);      // qsort(_compareLLonGC)
        if (p.cand[s].size() > 10) p.cand[s].resize(10);
    }
    auto pair = [&](int pa, int ma) { if (!checkMatchingCC(p, pa, ma, kid)) add_choice(p, b, PF_PROP_TRIO, pa, ma, -1); };
    for (auto &m : p.cand[1]) {
        for (auto &f : p.cand[2]) pair(m.first, f.first);
        for (auto &u : p.cand[0]) pair(m.first, u.first);
    }
    for (auto &f : p.cand[2])
        for (auto &u : p.cand[0]) pair(u.first, f.first);
    for (size_t i = 0; i + 1 < p.cand[0].size(); i++)
        for (size_t j = i + 1; j < p.cand[0].size(); j++) pair(p.cand[0][i].first, p.cand[0][j].first);
    // prong proposals for marked marriages. The reference scans every marriage slot in order and tests
    // "active", the age window and the mark in that order, so a mark survives on a marriage that is
    // inactive or outside this kid's window; the marks are few, so they are kept in a list (p.marked)
    // and only those slots are visited, in slot order, with the same tests
    if (!p.marked.empty()) {
        std::sort(p.marked.begin(), p.marked.end());
        p.marked.erase(std::unique(p.marked.begin(), p.marked.end()), p.marked.end());
        size_t kept = 0;
        for (size_t q = 0; q < p.marked.size(); q++) {
            const int m = p.marked[q];
            Marr &mm = p.marr[m];
            if (!mm.mark) continue;                            // cleared since (marriage dissolved or slot reused)
            bool stays = !p.marrActive[m] || m >= p.nMarrUsed;
            if (!stays) {
                const double g1 = p.indiv[mm.pa].genTime - kt, g2 = p.indiv[mm.ma].genTime - kt;
                stays = g1 > p.maxAge || p.minAge > g1 || g2 > p.maxAge || p.minAge > g2;
            }
            if (stays) { p.marked[kept++] = m; continue; }
            mm.mark = false;
            if (p.cc[mm.pa] == ck) continue;
            add_choice(p, b, PF_PROP_PRONG, -1, -1, m);
        }
        p.marked.resize(kept);
    }
    flush(p, kid, b);
    FinishSwaps(p, sp);
}

// _AddObsFracPrior@0x100019800
static void AddObsFracPrior(Ped &p, int kid)
{
    const int g = p.indiv[kid].gen + 1;
    if (g >= (int)p.obsFrac.size()) return;
    const double f = p.obsFrac[g];
    if (f == 0 || p.nObsPerGen[g] == 0) return;
    double nhat = ceil((double)p.nObsPerGen[g] / f);
    int u = (int)(nhat - (double)p.nIndivPerGen[g]);
    if ((double)p.nIndivPerGen[g] > nhat) { nhat = (double)p.nIndivPerGen[g]; u = 0; }
    const float aa = (float)((1.0 + f) + (double)p.nIndivPerGen[g]);
    const float bb = (float)((1.0 + (1.0 - f)) + (double)u);
    const double q = rng::genbet(aa, bb), nq = 1.0 - q;
    for (size_t c = 0; c < p.ch.size(); c++) {
        const int pa = p.ch[c].pa, ma = p.ch[c].ma;
        double add;
        if (pa == -1 && ma == -1) add = 2.0 * log(nq);
        else if (pa == -1 || ma == -1) {
            const int given = pa == -1 ? ma : pa;
            add = given < p.nObs ? (log(nq) + log(2.0)) + log(q) : (log(nq) + log(2.0)) + log(nq);
        } else if (pa >= p.nObs && ma >= p.nObs) add = 2.0 * log(nq);
        else if (pa < p.nObs && ma < p.nObs) add = 2.0 * log(q);
        else add = (log(nq) + log(2.0)) + log(q);
        p.chll[c] = add + p.chll[c];
    }
}

// _NormalizeAndSelect@0x100016660
static void NormalizeAndSelect(Ped &p)
{
    const int n = (int)p.ch.size();
    if (n == 0) { p.neu[0] = p.prev[0]; p.neu[1] = p.prev[1]; p.neu[2] = -1; return; }
    const double thresh = log(pow(10.0, -15.0)) - log((double)n);
    double mx = p.chll[0];
    for (int i = 0; i < n; i++) if (p.chll[i] > mx) mx = p.chll[i];
    std::vector<double> w(n);
    for (int i = 0; i < n; i++) { const double d = p.chll[i] - mx; w[i] = d < thresh ? 0.0 : exp(d); }
    double sum = 0, c = 0;
    for (int i = 0; i < n; i++) {
        if (w[i] == 0) continue;
        const double y = w[i] - c, t = sum + y;
        c = (t - sum) - y; sum = t;
    }
    const double u = (double)rng::genunf(0.0f, 1.0f);
    int pick = -1;
    double cum = 0; c = 0;
    for (int i = 0; i < n; i++) {
        if (w[i] == 0) continue;
        const double y = w[i] / sum - c, t = cum + y;
        c = (t - cum) - y; cum = t;
        if (!(cum < u)) { pick = i; break; }
    }
    if (pick == -1) { fprintf(stderr, "Error on calculating a series of cumulative likelihoods \n"); exit(-1); }
    p.picked = pick;
    p.neu[0] = p.ch[pick].pa; p.neu[1] = p.ch[pick].ma; p.neu[2] = p.ch[pick].mnode; p.neu[3] = p.ch[pick].selfing;
}

// _GetEmptyIndivIndx@0x100012ad0 / _GetEmptyMarrIndx@0x100012b80
static int GetEmptyIndivIndx(Ped &p, int gen)
{
    p.nIndivPerGen[gen]++;
    if (!p.freeIndiv.empty()) { const int s = p.freeIndiv.back(); p.freeIndiv.pop_back(); return s; }
    if (p.nIndivUsed >= p.capIndiv) { fprintf(stderr, "out of individual slots (%d)\n", p.capIndiv); exit(-1); }
    return p.nIndivUsed++;
}
static int GetEmptyMarrIndx(Ped &p)
{
    if (!p.freeMarr.empty()) { const int s = p.freeMarr.back(); p.freeMarr.pop_back(); return s; }
    if (p.nMarrUsed >= p.capMarr) { fprintf(stderr, "out of marriage slots (%d)\n", p.capMarr); exit(-1); }
    return p.nMarrUsed++;
}
// _AttachUnobsIndiv@0x100012c00
static void AttachUnobsIndiv(Ped &p, int slot, int kid, int m, int edge)
{
    Indiv &x = p.indiv[slot];
    // the reference re-initialises only sex, gen, genTime, founder, observed, ids and the marriage
    // list of a recycled slot (@0x100012c15-0x100012d67): the previous occupant's unobsDegree,
    // `reductive` flag and sexWasUnknown stay (a fresh slot has 0 / false from _AllocIndiv), and the
    // degree of the spouse created just before is what UpdateUnobsDegree then reads
    const int staleDegree = x.unobsDegree, staleSexUnknown = x.sexWasUnknown;
    const bool staleReductive = x.reductive, staleQueued = x.queued;
    x = Indiv();
    x.unobsDegree = staleDegree; x.sexWasUnknown = staleSexUnknown; x.reductive = staleReductive; x.queued = staleQueued;
    x.sex = edge + 1;
    x.gen = p.indiv[kid].gen + 1;
    x.genTime = p.indiv[kid].genTime + p.minAge;
    x.founder = true; x.observed = false;
    x.userID = p.maxID + slot;
    p.indivActive[slot] = 1;
    p.cc[slot] = p.cc[kid];
    if (p.cc[slot] >= 0) p.ccMem[p.cc[slot]].push_back(slot);
    x.M = {m}; x.edge = {edge};
}

// _PruneIndiv@0x1000141f0
static void PruneIndiv(Ped &p, int id)
{
    hist_node(p, 2, id);
    Indiv &x = p.indiv[id];
    int m = -1;
    for (size_t k = 0; k < x.M.size(); k++) if (x.edge[k] > 1) { m = x.M[k]; break; }
    x.founder = false;
    if (m == -1) {
        p.prev[0] = p.prev[1] = -1; p.prev[2] = -1;
        mark_dirty(p, p.cc[id]);
        return;
    }
    Marr &mm = p.marr[m];
    const int pa = mm.pa, ma = mm.ma;
    p.prev[0] = pa; p.prev[1] = ma;
    hist_edge(p, 0, m, id);
    if (mm.nEdges() == 3) {
        p.marrActive[m] = 0; mm.mark = false;
        p.freeMarr.push_back(m);
        hist_edge(p, 0, m, pa);
        if (!mm.selfing) hist_edge(p, 0, m, ma);
        hist_mnode(p, 0, m);
        DissociateParentMnode(p, pa, m);
        if (!mm.selfing) DissociateParentMnode(p, ma, m);
        DissociateParentMnode(p, id, m);
        mm.kids.clear();
        p.prev[2] = -1;
    } else {
        p.prev[2] = m;
        DissociateSibMnode(p, id, m);
        DissociateParentMnode(p, id, m);
    }
    UpdateConComp(p, id);
    UpdateUnobsDegree(p, pa);
    UpdateUnobsDegree(p, ma);
    p.indiv[pa].reductive = isReductiveParent(p, pa);
    p.indiv[ma].reductive = isReductiveParent(p, ma);
}

// The swap branches of _PickAndUpdate@0x100017f5a-0x100019281. The kid has just been re-attached
// to its previous parents in marriage m (sp, sq = the selected pa, ma); now the fathers (kind 1),
// mothers (kind 2) or both (kind 3) of m and of marr_z change places, and when the choice carries z
// (aux[1] >= 0) the two sibships do too: m keeps the kid and gets z's siblings, marr_z keeps z and
// gets the kid's siblings.
static void ApplySwap(Ped &p, const Choice &c, int kid, int m, int sp, int sq)
{
    const int mz = c.marrZ;
    auto retarget_last = [&](int who, int from, int to) {      // search from the end (@0x100017fda-0x100018068)
        Indiv &w = p.indiv[who];
        for (int k = (int)w.M.size() - 1; k >= 0; k--) if (w.M[k] == from) { w.M[k] = to; return; }
    };
    auto retarget_first = [&](int who, int from, int to) {     // search from the start (@0x10001808e-0x100018135)
        Indiv &w = p.indiv[who];
        for (size_t k = 0; k < w.M.size(); k++) if (w.M[k] == from) { w.M[k] = to; return; }
    };
    const int paz = p.marr[mz].pa, maz = p.marr[mz].ma;
    if (c.swapKind == 1) {
        retarget_last(sp, m, mz); p.marr[m].pa = paz;
        retarget_first(paz, mz, m); p.marr[mz].pa = sp;
    } else if (c.swapKind == 2) {
        retarget_last(sq, m, mz); p.marr[m].ma = maz;
        retarget_first(maz, mz, m); p.marr[mz].ma = sq;
    } else {
        retarget_last(sp, m, mz); p.marr[m].pa = paz;
        retarget_last(sq, m, mz); p.marr[m].ma = maz;
        retarget_first(maz, mz, m); retarget_first(paz, mz, m);
        p.marr[mz].pa = sp; p.marr[mz].ma = sq;
    }
    if (c.swapZ < 0) return;
    const int z = c.swapZ;                                      // @0x100018b9e-0x10001927f
    Marr &A = p.marr[m], &B = p.marr[mz];
    const int nA = (int)A.kids.size(), nB = (int)B.kids.size();    // vf0 - 2, vf4 - 2
    A.kids.resize(std::max(nA, nB) + 1); B.kids.resize(nA + nB + 1);
    auto set_entry = [&](int who, int from, int to, int edge) {
        Indiv &w = p.indiv[who];
        for (size_t k = 0; k < w.M.size(); k++) if (w.M[k] == from) { w.M[k] = to; w.edge[k] = edge; return; }
    };
    int moved = 0;
    for (int i = 0; i < nA - 1; i++) {                         // the kid's siblings queue up behind marr_z's kids
        if (A.kids[i] == kid) continue;
        B.kids[nB + moved++] = A.kids[i];
    }
    if (nA > 1) {
        Indiv &w = p.indiv[kid];
        for (int k = (int)w.M.size() - 1; k >= 0; k--) if (w.M[k] == m) { w.edge[k] = 2; break; }
        A.kids[0] = kid;
    }
    if (nB > 1) {
        int pos = 0;
        for (int i = 0; i < nB; i++) {
            const int y = B.kids[i];
            if (y == z) { set_entry(z, mz, mz, 2); B.kids[0] = z; }
            else { pos++; set_entry(y, mz, m, pos + 2); A.kids[pos] = y; }
        }
    }
    for (int j = 0; j < moved; j++) {
        const int y = B.kids[nB + j];
        set_entry(y, m, mz, j + 1 + 2);
        B.kids[j + 1] = y;
    }
    A.kids.resize(nB); B.kids.resize(nA);
}

// _PickAndUpdate@0x100016a60
static void PickAndUpdate(Ped &p, int kid)
{
    NormalizeAndSelect(p);
    bool draw = true;
    if (p.neu[0] != -1 && p.indiv[p.neu[0]].sex != 0) draw = false;
    if (p.neu[1] != -1 && p.indiv[p.neu[1]].sex != 0) draw = false;
    if (draw) {
        const double u = (double)rng::genunf(0.0f, 1.0f);
        if (!(u < 0.5)) std::swap(p.neu[0], p.neu[1]);
    }
    fprintf(p.fLL, "%d %d %f\n", p.iter, kid, p.picked >= 0 ? p.chll[p.picked] : 0.0);
    const bool swap = p.picked >= 0 && p.ch[p.picked].kind == 3;
    if (!swap) {                                               // @0x100016be2: a swap keeps the previous parents
        // Trimming an ex-parent cascades to its unobserved ancestors, and only the ex-parents themselves
        // are hidden from the enumeration: the choice can be such an ancestor (or the marriage of two).
        // The reference trims first and then links the kid to the freed slot (_PickAndUpdate@0x100016cf5
        // uses new[] unchecked). Here the chosen individuals and marriage survive the trimming, with
        // their own ancestors, so the pedigree after the step is the one that was scored.
        p.keep[0] = p.neu[0]; p.keep[1] = p.neu[1]; p.keep[2] = p.neu[2];
        if (p.neu[0] != p.prev[0] && p.prev[0] > -1) {
            TrimPed(p, p.prev[0]); UpdateConComp(p, p.prev[0]); RevertSex(p, p.prev[0]);
        }
        if (p.prev[0] != p.prev[1] && p.neu[1] != p.prev[1] && p.prev[1] > -1) {
            TrimPed(p, p.prev[1]); UpdateConComp(p, p.prev[1]); RevertSex(p, p.prev[0]);   // sic: prev.pa
        }
        p.keep[0] = p.keep[1] = p.keep[2] = -1;
    }
    const int sel_pa = p.neu[0], sel_ma = p.neu[1];            // v54 / v58 of the swap rewiring below
    int m = p.neu[2], pa = p.neu[0], ma = p.neu[1];
    Indiv &k = p.indiv[kid];
    if (m != -1) {
        Marr &mm = p.marr[m];
        mm.kids.push_back(kid);
        k.M.push_back(m); k.edge.push_back(mm.nEdges() - 1);
        pa = mm.pa; ma = mm.ma;
        AssignConComp(p, p.cc[pa], p.cc[kid]);
        hist_edge(p, 1, m, kid);
    } else {
        m = GetEmptyMarrIndx(p);
        Marr &mm = p.marr[m];
        mm = Marr();
        const bool self = p.neu[3] == 1;
        auto attach_parent = [&](int &par, int edge) {
            if (par == -1) {
                par = GetEmptyIndivIndx(p, k.gen + 1);
                AttachUnobsIndiv(p, par, kid, m, edge);
                hist_node(p, 1, par);
            } else {
                Indiv &ip = p.indiv[par];
                ip.M.push_back(m); ip.edge.push_back(edge);
                if (!self && ip.sex == 0) ip.sex = edge + 1;
                AssignConComp(p, p.cc[par], p.cc[kid]);
            }
        };
        attach_parent(pa, 0);
        if (self) ma = pa; else attach_parent(ma, 1);
        k.M.push_back(m); k.edge.push_back(2);
        mm.pa = pa; mm.ma = ma; mm.kids = {kid}; mm.selfing = self; mm.mark = false;
        p.marrActive[m] = 1;
        hist_mnode(p, 1, m);
        hist_edge(p, 1, m, pa);
        if (!self) hist_edge(p, 1, m, ma);
        hist_edge(p, 1, m, kid);
    }
    hist_node(p, 3, kid);
    if (swap) { pa = sel_pa; ma = sel_ma; ApplySwap(p, p.ch[p.picked], kid, m, pa, ma); p.n_swaps_picked++; }
    mark_dirty(p, p.cc[kid]);
    p.neu[0] = pa; p.neu[1] = ma; p.neu[2] = m;
    p.indiv[pa].sex = 1; p.indiv[ma].sex = 2;
    UpdateUnobsDegree(p, pa); UpdateUnobsDegree(p, ma); UpdateUnobsDegree(p, kid);
    if (swap) {                                                // @0x10001938b-0x1000195f3: one row per kid of both families
        auto family = [&](int mm_, bool parents) {
            const Marr &f = p.marr[mm_];
            if (parents) { UpdateUnobsDegree(p, f.pa); UpdateUnobsDegree(p, f.ma); }
            for (int c : f.kids) {
                UpdateUnobsDegree(p, c);
                fprintf(p.fPed, "%d %d %d %d\n", p.iter, p.indiv[c].userID, p.indiv[f.pa].userID, p.indiv[f.ma].userID);
            }
        };
        family(m, false);
        family(p.ch[p.picked].marrZ, true);
    } else
    fprintf(p.fPed, "%d %d %d %d\n", p.iter, k.userID, p.indiv[pa].userID, p.indiv[ma].userID);
    // the reference calls Propagating_msg here; the dirty flags carry over to the sweep at the
    // start of the next Gibbs step instead (see the header of this file)
}

// PEDFAC_DUMP=path: the pedigree state as text (debug aid; read back by tools/replay_state.py)
static void DumpState(Ped &p, const char *path)
{
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "T %.17g %d\n", p.LLtot, p.nCC);
    for (int i = 0; i < p.nIndivUsed; i++)
        if (p.indivActive[i]) fprintf(f, "I %d %d %d %d %d %d %d\n", i, p.indiv[i].userID, p.indiv[i].observed ? 1 : 0, p.indiv[i].founder ? 1 : 0, p.cc[i], p.indiv[i].gen, p.indiv[i].sex);
    for (int m = 0; m < p.nMarrUsed; m++)
        if (p.marrActive[m]) {
            fprintf(f, "M %d %d %d %d", m, p.marr[m].pa, p.marr[m].ma, p.marr[m].selfing ? 1 : 0);
            for (int k : p.marr[m].kids) fprintf(f, " %d", k);
            fprintf(f, "\n");
        }
    for (int c = 0; c < p.nCC; c++) fprintf(f, "C %d %.17g\n", c, p.LLcc[c]);
    fclose(f);
}

// PEDFAC_CHECK: structural invariants of the pedigree state (every member of an active marriage is
// active, carries the marriage in its own list with the right edge index and shares its component)
static void CheckInvariants(Ped &p, const char *where, int id)
{
    auto bad = [&](const char *what, int a, int b) {
        fprintf(stderr, "PEDFAC_CHECK: %s after %s of %d (iter %d): %d %d\n", what, where, id, p.iter, a, b);
        exit(4);
    };
    for (int m = 0; m < p.nMarrUsed; m++) {
        if (!p.marrActive[m]) continue;
        const Marr &mm = p.marr[m];
        const int ne = mm.nEdges();
        for (int ed = 0; ed < ne; ed++) {
            if (ed == 1 && mm.selfing) continue;
            const int x = ed == 0 ? mm.pa : ed == 1 ? mm.ma : mm.kids[ed - 2];
            if (x < 0 || x >= p.nIndivUsed || !p.indivActive[x]) bad("inactive member of marriage", m, x);
            if (p.cc[x] != p.cc[mm.pa]) bad("member in another component than the father of marriage", m, x);
            bool found = false;
            for (size_t k = 0; k < p.indiv[x].M.size(); k++) if (p.indiv[x].M[k] == m && p.indiv[x].edge[k] == ed) found = true;
            if (!found) bad("marriage missing from its member's list", m, x);
        }
    }
    for (int i = 0; i < p.nIndivUsed; i++) {
        if (!p.indivActive[i]) continue;
        if (p.cc[i] < 0 || p.cc[i] >= p.nCC) bad("component label out of range", i, p.cc[i]);
        for (size_t k = 0; k < p.indiv[i].M.size(); k++) {
            const int m = p.indiv[i].M[k], ed = p.indiv[i].edge[k];
            if (m < 0 || m >= p.nMarrUsed || !p.marrActive[m]) bad("individual lists an inactive marriage", i, m);
            const Marr &mm = p.marr[m];
            const int x = ed == 0 ? mm.pa : ed == 1 ? mm.ma : (ed - 2 < (int)mm.kids.size() ? mm.kids[ed - 2] : -1);
            if (x != i) bad("edge index does not point back", i, m);
        }
    }
    // the member lists hold exactly the active individuals of every label
    long listed = 0, active = 0;
    for (int c = 0; c < (int)p.ccMem.size(); c++) {
        if (c >= p.nCC && !p.ccMem[c].empty()) bad("members listed under a retired label", c, (int)p.ccMem[c].size());
        for (int i : p.ccMem[c]) { if (!p.indivActive[i] || p.cc[i] != c) bad("stale entry in a component's member list", c, i); listed++; }
    }
    for (int i = 0; i < p.nIndivUsed; i++) active += p.indivActive[i] ? 1 : 0;
    if (listed != active) bad("member lists and active individuals differ in number", (int)listed, (int)active);
}

// _MCMC_sampling@0x100019ec0
static void MCMC_sampling(Ped &p, const Opts &o, int id)
{
    static const bool dbg_calls = getenv("PEDFAC_DEBUG_CALLS") != nullptr;
    if (dbg_calls) fprintf(stderr, "call %d gen %d deg %d\n", id, p.indiv[id].gen, p.indiv[id].unobsDegree);
    static const bool dbg_state = getenv("PEDFAC_DEBUG_STATE") != nullptr, dbg_ll = getenv("PEDFAC_DEBUG_LL") != nullptr, dbg_step = getenv("PEDFAC_DEBUG") != nullptr;
    static const char *dump_pre = getenv("PEDFAC_DUMP_PRE"), *dump_kid = getenv("PEDFAC_DUMP_KID");
    if (dbg_state) {                                          // same format as oracle/macho_run.c's PEDFAC_REF_DUMP
        fprintf(stderr, "state before %d nCC %d\n", id, p.nCC);
        for (int i = 0; i < p.nIndivUsed; i++) {
            if (!p.indivActive[i]) continue;
            const Indiv &x = p.indiv[i];
            fprintf(stderr, "I %d gen %d sex %d deg %d founder %d red %d queued %d cc %d :", i, x.gen, x.sex, x.unobsDegree, x.founder ? 1 : 0, x.reductive ? 1 : 0, x.queued ? 1 : 0, p.cc[i]);
            for (size_t k = 0; k < x.M.size(); k++) fprintf(stderr, " %d/%d", x.M[k], x.edge[k]);
            fprintf(stderr, "\n");
        }
        for (int m = 0; m < p.nMarrUsed; m++) {
            if (!p.marrActive[m]) continue;
            fprintf(stderr, "M %d pa %d ma %d self %d kids", m, p.marr[m].pa, p.marr[m].selfing ? -1 : p.marr[m].ma, p.marr[m].selfing ? 1 : 0);
            for (int k : p.marr[m].kids) fprintf(stderr, " %d", k);
            fprintf(stderr, "\n");
        }
    }
    if (!(p.indiv[id].gen < p.nGen - 1)) return;
    if (!(p.indiv[id].unobsDegree <= p.maxUnobsLayer)) return;
    p.t++;
    p.n_steps++;
    const double t_0 = now_s();
    PruneIndiv(p, id);
    if (p.check) CheckInvariants(p, "PruneIndiv", id);
    const double t_1 = now_s();
    EvalMoves(p, o, id);                                       // runs the pending sweep together with its first batch
    const double t_2 = now_s();
    p.t_prune += t_1 - t_0; p.t_moves += t_2 - t_1;
    p.t++;
    if (p.check) p.chll_data = p.chll;
    if (dump_pre && dump_kid && atoi(dump_kid) == id) DumpState(p, dump_pre);   // last visit wins
    if (dbg_ll) {                          // same format as oracle/macho_run.c's PEDFAC_REF_DUMP_LL
        fprintf(stderr, "choices kid %d n %zu\n", id, p.ch.size());
        for (size_t k = 0; k < p.ch.size(); k++)
            fprintf(stderr, "c %zu ids %d %d %d %d %.17g\n", k, p.ch[k].mnode, p.ch[k].pa, p.ch[k].ma, p.ch[k].selfing, p.chll[k]);
    }
    AddObsFracPrior(p, id);
    const int dbg_prev[3] = {p.prev[0], p.prev[1], p.prev[2]};
    PickAndUpdate(p, id);
    p.t_pick += now_s() - t_2;
    if (dbg_step) fprintf(stderr, "step iter %d kid %d prev %d %d %d -> new %d %d %d self %d picked %d kind %d\n", p.iter, id, dbg_prev[0], dbg_prev[1], dbg_prev[2], p.neu[0], p.neu[1], p.neu[2], p.neu[3], p.picked, p.picked >= 0 ? p.ch[p.picked].kind : -1);
    if (p.check) CheckInvariants(p, "PickAndUpdate", id);
    if (p.check && p.picked >= 0 && !p.ch.empty()) {
        // the log-likelihood a proposal was scored with is the total the pedigree has once the move
        // is made: re-propagate now and compare (every move kind, swaps included)
        Propagating_msg(p);
        const double want = p.chll_data[p.picked], got = p.LLtot;
        bool in_range = true;                                  // components beyond the double range cannot be checked
        for (int c = 0; c < p.nCC; c++) in_range = in_range && !(p.LLcc[c] < kUnderflowMeanPerLocus * p.S);
        if (in_range && !(fabs(want - got) <= 1e-9 * std::max(1.0, fabs(got)))) {
            if (getenv("PEDFAC_DUMP")) DumpState(p, getenv("PEDFAC_DUMP"));
            fprintf(stderr, "PEDFAC_CHECK: iter %d kid %d choice %d (kind %d swap %d/%d): scored %.12f, pedigree has %.12f\n",
                    p.iter, id, p.picked, p.ch[p.picked].kind, p.ch[p.picked].swapKind, p.ch[p.picked].swapZ, want, got);
            exit(3);
        }
    }
    const int pa = p.neu[0], ma = p.neu[1], self = p.neu[3];
    auto enqueue = [&](int x) { p.unobsQueue.push_back(x); p.indiv[x].queued = true; };
    const bool qa = pa >= p.nObs && !p.indiv[pa].queued, qm = ma >= p.nObs && !p.indiv[ma].queued;
    if (qa && qm) {
        if (self == 1) enqueue(pa);
        else {
            const double u = (double)rng::genunf(0.0f, 1.0f);
            if (u > 0.5) { enqueue(pa); enqueue(ma); } else { enqueue(ma); enqueue(pa); }
        }
    } else if (qa) enqueue(pa);
    else if (qm) enqueue(ma);
}

// ---------------------------------------------------------------------------------------------
// input (_ReadPedfile@0x1000094f0, _CalculateMaxSize@0x1000079c0)

// ---------------------------------------------------------------------------------------------
// Several GPUs (no counterpart in the reference, which is one thread on one core): PEDFAC_GPUS=N makes
// `pedigraph` fork one process per GPU. Every process runs the identical chain — same seed, same host
// logic — on its own slice of the loci; libpedfac_cuda adds the slices' partial sums across the GPUs
// inside its own kernels (pf_comm_*), so every process sees the same log-likelihoods and makes the
// same moves. Rank 0 writes the output files and stdout; the others write under out/.shard<r>/.
static int g_shard_rank = 0, g_shard_world = 1;
static int g_n_chains = 1;               // PEDFAC_CHAINS: independent chains batched on this GPU
static int g_n_threads = 1;              // PEDFAC_THREADS: scheduler threads the chains are dealt to, one engine each
// what was uploaded to the first engine, kept to fill the engines of the other scheduler threads
struct Upload { int slot, kind; std::vector<uint8_t> codes; std::vector<uint16_t> num; std::vector<double> lik; };
static std::vector<Upload> g_uploads;
static pf_config g_cfg;
static int chains_of_thread(int t) { return (g_n_chains * (t + 1)) / g_n_threads - (g_n_chains * t) / g_n_threads; }
static std::string g_shard_dir;          // where the ranks exchange their mailbox handles

#ifndef PF_NO_COMM
static void ConnectShards(Ped &p)
{
    if (g_shard_world <= 1) return;
    char mine[PF_COMM_HANDLE_BYTES];
    if (PFN(comm_local_handle)(p.eng, g_shard_world, mine) != PF_OK) die("cannot create the cross-GPU mailbox");
    const char *nonce = getenv("PEDFAC_SHARD_NONCE") ? getenv("PEDFAC_SHARD_NONCE") : "0";
    auto path = [&](int r, bool tmp) { return g_shard_dir + "/.pf_comm_" + nonce + "_" + std::to_string(r) + (tmp ? ".tmp" : ""); };
    {
        FILE *f = fopen(path(g_shard_rank, true).c_str(), "wb");
        if (!f || fwrite(mine, 1, sizeof mine, f) != sizeof mine) die("cannot write the mailbox handle");
        fclose(f);
        rename(path(g_shard_rank, true).c_str(), path(g_shard_rank, false).c_str());
    }
    std::vector<char> all((size_t)g_shard_world * PF_COMM_HANDLE_BYTES);
    for (int r = 0; r < g_shard_world; r++) {
        FILE *f = nullptr;
        for (int tries = 0; tries < 6000 && !(f = fopen(path(r, false).c_str(), "rb")); tries++) usleep(10000);
        if (!f || fread(&all[(size_t)r * PF_COMM_HANDLE_BYTES], 1, PF_COMM_HANDLE_BYTES, f) != PF_COMM_HANDLE_BYTES) die("a GPU shard did not come up");
        fclose(f);
    }
    if (PFN(comm_connect)(p.eng, g_shard_rank, g_shard_world, all.data()) != PF_OK) die("cannot connect the GPU shards (peer access between the GPUs is required)");
}
#else
static void ConnectShards(Ped &) { if (g_shard_world > 1) die("this build has one likelihood engine: PEDFAC_GPUS is for the CUDA build"); }
#endif

static std::vector<std::string> split(const std::string &s, char sep)
{
    std::vector<std::string> out;
    size_t i = 0;
    while (i <= s.size()) {
        size_t j = s.find(sep, i);
        if (j == std::string::npos) j = s.size();
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j + 1;
    }
    return out;
}

static bool read_line(FILE *f, std::string &line)
{
    line.clear();
    int c;
    bool any = false;
    while ((c = fgetc(f)) != EOF) { any = true; if (c == '\n') break; if (c != '\r') line.push_back((char)c); }
    return any;
}

static void ReadPrior(Ped &p, const std::string &path)
{
    FILE *f = fopen(path.c_str(), "r");
    if (!f) { fprintf(stderr, "Failed to open prior.txt for reading\n"); exit(-1); }
    std::string line;
    while (read_line(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        auto tok = split(line, ' ');
        if (tok.empty()) continue;
        const std::string &k = tok[0];
        auto ival = [&]() { return tok.size() > 1 ? atoi(tok[1].c_str()) : 0; };
        auto dvec = [&](size_t n) { std::vector<double> v(n, 0.0); for (size_t i = 0; i < n && i + 1 < tok.size(); i++) v[i] = atof(tok[i + 1].c_str()); return v; };
        if (k == "nIndiv") p.nIndivDecl = ival();
        else if (k == "nObsIndiv") p.nObs = ival();
        else if (k == "nGen") p.nGen = ival();
        else if (k == "nSNP") p.S = ival();
        else if (k == "nMar") p.nMarDecl = ival();
        else if (k == "maxID") p.maxID = ival();
        else if (k == "aFreq") p.afreq = dvec(p.S);
        else if (k == "epsilon") p.eps = dvec(p.S);
        else if (k == "obsFrac") p.obsFrac = dvec(p.nGen);
        else if (k == "maxAge") p.maxAge = tok.size() > 1 ? atof(tok[1].c_str()) : 1.0;
        else if (k == "minAge") p.minAge = tok.size() > 1 ? atof(tok[1].c_str()) : 1.0;
        else if (k == "maxUnobsLayer") p.maxUnobsLayer = ival();
        else if (k == "maxMarrGap") { /* parsed and ignored by the reference (@0x100009d47) */ }
        else { fprintf(stderr, "In prior.txt, I cannot recognize this keyword: %s. \n", k.c_str()); exit(-1); }
    }
    fclose(f);
    if (p.S <= 0 || p.nGen <= 0 || (int)p.afreq.size() != p.S || (int)p.eps.size() != p.S) {
        fprintf(stderr, "prior.txt: nSNP, nGen, aFreq and epsilon are required\n"); exit(-1);
    }
    if ((int)p.obsFrac.size() != p.nGen) p.obsFrac.assign(p.nGen, 0.0);
}

// exact decimal numerators: "0.9800" -> 9800 over 10000 when the token has <= 4 fractional digits
static bool parse_dec4(const char *s, unsigned &num)
{
    unsigned long ip = 0; int nd = 0, nf = 0; unsigned long fp = 0;
    const char *c = s;
    while (*c >= '0' && *c <= '9') { ip = ip * 10 + (*c - '0'); c++; nd++; if (ip > 6) return false; }
    if (*c == '.') { c++; while (*c >= '0' && *c <= '9') { if (nf == 4) { if (*c != '0') return false; } else { fp = fp * 10 + (*c - '0'); nf++; } c++; } }
    if (*c != 0 || (nd == 0 && nf == 0)) return false;
    while (nf < 4) { fp *= 10; nf++; }
    const unsigned long v = ip * 10000 + fp;
    if (v > 65535) return false;
    num = (unsigned)v;
    return true;
}

// geno.txt is the one big input (tens of MB: one token per individual and locus): its lines are read
// with getline() and tokenised in place; the generic read_line()/split() pair above allocates a
// string per token and took most of the sampler's start-up time.
struct LineReader {
    FILE *f; char *buf = nullptr; size_t cap = 0; size_t len = 0;
    explicit LineReader(FILE *f_) : f(f_) {}
    ~LineReader() { free(buf); }
    bool next()                                 // like read_line: drops the newline and every '\r'
    {
        const ssize_t n = getline(&buf, &cap, f);
        if (n < 0) return false;
        size_t w = 0;
        for (ssize_t i = 0; i < n; i++) if (buf[i] != '\n' && buf[i] != '\r') buf[w++] = buf[i];
        buf[w] = 0; len = w;
        return true;
    }
};
// next run of non-blank characters at or after c; returns false at the end of the line
static bool next_token(const char *&c, const char *&b, const char *&e)
{
    while (*c == ' ') c++;
    if (!*c) return false;
    b = c;
    while (*c && *c != ' ') c++;
    e = c;
    return true;
}
static double tok_strtod(const char *b, const char *e)
{
    char tmp[64];
    const size_t n = std::min<size_t>((size_t)(e - b), sizeof tmp - 1);
    memcpy(tmp, b, n); tmp[n] = 0;
    return strtod(tmp, nullptr);
}
static bool parse_dec4(const char *b, const char *e, unsigned &num)
{
    char tmp[32];
    if ((size_t)(e - b) >= sizeof tmp) return false;
    memcpy(tmp, b, (size_t)(e - b)); tmp[e - b] = 0;
    return parse_dec4(tmp, num);
}

static void ReadGeno(Ped &p, const std::string &path)
{
    // pass 1 (_CalculateMaxSize): observed individuals per generation -> capacities
    FILE *f = fopen(path.c_str(), "r");
    if (!f) { fprintf(stderr, "Failed to open geno.txt for reading\n"); exit(-1); }
    LineReader lr(f);
    std::vector<int> nPerGen(p.nGen, 0);
    while (lr.next()) {
        if (lr.len == 0 || lr.buf[0] == '#') continue;
        const char *c = lr.buf, *b, *e;
        const char *tb[4], *te[4];
        int nt = 0;
        while (nt < 4 && next_token(c, b, e)) { tb[nt] = b; te[nt] = e; nt++; }
        if (nt < 4 || (int)tok_strtod(tb[1], te[1]) == 0) continue;      // atoi semantics for plain integers
        const int g = (int)tok_strtod(tb[3], te[3]);
        if (g >= 0 && g < p.nGen) nPerGen[g]++;
    }
    p.capIndiv = 0; p.capMarr = 0;
    for (int g = 0; g < p.nGen; g++) {
        p.capIndiv += nPerGen[g] * (int)(pow(2.0, (double)(p.nGen - g)) - 1.0);
        p.capMarr += nPerGen[g] * (int)(pow(2.0, (double)(p.nGen - g - 1)) - 1.0);
    }
    p.capMarr += 1;
    p.capIndiv = std::max(p.capIndiv, p.nIndivDecl) + 1;
    printf("# of marriage allowed/allocated for: %d\n", p.capMarr);
    printf("# of individuals allocated for : %d\n", p.capIndiv);
    p.indiv.assign(p.capIndiv, Indiv());
    p.marr.assign(p.capMarr, Marr());
    p.indivActive.assign(p.capIndiv, 0); p.marrActive.assign(p.capMarr, 0);
    p.cc.assign(p.capIndiv, -1); p.ccDirty.assign(p.capIndiv + 1, 0); p.LLcc.assign(p.capIndiv + 1, 0.0);
    p.nObsPerGen.assign(p.nGen + 1, 0); p.nIndivPerGen.assign(p.nGen + 1, 0);
    p.parentCand.assign(p.capIndiv, {});

    pf_config cfg{};
    cfg.n_loci = p.S; cfg.locus_begin = 0; cfg.locus_end = p.S; cfg.cap_indiv = p.capIndiv; cfg.cap_marr = p.capMarr;
    cfg.n_chains = chains_of_thread(0); cfg.device = getenv("PEDFAC_DEVICE") ? atoi(getenv("PEDFAC_DEVICE")) : 0;
    if (g_shard_world > 1) {                                  // this process owns one slice of the loci (see main)
        cfg.locus_begin = (int)((long long)p.S * g_shard_rank / g_shard_world);
        cfg.locus_end = (int)((long long)p.S * (g_shard_rank + 1) / g_shard_world);
        if (cfg.locus_end <= cfg.locus_begin) die("more GPU shards than loci");
    }
    g_cfg = cfg;
    if (PFN(create)(&p.eng, &cfg) != PF_OK) die("cannot create the likelihood engine");
    ConnectShards(p);
    if (PFN(set_locus_params)(p.eng, p.eps.data(), p.afreq.data()) != PF_OK) die("set_locus_params");

    // pass 2: rows. Observed rows fill slots 0..nObs-1 in file order, unobserved rows follow
    rewind(f);
    int nextObs = 0, nextUnobs = p.nObs, row = 0;
    std::vector<uint8_t> codes(p.S);
    std::vector<double> lik((size_t)p.S * 3);
    std::vector<uint16_t> num((size_t)p.S * 3);
    while (lr.next()) {
        if (lr.len == 0 || lr.buf[0] == '#') continue;
        const char *c = lr.buf, *b, *e;
        const char *tb[5], *te[5];
        int nt = 0;
        while (nt < 5 && next_token(c, b, e)) { tb[nt] = b; te[nt] = e; nt++; }
        if (nt < 5) continue;
        auto tok_int = [&](int k) { char tmp[32]; const size_t n = std::min<size_t>((size_t)(te[k] - tb[k]), sizeof tmp - 1); memcpy(tmp, tb[k], n); tmp[n] = 0; return atoi(tmp); };
        const bool obs = tok_int(1) != 0;
        int slot;
        if (obs) {
            if (nextObs >= p.nObs) { fprintf(stderr, "There are more observed individuals in your genoInfo.txt than what you have declared\n"); exit(-1); }
            slot = nextObs++;
        } else {
            if (nextUnobs >= p.capIndiv) { fprintf(stderr, "There are more individuals in your genoInfo.txt than what we can allocate\n"); exit(-1); }
            slot = nextUnobs++;
        }
        Indiv &x = p.indiv[slot];
        x.userID = tok_int(0);
        x.observed = obs;
        x.sex = tok_int(2); x.sexWasUnknown = x.sex == 0;
        x.genTime = tok_strtod(tb[3], te[3]);
        x.gen = (int)(x.genTime / p.minAge);
        if (x.gen >= p.nGen || x.gen < 0) {
            fprintf(stderr, "In genoInfo.txt, the assigned generation level of %d-th individual is %d, beyond nGen\n", row, x.gen); exit(0);
        }
        p.nIndivPerGen[x.gen]++;
        if (obs) p.nObsPerGen[x.gen]++;
        x.founder = tok_int(4) != 0;
        p.indivActive[slot] = 1;
        if (obs) {
            bool soft = false, dec_ok = true;
            int s = 0;
            for (; s < p.S && next_token(c, b, e); s++) {
                const char *comma = (const char *)memchr(b, ',', (size_t)(e - b));
                if (comma) {
                    soft = true;
                    // three non-empty parts separated by commas (empty parts are skipped, as split() does)
                    const char *pb[3], *pe[3];
                    int np = 0;
                    const char *q = b;
                    while (q < e) {
                        const char *r = (const char *)memchr(q, ',', (size_t)(e - q));
                        if (!r) r = e;
                        if (r > q) { if (np < 3) { pb[np] = q; pe[np] = r; } np++; }
                        q = r + 1;
                    }
                    if (np != 3) { fprintf(stderr, "Expect three genotype posterior prob on the geno.txt\n"); exit(1); }
                    for (int g = 0; g < 3; g++) {
                        unsigned nn = 0;
                        if (parse_dec4(pb[g], pe[g], nn)) {
                            lik[(size_t)s * 3 + g] = (double)nn / 10000.0;        // = strtod of the token: both are the correctly rounded value
                            if (dec_ok) num[(size_t)s * 3 + g] = (uint16_t)nn;
                        } else {
                            lik[(size_t)s * 3 + g] = tok_strtod(pb[g], pe[g]);
                            dec_ok = false;
                        }
                    }
                    codes[s] = 3;
                } else {
                    const int cde = (e - b == 1 && *b >= '0' && *b <= '9') ? *b - '0' : (int)tok_strtod(b, e);
                    if (cde < 0 || cde > 3) {
                        fprintf(stderr, "Acceptable genotype values are 0, 1, 2, or 3.\nIf you are uncertain about the genotype value, you can replace that value with 3.\n"); exit(-1);
                    }
                    codes[s] = (uint8_t)cde;
                    const double ee = p.eps[s];
                    for (int g = 0; g < 3; g++) lik[(size_t)s * 3 + g] = cde == 3 ? 1.0 : (g == cde ? 1.0 - ee : ee / 2.0);
                    dec_ok = false;
                }
            }
            if (s < p.S) {
                fprintf(stderr, "error on geno.txt: Individual %d has fewer SNPs than reported. \nnum of SNP - obs: %d , exp: %d", row, s, p.S); exit(-1);
            }
            int rc;
            if (!soft) rc = PFN(set_genotypes_hard)(p.eng, slot, codes.data());
            else if (dec_ok) rc = PFN(set_genotypes_soft_dec16)(p.eng, slot, num.data(), 10000);
            else rc = PFN(set_genotypes_soft_f64)(p.eng, slot, lik.data());   // mixed rows: exact doubles
            if (rc != PF_OK) die("genotype upload failed");
            if (g_n_threads > 1) {
                Upload u; u.slot = slot; u.kind = !soft ? 0 : dec_ok ? 1 : 2;
                if (u.kind == 0) u.codes = codes; else if (u.kind == 1) u.num = num; else u.lik = lik;
                g_uploads.push_back(std::move(u));
            }
        }
        row++;
    }
    fclose(f);
    if (nextObs < p.nObs) { fprintf(stderr, "There are fewer observed individuals in your genoInfo.txt than what you have declared\n"); exit(-1); }
    p.nIndivUsed = nextUnobs;
}

static void ReadMarriages(Ped &p, const std::string &path)
{
    FILE *f = fopen(path.c_str(), "r");
    if (!f) { printf("optional marriage.txt is not found; will build marriage from ground-up\n"); return; }
    std::vector<int> byUser;
    auto find_slot = [&](int uid) { for (int i = 0; i < p.nIndivUsed; i++) if (p.indivActive[i] && p.indiv[i].userID == uid) return i; return -1; };
    std::string line;
    std::vector<std::pair<std::pair<int, int>, int>> couples;   // (pa, ma) -> marriage slot
    while (read_line(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        auto tok = split(line, ' ');
        if (tok.size() < 3) continue;
        const int kid = find_slot(atoi(tok[0].c_str())), pa = find_slot(atoi(tok[1].c_str())), ma = find_slot(atoi(tok[2].c_str()));
        if (kid < 0) { fprintf(stderr, "The genotype information for this kid's ID %s from marriage.txt cannot be found\n", tok[0].c_str()); exit(-1); }
        if (pa < 0) { fprintf(stderr, "The genotype information for the parent 0's ID %s from marriage.txt cannot be found\n", tok[1].c_str()); exit(-1); }
        if (ma < 0) { fprintf(stderr, "The genotype information for the parent 1's ID %s from marriage.txt cannot be found\n", tok[2].c_str()); exit(-1); }
        int m = -1;
        for (auto &c : couples) if (c.first.first == pa && c.first.second == ma) m = c.second;
        if (m < 0) {
            m = GetEmptyMarrIndx(p);
            Marr &mm = p.marr[m];
            mm = Marr(); mm.pa = pa; mm.ma = ma; mm.selfing = pa == ma;
            p.marrActive[m] = 1;
            p.indiv[pa].M.push_back(m); p.indiv[pa].edge.push_back(0);
            if (pa != ma) { p.indiv[ma].M.push_back(m); p.indiv[ma].edge.push_back(1); }
            couples.push_back({{pa, ma}, m});
        }
        Marr &mm = p.marr[m];
        mm.kids.push_back(kid);
        p.indiv[kid].M.push_back(m); p.indiv[kid].edge.push_back(mm.nEdges() - 1);
    }
    fclose(f);
    if (p.nMarDecl != 0 && (int)couples.size() != p.nMarDecl) {
        fprintf(stderr, "The number of marriage observed in marriage.txt doesn't match up with nMar in prior.txt\n"); exit(-1);
    }
}

// _CheckPed@0x100012e10: label components; all start dirty. Loops abort (acyclic model).
static void CheckPed(Ped &p)
{
    p.nCC = 0; p.nDirty = 0;
    p.ccMem.assign(p.capIndiv + 2, {});
    for (int i = 0; i < p.nIndivUsed; i++) {
        if (!p.indivActive[i] || p.cc[i] != -1) continue;
        label_component(p, i, p.nCC);
        p.ccDirty[p.nCC] = 1; p.nDirty++;
        p.nCC++;
    }
    // forest test: edges = nodes - components
    long edges = 0, nodes = 0;
    for (int i = 0; i < p.nIndivUsed; i++) nodes += p.indivActive[i] ? 1 : 0;
    for (int m = 0; m < p.nMarrUsed; m++) if (p.marrActive[m]) { nodes++; edges += p.marr[m].nEdges() - (p.marr[m].selfing ? 1 : 0); }
    if (edges != nodes - p.nCC) { fprintf(stderr, "Initial pedigree configuration contains loops."); exit(-1); }
}

// _RefineParentOffPair@0x100014d20: top-15 single-parent candidates per observed kid, scored on
// the engine in one dense scan per group of kids sharing a generation time
static void RefineParentOffPair(Ped &p)
{
    std::vector<int> order(p.nObs);
    for (int i = 0; i < p.nObs; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return p.indiv[a].genTime < p.indiv[b].genTime; });
    size_t i0 = 0;
    while (i0 < order.size()) {
        size_t i1 = i0;
        while (i1 < order.size() && p.indiv[order[i1]].genTime == p.indiv[order[i0]].genTime) i1++;
        const double kt = p.indiv[order[i0]].genTime;
        std::vector<int> kids(order.begin() + i0, order.begin() + i1), cands;
        std::sort(kids.begin(), kids.end());
        std::vector<double> bias;
        for (int j = 0; j < p.nObs; j++) {
            const double gap = p.indiv[j].genTime - kt;
            if (gap > p.maxAge || p.minAge > gap) continue;
            cands.push_back(j); bias.push_back(-p.LLcc[p.cc[j]]);
        }
        if (!cands.empty()) {
            // _CalculateLLByStubs@0x10001eef0 subtracts LLcc[cc(candidate)] only when the candidate is
            // not in the kid's own component. From the all-singletons start that is every pair, and the
            // whole generation goes down in one call with a per-candidate bias; a kid that shares its
            // component with a candidate (possible only with a seeded marriage.txt) gets a call of
            // its own with that candidate's bias cleared.
            const int K = 15;
            auto run = [&](const std::vector<int> &ks, const std::vector<double> &bs) {
                std::vector<uint8_t> role(ks.size());
                for (size_t k = 0; k < ks.size(); k++) role[k] = p.indiv[ks[k]].sex == 1 ? 1 : 0;
                std::vector<int> slot(ks.size() * K);
                std::vector<double> score(ks.size() * K);
                if (PFN(prefilter)(p.eng, p.chain, (int)ks.size(), ks.data(), role.data(), (int)cands.size(), cands.data(),
                                   bs.data(), K, slot.data(), score.data()) != PF_OK) die("prefilter failed");
                for (size_t k = 0; k < ks.size(); k++) {
                    auto &pc = p.parentCand[ks[k]];
                    pc.clear();
                    for (int t = 0; t < K; t++) if (slot[k * K + t] >= 0) pc.push_back(slot[k * K + t]);
                    if (getenv("PEDFAC_DEBUG_CAND")) {
                        fprintf(stderr, "cand kid %d:", ks[k]);
                        for (int t = 0; t < K; t++) if (slot[k * K + t] >= 0) fprintf(stderr, " %d(%.17g)", slot[k * K + t], score[k * K + t]);
                        fprintf(stderr, "\n");
                    }
                }
            };
            std::vector<int> plain;
            for (int kd : kids) {
                bool shares = false;
                for (int j : cands) shares = shares || p.cc[j] == p.cc[kd];
                if (!shares) { plain.push_back(kd); continue; }
                std::vector<double> own(bias);
                for (size_t c = 0; c < cands.size(); c++) if (p.cc[cands[c]] == p.cc[kd]) own[c] = 0.0;
                run({kd}, own);
            }
            if (!plain.empty()) run(plain, bias);
        }
        i0 = i1;
    }
}

// ---------------------------------------------------------------------------------------------
// CLI (_GetCommLineOpts@0x10000f450: the seven functional flags of the ECA_Opt3 parser)

static void usage()
{
    fprintf(stderr, "pedigraph -d/--directory-info DIR [-r/--random-file FILE] [-n/--num-iter N] [-c/--cyclic-choice J] [-s/--selfing] [-w/--diswap] [-l/--LBP]\n");
}

static Opts GetCommLineOpts(int argc, char **argv)
{
    Opts o;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&]() -> const char * { if (i + 1 >= argc) { usage(); exit(1); } return argv[++i]; };
        if (a == "-d" || a == "--directory-info") o.dir = need();
        else if (a == "-r" || a == "--random-file") o.seedfile = need();
        else if (a == "-n" || a == "--num-iter") o.nIter = atoi(need());
        else if (a == "-c" || a == "--cyclic-choice") o.cyclic = atoi(need());
        else if (a == "-s" || a == "--selfing") o.selfing = true;
        else if (a == "-w" || a == "--diswap") o.diswap = true;
        else if (a == "-l" || a == "--LBP") o.loopy = true;
        else if (a == "--help" || a == "--help-full") { usage(); exit(0); }
        else if (a == "--version") { printf("pedigraph (pedfac-b200)\n"); exit(0); }
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); exit(1); }
    }
    if (o.dir.empty()) { usage(); exit(1); }
    // -c 1 (throttle): observed on the reference binary, its trace equals the -c 0 trace
    // until a loop-closing throttle proposal is picked, at which point its exact propagation aborts
    // ("propagation step exceeds the acyclic upper limit"). This sampler runs -c 1 as -c 0: the same
    // trace, without the abort. -l is handled where the reference aborts under it (main).
    if (o.cyclic != 0) fprintf(stderr, "pedigraph: -c %d runs as -c 0 (loop-closing proposals are dropped; the reference aborts once it picks one)\n", o.cyclic);
    return o;
}

static void SampleLoop(Ped &p, const Opts &o)
{
    printf("sample for %d iterations: -- \n", o.nIter);
    // measurement aids (not part of the reference contract): PEDFAC_STATS=1 prints per-iteration
    // counters and wall time to stderr; PEDFAC_MAX_STEPS=K stops after K Gibbs steps in total
    const bool stats = getenv("PEDFAC_STATS") != nullptr;
    p.check = getenv("PEDFAC_CHECK") != nullptr;             // PEDFAC_CHECK=1: re-propagate after every move and compare with its score
    const long max_steps = getenv("PEDFAC_MAX_STEPS") ? atol(getenv("PEDFAC_MAX_STEPS")) : -1;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    for (p.iter = 0; p.iter < o.nIter; p.iter++) {             // _main@0x10001a812-0x10001aaea
        const double t_iter = now();
        const long props0 = p.n_proposals, sweeps0 = p.n_sweeps, steps0 = p.n_steps, swaps0 = p.n_swaps, picked0 = p.n_swaps_picked;
        uint64_t io0[2] = {0, 0};
        PFN(io_bytes)(p.eng, &io0[0], &io0[1]);
        p.unobsQueue.clear();
        int qBeg = 0, qEnd = -1;
        for (int k = 0; k < p.nObs; k++) {
            const int id = p.observedSorted[k];
            if (max_steps >= 0 && p.n_steps >= max_steps) break;
            MCMC_sampling(p, o, id);
            const bool last = k == p.nObs - 1;
            if (!last && !(p.indiv[id].gen < p.indiv[p.observedSorted[k + 1]].gen)) continue;
            const int gap = last ? p.nGen - p.indiv[id].gen : p.indiv[p.observedSorted[k + 1]].gen - p.indiv[id].gen;
            for (int g = 1; g <= gap; g++) {
                if (getenv("PEDFAC_DEBUG_CALLS")) fprintf(stderr, "pass g %d of %d window %d..%d queue %zu\n", g, gap, qBeg, qEnd, p.unobsQueue.size());
                for (int q = qBeg; q < qEnd + 1; q++) {
                    const int u = p.unobsQueue[q];
                    if (!p.indiv[u].queued) continue;
                    MCMC_sampling(p, o, u);
                    p.indiv[u].queued = false;
                }
                if ((int)p.unobsQueue.size() - qEnd > 1) { qBeg = qEnd + 1; qEnd = (int)p.unobsQueue.size() - 1; }
            }
        }
        printf("..%i\n", p.iter);
        fflush(stdout);
        if (stats)
        {
            uint64_t h2d = 0, d2h = 0;
            PFN(io_bytes)(p.eng, &h2d, &d2h);
            fprintf(stderr, "pedigraph-stats iter %d steps %ld proposals %ld sweeps %ld swaps %ld swaps_picked %ld seconds %.6f h2d_bytes %llu d2h_bytes %llu chain %d\n", p.iter,
                    p.n_steps - steps0, p.n_proposals - props0, p.n_sweeps - sweeps0, p.n_swaps - swaps0, p.n_swaps_picked - picked0, now() - t_iter,
                    (unsigned long long)(h2d - io0[0]), (unsigned long long)(d2h - io0[1]), p.chain_id);
            if (p.chain_id == 0) fprintf(stderr, "pedigraph-time iter %d sweep_call %.6f eval_call %.6f view_build %.6f prune %.6f moves %.6f pick %.6f\n", p.iter, p.t_sweep, p.t_eval, p.t_view, p.t_prune, p.t_moves, p.t_pick);
            p.t_sweep = p.t_eval = p.t_view = p.t_prune = p.t_moves = p.t_pick = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PEDFAC_CHAINS: the scheduler of the chain fibers (see namespace mc above)
#ifndef PF_NO_COMM
static std::vector<Ped> *g_peds = nullptr;
static const Opts *g_opts = nullptr;

static void fiber_entry(int c)                     // c: global chain number
{
    SampleLoop((*g_peds)[c], *g_opts);
    mc::Fiber &f = mc::fibers[c - mc::base];
    f.done = true;
    f.req.kind = mc::NONE;
    swapcontext(&f.ctx, &mc::sched_ctx);
}

static bool open_outputs(Ped &p, const std::string &out)
{
    mkdir(out.c_str(), 0777);
    p.fEdge = fopen((out + "/historyEdge.csv").c_str(), "w");
    p.fNode = fopen((out + "/historyNode.csv").c_str(), "w");
    p.fPed = fopen((out + "/ped.txt").c_str(), "w");
    p.fGeno = fopen((out + "/geno.txt").c_str(), "w");
    p.fLL = fopen((out + "/ll.txt").c_str(), "w");
    return p.fEdge && p.fNode && p.fPed && p.fGeno && p.fLL;
}

static void copy_file_so_far(FILE *from, const std::string &path_from, FILE *to)
{
    fflush(from);
    if (FILE *f = fopen(path_from.c_str(), "r")) {
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) fwrite(buf, 1, n, to);
        fclose(f);
    }
}

// One scheduler: runs the chains c0 .. c1-1 (fibers of this thread) to the end on `eng`, batching the
// requests of a round into one engine submission per kind.
static void RunGroup(PF_ENGINE_T *eng, int c0, int c1, long seed_a, long seed_b)
{
    const int C = c1 - c0;
    mc::base = c0;
    mc::fibers.clear();
    mc::fibers.resize(C);
    const size_t stack_bytes = (size_t)16 << 20;
    for (int k = 0; k < C; k++) {
        mc::Fiber &f = mc::fibers[k];
        f.chain = c0 + k;
        f.stack = mmap(nullptr, stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (f.stack == MAP_FAILED) { perror("mmap"); exit(-1); }
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = stack_bytes; f.ctx.uc_link = &mc::sched_ctx;
        makecontext(&f.ctx, (void (*)())fiber_entry, 1, c0 + k);
        rng::setall(seed_a + c0 + k, seed_b + c0 + k);
        f.s1 = rng::s1; f.s2 = rng::s2;
    }
    mc::active = true;
    std::vector<int> idx, chains, kids, pb, sb;
    std::vector<pf_swap> swaps;
    std::vector<pf_graph_view> views;
    std::vector<pf_proposal> props;
    std::vector<double> ll, lsum;
    for (;;) {
        int alive = 0;
        for (int k = 0; k < C; k++) {
            mc::Fiber &f = mc::fibers[k];
            if (f.done) continue;
            mc::current = k; g_quiet = c0 + k != 0;
            rng::s1 = f.s1; rng::s2 = f.s2;
            swapcontext(&mc::sched_ctx, &f.ctx);
            f.s1 = rng::s1; f.s2 = rng::s2;
            if (!f.done) alive++;
        }
        g_quiet = c0 != 0;
        if (alive == 0) break;
        // the sweeps of this round (SWEEP and the sweep half of SWEEP_EVAL) as one batch
        idx.clear(); chains.clear(); views.clear();
        size_t n_ll = 0;
        for (int k = 0; k < C; k++) {
            mc::Req &r = mc::fibers[k].req;
            if (mc::fibers[k].done || (r.kind != mc::SWEEP && r.kind != mc::SWEEP_EVAL)) continue;
            idx.push_back(k); chains.push_back(k); views.push_back(*r.view); n_ll += (size_t)std::max(r.view->n_cc, 0);
        }
        if (!idx.empty()) {
            ll.assign(n_ll + 1, 0.0);
            const int rc = PFN(sweep_batch)(eng, (int)idx.size(), chains.data(), views.data(), ll.data());
            size_t at = 0;
            for (int k : idx) {
                mc::Req &r = mc::fibers[k].req;
                r.rc = rc;
                for (int j = 0; j < r.view->n_cc; j++) r.ll[j] = ll[at + j];
                at += (size_t)std::max(r.view->n_cc, 0);
            }
        }
        // the proposal lists of this round (EVAL and the eval half of SWEEP_EVAL) as one batch
        idx.clear(); chains.clear(); kids.clear(); pb.assign(1, 0); props.clear();
        for (int k = 0; k < C; k++) {
            mc::Req &r = mc::fibers[k].req;
            if (mc::fibers[k].done || (r.kind != mc::EVAL && r.kind != mc::SWEEP_EVAL) || r.n <= 0) continue;
            idx.push_back(k); chains.push_back(k); kids.push_back(r.kid);
            props.insert(props.end(), r.props, r.props + r.n);
            pb.push_back((int)props.size());
        }
        if (!idx.empty()) {
            lsum.assign(props.size() + 1, 0.0);
            const int rc = PFN(eval_batch)(eng, (int)idx.size(), chains.data(), kids.data(), pb.data(), props.data(), lsum.data());
            for (size_t q = 0; q < idx.size(); q++) {
                mc::Req &r = mc::fibers[idx[q]].req;
                if (rc != PF_OK) r.rc = rc;
                for (int j = 0; j < r.n; j++) r.lsum[j] = lsum[(size_t)pb[q] + j];
            }
        }
        // the swap blocks of this round as one batch
        idx.clear(); chains.clear(); views.clear(); swaps.clear(); sb.assign(1, 0);
        for (int k = 0; k < C; k++) {
            mc::Req &r = mc::fibers[k].req;
            if (mc::fibers[k].done || r.kind != mc::SWAPS) continue;
            idx.push_back(k); chains.push_back(k); views.push_back(*r.view);
            swaps.insert(swaps.end(), r.swaps, r.swaps + r.n);
            sb.push_back((int)swaps.size());
        }
        if (!idx.empty()) {
            lsum.assign(swaps.size() + 1, 0.0);
            const int rc = PFN(eval_swaps_batch)(eng, (int)idx.size(), chains.data(), views.data(), sb.data(), swaps.data(), lsum.data());
            for (size_t q = 0; q < idx.size(); q++) {
                mc::Req &r = mc::fibers[idx[q]].req;
                r.rc = rc;
                for (int j = 0; j < r.n; j++) r.lsum[j] = lsum[(size_t)sb[q] + j];
            }
        }
        for (int k = 0; k < C; k++)
            if (!mc::fibers[k].done && mc::fibers[k].req.kind == mc::EVAL && mc::fibers[k].req.n <= 0) mc::fibers[k].req.rc = PF_OK;
    }
    mc::active = false;
    for (int k = 0; k < C; k++) munmap(mc::fibers[k].stack, stack_bytes);
    mc::fibers.clear();
}

// Runs all chains to the end. `p0` is chain 0, set up to the point where sampling starts (files read,
// initial sweep and prefilter done); the other chains start as copies of it with their own row block in
// an engine (everything dirty, so their first Gibbs step propagates it), RNG stream and output files
// under out/chain<c>/. PEDFAC_THREADS=T deals the chains to T scheduler threads, each with an engine of
// its own on the same GPU (the genotype tables are uploaded to each): host chain logic of one group
// overlaps the kernels of the others.
static void RunChains(Ped &p0, const Opts &o, long seed_a, long seed_b)
{
    const int C = g_n_chains, T = g_n_threads;
    std::vector<PF_ENGINE_T *> eng(T, nullptr);
    eng[0] = p0.eng;
    for (int t = 1; t < T; t++) {
        pf_config cfg = g_cfg;
        cfg.n_chains = chains_of_thread(t);
        if (PFN(create)(&eng[t], &cfg) != PF_OK) die("cannot create the likelihood engine of a scheduler thread");
        if (PFN(set_locus_params)(eng[t], p0.eps.data(), p0.afreq.data()) != PF_OK) die("set_locus_params");
        for (const Upload &u : g_uploads) {
            const int rc = u.kind == 0 ? PFN(set_genotypes_hard)(eng[t], u.slot, u.codes.data())
                         : u.kind == 1 ? PFN(set_genotypes_soft_dec16)(eng[t], u.slot, u.num.data(), 10000)
                                       : PFN(set_genotypes_soft_f64)(eng[t], u.slot, u.lik.data());
            if (rc != PF_OK) die("genotype upload failed");
        }
    }
    g_uploads.clear(); g_uploads.shrink_to_fit();
    std::vector<Ped> peds(C);
    const std::string out0 = o.dir + "/out";
    std::vector<int> first(T + 1, 0);
    for (int t = 0; t < T; t++) first[t + 1] = first[t] + chains_of_thread(t);
    for (int t = 0; t < T; t++)
        for (int c = first[t]; c < first[t + 1]; c++) {
            peds[c] = p0;
            Ped &q = peds[c];
            q.chain = c - first[t]; q.chain_id = c; q.eng = eng[t];
            if (c == 0) continue;
            if (!open_outputs(q, out0 + "/chain" + std::to_string(c))) { fprintf(stderr, "cannot open output files of chain %d\n", c); exit(-1); }
            copy_file_so_far(p0.fNode, out0 + "/historyNode.csv", q.fNode);
            copy_file_so_far(p0.fEdge, out0 + "/historyEdge.csv", q.fEdge);
            for (int k = 0; k < q.nCC; k++) mark_dirty(q, k);
        }
    g_peds = &peds; g_opts = &o;
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back([&, t]() { RunGroup(eng[t], first[t], first[t + 1], seed_a, seed_b); });
    RunGroup(eng[0], first[0], first[1], seed_a, seed_b);
    for (auto &x : th) x.join();
    for (int c = 1; c < C; c++) { Ped &q = peds[c]; fclose(q.fEdge); fclose(q.fNode); fclose(q.fPed); fclose(q.fGeno); fclose(q.fLL); }
    for (int t = 1; t < T; t++) PFN(destroy)(eng[t]);
    p0 = peds[0];
}
#endif

int main(int argc, char **argv)
{
    Opts o = GetCommLineOpts(argc, argv);
    if (getenv("PEDFAC_SHARD_WORLD")) {                        // a child of the launcher below
        g_shard_world = atoi(getenv("PEDFAC_SHARD_WORLD"));
        g_shard_rank = getenv("PEDFAC_SHARD_RANK") ? atoi(getenv("PEDFAC_SHARD_RANK")) : 0;
    } else if (getenv("PEDFAC_GPUS") && atoi(getenv("PEDFAC_GPUS")) > 1) {
        const int n = atoi(getenv("PEDFAC_GPUS"));
        mkdir((o.dir + "/out").c_str(), 0777);
        std::vector<pid_t> kids;
        const std::string nonce = std::to_string((long)getpid());
        for (int r = 0; r < n; r++) {
            const pid_t c = fork();
            if (c < 0) { perror("fork"); return -1; }
            if (c == 0) {
                setenv("PEDFAC_SHARD_WORLD", std::to_string(n).c_str(), 1);
                setenv("PEDFAC_SHARD_RANK", std::to_string(r).c_str(), 1);
                setenv("PEDFAC_SHARD_NONCE", nonce.c_str(), 1);
                setenv("PEDFAC_DEVICE", std::to_string(r).c_str(), 1);
                execv("/proc/self/exe", argv);
                perror("execv"); _exit(127);
            }
            kids.push_back(c);
        }
        int rc = 0;
        for (int r = 0; r < n; r++) {
            int st = 0;
            waitpid(kids[r], &st, 0);
            const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
            if (code != 0 && rc == 0) rc = code;
        }
        for (int r = 0; r < n; r++) unlink((o.dir + "/out/.pf_comm_" + nonce + "_" + std::to_string(r)).c_str());
        return rc;
    }
    if (getenv("PEDFAC_CHAINS") && atoi(getenv("PEDFAC_CHAINS")) > 1) g_n_chains = atoi(getenv("PEDFAC_CHAINS"));
    if (getenv("PEDFAC_THREADS")) g_n_threads = std::max(1, std::min(atoi(getenv("PEDFAC_THREADS")), g_n_chains));
    g_shard_dir = o.dir + "/out";
    if (g_shard_rank > 0 && !freopen("/dev/null", "w", stdout)) return -1;
    Ped p;
    long seed_a = 0, seed_b = 0;
    {   // _SeedFromFile@0x100027070
        long a = 0, b = 0;
        FILE *f = fopen(o.seedfile.c_str(), "r");
        if (!f) {
            a = (long)time(nullptr); b = a / 3;
            printf("SEEDFILE: No file \"%s\".  Using clock to set RNG: %ld %ld\n", o.seedfile.c_str(), a, b);
        } else {
            if (fscanf(f, "%ld %ld", &a, &b) != 2) { a = 1234567890L; b = 123456789L; }
            printf("SEEDFILE: Using contents of %s to set RNG: %ld %ld\n", o.seedfile.c_str(), a, b);
            fclose(f);
        }
        rng::setall(a, b);
        seed_a = a; seed_b = b;
    }
    {   // _PrepOutFiles@0x10001a3a0
        std::string out = o.dir + "/out";
        mkdir(out.c_str(), 0777);
        if (g_shard_rank > 0) { out += "/.shard" + std::to_string(g_shard_rank); mkdir(out.c_str(), 0777); }
        p.fEdge = fopen((out + "/historyEdge.csv").c_str(), "w");
        p.fNode = fopen((out + "/historyNode.csv").c_str(), "w");
        p.fPed = fopen((out + "/ped.txt").c_str(), "w");
        p.fGeno = fopen((out + "/geno.txt").c_str(), "w");
        p.fLL = fopen((out + "/ll.txt").c_str(), "w");
        if (!p.fEdge || !p.fNode || !p.fPed || !p.fGeno || !p.fLL) { fprintf(stderr, "cannot open output files under %s\n", out.c_str()); return -1; }
    }
    ReadPrior(p, o.dir + "/prior.txt");
    ReadGeno(p, o.dir + "/geno.txt");
    ReadMarriages(p, o.dir + "/marriage.txt");
    printf("\t\tcompleted allocation!\n");
    CheckPed(p);
    printf("Allow selfing: %d\n", o.selfing ? 1 : 0);
    printf("Allow loopy: %d\n", o.loopy ? 1 : 0);
    printf("propagating message\n");
    if (o.loopy) {
        // -l: the reference's loopy schedule never reports completion (_PassingIFtoIV@0x10001c8f0 returns
        // "all done" only when not loopy), so its very first propagation runs into the round limit and
        // exits — observed on every input. Same messages, same streams, same status.
        printf("propagation step exceeds the acyclic upper limit");
        fflush(stdout);
        fprintf(stderr, "exceeding the upper limits of sum-product recursion counter: %d\n", p.nIndivUsed + 2);
        exit(-1);
    }
    Propagating_msg(p);
    printf("prior: %f \n", p.LLtot);
    printf("finish propagating message\n");
    {   // _PrintInitalHistory@0x100012600
        fprintf(p.fNode, "t,action,id,label,level,group\n");
        fprintf(p.fEdge, "t,action,id,from,to\n");
        for (int i = 0; i < p.nIndivUsed; i++) if (p.indivActive[i]) hist_node(p, 1, i);
        for (int m = 0; m < p.nMarrUsed; m++) {
            if (!p.marrActive[m]) continue;
            hist_mnode(p, 1, m);
            hist_edge(p, 1, m, p.marr[m].pa);
            if (!p.marr[m].selfing) hist_edge(p, 1, m, p.marr[m].ma);
            for (int k : p.marr[m].kids) hist_edge(p, 1, m, k);
        }
    }
    for (int i = 0; i < p.nIndivUsed; i++) UpdateUnobsDegree(p, i);
    RefineParentOffPair(p);
    p.observedSorted.resize(p.nObs);
    for (int i = 0; i < p.nObs; i++) p.observedSorted[i] = i;
    std::stable_sort(p.observedSorted.begin(), p.observedSorted.end(), [&](int a, int b) { return p.indiv[a].gen < p.indiv[b].gen; });   // qsort(_CompareGen)
#ifndef PF_NO_COMM
    if (g_n_chains > 1) RunChains(p, o, seed_a, seed_b);
    else
#else
    if (g_n_chains > 1) { fprintf(stderr, "PEDFAC_CHAINS needs the CUDA build (batched engine calls)\n"); return -1; }
#endif
    SampleLoop(p, o);
    printf(".. DONE \n");
    fclose(p.fEdge); fclose(p.fNode); fclose(p.fPed); fclose(p.fGeno); fclose(p.fLL);
    PFN(destroy)(p.eng);
    return 0;
}
