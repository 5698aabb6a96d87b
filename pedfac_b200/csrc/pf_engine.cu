// pf_engine.cu — host side of libpedfac_cuda: device storage, genotype packing, the sweep plan
// builder (graph view -> level-ordered task list) and the C ABI of include/pedfac_cuda.h.
// There is no CPU fallback anywhere in this file: every entry point either runs the sm_100a
// kernels of pf_kernels.cu or returns an error.
#include "../../include/pedfac_cuda.h"
#include "pf_kernels.cuh"
#include "pf_prefilter.cuh"
#include "pf_sweep2.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

using namespace pf;

static thread_local char g_err[1024] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(PF_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#ifndef PF_EVAL_MIN_CHUNK
#define PF_EVAL_MIN_CHUNK 128
#endif
namespace {

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 2 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// ring of pinned staging buffers; a slot is reused only after the copy that read it completed
struct Staging {
    static constexpr int N = 4;
    void *p[N] = {};
    size_t cap[N] = {};
    cudaEvent_t ev[N] = {};
    bool used[N] = {};
    int next = 0;
    cudaError_t get(size_t bytes, void **out, int *slot)
    {
        int i = next; next = (next + 1) % N;
        if (used[i]) { cudaError_t e = cudaEventSynchronize(ev[i]); if (e != cudaSuccess) return e; used[i] = false; }
        if (!ev[i]) { cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
        if (bytes > cap[i]) {
            if (p[i]) cudaFreeHost(p[i]);
            p[i] = nullptr; cap[i] = 0;
            size_t want = bytes + bytes / 2 + 4096;
            cudaError_t e = cudaMallocHost(&p[i], want);
            if (e != cudaSuccess) return e;
            cap[i] = want;
        }
        *out = p[i]; *slot = i;
        return cudaSuccess;
    }
    cudaError_t mark(int slot, cudaStream_t st) { used[slot] = true; return cudaEventRecord(ev[slot], st); }
    void release()
    {
        for (int i = 0; i < N; i++) { if (p[i]) cudaFreeHost(p[i]); if (ev[i]) cudaEventDestroy(ev[i]); p[i] = nullptr; ev[i] = nullptr; }
    }
};

struct SoftChunk { char *base; size_t cap, used; };

} // namespace

struct pf_engine {
    // PF_HOST_PROF=1: seconds spent building sweep plans / staging+launching / waiting for results
    int force_shape = 0;              // PF_SWEEP_SHAPE=latency|throughput overrides the per-launch choice (measurement aid)
    bool sweep_clocks = false;        // PF_SWEEP_CLOCKS=1: per-level clocks of the largest group after every sweep
    bool host_prof = false; double t_plan = 0, t_submit = 0, t_wait = 0, t_evalprep = 0; long n_prof = 0, n_prof_calls = 0; double t_ph[6] = {0, 0, 0, 0, 0, 0};
    bool sweep_v2 = true;             // PF_SWEEP_V2=0: first-generation sweep kernel only (measurement aid)
    bool eval_singles = true;         // PF_EVAL_SINGLES=0: single-parent proposals through the general path (measurement aid)
    long sweeps_v2 = 0, sweeps_v1_fallback = 0;
    unsigned long long io_h2d = 0, io_d2h = 0;       // bytes copied per call path: plans / proposals up, results down (pf_io_bytes)
    bool no_mailbox = false;          // PEDFAC_NO_MAILBOX=1 or no mapped memory: results by copy + stream synchronise
    int sm_count = 0;
    int eval_capacity = 0;            // resident eval_kernel blocks on this device
    pf_config cfg;
    int Sl = 0, Sp = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    double *d_rows = nullptr; size_t n_rows = 0;
    double *d_eps = nullptr;
    uint32_t *d_hard = nullptr;           // [cap_indiv][Sp/16]
    uint8_t *d_kind = nullptr; void **d_eptr = nullptr; uint32_t *d_denom = nullptr;
    std::vector<uint8_t> h_kind; std::vector<void *> h_eptr; std::vector<uint32_t> h_denom;
    std::vector<double> h_rcp;            // 1 / h_denom per slot (the sweep plans carry it)
    std::vector<uint8_t> soft_kind_of_row;  // kind the slot's soft row was allocated for
    std::vector<void *> soft_rows_by_kind;  // [slot][kind]: rows already allocated for a slot, one per format
    bool tables_dirty = true;
    std::vector<SoftChunk> soft_chunks;
    std::vector<double> h_eps;            // local shard, for validation only

    DevBuf scratch, plan, partial, result, counters, clockbuf, sweep_counter;
    Staging staging;
    double *h_result = nullptr; size_t h_result_cap = 0;
    // mapped pinned mailbox for small results: [0] = sequence flag, [8..] = values (see publish_kernel)
    static constexpr int MAILBOX_MAX = 4096;
    unsigned long long *h_mail = nullptr, *d_mail = nullptr, mail_seq = 0;
    // results of swap proposals queued by pf_sweep_eval_swaps, delivered with the next host-result call
    DevBuf deferred; int n_deferred = 0, pub_armed_def = 0; std::vector<double> h_deferred;
    cudaEvent_t fetch_event = nullptr;
    int pub_want = 0;                 // set by a host-buffer entry point: doubles of the result vector the eval kernel should deliver itself
    unsigned long long pub_armed = 0; // sequence number the eval kernel was armed with (0: fetch_result launches publish_kernel)
    bool eval_zero_copy = false;      // PF_EVAL_ZEROCOPY=1 (measurement aid): the kernel reads small eval plans from the pinned buffer. Off:
                                      // in-kernel PCIe reads made the kernel 6 us longer on one box and 28 us longer on another
    static constexpr size_t EVAL_ZERO_COPY_MAX = 16384;
    int eval_min_chunk = PF_EVAL_MIN_CHUNK;   // PF_EVAL_MIN_CHUNK (measurement aid): fewest loci per proposal-kernel chunk
    bool swap_sites = true;           // PF_SWAP_SITES=0: every swap proposal as a hypothetical component of its own
    int inline_kids = FAM2_INLINE_KIDS, chunk_kids = CHUNK2_KIDS;   // PF_INLINE_KIDS / PF_CHUNK_KIDS (measurement aid): how sibships are cut into tasks
    DevBuf pubctr; bool pubctr_stale = true;
    int eval_small_group = 4, eval_small_props = 1024;   // PF_EVAL_GROUP / PF_EVAL_SMALL_PROPS
    bool eval_inline = true;          // PF_EVAL_INLINE=0: always copy the eval plan to device memory
    bool fused_publish = true;        // PF_FUSED_PUBLISH=0: always a separate publish launch

    // locus shards on several GPUs (pf_comm_*): peer-mapped mailboxes for the cross-GPU sum (xreduce_kernel)
    struct Comm {
        bool on = false;
        int rank = 0, world = 1;
        void *local = nullptr;                       // this rank's mailbox allocation: flags, then data
        void *peer_base[XREDUCE_MAX_WORLD] = {};     // every rank's allocation as mapped here (own: local)
        unsigned long long seq = 1;
        long n_reduce = 0;
        bool async = false;                          // pf_comm_set_async: sums of the _dev calls run on `stream2`
        cudaStream_t stream2 = nullptr;
        cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
        bool pending = false;                        // a sum was issued on stream2 since the last fence
    } comm;

    std::vector<int> loc;                 // slot -> index in the view being planned
    int64_t launches = 0;
    int32_t last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // see pf_plan_stats

    // optional per-launch CUDA-event timing (pf_profile): kind 0 sweep, 1 eval, 2 reduce, 3 prefilter
    bool prof_on = false;
    struct ProfRec { int kind; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    void prof_begin(int kind)
    {
        if (!prof_on) return;
        ProfRec r{kind, nullptr, nullptr};
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream);
        prof.push_back(r);
    }
    void prof_end() { if (prof_on && !prof.empty()) cudaEventRecord(prof.back().b, stream); }

    size_t row_elems() const { return (size_t)3 * Sp; }
    int stub_row(int chain, int slot) const { return ROW_CHAIN0 + chain * (cfg.cap_indiv + cfg.cap_marr) + slot; }
    int prong_row(int chain, int m) const { return ROW_CHAIN0 + chain * (cfg.cap_indiv + cfg.cap_marr) + cfg.cap_indiv + m; }
};

// ---------------------------------------------------------------------------------------------

extern "C" const char *pf_last_error(void) { return g_err; }

static double host_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

extern "C" int pf_destroy(pf_engine *e)
{
    if (e && e->host_prof && e->n_prof)
        fprintf(stderr, "pf-host-prof plan phases us: adjacency %.1f traversals %.1f heights %.1f rows+items+genotype-entries %.1f emission %.1f\n",
                1e6 * e->t_ph[0] / e->n_prof, 1e6 * e->t_ph[1] / e->n_prof, 1e6 * e->t_ph[2] / e->n_prof, 1e6 * e->t_ph[3] / e->n_prof, 1e6 * e->t_ph[4] / e->n_prof),
        fprintf(stderr, "pf-host-prof sweeps %ld plan_us %.1f submit_us %.1f evalprep_us %.1f wait_us %.1f (per sweep)\n", e->n_prof,
                1e6 * e->t_plan / e->n_prof, 1e6 * e->t_submit / e->n_prof, 1e6 * e->t_evalprep / e->n_prof, 1e6 * e->t_wait / e->n_prof);
    if (!e) return PF_OK;
    if (getenv("PF_SWEEP_STATS")) fprintf(stderr, "pf-sweep-stats v2 %ld v1-fallback %ld\n", e->sweeps_v2, e->sweeps_v1_fallback);
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_rows); cudaFree(e->d_eps); cudaFree(e->d_hard);
    cudaFree(e->d_kind); cudaFree(e->d_eptr); cudaFree(e->d_denom);
    for (auto &c : e->soft_chunks) cudaFree(c.base);
    e->scratch.release(); e->plan.release(); e->partial.release(); e->result.release(); e->counters.release(); e->sweep_counter.release();
    e->clockbuf.release(); e->pubctr.release(); e->deferred.release();
    if (e->fetch_event) cudaEventDestroy(e->fetch_event);
    for (auto &r : e->prof) { if (r.a) cudaEventDestroy(r.a); if (r.b) cudaEventDestroy(r.b); }
    e->prof.clear();
    e->staging.release();
    if (e->h_result) cudaFreeHost(e->h_result);
    if (e->h_mail) cudaFreeHost(e->h_mail);
    for (int r = 0; r < e->comm.world && e->comm.on; r++)
        if (r != e->comm.rank && e->comm.peer_base[r]) cudaIpcCloseMemHandle(e->comm.peer_base[r]);
    if (e->comm.local) cudaFree(e->comm.local);
    if (e->comm.stream2) cudaStreamDestroy(e->comm.stream2);
    if (e->comm.ev_main) cudaEventDestroy(e->comm.ev_main);
    if (e->comm.ev_comm) cudaEventDestroy(e->comm.ev_comm);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return PF_OK;
}

static __global__ void fill_kernel(double *p, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

extern "C" int pf_create(pf_engine **out, const pf_config *cfg)
{
    if (!out || !cfg) return fail(PF_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->n_loci <= 0 || cfg->locus_begin < 0 || cfg->locus_end > cfg->n_loci ||
        cfg->locus_begin >= cfg->locus_end || cfg->cap_indiv <= 0 || cfg->cap_marr < 0 || cfg->n_chains < 1)
        return fail(PF_EINVAL, "bad config (n_loci=%d shard=[%d,%d) cap_indiv=%d cap_marr=%d n_chains=%d)",
                    cfg->n_loci, cfg->locus_begin, cfg->locus_end, cfg->cap_indiv, cfg->cap_marr, cfg->n_chains);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(PF_ENODEVICE, "no CUDA device available (%s); libpedfac_cuda has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PF_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));

    pf_engine *e = new pf_engine();
    e->cfg = *cfg;
    e->Sl = cfg->locus_end - cfg->locus_begin;
    e->Sp = (e->Sl + 63) / 64 * 64;
    const size_t rows_total = (size_t)ROW_CHAIN0 + (size_t)cfg->n_chains * ((size_t)cfg->cap_indiv + cfg->cap_marr);
    if (rows_total >= (size_t)SCRATCH_BIT) { delete e; return fail(PF_EINVAL, "too many rows"); }
    e->n_rows = rows_total;

#define CUX(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { int rc = fail(_e == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); pf_destroy(e); return rc; } } while (0)
    CUX(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
    CUX(cudaMalloc(&e->d_rows, rows_total * e->row_elems() * sizeof(double)));
    CUX(cudaMalloc(&e->d_eps, e->Sp * sizeof(double)));
    CUX(cudaMalloc(&e->d_hard, (size_t)cfg->cap_indiv * (e->Sp / 16) * sizeof(uint32_t)));
    CUX(cudaMalloc(&e->d_kind, cfg->cap_indiv));
    CUX(cudaMalloc(&e->d_eptr, (size_t)cfg->cap_indiv * sizeof(void *)));
    CUX(cudaMalloc(&e->d_denom, (size_t)cfg->cap_indiv * sizeof(uint32_t)));
    CUX(cudaMemsetAsync(e->d_eps, 0, e->Sp * sizeof(double), e->stream));
    // prior and ones rows start as ones (also keeps padded loci harmless)
    fill_kernel<<<64, 256, 0, e->stream>>>(e->d_rows, 2 * e->row_elems(), 1.0);
    e->launches++;
    CUX(cudaGetLastError());
    CUX(cudaStreamSynchronize(e->stream));
#undef CUX
    e->h_kind.assign(cfg->cap_indiv, EMIS_NONE);
    e->h_eptr.assign(cfg->cap_indiv, nullptr);
    e->h_denom.assign(cfg->cap_indiv, 1);
    e->h_rcp.assign(cfg->cap_indiv, 1.0);
    e->soft_kind_of_row.assign(cfg->cap_indiv, EMIS_NONE);
    e->h_eps.assign(e->Sl, 0.0);
    e->loc.assign(cfg->cap_indiv, -1);
    e->no_mailbox = getenv("PEDFAC_NO_MAILBOX") != nullptr;
    e->host_prof = getenv("PF_HOST_PROF") != nullptr;
    e->sweep_clocks = getenv("PF_SWEEP_CLOCKS") != nullptr;
    if (const char *v2 = getenv("PF_SWEEP_V2")) e->sweep_v2 = v2[0] != '0';
    if (const char *v2 = getenv("PF_EVAL_SINGLES")) e->eval_singles = v2[0] != '0';
    if (const char *v2 = getenv("PF_EVAL_MIN_CHUNK")) e->eval_min_chunk = std::max(32, atoi(v2) / 32 * 32);
    if (const char *v2 = getenv("PF_SWAP_SITES")) e->swap_sites = v2[0] != '0';
    if (const char *v2 = getenv("PF_INLINE_KIDS")) e->inline_kids = std::min(FAM2_INLINE_KIDS, std::max(1, atoi(v2)));
    if (const char *v2 = getenv("PF_CHUNK_KIDS")) e->chunk_kids = std::min(CHUNK2_KIDS, std::max(1, atoi(v2)));
    if (const char *v2 = getenv("PF_EVAL_ZEROCOPY")) e->eval_zero_copy = v2[0] != '0';
    if (const char *v2 = getenv("PF_PDL")) set_eval_pdl(v2[0] != '0');
    if (const char *v2 = getenv("PF_EVAL_GROUP")) e->eval_small_group = std::max(1, std::min(4, atoi(v2)));
    if (const char *v2 = getenv("PF_EVAL_SMALL_PROPS")) e->eval_small_props = std::max(0, atoi(v2));
    if (const char *v2 = getenv("PF_EVAL_INLINE")) e->eval_inline = v2[0] != '0';
    if (const char *v2 = getenv("PF_FUSED_PUBLISH")) e->fused_publish = v2[0] != '0';
    if (const char *fs = getenv("PF_SWEEP_SHAPE")) e->force_shape = fs[0] == 't' ? 2 : fs[0] == 'l' ? 1 : 0;
    *out = e;
    return PF_OK;
}

extern "C" int pf_set_stream(pf_engine *e, void *cuda_stream)
{
    if (!e) return fail(PF_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->own_stream) { cudaStreamDestroy(e->stream); e->own_stream = false; }
    e->stream = static_cast<cudaStream_t>(cuda_stream);
    return PF_OK;
}

extern "C" int64_t pf_launch_count(const pf_engine *e) { return e ? e->launches : 0; }

extern "C" int pf_io_bytes(const pf_engine *e, uint64_t *h2d, uint64_t *d2h)
{
    if (!e || !h2d || !d2h) return fail(PF_EINVAL, "null argument");
    *h2d = e->io_h2d; *d2h = e->io_d2h;
    return PF_OK;
}

extern "C" int pf_plan_stats(const pf_engine *e, int32_t *out)
{
    if (!e || !out) return fail(PF_EINVAL, "null argument");
    memcpy(out, e->last_plan, sizeof e->last_plan);
    return PF_OK;
}

extern "C" int pf_profile(pf_engine *e, int32_t enable)
{
    if (!e) return fail(PF_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    for (auto &r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    e->prof.clear();
    e->prof_on = enable != 0;
    return PF_OK;
}

extern "C" int pf_profile_read(pf_engine *e, double *ms, int64_t *count)
{
    if (!e || !ms || !count) return fail(PF_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < 4; k++) { ms[k] = 0.0; count[k] = 0; }
    for (auto &r : e->prof) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.kind] += t; count[r.kind]++;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    e->prof.clear();
    return PF_OK;
}

extern "C" int pf_set_locus_params(pf_engine *e, const double *eps, const double *afreq)
{
    if (!e || !eps || !afreq) return fail(PF_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    const int Sl = e->Sl, Sp = e->Sp, s0 = e->cfg.locus_begin;
    std::vector<double> buf((size_t)4 * Sp, 1.0);
    for (int s = 0; s < Sl; s++) {
        // _ReadPedfile@0x1000099d2-0x100009ad4: q = 1 - a; [q*q, (2*(1-q))*q, (1-q)*(1-q)]
        const double q = 1.0 - afreq[s0 + s];
        buf[s] = q * q;
        buf[Sp + s] = (2.0 * (1.0 - q)) * q;
        buf[2 * Sp + s] = (1.0 - q) * (1.0 - q);
        buf[3 * Sp + s] = eps[s0 + s];
        e->h_eps[s] = eps[s0 + s];
    }
    for (int s = Sl; s < Sp; s++) buf[3 * Sp + s] = 0.0;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(e->d_rows + (size_t)ROW_PRIOR * e->row_elems(), buf.data(), (size_t)3 * Sp * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_eps, buf.data() + 3 * Sp, (size_t)Sp * sizeof(double), cudaMemcpyHostToDevice));
    return PF_OK;
}

static int soft_row(pf_engine *e, int slot, uint8_t kind, size_t elem_bytes, void **out)
{
    if (e->h_eptr[slot] && e->soft_kind_of_row[slot] == kind) { *out = e->h_eptr[slot]; return PF_OK; }
    // a slot that changes its format (f32 -> dec16 -> f32 ...) gets one row per format, kept for reuse
    if (e->soft_rows_by_kind.empty()) e->soft_rows_by_kind.assign((size_t)e->cfg.cap_indiv * 8, nullptr);
    if (void *kept = e->soft_rows_by_kind[(size_t)slot * 8 + (kind & 7)]) { *out = kept; e->soft_kind_of_row[slot] = kind; return PF_OK; }
    const size_t bytes = ((size_t)3 * e->Sp * elem_bytes + 255) & ~(size_t)255;
    if (e->soft_chunks.empty() || e->soft_chunks.back().used + bytes > e->soft_chunks.back().cap) {
        SoftChunk c{};
        c.cap = std::max<size_t>(bytes, (size_t)64 << 20);
        cudaError_t ce = cudaMalloc(&c.base, c.cap);
        if (ce != cudaSuccess) return fail(PF_ENOMEM, "cudaMalloc(%zu) for soft genotypes failed: %s", c.cap, cudaGetErrorString(ce));
        e->soft_chunks.push_back(c);
    }
    SoftChunk &c = e->soft_chunks.back();
    *out = c.base + c.used;
    c.used += bytes;
    e->soft_kind_of_row[slot] = kind;
    e->soft_rows_by_kind[(size_t)slot * 8 + (kind & 7)] = *out;
    return PF_OK;
}

template <typename T>
static int set_soft(pf_engine *e, int32_t slot, const T *lik, uint8_t kind, uint32_t denom)
{
    if (!e || !lik) return fail(PF_EINVAL, "null argument");
    if (slot < 0 || slot >= e->cfg.cap_indiv) return fail(PF_EINVAL, "slot %d out of range", slot);
    CU(cudaSetDevice(e->cfg.device));
    void *row = nullptr;
    int rc = soft_row(e, slot, kind, sizeof(T), &row);
    if (rc) return rc;
    const int Sl = e->Sl, Sp = e->Sp, s0 = e->cfg.locus_begin;
    std::vector<T> buf((size_t)3 * Sp, (T)1);
    for (int s = 0; s < Sl; s++)
        for (int g = 0; g < 3; g++) buf[(size_t)g * Sp + s] = lik[(size_t)(s0 + s) * 3 + g];
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(row, buf.data(), buf.size() * sizeof(T), cudaMemcpyHostToDevice));
    e->h_kind[slot] = kind; e->h_eptr[slot] = row; e->h_denom[slot] = denom; e->h_rcp[slot] = 1.0 / (double)denom;
    e->tables_dirty = true;
    return PF_OK;
}

extern "C" int pf_set_genotypes_soft(pf_engine *e, int32_t slot, const float *lik) { return set_soft<float>(e, slot, lik, EMIS_F32, 1); }
extern "C" int pf_set_genotypes_soft_f64(pf_engine *e, int32_t slot, const double *lik) { return set_soft<double>(e, slot, lik, EMIS_F64, 1); }
extern "C" int pf_set_genotypes_soft_dec16(pf_engine *e, int32_t slot, const uint16_t *num, uint32_t denom)
{
    if (denom == 0) return fail(PF_EINVAL, "denominator 0");
    return set_soft<uint16_t>(e, slot, num, EMIS_DEC16, denom);
}

extern "C" int pf_set_genotypes_hard(pf_engine *e, int32_t slot, const uint8_t *codes)
{
    if (!e || !codes) return fail(PF_EINVAL, "null argument");
    if (slot < 0 || slot >= e->cfg.cap_indiv) return fail(PF_EINVAL, "slot %d out of range", slot);
    CU(cudaSetDevice(e->cfg.device));
    const int Sl = e->Sl, words = e->Sp / 16, s0 = e->cfg.locus_begin;
    std::vector<uint32_t> buf(words, 0xffffffffu);   // padding = code 3 (missing)
    for (int s = 0; s < Sl; s++) {
        const uint32_t c = codes[s0 + s];
        if (c > 3) return fail(PF_EINVAL, "Acceptable genotype values are 0, 1, 2, or 3 (slot %d, locus %d: %u)", slot, s0 + s, c);
        buf[s >> 4] = (buf[s >> 4] & ~(3u << ((s & 15) * 2))) | (c << ((s & 15) * 2));
    }
    uint32_t *row = e->d_hard + (size_t)slot * words;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(row, buf.data(), (size_t)words * sizeof(uint32_t), cudaMemcpyHostToDevice));
    e->h_kind[slot] = EMIS_HARD; e->h_eptr[slot] = row; e->h_denom[slot] = 1; e->h_rcp[slot] = 1.0;
    e->soft_kind_of_row[slot] = EMIS_NONE;
    e->tables_dirty = true;
    return PF_OK;
}

extern "C" int pf_set_genotypes_unobserved(pf_engine *e, int32_t slot)
{
    if (!e) return fail(PF_EINVAL, "null engine");
    if (slot < 0 || slot >= e->cfg.cap_indiv) return fail(PF_EINVAL, "slot %d out of range", slot);
    if (e->h_kind[slot] != EMIS_NONE) { e->h_kind[slot] = EMIS_NONE; e->tables_dirty = true; }
    return PF_OK;
}

static int sync_tables(pf_engine *e)
{
    if (!e->tables_dirty) return PF_OK;
    CU(cudaMemcpyAsync(e->d_kind, e->h_kind.data(), e->h_kind.size(), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_eptr, e->h_eptr.data(), e->h_eptr.size() * sizeof(void *), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_denom, e->h_denom.data(), e->h_denom.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));   // the sources are pageable vectors that may change next
    e->tables_dirty = false;
    return PF_OK;
}

static DevView dev_view(pf_engine *e)
{
    DevView D;
    D.rows = e->d_rows; D.scratch = static_cast<double *>(e->scratch.p);
    D.eps = e->d_eps; D.ekind = e->d_kind; D.eptr = e->d_eptr; D.edenom = e->d_denom;
    D.Sl = e->Sl; D.Sp = e->Sp;
    return D;
}

// ---------------------------------------------------------------------------------------------
// Sweep plan builder

namespace {

struct PlanTask { int group, level, kind, off, cost; };     // cost: rough instruction count, orders the tasks of a crowded level

struct Plan {
    std::vector<PlanTask> tasks;
    std::vector<int> pool;
    std::vector<EmisEnt> emis;
    std::vector<SweepGroup> groups;
    int n_levels = 1, n_out = 0;
    int target_tasks_per_group = 1 << 30;
    int scratch_rows = 0;
    void reset()        // empties the plan but keeps the vectors' capacity (plans are built per call)
    {
        tasks.clear(); pool.clear(); emis.clear(); groups.clear();
        n_levels = 1; n_out = 0; target_tasks_per_group = 1 << 30; scratch_rows = 0;
    }

    // Whole components are appended to the open group; a new group starts once the open one has
    // reached the target size (components are never split: their tasks depend on each other).
    void begin_component()
    {
        if (groups.empty() || groups.back().task_end >= target_tasks_per_group) {
            SweepGroup g{};
            g.pool_begin = (int)pool.size();
            g.emis_begin = (int)emis.size();
            groups.push_back(g);
        }
    }
    SweepGroup &cur() { return groups.back(); }
};

struct ViewWork {           // scratch vectors reused across calls
    std::vector<int> adj_off, adj_m, adj_role;      // individual -> (marriage, role)
    std::vector<int> mem_off, mem_i, mem_role;      // marriage -> members (pa, ma, kids)
    std::vector<int> Ui, um, urole, depth_m, lvlA;
    std::vector<int> up_i_row, up_m_row, dn_row, in_u_row, emis_idx;
    std::vector<int> bfs_m, queue, cnt_a, cnt_b, cnt_c, tmp_row;
    std::vector<char> vis_i, vis_m, label_seen, in_chunk;
    std::vector<int> roots, comp_members_end, order_i, q;    // BFS order of individuals per component
    std::vector<int> level_off, start, cost_of;
    std::vector<std::pair<int, SweepTask>> sort_tmp;
    std::vector<int> it_off, it_cnt, it_type, it_first, it_n, it_kids, it_prow;   // sibship items per marriage (see item_factor)
};

// ---- second-generation plan (pf_sweep2.cuh): compiled task records ----
struct Plan2Task { int group, level; uint32_t off; int cost; };
struct Plan2 {
    std::vector<Plan2Task> tasks;
    std::vector<uint32_t> pool;
    std::vector<Emis2> emis;
    std::vector<Stage2> stage;
    std::vector<Sweep2Group> groups;
    int target_tasks_per_group = 1 << 30;
    int format = 0;             // Fmt2 of the launch (0 generic, 1 hard calls, 2 dec16 over one denominator)
    bool fits = true;           // false: some count exceeded a 16-bit field -> first-generation kernel
    void reset() { tasks.clear(); pool.clear(); emis.clear(); stage.clear(); groups.clear(); target_tasks_per_group = 1 << 30; fits = true; format = 0; }
    void begin_component()
    {
        // a group is one block's shared memory: whole components are appended to the open group until it
        // has reached its share of the tasks, or a quarter of the shared memory (the largest component
        // that fits at all needs about three quarters: it must still fit behind what is there)
        bool open_new = groups.empty() || groups.back().task_end >= target_tasks_per_group;
        if (!open_new) {
            Sweep2Group g = groups.back();
            g.task_begin = 0; g.pool_len = (int)pool.size() - g.pool_begin;
            open_new = sweep2_smem_layout(g, 64).total > SWEEP2_SMEM_MAX / 4;
        }
        if (open_new) {
            Sweep2Group g{};
            g.pool_begin = (int)pool.size();
            g.emis_begin = (int)emis.size();
            g.n_rows = 2;                       // row 0: ones, row 1: trash
            g.emis_bytes = format == 1 ? 16 : format == 2 ? 3 * TILE_LOCI * 2 : 0;   // tile 0: the constant "no data" tile
            pool.push_back(0);                  // word 0 of a group's pool is never a list (offset 0 = "no extra rows")
            groups.push_back(g);
        }
    }
    Sweep2Group &cur() { return groups.back(); }
};

} // namespace

static thread_local ViewWork g_work;

// Appends the tasks of one view to `plan`. Levels are assigned in two bands: FAMILY_UP tasks get
// their height above the leaves; ROOT/FAMILY_DOWN tasks get (band offset) + depth below the root.
// The band offset is fixed up by the caller once every view has been added. While a group is
// being filled, SweepGroup::task_end counts its tasks and n_rows / emis_count its transient rows
// and genotype entries.
// A swap site: the swap proposals of one Gibbs step that rewire the same two families (same x, z and
// marriages). The view handed to the planner is the hypothetical component of the first of them, listed
// with x first (the tree is rooted there, not at its centre) and with n_cc = the number of variants; the
// others differ only in which of the four parents and which of the two sibships go to which family, i.e.
// in the three family tasks next to the root: emit2 adds those per variant and shares everything below.
struct SwapVariantSpec { int za, zb, xa, xb, move; };      // parent slots of z's and x's new family; siblings moved?
struct SwapSite { int marr_x_slot = -1; std::vector<SwapVariantSpec> var; };

static int emit2(pf_engine *e, const pf_graph_view *v, int chain, int out_base, Plan2 &plan, ViewWork &w, int &max_up_levels, bool up_only, const SwapSite *site);

static int plan_add_view(pf_engine *e, const pf_graph_view *v, int chain, int out_base, Plan &plan, int &max_up_levels, bool up_only, const SweepShape &shape, int *loc = nullptr, Plan2 *plan2 = nullptr, const SwapSite *site = nullptr)
{
    if (!loc) loc = e->loc.data();
    if (!v) return fail(PF_EINVAL, "null view");
    const int nI = v->n_indiv, nM = v->n_marr;
    if (nI < 0 || nM < 0 || v->n_cc < 0) return fail(PF_EINVAL, "negative count in view");
    if (nI > 0 && (!v->indiv || !v->indiv_cc || !v->indiv_founder)) return fail(PF_EINVAL, "null individual arrays");
    if (nM > 0 && (!v->marr || !v->marr_pa || !v->marr_ma || !v->marr_selfing || !v->marr_kid_begin)) return fail(PF_EINVAL, "null marriage arrays");
    if (chain < 0 || chain >= e->cfg.n_chains) return fail(PF_EINVAL, "chain %d out of range", chain);
    ViewWork &w = g_work;
    int rc = PF_OK;
    int registered = 0;

    const bool hp_on = e->host_prof && e->n_prof_calls >= 100;
    double hp_t = hp_on ? host_now() : 0.0;
    auto hp_mark = [&](int ph) { if (hp_on) { const double t = host_now(); e->t_ph[ph] += t - hp_t; hp_t = t; } };
    // slot -> local index
    for (int i = 0; i < nI; i++) {
        const int slot = v->indiv[i];
        if (slot < 0 || slot >= e->cfg.cap_indiv || loc[slot] != -1) { rc = fail(PF_EINVAL, "bad or duplicate individual slot %d in view", slot); goto cleanup; }
        if (v->indiv_cc[i] < 0 || v->indiv_cc[i] >= v->n_cc) { rc = fail(PF_EINVAL, "component label %d of slot %d outside [0,%d)", v->indiv_cc[i], slot, v->n_cc); goto cleanup; }
        loc[slot] = i;
        registered = i + 1;
    }
    {
        // members of each marriage
        w.mem_off.assign(nM + 1, 0); w.mem_i.clear(); w.mem_role.clear();
        w.adj_off.assign(nI + 1, 0);
        for (int j = 0; j < nM; j++) {
            if (v->marr[j] < 0 || v->marr[j] >= e->cfg.cap_marr) { rc = fail(PF_EINVAL, "marriage slot %d out of range", v->marr[j]); goto cleanup; }
            const int kb = v->marr_kid_begin[j], ke = v->marr_kid_begin[j + 1];
            if (ke < kb) { rc = fail(PF_EINVAL, "marr_kid_begin not monotone"); goto cleanup; }
            const int ne = 2 + (ke - kb);
            for (int ed = 0; ed < ne; ed++) {
                if (ed == 1 && v->marr_selfing[j]) continue;
                const int slot = ed == 0 ? v->marr_pa[j] : ed == 1 ? v->marr_ma[j] : v->marr_kids[kb + ed - 2];
                if (slot < 0 || slot >= e->cfg.cap_indiv || loc[slot] < 0) { rc = fail(PF_EINVAL, "view not closed: marriage %d references unlisted individual %d", v->marr[j], slot); goto cleanup; }
                w.mem_i.push_back(loc[slot]);
                w.mem_role.push_back(ed < 2 ? ed : 2);
                w.adj_off[loc[slot] + 1]++;
            }
            w.mem_off[j + 1] = (int)w.mem_i.size();
        }
        for (int i = 0; i < nI; i++) w.adj_off[i + 1] += w.adj_off[i];
        w.adj_m.assign(w.mem_i.size(), 0); w.adj_role.assign(w.mem_i.size(), 0);
        std::vector<int> &fillpos = w.queue;
        fillpos.assign(w.adj_off.begin(), w.adj_off.end() - 1);
        for (int j = 0; j < nM; j++)
            for (int k = w.mem_off[j]; k < w.mem_off[j + 1]; k++) {
                const int i = w.mem_i[k];
                w.adj_m[fillpos[i]] = j; w.adj_role[fillpos[i]] = w.mem_role[k]; fillpos[i]++;
            }

        w.Ui.assign(nI, -1); w.um.assign(nM, -1); w.urole.assign(nM, 0); w.depth_m.assign(nM, 0);
        w.lvlA.assign(nM, 0); w.vis_i.assign(nI, 0); w.vis_m.assign(nM, 0); w.label_seen.assign(v->n_cc, 0);
        w.it_off.assign(nM, 0); w.it_cnt.assign(nM, 0); w.it_type.clear(); w.it_first.clear(); w.it_n.clear(); w.it_kids.clear(); w.it_prow.clear(); w.in_chunk.assign(nI, 0);
        w.up_i_row.assign(nI, -1); w.up_m_row.assign(nM, -1); w.in_u_row.assign(nM, -1); w.tmp_row.assign(nM, -1); w.dn_row.assign(nI, -1); w.emis_idx.assign(nI, -1);
        w.bfs_m.clear();
        std::vector<int> &roots = w.roots, &comp_members_end = w.comp_members_end, &order_i = w.order_i, &q = w.q;
        roots.clear(); comp_members_end.clear(); order_i.clear(); q.clear();
        const size_t no_m = (size_t)-1;
        // BFS over the bipartite tree from individual r: fills Ui / um / urole / depth_m, appends
        // the marriages to bfs_m, leaves the individuals in q. Returns an error code.
        auto bfs = [&](int r, int label) -> int {
            w.vis_i[r] = 1; w.Ui[r] = -1;
            q.clear(); q.push_back(r);
            for (size_t h = 0; h < q.size(); h++) {
                const int i = q[h];
                const int di = w.Ui[i] < 0 ? 0 : w.depth_m[w.Ui[i]] + 1;
                for (int a = w.adj_off[i]; a < w.adj_off[i + 1]; a++) {
                    const int m = w.adj_m[a];
                    if (m == w.Ui[i]) continue;                        // the edge we arrived by
                    if (w.vis_m[m]) return fail(PF_ECYCLE, "the graph view contains a loop through marriage %d", v->marr[m]);
                    w.vis_m[m] = 1; w.um[m] = i; w.urole[m] = w.adj_role[a]; w.depth_m[m] = di;
                    w.bfs_m.push_back(m);
                    for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                        const int y = w.mem_i[k];
                        if (y == i && w.mem_role[k] == w.adj_role[a]) continue;
                        if (w.vis_i[y]) return fail(PF_ECYCLE, "the graph view contains a loop through individual %d", v->indiv[y]);
                        if (v->indiv_cc[y] != label) return fail(PF_EINVAL, "individuals %d and %d are connected but carry different component labels", v->indiv[i], v->indiv[y]);
                        w.vis_i[y] = 1; w.Ui[y] = m;
                        q.push_back(y);
                    }
                }
            }
            return PF_OK;
        };
        auto unvisit = [&](size_t m_from) {
            for (int i : q) w.vis_i[i] = 0;
            for (size_t b = m_from; b < w.bfs_m.size(); b++) w.vis_m[w.bfs_m[b]] = 0;
            w.bfs_m.resize(m_from);
        };
        (void)no_m;
        hp_mark(0);
        for (int r0 = 0; r0 < nI; r0++) {
            if (w.vis_i[r0]) continue;
            const int label = v->indiv_cc[r0];
            if (w.label_seen[label]) { rc = fail(PF_EINVAL, "component label %d covers individuals that are not connected (slot %d)", label, v->indiv[r0]); goto cleanup; }
            w.label_seen[label] = 1;
            const size_t m_from = w.bfs_m.size();
            rc = bfs(r0, label);
            if (rc) goto cleanup;
            int r = r0;
            if (w.bfs_m.size() - m_from >= 3 && !site) {
                // re-root at the centre of the tree: the number of levels is the tree's radius
                // instead of up to its diameter. Two more traversals: farthest node from r0, then
                // the farthest from that one; the centre is half way along that path.
                const int a_end = q.back();
                unvisit(m_from);
                bfs(a_end, label);
                int c = q.back();
                const int diam = w.Ui[c] < 0 ? 0 : w.depth_m[w.Ui[c]] + 1;
                for (int step = 0; step < diam / 2; step++) c = w.um[w.Ui[c]];
                unvisit(m_from);
                bfs(c, label);
                r = c;
            }
            roots.push_back(r);
            order_i.insert(order_i.end(), q.begin(), q.end());
            comp_members_end.push_back((int)order_i.size());
        }
        for (int m = 0; m < nM; m++)
            if (!w.vis_m[m]) { rc = fail(PF_EINVAL, "marriage %d not reachable", v->marr[m]); goto cleanup; }
        hp_mark(1);
        if (plan2) { rc = emit2(e, v, chain, out_base, *plan2, w, max_up_levels, up_only, site); goto cleanup; }
        // heights of the FAMILY_UP tasks, children first (reverse BFS order)
        int up_levels = 0;
        for (int b = (int)w.bfs_m.size() - 1; b >= 0; b--) {
            const int m = w.bfs_m[b];
            int lv = 0;
            for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                const int x = w.mem_i[k];
                if (x == w.um[m]) continue;
                for (int a = w.adj_off[x]; a < w.adj_off[x + 1]; a++)
                    if (w.adj_m[a] != m) lv = std::max(lv, w.lvlA[w.adj_m[a]] + 1);
            }
            w.lvlA[m] = lv;
            up_levels = std::max(up_levels, lv + 1);
        }
        max_up_levels = std::max(max_up_levels, up_levels);

        hp_mark(2);
        // emit tasks component by component (BFS order keeps a component's marriages contiguous)
        std::vector<int> &pool = plan.pool;
        auto push_row = [&](int ref) { pool.push_back(ref); };     // transient rows: the kernel tells shared from spilled by the row number
        auto below_rows = [&](int x, int except_m) {
            int n = 0;
            const size_t at = pool.size();
            pool.push_back(0);
            for (int a = w.adj_off[x]; a < w.adj_off[x + 1]; a++) {
                const int m2 = w.adj_m[a];
                if (m2 == except_m || m2 == w.Ui[x]) continue;
                push_row(w.up_m_row[m2]); n++;
            }
            pool[at] = n;
        };
        auto member = [&](int x, int role) { return role | (v->indiv_founder[x] ? 4 : 0) | (w.emis_idx[x] << 3); };
        size_t bfs_pos = 0;
        for (size_t ri = 0; ri < roots.size(); ri++) {
            const int r = roots[ri];
            plan.begin_component();
            SweepGroup &G = plan.cur();
            const int gid = (int)plan.groups.size() - 1;
            const int ib = ri == 0 ? 0 : comp_members_end[ri - 1], ie = comp_members_end[ri];
            // the component's marriages: the run of bfs_m whose up-individual carries this label
            size_t mb = bfs_pos, me = bfs_pos;
            while (me < w.bfs_m.size() && v->indiv_cc[w.um[w.bfs_m[me]]] == v->indiv_cc[r]) me++;
            bfs_pos = me;
            // group-local transient rows. up_m (one per marriage) and up_i (only for individuals
            // with marriages below them: a leaf's up message is its local factor, recomputed from
            // the staged genotype tile) live for the whole sweep. in_u(m) and dn_m(x) are written
            // and consumed within two consecutive levels of the down band, so every depth reuses
            // the same small bank of rows.
            for (size_t b = mb; b < me; b++) w.up_m_row[w.bfs_m[b]] = SCRATCH_BIT | G.n_rows++;
            // sibship items of every marriage of the component: down kids that have marriages of
            // their own are single items, leaf kids are grouped FAM_CHUNK at a time (a left-over
            // single leaf is a single item)
            for (size_t b = mb; b < me; b++) {
                const int m = w.bfs_m[b];
                w.it_off[m] = (int)w.it_type.size();
                int run = 0;
                for (int pass = 0; pass < 2; pass++)
                    for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                        const int x = w.mem_i[k];
                        if (x == w.um[m] || w.mem_role[k] != 2) continue;
                        const bool leaf = w.adj_off[x + 1] - w.adj_off[x] <= 1;
                        if (pass == 0 && !leaf) { w.it_type.push_back(0); w.it_first.push_back((int)w.it_kids.size()); w.it_n.push_back(1); w.it_prow.push_back(-1); w.it_kids.push_back(k); }
                        if (pass == 1 && leaf) {
                            if (run == 0) { w.it_type.push_back(0); w.it_first.push_back((int)w.it_kids.size()); w.it_n.push_back(0); w.it_prow.push_back(-1); }
                            w.it_kids.push_back(k); w.it_n.back()++;
                            if (++run == FAM_CHUNK) run = 0;
                        }
                    }
                for (int it = w.it_off[m]; it < (int)w.it_type.size(); it++)
                    if (w.it_n[it] >= 2) {
                        w.it_type[it] = 1;
                        for (int j = 0; j < w.it_n[it]; j++) w.in_chunk[w.mem_i[w.it_kids[w.it_first[it] + j]]] = 1;
                    }
                w.it_cnt[m] = (int)w.it_type.size() - w.it_off[m];
            }
            if (!up_only) {
                // Down-band levels by actual dependency instead of two per tree depth: FAM_DOWN(m)
                // runs one level after the task that writes the down message of its up individual
                // u — FAM_DOWN of u's up marriage when u is a parent there (or its single kid),
                // that marriage's KID_CHUNK level when u is one of several kids. Going up through
                // ancestors therefore costs one level per marriage, going down through a sibship two.
                // (bfs_m lists a component's marriages parents-first.)
                auto has_tmp = [&](int m) { const int ni = w.it_cnt[m]; return ni >= 2 || (ni == 1 && w.it_type[w.it_off[m]] == 1); };
                int max_l = 0;
                for (size_t b = mb; b < me; b++) {
                    const int m = w.bfs_m[b], u = w.um[m];
                    int l = 0;
                    if (w.Ui[u] >= 0) {
                        const int mp = w.Ui[u];
                        int role = 2;
                        for (int k = w.mem_off[mp]; k < w.mem_off[mp + 1]; k++) if (w.mem_i[k] == u && w.mem_i[k] != w.um[mp]) role = w.mem_role[k];
                        l = w.depth_m[mp] + 1 + ((role == 2 && has_tmp(mp)) ? 1 : 0);
                    }
                    w.depth_m[m] = l;               // from here on depth_m holds the down-band level of FAM_DOWN(m)
                    max_l = std::max(max_l, l + 1);
                }
                std::vector<int> &cnt_d = w.cnt_b, &cnt_t = w.cnt_c;
                cnt_d.assign(max_l + 2, 0); cnt_t.assign(max_l + 2, 0);
                int bank_d = 0, bank_t = 0;
                for (size_t b = mb; b < me; b++) {
                    const int m = w.bfs_m[b], l = w.depth_m[m];
                    w.tmp_row[m] = -1;
                    // sibships of more than one single kid park a, b and one row pair per item; the rows
                    // are written at level l and read at l + 1: two banks by the parity of l
                    if (has_tmp(m)) { w.tmp_row[m] = cnt_t[l]; cnt_t[l] += 4 + 2 * w.it_cnt[m]; bank_t = std::max(bank_t, cnt_t[l]); }
                    for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                        const int x = w.mem_i[k];
                        if (x == w.um[m] || w.adj_off[x + 1] - w.adj_off[x] <= 1) continue;
                        const int lw = l + ((w.mem_role[k] == 2 && has_tmp(m)) ? 1 : 0);   // level that writes the message to x
                        w.dn_row[x] = (cnt_d[lw]++) * 2 + (lw & 1);        // read at lw + 1, reusable from lw + 2
                        bank_d = std::max(bank_d, 2 * cnt_d[lw]);
                    }
                }
                const int base_d = G.n_rows, base_t = base_d + bank_d;
                G.n_rows += bank_d + 2 * bank_t;
                for (size_t b = mb; b < me; b++) {
                    const int m = w.bfs_m[b];
                    if (w.tmp_row[m] >= 0) w.tmp_row[m] += base_t + (w.depth_m[m] & 1) * bank_t;      // plain row numbers (ld_i / st_i)
                }
                for (int oi = ib; oi < ie; oi++) {
                    const int i = order_i[oi];
                    if (w.dn_row[i] >= 0) w.dn_row[i] = SCRATCH_BIT | (base_d + w.dn_row[i]);
                }
            }
            // up_i rows: every non-leaf; then leaves with genotype data while shared memory has
            // room — a cached leaf is decoded once (UP_IN at level 0, straight from global memory,
            // so its genotype tile is not staged) instead of once per family task that folds it
            for (int oi = ib; oi < ie; oi++) {
                const int i = order_i[oi];
                const bool leaf = w.adj_off[i + 1] - w.adj_off[i] <= 1;
                if (w.Ui[i] >= 0 && !leaf) w.up_i_row[i] = SCRATCH_BIT | G.n_rows++;
            }
            for (int oi = ib; oi < ie; oi++) {
                const int i = order_i[oi];
                const bool leaf = w.adj_off[i + 1] - w.adj_off[i] <= 1;
                bool cached_leaf = false;
                if (w.Ui[i] >= 0 && leaf && !w.in_chunk[i] && e->h_kind[v->indiv[i]] != EMIS_NONE && G.n_rows < shape.leaf_row_budget) {
                    w.up_i_row[i] = SCRATCH_BIT | G.n_rows++;
                    cached_leaf = true;
                }
                // genotype entry
                const int slot = v->indiv[i];
                EmisEnt E{};
                E.kind = e->h_kind[slot]; E.denom = e->h_denom[slot]; E.rcp = e->h_rcp[slot];
                const unsigned long long pp = (unsigned long long)(uintptr_t)e->h_eptr[slot];
                E.ptr_lo = (unsigned)(pp & 0xffffffffu); E.ptr_hi = (unsigned)(pp >> 32);
                const int bytes = emis_tile_bytes(E.kind);
                if (!cached_leaf && bytes > 0 && G.emis_bytes + bytes <= shape.emis_smem_max) { E.smem_off = (unsigned)G.emis_bytes; G.emis_bytes += (bytes + 7) & ~7; }
                else E.smem_off = EMIS_NOT_STAGED;
                w.emis_idx[i] = G.emis_count++;
                plan.emis.push_back(E);
            }
            // rows of the chunk products (written by KFOLD at level 0, read by FAM_UP and FAM_DOWN): last,
            // so that they are the first to spill when shared memory is short
            for (size_t b = mb; b < me; b++) {
                const int m = w.bfs_m[b];
                for (int it = w.it_off[m]; it < w.it_off[m] + w.it_cnt[m]; it++)
                    if (w.it_type[it] == 1) { w.it_prow[it] = G.n_rows; G.n_rows += 2; }
            }
            // reference to the up message of member x: its row, or (leaf) LEAF_REF_BIT | member
            auto push_up_i = [&](int x, int role) {
                if (w.up_i_row[x] >= 0) push_row(w.up_i_row[x]);
                else pool.push_back((int)(LEAF_REF_BIT | (unsigned)member(x, role)));
            };
            hp_mark(3);
            auto add_task = [&](int level, int kind, int cost) { plan.tasks.push_back(PlanTask{gid, level, kind, (int)pool.size(), cost}); G.task_end++; };
            // levels: >= 0 = up band; -1 = first level of the down band; -(2 + k) = down band level k
            const bool has_m = w.adj_off[r + 1] > w.adj_off[r];
            add_task(has_m ? -1 : 0, TASK_ROOT, 8);
            pool.push_back(member(r, 0));
            if (up_only) push_row(SCRATCH_BIT | G.n_rows++);        // nothing persisted: the root's belief goes to a transient row
            else pool.push_back(e->stub_row(chain, v->indiv[r]));
            pool.push_back(out_base + v->indiv_cc[r]);
            below_rows(r, -1);
            // cached leaves are decoded once, at level 0
            for (int oi = ib; oi < ie; oi++) {
                const int i = order_i[oi];
                const bool leaf = w.adj_off[i + 1] - w.adj_off[i] <= 1;
                if (!leaf || w.up_i_row[i] < 0) continue;
                add_task(0, TASK_LEAF_IN, 5);
                pool.push_back(member(i, 0)); push_row(w.up_i_row[i]);
            }
            // chunk products, also at level 0
            for (size_t b = mb; b < me; b++) {
                const int m = w.bfs_m[b];
                for (int it = w.it_off[m]; it < w.it_off[m] + w.it_cnt[m]; it++) {
                    if (w.it_type[it] != 1) continue;
                    add_task(0, TASK_KFOLD, 4 + 6 * w.it_n[it]);
                    pool.push_back(w.it_prow[it]); pool.push_back(w.it_n[it]);
                    for (int j = 0; j < w.it_n[it]; j++) push_up_i(w.mem_i[w.it_kids[w.it_first[it] + j]], 2);
                }
            }
            // the sibship items of marriage m as {type, value} pairs
            auto push_item_list = [&](int m) {
                for (int it = w.it_off[m]; it < w.it_off[m] + w.it_cnt[m]; it++) {
                    pool.push_back(w.it_type[it]);
                    if (w.it_type[it] == 1) pool.push_back(w.it_prow[it]);
                    else push_up_i(w.mem_i[w.it_kids[w.it_first[it]]], 2);
                }
            };
            for (size_t b = mb; b < me; b++) {
                const int m = w.bfs_m[b];
                const int flags = (v->marr_selfing[m] ? 1 : 0) | (w.urole[m] << 1);
                const int u = w.um[m];
                int nd = 0;
                for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) if (w.mem_i[k] != u) nd++;
                add_task(1 + w.lvlA[m], TASK_FAM_UP, 10 + 4 * nd + 3 * w.it_cnt[m]);
                pool.push_back(flags); push_row(w.up_m_row[m]);
                {
                    const size_t nc_at = pool.size(); pool.push_back(0);
                    int ncomp = 0;
                    for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                        const int x = w.mem_i[k];
                        if (x == u || w.adj_off[x + 1] - w.adj_off[x] <= 1) continue;
                        pool.push_back(member(x, w.mem_role[k])); push_row(w.up_i_row[x]);
                        below_rows(x, m);
                        ncomp++;
                    }
                    pool[nc_at] = ncomp;
                }
                int pa_k = -1, ma_k = -1, nkid = 0;
                for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                    if (w.mem_i[k] == u) continue;
                    if (w.mem_role[k] == 0) pa_k = k; else if (w.mem_role[k] == 1) ma_k = k; else nkid++;
                }
                if (pa_k >= 0) push_up_i(w.mem_i[pa_k], 0); else pool.push_back(NO_REF);
                if (ma_k >= 0) push_up_i(w.mem_i[ma_k], 1); else pool.push_back(NO_REF);
                (void)nkid;
                pool.push_back(w.it_cnt[m]);
                push_item_list(m);
                if (up_only) continue;                              // log Z only: no down band
                add_task(-(2 + w.depth_m[m]), TASK_FAM_DOWN, 24 + 3 * w.it_cnt[m]);
                pool.push_back(flags); pool.push_back(e->prong_row(chain, v->marr[m]));
                pool.push_back(member(u, w.urole[m]));
                if (w.Ui[u] < 0) pool.push_back(-1); else push_row(w.dn_row[u]);
                below_rows(u, m);
                auto down_entry = [&](int k) {
                    if (k < 0) { pool.push_back(NO_REF); pool.push_back(-1); pool.push_back(-1); return; }
                    const int x = w.mem_i[k];
                    push_up_i(x, w.mem_role[k]);
                    pool.push_back(e->stub_row(chain, v->indiv[x]));
                    if (w.dn_row[x] < 0) pool.push_back(-1); else push_row(w.dn_row[x]);
                };
                down_entry(pa_k); down_entry(ma_k);
                pool.push_back(w.it_cnt[m]);
                pool.push_back(w.tmp_row[m]);
                push_item_list(m);
                if (w.tmp_row[m] < 0) {
                    if (w.it_cnt[m] == 1) {         // one single kid, finished inside FAM_DOWN: {ref (above), stub_row, dn_row}
                        const int x = w.mem_i[w.it_kids[w.it_first[w.it_off[m]]]];
                        pool.push_back(e->stub_row(chain, v->indiv[x]));
                        if (w.dn_row[x] < 0) pool.push_back(-1); else push_row(w.dn_row[x]);
                    }
                    continue;
                }
                // beside FAM_DOWN, on another warp: the product of the other items' factors, per item
                if (w.it_cnt[m] >= 2) {
                    add_task(-(2 + w.depth_m[m]), TASK_FAM_EXCL, 4 + 8 * w.it_cnt[m]);
                    pool.push_back(w.it_cnt[m]);
                    pool.push_back(w.tmp_row[m] + 4);
                    push_item_list(m);
                }
                // one KID_CHUNK task per item on the next level
                for (int it = w.it_off[m], i = 0; it < w.it_off[m] + w.it_cnt[m]; it++, i++) {
                    add_task(-(2 + w.depth_m[m] + 1), TASK_KID_CHUNK, 6 + 12 * w.it_n[it]);
                    pool.push_back(flags);
                    pool.push_back(w.tmp_row[m]);
                    pool.push_back(w.it_cnt[m] >= 2 ? w.tmp_row[m] + 4 + 2 * i : -1);
                    pool.push_back(w.it_n[it]);
                    for (int j = 0; j < w.it_n[it]; j++) {
                        const int x = w.mem_i[w.it_kids[w.it_first[it] + j]];
                        push_up_i(x, 2);
                        pool.push_back(e->stub_row(chain, v->indiv[x]));
                        if (w.dn_row[x] < 0) pool.push_back(-1); else push_row(w.dn_row[x]);
                    }
                }
            }
            G.pool_len = (int)pool.size() - G.pool_begin;
            hp_mark(4);
        }
    }
cleanup:
    for (int i = 0; i < registered; i++) loc[v->indiv[i]] = -1;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Second-generation plan builder (record formats: pf_sweep2.cuh). Runs after plan_add_view's
// validation and traversal, on the same ViewWork: adjacency (adj_*), members (mem_*), the rooted
// tree (Ui, um, urole), the marriages in BFS order (bfs_m), roots / order_i / comp_members_end.

namespace {
struct Work2 {
    std::vector<int> kid_off, kid_x, n_inline, chunk_off, n_chunks, chunk_first, chunk_n, chunk_row, chunk_level;
    std::vector<int> lvl2, up_m_row, dn_row, emis_idx, depth2, cnt_d, cnt_t, rows;
    std::vector<char> in_chunk;
    std::vector<uint32_t> rec;
    std::vector<uint32_t> famup_off; std::vector<int> famup_len;      // where the FAMUP2 record of each marriage went (swap sites)
};
}
static thread_local Work2 g_work2;

static int emit2(pf_engine *e, const pf_graph_view *v, int chain, int out_base, Plan2 &plan, ViewWork &w, int &max_up_levels, bool up_only, const SwapSite *site)
{
    Work2 &x2 = g_work2;
    const int nI = v->n_indiv, nM = v->n_marr;
    const bool hp_on = e->host_prof && e->n_prof_calls >= 100;
    double hp_t = hp_on ? host_now() : 0.0;
    auto hp_mark = [&](int ph) { if (hp_on) { const double t = host_now(); e->t_ph[ph] += t - hp_t; hp_t = t; } };
    x2.kid_off.assign(nM + 1, 0); x2.kid_x.clear(); x2.n_inline.assign(nM, 0); x2.chunk_off.assign(nM, 0); x2.n_chunks.assign(nM, 0);
    x2.chunk_first.clear(); x2.chunk_n.clear(); x2.chunk_row.clear(); x2.chunk_level.clear();
    x2.lvl2.assign(nM, 0); x2.up_m_row.assign(nM, 0); x2.dn_row.assign(nI, -1); x2.emis_idx.assign(nI, -1);
    x2.depth2.assign(nM, 0); x2.in_chunk.assign(nI, 0);
    if (site) { x2.famup_off.assign(nM, 0); x2.famup_len.assign(nM, 0); }
    uint32_t root_off = 0; int root_len = 0;
    std::vector<uint32_t> &pool = plan.pool, &rec = x2.rec;
    auto n_marr_of = [&](int i) { return w.adj_off[i + 1] - w.adj_off[i]; };
    // kids of every marriage below it (the up individual excluded), kids with marriages of their own first
    for (int m = 0; m < nM; m++) {
        x2.kid_off[m] = (int)x2.kid_x.size();
        for (int pass = 0; pass < 2; pass++)
            for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                const int i = w.mem_i[k];
                if (w.mem_role[k] != 2 || i == w.um[m]) continue;
                if ((n_marr_of(i) > 1) == (pass == 0)) x2.kid_x.push_back(i);
            }
    }
    x2.kid_off[nM] = (int)x2.kid_x.size();

    hp_mark(2);
    size_t bfs_pos = 0;
    for (size_t ri = 0; ri < w.roots.size(); ri++) {
        const int r = w.roots[ri];
        plan.begin_component();
        const int gid = (int)plan.groups.size() - 1;
        Sweep2Group &G = plan.cur();
        const int ib = ri == 0 ? 0 : w.comp_members_end[ri - 1], ie = w.comp_members_end[ri];
        size_t mb = bfs_pos, me = bfs_pos;
        while (me < w.bfs_m.size() && v->indiv_cc[w.um[w.bfs_m[me]]] == v->indiv_cc[r]) me++;
        bfs_pos = me;

        // genotype entries of the component's individuals
        for (int oi = ib; oi < ie; oi++) {
            const int i = w.order_i[oi], slot = v->indiv[i];
            const uint32_t kind = e->h_kind[slot];
            if (kind == EMIS_NONE) continue;
            const int align = kind == EMIS_HARD ? 8 : 16;
            G.emis_bytes = (G.emis_bytes + align - 1) & ~(align - 1);
            x2.emis_idx[i] = plan.format == 0 ? G.emis_count : G.emis_bytes;      // table index, or the tile's offset itself
            G.emis_count++;
            if (plan.format == 0) G.emis_tab++;
            plan.emis.push_back(Emis2{(uint32_t)G.emis_bytes, e->h_denom[slot], e->h_rcp[slot]});
            plan.stage.push_back(Stage2{(unsigned long long)(uintptr_t)e->h_eptr[slot], kind, (uint32_t)G.emis_bytes});
            G.emis_bytes += emis_tile_bytes(kind);
        }
        // rows: one per marriage, one pair per chunk of a large sibship
        for (size_t b = mb; b < me; b++) {
            const int m = w.bfs_m[b];
            x2.up_m_row[m] = G.n_rows++;
            const int nk = x2.kid_off[m + 1] - x2.kid_off[m];
            x2.n_inline[m] = std::min(nk, e->inline_kids);
            x2.chunk_off[m] = (int)x2.chunk_first.size();
            for (int k0 = x2.n_inline[m]; k0 < nk; k0 += e->chunk_kids) {
                const int n = std::min(e->chunk_kids, nk - k0);
                x2.chunk_first.push_back(x2.kid_off[m] + k0); x2.chunk_n.push_back(n);
                x2.chunk_row.push_back(G.n_rows); G.n_rows += 2;
                x2.chunk_level.push_back(0);
                for (int j = 0; j < n; j++) x2.in_chunk[x2.kid_x[x2.kid_off[m] + k0 + j]] = 1;
            }
            x2.n_chunks[m] = (int)x2.chunk_first.size() - x2.chunk_off[m];
        }
        // up-band levels by dependency, children first
        int up_levels = 0;
        auto below_level = [&](int i, int m) {
            int l = 0;
            for (int a = w.adj_off[i]; a < w.adj_off[i + 1]; a++)
                if (w.adj_m[a] != m) l = std::max(l, x2.lvl2[w.adj_m[a]] + 1);
            return l;
        };
        for (size_t b = me; b-- > mb;) {
            const int m = w.bfs_m[b];
            int lv = 0;
            for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                const int i = w.mem_i[k];
                if (i == w.um[m] || x2.in_chunk[i]) continue;
                lv = std::max(lv, below_level(i, m));
            }
            for (int c = x2.chunk_off[m]; c < x2.chunk_off[m] + x2.n_chunks[m]; c++) {
                int cl = 0;
                for (int j = 0; j < x2.chunk_n[c]; j++) cl = std::max(cl, below_level(x2.kid_x[x2.chunk_first[c] + j], m));
                x2.chunk_level[c] = cl;
                lv = std::max(lv, cl + 1);
            }
            x2.lvl2[m] = lv;
            up_levels = std::max(up_levels, lv + 1);
        }
        max_up_levels = std::max(max_up_levels, up_levels);
        // down band: FAMDOWN2(m) and the KIDCHUNK2 tasks of m run one level after the task that writes
        // the message into m's up individual; message rows are written at level l and read at l + 1:
        // two banks by the parity of l
        if (!up_only) {
            int max_l = 0;
            for (size_t b = mb; b < me; b++) {
                const int m = w.bfs_m[b], u = w.um[m];
                const int l = w.Ui[u] >= 0 ? x2.depth2[w.Ui[u]] + 1 : 0;
                x2.depth2[m] = l;
                max_l = std::max(max_l, l + 1);
            }
            x2.cnt_d.assign(max_l + 2, 0);
            int bank_d = 0;
            for (size_t b = mb; b < me; b++) {
                const int m = w.bfs_m[b], l = x2.depth2[m];
                for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                    const int i = w.mem_i[k];
                    if (i == w.um[m] || n_marr_of(i) <= 1) continue;
                    x2.dn_row[i] = (x2.cnt_d[l]++) * 2 + (l & 1);
                    bank_d = std::max(bank_d, 2 * x2.cnt_d[l]);
                }
            }
            const int base_d = G.n_rows;
            G.n_rows += bank_d;
            for (int oi = ib; oi < ie; oi++) {
                const int i = w.order_i[oi];
                if (x2.dn_row[i] >= 0) x2.dn_row[i] += base_d;
            }
        }

        hp_mark(3);
        // ---- records ----
        bool rec_extras = false;                               // some operand of the record being built has more than three rows
        auto push_cmp = [&](int i, int m_ex, int first_row) {
            std::vector<int> &rows = x2.rows;
            rows.clear();
            if (first_row >= 0) rows.push_back(first_row);
            for (int a = w.adj_off[i]; a < w.adj_off[i + 1]; a++) {
                const int m2 = w.adj_m[a];
                if (m2 == m_ex || m2 == w.Ui[i]) continue;
                rows.push_back(x2.up_m_row[m2]);
            }
            const uint32_t kind = e->h_kind[v->indiv[i]];
            uint32_t w0 = v->indiv_founder[i] ? CMP_FOUNDER : 0u;
            if (plan.format == 0) w0 |= kind | (kind != EMIS_NONE ? (uint32_t)x2.emis_idx[i] << 3 : 0u);
            else if (kind != EMIS_NONE) w0 |= (uint32_t)x2.emis_idx[i];
            uint32_t extra = 0;
            if (rows.size() > 3) {
                extra = (uint32_t)(pool.size() - (size_t)G.pool_begin);
                if (extra > 0xffffu) plan.fits = false;
                pool.push_back((uint32_t)rows.size() - 3);
                for (size_t j = 3; j < rows.size(); j++) pool.push_back((uint32_t)rows[j]);
                rec_extras = true;
            }
            auto row_or_ones = [&](size_t j) { return j < rows.size() ? (uint32_t)rows[j] : ROW2_ONES; };
            rec.push_back(w0);
            rec.push_back(row_or_ones(0) | row_or_ones(1) << 16);
            rec.push_back(row_or_ones(2) | (extra & 0xffffu) << 16);
        };
        auto push_absent = [&]() { rec.push_back(0); rec.push_back(0); rec.push_back(0); };      // no data, rows 0: ones
        auto push_pairs = [&](int m) {
            const int np = x2.n_chunks[m];
            for (int j = 0; j < np; j += 2) {
                uint32_t wd = (uint32_t)x2.chunk_row[x2.chunk_off[m] + j];
                if (j + 1 < np) wd |= (uint32_t)x2.chunk_row[x2.chunk_off[m] + j + 1] << 16;
                rec.push_back(wd);
            }
        };
        auto add_task = [&](int level) {
            const uint32_t kind = rec[0] & 0xffu;
            if (rec_extras || (kind != T2_ROOT && (rec[0] & (1u << 8)))) rec[0] |= T2_GENERAL;     // selfing, or an operand with extra rows
            rec_extras = false;
            plan.tasks.push_back(Plan2Task{gid, level, (uint32_t)(pool.size() - (size_t)G.pool_begin), (int)rec.size()});
            pool.insert(pool.end(), rec.begin(), rec.end());
            G.task_end++;
        };
        auto row16 = [&](int row) { return row < 0 ? ROW2_TRASH : (uint32_t)row; };
        // kid entry of the down band: cmp, stub row, dn row
        auto push_kid_down = [&](int i, int m) {
            push_cmp(i, m, -1);
            rec.push_back((uint32_t)e->stub_row(chain, v->indiv[i]));
            rec.push_back(row16(x2.dn_row[i]));
        };

        // levels: >= 0 = up band; -1 = first level of the down band; -(2 + k) = down-band level k
        rec.clear();
        rec.push_back(T2_ROOT | (up_only ? 0u : 1u) << 8);
        rec.push_back(up_only ? 0u : (uint32_t)e->stub_row(chain, v->indiv[r]));
        rec.push_back((uint32_t)(out_base + v->indiv_cc[r]));
        rec.push_back(up_only ? (uint32_t)G.n_rows++ : 0u);
        push_cmp(r, -1, -1);
        add_task(-1);
        root_off = plan.tasks.back().off; root_len = plan.tasks.back().cost;

        for (size_t b = mb; b < me; b++) {
            const int m = w.bfs_m[b], u = w.um[m];
            const uint32_t flags = (v->marr_selfing[m] ? 1u : 0u) | (uint32_t)w.urole[m] << 1;
            int pa = -1, ma = -1;
            for (int k = w.mem_off[m]; k < w.mem_off[m + 1]; k++) {
                if (w.mem_i[k] == u && w.mem_role[k] == w.urole[m]) continue;
                if (w.mem_role[k] == 0) pa = w.mem_i[k]; else if (w.mem_role[k] == 1) ma = w.mem_i[k];
            }
            const int k0 = x2.kid_off[m], nki = x2.n_inline[m], np = x2.n_chunks[m];
            // chunk products (up band)
            for (int c = x2.chunk_off[m]; c < x2.chunk_off[m] + np; c++) {
                rec.clear();
                rec.push_back(T2_KFOLD | (uint32_t)x2.chunk_n[c] << 16);
                rec.push_back((uint32_t)x2.chunk_row[c]);
                for (int j = 0; j < x2.chunk_n[c]; j++) push_cmp(x2.kid_x[x2.chunk_first[c] + j], m, -1);
                add_task(x2.chunk_level[c]);
            }
            rec.clear();
            rec.push_back(T2_FAMUP | flags << 8 | (uint32_t)nki << 16 | (uint32_t)np << 24);
            rec.push_back((uint32_t)x2.up_m_row[m]);
            if (pa >= 0) push_cmp(pa, m, -1); else push_absent();
            if (ma >= 0) push_cmp(ma, m, -1); else push_absent();
            for (int j = 0; j < nki; j++) push_cmp(x2.kid_x[k0 + j], m, -1);
            push_pairs(m);
            add_task(x2.lvl2[m]);
            if (site) { x2.famup_off[m] = plan.tasks.back().off; x2.famup_len[m] = plan.tasks.back().cost; }
            if (up_only) continue;

            const int l = x2.depth2[m];
            const int u_dn = w.Ui[u] >= 0 ? x2.dn_row[u] : -1;
            rec.clear();
            rec.push_back(T2_FAMDOWN | flags << 8 | (pa >= 0 ? T2_HAS_A : 0u) | (ma >= 0 ? T2_HAS_B : 0u) | (uint32_t)nki << 16 | (uint32_t)np << 24);
            rec.push_back((uint32_t)e->prong_row(chain, v->marr[m]));
            push_cmp(u, m, u_dn);
            if (pa >= 0) { push_cmp(pa, m, -1); rec.push_back((uint32_t)e->stub_row(chain, v->indiv[pa])); } else { push_absent(); rec.push_back(0); }
            if (ma >= 0) { push_cmp(ma, m, -1); rec.push_back((uint32_t)e->stub_row(chain, v->indiv[ma])); } else { push_absent(); rec.push_back(0); }
            rec.push_back(row16(pa >= 0 ? x2.dn_row[pa] : -1) | row16(ma >= 0 ? x2.dn_row[ma] : -1) << 16);
            for (int j = 0; j < nki; j++) push_kid_down(x2.kid_x[k0 + j], m);
            push_pairs(m);
            add_task(-(2 + l));
            for (int ci = 0; ci < np; ci++) {
                const int c = x2.chunk_off[m] + ci;
                rec.clear();
                rec.push_back(T2_KIDCHUNK | flags << 8 | (pa >= 0 ? T2_HAS_A : 0u) | (ma >= 0 ? T2_HAS_B : 0u) | (uint32_t)x2.chunk_n[c] << 16 | (uint32_t)np << 24);
                rec.push_back((uint32_t)nki | (uint32_t)ci << 8);
                push_cmp(u, m, u_dn);
                if (pa >= 0) push_cmp(pa, m, -1); else push_absent();
                if (ma >= 0) push_cmp(ma, m, -1); else push_absent();
                for (int j = 0; j < x2.chunk_n[c]; j++) push_kid_down(x2.kid_x[x2.chunk_first[c] + j], m);
                for (int j = 0; j < nki; j++) push_cmp(x2.kid_x[k0 + j], m, -1);
                push_pairs(m);
                add_task(-(2 + l));
            }
        }
        if (site && site->var.size() > 1) {
            // the other variants of a swap site: three family tasks and a root each, put together from
            // the operands of the first variant's records (which is the view itself)
            const int mZ = nM - 2, mX = nM - 1;
            int mx = -1;
            for (int j = 0; j < nM - 2; j++) if (v->marr[j] == site->marr_x_slot) mx = j;
            if (!up_only || w.roots.size() != 1 || nM < 3 || mx < 0 || w.um[mX] != r || w.um[mx] != r || w.urole[mZ] != 2 || w.urole[mX] != 2 ||
                w.Ui[w.um[mZ]] != mx || w.urole[mx] > 1 || v->marr_selfing[mZ] || v->marr_selfing[mX] || v->marr_selfing[mx])
                return fail(PF_EINVAL, "internal: malformed swap site");
            const size_t pb = (size_t)G.pool_begin;
            auto grab = [&](uint32_t off, int len) { return std::vector<uint32_t>(pool.begin() + pb + off, pool.begin() + pb + off + len); };
            const std::vector<uint32_t> recZ = grab(x2.famup_off[mZ], x2.famup_len[mZ]), recX = grab(x2.famup_off[mX], x2.famup_len[mX]),
                                        recx = grab(x2.famup_off[mx], x2.famup_len[mx]), recR = grab(root_off, root_len);
            const SwapVariantSpec &v0 = site->var[0];
            auto parent_cmp = [&](int slot) -> const uint32_t * {
                if (slot == v0.za) return &recZ[2];
                if (slot == v0.zb) return &recZ[2 + CMP_WORDS];
                if (slot == v0.xa) return &recX[2];
                if (slot == v0.xb) return &recX[2 + CMP_WORDS];
                return nullptr;
            };
            bool bad = false;
            // replaces row `from` by `to` among the rows of the cmp at rec[pos..pos+2]
            auto swap_row = [&](std::vector<uint32_t> &rc2, int pos, uint32_t from, uint32_t to) {
                uint32_t &w1 = rc2[pos + 1], &w2 = rc2[pos + 2];
                if ((w1 & 0xffffu) == from) { w1 = (w1 & 0xffff0000u) | to; return; }
                if ((w1 >> 16) == from) { w1 = (w1 & 0xffffu) | to << 16; return; }
                if ((w2 & 0xffffu) == from) { w2 = (w2 & 0xffff0000u) | to; return; }
                const uint32_t extra = w2 >> 16;
                if (!extra) { bad = true; return; }
                const uint32_t n_more = pool[pb + extra];
                const uint32_t at = (uint32_t)(pool.size() - pb);
                if (at > 0xffffu) { plan.fits = false; return; }
                bool found = false;
                pool.push_back(n_more);
                for (uint32_t j = 0; j < n_more; j++) {
                    uint32_t row = pool[pb + extra + 1 + j];
                    if (row == from) { row = to; found = true; }
                    pool.push_back(row);
                }
                if (!found) bad = true;
                w2 = (w2 & 0xffffu) | at << 16;
            };
            const uint32_t gen = (recZ[0] | recX[0]) & T2_GENERAL;
            const int Lf = std::max(x2.lvl2[mZ], x2.lvl2[mX]), Lx = std::max(x2.lvl2[mx], Lf + 1);
            max_up_levels = std::max(max_up_levels, Lx + 1);
            const int kids_at = 2 + 2 * CMP_WORDS;
            for (size_t vi = 1; vi < site->var.size(); vi++) {
                const SwapVariantSpec &sv = site->var[vi];
                const bool same = (sv.move != 0) == (v0.move != 0);
                const std::vector<uint32_t> &kz = same ? recZ : recX, &kx = same ? recX : recZ;      // whose sibship goes where
                const uint32_t rowZ = (uint32_t)G.n_rows++, rowX = (uint32_t)G.n_rows++, rowx = (uint32_t)G.n_rows++;
                const uint32_t *cza = parent_cmp(sv.za), *czb = parent_cmp(sv.zb), *cxa = parent_cmp(sv.xa), *cxb = parent_cmp(sv.xb);
                if (!cza || !czb || !cxa || !cxb) return fail(PF_EINVAL, "internal: swap variant names an unknown parent");
                auto family = [&](const std::vector<uint32_t> &kids, uint32_t dest, const uint32_t *ca, const uint32_t *cb) {
                    rec.clear();
                    rec.push_back(T2_FAMUP | (2u << 1) << 8 | (kids[0] & 0xffff0000u) | gen);      // the up individual is a kid; nk and np of the sibship
                    rec.push_back(dest);
                    rec.insert(rec.end(), ca, ca + CMP_WORDS);
                    rec.insert(rec.end(), cb, cb + CMP_WORDS);
                    rec.insert(rec.end(), kids.begin() + kids_at, kids.end());
                    add_task(Lf);
                };
                family(kz, rowZ, cza, czb);
                family(kx, rowX, cxa, cxb);
                {
                    std::vector<uint32_t> r2 = recx;
                    r2[1] = rowx;
                    swap_row(r2, w.urole[mx] == 0 ? 2 + CMP_WORDS : 2, recZ[1], rowZ);          // z is the parent that is not x
                    rec = r2;
                    add_task(Lx);
                }
                {
                    std::vector<uint32_t> r2 = recR;
                    r2[2] = (uint32_t)(out_base + (int)vi);
                    r2[3] = ROW2_TRASH;
                    swap_row(r2, 4, recX[1], rowX);
                    swap_row(r2, 4, recx[1], rowx);
                    rec = r2;
                    add_task(-1);
                }
                if (bad) return fail(PF_EINVAL, "internal: swap site rows not found");
            }
        }
        G.pool_len = (int)pool.size() - G.pool_begin;
        hp_mark(4);
        if (G.n_rows >= 0xffff) plan.fits = false;
        for (size_t b = mb; b < me; b++) if (x2.n_chunks[w.bfs_m[b]] > 255) plan.fits = false;
    }
    return PF_OK;
}

// Plans the views with the second-generation builder and launches sweep2_kernel. Returns
// PF2_FALLBACK (nothing launched) when some group does not fit the kernel's shared-memory or field
// limits: the caller then runs the first-generation kernel, which can spill rows.
constexpr int PF2_FALLBACK = 1000;

static int run_sweeps2(pf_engine *e, int n, const int32_t *chains, const pf_graph_view *views, double *out_dev, int *n_out_total, bool up_only, const SwapSite *sites = nullptr)
{
    const int n_tiles = (e->Sl + TILE_LOCI - 1) / TILE_LOCI;
    const double t_begin = e->host_prof ? host_now() : 0.0;
    static thread_local Plan2 plan_storage;
    static thread_local Plan dummy;
    Plan2 &plan = plan_storage;
    plan.reset();
    long long est_tasks = 0, est_comp = 0;
    for (int i = 0; i < n; i++) { est_tasks += (long long)std::max(views[i].n_cc, 0) + 2LL * std::max(views[i].n_marr, 0); est_comp += std::max(views[i].n_cc, 0); }
    {
        // enough groups that (tiles x groups) fills the GPU, but no more than 256 tasks per group
        const long long want_groups = std::max<long long>((592 + n_tiles - 1) / n_tiles, (est_tasks + 255) / 256);
        const long long Gn = std::max<long long>(1, std::min<long long>(std::min<long long>(want_groups, est_comp), 65535));
        plan.target_tasks_per_group = (int)std::max<long long>(1, (est_tasks + Gn - 1) / Gn);
    }
    int sm_count = e->sm_count;
    if (sm_count <= 0) { cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, e->cfg.device); e->sm_count = sm_count = sm_count > 0 ? sm_count : 148; }
    // genotype format of the launch (pf_sweep2.cu: Fmt2): all hard calls, or all 16-bit numerators over
    // one denominator (individuals without data fit both), else generic
    uint32_t dec16_denom = 0;
    {
        bool hard = false, dec16 = false, other = false, denoms_differ = false;
        for (int i = 0; i < n; i++) {
            const pf_graph_view &v = views[i];
            if (v.n_indiv <= 0 || !v.indiv) continue;
            for (int k = 0; k < v.n_indiv; k++) {
                const int slot = v.indiv[k];
                if (slot < 0 || slot >= e->cfg.cap_indiv) continue;          // reported by plan_add_view
                const uint8_t kd = e->h_kind[slot];
                if (kd == EMIS_HARD) hard = true;
                else if (kd == EMIS_DEC16) {
                    if (dec16 && e->h_denom[slot] != dec16_denom) denoms_differ = true;
                    dec16 = true; dec16_denom = e->h_denom[slot];
                } else if (kd != EMIS_NONE) other = true;
            }
        }
        if (!other && hard && !dec16) plan.format = 1;
        else if (!other && dec16 && !hard && !denoms_differ && dec16_denom <= 0xffffu) plan.format = 2;
        else if (!other && !hard && !dec16) plan.format = 1;               // nobody carries data
    }
    int max_up = 0, out_base = 0;
    if (n >= 16) {
        const int T = (int)std::min<unsigned>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())), (unsigned)n / 4u);
        std::vector<int> bases(n + 1, 0);
        for (int i = 0; i < n; i++) bases[i + 1] = bases[i] + std::max(views[i].n_cc, 0);
        struct Part { Plan2 plan; int max_up = 0, rc = PF_OK; std::string err; };
        std::vector<Part> parts(T);
        std::vector<std::thread> th;
        const int target = plan.target_tasks_per_group;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                static thread_local std::vector<int> tl_loc;
                static thread_local Plan tl_dummy;
                if ((int)tl_loc.size() < e->cfg.cap_indiv) tl_loc.assign(e->cfg.cap_indiv, -1);
                Part &P = parts[t];
                P.plan.target_tasks_per_group = target;
                P.plan.format = plan.format;
                const int i0 = (int)((long long)n * t / T), i1 = (int)((long long)n * (t + 1) / T);
                for (int i = i0; i < i1 && P.rc == PF_OK; i++) {
                    P.rc = plan_add_view(e, &views[i], chains ? chains[i] : 0, bases[i], tl_dummy, P.max_up, up_only, SWEEP_LATENCY, tl_loc.data(), &P.plan, sites ? &sites[i] : nullptr);
                    if (P.rc) P.err = g_err;
                }
            });
        for (auto &x : th) x.join();
        for (int t = 0; t < T; t++) {
            Part &P = parts[t];
            if (P.rc) return fail(P.rc, "%s", P.err.c_str());
            const int gb = (int)plan.groups.size(), pb = (int)plan.pool.size(), eb = (int)plan.emis.size();
            for (auto &tk : P.plan.tasks) plan.tasks.push_back(Plan2Task{tk.group + gb, tk.level, tk.off, tk.cost});
            for (auto &sg : P.plan.groups) { Sweep2Group g2 = sg; g2.pool_begin += pb; g2.emis_begin += eb; plan.groups.push_back(g2); }
            plan.pool.insert(plan.pool.end(), P.plan.pool.begin(), P.plan.pool.end());
            plan.emis.insert(plan.emis.end(), P.plan.emis.begin(), P.plan.emis.end());
            plan.stage.insert(plan.stage.end(), P.plan.stage.begin(), P.plan.stage.end());
            plan.fits = plan.fits && P.plan.fits;
            max_up = std::max(max_up, P.max_up);
        }
        out_base = bases[n];
    } else {
        for (int i = 0; i < n; i++) {
            int rc = plan_add_view(e, &views[i], chains ? chains[i] : 0, out_base, dummy, max_up, up_only, SWEEP_LATENCY, nullptr, &plan, sites ? &sites[i] : nullptr);
            if (rc) return rc;
            out_base += views[i].n_cc;
        }
    }
    *n_out_total = out_base;
    if (plan.tasks.empty()) {
        if (out_base > 0) CU(cudaMemsetAsync(out_dev, 0, (size_t)out_base * sizeof(double), e->stream));
        return PF_OK;
    }
    if (!plan.fits) return PF2_FALLBACK;

    int n_levels = 1;
    for (auto &t : plan.tasks) {
        if (t.level == -1) t.level = max_up;
        else if (t.level <= -2) t.level = max_up + (-t.level - 2);
        n_levels = std::max(n_levels, t.level + 1);
    }
    int G = (int)plan.groups.size();
    if (G > 65535) return PF2_FALLBACK;
    int smem_bytes = 0;
    for (int g = 0; g < G; g++) {
        Sweep2Group sg = plan.groups[g];
        sg.task_begin = 0;                      // task_end still counts the group's tasks
        smem_bytes = std::max(smem_bytes, sweep2_smem_layout(sg, n_levels).total);
    }
    if (smem_bytes > SWEEP2_SMEM_MAX) return PF2_FALLBACK;
    // heaviest groups first (see run_sweeps)
    if (G > 1 && (long long)G * n_tiles > sm_count) {
        std::vector<int> &perm = g_work.cost_of, &inv = g_work.start;
        perm.resize(G); inv.resize(G);
        for (int g = 0; g < G; g++) perm[g] = g;
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return plan.groups[a].task_end > plan.groups[b].task_end; });
        bool moved = false;
        for (int k = 0; k < G; k++) { inv[perm[k]] = k; moved = moved || perm[k] != k; }
        if (moved) {
            std::vector<Sweep2Group> sorted(G);
            for (int k = 0; k < G; k++) sorted[k] = plan.groups[perm[k]];
            plan.groups.swap(sorted);
            for (auto &t : plan.tasks) t.group = inv[t.group];
        }
    }
    // counting sort by (group, level)
    std::vector<int> &level_off = g_work.level_off, &start = g_work.start;
    level_off.assign((size_t)G * (n_levels + 1) + 1, 0);
    for (auto &t : plan.tasks) level_off[(size_t)t.group * (n_levels + 1) + t.level + 1]++;
    start.assign((size_t)G * n_levels, 0);
    {
        int base = 0;
        for (int g = 0; g < G; g++) {
            int *lo = &level_off[(size_t)g * (n_levels + 1)];
            int run = base;
            for (int l = 0; l < n_levels; l++) { const int c = lo[l + 1]; start[(size_t)g * n_levels + l] = run; lo[l] = run; run += c; }
            lo[n_levels] = run;
            plan.groups[g].task_begin = base;
            plan.groups[g].task_end = run;
            base = run;
        }
    }
    const double t_planned = e->host_prof ? host_now() : 0.0;
    const size_t bytes_tasks = plan.tasks.size() * sizeof(Sweep2Task);
    const size_t bytes_pool = plan.pool.size() * sizeof(uint32_t);
    const size_t bytes_loff = (size_t)G * (n_levels + 1) * sizeof(int);
    const size_t bytes_emis = plan.emis.size() * sizeof(Emis2);
    const size_t bytes_stage = plan.stage.size() * sizeof(Stage2);
    const size_t bytes_groups = (size_t)G * sizeof(Sweep2Group);
    const size_t off_pool = (bytes_tasks + 15) & ~(size_t)15;
    const size_t off_loff = (off_pool + bytes_pool + 15) & ~(size_t)15;
    const size_t off_emis = (off_loff + bytes_loff + 15) & ~(size_t)15;
    const size_t off_stage = (off_emis + bytes_emis + 15) & ~(size_t)15;
    const size_t off_groups = (off_stage + bytes_stage + 15) & ~(size_t)15;
    const size_t blob = off_groups + bytes_groups;

    void *hp = nullptr; int slot = 0;
    CU(e->staging.get(blob, &hp, &slot));
    Sweep2Task *ht = static_cast<Sweep2Task *>(hp);
    {
        std::vector<int> &cost = g_work.cost_of;
        cost.resize(plan.tasks.size());
        for (auto &t : plan.tasks) {
            const int pos = start[(size_t)t.group * n_levels + t.level]++;
            ht[pos].off = t.off; cost[pos] = t.cost;
        }
        // a level with more tasks than warps runs in rounds: longest first
        std::vector<std::pair<int, SweepTask>> &tmp = g_work.sort_tmp;
        for (int g = 0; g < G; g++)
            for (int l = 0; l < n_levels; l++) {
                const int b0 = level_off[(size_t)g * (n_levels + 1) + l], b1 = level_off[(size_t)g * (n_levels + 1) + l + 1];
                if (b1 - b0 <= SWEEP2_WARPS_SMALL) continue;
                tmp.clear();
                for (int k = b0; k < b1; k++) tmp.push_back({cost[k], SweepTask{0, (int)ht[k].off}});
                std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, SweepTask> &x, const std::pair<int, SweepTask> &y) { return x.first > y.first; });
                for (int k = b0; k < b1; k++) ht[k].off = (uint32_t)tmp[k - b0].second.off;
            }
    }
    memcpy(static_cast<char *>(hp) + off_pool, plan.pool.data(), bytes_pool);
    memcpy(static_cast<char *>(hp) + off_loff, level_off.data(), bytes_loff);
    memcpy(static_cast<char *>(hp) + off_emis, plan.emis.data(), bytes_emis);
    memcpy(static_cast<char *>(hp) + off_stage, plan.stage.data(), bytes_stage);
    memcpy(static_cast<char *>(hp) + off_groups, plan.groups.data(), bytes_groups);

    CU(e->plan.reserve(blob));
    CU(e->partial.reserve((size_t)n_tiles * std::max(out_base, 1) * sizeof(double)));
    CU(cudaMemcpyAsync(e->plan.p, hp, blob, cudaMemcpyHostToDevice, e->stream)); e->io_h2d += blob;
    CU(e->staging.mark(slot, e->stream));
    // a component label without individuals has no ROOT task: its sum is 0, as in the oracle. Otherwise
    // every partial sum is written by its ROOT task and the memset (a launch on the critical path) is not needed
    {
        int n_roots = 0;
        for (auto &t : plan.tasks) n_roots += (plan.pool[(size_t)plan.groups[t.group].pool_begin + t.off] & 0xffu) == T2_ROOT ? 1 : 0;
        if (n_roots != out_base) CU(cudaMemsetAsync(e->partial.p, 0, (size_t)n_tiles * out_base * sizeof(double), e->stream));
    }

    Sweep2Params P;
    P.D = dev_view(e);
    char *dp = static_cast<char *>(e->plan.p);
    P.tasks = reinterpret_cast<const Sweep2Task *>(dp);
    P.pool = reinterpret_cast<const uint32_t *>(dp + off_pool);
    P.level_off = reinterpret_cast<const int *>(dp + off_loff);
    P.emis = reinterpret_cast<const Emis2 *>(dp + off_emis);
    P.stage = reinterpret_cast<const Stage2 *>(dp + off_stage);
    P.groups = reinterpret_cast<const Sweep2Group *>(dp + off_groups);
    P.n_groups = G; P.n_levels = n_levels;
    P.partial = static_cast<double *>(e->partial.p);
    P.n_out = out_base;
    P.clocks = nullptr; P.clock_group = 0;
    const bool fused_reduce = out_base > 0 && out_base <= 256;
    if (e->sweep_counter.cap == 0) {
        CU(e->sweep_counter.reserve(256));
        CU(cudaMemsetAsync(e->sweep_counter.p, 0, e->sweep_counter.cap, e->stream));
    }
    P.out = fused_reduce ? out_dev : nullptr;
    P.counter = static_cast<unsigned *>(e->sweep_counter.p);
    P.debug_flags = getenv("PF_V2_NOPERSIST") ? 1 : 0;                             // measurement aid: skip the stub / prong stores
    P.format = plan.format; P.dec16_denom = dec16_denom ? dec16_denom : 1u; P.dec16_rcp = 1.0 / (double)P.dec16_denom;
    if (e->sweep_clocks) {
        CU(e->clockbuf.reserve((size_t)(n_levels + 2) * sizeof(long long)));
        P.clocks = static_cast<long long *>(e->clockbuf.p);
        int best = 0;
        for (int g = 0; g < G; g++) if (plan.groups[g].task_end - plan.groups[g].task_begin > plan.groups[best].task_end - plan.groups[best].task_begin) best = g;
        P.clock_group = best;
    }
    {
        int rows = 0;
        for (auto &sg : plan.groups) rows += sg.n_rows;
        const int32_t st[8] = {n_levels, G, (int)plan.tasks.size(), smem_bytes, rows, rows, max_up, n_tiles};
        memcpy(e->last_plan, st, sizeof st);
    }
    e->prof_begin(0);
    launch_sweep2(P, n_tiles, smem_bytes, e->stream);
    e->prof_end();
    e->launches++;
    e->sweeps_v2++;
    CU(cudaGetLastError());
    if (P.clocks) {
        std::vector<long long> hc(n_levels + 1);
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaMemcpy(hc.data(), P.clocks, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const int *lo = &level_off[(size_t)P.clock_group * (n_levels + 1)];
        int n_general = 0, hist[8] = {0, 0, 0, 0, 0, 0, 0, 0}, hg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, self = 0, npn = 0;
        for (auto &t : plan.tasks) {
            const uint32_t w0 = plan.pool[(size_t)plan.groups[t.group].pool_begin + t.off];
            hist[w0 & 7]++;
            if (w0 & T2_GENERAL) { n_general++; hg[w0 & 7]++; if (w0 & 256u) self++; if (w0 >> 24) npn++; }
        }
        fprintf(stderr, "sweep2-general-tasks %d of %zu; by kind %d/%d %d/%d %d/%d %d/%d %d/%d %d/%d; bit8 %d np %d\n", n_general, plan.tasks.size(),
                hg[0], hist[0], hg[1], hist[1], hg[2], hist[2], hg[3], hist[3], hg[4], hist[4], hg[5], hist[5], self, npn);
        fprintf(stderr, "sweep2-clocks fmt %d levels %d group-tasks %d smem %d:", plan.format, n_levels, lo[n_levels] - lo[0], smem_bytes);
        for (int l = 0; l < n_levels; l++) fprintf(stderr, " %d:%lld", lo[l + 1] - lo[l], hc[l + 1] - hc[l]);
        fprintf(stderr, "\n");
    }
    if (e->host_prof && ++e->n_prof_calls > 100) { const double t = host_now(); e->t_plan += t_planned - t_begin; e->t_submit += t - t_planned; e->n_prof++; }
    if (out_base > 0 && !fused_reduce) {
        ReduceParams R{static_cast<const double *>(e->partial.p), out_dev, n_tiles, out_base};
        e->prof_begin(2);
        launch_reduce(R, e->stream);
        e->prof_end();
        e->launches++;
        CU(cudaGetLastError());
    }
    return PF_OK;
}

// Upload the plan and launch sweep + reduce. out_dev receives n_out doubles.
static int run_sweeps(pf_engine *e, int n, const int32_t *chains, const pf_graph_view *views, double *out_dev, int *n_out_total, bool up_only = false)
{
    CU(cudaSetDevice(e->cfg.device));
    int rc = sync_tables(e);
    if (rc) return rc;
    if (e->sweep_v2) {
        rc = run_sweeps2(e, n, chains, views, out_dev, n_out_total, up_only);
        if (rc != PF2_FALLBACK) return rc;
        e->sweeps_v1_fallback++;
    }
    const int n_tiles = (e->Sl + TILE_LOCI - 1) / TILE_LOCI;
    const double t_begin = e->host_prof ? host_now() : 0.0;
    static thread_local Plan plan_storage;
    Plan &plan = plan_storage;
    plan.reset();
    // tasks = one ROOT per component + two per marriage: known before any traversal
    long long est_tasks = 0, est_comp = 0;
    for (int i = 0; i < n; i++) { est_tasks += (long long)std::max(views[i].n_cc, 0) + 2LL * std::max(views[i].n_marr, 0); est_comp += std::max(views[i].n_cc, 0); }
    {
        // enough groups that (tiles x groups) fills the GPU, but no more than 256 tasks per group
        const long long want_groups = std::max<long long>((592 + n_tiles - 1) / n_tiles, (est_tasks + 255) / 256);
        const long long G = std::max<long long>(1, std::min<long long>(std::min<long long>(want_groups, est_comp), 65535));
        plan.target_tasks_per_group = (int)std::max<long long>(1, (est_tasks + G - 1) / G);
    }
    // Shape of the launch (see SweepShape), decided from the views before any traversal: a component
    // is heavy when it has at least 40 individuals and a quarter of the largest one's; a block of a
    // heavy component runs ~100 us whatever the number of tiles, so with more heavy blocks than SMs the
    // throughput shape overlaps two per SM — unless the genotype tiles are large (soft calls:
    // 192+ B per individual): then half the shared memory leaves too few message rows (measured on
    // config D the small shape loses, on config C, hard calls at 8 B per individual, it wins 21 %).
    int sm_count = e->sm_count;
    if (sm_count <= 0) { cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, e->cfg.device); e->sm_count = sm_count = sm_count > 0 ? sm_count : 148; }
    bool throughput_shape = e->force_shape == 2;
    if (e->force_shape == 0) {
        std::vector<int> &cnt = g_work.cnt_a, &bytes = g_work.cnt_b;
        int heavy = 0, biggest = 0, emis_bytes = 0;
        for (int pass = 0; pass < 2; pass++)
            for (int i = 0; i < n; i++) {
                const pf_graph_view &v = views[i];
                if (v.n_cc <= 0 || v.n_indiv <= 0 || !v.indiv || !v.indiv_cc) continue;
                if (pass == 0) {
                    cnt.assign(v.n_cc, 0); bytes.assign(v.n_cc, 0);
                    for (int k = 0; k < v.n_indiv; k++) {
                        const int c = v.indiv_cc[k], slot = v.indiv[k];
                        if (c < 0 || c >= v.n_cc || slot < 0 || slot >= e->cfg.cap_indiv) continue;    // reported by plan_add_view
                        cnt[c]++; bytes[c] += emis_tile_bytes(e->h_kind[slot]);
                    }
                    for (int c = 0; c < v.n_cc; c++) biggest = std::max(biggest, cnt[c]);
                    if (n > 1) continue;        // several views: second pass recounts (rare: batched chains)
                }
                if (pass == 1 && n > 1) {
                    cnt.assign(v.n_cc, 0); bytes.assign(v.n_cc, 0);
                    for (int k = 0; k < v.n_indiv; k++) {
                        const int c = v.indiv_cc[k], slot = v.indiv[k];
                        if (c < 0 || c >= v.n_cc || slot < 0 || slot >= e->cfg.cap_indiv) continue;
                        cnt[c]++; bytes[c] += emis_tile_bytes(e->h_kind[slot]);
                    }
                }
                if (pass == 1 || n == 1)
                    for (int c = 0; c < v.n_cc; c++)
                        if (cnt[c] >= 40 && 4 * cnt[c] >= biggest) { heavy++; emis_bytes = std::max(emis_bytes, bytes[c]); }
                if (n == 1) break;
            }
        throughput_shape = (long long)heavy * n_tiles > sm_count && emis_bytes <= 16 * 1024;
    }
    const SweepShape shape = throughput_shape ? SWEEP_THROUGHPUT : SWEEP_LATENCY;
    int n_levels = 1, G = 0, smem_bytes = 0, scratch_rows = 0, max_up = 0;
    bool all_staged = true;
    std::vector<int> &level_off = g_work.level_off, &start = g_work.start;
    {
        int out_base = 0;
        if (n >= 16) {
            // a batch of chains: the views are planned by a few threads (each with its own work
            // buffers and slot map) and the partial plans are concatenated; 64 plans of ~50 us each
            // built one after the other made the batched sweeps host-bound
            const int T = (int)std::min<unsigned>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())), (unsigned)n / 4u);
            std::vector<int> bases(n + 1, 0);
            for (int i = 0; i < n; i++) bases[i + 1] = bases[i] + std::max(views[i].n_cc, 0);
            struct Part { Plan plan; int max_up = 0, rc = PF_OK; std::string err; };
            std::vector<Part> parts(T);
            std::vector<std::thread> th;
            const int target = plan.target_tasks_per_group;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t]() {
                    static thread_local std::vector<int> tl_loc;
                    if ((int)tl_loc.size() < e->cfg.cap_indiv) tl_loc.assign(e->cfg.cap_indiv, -1);
                    Part &P = parts[t];
                    P.plan.target_tasks_per_group = target;
                    const int i0 = (int)((long long)n * t / T), i1 = (int)((long long)n * (t + 1) / T);
                    for (int i = i0; i < i1 && P.rc == PF_OK; i++) {
                        P.rc = plan_add_view(e, &views[i], chains ? chains[i] : 0, bases[i], P.plan, P.max_up, up_only, shape, tl_loc.data());
                        if (P.rc) P.err = g_err;
                    }
                });
            for (auto &x : th) x.join();
            for (int t = 0; t < T; t++) {
                Part &P = parts[t];
                if (P.rc) return fail(P.rc, "%s", P.err.c_str());
                const int gb = (int)plan.groups.size(), pb = (int)plan.pool.size(), eb = (int)plan.emis.size();
                for (auto &tk : P.plan.tasks) plan.tasks.push_back(PlanTask{tk.group + gb, tk.level, tk.kind, tk.off + pb, tk.cost});
                for (auto &sg : P.plan.groups) { SweepGroup g2 = sg; g2.pool_begin += pb; g2.emis_begin += eb; plan.groups.push_back(g2); }
                plan.pool.insert(plan.pool.end(), P.plan.pool.begin(), P.plan.pool.end());
                plan.emis.insert(plan.emis.end(), P.plan.emis.begin(), P.plan.emis.end());
                max_up = std::max(max_up, P.max_up);
            }
            out_base = bases[n];
        } else {
            for (int i = 0; i < n; i++) {
                rc = plan_add_view(e, &views[i], chains ? chains[i] : 0, out_base, plan, max_up, up_only, shape);
                if (rc) return rc;
                out_base += views[i].n_cc;
            }
        }
        plan.n_out = out_base;
        *n_out_total = out_base;
        if (plan.tasks.empty()) {
            if (out_base > 0) CU(cudaMemsetAsync(out_dev, 0, (size_t)out_base * sizeof(double), e->stream));
            return PF_OK;
        }

        // levels: 0 = leaf caching, up band [1, 1 + max_up), down band starts at 1 + max_up
        n_levels = 1;
        for (auto &t : plan.tasks) {
            if (t.level == -1) t.level = 1 + max_up;
            else if (t.level <= -2) t.level = 1 + max_up + (-t.level - 2);
            n_levels = std::max(n_levels, t.level + 1);
        }
        plan.n_levels = n_levels;
        G = (int)plan.groups.size();
        if (G > 65535) return fail(PF_EINVAL, "too many task groups (%d)", G);
        // heaviest groups first: blockIdx.y is the group, and when the grid is larger than the GPU the
        // blocks of the last groups only start once earlier ones have finished — the long-running
        // blocks must not be the ones that wait
        if (G > 1 && (long long)G * n_tiles > sm_count) {
            std::vector<int> &perm = g_work.cost_of, &inv = g_work.start;
            perm.resize(G); inv.resize(G);
            for (int g = 0; g < G; g++) perm[g] = g;
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return plan.groups[a].task_end > plan.groups[b].task_end; });   // task_end still counts the group's tasks
            bool moved = false;
            for (int k = 0; k < G; k++) { inv[perm[k]] = k; moved = moved || perm[k] != k; }
            if (moved) {
                std::vector<SweepGroup> sorted(G);
                for (int k = 0; k < G; k++) sorted[k] = plan.groups[perm[k]];
                plan.groups.swap(sorted);
                for (auto &t : plan.tasks) t.group = inv[t.group];
            }
        }

        // counting sort by (group, level); groups own contiguous task ranges in emission order already
        level_off.assign((size_t)G * (n_levels + 1) + 1, 0);
        for (auto &t : plan.tasks) level_off[(size_t)t.group * (n_levels + 1) + t.level + 1]++;
        start.assign((size_t)G * n_levels, 0);
        {
            int base = 0;
            for (int g = 0; g < G; g++) {
                int *lo = &level_off[(size_t)g * (n_levels + 1)];
                int run = base;
                for (int l = 0; l < n_levels; l++) { const int c = lo[l + 1]; start[(size_t)g * n_levels + l] = run; lo[l] = run; run += c; }
                lo[n_levels] = run;
                plan.groups[g].task_begin = base;
                plan.groups[g].task_end = run;
                base = run;
            }
        }
        // shared-memory budget per group: stage the plan slice if it is small, then as many transient
        // rows as fit; the launch uses the largest layout
        smem_bytes = 0; scratch_rows = 0; all_staged = true;
        for (int g = 0; g < G; g++) {
            const int slice0 = (plan.groups[g].task_end - plan.groups[g].task_begin) * 8 + plan.groups[g].pool_len * 4 + plan.groups[g].emis_count * (int)sizeof(EmisEnt);
            if (slice0 > shape.plan_smem_max) all_staged = false;
        }
        for (int g = 0; g < G; g++) {
            SweepGroup &sg = plan.groups[g];
            sg.staged = all_staged ? 1 : 0;
            sg.n_rows_smem = 0;
            const SweepSmem L0 = sweep_smem_layout(sg, n_levels);
            if (L0.total > shape.smem_max) return fail(PF_EINVAL, "sweep plan too deep for shared memory (%d levels)", n_levels);
            sg.n_rows_smem = std::min(sg.n_rows, (shape.smem_max - L0.total) / SMEM_ROW_BYTES);
            sg.scratch_base = scratch_rows;
            scratch_rows += sg.n_rows;
            smem_bytes = std::max(smem_bytes, sweep_smem_layout(sg, n_levels).total);
        }
    }
    const double t_planned = e->host_prof ? host_now() : 0.0;
    const size_t bytes_tasks = plan.tasks.size() * sizeof(SweepTask);
    const size_t bytes_pool = plan.pool.size() * sizeof(int);
    const size_t bytes_loff = (size_t)G * (n_levels + 1) * sizeof(int);
    const size_t bytes_emis = plan.emis.size() * sizeof(EmisEnt);
    const size_t bytes_groups = (size_t)G * sizeof(SweepGroup);
    const size_t off_pool = (bytes_tasks + 15) & ~(size_t)15;
    const size_t off_loff = (off_pool + bytes_pool + 15) & ~(size_t)15;
    const size_t off_emis = (off_loff + bytes_loff + 15) & ~(size_t)15;
    const size_t off_groups = (off_emis + bytes_emis + 15) & ~(size_t)15;
    const size_t blob = off_groups + bytes_groups;

    void *hp = nullptr; int slot = 0;
    CU(e->staging.get(blob, &hp, &slot));
    SweepTask *ht = static_cast<SweepTask *>(hp);
    {
        std::vector<int> &cost = g_work.cost_of;
        cost.resize(plan.tasks.size());
        for (auto &t : plan.tasks) {
            const int pos = start[(size_t)t.group * n_levels + t.level]++;
            ht[pos].kind = t.kind; ht[pos].off = t.off; cost[pos] = t.cost;
        }
        // a level with more tasks than warps runs in rounds (warp w takes tasks w, w + warps, ...):
        // longest first, so that the last round holds the short ones
        std::vector<std::pair<int, SweepTask>> &tmp = g_work.sort_tmp;
        for (int g = 0; g < G; g++)
            for (int l = 0; l < n_levels; l++) {
                const int b0 = level_off[(size_t)g * (n_levels + 1) + l], b1 = level_off[(size_t)g * (n_levels + 1) + l + 1];
                if (b1 - b0 <= shape.warps) continue;
                tmp.clear();
                for (int k = b0; k < b1; k++) tmp.push_back({cost[k], ht[k]});
                std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, SweepTask> &x, const std::pair<int, SweepTask> &y) { return x.first > y.first; });
                for (int k = b0; k < b1; k++) ht[k] = tmp[k - b0].second;
            }
    }
    memcpy(static_cast<char *>(hp) + off_pool, plan.pool.data(), bytes_pool);
    memcpy(static_cast<char *>(hp) + off_loff, level_off.data(), bytes_loff);
    memcpy(static_cast<char *>(hp) + off_emis, plan.emis.data(), bytes_emis);
    memcpy(static_cast<char *>(hp) + off_groups, plan.groups.data(), bytes_groups);

    CU(e->plan.reserve(blob));
    CU(e->scratch.reserve((size_t)std::max(scratch_rows, 1) * e->row_elems() * sizeof(double)));
    CU(e->partial.reserve((size_t)n_tiles * std::max(plan.n_out, 1) * sizeof(double)));
    CU(cudaMemcpyAsync(e->plan.p, hp, blob, cudaMemcpyHostToDevice, e->stream)); e->io_h2d += blob;
    CU(e->staging.mark(slot, e->stream));
    // a component label without individuals has no ROOT task: its sum is 0, as in the oracle
    CU(cudaMemsetAsync(e->partial.p, 0, (size_t)n_tiles * plan.n_out * sizeof(double), e->stream));

    SweepParams P;
    P.D = dev_view(e);
    char *dp = static_cast<char *>(e->plan.p);
    P.tasks = reinterpret_cast<const SweepTask *>(dp);
    P.pool = reinterpret_cast<const int *>(dp + off_pool);
    P.level_off = reinterpret_cast<const int *>(dp + off_loff);
    P.emis = reinterpret_cast<const EmisEnt *>(dp + off_emis);
    P.groups = reinterpret_cast<const SweepGroup *>(dp + off_groups);
    P.n_groups = G; P.n_levels = n_levels;
    P.partial = static_cast<double *>(e->partial.p);
    P.n_out = plan.n_out;
    P.clocks = nullptr; P.clock_group = 0;
    // few outputs: the sweep's last block reduces them itself; many (wide sweeps): separate launch
    const bool fused_reduce = plan.n_out > 0 && plan.n_out <= 256;
    if (e->sweep_counter.cap == 0) {
        CU(e->sweep_counter.reserve(256));
        CU(cudaMemsetAsync(e->sweep_counter.p, 0, e->sweep_counter.cap, e->stream));
    }
    P.out = fused_reduce ? out_dev : nullptr;
    P.counter = static_cast<unsigned *>(e->sweep_counter.p);
    if (e->sweep_clocks) {
        // measurement aid: per-level clocks of the largest group, printed after the launch
        CU(e->clockbuf.reserve((size_t)(n_levels + 2) * sizeof(long long)));
        P.clocks = static_cast<long long *>(e->clockbuf.p);
        int best = 0;
        for (int g = 0; g < G; g++) if (plan.groups[g].task_end - plan.groups[g].task_begin > plan.groups[best].task_end - plan.groups[best].task_begin) best = g;
        P.clock_group = best;
    }
    {
        int rows = 0, rows_smem = 0;
        for (auto &sg : plan.groups) { rows += sg.n_rows; rows_smem += sg.n_rows_smem; }
        const int32_t st[8] = {n_levels, G, (int)plan.tasks.size(), smem_bytes, rows, rows_smem, max_up, n_tiles};
        memcpy(e->last_plan, st, sizeof st);
    }
    e->prof_begin(0);
    launch_sweep(P, n_tiles, smem_bytes, all_staged, throughput_shape, e->stream);
    e->prof_end();
    e->launches++;
    CU(cudaGetLastError());
    if (P.clocks) {
        std::vector<long long> hc(n_levels + 1);
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaMemcpy(hc.data(), P.clocks, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const int *lo = &level_off[(size_t)P.clock_group * (n_levels + 1)];
        fprintf(stderr, "sweep-clocks levels %d group-tasks %d:", n_levels, lo[n_levels] - lo[0]);
        for (int l = 0; l < n_levels; l++) fprintf(stderr, " %d:%lld", lo[l + 1] - lo[l], hc[l + 1] - hc[l]);
        fprintf(stderr, "\n");
    }
    if (e->host_prof && ++e->n_prof_calls > 100) { const double t = host_now(); e->t_plan += t_planned - t_begin; e->t_submit += t - t_planned; e->n_prof++; }   // the first 100 sweeps are warm-up
    if (plan.n_out > 0 && !fused_reduce) {
        ReduceParams R{static_cast<const double *>(e->partial.p), out_dev, n_tiles, plan.n_out};
        e->prof_begin(2);
        launch_reduce(R, e->stream);
        e->prof_end();
        e->launches++;
        CU(cudaGetLastError());
    }
    return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// Locus shards on several GPUs of one node, one process per GPU. Every rank runs the
// identical call sequence on its own slice of the loci; results that are sums over loci are added
// up across the ranks by xreduce_kernel through peer-mapped mailboxes before they are delivered, so
// a sharded engine answers every call exactly as an unsharded one does.

static constexpr size_t COMM_FLAG_BYTES = 256;
static size_t comm_bytes(int world) { return COMM_FLAG_BYTES + (size_t)2 * world * XREDUCE_MAX * sizeof(double); }

extern "C" int pf_comm_local_handle(pf_engine *e, int32_t world, void *handle_out)
{
    if (!e || !handle_out) return fail(PF_EINVAL, "null argument");
    if (world < 1 || world > XREDUCE_MAX_WORLD) return fail(PF_EINVAL, "world size %d outside [1,%d]", world, XREDUCE_MAX_WORLD);
    if (e->comm.on) return fail(PF_EINVAL, "communicator already connected");
    CU(cudaSetDevice(e->cfg.device));
    if (e->comm.local) { CU(cudaFree(e->comm.local)); e->comm.local = nullptr; }
    CU(cudaMalloc(&e->comm.local, comm_bytes(world)));
    CU(cudaMemset(e->comm.local, 0, comm_bytes(world)));
    e->comm.world = world;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->comm.local));
    static_assert(sizeof(cudaIpcMemHandle_t) == PF_COMM_HANDLE_BYTES, "handle size");
    memcpy(handle_out, &h, sizeof h);
    return PF_OK;
}

extern "C" int pf_comm_connect(pf_engine *e, int32_t rank, int32_t world, const void *handles)
{
    if (!e || !handles) return fail(PF_EINVAL, "null argument");
    if (!e->comm.local || world != e->comm.world) return fail(PF_EINVAL, "pf_comm_local_handle(world = %d) must come first", world);
    if (rank < 0 || rank >= world) return fail(PF_EINVAL, "rank %d outside [0,%d)", rank, world);
    CU(cudaSetDevice(e->cfg.device));
    for (int r = 0; r < world; r++) {
        if (r == rank) { e->comm.peer_base[r] = e->comm.local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * PF_COMM_HANDLE_BYTES, sizeof h);
        const cudaError_t ce = cudaIpcOpenMemHandle(&e->comm.peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) return fail(PF_ECUDA, "cannot map the mailbox of rank %d (needs peer access between the GPUs): %s", r, cudaGetErrorString(ce));
    }
    e->comm.rank = rank; e->comm.on = world > 1;
    return PF_OK;
}

extern "C" int64_t pf_comm_reduce_count(const pf_engine *e) { return e ? e->comm.n_reduce : 0; }

// Asynchronous sums: the cross-GPU sums of the _dev entry points are issued on a second stream (after
// the kernels that produce their input), so the engine's stream goes on with the next sweep while
// the exchange — which is latency, and waits for the slowest rank — is in flight. The caller owns the
// output buffers and must not touch them before pf_comm_fence, which makes the engine's stream wait
// for every sum issued so far. Host-buffer entry points are unaffected (they deliver complete sums).
extern "C" int pf_comm_set_async(pf_engine *e, int32_t enable)
{
    if (!e) return fail(PF_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    if (enable && !e->comm.stream2) {
        CU(cudaStreamCreateWithFlags(&e->comm.stream2, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&e->comm.ev_main, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&e->comm.ev_comm, cudaEventDisableTiming));
    }
    if (!enable && e->comm.pending) { CU(cudaStreamSynchronize(e->comm.stream2)); e->comm.pending = false; }
    e->comm.async = enable != 0;
    return PF_OK;
}

extern "C" int pf_comm_fence(pf_engine *e)
{
    if (!e) return fail(PF_EINVAL, "null engine");
    if (!e->comm.pending) return PF_OK;
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventRecord(e->comm.ev_comm, e->comm.stream2));
    CU(cudaStreamWaitEvent(e->stream, e->comm.ev_comm, 0));
    e->comm.pending = false;
    return PF_OK;
}

// Sum buf[0..n) over the shards, in place; optionally deliver the sums to the mapped host mailbox in
// the same kernel. The exchanges of one engine are totally ordered (their sequence numbers drive the
// mailbox protocol): they all run on stream2 once asynchronous sums are enabled, else on the engine's
// stream.
static int comm_reduce(pf_engine *e, double *buf, int n, double *host_dst = nullptr, unsigned long long *host_flag = nullptr, unsigned long long host_seq = 0, bool may_defer = false)
{
    if (n <= 0) return PF_OK;
    cudaStream_t st = e->stream;
    if (e->comm.stream2 && (e->comm.async || e->comm.pending)) {
        CU(cudaEventRecord(e->comm.ev_main, e->stream));
        CU(cudaStreamWaitEvent(e->comm.stream2, e->comm.ev_main, 0));
        st = e->comm.stream2;
        e->comm.pending = true;
    }
    XReduceParams X{};
    X.src = buf; X.dst = buf; X.n = n; X.rank = e->comm.rank; X.world = e->comm.world; X.seq0 = e->comm.seq;
    for (int r = 0; r < e->comm.world; r++) {
        X.peer_flags[r] = static_cast<unsigned long long *>(e->comm.peer_base[r]);
        X.peer_mail[r] = reinterpret_cast<double *>(static_cast<char *>(e->comm.peer_base[r]) + COMM_FLAG_BYTES);
    }
    X.host_dst = host_dst; X.host_flag = host_flag; X.host_seq = host_seq;
    e->comm.seq += (unsigned long long)((n + XREDUCE_MAX - 1) / XREDUCE_MAX);
    if (st == e->stream) e->prof_begin(2);
    launch_xreduce(X, st);
    if (st == e->stream) e->prof_end();
    e->launches++; e->comm.n_reduce++;
    CU(cudaGetLastError());
    if (st != e->stream && !(may_defer && e->comm.async)) { int rc = pf_comm_fence(e); if (rc) return rc; }
    return PF_OK;
}

static bool ensure_mailbox(pf_engine *e)
{
    if (e->no_mailbox) return false;
    if (!e->h_mail) {
        void *hp = nullptr, *dp = nullptr;
        if (cudaHostAlloc(&hp, (size_t)(pf_engine::MAILBOX_MAX + 8) * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess) {
            cudaGetLastError();
            if (hp) cudaFreeHost(hp);
            e->no_mailbox = true;
            return false;
        }
        e->h_mail = static_cast<unsigned long long *>(hp); e->d_mail = static_cast<unsigned long long *>(dp);
        e->h_mail[0] = 0;
    }
    return true;
}

// Delivery of a call's results to the host, in two halves so that more work can be queued behind the
// kernels of the call before the host starts waiting. The result vector is [own results (n) | results
// deferred by an earlier pf_sweep_eval_swaps (n_def)]: small vectors go through the mapped mailbox
// (written by the eval kernel's last block when run_evals armed it, by the cross-GPU sum on a sharded
// engine, else by publish_kernel), large ones through a copy.
struct Fetch { int mode = 0, n = 0, n_def = 0; unsigned long long seq = 0; };

static int reserve_result(pf_engine *e, int n_own)
{
    CU(e->result.reserve((size_t)std::max(n_own + e->n_deferred, 1) * sizeof(double)));
    return PF_OK;
}

static int fetch_begin(pf_engine *e, int n, Fetch &f)
{
    e->pub_want = 0;
    const unsigned long long armed = e->pub_armed;
    e->pub_armed = 0;
    f.n = n;
    f.n_def = armed ? e->pub_armed_def : e->n_deferred;
    const int tot = n + f.n_def;
    if (tot <= 0) { f.mode = 0; return PF_OK; }
    if (!armed && f.n_def > 0)      // no kernel gathers the two pieces: append the deferred results behind this call's
        CU(cudaMemcpyAsync(static_cast<double *>(e->result.p) + n, e->deferred.p, (size_t)f.n_def * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    if (tot <= pf_engine::MAILBOX_MAX && ensure_mailbox(e)) {
        f.mode = 1;
        if (armed) {                      // the eval kernel's last block delivers: nothing to launch
            if (armed != e->mail_seq) return fail(PF_ECUDA, "internal: mailbox sequence out of order");
            f.seq = armed;
        } else if (e->comm.on) {          // cross-GPU sum and delivery in one kernel
            f.seq = ++e->mail_seq;
            int rc = comm_reduce(e, static_cast<double *>(e->result.p), tot, reinterpret_cast<double *>(e->d_mail + 8), e->d_mail, f.seq);
            if (rc) return rc;
        } else {
            f.seq = ++e->mail_seq;
            launch_publish(static_cast<const double *>(e->result.p), reinterpret_cast<double *>(e->d_mail + 8), tot, e->d_mail, f.seq, e->stream);
            e->launches++;
            CU(cudaGetLastError());
        }
        e->io_d2h += (size_t)tot * sizeof(double) + 8;
        return PF_OK;
    }
    if (armed) return fail(PF_ECUDA, "internal: armed delivery without a mailbox");
    f.mode = 2;
    if (e->comm.on) { int rc = comm_reduce(e, static_cast<double *>(e->result.p), tot); if (rc) return rc; }
    if ((size_t)tot * sizeof(double) > e->h_result_cap) {
        if (e->h_result) CU(cudaFreeHost(e->h_result));
        e->h_result = nullptr; e->h_result_cap = 0;
        const size_t want = (size_t)tot * sizeof(double) * 2 + 4096;
        CU(cudaMallocHost(&e->h_result, want));
        e->h_result_cap = want;
    }
    CU(cudaMemcpyAsync(e->h_result, e->result.p, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (!e->fetch_event) CU(cudaEventCreateWithFlags(&e->fetch_event, cudaEventDisableTiming));
    CU(cudaEventRecord(e->fetch_event, e->stream));
    e->io_d2h += (size_t)tot * sizeof(double);
    return PF_OK;
}

static int fetch_end(pf_engine *e, const Fetch &f, double *host_out, int n1, double *host_out2 = nullptr, int n2 = 0)
{
    const double *vals = nullptr;
    if (f.mode == 0) { CU(cudaStreamSynchronize(e->stream)); return PF_OK; }
    if (f.mode == 1) {
        const double t_w0 = e->host_prof ? host_now() : 0.0;
        volatile unsigned long long *flag = e->h_mail;
        for (unsigned spins = 1; *flag != f.seq; spins++) {
            __builtin_ia32_pause();
            if ((spins & 0x3fff) == 0) {                 // a faulting kernel never raises the flag
                const cudaError_t q = cudaStreamQuery(e->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) return fail(PF_ECUDA, "CUDA error while waiting for results: %s", cudaGetErrorString(q));
                if (q == cudaSuccess && *flag != f.seq) { CU(cudaStreamSynchronize(e->stream)); break; }
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if (*flag != f.seq) {                              // the stream is idle and nobody delivered: never hand out stale values
            e->pubctr_stale = true;
            return fail(PF_ECUDA, "internal: the results of this call were not delivered (mailbox sequence %llu, expected %llu)", (unsigned long long)*flag, f.seq);
        }
        if (e->host_prof && e->n_prof_calls > 100) e->t_wait += host_now() - t_w0;
        vals = reinterpret_cast<const double *>(e->h_mail + 8);
    } else {
        CU(cudaEventSynchronize(e->fetch_event));
        vals = e->h_result;
    }
    if (n1 > 0) memcpy(host_out, vals, (size_t)n1 * sizeof(double));
    if (n2 > 0) memcpy(host_out2, vals + n1, (size_t)n2 * sizeof(double));
    if (f.n_def > 0) {
        e->h_deferred.assign(vals + f.n, vals + f.n + f.n_def);
        e->n_deferred = 0;
    }
    return PF_OK;
}

static int fetch_result(pf_engine *e, double *host_out, int n, double *host_out2 = nullptr, int n2 = 0)
{
    Fetch f;
    int rc = fetch_begin(e, n + n2, f);
    if (rc) return rc;
    return fetch_end(e, f, host_out, n, host_out2, n2);
}

static int total_cc(int n, const pf_graph_view *views)
{
    long long t = 0;
    for (int i = 0; i < n; i++) t += views[i].n_cc > 0 ? views[i].n_cc : 0;
    return (int)t;
}

extern "C" int pf_sweep_batch(pf_engine *e, int32_t n, const int32_t *chains, const pf_graph_view *views, double *ll_cc)
{
    if (!e || !views || n < 0) return fail(PF_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    const int tot = total_cc(n, views);
    if (tot > 0 && !ll_cc) return fail(PF_EINVAL, "null output");
    int n_out = 0;
    int rc = reserve_result(e, tot);
    if (rc == PF_OK) rc = run_sweeps(e, n, chains, views, static_cast<double *>(e->result.p), &n_out);
    if (rc) return rc;
    return fetch_result(e, ll_cc, n_out);
}

extern "C" int pf_sweep_batch_dev(pf_engine *e, int32_t n, const int32_t *chains, const pf_graph_view *views, double *ll_cc_dev)
{
    if (!e || !views || n < 0) return fail(PF_EINVAL, "null argument");
    int n_out = 0;
    int rc = run_sweeps(e, n, chains, views, ll_cc_dev, &n_out);
    if (rc == PF_OK && e->comm.on) rc = comm_reduce(e, ll_cc_dev, n_out, nullptr, nullptr, 0, true);
    return rc;
}

extern "C" int pf_sweep(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc)
{
    return pf_sweep_batch(e, 1, &chain, view, ll_cc);
}

extern "C" int pf_sweep_dev(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc_dev)
{
    if (!e || !view) return fail(PF_EINVAL, "null argument");
    int n_out = 0;
    int rc = run_sweeps(e, 1, &chain, view, ll_cc_dev, &n_out);
    if (rc == PF_OK && e->comm.on) rc = comm_reduce(e, ll_cc_dev, n_out, nullptr, nullptr, 0, true);
    return rc;
}

// forward: defined with the proposal evaluation below
static int run_evals(pf_engine *e, int n, const int32_t *chains, const int32_t *kids, const int32_t *prop_begin,
                     const pf_proposal *props, double *out_dev);

// One Gibbs step's first round trip in a single submission: the sweep of the dirty components and
// the evaluation of the focal individual's single-parent proposals are enqueued back to back and
// the host waits once (the proposal list depends on the graph only, not on the sweep's results).
static int run_swaps(pf_engine *e, int chain, const pf_graph_view *view, int n, const pf_swap *swaps, double *out_dev);

static int sweep_eval_swaps(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc, int32_t kid, int32_t n,
                            const pf_proposal *props, double *lsum, const pf_graph_view *swap_view, int32_t n_swaps, const pf_swap *swaps)
{
    if (!e || !view || n < 0 || n_swaps < 0) return fail(PF_EINVAL, "null argument");
    if ((view->n_cc > 0 && !ll_cc) || (n > 0 && (!props || !lsum)) || (n_swaps > 0 && (!swap_view || !swaps))) return fail(PF_EINVAL, "null argument");
    if (n_swaps > 0 && (e->n_deferred > 0 || !e->h_deferred.empty())) return fail(PF_EINVAL, "the swap results of an earlier pf_sweep_eval_swaps were not collected (pf_deferred_results)");
    CU(cudaSetDevice(e->cfg.device));
    const int n_cc = std::max(view->n_cc, 0);
    int rc = reserve_result(e, n_cc + n);
    if (rc) return rc;
    if (n_swaps > 0) CU(e->deferred.reserve((size_t)n_swaps * sizeof(double)));
    int n_out = 0;
    rc = run_sweeps(e, 1, &chain, view, static_cast<double *>(e->result.p), &n_out);
    if (rc) return rc;
    if (n > 0) {
        const int32_t pb[2] = {0, n};
        e->pub_want = n_out + n;          // the eval kernel delivers [ll_cc | lsum] itself
        rc = run_evals(e, 1, &chain, &kid, pb, props, static_cast<double *>(e->result.p) + n_out);
        e->pub_want = 0;
        if (rc) { e->pub_armed = 0; e->pubctr_stale = true; return rc; }
    }
    Fetch f;
    rc = fetch_begin(e, n_out + n, f);
    if (rc) return rc;
    // the swap proposals are planned while the sweep and the proposal kernel run, and run behind them;
    // their values travel with the results of the next call
    int rc_swaps = PF_OK;
    if (n_swaps > 0) {
        rc_swaps = run_swaps(e, chain, swap_view, n_swaps, swaps, static_cast<double *>(e->deferred.p));
        if (rc_swaps == PF_OK) e->n_deferred = n_swaps;
    }
    rc = fetch_end(e, f, ll_cc, n_out, lsum, n);
    return rc ? rc : rc_swaps;
}

extern "C" int pf_sweep_eval(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc,
                             int32_t kid, int32_t n, const pf_proposal *props, double *lsum)
{
    return sweep_eval_swaps(e, chain, view, ll_cc, kid, n, props, lsum, nullptr, 0, nullptr);
}

extern "C" int pf_sweep_eval_swaps(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc, int32_t kid, int32_t n,
                                   const pf_proposal *props, double *lsum, const pf_graph_view *swap_view, int32_t n_swaps, const pf_swap *swaps)
{
    return sweep_eval_swaps(e, chain, view, ll_cc, kid, n, props, lsum, swap_view, n_swaps, swaps);
}

extern "C" int pf_deferred_results(pf_engine *e, int32_t n, double *lsum_swaps)
{
    if (!e || n < 0 || (n > 0 && !lsum_swaps)) return fail(PF_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    if (e->n_deferred > 0) {                       // no call has carried them yet: deliver them now
        int rc = reserve_result(e, 0);
        if (rc == PF_OK) rc = fetch_result(e, nullptr, 0);
        if (rc) return rc;
    }
    if ((size_t)n != e->h_deferred.size()) return fail(PF_EINVAL, "%d deferred results asked for, %d are held", n, (int)e->h_deferred.size());
    if (n > 0) memcpy(lsum_swaps, e->h_deferred.data(), (size_t)n * sizeof(double));
    e->h_deferred.clear();
    return PF_OK;
}

// the same with the results left on the device: out_dev = [ll_cc (view->n_cc) | lsum (n)], one
// cross-GPU sum over both on a sharded engine
extern "C" int pf_sweep_eval_dev(pf_engine *e, int32_t chain, const pf_graph_view *view, int32_t kid, int32_t n, const pf_proposal *props, double *out_dev)
{
    if (!e || !view || n < 0 || (n > 0 && !props)) return fail(PF_EINVAL, "null argument");
    if ((view->n_cc > 0 || n > 0) && !out_dev) return fail(PF_EINVAL, "null output");
    int n_out = 0;
    int rc = run_sweeps(e, 1, &chain, view, out_dev, &n_out);
    if (rc) return rc;
    if (n > 0) {
        const int32_t pb[2] = {0, n};
        rc = run_evals(e, 1, &chain, &kid, pb, props, out_dev + n_out);
        if (rc) return rc;
    }
    if (e->comm.on) rc = comm_reduce(e, out_dev, n_out + n, nullptr, nullptr, 0, true);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Swap proposals (_ProposeSwapping@0x1000223d0). The reference pushes two "holder" messages along
// the path parents(z) -> z -> marr_x -> x using the per-edge messages it keeps on every marriage;
// the value, sum_s log sum_g h0*h1, is the log partition value of the component that results from
// the rewiring _PickAndUpdate@0x100017f5a-0x100019281 would perform. This engine persists no
// per-edge messages: it builds that hypothetical component and runs the upward half of a sweep
// over it (log Z only, nothing persisted is written), one task group per proposal, one launch.

namespace {
struct HypoView {
    std::vector<int> indiv, icc, marr, mpa, mma, kb, kids;
    std::vector<uint8_t> founder, selfing;
    pf_graph_view v;
    int paz = -1, maz = -1;             // z's parents in the base view
    void finish()
    {
        v.n_indiv = (int)indiv.size(); v.indiv = indiv.data(); v.indiv_cc = icc.data(); v.indiv_founder = founder.data();
        v.n_marr = (int)marr.size(); v.marr = marr.data(); v.marr_pa = mpa.data(); v.marr_ma = mma.data();
        v.marr_selfing = selfing.data(); v.marr_kid_begin = kb.data(); v.marr_kids = kids.data(); v.n_cc = 1;
    }
};
} // namespace

static int build_swap_view(pf_engine *e, const pf_graph_view *b, const pf_swap &w, int q, HypoView &H)
{
    auto label_of = [&](int slot) { return slot >= 0 && slot < e->cfg.cap_indiv && e->loc[slot] >= 0 ? b->indiv_cc[e->loc[slot]] : -1; };
    const int lx = label_of(w.x), lp = label_of(w.prev_pa), lq = label_of(w.prev_ma);
    if (lx < 0 || lp < 0 || lq < 0 || label_of(w.z) != lx) return fail(PF_EINVAL, "swap %d: x, z, prev_pa or prev_ma not in the view (or x and z in different components)", q);
    if (lp == lx || lq == lx) return fail(PF_EINVAL, "swap %d: a previous parent is in the component of x", q);
    if (w.kind < 1 || w.kind > 3 || (w.kind == 3 && !w.move_kids)) return fail(PF_EINVAL, "swap %d: bad kind", q);
    auto in_set = [&](int l) { return l == lx || l == lp || l == lq; };
    H.indiv.clear(); H.icc.clear(); H.founder.clear(); H.marr.clear(); H.mpa.clear(); H.mma.clear(); H.selfing.clear(); H.kids.clear(); H.kb.assign(1, 0);
    // x first: a swap site's tree is rooted at the first individual listed
    H.indiv.push_back(w.x); H.icc.push_back(0); H.founder.push_back(b->indiv_founder[e->loc[w.x]]);
    for (int i = 0; i < b->n_indiv; i++)
        if (in_set(b->indiv_cc[i]) && b->indiv[i] != w.x) { H.indiv.push_back(b->indiv[i]); H.icc.push_back(0); H.founder.push_back(b->indiv_founder[i]); }
    int paz = -1, maz = -1;
    bool have_z = false, have_x = false, have_0 = w.prev_marr == -1;
    std::vector<int> Kz, K0;
    for (int j = 0; j < b->n_marr; j++) {
        const int lj = label_of(b->marr_pa[j]);
        if (lj < 0) return fail(PF_EINVAL, "view not closed: marriage %d", b->marr[j]);
        if (!in_set(lj)) continue;
        const int kb = b->marr_kid_begin[j], ke = b->marr_kid_begin[j + 1];
        const int slot = b->marr[j];
        if (slot == w.marr_z || (slot == w.prev_marr && w.prev_marr != -1)) {
            if (b->marr_selfing[j]) return fail(PF_EINVAL, "swap %d touches a selfing marriage", q);
            if (slot == w.marr_z) {
                paz = b->marr_pa[j]; maz = b->marr_ma[j];
                bool found = false;
                for (int k = kb; k < ke; k++) { if (b->marr_kids[k] == w.z) found = true; else Kz.push_back(b->marr_kids[k]); }
                if (!found) return fail(PF_EINVAL, "swap %d: z is not a kid of marr_z", q);
                have_z = true;
            } else {
                if (b->marr_pa[j] != w.prev_pa || b->marr_ma[j] != w.prev_ma) return fail(PF_EINVAL, "swap %d: prev_marr does not join prev_pa and prev_ma", q);
                for (int k = kb; k < ke; k++) { if (b->marr_kids[k] == w.x) return fail(PF_EINVAL, "swap %d: x is still a kid of prev_marr", q); K0.push_back(b->marr_kids[k]); }
                have_0 = true;
            }
            continue;
        }
        if (slot == w.marr_x) {
            if (b->marr_selfing[j]) return fail(PF_EINVAL, "swap %d touches a selfing marriage", q);
            const bool ok = (b->marr_pa[j] == w.x && b->marr_ma[j] == w.z) || (b->marr_pa[j] == w.z && b->marr_ma[j] == w.x);
            if (!ok) return fail(PF_EINVAL, "swap %d: marr_x does not join x and z", q);
            have_x = true;
        }
        H.marr.push_back(slot); H.mpa.push_back(b->marr_pa[j]); H.mma.push_back(b->marr_ma[j]); H.selfing.push_back(b->marr_selfing[j]);
        H.kids.insert(H.kids.end(), b->marr_kids + kb, b->marr_kids + ke);
        H.kb.push_back((int)H.kids.size());
    }
    if (!have_z || !have_x || !have_0) return fail(PF_EINVAL, "swap %d: marr_z, marr_x or prev_marr not in the view", q);
    // the two rewired families (see pf_swap in the header)
    const int P = w.prev_pa, Q = w.prev_ma;
    int za, zb, xa, xb;
    if (w.kind == 1) { za = P; zb = maz; xa = paz; xb = Q; }
    else if (w.kind == 2) { za = paz; zb = Q; xa = P; xb = maz; }
    else { za = P; zb = Q; xa = paz; xb = maz; }
    const std::vector<int> &kids_z = w.move_kids ? K0 : Kz, &kids_x = w.move_kids ? Kz : K0;
    auto emit = [&](int slot, int pa, int ma, const std::vector<int> &sibs, int extra) {
        H.marr.push_back(slot); H.mpa.push_back(pa); H.mma.push_back(ma); H.selfing.push_back(0);
        H.kids.insert(H.kids.end(), sibs.begin(), sibs.end()); H.kids.push_back(extra);
        H.kb.push_back((int)H.kids.size());
    };
    emit(w.marr_z, za, zb, kids_z, w.z);
    emit(w.prev_marr == -1 ? w.marr_z : w.prev_marr, xa, xb, kids_x, w.x);   // the slot only names a prong row, which an upward sweep never writes
    H.paz = paz; H.maz = maz;
    H.finish();
    return PF_OK;
}

static SwapVariantSpec swap_variant(const pf_swap &w, int paz, int maz)
{
    const int P = w.prev_pa, Q = w.prev_ma;
    if (w.kind == 1) return SwapVariantSpec{P, maz, paz, Q, w.move_kids};
    if (w.kind == 2) return SwapVariantSpec{paz, Q, P, maz, w.move_kids};
    return SwapVariantSpec{P, Q, paz, maz, w.move_kids};
}

// n_sets lists of swap proposals, set s on chain chains[s] (null: chain 0) and view views[s]: one launch
static int run_swap_sets(pf_engine *e, int n_sets, const int32_t *chains_in, const pf_graph_view *views, const int32_t *swap_begin,
                         const pf_swap *swaps, double *out_dev)
{
    if (!e || n_sets < 0 || (n_sets > 0 && (!views || !swap_begin))) return fail(PF_EINVAL, "null argument");
    const int n = n_sets > 0 ? swap_begin[n_sets] : 0;
    if (n < 0 || (n > 0 && !swaps)) return fail(PF_EINVAL, "null argument");
    if (n == 0) return PF_OK;
    std::vector<HypoView> H(n);
    std::vector<pf_graph_view> hv(n);
    std::vector<int32_t> chains(n, 0);
    std::vector<int> site_first;                 // swaps that start a site (a run of swaps on the same two families)
    auto same_site = [](const pf_swap &a, const pf_swap &b) {
        return a.x == b.x && a.z == b.z && a.marr_z == b.marr_z && a.marr_x == b.marr_x && a.prev_pa == b.prev_pa && a.prev_ma == b.prev_ma && a.prev_marr == b.prev_marr;
    };
    const bool use_sites = e->sweep_v2 && e->swap_sites;
    for (int s = 0; s < n_sets; s++) {
        const pf_graph_view *view = &views[s];
        const int chain = chains_in ? chains_in[s] : 0;
        if (swap_begin[s + 1] < swap_begin[s]) return fail(PF_EINVAL, "swap_begin not monotone");
        if (swap_begin[s + 1] == swap_begin[s]) continue;
        if (view->n_indiv < 0 || view->n_marr < 0 || (view->n_indiv > 0 && (!view->indiv || !view->indiv_cc || !view->indiv_founder)) ||
            (view->n_marr > 0 && (!view->marr || !view->marr_pa || !view->marr_ma || !view->marr_selfing || !view->marr_kid_begin)))
            return fail(PF_EINVAL, "malformed view");
        int rc = PF_OK, registered = 0;
        for (int i = 0; i < view->n_indiv; i++) {
            const int slot = view->indiv[i];
            if (slot < 0 || slot >= e->cfg.cap_indiv || e->loc[slot] != -1) { rc = fail(PF_EINVAL, "bad or duplicate individual slot %d in view", slot); break; }
            e->loc[slot] = i; registered = i + 1;
        }
        for (int q = swap_begin[s]; q < swap_begin[s + 1] && rc == PF_OK; q++) {
            chains[q] = chain;
            if (use_sites && q > swap_begin[s] && same_site(swaps[q], swaps[q - 1])) {      // a variant of the site before it: only its arguments are checked
                const pf_swap &sw = swaps[q];
                if (sw.kind < 1 || sw.kind > 3 || (sw.kind == 3 && !sw.move_kids)) rc = fail(PF_EINVAL, "swap %d: bad kind", q);
                continue;
            }
            rc = build_swap_view(e, view, swaps[q], q, H[q]);
            site_first.push_back(q);
        }
        for (int i = 0; i < registered; i++) e->loc[view->indiv[i]] = -1;
        if (rc) return rc;
    }
    int n_out = 0;
    if (use_sites) {
        const int ns = (int)site_first.size();
        std::vector<SwapSite> sites(ns);
        std::vector<pf_graph_view> sv(ns);
        std::vector<int32_t> sc(ns);
        for (int k = 0; k < ns; k++) {
            const int q0 = site_first[k], q1 = k + 1 < ns ? site_first[k + 1] : n;
            sites[k].marr_x_slot = swaps[q0].marr_x;
            for (int q = q0; q < q1; q++) sites[k].var.push_back(swap_variant(swaps[q], H[q0].paz, H[q0].maz));
            sv[k] = H[q0].v;
            sv[k].n_cc = q1 - q0;                 // one output per variant
            sc[k] = chains[q0];
        }
        int rc = sync_tables(e);
        if (rc) return rc;
        rc = run_sweeps2(e, ns, sc.data(), sv.data(), out_dev, &n_out, true, sites.data());
        if (rc != PF2_FALLBACK) return rc;
        e->sweeps_v1_fallback++;
        // does not fit the compiled kernel: one view per swap through the first-generation kernel
        std::vector<char> have(n, 0);
        for (int q : site_first) have[q] = 1;
        // rebuild the views of the variants (their base views are no longer registered: redo set by set)
        for (int s = 0; s < n_sets; s++) {
            const pf_graph_view *view = &views[s];
            bool any = false;
            for (int q = swap_begin[s]; q < swap_begin[s + 1]; q++) any = any || !have[q];
            if (!any) continue;
            for (int i = 0; i < view->n_indiv; i++) e->loc[view->indiv[i]] = i;
            int rc2 = PF_OK;
            for (int q = swap_begin[s]; q < swap_begin[s + 1] && rc2 == PF_OK; q++) if (!have[q]) rc2 = build_swap_view(e, view, swaps[q], q, H[q]);
            for (int i = 0; i < view->n_indiv; i++) e->loc[view->indiv[i]] = -1;
            if (rc2) return rc2;
        }
    }
    for (int q = 0; q < n; q++) hv[q] = H[q].v;
    return run_sweeps(e, n, chains.data(), hv.data(), out_dev, &n_out, true);
}

static int run_swaps(pf_engine *e, int chain, const pf_graph_view *view, int n, const pf_swap *swaps, double *out_dev)
{
    if (!e || !view || n < 0 || (n > 0 && !swaps)) return fail(PF_EINVAL, "null argument");
    const int32_t sb[2] = {0, n};
    return run_swap_sets(e, 1, &chain, view, sb, swaps, out_dev);
}

extern "C" int pf_eval_swaps_batch(pf_engine *e, int32_t n_sets, const int32_t *chains, const pf_graph_view *views,
                                   const int32_t *swap_begin, const pf_swap *swaps, double *lsum)
{
    if (!e || n_sets < 0 || (n_sets > 0 && !swap_begin)) return fail(PF_EINVAL, "null argument");
    const int n = n_sets > 0 ? swap_begin[n_sets] : 0;
    if (n > 0 && !lsum) return fail(PF_EINVAL, "null output");
    CU(cudaSetDevice(e->cfg.device));
    int rc = reserve_result(e, n);
    if (rc == PF_OK) rc = run_swap_sets(e, n_sets, chains, views, swap_begin, swaps, static_cast<double *>(e->result.p));
    if (rc) return rc;
    return fetch_result(e, lsum, n);
}

extern "C" int pf_eval_swaps(pf_engine *e, int32_t chain, const pf_graph_view *view, int32_t n, const pf_swap *swaps, double *lsum)
{
    if (!e) return fail(PF_EINVAL, "null argument");
    if (n > 0 && !lsum) return fail(PF_EINVAL, "null output");
    CU(cudaSetDevice(e->cfg.device));
    int rc = reserve_result(e, n);
    if (rc == PF_OK) rc = run_swaps(e, chain, view, n, swaps, static_cast<double *>(e->result.p));
    if (rc) return rc;
    return fetch_result(e, lsum, n);
}

extern "C" int pf_eval_swaps_dev(pf_engine *e, int32_t chain, const pf_graph_view *view, int32_t n, const pf_swap *swaps, double *lsum_dev)
{
    if (!e) return fail(PF_EINVAL, "null argument");
    if (n > 0 && !lsum_dev) return fail(PF_EINVAL, "null output");
    int rc = run_swaps(e, chain, view, n, swaps, lsum_dev);
    if (rc == PF_OK && e->comm.on) rc = comm_reduce(e, lsum_dev, n, nullptr, nullptr, 0, true);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Proposal evaluation

static int run_evals(pf_engine *e, int n, const int32_t *chains, const int32_t *kids, const int32_t *prop_begin,
                     const pf_proposal *props, double *out_dev)
{
    CU(cudaSetDevice(e->cfg.device));
    const int n_props = prop_begin[n];
    if (n_props <= 0) return PF_OK;
    const double t_e0 = e->host_prof ? host_now() : 0.0;
    // Single-parent proposals (one given parent or none, the other an HWE founder) go eight to a warp
    // through the one-row form, everything else four to a warp: the kernel's proposal array is in that
    // order, every entry carrying its index in the caller's list.
    for (int i = 0; i < n; i++)
        if (prop_begin[i + 1] < prop_begin[i]) return fail(PF_EINVAL, "prop_begin not monotone");
    auto is_single = [&](const pf_proposal &q) { return e->eval_singles && q.kind == PF_PROP_TRIO && (q.pa < 0 || q.ma < 0); };
    // proposals per warp: 4, or fewer for a short list (PF_EVAL_GROUP, measurement aid) so that more warps share it
    const int per_s = n_props <= e->eval_small_props ? std::min(e->eval_small_group, EVAL_GS) : EVAL_GS;
    const int per_g = n_props <= e->eval_small_props ? std::min(e->eval_small_group, EVAL_G) : EVAL_G;
    int n_groups = 0;
    for (int i = 0; i < n; i++) {
        int ns = 0;
        for (int p = prop_begin[i]; p < prop_begin[i + 1]; p++) ns += is_single(props[p]) ? 1 : 0;
        const int ng = prop_begin[i + 1] - prop_begin[i] - ns;
        n_groups += (ns + per_s - 1) / per_s + (ng + per_g - 1) / per_g;
    }
    const size_t bytes_groups = (size_t)n_groups * sizeof(EvalGroup);
    const size_t off_props = (bytes_groups + 15) & ~(size_t)15;
    const size_t blob = off_props + (size_t)n_props * sizeof(EvalProp);
    void *hp = nullptr; int slot = 0;
    CU(e->staging.get(blob, &hp, &slot));
    EvalGroup *hg = static_cast<EvalGroup *>(hp);
    EvalProp *hq = reinterpret_cast<EvalProp *>(static_cast<char *>(hp) + off_props);
    int gi = 0, pos = 0;
    for (int i = 0; i < n; i++) {
        const int chain = chains ? chains[i] : 0, kid = kids[i];
        if (chain < 0 || chain >= e->cfg.n_chains) return fail(PF_EINVAL, "chain %d out of range", chain);
        if (kid < 0 || kid >= e->cfg.cap_indiv) return fail(PF_EINVAL, "kid slot %d out of range", kid);
        for (int pass = 0; pass < 2; pass++) {
            const int first = pos;
            for (int p = prop_begin[i]; p < prop_begin[i + 1]; p++) {
                const pf_proposal &q = props[p];
                if (q.kind != PF_PROP_TRIO && q.kind != PF_PROP_SELFING && q.kind != PF_PROP_PRONG) return fail(PF_EINVAL, "proposal %d: unknown kind %d", p, q.kind);
                if (is_single(q) != (pass == 0)) continue;
                EvalProp d{q.kind, ROW_PRIOR, ROW_PRIOR, p};
                if (q.kind == PF_PROP_TRIO || q.kind == PF_PROP_SELFING) {
                    if (q.pa < -1 || q.pa >= e->cfg.cap_indiv || q.ma < -1 || q.ma >= e->cfg.cap_indiv)
                        return fail(PF_EINVAL, "proposal %d: parent slot out of range", p);
                    if (q.pa >= 0) d.row_p = e->stub_row(chain, q.pa);
                    if (q.kind == PF_PROP_TRIO && q.ma >= 0) d.row_m = e->stub_row(chain, q.ma);
                    if (pass == 0 && q.pa < 0) { d.row_p = d.row_m; d.row_m = ROW_PRIOR; }      // the given parent's row, whichever role
                } else {
                    if (q.marr < 0 || q.marr >= e->cfg.cap_marr) return fail(PF_EINVAL, "proposal %d: marriage slot out of range", p);
                    d.row_p = e->prong_row(chain, q.marr);
                }
                hq[pos++] = d;
            }
            const int per = pass == 0 ? per_s : per_g;
            for (int b = first; b < pos; b += per)
                hg[gi++] = EvalGroup{e->stub_row(chain, kid), b, std::min(per, pos - b), pass == 0 ? 1 : 0};
        }
    }
    // chunking: (chunks x group blocks) fills one wave of resident blocks without spilling into a
    // second, mostly empty one (measured at config C: a 1.45-wave grid 91 us, one wave 80 us); at
    // least 4 lane-iterations per chunk, at most 128 (the running mantissa product needs
    // renormalising only every 512 factors)
    const int group_blocks = (n_groups + EVAL_WARPS - 1) / EVAL_WARPS;
    const int max_chunks = std::max(1, (e->Sl + e->eval_min_chunk - 1) / e->eval_min_chunk);
    if (e->eval_capacity <= 0) e->eval_capacity = eval_resident_blocks();
    const int waves = (group_blocks + e->eval_capacity - 1) / e->eval_capacity;    // > 1 only for huge proposal lists
    int n_chunks = std::max(1, std::min(max_chunks, waves * e->eval_capacity / group_blocks));
    int chunk_loci = ((e->Sl + n_chunks - 1) / n_chunks + 31) / 32 * 32;
    chunk_loci = std::min(chunk_loci, 4096);
    n_chunks = (e->Sl + chunk_loci - 1) / chunk_loci;

    int rc = sync_tables(e);
    if (rc) return rc;
    CU(e->plan.reserve(blob));
    CU(e->partial.reserve((size_t)n_chunks * n_props * sizeof(double)));
    if ((size_t)group_blocks * sizeof(unsigned) > e->counters.cap) {
        CU(e->counters.reserve((size_t)group_blocks * sizeof(unsigned) * 2));
        CU(cudaMemsetAsync(e->counters.p, 0, e->counters.cap, e->stream));
    }
    // the plan is copied to device memory in front of the kernel (PF_EVAL_ZEROCOPY=1: small lists are read
    // by the kernel straight from the pinned staging slot instead)
    // small plans travel in the kernel's parameter space (launch_eval_inline): nothing to copy
    const bool inline_plan = e->eval_inline && blob <= (size_t)EVAL_INLINE_MAX;
    const bool zero_copy = !inline_plan && e->eval_zero_copy && blob <= pf_engine::EVAL_ZERO_COPY_MAX;
    const char *plan_dev = static_cast<const char *>(zero_copy ? hp : e->plan.p);
    if (!zero_copy && !inline_plan) CU(cudaMemcpyAsync(e->plan.p, hp, blob, cudaMemcpyHostToDevice, e->stream));
    e->io_h2d += blob;
    EvalParams P;
    P.D = dev_view(e);
    P.groups = reinterpret_cast<const EvalGroup *>(plan_dev);
    P.props = reinterpret_cast<const EvalProp *>(plan_dev + off_props);
    P.pub_src = nullptr; P.pub_dst = nullptr; P.pub_n = 0; P.pub_flag = nullptr; P.pub_seq = 0; P.pub_counter = nullptr;
    P.pub_src2 = nullptr; P.pub_n2 = 0;
    if (e->pub_want > 0 && e->fused_publish && !e->comm.on && e->pub_want + e->n_deferred <= pf_engine::MAILBOX_MAX && ensure_mailbox(e)) {
        if (!e->pubctr.p) { CU(e->pubctr.reserve(64)); e->pubctr_stale = true; }
        if (e->pubctr_stale) { CU(cudaMemsetAsync(e->pubctr.p, 0, 64, e->stream)); e->pubctr_stale = false; }   // after a failed call the arrival counter may be anywhere
        P.pub_src = static_cast<const double *>(e->result.p);
        P.pub_dst = reinterpret_cast<double *>(e->d_mail + 8);
        P.pub_n = e->pub_want;
        P.pub_src2 = static_cast<const double *>(e->deferred.p); P.pub_n2 = e->pub_armed_def = e->n_deferred;
        P.pub_flag = e->d_mail;
        P.pub_seq = e->pub_armed = ++e->mail_seq;
        P.pub_counter = static_cast<unsigned *>(e->pubctr.p);
    }
    P.n_groups = n_groups; P.n_props = n_props; P.chunk_loci = chunk_loci;
    P.partial = static_cast<double *>(e->partial.p);
    P.out = out_dev;
    P.counters = static_cast<unsigned *>(e->counters.p);
    e->prof_begin(1);
    if (!inline_plan || !launch_eval_inline(P, hp, blob, off_props, n_chunks, e->stream)) launch_eval(P, n_chunks, e->stream);
    e->prof_end();
    e->launches++;
    CU(e->staging.mark(slot, e->stream));
    if (e->host_prof && e->n_prof_calls > 100) e->t_evalprep += host_now() - t_e0;
    CU(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_eval_batch(pf_engine *e, int32_t n, const int32_t *chains, const int32_t *kids,
                             const int32_t *prop_begin, const pf_proposal *props, double *lsum)
{
    if (!e || n < 0 || !kids || !prop_begin) return fail(PF_EINVAL, "null argument");
    const int n_props = prop_begin[n];
    if (n_props > 0 && (!props || !lsum)) return fail(PF_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    int rc = reserve_result(e, n_props);
    if (rc) return rc;
    e->pub_want = n_props;
    rc = run_evals(e, n, chains, kids, prop_begin, props, static_cast<double *>(e->result.p));
    e->pub_want = 0;
    if (rc) { e->pub_armed = 0; e->pubctr_stale = true; return rc; }
    return fetch_result(e, lsum, n_props);
}

extern "C" int pf_eval_batch_dev(pf_engine *e, int32_t n, const int32_t *chains, const int32_t *kids,
                                 const int32_t *prop_begin, const pf_proposal *props, double *lsum_dev)
{
    if (!e || n < 0 || !kids || !prop_begin) return fail(PF_EINVAL, "null argument");
    if (prop_begin[n] > 0 && (!props || !lsum_dev)) return fail(PF_EINVAL, "null argument");
    int rc = run_evals(e, n, chains, kids, prop_begin, props, lsum_dev);
    if (rc == PF_OK && e->comm.on) rc = comm_reduce(e, lsum_dev, prop_begin[n], nullptr, nullptr, 0, true);
    return rc;
}

extern "C" int pf_eval(pf_engine *e, int32_t chain, int32_t kid, int32_t n, const pf_proposal *props, double *lsum)
{
    const int32_t pb[2] = {0, n};
    return pf_eval_batch(e, 1, &chain, &kid, pb, props, lsum);
}

extern "C" int pf_eval_dev(pf_engine *e, int32_t chain, int32_t kid, int32_t n, const pf_proposal *props, double *lsum_dev)
{
    if (!e || n < 0 || (n > 0 && (!props || !lsum_dev))) return fail(PF_EINVAL, "null argument");
    const int32_t pb[2] = {0, n};
    int rc = run_evals(e, 1, &chain, &kid, pb, props, lsum_dev);
    if (rc == PF_OK && e->comm.on) rc = comm_reduce(e, lsum_dev, n, nullptr, nullptr, 0, true);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Prefilter (_RefineParentOffPair@0x100014d20)

// stage the kid / candidate row tables and the bias, fill the kernel parameters
static int prefilter_setup(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids, int32_t n_cands, const int32_t *cands,
                           const double *cand_bias, PrefilterParams &P)
{
    if (chain < 0 || chain >= e->cfg.n_chains) return fail(PF_EINVAL, "chain %d out of range", chain);
    CU(cudaSetDevice(e->cfg.device));
    for (int k = 0; k < n_kids; k++)
        if (kids[k] < 0 || kids[k] >= e->cfg.cap_indiv) return fail(PF_EINVAL, "kid slot %d out of range", kids[k]);
    for (int j = 0; j < n_cands; j++)
        if (cands[j] < 0 || cands[j] >= e->cfg.cap_indiv) return fail(PF_EINVAL, "candidate slot %d out of range", cands[j]);
    int rc = sync_tables(e);
    if (rc) return rc;
    const size_t b_kid = (size_t)n_kids * sizeof(int), b_cand = (size_t)n_cands * sizeof(int);
    const size_t o_cand = (b_kid + 15) & ~(size_t)15, o_bias = (o_cand + b_cand + 15) & ~(size_t)15;
    const size_t blob = o_bias + (size_t)n_cands * sizeof(double);
    void *hp = nullptr; int slot = 0;
    CU(e->staging.get(blob, &hp, &slot));
    char *h = static_cast<char *>(hp);
    int *hk = reinterpret_cast<int *>(h), *hc = reinterpret_cast<int *>(h + o_cand);
    double *hbias = reinterpret_cast<double *>(h + o_bias);
    for (int k = 0; k < n_kids; k++) hk[k] = e->stub_row(chain, kids[k]);
    for (int j = 0; j < n_cands; j++) { hc[j] = e->stub_row(chain, cands[j]); hbias[j] = cand_bias ? cand_bias[j] : 0.0; }
    CU(e->plan.reserve(blob));
    CU(cudaMemcpyAsync(e->plan.p, hp, blob, cudaMemcpyHostToDevice, e->stream)); e->io_h2d += blob;
    CU(e->staging.mark(slot, e->stream));
    P.D = dev_view(e);
    char *d = static_cast<char *>(e->plan.p);
    P.kid_rows = reinterpret_cast<const int *>(d);
    P.cand_rows = reinterpret_cast<const int *>(d + o_cand);
    P.cand_bias = reinterpret_cast<const double *>(d + o_bias);
    P.n_kids = n_kids; P.n_cands = n_cands; P.top_k = 0;
    P.scores = nullptr; P.out_score = nullptr; P.out_idx = nullptr;
    return PF_OK;
}

// rank the dense scores on the device and bring the top_k per kid back to the host
static int prefilter_rank(pf_engine *e, PrefilterParams &P, const int32_t *cands, int32_t top_k, int32_t *out_slot, double *out_score)
{
    const int n_kids = P.n_kids;
    const size_t out_bytes = (size_t)n_kids * top_k * (sizeof(double) + sizeof(int));
    CU(e->result.reserve(out_bytes));
    P.top_k = top_k;
    P.out_score = static_cast<double *>(e->result.p);
    P.out_idx = reinterpret_cast<int *>(static_cast<char *>(e->result.p) + (size_t)n_kids * top_k * sizeof(double));
    e->prof_begin(3);
    e->launches += launch_prefilter_topk(P, e->stream);
    e->prof_end();
    CU(cudaGetLastError());
    std::vector<char> hout(out_bytes);
    CU(cudaMemcpyAsync(hout.data(), e->result.p, out_bytes, cudaMemcpyDeviceToHost, e->stream));
    e->io_d2h += out_bytes;
    CU(cudaStreamSynchronize(e->stream));
    const double *sc = reinterpret_cast<const double *>(hout.data());
    const int *ix = reinterpret_cast<const int *>(hout.data() + (size_t)n_kids * top_k * sizeof(double));
    for (int k = 0; k < n_kids; k++)
        for (int t = 0; t < top_k; t++) {
            const int j = ix[(size_t)k * top_k + t];
            out_slot[(size_t)k * top_k + t] = j < 0 ? -1 : cands[j];
            out_score[(size_t)k * top_k + t] = j < 0 ? -INFINITY : sc[(size_t)k * top_k + t];
        }
    return PF_OK;
}

extern "C" int pf_prefilter(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                            const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                            const double *cand_bias, int32_t top_k, int32_t *out_slot, double *out_score)
{
    if (!e || n_kids < 0 || n_cands < 0 || !kids || !as_father || !out_slot || !out_score) return fail(PF_EINVAL, "null argument");
    if (n_cands > 0 && !cands) return fail(PF_EINVAL, "null candidates");
    if (top_k <= 0 || top_k > PREF_MAX_TOPK) return fail(PF_EINVAL, "top_k must be in [1, %d]", PREF_MAX_TOPK);
    if (e->Sl != e->cfg.n_loci && !e->comm.on) return fail(PF_EINVAL, "pf_prefilter on a locus shard needs the communicator (pf_comm_connect): ranking requires the scores summed over all loci");
    if (n_kids == 0) return PF_OK;
    PrefilterParams P;
    // shards: the bias is a per-candidate constant, added by rank 0 only; the dense partial scores are
    // summed across the ranks before anyone ranks them
    int rc = prefilter_setup(e, chain, n_kids, kids, n_cands, cands, (e->comm.on && e->comm.rank != 0) ? nullptr : cand_bias, P);
    if (rc) return rc;
    CU(e->partial.reserve(std::max<size_t>((size_t)n_kids * n_cands, 1) * sizeof(double)));
    P.scores = static_cast<double *>(e->partial.p);
    e->prof_begin(3);
    e->launches += launch_prefilter_scores(P, e->stream);
    e->prof_end();
    CU(cudaGetLastError());
    if (e->comm.on) {
        const long long tot = (long long)n_kids * n_cands;
        for (long long o = 0; o < tot; o += 1 << 24) {          // int-sized pieces
            rc = comm_reduce(e, P.scores + o, (int)std::min<long long>(1 << 24, tot - o));
            if (rc) return rc;
        }
    }
    return prefilter_rank(e, P, cands, top_k, out_slot, out_score);
}

// The two halves of pf_prefilter for engines that hold a locus shard: dense partial scores first,
// the ranking after the caller has summed them over the shards (NCCL all-reduce on this stream).
extern "C" int pf_prefilter_scores_dev(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                                       const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                                       const double *cand_bias, double *scores_dev)
{
    if (!e || n_kids < 0 || n_cands < 0 || (n_kids > 0 && (!kids || !as_father)) || (n_cands > 0 && !cands)) return fail(PF_EINVAL, "null argument");
    if (n_kids == 0 || n_cands == 0) return PF_OK;
    if (!scores_dev) return fail(PF_EINVAL, "null output");
    PrefilterParams P;
    int rc = prefilter_setup(e, chain, n_kids, kids, n_cands, cands, cand_bias, P);
    if (rc) return rc;
    P.scores = scores_dev;
    e->prof_begin(3);
    e->launches += launch_prefilter_scores(P, e->stream);
    e->prof_end();
    CU(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_prefilter_topk_dev(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids, int32_t n_cands,
                                     const int32_t *cands, double *scores_dev, int32_t top_k, int32_t *out_slot, double *out_score)
{
    if (!e || n_kids < 0 || n_cands < 0 || (n_kids > 0 && (!kids || !out_slot || !out_score)) || (n_cands > 0 && (!cands || !scores_dev))) return fail(PF_EINVAL, "null argument");
    if (top_k <= 0 || top_k > PREF_MAX_TOPK) return fail(PF_EINVAL, "top_k must be in [1, %d]", PREF_MAX_TOPK);
    if (n_kids == 0) return PF_OK;
    PrefilterParams P;
    int rc = prefilter_setup(e, chain, n_kids, kids, n_cands, cands, nullptr, P);
    if (rc) return rc;
    P.scores = scores_dev;
    return prefilter_rank(e, P, cands, top_k, out_slot, out_score);
}

// ---------------------------------------------------------------------------------------------
// Read-back

static int get_row(pf_engine *e, int row, double *out)
{
    CU(cudaSetDevice(e->cfg.device));
    std::vector<double> buf(e->row_elems());
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(buf.data(), e->d_rows + (size_t)row * e->row_elems(), buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int s = 0; s < e->Sl; s++)
        for (int g = 0; g < 3; g++) out[(size_t)s * 3 + g] = buf[(size_t)g * e->Sp + s];
    return PF_OK;
}

extern "C" int pf_get_stub(pf_engine *e, int32_t chain, int32_t slot, double *out)
{
    if (!e || !out) return fail(PF_EINVAL, "null argument");
    if (chain < 0 || chain >= e->cfg.n_chains || slot < 0 || slot >= e->cfg.cap_indiv) return fail(PF_EINVAL, "index out of range");
    return get_row(e, e->stub_row(chain, slot), out);
}

extern "C" int pf_get_prong(pf_engine *e, int32_t chain, int32_t marr, double *out)
{
    if (!e || !out) return fail(PF_EINVAL, "null argument");
    if (chain < 0 || chain >= e->cfg.n_chains || marr < 0 || marr >= e->cfg.cap_marr) return fail(PF_EINVAL, "index out of range");
    return get_row(e, e->prong_row(chain, marr), out);
}
