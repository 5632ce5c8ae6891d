// capi.cu -- the C ABI of libdnascent_b200.so (include/dnascent_b200.h): contexts, model tables, batch staging
// (pinned host <-> HBM), the two-phase device pipeline and result hand-out.
//
// Host-side counterpart of the reference's per-read call
//     normaliseEvents(r, false)                      /root/reference/src/detect.cpp:876
// turned into a batch: the OpenMP read loop (detect.cpp:852) hands a buffer of reads to dnb_submit.
//
// Pipeline of one batch (all on the batch's own stream):
//   phase A   segmentation -> k-mer ranks -> quantile scaling -> event scaling
//   host      n_events comes back (one small D2H); the per-read transition constants need glibc's log/exp to be
//             bit-identical with the reference (event_handling.cpp:174-183), and the trace/alignment workspaces
//             are sized from n_events, so both are done here and sent down
//   phase B   banded DP -> backtrace/QC -> Theil-Sen
//   fetch     alignment compaction -> D2H of events, alignment pairs and per-read scalars
//
// Memory classes of a batch:
//   inputs     signal, sequences, offsets: allocated by upload, resident until release   (~2.6 B/sample as int16)
//   workspace  everything the kernels write: allocated by run, freed by drop_workspace    (~27 B/sample)
//   results    pinned host copies handed out by dnb_result: allocated by fetch, from the context's pinned pool
// Device memory comes from the context's own cache of whole cudaMalloc blocks (DevCache below), so a freed workspace is
// what the next batch's run gets back without touching the driver; dnb_config.workspace_bytes caps what stays cached.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <chrono>
#include <vector>
#include <omp.h>

#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

namespace {

thread_local std::string g_last_error;

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t _e = (call);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(_e);                                \
            cudaGetLastError();                                                                               \
            return DNB_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)

#define TRY(x)                         \
    do {                               \
        int _rc = (x);                 \
        if (_rc != DNB_OK) return _rc; \
    } while (0)

// Host-side phase accounting of the batch pipeline: every phase of upload / run / fetch adds its wall time (summed over
// the calling threads) to a process-wide table that dnb_host_stats() reads; DNB_TRACE_HOST=1 additionally prints every
// phase to stderr as it ends (diagnostics only).
enum HostPhase {
    PH_UP_GATE_PACK, PH_UP_SHAPES, PH_UP_PACK, PH_UP_GATE_H2D, PH_UP_ENQUEUE, PH_UP_WAIT,
    PH_RUN_GATE, PH_RUN_ALLOC_A, PH_RUN_SEG_LAUNCH, PH_RUN_WAIT_A, PH_RUN_HOST, PH_RUN_ALLOC_B, PH_RUN_ENQUEUE_B, PH_RUN_WAIT_B,
    PH_FETCH_GATE, PH_FETCH_SCALARS, PH_FETCH_WAIT, PH_N
};
const char *const kHostPhaseName[PH_N] = {
    "upload: wait for the pack slot", "upload: shapes, order, tile tables", "upload: pack into pinned staging",
    "upload: wait for the H2D slot", "upload: device alloc + enqueue copies", "upload: wait for H2D",
    "run: wait for a compute slot", "run: workspace A + pinned counters", "run: segmentation scratch + launches",
    "run: wait for phase A (GPU)", "run: transition constants, offsets (glibc)", "run: workspace B",
    "run: enqueue phase B", "run: wait for phase B (GPU)",
    "fetch: wait for the fetch slot", "fetch: per-read scalars (D2H + wait)", "fetch: compaction kernels + D2H + wait"};
struct HostStats {
    std::atomic<uint64_t> ns[PH_N];
    std::atomic<uint64_t> n_cuda_malloc{0}, n_cuda_malloc_host{0}, n_batches{0}, n_direct{0};
    HostStats() { for (auto &x : ns) x = 0; }
};
HostStats g_host_stats;

struct HostTrace {
    bool on;
    double prev;
    const char *who;
    explicit HostTrace(const char *w) : who(w) {
        static const bool enabled = getenv("DNB_TRACE_HOST") != nullptr;
        on = enabled;
        prev = omp_get_wtime();
    }
    void tick(int phase) {
        const double now = omp_get_wtime();
        g_host_stats.ns[phase].fetch_add((uint64_t)(1e9 * (now - prev)), std::memory_order_relaxed);
        if (on) fprintf(stderr, "[dnb host] %-6s %-44s %8.2f ms  (t=%.3f)\n", who, kHostPhaseName[phase], 1e3 * (now - prev), now);
        prev = now;
    }
};

struct ModelHost {
    bool loaded = false;
    double *d_mean = nullptr, *d_stdv = nullptr, *d_sorted = nullptr;
    uint32_t *d_order = nullptr;
    DnbModelDev dev() const { return DnbModelDev{d_mean, d_stdv, d_order, d_sorted}; }
};

// pinned host buffers are expensive to create (page locking): keep them for the life of the context
struct PinnedPool {
    struct Block { void *p; size_t n; bool used; };
    std::vector<Block> blocks;
    std::mutex mu;
    void *acquire(size_t n) {
        if (n == 0) n = 1;
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (size_t i = 0; i < blocks.size(); i++)
            if (!blocks[i].used && blocks[i].n >= n && (best < 0 || blocks[i].n < blocks[best].n)) best = (int)i;
        if (best >= 0 && blocks[best].n <= 4 * n + (1u << 20)) { blocks[best].used = true; return blocks[best].p; }
        const size_t gran = 1u << 20;
        const size_t sz = (n + gran - 1) / gran * gran;
        void *p = nullptr;
        static const bool trace = getenv("DNB_TRACE_HOST") != nullptr;
        const double t0 = trace ? omp_get_wtime() : 0.0;
        g_host_stats.n_cuda_malloc_host++;
        if (cudaMallocHost(&p, sz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (trace) fprintf(stderr, "[dnb host] cudaMallocHost %.1f MB took %.2f ms (t=%.3f)\n", sz / 1e6, 1e3 * (omp_get_wtime() - t0), omp_get_wtime());
        blocks.push_back({p, sz, true});
        return p;
    }
    void release(void *p) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &b : blocks) if (b.p == p) { b.used = false; return; }
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        size_t k = 0;
        for (auto &b : blocks) {
            if (b.used) blocks[k++] = b;
            else cudaFreeHost(b.p);
        }
        blocks.resize(k);
    }
    void destroy() {
        for (auto &b : blocks) cudaFreeHost(b.p);
        blocks.clear();
    }
};

// Device blocks are cached per context as whole cudaMalloc allocations, sizes rounded up to a coarse grid (1/16 of
// the next power of two, at least 2 MiB) so that bins of similar size map to identical block sizes and are reused
// without touching the driver.  (The stream-ordered pool fragmented under the 10-20 GB trace buffers: profiles/
// r1 host trace showed 0.1-1.1 s stalls in cudaMallocAsync.)  Blocks are only released after their stream has been
// synchronised, so reuse by another batch/stream is safe.
struct DevCache {
    struct Block { void *p; size_t n; bool used; };
    std::vector<Block> blocks;
    std::mutex mu;
    size_t idle_cap = 0;   // dnb_config.workspace_bytes: idle bytes kept cached (0 = no cap)
    static size_t grid(size_t n) {
        size_t g = (size_t)2 << 20;
        while (g * 16 < n) g <<= 1;
        return (n + g - 1) / g * g;
    }
    void *acquire(size_t n, cudaError_t *err) {
        n = grid(n ? n : 1);
        *err = cudaSuccess;
        {
            // best fit among the idle blocks of at most 1.25x the (grid-rounded) size: bins of similar size reuse each
            // other's blocks instead of going to the driver (cudaMalloc stalls every stream of the device)
            std::lock_guard<std::mutex> lk(mu);
            Block *best = nullptr;
            for (auto &b : blocks)
                if (!b.used && b.n >= n && b.n <= n + n / 4 && (!best || b.n < best->n)) best = &b;
            if (best) { best->used = true; return best->p; }
        }
        void *p = nullptr;
        static const bool trace = getenv("DNB_TRACE_HOST") != nullptr;
        const double t0 = trace ? omp_get_wtime() : 0.0;
        g_host_stats.n_cuda_malloc++;
        cudaError_t e = cudaMalloc(&p, n);
        if (trace) fprintf(stderr, "[dnb host] cudaMalloc %.1f MB took %.2f ms (t=%.3f)\n", n / 1e6, 1e3 * (omp_get_wtime() - t0), omp_get_wtime());
        // Out of memory: give the cached-but-idle blocks back and retry.  With several submissions in flight the shortage
        // is usually transient (another batch is about to release its workspace, and a single retry can lose the freed
        // memory to a concurrent caller), so keep trimming and retrying for up to ~20 s before reporting the error.
        for (int attempt = 0; e == cudaErrorMemoryAllocation && attempt < 400; attempt++) {
            cudaGetLastError();
            trim();
            e = cudaMalloc(&p, n);
            if (e == cudaErrorMemoryAllocation) std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
        if (e != cudaSuccess) { cudaGetLastError(); *err = e; return nullptr; }
        std::lock_guard<std::mutex> lk(mu);
        blocks.push_back({p, n, true});
        return p;
    }
    void release(void *p) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &b : blocks) if (b.p == p) { b.used = false; break; }
        if (!idle_cap) return;
        // another user of the same GPU (the reference's TensorFlow session) cannot ask us to trim: keep at most
        // idle_cap bytes of idle blocks, largest ones go first
        for (;;) {
            size_t idle = 0; int big = -1;
            for (size_t i = 0; i < blocks.size(); i++)
                if (!blocks[i].used) { idle += blocks[i].n; if (big < 0 || blocks[i].n > blocks[big].n) big = (int)i; }
            if (idle <= idle_cap || big < 0) return;
            cudaFree(blocks[big].p);
            blocks.erase(blocks.begin() + big);
        }
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        size_t k = 0;
        for (auto &b : blocks) {
            if (b.used) blocks[k++] = b;
            else cudaFree(b.p);
        }
        blocks.resize(k);
    }
    void destroy() {
        for (auto &b : blocks) cudaFree(b.p);
        blocks.clear();
    }
};

}  // namespace

// counting semaphore: how many batches may be in one pipeline stage of dnb_submit at a time
struct StageGate {
    std::mutex mu;
    std::condition_variable cv;
    int free_slots;
    explicit StageGate(int n) : free_slots(n) {}
    void enter() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return free_slots > 0; }); free_slots--; }
    void leave() { { std::lock_guard<std::mutex> lk(mu); free_slots++; } cv.notify_one(); }
};
struct StageHold {
    StageGate &g;
    explicit StageHold(StageGate &gate) : g(gate) { g.enter(); }
    ~StageHold() { g.leave(); }
};

// streams and events are recycled: creating them per batch cost ~30 ms per dnb_submit under load (driver lock)
struct StreamSet {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[10] = {};
    cudaEvent_t sync_ev = nullptr;   // cudaEventBlockingSync: a waiting host thread sleeps instead of spinning
};
struct StreamPool {
    std::vector<StreamSet> idle;
    std::mutex mu;
    bool acquire(StreamSet *out) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!idle.empty()) { *out = idle.back(); idle.pop_back(); return true; }
        }
        StreamSet s;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
        for (auto &e : s.ev)
            if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return false; }
        if (cudaEventCreateWithFlags(&s.sync_ev, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        *out = s;
        return true;
    }
    void release(const StreamSet &s) {
        std::lock_guard<std::mutex> lk(mu);
        idle.push_back(s);
    }
    void destroy() {
        for (auto &s : idle) {
            for (auto &e : s.ev) if (e) cudaEventDestroy(e);
            if (s.sync_ev) cudaEventDestroy(s.sync_ev);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        idle.clear();
    }
};

struct dnb_ctx {
    dnb_config cfg;
    ModelHost model[3];
    double emit_const = 0.0;
    std::mutex mu;
    PinnedPool pinned;
    DevCache dev;
    // dnb_submit from several host threads forms a software pipeline: one batch packing, up to two enqueueing /
    // copying in (so that the copy engine always has the next batch's copies queued behind the current one's), up to
    // three computing (one warp per read and a serial band chain per read: the tail of one submission's alignment launch
    // leaves most warp slots idle, which the next submissions' kernels fill; measured at 100k reads, 8e8-sample
    // submissions: 2 computing 7 750, 3 computing 8 280, 4 computing 8 220 Msamples/s end to end; DNB_COMPUTE_SLOTS
    // overrides), one copying out.  Without the gates concurrent callers fall into lockstep (all upload, then all
    // compute, then all fetch) and the GPU idles during the copy phases.
    StageGate gate_pack{1}, gate_h2d{2}, gate_compute{3}, gate_fetch{1};
    StreamPool streams;
    cudaMemPool_t pool = nullptr;        // stream-ordered scratch of the non-batch entry points (private: no global side effect)
    // n_devices > 1: this context is the front of a set; peers[k] drives cfg.devices[k + 1]
    std::vector<dnb_ctx *> peers;
    std::atomic<uint64_t> inflight{0};   // samples submitted and not yet released (the dealing rule's load measure)
};

namespace {

// device workspace: everything the kernels write (re-created by every run, dropped on request)
struct Work {
    uint32_t *et_n, *n_events, *ev_start;
    float *ev_mean;
    int *status;
    uint64_t *et_start;
    float *et_length, *et_mean, *et_stdv;
    double *mu_q, *x_e, *rough_shift, *rough_scale, *lp, *tot_sum;
    uint32_t *rank_ref, *redo;
    uint64_t *band_off, *al_off, *cl_off, *out_off;
    uint8_t *trace;
    uint32_t *moves, *rcum;
    int32_t *end_event, *end_ll;
    float *end_score;
    unsigned long long *cells;   // [3]: DP cells filled, warp cycles in the band fill, warp cycles in the backtrace
    uint32_t *al_rev, *n_align, *cl_rank, *n_cleaned, *out_pairs, *cev_start;
    float *cev_mean;
    uint32_t *n_long;            // events of >= 255 samples per read (escapes of the compact event coding)
    double *cl_signal, *avg, *shift, *scale;
    int *spanned, *max_gap;
    uint32_t *ts_nan_list;       // Theil-Sen reads with a 0/0 slope (theilsen.cu): entries, values, then the counter
    double *ts_nan_scratch;      // DNB_TS_NAN_SLOTS slots of slopes + position lists
};

// pinned host results
struct HostRes {
    uint32_t *et_n, *n_events, *n_align, *n_cleaned, *redo;
    int *status, *spanned, *max_gap;
    double *rough_shift, *rough_scale, *shift, *scale, *avg;
    uint32_t *ev_start, *out_pairs, *cl_rank;
    float *ev_mean;
    double *cl_signal;
    uint64_t *et_start;
    float *et_length, *et_mean, *et_stdv;
    // DNB_RESULT_COMPACT
    uint32_t *n_long, *esc, *first_ev, *first_pair, *bad_steps;
    uint8_t *len8, *steps;
};

}  // namespace

// results of the resident eventalign + tensor stage (dnb_batch_eventalign_features), pinned host
struct Stage2 {
    bool done = false, want_records = false;
    std::vector<uint64_t> rec_off, pos_off;
    dnb_eventalign_rec *h_recs = nullptr;
    uint32_t *h_n_rec = nullptr, *h_n_pos = nullptr;
    int *h_status = nullptr;
    float *h_signal = nullptr, *h_core = nullptr, *h_resid = nullptr;
    uint32_t *h_coords = nullptr, *h_ri = nullptr, *h_qi = nullptr;
    int32_t *h_qual = nullptr;
    double ms[2] = {0.0, 0.0};
    uint64_t bytes[2] = {0, 0};
};

// results of the resident analogue stage (dnb_batch_analogue_llr), pinned host
struct Stage3 {
    bool done = false;
    std::vector<uint64_t> poi_off;
    uint32_t *h_n_poi = nullptr, *h_pos = nullptr, *h_nev = nullptr;
    double *h_a = nullptr, *h_t = nullptr;
    double ms[2] = {0.0, 0.0};
    uint64_t counts[4] = {0, 0, 0, 0};
};

struct dnb_batch {
    dnb_ctx *ctx = nullptr;
    size_t R = 0;
    uint64_t load = 0;           // what this batch added to ctx->inflight (multi-device dealing)
    cudaStream_t stream = nullptr;
    bool want_table = false;     // dnb_detect_events: keep the full scrappie table, segmentation only
    bool uploaded = false, ran = false, fetched = false, have_work = false;
    // ---- host-side shapes ----
    std::vector<uint64_t> raw_off, q_off, r_off, ev_off, cev_off, band_off, al_off, cl_off, out_off, ck_off, esc_off, step_off;
    bool compact = false, have_dev_pairs = false;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;   // PCIe payload of the last upload / fetch
    std::vector<uint32_t> n_samples, order, qlen, rlen, tile_off, tile_read;
    std::vector<double> lp;
    bool i16 = false, direct_dma = false;
    uint64_t tot_raw = 0, tot_q = 0, tot_r = 0, tot_ev = 0, tot_bands = 0, tot_al = 0, tot_cl = 0, tot_out = 0;
    // ---- device: resident inputs ----
    void *d_raw = nullptr;
    float *d_dac_off = nullptr, *d_dac_scl = nullptr;
    char *d_query = nullptr, *d_ref = nullptr;
    int32_t *d_q2r = nullptr;
    uint64_t *d_raw_off = nullptr, *d_q_off = nullptr, *d_r_off = nullptr, *d_ev_off = nullptr, *d_ck_off = nullptr;
    uint32_t *d_n_samples = nullptr, *d_order = nullptr, *d_tile_off = nullptr, *d_tile_read = nullptr;
    Work w = {};
    HostRes h = {};
    // ---- timings ----
    cudaEvent_t ev[10] = {};     // [8], [9]: after the checkpoint / tile kernel of the segmentation
    cudaEvent_t sync_ev = nullptr;
    double ms[8] = {};
    double seg_ms[3] = {};       // checkpoint (or scan), tiles, stitch + events + serial redo
    uint64_t counts[8] = {};
    unsigned long long h_cells[3] = {};
    std::vector<void *> input_allocs, work_allocs, res_allocs;
    Stage2 s2;
    Stage3 s3;
};

namespace {

// Wait for everything enqueued on the batch's stream.  Several batches are in flight from several host threads (and
// under torchrun several ranks share the box's cores), so a waiting thread must sleep: cudaStreamSynchronize spins
// by default and N waiting threads would take N cores away from the threads that still have host work.
// DNB_SPIN_SYNC=1 restores the spinning wait (lowest latency for a single caller).
cudaError_t wait_stream(dnb_batch *b) {
    static const bool spin = getenv("DNB_SPIN_SYNC") != nullptr && getenv("DNB_SPIN_SYNC")[0] == '1';
    if (spin || !b->sync_ev) return cudaStreamSynchronize(b->stream);
    cudaError_t e = cudaEventRecord(b->sync_ev, b->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(b->sync_ev);
}

template <class T>
int dev_alloc(dnb_batch *b, std::vector<void *> &owner, T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e;
    void *q = b->ctx->dev.acquire(n * sizeof(T), &e);
    if (!q) {
        g_last_error = std::string("cudaMalloc(") + std::to_string(n * sizeof(T)) + " B): " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? DNB_ERR_NOMEM : DNB_ERR_CUDA;
    }
    owner.push_back(q);
    *p = (T *)q;
    return DNB_OK;
}
template <class T> int ialloc(dnb_batch *b, T **p, size_t n) { return dev_alloc(b, b->input_allocs, p, n); }
template <class T> int walloc(dnb_batch *b, T **p, size_t n) { return dev_alloc(b, b->work_allocs, p, n); }

// A group of device arrays carved out of ONE cached block: a batch then takes a handful of blocks of a few coarse
// size classes instead of ~60 exact-size ones, and concurrent batches of similar size find each other's blocks in the
// cache (a cache miss is a cudaMalloc, which stalls every stream of the device).
struct Arena {
    struct Item { void **p; size_t bytes; };
    std::vector<Item> items;
    template <class T> void add(T **p, size_t n) { *p = nullptr; items.push_back({(void **)p, (n ? n : 1) * sizeof(T)}); }
    int commit(dnb_batch *b, std::vector<void *> &owner) {
        size_t tot = 0;
        for (auto &it : items) tot += (it.bytes + 255) & ~(size_t)255;
        uint8_t *base = nullptr;
        TRY(dev_alloc(b, owner, &base, tot));
        size_t off = 0;
        for (auto &it : items) { *it.p = base + off; off += (it.bytes + 255) & ~(size_t)255; }
        items.clear();
        return DNB_OK;
    }
};

template <class T>
int ralloc(dnb_batch *b, T **p, size_t n) {   // pinned host, from the context pool
    void *q = b->ctx->pinned.acquire(n * sizeof(T));
    *p = (T *)q;
    if (!q) { g_last_error = "pinned host allocation failed"; return DNB_ERR_NOMEM; }
    b->res_allocs.push_back(q);
    return DNB_OK;
}

template <class T>
int h2d(dnb_batch *b, T *dst, const T *src, size_t n) {
    if (n == 0) return DNB_OK;
    CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, b->stream));
    return DNB_OK;
}
template <class T>
int d2h(dnb_batch *b, T *dst, const T *src, size_t n) {
    if (n == 0) return DNB_OK;
    CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, b->stream));
    return DNB_OK;
}

DnbBatchView make_view(const dnb_batch *b) {
    DnbBatchView v;
    memset(&v, 0, sizeof(v));
    v.n_reads = (uint32_t)b->R;
    v.order = b->d_order;
    v.raw_f32 = b->i16 ? nullptr : (const float *)b->d_raw;
    v.raw_i16 = b->i16 ? (const int16_t *)b->d_raw : nullptr;
    v.dac_offset = b->d_dac_off;
    v.dac_scale = b->d_dac_scl;
    v.raw_off = b->d_raw_off;
    v.n_samples = b->d_n_samples;
    v.query = b->d_query;
    v.ref = b->d_ref;
    v.q_off = b->d_q_off;
    v.r_off = b->d_r_off;
    v.q2r = b->d_q2r;
    v.et_n = b->w.et_n;
    v.n_events = b->w.n_events;
    v.ev_off = b->d_ev_off;
    v.ev_start = b->w.ev_start;
    v.ev_mean = b->w.ev_mean;
    v.status = b->w.status;
    v.et_start = b->w.et_start;
    v.et_length = b->w.et_length;
    v.et_mean = b->w.et_mean;
    v.et_stdv = b->w.et_stdv;
    return v;
}

// callers make sure the batch's stream is idle (every run/fetch ends with a synchronize)
void drop_work(dnb_batch *b) {
    for (void *p : b->work_allocs) b->ctx->dev.release(p);
    b->work_allocs.clear();
    b->w = Work{};
    b->have_work = false;
    b->have_dev_pairs = false;
}
void drop_results(dnb_batch *b) {
    for (void *p : b->res_allocs) b->ctx->pinned.release(p);
    b->res_allocs.clear();
    b->h = HostRes{};
    b->s2 = Stage2{};
    b->s3 = Stage3{};
    b->fetched = false;
}

void free_batch(dnb_batch *b) {
    if (!b) return;
    if (b->stream) wait_stream(b);
    drop_work(b);
    drop_results(b);
    for (void *p : b->input_allocs) b->ctx->dev.release(p);
    b->ctx->inflight.fetch_sub(b->load, std::memory_order_relaxed);
    if (b->stream) {
        StreamSet ss;
        ss.stream = b->stream;
        for (int i = 0; i < 10; i++) ss.ev[i] = b->ev[i];
        ss.sync_ev = b->sync_ev;
        b->ctx->streams.release(ss);
    }
    delete b;
}

// ---- caller memory known to be page-locked: ranges registered / allocated through this library ----------------
struct HostRegistry {
    std::mutex mu;
    std::vector<std::pair<uintptr_t, uintptr_t>> ranges;   // [lo, hi)
    void add(const void *p, size_t n) {
        std::lock_guard<std::mutex> lk(mu);
        ranges.emplace_back((uintptr_t)p, (uintptr_t)p + n);
    }
    bool remove(const void *p) {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < ranges.size(); i++)
            if (ranges[i].first == (uintptr_t)p) { ranges.erase(ranges.begin() + i); return true; }
        return false;
    }
    bool covers(const void *p, size_t n) {
        const uintptr_t lo = (uintptr_t)p, hi = lo + n;
        std::lock_guard<std::mutex> lk(mu);
        for (auto &r : ranges) if (lo >= r.first && hi <= r.second) return true;
        return false;
    }
};
HostRegistry g_host_reg;

// true when [p, p+n) can be the source of an asynchronous DMA: registered here, or page-locked by the caller's own
// cudaHostAlloc / cudaHostRegister (asked from the driver: first and last byte)
bool is_page_locked(const void *p, size_t n) {
    if (n == 0) return true;
    if (g_host_reg.covers(p, n)) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost) return false;
    if (cudaPointerGetAttributes(&a, (const char *)p + n - 1) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int upload(dnb_ctx *ctx, const dnb_read_desc *reads, size_t R, bool want_table, dnb_batch **out, bool gated = false) {
    if (!ctx || (!reads && R) || !out) return DNB_ERR_ARG;
    if (!ctx->model[DNB_MODEL_PORE].loaded && !want_table) return DNB_ERR_MODEL;
    if (ctx->cfg.use_fit_pore_model) {
        g_last_error = "use_fit_pore_model=1 is not implemented on the device path (all reference callers pass false)";
        return DNB_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->cfg.device));
    // dnb_submit pipelines concurrent callers: one batch packs on the host cores while another one's copy is on the bus
    struct Gate {
        StageGate *g = nullptr;
        void enter(StageGate *gate) { leave(); g = gate; if (g) g->enter(); }
        void leave() { if (g) g->leave(); g = nullptr; }
        ~Gate() { leave(); }
    } gate;
    HostTrace ht("upload");
    if (gated) gate.enter(&ctx->gate_pack);
    ht.tick(PH_UP_GATE_PACK);
    dnb_batch *b = new dnb_batch();
    b->ctx = ctx;
    b->R = R;
    b->want_table = want_table;
    cudaError_t e = cudaSuccess;
    {
        StreamSet ss;
        if (!ctx->streams.acquire(&ss)) { g_last_error = "stream/event creation failed"; delete b; return DNB_ERR_CUDA; }
        b->stream = ss.stream;
        for (int i = 0; i < 10; i++) b->ev[i] = ss.ev[i];
        b->sync_ev = ss.sync_ev;
    }

    // ---- shapes ----
    b->raw_off.resize(R + 1); b->q_off.resize(R + 1); b->r_off.resize(R + 1); b->ev_off.resize(R + 1);
    b->n_samples.resize(R); b->qlen.resize(R); b->rlen.resize(R); b->order.resize(R);
    bool any_i16 = false, any_f32 = false, any_dense = false, any_runs = false, same_seq = !want_table;
    uint64_t ro = 0, qo = 0, fo = 0, eo = 0, n_runs = 0;
    const double cap_per_sample = ctx->cfg.event_capacity_per_sample > 0 ? ctx->cfg.event_capacity_per_sample : 0.4;
    for (size_t i = 0; i < R; i++) {
        const dnb_read_desc &d = reads[i];
        if ((!d.raw_pA && !d.raw_dac) || d.n_samples == 0 || d.n_samples >= (1ull << 31) || !d.query || !d.ref ||
            (!want_table && !d.query_to_ref && !d.q2r_runs)) {
            free_batch(b);
            g_last_error = "read descriptor " + std::to_string(i) + " is incomplete";
            return DNB_ERR_ARG;
        }
        (d.raw_pA ? any_f32 : any_i16) = true;
        if (!want_table) { (d.query_to_ref ? any_dense : any_runs) = true; n_runs += d.query_to_ref ? 0 : d.n_q2r_runs; }
        same_seq = same_seq && d.ref == d.query && d.ref_len == d.query_len;
        b->raw_off[i] = ro; b->q_off[i] = qo; b->r_off[i] = fo; b->ev_off[i] = eo;
        b->n_samples[i] = (uint32_t)d.n_samples; b->qlen[i] = d.query_len; b->rlen[i] = d.ref_len;
        ro += (d.n_samples + 31) & ~31ull;                 // 128-byte aligned read starts (float4 loads)
        qo += d.query_len; fo += d.ref_len;
        eo += (uint64_t)(cap_per_sample * (double)d.n_samples) + 16;
    }
    if (any_i16 && any_f32) { free_batch(b); g_last_error = "mixing raw_pA and raw_dac reads in one batch"; return DNB_ERR_ARG; }
    if (any_dense && any_runs) { free_batch(b); g_last_error = "mixing query_to_ref and q2r_runs reads in one batch"; return DNB_ERR_ARG; }
    b->i16 = any_i16;
    b->raw_off[R] = ro; b->q_off[R] = qo; b->r_off[R] = fo; b->ev_off[R] = eo;
    b->tot_raw = ro; b->tot_q = qo; b->tot_r = fo; b->tot_ev = eo;
    std::iota(b->order.begin(), b->order.end(), 0u);
    std::stable_sort(b->order.begin(), b->order.end(),
                     [&](uint32_t x, uint32_t y) { return b->n_samples[x] > b->n_samples[y]; });   // longest first

    // ---- tile bookkeeping of the tiled segmentation ----
    b->tile_off.resize(R + 1); b->ck_off.resize(R + 1);
    {
        uint64_t to = 0, co = 0;
        for (size_t i = 0; i < R; i++) {
            b->tile_off[i] = (uint32_t)to; b->ck_off[i] = co;
            to += (b->n_samples[i] + DNB_SEG_TILE - 1) / DNB_SEG_TILE;
            co += (b->n_samples[i] + DNB_SEG_CK - 1) / DNB_SEG_CK + 1;
        }
        if (to >= (1ull << 32)) { free_batch(b); g_last_error = "batch too large (tile index overflows 32 bits)"; return DNB_ERR_ARG; }
        b->tile_off[R] = (uint32_t)to; b->ck_off[R] = co;
        b->tile_read.resize(to);
        for (size_t i = 0; i < R; i++)
            for (uint32_t g = b->tile_off[i]; g < b->tile_off[i + 1]; g++) b->tile_read[g] = (uint32_t)i;
    }

    // ---- zero-staging ingest: is every read's signal in page-locked caller memory? ----
    const size_t esz = b->i16 ? 2 : 4;
    static const bool no_direct = getenv("DNB_NO_DIRECT_DMA") != nullptr && getenv("DNB_NO_DIRECT_DMA")[0] == '1';
    bool direct = !no_direct && R > 0;
    for (size_t i = 0; direct && i < R; i++)
        direct = is_page_locked(b->i16 ? (const void *)reads[i].raw_dac : (const void *)reads[i].raw_pA, reads[i].n_samples * esz);
    b->direct_dma = direct;
    ht.tick(PH_UP_SHAPES);

    // ---- pinned staging + device inputs ----
    int rc = DNB_OK;
    if (same_seq) { fo = 0; b->r_off = b->q_off; b->tot_r = qo; }      // ref IS the query: one copy serves both
    uint8_t *h_raw = direct ? nullptr : (uint8_t *)ctx->pinned.acquire(ro * esz);
    char *h_q = (char *)ctx->pinned.acquire(qo), *h_r = same_seq ? nullptr : (char *)ctx->pinned.acquire(fo);
    int32_t *h_q2r = any_dense ? (int32_t *)ctx->pinned.acquire(qo * 4) : nullptr;
    // per-run tables of the compact queryToRef: the run, its read's query offset and length
    dnb_q2r_run *h_runs = any_runs ? (dnb_q2r_run *)ctx->pinned.acquire(n_runs * (sizeof(dnb_q2r_run) + 8 + 4) + 64) : nullptr;
    uint64_t *h_run_base = h_runs ? (uint64_t *)(h_runs + n_runs) : nullptr;
    uint32_t *h_run_len = h_runs ? (uint32_t *)(h_run_base + n_runs) : nullptr;
    // per-read tables travel in one pinned block (pageable sources would make every copy synchronous)
    const size_t nt = b->tile_read.size();
    const size_t meta_bytes = 8 * (R + 1) * 5 + 4 * R * 2 + 4 * (R + 1) + 4 * nt + 4 * R * 2 + 256;
    uint8_t *h_meta = (uint8_t *)ctx->pinned.acquire(meta_bytes);
    auto release_staging = [&]() {
        for (void *p : {(void *)h_raw, (void *)h_q, (void *)h_r, (void *)h_q2r, (void *)h_meta, (void *)h_runs})
            if (p) ctx->pinned.release(p);
    };
#define TRYF(x) do { rc = (x); if (rc != DNB_OK) { wait_stream(b); release_staging(); free_batch(b); return rc; } } while (0)
    if ((!direct && !h_raw) || !h_q || (!same_seq && !h_r) || (any_dense && !h_q2r) || (any_runs && !h_runs) || !h_meta) {
        g_last_error = "pinned staging allocation failed"; TRYF(DNB_ERR_NOMEM);
    }
    float *h_doff = (float *)h_meta, *h_dscl = h_doff + R;
    size_t meta_used = 8 * R;
    std::vector<uint64_t> run_off;
    if (any_runs) {
        run_off.resize(R + 1);
        uint64_t k = 0;
        for (size_t i = 0; i < R; i++) { run_off[i] = k; k += reads[i].n_q2r_runs; }
        run_off[R] = k;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < R; i++) {
        const dnb_read_desc &d = reads[i];
        if (!direct) {
            uint8_t *dst = h_raw + b->raw_off[i] * esz;
            const size_t padded = (size_t)(b->raw_off[i + 1] - b->raw_off[i]);
            if (b->i16) memcpy(dst, d.raw_dac, d.n_samples * 2); else memcpy(dst, d.raw_pA, d.n_samples * 4);
            memset(dst + d.n_samples * esz, 0, (padded - d.n_samples) * esz);
        }
        memcpy(h_q + b->q_off[i], d.query, d.query_len);
        if (!same_seq) memcpy(h_r + b->r_off[i], d.ref, d.ref_len);
        if (d.query_to_ref) memcpy(h_q2r + b->q_off[i], d.query_to_ref, (size_t)d.query_len * 4);
        else if (any_runs)
            for (uint32_t j = 0; j < d.n_q2r_runs; j++) {
                h_runs[run_off[i] + j] = d.q2r_runs[j];
                h_run_base[run_off[i] + j] = b->q_off[i];
                h_run_len[run_off[i] + j] = d.query_len;
            }
        h_doff[i] = d.dac_offset; h_dscl[i] = d.dac_scale;
    }
    ht.tick(PH_UP_PACK);
    if (gated) gate.enter(&ctx->gate_h2d);
    ht.tick(PH_UP_GATE_H2D);
    Arena ar;
    ar.add((uint8_t **)&b->d_raw, ro * esz);
    ar.add(&b->d_query, qo);
    if (!same_seq) ar.add(&b->d_ref, fo);
    if (!want_table) ar.add(&b->d_q2r, qo);
    ar.add(&b->d_dac_off, R); ar.add(&b->d_dac_scl, R);
    ar.add(&b->d_raw_off, R + 1); ar.add(&b->d_q_off, R + 1);
    if (!same_seq) ar.add(&b->d_r_off, R + 1);
    ar.add(&b->d_ev_off, R + 1); ar.add(&b->d_n_samples, R); ar.add(&b->d_order, R);
    ar.add(&b->d_tile_off, R + 1); ar.add(&b->d_tile_read, nt);
    ar.add(&b->d_ck_off, R + 1);
    dnb_q2r_run *d_runs = nullptr; uint64_t *d_run_base = nullptr; uint32_t *d_run_len = nullptr;
    if (any_runs) { ar.add(&d_runs, n_runs); ar.add(&d_run_base, n_runs); ar.add(&d_run_len, n_runs); }
    TRYF(ar.commit(b, b->input_allocs));
    if (same_seq) { b->d_ref = b->d_query; b->d_r_off = b->d_q_off; }
    uint64_t sent = 0;
    auto send = [&](void *dst, const void *src, size_t bytes) -> int {
        if (bytes == 0) return DNB_OK;
        sent += bytes;
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, b->stream));
        return DNB_OK;
    };
    auto send_table = [&](void *dst, const void *src, size_t bytes) -> int {   // through the pinned meta block
        if (bytes == 0) return DNB_OK;
        meta_used = (meta_used + 15) & ~(size_t)15;
        memcpy(h_meta + meta_used, src, bytes);
        const int rc2 = send(dst, h_meta + meta_used, bytes);
        meta_used += bytes;
        return rc2;
    };
    // small tables first: the kernels that expand / pad need them, and the long signal copy then has the bus to itself
    TRYF(send(b->d_dac_off, h_doff, R * 4)); TRYF(send(b->d_dac_scl, h_dscl, R * 4));
    TRYF(send_table(b->d_raw_off, b->raw_off.data(), 8 * (R + 1))); TRYF(send_table(b->d_q_off, b->q_off.data(), 8 * (R + 1)));
    if (!same_seq) TRYF(send_table(b->d_r_off, b->r_off.data(), 8 * (R + 1)));
    TRYF(send_table(b->d_ev_off, b->ev_off.data(), 8 * (R + 1)));
    TRYF(send_table(b->d_n_samples, b->n_samples.data(), 4 * R)); TRYF(send_table(b->d_order, b->order.data(), 4 * R));
    TRYF(send_table(b->d_tile_off, b->tile_off.data(), 4 * (R + 1)));
    TRYF(send_table(b->d_tile_read, b->tile_read.data(), 4 * nt));
    TRYF(send_table(b->d_ck_off, b->ck_off.data(), 8 * (R + 1)));
    TRYF(send(b->d_query, h_q, qo));
    if (!same_seq) TRYF(send(b->d_ref, h_r, fo));
    if (any_dense) TRYF(send(b->d_q2r, h_q2r, qo * 4));
    else if (!want_table) {
        // compact queryToRef: 16 B per run over PCIe, expanded to the dense array the backtrace reads in HBM
        if (qo) { cudaError_t em = cudaMemsetAsync(b->d_q2r, 0xFF, qo * 4, b->stream); if (em != cudaSuccess) { g_last_error = cudaGetErrorString(em); TRYF(DNB_ERR_CUDA); } }
        TRYF(send(d_runs, h_runs, n_runs * sizeof(dnb_q2r_run)));
        TRYF(send(d_run_base, h_run_base, n_runs * 8)); TRYF(send(d_run_len, h_run_len, n_runs * 4));
        dnb_launch_expand_q2r(d_runs, d_run_base, d_run_len, n_runs, b->d_q2r, b->stream);
    }
    if (direct) {
        // one DMA per read, from where the caller has it to its padded slot; nothing touches the samples on the host
        uint8_t *base = (uint8_t *)b->d_raw;
        // all of them in one driver call where the runtime has cudaMemcpyBatchAsync (CUDA >= 12.8): a per-read
        // cudaMemcpyAsync costs 5-20 us of host time under load, which at 10^5 reads per step is the H2D stage's time
        static const bool no_batch = getenv("DNB_NO_BATCH_MEMCPY") != nullptr && getenv("DNB_NO_BATCH_MEMCPY")[0] == '1';
        bool batched = false;
        if (!no_batch && R > 1) {
            std::vector<void *> dsts(R), srcs(R);
            std::vector<size_t> sizes(R);
            uint64_t tot = 0;
            for (size_t i = 0; i < R; i++) {
                dsts[i] = base + b->raw_off[i] * esz;
                srcs[i] = b->i16 ? (void *)reads[i].raw_dac : (void *)reads[i].raw_pA;
                sizes[i] = (size_t)reads[i].n_samples * esz;
                tot += sizes[i];
            }
            cudaMemcpyAttributes at;
            memset(&at, 0, sizeof(at));
            at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;     // the caller's buffers stay valid until we return
            size_t attr_idx = 0, fail_idx = 0;
            const cudaError_t eb = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), R, &at, &attr_idx, 1, &fail_idx, b->stream);
            if (eb == cudaSuccess) { batched = true; sent += tot; }
            else cudaGetLastError();                                // e.g. an older driver: fall back to one call per read
        }
        if (!batched)
            for (size_t i = 0; i < R; i++)
                TRYF(send(base + b->raw_off[i] * esz, b->i16 ? (const void *)reads[i].raw_dac : (const void *)reads[i].raw_pA,
                          (size_t)reads[i].n_samples * esz));
        dnb_launch_zero_padding(make_view(b), (uint32_t)esz, b->stream);
    } else {
        TRYF(send(b->d_raw, h_raw, ro * esz));
    }
    b->h2d_bytes = sent;
    ht.tick(PH_UP_ENQUEUE);
    e = wait_stream(b);
    if (e == cudaSuccess) e = cudaGetLastError();
    ht.tick(PH_UP_WAIT);
    g_host_stats.n_batches++;
    if (direct) g_host_stats.n_direct++;
    release_staging();
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); free_batch(b); return DNB_ERR_CUDA; }
    b->uploaded = true;
    *out = b;
    return DNB_OK;
#undef TRYF
}

// workspace that does not depend on the event counts
int alloc_work_a(dnb_batch *b) {
    const size_t R = b->R, eo = b->tot_ev, qo = b->tot_q, fo = b->tot_r;
    Work &w = b->w;
    Arena ar;
    ar.add(&w.et_n, R); ar.add(&w.n_events, R); ar.add(&w.status, R);
    ar.add(&w.ev_start, eo + R); ar.add(&w.ev_mean, eo);
    ar.add(&w.tot_sum, R); ar.add(&w.redo, R);
    if (b->want_table) {
        ar.add(&w.et_start, eo + R); ar.add(&w.et_length, eo + R);
        ar.add(&w.et_mean, eo + R); ar.add(&w.et_stdv, eo + R);
    } else {
        ar.add(&w.mu_q, qo); ar.add(&w.rank_ref, fo); ar.add(&w.x_e, eo);
        ar.add(&w.rough_shift, R); ar.add(&w.rough_scale, R); ar.add(&w.lp, 4 * R);
        ar.add(&w.band_off, R + 1); ar.add(&w.al_off, R + 1); ar.add(&w.cl_off, R + 1);
        ar.add(&w.out_off, R + 1);
        ar.add(&w.end_event, R); ar.add(&w.end_ll, R); ar.add(&w.end_score, R);
        ar.add(&w.cells, 3);
        ar.add(&w.n_align, R); ar.add(&w.n_cleaned, R); ar.add(&w.avg, R);
        ar.add(&w.shift, R); ar.add(&w.scale, R); ar.add(&w.spanned, R); ar.add(&w.max_gap, R);
        ar.add(&w.n_long, R);
    }
    TRY(ar.commit(b, b->work_allocs));
    if (!b->want_table) {
        // reads that stop early (UNDEFINED, OVERFLOW) never write their scalars and the blocks are recycled: hand out
        // zeros, not another batch's values
        for (void *p : {(void *)w.rough_shift, (void *)w.rough_scale, (void *)w.shift, (void *)w.scale, (void *)w.avg})
            CK(cudaMemsetAsync(p, 0, R * 8, b->stream));
        for (void *p : {(void *)w.n_align, (void *)w.n_cleaned, (void *)w.spanned, (void *)w.max_gap, (void *)w.n_long})
            CK(cudaMemsetAsync(p, 0, R * 4, b->stream));
    }
    b->have_work = true;
    return DNB_OK;
}

int run(dnb_batch *b) {
    if (!b || !b->uploaded) return DNB_ERR_STATE;
    dnb_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->cfg.device));
    const size_t R = b->R;
    cudaStream_t s = b->stream;
    const double wall0 = omp_get_wtime();
    HostTrace ht("run");
    auto tick = [&](int phase) { ht.tick(phase); };
    drop_work(b);
    drop_results(b);
    TRY(alloc_work_a(b));
    // the few per-read counters the host reads in the middle of the pipeline
    TRY(ralloc(b, &b->h.n_events, R)); TRY(ralloc(b, &b->h.et_n, R)); TRY(ralloc(b, &b->h.status, R));
    TRY(ralloc(b, &b->h.redo, R));
    tick(PH_RUN_ALLOC_A);
    DnbBatchView v = make_view(b);
    DnbDetector det = {ctx->cfg.window_length1, ctx->cfg.window_length2, ctx->cfg.threshold1, ctx->cfg.threshold2,
                       ctx->cfg.peak_height};
    uint64_t launches = 0;
    // scratch of the tiled segmentation: back in the cache at the mid-pipeline sync, before the DP workspace is taken
    std::vector<void *> seg_scratch;
    struct ScratchGuard {
        dnb_batch *b; std::vector<void *> &v;
        ~ScratchGuard() { if (!v.empty()) { wait_stream(b); for (void *p : v) b->ctx->dev.release(p); v.clear(); } }
    } scratch_guard{b, seg_scratch};
    CK(cudaEventRecord(b->ev[0], s));
    if (b->want_table) {
        dnb_launch_segmentation_serial(v, det, nullptr, s); launches++;
    } else {
        // per-run scratch of the tiled segmentation (back in the pool before the DP workspace is taken)
        const size_t nt = b->tile_off[R], nck = b->ck_off[R];
        uint8_t *scratch[10] = {};
        Arena sar;
        auto sal = [&](int i, size_t bytes) { sar.add(&scratch[i], bytes); };
        sal(0, nck * 8); sal(1, nck * 8); sal(2, nt * DNB_SEG_PEAK_CAP * 4); sal(3, nt * DNB_SEG_PEAK_CAP * 8);
        sal(4, nt * 4); sal(5, nt * DNB_SEG_BOUNDARY_BYTES); sal(6, nt * DNB_SEG_BOUNDARY_BYTES);
        sal(7, nt * 4); sal(8, nt * 4); sal(9, nt * 8);
        TRY(sar.commit(b, seg_scratch));
        DnbSegTiles t;
        memset(&t, 0, sizeof(t));
        t.n_tiles = (uint32_t)nt; t.tile_off = b->d_tile_off; t.tile_read = b->d_tile_read; t.ck_off = b->d_ck_off;
        t.tot_sum = b->w.tot_sum; t.redo = b->w.redo;
        t.ck_sum = (double *)scratch[0]; t.ck_sq = (double *)scratch[1]; t.pk_pos = (uint32_t *)scratch[2];
        t.pk_sum = (double *)scratch[3]; t.pk_count = (uint32_t *)scratch[4]; t.b_start = (SegBoundary *)scratch[5];
        t.b_end = (SegBoundary *)scratch[6]; t.tile_prefix = (uint32_t *)scratch[7]; t.tile_prev_pos = (uint32_t *)scratch[8];
        t.tile_prev_sum = (double *)scratch[9];
        dnb_launch_segmentation_tiled(v, det, t, s, b->ev[8], b->ev[9]); launches += 5;
        tick(PH_RUN_SEG_LAUNCH);
    }
    CK(cudaEventRecord(b->ev[1], s));
    b->compact = ctx->cfg.result_format == DNB_RESULT_COMPACT && !b->want_table;
    if (b->compact) { dnb_launch_count_long_events(v, b->w.n_long, s); launches++; }
    TRY(d2h(b, b->h.n_events, b->w.n_events, R));
    TRY(d2h(b, b->h.et_n, b->w.et_n, R));
    if (b->want_table) {
        TRY(d2h(b, b->h.status, b->w.status, R));
        CK(wait_stream(b));
        CK(cudaGetLastError());
        b->ran = true;
        return DNB_OK;
    }
    TRY(d2h(b, b->h.redo, b->w.redo, R));
    const DnbModelDev pore = ctx->model[DNB_MODEL_PORE].dev();
    dnb_launch_ranks(v, pore, b->w.mu_q, b->w.rank_ref, s); launches++;
    dnb_launch_quantile_scaling(v, pore, b->w.rank_ref, b->w.rough_shift, b->w.rough_scale, s); launches++;
    dnb_launch_scale_events(v, b->w.rough_shift, b->w.rough_scale, b->w.x_e, s); launches++;
    CK(cudaEventRecord(b->ev[2], s));
    CK(wait_stream(b));   // n_events is on the host (its copy was enqueued before the prep kernels)
    CK(cudaGetLastError());
    for (void *p : seg_scratch) ctx->dev.release(p);
    seg_scratch.clear();
    const double wall_gap0 = omp_get_wtime();
    tick(PH_RUN_WAIT_A);

    // ---- host step: transition constants with glibc (event_handling.cpp:174-183) + workspace shapes ----
    b->lp.resize(4 * R);
    b->band_off.assign(R + 1, 0); b->al_off.assign(R + 1, 0); b->cl_off.assign(R + 1, 0);
    uint64_t bo = 0, ao = 0, co = 0, n_ev = 0, n_km = 0, n_redo = 0;
    for (size_t i = 0; i < R; i++) {
        const uint32_t E = b->h.n_events[i];
        const int64_t K = (int64_t)b->qlen[i] - DNB_K + 1;
        b->band_off[i] = bo; b->al_off[i] = ao; b->cl_off[i] = co;
        n_redo += b->h.redo[i] ? 1 : 0;
        if (K >= 1 && E >= 1) {
            const double epk = (double)E / (double)K;
            const double p_stay = 1 - (1 / (epk + 1));
            const double lp_skip = log(1e-30), lp_stay = log(p_stay);
            const double lp_step = log(1.0 - exp(lp_skip) - exp(lp_stay)), lp_trim = log(0.01);
            b->lp[4 * i + 0] = lp_skip; b->lp[4 * i + 1] = lp_stay; b->lp[4 * i + 2] = lp_step; b->lp[4 * i + 3] = lp_trim;
            const uint64_t nb = (uint64_t)E + (uint64_t)K + 2;
            bo += nb; ao += nb; co += (uint64_t)K;
            n_ev += E; n_km += (uint64_t)K;
        }
    }
    b->band_off[R] = bo; b->al_off[R] = ao; b->cl_off[R] = co;
    b->tot_bands = bo; b->tot_al = ao; b->tot_cl = co;
    Work &w = b->w;
    tick(PH_RUN_HOST);
    {
        Arena ar;
        ar.add(&w.trace, bo * DNB_TRACE_ROW + 64);
        ar.add(&w.moves, (bo >> 5) + R + 2); ar.add(&w.rcum, (bo >> 5) + R + 2);
        ar.add(&w.al_rev, 2 * ao); ar.add(&w.cl_signal, co); ar.add(&w.cl_rank, co);
        ar.add(&w.ts_nan_list, 8 * (size_t)DNB_TS_NAN_CAP + 2);
        ar.add(&w.ts_nan_scratch, (size_t)DNB_TS_NAN_SLOTS * DNB_TS_NAN_SLOT_DOUBLES);
        TRY(ar.commit(b, b->work_allocs));
    }
    tick(PH_RUN_ALLOC_B);
    TRY(h2d(b, w.lp, b->lp.data(), 4 * R));
    TRY(h2d(b, w.band_off, b->band_off.data(), R + 1));
    TRY(h2d(b, w.al_off, b->al_off.data(), R + 1));
    TRY(h2d(b, w.cl_off, b->cl_off.data(), R + 1));
    CK(cudaMemsetAsync(w.cells, 0, 3 * sizeof(unsigned long long), s));

    DnbDpArgs dp;
    dp.x_e = w.x_e; dp.mu_q = w.mu_q; dp.lp = w.lp; dp.emit_const = ctx->emit_const; dp.inv_sigma = 1.0 / 0.14;
    dp.band_off = w.band_off; dp.trace = w.trace; dp.moves = w.moves; dp.rcum = w.rcum; dp.end_event = w.end_event; dp.end_ll_event = w.end_ll;
    dp.end_score = w.end_score; dp.cells = w.cells;
    DnbBtArgs bt;
    bt.dp = dp; bt.rank_ref = w.rank_ref; bt.al_off = w.al_off; bt.al_pairs_rev = w.al_rev; bt.n_align = w.n_align;
    bt.cl_off = w.cl_off; bt.cl_signal = w.cl_signal; bt.cl_rank = w.cl_rank; bt.n_cleaned = w.n_cleaned;
    bt.avg_log_emission = w.avg; bt.spanned = w.spanned; bt.max_gap = w.max_gap;
    bt.min_avg_log_emission = ctx->cfg.min_average_log_emission; bt.max_gap_threshold = ctx->cfg.max_gap_threshold;
    bt.phase_cycles = w.cells + 1;
    // DNB_SPLIT_ALIGN=1 launches the band fill and the backtrace as two kernels (same device code) so that a profiler
    // sees them separately; production is the single fused launch
    static const bool split_align = getenv("DNB_SPLIT_ALIGN") != nullptr;
    CK(cudaEventRecord(b->ev[3], s));
    const double wall_gap1 = omp_get_wtime();
    tick(PH_RUN_ENQUEUE_B);
    if (split_align) {
        dnb_launch_align(v, bt, 1, s); launches++;
        CK(cudaEventRecord(b->ev[4], s));
        dnb_launch_align(v, bt, 2, s); launches++;
    } else {
        dnb_launch_align(v, bt, 0, s); launches++;
        CK(cudaEventRecord(b->ev[4], s));
    }
    CK(cudaEventRecord(b->ev[5], s));
    DnbTsArgs ts;
    ts.cl_off = w.cl_off; ts.cl_signal = w.cl_signal; ts.cl_rank = w.cl_rank; ts.n_cleaned = w.n_cleaned;
    ts.rough_shift = w.rough_shift; ts.rough_scale = w.rough_scale; ts.shift = w.shift; ts.scale = w.scale;
    static const int ts_mode = getenv("DNB_TS_MODE") ? atoi(getenv("DNB_TS_MODE")) : 0;
    ts.mode = ts_mode;
    static const bool ts_nan_path = !(getenv("DNB_TS_NAN_PATH") && atoi(getenv("DNB_TS_NAN_PATH")) == 0);   // 0: "NaN last" only (cross-check)
    ts.nan_list = ts_nan_path ? w.ts_nan_list : nullptr;
    ts.nan_count = w.ts_nan_list + 8 * (size_t)DNB_TS_NAN_CAP;
    ts.nan_cap = DNB_TS_NAN_CAP; ts.nan_scratch = w.ts_nan_scratch; ts.nan_slots = DNB_TS_NAN_SLOTS;
    dnb_launch_theil_sen(v, pore, ts, s); launches += ts_nan_path ? 3 : 1;
    CK(cudaEventRecord(b->ev[6], s));
    CK(cudaMemcpyAsync(b->h_cells, w.cells, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    TRY(d2h(b, b->h.status, w.status, R));     // the outcome of every read is known to the host when run() returns
    CK(wait_stream(b));
    CK(cudaGetLastError());
    tick(PH_RUN_WAIT_B);
    float t;
    cudaEventElapsedTime(&t, b->ev[0], b->ev[1]); b->ms[0] = t;
    cudaEventElapsedTime(&t, b->ev[1], b->ev[2]); b->ms[1] = t;
    cudaEventElapsedTime(&t, b->ev[3], b->ev[4]); b->ms[2] = t;
    cudaEventElapsedTime(&t, b->ev[4], b->ev[5]); b->ms[3] = t;
    if (!split_align) {
        // one fused launch: attribute its duration to the two phases by the warp cycles each one took
        const double cyc = (double)b->h_cells[1] + (double)b->h_cells[2];
        const double share = cyc > 0 ? (double)b->h_cells[1] / cyc : 1.0;
        b->ms[3] = b->ms[2] * (1.0 - share);
        b->ms[2] = b->ms[2] * share;
    }
    cudaEventElapsedTime(&t, b->ev[5], b->ev[6]); b->ms[4] = t;
    cudaEventElapsedTime(&t, b->ev[0], b->ev[6]); b->ms[5] = t;
    cudaEventElapsedTime(&t, b->ev[0], b->ev[8]); b->seg_ms[0] = t;
    cudaEventElapsedTime(&t, b->ev[8], b->ev[9]); b->seg_ms[1] = t;
    cudaEventElapsedTime(&t, b->ev[9], b->ev[1]); b->seg_ms[2] = t;
    b->ms[6] = 1e3 * (wall_gap1 - wall_gap0);
    b->ms[7] = 1e3 * (omp_get_wtime() - wall0);
    uint64_t n_samp = 0, n_fail = 0;
    for (size_t i = 0; i < R; i++) { n_samp += b->n_samples[i]; n_fail += b->h.status[i] != DNB_READ_OK; }
    b->counts[0] = n_samp; b->counts[1] = n_ev; b->counts[2] = n_km; b->counts[3] = bo; b->counts[4] = b->h_cells[0];
    b->counts[5] = launches; b->counts[6] = n_redo; b->counts[7] = n_fail;
    b->ran = true;
    b->fetched = false;
    return DNB_OK;
}

int fetch(dnb_batch *b) {
    if (!b || !b->ran || !b->have_work) return DNB_ERR_STATE;
    if (b->fetched) return DNB_OK;
    dnb_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->cfg.device));
    const size_t R = b->R;
    cudaStream_t s = b->stream;
    HostRes &h = b->h;
    Work &w = b->w;
    HostTrace ht("fetch");
    if (b->want_table) {
        TRY(ralloc(b, &h.ev_start, b->tot_ev + R)); TRY(ralloc(b, &h.ev_mean, b->tot_ev));
        TRY(d2h(b, h.ev_start, w.ev_start, b->tot_ev + R));
        TRY(d2h(b, h.ev_mean, w.ev_mean, b->tot_ev));
        b->cev_off = b->ev_off;
        TRY(ralloc(b, &h.et_start, b->tot_ev + R)); TRY(ralloc(b, &h.et_length, b->tot_ev + R));
        TRY(ralloc(b, &h.et_mean, b->tot_ev + R)); TRY(ralloc(b, &h.et_stdv, b->tot_ev + R));
        TRY(d2h(b, h.et_start, w.et_start, b->tot_ev + R)); TRY(d2h(b, h.et_length, w.et_length, b->tot_ev + R));
        TRY(d2h(b, h.et_mean, w.et_mean, b->tot_ev + R)); TRY(d2h(b, h.et_stdv, w.et_stdv, b->tot_ev + R));
        CK(wait_stream(b));
        b->fetched = true;
        return DNB_OK;
    }
    // per-read scalars first: the host sizes the dense result arrays from them
    TRY(ralloc(b, &h.n_align, R)); TRY(ralloc(b, &h.n_cleaned, R)); TRY(ralloc(b, &h.spanned, R));
    TRY(ralloc(b, &h.max_gap, R)); TRY(ralloc(b, &h.rough_shift, R)); TRY(ralloc(b, &h.rough_scale, R));
    TRY(ralloc(b, &h.shift, R)); TRY(ralloc(b, &h.scale, R)); TRY(ralloc(b, &h.avg, R));
    TRY(d2h(b, h.status, w.status, R)); TRY(d2h(b, h.n_align, w.n_align, R));
    TRY(d2h(b, h.n_cleaned, w.n_cleaned, R)); TRY(d2h(b, h.spanned, w.spanned, R));
    TRY(d2h(b, h.max_gap, w.max_gap, R)); TRY(d2h(b, h.rough_shift, w.rough_shift, R));
    TRY(d2h(b, h.rough_scale, w.rough_scale, R)); TRY(d2h(b, h.shift, w.shift, R));
    TRY(d2h(b, h.scale, w.scale, R)); TRY(d2h(b, h.avg, w.avg, R));
    if (b->compact) { TRY(ralloc(b, &h.n_long, R)); TRY(d2h(b, h.n_long, w.n_long, R)); }
    CK(wait_stream(b));
    ht.tick(PH_FETCH_SCALARS);
    // dense layouts: events (the device slots are capacity-strided) and forward alignment pairs (the device list is
    // in backtrace order): only what dnb_result hands out crosses PCIe
    b->out_off.assign(R + 1, 0);
    b->cev_off.assign(R + 1, 0);
    uint64_t oo = 0, ce = 0, n_fail = 0;
    for (size_t i = 0; i < R; i++) {
        b->out_off[i] = oo; oo += h.n_align[i];
        b->cev_off[i] = ce; ce += h.status[i] == DNB_READ_OVERFLOW ? 0 : h.n_events[i];
        n_fail += h.status[i] != DNB_READ_OK;
    }
    b->out_off[R] = oo; b->cev_off[R] = ce;
    b->tot_out = oo;
    b->counts[7] = n_fail;
    uint64_t *d_cev_off = nullptr;
    Arena ar;
    ar.add(&d_cev_off, R + 1);
    if (!b->compact) {
        ar.add(&w.out_pairs, 2 * oo); ar.add(&w.cev_start, ce + R); ar.add(&w.cev_mean, ce);
        TRY(ar.commit(b, b->work_allocs));
        TRY(h2d(b, w.out_off, b->out_off.data(), R + 1));
        TRY(h2d(b, d_cev_off, b->cev_off.data(), R + 1));
        TRY(ralloc(b, &h.out_pairs, 2 * oo)); TRY(ralloc(b, &h.ev_start, ce + R)); TRY(ralloc(b, &h.ev_mean, ce));
        dnb_launch_compact_events(make_view(b), d_cev_off, w.cev_start, w.cev_mean, s);
        TRY(d2h(b, h.ev_start, w.cev_start, ce + R));
        TRY(d2h(b, h.ev_mean, w.cev_mean, ce));
        dnb_launch_compact_alignment(make_view(b), w.al_off, w.al_rev, w.n_align, w.out_off, w.out_pairs, s);
        b->have_dev_pairs = true;
        TRY(d2h(b, h.out_pairs, w.out_pairs, 2 * oo));
        b->d2h_bytes = 8 * ce + 4 * R + 8 * oo + 84 * R;
    } else {
        // compact wire format (pack.cu): 1 B length + 4 B mean per event, 2 bits per alignment step
        b->esc_off.assign(R + 1, 0); b->step_off.assign(R + 1, 0);
        uint64_t xo = 0, so = 0;
        for (size_t i = 0; i < R; i++) {
            b->esc_off[i] = xo; xo += h.status[i] == DNB_READ_OVERFLOW ? 0 : h.n_long[i];
            b->step_off[i] = so; so += h.n_align[i] > 1 ? (((uint64_t)h.n_align[i] - 1 + 3) / 4 + 3) & ~3ull : 0;
        }
        b->esc_off[R] = xo; b->step_off[R] = so;
        uint64_t *d_esc_off = nullptr, *d_step_off = nullptr;
        uint8_t *d_len8 = nullptr, *d_steps = nullptr;
        uint32_t *d_esc = nullptr, *d_first_ev = nullptr, *d_first_pair = nullptr, *d_bad = nullptr;
        ar.add(&d_esc_off, R + 1); ar.add(&d_step_off, R + 1);
        ar.add(&d_len8, ce); ar.add(&w.cev_mean, ce); ar.add(&d_esc, xo);
        ar.add(&d_first_ev, R); ar.add(&d_steps, so); ar.add(&d_first_pair, 2 * R);
        ar.add(&d_bad, R);
        TRY(ar.commit(b, b->work_allocs));
        TRY(h2d(b, w.out_off, b->out_off.data(), R + 1));
        TRY(h2d(b, d_cev_off, b->cev_off.data(), R + 1));
        TRY(ralloc(b, &h.len8, ce)); TRY(ralloc(b, &h.ev_mean, ce)); TRY(ralloc(b, &h.esc, xo));
        TRY(ralloc(b, &h.first_ev, R)); TRY(ralloc(b, &h.steps, so)); TRY(ralloc(b, &h.first_pair, 2 * R));
        TRY(ralloc(b, &h.bad_steps, R));
        TRY(h2d(b, d_esc_off, b->esc_off.data(), R + 1));
        TRY(h2d(b, d_step_off, b->step_off.data(), R + 1));
        CK(cudaMemsetAsync(d_bad, 0, R * 4, s));
        dnb_launch_compact_events8(make_view(b), d_cev_off, d_esc_off, d_len8, w.cev_mean, d_esc, d_first_ev, s);
        TRY(d2h(b, h.len8, d_len8, ce)); TRY(d2h(b, h.ev_mean, w.cev_mean, ce)); TRY(d2h(b, h.esc, d_esc, xo));
        TRY(d2h(b, h.first_ev, d_first_ev, R));
        dnb_launch_compact_steps(make_view(b), w.al_off, w.al_rev, w.n_align, d_step_off, d_steps, d_first_pair, d_bad, s);
        TRY(d2h(b, h.steps, d_steps, so)); TRY(d2h(b, h.first_pair, d_first_pair, 2 * R)); TRY(d2h(b, h.bad_steps, d_bad, R));
        b->d2h_bytes = 5 * ce + 4 * xo + so + 16 * R + 88 * R;
    }
    if (ctx->cfg.keep_debug) {
        TRY(ralloc(b, &h.cl_signal, b->tot_cl)); TRY(ralloc(b, &h.cl_rank, b->tot_cl));
        TRY(d2h(b, h.cl_signal, w.cl_signal, b->tot_cl));
        TRY(d2h(b, h.cl_rank, w.cl_rank, b->tot_cl));
    }
    CK(wait_stream(b));
    CK(cudaGetLastError());
    if (b->compact)
        for (size_t i = 0; i < R; i++)
            if (h.bad_steps[i]) { g_last_error = "alignment of read " + std::to_string(i) + " is not a monotone path"; return DNB_ERR_STATE; }
    ht.tick(PH_FETCH_WAIT);
    b->fetched = true;
    return DNB_OK;
}

template <class T>
cudaError_t pool_alloc(dnb_ctx *ctx, T **p, size_t bytes, cudaStream_t s) {
    return ctx->pool ? cudaMallocFromPoolAsync((void **)p, bytes, ctx->pool, s) : cudaMallocAsync((void **)p, bytes, s);
}

// one context driving several GPUs: a submission goes to the device with the fewest samples in flight
dnb_ctx *pick_device(dnb_ctx *front, const dnb_read_desc *reads, size_t R, uint64_t *load) {
    uint64_t n = 0;
    for (size_t i = 0; reads && i < R; i++) n += reads[i].n_samples;
    *load = n;
    dnb_ctx *best = front;
    for (dnb_ctx *p : front->peers)
        if (p->inflight.load(std::memory_order_relaxed) < best->inflight.load(std::memory_order_relaxed)) best = p;
    best->inflight.fetch_add(n, std::memory_order_relaxed);
    return best;
}

}  // namespace

// ================================================== exported C ABI ==================================================
extern "C" {

void dnb_default_config(dnb_config *c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->device = 0;
    c->window_length1 = 3; c->window_length2 = 6;                   // event_detection.h:19-25
    c->threshold1 = 1.4f; c->threshold2 = 9.0f; c->peak_height = 0.2f;
    c->min_average_log_emission = -2.0; c->max_gap_threshold = 5; c->bandwidth = DNB_BW;   // config.h:41
    c->use_fit_pore_model = 0;
    c->event_capacity_per_sample = 0.40f;
    c->keep_debug = 0;
    c->workspace_bytes = 0;
}

const char *dnb_strerror(int code) {
    switch (code) {
        case DNB_OK: return "ok";
        case DNB_ERR_ARG: return "invalid argument";
        case DNB_ERR_CUDA: return "CUDA error (see dnb_last_error)";
        case DNB_ERR_NOMEM: return "out of memory";
        case DNB_ERR_MODEL: return "pore model table not loaded";
        case DNB_ERR_STATE: return "call order violated";
        case DNB_ERR_NEGATIVE_LOG: return "Negative value passed to natural log function.";
        default: return "unknown error";
    }
}

const char *dnb_last_error(void) { return g_last_error.c_str(); }

static int create_one(dnb_ctx **out, const dnb_config &c) {
    CK(cudaSetDevice(c.device));
    dnb_ctx *ctx = new dnb_ctx();
    if (const char *e = getenv("DNB_COMPUTE_SLOTS")) {       // submissions computing at the same time (default 2, see gate_compute)
        const int n = atoi(e);
        if (n >= 1 && n <= 8) ctx->gate_compute.free_slots = n;
    }
    ctx->cfg = c;
    ctx->cfg.n_devices = 1;
    ctx->dev.idle_cap = c.workspace_bytes;
    // stream-ordered scratch for dnb_sequence_probability_batch / dnb_eventalign_batch: a pool of our own that keeps
    // what it has been given (the process-wide default pool is left as the application configured it)
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = c.device;
    if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else {
        cudaGetLastError();
        ctx->pool = nullptr;
    }
    // log_inv_sqrt_2pi is a float in the reference (event_handling.cpp:134); sigma is 0.14 for every k-mer
    const float log_inv_sqrt_2pi = (float)log(0.3989422804014327);
    ctx->emit_const = (double)log_inv_sqrt_2pi - log(0.14);
    *out = ctx;
    return DNB_OK;
}

int dnb_create(dnb_ctx **out, const dnb_config *cfg) {
    if (!out) return DNB_ERR_ARG;
    *out = nullptr;
    dnb_config c;
    if (cfg) c = *cfg; else dnb_default_config(&c);
    if (c.bandwidth != DNB_BW || c.window_length2 > 7 || c.window_length1 > c.window_length2 || c.window_length1 < 1) {
        g_last_error = "unsupported configuration (bandwidth must be 100, window lengths <= 7)";
        return DNB_ERR_ARG;
    }
    if (c.result_format != DNB_RESULT_DENSE && c.result_format != DNB_RESULT_COMPACT) return DNB_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_last_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libdnascent_b200 has no CPU fallback)";
        cudaGetLastError();
        return DNB_ERR_CUDA;
    }
    if (c.n_devices < 0 || c.n_devices > DNB_MAX_DEVICES) return DNB_ERR_ARG;
    if (c.n_devices > 1) {
        for (int k = 0; k < c.n_devices; k++) {
            if (c.devices[k] < 0 || c.devices[k] >= n) { g_last_error = "dnb_config.devices: no such CUDA device"; return DNB_ERR_ARG; }
            for (int j = 0; j < k; j++) if (c.devices[j] == c.devices[k]) { g_last_error = "dnb_config.devices: duplicate"; return DNB_ERR_ARG; }
        }
        c.device = c.devices[0];
    } else if (c.n_devices == 1) c.device = c.devices[0];
    if (c.device < 0 || c.device >= n) return DNB_ERR_ARG;
    dnb_ctx *front = nullptr;
    TRY(create_one(&front, c));
    for (int k = 1; k < c.n_devices; k++) {
        dnb_config ck = c;
        ck.device = c.devices[k];
        dnb_ctx *peer = nullptr;
        const int rc = create_one(&peer, ck);
        if (rc != DNB_OK) { dnb_destroy(front); return rc; }
        front->peers.push_back(peer);
    }
    front->cfg.n_devices = c.n_devices > 1 ? c.n_devices : 1;
    *out = front;
    return DNB_OK;
}

void dnb_destroy(dnb_ctx *ctx) {
    if (!ctx) return;
    for (dnb_ctx *p : ctx->peers) dnb_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->cfg.device);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    for (auto &m : ctx->model) {
        cudaFree(m.d_mean); cudaFree(m.d_stdv); cudaFree(m.d_sorted); cudaFree(m.d_order);
    }
    ctx->pinned.destroy();
    ctx->dev.destroy();
    ctx->streams.destroy();
    delete ctx;
}

int dnb_load_model(dnb_ctx *ctx, int which, const double *mean, const double *stdv, size_t n) {
    if (!ctx || which < 0 || which > 2 || !mean || n != DNB_N_KMERS) return DNB_ERR_ARG;
    if (which == DNB_MODEL_PORE && stdv) {
        // the alignment and eventalign kernels have the ONT table's static sigma compiled in as the reference has it
        // (import_poreModel_staticStdv, data_IO.cpp:170-173): any other value would silently change the emissions
        for (size_t i = 0; i < n; i++)
            if (stdv[i] != 0.14) { g_last_error = "DNB_MODEL_PORE: stdv must be the static 0.14 of data_IO.cpp:173 (or NULL)"; return DNB_ERR_ARG; }
    }
    for (dnb_ctx *p : ctx->peers) TRY(dnb_load_model(p, which, mean, stdv, n));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->cfg.device));
    ModelHost &m = ctx->model[which];
    if (!m.d_mean) {
        CK(cudaMalloc(&m.d_mean, n * 8)); CK(cudaMalloc(&m.d_stdv, n * 8));
        CK(cudaMalloc(&m.d_sorted, n * 8)); CK(cudaMalloc(&m.d_order, n * 4));
    }
    std::vector<uint32_t> idx(n), order(n);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return mean[a] < mean[b]; });
    std::vector<double> sorted(n), sd(n, 0.14);
    for (size_t i = 0; i < n; i++) { sorted[i] = mean[idx[i]]; order[idx[i]] = (uint32_t)i; }
    if (stdv) sd.assign(stdv, stdv + n);
    CK(cudaMemcpy(m.d_mean, mean, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.d_stdv, sd.data(), n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.d_sorted, sorted.data(), n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.d_order, order.data(), n * 4, cudaMemcpyHostToDevice));
    m.loaded = true;
    return DNB_OK;
}

int dnb_batch_upload(dnb_ctx *ctx, const dnb_read_desc *reads, size_t n_reads, dnb_batch **batch) {
    if (!ctx) return DNB_ERR_ARG;
    uint64_t load = 0;
    dnb_ctx *dev = pick_device(ctx, reads, n_reads, &load);
    const int rc = upload(dev, reads, n_reads, false, batch);
    if (rc == DNB_OK) (*batch)->load = load; else dev->inflight.fetch_sub(load, std::memory_order_relaxed);
    return rc;
}
int dnb_batch_run(dnb_batch *batch) { return run(batch); }
int dnb_batch_fetch(dnb_batch *batch) { return fetch(batch); }
int dnb_batch_drop_workspace(dnb_batch *b) {
    if (!b) return DNB_ERR_ARG;
    CK(cudaSetDevice(b->ctx->cfg.device));
    drop_work(b);
    return DNB_OK;
}

int dnb_submit(dnb_ctx *ctx, const dnb_read_desc *reads, size_t n_reads, dnb_batch **batch) {
    dnb_batch *b = nullptr;
    if (!ctx || !batch) return DNB_ERR_ARG;
    uint64_t load = 0;
    ctx = pick_device(ctx, reads, n_reads, &load);
    int rc = upload(ctx, reads, n_reads, false, &b, /*gated=*/true);
    if (rc != DNB_OK) { ctx->inflight.fetch_sub(load, std::memory_order_relaxed); return rc; }
    b->load = load;
    {
        HostTrace hg("submit");
        StageHold hold(ctx->gate_compute);
        hg.tick(PH_RUN_GATE);
        rc = run(b);
    }
    if (rc == DNB_OK) {
        HostTrace hg("submit");
        StageHold hold(ctx->gate_fetch);
        hg.tick(PH_FETCH_GATE);
        rc = fetch(b);
    }
    if (rc != DNB_OK) { free_batch(b); return rc; }
    drop_work(b);   // results are on the host: give the HBM workspace back to the pool for the next batch
    *batch = b;
    return DNB_OK;
}

int dnb_wait(dnb_batch *b) {
    if (!b) return DNB_ERR_ARG;
    CK(cudaSetDevice(b->ctx->cfg.device));
    CK(wait_stream(b));
    return DNB_OK;
}

int dnb_result(dnb_batch *b, size_t i, dnb_read_result *o) {
    if (!b || !o || i >= b->R) return DNB_ERR_ARG;
    if (!b->fetched || b->want_table) return DNB_ERR_STATE;
    const HostRes &h = b->h;
    memset(o, 0, sizeof(*o));
    o->status = h.status[i];
    o->et_n = h.et_n[i];
    o->n_events = h.n_events[i];
    if (o->status == DNB_READ_OVERFLOW) o->n_events = 0;
    o->event_mean = h.ev_mean + b->cev_off[i];
    o->n_align = o->status == DNB_READ_OK ? h.n_align[i] : 0;
    if (!b->compact) {
        o->event_start = h.ev_start + b->cev_off[i] + i;
        o->align_pairs = h.out_pairs + 2 * b->out_off[i];
    } else {
        o->event_first = h.first_ev[i];
        o->event_len8 = h.len8 + b->cev_off[i];
        o->event_len_escape = h.esc + b->esc_off[i];
        o->n_event_len_escape = (uint32_t)(b->esc_off[i + 1] - b->esc_off[i]);
        o->align_first[0] = h.first_pair[2 * i]; o->align_first[1] = h.first_pair[2 * i + 1];
        o->align_steps = h.steps + b->step_off[i];
    }
    o->rough_shift = h.rough_shift[i]; o->rough_scale = h.rough_scale[i];
    o->shift = h.shift[i]; o->scale = h.scale[i];
    const int64_t denom = (int64_t)b->qlen[i] - DNB_K;
    o->events_per_base = (double)o->et_n / (double)denom;                      // event_handling.cpp:606 (quirk Q4)
    o->avg_log_emission = h.avg[i]; o->spanned = h.spanned[i]; o->max_gap = h.max_gap[i];
    if (b->ctx->cfg.keep_debug && h.cl_signal) {
        o->n_cleaned = h.n_cleaned[i];
        o->cleaned_signal = h.cl_signal + b->cl_off[i];
        o->cleaned_rank = h.cl_rank + b->cl_off[i];
    }
    return DNB_OK;
}

void dnb_release(dnb_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->cfg.device);
    free_batch(b);
}

int dnb_expand_events(const dnb_read_result *r, uint32_t *event_start) {
    if (!r || (!event_start && r->n_events)) return DNB_ERR_ARG;
    if (r->n_events == 0) return DNB_OK;
    if (r->event_start) { memcpy(event_start, r->event_start, 4 * ((size_t)r->n_events + 1)); return DNB_OK; }
    if (!r->event_len8) return DNB_ERR_STATE;
    uint32_t at = r->event_first, k = 0;
    event_start[0] = at;
    for (uint32_t j = 0; j < r->n_events; j++) {
        uint32_t d = r->event_len8[j];
        if (d == 255u) {
            if (k >= r->n_event_len_escape) return DNB_ERR_STATE;
            d = r->event_len_escape[k++];
        }
        at += d;
        event_start[j + 1] = at;
    }
    return DNB_OK;
}

int dnb_expand_alignment(const dnb_read_result *r, uint32_t *align_pairs) {
    if (!r || (!align_pairs && r->n_align)) return DNB_ERR_ARG;
    if (r->n_align == 0) return DNB_OK;
    if (r->align_pairs) { memcpy(align_pairs, r->align_pairs, 8 * (size_t)r->n_align); return DNB_OK; }
    if (r->n_align > 1 && !r->align_steps) return DNB_ERR_STATE;
    uint32_t e = r->align_first[0], k = r->align_first[1];
    align_pairs[0] = e; align_pairs[1] = k;
    for (uint32_t t = 0; t + 1 < r->n_align; t++) {
        const uint32_t code = (r->align_steps[t >> 2] >> (2 * (t & 3))) & 3u;
        e += code != DNB_FROM_L;          // D and U advance the event, D and L the k-mer
        k += code != DNB_FROM_U;
        align_pairs[2 * t + 2] = e; align_pairs[2 * t + 3] = k;
    }
    return DNB_OK;
}

int dnb_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return DNB_ERR_ARG;
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    g_host_reg.add(p, bytes);
    return DNB_OK;
}
int dnb_host_unregister(void *p) {
    if (!p) return DNB_ERR_ARG;
    g_host_reg.remove(p);
    CK(cudaHostUnregister(p));
    return DNB_OK;
}
int dnb_host_alloc(void **p, size_t bytes) {
    if (!p || !bytes) return DNB_ERR_ARG;
    *p = nullptr;
    cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); cudaGetLastError(); return e == cudaErrorMemoryAllocation ? DNB_ERR_NOMEM : DNB_ERR_CUDA; }
    g_host_reg.add(*p, bytes);
    return DNB_OK;
}
void dnb_host_free(void *p) {
    if (!p) return;
    g_host_reg.remove(p);
    cudaFreeHost(p);
}

int dnb_trim(dnb_ctx *ctx) {
    if (!ctx) return DNB_ERR_ARG;
    for (dnb_ctx *p : ctx->peers) TRY(dnb_trim(p));
    CK(cudaSetDevice(ctx->cfg.device));
    ctx->dev.trim();
    ctx->pinned.trim();
    return DNB_OK;
}

int dnb_host_stats(int reset, double seconds[DNB_N_HOST_PHASES], uint64_t counts[4]) {
    static_assert(PH_N == DNB_N_HOST_PHASES, "dnb_host_stats phase count");
    for (int i = 0; i < PH_N; i++) {
        if (seconds) seconds[i] = 1e-9 * (double)g_host_stats.ns[i].load();
        if (reset) g_host_stats.ns[i] = 0;
    }
    if (counts) {
        counts[0] = g_host_stats.n_batches; counts[1] = g_host_stats.n_direct;
        counts[2] = g_host_stats.n_cuda_malloc; counts[3] = g_host_stats.n_cuda_malloc_host;
    }
    if (reset) { g_host_stats.n_batches = 0; g_host_stats.n_direct = 0; g_host_stats.n_cuda_malloc = 0; g_host_stats.n_cuda_malloc_host = 0; }
    return DNB_OK;
}
const char *dnb_host_phase_name(int i) { return (i >= 0 && i < PH_N) ? kHostPhaseName[i] : ""; }

int dnb_batch_device(dnb_batch *b) { return b ? b->ctx->cfg.device : -1; }

int dnb_batch_seg_timings(dnb_batch *b, double ms[3]) {
    if (!b || !b->ran || !ms) return DNB_ERR_STATE;
    for (int i = 0; i < 3; i++) ms[i] = b->seg_ms[i];
    return DNB_OK;
}

int dnb_batch_timings(dnb_batch *b, double ms[8], uint64_t counts[8]) {
    if (!b || !b->ran) return DNB_ERR_STATE;
    for (int i = 0; i < 8; i++) if (ms) ms[i] = b->ms[i];
    for (int i = 0; i < 8; i++) if (counts) counts[i] = b->counts[i];
    return DNB_OK;
}

int dnb_batch_io_bytes(dnb_batch *b, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
    if (!b) return DNB_ERR_ARG;
    if (h2d_bytes) *h2d_bytes = b->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = b->d2h_bytes;
    return DNB_OK;
}

int dnb_detect_events(dnb_ctx *ctx, const float *raw_pA, size_t n, dnb_event_t *events, size_t cap, size_t *n_events) {
    if (!ctx || !raw_pA || n == 0 || !n_events) return DNB_ERR_ARG;
    dnb_read_desc d;
    memset(&d, 0, sizeof(d));
    d.raw_pA = raw_pA; d.n_samples = n; d.query = ""; d.ref = "";
    dnb_batch *b = nullptr;
    TRY(upload(ctx, &d, 1, true, &b));
    int rc = run(b);
    if (rc == DNB_OK) rc = fetch(b);
    if (rc == DNB_OK) {
        const size_t ne = b->h.et_n[0];
        *n_events = ne;
        if (b->h.status[0] == DNB_READ_OVERFLOW) rc = DNB_ERR_NOMEM;
        for (size_t i = 0; rc == DNB_OK && i < ne && i < cap; i++) {
            events[i].start = b->h.et_start[i]; events[i].length = b->h.et_length[i];
            events[i].mean = b->h.et_mean[i]; events[i].stdv = b->h.et_stdv[i];
            events[i].pos = -1; events[i].state = -1;                         // event_detection.c:219-220
        }
    }
    free_batch(b);
    return rc;
}

// ---- probability.cpp drop-ins: scalar host functions, semantics of src/probability.cpp:23-154 ----
double dnb_eexp(double x) { return std::isnan(x) ? 0.0 : exp(x); }
int dnb_eln(double x, double *out) {
    if (!out) return DNB_ERR_ARG;
    if (x == 0.0) { *out = NAN; return DNB_OK; }
    if (x > 0.0) { *out = log(x); return DNB_OK; }
    return DNB_ERR_NEGATIVE_LOG;
}
static double eln_nothrow(double x) { double o = NAN; dnb_eln(x, &o); return o; }
double dnb_lnSum(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) {
        if (std::isnan(a) && std::isnan(b)) return NAN;
        return std::isnan(a) ? b : a;
    }
    if (a > b) return a + eln_nothrow(1.0 + dnb_eexp(b - a));
    return b + eln_nothrow(1.0 + dnb_eexp(a - b));
}
double dnb_lnProd(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : a + b; }
int dnb_lnGreaterThan(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) {
        if (std::isnan(a) || !std::isnan(b)) return 0;
        return 1;
    }
    return a > b ? 1 : 0;
}
double dnb_uniformPDF(double lb, double ub, double x) { return (x >= lb && x <= ub) ? 1.0 / (ub - lb) : 0.0; }
double dnb_normalPDF(double mu, double sigma, double x) {
    return (1.0 / sqrt(2.0 * pow(sigma, 2.0) * M_PI)) * exp(-pow(x - mu, 2.0) / (2.0 * pow(sigma, 2.0)));
}
double dnb_cauchyPDF(double loc, double scale, double x) { return 1. / ((scale * M_PI) * (1. + pow((x - loc) / scale, 2.))); }

int dnb_sequence_probability_batch(dnb_ctx *ctx, const double *obs, const uint64_t *obs_off, const char *seq,
                                   const double *shift, const double *scale, const double *epb, size_t n_sites,
                                   uint32_t window, double *out_analogue, double *out_thymidine) {
    if (!ctx || !obs_off || !seq || !shift || !scale || !epb || !out_analogue || !out_thymidine) return DNB_ERR_ARG;
    if (window < 5 || window > 16) return DNB_ERR_ARG;
    if (!ctx->model[DNB_MODEL_UNLABELLED].loaded || !ctx->model[DNB_MODEL_ANALOGUE].loaded) return DNB_ERR_MODEL;
    if (n_sites == 0) return DNB_OK;
    CK(cudaSetDevice(ctx->cfg.device));
    const size_t n_obs = obs_off[n_sites], snip = 2 * (size_t)window + DNB_K;
    double *d_obs = nullptr, *d_shift = nullptr, *d_scale = nullptr, *d_epb = nullptr, *d_oa = nullptr, *d_ot = nullptr;
    uint64_t *d_off = nullptr;
    char *d_seq = nullptr;
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = DNB_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == DNB_OK) { g_last_error = cudaGetErrorString(e); rc = DNB_ERR_CUDA; } };
    fail(pool_alloc(ctx, &d_obs, (n_obs ? n_obs : 1) * 8, s)); fail(pool_alloc(ctx, &d_off, (n_sites + 1) * 8, s));
    fail(pool_alloc(ctx, &d_seq, n_sites * snip, s)); fail(pool_alloc(ctx, &d_shift, n_sites * 8, s));
    fail(pool_alloc(ctx, &d_scale, n_sites * 8, s)); fail(pool_alloc(ctx, &d_epb, n_sites * 8, s));
    fail(pool_alloc(ctx, &d_oa, n_sites * 8, s)); fail(pool_alloc(ctx, &d_ot, n_sites * 8, s));
    if (rc == DNB_OK) {
        if (n_obs) fail(cudaMemcpyAsync(d_obs, obs, n_obs * 8, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_off, obs_off, (n_sites + 1) * 8, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_seq, seq, n_sites * snip, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_shift, shift, n_sites * 8, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_scale, scale, n_sites * 8, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_epb, epb, n_sites * 8, cudaMemcpyHostToDevice, s));
        dnb_launch_hmm_forward(d_obs, d_off, d_seq, d_shift, d_scale, d_epb, n_sites, window,
                               ctx->model[DNB_MODEL_UNLABELLED].dev(), ctx->model[DNB_MODEL_ANALOGUE].dev(), d_oa, d_ot, s);
        fail(cudaMemcpyAsync(out_analogue, d_oa, n_sites * 8, cudaMemcpyDeviceToHost, s));
        fail(cudaMemcpyAsync(out_thymidine, d_ot, n_sites * 8, cudaMemcpyDeviceToHost, s));
        fail(cudaStreamSynchronize(s));
        fail(cudaGetLastError());
    }
    for (void *p : {(void *)d_obs, (void *)d_off, (void *)d_seq, (void *)d_shift, (void *)d_scale, (void *)d_epb, (void *)d_oa, (void *)d_ot})
        if (p) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

// ---- estimateScaling_theilSen alone (event_handling.cpp:24-110): the Theil-Sen kernels of run() on caller-supplied vectors ----
int dnb_theil_sen_batch(dnb_ctx *ctx, const double *signals, const uint32_t *ranks, const uint64_t *off, size_t n_reads,
                        const double *rough_shift, const double *rough_scale, double *shift, double *scale) {
    if (!ctx || !off || !rough_shift || !rough_scale || !shift || !scale) return DNB_ERR_ARG;
    if (!ctx->model[DNB_MODEL_PORE].loaded) return DNB_ERR_MODEL;
    if (n_reads == 0) return DNB_OK;
    if (n_reads >= (1ull << 32)) return DNB_ERR_ARG;
    const size_t n_pts = off[n_reads];
    if (n_pts && (!signals || !ranks)) return DNB_ERR_ARG;
    for (size_t i = 0; i < n_pts; i++)
        if (ranks[i] >= DNB_N_KMERS) { g_last_error = "dnb_theil_sen_batch: k-mer rank out of range"; return DNB_ERR_ARG; }
    CK(cudaSetDevice(ctx->cfg.device));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = DNB_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == DNB_OK) { g_last_error = cudaGetErrorString(e); rc = DNB_ERR_CUDA; } };
    double *d_sig = nullptr, *d_rs = nullptr, *d_rc = nullptr, *d_shift = nullptr, *d_scale = nullptr, *d_scratch = nullptr;
    uint32_t *d_rank = nullptr, *d_n = nullptr, *d_order = nullptr, *d_nan = nullptr;
    uint64_t *d_off = nullptr;
    int *d_status = nullptr;
    std::vector<uint32_t> n_cl(n_reads), order(n_reads);
    for (size_t i = 0; i < n_reads; i++) { n_cl[i] = (uint32_t)(off[i + 1] - off[i]); order[i] = (uint32_t)i; }
    fail(pool_alloc(ctx, &d_sig, (n_pts ? n_pts : 1) * 8, s)); fail(pool_alloc(ctx, &d_rank, (n_pts ? n_pts : 1) * 4, s));
    fail(pool_alloc(ctx, &d_off, (n_reads + 1) * 8, s)); fail(pool_alloc(ctx, &d_n, n_reads * 4, s));
    fail(pool_alloc(ctx, &d_order, n_reads * 4, s)); fail(pool_alloc(ctx, &d_status, n_reads * 4, s));
    fail(pool_alloc(ctx, &d_rs, n_reads * 8, s)); fail(pool_alloc(ctx, &d_rc, n_reads * 8, s));
    fail(pool_alloc(ctx, &d_shift, n_reads * 8, s)); fail(pool_alloc(ctx, &d_scale, n_reads * 8, s));
    fail(pool_alloc(ctx, &d_nan, (8 * (size_t)DNB_TS_NAN_CAP + 2) * 4, s));
    fail(pool_alloc(ctx, &d_scratch, (size_t)DNB_TS_NAN_SLOTS * DNB_TS_NAN_SLOT_DOUBLES * 8, s));
    if (rc == DNB_OK) {
        if (n_pts) {
            fail(cudaMemcpyAsync(d_sig, signals, n_pts * 8, cudaMemcpyHostToDevice, s));
            fail(cudaMemcpyAsync(d_rank, ranks, n_pts * 4, cudaMemcpyHostToDevice, s));
        }
        fail(cudaMemcpyAsync(d_off, off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_n, n_cl.data(), n_reads * 4, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_order, order.data(), n_reads * 4, cudaMemcpyHostToDevice, s));
        fail(cudaMemsetAsync(d_status, 0, n_reads * 4, s));
        fail(cudaMemcpyAsync(d_rs, rough_shift, n_reads * 8, cudaMemcpyHostToDevice, s));
        fail(cudaMemcpyAsync(d_rc, rough_scale, n_reads * 8, cudaMemcpyHostToDevice, s));
        DnbBatchView v{};
        v.n_reads = (uint32_t)n_reads; v.order = d_order; v.status = d_status;
        DnbTsArgs ts{};
        ts.cl_off = d_off; ts.cl_signal = d_sig; ts.cl_rank = d_rank; ts.n_cleaned = d_n;
        ts.rough_shift = d_rs; ts.rough_scale = d_rc; ts.shift = d_shift; ts.scale = d_scale;
        ts.mode = getenv("DNB_TS_MODE") ? atoi(getenv("DNB_TS_MODE")) : 0;
        const bool nan_path = !(getenv("DNB_TS_NAN_PATH") && atoi(getenv("DNB_TS_NAN_PATH")) == 0);
        ts.nan_list = nan_path ? d_nan : nullptr;
        ts.nan_count = d_nan + 8 * (size_t)DNB_TS_NAN_CAP;
        ts.nan_cap = DNB_TS_NAN_CAP; ts.nan_scratch = d_scratch; ts.nan_slots = DNB_TS_NAN_SLOTS;
        dnb_launch_theil_sen(v, ctx->model[DNB_MODEL_PORE].dev(), ts, s);
        fail(cudaMemcpyAsync(shift, d_shift, n_reads * 8, cudaMemcpyDeviceToHost, s));
        fail(cudaMemcpyAsync(scale, d_scale, n_reads * 8, cudaMemcpyDeviceToHost, s));
        fail(cudaStreamSynchronize(s));
        fail(cudaGetLastError());
    }
    for (void *p : {(void *)d_sig, (void *)d_rank, (void *)d_off, (void *)d_n, (void *)d_order, (void *)d_status, (void *)d_rs, (void *)d_rc,
                    (void *)d_shift, (void *)d_scale, (void *)d_nan, (void *)d_scratch})
        if (p) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

// ---- eventalign (SURVEY s.8 row f1) ------------------------------------------------------------------------------
// eventalign runs window-parallel (eventalign_wp.cu: one warp per window, speculative chain + verification rounds), which
// makes a read's latency one window instead of its length; DNB_EA_WINDOW_PARALLEL=0 selects the read-serial kernel
// (eventalign.cu: one warp per read), kept as the cross-check both are tested against
static bool ea_window_parallel() {
    static const bool on = !(getenv("DNB_EA_WINDOW_PARALLEL") != nullptr && getenv("DNB_EA_WINDOW_PARALLEL")[0] == '0');
    return on;
}

static thread_local double g_ea_kernel_ms = 0.0;
double dnb_eventalign_last_kernel_ms(void) { return g_ea_kernel_ms; }

static thread_local double g_ft_kernel_ms = 0.0;
double dnb_features_last_kernel_ms(void) { return g_ft_kernel_ms; }

// eventalign, and (feats != NULL) the DNN input tensors from its records in the same device pass (row f2)
static int eventalign_impl(dnb_ctx *ctx, const dnb_eventalign_desc *reads, const dnb_feature_desc *feats, size_t n_reads,
                           uint32_t window, dnb_eventalign_rec *recs, const uint64_t *rec_off, uint32_t *n_recs,
                           int *status, const dnb_feature_tensors *out, const uint64_t *pos_off, uint32_t *n_pos) {
    if (!ctx || (!reads && n_reads) || !rec_off || !n_recs || !status) return DNB_ERR_ARG;
    if (!feats && !recs && n_reads && rec_off[n_reads]) return DNB_ERR_ARG;
    if (feats && (!out || !pos_off || !n_pos)) return DNB_ERR_ARG;
    if (feats && n_reads && pos_off[n_reads] &&
        (!out->signal || !out->core || !out->residual || !out->coords || !out->ref_index || !out->query_index || !out->quality))
        return DNB_ERR_ARG;
    if (window < DNB_K + 2 || window > 60) { g_last_error = "eventalign window must be in [11, 60]"; return DNB_ERR_ARG; }
    if (!ctx->model[DNB_MODEL_PORE].loaded) return DNB_ERR_MODEL;
    if (n_reads == 0) return DNB_OK;
    if (n_reads >= (1ull << 32)) return DNB_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    const size_t R = n_reads;
    // ---- pack (host) ----
    std::vector<uint64_t> ref_off(R + 1, 0), al_off(R + 1, 0), ev_off(R + 1, 0);
    for (size_t i = 0; i < R; i++) {
        const dnb_eventalign_desc &d = reads[i];
        if ((d.ref_len && (!d.ref || !d.ref_to_query)) || (d.n_align && !d.align_pairs) || (d.n_events && !d.event_mean)) {
            g_last_error = "eventalign descriptor " + std::to_string(i) + " is incomplete";
            return DNB_ERR_ARG;
        }
        ref_off[i + 1] = ref_off[i] + d.ref_len;
        al_off[i + 1] = al_off[i] + d.n_align;
        ev_off[i + 1] = ev_off[i] + d.n_events;
    }
    const uint64_t tot_ref = ref_off[R], tot_al = al_off[R], tot_ev = ev_off[R], tot_rec = rec_off[R];
    std::vector<char> h_ref(tot_ref ? tot_ref : 1);
    std::vector<int32_t> h_r2q(tot_ref ? tot_ref : 1);
    std::vector<uint32_t> h_pairs(tot_al ? 2 * tot_al : 2);
    std::vector<float> h_evm(tot_ev ? tot_ev : 1);
    std::vector<double> h_shift(R), h_scale(R), h_trans(4 * R);
    std::vector<int> h_status(R, DNB_READ_OK);
    // eln() of HMM_TransitionProbs_DNA_R10 {0.3, 0.7, 0.999, 0.0025, 0.001, 0.001} (src/config.h:42, alignment.cpp:199-204)
    const double d2d = log(0.3), d2m = log(0.7), i2m = log(0.999), m2d = log(0.0025), m2i = log(0.001), i2i = log(0.001);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)R; ii++) {
        const size_t i = (size_t)ii;
        const dnb_eventalign_desc &d = reads[i];
        if (d.ref_len) { memcpy(&h_ref[ref_off[i]], d.ref, d.ref_len); memcpy(&h_r2q[ref_off[i]], d.ref_to_query, 4ull * d.ref_len); }
        if (d.n_align) memcpy(&h_pairs[2 * al_off[i]], d.align_pairs, 8ull * d.n_align);
        if (d.n_events) memcpy(&h_evm[ev_off[i]], d.event_mean, 4ull * d.n_events);
        h_shift[i] = d.shift; h_scale[i] = d.scale;
        // alignment.cpp:207-210.  eln throws NegativeLog for x < 0 and the next eln throws for a NaN argument (x == 0)
        const double x = 1. - (1. / d.events_per_base);
        bool ok = x > 0.0 && d.ref_len >= DNB_K;
        // every (event, k-mer) the kernel dereferences must exist in the caller's arrays
        for (uint32_t j = 0; ok && j < d.n_align; j++) ok = d.align_pairs[2 * j] < d.n_events;
        if (!ok) { h_status[i] = DNB_READ_UNDEFINED; h_trans[4 * i] = h_trans[4 * i + 1] = h_trans[4 * i + 2] = h_trans[4 * i + 3] = 0.0; continue; }
        const double m12m1_int = log(x);
        const double m12m1_ext = eln_nothrow(1.0 - m2d - m2i - m12m1_int);        // quirk Q10: logs inside, on purpose
        h_trans[4 * i + 0] = m12m1_int;
        h_trans[4 * i + 1] = m12m1_ext;
        h_trans[4 * i + 2] = dnb_lnSum(m12m1_ext, m12m1_int);                      // externalOrInternalM12M1
        h_trans[4 * i + 3] = dnb_lnSum(m12m1_ext, m2d);                            // externalM12M1orD
    }
    // ---- feature inputs (row f2): offsets only; the signal and the event starts go up straight from the caller's arrays
    std::vector<uint64_t> raw_off, called_off;
    std::vector<uint8_t> h_kind, h_rev;
    std::vector<float> h_doff, h_dscl;
    std::vector<uint32_t> h_rstart, h_rend;
    uint64_t tot_f32 = 0, tot_i16 = 0, tot_called = 0, tot_pos = 0;
    if (feats) {
        raw_off.assign(R, 0); called_off.assign(R + 1, 0);
        h_kind.assign(R, 0); h_rev.assign(R, 0); h_doff.assign(R, 0.f); h_dscl.assign(R, 1.f);
        h_rstart.assign(R, 0); h_rend.assign(R, 0);
        for (size_t i = 0; i < R; i++) {
            const dnb_feature_desc &f = feats[i];
            if ((f.n_samples && !f.raw_pA && !f.raw_dac) || (reads[i].n_events && !f.event_start) || (f.n_called && !f.called)) {
                g_last_error = "feature descriptor " + std::to_string(i) + " is incomplete";
                return DNB_ERR_ARG;
            }
            h_kind[i] = f.raw_pA ? 0 : 1;
            // 16-byte aligned read starts in both arrays
            if (f.raw_pA) { raw_off[i] = tot_f32; tot_f32 += (f.n_samples + 3) & ~3ull; }
            else { raw_off[i] = tot_i16; tot_i16 += (f.n_samples + 7) & ~7ull; }
            h_doff[i] = f.dac_offset; h_dscl[i] = f.dac_scale;
            h_rev[i] = f.is_reverse ? 1 : 0; h_rstart[i] = f.ref_start; h_rend[i] = f.ref_end;
            called_off[i + 1] = called_off[i] + f.n_called;
            // every sample an event of this read can address must exist (event_start is non-decreasing in a valid read)
            if (h_status[i] == DNB_READ_OK && reads[i].n_events) {
                bool ok = f.event_start[reads[i].n_events] <= f.n_samples;
                for (uint32_t j = 0; ok && j < reads[i].n_events; j++) ok = f.event_start[j] <= f.event_start[j + 1];
                if (!ok) h_status[i] = DNB_READ_UNDEFINED;
            }
        }
        tot_called = called_off[R];
        tot_pos = pos_off[R];
    }
    // ---- device ----
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = DNB_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == DNB_OK) { g_last_error = cudaGetErrorString(e); rc = DNB_ERR_CUDA; } };
    std::vector<void *> owned;
    auto dalloc = [&](size_t bytes) -> void * { void *p = nullptr; fail(pool_alloc(ctx, &p, bytes ? bytes : 16, s)); if (p) owned.push_back(p); return p; };
    DnbEaArgs a = {};
    a.n_reads = (uint32_t)R; a.window = window; a.t_max = 4096;
    unsigned grid = dnb_eventalign_grid(ctx->cfg.device);
    const unsigned wpb = dnb_eventalign_warps_per_block();
    if ((size_t)grid * wpb > R) grid = (unsigned)((R + wpb - 1) / wpb);
    const size_t warps = (size_t)grid * wpb;
    uint64_t *d_ref_off = (uint64_t *)dalloc((R + 1) * 8), *d_al_off = (uint64_t *)dalloc((R + 1) * 8);
    uint64_t *d_ev_off = (uint64_t *)dalloc((R + 1) * 8), *d_rec_off = (uint64_t *)dalloc((R + 1) * 8);
    char *d_ref = (char *)dalloc(tot_ref);
    int32_t *d_r2q = (int32_t *)dalloc(tot_ref * 4);
    uint32_t *d_pairs = (uint32_t *)dalloc(tot_al * 8);
    float *d_evm = (float *)dalloc(tot_ev * 4);
    double *d_shift = (double *)dalloc(R * 8), *d_scale = (double *)dalloc(R * 8), *d_trans = (double *)dalloc(R * 32);
    dnb_eventalign_rec *d_recs = (dnb_eventalign_rec *)dalloc(tot_rec * sizeof(dnb_eventalign_rec));
    uint32_t *d_nrec = (uint32_t *)dalloc(R * 4);
    int *d_status = (int *)dalloc(R * 4);
    unsigned int *d_next = (unsigned int *)dalloc(4);
    a.scratch_obs = (double *)dalloc(warps * a.t_max * 8);
    a.scratch_ev = (uint32_t *)dalloc(warps * a.t_max * 4);
    a.scratch_bt = (uint8_t *)dalloc(warps * a.t_max * dnb_eventalign_bt_row_bytes());
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    fail(cudaEventCreate(&e0)); fail(cudaEventCreate(&e1)); fail(cudaEventCreate(&e2)); fail(cudaEventCreate(&e3));
    if (rc == DNB_OK) {
        auto up = [&](void *dst, const void *src, size_t bytes) { if (bytes) fail(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s)); };
        up(d_ref_off, ref_off.data(), (R + 1) * 8); up(d_al_off, al_off.data(), (R + 1) * 8);
        up(d_ev_off, ev_off.data(), (R + 1) * 8); up(d_rec_off, rec_off, (R + 1) * 8);
        up(d_ref, h_ref.data(), tot_ref); up(d_r2q, h_r2q.data(), tot_ref * 4); up(d_pairs, h_pairs.data(), tot_al * 8);
        up(d_evm, h_evm.data(), tot_ev * 4); up(d_shift, h_shift.data(), R * 8); up(d_scale, h_scale.data(), R * 8);
        up(d_trans, h_trans.data(), R * 32); up(d_status, h_status.data(), R * 4);
        fail(cudaMemsetAsync(d_next, 0, 4, s));
        a.ref_off = d_ref_off; a.ref = d_ref; a.r2q = d_r2q; a.al_off = d_al_off; a.pairs = reinterpret_cast<const uint2 *>(d_pairs);
        a.ev_off = d_ev_off; a.ev_mean = d_evm; a.shift = d_shift; a.scale = d_scale; a.trans = d_trans;
        a.model_mean = ctx->model[DNB_MODEL_PORE].d_mean;
        // normalPDF's constants for the static sigma of the ONT table (data_IO.cpp:170, probability.cpp:147), host libm
        a.two_sigma2 = 2.0 * pow(0.14, 2.0);
        a.inv_two_sigma2 = 1.0 / a.two_sigma2;
        a.c = 1.0 / sqrt(2.0 * pow(0.14, 2.0) * M_PI);
        a.ln_c = log(a.c);
        a.d2d = d2d; a.d2m = d2m; a.i2m = i2m; a.m2d = m2d; a.m2i = m2i; a.i2i = i2i;
        a.rec_off = d_rec_off; a.recs = d_recs; a.n_rec = d_nrec; a.status = d_status; a.next_read = d_next;
        fail(cudaEventRecord(e0, s));
        if (ea_window_parallel()) {
            struct PoolUser { dnb_ctx *ctx; cudaStream_t s; std::vector<void *> *owned; } pu{ctx, s, &owned};
            DnbAlloc al{[](void *u, size_t bytes) -> void * {
                            PoolUser *q = (PoolUser *)u;
                            void *p = nullptr;
                            if (pool_alloc(q->ctx, &p, bytes, q->s) != cudaSuccess) { cudaGetLastError(); return nullptr; }
                            q->owned->push_back(p);
                            return p;
                        }, &pu};
            fail(dnb_run_eventalign_wp(a, tot_ref, tot_al, ctx->cfg.device, s, al, nullptr));
        }
        else dnb_launch_eventalign(a, grid, s);
        fail(cudaEventRecord(e1, s));
        fail(cudaGetLastError());
        if (feats && rc == DNB_OK) {
            DnbFeatArgs f = {};
            f.n_reads = (uint32_t)R;
            f.rec_off = d_rec_off; f.recs = d_recs; f.n_rec = d_nrec; f.status = d_status;
            f.ref_off = d_ref_off; f.ref = d_ref; f.r2q = d_r2q;
            uint64_t *d_raw_off = (uint64_t *)dalloc(R * 8), *d_called_off = (uint64_t *)dalloc((R + 1) * 8);
            uint64_t *d_pos_off = (uint64_t *)dalloc((R + 1) * 8);
            uint8_t *d_kind = (uint8_t *)dalloc(R), *d_rev = (uint8_t *)dalloc(R);
            float *d_doff = (float *)dalloc(R * 4), *d_dscl = (float *)dalloc(R * 4);
            uint32_t *d_rstart = (uint32_t *)dalloc(R * 4), *d_rend = (uint32_t *)dalloc(R * 4);
            float *d_f32 = (float *)dalloc(tot_f32 * 4);
            int16_t *d_i16 = (int16_t *)dalloc(tot_i16 * 2);
            uint32_t *d_es = (uint32_t *)dalloc((tot_ev + R) * 4);
            uint32_t *d_called = (uint32_t *)dalloc(tot_called * 4);
            float *d_signal = (float *)dalloc(tot_pos * DNB_RAWDEPTH * 4);
            float *d_core = (float *)dalloc(tot_pos * 4), *d_resid = (float *)dalloc(tot_pos * 4);
            uint32_t *d_coords = (uint32_t *)dalloc(tot_pos * 4), *d_ri = (uint32_t *)dalloc(tot_pos * 4), *d_qi = (uint32_t *)dalloc(tot_pos * 4);
            int32_t *d_qual = (int32_t *)dalloc(tot_pos * 4);
            uint32_t *d_npos = (uint32_t *)dalloc(R * 4);
            unsigned int *d_next2 = (unsigned int *)dalloc(4);
            if (rc == DNB_OK) {
                up(d_raw_off, raw_off.data(), R * 8); up(d_called_off, called_off.data(), (R + 1) * 8);
                up(d_pos_off, pos_off, (R + 1) * 8);
                up(d_kind, h_kind.data(), R); up(d_rev, h_rev.data(), R);
                up(d_doff, h_doff.data(), R * 4); up(d_dscl, h_dscl.data(), R * 4);
                up(d_rstart, h_rstart.data(), R * 4); up(d_rend, h_rend.data(), R * 4);
                for (size_t i = 0; i < R; i++) {
                    const dnb_feature_desc &fd = feats[i];
                    if (fd.raw_pA) up(d_f32 + raw_off[i], fd.raw_pA, fd.n_samples * 4);
                    else up(d_i16 + raw_off[i], fd.raw_dac, fd.n_samples * 2);
                    if (reads[i].n_events) up(d_es + ev_off[i] + i, fd.event_start, (reads[i].n_events + 1ull) * 4);
                    if (fd.n_called) up(d_called + called_off[i], fd.called, fd.n_called * 4ull);
                }
                fail(cudaMemsetAsync(d_next2, 0, 4, s));
                f.raw_off = d_raw_off; f.raw_kind = d_kind; f.raw_f32 = d_f32; f.raw_i16 = d_i16;
                f.dac_offset = d_doff; f.dac_scale = d_dscl; f.ev_off = d_ev_off; f.ev_start = d_es;
                f.shift = d_shift; f.scale = d_scale; f.is_reverse = d_rev; f.ref_start = d_rstart; f.ref_end = d_rend;
                f.called_off = d_called_off; f.called = d_called; f.pos_off = d_pos_off;
                f.signal = d_signal; f.core = d_core; f.residual = d_resid; f.coords = d_coords; f.ref_index = d_ri;
                f.query_index = d_qi; f.quality = d_qual; f.n_pos = d_npos; f.next_read = d_next2;
                fail(cudaEventRecord(e2, s));
                dnb_launch_features(f, dnb_features_grid(ctx->cfg.device), s);
                fail(cudaEventRecord(e3, s));
                fail(cudaGetLastError());
                auto down = [&](void *dst, const void *src, size_t bytes) { if (bytes) fail(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s)); };
                down(out->signal, d_signal, tot_pos * DNB_RAWDEPTH * 4);
                down(out->core, d_core, tot_pos * 4); down(out->residual, d_resid, tot_pos * 4);
                down(out->coords, d_coords, tot_pos * 4); down(out->ref_index, d_ri, tot_pos * 4);
                down(out->query_index, d_qi, tot_pos * 4); down(out->quality, d_qual, tot_pos * 4);
                down(n_pos, d_npos, R * 4);
            }
        }
        if (tot_rec && recs) fail(cudaMemcpyAsync(recs, d_recs, tot_rec * sizeof(dnb_eventalign_rec), cudaMemcpyDeviceToHost, s));
        fail(cudaMemcpyAsync(n_recs, d_nrec, R * 4, cudaMemcpyDeviceToHost, s));
        fail(cudaMemcpyAsync(status, d_status, R * 4, cudaMemcpyDeviceToHost, s));
        fail(cudaStreamSynchronize(s));
        float ms = 0.f;
        if (rc == DNB_OK && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) g_ea_kernel_ms = ms;
        if (rc == DNB_OK && feats && cudaEventElapsedTime(&ms, e2, e3) == cudaSuccess) g_ft_kernel_ms = ms;
    }
    for (void *p : owned) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    if (e3) cudaEventDestroy(e3);
    cudaStreamDestroy(s);
    return rc;
}

// ---- resident form (rows f1 + f2 without a host round trip): everything normaliseEvents left in HBM is read in place
int dnb_batch_eventalign_features(dnb_batch *b, const dnb_read_extra *extra, uint32_t window, int want_records) {
    if (!b || (!extra && b->R)) return DNB_ERR_ARG;
    if (window < DNB_K + 2 || window > 60) { g_last_error = "eventalign window must be in [11, 60]"; return DNB_ERR_ARG; }
    if (!b->ran || b->want_table || !b->have_work) return DNB_ERR_STATE;
    dnb_ctx *ctx = b->ctx;
    if (!ctx->model[DNB_MODEL_PORE].loaded) return DNB_ERR_MODEL;
    TRY(fetch(b));                       // per-read scalars on the host, (dense format) forward alignment pairs on the device
    CK(cudaSetDevice(ctx->cfg.device));
    if (!b->have_dev_pairs) {            // compact result format: eventalign still reads dense pairs, in HBM only
        TRY(walloc(b, &b->w.out_pairs, 2 * b->tot_out));
        dnb_launch_compact_alignment(make_view(b), b->w.al_off, b->w.al_rev, b->w.n_align, b->w.out_off, b->w.out_pairs, b->stream);
        b->have_dev_pairs = true;
    }
    const size_t R = b->R;
    cudaStream_t s = b->stream;
    HostRes &h = b->h;
    Work &w = b->w;
    Stage2 &S = b->s2;
    S = Stage2{};
    S.want_records = want_records != 0;
    for (size_t i = 0; i < R; i++)
        if ((b->rlen[i] && !extra[i].ref_to_query) || (extra[i].n_called && !extra[i].called)) {
            g_last_error = "dnb_read_extra " + std::to_string(i) + " is incomplete";
            return DNB_ERR_ARG;
        }
    // ---- host: per-read constants and shapes ----
    const double d2d = log(0.3), d2m = log(0.7), i2m = log(0.999), m2d = log(0.0025), m2i = log(0.001), i2i = log(0.001);
    std::vector<double> h_trans(4 * R, 0.0);
    std::vector<int> h_status(R);
    std::vector<uint8_t> h_rev(R);
    std::vector<uint32_t> h_rstart(R), h_rend(R);
    std::vector<uint64_t> called_off(R + 1, 0);
    S.rec_off.assign(R + 1, 0); S.pos_off.assign(R + 1, 0);
    for (size_t i = 0; i < R; i++) {
        int st = h.status[i];
        if (st == DNB_READ_OK) {
            // r.scalings.eventsPerBase (event_handling.cpp:606) -> alignment.cpp:207-210; eln throws for x <= 0
            const double epb = (double)h.et_n[i] / (double)((int64_t)b->qlen[i] - DNB_K);
            const double x = 1. - (1. / epb);
            if (!(x > 0.0) || b->rlen[i] < DNB_K) st = DNB_READ_UNDEFINED;
            else {
                const double m12m1_int = log(x);
                const double m12m1_ext = eln_nothrow(1.0 - m2d - m2i - m12m1_int);    // quirk Q10
                h_trans[4 * i + 0] = m12m1_int;
                h_trans[4 * i + 1] = m12m1_ext;
                h_trans[4 * i + 2] = dnb_lnSum(m12m1_ext, m12m1_int);
                h_trans[4 * i + 3] = dnb_lnSum(m12m1_ext, m2d);
            }
        }
        h_status[i] = st;
        h_rev[i] = extra[i].is_reverse ? 1 : 0; h_rstart[i] = extra[i].ref_start; h_rend[i] = extra[i].ref_end;
        called_off[i + 1] = called_off[i] + extra[i].n_called;
        S.rec_off[i + 1] = S.rec_off[i] + (st == DNB_READ_OK ? (uint64_t)h.n_align[i] + 64 : 0);
        S.pos_off[i + 1] = S.pos_off[i] + (st == DNB_READ_OK ? (uint64_t)b->rlen[i] - DNB_K + 2 : 0);
    }
    const uint64_t tot_rec = S.rec_off[R], tot_pos = S.pos_off[R], tot_called = called_off[R], tot_r = b->tot_r;
    // refToQuery / called keys through pinned staging (the only per-base upload of this stage)
    int32_t *st_r2q = (int32_t *)ctx->pinned.acquire((tot_r ? tot_r : 1) * 4);
    uint32_t *st_called = (uint32_t *)ctx->pinned.acquire((tot_called ? tot_called : 1) * 4);
    struct StagingGuard {
        dnb_batch *b; void *p0, *p1;
        ~StagingGuard() { wait_stream(b); if (p0) b->ctx->pinned.release(p0); if (p1) b->ctx->pinned.release(p1); }
    } staging_guard{b, st_r2q, st_called};
    if (!st_r2q || !st_called) { g_last_error = "pinned staging allocation failed"; return DNB_ERR_NOMEM; }
#pragma omp parallel for schedule(dynamic, 16)
    for (long long ii = 0; ii < (long long)R; ii++) {
        const size_t i = (size_t)ii;
        if (b->rlen[i]) memcpy(st_r2q + b->r_off[i], extra[i].ref_to_query, 4ull * b->rlen[i]);
        if (extra[i].n_called) memcpy(st_called + called_off[i], extra[i].called, 4ull * extra[i].n_called);
    }
    // ---- device ----
    int32_t *d_r2q; uint32_t *d_called, *d_rstart, *d_rend, *d_nrec, *d_npos; uint8_t *d_rev, *d_kind;
    uint64_t *d_called_off, *d_rec_off, *d_pos_off; double *d_trans; int *d_status; unsigned int *d_next;
    dnb_eventalign_rec *d_recs;
    float *d_signal, *d_core, *d_resid; uint32_t *d_coords, *d_ri, *d_qi; int32_t *d_qual;
    Arena ar;
    ar.add(&d_r2q, tot_r); ar.add(&d_called, tot_called); ar.add(&d_called_off, R + 1);
    ar.add(&d_rstart, R); ar.add(&d_rend, R); ar.add(&d_rev, R); ar.add(&d_kind, R);
    ar.add(&d_rec_off, R + 1); ar.add(&d_pos_off, R + 1); ar.add(&d_trans, 4 * R);
    ar.add(&d_status, R); ar.add(&d_nrec, R); ar.add(&d_npos, R); ar.add(&d_next, 2);
    ar.add(&d_recs, tot_rec);
    ar.add(&d_signal, tot_pos * DNB_RAWDEPTH); ar.add(&d_core, tot_pos); ar.add(&d_resid, tot_pos);
    ar.add(&d_coords, tot_pos); ar.add(&d_ri, tot_pos); ar.add(&d_qi, tot_pos); ar.add(&d_qual, tot_pos);
    DnbEaArgs a = {};
    a.n_reads = (uint32_t)R; a.window = window; a.t_max = 4096;
    unsigned grid = dnb_eventalign_grid(ctx->cfg.device);
    const unsigned wpb = dnb_eventalign_warps_per_block();
    if ((size_t)grid * wpb > R) grid = (unsigned)((R + wpb - 1) / wpb);
    const size_t warps = (size_t)grid * wpb;
    if (!ea_window_parallel()) {       // per-warp scratch rows of the read-serial kernel (the window-parallel one sizes its own)
        ar.add(&a.scratch_obs, warps * a.t_max); ar.add(&a.scratch_ev, warps * a.t_max);
        ar.add(&a.scratch_bt, warps * a.t_max * dnb_eventalign_bt_row_bytes());
    }
    TRY(ar.commit(b, b->work_allocs));
    TRY(h2d(b, d_r2q, st_r2q, tot_r)); TRY(h2d(b, d_called, st_called, tot_called));
    TRY(h2d(b, d_called_off, called_off.data(), R + 1)); TRY(h2d(b, d_rstart, h_rstart.data(), R));
    TRY(h2d(b, d_rend, h_rend.data(), R)); TRY(h2d(b, d_rev, h_rev.data(), R));
    TRY(h2d(b, d_rec_off, S.rec_off.data(), R + 1)); TRY(h2d(b, d_pos_off, S.pos_off.data(), R + 1));
    TRY(h2d(b, d_trans, h_trans.data(), 4 * R)); TRY(h2d(b, d_status, h_status.data(), R));
    CK(cudaMemsetAsync(d_kind, b->i16 ? 1 : 0, R, s));
    CK(cudaMemsetAsync(d_next, 0, 8, s));
    S.bytes[0] = 4 * tot_r + 4 * tot_called + 8 * 3 * (R + 1) + R * (4 + 4 + 1 + 32 + 4);
    a.ref_off = b->d_r_off; a.ref = b->d_ref; a.r2q = d_r2q;
    a.al_off = w.out_off; a.pairs = reinterpret_cast<const uint2 *>(w.out_pairs);
    a.ev_off = b->d_ev_off; a.ev_mean = w.ev_mean; a.shift = w.shift; a.scale = w.scale; a.trans = d_trans;
    a.model_mean = ctx->model[DNB_MODEL_PORE].d_mean;
    a.two_sigma2 = 2.0 * pow(0.14, 2.0);
        a.inv_two_sigma2 = 1.0 / a.two_sigma2;
    a.c = 1.0 / sqrt(2.0 * pow(0.14, 2.0) * M_PI);
    a.ln_c = log(a.c);
    a.d2d = d2d; a.d2m = d2m; a.i2m = i2m; a.m2d = m2d; a.m2i = m2i; a.i2i = i2i;
    a.rec_off = d_rec_off; a.recs = d_recs; a.n_rec = d_nrec; a.status = d_status; a.next_read = d_next;
    a.order = b->d_order;
    CK(cudaEventRecord(b->ev[0], s));
    if (ea_window_parallel()) {
        // workspace from the batch's device cache (dropped with the rest of the workspace), sleeping waits
        DnbAlloc al{[](void *u, size_t bytes) -> void * {
                        dnb_batch *bb = (dnb_batch *)u;
                        uint8_t *p = nullptr;
                        return walloc(bb, &p, bytes) == DNB_OK ? (void *)p : nullptr;
                    }, b};
        CK(dnb_run_eventalign_wp(a, tot_r, b->tot_out, ctx->cfg.device, s, al, b->sync_ev));
    }
    else dnb_launch_eventalign(a, grid, s);
    CK(cudaEventRecord(b->ev[1], s));
    DnbFeatArgs f = {};
    f.n_reads = (uint32_t)R;
    f.rec_off = d_rec_off; f.recs = d_recs; f.n_rec = d_nrec; f.status = d_status;
    f.ref_off = b->d_r_off; f.ref = b->d_ref; f.r2q = d_r2q;
    f.raw_off = b->d_raw_off; f.raw_kind = d_kind;
    f.raw_f32 = b->i16 ? nullptr : (const float *)b->d_raw; f.raw_i16 = b->i16 ? (const int16_t *)b->d_raw : nullptr;
    f.dac_offset = b->d_dac_off; f.dac_scale = b->d_dac_scl;
    f.ev_off = b->d_ev_off; f.ev_start = w.ev_start; f.shift = w.shift; f.scale = w.scale;
    f.is_reverse = d_rev; f.ref_start = d_rstart; f.ref_end = d_rend; f.called_off = d_called_off; f.called = d_called;
    f.pos_off = d_pos_off; f.signal = d_signal; f.core = d_core; f.residual = d_resid; f.coords = d_coords;
    f.ref_index = d_ri; f.query_index = d_qi; f.quality = d_qual; f.n_pos = d_npos; f.next_read = d_next + 1;
    f.order = b->d_order;
    CK(cudaEventRecord(b->ev[2], s));
    dnb_launch_features(f, dnb_features_grid(ctx->cfg.device), s);
    CK(cudaEventRecord(b->ev[3], s));
    CK(cudaGetLastError());
    // ---- results -> pinned host ----
    TRY(ralloc(b, &S.h_n_rec, R)); TRY(ralloc(b, &S.h_n_pos, R)); TRY(ralloc(b, &S.h_status, R));
    TRY(ralloc(b, &S.h_signal, tot_pos * DNB_RAWDEPTH)); TRY(ralloc(b, &S.h_core, tot_pos)); TRY(ralloc(b, &S.h_resid, tot_pos));
    TRY(ralloc(b, &S.h_coords, tot_pos)); TRY(ralloc(b, &S.h_ri, tot_pos)); TRY(ralloc(b, &S.h_qi, tot_pos));
    TRY(ralloc(b, &S.h_qual, tot_pos));
    TRY(d2h(b, S.h_n_rec, d_nrec, R)); TRY(d2h(b, S.h_n_pos, d_npos, R)); TRY(d2h(b, S.h_status, d_status, R));
    TRY(d2h(b, S.h_signal, d_signal, tot_pos * DNB_RAWDEPTH)); TRY(d2h(b, S.h_core, d_core, tot_pos));
    TRY(d2h(b, S.h_resid, d_resid, tot_pos)); TRY(d2h(b, S.h_coords, d_coords, tot_pos)); TRY(d2h(b, S.h_ri, d_ri, tot_pos));
    TRY(d2h(b, S.h_qi, d_qi, tot_pos)); TRY(d2h(b, S.h_qual, d_qual, tot_pos));
    S.bytes[1] = tot_pos * (4ull * DNB_RAWDEPTH + 28) + 12 * R;
    if (S.want_records) {
        TRY(ralloc(b, &S.h_recs, tot_rec));
        TRY(d2h(b, S.h_recs, d_recs, tot_rec));
        S.bytes[1] += tot_rec * sizeof(dnb_eventalign_rec);
    }
    CK(wait_stream(b));
    CK(cudaGetLastError());
    float t;
    cudaEventElapsedTime(&t, b->ev[0], b->ev[1]); S.ms[0] = t;
    cudaEventElapsedTime(&t, b->ev[2], b->ev[3]); S.ms[1] = t;
    S.done = true;
    return DNB_OK;
}

// dnb_submit with the resident stage appended: the pipelined (thread-safe, gated) form of the whole chain
int dnb_submit_chain(dnb_ctx *ctx, const dnb_read_desc *reads, const dnb_read_extra *extra, size_t n_reads,
                     uint32_t window, int want_records, dnb_batch **batch) {
    dnb_batch *b = nullptr;
    if (!ctx || !batch || (!extra && n_reads)) return DNB_ERR_ARG;
    uint64_t load = 0;
    ctx = pick_device(ctx, reads, n_reads, &load);
    int rc = upload(ctx, reads, n_reads, false, &b, /*gated=*/true);
    if (rc != DNB_OK) { ctx->inflight.fetch_sub(load, std::memory_order_relaxed); return rc; }
    b->load = load;
    {
        HostTrace hg("submit");
        StageHold hold(ctx->gate_compute);
        hg.tick(PH_RUN_GATE);
        rc = run(b);
    }
    if (rc == DNB_OK) {
        HostTrace hg("submit");
        StageHold hold(ctx->gate_fetch);
        hg.tick(PH_FETCH_GATE);
        rc = fetch(b);
    }
    if (rc == DNB_OK) {
        StageHold hold(ctx->gate_compute);
        rc = dnb_batch_eventalign_features(b, extra, window, want_records);
    }
    if (rc != DNB_OK) { free_batch(b); return rc; }
    drop_work(b);
    *batch = b;
    return DNB_OK;
}

// ---- resident analogue stage (SURVEY s.8 row a15; detect.cpp:885): sites, event ranges and forward passes on the device
int dnb_batch_analogue_llr(dnb_batch *b, const dnb_read_extra *extra, uint32_t window) {
    if (!b || (!extra && b->R)) return DNB_ERR_ARG;
    if (window < 5 || window > 16) { g_last_error = "llAcrossRead window must be in [5, 16]"; return DNB_ERR_ARG; }
    if (!b->ran || b->want_table || !b->have_work) return DNB_ERR_STATE;
    dnb_ctx *ctx = b->ctx;
    if (!ctx->model[DNB_MODEL_UNLABELLED].loaded || !ctx->model[DNB_MODEL_ANALOGUE].loaded) return DNB_ERR_MODEL;
    CK(cudaSetDevice(ctx->cfg.device));
    const size_t R = b->R;
    cudaStream_t s = b->stream;
    Work &w = b->w;
    Stage3 &S = b->s3;
    S = Stage3{};
    for (size_t i = 0; i < R; i++)
        if (b->rlen[i] && !extra[i].ref_to_query) { g_last_error = "dnb_read_extra " + std::to_string(i) + " is incomplete"; return DNB_ERR_ARG; }
    const uint64_t tot_r = b->tot_r;
    int32_t *st_r2q = (int32_t *)ctx->pinned.acquire((tot_r ? tot_r : 1) * 4);
    uint8_t *st_rev = (uint8_t *)ctx->pinned.acquire(R ? R : 1);
    struct StagingGuard {
        dnb_batch *b; void *p0, *p1;
        ~StagingGuard() { wait_stream(b); if (p0) b->ctx->pinned.release(p0); if (p1) b->ctx->pinned.release(p1); }
    } staging_guard{b, st_r2q, st_rev};
    if (!st_r2q || !st_rev) { g_last_error = "pinned staging allocation failed"; return DNB_ERR_NOMEM; }
#pragma omp parallel for schedule(dynamic, 16)
    for (long long ii = 0; ii < (long long)R; ii++) {
        const size_t i = (size_t)ii;
        if (b->rlen[i]) memcpy(st_r2q + b->r_off[i], extra[i].ref_to_query, 4ull * b->rlen[i]);
        st_rev[i] = extra[i].is_reverse ? 1 : 0;
    }
    DnbLlrArgs a;
    memset(&a, 0, sizeof(a));
    int32_t *d_r2q = nullptr; uint8_t *d_rev = nullptr; uint64_t *d_poi_off = nullptr;
    {
        Arena ar;
        ar.add(&d_r2q, tot_r); ar.add(&d_rev, R); ar.add(&a.poi, tot_r); ar.add(&a.j_begin, tot_r); ar.add(&a.j_end, tot_r);
        ar.add(&a.n_poi, R); ar.add(&d_poi_off, R + 1); ar.add(&a.next_site, 1);
        TRY(ar.commit(b, b->work_allocs));
    }
    TRY(h2d(b, d_r2q, st_r2q, tot_r)); TRY(h2d(b, d_rev, st_rev, R));
    a.window = window; a.r2q = d_r2q; a.is_reverse = d_rev;
    a.al_off = w.al_off; a.al_pairs_rev = w.al_rev; a.n_align = w.n_align; a.shift = w.shift; a.scale = w.scale;
    const DnbBatchView v = make_view(b);
    CK(cudaEventRecord(b->ev[0], s));
    dnb_launch_llr_sites(v, a, s);
    CK(cudaEventRecord(b->ev[1], s));
    TRY(ralloc(b, &S.h_n_poi, R));
    TRY(d2h(b, S.h_n_poi, a.n_poi, R));
    CK(wait_stream(b));
    CK(cudaGetLastError());
    S.poi_off.assign(R + 1, 0);
    for (size_t i = 0; i < R; i++) S.poi_off[i + 1] = S.poi_off[i] + S.h_n_poi[i];
    const uint64_t n_sites = S.poi_off[R];
    {
        Arena ar;
        ar.add(&a.out_pos, n_sites); ar.add(&a.out_n_events, n_sites); ar.add(&a.out_a, n_sites); ar.add(&a.out_t, n_sites);
        TRY(ar.commit(b, b->work_allocs));
    }
    TRY(h2d(b, d_poi_off, S.poi_off.data(), R + 1));
    a.poi_off = d_poi_off;
    CK(cudaMemsetAsync(a.next_site, 0, sizeof(unsigned long long), s));
    CK(cudaEventRecord(b->ev[2], s));
    dnb_launch_llr_forward(v, a, n_sites, ctx->model[DNB_MODEL_UNLABELLED].dev(), ctx->model[DNB_MODEL_ANALOGUE].dev(),
                           ctx->cfg.device, s);
    CK(cudaEventRecord(b->ev[3], s));
    TRY(ralloc(b, &S.h_pos, n_sites)); TRY(ralloc(b, &S.h_nev, n_sites)); TRY(ralloc(b, &S.h_a, n_sites)); TRY(ralloc(b, &S.h_t, n_sites));
    TRY(d2h(b, S.h_pos, a.out_pos, n_sites)); TRY(d2h(b, S.h_nev, a.out_n_events, n_sites));
    TRY(d2h(b, S.h_a, a.out_a, n_sites)); TRY(d2h(b, S.h_t, a.out_t, n_sites));
    CK(wait_stream(b));
    CK(cudaGetLastError());
    float t;
    cudaEventElapsedTime(&t, b->ev[0], b->ev[1]); S.ms[0] = t;
    cudaEventElapsedTime(&t, b->ev[2], b->ev[3]); S.ms[1] = t;
    uint64_t calls = 0, obs = 0;
    for (uint64_t k = 0; k < n_sites; k++) { calls += S.h_nev[k] > 0; obs += S.h_nev[k]; }
    S.counts[0] = n_sites; S.counts[1] = calls; S.counts[2] = obs; S.counts[3] = 24 * n_sites + 4 * R;
    S.done = true;
    return DNB_OK;
}

int dnb_batch_analogue_result(dnb_batch *b, size_t i, dnb_analogue_result *o) {
    if (!b || !o || i >= b->R) return DNB_ERR_ARG;
    if (!b->s3.done) return DNB_ERR_STATE;
    const Stage3 &S = b->s3;
    memset(o, 0, sizeof(*o));
    o->status = b->h.status ? b->h.status[i] : DNB_READ_OK;
    const uint64_t lo = S.poi_off[i];
    o->n_sites = (uint32_t)(S.poi_off[i + 1] - lo);
    o->pos_on_ref = S.h_pos + lo; o->n_events = S.h_nev + lo;
    o->log_analogue = S.h_a + lo; o->log_thymidine = S.h_t + lo;
    return DNB_OK;
}

int dnb_batch_analogue_timings(dnb_batch *b, double ms[2], uint64_t counts[4]) {
    if (!b) return DNB_ERR_ARG;
    if (!b->s3.done) return DNB_ERR_STATE;
    for (int i = 0; i < 2; i++) if (ms) ms[i] = b->s3.ms[i];
    for (int i = 0; i < 4; i++) if (counts) counts[i] = b->s3.counts[i];
    return DNB_OK;
}

int dnb_submit_llr(dnb_ctx *ctx, const dnb_read_desc *reads, const dnb_read_extra *extra, size_t n_reads, uint32_t window,
                   dnb_batch **batch) {
    dnb_batch *b = nullptr;
    if (!ctx || !batch || (!extra && n_reads)) return DNB_ERR_ARG;
    uint64_t load = 0;
    ctx = pick_device(ctx, reads, n_reads, &load);
    int rc = upload(ctx, reads, n_reads, false, &b, /*gated=*/true);
    if (rc != DNB_OK) { ctx->inflight.fetch_sub(load, std::memory_order_relaxed); return rc; }
    b->load = load;
    {
        StageHold hold(ctx->gate_compute);
        rc = run(b);
        if (rc == DNB_OK) rc = dnb_batch_analogue_llr(b, extra, window);
    }
    if (rc == DNB_OK) {
        StageHold hold(ctx->gate_fetch);
        rc = fetch(b);
    }
    if (rc != DNB_OK) { free_batch(b); return rc; }
    drop_work(b);
    *batch = b;
    return DNB_OK;
}

int dnb_batch_feature_result(dnb_batch *b, size_t i, dnb_feature_result *o) {
    if (!b || !o || i >= b->R) return DNB_ERR_ARG;
    if (!b->s2.done) return DNB_ERR_STATE;
    const Stage2 &S = b->s2;
    memset(o, 0, sizeof(*o));
    o->status = S.h_status[i];
    const bool ok = o->status == DNB_READ_OK;
    const uint64_t lo = S.pos_off[i];
    o->n_pos = ok ? S.h_n_pos[i] : 0;
    o->signal = S.h_signal + lo * DNB_RAWDEPTH;
    o->core = S.h_core + lo; o->residual = S.h_resid + lo;
    o->coords = S.h_coords + lo; o->ref_index = S.h_ri + lo; o->query_index = S.h_qi + lo; o->quality = S.h_qual + lo;
    if (S.want_records) {
        o->n_recs = ok ? S.h_n_rec[i] : 0;
        o->recs = S.h_recs + S.rec_off[i];
    }
    return DNB_OK;
}

int dnb_batch_stage2_timings(dnb_batch *b, double ms[2], uint64_t bytes[2]) {
    if (!b) return DNB_ERR_ARG;
    if (!b->s2.done) return DNB_ERR_STATE;
    for (int i = 0; i < 2; i++) { if (ms) ms[i] = b->s2.ms[i]; if (bytes) bytes[i] = b->s2.bytes[i]; }
    return DNB_OK;
}

// ---- int16 ingest: Dorado slice (src/pod5.cpp:76-93) ----------------------------------------------------------------
int dnb_dorado_slice(uint64_t n_total, int64_t signal_length, int64_t signal_trim, int64_t signal_start_coord,
                     int is_split, uint64_t *first, uint64_t *count) {
    if (!first || !count) return DNB_ERR_ARG;
    *first = 0; *count = 0;
    if (n_total == 0) { g_last_error = "empty signal (the reference exits, pod5.cpp:64-73)"; return DNB_ERR_ARG; }
    if (signal_length <= 0) { *count = n_total; return DNB_OK; }              // not a Dorado BAM: the whole record
    // size_t arithmetic of pod5.cpp:82-83 / 90-91 (negative tags wrap there; here they are rejected)
    if (signal_trim < 0 || signal_start_coord < 0) { g_last_error = "negative Dorado signal tag"; return DNB_ERR_ARG; }
    const uint64_t start = (is_split ? (uint64_t)signal_start_coord : 0ull) + (uint64_t)signal_trim;
    const uint64_t end = (is_split ? (uint64_t)signal_start_coord : 0ull) + (uint64_t)signal_length;
    // erase(begin, begin + start) needs start <= size; erase(begin + (end - start), end()) needs start <= end <= size
    if (start > end || end > n_total) { g_last_error = "Dorado slice outside the record (undefined in the reference)"; return DNB_ERR_ARG; }
    *first = start; *count = end - start;
    return DNB_OK;
}

int dnb_eventalign_batch(dnb_ctx *ctx, const dnb_eventalign_desc *reads, size_t n_reads, uint32_t window,
                         dnb_eventalign_rec *recs, const uint64_t *rec_off, uint32_t *n_recs, int *status) {
    return eventalign_impl(ctx, reads, nullptr, n_reads, window, recs, rec_off, n_recs, status, nullptr, nullptr, nullptr);
}

int dnb_eventalign_features_batch(dnb_ctx *ctx, const dnb_eventalign_desc *reads, const dnb_feature_desc *feats,
                                  size_t n_reads, uint32_t window, dnb_eventalign_rec *recs, const uint64_t *rec_off,
                                  uint32_t *n_recs, int *status, const dnb_feature_tensors *out, const uint64_t *pos_off,
                                  uint32_t *n_pos) {
    if (!feats && n_reads) return DNB_ERR_ARG;
    return eventalign_impl(ctx, reads, feats, n_reads, window, recs, rec_off, n_recs, status, out, pos_off, n_pos);
}

}  // extern "C"
