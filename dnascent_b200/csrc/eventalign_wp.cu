// eventalign_wp.cu -- eventalign with the windows of a read computed in PARALLEL (experimental).
//
// STATUS: written at the end of round 1 with the last seconds of the round's GPU budget.  What has run on a B200:
// scripts/wp_check.py (records identical to the unmodified reference's for all ten golden reads, indel / soft-clip CIGARs
// included) and scripts/wp_perf.py (200 reads: 3.0 ms against 11.9 ms for the read-serial kernel, records identical).
// What has NOT: the rest of the GPU suite (edge cases, the tensor chain, the shim), long reads, a saturated batch, ncu.
// Until then it is reachable only with DNB_EA_WINDOW_PARALLEL=1; the default is the read-serial kernel of eventalign.cu.
//
// Same contract as eventalign.cu (src/alignment.cpp:547-744): records (event, ref_pos, indelScore, label) per read.
//
// Why.  The reference walks a read's windows serially: window w+1 starts one past the last match state of window w
// (alignment.cpp:740-741).  With one warp per read that chain makes a read's latency proportional to its length
// (34 ms per 10 kb) and the longest read bounds a batch.  But in ~99.7 % of the windows the path ends with a match in
// the last state on the last event, and then the next window depends on the reference alone: its length comes from the
// breakpoint rule at its start, its events are the aligned events whose k-mer lies in its query range (readHead is then
// only a lower bound of the first of them).  So the chain can be guessed without running a single Viterbi pass.
//
// Algorithm (rounds, all on the device; the host only reads two counters per round):
//   walk kernel     one warp per unfinished read follows the TRUE chain as far as window results exist.  A window is
//                   looked up by its reference index (win_at[], when readHead does not matter) or in the read's short
//                   list of readHead-dependent windows.  A missing window is created (its gather counts are computed on
//                   the spot), queued, and the walk continues behind it on the assumption that it advances fully.
//                   While everything was found the walk also copies the windows' records to the read's output.
//   window kernel   one warp per queued window: gather, builtinViterbi forward pass, traceback, records into the
//                   window's own block -- the code of eventalign.cu's loop body, with the loop turned inside out.
//   Round 0 queues every window of the guessed chain; later rounds queue the few the true chain needed instead (it
//   usually rejoins the guessed chain one window later).  The loop ends when every read was walked to its end.
//
// Requires the alignment's k-mer indices to be non-decreasing (true for normaliseEvents' output); a read that violates
// this is reported DNB_READ_UNDEFINED here (the read-serial kernel has no such restriction).
#include <algorithm>
#include <cstring>
#include <vector>
#include "eventalign_core.cuh"

namespace {

#define EA_NO_EVENT 0xffffffffu

struct EaWin {
    uint32_t read, ri, wl;
    uint32_t rh_in;                  // readHead the gather starts from
    uint32_t j0;                     // index of the first aligned event in range = readHead after the gather; EA_NO_EVENT if none
    uint32_t n_good;                 // observations: events in range that pass the 0 < mean < 250 guard
    uint32_t last_m_ref, last_m_ev;  // out (0 when no match state is on the path, as in the reference)
    uint32_t found;                  // out: a match state is on the path
    uint32_t state;                  // 0 = queued, 1 = done (windows with n_good < 2 are born done: skipped, :641)
    int32_t next_dep;                // next readHead-dependent window of the same read, -1 = none
    uint32_t pad;
    unsigned long long rec_base;     // the window's block in the scratch records, n_good slots
};

struct WpArgs {
    DnbEaArgs a;
    EaWin *wins;
    uint32_t win_cap;
    uint32_t *n_wins;
    int32_t *win_at;                 // indexed like ref: window starting at this reference index whose events do not depend on readHead
    int32_t *dep_head;               // [R]
    dnb_eventalign_rec *scratch;
    unsigned long long scratch_cap;
    unsigned long long *scratch_used;
    uint32_t *work;                  // windows queued in this round
    uint32_t work_cap;
    uint32_t *n_work;
    uint32_t *read_done;             // [R]
    uint32_t *n_incomplete;
    unsigned int *next_item;         // work counters of the two kernels
};

__device__ __forceinline__ EaRead load_read(const DnbEaArgs &a, uint32_t r) {
    EaRead rd;
    rd.ref = a.ref + a.ref_off[r];
    rd.rlen = (uint32_t)(a.ref_off[r + 1] - a.ref_off[r]);
    rd.r2q = a.r2q + a.ref_off[r];
    rd.pairs = a.pairs + a.al_off[r];
    rd.n_align = (uint32_t)(a.al_off[r + 1] - a.al_off[r]);
    rd.evm = a.ev_mean + a.ev_off[r];
    rd.shift = a.shift[r]; rd.scale = a.scale[r];
    rd.m12m1_int = a.trans[4 * (size_t)r + 0]; rd.m12m1_ext = a.trans[4 * (size_t)r + 1];
    rd.m12m1_ext_or_int = a.trans[4 * (size_t)r + 2]; rd.m12m1_ext_or_d = a.trans[4 * (size_t)r + 3];
    return rd;
}

// windowLength at reference index ri (alignment.cpp:556-604); false = the window is skipped (undefined base), advance by wl
__device__ __forceinline__ bool window_len(const DnbEaArgs &a, const EaRead &rd, unsigned ri, int lane, unsigned &wl) {
    const unsigned k = DNB_K, W = a.window;
    const unsigned bases_to_end = rd.rlen - ri;
    wl = min(bases_to_end, W);
    const char *ws = rd.ref + ri;
    if ((double)bases_to_end > 1.5 * (double)W) {
        const unsigned bl = (unsigned)(1.5 * (double)wl);
        if (!warp_defined(ws, bl, lane)) return false;
        const double lim = 1.5 * (double)wl - (double)k - 1.0;
        for (unsigned base = wl; (double)base < lim; base += 32) {
            const unsigned i = base + lane;
            bool hit = false;
            if ((double)i < lim) {
                const double m = a.model_mean[kmer_rank(ws + i)];
                const double mb = a.model_mean[kmer_rank(ws + i - 1)];
                const double mf = a.model_mean[kmer_rank(ws + i + 1)];
                hit = fabs(dSub(m, mf)) > 0.75 && fabs(dSub(m, mb)) > 0.75;
            }
            const unsigned hm = __ballot_sync(FULL, hit);
            if (hm) { wl = base + (unsigned)(__ffs(hm) - 1) + k; break; }
        }
    }
    return warp_defined(ws, wl, lane);
}

// The gather of alignment.cpp:611-632 from readHead rh: returns the number of observations and the updated readHead;
// with obs != nullptr also stores the scaled observations and their event indices (at most t_max of them).
__device__ __forceinline__ uint32_t gather(const DnbEaArgs &a, const EaRead &rd, uint32_t lo, uint32_t hi, uint32_t rh, int lane,
                                           uint32_t &rh_out, double *obs, uint32_t *obs_ev) {
    const unsigned lt = (1u << lane) - 1u;
    uint32_t ns = 0;
    bool first = true;
    long read_head = (long)rh;
    for (long j0 = (long)rh; j0 < (long)rd.n_align; j0 += 32) {
        const long j = j0 + lane;
        const bool have = j < (long)rd.n_align;
        const uint2 pr = have ? rd.pairs[j] : make_uint2(0u, 0u);
        const bool stop = have && pr.y >= hi;
        const unsigned stopmask = __ballot_sync(FULL, stop);
        const unsigned before = stopmask ? ((1u << (__ffs(stopmask) - 1)) - 1u) : FULL;
        const unsigned inmask = __ballot_sync(FULL, have && lo <= pr.y && pr.y < hi) & before;
        if (first && inmask) { read_head = j0 + (__ffs(inmask) - 1); first = false; }
        const bool in = (inmask >> lane) & 1u;
        const double em = in ? (double)rd.evm[pr.x] : 0.0;
        const bool good = in && 0. < em && em < 250.;                                       // :623
        const unsigned gm = __ballot_sync(FULL, good);
        const uint32_t pos = ns + __popc(gm & lt);
        if (obs && good && pos < a.t_max) { obs[pos] = dDiv(dSub(em, rd.shift), rd.scale); obs_ev[pos] = pr.x; }
        ns += __popc(gm);
        if (stopmask) break;
    }
    rh_out = first ? EA_NO_EVENT : (uint32_t)read_head;       // readHead moves only if an event was in range (:617-620)
    return ns;
}

// index of the first aligned event whose k-mer is >= lo (k-mer indices are non-decreasing)
__device__ __forceinline__ uint32_t first_at_least(const EaRead &rd, uint32_t lo) {
    uint32_t a = 0, b = rd.n_align;
    while (a < b) {
        const uint32_t m = (a + b) >> 1;
        if (rd.pairs[m].y < lo) a = m + 1; else b = m;
    }
    return a;
}

__device__ __forceinline__ bool kmers_sorted(const EaRead &rd, int lane) {
    bool ok = true;
    for (uint32_t j = lane + 1; j < rd.n_align; j += 32) ok = ok && rd.pairs[j - 1].y <= rd.pairs[j].y;
    return __all_sync(FULL, ok);
}

// ---- walk kernel: one warp per unfinished read -----------------------------------------------------------------------
__global__ void __launch_bounds__(128) wp_walk_kernel(WpArgs p) {
    const DnbEaArgs &a = p.a;
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(p.next_item, 1u);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        if (p.read_done[r]) continue;
        if (a.status[r] != DNB_READ_OK) { if (lane == 0) { a.n_rec[r] = 0; p.read_done[r] = 1; } continue; }
        const EaRead rd = load_read(a, r);
        if (!kmers_sorted(rd, lane)) { if (lane == 0) { a.status[r] = DNB_READ_UNDEFINED; a.n_rec[r] = 0; p.read_done[r] = 1; } continue; }
        dnb_eventalign_rec *recs = a.recs + a.rec_off[r];
        const uint64_t cap = a.rec_off[r + 1] - a.rec_off[r];
        const unsigned k = DNB_K;
        uint64_t nrec = 0;
        uint32_t rh = 0;
        unsigned ri = 0;
        bool ok = true;
        int status = DNB_READ_OK;
        while (ri < rd.rlen - k + 1) {
            unsigned wl;
            if (!window_len(a, rd, ri, lane, wl)) { ri += wl; continue; }
            const uint32_t lo = (uint32_t)rd.r2q[ri], hi = (uint32_t)rd.r2q[ri + wl - k + 1];
            const bool indep = rh <= first_at_least(rd, lo);
            int32_t w = -1;
            if (indep) w = p.win_at[a.ref_off[r] + ri];
            else
                for (int32_t d = p.dep_head[r]; d >= 0; d = p.wins[d].next_dep)
                    if (p.wins[d].ri == ri && p.wins[d].rh_in == rh) { w = d; break; }
            if (w < 0) {
                // not computed yet: create it, queue it, and go on behind it as if it advanced fully
                ok = false;
                uint32_t j0;
                const uint32_t ns = gather(a, rd, lo, hi, rh, lane, j0, nullptr, nullptr);
                uint32_t idx = 0;
                unsigned long long base = 0;
                if (lane == 0) {
                    idx = atomicAdd(p.n_wins, 1u);
                    if (ns >= 2) base = atomicAdd(p.scratch_used, (unsigned long long)ns);
                }
                idx = __shfl_sync(FULL, idx, 0);
                base = __shfl_sync(FULL, base, 0);
                if (idx >= p.win_cap || base + ns > p.scratch_cap || ns > a.t_max) { status = DNB_READ_OVERFLOW; break; }
                if (lane == 0) {
                    EaWin nw;
                    nw.read = r; nw.ri = ri; nw.wl = wl; nw.rh_in = rh; nw.j0 = j0; nw.n_good = ns;
                    nw.last_m_ref = 0; nw.last_m_ev = 0; nw.found = 0; nw.state = ns < 2 ? 1u : 0u;
                    nw.next_dep = -1; nw.pad = 0; nw.rec_base = base;
                    if (!indep) { nw.next_dep = p.dep_head[r]; }
                    p.wins[idx] = nw;
                    __threadfence();
                    if (indep) p.win_at[a.ref_off[r] + ri] = (int32_t)idx; else p.dep_head[r] = (int32_t)idx;
                    if (ns >= 2) {
                        const uint32_t slot = atomicAdd(p.n_work, 1u);
                        if (slot < p.work_cap) p.work[slot] = idx;
                    }
                }
                __syncwarp();
                if (j0 != EA_NO_EVENT) rh = j0;                                             // the gather moves readHead even if the window is skipped
                if (ns < 2) { ri += wl; continue; }                                         // :641
                rh = j0 + ns;                                                               // guess: every event consumed ...
                ri += wl - k + 1;                                                           // ... and the last state matched
                continue;
            }
            const EaWin W = p.wins[w];
            if (W.j0 != EA_NO_EVENT) rh = W.j0;
            if (W.n_good < 2) { ri += wl; continue; }                                       // :641
            if (W.found) {
                const uint64_t cnt = (uint64_t)W.last_m_ev + 1;
                if (ok) {
                    if (nrec + cnt > cap) { status = DNB_READ_OVERFLOW; break; }
                    const dnb_eventalign_rec *src = p.scratch + W.rec_base;
                    for (uint64_t t = lane; t < cnt; t += 32) recs[nrec + t] = src[t];
                }
                nrec += cnt;
            }
            rh = W.j0 + W.last_m_ev + 1;                                                    // :740-741
            ri += W.last_m_ref + 1;
        }
        __syncwarp();
        if (status != DNB_READ_OK) {
            if (lane == 0) { a.status[r] = status; a.n_rec[r] = 0; p.read_done[r] = 1; }
        } else if (ok) {
            if (lane == 0) { a.n_rec[r] = (uint32_t)nrec; p.read_done[r] = 1; }
        } else if (lane == 0) {
            atomicAdd(p.n_incomplete, 1u);
        }
    }
}

// ---- window kernel: one warp per queued window (the loop body of eventalign.cu's eventalign_read) ----------------------
__global__ void __launch_bounds__(EA_WARPS * 32) wp_window_kernel(WpArgs p, uint32_t n_work) {
    const DnbEaArgs &a = p.a;
    const int lane = threadIdx.x & 31;
    const uint32_t wslot = blockIdx.x * EA_WARPS + (threadIdx.x >> 5);
    double *obs = a.scratch_obs + (size_t)wslot * a.t_max;
    uint32_t *obs_ev = a.scratch_ev + (size_t)wslot * a.t_max;
    uint8_t *bt = a.scratch_bt + (size_t)wslot * a.t_max * (EA_SLOTS * 32);
    const unsigned k = DNB_K;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(p.next_item, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_work) break;
        const uint32_t w = p.work[item];
        const EaWin W = p.wins[w];
        const EaRead rd = load_read(a, W.read);
        const unsigned ri = W.ri, wl = W.wl;
        const char *ws = rd.ref + ri;
        const uint32_t lo = (uint32_t)rd.r2q[ri], hi = (uint32_t)rd.r2q[ri + wl - k + 1];
        uint32_t j0;
        const uint32_t ns = gather(a, rd, lo, hi, W.rh_in, lane, j0, obs, obs_ev);          // == W.n_good (2 <= ns <= t_max)
        const int indel = ((int)hi - (int)lo) - (int)(wl - k + 1);                          // :635-638
        __syncwarp();

        // ---- builtinViterbi forward pass (alignment.cpp:193-428) ----
        const int n = (int)(wl - k + 1);
        double mu[EA_SLOTS], I[EA_SLOTS], M[EA_SLOTS], D[EA_SLOTS];
#pragma unroll
        for (int s = 0; s < EA_SLOTS; s++) {
            const int i = 32 * s + lane;
            mu[s] = i < n ? a.model_mean[kmer_rank(ws + i)] : 0.0;
            I[s] = NEG_INF; M[s] = NEG_INF; D[s] = NEG_INF;
        }
        {   // :241-250  D_prev[0] = 0 + m2d, D_prev[i] = D_prev[i-1] + d2d (repeated addition, not a multiply)
            double v = a.m2d;
            for (int i = 0; i < n; i++) {
                if ((i & 31) == lane) {
                    if ((i >> 5) == 0) D[0] = v;
                    if ((i >> 5) == 1) D[1] = v;
                    if ((i >> 5) == 2) D[2] = v;
                }
                v = dAdd(v, a.d2d);
            }
        }
        const int nslots = (n + 31) >> 5;
        if (nslots == 1) viterbi_forward<1>(a, rd, obs, ns, n, mu, I, M, D, bt, lane);
        else if (nslots == 2) viterbi_forward<2>(a, rd, obs, ns, n, mu, I, M, D, bt, lane);
        else viterbi_forward<3>(a, rd, obs, ns, n, mu, I, M, D, bt, lane);
        // ---- termination (:447-474) ----
        int end_ty;
        {
            const int sl = (n - 1) >> 5, ll = (n - 1) & 31;
            const double dl = sl == 0 ? D[0] : sl == 1 ? D[1] : D[2];
            const double ml = sl == 0 ? M[0] : sl == 1 ? M[1] : M[2];
            const double il = sl == 0 ? I[0] : sl == 1 ? I[1] : I[2];
            const double e1 = dAdd(ml, rd.m12m1_ext_or_d), e2 = dAdd(il, a.i2m);
            double m = dl; int ty = 0;                    // 0 = D, 1 = M, 2 = I
            if (e1 > m) { m = e1; ty = 1; }
            if (e2 > m) { m = e2; ty = 2; }
            end_ty = __shfl_sync(FULL, ty, ll);
        }
        __syncwarp();
        // ---- traceback (:476-505) and the two passes over the state labels (:655-736), by lane 0: the record of
        // observation t (kept when it is not after the last match) goes to slot t of the window's block
        if (lane == 0) {
            dnb_eventalign_rec *recs = p.scratch + W.rec_base;
            uint32_t last_m_ev = 0, last_m_ref = 0;
            int found = 0;
            int ty = end_ty, i = n - 1;
            long t = (long)ns;
            for (;;) {
                if (ty != 0) {
                    if (t == 0) { ty = 0; i = 0; continue; }      // never-written column 0 of an M/I row reads as "D 0, t 0"
                    const uint32_t ev = (uint32_t)(t - 1);
                    if (ty == 1 && !found) { found = 1; last_m_ev = ev; last_m_ref = (uint32_t)i; }
                    if (found) {
                        dnb_eventalign_rec rc;
                        rc.event = obs_ev[ev]; rc.ref_pos = ri + (uint32_t)i; rc.indel_score = indel; rc.label = (uint32_t)ty;
                        recs[ev] = rc;
                    }
                    const uint32_t cd = bt[(size_t)ev * (EA_SLOTS * 32) + i];
                    t--;
                    if (ty == 1) {
                        const uint32_t c = (cd >> 2) & 3u;
                        if (i == 0) { if (c == 1) break; /* c == 0: M(0) again */ }
                        else if (c == 0) { ty = 2; i--; } else if (c == 1) { i--; } else if (c == 3) { ty = 0; i--; }
                    } else {
                        const uint32_t c = cd & 3u;
                        if (c == 1) ty = 1; else if (c == 2) break;
                    }
                } else {
                    if (i == 0) break;                             // :242-243, 326
                    if (t == 0) { i--; continue; }                 // :248-250
                    const uint32_t c = (bt[(size_t)(t - 1) * (EA_SLOTS * 32) + i] >> 4) & 1u;
                    ty = c ? 0 : 1;
                    i--;
                }
            }
            p.wins[w].j0 = j0;
            p.wins[w].last_m_ref = last_m_ref;
            p.wins[w].last_m_ev = last_m_ev;
            p.wins[w].found = (uint32_t)found;
            __threadfence();
            p.wins[w].state = 1u;
        }
        __syncwarp();
    }
}

// reads the round limit left unfinished (never seen on the CPU prototype: two rounds suffice there)
__global__ void wp_fail_unfinished_kernel(WpArgs p) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p.a.n_reads && !p.read_done[r]) { p.a.status[r] = DNB_READ_OVERFLOW; p.a.n_rec[r] = 0; }
}

}  // namespace

// Runs the whole thing on stream s (synchronises it once per round to read the two counters).  `a` is the argument block
// the read-serial launch would get (its scratch_* fields are ignored: the workspace here is sized for this kernel's grid).
cudaError_t dnb_run_eventalign_wp(const DnbEaArgs &a_in, uint64_t tot_ref, uint64_t tot_align, int device, cudaStream_t s,
                                  const DnbAlloc &alloc, cudaEvent_t sync_ev) {
    if (a_in.n_reads == 0) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    // the workspace comes from the caller's allocator (the batch's device cache or the context's pool) and stays the
    // caller's to free
    auto dalloc = [&](size_t bytes) -> void * {
        void *q = nullptr;
        if (err == cudaSuccess) { q = alloc.fn(alloc.user, bytes ? bytes : 16); if (!q) err = cudaErrorMemoryAllocation; }
        return q;
    };
    WpArgs p;
    memset(&p, 0, sizeof(p));
    p.a = a_in;
    const uint64_t R = a_in.n_reads;
    // windows advance by >= window - 8 bases on the guessed chain; repairs add a few per cent
    const uint64_t step = a_in.window > 16 ? a_in.window - 8 : 8;
    p.win_cap = (uint32_t)std::min<uint64_t>((tot_ref / step + 2 * R) * 3 / 2 + 1024, 0xfffffff0ull);
    p.work_cap = p.win_cap;
    p.scratch_cap = tot_align * 3 / 2 + 4096;
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wp_window_kernel, EA_WARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    const unsigned grid_win = (unsigned)(sms * per_sm);
    const size_t warps = (size_t)grid_win * EA_WARPS;
    p.wins = (EaWin *)dalloc((size_t)p.win_cap * sizeof(EaWin));
    p.win_at = (int32_t *)dalloc((tot_ref ? tot_ref : 1) * 4);
    p.dep_head = (int32_t *)dalloc(R * 4);
    p.scratch = (dnb_eventalign_rec *)dalloc(p.scratch_cap * sizeof(dnb_eventalign_rec));
    p.work = (uint32_t *)dalloc((size_t)p.work_cap * 4);
    p.read_done = (uint32_t *)dalloc(R * 4);
    uint32_t *counters = (uint32_t *)dalloc(64);          // [0] n_wins [1] n_work [2] n_incomplete [3] next_item; [4..5] scratch_used
    p.a.scratch_obs = (double *)dalloc(warps * a_in.t_max * 8);
    p.a.scratch_ev = (uint32_t *)dalloc(warps * a_in.t_max * 4);
    p.a.scratch_bt = (uint8_t *)dalloc(warps * a_in.t_max * (EA_SLOTS * 32));
    if (err == cudaSuccess) {
        p.n_wins = counters; p.n_work = counters + 1; p.n_incomplete = counters + 2; p.next_item = counters + 3;
        p.scratch_used = (unsigned long long *)(counters + 4);
        auto ck = [&](cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; };
        ck(cudaMemsetAsync(counters, 0, 64, s));
        ck(cudaMemsetAsync(p.win_at, 0xFF, (tot_ref ? tot_ref : 1) * 4, s));
        ck(cudaMemsetAsync(p.dep_head, 0xFF, R * 4, s));
        ck(cudaMemsetAsync(p.read_done, 0, R * 4, s));
        const unsigned grid_walk = (unsigned)std::min<uint64_t>((R + 3) / 4, (uint64_t)sms * 8);
        uint32_t h[3] = {0, 0, 0};
        for (int round = 0; round < 64 && err == cudaSuccess; round++) {
            ck(cudaMemsetAsync(counters + 1, 0, 12, s));                    // n_work, n_incomplete, next_item
            wp_walk_kernel<<<grid_walk, 128, 0, s>>>(p);
            ck(cudaGetLastError());
            ck(cudaMemcpyAsync(h, counters, 12, cudaMemcpyDeviceToHost, s));
            if (sync_ev) { ck(cudaEventRecord(sync_ev, s)); ck(cudaEventSynchronize(sync_ev)); }   // a sleeping wait
            else ck(cudaStreamSynchronize(s));
            if (err != cudaSuccess) break;
            const uint32_t n_work = h[1] < p.work_cap ? h[1] : p.work_cap;
            if (n_work) {
                ck(cudaMemsetAsync(counters + 3, 0, 4, s));
                wp_window_kernel<<<grid_win, EA_WARPS * 32, 0, s>>>(p, n_work);
                ck(cudaGetLastError());
            }
            if (h[2] == 0) break;                                           // every read was walked to its end
        }
        if (err == cudaSuccess && h[2] != 0) {
            wp_fail_unfinished_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(p);
            ck(cudaGetLastError());
        }
    }
    return err;
}
