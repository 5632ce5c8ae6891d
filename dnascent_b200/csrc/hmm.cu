// hmm.cu -- analogue log-likelihood: the forward algorithm of sequenceProbability, one WARP per site, analogue and
// thymidine passes side by side, and the per-read gathering of llAcrossRead done on the device.
//
// Replaces (reference, paths relative to /root/reference):
//   sequenceProbability      src/detect.cpp:235-378   (called twice per T site by llAcrossRead, :546-547)
//   getPOIs + llAcrossRead   src/detect.cpp:381-390, 393-574  (which sites, which events)
//   eexp/eln/lnSum/lnProd    src/probability.cpp:23-88
//   normalPDF                src/probability.cpp:145-148
//
// Forward pass.  Lanes are the 2*window positions (24 for window 12), the {I, M, D} values of a position live in the
// lane's registers, observations are consumed one after the other.  The reference keeps "log 0" as NaN with
// NaN-skipping lnSum / NaN-absorbing lnProd; here log 0 is -inf, for which IEEE arithmetic does the same thing
// (x + -inf = -inf, exp(-inf) = 0), and the result is turned back into NaN at the end.  The deletion chain
//     D[i] = lnSum(M[i-1] + m2d, D[i-1] + d2d)
// runs along the position axis inside a time step; with u[i] = D[i] - i*d2d it is a plain log-sum-exp prefix
// u[i] = lse(u[i-1], M[i-1] + m2d - i*d2d), done as a warp scan on (max, scaled sum) pairs: 5 exp + 1 log per lane
// instead of 23 dependent lnSums.  The two passes differ only in the emission of the positions inside
// [BrdUStart, BrdUEnd] whose 9-mer contains a T (:317-323).  Tolerance of this path is 1e-4 relative
// (BASELINE.json), so CUDA's double exp/log are used and sums are re-associated.
//
// Gathering (resident form).  llAcrossRead walks the T positions of the reference (descending for reverse reads) and,
// for each, scans r.eventAlignment from a moving `readHead` for the events whose k-mer lies in
// [refToQuery[p-w], refToQuery[p+w]).  k-mer indices are non-decreasing along the alignment, so the events of a site
// are an index RANGE found by two binary searches, and readHead is a one-variable recurrence over the sites that one
// thread per read evaluates (it matters only where refToQuery has gaps: operator[] reads a missing key as 0,
// :446, :461, :488).  Sites, ranges and forward passes never leave the device; only (position, two log-probabilities)
// per site come back.
#include <cmath>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define HMM_MAXW 16   // window <= 16 -> <= 32 positions (the reference uses windowLength = 12)
#define HMM_FULL 0xffffffffu
#define HMM_WARPS 4

namespace {

__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000ll); }
__device__ __forceinline__ double d_ninf() { return __longlong_as_double(0xfff0000000000000ll); }

// The double-precision exp / log / log1p of CUDA's math library inline ~60 instructions per call site; a forward step has
// 26 of them per pass, and with everything inlined the loop body overflowed the instruction cache (ncu: "no
// instruction" was the second largest stall).  One out-of-line copy of each keeps the loop body small.
__device__ __noinline__ double exp_nl(double x) { return exp(x); }
__device__ __noinline__ double log_nl(double x) { return log(x); }
__device__ __noinline__ double log1p_nl(double x) { return log1p(x); }

__device__ __forceinline__ double lse2(double a, double b) {
    const double m = fmax(a, b);
    if (m == d_ninf()) return m;
    return m + log1p_nl(exp_nl(-fabs(a - b)));       // one of them -inf: exp(-inf) = 0
}
__device__ __forceinline__ double lse3(double a, double b, double c) {
    const double m = fmax(fmax(a, b), c);
    if (m == d_ninf()) return m;
    return m + log_nl(exp_nl(a - m) + exp_nl(b - m) + exp_nl(c - m));
}
__device__ __forceinline__ double lse4(double a, double b, double c, double d) {
    const double m = fmax(fmax(a, b), fmax(c, d));
    if (m == d_ninf()) return m;
    return m + log_nl(exp_nl(a - m) + exp_nl(b - m) + exp_nl(c - m) + exp_nl(d - m));
}

// per-lane emission model of one pass: eln(normalPDF(mu, sigma, x))
struct Emit {
    double mu, inv2s2, lnc, c;
    __device__ __forceinline__ void set(double m, double s) {
        mu = m;
        const double s2 = s * s;
        c = 1.0 / sqrt(2.0 * s2 * 3.14159265358979323846);
        lnc = log(c);
        inv2s2 = 1.0 / (2.0 * s2);
    }
    __device__ __forceinline__ double at(double x) const {
        const double dx = x - mu;
        const double y = -(dx * dx) * inv2s2;
        if (y > -700.0) return lnc + y;
        const double p = c * exp_nl(y);        // the reference's linear-space value underflows to 0 -> log 0
        return p > 0.0 ? log_nl(p) : d_ninf();
    }
};

struct Trans {
    double D2D, D2M, I2M, M2D, M2I, I2I, M2M_int, M2M_ext, l25, l5;
};
__device__ __forceinline__ Trans make_trans(double epb) {
    Trans t;
    // HMM_TransitionProbs_DNA_R10 {0.3, 0.7, 0.999, 0.0025, 0.001, 0.001}, src/config.h:42
    t.D2D = log(0.3); t.D2M = log(0.7); t.I2M = log(0.999); t.M2D = log(0.0025); t.M2I = log(0.001); t.I2I = log(0.001);
    const double a = 1. - (1. / epb);
    t.M2M_int = a > 0.0 ? log(a) : d_ninf();                 // eln(0) is log 0; the reference throws for a < 0
    const double e = 1.0 - t.M2D - t.M2I - t.M2M_int;        // quirk Q10: log-space values inside, on purpose (:254-255)
    t.M2M_ext = e > 0.0 ? log(e) : d_ninf();
    t.l25 = log(0.25); t.l5 = log(0.5);
    return t;
}

// state of one pass: this lane's position
struct PassState {
    double I, M, D, firstI;
    __device__ __forceinline__ void init(int lane, const Trans &t) {
        I = d_ninf(); M = d_ninf(); firstI = d_ninf();
        D = t.l25 + (double)lane * t.D2D;                    // detect.cpp:265-271
    }
};

// one observation of one pass (detect.cpp:277-356); start_prev is 0 for the first observation, log 0 afterwards
__device__ __forceinline__ void hmm_step(PassState &p, const Emit &em, double x, double start_prev, const Trans &t,
                                         int lane, int n) {
    const double match = em.at(x);
    const double Iup = __shfl_up_sync(HMM_FULL, p.I, 1), Mup = __shfl_up_sync(HMM_FULL, p.M, 1), Dup = __shfl_up_sync(HMM_FULL, p.D, 1);
    const double firstI_cur = lse2(start_prev + t.l25, p.firstI + t.l25);
    const double Ic = lse2(p.I + t.I2I, p.M + t.M2I);
    double Mc;
    if (lane == 0) Mc = match + lse3(p.firstI + t.l5, p.M + t.M2M_int, start_prev + t.l5);
    else Mc = match + lse4(Iup + t.I2M, Mup + t.M2M_ext, p.M + t.M2M_int, Dup + t.D2M);
    // deletion chain: inclusive log-sum-exp scan of v along the lanes, on (max, scaled sum) pairs
    const double Mc_up = __shfl_up_sync(HMM_FULL, Mc, 1);
    const double shiftD = (double)lane * t.D2D;
    double m = lane == 0 ? firstI_cur + t.l25 : (Mc_up + t.M2D) - shiftD;
    if (lane >= n) m = d_ninf();
    double s = 1.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double m2 = __shfl_up_sync(HMM_FULL, m, d), s2 = __shfl_up_sync(HMM_FULL, s, d);
        if (lane >= d && m2 != d_ninf()) {
            if (m == d_ninf()) { m = m2; s = s2; }
            else if (m >= m2) s = s + s2 * exp_nl(m2 - m);
            else { s = s * exp_nl(m - m2) + s2; m = m2; }
        }
    }
    p.D = (m == d_ninf()) ? m : m + log_nl(s) + shiftD;
    p.I = Ic; p.M = Mc; p.firstI = firstI_cur;
}

__device__ __forceinline__ double hmm_finish(const PassState &p, const Trans &t, int n) {
    // termination, read from the last position's lane (:361-367)
    const double v = lse3(p.D, p.M + lse2(t.M2M_ext, t.M2D), p.I + t.I2M);
    const double out = __shfl_sync(HMM_FULL, v, n - 1);
    return out == d_ninf() ? d_nan() : out;
}

// the per-lane emission models of a (2*window + 9)-base snippet (lane = position); false: a base outside ACGT
__device__ __forceinline__ bool site_models(const char *sn, uint32_t window, int lane, int n, const DnbModelDev &unl,
                                            const DnbModelDev &ana, Emit &et, Emit &ea) {
    uint32_t rk = 0;
    bool hasT = false, defined = true;
    if (lane < n) {
        for (int j = 0; j < DNB_K; j++) {
            const char c = sn[lane + j];
            rk = rk * 4u + dnb_base_code(c);
            hasT |= c == 'T';
            defined &= (c == 'A' || c == 'T' || c == 'G' || c == 'C');
        }
    }
    const uint32_t a_start = window - DNB_K / 2, a_end = window + DNB_K / 2;   // detect.cpp:544-545
    double mu_t = 0.0, sg_t = 1.0, mu_a = 0.0, sg_a = 1.0;
    if (lane < n) {
        mu_t = unl.mean[rk]; sg_t = unl.stdv[rk];
        const bool use_a = lane >= 1 && a_start <= (uint32_t)lane && (uint32_t)lane <= a_end && hasT;   // position 0 always unlabelled (:281)
        mu_a = use_a ? ana.mean[rk] : mu_t;
        sg_a = use_a ? ana.stdv[rk] : sg_t;
    }
    et.set(mu_t, sg_t);
    ea.set(mu_a, sg_a);
    return __all_sync(HMM_FULL, defined);
}

// ---- flat form: observations given by the caller (dnb_sequence_probability_batch) ------------------------------------
__global__ void __launch_bounds__(HMM_WARPS * 32) hmm_sites_kernel(const double *obs, const uint64_t *obs_off, const char *seq,
                                                                  const double *shift, const double *scale, const double *epb,
                                                                  size_t n_sites, uint32_t window, DnbModelDev unl,
                                                                  DnbModelDev ana, double *out_a, double *out_t) {
    const int lane = threadIdx.x & 31;
    const size_t s = (size_t)blockIdx.x * HMM_WARPS + (threadIdx.x >> 5);
    if (s >= n_sites) return;
    const int n = 2 * (int)window;
    Emit et, ea;
    site_models(seq + s * (size_t)(2 * window + DNB_K), window, lane, n, unl, ana, et, ea);
    const Trans t = make_trans(epb[s]);
    PassState pa, pt;
    pa.init(lane, t); pt.init(lane, t);
    const double *o = obs + obs_off[s];
    const uint32_t n_obs = (uint32_t)(obs_off[s + 1] - obs_off[s]);
    const double sh = shift[s], sc = scale[s];
    double start_prev = 0.0;
    for (uint32_t t0 = 0; t0 < n_obs; t0 += 32) {
        const double mine = (t0 + lane < n_obs) ? (o[t0 + lane] - sh) / sc : 0.0;
        const int cnt = (int)min(32u, n_obs - t0);
        for (int k = 0; k < cnt; k++) {
            const double x = __shfl_sync(HMM_FULL, mine, k);
            hmm_step(pa, ea, x, start_prev, t, lane, n);
            hmm_step(pt, et, x, start_prev, t, lane, n);
            start_prev = d_ninf();                              // start_curr is never set (:249, :356)
        }
    }
    const double ra = hmm_finish(pa, t, n), rt = hmm_finish(pt, t, n);
    if (lane == 0) {
        out_a[s] = n_obs ? ra : d_nan();
        out_t[s] = n_obs ? rt : d_nan();
    }
}

// ---- resident form, step 1: the T positions of every read and their alignment index ranges --------------------------
// forward alignment pair j of read r = rev[n - 1 - j]; its k-mer index is non-decreasing in j
__device__ __forceinline__ uint32_t kmer_at(const uint2 *rev, uint32_t n, uint32_t j) { return rev[n - 1 - j].y; }
__device__ __forceinline__ uint32_t lower_bound_kmer(const uint2 *rev, uint32_t n, uint32_t key) {   // first j with k_j >= key
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (kmer_at(rev, n, mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) llr_sites_kernel(DnbBatchView v, DnbLlrArgs a) {
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t base;
    const uint32_t r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t w = a.window;
    const uint32_t L = (uint32_t)(v.r_off[r + 1] - v.r_off[r]);
    const char *ref = v.ref + v.r_off[r];
    const int32_t *r2q = a.r2q + v.r_off[r];
    uint32_t *poi = a.poi + v.r_off[r];
    int32_t *jb = a.j_begin + v.r_off[r], *je = a.j_end + v.r_off[r];
    if (tid == 0) base = 0;
    __syncthreads();
    const bool live = v.status[r] == DNB_READ_OK && a.n_align[r] > 0 && L >= 4 * w + 1;
    // ---- the T's of [2w, L - 2w) in ascending order (getPOIs, :381-390) ----
    if (live) {
        for (uint32_t p0 = 2 * w; p0 < L - 2 * w; p0 += 256) {
            const uint32_t p = p0 + tid;
            const bool isT = p < L - 2 * w && ref[p] == 'T';
            const unsigned bal = __ballot_sync(HMM_FULL, isT);
            if (lane == 0) warp_cnt[wid] = __popc(bal);
            __syncthreads();
            uint32_t before = base;
            for (int k = 0; k < wid; k++) before += warp_cnt[k];
            if (isT) poi[before + __popc(bal & ((1u << lane) - 1u))] = p;
            __syncthreads();
            if (tid == 0) { uint32_t tot = 0; for (int k = 0; k < 8; k++) tot += warp_cnt[k]; base += tot; }
            __syncthreads();
        }
    }
    const uint32_t np = base;
    if (tid == 0) a.n_poi[r] = np;
    if (np == 0) return;
    // ---- per site: is the snippet A/C/G/T only (:423-442), and the two lower bounds of its k-mer window ----
    const uint2 *rev = reinterpret_cast<const uint2 *>(a.al_pairs_rev) + a.al_off[r];
    const uint32_t na = a.n_align[r];
    for (uint32_t i = tid; i < np; i += 256) {
        const uint32_t p = poi[i];
        bool defined = (uint64_t)p + w + DNB_K <= L;
        for (uint32_t c = p - w; defined && c < p + w + DNB_K; c++) {
            const char ch = ref[c];
            defined = ch == 'A' || ch == 'T' || ch == 'G' || ch == 'C';
        }
        if (!defined) { jb[i] = -2; je[i] = -2; continue; }          // skipped before the event loop: readHead untouched
        const uint32_t lo = (uint32_t)r2q[p - w], hi = (uint32_t)r2q[p + w];
        jb[i] = (int32_t)lower_bound_kmer(rev, na, lo);              // LB(lo)
        je[i] = (int32_t)lower_bound_kmer(rev, na, hi);              // LB(hi)
    }
    __syncthreads();
    // ---- readHead in visit order (ascending sites for forward reads, descending for reverse reads) ----
    // readHead is a one-variable recurrence over the sites.  parseCigar makes refToQuery non-decreasing, and then the
    // lower bounds are monotone along the visit order and readHead never binds: every site's range is [LB(lo), LB(hi))
    // and all of them are written in parallel.  A read whose window bounds are not monotone (refToQuery with gaps
    // that read as 0) takes the literal sequential walk on one thread.
    const bool rev_strand = a.is_reverse[r] != 0;
    bool mono = true;
    for (uint32_t i = tid; i < np; i += 256) {
        if (jb[i] == -2) continue;
        // previous defined site in ascending order
        uint32_t k = i;
        while (k > 0 && jb[k - 1] == -2) k--;
        if (k > 0) mono = mono && jb[k - 1] <= jb[i] && je[k - 1] <= je[i];
    }
    mono = __syncthreads_and(mono);
    if (mono) {
        for (uint32_t i = tid; i < np; i += 256) {
            const int32_t lbl = jb[i], lbh = je[i];
            if (lbl == -2) { jb[i] = 0; je[i] = 0; continue; }
            if (!rev_strand) { jb[i] = lbl; je[i] = lbh > lbl ? lbh : lbl; }
            else if (lbh - 1 >= lbl) { jb[i] = lbl; je[i] = lbl == 0 ? -lbh - 1 : lbh; }   // bottom == 0: no `break`, order stays descending
            else { jb[i] = 0; je[i] = 0; }
        }
        return;
    }
    if (tid == 0) {
        if (!rev_strand) {
            int32_t h = 0;
            for (uint32_t i = 0; i < np; i++) {
                const int32_t lbl = jb[i], lbh = je[i];
                if (lbl == -2) { jb[i] = 0; je[i] = 0; continue; }
                const int32_t b0 = max(h, lbl), e0 = max(h, lbh);   // in-window indices [b0, e0)
                jb[i] = b0; je[i] = e0 > b0 ? e0 : b0;
                if (e0 > b0) h = b0;
            }
        } else {
            int32_t h = (int32_t)na - 1;
            for (uint32_t ii = np; ii-- > 0;) {
                const int32_t lbl = jb[ii], lbh = je[ii];
                if (lbl == -2) { jb[ii] = 0; je[ii] = 0; continue; }
                const int32_t top = min(h, lbh - 1), bottom = lbl;     // in-window indices [bottom, top]
                if (top >= bottom) {
                    // pushed from top down; reversed into ascending order only if the loop met a k-mer below the
                    // window (the `break` at :470-474), i.e. if there is an alignment entry before `bottom`
                    jb[ii] = bottom; je[ii] = top + 1;
                    if (bottom == 0) je[ii] = -(top + 1) - 1;          // no break: order stays descending (encoded < 0)
                    h = top;
                } else { jb[ii] = 0; je[ii] = 0; }
            }
        }
    }
}

// ---- resident form, step 2: one warp per site pulls work from a counter -----------------------------------------------
__global__ void __launch_bounds__(HMM_WARPS * 32) llr_forward_kernel(DnbBatchView v, DnbLlrArgs a, uint64_t n_sites_total,
                                                                    DnbModelDev unl, DnbModelDev ana) {
    const int lane = threadIdx.x & 31;
    const int n = 2 * (int)a.window;
    const uint32_t min_events = 2 * a.window - DNB_K;                  // :515
    for (;;) {
        unsigned long long s = 0;
        if (lane == 0) s = atomicAdd(a.next_site, 1ull);
        s = __shfl_sync(HMM_FULL, s, 0);
        if (s >= n_sites_total) return;
        // site -> (read, index): poi_off is the exclusive prefix of n_poi
        uint32_t lo = 0, hi = v.n_reads;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.poi_off[mid] <= s) lo = mid; else hi = mid; }
        const uint32_t r = lo;
        const uint32_t i = (uint32_t)(s - a.poi_off[r]);
        const uint32_t p = a.poi[v.r_off[r] + i];
        const int32_t b0 = a.j_begin[v.r_off[r] + i];
        int32_t e0 = a.j_end[v.r_off[r] + i];
        const bool descending = e0 < 0;
        if (descending) e0 = -(e0 + 1);
        if (lane == 0) { a.out_pos[s] = p; a.out_n_events[s] = 0; a.out_a[s] = d_nan(); a.out_t[s] = d_nan(); }
        if (e0 <= b0) continue;
        const uint2 *rev = reinterpret_cast<const uint2 *>(a.al_pairs_rev) + a.al_off[r];
        const uint32_t na = a.n_align[r];
        const float *evm = v.ev_mean + v.ev_off[r];
        // events of the window with 0 < mean < 250 (:456-459); too few -> no call (:515)
        uint32_t n_ev = 0;
        for (int32_t j0 = b0; j0 < e0; j0 += 32) {
            const int32_t j = j0 + lane;
            bool ok = false;
            if (j < e0) { const float mv = evm[rev[na - 1 - (uint32_t)j].x]; ok = (double)mv > 0. && (double)mv < 250.0; }
            n_ev += __popc(__ballot_sync(HMM_FULL, ok));
        }
        if (n_ev < min_events) continue;
        Emit et, ea;
        site_models(v.ref + v.r_off[r] + p - a.window, a.window, lane, n, unl, ana, et, ea);
        const double epb = (double)v.et_n[r] / (double)((int64_t)(v.q_off[r + 1] - v.q_off[r]) - DNB_K);   // event_handling.cpp:606
        const Trans t = make_trans(epb);
        PassState pa, pt;
        pa.init(lane, t); pt.init(lane, t);
        const double sh = a.shift[r], sc = a.scale[r];
        double start_prev = 0.0;
        const int32_t span = e0 - b0;
        for (int32_t q0 = 0; q0 < span; q0 += 32) {
            const int32_t q = q0 + lane;
            double mine = 0.0;
            bool ok = false;
            if (q < span) {
                const int32_t j = descending ? e0 - 1 - q : b0 + q;
                const float mv = evm[rev[na - 1 - (uint32_t)j].x];
                ok = (double)mv > 0. && (double)mv < 250.0;
                mine = ((double)mv - sh) / sc;
            }
            unsigned todo = __ballot_sync(HMM_FULL, ok);
            while (todo) {
                const int k = __ffs(todo) - 1;
                todo &= todo - 1;
                const double x = __shfl_sync(HMM_FULL, mine, k);
                hmm_step(pa, ea, x, start_prev, t, lane, n);
                hmm_step(pt, et, x, start_prev, t, lane, n);
                start_prev = d_ninf();
            }
        }
        const double ra = hmm_finish(pa, t, n), rt = hmm_finish(pt, t, n);
        if (lane == 0) { a.out_a[s] = ra; a.out_t[s] = rt; a.out_n_events[s] = n_ev; }
    }
}

}  // namespace

void dnb_launch_hmm_forward(const double *obs, const uint64_t *obs_off, const char *seq, const double *shift,
                            const double *scale, const double *epb, size_t n_sites, uint32_t window,
                            const DnbModelDev &unl, const DnbModelDev &ana, double *out_analogue, double *out_thymidine,
                            cudaStream_t s) {
    if (n_sites == 0) return;
    hmm_sites_kernel<<<(unsigned)((n_sites + HMM_WARPS - 1) / HMM_WARPS), HMM_WARPS * 32, 0, s>>>(
        obs, obs_off, seq, shift, scale, epb, n_sites, window, unl, ana, out_analogue, out_thymidine);
}

void dnb_launch_llr_sites(const DnbBatchView &v, const DnbLlrArgs &a, cudaStream_t s) {
    if (v.n_reads == 0) return;
    llr_sites_kernel<<<v.n_reads, 256, 0, s>>>(v, a);
}

void dnb_launch_llr_forward(const DnbBatchView &v, const DnbLlrArgs &a, uint64_t n_sites_total, const DnbModelDev &unl,
                            const DnbModelDev &ana, int device, cudaStream_t s) {
    if (n_sites_total == 0) return;
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, llr_forward_kernel, HMM_WARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)sms * per_sm;
    const uint64_t need = (n_sites_total + HMM_WARPS - 1) / HMM_WARPS;
    if (grid > need) grid = need;
    llr_forward_kernel<<<(unsigned)grid, HMM_WARPS * 32, 0, s>>>(v, a, n_sites_total, unl, ana);
}
