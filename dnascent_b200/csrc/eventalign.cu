// eventalign.cu -- windowed Viterbi re-alignment of events to the reference.
//
// Replaces (reference, paths relative to /root/reference):
//   builtinViterbi    src/alignment.cpp:193-516   3-state (D, M, I) profile HMM over one reference window
//   eventalign        src/alignment.cpp:547-744   the window chain of one read, up to (not including) the text
//                                                 formatting and r.addSignal, which stay on the host (shim)
//
// Shape of the work.  A read is a SERIAL chain of windows: where window w+1 starts (on the reference and in
// r.eventAlignment) depends on the last match state of window w's Viterbi path (alignment.cpp:740-741), so a read is
// one unit of work and reads are the parallel axis -- one warp per read, persistent warps pulling reads from an
// atomic counter.  Inside a window the lanes are the HMM states (n = windowLength - 8 <= 67 states -> up to three
// register slots per lane, state i = 32*slot + lane); time steps are sequential.
//
// The recurrences are evaluated with the reference's own operation order (each candidate is (prev + transition) +
// emission, alignment.cpp:350-381), with -inf standing for the reference's NaN == log(0): every use of those values
// in builtinViterbi is `+` (NaN/-inf absorbing) or lnGreaterThan (NaN is smaller than everything, two NaNs are not
// greater than each other), so the two conventions take the same decisions.
//
// The deletion chain D[i] = max(M[i-1] + m2d, D[i-1] + d2d) runs along the state axis inside one time step.  It is
// solved exactly without serialising the warp: F(a) = fl(a + d2d) is monotone, so D[i] = max_d F^d(v0[i-d]) with
// v0[i] = M[i-1] + m2d, a max-scan that is evaluated by distance doubling (k = 1, 2, 4, ...), F^k being k literal
// roundings.  (ncu showed chains as long as the window at almost every step: one good match state beats every later
// state's own v0, so a sweep-until-stable loop needed ~n/2 warp-wide sweeps; doubling needs log2 n rounds.)
//
// Emission (alignment.cpp:344): eln(normalPDF(mu, 0.14, x)) = log(c * exp(y)), y = -(x-mu)^2 / (2*0.14^2).  For
// y >= -700 this is computed as log(c) + y (log(c), c and 2*sigma^2 come from the host's libm); below that, where
// exp() is subnormal or underflows and the reference's value really does depart from log(c) + y (it becomes NaN at
// y < -745.13), the literal exp/log form is used.  Either way the emission can differ from glibc's by a few ulps:
// the state path is a discrete result and is identical unless two candidates tie to ~1e-15 (tests compare the
// reference's text byte for byte); see DESIGN.md s.4.6.
#include <cfloat>
#include <cmath>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define EA_WARPS 4
#define EA_SLOTS 3
#define FULL 0xffffffffu
#define NEG_INF (-INFINITY)

namespace {

__device__ __forceinline__ bool base_defined(char c) { return c == 'A' || c == 'T' || c == 'G' || c == 'C'; }

// referenceDefined (alignment.cpp:519-544) over ref[0, len): len <= 96
__device__ __forceinline__ bool warp_defined(const char *ref, unsigned len, int lane) {
    bool ok = true;
    for (unsigned i = lane; i < len; i += 32) ok = ok && base_defined(ref[i]);
    return __all_sync(FULL, ok);
}

__device__ __forceinline__ uint32_t kmer_rank(const char *s) {      // kmer2index, src/data_IO.cpp:129-141
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < DNB_K; i++) r = r * 4u + dnb_base_code(s[i]);
    return r;
}

// c[s] = F^k(c[s]) for the slots in use, F(a) = fl(a + d): K literal roundings per slot, straight-line code (the
// slots' chains are independent, so they overlap in the FP64 pipe)
template <int K, int NS>
__device__ __forceinline__ void add_chain(double (&c)[EA_SLOTS], double d) {
#pragma unroll
    for (int j = 0; j < K; j++) {
#pragma unroll
        for (int s = 0; s < NS; s++) c[s] = dAdd(c[s], d);
    }
}
struct EaRead {
    const char *ref;
    uint32_t rlen;
    const int32_t *r2q;
    const uint2 *pairs;
    uint32_t n_align;
    const float *evm;
    double shift, scale;
    double m12m1_int, m12m1_ext, m12m1_ext_or_int, m12m1_ext_or_d;   // per-read transitions (host libm)
};

// value of state i-k for the NS slots in use (slot s, lane l <-> i = 32 s + l), -inf where i < k; k is warp-uniform,
// 1..31, 32 or 64.  One rotation per slot serves both the in-slot and the wrapped-in-from-the-previous-slot lanes.
template <int NS>
__device__ __forceinline__ void shift_states(const double (&X)[EA_SLOTS], double (&out)[EA_SLOTS], int lane, int k) {
    if (k < 32) {
        const int src = (lane - k) & 31;
        double rot[EA_SLOTS];
#pragma unroll
        for (int s = 0; s < NS; s++) rot[s] = __shfl_sync(FULL, X[s], src);
#pragma unroll
        for (int s = 0; s < NS; s++) out[s] = lane >= k ? rot[s] : (s ? rot[s ? s - 1 : 0] : NEG_INF);
    } else if (k == 32) {
#pragma unroll
        for (int s = 0; s < NS; s++) out[s] = s ? X[s ? s - 1 : 0] : NEG_INF;
    } else {
#pragma unroll
        for (int s = 0; s < NS; s++) out[s] = s >= 2 ? X[s >= 2 ? s - 2 : 0] : NEG_INF;
    }
}

// builtinViterbi's forward pass (alignment.cpp:252-440) over the ns observations of one window, compiled per number
// of register slots in use so that no instruction is spent on states the window does not have.
template <int NS>
__device__ __forceinline__ void viterbi_forward(const DnbEaArgs &a, const EaRead &rd, const double *obs, uint32_t ns, int n,
                                                const double (&mu)[EA_SLOTS], double (&I)[EA_SLOTS], double (&M)[EA_SLOTS],
                                                double (&D)[EA_SLOTS], uint8_t *bt, int lane) {
    double start_prev = 0.0;
    const bool first_state = lane == 0;                                                     // i == 0 lives in slot 0, lane 0
    // M and D of the previous step shifted by one state: both are by-products of the previous step's deletion scan
    double Mm1[EA_SLOTS], Dm1[EA_SLOTS];
    shift_states<NS>(M, Mm1, lane, 1); shift_states<NS>(D, Dm1, lane, 1);
    double x_next = obs[0];
    for (uint32_t t = 0; t < ns; t++) {
        const double x = x_next;
        if (t + 1 < ns) x_next = obs[t + 1];                                                // off the critical path
        double Im1[EA_SLOTS];
        shift_states<NS>(I, Im1, lane, 1);
        double In[EA_SLOTS], Mn[EA_SLOTS];
        uint32_t code[EA_SLOTS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int i = 32 * s + lane;
            const double d = dSub(x, mu[s]);
            // -(x-mu)^2 / (2 sigma^2) as a multiply by the rounded reciprocal: at most 1.5 ulp from the IEEE quotient,
            // inside the few-ulp difference log(c) + y already has to glibc's log(c * exp(y)) (see the header)
            const double y = dMul(-dMul(d, d), a.inv_two_sigma2);
            double mp;
            if (y >= -700.0) mp = dAdd(a.ln_c, y);
            else { const double v = dMul(a.c, exp(y)); mp = v == 0.0 ? NEG_INF : log(v); }
            // insertion (:276-300 for i == 0, :350-369 otherwise).  The reference adds insProb = 0.0 to every candidate;
            // that only turns -0.0 into +0.0, which no comparison or later sum can tell apart, so it is not evaluated.
            const double i0 = dAdd(I[s], a.i2i), i1 = dAdd(M[s], a.m2i);
            double m = i0; uint32_t ai = 0;
            if (i1 > m) { m = i1; ai = 1; }
            if (s == 0) {                                                                   // third candidate of state 0 only
                const double i2 = first_state ? dAdd(start_prev, a.m2i) : NEG_INF;
                if (i2 > m) { m = i2; ai = 2; }
            }
            In[s] = m;
            // match (:303-322 for i == 0, :372-401 otherwise): state 0 has two candidates (stay, enter from the start
            // state), the others four (from I, M, M-stay, D of the previous state); evaluated as one select chain with
            // state 0's operands swapped in and its missing candidates at -inf (a strict > never picks them)
            double p0 = Im1[s], c0 = a.i2m, p1 = Mm1[s], c1 = rd.m12m1_ext, p2 = M[s], p3 = Dm1[s];
            if (s == 0 && first_state) { p0 = M[s]; c0 = rd.m12m1_int; p1 = start_prev; c1 = rd.m12m1_ext_or_int; p2 = NEG_INF; p3 = NEG_INF; }
            const double m0 = dAdd(dAdd(p0, c0), mp), m1 = dAdd(dAdd(p1, c1), mp);
            const double m2 = dAdd(dAdd(p2, rd.m12m1_int), mp), m3 = dAdd(dAdd(p3, a.d2m), mp);
            uint32_t am = 0;
            m = m0;
            if (m1 > m) { m = m1; am = 1; }
            if (m2 > m) { m = m2; am = 2; }
            if (m3 > m) { m = m3; am = 3; }
            Mn[s] = m;
            code[s] = ai | (am << 2);
            if (i >= n) { In[s] = NEG_INF; Mn[s] = NEG_INF; }
        }
        // deletion (:325-327, 405-428): D[i] = max(v0[i], F(D[i-1])) with v0[i] = M_curr[i-1] + m2d and
        // F(a) = fl(a + d2d).  F is monotone, so D[i] = max over d of F^d(v0[i-d]) -- and a strong match state
        // usually wins for EVERY state after it (d2d = log 0.3 costs less than a mismatching emission), so the chain
        // is as long as the window.  Exact scan by doubling: after the round with distance k every D[i] holds the
        // max over d < 2k; F^k is k dependent roundings, done literally (total <= n adds per lane).  A round that
        // changes nothing proves the fixed point (any longer chain factors through states that did not grow).
        double v0[EA_SLOTS], Dn[EA_SLOTS];
        shift_states<NS>(Mn, Mm1, lane, 1);                                                 // also the next step's M[i-1]
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int i = 32 * s + lane;
            v0[s] = (i > 0 && i < n) ? dAdd(Mm1[s], a.m2d) : NEG_INF;
            Dn[s] = v0[s];
        }
        {
            bool more = true;
#define EA_ROUND(KK)                                                                                                  \
            if (more && (KK) <= n - 2) {                                                                              \
                double c[EA_SLOTS];                                                                                   \
                shift_states<NS>(Dn, c, lane, (KK));                                                                  \
                add_chain<(KK), NS>(c, a.d2d);                                                                        \
                bool grew = false;                                                                                    \
                _Pragma("unroll") for (int s = 0; s < NS; s++) {                                                      \
                    const int i = 32 * s + lane;                                                                      \
                    if (i < n && c[s] > Dn[s]) { Dn[s] = c[s]; grew = true; }        /* i == 0: c is -inf */          \
                }                                                                                                     \
                more = __any_sync(FULL, grew);                                                                        \
            }
            EA_ROUND(1) EA_ROUND(2) EA_ROUND(4) EA_ROUND(8) EA_ROUND(16) EA_ROUND(32)
            if (NS == 3) { EA_ROUND(64) }
#undef EA_ROUND
        }
        // lnArgMax: D over M only when strictly greater; the shifted D is also the next step's D[i-1]
        shift_states<NS>(Dn, Dm1, lane, 1);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int i = 32 * s + lane;
            if (i < n && dAdd(Dm1[s], a.d2d) > v0[s]) code[s] |= 1u << 4;                  // i == 0: -inf > -inf is false
        }
        uint8_t *row = bt + (size_t)t * (EA_SLOTS * 32);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            row[32 * s + lane] = (uint8_t)code[s];
            I[s] = In[s]; M[s] = Mn[s]; D[s] = Dn[s];
        }
        start_prev = NEG_INF;                                                               // :433 (start_curr = NAN)
    }
}


__device__ void eventalign_read(const DnbEaArgs &a, const EaRead &rd, uint32_t r, int lane, double *obs, uint32_t *obs_ev,
                                uint8_t *bt) {
    const unsigned k = DNB_K, W = a.window;
    const unsigned lt = (1u << lane) - 1u;
    dnb_eventalign_rec *recs = a.recs + a.rec_off[r];
    const uint64_t cap = a.rec_off[r + 1] - a.rec_off[r];
    uint64_t nrec = 0;
    long read_head = 0;
    unsigned ri = 0;
    int status = DNB_READ_OK;

    while (ri < rd.rlen - k + 1) {
        const unsigned bases_to_end = rd.rlen - ri;
        unsigned wl = min(bases_to_end, W);
        const char *ws = rd.ref + ri;
        if ((double)bases_to_end > 1.5 * (double)W) {                                       // alignment.cpp:565-595
            const unsigned bl = (unsigned)(1.5 * (double)wl);
            if (!warp_defined(ws, bl, lane)) { ri += wl; continue; }
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
        if (!warp_defined(ws, wl, lane)) { ri += wl; continue; }                           // :597-604
        const uint32_t lo = (uint32_t)rd.r2q[ri], hi = (uint32_t)rd.r2q[ri + wl - k + 1];

        // ---- events aligned into the window (:611-632) ----
        uint32_t ns = 0;
        bool first = true;
        for (long j0 = read_head; j0 < (long)rd.n_align; j0 += 32) {
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
            const bool good = in && 0. < em && em < 250.;                                   // :623
            const unsigned gm = __ballot_sync(FULL, good);
            const uint32_t pos = ns + __popc(gm & lt);
            // the scaled observation (:272, 344) is computed here, once per event, instead of by every lane at its time step
            if (good && pos < a.t_max) { obs[pos] = dDiv(dSub(em, rd.shift), rd.scale); obs_ev[pos] = pr.x; }
            ns += __popc(gm);
            if (stopmask) break;
        }
        const int indel = ((int)hi - (int)lo) - (int)(wl - k + 1);                          // :635-638
        if (ns < 2) { ri += wl; continue; }                                                 // :641
        if (ns > a.t_max) { status = DNB_READ_OVERFLOW; break; }
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
        const int nslots = (n + 31) >> 5;                                                   // register slots in use (1..3)
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
        // ---- traceback (:476-505) and the two passes over the state labels (:655-736), by lane 0.  Every
        // observation is consumed by exactly one M or I state, so the record of observation t (kept when it is not
        // after the last match) goes straight to slot t of this window's record block.
        uint32_t last_m_ev = 0, last_m_ref = 0;
        int found = 0, overflow = 0;
        if (lane == 0) {
            int ty = end_ty, i = n - 1;
            long t = (long)ns;
            for (;;) {
                if (ty != 0) {
                    if (t == 0) { ty = 0; i = 0; continue; }      // never-written column 0 of an M/I row reads as "D 0, t 0"
                    const uint32_t ev = (uint32_t)(t - 1);
                    if (ty == 1 && !found) { found = 1; last_m_ev = ev; last_m_ref = (uint32_t)i; if (nrec + ev + 1 > cap) { overflow = 1; break; } }
                    if (found) {
                        dnb_eventalign_rec rc;
                        rc.event = obs_ev[ev]; rc.ref_pos = ri + (uint32_t)i; rc.indel_score = indel; rc.label = (uint32_t)ty;
                        recs[nrec + ev] = rc;
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
        }
        overflow = __shfl_sync(FULL, overflow, 0);
        if (overflow) { status = DNB_READ_OVERFLOW; break; }
        found = __shfl_sync(FULL, found, 0);
        last_m_ev = __shfl_sync(FULL, last_m_ev, 0);
        last_m_ref = __shfl_sync(FULL, last_m_ref, 0);
        if (found) nrec += (uint64_t)last_m_ev + 1;
        read_head += (long)last_m_ev + 1;                                                   // :740-741
        ri += last_m_ref + 1;
        __syncwarp();
    }
    if (lane == 0) {
        a.status[r] = status;
        a.n_rec[r] = status == DNB_READ_OK ? (uint32_t)nrec : 0u;
    }
}

__global__ void __launch_bounds__(EA_WARPS * 32) eventalign_kernel(DnbEaArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t wslot = blockIdx.x * EA_WARPS + (threadIdx.x >> 5);
    double *obs = a.scratch_obs + (size_t)wslot * a.t_max;
    uint32_t *obs_ev = a.scratch_ev + (size_t)wslot * a.t_max;
    uint8_t *bt = a.scratch_bt + (size_t)wslot * a.t_max * (EA_SLOTS * 32);
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(a.next_read, 1u);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        if (a.order) r = a.order[r];
        if (a.status[r] != DNB_READ_OK) { if (lane == 0) a.n_rec[r] = 0; continue; }   // rejected by the host (see capi.cu)
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
        eventalign_read(a, rd, r, lane, obs, obs_ev, bt);
        __syncwarp();
    }
}

}  // namespace

unsigned dnb_eventalign_grid(int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eventalign_kernel, EA_WARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    return (unsigned)(sms * per_sm);
}
unsigned dnb_eventalign_warps_per_block(void) { return EA_WARPS; }
size_t dnb_eventalign_bt_row_bytes(void) { return EA_SLOTS * 32; }

void dnb_launch_eventalign(const DnbEaArgs &a, unsigned grid, cudaStream_t s) {
    if (a.n_reads == 0) return;
    eventalign_kernel<<<grid, EA_WARPS * 32, 0, s>>>(a);
}
