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
// reference's text byte for byte); see DESIGN.md s.4.5.
#include "eventalign_core.cuh"

namespace {

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
