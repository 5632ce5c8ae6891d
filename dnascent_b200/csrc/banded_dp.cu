// banded_dp.cu -- adaptive banded event-to-k-mer alignment: band fill (one warp per read) and backtrace/QC.
//
// Replaces (reference, paths relative to /root/reference):
//   logProbabilityMatch                          src/event_handling.cpp:116-137
//   adaptive_banded_simple_event_align, fill     src/event_handling.cpp:148-312
//   end-cell scan                                src/event_handling.cpp:321-340
//   backtrace, cleaned (signal,rank) vectors, QC src/event_handling.cpp:347-443
//
// Layout.  The reference allocates n_bands x 100 floats + n_bands x 100 bytes per read.  Here a band lives in
// registers: lane L of the warp owns offsets 4L..4L+3 (lanes 25..31 own nothing and stay at -inf), with the two
// previous bands, the scaled event level x_e and the model level mu_k of each owned cell.  A band step is
// warp-uniform: Suzuki's rule is evaluated from cells 0 and 99 (two shuffles), then either the k-mer registers
// shift one cell down (right move) or the event registers shift one cell up (down move) -- one lane-to-lane
// shuffle each -- and the three neighbours of every cell are register moves plus one edge shuffle.  Only
// 2 bits of trace per cell and 1 move bit per band go to HBM (one 32-byte row per band), plus the running
// end-cell maximum; the band scores themselves are never stored.
#include <cfloat>
#include <cmath>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define DP_WARPS 4
#define FULL 0xffffffffu

namespace {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double shfl_down_d(double v) { return __shfl_down_sync(FULL, v, 1); }
__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(FULL, v, 1); }

// event_handling.cpp:133-136 with sigma == 0.14 (import_poreModel_staticStdv, data_IO.cpp:170)
__device__ __forceinline__ float emission_static(double x, double mu, double emit_const) {
    float a = d2f(dDiv(dSub(x, mu), 0.14));
    float t = fMul(fMul(-0.5f, a), a);
    return d2f(dAdd(emit_const, (double)t));
}

__global__ void __launch_bounds__(DP_WARPS * 32) banded_dp_kernel(DnbBatchView v, DnbDpArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * DP_WARPS + (threadIdx.x >> 5);
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    if (v.status[r] != 0) {
        if (lane == 0) { a.end_event[r] = -1; a.end_ll_event[r] = 0; a.end_score[r] = -INFINITY; }
        return;
    }
    const int E = (int)v.n_events[r];
    const int K = (int)(v.q_off[r + 1] - v.q_off[r]) - DNB_K + 1;
    const double *__restrict__ x = a.x_e + v.ev_off[r];
    const double *__restrict__ mu = a.mu_q + v.q_off[r];
    const double lp_skip = a.lp[4 * r + 0], lp_stay = a.lp[4 * r + 1], lp_step = a.lp[4 * r + 2], lp_trim = a.lp[4 * r + 3];
    const double emit_const = a.emit_const;
    uint8_t *rows = a.trace + a.band_off[r] * DNB_TRACE_ROW;
    const int n_bands = E + K + 2;
    const float NINF = -INFINITY;

    // ---- bands 0 and 1 (event_handling.cpp:213-228) ----
    float p1[4], p2[4];
    double xe[4], mk[4];
    int ll_e = DNB_BW / 2, ll_k = -1 - DNB_BW / 2;     // lower-left of band 1 = move_down(band 0)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int o = lane * 4 + j;
        p2[j] = (o == DNB_BW / 2) ? 0.0f : NINF;                 // bands[0][50] = 0
        p1[j] = (o == DNB_BW / 2) ? d2f(lp_trim) : NINF;         // bands[1][50] = lp_trim
        const int e = ll_e - o, km = ll_k + o;
        xe[j] = (e >= 0 && e < E) ? x[e] : 0.0;
        mk[j] = (km >= 0 && km < K) ? mu[km] : 0.0;
    }
    rows[lane] = 0;                                                               // band 0: no trace, move 0
    rows[DNB_TRACE_ROW + lane] = (lane == 12) ? (uint8_t)(DNB_FROM_U << 4) : 0;   // trace[1][50] = FROM_U

    // coalesced 32-wide look-ahead of the next events / k-mers entering the band (double buffered)
    int xbase = ll_e + 1, mbase = ll_k + DNB_BW;     // next event index on a down move / next k-mer on a right move
    double xbuf = (xbase + lane < E) ? x[xbase + lane] : 0.0;
    double xnxt = (xbase + 32 + lane < E) ? x[xbase + 32 + lane] : 0.0;
    double mbuf = (mbase + lane < K) ? mu[mbase + lane] : 0.0;
    double mnxt = (mbase + 32 + lane < K) ? mu[mbase + 32 + lane] : 0.0;

    bool prev_right = false;
    float best_s = NINF;
    int best_e = 0x7fffffff, best_lle = 0;
    unsigned long long fills = 0;

    for (int b = 2; b < n_bands; b++) {
        // Suzuki's rule (event_handling.cpp:237-253)
        const float ll = __shfl_sync(FULL, p1[0], 0);
        const float ur = __shfl_sync(FULL, p1[3], (DNB_BW - 1) / 4);
        const bool right = (ll == NINF && ur == NINF) ? ((b & 1) == 1) : (ll < ur);

        float up[4], left[4], diag[4];
        if (right) {
            ll_k++;
            // k-mer at offset o becomes the old k-mer at o+1; cell 99 receives k-mer ll_k+99
            const double in = shfl_down_d(mk[0]);
            mk[0] = mk[1]; mk[1] = mk[2]; mk[2] = mk[3]; mk[3] = in;
            const int need = ll_k + DNB_BW - 1;
            if (need - mbase == 32) {
                mbuf = mnxt; mbase += 32;
                mnxt = (mbase + 32 + lane < K) ? mu[mbase + 32 + lane] : 0.0;
            }
            const double fresh = shfl_d(mbuf, need - mbase);
            if (lane == (DNB_BW - 1) / 4) mk[3] = fresh;
            // up = band[b-1][o+1], left = band[b-1][o]
            const float e1 = __shfl_down_sync(FULL, p1[0], 1);
            up[0] = p1[1]; up[1] = p1[2]; up[2] = p1[3]; up[3] = e1;
            left[0] = p1[0]; left[1] = p1[1]; left[2] = p1[2]; left[3] = p1[3];
        } else {
            ll_e++;
            // event at offset o becomes the old event at o-1; cell 0 receives event ll_e
            const double in = shfl_up_d(xe[3]);
            xe[3] = xe[2]; xe[2] = xe[1]; xe[1] = xe[0]; xe[0] = in;
            const int need = ll_e;
            if (need - xbase == 32) {
                xbuf = xnxt; xbase += 32;
                xnxt = (xbase + 32 + lane < E) ? x[xbase + 32 + lane] : 0.0;
            }
            const double fresh = shfl_d(xbuf, need - xbase);
            if (lane == 0) xe[0] = fresh;
            // up = band[b-1][o], left = band[b-1][o-1]
            float e1 = __shfl_up_sync(FULL, p1[3], 1);
            if (lane == 0) e1 = NINF;
            up[0] = p1[0]; up[1] = p1[1]; up[2] = p1[2]; up[3] = p1[3];
            left[0] = e1; left[1] = p1[0]; left[2] = p1[1]; left[3] = p1[2];
        }
        // diag = band[b-2][o + d], d = +1 (right,right), 0 (mixed), -1 (down,down)
        if (prev_right && right) {
            const float e2 = __shfl_down_sync(FULL, p2[0], 1);
            diag[0] = p2[1]; diag[1] = p2[2]; diag[2] = p2[3]; diag[3] = e2;
        } else if (!prev_right && !right) {
            float e2 = __shfl_up_sync(FULL, p2[3], 1);
            if (lane == 0) e2 = NINF;
            diag[0] = e2; diag[1] = p2[0]; diag[2] = p2[1]; diag[3] = p2[2];
        } else {
            diag[0] = p2[0]; diag[1] = p2[1]; diag[2] = p2[2]; diag[3] = p2[3];
        }

        // fill range (event_handling.cpp:269-278) and trim cell (:256-265)
        int lo = max(max(-ll_k, ll_e - (E - 1)), 0);
        int hi = min(min(K - ll_k, ll_e + 1), DNB_BW);
        const int o_trim = -1 - ll_k;
        const int ev_trim = ll_e - o_trim;
        const bool trim_ok = o_trim >= 0 && o_trim < DNB_BW && ev_trim >= 0 && ev_trim < E;
        if (lane == 0 && hi > lo) fills += (unsigned long long)(hi - lo);

        float nb[4];
        uint32_t tb = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = lane * 4 + j;
            float m = NINF;
            uint32_t from = 0;
            if (o >= lo && o < hi) {
                const float em = emission_static(xe[j], mk[j], emit_const);
                const float sd = d2f(dAdd(dAdd((double)diag[j], lp_step), (double)em));   // :296
                const float su = d2f(dAdd(dAdd((double)up[j], lp_stay), (double)em));     // :297
                const float sl = d2f(dAdd((double)left[j], lp_skip));                     // :298
                m = sd; from = DNB_FROM_D;                                                // :300-306
                m = su > m ? su : m; from = (m == su) ? DNB_FROM_U : from;
                m = sl > m ? sl : m; from = (m == sl) ? DNB_FROM_L : from;
            } else if (o == o_trim && trim_ok) {
                m = d2f(dMul(lp_trim, (double)((uint32_t)ev_trim + 1u)));                 // :260
                from = DNB_FROM_U;
            }
            nb[j] = m;
            tb |= from << (2 * j);
        }
        uint32_t rowbyte = tb;
        if (lane == 25) rowbyte = right ? 1u : 0u;
        if (lane > 25) rowbyte = 0;
        rows[(size_t)b * DNB_TRACE_ROW + lane] = (uint8_t)rowbyte;

        // end-cell candidate of this band: cell (event b-K-1, last k-mer) (event_handling.cpp:329-340)
        {
            const int e = b - K - 1;
            const int o = (K - 1) - ll_k;
            if (e >= 0 && e < E && o >= 0 && o < DNB_BW && (o >> 2) == lane) {
                const int j = o & 3;
                const float val = j == 0 ? nb[0] : j == 1 ? nb[1] : j == 2 ? nb[2] : nb[3];
                const float s = d2f(dAdd((double)val, dMul((double)(unsigned long long)(E - e), lp_trim)));
                if (s > best_s) { best_s = s; best_e = e; best_lle = ll_e; }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) { p2[j] = p1[j]; p1[j] = nb[j]; }
        prev_right = right;
    }

    // first event index attaining the maximum (strict '>' in ascending event order, :335)
    float gmax = best_s;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(FULL, gmax, d));
    int cand = (best_s == gmax && gmax != NINF) ? best_e : 0x7fffffff;
    int gmin = cand;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) gmin = min(gmin, __shfl_xor_sync(FULL, gmin, d));
    const unsigned who = __ballot_sync(FULL, cand == gmin && gmin != 0x7fffffff);
    int lle = 0;
    if (who) lle = __shfl_sync(FULL, best_lle, __ffs(who) - 1);
    if (lane == 0) {
        if (gmin == 0x7fffffff) {
            a.end_event[r] = -1; a.end_ll_event[r] = 0; a.end_score[r] = NINF;
            v.status[r] = DNB_READ_UNDEFINED;   // the reference would backtrace from an out-of-band cell
        } else {
            a.end_event[r] = gmin; a.end_ll_event[r] = lle; a.end_score[r] = gmax;
        }
        atomicAdd(a.cells, fills);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backtrace + QC.  One warp per read: all lanes stage a 64-row window of the trace into shared memory with
// 16-byte loads, lane 0 walks it (the walk is a serial pointer chase; the accumulations must keep the
// reference's order: sum_emission adds floats into a double in walk order, vectorMean sums in push order).
// ---------------------------------------------------------------------------------------------------------------
#define BT_WARPS 4
#define BT_ROWS 64

__global__ void __launch_bounds__(BT_WARPS * 32) backtrace_kernel(DnbBatchView v, DnbBtArgs a) {
    __shared__ __align__(16) uint8_t win[BT_WARPS][BT_ROWS * DNB_TRACE_ROW];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * BT_WARPS + w;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    if (v.status[r] != 0) {
        if (lane == 0) {
            a.n_align[r] = 0; a.n_cleaned[r] = 0; a.avg_log_emission[r] = 0.0; a.spanned[r] = 0; a.max_gap[r] = 0;
        }
        return;
    }
    const int E = (int)v.n_events[r];
    const int K = (int)(v.q_off[r + 1] - v.q_off[r]) - DNB_K + 1;
    const uint32_t Kref = (uint32_t)((v.r_off[r + 1] - v.r_off[r]) - DNB_K + 1);
    const double *__restrict__ x = a.dp.x_e + v.ev_off[r];
    const double *__restrict__ mu = a.dp.mu_q + v.q_off[r];
    const float *__restrict__ evm = v.ev_mean + v.ev_off[r];
    const int32_t *__restrict__ q2r = v.q2r + v.q_off[r];
    const uint32_t *__restrict__ rr = a.rank_ref + v.r_off[r];
    const uint8_t *rows = a.dp.trace + a.dp.band_off[r] * DNB_TRACE_ROW;
    uint32_t *pairs = a.al_pairs_rev + 2 * a.al_off[r];
    double *cls = a.cl_signal + a.cl_off[r];
    uint32_t *clr = a.cl_rank + a.cl_off[r];
    const double emit_const = a.dp.emit_const;
    (void)E;

    int e = a.dp.end_event[r], k = K - 1, lle = a.dp.end_ll_event[r];
    double sum_em = 0.0, n_aligned = 0.0, buf_total = 0.0;
    uint32_t buf_n = 0, na = 0, nc = 0;
    int gap = 0, max_gap = 0, last_k = -1;
    bool bad = false, done = false;

    while (!done) {
        const int hi = e + k + 2;
        const int lo = max(hi - (BT_ROWS - 1), 0);
        const int nrows = hi - lo + 1;
        // stage rows [lo, hi]
        const uint4 *src = reinterpret_cast<const uint4 *>(rows + (size_t)lo * DNB_TRACE_ROW);
        uint4 *dst = reinterpret_cast<uint4 *>(win[w]);
#pragma unroll
        for (int i = 0; i < (BT_ROWS * DNB_TRACE_ROW / 16) / 32; i++) {
            const int c = lane + 32 * i;
            if ((c >> 1) < nrows) dst[c] = src[c];
        }
        __syncwarp();
        if (lane == 0) {
            const uint8_t *wb = win[w];
            while (k >= 0 && e >= 0) {
                const int b = e + k + 2;
                if (b - 1 < lo) break;                         // next window
                pairs[2 * na] = (uint32_t)e; pairs[2 * na + 1] = (uint32_t)k; na++;       // :359
                last_k = k;
                const float lp = emission_static(x[e], mu[k], emit_const);                // :363
                sum_em = dAdd(sum_em, (double)lp);
                n_aligned = dAdd(n_aligned, 1.0);
                const int off = lle - e;
                if (off < 0 || off >= DNB_BW) { bad = true; break; }
                const uint8_t *row = wb + (size_t)(b - lo) * DNB_TRACE_ROW;
                const uint32_t from = (row[off >> 2] >> (2 * (off & 3))) & 3u;
                const int down_b = row[25] ? 0 : 1;            // band b was placed by a down move
                if (from == DNB_FROM_D) {
                    buf_total = dAdd(buf_total, (double)evm[e]); buf_n++;
                    const int32_t qr = q2r[k];
                    if (qr >= 0 && (uint32_t)qr < Kref) {                                 // :386-393
                        clr[nc] = rr[qr];
                        cls[nc] = dDiv(buf_total, (double)buf_n);                         // vectorMean, common.h:184
                        nc++;
                    }
                    buf_total = 0.0; buf_n = 0;
                    const int down_b1 = (wb + (size_t)(b - 1 - lo) * DNB_TRACE_ROW)[25] ? 0 : 1;
                    lle -= down_b + down_b1;
                    k--; e--; gap = 0;
                } else if (from == DNB_FROM_U) {
                    buf_total = dAdd(buf_total, (double)evm[e]); buf_n++;
                    lle -= down_b;
                    e--; gap = 0;
                } else {
                    lle -= down_b;
                    k--; gap++;
                    max_gap = max(max_gap, gap);
                }
            }
            if (bad || k < 0 || e < 0) done = true;
        }
        done = __shfl_sync(FULL, done, 0);
        e = __shfl_sync(FULL, e, 0);
        k = __shfl_sync(FULL, k, 0);
        __syncwarp();
    }
    if (lane == 0) {
        if (bad) {
            v.status[r] = DNB_READ_UNDEFINED;
            a.n_align[r] = 0; a.n_cleaned[r] = 0; a.avg_log_emission[r] = 0.0; a.spanned[r] = 0; a.max_gap[r] = 0;
            return;
        }
        const double avg = dDiv(sum_em, n_aligned);                                        // :420
        const bool spanned = na > 0 && last_k == 0;     // front().second == 0; back().second == K-1 holds by construction
        a.avg_log_emission[r] = avg;
        a.spanned[r] = spanned ? 1 : 0;
        a.max_gap[r] = max_gap;
        a.n_cleaned[r] = nc;
        bool fail = avg < a.min_avg_log_emission || !spanned || max_gap > a.max_gap_threshold;  // :433
        if (!fail && nc < 1000) fail = true;                                                // :438
        if (fail) { v.status[r] = DNB_READ_QC_FAIL; a.n_align[r] = 0; }
        else a.n_align[r] = na;
    }
}

__global__ void compact_alignment_kernel(DnbBatchView v, const uint64_t *al_off, const uint32_t *al_pairs_rev,
                                         const uint32_t *n_align, const uint64_t *out_off, uint32_t *out_pairs) {
    const uint32_t r = blockIdx.x;
    const uint32_t n = n_align[r];
    const uint2 *src = reinterpret_cast<const uint2 *>(al_pairs_rev) + al_off[r];
    uint2 *dst = reinterpret_cast<uint2 *>(out_pairs) + out_off[r];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[n - 1 - i];   // std::reverse, :413
}

}  // namespace

void dnb_launch_banded_dp(const DnbBatchView &v, const DnbDpArgs &a, cudaStream_t s) {
    if (v.n_reads == 0) return;
    banded_dp_kernel<<<(v.n_reads + DP_WARPS - 1) / DP_WARPS, DP_WARPS * 32, 0, s>>>(v, a);
}

void dnb_launch_backtrace(const DnbBatchView &v, const DnbBtArgs &a, cudaStream_t s) {
    if (v.n_reads == 0) return;
    backtrace_kernel<<<(v.n_reads + BT_WARPS - 1) / BT_WARPS, BT_WARPS * 32, 0, s>>>(v, a);
}

void dnb_launch_compact_alignment(const DnbBatchView &v, const uint64_t *al_off, const uint32_t *al_pairs_rev,
                                  const uint32_t *n_align, const uint64_t *out_off, uint32_t *out_pairs,
                                  cudaStream_t s) {
    if (v.n_reads == 0) return;
    compact_alignment_kernel<<<v.n_reads, 128, 0, s>>>(v, al_off, al_pairs_rev, n_align, out_off, out_pairs);
}
