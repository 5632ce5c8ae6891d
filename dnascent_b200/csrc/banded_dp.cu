// banded_dp.cu -- adaptive banded event-to-k-mer alignment: band fill and backtrace/QC, one warp per read, one kernel.
//
// Replaces (reference, paths relative to /root/reference):
//   logProbabilityMatch                          src/event_handling.cpp:116-137
//   adaptive_banded_simple_event_align, fill     src/event_handling.cpp:148-312
//   end-cell scan                                src/event_handling.cpp:321-340
//   backtrace, cleaned (signal,rank) vectors, QC src/event_handling.cpp:347-443
//
// Layout.  The reference allocates n_bands x 100 floats + n_bands x 100 bytes per read.  Here a band lives in
// registers, indexed by EVENT, not by band offset: the cell of the current band whose event index is e sits in slot
// e mod 128 = (lane, register) = ((e >> 2) & 31, e & 3).  With that indexing the three neighbours of a cell are at
// fixed places whatever the band did -- left = same slot of band b-1, up = slot-1 of band b-1, diag = slot-1 of
// band b-2 -- so a band step needs no register shuffling of scores: one lane-to-lane shuffle per band array for the
// slot-1 edge, the two previous bands swap roles (loop unrolled by two).  The scaled event level x_e of a slot never
// moves (a new event is dropped into its slot on a down move); the model level mu_k moves one slot per band like a
// conveyor (a new k-mer is dropped in on a right move).  Suzuki's rule reads the two end cells with two shuffles.
//
// Arithmetic.  Scores are kept as doubles that hold float values.  The reference's `float = double expression`
// roundings are done with the exponent-aligned magic-constant add (exact round-to-nearest-even at 24 bits, 2 integer
// + 2 FP64 instructions) instead of F2F conversions, which run at quarter rate on the XU pipe (50 % busy in the
// first version, profiles/r1_captureA).  (x - mu)/0.14 is a multiply by the correctly rounded reciprocal, guarded:
// if the product lies within 8 ulps of a float rounding boundary the whole band redoes its four quotients with
// the IEEE division (probability ~3e-6 per band).  -inf is represented by -FLT_MAX, which is a fixed point of every
// update and compares like -inf does.
//
// HBM traffic per band: one 32-byte row (2-bit trace code per slot) + 1 move bit; scores never leave registers.
#include <cfloat>
#include <cmath>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

// One warp per read and ONE WARP PER CTA: warps of a CTA do not cooperate, and a CTA's slot is only handed to the next
// CTA when its last warp retires, so 4-warp CTAs idled slots behind their longest read (measured at 30k reads:
// 1-warp CTAs alone -6 % on the fill).  DP_MIN_BLOCKS caps the registers: 24 resident warps per SM at 80 registers
// without spills, against 16 at 128 (fill 858 -> 792 ms, backtrace 164 -> 128 ms per 30k-read step, profiles/r2s_*).
#ifndef DP_WARPS
#define DP_WARPS 1
#endif
#ifndef DP_MIN_BLOCKS
#define DP_MIN_BLOCKS 24     // resident CTAs per SM the register allocation must allow (DP_WARPS warps each)
#endif
#ifndef DP_LEAN
#define DP_LEAN 1            // band bookkeeping with carried counters / unsigned range tests (0: the first, literal forms)
#endif
#define FULL 0xffffffffu
#define NEG_SENT (-3.4028234663852886e38)   /* (double)(-FLT_MAX): stands for -INFINITY */

namespace {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// event_handling.cpp:133-136 with sigma == 0.14 (import_poreModel_staticStdv, data_IO.cpp:170): literal form
__device__ __forceinline__ float emission_static(double x, double mu, double emit_const) {
    float a = d2f(dDiv(dSub(x, mu), 0.14));
    float t = fMul(fMul(-0.5f, a), a);
    return d2f(dAdd(emit_const, (double)t));
}

// (double)(float)v for |v| in the normal float range (or 0, or the sentinel): round to nearest even at 24 bits by
// adding and subtracting 1.5 * 2^(exponent(v) + 29), built from v's own sign and exponent field
__device__ __forceinline__ double rn24(double v) {
    const int mhi = (__double2hiint(v) & 0xFFF00000) + 0x01D80000;
    const double M = __hiloint2double(mhi, 0);
    return __dsub_rn(__dadd_rn(v, M), M);
}

// The same rounding by Veltkamp's split with C = 2^29 + 1: p = v*C, result = p - (p - v).  Three FP64 instructions
// and no integer ones (the magic-constant form is 2 FP64 + 3 integer), so the two forms trade FP64-pipe cycles
// against issue slots; DP_VELTKAMP picks how many of a cell's six roundings use this one.  Exact RN-even on 24 bits
// including ties (checked against the (float) cast on 2*10^8 doubles, 2.5*10^7 of them exact ties).
#ifndef DP_VELTKAMP
#define DP_VELTKAMP 6
#endif
__device__ __forceinline__ double rn24v(double v) {
    const double p = __dmul_rn(v, 536870913.0);
    return __dsub_rn(p, __dsub_rn(p, v));
}
template <int kSite>
__device__ __forceinline__ double rn24s(double v) { return kSite < DP_VELTKAMP ? rn24v(v) : rn24(v); }

struct DpConst {
    double lp_skip, lp_stay, lp_step, lp_trim, emit_const, inv_sigma;
};

// Which of a cell's six float roundings go through the conversion unit (F2F on the XU pipe: one instruction, but an
// eighth of the FP64 rate) instead of three FP64 instructions: bit 0 = the emission's a and a*a (a*a is then a
// native FP32 multiply), bit 1 = the three candidate scores (the maximum and the tie rule are then FP32 compares and
// single-register selects, and the winner is widened once).  The kernel is issue-bound, so what counts is the
// instruction total as long as no single pipe saturates (XU: 8 cycles per warp instruction and scheduler).
// DP_XU holds one such 2-bit choice per register slot j (nibble j), so the XU load can be set in steps of one cell.
#ifndef DP_XU
#define DP_XU 0x3332         // measured with 24 resident warps per SM: 0x3333 792 ms, 0x3332 780, 0x3331 780 per 30k-read step
#endif

// one cell: q = (x - mu)/sigma still in double; returns the new score, sets `from`
template <int kXU>
__device__ __forceinline__ double cell_update(double diag, double up, double left, double q, const DpConst &c, uint32_t &from) {
    double r;
    if (kXU & 1) {
        const float a = d2f(q);                                             // event_handling.cpp:133
        r = (double)fMul(a, a);                                             // (-0.5f*a)*a == -0.5f*float(a*a)
    } else {
        const double a = rn24s<0>(q);
        r = rn24s<1>(dMul(a, a));
    }
    const double em = rn24s<2>(__fma_rn(r, -0.5, c.emit_const));            // product exact: one rounding, as C + (double)t
    const double xd = dAdd(dAdd(diag, c.lp_step), em);                      // event_handling.cpp:296
    const double xu = dAdd(dAdd(up, c.lp_stay), em);                        // :297
    const double xl = dAdd(left, c.lp_skip);                                // :298
    if (kXU & 2) {
        const float sd = d2f(xd), su = d2f(xu), sl = d2f(xl);
        float m = sd;                                                       // :300-306, ties: L over U over D
        from = DNB_FROM_D;
        if (su >= m) { m = su; from = DNB_FROM_U; }
        if (sl >= m) { m = sl; from = DNB_FROM_L; }
        return (double)m;
    }
    const double sd = rn24s<4>(xd), su = rn24s<5>(xu), sl = rn24s<3>(xl);
    double m = sd;
    from = DNB_FROM_D;
    if (su >= m) { m = su; from = DNB_FROM_U; }
    if (sl >= m) { m = sl; from = DNB_FROM_L; }
    return m;
}

// ---------------------------------------------------------------------------------------------------------------
// Band fill of one read by one warp (event_handling.cpp:213-312 and the end-cell scan :321-340).
// ---------------------------------------------------------------------------------------------------------------
struct DpEnd {
    int event;        // event index of the best end cell, -1 if none
    int ll_event;     // band_lower_left.event_idx of that band
    float score;
};

struct DpWarp {
    // per read
    int E, K, lane;
    const double *__restrict__ x;
    const double *__restrict__ mu;
    DpConst c;
    uint8_t *rows;
    uint32_t *moves, *rcum;
    double *sc;                 // shared memory, 128 doubles: the band just computed, by event slot
    // state
    double xe[4], mk[4];        // event level / k-mer level of this lane's four slots
    int ll_e, ll_k;             // lower-left cell of the newest band
    int xbase, mbase;           // coalesced 32-wide look-ahead of the events / k-mers entering the band
    double xbuf, xnxt, mbuf, mnxt;
    double v_ll, v_ur;          // cells 0 and 99 of the newest band (Suzuki's rule reads them), warp-uniform
    double best_s;              // end-cell scan, warp-uniform
    int best_e, best_lle;
    unsigned long long fills;
    uint32_t mvword, rights;
    int e_cand;                 // event of the end-cell candidate of the band being computed (b - K - 1)
    unsigned k_lim, e_lim;      // steady-state limits: max(K - 99, 0), max(E - 99, 0)
};

// value of slot s (warp-uniform) in the four registers of lane s >> 2; only that lane's result is meaningful.
// Written as predicated moves: the `?:` form compiled to 16 SEL + predicate logic per call (58 instructions on the
// right-move path of a band, static SASS of round 1), this is 4 compares + 8 predicated 32-bit moves.
__device__ __forceinline__ void put4(double (&p)[4], int lane, int s, double v) {
    const int key = (lane == (s >> 2)) ? (s & 3) : -1;
    asm("{\n\t.reg .pred q;\n\t"
        "setp.eq.s32 q, %4, 0;\n\t@q mov.f64 %0, %5;\n\t"
        "setp.eq.s32 q, %4, 1;\n\t@q mov.f64 %1, %5;\n\t"
        "setp.eq.s32 q, %4, 2;\n\t@q mov.f64 %2, %5;\n\t"
        "setp.eq.s32 q, %4, 3;\n\t@q mov.f64 %3, %5;\n\t}"
        : "+d"(p[0]), "+d"(p[1]), "+d"(p[2]), "+d"(p[3]) : "r"(key), "d"(v));
}
// two arrays under the same slot (a down move: the new bottom cell's event level and its out-of-band score)
__device__ __forceinline__ void put4x2(double (&p)[4], double (&q)[4], int lane, int s, double vp, double vq) {
    const int key = (lane == (s >> 2)) ? (s & 3) : -1;
    asm("{\n\t.reg .pred q;\n\t"
        "setp.eq.s32 q, %8, 0;\n\t@q mov.f64 %0, %9;\n\t@q mov.f64 %4, %10;\n\t"
        "setp.eq.s32 q, %8, 1;\n\t@q mov.f64 %1, %9;\n\t@q mov.f64 %5, %10;\n\t"
        "setp.eq.s32 q, %8, 2;\n\t@q mov.f64 %2, %9;\n\t@q mov.f64 %6, %10;\n\t"
        "setp.eq.s32 q, %8, 3;\n\t@q mov.f64 %3, %9;\n\t@q mov.f64 %7, %10;\n\t}"
        : "+d"(p[0]), "+d"(p[1]), "+d"(p[2]), "+d"(p[3]), "+d"(q[0]), "+d"(q[1]), "+d"(q[2]), "+d"(q[3])
        : "r"(key), "d"(vp), "d"(vq));
}

// One band b.  P1 = band b-1, P2 = band b-2 (overwritten with band b).
//
// Out-of-band neighbours must read as -inf (event_handling.cpp:283-294).  With event-indexed slots the only
// out-of-band slots a band can read are the one just above its top cell (as `up`, after a right move; as `diag`
// after two) and the one its new bottom cell enters (as `left`, after a down move; as `diag` after two).  Both
// belong to band b-1's array and are set to the sentinel here, once per band, which also covers the diagonal
// reads of band b+1.  Away from the ends of the read (kSteady) every one of the 100 cells is inside the event and
// k-mer ranges, so nothing else needs masking: the 28 slots outside the band compute garbage nobody reads.
template <bool kSteady>
__device__ __forceinline__ void dp_cells(DpWarp &w, double (&P1)[4], double (&P2)[4], const int b) {
    const int lane = w.lane;
    const DpConst &c = w.c;
    // slot-1 edges of the two previous bands
    const double e1 = shfl_d(P1[3], (lane + 31) & 31);
    const double e2 = shfl_d(P2[3], (lane + 31) & 31);
    // emissions: a = float((x - mu) / 0.14) via the guarded reciprocal multiply
    double q[4];
    bool near = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        q[j] = dMul(dSub(w.xe[j], w.mk[j]), c.inv_sigma);
        near |= (((unsigned)__double2loint(q[j]) & 0x1FFFFFFFu) - 0x0FFFFFF8u) <= 16u;
    }
    if (__any_sync(FULL, near)) {
#pragma unroll
        for (int j = 0; j < 4; j++) q[j] = dDiv(dSub(w.xe[j], w.mk[j]), 0.14);
    }
    uint32_t tb = 0;
    if (kSteady) {
#define DP_CELL(j)                                                                                                  \
        {                                                                                                           \
            uint32_t from;                                                                                          \
            const double m = cell_update<(DP_XU >> (4 * (j))) & 3>((j) ? P2[(j) ? (j) - 1 : 0] : e2,                \
                                                                   (j) ? P1[(j) ? (j) - 1 : 0] : e1, P1[j], q[j], c, from); \
            P2[j] = m;                                                                                              \
            tb |= from << (2 * (j));                                                                                \
        }
        DP_CELL(3) DP_CELL(2) DP_CELL(1) DP_CELL(0)
#undef DP_CELL
        w.fills += DNB_BW;
    } else {
        // fill range (event_handling.cpp:269-278) and trim cell (:256-265), as band offsets o = ll_e - event
        const int lo = max(max(-w.ll_k, w.ll_e - (w.E - 1)), 0);
        const int hi = min(min(w.K - w.ll_k, w.ll_e + 1), DNB_BW);
        const unsigned span = hi > lo ? (unsigned)(hi - lo) : 0u;
        const int o_trim = -1 - w.ll_k;
        const int ev_trim = w.ll_e - o_trim;
        const bool trim_ok = o_trim >= 0 && o_trim < DNB_BW && ev_trim >= 0 && ev_trim < w.E;
        const double trim_val = rn24(dMul(c.lp_trim, (double)((uint32_t)ev_trim + 1u)));   // :260
        w.fills += span;
#pragma unroll
        for (int j = 3; j >= 0; j--) {
            const int o = (w.ll_e - (lane * 4 + j)) & 127;
            uint32_t from;
            double m = cell_update<DP_XU & 3>(j ? P2[j ? j - 1 : 0] : e2, j ? P1[j ? j - 1 : 0] : e1, P1[j], q[j], c, from);
            const bool valid = (unsigned)(o - lo) < span;
            const bool trim = (o == o_trim) && trim_ok;
            m = valid ? m : (trim ? trim_val : NEG_SENT);
            from = valid ? from : (trim ? (uint32_t)DNB_FROM_U : 0u);
            P2[j] = m;
            tb |= from << (2 * j);
        }
    }
    // publish the band by event slot: the next move decision and the end-cell scan read single cells from it
    __syncwarp();
    *reinterpret_cast<double2 *>(w.sc + 4 * lane) = make_double2(P2[0], P2[1]);
    *reinterpret_cast<double2 *>(w.sc + 4 * lane + 2) = make_double2(P2[2], P2[3]);
    __syncwarp();
    w.v_ll = w.sc[w.ll_e & 127];
    w.v_ur = w.sc[(w.ll_e - (DNB_BW - 1)) & 127];
    w.rows[(size_t)b * DNB_TRACE_ROW + lane] = (uint8_t)tb;
    if ((b & 31) == 31) {
        if (lane == 0) { w.moves[b >> 5] = w.mvword; w.rcum[b >> 5] = w.rights; }
        w.rights += __popc(w.mvword);
        w.mvword = 0;
    }
    // end-cell candidate of this band: (event b-K-1, last k-mer) (event_handling.cpp:329-340); strict '>' in
    // ascending event order keeps the first maximum
    {
#if DP_LEAN
        const int e = w.e_cand++;                                   // b - K - 1, carried instead of recomputed
        if ((unsigned)e < (unsigned)w.E && (unsigned)(w.ll_e - e) < (unsigned)DNB_BW) {
#else
        const int e = b - w.K - 1;
        const int o = w.ll_e - e;
        if (e >= 0 && e < w.E && o >= 0 && o < DNB_BW) {
#endif
            const double val = w.sc[e & 127];
            const double s = rn24(dAdd(val, dMul((double)(unsigned long long)(w.E - e), c.lp_trim)));
            if (s > w.best_s) { w.best_s = s; w.best_e = e; w.best_lle = w.ll_e; }
        }
    }
}

__device__ __forceinline__ void dp_band(DpWarp &w, double (&P1)[4], double (&P2)[4], const int b) {
    const int lane = w.lane;
    // Suzuki's rule on cells 0 and 99 of band b-1 (event_handling.cpp:237-253)
    const bool right = (w.v_ll == NEG_SENT && w.v_ur == NEG_SENT) ? ((b & 1) != 0) : (w.v_ll < w.v_ur);
    // the k-mer level conveyor advances one slot every band
    {
        const double in = shfl_d(w.mk[3], (lane + 31) & 31);
        w.mk[3] = w.mk[2]; w.mk[2] = w.mk[1]; w.mk[1] = w.mk[0]; w.mk[0] = in;
    }
    if (right) {
        w.ll_k++;
        const int need = w.ll_k + DNB_BW - 1;
        if (need - w.mbase == 32) {
            w.mbuf = w.mnxt; w.mbase += 32;
            w.mnxt = (w.mbase + 32 + lane < w.K) ? w.mu[w.mbase + 32 + lane] : 0.0;
        }
        const double fresh = shfl_d(w.mbuf, need - w.mbase);
        put4(w.mk, lane, (w.ll_e - (DNB_BW - 1)) & 127, fresh);       // the new top cell's k-mer
        put4(P1, lane, (w.ll_e - DNB_BW) & 127, NEG_SENT);            // above the top: out of band b-1
        // (a warp-uniform 4-way branch on the slot's register index instead of these per-register selects was measured:
        // 21 instead of 58 instructions on this path, but 2 % SLOWER overall -- the branches cost more than the selects)
        w.mvword |= 1u << (b & 31);
    } else {
        w.ll_e++;
        const int need = w.ll_e;
        if (need - w.xbase == 32) {
            w.xbuf = w.xnxt; w.xbase += 32;
            w.xnxt = (w.xbase + 32 + lane < w.E) ? w.x[w.xbase + 32 + lane] : 0.0;
        }
        const double fresh = shfl_d(w.xbuf, need - w.xbase);
        // the new bottom cell: its event level enters, and its slot was out of band b-1
        put4x2(w.xe, P1, lane, w.ll_e & 127, fresh, NEG_SENT);
    }
#if DP_LEAN
    // 0 <= ll_k <= K - 100 and 99 <= ll_e <= E - 1 as two unsigned compares (the limits are 0 for reads too short to
    // ever hold a whole band, which makes the test false)
    const bool steady = (unsigned)w.ll_k < w.k_lim && (unsigned)(w.ll_e - (DNB_BW - 1)) < w.e_lim;
#else
    const bool steady = w.ll_k >= 0 && w.ll_e >= DNB_BW - 1 && w.ll_e <= w.E - 1 && w.ll_k + DNB_BW <= w.K;
#endif
    if (steady) dp_cells<true>(w, P1, P2, b);
    else dp_cells<false>(w, P1, P2, b);
}

__device__ __forceinline__ DpEnd dp_fill_warp(const DnbBatchView &v, const DnbDpArgs &a, uint32_t r, int lane, double *sc) {
    DpWarp w;
    w.lane = lane;
    w.E = (int)v.n_events[r];
    w.K = (int)(v.q_off[r + 1] - v.q_off[r]) - DNB_K + 1;
    w.x = a.x_e + v.ev_off[r];
    w.mu = a.mu_q + v.q_off[r];
    w.c.lp_skip = a.lp[4 * r + 0]; w.c.lp_stay = a.lp[4 * r + 1]; w.c.lp_step = a.lp[4 * r + 2]; w.c.lp_trim = a.lp[4 * r + 3];
    w.c.emit_const = a.emit_const; w.c.inv_sigma = a.inv_sigma;
    w.rows = a.trace + a.band_off[r] * DNB_TRACE_ROW;
    w.moves = a.moves + (a.band_off[r] >> 5) + r;
    w.rcum = a.rcum + (a.band_off[r] >> 5) + r;
    w.sc = sc;
    const int E = w.E, K = w.K;
    const int n_bands = E + K + 2;

    // ---- bands 0 and 1 (event_handling.cpp:213-228): A = band 0, B = band 1 ----
    double A[4], B[4];
    w.ll_e = DNB_BW / 2; w.ll_k = -1 - DNB_BW / 2;     // lower-left of band 1 = move_down(band 0)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int s = lane * 4 + j;
        A[j] = (s == 127) ? 0.0 : NEG_SENT;                        // bands[0][50] = 0: event -1 -> slot 127
        B[j] = (s == 0) ? rn24(w.c.lp_trim) : NEG_SENT;            // bands[1][50] = lp_trim: event 0 -> slot 0
        const int o = (w.ll_e - s) & 127;
        const int e = w.ll_e - o, km = w.ll_k + o;
        w.xe[j] = (e >= 0 && e < E) ? w.x[e] : 0.0;
        w.mk[j] = (km >= 0 && km < K) ? w.mu[km] : 0.0;
    }
    w.rows[lane] = 0;                                                            // band 0: no trace
    w.rows[DNB_TRACE_ROW + lane] = (lane == 0) ? (uint8_t)DNB_FROM_U : 0;       // trace[1][50] = FROM_U (slot 0)
    // cells 0 and 99 of band 1: offsets 0 and 99 are events 50 and -49 -> both -inf (only offset 50 is set)
    w.v_ll = NEG_SENT; w.v_ur = NEG_SENT;

    w.xbase = w.ll_e + 1; w.mbase = w.ll_k + DNB_BW;
    w.xbuf = (w.xbase + lane < E) ? w.x[w.xbase + lane] : 0.0;
    w.xnxt = (w.xbase + 32 + lane < E) ? w.x[w.xbase + 32 + lane] : 0.0;
    w.mbuf = (w.mbase + lane < K) ? w.mu[w.mbase + lane] : 0.0;
    w.mnxt = (w.mbase + 32 + lane < K) ? w.mu[w.mbase + 32 + lane] : 0.0;

    w.best_s = NEG_SENT; w.best_e = 0x7fffffff; w.best_lle = 0;
    w.fills = 0; w.mvword = 0; w.rights = 0;
    w.e_cand = 2 - K - 1;
    w.k_lim = (unsigned)max(K - (DNB_BW - 1), 0); w.e_lim = (unsigned)max(E - (DNB_BW - 1), 0);

    int b = 2;
    for (; b + 1 < n_bands; b += 2) {
        dp_band(w, B, A, b);          // band b   : P1 = B (b-1), P2 = A (b-2) -> A becomes band b
        dp_band(w, A, B, b + 1);      // band b+1 : P1 = A,       P2 = B       -> B becomes band b+1
    }
    if (b < n_bands) { dp_band(w, B, A, b); b++; }
    if (lane == 0 && (b & 31) != 0) { w.moves[b >> 5] = w.mvword; w.rcum[b >> 5] = w.rights; }
    if (lane == 0) atomicAdd(a.cells, w.fills);
    DpEnd end;
    if (w.best_e == 0x7fffffff) { end.event = -1; end.ll_event = 0; end.score = -INFINITY; }
    else { end.event = w.best_e; end.ll_event = w.best_lle; end.score = (float)w.best_s; }
    return end;
}

// ---------------------------------------------------------------------------------------------------------------
// Backtrace + QC of one read by one warp (event_handling.cpp:347-443).
//
// The walk itself is a serial pointer chase, but it only needs the 2-bit trace codes: lane 0 follows them through
// a 64-band window of the trace staged in shared memory (<= 32 steps per round, nothing else on its dependency
// chain).  Everything that hangs off the path is done by all lanes afterwards, one step per lane: the (event,
// k-mer) coordinates come from prefix counts of the codes (D: -1,-1  U: -1,0  L: 0,-1), then the alignment pairs,
// the emission of every aligned event, and for every diagonal step the k-mer's cleaned signal (mean of the event
// means buffered since the previous diagonal step, summed in push order) and its reference rank.  Only two
// quantities depend on the walk order through floating-point rounding or state, and both stay with lane 0:
// sum_emission (float emissions added into a double in walk order; added one round late, from shared memory, in
// the shadow of the next round's pointer chase) and the running gap length.
// ---------------------------------------------------------------------------------------------------------------
#define BT_ROWS 64
#define BT_STEPS 32

struct BtSmem {
    __align__(16) uint8_t win[BT_ROWS * DNB_TRACE_ROW];
    double lp[BT_STEPS];      // the round's emissions (float values, widened by the lanes that computed them)
};

__device__ __forceinline__ void backtrace_warp(const DnbBatchView &v, const DnbBtArgs &a, uint32_t r, int lane,
                                               int end_event, BtSmem &sm) {
    const int K = (int)(v.q_off[r + 1] - v.q_off[r]) - DNB_K + 1;
    const uint32_t Kref = (uint32_t)((v.r_off[r + 1] - v.r_off[r]) - DNB_K + 1);
    const double *__restrict__ x = a.dp.x_e + v.ev_off[r];
    const double *__restrict__ mu = a.dp.mu_q + v.q_off[r];
    const float *__restrict__ evm = v.ev_mean + v.ev_off[r];
    const int32_t *__restrict__ q2r = v.q2r + v.q_off[r];
    const uint32_t *__restrict__ rr = a.rank_ref + v.r_off[r];
    const uint8_t *rows = a.dp.trace + a.dp.band_off[r] * DNB_TRACE_ROW;
    const uint32_t *moves = a.dp.moves + (a.dp.band_off[r] >> 5) + r;
    const uint32_t *rcum = a.dp.rcum + (a.dp.band_off[r] >> 5) + r;
    uint2 *pairs = reinterpret_cast<uint2 *>(a.al_pairs_rev) + a.al_off[r];
    double *cls = a.cl_signal + a.cl_off[r];
    uint32_t *clr = a.cl_rank + a.cl_off[r];
    const double emit_const = a.dp.emit_const;
    const unsigned lt = (1u << lane) - 1u;

    int e = end_event, k = K - 1;          // warp-uniform position of the walk
    int e_prev_d = e + 1;                  // event of the previous diagonal step (+1 before the first: empty buffer)
    uint32_t na = 0, nc = 0;
    int last_k = -1;
    bool bad = false;
    // lane 0 only
    double sum_em = 0.0;
    int gap = 0, max_gap = 0, n_prev = 0;

    while (k >= 0 && e >= 0) {
        const int hi = e + k + 2;
        const int lo = max(hi - (BT_ROWS - 1), 0);
        {   // stage rows [lo, hi]
            const uint4 *src = reinterpret_cast<const uint4 *>(rows + (size_t)lo * DNB_TRACE_ROW);
            uint4 *dst = reinterpret_cast<uint4 *>(sm.win);
            const int n16 = (hi - lo + 1) * (DNB_TRACE_ROW / 16);
#pragma unroll
            for (int i = 0; i < (BT_ROWS * DNB_TRACE_ROW / 16) / 32; i++) {
                const int q = lane + 32 * i;
                if (q < n16) dst[q] = src[q];
            }
        }
        __syncwarp();
        // The chase.  ncu showed it to be ~75 % of the backtrace's instructions (38 per step, all on one lane), so the
        // common round -- 32 steps that can neither reach the start of the read nor wrap the 128-slot ring -- is a
        // fixed, unrolled loop of ~13 instructions per step: one shared-memory word, a 2-bit extract, the cell index
        // moved by a packed per-code delta (D: two rows up and one slot back = 257 cells, U: 129, L: 128); the codes
        // stay in two registers and are broadcast, the previous round's emissions are added in the loop's shadow.
        int ns = 0;
        uint32_t c_lo = 0, c_hi = 0;           // 2-bit codes of steps 0-15 / 16-31
        const bool fast = (e & 127) >= BT_STEPS && e >= BT_STEPS && k >= BT_STEPS;
        if (lane == 0) {
            if (fast) {
                const uint32_t *w32 = reinterpret_cast<const uint32_t *>(sm.win);
                uint32_t c = (uint32_t)(hi - lo) * 128u + (uint32_t)(e & 127);
#pragma unroll
                for (int t = 0; t < BT_STEPS; t++) {
                    const uint32_t code = (w32[c >> 4] >> ((c & 15u) * 2u)) & 3u;
                    if (t < 16) c_lo |= code << (2 * t); else c_hi |= code << (2 * (t - 16));
                    c -= (0x2010301u >> (9u * code)) & 0x1ffu;
                    if (t < n_prev) sum_em = dAdd(sum_em, sm.lp[t]);              // previous round's emissions, in order
                }
                ns = BT_STEPS;
            } else {
                int ee = e, kk = k;
                while (ns < BT_STEPS && kk >= 0 && ee >= 0) {
                    const int bb = ee + kk + 2;
                    const uint32_t code = (sm.win[(bb - lo) * DNB_TRACE_ROW + ((ee & 127) >> 2)] >> (2 * (ee & 3))) & 3u;
                    if (ns < 16) c_lo |= code << (2 * ns); else c_hi |= code << (2 * (ns - 16));
                    if (ns < n_prev) sum_em = dAdd(sum_em, sm.lp[ns]);
                    ee -= (code != DNB_FROM_L);
                    kk -= (code != DNB_FROM_U);
                    ns++;
                }
                for (int t = ns; t < n_prev; t++) sum_em = dAdd(sum_em, sm.lp[t]);
            }
            n_prev = ns;
        }
        ns = __shfl_sync(FULL, ns, 0);
        c_lo = __shfl_sync(FULL, c_lo, 0);
        c_hi = __shfl_sync(FULL, c_hi, 0);
        __syncwarp();

        // ---- one step per lane ----
        const bool act = lane < ns;
        const uint32_t code = act ? (((lane < 16 ? c_lo : c_hi) >> (2 * (lane & 15))) & 3u) : 3u;
        {   // running gap length (consecutive L moves): L moves are rare, so normally one compare per round
            const unsigned m_l = __ballot_sync(FULL, act && code == DNB_FROM_L);
            if (lane == 0 && ns > 0) {
                if (m_l == 0) gap = 0;
                else
                    for (int t = 0; t < ns; t++) {
                        gap = ((m_l >> t) & 1u) ? gap + 1 : 0;
                        max_gap = max(max_gap, gap);
                    }
            }
        }
        const unsigned m_e = __ballot_sync(FULL, act && code != DNB_FROM_L);
        const unsigned m_k = __ballot_sync(FULL, act && code != DNB_FROM_U);
        const unsigned m_d = __ballot_sync(FULL, act && code == DNB_FROM_D);
        const int ei = e - __popc(m_e & lt), ki = k - __popc(m_k & lt);
        if (act) {
            // the cell must lie inside its band (the reference would index trace[][] out of range otherwise)
            const int bi = ei + ki + 2;
            const uint32_t w = moves[bi >> 5];
            const int rights = (int)rcum[bi >> 5] + __popc(w & ((2u << (bi & 31)) - 1u));
            const int off = DNB_BW / 2 + (bi - 1) - rights - ei;
            if (off < 0 || off >= DNB_BW) bad = true;
            pairs[na + lane] = make_uint2((uint32_t)ei, (uint32_t)ki);                         // :359
            sm.lp[lane] = (double)emission_static(x[ei], mu[ki], emit_const);                  // :363
        }
        // previous diagonal step: inside this round or carried over
        const unsigned pm = m_d & lt;
        const int e_round = __shfl_sync(FULL, ei, pm ? 31 - __clz(pm) : 0);
        const int e_p = pm ? e_round : e_prev_d;
        const bool is_d = act && code == DNB_FROM_D;
        int32_t qr = -1;
        if (is_d) qr = q2r[ki];
        const bool emit = is_d && qr >= 0 && (uint32_t)qr < Kref;                              // :386-393
        const unsigned m_out = __ballot_sync(FULL, emit);
        if (emit) {
            double tot = 0.0;
            for (int j = e_p - 1; j >= ei; j--) tot = dAdd(tot, (double)evm[j]);               // push order
            const uint32_t idx = nc + __popc(m_out & lt);
            cls[idx] = dDiv(tot, (double)(uint32_t)(e_p - ei));                                // vectorMean, common.h:184
            clr[idx] = rr[qr];
        }
        nc += __popc(m_out);
        if (m_d) e_prev_d = __shfl_sync(FULL, ei, 31 - __clz(m_d));
        if (ns > 0) last_k = __shfl_sync(FULL, ki, ns - 1);
        na += ns;
        e -= __popc(m_e);
        k -= __popc(m_k);
        if (__any_sync(FULL, bad)) { bad = true; break; }
        __syncwarp();
    }
    __syncwarp();
    if (lane == 0) {
        if (bad) {
            v.status[r] = DNB_READ_UNDEFINED;
            a.n_align[r] = 0; a.n_cleaned[r] = 0; a.avg_log_emission[r] = 0.0; a.spanned[r] = 0; a.max_gap[r] = 0;
            return;
        }
        for (int t = 0; t < n_prev; t++) sum_em = dAdd(sum_em, sm.lp[t]);
        const double avg = dDiv(sum_em, (double)na);                                       // :420 (n_aligned counts steps)
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

// kMode 0: fill + backtrace (production)   1: fill only   2: backtrace only (the split pair is for profiling the two
// phases as separate launches; same device code)
#ifndef BT_MIN_BLOCKS
#define BT_MIN_BLOCKS 8      // the backtrace-only launch is a latency machine: cap its registers for 32 resident warps per SM
#endif
template <int kMode>
__global__ void __launch_bounds__(DP_WARPS * 32, kMode == 2 ? BT_MIN_BLOCKS : DP_MIN_BLOCKS) align_kernel(DnbBatchView v, DnbBtArgs a) {
    __shared__ BtSmem sm[DP_WARPS];      // the band fill's 1 KB slot copy aliases the backtrace window (used after it)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * DP_WARPS + w;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    if (v.status[r] != 0) {
        if (lane == 0 && kMode != 2) { a.dp.end_event[r] = -1; a.dp.end_ll_event[r] = 0; a.dp.end_score[r] = -INFINITY; }
        if (lane == 0 && kMode != 1) {
            a.n_align[r] = 0; a.n_cleaned[r] = 0; a.avg_log_emission[r] = 0.0; a.spanned[r] = 0; a.max_gap[r] = 0;
        }
        return;
    }
    int end_event;
    if (kMode != 2) {
        const long long t0 = clock64();
        const DpEnd end = dp_fill_warp(v, a.dp, r, lane, reinterpret_cast<double *>(sm[w].win));
        if (lane == 0) {
            a.dp.end_event[r] = end.event; a.dp.end_ll_event[r] = end.ll_event; a.dp.end_score[r] = end.score;
            if (end.event < 0) v.status[r] = DNB_READ_UNDEFINED;   // the reference would backtrace from an out-of-band cell
            atomicAdd(&a.phase_cycles[0], (unsigned long long)(clock64() - t0));
        }
        end_event = end.event;
        __syncwarp();      // the trace rows written above are read back by other lanes below
    } else {
        end_event = a.dp.end_event[r];
    }
    if (kMode != 1) {
        if (end_event < 0) {
            if (lane == 0) { a.n_align[r] = 0; a.n_cleaned[r] = 0; a.avg_log_emission[r] = 0.0; a.spanned[r] = 0; a.max_gap[r] = 0; }
            return;
        }
        const long long t1 = clock64();
        backtrace_warp(v, a, r, lane, end_event, sm[w]);
        if (lane == 0) atomicAdd(&a.phase_cycles[1], (unsigned long long)(clock64() - t1));
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

__global__ void compact_events_kernel(DnbBatchView v, const uint64_t *dense_off, uint32_t *out_start, float *out_mean) {
    const uint32_t r = blockIdx.x;
    const uint32_t n = (uint32_t)(dense_off[r + 1] - dense_off[r]);
    if (n == 0) return;
    const uint32_t *ss = v.ev_start + v.ev_off[r] + r;
    const float *ms = v.ev_mean + v.ev_off[r];
    uint32_t *ds = out_start + dense_off[r] + r;
    float *dm = out_mean + dense_off[r];
    for (uint32_t i = threadIdx.x; i <= n; i += blockDim.x) {
        ds[i] = ss[i];
        if (i < n) dm[i] = ms[i];
    }
}

}  // namespace

void dnb_launch_compact_events(const DnbBatchView &v, const uint64_t *dense_off, uint32_t *out_start, float *out_mean,
                               cudaStream_t s) {
    if (v.n_reads == 0) return;
    compact_events_kernel<<<v.n_reads, 256, 0, s>>>(v, dense_off, out_start, out_mean);
}

void dnb_launch_align(const DnbBatchView &v, const DnbBtArgs &a, int mode, cudaStream_t s) {
    if (v.n_reads == 0) return;
    const unsigned grid = (v.n_reads + DP_WARPS - 1) / DP_WARPS;
    if (mode == 0) align_kernel<0><<<grid, DP_WARPS * 32, 0, s>>>(v, a);
    else if (mode == 1) align_kernel<1><<<grid, DP_WARPS * 32, 0, s>>>(v, a);
    else align_kernel<2><<<grid, DP_WARPS * 32, 0, s>>>(v, a);
}

void dnb_launch_compact_alignment(const DnbBatchView &v, const uint64_t *al_off, const uint32_t *al_pairs_rev,
                                  const uint32_t *n_align, const uint64_t *out_off, uint32_t *out_pairs,
                                  cudaStream_t s) {
    if (v.n_reads == 0) return;
    compact_alignment_kernel<<<v.n_reads, 128, 0, s>>>(v, al_off, al_pairs_rev, n_align, out_off, out_pairs);
}
