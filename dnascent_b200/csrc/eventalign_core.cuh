// eventalign_core.cuh -- the pieces of the windowed Viterbi re-alignment shared by the read-serial kernel
// (eventalign.cu) and the window-parallel one (eventalign_wp.cu): alphabet helpers, the exact deletion-chain scan,
// builtinViterbi's forward pass (src/alignment.cpp:193-440).  See eventalign.cu's header for the arithmetic notes.
#pragma once
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


}  // namespace
