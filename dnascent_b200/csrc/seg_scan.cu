// seg_scan.cu -- the (sum, sumsq) checkpoints of the segmentation WITHOUT a per-sample serial chain (experimental).
//
// STATUS: written at the end of round 1 after the round's GPU budget was spent: it compiles for sm_100a and mirrors, step
// by step, a CPU prototype that is bit-identical to the sequential loop (tests/helpers/proto_exact_prefix.py, block_chain;
// tests/test_host_logic.py), but THIS KERNEL HAS NEVER RUN ON A GPU.  It replaces seg.cu's checkpoint kernel only when
// DNB_SEG_PARITY_SCAN=1 is set; every bit-exact GPU test of the segmentation is its acceptance test.
//
// What it replaces: seg_checkpoint_kernel (seg.cu), i.e. scrappie's compute_sum_sumsq
// (/root/reference/src/scrappie/event_detection.c:35-48) evaluated up to every 64th sample:
//     sum[i+1] = fl(sum[i] + x[i]),   sumsq[i+1] = fl(sumsq[i] + x[i]*x[i])
// Both round at every step, so the values depend on the order; the checkpoint kernel therefore walks each read with
// one lane (~31 cycles per sample: 65 ms for a 4*10^6-sample read, the critical path of a short batch).
//
// How.  While the running sum s stays inside one binade (ulp u), fl(s + a) = (S + q + c) u with s = S u, a = q u + r,
// 0 <= r < u, and c = [r > u/2], or (S + q) mod 2 on a tie: the step is S -> S + d[S mod 2] for a pair of integers
// (d[0], d[1]) that depends on a and u only, and such maps compose associatively
//     (g o f).d[p] = f.d[p] + g.d[p xor (f.d[p] & 1)].
// One warp per read:
//   A  every lane takes 64-sample blocks: sum x, sum |x|, sum x^2 of the block (plain doubles: only used to PREDICT)
//   B  warp scan of the block sums -> predicted running sums at the block starts -> predicted binade of each block
//   C  every lane folds its blocks' 64 samples into one map per chain under the predicted binade (integer arithmetic)
//   D  lane 0 walks the blocks: if the ACTUAL running sum is in the predicted binade and sum |a| of the block cannot
//      take it out, one integer step replaces 64 roundings; otherwise the block is summed literally.  Correctness never
//      rests on the prediction: a wrong one only costs the literal loop.
// The serial chain shrinks from N to N/64 steps (plus ~1 % literal blocks on POD5-like data).
#include <climits>
#include <vector>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define PS_BLOCK DNB_SEG_CK          // 64 samples: one checkpoint per block
#define PS_FULL 0xffffffffu
#define PS_NO_MAP INT_MIN

namespace {

template <bool kI16>
struct PsReader {
    const float *f32;
    const int16_t *i16;
    float dac_off, dac_scl;
    __device__ __forceinline__ float at(uint64_t idx) const {
        if (kI16) return fMul(fAdd((float)i16[idx], dac_off), dac_scl);      // src/pod5.cpp:60
        return f32[idx];
    }
};

__device__ __forceinline__ int binade_of(double s) {                         // s > 0, normal: s in [2^e, 2^(e+1))
    return ((__double2hiint(s) >> 20) & 0x7ff) - 1023;
}
__device__ __forceinline__ double pow2(int e) {                              // 2^e for -1022 <= e <= 1023
    return __hiloint2double((e + 1023) << 20, 0);
}

// fold the addend `a` into the map (d0, d1) of running sums with ulp 2^sh; false = the addend does not fit the scheme
__device__ __forceinline__ bool fold(double a, int sh, long long &d0, long long &d1) {
    const double scaled = dMul(a, pow2(-sh));                                // exact: a power of two, no under/overflow here
    if (!(fabs(scaled) < 9007199254740992.0)) return false;                  // |a| >= 2^53 ulps: leaves the binade anyway
    const double q = floor(scaled);                                          // floor also for negative addends: r in [0, 1)
    const double r = dSub(scaled, q);                                        // exact
    const long long qi = (long long)q;
    long long g0, g1;
    if (r == 0.5) {                                                          // tie: to even
        const bool odd = (qi & 1) != 0;
        g0 = odd ? qi + 1 : qi;
        g1 = odd ? qi : qi + 1;
    } else {
        g0 = g1 = qi + (r > 0.5 ? 1 : 0);
    }
    d0 += (d0 & 1) ? g1 : g0;                                                // this element after the block so far
    d1 += (d1 & 1) ? g0 : g1;
    return true;
}

struct PsScratch {
    double *bs, *ba, *bq;            // per block: sum x, sum |x|, sum x^2
    double *ps, *pq;                 // predicted running sums at the block start
    long long *s0, *s1, *q0, *q1;    // maps of the two chains
    int *es, *eq;                    // binade the map was built for, PS_NO_MAP = none
};

template <bool kI16>
__global__ void __launch_bounds__(128) seg_parity_scan_kernel(DnbBatchView v, DnbSegTiles t, PsScratch w) {
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    const uint32_t N = v.n_samples[r];
    const uint64_t base = v.raw_off[r];
    PsReader<kI16> rd{v.raw_f32, v.raw_i16, 0.f, 1.f};
    if (kI16) { rd.dac_off = v.dac_offset[r]; rd.dac_scl = v.dac_scale[r]; }
    const uint64_t ck = t.ck_off[r];
    const uint32_t nb = (N + PS_BLOCK - 1) / PS_BLOCK;

    // ---- A: block sums (and the precondition of the tile kernel's fast path, as the checkpoint kernel checks it) ----
    uint32_t out_of_range = 0;
    for (uint32_t b = lane; b < nb; b += 32) {
        const uint32_t j0 = b * PS_BLOCK, j1 = min(j0 + PS_BLOCK, N);
        double s = 0.0, a = 0.0, q = 0.0;
        for (uint32_t j = j0; j < j1; j++) {
            const float xf = rd.at(base + j);
            const double x = (double)xf;
            const uint32_t ax = __float_as_uint(xf) & 0x7fffffffu;
            out_of_range |= (ax - 0x35800000u >= 0x19000000u) && ax != 0u;
            s = dAdd(s, x);
            a = dAdd(a, fabs(x));
            q = dAdd(q, dMul(x, x));
        }
        w.bs[ck + b] = s; w.ba[ck + b] = a; w.bq[ck + b] = q;
    }
    out_of_range = __any_sync(PS_FULL, out_of_range != 0) ? 1u : 0u;
    __syncwarp();

    // ---- B: predicted running sums at the block starts (approximate: any order will do) ----
    {
        double carry_s = 0.0, carry_q = 0.0;
        for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
            const uint32_t b = b0 + lane;
            const double vs = b < nb ? w.bs[ck + b] : 0.0, vq = b < nb ? w.bq[ck + b] : 0.0;
            double is = vs, iq = vq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double us = __shfl_up_sync(PS_FULL, is, d), uq = __shfl_up_sync(PS_FULL, iq, d);
                if (lane >= d) { is += us; iq += uq; }
            }
            if (b < nb) { w.ps[ck + b] = carry_s + (is - vs); w.pq[ck + b] = carry_q + (iq - vq); }
            carry_s += __shfl_sync(PS_FULL, is, 31);
            carry_q += __shfl_sync(PS_FULL, iq, 31);
        }
    }
    __syncwarp();

    // ---- C: one map per block and chain, under the predicted binade ----
    for (uint32_t b = lane; b < nb; b += 32) {
        const uint32_t j0 = b * PS_BLOCK, j1 = min(j0 + PS_BLOCK, N);
        const double ps = w.ps[ck + b], pq = w.pq[ck + b];
        int es = PS_NO_MAP, eq = PS_NO_MAP;
        if (ps > 0.0 && ps < 1.0e300) { es = binade_of(ps); if (es < -900) es = PS_NO_MAP; }
        if (pq > 0.0 && pq < 1.0e300) { eq = binade_of(pq); if (eq < -900) eq = PS_NO_MAP; }
        long long s0 = 0, s1 = 0, q0 = 0, q1 = 0;
        if (es != PS_NO_MAP || eq != PS_NO_MAP) {
            for (uint32_t j = j0; j < j1; j++) {
                const double x = (double)rd.at(base + j);
                if (es != PS_NO_MAP && !fold(x, es - 52, s0, s1)) es = PS_NO_MAP;
                if (eq != PS_NO_MAP && !fold(dMul(x, x), eq - 52, q0, q1)) eq = PS_NO_MAP;
            }
        }
        w.s0[ck + b] = s0; w.s1[ck + b] = s1; w.q0[ck + b] = q0; w.q1[ck + b] = q1;
        w.es[ck + b] = es; w.eq[ck + b] = eq;
    }
    __syncwarp();
    __threadfence_block();

    // ---- D: the chain, one step per block; the block's parameters are loaded one block ahead of their use ----
    if (lane == 0) {
        struct Par { int es, eq; double ba, bq; long long s0, s1, q0, q1; };
        auto load = [&](uint32_t b) {
            Par p;
            p.es = w.es[ck + b]; p.eq = w.eq[ck + b]; p.ba = w.ba[ck + b]; p.bq = w.bq[ck + b];
            p.s0 = w.s0[ck + b]; p.s1 = w.s1[ck + b]; p.q0 = w.q0[ck + b]; p.q1 = w.q1[ck + b];
            return p;
        };
        double *cs = t.ck_sum + ck, *cq = t.ck_sq + ck;
        double s = 0.0, q = 0.0;
        Par nxt = {PS_NO_MAP, PS_NO_MAP, 0.0, 0.0, 0, 0, 0, 0};
        if (nb) nxt = load(0);
        for (uint32_t b = 0; b < nb; b++) {
            const Par cur = nxt;
            if (b + 1 < nb) nxt = load(b + 1);
            cs[b] = s; cq[b] = q;
            const uint32_t j0 = b * PS_BLOCK, j1 = min(j0 + PS_BLOCK, N);
            // sum chain: signed addends, so the block must not be able to leave the binade in either direction
            bool done = false;
            if (cur.es != PS_NO_MAP && s > 0.0 && binade_of(s) == cur.es) {
                const int e = cur.es;
                const double lo = pow2(e), hi = pow2(e + 1), span = dMul(cur.ba, 1.0000001);
                if (dSub(s, span) >= lo && dAdd(s, span) < hi) {
                    long long S = (long long)dMul(s, pow2(52 - e));                         // exact: s is a multiple of its ulp
                    S += (S & 1) ? cur.s1 : cur.s0;
                    if (S >= (1ll << 52) && S < (1ll << 53)) { s = dMul((double)S, pow2(e - 52)); done = true; }
                }
            }
            if (!done)
                for (uint32_t j = j0; j < j1; j++) s = dAdd(s, (double)rd.at(base + j));    // event_detection.c:45
            // sumsq chain: non-negative addends, the sum only grows
            done = false;
            if (cur.eq != PS_NO_MAP && q > 0.0 && binade_of(q) == cur.eq) {
                const int e = cur.eq;
                const double hi = pow2(e + 1), span = dMul(cur.bq, 1.0000001);
                if (dAdd(q, span) < hi) {
                    long long S = (long long)dMul(q, pow2(52 - e));
                    S += (S & 1) ? cur.q1 : cur.q0;
                    if (S >= (1ll << 52) && S < (1ll << 53)) { q = dMul((double)S, pow2(e - 52)); done = true; }
                }
            }
            if (!done)
                for (uint32_t j = j0; j < j1; j++) { const double x = (double)rd.at(base + j); q = dAdd(q, dMul(x, x)); }   // :46
        }
        t.tot_sum[r] = s;
        t.redo[r] = out_of_range;      // the stitch kernel ORs its own verdict into this
    }
}

}  // namespace

// Drop-in for the two seg_checkpoint_kernel launches of dnb_launch_segmentation_tiled (same outputs: ck_sum, ck_sq,
// tot_sum, redo).  The scratch (11 arrays of one entry per 64-sample block) comes from the caller's device cache.
static size_t ps_slots(const DnbSegTiles &t, uint32_t n_reads) {
    // checkpoint slots of the batch: ceil(N/64) + 1 per read <= 8 tiles' worth + 2
    return (size_t)t.n_tiles * (DNB_SEG_TILE / DNB_SEG_CK) + 2 * (size_t)n_reads + 64;
}
size_t dnb_seg_parity_scan_scratch_bytes(const DnbSegTiles &t, uint32_t n_reads) { return ps_slots(t, n_reads) * (9 * 8 + 2 * 4) + 256; }

cudaError_t dnb_launch_seg_parity_scan(const DnbBatchView &v, const DnbSegTiles &t, void *scratch, cudaStream_t s) {
    if (v.n_reads == 0) return cudaSuccess;
    if (!scratch) return cudaErrorInvalidValue;
    const size_t slots = ps_slots(t, v.n_reads);
    uint8_t *p = (uint8_t *)scratch;
    auto take = [&](size_t bytes) { void *q = p; p += bytes; return q; };
    PsScratch w;
    w.bs = (double *)take(slots * 8); w.ba = (double *)take(slots * 8); w.bq = (double *)take(slots * 8);
    w.ps = (double *)take(slots * 8); w.pq = (double *)take(slots * 8);
    w.s0 = (long long *)take(slots * 8); w.s1 = (long long *)take(slots * 8);
    w.q0 = (long long *)take(slots * 8); w.q1 = (long long *)take(slots * 8);
    w.es = (int *)take(slots * 4); w.eq = (int *)take(slots * 4);
    const unsigned grid = (v.n_reads + 3) / 4;                               // one warp per read, 4 warps per CTA
    if (v.raw_i16) seg_parity_scan_kernel<true><<<grid, 128, 0, s>>>(v, t, w);
    else seg_parity_scan_kernel<false><<<grid, 128, 0, s>>>(v, t, w);
    return cudaGetLastError();
}
