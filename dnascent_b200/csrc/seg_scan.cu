// seg_scan.cu -- the exact (sum, sumsq) checkpoints of the segmentation as a streaming warp scan.
//
// What it replaces: seg_checkpoint_kernel (seg.cu), i.e. scrappie's compute_sum_sumsq
// (/root/reference/src/scrappie/event_detection.c:35-48) evaluated up to every 64th sample:
//     sum[i+1] = fl(sum[i] + x[i]),   sumsq[i+1] = fl(sumsq[i] + x[i]*x[i])
// Both chains round, so in general their values depend on the order of the additions; the checkpoint kernel therefore
// walks each read with ONE lane (~31 cycles per sample: 100 ms of a 30k-read bench step, 60 ms for one 4*10^6-sample
// read).  Here one warp streams a read in chunks of 32 blocks x 64 samples, every lane working on its own block, and
// the values written are bit-identical to the sequential loop's:
//
//   sum     The samples are float32 values (src/pod5.cpp:60).  If q is the smallest power of two every sample is a
//           multiple of (2^-23 of the smallest nonzero |x|'s binade) and sum|x| < 2^53 q, EVERY partial sum of ANY
//           subset is exactly representable, so no addition rounds and the order does not matter: block sums + a
//           warp scan.  The condition is checked per read (it holds for any realistic signal: 49 of the 53 bits are
//           used by 10^7 samples of ~100 pA); a read that violates it is flagged for the literal serial kernel.
//   sumsq   x*x is exact (48 bits) but the running sum rounds at every step.  While the sum s = S u stays inside one
//           binade (ulp u), fl(s + a) = (S + g[S mod 2]) u where g[0] = RN-even(a/u) and g[1] is the same except on
//           an exact tie (then the other neighbour): the step is a two-state transducer on the parity of S, and
//           transducers compose associatively,  (g o f)[p] = f[p] + g[p xor (f[p] & 1)].
//           Every lane folds its block's 64 addends into one such map under the binade PREDICTED from the running
//           sum at the chunk start plus a plain prefix of the block sums (phase C), a warp scan composes the 32 maps
//           (phase D), and the result is accepted only if every block provably stayed inside its binade: start
//           state in [2^52, 2^53) ulps of the binade the map was built for and end state < 2^53 (addends are
//           non-negative, so the chain is monotone and the as-if-one-binade sum passes 2^53 exactly when the real
//           chain leaves the binade).  Otherwise -- the ~40 chunks per read in which the sum crosses a power of two,
//           and the first chunk -- the chunk is walked block by block, applying the maps that verify and adding the
//           64 samples of a block literally where they do not.  Correctness never rests on the prediction.
//
// Cost: ~25 instructions per sample, all lanes busy and loads coalesced into 16-byte vectors, against a 31-cycle
// dependent chain per sample on one lane.  Enabled by default (DNB_SEG_PARITY_SCAN=0 selects the checkpoint kernel);
// every bit-exact GPU test of the segmentation is its acceptance test, tests/test_segmentation_scan_gpu.py adds
// adversarial binade-crossing signals.  The algebra was first checked on the CPU (tests/helpers/proto_exact_prefix.py).
#include <climits>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define PS_BLOCK DNB_SEG_CK          // 64 samples: one checkpoint per block
#define PS_FULL 0xffffffffu
#define PS_NO_MAP INT_MIN
#define PS_TWO52 4503599627370496.0

namespace {

__device__ __forceinline__ int binade_of(double s) {                         // s > 0, normal: s in [2^e, 2^(e+1))
    return ((__double2hiint(s) >> 20) & 0x7ff) - 1023;
}
__device__ __forceinline__ double pow2(int e) {                              // 2^e for -1022 <= e <= 1023
    return __hiloint2double((e + 1023) << 20, 0);
}

// the samples [j0, j0 + 8) of a read as floats; `base` is the read's first element (32-element aligned)
template <bool kI16>
__device__ __forceinline__ void load8(const DnbBatchView &v, uint64_t base, uint32_t j0, float dac_off, float dac_scl, float (&x)[8]) {
    if (kI16) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(v.raw_i16 + base + j0));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            x[2 * i] = fMul(fAdd((float)(short)(w[i] & 0xffffu), dac_off), dac_scl);      // src/pod5.cpp:60
            x[2 * i + 1] = fMul(fAdd((float)(short)(w[i] >> 16), dac_off), dac_scl);
        }
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(v.raw_f32 + base + j0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(v.raw_f32 + base + j0 + 4));
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
}
template <bool kI16>
__device__ __forceinline__ float load1(const DnbBatchView &v, uint64_t idx, float dac_off, float dac_scl) {
    if (kI16) return fMul(fAdd((float)v.raw_i16[idx], dac_off), dac_scl);
    return v.raw_f32[idx];
}
// eight samples of a block, by vector where the padded read allows it
template <bool kI16>
__device__ __forceinline__ void load_group(const DnbBatchView &v, uint64_t base, uint32_t j, uint32_t N, uint32_t padded,
                                           float dac_off, float dac_scl, float (&x)[8]) {
    if (j + 8 <= padded) load8<kI16>(v, base, j, dac_off, dac_scl, x);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = (j + i < N) ? load1<kI16>(v, base + j + i, dac_off, dac_scl) : 0.0f;
    }
}

// one addend of the sumsq chain folded into the map (d0, d1) of a running sum with ulp 2^-k (inv_u = 2^k):
// g0 = RN-even(a / u) by the 2^52 trick, g1 differs only on an exact tie.  false: a/u >= 2^52, no map for this block.
__device__ __forceinline__ bool fold_sq(double a, double inv_u, double &d0, double &d1) {
    const double A = dMul(a, inv_u);                          // exact (power of two; no underflow for |x| >= 2^-20)
    if (!(A < PS_TWO52)) return false;
    const double g0 = dSub(dAdd(A, PS_TWO52), PS_TWO52);      // round half to even, integer valued
    const double diff = dSub(g0, A);                          // exact, |diff| <= 0.5
    if (fabs(diff) == 0.5) {                                  // exact tie (rare): the odd-parity state takes the other neighbour
        const double g1 = dSub(g0, dAdd(diff, diff));
        const bool odd0 = (__double2loint(dAdd(d0, PS_TWO52)) & 1) != 0;
        const bool odd1 = (__double2loint(dAdd(d1, PS_TWO52)) & 1) != 0;
        d0 = dAdd(d0, odd0 ? g1 : g0);
        d1 = dAdd(d1, odd1 ? g0 : g1);
    } else {
        d0 = dAdd(d0, g0);
        d1 = dAdd(d1, g0);
    }
    return true;
}

template <bool kI16>
__global__ void __launch_bounds__(128) seg_scan_kernel(DnbBatchView v, DnbSegTiles t) {
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    const uint32_t N = v.n_samples[r];
    const uint64_t base = v.raw_off[r];
    const uint32_t padded = (uint32_t)(v.raw_off[r + 1] - base);            // multiple of 32: vector loads stay inside
    float dac_off = 0.f, dac_scl = 1.f;
    if (kI16) { dac_off = v.dac_offset[r]; dac_scl = v.dac_scale[r]; }
    double *cs = t.ck_sum + t.ck_off[r], *cq = t.ck_sq + t.ck_off[r];
    const uint32_t nb = (N + PS_BLOCK - 1) / PS_BLOCK;

    double s_act = 0.0, q_act = 0.0;       // the chains at the chunk start (warp-uniform, the sequential loop's values)
    double abs_sum = 0.0;                  // this lane's share of sum |x| (bound for the exactness condition of `sum`)
    uint32_t min_exp = 0xffu;              // smallest biased exponent among this lane's nonzero samples
    uint32_t out_of_range = 0;             // precondition of tstat_fast: every sample is 0 or 2^-20 <= |x| < 2^30

    for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
        const uint32_t b = b0 + lane;
        const bool have = b < nb;
        const uint32_t j0 = b * PS_BLOCK;
        const uint32_t cnt = have ? min((uint32_t)PS_BLOCK, N - j0) : 0u;
        // ---- A: block sums (exact sum x; plain sum x^2 for the prediction) ----
        double S = 0.0, Q = 0.0;
        for (uint32_t g = 0; g < cnt; g += 8) {
            float x[8];
            load_group<kI16>(v, base, j0 + g, N, padded, dac_off, dac_scl, x);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (g + i < cnt) {
                    const uint32_t ax = __float_as_uint(x[i]) & 0x7fffffffu;
                    out_of_range |= (ax - 0x35800000u >= 0x19000000u) && ax != 0u;
                    if (ax != 0u) min_exp = min(min_exp, ax >> 23);
                    const double xd = (double)x[i];
                    S = dAdd(S, xd);
                    abs_sum = dAdd(abs_sum, fabs(xd));
                    Q = dAdd(Q, dMul(xd, xd));
                }
            }
        }
        // ---- B: exclusive prefixes inside the chunk ----
        double inS = S, inQ = Q;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double us = __shfl_up_sync(PS_FULL, inS, d), uq = __shfl_up_sync(PS_FULL, inQ, d);
            if (lane >= d) { inS = dAdd(inS, us); inQ = dAdd(inQ, uq); }
        }
        const double exS = dSub(inS, S);                                     // exact under the read's condition
        const double pq = dAdd(q_act, dSub(inQ, Q));                         // predicted sumsq at this block's start
        if (have) cs[b] = dAdd(s_act, exS);
        s_act = dAdd(s_act, __shfl_sync(PS_FULL, inS, 31));
        // ---- C: this block's transducer under the predicted binade ----
        int eq = PS_NO_MAP;
        double d0 = 0.0, d1 = 0.0;
        if (have && pq > 0.0 && pq < 1.0e300) {
            eq = binade_of(pq);
            if (eq < -900) eq = PS_NO_MAP;
        }
        if (eq != PS_NO_MAP) {
            const double inv_u = pow2(52 - eq);
            bool ok = true;
            for (uint32_t g = 0; g < cnt && ok; g += 8) {
                float x[8];
                load_group<kI16>(v, base, j0 + g, N, padded, dac_off, dac_scl, x);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (g + i < cnt) { const double xd = (double)x[i]; ok = ok && fold_sq(dMul(xd, xd), inv_u, d0, d1); }
            }
            if (!ok) eq = PS_NO_MAP;
        }
        const long long m0 = (long long)d0, m1 = (long long)d1;             // integers (below 2^53 wherever they get used)
        // ---- D: compose.  Fast path: every block of the chunk was folded for the binade the chain is in now ----
        const int e_act = (q_act > 0.0 && q_act < 1.0e300) ? binade_of(q_act) : PS_NO_MAP;
        const bool uniform = e_act != PS_NO_MAP && __all_sync(PS_FULL, !have || eq == e_act);
        bool done = false;
        if (uniform) {
            const long long S0 = (long long)dMul(q_act, pow2(52 - e_act));   // exact: q_act is a multiple of its ulp
            long long f0 = have ? m0 : 0, f1 = have ? m1 : 0;               // inclusive composition of lanes 0..lane
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long a0 = __shfl_up_sync(PS_FULL, f0, d), a1 = __shfl_up_sync(PS_FULL, f1, d);
                if (lane >= d) {                                             // (earlier blocks: a) then (these: f)
                    const long long h0 = a0 + ((a0 & 1) ? f1 : f0);
                    const long long h1 = a1 + ((a1 & 1) ? f0 : f1);
                    f0 = h0; f1 = h1;
                }
            }
            long long x0 = __shfl_up_sync(PS_FULL, f0, 1), x1 = __shfl_up_sync(PS_FULL, f1, 1);
            if (lane == 0) { x0 = 0; x1 = 0; }                               // exclusive: identity for the first block
            const long long Sb = S0 + ((S0 & 1) ? x1 : x0);                  // state at this block's start
            const long long Se = S0 + ((S0 & 1) ? f1 : f0);                  // ... and at its end
            const bool fine = !have || (m0 >= 0 && m1 >= 0 && m0 < (1ll << 52) && m1 < (1ll << 52) && Se < (1ll << 53));
            if (__all_sync(PS_FULL, fine)) {
                const double u = pow2(e_act - 52);
                if (have) cq[b] = dMul((double)Sb, u);
                const int last = (int)min(31u, nb - 1 - b0);
                q_act = dMul((double)__shfl_sync(PS_FULL, Se, last), u);
                done = true;
            }
        }
        if (!done) {
            // block by block, every lane computing the same (warp-uniform) chain
            const int n_here = (int)min(32u, nb - b0);
            for (int L = 0; L < n_here; L++) {
                const int eL = __shfl_sync(PS_FULL, eq, L);
                const long long a0 = __shfl_sync(PS_FULL, m0, L), a1 = __shfl_sync(PS_FULL, m1, L);
                if (lane == 0) cq[b0 + L] = q_act;
                bool applied = false;
                if (eL != PS_NO_MAP && a0 >= 0 && a1 >= 0 && a0 < (1ll << 52) && a1 < (1ll << 52) && q_act > 0.0 &&
                    q_act < 1.0e300 && binade_of(q_act) == eL) {
                    long long Sx = (long long)dMul(q_act, pow2(52 - eL));
                    Sx += (Sx & 1) ? a1 : a0;
                    if (Sx < (1ll << 53)) { q_act = dMul((double)Sx, pow2(eL - 52)); applied = true; }
                }
                if (!applied) {
                    const uint32_t k0 = (b0 + L) * PS_BLOCK, k1 = min(k0 + PS_BLOCK, N);
                    for (uint32_t j = k0; j < k1; j++) {
                        const double xd = (double)load1<kI16>(v, base + j, dac_off, dac_scl);
                        q_act = dAdd(q_act, dMul(xd, xd));                   // event_detection.c:46
                    }
                }
            }
        }
    }
    // ---- the read-level condition under which `sum` was exact in any order ----
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        abs_sum = dAdd(abs_sum, __shfl_xor_sync(PS_FULL, abs_sum, d));
        min_exp = min(min_exp, __shfl_xor_sync(PS_FULL, min_exp, d));
        out_of_range |= __shfl_xor_sync(PS_FULL, out_of_range, d);
    }
    if (lane == 0) {
        bool exact = true;
        if (min_exp != 0xffu) {
            // every sample is a multiple of 2^(min_exp - 127 - 23); all partial sums are below 1.0001 * sum|x|
            exact = min_exp >= 1u && dMul(abs_sum, 1.0001) < pow2((int)min_exp - 150 + 53);
        }
        t.tot_sum[r] = s_act;
        t.redo[r] = (out_of_range != 0u || !exact) ? 1u : 0u;     // the stitch kernel ORs its own verdict into this
    }
}

}  // namespace

// Drop-in for the seg_checkpoint_kernel launch of dnb_launch_segmentation_tiled (same outputs: ck_sum, ck_sq,
// tot_sum, redo); needs no scratch.
cudaError_t dnb_launch_seg_parity_scan(const DnbBatchView &v, const DnbSegTiles &t, cudaStream_t s) {
    if (v.n_reads == 0) return cudaSuccess;
    const unsigned grid = (v.n_reads + 3) / 4;                               // one warp per read, 4 warps per CTA
    if (v.raw_i16) seg_scan_kernel<true><<<grid, 128, 0, s>>>(v, t);
    else seg_scan_kernel<false><<<grid, 128, 0, s>>>(v, t);
    return cudaGetLastError();
}
