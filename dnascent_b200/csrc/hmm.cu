// hmm.cu -- analogue log-likelihood: batched forward algorithm of sequenceProbability, analogue and thymidine
// passes fused per site.
//
// Replaces (reference, paths relative to /root/reference):
//   sequenceProbability      src/detect.cpp:235-378   (called twice per T site by llAcrossRead, :546-547)
//   eexp/eln/lnSum/lnProd    src/probability.cpp:23-88  (NaN == log 0 convention, kept verbatim)
//   normalPDF                src/probability.cpp:145-148
//
// One thread per site: 2*window positions x {I, M, D} states in per-thread arrays, observations streamed.  The
// two passes differ only in the emission of the positions inside [BrdUStart, BrdUEnd] whose 9-mer contains a T
// (:317-323), so the unlabelled emissions are computed once and shared.  Tolerance for this path is 1e-4
// relative (BASELINE.json), so CUDA's log/exp are used.
#include <cmath>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define HMM_MAXW 16   // window <= 16 -> <= 32 positions (the reference uses windowLength = 12)

namespace {

__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000ll); }
__device__ __forceinline__ double eexp_d(double x) { return isnan(x) ? 0.0 : exp(x); }
// eln of a non-negative argument (the reference throws for negatives; callers here never produce them
// except through NaN emissions, which propagate as NaN)
__device__ __forceinline__ double eln_d(double x) { return x == 0.0 ? d_nan() : log(x); }
__device__ __forceinline__ double lnSum_d(double a, double b) {
    const bool na = isnan(a), nb = isnan(b);
    if (na || nb) return (na && nb) ? d_nan() : (na ? b : a);
    return a > b ? a + eln_d(1.0 + eexp_d(b - a)) : b + eln_d(1.0 + eexp_d(a - b));
}
__device__ __forceinline__ double lnProd_d(double a, double b) { return (isnan(a) || isnan(b)) ? d_nan() : a + b; }
__device__ __forceinline__ double normalPDF_d(double mu, double sigma, double x) {
    const double s2 = sigma * sigma;   // pow(sigma, 2.0) == sigma*sigma exactly
    const double dx = x - mu;
    return (1.0 / sqrt(2.0 * s2 * 3.14159265358979323846)) * exp(-(dx * dx) / (2.0 * s2));
}

struct Trans {
    double D2D, D2M, I2M, M2D, M2I, I2I, M2M_int, M2M_ext, l25, l5;
};

__device__ void forward_pass(const double *obs, uint32_t n_obs, double shift, double scale, const double *mu,
                             const double *sg, uint32_t n, const Trans &t, double *out) {
    double Ip[2 * HMM_MAXW], Mp[2 * HMM_MAXW], Dp[2 * HMM_MAXW], Ic[2 * HMM_MAXW], Mc[2 * HMM_MAXW], Dc[2 * HMM_MAXW];
    const double NaN = d_nan();
    for (uint32_t i = 0; i < n; i++) { Ip[i] = Mp[i] = NaN; Ic[i] = Mc[i] = Dc[i] = NaN; }
    double firstI_curr = NaN, firstI_prev = NaN, start_prev = 0.0;
    Dp[0] = lnProd_d(start_prev, t.l25);                       // detect.cpp:265
    for (uint32_t i = 1; i < n; i++) Dp[i] = lnProd_d(Dp[i - 1], t.D2D);
    for (uint32_t ti = 0; ti < n_obs; ti++) {
        const double xo = (obs[ti] - shift) / scale;
        double match = eln_d(normalPDF_d(mu[0], sg[0], xo));
        firstI_curr = NaN;
        firstI_curr = lnSum_d(firstI_curr, lnProd_d(lnProd_d(start_prev, t.l25), 0.0));
        firstI_curr = lnSum_d(firstI_curr, lnProd_d(lnProd_d(firstI_prev, t.l25), 0.0));
        Ic[0] = lnSum_d(NaN, lnProd_d(lnProd_d(Ip[0], t.I2I), 0.0));
        Ic[0] = lnSum_d(Ic[0], lnProd_d(lnProd_d(Mp[0], t.M2I), 0.0));
        Mc[0] = lnSum_d(NaN, lnProd_d(lnProd_d(firstI_prev, t.l5), match));
        Mc[0] = lnSum_d(Mc[0], lnProd_d(lnProd_d(Mp[0], t.M2M_int), match));
        Mc[0] = lnSum_d(Mc[0], lnProd_d(lnProd_d(start_prev, t.l5), match));
        Dc[0] = lnSum_d(NaN, lnProd_d(NaN, t.l25));
        Dc[0] = lnSum_d(Dc[0], lnProd_d(firstI_curr, t.l25));
        for (uint32_t i = 1; i < n; i++) {
            match = eln_d(normalPDF_d(mu[i], sg[i], xo));
            Ic[i] = lnSum_d(NaN, lnProd_d(lnProd_d(Ip[i], t.I2I), 0.0));
            Ic[i] = lnSum_d(Ic[i], lnProd_d(lnProd_d(Mp[i], t.M2I), 0.0));
            double mc = lnSum_d(NaN, lnProd_d(lnProd_d(Ip[i - 1], t.I2M), match));
            mc = lnSum_d(mc, lnProd_d(lnProd_d(Mp[i - 1], t.M2M_ext), match));
            mc = lnSum_d(mc, lnProd_d(lnProd_d(Mp[i], t.M2M_int), match));
            mc = lnSum_d(mc, lnProd_d(lnProd_d(Dp[i - 1], t.D2M), match));
            Mc[i] = mc;
        }
        for (uint32_t i = 1; i < n; i++) {
            double dc = lnSum_d(NaN, lnProd_d(Mc[i - 1], t.M2D));
            Dc[i] = lnSum_d(dc, lnProd_d(Dc[i - 1], t.D2D));
        }
        for (uint32_t i = 0; i < n; i++) { Ip[i] = Ic[i]; Mp[i] = Mc[i]; Dp[i] = Dc[i]; }
        firstI_prev = firstI_curr;
        start_prev = NaN;                                       // start_curr is never set (:249, :356)
    }
    double fwd = NaN;
    fwd = lnSum_d(fwd, lnProd_d(Dc[n - 1], eln_d(1.0)));
    fwd = lnSum_d(fwd, lnProd_d(Mc[n - 1], lnSum_d(t.M2M_ext, t.M2D)));
    fwd = lnSum_d(fwd, lnProd_d(Ic[n - 1], t.I2M));
    *out = fwd;
}

__global__ void __launch_bounds__(128) hmm_forward_kernel(const double *obs, const uint64_t *obs_off, const char *seq,
                                                         const double *shift, const double *scale, const double *epb,
                                                         size_t n_sites, uint32_t window, DnbModelDev unl,
                                                         DnbModelDev ana, double *out_a, double *out_t) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sites) return;
    const uint32_t n = 2 * window;
    const char *sn = seq + s * (size_t)(2 * window + DNB_K);
    double mu_t[2 * HMM_MAXW], sg_t[2 * HMM_MAXW], mu_a[2 * HMM_MAXW], sg_a[2 * HMM_MAXW];
    const uint32_t a_start = window - DNB_K / 2, a_end = window + DNB_K / 2;   // detect.cpp:544-545
    for (uint32_t i = 0; i < n; i++) {
        uint32_t rk = 0;
        bool hasT = false;
        for (int j = 0; j < DNB_K; j++) {
            const char c = sn[i + j];
            rk = rk * 4u + dnb_base_code(c);
            hasT |= c == 'T';
        }
        mu_t[i] = unl.mean[rk]; sg_t[i] = unl.stdv[rk];
        const bool use_a = i >= 1 && a_start <= i && i <= a_end && hasT;   // position 0 always unlabelled (:281)
        mu_a[i] = use_a ? ana.mean[rk] : mu_t[i];
        sg_a[i] = use_a ? ana.stdv[rk] : sg_t[i];
    }
    Trans t;
    // HMM_TransitionProbs_DNA_R10 {0.3, 0.7, 0.999, 0.0025, 0.001, 0.001}, src/config.h:42
    t.D2D = log(0.3); t.D2M = log(0.7); t.I2M = log(0.999); t.M2D = log(0.0025); t.M2I = log(0.001); t.I2I = log(0.001);
    t.M2M_int = eln_d(1. - (1. / epb[s]));
    t.M2M_ext = eln_d(1.0 - t.M2D - t.M2I - t.M2M_int);     // quirk Q10: log-space values, on purpose
    t.l25 = log(0.25); t.l5 = log(0.5);
    const double *o = obs + obs_off[s];
    const uint32_t n_obs = (uint32_t)(obs_off[s + 1] - obs_off[s]);
    forward_pass(o, n_obs, shift[s], scale[s], mu_a, sg_a, n, t, &out_a[s]);
    forward_pass(o, n_obs, shift[s], scale[s], mu_t, sg_t, n, t, &out_t[s]);
}

}  // namespace

void dnb_launch_hmm_forward(const double *obs, const uint64_t *obs_off, const char *seq, const double *shift,
                            const double *scale, const double *epb, size_t n_sites, uint32_t window,
                            const DnbModelDev &unl, const DnbModelDev &ana, double *out_analogue, double *out_thymidine,
                            cudaStream_t s) {
    if (n_sites == 0) return;
    hmm_forward_kernel<<<(unsigned)((n_sites + 127) / 128), 128, 0, s>>>(obs, obs_off, seq, shift, scale, epb, n_sites,
                                                                         window, unl, ana, out_analogue, out_thymidine);
}
