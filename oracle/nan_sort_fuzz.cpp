// TEST INFRASTRUCTURE: targeted fuzz of dnascent_b200/csrc/nan_sort_path.cuh (nsp_follow_lists, the form the device runs) against the real
// std::sort -- see oracle/nan_sort_check.cpp for the randomised check.  Build and run:
//   g++ -O2 -std=c++14 -x c++ -I dnascent_b200/csrc oracle/nan_sort_fuzz.cpp -o oracle/_build/nan_sort_fuzz && oracle/_build/nan_sort_fuzz
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include "nan_sort_path.cuh"
static bool same(double a, double b) { return (a != a && b != b) || (a == b && std::signbit(a) == std::signbit(b)); }
static long bad = 0, checked = 0, notemu = 0;
static void check(const std::vector<double> &v, long full) {
    const long n = (long)v.size();
    std::vector<double> real = v, e2 = v, clean;
    std::sort(real.begin(), real.end());
    for (double x : v) if (x == x) clean.push_back(x);
    std::sort(clean.begin(), clean.end());
    std::vector<int> la(n), ld(n);
    const NspResult r = nsp_follow_lists(e2.data(), n, la.data(), ld.data(), full);
    if (!r.ok) { notemu++; return; }
    checked++;
    for (long m = 0; m < n; m++) {      // EVERY index, not only around the median
        const double want = real[m], got = (m >= r.f && m < r.l) ? e2[m] : clean[m < r.f ? m : m - 1];
        if (!same(want, got)) { if (bad++ < 5) printf("MISMATCH n %ld full %ld index %ld [%ld,%ld) want %.17g got %.17g\n", n, full, m, r.f, r.l, want, got); return; }
    }
}
int main() {
    std::mt19937_64 rng(777);
    std::normal_distribution<double> nd(1.0, 0.05);
    const long fulls[3] = {16, 256, 2048};
    for (long n = 17; n <= 400; n += (n < 80 ? 1 : 7))
        for (int kind = 0; kind < 4; kind++) {
            std::vector<double> base(n);
            for (long i = 0; i < n; i++) base[i] = kind == 0 ? nd(rng) : kind == 1 ? (double)(rng() % 5) : kind == 2 ? (double)i : (double)(n - i);
            for (long at = 0; at < n; at++) {
                std::vector<double> v = base;
                v[at] = -std::nan("");
                check(v, fulls[(n + at) % 3]);
            }
        }
    // larger arrays: NaN at the positions the top-level median-of-3 looks at, and around them
    for (int t = 0; t < 300; t++) {
        const long n = 3000 + (long)(rng() % 200000);
        std::vector<double> base(n);
        const int kind = t % 4;
        for (long i = 0; i < n; i++) base[i] = kind == 0 ? nd(rng) : kind == 1 ? (double)(rng() % 50) * 0.5 : kind == 2 ? (double)i : (double)(n - i);
        const long probes[8] = {0, 1, 2, n / 2 - 1, n / 2, n / 2 + 1, n - 2, n - 1};
        for (long at : probes) {
            std::vector<double> v = base;
            v[at] = -std::nan("");
            check(v, fulls[t % 3]);
        }
    }
    printf("%s: %ld arrays checked at every index, %ld mismatching, %ld not emulated (pivot-NaN range > %d)\n", bad ? "FAIL" : "ok", checked, bad, notemu, NSP_PIVOT_MAX);
    return bad != 0;
}
