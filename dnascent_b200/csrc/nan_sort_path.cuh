// nan_sort_path.cuh -- where ONE NaN ends up when libstdc++'s std::sort runs over an array of doubles, and what the
// array looks like around it.
//
// Why this exists.  estimateScaling_theilSen (/root/reference/src/event_handling.cpp:67-78) pushes every pairwise slope
// dy/dx into a vector, std::sorts it and takes element [size/2].  Two cleaned points with identical signal and model
// level give 0/0 = NaN (about one read in 1000), and std::sort with a NaN violates its strict-weak-ordering
// precondition: the result is whatever the algorithm happens to do.  It is still deterministic -- libstdc++'s
// introsort (median-of-3 pivot moved to the front, unguarded Hoare partition, insertion sort below 16 elements) treats
// the NaN as "equal" to every pivot, so it drifts through the partitions, and in the final insertion sort it is a
// barrier nothing crosses -- but WHERE it ends up (before or after the median) depends on the exact sequence of swaps.
// The 2000-read statistical run of round 2 found one read where it lands after the median and one where it lands
// before, so neither "NaN first" nor "NaN last" reproduces the reference.
//
// What is emulated.  Only the branch of the recursion that contains the NaN has to be followed literally: every
// partition of an enclosing range is applied to the actual data (O(range) each, ~2n steps in total), the other
// branches are never sorted.  When the range is small (<= NSP_FULL elements) or the NaN itself is chosen as a pivot
// (then the partition orders nothing and the two halves only get merged by the final insertion sort, across which the
// NaN is a barrier), the whole remaining range is sorted literally: introsort recursion with the remaining depth
// budget, then the insertion pass.  Outside that range every other element is correctly ordered against it, so the
// caller gets: the range [f, l) with its exact final contents, and the rule "index m < f is the m-th smallest non-NaN
// value, index m >= l the (m-1)-th".  Transcribed from GCC 13's bits/stl_algo.h (std::__introsort_loop,
// std::__unguarded_partition_pivot, std::__move_median_to_first, std::__unguarded_partition,
// std::__final_insertion_sort, std::__unguarded_linear_insert); checked against the real std::sort on the host
// (tests/test_host_logic.py::test_nan_sort_path_matches_std_sort, oracle/nan_sort_check.cpp).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define NSP_HD __host__ __device__
#else
#define NSP_HD
#endif

#define NSP_THRESHOLD 16     // std::_S_threshold
#ifndef NSP_FULL
#define NSP_FULL 2048        // ranges up to this size are sorted literally (any value >= NSP_THRESHOLD gives the same result)
#endif
#define NSP_PIVOT_MAX 4096   // largest range that is sorted literally when the NaN itself becomes the pivot

struct NspResult {
    long f, l;        // the literally sorted range [f, l) (final contents are in the array)
    bool ok;          // false: the depth limit was reached on the NaN's branch (heapsort fallback not emulated)
};

NSP_HD inline int nsp_lg(long n) { int k = 0; while (n > 1) { n >>= 1; k++; } return k; }   // std::__lg
NSP_HD inline void nsp_swap(double *v, long a, long b) { const double t = v[a]; v[a] = v[b]; v[b] = t; }

// std::__move_median_to_first(result, a, b, c) with operator<
NSP_HD inline void nsp_median_to_first(double *v, long result, long a, long b, long c) {
    if (v[a] < v[b]) {
        if (v[b] < v[c]) nsp_swap(v, result, b);
        else if (v[a] < v[c]) nsp_swap(v, result, c);
        else nsp_swap(v, result, a);
    } else if (v[a] < v[c]) nsp_swap(v, result, a);
    else if (v[b] < v[c]) nsp_swap(v, result, c);
    else nsp_swap(v, result, b);
}
// std::__unguarded_partition(first, last, pivot)
NSP_HD inline long nsp_partition(double *v, long first, long last, long pivot) {
    for (;;) {
        while (v[first] < v[pivot]) ++first;
        --last;
        while (v[pivot] < v[last]) --last;
        if (!(first < last)) return first;
        nsp_swap(v, first, last);
        ++first;
    }
}
// std::__unguarded_partition_pivot(first, last)
NSP_HD inline long nsp_partition_pivot(double *v, long first, long last) {
    const long mid = first + (last - first) / 2;
    nsp_median_to_first(v, first, first + 1, mid, last - 1);
    return nsp_partition(v, first + 1, last, first);
}

// literal std::__introsort_loop over [first, last) with an explicit stack; false if the depth limit is reached
NSP_HD inline bool nsp_introsort_full(double *v, long first, long last, int depth_limit) {
    long stack_first[96], stack_last[96];
    int stack_depth[96];
    int sp = 0;
    stack_first[sp] = first; stack_last[sp] = last; stack_depth[sp] = depth_limit; sp++;
    while (sp > 0) {
        sp--;
        long f = stack_first[sp], l = stack_last[sp];
        int d = stack_depth[sp];
        while (l - f > NSP_THRESHOLD) {
            if (d == 0) return false;
            --d;
            const long cut = nsp_partition_pivot(v, f, l);
            if (sp >= 95) return false;
            stack_first[sp] = cut; stack_last[sp] = l; stack_depth[sp] = d; sp++;     // __introsort_loop(cut, last, depth)
            l = cut;
        }
    }
    return true;
}

// the insertion pass of std::__final_insertion_sort restricted to [f, l), where every element left of f is <= and every
// element right of l is >= the elements of the range (so nothing enters or leaves it); `global_first16` tells whether
// the guarded form applies (indices < 16 of the whole array)
NSP_HD inline bool nsp_insertion_pass(double *v, long f, long l) {
    long budget = 200000000;      // a pivot-NaN partition of a huge range makes this pass quadratic (the reference too): give up
    for (long i = f; i < l; i++) {
        const double val = v[i];
        if (i < NSP_THRESHOLD) {
            // std::__insertion_sort on the first 16 elements of the whole array: guarded
            if (i == 0) continue;
            if (val < v[0]) {
                for (long j = i; j > 0; j--) v[j] = v[j - 1];      // move_backward(first, i, i + 1)
                v[0] = val;
                continue;
            }
        }
        long last = i, next = i - 1;                                // std::__unguarded_linear_insert
        while (next >= f && val < v[next]) {                        // left of f everything is <= val: the walk stops there
            v[last] = v[next];
            last = next;
            --next;
            if (--budget < 0) return false;
        }
        v[last] = val;
    }
    return true;
}

// ---- the partition as the device runs it ---------------------------------------------------------------------------
// std::__unguarded_partition is a two-pointer walk, serial as written.  Its outcome has a closed form, which a CTA can
// evaluate with prefix counts.  Let A = a_1 < a_2 < ... be the positions in [first, last) whose value is NOT less than
// the pivot (where the left pointer stops), D = d_1 > d_2 > ... the positions whose value the pivot is NOT less than
// (where the right pointer stops), both taken on the array as it is before the partition; a NaN and a copy of the
// pivot are in both.  Until the pointers cross they only ever read untouched elements, so round k swaps a_k with d_k,
// for k = 1..m with m = the number of k for which a_k < d_k (a_k grows, d_k falls: the condition is monotone).  The
// value returned is min(a_{m+1}, d_m) (a_{m+1} = the next untouched stop of the left pointer if there is one, d_m =
// the swapped-in element that stops it otherwise; for m = 0 it is a_1).  A position is never in two swaps.
// listA / listD receive the positions in ASCENDING order (capacity last - first each); returns the cut.
NSP_HD inline long nsp_partition_lists(double *v, long first, long last, long pivot, int *listA, int *listD) {
    const double p = v[pivot];
    long ta = 0, td = 0;
    for (long i = first; i < last; i++) {                    // device: one pass of block-wide prefix counts
        const double x = v[i];
        if (!(x < p)) listA[ta++] = (int)i;
        if (!(p < x)) listD[td++] = (int)i;
    }
    long lo = 0, hi = ta < td ? ta : td;                     // m = number of k in 1..min(ta,td) with a_k < d_k
    while (lo < hi) {
        const long k = (lo + hi + 1) / 2;
        if (listA[k - 1] < listD[td - k]) lo = k; else hi = k - 1;
    }
    const long m = lo;
    for (long k = 0; k < m; k++) nsp_swap(v, listA[k], listD[td - 1 - k]);       // device: in parallel
    long cut = m >= 1 ? (long)listD[td - m] : last;
    if (m < ta && listA[m] < cut) cut = listA[m];
    return cut;
}

// v[0..n): exactly one NaN.  Emulates std::sort(v, v + n) along the NaN's branch; on return [f, l) holds its final contents.
NSP_HD inline NspResult nsp_follow(double *v, long n) {
    NspResult r;
    r.ok = true;
    long first = 0, last = n;
    int depth = nsp_lg(n) * 2;
    long pos = -1;
    for (long i = 0; i < n; i++) if (v[i] != v[i]) { pos = i; break; }
    while (last - first > NSP_FULL) {
        if (depth == 0) { r.ok = false; break; }
        --depth;
        const long mid = first + (last - first) / 2;
        nsp_median_to_first(v, first, first + 1, mid, last - 1);
        if (v[first] != v[first]) {
            // the NaN is the pivot: this partition orders nothing; sort the whole range literally from here
            const long cut = nsp_partition(v, first + 1, last, first);
            bool ok = nsp_introsort_full(v, cut, last, depth);
            ok = nsp_introsort_full(v, first, cut, depth) && ok;
            ok = nsp_insertion_pass(v, first, last) && ok;
            r.f = first; r.l = last; r.ok = ok;
            return r;
        }
        const long cut = nsp_partition(v, first + 1, last, first);
        // which side holds the NaN now?
        pos = -1;
        // the NaN moved at most once per partition; find it on the smaller side first
        const long nl = cut - first, nr = last - cut;
        if (nl <= nr) {
            for (long i = first; i < cut; i++) if (v[i] != v[i]) { pos = i; break; }
            if (pos >= 0) last = cut; else first = cut;
        } else {
            for (long i = cut; i < last; i++) if (v[i] != v[i]) { pos = i; break; }
            if (pos >= 0) first = cut; else last = cut;
        }
    }
    if (r.ok) r.ok = nsp_introsort_full(v, first, last, depth);
    if (r.ok) r.ok = nsp_insertion_pass(v, first, last);
    r.f = first; r.l = last;
    return r;
}

// nsp_follow with the partition in its closed form (what theilsen_nan.cu runs; checked against nsp_follow and the real
// std::sort by oracle/nan_sort_check.cpp).  A pivot-NaN range larger than NSP_PIVOT_MAX is not emulated (ok = false).
NSP_HD inline NspResult nsp_follow_lists(double *v, long n, int *listA, int *listD, long full) {
    NspResult r;
    r.ok = true;
    long first = 0, last = n;
    int depth = nsp_lg(n) * 2;
    long pos = -1;
    for (long i = 0; i < n; i++) if (v[i] != v[i]) { pos = i; break; }
    while (last - first > full) {
        if (depth == 0) { r.ok = false; break; }
        --depth;
        const long mid = first + (last - first) / 2;
        nsp_median_to_first(v, first, first + 1, mid, last - 1);
        if (v[first] != v[first]) {
            r.f = first; r.l = last;
            if (last - first > NSP_PIVOT_MAX) { r.ok = false; return r; }
            const long cut = nsp_partition(v, first + 1, last, first);
            bool ok = nsp_introsort_full(v, cut, last, depth);
            ok = nsp_introsort_full(v, first, cut, depth) && ok;
            ok = nsp_insertion_pass(v, first, last) && ok;
            r.ok = ok;
            return r;
        }
        if (v[first + 1] != v[first + 1]) pos = first + 1;       // the median-of-3 may have moved the NaN
        else if (v[mid] != v[mid]) pos = mid;
        else if (v[last - 1] != v[last - 1]) pos = last - 1;
        const long cut = nsp_partition_lists(v, first + 1, last, first, listA, listD);
        if (!(v[pos] != v[pos])) {                               // it was swapped: its partner is where it is now
            long q = -1;
            for (long i = first + 1; i < last; i++) if (v[i] != v[i]) { q = i; break; }   // device: recorded by the swapping thread
            pos = q;
        }
        if (pos < cut) last = cut; else first = cut;
    }
    if (r.ok) r.ok = nsp_introsort_full(v, first, last, depth);
    if (r.ok) r.ok = nsp_insertion_pass(v, first, last);
    r.f = first; r.l = last;
    return r;
}
