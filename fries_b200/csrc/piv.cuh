// Pivotal compression family: piv_samp_serial / adjust_probs / piv_budget / piv_comp_parallel,
// FRIES/compress_utils.cpp:354-681.
//
// The reference's piv_samp_serial is a sequential sweep: the cumulative line of the non-preserved magnitudes is cut
// into n_samp units of width u = seg_norm / n_samp; in unit k it (1) draws a candidate H_k among the element carried
// over from unit k-1 and the elements that end inside the unit, with probability proportional to the weight each
// has inside the unit, and (2) with probability a_k / (u - b_k) takes the element that straddles the unit's upper
// border as the sample and carries H_k on, otherwise takes H_k and carries the straddling element on (a_k / b_k = the
// straddling element's weight below / above the border).  Two mt19937 draws per unit.
//
// Parallel restatement (what the kernels in piv.cu do).  Every quantity of unit k except the IDENTITY of the carried
// element is a function of the inclusive prefix sums E[] alone:
//   * the straddling element cross[k] is the one with E[c-1] < b(k+1) <= E[c], b(m) = m * u;
//   * step (1) is "the element that contains the point p = b(k) + r1 * (E[c-1] - b(k))" -- if that is cross[k-1], the
//     candidate is the carried element, whoever that is;
//   * step (2) compares r2 with a_k / (u - b_k), a_k = b(k+1) - E[c-1], b_k = E[c] - b(k+1).
// The carried element of unit k+1 is either a constant (the straddling element, or an interior candidate) or -- when
// the candidate was the carried element and it is carried again -- the same as in unit k.  So each unit emits
// (sample, carry) with the markers PIV_CARRIED / PIV_SAME, and a unit whose sample is "the carried one" walks back
// over its predecessors to the last constant (a geometric number of steps).  One thread per unit, no sequential sweep.
// Differences from the sequential arithmetic are FP ties only: the reference restarts its running sum in every unit,
// here all comparisons are made on the global prefix sums, and the very last unit closes with or without a straddling
// element depending on whether the prefix total reaches fl(n_samp * u).
#pragma once
#include <cstdint>
#include <cmath>
#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define PIV_NONE 0xffffffffu     // no element
#define PIV_CARRIED 0xfffffffeu  // sample: "the element carried into this unit"
#define PIV_SAME 0xfffffffdu     // carry: "the same element that was carried into this unit"

struct PivGrid {
    double unit, inv;
    uint32_t n_samp;
    __host__ __device__ __forceinline__ double border(uint64_t m) const { return (double)m * unit; }
    // number of borders m = 1 .. n_samp with border(m) <= x
    __host__ __device__ __forceinline__ uint32_t borders_le(double x) const {
        if (!(x > 0) || n_samp == 0) return 0;
        double q = x * inv;
        uint64_t k = q >= (double)n_samp ? n_samp : (uint64_t)q;
        while (k > 0 && !(border(k) <= x)) k--;
        while (k < n_samp && border(k + 1) <= x) k++;
        return (uint32_t)k;
    }
};

__host__ __device__ __forceinline__ PivGrid piv_grid(double seg_norm, uint32_t n_samp) {
    PivGrid g;
    g.n_samp = n_samp;
    g.unit = n_samp ? seg_norm / n_samp : 0.0;
    g.inv = g.unit > 0 ? 1.0 / g.unit : 0.0;
    return g;
}

// Element i (prefix sums s = E[i-1], e = E[i]) straddles the borders m in (borders_le(s), borders_le(e)]: it is
// cross[m - 1].
template <class Store>
__host__ __device__ __forceinline__ void piv_mark_cross(const PivGrid &g, double s, double e, uint32_t i, Store &&store) {
    if (!(e > s)) return;
    uint32_t lo = g.borders_le(s), hi = g.borders_le(e);
    for (uint32_t m = lo + 1; m <= hi; m++) store(m - 1, i);
}

// units that exist: one per crossed border, plus a closing unit without straddling element when borders are left and
// elements remain behind the last straddling one (the reference's loop condition, compress_utils.cpp:414)
__host__ __device__ __forceinline__ uint32_t piv_n_units(const PivGrid &g, uint32_t n_crossed, uint32_t last_cross,
                                                         size_t n) {
    if (n == 0 || g.n_samp == 0) return 0;
    if (n_crossed >= g.n_samp) return g.n_samp;
    size_t next = n_crossed ? (size_t)last_cross + 1 : 0;
    return n_crossed + (next < n ? 1u : 0u);
}

// first j in [lo, hi] with E[j] >= p (E is non-decreasing; E[hi] >= p is guaranteed by the caller)
template <class LoadE>
__host__ __device__ __forceinline__ uint32_t piv_search(LoadE &&E, uint32_t lo, uint32_t hi, double p) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (E(mid) >= p) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// Unit k: E(j) loads a prefix sum, cross(k) a straddling element.  r1, r2 in [0, 1).
template <class LoadE, class LoadC>
__host__ __device__ __forceinline__ void piv_unit(const PivGrid &g, uint32_t k, uint32_t n_crossed, size_t n, LoadE &&E,
                                                  LoadC &&cross, double r1, double r2, uint32_t &sample,
                                                  uint32_t &carry) {
    const bool closing = k >= n_crossed;  // no straddling element: everything up to the end of the vector is interior
    const uint32_t c = closing ? (uint32_t)n : cross(k);
    const uint32_t prev = k ? cross(k - 1) : PIV_NONE;
    if (!closing && c == prev) {  // an element wider than a unit (outside the contract): it is this unit's sample too
        sample = c;
        carry = PIV_SAME;
        return;
    }
    const double b_lo = g.border(k), b_hi = g.border((uint64_t)k + 1);
    const double s_c = c ? E(c - 1) : 0.0;  // start of the straddling element = end of the interior ones
    double span = s_c - b_lo;               // carried share + interior weights
    if (!(span > 0)) span = 0;
    double p = b_lo + r1 * span;
    if (p > s_c) p = s_c;
    uint32_t cand;
    if (k == 0) {
        cand = (p > 0 && c > 0) ? piv_search(E, 0u, c - 1, p) : PIV_NONE;
    } else {
        uint32_t j = piv_search(E, prev, c - 1, p);  // c > prev here
        cand = j == prev ? PIV_CARRIED : j;
    }
    bool take_border = false;
    if (!closing) {
        double under = b_hi - s_c, over = E(c) - b_hi;
        take_border = r2 < under / (g.unit - over);
    }
    if (take_border) {
        sample = c;
        carry = cand == PIV_CARRIED ? PIV_SAME : cand;
    } else {
        sample = cand;
        carry = closing ? PIV_NONE : c;
    }
}

// the element carried INTO unit k: the last constant carry of the units before it
template <class LoadCarry>
__host__ __device__ __forceinline__ uint32_t piv_resolve(uint32_t k, LoadCarry &&carry) {
    while (k > 0) {
        uint32_t v = carry(--k);
        if (v != PIV_SAME) return v;
    }
    return PIV_NONE;
}

// final state of element i.  flag in: 0 resampled, 1 preserved, 2 resampled + drawn; out: 1 = zeroed element
__host__ __device__ __forceinline__ void piv_finish(double unit, uint32_t n_samp, double &v, uint8_t &flag) {
    if (flag == 2) {
        v = v > 0 ? unit : -unit;
        flag = 0;
    } else if (flag == 1) {
        flag = (n_samp == 0 && v == 0) ? 1 : 0;  // compress_utils.cpp:392-404
    } else {
        v = 0;
        flag = 1;
    }
}

// ---- adjust_probs (compress_utils.cpp:617-681) ----------------------------------------------------------------------
// The reference walks the non-preserved elements with a running `counter` and budget and stops at the first element
// where they meet.  g = counter - budget is a prefix sum of per-element increments that all have the same sign, so an
// element knows from its exclusive prefix whether the walk reaches it, and from the inclusive one whether it is last.
struct PivAdjust {
    double unit, frac, loc_norm, exp_loc;
    uint32_t n_loc;
    int up;  // the budget was rounded up
    __host__ __device__ __forceinline__ double g0() const { return exp_loc - (double)n_loc; }
    // increment of g and of the "made exact" count for an element of magnitude a
    __host__ __device__ __forceinline__ void delta(double a, double &dg, unsigned long long &dk) const {
        double pi = a / unit;
        dk = 0;
        if (up) {
            if (pi < frac) dg = pi / frac - pi;
            else {
                dg = 1.0 - pi;
                dk = 1;
            }
        } else {
            dg = pi > frac ? (pi - frac) / (1 - frac) - pi : -pi;
        }
    }
    __host__ __device__ __forceinline__ bool reached(double g_before) const { return up ? g_before < 0 : g_before > 0; }
    __host__ __device__ __forceinline__ bool last(double g_after) const { return up ? g_after >= 0 : g_after <= 0; }
    // transformed value; exact = the element now carries a whole unit and is preserved (up branch only)
    __host__ __device__ __forceinline__ double apply(double v, bool &exact) const {
        double a = fabs(v), pi = a / unit, sgn = v > 0 ? 1.0 : -1.0;
        exact = false;
        if (up) {
            if (pi < frac) return v / frac;
            exact = true;
            return sgn * unit;
        }
        return pi > frac ? sgn * ((pi - frac) / (1 - frac)) * unit : 0.0;
    }
};

__host__ __device__ __forceinline__ PivAdjust piv_adjust_setup(uint32_t n_loc, double exp_loc, uint32_t n_tot,
                                                               double tot_norm) {
    PivAdjust a;
    a.unit = tot_norm / n_tot;
    a.frac = exp_loc - (unsigned int)exp_loc;
    a.loc_norm = exp_loc * a.unit;
    a.exp_loc = exp_loc;
    a.n_loc = n_loc;
    a.up = (double)n_loc > exp_loc;
    return a;
}
