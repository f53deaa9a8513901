// puct.cuh — PUCTAgent's root rule (agents/mcts.py:276-315), __host__ __device__ so that the CPU
// tests can check it against vectors generated from the reference.
//
//   puct_a = q_hat_a + c_puct * pi_a * sqrt(N + 1e-9) / (1 + n_a)                       (:301)
//   q_hat  = clip((q - min) / (max - min), 0, 1),  q_a = mean outcome of a, or "mean" if unvisited (:299-300)
//   (max, min, "mean") = (0, -10, -5) while fewer than 10 outcomes exist, else the max, min and
//   MEDIAN of all outcomes so far                                                        (:304-315)
//   choice = first index with the strictly largest puct; NaNs never win, so when every outcome is equal
//   (0/0) the first card is chosen                                                        (:286-293)
// Outcomes are integers in [-171, 0]; their multiset is kept as a histogram of magnitudes.
#pragma once
#include <cstdint>
#include "game.cuh"

namespace nimmt {

constexpr int kOutcomeBins = 172;   // bull heads in the deck: 171

struct RootStats {
    int count[10];
    int sum[10];          // sum of outcomes (<= 0)
    long long sumsq[10];
    uint16_t hist[kOutcomeBins];   // hist[m] = number of outcomes equal to -m
    int total;
    int min_mag, max_mag;          // smallest / largest magnitude seen (= max / min outcome); valid when total > 0
};

NIMMT_HD void root_stats_clear(RootStats& s) {
    for (int a = 0; a < 10; ++a) { s.count[a] = 0; s.sum[a] = 0; s.sumsq[a] = 0; }
    for (int m = 0; m < kOutcomeBins; ++m) s.hist[m] = 0;
    s.total = 0;
    s.min_mag = kOutcomeBins;
    s.max_mag = -1;
}

NIMMT_HD void root_stats_add(RootStats& s, int action_index, int outcome) {
    s.count[action_index] += 1;
    s.sum[action_index] += outcome;
    s.sumsq[action_index] += (long long)outcome * outcome;
    s.hist[-outcome] += 1;
    s.total += 1;
    s.min_mag = -outcome < s.min_mag ? -outcome : s.min_mag;
    s.max_mag = -outcome > s.max_mag ? -outcome : s.max_mag;
}

// np.median of the multiset of outcomes: mean of the two middle order statistics.  Outcomes cluster near
// zero, so the scan runs over magnitudes from 0 upwards (k-th smallest outcome = (total - 1 - k)-th smallest
// magnitude) and stops at the upper middle one, typically after a dozen bins.
NIMMT_HD double median_outcome(const RootStats& s) {
    const int k_lo = s.total - 1 - s.total / 2, k_hi = s.total - 1 - (s.total - 1) / 2;   // k_lo <= k_hi, magnitude ranks
    int seen = 0, m_lo = -1;
    for (int m = s.min_mag; m <= s.max_mag; ++m) {
        seen += s.hist[m];
        if (m_lo < 0 && seen > k_lo) m_lo = m;
        if (seen > k_hi) return -0.5 * ((double)m_lo + (double)m);
    }
    return 0.0;
}

// (max, min, "mean") of _normalize_q (agents/mcts.py:304-315).
struct PuctBounds { double mx, mn, md, root_n; };
NIMMT_HD PuctBounds puct_bounds(const RootStats& s) {
    PuctBounds b;
    if (s.total < 10) {
        b.mx = 0.0; b.mn = -10.0; b.md = -5.0;
    } else {
        b.mn = (double)-s.max_mag;
        b.mx = (double)-s.min_mag;
        b.md = median_outcome(s);   // np.median
    }
    b.root_n = sqrt((double)s.total + 1.0e-9);
    return b;
}

// PUCT value of legal card a (agents/mcts.py:295-302).
NIMMT_HD double puct_value(const RootStats& s, const PuctBounds& b, int a, float prob, float c_puct) {
    const double q = s.count[a] > 0 ? (double)s.sum[a] / (double)s.count[a] : b.md;
    double qn = (q - b.mn) / (b.mx - b.mn);        // 0/0 -> NaN when all outcomes are equal: kept
    qn = qn < 0.0 ? 0.0 : (qn > 1.0 ? 1.0 : qn);   // np.clip; NaN compares false twice and passes through
    const float cp = c_puct * prob;                // float32 product first, as numpy does (python float * float32 array)
    return qn + (double)cp * b.root_n / (1.0 + (double)s.count[a]);
}

// Returns the index (into the n legal cards, ascending) PUCT selects; pucts[] receives the values.
NIMMT_HD int puct_choose(const RootStats& s, const float* probs, int n, float c_puct, double* pucts) {
    const PuctBounds b = puct_bounds(s);
    int choice = 0;
    double best = -INFINITY;
    for (int a = 0; a < n; ++a) {
        const double p = puct_value(s, b, a, probs[a], c_puct);
        if (pucts) pucts[a] = p;
        if (p > best) { best = p; choice = a; }
    }
    return choice;
}

#ifdef __CUDACC__
// The same rule as the search kernel evaluates it (k_policy_rollouts, and k_puct_cases behind nimmt_puct_choose): the
// decision's hand slots sit in lanes gbase .. gbase + 9 of a warp; lane `slot` computes the PUCT value of card `slot`
// (agents/mcts.py:295-302) and every lane of the decision then runs the strict-'>' scan over the gathered values in
// ascending card order (:286-293; NaN never wins, so 0/0 everywhere selects the first card).  All 32 lanes must call it;
// `active` = this lane holds a legal card of a decision that is choosing.  Returns the chosen hand slot; `mine` receives this
// lane's own PUCT value (-inf for inactive lanes).
__device__ __forceinline__ int puct_choose_lanes(const RootStats& s, bool active, int slot, int h, int gbase, float prob, float c_puct,
                                                 double& mine) {
    mine = -INFINITY;
    if (active && slot < h) mine = puct_value(s, puct_bounds(s), slot, prob, c_puct);
    double best = -INFINITY;
    int choice = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const double p = __shfl_sync(0xffffffffu, mine, gbase + i);
        if (i < h && p > best) { best = p; choice = i; }
    }
    return choice;
}
#endif

}  // namespace nimmt
