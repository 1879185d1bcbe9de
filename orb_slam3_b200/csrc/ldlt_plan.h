// Host-side planning of the two-sided reduced solve (lba.cu: rev_gather_kernel / ldlt_win_kernel modes 1 and 2).
// Plain C++, no CUDA: the same code runs inside lba_solve and behind lba_debug_two_sided_plan, which
// tests/test_ldlt_plan.py drives on the CPU together with a numpy emulation of the elimination.
#pragma once
#include <algorithm>
#include <climits>
#include <vector>

namespace orbb200 {

struct TwoSidedPlan {
  bool ok = false;
  int m = 0;    // side 0 eliminates columns [0, m)
  int e2 = 0;   // side 1 eliminates columns [e2, n); the separator is [m, e2)
  int p0 = 0, p1 = 0;  // panels per side (m = panel * p0, n - e2 = panel * p1)
  int w = 0;    // e2 - m
  int R0 = 0;   // last row of side 0's window when it stops (caller's numbering)
  int R1 = 0;   // last row of side 1's window when it stops (reversed numbering)
  std::vector<int> first1, reach1;  // row envelope of P S P (side 1's tables)
};

// env_first[i] = first column of row i's envelope, env_reach[c] = last row whose envelope holds a column <= c
// (non-decreasing).  panel = 8, max_rows = rows a window may hold (WIN_ROWS).
inline TwoSidedPlan plan_two_sided(int n, const std::vector<int>& env_reach, int panel, int max_rows) {
  TwoSidedPlan P;
  P.first1.resize(n); P.reach1.assign(n, 0);
  for (int a = 0; a < n; a++) P.first1[a] = n - 1 - env_reach[n - 1 - a];  // column envelope of S = row envelope of P S P
  for (int a = 0; a < n; a++) P.reach1[P.first1[a]] = std::max(P.reach1[P.first1[a]], a);
  for (int c = 1; c < n; c++) P.reach1[c] = std::max(P.reach1[c], P.reach1[c - 1]);
  int rows1 = 0;
  for (int k0 = 0; k0 < n; k0 += panel) {
    const int nb = std::min(panel, n - k0);
    const int R = std::min(std::max(P.reach1[k0 + nb - 1], k0 + nb - 1), n - 1);
    int jmin = k0;
    for (int r = 0; r < nb; r++) jmin = std::min(jmin, P.first1[k0 + r]);
    rows1 = std::max(rows1, std::max(R - k0 + 1, k0 - jmin + nb));
  }
  if (rows1 > max_rows) return P;
  int best = INT_MAX;
  for (int pm = 1; panel * pm < n; pm++) {
    const int m = panel * pm, e = env_reach[m - 1] + 1;  // rows >= e do not reach a column < m
    const int p1 = (n - e) / panel;
    if (p1 < 1) break;
    const int e2 = n - panel * p1, w = e2 - m;
    if (w < panel || w > max_rows - panel) continue;
    if (std::max(pm, p1) < best) {
      best = std::max(pm, p1);
      P.m = m; P.e2 = e2; P.p0 = pm; P.p1 = p1; P.w = w;
    }
  }
  if (best == INT_MAX) return P;
  P.ok = true;
  P.R0 = std::min(std::max(env_reach[P.m - 1], P.m - 1), n - 1);
  const int m1 = panel * P.p1;
  P.R1 = std::min(std::max(P.reach1[m1 - 1], m1 - 1), n - 1);
  return P;
}

}  // namespace orbb200
