// Exhaustive host check of csrc/glibc_sincosf.h against the libm of this box: every float of [0, 2*pi*1.001]
// (the extractor's angle * pi/180 lies in [0, 2*pi]), cosf and sinf, fused and unfused variants.  Also counts
// how often the device's former `(float)cos((double)x)` formulation differs, and for how many of those angles
// a rotated rBRIEF pattern coordinate moves (the reason the exact restatement exists).
// Prints: n_floats mismatch_fused mismatch_unfused double_rounding_differs pattern_coordinate_moves
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../orb_slam3_b200/csrc/glibc_sincosf.h"

static const int kPattern[1024] = {
#include "../../orb_slam3_b200/csrc/pattern_31.inc"
};

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 8;
  const float hi = 6.2832f * 1.001f;
  uint32_t top;
  memcpy(&top, &hi, 4);
  std::atomic<long long> bad_f{0}, bad_u{0}, dbl{0}, moves{0};
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      long long bf = 0, bu = 0, d = 0, mv = 0;
      for (uint64_t u = t; u <= top; u += T) {
        const uint32_t b = (uint32_t)u;
        float x;
        memcpy(&x, &b, 4);
        const float c = cosf(x), s = sinf(x);
        if (c != glibc_sincosf::cosf_exact<true>(x) || s != glibc_sincosf::sinf_exact<true>(x)) bf++;
        if (c != glibc_sincosf::cosf_exact<false>(x) || s != glibc_sincosf::sinf_exact<false>(x)) bu++;
        const float c2 = (float)cos((double)x), s2 = (float)sin((double)x);
        if (c != c2 || s != s2) {
          d++;
          for (int i = 0; i < 512; i++) {
            const float px = (float)kPattern[2 * i], py = (float)kPattern[2 * i + 1];
            if (lrintf(px * s + py * c) != lrintf(px * s2 + py * c2) || lrintf(px * c - py * s) != lrintf(px * c2 - py * s2)) {
              mv++;
              break;
            }
          }
        }
      }
      bad_f += bf; bad_u += bu; dbl += d; moves += mv;
    });
  for (auto& x : th) x.join();
  printf("%u %lld %lld %lld %lld\n", top + 1, (long long)bad_f, (long long)bad_u, (long long)dbl, (long long)moves);
  return 0;
}
