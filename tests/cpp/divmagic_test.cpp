// CPU check of fiesta_b200/csrc/fb_divmagic.h: the multiply-high division k_x_relax uses to turn a voxel index into coordinates.
// For every divisor a grid can have (z pitch and y size are <= 1024; checked up to 2050) the computed quotient is compared with
// n / d on both sides of EVERY multiple of d below 2^31.  Both are non-decreasing step functions of n, so equality at all step
// edges is equality everywhere.
#include <cstdio>
#include "../../fiesta_b200/csrc/fb_divmagic.h"
int main() {
  unsigned long long checks = 0;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : checks) reduction(| : bad)
  for (unsigned d = 1; d <= 2050; ++d) {
    unsigned m, s;
    fb_div_make(d, m, s);
    for (unsigned long long k = 0;; ++k) {
      const unsigned long long lo = k * d, hi = lo + d - 1;            // the run of n with n / d == k
      if (lo >= (1ull << 31)) break;
      const unsigned n1 = (unsigned)lo, n2 = (unsigned)(hi < (1ull << 31) ? hi : (1ull << 31) - 1);
      if (fb_div_apply(n1, m, s) != n1 / d || fb_div_apply(n2, m, s) != n2 / d) { std::printf("FAIL d=%u n=%u/%u\n", d, n1, n2); bad = 1; break; }
      checks += 2;
      if (d == 1 && k > (1u << 20)) break;                            // d == 1 is the identity
    }
  }
  if (bad) return 1;
  std::printf("OK %llu checks\n", checks);
  return 0;
}
