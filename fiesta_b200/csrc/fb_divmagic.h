// Division of a 31-bit unsigned number by a run-time constant as multiply-high + shift (host side: the constants; the device
// side applies them with __umulhi, k_x_relax's x_div).  Plain C++, no CUDA types: tests/cpp/divmagic_test.cpp checks it on the CPU.
//
// floor(n / d) == umulhi(n, m) >> s for every n < 2^31:  l = ceil(log2 d), m = floor(2^(31+l) / d) + 1 (< 2^32), s = l - 1.
// (m d = 2^(31+l) + e with 0 < e <= d <= 2^l, so n m / 2^(31+l) exceeds n / d by less than 1 / d.)   d == 1: m = 0 means "n itself".
#ifndef FB_DIVMAGIC_H_
#define FB_DIVMAGIC_H_
static inline void fb_div_make(unsigned d, unsigned &m, unsigned &s) {
  if (d <= 1u) { m = 0; s = 0; return; }
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  m = (unsigned)((1ull << (31 + l)) / d + 1ull);
  s = l - 1u;
}
// what the device computes: (n * m) >> 32 >> s
static inline unsigned fb_div_apply(unsigned n, unsigned m, unsigned s) {
  return m ? (unsigned)(((unsigned long long)n * m) >> 32) >> s : n;
}
#endif
