/* Check of the multiply-shift division used by the CSR kernels (ctp_prm.cuh, ctp_div_make / ctp_div_u31):
 * floor(x / d) == (x * mul) >> sh for 0 <= x < 2^31, 1 <= d < 2^31.  Every d up to LIMIT against the x values
 * where a wrong multiplier shows first (multiples of d and their predecessors, near 2^31), plus random pairs.
 * Prints "bad total". */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef struct { uint32_t d, mul; int sh; } CtpDiv;
static CtpDiv make(uint32_t d) {
  int s = 0;
  while ((1ull << s) < d) s++;
  CtpDiv r;
  r.d = d;
  r.mul = (uint32_t)(((1ull << (31 + s)) / d) + 1ull);
  r.sh = 31 + s;
  return r;
}
static uint32_t div31(uint32_t x, CtpDiv v) { return (uint32_t)(((unsigned long long)x * v.mul) >> v.sh); }
int main(int argc, char **argv) {
  const uint32_t limit = argc > 1 ? (uint32_t)atoll(argv[1]) : 200000u;
  const uint32_t top = 0x7fffffffu;
  long long bad = 0, total = 0;
  uint64_t rng = 88172645463325252ull;
  for (uint32_t d = 1; d <= limit; d++) {
    const CtpDiv v = make(d);
    if ((((1ull << (31 + (v.sh - 31))) / d) + 1ull) > 0xffffffffull) bad++;  /* multiplier must fit 32 bits */
    const uint32_t qmax = top / d;
    const uint32_t qs[6] = {0, 1, qmax / 2, qmax > 0 ? qmax - 1 : 0, qmax, qmax / 3};
    for (int a = 0; a < 6; a++)
      for (int off = -1; off <= 1; off++) {
        const int64_t x = (int64_t)qs[a] * d + off;
        if (x < 0 || x > top) continue;
        bad += div31((uint32_t)x, v) != (uint32_t)x / d;
        total++;
      }
    bad += div31(top, v) != top / d;
    total++;
  }
  for (int t = 0; t < 20000000; t++) {
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    uint32_t d = (uint32_t)(rng >> 33) >> (rng & 31 ? (rng & 31) % 31 : 0);
    if (d == 0) d = 1;
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    const uint32_t x = (uint32_t)(rng >> 33);
    bad += div31(x, make(d)) != x / d;
    total++;
  }
  printf("%lld %lld\n", bad, total);
  return bad != 0;
}
