/* Exhaustive check of the division shortcut used by k_edge_flat (ctp_device.cuh, ctp_div_small):
 * for all integers 1 <= idx < npts <= LIMIT,
 *     r  = 1.0 / npts                  (correctly rounded)
 *     q0 = idx * r
 *     q1 = fma(fma(-q0, npts, idx), r, q0)
 * equals the correctly rounded quotient (double)idx / npts that the reference computes
 * (src/base_map.cpp:65, "(double)index / num_points_path").  Prints the number of mismatches. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv) {
  const int limit = argc > 1 ? atoi(argv[1]) : 16384;
  long long bad = 0, total = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad, total)
  for (int n = 2; n <= limit; n++) {
    const double np = (double)n;
    const volatile double r = 1.0 / np;
    for (int i = 1; i < n; i++) {
      const double a = (double)i;
      const double q0 = a * r;
      const double q1 = fma(fma(-q0, np, a), r, q0);
      const volatile double want = a / np;
      bad += (q1 != want);
      total++;
    }
  }
  printf("%lld %lld\n", bad, total);
  return bad != 0;
}
