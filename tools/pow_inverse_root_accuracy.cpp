// CPU harness for lgar::pow_inverse_root (lgar_pow.cuh): prints u m se shortcut full glibc as hex for random node
// arguments; checked against mpmath by tools/pow_inverse_root_accuracy.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../lgar-py_b200/csrc/lgar_pow.cuh"
int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 100000;
  srand48(777);
  long skipped = 0;
  for (long i = 0; i < n; i++) {
    const double nn = 1.1 + 1.9 * drand48();
    const double m = 1.0 - (1.0 / nn), inv_m = 1.0 / m;
    double p;
    switch (i % 4) {
      case 0: p = exp(log(1e-14) + drand48() * log(1e27)); break;   // (alpha h)^n over the whole range
      case 1: p = exp(log(1e-4) + drand48() * log(1e8)); break;
      case 2: p = drand48() * 1e-6; break;                           // nearly saturated
      default: p = exp(drand48() * log(1e13)); break;
    }
    const double u = 1.0 + p;
    const double xv[1] = {u}, yv[1] = {m};
    double d[1], lg[1], rs[1];
    bool ok[1];
    lgar::pow_core_v<1>(xv, yv, d, ok, lg, rs);
    if (!ok[0] || std::isnan(rs[0])) { skipped++; continue; }
    const double se = 1.0 / d[0];
    const double fast = lgar::pow_inverse_root(u, d[0], se, rs[0], lg[0], m, inv_m);
    double full;
    if (!lgar::pow_fast(se, inv_m, &full)) full = pow(se, inv_m);
    printf("%a %a %a %a %a %a\n", u, m, se, fast, full, pow(se, inv_m));
  }
  fprintf(stderr, "skipped: %ld of %ld\n", skipped, n);
  return 0;
}
