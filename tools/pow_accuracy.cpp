// CPU accuracy harness for lgar_pow.cuh: prints x y fast(x,y) glibc(x,y) as hex for random inputs
// in the domains the LGAR closures use.  Checked against mpmath by tools/pow_accuracy.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../lgar-py_b200/csrc/lgar_pow.cuh"
int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 100000;
  srand48(12345);
  long fallback = 0;
  for (long i = 0; i < n; i++) {
    double x, y;
    switch (i % 6) {
      case 0: x = exp(log(1e-6) + drand48() * log(1e10)); y = 1.1 + 1.9 * drand48(); break;      // (alpha h)^n
      case 1: x = 1.0 + exp(log(1e-12) + drand48() * log(1e16)); y = 1.0 - 1.0 / (1.1 + 1.9 * drand48()); break;  // (1+..)^m
      case 2: x = drand48(); y = 1.0 / (1.0 - 1.0 / (1.1 + 1.9 * drand48())); break;              // Se^(1/m)
      case 3: x = exp(log(1e-12) + drand48() * log(1e12)); y = 1.0 - 1.0 / (1.1 + 1.9 * drand48()); break;  // base^m
      case 4: x = drand48(); y = -1.0 / (1.0 - 1.0 / (1.1 + 1.9 * drand48())); break;             // Se^(-1/m)
      default: x = exp(log(1e-12) + drand48() * log(1e24)); y = 1.0 / (1.1 + 1.9 * drand48()); break;  // base^(1/n)
    }
    if (i % 97 == 0) x = 1.0 + (drand48() - 0.5) * 1e-3;
    if (i % 101 == 0) x = 1.0 - drand48() * 1e-9;
    double r;
    if (!lgar::pow_fast(x, y, &r)) { fallback++; r = pow(x, y); }
    printf("%a %a %a %a\n", x, y, r, pow(x, y));
  }
  fprintf(stderr, "fallbacks: %ld of %ld\n", fallback, n);
  return 0;
}
