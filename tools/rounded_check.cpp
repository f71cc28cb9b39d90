// CPU check of lgar::advance_rounded (lgar-py_b200/csrc/lgar_rounded.cuh) against the literal chain of rounded
// additions it replaces.  Prints the number of cases and mismatches; exit code 1 on any mismatch.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../lgar-py_b200/csrc/lgar_rounded.cuh"

static double literal(double x, double s, long long k, bool* stayed_positive) {
  const bool neg = x < 0.0;
  *stayed_positive = true;
  for (long long i = 0; i < k; i++) {
    x = x + s;
    if (neg ? !(x < 0.0) : !(x > 0.0)) { *stayed_positive = false; return x; }
  }
  return x;
}

int main(int argc, char** argv) {
  const long ncases = argc > 1 ? atol(argv[1]) : 200000;
  srand48(2024);
  long bad = 0, checked = 0, left = 0;
  for (long i = 0; i < ncases; i++) {
    double x, s;
    long long k;
    switch (i % 8) {
      case 0: x = 2000.0 * drand48() + 0.5; s = -0.1 * pow(10.0, -(double)(lrand48() % 6)); k = 1 + lrand48() % 20000; break;  // psi runs
      case 1: x = 2000.0 * drand48() + 0.5; s = 0.1 * pow(10.0, -(double)(lrand48() % 6)); k = 1 + lrand48() % 20000; break;
      case 2: x = ldexp(1.0, (int)(lrand48() % 12) - 2) * (1.0 + 1e-9 * drand48()); s = -x * 1e-4 * drand48(); k = 1 + lrand48() % 30000; break;  // binade edges
      case 3: x = 200.0 * drand48() + 1e-3; s = (drand48() - 0.5) * 0.05; k = 1 + lrand48() % 5000; break;  // Geff node chains (dh of either sign)
      case 4: x = -(2000.0 * drand48() + 0.5); s = 0.01 * (drand48() - 0.3); k = 1 + lrand48() % 20000; break;  // negative values mirror
      case 5: x = 1.0 + drand48(); s = ldexp(1.0, -53) * (double)(1 + lrand48() % 7) * 0.5; k = 1 + lrand48() % 4000; break;  // half-ulp ties
      case 6: x = 100.0 * drand48() + 1.0; s = -0.01 * 0.001; k = 1 + lrand48() % 300000; break;  // check_column_mass steps
      default: x = exp(log(1e-6) + drand48() * log(1e12)); s = x * (drand48() - 0.5) * 1e-3; k = 1 + lrand48() % 10000; break;
    }
    bool pos;
    const double want = literal(x, s, k, &pos);
    const double got = lgar::advance_rounded(x, s, k);
    if (pos) {
      checked++;
      if (!(got == want)) {
        if (bad < 10) fprintf(stderr, "MISMATCH x=%a s=%a k=%lld: got %a want %a\n", x, s, k, got, want);
        bad++;
      }
    } else {  // the chain leaves the sign of x: the callers only rely on the result not being accepted
      left++;
      const bool rejected = (x < 0.0) ? !(got < 0.0) : !(got > 0.0);
      if (!rejected && !(got == want)) bad++;
    }
  }
  printf("cases %ld, exact comparisons %ld, chains that changed sign %ld, mismatches %ld\n", ncases, checked, left, bad);
  return bad ? 1 : 0;
}
