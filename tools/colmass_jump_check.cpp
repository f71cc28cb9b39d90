// CPU check of the search of Layer.check_column_mass (reference: dpLGAR/models/physics/layers/Layer.py:681-701) as the
// CUDA path runs it: a host RESTATEMENT of the jumping loop of Column::check_column_mass (lgar-py_b200/csrc/
// lgar_device.cuh) -- doubling / halving probes on depths produced by advance_rounded() (the real header,
// lgar-py_b200/csrc/lgar_rounded.cuh), including the clause that crosses runs whose end points are both far from the
// stopping window (the FLAT case: the reference never leaves the loop) -- against the literal loop of the reference,
// on column-mass functions of the shape mass_balance() has around the free-drainage front:
//     mass(d) = c + (d - p) * theta_fd + (q - d) * theta_next          (slope theta_fd - theta_next >= 0)
// evaluated in floating point like the kernel does (every product and sum rounded).  For every case the two loops must
// agree on: capped or not; if not capped, the final depth (bit for bit) and the iteration count.
// Prints the number of cases and mismatches; exit code 1 on any mismatch.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include "../lgar-py_b200/csrc/lgar_rounded.cuh"

struct Mass {
  double c, p, q, th_fd, th_next;
  double operator()(double d) const {
    double m = 0.0 + c;
    m = m + (d - p) * th_fd;
    m = m + (q - d) * th_next;
    return m;
  }
};
struct Result {
  double depth;
  long long it;
  bool capped;
  long long evals;
};

// Layer.py:681-701, one step per iteration
static Result literal(const Mass& M, double depth, double target, long long cap) {
  double current_mass = M(depth);
  double err = fabs(current_mass - target);
  bool switched = false;
  double factor = 1.0;
  double depth_new = depth;
  Result r{depth, 0, false, 1};
  while (fabs(err - 1e-12) > 1e-12) {
    if (++r.it > cap) {
      r.capped = true;
      break;
    }
    if (current_mass < target) {
      depth_new = depth_new + 0.01 * factor;
      switched = false;
    } else {
      if (!switched) {
        switched = true;
        factor = factor * 0.001;
      }
      depth_new = depth_new - (0.01 * factor);
    }
    current_mass = M(depth_new);
    r.evals++;
    err = fabs(current_mass - target);
  }
  r.depth = depth_new;
  return r;
}

// Column::check_column_mass (lgar_device.cuh), values only
static Result jumping(const Mass& M, double depth, double target, long long cap) {
  double current_mass = M(depth);
  double err = fabs(current_mass - target);
  bool switched = false;
  double factor = 1.0;
  double depth_new = depth;
  long long it = 0;
  int run_len = 0;
  bool run_up = false;
  Result r{depth, 0, false, 1};
  while (fabs(err - 1e-12) > 1e-12) {
    if (++it > cap) {
      r.capped = true;
      break;
    }
    const bool up = current_mass < target;
    const double fac_before = factor;
    if (up) {
      depth_new = depth_new + 0.01 * factor;
      switched = false;
    } else {
      if (!switched) {
        switched = true;
        factor = factor * 0.001;
      }
      depth_new = depth_new - (0.01 * factor);
    }
    current_mass = M(depth_new);
    r.evals++;
    err = fabs(current_mass - target);
    if (up == run_up && factor == fac_before) run_len++;
    else {
      run_len = 1;
      run_up = up;
    }
    if (run_len >= 3 && fabs(err - 1e-12) > 1e-12 && (current_mass < target) == run_up &&
        fabs(depth_new) >= 64.0 * (0.01 * factor)) {
      run_len = 0;
      const double step = 0.01 * factor;
      long long stride = 4;
      bool shrinking = false;
      while (stride >= 2) {
        const double cand = lgar::advance_rounded(depth_new, up ? step : -step, stride);
        const double m = M(cand);
        r.evals++;
        const double e = fabs(m - target);
        const bool cont = (up ? (m < target) : !(m < target)) && (fabs(e - 1e-12) > 1e-12) && (fabs(cand) >= 64.0 * step) &&
                          ((cand < 0.0) == (depth_new < 0.0)) &&
                          (fabs(m - current_mass) >= 1e-13 * (double)stride || (e > 1e-11 && err > 1e-11) || cand == depth_new);
        if (cont) {
          depth_new = cand;
          current_mass = m;
          err = e;
          it += stride;
          if (it > cap) break;
          stride = shrinking ? (stride >> 1) : (stride << 1);
          if (stride > (1LL << 40)) stride = 1LL << 40;
        } else {
          shrinking = true;
          stride >>= 1;
        }
      }
    }
  }
  r.depth = depth_new;
  r.it = it;
  return r;
}

int main(int argc, char** argv) {
  const long ncases = argc > 1 ? atol(argv[1]) : 4000;
  const long long cap = argc > 2 ? atoll(argv[2]) : 200000;
  srand48(20261019);
  long bad = 0, capped = 0, flat = 0, converged = 0;
  long long ev_lit = 0, ev_jmp = 0, max_jmp_capped = 0;
  for (long i = 0; i < ncases; i++) {
    Mass M;
    M.p = 40.0 * drand48();
    const double d0 = M.p + 0.5 + 120.0 * drand48();
    M.q = d0 + 0.5 + 30.0 * drand48();
    M.th_fd = 0.30 + 0.18 * drand48();
    const int kind = (int)(i % 8);
    // slope of the mass in the depth: flat (theta equal: the reference's loop never ends), tiny, small, ordinary
    const double slope = kind == 0 ? 0.0 : (kind == 1 ? ldexp(1.0, -40 - (int)(lrand48() % 12)) : exp(log(1e-9) + drand48() * log(0.3e9)));
    M.th_next = M.th_fd - slope;
    M.c = 20.0 * drand48();
    // distance of the target from the starting mass: both signs, 1e-13 .. 3 cm (a flat or nearly flat column cannot get there)
    const double dist = exp(log(1e-13) + drand48() * log(3e13)) * ((lrand48() & 1) ? 1.0 : -1.0);
    const double target = M(d0) + dist;
    const Result a = literal(M, d0, target, cap);
    const Result b = jumping(M, d0, target, cap);
    ev_lit += a.evals;
    ev_jmp += b.evals;
    if (slope == 0.0) flat++;
    bool ok = a.capped == b.capped;
    if (ok && !a.capped) ok = (memcmp(&a.depth, &b.depth, 8) == 0) && a.it == b.it;
    if (a.capped) {
      capped++;
      if (b.evals > max_jmp_capped) max_jmp_capped = b.evals;
    } else converged++;
    if (!ok) {
      if (bad < 10)
        fprintf(stderr, "MISMATCH case %ld slope %g dist %g: literal depth %a it %lld capped %d | jumping depth %a it %lld capped %d\n", i, slope,
                dist, a.depth, a.it, (int)a.capped, b.depth, b.it, (int)b.capped);
      bad++;
    }
  }
  printf("cases %ld (flat %ld), converged %ld, capped %ld, mass evaluations literal %lld jumping %lld, most evaluations of a capped "
         "search with jumps %lld, mismatches %ld\n",
         ncases, flat, converged, capped, ev_lit, ev_jmp, max_jmp_capped, bad);
  return bad ? 1 : 0;
}
