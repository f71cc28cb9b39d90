// =====================================================================================
// oracle/lgar_oracle.cpp -- CPU restatement of the dpLGAR time-stepping core.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lgar-py_b200/) may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
//
// Parity pinning: the Python reference has no tests or golden vectors of its own
// (SURVEY.md 8c), so this oracle is pinned against outputs of the reference itself,
// generated in the build container by tests/golden/make_golden.py (reference imported
// unmodified from /root/reference) and committed as tests/golden/*.npz.  See
// tests/test_oracle_vs_golden.py.
//
// The code deliberately mirrors the reference's data structures (one list of wetting
// fronts per layer + a `previous_state` snapshot per layer, fronts identified by object
// identity or by the value-equality `is_equal`) and its evaluation ORDER, including the
// quirks Q1-Q21 listed in SURVEY.md.  Every function cites the reference file:line it
// follows (paths relative to /root/reference/dpLGAR/).
//
// The scalar type is a template parameter: `double` for the plain forward pass and
// `Dual` (value + NT forward-mode tangents) to obtain d(outputs)/d(alpha,n,ksat) with
// the same local-derivative conventions as torch.autograd (Q13/Q14), which serves as an
// independent check of the hand-written reverse-mode CUDA kernel.
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>
#ifdef LGAR_ORACLE_DEVICE_POW
#include "../lgar-py_b200/csrc/lgar_pow.cuh"
#endif

#define LGAR_LMAX 8
#define LGAR_FMAX 16
#define LGAR_NOUT 10
#define LGAR_NGIUH 8

// status codes (same numbering as include/lgar_b200.h)
enum {
  ST_OK = 0,
  ST_NEG_POW = 1,         // ValueError: negative base in safe_pow (physics/utils.py:25-27)
  ST_NAN = 2,             // ValueError: NaN in pow input / result (physics/utils.py:17-19,181-183)
  ST_THETA_ORDER = 3,     // ValueError: theta_1 > theta_2 in layer 0 (Layer.py:1206-1208)
  ST_BOTTOM_REACHED = 4,  // AttributeError in recalibrate on the last layer (Layer.py:980, Q9)
  ST_NULL_NEIGHBOUR = 5,  // AttributeError/UnboundLocalError on a missing neighbour (Q10)
  ST_FRONT_OVERFLOW = 6,  // more than LGAR_FMAX fronts (capacity of the dump; not a reference state)
  ST_ITER_CAP = 7,        // a root finder exceeded the iteration cap (reference would spin)
  ST_INDEX_ERROR = 8,     // IndexError (Layer.py:1115 or list index)
};

struct RefError {
  int code;
};

extern "C" {
typedef struct {
  int32_t num_layers;
  int32_t nint;
  int32_t num_subcycles;
  int32_t num_giuh;
  double dt_h;  // subcycle_length_h
  double initial_psi;
  double wilting_point_psi;
  double ponded_depth_max;
  double frozen_factor;
  double thickness[LGAR_LMAX];
  double theta_r[LGAR_LMAX];
  double theta_e[LGAR_LMAX];
  double alpha[LGAR_LMAX];
  double n[LGAR_LMAX];
  double ksat[LGAR_LMAX];  // already multiplied by frozen_factor (models/dpLGAR.py:57)
  double giuh[LGAR_NGIUH];
  int64_t iter_cap;  // cap for the two root finders (0 -> default 2,000,000)
  int64_t use_closed_form_G;  // cfg.data.use_closed_form_G (green_ampt.py:85-98)
} lgar_oracle_cfg;
}

// ------------------------------------------------------------------------------------
// scalar abstraction
// ------------------------------------------------------------------------------------
static thread_local long long g_cnt[12];  // 0 geff, 1 theta_from_h, 2 h_from_se, 3 k_from_se, 4 se_from_h, 5 rootfind iters, 6 colmass iters,
                                         // 7 unused; branch coverage: 8 dry-over-wet fixes (A17), 9 insert_water equality (Q8),
                                         // 10 calc_bottom_sum_f_p with the free-drainage front in layer >= 2 (Q18), 11 domain-boundary pops (A16)

#ifndef LGAR_NT
#define LGAR_NT 9
#endif

struct Dual {
  double v;
  double d[LGAR_NT];
  Dual() : v(0.0) { std::memset(d, 0, sizeof(d)); }
  Dual(double x) : v(x) { std::memset(d, 0, sizeof(d)); }
};

static inline double val(double x) { return x; }
static inline double val(const Dual& x) { return x.v; }

static inline Dual operator+(const Dual& a, const Dual& b) {
  Dual r; r.v = a.v + b.v;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
static inline Dual operator-(const Dual& a, const Dual& b) {
  Dual r; r.v = a.v - b.v;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
static inline Dual operator-(const Dual& a) {
  Dual r; r.v = -a.v;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = -a.d[i];
  return r;
}
static inline Dual operator*(const Dual& a, const Dual& b) {
  Dual r; r.v = a.v * b.v;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
static inline Dual operator/(const Dual& a, const Dual& b) {
  // torch: grad_self = g / other ; grad_other = -g * self / (other*other)
  Dual r; r.v = a.v / b.v;
  double ia = 1.0 / b.v, ib = -a.v / (b.v * b.v);
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] * ia + b.d[i] * ib;
  return r;
}
static inline Dual operator+(const Dual& a, double b) { Dual r = a; r.v = a.v + b; return r; }
static inline Dual operator+(double a, const Dual& b) { Dual r = b; r.v = a + b.v; return r; }
static inline Dual operator-(const Dual& a, double b) { Dual r = a; r.v = a.v - b; return r; }
static inline Dual operator-(double a, const Dual& b) { Dual r = -b; r.v = a - b.v; return r; }
static inline Dual operator*(const Dual& a, double b) {
  Dual r; r.v = a.v * b;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] * b;
  return r;
}
static inline Dual operator*(double a, const Dual& b) { return b * a; }
static inline Dual operator/(const Dual& a, double b) {
  Dual r; r.v = a.v / b;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] / b;
  return r;
}
static inline Dual operator/(double a, const Dual& b) { return Dual(a) / b; }

// torch.pow(base, exponent) with torch's backward conventions
// (pow_backward_self: 0 where exponent==0 ; pow_backward_exponent: 0 where base==0 && exponent>=0)
#ifdef LGAR_ORACLE_DEVICE_POW
// Second build of the oracle (liblgar_oracle_devpow.so, tests only): pow through the CUDA path's own table-driven
// routine (lgar-py_b200/csrc/lgar_pow.cuh compiles for the host with the same operations) instead of glibc's.  The two
// pows differ in the last bit in ~0.035 % of calls; LGAR's discrete decisions (front creation, merging, root-finder
// exits) can amplify one such bit into a different trajectory for an ill-conditioned column.  With this build the
// oracle and the CUDA kernels execute the same arithmetic, so every remaining difference would be a logic defect.
static inline double pow_(double a, double b) {
  if (b == 2.0) return a * a;  // the CUDA closures square by multiplication (== glibc's pow(x, 2.0) bit for bit)
  double r;
  return lgar::pow_fast(a, b, &r) ? r : std::pow(a, b);
}
#else
static inline double pow_(double a, double b) { return std::pow(a, b); }
#endif
static inline Dual pow_(const Dual& a, const Dual& b) {
  Dual r; r.v = pow_(a.v, b.v);
  double da = (b.v == 0.0) ? 0.0 : b.v * std::pow(a.v, b.v - 1.0);
  double db = (a.v == 0.0 && b.v >= 0.0) ? 0.0 : r.v * std::log(a.v);
  for (int i = 0; i < LGAR_NT; i++) {
    double t = 0.0;
    if (a.d[i] != 0.0) t += a.d[i] * da;
    if (b.d[i] != 0.0) t += b.d[i] * db;
    r.d[i] = t;
  }
  return r;
}
static inline Dual pow_(const Dual& a, double b) { return pow_(a, Dual(b)); }
static inline double sqrt_(double a) { return std::sqrt(a); }
static inline Dual sqrt_(const Dual& a) {
  Dual r; r.v = std::sqrt(a.v);
  double g = 1.0 / (2.0 * r.v);
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = (a.d[i] != 0.0) ? a.d[i] * g : 0.0;
  return r;
}
static inline double abs_(double a) { return std::fabs(a); }
static inline Dual abs_(const Dual& a) {
  // torch abs backward: grad * sgn(self)  (0 at 0)
  Dual r; r.v = std::fabs(a.v);
  double s = (a.v > 0.0) ? 1.0 : ((a.v < 0.0) ? -1.0 : 0.0);
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = a.d[i] * s;
  return r;
}
// torch.min / torch.minimum (binary) PROPAGATE NaN (unlike fmin): calc_dry_depth hands torch.min a NaN when the top
// front sits at theta_e exactly (Layer.py:1318-1333: delta_theta = 0, tau = inf, geff = 0, tau * geff = NaN) -- golden
// nan_dry_depth_col185
static inline double min_(double a, double b) {
  if (std::isnan(a) || std::isnan(b)) return a + b;
  return (b < a) ? b : a;
}
static inline Dual min_(const Dual& a, const Dual& b) {
  // gradient split 0.5/0.5 at ties (Q13)
  if (std::isnan(a.v) || std::isnan(b.v)) return Dual(a.v + b.v);
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  Dual r; r.v = a.v;
  for (int i = 0; i < LGAR_NT; i++) r.d[i] = 0.5 * a.d[i] + 0.5 * b.d[i];
  return r;
}
static inline Dual min_(const Dual& a, double b) { return min_(a, Dual(b)); }
static inline Dual min_(double a, const Dual& b) { return min_(Dual(a), b); }
// torch.clamp(x, min=lo): gradient 1 where x >= lo (boundary included, Q13)
static inline double clamp_min_(double a, double lo) { return (a < lo) ? lo : a; }
static inline Dual clamp_min_(const Dual& a, double lo) {
  if (a.v >= lo) return a;
  return Dual(lo);
}
// torch.clamp(x, min=lo, max=hi): value min(max(x,lo),hi); gradient 1 where lo <= x <= hi
static inline double clamp_(double a, double lo, double hi) {
  double r = (a < lo) ? lo : a;
  return (r > hi) ? hi : r;
}
static inline Dual clamp_(const Dual& a, double lo, double hi) {
  double r = (a.v < lo) ? lo : a.v;
  r = (r > hi) ? hi : r;
  if (a.v >= lo && a.v <= hi) { Dual o = a; o.v = r; return o; }
  return Dual(r);
}

// ------------------------------------------------------------------------------------
// van Genuchten closures -- physics/utils.py
// ------------------------------------------------------------------------------------
template <class T>
static inline T safe_pow(const T& base, const T& e) {  // utils.py:12-32
  if (std::isnan(val(base)) || std::isnan(val(e))) throw RefError{ST_NAN};
  if (val(base) < 0) throw RefError{ST_NEG_POW};
  return pow_(base, e);
}
template <class T>
static inline T error_check(const T& r) {  // utils.py:177-185
  if (std::isnan(val(r))) throw RefError{ST_NAN};
  return r;
}
template <class T>
struct Soil {  // one layer's parameter set: Layer.attributes + alpha/n/ksat (Layer.py:48-57)
  T alpha, n, ksat, m;
  T theta_e, theta_r;
  T bc_lambda, bc_psib;  // Brooks-Corey estimates (utils.py:54-99), only read by the closed-form Geff
};

template <class T>
static T theta_from_h(const T& h, const Soil<T>& s) {  // utils.py:35-51
  g_cnt[1]++;
  T alpha_pow = safe_pow(s.alpha * h, s.n);
  T outer = safe_pow(1.0 + alpha_pow, s.m);
  T result = (1.0 / outer * (s.theta_e - s.theta_r)) + s.theta_r;
  return error_check(result);
}
template <class T>
static T se_from_theta(const T& theta, const Soil<T>& s) {  // utils.py:102-112
  return error_check((theta - s.theta_r) / (s.theta_e - s.theta_r));
}
template <class T>
static T se_from_h(const T& h, const Soil<T>& s) {  // utils.py:115-131
  g_cnt[4]++;
  if (std::fabs(val(h)) < 1.0e-01) return T(1.0);  // constant: zero gradient (Q12)
  T internal = safe_pow(s.alpha * h, s.n);
  T result = 1.0 / safe_pow(1.0 + internal, s.m);
  return error_check(result);
}
static inline bool isclose_zero_1e8(double base) {
  // torch.isclose(base, 0, threshold): |base - 0| <= atol(1e-8) + rtol*|0|  (Q2)
  return std::fabs(base) <= 1e-8;
}
template <class T>
static T k_from_se(const T& se, const T& ksat, const T& m) {  // utils.py:134-156
  g_cnt[3]++;
  T se_pow = safe_pow(se, 1.0 / m);
  T base = 1.0 - se_pow;
  if (isclose_zero_1e8(val(base))) base = base + 1e-12;
  T outside = safe_pow(base, m);
  T result = ksat * sqrt_(se) * safe_pow(1.0 - outside, T(2.0));
  return error_check(result);
}
template <class T>
static T h_from_se(const T& se, const Soil<T>& s) {  // utils.py:159-174
  g_cnt[2]++;
  T se_pow = safe_pow(se, (-1.0 / s.m));
  T base = se_pow - 1.0;
  if (isclose_zero_1e8(val(base))) base = base + 1e-12;
  T outside = safe_pow(base, (1.0 / s.n));
  T result = 1.0 / s.alpha * outside;
  return error_check(result);
}

// ------------------------------------------------------------------------------------
// Geff -- physics/lgar/green_ampt.py:19-99 (trapezoid branch; use_closed_form_G=False)
// ------------------------------------------------------------------------------------
// closed form (use_closed_form_G=True, green_ampt.py:85-98).  Kept literally, including the swapped roles of
// theta_1 / theta_2 and the operator precedence of the published line (sic):
//   geff = h_c * se_i^e - se_f^e / (1 - se_f^e),  e = 3 + 1/lambda;  inf or NaN -> h_c
template <class T>
static T calc_geff_closed(const T& theta_1, const T& theta_2, const Soil<T>& s) {
  T se_f = se_from_theta(theta_1, s);
  T se_i = se_from_theta(theta_2, s);
  T h_c = s.bc_psib * (2.0 + 3.0 * s.bc_lambda) / (1.0 + 3.0 * s.bc_lambda);
  T e = 3.0 + 1.0 / s.bc_lambda;
  T geff = h_c * safe_pow(se_i, e) - safe_pow(se_f, e) / (1.0 - safe_pow(se_f, e));
  if (std::isinf(val(geff)) || std::isnan(val(geff))) geff = h_c;
  return geff;
}
template <class T>
static T calc_geff(const T& theta_1, const T& theta_2, const Soil<T>& s, int nint) {
  if (nint < 0) return calc_geff_closed(theta_1, theta_2, s);  // nint < 0 encodes use_closed_form_G (see init)
  g_cnt[0]++;
  T se_i = se_from_theta(theta_1, s);
  T se_f = se_from_theta(theta_2, s);
  T h_i = h_from_se(se_i, s);
  T h_f = h_from_se(se_f, s);
  (void)se_from_h(h_i, s);  // "Checkpoint" calls (green_ampt.py:61-63): only their guards matter
  (void)se_from_h(h_f, s);
  T dh = (h_f - h_i) / (double)nint;
  T geff(0.0);
  T k1 = k_from_se(se_i, s.ksat, s.m);
  T h2 = h_i + dh;
  for (int i = 0; i < nint; i++) {
    T se2 = se_from_h(h2, s);
    T k2 = k_from_se(se2, s.ksat, s.m);
    geff = geff + ((k1 + k2) * (dh / 2.0));
    k1 = k2;
    h2 = h2 + dh;
  }
  return abs_(geff / s.ksat);
}

// ------------------------------------------------------------------------------------
// column state
// ------------------------------------------------------------------------------------
template <class T>
struct Front {  // layers/WettingFront.py:18-49
  T depth, theta, psi, k, dzdt;
  int layer_num;
  bool to_bottom;
  bool is_equal(const Front& o) const {  // WettingFront.py:76-84 (value equality, Q3)
    return val(o.depth) == val(depth) && val(o.psi) == val(psi) && val(o.dzdt) == val(dzdt);
  }
};

template <class T>
struct Column;

template <class T>
struct Layer {
  int l;
  Soil<T> s;
  double thick, cum;
  std::vector<Front<T>*> wf;    // Layer.wetting_fronts
  std::vector<Front<T>> prev;   // Layer.previous_state (value copies)
};

template <class T>
struct Nb {  // get_neighboring_fronts / get_extended_neighbors
  Front<T>* cur = nullptr;
  const Front<T>* pcur = nullptr;
  Front<T>* next = nullptr;
  const Front<T>* pnext = nullptr;
  Front<T>* n2n = nullptr;
};

template <class T>
struct Column {
  lgar_oracle_cfg cfg;
  int L;
  std::vector<Layer<T>> layers;
  std::vector<std::unique_ptr<Front<T>>> pool;  // owns every front ever created
  Front<T>* fd = nullptr;                       // wf_free_drainage_demand (object reference)
  T ending_volume, ponded_water, previous_precip;
  T precip, PET, AET, infiltration, runoff, percolation, giuh_runoff, discharge;
  T giuh_queue[LGAR_NGIUH];
  long long iter_cap;

  Front<T>* new_front() {
    pool.emplace_back(new Front<T>());
    return pool.back().get();
  }

  // ---- construction: dpLGAR.set_internal_states (models/dpLGAR.py:97-147),
  //      generate_soil_metrics (data/utils.py:40-105), Layer.__init__ (Layer.py:22-90)
  void init(const lgar_oracle_cfg& c, const T* alpha, const T* n, const T* ksat) {
    cfg = c;
    if (c.use_closed_form_G) cfg.nint = -1;  // every calc_geff call site passes cfg.nint
    L = c.num_layers;
    iter_cap = c.iter_cap > 0 ? c.iter_cap : 2000000;
    layers.resize(L);
    double cum = 0.0;
    for (int l = 0; l < L; l++) {
      Layer<T>& ly = layers[l];
      ly.l = l;
      ly.s.alpha = alpha[l];
      ly.s.n = n[l];
      ly.s.ksat = ksat[l];
      ly.s.m = 1.0 - (1.0 / n[l]);  // calc_m utils.py:72-74
      {  // calc_bc_lambda / calc_bc_psib (utils.py:54-99), as generate_soil_metrics computes them (data/utils.py:85-87)
        T p_ = 1.0 + (2.0 / ly.s.m);
        ly.s.bc_lambda = 2.0 / (p_ - 3.0);
        ly.s.bc_psib = (p_ + 3.0) * (147.8 + 8.1 * p_ + 0.092 * p_ * p_) /
                       (2.0 * ly.s.alpha * p_ * (p_ - 1.0) * (55.6 + 7.4 * p_ + p_ * p_));
        if (c.use_closed_form_G) error_check(ly.s.bc_psib);
      }
      ly.s.theta_e = T(c.theta_e[l]);
      ly.s.theta_r = T(c.theta_r[l]);
      ly.thick = c.thickness[l];
      cum = (l == 0) ? c.thickness[0] : cum + c.thickness[l];  // GlobalParams.py:106-110
      ly.cum = cum;
      T theta_init = theta_from_h(T(c.initial_psi), ly.s);  // data/utils.py:82-84
      Front<T>* f = new_front();                            // WettingFront.__init__
      f->depth = T(ly.cum);
      f->layer_num = l;
      f->theta = theta_init;
      f->dzdt = T(0.0);
      T se = se_from_theta(f->theta, ly.s);
      f->psi = T(c.initial_psi);
      f->k = k_from_se(se, ly.s.ksat, ly.s.m);
      f->to_bottom = true;
      ly.wf.push_back(f);
    }
    copy_states();
    ending_volume = mass_balance();
    ponded_water = T(0.0);
    previous_precip = T(0.0);
    precip = PET = AET = infiltration = runoff = percolation = giuh_runoff = discharge = T(0.0);
    for (int i = 0; i < LGAR_NGIUH; i++) giuh_queue[i] = T(0.0);
  }

  void copy_states() {  // Layer.copy_states / deepcopy (Layer.py:110-132)
    for (auto& ly : layers) {
      ly.prev.clear();
      for (auto* f : ly.wf) ly.prev.push_back(*f);
    }
  }
  Front<T>* first_front(int l) {  // layers[l].wetting_fronts[0]; IndexError on an empty list
    if (layers[l].wf.empty()) throw RefError{ST_INDEX_ERROR};
    return layers[l].wf[0];
  }
  int num_fronts() const {  // Layer.calc_num_wetting_fronts (Layer.py:171-175)
    int n = 0;
    for (auto& ly : layers) n += (int)ly.wf.size();
    return n;
  }
  int len_layers(const Layer<T>& ly) const {  // Layer.get_len_layers (Layer.py:1145-1155)
    return (ly.l < L - 1) ? (int)ly.wf.size() : (int)ly.wf.size() - 1;
  }

  // ---- Layer.mass_balance (Layer.py:795-824).  Association: S0 + (S1 + (S2 ...)).
  T mass_balance_from(int l) {
    Layer<T>& ly = layers[l];
    T sum(0.0);
    double base = (l == 0) ? 0.0 : (ly.cum - ly.thick);
    int nf = (int)ly.wf.size();
    if (nf > 1) {
      for (int i = 0; i < nf - 1; i++)
        sum = sum + (ly.wf[i]->depth - base) * (ly.wf[i]->theta - ly.wf[i + 1]->theta);
      sum = sum + (ly.wf[nf - 1]->depth - base) * ly.wf[nf - 1]->theta;
    } else {
      if (nf < 1) throw RefError{ST_INDEX_ERROR};
      sum = sum + (ly.wf[0]->depth - base) * ly.wf[0]->theta;
    }
    if (l < L - 1) return sum + mass_balance_from(l + 1);
    return sum;
  }
  T mass_balance() { return mass_balance_from(0); }

  // ---- free-drainage front: models/dpLGAR.py:328-338 + Layer.py:134-162
  Front<T>* calc_wetting_front_free_drainage() {
    Front<T>* w = first_front(0);
    T psi = w->psi;
    for (auto& ly : layers) {
      for (auto* cf : ly.wf) {
        if (val(cf->psi) <= val(psi)) {
          psi = cf->psi;
          w = cf;
        } else {
          // torch.isclose(cf.psi, psi, atol=1e-8): |a-b| <= 1e-8 + 1e-5*|b|
          double a = val(cf->psi), b = val(psi);
          if (std::isfinite(a) && std::isfinite(b) ? (std::fabs(a - b) <= 1e-8 + 1e-5 * std::fabs(b))
                                                   : (a == b)) {
            psi = cf->psi;
            w = cf;
          }
        }
      }
    }
    return w;
  }

  // ---- neighbour lookups: Layer.py:703-758
  Nb<T> neighbors(Layer<T>& ly, int i) {
    Nb<T> nb;
    int nf = (int)ly.wf.size();
    if (i < 0 || i >= nf) throw RefError{ST_INDEX_ERROR};
    nb.cur = ly.wf[i];
    if (i >= (int)ly.prev.size()) throw RefError{ST_INDEX_ERROR};
    nb.pcur = &ly.prev[i];
    if (i < nf - 1) {
      nb.next = ly.wf[i + 1];
      if (i + 1 >= (int)ly.prev.size()) throw RefError{ST_INDEX_ERROR};
      nb.pnext = &ly.prev[i + 1];
    }
    if (ly.l < L - 1 && i == nf - 1) {
      nb.next = first_front(ly.l + 1);
      nb.pnext = &ly.prev[0];  // sic (Q4): previous_state[0] of the SAME layer
    }
    return nb;
  }
  Nb<T> ext_neighbors(Layer<T>& ly, int i) {  // Layer.py:733-758
    Nb<T> nb = neighbors(ly, i);
    int nf = (int)ly.wf.size();
    if (i < nf - 2) {
      nb.n2n = ly.wf[i + 2];
    } else if (nb.next != nullptr) {
      if (nb.next->layer_num != nb.cur->layer_num) {
        if (ly.l >= L - 1) throw RefError{ST_NULL_NEIGHBOUR};  // self.next_layer is None
        Layer<T>& nl = layers[ly.l + 1];
        if (nl.wf.size() > 1) nb.n2n = nl.wf[1];
        else if (nl.l < L - 1) nb.n2n = first_front(nl.l + 1);
      } else {
        if (ly.l < L - 1) nb.n2n = first_front(ly.l + 1);
      }
    }
    return nb;
  }

  // ---- AET: Layer.calc_aet (Layer.py:760-783) -> calc_aet (lgar/aet.py:17-51)
  T calc_aet(double pet, double dt) {
    const Soil<T>& s = layers[0].s;
    const T& psi_cm = first_front(0)->psi;
    T theta_fc = (s.theta_e - s.theta_r) * 0.75 + s.theta_r;  // GlobalParams.py:75
    T wp_head_theta = theta_from_h(T(cfg.wilting_point_psi), s);
    T theta_wp = (theta_fc - wp_head_theta) * 0.5 + wp_head_theta;
    T se = se_from_theta(theta_wp, s);
    T psi_wp = h_from_se(se, s);
    T h_ratio = 1.0 + safe_pow(psi_cm / psi_wp, T(3.0));
    T aet_ = pet * (1.0 / h_ratio) * dt;
    return clamp_(aet_, 0.0, pet);  // upper clamp is the RATE pet (sic)
  }

  // ---- root finder: Layer.theta_mass_balance (Layer.py:242-318) + recalculate_mass (:211-240)
  T theta_mass_balance(Layer<T>& ly, T psi_cm, T new_mass, T prior_mass, const std::vector<T>& dtheta,
                       const std::vector<T>& dthick) {
    const double tol = 1e-12;
    T delta_mass = abs_(new_mass - prior_mass);
    bool switched = false;
    double factor = 1.0;
    T theta(0.0);
    T psi_prev = psi_cm;
    T delta_mass_prev = delta_mass;
    int count_no_mass_change = 0;
    if (val(delta_mass) <= tol) return theta_from_h(psi_cm, ly.s);
    long long it = 0;
    int nupper = (int)dthick.size() - 1;
    while (val(delta_mass) > tol) {
      if (++it > iter_cap) throw RefError{ST_ITER_CAP};
      g_cnt[5]++;
      if (val(new_mass) > val(prior_mass)) {
        psi_cm = psi_cm + (0.1 * factor);
        switched = false;
      } else {
        if (!switched) {
          switched = true;
          factor = factor * 0.1;
        }
        psi_prev = psi_cm;
        psi_cm = psi_cm - (0.1 * factor);
        if (val(psi_cm) < 0 && val(psi_prev) != 0) psi_cm = psi_prev * 0.1;
      }
      theta = theta_from_h(psi_cm, ly.s);
      T mass_layers = T(0.0) + (dthick[ly.l] * (theta - dtheta[ly.l]));
      for (int k = 0; k < nupper; k++) {  // recalculate_mass over layers 0..len-2
        T theta_layer = theta_from_h(psi_cm, layers[k].s);
        mass_layers = mass_layers + dthick[k] * (theta_layer - dtheta[k]);
      }
      new_mass = mass_layers;
      delta_mass = abs_(new_mass - prior_mass);
      if (std::fabs(val(psi_cm) - val(psi_prev)) < 1e-15 && factor < 1e-13) break;
      if (std::fabs(val(delta_mass) - val(delta_mass_prev)) < 1e-15) count_no_mass_change++;
      else count_no_mass_change = 0;
      if (count_no_mass_change == 5) break;
      if (val(psi_cm) <= 0 && val(psi_prev) < 1e-50) break;
      delta_mass_prev = delta_mass;
    }
    return theta;
  }

  // ---- Layer.base_case (Layer.py:320-387) + populate_delta_thickness (:177-209)
  void base_case(Layer<T>& ly, const T& percolation, const T& aet, double dt, Nb<T>& nb) {
    Front<T>* cur = nb.cur;
    const Front<T>* pcur = nb.pcur;
    cur->depth = cur->depth + cur->dzdt * dt;
    std::vector<T> dtheta(L, T(0.0)), dthick(L, T(0.0));
    T psi_old = pcur->psi;
    T psi_cm = cur->psi;
    double base = (ly.l > 0) ? layers[ly.l - 1].cum : 0.0;
    T prior_mass = (pcur->depth - base) * (pcur->theta - 0.0);
    T new_mass = (cur->depth - base) * (cur->theta - 0.0);
    // populate_delta_thickness starts at the top layer; needs a next layer (Q10)
    int k = 0;
    while (true) {
      Layer<T>& lk = layers[k];
      T theta_old = theta_from_h(psi_old, lk.s);
      prior_mass = prior_mass + lk.thick * (theta_old - 0.0);
      T theta = theta_from_h(psi_cm, lk.s);
      new_mass = new_mass + lk.thick * (theta - 0.0);
      dtheta[k] = T(0.0);
      dthick[k] = T(lk.thick);
      if (k + 1 >= L) throw RefError{ST_NULL_NEIGHBOUR};
      if (k + 1 < L - 1) k++;
      else break;
    }
    dthick[ly.l] = cur->depth - base;
    if (fd->layer_num == ly.l) prior_mass = prior_mass + percolation - (0.0 + aet);
    T theta_new = theta_mass_balance(ly, psi_cm, new_mass, prior_mass, dtheta, dthick);
    cur->theta = min_(theta_new, ly.s.theta_e);
    T se = se_from_theta(cur->theta, ly.s);
    cur->psi = h_from_se(se, ly.s);
  }

  // ---- Layer.deepest_layer_front (Layer.py:389-418)
  void deepest_layer_front(Layer<T>& ly, Nb<T>& nb) {
    if (!nb.next) throw RefError{ST_NULL_NEIGHBOUR};
    nb.cur->theta = theta_from_h(nb.next->psi, ly.s);
    nb.cur->psi = nb.next->psi;
  }

  // ---- Layer.wetting_front_in_layer (Layer.py:420-547) + compute_wetting_front_mass (:561-644)
  void wetting_front_in_layer(Layer<T>& ly, const T& infiltration, const T& aet, Nb<T>& nb, double dt) {
    Front<T>* cur = nb.cur;
    Front<T>* next = nb.next;
    const Front<T>* pcur = nb.pcur;
    const Front<T>* pnext = nb.pnext;
    if (!next || !pnext) throw RefError{ST_NULL_NEIGHBOUR};
    if (ly.l == 0) {
      T prior_mass = pcur->depth * (pcur->theta - pnext->theta);
      if (fd->is_equal(*cur)) prior_mass = prior_mass + (infiltration - (0.0 + aet));
      cur->depth = cur->depth + (cur->dzdt * dt);
      cur->depth = min_(cur->depth, layers[L - 1].cum);
      bool zero_dzdt = std::fabs(val(cur->dzdt)) <= 1e-8;  // isclose(dzdt, 0, rtol=1e-8) (Q2)
      if (zero_dzdt && cur->to_bottom == false) {
        // a new front was just created: leave theta
      } else {
        T potential_theta = (prior_mass / cur->depth) + next->theta;
        cur->theta = min_(ly.s.theta_e, potential_theta);
      }
    } else {
      double plt = layers[ly.l - 1].cum;
      cur->depth = cur->depth + (cur->dzdt * dt);
      T psi_old = pcur->psi, psi_below_old = pnext->psi;
      T psi_cm = cur->psi, psi_below = next->psi;
      T prior_mass = (pcur->depth - plt) * (pcur->theta - pnext->theta);
      T new_mass = (cur->depth - plt) * (cur->theta - next->theta);
      std::vector<T> dtheta(ly.l + 1, T(0.0)), dthick(ly.l + 1, T(0.0));
      for (int k = 0; k < ly.l; k++) {  // compute_wetting_front_mass, layers above
        Layer<T>& lk = layers[k];
        T theta_old = theta_from_h(psi_old, lk.s);
        T theta_below_old = theta_from_h(psi_below_old, lk.s);
        T local_delta_old = theta_old - theta_below_old;
        double layer_thickness = lk.cum - 0.0;  // sic (Q5): cumulative thickness
        prior_mass = prior_mass + (layer_thickness * local_delta_old);
        T theta = theta_from_h(psi_cm, lk.s);
        T theta_below = theta_from_h(psi_below, lk.s);
        new_mass = new_mass + (layer_thickness * (theta - theta_below));
        dtheta[k] = theta_below;
        dthick[k] = T(layer_thickness);
      }
      dtheta[ly.l] = next->theta;
      dthick[ly.l] = cur->depth - plt;
      if (fd->is_equal(*cur)) prior_mass = prior_mass + infiltration - (0.0 + aet);
      T theta_new = theta_mass_balance(ly, psi_cm, new_mass, prior_mass, dtheta, dthick);
      cur->theta = min_(theta_new, ly.s.theta_e);
    }
    T se = se_from_theta(cur->theta, ly.s);
    cur->psi = h_from_se(se, ly.s);
  }

  // ---- Layer.check_column_mass (Layer.py:655-701)
  void check_column_mass(const T& old_mass, const T& percolation, const T& aet) {
    const T& theta_e_k1 = layers[fd->layer_num].s.theta_e;
    T mass_timestep = (old_mass + percolation) - (aet + 0.0);
    if (std::fabs(val(fd->theta) - val(theta_e_k1)) < 1e-12) {
      T current_mass = mass_balance();
      T err = abs_(current_mass - mass_timestep);
      bool switched = false;
      double factor = 1.0;
      T depth_new = fd->depth;
      long long it = 0;
      while (std::fabs(val(err) - 1e-12) > 1e-12) {
        if (++it > iter_cap) throw RefError{ST_ITER_CAP};
        g_cnt[6]++;
        if (val(current_mass) < val(mass_timestep)) {
          depth_new = depth_new + 0.01 * factor;
          switched = false;
        } else {
          if (!switched) {
            switched = true;
            factor = factor * 0.001;
          }
          depth_new = depth_new - (0.01 * factor);
        }
        fd->depth = depth_new;
        current_mass = mass_balance();
        err = abs_(current_mass - mass_timestep);
      }
    }
  }

  // ---- Layer.move_wetting_fronts (Layer.py:1254-1307): deepest -> top sweep
  void move_wetting_fronts(const T& infiltration, const T& aet, const T& old_mass, int num_wf, double dt) {
    int count = num_wf;
    for (int l = L - 1; l >= 0; l--) {
      Layer<T>& ly = layers[l];
      bool is_bottom = (l == L - 1);
      int nf = (int)ly.wf.size();
      for (int i = nf - 1; i >= 0; i--) {
        Nb<T> nb = neighbors(ly, i);
        if (count < num_wf) {
          if (nb.cur->is_equal(*ly.wf.back())) deepest_layer_front(ly, nb);
          else wetting_front_in_layer(ly, infiltration, aet, nb, dt);
        }
        if (num_wf == L && is_bottom) base_case(ly, infiltration, aet, dt, nb);
        if (count == 1) check_column_mass(old_mass, infiltration, aet);
        count--;
      }
    }
  }

  // ---- merging: Layer.merge_wetting_fronts (:838-867), is_passing (:826-836), pass_front (:869-886)
  void merge_wetting_fronts() {
    for (auto& ly : layers) {
      int lf = len_layers(ly);
      for (int i = 0; i < lf; i++) {
        Nb<T> e = ext_neighbors(ly, i);
        if (!e.next) throw RefError{ST_NULL_NEIGHBOUR};
        bool passing = (val(e.cur->depth) > val(e.next->depth)) && (e.cur->layer_num == e.next->layer_num) &&
                       !e.next->to_bottom;
        if (passing) {
          if (!e.n2n) throw RefError{ST_NULL_NEIGHBOUR};  // Q10
          Front<T>* cur = e.cur;
          T mass = cur->depth * (cur->theta - e.next->theta) + e.next->depth * (e.next->theta - e.n2n->theta);
          cur->depth = mass / (cur->theta - e.n2n->theta);
          T se = se_from_theta(cur->theta, ly.s);
          cur->psi = h_from_se(se, ly.s);
          cur->k = k_from_se(se, ly.s.ksat, ly.s.m);
          for (size_t j = 0; j < ly.wf.size(); j++) {  // delete_front (:888-892): value equality
            if (ly.wf[j]->is_equal(*e.next)) {
              ly.wf.erase(ly.wf.begin() + j);
              break;
            }
          }
          break;
        }
      }
    }
  }

  // ---- layer-boundary crossing: Layer.py:894-1008
  void wetting_fronts_cross_layer_boundary() {
    for (auto& ly : layers) {
      int lf = len_layers(ly);
      for (int i = 0; i < lf; i++) {
        Nb<T> e = ext_neighbors(ly, i);
        Front<T>* cur = e.cur;
        Front<T>* next = e.next;
        if (!next) throw RefError{ST_NULL_NEIGHBOUR};
        bool deeper = val(cur->depth) > ly.cum;
        bool next_at_boundary = val(next->depth) == ly.cum;
        if (deeper && next_at_boundary) {
          T overshot = cur->depth - next->depth;
          T se = se_from_theta(cur->theta, ly.s);
          cur->psi = h_from_se(se, ly.s);
          cur->k = k_from_se(se, ly.s.ksat, ly.s.m);
          // recalibrate (:965-1008)
          if (ly.l >= L - 1) throw RefError{ST_BOTTOM_REACHED};  // self.next_layer is None (Q9)
          const Soil<T>& ns = layers[ly.l + 1].s;
          T theta_new = theta_from_h(cur->psi, ns);
          T mbal = overshot * (cur->theta - next->theta);
          if (!e.n2n) throw RefError{ST_NULL_NEIGHBOUR};
          T mbal_z = mbal / (theta_new - e.n2n->theta);
          T depth_new = ly.cum + mbal_z;
          cur->depth = T(ly.cum);
          next->theta = theta_new;
          next->psi = cur->psi;
          next->depth = depth_new;
          next->layer_num = ly.l + 1;
          next->dzdt = cur->dzdt;
          cur->dzdt = T(0.0);
          cur->to_bottom = true;
          next->to_bottom = false;
        }
      }
    }
    // update_wetting_fronts / check_wetting_front (:939-963)
    for (int l = 0; l < L; l++) check_wetting_front(layers[l]);
  }
  void check_wetting_front(Layer<T>& ly) {
    for (size_t i = 0; i < ly.wf.size(); i++) {
      if (ly.wf[i]->layer_num > ly.l) {
        Front<T>* popped = ly.wf[i];
        ly.wf.erase(ly.wf.begin() + i);
        if (ly.l >= L - 1) throw RefError{ST_NULL_NEIGHBOUR};
        Layer<T>& nl = layers[ly.l + 1];
        nl.wf.insert(nl.wf.begin(), popped);
        if (i >= ly.prev.size()) throw RefError{ST_INDEX_ERROR};
        Front<T> pp = ly.prev[i];
        ly.prev.erase(ly.prev.begin() + i);
        nl.prev.insert(nl.prev.begin(), pp);
        check_wetting_front(ly);
        break;
      }
    }
  }

  // ---- lower boundary: Layer.wetting_front_cross_domain_boundary (:1010-1053)
  T wetting_front_cross_domain_boundary_from(int l) {
    Layer<T>& ly = layers[l];
    T flux(0.0);
    int lf = len_layers(ly);
    for (int i = 0; i < lf; i++) {
      Nb<T> e = ext_neighbors(ly, i);  // IndexError if a previous pop shortened the list
      T tmp(0.0);
      if (e.n2n == nullptr) {
        if (val(e.cur->depth) > ly.cum) {
          if (!e.next) throw RefError{ST_NULL_NEIGHBOUR};
          g_cnt[11]++;
          tmp = (e.cur->theta - e.next->theta) * (e.cur->depth - e.next->depth);
          e.next->theta = e.cur->theta;
          T se_k = se_from_theta(e.cur->theta, ly.s);
          e.next->psi = h_from_se(se_k, ly.s);
          e.next->k = k_from_se(se_k, ly.s.ksat, ly.s.m);
          ly.wf.erase(ly.wf.begin() + i);
        }
      }
      flux = flux + tmp;
    }
    if (l < L - 1) return flux + wetting_front_cross_domain_boundary_from(l + 1);
    return flux;
  }

  // ---- dry-over-wet: Layer.py:1055-1143
  T fix_dry_over_wet_from(int l) {
    Layer<T>& ly = layers[l];
    T mass_change(0.0);
    for (int i = 0; i < (int)ly.wf.size(); i++) {
      Nb<T> nb = neighbors(ly, i);
      if (nb.next != nullptr) {
        bool theta_less = val(nb.cur->theta) <= val(nb.next->theta);
        bool same_layer = nb.cur->layer_num == nb.next->layer_num;
        if (theta_less && same_layer) {
          g_cnt[8]++;
          T mass_before = mass_balance();
          Front<T>* popped = ly.wf[i];
          ly.wf.erase(ly.wf.begin() + i);
          if (popped->layer_num > 0) cleanup_wetting_fronts(nb.next);
          T mass_after = mass_balance();
          mass_change = mass_change + abs_(mass_after - mass_before);
          break;
        }
      }
    }
    if (l < L - 1) return mass_change + fix_dry_over_wet_from(l + 1);
    return mass_change;
  }
  void cleanup_wetting_fronts(Front<T>* next_front) {  // :1098-1115 (search from the top layer)
    for (auto& ly : layers) {
      for (auto* cf : ly.wf) {
        if (cf->is_equal(*next_front)) {
          T se_k = se_from_theta(cf->theta, ly.s);
          cf->psi = h_from_se(se_k, ly.s);
          update_layer_fronts(cf);
          return;
        }
      }
    }
    throw RefError{ST_INDEX_ERROR};
  }
  void update_layer_fronts(Front<T>* dry) {  // :1117-1143 (Q15)
    for (auto& ly : layers) {
      if (ly.l < dry->layer_num) {
        for (auto* cf : ly.wf) {
          T se_l = se_from_theta(dry->theta, ly.s);
          cf->psi = h_from_se(se_l, ly.s);
          cf->theta = theta_from_h(dry->psi, ly.s);
        }
        if (ly.l >= L - 1) throw RefError{ST_NULL_NEIGHBOUR};
      } else {
        return;
      }
    }
  }

  // ---- Layer.update_psi (:1157-1174)
  void update_psi() {
    for (auto& ly : layers) {
      int lf = len_layers(ly);
      for (int i = 0; i < lf; i++) {
        Front<T>* cf = ly.wf[i];
        T se = se_from_theta(cf->theta, ly.s);
        cf->psi = h_from_se(se, ly.s);
        cf->k = k_from_se(se, ly.s.ksat, ly.s.m);
      }
    }
  }

  // ---- Layer.calc_bottom_sum (:1557-1582), started at the top layer
  T calc_bottom_sum(int l0, T bottom_sum, const Front<T>* front) {
    for (int k = l0;; k++) {
      Layer<T>& lk = layers[k];
      T theta_prev = theta_from_h(front->psi, lk.s);
      T se_prev = se_from_theta(theta_prev, lk.s);
      T kk = k_from_se(se_prev, lk.s.ksat, lk.s.m);
      double plt = (k != 0) ? layers[k - 1].cum : 0.0;
      bottom_sum = bottom_sum + ((lk.cum - plt) / kk);
      if (k + 1 >= L) throw RefError{ST_NULL_NEIGHBOUR};
      if (layers[k + 1].l == front->layer_num) return bottom_sum;
    }
  }

  // ---- Layer.calc_dzdt (:1176-1252)
  void calc_dzdt(const T& h_p) {
    for (auto& ly : layers) {
      int lf = len_layers(ly);
      for (int i = 0; i < lf; i++) {
        Nb<T> nb = neighbors(ly, i);
        Front<T>* cur = nb.cur;
        Front<T>* next = nb.next;
        if (!next) throw RefError{ST_NULL_NEIGHBOUR};
        T bottom_sum(0.0);
        T theta_1 = next->theta, theta_2 = cur->theta;
        if (cur->to_bottom) {
          cur->dzdt = T(0.0);
          continue;
        }
        if (cur->layer_num > 0) {
          if (ly.l == 0) throw RefError{ST_NULL_NEIGHBOUR};  // self.previous_layer is None
          bottom_sum = bottom_sum + (cur->depth - layers[ly.l - 1].cum) / cur->k;
        } else {
          if (val(theta_1) > val(theta_2)) throw RefError{ST_THETA_ORDER};
        }
        T geff = calc_geff(theta_1, theta_2, ly.s, cfg.nint);
        T delta_theta = cur->theta - next->theta;
        T dzdt(0.0);
        if (cur->layer_num == 0) {
          if (val(delta_theta) > 0)
            dzdt = 1.0 / delta_theta * (ly.s.ksat * (geff + h_p) / cur->depth + cur->k);
        } else {
          T denominator = calc_bottom_sum(0, bottom_sum, cur);
          T numerator = cur->depth;
          if (val(delta_theta) > 0)
            dzdt = (1.0 / delta_theta) * ((numerator / denominator) + ly.s.ksat * (geff + h_p) / cur->depth);
        }
        cur->dzdt = dzdt;
      }
    }
  }

  // ---- Layer.calc_dry_depth (:1309-1334)
  T calc_dry_depth(double dt) {
    Layer<T>& ly = layers[0];
    Front<T>* cur = ly.wf[0];
    T theta_1 = cur->theta, theta_2 = ly.s.theta_e;
    T delta_theta = ly.s.theta_e - cur->theta;
    T tau = dt * ly.s.ksat / delta_theta;
    T geff = calc_geff(theta_1, theta_2, ly.s, cfg.nint);
    T dry = 0.5 * (tau + sqrt_(tau * tau + 4.0 * tau * geff));
    return min_(T(ly.cum), dry);
  }

  // ---- Layer.create_surficial_front (:1336-1416)
  void create_surficial_front(const T& dry_depth, T& ponded_depth, T& infiltration) {
    Layer<T>& ly = layers[0];
    Front<T>* cur = ly.wf[0];
    T delta_theta = ly.s.theta_e - cur->theta;
    Front<T>* nf = new_front();
    nf->layer_num = 0;
    T theta_new;
    if (val(dry_depth * delta_theta) > val(ponded_depth)) {
      infiltration = ponded_depth;
      theta_new = min_(cur->theta + ponded_depth / dry_depth, ly.s.theta_e);
      nf->theta = theta_new;
      nf->depth = dry_depth;
      nf->to_bottom = false;
      ponded_depth = T(0.0);
    } else {
      infiltration = dry_depth * delta_theta;
      ponded_depth = ponded_depth - (dry_depth * delta_theta);
      theta_new = ly.s.theta_e;
      nf->depth = dry_depth;
      nf->theta = ly.s.theta_e;
      nf->to_bottom = !(val(dry_depth) < ly.cum);
    }
    ly.wf.insert(ly.wf.begin(), nf);
    T se = se_from_theta(theta_new, ly.s);
    nf->psi = h_from_se(se, ly.s);
    nf->k = k_from_se(se, ly.s.ksat, ly.s.m) * cfg.frozen_factor;
    nf->dzdt = T(0.0);
  }

  // ---- Layer.insert_water (:1418-1536) incl. get_drainage_neighbors (:1584-1607, Q6),
  //      calc_bottom_sum_f_p (:1538-1555, Q18)
  void insert_water(double dt, double precip_sub, T& ponded_depth, T& infiltration, T& runoff_out) {
    T h_p = clamp_min_((ponded_depth - precip_sub) * dt, 0.0);
    int lfp = fd->layer_num;
    Layer<T>& fl = layers[lfp];
    Front<T>* current_front = first_front(lfp);
    Front<T>* next_fd;
    if (fl.wf.size() > 1) next_fd = fl.wf[1];
    else {
      if (lfp >= L - 1) throw RefError{ST_NULL_NEIGHBOUR};
      next_fd = first_front(lfp + 1);
    }
    int nwf = num_fronts();
    T geff(0.0);
    T fd_ksat(0.0);
    bool have_fd_ksat = false;
    if (nwf != L) {
      T theta_1 = next_fd->theta, theta_2 = fl.s.theta_e;
      fd_ksat = fl.s.ksat * cfg.frozen_factor;
      have_fd_ksat = true;
      geff = calc_geff(theta_1, theta_2, fl.s, cfg.nint);
    }
    T f_p(0.0);
    if (lfp == 0) {
      f_p = layers[0].s.ksat * (1.0 + (geff + h_p) / fd->depth);
    } else {
      if (!have_fd_ksat) throw RefError{ST_NULL_NEIGHBOUR};  // UnboundLocalError in the reference
      double plt = layers[lfp - 1].cum;
      T bottom_sum = (fd->depth - plt) / fd_ksat;
      // calc_bottom_sum_f_p on the top layer
      T k0 = layers[0].s.ksat * cfg.frozen_factor;
      bottom_sum = bottom_sum + ((layers[0].cum - 0.0) / k0);
      if (L < 2) throw RefError{ST_NULL_NEIGHBOUR};
      if (layers[1].l != fd->layer_num) {
        g_cnt[10]++;
        bottom_sum = calc_bottom_sum(1, bottom_sum, fd);
      }
      f_p = (fd->depth / bottom_sum) + ((geff + h_p) * fd_ksat / fd->depth);
    }
    (void)current_front;  // theta_e1 / layer_nums_equal guard can never fire (Q6)
    T ponded_temp = clamp_min_(ponded_depth - f_p * dt - 0.0, 0.0);
    T fp_cm = f_p * dt + 0.0 / dt;
    const double pdm = cfg.ponded_depth_max;
    if (pdm > 0.0) {
      if (val(ponded_temp) < pdm) {
        infiltration = min_(ponded_depth, fp_cm);
        ponded_depth = ponded_depth - infiltration;
      } else if (val(ponded_temp) > pdm) {
        ponded_depth = T(pdm);
        infiltration = fp_cm;
      } else {
        g_cnt[9]++;
      }
      runoff_out = clamp_min_(ponded_temp - pdm, 0.0);
    } else {
      infiltration = min_(ponded_depth, fp_cm);
      T r = ponded_depth - infiltration;
      ponded_depth = T(pdm);
      runoff_out = clamp_min_(r, 0.0);
    }
  }

  // ---- dpLGAR.move_wetting_front (models/dpLGAR.py:340-367)
  T move_wetting_front(const T& infiltration, T& AET_sub, const T& old_mass, double dt) {
    int nwf = num_fronts();
    move_wetting_fronts(infiltration, AET_sub, old_mass, nwf, dt);
    merge_wetting_fronts();
    wetting_fronts_cross_layer_boundary();
    merge_wetting_fronts();
    T bottom_flux = T(0.0) + wetting_front_cross_domain_boundary_from(0);
    T mass_change = fix_dry_over_wet_from(0);
    if (std::fabs(val(mass_change)) > 1e-7) AET_sub = AET_sub - mass_change;
    update_psi();
    return bottom_flux;
  }

  // ---- dpLGAR.forward (models/dpLGAR.py:154-299): one forcing step = num_subcycles sub-steps
  void forward(double precip_rate, double pet_rate) {
    const double dt = cfg.dt_h;
    const double pdm = cfg.ponded_depth_max;
    T ending_volume_sub = ending_volume;
    for (int sc = 0; sc < cfg.num_subcycles; sc++) {
      copy_states();
      double precip_sub = precip_rate * dt;
      double pet_sub = pet_rate * dt;
      T previous_precip_sub = previous_precip;
      T ponded_depth_sub = precip_sub + ponded_water;
      T ponded_water_sub(0.0), percolation_sub(0.0), runoff_sub(0.0), infiltration_sub(0.0), AET_sub(0.0);
      bool create = (val(previous_precip_sub) == 0.0) && (precip_sub > 0.0) && (val(ponded_water) == 0.0);
      fd = calc_wetting_front_free_drainage();
      bool saturated = val(first_front(0)->theta) >= val(layers[0].s.theta_e);
      if (pet_rate > 0.0) AET_sub = calc_aet(pet_rate, dt);
      precip = precip + precip_sub;
      PET = PET + std::fmax(pet_sub, 0.0);
      T starting_volume_sub = mass_balance();
      (void)starting_volume_sub;
      if (create && !saturated) {
        (void)move_wetting_front(T(0.0), AET_sub, ending_volume_sub, dt);  // bottom flux dropped (Q7)
        T dry_depth = calc_dry_depth(dt);
        create_surficial_front(dry_depth, ponded_depth_sub, infiltration_sub);
        copy_states();
        infiltration = infiltration + infiltration_sub;
      }
      if (!create && val(ponded_depth_sub) > 0) {
        insert_water(dt, precip_sub, ponded_depth_sub, infiltration_sub, runoff_sub);
        infiltration = infiltration + infiltration_sub;
        runoff = runoff + runoff_sub;
        percolation_sub = infiltration_sub;
        ponded_water_sub = ponded_depth_sub;
      } else {
        // update_ponded_depth (models/dpLGAR.py:369-382)
        if (val(ponded_depth_sub) < pdm) {
          runoff_sub = T(0.0);
          runoff = runoff + runoff_sub;
          ponded_water_sub = ponded_depth_sub;
          ponded_depth_sub = T(0.0);
        } else {
          runoff_sub = ponded_depth_sub - pdm;
          ponded_depth_sub = T(pdm);
          ponded_water_sub = ponded_depth_sub;
          runoff = runoff + runoff_sub;
        }
      }
      if (!create) {
        T infiltration_temp = infiltration_sub;
        infiltration_sub = move_wetting_front(infiltration_sub, AET_sub, ending_volume_sub, dt);
        percolation_sub = infiltration_sub;
        percolation = percolation + percolation_sub;
        infiltration_sub = infiltration_temp;
      }
      calc_dzdt(ponded_depth_sub);
      ending_volume_sub = mass_balance();
      previous_precip = T(precip_sub);
      ending_volume = ending_volume_sub;
      AET = AET + AET_sub;
      ponded_water = ponded_water_sub;
      // GIUH (lgar/giuh.py:8-20)
      T qsum(0.0);
      for (int i = 0; i < cfg.num_giuh; i++) qsum = qsum + giuh_queue[i];
      if (val(qsum) > 0 || val(runoff_sub) > 0) {
        for (int i = 0; i < cfg.num_giuh; i++) giuh_queue[i] = giuh_queue[i] + (cfg.giuh[i] * runoff_sub);
        T now = giuh_queue[0];
        for (int i = 0; i + 1 < cfg.num_giuh; i++) giuh_queue[i] = giuh_queue[i + 1];
        giuh_queue[cfg.num_giuh - 1] = T(0.0);
        giuh_runoff = giuh_runoff + now;
        discharge = discharge + now;
      }
    }
  }

  void reset_accumulators() {  // MassBalance.change_mass (physics/MassBalance.py:45-53)
    precip = PET = AET = infiltration = runoff = percolation = giuh_runoff = discharge = T(0.0);
  }
};

// ------------------------------------------------------------------------------------
// C ABI for ctypes
// ------------------------------------------------------------------------------------
template <class T>
static void store_out(const Column<T>& c, double* o) {
  o[0] = val(c.runoff); o[1] = val(c.percolation); o[2] = val(c.AET); o[3] = val(c.infiltration);
  o[4] = val(c.ending_volume); o[5] = val(c.ponded_water); o[6] = val(c.giuh_runoff);
  o[7] = val(c.precip); o[8] = val(c.PET); o[9] = val(c.discharge);
}
static void store_tan(const Column<Dual>& c, double* o, int np) {
  const Dual* q[LGAR_NOUT] = {&c.runoff, &c.percolation, &c.AET, &c.infiltration, &c.ending_volume,
                              &c.ponded_water, &c.giuh_runoff, &c.precip, &c.PET, &c.discharge};
  for (int k = 0; k < LGAR_NOUT; k++)
    for (int p = 0; p < np; p++) o[k * np + p] = q[k]->d[p];
}

template <class T>
static int dump_fronts(Column<T>& c, double* fronts, int8_t* flayer, int8_t* ftb, double* dfronts, int np) {
  int j = 0;
  for (auto& ly : c.layers)
    for (auto* f : ly.wf) {
      if (j < LGAR_FMAX) {
        if (fronts) {
          double* r = fronts + j * 5;
          r[0] = val(f->depth); r[1] = val(f->theta); r[2] = val(f->psi); r[3] = val(f->k); r[4] = val(f->dzdt);
        }
        if (flayer) flayer[j] = (int8_t)f->layer_num;
        if (ftb) ftb[j] = (int8_t)f->to_bottom;
        (void)dfronts; (void)np;
      }
      j++;
    }
  return j;
}

extern "C" {

// out[T][NOUT]; fronts[T][FMAX][5] / flayer[T][FMAX] / ftb[T][FMAX] optional (NULL);
// nfronts[T]; returns status (0 = OK); *crash_step = forcing step at which the reference raises.
int lgar_oracle_forward(const lgar_oracle_cfg* cfg, const double* forcing, int T_, double* out, double* fronts,
                        int8_t* flayer, int8_t* ftb, int32_t* nfronts, int32_t* crash_step, long long* counters) {
  std::memset(g_cnt, 0, sizeof(g_cnt));
  Column<double> col;
  int status = ST_OK;
  if (crash_step) *crash_step = -1;
  try {
    col.init(*cfg, cfg->alpha, cfg->n, cfg->ksat);
  } catch (RefError& e) {
    if (crash_step) *crash_step = 0;
    return e.code;
  }
  for (int t = 0; t < T_; t++) {
    try {
      col.forward(forcing[2 * t], forcing[2 * t + 1]);
    } catch (RefError& e) {
      status = e.code;
      if (crash_step) *crash_step = t;
      break;
    }
    if (out) store_out(col, out + (size_t)t * LGAR_NOUT);
    int nf = dump_fronts(col, fronts ? fronts + (size_t)t * LGAR_FMAX * 5 : nullptr,
                         flayer ? flayer + (size_t)t * LGAR_FMAX : nullptr,
                         ftb ? ftb + (size_t)t * LGAR_FMAX : nullptr, nullptr, 0);
    if (nfronts) nfronts[t] = nf;
    col.reset_accumulators();
  }
  if (counters) std::memcpy(counters, g_cnt, sizeof(g_cnt));
  return status;
}

// Forward-mode tangents w.r.t. (alpha[L], n[L], ksat[L]) -> np = 3L <= LGAR_NT.
// out[T][NOUT], dout[T][NOUT][np]
int lgar_oracle_forward_tangent(const lgar_oracle_cfg* cfg, const double* forcing, int T_, double* out, double* dout,
                                int32_t* crash_step) {
  std::memset(g_cnt, 0, sizeof(g_cnt));
  int L = cfg->num_layers;
  int np = 3 * L;
  if (np > LGAR_NT) return -1;
  Dual a[LGAR_LMAX], n[LGAR_LMAX], k[LGAR_LMAX];
  for (int l = 0; l < L; l++) {
    a[l] = Dual(cfg->alpha[l]); a[l].d[l] = 1.0;
    n[l] = Dual(cfg->n[l]); n[l].d[L + l] = 1.0;
    k[l] = Dual(cfg->ksat[l]); k[l].d[2 * L + l] = 1.0;
  }
  Column<Dual> col;
  int status = ST_OK;
  if (crash_step) *crash_step = -1;
  try {
    col.init(*cfg, a, n, k);
  } catch (RefError& e) {
    if (crash_step) *crash_step = 0;
    return e.code;
  }
  for (int t = 0; t < T_; t++) {
    try {
      col.forward(forcing[2 * t], forcing[2 * t + 1]);
    } catch (RefError& e) {
      status = e.code;
      if (crash_step) *crash_step = t;
      break;
    }
    if (out) store_out(col, out + (size_t)t * LGAR_NOUT);
    if (dout) store_tan(col, dout + (size_t)t * LGAR_NOUT * np, np);
    col.reset_accumulators();
  }
  return status;
}

// Many columns, std::thread workers over columns (the CPU baseline of bench.py and the wide parity tests).
// cfgs[B]; column b reads the forcing record forcing + site[b] * T * 2 (site == NULL: all use record 0);
// sums[B][NOUT] = per-column sums over time of the per-step outputs of the COMPLETED steps (ending_volume and
// ponded_water: last value); status[B]; crash_step[B] = forcing step at which the reference raises (-1 = none).
// dsums (optional) [B][NOUT][3L]: forward-mode tangents of `sums` w.r.t. (alpha[L], n[L], ksat[L]) -- the
// reference-autograd semantics of the Dual scalar; the forward+gradient CPU baseline.
}  // extern "C"
template <class T>
static void batch_column(const lgar_oracle_cfg& cfg, const double* forcing, int T_, double* sums, double* dsums,
                         int32_t* status, int32_t* crash_step) {
  const int L = cfg.num_layers, np = 3 * L;
  Column<T> col;
  int st = ST_OK, crash = -1;
  double acc[LGAR_NOUT] = {0};
  std::vector<double> dacc, dtmp;
  constexpr bool TAN = !std::is_same<T, double>::value;
  if (TAN) {
    dacc.assign((size_t)LGAR_NOUT * np, 0.0);
    dtmp.assign((size_t)LGAR_NOUT * np, 0.0);
  }
  int t = 0;
  try {
    if constexpr (TAN) {
      Dual a[LGAR_LMAX], n[LGAR_LMAX], k[LGAR_LMAX];
      for (int l = 0; l < L; l++) {
        a[l] = Dual(cfg.alpha[l]); a[l].d[l] = 1.0;
        n[l] = Dual(cfg.n[l]); n[l].d[L + l] = 1.0;
        k[l] = Dual(cfg.ksat[l]); k[l].d[2 * L + l] = 1.0;
      }
      col.init(cfg, a, n, k);
    } else {
      col.init(cfg, cfg.alpha, cfg.n, cfg.ksat);
    }
    for (t = 0; t < T_; t++) {
      col.forward(forcing[2 * t], forcing[2 * t + 1]);
      double o[LGAR_NOUT];
      store_out(col, o);
      for (int k = 0; k < LGAR_NOUT; k++) acc[k] = (k == 4 || k == 5) ? o[k] : acc[k] + o[k];
      if constexpr (TAN) {
        store_tan(col, dtmp.data(), np);
        for (int k = 0; k < LGAR_NOUT; k++)
          for (int q = 0; q < np; q++)
            dacc[k * np + q] = (k == 4 || k == 5) ? dtmp[k * np + q] : dacc[k * np + q] + dtmp[k * np + q];
      }
      col.reset_accumulators();
    }
  } catch (RefError& e) {
    st = e.code;
    crash = t;
  }
  if (sums) std::memcpy(sums, acc, sizeof(acc));
  if (TAN && dsums) std::memcpy(dsums, dacc.data(), dacc.size() * sizeof(double));
  if (status) *status = st;
  if (crash_step) *crash_step = crash;
}

extern "C" {
int lgar_oracle_forward_batch_ex(const lgar_oracle_cfg* cfgs, int B, const double* forcing, const int32_t* site, int T_,
                                 double* sums, double* dsums, int32_t* status, int32_t* crash_step, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) return;
      const double* f = forcing + (site ? (size_t)site[b] * T_ * 2 : 0);
      const int np = 3 * cfgs[b].num_layers;
      if (dsums)
        batch_column<Dual>(cfgs[b], f, T_, sums ? sums + (size_t)b * LGAR_NOUT : nullptr, dsums + (size_t)b * LGAR_NOUT * np,
                           status ? status + b : nullptr, crash_step ? crash_step + b : nullptr);
      else
        batch_column<double>(cfgs[b], f, T_, sums ? sums + (size_t)b * LGAR_NOUT : nullptr, nullptr,
                             status ? status + b : nullptr, crash_step ? crash_step + b : nullptr);
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return 0;
}

int lgar_oracle_forward_batch(const lgar_oracle_cfg* cfgs, int B, const double* forcing, int T_, double* sums,
                              int32_t* status, int nthreads) {
  return lgar_oracle_forward_batch_ex(cfgs, B, forcing, nullptr, T_, sums, nullptr, status, nullptr, nthreads);
}

int lgar_oracle_sizeof_cfg(void) { return (int)sizeof(lgar_oracle_cfg); }

}  // extern "C"
