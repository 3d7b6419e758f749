// kbpoly.cu -- piecewise-polynomial form of the Kaiser-Bessel window for on-the-fly evaluation.
//
// The reference evaluates PHI (include/infft.h:209-215) with sinh/sqrt/division for each of the
// d*(2m+2) window values of a node (nfft.c:4896-4908).  On the GPU that is ~400 instructions per
// value through the FP64 pipe, as much work as the tap sums themselves (profiles/r01c).  The window
//     phi(t) = (b/pi) * sinhc(b sqrt(m^2 - t^2))
// is an entire function of t (the sinh and sin branches are the same power series in
// s = m^2 - t^2), so on each unit interval it is approximated to rounding level by a low-degree
// polynomial: for tap l of a node with fractional offset frac = n x - floor(n x) in [0,1),
//     psi_l = phi(frac + m - l) = P_l(y),   y = 2 frac - 1 in [-1,1].
// P_l is the Chebyshev interpolant of degree p on [-1,1], built here in long double and converted
// to monomial coefficients; p is raised until the maximum deviation from the long-double window
// at 64 check points per tap is below 3e-15 of the window's peak (p = 13..16 in practice; capped at kKbPolyDeg = 16), so the
// device evaluates a value with p FMAs (Horner).  If that cannot be reached the
// plan keeps the closed form (kb_phi in common.cuh).
#include "common.cuh"

#include <mutex>

#include <math.h>

namespace nfftcu {

namespace {

const long double kPiL = 3.141592653589793238462643383279502884L;

long double phi_ld(long double t, long double m, long double b, int window) {
  if (window == NFFTCU_WINDOW_GAUSSIAN) return expl(-t * t / b) / sqrtl(kPiL * b);   // include/infft.h:155-156
  const long double s = m * m - t * t;
  if (s > 0) { const long double r = sqrtl(s); return sinhl(b * r) / (kPiL * r); }
  if (s < 0) { const long double r = sqrtl(-s); return sinl(b * r) / (kPiL * r); }
  return b / kPiL;
}

// monomial coefficients (in y) of the degree-p Chebyshev interpolant of f on [-1,1]
void cheb_fit_monomial(int p, long double m, long double b, int window, int l, std::vector<long double> &mono) {
  std::vector<long double> c((size_t) p + 1), fj((size_t) p + 1);
  for (int j = 0; j <= p; j++) {
    const long double yj = cosl(kPiL * (j + 0.5L) / (p + 1));
    fj[(size_t) j] = phi_ld((yj + 1) / 2 + m - l, m, b, window);
  }
  for (int k = 0; k <= p; k++) {
    long double s = 0;
    for (int j = 0; j <= p; j++) s += fj[(size_t) j] * cosl(kPiL * k * (j + 0.5L) / (p + 1));
    c[(size_t) k] = s * 2 / (p + 1);
  }
  c[0] /= 2;
  // sum_k c_k T_k(y) -> monomials via T_{k+1} = 2 y T_k - T_{k-1}
  std::vector<long double> tkm1((size_t) p + 1, 0.0L), tk((size_t) p + 1, 0.0L), tkp1((size_t) p + 1);
  mono.assign((size_t) p + 1, 0.0L);
  tkm1[0] = 1;                       // T_0
  for (int q = 0; q <= p; q++) mono[(size_t) q] += c[0] * tkm1[(size_t) q];
  if (p >= 1) {
    tk[1] = 1;                       // T_1
    for (int q = 0; q <= p; q++) mono[(size_t) q] += c[1] * tk[(size_t) q];
  }
  for (int k = 2; k <= p; k++) {
    for (int q = 0; q <= p; q++)
      tkp1[(size_t) q] = (q > 0 ? 2 * tk[(size_t) q - 1] : 0.0L) - tkm1[(size_t) q];
    for (int q = 0; q <= p; q++) mono[(size_t) q] += c[(size_t) k] * tkp1[(size_t) q];
    tkm1 = tk;
    tk = tkp1;
  }
}

}  // namespace

// Fills c->kbpoly (host, double) with layout coef[(t*(kKbPolyDeg+1) + k)*W + l], k = power of y, and
// uploads it.  Returns NFFTCU_OK; c->kbpoly_deg stays -1 when no adequate polynomial was found.
// The fit depends on (precision, m, b_t) only and costs ~20 ms of long-double arithmetic: plan-per-coil callers create
// many identical plans, so the results are cached per process.
struct FitKey {
  int prec, d, window;
  long long m;
  double b[NFFTCU_MAX_D], ws[NFFTCU_MAX_D];
  bool operator==(const FitKey &o) const {
    if (prec != o.prec || d != o.d || m != o.m || window != o.window) return false;
    for (int t = 0; t < d; t++) if (b[t] != o.b[t] || ws[t] != o.ws[t]) return false;
    return true;
  }
};
struct FitEntry { FitKey key; int fit; std::vector<double> coef; };
static std::mutex g_fit_mutex;
static std::vector<FitEntry> g_fit_cache;

static int upload_poly(nfftcu_ctx *c) {
  if (c->kbpoly_dev) pool_free(c->kbpoly_dev);
  c->kbpoly_dev = nullptr;
  NFFTCU_CUDA(pool_malloc(&c->kbpoly_dev, sizeof(double) * c->kbpoly_host.size()));
  NFFTCU_CUDA(cudaMemcpy(c->kbpoly_dev, c->kbpoly_host.data(), sizeof(double) * c->kbpoly_host.size(),
                         cudaMemcpyHostToDevice));
  return NFFTCU_OK;
}

int build_kb_poly(nfftcu_ctx *c) {
  const int W = 2 * (int) c->m + 2;
  const long double m = (long double) c->m;
  c->kbpoly_deg = -1;
  c->kbpoly_fit = -1;
  FitKey key;
  key.prec = c->prec; key.d = c->d; key.m = c->m; key.window = c->window;
  for (int t = 0; t < c->d; t++) { key.b[t] = c->b[t]; key.ws[t] = c->wscale[t]; }
  {
    std::lock_guard<std::mutex> lock(g_fit_mutex);
    for (const FitEntry &e : g_fit_cache)
      if (e.key == key) {
        if (e.fit < 0) return NFFTCU_OK;
        c->kbpoly_deg = kKbPolyDeg;
        c->kbpoly_fit = e.fit;
        c->kbpoly_host = e.coef;
        return upload_poly(c);
      }
  }
  // The device keeps the coefficients of one tap in registers and runs a fixed-length Horner loop, so
  // the table is always stored with kKbPolyDeg+1 coefficients per tap (higher ones zero).
  // fp32 plans need the window to ~1e-9 of its peak only (their results carry 1e-7): about half the degree
  const long double tol = c->prec == NFFTCU_DOUBLE ? 3e-15L : 2e-9L;
  for (int p = c->prec == NFFTCU_DOUBLE ? 10 : 4; p <= kKbPolyDeg; p++) {
    std::vector<double> coef((size_t) c->d * (kKbPolyDeg + 1) * W, 0.0);
    long double worst = 0;
    for (int t = 0; t < c->d; t++) {
      const long double b = (long double) c->b[t];
      const long double peak = phi_ld(0, m, b, c->window);
      const long double ws = (long double) c->wscale[t];   // exact power of two, folded into the coefficients
      for (int l = 0; l < W; l++) {
        std::vector<long double> mono;
        cheb_fit_monomial(p, m, b, c->window, l, mono);
        for (int k = 0; k <= p; k++) coef[((size_t) t * (kKbPolyDeg + 1) + k) * W + l] = (double) (mono[(size_t) k] * ws);
        for (int q = 0; q <= 64; q++) {   // check the double-precision Horner value itself
          const double y = -1.0 + 2.0 * q / 64.0;
          double acc = coef[((size_t) t * (kKbPolyDeg + 1) + kKbPolyDeg) * W + l];
          for (int k = kKbPolyDeg - 1; k >= 0; k--) acc = fma(acc, y, coef[((size_t) t * (kKbPolyDeg + 1) + k) * W + l]);
          const long double ref = phi_ld(((long double) y + 1) / 2 + m - l, m, b, c->window) * ws;
          const long double e = fabsl((long double) acc - ref) / (peak * ws);
          if (e > worst) worst = e;
        }
      }
    }
    if (worst < tol) {
      c->kbpoly_deg = kKbPolyDeg;   // stored (padded) degree; the fitted degree is p
      c->kbpoly_fit = p;
      c->kbpoly_host = coef;
      break;
    }
  }
  {
    std::lock_guard<std::mutex> lock(g_fit_mutex);
    if (g_fit_cache.size() < 64) g_fit_cache.push_back(FitEntry{key, c->kbpoly_fit, c->kbpoly_host});
  }
  if (c->kbpoly_deg < 0) return NFFTCU_OK;
  return upload_poly(c);
}

}  // namespace nfftcu
