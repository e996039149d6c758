// TEST INFRASTRUCTURE (oracle/): stand-in for Boost.Math's chi_squared (absent
// from this image) so /root/reference/src/entropy.cc compiles unchanged.
// cdf = regularised lower incomplete gamma P(k/2, x/2) (series / continued
// fraction, Numerical-Recipes style); quantile(complement) by bisection.
// Check: dof 1, sig 0.05 -> 3.841458820694124.
#ifndef HMC_ORACLE_SHIM_CHI_SQUARED_HPP
#define HMC_ORACLE_SHIM_CHI_SQUARED_HPP
#include <cmath>
namespace boost {
namespace math {
struct chi_squared {
  double dof;
  explicit chi_squared(double d) : dof(d) {}
};
namespace detail_shim {
inline double gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  const double gln = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; n++) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - gln);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - gln) * h;
}
} // namespace detail_shim
inline double cdf(const chi_squared &c, double x) {
  return detail_shim::gamma_p(0.5 * c.dof, 0.5 * x);
}
struct chi_complement {
  chi_squared dist;
  double q;
};
inline chi_complement complement(const chi_squared &c, double q) {
  return chi_complement{c, q};
}
inline double quantile(const chi_complement &c) {
  // find x with 1 - cdf(x) == q
  double lo = 0.0, hi = 1.0;
  while (1.0 - cdf(c.dist, hi) > c.q) hi *= 2.0;
  for (int i = 0; i < 200; i++) {
    const double mid = 0.5 * (lo + hi);
    if (1.0 - cdf(c.dist, mid) > c.q)
      lo = mid;
    else
      hi = mid;
  }
  return 0.5 * (lo + hi);
}
} // namespace math
} // namespace boost
#endif
