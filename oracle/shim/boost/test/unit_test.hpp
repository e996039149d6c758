// TEST INFRASTRUCTURE (oracle/): minimal Boost.Test stand-in (Boost is absent
// from this image) so /root/reference/src/hybridmc_test.cc compiles unchanged
// and its 22 known-answer cases can pin the shim build of the reference.
#ifndef HMC_ORACLE_SHIM_BOOST_TEST_HPP
#define HMC_ORACLE_SHIM_BOOST_TEST_HPP
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <vector>

namespace hmc_shim_test {
struct Case {
  const char *name;
  void (*fn)();
};
inline std::vector<Case> &cases() {
  static std::vector<Case> c;
  return c;
}
inline unsigned long &checks() {
  static unsigned long n = 0;
  return n;
}
inline unsigned long &failures() {
  static unsigned long n = 0;
  return n;
}
struct Reg {
  Reg(const char *n, void (*f)()) { cases().push_back({n, f}); }
};
inline void report(bool ok, const char *file, int line, const std::string &what) {
  checks()++;
  if (!ok) {
    failures()++;
    std::cerr << file << ":" << line << ": check failed: " << what << std::endl;
  }
}
template <class A, class B> bool close_frac(A a, B b, double tol) {
  const double x = double(a), y = double(b);
  const double d = std::fabs(x - y);
  // Boost "close fraction" (strong): d/|x| <= tol and d/|y| <= tol
  if (d == 0) return true;
  return d <= tol * std::fabs(x) && d <= tol * std::fabs(y);
}
} // namespace hmc_shim_test

#define BOOST_AUTO_TEST_CASE(name)                                             \
  static void name##_impl();                                                   \
  static hmc_shim_test::Reg name##_reg(#name, &name##_impl);                   \
  static void name##_impl()

#define BOOST_CHECK(expr)                                                      \
  hmc_shim_test::report(bool(expr), __FILE__, __LINE__, #expr)
#define BOOST_CHECK_MESSAGE(expr, msg)                                         \
  do {                                                                         \
    std::ostringstream hmc_os_;                                                \
    hmc_os_ << #expr << " : " << msg;                                          \
    hmc_shim_test::report(bool(expr), __FILE__, __LINE__, hmc_os_.str());      \
  } while (0)
#define BOOST_CHECK_EQUAL(a, b)                                                \
  hmc_shim_test::report((a) == (b), __FILE__, __LINE__, #a " == " #b)
#define BOOST_REQUIRE_EQUAL(a, b)                                              \
  do {                                                                         \
    const bool hmc_ok_ = ((a) == (b));                                         \
    hmc_shim_test::report(hmc_ok_, __FILE__, __LINE__, #a " == " #b);          \
    if (!hmc_ok_) return;                                                      \
  } while (0)
#define BOOST_CHECK_SMALL(a, tol)                                              \
  hmc_shim_test::report(std::fabs(double(a)) <= (tol), __FILE__, __LINE__,     \
                        "|" #a "| <= " #tol)
#define BOOST_CHECK_CLOSE_FRACTION(a, b, tol)                                  \
  hmc_shim_test::report(hmc_shim_test::close_frac((a), (b), (tol)), __FILE__,  \
                        __LINE__, #a " ~ " #b)
#define BOOST_CHECK_EQUAL_COLLECTIONS(b1, e1, b2, e2)                          \
  do {                                                                         \
    auto hmc_i1_ = (b1);                                                       \
    auto hmc_i2_ = (b2);                                                       \
    bool hmc_ok_ = true;                                                       \
    for (; hmc_i1_ != (e1) && hmc_i2_ != (e2); ++hmc_i1_, ++hmc_i2_)           \
      if (!(*hmc_i1_ == *hmc_i2_)) hmc_ok_ = false;                            \
    if (hmc_i1_ != (e1) || hmc_i2_ != (e2)) hmc_ok_ = false;                   \
    hmc_shim_test::report(hmc_ok_, __FILE__, __LINE__, "collections equal");   \
  } while (0)

#ifdef BOOST_TEST_MODULE
int main() {
  for (auto &c : hmc_shim_test::cases()) {
    const unsigned long f0 = hmc_shim_test::failures();
    c.fn();
    std::printf("case %-48s %s\n", c.name,
                hmc_shim_test::failures() == f0 ? "ok" : "FAILED");
  }
  std::printf("cases %zu checks %lu failures %lu\n",
              hmc_shim_test::cases().size(), hmc_shim_test::checks(),
              hmc_shim_test::failures());
  return hmc_shim_test::failures() ? 1 : 0;
}
#endif
#endif
