// Minimal googletest-compatible macros -- NOT googletest.  googletest is fetched from the network by the
// reference's build (cmake/gtest.cmake:7-8) and is not available in this image; this header provides TEST,
// ASSERT_* / EXPECT_* and a main() so that the reference's own tests/*.cpp compile unchanged against the
// B200 host library.
#ifndef DPGO_GTEST_SHIM_H
#define DPGO_GTEST_SHIM_H

#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
struct TestInfo {
  const char *suite, *name;
  void (*fn)(bool &);
};
inline std::vector<TestInfo> &registry() {
  static std::vector<TestInfo> r;
  return r;
}
struct Registrar {
  Registrar(const char *s, const char *n, void (*f)(bool &)) { registry().push_back({s, n, f}); }
};
inline void InitGoogleTest(int *, char **) {}
}  // namespace testing

#define TEST(suite, name)                                                                          \
  static void suite##_##name##_body(bool &gtest_failed_);                                          \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name, &suite##_##name##_body);         \
  static void suite##_##name##_body(bool &gtest_failed_)

#define GTEST_SHIM_CHECK_(cond, text, fatal)                                                       \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      std::cerr << __FILE__ << ":" << __LINE__ << ": Failure\n  " << text << std::endl;            \
      gtest_failed_ = true;                                                                        \
      if (fatal) return;                                                                           \
    }                                                                                              \
  } while (0)

#define GTEST_SHIM_CMP_(a, op, b, fatal)                                                           \
  do {                                                                                             \
    auto gtest_a_ = (a);                                                                           \
    auto gtest_b_ = (b);                                                                           \
    if (!(gtest_a_ op gtest_b_)) {                                                                 \
      std::ostringstream gtest_os_;                                                                \
      gtest_os_ << "Expected: (" #a ") " #op " (" #b "), actual: " << gtest_a_ << " vs " << gtest_b_; \
      std::cerr << __FILE__ << ":" << __LINE__ << ": Failure\n  " << gtest_os_.str() << std::endl; \
      gtest_failed_ = true;                                                                        \
      if (fatal) return;                                                                           \
    }                                                                                              \
  } while (0)

#define ASSERT_TRUE(c) GTEST_SHIM_CHECK_((c), "Expected true: " #c, true)
#define ASSERT_FALSE(c) GTEST_SHIM_CHECK_(!(c), "Expected false: " #c, true)
#define EXPECT_TRUE(c) GTEST_SHIM_CHECK_((c), "Expected true: " #c, false)
#define EXPECT_FALSE(c) GTEST_SHIM_CHECK_(!(c), "Expected false: " #c, false)
#define ASSERT_EQ(a, b) GTEST_SHIM_CMP_(a, ==, b, true)
#define ASSERT_NE(a, b) GTEST_SHIM_CMP_(a, !=, b, true)
#define ASSERT_LE(a, b) GTEST_SHIM_CMP_(a, <=, b, true)
#define ASSERT_LT(a, b) GTEST_SHIM_CMP_(a, <, b, true)
#define ASSERT_GE(a, b) GTEST_SHIM_CMP_(a, >=, b, true)
#define ASSERT_GT(a, b) GTEST_SHIM_CMP_(a, >, b, true)
#define EXPECT_EQ(a, b) GTEST_SHIM_CMP_(a, ==, b, false)
#define EXPECT_LE(a, b) GTEST_SHIM_CMP_(a, <=, b, false)
#define EXPECT_LT(a, b) GTEST_SHIM_CMP_(a, <, b, false)
#define ASSERT_NEAR(a, b, tol) GTEST_SHIM_CHECK_(std::fabs((a) - (b)) <= (tol), "Expected |" #a " - " #b "| <= " #tol, true)
#define EXPECT_NEAR(a, b, tol) GTEST_SHIM_CHECK_(std::fabs((a) - (b)) <= (tol), "Expected |" #a " - " #b "| <= " #tol, false)

inline int RUN_ALL_TESTS() {
  int failed = 0;
  for (const auto &t : ::testing::registry()) {
    std::printf("[ RUN      ] %s.%s\n", t.suite, t.name);
    bool f = false;
    t.fn(f);
    std::printf(f ? "[  FAILED  ] %s.%s\n" : "[       OK ] %s.%s\n", t.suite, t.name);
    failed += f ? 1 : 0;
  }
  std::printf("[==========] %zu tests ran, %d failed.\n", ::testing::registry().size(), failed);
  return failed ? 1 : 0;
}

#ifdef GTEST_SHIM_MAIN
int main(int argc, char **argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
#endif
#endif
