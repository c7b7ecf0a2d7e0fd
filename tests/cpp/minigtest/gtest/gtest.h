// Minimal single-header stand-in for the subset of GoogleTest that the reference's own test files use (TEST, TEST_F,
// EXPECT_*/ASSERT_* with streamed messages, GTEST_SKIP, global environments). GoogleTest itself is fetched from the
// network by the reference's CMake (tests/CMakeLists.txt FetchContent) and is not available offline; this header lets
// the UNMODIFIED reference tests (compiled in place from /root/reference/tests) run against libaxiom_cuda.so.
// Test infrastructure only. Output format follows gtest ("[ RUN ]", "[ OK ]", "[ FAILED ]", "[ SKIPPED ]").
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

class Message {
  public:
    template <typename T> Message &operator<<(const T &v) { ss_ << v; return *this; }
    Message &operator<<(std::ostream &(*f)(std::ostream &)) { ss_ << f; return *this; }
    std::string str() const { return ss_.str(); }
  private:
    std::ostringstream ss_;
};

// declared for axiom_test_utils.hpp; the typed / value-parameterised test files themselves are not built
template <typename T> struct TestParamInfo {
    T param;
    size_t index;
};
template <typename... Ts> struct Types {};

class Environment {
  public:
    virtual ~Environment() = default;
    virtual void SetUp() {}
    virtual void TearDown() {}
};

class Test {
  public:
    virtual ~Test() = default;
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};

namespace internal {

struct State {
    bool failed = false, skipped = false;
    int n_failed_checks = 0;
};
inline State &state() { static State s; return s; }

struct TestInfo {
    std::string suite, name;
    std::function<Test *()> make;
};
inline std::vector<TestInfo> &registry() { static std::vector<TestInfo> r; return r; }
inline std::vector<Environment *> &environments() { static std::vector<Environment *> e; return e; }
inline int register_test(const char *suite, const char *name, std::function<Test *()> make) {
    registry().push_back({suite, name, std::move(make)});
    return 0;
}

enum Kind { kNonFatal, kFatal, kSkip };
// Receives the streamed message; reports when destroyed at the end of the full expression
class Reporter {
  public:
    Reporter(Kind k, const char *file, int line, std::string what) : kind_(k), file_(file), line_(line), what_(std::move(what)) {}
    Reporter(const Reporter &) = delete;
    ~Reporter() {
        if (kind_ == kSkip) {
            state().skipped = true;
            std::printf("%s:%d: Skipped\n%s\n", file_, line_, msg_.str().c_str());
        } else {
            state().failed = true;
            state().n_failed_checks++;
            std::printf("%s:%d: Failure\n%s\n%s\n", file_, line_, what_.c_str(), msg_.str().c_str());
        }
        std::fflush(stdout);
    }
    template <typename T> Reporter &operator<<(const T &v) { msg_ << v; return *this; }
    Reporter &operator<<(std::ostream &(*f)(std::ostream &)) { msg_ << f; return *this; }
  private:
    Kind kind_;
    const char *file_;
    int line_;
    std::string what_;
    Message msg_;
};
struct Voidify {
    void operator&(const Reporter &) const {}
};

template <typename T> std::string show(const T &v) {
    if constexpr (requires(std::ostream &o, const T &x) { o << x; }) {
        std::ostringstream ss;
        ss << v;
        return ss.str();
    } else if constexpr (std::is_enum_v<T>) {
        return std::to_string(static_cast<long long>(v));
    } else {
        return "<value>";
    }
}
inline std::string show(bool v) { return v ? "true" : "false"; }
inline std::string show(std::nullptr_t) { return "nullptr"; }

template <typename A, typename B, typename Op>
std::string cmp(const A &a, const B &b, Op op, const char *ea, const char *eb, const char *opname) {
    if (op(a, b))
        return "";
    return std::string("Expected: (") + ea + ") " + opname + " (" + eb + "), actual: " + show(a) + " vs " + show(b);
}
inline std::string near(double a, double b, double tol, const char *ea, const char *eb) {
    if (std::fabs(a - b) <= tol)
        return "";
    std::ostringstream ss;
    ss.precision(17);
    ss << "The difference between " << ea << " and " << eb << " is " << std::fabs(a - b) << ", which exceeds " << tol << " (" << a
       << " vs " << b << ")";
    return ss.str();
}
template <typename F> std::string float_eq(F a, F b, const char *ea, const char *eb) {
    // 4-ULP comparison like gtest's FloatingPoint::AlmostEquals
    if (a == b)
        return "";
    if (std::isnan(a) || std::isnan(b))
        return std::string(ea) + " vs " + eb + ": NaN";
    using I = std::conditional_t<sizeof(F) == 4, int32_t, int64_t>;
    using U = std::conditional_t<sizeof(F) == 4, uint32_t, uint64_t>;
    auto biased = [](F f) {
        U u;
        std::memcpy(&u, &f, sizeof f);
        const U sign = U(1) << (sizeof(F) * 8 - 1);
        return (u & sign) ? ~u + 1 : sign | u;
    };
    const U x = biased(a), y = biased(b);
    const U d = x > y ? x - y : y - x;
    (void)sizeof(I);
    if (d <= 4)
        return "";
    std::ostringstream ss;
    ss.precision(17);
    ss << "Expected equality of " << ea << " (" << a << ") and " << eb << " (" << b << ")";
    return ss.str();
}

} // namespace internal

inline Environment *AddGlobalTestEnvironment(Environment *e) {
    internal::environments().push_back(e);
    return e;
}
inline void InitGoogleTest(int *, char **) {}

} // namespace testing

#define GTEST_CHECK_(kind, failmsg_expr)                                                                          \
    if (const std::string gtest_msg_ = (failmsg_expr); gtest_msg_.empty())                                         \
        ;                                                                                                          \
    else                                                                                                           \
        GTEST_ON_FAIL_##kind ::testing::internal::Voidify() &                                                      \
            ::testing::internal::Reporter(::testing::internal::kind, __FILE__, __LINE__, gtest_msg_)
#define GTEST_ON_FAIL_kNonFatal
#define GTEST_ON_FAIL_kFatal return
#define GTEST_ON_FAIL_kSkip return

#define GTEST_BOOL_(c, want) (static_cast<bool>(c) == want ? std::string() : std::string("Value of: " #c "\n  Expected: " #want))
#define EXPECT_TRUE(c) GTEST_CHECK_(kNonFatal, GTEST_BOOL_(c, true))
#define EXPECT_FALSE(c) GTEST_CHECK_(kNonFatal, GTEST_BOOL_(c, false))
#define ASSERT_TRUE(c) GTEST_CHECK_(kFatal, GTEST_BOOL_(c, true))
#define ASSERT_FALSE(c) GTEST_CHECK_(kFatal, GTEST_BOOL_(c, false))

#define GTEST_CMP_(a, b, op, name) ::testing::internal::cmp((a), (b), [](const auto &x, const auto &y) { return x op y; }, #a, #b, name)
#define EXPECT_EQ(a, b) GTEST_CHECK_(kNonFatal, GTEST_CMP_(a, b, ==, "=="))
#define EXPECT_NE(a, b) GTEST_CHECK_(kNonFatal, GTEST_CMP_(a, b, !=, "!="))
#define EXPECT_LT(a, b) GTEST_CHECK_(kNonFatal, GTEST_CMP_(a, b, <, "<"))
#define EXPECT_LE(a, b) GTEST_CHECK_(kNonFatal, GTEST_CMP_(a, b, <=, "<="))
#define EXPECT_GT(a, b) GTEST_CHECK_(kNonFatal, GTEST_CMP_(a, b, >, ">"))
#define EXPECT_GE(a, b) GTEST_CHECK_(kNonFatal, GTEST_CMP_(a, b, >=, ">="))
#define ASSERT_EQ(a, b) GTEST_CHECK_(kFatal, GTEST_CMP_(a, b, ==, "=="))
#define ASSERT_NE(a, b) GTEST_CHECK_(kFatal, GTEST_CMP_(a, b, !=, "!="))
#define ASSERT_LT(a, b) GTEST_CHECK_(kFatal, GTEST_CMP_(a, b, <, "<"))
#define ASSERT_LE(a, b) GTEST_CHECK_(kFatal, GTEST_CMP_(a, b, <=, "<="))
#define ASSERT_GT(a, b) GTEST_CHECK_(kFatal, GTEST_CMP_(a, b, >, ">"))
#define ASSERT_GE(a, b) GTEST_CHECK_(kFatal, GTEST_CMP_(a, b, >=, ">="))
#define EXPECT_NEAR(a, b, t) GTEST_CHECK_(kNonFatal, ::testing::internal::near((double)(a), (double)(b), (double)(t), #a, #b))
#define ASSERT_NEAR(a, b, t) GTEST_CHECK_(kFatal, ::testing::internal::near((double)(a), (double)(b), (double)(t), #a, #b))
#define EXPECT_FLOAT_EQ(a, b) GTEST_CHECK_(kNonFatal, ::testing::internal::float_eq<float>((float)(a), (float)(b), #a, #b))
#define ASSERT_FLOAT_EQ(a, b) GTEST_CHECK_(kFatal, ::testing::internal::float_eq<float>((float)(a), (float)(b), #a, #b))
#define EXPECT_DOUBLE_EQ(a, b) GTEST_CHECK_(kNonFatal, ::testing::internal::float_eq<double>((double)(a), (double)(b), #a, #b))
#define ASSERT_DOUBLE_EQ(a, b) GTEST_CHECK_(kFatal, ::testing::internal::float_eq<double>((double)(a), (double)(b), #a, #b))
#define EXPECT_STREQ(a, b) GTEST_CHECK_(kNonFatal, (std::strcmp((a), (b)) == 0 ? std::string() : std::string("Expected equal strings: " #a " vs " #b)))

#define GTEST_THROW_(stmt, extype, kind)                                                                           \
    GTEST_CHECK_(kind, ([&]() -> std::string {                                                                     \
                     try {                                                                                         \
                         stmt;                                                                                     \
                     } catch (const extype &) {                                                                    \
                         return "";                                                                                \
                     } catch (...) {                                                                               \
                         return "Expected: " #stmt " throws " #extype ".\n  Actual: it throws a different type.";    \
                     }                                                                                             \
                     return "Expected: " #stmt " throws " #extype ".\n  Actual: it throws nothing.";                \
                 })())
#define EXPECT_THROW(stmt, extype) GTEST_THROW_(stmt, extype, kNonFatal)
#define ASSERT_THROW(stmt, extype) GTEST_THROW_(stmt, extype, kFatal)
#define GTEST_NO_THROW_(stmt, kind)                                                                                \
    GTEST_CHECK_(kind, ([&]() -> std::string {                                                                     \
                     try {                                                                                         \
                         stmt;                                                                                     \
                     } catch (const std::exception &e) {                                                           \
                         return std::string("Expected: " #stmt " doesn't throw.\n  Actual: it throws ") + e.what(); \
                     } catch (...) {                                                                               \
                         return "Expected: " #stmt " doesn't throw.\n  Actual: it throws.";                         \
                     }                                                                                             \
                     return "";                                                                                    \
                 })())
#define EXPECT_NO_THROW(stmt) GTEST_NO_THROW_(stmt, kNonFatal)
#define ASSERT_NO_THROW(stmt) GTEST_NO_THROW_(stmt, kFatal)
#define GTEST_ANY_THROW_(stmt, kind)                                                                               \
    GTEST_CHECK_(kind, ([&]() -> std::string {                                                                     \
                     try {                                                                                         \
                         stmt;                                                                                     \
                     } catch (...) {                                                                               \
                         return "";                                                                                \
                     }                                                                                             \
                     return "Expected: " #stmt " throws.\n  Actual: it doesn't.";                                   \
                 })())
#define EXPECT_ANY_THROW(stmt) GTEST_ANY_THROW_(stmt, kNonFatal)
#define ASSERT_ANY_THROW(stmt) GTEST_ANY_THROW_(stmt, kFatal)

#define GTEST_SKIP() GTEST_CHECK_(kSkip, std::string("skipped"))
#define FAIL() GTEST_CHECK_(kFatal, std::string("Failed"))
#define ADD_FAILURE() GTEST_CHECK_(kNonFatal, std::string("Failed"))
#define SUCCEED() ::testing::Message()
#define SCOPED_TRACE(msg) ((void)0)

#define GTEST_TEST_CLASS_(suite, name) suite##_##name##_Test
#define GTEST_DEFINE_(suite, name, base)                                                                           \
    class GTEST_TEST_CLASS_(suite, name) : public base {                                                           \
      public:                                                                                                      \
        void TestBody() override;                                                                                  \
    };                                                                                                             \
    static int gtest_reg_##suite##_##name =                                                                        \
        ::testing::internal::register_test(#suite, #name, [] { return (::testing::Test *)new GTEST_TEST_CLASS_(suite, name); }); \
    void GTEST_TEST_CLASS_(suite, name)::TestBody()
#define TEST(suite, name) GTEST_DEFINE_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_DEFINE_(fixture, name, fixture)

#ifdef MINIGTEST_MAIN
// usage: runner [--gtest_filter=SUBSTR[:SUBSTR...]] ; '-' prefix inside the filter excludes (substring match on Suite.Name)
int main(int argc, char **argv) {
    std::vector<std::string> inc, exc;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        const std::string key = "--gtest_filter=";
        if (a.rfind(key, 0) == 0) {
            std::string f = a.substr(key.size()), neg;
            const size_t dash = f.find('-');
            if (dash != std::string::npos) {
                neg = f.substr(dash + 1);
                f = f.substr(0, dash);
            }
            auto split = [](const std::string &s, std::vector<std::string> &out) {
                size_t p = 0;
                while (p <= s.size()) {
                    size_t q = s.find(':', p);
                    if (q == std::string::npos)
                        q = s.size();
                    std::string t = s.substr(p, q - p);
                    while (!t.empty() && t.back() == '*') t.pop_back();
                    while (!t.empty() && t.front() == '*') t.erase(t.begin());
                    if (!t.empty())
                        out.push_back(t);
                    p = q + 1;
                }
            };
            split(f, inc);
            split(neg, exc);
        }
    }
    for (auto *e : ::testing::internal::environments())
        e->SetUp();
    int n_run = 0, n_ok = 0, n_skip = 0;
    std::vector<std::string> failed;
    for (auto &t : ::testing::internal::registry()) {
        const std::string full = t.suite + "." + t.name;
        bool sel = inc.empty();
        for (auto &s : inc)
            sel = sel || full.find(s) != std::string::npos;
        for (auto &s : exc)
            sel = sel && full.find(s) == std::string::npos;
        if (!sel)
            continue;
        std::printf("[ RUN      ] %s\n", full.c_str());
        std::fflush(stdout);
        auto &st = ::testing::internal::state();
        st.failed = st.skipped = false;
        ++n_run;
        try {
            std::unique_ptr<::testing::Test> obj(t.make());
            obj->SetUp();
            if (!st.skipped && !st.failed)
                obj->TestBody();
            obj->TearDown();
        } catch (const std::exception &e) {
            st.failed = true;
            std::printf("unknown file: Failure\nC++ exception with description \"%s\" thrown in the test body.\n", e.what());
        } catch (...) {
            st.failed = true;
            std::printf("unknown file: Failure\nUnknown C++ exception thrown in the test body.\n");
        }
        if (st.failed) {
            failed.push_back(full);
            std::printf("[  FAILED  ] %s\n", full.c_str());
        } else if (st.skipped) {
            ++n_skip;
            std::printf("[  SKIPPED ] %s\n", full.c_str());
        } else {
            ++n_ok;
            std::printf("[       OK ] %s\n", full.c_str());
        }
        std::fflush(stdout);
    }
    for (auto *e : ::testing::internal::environments())
        e->TearDown();
    std::printf("[==========] %d tests ran.\n[  PASSED  ] %d tests.\n", n_run, n_ok);
    if (n_skip)
        std::printf("[  SKIPPED ] %d tests.\n", n_skip);
    if (!failed.empty()) {
        std::printf("[  FAILED  ] %zu tests, listed below:\n", failed.size());
        for (auto &f : failed)
            std::printf("[  FAILED  ] %s\n", f.c_str());
    }
    std::printf("MINIGTEST_SUMMARY {\"ran\": %d, \"passed\": %d, \"skipped\": %d, \"failed\": %zu}\n", n_run, n_ok, n_skip, failed.size());
    return failed.empty() ? 0 : 1;
}
#endif
