// Minimal stand-in for the slice of googletest 1.8.0 that PA4/handout/test/{main.cpp,test_spmm.cu} use
// (testing::Test, TEST_F, ASSERT_LT, InitGoogleTest, RUN_ALL_TESTS). googletest itself is not installed and the
// handout fetches it over the network (PA4/handout/cmake/googletest-download.cmake:12-15). TEST INFRASTRUCTURE.
#ifndef B200_GTEST_SHIM_H_
#define B200_GTEST_SHIM_H_
#include <chrono>
#include <cstdio>
#include <functional>
#include <string>
#include <utility>
#include <vector>

namespace testing {
class Test {
   public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};
struct Registry {
    std::vector<std::pair<std::string, std::function<Test *()>>> tests;
    bool failed_now = false;
    static Registry &get() {
        static Registry r;
        return r;
    }
    static bool add(const char *name, std::function<Test *()> make) {
        get().tests.emplace_back(name, std::move(make));
        return true;
    }
};
inline void InitGoogleTest(int *, char **) {}
inline int run_all() {
    Registry &r = Registry::get();
    int failed = 0;
    std::printf("[==========] Running %zu tests from 1 test case.\n", r.tests.size());
    for (auto &t : r.tests) {
        std::printf("[ RUN      ] %s\n", t.first.c_str());
        std::fflush(stdout);
        r.failed_now = false;
        const auto t0 = std::chrono::steady_clock::now();
        Test *obj = t.second();
        obj->SetUp();
        obj->TestBody();
        obj->TearDown();
        delete obj;
        const long ms = (long)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
        std::printf(r.failed_now ? "[  FAILED  ] %s (%ld ms)\n" : "[       OK ] %s (%ld ms)\n", t.first.c_str(), ms);
        failed += r.failed_now;
    }
    if (failed) std::printf("[  FAILED  ] %d tests.\n", failed);
    else std::printf("[  PASSED  ] %zu tests.\n", r.tests.size());
    return failed ? 1 : 0;
}
}  // namespace testing

#define TEST_F(fixture, name)                                                                      \
    class fixture##_##name##_Test : public fixture {                                               \
       public:                                                                                     \
        void TestBody() override;                                                                  \
    };                                                                                             \
    static bool fixture##_##name##_registered =                                                    \
        ::testing::Registry::add(#fixture "." #name, []() -> ::testing::Test * { return new fixture##_##name##_Test; }); \
    void fixture##_##name##_Test::TestBody()

#define ASSERT_LT(a, b)                                                                            \
    do {                                                                                           \
        const auto va_ = (a);                                                                      \
        const auto vb_ = (b);                                                                      \
        if (!(va_ < vb_)) {                                                                        \
            std::printf("%s:%d: Failure\nExpected: (%s) < (%s), actual: %lld vs %lld\n", __FILE__, __LINE__, #a, #b, \
                        (long long)va_, (long long)vb_);                                           \
            ::testing::Registry::get().failed_now = true;                                          \
            return;                                                                                \
        }                                                                                          \
    } while (0)

#define RUN_ALL_TESTS() ::testing::run_all()
#endif
