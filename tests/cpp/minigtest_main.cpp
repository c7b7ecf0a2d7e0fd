// main() of the reference-test runners (see minigtest/gtest/gtest.h)
#define MINIGTEST_MAIN
#include <gtest/gtest.h>
