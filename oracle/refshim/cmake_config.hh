// oracle/refshim/cmake_config.hh - TEST INFRASTRUCTURE.  Hand-written stand-in for the header the reference's CMake
// generates from include/cmake_config.hh.in: the reference's default build options (CODE_PRECISION=DOUBLE,
// CONVOLVER_TYPE=ORSZAG, PLT off; CMakeLists.txt:89-96).
#pragma once
#define USE_PRECISION_DOUBLE
#ifdef MFB_TEST_CONVOLVER_NAIVE   /* CONVOLVER_TYPE=NAIVE builds of tests/cpp/Makefile */
#define USE_CONVOLVER_NAIVE
#else
#define USE_CONVOLVER_ORSZAG
#endif
#ifdef __cplusplus
constexpr char CMAKE_BUILDTYPE_STR[] = "oracle-shim";
constexpr char CMAKE_PRECISION_STR[] = "double";
#ifdef MFB_TEST_CONVOLVER_NAIVE
constexpr char CMAKE_CONVOLVER_STR[] = "naive";
#else
constexpr char CMAKE_CONVOLVER_STR[] = "Orszag3/2";
#endif
constexpr char CMAKE_PLT_STR[] = "PLT corr. off";
extern "C" {
extern const char* GIT_TAG;
extern const char* GIT_REV;
extern const char* GIT_BRANCH;
}
#endif
