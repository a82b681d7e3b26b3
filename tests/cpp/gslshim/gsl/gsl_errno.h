// tests/cpp/gslshim - TEST INFRASTRUCTURE: declarations of the few GSL entry points the reference's headers and plugins
// name, so that its plugin sources can be COMPILED against the drop-in headers of include/mfb/ in a container without GSL
// (tests/test_plugin_dropin.py).  No definitions: nothing here computes.
#pragma once
extern "C" {
typedef void gsl_error_handler_t(const char* reason, const char* file, int line, int gsl_errno);
gsl_error_handler_t* gsl_set_error_handler_off(void);
gsl_error_handler_t* gsl_set_error_handler(gsl_error_handler_t* h);
const char* gsl_strerror(const int gsl_errno);
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1 };
}
