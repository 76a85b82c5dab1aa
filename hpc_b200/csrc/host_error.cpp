// host_error.cpp — error plumbing of the host-only graph library (libspmm_b200_graph.so): the same two symbols
// capi.cu provides inside libspmm_b200.so, without any CUDA dependency.
#include <stdarg.h>
#include <stdio.h>

namespace spmm_b200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace spmm_b200

extern "C" const char *spmm_b200_last_error(void) { return spmm_b200::g_err; }
