#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "no error";

void d3f_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int d3f_version(void) { return 100; }
extern "C" const char* d3f_last_error_string(void) { return g_err; }

unsigned long long g_d3f_launches = 0;
extern "C" unsigned long long d3f_launch_count(void) { return g_d3f_launches; }
