// Library-wide entry points of libprobingrag.so: version and thread-local error text.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_pr_error[512] = "";

void pr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_pr_error, sizeof(g_pr_error), fmt, ap);
    va_end(ap);
}

extern "C" int pr_version(void) { return PR_VERSION; }

extern "C" const char *pr_last_error(void) { return g_pr_error; }
