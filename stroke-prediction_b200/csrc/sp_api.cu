// sp_api.cu — library-level entry points: version and thread-local error message.
#include "sp_common.cuh"
#include <math.h>

static thread_local char g_sp_error[512] = "";

void sp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_sp_error, sizeof(g_sp_error), fmt, ap);
    va_end(ap);
}

extern "C" int sp_version(void) { return SP_VERSION; }
extern "C" const char* sp_last_error(void) { return g_sp_error; }
