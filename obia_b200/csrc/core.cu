// Error plumbing + version for the obia_b200 C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace obia {

char *err_buf()
{
    static thread_local char buf[512] = {0};
    return buf;
}

int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace obia

extern "C" const char *obia_b200_last_error(void) { return obia::err_buf(); }
extern "C" int obia_b200_version(void) { return 100; }
