// Library-wide state of the C ABI: version and the last error string.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace amqb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace amqb

extern "C" {
const char* amqb_last_error_string(void) { return amqb::g_err; }
int amqb_version(void) { return 100; }
}
