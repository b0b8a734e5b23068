// Error reporting and version for the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void lcao_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

unsigned long long g_lcao_launches = 0;

bool lcao_pdl_enabled() {
  static const bool on = [] { const char* s = getenv("LCAO_PDL"); return !(s && s[0] == '0'); }();
  return on;
}

extern "C" int lcao_version(void) { return 100; }
extern "C" int64_t lcao_launch_count(void) { return (int64_t)g_lcao_launches; }
extern "C" const char* lcao_last_error(void) { return g_err; }
