// Library-wide state: ABI version and the thread-local error string.
#include "common.cuh"
#include <stdarg.h>
#include "../../include/calm_b200.h"

static thread_local char g_err[512] = "";

void calm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int* g_calm_err_flag = nullptr;  // bring-up only: int the kernels write a barrier id into before trapping on a protocol time-out
#ifdef CALM_BRINGUP
extern "C" int32_t calm_set_error_flag_buffer(int32_t* device_int) { g_calm_err_flag = device_int; return CALM_OK; }
#endif

extern "C" int32_t calm_abi_version(void) { return CALM_ABI_VERSION; }
extern "C" const char* calm_last_error(void) { return g_err; }
