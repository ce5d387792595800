#include <string.h>

#include "../../include/sdrm_b200.h"
#include "host_util.h"

static thread_local char g_last_error[512] = "";

int sdrm_fail(int code, const char* msg) {
  strncpy(g_last_error, msg ? msg : "", sizeof g_last_error - 1);
  g_last_error[sizeof g_last_error - 1] = 0;
  return code;
}

extern "C" const char* sdrm_last_error(void) { return g_last_error; }
