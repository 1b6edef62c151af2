// Library identification and error text (C-ABI: include/dyt_b200.h).
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"

extern "C" int dyt_version(void) { return DYT_ABI_VERSION; }
extern "C" const char* dyt_last_error(void) { return dyt::last_error_buf(); }
