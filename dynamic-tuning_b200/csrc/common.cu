// Library identification and error text (C-ABI: include/dyt_b200.h).
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"

extern "C" int dyt_version(void) { return DYT_ABI_VERSION; }
extern "C" const char* dyt_last_error(void) { return dyt::last_error_buf(); }

extern "C" int dyt_configure(int option, int value) {
  switch (option) {
    case DYT_OPT_PDL:
      dyt::pdl_option().store(value != 0 ? 1 : 0);
      return dyt::DYT_OK;
    case DYT_OPT_GEMM_TAIL_SPLIT:
      dyt::tail_split_option().store(value != 0 ? 1 : 0);
      return dyt::DYT_OK;
    case DYT_OPT_FUSE_ADAPTER_DOWN:
      dyt::fuse_down_option().store(value != 0 ? 1 : 0);
      return dyt::DYT_OK;
    case DYT_OPT_ATTN_SPLIT:
      dyt::attn_split_option().store(value != 0 ? 1 : 0);
      return dyt::DYT_OK;
    case DYT_OPT_FUSE_ADAPTER_UP:
      dyt::fuse_up_option().store(value != 0 ? 1 : 0);
      return dyt::DYT_OK;
    case DYT_OPT_SM_LIMIT:
      dyt::sm_limit_option().store(value < 0 ? 0 : value);
      return dyt::DYT_OK;
    case DYT_OPT_SIDE_PLAN:
      dyt::side_plan_option().store(value & 7);
      return dyt::DYT_OK;
    case DYT_OPT_TILE_ORDER:
      dyt::tile_order_option().store(value & 15);
      return dyt::DYT_OK;
    default:
      return dyt::fail(dyt::DYT_EINVAL, "dyt_configure: unknown option %d", option);
  }
}
