// placeholder, replaced below
#include "ifb_internal.h"
extern "C" {
int ifb_fit_device(int32_t, const float *, int64_t, int32_t, int64_t, int32_t, const ifb_fit_params *, ifb_forest **, void *) { ifb::set_error("fit not built"); return IFB_ESTATE; }
int ifb_fit_host(int32_t, const float *, int64_t, int32_t, int64_t, int32_t, const ifb_fit_params *, ifb_forest **) { ifb::set_error("fit not built"); return IFB_ESTATE; }
}
