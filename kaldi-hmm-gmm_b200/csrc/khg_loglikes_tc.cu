// placeholder: tcgen05 dense log-likelihood kernel (filled in next)
#include "khg_internal.h"
namespace khg {
bool tc_supported(const khg_model *) { return false; }
khg_status tc_pack_build(khg_model *) { set_error("tcgen05 kernel not built"); return KHG_ERR_UNSUPPORTED; }
void tc_pack_free(khg_model *) {}
khg_status tc_loglikes(khg_model *, const float *, int64_t, float, float *, int64_t) { set_error("tcgen05 kernel not built"); return KHG_ERR_UNSUPPORTED; }
}
