// pldp.cu - placeholder (filled in below in the same round)
#include "wg_common.h"
void wg_pldp_release(wg_ctx *) {}
