// xy_fields (fields.rs:63-127) -- implemented in a later milestone.
#include "../../include/fem2d.h"
extern "C" int fem2d_xy_fields(const fem2d_domain_view*, int, int, uint32_t, const double*, uint64_t, uint64_t*, uint32_t*, double*, double*) {
    return FEM2D_ERR_UNSUPPORTED;
}
