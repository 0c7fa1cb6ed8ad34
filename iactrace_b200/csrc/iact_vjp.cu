// K6: vector-Jacobian product of render (placeholder until the backward kernel lands).
#include "iact_common.cuh"

extern "C" int iact_render_vjp(const IactScene*, const IactFacets*, const float*, const float*, int, int,
                               const float*, const IactGrads*, void*) {
    iact_set_error("iact_render_vjp: not implemented yet");
    return IACT_ERR_UNSUPPORTED;
}
