// hyperbo_b200: kernels + launch code of the fp64 (DMMA tile products) engine.
// The .inc files are shared with the other precision; only `Real` differs.
#include "hb_internal.cuh"

#define HB_F64 1
#define HB_MIN_CTAS 2
namespace hb {
namespace f64 {
using namespace hb::host;
using Real = double;
using Real2 = double2;
__device__ __forceinline__ Real2 make_real2(Real a, Real b) { return make_double2(a, b); }
#include "hb_device.inc"
#include "hb_kernels.inc"
#include "hb_fused.inc"
#include "hb_host.inc"
#include "hb_mrhs.inc"
}  // namespace f64
}  // namespace hb
