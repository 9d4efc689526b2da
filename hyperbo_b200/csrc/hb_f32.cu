// hyperbo_b200: kernels + launch code of the fp32 (3xTF32 tile products) engine.
// The .inc files are shared with the other precision; only `Real` differs.
#include "hb_internal.cuh"

#define HB_F64 0
#define HB_MIN_CTAS 3
namespace hb {
namespace f32 {
using namespace hb::host;
using Real = float;
using Real2 = float2;
__device__ __forceinline__ Real2 make_real2(Real a, Real b) { return make_float2(a, b); }
#include "hb_device.inc"
#include "hb_kernels.inc"
#include "hb_fused.inc"
#include "hb_host.inc"
#include "hb_mrhs.inc"
}  // namespace f32
}  // namespace hb
