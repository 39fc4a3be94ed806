// fp64 "compat" kernels: compiled with -fmad=false so that every operation is a separately
// rounded IEEE operation, as in the reference's Python floats
#include "ppb_launch_impl.cuh"
namespace ppb {
template struct Launch<double>;
}  // namespace ppb
