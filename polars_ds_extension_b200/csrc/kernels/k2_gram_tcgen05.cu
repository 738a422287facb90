// K2b — placeholder until the tcgen05 kernel lands (replaced in the next milestone).
#include "../common.h"
#include "kernels.h"
namespace pdsb {
bool moments_tcgen05_supported(const float*, int64_t, const float*, int64_t, int64_t, int, int) { return false; }
int moments_tcgen05_f32(const float*, int64_t, const float*, int64_t, const float*, int64_t, int, int, double*, cudaStream_t) { return -1; }
}  // namespace pdsb
