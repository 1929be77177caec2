// K2, impl 2: TMA + mbarrier + DMMA Gram.  (placeholder until the cp.async version is validated on hardware)
#include "kernels.cuh"
namespace ruviii {
int launch_gram_tma_prepare(const GramPlan&, const double*, int64_t, void*) {
    throw Error(6, "gram impl 2 (TMA) is not built into this library yet");
}
int launch_gram_tma(const GramPlan&, const void*, const int2*, double*, cudaStream_t) {
    throw Error(6, "gram impl 2 (TMA) is not built into this library yet");
}
}  // namespace ruviii
