// rg_rollout_fast.cu -- RG_MATH_FAST rollout (placeholder: routes to the exact kernel until the
// CTA-batched fp32 kernel lands).
#include "rg_host.h"
namespace rg {
cudaError_t launch_rollout_fast(const rg_env_impl &e, const PolicyDev &pol, long long user_begin,
                                long long n_organic_users, long long n_total, const RolloutOutDev &out,
                                cudaStream_t stream) {
    return launch_rollout_exact(e, pol, user_begin, n_organic_users, n_total, out, stream);
}
}  // namespace rg
