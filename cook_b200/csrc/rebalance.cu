// rebalance.cu — rebalancer preemption-victim search.
// Placeholder until the kernels land: returns an explicit error (never a CPU
// fallback).
#include "common.cuh"

extern "C" int32_t cook_rebalance(cook_pool* pool, const cook_running_soa*, const cook_jobs_soa*,
                                  const int64_t*, const int32_t*, const cook_host_table*,
                                  const cook_groups*, const cook_user_table*,
                                  const cook_rebalance_params*, cook_decision*, int32_t*, int32_t*) {
  return set_err(pool, COOK_E_UNSUPPORTED_CONSTRAINT, "cook_rebalance: not implemented yet");
}
