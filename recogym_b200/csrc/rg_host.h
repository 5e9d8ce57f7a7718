// rg_host.h -- host-side handle behind the opaque rg_env of include/recogym_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "rg_device.cuh"
#include "rg_tensor.cuh"

struct rg_env_impl {
    int device = 0;
    int num_sms = 0;
    rg::EnvDev dev{};
    // owned device allocations
    double *d_GammaT = nullptr, *d_mu_o = nullptr, *d_beta = nullptr, *d_mu_b = nullptr, *d_GammaB = nullptr;
    float *d_GammaT32 = nullptr, *d_mu_o32 = nullptr;
    float *d_Gamma32 = nullptr;          // [nsub][K+1][128] fp32 tiles (fast path)
    rg::tc::opnd_t *d_GTC = nullptr;     // [tiles][nslab][2][KCs][TILE_N/8][8][EPC] hi/lo UMMA operand tiles (rg_tensor.cuh)
    bool tc_operands_ok = true;          // Gamma', mu' inside the operand element range (else: CUDA-core contraction)
    double *d_om_scratch = nullptr;      // RG_MATH_TENSOR: omega rows of the resident slots
    size_t om_scratch_bytes = 0;
    double *d_csum_scratch = nullptr;    // CUDA-core form with large P x K: re-sampling partial sums spilled to global memory
    size_t csum_scratch_bytes = 0;
    float *d_gsum = nullptr;             // RG_MATH_TENSOR: pass-1 mass sums per 32-product group of the resident slots
    size_t gsum_bytes = 0;
    int2 *d_views = nullptr;             // RG_POLICY_USER_VIEWS: sorted (product, views) lists of the resident users
    size_t views_lists = 0;              // lists allocated (each view_cap entries)
    int view_cap = 0;
    unsigned long long *diag = nullptr;  // [4] device diagnostics (f64-ambiguous samples, ...)
    char *d_host_arena = nullptr;        // rg_rollout_host: device staging of the table and the outputs, grown on demand
    size_t host_arena_bytes = 0;
    double *d_hole_cdf = nullptr;        // RG_POLICY_EPS_TABLE: [P] table of the exploring draw's cdf (built on first use)
    double *d_t4 = nullptr;              // tier 4 (oracle-order re-draw): [kT4Slots][Ppad] masses + kT4Slots int locks
};

struct rg_env {
    rg_env_impl impl;
};

namespace rg {
cudaError_t launch_rollout_exact(const rg_env_impl &e, const PolicyDev &pol, long long user_begin,
                                 long long n_organic_users, long long n_total, const RolloutOutDev &out,
                                 cudaStream_t stream);
cudaError_t launch_rollout_fast(rg_env_impl &e, const PolicyDev &pol, long long user_begin,
                                long long n_organic_users, long long n_total, const RolloutOutDev &out,
                                cudaStream_t stream, bool tensor);
cudaError_t launch_batch(const rg_env_impl &e, int mode, long long N, const long long *user_id,
                         const int32_t *action, int32_t *state, int32_t *t, int32_t *last_view, double *omega,
                         int32_t *reward, int32_t *sess_v, int32_t *sess_len, int sess_cap, int32_t *n_unfinished,
                         cudaStream_t stream);
cudaError_t launch_tc_logits(const rg_env_impl &e, const double *omega, const float *shift, float *out,
                             cudaStream_t stream);
cudaError_t launch_row_offsets(const int32_t *n_events, long long n, long long n_organic_users, long long *row_offset,
                               long long *tile_scratch, cudaStream_t stream);
cudaError_t launch_order(const int4 *log, const unsigned long long *log_count, long long log_capacity,
                         const long long *row_offset, int4 *ordered, long long ordered_capacity, int num_sms,
                         cudaStream_t stream);
cudaError_t launch_columns(const int4 *ordered, const long long *row_offset, long long n_total,
                           long long n_organic_users, long long user_begin, const int32_t *last_view, int prev_last_view,
                           int policy_kind, int P, const double *ps_table, const int32_t *table, double epsilon,
                           int time_mode, uint64_t seed, double time_mu, double time_sigma,
                           const rg_log_columns &c, int train_mode, int32_t *carry_out, int num_sms, cudaStream_t stream);
cudaError_t launch_count_stats(const rg_log_columns &c, long long rows, int P, int train_mode, int64_t *co,
                               int64_t *pulls, int64_t *clicks, int num_sms, cudaStream_t stream);
cudaError_t launch_row_argmax(int kind, int P, const int64_t *A, const int64_t *B, int32_t *table, double *ps,
                              cudaStream_t stream);
cudaError_t launch_sparse_count_stats(const rg_log_columns &c, long long rows, int train_mode, const rg_sparse_counts *co,
                                      const rg_sparse_counts *bandit, int num_sms, cudaStream_t stream);
cudaError_t launch_sparse_rehash(const rg_sparse_counts &from, const rg_sparse_counts &to, int num_sms, cudaStream_t stream);
cudaError_t launch_sparse_policy(int kind, int P, const rg_sparse_counts &s, int32_t *table, double *ps,
                                 unsigned long long *scratch, int num_sms, cudaStream_t stream);
cudaError_t launch_ips(const rg_log_columns &c, long long rows, long long u_end, const int32_t *target, double *part,
                       int nblocks, double *out, cudaStream_t stream);
cudaError_t launch_recall(const rg_log_columns &c, long long rows, long long u_end, const int32_t *target,
                          const int32_t *topk, int k, unsigned long long *out2, int num_sms, cudaStream_t stream);
cudaError_t launch_spec_eval(int kind, long long n, const double *in, double *out, cudaStream_t stream);
}  // namespace rg

namespace rg {
int set_error(int code, const char *msg);   // sets rg_last_error()'s text (rg_capi.cu), returns code
}
