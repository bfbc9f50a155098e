// Internal host-side declarations shared by the translation units of libpb2.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/pb2.h"
#include <vector>

#include "pb2_chain_kernel.cuh"
#include "pb2_user_target.cuh"
#include "pb2_targets.cuh"

struct pb2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  int max_smem_optin = 227 * 1024;
  long long launches = 0;
  int dense_variant = 0;           // A/B switch of the dense-Gaussian kernels, see pb2_ctx_set_int in pb2.h
  int synchronous_run = 0;         // 1: pb2_run waits for the stream before returning (default: enqueue and return)
  uint32_t* h_keys = nullptr;      // pinned staging of the per-transition seeds (the H2D copy outlives pb2_run)
  size_t h_keys_bytes = 0;
  cudaEvent_t keys_copied = nullptr;
  std::string err;
  int* d_queue = nullptr;          // dynamic chain queue counter
  float* d_ckpt = nullptr;         // global checkpoint scratch (block-group targets)
  size_t ckpt_bytes = 0;
  uint32_t* d_sched = nullptr;     // key schedule scratch
  size_t sched_bytes = 0;
  uint32_t* d_step_keys = nullptr;
  size_t step_keys_bytes = 0;
  float* d_step_seq = nullptr;     // per-transition scalar step sizes (dual averaging)
  size_t step_seq_bytes = 0;
  float* d_partial = nullptr;      // [2] dual-averaging partial, followed by [2 * comm_size] gathered partials
  // multi-GPU group (pb2_comm.cu): NCCL communicator of this context (ncclComm_t), nullptr on a single GPU
  void* comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  float* d_rs = nullptr;           // row-sharded leapfrog scratch: v [B, D] + packed [B, D + 1]
  size_t rs_bytes = 0;
  // peer-memory group of the row-sharded leapfrog (pb2_comm.cu PeerGroup): every rank's packed gradient buffer mapped
  // into every process (CUDA IPC), so the cross-rank sum is done by the consuming kernel itself
  void* peer = nullptr;
  int rowshard_collective = 1;     // 0: NCCL all-reduce between the kernels; 1: peer-memory reduction in the step kernel
};

struct pb2_target {
  pb2_ctx* ctx = nullptr;
  int kind = 0;
  int dim = 0;
  int n_rows = 0;
  float* d_a = nullptr;
  float* d_b = nullptr;
  float scalar = 0.f;
  unsigned char* d_tc = nullptr;   // logistic: X~ pre-split into the tensor-core operand planes (pb2_logistic_tc.cu)
  size_t tc_bytes = 0;
  // user-defined target (pb2_user.cu): source + the loaded run-time builds, one per variant (plain / ScaledT /
  // TransformedT, built on first use), and their kernels indexed by Mode
  std::string user_source, user_include;
  int user_flags = 0;
  void* user_lib[3] = {nullptr, nullptr, nullptr};
  void* user_kernels[3][4] = {};
};

namespace pb2 {

int set_error(pb2_ctx* ctx, int code, const std::string& msg);
int check_cuda(pb2_ctx* ctx, cudaError_t e, const char* what);

// pb2_chain_kernels.cu
int launch_chain(pb2_ctx* ctx, const pb2_target* tgt, int mode, ChainParams& p, const PrimIO& io);

// pb2_tile.cu (tcgen05 tile kernels, dense Gaussian: dispatch + the 128-chain HMC kernel)
bool tile_path_supported(const pb2_ctx* ctx, const pb2_target* tgt, int mode, const ChainParams& p);
int launch_tile_chain(pb2_ctx* ctx, const pb2_target* tgt, int mode, ChainParams& p);
// pb2_tile_nuts.cu (64-chain tiles: lock-step and asynchronous-lane NUTS)
int launch_tile_nuts(pb2_ctx* ctx, const pb2_target* tgt, ChainParams& p);

// pb2_user.cu (user-defined targets: NVRTC build + load)
int user_elements_per_lane(int dim);
constexpr int kUserVariants = 3;
int user_compile(pb2_ctx* ctx, const char* source, int dim, const char* include_dir, int variant, int flags,
                 std::vector<char>& cubin);
int user_target_ensure(pb2_ctx* ctx, pb2_target* t, int variant);
void user_target_unload(pb2_target* t);

// pb2_logistic_tc.cu
int launch_logistic_tc(pb2_ctx* ctx, pb2_target* tgt, int B, const float* d_x, float* d_lp, float* d_g);
size_t rowshard_tc_planes_bytes(int N);
int launch_rowshard_tc_prepare(pb2_ctx* ctx, const float* d_X, int N, int D, int DP, unsigned char* d_planes);
int launch_rowshard_tc(pb2_ctx* ctx, const unsigned char* d_planes, const float* d_y, int N, int D, const float* d_theta,
                       int B, float* d_packed);

// pb2_comm.cu (no-ops / errors without a communicator)
int comm_allreduce_sum(pb2_ctx* ctx, float* d_buf, size_t n);
int comm_allgather(pb2_ctx* ctx, const float* d_send, float* d_recv, size_t n_per_rank);

// pb2_misc.cu
int launch_hmc_sched(pb2_ctx* ctx, const uint32_t* d_step_keys, int T, int n_parts, int layout, uint32_t* d_out);
int launch_nuts_sched(pb2_ctx* ctx, const uint32_t* d_step_keys, int T, int n_parts, int max_depth, int layout,
                      uint32_t* d_out);
inline int hmc_sched_stride(int n_parts) { return 2 * (n_parts + 1); }
inline int nuts_sched_stride(int n_parts, int max_depth) {
  return 2 * n_parts + 6 * max_depth + 2 * ((1 << max_depth) - 1);
}
int launch_da_partial(pb2_ctx* ctx, const float* d_lar, int B, float* d_partial);
int launch_da_apply(pb2_ctx* ctx, const float* d_partials, int n, long long B_global, float* d_state,
                    float* d_step_out, float* d_step_seq_next);
int launch_fill_step_seq(pb2_ctx* ctx, float* d_seq, const float* d_step, int n);
int launch_scale_rows(pb2_ctx* ctx, float* d_a, size_t rows, int D, const float* d_s, int div);
int launch_running_moments(pb2_ctx* ctx, const float* d_x, long long rows, int D, float* d_state);
int launch_rng_fill(pb2_ctx* ctx, Key key, Key key_hi, long long n, int layout, int what, float lo, float hi,
                    int ilo, int ihi, void* out);
int launch_ess(pb2_ctx* ctx, const float* d_states, int N, int B, int D, float thr, int use_thr, int max_lag,
               int pairs, int cross, float* d_mean_scratch, float* d_out);
int launch_rhat(pb2_ctx* ctx, const float* d_states, int N, int B, int D, int split, float* d_out);

}  // namespace pb2
