// Multi-GPU group of the C ABI (SURVEY 8b `xx_comm_init`): one NCCL communicator per context, used by
//   * pb2_run          : cross-rank dual-averaging statistics (2 floats per rank and adapting transition), the analogue of
//                        experimental_reduce_chain_axis_names (tfp/mcmc/dual_averaging_step_size_adaptation.py:259-261,
//                        422-435; internal/distribute_lib.py:147-162 reduce_logsumexp);
//   * pb2_rowshard_leapfrog : SimpleLeapfrogIntegrator (tfp/mcmc/internal/leapfrog_integrator.py:280-355) for a target
//                        whose rows are sharded over the ranks: the per-leapfrog gradient psum of
//                        internal/distribute_lib.py:179-242, enqueued from C between the gradient kernel and the fused
//                        prior + kick kernel -- no host code between the L leapfrogs.
// NCCL is bound at run time (dlopen): a process that already carries an NCCL (torch.distributed's) shares it, a plain C
// host gets the system library.  Single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "pb2_internal.h"

namespace pb2 {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  if (api.handle || !api.err.empty()) return &api;
  // prefer an NCCL this process has already loaded (torch's bundled one), then the system library
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
    if (api.handle) break;
  }
  if (!api.handle)
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
  if (!api.handle) {
    const char* e = dlerror();
    api.err = std::string("cannot load libnccl.so.2: ") + (e ? e : "?");
    return &api;
  }
  auto sym = [&](const char* s) -> void* {
    void* p = dlsym(api.handle, s);
    if (!p && api.err.empty()) api.err = std::string("libnccl has no symbol ") + s;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return &api;
}

static int check_nccl(pb2_ctx* ctx, NcclApi* api, ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return PB2_OK;
  return set_error(ctx, PB2_ERR_CUDA, std::string(what) + ": " + (api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
}

int comm_allreduce_sum(pb2_ctx* ctx, float* d_buf, size_t n) {
  if (!ctx->comm) return PB2_OK;
  NcclApi* api = nccl_api();
  return check_nccl(ctx, api, api->AllReduce(d_buf, d_buf, n, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream),
                    "ncclAllReduce");
}

int comm_allgather(pb2_ctx* ctx, const float* d_send, float* d_recv, size_t n_per_rank) {
  NcclApi* api = nccl_api();
  return check_nccl(ctx, api, api->AllGather(d_send, d_recv, n_per_rank, ncclFloat32, (ncclComm_t)ctx->comm, ctx->stream),
                    "ncclAllGather");
}

// ---- row-sharded leapfrog: fused prior + kick + drift around the all-reduced likelihood gradient
//   first: v = m + (eps/2) g ; x += eps v                                            (leapfrog_integrator.py:280-287)
//   step : g = -x + G ; lp = log N(x; 0, I) + loglik ; v += eps g ; x += eps v  (last: m_out = v - (eps/2) g instead)
__global__ void rowshard_first_kernel(int B, int D, const float* step, int step_kind, const float* m, const float* g,
                                      const float* x_in, float* v, float* x) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * D) return;
  const int b = (int)(i / D), d = (int)(i - (size_t)b * D);
  const float eps = step_kind == 0 ? step[0] : (step_kind == 1 ? step[d] : step[b]);
  const float vv = m[i] + (0.5f * eps) * g[i];
  v[i] = vv;
  x[i] = x_in[i] + eps * vv;
}

__global__ void rowshard_step_kernel(int last, int B, int D, const float* step, int step_kind, const float* packed,
                                     float* v, float* x, float* grad, float* logp, float* m_out) {
  const int b = blockIdx.x;
  float part = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const size_t i = (size_t)b * D + d;
    const float eps = step_kind == 0 ? step[0] : (step_kind == 1 ? step[d] : step[b]);
    const float t = x[i];
    const float gn = -t + packed[(size_t)b * (D + 1) + d];
    grad[i] = gn;
    part += -0.5f * t * t - kHalfLog2Pi;
    const float vv = v[i] + eps * gn;
    v[i] = vv;
    if (last) m_out[i] = vv - (0.5f * eps) * gn;
    else x[i] = t + eps * vv;
  }
  part = warp_sum(part);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
    logp[b] = s + packed[(size_t)b * (D + 1) + D];
  }
}

}  // namespace pb2

using namespace pb2;

extern "C" {

int pb2_comm_unique_id(void* out_id) {
  if (!out_id) return set_error(nullptr, PB2_ERR_INVALID, "pb2_comm_unique_id: NULL argument");
  NcclApi* api = nccl_api();
  if (!api->err.empty()) return set_error(nullptr, PB2_ERR_UNSUPPORTED, api->err);
  ncclUniqueId id;
  if (int rc = check_nccl(nullptr, api, api->GetUniqueId(&id), "ncclGetUniqueId")) return rc;
  static_assert(sizeof(ncclUniqueId) == PB2_COMM_ID_BYTES, "pb2.h declares the size of the NCCL unique id");
  std::memcpy(out_id, &id, sizeof(id));
  return PB2_OK;
}

int pb2_comm_init(pb2_ctx* ctx, int nranks, int rank, const void* nccl_unique_id) {
  if (!ctx || !nccl_unique_id || nranks < 1 || rank < 0 || rank >= nranks)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_comm_init: bad argument");
  if (ctx->comm) return set_error(ctx, PB2_ERR_INVALID, "pb2_comm_init: this context already has a communicator");
  NcclApi* api = nccl_api();
  if (!api->err.empty()) return set_error(ctx, PB2_ERR_UNSUPPORTED, api->err);
  cudaSetDevice(ctx->device);
  ncclUniqueId id;
  std::memcpy(&id, nccl_unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  if (int rc = check_nccl(ctx, api, api->CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank")) return rc;
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_size = nranks;
  return PB2_OK;
}

int pb2_comm_destroy(pb2_ctx* ctx) {
  if (!ctx) return PB2_ERR_INVALID;
  if (!ctx->comm) return PB2_OK;
  NcclApi* api = nccl_api();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ncclResult_t r = api->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_size = 1;
  ctx->comm_rank = 0;
  return check_nccl(ctx, api, r, "ncclCommDestroy");
}

int pb2_comm_size(pb2_ctx* ctx) { return ctx ? ctx->comm_size : PB2_ERR_INVALID; }
int pb2_comm_rank(pb2_ctx* ctx) { return ctx ? ctx->comm_rank : PB2_ERR_INVALID; }

int pb2_comm_allreduce_sum(pb2_ctx* ctx, float* d_buf, long long n) {
  if (!ctx || !d_buf || n < 0) return set_error(ctx, PB2_ERR_INVALID, "pb2_comm_allreduce_sum: bad argument");
  if (!ctx->comm) return set_error(ctx, PB2_ERR_INVALID, "pb2_comm_allreduce_sum: no communicator (pb2_comm_init)");
  cudaSetDevice(ctx->device);
  return comm_allreduce_sum(ctx, d_buf, (size_t)n);
}

int pb2_rowshard_leapfrog(pb2_ctx* ctx, const void* d_planes, const float* d_X, const float* d_y, int N, int D, int DP,
                          int B, const float* d_m, const float* d_x, const float* d_logp, const float* d_grad,
                          const float* d_step, int step_kind, int num_steps, int reduce_over_ranks, float* d_m_out,
                          float* d_x_out, float* d_logp_out, float* d_grad_out) {
  if (reduce_over_ranks && ctx && !ctx->comm)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_rowshard_leapfrog: reduce_over_ranks needs a communicator (pb2_comm_init)");
  if (!ctx || (!d_planes && !d_X) || !d_y || !d_m || !d_x || !d_logp || !d_grad || !d_step || !d_m_out || !d_x_out ||
      !d_logp_out || !d_grad_out || N < 0 || B < 1 || D < 1 || num_steps < 1 || step_kind < 0 || step_kind > 2)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_rowshard_leapfrog: bad argument");
  cudaSetDevice(ctx->device);
  const size_t nBD = (size_t)B * D;
  const size_t need = sizeof(float) * (nBD + (size_t)B * (D + 1));
  if (need > ctx->rs_bytes) {
    if (ctx->d_rs) cudaFree(ctx->d_rs);
    ctx->d_rs = nullptr;
    ctx->rs_bytes = 0;
    if (int rc = check_cuda(ctx, cudaMalloc((void**)&ctx->d_rs, need), "cudaMalloc(rowshard leapfrog scratch)")) return rc;
    ctx->rs_bytes = need;
  }
  float* v = ctx->d_rs;
  float* packed = v + nBD;
  const unsigned nb = (unsigned)((nBD + 255) / 256);
  rowshard_first_kernel<<<nb, 256, 0, ctx->stream>>>(B, D, d_step, step_kind, d_m, d_grad, d_x, v, d_x_out);
  ctx->launches += 1;
  for (int l = 0; l < num_steps; ++l) {
    int rc = d_planes ? pb2_rowshard_logistic_grad_tc(ctx, d_planes, d_y, N, D, d_x_out, B, packed)
                      : pb2_rowshard_logistic_grad(ctx, d_X, d_y, N, D, DP, d_x_out, B, packed);
    if (rc) return rc;
    if (reduce_over_ranks)   // psum over the data axis
      if (int rc2 = comm_allreduce_sum(ctx, packed, (size_t)B * (D + 1))) return rc2;
    rowshard_step_kernel<<<B, 128, 0, ctx->stream>>>(l + 1 == num_steps ? 1 : 0, B, D, d_step, step_kind, packed, v,
                                                     d_x_out, d_grad_out, d_logp_out, d_m_out);
    ctx->launches += 1;
  }
  return check_cuda(ctx, cudaGetLastError(), "pb2_rowshard_leapfrog kernels");
}

}  // extern "C"
