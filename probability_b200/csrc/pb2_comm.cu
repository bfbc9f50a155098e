// Multi-GPU group of the C ABI (SURVEY 8b `xx_comm_init`): one NCCL communicator per context, used by
//   * pb2_run          : cross-rank dual-averaging statistics (2 floats per rank and adapting transition), the analogue of
//                        experimental_reduce_chain_axis_names (tfp/mcmc/dual_averaging_step_size_adaptation.py:259-261,
//                        422-435; internal/distribute_lib.py:147-162 reduce_logsumexp);
//   * pb2_rowshard_leapfrog : SimpleLeapfrogIntegrator (tfp/mcmc/internal/leapfrog_integrator.py:280-355) for a target
//                        whose rows are sharded over the ranks: the per-leapfrog gradient psum of
//                        internal/distribute_lib.py:179-242, enqueued from C between the gradient kernel and the fused
//                        prior + kick kernel -- no host code between the L leapfrogs.  By default the psum is not a
//                        separate collective at all: every rank's packed gradient buffer is mapped into every process
//                        (CUDA IPC over NVLink), the gradient pass publishes into it, and the fused prior + kick + drift
//                        kernel sums the ranks' buffers itself (fixed rank order: replicas stay bit-identical), waiting on
//                        per-rank sequence flags written over NVLink.  NCCL stays as the selectable baseline
//                        (pb2_ctx_set_int "rowshard_collective" 0) and carries the one-time exchange of the IPC handles.
// NCCL is bound at run time (dlopen): a process that already carries an NCCL (torch.distributed's) shares it, a plain C
// host gets the system library.  Single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

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

// ---- peer-memory group: rank r's buffer = [2 slots][cap floats] packed gradients + [kMaxPeers] sequence flags
constexpr int kMaxPeers = 8;
struct PeerView {
  const float* data[kMaxPeers];   // every rank's buffer, mapped here (own rank: the local allocation)
  const unsigned* flags;          // MY flags: flags[r] = last sequence number rank r has published
  unsigned* err;                  // MY error word (set when a wait times out)
  int n;
  unsigned seq;                   // the sequence number this leapfrog waits for
  size_t slot_off;                // float offset of the slot in use
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// after the gradient pass has written this rank's slot: tell every rank (one flag per peer, over NVLink)
__global__ void peer_signal_kernel(unsigned* const* peer_flags, int n, int rank, unsigned seq) {
  const int r = threadIdx.x;
  if (r < n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[r] + rank), "r"(seq) : "memory");
  }
}

template <bool kPeer>
__global__ void rowshard_step_kernel(int last, int B, int D, const float* step, int step_kind, const float* packed,
                                     const PeerView pv, float* v, float* x, float* grad, float* logp, float* m_out) {
  const int b = blockIdx.x;
  if constexpr (kPeer) {
    // wait until every rank has published this leapfrog's gradient (bounded: a lost peer sets the error word instead of
    // hanging the GPU)
    if (threadIdx.x < pv.n) {
      long long spins = 0;
      while ((int)(ld_acquire_sys(pv.flags + threadIdx.x) - pv.seq) < 0) {
        if (++spins > (1ll << 23)) { atomicExch(pv.err, 1u); break; }   // seconds: far beyond any skew between ranks
        __nanosleep(64);
      }
    }
    __syncthreads();
  }
  auto summed = [&](int col) -> float {   // psum over the data axis, rank order 0 .. n-1 on every rank
    if constexpr (kPeer) {
      const size_t o = pv.slot_off + (size_t)b * (D + 1) + col;
      float s = ld_relaxed_sys(pv.data[0] + o);
      for (int r = 1; r < pv.n; ++r) s += ld_relaxed_sys(pv.data[r] + o);
      return s;
    } else {
      return packed[(size_t)b * (D + 1) + col];
    }
  };
  float part = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const size_t i = (size_t)b * D + d;
    const float eps = step_kind == 0 ? step[0] : (step_kind == 1 ? step[d] : step[b]);
    const float t = x[i];
    const float gn = -t + summed(d);
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
    logp[b] = s + summed(D);
  }
}

struct PeerGroup {
  int n = 0, rank = 0;
  size_t cap = 0;                       // floats per slot
  unsigned char* local = nullptr;       // [2][cap] floats | flags [kMaxPeers] | error word
  float* data[kMaxPeers] = {};          // mapped base pointers of every rank's allocation
  unsigned** d_peer_flags = nullptr;    // device array [n]: every rank's flag array (for peer_signal_kernel)
  unsigned seq = 0;
  unsigned* h_err = nullptr;            // pinned mirror of the error word, refreshed after every call
  cudaEvent_t err_ready = nullptr;      // (reserved)
  size_t bytes() const { return 2 * cap * sizeof(float) + (kMaxPeers + 1) * sizeof(unsigned); }
  unsigned* flags_of(int r) const { return reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(data[r]) + 2 * cap * sizeof(float)); }
};

// every rank has finished the kernels it enqueued before this point (a peer may still be reading my buffer)
static void peer_barrier(pb2_ctx* ctx) {
  NcclApi* api = nccl_api();
  float* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(float)) != cudaSuccess) return;
  cudaMemsetAsync(d, 0, sizeof(float), ctx->stream);
  api->AllReduce(d, d, 1, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
}

static void peer_destroy(pb2_ctx* ctx) {
  PeerGroup* g = static_cast<PeerGroup*>(ctx->peer);
  if (!g) return;
  if (g->n > 1 && g->data[g->n - 1]) peer_barrier(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (g->h_err) cudaFreeHost(g->h_err);
  if (g->err_ready) cudaEventDestroy(g->err_ready);
  for (int r = 0; r < g->n; ++r)
    if (r != g->rank && g->data[r]) cudaIpcCloseMemHandle(g->data[r]);
  cudaFree(g->local);
  cudaFree(g->d_peer_flags);
  delete g;
  ctx->peer = nullptr;
}

// Collective over the communicator: (re)build the peer group with room for `floats` per slot.
static int peer_ensure(pb2_ctx* ctx, size_t floats) {
  PeerGroup* g = static_cast<PeerGroup*>(ctx->peer);
  if (g && g->cap >= floats) return PB2_OK;
  if (ctx->comm_size > kMaxPeers)
    return set_error(ctx, PB2_ERR_UNSUPPORTED, "peer-memory reduction: at most 8 ranks (pb2_ctx_set_int rowshard_collective 0)");
  peer_destroy(ctx);
  NcclApi* api = nccl_api();
  g = new PeerGroup();
  g->n = ctx->comm_size;
  g->rank = ctx->comm_rank;
  g->cap = (floats + 1023) & ~(size_t)1023;
  int rc = check_cuda(ctx, cudaMalloc((void**)&g->local, g->bytes()), "cudaMalloc(peer buffer)");
  if (!rc) rc = check_cuda(ctx, cudaMemsetAsync(g->local, 0, g->bytes(), ctx->stream), "memset(peer buffer)");
  cudaIpcMemHandle_t mine;
  if (!rc) rc = check_cuda(ctx, cudaIpcGetMemHandle(&mine, g->local), "cudaIpcGetMemHandle");
  unsigned char* d_h = nullptr;
  std::vector<cudaIpcMemHandle_t> all(g->n);
  if (!rc) rc = check_cuda(ctx, cudaMalloc((void**)&d_h, sizeof(mine) * (g->n + 1)), "cudaMalloc(ipc handles)");
  if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream), "memcpy(ipc handle)");
  if (!rc)
    rc = check_nccl(ctx, api, api->AllGather(d_h, d_h + sizeof(mine), sizeof(mine), ncclUint8, (ncclComm_t)ctx->comm, ctx->stream),
                    "ncclAllGather(ipc handles)");
  if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(all.data(), d_h + sizeof(mine), sizeof(mine) * g->n, cudaMemcpyDeviceToHost, ctx->stream), "memcpy(ipc handles)");
  if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync(ipc handles)");
  cudaFree(d_h);
  for (int r = 0; r < g->n && !rc; ++r) {
    if (r == g->rank) { g->data[r] = reinterpret_cast<float*>(g->local); continue; }
    void* p = nullptr;
    rc = check_cuda(ctx, cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer access over NVLink)");
    g->data[r] = static_cast<float*>(p);
  }
  if (!rc) {
    unsigned* fl[kMaxPeers] = {};
    for (int r = 0; r < g->n; ++r) fl[r] = g->flags_of(r);
    rc = check_cuda(ctx, cudaMalloc((void**)&g->d_peer_flags, sizeof(fl)), "cudaMalloc(peer flag table)");
    if (!rc) rc = check_cuda(ctx, cudaMemcpy(g->d_peer_flags, fl, sizeof(fl), cudaMemcpyHostToDevice), "memcpy(peer flag table)");
  }
  if (!rc) rc = check_cuda(ctx, cudaMallocHost((void**)&g->h_err, sizeof(unsigned)), "cudaMallocHost(peer error word)");
  if (!rc) {
    *g->h_err = 0;
    rc = check_cuda(ctx, cudaEventCreateWithFlags(&g->err_ready, cudaEventDisableTiming), "cudaEventCreate");
  }
  ctx->peer = g;
  if (rc) { peer_destroy(ctx); return rc; }
  return PB2_OK;
}

// a wait that timed out in an earlier call (a peer that never published) is reported by the next call / by destroy
static int peer_check(pb2_ctx* ctx, PeerGroup* g) {
  if (g && g->h_err && *(volatile unsigned*)g->h_err)
    return set_error(ctx, PB2_ERR_CUDA, "pb2_rowshard_leapfrog: timed out waiting for a peer's gradient in an earlier call");
  return PB2_OK;
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
  int prc = PB2_OK;
  if (ctx->peer) {
    cudaStreamSynchronize(ctx->stream);
    prc = peer_check(ctx, static_cast<PeerGroup*>(ctx->peer));
  }
  peer_destroy(ctx);
  ncclResult_t r = api->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_size = 1;
  ctx->comm_rank = 0;
  if (prc) return prc;
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
  const size_t npk = (size_t)B * (D + 1);
  const bool peer = reduce_over_ranks && ctx->rowshard_collective == 1 && ctx->comm_size > 1;
  PeerGroup* g = nullptr;
  if (peer) {
    if (int rc = peer_ensure(ctx, npk)) return rc;
    g = static_cast<PeerGroup*>(ctx->peer);
    if (int rc = peer_check(ctx, g)) return rc;
  }
  const unsigned nb = (unsigned)((nBD + 255) / 256);
  rowshard_first_kernel<<<nb, 256, 0, ctx->stream>>>(B, D, d_step, step_kind, d_m, d_grad, d_x, v, d_x_out);
  ctx->launches += 1;
  for (int l = 0; l < num_steps; ++l) {
    PeerView pv{};
    float* dst = packed;
    if (peer) {   // this leapfrog's slot of my peer buffer: the gradient pass publishes straight into it
      g->seq += 1;
      pv.n = g->n;
      pv.seq = g->seq;
      pv.slot_off = (size_t)(g->seq & 1u) * g->cap;
      for (int r = 0; r < g->n; ++r) pv.data[r] = g->data[r];
      pv.flags = g->flags_of(g->rank);
      pv.err = g->flags_of(g->rank) + kMaxPeers;
      dst = g->data[g->rank] + pv.slot_off;
    }
    int rc = d_planes ? pb2_rowshard_logistic_grad_tc(ctx, d_planes, d_y, N, D, d_x_out, B, dst)
                      : pb2_rowshard_logistic_grad(ctx, d_X, d_y, N, D, DP, d_x_out, B, dst);
    if (rc) return rc;
    const int last = l + 1 == num_steps ? 1 : 0;
    if (peer) {
      peer_signal_kernel<<<1, 32, 0, ctx->stream>>>(g->d_peer_flags, g->n, g->rank, g->seq);
      rowshard_step_kernel<true><<<B, 128, 0, ctx->stream>>>(last, B, D, d_step, step_kind, nullptr, pv, v, d_x_out, d_grad_out,
                                                             d_logp_out, d_m_out);
      ctx->launches += 2;
    } else {
      if (reduce_over_ranks)   // psum over the data axis as a separate collective
        if (int rc2 = comm_allreduce_sum(ctx, packed, npk)) return rc2;
      rowshard_step_kernel<false><<<B, 128, 0, ctx->stream>>>(last, B, D, d_step, step_kind, packed, pv, v, d_x_out, d_grad_out,
                                                              d_logp_out, d_m_out);
      ctx->launches += 1;
    }
  }
  if (peer)   // refresh the host mirror of the error word (stream-ordered, no host synchronisation)
    if (int rc = check_cuda(ctx, cudaMemcpyAsync(g->h_err, g->flags_of(g->rank) + kMaxPeers, sizeof(unsigned),
                                                 cudaMemcpyDeviceToHost, ctx->stream), "memcpy(peer error word)"))
      return rc;
  return check_cuda(ctx, cudaGetLastError(), "pb2_rowshard_leapfrog kernels");
}

}  // extern "C"
