// extern "C" entry points of libpb2 (see include/pb2.h for the contract and the
// reference interfaces each entry point replaces).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pb2_internal.h"

static std::string g_last_error;

namespace pb2 {

int set_error(pb2_ctx* ctx, int code, const std::string& msg) {
  g_last_error = msg;
  if (ctx) ctx->err = msg;
  return code;
}

int check_cuda(pb2_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PB2_OK;
  return set_error(ctx, PB2_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static int ensure(pb2_ctx* ctx, void** ptr, size_t* cap, size_t need, const char* what) {
  if (need <= *cap) return PB2_OK;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *cap = 0;
  if (int rc = check_cuda(ctx, cudaMalloc(ptr, need), what)) return rc;
  *cap = need;
  return PB2_OK;
}

__global__ void da_advance_kernel(float* st, int k) {
  if (threadIdx.x == 0 && blockIdx.x == 0) st[3] = st[3] + (float)k;
}

static void host_split(const uint32_t key[2], int n, int layout, uint32_t* out) {
  Key k{key[0], key[1]};
  for (int j = 0; j < n; ++j) {
    Key c = split_at(k, (uint32_t)j, (uint32_t)n, layout);
    out[2 * j] = c.k0;
    out[2 * j + 1] = c.k1;
  }
}

}  // namespace pb2

using namespace pb2;

extern "C" {

int pb2_version(void) { return PB2_VERSION; }

const char* pb2_last_error(pb2_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int pb2_ctx_create(int device, pb2_ctx** out) {
  if (!out) return set_error(nullptr, PB2_ERR_INVALID, "pb2_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_error(nullptr, PB2_ERR_CUDA,
                     std::string("pb2_ctx_create: no CUDA device (") + cudaGetErrorString(e) + ")");
  if (device < 0 || device >= ndev) return set_error(nullptr, PB2_ERR_INVALID, "pb2_ctx_create: bad device index");
  if (int rc = check_cuda(nullptr, cudaSetDevice(device), "cudaSetDevice")) return rc;
  cudaDeviceProp prop;
  if (int rc = check_cuda(nullptr, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return rc;
  if (prop.major != 10)
    return set_error(nullptr, PB2_ERR_UNSUPPORTED, "libpb2 is built for sm_100a (Blackwell B200) only");
  pb2_ctx* c = new pb2_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (const char* v = getenv("PB2_DENSE_VARIANT")) c->dense_variant = atoi(v);
  if (const char* v = getenv("PB2_ROWSHARD_COLLECTIVE")) c->rowshard_collective = atoi(v) ? 1 : 0;   // A/B runs
  if (int rc = check_cuda(c, cudaMalloc(&c->d_queue, 64), "cudaMalloc(queue)")) { delete c; return rc; }
  if (int rc = check_cuda(c, cudaMalloc(&c->d_partial, 4096), "cudaMalloc(partial)")) { delete c; return rc; }
  *out = c;
  return PB2_OK;
}

int pb2_ctx_destroy(pb2_ctx* ctx) {
  if (!ctx) return PB2_OK;
  cudaSetDevice(ctx->device);
  pb2_comm_destroy(ctx);
  cudaFree(ctx->d_rs);
  cudaFree(ctx->d_queue);
  cudaFree(ctx->d_partial);
  cudaFree(ctx->d_ckpt);
  cudaFree(ctx->d_sched);
  cudaFree(ctx->d_step_keys);
  cudaFree(ctx->d_step_seq);
  if (ctx->h_keys) cudaFreeHost(ctx->h_keys);
  if (ctx->keys_copied) cudaEventDestroy(ctx->keys_copied);
  delete ctx;
  return PB2_OK;
}

int pb2_ctx_set_stream(pb2_ctx* ctx, void* s) {
  if (!ctx) return PB2_ERR_INVALID;
  ctx->stream = (cudaStream_t)s;
  return PB2_OK;
}

int pb2_ctx_synchronize(pb2_ctx* ctx) {
  if (!ctx) return PB2_ERR_INVALID;
  return check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
}

long long pb2_launch_count(pb2_ctx* ctx) { return ctx ? ctx->launches : 0; }

int pb2_ctx_set_int(pb2_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return PB2_ERR_INVALID;
  if (std::strcmp(name, "dense_variant") == 0) {
    ctx->dense_variant = value;
    return PB2_OK;
  }
  if (std::strcmp(name, "synchronous_run") == 0) {
    ctx->synchronous_run = value ? 1 : 0;
    return PB2_OK;
  }
  if (std::strcmp(name, "rowshard_collective") == 0) {
    if (value != 0 && value != 1) return set_error(ctx, PB2_ERR_INVALID, "rowshard_collective: 0 (NCCL) or 1 (peer memory)");
    ctx->rowshard_collective = value;
    return PB2_OK;
  }
  return set_error(ctx, PB2_ERR_INVALID, std::string("pb2_ctx_set_int: unknown option ") + name);
}

// ------------------------------------------------------------------ targets
int pb2_target_create(pb2_ctx* ctx, const pb2_target_desc* d, pb2_target** out) {
  if (!ctx || !d || !out) return set_error(ctx, PB2_ERR_INVALID, "pb2_target_create: NULL argument");
  cudaSetDevice(ctx->device);
  size_t na = 0, nb = 0;
  switch (d->kind) {
    case PB2_TARGET_EIGHT_SCHOOLS:
      if (d->n_rows < 1 || d->n_rows > 30 || d->dim != d->n_rows + 2)
        return set_error(ctx, PB2_ERR_INVALID, "eight schools: need 1 <= J <= 30 and dim == J + 2");
      na = nb = d->n_rows;
      break;
    case PB2_TARGET_DENSE_GAUSSIAN:
      if (d->dim < 1 || d->dim > 128) return set_error(ctx, PB2_ERR_UNSUPPORTED, "dense Gaussian: need 1 <= D <= 128");
      na = (size_t)d->dim * d->dim;
      nb = d->dim;
      break;
    case PB2_TARGET_LOGISTIC:
      if (d->dim < 1 || d->dim > 32 || d->n_rows < 1)
        return set_error(ctx, PB2_ERR_UNSUPPORTED, "logistic (shared-memory path): need 1 <= D <= 32, N >= 1");
      na = (size_t)d->n_rows * d->dim;
      nb = d->n_rows;
      break;
    case PB2_TARGET_STOCH_VOL:
    case PB2_TARGET_STOCH_VOL_CONSTRAINED:
    case PB2_TARGET_STOCH_VOL_CENTERED:
    case PB2_TARGET_STOCH_VOL_CENTERED_CONSTRAINED:
      if (d->n_rows < 1 || d->dim != d->n_rows + 3 || d->dim > 2560)
        return set_error(ctx, PB2_ERR_UNSUPPORTED, "stochastic volatility: need dim == T + 3 <= 2560");
      na = d->n_rows;
      nb = 0;
      break;
    default:
      return set_error(ctx, PB2_ERR_INVALID, "pb2_target_create: unknown kind");
  }
  if (!d->h_a) return set_error(ctx, PB2_ERR_INVALID, "pb2_target_create: h_a is NULL");
  pb2_target* t = new pb2_target();
  t->ctx = ctx;
  t->kind = d->kind;
  t->dim = d->dim;
  t->n_rows = d->n_rows;
  t->scalar = d->scalar;
  int rc = check_cuda(ctx, cudaMalloc(&t->d_a, na * sizeof(float)), "cudaMalloc(target a)");
  if (!rc) rc = check_cuda(ctx, cudaMemcpy(t->d_a, d->h_a, na * sizeof(float), cudaMemcpyHostToDevice), "memcpy(target a)");
  if (!rc && nb) {
    rc = check_cuda(ctx, cudaMalloc(&t->d_b, nb * sizeof(float)), "cudaMalloc(target b)");
    if (!rc) {
      if (d->h_b) rc = check_cuda(ctx, cudaMemcpy(t->d_b, d->h_b, nb * sizeof(float), cudaMemcpyHostToDevice), "memcpy(target b)");
      else if (d->kind == PB2_TARGET_DENSE_GAUSSIAN) rc = check_cuda(ctx, cudaMemset(t->d_b, 0, nb * sizeof(float)), "memset(loc)");
      else rc = set_error(ctx, PB2_ERR_INVALID, "pb2_target_create: h_b is NULL");
    }
  }
  if (rc) {
    cudaFree(t->d_a);
    cudaFree(t->d_b);
    delete t;
    return rc;
  }
  *out = t;
  return PB2_OK;
}

int pb2_user_target_check(const char* cuda_source, int dim, int flags, const char* include_dir) {
  // all three builds: as written, diagonally preconditioned, behind bijectors
  for (int variant = 0; variant < kUserVariants; ++variant) {
    std::vector<char> cubin;
    if (int rc = user_compile(nullptr, cuda_source, dim, include_dir, variant, flags, cubin)) return rc;
  }
  return PB2_OK;
}

int pb2_target_create_user(pb2_ctx* ctx, int dim, const char* cuda_source, int flags, const float* h_data,
                           long long n_data, const char* include_dir, pb2_target** out) {
  if (!ctx || !out || n_data < 0 || n_data > 0x7fffffffll || (n_data > 0 && !h_data))
    return set_error(ctx, PB2_ERR_INVALID, "pb2_target_create_user: bad argument");
  cudaSetDevice(ctx->device);
  if (!cuda_source || !include_dir) return set_error(ctx, PB2_ERR_INVALID, "pb2_target_create_user: NULL source");
  pb2_target* t = new pb2_target();
  t->ctx = ctx;
  t->kind = PB2_TARGET_USER;
  t->dim = dim;
  t->n_rows = (int)n_data;
  t->user_source = cuda_source;
  t->user_include = include_dir;
  t->user_flags = flags;
  int rc = user_target_ensure(ctx, t, 0);   // the preconditioned / transformed variants are built on first use
  if (!rc && n_data) {
    rc = check_cuda(ctx, cudaMalloc(&t->d_a, (size_t)n_data * sizeof(float)), "cudaMalloc(user data)");
    if (!rc) rc = check_cuda(ctx, cudaMemcpy(t->d_a, h_data, (size_t)n_data * sizeof(float), cudaMemcpyHostToDevice), "memcpy(user data)");
  }
  if (rc) {
    pb2_target_destroy(t);
    return rc;
  }
  *out = t;
  return PB2_OK;
}

int pb2_target_destroy(pb2_target* t) {
  if (!t) return PB2_OK;
  user_target_unload(t);
  cudaFree(t->d_a);
  cudaFree(t->d_b);
  cudaFree(t->d_tc);
  delete t;
  return PB2_OK;
}

int pb2_target_dim(const pb2_target* t) { return t ? t->dim : PB2_ERR_INVALID; }

// ------------------------------------------------------------------ RNG
int pb2_rng_split(const uint32_t key[2], int n, int layout, uint32_t* h_out) {
  if (!key || !h_out || n < 0) return set_error(nullptr, PB2_ERR_INVALID, "pb2_rng_split: bad argument");
  if (layout < PB2_LAYOUT_PARTITIONABLE || layout > PB2_LAYOUT_PHILOX)
    return set_error(nullptr, PB2_ERR_INVALID, "pb2_rng_split: unknown generator layout");
  host_split(key, n, layout, h_out);
  return PB2_OK;
}

int pb2_rng_fold_in(const uint32_t key[2], uint32_t data, uint32_t out[2]) {
  if (!key || !out) return set_error(nullptr, PB2_ERR_INVALID, "pb2_rng_fold_in: bad argument");
  Key k = fold_in(Key{key[0], key[1]}, data);
  out[0] = k.k0;
  out[1] = k.k1;
  return PB2_OK;
}

static int rng_fill(pb2_ctx* ctx, const uint32_t key[2], long long n, int layout, int what, float lo, float hi,
                    int ilo, int ihi, void* out) {
  if (!ctx || !key || (!out && n > 0) || n < 0) return set_error(ctx, PB2_ERR_INVALID, "pb2_rng_*: bad argument");
  if (layout < PB2_LAYOUT_PARTITIONABLE || layout > PB2_LAYOUT_PHILOX)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_rng_*: unknown generator layout");
  if (layout == PB2_LAYOUT_ORIGINAL && n >= (1ll << 32))
    return set_error(ctx, PB2_ERR_UNSUPPORTED, "original threefry layout: n must be < 2^32");
  Key k{key[0], key[1]}, kh{0, 0};
  if (what == 3) {  // randint: k1 (high draw), k2 (low draw) = split(key)
    if (ihi <= ilo) return set_error(ctx, PB2_ERR_INVALID, "pb2_rng_randint: need hi > lo");
    kh = split_at(k, 0, 2, layout);
    k = split_at(Key{key[0], key[1]}, 1, 2, layout);
  }
  cudaSetDevice(ctx->device);
  return launch_rng_fill(ctx, k, kh, n, layout, what, lo, hi, ilo, ihi, out);
}

int pb2_rng_bits(pb2_ctx* ctx, const uint32_t key[2], long long n, int layout, uint32_t* d_out) {
  return rng_fill(ctx, key, n, layout, 0, 0.f, 1.f, 0, 0, d_out);
}
int pb2_rng_uniform(pb2_ctx* ctx, const uint32_t key[2], long long n, float lo, float hi, int layout, float* d_out) {
  return rng_fill(ctx, key, n, layout, 1, lo, hi, 0, 0, d_out);
}
int pb2_rng_normal(pb2_ctx* ctx, const uint32_t key[2], long long n, int layout, float* d_out) {
  return rng_fill(ctx, key, n, layout, 2, 0.f, 1.f, 0, 0, d_out);
}
int pb2_rng_randint(pb2_ctx* ctx, const uint32_t key[2], long long n, int lo, int hi, int layout, int32_t* d_out) {
  return rng_fill(ctx, key, n, layout, 3, 0.f, 1.f, lo, hi, d_out);
}

// ------------------------------------------------------------------ primitives
static void base_params(ChainParams& p, const pb2_target* tgt, int B) {
  std::memset(&p, 0, sizeof(p));
  p.B = B;
  p.D = tgt->dim;
  p.B_global = B;
  p.n_parts = 1;
  p.part_off[0] = 0;
  p.part_off[1] = tgt->dim;
  p.unrolled = 1;
}

int pb2_logp_grad(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, float* d_logp, float* d_grad) {
  if (!ctx || !tgt || !d_x || !d_logp || !d_grad || B < 0)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_logp_grad: bad argument");
  if (B == 0) return PB2_OK;
  cudaSetDevice(ctx->device);
  ChainParams p;
  base_params(p, tgt, B);
  PrimIO io{};
  io.x_in = d_x;
  io.lp_out = d_logp;
  io.g_out = d_grad;
  return launch_chain(ctx, tgt, kModeLogpGrad, p, io);
}

int pb2_logp_grad_transformed(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, const int32_t* d_kind,
                              const float* d_lo, const float* d_hi, float* d_logp, float* d_grad) {
  if (!ctx || !tgt || !d_x || !d_logp || !d_grad || !d_kind || !d_lo || !d_hi || B < 0)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_logp_grad_transformed: bad argument");
  if (B == 0) return PB2_OK;
  cudaSetDevice(ctx->device);
  ChainParams p;
  base_params(p, tgt, B);
  p.bij_kind = d_kind;
  p.bij_lo = d_lo;
  p.bij_hi = d_hi;
  PrimIO io{};
  io.x_in = d_x;
  io.lp_out = d_logp;
  io.g_out = d_grad;
  return launch_chain(ctx, tgt, kModeLogpGrad, p, io);
}

int pb2_leapfrog(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_m, const float* d_x, const float* d_logp,
                 const float* d_grad, const float* d_step, int step_kind, int num_steps, float* d_m_out,
                 float* d_x_out, float* d_logp_out, float* d_grad_out) {
  if (!ctx || !tgt || !d_m || !d_x || !d_logp || !d_grad || !d_step || !d_m_out || !d_x_out || !d_logp_out ||
      !d_grad_out || B < 0 || num_steps < 0 || step_kind < 0 || step_kind > 2)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_leapfrog: bad argument");
  if (B == 0) return PB2_OK;
  cudaSetDevice(ctx->device);
  ChainParams p;
  base_params(p, tgt, B);
  p.step = d_step;
  p.step_kind = step_kind;
  PrimIO io{d_m, d_x, d_logp, d_grad, d_m_out, d_x_out, d_logp_out, d_grad_out, num_steps};
  return launch_chain(ctx, tgt, kModeLeapfrog, p, io);
}

// ------------------------------------------------------------------ transitions / sample_chain
int pb2_run(pb2_ctx* ctx, const pb2_target* tgt, const pb2_chain_layout* lay, const pb2_run_cfg* cfg,
            uint32_t h_seed[2], uint32_t* h_step_seeds, float* d_x, float* d_logp, float* d_grad,
            float* d_step_size, const pb2_da* da, const pb2_trace* trace, unsigned long long* d_leapfrog_total) {
  if (!ctx || !tgt || !lay || !cfg || !h_seed || !d_x || !d_logp || !d_grad || !d_step_size)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: NULL argument");
  if (lay->B < 1 || lay->B_global < lay->B || lay->chain_offset < 0 || lay->chain_offset + lay->B > lay->B_global)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: inconsistent chain layout");
  if (lay->rng_layout < PB2_LAYOUT_PARTITIONABLE || lay->rng_layout > PB2_LAYOUT_PHILOX)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: unknown generator layout");
  if (lay->n_parts < 1 || lay->n_parts > kMaxParts) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: need 1 <= n_parts <= 8");
  int sum = 0;
  for (int q = 0; q < lay->n_parts; ++q) {
    if (lay->part_sizes[q] < 1) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: empty state part");
    sum += lay->part_sizes[q];
  }
  if (sum != tgt->dim) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: state part sizes must sum to the target dimension");
  if (cfg->kind != PB2_KERNEL_HMC && cfg->kind != PB2_KERNEL_NUTS) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: bad kernel kind");
  if (cfg->num_results < 1 || cfg->num_burnin_steps < 0 || cfg->num_steps_between_results < 0)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: bad num_results / burnin / thinning");
  if (cfg->step_kind < 0 || cfg->step_kind > 2) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: bad step_kind");
  if (cfg->kind == PB2_KERNEL_HMC && cfg->num_leapfrog_steps < 1) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: num_leapfrog_steps < 1");
  if (cfg->kind == PB2_KERNEL_NUTS && (cfg->max_tree_depth < 1 || cfg->max_tree_depth > 12 || cfg->unrolled_leapfrog_steps < 1))
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: need 1 <= max_tree_depth <= 12 and unrolled_leapfrog_steps >= 1");
  const bool use_da = da && da->enabled;
  if (use_da && (cfg->step_kind != PB2_STEP_SCALAR || !da->d_state))
    return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_run: fused dual averaging needs a scalar step size");
  const bool da_ranks = use_da && da->reduce_over_ranks && lay->B_global > lay->B;
  if (da_ranks && (!ctx->comm || ctx->comm_size < 2))
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: da->reduce_over_ranks needs a communicator (pb2_comm_init)");
  if (da_ranks && ctx->comm_size > 500) return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_run: more than 500 ranks");
  cudaSetDevice(ctx->device);

  const long long n_steps_ll = (long long)cfg->num_burnin_steps + 1 +
                               (long long)(cfg->num_results - 1) * (1 + cfg->num_steps_between_results);
  if (n_steps_ll > (1ll << 30)) return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_run: too many transitions");
  const int n_steps = (int)n_steps_ll;

  // sample.py:344-349: step_seed, seed = split(seed) per transition
  std::vector<uint32_t> keys((size_t)2 * n_steps);
  if (cfg->explicit_step_seeds) {
    if (!h_step_seeds) return set_error(ctx, PB2_ERR_INVALID, "pb2_run: explicit_step_seeds needs h_step_seeds");
    std::memcpy(keys.data(), h_step_seeds, keys.size() * sizeof(uint32_t));
  } else {
    uint32_t cur[2] = {h_seed[0], h_seed[1]};
    uint32_t two[4];
    for (int t = 0; t < n_steps; ++t) {
      host_split(cur, 2, lay->rng_layout, two);
      keys[2 * t] = two[0];
      keys[2 * t + 1] = two[1];
      cur[0] = two[2];
      cur[1] = two[3];
    }
    h_seed[0] = cur[0];
    h_seed[1] = cur[1];
  }
  if (h_step_seeds && !cfg->explicit_step_seeds) std::memcpy(h_step_seeds, keys.data(), keys.size() * sizeof(uint32_t));
  if (lay->B == 0) return PB2_OK;

  ChainParams p;
  std::memset(&p, 0, sizeof(p));
  p.B = lay->B;
  p.D = tgt->dim;
  p.B_global = lay->B_global;
  p.chain_offset = lay->chain_offset;
  p.layout = lay->rng_layout;
  p.n_parts = lay->n_parts;
  p.part_off[0] = 0;
  for (int q = 0; q < lay->n_parts; ++q) p.part_off[q + 1] = p.part_off[q] + lay->part_sizes[q];
  p.x = d_x;
  p.lp = d_logp;
  p.g = d_grad;
  p.step = d_step_size;
  p.step_kind = cfg->step_kind;
  p.step_seq_stride = 0;
  p.burnin = cfg->num_burnin_steps;
  p.thin = cfg->num_steps_between_results;
  p.n_results = cfg->num_results;
  p.leapfrog_total = d_leapfrog_total;
  p.L = cfg->num_leapfrog_steps;
  p.max_depth = cfg->max_tree_depth;
  p.max_energy_diff = cfg->max_energy_diff;
  p.unrolled = cfg->unrolled_leapfrog_steps;
  p.scale = cfg->d_momentum_scale;
  p.bij_kind = cfg->d_bijector_kind;
  p.bij_lo = cfg->d_bijector_low;
  p.bij_hi = cfg->d_bijector_high;
  if (p.bij_kind && (!p.bij_lo || !p.bij_hi))
    return set_error(ctx, PB2_ERR_INVALID, "pb2_run: d_bijector_kind needs d_bijector_low and d_bijector_high");
  if (p.scale && cfg->step_kind == PB2_STEP_PER_DIM)
    return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_run: per-dimension step sizes together with a momentum scale");
  if (trace) {
    Trace& tr = p.tr;
    tr.states = trace->d_states;
    tr.target_log_prob = trace->d_target_log_prob;
    tr.grads = trace->d_grads_target_log_prob;
    tr.log_accept_ratio = trace->d_log_accept_ratio;
    tr.is_accepted = trace->d_is_accepted;
    tr.step_size = trace->d_step_size;
    tr.proposed_state = trace->d_proposed_state;
    tr.proposed_target_log_prob = trace->d_proposed_target_log_prob;
    tr.proposed_grads = trace->d_proposed_grads;
    tr.log_acceptance_correction = trace->d_log_acceptance_correction;
    tr.initial_momentum = trace->d_initial_momentum;
    tr.final_momentum = trace->d_final_momentum;
    tr.leapfrogs_taken = trace->d_leapfrogs_taken;
    tr.has_divergence = trace->d_has_divergence;
    tr.reach_max_depth = trace->d_reach_max_depth;
    tr.energy = trace->d_energy;
  }
  const bool nuts = cfg->kind == PB2_KERNEL_NUTS;
  const int stride = nuts ? nuts_sched_stride(lay->n_parts, cfg->max_tree_depth) : hmc_sched_stride(lay->n_parts);
  p.sched_stride = stride;
  const int mode = nuts ? kModeNUTS : kModeHMC;

  // dual-averaging bookkeeping: how many of the coming transitions still adapt
  int da_step = 0, da_nadapt = 0;
  if (use_da) {
    float hst[16];
    if (int rc = check_cuda(ctx, cudaMemcpyAsync(hst, da->d_state, sizeof(hst), cudaMemcpyDeviceToHost, ctx->stream), "memcpy(da state)")) return rc;
    if (int rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync(da state)")) return rc;
    da_step = (int)hst[3];
    da_nadapt = (int)hst[4];
    if (int rc = ensure(ctx, (void**)&ctx->d_step_seq, &ctx->step_seq_bytes, sizeof(float) * (size_t)lay->B, "cudaMalloc(lar scratch)")) return rc;
    p.lar_last = ctx->d_step_seq;
  }

  // diagonal preconditioning: the kernels run on u = x / s with grad_u = s grad_x (converted back below)
  if (p.scale) {
    if (int rc = launch_scale_rows(ctx, d_x, (size_t)lay->B, p.D, p.scale, 1)) return rc;
    if (int rc = launch_scale_rows(ctx, d_grad, (size_t)lay->B, p.D, p.scale, 0)) return rc;
  }
  const int chunk_max = nuts ? std::max(1, (int)((64ull << 20) / (4ull * stride))) : (1 << 20);
  PrimIO io{};
  int t = 0;
  while (t < n_steps) {
    const int t_end = std::min(n_steps, t + chunk_max);
    const int T = t_end - t;
    if (int rc = ensure(ctx, (void**)&ctx->d_step_keys, &ctx->step_keys_bytes, sizeof(uint32_t) * 2 * (size_t)T, "cudaMalloc(step keys)")) return rc;
    if (int rc = ensure(ctx, (void**)&ctx->d_sched, &ctx->sched_bytes, sizeof(uint32_t) * (size_t)stride * T, "cudaMalloc(key schedule)")) return rc;
    // the seeds go through pinned memory owned by the context, so that the copy may outlive this call
    const size_t kbytes = sizeof(uint32_t) * 2 * (size_t)T;
    if (ctx->keys_copied) cudaEventSynchronize(ctx->keys_copied);   // an earlier call's copy has left the staging buffer
    if (kbytes > ctx->h_keys_bytes) {
      if (ctx->h_keys) cudaFreeHost(ctx->h_keys);
      ctx->h_keys = nullptr;
      ctx->h_keys_bytes = 0;
      if (int rc = check_cuda(ctx, cudaMallocHost((void**)&ctx->h_keys, kbytes), "cudaMallocHost(step keys)")) return rc;
      ctx->h_keys_bytes = kbytes;
    }
    if (!ctx->keys_copied)
      if (int rc = check_cuda(ctx, cudaEventCreateWithFlags(&ctx->keys_copied, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    std::memcpy(ctx->h_keys, keys.data() + 2 * (size_t)t, kbytes);
    if (int rc = check_cuda(ctx, cudaMemcpyAsync(ctx->d_step_keys, ctx->h_keys, kbytes, cudaMemcpyHostToDevice, ctx->stream), "memcpy(step keys)")) return rc;
    if (int rc = check_cuda(ctx, cudaEventRecord(ctx->keys_copied, ctx->stream), "cudaEventRecord")) return rc;
    int rc = nuts ? launch_nuts_sched(ctx, ctx->d_step_keys, T, lay->n_parts, cfg->max_tree_depth, lay->rng_layout, ctx->d_sched)
                  : launch_hmc_sched(ctx, ctx->d_step_keys, T, lay->n_parts, lay->rng_layout, ctx->d_sched);
    if (rc) return rc;
    p.sched = ctx->d_sched;
    p.t_sched0 = t;
    int u = t;
    while (u < t_end) {
      const bool adapting = use_da && da_step < da_nadapt;
      const int u_end = adapting ? u + 1 : t_end;
      p.t0 = u;
      p.t1 = u_end;
      if (int rc2 = launch_chain(ctx, tgt, mode, p, io)) return rc2;
      if (adapting) {
        if (int rc2 = launch_da_partial(ctx, p.lar_last, lay->B, ctx->d_partial)) return rc2;
        if (da_ranks) {
          // every rank gathers all fixed-point partial sums and adds them: exact, the same bits everywhere
          if (int rc2 = comm_allgather(ctx, ctx->d_partial, ctx->d_partial + 2, 2)) return rc2;
          if (int rc2 = launch_da_apply(ctx, ctx->d_partial + 2, ctx->comm_size, lay->B_global, da->d_state, d_step_size, nullptr)) return rc2;
        } else if (int rc2 = launch_da_apply(ctx, ctx->d_partial, 1, lay->B, da->d_state, d_step_size, nullptr)) return rc2;
        da_step += 1;
      } else if (use_da) {
        da_advance_kernel<<<1, 32, 0, ctx->stream>>>(da->d_state, u_end - u);
        ctx->launches += 1;
        da_step += u_end - u;
      }
      u = u_end;
    }
    // the host key buffer / schedule scratch are reused by the next chunk
    if (t_end < n_steps)
      if (int rc2 = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync(chunk)")) return rc2;
    t = t_end;
  }
  if (p.scale) {   // back to the original coordinates: x = s u, grad_x = grad_u / s, momentum m_x = m_u / s
    const size_t R = (size_t)cfg->num_results * (size_t)lay->B;
    int rc = launch_scale_rows(ctx, d_x, (size_t)lay->B, p.D, p.scale, 0);
    if (!rc) rc = launch_scale_rows(ctx, d_grad, (size_t)lay->B, p.D, p.scale, 1);
    if (!rc) rc = launch_scale_rows(ctx, p.tr.states, R, p.D, p.scale, 0);
    if (!rc) rc = launch_scale_rows(ctx, p.tr.grads, R, p.D, p.scale, 1);
    if (!rc) rc = launch_scale_rows(ctx, p.tr.proposed_state, R, p.D, p.scale, 0);
    if (!rc) rc = launch_scale_rows(ctx, p.tr.proposed_grads, R, p.D, p.scale, 1);
    if (!rc) rc = launch_scale_rows(ctx, p.tr.initial_momentum, R, p.D, p.scale, 1);
    if (!rc) rc = launch_scale_rows(ctx, p.tr.final_momentum, R, p.D, p.scale, 1);
    if (rc) return rc;
  }
  // everything above is enqueued on the context's stream; the caller's later work on that stream is ordered after it
  if (ctx->synchronous_run) return check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync(run)");
  return check_cuda(ctx, cudaGetLastError(), "pb2_run");
}

// ------------------------------------------------------------------ dual averaging
int pb2_da_init(pb2_ctx* ctx, float step_size, int num_adaptation_steps, float target_accept_prob,
                float exploration_shrinkage, float step_count_smoothing, float decay_rate, float log_shrinkage_target,
                int step, float error_sum, float log_averaging_step, float* d_state) {
  if (!ctx || !d_state) return set_error(ctx, PB2_ERR_INVALID, "pb2_da_init: NULL argument");
  if (!(step_size > 0.f)) return set_error(ctx, PB2_ERR_INVALID, "pb2_da_init: step_size must be positive");
  float h[16] = {0};
  h[0] = error_sum;
  h[1] = log_averaging_step;
  // dual_averaging_step_size_adaptation.py:589-591: log(10) + log(step_size)
  h[2] = std::isnan(log_shrinkage_target) ? (2.302585092994046f + logf(step_size)) : log_shrinkage_target;
  h[3] = (float)step;
  h[4] = (float)num_adaptation_steps;
  h[5] = target_accept_prob;
  h[6] = exploration_shrinkage;
  h[7] = step_count_smoothing;
  h[8] = decay_rate;
  h[9] = step_size;
  cudaSetDevice(ctx->device);
  if (int rc = check_cuda(ctx, cudaMemcpyAsync(d_state, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream), "memcpy(da init)")) return rc;
  return check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync(da init)");
}

int pb2_da_partial(pb2_ctx* ctx, const float* d_lar, int B, float* d_partial) {
  if (!ctx || !d_lar || !d_partial || B < 1) return set_error(ctx, PB2_ERR_INVALID, "pb2_da_partial: bad argument");
  cudaSetDevice(ctx->device);
  return launch_da_partial(ctx, d_lar, B, d_partial);
}

int pb2_da_apply(pb2_ctx* ctx, const float* d_partials, int n, long long B_global, float* d_state, float* d_step_out) {
  if (!ctx || !d_partials || !d_state || n < 1 || B_global < 1) return set_error(ctx, PB2_ERR_INVALID, "pb2_da_apply: bad argument");
  cudaSetDevice(ctx->device);
  return launch_da_apply(ctx, d_partials, n, B_global, d_state, d_step_out, nullptr);
}

// ------------------------------------------------------------------ streaming moments
int pb2_running_moments_update(pb2_ctx* ctx, const float* d_x, long long rows, int D, float* d_state) {
  if (!ctx || !d_x || !d_state || rows < 0 || D < 1)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_running_moments_update: bad argument");
  cudaSetDevice(ctx->device);
  return launch_running_moments(ctx, d_x, rows, D, d_state);
}

// ------------------------------------------------------------------ diagnostics
int pb2_ess(pb2_ctx* ctx, const float* d_states, int N, int B, int D, float filter_threshold, int filter_beyond_lag,
            int filter_beyond_positive_pairs, int cross_chain, float* d_out) {
  if (!ctx || !d_states || !d_out || N < 2 || B < 1 || D < 1) return set_error(ctx, PB2_ERR_INVALID, "pb2_ess: bad argument");
  if (cross_chain && B < 2)
    return set_error(ctx, PB2_ERR_INVALID, "When `cross_chain_dims` is not `None`, there must be > 1 chain in `states`.");
  cudaSetDevice(ctx->device);
  int max_lag = N - 1;
  if (filter_beyond_lag >= 0) max_lag = std::min(N - 1, filter_beyond_lag);
  const int use_thr = std::isnan(filter_threshold) ? 0 : 1;
  float* mean = nullptr;
  if (int rc = check_cuda(ctx, cudaMallocAsync(&mean, sizeof(float) * (size_t)B * D, ctx->stream), "cudaMallocAsync(ess mean)")) return rc;
  int rc = launch_ess(ctx, d_states, N, B, D, filter_threshold, use_thr, max_lag, filter_beyond_positive_pairs, cross_chain, mean, d_out);
  cudaFreeAsync(mean, ctx->stream);
  return rc;
}

int pb2_rhat(pb2_ctx* ctx, const float* d_states, int N, int B, int D, int split_chains, float* d_out) {
  if (!ctx || !d_states || !d_out || B < 1 || D < 1) return set_error(ctx, PB2_ERR_INVALID, "pb2_rhat: bad argument");
  if (split_chains && N < 4) return set_error(ctx, PB2_ERR_INVALID, "Must provide at least 4 samples when splitting chains.");
  if (!split_chains && N < 2) return set_error(ctx, PB2_ERR_INVALID, "Must provide at least 2 samples.");
  cudaSetDevice(ctx->device);
  return launch_rhat(ctx, d_states, N, B, D, split_chains, d_out);
}

}  // extern "C"
