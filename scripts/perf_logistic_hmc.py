"""HMC on the logistic-regression target (1000 x 25, C3's model): FP32 warp-per-chain fused kernel vs lock-step leapfrogs
on tcgen05 (LogisticRegression(tensor_core_transitions=True))."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
dev = torch.device('cuda', 0)
X, y = tfp.targets.synthetic_logistic_data(1000, 24, seed=1)
L = 10
for B in (8192, 65536):
  x0 = torch.tensor((0.1 * np.random.default_rng(0).standard_normal((B, 25))).astype(np.float32), device=dev)
  for name, tg in (('fp32 warp-per-chain (fused sample_chain)', tfp.targets.LogisticRegression(X, y)),
                   ('tcgen05 lock-step leapfrogs', tfp.targets.LogisticRegression(X, y, tensor_core_transitions=True))):
    k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.02, num_leapfrog_steps=L)
    tfp.mcmc.sample_chain(2, x0, kernel=k, seed=1, trace_fn=None)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      r = tfp.mcmc.sample_chain(10, x0, kernel=k, seed=2, trace_fn=lambda _, kr: kr.is_accepted)
      e1.record(); torch.cuda.synchronize()
      best = min(best, e0.elapsed_time(e1))
    print('B=%6d %-42s %.2f ms per transition -> %.3e grad-evals/s (accept %.2f)' % (
        B, name, best / 10, B * L * 10 / best * 1e3, r.trace.float().mean().item()), flush=True)
