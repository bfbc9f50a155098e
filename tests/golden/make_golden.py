"""Generates the golden fixtures in this directory FROM THE REFERENCE (run in the build container only;
/root/reference does not exist on the GPU box).  The reference's ground-truth modules and the embedded
S&P 500 series are plain NumPy/Python files, importable without tensorflow/jax:

  eight_schools_ground_truth.json   <- spinoffs/inference_gym/inference_gym/targets/ground_truth/eight_schools.py
  sv_sp500.npz                      <- .../internal/datasets/sp500_closing_prices.py (centred returns as
                                       internal/data.py:575-579 computes them) and
                                       .../targets/ground_truth/stochastic_volatility_sp500.py
  nuts_tables_depth4.json           <- the literal pins of tensorflow_probability/python/mcmc/nuts_test.py:191-215
  dual_averaging_pins.json          <- tensorflow_probability/python/mcmc/dual_averaging_step_size_adaptation_test.py:51-57
"""
import importlib.util
import json
import os
import re

import numpy as np

REF = '/root/reference'
GYM = os.path.join(REF, 'spinoffs/inference_gym/inference_gym')
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
  spec = importlib.util.spec_from_file_location(name, path)
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m


def main():
  gt = load(os.path.join(GYM, 'targets/ground_truth/eight_schools.py'), 'gt_es')
  out = {k: np.asarray(getattr(gt, k)).tolist() for k in dir(gt) if k.startswith('IDENTITY_')}
  json.dump(out, open(os.path.join(HERE, 'eight_schools_ground_truth.json'), 'w'), indent=1)

  prices = np.asarray(load(os.path.join(GYM, 'internal/datasets/sp500_closing_prices.py'), 'sp').CLOSING_PRICES)
  returns = np.diff(prices)
  centered = returns - np.mean(returns)
  gts = load(os.path.join(GYM, 'targets/ground_truth/stochastic_volatility_sp500.py'), 'gt_sv')
  arrays = {k.lower(): np.asarray(getattr(gts, k)) for k in dir(gts) if k.startswith('IDENTITY_')}
  np.savez_compressed(os.path.join(HERE, 'sv_sp500.npz'), centered_returns=centered, **arrays)

  src = open(os.path.join(REF, 'tensorflow_probability/python/mcmc/nuts_test.py')).read()
  blk = src[src.index('def testCorrectReadWriteInstruction'):]
  blk = blk[:blk.index('\n  def ', 10)]
  nums = [int(v) for v in re.findall(r'(?<![\w.])\d+(?![\w.])', blk[blk.index('_write_instruction'):])]
  write, read = nums[:16], np.asarray(nums[16:16 + 32]).reshape(16, 2).tolist()
  json.dump({'max_tree_depth': 4, 'write_instruction': write, 'read_instruction': read},
            open(os.path.join(HERE, 'nuts_tables_depth4.json'), 'w'))

  src = open(os.path.join(REF, 'tensorflow_probability/python/mcmc/dual_averaging_step_size_adaptation_test.py')).read()
  pins = {m.group(1): float(m.group(2)) for m in re.finditer(r'^(_UPDATE_\w+) = ([0-9.]+)', src, re.M)}
  pins['_INITIAL_T'] = float(re.search(r'^_INITIAL_T = ([0-9.]+)', src, re.M).group(1))
  pins['_EXPLORATION_SHRINKAGE'] = float(re.search(r'^_EXPLORATION_SHRINKAGE = ([0-9.]+)', src, re.M).group(1))
  json.dump(pins, open(os.path.join(HERE, 'dual_averaging_pins.json'), 'w'), indent=1)
  print('golden fixtures written to', HERE)


if __name__ == '__main__':
  main()
