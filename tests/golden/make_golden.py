"""Generates the golden fixtures in this directory FROM THE REFERENCE (run in the build container only;
/root/reference does not exist on the GPU box).  The reference's ground-truth modules and the embedded
S&P 500 series are plain NumPy/Python files, importable without tensorflow/jax:

  eight_schools_ground_truth.json   <- spinoffs/inference_gym/inference_gym/targets/ground_truth/eight_schools.py
  sv_sp500.npz                      <- .../internal/datasets/sp500_closing_prices.py (centred returns as
                                       internal/data.py:575-579 computes them) and
                                       .../targets/ground_truth/stochastic_volatility_sp500.py
  nuts_tables_depth4.json           <- the literal pins of tensorflow_probability/python/mcmc/nuts_test.py:191-215
  dual_averaging_pins.json          <- tensorflow_probability/python/mcmc/dual_averaging_step_size_adaptation_test.py:51-57
  jax_notebook_rng.json             <- the executed cells of tensorflow_probability/examples/jupyter_notebooks/
                                       TensorFlow_Probability_on_JAX.ipynb: outputs of a LIVE JAX run that the
                                       reference keeps (PRNGKey(0), random.split, random.normal on the key and on
                                       both split keys, tfd.Normal(0, 1).sample(seed=PRNGKey(0)), and a jitted
                                       split -> normal -> exp chain) -- the reference-held pin of the threefry stream,
                                       the key split and the uniform -> normal transform of the seed contract (R1)
  tf_notebook_hmc.json              <- the executed HMC cell of .../jupyter_notebooks/TFP_Release_Notebook_0_11_0.ipynb:
                                       all_states / kernel results of a sample_chain run whose dynamics do not depend
                                       on the generator (pins leapfrog + accept + burn-in indexing on a reference run)
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
  # ---- JAX outputs held by the reference's own notebook (cell source -> printed output)
  nb_path = 'tensorflow_probability/examples/jupyter_notebooks/TensorFlow_Probability_on_JAX.ipynb'
  nb = json.load(open(os.path.join(REF, nb_path)))

  def output_of(fragment):
    for i, c in enumerate(nb['cells']):
      if c['cell_type'] == 'code' and fragment in ''.join(c['source']):
        outs = []
        for o in c.get('outputs', []):
          if 'text' in o:
            outs.append(''.join(o['text']))
          elif 'text/plain' in o.get('data', {}):
            outs.append(''.join(o['data']['text/plain']))
        return i, ' '.join(outs)
    raise KeyError(fragment)

  num = r'-?\d+\.?\d*(?:e[+-]?\d+)?'
  fl = lambda txt: [float(v) for v in re.findall(num, txt)]
  rng = {'source': nb_path, 'cells': {}}
  for name, frag, parse in [
      ('prng_key_0', 'key = random.PRNGKey(0)  # Creates a key', lambda t: [int(v) for v in re.findall(r'\d+', t)]),
      ('normal_key_0', 'print(random.normal(key))', lambda t: fl(t)[0]),
      ('split_key_0', 'key1, key2 = random.split(key, num=2)',
       lambda t: np.asarray([int(v) for v in re.findall(r'\d+', t)]).reshape(2, 2).tolist()),
      ('normal_split_keys', 'print(random.normal(key1), random.normal(key2))', lambda t: fl(t)[:2]),
      ('tfd_normal_sample_key_0', 'tfd.Normal(0., 1.).sample(seed=random.PRNGKey(0))',
       lambda t: fl(t.replace('float32', ''))[0]),
      ('split_normal_exp_mean_variance', 'def random_distribution(key):', lambda t: fl(t)[:2])]:
    cell, txt = output_of(frag)
    rng[name] = parse(txt)
    rng['cells'][name] = cell
  # two more executed notebooks: multi-element draws (the counter layout of a stream) and a 2-d shape
  def notebook_output(nb_rel, fragment):
    nb2 = json.load(open(os.path.join(REF, nb_rel)))
    for i, c in enumerate(nb2['cells']):
      if c['cell_type'] == 'code' and fragment in ''.join(c['source']):
        outs = []
        for o in c.get('outputs', []):
          if 'text' in o:
            outs.append(''.join(o['text']))
          elif 'text/plain' in o.get('data', {}):
            outs.append(''.join(o['data']['text/plain']))
        return i, ' '.join(outs)
    raise KeyError(fragment)

  more = {}
  nb_a = 'discussion/examples/TFP_and_Jax.ipynb'
  cell, txt = notebook_output(nb_a, 'tf.random.stateless_uniform([1, 2], seed=random.PRNGKey(0))')
  more['uniform_1x2_key_0'] = {'source': nb_a, 'cell': cell, 'value': fl(txt.replace('float32', ''))[:2]}
  cell, txt = notebook_output(nb_a, 'tfd.MultivariateNormalDiag(tf.zeros(5), tf.ones(5)),\n    tfb.Exp())')
  more['exp_normal_2x5_key_0'] = {'source': nb_a, 'cell': cell,
                                  'value': np.asarray(fl(txt)[:10]).reshape(2, 5).tolist()}
  nb_b = 'tensorflow_probability/examples/jupyter_notebooks/Distributed_Inference_with_JAX.ipynb'
  cell, txt = notebook_output(nb_b, 'dist = tfd.Sample(tfd.Normal(0., 1.), jax.device_count())')
  more['normal_8_key_0_tpu'] = {'source': nb_b, 'cell': cell, 'value': fl(txt.replace('float32', ''))[:8],
                                'note': 'run on 8 TPU cores: the erfinv of that platform differs from the CPU one in '
                                        'the last digits (the same three notebooks print normal(PRNGKey(0)) as '
                                        '-0.20584226, -0.20584235 and -0.20584236)'}
  rng['more'] = more
  # ---- a reference-held HMC run (TF substrate, eager): sample_chain(num_results=5, num_burnin_steps=100) of
  # HamiltonianMonteCarlo(lambda x: -(x - .2)**2, step_size=1., num_leapfrog_steps=2) from zeros([3]).  For this
  # target two unit leapfrog steps map x -> 0.4 - x and m -> -m whatever the momentum, so the trajectory does not
  # depend on the generator: it pins the leapfrog arithmetic, the accept step and sample_chain's burn-in indexing.
  nb_c = 'tensorflow_probability/examples/jupyter_notebooks/TFP_Release_Notebook_0_11_0.ipynb'
  cell, txt = notebook_output(nb_c, 'kernel = tfp.mcmc.HamiltonianMonteCarlo(lambda x: -(x - .2)**2')
  first = txt[:txt.index('And again')]
  blocks = re.findall(r'array\(\[\[(.*?)\]\]', first, re.S)
  rows = lambda blk: [[float(v) for v in re.findall(num, r)] for r in blk.split('],')]
  hmc = {'source': nb_c, 'cell': cell, 'step_size': 1.0, 'num_leapfrog_steps': 2, 'num_results': 5,
         'num_burnin_steps': 100, 'all_states': rows(blocks[0]), 'log_acceptance_correction': rows(blocks[1])}
  tl = re.search(r'target_log_prob=.*?array\(\[\[(.*?)\]', first, re.S).group(1)
  hmc['target_log_prob_first_result'] = [float(v) for v in re.findall(num, tl)]
  json.dump(hmc, open(os.path.join(HERE, 'tf_notebook_hmc.json'), 'w'), indent=1)
  json.dump(rng, open(os.path.join(HERE, 'jax_notebook_rng.json'), 'w'), indent=1)
  print('golden fixtures written to', HERE)


if __name__ == '__main__':
  main()
