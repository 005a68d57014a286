"""CPU: the self-contained CPU-arm workload (oracle/synth.py) agrees with the product's config / parameter specs, runs,
and the `--impl reference` arm of bench.py never loads the CUDA library."""

import dataclasses
import json
import os
import subprocess
import sys

import torch

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_literals_match_named_configs():
    from jaxrl_b200.config import NAMED
    assert set(synth.WORKLOADS) == set(NAMED)
    for name, (cfg, hyp, agents) in synth.WORKLOADS.items():
        run = NAMED[name]
        for f in dataclasses.fields(run.model):
            assert getattr(cfg, f.name) == getattr(run.model, f.name), (name, f.name)
        for f in dataclasses.fields(run.ppo):
            assert getattr(hyp, f.name) == getattr(run.ppo, f.name), (name, f.name)
        assert agents == run.num_agents


def test_param_specs_and_init_match_param_store():
    from jaxrl_b200.config import NAMED
    from jaxrl_b200.params import param_specs, _init
    for name, (cfg, _, _) in synth.WORKLOADS.items():
        assert synth.param_specs(cfg) == param_specs(NAMED[name].model), name
    cfg = synth.WORKLOADS["test"][0]
    g1, g2 = torch.Generator().manual_seed(4), torch.Generator().manual_seed(4)
    for _, shape, kind in synth.param_specs(cfg):
        assert torch.equal(synth.init_tensor(shape, kind, g1), _init(shape, kind, g2))


def test_cpu_update_runs_on_a_tiny_sample():
    cfg, hyp, _ = synth.WORKLOADS["test"]
    cfg = type(cfg)(**{**vars(cfg), "max_seq_length": 8})
    p = synth.init_params(cfg, 0)
    obs, reward, gumbel, tasks = synth.synthetic_inputs(cfg, 4, 0)
    grads, m = synth.ppo_update(p, cfg, hyp, obs, reward, gumbel, tasks)
    assert m == 2 and set(grads) == set(p)
    assert all(torch.isfinite(g).all() for g in grads.values())
    assert grads["layers.1.ffn.down_proj.kernel"].abs().sum() > 0


def test_package_import_is_lazy_and_reference_arm_never_loads_the_library():
    code = ("import sys, jaxrl_b200, jaxrl_b200.config, jaxrl_b200.params, jaxrl_b200.dist; "
            "assert 'jaxrl_b200._lib' not in sys.modules; print('lazy ok')")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "lazy ok" in out.stdout, out.stderr[-2000:]
    # bench.py --impl reference: prints one JSON line, no libmapox_b200.so mapped, no jaxrl_b200 module imported
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--config', 'test', '--cpu-agents', '2', "
            "'--steps', '1', '--warmup', '0']; runpy.run_path('bench.py', run_name='__main__'); "
            "maps = open('/proc/self/maps').read(); assert 'libmapox_b200' not in maps; "
            "assert not [m for m in sys.modules if m.startswith('jaxrl_b200')], 'CPU arm imported the product package'")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == "ppo_env_steps_per_sec"
