"""CPU: weight hand-off round trip (params.py save_npz / load_npz, the format scripts/export_reference_params.py
writes on the reference side; checkpointer.py:15-34) - names, shapes and values survive, and the oracle computes the
same outputs from the re-loaded parameters."""

import numpy as np
import pytest
import torch

from jaxrl_b200.config import ModelConfig
from jaxrl_b200.params import ParamStore, param_specs
from oracle import oracle as O

CFG = ModelConfig(num_layers=2, max_seq_length=16, action_dim=5, task_count=3)


def test_npz_round_trip_and_oracle_forward(tmp_path):
    a = ParamStore(CFG, "cpu", seed=1, bias_noise=0.1)
    f = str(tmp_path / "params.npz")
    a.save_npz(f, action_dim=CFG.action_dim, obs_max_value=CFG.obs_max_value)
    with np.load(f) as z:
        names = sorted(k[6:] for k in z.files if k.startswith("param/"))
        assert names == sorted(n for n, _, _ in param_specs(CFG))
        assert z["param/layers.0.history.query_proj.kernel"].shape == (128, 4, 32)
        assert z["param/layers.1.history.out.kernel"].shape == (4, 32, 128)
        assert z["param/obs_encoder.layers.0.kernel"].shape == (3, 3, 10, 16)
    b = ParamStore(CFG, "cpu", seed=2)
    assert not torch.equal(a.flat, b.flat)
    meta = b.load_npz(f)
    assert int(meta["action_dim"]) == 5 and tuple(meta["obs_max_value"]) == (5, 5)
    for n in a.views:
        assert torch.equal(a.views[n], b.views[n]), n
    gen = torch.Generator().manual_seed(0)
    B, T = 2, 8
    obs = torch.stack([torch.randint(0, n, (B, T, 11, 11), generator=gen) for n in CFG.obs_max_value], -1).to(torch.uint8)
    ts = O.TimeStep(obs=obs, time=torch.arange(T, dtype=torch.int32)[None], terminated=torch.zeros(B, T, dtype=torch.bool),
                    last_action=torch.randint(0, 5, (B, T), generator=gen).to(torch.int32),
                    reward=torch.randn(B, T, generator=gen), action_mask=None,
                    task_ids=torch.randint(0, 3, (B, T), generator=gen).to(torch.int32))
    va, la, _ = O.model_forward(a.oracle_dict(), CFG, ts)
    vb, lb, _ = O.model_forward(b.oracle_dict(), CFG, ts)
    assert torch.equal(va, vb) and torch.equal(la, lb)


def test_load_rejects_wrong_names_and_shapes(tmp_path):
    a = ParamStore(CFG, "cpu", seed=1)
    d = a.oracle_dict()
    bad = dict(d)
    bad.pop("output_norm.scale")
    with pytest.raises(KeyError):
        ParamStore(CFG, "cpu").load_dict(bad)
    bad = dict(d)
    bad["layers.0.ffn.up_proj.kernel"] = torch.zeros(768, 128)
    with pytest.raises(ValueError):
        ParamStore(CFG, "cpu").load_dict(bad)
    ParamStore(CFG, "cpu").load_dict({"output_norm.scale": torch.ones(128)}, strict=False)


def test_export_script_flattens_nnx_paths():
    """The exporter's path flattening, exercised on a stand-in for nnx.State.flat_state()."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "export_reference_params", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                                                "export_reference_params.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class Leaf:
        def __init__(self, v):
            self.value = v

    class State:
        def flat_state(self):
            return [(("layers", 0, "history", "query_proj", "kernel"), Leaf(np.ones((128, 4, 32), np.float32))),
                    (("output_norm", "scale", "value"), np.ones(128, np.float32))]
    flat = mod.flatten_state(State())
    assert set(flat) == {"layers.0.history.query_proj.kernel", "output_norm.scale"}
    assert flat["layers.0.history.query_proj.kernel"].shape == (128, 4, 32)
