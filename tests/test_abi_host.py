"""CPU suite: the C-ABI library loads and exports every symbol include/qpnet_b200.h declares;
host-side logic of the Python mirror (no compute calls -- there is no GPU here)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "qpnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from qpnet_b200 import _lib
    declared = _header_symbols()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (qp_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    # the ctypes table binds exactly the declared symbols
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.lib.qp_abi_version() == 3


def test_no_cpu_fallback_without_gpu():
    from qpnet_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib.qp_device_ok() == _lib.QP_EARCH
    from qpnet_b200.qpnet import QPNet
    m = QPNet(n_resch=32, n_skipch=16)
    x = torch.zeros(1, 2000, dtype=torch.long)
    with pytest.raises(RuntimeError):
        m(x, torch.zeros(1, 39, 19), torch.ones(1, 2000), torch.tensor([100]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "qpnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_state_dict_matches_reference_layout():
    from oracle import qpnet_oracle as orc
    from qpnet_b200.qpnet import QPNet, initialize
    for kw in (dict(), dict(n_resch=32, n_skipch=16), dict(n_resch=64, n_skipch=32, dilationF_depth=3,
                                                           dilationF_repeat=2, dilationA_depth=2, dilationA_repeat=2)):
        m = QPNet(**kw)
        spec = orc.state_dict_spec(orc.Arch(**kw))
        sd = m.state_dict()
        assert [k for k in sd] == [n for n, _ in spec]
        assert all(tuple(sd[n].shape) == s for n, s in spec)
        assert [n for n, _ in m.named_parameters()] == [n for n, _ in spec]
    m = QPNet()
    assert sum(p.numel() for p in m.parameters()) == 24151151
    assert (m.receptiveCausal_field, m.receptiveF_field, m.receptiveA_field) == (1, 45, 15)
    m = QPNet(n_resch=32, n_skipch=16)
    m.apply(initialize)
    assert float(m.upsampling.conv.weight.min()) == 1.0 and float(m.causal.conv.bias.abs().max()) == 0.0
    p = orc.init_params(orc.Arch(n_resch=32, n_skipch=16), 1, 0.1)
    m.load_state_dict(p)            # oracle tensors load by name


def test_constructor_rejects_unsupported():
    from qpnet_b200.qpnet import QPNet
    with pytest.raises(ValueError):
        QPNet(kernel_size=3)


def test_synth_is_deterministic_and_in_range():
    from qpnet_b200 import synth
    a, f0a, n = synth.utterance(20, 3, 0.5)
    b, f0b, _ = synth.utterance(20, 3, 0.5)
    assert np.array_equal(a, b) and np.array_equal(f0a, f0b) and n == 20 * 110 - 1
    f0 = synth.f0_contour(500, 7)
    assert f0.min() >= 45.0 and f0.max() <= 450.0


def test_reference_checkpoint_round_trip(tmp_path):
    """a checkpoint and model.conf written by the reference's trainer (fixtures produced by the
    unmodified reference class + torch.optim.Adam, make_golden.py: golden_checkpoint) load into the new class, resume the
    optimizer, and save back in the same format."""
    from qpnet_b200 import checkpoint as ck
    from qpnet_b200.qpnet import QPNet
    gold = os.path.join(ROOT, "tests", "golden")
    kw = ck.load_config(os.path.join(gold, "model.conf"))
    assert kw["n_resch"] == 8 and kw["upsampling_factor"] == 110
    m = QPNet(**kw)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    it = ck.load_checkpoint(os.path.join(gold, "checkpoint-7.pkl"), m, opt)
    assert it == 7
    ref = torch.load(os.path.join(gold, "checkpoint-7.pkl"), weights_only=False)
    sd = m.state_dict()
    assert list(sd) == list(ref["model"])
    assert all(torch.equal(sd[k], ref["model"][k]) for k in sd)
    assert opt.state_dict()["state"][0]["step"] == ref["optimizer"]["state"][0]["step"]
    path = ck.save_checkpoint(str(tmp_path), m, opt, it + 1)
    assert os.path.basename(path) == "checkpoint-8.pkl"
    again = torch.load(path, weights_only=False)
    assert set(again) == {"model", "optimizer", "iterations"} and again["iterations"] == 8
    # a DataParallel checkpoint ("module." prefix, qpnet_train.py:416-423) loads too
    torch.save({"model": {"module." + k: v for k, v in ref["model"].items()}, "iterations": 3}, str(tmp_path / "dp.pkl"))
    assert ck.load_checkpoint(str(tmp_path / "dp.pkl"), QPNet(**kw)) == 3
