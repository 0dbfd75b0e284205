"""Checkpoint / config compatibility with the reference's training scripts.

* ``save_checkpoint`` writes what ``_save_checkpoint`` writes (qpnet_train.py:336-352):
  ``{"model": state_dict, "optimizer": state_dict, "iterations": int}`` under ``checkpoint-<iterations>.pkl``.
* ``load_checkpoint`` restores it the way ``main`` resumes (qpnet_train.py:481-489) and the decoder loads it
  (qpnet_decode.py:286-289): tensors are mapped to the CPU first, ``nn.DataParallel``'s ``module.`` prefix is accepted.
* ``load_config`` reads ``model.conf``, the pickled ``argparse.Namespace`` the trainer saves with ``torch.save``
  (qpnet_train.py:389) and returns the keyword arguments of ``QPNet``.

The parameter names and shapes of ``qpnet_b200.qpnet.QPNet`` are the reference's, so a checkpoint moves both ways.
"""
from __future__ import annotations

import os

import torch

_ARCH_KEYS = ("n_quantize", "n_aux", "n_resch", "n_skipch", "dilationF_depth", "dilationF_repeat", "dilationA_depth",
              "dilationA_repeat", "kernel_size", "upsampling_factor")


def save_checkpoint(checkpoint_dir, model, optimizer, iterations):
    """qpnet_train.py:336-352."""
    checkpoint = {"model": model.state_dict(), "optimizer": optimizer.state_dict(), "iterations": int(iterations)}
    os.makedirs(checkpoint_dir, exist_ok=True)
    path = os.path.join(checkpoint_dir, "checkpoint-%d.pkl" % iterations)
    torch.save(checkpoint, path)
    return path


def load_checkpoint(path, model, optimizer=None):
    """Restore ``model`` (and ``optimizer`` when given) from a reference-format checkpoint; returns ``iterations``."""
    checkpoint = torch.load(path, map_location=lambda storage, loc: storage, weights_only=False)
    state = checkpoint["model"]
    if all(k.startswith("module.") for k in state):              # saved from nn.DataParallel (qpnet_train.py:416-423)
        state = {k[len("module."):]: v for k, v in state.items()}
    model.load_state_dict(state)
    if optimizer is not None and "optimizer" in checkpoint:
        optimizer.load_state_dict(checkpoint["optimizer"])
    return int(checkpoint.get("iterations", 0))


def load_config(path):
    """``model.conf`` (pickled Namespace, qpnet_train.py:389) -> dict of QPNet constructor arguments."""
    cfg = torch.load(path, weights_only=False)
    ns = vars(cfg) if not isinstance(cfg, dict) else cfg
    return {k: ns[k] for k in _ARCH_KEYS if k in ns}
