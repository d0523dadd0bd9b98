"""Bridge between the device-resident training examples of BatchedSelfPlay and the reference's example
format.

The reference keeps ``trainExamplesHistory``: a list (one entry per iteration) of deques of
``(canonicalBoard, pi, v)`` tuples (Coach.py:66-72,96-106), pickles it to ``<checkpoint>.examples``
(Coach.py:144-152) and reads it back on resume (Coach.py:154-170); ``NNetWrapper.train`` turns the boards
into feature rows with ``state_to_vec`` (yacht/NNet.py:118-125).  Here the boards are YachtBoard objects
(same attributes, so ``state_to_vec`` and the reference's trainer accept them) and the conversion back to
tensors runs the feature kernel on the packed states.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

from . import _lib
from .engine import ACTION_SIZE, FEATURE_SIZE
from .layout import boards_to_planes, pack_state, planes_to_boards


def to_reference_examples(ex):
    """dict from BatchedSelfPlay(record_states=True).execute_episodes() -> list of (YachtBoard, pi, v) in the order
    Coach.executeEpisode emits them for game 0, then game 1, ... (pi: float64 list of 3226, v: float)."""
    if "states" not in ex:
        raise ValueError("examples carry no boards: run BatchedSelfPlay with record_states=True")
    states = ex["states"].cpu().numpy()                                   # [T, 2, n, 4]
    actions = ex["actions"].cpu().numpy().astype(np.int64) & 0xFFFF       # [T, n, k]
    counts = ex["counts"].cpu().numpy().astype(np.float64)
    value = ex["value"].cpu().numpy()
    plies, _, n, _ = states.shape
    out = []
    boards = [planes_to_boards(states[t]) for t in range(plies)]
    for g in range(n):
        for t in range(plies):
            pi = np.zeros(ACTION_SIZE, dtype=np.float64)
            np.add.at(pi, actions[t, g], counts[t, g])
            total = pi.sum()
            out.append((boards[t][g], (pi / total).tolist() if total > 0 else pi.tolist(), float(value[t, g])))
    return out


def from_reference_examples(examples, device="cuda"):
    """Iterable of (board, pi, v) -- boards: YachtBoard or the reference's YachtState -- -> dict(features
    float32[m, 59] (the ya_features kernel = state_to_vec), pi float32[m, 3226], value float32[m]) on the device:
    the tensors NNetWrapper.train builds per batch."""
    examples = list(examples)
    dev = torch.device(device)
    lib = _lib.load()
    m = len(examples)
    planes = torch.from_numpy(boards_to_planes([pack_state(b) for b, _, _ in examples]).view(np.int32)).to(dev)
    feats = torch.empty((m, FEATURE_SIZE), dtype=torch.float32, device=dev)
    if m:
        with torch.cuda.device(dev):
            _lib.check(lib.ya_features(_lib.ptr(planes), m, _lib.ptr(feats), m, _lib.current_stream()), "ya_features")
    pi = torch.tensor(np.asarray([p for _, p, _ in examples], dtype=np.float32).reshape(m, ACTION_SIZE), device=dev)
    v = torch.tensor(np.asarray([x for _, _, x in examples], dtype=np.float32), device=dev)
    return {"features": feats, "pi": pi, "value": v}


def save_train_examples(path, history):
    """Coach.saveTrainExamples (Coach.py:144-152): pickle of the list of per-iteration example collections."""
    with open(path, "wb+") as f:
        pickle.Pickler(f).dump(history)


def load_train_examples(path):
    """Coach.loadTrainExamples (Coach.py:154-170)."""
    with open(path, "rb") as f:
        return pickle.Unpickler(f).load()
