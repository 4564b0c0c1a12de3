"""Injected-prior models for search parity (TEST INFRASTRUCTURE ONLY).

Each class follows the reference's ``nn_model`` protocol (new_mcts_alpha.py:161):
``predict(X f32[B,3,15,15]) -> (probs f32[B,225], values f32[B,1])``.  Priors are
pure functions of the position built from integer hashing and one correctly
rounded division, so they are identical on every machine (no libm involved).
Values are arbitrary on purpose: the reference never backs them up.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def position_hash(X: np.ndarray) -> np.ndarray:
    """uint64[B]: hash of the two stone planes of each row."""
    B = X.shape[0]
    s = (X[:, 0].reshape(B, -1) > 0.5).astype(np.uint64) + np.uint64(2) * (X[:, 1].reshape(B, -1) > 0.5).astype(np.uint64)
    h = np.full(B, 0x1234567, dtype=np.uint64)
    for i in range(s.shape[1]):
        h = _mix(h ^ (s[:, i] + np.uint64(3 * i + 1)))
    return h


class _Base:
    def __init__(self):
        self.rows = 0
        self.calls = 0

    def weights(self, X):          # uint64[B,225] positive integer weights
        raise NotImplementedError

    def predict(self, X):
        X = np.asarray(X)
        w = self.weights(X).astype(np.float64)
        probs = (w / w.sum(axis=1, keepdims=True)).astype(np.float32)
        h = position_hash(X)
        values = (((h >> np.uint64(11)) % np.uint64(2001)).astype(np.float64) / 1000.0 - 1.0).astype(np.float32)
        self.rows += X.shape[0]
        self.calls += 1
        return probs, values.reshape(-1, 1)


class Uniform(_Base):
    """probs = 1/225 everywhere (the SURVEY 8c known-answer model)."""

    def predict(self, X):
        X = np.asarray(X)
        self.rows += X.shape[0]
        self.calls += 1
        return (np.full((X.shape[0], 225), 1.0 / 225.0, dtype=np.float32),
                np.zeros((X.shape[0], 1), dtype=np.float32))


class Hashed(_Base):
    """Smooth pseudo-random priors: weight in 1..1000 per (position, action)."""

    def weights(self, X):
        h = position_hash(X)[:, None]
        a = np.arange(225, dtype=np.uint64)[None, :]
        return _mix(h ^ (a * np.uint64(0x100000001B3))) % np.uint64(1000) + np.uint64(1)


class Spiky(_Base):
    """A few actions carry almost all the mass (forces deep, narrow trees and
    the masked-sum < 1e-8 fallback once the spikes are occupied)."""

    def weights(self, X):
        h = position_hash(X)[:, None] >> np.uint64(40)      # spikes move slowly with the position
        a = np.arange(225, dtype=np.uint64)[None, :]
        hit = (_mix(h ^ (a * np.uint64(0x9E3779B1))) % np.uint64(61)) == 0
        return np.where(hit, np.uint64(10 ** 12), np.uint64(1))


class FixedSpike(_Base):
    """All mass on three fixed cells; once they are taken the masked prior sums
    below 1e-8 and the uniform-legal fallback fires (new_mcts_alpha.py:167-168)."""

    def weights(self, X):
        w = np.ones((X.shape[0], 225), dtype=np.uint64)
        w[:, [112, 113, 97]] = np.uint64(10 ** 15)
        return w


BY_NAME = {"uniform": Uniform, "hashed": Hashed, "spiky": Spiky, "fixedspike": FixedSpike}
