"""Shared helpers for the parity tests (fixture loading into oracle containers)."""
import numpy as np


def load_mpo(g, prefix="mpo"):
    return [g[f"{prefix}_{i}"] for i in range(int(g[prefix + "_n"]))]


def load_oracle_mps(g, prefix, meta="mps0"):
    from oracle.sweep import Mps
    n = int(g[prefix + "_n"])
    return Mps([g[f"{prefix}_{i}"] for i in range(n)],
               [g[f"{meta}_qn_{i}"] for i in range(n + 1)],
               [g[f"{meta}_sigmaqn_{i}"] for i in range(n)],
               g[meta + "_qntot"], int(g[meta + "_qnidx"]), bool(g[meta + "_to_right"]))


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def load_oracle_mpo(g, prefix="mpo"):
    from oracle.sweep import Mpo
    n = int(g[prefix + "_n"])
    return Mpo([g[f"{prefix}_{i}"] for i in range(n)], [g[f"{prefix}_qn_{i}"] for i in range(n + 1)],
               g[prefix + "_qntot"], int(g[prefix + "_qnidx"]))
