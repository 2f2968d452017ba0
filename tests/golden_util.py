"""Loaders for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py
from the unmodified reference C)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mp_cases():
    z = np.load(os.path.join(GOLD, "mp_cases.npz"))
    for i in range(int(z["num_cases"][0])):
        p = f"c{i:03d}_"
        n, k, dv, dc, its = (int(v) for v in z[p + "dims"])
        yield dict(n=n, k=k, dv=dv, dc=dc, its=its, v2c=z[p + "v2c"], c2v=z[p + "c2v"], inp=z[p + "in"],
                   prefill=z[p + "prefill"], out=z[p + "out"], errors=z[p + "errors"], ret=int(z[p + "ret"][0]))


def mc_cases():
    z = np.load(os.path.join(GOLD, "mc_cases.npz"))
    for i in range(int(z["num_cases"][0])):
        p = f"m{i:02d}_"
        n, k, dv, dc, its, exp, frames, first, seed = (int(v) for v in z[p + "dims"])
        yield dict(n=n, k=k, dv=dv, dc=dc, its=its, expurgation=exp, frames=frames, first=first, seed=seed,
                   eps=z[p + "eps"], v2c=z[p + "v2c"], c2v=z[p + "c2v"], counters=z[p + "counters"])


def rng_words():
    z = np.load(os.path.join(GOLD, "rng_words.npz"))
    for t in range(int(z["num"][0])):
        seed, point, q = (int(v) for v in z[f"w{t}_key"])
        yield seed, point, q, float(z[f"w{t}_eps"][0]), z[f"w{t}_words"]
