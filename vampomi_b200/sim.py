"""Seeded synthetic data in the reference's file formats.

Same generative model as the reference's example simulator (simulation/data_sim.py:35-49: X ~ N(0,1) i.i.d.,
CM = int(M*lam) causal markers with effects N(0, h2/CM), y = X beta + N(0, 1-h2)) and the same three files
(README.md:16-19): ``<name>.bin`` marker-major FP64, ``<name>.phen`` PLINK-style "FID IID value" with %0.10f
(data_sim.py:68), ``<name>_ts.bin`` true effects. Unlike the reference script it takes a seed, and it can emit a
binary (case/control) outcome for the probit model.
"""
import os

import numpy as np


def simulate(N, M, lam=0.1, h2=0.8, seed=1234, binary=False, dtype=np.float64):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((M, N)).astype(dtype)      # marker-major: row j is marker j over the N samples
    CM = max(int(M * lam), 1)
    idx = rng.choice(M, size=CM, replace=False)
    beta = np.zeros(M)
    beta[idx] = rng.normal(0.0, np.sqrt(h2 / CM), CM)
    g = beta @ X
    y = g + rng.normal(0.0, np.sqrt(1.0 - h2), N)
    if binary:
        y = (y > 0).astype(np.float64)
    return X, y, beta


def simulate_covariates(N, C, seed, y=None, effects=None, binary=False):
    """C raw covariates per individual (seeded, with a mean and a scale of their own so that standardisation matters) and,
    when y is given, a phenotype that carries their effects: continuous y gets +cov @ effects; a binary outcome is re-drawn
    from the shifted liability's sign (the original 0/1 flipped where the covariate shift dominates)."""
    rng = np.random.default_rng(seed + 7919)
    cov = rng.standard_normal((N, C)) * rng.uniform(0.5, 3.0, C) + rng.uniform(-2.0, 2.0, C)
    if y is None:
        return cov
    eff = np.asarray(effects if effects is not None else np.linspace(0.4, -0.3, C))
    shift = (cov - cov.mean(0)) / cov.std(0) @ eff
    if binary:
        liab = (2 * y - 1) * np.abs(rng.standard_normal(N)) + shift
        return cov, (liab > 0).astype(np.float64)
    return cov, y + shift


def write_covariates(path, cov):
    """The covariate file format of data::read_covariates (src/data.cpp:159-227): a header line, then `IID FID c1 .. cC`."""
    with open(path, "w") as f:
        f.write("IID FID " + " ".join(f"cov{j}" for j in range(cov.shape[1])) + "\n")
        for i, row in enumerate(cov):
            f.write("%d %d " % (i, i) + " ".join("%0.10f" % v for v in row) + "\n")


def write_phen(path, y):
    with open(path, "w") as f:
        for i, v in enumerate(y):
            f.write("%d %d %0.10f\n" % (i, i, v))


def write_dataset(out_dir, name, N, M, lam=0.1, h2=0.8, seed=1234, binary=False):
    os.makedirs(out_dir, exist_ok=True)
    X, y, beta = simulate(N, M, lam, h2, seed, binary)
    X.tofile(os.path.join(out_dir, name + ".bin"))
    write_phen(os.path.join(out_dir, name + ".phen"), y)
    beta.tofile(os.path.join(out_dir, name + "_ts.bin"))
    return X, y, beta


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--out-name", default="example")
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--M", type=int, default=2000)
    ap.add_argument("--lam", type=float, default=0.1)
    ap.add_argument("--h2", type=float, default=0.8)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--binary", action="store_true")
    a = ap.parse_args()
    write_dataset(a.out_dir, a.out_name, a.N, a.M, a.lam, a.h2, a.seed, a.binary)
