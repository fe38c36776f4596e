#!/usr/bin/env python
"""Generates tests/golden/oracle_<case>.npz: the CPU oracle (oracle/oracle.c) run ONCE on the seeded
parameter vectors of every case of tests/parity_cases.py -- BASELINE configs 2-5 at their full sizes with
256 vectors each, and the option / metal-branch cases with 8 -- so that the GPU parity tests compare per-bin
predictions and chi-square at 1e-10 on hundreds of vectors without paying the oracle's brute-force k->r
quadrature (about one second on eight host threads per vector) on every run.

    python tools/gen_oracle_fixtures.py [--cases cfg2,cfg3,...|all] [--threads N]

Every file holds: params [B][npar], pred [n_data][B], chi2 [B], status [B], names, and for the synthetic
data sets (cfg3 and cfg5: the reference ships no covariance for them) the oracle's xi(p0) from which
workloads.synthetic_dataset rebuilds the data vector and covariance deterministically.
tests/test_oracle_fixtures.py reads them (GPU: parity; CPU: the oracle re-run on two vectors per file).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import parity_cases as PC  # noqa: E402


def oracle_predict(m, d, p):
    return oracle.predict(oracle.OracleModel(m), oracle.OracleData(d), p[0])[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="all")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    oracle.build()
    cases = PC.ALL if a.cases == "all" else a.cases.split(",")
    for name in cases:
        t0 = time.perf_counter()
        m, d, sweep, B, xi0 = PC.build(name, predict=oracle_predict)
        params = PC.make_params(name, m, sweep, B)
        om, od = oracle.OracleModel(m), oracle.OracleData(d)
        chi2, status, pred = oracle.chi2_batch(om, od, params, want_pred=True, nthreads=a.threads)
        extra = {} if xi0 is None else {"xi0": xi0}
        np.savez_compressed(PC.fixture_path(name), params=params, pred=pred, chi2=chi2, status=status,
                            names=np.array(m.names), n_data=d.n_data, **extra)
        print("%s: B=%d n_data=%d npar=%d -> %s (%.0f s, %d flagged vectors)" %
              (name, B, d.n_data, m.desc.npar, PC.fixture_path(name), time.perf_counter() - t0, int((status != 0).sum())), flush=True)


if __name__ == "__main__":
    main()
