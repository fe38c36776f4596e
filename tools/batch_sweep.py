"""Latency and throughput of bf_chi2_batch (host buffers in, chi2 out) against the batch size B
(B in {1, 32, 256, 4096, 65536}) for the headline k-space workload and the
r-space QSO workload. Prints one JSON object."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from baofit_b200 import capi, workloads as W  # noqa: E402

ctx = capi.Context(0)
out = {}
for name in ("qso_kspace", "qso_rspace"):
    m, d = getattr(W, name)()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    rng = np.random.Generator(np.random.PCG64(3))
    rows = []
    for B in (1, 32, 256, 4096, 65536):
        p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.7, 1.3)})
        for _ in range(5):
            capi.chi2_batch(ctx, gm, gd, p)
        reps = 200 if B <= 4096 else 20
        t0 = time.perf_counter()
        for _ in range(reps):
            capi.chi2_batch(ctx, gm, gd, p)
        dt = (time.perf_counter() - t0) / reps
        # the same call on preallocated buffers through raw addresses (what a C caller pays: no numpy allocation / conversion per call)
        pc = np.ascontiguousarray(p)
        chi2, st = np.empty(B), np.empty(B, dtype=np.int32)
        for _ in range(5):
            capi.chi2_batch_raw(ctx, gm, gd, pc.ctypes.data, B, chi2.ctypes.data, st.ctypes.data)
        t0 = time.perf_counter()
        for _ in range(reps):
            capi.chi2_batch_raw(ctx, gm, gd, pc.ctypes.data, B, chi2.ctypes.data, st.ctypes.data)
        dt_raw = (time.perf_counter() - t0) / reps
        rows.append({"B": B, "us_per_call": dt * 1e6, "us_per_call_raw_buffers": dt_raw * 1e6, "evals_per_s": B / dt})
    out[name] = rows
print(json.dumps(out, indent=1))
