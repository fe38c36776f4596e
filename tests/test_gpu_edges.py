"""Edge cases of the C ABI on the GPU: ragged batch sizes, single vectors, out-of-range bins,
argument errors. Every call goes through libbaofit_b200.so."""
import os

import numpy as np
import pytest

import oracle
from baofit_b200 import capi, workloads as W
from baofit_b200.desc import BF_STATUS_FFT_RANGE
from util import RTOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 3, 65, 130])
def test_ragged_batches_rspace_and_gram(ctx, B):
    m, d = W.qso_rspace()
    rng = np.random.Generator(np.random.PCG64(B))
    params = m.random_batch(B, rng)
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, st = capi.chi2_batch(ctx, gm, gd, params)
    ref, rst = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, nthreads=8)
    assert np.array_equal(st, rst) and rel_err(chi2, ref) < RTOL
    # one probe group per vector (G = 1): gram[f][0][0] is chi2
    gram, gst = capi.gram_batch(ctx, gm, gd, params, 1)
    assert np.all(gst == 0) and rel_err(gram[:, 0, 0], ref) < RTOL
    gm.close(), gd.close()


def test_fft_bins_outside_the_kept_plane_are_flagged(ctx):
    """A dilation that carries bins beyond fft_rmax flags BF_STATUS_FFT_RANGE and yields NaN, like the
    dilation limits of the k-space model; vectors inside are unaffected. Bit-exact status vs oracle."""
    m, coords = W.dr11_auto_fft(ngrid=128, spacing=5., nbins=40, fft_rmax=200.)
    om = oracle.OracleModel(m)
    d, _ = W.synthetic_dataset(m, coords, lambda mm, dd, p: oracle.predict(om, oracle.OracleData(dd), p[0])[0])
    params = np.tile(m.values, (5, 1))
    params[:, m.index("BAO alpha-parallel")] = [1.0, 1.05, 1.3, 0.9, 1.5]
    params[:, m.index("BAO alpha-perp")] = [1.0, 1.1, 1.0, 1.4, 0.7]
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, st = capi.chi2_batch(ctx, gm, gd, params)
    pred, st2 = capi.eval_batch(ctx, gm, gd, params)
    ref, rst = oracle.chi2_batch(om, oracle.OracleData(d), params, nthreads=8)
    assert np.array_equal(st, rst) and np.array_equal(st2, rst)
    assert set(np.unique(rst)) == {0, BF_STATUS_FFT_RANGE}
    ok = rst == 0
    assert rel_err(chi2[ok], ref[ok]) < RTOL and np.all(np.isnan(chi2[~ok])) and np.all(np.isnan(pred[:, ~ok]))
    gm.close(), gd.close()


def test_argument_errors_are_reported_not_crashed(ctx):
    m, d = W.qso_rspace()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    with pytest.raises(capi.BaofitError):
        capi.gram_batch(ctx, gm, gd, np.tile(m.values, (200, 1)), 100)  # G > 96
    truth = np.zeros(gd.n)
    toys = capi.Toys(ctx, gd, truth, seed=1, first_toy=0, ntoy=3)
    with pytest.raises(capi.BaofitError):
        capi.chi2_batch_toys(ctx, gm, gd, m.values[None, :], toys, np.array([3]))  # toy index out of range
    with pytest.raises(capi.BaofitError):
        gm.fft_planes(m.values[None, :], 2.3)  # not an FFT model
    # multipole data refuse cross-correlation models (the reference's multipole overload applies no velocity shift)
    _, dm = W.demo_multi(lmax=4)
    gdm = capi.Data(ctx, dm)
    with pytest.raises(capi.BaofitError):
        capi.chi2_batch(ctx, gm, gdm, m.values[None, :])
    toys.close(), gm.close(), gd.close(), gdm.close()


def test_fft_model_rejects_bad_grids(ctx):
    m, _ = W.dr11_auto_fft(ngrid=4, spacing=4.)
    with pytest.raises(capi.BaofitError):
        capi.Model(ctx, m)
    m, _ = W.dr11_auto_fft(ngrid=64, spacing=-1.)
    with pytest.raises(capi.BaofitError):
        capi.Model(ctx, m)


def test_two_devices_in_one_process(ctx):
    """Contexts on two GPUs of one process give identical results (per-device kernel attributes such as the
    dynamic shared-memory opt-in must be set on each device). Needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.Generator(np.random.PCG64(31))
    ctx1 = capi.Context(1)
    for m, d in (W.qso_kspace(), W.qso_rspace()):
        p = m.random_batch(300, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.7, 1.3)})
        g0, d0 = capi.Model(ctx, m), capi.Data(ctx, d)
        g1, d1 = capi.Model(ctx1, m), capi.Data(ctx1, d)
        a, sa = capi.chi2_batch(ctx, g0, d0, p)
        b, sb = capi.chi2_batch(ctx1, g1, d1, p)
        assert np.array_equal(a, b) and np.array_equal(sa, sb)
        for h in (g0, d0, g1, d1):
            h.close()
    ctx1.close()


def test_plans_do_not_outlive_their_model(ctx):
    """A (model, data) plan must never be found again by a *different* model that happens to reuse the address
    of a destroyed one: create A (no distortion matrix, no metals), evaluate, destroy it, create B (distortion
    matrix + metals) against the same Data object -- B's results must equal those of B on a fresh Data."""
    mA, dA = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    mB, dB = W.qso_kspace()
    rng = np.random.Generator(np.random.PCG64(77))
    pA, pB = mA.random_batch(5, rng), mB.random_batch(5, rng)
    gd = capi.Data(ctx, dB)  # carries the grid: usable by both models
    fresh = capi.Data(ctx, dB)
    for rep in range(6):  # several rounds so that the allocator does hand an old address out again
        gA = capi.Model(ctx, mA)
        cA, _ = capi.chi2_batch(ctx, gA, gd, pA)
        gA.close()
        gB = capi.Model(ctx, mB)
        cB, _ = capi.chi2_batch(ctx, gB, gd, pB)
        cB_fresh, _ = capi.chi2_batch(ctx, gB, fresh, pB)
        gB.close()
        assert np.array_equal(cB, cB_fresh)
        if rep == 0:
            cA0, cB0 = cA, cB
        assert np.array_equal(cA, cA0) and np.array_equal(cB, cB0)
    ref, _ = oracle.chi2_batch(oracle.OracleModel(mB), oracle.OracleData(dB), pB[:1], nthreads=8)
    assert rel_err(cB0[:1], ref) < RTOL
    gd.close(), fresh.close()


def test_two_contexts_and_a_shared_context_from_two_threads(ctx):
    """Runs on a one-GPU box: (1) two host threads, each with its own context / model / data on device 0,
    evaluate concurrently; (2) two host threads share ONE context (calls are serialised by the context's lock).
    Either way every thread gets exactly the single-threaded answer."""
    import threading
    m, d = W.qso_kspace()
    rng = np.random.Generator(np.random.PCG64(3))
    batches = [m.random_batch(700, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)}) for _ in range(2)]
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    want = [capi.chi2_batch(ctx, gm, gd, b)[0] for b in batches]
    got, errs = [None, None], []

    def own_context(k):
        try:
            c = capi.Context(0)
            lm, ld = capi.Model(c, m), capi.Data(c, d)
            for _ in range(8):
                got[k] = capi.chi2_batch(c, lm, ld, batches[k])[0]
            lm.close(), ld.close(), c.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def shared_context(k):
        try:
            for _ in range(8):
                got[k] = capi.chi2_batch(ctx, gm, gd, batches[k])[0]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    for target in (own_context, shared_context):
        got[0] = got[1] = None
        th = [threading.Thread(target=target, args=(k,)) for k in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for k in range(2):
            assert np.array_equal(got[k], want[k], equal_nan=True), target.__name__
    gm.close(), gd.close()


def test_device_entry_points_do_not_wait_for_the_gpu(ctx):
    """bf_chi2_batch_dev is asynchronous on the context's stream: which k-space chain runs (shared basis tables or
    per-vector transforms) and whether the tables are rebuilt is decided on the device, so the call returns while
    the GPU is still working -- for both kinds of batch -- and the results are those of the host-buffer call."""
    import time
    import torch
    m, d = W.qso_kspace()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    B = 37888
    rng = np.random.Generator(np.random.PCG64(12))
    shared = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)})
    mixed = shared.copy()
    mixed[::2, m.index("SigmaNL-perp")] *= 1.05  # two keys in one batch: per-vector transforms
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for name, p in (("shared", shared), ("mixed", mixed)):
        want, want_st = capi.chi2_batch(ctx, gm, gd, p)  # also warms up workspaces (growing them may synchronise)
        d_p = torch.from_numpy(p).to(dev)
        d_chi2 = torch.empty(B, dtype=torch.float64, device=dev)
        d_st = torch.empty(B, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        t0 = time.perf_counter()
        for _ in range(4):
            capi.chi2_batch_dev(ctx, gm, gd, d_p.data_ptr(), B, d_chi2.data_ptr(), d_st.data_ptr())
        host_s = time.perf_counter() - t0
        busy_after_return = not ext.query()
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        gpu_s = e0.elapsed_time(e1) * 1e-3
        assert busy_after_return, name
        assert host_s < 0.5 * gpu_s, (name, host_s, gpu_s)
        got = d_chi2.cpu().numpy()
        assert np.array_equal(d_st.cpu().numpy(), want_st)
        ok = want_st == 0
        # the host-buffer call evaluates a large batch in pieces (upload overlapped with the kernels); a piece can fall
        # below the batch size at which the chi2 GEMM splits its row tiles over CTAs, so sums may differ in the last bits
        rel = float(np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok])))
        assert rel < 1e-13 and np.all(np.isnan(got[~ok])), (name, int(np.sum(got[ok] != want[ok])), rel)
    gm.close(), gd.close()


@pytest.mark.gpu
def test_submit_wait_pipeline_matches_the_synchronous_call(ctx):
    """bf_chi2_batch_submit / bf_chi2_batch_wait: two batches in flight, results identical to bf_chi2_batch; a third
    submit without a wait and a stale ticket are refused."""
    import torch
    m, d = W.qso_kspace()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    rng = np.random.Generator(np.random.PCG64(21))
    B = 5000
    batches = [m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.7, 1.3)}) for _ in range(5)]
    want = [capi.chi2_batch(ctx, gm, gd, p) for p in batches]
    hp = [torch.from_numpy(p).pin_memory() for p in batches]
    hc = [torch.empty(B, dtype=torch.float64).pin_memory() for _ in batches]
    hs = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in batches]
    tickets = []
    for i in range(5):
        tickets.append(capi.chi2_batch_submit(ctx, gm, gd, hp[i].data_ptr(), B, hc[i].data_ptr(), hs[i].data_ptr()))
        if len(tickets) == 2:
            if i == 1:
                with pytest.raises(capi.BaofitError):  # a third call in flight
                    capi.chi2_batch_submit(ctx, gm, gd, hp[2].data_ptr(), B, hc[2].data_ptr(), hs[2].data_ptr())
            capi.chi2_batch_wait(ctx, tickets.pop(0))
    last = tickets.pop(0)
    capi.chi2_batch_wait(ctx, last)
    with pytest.raises(capi.BaofitError):
        capi.chi2_batch_wait(ctx, last)  # already collected
    for i in range(5):
        chi2, st = want[i]
        assert np.array_equal(hs[i].numpy(), st)
        ok = st == 0
        assert np.array_equal(hc[i].numpy()[ok], chi2[ok])
    gm.close(), gd.close()


@pytest.mark.gpu
def test_reference_mid_sweep_retransform_is_flagged_not_silently_different(ctx):
    """Where the reference would re-run its k-space transforms in the middle of a sweep over the bins (consecutive bins
    whose redshifts differ by more than dzmin, BaoKSpaceCorrelationModel.cc:317-319) and D(k, mu) depends on z through
    gamma-beta, the answer depends on the bin order and is not reproduced: such vectors come back with
    BF_STATUS_Z_RETRIGGER and NaN; with gamma-beta = 0 (as config/BOSSDR9LyaF_k.ini fixes it), or dzmin above the jumps
    between DR9's three z bins, nothing is flagged."""
    from baofit_b200.desc import BF_STATUS_Z_RETRIGGER
    m, d = W.dr9_kspace()
    rng = np.random.Generator(np.random.PCG64(31))
    p = m.random_batch(6, rng)
    jg = m.index("gamma-beta")
    p[:3, jg] = 0.0
    p[3:, jg] = 0.7
    m.desc.dzmin = 2.0  # no jump between consecutive bins is larger
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, st = capi.chi2_batch(ctx, gm, gd, p)
    assert np.all(st == 0) and np.all(np.isfinite(chi2))
    gm.close(), gd.close()
    m.desc.dzmin = 0.5  # the reference's default: the sweep over DR9's bins crosses from one z bin to another
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2b, stb = capi.chi2_batch(ctx, gm, gd, p)
    assert np.all(stb[:3] == 0) and np.array_equal(chi2b[:3], chi2[:3])
    assert np.all(stb[3:] == BF_STATUS_Z_RETRIGGER) and np.all(np.isnan(chi2b[3:]))
    gm.close(), gd.close()
    m.desc.dzmin = 0.5


@pytest.mark.gpu
def test_small_call_graphs_replay_the_same_answers(ctx):
    """Small host-buffer chi2 calls (B <= 256) are captured into a CUDA graph on their second occurrence and replayed from
    the third on, as long as the key parameters stay those the device tables were built for. Same answers as a context
    that never uses graphs (BAOFIT_NO_GRAPHS), across a change of the key and back, for the k-space and r-space models."""
    os.environ["BAOFIT_NO_GRAPHS"] = "1"
    try:
        plain = capi.Context(0)
    finally:
        os.environ.pop("BAOFIT_NO_GRAPHS", None)
    rng = np.random.Generator(np.random.PCG64(77))
    for m, d in (W.qso_kspace(), W.qso_rspace()):
        gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
        pm, pd_ = capi.Model(plain, m), capi.Data(plain, d)
        kspace = "SigmaNL-perp" in m.names
        for B in (1, 7):
            batches = [m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.9, 1.1), "BAO alpha-perp": (0.9, 1.1)}) for _ in range(9)]
            if kspace:
                for q in (4, 5):  # another key for two calls, then back
                    batches[q][:, m.index("SigmaNL-perp")] *= 1.07
            l0 = ctx.launches
            per_call = []
            for p in batches:
                before = ctx.launches
                got, st = capi.chi2_batch(ctx, gm, gd, p)
                per_call.append(ctx.launches - before)
                want, wst = capi.chi2_batch(plain, pm, pd_, p)
                assert np.array_equal(st, wst)
                ok = wst == 0
                assert np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok])) < 1e-13, (B, per_call)
            # the launch counter keeps counting the kernels a replayed graph runs
            assert per_call[-1] == per_call[2] and per_call[-1] > 0, per_call
        gm.close(), gd.close(), pm.close(), pd_.close()
    plain.close()
