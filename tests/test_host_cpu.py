"""Host-side logic that needs no GPU: sweep validation, refusal of initialisations the path does not have, seed lists."""
import numpy as np
import pytest


def test_sweep_member_supported_limits():
    from demethify_b200.ic import sweep_member_supported
    assert sweep_member_supported(5, 25)             # the fixture: 5 known types, the reference's whole 1..25 sweep
    assert sweep_member_supported(6, 26)
    assert not sweep_member_supported(7, 25)         # even(7) + even(25) = 34 > 32
    assert sweep_member_supported(7, 24)
    assert sweep_member_supported(25, 6) and not sweep_member_supported(25, 7)
    assert sweep_member_supported(6, 0)


@pytest.mark.parametrize("option,exc", [("ICA", NotImplementedError), ("svd", ValueError), ("typo", ValueError)])
def test_bootstrap_refuses_inits_it_does_not_have(option, exc):
    """bootstrap_fits must not silently replace an initialisation (it used to fall back to uniform_): the same errors as the
    point-estimate path, raised before any resample is drawn or any device is touched."""
    from demethify_b200.bootstrap import bootstrap_fits
    X = np.random.RandomState(0).uniform(size=(20, 6))
    with pytest.raises(exc):
        bootstrap_fits(3, 1, X, np.ones_like(X), np.zeros((20, 2)), option, 5, 5, 1e-3, None, 1)


def test_bootstrap_n_u_above_samples_falls_back_like_the_reference():
    """n_u > N turns every option into uniform_ (deconvolution.py:44-45) before the dispatch: no error for ICA then; the call goes
    on to the device layer, which refuses without a GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from demethify_b200 import _lib
    from demethify_b200.bootstrap import bootstrap_fits
    X = np.random.RandomState(0).uniform(size=(20, 2))
    with pytest.raises(_lib.DmfError):
        bootstrap_fits(2, 3, X, np.ones_like(X), np.zeros((20, 2)), "ICA", 5, 5, 1e-3, None, 1)


def test_bootstrap_seed_sequence_and_restart_jobs():
    from demethify_b200.bootstrap import bootstrap_seeds
    assert bootstrap_seeds(1, 6) == [1, 2, 4, 7, 11, 16]          # bootstrap.py:27, SURVEY Q3
    with pytest.raises(TypeError):
        bootstrap_seeds([5], 3)                                    # `--seed 5` reaches bt_ci as a list (Q1)


def test_bench_clock_sampler_window():
    """bench.py's nvidia-smi sampler: only samples inside the timed region count; a region shorter than the sampling period falls
    back to all samples (warm-up ran the same kernels) and says so."""
    import importlib.util
    import os
    import time
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self):
            pass
    row = lambda mhz, cap: [str(mhz), "1965", "0x4", "Not Active", "Not Active", "Not Active", cap]
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    t = time.perf_counter()
    s.rows = [(t - 2.0, row(500, "Not Active")), (t - 1.0, row(1200, "Not Active"))]      # before the timed region
    s.mark_begin()
    s.rows += [(time.perf_counter(), row(1950, "Active")), (time.perf_counter(), row(1965, "Not Active"))]
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region"
    s2 = bench.ClockSampler(0)
    s2.proc = FakeProc()
    s2.rows = [(time.perf_counter() - 1.0, row(1900, "Not Active"))]
    s2.mark_begin()
    out2 = s2.stop()
    assert out2["samples"] == 1 and out2["sm_mhz"] == 1900.0 and out2["window"].startswith("warm-up")
