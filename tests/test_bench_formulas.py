"""bench.py's algorithmic-work table: every compute entry point of the C ABI has a formula, so that no
launch enters the step roofline with zero work (a missing formula reads as "infinitely far from its roof")."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# not launches of their own on a training step: queries, workspace sizes, ops that are not on any step
NOT_ON_A_STEP = ("_eligible", "_workspace_bytes", "pasnl_version", "pasnl_error_string", "pasnl_grid_subsample",
                 "pasnl_prob_sample", "pasnl_selection_sort")


def test_every_step_entry_point_has_algorithmic_work():
    bench = _bench()
    from pointasnl_b200 import _lib
    args = (64, 32, 16, 8, 4, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1)
    missing = []
    for name in _lib.SIGNATURES:
        if any(tag in name for tag in NOT_ON_A_STEP):
            continue
        nbytes, flops, bound = bench.algorithmic_work(name, args)
        assert bound in ("hbm", "tensor", "fp32", "latency"), (name, bound)
        if nbytes <= 0:
            missing.append(name)
    assert not missing, f"bench.algorithmic_work has no formula for {missing}"


def test_roof_seconds_takes_the_larger_of_the_two_roofs():
    bench = _bench()
    peaks = {"hbm_gbs": 6574.8, "bf16_tflops": 1724.4}
    assert abs(bench.roof_seconds(6574.8e9, 0, "hbm", peaks) - 1.0) < 1e-12
    # 3xTF32: three passes at half the bf16 rate
    assert abs(bench.roof_seconds(0, 1724.4e12 / 6, "tensor", peaks) - 1.0) < 1e-12
    assert bench.roof_seconds(6574.8e9, 1724.4e12, "tensor", peaks) > 5.9
    assert abs(bench.roof_seconds(0, bench.FP32_SIMT_TFLOPS * 1e12, "fp32", peaks) - 1.0) < 1e-12
