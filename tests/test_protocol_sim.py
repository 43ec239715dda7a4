"""The mbarrier protocols of the tcgen05 attention kernels, replayed under random schedules by
tools/mbar_sim.py (CPU only).  The shipped protocols must never deadlock or read a stale buffer; the
two known-broken variants kept in the model must be caught (a self-test of the model's power)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mbar_sim", os.path.join(ROOT, "tools", "mbar_sim.py"))
mbar_sim = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mbar_sim)


@pytest.mark.parametrize("variant", ["gemm", "fwd_v2", "dq", "dkdv"])
def test_shipped_protocols_are_clean(variant):
    assert mbar_sim.check(variant, trials=300) is None


@pytest.mark.parametrize("variant", ["fwd_lazy", "fwd_pbuf2", "dkdv_pbuf2"])
def test_prepared_protocols_are_clean(variant):
    assert mbar_sim.check(variant, trials=300) is None


@pytest.mark.parametrize("variant", ["gemm_unpaced_release", "fwd_lazy_bad", "fwd_pbuf2_badfinal", "dkdv_pbuf2_bad"])
def test_model_catches_known_bugs(variant):
    bad = mbar_sim.check(variant, trials=1500)
    assert bad is not None, "the model no longer finds the known deadlock / race"
