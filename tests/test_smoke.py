"""The driver's smoke() as a regression test: one small garment through the whole CUDA hot path, every stage checked against
the oracle (it caught an accuracy regression of the tensor-core Linear block that the per-op tests did not)."""
import pytest


@pytest.mark.gpu
def test_graft_entry_smoke(dev):
    import __graft_entry__ as g
    g.smoke()
