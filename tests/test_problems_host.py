"""Host-side initial conditions (pluto_b200/problems.py): div B at round-off and
agreement with the reference's own start-up dump (golden s0) to 1e-12."""
import numpy as np
import pytest

from pluto_b200 import problems
from tests.util import Golden, divb_max

CASES = [("ot2d_plm_hlld", "ot"), ("ot3d_plm_hlld", "ot"), ("blast3d_plm_hlld", "blast"),
         ("blast2d_plm_hlld", "blast"), ("rotor2d_ppm_roe", "rotor"), ("turb3d_plm_hlld", "turb")]


@pytest.mark.parametrize("name,problem", CASES)
def test_ic_matches_reference_startup(name, problem):
    g = Golden(name)
    st, meta = problems.make(problem, g.dims, g.n)
    ref = g.states[0]
    for k, v in ref.items():
        assert st[k].shape == v.shape, k
        scale = max(np.abs(v).max(), 1.0)
        assert np.abs(st[k] - v).max() <= 2e-12 * scale, (name, k, np.abs(st[k] - v).max())
    assert np.allclose(meta["dx"], g.dx, rtol=1e-15)
    bscale = max(np.abs(st["Bx1s"]).max(), 1e-30) / min(g.dx)
    assert divb_max(st, g.dims, g.dx_zones) < 1e-12 * bscale


def test_subblock_generation_is_consistent():
    st, _ = problems.make("turb", 3, (8, 12, 16))
    sub, _ = problems.make("turb", 3, (8, 12, 16), offset=(2, 4, 8), count=(6, 8, 8))
    assert np.array_equal(sub["rho"], st["rho"][8:16, 4:12, 2:8])
    assert np.allclose(sub["vx2"], st["vx2"][8:16, 4:12, 2:8], rtol=0, atol=1e-15)
    assert np.allclose(sub["Bx3s"], st["Bx3s"][8:17, 4:12, 2:8], rtol=0, atol=1e-13)
