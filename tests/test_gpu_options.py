"""gsr_set_option knobs (include/gsr_b200.h): every configuration must produce the very same bits as the default one.

* ``blend_persist`` — the persistent blend draws its work items from ``gsr_counters.blend_next``;
* ``sort_single_pass`` — tiles of <= 2048 instances are sorted from one read of their keys (default) or by the three-pass path;
  the cases hold tiles above and below that size, equal depths (generic-network fallback) and empty tiles."""
import pytest
import torch

from tests import helpers as Hh
from autovfx_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["config1", "small_sh", "big_splats", "dense_tile", "coplanar"]


def _set(name, value):
    _lib.check(_lib.lib.gsr_set_option(name.encode(), int(value)), "gsr_set_option")


def _snapshot(a, for_backward):
    o = Hh.run_ours(a, for_backward=for_backward, sorted_keys=True)
    n = int(o["stats"]["num_rendered"])
    out = {k: o[k].clone() for k in ("color", "depth", "alpha", "radii")}
    out["point_list"] = o["views"]["point_list"][:n].clone()
    out["ranges"] = o["views"]["ranges"].clone()
    out["sorted_keys"] = o["views"]["sorted_keys"][:n].clone()
    if for_backward:
        out["n_contrib"] = o["views"]["n_contrib"].clone()
    return out


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("option,value", [("blend_persist", 5), ("blend_persist", 1), ("sort_single_pass", 0)])
def test_option_does_not_change_a_bit(name, option, value):
    a = Hh.resolve(Hh.case_inputs(name), DEV)
    ref = _snapshot(a, True)
    try:
        _set(option, value)
        got = _snapshot(a, True)
    finally:
        _set("blend_persist", 0)
        _set("sort_single_pass", 1)
    for k, v in ref.items():
        assert torch.equal(got[k], v), (option, value, k)


def test_persistent_blend_second_pass_on_reused_geometry():
    """GSR_FLAG_REUSE_GEOMETRY re-blends on the first pass's workspaces: the work cursor must be cleared again."""
    from autovfx_b200 import rasterizer as R
    a = Hh.resolve(Hh.case_inputs("config1"), DEV)
    extra = torch.rand(a["means3D"].shape[0], 3, generator=torch.Generator().manual_seed(5)).to(DEV)
    s = Hh.settings_from(a)

    def two_pass():
        with torch.no_grad():
            c1 = R.forward_raw(a["means3D"], a["shs"], None, a["opacities"], a["scales"], a["rotations"], None, s, sync=True)[0].clone()
            c2 = R.forward_raw(a["means3D"], None, extra, a["opacities"], a["scales"], a["rotations"], None, s, sync=True)[0].clone()
        return c1, c2
    ref = two_pass()
    try:
        _set("blend_persist", 4)
        got = two_pass()
    finally:
        _set("blend_persist", 0)
    assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
