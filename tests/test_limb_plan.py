"""Host-side logic of the limb (path-decomposed) kernels: eligibility and the lane plan the engine derives
from a model (abr_limb_plan_host needs no device)."""
import numpy as np

from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from tests.conftest import MODELS


def _plan(name, **opt):
    mj = load_mj_model_from_file(MODELS[name][0])
    o = mj.opt.replace(**opt) if opt else None
    return mj, mjx.limb_plan(mj, o)


def test_quadruped_is_flat_four_lanes():
    mj, p = _plan("barkour")
    assert p["eligible"] and p["lanes"] == 4 and (p["NL"], p["NC"]) == (3, 1) and p["pattern"] == 2 and p["lanes_used"] == 4
    legs = set()
    for g, path in enumerate(p["paths"]):
        assert path[0] == 1 and len(path) == 4  # trunk, then one leg
        assert all(mj.body_parentid[path[k + 1]] == path[k] for k in range(3))
        legs.add(tuple(path[1:]))
        assert p["level"][g] == [2, 0, 0, 0] and p["own"][g][1:] == [1, 1, 1]
    assert len(legs) == 4 and sorted(b for leg in legs for b in leg) == list(range(2, 14))  # every body on exactly one path
    assert [o[0] for o in p["own"]] == [1, 0, 0, 0]  # the trunk has one owner


def test_biped_shares_the_torso_between_the_arm_paths():
    mj, p = _plan("biped")
    assert p["eligible"] and p["lanes"] == 4 and (p["NL"], p["NC"]) == (6, 4) and p["pattern"] == 86
    paths = p["paths"]
    shared = [g for g in range(4) if p["level"][g][1] == 1]
    assert len(shared) == 2 and shared[1] == shared[0] + 1 and shared[0] % 2 == 0  # an aligned pair of lanes
    a, b = shared
    assert paths[a][:4] == paths[b][:4] and paths[a][4:] != paths[b][4:]  # same torso, different arms
    assert p["own"][a][1:4] == [1, 1, 1] and p["own"][b][1:4] == [0, 0, 0]
    covered = sorted({body for path in paths for body in path if body > 0})
    assert covered == list(range(1, mj.nbody))
    assert p["ncon"] == 8 and p["nefc"] == 53


def test_tripod_nested_sharing_and_padding():
    mj, p = _plan("tripod")
    assert p["eligible"] and p["lanes"] == 4 and p["lanes_used"] == 4 and p["pattern"] not in (2, 86)
    assert (p["NL"], p["NC"]) == (6, 4)  # two contacts on one path (belly + foot) need the wider kernel
    assert sum(1 for lv in p["level"] if lv[1] == 1) == 2  # the fork's stem is shared by its two branches
    assert any(path[1] > 0 and path[-1] < 0 for path in p["paths"])  # short limbs are padded
    covered = sorted({body for path in p["paths"] for body in path if body > 0})
    assert covered == list(range(1, mj.nbody))


def test_three_limbs_leave_a_dummy_lane():
    mj, p = _plan("tripod3")
    assert p["eligible"] and p["lanes"] == 4 and p["lanes_used"] == 3
    assert sum(1 for path in p["paths"] if all(b < 0 for b in path[1:])) == 1  # the dummy lane carries the trunk only
    assert [o[0] for o in p["own"]].count(1) == 1
    covered = sorted({body for path in p["paths"] for body in path if body > 0})
    assert covered == list(range(1, mj.nbody))


def test_ineligible_models_and_options():
    assert not _plan("pendulum")[1]["eligible"]  # fixed base
    assert not _plan("bh280")[1]["eligible"]  # fixed base, joint equalities
    assert not _plan("barkour", solver=1)[1]["eligible"]  # CG
    assert not _plan("barkour", integrator=1)[1]["eligible"]  # RK4
    assert _plan("barkour", disableflags=16)[1]["eligible"]  # contacts off is fine
    assert _plan("barkour", disableflags=16)[1]["ncon"] == 0
