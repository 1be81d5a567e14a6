import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


MODELS = {
    "pendulum": ("models/pendulum/scene.xml", None),
    "bh280": ("models/barrett_hand/bh280.xml", None),
    "barkour": ("models/barkour_standin/barkour_vb_standin.xml", "home"),
    "biped": ("models/biped_standin/biped_exo_standin.xml", "stand"),
    "exolegs": ("models/biped_standin/exo_legs_standin.xml", "stand"),  # legs only: the flat 2-lane class of the limb kernels
    # test-only fixture for the limb kernels: 3 limbs (dummy lane), a forked limb (nested sharing), padding,
    # slide joint, condim-1 contacts, a trunk contact, a tilted floor
    "tripod": (str(ROOT / "tests/models/tripod.xml"), "home"),
    "tripod3": (str(ROOT / "tests/models/tripod3.xml"), "home"),
    "fixedbase": (str(ROOT / "tests/models/fixedbase.xml"), None),
    "boxbot": (str(ROOT / "tests/models/boxbot.xml"), "home"),  # plane - convex collision: a box and a convex mesh foot (4 contacts each) + a sphere  # hand kernels: static base, three chains, four joint equalities  # three leaf paths: the 4-lane group carries a dummy lane
    "bh280_hulls": (str(ROOT / "tests/models/bh280_hulls.xml"), None),  # the Barrett hand with the small hulls of its real collision geometry
    "blocks": (str(ROOT / "tests/models/blocks.xml"), "home"),  # sphere / capsule / convex - convex collision: box, capsule, sphere and mesh geoms over a static box and a static wedge
}


@pytest.fixture(scope="session")
def load_model():
    from ambersim_b200.utils.io_utils import load_mj_model_from_file

    cache = {}

    def _load(name, **kw):
        key = (name, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = load_mj_model_from_file(MODELS[name][0], **kw)
        return cache[key]

    return _load
