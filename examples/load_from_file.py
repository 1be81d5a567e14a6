"""Loading robots from URDF / MJCF files into the engine (the reference's examples/load_from_file.py with the import line changed).

All of the following work: (1) a global path, (2) a path relative to the working directory, (3) a path relative to the package root;
strings or Path objects; `force_float=True` floats the base. Needs a CUDA device (mjx.make_data allocates on it: the engine has no CPU path).
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ambersim_b200 import ROOT
from ambersim_b200.utils.io_utils import load_mjx_model_and_data_from_file

mjx_model1, mjx_data1 = load_mjx_model_and_data_from_file(ROOT + "/models/pendulum/pendulum.urdf")  # (1)
mjx_model3, mjx_data3 = load_mjx_model_and_data_from_file("models/pendulum/pendulum.urdf")  # (3)
hand, _ = load_mjx_model_and_data_from_file(Path("models/barrett_hand/bh280.urdf"))
floating_hand, _ = load_mjx_model_and_data_from_file("models/barrett_hand/bh280.xml", force_float=True)
print("pendulum:", mjx_model1.nq, "dof,", mjx_model1.nu, "actuator;  bh280:", hand.nq, "dofs,", hand.nu, "motors from the URDF's transmissions,",
      hand.neq, "joint equalities from its mimic joints;  floated:", floating_hand.nq, "qpos entries")
