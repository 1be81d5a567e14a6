import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.rl.base import VectorEnvStepper
from ambersim_b200.utils.io_utils import load_mj_model_from_file
mj = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml"); m = mjx.device_put(mj)
for E in (1024, 8192):
    q0 = torch.tensor(mj.key_qpos("home"), dtype=torch.float32, device="cuda").repeat(E,1); v0 = torch.zeros(E, mj.nv, device="cuda")
    ctrl = torch.tensor(mj.key_ctrl("home"), dtype=torch.float32, device="cuda").repeat(E,1)
    for ns in (1,2,4,8,16):
        st = VectorEnvStepper(m, q0, v0, nsubsteps=ns)
        for _ in range(20): st.step(ctrl)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        n=200; e0.record()
        for _ in range(n): st.step(ctrl)
        e1.record(); torch.cuda.synchronize()
        print(f"E={E} nsubsteps={ns}: {e0.elapsed_time(e1)/n*1e3:.2f} us per launch")
