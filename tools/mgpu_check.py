"""Multi-GPU check, run under torchrun on N GPUs of one box:
   sharded predictive sampling (samples split by global id, one all-gather of the per-rank winners)
   must reproduce the single-GPU winner bit for bit on every rank.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, torch.distributed as dist
from ambersim_b200 import mjx
from ambersim_b200.parallel import sharded_optimize
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams
from ambersim_b200.utils.io_utils import load_mj_model_from_file

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mj = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
m = mjx.device_put(mj)
nx = mj.nq + mj.nv
q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
f = dict(dtype=torch.float32, device=dev)
ok = True
for S in (1000, 4096, 65536):
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=S, stdev=0.1)
    prm = VanillaPredictiveSamplerParams(key=3, x0=torch.tensor(q0, **f), us_guess=torch.tensor(mj.key_ctrl("home"), **f).repeat(32, 1))
    xs1, us1, info = ps.optimize(prm, return_info=True)          # every rank solves the whole problem alone ...
    for _ in range(2): sharded_optimize(ps, prm)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); xsN, usN = sharded_optimize(ps, prm); e1.record(); torch.cuda.synchronize()   # ... and its shard of it
    same = bool(torch.equal(xs1, xsN) and torch.equal(us1, usN))
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = ok and same
    if rank == 0: print(f"S={S}: sharded over {dist.get_world_size()} GPUs == single GPU: {same}; best_idx {int(info['best_idx'])}; sharded solve {float(ms):.3f} ms", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
