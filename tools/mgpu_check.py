"""Multi-GPU check, run under torchrun on N GPUs of one box:
   sharded predictive sampling (samples split by global id, ONE exchange of the per-rank winners) must reproduce the
   single-GPU winner bit for bit on every rank, through both exchange paths: the NCCL all-gather (parallel.merge_best)
   and the NVLink peer-memory kernel (parallel.PeerExchange, csrc/abr_xchg.cu). The start state is off the goal and the
   guess is poor, so the winner is a noised sample that lives on a non-zero rank for some of the sizes.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, torch.distributed as dist
from ambersim_b200 import mjx
from ambersim_b200.parallel import PeerExchange, shard_range, sharded_optimize
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams
from ambersim_b200.utils.io_utils import load_mj_model_from_file

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
R = dist.get_world_size()
mj = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
m = mjx.device_put(mj)
nx, N = mj.nq + mj.nv, 32
q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
f = dict(dtype=torch.float32, device=dev)
rng = np.random.default_rng(0)
x0 = q0.copy(); x0[7:19] += rng.uniform(-0.15, 0.15, 12)
guess = np.clip(mj.key_ctrl("home") + 0.3 * rng.standard_normal((N, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
xch = PeerExchange(dev, capacity=4 * (2 + N * mj.nu + (N + 1) * nx))
ok = True


def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return out, float(ms)


for S in (1000, 4096, 65536, 1048576):
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=S, stdev=0.2)
    prm = VanillaPredictiveSamplerParams(key=3, x0=torch.tensor(x0, **f), us_guess=torch.tensor(guess, **f))
    xs1, us1, info = ps.optimize(prm, return_info=True)          # every rank solves the whole problem alone ...
    (xsN, usN), ms_nccl = timed(lambda: sharded_optimize(ps, prm))  # ... and its shard of it, exchanged through NCCL
    (xsP, usP), ms_p2p = timed(lambda: sharded_optimize(ps, prm, exchange=xch))  # ... or through the peer-memory kernel
    best = int(info["best_idx"])
    owner = next(r for r in range(R) if shard_range(S, r, R)[0] <= best < shard_range(S, r, R)[1])
    same = bool(torch.equal(xs1, xsN) and torch.equal(us1, usN) and torch.equal(xs1, xsP) and torch.equal(us1, usP))
    ok = ok and same and not xch.timed_out()
    if rank == 0:
        print(f"S={S}: sharded over {R} GPUs == single GPU (NCCL and peer-memory exchange): {same}; best_idx {best} (rank {owner}); "
              f"sharded solve {ms_nccl:.3f} ms with the NCCL all-gather, {ms_p2p:.3f} ms with the peer-memory kernel", flush=True)
# batched problems (B = 3) through the peer-memory exchange
ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=2048, stdev=0.2)
xb = torch.tensor(np.stack([x0, q0, x0 + 0.01]), **f)
gb = torch.tensor(np.stack([guess, guess * 0 + mj.key_ctrl("home"), guess]), **f)
prm = VanillaPredictiveSamplerParams(key=5, x0=xb, us_guess=gb)
xs1, us1 = ps.optimize(prm)
xsP, usP = sharded_optimize(ps, prm, exchange=xch)
same = bool(torch.equal(xs1, xsP) and torch.equal(us1, usP))
ok = ok and same and not xch.timed_out()
if rank == 0: print(f"B=3 batched problems through the peer-memory exchange == single GPU: {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
xch.close()
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
