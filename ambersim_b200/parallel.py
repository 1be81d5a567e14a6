"""Multi-GPU sharding of predictive sampling: one process per GPU, torch.distributed plumbing.

Samples are independent, so rank r of R owns the contiguous global sample ids
[r*S/R, (r+1)*S/R); the noise is keyed by the GLOBAL sample id, hence the union over ranks equals
the single-GPU result. The only exchange is one all_gather of a small per-rank record
{best_cost, best_idx, us*[N,nu], xs*[N+1,nx]}; every rank then takes the first minimum (NaN counts
as minimal, lower global index wins ties) so all ranks end with identical winners and no dependent
broadcast is needed. Rollout / env-step throughput needs no collective at all (replicas).

Two implementations of that exchange: `merge_best` (torch.distributed all_gather: NCCL on GPUs, gloo in
the CPU tests) and `PeerExchange` (ONE kernel launch per rank that stores the record into every peer's
buffer over NVLink P2P and selects the winner, csrc/abr_xchg.cu; GPUs of one box only).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(nsamples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous global-sample range [lo, hi) owned by `rank` (sample 0, the guess, is on rank 0)."""
    per = (nsamples + world_size - 1) // world_size
    lo = min(nsamples, rank * per)
    return lo, min(nsamples, lo + per)


def first_min_index(costs: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Index (along dim 0) of the record with minimal cost; NaN is minimal; ties -> lowest idx.

    costs, idx: (R, B). Returns (B,) positions into the R records."""
    key = torch.where(torch.isnan(costs), torch.full_like(costs, -float("inf")), costs)
    best = key.min(dim=0, keepdim=True).values
    cand = torch.where(key == best, idx, torch.full_like(idx, torch.iinfo(idx.dtype).max))
    return cand.argmin(dim=0)


def merge_best(best_cost: torch.Tensor, best_idx: torch.Tensor, xs_star: torch.Tensor, us_star: torch.Tensor, group=None):
    """All-gather the per-rank winners and pick the global one on every rank.

    best_cost (B,), best_idx (B,) global sample ids, xs_star (B,N+1,nx), us_star (B,N,nu).
    Returns (xs_star, us_star, best_idx, best_cost) of the global winner, identical on all ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return xs_star, us_star, best_idx, best_cost
    R = dist.get_world_size(group)
    B = best_cost.shape[0]
    # the global sample id travels as its int32 BITS inside the float record (exact for every id; a float32 value would
    # collide above 2^24 samples), as the peer-memory kernel does (abr_xchg.cu)
    rec = torch.cat((best_cost.reshape(B, 1).float(), best_idx.reshape(B, 1).to(torch.int32).view(torch.float32),
                     us_star.reshape(B, -1).float(), xs_star.reshape(B, -1).float()), dim=1).contiguous()
    flat = torch.empty((R * B, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(flat, rec, group=group)  # concatenated layout works on nccl and gloo
    out = flat.reshape(R, B, rec.shape[1])
    costs = out[:, :, 0]
    idx = out[:, :, 1].contiguous().view(torch.int32).to(torch.int64)
    pick = first_min_index(costs, idx)  # (B,)
    sel = out[pick, torch.arange(B, device=rec.device)]  # (B, rec)
    nu_n = us_star[0].numel()
    return (sel[:, 2 + nu_n:].reshape(xs_star.shape).to(xs_star.dtype), sel[:, 2:2 + nu_n].reshape(us_star.shape).to(us_star.dtype),
            sel[:, 1].contiguous().view(torch.int32).to(best_idx.dtype), sel[:, 0].to(best_cost.dtype))


class PeerExchange:
    """The winner exchange as one kernel over NVLink peer memory (abr_xchg_*). Collective: build it on every
    rank of `group` (ranks = GPUs of one box), then pass it to `sharded_optimize(..., exchange=...)`.

    `capacity` bounds B * (2 + N*nu + (N+1)*nx) floats per call."""

    def __init__(self, device: torch.device, capacity: int, group=None):
        import ctypes as C

        from ambersim_b200 import _lib

        self._L = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self._h = C.c_void_p()
        if self.device.type != "cuda":
            raise ValueError("PeerExchange needs a CUDA device")
        dev_index = torch.cuda.current_device() if self.device.index is None else self.device.index
        self.device = torch.device("cuda", dev_index)
        handle = C.create_string_buffer(_lib.xchg_handle_bytes())
        _lib.check(self._L.abr_xchg_create(dev_index, self.world, self.rank, int(capacity), C.byref(self._h), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        _lib.check(self._L.abr_xchg_connect(self._h, b"".join(handles)))  # ABR_EINVAL when two ranks share one GPU
        self._group = group
        dist.barrier(group)

    def merge_best(self, best_cost, best_idx, xs_star, us_star, check: bool = True):
        """check=True reads the winner ids back (one small device->host sync) and raises `TimeoutError` when a peer
        failed to arrive (the kernel then returns the sentinel id -1 / cost +inf, never stale memory). Latency-critical
        loops pass check=False and poll `timed_out()` at their own pace."""
        import ctypes as C

        from ambersim_b200 import _lib

        B = best_cost.shape[0]
        cost, idx = best_cost.reshape(B).float().contiguous(), best_idx.reshape(B).to(torch.int32).contiguous()
        xs, us = xs_star.reshape(B, -1).float().contiguous(), us_star.reshape(B, -1).float().contiguous()
        xs_o, us_o, idx_o, cost_o = torch.empty_like(xs), torch.empty_like(us), torch.empty_like(idx), torch.empty_like(cost)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self._L.abr_xchg_merge_best_dev(self._h, p(cost), p(idx), p(xs), p(us), B, xs.shape[1], us.shape[1], p(xs_o), p(us_o),
                                                   p(idx_o), p(cost_o), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        if check and bool((idx_o < 0).any()):
            self.timed_out()  # clears the status word; the handle now refuses further exchanges
            raise TimeoutError("PeerExchange.merge_best: a peer did not publish its record within ~2 s; re-create the exchange on every rank")
        return xs_o.reshape(xs_star.shape).to(xs_star.dtype), us_o.reshape(us_star.shape).to(us_star.dtype), idx_o.to(best_idx.dtype), cost_o

    def timed_out(self) -> bool:
        import ctypes as C

        from ambersim_b200 import _lib

        torch.cuda.synchronize(self.device)
        flag = C.c_int()
        _lib.check(self._L.abr_xchg_timed_out(self._h, C.byref(flag)))
        return bool(flag.value)

    def close(self):
        if self._h:
            torch.cuda.synchronize(self.device)
            dist.barrier(self._group)  # nobody unmaps a buffer a peer may still be writing
            self._L.abr_xchg_destroy(self._h)
            self._h = None


def sharded_optimize(sampler, params, group=None, exchange: "PeerExchange" = None, check: bool = True, return_info: bool = False):
    """VanillaPredictiveSampler.optimize with `sampler.nsamples` samples split over the ranks.
    `exchange`: a `PeerExchange` for the winner exchange (default: all_gather through torch.distributed); `check` as in
    `PeerExchange.merge_best`. return_info adds {"best_idx", "best_cost"} of the global winner."""
    import dataclasses

    if not dist.is_initialized():
        return sampler.optimize(params, return_info=return_info)
    rank, R = dist.get_rank(group), dist.get_world_size(group)
    S = int(sampler.nsamples)
    lo, hi = shard_range(S, rank, R)
    local = dataclasses.replace(sampler, nsamples=max(1, hi - lo))
    ug = params.us_guess
    batched = ug.dim() == 3
    xs, us, info = local.optimize(params, sample_offset=min(lo, S - 1), nsamples_total=S, return_info=True)
    if not batched:
        xs, us = xs[None], us[None]
        info = {k: (v[None] if hasattr(v, "dim") else v) for k, v in info.items()}
    if exchange is not None:
        xs, us, idx, cost = exchange.merge_best(info["best_cost"].reshape(-1), info["best_idx"].reshape(-1), xs, us, check=check)
    else:
        xs, us, idx, cost = merge_best(info["best_cost"].reshape(-1), info["best_idx"].reshape(-1), xs, us, group)
    if not batched:
        xs, us, idx, cost = xs[0], us[0], idx[0], cost[0]
    if return_info:
        return xs, us, {"best_idx": idx, "best_cost": cost}
    return xs, us
