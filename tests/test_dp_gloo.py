"""Data-parallel TD3 step, host-side logic, on CPU with 2 `gloo` ranks (no GPU needed).

What the product does at N>1 (sgrl_b200/agent.py `_allreduce` + `FusedAdam.step(world_size=)`): every rank computes the gradients of
its LOCAL mean loss, the flat gradient arena is all-reduced with SUM, and the fused clip+Adam kernel multiplies by
grad_scale = 1/world before clipping.  The claim under test: that equals the single-device step on the
concatenated batch.  Gradients are produced by the oracle here (the CUDA kernels are checked against the oracle in the -m gpu tests);
the reduction goes through the agent's own `_world` / `_allreduce` methods.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_grads(pc, batch, g, lo, hi):
    from oracle import set_oracle as O
    p = {k: v.clone().requires_grad_(not O.is_dead(k)) for k, v in pc.items()}
    q1, q2 = O.critic_forward(p, batch["obs"][lo:hi], batch["action"][lo:hi], g)
    tgt = batch["reward"][lo:hi].expand_as(q1)            # any fixed target: the test is about the reduction
    loss = ((q1 - tgt) ** 2).mean() + ((q2 - tgt) ** 2).mean()
    loss.backward()
    names = [k for k in p if p[k].grad is not None]
    return names, torch.cat([p[k].grad.reshape(-1) for k in names])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ref_loader, set_oracle as O
        from sgrl_b200 import graph as G, morphologies as M, synth
        from sgrl_b200.agent import Agent
        par = M.ALL["3d_hopper_3_shin"] if "3d_hopper_3_shin" in M.ALL else next(iter(M.ALL.values()))
        g = G.build_graph(par)
        B = 8
        batch = synth.make_batch(B, len(par), seed=5)
        pc = {"critic1." + k: v for k, v in O.synth_params("critic", 12).items()}
        pc.update({"critic2." + k: v for k, v in O.synth_params("critic", 13).items()})
        per = B // world
        names, flat = _shard_grads(pc, batch, g, rank * per, (rank + 1) * per)
        agent = Agent(ref_loader.default_args())          # CPU construction only: no kernels are launched
        assert agent._world() == world
        agent._allreduce(flat, agent._world())             # SUM over ranks, as before the fused clip+Adam
        flat *= 1.0 / world                                # grad_scale of sgrl_adam_clip
        _, full = _shard_grads(pc, batch, g, 0, B)
        rel = ((flat - full).norm() / full.norm()).item()
        # replicas identical after the reduction
        chk = flat.double().sum().reshape(1).clone()
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        if rank == 0:
            out.put((rel, float((gathered[0] - gathered[1]).abs().item()), len(names)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gradient_allreduce_equals_concatenated_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(540)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    rel, spread, n = out.get(timeout=10)
    assert n > 100
    assert rel < 1e-5, rel          # fp32 mean-of-means vs full mean (equal shard sizes)
    assert spread == 0.0            # every rank holds the identical reduced gradient
