"""Data parallelism on real NCCL hardware (SURVEY.md §4 tier 5; needs >= 2 GPUs, skipped otherwise — run with
`gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py -m gpu`).  Two ranks, one process each:

* one data-parallel TD3 step (each rank its half of the batch, gradient arena all-reduced inside the step) equals the
  single-GPU step on the concatenated batch: gradients within 1e-4, in eager mode and as a replayed CUDA graph;
* after 10 graph-replayed steps on different per-rank batches the replicas' parameter arenas (live and target nets) are
  bit-identical.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from sgrl_b200 import graph as G, morphologies as M, synth
        import parity
        from test_agent_gpu import make_agent
        par = M.ALL["3d_humanoid_9_full"]
        g = G.build_graph(par, device=dev)
        B, N = 64, len(par)
        full = {k: v.to(dev) for k, v in synth.make_batch(world * B, N, seed=9).items()}
        noise = (torch.randn(world * B, 3 * N, generator=torch.Generator().manual_seed(5)) * 0.2).to(dev)
        mine = {k: v[rank * B:(rank + 1) * B].contiguous() for k, v in full.items()}
        res = {}
        for mode in ("eager", "graph"):
            dp, _, _ = make_agent()
            one, _, _ = make_agent()
            one.data_parallel = False
            for ag in (dp, one):
                ag.change_morphology(g)
                ag.use_graphs = mode == "graph"
            reps = 1 if mode == "eager" else 3         # graph mode: the third update of each kind is a replay; state re-synchronised before it
            for r in range(reps):
                if r:
                    for ag in (dp, one):
                        fresh, _, _ = make_agent()
                        ag.load_state_dict(fresh.state_dict())
                        ag.critic_optimizer.load_state_dict(fresh.critic_optimizer.state_dict())
                        ag.actor_optimizer.load_state_dict(fresh.actor_optimizer.state_dict())
                dp.update(mine, 0, noise=noise[rank * B:(rank + 1) * B])
                one.update(full, 0, noise=noise)
            torch.cuda.synchronize()
            # the DP arena holds the SUM over ranks of the local mean-loss gradients; 1/world is folded into the fused Adam
            ec = parity.rel_err(dp.critic.grad_arena() / world, one.critic.grad_arena())
            ea = parity.rel_err(dp.actor.grad_arena() / world, one.actor.grad_arena())
            ep = parity.rel_err(dp.critic.live_arena - make_agent()[0].critic.live_arena, one.critic.live_arena - make_agent()[0].critic.live_arena)
            res[mode] = (ec, ea, ep)
        # replicas after 10 replayed steps on rank-specific batches
        ag, _, _ = make_agent()
        ag.change_morphology(g)
        for it in range(10):
            b = {k: v.to(dev) for k, v in synth.make_batch(B, N, seed=1000 + 31 * rank + it).items()}
            ag.update(b, it)
        torch.cuda.synchronize()
        sums = []
        for m in (ag.actor, ag.critic, ag.actor_target, ag.critic_target):
            bits = m.full_arena.view(torch.int32).to(torch.int64)
            sums += [bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()]
        mine_sum = torch.stack(sums)
        allv = [torch.zeros_like(mine_sum) for _ in range(world)]
        dist.all_gather(allv, mine_sum)
        same = all(bool((v == allv[0]).all()) for v in allv)
        moved = parity.rel_err(ag.critic.live_arena, make_agent()[0].critic.live_arena)
        if rank == 0:
            out.put((res, same, moved))
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        os._exit(0)         # captured graphs hold NCCL work: leave without tearing the process group down (see bench.py)


@pytest.mark.timeout(900)
def test_two_rank_nccl_step_equals_concatenated_batch_and_replicas_stay_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res, same, moved = out.get(timeout=800)
    for p in procs:
        p.join(60)
    for mode, (ec, ea, ep) in res.items():
        assert ec < 1e-4 and ea < 1e-4, (mode, ec, ea)          # all-reduced gradients == gradients of the concatenated batch
        assert ep < 2e-3, (mode, ep)                             # and so is the clipped Adam step
    assert same, "replicas diverged"
    assert moved > 1e-6                                          # the 10 steps did change the weights
