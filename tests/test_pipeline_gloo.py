"""CPU: the N > 1 host logic (llama.cpp-omni_b200/pipeline.py) — byte-balanced contiguous layer partition and the one-hop-per-boundary
hidden-state hand-off — exercised with world_size 2 and 3 over the gloo backend (no GPU), as bench.py drives it over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from __graft_entry__ import load_package


def _pipeline():
    return load_package().pipeline


def test_partition_covers_every_layer_once_and_balances_bytes():
    P = _pipeline()
    dec = load_package().decode
    cfg = dec.LLMConfig()
    lb = [dec.weight_bytes_per_token(cfg, range(i, i + 1), with_head=False) for i in range(cfg.n_layer)]
    head = dec.weight_bytes_per_token(cfg, range(0), with_head=True)
    for world in (1, 2, 3, 4, 8):
        rg = P.partition_layers(lb, head, world)
        assert len(rg) == world
        flat = [i for r in rg for i in r]
        assert flat == list(range(cfg.n_layer))                      # contiguous, in order, each layer exactly once
        cost = [sum(lb[i] for i in r) + (head if k == P.head_rank(rg) else 0) for k, r in enumerate(rg)]
        total = sum(lb) + head
        assert max(cost) <= total / world + max(max(lb), head)       # within one indivisible unit of the ideal
        # never worse than the equal-count split the reference uses
        per = (cfg.n_layer + world - 1) // world
        naive = [sum(lb[i] for i in range(k * per, min(cfg.n_layer, (k + 1) * per))) for k in range(world)]
        naive[max(k for k in range(world) if k * per < cfg.n_layer)] += head
        assert max(cost) <= max(naive)
    # degenerate shapes
    assert [list(r) for r in P.partition_layers([5], 1, 3)][0:1] == [[0]] or sum(len(r) for r in P.partition_layers([5], 1, 3)) == 1
    assert sum(len(r) for r in P.partition_layers([], 7, 2)) == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_layer, n_steps, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = _pipeline()
    rng = np.random.default_rng(0)
    E = 64
    mats = [torch.from_numpy(rng.standard_normal((E, E)).astype(np.float32) / np.float32(8.0)) for _ in range(n_layer)]
    lb = [1000 + 37 * (i % 5) for i in range(n_layer)]
    rg = P.partition_layers(lb, 2500, world)
    pipe = P.Pipeline(rank, world, rg, dist)
    hidden = torch.zeros(E)
    outs = []
    for step in range(n_steps):
        if rank == pipe.first:
            hidden.copy_(torch.full((E,), 0.01 * (step + 1)))

        def stage():
            x = hidden.clone()
            for il in rg[rank]:
                x = torch.tanh(mats[il] @ x) + x
            hidden.copy_(x)
        pipe.stage_step(hidden, hidden, stage)
        if rank == pipe.last:
            outs.append(hidden.clone())
    if rank == pipe.last:
        torch.save(torch.stack(outs), out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pipeline_hand_off_matches_single_process(world, tmp_path):
    n_layer, n_steps, E = 7, 3, 64
    out = tmp_path / "out.pt"
    mp.spawn(_worker, args=(world, _free_port(), n_layer, n_steps, str(out)), nprocs=world, join=True)
    got = torch.load(out)
    rng = np.random.default_rng(0)
    mats = [torch.from_numpy(rng.standard_normal((E, E)).astype(np.float32) / np.float32(8.0)) for _ in range(n_layer)]
    for step in range(n_steps):
        x = torch.full((E,), 0.01 * (step + 1))
        for m in mats:
            x = torch.tanh(m @ x) + x
        assert torch.equal(got[step], x)              # same f32 ops in the same order on every path: bit-exact
