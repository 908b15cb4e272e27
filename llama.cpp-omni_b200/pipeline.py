"""Layer-split pipeline host logic: which contiguous layer range a rank owns, and the one hidden-state hop per stage boundary.

Mirrors the reference's LLAMA_SPLIT_MODE_LAYER (src/llama-model.cpp:2130-2185: contiguous layer ranges per device, the output layer with the
last one; KV cache of layer il lives with layer il, src/llama-context.cpp:313) for the one-process-per-GPU launch bench.py uses.  The only
exchange on the decode path is the residual stream `l_out` ([n_embd] F32 per token, 16 KiB) between consecutive stages: a point-to-point
send/recv (NCCL over NVLink on GPUs, gloo in the CPU tests).  There is no all-reduce on this path and none is invented.

Unlike the reference (equal layer counts scaled by free memory), stages are balanced by a per-layer COST the caller supplies, lm_head
included.  bench.py uses streamed bytes + the fixed latency of a layer's dependent phase transitions (in bytes of streaming time): by bytes
alone lm_head (0.51 GB of Q6_K) weighs 4.4 transformer layers, by measured time 1.5 — balancing by the latter took 2 GPUs from 771 to 877 tok/s.
"""
from __future__ import annotations

from typing import Callable, Sequence


def partition_layers(layer_bytes: Sequence[int], head_bytes: int, world: int) -> list[range]:
    """Contiguous ranges, one per rank (possibly empty at the tail when world > n_layer + 1), minimising the largest stage; the last NON-EMPTY
    stage also carries `head_bytes`.  Exact dynamic programme over prefix sums (n_layer * world states)."""
    n = len(layer_bytes)
    if world <= 0:
        raise ValueError("world must be positive")
    pre = [0]
    for b in layer_bytes:
        pre.append(pre[-1] + int(b))
    INF = float("inf")
    # best[s][i]: minimal max-stage cost splitting layers [0, i) into s stages, none of which carries the head
    best = [[INF] * (n + 1) for _ in range(world + 1)]
    cut = [[0] * (n + 1) for _ in range(world + 1)]
    best[0][0] = 0
    for s in range(1, world + 1):
        for i in range(0, n + 1):
            for j in range(0, i + 1):
                if best[s - 1][j] == INF:
                    continue
                c = max(best[s - 1][j], pre[i] - pre[j])
                if c < best[s][i]:
                    best[s][i], cut[s][i] = c, j
    # the last stage takes layers [j, n) + head; earlier s-1 stages split [0, j)
    choice, cost = (1, 0), INF
    for s in range(1, world + 1):
        for j in range(0, n + 1):
            if best[s - 1][j] == INF:
                continue
            c = max(best[s - 1][j], pre[n] - pre[j] + head_bytes)
            if c < cost or (c == cost and s > choice[0]):
                cost, choice = c, (s, j)
    s, j = choice
    bounds = [n, j]
    for t in range(s - 1, 0, -1):
        j = cut[t][j]
        bounds.append(j)
    bounds.reverse()                                     # [0, ..., n]
    ranges = [range(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
    ranges += [range(n, n)] * (world - len(ranges))
    return ranges


def head_rank(ranges: Sequence[range]) -> int:
    """The stage that applies output_norm + lm_head: the last one that owns layers (or rank 0 for a model without layers)."""
    owners = [r for r, rg in enumerate(ranges) if len(rg) > 0]
    return owners[-1] if owners else 0


class Pipeline:
    """One hop per boundary: stage r receives the hidden state from r-1, runs its layers, sends to r+1.  `dist` is torch.distributed (or a
    stand-in with send/recv); with world == 1 every call is a no-op."""

    def __init__(self, rank: int, world: int, ranges: Sequence[range], dist=None):
        self.rank, self.world, self.ranges, self.dist = rank, world, list(ranges), dist
        self.active = [r for r, rg in enumerate(ranges) if len(rg) > 0] or [0]
        self.first, self.last = self.active[0], self.active[-1]
        self.is_active = rank in self.active
        i = self.active.index(rank) if self.is_active else -1
        self.prev = self.active[i - 1] if i > 0 else None
        self.next = self.active[i + 1] if 0 <= i < len(self.active) - 1 else None

    def stage_step(self, hidden_in, hidden_out, stage_fn: Callable[[], None]) -> None:
        """recv (unless first) -> stage_fn() -> send (unless last).  stage_fn reads hidden_in and writes hidden_out in place."""
        if not self.is_active:
            return
        if self.prev is not None:
            self.dist.recv(hidden_in, src=self.prev)
        stage_fn()
        if self.next is not None:
            self.dist.send(hidden_out, dst=self.next)
