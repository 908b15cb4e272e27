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


class PeerHop:
    """The same hop ON THE DEVICE (csrc/hop.cu): the producer stage writes the hidden state straight into the consumer's input slot over NVLink
    (cudaIpc-mapped peer memory) and releases a sequence word; the consumer's stream waits for it with a 1-CTA kernel in front of its decode
    step and acks the slot afterwards.  Everything is stream-ordered kernels, so a stage's whole step (wait -> copy-in -> stage -> ack -> send) can
    be captured into ONE CUDA graph per slot and replayed without host code between two stages' kernels.  `ops` is the ctypes mirror of the
    C-ABI, `dist` torch.distributed (only used once, to exchange the 64-byte IPC handles)."""

    SLOT_ALIGN = 256

    def __init__(self, pipe: Pipeline, n_embd: int, ops, torch, dist, device, n_slots: int = 2):
        import ctypes as C
        self.pipe, self.ops, self.torch, self.C, self.n, self.n_slots = pipe, ops, torch, C, n_embd, n_slots
        L = ops.lib()
        self.slot_bytes = (n_embd * 4 + self.SLOT_ALIGN - 1) // self.SLOT_ALIGN * self.SLOT_ALIGN
        inbox_bytes = n_slots * (self.slot_bytes + self.SLOT_ALIGN)        # data slots, then one ready word per 256-byte line
        ack_bytes = n_slots * self.SLOT_ALIGN
        self.inbox, self.ackbox = C.c_void_p(), C.c_void_p()
        h_in, h_ack = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
        ops.check(L.b200_ipc_alloc(inbox_bytes, C.byref(self.inbox), h_in))
        ops.check(L.b200_ipc_alloc(ack_bytes, C.byref(self.ackbox), h_ack))
        mine = torch.tensor(list(bytes(h_in)) + list(bytes(h_ack)), dtype=torch.uint8, device=device)
        allh = [torch.zeros_like(mine) for _ in range(pipe.world)]
        dist.all_gather(allh, mine)
        self.peer_inbox, self.peer_ack = C.c_void_p(), C.c_void_p()          # next stage's inbox, previous stage's ack words
        if pipe.is_active and pipe.next is not None:
            hb = (C.c_uint8 * 64)(*allh[pipe.next][:64].cpu().tolist())
            ops.check(L.b200_ipc_open(hb, C.byref(self.peer_inbox)))
        if pipe.is_active and pipe.prev is not None:
            hb = (C.c_uint8 * 64)(*allh[pipe.prev][64:].cpu().tolist())
            ops.check(L.b200_ipc_open(hb, C.byref(self.peer_ack)))
        dist.barrier()
        self.state = torch.zeros((3, n_slots, 2), dtype=torch.int32, device=device)      # [wait | ack | send][slot] -> (count, error)

    # addresses inside a mapped inbox / ack box
    def _data(self, base, slot): return base.value + slot * self.slot_bytes
    def _ready(self, base, slot): return base.value + self.n_slots * self.slot_bytes + slot * self.SLOT_ALIGN
    def _ack(self, base, slot): return base.value + slot * self.SLOT_ALIGN
    def _state(self, which, slot): return self.state[which, slot].data_ptr()

    def enqueue_recv(self, slot: int, x_in) -> int:
        """wait for the producer's release of `slot`, then bring the hidden state into the stage's input vector.  Returns #launches."""
        if not self.pipe.is_active or self.pipe.prev is None:
            return 0
        C, ops, L = self.C, self.ops, self.ops.lib()
        ops.check(L.b200_hop_wait(C.c_void_p(self._ready(self.inbox, slot)), C.c_void_p(self._state(0, slot)), ops.stream()))
        src = ops.Tensor(); dst = ops.T(x_in)
        src.data, src.type, src.layout = self._data(self.inbox, slot), ops.F32, ops.LAYOUT_NATIVE
        for i in range(4):
            src.ne[i], src.nb[i] = dst.ne[i], dst.nb[i]
        ops.check(L.b200_cpy(C.byref(src), C.byref(dst), ops.stream()))
        return 2

    def enqueue_send(self, slot: int, x_out) -> int:
        """after the stage's step: ack the slot just consumed (to the previous stage) and publish x_out into the next stage's slot."""
        if not self.pipe.is_active:
            return 0
        C, ops, L = self.C, self.ops, self.ops.lib()
        n = 0
        if self.pipe.prev is not None:
            ops.check(L.b200_hop_ack(C.c_void_p(self._ack(self.peer_ack, slot)), C.c_void_p(self._state(1, slot)), ops.stream()))
            n += 1
        if self.pipe.next is not None:
            ops.check(L.b200_hop_send(C.c_void_p(x_out.data_ptr()), C.c_void_p(self._data(self.peer_inbox, slot)), self.n,
                                      C.c_void_p(self._ready(self.peer_inbox, slot)), C.c_void_p(self._ack(self.ackbox, slot)),
                                      C.c_void_p(self._state(2, slot)), ops.stream()))
            n += 1
        return n

    def errors(self) -> int:
        return int(self.state[:, :, 1].abs().sum().item())

    def close(self):
        L = self.ops.lib()
        for p in (self.peer_inbox, self.peer_ack):
            if p.value:
                L.b200_ipc_close(p)
        self.torch.cuda.synchronize()
        for p in (self.inbox, self.ackbox):
            if p.value:
                L.b200_ipc_free(p)
