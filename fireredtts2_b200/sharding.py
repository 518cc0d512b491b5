"""Multi-GPU partitioning of the codec decode path.

Units (utterances, dialogue turns) are independent — a turn's decode depends only on its own tokens, the LLM
carries the cross-turn context (reference fireredtts2.py:379-394) — so the path shards with no collective in
the compute: one process per GPU, a full weight replica each, a static longest-first assignment of units to
ranks.  The only exchange step is the gather of waveform chunks to one rank in unit order (config 4: the turns
of a dialogue are concatenated on the time axis, reference fireredtts2.py:401), done with point-to-point
``torch.distributed`` send/recv (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

SAMPLES_PER_TOKEN = 1920


def partition_units(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first greedy assignment.  Returns, per rank, the unit indices it owns
    (ascending).  Deterministic, so every rank computes the same plan without communicating."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(lengths[i])
    for p in plan:
        p.sort()
    return plan


def make_batches(unit_ids: Sequence[int], lengths: Sequence[int], max_batch: int, max_tokens: int) -> List[List[int]]:
    """Group a rank's units into padded batches of similar length (longest first) bounded by `max_batch`
    items and `max_tokens` padded tokens."""
    ids = sorted(unit_ids, key=lambda i: (-int(lengths[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    for i in ids:
        if cur:
            L = int(lengths[cur[0]])
            if len(cur) + 1 > max_batch or (len(cur) + 1) * L > max_tokens:
                batches.append(cur)
                cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


DecodeFn = Callable[[torch.Tensor, Optional[torch.Tensor]], torch.Tensor]


def decode_units_local(decode_fn: DecodeFn, units: Sequence[torch.Tensor], unit_ids: Sequence[int],
                       device: torch.device, max_batch: int = 64, max_tokens: int = 64 * 375,
                       samples_per_token: int = SAMPLES_PER_TOKEN) -> dict:
    """Decode the given units (each a (nq, L_i) integer tensor) in padded var-len batches.
    Returns {unit_id: waveform (samples_per_token*L_i,) on `device`}."""
    lengths = [int(u.shape[1]) for u in units]
    out = {}
    for batch in make_batches(unit_ids, lengths, max_batch, max_tokens):
        L = max(lengths[i] for i in batch)
        nq = units[batch[0]].shape[0]
        tok = torch.zeros((len(batch), nq, L), dtype=units[batch[0]].dtype, device=device)
        for k, i in enumerate(batch):
            tok[k, :, :lengths[i]] = units[i].to(device)
        lens = torch.tensor([lengths[i] for i in batch], dtype=torch.int32, device=device)
        audio = decode_fn(tok, lens)
        for k, i in enumerate(batch):
            out[i] = audio[k, :samples_per_token * lengths[i]]
    return out


def decode_sharded(decode_fn: DecodeFn, units: Sequence[torch.Tensor], device: torch.device,
                   group: Optional[dist.ProcessGroup] = None, dst: int = 0, max_batch: int = 64,
                   max_tokens: int = 64 * 375, samples_per_token: int = SAMPLES_PER_TOKEN,
                   gather: bool = True) -> Optional[List[torch.Tensor]]:
    """Decode `units` (the same list on every rank) sharded over the ranks of `group`.

    With ``gather`` the waveforms are collected on rank `dst` and returned there in unit order (other ranks
    return None); without it every rank returns only its own units (None elsewhere in the list)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = [int(u.shape[1]) for u in units]
    plan = partition_units(lengths, world)
    mine = decode_units_local(decode_fn, units, plan[rank], device, max_batch, max_tokens, samples_per_token)
    if world == 1 or not gather:
        return [mine.get(i) for i in range(len(units))]
    # ---- the one exchange step: waveform chunks to `dst`, one flat message per rank ----
    if rank != dst:
        if plan[rank]:
            flat = torch.cat([mine[i].reshape(-1) for i in plan[rank]]).contiguous()
            dist.send(flat, dst=dst, group=group)
        return None
    result: List[Optional[torch.Tensor]] = [None] * len(units)
    for i in plan[dst]:
        result[i] = mine[i]
    for r in range(world):
        if r == dst or not plan[r]:
            continue
        n = sum(samples_per_token * lengths[i] for i in plan[r])
        buf = torch.empty((n,), dtype=torch.float32, device=device)
        dist.recv(buf, src=r, group=group)
        off = 0
        for i in plan[r]:
            k = samples_per_token * lengths[i]
            result[i] = buf[off:off + k]
            off += k
    return result  # type: ignore[return-value]


def unit_offsets(lengths: Sequence[int], samples_per_token: int = SAMPLES_PER_TOKEN) -> List[int]:
    """Sample offset of every unit in the concatenation of all units in unit order (the gathered buffer's layout);
    one extra entry = the total."""
    off = [0]
    for n in lengths:
        off.append(off[-1] + samples_per_token * int(n))
    return off


class PeerBuffer:
    """A waveform buffer on rank ``dst``'s GPU that every rank of a one-box group can write with plain kernel stores.

    ``dst`` allocates it in the native library (frt2_peer_alloc: cudaMalloc + CUDA IPC handle), the 64-byte handle is
    broadcast through ``torch.distributed`` (plumbing), the other ranks map it (frt2_peer_open).  ``ptr`` is the device
    address to hand to ``RedCodecB200.decode_into`` — on ``dst`` the allocation itself, elsewhere the NVLink peer
    mapping.  ``tensor()`` (owner only) views the memory as a torch tensor without copying."""

    def __init__(self, numel: int, dtype: torch.dtype, device: torch.device, group=None, dst: int = 0):
        from . import _native as N
        self._N = N
        self._lib = N.load()
        self.numel, self.dtype, self.device, self.dst = int(numel), dtype, device, dst
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.owner = self.rank == dst
        self.ptr = 0
        itemsize = torch.empty((), dtype=dtype).element_size()
        handle = C.create_string_buffer(N.PEER_HANDLE_BYTES)
        p = C.c_void_p()
        err = None
        if self.owner:
            try:
                N.check(self._lib.frt2_peer_alloc(device.index, max(self.numel, 1) * itemsize, C.byref(p), handle))
                self.ptr = p.value
            except Exception as e:   # reported to every rank below: all ranks must take the same path
                err = repr(e)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            box = [(handle.raw, err)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
            raw, err = box[0]
            if err is None and not self.owner:
                try:
                    q = C.c_void_p()
                    N.check(self._lib.frt2_peer_open(device.index, raw, C.byref(q)))
                    self.ptr = q.value
                except Exception as e:
                    err = repr(e)
            errs = [None] * dist.get_world_size(group)
            dist.all_gather_object(errs, err, group=group)
            err = next((e for e in errs if e is not None), None)
        if err is not None:
            self.close()
            raise RuntimeError(f"peer memory unavailable: {err}")

    @property
    def __cuda_array_interface__(self):
        typestr = {torch.float32: "<f4", torch.int16: "<i2"}[self.dtype]
        return {"shape": (self.numel,), "typestr": typestr, "data": (self.ptr, False), "version": 2}

    def tensor(self) -> torch.Tensor:
        assert self.owner, "only the owning rank reads the gathered buffer"
        return torch.as_tensor(self, device=self.device)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if not self.ptr:
            return
        if self.owner:
            self._lib.frt2_peer_free(self.device.index, C.c_void_p(self.ptr))
        else:
            self._lib.frt2_peer_close(self.device.index, C.c_void_p(self.ptr))
        self.ptr = 0


def decode_sharded_peer(codec, units: Sequence[torch.Tensor], device: torch.device, group=None, dst: int = 0,
                        max_batch: int = 64, max_tokens: int = 64 * 375, pcm16: bool = False,
                        buffer: Optional[PeerBuffer] = None):
    """`decode_sharded` with the gather fused into the decode: every rank's overlap-add kernel writes its units
    straight to their place in ONE concatenated waveform on rank `dst` (peer stores over NVLink / NVSwitch,
    frt2_decode_scatter), so the exchange step overlaps the rest of the rank's batches and no collective runs after
    the compute.  Returns ``(flat, offsets, buffer)`` on `dst` — `flat` the (total samples,) waveform in unit order
    (== the dialogue, reference fireredtts2.py:399-401), `offsets` from `unit_offsets` — and ``(None, offsets,
    buffer)`` elsewhere.  The caller keeps `buffer` alive while it reads `flat` and closes it afterwards (or passes it
    back in for the next call).  `codec` is a RedCodecB200 (`decode_into`)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = [int(u.shape[1]) for u in units]
    spt = codec.cfg.samples_per_token
    offs = unit_offsets(lengths, spt)
    dtype = torch.int16 if pcm16 else torch.float32
    if buffer is None:
        buffer = PeerBuffer(offs[-1], dtype, device, group, dst)
    assert buffer.numel >= offs[-1] and buffer.dtype == dtype
    plan = partition_units(lengths, world)
    for batch in make_batches(plan[rank], lengths, max_batch, max_tokens):
        L = max(lengths[i] for i in batch)
        nq = units[batch[0]].shape[0]
        tok = torch.zeros((len(batch), nq, L), dtype=units[batch[0]].dtype, device=device)
        for k, i in enumerate(batch):
            tok[k, :, :lengths[i]] = units[i].to(device)
        lens = torch.tensor([lengths[i] for i in batch], dtype=torch.int32, device=device)
        out_off = torch.tensor([offs[i] for i in batch], dtype=torch.int64, device=device)
        codec.decode_into(tok, buffer.ptr, out_off, lens, pcm16=pcm16)
    torch.cuda.synchronize(device)      # this rank's peer stores are complete ...
    if world > 1:
        dist.barrier(group=group)       # ... and so are everybody else's
    return (buffer.tensor()[:offs[-1]] if rank == dst else None), offs, buffer


def dialogue_turn_lengths(total_tokens: int = 2250, turns: int = 24, seed: int = 0, max_len: int = 375) -> List[int]:
    """BASELINE.json configs[3]: a 3-minute 4-speaker dialogue = `turns` turns whose lengths are drawn
    U[2 s, 15 s] and rescaled to sum to `total_tokens` (each <= 30 s, reference fireredtts2.py:383)."""
    g = torch.Generator().manual_seed(seed)
    raw = 25 + torch.rand(turns, generator=g) * (187.5 - 25)
    lens = torch.clamp((raw * (total_tokens / raw.sum())).round().long(), 1, max_len)
    diff = total_tokens - int(lens.sum())
    i = 0
    while diff != 0:   # distribute the rounding remainder
        step = 1 if diff > 0 else -1
        if 1 <= int(lens[i % turns]) + step <= max_len:
            lens[i % turns] += step
            diff -= step
        i += 1
    return [int(x) for x in lens]
