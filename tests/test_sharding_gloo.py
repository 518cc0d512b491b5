"""N > 1 host logic on CPU: world_size-2 gloo processes shard dialogue turns, decode them (the oracle stands in
for the CUDA codec, same call signature) and gather the waveform chunks to rank 0 in turn order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fireredtts2_b200.config import TINY  # noqa: E402
from fireredtts2_b200.sharding import (decode_sharded, dialogue_turn_lengths, make_batches,  # noqa: E402
                                       partition_units)
from fireredtts2_b200.weights import synthetic_state_dict  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _units(n, seed=0):
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in rng.integers(2, 9, size=n)]
    return [torch.from_numpy(rng.integers(0, TINY.codebook_size, size=(TINY.num_quantizers, L))) for L in lens]


def _oracle_decode_fn(sd):
    from oracle import codec_oracle as O

    def fn(tokens, lengths):
        tok = tokens.numpy()
        out = np.zeros((tok.shape[0], TINY.samples_per_token * tok.shape[2]), dtype=np.float32)
        for b in range(tok.shape[0]):   # item b == standalone decode of its first lengths[b] tokens
            L = int(lengths[b]) if lengths is not None else tok.shape[2]
            y = O.decode(sd, tok[b:b + 1, :, :L], TINY.num_heads, TINY.hop_length)
            out[b, :y.shape[1]] = y[0]
        return torch.from_numpy(out)
    return fn


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        sd = synthetic_state_dict(TINY, 0)
        units = _units(7)
        res = decode_sharded(_oracle_decode_fn(sd), units, torch.device("cpu"), max_batch=3, max_tokens=20)
        if rank == 0:
            q.put([r.numpy().copy() for r in res])
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process result for the same units
    from oracle import codec_oracle as O
    sd = synthetic_state_dict(TINY, 0)
    units = _units(7)
    assert len(got) == len(units)
    for u, g in zip(units, got):
        ref = O.decode(sd, u.numpy()[None], TINY.num_heads, TINY.hop_length)[0]
        assert g.shape == ref.shape
        assert np.abs(g - ref).max() < 2e-6


def test_partition_is_balanced_and_complete():
    lens = dialogue_turn_lengths()
    assert sum(lens) == 2250 and len(lens) == 24 and max(lens) <= 375 and min(lens) >= 1
    for world in (1, 2, 4, 8):
        plan = partition_units(lens, world)
        flat = sorted(i for p in plan for i in p)
        assert flat == list(range(24))
        loads = [sum(lens[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(lens)     # LPT bound
    assert partition_units([], 4) == [[], [], [], []]
    assert partition_units([5], 2) == [[0], []]


def test_batches_respect_limits():
    lens = [9, 3, 7, 7, 2, 8, 1]
    bs = make_batches(range(len(lens)), lens, max_batch=3, max_tokens=20)
    assert sorted(i for b in bs for i in b) == list(range(len(lens)))
    for b in bs:
        assert len(b) <= 3
        assert len(b) * max(lens[i] for i in b) <= 20 or len(b) == 1


def test_single_process_path_without_dist():
    sd = synthetic_state_dict(TINY, 0)
    units = _units(3, seed=1)
    res = decode_sharded(_oracle_decode_fn(sd), units, torch.device("cpu"))
    assert [r.shape[0] for r in res] == [TINY.samples_per_token * u.shape[1] for u in units]


def test_unit_offsets_is_the_concatenation_layout():
    from fireredtts2_b200.sharding import unit_offsets
    lens = dialogue_turn_lengths()
    offs = unit_offsets(lens)
    assert offs[0] == 0 and offs[-1] == 1920 * 2250 and len(offs) == len(lens) + 1
    assert all(offs[i + 1] - offs[i] == 1920 * lens[i] for i in range(len(lens)))


def test_partition_and_batches_properties():
    """Property tests of the host-side planning (hypothesis): every unit is owned exactly once, the plan is deterministic
    (every rank computes it without communicating), the greedy bound holds (max load <= mean load + longest unit), batches
    respect both limits, keep every unit, and are ordered longest first."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from fireredtts2_b200.sharding import unit_offsets

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(1, 375), min_size=0, max_size=60), st.integers(1, 8))
    def plan_props(lens, world):
        plan = partition_units(lens, world)
        assert plan == partition_units(list(lens), world) and len(plan) == world
        assert sorted(i for p in plan for i in p) == list(range(len(lens)))
        assert all(p == sorted(p) for p in plan)
        loads = [sum(lens[i] for i in p) for p in plan]
        if lens:
            assert max(loads) <= sum(lens) / world + max(lens)
            assert max(loads) - min(loads) <= max(lens)
        offs = unit_offsets(lens, 8)
        assert len(offs) == len(lens) + 1 and offs[0] == 0 and offs[-1] == 8 * sum(lens)
        assert all(b - a == 8 * n for a, b, n in zip(offs, offs[1:], lens))

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(1, 375), min_size=1, max_size=60), st.integers(1, 64), st.integers(1, 64 * 375))
    def batch_props(lens, max_batch, max_tokens):
        bs = make_batches(range(len(lens)), lens, max_batch, max_tokens)
        assert sorted(i for b in bs for i in b) == list(range(len(lens)))
        flat = [lens[i] for b in bs for i in b]
        assert flat == sorted(flat, reverse=True)                       # longest first, across and inside the batches
        for b in bs:
            padded = len(b) * max(lens[i] for i in b)
            assert 1 <= len(b) <= max_batch and (padded <= max_tokens or len(b) == 1)

    plan_props()
    batch_props()
