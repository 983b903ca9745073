"""N > 1 host logic on CPU: a world-size-2 gloo group shards a batch by world index, every rank steps its shard
(here with the oracle — test infrastructure — standing in for its GPU), no collective while stepping, one final
gather; the result must equal the unsharded run bit for bit (worlds are independent, src/Physics.elm:522-525)."""
import os
import socket

import numpy as np
import pytest

from elm_physics_b200 import abi, scenes, sharding

N_WORLDS, STEPS = 7, 25     # odd on purpose: ragged shards


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 8192, 65536):
        for ws in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            for k, (a, b) in enumerate(r):
                assert all(sharding.owner_of(w, n, ws) == k for w in range(a, b))
            if n >= ws:
                sizes = [b - a for a, b in r]
                assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, q):
    import torch.distributed as dist
    import helpers as H
    from oracle import binding as oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        w0, w1 = sharding.shard_range(N_WORLDS, rank, world_size)
        shapes, bodies, params = scenes.mixed_worlds(w1 - w0, bodies_per_world=8, first_world=w0)
        out = H.contacts_for(bodies)
        oracle.step(shapes, bodies, params, None, out, STEPS)
        full = sharding.gather_body_state(bodies, N_WORLDS)
        if rank == 0:
            q.put({k: v.tobytes() for k, v in full.items()})
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_the_unsharded_batch(oracle):
    import torch.multiprocessing as mp
    import helpers as H
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shapes, bodies, params = scenes.mixed_worlds(N_WORLDS, bodies_per_world=8)
    out = H.contacts_for(bodies)
    oracle.step(shapes, bodies, params, None, out, STEPS)
    for name in abi.BODY_INOUT:
        assert got[name] == getattr(bodies, name).tobytes(), name


def test_bench_end_to_end_driving_policy():
    """How bench.py drives a shard end to end (measured: profiles/r02_e2e_contexts_sweep.txt, r02_e2e_host_wait_probe.txt):
    one context per GPU up to 4096 worlds, two above; spinning host waits only with four cores per driving thread."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert [bench.e2e_contexts_for(w) for w in (1, 1024, 4096, 4097, 8192)] == [1, 1, 1, 2, 2]
    assert bench.e2e_contexts_for(8192, 3) == 3 and bench.e2e_contexts_for(2, 4) == 2 and bench.e2e_contexts_for(1, 2) == 1
    assert bench.host_waits_spin(8192, 1, 1, 16)            # one GPU, two threads, 16 cores
    assert bench.host_waits_spin(8192, 8, 8, 32)            # eight ranks x one context on 32 cores
    assert not bench.host_waits_spin(8192, 8, 8, 16)
    assert not bench.host_waits_spin(65536, 8, 8, 32)       # eight ranks x two contexts
