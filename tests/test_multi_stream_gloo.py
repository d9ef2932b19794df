"""N>1 path on CPU: world_size-2 gloo run of the stream fan-out plumbing bench.py uses
(independent streams per rank, barrier + MAX-over-ranks timing; no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videorenderer_b200 import streams


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w, l = streams.rank_info()
        mine = streams.assign_streams(8, w)[r]
        # every rank "renders" its own streams with the planning-only processor (host logic only, no GPU)
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        import harness as H
        import orclib as O
        plans = []
        for s in mine:
            case = H.Case(O.CF["P010"], 64 + 2 * s, 36, 128 + 4 * s, 72, H.EX_2020_PQ, up=O.UPSCALE["Lanczos3"])
            vp = case.make_processor(device=-1)
            plans.append((s, vp.debug_plan().taps_x, vp.GetVideoRect()))
        streams.barrier(dist)
        my_ms = 10.0 * (rank + 1)                      # rank 1 is the slow one
        ms = streams.max_over_ranks(my_ms, dist)
        fps = streams.aggregate_fps(100, w, ms)
        q.put((rank, mine, plans, ms, fps))
    finally:
        dist.destroy_process_group()


def test_two_rank_stream_fanout_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]        # disjoint, covering
    assert res[0][3] == res[1][3] == 20.0                                  # MAX over ranks
    assert res[0][4] == res[1][4] == pytest.approx(2 * 100 / 0.020)
    for _, _, plans, _, _ in res:
        for s, taps, rect in plans:
            assert taps == 6 and rect == (0, 0, 128 + 4 * s, 72)


def test_assign_streams_shapes():
    assert streams.assign_streams(8, 8) == [[i] for i in range(8)]
    assert streams.assign_streams(8, 1) == [list(range(8))]
    assert streams.assign_streams(3, 2) == [[0, 2], [1]]
    assert streams.max_over_ranks(3.5, None) == 3.5
