"""Multi-GPU path on CPU: frame ranges are split contiguously over ranks with no data-path collective.
world_size-2 gloo processes each run their range (through the oracle here - no GPU in this container) and
only exchange checksums to prove that the union is the whole clip, with no overlap."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pysangnom.shard import frame_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("total,world", [(0, 2), (1, 2), (7, 2), (256, 8), (257, 8), (5, 8), (1000, 3)])
def test_ranges_partition_the_clip(total, world):
    ranges = [frame_range(total, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
        assert e0 == b1 and b0 <= e0
    sizes = [e - b for b, e in ranges]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_bad_rank():
    with pytest.raises(ValueError):
        frame_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200")]
    from oracle import oracle as O
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fmt = FORMATS["YV12"]
    b, e = frame_range(total, rank, world)
    sums = torch.zeros(total, dtype=torch.int64)
    for n in range(b, e):
        out = O.oracle_frame(make_frame(9, 64, 32, fmt, "noise", n), 8, order=0, aa=48, aac=48, parity=(n % 2 == 0))
        h = hashlib.sha256(b"".join(p.tobytes() for p in out)).digest()
        sums[n] = int.from_bytes(h[:7], "little")
    dist.barrier()
    # control-plane only: gather per-frame checksums; frames a rank does not own stay 0
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    owned = torch.zeros(total, dtype=torch.int64)
    owned[b:e] = 1
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((sums.tolist(), owned.tolist()))
    dist.destroy_process_group()


def test_two_ranks_cover_the_clip_once():
    total, world = 9, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sums, owned = q.get()
    assert owned == [1] * total
    # the same frames on one rank
    sys.path[:0] = [ROOT]
    from oracle import oracle as O
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    for n in range(total):
        out = O.oracle_frame(make_frame(9, 64, 32, FORMATS["YV12"], "noise", n), 8, order=0, aa=48, aac=48, parity=(n % 2 == 0))
        h = hashlib.sha256(b"".join(p.tobytes() for p in out)).digest()
        assert sums[n] == int.from_bytes(h[:7], "little")


@pytest.mark.parametrize("nframes,chunk,ndev", [(0, 4, 2), (1, 4, 2), (37, 9, 2), (592, 37, 8), (100, 7, 3), (5, 100, 4)])
def test_chunks_of_a_batch_are_dealt_round_robin_over_the_devices(nframes, chunk, ndev):
    """The multi-device context (one pipeline per GPU behind one sn_ctx): every frame of every batch lands in exactly
    one chunk, chunks are runs of consecutive frames, and consecutive chunks - across batches too - go to consecutive
    devices. The dealing is the pure function the library's submit uses (csrc/sangnom_plan.h plan_chunks)."""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-f", os.path.join(here, "emul", "Makefile")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(here, "emul", "libkernel_emul.so"))
    L.emul_deal_chunks.restype = C.c_int
    L.emul_deal_chunks.argtypes = [C.c_longlong, C.c_longlong, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
    nxt = C.c_longlong(0)
    seen_devices = []
    for batch in range(3):
        out = (C.c_longlong * (3 * 1024))()
        n = L.emul_deal_chunks(nframes, chunk, ndev, C.byref(nxt), out, 1024)
        deals = [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(n)]
        assert n == (nframes + chunk - 1) // chunk
        covered = []
        for first, last, dev in deals:
            assert 0 <= dev < ndev and first < last <= nframes and last - first <= chunk
            covered += list(range(first, last))
            seen_devices.append(dev)
        assert covered == list(range(nframes))
    assert seen_devices == [i % ndev for i in range(len(seen_devices))]
