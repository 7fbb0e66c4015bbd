"""N>1 host logic on CPU: two gloo ranks shard a frame's blocks, compress their shards with the oracle,
and the gathered records form one valid frame whose marks agree with the single-writer frame."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.oracle import Port
    from oracle import frame_oracle as F
    from plz4_b200.shard import record_offsets, shard_blocks
    from tests.datagen import logtext
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = Port()
    bsz, nblk = 65536, 11
    data = logtext(bsz * nblk - 1000)
    b0, b1 = shard_blocks(nblk, world)[rank]
    recs = [port.block_record(data[b * bsz:(b + 1) * bsz], bsz, True) for b in range(b0, b1)]
    lens = [len(r) for r in recs]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lens, b"".join(recs)))        # frame assembly only; no data-path collective
    # timing rule of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        hdr = F.write_header(F.Opts(block_idx=4, block_checksum=True, content_checksum=False), port.xxh32)
        frame = hdr + b"".join(g[1] for g in gathered) + b"\0\0\0\0"
        marks = []
        ref = F.write_frame(data, F.Opts(block_idx=4, block_checksum=True, content_checksum=False), port,
                            progress=lambda s, d: marks.append(d))
        offs = [o for per in record_offsets([g[0] for g in gathered], len(hdr)) for o in per]
        q.put((frame == ref, F.read_frames(frame, port) == data, offs == marks[:-1], float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_builds_the_same_frame():
    from plz4_b200.shard import shard_blocks
    assert shard_blocks(11, 2) == [(0, 6), (6, 11)]
    assert shard_blocks(3, 8)[:4] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert shard_blocks(0, 4) == [(0, 0)] * 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, decodes, marks_ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and decodes and marks_ok and tmax == 2.0
