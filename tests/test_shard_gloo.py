"""CPU tier, world_size 2 over gloo: the N>1 path is pure partitioning (no data-path collective),
so the test checks that two ranks agree on a disjoint, complete assignment of streams and of
key-frame-delimited segments, that segments decode independently (host parser -> oracle equals the
unsplit decode), and that the max-over-ranks reduction bench.py uses works."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ivf_path, q):
    sys.path.insert(0, helpers.ROOT)
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vp8_b200
    from vp8_b200 import shard
    # streams
    mine = shard.streams_of_rank(7, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # segments of one stream, decoded by this rank only
    _, payloads = vp8_b200.read_ivf(ivf_path)
    segs = shard.split_at_key_frames(payloads)
    md5s = {}
    for g, (first, frames) in shard.segments_of_rank(segs, rank, world):
        ps, orc = vp8_b200.Parser(), helpers.Oracle()
        out = []
        for p in frames:
            fr = ps.parse(p)
            img = orc.decode(fr)
            if fr.desc().hdr.show_frame:
                out.append(helpers.md5(img))
            fr.close()
        md5s[first] = out
        orc.close()
    all_md5 = [None] * world
    dist.all_gather_object(all_md5, md5s)
    # the timing reduction of bench.py: max over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((gathered, all_md5, float(t.item()), len(segs)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_streams_and_gop_segments(built, tmp_path):
    ivf = helpers.synth_stream("--width 176 --height 144 --frames 18 --seed 21 --key-interval 6 --hidden-altref 0")
    path = tmp_path / "gop.ivf"
    path.write_bytes(ivf)
    whole = [helpers.md5(f) for f in helpers.oracle_decode_ivf(ivf)]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(path), q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, all_md5, tmax, n_segs = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(7)) and not set(gathered[0]) & set(gathered[1])
    assert n_segs == 3
    merged = {}
    for d in all_md5:
        assert not set(d) & set(merged)
        merged.update(d)
    stitched = [m for first in sorted(merged) for m in merged[first]]
    assert stitched == whole  # segments decoded on different ranks == the unsplit decode
    assert tmax == 11.0
