"""CPU tests of the multi-GPU host logic (N > 1 path): frame sharding + the single counter all-reduce,
world_size 2 over gloo.  The per-rank decode is stood in for by the oracle here (no GPU in this container);
on the GPU box tests/test_gpu_parity.py::test_sharded_equals_whole runs the same split through the kernel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_frames_partitions_exactly():
    from iib_project_ldpc_codes_b200.dist import shard_frames
    for total in (0, 1, 127, 128, 129, 1000, 10 ** 7, 10 ** 9 + 7):
        for world in (1, 2, 3, 4, 8):
            pos = 256
            for r in range(world):
                lo, cnt = shard_frames(total, r, world, first_frame=256)
                assert lo % 128 == 0 or cnt == 0
                assert lo == pos or cnt == 0
                pos = lo + cnt if cnt else pos
            assert pos == 256 + total
    with pytest.raises(ValueError):
        shard_frames(10, 0, 2, first_frame=5)


def test_shard_items_covers_every_point_and_frame_once():
    """(point, block) partitioning: every (point, frame) belongs to exactly one rank, blocks 128-aligned."""
    from iib_project_ldpc_codes_b200.dist import shard_items
    for npoints in (1, 2, 6, 20):
        for total in (1, 127, 1000, 12345, 10 ** 7):
            for world in (1, 2, 3, 4, 8):
                cover = {p: [] for p in range(npoints)}
                counts = []
                for r in range(world):
                    items = shard_items(npoints, total, r, world, first_frame=384)
                    counts.append(sum(c for _, _, c in items))
                    for p, lo, cnt in items:
                        assert lo % 128 == 0 and cnt > 0
                        cover[p].append((lo, cnt))
                for p, segs in cover.items():
                    pos = 384
                    for lo, cnt in sorted(segs):
                        assert lo == pos
                        pos += cnt
                    assert pos == 384 + total
                if total >= 10 ** 7:
                    assert max(counts) - min(counts) <= 128 * npoints + total // 4   # dealt evenly up to one block


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    from iib_project_ldpc_codes_b200 import dist as D
    from oracle import oracle as O
    from golden_util import mc_cases
    r, w, _ = D.init_from_env(backend="gloo")
    outs = []
    for c in list(mc_cases())[:3]:
        lo, cnt = D.shard_frames(c["frames"], r, w, c["first"])
        part = O.restatement().mc_run(c["v2c"], c["c2v"], c["n"], c["k"], c["dv"], c["dc"], c["eps"], c["seed"], lo, cnt,
                                      c["its"], c["expurgation"], nthreads=1)
        t = torch.from_numpy(part.view(np.int64).copy())
        D.allreduce_counters(t)
        outs.append(t.numpy().view(np.uint64))
    if r == 0:
        q.put(outs)
    torch.distributed.destroy_process_group()


def test_world_size_2_gloo_allreduce_reproduces_single_process_counters(oracle):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import mc_cases
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for got, c in zip(outs, list(mc_cases())[:3]):
        assert np.array_equal(got, c["counters"])   # golden = reference C over the whole frame range
