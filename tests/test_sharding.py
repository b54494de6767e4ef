"""CPU: the N>1 path's host logic — contiguous index sharding and the host-side gather of the
polygon lists (world_size 2, gloo backend, no GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from ocr_rs_b200.sharding import shard_range
    for n in (0, 1, 7, 16, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_document_shards_are_a_function_of_the_global_index():
    from ocr_rs_b200 import synth
    whole = synth.document_image_shard(0, 6, 128, 160, seed=3, unique=4)
    a = synth.document_image_shard(0, 3, 128, 160, seed=3, unique=4)
    b = synth.document_image_shard(3, 3, 128, 160, seed=3, unique=4)
    assert (np.concatenate([a, b]) == whole).all()
    assert not (whole[0] == whole[4]).all()  # second pass over the base set is shifted


def _fake_result(first, count):
    """Deterministic per-image polygons as a function of the global image index."""
    from ocr_rs_b200._ffi import Polygons
    io, po, xy, sc = [0], [0], [], []
    for i in range(first, first + count):
        for k in range(i % 3):
            m = 4 + (i + k) % 3
            xy.append(np.arange(2 * m, dtype=np.uint32).reshape(m, 2) + np.uint32(10 * i + k))
            sc.append(0.7 + 0.001 * i + 0.01 * k)
            po.append(po[-1] + m)
        io.append(len(po) - 1)
    return Polygons(arrays=(np.array(io, np.int64), np.array(po, np.int64),
                            np.concatenate(xy) if xy else np.zeros((0, 2), np.uint32), np.array(sc, np.float64),
                            np.arange(5 * first, 5 * (first + count), dtype=np.int64).reshape(count, 5)))


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ocr_rs_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_range(n, rank, world)
    mine = _fake_result(first, count)
    got = sharding.gather_polygons(mine)
    lists = sharding.gather_polygon_scores(mine.polygons, mine.scores)
    if rank == 0:
        q.put((got.arrays(), [[p.tolist() for p in img] for img in lists[0]]))
    else:
        assert got is None and lists is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    n, world, port = 11, 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    arrays, lists = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = _fake_result(0, n)
    for a, b in zip(arrays, whole.arrays()):
        assert a.shape == b.shape and (a == b).all()
    assert lists == [[p.tolist() for p in img] for img in whole.polygons]


def test_c_shard_range_matches_python():
    """ocrb_shard_range (the split ocrb_detect_and_recognize_sharded uses inside the library) == sharding.shard_range."""
    import ctypes as C

    from ocr_rs_b200 import _ffi
    from ocr_rs_b200.sharding import shard_range
    L = _ffi.lib()
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            for r in range(world):
                f, c = C.c_int64(), C.c_int64()
                assert L.ocrb_shard_range(n, r, world, C.byref(f), C.byref(c)) == 0
                assert (f.value, c.value) == shard_range(n, r, world)
    assert L.ocrb_shard_range(5, 3, 3, C.byref(f), C.byref(c)) != 0


def _shm_worker(rank, world, key, n, steps, q):
    sys.path.insert(0, ROOT)
    from ocr_rs_b200 import sharding
    g = sharding.ShmGather(rank, world, key, cap_bytes=1 << 20)
    first, count = sharding.shard_range(n, rank, world)
    out = []
    for step in range(steps):
        mine = _fake_result(first + step, count)  # a different payload every step
        g.publish(mine, step)
        if rank == 0:
            out.append(g.collect(step).arrays())
    if rank == 0:
        q.put(out)
    else:
        import time
        time.sleep(0.5)  # keep the segment alive until rank 0 has read the last step
    g.close()


def test_shm_gather_world_size_3():
    """The shared-memory gather the bench's e2e leg uses at N > 1: three ranks, five steps (slots are reused:
    the writer must wait for the reader's acknowledgement), results in rank order."""
    import multiprocessing as mp
    n, world, steps, key = 11, 3, 5, f"test{os.getpid()}"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shm_worker, args=(r, world, key, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from ocr_rs_b200.sharding import shard_range
    from ocr_rs_b200._ffi import Polygons
    for step in range(steps):
        want = Polygons.concat([_fake_result(shard_range(n, r, world)[0] + step, shard_range(n, r, world)[1]) for r in range(world)])
        for a, b in zip(out[step], want.arrays()):
            assert a.shape == b.shape and (a == b).all()


def _shm_worker_deferred(rank, world, key, n, steps, q):
    sys.path.insert(0, ROOT)
    from ocr_rs_b200 import sharding
    g = sharding.ShmGather(rank, world, key, cap_bytes=1 << 20)
    first, count = sharding.shard_range(n, rank, world)
    out = []
    for step in range(steps):
        g.publish(_fake_result(first + step, count), step)
        if rank == 0 and step >= 1:
            out.append(g.collect(step - 1).arrays())  # one step behind, as bench.py's e2e leg does
    if rank == 0:
        out.append(g.collect(steps - 1).arrays())  # the flush at the end of the timed region
        q.put(out)
    else:
        import time
        time.sleep(0.5)
    g.close()


def test_shm_gather_collect_one_step_behind():
    """bench.py's e2e leg at N > 1: rank 0 collects step i - 1 after publishing step i (it never waits for the slowest
    rank) and flushes the last step at the end.  Two slots per segment: a writer two steps ahead of the reader's
    acknowledgement waits, so no shard is overwritten before it is read."""
    import multiprocessing as mp
    n, world, steps, key = 13, 3, 6, f"testd{os.getpid()}"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shm_worker_deferred, args=(r, world, key, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from ocr_rs_b200.sharding import shard_range
    from ocr_rs_b200._ffi import Polygons
    assert len(out) == steps
    for step in range(steps):
        want = Polygons.concat([_fake_result(shard_range(n, r, world)[0] + step, shard_range(n, r, world)[1]) for r in range(world)])
        for a, b in zip(out[step], want.arrays()):
            assert a.shape == b.shape and (a == b).all()
