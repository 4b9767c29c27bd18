"""CPU tests of the multi-GPU host logic with the gloo backend (world_size 2 and 4): batch
sharding arithmetic and the slab-decomposed 3-D transform with its two all-to-all exchanges.
The local compute is the oracle here; on GPUs it is libjwave_cuda.so (tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jwave_b200.distributed import FORWARD, FWT, REVERSE, WPT, SlabVolumeTransform, shard_range
from oracle import c_oracle as co

CLS = "Coiflet2"


def oracle_axis_fn(kind, direction, x, outer, n, inner, level):
    a = x.numpy().reshape(outer, n, inner).transpose(0, 2, 1).reshape(-1, n)
    y = co.batch_1d(kind, direction, CLS, np.ascontiguousarray(a), level, threads=1)
    return torch.from_numpy(np.ascontiguousarray(y.reshape(outer, inner, n).transpose(0, 2, 1))).reshape(x.shape)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, levels, kind, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, Q, R = shape
        vol = np.random.default_rng(5).standard_normal(shape)
        lo, hi = shard_range(P, rank, world)
        slab = torch.from_numpy(vol[lo:hi].copy())
        t = SlabVolumeTransform(oracle_axis_fn, kind=kind)
        f = t.forward(slab, P, *levels)
        r = t.reverse(f, P, *levels)
        parts_f = [torch.empty_like(f) for _ in range(world)]
        parts_r = [torch.empty_like(r) for _ in range(world)]
        dist.all_gather(parts_f, f.contiguous())
        dist.all_gather(parts_r, r.contiguous())
        if rank == 0:
            np.save(os.path.join(out_dir, "f.npy"), torch.cat(parts_f).numpy())
            np.save(os.path.join(out_dir, "r.npy"), torch.cat(parts_r).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,levels,kind", [
    (2, (8, 8, 16), (3, 4, 3), FWT),      # (lvlP, lvlQ, lvlR): Q=8 gets lvlP, R=16 gets lvlQ, P=8 gets lvlR (F5)
    (4, (16, 8, 8), (2, 3, 4), FWT),
    (2, (4, 8, 8), (3, 2, 1), WPT),
])
def test_slab_volume_transform_equals_oracle_3d(tmp_path, world, shape, levels, kind):
    mp.spawn(_worker, args=(world, _free_port(), shape, levels, kind, str(tmp_path)), nprocs=world, join=True)
    vol = np.random.default_rng(5).standard_normal(shape)
    ref_f = co.transform_3d(kind, co.FORWARD, CLS, vol, *levels)
    ref_r = co.transform_3d(kind, co.REVERSE, CLS, ref_f, *levels)
    # same arithmetic, only the data distribution differs: bit-equal
    assert np.array_equal(np.load(tmp_path / "f.npy"), ref_f)
    assert np.array_equal(np.load(tmp_path / "r.npy"), ref_r)


def test_shard_range_covers_batch_without_overlap():
    for total in (0, 1, 7, 64, 65536, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_slab_transform_is_the_plain_3d_transform():
    vol = np.random.default_rng(9).standard_normal((8, 4, 8))
    t = SlabVolumeTransform(oracle_axis_fn)
    f = t.forward(torch.from_numpy(vol), 8, 2, 3, 3)
    assert np.array_equal(f.numpy(), co.transform_3d(co.FWT, co.FORWARD, CLS, vol, 2, 3, 3))
