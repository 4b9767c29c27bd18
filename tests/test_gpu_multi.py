"""Multi-GPU tests (need >= 2 GPUs; skipped on a single-GPU box): the slab-decomposed 3-D
transform over NCCL with libjwave_cuda.so as the local compute equals the oracle's 3-D result."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, levels, out_dir):
    import torch.distributed as dist
    import jwave_b200 as jw
    from jwave_b200.device import DeviceTransforms
    from jwave_b200.distributed import FWT, SlabVolumeTransform, device_axis_fn, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        P = shape[0]
        vol = np.random.default_rng(5).standard_normal(shape)
        lo, hi = shard_range(P, rank, world)
        dev = DeviceTransforms(jw.WaveletBuilder.create("Coiflet5"), rank)
        t = SlabVolumeTransform(device_axis_fn(dev), kind=FWT)
        slab = torch.from_numpy(vol[lo:hi].copy()).cuda()
        f = t.forward(slab, P, *levels)
        r = t.reverse(f, P, *levels)
        np.save(os.path.join(out_dir, f"f{rank}.npy"), f.cpu().numpy())
        np.save(os.path.join(out_dir, f"r{rank}.npy"), r.cpu().numpy())
        # the same volume with the exchanges folded into the kernels' stores (peer-mapped slabs)
        from jwave_b200.distributed import PeerSlabVolumeTransform
        pt = PeerSlabVolumeTransform(dev, *shape)
        for it in range(2):  # twice: buffer reuse across calls must not race
            pf = pt.forward(slab, P, *levels)
            pr = pt.reverse(pf, P, *levels)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"pf{rank}.npy"), pf.cpu().numpy())
        np.save(os.path.join(out_dir, f"pr{rank}.npy"), pr.cpu().numpy())
        ct = PeerSlabVolumeTransform(dev, *shape, exchange="copies")
        for it in range(2):
            cf = ct.forward(slab, P, *levels)
            cr = ct.reverse(cf, P, *levels)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"cf{rank}.npy"), cf.cpu().numpy())
        np.save(os.path.join(out_dir, f"cr{rank}.npy"), cr.cpu().numpy())
        # coefficients left in the j-slab layout: one re-cut per direction, the reverse rebuilds axis i first.
        # Results live in internal buffers until the next call, which may take them as its input (chained).
        for it in range(2):
            tf = ct.forward_t(slab, P, *levels)
            tf_dense = ct.t_to_dense(tf).cpu().numpy()
            tr = ct.reverse_t(tf, P, *levels)
            tr_host = tr.cpu().numpy()
            tf2 = ct.forward_t(tr, P, *levels)
            tf2_dense = ct.t_to_dense(tf2).cpu().numpy()
        np.save(os.path.join(out_dir, f"tf{rank}.npy"), tf_dense)
        np.save(os.path.join(out_dir, f"tr{rank}.npy"), tr_host)
        np.save(os.path.join(out_dir, f"tg{rank}.npy"), tf2_dense)
    finally:
        dist.destroy_process_group()


def test_slab_volume_transform_on_gpus(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    shape, levels = (64, 32, 64), (5, 6, 6)  # (lvlP, lvlQ, lvlR): Q = 32 gets 5, R = 64 gets 6, P = 64 gets 6
    mp.spawn(_worker, args=(world, _free_port(), shape, levels, str(tmp_path)), nprocs=world, join=True)
    vol = np.random.default_rng(5).standard_normal(shape)
    ref_f = co.transform_3d(co.FWT, co.FORWARD, "Coiflet5", vol, *levels)
    ref_r = co.transform_3d(co.FWT, co.REVERSE, "Coiflet5", ref_f, *levels)
    got_f = np.concatenate([np.load(tmp_path / f"f{r}.npy") for r in range(world)])
    got_r = np.concatenate([np.load(tmp_path / f"r{r}.npy") for r in range(world)])
    assert np.abs(got_f - ref_f).max() <= 1e-12 * np.abs(vol).max()
    assert np.abs(got_r - ref_r).max() <= 1e-12 * np.abs(ref_f).max()
    peer_f = np.concatenate([np.load(tmp_path / f"pf{r}.npy") for r in range(world)])
    peer_r = np.concatenate([np.load(tmp_path / f"pr{r}.npy") for r in range(world)])
    assert np.array_equal(peer_f, got_f)  # same kernels, same arithmetic: only the stores differ
    assert np.array_equal(peer_r, got_r)
    for tag, want in (("cf", got_f), ("cr", got_r)):  # peer-mapped slabs filled by strided device copies
        assert np.array_equal(np.concatenate([np.load(tmp_path / f"{tag}{r}.npy") for r in range(world)]), want)
    # transposed coefficient layout: rank g holds coef[:, g q:(g+1) q, :]
    t_f = np.concatenate([np.load(tmp_path / f"tf{r}.npy") for r in range(world)], axis=1)
    t_r = np.concatenate([np.load(tmp_path / f"tr{r}.npy") for r in range(world)])
    t_g = np.concatenate([np.load(tmp_path / f"tg{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(t_f, got_f)                                      # same passes, one re-cut fewer
    assert np.abs(t_r - ref_r).max() <= 1e-12 * np.abs(ref_f).max()        # i-first order: rounding-level difference
    assert np.abs(t_r - co.parallel_3d(co.FWT, co.REVERSE, "Coiflet5", got_f, *levels)).max() <= 1e-12 * np.abs(ref_f).max()
    assert np.abs(t_g - co.transform_3d(co.FWT, co.FORWARD, "Coiflet5", t_r, *levels)).max() <= 1e-12 * np.abs(t_r).max()


def test_device_group_through_the_c_abi():
    """jwc_create_multi: ONE context, one process, no torch.distributed - the batched host
    entry points shard over the GPUs and jwc_fwt3d slab-decomposes the volume internally.  Checked against the
    oracle through the host-side mirror classes, i.e. through ctypes and the C ABI only."""
    import jwave_b200 as jw
    from jwave_b200.transforms import CudaContext
    ndev = min(torch.cuda.device_count(), 4)
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    ctx = CudaContext(list(range(ndev)))
    assert ctx.device_count() == ndev
    rng = np.random.default_rng(21)
    fwt = jw.CudaFastWaveletTransform(jw.WaveletBuilder.create("Coiflet5"), context=ctx)
    # 3-D: cube, a non-cubic volume with the reference's level shift, and a shape that cannot be slab-cut (falls
    # back to device 0)
    for shape, levels in (((64, 64, 64), ()), ((64, 32, 128), (5, 7, 6)), ((8, 8, 16), (3, 4, 3)), ((2, 8, 8), (3, 3, 1))):
        vol = rng.standard_normal(shape)
        want = co.transform_3d(co.FWT, co.FORWARD, "Coiflet5", vol, *levels)
        got = fwt.forward(vol, *levels)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(vol).max(), shape
        back = fwt.reverse(want, *levels)   # axis i first on a group: rounding-level difference to the reference order
        assert np.abs(back - co.transform_3d(co.FWT, co.REVERSE, "Coiflet5", want, *levels)).max() <= 1e-12 * np.abs(want).max()
    wpt = jw.CudaWaveletPacketTransform(jw.WaveletBuilder.create("Daubechies4"), context=ctx)
    vol = rng.standard_normal((32, 32, 32))
    assert np.abs(wpt.forward(vol) - co.transform_3d(co.WPT, co.FORWARD, "Daubechies4", vol)).max() <= 1e-12 * np.abs(vol).max()
    # batched 1-D and 2-D: contiguous blocks per GPU, including a batch that does not divide evenly
    x = rng.standard_normal((1001, 512))
    assert np.abs(fwt.forwardBatch(x) - co.batch_1d(co.FWT, co.FORWARD, "Coiflet5", x, 9)).max() <= 1e-12 * np.abs(x).max()
    c = co.batch_1d(co.WPT, co.FORWARD, "Daubechies4", x, 4)
    assert np.abs(wpt.reverseBatch(c, 4) - co.batch_1d(co.WPT, co.REVERSE, "Daubechies4", c, 4)).max() <= 1e-12 * np.abs(c).max()
    imgs = rng.standard_normal((7, 64, 32))
    want = np.stack([co.transform_2d(co.FWT, co.FORWARD, "Coiflet5", m) for m in imgs])
    assert np.abs(fwt.forwardBatch2D(imgs) - want).max() <= 1e-12 * np.abs(imgs).max()
    # failure classes are those of the single-device path
    with pytest.raises(jw.JWaveFailure):
        fwt.forward(rng.standard_normal((64, 48, 64)))
    with pytest.raises(jw.JWaveFailure):
        fwt.forward(rng.standard_normal((64, 64, 64)), 7, 6, 6)
    ctx.close()
