"""Multi-GPU sharding: clip batches split over ranks, no data-path collective, optional final
gather.  The host logic runs here with world_size 2 and 3 on the gloo backend (CPU), the per-shard compute
replaced by the oracle; the `gpu` tests run the real CUDA path under NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from soxr import distributed as D   # noqa: E402


def test_shard_bounds_cover_the_batch_contiguously():
    for n in (0, 1, 2, 5, 7, 8, 1024, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            b = [D.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert sizes == D.shard_sizes(n, world) and max(sizes) - min(sizes) <= 1
            assert sizes == sorted(sizes, reverse=True)
    assert D.shard_sizes(4096, 8) == [512] * 8          # BASELINE config 4: 512 clips per GPU
    assert D.shard_sizes(1024, 8) == [128] * 8
    with pytest.raises(ValueError):
        D.shard_bounds(4, 2, 2)
    with pytest.raises(ValueError):
        D.shard_bounds(4, 0, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_compute(x, in_rate, out_rate, quality):
    import soxr_oracle as O
    q = {'HQ': O.HQ, 'VHQ': O.VHQ}[quality]
    x = np.asarray(x)
    n_out = O.num_out(x.shape[1], in_rate / out_rate)
    y = np.zeros((x.shape[0], n_out) + x.shape[2:], dtype=np.float64)
    for i in range(x.shape[0]):
        y[i] = O.resample_f64(x[i], in_rate, out_rate, q)
    return y.astype(x.dtype)


def _gloo_worker(rank, world, port, n_clips, channels, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "python-soxr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from soxr import distributed as DD
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        shape = (n_clips, 600, channels) if channels else (n_clips, 600)
        x = np.random.default_rng(7).standard_normal(shape).astype(np.float32)
        # 1. sharded, no collective: this rank's slice only
        y_local = DD.resample_batch_sharded(x, 48000, 44100, 'HQ', gather=None, _compute=_oracle_compute)
        lo, hi = DD.shard_bounds(n_clips, world, rank)
        assert y_local.shape[0] == hi - lo
        # 2. all_gather: every rank gets the whole batch
        y_all = DD.resample_batch_sharded(x, 48000, 44100, 'HQ', gather='all', _compute=_oracle_compute)
        # 3. gather to the last rank, from local shards
        y_root = DD.resample_batch_sharded(x[lo:hi], 48000, 44100, 'HQ', gather=world - 1, x_is_local=True,
                                           _compute=_oracle_compute)
        assert (y_root is None) == (rank != world - 1)
        np.save(os.path.join(out_dir, "all_%d.npy" % rank), y_all)
        np.save(os.path.join(out_dir, "local_%d.npy" % rank), y_local)
        if y_root is not None:
            np.save(os.path.join(out_dir, "root.npy"), y_root)
        # 4. pipelined resample + all_gather (equal shards only): sub-batches gathered while the next one is computed
        if n_clips % world == 0:
            import torch
            y_pipe = DD.resample_batch_gathered(torch.from_numpy(x[lo:hi]), 48000, 44100, 'HQ', sub_batches=2,
                                                _compute=_oracle_compute)
            np.save(os.path.join(out_dir, "pipe_%d.npy" % rank), y_pipe.numpy())
        else:
            import torch
            with pytest.raises(ValueError):
                DD.resample_batch_gathered(torch.from_numpy(x[lo:hi]), 48000, 44100, 'HQ', _compute=_oracle_compute)
        # a shard that does not follow shard_bounds() is refused
        if world == 2 and n_clips == 5:
            bad = x[:1] if rank == 0 else x[1:]
            with pytest.raises(ValueError):
                DD.resample_batch_sharded(bad, 48000, 44100, 'HQ', gather='all', x_is_local=True, _compute=_oracle_compute)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clips,channels", [(2, 4, 2), (2, 5, 0), (3, 2, 1), (3, 7, 2), (3, 6, 1)])
def test_sharded_batch_matches_single_process_gloo(tmp_path, world, n_clips, channels):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(world, port, n_clips, channels, str(tmp_path)), nprocs=world, join=True)
    shape = (n_clips, 600, channels) if channels else (n_clips, 600)
    x = np.random.default_rng(7).standard_normal(shape).astype(np.float32)
    want = _oracle_compute(x, 48000, 44100, 'HQ')
    for r in range(world):
        got = np.load(tmp_path / ("all_%d.npy" % r))
        assert got.shape == want.shape and np.array_equal(got, want)
        lo, hi = D.shard_bounds(n_clips, world, r)
        assert np.array_equal(np.load(tmp_path / ("local_%d.npy" % r)), want[lo:hi])
        if n_clips % world == 0:
            assert np.array_equal(np.load(tmp_path / ("pipe_%d.npy" % r)), want)
    assert np.array_equal(np.load(tmp_path / "root.npy"), want)


def test_not_initialised_raises():
    x = np.zeros((2, 100), np.float32)
    with pytest.raises(RuntimeError):
        D.resample_batch_sharded(x, 48000, 44100)


# ---- CUDA / NCCL ---------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "python-soxr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import soxr
    from soxr import distributed as DD
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    try:
        n_clips = 6 * world + 1                       # ragged on purpose
        g = torch.Generator(device="cpu")
        g.manual_seed(11)
        x_cpu = torch.randn((n_clips, 9000, 2), generator=g)
        x = x_cpu.to(dev)
        y_all = DD.resample_batch_sharded(x, 48000, 44100, 'VHQ', gather='all')
        y_one = soxr.resample_batch(x, 48000, 44100, 'VHQ')
        assert y_all.shape == y_one.shape
        assert torch.equal(y_all, y_one)            # a clip's result does not depend on its batch neighbours
        lo, hi = DD.shard_bounds(n_clips, world, rank)
        y_loc = DD.resample_batch_sharded(x[lo:hi].contiguous(), 48000, 44100, 'VHQ', gather=0, x_is_local=True)
        # pipelined resample + all_gather on equal shards: sub-batch gathers overlap the next sub-batch's kernels
        xe = x[:6 * world]
        y_pipe = DD.resample_batch_gathered(xe[6 * rank:6 * rank + 6].contiguous(), 48000, 44100, 'VHQ', sub_batches=3)
        torch.cuda.synchronize()
        assert torch.equal(y_pipe, y_one[:6 * world])
        if rank == 0:
            assert torch.equal(y_loc, y_one)
            np.save(os.path.join(out_dir, "y.npy"), y_one[:2].cpu().numpy())
            np.save(os.path.join(out_dir, "x.npy"), x_cpu[:2].numpy())
    finally:
        dist.destroy_process_group()


def _run_nccl(tmp_path, world):
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import soxr_oracle as O
    x, y = np.load(tmp_path / "x.npy"), np.load(tmp_path / "y.npy")
    for i in range(x.shape[0]):
        want = O.resample_f64(x[i], 48000, 44100, O.VHQ)
        err = float(np.sqrt(np.mean((y[i].astype(np.float64) - want) ** 2)))
        assert err <= 1e-6, err                       # north_star tolerance for float32


@pytest.mark.gpu
def test_sharded_batch_nccl_one_gpu(tmp_path):
    _run_nccl(tmp_path, 1)


@pytest.mark.gpu
def test_sharded_batch_nccl_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_nccl(tmp_path, 2)
