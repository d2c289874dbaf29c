"""soxr.distributed -- batches of independent clips sharded over the GPUs of one box.

Clips (and channels) never interact in the resampling path: the reference keeps one
resampler state per channel (libsoxr/src/soxr.c:547-560) and python-soxr creates a fresh
handle per call (src/soxr_ext.cpp:230), so a batch shards by clip with NO data-path
collective.  One process per GPU (torchrun); `torch.distributed` is used only for

  * the optional final gather of the output shards (NCCL all_gather / gather over NVLink,
    gloo in the CPU tests), and
  * barriers / max-over-ranks timing in bench.py.

A single stream (ResampleStream) does not shard: replicas only.

    lo, hi = shard_bounds(n_clips, world, rank)          # contiguous, sizes differ by at most 1
    y_local = resample_shard(x[lo:hi], 48000, 44100, 'VHQ')
    y = gather_shards(y_local, n_clips, dst=None)        # every rank gets [n_clips, frames, ch]

or in one call `resample_batch_sharded(x_full_or_local, ...)`.  `resample_batch_gathered` is the pipelined form of
resample + all_gather: the shard is cut into sub-batches and the gather of sub-batch i runs on a side stream while
sub-batch i + 1 is being resampled.
"""
import numpy as np

from . import soxr_ext


def shard_bounds(n_clips, world, rank):
    """Contiguous clip range [lo, hi) of `rank`: the first n_clips % world ranks hold one clip more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('rank {} outside world of size {}'.format(rank, world))
    if n_clips < 0:
        raise ValueError('negative clip count')
    q, r = divmod(int(n_clips), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_sizes(n_clips, world):
    return [hi - lo for lo, hi in (shard_bounds(n_clips, world, r) for r in range(world))]


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised (launch with torchrun, one process per GPU)')
    return dist


def resample_shard(x_local, in_rate, out_rate, quality='HQ', out=None, stream=None, _compute=None):
    """Resample this rank's clips ([clip, frame] or [clip, frame, channel]) on its own GPU.
    `_compute` replaces the GPU call in the CPU (gloo) tests; the product default has no CPU fallback."""
    if _compute is not None:
        return _compute(x_local, in_rate, out_rate, quality)
    from . import resample_batch
    if int(x_local.shape[0]) == 0:      # an empty shard (more ranks than clips) launches nothing
        total = soxr_ext.output_frames(in_rate, out_rate, int(x_local.shape[1]))
        shape = (0, total) + tuple(x_local.shape[2:])
        if type(x_local).__module__.startswith('torch'):
            import torch
            return torch.empty(shape, dtype=x_local.dtype, device=x_local.device)
        return np.empty(shape, dtype=x_local.dtype)
    return resample_batch(x_local, in_rate, out_rate, quality, out=out, stream=stream)


def gather_shards(y_local, n_clips, dst=None, group=None):
    """Final gather of the output shards, in clip order.

    dst=None: every rank receives the whole [n_clips, ...] result (all_gather);
    dst=r:    only rank r does (others get None).
    torch tensors stay on their device (NCCL for CUDA tensors); numpy shards go through
    CPU tensors (gloo).  Equal shards use one all_gather_into_tensor; ragged ones (n_clips not a
    multiple of the world size) are padded to the largest shard and trimmed afterwards.
    """
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_clips, world)
    is_np = isinstance(y_local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(y_local)) if is_np else y_local.contiguous()
    if int(t.shape[0]) != sizes[rank]:
        raise ValueError('rank {} holds {} clips, its shard of {} is {}'.format(rank, int(t.shape[0]), n_clips, sizes[rank]))
    tail = tuple(t.shape[1:])
    big = max(sizes)
    if big == 0:
        full = t.new_empty((0,) + tail)
        return full.numpy() if is_np else full
    send = t
    if sizes[rank] != big:
        send = t.new_zeros((big,) + tail)
        send[:sizes[rank]] = t
    if dst is None:
        buf = t.new_empty((world * big,) + tail)
        dist.all_gather_into_tensor(buf, send, group=group)
    else:
        parts = [t.new_empty((big,) + tail) for _ in range(world)] if rank == dst else None
        dist.gather(send, parts, dst=dst, group=group)
        if rank != dst:
            return None
        buf = torch.cat(parts, 0)
    if all(s == big for s in sizes):
        full = buf
    else:
        full = torch.cat([buf[r * big:r * big + sizes[r]] for r in range(world)], 0)
    return full.numpy() if is_np else full


def resample_batch_sharded(x, in_rate, out_rate, quality='HQ', gather=None, x_is_local=False, group=None,
                           _compute=None):
    """Resample a batch of equal-length clips over all ranks of the process group.

    x: the whole batch (every rank passes the same array and works on its own slice) or, with
       x_is_local=True, this rank's shard only.
    gather: None  -> return this rank's output shard (no collective at all);
            'all' -> all_gather, every rank returns the whole result;
            int r -> gather to rank r (others return None).
    """
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if x_is_local:
        x_local = x
        import torch
        n = torch.tensor([int(x.shape[0])], dtype=torch.int64,
                         device=x.device if type(x).__module__.startswith('torch') else 'cpu')
        dist.all_reduce(n, group=group)
        n_clips = int(n.item())
        if shard_sizes(n_clips, world)[rank] != int(x.shape[0]):
            raise ValueError('local shards must follow shard_bounds(): contiguous, sizes differing by at most one')
    else:
        n_clips = int(x.shape[0])
        lo, hi = shard_bounds(n_clips, world, rank)
        x_local = x[lo:hi]
    y_local = resample_shard(x_local, in_rate, out_rate, quality, _compute=_compute)
    if gather is None:
        return y_local
    return gather_shards(y_local, n_clips, dst=None if gather == 'all' else int(gather), group=group)


_SYMM_CACHE = {}


def _symmetric_result(shape, dtype, device, group):
    """A result buffer every rank can write into over NVLink peer mappings (torch symmetric memory), cached per shape:
    the rendezvous is a collective and the mapping is reused by every later call.  None where peer access is not
    available (the caller then exchanges through NCCL)."""
    dist = _dist()
    ranks = tuple(dist.get_process_group_ranks(group if group is not None else dist.group.WORLD))
    key = (tuple(shape), str(dtype), str(device), ranks)
    if key in _SYMM_CACHE:
        return _SYMM_CACHE[key]
    try:
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        buf = symm_mem.empty(*shape, dtype=dtype, device=device)
        hdl = symm_mem.rendezvous(buf, grp)
        world = dist.get_world_size(group)
        views = [hdl.get_buffer(r, tuple(shape), dtype) for r in range(world)]
        entry = (buf, hdl, views)
    except Exception:      # no peer access / an older torch: every rank takes the same branch (rendezvous is collective)
        entry = None
    _SYMM_CACHE[key] = entry
    return entry


def resample_batch_gathered(x_local, in_rate, out_rate, quality='HQ', sub_batches=2, group=None, _compute=None,
                            method='auto'):
    """Resample this rank's clips and gather the outputs on every rank, overlapped: returns [world * n_local, ...].

    The local shard (torch tensor; equal size on every rank) is processed in `sub_batches` pieces, each written
    straight into its final place in the gathered result.  While the next piece is computed, the finished one travels:

    * method 'peer' (CUDA, the default where peer access exists): the result lives in symmetric memory and every rank
      PUSHES its piece into the same place of every peer's result with device-to-device copies over NVLink
      (copy engines, no SMs: the persistent resampling kernel keeps the whole GPU, and a NCCL kernel queued beside it
      would wait for it).  The returned tensor is that cached buffer: it is overwritten by the next call with the
      same shapes.
    * method 'nccl': NCCL send / receive pairs on a side stream, every peer receiving into the final place too.

    Either way the exchange is bound by the link: every rank receives (world - 1) shards.
    """
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = int(x_local.shape[0])
    sizes = torch.tensor([n_local], dtype=torch.int64, device=x_local.device)
    lo_hi = torch.stack([sizes.clone(), -sizes.clone()])
    dist.all_reduce(lo_hi, op=dist.ReduceOp.MAX, group=group)
    if int(lo_hi[0].item()) != n_local or int(-lo_hi[1].item()) != n_local:
        raise ValueError('resample_batch_gathered needs equal shards on every rank (use gather_shards for ragged ones)')
    total = soxr_ext.output_frames(in_rate, out_rate, int(x_local.shape[1]))
    tail = (total,) + tuple(x_local.shape[2:])
    cuda = x_local.is_cuda
    symm = None
    if cuda and method in ('auto', 'peer') and n_local and world > 1 and _compute is None:
        symm = _symmetric_result((world * n_local,) + tail, x_local.dtype, x_local.device, group)
        if symm is None and method == 'peer':
            raise RuntimeError('resample_batch_gathered: no peer access between the GPUs of this group')
    if symm is not None:
        full, hdl, views = symm
        hdl.barrier()            # every rank is done reading the previous result held in this buffer
    else:
        full = x_local.new_empty((world * n_local,) + tail)
    if n_local == 0:
        return full
    sub_batches = max(1, min(int(sub_batches), n_local))
    bounds = [shard_bounds(n_local, sub_batches, j) for j in range(sub_batches)]
    comm = torch.cuda.Stream(device=x_local.device) if cuda else None
    works = []
    for lo, hi in bounds:
        y_j = full[rank * n_local + lo:rank * n_local + hi]
        if _compute is not None:
            y_j.copy_(torch.as_tensor(_compute(x_local[lo:hi], in_rate, out_rate, quality)))
        else:
            from . import resample_batch
            resample_batch(x_local[lo:hi], in_rate, out_rate, quality, out=y_j)
        outs = [full[r * n_local + lo:r * n_local + hi] for r in range(world)]
        if cuda:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev)          # the exchange starts when THIS piece is done, not the whole shard
                if symm is not None:
                    for k in range(1, world):       # staggered: at any time every rank writes to a different peer
                        r = (rank + k) % world
                        views[r][rank * n_local + lo:rank * n_local + hi].copy_(y_j, non_blocking=True)
                else:
                    # (all_gather with a list of outputs goes through a flattened staging buffer: 5.0 ms per step at
                    # N = 2 against 1.1 ms for one all_gather_into_tensor of the whole shard)
                    ops = []
                    for r in range(world):
                        if r != rank:
                            peer = r if group is None else dist.get_global_rank(group, r)
                            ops.append(dist.P2POp(dist.isend, y_j, peer, group))
                            ops.append(dist.P2POp(dist.irecv, outs[r], peer, group))
                    if ops:
                        works.extend(dist.batch_isend_irecv(ops))
        else:
            works.append(dist.all_gather(outs, y_j.clone(), group=group, async_op=True))
    for w in works:
        w.wait()
    if cuda:
        torch.cuda.current_stream(x_local.device).wait_stream(comm)
    if symm is not None:
        hdl.barrier()            # every peer's pushes into this rank's buffer have landed
    return full
