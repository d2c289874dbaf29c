"""soxr -- B200-native drop-in for python-soxr's constant-rate resampling path.

Public surface and behaviour follow /root/reference/src/soxr/__init__.py:
`resample` (:182-231), `ResampleStream` (:61-179), `_resample_oneshot` (:234-249),
the quality constants `QQ LQ MQ HQ VHQ`, `__libsoxr_version__` and the same
ValueError / TypeError conditions (pinned by the reference's
tests/test_resample.py:20-43).  All sample processing happens on the GPU in
libb200soxr.so through the ctypes shim `soxr_ext`; there is no CPU fallback.

Extension (not in the reference): `resample_batch` for many equal-length clips
in one launch sequence, accepting numpy arrays or torch CUDA tensors.
"""
import numpy as np

from . import soxr_ext
from . import distributed  # noqa: F401  (clip batches sharded over GPUs; torch is imported lazily)
from .soxr_ext import QQ, LQ, MQ, HQ, VHQ
from ._version import version as __version__

__libsoxr_version__ = soxr_ext.libsoxr_version()

# one resampler state per channel is allocated: bound the channel count like the reference (:20-22)
_CH_LIMIT = 65536

_PRESETS = {'qq': QQ, 'lq': LQ, 'mq': MQ, 'hq': HQ, 'vhq': VHQ}
_DTYPE_CODES = {np.dtype(np.float32): soxr_ext.SOXR_FLOAT32_I, np.dtype(np.float64): soxr_ext.SOXR_FLOAT64_I,
                np.dtype(np.int32): soxr_ext.SOXR_INT32_I, np.dtype(np.int16): soxr_ext.SOXR_INT16_I}


def _quality_to_enum(q):
    if isinstance(q, str):
        key = q.lower()
        if key.startswith('soxr_'):
            key = key[5:]
        if key in _PRESETS:
            return _PRESETS[key]
    elif isinstance(q, (int, np.integer)) and not isinstance(q, bool) and int(q) in _PRESETS.values():
        return int(q)
    raise ValueError("Quality must be one of [QQ, LQ, MQ, HQ, VHQ]")


def _to_soxr_datatype(ntype):
    try:
        return _DTYPE_CODES[np.dtype(ntype)]
    except (KeyError, TypeError):
        raise TypeError('Data type must be one of [float32, float64, int16, int32], not {}'.format(ntype))


def _check_rates(in_rate, out_rate):
    if in_rate <= 0 or out_rate <= 0:
        raise ValueError('Sample rate should be over 0')


def _check_channels(n):
    if n < 1 or _CH_LIMIT < n:
        raise ValueError('Channel num({}) out of limit. Should be in [1, {}]'.format(n, _CH_LIMIT))


class ResampleStream:
    """Streaming resampler for real-time use or very long signals.

    Parameters: in_rate, out_rate (Hz, > 0), num_channels, dtype (float32 /
    float64 / int16 / int32), quality (QQ..VHQ or their names).  `vr=True`
    selects the variable-rate engine (the constructor's rates then give the MAXIMUM in/out ratio;
    `set_io_ratio` changes it between chunks, at once or over `slew_len` output samples).
    """

    def __init__(self, in_rate, out_rate, num_channels, dtype='float32', quality='HQ', vr=False, low_latency=False):
        _check_rates(in_rate, out_rate)
        _check_channels(num_channels)
        self._type = np.dtype(dtype)
        stype = _to_soxr_datatype(self._type)
        q = _quality_to_enum(quality)
        # low_latency (B200 extension): return every output whose filter support has arrived (about half a filter
        # length of delay) instead of the reference's FFT-block bursts; same filters, within the parity budget of
        # the default mode but not bit-identical to it
        self._csoxr = soxr_ext.CSoxr(in_rate, out_rate, num_channels, stype, q, vr, low_latency=low_latency)
        self._process = getattr(self._csoxr, 'process_{}'.format(self._type))

    def resample_chunk(self, x, last=False):
        """Feed one chunk (1-D mono or 2-D [frame, channel], dtype as constructed);
        returns whatever output is ready.  Pass last=True exactly once, on the final chunk."""
        if type(x) != np.ndarray or x.dtype != self._type:
            raise TypeError('Input should be a `np.ndarray` with matching dtype for ResampleStream({}).'
                            .format(self._type))
        if x.ndim == 1:
            return np.squeeze(self._process(x[:, np.newaxis], last), axis=1)
        if x.ndim == 2:
            return self._process(x, last)
        raise ValueError('Input must be 1-D or 2-D array')

    def num_clips(self):
        """Count of clipped (saturated) samples so far; integer I/O only."""
        return self._csoxr.num_clips()

    def delay(self):
        """Length of the pending output, in output samples."""
        return self._csoxr.delay()

    def clear(self):
        """Reset for a fresh signal with the same configuration."""
        self._csoxr.clear()

    def engine(self):
        return self._csoxr.engine()

    def set_io_ratio(self, in_rate, out_rate, slew_len=0):
        self._csoxr.set_io_ratio(in_rate / out_rate, slew_len)


class ResampleStreamBatch:
    """B200 extension: `n_streams` independent ResampleStreams of `num_channels` channels each, fed in lockstep
    through ONE library handle (one launch per chunk for all of them instead of one per stream).

    resample_chunk(x, last=False): x is [stream, frame, channel] (dtype as constructed); returns
    [stream, frames_out, channel].  Per stream the samples, the per-call frame counts, delay() and the int16
    dither are exactly those of a separate ResampleStream fed that stream's chunks.
    """

    def __init__(self, in_rate, out_rate, n_streams, num_channels, dtype='float32', quality='HQ', low_latency=False):
        _check_rates(in_rate, out_rate)
        _check_channels(num_channels)
        if n_streams < 1 or n_streams * num_channels > _CH_LIMIT * 64:
            raise ValueError('Stream num({}) out of limit'.format(n_streams))
        self._type = np.dtype(dtype)
        self._n = int(n_streams)
        stype = _to_soxr_datatype(self._type)
        q = _quality_to_enum(quality)
        self._csoxr = soxr_ext.CSoxr(in_rate, out_rate, num_channels, stype, q, False, n_streams=max(2, self._n),
                                     low_latency=low_latency) if self._n > 1 else \
            soxr_ext.CSoxr(in_rate, out_rate, num_channels, stype, q, False, low_latency=low_latency)
        self._process = getattr(self._csoxr, 'process_{}'.format(self._type))

    def resample_chunk(self, x, last=False):
        if type(x) != np.ndarray or x.dtype != self._type:
            raise TypeError('Input should be a `np.ndarray` with matching dtype for ResampleStreamBatch({}).'
                            .format(self._type))
        if x.ndim != 3 or x.shape[0] != self._n:
            raise ValueError('Input must be a [stream, frame, channel] array with {} streams'.format(self._n))
        if self._n == 1:
            return self._process(x[0], last)[np.newaxis]
        return self._process(x, last)

    def num_clips(self):
        """Clipped samples so far, summed over the streams."""
        return self._csoxr.num_clips()

    def delay(self):
        return self._csoxr.delay()

    def clear(self):
        self._csoxr.clear()

    def engine(self):
        return self._csoxr.engine()


def resample(x, in_rate, out_rate, quality='HQ'):
    """Resample a whole signal.

    x: 1-D (mono) or 2-D [frame, channel]; non-ndarray input is converted to
    float32.  Returns an ndarray with the same ndim and dtype as the input and
    floor(len * out_rate / in_rate + .5) frames.
    """
    _check_rates(in_rate, out_rate)
    if type(x) != np.ndarray:
        x = np.asarray(x, dtype=np.float32)
    # memory-layout dispatch of the reference (:212-217): unit stride along frames => split-channel driver
    kind = 'split_ch' if x.strides[0] == x.itemsize else 'divide_proc'
    proc = getattr(soxr_ext, 'csoxr_{}_{}'.format(kind, x.dtype), None)
    if proc is None:
        raise TypeError('Data type must be one of [float32, float64, int16, int32], not {}'.format(x.dtype))
    q = _quality_to_enum(quality)
    if x.ndim == 1:
        return np.squeeze(proc(in_rate, out_rate, x[:, np.newaxis], q), axis=1)
    if x.ndim == 2:
        _check_channels(x.shape[1])
        return proc(in_rate, out_rate, x, q)
    raise ValueError('Input must be 1-D or 2-D array')


def _resample_oneshot(x, in_rate, out_rate, quality='HQ'):
    """Single-call variant kept for the reference's tests (`soxr_oneshot`)."""
    oneshot = getattr(soxr_ext, 'csoxr_oneshot_{}'.format(x.dtype), None)
    if oneshot is None:
        raise TypeError('Data type must be one of [float32, float64, int16, int32], not {}'.format(x.dtype))
    q = _quality_to_enum(quality)
    if x.ndim == 1:
        return np.squeeze(oneshot(in_rate, out_rate, x[:, np.newaxis], q), axis=1)
    return oneshot(in_rate, out_rate, x, q)


def _check_batch_out(out, shape, dtype, like=None):
    """A caller-supplied `out` is written through raw pointers: it must be exactly what the call would allocate."""
    if tuple(out.shape) != tuple(shape):
        raise ValueError('out has shape {}, expected {}'.format(tuple(out.shape), tuple(shape)))
    if out.dtype != dtype:
        raise ValueError('out has dtype {}, expected {}'.format(out.dtype, dtype))
    if like is None:
        if type(out) != np.ndarray or not out.flags.c_contiguous or not out.flags.writeable:
            raise ValueError('out must be a writeable C-contiguous numpy.ndarray')
    else:
        if not out.is_cuda or out.device != like.device:
            raise ValueError('out must be a CUDA tensor on the same device as the input ({})'.format(like.device))
        if not out.is_contiguous():
            raise ValueError('out must be contiguous')


def resample_batch(x, in_rate, out_rate, quality='HQ', out=None, stream=None, exact_grid=False):
    """B200 extension: resample a batch of equal-length clips in one go.

    x: [clip, frame] (mono clips) or [clip, frame, channel], C-contiguous;
    a numpy array (host; copies included) or a torch CUDA tensor (device
    resident; asynchronous on `stream` -- an integer cudaStream_t handle -- or
    torch's current stream).  `out`, if given, must have exactly the shape,
    dtype, device and C-contiguous layout the call would allocate.
    Returns the same kind of container with floor(frames*out/in + .5) frames.

    exact_grid=True keeps the reference's block grid: every clip is then
    bit-identical to `soxr.resample` / a ResampleStream on that clip (int16
    included: the dither replays the per-clip call partition either way).  By
    default whole-clip float32 arithmetic at the ratio families of the
    single-pass frequency-domain kernel (48000:44100, 2:1, 3:1, ...) runs through
    that kernel, which agrees with the exact grid to ~2e-7 RMS (float32 rounding).
    """
    _check_rates(in_rate, out_rate)
    q = _quality_to_enum(quality)
    is_torch = type(x).__module__.startswith('torch')
    ndim = x.ndim
    if ndim not in (2, 3):
        raise ValueError('Input must be [clip, frame] or [clip, frame, channel]')
    if stream is not None and (isinstance(stream, bool) or not isinstance(stream, (int, np.integer))):
        raise TypeError('stream must be an integer cudaStream_t handle (e.g. torch.cuda.Stream.cuda_stream)')
    n_clips, ilen = int(x.shape[0]), int(x.shape[1])
    ch = int(x.shape[2]) if ndim == 3 else 1
    _check_channels(ch)
    total = soxr_ext.output_frames(in_rate, out_rate, ilen)
    shape = (n_clips, total, ch) if ndim == 3 else (n_clips, total)
    if n_clips == 0:
        if out is not None:
            _check_batch_out(out, shape, x.dtype, like=x if is_torch else None)
            return out
        if is_torch:
            import torch
            return torch.empty(shape, dtype=x.dtype, device=x.device)
        return np.empty(shape, dtype=x.dtype)
    # int16 dither replays the call partition `soxr.resample` would make for each clip (soxr_ext.cpp:239-257)
    flags = soxr_ext.BATCH_EMULATE_CALLS | (soxr_ext.BATCH_EXACT_GRID if exact_grid else 0)
    if is_torch:
        import torch
        if not x.is_cuda:
            raise ValueError('torch input must be a CUDA tensor (use numpy for host data)')
        x = x.contiguous()
        code = _to_soxr_datatype(str(x.dtype).replace('torch.', ''))
        if out is None:
            out = torch.empty(shape, dtype=x.dtype, device=x.device)
        else:
            _check_batch_out(out, shape, x.dtype, like=x)
        s = int(stream) if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
        soxr_ext.resample_batch_raw(x.data_ptr(), out.data_ptr(), in_rate=in_rate, out_rate=out_rate, quality=q,
                                    dtype_code=code, n_clips=n_clips, channels=ch, ilen=ilen,
                                    in_strides=(ilen * ch, ch, 1), out_strides=(total * ch, ch, 1), olen=total,
                                    flags=soxr_ext.BATCH_DEVICE_PTRS | flags, stream=s, want_clips=False)
        return out
    x = np.ascontiguousarray(x)
    code = _to_soxr_datatype(x.dtype)
    if out is None:
        out = np.empty(shape, dtype=x.dtype)
    else:
        _check_batch_out(out, shape, x.dtype)
    soxr_ext.resample_batch_raw(x.ctypes.data, out.ctypes.data, in_rate=in_rate, out_rate=out_rate, quality=q,
                                dtype_code=code, n_clips=n_clips, channels=ch, ilen=ilen,
                                in_strides=(ilen * ch, ch, 1), out_strides=(total * ch, ch, 1), olen=total,
                                flags=flags, stream=None if stream is None else int(stream))
    return out
