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
    (variable rate) is not provided by the B200 library and raises.
    """

    def __init__(self, in_rate, out_rate, num_channels, dtype='float32', quality='HQ', vr=False):
        _check_rates(in_rate, out_rate)
        _check_channels(num_channels)
        self._type = np.dtype(dtype)
        stype = _to_soxr_datatype(self._type)
        q = _quality_to_enum(quality)
        self._csoxr = soxr_ext.CSoxr(in_rate, out_rate, num_channels, stype, q, vr)
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


def resample_batch(x, in_rate, out_rate, quality='HQ', out=None, stream=None, exact_grid=False):
    """B200 extension: resample a batch of equal-length clips in one go.

    x: [clip, frame] (mono clips) or [clip, frame, channel], C-contiguous;
    a numpy array (host; copies included) or a torch CUDA tensor (device
    resident; asynchronous on `stream` or torch's current stream).
    Returns the same kind of container with floor(frames*out/in + .5) frames.

    exact_grid=True keeps the reference's block grid (results bit-identical to a
    ResampleStream fed the same clip); by default whole-clip float32 VHQ work at
    48000:44100-type ratios runs through the single-pass frequency-domain kernel,
    which agrees with it to ~2e-7 RMS (float32 rounding).
    """
    _check_rates(in_rate, out_rate)
    q = _quality_to_enum(quality)
    is_torch = type(x).__module__.startswith('torch')
    ndim = x.ndim
    if ndim not in (2, 3):
        raise ValueError('Input must be [clip, frame] or [clip, frame, channel]')
    n_clips, ilen = int(x.shape[0]), int(x.shape[1])
    ch = int(x.shape[2]) if ndim == 3 else 1
    _check_channels(ch)
    total = soxr_ext.output_frames(in_rate, out_rate, ilen)
    if is_torch:
        import torch
        if not x.is_cuda:
            raise ValueError('torch input must be a CUDA tensor (use numpy for host data)')
        x = x.contiguous()
        code = _to_soxr_datatype(str(x.dtype).replace('torch.', ''))
        shape = (n_clips, total, ch) if ndim == 3 else (n_clips, total)
        if out is None:
            out = torch.empty(shape, dtype=x.dtype, device=x.device)
        s = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
        soxr_ext.resample_batch_raw(x.data_ptr(), out.data_ptr(), in_rate=in_rate, out_rate=out_rate, quality=q,
                                    dtype_code=code, n_clips=n_clips, channels=ch, ilen=ilen,
                                    in_strides=(ilen * ch, ch, 1), out_strides=(total * ch, ch, 1), olen=total,
                                    flags=soxr_ext.BATCH_DEVICE_PTRS | (soxr_ext.BATCH_EXACT_GRID if exact_grid else 0), stream=s)
        return out
    x = np.ascontiguousarray(x)
    code = _to_soxr_datatype(x.dtype)
    shape = (n_clips, total, ch) if ndim == 3 else (n_clips, total)
    if out is None:
        out = np.empty(shape, dtype=x.dtype)
    soxr_ext.resample_batch_raw(x.ctypes.data, out.ctypes.data, in_rate=in_rate, out_rate=out_rate, quality=q,
                                dtype_code=code, n_clips=n_clips, channels=ch, ilen=ilen,
                                in_strides=(ilen * ch, ch, 1), out_strides=(total * ch, ch, 1), olen=total,
                                flags=soxr_ext.BATCH_EXACT_GRID if exact_grid else 0)
    return out
