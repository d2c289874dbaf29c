"""ctypes shim over libb200soxr.so presenting the names of python-soxr's
compiled extension module `soxr_ext` (reference: /root/reference/src/soxr_ext.cpp,
module table :405-452), so the python front-end in soxr/__init__.py drives the
B200 library exactly the way the reference front-end drives its nanobind module.

  CSoxr                      <- class CSoxr            soxr_ext.cpp:49-205
  csoxr_divide_proc_{dtype}  <- csoxr_divide_proc<T>   soxr_ext.cpp:210-273
  csoxr_split_ch_{dtype}     <- csoxr_split_ch<T>      soxr_ext.cpp:277-359
  csoxr_oneshot_{dtype}      <- csoxr_oneshot<T>       soxr_ext.cpp:362-402
  libsoxr_version            <- csoxr_version.cpp

The GIL is released for the duration of every native call (ctypes does this for
CDLL functions), like nb::gil_scoped_release in the reference (:151,222,297,381).
There is no CPU fallback: if the CUDA library is missing or no GPU is usable the
calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libb200soxr.so")

QQ, LQ, MQ, HQ, VHQ = 0, 1, 2, 4, 6                                  # soxr.h:284-293
SOXR_FLOAT32_I, SOXR_FLOAT64_I, SOXR_INT32_I, SOXR_INT16_I = 0, 1, 2, 3   # soxr.h:211-214
_SPLIT = 4

BATCH_DEVICE_PTRS, BATCH_NO_DITHER, BATCH_DOUBLE = 1, 2, 4
BATCH_EMULATE_CALLS, BATCH_ONESHOT_CALLS, BATCH_SPLIT_SEMANTICS = 8, 16, 32
# Reference block grid only: bit-identical to a ResampleStream fed the same input.  The three whole-signal drivers
# below (divide_proc, split_ch, oneshot) always set it, so `resample == _resample_oneshot == ResampleStream`
# holds bit for bit at every ratio like in the reference (tests/test_resample.py:46-116); the faster single-pass
# frequency-domain kernel is what `soxr.resample_batch` / b200soxr_resample_batch use by default.
BATCH_EXACT_GRID = 64
RT_DEVICE_IO = 0x100
RT_LOW_LATENCY = 0x200


class _IoSpec(C.Structure):
    _fields_ = [("itype", C.c_int), ("otype", C.c_int), ("scale", C.c_double),
                ("e", C.c_void_p), ("flags", C.c_ulong)]


class _QualitySpec(C.Structure):
    _fields_ = [("precision", C.c_double), ("phase_response", C.c_double),
                ("passband_end", C.c_double), ("stopband_begin", C.c_double),
                ("e", C.c_void_p), ("flags", C.c_ulong)]


class _RuntimeSpec(C.Structure):
    _fields_ = [("log2_min_dft_size", C.c_uint), ("log2_large_dft_size", C.c_uint),
                ("coef_size_kbytes", C.c_uint), ("num_threads", C.c_uint),
                ("e", C.c_void_p), ("flags", C.c_ulong)]


class BatchDesc(C.Structure):
    _fields_ = [("in_rate", C.c_double), ("out_rate", C.c_double), ("quality", C.c_ulong),
                ("dtype", C.c_int), ("n_clips", C.c_uint), ("channels", C.c_uint), ("ilen", C.c_size_t),
                ("in_clip_stride", C.c_longlong), ("in_frame_stride", C.c_longlong), ("in_chan_stride", C.c_longlong),
                ("out_clip_stride", C.c_longlong), ("out_frame_stride", C.c_longlong), ("out_chan_stride", C.c_longlong),
                ("olen", C.c_size_t), ("flags", C.c_uint)]


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            "libb200soxr.so is not built (%s). Run `python __graft_entry__.py` at the repo root; "
            "this package has no CPU fallback." % _LIB_PATH)
    L = C.CDLL(_LIB_PATH)
    L.b200soxr_version.restype = C.c_char_p
    L.b200soxr_quality_spec.restype = _QualitySpec
    L.b200soxr_quality_spec.argtypes = [C.c_ulong, C.c_ulong]
    L.b200soxr_io_spec.restype = _IoSpec
    L.b200soxr_io_spec.argtypes = [C.c_int, C.c_int]
    L.b200soxr_runtime_spec.restype = _RuntimeSpec
    L.b200soxr_runtime_spec.argtypes = [C.c_uint]
    L.b200soxr_create.restype = C.c_void_p
    L.b200soxr_create.argtypes = [C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_char_p),
                                  C.POINTER(_IoSpec), C.POINTER(_QualitySpec), C.POINTER(_RuntimeSpec)]
    L.b200soxr_create_batch.restype = C.c_void_p
    L.b200soxr_create_batch.argtypes = [C.c_double, C.c_double, C.c_uint, C.c_uint, C.POINTER(C.c_char_p),
                                        C.POINTER(_IoSpec), C.POINTER(_QualitySpec), C.POINTER(_RuntimeSpec)]
    L.b200soxr_process.restype = C.c_char_p
    L.b200soxr_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.b200soxr_oneshot.restype = C.c_char_p
    L.b200soxr_oneshot.argtypes = [C.c_double, C.c_double, C.c_uint, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.POINTER(_IoSpec), C.POINTER(_QualitySpec), C.POINTER(_RuntimeSpec)]
    L.b200soxr_delay.restype = C.c_double
    L.b200soxr_delay.argtypes = [C.c_void_p]
    L.b200soxr_num_clips.restype = C.POINTER(C.c_size_t)
    L.b200soxr_num_clips.argtypes = [C.c_void_p]
    L.b200soxr_engine.restype = C.c_char_p
    L.b200soxr_engine.argtypes = [C.c_void_p]
    L.b200soxr_clear.restype = C.c_char_p
    L.b200soxr_clear.argtypes = [C.c_void_p]
    L.b200soxr_set_io_ratio.restype = C.c_char_p
    L.b200soxr_set_io_ratio.argtypes = [C.c_void_p, C.c_double, C.c_size_t]
    L.b200soxr_delete.restype = None
    L.b200soxr_delete.argtypes = [C.c_void_p]
    L.b200soxr_resample_batch.restype = C.c_char_p
    L.b200soxr_resample_batch.argtypes = [C.POINTER(BatchDesc), C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p]
    L.b200soxr_result_alloc.restype = C.c_void_p
    L.b200soxr_result_alloc.argtypes = [C.c_size_t]
    L.b200soxr_result_free.restype = None
    L.b200soxr_result_free.argtypes = [C.c_void_p]
    L.b200soxr_output_frames.restype = C.c_size_t
    L.b200soxr_output_frames.argtypes = [C.c_double, C.c_double, C.c_size_t]
    L.b200soxr_plan_dump.restype = C.c_size_t
    L.b200soxr_plan_dump.argtypes = [C.c_double, C.c_double, C.c_ulong, C.c_ulong, C.c_char_p, C.c_size_t]
    L.b200soxr_emulate_calls.restype = C.c_char_p
    L.b200soxr_emulate_calls.argtypes = [C.c_double, C.c_double, C.c_ulong, C.c_ulong,
                                         C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
    L.b200soxr_quantize.restype = C.c_char_p
    L.b200soxr_quantize.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t,
                                    C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.c_void_p,
                                    C.POINTER(C.c_size_t)]
    L.b200soxr_describe_path.restype = C.c_size_t
    L.b200soxr_describe_path.argtypes = [C.c_double, C.c_double, C.c_ulong, C.c_int, C.c_char_p, C.c_size_t]
    L.b200soxr_launch_count.restype = C.c_ulonglong
    L.b200soxr_device_available.restype = C.c_int
    return L


_lib = _load()

_DT_CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int16): 3}
_DT_NAME = {0: 'float32', 1: 'float64', 2: 'int32', 3: 'int16'}


def libsoxr_version():
    return _lib.b200soxr_version().decode()


def launch_count():
    """Number of CUDA kernels launched by the library so far (bench `gpu_launches`)."""
    return int(_lib.b200soxr_launch_count())


def device_available():
    return bool(_lib.b200soxr_device_available())


def output_frames(in_rate, out_rate, ilen):
    return int(_lib.b200soxr_output_frames(in_rate, out_rate, ilen))


def plan_dump(in_rate, out_rate, quality, q_flags=0):
    """SOXR_TRACE=1-compatible plan lines (host only; works without a GPU)."""
    n = _lib.b200soxr_plan_dump(in_rate, out_rate, quality, q_flags, None, 0)
    buf = C.create_string_buffer(n)
    _lib.b200soxr_plan_dump(in_rate, out_rate, quality, q_flags, buf, n)
    return buf.value.decode().splitlines()


def describe_path(in_rate, out_rate, quality, dtype='float32'):
    """One line naming the kernels that serve this conversion (exact grid and resample_batch default)."""
    code = _DT_CODE[np.dtype(dtype)]
    n = _lib.b200soxr_describe_path(in_rate, out_rate, int(quality), code, None, 0)
    buf = C.create_string_buffer(n)
    _lib.b200soxr_describe_path(in_rate, out_rate, int(quality), code, buf, n)
    return buf.value.decode()


def emulate_calls(in_rate, out_rate, quality, chunk_lens, olen_each=1 << 40, q_flags=0):
    """Output frames (and delay) the reference would return for each soxr_process
    call of the given partition; 0 in chunk_lens means a flush call (host only)."""
    n = len(chunk_lens)
    arr = (C.c_size_t * n)(*chunk_lens)
    od = (C.c_size_t * n)()
    dl = (C.c_double * n)()
    err = _lib.b200soxr_emulate_calls(in_rate, out_rate, quality, q_flags, arr, n, olen_each, od, dl)
    if err:
        raise RuntimeError(err.decode())
    return list(od), list(dl)


def resample_batch_raw(in_ptr, out_ptr, *, in_rate, out_rate, quality, dtype_code, n_clips, channels, ilen,
                       in_strides, out_strides, olen, flags=0, stream=None, want_clips=True):
    """Thin call into b200soxr_resample_batch with raw pointers (host or device).  want_clips=False passes
    clips=NULL: an integer batch on device pointers then returns without waiting, like a float one."""
    d = BatchDesc(in_rate, out_rate, quality, dtype_code, n_clips, channels, ilen,
                  in_strides[0], in_strides[1], in_strides[2],
                  out_strides[0], out_strides[1], out_strides[2], olen, flags)
    odone = C.c_size_t(0)
    clips = C.c_size_t(0)
    err = _lib.b200soxr_resample_batch(C.byref(d), in_ptr, out_ptr, C.byref(odone),
                                       C.byref(clips) if want_clips else None, stream)
    if err:
        raise RuntimeError(err.decode())
    return odone.value, clips.value


def quantize(planar, out_dtype, call_lens=None, seed=None, split=False):
    """Device K7 in isolation (b200soxr_quantize).  planar: [ch][n] float32/float64.
    seed None => no dither.  Returns (interleaved [n][ch], clips, new_seed)."""
    planar = np.ascontiguousarray(planar)
    ch, n = planar.shape
    out_dtype = np.dtype(out_dtype)
    calls = [n] if call_lens is None else list(call_lens)
    arr = (C.c_size_t * len(calls))(*calls)
    out = np.zeros((n, ch), dtype=out_dtype)
    sd = C.c_ulonglong(0 if seed is None else seed & 0xFFFFFFFFFFFFFFFF)
    clips = C.c_size_t(0)
    err = _lib.b200soxr_quantize(planar.ctypes.data, int(planar.dtype == np.float64), ch, n, arr, len(calls),
                                 _DT_CODE[out_dtype], int(seed is not None), int(split), C.byref(sd),
                                 out.ctypes.data, C.byref(clips))
    if err:
        raise RuntimeError(err.decode())
    return out, clips.value, (None if seed is None else sd.value)


def _div_len(in_rate, out_rate):
    return int(max(1000., 48000 * in_rate / out_rate))     # soxr_ext.cpp:71,239


class CSoxr:
    """Stream handle; mirrors class CSoxr of the reference binding (soxr_ext.cpp:49-205).

    Extensions: `n_streams` > 1 makes ONE handle drive that many independent streams fed in lockstep
    (b200soxr_create_batch; arrays are then [stream, frame, channel]); `low_latency` returns every output whose
    filter support has arrived instead of the reference's FFT-block bursts (B200SOXR_RT_LOW_LATENCY)."""

    def __init__(self, in_rate, out_rate, num_channels, ntype, quality, vr=False, n_streams=1, low_latency=False,
                 q_flags=0):
        self.in_rate = float(in_rate)
        self.out_rate = float(out_rate)
        self.ntype = int(ntype)
        self.channels = int(num_channels)
        self.n_streams = int(n_streams)
        self.ended = False
        self._oi_ratio = self.out_rate / self.in_rate
        self._div_len = _div_len(self.in_rate, self.out_rate)
        self._dtype = np.dtype(_DT_NAME[self.ntype])
        self._fsz = self._dtype.itemsize * self.channels
        io = _lib.b200soxr_io_spec(self.ntype, self.ntype)
        # `quality` is a libsoxr recipe (QQ..VHQ, optionally | 0x40 SOXR_STEEP_FILTER), q_flags its flags (roll-off ...)
        qs = _lib.b200soxr_quality_spec(int(quality), (32 if vr else 0) | int(q_flags))
        rt = _lib.b200soxr_runtime_spec(1)
        if low_latency:
            rt.flags |= RT_LOW_LATENCY
        err = C.c_char_p()
        if self.n_streams == 1:
            self._h = _lib.b200soxr_create(self.in_rate, self.out_rate, self.channels, C.byref(err),
                                           C.byref(io), C.byref(qs), C.byref(rt))
        else:
            self._h = _lib.b200soxr_create_batch(self.in_rate, self.out_rate, self.n_streams, self.channels,
                                                 C.byref(err), C.byref(io), C.byref(qs), C.byref(rt))
        if err.value:
            self._h = None
            raise RuntimeError(err.value.decode())
        self._odone = C.c_size_t(0)
        self._odone_ref = C.byref(self._odone)
        self._vr = bool(vr)
        self._ybytes = 0      # variable-rate handles: size of the reference's reusable output buffer (see _vr_olen)

    def _vr_olen(self, req_len=None):
        """Frames the reference binding's reusable output buffer holds (soxr_ext.cpp:89-107): it only ever grows,
        to the next power of two of the request, or doubles (req_len None).  The variable-rate engine's output
        of a call depends on the room it is offered, so a drop-in has to offer the same."""
        if req_len is None:
            self._ybytes *= 2
        else:
            req = self._fsz * req_len
            if not (self._ybytes and req < self._ybytes):
                n = 1024
                while n < req:
                    n <<= 1
                self._ybytes = n
        return self._ybytes // self._fsz

    def __del__(self, _delete=_lib.b200soxr_delete):   # (bound at definition: module globals may be gone at exit)
        h = getattr(self, "_h", None)
        if h:
            _delete(h)
            self._h = None

    def _process(self, x, last, code):
        B = self.n_streams
        if type(x) is not np.ndarray or x.ndim != (2 if B == 1 else 3):
            raise TypeError("process(): incompatible function arguments (need a %d-D ndarray)" % (2 if B == 1 else 3))
        if self.ended:
            raise RuntimeError("Input after last input")
        if x.shape[-1] != self.channels or (B > 1 and x.shape[0] != B):
            raise ValueError("Channel num mismatch")
        if code != self.ntype or x.dtype != self._dtype:
            raise TypeError("Data type mismatch")
        if not x.flags.c_contiguous:
            x = np.ascontiguousarray(x)
        ilen = x.shape[-2]
        fsz = self._fsz
        h = self._h
        process = _lib.b200soxr_process
        x_addr = x.__array_interface__['data'][0]      # (ndarray.ctypes builds a helper object per access: 3x slower)
        # lowest-latency sizing of the reference (soxr_ext.cpp:155-158)
        olen = int(_lib.b200soxr_delay(h) + ilen * self._oi_ratio + 1)
        if self._vr:
            olen = self._vr_olen(olen)
        odone = self._odone
        if B > 1:
            return self._process_batch(x, x_addr, ilen, olen, last)
        y = np.empty((max(olen, 1), self.channels), dtype=self._dtype)
        y_addr = y.__array_interface__['data'][0]
        out_pos = 0
        for idx in range(0, ilen, self._div_len):
            err = process(h, x_addr + idx * fsz, min(self._div_len, ilen - idx), None,
                          y_addr + out_pos * fsz, olen - out_pos, self._odone_ref)
            if err:
                raise RuntimeError(err.decode())
            out_pos += odone.value
            if olen <= out_pos:
                # variable-rate mode can fill the buffer: keep pulling, without new input, until nothing comes
                # (soxr_ext.cpp:170-173 -> _flush with a non-null input pointer and ilen 0)
                while True:
                    if olen <= out_pos:
                        grow = (self._vr_olen() - olen) if self._vr else max(olen, 1024)
                        y = np.concatenate([y[:out_pos], np.empty((olen - out_pos + grow, self.channels), dtype=self._dtype)])
                        y_addr = y.__array_interface__['data'][0]
                        olen = y.shape[0]
                    err = process(h, x_addr + idx * fsz, 0, None, y_addr + out_pos * fsz, olen - out_pos, self._odone_ref)
                    if err:
                        raise RuntimeError(err.decode())
                    out_pos += odone.value
                    if odone.value == 0:
                        break
        if last:
            self.ended = True
            while True:       # flush until no more output (soxr_ext.cpp:109-127)
                if olen <= out_pos:
                    grow = (self._vr_olen() - olen) if self._vr else max(olen, 1024)
                    y = np.concatenate([y[:out_pos], np.empty((olen - out_pos + grow, self.channels), dtype=self._dtype)])
                    y_addr = y.__array_interface__['data'][0]
                    olen = y.shape[0]
                err = process(h, None, 0, None, y_addr + out_pos * fsz, olen - out_pos, self._odone_ref)
                if err:
                    raise RuntimeError(err.decode())
                out_pos += odone.value
                if odone.value == 0:
                    break
        return y[:out_pos]      # a fresh array per call (the reference copies out of its reusable buffer, soxr_ext.cpp:187)

    def _process_batch(self, x, x_addr, ilen, olen, last):
        """n_streams > 1: one b200soxr_process call per chunk (and per flush round); the library packs the result
        as [stream][odone][channel], so every call's piece is a view of the flat buffer it was written into."""
        B, ch, h = self.n_streams, self.channels, self._h
        odone = self._odone
        pieces = []
        if ilen:
            buf = np.empty(B * max(olen, 1) * ch, dtype=self._dtype)
            err = _lib.b200soxr_process(h, x_addr, ilen, None, buf.__array_interface__['data'][0], olen,
                                        self._odone_ref)
            if err:
                raise RuntimeError(err.decode())
            pieces.append(buf[:B * odone.value * ch].reshape(B, odone.value, ch))
        if last:
            self.ended = True
            while True:
                cap = max(olen, 1024)
                buf = np.empty(B * cap * ch, dtype=self._dtype)
                err = _lib.b200soxr_process(h, None, 0, None, buf.__array_interface__['data'][0], cap, self._odone_ref)
                if err:
                    raise RuntimeError(err.decode())
                if odone.value == 0:
                    break
                pieces.append(buf[:B * odone.value * ch].reshape(B, odone.value, ch))
        if not pieces:
            return np.empty((B, 0, ch), dtype=self._dtype)
        return pieces[0] if len(pieces) == 1 else np.concatenate(pieces, axis=1)

    def process_float32(self, x, last=False): return self._process(x, last, 0)
    def process_float64(self, x, last=False): return self._process(x, last, 1)
    def process_int32(self, x, last=False): return self._process(x, last, 2)
    def process_int16(self, x, last=False): return self._process(x, last, 3)

    def num_clips(self):
        return int(_lib.b200soxr_num_clips(self._h)[0])

    def delay(self):
        return float(_lib.b200soxr_delay(self._h))

    def engine(self):
        return _lib.b200soxr_engine(self._h).decode()

    def clear(self):
        err = _lib.b200soxr_clear(self._h)
        if err:
            raise RuntimeError(err.decode())
        self.ended = False

    def set_io_ratio(self, io_ratio, slew_len=0):
        # soxr_ext.cpp:200-203
        err = _lib.b200soxr_set_io_ratio(self._h, float(io_ratio), int(slew_len))
        if err:
            raise RuntimeError(err.decode())


class _PinnedResult:
    """Owner of one page-locked result buffer of the library (b200soxr_result_alloc); the ndarray built over it keeps
    it alive through `base`, and it goes back to the library's cache when the last view is gone -- the role of the
    capsule that owns `new T[olen]` in the reference binding (soxr_ext.cpp:268-272)."""
    __slots__ = ('ptr', '__array_interface__')

    def __init__(self, ptr, shape, dtype):
        self.ptr = ptr
        self.__array_interface__ = {'shape': tuple(shape), 'typestr': np.dtype(dtype).str, 'data': (ptr, False),
                                    'version': 3}

    def __del__(self):
        try:
            _lib.b200soxr_result_free(self.ptr)
        except Exception:       # interpreter shutdown: the process is going away with the buffer
            pass


_RESULT_PINNED_MIN = 128 << 10


def _result_empty(shape, dtype):
    """Uninitialised result array: in page-locked memory of the library from 128 kB (the device-to-host copy then lands
    in it directly), ordinary memory below that or when the library's budget is spent."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes >= _RESULT_PINNED_MIN:
        ptr = _lib.b200soxr_result_alloc(nbytes)
        if ptr:
            return np.asarray(_PinnedResult(ptr, shape, dtype))
    return np.empty(shape, dtype=dtype)


def _check_2d(x, code):
    if type(x) is not np.ndarray or x.ndim != 2 or _DT_CODE.get(x.dtype) != code:
        raise TypeError("incompatible function arguments: need a 2-D %s ndarray" % _DT_NAME[code])


def _make_divide_proc(code):
    def csoxr_divide_proc(in_rate, out_rate, x, quality):
        _check_2d(x, code)
        x = np.ascontiguousarray(x)
        ilen, ch = x.shape
        total = output_frames(in_rate, out_rate, ilen)
        y = _result_empty((total, ch), x.dtype)
        resample_batch_raw(x.ctypes.data, y.ctypes.data, in_rate=in_rate, out_rate=out_rate, quality=int(quality),
                           dtype_code=code, n_clips=1, channels=ch, ilen=ilen,
                           in_strides=(0, ch, 1), out_strides=(0, ch, 1), olen=total,
                           flags=BATCH_EMULATE_CALLS | BATCH_EXACT_GRID)
        return y
    csoxr_divide_proc.__name__ = "csoxr_divide_proc_" + _DT_NAME[code]
    return csoxr_divide_proc


def _make_split_ch(code):
    def csoxr_split_ch(in_rate, out_rate, x, quality):
        _check_2d(x, code)
        if in_rate <= 0 or out_rate <= 0:
            raise ValueError("Sample rate should be over 0")
        ilen, ch = x.shape
        if ilen != 0 and x.strides[0] != x.itemsize:
            raise ValueError("Data not contiguous")
        if ch > 1 and ilen and (x.strides[1] % x.itemsize or x.strides[1] < 0):
            x = np.asfortranarray(x)
        olen = int(ilen * out_rate / in_rate + 1)          # soxr_ext.cpp:284
        ybuf = _result_empty((olen * ch,), x.dtype)   # (every element of the view returned below is written)
        st = (x.strides[1] // x.itemsize) if (ch > 1 and ilen) else max(ilen, 1)
        out_pos, _ = resample_batch_raw(x.ctypes.data, ybuf.ctypes.data, in_rate=in_rate, out_rate=out_rate,
                                        quality=int(quality), dtype_code=code, n_clips=1, channels=ch, ilen=ilen,
                                        in_strides=(0, 1, st), out_strides=(0, 1, olen), olen=olen,
                                        flags=BATCH_EMULATE_CALLS | BATCH_SPLIT_SEMANTICS | BATCH_EXACT_GRID)
        # strided view over the over-allocated buffer, like the reference (soxr_ext.cpp:358)
        return np.lib.stride_tricks.as_strided(ybuf, shape=(out_pos, ch),
                                               strides=(x.itemsize, olen * x.itemsize))
    csoxr_split_ch.__name__ = "csoxr_split_ch_" + _DT_NAME[code]
    return csoxr_split_ch


def _make_oneshot(code):
    def csoxr_oneshot(in_rate, out_rate, x, quality):
        _check_2d(x, code)
        x = np.ascontiguousarray(x)
        ilen, ch = x.shape
        total = output_frames(in_rate, out_rate, ilen)
        y = _result_empty((total, ch), x.dtype)
        resample_batch_raw(x.ctypes.data, y.ctypes.data, in_rate=in_rate, out_rate=out_rate, quality=int(quality),
                           dtype_code=code, n_clips=1, channels=ch, ilen=ilen,
                           in_strides=(0, ch, 1), out_strides=(0, ch, 1), olen=total,
                           flags=BATCH_ONESHOT_CALLS | BATCH_EXACT_GRID)
        return y
    csoxr_oneshot.__name__ = "csoxr_oneshot_" + _DT_NAME[code]
    return csoxr_oneshot


for _code, _name in _DT_NAME.items():
    globals()["csoxr_divide_proc_" + _name] = _make_divide_proc(_code)
    globals()["csoxr_split_ch_" + _name] = _make_split_ch(_code)
    globals()["csoxr_oneshot_" + _name] = _make_oneshot(_code)
del _code, _name
