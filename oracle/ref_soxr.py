"""oracle/ref_soxr.py -- TEST INFRASTRUCTURE ONLY.

ctypes driver over oracle/_ref/libsoxr_ref.so, i.e. the UNMODIFIED reference
libsoxr compiled from /root/reference/libsoxr/src by oracle/Makefile.  It
re-states the three drivers of the reference's nanobind binding so that the
per-call block partition (which int16 dither depends on) is the reference's:

  * divide_proc  <- /root/reference/src/soxr_ext.cpp:210-273 (csoxr_divide_proc)
  * split_ch     <- /root/reference/src/soxr_ext.cpp:277-359 (csoxr_split_ch)
  * oneshot      <- /root/reference/src/soxr_ext.cpp:362-402 (csoxr_oneshot)
  * RefStream    <- /root/reference/src/soxr_ext.cpp:49-205  (class CSoxr)

and the python front-end's dispatch (/root/reference/src/soxr/__init__.py:182-249)
in `resample` / `resample_oneshot`.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.  Nothing in the product package does.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libsoxr_ref.so")

QQ, LQ, MQ, HQ, VHQ = 0, 1, 2, 4, 6            # soxr.h:284-293
FLOAT32_I, FLOAT64_I, INT32_I, INT16_I = 0, 1, 2, 3   # soxr.h:211-214
SPLIT = 4
NO_DITHER = 8                                   # soxr.h:235
DOUBLE_PRECISION = 16                           # soxr.h:253

_Q = {'vhq': VHQ, 'hq': HQ, 'mq': MQ, 'lq': LQ, 'qq': QQ,
      'soxr_vhq': VHQ, 'soxr_hq': HQ, 'soxr_mq': MQ, 'soxr_lq': LQ, 'soxr_qq': QQ}


class IoSpec(C.Structure):          # soxr.h:226-232
    _fields_ = [("itype", C.c_int), ("otype", C.c_int), ("scale", C.c_double),
                ("e", C.c_void_p), ("flags", C.c_ulong)]


class QualitySpec(C.Structure):     # soxr.h:239-246
    _fields_ = [("precision", C.c_double), ("phase_response", C.c_double),
                ("passband_end", C.c_double), ("stopband_begin", C.c_double),
                ("e", C.c_void_p), ("flags", C.c_ulong)]


class RuntimeSpec(C.Structure):     # soxr.h:258-265
    _fields_ = [("log2_min_dft_size", C.c_uint), ("log2_large_dft_size", C.c_uint),
                ("coef_size_kbytes", C.c_uint), ("num_threads", C.c_uint),
                ("e", C.c_void_p), ("flags", C.c_ulong)]


_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        L.soxr_version.restype = C.c_char_p
        L.soxr_io_spec.restype = IoSpec
        L.soxr_io_spec.argtypes = [C.c_int, C.c_int]
        L.soxr_quality_spec.restype = QualitySpec
        L.soxr_quality_spec.argtypes = [C.c_ulong, C.c_ulong]
        L.soxr_runtime_spec.restype = RuntimeSpec
        L.soxr_runtime_spec.argtypes = [C.c_uint]
        L.soxr_create.restype = C.c_void_p
        L.soxr_create.argtypes = [C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_char_p),
                                  C.POINTER(IoSpec), C.POINTER(QualitySpec), C.POINTER(RuntimeSpec)]
        L.soxr_process.restype = C.c_char_p
        L.soxr_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.soxr_oneshot.restype = C.c_char_p
        L.soxr_oneshot.argtypes = [C.c_double, C.c_double, C.c_uint,
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.POINTER(IoSpec), C.POINTER(QualitySpec), C.POINTER(RuntimeSpec)]
        L.soxr_delay.restype = C.c_double
        L.soxr_delay.argtypes = [C.c_void_p]
        L.soxr_num_clips.restype = C.POINTER(C.c_size_t)
        L.soxr_num_clips.argtypes = [C.c_void_p]
        L.soxr_engine.restype = C.c_char_p
        L.soxr_engine.argtypes = [C.c_void_p]
        L.soxr_clear.restype = C.c_char_p
        L.soxr_clear.argtypes = [C.c_void_p]
        L.soxr_set_io_ratio.restype = C.c_char_p
        L.soxr_set_io_ratio.argtypes = [C.c_void_p, C.c_double, C.c_size_t]
        L.soxr_delete.restype = None
        L.soxr_delete.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _dt(dtype):
    dtype = np.dtype(dtype)
    try:
        return {np.dtype(np.float32): FLOAT32_I, np.dtype(np.float64): FLOAT64_I,
                np.dtype(np.int32): INT32_I, np.dtype(np.int16): INT16_I}[dtype]
    except KeyError:
        raise TypeError(f"bad dtype {dtype}")


def _q(quality):
    if isinstance(quality, str):
        return _Q[quality.lower()]
    return int(quality)


def _create(in_rate, out_rate, channels, itype, otype, quality, io_flags=0, q_flags=0):
    L = lib()
    io = L.soxr_io_spec(itype, otype)
    io.flags = io_flags
    qs = L.soxr_quality_spec(_q(quality), q_flags)
    err = C.c_char_p()
    h = L.soxr_create(in_rate, out_rate, channels, C.byref(err), C.byref(io), C.byref(qs), None)
    if err.value:
        raise RuntimeError(err.value.decode())
    return h


def _div_len(in_rate, out_rate):
    return int(max(1000., 48000 * in_rate / out_rate))     # soxr_ext.cpp:71,239


def divide_proc(in_rate, out_rate, x, quality=HQ, io_flags=0, q_flags=0, return_calls=False,
                out_dtype=None, div_len=None):
    """x: C-contiguous (frames, channels).  Mirrors csoxr_divide_proc incl. its
    `div_len` call partition (soxr_ext.cpp:238-257).  `return_calls` also
    returns the list of per-call odone values (the dither block partition)."""
    L = lib()
    x = np.ascontiguousarray(x)
    ilen, ch = x.shape
    t = _dt(x.dtype)
    out_dtype = x.dtype if out_dtype is None else np.dtype(out_dtype)
    h = _create(in_rate, out_rate, ch, t, _dt(out_dtype), quality, io_flags, q_flags)
    olen = int(ilen * out_rate / in_rate + 1)
    if div_len is None:
        div_len = _div_len(in_rate, out_rate)
    y = np.zeros((olen, ch), dtype=out_dtype)
    isz = x.dtype.itemsize * ch
    osz = out_dtype.itemsize * ch
    out_pos = 0
    odone = C.c_size_t(0)
    calls = []
    for idx in range(0, ilen, div_len):
        n = min(div_len, ilen - idx)
        err = L.soxr_process(h, x.ctypes.data + idx * isz, n, None,
                             y.ctypes.data + out_pos * osz, olen - out_pos, C.byref(odone))
        if err:
            L.soxr_delete(h)
            raise RuntimeError(err.decode())
        out_pos += odone.value
        calls.append(odone.value)
    err = L.soxr_process(h, None, 0, None, y.ctypes.data + out_pos * osz, olen - out_pos, C.byref(odone))
    out_pos += odone.value
    calls.append(odone.value)
    clips = L.soxr_num_clips(h)[0]
    L.soxr_delete(h)
    if err:
        raise RuntimeError(err.decode())
    if return_calls:
        return y[:out_pos], calls, clips
    return y[:out_pos]


def split_ch(in_rate, out_rate, x, quality=HQ):
    """x: (frames, channels) with x.strides[0]==itemsize (Fortran order).
    Mirrors csoxr_split_ch (soxr_ext.cpp:277-359); returns a strided view
    {1, olen} like the reference (soxr_ext.cpp:358)."""
    L = lib()
    ilen, ch = x.shape
    if ilen and x.strides[0] != x.itemsize:
        raise ValueError("Data not contiguous")
    t = _dt(x.dtype) | SPLIT
    h = _create(in_rate, out_rate, ch, t, t, quality)
    olen = int(ilen * out_rate / in_rate + 1)
    div_len = _div_len(in_rate, out_rate)
    ybuf = np.zeros(olen * ch, dtype=x.dtype)
    st = x.strides[1] // x.itemsize if ch > 1 else ilen
    item = x.itemsize
    base = x.ctypes.data
    PtrArr = C.c_void_p * ch
    out_pos = 0
    odone = C.c_size_t(0)
    err = None
    for idx in range(0, ilen, div_len):
        ib = PtrArr(*[base + (st * c + idx) * item for c in range(ch)])
        ob = PtrArr(*[ybuf.ctypes.data + (olen * c + out_pos) * item for c in range(ch)])
        err = L.soxr_process(h, ib, min(div_len, ilen - idx), None, ob, olen - out_pos, C.byref(odone))
        out_pos += odone.value
    ob = PtrArr(*[ybuf.ctypes.data + (olen * c + out_pos) * item for c in range(ch)])
    err = L.soxr_process(h, None, 0, None, ob, olen - out_pos, C.byref(odone))
    out_pos += odone.value
    L.soxr_delete(h)
    if err:
        raise RuntimeError(err.decode())
    return np.lib.stride_tricks.as_strided(ybuf, shape=(out_pos, ch), strides=(item, olen * item))


def oneshot(in_rate, out_rate, x, quality=HQ, io_flags=0, q_flags=0):
    """Mirrors csoxr_oneshot (soxr_ext.cpp:362-402): a single soxr_oneshot call."""
    L = lib()
    x = np.ascontiguousarray(x)
    ilen, ch = x.shape
    t = _dt(x.dtype)
    io = L.soxr_io_spec(t, t)
    io.flags = io_flags
    qs = L.soxr_quality_spec(_q(quality), q_flags)
    olen = int(ilen * out_rate / in_rate + 1)
    y = np.zeros((olen, ch), dtype=x.dtype)
    odone = C.c_size_t(0)
    err = L.soxr_oneshot(in_rate, out_rate, ch, x.ctypes.data, ilen, None,
                         y.ctypes.data, olen, C.byref(odone), C.byref(io), C.byref(qs), None)
    if err:
        raise RuntimeError(err.decode())
    return y[:odone.value]


class RefStream:
    """Mirrors class CSoxr (soxr_ext.cpp:49-205) + ResampleStream (__init__.py:61-179)."""

    def __init__(self, in_rate, out_rate, num_channels, dtype='float32', quality='HQ',
                 io_flags=0, q_flags=0, vr=False):
        if vr:
            q_flags |= 32      # SOXR_VR, soxr.h:254 (soxr_ext.cpp:74)
        self._vr = vr
        self._ybytes = 0
        self._type = np.dtype(dtype)
        self._t = _dt(self._type)
        self._ch = num_channels
        self._oi = out_rate / in_rate
        self._div_len = _div_len(in_rate, out_rate)
        self._h = _create(in_rate, out_rate, num_channels, self._t, self._t, quality, io_flags, q_flags)
        self._ended = False

    def __del__(self):
        if getattr(self, "_h", None):
            lib().soxr_delete(self._h)
            self._h = None

    def resample_chunk(self, x, last=False):
        L = lib()
        squeeze = x.ndim == 1
        if squeeze:
            x = x[:, None]
        x = np.ascontiguousarray(x)
        if self._ended:
            raise RuntimeError("Input after last input")
        ilen, ch = x.shape
        isz = x.dtype.itemsize * ch
        olen = int(L.soxr_delay(self._h) + ilen * self._oi + 1) + 16
        y = np.zeros((olen, ch), dtype=x.dtype)
        out_pos = 0
        odone = C.c_size_t(0)
        for idx in range(0, ilen, self._div_len):
            err = L.soxr_process(self._h, x.ctypes.data + idx * isz, min(self._div_len, ilen - idx), None,
                                 y.ctypes.data + out_pos * isz, olen - out_pos, C.byref(odone))
            if err:
                raise RuntimeError(err.decode())
            out_pos += odone.value
        if last:
            self._ended = True
            while True:
                if olen <= out_pos:
                    y = np.concatenate([y, np.zeros_like(y)])
                    olen = y.shape[0]
                err = L.soxr_process(self._h, None, 0, None,
                                     y.ctypes.data + out_pos * isz, olen - out_pos, C.byref(odone))
                if err:
                    raise RuntimeError(err.decode())
                out_pos += odone.value
                if odone.value == 0:
                    break
        y = y[:out_pos].copy()
        return y[:, 0] if squeeze else y

    def set_io_ratio(self, in_rate, out_rate, slew_len=0):
        """__init__.py:162-179 -> soxr_ext.cpp:200-203"""
        err = lib().soxr_set_io_ratio(self._h, in_rate / out_rate, slew_len)
        if err:
            raise RuntimeError(err.decode())

    def _olen_buf(self, fsz, req_len=None):
        """_resize_ybuf (soxr_ext.cpp:89-107): frames the binding's reusable buffer holds"""
        if req_len is None:
            self._ybytes *= 2
        else:
            req = fsz * req_len
            if not (self._ybytes and req < self._ybytes):
                n = 1024
                while n < req:
                    n <<= 1
                self._ybytes = n
        return self._ybytes // fsz

    def process_vr(self, x, last=False):
        """CSoxr::process exactly as the binding drives it (soxr_ext.cpp:129-188), including the size of the
        reusable output buffer and the mid-call pulls of variable-rate mode (:170-173): what a variable-rate
        stream returns per call depends on the room each soxr_process call is offered."""
        L = lib()
        squeeze = x.ndim == 1
        if squeeze:
            x = x[:, None]
        x = np.ascontiguousarray(x)
        if self._ended:
            raise RuntimeError("Input after last input")
        ilen, ch = x.shape
        isz = x.dtype.itemsize * ch
        olen = self._olen_buf(isz, int(L.soxr_delay(self._h) + ilen * self._oi + 1))
        y = np.zeros((olen, ch), dtype=x.dtype)
        out_pos = 0
        odone = C.c_size_t(0)

        def pull(in_ptr):
            nonlocal y, olen, out_pos
            while True:
                if olen <= out_pos:
                    new_olen = self._olen_buf(isz)
                    y = np.concatenate([y, np.zeros((new_olen - olen, ch), dtype=x.dtype)])
                    olen = new_olen
                err = L.soxr_process(self._h, in_ptr, 0, None, y.ctypes.data + out_pos * isz, olen - out_pos,
                                     C.byref(odone))
                if err:
                    raise RuntimeError(err.decode())
                out_pos += odone.value
                if odone.value == 0:
                    break

        for idx in range(0, ilen, self._div_len):
            err = L.soxr_process(self._h, x.ctypes.data + idx * isz, min(self._div_len, ilen - idx), None,
                                 y.ctypes.data + out_pos * isz, olen - out_pos, C.byref(odone))
            if err:
                raise RuntimeError(err.decode())
            out_pos += odone.value
            if olen <= out_pos:
                pull(x.ctypes.data + idx * isz)
        if last:
            self._ended = True
            pull(None)
        y = y[:out_pos].copy()
        return y[:, 0] if squeeze else y

    def delay(self):
        return lib().soxr_delay(self._h)

    def num_clips(self):
        return lib().soxr_num_clips(self._h)[0]

    def engine(self):
        return lib().soxr_engine(self._h).decode()

    def clear(self):
        err = lib().soxr_clear(self._h)
        if err:
            raise RuntimeError(err.decode())
        self._ended = False


def resample(x, in_rate, out_rate, quality='HQ'):
    """Front-end dispatch of /root/reference/src/soxr/__init__.py:182-231."""
    if in_rate <= 0 or out_rate <= 0:
        raise ValueError('Sample rate should be over 0')
    if type(x) != np.ndarray:
        x = np.asarray(x, dtype=np.float32)
    fn = split_ch if x.strides[0] == x.itemsize else divide_proc
    if x.ndim == 1:
        return np.squeeze(fn(in_rate, out_rate, x[:, None], quality), axis=1)
    return fn(in_rate, out_rate, x, quality)


def resample_oneshot(x, in_rate, out_rate, quality='HQ'):
    if x.ndim == 1:
        return np.squeeze(oneshot(in_rate, out_rate, x[:, None], quality), axis=1)
    return oneshot(in_rate, out_rate, x, quality)


def trace_plan(in_rate, out_rate, quality, dtype='float32', q_flags=0, env=None):
    """Run soxr_create under SOXR_TRACE=1 in a subprocess and return the
    stderr plan lines (formats: SURVEY.md Appendix D; cr.c:232-233,440-441,484-485)."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import ref_soxr as r; "
            "h = r._create(%r, %r, 1, r._dt(%r), r._dt(%r), %r, 0, %r); r.lib().soxr_delete(h)"
            % (_HERE, float(in_rate), float(out_rate), str(dtype), str(dtype), quality, int(q_flags)))
    env = dict(os.environ, SOXR_TRACE="2", **(env or {}))
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    return p.stderr.splitlines()
