"""ctypes binding of libtnac4o_b200.so (the C ABI declared in include/tnac4o_b200.h).

There is no CPU fallback: importing this module without the built library, or calling into it without a
CUDA device, raises.  torch is used only for device memory and streams.
"""
import ctypes
import os
import threading
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# TNAC4O_B200_LIB selects another build of the same library (tools/microbench: the phase-timer build)
LIB_PATH = os.environ.get('TNAC4O_B200_LIB') or os.path.join(_HERE, 'lib', 'libtnac4o_b200.so')

if not os.path.exists(LIB_PATH):
    raise ImportError('tnac4o_b200: %s is missing -- build it with `python tnac4o_b200/build.py` '
                      '(there is no CPU fallback)' % LIB_PATH)
lib = ctypes.CDLL(LIB_PATH)


class TnSite(ctypes.Structure):
    """mirror of `struct tn_site`"""
    _fields_ = [('nS', c_int), ('nl', c_int), ('nd', c_int), ('nr', c_int), ('nu', c_int),
                ('Wlu', c_void_p), ('Wtr', c_void_p), ('dmap', c_void_p), ('rmap', c_void_p),
                ('Es', c_void_p), ('Esl', c_void_p), ('Esu', c_void_p)]


P = c_void_p
_SIGS = {
    'tn_version': (c_int, []),
    'tn_last_error': (c_char_p, []),
    'tn_create': (c_int, [c_int, POINTER(c_void_p)]),
    'tn_destroy': (c_int, [P]),
    'tn_launch_count': (c_int64, [P]),
    'tn_profile': (c_int, [P, c_int]),
    'tn_profile_read': (c_int, [P, P, c_int]),
    'tn_set_blocking_sync': (c_int, [c_int]),
    'tn_set_throughput_mode': (c_int, [c_int]),
    'tn_gemm': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_double, P, c_int, c_int64, P, c_int, c_int64,
                        c_double, P, c_int, c_int64, c_int]),
    'tn_transpose': (c_int, [P, P, c_int, c_int, P, c_int, P, c_int]),
    'tn_qr_pos': (c_int, [P, P, c_int, c_int, P, c_int, P, c_int, P, c_int, P]),
    'tn_maxabs': (c_int, [P, P, P, c_int64, P]),
    'tn_pow2_scale': (c_int, [P, P, P, c_int64, P, P]),
    'tn_svd': (c_int, [P, P, c_int, c_int, P, c_int, P, c_int, P, P, c_int, c_int, POINTER(c_int)]),
    'tn_truncation_rank': (c_int, [P, P, P, c_int, c_double, c_int, POINTER(c_int), POINTER(c_double)]),
    'tn_mpo_apply': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    'tn_diff_norm': (c_int, [P, P, P, P, c_int, P]),
    'tn_row_compress': (c_int, [P, P, c_int, P, P, P, P, P, P, P, P, c_int, c_double, c_double, c_double, c_int, c_int,
                                POINTER(c_void_p)]),
    'tn_row_shapes': (c_int, [P, P, P]),
    'tn_row_fetch': (c_int, [P, P, POINTER(c_double), P, POINTER(c_double)]),
    'tn_row_free': (c_int, [P]),
    'tn_build_site_tables': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P]),
    'tn_rr_level': (c_int, [P, P, POINTER(TnSite), c_int, c_int, c_int, P, P, P, c_int, P]),
    'tn_marginals': (c_int, [P, P, POINTER(TnSite), c_int, c_int, P, P, P, P, c_int, c_int, P, P, P, P, P]),
    'tn_select': (c_int, [P, P, P, c_int64, P, c_double, P, P, P, POINTER(c_int)]),
    'tn_expand': (c_int, [P, P, POINTER(TnSite), c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, P, P, P, P, P, P,
                          P, P, P]),
    'tn_merge': (c_int, [P, P, c_int, P, P, P, P, P, P, P, c_double, P, P, P, P, P, P, POINTER(c_int)]),
    'tn_topm': (c_int, [P, P, c_int, c_int, P, P, P, P, P, P]),
    'tn_materialise': (c_int, [P, P, POINTER(TnSite), c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P,
                               P, P, P, P, P, P, P, P, P, P, P, P, P]),
    'tn_row_shift': (c_int, [P, P, c_int, c_int, P]),
    'tn_sample': (c_int, [P, P, POINTER(TnSite), c_int, c_int, c_int, c_int, P, P, P, c_int, P, P, P, P]),
    'tn_search_ground_state': (c_int, [P, P, c_int, c_int, P, P, P, P, c_int, c_double, c_double, P, P, P, P,
                                       POINTER(c_int), POINTER(c_double), POINTER(c_double), POINTER(c_int64)]),
    'tn_sort_keys': (c_int, [P, P, P, P, P, c_int]),
    'tn_sort_capacity_for': (c_int, [c_int]),
    'tn_xor_diff': (c_int, [P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, P]),
    'tn_apply_droplets': (c_int, [P, P, c_int, c_int, P, P, P, P, P, P, P]),
    'tn_energy_ising': (c_int, [P, P, c_int, c_int, P, c_int64, P, P, P, P]),
    'tn_book_create': (c_int, [P, P, c_int, c_int, POINTER(c_void_p)]),
    'tn_book_site': (c_int, [P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P, c_double, c_int]),
    'tn_book_sizes': (c_int, [P, P, P]),
    'tn_book_export': (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P]),
    'tn_book_free': (c_int, [P]),
    'tn_decode_enumerate': (c_int, [P, P, c_int, c_int, P, P, P, P, P, P, c_double, c_int64, POINTER(c_void_p),
                                    POINTER(c_int64)]),
    'tn_decode_fetch': (c_int, [P, P, c_int64, P, P, P, P, P, P]),
    'tn_decode_free': (c_int, [P]),
}
EXPORTED = sorted(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)          # AttributeError here = the library does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args


class NativeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise NativeError('tnac4o_b200 native call failed (%d): %s' % (rc, lib.tn_last_error().decode()))


class Context:
    """One library context (tn_create / tn_destroy) per CUDA device and host thread: a context owns its scratch
    memory, so concurrent solver instances -- one host thread and one CUDA stream each -- never share one."""

    _by_key = {}
    _retired = {}
    _lock = threading.Lock()

    def __init__(self, device_index):
        if not torch.cuda.is_available():
            raise NativeError('tnac4o_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', device_index)
        handle = c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream()           # make sure the primary context exists
            check(lib.tn_create(device_index, byref(handle)))
        self.handle = handle

    @classmethod
    def get(cls, device=None):
        if device is None or torch.device(device).index is None:
            # torch.device('cuda') has no index: it means the CURRENT device, not device 0
            index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        else:
            index = torch.device(device).index
        key = (index, threading.get_ident())
        ctx = cls._by_key.get(key)
        if ctx is None:
            with cls._lock:
                ctx = cls._by_key.get(key)
                if ctx is None:
                    ctx = cls._by_key[key] = Context(index)
        return ctx

    @classmethod
    def total_launches(cls, device=None):
        index = torch.device(device).index if device is not None else None
        if index is None:
            index = torch.cuda.current_device()
        with cls._lock:
            return sum(c.launch_count() for (i, _), c in cls._by_key.items() if i == index) + cls._retired.get(index, 0)

    @classmethod
    def release_thread(cls):
        """destroy the contexts of the calling host thread (scratch, pinned buffer, memory pool); their launch counts stay
        in the per-device total.  Called by worker threads of parallel.StreamPool before they exit."""
        me = threading.get_ident()
        with cls._lock:
            mine = [k for k in cls._by_key if k[1] == me]
            for k in mine:
                c = cls._by_key.pop(k)
                cls._retired[k[0]] = cls._retired.get(k[0], 0) + c.launch_count()
                lib.tn_destroy(c.handle)

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def launch_count(self):
        return int(lib.tn_launch_count(self.handle))

    PROFILE_CATEGORIES = ('gemm', 'qr', 'svd', 'mps_other', 'right_env', 'marginals', 'select_merge')

    def profile(self, on=True):
        """start (and reset) / stop the per-primitive timing of the native drivers on this context"""
        check(lib.tn_profile(self.handle, int(bool(on))))

    def profile_read(self):
        """{category: {'seconds', 'flops', 'bytes', 'calls'}} accumulated since profile(True); synchronises"""
        n = len(self.PROFILE_CATEGORIES)
        buf = (c_double * (4 * n))()
        check(lib.tn_profile_read(self.handle, buf, n))
        return {name: dict(zip(('seconds', 'flops', 'bytes', 'calls'), buf[4 * i:4 * i + 4]))
                for i, name in enumerate(self.PROFILE_CATEGORIES)}


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()
