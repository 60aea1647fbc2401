"""Device plumbing for the B200 path: buffers are torch CUDA tensors (torch is used for
memory ownership and stream identity only); all arithmetic goes through the C ABI of
libffthom_b200.so.  There is no CPU fallback — without a CUDA device every operation
raises."""
import ctypes as C

import numpy as np

from . import _lib as L

_state = {'ready': False, 'plans': {}}


def _torch():
    import torch
    return torch


def init(device=0):
    """Initialise the library on `device` (idempotent)."""
    if _state['ready']:
        return
    torch = _torch()
    if not torch.cuda.is_available():
        raise L.FhError('ffthompy_b200 needs a CUDA device (sm_100a); none is visible and there '
                        'is no CPU fallback')
    lib = L.load()
    torch.cuda.set_device(device)
    L.check(lib.fh_init(int(device)))
    L.check(lib.fh_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    _state['device'] = torch.device('cuda', device)
    _state['ready'] = True


def lib():
    init()
    return L.load()


def device():
    init()
    return _state['device']


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def empty(shape, complex_=False):
    torch = _torch()
    return torch.empty(tuple(int(s) for s in shape), dtype=torch.complex128 if complex_ else torch.float64,
                       device=device())


def zeros(shape, complex_=False):
    torch = _torch()
    return torch.zeros(tuple(int(s) for s in shape), dtype=torch.complex128 if complex_ else torch.float64,
                       device=device())


def upload(a):
    """numpy -> device (fp64 / complex128, C-contiguous)."""
    torch = _torch()
    a = np.asarray(a)
    if np.iscomplexobj(a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
    else:
        a = np.ascontiguousarray(a, dtype=np.float64)
    return torch.from_numpy(a).to(device())


def download(t):
    """device -> host (a fresh pageable array, as `Tensor.val` hands out a new NumPy array every time).  Measured for the
    0.8 GB solution of a 256^3 elasticity solve with a plain `cudaMemcpy`: 0.36 s, dominated by the page faults of the
    fresh destination taken by one thread; a page-locked destination from torch's host allocator is slower the first time
    (0.44 s: pinning 0.8 GB).  Large results therefore go through `fh_download`."""
    nbytes = t.numel()*t.element_size()
    if nbytes >= (4 << 20) and t.is_contiguous():
        # large results: pinned staging ring + parallel first touch of the destination (csrc/fh_host.cu)
        out = np.empty(tuple(t.shape), dtype=np.complex128 if t.dtype.is_complex else np.float64)
        L.check(lib().fh_download(C.c_void_p(out.ctypes.data), ptr(t), C.c_int64(nbytes)))
        return out
    return t.cpu().numpy()


def is_complex(t):
    return t.dtype.is_complex


def ndoubles(t):
    return int(t.numel())*(2 if t.dtype.is_complex else 1)


def plan(N):
    """Cached fh_plan handle for the grid N."""
    init()
    key = tuple(int(n) for n in N)
    p = _state['plans'].get(key)
    if p is None:
        p = C.c_void_p()
        L.check(L.load().fh_plan_create(C.byref(p), len(key), L.i64arr(key)))
        _state['plans'][key] = p
    return p


FORM_CODE = {0: 0, 'r': 1, 'c': 2}


def form_code(fft_form):
    return FORM_CODE[fft_form]


def launch_count():
    return int(L.load().fh_launch_count())


def synchronize():
    L.check(lib().fh_sync())
