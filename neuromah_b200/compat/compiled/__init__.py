"""Drop-in for the reference's native seam ``src/layers/compiled/``.

The reference imports two pybind11 modules from that package -- ``_Conv2DBackend_cpp`` (Conv_wrapepr.py:3, used at :91;
built from CPU_kernels/Conv2D.cpp:242-245) and ``_MaxPooling2DBackend_cpp`` (MaxPooling2D.py:2, used at :53-58 and
:73-80; built from maxpooling_backend.cpp:240-247).  Copy this file to ``src/layers/compiled/__init__.py`` and the
reference's own ``Layer_Conv2D`` / ``Layer_MaxPooling2D`` run on ``libneuromah_b200.so`` unchanged: same function names,
argument order, return shapes / dtypes and error types (std::runtime_error -> RuntimeError).

Only ``ctypes`` + ``numpy`` and the C ABI of include/neuromah_b200.h are used -- nothing else of the neuromah_b200
package -- so the file is self-contained inside the reference tree.  The seam is per call and host-to-host, so every
call pays an H2D + D2H copy: it is the compatibility path (and a way to run the reference's own tests against the
kernels), not the fast path.  No CPU fallback: without the library / a GPU every call raises.
"""
import ctypes as C
import math
import os
import types

import numpy as np

_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_state = {"lib": None}


def _lib():
    if _state["lib"] is None:
        here = os.path.dirname(os.path.abspath(__file__))
        cands = [os.environ.get("NM_LIB_PATH"), os.path.join(here, "libneuromah_b200.so"),
                 os.path.join(here, "..", "..", "libneuromah_b200.so"), "libneuromah_b200.so"]
        path = next((p for p in cands if p and (os.path.sep not in p or os.path.exists(p))), "libneuromah_b200.so")
        L = C.CDLL(path)
        L.nm_last_error.restype = C.c_char_p
        L.nm_conv2d_wgrad_f32_workspace.restype = _sz
        L.nm_malloc.argtypes = [C.POINTER(_vp), _sz]
        L.nm_free.argtypes = [_vp]
        L.nm_h2d.argtypes = L.nm_d2h.argtypes = [_vp, _vp, _sz, _vp]
        L.nm_stream_sync.argtypes = [_vp]
        L.nm_conv2d_fprop_f32.argtypes = [_vp] * 4 + [_i] * 12 + [_vp]
        L.nm_conv2d_dgrad_f32.argtypes = [_vp] * 3 + [_i] * 11 + [_vp]
        L.nm_conv2d_wgrad_f32_workspace.argtypes = [_i] * 11
        L.nm_conv2d_wgrad_f32.argtypes = [_vp] * 5 + [_sz] + [_i] * 11 + [_vp]
        L.nm_maxpool2d_fwd_f32.argtypes = L.nm_maxpool2d_bwd_f32.argtypes = [_vp] * 3 + [_i] * 12 + [_vp]
        _state["lib"] = L
        _chk(L.nm_init(int(os.environ.get("LOCAL_RANK", "0"))))
    return _state["lib"]


def _chk(rc):
    if rc:
        raise RuntimeError(_state["lib"].nm_last_error().decode("utf-8", "replace"))


class _Dev:
    """Device buffers of one call; released on exit whatever happens."""

    def __init__(self):
        self.ptrs = []

    def alloc(self, nbytes):
        p = _vp()
        _chk(_lib().nm_malloc(C.byref(p), max(int(nbytes), 16)))
        self.ptrs.append(p)
        return p

    def upload(self, a, dtype=np.float32):
        a = np.ascontiguousarray(a, dtype=dtype)
        p = self.alloc(a.nbytes)
        if a.nbytes:
            _chk(_lib().nm_h2d(p, a.ctypes.data_as(_vp), a.nbytes, None))
            _chk(_lib().nm_stream_sync(None))          # pageable source: the copy has landed before `a` can go away
        return p

    def download(self, p, shape, dtype=np.float32):
        out = np.empty(shape, dtype)
        if out.nbytes:
            _chk(_lib().nm_d2h(out.ctypes.data_as(_vp), p, out.nbytes, None))
        _chk(_lib().nm_stream_sync(None))
        return out

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        for p in self.ptrs:
            _lib().nm_free(p)
        return False


# ------------------------------------------------------------------ _Conv2DBackend_cpp ----
def conv2d_cpu(input, kernel):
    """Conv2D.cpp:83-146: stride-1 cross-correlation of an ALREADY padded NCHW input (Conv_wrapepr.py:78-84 pads),
    (N,C,H,W) x (OC,C,kH,kW) -> float32 (N,OC,H-kH+1,W-kW+1)."""
    input, kernel = np.asarray(input), np.asarray(kernel)
    if input.ndim != 4 or kernel.ndim != 4:
        raise RuntimeError("Input and kernel must be 4D tensors")                    # Conv2D.cpp:86-87
    n, c, h, w = input.shape
    oc, kc, kh, kw = kernel.shape
    if kc != c:
        raise RuntimeError("Input channels must match kernel channels")              # Conv2D.cpp:101-102
    oh, ow = h - kh + 1, w - kw + 1
    if oh <= 0 or ow <= 0:
        raise RuntimeError("Kernel larger than the (padded) input")
    with _Dev() as d:
        x, k, y = d.upload(input), d.upload(kernel), d.alloc(4 * n * oc * oh * ow)
        _chk(_lib().nm_conv2d_fprop_f32(x, k, None, y, n, c, h, w, oc, kh, kw, 0, 0, 0, 0, 0, None))
        return d.download(y, (n, oc, oh, ow))


def conv2d_backward_cpu(input, kernel, doutput):
    """Conv2D.cpp:150-240 -> (dinputs (N,C,H,W), dweights (OC,C,kH,kW)); dweights summed over the batch, no 1/B."""
    input, kernel, doutput = np.asarray(input), np.asarray(kernel), np.asarray(doutput)
    if input.ndim != 4 or kernel.ndim != 4 or doutput.ndim != 4:
        raise RuntimeError("Input, kernel and doutput must be 4D tensors")
    n, c, h, w = input.shape
    oc, _, kh, kw = kernel.shape
    if doutput.shape != (n, oc, h - kh + 1, w - kw + 1):
        raise RuntimeError("doutput has the wrong shape")
    L = _lib()
    with _Dev() as d:
        x, k, dy = d.upload(input), d.upload(kernel), d.upload(doutput)
        dx, dw = d.alloc(4 * input.size), d.alloc(4 * kernel.size)
        nws = L.nm_conv2d_wgrad_f32_workspace(n, c, h, w, oc, kh, kw, 0, 0, 0, 0)
        ws = d.alloc(nws)
        _chk(L.nm_conv2d_dgrad_f32(dy, k, dx, n, c, h, w, oc, kh, kw, 0, 0, 0, 0, None))
        _chk(L.nm_conv2d_wgrad_f32(x, dy, dw, None, ws, nws, n, c, h, w, oc, kh, kw, 0, 0, 0, 0, None))
        return d.download(dx, input.shape), d.download(dw, kernel.shape)


# ------------------------------------------------------------------ _MaxPooling2DBackend_cpp ----
def _pool_geometry(h, w, ph, pw, sh, sw, padding):
    """maxpooling_backend.cpp:37-52."""
    if padding == "same":
        oh, ow = int(math.ceil(h / sh)), int(math.ceil(w / sw))
        return oh, ow, max((oh - 1) * sh + ph - h, 0) // 2, max((ow - 1) * sw + pw - w, 0) // 2
    if padding != "valid":
        raise RuntimeError("padding must be 'valid' or 'same'")
    return (h - ph) // sh + 1, (w - pw) // sw + 1, 0, 0


def maxpool2d_forward_cpu(input, pool_height, pool_width, stride_height, stride_width, padding):
    """maxpooling_backend.cpp:24-133 -> (output float32 (N,C,OH,OW), max_indices int32 (N,C,OH,OW,2))."""
    input = np.asarray(input)
    if input.ndim != 4:
        raise RuntimeError("Input must be a 4D tensor")
    n, c, h, w = input.shape
    oh, ow, pt, pl = _pool_geometry(h, w, pool_height, pool_width, stride_height, stride_width, padding)
    if oh <= 0 or ow <= 0:
        raise RuntimeError("pool window larger than the input")
    with _Dev() as d:
        x, y, idx = d.upload(input), d.alloc(4 * n * c * oh * ow), d.alloc(8 * n * c * oh * ow)
        _chk(_lib().nm_maxpool2d_fwd_f32(x, y, idx, n, c, h, w, pool_height, pool_width, stride_height, stride_width,
                                         pt, pl, oh, ow, None))
        return d.download(y, (n, c, oh, ow)), d.download(idx, (n, c, oh, ow, 2), np.int32)


def maxpool2d_backward_cpu(dvalues, max_indices, input, pool_height, pool_width, stride_height, stride_width, padding):
    """maxpooling_backend.cpp:145-238 -> dinputs float32 (N,C,H,W): scatter-add of dvalues to the saved arg-max."""
    dvalues, input = np.asarray(dvalues), np.asarray(input)
    n, c, h, w = input.shape
    oh, ow, pt, pl = _pool_geometry(h, w, pool_height, pool_width, stride_height, stride_width, padding)
    if dvalues.shape != (n, c, oh, ow) or np.asarray(max_indices).shape != (n, c, oh, ow, 2):
        raise RuntimeError("dvalues / max_indices do not match the pooled shape")
    with _Dev() as d:
        dy, idx, dx = d.upload(dvalues), d.upload(max_indices, np.int32), d.alloc(4 * input.size)
        _chk(_lib().nm_maxpool2d_bwd_f32(dy, idx, dx, n, c, h, w, pool_height, pool_width, stride_height, stride_width,
                                         pt, pl, oh, ow, None))
        return d.download(dx, input.shape)


_Conv2DBackend_cpp = types.SimpleNamespace(conv2d_cpu=conv2d_cpu, conv2d_backward_cpu=conv2d_backward_cpu)
_MaxPooling2DBackend_cpp = types.SimpleNamespace(maxpool2d_forward_cpu=maxpool2d_forward_cpu,
                                                 maxpool2d_backward_cpu=maxpool2d_backward_cpu)
