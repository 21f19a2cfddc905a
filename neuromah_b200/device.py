"""Device memory for the hot path: a small array type over the C ABI.

``DeviceArray`` is what the layers' ``output`` / ``dinputs`` / ``weights`` /
gradient attributes hold.  It quacks enough like ``numpy.ndarray`` for the
reference's scripts (``shape``, ``dtype``, ``len``, leading-axis slicing,
``np.asarray(x)`` via ``__array__``) while the data stays in HBM.  PyTorch is
not involved; memory comes from ``nm_malloc`` and every launch goes to one
process-wide non-blocking stream.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import gc
import os
import warnings

import numpy as np

from ._lib import lib

_state = {"device": None, "stream": None}


def init(device: int | None = None):
    """Select the GPU (default: LOCAL_RANK or 0) and create the library stream."""
    if _state["device"] is not None and (device is None or device == _state["device"]):
        return _state["device"]
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    n = C.c_int(0)
    lib.nm_device_count(C.byref(n))
    if n.value <= 0:
        raise RuntimeError("neuromah_b200: no CUDA device visible (there is no CPU fallback)")
    lib.nm_init(device % n.value)
    s = C.c_void_p()
    lib.nm_stream_create(C.byref(s))
    _state["device"], _state["stream"] = device % n.value, s
    return _state["device"]


def stream():
    if _state["stream"] is None:
        init()
    return _state["stream"]


def synchronize():
    lib.nm_stream_sync(stream())


def launch_count() -> int:
    n = C.c_ulonglong(0)
    lib.nm_launch_count(C.byref(n))
    return int(n.value)


# ---- releases are never issued while a stream capture is open -------------------------------------------------
# cudaFree / cudaFreeHost synchronise the device: issued by a finalizer (Python's cyclic GC can run one at any
# allocation) while a CUDA-graph capture is open on this thread they fail AND invalidate the capture.  While
# ``graph_capture`` is active every release is queued and drained after cudaStreamEndCapture.
_capture = {"depth": 0, "device": [], "host": []}


def capturing() -> bool:
    return _capture["depth"] > 0


def _release(kind: str, ptr: int):
    if _capture["depth"] > 0:
        _capture[kind].append(ptr)
        return
    try:
        (lib.nm_free if kind == "device" else lib.nm_host_free)(C.c_void_p(ptr))
    except Exception as e:                      # a finalizer cannot raise: make the failure visible instead
        warnings.warn(f"neuromah_b200: releasing {kind} block 0x{ptr:x} failed: {e}", RuntimeWarning)


def drain_deferred_frees():
    """Release what finalizers queued while a capture was open (no-op inside a capture)."""
    if _capture["depth"] > 0:
        return
    for kind in ("device", "host"):
        while _capture[kind]:
            _release(kind, _capture[kind].pop())


@contextlib.contextmanager
def graph_capture(s=None):
    """Capture the launches issued inside the ``with`` body on stream ``s`` into a CUDA graph.

    ``with graph_capture() as cap: ...`` leaves the instantiated executable graph in ``cap.graph`` (a ``c_void_p``
    for ``nm_graph_launch``).  Inside the body the cyclic GC is paused and frees are deferred, nothing may allocate
    device memory.  If the body (or the capture itself) fails, the capture is closed, the partial graph discarded and
    the ORIGINAL exception propagates -- no kernel of the body has run."""
    s = stream() if s is None else s

    class _Cap:
        graph = None
    cap = _Cap()
    gc_was_on = gc.isenabled()
    gc.disable()
    _capture["depth"] += 1
    try:
        lib.nm_graph_begin(s)
    except BaseException:
        _capture["depth"] -= 1
        gc.enable() if gc_was_on else None
        raise
    try:
        yield cap
    except BaseException:
        try:                                    # close the (possibly invalidated) capture; its own error is secondary
            lib.nm_graph_abort(s)
        except Exception:
            pass
        raise
    else:
        g = C.c_void_p()
        lib.nm_graph_end(s, C.byref(g))
        cap.graph = g
    finally:
        _capture["depth"] -= 1
        if gc_was_on:
            gc.enable()
        drain_deferred_frees()


class _Block:
    """Owner of one cudaMalloc'd block."""
    __slots__ = ("ptr", "nbytes")

    def __init__(self, nbytes: int):
        init()
        if _capture["depth"] > 0:
            raise RuntimeError("neuromah_b200: device allocation inside a CUDA-graph capture (allocate before capturing)")
        p = C.c_void_p()
        lib.nm_malloc(C.byref(p), max(int(nbytes), 16))
        self.ptr, self.nbytes = p.value, int(nbytes)

    def __del__(self):
        ptr = getattr(self, "ptr", None)      # __init__ may have raised before the block existed
        self.ptr = None
        if ptr and _release is not None:          # module globals are gone at interpreter shutdown: the driver reclaims the block
            _release("device", ptr)


_pinned_ranges = []      # (start, end) of live page-locked host blocks


def pinned_address(arr) -> int | None:
    """Host address of ``arr`` if it is a C-contiguous NumPy view into page-locked memory."""
    if not isinstance(arr, np.ndarray) or not arr.flags["C_CONTIGUOUS"]:
        return None
    a = arr.ctypes.data
    for lo, hi in _pinned_ranges:
        if lo <= a and a + arr.nbytes <= hi:
            return a
    return None


def new_stream():
    s = C.c_void_p()
    lib.nm_stream_create(C.byref(s))
    return s


def new_event():
    e = C.c_void_p()
    lib.nm_event_create(C.byref(e))
    return e


class PinnedBuffer:
    """Page-locked host buffer exposed as a NumPy array (for async H2D / D2H)."""

    def __init__(self, shape, dtype=np.float32):
        init()
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        count = int(np.prod(self.shape))
        self.nbytes = count * self.dtype.itemsize
        p = C.c_void_p()
        lib.nm_host_alloc(C.byref(p), max(self.nbytes, 16))
        self.ptr = p.value
        _pinned_ranges.append((self.ptr, self.ptr + max(self.nbytes, 16)))
        buf = (C.c_byte * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=count).reshape(self.shape)

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr and _release is not None and _pinned_ranges is not None:
            self.array = None
            _pinned_ranges[:] = [r for r in _pinned_ranges if r[0] != ptr]
            _release("host", ptr)


class DeviceArray:
    """Contiguous row-major tensor in HBM (float32 / int32 / uint8 / uint16=bf16 bits / float64)."""

    __array_priority__ = 100

    def __init__(self, shape, dtype=np.float32, *, block: _Block | None = None, offset: int = 0):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._block = block if block is not None else _Block(self.nbytes)
        self._offset = int(offset)
        if self._offset + self.nbytes > max(self._block.nbytes, 16):
            raise ValueError("DeviceArray view exceeds its block")

    # -- basic properties -----------------------------------------------------
    @property
    def ptr(self) -> int:
        return self._block.ptr + self._offset

    @property
    def p(self):
        return C.c_void_p(self.ptr)

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, ptr=0x{self.ptr:x})"

    # -- construction -----------------------------------------------------------
    @classmethod
    def from_numpy(cls, arr, dtype=None):
        a = np.ascontiguousarray(arr, dtype=dtype)
        out = cls(a.shape, a.dtype)
        out.copy_from(a)
        return out

    @classmethod
    def zeros(cls, shape, dtype=np.float32):
        out = cls(shape, dtype)
        out.fill_zero()
        return out

    # -- data movement ------------------------------------------------------------
    def copy_from(self, arr):
        """Host (or device) -> this array.  Pageable host sources are synchronous."""
        if isinstance(arr, DeviceArray):
            if arr.nbytes != self.nbytes:
                raise ValueError("size mismatch")
            lib.nm_d2d(self.p, arr.p, self.nbytes, stream())
            return self
        a = np.ascontiguousarray(arr, dtype=self.dtype)
        if a.size != self.size:
            raise ValueError(f"cannot copy array of shape {a.shape} into DeviceArray of shape {self.shape}")
        lib.nm_h2d(self.p, a.ctypes.data_as(C.c_void_p), self.nbytes, stream())
        synchronize()          # pageable source: keep `a` alive until the copy has landed
        return self

    def copy_from_pinned(self, pinned: PinnedBuffer, nbytes: int | None = None):
        """Asynchronous host -> device from page-locked memory (no sync)."""
        lib.nm_h2d(self.p, C.c_void_p(pinned.ptr), self.nbytes if nbytes is None else nbytes, stream())
        return self

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            lib.nm_d2h(out.ctypes.data_as(C.c_void_p), self.p, self.nbytes, stream())
            synchronize()
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def fill_zero(self):
        lib.nm_memset(self.p, 0, self.nbytes, stream())
        return self

    def copy(self):
        out = DeviceArray(self.shape, self.dtype)
        lib.nm_d2d(out.p, self.p, self.nbytes, stream())
        return out

    # -- views ------------------------------------------------------------------
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1])) or 1
            shape[shape.index(-1)] = self.size // known
        if int(np.prod(shape)) != self.size:
            raise ValueError(f"cannot reshape {self.shape} into {tuple(shape)}")
        return DeviceArray(shape, self.dtype, block=self._block, offset=self._offset)

    def __getitem__(self, key):
        # contiguous leading-axis slices stay on the device (batch slicing, Model.py:150-151)
        if isinstance(key, slice) and self.ndim >= 1:
            start, stop, step = key.indices(self.shape[0])
            if step == 1:
                rows = max(stop - start, 0)
                row_bytes = (self.nbytes // self.shape[0]) if self.shape[0] else 0
                return DeviceArray((rows,) + self.shape[1:], self.dtype, block=self._block,
                                   offset=self._offset + start * row_bytes)
        return self.numpy()[key]

    @property
    def T(self):
        return self.numpy().T

    def astype(self, dtype, copy=True):
        return self.numpy().astype(dtype)

    def tolist(self):
        return self.numpy().tolist()


def as_device(x, dtype=np.float32) -> DeviceArray:
    """NumPy (or anything array-like) -> DeviceArray of ``dtype``; DeviceArray passes through."""
    if isinstance(x, DeviceArray):
        if x.dtype != np.dtype(dtype):
            raise TypeError(f"expected a {np.dtype(dtype)} DeviceArray, got {x.dtype}")
        return x
    return DeviceArray.from_numpy(np.asarray(x), dtype=dtype)


def input_to_device(x) -> DeviceArray:
    """Model inputs: float32 on the device, except raw ``uint8`` pixels, which stay uint8 (a quarter of the bytes over
    PCIe) and are normalised on the device -- ``X.astype(np.float32) / 255.0`` of examples/test_Conv2D_MNIST.py:27."""
    if isinstance(x, DeviceArray):
        if x.dtype not in (np.dtype(np.float32), np.dtype(np.uint8)):
            raise TypeError(f"expected a float32 or uint8 DeviceArray, got {x.dtype}")
        return x
    a = np.asarray(x)
    return DeviceArray.from_numpy(a, dtype=np.uint8 if a.dtype == np.uint8 else np.float32)


def is_array(x) -> bool:
    return isinstance(x, (np.ndarray, DeviceArray))


def labels_to_device(y) -> DeviceArray:
    """Class ids (int32) on the device from sparse labels or one-hot rows.

    One-hot rows are reduced with argmax exactly as the reference's fused
    gradient does (Activation_Softmax_Loss_CategoricalCrossentropy.py:11-12).
    """
    if isinstance(y, DeviceArray):
        if y.dtype == np.int32 and y.ndim == 1:
            return y
        y = y.numpy()
    y = np.asarray(y)
    if y.ndim == 2:
        y = np.argmax(y, axis=1)
    elif y.ndim != 1:
        raise ValueError("y_true must be either sparse labels or one-hot encoded")
    return DeviceArray.from_numpy(y, dtype=np.int32)


class Scratch:
    """Grow-only device scratch buffer (split-K partials, loss staging)."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes: int) -> DeviceArray:
        if self._buf is None or self._buf.nbytes < nbytes:
            self._buf = DeviceArray((max(int(nbytes), 16),), np.uint8)
        return self._buf
