"""Flat parameter arena: weights | grads | optimizer-state slabs in HBM.

One contiguous FP32 slab per role with a per-layer offset table, so that
  (i)   Adam / SGD update the whole model in one vectorised launch,
  (ii)  the gradient slab IS the NCCL all-reduce buffer (data parallel),
  (iii) FedAvg is one scale + all-reduce over the weight slab.
Layers keep their reference attribute names; after ``adopt`` those attributes are
views into the slabs.  Every tensor starts on a 16-byte boundary (float4 access);
the padding elements are zero in every slab and stay zero under Adam/SGD.
"""
from __future__ import annotations

import numpy as np

from .device import DeviceArray

_ALIGN = 4   # floats


def _param_attr_names(layer):
    """(param attr, grad attr) pairs in get_parameters() order: weights first, then biases."""
    if hasattr(layer, "weight_gradients"):           # Layer_Conv2D (Conv_wrapepr.py:113-115)
        return [("weights", "weight_gradients"), ("biases", "bias_gradients")]
    return [("weights", "dweights"), ("biases", "dbiases")]    # Layer_Dense (Dense.py:97-98)


class ParamArena:
    def __init__(self, layers):
        """``layers``: unique trainable layers in model order."""
        self.layers = list(layers)
        self.table = []               # (layer, name, offset, size, shape)
        off = 0
        for layer in self.layers:
            for pname, _ in _param_attr_names(layer):
                arr = getattr(layer, pname)
                self.table.append((layer, pname, off, arr.size, arr.shape))
                off += (arr.size + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.n_params = sum(t[3] for t in self.table)
        # two extra floats at the end of the gradient slab carry loss-sum / correct-count
        self.params = DeviceArray.zeros((self.total,))
        self.grads = DeviceArray.zeros((self.total,))
        self.state = []
        self._adopt()

    def _view(self, slab, off, size, shape):
        return DeviceArray(shape, np.float32, block=slab._block, offset=slab._offset + off * 4)

    def _adopt(self):
        for layer, pname, off, size, shape in self.table:
            gname = dict(_param_attr_names(layer))[pname]
            pview = self._view(self.params, off, size, shape)
            gview = self._view(self.grads, off, size, shape)
            pview.copy_from(getattr(layer, pname))
            gview.copy_from(getattr(layer, gname))
            setattr(layer, pname, pview)
            setattr(layer, gname, gview)

    def ensure_state(self, n):
        while len(self.state) < n:
            self.state.append(DeviceArray.zeros((self.total,)))

    def entries(self):
        """Yield (layer, name, view_of(slab)) for optimizers that expose per-tensor state views."""
        for layer, pname, off, size, shape in self.table:
            yield layer, pname, (lambda slab, o=off, s=size, sh=shape: self._view(slab, o, s, sh))

    def host_params(self) -> np.ndarray:
        """Packed (unpadded) FP32 copy of all parameters, table order."""
        flat = self.params.numpy()
        return np.concatenate([flat[o:o + s] for _, _, o, s, _ in self.table])

    def host_grads(self) -> np.ndarray:
        flat = self.grads.numpy()
        return np.concatenate([flat[o:o + s] for _, _, o, s, _ in self.table])

    def load_host_params(self, packed):
        flat = np.zeros((self.total,), np.float32)
        i = 0
        for _, _, o, s, _ in self.table:
            flat[o:o + s] = packed[i:i + s]
            i += s
        self.params.copy_from(flat)
