"""TensorMonitor: per-epoch JSON training log (reference: src/utils/TensorMonitor.py:7-136).

Same hooks (start_run / start_epoch / log_scalar / log_histogram / log_layer_parameters / end_epoch / save_logs),
same file name pattern (``<log_dir>/run-<timestamp>.json``) and the same JSON shape
(``metadata{start_time, layers, optimizer, loss_function, accuracy_metric, end_time}``, ``epochs[{epoch, time_start,
metrics{scalars}, parameters{histograms{tag: {histogram, bin_edges}}}, time_end, time_elapsed}]``), so the reference's
log viewer reads these files.  Parameters live in HBM here: a histogram of a device tensor is computed ON THE DEVICE
(min/max pass + integer bin counts with NumPy's own index arithmetic, nm_minmax_f32 / nm_histogram_f32: the counts equal
np.histogram's exactly) and only 2 floats + `bins` counters cross PCIe; the metadata records shapes instead of
``tolist()``-ing every weight into the JSON (TensorMonitor.py:42-44).
"""
import json
import os
import time

import numpy as np

_SIMPLE = (int, float, str, bool, type(None))


def _describe(obj):
    """JSON-friendly description of an object's attributes (scalars kept, tensors summarised, the rest str())."""
    out = {}
    for key, value in vars(obj).items():
        if key.startswith("_") or key in ("prev", "next", "model", "xp"):
            continue
        if isinstance(value, _SIMPLE):
            out[key] = value
        elif isinstance(value, (list, tuple)) and all(isinstance(v, _SIMPLE) for v in value):
            out[key] = list(value)
        elif hasattr(value, "shape") and hasattr(value, "dtype"):
            out[key] = {"shape": list(value.shape), "dtype": str(value.dtype)}
        else:
            out[key] = type(value).__name__
    return out


def _on_device(values):
    from ...device import DeviceArray
    return isinstance(values, DeviceArray) and values.dtype == np.float32 and values.size > 0


def device_histogram(values, bins=10):
    """np.histogram(values.flatten(), bins) for a float32 DeviceArray, computed on the device (same counts, same edges)."""
    import ctypes as C
    from ..._lib import lib
    from ...device import DeviceArray, stream
    mm = DeviceArray((2,), np.float32)
    lib.nm_minmax_f32(values.p, values.size, mm.p, stream())
    lo, hi = (float(v) for v in mm.numpy())
    if not (np.isfinite(lo) and np.isfinite(hi)):
        raise ValueError(f"autodetected range of [{lo}, {hi}] is not finite")          # np.histogram's own error
    first, last = (lo - 0.5, hi + 0.5) if lo == hi else (lo, hi)                       # numpy widens an empty range
    counts = DeviceArray((int(bins),), np.uint64)
    lib.nm_histogram_f32(values.p, values.size, C.c_double(first), C.c_double(last), int(bins), counts.p, stream())
    return counts.numpy().astype(np.int64), np.linspace(first, last, int(bins) + 1)


class TensorMonitor:
    def __init__(self, log_dir='tensor_logs'):
        self.log_dir = log_dir
        os.makedirs(self.log_dir, exist_ok=True)
        self.run_id = time.strftime("%Y%m%d-%H%M%S")
        self.log_file_path = os.path.join(self.log_dir, f'run-{self.run_id}.json')
        self.log_data = {'metadata': {'start_time': time.strftime("%Y-%m-%d %H:%M:%S"), 'layers': [], 'optimizer': '',
                                      'loss_function': '', 'accuracy_metric': ''},
                         'epochs': []}
        self.epoch_data = None
        self.epoch_count = 0

    def start_run(self, model):
        meta = self.log_data['metadata']
        meta['layers'] = [{'name': type(layer).__name__, 'config': _describe(layer)}
                          for layer in model.layers if hasattr(layer, 'weights')]
        if model.optimizer:
            meta['optimizer'] = type(model.optimizer).__name__ + " " + str(_describe(model.optimizer))
        if model.loss:
            meta['loss_function'] = type(model.loss).__name__
        if model.accuracy:
            meta['accuracy_metric'] = type(model.accuracy).__name__

    def start_epoch(self, epoch_num):
        self.epoch_count = epoch_num
        self.epoch_data = {'epoch': epoch_num, 'time_start': time.time(), 'metrics': {}, 'parameters': {}}

    def log_scalar(self, tag, value):
        if self.epoch_data is not None:
            self.epoch_data['metrics'].setdefault('scalars', {})[tag] = float(value)

    def log_histogram(self, tag, values, bins=10):
        if self.epoch_data is not None:
            hist, edges = device_histogram(values, bins) if _on_device(values) else np.histogram(np.asarray(values).ravel(), bins=bins)
            self.epoch_data['parameters'].setdefault('histograms', {})[tag] = {'histogram': hist.tolist(),
                                                                                'bin_edges': edges.tolist()}

    def log_layer_parameters(self, layer):
        """weights / biases and -- where the layer names them so (Layer_Conv2D) -- weight_gradients / bias_gradients
        (TensorMonitor.py:108-117; Dense's dweights/dbiases are not logged by the reference either)."""
        if not hasattr(layer, 'weights'):
            return
        name = type(layer).__name__
        self.log_histogram(f'{name}/weights', layer.weights)
        if hasattr(layer, 'weight_gradients'):
            self.log_histogram(f'{name}/dweights', layer.weight_gradients)
        if hasattr(layer, 'biases'):
            self.log_histogram(f'{name}/biases', layer.biases)
            if hasattr(layer, 'bias_gradients'):
                self.log_histogram(f'{name}/dbiases', layer.bias_gradients)

    def end_epoch(self):
        if self.epoch_data is not None:
            self.epoch_data['time_end'] = time.time()
            self.epoch_data['time_elapsed'] = self.epoch_data['time_end'] - self.epoch_data['time_start']
            self.log_data['epochs'].append(self.epoch_data)
            self.epoch_data = None

    def save_logs(self):
        self.log_data['metadata']['end_time'] = time.strftime("%Y-%m-%d %H:%M:%S")
        with open(self.log_file_path, 'w') as f:
            json.dump(self.log_data, f, indent=4, default=str)
        print(f"TensorMonitor logs saved to: {self.log_file_path}")
