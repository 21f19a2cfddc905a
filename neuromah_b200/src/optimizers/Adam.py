"""Optimizer_Adam on the device (reference: src/optimizers/Adam.py:4-119).

Same attributes and the same pre/update/post call triple.  ``update_params(layer)``
launches nm_adam_step_f32 per parameter tensor; ``update_arena`` covers the whole
model in ONE 128-bit-vectorised launch over the flat parameter slab, applying the
update ``n_apply`` times in registers (Model registers each layer twice).
"""
import numpy as np

from ... import ops
from ...device import DeviceArray


class Optimizer_Adam:
    def __init__(self, learning_rate=0.001, decay=0., epsilon=1e-7, beta_1=0.9, beta_2=0.999, xp=np):
        if not (0 <= beta_1 < 1) or not (0 <= beta_2 < 1):
            raise ValueError("beta values must be in the interval [0 , 1)")
        self.learning_rate = learning_rate
        self.current_learning_rate = learning_rate
        self.decay = decay
        self.iterations = 0
        self.epsilon = epsilon
        self.beta_1 = beta_1
        self.beta_2 = beta_2
        self.momentums = {}
        self.caches = {}
        if xp is None:
            raise AttributeError("The Adam Optimizer expects a numpy or cupy object")
        self.xp = np

    def pre_update_params(self):
        if self.decay:
            self.current_learning_rate = self.learning_rate * (1. / (1. + self.decay * self.iterations))

    def _hp(self):
        return dict(lr=self.current_learning_rate, beta_1=self.beta_1, beta_2=self.beta_2,
                    eps=self.epsilon, t=self.iterations)

    def update_params(self, layer):
        if not hasattr(layer, "get_parameters") or not callable(getattr(layer, "get_parameters")):
            raise AttributeError("Layer must have a get_parameters method")
        parameters = layer.get_parameters()
        if not isinstance(parameters, dict):
            raise TypeError("The get_parameters method must return a dictionary")
        for original_name, values in parameters.items():
            key = f"{id(layer)}_{original_name}"
            if not isinstance(values, tuple) or len(values) != 2:
                raise ValueError("The parameter values must be in tuples")
            param, grad = values
            if not isinstance(param, DeviceArray) or not isinstance(grad, DeviceArray):
                raise ValueError("The parameters and gradients must be device arrays")
            if param.shape != grad.shape:
                raise ValueError("The parameter and gradient must have the same shape")
            if key not in self.momentums:
                self.momentums[key] = DeviceArray.zeros(param.shape)
                self.caches[key] = DeviceArray.zeros(param.shape)
            ops.adam_step(param, grad, self.momentums[key], self.caches[key], n_apply=1, **self._hp())

    def bind_arena(self, arena):
        """Point momentums/caches at two zeroed slabs laid out like the parameter arena."""
        arena.ensure_state(2)
        for (layer, name, view_of) in arena.entries():
            key = f"{id(layer)}_{name}"
            self.momentums[key] = view_of(arena.state[0])
            self.caches[key] = view_of(arena.state[1])

    def update_arena(self, arena, n_apply=1):
        ops.adam_step(arena.params, arena.grads, arena.state[0], arena.state[1], n_apply=n_apply, **self._hp())

    # -- graph-capturable variant: hyper-parameters live in a device struct advanced on the device ----------
    def _host_snapshot(self):
        return (self.learning_rate, self.decay, self.beta_1, self.beta_2, self.epsilon, self.iterations)

    def sync_device_state(self):
        """(Re)initialise the device-side nm_adam_state from the host attributes if they changed behind its back."""
        import ctypes as C
        from ..._lib import lib
        from ...device import stream
        if getattr(self, "_dev_state", None) is None:
            self._dev_state = DeviceArray.zeros((128,), np.uint8)         # sizeof(nm_adam_state) = 88
            self._dev_snapshot = None
        if self._dev_snapshot != self._host_snapshot():
            lib.nm_adam_state_init(self._dev_state.p, float(self.learning_rate), float(self.decay), float(self.beta_1),
                                   float(self.beta_2), float(self.epsilon), C.c_longlong(int(self.iterations)), stream())
            self._dev_snapshot = self._host_snapshot()      # a second call (e.g. inside a graph capture) must not launch again
        return self._dev_state

    def update_arena_dev(self, arena, n_apply=1):
        """One launch: Adam over the slab + (last block) the on-device advance of (iterations, bias corrections, decayed
        lr).  No host values are baked into the launch, so it can be replayed from a CUDA graph."""
        from ..._lib import lib
        from ...device import stream
        lib.nm_adam_step_advance_dev_f32(arena.params.p, arena.grads.p, arena.state[0].p, arena.state[1].p, arena.total,
                                         self._dev_state.p, n_apply, stream())

    def update_arena_dp(self, arena, comm, n_apply=1, pushed_from=None):
        """Data-parallel step: gradient sum over ranks (peer memory over NVLink) + Adam in ONE launch, then the
        on-device advance.  ``arena.grads`` holds the summed gradient afterwards, as after an in-place all-reduce.
        ``pushed_from``: first slab chunk the backward pass already delivered to the peers (nm_dp_push_f32)."""
        from ..._lib import lib
        from ...device import stream
        if pushed_from is None:
            lib.nm_dp_allreduce_adam_dev_f32(comm.peer, arena.params.p, arena.grads.p, arena.state[0].p, arena.state[1].p,
                                             arena.total, self._dev_state.p, n_apply, 1, stream())
        else:
            lib.nm_dp_collect_adam_dev_f32(comm.peer, arena.params.p, arena.grads.p, arena.state[0].p, arena.state[1].p,
                                           arena.total, self._dev_state.p, n_apply, 1, int(pushed_from), stream())

    def host_advance(self):
        """Host mirror of the on-device advance (keeps ``iterations`` / ``current_learning_rate`` readable)."""
        self.pre_update_params()          # the rate the step just used (Adam.py:41-47 runs before the update)
        self.iterations += 1
        self._dev_snapshot = self._host_snapshot()

    def post_update_params(self):
        self.iterations += 1

    def get_state(self):
        return {"momentums": {k: np.asarray(v) for k, v in self.momentums.items()},
                "caches": {k: np.asarray(v) for k, v in self.caches.items()},
                "iterations": self.iterations}

    def load_state(self, state):
        for k, v in state["momentums"].items():
            self.momentums[k].copy_from(v) if k in self.momentums else self.momentums.__setitem__(k, DeviceArray.from_numpy(v, np.float32))
        for k, v in state["caches"].items():
            self.caches[k].copy_from(v) if k in self.caches else self.caches.__setitem__(k, DeviceArray.from_numpy(v, np.float32))
        self.iterations = state["iterations"]

    laod_state = load_state          # the reference's spelling (Adam.py:117)
