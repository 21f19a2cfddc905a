"""Optimizer_RMSprop on the device (reference: src/optimizers/RMSprop.py:4-115)."""
from ..._lib import lib
from ...device import stream
from ._slab import NonFiniteFlag, checked_parameters, named_state


class Optimizer_RMSprop:
    def __init__(self, learning_rate=0.001, decay=0., epsilon=1e-7, rho=0.9):
        if not isinstance(learning_rate, (float, int)) or learning_rate <= 0:
            raise ValueError("learning_rate must be a positive float.")
        if not isinstance(decay, (float, int)) or decay < 0:
            raise ValueError("decay must be a non-negative float.")
        if not isinstance(epsilon, (float, int)) or epsilon <= 0:
            raise ValueError("epsilon must be a positive float.")
        if not isinstance(rho, (float, int)) or rho < 0 or rho >= 1:
            raise ValueError("rho must be a float in the range [0, 1).")
        self.learning_rate = learning_rate
        self.current_learning_rate = learning_rate
        self.decay = decay
        self.iterations = 0
        self.epsilon = epsilon
        self.rho = rho
        self.caches = {}
        self.check_gradients = True
        self._nf = NonFiniteFlag()

    def pre_update_params(self):
        if self.decay:
            self.current_learning_rate = self.learning_rate * (1. / (1. + self.decay * self.iterations))

    def _step(self, p, g, cache, n, n_apply):
        lib.nm_rmsprop_step_f32(p.p, g.p, cache.p, n, float(self.current_learning_rate), float(self.rho), float(self.epsilon),
                                int(n_apply), self._nf.ptr() if self.check_gradients else None, stream())

    def update_params(self, layer):
        for name, param, grad in checked_parameters(layer):
            self._step(param, grad, named_state(self.caches, name, param), param.size, 1)
            if self.check_gradients:
                self._nf.check(name)

    def bind_arena(self, arena):
        arena.ensure_state(1)

    def update_arena(self, arena, n_apply=1):
        self._step(arena.params, arena.grads, arena.state[0], arena.total, n_apply)

    def check_nonfinite(self):
        """Arena path: raise the reference's ValueError if the kernel flagged a NaN/Inf gradient since the last check
        (one device synchronisation: Model.train calls it once per epoch beside the loss read-back)."""
        if self.check_gradients:
            self._nf.check("arena")

    def post_update_params(self):
        self.iterations += 1
