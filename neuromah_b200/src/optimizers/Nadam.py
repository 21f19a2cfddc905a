"""Optimizer_Nadam on the device (reference: src/optimizers/Nadam.py).  Note the reference's own twists, kept here:
epsilon defaults to 1e-8 and is added to the bias-correction denominators as well as to sqrt(v_hat)."""
from ..._lib import lib
from ...device import stream
from ._slab import checked_parameters, named_state


class Optimizer_Nadam:
    def __init__(self, learning_rate=1e-3, decay=0.0, beta1=0.9, beta2=0.999, epsilon=1e-8):
        if not isinstance(learning_rate, (int, float)):
            raise TypeError("learning_rate must be a number")
        if learning_rate <= 0:
            raise ValueError("learning_rate must be positive")
        if not isinstance(decay, (int, float)):
            raise TypeError("decay must be a number")
        if decay < 0:
            raise ValueError("decay must be non-negative")
        if not isinstance(beta1, float):
            raise TypeError("beta1 must be a float")
        if not 0 < beta1 < 1:
            raise ValueError("beta1 must be in range (0, 1)")
        if not isinstance(beta2, float):
            raise TypeError("beta2 must be a float")
        if not 0 < beta2 < 1:
            raise ValueError("beta2 must be in range (0, 1)")
        if not isinstance(epsilon, float):
            raise TypeError("epsilon must be a float")
        if epsilon <= 0:
            raise ValueError("epsilon must be positive")
        self.learning_rate = learning_rate
        self.current_learning_rate = learning_rate
        self.decay = decay
        self.iterations = 0
        self.epsilon = epsilon
        self.beta1 = beta1
        self.beta2 = beta2
        self.momentums = {}
        self.caches = {}

    def pre_update_params(self):
        if self.decay:
            self.current_learning_rate = self.learning_rate * (1.0 / (1.0 + self.decay * self.iterations))

    def _step(self, p, g, m, v, n, n_apply):
        bc1 = 1.0 - self.beta1 ** (self.iterations + 1)
        bc2 = 1.0 - self.beta2 ** (self.iterations + 1)
        lib.nm_nadam_step_f32(p.p, g.p, m.p, v.p, n, float(self.current_learning_rate), float(self.beta1), float(self.beta2),
                              float(self.epsilon), float(bc1), float(bc2), int(n_apply), stream())

    def update_params(self, layer):
        for name, param, grad in checked_parameters(layer, require_nonempty=False):
            self._step(param, grad, named_state(self.momentums, name, param), named_state(self.caches, name, param),
                       param.size, 1)

    def bind_arena(self, arena):
        arena.ensure_state(2)

    def update_arena(self, arena, n_apply=1):
        self._step(arena.params, arena.grads, arena.state[0], arena.state[1], arena.total, n_apply)

    def post_update_params(self):
        self.iterations += 1
