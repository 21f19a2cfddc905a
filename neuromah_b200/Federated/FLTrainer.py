"""FLTrainer (reference: Federated/FLTrainer.py:3-31), with REAL local training.

The reference "simulates" local training by adding N(0, 0.01) noise to a weight list and
reporting 100 samples (:25-31).  The API shell is kept -- construct from
``{"weights": [...]}``, ``train_local_model()`` returns ``({"weights": list}, num_samples)`` --
but when a Model and a data shard are attached the body runs E local steps of the B200
training path and returns the trained weights; aggregation is Federated.utils.
"""
import numpy as np


class FLTrainer:
    def __init__(self, global_model, model=None, data=None, batch_size=None, local_epochs=1):
        self.model_weights = np.array(global_model["weights"])
        self.model, self.data = model, data
        self.batch_size, self.local_epochs = batch_size, local_epochs

    def train_local_model(self):
        if self.model is None:           # the reference's simulated update
            new_weights = self.model_weights + np.random.normal(0, 0.01, size=self.model_weights.shape)
            return {"weights": new_weights.tolist()}, 100
        X, y = self.data
        self.model.arena.load_host_params(np.asarray(self.model_weights, dtype=np.float32))
        self.model.train(X, y, epochs=self.local_epochs, batch_size=self.batch_size, verbose=0)
        return {"weights": self.model.arena.host_params().tolist()}, len(X)
