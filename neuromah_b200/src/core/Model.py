"""Model: the Keras-like driver, GPU resident (reference: src/core/Model.py:11-417).

add / set / finalize / train / evaluate / predict / forward / backward /
get_parameters / set_parameters / save(_parameters) / load(_parameters) keep the
reference's signatures and step order (Model.py:154-174).  Differences, all to keep
the host out of the hot loop:
  * the data set and labels are staged in HBM once per train() call (or streamed
    through pinned double buffers when ``stream_batches=True``); batch slicing
    (Model.py:150-151) is a pointer offset
  * the running loss / accuracy sums live in a device accumulator and are read
    back once per epoch
  * finalize() adopts all parameters into a flat arena (neuromah_b200/arena.py) so
    the optimizer is one launch and the gradient slab is the all-reduce buffer
  * the end-of-epoch "forward over the entire X" (Model.py:199) runs in batches.
finalize() registers every trainable layer twice exactly as the reference does
(Model.py:56-59 and :82-83) unless config.duplicate_trainable_registration
is cleared.
"""
import copy
import os
import pickle
import time

import numpy as np

from ... import config, ops
from ...arena import ParamArena
import ctypes as C

from ..._lib import lib
from ...device import (DeviceArray, PinnedBuffer, as_device, graph_capture, input_to_device, labels_to_device, new_event,
                       new_stream, pinned_address, stream, synchronize)
from ..layers.Input import Layer_Input
from ..layers.Conv_wrapepr import Layer_Conv2D


class Model:
    xp = np

    def __init__(self, device=None):
        self.layers = []
        self.softmax_classifier_output = None
        self.loss = None
        self.accuracy = None
        self.optimizer = None
        self.tensorMonitor = None
        self.arena = None
        self.comm = None                   # data-parallel communicator (neuromah_b200.parallel)
        self.epoch_end_accuracy = True     # Model.py:199-201: forward over the whole X after each epoch
        self.uint8_denominator = 255.0     # uint8 inputs mean X / 255 (the example scripts' preprocessing, done on the device)
        self.device = device

    def add(self, layer):
        layer.xp = Model.xp
        layer.model = self
        self.layers.append(layer)

    def set(self, *, loss=None, optimizer=None, accuracy=None, tensorMonitor=None):
        if loss is None:
            raise ValueError("A loss object needs to be passed to the set contructor")
        if optimizer is None:
            raise ValueError("An optimizer object needs to be passed to the set contructor")
        if accuracy is None:
            raise ValueError("A metrics object needs to be passed to the set contructor")
        self.loss, self.optimizer, self.accuracy = loss, optimizer, accuracy
        self.tensorMonitor = tensorMonitor

    def finalize(self, xp=np, mode=None):
        """``mode``: None / "fp32" = FP32-exact CUDA-core kernels (per-layer launches, reference layouts);
        "bf16" = tensor-core plan for the supported layer patterns (neuromah_b200/tc_plan.py: the MNIST CNN;
        neuromah_b200/tc_plan_gen.py: VGG-style conv stacks; neuromah_b200/tc_plan_mlp.py: multi-layer perceptrons);
        "bf16x3" = FP32-exact on the tensor cores (three bf16 planes per operand, six products per contraction):
        conv stacks (tc_plan_gen.py, which covers the MNIST CNN as well) and multi-layer perceptrons (tc_plan_mlp.py),
        same tolerance against the NumPy path as "fp32"."""
        from ..activations import Activation_Softmax
        from ..losses import Loss_CategoricalCrossentropy
        from ..losses.Activation_Softmax_Loss_CategoricalCrossentropy import \
            Activation_Softmax_Loss_CategoricalCrossentropy

        self.input_layer = Layer_Input()
        n = len(self.layers)
        if n < 2:
            raise IndexError("Model.finalize needs at least two layers")     # Model.py:67 indexes layers[i+1]
        unique = [l for l in self.layers if hasattr(l, "weights")]
        self.trainable_layers = list(unique) if config.duplicate_trainable_registration else []
        for i, layer in enumerate(self.layers):
            layer.prev = self.input_layer if i == 0 else self.layers[i - 1]
            if i < n - 1:
                layer.next = self.layers[i + 1]
            else:
                layer.next = self.loss
                self.output_layer_activation = layer
            if hasattr(layer, "weights"):
                self.trainable_layers.append(layer)
        if isinstance(self.loss, type):
            self.loss = self.loss(model=self)
            # the last layer's `next` is the loss OBJECT whose dinputs the generic backward reads (Model.py:362-366); the
            # reference leaves the class there (:74 runs before :91), which only the fused softmax path never notices
            self.layers[-1].next = self.loss
        if isinstance(self.accuracy, type):
            self.accuracy = self.accuracy(model=self)
        if isinstance(self.layers[-1], Activation_Softmax) and isinstance(self.loss, Loss_CategoricalCrossentropy):
            self.softmax_classifier_output = Activation_Softmax_Loss_CategoricalCrossentropy()
        # no consumer for the first layer's input gradient: skip the first conv dgrad
        if isinstance(self.layers[0], Layer_Conv2D):
            self.layers[0].compute_dinputs = False
        # flat arena + optimizer state slabs
        self.arena = ParamArena(unique)
        self._n_apply = {id(l): sum(1 for t in self.trainable_layers if t is l) for l in unique}
        if hasattr(self.optimizer, "bind_arena"):
            self.optimizer.bind_arena(self.arena)
        self._accum = DeviceArray.zeros((2,), np.float64)        # {sum of losses, #correct}
        # per-step read-back of the running sums (streamed training): a 16-byte device->host copy that depends only on the
        # forward pass, issued on its own stream -- a sibling branch of the backward pass when the step is a CUDA graph
        self._step_readback = None                               # PinnedBuffer while Model.train streams batches
        self._rb_stream, self._rb_fork, self._rb_join = new_stream(), new_event(), new_event()
        self.mode = mode or "fp32"
        self.plan = None
        if self.mode == "bf16":
            from ...tc_plan import TcCnnPlan
            from ...tc_plan_gen import TcStackPlan
            from ...tc_plan_mlp import TcMlpPlan
            # the MNIST pattern has its own fully fused kernels; any other VGG-style stack takes the general conv plan;
            # Dense-only stacks (the MLP of examples/test_Dense_MNIST.py) the general tcgen05 Dense plan
            for plan_cls in (TcCnnPlan, TcStackPlan, TcMlpPlan):
                if plan_cls.match(self):
                    self.plan = plan_cls(self)
                    break
            else:
                raise ValueError("mode='bf16' supports these layer patterns:\n  " +
                                 "\n  ".join(c.PATTERN for c in (TcCnnPlan, TcStackPlan, TcMlpPlan)))
        elif self.mode == "bf16x3":
            from ...tc_plan_gen import TcStackPlan
            from ...tc_plan_mlp import TcMlpPlan
            for plan_cls in (TcStackPlan, TcMlpPlan):
                if plan_cls.match(self, True) if plan_cls is TcStackPlan else plan_cls.match(self):
                    self.plan = plan_cls(self, split=True)
                    break
            else:
                raise ValueError("mode='bf16x3' supports these layer patterns:\n  " +
                                 "\n  ".join(c.PATTERN for c in (TcStackPlan, TcMlpPlan)))
        elif self.mode != "fp32":
            raise ValueError("mode must be None, 'fp32', 'bf16' or 'bf16x3'")
        # raw uint8 pixel batches: normalised on the device by 1 / uint8_denominator (examples/test_Conv2D_MNIST.py:27);
        # a tensor-core plan reads the bytes itself, the FP32-exact path converts them in Layer_Input
        self.input_layer.keep_uint8 = self.plan is not None
        self.input_layer.uint8_denominator = self.uint8_denominator
        if self.tensorMonitor:
            self.tensorMonitor.start_run(self)

    # ------------------------------------------------------------------ steps ----
    def _steps(self, n, batch_size):
        if batch_size is None:
            return 1
        steps = n // batch_size
        return steps + 1 if steps * batch_size < n else steps

    def _optimizer_step(self):
        """pre_update / update_params x layers / post_update (Model.py:171-174)."""
        opt = self.optimizer
        opt.pre_update_params()
        counts = set(self._n_apply.values())
        if hasattr(opt, "update_arena") and len(counts) == 1:
            opt.update_arena(self.arena, n_apply=counts.pop())
        else:
            for layer in self.trainable_layers:
                opt.update_params(layer)
        opt.post_update_params()

    def _device_loss_acc(self, output, labels):
        """Accumulate sum of CCE losses and #correct for this batch on the device."""
        ops.softmax_ce(output, labels, want_grad=False, accum=self._accum, holder=self)

    # ------------------------------------------------------------ CUDA-graph replay of the fused step ----
    graph_steps = True          # replay the tensor-core plan's training step from a CUDA graph when possible
    graph_fork = True           # ... with the Dense wgrad as a sibling branch of the conv backward chain

    def _graph_train_step(self, x, labels):
        """The whole step (re-pack, forward, loss sums, backward, Adam, on-device advance) as ONE graph launch.
        A graph is captured per (input pointer, label pointer, batch) the second time that triple is seen -- the
        streamed feeder alternates between two device slots, so two graphs serve a whole epoch."""
        opt = self.optimizer
        # plan.generation changes whenever the plan re-allocates its buffers (a larger batch): graphs captured before
        # that hold the old device pointers and must never be replayed
        rb = self._step_readback
        key = (x.ptr, labels.ptr, x.shape[0], self.softmax_classifier_output.grad_batch, getattr(self.plan, "generation", 0),
               None if rb is None else rb.ptr)
        graphs = self.__dict__.setdefault("_graphs", {})
        seen = self.__dict__.setdefault("_graph_seen", set())
        g = graphs.get(key)
        if g is None and (key not in seen or x.shape[0] > self.plan.B):
            if len(seen) >= 64:                            # bounded: a stream of one-off pointers must not grow this set
                seen.clear()
            seen.add(key)
            return None                                    # first sighting (buffers, lazy allocations): run eagerly
        if g is None and len(graphs) >= 16:                # bounded cache: drop the oldest graph
            lib.nm_graph_destroy(graphs.pop(next(iter(graphs))))
        dev_state = opt.sync_device_state()
        n_apply = next(iter(set(self._n_apply.values())))
        if g is None:
            gb = self.softmax_classifier_output.grad_batch
            try:
                with graph_capture(stream()) as cap:
                    self.plan.forward(x, labels, accum=self._accum, inv_batch=None if gb is None else 1.0 / gb)
                    self._readback_fork()
                    self.plan.backward(gb, fork=self.graph_fork, dp=self.comm)
                    if self.comm is not None:
                        opt.update_arena_dp(self.arena, self.comm, n_apply=n_apply, pushed_from=self.plan.pushed_from)
                    else:
                        opt.update_arena_dev(self.arena, n_apply=n_apply)
                    self._readback_join()
            except Exception as e:
                # nothing of the body has executed: drop graph replay for this model and run the step eagerly
                import warnings
                warnings.warn(f"neuromah_b200: CUDA-graph capture of the training step failed ({e}); "
                              "continuing with eager launches", RuntimeWarning)
                self.graph_steps = False
                return None
            g = graphs[key] = cap.graph
        lib.nm_graph_launch(g, stream())
        opt.host_advance()
        return self.plan.probs[0:x.shape[0]]

    def _readback_fork(self):
        """Streamed training: copy the running {loss sum, #correct} to pinned host memory as soon as the forward pass has
        updated them, on the read-back stream (captured: a sibling branch of the backward chain)."""
        if self._step_readback is not None:
            lib.nm_event_record(self._rb_fork, stream())
            lib.nm_stream_wait_event(self._rb_stream, self._rb_fork)
            lib.nm_d2h(C.c_void_p(self._step_readback.ptr), self._accum.p, 16, self._rb_stream)
            lib.nm_event_record(self._rb_join, self._rb_stream)

    def _readback_join(self):
        if self._step_readback is not None:
            lib.nm_stream_wait_event(stream(), self._rb_join)

    def _graph_capable(self):
        opt = self.optimizer
        return (self.graph_steps and self.plan is not None and (self.comm is None or self._dp_fused_ok())
                and ops.timers is None and hasattr(opt, "update_arena_dev") and len(set(self._n_apply.values())) == 1)

    def _graph_ok(self):
        return self._graph_capable() and getattr(self.plan, "B", 0) > 0

    def _dp_fused_ok(self):
        """Data parallel with the peer-memory exchange: gradient sum + Adam are one kernel (nm_dp_allreduce_adam)."""
        return (self.comm is not None and getattr(self.comm, "peer", None) is not None
                and hasattr(self.optimizer, "update_arena_dp") and len(set(self._n_apply.values())) == 1)

    def _dp_fused_step(self, pushed_from=None):
        opt = self.optimizer
        opt.sync_device_state()
        opt.update_arena_dp(self.arena, self.comm, n_apply=next(iter(set(self._n_apply.values()))), pushed_from=pushed_from)
        opt.host_advance()

    def train_step(self, batch_X, batch_y):
        """One step of Model.train's loop body with device-side loss/accuracy sums."""
        if not self._fast_path_ok():
            return self._generic_step(batch_X, batch_y)
        labels = labels_to_device(batch_y)
        if self._graph_ok() and isinstance(batch_X, DeviceArray):
            out = self._graph_train_step(batch_X, labels)
            if out is not None:
                return out
        if self.plan is not None:
            gb = self.softmax_classifier_output.grad_batch
            output = self.plan.forward(input_to_device(batch_X), labels, accum=self._accum,
                                       inv_batch=None if gb is None else 1.0 / gb)
            self._readback_fork()
            fused = self._dp_fused_ok()
            self.plan.backward(gb, comm=None if fused else self.comm, dp=self.comm if fused else None)
            if fused:
                self._dp_fused_step(self.plan.pushed_from)
                self._readback_join()
                return output
            if self.comm is not None:
                self.comm.wait()
        else:
            output = self.forward(batch_X, training=True)
            self._device_loss_acc(output, labels)
            self._readback_fork()
            self.backward(output, labels)
            if self._dp_fused_ok():
                self._dp_fused_step()
                self._readback_join()
                return output
            if self.comm is not None:
                self.comm.allreduce_grads(self.arena)
        self._optimizer_step()
        self._readback_join()
        return output

    def _fast_path_ok(self):
        from ..losses import Loss_CategoricalCrossentropy
        return self.softmax_classifier_output is not None and isinstance(self.loss, Loss_CategoricalCrossentropy)

    def _read_accum(self, reset=True):
        s = self._accum.numpy()
        if reset:
            self._accum.fill_zero()
        return float(s[0]), float(s[1])

    def train(self, X, y, *, epochs=1, batch_size=None, verbose=1, validation_data=None, xp=np,
              stream_batches=False):
        self.accuracy.init(y)
        n = len(X)
        train_steps = self._steps(n, batch_size)
        bs = n if batch_size is None else batch_size
        if not self._fast_path_ok():
            return self._train_generic(X, y, epochs=epochs, batch_size=batch_size, verbose=verbose,
                                       validation_data=validation_data)
        feeder = _BatchFeeder(X, _host_labels(y), bs, stream_batches, cache=self.__dict__.setdefault("_feeder_cache", {}))
        self.h2d_bytes = self.d2h_bytes = 0
        # streamed mode: every step's running loss / accuracy sums go back to the host (16 bytes, see _readback_fork)
        if feeder.stream and self.__dict__.get("_rb_pinned") is None:
            self._rb_pinned = PinnedBuffer((2,), np.float64)
        self._step_readback = self._rb_pinned if feeder.stream else None
        for epoch in range(1, epochs + 1):
            start_time = time.time()
            self.loss.new_pass()
            self.accuracy.new_pass()
            self._accum.fill_zero()
            if self.tensorMonitor:
                self.tensorMonitor.start_epoch(epoch)
            feeder.rewind() if feeder.stream else None
            fixed = self._graph_capable()     # staged batches go through a fixed device slot so that one graph serves all
            for step in range(train_steps):
                self.train_step(*feeder.batch(step, fixed))
                feeder.release(step, train_steps)
                if self._step_readback is not None:
                    self.d2h_bytes += 16
            if feeder.stream and (self.epoch_end_accuracy or epoch < epochs):
                feeder.prefetch_next_pass()       # keep the copy engine busy across the end-of-pass synchronisation
            loss_sum, _ = self._read_accum()
            if hasattr(self.optimizer, "check_nonfinite"):
                self.optimizer.check_nonfinite()          # NaN/Inf gradients raise ValueError as in SGD.py:80-83
            if self.comm is not None and getattr(self.comm, "peer", None) is not None:
                self.comm.check()         # a peer that never delivered its gradients raises here instead of going unnoticed
            self.h2d_bytes = feeder.h2d_bytes
            epoch_time = time.time() - start_time
            self.loss.accumulated_sum, self.loss.accumulated_count = loss_sum, n
            epoch_losses = self.loss.calculate_accumulated(include_regularization=True)
            epoch_data_loss = epoch_losses["data_loss"]
            epoch_regularization_loss = epoch_losses.get("regularization_loss", 0)
            epoch_loss = epoch_data_loss + epoch_regularization_loss
            # accuracy of the post-epoch weights over the whole training set (Model.py:199-201), batched
            if self.epoch_end_accuracy:
                feeder.rewind() if feeder.stream else None
                for step in range(train_steps):
                    xb, lb = feeder.batch(step)
                    self._device_loss_acc(self.forward(xb, training=False), lb)
                    feeder.release(step, train_steps)
                if feeder.stream and epoch < epochs:
                    feeder.prefetch_next_pass()
                _, correct = self._read_accum()
            else:
                correct = float("nan")
            epoch_accuracy = correct / n
            self.accuracy.accumulated_sum += correct
            self.accuracy.accumulated_count += n
            if self.tensorMonitor:
                self.tensorMonitor.log_scalar("loss/epoch_training", epoch_loss)
                self.tensorMonitor.log_scalar("accuracy/epoch_training", epoch_accuracy)
                for layer in self.trainable_layers:
                    self.tensorMonitor.log_layer_parameters(layer)
                self.tensorMonitor.end_epoch()
            if verbose:
                print(f"time: {epoch_time:.2f}s, acc: {epoch_accuracy:.3f}, loss: {epoch_loss:.3f} ("
                      f"data_loss: {epoch_data_loss:.3f}, reg_loss: {epoch_regularization_loss:.3f}), "
                      f"lr: {self.optimizer.current_learning_rate:.7f}")
            self.last_epoch = {"time": epoch_time, "acc": epoch_accuracy, "loss": epoch_loss,
                               "data_loss": epoch_data_loss}
            if validation_data is not None:
                self.evaluate(*validation_data, batch_size=batch_size, verbose=verbose)
        self._step_readback = None
        if self.tensorMonitor:
            self.tensorMonitor.save_logs()

    def _generic_step(self, batch_X, batch_y):
        """Model.py:154-174 for ANY loss / output activation (MSE, MAE, binary cross-entropy, a softmax without the
        fused gradient, ...): the layers, the loss and its gradient run on the device, the per-sample losses and the
        predictions are reduced on the host as the reference does."""
        output = self.forward(batch_X, training=True)
        losses = self.loss.calculate(output, batch_y, include_regularization=True)
        predictions = self.output_layer_activation.predictions(output)
        acc = self.accuracy.calculate(predictions, batch_y)
        self.last_step = {"loss": float(losses["data_loss"]), "acc": float(acc)}
        self.backward(output, batch_y)
        if self.comm is not None:
            self.comm.allreduce_grads(self.arena)
        self._optimizer_step()
        return output

    def _train_generic(self, X, y, *, epochs, batch_size, verbose, validation_data):
        """The reference's loop verbatim in structure (any loss / last layer); host-side sums."""
        train_steps = self._steps(len(X), batch_size)
        for epoch in range(1, epochs + 1):
            start_time = time.time()
            self.loss.new_pass()
            self.accuracy.new_pass()
            if self.tensorMonitor:
                self.tensorMonitor.start_epoch(epoch)
            for step in range(train_steps):
                sl = slice(None) if batch_size is None else slice(step * batch_size, (step + 1) * batch_size)
                self._generic_step(X[sl], y[sl])
            epoch_time = time.time() - start_time
            losses = self.loss.calculate_accumulated(include_regularization=True)
            epoch_loss = losses['data_loss'] + losses.get('regularization_loss', 0)
            output = self.forward(X, training=False)
            acc = self.accuracy.calculate(self.output_layer_activation.predictions(output), y)
            if self.tensorMonitor:                      # the same per-epoch hooks as the fused path (Model.py:203-223)
                self.tensorMonitor.log_scalar("loss/epoch_training", epoch_loss)
                self.tensorMonitor.log_scalar("accuracy/epoch_training", acc)
                for layer in self.trainable_layers:
                    self.tensorMonitor.log_layer_parameters(layer)
                self.tensorMonitor.end_epoch()
            if verbose:
                print(f"time: {epoch_time:.2f}s, acc: {acc:.3f}, loss: {epoch_loss:.3f} ("
                      f"data_loss: {losses['data_loss']:.3f}, reg_loss: {losses.get('regularization_loss', 0):.3f}), "
                      f"lr: {self.optimizer.current_learning_rate:.7f}")
            self.last_epoch = {"time": epoch_time, "acc": float(acc), "loss": float(epoch_loss),
                               "data_loss": float(losses['data_loss'])}
            if validation_data is not None:
                self.evaluate(*validation_data, batch_size=batch_size, verbose=verbose)
        if self.tensorMonitor:
            self.tensorMonitor.save_logs()

    def evaluate(self, X_val, y_val, *, batch_size=None, verbose=1):
        n = len(X_val)
        steps = self._steps(n, batch_size)
        bs = n if batch_size is None else batch_size
        self.loss.new_pass()
        self.accuracy.new_pass()
        if self._fast_path_ok():
            feeder = _BatchFeeder(X_val, _host_labels(y_val), bs, False)
            self._accum.fill_zero()
            for step in range(steps):
                xb, lb = feeder.batch(step)
                self._device_loss_acc(self.forward(xb, training=False), lb)
            loss_sum, correct = self._read_accum()
            self.loss.accumulated_sum, self.loss.accumulated_count = loss_sum, n
            self.accuracy.accumulated_sum, self.accuracy.accumulated_count = correct, n
        else:
            for step in range(steps):
                sl = slice(None) if batch_size is None else slice(step * batch_size, (step + 1) * batch_size)
                output = self.forward(X_val[sl], training=False)
                self.loss.calculate(output, y_val[sl])
                self.accuracy.calculate(self.output_layer_activation.predictions(output), y_val[sl])
        validation_data_loss = self.loss.calculate_accumulated()["data_loss"]
        validation_accuracy = self.accuracy.calculate_accumulated()
        if verbose:
            print(f"validation, acc: {validation_accuracy:.3f}, loss: {validation_data_loss:.3f}")
        return {"acc": float(validation_accuracy), "loss": float(validation_data_loss)}

    def predict(self, X, *, batch_size=None):
        steps = self._steps(len(X), batch_size)
        output = []
        for step in range(steps):
            batch_X = X if batch_size is None else X[step * batch_size:(step + 1) * batch_size]
            output.append(np.asarray(self.forward(batch_X, training=False)))
        return np.vstack(output)

    # --------------------------------------------------- forward / backward ----
    def forward(self, X, training):
        self.input_layer.forward(X, training)
        if self.plan is not None:
            if getattr(self.plan, "drops", None):          # a plan with fused Dropout layers needs to know the phase
                return self.plan.forward(self.input_layer.output, training=training)
            return self.plan.forward(self.input_layer.output)
        for layer in self.layers:
            layer.forward(layer.prev.output, training)
        return layer.output

    def backward(self, output, y):
        if self.plan is not None:
            # public-API backward after a label-free forward: fused (p - onehot)/B, then the plan's backward
            gb = self.softmax_classifier_output.grad_batch
            self.softmax_classifier_output.backward(output, y)
            bsz = output.shape[0]
            self.plan.dlogits[0:bsz].copy_from(self.softmax_classifier_output.dinputs)
            self.plan.dl_hilo_valid = False          # the bf16 hi/lo copy must be rebuilt from these dlogits
            self.plan.backward(gb)
            return
        if self.softmax_classifier_output is not None:
            self.softmax_classifier_output.backward(output, y)
            self.layers[-1].dinputs = self.softmax_classifier_output.dinputs
            for layer in reversed(self.layers[:-1]):
                layer.backward(layer.next.dinputs)
            return
        self.loss.backward(output, y)
        for layer in reversed(self.layers):
            layer.backward(layer.next.dinputs)

    # ------------------------------------------------------------ parameters ----
    def get_parameters(self):
        return [layer.get_parameters() for layer in self.trainable_layers]

    def set_parameters(self, parameters):
        """Accepts the reference's list of {"weights": (W, dW), "biases": (b, db)} dicts
        (the reference's own ``layer.set_parameters(*parameter_set)`` unpacks the dict KEYS,
        Model.py:380-381, and cannot load what save_parameters wrote; we read the values)."""
        for parameter_set, layer in zip(parameters, self.trainable_layers):
            if isinstance(parameter_set, dict):
                w, b = parameter_set["weights"], parameter_set["biases"]
                w = w[0] if isinstance(w, tuple) else w
                b = b[0] if isinstance(b, tuple) else b
                layer.set_parameters(np.asarray(w), np.asarray(b))
            else:
                layer.set_parameters(*parameter_set)

    def save_parameters(self, path):
        host = [{k: (np.asarray(p), np.asarray(g)) for k, (p, g) in d.items()} for d in self.get_parameters()]
        with open(path, "wb") as f:
            pickle.dump(host, f)

    def load_parameters(self, path):
        with open(path, "rb") as f:
            self.set_parameters(pickle.load(f))

    def save(self, path):
        """Whole-model checkpoint (Model.py:396-411): layers + parameters + optimizer state, as a host-side description
        (see neuromah_b200/checkpoint.py for why it is not a pickle of the live object)."""
        from ... import checkpoint
        checkpoint.save(self, path)

    @staticmethod
    def load(path):
        """Rebuild and finalize the model stored by ``save`` (Model.py:413-417)."""
        from ... import checkpoint
        return checkpoint.load(path)


def _host_labels(y):
    """Class ids from sparse labels or one-hot rows (argmax, as the reference's fused gradient does)."""
    if isinstance(y, DeviceArray):
        return y if (y.dtype == np.int32 and y.ndim == 1) else _host_labels(y.numpy())
    y = np.asarray(y)
    if y.ndim == 2:
        return np.argmax(y, axis=1).astype(np.int32)
    if y.ndim != 1:
        raise ValueError("y_true must be either sparse labels or one-hot encoded")
    return y.astype(np.int32)


class _BatchFeeder:
    """Supplies batch ``step`` as (X DeviceArray, labels DeviceArray).

    staged   (default)       : the whole data set and the labels are uploaded once; a batch is a
                               pointer offset (Model.py:150-151 becomes free).
    streamed (stream_batches): every step's inputs are copied host->device inside the step, on a
                               separate copy stream, double-buffered so the upload of step k+1
                               overlaps the compute of step k.  Page-locked sources
                               (device.PinnedBuffer.array) are DMA'd in place; pageable sources are
                               staged through two pinned buffers first.
    """

    def __init__(self, X, labels_host, bs, stream_batches, cache=None):
        self.bs, self.n = bs, len(X)
        self.stream = bool(stream_batches) and not isinstance(X, DeviceArray)
        self.h2d_bytes = 0
        self._cache = cache
        if not self.stream:
            self.X = input_to_device(X)
            self.labels = labels_to_device(labels_host)
            return
        self.X = np.asarray(X)
        if self.X.dtype not in (np.dtype(np.float32), np.dtype(np.uint8)):     # uint8 pixels travel as bytes
            self.X = self.X.astype(np.float32)
        dt = self.X.dtype
        self.labels_host = np.ascontiguousarray(labels_host, dtype=np.int32)
        shape = (bs,) + self.X.shape[1:] + (str(dt),)
        # the two device slots, their pinned label staging, the copy stream and the events persist across train() calls
        # (same pointers => the CUDA graphs captured for them stay valid)
        res = None if cache is None else cache.get(shape)
        if res is None:
            ns = self.SLOTS
            res = dict(dev=[DeviceArray(shape[:-1], dt) for _ in range(ns)],
                       dev_lab=[DeviceArray((bs,), np.int32) for _ in range(ns)],
                       copy_stream=new_stream(), uploaded=[new_event() for _ in range(ns)],
                       consumed=[new_event() for _ in range(ns)], labels_ready=new_event(), pin_x=None,
                       pin_labels=None, dev_labels=None)
            if cache is not None:
                cache[shape] = res
        else:
            synchronize()                                  # a previous pass may still be reading the slots
        self._res = res
        self.dev, self.dev_lab = res["dev"], res["dev_lab"]
        self.pin_x = res["pin_x"]
        self.copy_stream, self.uploaded, self.consumed = res["copy_stream"], res["uploaded"], res["consumed"]
        self.issued = -1          # last step (of the current pass) whose upload has been enqueued
        self.uploads = 0          # uploads ever enqueued; slot = uploads % SLOTS
        self.slot_of = {}
        self.prefetched = False   # step 0 of the NEXT pass is already on its way (prefetch_next_pass)
        # The labels of the whole data set go up in ONE copy (they are 4 bytes per sample); every step then fills its
        # slot's label buffer with a 16 KB device-to-device copy on the compute stream.  A per-step 16 KB host copy
        # behind each 12.8 MB image copy cost the copy engine ~7 us of turnaround per step (scratch/e2e_probe3.py).
        if res["pin_labels"] is None or res["pin_labels"].array.shape[0] < self.n:
            res["pin_labels"] = PinnedBuffer((self.n,), np.int32)
            res["dev_labels"] = DeviceArray((self.n,), np.int32)
        self.pin_labels, self.dev_labels, self.labels_ready = res["pin_labels"], res["dev_labels"], res["labels_ready"]
        self.pin_labels.array[:self.n] = self.labels_host
        lib.nm_h2d(self.dev_labels.p, C.c_void_p(self.pin_labels.ptr), 4 * self.n, self.copy_stream)
        lib.nm_event_record(self.labels_ready, self.copy_stream)
        self.h2d_bytes += 4 * self.n

    # Device slots (double buffering).  Three slots were measured too (NM_FEED_SLOTS=3): 0.259 ms per step against 0.254
    # with two on the B200 box, so two it is.
    SLOTS = int(os.environ.get("NM_FEED_SLOTS", "2"))

    def rewind(self):
        """Start a new pass over the data (new epoch / evaluation sweep)."""
        if self.prefetched:
            self.prefetched = False       # step 0 of this pass was enqueued at the end of the previous one
        else:
            self.issued = -1

    def prefetch_next_pass(self):
        """Enqueue the upload of step 0 of the next pass over the same data NOW, so the copy engine keeps going while
        the host synchronises on the end-of-pass loss read (otherwise every pass boundary drains the pipeline: one
        copy time, ~0.24 ms, per epoch)."""
        if self.stream and not self.prefetched:
            self.issued = -1
            self._upload(0)
            self.prefetched = True

    def _upload(self, step):
        slot = self.uploads % self.SLOTS
        lo, hi = step * self.bs, min((step + 1) * self.bs, self.n)
        rows = hi - lo
        if self.uploads >= self.SLOTS:
            lib.nm_stream_wait_event(self.copy_stream, self.consumed[slot])   # device slot consumed by its previous user
            if self.pin_x is not None:
                lib.nm_event_sync(self.uploaded[slot])                        # pinned staging slot drained
        src = self.X[lo:hi]
        addr = pinned_address(src)
        if addr is None:                              # pageable: stage through pinned memory
            if self.pin_x is None:
                self.pin_x = self._res["pin_x"] = [PinnedBuffer((self.bs,) + self.X.shape[1:], self.X.dtype)
                                                   for _ in range(self.SLOTS)]
                if self.uploads >= self.SLOTS:
                    lib.nm_event_sync(self.uploaded[slot])
            self.pin_x[slot].array[:rows] = src
            addr = self.pin_x[slot].ptr
        xb = self.dev[slot][0:rows]
        lib.nm_h2d(xb.p, C.c_void_p(addr), xb.nbytes, self.copy_stream)
        # the slot's labels: a 16 KB device-to-device copy out of the label array that went up once, on the COPY stream (it
        # is ordered behind that upload there), so the compute stream carries nothing but the step between two events
        lib.nm_d2d(self.dev_lab[slot].p, C.c_void_p(self.dev_labels.ptr + 4 * lo), 4 * rows, self.copy_stream)
        lib.nm_event_record(self.uploaded[slot], self.copy_stream)
        self.h2d_bytes += xb.nbytes
        self.issued = step
        self.uploads += 1
        self.slot_of[step] = slot

    def _fixed_slot(self):
        """One persistent (inputs, labels) device slot per batch shape: a staged batch is a different pointer every
        step, and a CUDA graph is bound to the pointers it was captured with -- copying the batch into a fixed slot
        (2 x 12.8 MB of HBM traffic at batch 4096, ~5 us) lets ONE graph serve every step of every epoch."""
        key = ("staged", (self.bs,) + tuple(self.X.shape[1:]), str(self.X.dtype))
        store = self._cache if self._cache is not None else self.__dict__.setdefault("_own", {})
        if key not in store:
            store[key] = (DeviceArray((self.bs,) + tuple(self.X.shape[1:]), self.X.dtype), DeviceArray((self.bs,), np.int32))
        return store[key]

    def batch(self, step, fixed=False):
        lo, hi = step * self.bs, min((step + 1) * self.bs, self.n)
        if not self.stream:
            if not fixed:
                return self.X[lo:hi], self.labels[lo:hi]
            sx, sl = self._fixed_slot()
            xb, lb = sx[0:hi - lo], sl[0:hi - lo]
            xb.copy_from(self.X[lo:hi])
            lb.copy_from(self.labels[lo:hi])
            return xb, lb
        while self.issued < step:
            self._upload(self.issued + 1)
        slot = self.slot_of[step]
        lib.nm_stream_wait_event(stream(), self.uploaded[slot])
        return self.dev[slot][0:hi - lo], self.dev_lab[slot][0:hi - lo]

    def release(self, step, n_steps):
        """Call after step's kernels are enqueued: marks its slot reusable and keeps SLOTS-1 uploads in flight."""
        if not self.stream:
            return
        lib.nm_event_record(self.consumed[self.slot_of[step]], stream())
        while self.issued < min(step + self.SLOTS - 1, n_steps - 1):
            self._upload(self.issued + 1)
