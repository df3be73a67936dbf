"""Evaluation metric of the encoder: MSE over every prediction seen since `reset()` (reference
LBDRNperformance.py:13-21).  Accumulates sums instead of concatenating tensors; on the fused path the evaluator
fills it from the device-side squared-error kernel.  No ignite dependency: `attach` works with the Engine of
modified_ignite_engine.py (and duck-types ignite's when that package is present)."""
import torch


class LBDRNPerformance:
    def __init__(self):
        self.reset()

    def reset(self):
        self._sse, self._n = 0.0, 0

    def update(self, output):
        y_pred, y = output
        self._sse += float(((y_pred.double() - y.double()) ** 2).sum())
        self._n += y.numel()

    def add_sse(self, sse, n):
        self._sse += float(sse)
        self._n += int(n)

    def compute(self):
        return {'MSE': self._sse / max(1, self._n)}

    def attach(self, engine, name):
        from modified_ignite_engine import Events
        engine.add_event_handler(Events.EPOCH_STARTED, lambda e: self.reset())
        engine.add_event_handler(Events.ITERATION_COMPLETED, lambda e: self.update(e.state.output)
                                 if isinstance(e.state.output, (tuple, list)) else None)

        def done(e):
            res = self.compute()
            e.state.metrics.update(res)
            e.state.metrics[name] = res
        engine.add_event_handler(Events.EPOCH_COMPLETED, done)
