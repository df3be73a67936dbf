"""TEST INFRASTRUCTURE ONLY -- minimal `ignite.metrics.metric.Metric` (see package docstring)."""
from collections.abc import Mapping

from ..engine.engine import Events


class Metric:
    def __init__(self, output_transform=lambda x: x, device=None):
        self._output_transform = output_transform
        self.reset()

    def reset(self):
        raise NotImplementedError

    def update(self, output):
        raise NotImplementedError

    def compute(self):
        raise NotImplementedError

    def started(self, engine):
        self.reset()

    def iteration_completed(self, engine):
        self.update(self._output_transform(engine.state.output))

    def completed(self, engine, name):
        result = self.compute()
        if isinstance(result, Mapping):
            for key, value in result.items():
                engine.state.metrics[key] = value
        engine.state.metrics[name] = result

    def attach(self, engine, name):
        engine.add_event_handler(Events.EPOCH_STARTED, self.started)
        engine.add_event_handler(Events.ITERATION_COMPLETED, self.iteration_completed)
        engine.add_event_handler(Events.EPOCH_COMPLETED, self.completed, name)
