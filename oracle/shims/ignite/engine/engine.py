"""TEST INFRASTRUCTURE ONLY -- minimal `ignite.engine.Engine` (see package docstring)."""
import enum


class Events(enum.Enum):
    STARTED = "started"
    EPOCH_STARTED = "epoch_started"
    ITERATION_STARTED = "iteration_started"
    ITERATION_COMPLETED = "iteration_completed"
    EPOCH_COMPLETED = "epoch_completed"
    COMPLETED = "completed"


class State:
    def __init__(self):
        self.iteration = 0
        self.epoch = 0
        self.max_epochs = None
        self.output = None
        self.batch = None
        self.metrics = {}
        self.dataloader = None


class Engine:
    def __init__(self, process_function):
        self._process_function = process_function
        self._handlers = {e: [] for e in Events}
        self.state = State()

    def add_event_handler(self, event, handler, *args, **kwargs):
        self._handlers[event].append((handler, args, kwargs))

    def on(self, event, *args, **kwargs):
        def deco(fn):
            self.add_event_handler(event, fn, *args, **kwargs)
            return fn
        return deco

    def _fire(self, event):
        for handler, args, kwargs in list(self._handlers[event]):
            handler(self, *args, **kwargs)

    def run(self, data, max_epochs=1):
        self.state = State()
        self.state.max_epochs = max_epochs
        self.state.dataloader = data
        self._fire(Events.STARTED)
        while self.state.epoch < max_epochs:
            self.state.epoch += 1
            self._fire(Events.EPOCH_STARTED)
            for batch in data:            # a fresh iterator per epoch
                self.state.batch = batch
                self._fire(Events.ITERATION_STARTED)
                self.state.output = self._process_function(self, batch)
                self.state.iteration += 1
                self._fire(Events.ITERATION_COMPLETED)
            self._fire(Events.EPOCH_COMPLETED)
        self._fire(Events.COMPLETED)
        return self.state
