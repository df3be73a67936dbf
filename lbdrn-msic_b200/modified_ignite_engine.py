"""Trainer / evaluator factories -- drop-in for the reference's modified_ignite_engine.py (same function names and
call signatures: `create_supervised_trainer(model, optimizer, loss_fn, device)`,
`create_supervised_evaluator(model, metrics, device)`), without the pytorch-ignite dependency.

When the engine is run over a DataLoader whose dataset is an `LBDRNDataset` (scene resident on the GPU), an epoch is
ONE persistent fused kernel launch (gather + forward + MSE + backward + Adam for every batch; lbdrn_train_steps)
instead of the reference's per-batch H2D copy + ~35 library launches + host sync (modified_ignite_engine.py:18-27).
The optimiser object supplies lr / betas / eps (so an attached StepLR keeps working); its state is not used.
Any other iterable of (x, y) batches takes the plain torch step (cold path, same arithmetic as the reference).
"""
import enum

import torch


class Events(enum.Enum):
    STARTED = 'started'
    EPOCH_STARTED = 'epoch_started'
    ITERATION_STARTED = 'iteration_started'
    ITERATION_COMPLETED = 'iteration_completed'
    EPOCH_COMPLETED = 'epoch_completed'
    COMPLETED = 'completed'


class State:
    def __init__(self):
        self.iteration = self.epoch = 0
        self.max_epochs = self.output = self.batch = None
        self.metrics = {}


class Engine:
    """Minimal event loop with the slice of ignite.engine.Engine's surface the codec uses."""

    def __init__(self, process_function, fused_epoch=None):
        self._fn, self._fused_epoch = process_function, fused_epoch
        self._handlers = {e: [] for e in Events}
        self.state = State()

    def add_event_handler(self, event, handler, *args, **kwargs):
        self._handlers[event].append((handler, args, kwargs))

    def on(self, event):
        def deco(fn):
            self.add_event_handler(event, fn)
            return fn
        return deco

    def _fire(self, event):
        for fn, a, k in list(self._handlers[event]):
            fn(self, *a, **k)

    def run(self, data, max_epochs=1):
        self.state = State()
        self.state.max_epochs = max_epochs
        self._fire(Events.STARTED)
        for _ in range(max_epochs):
            self.state.epoch += 1
            self._fire(Events.EPOCH_STARTED)
            if self._fused_epoch is not None and _fused_dataset(data) is not None:
                self._fused_epoch(self, data)
            else:
                for batch in data:
                    self.state.batch = batch
                    self._fire(Events.ITERATION_STARTED)
                    self.state.output = self._fn(self, batch)
                    self.state.iteration += 1
                    self._fire(Events.ITERATION_COMPLETED)
            self._fire(Events.EPOCH_COMPLETED)
        self._fire(Events.COMPLETED)
        return self.state


def _fused_dataset(data):
    ds = getattr(data, 'dataset', None)
    return ds if (ds is not None and hasattr(ds, 'scene')) else None


def _prepare_batch(batch, device=None, non_blocking=False):
    x, y = batch
    return x.to(device, non_blocking=non_blocking), y.to(device, non_blocking=non_blocking)


def create_supervised_trainer(model, optimizer, loss_fn, device=None, non_blocking=False,
                              prepare_batch=_prepare_batch):
    if device:
        model.to(device)
    ctx = {}

    def _update(engine, batch):                      # cold path: explicit (x, y) batches
        optimizer.zero_grad()
        model.train()
        x, y = prepare_batch(batch, device=device, non_blocking=non_blocking)
        loss = loss_fn(model(x), y)
        loss.backward()
        optimizer.step()
        return loss

    def _fused_epoch(engine, loader):
        import lbdrn_fused
        ds = _fused_dataset(loader)
        tr = ctx.get('trainer')
        if tr is None:
            g = optimizer.param_groups[0]
            if g.get('weight_decay', 0) or g.get('amsgrad', False):
                raise NotImplementedError("the fused step implements Adam without weight decay / amsgrad (encode.py:84)")
            tr = lbdrn_fused.FusedTrainer(model, ds.scene, ds.D, g['lr'], loader.batch_size, engine.state.max_epochs,
                                          flags=ds.flags, betas=g['betas'], eps=g['eps'])
            tr.begin()
            ctx['trainer'] = tr
        # same default-generator draws as iter(DataLoader(shuffle=True)) in the reference (encode.py:69-70)
        seed = lbdrn_fused.draw_loader_seeds()
        n = ds.n_pixels
        shuffle = isinstance(getattr(loader, 'sampler', None), torch.utils.data.RandomSampler)
        perm = lbdrn_fused.permutation_from_seed(n, seed) if shuffle else torch.arange(n)
        losses = tr.train_epoch(perm.to(tr.dev), optimizer.param_groups[0]['lr'])
        model.load_flat_params(tr.current_params())
        losses = losses.cpu()
        replay = bool(engine._handlers[Events.ITERATION_COMPLETED] or engine._handlers[Events.ITERATION_STARTED])
        for v in (losses if replay else losses[-1:]):
            engine.state.output = v
            if replay:
                engine._fire(Events.ITERATION_STARTED)
            engine.state.iteration += 1 if replay else len(losses)
            if replay:
                engine._fire(Events.ITERATION_COMPLETED)

    return Engine(_update, _fused_epoch)


def create_supervised_evaluator(model, metrics, device=None, non_blocking=False, prepare_batch=_prepare_batch,
                                output_transform=lambda x, y, y_pred: (y_pred, y)):
    metrics = metrics or {}

    def _inference(engine, batch):                   # cold path
        model.eval()
        with torch.no_grad():
            x, y = prepare_batch(batch, device=device, non_blocking=non_blocking)
            return output_transform(x, y, model(x))

    def _fused_epoch(engine, loader):
        import lbdrn_fused
        ds = _fused_dataset(loader)
        lbdrn_fused.draw_loader_seeds()              # evaluator.run(train_loader) also re-iterates the loader
        sc = ds.scene
        params = model.flat_params().to(sc.msb.device).contiguous()
        mse = lbdrn_fused.eval_mse(sc, params, ds.D, model.dim_hidden, model.num_layers, ds.flags, model._relu, model.w0)
        n = sc.C * sc.H * sc.W
        for m in metrics.values():
            if hasattr(m, 'add_sse'):
                m.add_sse(mse * n, n)
        engine.state.iteration += len(loader)

    evaluator = Engine(_inference, _fused_epoch)
    for name, metric in metrics.items():
        metric.attach(evaluator, name)
    return evaluator
