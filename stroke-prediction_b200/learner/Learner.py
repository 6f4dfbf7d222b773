"""Training procedure skeleton (API of the reference's learner/Learner.py:16-226).

Hot path = ``train_batch`` / ``validate_batch`` (Learner.py:116-142): forward through the fused engine, the
subclass's ``loss_step``, ``zero_grad`` / ``backward`` / fused Adam step, one scalar D2H for the loss.  A
``torch.optim.Adam`` handed in by a reference-style script is adopted by :class:`FusedAdam` (same param_groups and
state objects).  Plotting (matplotlib) and jsonpickle history files are host-side orchestration outside the hot-path
scope; they are used when those packages are importable and skipped otherwise.
"""
import copy
import json
from abc import abstractmethod

import numpy
import torch

from ..common.dto.Dto import Dto
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.MetricMeasuresDto import MetricMeasuresDto
from ..common.inference.Inference import Inference
from ..optim import FusedAdam


def _history_to_json(history):
    def enc(d):
        return {k: (enc(v) if isinstance(v, Dto) else (None if v is None else float(v))) for k, v in d}
    return json.dumps({phase: [enc(m) for m in metrics] for phase, metrics in history.items()})


def _history_from_json(text):
    def dec(d):
        b = lambda x: MetricMeasuresDtoInit.BinaryMeasuresDto(x['dc'], x['hd'], x['assd'], x['precision'],
                                                                x['sensitivity'], x['specificity'])
        return MetricMeasuresDto(d['loss'], b(d['core']), b(d['penu']), b(d['lesion']))
    raw = json.loads(text)
    return {phase: [dec(m) for m in metrics] for phase, metrics in raw.items()}


class Learner(Inference):
    FNB_MODEL = 'model'
    FNB_OPTIM = 'optimizer'
    FNB_TRAIN = 'training'
    FNB_PLOTS = 'plots'
    FNB_IMAGE = 'visual'
    FNB_MARKS = '_learner'
    EXT_MODEL = '.model'
    EXT_OPTIM = '.optim'
    EXT_TRAIN = '.json'
    EXT_IMAGE = '.png'

    def __init__(self, dataloader_training, dataloader_validation, model, optimizer, scheduler, n_epochs: int,
                 path_previous_base: str = None, path_outputs_base: str = '/tmp/stroke-prediction'):
        Inference.__init__(self, model)

        assert dataloader_training is None or dataloader_training.batch_size > 1, \
            'For normalization layers batch_size > 1 is required.'
        self._dataloader_training = dataloader_training
        self._dataloader_validation = dataloader_validation
        self._optimizer = FusedAdam.from_torch(optimizer) if isinstance(optimizer, torch.optim.Adam) else optimizer
        self._scheduler = scheduler
        self._n_epochs = n_epochs
        self._grad_sync = None   # data-parallel hook (parallel.GradientAllReduce), called between backward and step
        if isinstance(self._optimizer, FusedAdam) and self.is_cuda:
            # gradients accumulate straight into one flat buffer and are cleared inside the Adam kernel
            self._optimizer.attach_grad_sink()
            self._optimizer.fuse_zero_grad = True

        self._path_outputs_base = path_outputs_base
        self._path_previous_base = path_previous_base

        if path_previous_base is not None:
            self.load_model(self.is_cuda)
            self.load_training()
            print('Continue training', path_previous_base, '...')
        else:
            self._metric_dtos = {'training': [], 'validate': []}
        assert len(self._metric_dtos['training']) == len(self._metric_dtos['validate']), 'Incomplete training data!'

    def path(self, mode: str, type: str, suffix: str = ''):
        base_path = {'load': self._path_previous_base, 'save': self._path_outputs_base}.get(mode)
        ext = {self.FNB_MODEL: self.EXT_MODEL, self.FNB_OPTIM: self.EXT_OPTIM, self.FNB_TRAIN: self.EXT_TRAIN,
               self.FNB_PLOTS: self.EXT_IMAGE, self.FNB_IMAGE: self.EXT_IMAGE}.get(type)
        if base_path is None or ext is None:
            return None
        return base_path + self.FNB_MARKS + suffix + ext

    @abstractmethod
    def loss_step(self, dto: Dto, epoch):
        pass

    def get_start_epoch(self):
        return 0

    def get_start_min_loss(self):
        return numpy.inf

    def load_model(self, cuda=True):
        """Learner.py:96-101.  The reference replaces ``self._model`` by the unpickled module although the optimizer it
        was handed still holds the parameters of the script-built one, so a "continued" training never updates the loaded
        weights (defect D11).  When the checkpoint has the structure of the model the optimizer was built on, its
        tensors are copied INTO that model (optimizer / gradient-sink links stay valid: the resume really resumes);
        otherwise the module is replaced like in the reference."""
        model = torch.load(self.path('load', self.FNB_MODEL), weights_only=False)
        self._model = self.adopt_checkpoint(getattr(self, '_model', None), model, cuda)

    @staticmethod
    def adopt_checkpoint(live, loaded, cuda=True):
        if live is not None and type(live) is type(loaded):
            have, want = live.state_dict(), loaded.state_dict()
            if have.keys() == want.keys() and all(have[k].shape == want[k].shape for k in have):
                with torch.no_grad():
                    for k, t in have.items():
                        t.copy_(want[k])
                from .. import engine
                engine.bump_weights_epoch()
                return live
        return loaded.cuda() if cuda else loaded

    def load_training(self):
        path_training = self.path('load', self.FNB_TRAIN)
        path_optimizer = self.path('load', self.FNB_OPTIM)
        print('Loading:', path_training, path_optimizer)
        self._optimizer.load_state_dict(torch.load(path_optimizer, weights_only=False))
        with open(path_training, 'r') as fp:
            text = fp.read()
        try:
            import jsonpickle
            self._metric_dtos = jsonpickle.decode(text)
        except ImportError:
            self._metric_dtos = _history_from_json(text)

    def save_training(self):
        torch.save(self._optimizer.state_dict(), self.path('save', self.FNB_OPTIM))
        try:
            import jsonpickle
            text = jsonpickle.encode(self._metric_dtos)
        except ImportError:
            text = _history_to_json(self._metric_dtos)
        with open(self.path('save', self.FNB_TRAIN), 'w') as fp:
            fp.write(text)

    @staticmethod
    def cpu_snapshot(module):
        """A CPU deep copy of `module` for checkpointing (the reference pickles ``model.cpu()`` and moves it back,
        Learner.py:112-114).  The LIVE module is never moved: ``Module._apply`` re-binds the storage of every
        ``param.grad``, which would silently cut the gradients loose from the flat gradient buffer the backward kernels,
        the all-reduce and the fused Adam share (engine.GradSink).  Gradients are not part of a checkpoint."""
        params = list(module.parameters())
        grads = [p.grad for p in params]
        for p in params:
            p.grad = None
        try:
            clone = copy.deepcopy(module)
        finally:
            for p, g in zip(params, grads):
                p.grad = g
        return clone.cpu()

    def save_model(self, suffix=''):
        torch.save(self.cpu_snapshot(self._model), self.path('save', self.FNB_MODEL, suffix))

    # ------------------------------------------------------------------------------------------ hot path
    OVERLAP_METRICS = True      # per-batch metrics on a side stream, concurrent with backward + optimizer step

    def train_batch(self, batch: dict, epoch) -> MetricMeasuresDto:
        dto = self.inference_step(batch)
        loss = self.loss_step(dto, epoch)

        # The per-batch metrics (Learner.py:127) only read the forward outputs: their kernels are launched on a side stream
        # before the backward pass and their single device-to-host read happens after the optimizer step.
        pending = side = None
        if self.OVERLAP_METRICS and loss.is_cuda:
            from ..common import metrics
            main = torch.cuda.current_stream()
            side = self.__dict__.get('_metric_stream')
            if side is None:
                side = self._metric_stream = torch.cuda.Stream()
            side.wait_stream(main)
            with torch.cuda.stream(side), metrics.deferred() as pending:
                batch_metrics = self.batch_metrics_step(dto, epoch)

        self._optimizer.zero_grad()
        loss.backward()
        if self._grad_sync is not None:
            self._grad_sync()
        self._optimizer.step()

        if pending is None:
            batch_metrics = self.batch_metrics_step(dto, epoch)
        else:
            with torch.cuda.stream(side):
                pending.finish()
            torch.cuda.current_stream().wait_stream(side)       # the forward outputs may be recycled from here on
        batch_metrics.loss = float(loss.detach().reshape(-1)[0].cpu())   # the step's single D2H read on the main stream

        del loss
        del dto
        return batch_metrics

    # keys of a batch dict whose tensors are only consumed on the device (the clinical scalars are read on the host)
    PREFETCH_KEYS = ('images', 'labels')

    def prefetched(self, loader):
        """Iterate `loader` one batch ahead: the host-to-device copy of batch i + 1 (its `images` / `labels` tensors; pinned host
        memory makes the copy asynchronous) is issued on a copy stream while batch i trains, so the epoch loop of `run_training`
        (Learner.py:161-163) never waits for PCIe.  `inference_step` takes tensors that are already on the device as they are."""
        dev = self.device if self.is_cuda else None
        if dev is None:
            for batch in loader:
                yield batch
            return
        copy = self.__dict__.get('_copy_stream')
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(dev)
            self._copy_bufs = {}                    # (key, slot) -> persistent device staging tensor (no allocator traffic per step)
        bufs = self._copy_bufs
        free = [None, None]                         # main-stream event: the step that consumed this slot has been queued

        def stage(batch, slot):
            out = dict(batch)
            if free[slot] is not None:
                copy.wait_event(free[slot])         # do not overwrite a buffer the main stream still reads
            with torch.cuda.stream(copy):
                for k in self.PREFETCH_KEYS:
                    v = out.get(k)
                    if torch.is_tensor(v) and not v.is_cuda and v.numel() >= 1 << 16:
                        b = bufs.get((k, slot))
                        if b is None or b.shape != v.shape or b.dtype != v.dtype:
                            b = bufs[(k, slot)] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                        b.copy_(v, non_blocking=v.is_pinned())
                        out[k] = b
            return out, copy.record_event()

        it = iter(loader)
        try:
            nxt = stage(next(it), 0)
        except StopIteration:
            return
        slot = 0
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = stage(next(it), slot ^ 1)
            except StopIteration:
                nxt = None
            main = torch.cuda.current_stream(dev)
            main.wait_event(ev)
            yield cur                               # the caller trains on `cur` (main stream) before asking for the next batch
            free[slot] = torch.cuda.current_stream(dev).record_event()
            slot ^= 1

    def train_batches(self, loader, epoch):
        """`train_batch` over `loader` with the next batch's H2D copy overlapped (what `run_training` runs per epoch)."""
        for batch in self.prefetched(loader):
            yield self.train_batch(batch, epoch)

    def enable_data_parallel(self, process_group=None):
        """Batch-sharded training over torch.distributed (one process per GPU): gradients are summed with one
        all-reduce of the flat gradient buffer and averaged inside the Adam kernel (DDP semantics)."""
        from ..parallel import GradientAllReduce
        self._grad_sync = GradientAllReduce(self._optimizer, process_group)
        return self._grad_sync

    def validate_batch(self, batch: dict, epoch) -> MetricMeasuresDto:
        with torch.no_grad():   # the reference builds a graph here for nothing
            dto = self.inference_step(batch)
            loss = self.loss_step(dto, epoch)
            batch_metrics = self.batch_metrics_step(dto, epoch)
            batch_metrics.loss = float(loss.detach().reshape(-1)[0].cpu())
        del loss
        del dto
        return batch_metrics

    def batch_metrics_step(self, dto: Dto, epoch) -> MetricMeasuresDto:
        return MetricMeasuresDtoInit.init_dto()

    def print_epoch(self, epoch, phase, epoch_metrics: MetricMeasuresDto):
        pass

    def plot_epoch(self, plotter, epochs):
        pass

    def visualize_epoch(self, epoch):
        pass

    def adapt_lr(self, epoch):
        if self._scheduler is not None:
            self._scheduler.step()

    def adapt_betas(self, epoch):
        pass

    def run_training(self):
        min_loss = self.get_start_min_loss()
        epoch = self.get_start_epoch()
        for epoch in range(self.get_start_epoch(), self._n_epochs):
            self.adapt_lr(epoch)
            self.adapt_betas(epoch)

            # (1) training
            self._model.train()
            epoch_metrics = MetricMeasuresDtoInit.init_dto()
            for batch_metrics in self.train_batches(self._dataloader_training, epoch):
                epoch_metrics.add(batch_metrics)
            epoch_metrics.div(len(self._dataloader_training))
            self.print_epoch(epoch, 'training', epoch_metrics)
            self._metric_dtos['training'].append(epoch_metrics)

            # (2) validation
            self._model.eval()
            if self._dataloader_validation is None:
                epoch_metrics = MetricMeasuresDtoInit.init_dto(*([0.0] * 13))
            else:
                epoch_metrics = MetricMeasuresDtoInit.init_dto()
                for batch in self._dataloader_validation:
                    epoch_metrics.add(self.validate_batch(batch, epoch))
                epoch_metrics.div(len(self._dataloader_validation))
            self.print_epoch(epoch, 'validate', epoch_metrics)
            self._metric_dtos['validate'].append(epoch_metrics)

            # (3) checkpoint on a new validation optimum
            if self._metric_dtos['validate'] and self._metric_dtos['validate'][-1].loss < min_loss:
                min_loss = self._metric_dtos['validate'][-1].loss
                self.save_model()
                self.save_training()
                print('(New optimum: Training saved)', end=' ')
                self.visualize_epoch(epoch)
            if epoch % 50 == 0:
                self.visualize_epoch(epoch)

            # (4) curves
            if epoch > 0:
                try:
                    import matplotlib.pyplot as plt
                except ImportError:
                    plt = None
                if plt is not None:
                    fig, plot = plt.subplots()
                    self.plot_epoch(plot, range(1, epoch + 2))
                    fig.savefig(self._path_outputs_base + self.FN_VIS_BASE + 'plots.png', bbox_inches='tight', dpi=300)
                    plt.close(fig)

        # (5) final model
        self.save_model('_final')
        self.visualize_epoch(epoch)
