"""CPU, world_size 2 over gloo: the data-parallel host logic (batch sharding, flat-gradient all-reduce, 1/world
averaging folded into the update, parameter broadcast).  The arithmetic kernels are not involved (no GPU here): the
"local gradient" of each rank is produced by the CPU oracle on that rank's shard, which is exactly the DDP semantics
the design claims (SURVEY §8e mode 1): reference run on each shard, gradients averaged."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import stroke_oracle as O
from stroke_prediction_b200 import parallel
from stroke_prediction_b200.common import data
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D

CH = [1, 4, 6, 8, 10, 12, 1]
SIZE = (28, 56, 56)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _local_grads(sd, batch):
    labels = batch[data.KEY_LABELS]
    step = O.time_to_treatment(batch[data.KEY_GLOBAL])
    lat, rec = O.cae_forward(sd, CH, 1.0, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    loss = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], 60)
    return O.grads_of(loss, sd)


class _FakeOpt:
    """Stands in for FusedAdam on CPU: parameters whose .grad are views of one flat buffer (the GradSink contract)."""

    def __init__(self, params):
        self.param_groups = [{"params": params}]
        self.grad_scale = 1.0
        pad = lambda n: (n + 3) // 4 * 4                 # engine.GradSink keeps every view 16-byte aligned
        self.flat = torch.zeros(sum(pad(p.numel()) for p in params))
        self.offsets = {}
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view(p.shape)
            self.offsets[id(p)] = off
            off += pad(p.numel())
        self.params = params

    def attach_grad_sink(self):
        class S:
            pass
        s = S()
        s.flat, s._offsets, s.params = self.flat, self.offsets, self.params
        return s


class _FakePlan:
    """What engine.plan_backward_hooks hands over: something with .params() (engine.SeqPlan)."""

    def __init__(self, module):
        self._params = list(module.parameters())

    def params(self):
        return self._params


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    torch.manual_seed(100 + rank)        # different init per rank on purpose: broadcast must fix it
    cae = Cae3D(Enc3D(SIZE[1], SIZE[0], CH, 5, 1.0), Dec3D(SIZE[1], SIZE[0], CH, 5, 1.0))
    parallel.broadcast_parameters(cae, src=0)
    full = data.synthetic_cae_batch(4, size=SIZE, seed=9)
    shard = parallel.shard_batch(full, rank, world)
    assert shard[data.KEY_LABELS].shape[0] == 2 and shard[data.KEY_GLOBAL].shape[0] == 2
    params = [p for p in cae.parameters()]
    opt = _FakeOpt(params)
    sync = parallel.GradientAllReduce(opt)
    assert opt.grad_scale == 0.5 and sync.world == 2
    sd = O.clone_state(cae.state_dict(), requires_grad=True)
    g = _local_grads(sd, shard)
    names = [n for n, _ in cae.named_parameters()]
    # backward order of the CAE: the decoder's gradients are final first -> its slice is reduced while the "encoder backward"
    # (here: the copy of the encoder gradients) still runs; the call between backward and step covers the rest
    dec_plan, enc_plan = _FakePlan(cae.dec.decoder), _FakePlan(cae.enc.encoder)
    dec_ids = {id(p) for p in dec_plan.params()}
    with torch.no_grad():
        for n, p in zip(names, params):
            if id(p) in dec_ids:
                p.grad.copy_(g[n])
    lo, hi = sync.plan_range(dec_plan)
    assert hi - lo >= sum(p.numel() for p in dec_plan.params()) and sync.plan_range(enc_plan)[1] == lo
    sync.plan_ready(dec_plan)
    sync.plan_ready(dec_plan)                # a second notification must not reduce the slice twice
    assert sync.early_elements == hi - lo
    with torch.no_grad():
        for n, p in zip(names, params):
            if id(p) not in dec_ids:
                p.grad.copy_(g[n])
    sync()
    assert sync.last_early_elements == hi - lo and sync.early_elements == 0 and not sync._done and not sync._work
    averaged = {n: (p.grad * opt.grad_scale).clone() for n, p in zip(names, params)}
    # a second step without any early notification: one collective over the whole buffer, same arithmetic
    with torch.no_grad():
        for n, p in zip(names, params):
            p.grad.copy_(g[n])
    sync()
    for n, p in zip(names, params):
        assert torch.equal(p.grad * opt.grad_scale, averaged[n]), n
    assert sync._uncovered(0, 10) == [(0, 10)]
    sync._done = [(2, 4), (6, 8)]
    assert sync._uncovered(0, 10) == [(0, 2), (4, 6), (8, 10)] and sync._uncovered(3, 7) == [(4, 6)]
    sync._done = []
    if rank == 0:
        torch.save({"avg": averaged, "sd": {k: v.detach() for k, v in cae.state_dict().items()}}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_mean_of_shard_gradients(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    full = data.synthetic_cae_batch(4, size=SIZE, seed=9)
    ref = None
    for r in range(2):
        sd = O.clone_state(got["sd"], requires_grad=True)
        g = _local_grads(sd, parallel.shard_batch(full, r, 2))
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    for k, v in got["avg"].items():
        want = ref[k] / 2
        # the workers ran the oracle with 2 threads, this process with all of them: fp32 reduction order differs
        assert (v - want).abs().max().item() <= 1e-3 * max(1e-6, want.abs().max().item()) + 1e-9, k


def test_shard_range_covers_everything_once():
    for n in (1, 7, 8, 33):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
