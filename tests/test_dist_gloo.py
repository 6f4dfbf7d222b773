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
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n)
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()

    def attach_grad_sink(self):
        class S:
            pass
        s = S()
        s.flat = self.flat
        return s


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
    with torch.no_grad():
        for n, p in zip(names, params):
            p.grad.copy_(g[n])
    sync()
    averaged = {n: (p.grad * opt.grad_scale).clone() for n, p in zip(names, params)}
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
