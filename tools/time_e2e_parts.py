"""Where the end-to-end step (Learner.train_batch on a pinned host batch) spends its time beyond the resident step: variants with
the per-batch metrics switched off / serial / without the surface distances, and the bare H2D of the batch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import bench  # noqa: E402
from stroke_prediction_b200.common import metrics  # noqa: E402
from stroke_prediction_b200.learner.Learner import Learner  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = bench.Workload("cae200", 8, dev, 0, 1)
    print("resident step            %.2f ms" % timed(w.step_resident))
    print("train_batch (default)    %.2f ms" % timed(w.step_e2e))
    Learner.OVERLAP_METRICS = False
    print("  metrics serial         %.2f ms" % timed(w.step_e2e))
    Learner.OVERLAP_METRICS = True
    metrics.SURFACE_DISTANCES = False
    print("  no surface distances   %.2f ms" % timed(w.step_e2e))
    orig = w.learner.batch_metrics_step
    w.learner.batch_metrics_step = lambda dto, epoch: Learner.batch_metrics_step(w.learner, dto, epoch)
    print("  no metrics at all      %.2f ms" % timed(w.step_e2e))
    w.learner.batch_metrics_step = orig
    metrics.SURFACE_DISTANCES = True
    host = w.host
    def h2d():
        return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
    print("bare H2D of the batch    %.2f ms (%.1f MB)" % (timed(h2d), sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v)) / 1e6))


if __name__ == "__main__":
    main()
