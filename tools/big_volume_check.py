"""Scaled-volume sanity run (BASELINE configs[4]): one CAE train step at 1x60x256x256, batch 2, through the same Learner API;
checks that every tier's index arithmetic holds at 8.6x the voxels of the benchmark config and prints the step time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stroke_prediction_b200.common import data                                   # noqa: E402
from stroke_prediction_b200.common.metrics import BatchDiceLoss                  # noqa: E402
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D        # noqa: E402
from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner   # noqa: E402
from stroke_prediction_b200.optim import FusedAdam                               # noqa: E402


def main():
    D, HW, B = (int(a) for a in (sys.argv[1:4] + ["60", "256", "2"][len(sys.argv) - 1:]))
    ch = [1, 16, 24, 32, 100, 200, 1]
    torch.manual_seed(4)
    cae = Cae3D(Enc3D(HW, D, ch, 5, 1.0), Dec3D(HW, D, ch, 5, 1.0)).cuda().train()
    opt = FusedAdam(list(cae.parameters()), lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
    learner = CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/big", BatchDiceLoss([1.0]))
    batch = data.synthetic_cae_batch(B, size=(D, HW, HW), seed=4)
    batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        m = learner.train_batch(batch, 30)
        torch.cuda.synchronize()
        print("step %d: loss %.6f  lesion dc %.4f  %.1f ms  (%.2f volumes/s)" % (i, m.loss, m.lesion.dc, (time.time() - t0) * 1e3,
                                                                                 B / (time.time() - t0)))
        assert m.loss == m.loss and abs(m.loss) < 10.0
    print("peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 2**30))


if __name__ == "__main__":
    main()
