"""stroke-prediction_b200 — B200-native (sm_100a) hot path of multimodallearning/stroke-prediction.

Import name: ``stroke_prediction_b200`` (the directory carries the reference's hyphenated name; the sibling
``stroke_prediction_b200/`` package re-exports it).  The sub-packages ``common``, ``learner`` and ``tester`` mirror the
reference's module tree; :func:`install_reference_aliases` registers them under the reference's top-level names so
whole-module pickles written by the reference (``Learner.py:113``) resolve to these classes.
"""
import importlib
import sys

__version__ = "0.1.0"


def install_reference_aliases():
    """Make ``import common.model.Cae3D`` / ``learner...`` / ``tester...`` resolve to this package's drop-ins."""
    pkg = __name__
    for top in ("common", "learner", "tester"):
        mod = importlib.import_module(pkg + "." + top)
        sys.modules.setdefault(top, mod)
    for sub in ("common.dto", "common.dto.Dto", "common.dto.CaeDto", "common.dto.UnetDto",
                "common.dto.MetricMeasuresDto", "common.model", "common.model.Cae3D", "common.model.Unet3D",
                "common.metrics", "common.data", "common.inference", "common.inference.Inference",
                "common.inference.CaeInference", "common.inference.UnetInference",
                "common.inference.CaeEncInference", "learner.Learner", "learner.CaeReconstructionLearner",
                "learner.UnetSegmentationLearner", "learner.CaeStepLearner", "learner.CaePredictionLearner",
                "tester.Tester"):
        sys.modules.setdefault(sub, importlib.import_module(pkg + "." + sub))
