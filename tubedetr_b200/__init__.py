"""tubedetr_b200: B200-native (sm_100a) implementation of TubeDETR's forward/backward hot path.

Drop-in for the reference's `models` package:  `from tubedetr_b200 import build_model`
(reference models/__init__.py:3-4).  The CUDA library csrc/libtdb.so is required; there is no fallback path.
"""
__all__ = ["build_model", "TubeDETR", "SetCriterion", "NestedTensor"]


def __getattr__(name):
    if name in ("build_model", "TubeDETR", "SetCriterion", "NestedTensor"):
        from . import model
        return {"build_model": model.build, "TubeDETR": model.TubeDETR, "SetCriterion": model.SetCriterion,
                "NestedTensor": model.NestedTensor}[name]
    raise AttributeError(name)
