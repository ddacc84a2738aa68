"""TEST INFRASTRUCTURE ONLY -- import shims for the *unmodified* reference tree.

Loads dotchen/LearningByCheating from /root/reference (read-only mount, present
only in the build container, never on the GPU box) so that
  * oracle/make_golden.py can dump golden vectors from the reference itself, and
  * tests can cross-check oracle/lbc_oracle.py (the CPU restatement) against it.

Nothing in learningbycheating_b200/ may import this module.

Shims; every one lives outside /root/reference:
  1. MagicMock stubs for carla, pygame, lmdb, imgaug, imageio, tensorboardX, removed from sys.modules again
     once the reference is imported (import chain: bird_view/models/agent.py:5, utils/carla_utils.py:12-15, ...)
  2. sys.modules['train_util'] = utils.train_utils   (training/train_image_phase0.py:25)
  3. sys.path += bird_view/, training/, PythonAPI/
  4. torch.Tensor.cuda = identity on a GPU-less host (bird_view/models/common.py:105-106)
  5. loguru.logger._handlers = {}  (bird_view/utils/bz_utils/saver.py:63)
  6. torch.distributions validate_args off (training/train_image_phase1.py:172)
"""
import os
import sys
import types
from unittest import mock

REF_ROOT = os.environ.get("LBC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "bird_view", "models"))


_loaded = {}


def load():
    """Returns a namespace with the reference's model classes and training modules."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch

    stubbed = []
    for name in ["carla", "pygame", "pygame.locals", "lmdb", "imgaug",
                 "imgaug.augmenters", "imageio", "tensorboardX"]:
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            m.__name__ = name
            m.__spec__ = None
            sys.modules[name] = m
            stubbed.append(name)
    for sub in ["bird_view", "training", "PythonAPI", ""]:
        p = os.path.join(REF_ROOT, sub) if sub else REF_ROOT
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        import loguru
        if not hasattr(loguru.logger, "_handlers"):
            loguru.logger._handlers = {}
    except ImportError:
        pass
    import torch.distributions
    torch.distributions.Distribution.set_default_validate_args(False)

    import utils.train_utils as train_utils
    sys.modules["train_util"] = train_utils
    from models.image import ImagePolicyModelSS
    from models.birdview import BirdViewPolicyModelSS
    from models import common
    import train_image_phase0
    import train_image_phase1
    import train_birdview

    ns = types.SimpleNamespace(
        ImagePolicyModelSS=ImagePolicyModelSS,
        BirdViewPolicyModelSS=BirdViewPolicyModelSS,
        common=common,
        one_hot=train_utils.one_hot,
        phase0=train_image_phase0,
        phase1=train_image_phase1,
        birdview=train_birdview,
    )
    # the reference modules keep their own references to the stubs; anything imported later (the package's optional
    # `import lmdb`, tests of its "module missing" paths) must not find a MagicMock posing as the real module
    for name in stubbed:
        sys.modules.pop(name, None)
    _loaded["ns"] = ns
    return ns
