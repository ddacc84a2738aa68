"""learningbycheating_b200 -- B200-native (sm_100a) training hot path of LearningByCheating's image agent.

Public surface mirrors the reference: ``ImagePolicyModelSS`` (bird_view/models/image.py),
``BirdViewPolicyModelSS`` (bird_view/models/birdview.py), ``one_hot`` (bird_view/utils/train_utils.py),
and the ``train_image_phase0`` / ``train_image_phase1`` / ``train_birdview`` modules with the reference's
``train`` / ``train_or_eval`` / ``CoordConverter`` / ``LocationLoss`` entry points.
"""
from ._lib import LbcError  # noqa: F401
from .image import ImagePolicyModelSS  # noqa: F401
from .birdview import BirdViewPolicyModelSS  # noqa: F401
from .train_utils import one_hot  # noqa: F401
from .optim import Adam  # noqa: F401
