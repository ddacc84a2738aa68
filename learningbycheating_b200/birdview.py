"""Drop-in for bird_view/models/birdview.py:47-79 -- ``BirdViewPolicyModelSS`` (privileged teacher).

Used frozen (eval + no_grad) as the target generator of train_image_phase0/1 and trained itself by
train_birdview.py (config 5).  ``BirdViewAgent`` (birdview.py:82-174) is out of scope.
"""
from . import common

STEPS = 5
COMMANDS = 4


class BirdViewPolicyModelSS(common.PolicyNetBase):
    _lbc_kind = common.KIND_BIRDVIEW_RESNET18
    _lbc_input_shape = (7, 192, 192)

    def __init__(self, backbone="resnet18", input_channel=7, n_step=5, all_branch=False, **kwargs):
        if backbone != "resnet18" or input_channel != 7 or n_step != 5:
            raise ValueError("the B200 hot path implements the ResNet-18 / 7-channel / 5-step teacher only "
                             "(training/train_birdview.py:28)")
        super().__init__(backbone, input_channel=input_channel, bias_first=False,
                         precision=kwargs.pop("lbc_precision", None))
        self.deconv = common._decoder_params()
        self.location_pred = common._head_params(48, 48, STEPS, COMMANDS)
        self.all_branch = all_branch
