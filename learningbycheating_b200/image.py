"""Drop-in for bird_view/models/image.py:22-89 -- ``ImagePolicyModelSS`` (the camera student).

Same constructor signature (extra kwargs swallowed: benchmark_agent.py:36 passes the whole model_args
dict), same ``state_dict`` keys, same ``forward(image, velocity, command)`` contract; the compute is the
native sm_100a engine (include/lbc_b200.h), never torch ops.  ``ImageAgent`` (CARLA-side PID control,
image.py:93-219) is out of scope.
"""
from . import common

STEPS = 5
COMMANDS = 4
CROP_SIZE = 192
PIXELS_PER_METER = 5


class ImagePolicyModelSS(common.PolicyNetBase):
    _lbc_kind = common.KIND_IMAGE_RESNET34
    _lbc_input_shape = (3, 160, 384)

    def __init__(self, backbone, warp=False, pretrained=False, all_branch=False, **kwargs):
        if backbone != "resnet34":
            raise ValueError("the B200 hot path implements the ResNet-34 student only (BACKBONE='resnet34', "
                             "training/train_image_phase0.py:28); got %r" % (backbone,))
        if warp:
            raise NotImplementedError("warp=True is dead code in the reference (image.py:65-68 uses undefined names)")
        if pretrained:
            raise RuntimeError("ImageNet weights need a URL download (resnet.py:175-178); load a state_dict instead")
        super().__init__(backbone, input_channel=3, bias_first=False, precision=kwargs.pop("lbc_precision", None))
        self.c = 512
        self.warp = warp
        # rgb_transform (common.NormalizeV2, image.py:32-35) holds no parameters/buffers; the engine applies it
        self.deconv = common._decoder_params()
        ow, oh = 96, 40
        self.location_pred = common._head_params(ow, oh, STEPS, COMMANDS)
        self.all_branch = all_branch
