"""Drop-in for training/train_image_phase1.py -- phase-1 distillation on all four command branches.

``CoordConverter`` (:35-64, differentiable image->map transform), ``LocationLoss`` (:66-70),
``train_or_eval`` (:157-229).  Speed noise / batch_aug reshuffling (:172,180-189) are kept as host logic.
"""
import time

import torch

from . import losses
from .log import log as _log
from .train_utils import one_hot

BACKBONE = 'resnet34'
N_STEP = 5
PIXELS_PER_METER = 5
CROP_SIZE = 192
SAVE_EPOCHS = [1, 2, 4, 8, 16, 32, 64, 128, 192, 256]


class CoordConverter():
    """train_image_phase1.py:35-64.  Singular at the image horizon (yt -> 0) exactly like the reference."""

    def __init__(self, w=384, h=160, fov=90, world_y=1.4, fixed_offset=2.0, device='cuda'):
        self._w, self._h, self._fov = w, h, fov
        self._world_y, self._fixed_offset = world_y, fixed_offset
        self._img_size = torch.FloatTensor([w, h]).to(device)

    def __call__(self, camera_locations):
        return losses.phase1_convert(camera_locations, self._w, self._h, self._fov, self._world_y, self._fixed_offset)


class LocationLoss(torch.nn.Module):
    """train_image_phase1.py:66-70: mean_{branch,step,xy} |pred/96 - 1 - teac| -> [B]."""

    def forward(self, pred_locations, teac_locations):
        return losses.l1_location_loss(pred_locations, teac_locations, 1.0 / (0.5 * CROP_SIZE), -1.0, 1.0, 1.0, 0.0)


def repeat(a, repeats, dim=0):
    """train_image_phase1.py:131-154 == numpy.repeat along ``dim``."""
    return torch.repeat_interleave(a, repeats, dim=dim)


def train_or_eval(coord_converter, criterion, net, teacher_net, data, optim, is_train, config, is_first_epoch):
    if is_train:
        net.train()
    else:
        net.eval()
    tick = time.time()
    losses_seen = []
    speed_noise = float(config.get('speed_noise', 0.0))
    for i, (rgb_image, birdview, location, command, speed) in enumerate(data):
        dev = config['device']
        rgb_image = rgb_image.to(dev, non_blocking=True)
        birdview = birdview.to(dev, non_blocking=True)
        command = one_hot(command).to(dev, non_blocking=True)
        speed = speed.to(dev, non_blocking=True)

        if is_train and speed_noise > 0:
            speed = torch.clamp(speed + torch.randn(speed.size(), device=speed.device) * speed_noise, 0, 10)

        if len(rgb_image.size()) > 4:
            B, batch_aug, c, h, w = rgb_image.size()
            rgb_image = rgb_image.view(B * batch_aug, c, h, w)
            birdview = repeat(birdview, batch_aug)
            command = repeat(command, batch_aug)
            speed = repeat(speed, batch_aug)

        with torch.no_grad():
            _teac_location, _teac_locations = teacher_net(birdview, speed, command)

        _pred_location, _pred_locations = net(rgb_image, speed, command)
        pred_locations = coord_converter(_pred_locations)

        loss = criterion(pred_locations, _teac_locations)
        loss_mean = loss.mean()

        if is_train and not is_first_epoch:
            optim.zero_grad()
            loss_mean.backward()
            optim.step()

        should_log = (i % int(config['log_iterations']) == 0) or (not is_train) or is_first_epoch
        if should_log:
            _log.scalar(is_train=is_train, loss_mean=loss_mean.item())
        losses_seen.append(loss_mean.detach())
        _log.scalar(is_train=is_train, fps=1.0 / max(time.time() - tick, 1e-9))
        tick = time.time()
        if is_first_epoch and i == 10:
            break
    return losses_seen


def train(config, data_train=None, data_val=None, net=None, teacher_net=None, optim=None):
    """train_image_phase1.py:232-263.  ``train(config)`` alone follows the reference (student from ``config['phase0_ckpt']``,
    teacher from ``config['teacher_args']['model_path']``, data through the dataset hook, checkpoints on SAVE_EPOCHS);
    anything passed explicitly is used as is."""
    from . import _train_common as tc
    from .birdview import BirdViewPolicyModelSS
    from .image import ImagePolicyModelSS
    from .optim import Adam
    if data_train is None or data_val is None:
        data_train, data_val = tc.load_data(config)
    criterion = LocationLoss()
    if net is None:
        net = ImagePolicyModelSS(config['model_args']['backbone'],
                                 pretrained=config['model_args'].get('imagenet_pretrained', False),
                                 all_branch=True).to(config['device'])
        net.load_state_dict(torch.load(config['phase0_ckpt'], map_location=config['device']))
    if teacher_net is None:
        teacher_net = BirdViewPolicyModelSS(tc.teacher_backbone(config), all_branch=True).to(config['device'])
        teacher_net.load_state_dict(torch.load(config['teacher_args']['model_path'], map_location=config['device']))
    teacher_net.eval()
    coord_converter = CoordConverter(device=config['device'], **config['agent_args']['camera_args'])
    optim = optim or Adam(net.parameters(), lr=config['optimizer_args']['lr'])
    for epoch in range(int(config['max_epoch']) + 1):
        train_or_eval(coord_converter, criterion, net, teacher_net, data_train, optim, True, config, epoch == 0)
        train_or_eval(coord_converter, criterion, net, teacher_net, data_val, None, False, config, epoch == 0)
        tc.save_checkpoint(net, config, epoch, SAVE_EPOCHS)
        _log.end_epoch()
    return net
