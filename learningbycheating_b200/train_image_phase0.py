"""Drop-in for training/train_image_phase0.py -- phase-0 warm-up of the camera student.

Mirrors the reference's entry points: ``CoordConverter`` (:36-79), ``LocationLoss`` (:81-89),
``train_or_eval`` (:152-209) with the same signature and per-iteration order of operations (one_hot ->
teacher under no_grad -> student -> target transform -> per-sample L1 -> mean -> zero_grad / backward / step).
Dataloading stays whatever iterable the caller passes; the visual logging (:91-147, cv2) is
outside the hot path.
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
SAVE_EPOCHS = [1, 2, 4, 8, 16, 32, 64, 128, 256, 384, 512, 768, 1000]


class CoordConverter():
    """train_image_phase0.py:36-79: teacher map coords [-1,1] -> image pixel targets [B,5,2]."""

    def __init__(self, w=384, h=160, fov=90, world_y=1.4, fixed_offset=4.0, device='cuda'):
        self._w, self._h, self._fov = w, h, fov
        self._world_y, self._fixed_offset = world_y, fixed_offset
        self._img_size = torch.FloatTensor([w, h]).to(device)

    def __call__(self, map_locations):
        return losses.phase0_target(map_locations, self._w, self._h, self._fov, self._world_y, self._fixed_offset)


class LocationLoss(torch.nn.Module):
    """train_image_phase0.py:81-89: mean_{step,xy} |pred - (loc/(0.5*[w,h]) - 1)| -> [B]."""

    def __init__(self, w=384, h=160, device='cuda', **kwargs):
        super().__init__()
        self._w, self._h = float(w), float(h)
        self._img_size = torch.FloatTensor([w, h]).to(device)

    def forward(self, pred_locations, locations):
        locations = locations.to(pred_locations.device)
        return losses.l1_location_loss(pred_locations, locations, 1.0, 0.0, 2.0 / self._w, 2.0 / self._h, -1.0)


def train_or_eval(coord_converter, criterion, net, teacher_net, data, optim, is_train, config, is_first_epoch):
    if is_train:
        net.train()
    else:
        net.eval()
    tick = time.time()
    losses_seen = []
    for i, (rgb_image, birdview, location, command, speed) in enumerate(data):
        dev = config['device']
        rgb_image = rgb_image.to(dev, non_blocking=True)
        birdview = birdview.to(dev, non_blocking=True)
        command = one_hot(command).to(dev, non_blocking=True)
        speed = speed.to(dev, non_blocking=True)

        with torch.no_grad():
            _teac_location = teacher_net(birdview, speed, command)
            if isinstance(_teac_location, tuple):
                _teac_location = _teac_location[0]

        _pred_location = net(rgb_image, speed, command)
        if isinstance(_pred_location, tuple):
            _pred_location = _pred_location[0]
        teac_location = coord_converter(_teac_location)

        loss = criterion(_pred_location, teac_location)
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
    """train_image_phase0.py:214-242.  ``train(config)`` alone follows the reference: models from ``config['model_args']`` /
    ``config['teacher_args']['model_path']``, data through the dataset hook (_train_common.load_data), checkpoints on
    SAVE_EPOCHS into ``config['log_dir']``.  Anything passed explicitly (iterables, models, optimiser) is used as is."""
    from . import _train_common as tc
    from .birdview import BirdViewPolicyModelSS
    from .image import ImagePolicyModelSS
    from .optim import Adam
    if data_train is None or data_val is None:
        data_train, data_val = tc.load_data(config)
    criterion = LocationLoss(device=config['device'], **config['camera_args'])
    if net is None:
        net = ImagePolicyModelSS(config['model_args']['backbone'],
                                 pretrained=config['model_args'].get('imagenet_pretrained', False)).to(config['device'])
    if teacher_net is None:
        teacher_net = BirdViewPolicyModelSS(tc.teacher_backbone(config)).to(config['device'])
        teacher_net.load_state_dict(torch.load(config['teacher_args']['model_path'], map_location=config['device']))
    teacher_net.eval()
    coord_converter = CoordConverter(device=config['device'], **config['camera_args'])
    optim = optim or Adam(net.parameters(), lr=config['optimizer_args']['lr'])
    for epoch in range(int(config['max_epoch']) + 1):
        train_or_eval(coord_converter, criterion, net, teacher_net, data_train, optim, True, config, epoch == 0)
        train_or_eval(coord_converter, criterion, net, teacher_net, data_val, None, False, config, epoch == 0)
        tc.save_checkpoint(net, config, epoch, SAVE_EPOCHS)
        _log.end_epoch()
    return net
