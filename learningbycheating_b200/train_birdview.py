"""Drop-in for training/train_birdview.py (config 5): trains the privileged BirdViewPolicyModelSS itself.
``LocationLoss`` (:33-54, 'l1' as selected at :158), ``train_or_eval`` (:102-153)."""
import time

import torch

from . import losses
from .log import log as _log
from .train_utils import one_hot

BACKBONE = 'resnet18'
N_STEP = 5
SAVE_EPOCHS = [1, 2, 4, 8, 16, 32, 64, 128, 256, 384, 512, 768, 1000]


class LocationLoss(torch.nn.Module):
    def __init__(self, w=192, h=192, choice='l2'):
        super().__init__()
        if choice != 'l1':
            raise NotImplementedError("train_birdview.py:158 trains with choice='l1'; 'l2' is unused")
        self._w, self._h = float(w), float(h)

    def forward(self, pred_location, gt_location):
        return losses.l1_location_loss(pred_location, gt_location.float(), 1.0, 0.0, 2.0 / self._w, 2.0 / self._h, -1.0)


def train_or_eval(criterion, net, data, optim, is_train, config, is_first_epoch):
    if is_train:
        net.train()
    else:
        net.eval()
    tick = time.time()
    losses_seen = []
    for i, (birdview, location, command, speed) in enumerate(data):
        dev = config['device']
        birdview = birdview.to(dev, non_blocking=True)
        command = one_hot(command).to(dev, non_blocking=True)
        speed = speed.to(dev, non_blocking=True)
        location = location.float().to(dev, non_blocking=True)

        pred_location = net(birdview, speed, command)
        if isinstance(pred_location, tuple):
            pred_location = pred_location[0]
        loss = criterion(pred_location, location)
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


def train(config, data_train=None, data_val=None, net=None, optim=None):
    """train_birdview.py:155-181.  ``train(config)`` alone follows the reference (model from ``config['model_args']``, resume
    from the newest ``model-*.th`` in ``config['log_dir']`` when ``config['resume']``, data through the dataset hook,
    checkpoints on SAVE_EPOCHS); anything passed explicitly is used as is."""
    import glob
    import os

    from . import _train_common as tc
    from .birdview import BirdViewPolicyModelSS
    from .optim import Adam
    if data_train is None or data_val is None:
        data_train, data_val = tc.load_data(config)
    criterion = LocationLoss(w=192, h=192, choice='l1')
    if net is None:
        net = BirdViewPolicyModelSS(config['model_args']['backbone']).to(config['device'])
        if config.get('resume'):
            ckpts = sorted(glob.glob(os.path.join(str(config['log_dir']), 'model-*.th')),
                           key=lambda p_: int(os.path.basename(p_)[6:-3]))
            net.load_state_dict(torch.load(ckpts[-1], map_location=config['device']))
    optim = optim or Adam(net.parameters(), lr=config['optimizer_args']['lr'])
    for epoch in range(int(config['max_epoch']) + 1):
        train_or_eval(criterion, net, data_train, optim, True, config, epoch == 0)
        train_or_eval(criterion, net, data_val, None, False, config, epoch == 0)
        tc.save_checkpoint(net, config, epoch, SAVE_EPOCHS)
        _log.end_epoch()
    return net
