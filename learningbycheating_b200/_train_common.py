"""Shared pieces of the three ``train(config)`` entry points (training/train_image_phase0.py:214-242,
train_image_phase1.py:232-263, train_birdview.py:155-181): dataset hook, teacher loading, checkpoint cadence."""
import json
import os

import torch

from . import _lib


def load_data(config):
    """The reference builds its loaders with ``load_data(**config['data_args'])`` over LMDB shards (image_lmdb.py:222-281,
    birdview_lmdb.py:169-230) -- outside this hot path.  The host application passes the iterables to ``train`` directly or
    provides ``config['data_loader']``, a callable taking ``**config['data_args']`` and returning (data_train, data_val)."""
    loader = config.get('data_loader')
    if loader is None:
        # the package's own reader of the collector's LMDB episodes (records.py; no imgaug augmentation)
        args = config.get('data_args', {})
        try:
            import lmdb  # noqa: F401
            have_lmdb = True
        except ImportError:
            have_lmdb = False
        if have_lmdb and args.get('dataset_dir'):
            from . import records
            if config.get('model_args', {}).get('model', 'image_ss') == 'image_ss':
                return records.get_image(**args)
            return records.get_birdview(**args)      # the privileged agent's dataset, with its rotation / shift jitter
        raise _lib.LbcError("train(config): no dataset -- pass data_train/data_val, or set config['data_loader'] to a callable "
                            "(**data_args) -> (data_train, data_val); records.get_image reads the collector's LMDB episodes "
                            "when the `lmdb` module is installed")
    return loader(**config.get('data_args', {}))


def teacher_backbone(config):
    """bzu.log.load_config(model_path): the teacher run's config.json next to its checkpoint (saver.py); default resnet18."""
    path = config['teacher_args']['model_path']
    cfg = os.path.join(os.path.dirname(str(path)), 'config.json')
    if os.path.exists(cfg):
        try:
            return json.load(open(cfg))['model_args']['backbone']
        except (KeyError, ValueError):
            pass
    return config['teacher_args'].get('backbone', 'resnet18')


def save_checkpoint(net, config, epoch, save_epochs):
    """torch.save(net.state_dict(), log_dir/model-<epoch>.th) on the reference's epochs."""
    if epoch in save_epochs and config.get('log_dir'):
        os.makedirs(str(config['log_dir']), exist_ok=True)
        torch.save(net.state_dict(), os.path.join(str(config['log_dir']), 'model-%d.th' % epoch))
