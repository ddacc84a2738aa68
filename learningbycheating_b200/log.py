"""Minimal stand-in for ``bz_utils.log`` (bird_view/utils/bz_utils/saver.py:51-136): the training loops only
need ``scalar`` / ``image`` / ``end_epoch``; tensorboard / loguru output is host-side tooling outside the hot path
(SURVEY.md 5.5).  Scalars are accumulated per epoch and summarised by ``end_epoch``."""
import collections


class Experiment:
    def __init__(self):
        self.scalars = {True: collections.defaultdict(list), False: collections.defaultdict(list)}
        self.epoch = 0
        self.history = []

    def init(self, log_dir=None):
        self.log_dir = log_dir

    def scalar(self, is_train=True, **kwargs):
        for k, v in kwargs.items():
            self.scalars[bool(is_train)][k].append(float(v))

    def image(self, is_train=True, **kwargs):
        pass

    def end_epoch(self):
        summary = {}
        for mode in (True, False):
            for k, v in self.scalars[mode].items():
                if v:
                    summary[("train/" if mode else "val/") + k] = sum(v) / len(v)
            self.scalars[mode].clear()
        self.history.append(summary)
        self.epoch += 1
        return summary


log = Experiment()
