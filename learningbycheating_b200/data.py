"""Host->device staging for the training loops: copy batch i+1 on a side stream while batch i computes.

The reference's loop does `.to(device)` on the compute stream for every tensor (training/train_image_phase0.py:167-170),
so a 189 MB fp32 RGB batch (B=256) costs a PCIe transfer in the critical path of every step.  `CudaPrefetcher` keeps the
reference's batch contract (tuples of CPU tensors in, the same tuples of CUDA tensors out) and only moves the copy
off the critical path; tensors that are not pinned are pinned once per batch.
"""
import torch


class CudaPrefetcher:
    def __init__(self, iterable, device, transform=None):
        self.iterable = iterable
        self.device = torch.device(device)
        self.transform = transform          # optional host-side hook applied to each CPU batch (e.g. one_hot)
        self.stream = torch.cuda.Stream(self.device)

    def __len__(self):
        return len(self.iterable)

    def _stage(self, batch):
        if self.transform is not None:
            batch = self.transform(batch)
        out = []
        with torch.cuda.stream(self.stream):
            for t in batch:
                if torch.is_tensor(t):
                    if not t.is_cuda and not t.is_pinned():
                        t = t.pin_memory()
                    t = t.to(self.device, non_blocking=True)
                out.append(t)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return tuple(out), ev

    def __iter__(self):
        it = iter(self.iterable)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in cur:
                if torch.is_tensor(t):
                    t.record_stream(torch.cuda.current_stream(self.device))
            yield cur
