"""Host->device staging for the training loops: copy batch i+1 on a side stream while batch i computes.

The reference's loop does `.to(device)` on the compute stream for every tensor (training/train_image_phase0.py:167-170),
so a 189 MB fp32 RGB batch (B=256) costs a PCIe transfer in the critical path of every step.  `CudaPrefetcher` keeps the
reference's batch contract (tuples of CPU tensors in, the same tuples of CUDA tensors out) and only moves the copy
off the critical path; tensors that are not pinned are pinned once per batch.
"""
import torch


class CudaPrefetcher:
    """Device staging is two fixed slots of flat byte buffers (grown on demand, never per batch): the steady state makes no
    allocator call, so no cudaMalloc / cross-stream block recycling can stall the loop (per-batch `.to(device)` on a side
    stream did exactly that: 20-40 % swings of the end-to-end rate on some boxes).  A yielded batch lives in a slot that is
    overwritten when the batch after the next one is staged: consume it within the iteration (clone what must outlive it)."""

    SLOTS = 2

    def __init__(self, iterable, device, transform=None):
        self.iterable = iterable
        self.device = torch.device(device)
        self.transform = transform          # optional host-side hook applied to each CPU batch (e.g. one_hot)
        self.stream = torch.cuda.Stream(self.device)
        self._bufs = [dict() for _ in range(self.SLOTS)]      # slot -> {position in the batch tuple: flat uint8 buffer}
        self._done = [None] * self.SLOTS                      # compute-stream event: the slot's last consumer was enqueued
        self._n = 0

    def __len__(self):
        return len(self.iterable)

    def _buffer(self, slot, pos, nbytes):
        buf = self._bufs[slot].get(pos)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                buf.record_stream(self.stream)
            buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
            self._bufs[slot][pos] = buf
        return buf

    def _stage(self, batch):
        if self.transform is not None:
            batch = self.transform(batch)
        slot = self._n % self.SLOTS
        self._n += 1
        if self._done[slot] is not None:
            self.stream.wait_event(self._done[slot])          # the step that read this slot has finished before it is overwritten
        out = []
        for pos, t in enumerate(batch):
            if torch.is_tensor(t) and not t.is_cuda:
                if not t.is_pinned():
                    t = t.pin_memory()
                t = t.contiguous()
                nbytes = t.numel() * t.element_size()
                dst = self._buffer(slot, pos, nbytes)[:nbytes].view(t.dtype).view(t.shape)
                with torch.cuda.stream(self.stream):
                    dst.copy_(t, non_blocking=True)
                t = dst
            out.append(t)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return tuple(out), ev, slot

    def __iter__(self):
        it = iter(self.iterable)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev, slot = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            try:
                yield cur
            finally:                          # also when the consumer leaves the loop early
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))   # the consumer has enqueued its work on `cur`
                self._done[slot] = done
