"""Data-parallel plumbing for the training step (new functionality: the reference is single-GPU).

One process per GPU (torchrun), parameters / Adam moments / BN buffers replicated, the batch sharded.  The only
exchange of the path is the all-reduce (sum) of the flat gradient array over NCCL/NVLink; the 1/world_size average is
folded into the native Adam kernel (``Adam.grad_scale``).  BatchNorm statistics stay per replica (reference
semantics = single-replica BN); rank 0's running buffers are the ones checkpointed.

``DataParallel.backward(loss)`` + ``step_after_backward()``: with ``overlap=True`` the gradient is reduced in five
buckets in the order backward finishes them (heads + decoder, layer4, layer3, layer2, layer1 + stem): the native
backward records a CUDA event per bucket, a side stream waits for each event and launches that bucket's ncclAllReduce
while the remaining layers are still being differentiated; Adam waits for the side stream.  ``conv.fc.*`` (never
trained) is in no bucket.  ``overlap=False`` (default): one all-reduce of the whole array after backward.

Measured on 2 x B200 (round 2, config 2, B = 256 per GPU): 14.45 ms on one GPU, 14.66 ms with the single all-reduce,
14.72 ms with the bucketed overlapped one -- NCCL's CTAs take SMs (and shared memory) away from the persistent
one-CTA-per-SM GEMM grids that run next to them, which costs as much as the hidden transfer saves; hence the default.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def broadcast_parameters(net, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and BN buffers (flat arrays: two broadcasts)."""
    st = net.lbc_flat_state()
    dist.broadcast(st.flat_params, src, group=group)
    dist.broadcast(st.flat_bufs, src, group=group)


def allreduce_gradients(net, group=None, async_op=False):
    """Sum the flat gradient array over ranks (call between backward() and optimizer.step())."""
    st = net._lbc
    return dist.all_reduce(st.flat_grads, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def grad_buckets(net):
    """[(offset, numel)] of the flat gradient array, in the order backward completes them."""
    st = net.lbc_flat_state()
    L = _lib.lib()
    out = []
    for k in range(L.lbc_net_num_grad_buckets(st.handle)):
        off, n = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.lbc_net_grad_bucket(st.handle, k, ctypes.byref(off), ctypes.byref(n)))
        out.append((off.value, n.value))
    return out


class DataParallel:
    """Wraps (net, optimizer): ``backward(loss)`` then ``step_after_backward()`` = all-reduce + Adam (grad_scale = 1/world)."""

    def __init__(self, net, optimizer, group=None, overlap=False):
        self.net, self.optimizer, self.group = net, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        optimizer.grad_scale = 1.0 / self.world
        self.overlap = bool(overlap) and self.world > 1
        self._side = None
        self._buckets = None
        self._handle = None
        self._reduced = False

    def sync_initial_state(self):
        if self.world > 1:
            broadcast_parameters(self.net, 0, self.group)

    def _prepare(self):
        st = self.net.lbc_flat_state()
        if self._handle is not st.handle:       # (re)created engine: events live in the native object
            _lib.check(_lib.lib().lbc_net_enable_grad_events(st.handle, 1))
            self._buckets = grad_buckets(self.net)
            self._handle = st.handle
        if self._side is None:
            self._side = torch.cuda.Stream(st.device)
        return st

    def backward(self, loss):
        """loss.backward() with the bucketed all-reduce launched underneath it."""
        if not self.overlap:
            loss.backward()
            self._reduced = False
            return
        if _lib.is_host_emulation():            # CPU unit tests (gloo): same buckets, no streams / events
            loss.backward()
            st = self.net._lbc
            for off, n in grad_buckets(self.net):
                dist.all_reduce(st.flat_grads[off:off + n], op=dist.ReduceOp.SUM, group=self.group)
            self._reduced = True
            return
        st = self._prepare()
        loss.backward()                         # enqueues every backward kernel and records the bucket events
        L = _lib.lib()
        main = torch.cuda.current_stream(st.device)
        side_ptr = ctypes.c_void_p(self._side.cuda_stream)
        with torch.cuda.stream(self._side):
            for k, (off, n) in enumerate(self._buckets):
                _lib.check(L.lbc_net_stream_wait_grads(st.handle, k, side_ptr))
                dist.all_reduce(st.flat_grads[off:off + n], op=dist.ReduceOp.SUM, group=self.group)
        main.wait_stream(self._side)            # Adam (next on the main stream) sees the reduced gradient
        self._reduced = True

    def step_after_backward(self):
        if self.world > 1 and not self._reduced:
            allreduce_gradients(self.net, self.group)
        self._reduced = False
        self.optimizer.step()
