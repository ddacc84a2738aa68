"""Data-parallel plumbing for the training step (new functionality: the reference is single-GPU).

One process per GPU (torchrun), parameters / Adam moments / BN buffers replicated, the batch sharded.  The only
exchange of the path is ONE all-reduce (sum) of the flat 23.1 M-element gradient array per step over NCCL/NVLink;
the 1/world_size average is folded into the native Adam kernel (``Adam.grad_scale``).  BatchNorm statistics stay
per replica (reference semantics = single-replica BN); rank 0's running buffers are the ones checkpointed.
"""
import torch
import torch.distributed as dist


def broadcast_parameters(net, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and BN buffers (flat arrays: two broadcasts)."""
    st = net.lbc_flat_state()
    dist.broadcast(st.flat_params, src, group=group)
    dist.broadcast(st.flat_bufs, src, group=group)


def allreduce_gradients(net, group=None, async_op=False):
    """Sum the flat gradient array over ranks (call between backward() and optimizer.step())."""
    st = net._lbc
    return dist.all_reduce(st.flat_grads, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class DataParallel:
    """Wraps (net, optimizer): ``step_after_backward()`` = all-reduce + Adam with grad_scale = 1/world."""

    def __init__(self, net, optimizer, group=None):
        self.net, self.optimizer, self.group = net, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        optimizer.grad_scale = 1.0 / self.world

    def sync_initial_state(self):
        if self.world > 1:
            broadcast_parameters(self.net, 0, self.group)

    def step_after_backward(self):
        if self.world > 1:
            allreduce_gradients(self.net, self.group)
        self.optimizer.step()
