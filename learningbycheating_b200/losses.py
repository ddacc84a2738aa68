"""Native loss kernels behind the reference's CoordConverter / LocationLoss call signatures.

All tensors stay on the device: the reference's phase-0 target transform leaves the graph through numpy and
cv2.projectPoints on the host every step (training/train_image_phase0.py:67-79); here it is one small kernel
(lbc_phase0_target) computing the same pinhole projection in float64.
"""
import torch

from . import _lib


class _L1LossFn(torch.autograd.Function):
    """loss_b[n] = mean_d |a*sa+ta - (b*sb+tb)|; gradient only wrt ``a`` (targets are constants in all three
    training scripts: teacher outputs are produced under no_grad)."""

    @staticmethod
    def forward(ctx, a, b, sa, ta, sbx, sby, tb):
        n = a.shape[0]
        d = a[0].numel()
        a_c, b_c = a.contiguous(), b.contiguous()
        loss_b = torch.empty(n, dtype=torch.float32, device=a.device)
        _lib.check(_lib.lib().lbc_l1_loss(_lib.ptr(a_c), _lib.ptr(b_c), n, d, sa, ta, sbx, sby, tb, None,
                                          _lib.ptr(loss_b), None, _lib.stream_ptr(a.device)))
        ctx.save_for_backward(a_c, b_c)
        ctx.consts = (sa, ta, sbx, sby, tb)
        return loss_b

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        sa, ta, sbx, sby, tb = ctx.consts
        n = a.shape[0]
        d = a[0].numel()
        da = torch.empty_like(a)
        g = gout.contiguous().float()
        _lib.check(_lib.lib().lbc_l1_loss(_lib.ptr(a), _lib.ptr(b), n, d, sa, ta, sbx, sby, tb, _lib.ptr(g),
                                          None, _lib.ptr(da), _lib.stream_ptr(a.device)))
        return da, None, None, None, None, None, None


def l1_location_loss(a, b, sa=1.0, ta=0.0, sbx=1.0, sby=1.0, tb=0.0):
    if a.shape != b.shape or a.dtype != torch.float32 or b.dtype != torch.float32:
        raise _lib.LbcError("l1_location_loss: shape/dtype mismatch %s %s vs %s %s" % (a.shape, a.dtype, b.shape, b.dtype))
    if a[0].numel() % 2 != 0:
        raise _lib.LbcError("l1_location_loss: trailing dimension must hold (x, y) pairs")
    return _L1LossFn.apply(a, b, float(sa), float(ta), float(sbx), float(sby), float(tb))


def phase0_target(teacher_pred, w, h, fov, world_y, fixed_offset):
    t = teacher_pred.detach().contiguous().float()
    out = torch.empty_like(t)
    _lib.check(_lib.lib().lbc_phase0_target(_lib.ptr(t), _lib.ptr(out), t.numel() // 2, float(w), float(h), float(fov),
                                            float(world_y), float(fixed_offset), _lib.stream_ptr(t.device)))
    return out


class _Phase1ConvertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, w, h, fov, world_y, fixed_offset):
        pc = p.contiguous()
        out = torch.empty_like(pc)
        _lib.check(_lib.lib().lbc_phase1_convert_fwd(_lib.ptr(pc), _lib.ptr(out), pc.numel() // 2, w, h, fov, world_y,
                                                     fixed_offset, _lib.stream_ptr(pc.device)))
        ctx.save_for_backward(pc)
        ctx.consts = (w, h, fov, world_y, fixed_offset)
        return out

    @staticmethod
    def backward(ctx, dout):
        (pc,) = ctx.saved_tensors
        w, h, fov, world_y, fixed_offset = ctx.consts
        dp = torch.empty_like(pc)
        d = dout.contiguous().float()
        _lib.check(_lib.lib().lbc_phase1_convert_bwd(_lib.ptr(pc), _lib.ptr(d), _lib.ptr(dp), pc.numel() // 2, w, h, fov,
                                                     world_y, fixed_offset, _lib.stream_ptr(pc.device)))
        return dp, None, None, None, None, None


def phase1_convert(p, w, h, fov, world_y, fixed_offset):
    if p.dtype != torch.float32 or p.shape[-1] != 2:
        raise _lib.LbcError("phase1_convert expects float32 [...,2]")
    return _Phase1ConvertFn.apply(p, float(w), float(h), float(fov), float(world_y), float(fixed_offset))


def phase2_weight(learner_map, teacher_map):
    """training/phase2_utils.py:50-59 ``get_weight`` (replay-buffer resampling weight, no gradient)."""
    a = learner_map.detach().contiguous().float()
    b = teacher_map.detach().contiguous().float()
    if a.shape != b.shape or a.dim() != 3 or tuple(a.shape[1:]) != (5, 2):
        raise _lib.LbcError("phase2_weight expects two [N,5,2] tensors")
    out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().lbc_phase2_weight(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.shape[0], _lib.stream_ptr(a.device)))
    return out
