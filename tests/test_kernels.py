"""Every kernel family the bf16 training step launches, alone, against torch on the same bf16-rounded operands, at the
shapes of the path -- and each test reads the library's launch trace, so it FAILS when the correctness-first kernel ran
instead of the kernel under test (lbc_trace_enable / lbc_trace_dump).  The cpu variants run the same entry points on the
host-emulation build (fp32, and bf16 storage through the correctness-first bodies): they pin the torch references of
this file and the wrapper logic, not the fast kernels.

Reference ops: nn.BatchNorm2d train mode (+residual, +ReLU) and its backward (resnet.py:41-53), nn.MaxPool2d after
BN+ReLU (resnet.py:150-152), nn.ConvTranspose2d+bias+ReLU (image.py:39-46), the downsample branch's input gradient
(resnet.py:48-52), the four BN->1x1->SpatialSoftmax heads (image.py:54-60, common.py:136-152), NormalizeV2 + the 7x7/s2
stem (common.py:101-109, resnet.py:102,148)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _L():
    from learningbycheating_b200 import _lib
    return _lib


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _r(t, precision):
    """round to the storage dtype of the path under test"""
    return t.bfloat16().float() if precision == 1 else t


def _err(got, ref):
    return (got.double() - ref.double()).abs().max().item() / max(ref.double().abs().max().item(), 1e-30)


REF_TAGS = ("k_head_dlogits", "k_conv_fwd", "k_conv_dgrad", "k_conv_wgrad_part", "k_bn_sum_part", "k_bn_var_part", "k_bn_apply",
            "k_bn_bwd_part", "k_bn_bwd_apply", "k_maxpool_fwd", "k_maxpool_bwd", "k_relu_mask", "k_add]", "k_head_logits",
            "k_head_softmax", "k_head_s_part", "k_head_dh", "k_colsum_part", "k_input_nhwc")


class Traced:
    """with Traced(dev, must=[kernel names], never=REF_TAGS): ... -- on the GPU asserts the named fast kernels ran and none
    of the correctness-first compute kernels did; on the host-emulation build it only records."""

    def __init__(self, dev, must=(), never=REF_TAGS):
        self.dev, self.must, self.never = dev, must, never

    def __enter__(self):
        _L().trace(True)
        return self

    def __exit__(self, et, ev, tb):
        self.counts = _L().trace_counts()
        _L().trace(False)
        if et is None and self.dev == "cuda":
            for m in self.must:
                assert any(k == m or k.startswith(m) for k in self.counts), (m, sorted(self.counts))
            for tag in self.never:
                bad = [k for k in self.counts if tag in k]
                assert not bad, ("correctness-first kernel ran instead of the fast kernel", bad, sorted(self.counts))
        return False


# ------------------------------------------------------------------ BatchNorm forward
def _bn_fwd_case(dev, M, C, relu, with_res, precision, shift):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    x = _r(torch.randn(M, C, generator=g) * 2 + 0.5, precision)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = _r(torch.randn(M, C, generator=g), precision) if with_res else None
    y_ref = F.batch_norm(x.t().reshape(1, C, M, 1), None, None, gamma, beta, True, 0.1, 1e-5).reshape(C, M).t()
    if with_res:
        y_ref = y_ref + res
    if relu:
        y_ref = y_ref.clamp_min(0)
    d = lambda t: None if t is None else t.contiguous().to(dev)
    rm0, rv0 = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    ns = (torch.randn(C, generator=g) * 0.3) if shift else None       # the stored tensor is (x + negshift)
    xs = _r(x + ns, precision) if shift else x
    xd, gd, bd, rd = d(xs), d(gamma), d(beta), d(res)
    rm, rv, nsd = d(rm0.clone()), d(rv0.clone()), d(ns)
    y, mean, var = torch.empty(M, C, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev)
    bits = torch.zeros(M, C // 8, dtype=torch.uint8, device=dev) if (precision == 1 and dev == "cuda") else None
    must = ["bn_stats_kernel", "bn_finalize_kernel", "bn_apply_kernel"] if precision == 1 else []
    with Traced(dev, must, REF_TAGS if precision == 1 else ()):
        _lib.check(L.lbc_op_bn_train(_lib.ptr(xd), _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(rd), int(relu), _lib.ptr(y),
                                     _lib.ptr(mean), _lib.ptr(var), M, C, precision, _lib.ptr(rm), _lib.ptr(rv),
                                     _lib.ptr(nsd), _lib.ptr(bits), None))
    if bits is not None:     # the (y > 0) bits the backward kernels use instead of re-reading the activation
        yb = (y.cpu() > 0).reshape(M, C // 8, 8).to(torch.int32)
        ref_bits = sum(yb[:, :, j] << j for j in range(8)).to(torch.uint8)
        assert torch.equal(bits.cpu(), ref_bits)
    tol = 1.2e-2 if precision == 1 else 2e-5
    if shift:       # statistics of the ROUNDED shifted tensor; the apply un-shifts implicitly
        xs64 = xs.double()
        y_ref = F.batch_norm(xs.t().reshape(1, C, M, 1), None, None, gamma, beta, True, 0.1, 1e-5).reshape(C, M).t()
        if with_res:
            y_ref = y_ref + res
        if relu:
            y_ref = y_ref.clamp_min(0)
        true_mean = xs64.mean(0) - ns.double()
        assert (nsd.cpu().double() + true_mean).abs().max() < 1e-4          # negshift <- -(batch mean of the true x)
        assert (mean.cpu().double() - xs64.mean(0)).abs().max() < 1e-4      # saved mean is that of the stored tensor
    else:
        true_mean = x.double().mean(0)
        assert (mean.cpu().double() - true_mean).abs().max() < 1e-4
    assert _err(y.cpu(), y_ref) < tol
    xv = (xs if shift else x).double()
    assert (var.cpu().double() - xv.var(0, unbiased=False)).abs().max() < 1e-3
    # running buffers: momentum 0.1, unbiased variance (nn.BatchNorm2d)
    assert (rm.cpu().double() - (0.9 * rm0.double() + 0.1 * true_mean)).abs().max() < 1e-4
    assert (rv.cpu().double() - (0.9 * rv0.double() + 0.1 * xv.var(0, unbiased=True))).abs().max() < 1e-3


BN_SHAPES = [(3 * 40 * 96, 64, True, False), (2 * 20 * 48, 128, True, True), (3 * 5 * 12, 640, False, False),
             (777, 256, False, True), (5 * 5 * 12, 512, True, True)]


@pytest.mark.parametrize("M,C,relu,res", BN_SHAPES[:3])
@pytest.mark.parametrize("precision", [0, 1])
def test_bn_forward_cpu(backend, M, C, relu, res, precision):
    _bn_fwd_case(backend, min(M, 1500), C, relu, res, precision, False)


@pytest.mark.gpu
@pytest.mark.parametrize("M,C,relu,res", BN_SHAPES)
@pytest.mark.parametrize("shift", [False, True])
def test_bn_forward_kernels_gpu(backend, M, C, relu, res, shift):
    _bn_fwd_case("cuda", M, C, relu, res, 1, shift)


# ------------------------------------------------------------------ BatchNorm backward
def _bn_bwd_case(dev, M, C, mode, precision, use_bits=False):
    """mode: 'plain' | 'mask' (a separate activation's ReLU mask on dy) | 'own' (mask = relu of this BN's own output)"""
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(7)
    x = _r(torch.randn(M, C, generator=g) * 1.5 + 0.3, precision)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5
    dy = _r(torch.randn(M, C, generator=g), precision)
    if mode == "own":
        # the fast kernels recompute the mask (bn(x) > 0) from x: move the few x whose BN output sits within rounding
        # distance of zero (bf16-exact steps), so that the reference and the kernel cannot disagree on any mask bit
        for _ in range(8):
            y0 = F.batch_norm(x.t().reshape(1, C, M, 1), None, None, gamma, beta, True, 0.1, 1e-5).reshape(C, M).t()
            near = y0.abs() < 4e-3
            if not bool(near.any()):
                break
            x = _r(torch.where(near, x + 0.0625, x), 1)
        assert not bool(near.any())
    xt = x.t().reshape(1, C, M, 1).clone().requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.batch_norm(xt, None, None, gp, bp, True, 0.1, 1e-5)
    act = None
    if mode == "mask":
        act = _r(torch.randn(M, C, generator=g), precision)
    elif mode == "own":
        act = y.detach().reshape(C, M).t().clamp_min(0)
    dym = dy if act is None else dy * (act > 0).float()
    y.backward(dym.t().reshape(1, C, M, 1))
    d = lambda t: None if t is None else t.contiguous().to(dev)
    xd, dyd, gd, bd, ad = d(x), d(dy), d(gamma), d(beta), d(act)
    dg, db, dx = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty(M, C, device=dev)
    own = mode == "own"
    if precision == 1:
        must = ["bn_bwd_reduce_kernel", "bn_bwd_apply_kernel"]
        if own:
            must = [m + "<own>" for m in must]
    else:
        must = []
    never = tuple(t for t in REF_TAGS if t not in ("k_bn_sum_part", "k_bn_var_part")) if precision == 1 else ()
    with Traced(dev, must, never):     # (the op wrapper recomputes the batch statistics with the correctness-first pass)
        _lib.check(L.lbc_op_bn_bwd(_lib.ptr(dyd), _lib.ptr(xd), _lib.ptr(gd), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(dx), M, C,
                                   precision, _lib.ptr(ad), _lib.ptr(bd), 1 if own else 0, 1 if use_bits else 0, None))
    # dgamma / dbeta are fp32 sums of exact (bf16 x bf16) products on both paths
    scale = max(1.0, gp.grad.abs().max().item(), bp.grad.abs().max().item())
    assert (dg.cpu() - gp.grad).abs().max() <= 2e-4 * scale, (dg.cpu() - gp.grad).abs().max()
    assert (db.cpu() - bp.grad).abs().max() <= 2e-4 * scale
    assert _err(dx.cpu(), xt.grad.reshape(C, M).t()) < (1.2e-2 if precision == 1 else 1e-4)     # dx is stored in bf16


@pytest.mark.parametrize("mode", ["plain", "mask", "own"])
@pytest.mark.parametrize("precision", [0, 1])
def test_bn_backward_cpu(backend, mode, precision):
    _bn_bwd_case(backend, 900, 64, mode, precision)


@pytest.mark.gpu
@pytest.mark.parametrize("M,C", [(3 * 40 * 96, 64), (2 * 20 * 48, 128), (9 * 10 * 24, 256), (33 * 5 * 12, 512), (60, 640)])
@pytest.mark.parametrize("mode", ["plain", "mask", "own"])
def test_bn_backward_kernels_gpu(backend, M, C, mode):
    _bn_bwd_case("cuda", M, C, mode, 1)
    if mode == "mask":       # the same mask handed over as bits (the block-final ReLU of the training step)
        _bn_bwd_case("cuda", M, C, mode, 1, use_bits=True)


# ------------------------------------------------------------------ masked adds
def _ew_case(dev, precision, use_bits=False):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(11)
    n = 3 * 20 * 48 * 128
    for mode in (0, 1, 2):
        a, b, m = (_r(torch.randn(n, generator=g), precision) for _ in range(3))
        ref = a + b if mode == 0 else (a + b * (m > 0).float() if mode == 1 else a * (m > 0).float())
        ad, bd, md = a.clone().to(dev), b.to(dev), m.to(dev)
        with Traced(dev, ["ew_kernel"] if precision == 1 else [], REF_TAGS if precision == 1 else ()):
            _lib.check(L.lbc_op_ew(_lib.ptr(ad), _lib.ptr(bd) if mode != 2 else None, _lib.ptr(md) if mode != 0 else None, n, mode,
                                   precision, 1 if use_bits else 0, None))
        assert _err(ad.cpu(), ref) < (8e-3 if precision == 1 else 1e-6), mode


@pytest.mark.parametrize("precision", [0, 1])
def test_masked_adds_cpu(backend, precision):
    _ew_case(backend, precision)


@pytest.mark.gpu
def test_masked_add_kernel_gpu(backend):
    _ew_case("cuda", 1)
    _ew_case("cuda", 1, use_bits=True)


# ------------------------------------------------------------------ speed fusion (image.py:77-79) and its slice in backward
def _copy_channels_case(dev, precision):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(13)
    N, HW = 5, 60
    trunk = _r(torch.randn(N * HW, 512, generator=g), precision)
    speed = torch.rand(N, generator=g) * 10
    wd, td, sd = torch.empty(N * HW, 640).to(dev), trunk.to(dev), speed.to(dev)   # (named: the pointers must stay alive)
    with Traced(dev, ["copy_channels_kernel"] if precision == 1 else [], REF_TAGS if precision == 1 else ()):
        _lib.check(L.lbc_op_copy_channels(_lib.ptr(td), _lib.ptr(wd), N * HW, 640, 512, _lib.ptr(sd), HW, precision, None))
    ref = torch.cat([trunk, _r(speed, precision).repeat_interleave(HW)[:, None].expand(-1, 128)], 1)
    assert torch.equal(wd.cpu(), ref)
    gwide = _r(torch.randn(N * HW, 640, generator=g), precision)
    nd, gd = torch.empty(N * HW, 512).to(dev), gwide.to(dev)
    with Traced(dev, ["copy_channels_kernel"] if precision == 1 else [], REF_TAGS if precision == 1 else ()):
        _lib.check(L.lbc_op_copy_channels(_lib.ptr(gd), _lib.ptr(nd), N * HW, 512, 640, None, 1, precision, None))
    assert torch.equal(nd.cpu(), gwide[:, :512])


@pytest.mark.parametrize("precision", [0, 1])
def test_copy_channels_cpu(backend, precision):
    _copy_channels_case(backend, precision)


@pytest.mark.gpu
def test_copy_channels_kernel_gpu(backend):
    _copy_channels_case("cuda", 1)


# ------------------------------------------------------------------ residual add fused with the next BatchNorm's reduce pass
def _resid_bn_case(dev, M, C, precision):
    """dst += src * (act > 0); BatchNorm backward of dst * (act_prev > 0) wrt x  (the d(out) chain between two BasicBlocks)"""
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(31)
    dst, src, act, act_prev = (_r(torch.randn(M, C, generator=g), precision) for _ in range(4))
    x = _r(torch.randn(M, C, generator=g) * 1.5 + 0.3, precision)
    gamma = torch.rand(C, generator=g) + 0.5
    summed = _r(dst + src * (act > 0).float(), precision)             # the stored (rounded) sum is what the BN consumes
    xt = x.t().reshape(1, C, M, 1).clone().requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), torch.zeros(C, requires_grad=True)
    y = F.batch_norm(xt, None, None, gp, bp, True, 0.1, 1e-5)
    y.backward((summed * (act_prev > 0).float()).t().reshape(1, C, M, 1))
    d = lambda t: t.contiguous().to(dev)
    dd, sd, ad, apd, xd, gd = d(dst.clone()), d(src), d(act), d(act_prev), d(x), d(gamma)
    dg, db, dx = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty(M, C, device=dev)
    must = ["bn_bwd_reduce_kernel<resid>", "col_finalize_kernel", "bn_bwd_apply_kernel"] if precision == 1 else []
    never = tuple(t for t in REF_TAGS if t not in ("k_bn_sum_part", "k_bn_var_part")) if precision == 1 else ()
    with Traced(dev, must, never):
        _lib.check(L.lbc_op_resid_bn_bwd(_lib.ptr(dd), _lib.ptr(sd), _lib.ptr(ad), _lib.ptr(xd), _lib.ptr(apd), _lib.ptr(gd),
                                         _lib.ptr(dg), _lib.ptr(db), _lib.ptr(dx), M, C, precision, None))
    assert _err(dd.cpu(), summed) < (1e-6 if precision == 0 else 1e-6)      # same rounding on both sides
    scale = max(1.0, gp.grad.abs().max().item(), bp.grad.abs().max().item())
    assert (dg.cpu() - gp.grad).abs().max() <= 2e-4 * scale
    assert (db.cpu() - bp.grad).abs().max() <= 2e-4 * scale
    assert _err(dx.cpu(), xt.grad.reshape(C, M).t()) < (1.2e-2 if precision == 1 else 1e-4)


@pytest.mark.parametrize("precision", [0, 1])
def test_resid_bn_backward_cpu(backend, precision):
    _resid_bn_case(backend, 900, 64, precision)


@pytest.mark.gpu
@pytest.mark.parametrize("M,C", [(3 * 40 * 96, 64), (2 * 20 * 48, 128), (9 * 10 * 24, 256), (33 * 5 * 12, 512)])
def test_resid_bn_backward_kernel_gpu(backend, M, C):
    _resid_bn_case("cuda", M, C, 1)


# ------------------------------------------------------------------ stem tail: BN + ReLU + MaxPool and its backward
def _pool_case(dev, N, H, W, C, precision):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(13)
    x = _r(torch.randn(N, C, H, W, generator=g) * 1.3, precision)
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.7
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    sc = gamma * rstd
    sh = beta - mean * sc
    z = (x * sc.view(1, C, 1, 1) + sh.view(1, C, 1, 1)).requires_grad_(True)     # BN output (given statistics)
    a = _r(z.clamp_min(0), precision) if precision == 1 else z.clamp_min(0)       # the bf16 path pools the ROUNDED activation
    a = z.clamp_min(0) + (a - z.clamp_min(0)).detach()
    y_ref = F.max_pool2d(a, 3, 2, 1)
    dy = _r(torch.randn(y_ref.shape, generator=g), precision)
    (y_ref * dy).sum().backward()
    d = lambda t: t.contiguous().to(dev)
    xd, dyd = d(_nhwc(x)), d(_nhwc(dy))
    md, rd, gd, bd = d(mean), d(rstd), d(gamma), d(beta)
    OH, OW = y_ref.shape[2:]
    y, dx = torch.empty(N, OH, OW, C, device=dev), torch.empty(N, H, W, C, device=dev)
    must = ["bn_relu_maxpool_kernel", "maxpool_relu_bwd_kernel"] if precision == 1 else []
    with Traced(dev, must, REF_TAGS if precision == 1 else ()):
        _lib.check(L.lbc_op_bn_relu_maxpool(_lib.ptr(xd), _lib.ptr(md), _lib.ptr(rd), _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(y),
                                            _lib.ptr(dyd), _lib.ptr(dx), N, H, W, C, precision, None))
    assert _err(_nchw(y.cpu()), y_ref.detach()) < (8e-3 if precision == 1 else 1e-5)
    # ties are only possible at exact zeros, where the ReLU mask kills the gradient; bf16 rounding of the
    # activation can also create ties between positive neighbours: compare where the reference window max is unique
    ref = z.grad
    got = _nchw(dx.cpu())
    if precision == 1:
        bad = ((got - ref).abs() > 1e-2 * ref.abs().max()).float().mean().item()
        assert bad < 2e-3, bad                       # isolated bf16 ties only
        assert abs(float(got.double().sum()) - float(ref.double().sum())) < 2e-2 * float(ref.double().abs().sum()) ** 0.5 + 1e-3 * float(ref.double().abs().sum())
    else:
        assert torch.allclose(got, ref, atol=1e-5)


@pytest.mark.parametrize("precision", [0, 1])
def test_bn_relu_maxpool_cpu(backend, precision):
    _pool_case(backend, 2, 10, 12, 8, precision)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W", [(2, 80, 192), (3, 96, 96), (1, 12, 16)])
def test_bn_relu_maxpool_kernels_gpu(backend, N, H, W):
    _pool_case("cuda", N, H, W, 64, 1)


def _stem_tail_case(dev, N, H, W, C, precision):
    """train-mode BN -> ReLU -> MaxPool and its whole backward (dgamma, dbeta, d raw) vs torch autograd"""
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(41)
    x = _r(torch.randn(N, C, H, W, generator=g) * 1.4 + 0.2, precision).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.3).requires_grad_(True)
    z = F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5)
    a = z.clamp_min(0)
    if precision == 1:       # the bf16 path pools the ROUNDED activation
        a = a + (_r(a.detach(), 1) - a.detach())
    y_ref = F.max_pool2d(a, 3, 2, 1)
    dy = _r(torch.randn(y_ref.shape, generator=g), precision)
    (y_ref * dy).sum().backward()
    d = lambda t: t.detach().contiguous().to(dev)
    xd, dyd, gd, bd = d(_nhwc(x)), d(_nhwc(dy)), d(gamma), d(beta)
    OH, OW = y_ref.shape[2:]
    y, dx = torch.empty(N, OH, OW, C, device=dev), torch.empty(N, H, W, C, device=dev)
    dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
    must = ["bn_relu_maxpool_kernel", "maxpool_relu_bwd_kernel", "bn_bwd_reduce_kernel", "bn_bwd_apply_kernel"] if precision == 1 else []
    never = tuple(t for t in REF_TAGS if t not in ("k_bn_sum_part", "k_bn_var_part")) if precision == 1 else ()
    with Traced(dev, must, never):
        _lib.check(L.lbc_op_stem_tail(_lib.ptr(xd), _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(y), _lib.ptr(dyd), _lib.ptr(dg), _lib.ptr(db),
                                      _lib.ptr(dx), N, H, W, C, precision, None))
    assert _err(_nchw(y.cpu()), y_ref.detach()) < (8e-3 if precision == 1 else 1e-5)
    scale = max(1.0, gamma.grad.abs().max().item(), beta.grad.abs().max().item())
    tol = 5e-3 if precision == 1 else 1e-4       # bf16: a tie between equal rounded activations may move a gradient to a neighbour
    assert (dg.cpu() - gamma.grad).abs().max() <= tol * scale
    assert (db.cpu() - beta.grad).abs().max() <= tol * scale
    ref, got = x.grad, _nchw(dx.cpu())
    if precision == 1:
        bad = ((got - ref).abs() > 2e-2 * ref.abs().max()).float().mean().item()
        assert bad < 2e-3, bad
    else:
        assert _err(got, ref) < 1e-4


@pytest.mark.parametrize("precision", [0, 1])
def test_stem_tail_cpu(backend, precision):
    _stem_tail_case(backend, 2, 10, 12, 8, precision)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W", [(2, 80, 192), (3, 96, 96)])
def test_stem_tail_kernels_gpu(backend, N, H, W):
    _stem_tail_case("cuda", N, H, W, 64, 1)


# ------------------------------------------------------------------ the four heads
def _head_ref(hn, gamma, beta, w, bias, H, W):
    """image.py:54-60 + common.py:136-152 restated with torch ops; hn [N,64,H,W] -> logits [N,20,HW], preds [N,4,5,2]"""
    N = hn.shape[0]
    idx = torch.arange(H * W)
    px = (-1 + 2 * (idx % W).double() / (W - 1)).float()
    py = (-1 + 2 * (idx // W).double() / (H - 1)).float()
    logits, preds = [], []
    for k in range(4):
        z = F.batch_norm(hn, None, None, gamma[k], beta[k], True, 0.1, 1e-5)
        lg = F.conv2d(z, w[k].reshape(5, 64, 1, 1), bias[k]).reshape(N, 5, H * W)
        sm = torch.softmax(lg, -1)
        logits.append(lg)
        preds.append(torch.stack([(sm * px).sum(-1), (sm * py).sum(-1)], -1))
    return torch.cat(logits, 1), torch.stack(preds, 1)


def _head_case(dev, N, H, W, precision, use_pred):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(17)
    h = _r(torch.randn(N, 64, H, W, generator=g).clamp_min(0) * 1.7, precision)        # post-ReLU decoder output
    gamma, beta = torch.rand(4, 64, generator=g) + 0.5, torch.randn(4, 64, generator=g) * 0.2
    w, bias = torch.randn(4, 5, 64, generator=g) * 0.3, torch.randn(4, 5, generator=g) * 0.1
    cmd = torch.randint(0, 4, (N,), generator=g)
    onehot = F.one_hot(cmd, 4).float()
    d_preds = torch.randn(N, 4, 5, 2, generator=g)
    d_pred = torch.randn(N, 5, 2, generator=g) if use_pred else None
    hn = h.clone().requires_grad_(True)
    P = [t.clone().requires_grad_(True) for t in (gamma, beta, w, bias)]
    logits_ref, preds_ref = _head_ref(hn, *P, H, W)
    loss = (preds_ref * d_preds).sum()
    if use_pred:
        loss = loss + ((onehot.view(N, 4, 1, 1) * preds_ref).sum(1) * d_pred).sum()
    loss.backward()
    d = lambda t: None if t is None else t.contiguous().to(dev)
    hd = d(_nhwc(h).reshape(N, H * W, 64))
    gd, bd, wd, bid, ohd, dpsd, dpd = d(gamma), d(beta), d(w), d(bias), d(onehot), d(d_preds), d(d_pred)
    rm, rv = torch.zeros(4, 64, device=dev), torch.ones(4, 64, device=dev)
    logits, preds = torch.empty(N, 20, H * W, device=dev), torch.empty(N, 4, 5, 2, device=dev)
    dg, db, dw, dbi = (torch.empty_like(t, device=dev) for t in (gamma, beta, w, bias))
    dh = torch.empty(N, H * W, 64, device=dev)
    masked = ctypes.c_int(0)
    must = ["bn_stats_kernel", "head_logits", "head_softmax_kernel", "head_dlogits_kernel", "head_s", "head_dh"] if precision == 1 else []
    with Traced(dev, must, REF_TAGS if precision == 1 else ()):
        _lib.check(L.lbc_op_head(_lib.ptr(hd), _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(wd), _lib.ptr(bid), N, H, W, _lib.ptr(rm),
                                 _lib.ptr(rv), _lib.ptr(logits), _lib.ptr(preds), _lib.ptr(ohd), _lib.ptr(dpd), _lib.ptr(dpsd),
                                 _lib.ptr(dg), _lib.ptr(db), _lib.ptr(dw), _lib.ptr(dbi), _lib.ptr(dh), ctypes.byref(masked),
                                 precision, None))
    # forward: h is given exactly (bf16-representable), everything after it is fp32 in both paths
    assert _err(logits.cpu(), logits_ref.detach()) < 2e-5
    assert (preds.cpu() - preds_ref.detach()).abs().max() < 2e-5
    # running buffers of all four heads see the same batch statistics (image.py:56)
    hm = h.permute(1, 0, 2, 3).reshape(64, -1).double()
    for k in range(4):
        assert (rm[k].cpu().double() - 0.1 * hm.mean(1)).abs().max() < 1e-5
        assert (rv[k].cpu().double() - (0.9 + 0.1 * hm.var(1, unbiased=True))).abs().max() < 1e-4
    tol = 2e-4
    for got, ref, name in ((dg, P[0].grad, "dgamma"), (dw, P[2].grad, "dw")):
        assert _err(got.cpu(), ref) < tol, name
    # softmax shift invariance: beta / bias receive exactly zero -- rounding noise only
    scale = max(P[0].grad.abs().max().item(), 1.0)
    assert db.cpu().abs().max() < 1e-4 * scale and dbi.cpu().abs().max() < 1e-4 * scale
    dh_ref = _nhwc(hn.grad).reshape(N, H * W, 64)
    if masked.value:
        dh_ref = dh_ref * (hd.cpu() > 0).float()
    assert _err(dh.cpu(), dh_ref) < (1e-2 if precision == 1 else 2e-4)     # dh is stored in bf16 on that path


@pytest.mark.parametrize("precision", [0, 1])
def test_heads_cpu(backend, precision):
    _head_case(backend, 2, 6, 8, precision, True)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W", [(3, 40, 96), (2, 48, 48)])
@pytest.mark.parametrize("use_pred", [False, True])
def test_head_kernels_gpu(backend, N, H, W, use_pred):
    _head_case("cuda", N, H, W, 1, use_pred)


@pytest.mark.gpu
def test_spatial_softmax_kernel_gpu(backend):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(19)
    for H, W in ((40, 96), (48, 48)):
        logits = torch.randn(40, H * W, generator=g) * 4
        out = torch.empty(40, 2, device="cuda")
        ld = logits.cuda()
        with Traced("cuda", ["head_softmax_kernel"]):
            _lib.check(L.lbc_op_spatial_softmax(_lib.ptr(ld), _lib.ptr(out), 40, H, W, 1, None))
        idx = torch.arange(H * W)
        px, py = -1 + 2 * (idx % W).double() / (W - 1), -1 + 2 * (idx // W).double() / (H - 1)
        sm = torch.softmax(logits.double(), -1)
        ref = torch.stack([(sm * px).sum(-1), (sm * py).sum(-1)], -1)
        assert (out.cpu().double() - ref).abs().max() < 2e-6


# ------------------------------------------------------------------ stem: normalise + pad, 7x7/s2 conv, statistics, weight gradient
_MEAN, _STD = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])


def _stem_case(dev, N, C, H, W, precision, frames, s2d=True):
    """frames: 'f32' | 'u8' ([N,C,H,W] uint8) | 'u8hwc' (the on-disk [N,H,W,C])"""
    _lib = _L()
    L = _lib.lib()
    _lib.check(L.lbc_set_fast_kernels(1 | (256 if s2d else 512)))
    try:
        _stem_case_body(_lib, L, dev, N, C, H, W, precision, frames, s2d)
    finally:
        _lib.check(L.lbc_set_fast_kernels(1 | 256))


def _stem_case_body(_lib, L, dev, N, C, H, W, precision, frames, s2d):
    g = torch.Generator().manual_seed(23)
    u8 = torch.randint(0, 256, (N, C, H, W), dtype=torch.uint8, generator=g)
    img = u8.float() / 255
    wgt = torch.randn(64, C, 7, 7, generator=g) * (2.0 / (49 * 64)) ** 0.5
    normalize = C == 3
    xn = (img - _MEAN.view(1, 3, 1, 1)) / _STD.view(1, 3, 1, 1) if normalize else img
    xr, wr = _r(xn, precision), _r(wgt, precision)
    if precision == 2:
        xr, wr = xn.double(), wgt.double()
    y_ref = F.conv2d(xr, wr, None, 2, 3)
    OH, OW = y_ref.shape[2:]
    dy = _r(torch.randn(N, 64, OH, OW, generator=g), precision)
    dw_ref = torch.nn.grad.conv2d_weight(xr, wgt.shape, dy.to(xr.dtype), 2, 3)
    d = lambda t: None if t is None else t.contiguous().to(dev)
    imgd = d(img) if frames == "f32" else None
    u8d = d(u8) if frames == "u8" else (d(u8.permute(0, 2, 3, 1)) if frames == "u8hwc" else None)
    wd, dyd = d(wgt), d(_nhwc(dy))
    direct = precision == 1 and C <= 8        # padded NHWC4 (camera) / NHWC8 (7-channel bird's-eye view) operand
    layout = L.lbc_stem_layout(C, W, int(normalize))          # 16: NHWC4 with every 2x2 pixel block contiguous (space to depth)
    assert layout == (8 if C > 4 else 16 if (s2d and W % 4 == 0) else 4) or not direct
    CH = 4 if C <= 4 else 8
    x4 = torch.empty(N, H + 6, W + 8, CH, device=dev) if direct else None
    y, dw = torch.empty(N, OH, OW, 64, device=dev), torch.empty(64, C, 7, 7, device=dev)
    stats = torch.empty(128, device=dev) if direct else None
    if precision == 1:
        tag = {4: "", 8: "<8ch>", 16: "<s2d>"}.get(layout, "")
        must = ["stem_conv_kernel" + tag, "stem_wgrad_kernel" + tag] if direct else ["stem_im2col_kernel", "conv_gemm_kernel<64>", "wgrad_gemm_kernel<64>"]
    elif precision == 2:
        must = ["tc_stem_im2col_kernel", "conv_gemm_kernel<64,f32>", "wgrad_gemm_kernel<64>"]
    else:
        must = []
    with Traced(dev, must, (REF_TAGS if precision == 1 else (("k_conv_fwd", "k_conv_wgrad_part") if precision == 2 else ()))):
        _lib.check(L.lbc_op_stem(_lib.ptr(imgd), _lib.ptr(u8d), 1 if frames == "u8hwc" else 0, _lib.ptr(wd), int(normalize), N, C,
                                 H, W, _lib.ptr(x4), _lib.ptr(y), _lib.ptr(stats), _lib.ptr(dyd), _lib.ptr(dw), precision, None))
    if direct:
        # a2: NormalizeV2 fused into the writer of the padded operand -- bit-exact bf16 rounding of (x-mean)/std
        pad = torch.zeros(N, H + 6, W + 8, CH)
        pad[:, 3:3 + H, 4:4 + W, :C] = _nhwc(xr)
        got = x4.cpu()
        if layout == 16:   # [N][(H+6)/2][(W+8)/2][row parity][column parity][4] -> [N][H+6][W+8][4]
            got = got.reshape(N, (H + 6) // 2, (W + 8) // 2, 2, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(N, H + 6, W + 8, 4)
        assert (got - pad).abs().max() <= 2 ** -7 * 3, (got - pad).abs().max()        # at most one bf16 ulp (division rounding)
        assert ((got - pad).abs() > 0).float().mean() < 1e-3
        yb = y.cpu()
        assert (stats[:64].cpu().double() - yb.double().sum((0, 1, 2))).abs().max() < 2e-3 * max(1.0, float(yb.abs().sum((0, 1, 2)).max()))
        assert (stats[64:].cpu().double() - (yb.double() ** 2).sum((0, 1, 2))).abs().max() < 2e-3 * float((yb.double() ** 2).sum((0, 1, 2)).max())
    assert _err(_nchw(y.cpu()), y_ref) < {0: 2e-5, 1: 1.2e-2, 2: 1e-5}[precision]
    assert _err(dw.cpu(), dw_ref) < {0: 2e-5, 1: 5e-3, 2: 3e-5}[precision]


@pytest.mark.parametrize("C", [3, 7])
def test_stem_cpu(backend, C):
    _stem_case(backend, 2, C, 12, 20, 0, "f32")
    _stem_case(backend, 2, C, 12, 20, 0, "u8hwc")


@pytest.mark.gpu
@pytest.mark.parametrize("s2d", [True, False], ids=["s2d", "rows7"])
@pytest.mark.parametrize("frames", ["f32", "u8", "u8hwc"])
def test_stem_kernels_student_gpu(backend, frames, s2d):
    _stem_case("cuda", 3, 3, 160, 384, 1, frames, s2d)


@pytest.mark.gpu
def test_stem_kernels_odd_shapes_gpu(backend):
    """batch tail (N not a multiple of the tile's image count) in both operand layouts; 4 input channels"""
    _stem_case("cuda", 5, 3, 32, 64, 1, "u8")
    _stem_case("cuda", 5, 3, 32, 64, 1, "u8", s2d=False)
    _stem_case("cuda", 2, 4, 32, 64, 1, "f32")


@pytest.mark.gpu
@pytest.mark.parametrize("frames", ["f32", "u8"])
def test_stem_kernels_teacher_gpu(backend, frames):
    _stem_case("cuda", 3, 7, 192, 192, 1, frames)


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,W", [(3, 160, 384), (7, 192, 192)])
def test_stem_fp32tc_gpu(backend, C, H, W):
    _stem_case("cuda", 2, C, H, W, 2, "f32")


# ------------------------------------------------------------------ decoder: ConvTranspose2d + bias + ReLU ; block-entry data gradient
def _deconv_case(dev, N, h, w, Cin, Cout, precision):
    """nn.ConvTranspose2d(Cin, Cout, 3, 2, 1, 1) + bias -> ReLU (image.py:39-46) through lbc_op_conv_dgrad (conv roles:
    Co = Cin, Ci = Cout, input plane 2h x 2w)"""
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(29)
    x = _r(torch.randn(N, Cin, h, w, generator=g), precision)
    wt = _r(torch.randn(Cin, Cout, 3, 3, generator=g) / (Cin * 2.25) ** 0.5, precision)
    bias = torch.randn(Cout, generator=g) * 0.2
    ref = F.relu(F.conv_transpose2d(x.double(), wt.double(), bias.double(), 2, 1, 1)) if precision == 2 else \
        F.relu(F.conv_transpose2d(x, wt, bias, 2, 1, 1))
    xd, wd, bd = _nhwc(x).to(dev), wt.contiguous().to(dev), bias.to(dev)
    y = torch.empty(N, 2 * h, 2 * w, Cout, device=dev)
    must = ["conv_gemm_kernel"] if precision >= 1 else []
    with Traced(dev, must, REF_TAGS if precision >= 1 else ()):
        _lib.check(L.lbc_op_conv_dgrad(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(y), N, 2 * h, 2 * w, Cout, Cin, 3, 2, 1, precision,
                                       _lib.ptr(bd), 1, None))
    assert _err(_nchw(y.cpu()), ref) < {0: 2e-5, 1: 1.2e-2, 2: 2e-5}[precision]


def test_deconv_bias_relu_cpu(backend):
    _deconv_case(backend, 2, 3, 4, 128, 64, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,h,w,Cin,Cout", [(2, 5, 12, 640, 256), (2, 10, 24, 256, 128), (3, 20, 48, 128, 64), (2, 6, 6, 640, 256)])
def test_deconv_bias_relu_kernels_gpu(backend, N, h, w, Cin, Cout):
    _deconv_case("cuda", N, h, w, Cin, Cout, 1)
    _deconv_case("cuda", N, h, w, Cin, Cout, 2)


def _block_dgrad_case(dev, N, H, W, Ci, Co, precision):
    _lib = _L()
    L = _lib.lib()
    g = torch.Generator().manual_seed(31)
    w1 = _r(torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5, precision)
    wd = _r(torch.randn(Co, Ci, 1, 1, generator=g) / Ci ** 0.5, precision)
    dy1 = _r(torch.randn(N, Co, H // 2, W // 2, generator=g), precision)
    dy2 = _r(torch.randn(N, Co, H // 2, W // 2, generator=g), precision)
    cv = (lambda t: t.double()) if precision == 2 else (lambda t: t)
    ref = torch.nn.grad.conv2d_input((N, Ci, H, W), cv(w1), cv(dy1), 2, 1) + torch.nn.grad.conv2d_input((N, Ci, H, W), cv(wd), cv(dy2), 2, 0)
    a, b, w1d, wdd = _nhwc(dy1).to(dev), _nhwc(dy2).to(dev), w1.contiguous().to(dev), wd.contiguous().to(dev)
    dx = torch.empty(N, H, W, Ci, device=dev)
    with Traced(dev, ["conv_gemm_kernel"] if precision >= 1 else [], REF_TAGS if precision >= 1 else ()):
        _lib.check(L.lbc_op_block_dgrad_ds(_lib.ptr(a), _lib.ptr(b), _lib.ptr(w1d), _lib.ptr(wdd), _lib.ptr(dx), N, H, W, Ci, Co,
                                           precision, None))
    assert _err(_nchw(dx.cpu()), ref) < {0: 2e-5, 1: 1.2e-2, 2: 3e-5}[precision]


def test_block_entry_dgrad_cpu(backend):
    _block_dgrad_case(backend, 2, 8, 12, 64, 128, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W,Ci,Co", [(3, 40, 96, 64, 128), (5, 20, 48, 128, 256), (9, 10, 24, 256, 512), (2, 48, 48, 64, 128)])
def test_block_entry_dgrad_kernels_gpu(backend, N, H, W, Ci, Co):
    _block_dgrad_case("cuda", N, H, W, Ci, Co, 1)
    _block_dgrad_case("cuda", N, H, W, Ci, Co, 2)


# ------------------------------------------------------------------ conv epilogue: centring shift + BatchNorm statistics partials
@pytest.mark.gpu
@pytest.mark.parametrize("case", [(3, 40, 96, 64, 64, 3, 1), (3, 40, 96, 64, 128, 3, 2), (3, 40, 96, 64, 128, 1, 2),
                                  (9, 10, 24, 256, 256, 3, 1), (33, 5, 12, 512, 512, 3, 1), (2, 48, 48, 64, 64, 3, 1),
                                  (3, 20, 48, 128, 128, 3, 1, "row"), (9, 10, 24, 256, 256, 3, 1, "row"),
                                  (3, 24, 24, 128, 128, 3, 1, "row"), (300, 20, 48, 128, 128, 3, 1, "row"),
                                  (3, 40, 96, 64, 64, 3, 1, "row64"), (40, 40, 96, 64, 64, 3, 1, "row64")])
def test_conv_epilogue_statistics_gpu(backend, case):
    """sum / sum of squares of the STORED bf16 output, emitted by the GEMM epilogue, vs the column sums of that output;
    the per-channel shift is added before rounding"""
    _lib = _L()
    L = _lib.lib()
    row = len(case) == 8   # the shared-row CTA-pair kernel (register statistics; 300 images: several tiles per CTA)
    N, H, W, Ci, Co, K, s = case[:7]
    p = K // 2
    g = torch.Generator().manual_seed(37)
    x = _r(torch.randn(N, Ci, H, W, generator=g), 1)
    wt = _r(torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5, 1)
    shift = torch.randn(Co, generator=g) * 0.5
    ref = F.conv2d(x, wt, shift, s, p)
    xd, wd, sd = _nhwc(x).cuda(), wt.contiguous().cuda(), shift.cuda()
    y = torch.empty(N, ref.shape[2], ref.shape[3], Co, device="cuda")
    stats = torch.empty(2 * Co, device="cuda")
    from test_ops import _variant_default
    row64 = row and case[7] == "row64"
    _lib.check(L.lbc_set_fast_kernels(1 | (65536 if row64 else 131072) | ((1024 | 4096) if row else (2048 | 8192))))
    try:
        with Traced("cuda", ["conv_row_kernel<64>" if row64 else "conv_row_kernel<128>" if row else
                             "conv3x3_c64_kernel" if (Ci == Co == 64 and K == 3) else "conv_gemm_kernel", "col_finalize_kernel"]):
            _lib.check(L.lbc_op_conv_fwd(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(y), N, H, W, Ci, Co, K, s, p, 1, _lib.ptr(sd),
                                         _lib.ptr(stats), None))
    finally:
        _lib.check(L.lbc_set_fast_kernels(1 | _variant_default()))
    yb = y.cpu().double()
    assert _err(_nchw(y.cpu()), ref) < 1.2e-2
    assert (stats[:Co].cpu().double() - yb.sum((0, 1, 2))).abs().max() < 1e-3 * max(1.0, float(yb.abs().sum((0, 1, 2)).max()))
    assert (stats[Co:].cpu().double() - (yb ** 2).sum((0, 1, 2))).abs().max() < 1e-3 * float((yb ** 2).sum((0, 1, 2)).max())
