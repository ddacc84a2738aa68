"""Whole-step parity: the drop-in classes + native engine against golden vectors from the reference
(tests/golden/, made by oracle/make_golden.py).  fp32 = parity mode, tolerance 1e-3 relative (north_star);
gradients are compared at the tolerance the fp32 oracle itself holds against an fp64 run of the same graph
(per-tensor ~5e-3: see DESIGN.md "parity").  CPU variants run the host-emulation build; gpu variants the B200."""
import os

import numpy as np
import pytest
import torch

from lbc_testing import (batch_on, build_models, check_init, check_step0_against_golden, gold, rel_err,
                         run_student_steps)

OUT_TOL = 1e-3       # north_star: <= 1e-3 relative fp32
LOSS_TOL = 1e-3


def _student_case(device, B, phase, precision="fp32"):
    g = gold("student_B%d_phase%d.npz" % (B, phase))
    s, t, recs = run_student_steps(device, precision, B, phase, 2)
    if precision == "fp32tc":
        # tensor-core parity mode: the north-star tolerance on every output and on the loss; gradients at the distance the
        # fp32 oracle itself keeps from an fp64 run of the same graph
        print("fp32tc B=%d phase %d: pred %.2e preds %.2e loss %.2e" % (
            B, phase, rel_err(recs[0]["pred"], g["step0/pred"]), rel_err(recs[0]["preds"], g["step0/preds"]),
            abs(recs[0]["loss_mean"] - float(g["step0/loss_mean"])) / float(g["step0/loss_mean"])))
        if phase == 0:
            check_step0_against_golden(recs[0], g, OUT_TOL, LOSS_TOL, 5e-3, 5e-2)
        else:
            check_step0_against_golden(recs[0], g, OUT_TOL, 2e-3 + LOSS_TOL, 2e-2, 2e-1)
    elif phase == 0:
        check_step0_against_golden(recs[0], g, 1e-4, 1e-5, 1e-3, 2e-2)
    else:   # phase-1 transform is singular at the horizon (train_image_phase1.py:55): looser
        check_step0_against_golden(recs[0], g, 1e-4, 2e-3, 2e-2, 1e-1)
    # BN running buffers after one step (a21) and num_batches_tracked
    post = recs[0]["post"]
    for k, v in post.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(v.numpy(), g["step0/post/%s/full" % k], rtol=1e-3 if precision == "fp32tc" else 1e-4,
                                       atol=1e-5 if precision == "fp32tc" else 1e-6, err_msg=k)
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(g["step0/post/" + k]) == 1
    # Adam step 1 moves every trained element by ~lr*sign(g) (bias-corrected first step): bounded difference
    for k, v in post.items():
        key = "step0/post/%s/full" % k
        if key in g.files and not k.endswith(("running_mean", "running_var")):
            assert np.abs(v.numpy() - g[key]).max() <= 2.05e-4, k
    # second step: loss trajectory stays within Adam's sign-flip chaos
    assert abs(recs[1]["loss_mean"] - float(g["step1/loss_mean"])) <= (5e-3 if phase == 0 else 1e-1) * float(g["step1/loss_mean"])
    return s, t


def test_student_phase0_B2_cpu(backend):
    _student_case(backend, 2, 0)


def test_eval_mode_and_state_dict_roundtrip_cpu(backend):
    dev = backend
    s, t = build_models(dev, "fp32")
    check_init(s, "student")
    b = batch_on(dev, 2)
    import learningbycheating_b200 as lbc
    oh = lbc.one_hot(b["command"].cpu()).to(dev)
    s.eval()
    with torch.no_grad():
        p1, ps1 = s(b["rgb"], b["speed"], oh)
    # eval BN uses running stats: compare with the oracle's eval path
    import lbc_oracle as orc
    sd = {k: v.cpu() for k, v in s.state_dict().items()}
    po, pso, _ = orc.policy_forward(sd, b["rgb"].cpu(), b["speed"].cpu(), oh.cpu(), "resnet34", False, True)
    assert rel_err(p1.cpu(), po) < 1e-4
    assert rel_err(ps1.cpu(), pso) < 1e-4
    # load_state_dict keeps the flat views; a fresh module loaded from the dict reproduces the output
    s2 = lbc.ImagePolicyModelSS('resnet34', all_branch=True, lbc_precision="fp32").to(dev)
    s2.load_state_dict(s.state_dict())
    s2.eval()
    with torch.no_grad():
        p2, _ = s2(b["rgb"], b["speed"], oh)
    assert torch.equal(p1, p2)
    # all_branch=False returns a single tensor (image.py:86-89)
    s2.all_branch = False
    with torch.no_grad():
        assert torch.is_tensor(s2(b["rgb"], b["speed"], oh))


def test_uint8_frames_match_float_frames_cpu(backend):
    """frames fed as uint8 ([B,C,H,W] or the on-disk [B,H,W,C]) give bit-identical results to
    torchvision ToTensor (u8/255) done on the host."""
    import learningbycheating_b200 as lbc
    dev = backend
    s, _ = build_models(dev, "fp32")
    s.eval()
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 3, 160, 384), dtype=torch.uint8, generator=g)
    speed, cmd = torch.rand(2, generator=g) * 10, torch.tensor([1., 3.])
    oh = lbc.one_hot(cmd).to(dev)
    with torch.no_grad():
        pf, _ = s((u8.float() / 255).to(dev), speed.to(dev), oh)
        p0, _ = s(u8.to(dev), speed.to(dev), oh)
        p1, _ = s(u8.permute(0, 2, 3, 1).contiguous().to(dev), speed.to(dev), oh)
    assert torch.equal(pf, p0) and torch.equal(pf, p1)


def test_torch_adam_also_works_cpu(backend):
    """Parameters are ordinary leaf nn.Parameters: the reference's own torch.optim.Adam drives the module too
    and agrees with the fused native Adam."""
    _, _, ra = run_student_steps(backend, "fp32", 2, 0, 1, optimizer="lbc")
    _, _, rb = run_student_steps(backend, "fp32", 2, 0, 1, optimizer="torch")
    for k in ra[0]["post"]:
        a, b = ra[0]["post"][k], rb[0]["post"][k]
        if a.dtype.is_floating_point:
            assert (a - b).abs().max() <= 1e-7, k


def test_input_validation_cpu(backend):
    from learningbycheating_b200 import LbcError
    s, _ = build_models(backend, "fp32")
    with pytest.raises(LbcError):
        s(torch.zeros(2, 3, 100, 100, device=backend), torch.zeros(2, device=backend), torch.zeros(2, 4, device=backend))
    with pytest.raises(LbcError):
        s(torch.zeros(2, 3, 160, 384, device=backend), torch.zeros(3, device=backend), torch.zeros(2, 4, device=backend))


def test_birdview_training_cpu(backend):
    """config 5: train_birdview.train_or_eval over the teacher architecture."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_birdview as tb
    dev = backend
    _, t = build_models(dev, "fp32", teacher_all_branch=False)
    g = gold("birdview_B2.npz")
    b = batch_on(dev, 2)
    opt = lbc.Adam(t.parameters(), lr=1e-4)
    data = [(b["birdview"], b["location"], b["command"].cpu(), b["speed"])] * 2
    ls = tb.train_or_eval(tb.LocationLoss(choice='l1'), t, data, opt, True, dict(device=dev, log_iterations=1000), False)
    assert abs(float(ls[0]) - float(g["step0/loss_mean"])) < 1e-5 * float(g["step0/loss_mean"])
    assert abs(float(ls[1]) - float(g["step1/loss_mean"])) < 5e-3 * float(g["step1/loss_mean"])


def test_stale_forward_and_copies_cpu(backend):
    """The engine keeps ONE set of activations: backward through an older train-mode forward, or twice through the same
    one, must raise instead of differentiating the wrong activations.  deepcopy / torch.save of a module that has already
    run drop the native handle and give an independent, working copy."""
    import copy
    import io
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import LbcError
    dev = backend
    s, _ = build_models(dev, "fp32")
    s.train()
    b = batch_on(dev, 2)
    oh = lbc.one_hot(b["command"].cpu()).to(dev)
    p1, _ = s(b["rgb"], b["speed"], oh)
    p2, _ = s(b["rgb"], b["speed"], oh)
    with pytest.raises(LbcError, match="activations are gone"):
        p1.sum().backward()
    p2.sum().backward()
    with pytest.raises(LbcError, match="twice"):
        p2.sum().backward()
    c = copy.deepcopy(s)
    buf = io.BytesIO()
    torch.save(s, buf)
    buf.seek(0)
    s3 = torch.load(buf, weights_only=False)
    for m in (s, c, s3):
        m.eval()
    with torch.no_grad():
        a, b2, c2 = (m(b["rgb"], b["speed"], oh)[0] for m in (s, c, s3))
    assert torch.equal(a, b2) and torch.equal(a, c2)
    assert c.conv.conv1.weight.data_ptr() != s.conv.conv1.weight.data_ptr()
    # Adam state: moments loaded on another device follow the parameters; a foreign-size state is rejected
    opt = lbc.Adam(s.parameters(), lr=1e-4)
    s.train()
    s(b["rgb"], b["speed"], oh)[0].sum().backward()
    opt.step()
    sd = opt.state_dict()
    opt2 = lbc.Adam(s.parameters(), lr=1e-4)
    opt2.load_state_dict(dict(sd, exp_avg=sd["exp_avg"].clone(), exp_avg_sq=sd["exp_avg_sq"].clone()))
    assert opt2.step_count == 1
    opt3 = lbc.Adam(s.parameters(), lr=1e-4)
    opt3.load_state_dict(dict(sd, exp_avg=sd["exp_avg"][:10].clone(), exp_avg_sq=sd["exp_avg_sq"][:10].clone()))
    s.zero_grad()
    s(b["rgb"], b["speed"], oh)[0].sum().backward()
    with pytest.raises(LbcError, match="moment buffers"):
        opt3.step()


# ------------------------------------------------------------------ tap by tap (lbc_net_read_tap vs the golden step0/tap/* entries)
# relative error allowed on the sampled values of each tap (fraction of the tap's RMS): fp32 parity mode, and the bf16
# throughput mode whose storage rounding is amplified layer by layer (DESIGN.md "bf16 mode: measured deviation")
def _tap_bound(name, precision):
    if precision == "fp32tc":
        return 1e-3
    if precision != "bf16":
        return 2e-4
    # measured on the B200 (worst of 24 sampled values / tap RMS, B = 2 and 4): stem 0.5 %, layer1 2.7 %, layer2 8 %,
    # layer3 15 %, layer4 38 %, decoder 34-93 %, logits 80 % -- bf16 storage noise amplified by the train-mode BatchNorms of
    # a randomly initialised net at these tiny batches; the same taps in fp32tc sit at <= 2.5e-4
    if name.startswith("stem"):
        return 0.01
    for key, b in (("layer1", 0.05), ("layer2", 0.13), ("layer3", 0.25), ("layer4", 0.60), ("deconv", 1.5), ("logits", 1.3)):
        if key in name:
            return b
    raise KeyError(name)


def _check_taps(net, g, precision):
    names = sorted({k.split("/")[2] for k in g.files if k.startswith("step0/tap/")})
    assert len(names) >= 22, names
    worst = []
    for name in names:
        shape = tuple(int(v) for v in g["step0/tap/%s/shape" % name])
        t = net.lbc_read_tap(name, int(np.prod(shape))).cpu().double().reshape(shape)
        rms = float(g["step0/tap/%s/sqmean" % name]) ** 0.5
        coords = g["step0/tap/%s/coords" % name]
        got = t[tuple(torch.from_numpy(coords.T))]
        e_vals = float((got - torch.from_numpy(g["step0/tap/%s/vals" % name])).abs().max()) / rms
        e_sq = abs(float((t * t).mean()) - float(g["step0/tap/%s/sqmean" % name])) / float(g["step0/tap/%s/sqmean" % name])
        e_abs = abs(float(t.abs().mean()) - float(g["step0/tap/%s/absmean" % name])) / float(g["step0/tap/%s/absmean" % name])
        b = _tap_bound(name, precision)
        worst.append((e_vals / b, name, e_vals, e_sq, e_abs))
    print("taps[%s]: sampled-value error / RMS per tap: %s" % (precision, ", ".join("%s %.2e" % (w[1], w[2]) for w in worst)))
    for _, name, e_vals, e_sq, e_abs in worst:
        b = _tap_bound(name, precision)
        assert e_vals <= b, (name, e_vals, b)
        assert e_sq <= b and e_abs <= b, (name, e_sq, e_abs, b)


def _tap_case(device, precision, B=2):
    import learningbycheating_b200 as lbc
    g = gold("student_B%d_phase0.npz" % B)
    s, _ = build_models(device, precision)
    s.train()
    b = batch_on(device, B)
    with torch.no_grad():
        s(b["rgb"], b["speed"], lbc.one_hot(b["command"].cpu()).to(device))
    _check_taps(s, g, precision)


def test_taps_match_golden_cpu(backend):
    _tap_case(backend, "fp32")


# ------------------------------------------------------------------ on the B200
@pytest.mark.gpu
@pytest.mark.parametrize("B,phase", [(2, 0), (4, 0), (2, 1), (4, 1)])
def test_student_step_gpu_fp32(backend, B, phase):
    assert backend == "cuda"
    _student_case("cuda", B, phase)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp32tc"])
def test_config1_plumbing_B8_gpu(backend, precision):
    """BASELINE config 1: the reference's phase-0 loop at batch 8 -- three DIFFERENT batches through
    train_image_phase0.train_or_eval of this package (engine + fused Adam) and through the oracle port of the reference loop
    (training/train_image_phase0.py:152-209): same loss sequence, same final weights up to Adam's sign noise."""
    import lbc_oracle as orc
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_image_phase0 as p0
    B, steps = 8, 3
    s, t = build_models("cuda", precision)
    sd0 = {k: v.detach().cpu().clone() for k, v in s.state_dict().items()}
    td0 = {k: v.detach().cpu().clone() for k, v in t.state_dict().items()}
    batches = [orc.synthetic_batch(B, seed=11 + i) for i in range(steps)]
    data = [(b["rgb"], b["birdview"], b["location"], b["command"], b["speed"]) for b in batches]
    opt = lbc.Adam(s.parameters(), lr=1e-4)
    t.eval()
    ls = p0.train_or_eval(p0.CoordConverter(device="cuda"), p0.LocationLoss(device="cuda"), s, t, data, opt, True,
                          dict(device="cuda", log_iterations=1000), False)
    ls = [float(x) for x in ls]
    osd, ost = orc.leafify(sd0), orc.new_adam_state()
    ref = [float(orc.train_step(osd, td0, b["rgb"], b["birdview"], b["speed"], b["command"], 0, adam_state=ost)["loss_mean"])
           for b in batches]
    print("config1 B=8 [%s] loss sequence: engine %s  reference loop %s" % (precision, ["%.6f" % x for x in ls], ["%.6f" % x for x in ref]))
    tol0 = 1e-5 if precision == "fp32" else LOSS_TOL
    assert abs(ls[0] - ref[0]) <= tol0 * ref[0]
    for a, b in zip(ls[1:], ref[1:]):      # later steps: Adam's first updates are lr*sign(g) -- sign noise on ~zero gradients
        assert abs(a - b) <= 5e-3 * b
    # final weights: the accumulated update of every trained tensor points the same way as the reference's
    post = {k: v.detach().cpu() for k, v in s.state_dict().items()}
    num = den_a = den_b = 0.0
    for k, w0 in sd0.items():
        if not w0.dtype.is_floating_point or k.endswith(("running_mean", "running_var")) or k.startswith("conv.fc"):
            continue
        if k.startswith("location_pred.") and (k.endswith(".0.bias") or k.endswith(".1.bias")):
            continue                         # exactly-zero gradients (softmax shift invariance): pure rounding noise
        da, db = (post[k] - w0).double().flatten(), (osd[k].detach() - w0).double().flatten()
        assert float((da - db).abs().max()) <= 2.05e-4 * steps, k
        num += float(da @ db)
        den_a += float(da @ da)
        den_b += float(db @ db)
    cos = num / (den_a * den_b) ** 0.5
    print("config1 B=8 [%s]: cosine(update, reference update) = %.5f" % (precision, cos))
    assert cos > 0.98
    for k in ("conv.bn1.running_mean", "conv.layer4.2.bn2.running_var", "deconv.0.running_mean"):
        # (buffers of steps 2-3 see weights that already differ by Adam's sign noise)
        np.testing.assert_allclose(post[k].numpy(), osd[k].detach().numpy(), rtol=2e-2, atol=1e-4, err_msg=k)
    assert int(post["conv.bn1.num_batches_tracked"]) == steps


@pytest.mark.gpu
@pytest.mark.parametrize("B,phase", [(2, 0), (4, 0), (2, 1), (4, 1)])
def test_student_step_gpu_fp32tc(backend, B, phase):
    """LBC_PREC_F32TC: the golden step through the tcgen05 split-precision convolutions at the north-star tolerance."""
    assert backend == "cuda"
    from learningbycheating_b200 import _lib
    from test_kernels import Traced
    with Traced("cuda", ["conv_gemm_kernel<", "wgrad", "tc_stem_im2col_kernel"], ("k_conv_fwd", "k_conv_dgrad", "k_conv_wgrad_part")) as tr:
        _student_case("cuda", B, phase, "fp32tc")
    assert not [k for k in tr.counts if k.startswith("conv_gemm_kernel<") and not k.endswith("f32>")], sorted(tr.counts)


@pytest.mark.gpu
def test_student_step_gpu_fp32tc_B32_vs_oracle(backend):
    """one B=32 train step (forward, phase-0 loss, backward) of the fp32tc engine against the CPU oracle"""
    import lbc_oracle as orc
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_image_phase0 as p0
    B = 32
    torch.manual_seed(0)
    s = lbc.ImagePolicyModelSS("resnet34", all_branch=True, lbc_precision="fp32tc")
    sd0 = {k: v.clone() for k, v in s.state_dict().items()}
    s = s.to("cuda").train()
    b = orc.synthetic_batch(B)
    oh = lbc.one_hot(b["command"])
    target = torch.rand(B, 5, 2, generator=torch.Generator().manual_seed(4)) * torch.tensor([384.0, 160.0])
    pred, preds = s(b["rgb"].cuda(), b["speed"].cuda(), oh.cuda())
    loss = p0.LocationLoss(device="cuda")(pred, target.cuda()).mean()
    loss.backward()
    osd = orc.leafify(sd0)
    op, ops, _ = orc.policy_forward(osd, b["rgb"], b["speed"], oh, "resnet34", True, True)
    ol = orc.phase0_loss(op, target).mean()
    ol.backward()
    e_pred, e_preds = rel_err(pred.detach().cpu(), op.detach()), rel_err(preds.detach().cpu(), ops.detach())
    e_loss = abs(loss.item() - ol.item()) / abs(ol.item())
    g_ref = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in osd.values() if v.requires_grad and v.grad is not None))
    g_ours = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in s.parameters() if p.grad is not None)).cpu()
    e_grad = abs(g_ours.item() - g_ref.item()) / g_ref.item()
    print("fp32tc B=32 vs oracle: pred %.2e preds %.2e loss %.2e grad-norm %.2e" % (e_pred, e_preds, e_loss, e_grad))
    assert e_pred <= OUT_TOL and e_preds <= OUT_TOL and e_loss <= LOSS_TOL and e_grad <= 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("precision,B", [("fp32", 2), ("fp32", 4), ("bf16", 2), ("bf16", 4), ("fp32tc", 2), ("fp32tc", 4)])
def test_taps_match_golden_gpu(backend, precision, B):
    """per-layer bound on every tapped activation (replaces a single loose bound on the waypoints for bf16)"""
    _tap_case("cuda", precision, B)


@pytest.mark.gpu
def test_eval_roundtrip_validation_gpu(backend):
    test_eval_mode_and_state_dict_roundtrip_cpu("cuda")
    test_uint8_frames_match_float_frames_cpu("cuda")
    test_input_validation_cpu("cuda")
    test_torch_adam_also_works_cpu("cuda")
    test_birdview_training_cpu("cuda")
    test_stale_forward_and_copies_cpu("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("train", [False, True])
def test_uint8_frames_bf16_direct_stem_gpu(backend, train):
    """bf16 fast path: uint8 frames go straight into the padded NHWC4 stem operand (no fp32 image on the device); the
    result must be bit-identical to feeding ToTensor'ed float frames."""
    assert backend == "cuda"
    import learningbycheating_b200 as lbc
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (3, 3, 160, 384), dtype=torch.uint8, generator=g)
    speed, cmd = torch.rand(3, generator=g) * 10, torch.tensor([1., 3., 4.])
    oh = lbc.one_hot(cmd).to("cuda")
    outs = []
    for frames in ((u8.float() / 255), u8, u8.permute(0, 2, 3, 1).contiguous()):
        # a fresh engine per variant: a train-mode forward moves the BN buffers and the engine's centring estimate
        s, _ = build_models("cuda", "bf16")
        s.train(train)
        with torch.no_grad():
            outs.append(s(frames.to("cuda"), speed.to("cuda"), oh)[0].clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [0, 1])
def test_student_step_gpu_bf16_deviation(backend, fast):
    """bf16 throughput mode: measured deviation from the fp32 reference is reported, and bounded loosely
    (stock bf16 autocast deviates ~9% on waypoints at random init)."""
    from learningbycheating_b200 import _lib
    _lib.check(_lib.lib().lbc_set_fast_kernels(fast))
    try:
        g = gold("student_B4_phase0.npz")
        _, _, recs = run_student_steps("cuda", "bf16", 4, 0, 1)
        e_pred = rel_err(recs[0]["pred"], g["step0/pred"])
        e_loss = abs(recs[0]["loss_mean"] - float(g["step0/loss_mean"])) / float(g["step0/loss_mean"])
        gl2 = sum(float(v.double().norm()) ** 2 for v in recs[0]["grads"].values() if v is not None) ** 0.5
        print("bf16 (fast=%d) deviation: pred %.3e  loss %.3e  grad_l2 %.4f vs %.4f" %
              (fast, e_pred, e_loss, gl2, float(g["step0/grad_global_l2"])))
        assert e_pred < 0.25 and e_loss < 0.05
        assert abs(gl2 - float(g["step0/grad_global_l2"])) < 0.2 * float(g["step0/grad_global_l2"])
    finally:
        _lib.check(_lib.lib().lbc_set_fast_kernels(1))


def _grads_with_fixed_upstream(precision, fast, B=4):
    """forward + backward with a FIXED upstream gradient (no sign() discontinuity of the L1 loss in the way)."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import _lib
    _lib.check(_lib.lib().lbc_set_fast_kernels(fast))
    try:
        s, _ = build_models("cuda", precision)
        s.train()
        b = batch_on("cuda", B)
        oh = lbc.one_hot(b["command"].cpu()).to("cuda")
        with torch.no_grad():
            s(b["rgb"], b["speed"], oh)     # warm-up pass: populates the bf16 path's per-channel centring estimates
        pred, preds = s(b["rgb"], b["speed"], oh)
        g = torch.Generator().manual_seed(5)
        r1 = torch.randn(pred.shape, generator=g).cuda()
        r2 = torch.randn(preds.shape, generator=g).cuda()
        torch.autograd.backward([pred, preds], [r1, r2])
        return pred.detach().cpu(), {k: p.grad.detach().clone().cpu() for k, p in s.named_parameters() if p.grad is not None}
    finally:
        _lib.check(_lib.lib().lbc_set_fast_kernels(1))


@pytest.mark.gpu
def test_bf16_fast_kernels_match_correctness_first_kernels_gpu(backend):
    """Same bf16 step through the tcgen05 / fused kernels and through the correctness-first kernels, and the fp32
    parity path as the yardstick: an indexing bug in any fast kernel gives O(1) differences, bf16 rounding does not."""
    p32, g32 = _grads_with_fixed_upstream("fp32", 1)
    p0, g0 = _grads_with_fixed_upstream("bf16", 0)
    p1, g1 = _grads_with_fixed_upstream("bf16", 1)
    print("pred deviation vs fp32: correctness-first bf16 %.3e, fast bf16 %.3e" % (rel_err(p0, p32), rel_err(p1, p32)))
    rows = []
    for k, ref in g32.items():
        n = float(ref.double().norm())
        if n < 1e-6:
            continue
        e0 = float((g0[k].double() - ref.double()).norm()) / n
        e1 = float((g1[k].double() - ref.double()).norm()) / n
        rows.append((e1, e0, k))
    rows.sort(reverse=True)
    print("per-tensor gradient error vs fp32 (fast bf16, correctness-first bf16, name), worst 6:")
    for r in rows[:6]:
        print("   %.3f %.3f %s" % r)
    import statistics
    med1 = statistics.median(r[0] for r in rows)
    med0 = statistics.median(r[1] for r in rows)
    print("median error: fast %.4f correctness-first %.4f" % (med1, med0))
    # Both bf16 kernel families sit at the SAME distance from fp32: that distance is the bf16 *storage* noise of this
    # ill-conditioned graph at random init and B=4 (DESIGN.md "bf16 mode: measured deviation"), not a kernel property.
    assert med1 < 1.25 * med0 + 0.05
    assert rows[0][0] < 1.5 * max(r[1] for r in rows) + 0.1
    assert rel_err(p1, p32) < 1.5 * rel_err(p0, p32) + 0.05


@pytest.mark.gpu
def test_birdview_training_bf16_gpu(backend):
    """config 5 architecture (ResNet-18, 7-channel 192x192 input, 48x48 heads) through the tcgen05 path: two Adam steps,
    first-step loss within bf16 noise of the fp32 golden, both kernel families agree, loss goes down."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import _lib, train_birdview as tb
    g = gold("birdview_B2.npz")
    out = {}
    for fast in (0, 1):
        _lib.check(_lib.lib().lbc_set_fast_kernels(fast))
        try:
            _, t = build_models("cuda", "bf16", teacher_all_branch=False)
            b = batch_on("cuda", 2)
            opt = lbc.Adam(t.parameters(), lr=1e-4)
            data = [(b["birdview"], b["location"], b["command"].cpu(), b["speed"])] * 3
            ls = tb.train_or_eval(tb.LocationLoss(choice='l1'), t, data, opt, True, dict(device="cuda", log_iterations=1000), False)
            out[fast] = [float(x) for x in ls]
        finally:
            _lib.check(_lib.lib().lbc_set_fast_kernels(1))
    ref0 = float(g["step0/loss_mean"])
    for fast in (0, 1):
        assert abs(out[fast][0] - ref0) < 0.02 * ref0, out
        assert out[fast][2] < out[fast][0], out
    assert abs(out[0][1] - out[1][1]) < 0.05 * out[0][1], out


@pytest.mark.gpu
def test_full_size_properties_gpu(backend):
    """BASELINE config 2 size (B=256, bf16): size-independent properties -- finite outputs in [-1,1], loss
    decreases over Adam steps on a fixed batch, BN counters advance, gradients finite, conv.fc untouched."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_image_phase0 as p0
    dev = "cuda"
    s, t = build_models(dev, "bf16")
    fc0 = s.conv.fc.weight.detach().clone()
    b = batch_on(dev, 256)
    opt = lbc.Adam(s.parameters(), lr=1e-4)
    data = [(b["rgb"], b["birdview"], b["location"], b["command"].cpu(), b["speed"])] * 6
    t.eval()
    ls = p0.train_or_eval(p0.CoordConverter(device=dev), p0.LocationLoss(device=dev), s, t, data, opt, True,
                          dict(device=dev, log_iterations=1000), False)
    ls = [float(x) for x in ls]
    assert all(np.isfinite(ls)), ls
    assert ls[-1] < ls[0], ls
    assert int(s.conv.bn1.num_batches_tracked) == 6
    for k, p in s.named_parameters():
        if k.startswith("conv.fc"):
            assert p.grad is None
        else:
            assert torch.isfinite(p.grad).all(), k
    assert torch.equal(s.conv.fc.weight, fc0)
    s.eval()
    with torch.no_grad():
        p, ps = s(b["rgb"], b["speed"], lbc.one_hot(b["command"].cpu()).to(dev))
    assert p.abs().max() <= 1.0 and ps.abs().max() <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("B", [8, 256])
def test_launch_schedule_does_not_change_results_gpu(backend, B):
    """lbc_set_schedule: weight gradients on the side stream (with and without the high-priority chain stream) and
    programmatic dependent launch reorder / overlap kernels, they must not change a single result beyond the fp32
    atomics of the split-K weight gradients: same batch, same weights, four schedules, three steps each (the second and
    third step start while the side stream may still hold work of the previous one unless the joins are right)."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import _lib, train_image_phase0 as p0
    dev = "cuda"
    b = batch_on(dev, B)
    oh = lbc.one_hot(b["command"].cpu()).to(dev)
    target = (torch.rand(B, 5, 2, generator=torch.Generator().manual_seed(5)) * torch.tensor([384.0, 160.0])).to(dev)

    def run(ovl, pdl):
        _lib.set_schedule(ovl, pdl)
        try:
            s, _ = build_models(dev, "bf16")
            s.train()
            opt = lbc.Adam(s.parameters(), lr=1e-4)
            out = []
            for _ in range(3):
                pred, preds = s(b["rgb"], b["speed"], oh)
                loss = p0.LocationLoss(device=dev)(pred, target).mean()
                opt.zero_grad()
                loss.backward()
                grads = {k: q.grad.detach().clone() for k, q in s.named_parameters() if q.grad is not None}
                opt.step()
                out.append((pred.detach().clone(), float(loss), grads))
            torch.cuda.synchronize()
            return out
        finally:
            _lib.set_schedule(int(os.environ.get("LBC_WGRAD_OVERLAP", "1")), int(os.environ.get("LBC_PDL", "1")))

    ref = run(0, 0)
    for ovl, pdl in ((1, 0), (2, 0), (0, 1), (2, 1)):
        got = run(ovl, pdl)
        for step, ((p0_, l0, g0), (p1_, l1, g1)) in enumerate(zip(ref, got)):
            tag = "schedule (%d, %d) step %d" % (ovl, pdl, step)
            # step 0 sees identical weights: only atomics order differs; later steps inherit that through Adam (lr 1e-4)
            tol = 1e-5 if step == 0 else 2e-2
            assert (p0_ - p1_).abs().max().item() <= tol, tag
            assert abs(l0 - l1) <= tol * max(1.0, abs(l0)), tag
            if step == 0:
                for k in g0:
                    d = (g0[k] - g1[k]).abs().max().item()
                    assert d <= 1e-4 * max(g0[k].abs().max().item(), 1e-6), (tag, k, d)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16", "fp32tc", "fp32"])
def test_inference_graph_gpu(backend, precision, monkeypatch):
    """small eval batches go through lbc_net_infer (one CUDA-graph replay per call).  Same bits as the
    eager eval forward; follows parameter changes made by load_state_dict, in-place edits and lbc.Adam; float and uint8
    frames; batch sizes interleaved."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import _lib
    dev = "cuda"
    s, _ = build_models(dev, precision)
    s.eval()
    b = batch_on(dev, 2)
    oh = lbc.one_hot(b["command"].cpu()).to(dev)
    u8 = (b["rgb"] * 255).round().to(torch.uint8)
    replays = lambda: _lib.lib().lbc_net_infer_replays(s._lbc.handle)

    def eager(x, n):
        monkeypatch.setenv("LBC_B200_INFER_GRAPH_MAX_B", "0")
        with torch.no_grad():
            out = [t.clone() for t in s(x[:n], b["speed"][:n], oh[:n])]
        monkeypatch.setenv("LBC_B200_INFER_GRAPH_MAX_B", "16")
        return out

    def graph(x, n):
        with torch.no_grad():
            return [t.clone() for t in s(x[:n], b["speed"][:n], oh[:n])]

    ref1 = eager(b["rgb"], 1)
    r0 = replays()
    first = graph(b["rgb"], 1)            # eager run + capture
    again = graph(b["rgb"], 1)            # replay
    assert replays() == r0 + 1, "the CUDA-graph path did not run"
    for a, c, d in zip(ref1, first, again):
        assert torch.equal(a, c) and torch.equal(a, d)
    # another input through the same graph; B = 2 gets its own graph; uint8 frames ([B,H,W,C]) too
    flipped = b["rgb"].flip(0)
    assert all(torch.equal(a, c) for a, c in zip(eager(flipped, 1), graph(flipped, 1)))
    graph(b["rgb"], 2)
    assert all(torch.equal(a, c) for a, c in zip(eager(b["rgb"], 2), graph(b["rgb"], 2)))
    hwc = u8.permute(0, 2, 3, 1).contiguous()
    graph(hwc, 1)
    assert all(torch.equal(a, c) for a, c in zip(eager(hwc, 1), graph(hwc, 1)))
    assert all(torch.equal(a, c) for a, c in zip(ref1, graph(b["rgb"], 1)))
    # parameter changes: in place, load_state_dict, the package's Adam after a training step
    with torch.no_grad():
        s.location_pred[0][1].weight.mul_(1.5)
    changed = graph(b["rgb"], 1)
    assert not torch.equal(changed[1], ref1[1]) and all(torch.equal(a, c) for a, c in zip(eager(b["rgb"], 1), changed))
    sd = {k: v.clone() for k, v in s.state_dict().items()}
    sd["conv.conv1.weight"] = sd["conv.conv1.weight"] * 0.5
    s.load_state_dict(sd)
    assert all(torch.equal(a, c) for a, c in zip(eager(b["rgb"], 1), graph(b["rgb"], 1)))
    opt = lbc.Adam(s.parameters(), lr=1e-3)
    s.train()
    pred, _ = s(b["rgb"], b["speed"], oh)
    pred.abs().mean().backward()
    s.eval()
    before = graph(b["rgb"], 1)
    opt.step()
    r1 = replays()
    after = graph(b["rgb"], 1)
    assert replays() == r1 + 1          # still the graph path (a larger batch re-creates the engine and its counter)
    assert not torch.equal(before[1], after[1]) and all(torch.equal(a, c) for a, c in zip(eager(b["rgb"], 1), after))
