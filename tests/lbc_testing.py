"""Shared checkers for the parity tests (engine vs golden vectors dumped from the reference)."""
import os

import numpy as np
import torch

import lbc_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def build_models(device, precision, all_branch=True, teacher_all_branch=True):
    """SURVEY 8(c) protocol: seed 0, student first, then teacher."""
    import learningbycheating_b200 as lbc
    torch.manual_seed(0)
    s = lbc.ImagePolicyModelSS('resnet34', all_branch=all_branch, lbc_precision=precision)
    t = lbc.BirdViewPolicyModelSS('resnet18', all_branch=teacher_all_branch, lbc_precision=precision)
    return s.to(device), t.to(device)


def batch_on(device, B, seed=1):
    b = orc.synthetic_batch(B, seed)
    return {k: v.to(device) for k, v in b.items()}


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def check_init(model, tag):
    init = gold("init_seed0.npz")
    sd = model.state_dict()
    assert list(sd.keys()) == list(init[tag + "/keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(init[tag + "/shapes"])
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            a = v.double().cpu().numpy().reshape(-1)
            assert abs(a.sum() - init["%s/%s/sum" % (tag, k)]) <= 1e-9 * max(1.0, np.abs(a).sum()), k
            np.testing.assert_allclose(a[:8], init["%s/%s/head" % (tag, k)], rtol=0, atol=1e-12)


def run_student_steps(device, precision, B, phase, n_steps, optimizer="lbc"):
    """Drives the package's own train loop pieces exactly like train_or_eval; returns per-step records."""
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_image_phase0 as p0, train_image_phase1 as p1
    s, t = build_models(device, precision)
    s.train()
    t.eval()
    b = batch_on(device, B)
    opt = lbc.Adam(s.parameters(), lr=1e-4) if optimizer == "lbc" else torch.optim.Adam(s.parameters(), lr=1e-4)
    if phase == 0:
        conv, crit = p0.CoordConverter(device=device), p0.LocationLoss(device=device)
    else:
        conv, crit = p1.CoordConverter(fixed_offset=4.0, device=device), p1.LocationLoss()
    recs = []
    for step in range(n_steps):
        oh = lbc.one_hot(b["command"].cpu()).to(device)
        with torch.no_grad():
            tp, tps = t(b["birdview"], b["speed"], oh)
        p, ps = s(b["rgb"], b["speed"], oh)
        if phase == 0:
            target = conv(tp)
            loss = crit(p, target)
        else:
            target = None
            loss = crit(conv(ps), tps)
        lm = loss.mean()
        opt.zero_grad()
        lm.backward()
        grads = {k: (q.grad.detach().clone().cpu() if q.grad is not None else None) for k, q in s.named_parameters()}
        opt.step()
        recs.append(dict(pred=p.detach().cpu(), preds=ps.detach().cpu(), t_pred=tp.cpu(), t_preds=tps.cpu(),
                         loss=loss.detach().cpu(), loss_mean=float(lm), grads=grads,
                         target=None if target is None else target.cpu(),
                         post={k: v.detach().clone().cpu() for k, v in s.state_dict().items()}))
    return s, t, recs


def check_step0_against_golden(rec, g, out_tol, loss_tol, grad_global_tol, grad_tensor_tol):
    assert rel_err(rec["t_pred"], g["step0/t_pred"]) <= out_tol
    assert rel_err(rec["t_preds"], g["step0/t_preds"]) <= out_tol
    assert rel_err(rec["pred"], g["step0/pred"]) <= out_tol, rel_err(rec["pred"], g["step0/pred"])
    assert rel_err(rec["preds"], g["step0/preds"]) <= out_tol
    assert rel_err(rec["loss"], g["step0/loss"]) <= loss_tol
    assert abs(rec["loss_mean"] - float(g["step0/loss_mean"])) <= loss_tol * abs(float(g["step0/loss_mean"]))
    if rec["target"] is not None:
        assert rel_err(rec["target"], g["step0/target_px"]) <= 1e-6
    gl2 = 0.0
    worst = (0.0, None)
    for k, gr in rec["grads"].items():
        key = "step0/grad/%s/l2" % k
        if gr is None:
            assert key not in g.files, "%s: engine produced no gradient" % k    # conv.fc.* only
            continue
        assert key in g.files, "%s: reference has grad None" % k
        l2 = float(gr.double().norm())
        gl2 += l2 * l2
        ref = float(g[key])
        if k.startswith("location_pred.") and (k.endswith(".0.bias") or k.endswith(".1.bias")):
            # head BN beta / 1x1-conv bias: mathematically zero gradient (softmax shift invariance);
            # both sides hold rounding noise only -> absolute bound relative to the global gradient norm
            assert l2 <= 1e-5 * max(1.0, float(g["step0/grad_global_l2"])), (k, l2)
            continue
        if ref < 1e-12:         # branch never selected in this batch (phase 0): exactly zero in the reference
            assert l2 < 1e-6, (k, l2)
            continue
        full = "step0/grad/%s/full" % k
        if full in g.files:
            e = float((gr.double() - torch.from_numpy(g[full]).double()).norm()) / ref
        else:
            e = abs(l2 - ref) / ref
        if e > worst[0]:
            worst = (e, k)
    assert abs(gl2 ** 0.5 - float(g["step0/grad_global_l2"])) <= grad_global_tol * float(g["step0/grad_global_l2"])
    assert worst[0] <= grad_tensor_tol, worst
    return worst
