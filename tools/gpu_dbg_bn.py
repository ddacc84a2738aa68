"""debug: lbc_op_bn_bwd bf16 fast vs correctness-first vs torch"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from learningbycheating_b200 import _lib
L = _lib.lib()
for (M, C) in ((60, 640), (1920, 128)):
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(M, C, generator=g) * 1.5 + 0.3).bfloat16().float()
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5
    dy = torch.randn(M, C, generator=g).bfloat16().float()
    xd = x.double(); mean = xd.mean(0); var = xd.var(0, unbiased=False); rstd = 1 / (var + 1e-5).sqrt()
    xh = (xd - mean) * rstd
    dg_ref = (dy.double() * xh).sum(0); db_ref = dy.double().sum(0)
    dg_nomean = (dy.double() * xd * rstd).sum(0)
    for fast in (1, 0):
        L.lbc_set_fast_kernels(fast)
        dg, db, dx = torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(M, C, device="cuda")
        xg, dyg, gg, bg = x.cuda(), dy.cuda(), gamma.cuda(), beta.cuda()
        _lib.trace(True)
        _lib.check(L.lbc_op_bn_bwd(_lib.ptr(dyg), _lib.ptr(xg), _lib.ptr(gg), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(dx), M, C, 1, None, None, 0, None))
        tr = _lib.trace_counts(); _lib.trace(False)
        print("M=%d C=%d fast=%d: dgamma err %.3e (vs no-mean formula %.3e)  dbeta err %.3e   kernels=%s" % (
            M, C, fast, float((dg.cpu().double() - dg_ref).abs().max()), float((dg.cpu().double() - dg_nomean).abs().max()),
            float((db.cpu().double() - db_ref).abs().max()), [k for k in tr if "bn_bwd" in k]))
        print("   first dgamma got", dg[:4].cpu().tolist(), "ref", dg_ref[:4].tolist(), "dbeta got", db[:4].cpu().tolist(), "ref", db_ref[:4].tolist())
    L.lbc_set_fast_kernels(1)
