"""Eval-mode forward latency of the student at small batches (ImageAgent.run_step is a B = 1 forward).
Prints one JSON line per batch size: median / p90 of per-call device time (CUDA events), calls back to back."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import learningbycheating_b200 as lbc  # noqa: E402

dev = "cuda:0"
for precision in ("bf16", "fp32tc"):
    torch.manual_seed(0)
    net = lbc.ImagePolicyModelSS("resnet34", all_branch=False, lbc_precision=precision).to(dev).eval()
    for B, mode in ((1, "eager"), (1, "graph"), (8, "eager"), (8, "graph"), (64, "eager")):
        os.environ["LBC_B200_INFER_GRAPH_MAX_B"] = "16" if mode == "graph" else "0"
        rgb = torch.rand(B, 3, 160, 384, device=dev)
        speed = torch.rand(B, device=dev) * 10
        oh = lbc.one_hot(torch.randint(1, 5, (B,)).float()).to(dev)
        with torch.no_grad():
            for _ in range(5):
                net(rgb, speed, oh)
            torch.cuda.synchronize()
            ts = []
            for _ in range(50):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = net(rgb, speed, oh)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
        print(json.dumps(dict(probe="eval_forward", precision=precision, B=B, mode=mode, ms_median=ts[len(ts) // 2], ms_p90=ts[int(len(ts) * .9)],
                              finite=bool(torch.isfinite(out).all()))))
