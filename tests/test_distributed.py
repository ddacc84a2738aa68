"""Data-parallel plumbing (new functionality): world_size-2 `gloo` run on the CPU (host-emulation engine):
flat-gradient all-reduce + Adam with the 1/world scale == averaging the two per-rank gradients by hand."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, overlap=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    os.environ["OMP_NUM_THREADS"] = "4"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learningbycheating_b200 import _lib, build
    _lib.use_library_for_tests(build.EMU_LIB)
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_image_phase0 as p0
    from learningbycheating_b200.distributed import DataParallel
    from lbc_testing import batch_on
    torch.manual_seed(0)
    net = lbc.ImagePolicyModelSS("resnet34", all_branch=True, lbc_precision="fp32").train()
    if rank == 1:   # start from different weights on purpose: sync_initial_state must broadcast rank 0's
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.01)
    opt = lbc.Adam(net.parameters(), lr=1e-4)
    b = batch_on("cpu", 2, seed=1 + rank)
    oh = lbc.one_hot(b["command"])
    tgt = torch.rand(2, 5, 2, generator=torch.Generator().manual_seed(7 + rank)) * torch.tensor([384.0, 160.0])
    dp = DataParallel(net, opt, overlap=overlap)
    net.lbc_flat_state(2)
    dp.sync_initial_state()
    pred, _ = net(b["rgb"], b["speed"], oh)
    loss = p0.LocationLoss(device="cpu")(pred, tgt).mean()
    opt.zero_grad()
    if rank == 0:       # the bucket table: contiguous ranges in backward order covering exactly the trained parameters
        from learningbycheating_b200.distributed import grad_buckets
        bk = grad_buckets(net)
        assert len(bk) == 5 and bk[-1][0] == 0
        covered = torch.zeros(net._lbc.flat_grads.numel(), dtype=torch.bool)
        for off, n in bk:
            assert not covered[off:off + n].any()
            covered[off:off + n] = True
        for p_, view, _, on_path in net._lbc.param_views:
            o = view.storage_offset()
            assert bool(covered[o:o + view.numel()].all()) == on_path, "bucket coverage != on-path parameters"
    # overlap=True: the bucketed path (on the GPU the same buckets ride a side stream under backward); False: one all-reduce
    local = {}
    hook = net._lbc
    import learningbycheating_b200.distributed as D
    orig = D.dist.all_reduce

    def spy(t, *a, **k):
        if "g" not in local:
            local["g"] = hook.flat_grads.clone()      # the local gradient, before the first bucket is summed
        return orig(t, *a, **k)
    D.dist.all_reduce = spy
    p_before = net._lbc.flat_params.clone()
    dp.backward(loss)
    dp.step_after_backward()
    D.dist.all_reduce = orig
    local_grad = local["g"]
    torch.save(dict(local_grad=local_grad, summed=net._lbc.flat_grads.clone(), p_before=p_before,
                    p_after=net._lbc.flat_params.clone(), loss=float(loss)), os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("overlap", [False, True], ids=["one_allreduce", "bucketed"])
def test_dp2_gloo_matches_manual_average(tmp_path, overlap):
    from learningbycheating_b200 import build
    build.build_hostemu()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), overlap), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    # identical starting point after the broadcast, identical result after the step
    assert torch.equal(r0["p_before"], r1["p_before"])
    assert torch.equal(r0["p_after"], r1["p_after"])
    # all-reduce = sum of the two local gradients
    assert torch.allclose(r0["summed"], r0["local_grad"] + r1["local_grad"], rtol=0, atol=1e-7)
    # Adam with grad_scale 1/2 on the summed gradient == torch.optim.Adam on the averaged gradient
    p = r0["p_before"].clone().requires_grad_(True)
    p.grad = 0.5 * (r0["local_grad"] + r1["local_grad"])
    mask = p.grad != 0          # conv.fc.* has no gradient and is skipped by the fused step
    ref = torch.optim.Adam([p], lr=1e-4)
    ref.step()
    d = (p.detach() - r0["p_after"]).abs()
    assert d[mask].max() < 1e-7
