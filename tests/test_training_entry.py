"""train(config) with the reference's signature (training/train_image_phase0.py:214-242): models built from the config,
teacher loaded from its checkpoint, data through the dataset hook, epoch-0 dry run, checkpoint on SAVE_EPOCHS."""
import os

import pytest
import torch

from lbc_testing import batch_on, build_models


def test_train_config_reference_signature_cpu(backend, tmp_path, monkeypatch):
    monkeypatch.setenv("LBC_B200_PRECISION", "fp32")       # models built inside train() take the package default
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import train_image_phase0 as p0
    dev = backend
    _, teacher = build_models(dev, "fp32", teacher_all_branch=False)
    tpath = tmp_path / "teacher" / "model-128.th"
    os.makedirs(tpath.parent)
    torch.save(teacher.state_dict(), tpath)
    b = batch_on("cpu", 2)
    batch = (b["rgb"], b["birdview"], b["location"], b["command"], b["speed"])
    calls = []

    def loader(**data_args):
        calls.append(data_args)
        return [batch], [batch]

    config = dict(log_dir=str(tmp_path / "run"), log_iterations=1000, max_epoch=1, device=torch.device(dev),
                  optimizer_args=dict(lr=1e-4), data_args=dict(dataset_dir="unused", batch_size=2), data_loader=loader,
                  model_args=dict(model="image_ss", imagenet_pretrained=False, backbone="resnet34"),
                  camera_args=dict(w=384, h=160, fov=90, world_y=1.4, fixed_offset=4.0),
                  teacher_args=dict(model_path=str(tpath)))
    torch.manual_seed(0)
    net = p0.train(config)
    assert calls == [dict(dataset_dir="unused", batch_size=2)]
    assert isinstance(net, lbc.ImagePolicyModelSS)
    # epoch 0 is the reference's dry run (no optimizer step), epoch 1 trains once: one BN update per train-mode forward
    assert int(net.conv.bn1.num_batches_tracked) == 2
    ck = tmp_path / "run" / "model-1.th"
    assert ck.exists()
    sd = torch.load(ck)
    assert list(sd.keys()) == list(net.state_dict().keys())
    # without a dataset the entry point says so instead of guessing
    import pytest
    cfg2 = dict(config)
    cfg2.pop("data_loader")
    with pytest.raises(lbc.LbcError, match="no dataset"):
        p0.train(cfg2)


@pytest.mark.gpu
def test_cuda_prefetcher_fixed_slots_gpu():
    """data.CudaPrefetcher: same tuples out as in, on the device; ragged last batch; staging buffers allocated once
    (no allocator growth in the steady state); early exit + re-iteration keeps the slot ordering safe."""
    from learningbycheating_b200.data import CudaPrefetcher
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randint(0, 256, (8 if i < 6 else 5, 3, 16, 24), dtype=torch.uint8, generator=g),
                torch.rand(8 if i < 6 else 5, generator=g), None, i) for i in range(7)]
    pf = CudaPrefetcher(batches, "cuda:0")
    seen = 0
    for i, (a, b, c, k) in enumerate(pf):
        assert a.is_cuda and b.is_cuda and c is None and k == i
        # consume with a kernel that takes a while so that the next copies really overlap
        assert torch.equal(a.cpu(), batches[i][0]) and torch.equal(b.cpu(), batches[i][1])
        seen += 1
    assert seen == 7
    torch.cuda.synchronize()
    m0 = torch.cuda.memory_reserved()
    for rep in range(3):
        for i, (a, b, c, k) in enumerate(pf):
            s = a.float().sum() + b.sum()
            if rep == 1 and i == 2:
                break
        assert torch.isfinite(s)
    # results stay right after an early exit
    for i, (a, b, c, k) in enumerate(pf):
        assert torch.equal(a.cpu(), batches[i][0])
    torch.cuda.synchronize()
    assert torch.cuda.memory_reserved() <= m0 + (4 << 20)


def test_phase1_batch_aug_and_speed_noise_cpu(backend, monkeypatch):
    """a17 (training/train_image_phase1.py:131-154,170-189): a [B, batch_aug, 3, H, W] frame tensor is flattened and the
    per-sample inputs are repeated batch_aug times (numpy.repeat order); in training the speed gets N(0, speed_noise)
    noise clamped to [0, 10] BEFORE it is repeated; evaluation passes the speed through."""
    import numpy as np
    from learningbycheating_b200 import train_image_phase1 as p1
    # repeat == numpy.repeat along dim 0 (the reference's helper is an index_select over arange(n).repeat)
    a = torch.arange(12.).reshape(3, 4)
    assert torch.equal(p1.repeat(a, 2), torch.from_numpy(np.repeat(a.numpy(), 2, axis=0)))
    assert torch.equal(p1.repeat(a, 3, dim=1), torch.from_numpy(np.repeat(a.numpy(), 3, axis=1)))

    dev = backend
    s, t = build_models(dev, "fp32")
    s.eval()
    t.eval()
    b = batch_on(dev, 2)
    B, aug = 1, 2
    rgb5 = b["rgb"].reshape(B, aug, 3, 160, 384)
    data = [(rgb5, b["birdview"][:B], b["location"][:B], b["command"][:B], b["speed"][:B])]
    conv, crit = p1.CoordConverter(fixed_offset=4.0, device=dev), p1.LocationLoss()
    config = dict(device=dev, log_iterations=1000, speed_noise=0.0)
    seen = []
    orig = type(t).forward

    def spy(self, birdview, speed, command):
        seen.append((birdview.clone(), speed.clone(), command.clone()))
        return orig(self, birdview, speed, command)
    monkeypatch.setattr(type(t), "forward", spy)
    got = p1.train_or_eval(conv, crit, s, t, data, None, False, config, False)
    # the same thing by hand on the flattened batch
    import learningbycheating_b200 as lbc
    oh = lbc.one_hot(b["command"][:B].cpu()).to(dev)
    bev2, sp2, oh2 = (torch.repeat_interleave(x, aug, 0) for x in (b["birdview"][:B], b["speed"][:B], oh))
    with torch.no_grad():
        _, tl = orig(t, bev2, sp2, oh2)
        _, pl = s(b["rgb"], sp2, oh2)
        want = crit(conv(pl), tl).mean()
    assert abs(float(got[0]) - float(want)) < 1e-6
    assert seen[0][0].shape[0] == B * aug and torch.equal(seen[0][1], sp2) and torch.equal(seen[0][2], oh2)

    # speed noise: training only, clamped, drawn per SAMPLE (before the repeat)
    seen.clear()
    monkeypatch.setattr(torch, "randn", lambda size, device=None: torch.tensor([100.0], device=device).expand(size))
    config["speed_noise"] = 1.0
    p1.train_or_eval(conv, crit, s, t, data, None, False, config, False)             # eval: untouched
    assert torch.equal(seen[-1][1], sp2)
    s.train()
    p1.train_or_eval(conv, crit, s, t, data, None, True, config, True)               # first-epoch dry run: no optimizer step
    assert torch.equal(seen[-1][1], torch.full_like(sp2, 10.0))
    monkeypatch.setattr(torch, "randn", lambda size, device=None: torch.tensor([-100.0], device=device).expand(size))
    p1.train_or_eval(conv, crit, s, t, data, None, True, config, True)
    assert torch.equal(seen[-1][1], torch.zeros_like(sp2))
