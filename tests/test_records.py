"""records.py: the collector's record format -> training samples (image_lmdb.py:128-214), over an in-memory shard."""
import math

import numpy as np
import pytest
import torch

from learningbycheating_b200 import records


def _episode(n, seed=0):
    g = np.random.default_rng(seed)
    d = {b"len": str(n).encode()}
    heading = 0.7
    pos = np.cumsum(np.stack([np.cos(heading + 0.02 * np.arange(n)), np.sin(heading + 0.02 * np.arange(n))], 1) * 0.6, 0) + (31.0, -12.0)
    for i in range(n):
        d[b"rgb_%04d" % i] = g.integers(0, 256, records.RGB_SHAPE, dtype=np.uint8).tobytes()
        d[b"birdview_%04d" % i] = (g.random(records.MAP_SHAPE) > 0.8).astype(np.uint8).__mul__(255).tobytes()
        m = np.zeros(17, np.float32)
        m[0:2] = pos[i]
        th = heading + 0.02 * i
        m[3:5] = (2.5 * math.cos(th), 2.5 * math.sin(th))          # un-normalised orientation vector, as the collector may store
        m[5:8] = (3.0, 4.0, 12.0)
        m[11] = 1 + i % 4
        d[b"measurements_%04d" % i] = m.tobytes()
    return d, pos


def test_record_samples_cpu(monkeypatch):
    import sys
    monkeypatch.setitem(sys.modules, "lmdb", None)      # the oracle's reference importer may have left a mock `lmdb` behind
    d, pos = _episode(40)
    data = records.ImageRecords([d], gap=5, n_step=5)
    assert len(data) == 40 - 25
    rgb, bev, loc, cmd, speed = data[3]
    assert rgb.dtype == torch.uint8 and tuple(rgb.shape) == (160, 384, 3)
    assert torch.equal(rgb, torch.from_numpy(np.frombuffer(d[b"rgb_0003"], np.uint8).reshape(160, 384, 3)))
    full = np.frombuffer(d[b"birdview_0003"], np.uint8).reshape(320, 320, 7)
    assert bev.dtype == torch.uint8 and torch.equal(bev, torch.from_numpy(full[58:250, 64:256].copy()))
    assert float(cmd) == 4.0 and abs(float(speed) - 13.0) < 1e-6
    # closed form of the location transform: (96 + lateral, 192 - forward) in crop pixels, 5 px per metre
    th = 0.7 + 0.02 * 3
    for k in range(5):
        dxy = pos[3 + 5 * (k + 1)] - pos[3]
        fwd = 5 * (dxy[0] * math.cos(th) + dxy[1] * math.sin(th))
        lat = 5 * (-dxy[0] * math.sin(th) + dxy[1] * math.cos(th))
        assert abs(float(loc[k, 0]) - (96 + lat)) < 2e-4 and abs(float(loc[k, 1]) - (192 - fwd)) < 2e-4
    assert float(loc[0, 1]) < 192 and float(loc[4, 1]) < float(loc[0, 1])          # the car drives "up" the crop
    # float frames == ToTensor; batch_aug stacks
    rgbf, bevf, loc2, _, _ = records.ImageRecords([d], frames="float", batch_aug=2)[3]
    assert tuple(rgbf.shape) == (2, 3, 160, 384) and torch.equal(rgbf[0], rgb.permute(2, 0, 1).float() / 255)
    assert tuple(bevf.shape) == (7, 192, 192) and set(bevf.unique().tolist()) <= {0.0, 1.0} and torch.equal(loc, loc2)
    # a default collate gives the batch the training loop takes
    batch = torch.utils.data.default_collate([data[0], data[1]])
    assert tuple(batch[0].shape) == (2, 160, 384, 3) and tuple(batch[2].shape) == (2, 5, 2) and batch[3].dtype == torch.float32
    # random-epoch wrapper and error paths
    ep = records.RandomEpoch(data, 4, 10)
    assert len(ep) == 40 and tuple(ep[0][0].shape) == (160, 384, 3)
    with pytest.raises(KeyError):
        records.ImageRecords([{b"rgb_0000": b""}])
    bad = dict(d)
    bad[b"measurements_0001"] = b"\0" * 8
    with pytest.raises(ValueError):
        records.ImageRecords([bad])[1]
    with pytest.raises(ImportError, match="lmdb"):
        records.open_shards("/nonexistent")


def test_uint8_records_feed_the_engine_cpu(backend):
    """uint8 [B,H,W,C] frames and map crops, as they sit in the records, give the same step as the ToTensor floats."""
    from lbc_testing import build_models
    import learningbycheating_b200 as lbc
    d, _ = _episode(30, seed=3)
    u8 = torch.utils.data.default_collate([records.ImageRecords([d])[i] for i in (0, 2)])
    fl = torch.utils.data.default_collate([records.ImageRecords([d], frames="float")[i] for i in (0, 2)])
    s, t = build_models(backend, "fp32")
    s.eval()
    t.eval()
    oh = lbc.one_hot(u8[3]).to(backend)
    with torch.no_grad():
        for net, a, b in ((s, u8[0], fl[0]), (t, u8[1], fl[1])):
            pa = net(a.to(backend), u8[4].to(backend), oh)[0]
            pb = net(b.to(backend), fl[4].to(backend), oh)[0]
            assert (pa - pb).abs().max() < 1e-5


# ------------------------------------------------------------------ bird's-eye dataset with jitter (birdview_lmdb.py:90-147)
def _warp_affine_reference(img, angle_deg, centre=(160.0, 260.0)):
    """cv2.warpAffine(img, cv2.getRotationMatrix2D(centre, angle, 1.0), flags=INTER_LINEAR), restated from OpenCV's documented
    matrix M = [[a, b, (1-a)cx - b cy], [-b, a, b cx + (1-a)cy]] (a = cos, b = sin): dst(x, y) = src(M^-1 (x, y, 1)), bilinear,
    zeros outside.  Independent of records.rotate_crop (general 3x3 inverse, per-pixel loops vectorised with numpy)."""
    a, b = math.cos(math.radians(angle_deg)), math.sin(math.radians(angle_deg))
    cx, cy = centre
    M = np.array([[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy], [0, 0, 1]], np.float64)
    Mi = np.linalg.inv(M)
    H, W, C = img.shape
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    u = Mi[0, 0] * xs + Mi[0, 1] * ys + Mi[0, 2]
    v = Mi[1, 0] * xs + Mi[1, 1] * ys + Mi[1, 2]
    u0, v0 = np.floor(u).astype(int), np.floor(v).astype(int)
    fu, fv = (u - u0)[..., None], (v - v0)[..., None]
    out = np.zeros((H, W, C), np.float64)
    for du, dv, w in ((0, 0, (1 - fu) * (1 - fv)), (1, 0, fu * (1 - fv)), (0, 1, (1 - fu) * fv), (1, 1, fu * fv)):
        uu, vv = u0 + du, v0 + dv
        ok = (uu >= 0) & (uu < W) & (vv >= 0) & (vv < H)
        val = img[np.clip(vv, 0, H - 1), np.clip(uu, 0, W - 1)].astype(np.float64)
        out += np.where(ok[..., None], val, 0.0) * w
    return out


def test_rotate_crop_matches_warp_affine_then_crop_cpu():
    g = np.random.default_rng(5)
    img = g.integers(0, 256, records.MAP_SHAPE, dtype=np.uint8)
    # no jitter: exactly the image dataset's crop
    assert torch.equal(records.rotate_crop(img, 0, 0, -10), torch.from_numpy(img[58:250, 64:256].copy()))
    for angle, dx, dy in ((5, 3, -7), (-4, -5, -10), (90, 0, -10), (17, 5, -5)):
        full = _warp_affine_reference(img, angle)
        ref = full[dy + 164 - 96:dy + 164 + 96, dx + 160 - 96:dx + 160 + 96]
        got = records.rotate_crop(img, angle, dx, dy).numpy().astype(np.float64)
        assert np.abs(got - ref).max() <= 0.5 + 1e-6, (angle, dx, dy)      # rounding of the same bilinear value
    # a batch with per-sample jitter == the samples one by one
    batch = torch.from_numpy(np.stack([img, img[::-1].copy()]))
    out = records.rotate_crop(batch, torch.tensor([5.0, -3.0]), torch.tensor([2, -4]), torch.tensor([-10, -6]))
    assert torch.equal(out[0], records.rotate_crop(batch[0], 5, 2, -10)) and torch.equal(out[1], records.rotate_crop(batch[1], -3, -4, -6))


def test_birdview_records_cpu():
    d, pos = _episode(40)
    img_data = records.ImageRecords([d], gap=5, n_step=5)
    # jitter switched off: the bird's-eye dataset is the image dataset's crop / locations
    data0 = records.BirdViewRecords([d], crop_x_jitter=0, crop_y_jitter=0, angle_jitter=0)
    assert len(data0) == len(img_data) == 15
    bev, loc, cmd, speed = data0[3]
    _, bev_i, loc_i, cmd_i, speed_i = img_data[3]
    assert torch.equal(bev, bev_i) and torch.allclose(loc, loc_i) and float(cmd) == float(cmd_i) and float(speed) == float(speed_i)
    # the jitter is drawn as the reference draws it (angle, dx, dy - 10) and reproducible with a Generator
    a = records.BirdViewRecords([d], rng=np.random.default_rng(3))
    b = records.BirdViewRecords([d], rng=np.random.default_rng(3))
    ja = [a.draw_jitter() for _ in range(50)]
    assert ja == [b.draw_jitter() for _ in range(50)]
    assert all(-5 <= j[0] <= 5 and -5 <= j[1] <= 5 and -10 <= j[2] <= -5 for j in ja) and len(set(ja)) > 20
    x, l, _, _ = records.BirdViewRecords([d], frames="float", rng=np.random.default_rng(1))[2]
    assert tuple(x.shape) == (7, 192, 192) and x.dtype == torch.float32 and tuple(l.shape) == (5, 2)
    # image and locations move together: a blob painted into the map where waypoint k sits (un-jittered map pixel) is found at
    # the returned location after rotation + shift.  (The reference rotates the map about the car at row 260 and the positions
    # about the measurement origin 10 px ahead of it, so the two agree to ~1 px at its +-5 degrees, not exactly.)
    th = 0.7 + 0.02 * 3
    for delta, dx, dy in ((5, 4, -6), (-5, -3, -10), (3, 0, -8)):
        for k in (0, 2, 4):
            dxy = pos[3 + 5 * (k + 1)] - pos[3]
            fwd = 5 * (dxy[0] * math.cos(th) + dxy[1] * math.sin(th))
            lat = 5 * (-dxy[0] * math.sin(th) + dxy[1] * math.cos(th))
            col, row = 160 + lat, 250 - fwd          # full-map pixel of the waypoint (crop (96 + lat, 192 - fwd) + crop origin (64, 58))
            m = np.zeros(records.MAP_SHAPE, np.uint8)
            yy, xx = np.mgrid[0:320, 0:320]
            m[..., 0] = (255 * np.exp(-((xx - col) ** 2 + (yy - row) ** 2) / 8.0)).astype(np.uint8)
            dd = dict(d)
            dd[b"birdview_0003"] = m.tobytes()
            crop, loc, _, _ = records.birdview_sample(records.Shard(dd.get), 3, delta, dx, dy)
            w = crop[..., 0].double()
            cy = float((w.sum(1) * torch.arange(192)).sum() / w.sum())
            cx = float((w.sum(0) * torch.arange(192)).sum() / w.sum())
            assert abs(cx - float(loc[k, 0])) < 1.5 and abs(cy - float(loc[k, 1])) < 1.5, (delta, dx, dy, k, cx, cy, loc[k])


def test_biased_birdview_records_cpu():
    d, _ = _episode(60)           # commands cycle 1..4, speed 13 m/s
    data = records.BiasedBirdViewRecords([d], left_ratio=0.5, right_ratio=0.5, straight_ratio=0.0, angle_jitter=0,
                                         crop_x_jitter=0, crop_y_jitter=0, rng=np.random.default_rng(2))
    assert sorted(len(v) for v in data.by_cmd.values()) == [8, 9, 9, 9] and sum(len(v) for v in data.by_cmd.values()) == len(data) == 35
    cmds = [float(data[0][2]) for _ in range(40)]
    assert set(cmds) == {1.0, 2.0}                       # only the two commands with non-zero ratio are ever drawn
    with pytest.raises(ImportError, match="lmdb"):
        import sys
        saved = sys.modules.get("lmdb", "absent")
        sys.modules["lmdb"] = None
        try:
            records.get_birdview("/nonexistent")
        finally:
            if saved == "absent":
                sys.modules.pop("lmdb", None)
            else:
                sys.modules["lmdb"] = saved
