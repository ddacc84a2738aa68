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
