"""The data collector's on-disk records and the samples the image-agent dataset builds from them.

Record format (``data_collector.py:234-252``): one LMDB environment per episode with the keys ``len`` (ASCII integer),
``rgb_%04d`` uint8[160,384,3], ``birdview_%04d`` uint8[320,320,7], ``measurements_%04d`` float32[17]
(x, y, z, ori_x, ori_y, vx, vy, vz, ax, ay, az, cmd, steer, throttle, brake, manual, gear), ``control_%04d`` float32[3].

Sample (``bird_view/utils/datasets/image_lmdb.py:128-214``): frame ``i`` -> (rgb, bird's-eye crop [192,192,7] ahead of the
car, the car's next ``n_step`` positions ``gap`` frames apart in crop pixels, command, speed).  What is different here:
**the frames stay uint8 in the layout they have on disk** (``[H,W,C]``); the engine's stem writer takes them as they are
(``lbc_net_forward_u8``: ``/255``, normalisation and the bf16 cast happen on the device), so a batch costs 47 MB of
host->device traffic instead of the 189 MB of ``ToTensor`` floats.  ``frames='float'`` gives the reference's tensors.

Anything with ``get(key: bytes) -> bytes`` works as a shard (an ``lmdb`` read transaction; a dict in the tests).  The
``lmdb`` module itself is only needed by :func:`open_shards`.  :class:`BirdViewRecords` is the privileged agent's dataset
(``birdview_lmdb.py:32-170``) with its rotation / translation jitter; the rotation + crop is one bilinear gather over the crop's
pixels (:func:`rotate_crop`), device-agnostic, so a batch of maps can be jittered on the GPU.  Not covered: the imgaug
augmenters (``augmenter.py``).

parity: unpinned -- the reference dataset needs lmdb, cv2 and imgaug, none of which exist in the build container; the test
(tests/test_records.py) checks the geometry against its closed form instead.
"""
import glob
import math
import os

import numpy as np
import torch

PIXELS_PER_METER = 5            # image_lmdb.py:17
PIXEL_OFFSET = 10               # image_lmdb.py:16
RGB_SHAPE = (160, 384, 3)
MAP_SHAPE = (320, 320, 7)
N_MEASUREMENTS = 17


def world_to_pixel(x, y, ox, oy, ori_ox, ori_oy, offset=(-80, 160), size=320):
    """image_lmdb.py:20-28: world position -> pixel of the 320x320 ego-centred map (car at (160, 260), heading up)."""
    dx, dy = (x - ox) * PIXELS_PER_METER, (y - oy) * PIXELS_PER_METER
    px = dx * ori_ox + dy * ori_oy
    py = -dx * ori_oy + dy * ori_ox
    return np.array([size - px + offset[0], py + offset[1]], dtype=np.float64)


class Shard:
    """One episode: ``get`` is the read transaction's (or any mapping's) lookup."""

    def __init__(self, get, name="<memory>"):
        self.get = get
        self.name = name
        raw = get(b"len")
        if raw is None:
            raise KeyError("%s: no 'len' record" % name)
        self.frames = int(raw)

    def _array(self, key, dtype, count):
        raw = self.get(key.encode())
        if raw is None:
            raise KeyError("%s: missing record %s" % (self.name, key))
        a = np.frombuffer(raw, dtype)
        if a.size != count:
            raise ValueError("%s: record %s has %d elements, expected %d" % (self.name, key, a.size, count))
        return a

    def rgb(self, i):
        return self._array("rgb_%04d" % i, np.uint8, 160 * 384 * 3).reshape(RGB_SHAPE)

    def birdview(self, i):
        return self._array("birdview_%04d" % i, np.uint8, 320 * 320 * 7).reshape(MAP_SHAPE)

    def measurements(self, i):
        return self._array("measurements_%04d" % i, np.float32, N_MEASUREMENTS)


def image_sample(shard, index, gap=5, n_step=5, crop_size=192, img_size=320):
    """image_lmdb.py:128-186 without augmentation (delta_angle = 0, dx = 0, dy = -PIXEL_OFFSET): uint8 views of the records."""
    m = shard.measurements(index)
    ox, oy = float(m[0]), float(m[1])
    angle = math.atan2(float(m[4]), float(m[3]))
    ori_ox, ori_oy = math.cos(angle), math.sin(angle)
    speed = float(np.linalg.norm(m[5:8].astype(np.float64)))
    # crop: the window centred at column 160, row 260 - crop/2, moved up by the pixel offset (image_lmdb.py:160-165)
    dx, dy = 0, -PIXEL_OFFSET
    cx, cy = 160, 260 - crop_size // 2
    h = crop_size // 2
    crop = shard.birdview(index)[dy + cy - h:dy + cy + h, dx + cx - h:dx + cx + h]
    locations = np.empty((n_step, 2), np.float32)
    for k, dt in enumerate(range(gap, gap * (n_step + 1), gap)):
        f = shard.measurements(index + dt)
        pixel_y, pixel_x = world_to_pixel(float(f[0]), float(f[1]), ox, oy, ori_ox, ori_oy, size=img_size)   # (sic: image_lmdb.py:178)
        pixel_x = pixel_x - (img_size - crop_size) // 2
        pixel_y = crop_size - (img_size - pixel_y) + 70
        locations[k] = (pixel_x - dx, pixel_y - dy)
    return shard.rgb(index), crop, locations, float(m[11]), speed


class ImageRecords(torch.utils.data.Dataset):
    """``ImageDataset`` of the reference over a list of shards.  Item: ``(rgb, birdview, locations, cmd, speed)``;
    ``frames='uint8'`` (default): rgb uint8 [160,384,3], birdview uint8 [192,192,7] -- the engine's uint8 entry takes both;
    ``frames='float'``: float32 [3,160,384] / [7,192,192] in [0, 1] like ``transforms.ToTensor``.
    ``batch_aug > 1`` stacks that many copies of the frame on a leading axis (the reference stacks augmented variants)."""

    def __init__(self, shards, gap=5, n_step=5, crop_size=192, img_size=320, frames="uint8", batch_aug=1):
        if frames not in ("uint8", "float"):
            raise ValueError("frames must be 'uint8' or 'float'")
        self.shards = [s if isinstance(s, Shard) else Shard(s.get if hasattr(s, "get") else s) for s in shards]
        self.gap, self.n_step, self.crop_size, self.img_size = gap, n_step, crop_size, img_size
        self.frames, self.batch_aug = frames, batch_aug
        self.index = []                       # (shard, frame): the last gap * n_step frames of an episode have no future
        for s in self.shards:
            self.index.extend((s, i) for i in range(max(0, s.frames - gap * n_step)))

    def __len__(self):
        return len(self.index)

    def __getitem__(self, idx):
        shard, i = self.index[idx]
        rgb, crop, loc, cmd, speed = image_sample(shard, i, self.gap, self.n_step, self.crop_size, self.img_size)
        rgb, crop = torch.from_numpy(np.array(rgb)), torch.from_numpy(np.array(crop))   # (writable copies of the record bytes)
        if self.frames == "float":
            rgb = rgb.permute(2, 0, 1).float().div(255)
            crop = crop.permute(2, 0, 1).float().div(255)
        if self.batch_aug > 1:
            rgb = torch.stack([rgb] * self.batch_aug)
        return rgb, crop, torch.from_numpy(loc), torch.tensor(cmd, dtype=torch.float32), torch.tensor(speed, dtype=torch.float32)


def rotate_crop(bev, delta_angle, dx, dy, crop_size=192):
    """birdview_lmdb.py:110-122 as ONE gather: ``cv2.warpAffine(bev, cv2.getRotationMatrix2D((160, 260), delta_angle, 1.0),
    flags=INTER_LINEAR)`` followed by the crop ``[dy + 164 - 96 : dy + 164 + 96, dx + 160 - 96 : dx + 160 + 96]`` -- only the
    crop_size^2 output pixels are interpolated, on whatever device ``bev`` lives (a batch of maps is rotated and cropped on the
    GPU with per-sample jitter).  ``bev``: uint8 [..., 320, 320, C]; ``delta_angle`` (degrees, counter-clockwise on screen as
    in OpenCV), ``dx``, ``dy``: numbers or tensors broadcastable to the batch.  Outside the map: 0 (BORDER_CONSTANT)."""
    t = bev if torch.is_tensor(bev) else torch.from_numpy(np.ascontiguousarray(bev))
    single = t.dim() == 3
    if single:
        t = t[None]
    B, H, W, C = t.shape
    dev = t.device

    def per_sample(v):
        return torch.as_tensor(v, dtype=torch.float64, device=dev).reshape(-1).expand(B)

    ang = per_sample(delta_angle) * (math.pi / 180.0)
    dxs, dys = per_sample(dx), per_sample(dy)
    h = crop_size // 2
    cx, cy = 160.0, 260.0                                           # the rotation centre: the car (birdview_lmdb.py:107-108)
    k = torch.arange(crop_size, device=dev, dtype=torch.float64)
    x = dxs[:, None, None] + (160 - h) + k[None, None, :]           # destination column / row of every crop pixel
    y = dys[:, None, None] + (260 - crop_size // 2 - h) + k[None, :, None]
    al, be = torch.cos(ang)[:, None, None], torch.sin(ang)[:, None, None]
    # getRotationMatrix2D: dst = A (src - c) + c with A = [[al, be], [-be, al]]  =>  src = A^T (dst - c) + c
    u = al * (x - cx) - be * (y - cy) + cx
    v = be * (x - cx) + al * (y - cy) + cy
    u0, v0 = torch.floor(u), torch.floor(v)
    fu, fv = (u - u0)[..., None], (v - v0)[..., None]
    flat = t.reshape(B, H * W, C)
    out = torch.zeros(B, crop_size, crop_size, C, dtype=torch.float64, device=dev)
    for du, dv, wgt in ((0, 0, (1 - fu) * (1 - fv)), (1, 0, fu * (1 - fv)), (0, 1, (1 - fu) * fv), (1, 1, fu * fv)):
        uu, vv = (u0 + du).long(), (v0 + dv).long()
        ok = ((uu >= 0) & (uu < W) & (vv >= 0) & (vv < H))[..., None]
        idx = (vv.clamp(0, H - 1) * W + uu.clamp(0, W - 1)).reshape(B, -1, 1).expand(-1, -1, C)
        val = torch.gather(flat, 1, idx).reshape(B, crop_size, crop_size, C).to(torch.float64)
        out += torch.where(ok, val, torch.zeros_like(val)) * wgt
    out = out.round().clamp_(0, 255).to(torch.uint8)
    return out[0] if single else out


def birdview_sample(shard, index, delta_angle=0, dx=0, dy=-PIXEL_OFFSET, gap=5, n_step=5, crop_size=192, img_size=320):
    """birdview_lmdb.py:90-147 with the jitter (delta_angle degrees, dx, dy pixels) given: the rotated / translated uint8 crop,
    the next ``n_step`` positions in ITS pixels (the heading the positions are expressed in is rotated by the same angle),
    command, speed.  ``dy`` already contains the -PIXEL_OFFSET of birdview_lmdb.py:105."""
    m = shard.measurements(index)
    ox, oy = float(m[0]), float(m[1])
    speed = float(np.linalg.norm(m[5:8].astype(np.float64)))
    crop = rotate_crop(shard.birdview(index), delta_angle, dx, dy, crop_size)
    angle = math.atan2(float(m[4]), float(m[3])) + math.radians(delta_angle)
    ori_ox, ori_oy = math.cos(angle), math.sin(angle)
    locations = np.empty((n_step, 2), np.float32)
    for k, dt in enumerate(range(gap, gap * (n_step + 1), gap)):
        f = shard.measurements(index + dt)
        pixel_y, pixel_x = world_to_pixel(float(f[0]), float(f[1]), ox, oy, ori_ox, ori_oy, size=img_size)   # (sic: :132)
        pixel_x = pixel_x - (img_size - crop_size) // 2
        pixel_y = crop_size - (img_size - pixel_y) + 70
        locations[k] = (pixel_x - dx, pixel_y - dy)
    return crop, locations, float(m[11]), speed


class BirdViewRecords(torch.utils.data.Dataset):
    """``BirdViewDataset`` of the reference (birdview_lmdb.py:32-170) over a list of shards: item ``(bird_view, locations, cmd,
    speed)`` with the reference's random jitter -- rotation by an integer angle in [-angle_jitter, angle_jitter] degrees about
    the car, crop shifted by dx in [-crop_x_jitter, crop_x_jitter] and dy in [0, crop_y_jitter] - 10 pixels.  ``frames='uint8'``
    keeps the crop as uint8 [192,192,7] (the engine's uint8 entry), ``'float'`` gives ToTensor's float32 [7,192,192].
    ``rng``: a ``numpy.random.Generator`` (default: the global ``np.random`` state, as the reference)."""

    def __init__(self, shards, gap=5, n_step=5, crop_size=192, img_size=320, crop_x_jitter=5, crop_y_jitter=5, angle_jitter=5,
                 frames="uint8", max_frames=None, rng=None):
        if frames not in ("uint8", "float"):
            raise ValueError("frames must be 'uint8' or 'float'")
        self.shards = [s if isinstance(s, Shard) else Shard(s.get if hasattr(s, "get") else s) for s in shards]
        self.gap, self.n_step, self.crop_size, self.img_size = gap, n_step, crop_size, img_size
        self.jitter = (int(angle_jitter), int(crop_x_jitter), int(crop_y_jitter))
        self.frames, self.rng = frames, rng
        self.index = []
        for s in self.shards:
            self.index.extend((s, i) for i in range(max(0, s.frames - gap * n_step)))
        if max_frames:
            self.index = self.index[:max_frames]

    def __len__(self):
        return len(self.index)

    def draw_jitter(self):
        """(delta_angle, dx, dy) in the order birdview_lmdb.py:103-105 draws them"""
        aj, xj, yj = self.jitter
        r = (lambda lo, hi: int(self.rng.integers(lo, hi))) if self.rng is not None else (lambda lo, hi: int(np.random.randint(lo, hi)))
        return r(-aj, aj + 1), r(-xj, xj + 1), r(0, yj + 1) - PIXEL_OFFSET

    def __getitem__(self, idx):
        shard, i = self.index[idx]
        delta, dx, dy = self.draw_jitter()
        crop, loc, cmd, speed = birdview_sample(shard, i, delta, dx, dy, self.gap, self.n_step, self.crop_size, self.img_size)
        if self.frames == "float":
            crop = crop.permute(2, 0, 1).float().div(255)
        return crop, torch.from_numpy(loc), torch.tensor(cmd, dtype=torch.float32), torch.tensor(speed, dtype=torch.float32)


class BiasedBirdViewRecords(BirdViewRecords):
    """birdview_lmdb.py:172-205: every item is drawn from the frames of a command chosen with the given ratios (left, right,
    straight, rest = follow); turning frames slower than 1 m/s count as 'follow'."""

    def __init__(self, shards, left_ratio=0.25, right_ratio=0.25, straight_ratio=0.25, **kwargs):
        super().__init__(shards, **kwargs)
        self.weights = [left_ratio, right_ratio, straight_ratio, 1 - left_ratio - right_ratio - straight_ratio]
        self.by_cmd = {c: [] for c in (1, 2, 3, 4)}
        for idx, (shard, i) in enumerate(self.index):
            m = shard.measurements(i)
            cmd, speed = int(m[11]), float(np.linalg.norm(m[5:8].astype(np.float64)))
            self.by_cmd[cmd if (cmd != 4 and speed > 1.0 and cmd in self.by_cmd) else 4].append(idx)

    def __getitem__(self, idx):
        draw = self.rng if self.rng is not None else np.random
        cmd = int(draw.choice([1, 2, 3, 4], p=self.weights))
        pool = self.by_cmd[cmd]
        if not pool:
            raise IndexError("BiasedBirdViewRecords: no frame with command %d in the dataset" % cmd)
        j = int(draw.integers(len(pool))) if self.rng is not None else int(np.random.randint(len(pool)))
        return super().__getitem__(pool[j])


class RandomEpoch(torch.utils.data.Dataset):
    """image_lmdb.py:246-256 (``Wrap``): an 'epoch' is batch_size * samples uniformly random draws, whatever the dataset size."""

    def __init__(self, data, batch_size, samples):
        self.data, self.n = data, batch_size * samples

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.data[np.random.randint(len(self.data))]


def open_shards(dataset_dir):
    """One read transaction per LMDB environment under ``dataset_dir`` (image_lmdb.py:100-112)."""
    try:
        import lmdb
    except ImportError as e:
        raise ImportError("reading LMDB episodes needs the `lmdb` module (not part of this image); "
                          "any object with get(bytes) -> bytes can be passed to ImageRecords instead") from e
    shards = []
    for path in sorted(glob.glob(os.path.join(str(dataset_dir), "**"))):
        env = lmdb.open(path, max_readers=1, readonly=True, lock=False, readahead=False, meminit=False)
        shards.append(Shard(env.begin(write=False).get, name=path))
    return shards


def get_image(dataset_dir, batch_size=32, num_workers=0, shuffle=True, augment=None, n_step=5, gap=5, batch_aug=1,
              frames="uint8"):
    """image_lmdb.py:265-290: (train, val) loaders over ``dataset_dir/train`` and ``dataset_dir/val``; 1000 / 10 random
    batches per epoch; pinned, drop_last.  ``augment`` (imgaug strategies) is not available here."""
    if augment not in (None, "None"):
        raise NotImplementedError("imgaug augmentation strategies are outside this package (augmenter.py)")

    def make(split, is_train):
        data = ImageRecords(open_shards(os.path.join(str(dataset_dir), split)), gap=gap, n_step=n_step, frames=frames,
                            batch_aug=batch_aug if is_train else 1)
        data = RandomEpoch(data, batch_size, 1000 if is_train else 10)
        return torch.utils.data.DataLoader(data, batch_size=batch_size, num_workers=num_workers if is_train else 0,
                                           shuffle=True, drop_last=True, pin_memory=torch.cuda.is_available())
    return make("train", True), make("val", False)


def get_birdview(dataset_dir, batch_size=32, num_workers=0, shuffle=True, crop_x_jitter=0, crop_y_jitter=0, angle_jitter=0,
                 n_step=5, gap=5, max_frames=None, cmd_biased=False, frames="uint8"):
    """birdview_lmdb.py:251-284: (train, val) loaders over ``dataset_dir/train`` and ``dataset_dir/val``; jitter, ``max_frames``
    and the command-biased sampling apply to the training split only; 1000 / 10 random batches per epoch."""
    def make(split, is_train):
        cls = BiasedBirdViewRecords if (is_train and cmd_biased) else BirdViewRecords
        data = cls(open_shards(os.path.join(str(dataset_dir), split)), gap=gap, n_step=n_step,
                   crop_x_jitter=crop_x_jitter if is_train else 0, crop_y_jitter=crop_y_jitter if is_train else 0,
                   angle_jitter=angle_jitter if is_train else 0, max_frames=max_frames if is_train else None, frames=frames)
        data = RandomEpoch(data, batch_size, 1000 if is_train else 10)
        return torch.utils.data.DataLoader(data, batch_size=batch_size, num_workers=num_workers if is_train else 0,
                                           shuffle=True, drop_last=True, pin_memory=torch.cuda.is_available())
    return make("train", True), make("val", False)
