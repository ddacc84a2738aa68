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
``lmdb`` module itself is only needed by :func:`open_shards`.  Not covered: the imgaug augmenters (``augmenter.py``) and
the rotation / translation jitter of the bird's-eye dataset (``birdview_lmdb.py:102-122``).

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
