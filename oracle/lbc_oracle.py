"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the LearningByCheating image-agent training step.

A plain-PyTorch (CPU, fp32 or fp64) *restatement* of the reference's algorithm, written
functionally over a ``state_dict`` so that it needs neither the reference tree nor any of
its dependencies (carla, lmdb, imgaug ...).  The arithmetic itself lives in third-party
``torch`` (reference pins pytorch=1.0.0, environment.yml:151; oracle = this image's torch
2.11 CPU kernels, SURVEY.md section 8(c) "version drift").

Parity status: PINNED against outputs of the reference itself -- ``oracle/make_golden.py``
imports the unmodified reference from /root/reference (via oracle/ref_import.py), checks
this restatement against it (forward, loss, gradients, Adam step, BN buffers) and commits
the resulting vectors to ``tests/golden/``.  The reference's own tests pin nothing on this
path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this file.  It is the checker, never the product.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

STEPS = 5          # bird_view/models/image.py:15
COMMANDS = 4       # bird_view/models/image.py:16
CROP_SIZE = 192    # bird_view/models/image.py:18
PIXELS_PER_METER = 5  # bird_view/models/image.py:19
BN_EPS = 1e-5      # torch.nn.BatchNorm2d default used at resnet.py:104, image.py:38
BN_MOMENTUM = 0.1

RESNET_LAYERS = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3]}  # resnet.py:163-164
RGB_MEAN = [0.485, 0.456, 0.406]   # image.py:32-35
RGB_STD = [0.229, 0.224, 0.225]


# ----------------------------------------------------------------------------------------
# host-side helpers
# ----------------------------------------------------------------------------------------
def one_hot(x, num_digits=4, start=1):
    """bird_view/utils/train_utils.py:33-40 -- class ids 1..4 -> one-hot, out-of-range ids clamp."""
    idx = torch.clamp(x.long() - start, 0, num_digits - 1)
    y = torch.zeros(x.shape[0], num_digits, dtype=torch.float32)
    y[torch.arange(x.shape[0]), idx] = 1.0
    return y


def spatial_grid(width, height, dtype=torch.float32):
    """common.py:127-134 as called from image.py:52,58 (SpatialSoftmax(ow, oh, STEPS)):
    pos_x = linspace(-1,1,width) varying along W, pos_y = linspace(-1,1,height) along H,
    both flattened row-major over (H, W)."""
    px, py = np.meshgrid(np.linspace(-1.0, 1.0, width), np.linspace(-1.0, 1.0, height))
    return (torch.from_numpy(px.reshape(-1)).float().to(dtype),
            torch.from_numpy(py.reshape(-1)).float().to(dtype))


# ----------------------------------------------------------------------------------------
# network forward (ImagePolicyModelSS / BirdViewPolicyModelSS share the topology)
# ----------------------------------------------------------------------------------------
class _BN:
    """Train/eval BatchNorm2d with explicit buffer side effects."""

    def __init__(self, sd, train, new_buffers):
        self.sd, self.train, self.new = sd, train, new_buffers

    def __call__(self, x, prefix):
        w, b = self.sd[prefix + ".weight"], self.sd[prefix + ".bias"]
        rm, rv = self.sd[prefix + ".running_mean"], self.sd[prefix + ".running_var"]
        if not self.train:
            return F.batch_norm(x, rm.to(x.dtype), rv.to(x.dtype), w, b, False, 0.0, BN_EPS)
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            mean = x.mean(dim=(0, 2, 3))
            var_b = x.var(dim=(0, 2, 3), unbiased=False)
            self.new[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm.to(x.dtype) + BN_MOMENTUM * mean
            self.new[prefix + ".running_var"] = ((1 - BN_MOMENTUM) * rv.to(x.dtype)
                                                 + BN_MOMENTUM * var_b * (n / max(n - 1, 1)))
            self.new[prefix + ".num_batches_tracked"] = self.sd[prefix + ".num_batches_tracked"] + 1
        return F.batch_norm(x, None, None, w, b, True, 0.0, BN_EPS)


def policy_forward(sd, x, velocity, command_onehot, backbone, train, normalize, taps=None):
    """ImagePolicyModelSS.forward (image.py:64-89) when normalize=True,
    BirdViewPolicyModelSS.forward (birdview.py:61-79) when normalize=False.

    sd: dict name->tensor with the reference's state_dict keys (requires_grad leaves allowed).
    Returns (location_pred [B,5,2], location_preds [B,4,5,2], new_buffers dict)."""
    new_buffers = {}
    bn = _BN(sd, train, new_buffers)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach()

    if normalize:   # common.py:108-109 via image.py:71
        mean = torch.tensor(RGB_MEAN, dtype=x.dtype).reshape(1, 3, 1, 1)
        std = torch.tensor(RGB_STD, dtype=x.dtype).reshape(1, 3, 1, 1)
        x = (x - mean) / std
    # stem, resnet.py:148-152
    h = F.conv2d(x, sd["conv.conv1.weight"], None, stride=2, padding=3)
    tap("stem.raw", h)
    h = F.relu(bn(h, "conv.bn1"))
    h = F.max_pool2d(h, 3, 2, 1)
    tap("stem.pool", h)
    # BasicBlocks, resnet.py:38-54 / _make_layer :132-146
    for li, nblocks in enumerate(RESNET_LAYERS[backbone]):
        for bi in range(nblocks):
            p = "conv.layer%d.%d" % (li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            identity = h
            out = F.conv2d(h, sd[p + ".conv1.weight"], None, stride=stride, padding=1)
            out = F.relu(bn(out, p + ".bn1"))
            out = F.conv2d(out, sd[p + ".conv2.weight"], None, stride=1, padding=1)
            out = bn(out, p + ".bn2")
            if (p + ".downsample.0.weight") in sd:
                identity = F.conv2d(h, sd[p + ".downsample.0.weight"], None, stride=stride)
                identity = bn(identity, p + ".downsample.1")
            h = F.relu(out + identity)
            tap(p, h)
    b, c, kh, kw = h.shape
    # late fusion of speed, image.py:77-79
    vel = velocity.to(h.dtype)[:, None, None, None].repeat(1, 128, kh, kw)
    h = torch.cat((h, vel), dim=1)
    # decoder, image.py:37-47: BN -> ConvTranspose2d(3,2,1,1)+bias -> ReLU, three times
    for bn_i, dc_i in ((0, 1), (3, 4), (6, 7)):
        h = bn(h, "deconv.%d" % bn_i)
        h = F.conv_transpose2d(h, sd["deconv.%d.weight" % dc_i], sd["deconv.%d.bias" % dc_i],
                               stride=2, padding=1, output_padding=1)
        h = F.relu(h)
        tap("deconv.%d" % dc_i, h)
    H, W = h.shape[2], h.shape[3]
    pos_x, pos_y = spatial_grid(W, H, h.dtype)
    preds = []
    logits_all = []
    for k in range(COMMANDS):   # heads, image.py:54-60 ; SpatialSoftmax common.py:136-152
        p = "location_pred.%d" % k
        z = bn(h, p + ".0")
        logit = F.conv2d(z, sd[p + ".1.weight"], sd[p + ".1.bias"])
        logits_all.append(logit)
        wgt = F.softmax(logit.reshape(-1, H * W), dim=-1)
        ex = (pos_x * wgt).sum(dim=1, keepdim=True)
        ey = (pos_y * wgt).sum(dim=1, keepdim=True)
        preds.append(torch.cat([ex, ey], 1).view(-1, STEPS, 2))
    tap("logits", torch.stack(logits_all, 1))
    location_preds = torch.stack(preds, dim=1)                       # image.py:83
    location_pred = (command_onehot.to(h.dtype)[:, :, None, None] * location_preds).sum(1)  # common.py:29-35
    return location_pred, location_preds, new_buffers


# ----------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------
def phase0_target(teacher_pred, w=384, h=160, fov=90, world_y=1.4, fixed_offset=4.0):
    """CoordConverter.__call__ + _project_image_xy, training/train_image_phase0.py:54-79.
    float64 numpy on the host, pinhole projection (what cv2.projectPoints does with zero
    rotation/translation and no distortion), clip to the image, back to float32 pixels."""
    t = teacher_pred.detach().cpu().numpy().astype(np.float32)
    t = (t + 1) * CROP_SIZE / 2
    t[:, :, 1] = CROP_SIZE - t[:, :, 1]
    t[:, :, 0] -= CROP_SIZE / 2
    t = t / PIXELS_PER_METER
    t[:, :, 1] += fixed_offset
    f = w / (2 * np.tan(fov * np.pi / 360))
    X = t[..., 0].astype(np.float64)
    Z = t[..., 1].astype(np.float64)
    u = np.clip(f * X / Z + w / 2, 0, w)
    v = np.clip(f * world_y / Z + h / 2, 0, h)
    return torch.from_numpy(np.stack([u, v], -1).astype(np.float32))


def phase0_loss(pred, target_px, w=384, h=160):
    """LocationLoss.forward, training/train_image_phase0.py:86-89 -> per-sample [B]."""
    img = torch.tensor([w, h], dtype=pred.dtype)
    loc = target_px.to(pred.dtype) / (0.5 * img) - 1
    return torch.mean(torch.abs(pred - loc), dim=(1, 2))


def phase1_convert(p, w=384, h=160, fov=90, world_y=1.4, fixed_offset=4.0):
    """CoordConverter.__call__, training/train_image_phase1.py:43-64 (differentiable)."""
    img = torch.tensor([w, h], dtype=p.dtype)
    cam = (p + 1) * img / 2
    f = w / (2 * np.tan(fov * np.pi / 360))
    xt = (cam[..., 0] - w / 2) / f
    yt = (cam[..., 1] - h / 2) / f
    world_z = world_y / yt
    world_x = world_z * xt
    mx = world_x * PIXELS_PER_METER + CROP_SIZE / 2
    my = CROP_SIZE - world_z * PIXELS_PER_METER + fixed_offset * PIXELS_PER_METER
    return torch.stack([mx, my], dim=-1)


def phase1_loss(pred_map, teac):
    """LocationLoss.forward, training/train_image_phase1.py:66-70 -> per-sample [B]."""
    return torch.mean(torch.abs(pred_map / (0.5 * CROP_SIZE) - 1 - teac), dim=(1, 2, 3))


def birdview_loss(pred, gt, w=192, h=192):
    """LocationLoss(choice='l1'), training/train_birdview.py:33-54 -> per-sample [B]."""
    img = torch.tensor([w, h], dtype=pred.dtype)
    return torch.mean(torch.abs(pred - (gt / (0.5 * img) - 1.0)), dim=(1, 2))


# ----------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam(lr=1e-4) single-tensor semantics of torch 2.11;
    called at training/train_image_phase0.py:231,185.  In place; tensors with grad None skipped."""
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        m, v = exp_avg[k], exp_avg_sq[k]
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------------------
# whole steps (what train_or_eval does per batch)
# ----------------------------------------------------------------------------------------
def leafify(sd, dtype=torch.float32):
    """Detached copy of a state_dict whose floating tensors are autograd leaves."""
    out = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            t = v.detach().clone().to(dtype)
            is_param = not (k.endswith("running_mean") or k.endswith("running_var")
                            or k.endswith("pos_x") or k.endswith("pos_y"))
            t.requires_grad_(is_param)
            out[k] = t
        else:
            out[k] = v.detach().clone()
    return out


def param_keys(sd):
    return [k for k, v in sd.items() if v.dtype.is_floating_point and v.requires_grad]


def train_step(student_sd, teacher_sd, rgb, birdview, speed, command, phase, adam_state=None,
               lr=1e-4, student_backbone="resnet34", teacher_backbone="resnet18", taps=None):
    """One iteration of train_or_eval: training/train_image_phase0.py:166-185 (phase=0),
    training/train_image_phase1.py:174-205 (phase=1).  student_sd must come from leafify().
    Mutates student_sd (params, BN buffers) when adam_state is given.  Returns dict."""
    oh = one_hot(command)
    with torch.no_grad():
        t_pred, t_preds, _ = policy_forward(teacher_sd, birdview, speed, oh, teacher_backbone,
                                            train=False, normalize=False)
    for k in param_keys(student_sd):
        student_sd[k].grad = None
    pred, preds, newbuf = policy_forward(student_sd, rgb, speed, oh, student_backbone,
                                         train=True, normalize=True, taps=taps)
    if phase == 0:
        target = phase0_target(t_pred)
        loss = phase0_loss(pred, target)
    else:
        loss = phase1_loss(phase1_convert(preds), t_preds)
    loss_mean = loss.mean()
    loss_mean.backward()
    grads = {k: student_sd[k].grad for k in param_keys(student_sd)}
    out = dict(pred=pred.detach(), preds=preds.detach(), t_pred=t_pred, t_preds=t_preds,
               loss=loss.detach(), loss_mean=loss_mean.detach(), grads=grads, new_buffers=newbuf)
    if adam_state is not None:
        adam_state["step"] += 1
        with torch.no_grad():
            pk = [k for k in param_keys(student_sd) if grads[k] is not None]
            for k in pk:
                if k not in adam_state["m"]:
                    adam_state["m"][k] = torch.zeros_like(student_sd[k])
                    adam_state["v"][k] = torch.zeros_like(student_sd[k])
            adam_step({k: student_sd[k] for k in pk}, grads, adam_state["m"], adam_state["v"],
                      adam_state["step"], lr=lr)
            for k, v in newbuf.items():
                student_sd[k] = v
    return out


def birdview_train_step(sd, birdview, location, speed, command, adam_state=None, lr=1e-4,
                        backbone="resnet18"):
    """One iteration of training/train_birdview.py:116-129 (config 5)."""
    oh = one_hot(command)
    for k in param_keys(sd):
        sd[k].grad = None
    pred, preds, newbuf = policy_forward(sd, birdview, speed, oh, backbone, train=True, normalize=False)
    loss = birdview_loss(pred, location)
    loss_mean = loss.mean()
    loss_mean.backward()
    grads = {k: sd[k].grad for k in param_keys(sd)}
    out = dict(pred=pred.detach(), preds=preds.detach(), loss=loss.detach(),
               loss_mean=loss_mean.detach(), grads=grads, new_buffers=newbuf)
    if adam_state is not None:
        adam_state["step"] += 1
        with torch.no_grad():
            pk = [k for k in param_keys(sd) if grads[k] is not None]
            for k in pk:
                if k not in adam_state["m"]:
                    adam_state["m"][k] = torch.zeros_like(sd[k])
                    adam_state["v"][k] = torch.zeros_like(sd[k])
            adam_step({k: sd[k] for k in pk}, grads, adam_state["m"], adam_state["v"],
                      adam_state["step"], lr=lr)
            for k, v in newbuf.items():
                sd[k] = v
    return out


def new_adam_state():
    return {"step": 0, "m": {}, "v": {}}


# ----------------------------------------------------------------------------------------
# synthetic batches (SURVEY 8(c) golden-vector protocol / 8(d) config 1)
# ----------------------------------------------------------------------------------------
def synthetic_batch(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randint(0, 256, (B, 3, 160, 384), dtype=torch.uint8, generator=g).float() / 255
    bev = (torch.rand(B, 7, 192, 192, generator=g) > 0.8).float()
    speed = torch.rand(B, generator=g) * 10
    cmd = torch.randint(1, 5, (B,), generator=g).float()
    loc = torch.rand(B, 5, 2, generator=g) * 192
    return dict(rgb=rgb, birdview=bev, speed=speed, command=cmd, location=loc)
