"""Host-side mirror of bird_view/models/common.py + resnet.py for the training hot path.

The classes here are *parameter containers*: they register exactly the tensors the reference
registers, under the same names, created by the same torch initialisers in the same order (so a
seeded construction yields bit-identical weights and ``state_dict()`` files are interchangeable,
SURVEY.md 8(b)).  None of them computes anything with torch: ``PolicyNetBase.forward`` hands the raw
device pointers to the native engine (include/lbc_b200.h) which runs the hand-written sm_100a kernels.
"""
import ctypes
import os
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib

KIND_IMAGE_RESNET34 = 0
KIND_BIRDVIEW_RESNET18 = 1
PRECISIONS = {"fp32": 0, "bf16": 1, "fp32tc": 2}

_RESNET_LAYERS = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3]}   # resnet.py:163-164


def default_precision():
    return os.environ.get("LBC_B200_PRECISION", "bf16")


def select_branch(branches, one_hot):
    """common.py:29-35 -- sum_k one_hot[b,k] * branches[b,k,...] (host helper; the engine fuses it)."""
    shape = [one_hot.shape[0], one_hot.shape[1]] + [1] * (branches.dim() - 2)
    return torch.sum(one_hot.reshape(shape) * branches, dim=1)


class SpatialSoftmaxBuffers(nn.Module):
    """Holds the pos_x / pos_y buffers of common.SpatialSoftmax (common.py:112-134) so that
    ``location_pred.k.2.pos_x`` / ``pos_y`` exist in the state_dict.  The soft-argmax itself runs in
    the fused head kernel."""

    def __init__(self, height, width, channel):
        super().__init__()
        self.height, self.width, self.channel = height, width, channel
        pos_x, pos_y = np.meshgrid(np.linspace(-1., 1., height), np.linspace(-1., 1., width))
        self.register_buffer("pos_x", torch.from_numpy(pos_x.reshape(height * width)).float())
        self.register_buffer("pos_y", torch.from_numpy(pos_y.reshape(height * width)).float())


class _BlockParams(nn.Module):
    """Parameters of resnet.BasicBlock (resnet.py:25-36)."""

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample


class _TrunkParams(nn.Module):
    """Parameters of resnet.ResNet with BasicBlocks (resnet.py:95-146), same creation order:
    stem, layer1..4 (downsample built before its block), avgpool, fc, then the kaiming pass."""

    def __init__(self, backbone, input_channel, bias_first):
        super().__init__()
        layers = _RESNET_LAYERS[backbone]
        self.conv1 = nn.Conv2d(input_channel, 64, 7, 2, 3, bias=bias_first)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        inplanes = 64
        for li, nblocks in enumerate(layers):
            planes = 64 << li
            stride = 1 if li == 0 else 2
            blocks = []
            for bi in range(nblocks):
                s = stride if bi == 0 else 1
                ds = None
                if s != 1 or inplanes != planes:
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes, 1, s, bias=False), nn.BatchNorm2d(planes))
                blocks.append(_BlockParams(inplanes, planes, s, ds))
                inplanes = planes
            setattr(self, "layer%d" % (li + 1), nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, 1000)     # constructed, never executed (resnet.py:112,148-159)
        for m in self.modules():           # resnet.py:114-119
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def _decoder_params():
    """image.py:37-47 / birdview.py:34-44: 3 x [BN, ConvTranspose2d(3,2,1,1), ReLU]."""
    return nn.Sequential(
        nn.BatchNorm2d(640), nn.ConvTranspose2d(640, 256, 3, 2, 1, 1), nn.ReLU(True),
        nn.BatchNorm2d(256), nn.ConvTranspose2d(256, 128, 3, 2, 1, 1), nn.ReLU(True),
        nn.BatchNorm2d(128), nn.ConvTranspose2d(128, 64, 3, 2, 1, 1), nn.ReLU(True),
    )


def _head_params(ow, oh, steps=5, commands=4):
    """image.py:54-60: per command BN(64) -> Conv2d(64,5,1) -> SpatialSoftmax(ow, oh, 5)."""
    return nn.ModuleList([
        nn.Sequential(nn.BatchNorm2d(64), nn.Conv2d(64, steps, 1, 1, 0), SpatialSoftmaxBuffers(ow, oh, steps))
        for _ in range(commands)
    ])


# id(nn.Parameter) -> weakref(owning module); kept outside the Parameter so that pickling / deepcopy of a module never meets
# a weakref (optim.Adam uses it to find the flat arrays its parameters are views of, and re-validates identity there)
_PARAM_OWNER = {}


def _infer_graph_max_b():
    """Eval-mode forwards of at most this many samples go through lbc_net_infer (one CUDA-graph replay); 0 disables."""
    try:
        return int(os.environ.get("LBC_B200_INFER_GRAPH_MAX_B", "16"))
    except ValueError:
        return 16


def owner_of(param):
    ref = _PARAM_OWNER.get(id(param))
    return ref() if ref is not None else None


class _NativeState:
    """Flat fp32 storage shared between the nn.Module (views) and the native engine (pointers)."""

    def __init__(self):
        self.handle = None
        self.device = None
        self.precision = None
        self.max_batch = 0
        self.flat_params = None
        self.flat_grads = None
        self.flat_grads_tmp = None
        self.flat_bufs = None
        self.param_views = None    # list of (param, data view, grad view, on_path)
        self.buffer_views = None   # list of (tensor, view)
        self.nbt = None
        self.table = None
        # the engine keeps the activations of the LATEST train-mode forward only: every such forward gets a generation
        # number, an autograd node may differentiate only the generation it produced, and only once
        self.fwd_gen = 0
        self.bwd_done_gen = -1
        # low-latency inference (lbc_net_infer): what the packed weight operands were last built from
        self.param_epoch = 0          # bumped by lbc.Adam (its kernel writes through the raw pointer)
        self.infer_seen = None        # (flat_params._version, param_epoch) at the previous graph replay

    def destroy(self):
        if self.handle is not None:
            try:
                _lib.lib().lbc_net_destroy(self.handle)
            except Exception:
                pass
            self.handle = None


class _PolicyFn(torch.autograd.Function):
    """Autograd node standing for the whole network: forward / backward are single native calls."""

    @staticmethod
    def forward(ctx, owner, image, velocity, command, anchor):
        pred, preds = owner._native_forward(image, velocity, command, True)
        ctx.owner = owner
        ctx.gen = owner._lbc.fwd_gen
        ctx.set_materialize_grads(False)
        return pred, preds

    @staticmethod
    def backward(ctx, d_pred, d_preds):
        ctx.owner._native_backward(d_pred, d_preds, ctx.gen)
        return None, None, None, None, None


class PolicyNetBase(nn.Module):
    """Shared machinery of ImagePolicyModelSS / BirdViewPolicyModelSS (common.ResnetBase, common.py:69-83)."""

    _lbc_kind = None
    _lbc_input_shape = None   # (C, H, W)

    def __init__(self, backbone, input_channel, bias_first, precision=None):
        super().__init__()
        if backbone not in _RESNET_LAYERS:
            raise ValueError("backbone %r is outside the hot path (resnet18/resnet34 BasicBlock nets only)" % backbone)
        self.conv = _TrunkParams(backbone, input_channel, bias_first)
        self.c = {"resnet18": -1, "resnet34": 512}[backbone]     # resnet.py:163-164 c_out quirk
        self.backbone = backbone
        self.input_channel = input_channel
        self.bias_first = bias_first
        self.lbc_precision = precision or default_precision()
        object.__setattr__(self, "_lbc", _NativeState())
        object.__setattr__(self, "_lbc_anchor", None)

    def __del__(self):
        st = self.__dict__.get("_lbc")
        if st is not None:
            if st.param_views and _PARAM_OWNER is not None:    # (None during interpreter shutdown)
                for p, _, _, _ in st.param_views:
                    _PARAM_OWNER.pop(id(p), None)
            st.destroy()

    # copy.deepcopy(net) / torch.save(net): the native handle (a ctypes pointer) and the flat views do not travel; the copy
    # owns plain tensors and builds its own engine lazily at its first forward
    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_lbc", None)
        d.pop("_lbc_anchor", None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        object.__setattr__(self, "_lbc", _NativeState())
        object.__setattr__(self, "_lbc_anchor", None)
        with torch.no_grad():
            for p in self.parameters():
                p.data = p.data.clone()
            for b in self.buffers():
                b.data = b.data.clone()

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__setstate__(copy.deepcopy(self.__getstate__(), memo))
        return new

    # ------------------------------------------------------------------ native state
    def _first_param(self):
        return next(self.parameters())

    def _ensure_native(self, B):
        st = self._lbc
        dev = self._first_param().device
        if self.lbc_precision not in PRECISIONS:
            raise ValueError("lbc_precision must be one of %s" % sorted(PRECISIONS))
        if not _lib.is_host_emulation() and dev.type != "cuda":
            raise _lib.LbcError("%s runs on CUDA (sm_100a) only; move the module with .to('cuda') -- there is no "
                                "CPU path" % type(self).__name__)
        recreate = (st.handle is None or st.device != dev or st.precision != self.lbc_precision
                    or B > st.max_batch)
        L = _lib.lib()
        if recreate:
            # keep current parameter values: views may still point into the old flat storage
            st.destroy()
            if dev.type == "cuda":
                torch.cuda.set_device(dev)
            max_batch = max(B, st.max_batch if st.device == dev else 0)
            h = ctypes.c_void_p()
            _lib.check(L.lbc_net_create(self._lbc_kind, PRECISIONS[self.lbc_precision], max_batch, ctypes.byref(h)))
            st.handle, st.device, st.precision, st.max_batch = h, dev, self.lbc_precision, max_batch
            st.table = self._read_table(h)
            st.flat_params = None
        elif dev.type == "cuda" and torch.cuda.current_device() != dev.index:
            # the library's scratch buffers / kernel attributes belong to the device that was current at first use: one
            # process drives ONE GPU; anything else must fail loudly, not corrupt memory
            raise _lib.LbcError("model lives on %s but the current CUDA device is cuda:%d -- liblbc_b200 drives one GPU per "
                                "process (torch.cuda.set_device first)" % (dev, torch.cuda.current_device()))
        if st.flat_params is None or not self._views_intact():
            self._flatten()
        return st

    def _read_table(self, h):
        L = _lib.lib()
        params, bufs = [], []
        for i in range(L.lbc_net_num_params(h)):
            name, ndim, shape = ctypes.c_char_p(), ctypes.c_int(), (ctypes.c_int * 4)()
            numel, off, onp = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
            _lib.check(L.lbc_net_param_info(h, i, ctypes.byref(name), ctypes.byref(ndim), ctypes.byref(shape),
                                            ctypes.byref(numel), ctypes.byref(off), ctypes.byref(onp)))
            params.append((name.value.decode(), tuple(shape[:ndim.value]), numel.value, off.value, bool(onp.value)))
        for i in range(L.lbc_net_num_buffers(h)):
            name, numel, off = ctypes.c_char_p(), ctypes.c_int64(), ctypes.c_int64()
            _lib.check(L.lbc_net_buffer_info(h, i, ctypes.byref(name), ctypes.byref(numel), ctypes.byref(off)))
            bufs.append((name.value.decode(), numel.value, off.value))
        return dict(params=params, buffers=bufs, n_params=L.lbc_net_total_params(h),
                    n_buffers=L.lbc_net_total_buffers(h))

    def _views_intact(self):
        st = self._lbc
        if st.param_views is None:
            return False
        for p, view, _, _ in st.param_views:
            if p.data_ptr() != view.data_ptr() or p.device != st.device:
                return False
        for t, view in st.buffer_views:
            if t.data_ptr() != view.data_ptr():
                return False
        return True

    def _flatten(self):
        """(Re)build the flat fp32 arrays and re-point every nn.Parameter / BN buffer at a view."""
        st = self._lbc
        dev = st.device
        named = list(self.named_parameters())
        table = st.table["params"]
        if [n for n, _ in named] != [t[0] for t in table]:
            raise _lib.LbcError("parameter table of the native engine does not match the module "
                                "(names/order differ) -- state_dict contract broken")
        flat = torch.zeros(st.table["n_params"], dtype=torch.float32, device=dev)
        grads = torch.zeros_like(flat)
        views = []
        with torch.no_grad():
            for (name, p), (_, shape, numel, off, on_path) in zip(named, table):
                if tuple(p.shape) != shape or p.dtype != torch.float32:
                    raise _lib.LbcError("parameter %s: expected fp32 %s, got %s %s" % (name, shape, p.dtype, tuple(p.shape)))
                v = flat[off:off + numel].view(shape)
                v.copy_(p.data)
                p.data = v
                _PARAM_OWNER[id(p)] = weakref.ref(self)
                views.append((p, v, grads[off:off + numel].view(shape), on_path))
            bufs = dict(self.named_buffers())
            fb = torch.zeros(st.table["n_buffers"], dtype=torch.float32, device=dev)
            bviews = []
            for name, numel, off in st.table["buffers"]:
                t = bufs[name]
                v = fb[off:off + numel]
                v.copy_(t.data)
                t.data = v
                bviews.append((t, v))
        st.flat_params, st.flat_grads, st.flat_bufs = flat, grads, fb
        st.flat_grads_tmp = None
        st.param_views, st.buffer_views = views, bviews
        st.nbt = [b for n, b in self.named_buffers() if n.endswith("num_batches_tracked")]
        _lib.check(_lib.lib().lbc_net_bind(st.handle, _lib.ptr(flat), _lib.ptr(grads), _lib.ptr(fb)))

    # ------------------------------------------------------------------ forward / backward
    def _check_inputs(self, x, velocity, command):
        C, H, W = self._lbc_input_shape
        if x.dtype == torch.uint8:
            # frames as stored on disk (u8, /255 applied on the device): [B,C,H,W] or [B,H,W,C]
            if x.dim() != 4 or (tuple(x.shape[1:]) != (C, H, W) and tuple(x.shape[1:]) != (H, W, C)):
                raise _lib.LbcError("expected uint8 input [B,%d,%d,%d] or [B,%d,%d,%d], got %s"
                                    % (C, H, W, H, W, C, tuple(x.shape)))
        elif x.dim() != 4 or tuple(x.shape[1:]) != (C, H, W):
            raise _lib.LbcError("expected input [B,%d,%d,%d], got %s" % (C, H, W, tuple(x.shape)))
        B = x.shape[0]
        if tuple(velocity.shape) != (B,) or tuple(command.shape) != (B, 4):
            raise _lib.LbcError("expected velocity [B] and one-hot command [B,4], got %s / %s"
                                % (tuple(velocity.shape), tuple(command.shape)))
        dev = self._first_param().device
        for t in (x, velocity, command):
            if t.device != dev:
                raise _lib.LbcError("input on %s but the model is on %s" % (t.device, dev))
            if t.dtype != torch.float32 and not (t is x and t.dtype == torch.uint8):
                raise _lib.LbcError("inputs must be float32 (or a uint8 frame tensor); got %s" % t.dtype)
        return B

    def _native_forward(self, x, velocity, command, train):
        B = self._check_inputs(x, velocity, command)
        st = self._ensure_native(B)
        x, velocity, command = x.contiguous(), velocity.contiguous(), command.contiguous()
        pred = torch.empty(B, 5, 2, dtype=torch.float32, device=st.device)
        preds = torch.empty(B, 4, 5, 2, dtype=torch.float32, device=st.device)
        if not train and st.device.type == "cuda" and B <= _infer_graph_max_b():
            # per-frame inference (ImageAgent.run_step): one CUDA-graph replay instead of ~130 launches
            seen = (st.flat_params._version, st.param_epoch)
            changed = seen != st.infer_seen
            st.infer_seen = seen
            u8 = x.dtype == torch.uint8
            layout = 1 if (u8 and tuple(x.shape[1:]) != tuple(self._lbc_input_shape)) else 0
            _lib.check(_lib.lib().lbc_net_infer(st.handle, None if u8 else _lib.ptr(x), _lib.ptr(x) if u8 else None, layout,
                                                _lib.ptr(velocity), _lib.ptr(command), B, 1 if changed else 0, _lib.ptr(pred),
                                                _lib.ptr(preds), _lib.stream_ptr(st.device)))
            return pred, preds
        if x.dtype == torch.uint8:
            layout = 0 if tuple(x.shape[1:]) == tuple(self._lbc_input_shape) else 1
            _lib.check(_lib.lib().lbc_net_forward_u8(st.handle, _lib.ptr(x), layout, _lib.ptr(velocity), _lib.ptr(command),
                                                     B, 1 if train else 0, _lib.ptr(pred), _lib.ptr(preds),
                                                     _lib.stream_ptr(st.device)))
        else:
            _lib.check(_lib.lib().lbc_net_forward(st.handle, _lib.ptr(x), _lib.ptr(velocity), _lib.ptr(command), B,
                                                  1 if train else 0, _lib.ptr(pred), _lib.ptr(preds),
                                                  _lib.stream_ptr(st.device)))
        if train:
            torch._foreach_add_(st.nbt, 1)     # BatchNorm2d.num_batches_tracked += 1
            st.fwd_gen += 1
        return pred, preds

    def _native_backward(self, d_pred, d_preds, gen=None):
        st = self._lbc
        if gen is not None and gen != st.fwd_gen:
            raise _lib.LbcError("backward() of a forward pass whose activations are gone: this model ran another train-mode "
                                "forward (generation %d -> %d) before backward.  The engine keeps one set of activations; "
                                "run logging / extra passes in eval() mode or after backward()" % (gen, st.fwd_gen))
        if gen is not None and st.bwd_done_gen == gen:
            raise _lib.LbcError("backward() called twice through the same forward pass (retain_graph is not supported: "
                                "the activations are consumed in place)")
        if not self._views_intact():
            raise _lib.LbcError("parameters were moved / replaced between forward and backward")
        fresh = all(p.grad is None for p, _, _, on in st.param_views if on)
        L = _lib.lib()
        if fresh:
            target = st.flat_grads
        else:
            if st.flat_grads_tmp is None:
                st.flat_grads_tmp = torch.zeros_like(st.flat_grads)
            target = st.flat_grads_tmp
        _lib.check(L.lbc_net_bind(st.handle, _lib.ptr(st.flat_params), _lib.ptr(target), _lib.ptr(st.flat_bufs)))
        dp = d_pred.contiguous().float() if d_pred is not None else None
        dps = d_preds.contiguous().float() if d_preds is not None else None
        _lib.check(L.lbc_net_backward(st.handle, _lib.ptr(dp), _lib.ptr(dps), _lib.stream_ptr(st.device)))
        if gen is not None:
            st.bwd_done_gen = gen
        if fresh:
            for p, _, gview, on in st.param_views:
                if on:
                    p.grad = gview
        else:   # gradient accumulation across several backward() calls without zero_grad()
            for p, _, gview, on in st.param_views:
                if not on:
                    continue
                off = gview.storage_offset()
                tv = target[off:off + gview.numel()].view(gview.shape)
                if p.grad is None:
                    p.grad = tv.clone()
                else:
                    p.grad += tv

    def forward(self, x, velocity, command):
        train = self.training
        if train and torch.is_grad_enabled():
            if self._lbc_anchor is None or self._lbc_anchor.device != x.device:
                object.__setattr__(self, "_lbc_anchor", torch.zeros((), device=x.device, requires_grad=True))
            pred, preds = _PolicyFn.apply(self, x, velocity, command, self._lbc_anchor)
        else:
            pred, preds = self._native_forward(x, velocity, command, train)
        if self.all_branch:
            return pred, preds
        return pred

    # ------------------------------------------------------------------ helpers for tests / tooling
    def lbc_flat_state(self, B=1):
        return self._ensure_native(max(B, self._lbc.max_batch or 1))

    def lbc_read_tap(self, name, numel):
        st = self._lbc
        out = torch.empty(numel, dtype=torch.float32, device=st.device)
        n = _lib.lib().lbc_net_read_tap(st.handle, name.encode(), _lib.ptr(out), numel, _lib.stream_ptr(st.device))
        if n < 0:
            raise _lib.LbcError(_lib.lib().lbc_last_error().decode())
        return out[:n]
