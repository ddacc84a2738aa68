"""``Adam`` -- drop-in for ``torch.optim.Adam(net.parameters(), lr=...)`` as used by the reference
(training/train_image_phase0.py:231,183-185): same constructor / ``zero_grad()`` / ``step()`` calls, but the
update of all 23 M parameters is ONE launch of the native multi-tensor kernel (lbc_adam_step, torch 2.11
single-tensor semantics) over the module's flat parameter / gradient arrays, with the
data-parallel 1/world_size gradient scale folded in.  ``torch.optim.Adam`` itself also still works on the
module (parameters are ordinary leaf nn.Parameters) -- this class is the fast path, not a requirement.
"""
import torch

from . import _lib


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the reference uses plain Adam (wd=0, no amsgrad)")
        self.param_list = list(params)
        if not self.param_list:
            raise ValueError("optimizer got an empty parameter list")
        self.defaults = dict(lr=lr, betas=betas, eps=eps)
        self.param_groups = [dict(params=self.param_list, lr=lr, betas=betas, eps=eps)]
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self.grad_scale = 1.0
        self._ranges = None

    # -- torch.optim.Optimizer protocol used by the training loops
    def zero_grad(self, set_to_none=True):
        for p in self.param_list:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _locate(self):
        """Find the module whose flat arrays the parameters are views of (PolicyNetBase._flatten)."""
        from .common import owner_of
        owner = owner_of(self.param_list[0])
        if owner is None:
            raise _lib.LbcError("Adam: parameters do not belong to an lbc model that has run a forward pass "
                                "(use torch.optim.Adam for foreign parameters)")
        st = owner._lbc
        if not owner._views_intact():
            raise _lib.LbcError("Adam: parameters were moved / replaced since the last forward pass")
        if len(self.param_list) != len(st.param_views) or any(
                a is not b[0] for a, b in zip(self.param_list, st.param_views)):
            raise _lib.LbcError("Adam: expected exactly net.parameters() of one lbc model")
        return st

    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closure is not used by the reference loops")
        st = self._locate()
        st.param_epoch += 1       # the parameters change below without torch noticing (see common.py: lbc_net_infer)
        flat, gflat = st.flat_params, st.flat_grads
        have = [p.grad is not None for p, _, _, on in st.param_views if on]
        if not any(have):
            return
        for p, _, gview, on in st.param_views:
            if on and (p.grad is None or p.grad.data_ptr() != gview.data_ptr()):
                raise _lib.LbcError("Adam: gradients are not the native flat gradient views (mixed / replaced .grad)")
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
        elif self.exp_avg.numel() != flat.numel():
            raise _lib.LbcError("Adam: moment buffers hold %d elements, the model has %d (state of another model?)"
                                % (self.exp_avg.numel(), flat.numel()))
        elif self.exp_avg.device != flat.device:     # e.g. a checkpoint loaded with map_location='cpu'
            self.exp_avg = self.exp_avg.to(flat.device)
            self.exp_avg_sq = self.exp_avg_sq.to(flat.device)
        if self._ranges is None:
            # contiguous runs of parameters that receive gradients (conv.fc.* never does: grad is None)
            runs = []
            for p, view, _, on in st.param_views:
                if not on:
                    continue
                off, n = view.storage_offset(), view.numel()
                if runs and runs[-1][0] + runs[-1][1] == off:
                    runs[-1][1] += n
                else:
                    runs.append([off, n])
            self._ranges = runs
        self.step_count += 1
        g = self.param_groups[0]
        L = _lib.lib()
        sp = _lib.stream_ptr(flat.device)
        for off, n in self._ranges:
            _lib.check(L.lbc_adam_step(_lib.ptr(flat[off:off + n]), _lib.ptr(gflat[off:off + n]),
                                       _lib.ptr(self.exp_avg[off:off + n]), _lib.ptr(self.exp_avg_sq[off:off + n]),
                                       n, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                       self.step_count, float(self.grad_scale), sp))

    def state_dict(self):
        """Own (flat) format -- NOT interchangeable with torch.optim.Adam.state_dict(): one exp_avg / exp_avg_sq array over
        the model's flat parameter array instead of per-parameter entries."""
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq,
                    param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups])

    def load_state_dict(self, sd):
        if "state" in sd and "exp_avg" not in sd:
            raise _lib.LbcError("Adam.load_state_dict: this is a torch.optim.Adam state_dict (per-parameter layout); "
                                "lbc.Adam stores flat moment arrays (see Adam.state_dict)")
        self.step_count = int(sd["step"])
        self.exp_avg, self.exp_avg_sq = sd["exp_avg"], sd["exp_avg_sq"]
        if (self.exp_avg is None) != (self.exp_avg_sq is None):
            raise _lib.LbcError("Adam.load_state_dict: exp_avg / exp_avg_sq must both be present")
        self._ranges = None
