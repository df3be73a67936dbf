"""LBDRN network module -- drop-in for the reference's LBDRNmodel.py (same class names, constructor arguments,
`state_dict` keys / shapes / order and RNG consumption at construction; reference LBDRNmodel.py:7-82).

`forward(x)` on an explicit feature matrix is the cold compatibility path (plain torch ops).  The product path
is image-level and never materialises features: `decode_image`, `predict_image` and `eval_mse` hand the base
layer and the flat parameter vector to the fused CUDA kernels of liblbdrn_b200 (see lbdrn_fused.py).
"""
import math

import torch
from torch import nn


class Sine(nn.Module):
    """x -> sin(w0 * x)   (reference LBDRNmodel.py:7-13)"""

    def __init__(self, w0=1.):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


class SirenLayer(nn.Module):
    """Linear + SIREN initialisation + activation (reference LBDRNmodel.py:16-43).

    Weight and bias are redrawn U(-b, b) after nn.Linear's own initialisation, b = 1/dim_in for the first layer
    and sqrt(c/dim_in)/w0 otherwise, in this order -- the order matters for fixed-seed reproducibility."""

    def __init__(self, dim_in, dim_out, w0=30., c=6., is_first=False, use_bias=True, activation=None):
        super().__init__()
        self.dim_in, self.is_first = dim_in, is_first
        self.linear = nn.Linear(dim_in, dim_out, bias=use_bias)
        bound = 1 / dim_in if is_first else math.sqrt(c / dim_in) / w0
        nn.init.uniform_(self.linear.weight, -bound, bound)
        if use_bias:
            nn.init.uniform_(self.linear.bias, -bound, bound)
        self.activation = activation if activation is not None else Sine(w0)

    def forward(self, x):
        return self.activation(self.linear(x))


class LBDRNModel(nn.Module):
    """`num_layers` SirenLayers of width `dim_hidden` followed by a Linear -> Sigmoid head
    (reference LBDRNmodel.py:46-82)."""

    def __init__(self, dim_in, dim_hidden, dim_out=4, num_layers=1, w0=30., w0_initial=30., use_bias=True,
                 activation=None, final_activation=None):
        super().__init__()
        self.dim_in, self.dim_hidden, self.dim_out, self.num_layers = dim_in, dim_hidden, dim_out, num_layers
        self.w0, self.w0_initial = w0, w0_initial
        self.net = nn.Sequential(*[
            SirenLayer(dim_in if i == 0 else dim_hidden, dim_hidden, w0=w0_initial if i == 0 else w0,
                       is_first=i == 0, use_bias=use_bias, activation=activation)
            for i in range(num_layers)])
        self.last_layer = SirenLayer(dim_hidden, dim_out, w0=w0, use_bias=use_bias,
                                     activation=final_activation if final_activation is not None else nn.Sigmoid())
        self._fused_ok = (use_bias and final_activation is None and w0 == w0_initial and
                          (activation is None or isinstance(activation, nn.ReLU)))
        self._relu = isinstance(activation, nn.ReLU)

    def forward(self, x):
        return self.last_layer(self.net(x))

    # ---- flat parameter vector in state_dict order: the nn sub-stream layout (encode.py:123-128) ----------
    def flat_params(self):
        return torch.cat([v.detach().reshape(-1) for v in self.state_dict().values()]).to(torch.float32)

    def load_flat_params(self, flat):
        """Inverse of `flat_params` (the slicing of decode.py:114-120)."""
        flat = torch.as_tensor(flat, dtype=torch.float32).reshape(-1)
        need = sum(v.numel() for v in self.state_dict().values())
        if flat.numel() != need:
            raise ValueError(f"parameter vector has {flat.numel()} values, model needs {need}")
        sd, k = {}, 0
        for name, ref in self.state_dict().items():
            n = ref.numel()
            sd[name] = flat[k:k + n].reshape(ref.shape).to(ref.device)
            k += n
        self.load_state_dict(sd)

    # ---- image-level fast paths (CUDA only; no CPU fallback) ----------------------------------------------
    def _require_fused(self):
        if not self._fused_ok:
            raise NotImplementedError("fused kernels implement Sine(w0)/ReLU hidden layers with a Sigmoid head and bias")

    def decode_image(self, base, K, D, flags=None, path="auto", device=None):
        """uint16 CHW reconstruction of the CHW base layer: (base << K) + round(y * (2^K-1))  (decode.py:122-134)."""
        import lbdrn_fused
        self._require_fused()
        return lbdrn_fused.decode_image(base, self.flat_params(), K, D, self.dim_hidden, self.num_layers, flags=flags,
                                        relu=self._relu, w0=self.w0, path=path, device=device)

    def predict_image(self, base, D, flags=None, device=None):
        """Network output y [H*W, C] float32 for every pixel of the CHW base layer (model(x) of decode.py:130)."""
        import lbdrn_fused
        self._require_fused()
        return lbdrn_fused.predict_image(base, self.flat_params(), D, self.dim_hidden, self.num_layers, flags=flags,
                                         relu=self._relu, w0=self.w0, device=device)
