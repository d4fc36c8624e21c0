"""Host-side mirror of the reference's implicit-network modules (same names, constructor
arguments, state_dict keys and forward contracts), with the sine-MLP arithmetic routed to the
native sm_100a kernels.

Reference interface mirrored here (paths relative to /root/reference):
  BatchLinear      modules.py:11-27     nn.Linear + MetaModule, forward(input, params=None)
  Sine             modules.py:30-38
  FCBlock          modules.py:40-119    forward(coords, params=None, **kwargs),
                                        forward_with_activations(coords, params=None, retain_grad=False)
  SingleBVPNet     modules.py:122-170   forward(model_input, params=None) -> {'model_in', 'model_out'}
  sine_init / first_layer_sine_init     modules.py:641-654 (and the other inits FCBlock selects,
                                        modules.py:602-638)

Where the fast path applies: nonlinearity 'sine', outermost_linear=True, hidden_features 256 (8..255: zero-padded),
1..8 hidden layers, in_features <= 16 (<= 256 without coordinate derivatives), out_features <= 8, fp32 CUDA tensors.  Anything else
(ReLU hypernetwork MLPs, CPU tensors, fp64, other widths) runs the same composed PyTorch ops
the reference runs.  On a CUDA device inside the envelope the native library must load.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import config, functional
from . import meta as _meta


# ------------------------------------------------------------------------------------------
# initialisers (same distributions as the reference)
# ------------------------------------------------------------------------------------------
def _is_linear(m):
    return isinstance(m, nn.Linear)


def init_weights_normal(m):
    if _is_linear(m) and hasattr(m, "weight"):
        nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity="relu", mode="fan_in")


def init_weights_selu(m):
    if _is_linear(m) and hasattr(m, "weight"):
        nn.init.normal_(m.weight, std=1.0 / math.sqrt(m.weight.size(-1)))


def init_weights_elu(m):
    if _is_linear(m) and hasattr(m, "weight"):
        nn.init.normal_(m.weight, std=math.sqrt(1.5505188080679277) / math.sqrt(m.weight.size(-1)))


def init_weights_xavier(m):
    if _is_linear(m) and hasattr(m, "weight"):
        nn.init.xavier_normal_(m.weight)


def sine_init(m):
    """U(-sqrt(6/fan_in)/30, +sqrt(6/fan_in)/30); the /30 is hard-coded (modules.py:646)."""
    with torch.no_grad():
        if hasattr(m, "weight"):
            bound = np.sqrt(6.0 / m.weight.size(-1)) / 30.0
            m.weight.uniform_(-bound, bound)


def first_layer_sine_init(m):
    """U(-1/fan_in, +1/fan_in) (modules.py:649-654)."""
    with torch.no_grad():
        if hasattr(m, "weight"):
            bound = 1.0 / m.weight.size(-1)
            m.weight.uniform_(-bound, bound)


class Sine(nn.Module):
    def __init__(self, w0=30):
        super().__init__()
        self.w0 = w0

    def forward(self, input):
        return torch.sin(self.w0 * input)


def _activation_table(w0):
    return {
        "sine": (lambda: Sine(w0=w0), sine_init, first_layer_sine_init),
        "relu": (lambda: nn.ReLU(inplace=True), init_weights_normal, None),
        "sigmoid": (lambda: nn.Sigmoid(), init_weights_xavier, None),
        "tanh": (lambda: nn.Tanh(), init_weights_xavier, None),
        "selu": (lambda: nn.SELU(inplace=True), init_weights_selu, None),
        "softplus": (lambda: nn.Softplus(), init_weights_normal, None),
        "elu": (lambda: nn.ELU(inplace=True), init_weights_elu, None),
    }


# ------------------------------------------------------------------------------------------
# class factory: the same classes can be built on this package's meta.py or on the reference's
# own torchmeta base classes (see siren_mri_b200.integration.patch_reference)
# ------------------------------------------------------------------------------------------
def build_classes(MetaModule, MetaSequential, get_subdict):

    class BatchLinear(nn.Linear, MetaModule):
        """Linear layer accepting batched weights ``[B, out, in]`` / biases ``[B, out]`` via ``params``."""
        __doc__ = nn.Linear.__doc__

        def forward(self, input, params=None):
            if params is None:
                params = OrderedDict(self.named_parameters())
            bias = params.get("bias", None)
            weight = params["weight"]
            output = input.matmul(weight.transpose(-1, -2))
            output += bias.unsqueeze(-2)     # like the reference, a missing bias is an error
            return output

    class FCBlock(MetaModule):
        """Fully connected network whose weights can be swapped per call (hypernetworks)."""

        def __init__(self, in_features, out_features, num_hidden_layers, hidden_features,
                     outermost_linear=False, nonlinearity="relu", weight_init=None, w0=30,
                     precision=None, coord_derivs=None, coords_grad=None, backend=None):
            super().__init__()
            self.first_layer_init = None
            make_nl, nl_weight_init, first_layer_init = _activation_table(w0)[nonlinearity]
            nl = make_nl()      # one shared activation module, as in the reference
            self.weight_init = weight_init if weight_init is not None else nl_weight_init

            layers = [MetaSequential(BatchLinear(in_features, hidden_features), nl)]
            for _ in range(num_hidden_layers):
                layers.append(MetaSequential(BatchLinear(hidden_features, hidden_features), nl))
            if outermost_linear:
                layers.append(MetaSequential(BatchLinear(hidden_features, out_features)))
            else:
                layers.append(MetaSequential(BatchLinear(hidden_features, out_features), nl))
            self.net = MetaSequential(*layers)
            if self.weight_init is not None:
                self.net.apply(self.weight_init)
            if first_layer_init is not None:
                self.net[0].apply(first_layer_init)

            # native-path bookkeeping (plain attributes: nothing here enters the state_dict)
            self._sine = nonlinearity == "sine"
            self._outermost_linear = bool(outermost_linear)
            self._w0 = float(w0) if self._sine else 0.0
            self._n_layers = num_hidden_layers + 2
            self.precision = precision          # None -> config default at call time
            self.coord_derivs = coord_derivs
            self.coords_grad = coords_grad
            self.backend = backend

        # -- option resolution ---------------------------------------------------------------
        def _opt(self, name):
            v = getattr(self, name, None)
            return config.get_defaults()[name] if v is None else v

        def _native_inputs(self, coords, params, fourier=None):
            """(coords3d, weights, biases, restore_shape) when the native path serves this call."""
            if not (self._sine and self._outermost_linear) or self._opt("backend") != "auto":
                return None
            if not torch.is_tensor(coords) or not coords.is_cuda or coords.dtype != torch.float32:
                return None
            try:
                weights = [params["net.%d.0.weight" % l] for l in range(self._n_layers)]
                biases = [params["net.%d.0.bias" % l] for l in range(self._n_layers)]
            except KeyError:
                return None
            hid = weights[0].shape[-2]
            if 8 <= hid < 256 and all(b is not None for b in biases):
                # narrower nets (the MRI script has a 64-wide block) run on the 256-wide kernels with zero-padded
                # weights: a padded unit computes sin(w0 * 0) = 0 and feeds zeros on, so values and gradients are those
                # of the narrow net (F.pad is differentiable: the gradients come back sliced)
                pad = 256 - hid
                last = self._n_layers - 1
                weights = [nn.functional.pad(W, (0, pad if l > 0 else 0, 0, pad if l < last else 0))
                           for l, W in enumerate(weights)]
                biases = [nn.functional.pad(b, (0, pad)) if l < last else b for l, b in enumerate(biases)]
            shape = None
            c3 = coords
            per_task = weights[0].dim() == 3
            if coords.dim() == 2 and not per_task:
                c3, shape = coords.unsqueeze(0), "2d"
            elif coords.dim() > 3 and not per_task:
                c3, shape = coords.reshape(1, -1, coords.shape[-1]), tuple(coords.shape[:-1])
            elif coords.dim() != 3:
                return None
            derivs = int(self._opt("coord_derivs")) if (coords.requires_grad and torch.is_grad_enabled()) else 0
            if fourier is not None:
                derivs = 0      # the lazy prologue carries no coordinate derivatives
            if not functional.native_supported(c3, weights, biases, derivs, fourier=fourier, precision=self._opt("precision"),
                                               coords_grad=bool(self._opt("coords_grad"))):
                return None
            return c3, weights, biases, shape, derivs

        def _dc_operands(self, dc, c3, weights, derivs):
            """``dc = (k0, mask, noise_lvl)`` with k0 / mask as data_consistency.DataConsistencyInKspace takes them
            (``[B, o, nx, ny]``, channel first): the tuple functional.siren_mlp wants when the kernels can apply the
            blend to this call's output, else None (the caller's own DataConsistencyInKspace then does)."""
            if dc is None or derivs or bool(self._opt("coords_grad")):
                return None
            k0, mask, noise_lvl = dc
            o = weights[-1].shape[-2]
            T, N = c3.shape[0], c3.shape[1]
            for t in (k0, mask):
                if (not torch.is_tensor(t) or not t.is_cuda or t.requires_grad or t.dim() < 3 or t.shape[0] != T
                        or t.shape[1] != o or t.numel() != T * N * o):
                    return None
            return k0, mask, noise_lvl, True

        def forward(self, coords, params=None, dc=None, **kwargs):
            if params is None:
                params = OrderedDict(self.named_parameters())
            if dc is None:           # handed over by a patched reference SingleBVPNet (integration._carry_tags)
                dc = getattr(self, "_siren_pending_dc", None)
            # raw coordinates tagged by features.GaussianFourierFeatureTransform(lazy=True): the kernels build the
            # features of the first layer themselves; any other path materialises them here (features.py:31-41)
            fourier = getattr(coords, "_siren_fourier", None)
            if fourier is None:      # handed over by a patched reference SingleBVPNet (integration._carry_fourier_tag)
                pend = getattr(self, "_siren_pending_fourier", None)
                if pend is not None and torch.is_tensor(coords) and coords.shape[-1] == pend.shape[0]:
                    fourier = pend
            nat = self._native_inputs(coords, params, fourier)
            if nat is None:
                if fourier is not None:
                    coords = functional.fourier_features(coords, fourier)
                    nat = self._native_inputs(coords, params, None)
                    fourier = None
                if nat is None:
                    return self.net(coords, params=get_subdict(params, "net"))
            c3, weights, biases, shape, derivs = nat
            dc_ops = self._dc_operands(dc, c3, weights, derivs) if shape is None else None
            if dc_ops is not None:
                out = functional.siren_mlp(c3, weights, biases, w0=self._w0, precision=self._opt("precision"),
                                           fourier=fourier, dc=dc_ops)
                # data_consistency.DataConsistencyInKspace (ours, or the reference's after patch_reference) returns a
                # tagged prediction as it is: the blend has been applied, with this noise level
                out._siren_dc_done = float(dc_ops[2] or 0.0)
                return out
            out = functional.siren_mlp(c3, weights, biases, w0=self._w0, precision=self._opt("precision"),
                                       coord_derivs=derivs, coords_grad=bool(self._opt("coords_grad")), fourier=fourier)
            fn, jets = getattr(out, "_siren_composed", None), getattr(out, "_siren_jets", 0)
            if shape == "2d":
                out = out.squeeze(0)
                if fn is not None:
                    functional.tag_composed(out, lambda: fn().squeeze(0), jets)
            elif shape is not None:
                full = shape + (out.shape[-1],)
                out = out.reshape(full)
                if fn is not None:
                    functional.tag_composed(out, lambda: fn().reshape(full), jets)
            return out

        def forward_with_activations(self, coords, params=None, retain_grad=False):
            """Model output plus every intermediate activation (composed path, as the reference)."""
            if params is None:
                params = OrderedDict(self.named_parameters())
            activations = OrderedDict()
            x = coords.clone().detach().requires_grad_(True)
            activations["input"] = x
            for i, layer in enumerate(self.net):
                subdict = get_subdict(params, "net.%d" % i)
                for j, sublayer in enumerate(layer):
                    if isinstance(sublayer, BatchLinear):
                        x = sublayer(x, params=get_subdict(subdict, "%d" % j))
                    else:
                        x = sublayer(x)
                    if retain_grad:
                        x.retain_grad()
                    activations["_".join((str(sublayer.__class__), "%d" % i))] = x
            return activations

    class SingleBVPNet(MetaModule):
        """Canonical representation network of a boundary value problem (mode='mlp')."""

        def __init__(self, out_features=1, type="sine", in_features=2, mode="mlp", hidden_features=256,
                     num_hidden_layers=3, w0=30, **kwargs):
            super().__init__()
            self.mode = mode
            if mode != "mlp":
                raise NotImplementedError(
                    "siren_mri_b200.modules.SingleBVPNet serves mode='mlp'; for 'rbf'/'nerf' keep the "
                    "reference class and call siren_mri_b200.integration.patch_reference(modules) so "
                    "that its FCBlock is the native one.")
            if kwargs.get("downsample", False):
                raise NotImplementedError("downsample=True is served by the reference class (see above)")
            self.net = FCBlock(in_features=in_features, out_features=out_features,
                               num_hidden_layers=num_hidden_layers, hidden_features=hidden_features,
                               outermost_linear=True, nonlinearity=type, w0=w0,
                               precision=kwargs.get("precision"), coord_derivs=kwargs.get("coord_derivs"),
                               coords_grad=kwargs.get("coords_grad"), backend=kwargs.get("backend"))

        def forward(self, model_input, params=None):
            if params is None:
                params = OrderedDict(self.named_parameters())
            # a fresh leaf so that derivatives w.r.t. the coordinates can be taken (modules.py:151)
            coords_org = model_input["coords"].clone().detach().requires_grad_(True)
            fourier = getattr(model_input["coords"], "_siren_fourier", None)
            if fourier is not None:      # lazy Fourier prologue (features.py): model_in stays the RAW coordinates
                coords_org._siren_fourier = fourier
            # fuse_dc (attribute, or the process default): the k-space data consistency the MRI models apply right
            # behind this network (meta_modules.py:217-219) is applied by the kernels' output epilogue instead
            dc = None
            fuse = getattr(self, "fuse_dc", None)
            if (config.get_defaults()["fuse_dc"] if fuse is None else fuse) and "img_sparse" in model_input \
                    and "dc_mask" in model_input:
                dc = (model_input["img_sparse"], model_input["dc_mask"], getattr(self, "dc_noise_lvl", None))
            output = self.net(coords_org, get_subdict(params, "net"), dc=dc) if dc is not None \
                else self.net(coords_org, get_subdict(params, "net"))
            return {"model_in": coords_org, "model_out": output}

        def forward_with_activations(self, model_input):
            coords = model_input["coords"].clone().detach().requires_grad_(True)
            activations = self.net.forward_with_activations(coords)
            return {"model_in": coords, "model_out": activations.popitem(), "activations": activations}

    return BatchLinear, FCBlock, SingleBVPNet


MetaModule = _meta.MetaModule
MetaSequential = _meta.MetaSequential
get_subdict = _meta.get_subdict
BatchLinear, FCBlock, SingleBVPNet = build_classes(MetaModule, MetaSequential, get_subdict)


class Conv2dResBlock(nn.Module):
    """modules.py:433-450: two 5x5 convolutions with ReLU and a skip connection.  (cuDNN convolutions: the encoder is
    outside the hand-written path; kept so that the neural-process models build from this package.)"""

    def __init__(self, in_channel, out_channel=128):
        super().__init__()
        self.convs = nn.Sequential(nn.Conv2d(in_channel, out_channel, 5, 1, 2), nn.ReLU(),
                                   nn.Conv2d(out_channel, out_channel, 5, 1, 2), nn.ReLU())
        self.final_relu = nn.ReLU()

    def forward(self, x):
        return self.final_relu(self.convs(x) + x)


class ConvImgEncoder(nn.Module):
    """modules.py:340-380: sparse k-space image ``[B, channel, nx, ny]`` -> latent ``[B, hidden_size]`` (input
    convolution, 3x3 convolution, ``num_conv_res_blocks`` residual blocks, 1x1 convolution, then one linear layer over
    the pixels of every channel).  Same constructor, attribute names and state_dict keys as the reference."""

    def __init__(self, channel, image_resolution, hidden_size=256, kernel_size=3, num_conv_res_blocks=4):
        super().__init__()
        self.hidden_size = hidden_size
        pad = kernel_size // 2
        self.conv_theta = nn.Conv2d(channel, hidden_size // 2, kernel_size, 1, pad)
        self.relu = nn.ReLU(inplace=True)
        layers = [nn.Conv2d(hidden_size // 2, hidden_size, kernel_size, 1, pad), nn.ReLU()]
        layers += [Conv2dResBlock(hidden_size, hidden_size) for _ in range(num_conv_res_blocks)]
        layers.append(nn.Conv2d(hidden_size, hidden_size, 1, 1, 0))
        self.cnn = nn.Sequential(*layers)
        self.relu_2 = nn.ReLU(inplace=True)
        self.fc = nn.Linear(image_resolution[0] * image_resolution[1], 1)
        self.image_resolution = image_resolution

    def forward(self, I):
        o = self.cnn(self.relu(self.conv_theta(I)))
        return self.fc(self.relu_2(o).view(o.shape[0], self.hidden_size, -1)).squeeze(-1)


class SineLayer(nn.Module):
    """The notebook's ``SineLayer`` (explore_siren.ipynb cell 3): ``sin(omega_0 * linear(x))`` with its two
    initialisations (first layer U(+-1/in), others U(+-sqrt(6/in)/omega_0))."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega_0=30):
        super().__init__()
        self.omega_0 = omega_0
        self.is_first = is_first
        self.in_features = in_features
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.init_weights()

    def init_weights(self):
        with torch.no_grad():
            if self.is_first:
                self.linear.weight.uniform_(-1 / self.in_features, 1 / self.in_features)
            else:
                bound = np.sqrt(6 / self.in_features) / self.omega_0
                self.linear.weight.uniform_(-bound, bound)

    def forward(self, input):
        return torch.sin(self.omega_0 * self.linear(input))

    def forward_with_intermediate(self, input):
        intermediate = self.omega_0 * self.linear(input)
        return torch.sin(intermediate), intermediate


class Siren(nn.Module):
    """The notebook's ``Siren(in_features, hidden_features, hidden_layers, out_features, outermost_linear=False,
    first_omega_0=30, hidden_omega_0=30.)`` (explore_siren.ipynb cell 3): same layer objects and state_dict keys
    (``net.{i}.linear.weight`` / ``net.{L}.weight``), ``forward(coords) -> (output, coords)`` with ``coords`` a fresh
    leaf that requires grad.

    The native kernels take the whole signature: a first layer with its own frequency is the same function as one
    with ``omega = hidden_omega_0`` and weights / bias scaled by ``first_omega_0 / hidden_omega_0`` (a [H, d]
    multiply, differentiable), and a sine on the outermost layer (``outermost_linear=False``) is applied to the
    kernels' [N, out] output.  Widths other than 256, bias-free layers and CPU tensors run the layers as written."""

    def __init__(self, in_features, hidden_features, hidden_layers, out_features, outermost_linear=False,
                 first_omega_0=30, hidden_omega_0=30., precision=None, coord_derivs=None, backend=None):
        super().__init__()
        net = [SineLayer(in_features, hidden_features, is_first=True, omega_0=first_omega_0)]
        for _ in range(hidden_layers):
            net.append(SineLayer(hidden_features, hidden_features, is_first=False, omega_0=hidden_omega_0))
        if outermost_linear:
            final_linear = nn.Linear(hidden_features, out_features)
            with torch.no_grad():
                bound = np.sqrt(6 / hidden_features) / hidden_omega_0
                final_linear.weight.uniform_(-bound, bound)
            net.append(final_linear)
        else:
            net.append(SineLayer(hidden_features, out_features, is_first=False, omega_0=hidden_omega_0))
        self.net = nn.Sequential(*net)
        self.outermost_linear = bool(outermost_linear)
        self.first_omega_0, self.hidden_omega_0 = float(first_omega_0), float(hidden_omega_0)
        self.precision, self.coord_derivs, self.backend = precision, coord_derivs, backend

    def _opt(self, name):
        v = getattr(self, name, None)
        return config.get_defaults()[name] if v is None else v

    def _linears(self):
        return [m.linear if isinstance(m, SineLayer) else m for m in self.net]

    def _native(self, coords):
        if self._opt("backend") != "auto" or not coords.is_cuda or coords.dtype != torch.float32:
            return None
        lins = self._linears()
        if any(l.bias is None for l in lins) or coords.dim() not in (2, 3):
            return None
        c3 = coords.unsqueeze(0) if coords.dim() == 2 else coords
        scale = self.first_omega_0 / self.hidden_omega_0
        weights = [lins[0].weight * scale if scale != 1.0 else lins[0].weight] + [l.weight for l in lins[1:]]
        biases = [lins[0].bias * scale if scale != 1.0 else lins[0].bias] + [l.bias for l in lins[1:]]
        derivs = int(self._opt("coord_derivs")) if (coords.requires_grad and torch.is_grad_enabled()) else 0
        if not functional.native_supported(c3, weights, biases, derivs):
            return None
        out = functional.siren_mlp(c3, weights, biases, w0=self.hidden_omega_0, precision=self._opt("precision"),
                                   coord_derivs=derivs)
        fn, jets = getattr(out, "_siren_composed", None), getattr(out, "_siren_jets", 0)
        if not self.outermost_linear:
            out = torch.sin(self.hidden_omega_0 * out)
        if coords.dim() == 2:
            out = out.squeeze(0)
        if fn is not None:
            functional.tag_composed(out, lambda: self.net(coords), jets)
        return out

    def forward(self, coords):
        coords = coords.clone().detach().requires_grad_(True)     # allows to take derivative w.r.t. input
        output = self._native(coords)
        if output is None:
            output = self.net(coords)
        return output, coords

    def forward_with_activations(self, coords, retain_grad=False):
        """Model output plus every intermediate activation (layers as written; visualisation only)."""
        activations = OrderedDict()
        count = 0
        x = coords.clone().detach().requires_grad_(True)
        activations["input"] = x
        for layer in self.net:
            if isinstance(layer, SineLayer):
                x, intermed = layer.forward_with_intermediate(x)
                if retain_grad:
                    x.retain_grad()
                    intermed.retain_grad()
                activations["_".join((str(layer.__class__), "%d" % count))] = intermed
                count += 1
            else:
                x = layer(x)
                if retain_grad:
                    x.retain_grad()
            activations["_".join((str(layer.__class__), "%d" % count))] = x
            count += 1
        return activations
