"""Producer of per-sample hypo-network parameters (the `params` operand of the native path).

Mirrors meta_modules.HyperNetwork (/root/reference/meta_modules.py:11-54) and its initialisers
hyper_weight_init / hyper_bias_init (meta_modules.py:303-321): one ReLU FCBlock per parameter of
the hypo module, whose output is reshaped to ``[B, *param_shape]``.  The ReLU MLPs themselves are
ordinary PyTorch (out of the hot path); what matters here is the output contract -- names in
``meta_named_parameters()`` order, leading task axis -- which feeds
``hypo_net(model_input, params=...)`` and receives the per-task weight gradients.
"""
import math
from collections import OrderedDict

import torch
from torch import nn

from . import config, functional, modules


def _hyper_last_layer_init(m, bias_bound=None):
    """He-normal weights scaled by 1e-2; bias ~ U(+-bias_bound) (default 1/fan_in of this layer)."""
    if hasattr(m, "weight"):
        nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity="relu", mode="fan_in")
        with torch.no_grad():
            m.weight.mul_(1.0e-2)
    if hasattr(m, "bias"):
        bound = bias_bound
        if bound is None:
            bound = 1.0 / nn.init._calculate_fan_in_and_fan_out(m.weight)[0]
        with torch.no_grad():
            m.bias.uniform_(-bound, bound)


def hyper_weight_init(m, in_features_main_net):
    """Last layer of a hypernetwork head predicting a weight matrix (meta_modules.py:303-310)."""
    _hyper_last_layer_init(m, 1.0 / in_features_main_net)


def hyper_bias_init(m):
    """Last layer of a hypernetwork head predicting a bias vector (meta_modules.py:313-321)."""
    _hyper_last_layer_init(m, None)


class HyperNetwork(nn.Module):
    """meta_modules.py:11-54.  ``native_heads`` (default None = when the hypo-network runs in the bf16 mode): the last
    linear of the heads that predict a HIDDEN weight matrix (65,536 outputs each: all but ~2 % of the hypernetwork's
    output) goes through ``siren_b200_hyper_head``, which writes the fp32 ``[B, 256, 256]`` entry of ``hypo_params``
    AND, from the same pass over the head's weights, the two 16-bit tensor-core operands the hypo-network's kernels
    read (so that call converts nothing) and ``sum W^2`` for ``hypo_weight_loss``.  Everything else is the reference's
    flow: ReLU FCBlocks, names and shapes in ``meta_named_parameters()`` order."""

    def __init__(self, hyper_in_features, hyper_hidden_layers, hyper_hidden_features, hypo_module, native_heads=None):
        super().__init__()
        self.names = []
        self.nets = nn.ModuleList()
        self.param_shapes = []
        self.native_heads = native_heads
        block = getattr(hypo_module, "net", None)
        self._hypo_w0 = float(getattr(block, "_w0", 30.0) or 30.0)
        self._hypo_precision = getattr(block, "precision", None)
        for name, param in hypo_module.meta_named_parameters():
            self.names.append(name)
            self.param_shapes.append(param.size())
            hn = modules.FCBlock(in_features=hyper_in_features, out_features=int(math.prod(param.size())),
                                 num_hidden_layers=hyper_hidden_layers, hidden_features=hyper_hidden_features,
                                 outermost_linear=True, nonlinearity="relu")
            self.nets.append(hn)
            if "weight" in name:
                fan = param.size()[-1]
                self.nets[-1].net[-1].apply(lambda m, fan=fan: hyper_weight_init(m, fan))
            elif "bias" in name:
                self.nets[-1].net[-1].apply(lambda m: hyper_bias_init(m))

    def _hidden_batched(self, z):
        """The ReLU trunks of ALL heads as batched products: the heads share their architecture (meta_modules.py:32-34),
        so level j of every head is one ``baddbmm`` over the stacked weights ``[P, h_out, h_in]`` instead of P small
        linears -- the same numbers (fp32, TF32 off) in a tenth of the launches, forward and backward."""
        trunks = [list(net.net)[:-1] for net in self.nets]
        x = z.unsqueeze(0).expand(len(trunks), -1, -1)
        for j in range(len(trunks[0])):
            W = torch.stack([t[j][0].weight for t in trunks])
            b = torch.stack([t[j][0].bias for t in trunks])
            x = torch.relu(torch.baddbmm(b.unsqueeze(1), x, W.transpose(1, 2)))
        return x

    def forward(self, z):
        """z: [B, hyper_in_features] -> OrderedDict name -> [B, *param_shape]."""
        params = OrderedDict()
        native = self.native_heads
        if native is None:
            native = (self._hypo_precision or config.get_defaults()["precision"]) == "bf16"
        if not (native and torch.is_tensor(z) and z.is_cuda and z.dim() == 2 and z.dtype == torch.float32):
            for name, net, shape in zip(self.names, self.nets, self.param_shapes):      # the reference's flow
                params[name] = net(z).reshape((-1,) + tuple(shape))
            return params
        hidden = self._hidden_batched(z)
        for p, (name, net, shape) in enumerate(zip(self.names, self.nets, self.param_shapes)):
            last = net.net[-1][0]
            h = hidden[p]
            if tuple(shape) == (256, 256) and functional.hyper_head_supported(h, last.weight, last.bias):
                W, wk, wt, ss = functional._HyperHeadFn.apply(h, last.weight, last.bias, self._hypo_w0)
                params[name] = functional.attach_ops(W, wk, wt, ss, self._hypo_w0)
            else:
                params[name] = torch.addmm(last.bias, h, last.weight.t()).reshape((-1,) + tuple(shape))
        return params


def hypo_weight_loss(model_output):
    """loss_functions.hypo_weight_loss (loss_functions.py:279-287): mean square of all predicted hypo parameters.  A
    weight that came out of the native head carries its sum of squares (formed in that kernel, differentiable); the
    remaining small tensors are squared and summed as ONE concatenated vector instead of one reduction each."""
    weight_sum, total, rest = 0, 0, []
    for weight in model_output["hypo_params"].values():
        ss = getattr(weight, "_siren_sumsq", None)
        if ss is not None:
            weight_sum = weight_sum + ss
        elif weight.numel() > (1 << 20):      # too large to be worth a copy
            weight_sum = weight_sum + torch.sum(weight ** 2)
        else:
            rest.append(weight.reshape(-1))
        total += weight.numel()
    if rest:
        flat = rest[0] if len(rest) == 1 else torch.cat(rest)
        weight_sum = weight_sum + torch.sum(flat ** 2)
    return weight_sum * (1 / total)


class ConvolutionalNeuralProcessImplicit2DHypernetFourierFeatures(nn.Module):
    """meta_modules.py:175-232, the model train_mri_neural_process_ddp.py:204 builds: ConvImgEncoder -> HyperNetwork ->
    per-sample SingleBVPNet on Fourier features -> k-space data consistency.  Same constructor arguments, attribute
    names (``encoder``, ``hypo_net``, ``hyper_net``, ``dc``), state_dict keys and output dict.  What differs is where
    the work of the hypo path runs (each switchable, all on by default on a CUDA device in the bf16 mode):

    * the heads of the hypernetwork that predict hidden weights emit the kernels' operands (``HyperNetwork``);
    * with a lazy ``features.GaussianFourierFeatureTransform`` the features are built inside the kernels;
    * ``fuse_dc``: the data consistency is applied by the kernels' output epilogue and ``self.dc`` passes the
      tagged prediction on.

    ``partial_conv=True`` (PartialConvImgEncoder) is not mirrored: build that model from the reference's class after
    ``integration.patch_reference``."""

    def __init__(self, in_features, out_features, image_resolution=None, partial_conv=False, fourier_features_size=512,
                 latent_dim=256, hidden_features=256, num_hidden_layers=5, hyper_hidden_features=512,
                 hyper_hidden_layers=1, conv_kernel_size=3, num_conv_res_blocks=4, w0=30, precision=None, fuse_dc=True,
                 native_heads=None):
        super().__init__()
        if partial_conv:
            raise NotImplementedError("partial_conv=True: use the reference class with integration.patch_reference")
        from . import data_consistency
        self.dc = data_consistency.DataConsistencyInKspace(noise_lvl=None)
        self.encoder = modules.ConvImgEncoder(channel=2, image_resolution=image_resolution, hidden_size=latent_dim,
                                              kernel_size=conv_kernel_size, num_conv_res_blocks=num_conv_res_blocks)
        self.hypo_net = modules.SingleBVPNet(out_features=out_features, type="sine", sidelength=image_resolution,
                                             in_features=fourier_features_size, hidden_features=hidden_features,
                                             num_hidden_layers=num_hidden_layers, w0=w0, precision=precision)
        self.hypo_net.fuse_dc = bool(fuse_dc)
        self.hyper_net = HyperNetwork(hyper_in_features=latent_dim, hyper_hidden_layers=hyper_hidden_layers,
                                      hyper_hidden_features=hyper_hidden_features, hypo_module=self.hypo_net,
                                      native_heads=native_heads)

    def forward(self, model_input):
        embedding = model_input.get("embedding", None)
        if embedding is None:
            embedding = self.encoder(model_input["img_sparse"])
        hypo_params = self.hyper_net(embedding)
        model_output = self.hypo_net(model_input, params=hypo_params)
        out = model_output["model_out"]
        if "img_sparse" in model_input:
            out = self.dc(out, model_input["img_sparse"], model_input["dc_mask"])
        return {"model_in": model_output["model_in"], "model_out": out, "latent_vec": embedding,
                "hypo_params": hypo_params}

    def get_hypo_net_weights(self, model_input):
        embedding = self.encoder(model_input["img_sparse"])
        return self.hyper_net(embedding), embedding

    def freeze_hypernet(self):
        for param in self.hyper_net.parameters():
            param.requires_grad = False
        for param in self.encoder.parameters():
            param.requires_grad = False
