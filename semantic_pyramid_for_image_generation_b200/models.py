"""Drop-in replacements of the reference's `models.py` classes, running on hand-written sm_100a kernels.

Same constructor arguments, forward keywords, attribute names and `state_dict` keys/shapes as the reference
(models.py:10-519): `Generator(out_channels, latent_dimensions, channels_factor, number_of_classes)`,
`Discriminator(in_channels, channel_factor, number_of_classes)`, `VGG16(path_to_pre_trained_model, return_output)`.
The modules own real FP32 `nn.Parameter`s (`*.weight_orig`, `*.bias`, `*.embedding.weight`, `gamma`, ...) and
buffers (`*.weight_u`, `*.weight_v`, `running_mean/var`, `num_batches_tracked`); gradients arrive in `.grad` through
autograd.  The forward/backward arithmetic is in engine.py / vgg_engine.py over the C-ABI library -- there is
no torch-op or CPU fallback: CPU tensors raise.
"""
from typing import List, Optional, Union

import torch
import torch.nn as nn

from . import engine, vgg_engine
from ._native import call
from .engine import GradArena
from .spectral import LayerSpec, SNSet, SpectralNormHolder

BF16 = torch.bfloat16


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: the B200 path needs CUDA tensors (got %s); there is no CPU fallback" % (what, t.device))


# ------------------------------------------------------------------------------------------------
# containers with the reference's module tree (for state_dict keys and repr); they hold state only
# ------------------------------------------------------------------------------------------------
class ConditionalBatchNorm(nn.Module):
    """State of models.py:469-506: BatchNorm2d(affine=False, momentum=0.001) + Embedding(classes, 2C) = [1..1|0..0]."""

    def __init__(self, num_features: int, number_of_classes: int = 365) -> None:
        super().__init__()
        self.batch_norm = nn.BatchNorm2d(num_features=num_features, momentum=0.001, affine=False)
        self.embedding = nn.Embedding(num_embeddings=number_of_classes, embedding_dim=num_features * 2)
        self.embedding.weight.data[:, :num_features].fill_(1.)
        self.embedding.weight.data[:, num_features:].zero_()


class GeneratorResidualBlock(nn.Module):
    """State of models.py:278-339."""

    def __init__(self, in_channels: int, out_channels: int, feature_channels: int, number_of_classes: int = 365) -> None:
        super().__init__()
        self.main_block = nn.ModuleList([
            ConditionalBatchNorm(in_channels, number_of_classes),
            nn.LeakyReLU(negative_slope=0.2),
            nn.UpsamplingBilinear2d(scale_factor=2),
            SpectralNormHolder(out_channels, in_channels, 3, 3),
            ConditionalBatchNorm(out_channels, number_of_classes),
            nn.LeakyReLU(negative_slope=0.2),
            SpectralNormHolder(out_channels, out_channels, 3, 3)])
        self.residual_mapping = nn.Sequential(nn.UpsamplingBilinear2d(scale_factor=2),
                                              SpectralNormHolder(out_channels, in_channels, 1, 1))
        self.masked_feature_mapping = SpectralNormHolder(out_channels, feature_channels, 3, 3)


class LinearBlock(nn.Module):
    """State of models.py:342-375."""

    def __init__(self, in_features: int, out_features: int, feature_size: int) -> None:
        super().__init__()
        self.main_block = nn.Sequential(nn.LeakyReLU(negative_slope=0.2), SpectralNormHolder(out_features, in_features))
        self.masked_feature_mapping = SpectralNormHolder(out_features, feature_size)


class SelfAttention(nn.Module):
    """State of models.py:219-275 (gamma initialised to 1)."""

    def __init__(self, channels: int) -> None:
        super().__init__()
        self.query_convolution = SpectralNormHolder(channels // 8, channels, 1, 1)
        self.key_convolution = SpectralNormHolder(channels // 8, channels, 1, 1)
        self.value_convolution = SpectralNormHolder(channels // 2, channels, 1, 1)
        self.attention_convolution = SpectralNormHolder(channels, channels // 2, 1, 1)
        self.max_pooling = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)
        self.gamma = nn.Parameter(torch.ones(1, dtype=torch.float32))


class DiscriminatorInputResidualBlock(nn.Module):
    """State of models.py:378-419."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.main_block = nn.Sequential(SpectralNormHolder(out_channels, in_channels, 3, 3),
                                        nn.LeakyReLU(negative_slope=0.2),
                                        SpectralNormHolder(out_channels, out_channels, 3, 3))
        self.residual_mapping = SpectralNormHolder(out_channels, in_channels, 1, 1)
        self.downsampling = nn.AvgPool2d(kernel_size=(2, 2))


class DiscriminatorResidualBlock(nn.Module):
    """State of models.py:422-466."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.main_block = nn.Sequential(nn.LeakyReLU(negative_slope=0.2),
                                        SpectralNormHolder(out_channels, in_channels, 3, 3),
                                        nn.LeakyReLU(negative_slope=0.2),
                                        SpectralNormHolder(out_channels, out_channels, 3, 3))
        self.residual_mapping = SpectralNormHolder(out_channels, in_channels, 1, 1)
        self.downsampling = nn.AvgPool2d(kernel_size=(2, 2))


def _sn_holders(module):
    return [(name, m) for name, m in module.named_modules() if isinstance(m, SpectralNormHolder)]


# ------------------------------------------------------------------------------------------------
# Generator
# ------------------------------------------------------------------------------------------------
class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, grad_on, z, class_id, nfeat, *rest):
        # grad_on: torch.is_grad_enabled() at the call site.  Inside Function.forward grad mode is always off and
        # ctx.needs_input_grad only mirrors requires_grad, so without it a no_grad() forward would save (and, for the
        # attention map, write) everything a backward needs.
        features, masks = rest[:nfeat], rest[nfeat:2 * nfeat]
        if grad_on and (ctx.needs_input_grad[2] or any(ctx.needs_input_grad[5:5 + nfeat])):
            # the reference back-propagates into z and the features and discards the results (model_wrapper.py:168-190);
            # this path does not compute them: refuse rather than hand back silently-zero gradients.  (Masks / class ids,
            # which the reference's loader flags as requiring grad out of habit, data.py:80,87, are ignored.)
            raise RuntimeError("Generator: gradients w.r.t. the latent input or the VGG features are not implemented "
                               "(detach them; the reference never uses these gradients)")
        save = grad_on and any(ctx.needs_input_grad)
        img, c = engine.generator_forward(module, z, features, masks, class_id, save)
        ctx.module, ctx.c = module, c
        return img

    @staticmethod
    def backward(ctx, g_img):
        module = ctx.module
        grad = engine.generator_backward(module, ctx.c, g_img)
        ctx.c = None
        nparams = len(module._ga.params)
        needs = ctx.needs_input_grad[-nparams:]
        module._last_grad_arena = grad
        pg = module._ga.views(grad, needs)
        return (None,) * (len(ctx.needs_input_grad) - nparams) + tuple(pg)


class Generator(nn.Module):
    '''
    Generator network (reference models.py:10-99)
    '''

    def __init__(self, out_channels: int = 3, latent_dimensions: int = 128,
                 channels_factor: Union[int, float] = 1, number_of_classes: int = 365) -> None:
        super(Generator, self).__init__()
        self.latent_dimensions = latent_dimensions
        ch = [int(512 // channels_factor), int(512 // channels_factor), int(512 // channels_factor),
              int(256 // channels_factor), int(128 // channels_factor), int(64 // channels_factor)]
        if any(c % 8 != 0 or c < 16 for c in ch):
            raise ValueError("channels_factor=%s gives channel counts %s; the tensor-core path needs multiples of 8 "
                             "(>= 16)" % (channels_factor, ch))
        self.linear_layer = SpectralNormHolder(latent_dimensions, latent_dimensions)
        self.linear_block_1 = LinearBlock(in_features=latent_dimensions, out_features=365, feature_size=365)
        self.linear_block_2 = LinearBlock(in_features=365, out_features=2048, feature_size=4096)
        self.convolution_layer = nn.Sequential(nn.LeakyReLU(negative_slope=0.2), SpectralNormHolder(ch[0], 128, 1, 1))
        self.main_path = nn.ModuleList([
            GeneratorResidualBlock(ch[0], ch[1], 513, number_of_classes),
            GeneratorResidualBlock(ch[1], ch[2], 513, number_of_classes),
            GeneratorResidualBlock(ch[2], ch[3], 257, number_of_classes),
            SelfAttention(channels=ch[3]),
            GeneratorResidualBlock(ch[3], ch[4], 129, number_of_classes),
            GeneratorResidualBlock(ch[4], ch[5], 65, number_of_classes)])
        self.final_block = nn.Sequential(
            nn.UpsamplingBilinear2d(scale_factor=2),
            nn.BatchNorm2d(ch[5]),
            nn.LeakyReLU(negative_slope=0.2),
            SpectralNormHolder(ch[5], ch[5], 3, 3),
            nn.LeakyReLU(negative_slope=0.2),
            SpectralNormHolder(out_channels, ch[5], 1, 1))
        self._ga = GradArena(self)
        specs = []
        for name, h in _sn_holders(self):
            if name.endswith("masked_feature_mapping") and len(h.shape) == 4:
                specs.append(LayerSpec(name, h, pack_cin=h.shape[1] - 1, stencil=True))
            elif len(h.shape) == 4 and name != "final_block.5":
                specs.append(LayerSpec(name, h, pack_cin=h.shape[1]))
            else:
                specs.append(LayerSpec(name, h))  # FP32 consumers (linear layers, the C->3 tail): sigma only
        self._sn = SNSet(specs, self._ga.offset)
        self._last_grad_arena = None

    def forward(self, input: torch.Tensor, features: List[torch.Tensor],
                masks: List[torch.Tensor] = None, class_id: torch.Tensor = None) -> torch.Tensor:
        '''
        Forward pass
        :param input: (torch.Tensor) Input latent tensor (B, latent_dimensions)
        :param features: (List[torch.Tensor]) List of the seven vgg16 features
        :param masks: (List[torch.Tensor]) List of the seven masks
        :param class_id: (torch.Tensor) Class one-hot tensor (B, number_of_classes)
        :return: (torch.Tensor) Generated output image (B, out_channels, 256, 256) in [-1, 1]
        '''
        _require_cuda(input, "Generator.forward")
        if masks is None or class_id is None:
            raise RuntimeError("Generator.forward needs masks and class_id (the reference dereferences both, "
                               "models.py:78,95)")
        if len(features) != 7 or len(masks) != 7:
            raise RuntimeError("Generator.forward expects 7 features and 7 masks")
        return _GeneratorFn.apply(self, torch.is_grad_enabled(), input, class_id, len(features), *features, *masks,
                                  *self._ga.params)


# ------------------------------------------------------------------------------------------------
# Discriminator
# ------------------------------------------------------------------------------------------------
class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, grad_on, img, class_id, *params):
        save = grad_on and any(ctx.needs_input_grad)
        out, c = engine.discriminator_forward(module, img, class_id, save)
        ctx.module, ctx.c = module, c
        return out

    @staticmethod
    def backward(ctx, g_out):
        module = ctx.module
        needs = ctx.needs_input_grad[4:]
        grad, g_img = engine.discriminator_backward(module, ctx.c, g_out, any(needs), ctx.needs_input_grad[2])
        ctx.c = None
        pg = [None] * len(needs)
        if grad is not None:
            ga, prev = module._ga, module._last_grad_arena
            p0 = ga.params[0]
            task = getattr(torch._C, "_current_graph_task_id", lambda: -1)()
            same_pass = prev is not None and task >= 0 and getattr(module, "_grad_task", -1) == task
            earlier_pass = prev is not None and p0.grad is not None and \
                p0.grad.data_ptr() == prev.data_ptr() + 4 * ga.offset(p0)
            cur = torch.cuda.current_stream(grad.device)
            if all(needs) and prev is not None and prev.device == grad.device and (same_pass or earlier_pass):
                # D(real) and D(fake) (model_wrapper.py:153-160) both feed every parameter.  The views of the first
                # arena are either still waiting in autograd's input buffers (same backward pass) or already are the
                # .grad tensors (an earlier pass): one flat add into that arena replaces 56 per-parameter additions.
                # The first pass may have run on another stream: order the add after it, and that stream's later work
                # (optimizer, all-reduce) after the add.
                owner = getattr(module, "_grad_stream", None)
                if owner is not None and owner != cur:
                    cur.wait_stream(owner)
                call("spyr_add_inplace", prev.data_ptr(), grad.data_ptr(), grad.numel())
                if owner is not None and owner != cur:
                    owner.wait_stream(cur)
            else:
                module._last_grad_arena = grad
                module._grad_task = task
                module._grad_stream = cur
                pg = module._ga.views(grad, needs)
        return (None, None, g_img, None) + tuple(pg)


class Discriminator(nn.Module):
    '''
    Discriminator network (reference models.py:102-155); returns a (B, B, 128) tensor like the reference
    '''

    def __init__(self, in_channels: int = 3, channel_factor: Union[int, float] = 1, number_of_classes: int = 365):
        super(Discriminator, self).__init__()
        if in_channels != 3:
            raise ValueError("the B200 discriminator path supports in_channels=3")
        ch = [int(64 // channel_factor), int(128 // channel_factor), int(256 // channel_factor),
              int(256 // channel_factor), int(256 // channel_factor), int(512 // channel_factor),
              int(768 // channel_factor)]
        if any(c % 8 != 0 or c < 16 for c in ch):
            raise ValueError("channel_factor=%s gives channel counts %s; the tensor-core path needs multiples of 8 "
                             "(>= 16)" % (channel_factor, ch))
        self.layers = nn.Sequential(
            DiscriminatorInputResidualBlock(in_channels, ch[0]),
            DiscriminatorResidualBlock(ch[0], ch[1]),
            DiscriminatorResidualBlock(ch[1], ch[2]),
            SelfAttention(channels=ch[2]),
            DiscriminatorResidualBlock(ch[2], ch[3]),
            DiscriminatorResidualBlock(ch[3], ch[4]),
            DiscriminatorResidualBlock(ch[4], ch[5]),
            DiscriminatorResidualBlock(ch[5], ch[6]),
            nn.LeakyReLU(negative_slope=0.2),
            nn.AdaptiveAvgPool2d(output_size=(1, 1)),
            nn.Flatten(start_dim=1),
            SpectralNormHolder(128, ch[6]),
            nn.LeakyReLU(negative_slope=0.2))
        self.classification = SpectralNormHolder(1, 128)
        self.embedding = SpectralNormHolder(number_of_classes, 128, bias=False, init="normal")
        self._ga = GradArena(self)
        specs = []
        for name, h in _sn_holders(self):
            if name == "layers.0.main_block.0":
                specs.append(LayerSpec(name, h, pack_cin=32, pack_mode=1))
            elif name == "layers.0.residual_mapping":
                specs.append(LayerSpec(name, h, pack_cin=8, pack_mode=1))
            elif len(h.shape) == 4:
                specs.append(LayerSpec(name, h, pack_cin=h.shape[1]))
            else:
                specs.append(LayerSpec(name, h))
        self._sn = SNSet(specs, self._ga.offset)
        self._last_grad_arena = None

    def forward(self, input: torch.Tensor, class_id: torch.Tensor) -> torch.Tensor:
        '''
        Forward pass
        :param input: (torch.Tensor) Image (B, 3, H, W), real or fake
        :param class_id: (torch.Tensor) Class one-hot tensor (B, number_of_classes)
        :return: (torch.Tensor) Prediction of shape (B, B, 128)
        '''
        _require_cuda(input, "Discriminator.forward")
        return _DiscriminatorFn.apply(self, torch.is_grad_enabled(), input, class_id, *self._ga.params)


# ------------------------------------------------------------------------------------------------
# VGG-16
# ------------------------------------------------------------------------------------------------
class _VGGFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, grad_on, img):
        pk = module._pack()
        pools, y7, y8, c = vgg_engine.vgg_forward(pk, img, grad_on and ctx.needs_input_grad[2])
        ctx.pk, ctx.c = pk, c
        return tuple(pools) + (y7, y8)

    @staticmethod
    def backward(ctx, *grads):
        from . import ops
        g_pools = [g.contiguous() if g is not None else None for g in grads[:5]]
        # split-BF16 mode: a gradient that did not come from this package's kernels has no lo plane -> rebuild the pair
        g_pools = [g if g is None or (g.dtype == BF16 and ops.has_planes(g)) else ops.to_act(g.float()) for g in g_pools]
        g_img = vgg_engine.vgg_backward(ctx.pk, ctx.c, g_pools, grads[5], grads[6])
        ctx.c = None
        return None, None, g_img


class _VGGTrainFn(torch.autograd.Function):
    """Classifier training path (vgg_16_train.py:142-165): logits only, dropout active, gradients for all 32 parameters."""

    @staticmethod
    def forward(ctx, module, img, dropout, *params):
        pk = module._pack()
        _, _, y8, c = vgg_engine.vgg_forward(pk, img, True, dropout=dropout)
        ctx.module, ctx.pk, ctx.c = module, pk, c
        return y8

    @staticmethod
    def backward(ctx, g8):
        module = ctx.module
        convs = [module.vgg16.features[i] for i in vgg_engine.CONV_IDX]
        fcs = [module.vgg16.classifier[i] for i in (0, 3, 6)]
        new = lambda p: torch.empty_like(p, dtype=torch.float32)
        wg = {"conv": [(new(m.weight), new(m.bias)) for m in convs], "fc": [(new(m.weight), new(m.bias)) for m in fcs],
              "keep": [], "image": ctx.needs_input_grad[1]}
        g_img = vgg_engine.vgg_backward(ctx.pk, ctx.c, [None] * 5, None, g8.contiguous().float(), wg=wg)
        ctx.c = None
        by_param = {}
        for m, (gw, gb) in list(zip(convs, wg["conv"])) + list(zip(fcs, wg["fc"])):
            by_param[id(m.weight)], by_param[id(m.bias)] = gw, gb
        grads = tuple(by_param[id(p)] if need else None
                      for p, need in zip(module.vgg16.parameters(), ctx.needs_input_grad[3:]))
        return (None, g_img, None) + grads


class VGG16(nn.Module):
    '''
    VGG-16 feature pyramid (reference models.py:158-216).  The five spatial taps are returned as NCHW-shaped views of
    NHWC BF16 storage (what Generator / SemanticReconstructionLoss consume without a copy); fc7 / logits are FP32.
    '''

    def __init__(self, path_to_pre_trained_model: Optional[str] = None, return_output: Optional[bool] = False) -> None:
        super(VGG16, self).__init__()
        import torchvision
        self.return_output = return_output
        if path_to_pre_trained_model is not None:
            self.vgg16 = torch.load(path_to_pre_trained_model, weights_only=False)
        else:
            self.vgg16 = torchvision.models.vgg16(weights=None)
            self.vgg16.classifier[-1] = nn.Linear(in_features=4096, out_features=365, bias=True)
        self.vgg16.features = nn.ModuleList(list(self.vgg16.features))
        self.vgg16.classifier = nn.ModuleList(list(self.vgg16.classifier))
        self._pk = None
        self._pk_sig = None
        self._dropout_calls = 0

    def _pack(self):
        from . import ops
        sig = tuple((p.data_ptr(), p._version) for p in self.vgg16.parameters()) + (ops.SPLIT,)
        if self._pk is None or sig != self._pk_sig:
            self._pk = vgg_engine.VGGPack(self.vgg16)
            self._pk_sig = sig
        return self._pk

    def forward(self, input: torch.Tensor) -> List[torch.Tensor]:
        '''
        Forward pass
        :param input: (torch.Tensor) Images (B, 1 or 3, H, W), H and W powers of two
        :return: (List[torch.Tensor]) pool1..pool5 taps, ReLU(fc7), logits
        '''
        _require_cuda(input, "VGG16.forward")
        if input.shape[1] == 1:
            input = input.repeat_interleave(3, dim=1)
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        params = list(self.vgg16.parameters())
        if self.return_output and torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # fine-tuning of the encoder as a Places365 classifier (vgg_16_train.py): dropout only in train() mode
            dropout = None
            if self.training:
                p_drop = float(self.vgg16.classifier[2].p)
                if p_drop > 0.0:
                    self._dropout_calls += 1
                    dropout = (p_drop, int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF, self._dropout_calls * (1 << 24))
            return _VGGTrainFn.apply(self, input, dropout, *params)
        if self.training:
            raise RuntimeError("VGG16 in train() mode is the classifier fine-tuning path: construct it with "
                               "return_output=True and trainable parameters (vgg_16_train.py), or call .eval() as "
                               "model_wrapper.py:113 does for the frozen feature pyramid")
        outs = _VGGFn.apply(self, torch.is_grad_enabled(), input)
        if self.return_output:
            return outs[6]
        return [t.permute(0, 3, 1, 2) for t in outs[:5]] + [outs[5], outs[6]]


def init_weights(module: nn.Module) -> None:
    """Reference models.py:509-519 initialises Conv/Linear weights (Xavier-uniform) and zero biases; the holders of
    this package already do so at construction, so applying it again re-draws the weights."""
    if isinstance(module, SpectralNormHolder) and module.bias is not None:
        nn.init.xavier_uniform_(module.weight_orig)
        module.bias.data.fill_(0.)
