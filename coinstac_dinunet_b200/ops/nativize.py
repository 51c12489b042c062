"""``nativize(model)`` - run a *user-defined* ``nn.Module`` on the hand-written sm_100a kernels.

The reference's whole point is that users fill ``self.nn[...]`` with their own modules
(coinstac_dinunet/nn/basetrainer.py:30-34).  ``nativize`` walks such a module, finds the layer patterns the kernels
implement and re-routes them, without touching parameters, buffers, module names or ``state_dict`` keys (the original
modules stay registered where they were; only the *forward* of the containing ``nn.Sequential`` changes):

==============================================================  =========================================================
pattern (consecutive children of an ``nn.Sequential``)          runs on
==============================================================  =========================================================
``Conv3d(k3,s1,p1) -> BatchNorm3d -> ReLU -> MaxPool3d(2)``     ``ops.vbm.ConvBnReluPoolFn`` (tcgen05 implicit-GEMM conv, BN
(or a child tagged ``native_pattern = 'conv_bn_relu_pool'``      statistics in the conv epilogue, fused BN+ReLU+pool fwd/bwd);
with ``.conv`` / ``.bn``)                                        channel counts the kernels are not instantiated for are
                                                                zero-padded up to the next instantiated pair
``Linear -> BatchNorm1d -> ReLU`` / ``Linear -> BatchNorm1d``   ``ops.linear.linear_bn_relu`` (one launch each way, M <= 32)
``Linear -> ReLU`` / ``Linear``                                 ``ops.linear.linear`` (small-batch kernels or tcgen05 GEMM)
anything else                                                   the original module, untouched
==============================================================  =========================================================

Consecutive conv blocks form one *stack*: activations stay channels-last bf16 between them and are converted to / from
the PyTorch layout once at the stack's ends.  On CPU tensors (or with ``COINN_DISABLE_NATIVE=1``) the original forward
runs, so a nativized model still works everywhere.  ``NNTrainer`` calls this for every model when
``cache['native_ops']`` is set.
"""
import torch as _torch
from torch import nn as _nn

# (C_in, C_out) pairs with fprop + dgrad + wgrad instantiations (ops/conv3d.py, ops/conv3d_wgrad.py)
_CONV_PAIRS = ((16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 256))


def _padded_pair(cin, cout):
    """Smallest instantiated (C_in', C_out') with C_in' >= cin and C_out' >= cout, or None."""
    best = None
    for ci, co in _CONV_PAIRS:
        if ci >= cin and co >= cout and (best is None or ci * co < best[0] * best[1]):
            best = (ci, co)
    return best


def _is_conv(m):
    return (isinstance(m, _nn.Conv3d) and tuple(m.kernel_size) == (3, 3, 3) and tuple(m.stride) == (1, 1, 1)
            and tuple(m.padding) == (1, 1, 1) and tuple(m.dilation) == (1, 1, 1) and m.groups == 1
            and m.padding_mode == 'zeros' and (m.in_channels == 1 and m.out_channels == 16
                                               or _padded_pair(m.in_channels, m.out_channels) is not None))


def _is_bn3(m, ch):
    return isinstance(m, _nn.BatchNorm3d) and m.affine and m.track_running_stats and m.num_features == ch


def _is_pool2(m):
    def trip(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v,) * 3
    return (isinstance(m, _nn.MaxPool3d) and trip(m.kernel_size) == (2, 2, 2) and trip(m.stride or m.kernel_size) == (2, 2, 2)
            and trip(m.padding) == (0, 0, 0) and trip(m.dilation) == (1, 1, 1) and not m.ceil_mode and not m.return_indices)


def _tagged_block(m):
    if getattr(m, 'native_pattern', None) == 'conv_bn_relu_pool' and hasattr(m, 'conv') and hasattr(m, 'bn'):
        if _is_conv(m.conv) and _is_bn3(m.bn, m.conv.out_channels):
            return m.conv, m.bn
    return None


def _plan(children):
    """Greedy left-to-right pattern match over the ordered children of a Sequential -> list of steps."""
    steps, i, n = [], 0, len(children)
    while i < n:
        m = children[i]
        tagged = _tagged_block(m)
        if tagged is not None:
            steps.append(('conv', tagged[0], tagged[1], (i,)))
            i += 1
        elif (i + 3 < n and _is_conv(m) and _is_bn3(children[i + 1], m.out_channels)
              and isinstance(children[i + 2], _nn.ReLU) and _is_pool2(children[i + 3])):
            steps.append(('conv', m, children[i + 1], (i, i + 1, i + 2, i + 3)))
            i += 4
        elif isinstance(m, _nn.Linear) and i + 1 < n and isinstance(children[i + 1], _nn.BatchNorm1d) \
                and children[i + 1].affine and children[i + 1].num_features == m.out_features:
            relu = i + 2 < n and isinstance(children[i + 2], _nn.ReLU)
            steps.append(('lbr', m, children[i + 1], relu, tuple(range(i, i + 2 + int(relu)))))
            i += 2 + int(relu)
        elif isinstance(m, _nn.Linear):
            relu = i + 1 < n and isinstance(children[i + 1], _nn.ReLU)
            steps.append(('lin', m, relu, tuple(range(i, i + 1 + int(relu)))))
            i += 1 + int(relu)
        else:
            steps.append(('mod', m, (i,)))
            i += 1
    # merge consecutive conv steps into stacks
    merged = []
    for st in steps:
        if st[0] == 'conv' and merged and merged[-1][0] == 'stack':
            merged[-1][1].append((st[1], st[2]))
        elif st[0] == 'conv':
            merged.append(('stack', [(st[1], st[2])]))
        else:
            merged.append(st)
    return merged


def _conv_stack(blocks, x, training):
    """x: [N, C, D, H, W] (or [N, D, H, W] when C == 1) in the PyTorch layout -> [N, C', D', H', W'] fp32 values in
    channels-last memory.  ``blocks``: [(Conv3d, BatchNorm3d), ...]."""
    from .vbm import ConvBnReluPoolFn
    F = _torch.nn.functional
    first_conv = blocks[0][0]
    if x.dim() == 4:
        x = x.unsqueeze(1)
    if first_conv.in_channels == 1 and first_conv.out_channels == 16:
        h = x[:, 0]
        if h.dtype not in (_torch.float32, _torch.bfloat16):
            h = h.float()
        h = h.contiguous()
        have = 1
    else:
        h = x.permute(0, 2, 3, 4, 1).to(_torch.bfloat16)          # NDHWC bf16 (copy)
        have = h.shape[-1]
    for conv, bn in blocks:
        cin, cout = conv.in_channels, conv.out_channels
        mom = bn.momentum if bn.momentum is not None else 0.1
        nbt = bn.num_batches_tracked if training else None
        # A conv bias in front of BatchNorm changes nothing but the tracked mean (batch statistics absorb it, its gradient is
        # zero): the kernels run bias-free and the running mean is shifted by the bias so state_dicts stay interchangeable.
        has_bias = conv.bias is not None
        rm_in = bn.running_mean
        if has_bias and not training:
            rm_in = bn.running_mean - conv.bias.detach()
        if h.dim() == 4:                                           # fused first block (1 -> 16)
            if has_bias and training:
                with _torch.no_grad():
                    bn.running_mean.sub_(conv.bias.detach())       # track the bias-free mean inside, restore below
            h = ConvBnReluPoolFn.apply(h, conv.weight, bn.weight, bn.bias, rm_in if not training else bn.running_mean,
                                       bn.running_var, bn.eps, mom, training, 'auto', nbt)
            if has_bias and training:
                with _torch.no_grad():
                    bn.running_mean.add_(conv.bias.detach())
            have = cout
            continue
        ci_p, co_p = _padded_pair(cin, cout)
        if have < ci_p:                                            # zero input channels up to the instantiated width
            h = F.pad(h, (0, ci_p - have))
        elif have > ci_p:                                          # leftover zero channels of a wider previous block
            h = h[..., :ci_p]
        h = h.contiguous()
        if (ci_p, co_p) == (cin, cout):
            if has_bias and training:
                with _torch.no_grad():
                    bn.running_mean.sub_(conv.bias.detach())
            h = ConvBnReluPoolFn.apply(h, conv.weight, bn.weight, bn.bias, rm_in if not training else bn.running_mean,
                                       bn.running_var, bn.eps, mom, training, 'auto', nbt)
            if has_bias and training:
                with _torch.no_grad():
                    bn.running_mean.add_(conv.bias.detach())
        else:
            w = F.pad(conv.weight, (0, 0, 0, 0, 0, 0, 0, ci_p - cin, 0, co_p - cout))
            extra = co_p - cout
            g = _torch.cat([bn.weight, bn.weight.new_ones(extra)]) if extra else bn.weight
            b = _torch.cat([bn.bias, bn.bias.new_zeros(extra)]) if extra else bn.bias
            shift = conv.bias.detach() if has_bias else 0.0
            rm = _torch.cat([bn.running_mean - shift, bn.running_mean.new_zeros(extra)])
            rv = _torch.cat([bn.running_var, bn.running_var.new_ones(extra)])
            h = ConvBnReluPoolFn.apply(h, w, g, b, rm, rv, bn.eps, mom, training, 'auto', nbt)
            if training:
                with _torch.no_grad():
                    bn.running_mean.copy_(rm[:cout] + shift)
                    bn.running_var.copy_(rv[:cout])
        have = co_p                                                # channels >= cout are exactly zero (gamma 1, beta 0, W 0)
        # keep the zero tail if the next block wants it, it is dropped at the end otherwise
    cout = blocks[-1][0].out_channels
    if h.shape[-1] != cout:
        h = h[..., :cout]
    return h.permute(0, 4, 1, 2, 3).float()


class NativeSequential(_nn.Sequential):
    """An ``nn.Sequential`` whose CUDA forward follows a plan of native steps (see module docstring).  Created in place
    by ``nativize`` (class swap): children, names, parameters and ``state_dict`` are those of the original container."""

    _native_plan = None

    def forward(self, x):
        from . import native_available
        if self._native_plan is None or not x.is_cuda or not native_available():
            return super().forward(x)
        from .linear import linear as _linear, linear_bn_relu as _lbr
        for st in self._native_plan:
            kind = st[0]
            if kind == 'stack':
                x = _conv_stack(st[1], x, self.training)
            elif kind == 'lbr':
                lead = x.shape[:-1]
                x = _lbr(x, st[1], st[2], relu=st[3]).reshape(*lead, st[1].out_features)
            elif kind == 'lin':
                lead = x.shape[:-1]
                x = _linear(x.reshape(-1, x.shape[-1]), st[1].weight, st[1].bias, st[2]).float().reshape(*lead, st[1].out_features)
            else:
                x = st[1](x)
        return x


def nativize(model, report=None):
    """Re-route every matching ``nn.Sequential`` inside ``model`` (including ``model`` itself) through the native kernels.
    Returns ``model``; ``report`` (a list) receives one ``(path, kind, detail)`` row per native step.  Models that
    already implement their own native path (``is_native``) are returned unchanged."""
    if getattr(model, 'is_native', False) and not isinstance(model, NativeSequential):
        return model
    found = 0
    for name, mod in list(model.named_modules()):
        if not isinstance(mod, _nn.Sequential):
            continue
        plan = _plan(list(mod.children()))
        native_steps = [st for st in plan if st[0] != 'mod']
        if not native_steps:
            continue
        if not isinstance(mod, NativeSequential):
            mod.__class__ = type('Native' + type(mod).__name__, (NativeSequential, type(mod)), {}) \
                if type(mod) is not _nn.Sequential else NativeSequential
        mod._native_plan = plan
        found += len(native_steps)
        if report is not None:
            for st in native_steps:
                if st[0] == 'stack':
                    report.append((name, 'conv_stack', [(c.in_channels, c.out_channels) for c, _ in st[1]]))
                elif st[0] == 'lbr':
                    report.append((name, 'linear_bn_relu' if st[3] else 'linear_bn', (st[1].in_features, st[1].out_features)))
                else:
                    report.append((name, 'linear_relu' if st[2] else 'linear', (st[1].in_features, st[1].out_features)))
    if found:
        try:
            model.is_native = True          # trainers skip autocast / input casts for natively routed models
        except Exception:
            pass
    return model
