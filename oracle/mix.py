"""Oracle for SURVEY row a3: the AdvMix per-pixel convex mix.

TEST INFRASTRUCTURE - see oracle/__init__.py.
Restates lib/core/function.py:137-146 verbatim with torch CPU fp32 ops.
"""
import torch
import torch.nn.functional as F


def mix_from_logits(inputs, logits):
    """inputs: list of K [B,C,H,W] fp32; logits [B,K,H,W] (the generator output).
    Returns (tmp, mix_weight) exactly as function.py:138-144 builds them."""
    mix_weight = F.softmax(logits, dim=1)
    return mix_from_weights(inputs, mix_weight), mix_weight


def mix_from_weights(inputs, mix_weight):
    tmp = inputs[0] * mix_weight[:, 0, ...].unsqueeze(dim=1)
    for k in range(1, len(inputs)):
        tmp += inputs[k] * mix_weight[:, k].unsqueeze(dim=1)
    return tmp


def mix_backward(inputs, logits, grad_out, through_softmax=True):
    """Autograd reference: d(sum(tmp*grad_out))/d(logits or weights)."""
    z = logits.clone().requires_grad_(True)
    w = F.softmax(z, dim=1) if through_softmax else z
    tmp = mix_from_weights([x.clone() for x in inputs], w)
    tmp.backward(grad_out)
    return z.grad
