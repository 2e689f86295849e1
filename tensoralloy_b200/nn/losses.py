"""
Loss arithmetic of the reference (tensoralloy/nn/losses.py), restated on torch
tensors for the training step:
  RMSE    = sqrt(mean((x - y)^2) + eps)                  losses.py:69-95  (eps = dtype eps)
  rRMSE   = mean(|x - y|_2 / |x|_2) over the rows        losses.py:53-66
  logcosh = mean(d + softplus(-2 d) - log 2), d = x - y  losses.py:44-50, 98-121
  ylogy   = mean(x (log max(x, 1e-12) - log max(y, 1e-12))^2)   losses.py:124-153
  energy: optionally per atom                            losses.py:204-282 (:250-253)
  forces: over the real atoms only                       losses.py:285-391
  stress: on the Voigt 6-vectors                         losses.py:394-456
  pressure: on the scalar total pressure (no rRMSE)      losses.py:459-504
  static or dynamic (linear / log-scaled in the global step) loss weights   losses.py:171-201
  total  = sum of the enabled weighted terms             nn/basic.py:626
`labels` come first, as in the reference.
"""
import math

import torch

from tensoralloy_b200.precision import get_float_dtype

METHODS = ('rmse', 'rrmse', 'logcosh', 'ylogy')


def _weights(sample_weight, like, normalized_weight):
    """losses.py:79-87: optional per-sample weights; `normalized_weight` divides them by their
    sum (and by the trailing extent of `like` for rank > 1), so that the weighted SUM below
    replaces the mean."""
    w = sample_weight
    if normalized_weight:
        w = w / torch.sum(w)
        if w.dim() > 1:
            w = w / float(like[0].numel())
    return w


def rmse(x, y, eps=None, sample_weight=None, normalized_weight=False):
    if eps is None:
        eps = get_float_dtype().eps
    if sample_weight is not None:
        mse = torch.sum((x - y) ** 2 * _weights(sample_weight, x, normalized_weight))
    else:
        mse = torch.mean((x - y) ** 2)
    return torch.sqrt(mse + eps)


def mae(x, y):
    """The MAE every loss of the reference reports beside its value."""
    return torch.mean(torch.abs(x - y))


def relative_rmse(labels, predictions):
    if labels.dim() == 1:
        labels, predictions = labels.reshape(-1, 1), predictions.reshape(-1, 1)
    upper = torch.linalg.norm(labels - predictions, dim=1)
    lower = torch.linalg.norm(labels, dim=1)
    return torch.mean(upper / lower)


def logcosh(labels, predictions, sample_weight=None, normalized_weight=False):
    d = labels - predictions
    v = d + torch.nn.functional.softplus(-2.0 * d) - math.log(2.0)
    if sample_weight is not None:
        return torch.sum(v * _weights(sample_weight, labels, normalized_weight))
    return torch.mean(v)


def ylogy(labels, predictions, sample_weight=None, normalized_weight=False):
    logx = torch.log(torch.clamp(labels, min=1e-12))
    logy = torch.log(torch.clamp(predictions, min=1e-12))
    v = (logx - logy) ** 2 * labels
    if sample_weight is not None:
        return torch.sum(v * _weights(sample_weight, labels, normalized_weight))
    return torch.mean(v)


def _raw(method, labels, predictions, allowed, sample_weight=None, normalized_weight=False):
    if method not in METHODS:
        raise KeyError(method)                    # LossMethod[options.method]
    if method not in allowed:
        raise ValueError(f"loss method '{method}' is not available for this property")
    if method == 'rrmse':                         # losses.py:53-66 takes no weights
        return relative_rmse(labels, predictions)
    return {'rmse': rmse, 'logcosh': logcosh, 'ylogy': ylogy}[method](
        labels, predictions, sample_weight=sample_weight, normalized_weight=normalized_weight)


def dynamic_weight(weight, global_step=0, max_train_steps=None, logscale=False):
    """losses.py:171-201: a float is a constant weight; a pair (w0, w1) moves from w0 to w1
    over `max_train_steps`, linearly or linearly in log10."""
    if isinstance(weight, (int, float)):
        return float(weight)
    w0, w1 = weight
    if max_train_steps is None:
        raise ValueError("a dynamic loss weight needs max_train_steps")
    if logscale:
        l0, l1 = math.log10(w0), math.log10(w1)
        return 10.0 ** (l0 + (l1 - l0) / max_train_steps * global_step)
    return w0 + (w1 - w0) / max_train_steps * global_step


def energy_loss(labels, predictions, n_atoms, per_atom_loss=True, weight=1.0, method='rmse',
                sample_weight=None, normalized_weight=False):
    """losses.py:204-282; `sample_weight` [batch] (e.g. `adaptive_sample_weight`)."""
    if per_atom_loss:
        n = n_atoms.to(labels.dtype)
        labels, predictions = labels / n, predictions / n
    return weight * _raw(method, labels, predictions, METHODS, sample_weight,
                         normalized_weight)


def forces_loss(labels, predictions, weight=1.0, method='rmse', sample_weight=None, sid=None,
                normalized_weight=True):
    """losses.py:285-391 (`_absolute_forces_loss`: rmse or logcosh; the reference asserts
    against rrmse).  labels / predictions: [total real atoms, 3] (padding already removed).
    `sample_weight` [batch] with `sid` [total atoms] = structure of every atom: each atom
    carries its structure's weight; normalised by (sum of the atom weights x 3), the weighted
    sum over atoms and components replaces the mean (losses.py:308-322)."""
    if method == 'rrmse':
        raise ValueError("loss method 'rrmse' is not available for the forces "
                         "(losses.py:297)")
    w = None
    if sample_weight is not None:
        w = sample_weight[sid] if sid is not None else sample_weight
        if normalized_weight:
            w = w / (torch.sum(w) * 3.0)
        w = w.unsqueeze(1)
    return weight * _raw(method, labels, predictions, ('rmse', 'logcosh'), w, False)


def stress_loss(labels, predictions, weight=1.0, method='rmse'):
    """labels / predictions: [batch, 6] Voigt, eV/A^3."""
    return weight * _raw(method, labels, predictions, ('rmse', 'rrmse', 'logcosh'))


def pressure_loss(labels, predictions, weight=1.0, method='rmse'):
    """labels / predictions: [batch] total pressure (losses.py:459-504: rmse or logcosh)."""
    if labels.dim() != 1 or predictions.dim() != 1:
        raise ValueError("pressure loss: rank-1 tensors expected")
    return weight * _raw(method, labels, predictions, ('rmse', 'logcosh'))


def l2_regularization_loss(variables, l2_weight, weight=0.0, decayed=True, decay_rate=0.99,
                           decay_steps=1000, global_step=0):
    """losses.py:507-550 with the layers' regulariser (convolutional.py:207-210, 262-265:
    `l2_regularizer(l2_weight)` on every kernel and bias = l2_weight * sum(w^2) / 2): the sum
    over `variables`, times the option weight, which decays exponentially in the global step
    (`L2LossOptions`, dataclasses.py:158-166).  Returns None for weight 0, as the reference."""
    if weight == 0.0:
        return None
    total = sum(0.5 * l2_weight * torch.sum(v * v) for v in variables)
    if decayed:
        weight = weight * decay_rate ** (float(global_step) / float(decay_steps))
    return weight * total


def adaptive_sample_weight(true_forces, sid, n_struct, metric, method, *args):
    """losses.py:553-586: a per-structure sample weight that shrinks for structures with very
    large forces.  true_forces [total atoms, 3] with `sid` [total atoms] = structure of each
    atom (the unpadded form of the reference's [batch, N + 1, 3]).
      metric 'norm'  f = sqrt(sum |F|^2 / n_atoms)      'fmax'  f = max |F_ia|
      method 'sigmoid', args = (slope, center, wmax, wmin):  wmin + wmax sigmoid(slope (center - f))"""
    if metric == 'norm':
        sq = torch.zeros(n_struct, dtype=true_forces.dtype, device=true_forces.device)
        sq = sq.index_add(0, sid, torch.sum(true_forces * true_forces, dim=1))
        n = torch.bincount(sid, minlength=n_struct).to(true_forces.dtype)
        f = torch.sqrt(torch.where(n > 0, sq / torch.clamp(n, min=1.0), torch.zeros_like(sq)))
    elif metric == 'fmax':
        f = torch.zeros(n_struct, dtype=true_forces.dtype, device=true_forces.device)
        f = f.scatter_reduce(0, sid, true_forces.abs().amax(dim=1), reduce='amax',
                             include_self=True)
    else:
        raise ValueError("Only the norm and fmax metric is implemented")
    if method != 'sigmoid':
        raise ValueError("Only the sigmoid method is implemented")
    slope, center, wmax, wmin = [float(a) for a in args]
    return torch.sigmoid(slope * (center - f)) * wmax + wmin
