"""
Loss arithmetic of the reference (tensoralloy/nn/losses.py), restated on torch
tensors for the training step:
  RMSE = sqrt(mean((x - y)^2) + eps)          losses.py:69-95  (eps = dtype eps)
  energy: optionally per atom                 losses.py:204-282 (:250-253)
  forces: over the real atoms only            losses.py:285-332
  stress: on the Voigt 6-vectors              losses.py:394-456
  total  = sum of the enabled weighted terms  nn/basic.py:626
"""
import torch

from tensoralloy_b200.precision import get_float_dtype


def rmse(x, y, eps=None):
    if eps is None:
        eps = get_float_dtype().eps
    return torch.sqrt(torch.mean((x - y) ** 2) + eps)


def energy_loss(labels, predictions, n_atoms, per_atom_loss=True, weight=1.0):
    if per_atom_loss:
        n = n_atoms.to(labels.dtype)
        return weight * rmse(labels / n, predictions / n)
    return weight * rmse(labels, predictions)


def forces_loss(labels, predictions, weight=1.0):
    """labels / predictions: [total real atoms, 3] (padding already removed)."""
    return weight * rmse(labels, predictions)


def stress_loss(labels, predictions, weight=1.0):
    """labels / predictions: [batch, 6] Voigt, eV/A^3."""
    return weight * rmse(labels, predictions)
