"""The reference's own known-answer tests for the loss arithmetic and the cutoff functions
(tensoralloy/nn/tests/test_losses.py:23-230, test_cutoff.py:23-95), re-run on the torch
restatements used by the training step (nn/losses.py), on the oracle's cutoffs and on the
torch cutoff of the filter-network path.  The "simple" formulas are the ones the reference
tests state."""
import numpy as np
import torch

from oracle import atomic as oat
from tensoralloy_b200.nn import losses
from tensoralloy_b200.nn.atomic.grap_nn import _cutoff
from tensoralloy_b200.precision import precision_scope


def test_energy_loss_per_atom_rmse():
    # test_losses.py:23-51 ('medium' precision, delta 1e-8 on the float32 value there;
    # the eps under the square root is the dtype eps, losses.py:69-95)
    rng = np.random.default_rng(0)
    x, y = rng.uniform(0.0, 3.0, 6), rng.uniform(0.0, 3.0, 6)
    n = rng.integers(1, 5, 6).astype(float)
    y_rmse = np.sqrt(np.mean((x / n - y / n) ** 2))
    with precision_scope('high'):
        got = losses.energy_loss(torch.tensor(x), torch.tensor(y), torch.tensor(n)).item()
        assert abs(got - np.sqrt(y_rmse ** 2 + 1e-14)) < 1e-14
        assert abs(got - y_rmse) < 1e-8
        total = losses.energy_loss(torch.tensor(x), torch.tensor(y), torch.tensor(n),
                                   per_atom_loss=False, weight=2.0).item()
        assert abs(total - 2.0 * np.sqrt(np.mean((x - y) ** 2))) < 1e-8
    with precision_scope('medium'):
        got32 = losses.energy_loss(torch.tensor(x, dtype=torch.float32),
                                   torch.tensor(y, dtype=torch.float32),
                                   torch.tensor(n, dtype=torch.float32)).item()
        assert abs(got32 - y_rmse) < 1e-6


def test_forces_loss_over_real_atoms_with_weight():
    # test_losses.py:90-130: RMSE over the real atoms of every structure, weight 10
    rng = np.random.default_rng(1)
    n_atoms = [5, 6, 4, 8, 5, 3]
    x = [rng.uniform(0.0, 3.0, (n, 3)) for n in n_atoms]
    y = [rng.uniform(0.0, 3.0, (n, 3)) for n in n_atoms]
    sq = np.concatenate([(a - b) ** 2 for a, b in zip(x, y)])
    y_rmse = np.sqrt(np.mean(sq))
    with precision_scope('high'):
        got = losses.forces_loss(torch.tensor(np.concatenate(x)),
                                 torch.tensor(np.concatenate(y)), weight=10.0).item()
    assert abs(got - 10.0 * y_rmse) < 1e-8


def test_stress_loss_voigt_rmse():
    # test_losses.py:187-204
    rng = np.random.default_rng(2)
    x, y = rng.uniform(0.1, 3.0, (8, 6)), rng.uniform(0.1, 3.0, (8, 6))
    with precision_scope('high'):
        got = losses.stress_loss(torch.tensor(x), torch.tensor(y)).item()
    assert abs(got - np.sqrt(np.mean((x - y) ** 2))) < 1e-8


def test_cutoff_functions_match_the_simple_forms():
    # test_cutoff.py:23-95: r = 1 .. 10 in 91 steps, rc = 6, gamma = 5
    rc, gamma = 6.0, 5.0
    r = np.linspace(1.0, 10.0, num=91, endpoint=True)
    cos_simple = np.where(r <= rc, 0.5 * (np.cos(np.pi * r / rc) + 1.0), 0.0)
    d = r / rc
    poly_simple = np.where(r <= rc, 1.0 + gamma * d ** (gamma + 1.0) - (gamma + 1.0) * d ** gamma,
                           0.0)
    t = torch.tensor(r)
    assert np.abs(oat.cosine_cutoff(t, rc).numpy() - cos_simple).max() < 1e-14
    assert np.abs(oat.polynomial_cutoff(t, rc).numpy() - poly_simple).max() < 1e-14
    assert np.abs(_cutoff('cosine', t, rc).numpy() - cos_simple).max() < 1e-14
    assert np.abs(_cutoff('polynomial', t, rc).numpy() - poly_simple).max() < 1e-14
    t32 = torch.tensor(r, dtype=torch.float32)
    assert np.abs(_cutoff('cosine', t32, rc).numpy() - cos_simple).max() < 1e-7   # the reference's delta


def test_other_loss_methods_and_dynamic_weights():
    """losses.py:44-66 (logcosh = keras form, relative RMSE), :124-153 (ylogy), :171-201
    (static / dynamic weights), :459-504 (pressure: rmse or logcosh only)."""
    import pytest
    rng = np.random.default_rng(3)
    x, y = rng.uniform(0.1, 3.0, (7, 3)), rng.uniform(0.1, 3.0, (7, 3))
    tx, ty = torch.tensor(x), torch.tensor(y)
    with precision_scope('high'):
        assert abs(losses.forces_loss(tx, ty, method='logcosh').item()
                   - np.mean(np.log(np.cosh(x - y)))) < 1e-12
        with pytest.raises(ValueError, match="not available"):
            losses.forces_loss(tx, ty, method='rrmse')      # losses.py:297 asserts against it
        assert abs(losses.stress_loss(tx, ty, method='rrmse').item()
                   - np.mean(np.linalg.norm(x - y, axis=1) / np.linalg.norm(x, axis=1))) < 1e-12
        e, p, n = tx[:, 0], ty[:, 0], torch.tensor(rng.integers(1, 5, 7).astype(float))
        xe, ye = (e / n).numpy(), (p / n).numpy()
        assert abs(losses.energy_loss(e, p, n, method='ylogy').item()
                   - np.mean(xe * (np.log(xe) - np.log(ye)) ** 2)) < 1e-12
        assert abs(losses.energy_loss(e, p, n, method='rrmse').item()
                   - np.mean(np.abs(xe - ye) / np.abs(xe))) < 1e-12
        assert abs(losses.mae(tx, ty).item() - np.mean(np.abs(x - y))) < 1e-15
        assert abs(losses.pressure_loss(e, p, weight=2.0).item()
                   - 2.0 * np.sqrt(np.mean((x[:, 0] - y[:, 0]) ** 2) + 1e-14)) < 1e-12
        with pytest.raises(ValueError, match="not available"):
            losses.pressure_loss(e, p, method='rrmse')
        with pytest.raises(ValueError, match="not available"):
            losses.stress_loss(tx, ty, method='ylogy')
        with pytest.raises(KeyError):
            losses.energy_loss(e, p, n, method='huber')
    assert losses.dynamic_weight(0.5) == 0.5
    assert abs(losses.dynamic_weight((1.0, 3.0), 250, 1000) - 1.5) < 1e-15
    assert abs(losses.dynamic_weight((1.0, 100.0), 500, 1000, logscale=True) - 10.0) < 1e-12
    with pytest.raises(ValueError, match="max_train_steps"):
        losses.dynamic_weight((1.0, 2.0), 3)


def test_l2_regularization_loss():
    # losses.py:507-550: weight 0 -> None; decayed weight = w * rate^(step / steps)
    ws = [torch.tensor([[1.0, -2.0], [0.5, 0.0]]), torch.tensor([3.0])]
    assert losses.l2_regularization_loss(ws, 0.01) is None
    raw = 0.5 * 0.01 * (1 + 4 + 0.25 + 9)
    got = losses.l2_regularization_loss(ws, 0.01, weight=0.2, decayed=False).item()
    assert abs(got - 0.2 * raw) < 1e-9
    got = losses.l2_regularization_loss(ws, 0.01, weight=0.2, decay_rate=0.5, decay_steps=100,
                                        global_step=200).item()
    assert abs(got - 0.2 * 0.25 * raw) < 1e-9


def test_adaptive_sample_weight():
    # losses.py:553-586
    import pytest
    rng = np.random.default_rng(5)
    n_atoms = [4, 2, 5]
    F = [rng.normal(scale=s, size=(n, 3)) for n, s in zip(n_atoms, (0.1, 3.0, 1.0))]
    sid = torch.tensor(np.repeat(np.arange(3), n_atoms))
    t = torch.tensor(np.concatenate(F))
    args = (2.0, 1.0, 0.9, 0.1)
    w = losses.adaptive_sample_weight(t, sid, 3, 'norm', 'sigmoid', *args).numpy()
    f = np.array([np.sqrt((x ** 2).sum() / len(x)) for x in F])
    assert np.allclose(w, 0.1 + 0.9 / (1 + np.exp(-2.0 * (1.0 - f))), atol=1e-14)
    w = losses.adaptive_sample_weight(t, sid, 3, 'fmax', 'sigmoid', *args).numpy()
    f = np.array([np.abs(x).max() for x in F])
    assert np.allclose(w, 0.1 + 0.9 / (1 + np.exp(-2.0 * (1.0 - f))), atol=1e-14)
    assert w[1] < w[2] < w[0]                          # large forces -> small weight
    with pytest.raises(ValueError, match="norm and fmax"):
        losses.adaptive_sample_weight(t, sid, 3, 'mean', 'sigmoid', *args)
    with pytest.raises(ValueError, match="sigmoid"):
        losses.adaptive_sample_weight(t, sid, 3, 'norm', 'linear', *args)


def test_sample_weights():
    """losses.py:69-153 (energy: weighted SUM of squares, weights optionally normalised by their
    sum) and :285-332 (forces: every atom carries its structure's weight, normalised by
    (sum of atom weights) x 3) -- with `adaptive_sample_weight` (:553-586) as the source."""
    rng = np.random.default_rng(9)
    with precision_scope('high'):
        e, p = torch.tensor(rng.normal(size=5)), torch.tensor(rng.normal(size=5))
        n = torch.tensor([3.0, 4.0, 2.0, 5.0, 3.0])
        w = torch.tensor(rng.uniform(0.2, 1.0, 5))
        d = (e / n - p / n).numpy()
        wn = (w / w.sum()).numpy()
        assert abs(losses.energy_loss(e, p, n, sample_weight=w, normalized_weight=True).item()
                   - np.sqrt(np.sum(d * d * wn) + 1e-14)) < 1e-14
        assert abs(losses.energy_loss(e, p, n, sample_weight=w).item()
                   - np.sqrt(np.sum(d * d * w.numpy()) + 1e-14)) < 1e-14
        assert abs(losses.energy_loss(e, p, n, method='logcosh', sample_weight=w,
                                      normalized_weight=True).item()
                   - np.sum(np.log(np.cosh(d)) * wn)) < 1e-13
        sid = torch.tensor([0] * 3 + [1] * 4 + [2] * 2 + [3] * 5 + [4] * 3)
        f, g = torch.tensor(rng.normal(size=(17, 3))), torch.tensor(rng.normal(size=(17, 3)))
        wa = w[sid].numpy()
        wa = wa / (wa.sum() * 3.0)
        df = (f - g).numpy()
        assert abs(losses.forces_loss(f, g, sample_weight=w, sid=sid).item()
                   - np.sqrt(np.sum(df * df * wa[:, None]) + 1e-14)) < 1e-14
        # equal weights, normalised = the plain mean
        one = torch.ones(5, dtype=torch.float64)
        assert abs(losses.forces_loss(f, g, sample_weight=one, sid=sid).item()
                   - losses.forces_loss(f, g).item()) < 1e-14
        aw = losses.adaptive_sample_weight(f, sid, 5, 'fmax', 'sigmoid', 2.0, 1.5, 1.0, 0.1)
        assert aw.shape == (5,) and float(aw.min()) > 0.1
