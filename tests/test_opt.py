"""nn/opt.py: learning-rate schedules (TF's closed forms), optimiser factory, the train op
with global step and moving averages (reference: nn/utils.py:77-150, nn/opt.py:89-166)."""
import math

import numpy as np
import pytest
import torch

from tensoralloy_b200.nn.opt import OptParameters, TrainOp, get_learning_rate, get_optimizer


def test_learning_rate_schedules():
    assert get_learning_rate(500, 0.01) == 0.01
    kw = dict(learning_rate=0.01, decay_rate=0.9, decay_steps=1000)
    assert abs(get_learning_rate(2500, decay_function='exponential', **kw)
               - 0.01 * 0.9 ** 2.5) < 1e-18
    assert abs(get_learning_rate(2500, decay_function='exponential', staircase=True, **kw)
               - 0.01 * 0.9 ** 2) < 1e-18
    assert abs(get_learning_rate(2500, decay_function='inverse_time', **kw)
               - 0.01 / (1 + 0.9 * 2.5)) < 1e-18
    assert abs(get_learning_rate(2500, decay_function='natural_exp', **kw)
               - 0.01 * math.exp(-0.9 * 2.5)) < 1e-18
    with pytest.raises(ValueError, match="not supported"):
        get_learning_rate(1, decay_function='cosine')


def test_optimizer_factory_defaults():
    p = [torch.zeros(3, requires_grad=True)]
    adam = get_optimizer(p, 0.01, 'adam', beta1=0.8)
    assert isinstance(adam, torch.optim.Adam) and adam.defaults['betas'] == (0.8, 0.999)
    assert adam.defaults['eps'] == 1e-8
    from tensoralloy_b200.nn.opt import DecoupledAdam
    adamw = get_optimizer(p, 0.01, 'adamw')
    assert isinstance(adamw, DecoupledAdam) and adamw.defaults['decay'] == 1e-4
    nadam = get_optimizer(p, 0.01, 'Nadam')
    assert isinstance(nadam, DecoupledAdam) and nadam.defaults['nesterov'] is True
    assert get_optimizer(p, 0.01, 'adadelta').defaults['rho'] == 0.95
    rms = get_optimizer(p, 0.01, 'rmsprop', momentum=0.5)
    from tensoralloy_b200.nn.opt import TfRMSprop
    assert isinstance(rms, TfRMSprop)
    assert rms.defaults['decay'] == 0.9 and rms.defaults['momentum'] == 0.5
    sgd = get_optimizer(p, 0.01, 'sgd')
    assert sgd.defaults['momentum'] == 0.9 and sgd.defaults['nesterov'] is True
    with pytest.raises(ValueError, match="Supported SGD optimizers"):
        get_optimizer(p, 0.01, 'lbfgs')


def test_train_op_step_learning_rate_and_moving_average():
    w = torch.tensor([1.0, -2.0], dtype=torch.float64, requires_grad=True)
    op = TrainOp([w], OptParameters(method='adam', learning_rate=0.1,
                                    decay_function='exponential', decay_rate=0.5,
                                    decay_steps=2))
    w0 = w.detach().clone()
    w.grad = torch.tensor([0.5, -0.25], dtype=torch.float64)
    lr = op.step()
    assert lr == 0.1 and op.global_step == 1
    # first Adam step: m_hat / (sqrt(v_hat) + eps) = sign(g) up to eps
    expect = w0 - 0.1 * torch.sign(w.grad)
    assert torch.allclose(w.detach(), expect, atol=1e-7)
    assert torch.allclose(op.shadow[0], 0.999 * w0 + 0.001 * w.detach(), atol=1e-15)
    lr2 = op.step()
    assert abs(lr2 - 0.1 * 0.5 ** 0.5) < 1e-15 and op.global_step == 2
    ema = op.ema_values()[0]
    op.swap_in_ema()
    assert torch.equal(w.detach(), ema)


def test_train_op_drives_a_trainer_like_object_to_the_minimum():
    """`trainer.train_step(optimizer)` only needs `.step()`: minimise a quadratic."""
    w = torch.tensor([3.0, -1.0], dtype=torch.float64, requires_grad=True)
    target = torch.tensor([0.5, 2.0], dtype=torch.float64)
    op = TrainOp([w], OptParameters(method='sgd', learning_rate=0.05,
                                    additional_kwargs=dict(momentum=0.5)))
    for _ in range(300):
        op.zero_grad()
        loss = ((w - target) ** 2).sum()
        loss.backward()
        op.step()
    assert torch.allclose(w.detach(), target, atol=1e-6) and op.global_step == 300


def test_tf_style_adamw_and_nadam_steps():
    """tf.contrib AdamWOptimizer: var -= decay * var, NOT scaled by the learning rate, then the
    TF Adam step (eps outside the bias correction); tf.contrib NadamOptimizer: the same Adam
    with the Nesterov numerator beta1 m + (1 - beta1) g and a constant beta1."""
    from tensoralloy_b200.nn.opt import DecoupledAdam
    x0, g = 2.0, 0.5
    lr, b1, b2, eps, decay = 0.01, 0.9, 0.999, 1e-8, 1e-2
    p = torch.tensor([x0], dtype=torch.float64, requires_grad=True)
    opt = DecoupledAdam([p], lr=lr, betas=(b1, b2), eps=eps, decay=decay)
    x, m, v = x0, 0.0, 0.0
    for t in range(1, 4):
        p.grad = torch.tensor([g], dtype=torch.float64)
        opt.step()
        x *= 1.0 - decay
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        x -= lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t) * m / (v ** 0.5 + eps)
        assert abs(p.item() - x) < 1e-14
    q = torch.tensor([x0], dtype=torch.float64, requires_grad=True)
    opt = DecoupledAdam([q], lr=lr, betas=(b1, b2), eps=eps, nesterov=True)
    x, m, v = x0, 0.0, 0.0
    for t in range(1, 4):
        q.grad = torch.tensor([g], dtype=torch.float64)
        opt.step()
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        x -= lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t) * (b1 * m + (1 - b1) * g) / (v ** 0.5 + eps)
        assert abs(q.item() - x) < 1e-14


def test_zero_grad_keeps_the_gradient_tensors():
    """A CUDA-graph training step replays into the gradient tensors of its capture:
    TrainOp.zero_grad must not detach them."""
    from tensoralloy_b200.nn.opt import TrainOp
    p = torch.zeros(3, requires_grad=True)
    op = TrainOp([p])
    p.grad = torch.ones(3)
    g = p.grad
    op.zero_grad()
    assert p.grad is g and float(g.abs().sum()) == 0.0


def test_rmsprop_follows_tf_arithmetic_under_a_decaying_learning_rate():
    """tf.train.RMSPropOptimizer (ApplyRMSProp): ms starts at 1, epsilon under the root, the
    learning rate INSIDE the momentum accumulator -- hand-rolled recurrence over 5 steps with a
    learning rate that halves every step."""
    w = torch.tensor([0.5, -1.5], dtype=torch.float64, requires_grad=True)
    op = TrainOp([w], OptParameters(method='rmsprop', learning_rate=0.1,
                                    decay_function='exponential', decay_rate=0.5, decay_steps=1,
                                    additional_kwargs={'decay': 0.8, 'momentum': 0.6}))
    x = w.detach().clone().numpy()
    ms, mom = np.ones(2), np.zeros(2)
    for k in range(5):
        g = 2.0 * x + np.array([0.3, -0.1]) * (k + 1)
        w.grad = torch.tensor(g)
        op.step()
        lr = 0.1 * 0.5 ** k
        ms = 0.8 * ms + 0.2 * g * g
        mom = 0.6 * mom + lr * g / np.sqrt(ms + 1e-10)
        x = x - mom
        assert np.allclose(w.detach().numpy(), x, rtol=0, atol=1e-15)
