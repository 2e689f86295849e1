"""
The optimiser side of the training step -- mirror of tensoralloy/nn/opt.py:89-166
(`get_train_op`), nn/utils.py:77-150 (`get_learning_rate`, `get_optimizer`) and
`OptParameters` (nn/dataclasses.py:298-308) on torch optimisers, for the trainers of this
package (`trainer.train_step(optimizer)` calls `optimizer.step()` after the gradients of the
step -- summed over the loss terms, opt.py:134-146, and averaged over the ranks -- are in
place).

  learning rate   constant, or TF's decay schedules evaluated at the global step BEFORE the
                  update: exponential lr * rate^p, inverse_time lr / (1 + rate p),
                  natural_exp lr * exp(-rate p), p = step / decay_steps (floored with
                  `staircase`)
  method          adam (beta1), adamw (decay), nadam (beta1), adadelta (rho), rmsprop (decay,
                  momentum), sgd (momentum 0.9, Nesterov) with TF's default epsilons
  moving average  shadow = d * shadow + (1 - d) * variable after every update, d = 0.999
                  (utils.py:416); `use_ema_variables` of the export path = `ema_values()`

`adam` runs torch.optim.Adam: it divides by sqrt(v_hat) + eps where TF divides by sqrt(v) + eps
with the bias correction folded into the step size; the two agree to O(eps).  `adamw`, `nadam`
and `rmsprop` are small optimisers of this module in TF's own arithmetic (`DecoupledAdam`,
`TfRMSprop`), because the torch classes of the same name differ materially (weight decay scaled
by lr, Dozat's momentum schedule, lr outside the momentum accumulator).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch


@dataclass
class OptParameters:
    """nn/dataclasses.py:298-308."""
    method: str = 'adam'
    learning_rate: float = 0.01
    decay_function: Optional[str] = None
    decay_rate: float = 0.99
    decay_steps: int = 1000
    staircase: bool = False
    additional_kwargs: Optional[Dict] = field(default=None)


def get_learning_rate(global_step, learning_rate=0.001, decay_function=None, decay_rate=0.99,
                      decay_steps=1000, staircase=False):
    """nn/utils.py:77-103 with tf.train.{exponential,inverse_time,natural_exp}_decay."""
    if decay_function is None:
        return float(learning_rate)
    p = float(global_step) / float(decay_steps)
    if staircase:
        p = math.floor(p)
    if decay_function == 'exponential':
        return learning_rate * decay_rate ** p
    if decay_function == 'inverse_time':
        return learning_rate / (1.0 + decay_rate * p)
    if decay_function == 'natural_exp':
        return learning_rate * math.exp(-decay_rate * p)
    raise ValueError("'{}' is not supported!".format(decay_function))


def get_optimizer(params, learning_rate, method='adam', **kwargs):
    """nn/utils.py:106-150 -> torch.optim."""
    params = list(params)
    m = method.lower()
    if m == 'adam':
        return torch.optim.Adam(params, lr=learning_rate,
                                betas=(kwargs.get('beta1', 0.9), 0.999), eps=1e-8)
    if m == 'adamw':
        # tf.contrib.opt.AdamWOptimizer (nn/utils.py:126-129) decays the variables by
        # `weight_decay * var` per step, NOT scaled by the learning rate; torch's AdamW applies
        # lr * weight_decay.  DecoupledAdam below keeps the reference's form.
        return DecoupledAdam(params, lr=learning_rate, eps=1e-8,
                             betas=(kwargs.get('beta1', 0.9), 0.999),
                             decay=kwargs.get('decay', 1e-4))
    if m == 'nadam':
        # tf.contrib.opt.NadamOptimizer = Adam with a Nesterov step and a CONSTANT beta1
        # (torch.optim.NAdam follows Dozat's momentum schedule mu_t, momentum_decay 4e-3)
        return DecoupledAdam(params, lr=learning_rate, eps=1e-8,
                             betas=(kwargs.get('beta1', 0.9), 0.999), decay=0.0,
                             nesterov=True)
    if m == 'adadelta':
        return torch.optim.Adadelta(params, lr=learning_rate, rho=kwargs.get('rho', 0.95),
                                    eps=1e-8)
    if m == 'rmsprop':
        # TF keeps the learning rate INSIDE the momentum accumulator and epsilon under the
        # square root (torch.optim.RMSprop: outside both): the two differ once lr decays
        return TfRMSprop(params, lr=learning_rate, decay=kwargs.get('decay', 0.9),
                         momentum=kwargs.get('momentum', 0.0), eps=1e-10)
    if m == 'sgd':
        return torch.optim.SGD(params, lr=learning_rate,
                               momentum=kwargs.get('momentum', 0.9),
                               nesterov=kwargs.get('use_nesterov', True))
    raise ValueError("Supported SGD optimizers: adam, nadam, adadelta, rmsprop.")


class DecoupledAdam(torch.optim.Optimizer):
    """Adam in the arithmetic of TF 1.15 (`training_ops.cc` ApplyAdam):
        lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t)
        m <- beta1 m + (1 - beta1) g ;  v <- beta2 v + (1 - beta2) g^2
        var <- var - lr_t m / (sqrt(v) + eps)             (eps OUTSIDE the bias correction)
    `nesterov=True` is tf.contrib's NadamOptimizer (use_nesterov): the step uses
    beta1 m + (1 - beta1) g instead of m.  `decay > 0` is tf.contrib's AdamWOptimizer
    (DecoupledWeightDecayExtension): var <- var - decay * var before the Adam step,
    independent of the learning rate."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, decay=0.0, nesterov=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, decay=decay,
                                      nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['t'] = 0
                    st['m'] = torch.zeros_like(p)
                    st['v'] = torch.zeros_like(p)
                st['t'] += 1
                t, m, v, g = st['t'], st['m'], st['v'], p.grad
                if group['decay']:
                    p.mul_(1.0 - group['decay'])
                m.mul_(b1).add_(g, alpha=1.0 - b1)
                v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
                lr_t = group['lr'] * (1.0 - b2 ** t) ** 0.5 / (1.0 - b1 ** t)
                num = m.mul(b1).add_(g, alpha=1.0 - b1) if group['nesterov'] else m
                p.addcdiv_(num, v.sqrt().add_(group['eps']), value=-lr_t)


class TfRMSprop(torch.optim.Optimizer):
    """tf.train.RMSPropOptimizer (`training_ops.cc` ApplyRMSProp, centered = False):
        ms  <- decay ms + (1 - decay) g^2          (ms starts at ONE, as TF initialises it)
        mom <- momentum mom + lr g / sqrt(ms + eps)
        var <- var - mom"""

    def __init__(self, params, lr=1e-2, decay=0.9, momentum=0.0, eps=1e-10):
        super().__init__(params, dict(lr=lr, decay=decay, momentum=momentum, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['ms'] = torch.ones_like(p)
                    st['mom'] = torch.zeros_like(p)
                ms, mom, g = st['ms'], st['mom'], p.grad
                ms.mul_(group['decay']).addcmul_(g, g, value=1.0 - group['decay'])
                mom.mul_(group['momentum']).addcdiv_(g, (ms + group['eps']).sqrt_(),
                                                     value=group['lr'])
                p.sub_(mom)


class TrainOp:
    """`get_train_op` (opt.py:89-166) as an object: `step()` = apply the gradients held in
    `p.grad` with the learning rate of the current global step, advance the global step,
    update the moving averages.  Pass it to `trainer.train_step(...)`."""

    ema_decay = 0.999                      # Defaults.variable_moving_average_decay

    def __init__(self, params, opt_parameters: OptParameters = None):
        self.params = list(params)
        self.hp = opt_parameters or OptParameters()
        self.global_step = 0
        self.optimizer = get_optimizer(self.params, self.learning_rate(), self.hp.method,
                                       **(self.hp.additional_kwargs or {}))
        self.shadow = [p.detach().clone() for p in self.params]

    def learning_rate(self, step=None):
        hp = self.hp
        return get_learning_rate(self.global_step if step is None else step,
                                 hp.learning_rate, hp.decay_function, hp.decay_rate,
                                 hp.decay_steps, hp.staircase)

    def step(self):
        lr = self.learning_rate()
        for group in self.optimizer.param_groups:
            group['lr'] = lr
        self.optimizer.step()
        self.global_step += 1
        d = self.ema_decay
        with torch.no_grad():
            for s, p in zip(self.shadow, self.params):
                s.mul_(d).add_(p.detach(), alpha=1.0 - d)
        return lr

    def zero_grad(self):
        # keep the tensors: a CUDA-graph training step (trainer.enable_graph) replays into the
        # gradient tensors of its capture
        self.optimizer.zero_grad(set_to_none=False)

    def ema_values(self):
        """The moving averages, in the order of `params` (what the reference exports with
        `use_ema_variables=True`, basic.py:1017-1060)."""
        return [s.clone() for s in self.shadow]

    def swap_in_ema(self):
        """Overwrite the parameters with their moving averages (before `sync_to_model` /
        export)."""
        with torch.no_grad():
            for s, p in zip(self.shadow, self.params):
                p.copy_(s)
