"""Fused temperature-dependent heads (tab_td_eval) against their torch formulation (three
chains of GEMMs + autograd for dF/dG) on synthetic descriptors (dev tool, run under gpurun):
python tools/td_heads_bench.py [n_atoms] -> JSON lines (float64, float32)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
from tensoralloy_b200.nn.atomic import SymmetryFunction                          # noqa: E402
from tensoralloy_b200.nn.atomic.finite_temperature import BeNN                   # noqa: E402
from tensoralloy_b200.precision import precision_scope, get_float_dtype          # noqa: E402
from tensoralloy_b200.transformer import UniversalTransformer                    # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for prec in ('high', 'medium'):
    with precision_scope(prec):
        nn = BeNN(['Be'], SymmetryFunction(['Be']), hidden_sizes=[64, 32],
                  finite_temperature=dict(activation='softplus', layers=[128, 128]))
        nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=True))
        nn.initialize_variables(seed=3)
        dim = nn._dim()
        rng = np.random.default_rng(1)
        G = torch.as_tensor(rng.random((n, dim)), device='cuda')
        T = torch.full((n,), 0.17, dtype=torch.float64, device='cuda')
        types = torch.zeros(n, dtype=torch.int32, device='cuda')
        dt = get_float_dtype()
        heads = nn._device_heads()
        U, S, F, dfdg = heads.eval(types, G, T, dt.tab_precision)
        tdtype = torch.float64 if prec == 'high' else torch.float32
        hd = nn._torch_heads(tdtype)['Be']

        def torch_path():
            x = G.to(tdtype).requires_grad_(True)
            H = nn._net(hd['H'], x if hd['xlo'] is None else
                        (hd['xhi'] - x) / (hd['xhi'] - hd['xlo']))
            t = T.to(tdtype)
            Ht = torch.cat([H, t[:, None]], dim=1)
            u = nn._net(hd['U'], Ht)[:, 0]
            s = nn._entropy(hd, Ht, t)
            f = u - t * s
            g = torch.autograd.grad(f.sum(), x)[0]
            return u, s, f, g

        u, s, f, g = torch_path()
        err = {"U": float((U - u).abs().max()), "S": float((S - s).abs().max()),
               "F": float((F - f).abs().max()), "dFdG": float((dfdg - g).abs().max()),
               "scale_dFdG": float(g.abs().max())}
        ms_k = timed(lambda: heads.eval(types, G, T, dt.tab_precision))
        ms_t = timed(torch_path)
        print(json.dumps({"precision": prec, "atoms": n, "dim": dim, "layers_H": [128, 128],
                          "hidden_S_U": [64, 32], "tab_td_eval_ms": ms_k,
                          "torch_gemm_autograd_ms": ms_t, "max_abs_diff": err}), flush=True)
