"""
Reader of the reference's exported model files: frozen TensorFlow `GraphDef`
`.pb` written by `BasicNN.export` (tensoralloy/nn/basic.py:1017-1153).

The TF graph is NOT executed.  What is read:
  * `Transformer/params`  JSON -> UniversalTransformer(**params)   (basic.py:1075-1080)
  * `Metadata/ops`        JSON property -> tensor name             (basic.py:1088-1092)
  * `Metadata/{timestamp, precision | api, variational_energy,
     is_finite_temperature}`; the legacy layout of the shipped
     test_files/models/*.pb (no `api`, has `precision`) is accepted too
     
  * model parameters stored as `Const` nodes:
       EAM|ADP/Shared/<section>/<param>                      (potentials.py:171-200)
       Atomic/<El>/Conv1d{k}/{kernel,bias}, Output/{kernel,bias}, xlo, xhi
  * which potential implements each function, from the node-name scopes the
    reference emits (`EAM/Rho/NiNi/Zjw04/Rho/Ni/...`).
Symmetry-function hyper-parameters are only present as anonymous graph constants
(`.../G2/{tau}/eta`); they are recovered by name pattern, and a clear error is
raised if that fails.
"""
import json
import re
from collections import namedtuple

import numpy as np

LoadedModel = namedtuple('LoadedModel', ['nn', 'precision', 'api_version', 'timestamp',
                                         'predict_properties', 'ops'])

# scope name emitted by each reference potential class -> registry name
_SCOPE_TO_POTENTIAL = {
    'Zjw04': 'zjw04', 'Zjw04xc': 'zjw04xc', 'Zjw04uxc': 'zjw04uxc',
    'Zjw04xcp': 'zjw04xcp', 'Sutton': 'sutton90', 'AgraBe': 'Be/1',
    'RWGrimes': 'grimes', 'MishinH': 'mishinh', 'Mash11': 'msah11',
}
_DT_FLOAT, _DT_DOUBLE, _DT_INT32, _DT_STRING, _DT_INT64 = 1, 2, 3, 7, 9


def parse_graph_def(path):
    try:
        from tensorboard.compat.proto import graph_pb2
    except Exception as exc:   # pragma: no cover
        raise ImportError("reading .pb models needs the GraphDef protos shipped "
                          "with `tensorboard`") from exc
    g = graph_pb2.GraphDef()
    with open(path, 'rb') as fp:
        g.ParseFromString(fp.read())
    return g


def const_value(node):
    """numpy value (or bytes) of a Const node."""
    t = node.attr['value'].tensor
    shape = [d.size for d in t.tensor_shape.dim]
    if t.dtype == _DT_STRING:
        return t.string_val[0] if t.string_val else b''
    dt = {_DT_FLOAT: np.float32, _DT_DOUBLE: np.float64, _DT_INT32: np.int32,
          _DT_INT64: np.int64}.get(t.dtype)
    if dt is None:
        return None
    if t.tensor_content:
        arr = np.frombuffer(t.tensor_content, dtype=dt)
    else:
        vals = {_DT_FLOAT: t.float_val, _DT_DOUBLE: t.double_val,
                _DT_INT32: t.int_val, _DT_INT64: t.int64_val}[t.dtype]
        arr = np.asarray(list(vals), dtype=dt)
        n = int(np.prod(shape)) if shape else 1
        if arr.size == 1 and n > 1:
            arr = np.full(n, arr[0], dtype=dt)
    return arr.reshape(shape) if shape else (arr.reshape(()) if arr.size == 1 else arr)


def load_graph_model(path) -> LoadedModel:
    from tensoralloy_b200.transformer import UniversalTransformer
    g = parse_graph_def(path)
    consts = {n.name: n for n in g.node if n.op == 'Const'}
    names = [n.name for n in g.node]

    def meta(key, default=None):
        node = consts.get(f'Metadata/{key}')
        if node is None:
            return default
        v = const_value(node)
        return v.decode('utf-8') if isinstance(v, bytes) else v

    if 'Transformer/params' not in consts:
        raise Exception("Validated Ops cannot be found")     # calculator.py:161
    params = json.loads(const_value(consts['Transformer/params']).decode('utf-8'))
    params.pop('predict_properties', None)
    cls = params.pop('class')
    if cls != 'UniversalTransformer':
        raise ValueError(f"Unsupported transformer: {cls}")  # calculator.py:142
    clf = UniversalTransformer(**params)
    ops = json.loads(meta('ops', '{}'))
    ops = {k: v for k, v in ops.items() if v.endswith(':0')}
    if not ops:
        raise Exception("Validated Ops cannot be found")
    # precision: api >= 1.1 infers it from the op dtype (calculator.py:154-159);
    # legacy files carry Metadata/precision
    precision = meta('precision')
    if precision is None:
        precision = 'high'
        first = next(iter(ops.values())).split(':')[0]
        for n in g.node:
            if n.name == first:
                t = n.attr.get('T') or n.attr.get('dtype')
                if t is not None and t.type == _DT_FLOAT:
                    precision = 'medium'
                break
    scopes = {x.split('/')[0] for x in names}
    if 'EAM' in scopes or 'ADP' in scopes:
        nn = _build_eam(g, consts, names, clf, 'ADP' if 'ADP' in scopes else 'EAM', ops)
    elif 'Atomic' in scopes:
        nn = _build_atomic(g, consts, names, clf, ops)
    else:
        raise ValueError(f"no supported model scope in {sorted(scopes)}")
    nn.attach_transformer(clf)
    props = [p for p in ops.keys()]
    return LoadedModel(nn=nn, precision=precision, api_version=meta('api', '1.0'),
                       timestamp=meta('timestamp'), predict_properties=props, ops=ops)


def _export_properties(ops):
    known = ('energy', 'forces', 'stress', 'total_pressure', 'hessian', 'elastic')
    return [p for p in known if p in ops]


def _build_eam(g, consts, names, clf, scope, ops):
    from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN, EamFsNN
    # which potential implements which function
    pat = re.compile(rf'^{scope}/(Rho|Phi|Embed|Dipole|Quadrupole)/([A-Za-z]+)/([A-Za-z0-9]+)/')
    custom = {}
    for x in names:
        m = pat.match(x)
        if not m:
            continue
        fn, key, pscope = m.group(1).lower(), m.group(2), m.group(3)
        if pscope in _SCOPE_TO_POTENTIAL:
            custom.setdefault(key, {})[fn] = _SCOPE_TO_POTENTIAL[pscope]
        elif pscope.startswith('Conv'):
            custom.setdefault(key, {})[fn] = 'nn'
    elements = clf.elements
    # alloy: rho keyed by k-body term in the graph (`Rho/NiNi`) but by element in
    # `custom_potentials`; FS keeps the k-body term
    rho_terms = {k for k, v in custom.items() if 'rho' in v}
    is_fs = False
    alloy_custom = {}
    for key, fns in custom.items():
        for fn, pot in fns.items():
            if fn == 'rho':
                from tensoralloy_b200.utils import get_elements_from_kbody_term
                els = get_elements_from_kbody_term(key)
                alloy_custom.setdefault(els[-1], {})['rho'] = pot
            elif fn == 'embed':
                alloy_custom.setdefault(key, {})['embed'] = pot
            else:
                alloy_custom.setdefault(key, {})[fn] = pot
    del rho_terms
    cls = AdpNN if scope == 'ADP' else (EamFsNN if is_fs else EamAlloyNN)
    nn = cls(elements=elements, custom_potentials=alloy_custom,
             export_properties=_export_properties(ops))
    prefix = f'{scope}/Shared/'
    for name, node in consts.items():
        if name.startswith(prefix) and name.count('/') == 3:
            v = const_value(node)
            if v is not None and np.size(v) == 1:
                nn.set_variable(name, float(np.asarray(v).reshape(-1)[0]))
    return nn


def _build_atomic(g, consts, names, clf, ops):
    from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
    elements = clf.elements

    def grid(kind, key):
        vals = {}
        pat = re.compile(rf'/{kind}/(\d+)/{key}$')
        for name, node in consts.items():
            m = pat.search(name)
            if m:
                vals[int(m.group(1))] = float(np.asarray(const_value(node)).reshape(-1)[0])
        return [vals[k] for k in sorted(vals)]

    eta_t, omega_t = grid('G2', 'eta'), grid('G2', 'omega')
    if not eta_t:
        raise ValueError("symmetry-function hyper-parameters could not be recovered "
                         "from the graph constants; pass them explicitly")
    # ParameterGrid order: eta outer, omega inner
    n_omega = len(set(omega_t)) or 1
    eta = eta_t[::n_omega]
    omega = omega_t[:n_omega]
    kw = dict(eta=eta, omega=omega)
    if clf.angular:
        b, gm, z = grid('G4', 'beta'), grid('G4', 'gamma'), grid('G4', 'zeta')
        nz = len(dict.fromkeys(z))
        ng = len(dict.fromkeys(gm))
        kw.update(beta=b[::nz * ng], gamma=gm[:nz * ng:nz], zeta=z[:nz])
    kw['cutoff_function'] = 'polynomial' if any('PolyCutoff' in x for x in names) \
        else 'cosine'
    hidden = {}
    act = 'softplus'
    for x in names:
        m = re.match(r'^Atomic/[A-Za-z]+/Conv1d1/(Softplus|Tanh|Relu|LeakyRelu|Sigmoid|'
                     r'Softsign|Elu|SquarePlus)', x)
        if m:
            act = {'Softplus': 'softplus', 'Tanh': 'tanh', 'Relu': 'relu',
                   'LeakyRelu': 'leaky_relu', 'Sigmoid': 'sigmoid',
                   'Softsign': 'softsign', 'Elu': 'elu',
                   'SquarePlus': 'squareplus'}[m.group(1)]
            break
    variables = {}
    for name, node in consts.items():
        if re.match(r'^Atomic/[A-Za-z]+/(Conv1d\d+|Output)/(kernel|bias)$', name) or \
                re.match(r'^Atomic/[A-Za-z]+/(xlo|xhi)$', name):
            variables[name] = np.asarray(const_value(node), dtype=np.float64)
    for el in elements:
        sizes = []
        k = 1
        while f'Atomic/{el}/Conv1d{k}/kernel' in variables:
            sizes.append(variables[f'Atomic/{el}/Conv1d{k}/kernel'].shape[-1])
            k += 1
        hidden[el] = sizes
    use_resnet = any(re.match(r'^Atomic/[A-Za-z]+/Res\d+', x) for x in names)
    nn = AtomicNN(elements, SymmetryFunction(elements, **kw), hidden_sizes=hidden,
                  activation=act, use_resnet_dt=use_resnet,
                  minmax_scale=any(k.endswith('/xlo') for k in variables),
                  use_atomic_static_energy=any(k.endswith('Output/bias')
                                               for k in variables),
                  export_properties=_export_properties(ops))
    for k, v in variables.items():
        nn.set_variable(k, v)
    return nn
