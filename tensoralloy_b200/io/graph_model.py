"""
Reader (and writer, see `write_graph_model`) of the reference's exported model files: frozen TensorFlow `GraphDef`
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
    if 'Model/params' in consts:          # written by `write_graph_model` below
        nn = _model_from_params(consts, clf, ops)
        return LoadedModel(nn=nn, precision=precision, api_version=meta('api', '1.0'),
                           timestamp=meta('timestamp'), predict_properties=list(ops),
                           ops=ops)
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


# --------------------------------------------------------------------------
# writer: `BasicNN.export` (basic.py:1017-1153)
# --------------------------------------------------------------------------
_OP_NAMES = {   # tensor names of the reference's output ops (basic.py:742-787)
    'energy': 'Output/Energy/energy:0', 'atomic': 'Output/Energy/atomic:0',
    'free_energy': 'Output/Energy/free_energy:0', 'eentropy': 'Output/Energy/eentropy:0',
    'forces': 'Output/Forces/forces:0', 'stress': 'Output/Stress/Voigt/stress:0',
    'total_pressure': 'Output/Pressure/pressure:0', 'hessian': 'Output/Hessian/hessian:0',
    'elastic': 'Output/Elastic/elastic:0',
}


def _const_node(g, name, value, np_dtype=None):
    from tensorboard.compat.proto import tensor_pb2, tensor_shape_pb2
    node = g.node.add()
    node.name = name
    node.op = 'Const'
    t = tensor_pb2.TensorProto()
    if isinstance(value, (str, bytes)):
        t.dtype = _DT_STRING
        t.string_val.append(value.encode('utf-8') if isinstance(value, str) else value)
        t.tensor_shape.CopyFrom(tensor_shape_pb2.TensorShapeProto())
    else:
        arr = np.ascontiguousarray(np.asarray(value, dtype=np_dtype))
        t.dtype = {np.dtype(np.float32): _DT_FLOAT, np.dtype(np.float64): _DT_DOUBLE,
                   np.dtype(np.int32): _DT_INT32, np.dtype(np.int64): _DT_INT64}[arr.dtype]
        for d in arr.shape:
            t.tensor_shape.dim.add().size = int(d)
        t.tensor_content = arr.tobytes()
    node.attr['dtype'].type = t.dtype
    node.attr['value'].tensor.CopyFrom(t)
    return node


def _jsonable(x):
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


def _shared_parameters(nn):
    """{potential: {section: {key: value}}} of the empirical potentials an EAM-family
    model uses, restricted to sections made of the model's own elements."""
    from tensoralloy_b200.utils import get_elements_from_kbody_term
    used = {name for fns in nn.potentials.values() for name in fns.values()
            if name != 'nn' and not name.startswith('spline@')}
    out = {}
    for pname in sorted(used):
        fn = nn._fn_of(*next((s, f) for s, fns in nn.potentials.items()
                             for f, v in fns.items() if v == pname))
        keep = {}
        for section, vals in fn.params.items():
            try:
                els = get_elements_from_kbody_term(section)
            except Exception:
                els = [section]
            if all(e in nn.elements for e in els):
                keep[section] = _jsonable(vals)
        out[pname] = keep
    return out


def write_graph_model(nn, output_graph_path, precision=None):
    """A frozen `.pb` in the reference's node naming that `load_graph_model` (and so
    `TensorAlloyCalculator(path)`) reads back: `Transformer/params`, `Metadata/*`
    (basic.py:1075-1092), every variable as a `Const` node under its reference name
    (`Atomic/<El>/Conv1d{k}/{kernel,bias}`, `EAM/Shared/<section>/<param>`, ...).

    It is a PARAMETER CONTAINER: there is no TensorFlow in this image, so the file holds
    no executable ops and the reference's TF calculator cannot run it (its
    `Metadata/tf_version` says so).  Two extra constants make the round trip exact
    without parsing op scopes: `Model/params` (JSON of `nn.as_dict()`) and, for the
    EAM family, `Model/shared` (the shared potential parameters per potential)."""
    from datetime import datetime
    from tensorboard.compat.proto import graph_pb2
    from tensoralloy_b200.precision import get_float_precision
    if nn.transformer is None:
        raise ValueError("A transformer must be attached before exporting to a pb file.")
    precision = precision or get_float_precision().name
    dt = np.float64 if precision == 'high' else np.float32
    clf = nn.transformer
    g = graph_pb2.GraphDef()
    g.versions.producer = 134           # TF 1.15 GraphDef version, for protobuf readers
    params = _jsonable(clf.as_dict())
    _const_node(g, 'Transformer/params', json.dumps(params))
    _const_node(g, 'Metadata/timestamp', str(datetime.today()))
    _const_node(g, 'Metadata/precision', precision)
    _const_node(g, 'Metadata/tf_version', 'none (tensoralloy_b200 parameter container)')
    _const_node(g, 'Metadata/variational_energy', nn.variational_energy)
    _const_node(g, 'Metadata/is_finite_temperature', int(nn.is_finite_temperature),
                np.int32)
    _const_node(g, 'Metadata/api', '1.1')
    props = list(nn.predict_properties)
    if nn.is_finite_temperature:
        props += [p for p in ('free_energy', 'eentropy') if p not in props]
    ops = {p: _OP_NAMES[p] for p in props if p in _OP_NAMES}
    _const_node(g, 'Metadata/ops', json.dumps(ops))
    _const_node(g, 'Model/params', json.dumps(_jsonable(nn.as_dict())))
    if hasattr(nn, 'potentials'):
        shared = _shared_parameters(nn)
        _const_node(g, 'Model/shared', json.dumps(shared))
        for pname, sections in shared.items():
            for section, vals in sections.items():
                for key, v in vals.items():
                    name = f'{nn.scope}/Shared/{section}/{key}'
                    if np.ndim(v) == 0 and not any(n.name == name for n in g.node):
                        _const_node(g, name, v, dt)
    if not getattr(nn, 'variables', None) and hasattr(nn, 'initialize_variables') \
            and not hasattr(nn, 'potentials'):
        nn.initialize_variables()
    for name, value in nn.variables.items():
        _const_node(g, name, value, dt)
    with open(output_graph_path, 'wb') as fp:
        fp.write(g.SerializeToString())


def _model_from_params(consts, clf, ops):
    """Rebuild the model of a file written by `write_graph_model`."""
    cfg = json.loads(const_value(consts['Model/params']).decode('utf-8'))
    cls_name = cfg.pop('class')
    from tensoralloy_b200.nn import atomic as _atomic, eam as _eam
    import tensoralloy_b200.nn.atomic.finite_temperature as _ft
    cls = getattr(_atomic, cls_name, None) or getattr(_eam, cls_name, None) or \
        getattr(_ft, cls_name, None)
    if cls is None:
        raise ValueError(f"unknown model class '{cls_name}'")
    cfg['export_properties'] = [p for p in cfg.get('export_properties', [])] or \
        _export_properties(ops)
    nn = cls(**cfg)
    nn.attach_transformer(clf)
    if 'Model/shared' in consts:
        shared = json.loads(const_value(consts['Model/shared']).decode('utf-8'))
        for pname, sections in shared.items():
            fn = nn._empirical_functions[pname]
            for section, vals in sections.items():
                for key, v in vals.items():
                    if np.ndim(v) == 0:
                        fn.set_param(section, key, v)
                    else:
                        fn.params.setdefault(section, {})[key] = v
        nn._model = None
    skip = ('Transformer/', 'Metadata/', 'Model/', f'{nn.scope}/Shared/')
    for name, node in consts.items():
        if name.startswith(skip):
            continue
        v = const_value(node)
        if v is not None and not isinstance(v, bytes):
            nn.set_variable(name, np.asarray(v, dtype=np.float64))
    return nn
