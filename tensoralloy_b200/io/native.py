"""
The `.npz` model-file layout of the reference: `AtomicNN.export_to_lammps_native`
(tensoralloy/nn/atomic/atomic.py:304-480), the file LAMMPS `pair_style
tensoralloy/native` loads.  It is one of the three model-file layouts on the
drop-in boundary; the other two are the frozen `.pb`
(io/graph_model.py) and setfl (io/lammps.py).

`write_lammps_native` writes exactly the reference's keys:

    rmax, nelt, masses, numbers (two character codes per element), tdnp,
    precision, use_fnn, descriptor::method (0 pexp, 1 morse, 2 density, 3 sf)
    + descriptor::<param> as PAIRED lists, nlayers, max_moment, actfn, fctype,
    layer_sizes, use_resnet_dt, apply_output_bias, is_T_symmetric,
    weights_{i}_{j}, biases_{i}_{j}

`read_lammps_native` is the inverse: an `AtomicNN` over a
`GenericRadialAtomicPotential` with the transformer attached, ready for
`TensorAlloyCalculator` (which accepts the `.npz` path directly).

Differences to the reference, all on the safe side:
  * the reference drops the min-max input scaling (`xlo`, `xhi`,
    atomic.py:176-200) when it writes this file -- a model trained with
    `minmax_scale=True` would evaluate differently in LAMMPS.  Here the affine
    map x' = (xhi - x) / (xhi - xlo) (0 where xhi = xlo, `div_no_nan`) is
    folded into the first layer (W1' = -W1 / (xhi - xlo),
    b1' = b1 + xhi / (xhi - xlo) . W1), which is exact (the first layer never carries a
    ResNet link, convolutional.py:272).
  * a descriptor whose `moment_tensors` skip a moment (e.g. [0, 2]) cannot be
    expressed (only `max_moment` is stored); the reference writes such a file
    without complaint, here it is a ValueError.
  * `use_fnn = 1` files (GRAP `nn` algorithm: `fnn::*` keys, atomic.py:408-438) are
    written and read; such a model is served by nn/atomic/grap_nn.py (GrapFilterTrainer),
    not by TensorAlloyCalculator.
  * the file does not record `legacy_mode`; a file with `is_T_symmetric = 1` or
    `max_moment = 3` (new-mode-only features, grap.py:434-457, 485-494) reads back as a
    new-mode descriptor, any other file as a legacy one (equal numbers there).
"""
import numpy as np

from tensoralloy_b200.atoms import atomic_numbers

FCTYPE = {"cosine": 0, "polynomial": 1}                      # atomic.py:325
ACTFN = {"relu": 0, "softplus": 1, "tanh": 2, "squareplus": 3}   # atomic.py:326
METHOD = {"pexp": 0, "morse": 1, "density": 2, "sf": 3}      # atomic.py:379-407
METHOD_KEYS = {"pexp": ("rl", "pl"), "morse": ("D", "gamma", "r0"),
               "density": ("A", "beta", "re"), "sf": ("eta", "omega")}

# standard atomic weights (the table `ase.data.atomic_masses` holds, IUPAC 2016)
_WEIGHTS = """H 1.008 He 4.002602 Li 6.94 Be 9.0121831 B 10.81 C 12.011 N 14.007 O 15.999
F 18.998403163 Ne 20.1797 Na 22.98976928 Mg 24.305 Al 26.9815385 Si 28.085 P 30.973761998
S 32.06 Cl 35.45 Ar 39.948 K 39.0983 Ca 40.078 Sc 44.955908 Ti 47.867 V 50.9415 Cr 51.9961
Mn 54.938044 Fe 55.845 Co 58.933194 Ni 58.6934 Cu 63.546 Zn 65.38 Ga 69.723 Ge 72.63
As 74.921595 Se 78.971 Br 79.904 Kr 83.798 Rb 85.4678 Sr 87.62 Y 88.90584 Zr 91.224
Nb 92.90637 Mo 95.95 Tc 97.90721 Ru 101.07 Rh 102.9055 Pd 106.42 Ag 107.8682 Cd 112.414
In 114.818 Sn 118.71 Sb 121.76 Te 127.6 I 126.90447 Xe 131.293 Cs 132.90545196 Ba 137.327
La 138.90547 Ce 140.116 Pr 140.90766 Nd 144.242 Pm 144.91276 Sm 150.36 Eu 151.964
Gd 157.25 Tb 158.92535 Dy 162.5 Ho 164.93033 Er 167.259 Tm 168.93422 Yb 173.054
Lu 174.9668 Hf 178.49 Ta 180.94788 W 183.84 Re 186.207 Os 190.23 Ir 192.217 Pt 195.084
Au 196.966569 Hg 200.592 Tl 204.38 Pb 207.2 Bi 208.9804 Po 208.98243 At 209.98715
Rn 222.01758 Fr 223.01974 Ra 226.02541 Ac 227.02775 Th 232.0377 Pa 231.03588
U 238.02891 Np 237.04817 Pu 244.06421""".split()
atomic_masses = {_WEIGHTS[i]: float(_WEIGHTS[i + 1]) for i in range(0, len(_WEIGHTS), 2)}


def _element_codes(elements):
    """atomic.py:359-366: two int32 character codes per element, a one-letter
    symbol is padded with 0."""
    chars = []
    for elt in elements:
        if len(elt) == 1:
            chars += [ord(elt[0]), 0]
        else:
            chars += [ord(c) for c in elt]
    return np.array(chars, dtype=np.int32)


def _decode_elements(numbers, nelt):
    numbers = np.asarray(numbers).reshape(-1)
    if numbers.size != 2 * nelt:
        raise ValueError("npz model: `numbers` must hold two codes per element")
    out = []
    for k in range(nelt):
        a, b = int(numbers[2 * k]), int(numbers[2 * k + 1])
        out.append(chr(a) + (chr(b) if b else ""))
    return out


def _folded_layers(nn, element):
    """The element's [W...], [b...] with the min-max input scaling folded into the
    first layer (see the module docstring)."""
    p = nn.mlp_params(element)
    W = [np.array(w, dtype=np.float64) for w in p['weights']]
    b = [None if x is None else np.array(x, dtype=np.float64) for x in p['biases']]
    if p['xlo'] is not None:
        # atomic.py:199: x' = div_no_nan(xhi - x, xhi - xlo)
        lo, hi = p['xlo'], p['xhi']
        den = hi - lo
        s = np.where(den == 0.0, 0.0, 1.0 / np.where(den == 0.0, 1.0, den))
        b0 = (b[0] if b[0] is not None else 0.0) + (hi * s) @ W[0]
        W[0], b[0] = -W[0] * s[:, None], np.asarray(b0, dtype=np.float64)
    return W, b


def lammps_native_dict(nn, dtype=np.float64):
    """The dict `np.savez` receives (atomic.py:368-478), key for key."""
    from tensoralloy_b200.nn.atomic.grap import GenericRadialAtomicPotential
    desc = nn.descriptor
    if not isinstance(desc, GenericRadialAtomicPotential):
        raise ValueError("The descriptor GenericRadialAtomicPotential is required")
    clf = nn.transformer
    if clf is None:
        raise ValueError("A transformer must be attached before exporting")
    elements = list(clf.elements)
    hs = nn.hidden_sizes
    layer_sizes = np.array(hs[elements[0]], dtype=int)
    for elt in elements[1:]:
        if len(hs[elt]) != len(layer_sizes) or not np.all(hs[elt] == layer_sizes):
            raise ValueError("Layer sizes of all elements must be the same")
    layer_sizes = np.append(layer_sizes, 1).astype(np.int32)
    if nn.activation not in ACTFN:
        raise KeyError(nn.activation)
    if list(desc.moments()) != list(range(desc.max_moment + 1)):
        # the file stores max_moment only and its reader evaluates 0..max_moment
        raise ValueError("the npz layout stores `max_moment` only: moment_tensors must "
                         f"be 0..{desc.max_moment}, got {list(desc.moments())}")
    for elt in elements:
        if elt not in atomic_masses:
            raise ValueError(f"no atomic mass for element {elt}")
    dtype = np.dtype(dtype).type
    data = {"rmax": dtype(clf.rcut),
            "nelt": np.int32(len(elements)),
            "masses": np.array([atomic_masses[e] for e in elements], dtype=dtype),
            "numbers": _element_codes(elements),
            "tdnp": np.int32(0),
            "precision": np.int32(64 if dtype == np.float64 else 32),
            "use_fnn": np.int32(0)}
    algo = desc.algorithm
    if algo == 'nn':
        # atomic.py:408-438: the filter network; kernels squeezed as the reference does
        # (the first one, [1, 1, 1, 1, h], becomes 1-D)
        from tensoralloy_b200.nn.atomic.grap_nn import filter_params
        a, fp = desc.algorithm_object, filter_params(nn)
        if a.activation not in ACTFN:
            raise KeyError(a.activation)
        data["use_fnn"] = np.int32(1)
        data["fnn::nlayers"] = np.int32(len(a.hidden_sizes) + 1)
        data["fnn::layer_sizes"] = np.array(list(a.hidden_sizes) + [a.num_filters],
                                            dtype=np.int32)
        data["fnn::num_filters"] = np.int32(a.num_filters)
        data["fnn::actfn"] = np.int32(ACTFN[a.activation])
        data["fnn::use_resnet_dt"] = np.int32(a.use_resnet_dt)
        data["fnn::apply_output_bias"] = np.int32(0)
        data["fnn::h_abck_modifier"] = np.int32(a.h_abck_modifier)
        nf = len(fp['weights'])
        for j in range(nf):
            data[f"fnn::weights_0_{j}"] = np.squeeze(fp['weights'][j]).astype(dtype)
            if j < nf - 1:
                data[f"fnn::biases_0_{j}"] = np.squeeze(fp['biases'][j]).astype(dtype)
    else:
        data["descriptor::method"] = np.int32(METHOD[algo])
        for key in METHOD_KEYS[algo]:                   # as_dict(convert_to_pairs=True)
            data[f"descriptor::{key}"] = np.array([row[key] for row in desc.grid],
                                                  dtype=dtype)
    data["nlayers"] = np.int32(len(layer_sizes))
    data["max_moment"] = np.int32(desc.max_moment)
    data["actfn"] = np.int32(ACTFN[nn.activation])
    data["fctype"] = np.int32(FCTYPE[desc.cutoff_function])
    data["layer_sizes"] = layer_sizes
    data["use_resnet_dt"] = np.int32(bool(nn.use_resnet_dt))
    data["apply_output_bias"] = np.int32(bool(nn.use_atomic_static_energy))
    data["is_T_symmetric"] = np.int32(bool(desc.is_T_symmetric))
    for i, elt in enumerate(elements):
        W, b = _folded_layers(nn, elt)
        last = len(layer_sizes) - 1
        for j in range(last):
            data[f"weights_{i}_{j}"] = W[j].astype(dtype)
            data[f"biases_{i}_{j}"] = np.asarray(b[j]).reshape(-1).astype(dtype)
        data[f"weights_{i}_{last}"] = W[last].reshape(-1).astype(dtype)
        if nn.use_atomic_static_energy and b[last] is not None:
            # np.squeeze of a one-element bias is 0-d in the reference's file
            data[f"biases_{i}_{last}"] = np.squeeze(np.asarray(b[last]).astype(dtype))
    return data


def write_lammps_native(nn, model_path, dtype=np.float64):
    np.savez(model_path, **lammps_native_dict(nn, dtype))


def read_lammps_native(model_path, export_properties=('energy', 'forces', 'stress')):
    """`.npz` written by the reference (or by `write_lammps_native`) -> AtomicNN with
    an attached UniversalTransformer."""
    from tensoralloy_b200.nn.atomic import AtomicNN
    from tensoralloy_b200.nn.atomic.grap import GenericRadialAtomicPotential
    from tensoralloy_b200.transformer import UniversalTransformer
    z = np.load(model_path)
    need = ("rmax", "nelt", "numbers", "descriptor::method", "nlayers", "max_moment",
            "actfn", "fctype", "layer_sizes")
    missing = [k for k in need if k not in z.files and k != "descriptor::method"]
    if missing:
        raise ValueError(f"npz model: missing keys {missing}")
    use_fnn = bool(int(z["use_fnn"])) if "use_fnn" in z.files else False
    if "tdnp" in z.files and int(z["tdnp"]):
        raise ValueError("npz model: temperature-dependent (tdnp) files are not supported")
    nelt = int(z["nelt"])
    elements = _decode_elements(z["numbers"], nelt)
    if sorted(elements) != elements:
        raise ValueError("npz model: elements must be sorted")
    for e in elements:
        if e not in atomic_numbers:
            raise ValueError(f"npz model: unknown element '{e}'")
    if use_fnn:
        algo = 'nn'
        fsizes = [int(x) for x in np.asarray(z["fnn::layer_sizes"]).reshape(-1)]
        if len(fsizes) != int(z["fnn::nlayers"]) or fsizes[-1] != int(z["fnn::num_filters"]):
            raise ValueError("npz model: inconsistent fnn::layer_sizes / nlayers / num_filters")
        if int(z["fnn::apply_output_bias"]) if "fnn::apply_output_bias" in z.files else 0:
            raise ValueError("npz model: a filter network with an output bias is not "
                             "something the reference writes (atomic.py:418)")
        params = dict(hidden_sizes=fsizes[:-1], num_filters=fsizes[-1],
                      activation={v: k for k, v in ACTFN.items()}[int(z["fnn::actfn"])],
                      use_resnet_dt=bool(int(z["fnn::use_resnet_dt"])),
                      h_abck_modifier=int(z["fnn::h_abck_modifier"])
                      if "fnn::h_abck_modifier" in z.files else 0)
    else:
        if "descriptor::method" not in z.files:
            raise ValueError("npz model: missing keys ['descriptor::method']")
        algo = {v: k for k, v in METHOD.items()}[int(z["descriptor::method"])]
        params = {k: np.asarray(z[f"descriptor::{k}"], dtype=np.float64).reshape(-1).tolist()
                  for k in METHOD_KEYS[algo]}
    max_moment = int(z["max_moment"])
    if max_moment > 5:
        raise ValueError("npz model: moments up to 5 are supported")       # grap.py:578-579
    fct = {v: k for k, v in FCTYPE.items()}[int(z["fctype"])]
    act = {v: k for k, v in ACTFN.items()}[int(z["actfn"])]
    sym = bool(int(z["is_T_symmetric"])) if "is_T_symmetric" in z.files else False
    # the file does not say which formulation trained the model; they agree for moments
    # <= 2 with the plain multiplicity tensor (test_grap.py:46-149), so only a traceless
    # T_dm or moment 3 (both exist in new mode only, grap.py:434-457, 485-494) select it
    desc = GenericRadialAtomicPotential(
        elements, algorithm=algo, parameters=params, param_space_method='pair',
        moment_tensors=list(range(max_moment + 1)), cutoff_function=fct, symmetric=sym,
        legacy_mode=not (sym or max_moment > 2 or use_fnn))
    sizes = [int(x) for x in np.asarray(z["layer_sizes"]).reshape(-1)]
    if len(sizes) != int(z["nlayers"]) or sizes[-1] != 1:
        raise ValueError("npz model: inconsistent layer_sizes / nlayers")
    hidden = sizes[:-1]
    out_bias = bool(int(z["apply_output_bias"])) if "apply_output_bias" in z.files else False
    nn = AtomicNN(elements, desc, hidden_sizes={e: list(hidden) for e in elements},
                  activation=act, minmax_scale=False,
                  use_resnet_dt=bool(int(z["use_resnet_dt"])) if "use_resnet_dt" in z.files
                  else False,
                  use_atomic_static_energy=out_bias,
                  export_properties=tuple(export_properties))
    clf = UniversalTransformer(elements, rcut=float(z["rmax"]), angular=False)
    nn.attach_transformer(clf)
    dim = desc.dimension(False)
    last = len(sizes) - 1
    for i, e in enumerate(elements):
        fan = dim
        for j in range(last):
            w = np.asarray(z[f"weights_{i}_{j}"], dtype=np.float64).reshape(fan, sizes[j])
            nn.set_variable(f"{nn.scope}/{e}/Conv1d{j + 1}/kernel", w[None])
            nn.set_variable(f"{nn.scope}/{e}/Conv1d{j + 1}/bias",
                            np.asarray(z[f"biases_{i}_{j}"], dtype=np.float64).reshape(-1))
            fan = sizes[j]
        w = np.asarray(z[f"weights_{i}_{last}"], dtype=np.float64).reshape(fan, 1)
        nn.set_variable(f"{nn.scope}/{e}/Output/kernel", w[None])
        if out_bias:
            nn.set_variable(f"{nn.scope}/{e}/Output/bias",
                            np.asarray(z[f"biases_{i}_{last}"], dtype=np.float64).reshape(1))
    if use_fnn:
        from tensoralloy_b200.nn.atomic.grap_nn import filter_scope
        fan = 1
        nf = len(fsizes)
        for j in range(nf):
            w = np.asarray(z[f"fnn::weights_0_{j}"], dtype=np.float64).reshape(fan, fsizes[j])
            name = f"Conv3d{j + 1}" if j < nf - 1 else "Output"
            nn.set_variable(f"{filter_scope(nn)}/{name}/kernel", w[None, None, None])
            if j < nf - 1:
                nn.set_variable(f"{filter_scope(nn)}/{name}/bias",
                                np.asarray(z[f"fnn::biases_0_{j}"],
                                           dtype=np.float64).reshape(-1))
            fan = fsizes[j]
    precision = 'medium' if ("precision" in z.files and int(z["precision"]) == 32) else 'high'
    return nn, precision
