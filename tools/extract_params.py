#!/usr/bin/env python
"""
One-off generator: pull the published potential PARAMETER TABLES (numbers from
Zhou-Johnson-Wadley PRB 69, 144113 etc.) out of the reference's python dict
literals with `ast` -- no reference code is executed or copied -- and write them
as JSON data files used by the product (`tensoralloy_b200/data/`) and,
independently, by the oracle (`oracle/data/`).

Run in the build container only (needs /root/reference):
    python tools/extract_params.py
"""
import ast
import json
import sys
from pathlib import Path

REF = Path('/root/reference/tensoralloy/nn/eam/potentials')
ROOT = Path(__file__).resolve().parent.parent


def literal_assignments(path, names):
    tree = ast.parse(path.read_text())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            t = node.targets[0]
            if isinstance(t, ast.Name) and t.id in names:
                out[t.id] = ast.literal_eval(node.value)
    return out


def dict_calls_in_method(path, cls, method):
    """Collect `params['X'] = dict(k=v, ...)` assignments (last one wins)."""
    tree = ast.parse(path.read_text())
    found = {}
    for c in ast.walk(tree):
        if isinstance(c, ast.ClassDef) and c.name == cls:
            for f in c.body:
                if isinstance(f, ast.FunctionDef) and f.name == method:
                    for node in ast.walk(f):
                        if (isinstance(node, ast.Assign)
                                and isinstance(node.targets[0], ast.Subscript)
                                and isinstance(node.value, ast.Call)
                                and getattr(node.value.func, 'id', '') == 'dict'):
                            key = ast.literal_eval(node.targets[0].slice)
                            found[key] = {
                                kw.arg: ast.literal_eval(kw.value)
                                for kw in node.value.keywords}
    return found


def main():
    zjw04 = literal_assignments(REF / 'zjw04.py', {'zjw04_defaults'})[
        'zjw04_defaults']
    xcp = dict_calls_in_method(REF / 'zjw04.py', 'Zjw04xcp', 'defaults')
    data = {'zjw04': zjw04, 'zjw04xcp_overrides': xcp}
    for target in (ROOT / 'tensoralloy_b200' / 'data' / 'zjw04.json',
                   ROOT / 'oracle' / 'data' / 'zjw04.json'):
        target.parent.mkdir(parents=True, exist_ok=True)
        target.write_text(json.dumps(data, indent=1, sort_keys=True) + '\n')
        print('wrote', target)


if __name__ == '__main__':
    sys.exit(main())
