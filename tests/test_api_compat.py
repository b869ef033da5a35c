"""Drop-in check at the boundary of SURVEY 8b: every callable the reference's train scripts / geometry modules use exists here under
the same module path and name, takes the reference's parameters in the reference's order with the reference's defaults (extra
TRAILING optional parameters are allowed, e.g. `device=`, `perms=`).  Reference signatures: tests/golden/api_signatures.json,
extracted from the reference source by tests/golden/make_golden_api.py (the train scripts themselves cannot be run in this
environment: nvdiffrast / xatlas / tiny-cuda-nn / kaolin are absent)."""
import ast
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIGS = json.load(open(os.path.join(HERE, "golden", "api_signatures.json")))


def _product_callable(key):
    rel, name = key.split("::")
    mod = importlib.import_module("gshell_b200." + rel[:-3].replace("/", "."))
    obj = mod
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj


def _value(text):
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


@pytest.mark.parametrize("key", sorted(SIGS))
def test_signature_is_drop_in(key):
    fn = _product_callable(key)
    want = SIGS[key]
    params = list(inspect.signature(fn).parameters.values())
    if want and want[0][0] == "self" and (not params or params[0].name != "self"):
        want = want[1:]                                   # bound / unbound difference only
    assert len(params) >= len(want), (key, [p.name for p in params])
    for p, (name, default) in zip(params, want):
        assert p.name == name, (key, p.name, name)
        if default is None:
            assert p.default is inspect.Parameter.empty, (key, name, "reference has no default")
        else:
            assert p.default is not inspect.Parameter.empty and p.default == _value(default), (key, name, p.default, default)
    for p in params[len(want):]:                          # additions must not break reference call sites
        assert p.default is not inspect.Parameter.empty or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD), (key, p.name)
