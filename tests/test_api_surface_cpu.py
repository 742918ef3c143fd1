"""CPU: the module offers every class, method and function of the reference's timemachine/lib/custom_ops.pyi with the same
positional / keyword argument names, the same number of required arguments and the same base classes
(tests/golden/custom_ops_api.json, generated from the stub by tests/golden/make_golden_api.py)."""

import inspect
import json
from pathlib import Path

import pytest

API = json.loads((Path(__file__).parent / "golden" / "custom_ops_api.json").read_text())


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def _positional(fn):
    ps = [p for p in inspect.signature(fn).parameters.values() if p.name != "self"]
    pos = [p for p in ps if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    return [p.name for p in pos], sum(1 for p in pos if p.default is p.empty)


def _check(name, fn, sig, problems):
    if sig["varargs"] or sig["kwargs"]:
        return  # "no constructor" placeholders of the abstract bases
    names, required = _positional(fn)
    if names != sig["args"]:
        problems.append(f"{name}: arguments {names}, the reference has {sig['args']}")
    elif required != sig["required"]:
        problems.append(f"{name}: {required} required arguments, the reference has {sig['required']}")


def test_every_name_of_the_reference_stub_exists_with_the_same_signature():
    o = ops()
    problems = []
    assert len(API["classes"]) == 50 and len(API["functions"]) == 12
    for cname, c in API["classes"].items():
        cls = getattr(o, cname, None)
        if cls is None:
            problems.append(f"class {cname} is missing")
            continue
        for base in c["bases"]:
            b = getattr(o, base.split(".")[-1], None)
            if b is not None and not issubclass(cls, b):
                problems.append(f"{cname} does not derive from {base}")
        for mname, sig in c["methods"].items():
            m = getattr(cls, mname, None)
            if m is None:
                problems.append(f"{cname}.{mname} is missing")
            else:
                _check(f"{cname}.{mname}", m, sig, problems)
    for fname, sig in API["functions"].items():
        f = getattr(o, fname, None)
        if f is None:
            problems.append(f"function {fname} is missing")
        else:
            _check(fname, f, sig, problems)
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("name", ["Potential", "Integrator", "Mover"])
def test_abstract_bases_have_no_constructor(name):
    with pytest.raises(TypeError):
        getattr(ops(), name)()
