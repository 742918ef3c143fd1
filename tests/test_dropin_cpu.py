"""CPU: the drop-in claim, first half.  tests/golden/dropin_calls.json is the list of `custom_ops` constructor calls the
reference's UNMODIFIED wrappers make (timemachine/potentials/potential.py:28-37, potentials.py:128-304,
lib/__init__.py:12-62; recorded by tests/golden/make_golden_dropin.py on a recording stand-in for the extension).  Here:
every class the wrappers look up exists in this repo's module, and every recorded argument list binds to its
constructor's signature.  tests/test_dropin_gpu.py replays the calls on the GPU."""

import inspect
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def trace():
    return json.loads((GOLDEN / "dropin_calls.json").read_text())["calls"]


def test_every_class_the_reference_wrappers_ask_for_exists(trace):
    from timemachine_b200 import custom_ops

    names = sorted({c["cls"] for c in trace})
    assert len(names) >= 29
    for n in names:
        assert hasattr(custom_ops, n), f"the reference's wrappers construct custom_ops.{n}; this module has no such class"
        assert inspect.isclass(getattr(custom_ops, n))
    # the base class the wrappers annotate with, and the install hook
    assert inspect.isclass(custom_ops.Potential)


def test_recorded_arguments_bind_to_the_constructors(trace):
    from timemachine_b200 import custom_ops

    for c in trace:
        sig = inspect.signature(getattr(custom_ops, c["cls"]).__init__)
        try:
            sig.bind(None, *c["args"], **c["kwargs"])
        except TypeError as e:
            pytest.fail(f"custom_ops.{c['cls']}{sig} does not accept the reference's call #{c['id']}: {e}")


def test_install_hook_aliases_the_module():
    import timemachine_b200
    from timemachine_b200 import custom_ops

    saved = {k: sys.modules.get(k) for k in ("timemachine.lib.custom_ops",)}
    try:
        timemachine_b200.install_as_timemachine_custom_ops()
        assert sys.modules["timemachine.lib.custom_ops"] is custom_ops
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not Path("/root/reference/timemachine/potentials/potentials.py").exists(), reason="needs the reference source tree")
def test_fixture_is_what_the_reference_wrappers_do_today(tmp_path):
    """Re-record from /root/reference (only where it exists: this container) and compare with the committed fixture."""
    code = (
        "import sys, json; sys.path.insert(0, 'tests/golden'); import make_golden_dropin as M; t = M.record(); "
        "print(json.dumps(t.calls))"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=Path(__file__).resolve().parents[1], check=True)
    live = json.loads(out.stdout.strip().splitlines()[-1])
    assert live == json.loads((GOLDEN / "dropin_calls.json").read_text())["calls"]


def test_jvp_rule_requests_only_traced_derivatives():
    """jax_interface.py:27-46: du_dx / du_dp are asked of the kernel only when their tangent is traced; a traced box is
    refused; the tangent is sum(du_dx * dx) + sum(du_dp * dp)."""
    from timemachine_b200 import jax_interface as J

    asked = []

    class Impl:
        def execute(self, x, p, box, want_dx=True, want_dp=True, want_u=True):
            asked.append((want_dx, want_dp, want_u))
            return (2.0 * x if want_dx else None, 3.0 * p if want_dp else None, np.float64(5.0))

    x, p, box = np.arange(6.0).reshape(2, 3), np.arange(4.0), np.eye(3)
    dx, dp = np.ones_like(x), np.full_like(p, 0.5)
    u, t = J.unbound_impl_jvp(Impl(), (x, p, box), (dx, dp, None))
    assert u == 5.0 and t == np.sum(2 * x) + np.sum(3 * p * 0.5) and asked[-1] == (True, True, True)
    u, t = J.unbound_impl_jvp(Impl(), (x, p, box), (None, dp, None))
    assert t == np.sum(3 * p * 0.5) and asked[-1] == (False, True, True)
    u, t = J.unbound_impl_jvp(Impl(), (x, p, box), (None, None, None))
    assert t == 0.0 and asked[-1] == (False, False, True)
    with pytest.raises(RuntimeError, match="box derivatives not supported"):
        J.unbound_impl_jvp(Impl(), (x, p, box), (dx, None, np.eye(3)))
    assert J.call_unbound_impl(Impl(), x, p, box) == 5.0 and asked[-1] == (False, False, True)


def test_mirror_dataclasses_make_the_same_calls_as_the_reference_wrappers(monkeypatch, trace):
    """`timemachine_b200.potentials` / `timemachine_b200.lib` (the host-side mirror of the reference's dataclasses) run
    through the SAME scenario on the SAME recording module: the constructor calls must equal, call for call and argument
    for argument (arrays bitwise), what the reference's own wrappers produced."""
    sys.path.insert(0, str(GOLDEN))
    import make_golden_dropin as M

    from timemachine_b200 import lib, potentials

    M.TRACE.calls.clear()
    M.TRACE.arrays.clear()
    rec = M.make_recording_module()
    monkeypatch.setattr(potentials, "custom_ops", rec)
    monkeypatch.setattr(lib, "custom_ops", rec)
    M.scenario(potentials, lib)
    mine, mine_arrays = list(M.TRACE.calls), dict(M.TRACE.arrays)
    golden_arrays = dict(np.load(GOLDEN / "dropin_calls.npz"))
    assert [c["cls"] for c in mine] == [c["cls"] for c in trace]

    def same(a, b, where):
        assert a.keys() == b.keys(), where
        if "array" in a:
            x, y = mine_arrays[a["array"]], golden_arrays[b["array"]]
            assert x.shape == y.shape and np.array_equal(x, y), where
        elif "list" in a:
            assert len(a["list"]) == len(b["list"]), where
            for i, (u, v) in enumerate(zip(a["list"], b["list"])):
                same(u, v, f"{where}[{i}]")
        else:
            assert a == b, where

    for c, g in zip(mine, trace):
        assert len(c["args"]) == len(g["args"]) and c["kwargs"].keys() == g["kwargs"].keys(), f"call #{g['id']} {g['cls']}"
        for i, (u, v) in enumerate(zip(c["args"], g["args"])):
            same(u, v, f"call #{g['id']} {g['cls']} arg {i}")
        for k in g["kwargs"]:
            same(c["kwargs"][k], g["kwargs"][k], f"call #{g['id']} {g['cls']} kwarg {k}")
