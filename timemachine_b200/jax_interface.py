"""Bridge from the GPU implementations to automatic differentiation: the counterpart of the reference's
`timemachine/potentials/jax_interface.py:12-66`.

The reference registers `call_unbound_impl(impl, conf, params, box) -> u` and `call_bound_impl(impl, conf, box) -> u`
as `jax.custom_jvp` functions whose JVP rule asks the kernel only for the derivatives that are being traced
(`impl.execute(x, p, box, compute_du_dx, compute_du_dp, True)`) and returns
`(u, sum(du_dx * dx) + sum(du_dp * dp))`; a traced box is refused.

jax is not part of this image, so the rule is stated twice here:

  * `unbound_impl_jvp` / `bound_impl_jvp`: the rule itself in NumPy - a tangent that is `None` plays the part of "not a
    Tracer" (that derivative is not requested from the kernel); this is what the GPU tests check against finite
    differences and against `execute`;
  * when `import jax` succeeds, `call_unbound_impl` / `call_bound_impl` are the same `jax.custom_jvp` functions as the
    reference's, built on those rules, so `jax.grad(wrapper)(conf, params, box)` works on this module like on the
    reference's.  Without jax they are plain functions returning the energy.
"""

from __future__ import annotations

import numpy as np


def unbound_impl_jvp(impl, primals, tangents):
    """(u, du) for `u = impl(conf, params, box)`; tangents = (dx | None, dp | None, dbox | None)."""
    x, p, box = primals
    dx, dp, dbox = tangents
    if dbox is not None:
        raise RuntimeError("box derivatives not supported")
    compute_du_dx = dx is not None
    compute_du_dp = dp is not None
    du_dx, du_dp, u = impl.execute(x, p, box, compute_du_dx, compute_du_dp, True)
    tangent_out = np.zeros_like(u)
    if compute_du_dx:
        tangent_out = tangent_out + np.sum(du_dx * np.asarray(dx))
    if compute_du_dp:
        tangent_out = tangent_out + np.sum(du_dp * np.asarray(dp).reshape(np.shape(du_dp)))
    return u, tangent_out


def bound_impl_jvp(impl, primals, tangents):
    """(u, du) for `u = bound_impl(conf, box)`; tangents = (dx | None, dbox | None)."""
    x, box = primals
    dx, dbox = tangents
    if dbox is not None:
        raise RuntimeError("box derivatives not supported")
    compute_du_dx = dx is not None
    du_dx, u = impl.execute(x, box, compute_du_dx, True)
    tangent_out = np.zeros_like(u)
    if compute_du_dx:
        tangent_out = tangent_out + np.sum(du_dx * np.asarray(dx))
    return u, tangent_out


def _plain_call_unbound_impl(impl, conf, params, box):
    _, _, u = impl.execute(conf, params, box, False, False, True)
    return u


def _plain_call_bound_impl(impl, conf, box):
    _, u = impl.execute(conf, box, compute_du_dx=False)
    return u


try:  # pragma: no cover - jax is absent from the build image
    import jax
    from jax.core import Tracer
    from functools import partial

    @partial(jax.custom_jvp, nondiff_argnums=(0,))
    def call_unbound_impl(impl, conf, params, box):
        return _plain_call_unbound_impl(impl, conf, params, box)

    @partial(jax.custom_jvp, nondiff_argnums=(0,))
    def call_bound_impl(impl, conf, box):
        return _plain_call_bound_impl(impl, conf, box)

    def _untraced(t):
        return t if isinstance(t, Tracer) else None

    @call_unbound_impl.defjvp
    def _(impl, primals, tangents):
        return unbound_impl_jvp(impl, primals, tuple(_untraced(t) for t in tangents))

    @call_bound_impl.defjvp
    def _(impl, primals, tangents):
        return bound_impl_jvp(impl, primals, tuple(_untraced(t) for t in tangents))

except ImportError:
    call_unbound_impl = _plain_call_unbound_impl
    call_bound_impl = _plain_call_bound_impl
