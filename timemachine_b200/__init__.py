"""timemachine_b200: B200-native (sm_100a) force evaluation + Langevin integration behind timemachine's custom_ops API.

Importing `timemachine_b200.custom_ops` loads the CUDA library; it raises ImportError if the library has not been built
(no CPU fallback exists).  `install_as_timemachine_custom_ops()` aliases the module under the reference's name so the
reference's own `timemachine.potentials` / `timemachine.lib` wrappers run on top of it.
"""

__version__ = "0.1.0"


def install_as_timemachine_custom_ops() -> None:
    """Make `import timemachine.lib.custom_ops` resolve to this implementation (call before importing timemachine)."""
    import sys

    from . import custom_ops

    sys.modules["timemachine.lib.custom_ops"] = custom_ops
    lib_pkg = sys.modules.get("timemachine.lib")
    if lib_pkg is not None:
        setattr(lib_pkg, "custom_ops", custom_ops)
