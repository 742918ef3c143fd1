"""Generate tests/golden/custom_ops_api.json: every class, method and module-level function of the reference's
timemachine/lib/custom_ops.pyi (the stub its build generates from the compiled module) with the argument names of each
callable - the API a drop-in `custom_ops` has to offer.

    python tests/golden/make_golden_api.py     (here, where /root/reference exists)
"""

import ast
import json
from pathlib import Path

REF = Path("/root/reference/timemachine/lib/custom_ops.pyi")
OUT = Path(__file__).resolve().parent / "custom_ops_api.json"


def arg_names(fn: ast.FunctionDef):
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    if names and names[0] == "self":
        names = names[1:]
    n_required = len(a.posonlyargs + a.args) - len(a.defaults) - (1 if (a.posonlyargs + a.args) and (a.posonlyargs + a.args)[0].arg == "self" else 0)
    return {"args": names, "required": max(0, n_required), "varargs": a.vararg is not None, "kwargs": a.kwarg is not None}


def main():
    tree = ast.parse(REF.read_text())
    api = {"classes": {}, "functions": {}}
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            methods = {}
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name not in ("__buffer__", "__release_buffer__"):
                    # overloads: keep the longest signature
                    sig = arg_names(f)
                    if f.name not in methods or len(sig["args"]) > len(methods[f.name]["args"]):
                        methods[f.name] = sig
            api["classes"][node.name] = {"bases": [ast.unparse(b) for b in node.bases], "methods": methods}
        elif isinstance(node, ast.FunctionDef):
            api["functions"][node.name] = arg_names(node)
    OUT.write_text(json.dumps(api, indent=1, sort_keys=True) + "\n")
    print(f"{len(api['classes'])} classes, {sum(len(c['methods']) for c in api['classes'].values())} methods, {len(api['functions'])} functions")


if __name__ == "__main__":
    main()
