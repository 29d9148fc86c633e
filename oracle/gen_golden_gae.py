"""Generate tests/golden/gae_golden.npz by EXECUTING the reference's own `discount` and
`generalized_advantage_estimate` / `value_target_estimate` (pure numpy) from their source files.

The modules themselves cannot be imported (top-level `import tensorflow`), so the function definitions
are lifted out of the source with `ast` and executed unmodified.  Run in the build container only
(`/root/reference` does not exist on the GPU box):  python oracle/gen_golden_gae.py
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lift(path, names, cls=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(picked) == len(names), (path, names)
    return ast.Module(body=picked, type_ignores=[])


def main():
    ns = {"np": np}
    exec(compile(lift(os.path.join(REF, "networks/utils.py"), ["discount"]), "ref_utils", "exec"), ns)
    exec(compile(lift(os.path.join(REF, "networks/actor_critic/a2c.py"),
                      ["generalized_advantage_estimate", "value_target_estimate"], cls="A2CNetwork"), "ref_a2c", "exec"), ns)
    rng = np.random.RandomState(33406)
    out = {}
    for i, (T, gamma, lambd) in enumerate([(1, 0.99, 0.95), (7, 0.99, 0.95), (368, 0.95, 0.95), (1000, 0.99, None), (64, 0.9, 1.0)]):
        obj = types.SimpleNamespace(gamma=gamma, gae_gamma=None if lambd is None else gamma * lambd)
        reward = rng.randn(T).astype(np.float32)
        value = (rng.randn(T + 1) * 3).astype(np.float32)
        adv = ns["generalized_advantage_estimate"](obj, list(reward), list(value))
        adv = np.asarray(adv)
        tgt = ns["value_target_estimate"](obj, value[:-1], adv)
        out[f"c{i}_reward"], out[f"c{i}_value"] = reward, value
        out[f"c{i}_gamma"], out[f"c{i}_gae_gamma"] = np.float64(gamma), np.float64(obj.gae_gamma or 0.0)
        out[f"c{i}_adv"], out[f"c{i}_adv_dtype"] = adv.astype(np.float64), str(adv.dtype)
        out[f"c{i}_target"] = np.asarray(tgt, dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gae_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("adv") or k.endswith("dtype")})


if __name__ == "__main__":
    main()
