"""Compact binary form of a telescope YAML config.

The reference's benchmark scenes (``configs/HESS/CT3.yaml``, ``CT5.yaml``) are ~0.5 MB of YAML.
``pack_config`` stores the same numbers losslessly (float64, as parsed) in a small ``.npz`` so the
scenes can travel with this repository; ``unpack_config`` rebuilds the exact config dict the YAML
loader would have produced, so ``build_telescope`` sees identical input either way.
"""
from __future__ import annotations

import io
import json
from pathlib import Path

import numpy as np

_DATA = Path(__file__).resolve().parent.parent / "data"


def pack_config(config: dict, path) -> None:
    """Config dict (as from ``yaml.safe_load``) -> ``.npz``: bulky numeric lists become arrays,
    everything else stays in a JSON skeleton."""
    arrays: dict[str, np.ndarray] = {}

    def strip(node, where):
        if isinstance(node, dict):
            return {k: strip(v, f"{where}/{k}") for k, v in node.items()}
        if isinstance(node, list):
            if node and all(isinstance(x, dict) for x in node) and where in ("/mirrors", "/obstructions"):
                return _pack_records(node, where, arrays)
            if len(node) > 16 and all(isinstance(x, (int, float)) for x in node):
                arrays[where] = np.asarray(node, np.float64)
                return {"__array__": where}
            return [strip(v, f"{where}/{i}") for i, v in enumerate(node)]
        return node

    skeleton = strip(config, "")
    buf = io.BytesIO()
    np.savez_compressed(buf, __skeleton__=np.frombuffer(json.dumps(skeleton).encode(), np.uint8), **arrays)
    Path(path).write_bytes(buf.getvalue())


def _pack_records(records, where, arrays):
    """List of homogeneous-ish dicts -> per-key arrays grouped by record 'shape'."""
    groups: dict[str, list] = {}
    order = []
    for i, r in enumerate(records):
        flat = _flatten(r)
        sig = json.dumps([(k, (len(v) if isinstance(v, list) else None) if not isinstance(v, str) else "s")
                          for k, v in flat])
        groups.setdefault(sig, []).append((i, flat))
        order.append(sig)
    out = {"__records__": where, "groups": []}
    for gi, (sig, items) in enumerate(groups.items()):
        keys = [k for k, _ in items[0][1]]
        meta = {"index": f"{where}#{gi}/__index__", "fields": []}
        arrays[meta["index"]] = np.asarray([i for i, _ in items], np.int64)
        for ki, k in enumerate(keys):
            vals = [fl[ki][1] for _, fl in items]
            if isinstance(vals[0], str):
                meta["fields"].append({"key": k, "str": vals})
            else:
                name = f"{where}#{gi}/{k}"
                arrays[name] = np.asarray(vals, np.float64)
                meta["fields"].append({"key": k, "array": name,
                                       "int": all(isinstance(v, int) and not isinstance(v, bool) for v in vals)})
        out["groups"].append(meta)
    return out


def _flatten(d, prefix=""):
    items = []
    for k, v in d.items():
        key = f"{prefix}{k}"
        if isinstance(v, dict):
            items += _flatten(v, key + ".")
        elif isinstance(v, list) and v and isinstance(v[0], list):
            items.append((key + "[]", [x for row in v for x in row] + [len(v[0])]))
        else:
            items.append((key, v))
    return items


def _unflatten(pairs):
    out: dict = {}
    for key, v in pairs:
        nested = key.endswith("[]")
        if nested:
            key = key[:-2]
            ncol = int(v[-1])
            body = v[:-1]
            v = [body[i:i + ncol] for i in range(0, len(body), ncol)]
        parts = key.split(".")
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


def unpack_config(path) -> dict:
    """Inverse of ``pack_config``."""
    z = np.load(path, allow_pickle=False)
    skeleton = json.loads(bytes(z["__skeleton__"]).decode())

    def build(node):
        if isinstance(node, dict):
            if "__array__" in node:
                return z[node["__array__"]].tolist()
            if "__records__" in node:
                n = sum(len(z[g["index"]]) for g in node["groups"])
                recs = [None] * n
                for g in node["groups"]:
                    idx = z[g["index"]].tolist()
                    cols = []
                    for f in g["fields"]:
                        if "str" in f:
                            cols.append(f["str"])
                        else:
                            a = z[f["array"]]
                            vals = a.astype(np.int64).tolist() if f.get("int") else a.tolist()
                            cols.append(vals)
                    for j, i in enumerate(idx):
                        recs[i] = _unflatten([(f["key"], c[j]) for f, c in zip(g["fields"], cols)])
                return recs
            return {k: build(v) for k, v in node.items()}
        if isinstance(node, list):
            return [build(v) for v in node]
        return node

    return build(skeleton)


def packaged_config(name: str) -> Path:
    """Path of a scene shipped with the package (``'CT3'``, ``'CT5'``)."""
    p = _DATA / f"{name}.npz"
    if not p.exists():
        raise FileNotFoundError(f"no packaged scene {name!r} (looked for {p})")
    return p


def load_packed_config(name_or_path) -> dict:
    p = Path(name_or_path)
    return unpack_config(p if p.suffix == ".npz" and p.exists() else packaged_config(str(name_or_path)))
