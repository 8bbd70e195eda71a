#!/usr/bin/env python
"""Records the config schema of the reference's model families on the B200 path (field names, order and literal defaults of
the dataclasses in mlx_vlm/models/<family>/config.py, read with `ast`) into tests/golden/config_schema_golden.json, so that
the product's config classes — which are generated from tables (models/config_schema.py), not written as dataclasses —
can be pinned field by field.  usage: python tests/golden/make_config_golden.py"""
import ast
import json
import os

REF = "/root/reference/mlx_vlm/models"
FAMILIES = ("llava", "llava_next", "idefics2", "idefics3", "smolvlm", "qwen2_5_vl", "qwen2_vl")


def default_of(node):
    if node is None:
        return "<required>"
    try:
        return ast.literal_eval(node)
    except Exception:
        pass
    # field(default_factory=lambda: <literal or Class()>)
    if isinstance(node, ast.Call) and getattr(node.func, "id", "") == "field":
        for kw in node.keywords:
            if kw.arg == "default_factory" and isinstance(kw.value, ast.Lambda):
                try:
                    return {"factory": ast.literal_eval(kw.value.body)}
                except Exception:
                    return {"factory": ast.unparse(kw.value.body)}
    return ast.unparse(node)


def main():
    out = {"_about": "dataclass fields of the reference's config.py files (ast; make_config_golden.py)"}
    for fam in FAMILIES:
        src = open(os.path.join(REF, fam, "config.py")).read()
        classes = {}
        for node in ast.parse(src).body:
            if isinstance(node, ast.ClassDef):
                classes[node.name] = [[st.target.id, default_of(st.value)] for st in node.body if isinstance(st, ast.AnnAssign)]
        out[fam] = classes
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_schema_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, {k: list(v) for k, v in out.items() if k != "_about"})


if __name__ == "__main__":
    main()
