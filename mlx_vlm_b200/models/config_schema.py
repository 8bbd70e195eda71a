"""Config classes from tables.

A checkpoint's `config.json` must load unchanged, so the field names, their order and their defaults are the reference's
(`mlx_vlm/models/<family>/config.py`) — that part is DATA.  Each family's config module keeps it as a small table

    name                      type                 default        (`-` = no default)

and `config_class` turns a table into a dataclass derived from `BaseModelConfig` (same `from_dict` filtering of unknown
keys); behaviour (kv-head defaults, derived sizes, validation, nested `from_dict`) is attached as plain functions.  The
tables are pinned field by field against the reference's dataclasses by tests/test_config_schema.py
(tests/golden/config_schema_golden.json)."""
from __future__ import annotations

import ast
import copy
import dataclasses
from dataclasses import field, make_dataclass
from typing import Any, Callable, Dict, List, Optional, Union  # noqa: F401  (names used by the table types)

from .base import BaseModelConfig

_TYPES = {"Optional": Optional, "List": List, "Dict": Dict, "Union": Union, "Any": Any, "int": int, "float": float,
          "str": str, "bool": bool, "list": list, "object": object}


def _parse(table: str):
    rows = []
    for line in table.strip().splitlines():
        line = line.split("#", 1)[0].strip()
        if not line:
            continue
        name, typ, default = line.split(None, 2)
        rows.append((name, eval(typ, dict(_TYPES)), default.strip()))   # noqa: S307  (our own literal tables)
    return rows


def config_class(name: str, module: str, table: str, post_init: Optional[Callable] = None,
                 members: Optional[Dict[str, Any]] = None, factories: Optional[Dict[str, Callable]] = None):
    """dataclass `name` (a BaseModelConfig) with the fields of `table`; `factories[field]()` builds a field's default"""
    fields = []
    for fname, typ, default in _parse(table):
        if factories and fname in factories:
            fields.append((fname, typ, field(default_factory=factories[fname])))
            continue
        if default == "-":
            fields.append((fname, typ))
            continue
        value = ast.literal_eval(default)
        if isinstance(value, (list, dict)):
            fields.append((fname, typ, field(default_factory=lambda v=value: copy.deepcopy(v))))
        else:
            fields.append((fname, typ, field(default=value)))
    ns = dict(members or {})
    if post_init is not None:
        ns["__post_init__"] = post_init
    cls = make_dataclass(name, fields, bases=(BaseModelConfig,), namespace=ns)
    cls.__module__ = module
    return cls


def nested_from_dict(**subs):
    """`ModelConfig.from_dict` of the families whose config.json nests `text_config` / `vision_config` / ...: nested dicts
    become their config classes, unknown keys are dropped, the argument is not modified."""
    def from_dict(cls, params):
        params = dict(params)
        for key, sub in subs.items():
            if isinstance(params.get(key), dict):
                params[key] = sub.from_dict(params[key])
        known = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in params.items() if k in known})
    return classmethod(from_dict)


def kv_heads_default(self):
    """`num_key_value_heads: null` means multi-head attention"""
    if self.num_key_value_heads is None:
        self.num_key_value_heads = self.num_attention_heads


def image_token_alias(self):
    """`image_token_index` falls back to `image_token_id` (Idefics2 / Idefics3 / SmolVLM)"""
    if self.image_token_index is None:
        self.image_token_index = self.image_token_id
