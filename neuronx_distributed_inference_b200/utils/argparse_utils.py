"""argparse helpers (reference utils/argparse_utils.py): ``--flag k=v k2=v2`` into dicts, comma lists, JSON-or-path values."""
from __future__ import annotations

import argparse
import json
import os


class StringOrIntegers(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        out = []
        for v in values if isinstance(values, list) else [values]:
            out.append(int(v) if str(v).lstrip("-").isdigit() else v)
        setattr(namespace, self.dest, out)


class KeyValueDict(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        d = {}
        for kv in values:
            k, v = kv.split("=", 1)
            try:
                d[k] = json.loads(v)
            except json.JSONDecodeError:
                d[k] = v
        setattr(namespace, self.dest, d)


def json_or_path(v: str):
    if os.path.exists(v):
        with open(v) as f:
            return json.load(f)
    return json.loads(v)


def comma_ints(v: str):
    return [int(x) for x in v.split(",") if x.strip()]
