"""Model-constant TOML files.

Same contract as the reference's ``mythos/input/toml.py:21-57``: string values are arithmetic expressions
(``"pi - 2.35"``, ``"296.15 * 0.1 / 300.0"``) evaluated to floats.  The reference goes through sympy
(32-digit evalf, rounded to float64); here a small AST walker evaluates the same grammar in float64
directly, which agrees to the last bit for every expression in the shipped defaults except possibly one ulp.
"""

from __future__ import annotations

import ast
import math
import operator
import tomllib
from pathlib import Path
from typing import Any

DEFAULTS_DIR = Path(__file__).resolve().parent.parent / "energy" / "defaults"

_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.Pow: operator.pow}
_NAMES = {"pi": math.pi, "e": math.e}


def _eval(node: ast.AST) -> float:
    if isinstance(node, ast.Expression):
        return _eval(node.body)
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return float(node.value)
    if isinstance(node, ast.Name) and node.id in _NAMES:
        return _NAMES[node.id]
    if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
        return _BIN[type(node.op)](_eval(node.left), _eval(node.right))
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = _eval(node.operand)
        return -v if isinstance(node.op, ast.USub) else v
    raise ValueError(f"unsupported expression node {ast.dump(node)}")


def parse_str(value: str) -> str | float:
    """A float, an arithmetic expression in ``pi``, or (unparseable) the string itself."""
    try:
        return float(value)
    except ValueError:
        try:
            return float(_eval(ast.parse(value, mode="eval")))
        except (ValueError, SyntaxError, ZeroDivisionError):
            return value


def parse_value(value: Any) -> Any:
    if isinstance(value, str):
        return parse_str(value)
    if isinstance(value, bool):
        return value
    if isinstance(value, (int, float)):
        return float(value)
    if isinstance(value, list):
        return [parse_value(v) for v in value]
    if isinstance(value, dict):
        return {k: parse_value(v) for k, v in value.items()}
    return value


def parse_toml(file_path: Path | str, key: str | None = None) -> dict[str, Any]:
    """Parse a TOML file into nested dicts of floats (optionally only the table ``key``)."""
    with Path(file_path).open("rb") as f:
        config = tomllib.load(f)
    if key is not None:
        if key not in config:
            raise ValueError(f"Missing entry {key} in TOML file")
        config = config[key]
    return parse_value(config)


def load_model_defaults(model: str) -> dict[str, Any]:
    """``{"energy": {...}, "simulation": {...}}`` of one of dna1 / dna2 / rna2 / na1."""
    path = DEFAULTS_DIR / f"{model}.toml"
    if not path.exists():
        raise ValueError(f"unknown model {model!r}")
    return parse_toml(path)
