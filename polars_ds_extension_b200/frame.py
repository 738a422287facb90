"""A minimal, Polars-shaped frame so the plugin expressions can be evaluated without a polars wheel.

Only what the lin_reg tests of the reference need: named columns (pyarrow arrays, nulls allowed), ``select``,
``filter`` / ``slice`` / ``limit``, and ``group_by(key).agg(expr)`` which — like Polars — gathers each
group's rows and calls the plugin symbol once per group.  ``group_by(...).agg(expr, fast=True)`` uses the additive
batched symbol ``pl_lr_by`` instead (one launch sequence for all groups).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from . import _harness


@dataclass(frozen=True)
class ColExpr:
    """Column reference with the few transformations expr_linear.py applies (cast, var, rechunk, alias)."""
    name: str
    cast_to: Optional[str] = None      # "f32" | "f64"
    agg: Optional[str] = None          # "var"
    alias_name: Optional[str] = None
    rechunked: bool = False
    shift_by: int = 0                  # pl.Expr.shift(n): rows move down by n, nulls enter at the top
    slice_offset: int = 0              # pl.Expr.slice(offset): drop the first `offset` rows (applied after the shift)

    def cast(self, dtype: str) -> "ColExpr":
        return replace(self, cast_to=dtype)

    def var(self) -> "ColExpr":
        return replace(self, agg="var")

    def alias(self, name: str) -> "ColExpr":
        return replace(self, alias_name=name)

    def rechunk(self) -> "ColExpr":
        return replace(self, rechunked=True)

    def shift(self, n: int = 1) -> "ColExpr":
        return replace(self, shift_by=self.shift_by + int(n))

    def slice(self, offset: int) -> "ColExpr":
        return replace(self, slice_offset=self.slice_offset + int(offset))

    @property
    def out_name(self) -> str:
        return self.alias_name if self.alias_name is not None else self.name


def col(name: str) -> ColExpr:
    return ColExpr(name)


@dataclass
class PluginExpr:
    """What ``polars.plugins.register_plugin_function`` would capture: symbol, input expressions, kwargs, flags."""
    symbol: str
    args: List[ColExpr]
    kwargs: Dict[str, Any]
    returns_scalar: bool = False
    changes_length: bool = False
    pass_name_to_apply: bool = True
    out_name: str = ""

    def alias(self, name: str) -> "PluginExpr":
        return replace(self, out_name=name)

    def to_polars(self):  # pragma: no cover - needs a polars wheel
        import polars as pl
        from pathlib import Path
        from polars.plugins import register_plugin_function

        def to_pl(c: ColExpr):
            e = pl.col(c.name)
            if c.cast_to:
                e = e.cast(pl.Float32 if c.cast_to == "f32" else pl.Float64)
            if c.agg == "var":
                e = e.var()
            if c.rechunked:
                e = e.rechunk()
            if c.shift_by:
                e = e.shift(c.shift_by)
            if c.slice_offset:
                e = e.slice(c.slice_offset)
            if c.alias_name:
                e = e.alias(c.alias_name)
            return e

        return register_plugin_function(
            plugin_path=Path(__file__).parent, args=[to_pl(a) for a in self.args], function_name=self.symbol,
            kwargs=self.kwargs, returns_scalar=self.returns_scalar, changes_length=self.changes_length,
            pass_name_to_apply=self.pass_name_to_apply,
        )


def _to_arrow(v) -> pa.ChunkedArray:
    if isinstance(v, pa.ChunkedArray):
        return v
    if isinstance(v, pa.Array):
        return pa.chunked_array([v])
    if isinstance(v, np.ndarray):
        return pa.chunked_array([pa.array(v)])
    return pa.chunked_array([pa.array(list(v))])


class Frame:
    def __init__(self, data: Dict[str, Any]):
        self.columns: Dict[str, pa.ChunkedArray] = {k: _to_arrow(v) for k, v in data.items()}
        lens = {len(v) for v in self.columns.values()}
        if len(lens) > 1:
            raise ValueError("columns have different lengths")
        self.height = lens.pop() if lens else 0

    def __len__(self) -> int:
        return self.height

    def __getitem__(self, name: str) -> pa.ChunkedArray:
        return self.columns[name]

    def with_columns(self, **cols) -> "Frame":
        d = dict(self.columns)
        d.update(cols)
        return Frame(d)

    def slice(self, offset: int, length: Optional[int] = None) -> "Frame":
        return Frame({k: v.slice(offset, length) for k, v in self.columns.items()})

    def limit(self, n: int) -> "Frame":
        return self.slice(0, n)

    def filter(self, mask) -> "Frame":
        m = pa.array(np.asarray(mask, dtype=bool))
        return Frame({k: v.filter(m) for k, v in self.columns.items()})

    def drop_nulls(self) -> "Frame":
        keep = np.ones(self.height, dtype=bool)
        for v in self.columns.values():
            keep &= np.asarray(pc.is_valid(v).to_numpy(zero_copy_only=False), dtype=bool)
        return self.filter(keep)

    # ---- expression evaluation -------------------------------------------------------------------------
    def _eval_arg(self, c: ColExpr) -> pa.ChunkedArray:
        a = self.columns[c.name]
        if c.cast_to:
            a = a.cast(pa.float32() if c.cast_to == "f32" else pa.float64())
        if c.agg == "var":
            vals = a.drop_null().to_numpy()
            v = float(np.var(vals.astype(np.float64), ddof=1)) if len(vals) > 1 else None
            a = pa.chunked_array([pa.array([v], type=a.type)])
        if c.rechunked:
            a = pa.chunked_array([a.combine_chunks()]) if a.num_chunks != 1 else a
        if c.shift_by:
            k = c.shift_by
            if k < 0:
                raise ValueError("only forward shifts are needed by the lin_reg callers")
            k = min(k, len(a))
            a = pa.chunked_array([pa.nulls(k, type=a.type)] + a.slice(0, len(a) - k).chunks, type=a.type)
        if c.slice_offset:
            a = a.slice(min(c.slice_offset, len(a)))
        return a

    def evaluate(self, e: PluginExpr) -> pa.Array:
        inputs = [self._eval_arg(a) for a in e.args]
        names = [a.out_name for a in e.args]
        return _harness.call_plugin(e.symbol, inputs, names, e.kwargs)

    def select(self, *exprs: PluginExpr) -> Dict[str, pa.Array]:
        return {(e.out_name or e.symbol): self.evaluate(e) for e in exprs}

    def group_by(self, key: str) -> "GroupBy":
        return GroupBy(self, key)


class GroupBy:
    def __init__(self, frame: Frame, key: str):
        self.frame = frame
        self.key = key

    def _groups(self):
        keys = self.frame.columns[self.key].to_numpy(zero_copy_only=False)
        uniq, first = np.unique(keys, return_index=True)
        order = np.argsort(first)              # maintain_order=True
        return keys, [uniq[i] for i in order]

    def agg(self, e: PluginExpr, fast: bool = False) -> Dict[str, Any]:
        keys, uniq = self._groups()
        if not fast:
            out = []
            for k in uniq:
                sub = self.frame.filter(keys == k)
                out.append(sub.evaluate(e))
            return {self.key: uniq, (e.out_name or e.symbol): out}
        # batched path: sort rows by key (stable), hand contiguous offsets to pl_lr_by
        base = e.symbol[:-4] if e.symbol.endswith("_f32") else e.symbol
        if base != "pl_lr":
            raise ValueError("fast group_by is only available for lin_reg coefficients")
        idx_of = {k: i for i, k in enumerate(uniq)}
        gid = np.fromiter((idx_of[k] for k in keys), dtype=np.int64, count=len(keys))
        order = np.argsort(gid, kind="stable")
        counts = np.bincount(gid, minlength=len(uniq))
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        take = pa.array(order)
        inputs = [pa.chunked_array([pa.array(offsets)])]
        names = ["__offsets__"]
        for a in e.args:
            inputs.append(pa.chunked_array([self.frame._eval_arg(a).combine_chunks().take(take)]))
            names.append(a.out_name)
        sym = "pl_lr_by" + ("_f32" if e.symbol.endswith("_f32") else "")
        res = _harness.call_plugin(sym, inputs, names, e.kwargs)
        return {self.key: uniq, (e.out_name or e.symbol): res}
