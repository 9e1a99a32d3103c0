"""Compile a chromadb-style ``where`` filter into what the fused scan kernels take.

In chromadb a ``where`` clause is resolved against the metadata store into an id allow-list that is
handed to the vector index.  Here the same clause becomes

* ``ns_code`` / ``label_code``: equality on the two columns that live next to the vectors on the
  device (namespace -- written by codd_dal/metrics/metrics_semantic_metadata_store.py:232 -- and
  one configurable label field), compared inside the scan kernel, and/or
* a row bitmask (one bit per row, LSB-first in u32 words) for everything else, evaluated on the host
  over the stored metadata and AND-ed inside the kernel.

Supported operators: implicit equality, ``$eq $ne $gt $gte $lt $lte $in $nin`` and ``$and`` /
``$or`` (chromadb's metadata filter grammar).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Optional, Sequence

import numpy as np

ANY = 0xFFFFFFFF

_CMP: dict[str, Callable[[Any, Any], bool]] = {
    "$eq": lambda a, b: a == b,
    "$ne": lambda a, b: a != b,
    "$gt": lambda a, b: a is not None and a > b,
    "$gte": lambda a, b: a is not None and a >= b,
    "$lt": lambda a, b: a is not None and a < b,
    "$lte": lambda a, b: a is not None and a <= b,
    "$in": lambda a, b: a in b,
    "$nin": lambda a, b: a not in b,
}


def validate_where(where: Any) -> None:
    if not isinstance(where, dict):
        raise ValueError(f"where must be a dict, got {type(where).__name__}")
    if len(where) != 1:
        raise ValueError(f"where must have exactly one operator or field, got {list(where)}")
    (key, val), = where.items()
    if key in ("$and", "$or"):
        if not isinstance(val, (list, tuple)) or len(val) < 2:
            raise ValueError(f"{key} needs a list of at least two clauses")
        for sub in val:
            validate_where(sub)
        return
    if key.startswith("$"):
        raise ValueError(f"unknown logical operator {key}")
    if isinstance(val, dict):
        if len(val) != 1:
            raise ValueError(f"field {key!r}: exactly one comparison operator expected")
        (op, operand), = val.items()
        if op not in _CMP:
            raise ValueError(f"field {key!r}: unknown operator {op}")
        if op in ("$in", "$nin") and not isinstance(operand, (list, tuple)):
            raise ValueError(f"{op} needs a list")
    elif not isinstance(val, (str, int, float, bool)):
        raise ValueError(f"field {key!r}: unsupported operand type {type(val).__name__}")


def evaluate(where: dict, metadata: Optional[dict]) -> bool:
    """Does one row's metadata satisfy `where`?"""
    (key, val), = where.items()
    if key == "$and":
        return all(evaluate(s, metadata) for s in val)
    if key == "$or":
        return any(evaluate(s, metadata) for s in val)
    md = metadata or {}
    if isinstance(val, dict):
        (op, operand), = val.items()
    else:
        op, operand = "$eq", val
    if key not in md:
        return op in ("$ne", "$nin")
    try:
        return bool(_CMP[op](md[key], operand))
    except TypeError:
        return False


def _as_equality(clause: dict):
    """-> (field, value) when the clause is a plain equality, else None."""
    (key, val), = clause.items()
    if key.startswith("$"):
        return None
    if isinstance(val, dict):
        (op, operand), = val.items()
        return (key, operand) if op == "$eq" else None
    return (key, val)


@dataclass
class CompiledWhere:
    ns_code: int = ANY
    label_code: int = ANY
    mask_bits: Optional[np.ndarray] = None   # bool [n], or None = every row
    empty: bool = False                      # provably matches nothing
    _words: Optional[np.ndarray] = None

    @property
    def mask_words(self) -> Optional[np.ndarray]:
        """The bitmask the kernels take: uint32 [ceil(n/32)], packed on first use."""
        if self.mask_bits is None:
            return None
        if self._words is None:
            self._words = pack_mask(self.mask_bits)
        return self._words

    @property
    def eligible(self) -> Optional[int]:
        return None if self.mask_bits is None else int(np.count_nonzero(self.mask_bits))


def pack_mask(bits: np.ndarray) -> np.ndarray:
    """bool [n] -> u32 words, bit (r & 31) of word (r >> 5) = bits[r]."""
    n = bits.shape[0]
    pad = (-n) % 32
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, dtype=bool)])
    return np.packbits(bits, bitorder="little").view(np.uint32)


_MISSING = object()


class MetadataColumns:
    """Dictionary-encoded metadata columns, one per field a `where` clause has touched, kept in step with the rows.

    A residual clause (anything but the fused namespace / label equalities) used to be evaluated by a Python loop
    over every row per query.  Here a field's values are coded once (code 0 = the row has no such key), a clause
    is evaluated once per DISTINCT value with the very comparison `evaluate` uses, and the per-row answer is one
    numpy gather `truth[codes]` -- identical results, milliseconds at a million rows.
    """

    def __init__(self):
        self._cols: dict[str, tuple[np.ndarray, list, dict]] = {}  # field -> (codes int32 [cap], values, code_of)

    def _code(self, col, value) -> int:
        codes, values, code_of = col
        key = (type(value).__name__, value)  # 1, 1.0 and True stay distinct values, as in `evaluate`
        c = code_of.get(key)
        if c is None:
            c = len(values)
            values.append(value)
            code_of[key] = c
        return c

    def column(self, field: str, metadatas: Sequence[Optional[dict]]):
        col = self._cols.get(field)
        n = len(metadatas)
        if col is None:
            # first touch of a field: one pass over the rows.  Values are pulled out with a list comprehension and coded
            # through a per-(class, value) memo, codes land in a Python list and reach numpy in one assignment -- a
            # third of the time of coding row by row into the array (1M rows: ~0.25 s instead of ~0.8 s).
            col = (np.zeros(max(n, 16), dtype=np.int32), [_MISSING], {})
            memo: dict = {}
            out = [0] * n
            for r, md in enumerate(metadatas):
                if md:
                    v = md.get(field, _MISSING)
                    if v is not _MISSING:
                        key = (v.__class__, v)
                        c = memo.get(key)
                        if c is None:
                            c = memo[key] = self._code(col, v)
                        out[r] = c
            col[0][:n] = out
            self._cols[field] = col
        elif col[0].shape[0] < n:  # rows appended since (set_row grows it too; this covers direct table edits)
            grown = np.zeros(max(n, 2 * col[0].shape[0]), dtype=np.int32)
            grown[: col[0].shape[0]] = col[0]
            col = (grown, col[1], col[2])
            self._cols[field] = col
        return col

    def set_row(self, row: int, metadata: Optional[dict]) -> None:
        """Row `row` now carries `metadata` (None = deleted): update every tracked column."""
        for field, col in list(self._cols.items()):
            if col[0].shape[0] <= row:
                grown = np.zeros(max(row + 1, 2 * col[0].shape[0]), dtype=np.int32)
                grown[: col[0].shape[0]] = col[0]
                col = (grown, col[1], col[2])
                self._cols[field] = col
            col[0][row] = self._code(col, metadata[field]) if metadata is not None and field in metadata else 0

    def evaluate_rows(self, where: dict, metadatas: Sequence[Optional[dict]]) -> np.ndarray:
        """bool [n]: `evaluate(where, metadatas[r])` for every row, vectorised."""
        n = len(metadatas)
        (key, val), = where.items()
        if key == "$and":
            out = np.ones(n, dtype=bool)
            for sub in val:
                out &= self.evaluate_rows(sub, metadatas)
            return out
        if key == "$or":
            out = np.zeros(n, dtype=bool)
            for sub in val:
                out |= self.evaluate_rows(sub, metadatas)
            return out
        if isinstance(val, dict):
            (op, operand), = val.items()
        else:
            op, operand = "$eq", val
        codes, values, _ = self.column(key, metadatas)
        truth = np.zeros(len(values), dtype=bool)
        truth[0] = op in ("$ne", "$nin")  # the row has no such key
        for c in range(1, len(values)):
            try:
                truth[c] = bool(_CMP[op](values[c], operand))
            except TypeError:
                truth[c] = False
        return truth[codes[:n]]


def compile_where(where: Optional[dict], metadatas: Sequence[Optional[dict]], ns_codes: dict, label_field: str,
                  label_codes: dict, alive: Optional[np.ndarray] = None,
                  columns: Optional[MetadataColumns] = None) -> CompiledWhere:
    """Split `where` into fused column predicates + a host-evaluated bitmask.

    metadatas: per-row metadata dicts (row order).  ns_codes / label_codes: value -> u32 code of the
    device columns.  alive: optional bool [n] of rows that exist (tombstones cleared).  columns: the collection's
    MetadataColumns (vectorised residual clauses); without it the clauses are evaluated row by row.
    """
    out = CompiledWhere()
    n = len(metadatas)
    residual: list[dict] = []
    if where:
        validate_where(where)
        (key, val), = where.items()
        clauses = list(val) if key == "$and" else [where]
        for cl in clauses:
            eq = _as_equality(cl)
            if eq is not None and eq[0] == "namespace" and out.ns_code == ANY:
                code = ns_codes.get(eq[1])
                if code is None:
                    out.empty = True
                    return out
                out.ns_code = code
            elif eq is not None and eq[0] == label_field and out.label_code == ANY:
                code = label_codes.get(eq[1])
                if code is None:
                    out.empty = True
                    return out
                out.label_code = code
            else:
                residual.append(cl)
    bits = None
    if residual:
        if columns is not None:
            bits = np.ones(n, dtype=bool)
            for cl in residual:
                bits &= columns.evaluate_rows(cl, metadatas)
        else:
            bits = np.fromiter((all(evaluate(cl, md) for cl in residual) for md in metadatas), dtype=bool, count=n)
    if alive is not None and not alive.all():
        bits = alive.copy() if bits is None else (bits & alive)
    if bits is not None:
        if not bits.any():
            out.empty = True
            return out
        out.mask_bits = bits
    return out
