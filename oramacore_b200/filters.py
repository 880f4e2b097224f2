"""FilterResult -> DocumentId bitmap (the host half — the device half is
`rows_ok_kernel` / `emb_apply_filter_kernel`, which turn this bitmap into per-row masks).

The reference hands the scorers a `FilterResult<DocumentId>`: a tree of And / Or / Not over plain
doc-id sets (`filter.rs:344-392`), consulted through `contains(doc)` once per candidate
(`embedding_field.rs:54-61`, `string_field.rs:66-69`).  Document ids are sequential per collection
(`write/collection_document_storage.rs:72-77`), so the same predicate is a bitmap over
`[0, n_bits)`: bit d set <=> `filter.contains(d)`.  `execute_filter` restates
`FilterContext::execute_filter` (`filter.rs:344-392`): no filter and no uncommitted deletes -> None;
no filter -> NOT(deleted); otherwise AND(filter, NOT(deleted)).  (The reference's plain sets may be
Bloom-backed, i.e. `contains` can have false positives; the bitmap is exact.)"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Union

import numpy as np


@dataclass(frozen=True)
class Ids:
    """FilterResult::Filter(PlainFilterResult): an explicit set of DocumentIds."""
    doc_ids: Sequence[int]


@dataclass(frozen=True)
class And:
    a: "FilterExpr"
    b: "FilterExpr"


@dataclass(frozen=True)
class Or:
    a: "FilterExpr"
    b: "FilterExpr"


@dataclass(frozen=True)
class Not:
    a: "FilterExpr"


FilterExpr = Union[Ids, And, Or, Not]


def _words(n_bits: int) -> int:
    return (int(n_bits) + 63) // 64


def _tail_mask(bits: np.ndarray, n_bits: int) -> np.ndarray:
    r = int(n_bits) & 63
    if r and bits.size:
        bits[-1] &= np.uint64((1 << r) - 1)      # ids >= n_bits do not exist: keep the padding bits clear
    return bits


def to_bitmap(expr: FilterExpr, n_bits: int) -> np.ndarray:
    """uint64 words, bit d of word d/64 set <=> expr.contains(d), for d in [0, n_bits)."""
    if isinstance(expr, Ids):
        bits = np.zeros(_words(n_bits), np.uint64)
        ids = np.asarray(list(expr.doc_ids) if not isinstance(expr.doc_ids, np.ndarray) else expr.doc_ids, np.uint64)
        ids = ids[ids < np.uint64(n_bits)]
        np.bitwise_or.at(bits, (ids >> np.uint64(6)).astype(np.int64), np.uint64(1) << (ids & np.uint64(63)))
        return bits
    if isinstance(expr, And):
        return to_bitmap(expr.a, n_bits) & to_bitmap(expr.b, n_bits)
    if isinstance(expr, Or):
        return to_bitmap(expr.a, n_bits) | to_bitmap(expr.b, n_bits)
    if isinstance(expr, Not):
        return _tail_mask(~to_bitmap(expr.a, n_bits), n_bits)
    raise TypeError(f"not a filter expression: {expr!r}")


def execute_filter(where: Optional[FilterExpr], uncommitted_deleted: Sequence[int], n_bits: int) -> Optional[np.ndarray]:
    """FilterContext::execute_filter (filter.rs:344-392) as a bitmap; None == no filtering needed."""
    deleted = list(uncommitted_deleted)
    if where is None:
        return None if not deleted else to_bitmap(Not(Ids(deleted)), n_bits)
    return to_bitmap(where if not deleted else And(where, Not(Ids(deleted))), n_bits)


def contains(bits: np.ndarray, doc_id: int) -> bool:
    w = int(doc_id) >> 6
    return w < bits.size and bool((int(bits[w]) >> (int(doc_id) & 63)) & 1)
