"""Host-side index building and query-term resolution (the callers either side of the hot path).

What the reference does here lives in the write side and in un-vendored crates
(`nlp::TextParser` tokenise+stem, `StringStorage` FST term dictionary with prefix /
Levenshtein expansion — SURVEY.md §8a9, §8f-3).  This module is the minimal stand-in needed
to drive the GPU path with real text: a lower-case alphanumeric tokenizer (no stemming), a
sorted term dictionary per field (term id = rank, so a prefix is a contiguous id range), and
a CSR builder producing `StringIndexData` (what compact() lays out, string_field.rs:186-191).

Unpinned constant: EXACT_MATCH_BOOST.  The reference's value lives in oramacore_fields 0.2.0
and is not visible; only its effect is pinned (an exact term outranks a prefix expansion,
src/tests/boost_integration.rs:449-490).  Any value > 1 reproduces that; 2.0 is assumed.
"""
from __future__ import annotations

import bisect
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from .types import FieldPostings, StringIndexData, TextQuery

EXACT_MATCH_BOOST = 2.0
_TOKEN_RE = re.compile(r"[0-9a-z]+")


def tokenize(text: str) -> List[str]:
    return _TOKEN_RE.findall(text.lower())


def _levenshtein_le(a: str, b: str, k: int) -> bool:
    if abs(len(a) - len(b)) > k:
        return False
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i] + [0] * len(b)
        for j, cb in enumerate(b, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb))
        if min(cur) > k:
            return False
        prev = cur
    return prev[-1] <= k


class HostStringIndex:
    """Accumulates StringFieldStorage::insert(doc_id, IndexedValue{field_length, terms}) calls
    (string_field.rs:155-177) per field and lays them out as CSR on commit()."""

    def __init__(self, field_names: Sequence[str]):
        self.field_names = list(field_names)
        self._docs: Dict[int, Dict[str, List[str]]] = {}
        self.terms: List[List[str]] = []      # per field: sorted vocabulary
        self.data: Optional[StringIndexData] = None

    def insert(self, doc_id: int, doc: Dict[str, str]):
        self._docs[int(doc_id)] = {f: tokenize(doc.get(f, "") or "") for f in self.field_names}

    def delete(self, doc_id: int):
        self._docs.pop(int(doc_id), None)

    def commit(self) -> StringIndexData:
        doc_ids = sorted(self._docs)
        row_of = {d: r for r, d in enumerate(doc_ids)}
        fields = []
        self.terms = []
        for f in self.field_names:
            vocab = sorted({t for d in doc_ids for t in self._docs[d][f]})
            tid = {t: i for i, t in enumerate(vocab)}
            self.terms.append(vocab)
            per_term: List[List] = [[] for _ in vocab]
            lens = []
            for d in doc_ids:
                toks = self._docs[d][f]
                if toks:
                    lens.append(len(toks))
                counts: Dict[str, int] = {}
                for t in toks:
                    counts[t] = counts.get(t, 0) + 1
                for t, c in counts.items():
                    per_term[tid[t]].append((row_of[d], min(c, 65535), min(len(toks), 65535)))
            offs = np.zeros(len(vocab) + 1, np.uint64)
            rows, tfs, fls = [], [], []
            for i, plist in enumerate(per_term):
                plist.sort()
                offs[i + 1] = offs[i] + np.uint64(len(plist))
                for (r, c, l) in plist:
                    rows.append(r); tfs.append(c); fls.append(l)
            avg = float(np.mean(lens)) if lens else 1.0
            fields.append(FieldPostings(avg, offs, np.asarray(rows, np.uint32), np.asarray(tfs, np.uint16),
                                        np.asarray(fls, np.uint16)))
        ident = doc_ids == list(range(len(doc_ids)))
        self.data = StringIndexData(fields, len(doc_ids), len(doc_ids),
                                    None if ident else np.asarray(doc_ids, np.uint64))
        return self.data

    # ---- query side: SearchParams{tokens, exact_match, boost, tolerance} (token_score.rs:235-242)
    def resolve(self, term: str, exact: bool = False, tolerance: Optional[int] = None,
                boost: Optional[Dict[str, float]] = None, properties: Optional[Sequence[str]] = None) -> TextQuery:
        toks = tokenize(term)
        if not toks:
            toks = [""]  # token_score.rs:206-209: the empty token matches every document
        props = [i for i, f in enumerate(self.field_names) if properties is None or f in properties]
        out = []
        for tok in toks:
            ents = []
            for fi in props:
                vocab = self.terms[fi]
                w = float((boost or {}).get(self.field_names[fi], 1.0))
                if exact:
                    i = bisect.bisect_left(vocab, tok)
                    if i < len(vocab) and vocab[i] == tok:
                        ents.append((fi, i, w * EXACT_MATCH_BOOST))
                elif tolerance is not None:
                    for i, v in enumerate(vocab):
                        if v == tok:
                            ents.append((fi, i, w * EXACT_MATCH_BOOST))
                        elif _levenshtein_le(tok, v, tolerance) or v.startswith(tok):
                            ents.append((fi, i, w))
                else:
                    lo = bisect.bisect_left(vocab, tok)
                    i = lo
                    while i < len(vocab) and vocab[i].startswith(tok):
                        ents.append((fi, i, w * (EXACT_MATCH_BOOST if vocab[i] == tok else 1.0)))
                        i += 1
            out.append(ents)
        return TextQuery.from_tokens(out)
