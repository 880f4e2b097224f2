"""Loader / refresher of the device-resident stores from the reference's own write-operation stream.

The reference's read side is fed by `Index::update_data(IndexWriteOperation)` (read/index/mod.rs:1436-1705):
  * `Index { doc_id, indexed_values }` — `document_count += 1`, then per value
      ScoreString2(field, IndexedValue{field_length: u16, terms: {term -> TermData{exact_positions, positions}}})
                                              -> StringFieldStorage::insert   (string_field.rs:155-177, mod.rs:1509-1515)
      FilterBool / FilterNumber / FilterString -> the filter fields the facets and filters read (mod.rs:1461-1497)
  * `IndexEmbedding { data: field -> [(doc_id, vectors)] }` -> EmbeddingFieldStorage::insert (mod.rs:1688-1698)
  * `DeleteDocuments { doc_ids }` — uncommitted deletes, excluded from every search at once (mod.rs:1346-1427)
and `commit` / `compact` lay the pending data out (`CURRENT` + `versions/<n>`, embedding_field.rs:91-95).

`IndexLoader.apply(op)` takes the same operations as plain dicts (the JSON shape of the reference's enum), resolves
terms to stable term ids through the native dictionary (oc_dict_*), and drives the C ABI: oc_str_insert /
oc_str_delete / oc_str_commit (snapshot swap: searches keep running on the previous version while a commit builds
the next), oc_emb_insert / oc_emb_delete (live).  `refresh_facets()` lays the accumulated filter fields out for
oc_search_facets.  tf of a term = number of positions (exact + stemmed), as StringStorage counts them."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .engine import (Context, EmbeddingFieldStorage, FacetStore, StringFieldStorage, TermDictionary, TokenScoreContext)


class IndexLoader:
    def __init__(self, ctx: Context, string_fields: Sequence[str], embedding_model: Optional[str] = None,
                 embedding_dim: Optional[int] = None, bool_fields: Sequence[str] = (), number_fields: Sequence[str] = (),
                 string_filter_fields: Sequence[str] = ()):
        self.ctx = ctx
        self.string_fields = list(string_fields)
        self.dict = TermDictionary(max(len(self.string_fields), 1))
        self.strs = StringFieldStorage.empty(ctx, max(len(self.string_fields), 1))
        self.emb = EmbeddingFieldStorage(ctx, embedding_model or "BGESmall", dim=embedding_dim) if (embedding_model or embedding_dim) else None
        self._bool = {f: ({}) for f in bool_fields}            # field -> {doc: bool}
        self._num = {f: ({}) for f in number_fields}           # field -> {doc: [numbers]}
        self._strf = {f: ({}) for f in string_filter_fields}   # field -> {doc: [keys]}
        self.document_count = 0
        self.max_doc_id = -1
        self._deleted: set = set()
        self.facets: Optional[FacetStore] = None

    # ---- Index::update_data
    def apply(self, op: Dict) -> None:
        kind = op["type"]
        if kind == "Index":
            d = int(op["doc_id"])
            self.document_count += 1
            self.max_doc_id = max(self.max_doc_id, d)
            self._deleted.discard(d)
            for v in op["indexed_values"]:
                t = v["type"]
                if t == "ScoreString2":
                    fi = self.string_fields.index(v["field"])
                    terms = v["terms"]                                  # {term: {"exact_positions": [...], "positions": [...]}}
                    names = list(terms)
                    ids = self.dict.add_terms(fi, names) if names else np.zeros(0, np.uint32)
                    tf = {int(i): max(1, len(terms[n].get("exact_positions", ())) + len(terms[n].get("positions", ()))) for i, n in zip(ids, names)}
                    self.strs.insert(d, fi, int(v["field_length"]), tf)
                elif t == "FilterBool":
                    self._bool[v["field"]][d] = bool(v["value"])
                elif t == "FilterNumber":
                    self._num[v["field"]].setdefault(d, []).append(float(v["value"]))
                elif t == "FilterString":
                    self._strf[v["field"]].setdefault(d, []).append(str(v["value"]))
                else:
                    raise ValueError(f"unsupported indexed value {t!r} (outside the search hot path)")
        elif kind == "IndexEmbedding":
            for d, vectors in op["data"]:
                self.max_doc_id = max(self.max_doc_id, int(d))
                self.emb.insert(int(d), vectors)
        elif kind == "DeleteDocuments":
            ids = [int(x) for x in op["doc_ids"]]
            self.strs.delete(ids)
            if self.emb is not None:
                self.emb.delete(ids)
            for d in ids:
                if d not in self._deleted:
                    self._deleted.add(d)
                    self.document_count -= 1
                for m in list(self._bool.values()) + list(self._num.values()) + list(self._strf.values()):
                    m.pop(d, None)
        else:
            raise ValueError(f"unsupported operation {kind!r}")

    def apply_all(self, ops: Iterable[Dict]) -> None:
        for op in ops:
            self.apply(op)

    def commit(self) -> None:
        """ReadSide::commit -> field compact(): publish the next snapshot of the string store (searches on the
        previous one keep running meanwhile) and refresh the facet layout."""
        self.strs.commit()
        # N of the idf is Index::document_count (mod.rs:1460: +1 per Index op, also for documents without string fields)
        self.strs.set_global(max(self.document_count, 0))
        self.refresh_facets()

    def refresh_facets(self) -> None:
        if not (self._bool or self._num or self._strf):
            return
        if self.facets is not None:
            self.facets.close()
        st = FacetStore(self.ctx, self.max_doc_id + 2)
        for f, m in self._bool.items():
            st.add_bool_field(f, [d for d, b in m.items() if b], [d for d, b in m.items() if not b])
        for f, m in self._num.items():
            docs = [d for d, vs in m.items() for _ in vs]
            vals = [x for vs in m.values() for x in vs]
            st.add_number_field(f, docs, vals)
        for f, m in self._strf.items():
            keys: Dict[str, List[int]] = {}
            for d, ks in m.items():
                for k in ks:
                    keys.setdefault(k, []).append(d)
            st.add_string_field(f, {k: keys[k] for k in sorted(keys)})
        self.facets = st

    def context(self) -> TokenScoreContext:
        return TokenScoreContext(self.ctx, self.emb, self.strs)

    def resolve(self, texts: Sequence[str], **kw):
        """token_score.rs:196-209 + the FST expansion: the packed query arrays oc_search takes."""
        return self.dict.resolve_batch(list(texts), **kw)

    def close(self):
        for x in (self.facets, self.emb, self.strs, self.dict):
            if x is not None:
                x.close()
