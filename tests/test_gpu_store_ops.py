"""String store op stream on the GPU: insert / re-insert / delete in call order, atomic commit, published
snapshot versions, and searches that keep running while a commit is in flight.

Reference semantics: StringFieldStorage::insert / delete / compact (string_field.rs:155-191), applied by
Index::update_data in op order (index/mod.rs:1436-1705); CURRENT + versions/<n> (embedding_field.rs:91-95)."""
import threading

import numpy as np
import pytest

import oramacore_b200 as ob
from oramacore_b200.types import TextQuery

pytestmark = pytest.mark.gpu


def _search(ctx, strs, term, limit=10):
    return ob.search(ctx, None, strs, "fulltext", texts=[TextQuery.single_terms([term])], limit=limit)[0]


def test_ops_apply_in_order(gpu_ctx):
    strs = ob.StringFieldStorage.empty(gpu_ctx, 1)
    v0 = strs.info()["version"]
    strs.insert(1, 0, 4, {0: 1, 1: 3})
    strs.insert(2, 0, 2, {1: 2})
    # same document twice before ONE commit, sharing term 1: the last insert wins (no duplicate posting, no union)
    strs.insert(1, 0, 3, {1: 1, 2: 2})
    # insert(X), delete(X), commit(): X never becomes searchable
    strs.insert(7, 0, 5, {1: 5})
    strs.delete(7)
    # delete(Y) then insert(Y): Y is a new document
    strs.delete(9)
    strs.insert(9, 0, 1, {2: 1})
    assert strs.info()["pending_postings"] == 7
    strs.commit()
    info = strs.info()
    assert info["version"] == v0 + 1 and info["pending_postings"] == 0 and info["total_documents"] == 3
    h = _search(gpu_ctx, strs, 1)
    assert sorted(h.doc_ids.tolist()) == [1, 2] and h.count == 2
    assert _search(gpu_ctx, strs, 0).count == 0           # doc 1's first version (term 0) was replaced
    assert sorted(_search(gpu_ctx, strs, 2).doc_ids.tolist()) == [1, 9]
    # delete of a committed document takes effect at once (tombstone), before any commit
    strs.delete(2)
    assert sorted(_search(gpu_ctx, strs, 1).doc_ids.tolist()) == [1]
    strs.commit()
    assert strs.info()["total_documents"] == 2
    strs.close()


def test_failed_commit_changes_nothing(gpu_ctx):
    strs = ob.StringFieldStorage.empty(gpu_ctx, 1)
    strs.insert(1, 0, 2, {0: 1, 1: 1})
    strs.commit()
    before = _search(gpu_ctx, strs, 1)
    v = strs.info()["version"]
    # a term listed twice in one insert is a caller error: the commit must fail and leave the store as it was
    t = np.asarray([1, 1], np.uint32)
    f = np.asarray([1, 2], np.uint16)
    import ctypes as C
    from oramacore_b200._lib import check, lib
    check(lib().oc_str_insert(strs._h, 0, 5, 2, 2, t.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p)))
    with pytest.raises(ob.OcError):
        strs.commit()
    assert strs.info()["version"] == v
    after = _search(gpu_ctx, strs, 1)
    assert after.count == before.count and np.array_equal(after.doc_ids, before.doc_ids) and np.array_equal(after.scores, before.scores)
    strs.close()


def test_searches_run_while_a_commit_is_in_flight(gpu_ctx, orc):
    # a big second batch makes the commit long enough for searches to overlap it; every search must see
    # either the old or the new snapshot, never a mixture, and the final state must equal a bulk load
    from oramacore_b200 import synth
    n, vocab = 120000, 2000
    data = synth.make_text_corpus(n, vocab, seed=31)
    f = data.fields[0]
    term_of = np.repeat(np.arange(vocab, dtype=np.uint32), np.diff(f.term_offsets.astype(np.int64)))
    order = np.argsort(f.post_row, kind="stable")
    rows, terms, tfs, lens = f.post_row[order], term_of[order], f.post_tf[order], f.post_len[order]
    starts = np.searchsorted(rows, np.arange(n + 1))
    strs = ob.StringFieldStorage.empty(gpu_ctx, 1)

    def feed(a, b):
        for d in range(a, b):
            lo, hi = starts[d], starts[d + 1]
            if hi > lo:
                strs.insert(d, 0, int(lens[lo]), dict(zip(terms[lo:hi].tolist(), tfs[lo:hi].tolist())))

    half = n // 2
    feed(0, half)
    strs.commit()
    probe = TextQuery.single_terms([3])
    old = ob.search(gpu_ctx, None, strs, "fulltext", texts=[probe], limit=10)[0]
    feed(half, n)
    seen, stop = [], threading.Event()

    def searcher():
        while not stop.is_set():
            seen.append(ob.search(gpu_ctx, None, strs, "fulltext", texts=[probe], limit=10)[0])

    th = threading.Thread(target=searcher)
    th.start()
    strs.commit()
    stop.set()
    th.join()
    new = ob.search(gpu_ctx, None, strs, "fulltext", texts=[probe], limit=10)[0]
    assert new.count > old.count
    for h in seen:
        assert h.count in (old.count, new.count)
        ref = old if h.count == old.count else new
        assert np.array_equal(h.doc_ids, ref.doc_ids) and np.array_equal(h.scores, ref.scores)
    # the incrementally built store answers exactly like the oracle on the bulk corpus
    ix = orc.StrIndex(data)
    sb = orc.SearchBatch(ix, None)
    sb.add(0, limit=10, text=probe)
    od, os_, on, oc = sb.run(1)
    assert new.count == int(oc[0]) and np.array_equal(new.scores, os_[0, :on[0]])
    strs.close()
