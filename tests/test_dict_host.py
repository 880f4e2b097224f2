"""Native term dictionary / batch query resolution (oc_dict_*, csrc/dict.h) against the Python
restatement in hostindex.resolve (same (field, term, weight) lists, same order) on random vocabularies
and on the tiny corpora of the reference's own tests (prefix: fulltext_search.rs:603-757, tolerance:
:956-1018, exact vs prefix boost: boost_integration.rs:449-490).  Host only: runs without a GPU."""
import time

import numpy as np
import pytest

import oramacore_b200 as ob
from oramacore_b200.hostindex import HostStringIndex


def _index(docs, fields=("text",)):
    h = HostStringIndex(fields)
    for d, doc in docs:
        h.insert(d, doc)
    h.commit()
    d = ob.TermDictionary(len(fields))
    for fi in range(len(fields)):
        ids = d.add_terms(fi, h.terms[fi])
        assert ids.tolist() == list(range(len(h.terms[fi])))      # sorted vocabulary: id == rank, like hostindex
    return h, d


def _same(q_native, q_host):
    assert np.array_equal(q_native.token_term_offsets, q_host.token_term_offsets), (q_native.token_term_offsets, q_host.token_term_offsets)
    assert np.array_equal(q_native.term_field, q_host.term_field)
    assert np.array_equal(q_native.term_id, q_host.term_id)
    assert np.array_equal(q_native.term_weight, q_host.term_weight)


def test_reference_pin_corpora():
    h, d = _index([(1, {"text": "Main Street"}), (2, {"text": "Maple Avenue"}), (3, {"text": "Another Street"})])
    for term, kw in (("Mxin", dict(tolerance=1)), ("Msple", dict(tolerance=1)), ("str", {}), ("street", dict(exact=True)),
                     ("", {}), ("ma", {}), ("zzz", {}), ("Main Street", dict(tolerance=2))):
        got = d.resolve_batch([term], **kw).query(0)
        _same(got, h.resolve(term, **kw))
    # tolerance 1: "mxin" reaches "main" only (fulltext_search.rs:956-1018)
    q = d.resolve_batch(["Mxin"], tolerance=1).query(0)
    assert [h.terms[0][i] for i in q.term_id] == ["main"]
    # exact term outranks its prefix expansion (boost_integration.rs:449-490)
    h2, d2 = _index([(1, {"text": "serve"}), (2, {"text": "server"})])
    q = d2.resolve_batch(["serve"]).query(0)
    w = dict(zip([h2.terms[0][i] for i in q.term_id], q.term_weight.tolist()))
    assert w["serve"] > w["server"]
    d.close(); d2.close()


def test_random_vocabulary_matches_hostindex_all_modes():
    rng = np.random.default_rng(5)
    alpha = "abcde"
    words = sorted({"".join(rng.choice(list(alpha), size=int(rng.integers(1, 7)))) for _ in range(3000)})
    docs = [(i, {"title": " ".join(rng.choice(words, size=3)), "body": " ".join(rng.choice(words, size=8))}) for i in range(400)]
    h, d = _index(docs, fields=("title", "body"))
    queries = [" ".join(rng.choice(words, size=int(rng.integers(1, 4)))) for _ in range(40)] + ["", "a", "zz", "abcdeabcde"]
    for kw in (dict(), dict(exact=True), dict(tolerance=1), dict(tolerance=2), dict(tolerance=0)):
        for boost, props in ((None, None), ({"title": 2.5}, None), (None, ["body"])):
            nb = None if boost is None else [boost.get(f, 1.0) for f in h.field_names]
            npp = None if props is None else [h.field_names.index(f) for f in props]
            batch = d.resolve_batch(queries, boost=nb, properties=npp, **kw)
            for i, q in enumerate(queries):
                _same(batch.query(i), h.resolve(q, boost=boost, properties=props, **kw))
    d.close()


def test_stable_ids_incremental_terms_and_stemmer_hook():
    d = ob.TermDictionary(1)
    a = d.add_terms(0, ["pear", "apple", "applesauce"])
    assert a.tolist() == [0, 1, 2]
    b = d.add_terms(0, ["apple", "banana", "app"])            # known term keeps its id, new ones append
    assert b.tolist() == [1, 3, 4] and d.size(0) == 5 and d.lookup(0, "banana") == 3 and d.lookup(0, "kiwi") is None
    q = d.resolve_batch(["app"]).query(0)                     # lexicographic emission order, stable ids
    assert q.term_id.tolist() == [4, 1, 2] and q.term_weight.tolist() == [2.0, 1.0, 1.0]
    # stems are flattened after their original unless exact (token_score.rs:196-204)
    d.set_stemmer(lambda t: t[:-1] if t.endswith("s") else None)
    q = d.resolve_batch(["pears"]).query(0)
    assert q.n_tokens == 2 and q.token_term_offsets.tolist() == [0, 0, 1] and q.term_id.tolist() == [0]
    q = d.resolve_batch(["pears"], exact=True).query(0)
    assert q.n_tokens == 1 and q.term_id.size == 0
    d.close()


def test_batch_resolution_is_fast():
    """256 queries x 3 tokens against a 200K-term vocabulary (the h1 shape): prefix mode must be far below the
    1.3 ms a batch spends on the GPU, tolerance 1 within a few batches' worth."""
    rng = np.random.default_rng(1)
    n = 200_000
    words = sorted({"".join(map(chr, rng.integers(97, 123, size=int(rng.integers(4, 11))))) for _ in range(int(n * 1.05))})[:n]
    d = ob.TermDictionary(1)
    d.add_terms(0, words)
    queries = [" ".join(words[int(i)] for i in rng.integers(0, n, size=3)) for _ in range(256)]
    d.resolve_batch(queries[:2])                               # builds the sorted index
    best = {}
    for name, kw in (("exact", dict(exact=True)), ("prefix", {}), ("tolerance1", dict(tolerance=1))):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            b = d.resolve_batch(queries, **kw)
            ts.append(time.perf_counter() - t0)
        best[name] = min(ts)
        assert b.n_queries == 256 and b.term_id.size >= 768
    print("resolve 256x3 over 200K terms:", {k: f"{v * 1e6:.0f} us" for k, v in best.items()})
    assert best["prefix"] < 5e-3 and best["exact"] < 5e-3      # includes the ctypes marshalling of 256 strings
    assert best["tolerance1"] < 2.0
    d.close()


SNOWBALL_SAMPLE = """consign consign consigned consign consigning consign consignment consign consist consist consisted consist
consistency consist consistent consist consistently consist consisting consist consists consist consolation consol
consolations consol consolatory consolatori console consol consoled consol consoles consol consolidate consolid
consolidated consolid consolidating consolid consoling consol consolingly consol consols consol consonant conson
consort consort consorted consort consorting consort conspicuous conspicu conspicuously conspicu conspiracy conspiraci
conspirator conspir conspirators conspir conspire conspir conspired conspir conspiring conspir constable constabl
constables constabl constance constanc constancy constanc constant constant knack knack knackeries knackeri knacks knack
knag knag knave knave knaves knave knavish knavish kneaded knead kneading knead knee knee kneel kneel kneeled kneel
kneeling kneel kneels kneel knees knee knell knell knelt knelt knew knew knick knick knif knif knife knife knight knight
knightly knight knights knight knit knit knits knit knitted knit knitting knit knives knive knob knob knobs knob
knock knock knocked knock knocker knocker knockers knocker knocking knock knocks knock knopp knopp knot knot knots knot
skies sky dying die cries cri ties tie gas gas this this gaps gap kiwis kiwi hopping hop hoping hope running run happy happi
generously generous communication communic relational relat conditional condit sensibility sensibl""".split()


def test_english_stemmer_reproduces_the_snowball_sample_vocabulary():
    """oc_stem_english (csrc/stem_en.h) on the sample pairs published with the Snowball English (Porter2) algorithm."""
    for w, e in zip(SNOWBALL_SAMPLE[0::2], SNOWBALL_SAMPLE[1::2]):
        assert ob.TermDictionary.stem_english(w) == e, (w, ob.TermDictionary.stem_english(w), e)
    d = ob.TermDictionary(1)
    d.add_terms(0, ["consol", "consolations", "knight"])
    d.use_english_stemmer()
    q = d.resolve_batch(["Consolations knightly"]).query(0)          # originals + stems, flattened (token_score.rs:196-204)
    assert q.n_tokens == 4
    assert [q.term_id[a:b].tolist() for a, b in zip(q.token_term_offsets[:-1], q.token_term_offsets[1:])] == [[1], [0, 1], [], [2]]
    q = d.resolve_batch(["Consolations knightly"], exact=True).query(0)
    assert q.n_tokens == 2 and q.term_id.tolist() == [1]
    d.close()
