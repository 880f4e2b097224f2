"""CPU-side checks: the CUDA library loads and exports every symbol the header declares
(no compute without a GPU), product code never touches the oracle, host logic works."""
import ctypes
import os
import re

import numpy as np
import pytest

import oramacore_b200 as ob
from oramacore_b200 import _lib, synth
from oramacore_b200.hostindex import HostStringIndex, tokenize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    ob.build()
    hdr = open(os.path.join(ROOT, "include", "oramacore_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(oc_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(ob.SO_PATH)
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/oramacore_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ob.OcError) as e:
        ob.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "oramacore_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "oracle/" not in src, f


def test_struct_layouts_match_header():
    # sizes the Rust/C side would see (repr(C)); guards the ctypes mirror against drift
    sizes = (ctypes.c_size_t * 4)()
    ob.lib().oc_abi_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(_lib.SearchParams), ctypes.sizeof(_lib.Timing),
                           ctypes.sizeof(_lib.EmbInfo), ctypes.sizeof(_lib.StrInfo)]


def test_tokenizer_and_index_builder():
    assert tokenize("The Pen, is ON the-table!") == ["the", "pen", "is", "on", "the", "table"]
    h = HostStringIndex(["text"])
    h.insert(10, {"text": "alpha beta beta"})
    h.insert(20, {"text": "beta gamma"})
    d = h.commit()
    f = d.fields[0]
    assert d.n_rows == 2 and d.row_doc_ids.tolist() == [10, 20]
    assert h.terms[0] == ["alpha", "beta", "gamma"]
    assert f.term_offsets.tolist() == [0, 1, 3, 4]
    assert f.post_row.tolist() == [0, 0, 1, 1] and f.post_tf.tolist() == [1, 2, 1, 1]
    assert f.post_len.tolist() == [3, 3, 2, 2] and abs(f.avg_field_len - 2.5) < 1e-6
    q = h.resolve("bet")
    assert q.n_tokens == 1 and q.term_id.tolist() == [1] and q.term_weight.tolist() == [1.0]
    q = h.resolve("beta")
    assert q.term_weight.tolist() == [2.0]
    assert h.resolve("").term_id.tolist() == [0, 1, 2]


def test_synth_shapes():
    rows = synth.make_vectors(1000, 64, seed=1)
    assert rows.shape == (1000, 64) and rows.dtype == np.float32
    q, j = synth.make_vector_queries(rows, 4, seed=2)
    assert q.shape == (4, 64)
    d = synth.make_text_corpus(2000, 300, seed=3)
    f = d.fields[0]
    f.validate()
    assert int(f.term_offsets[-1]) == f.post_row.shape[0] > 2000
    for t in (0, 5, 100):
        seg = f.post_row[int(f.term_offsets[t]):int(f.term_offsets[t + 1])]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)
    assert np.array_equal(synth.make_text_corpus(2000, 300, seed=3).fields[0].post_row, f.post_row)
    tq = synth.make_text_queries(300, 5, seed=4)
    assert len(tq) == 5 and all(t.n_tokens == 3 and len(set(t.term_id.tolist())) == 3 for t in tq)
