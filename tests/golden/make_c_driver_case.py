"""Generates tests/golden/c_driver_case.bin: inputs of one small hybrid search (synthetic, seeded) and the
ORACLE's answer, in the flat little-endian layout tests/c_driver/driver.c reads.  Run from the repo root:
    python tests/golden/make_c_driver_case.py
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc
from oramacore_b200 import synth
from oramacore_b200.engine import TextQueryBatch

n, dim, vocab, B, limit = 1500, 64, 150, 4, 10
rows = synth.make_vectors(n, dim, seed=101)
qv, _ = synth.make_vector_queries(rows, B, seed=102)
data = synth.make_text_corpus(n, vocab, seed=103, mean_len=12.0)
texts = synth.make_text_queries(vocab, B, seed=104)
f = data.fields[0]
tb = TextQueryBatch(texts)
sb = orc.SearchBatch(orc.StrIndex(data), orc.EmbStore(rows))
for i in range(B):
    sb.add(2, limit=limit, similarity=0.0, q_vec=qv[i], text=texts[i])
od, os_, on, oc = sb.run(1)
out = os.path.join(ROOT, "tests", "golden", "c_driver_case.bin")
with open(out, "wb") as fh:
    fh.write(struct.pack("<IIIIIIQII", 0x0C0DE001, n, dim, vocab, B, limit, f.post_row.shape[0],
                         tb.token_term_offsets.shape[0] - 1, tb.term_id.shape[0]))
    fh.write(struct.pack("<f", np.float32(f.avg_field_len)))
    for a, dt in ((rows, np.float32), (f.term_offsets, np.uint64), (f.post_row, np.uint32), (f.post_tf, np.uint16),
                  (f.post_len, np.uint16), (qv, np.float32), (tb.q_token_offsets, np.uint32), (tb.token_term_offsets, np.uint32),
                  (tb.term_field, np.uint32), (tb.term_id, np.uint32), (tb.term_weight, np.float32),
                  (od, np.uint64), (os_, np.float32), (on, np.uint32), (oc, np.uint64)):
        fh.write(np.ascontiguousarray(a, dt).tobytes())
print(out, os.path.getsize(out), "bytes")
