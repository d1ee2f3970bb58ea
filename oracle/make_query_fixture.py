#!/usr/bin/env python3
"""TEST INFRASTRUCTURE. Turns the reference's query log (test/test_data/queries: one query per line,
whitespace-separated term ids, queries.hpp:15-28) into tests/golden/queries.json so that the config-3
ragged-batch case runs where /root/reference is not mounted."""
import json, os, subprocess, sys, tempfile
REF = os.environ.get("DINT_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
qs = [[int(t) for t in line.split()] for line in open(os.path.join(REF, "test", "test_data", "queries")) if line.strip()]
out = os.path.join(HERE, "..", "tests", "golden", "queries.json")
# and_counts: what the reference's own and_query<false> (include/ds2i/queries.hpp:34-86) returns for every query on the
# frozen indexes of the test collection (oracle/_ref/ref_index and ..., built by `make -C oracle ref`); the three
# index types must agree
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import golden_data
G = golden_data.Golden()
counts = None
with tempfile.NamedTemporaryFile("w", suffix=".queries", delete=False) as f:
    f.write("".join(" ".join(map(str, q)) + "\n" for q in qs))
for type_ in golden_data.TYPES:
    r = subprocess.run([os.path.join(HERE, "_ref", "ref_index"), "and", type_, G.path(G.index_name(type_)), f.name],
                       check=True, capture_output=True, text=True)
    c = [int(x) for x in r.stdout.splitlines()[:len(qs)]]
    assert len(c) == len(qs) and (counts is None or c == counts), type_
    counts = c
os.unlink(f.name)
json.dump({"source": "test/test_data/queries", "n_queries": len(qs), "queries": qs,
           "and_counts": counts, "and_counts_source": "reference and_query<false> via oracle/_ref/ref_index and"},
          open(out, "w"), separators=(",", ":"))
print(f"and_query: {sum(counts)} results over {sum(1 for c in counts if c)} non-empty queries")
terms = sorted({t for q in qs for t in q})
print(f"{len(qs)} queries, {len(terms)} distinct terms, avg {sum(map(len, qs)) / len(qs):.2f} terms/query -> {out}")
