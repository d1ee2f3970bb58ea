#!/usr/bin/env python3
"""TEST INFRASTRUCTURE. Turns the reference's query log (test/test_data/queries: one query per line,
whitespace-separated term ids, queries.hpp:15-28) into tests/golden/queries.json so that the config-3
ragged-batch case runs where /root/reference is not mounted."""
import json, os, sys
REF = os.environ.get("DINT_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
qs = [[int(t) for t in line.split()] for line in open(os.path.join(REF, "test", "test_data", "queries")) if line.strip()]
out = os.path.join(HERE, "..", "tests", "golden", "queries.json")
json.dump({"source": "test/test_data/queries", "n_queries": len(qs), "queries": qs}, open(out, "w"), separators=(",", ":"))
terms = sorted({t for q in qs for t in q})
print(f"{len(qs)} queries, {len(terms)} distinct terms, avg {sum(map(len, qs)) / len(qs):.2f} terms/query -> {out}")
