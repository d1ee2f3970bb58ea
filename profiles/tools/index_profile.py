#!/usr/bin/env python3
"""one full scan of the frozen test_collection index per codec type (run under ncu for the launch list)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_data as g
from dint_b200 import driver
G = g.Golden()
for type_ in (sys.argv[1:] or ["single_packed_dint", "multi_packed_dint"]):
    idx = driver.IndexFile(type_, G.bytes(G.index_name(type_)))
    for _ in range(3):
        idx.scan()
    torch.cuda.synchronize()
