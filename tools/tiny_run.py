"""Smallest possible run of the decode path (debug aid): B rows, a few steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from flow_check import make, gen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g, embed = make({}, max_batch=max(B, 1), max_context=128)
o = gen(g, embed, [16, 7, 9, 5][:B], steps)
print("ids", o.ids[0][:3].tolist(), flush=True)
