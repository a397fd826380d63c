import cProfile, pstats, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_c1 import make_docs
import torch
from datasketch_b200 import MinHash
docs = make_docs(100_000, 256)
MinHash.bulk(docs[:64], num_perm=128)
for rep in range(2):
    t = time.perf_counter(); MinHash.bulk(docs, num_perm=128); print("bulk", time.perf_counter() - t)
pr = cProfile.Profile(); pr.enable()
MinHash.bulk(docs, num_perm=128)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
