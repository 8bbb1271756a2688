"""Known answers for curand_init(seed, subsequence, 0) from cuRAND ITSELF: the toolkit's curand_kernel.h host-compiled
into oracle/_ref/libref_host_*.so (oracle/ref_host_driver.cu: refh_curand_state).  Run where /root/reference and nvcc
exist (`make -C oracle ref_host` first); writes tests/golden/curand_subsequence_kat.json.

    python tests/golden/gen_curand_kat.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

O = entry.load_oracle()
rh = O.RefHost("oct_spl30")
cases = [(1984, s) for s in (0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 255, 256, 1199, 1200, 65535, 65536, 959999, 8294399, 33177599,
                              33177600, 2 ** 31 - 1, 2 ** 32 + 5, 2 ** 39 + 12345)]
cases += [(0, 1), (1, 1), (7, 123456789), (2 ** 32 + 1984, 42), (2 ** 40 + 17, 77), (2 ** 63 + 3, 1000003)]
out = {"source": "cuRAND 10.3.10 (CUDA 12.9 curand_kernel.h) host-compiled; state words d, v0..v4 after curand_init(seed, subsequence, 0)",
       "cases": [{"seed": s, "subsequence": q, "state": [int(x) for x in rh.curand_state(s, q)]} for s, q in cases]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "curand_subsequence_kat.json"), "w"), indent=1)
print(len(cases), "cases written")
