#!/usr/bin/env python
"""The 64-bit read-name hash stands in for the read name (fragment tracker, mate maps): counts collisions among N distinct names of
the shapes sequencers and this repo's generator write.   python tools/name_hash_collisions.py [--names 200000000]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnaseqc_b200 import api

ap = argparse.ArgumentParser(); ap.add_argument("--names", type=int, default=200_000_000); a = ap.parse_args()
H = api.host_lib()
H.rsqh_hash_qname_series.argtypes = [C.c_char_p, C.c_uint64, C.c_size_t, C.c_void_p]
out = {}
for prefix in (b"r", b"A00123:45:HXXXXXXXX:1:1101:", b"SRR1234567."):
    t = time.time()
    h = np.empty(a.names, np.uint64)
    chunk = 10_000_000
    for lo in range(0, a.names, chunk):
        n = min(chunk, a.names - lo)
        H.rsqh_hash_qname_series(prefix, lo, n, h[lo:].ctypes.data)
    h.sort()
    dup = int((h[1:] == h[:-1]).sum())
    out[prefix.decode() + "<id>"] = {"names": a.names, "colliding_pairs": dup, "expected": a.names ** 2 / 2 ** 65, "seconds": round(time.time() - t, 1)}
    print(prefix, out[prefix.decode() + "<id>"], flush=True)
print(json.dumps(out))
