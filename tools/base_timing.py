import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_algos_b200 import _lib
h = _lib.default_handle()
st = (C.c_longlong * 17)()
for rep in range(2):
    h.check(h.lib.gpk_debug_base_timing(h.h, C.addressof(st)))
    v = list(st)
    print("total cycles", v[16] - v[0])
    labels = {1: "load", 2: "potf2 cols 0-31", 3: "cols 32-63", 4: "cols 64-95", 5: "cols 96-127", 6: "store L", 7: "inv 8",
              8: "inv 16", 9: "inv 32", 10: "inv 64", 11: "inv 128", 16: "store Li"}
    prev = v[0]
    for i in sorted(labels):
        if i == 16:
            print(f"  {labels[i]:18s} {v[16] - v[15]:8d}")
        else:
            print(f"  {labels[i]:18s} {v[i] - prev:8d}")
            prev = v[i]
