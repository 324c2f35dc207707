"""One LU factorisation of a random K×K matrix (for ncu launch lists).  usage: python profiles/run_lu_one.py [K=2561] [reps=1]"""
import ctypes as C, json, sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
sk = ge.load_package()
L = sk._lib
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2561
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
A = torch.randn((K, K), dtype=torch.float64, device="cuda")
ms, tf = C.c_double(0), C.c_double(0)
L.check(L.lib.sk_lu_benchmark(C.c_void_p(A.data_ptr()), K, reps, C.byref(ms), C.byref(tf), None))
print(json.dumps({"K": K, "ms": ms.value, "tflops": tf.value}))
