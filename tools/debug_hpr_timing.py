import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_synth_once.py")).read())
from cloudaae_b200 import _capi
lib = ctypes.CDLL(_capi.LIB_PATH)
buf = np.zeros(8 * 256, np.int64)
lib.caae_debug_hpr_timing(buf.ctypes.data_as(ctypes.c_void_p))
t = buf.reshape(256, 8)[:128]
print("cycles mean  setup %.0f  phase1 %.0f  phase2 %.0f  phase3 %.0f | nsurv %.0f dirty %.1f" % tuple(t[:, k].mean() for k in range(6)))
print("cycles max   setup %.0f  phase1 %.0f  phase2 %.0f  phase3 %.0f | nsurv %.0f dirty %.0f" % tuple(t[:, k].max() for k in range(6)))

order = np.argsort(-t[:, 2])[:8]
cls = z["class_id"][sel]
for k in order:
    print("cloud", k, "class", cls[k], "p1 %d p2 %d p3 %d nsurv %d dirty %d slow %d" % tuple(t[k, 1:7]))
print("slow-path evaluations per (survivor x point): mean %.4f" % (t[:, 6].sum() / (t[:, 4] * 2048).sum()))
