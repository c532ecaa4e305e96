"""Dev tool: phase timestamps of ground_fwd_kernel (needs libnafae_b200_trace.so, -DNAFAE_TRACE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NAFAE_B200_LIB"] = os.path.join(ROOT, "tools", "_build", "libnafae_b200_trace.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
from nafae_b200 import synth, _C
from nafae_b200.pipeline import GroundingStep
c = synth.CONFIGS["cfg2"]
st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"], device="cuda:0")
b = synth.make_batch("cfg2", 1234)
print("lens", b["lens"])
st.load(b)
for _ in range(int(os.environ.get('ITERS', '400'))):
    st.run()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 256)()
_C.lib.nafae_debug_read_trace.argtypes = [ctypes.c_void_p]
_C.lib.nafae_debug_read_trace(buf)
t = np.array(buf[:], dtype=np.int64)
print("block0 (cycles): P1 done %d, after ticket %d" % (t[1] - t[0], t[2] - t[0]))
names = ["start", "loaded", "stats+gram", "-", "Sf+inv", "pairs", "end", "ticket"]
for a in range(c["Na"]):
    seg = t[16 + a * 8: 16 + a * 8 + 8]
    print("seg %d (cycles since P2 start):" % a, " ".join("%s %5d" % (n, x - seg[0]) for n, x in zip(names, seg)))
print("P3 cycles: loaded %d computed %d end %d" % (t[121] - t[120], t[122] - t[120], t[125] - t[120]))
print("seg3 gram iters (cycles since P2 start):", " ".join(str(x - t[16 + 3 * 8]) for x in t[202:212] if x > 0))
