"""Host side of wva_ingest_write (csrc/ingest_scatter.hpp through tests/host_emul): 1.44 M samples into the columns, in
registry order and in random order, serial against the two-pass radix partition over T threads."""
import ctypes as C, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "host_emul")], check=True)
lib = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libemul.so"))
f = lib.emul_ingest_scatter
f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
S = 1_437_822
g = np.random.default_rng(1)
out = {"samples": S}
for name in ("registry_order", "shuffled"):
    slot = np.arange(S, dtype=np.int32) if name == "registry_order" else g.permutation(S).astype(np.int32)
    val = g.random(S); col = np.zeros(S); has = np.zeros(S, np.uint8)
    res = {}
    for T in (1, 2, 4, 8, 16):
        ts = []
        for _ in range(9):
            t0 = time.perf_counter(); f(col.ctypes.data, has.ctypes.data, S, 1, S, slot.ctypes.data, val.ctypes.data, T)
            ts.append(time.perf_counter() - t0)
        res[f"threads_{T}_ms"] = round(1e3 * min(ts), 3)
    out[name] = res
try: out["cgroup_cpu_max"] = open("/sys/fs/cgroup/cpu.max").read().strip()
except OSError: out["cgroup_cpu_max"] = None
print(json.dumps(out))
