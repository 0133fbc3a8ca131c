"""Compare the sizer kernels on a large system (throughput regime)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
A = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = int(sys.argv[3]) if len(sys.argv) > 3 else 128
modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1, 2]
gang = int(sys.argv[5]) if len(sys.argv) > 5 else 1
d = pkg.synth.queue_system(S, A, N, n_classes=3, stream=3, R=N)
with pkg.Engine(0) as e:
    e.load_system(d)
    ref = None
    for mode in modes:
        natural = mode < 0                                                # negative mode: natural order, lanes refill one by one
        e.set_option(2, 0 if natural else (1 if gang >= 0 else -1))      # gang -1: library default for both options
        e.set_option(3, 0 if natural else gang)
        mode = abs(mode)
        e.set_option(1, mode)
        for rep in range(2):
            e.calculate()
        t = e.timing()
        c = e.candidates()
        if ref is None:
            ref = c
        same = all((c[k] == ref[k]).all() for k in ("state", "num_replicas", "cost", "itl", "ttft"))
        print(f"mode {mode}: calculate {t['calculate_ms']:.2f} ms, solves {t['chain_solves']}, states {t['chain_states']}, "
              f"{S*A/t['calculate_ms']*1e3:.3e} pairs/s, same_as_first={same}")
