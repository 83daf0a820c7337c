"""End-to-end timing of the drop-in CLI on a generated GECKO CSV (tooling): python tools/cli_e2e.py [n]"""
import os, subprocess, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from repkiller_b200 import gen
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = gen.scaled(gen.WORKLOADS["c2"], n)
rec = gen.generate(w)
inp, out = "/tmp/cli_in.csv", "/tmp/cli_out.csv"
O.write_input_csv(inp, rec, w.lx, w.ly)
print("csv bytes", os.path.getsize(inp))
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "repkiller_b200", "bin", "repkiller")
for i in range(2):
    t0 = time.perf_counter()
    p = subprocess.run([cli, inp, out, "0.05", "0.05"], capture_output=True, env=dict(os.environ, RK_TIMING="1"))
    dt = time.perf_counter() - t0
    print(f"run {i}: rc={p.returncode} wall {dt:.2f} s  out bytes {os.path.getsize(out)}")
    print(p.stderr.decode()[-600:])
if O.have_ref() and n <= 2_000_000:
    t0 = time.perf_counter(); info = O.run_ref(inp, "/tmp/ref_out.csv", 0.05, 0.05); print("reference", info, f"wall {time.perf_counter()-t0:.2f}")
    print("bytes identical:", open(out, "rb").read() == open("/tmp/ref_out.csv", "rb").read())
