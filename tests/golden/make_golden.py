"""Regenerates tests/golden/*.  Needs oracle/_ref/repkiller_ref, i.e. /root/reference (build container only).

The reference ships no golden vectors, so the fixtures are OUTPUTS OF THE REFERENCE ITSELF
(unmodified sources behind oracle/ref_driver.cpp) on
  * fuzz.json.gz   tiny adversarial CSV inputs (tests/fuzzgen.py) with the reference's full output bytes;
  * medium.json    generated workloads (repkiller_b200.gen, deterministic) with the md5 / group count of the
                   reference's output file, so large cases stay a few bytes in git.
Cases on which the reference has undefined behaviour (out-of-bounds bucket index) are skipped.
usage: python tests/golden/make_golden.py
"""
import gzip, hashlib, json, os, sys, tempfile
from dataclasses import asdict, replace

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O          # noqa: E402
from repkiller_b200 import gen          # noqa: E402
from fuzzgen import fuzz_case           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

MEDIUM = {
    "c1": gen.WORKLOADS["c1"],
    "c1_loose": replace(gen.WORKLOADS["c1"], name="c1_loose", len_ratio=0.5, pos_ratio=0.5),
    "c3_small": replace(gen.WORKLOADS["c1"], name="c3_small", n=300_000, p_rep=0.9, families=20, tandem_every=2, seed=7),
    "c2_small": gen.scaled(gen.WORKLOADS["c2"], 1_000_000),
    "dense": replace(gen.WORKLOADS["c1"], name="dense", n=200_000, lx=400_000, ly=300_000, families=50, seed=11),
}


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    tmp = tempfile.mkdtemp()
    cases, skipped = [], 0
    seed = 0
    while len(cases) < 120:
        text, lr, pr = fuzz_case(seed)
        seed += 1
        inp = os.path.join(tmp, "in.csv")
        with open(inp, "w", newline="") as f:
            f.write(text)
        try:
            rec, lx1, ly1, hdr = O.load_csv(inp)
            g = O.group(rec, lx1, ly1, lr, pr)
        except (ValueError, RuntimeError):
            skipped += 1          # reference UB / throws: not a parity case
            continue
        outp = os.path.join(tmp, "out.csv")
        O.run_ref(inp, outp, lr, pr)
        ref_out = open(outp, "rb").read()
        O.write_output(outp + ".o", hdr, rec, g)
        assert open(outp + ".o", "rb").read() == ref_out, f"oracle != reference on fuzz seed {seed - 1}"
        cases.append({"seed": seed - 1, "len_ratio": lr, "pos_ratio": pr, "csv": text,
                      "ref_out": ref_out.decode("latin1"), "n_kept": int(g.n_kept), "n_groups": int(g.n_groups)})
    with gzip.open(os.path.join(HERE, "fuzz.json.gz"), "wt", compresslevel=9) as f:
        json.dump(cases, f)
    print(f"fuzz: {len(cases)} cases kept, {skipped} skipped (reference UB)")

    med = {}
    for name, w in MEDIUM.items():
        rec = gen.generate(w)
        inp = os.path.join(tmp, name + ".csv")
        O.write_input_csv(inp, rec, w.lx, w.ly)
        outp = os.path.join(tmp, name + ".out")
        info = O.run_ref(inp, outp, w.len_ratio, w.pos_ratio)
        data = open(outp, "rb").read()
        med[name] = {"workload": asdict(w), "ref_md5": hashlib.md5(data).hexdigest(), "ref_bytes": len(data),
                     "n_groups": info["n_groups"], "n_frags": info["n_frags"],
                     "records_md5": hashlib.md5(rec.tobytes()).hexdigest()}
        print(name, med[name]["ref_md5"], info)
    with open(os.path.join(HERE, "medium.json"), "w") as f:
        json.dump(med, f, indent=1)


if __name__ == "__main__":
    main()
