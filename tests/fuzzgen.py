"""Adversarial tiny GECKO CSV inputs (SURVEY.md §4 fuzz classes): tiny sequences, duplicate and edge-of-bucket
centers, zero lengths, strands other than f/r, CSB rows, truncated rows, empty fields, rows in the dropped
last X bucket, header lengths below 100 (max_index == 0)."""
from __future__ import annotations

import random

from repkiller_b200.frags import make_header

RATIOS = [(0.05, 0.05), (0.5, 0.5), (1.0, 0.1), (0.1, 1.0), (2.5, 3.0), (0.01, 0.9)]


def fuzz_case(seed: int):
    """Returns (csv_text, len_ratio, pos_ratio)."""
    rnd = random.Random(seed)
    kind = seed % 6
    lx = rnd.choice([95, 130, 250, 999, 1000, 1001, 5000, 20000]) if kind != 5 else rnd.choice([5000, 20000])
    ly = rnd.choice([95, 180, 400, 1000, 2500, 20000]) if kind != 5 else rnd.choice([5000, 20000])
    nrows = rnd.randint(1, 400)
    maxlen = rnd.choice([0, 3, 20, 60, 150])
    rows = []
    # a few anchor points so that repeats/ties and bucket-edge centers are common
    anchors_x = [rnd.randrange(0, max(1, lx)) for _ in range(rnd.randint(1, 6))]
    anchors_y = [rnd.randrange(0, max(1, ly)) for _ in range(rnd.randint(1, 6))]
    for _ in range(nrows):
        ln = rnd.randint(0, maxlen)
        hx = max(0, lx - ln)   # keep xStart + length <= lx so that centers stay inside the occupation lists
        hy = max(0, ly - ln)
        mode = rnd.random()
        if mode < 0.45:
            x = min(hx, max(0, rnd.choice(anchors_x) + rnd.randint(-3, 3)))
            y = min(hy, max(0, rnd.choice(anchors_y) + rnd.randint(-3, 3)))
        elif mode < 0.6:
            # centers on a 100-bp bucket edge: residues 98, 99, 0, 1
            c = 100 * rnd.randint(0, max(0, lx // 100)) + rnd.choice([98, 99, 0, 1])
            x = min(hx, max(0, c - ln // 2))
            c = 100 * rnd.randint(0, max(0, ly // 100)) + rnd.choice([98, 99, 0, 1])
            y = min(hy, max(0, c - ln // 2))
        else:
            x = rnd.randint(0, hx)
            y = rnd.randint(0, hy)
        strand = rnd.choice("ffffrrrrx")
        sim = rnd.choice([0.0, 12.5, 87.92, 100.0, 66.666664, 33.0])
        ident = int(ln * sim / 100)
        score = 4 * ident
        row = f"Frag,{x},{y},{x + max(ln, 1) - 1},{y + max(ln, 1) - 1},{strand},{rnd.randint(0, 3)},{ln},{score},{ident},{sim:.9g},{sim:.9g},0,0"
        q = rnd.random()
        if q < 0.03:
            row = "CSB" + row[4:]
        elif q < 0.06:
            row = ",".join(row.split(",")[: rnd.randint(1, 13)])          # truncated row: padded by the reference
        elif q < 0.08:
            parts = row.split(",")
            parts[rnd.randint(1, 13)] = ""                                 # empty field: rejected
            row = ",".join(parts)
        elif q < 0.09:
            row = row + ","                                                # trailing comma
        elif q < 0.10:
            row = ""
        rows.append(row)
    text = make_header(lx, ly, nrows) + "\n".join(rows)
    if rnd.random() < 0.7:
        text += "\n"
    lr, pr = RATIOS[seed % len(RATIOS)]
    return text, lr, pr
