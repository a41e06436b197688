#!/usr/bin/env python
"""Processor-sharing model of one SM sub-partition running the attention softmax: N warps (one per chain) loop
[Z cycles that do not use the MUFU: wait for S, tcgen05.ld, tcgen05.st + arrive] -> [64 exponentials], and the time per
exponential of a warp depends on how many warps are in their exponentials at that moment (scratch/softmax_probe.cu on a B200:
11.8 / 17.45 / 25.6 / 33.5 cycles per element per warp for 1 / 2 / 3 / 4 warps). Calibration: the four-chain kernel measures
~2590 cycles per key step at S = 6272 (tools/attn_trace.py); the model gives 2522.

    python tools/xu_sharing_sim.py
"""
import random

PER_ELT = {1: 11.8, 2: 17.45, 3: 25.6, 4: 33.5, 5: 42.0, 6: 50.0}


def period(n_warps, z, steps=4000, seed=1):
    random.seed(seed)
    st = [("z", random.uniform(0, z)) for _ in range(n_warps)]
    t, done = 0.0, 0
    while done < steps:
        k = sum(1 for s in st if s[0] == "e")
        dt = min(s[1] if s[0] == "z" else s[1] * PER_ELT[k] for s in st)
        t += dt
        nxt = []
        for s in st:
            if s[0] == "z":
                r = s[1] - dt
                nxt.append(("e", 64.0) if r <= 1e-9 else ("z", r))
            else:
                r = s[1] - dt / PER_ELT[k]
                if r <= 1e-9:
                    nxt.append(("z", z))
                    done += 1
                else:
                    nxt.append(("e", r))
        st = nxt
    return t / done * n_warps


if __name__ == "__main__":
    cases = ((4, 1215, "4 chains, P aliases S (production): Z = wait for S 550 + ld 245 + st/arrive 270 + 150"),
             (4, 815, "4 chains if S(j+1) cost the chain nothing (needs 640 TMEM columns)"),
             (3, 815, "3 chains with P in its own TMEM columns (S(j+1) under the exponentials of step j)"),
             (3, 1215, "3 chains, production protocol"),
             (2, 815, "2 chains (128 keys per step would halve Z per key)"))
    for n, z, label in cases:
        p = period(n, z)
        print(f"{label:88s}: {p:6.0f} cycles per key step, {p / n:5.0f} per 128 x 64 tile, MUFU busy {n * 512 / p * 100:3.0f} % of the 8-cycle floor")
