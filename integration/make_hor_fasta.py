#!/usr/bin/env python
"""Synthetic alpha-satellite-like HOR arrays: an ancestral array of
HORs (12 monomers of 171 bp, ~22 % monomer-to-monomer divergence, ~2 % copy-to-copy divergence),
descendants mutated independently (0.4 % substitutions, 0.1 % 1-bp indels, optional whole-HOR indels).
usage: make_hor_fasta.py OUT.fa N_SEQS LENGTH_BP [SEED] [HOR_INDELS]"""
import sys

import numpy as np


def main():
    out, nseq, length = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    hor_indels = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)

    def mutate(seq, sub, indel=0.0):
        res = []
        for c in seq:
            u = rng.random()
            if u < indel / 2:
                continue
            if u < indel:
                res.append(alpha[rng.integers(4)])
            if rng.random() < sub:
                c = alpha[(np.searchsorted(alpha, c) + rng.integers(1, 4)) % 4]
            res.append(c)
        return np.asarray(res, dtype=np.uint8)

    mono = alpha[rng.integers(0, 4, 171)]
    hor = np.concatenate([mutate(mono, 0.22) for _ in range(12)])
    ncopies = max(1, length // len(hor))
    ancestor = [mutate(hor, 0.02) for _ in range(ncopies)]
    with open(out, "w") as f:
        for s in range(nseq):
            copies = list(ancestor)
            for _ in range(hor_indels):
                k = int(rng.integers(0, len(copies)))
                if rng.random() < 0.5 and len(copies) > 2:
                    del copies[k]
                else:
                    copies.insert(k, copies[k])
            seq = mutate(np.concatenate(copies), 0.004, 0.001)
            f.write(f">seq{s}\n")
            txt = seq.tobytes().decode()
            for i in range(0, len(txt), 80):
                f.write(txt[i:i + 80] + "\n")


if __name__ == "__main__":
    main()
