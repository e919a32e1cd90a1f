#!/usr/bin/env python3
"""A CPU MODEL (no GPU) of how many cell loads k_long_jump needs per match byte, by how far k_long_emit has resolved the cells
(not at all / inside a slice of N sequences / inside the block) and by how many bytes are in flight at once (the window).
Within a window all bytes take one step per round, reading the cells of the round before (synchronous pointer doubling);
everything below the window is final.  Data: a 24 MiB frame of the configs[2] generator, sequences from the oracle trace.
Results of round 1 are quoted in DESIGN.md section 7.  usage: python scripts/hop_model.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import corpus as cg
from oracle import pyszo
SIZE = 24<<20
c = cg.config3_single_frame(SIZE, 23)
want, tr = pyszo.decode_frame(c.frame(0), True)
N = len(want)
def build_cells(slice_seqs):
    """cells as k_long_emit leaves them: resolved inside the region (block, or slice of slice_seqs sequences)"""
    d = np.zeros(N, dtype=np.int64)
    for b in tr.blocks:
        if b.type != 2: continue
        pos = b.out_off
        region = pos
        for si, ((ll, ml, ofv), off) in enumerate(zip(b.sequences, b.real_offsets)):
            if slice_seqs and si % slice_seqs == 0: region = pos
            pos += ll
            m = np.arange(ml)
            imm = np.where(m < off, off, off*(m//off+1))
            if slice_seqs is None:
                d[pos:pos+ml] = imm
            else:
                src = pos + m - imm
                inside = src >= region
                add = np.where(inside, d[np.maximum(src,0)], 0)
                d[pos:pos+ml] = imm + add
            pos += ml
    return d
def simulate(d, W):
    d = d.copy(); loads = 0; rounds_tot = 0
    for a in range(0, N, W):
        idx = a + np.nonzero(d[a:a+W])[0]
        r = 0
        while len(idx):
            e = d[idx - d[idx]]
            loads += len(idx)
            act = e > 0
            d[idx[act]] += e[act]
            idx = idx[act]; r += 1
        rounds_tot += r
    return loads, rounds_tot
nmatch = None
for name, sl in (("none", None), ("block", 0), ("slice4096", 4096), ("slice1024", 1024), ("slice256", 256)):
    t=time.time(); d = build_cells(sl); nm = int((d>0).sum())
    for W in (1<<20, 3<<20):
        loads, rounds = simulate(d, W)
        print(f"emit resolution {name:10s} window {W>>20} MiB: loads per match byte {loads/nm:.2f}, rounds per window {rounds/((N+W-1)//W):.1f}  ({time.time()-t:.0f}s)")
