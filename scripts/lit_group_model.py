import sys, collections
sys.path.insert(0,'.')
from oracle import pyszo
from tools import corpus as cg
def entries(c, nmax):
    ent = []  # (slot, max_bits, streams, regen)
    slot = 0
    for i in range(min(c.nframes, nmax)):
        _, tr = pyszo.decode_frame(c.frame(i), want_trace=True)
        for b in tr.blocks:
            if b.type != 2 or b.lit_type < 2: continue
            if b.lit_type == 2: slot += 1
            ent.append((slot, b.huf_max_bits, b.lit_streams, b.lit_regen))
    return ent
def model(ent, order=None):
    if order: ent = sorted(ent, key=order)
    steps = 0; ideal = 0
    for g in range(0, len(ent), 8):
        grp = ent[g:g+8]
        i = 0
        while i < len(grp):
            used = 0; prev = None; mx = 0; j = i
            while j < len(grp):
                s, mb, st, rg = grp[j]
                if s != prev:
                    if used + (1 << mb) > 2048: break
                    used += 1 << mb; prev = s
                mx = max(mx, rg if st == 1 else (rg + 3) // 4)
                j += 1
            steps += mx; i = j
        for s, mb, st, rg in grp: ideal += rg
    return steps, ideal / 32
for name, c, n in (("mixed", cg.config5_mixed(256 << 20), 100000), ("literal", cg.config4_literal_heavy(64), 64), ("text", cg.config2_text_frames(256), 256)):
    ent = entries(c, n)
    s, i = model(ent)
    s2, _ = model(ent, order=lambda e: -(e[3] if e[2] == 1 else (e[3] + 3) // 4))
    s3, _ = model(ent, order=lambda e: (-e[1], -(e[3] if e[2] == 1 else (e[3] + 3) // 4)))
    print(name, "entries", len(ent), "warp steps: as is", s, "sorted by stream length", s2, "by (maxbits, length)", s3, "ideal", int(i), "utilisation %.2f / %.2f / %.2f" % (i / s, i / s2, i / s3))
print("---- more orders")
def famlen(ent):
    fl = collections.defaultdict(int)
    for s, mb, st, rg in ent:
        fl[s] = max(fl[s], rg if st == 1 else (rg + 3) // 4)
    return fl
for name, c, n in (("mixed", cg.config5_mixed(256 << 20), 100000), ("literal", cg.config4_literal_heavy(64), 64)):
    ent = entries(c, n)
    fl = famlen(ent)
    ln = lambda e: (e[3] if e[2] == 1 else (e[3] + 3) // 4)
    for label, key in (("(-maxbits, slot)", lambda e: (-e[1], e[0])),
                       ("(-maxbits, -famlen, slot, -len)", lambda e: (-e[1], -fl[e[0]], e[0], -ln(e))),
                       ("(-famlen, slot, -len)", lambda e: (-fl[e[0]], e[0], -ln(e))),
                       ("(-lenclass, -maxbits, slot)", lambda e: (-(fl[e[0]]).bit_length(), -e[1], e[0], -ln(e)))):
        s, i = model(ent, order=key)
        print(name, label, "utilisation %.2f" % (i / s))
print("---- lane packing: a warp takes entries while the streams fit 32 lanes and the distinct tables fit `cells`")
def model2(ent, order, cells=2048, maxent=32):
    ent = sorted(ent, key=order) if order else ent
    steps = 0; ideal = sum(e[3] for e in ent) / 32
    i = 0
    while i < len(ent):
        used = 0; prev = None; mx = 0; lanes = 0; j = i
        while j < len(ent) and j - i < maxent:
            s, mb, st, rg = ent[j]
            need = 0 if s == prev else (1 << mb)
            if used + need > cells or lanes + st > 32: break
            used += need; prev = s; lanes += st
            mx = max(mx, rg if st == 1 else (rg + 3) // 4)
            j += 1
        steps += mx; i = j
    return ideal / steps
for name, c, n in (("mixed", cg.config5_mixed(256 << 20), 100000), ("literal", cg.config4_literal_heavy(64), 64)):
    ent = entries(c, n)
    fl = famlen(ent)
    ln = lambda e: (e[3] if e[2] == 1 else (e[3] + 3) // 4)
    for label, key in (("as is", None), ("(-famlen, slot, -len)", lambda e: (-fl[e[0]], e[0], -ln(e))), ("(-len)", lambda e: -ln(e)), ("(-maxbits,-len)", lambda e: (-e[1], -ln(e)))):
        print(name, label, "4 KB: %.2f" % model2(ent, key), " 8 KB: %.2f" % model2(ent, key, 4096), " 16 KB: %.2f" % model2(ent, key, 8192))
