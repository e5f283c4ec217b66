#!/usr/bin/env python
"""Critical path of the persistent decode kernel from a raw phase trace (ZB_MEGA_TRACE_DUMP=file.npz tools/mega_trace.py).
SM clocks are not synchronised across SMs: every CTA's clock is aligned on the first stamp of the launch (all CTAs start
within the launch ramp) -- good to a few hundred ns, enough to see which phase of which CTA the next op waits for.

For every op: when did the last CTA finish it (its stamp 4), which CTA that was, and how that CTA spent the op
(wait for input + prologue, loop, epilogue).  usage: mega_critical.py trace.npz"""
import sys

import numpy as np


def main():
    z = np.load(sys.argv[1])
    t, kinds, ghz = z["t"].astype(np.float64), z["kinds"], float(z["ghz"])
    t[z["t"] == 0] = np.nan
    n, G, _ = t.shape
    t = (t - t[0, :, 0][None, :, None]) / (ghz * 1e3)   # us since the CTA's own first stamp
    names = {0: "embed", 2: "attn", 3: "final"}
    rows = []
    prev_end = 0.0
    agg = {}
    for i in range(n):
        k = int(kinds[i])
        name = names.get(k, f"gemv{(k - 100) % 1000}K{256 * (k // 1000)}")
        end = t[i, :, 4]
        if not np.isfinite(end).any():
            continue
        c = int(np.nanargmax(end))
        e = float(end[c])
        st = t[i, c]
        d = dict(op=i, name=name, cta=c, dur=e - prev_end, start_lag=st[0] - prev_end,
                 pro=st[1] - st[0], loop=st[3] - st[1], epi=st[4] - st[3], first_end=float(np.nanmin(end)), spread=e - float(np.nanmin(end)),
                 med_pro=float(np.nanmedian(t[i, :, 1] - t[i, :, 0])), med_loop=float(np.nanmedian(t[i, :, 3] - t[i, :, 1])))
        rows.append(d)
        a = agg.setdefault(name, [])
        a.append(d)
        prev_end = e
    print(f"{'class':>16} {'n':>4} {'dur':>7} {'lag':>6} {'pro':>6} {'loop':>6} {'epi':>6} {'spread':>7}   (us per op, mean over ops; of the CTA that finishes last)")
    tot = 0.0
    for name, a in agg.items():
        f = lambda key: np.nanmean([x[key] for x in a])
        tot += np.nansum([x["dur"] for x in a])
        print(f"{name:>16} {len(a):>4} {f('dur'):7.2f} {f('start_lag'):6.2f} {f('pro'):6.2f} {f('loop'):6.2f} {f('epi'):6.2f} {f('spread'):7.2f}   sum {np.nansum([x['dur'] for x in a]):8.1f}")
    print(f"sum {tot:.1f} us")
    if len(sys.argv) > 2:
        for d in rows[int(sys.argv[2]):int(sys.argv[2]) + 14]:
            print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.items()})


if __name__ == "__main__":
    main()
