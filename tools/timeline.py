#!/usr/bin/env python
"""Residency timeline of one wavefront render (development aid).

    RTB_TIMELINE=/path/file python tools/quick_bench.py --spp 128 --integrators 1 --traversal 3 --reps 2
    python tools/timeline.py /path/file

With RTB_TIMELINE set every warp of wf_raygen / wf_extend / wf_shade / wf_tail appends {start (globaltimer ns),
duration, kind, SM, lane, bounce}; this script rebuilds how many warps of each kind were resident on each SM over
time: mean residency, how often an SM holds few or no warps, and how long each kernel kind's warps live.  It is how
the kernel-boundary bubbles of the pipeline were found.
"""
import sys

import numpy as np

KINDS = ["raygen", "extend", "shade", "tail"]


def main(path, bins=8000):
    rec = np.fromfile(path, dtype=np.uint32).reshape(-1, 4)
    t0 = rec[:, 0].astype(np.uint64) | (rec[:, 1].astype(np.uint64) << np.uint64(32))
    dur = rec[:, 2].astype(np.uint64)
    kind = rec[:, 3] & 0xFF
    sm = (rec[:, 3] >> 8) & 0xFF
    lane = (rec[:, 3] >> 16) & 0xFF
    bounce = rec[:, 3] >> 24
    begin, end = t0.min(), (t0 + dur).max()
    span = float(end - begin)
    n_sm = int(sm.max()) + 1
    print(f"{len(rec)} warp records, {n_sm} SMs, span {span / 1e6:.3f} ms, lanes {sorted(set(lane.tolist()))}")
    # middle 80 % of the span (steady state)
    lo, hi = begin + np.uint64(span * 0.1), begin + np.uint64(span * 0.9)
    width = (hi - lo) / bins
    res = np.zeros((len(KINDS), n_sm, bins + 1), np.int32)
    for k in range(len(KINDS)):
        m = kind == k
        s = np.clip((t0[m].astype(np.float64) - float(lo)) / float(width), 0, bins).astype(np.int64)
        e = np.clip(((t0[m] + dur[m]).astype(np.float64) - float(lo)) / float(width), 0, bins).astype(np.int64)
        np.add.at(res[k], (sm[m], s), 1)
        np.add.at(res[k], (sm[m], e), -1)
    res = np.cumsum(res, axis=2)[:, :, :bins]  # warps of kind k resident on SM s in bin b
    total = res.sum(axis=0)
    print("mean resident warps per SM (steady state): " +
          ", ".join(f"{KINDS[k]} {res[k].mean():.1f}" for k in range(len(KINDS))) + f", all {total.mean():.1f} of 64")
    for thr in (1, 8, 16, 24, 32, 48):
        print(f"  SM-time with fewer than {thr:2d} resident warps: {(total < thr).mean() * 100:5.1f} %")
    print(f"  SM-time with no extend warp: {(res[1] == 0).mean() * 100:5.1f} %   extend warps when present: "
          f"{res[1][res[1] > 0].mean():.1f}")
    print(f"  SM-time with >= 32 extend warps: {(res[1] >= 32).mean() * 100:5.1f} %, >= 48: {(res[1] >= 48).mean() * 100:5.1f} %")
    for k in range(len(KINDS)):
        m = kind == k
        if m.any():
            d = dur[m].astype(np.float64) / 1e3
            print(f"{KINDS[k]:7s}: {m.sum():9d} warps, lifetime us mean {d.mean():8.1f} median {np.median(d):8.1f} "
                  f"p90 {np.percentile(d, 90):8.1f} max {d.max():8.1f}, warp-time share {d.sum() / (dur.sum() / 1e3) * 100:5.1f} %")
    # per-launch spread of extend: how long after the first warp of a launch ends does the last one end
    m = kind == 1
    key = (lane[m].astype(np.int64) << 8) | bounce[m]
    # launches repeat per batch: split by time gaps > 20 us between sorted starts of the same (lane, bounce)
    ends = (t0[m] + dur[m]).astype(np.float64)
    starts = t0[m].astype(np.float64)
    spreads, lives = [], []
    for kk in np.unique(key):
        idx = np.nonzero(key == kk)[0]
        order = idx[np.argsort(starts[idx])]
        cuts = np.nonzero(np.diff(starts[order]) > 50e3)[0] + 1
        for grp in np.split(order, cuts):
            if len(grp) >= 64:
                spreads.append((ends[grp].max() - np.percentile(ends[grp], 10)) / 1e3)
                lives.append((ends[grp].max() - starts[grp].min()) / 1e3)
    # per lane: how much of the time does the lane have NO warp running (its next kernel has not started yet)
    fine = 20000
    for ln in sorted(set(lane.tolist())):
        m = lane == ln
        s_ = np.clip((t0[m].astype(np.float64) - float(lo)) / (float(hi - lo) / fine), 0, fine).astype(np.int64)
        e_ = np.clip(((t0[m] + dur[m]).astype(np.float64) - float(lo)) / (float(hi - lo) / fine), 0, fine).astype(np.int64)
        d = np.zeros(fine + 1, np.int64)
        np.add.at(d, s_, 1)
        np.add.at(d, e_, -1)
        live = np.cumsum(d)[:fine]
        print(f"lane {ln}: no warp running {(live == 0).mean() * 100:5.1f} % of the time, fewer than 64 warps {(live < 64).mean() * 100:5.1f} %")
    # the launches of lane 0, in order: when the first warp started, when the last one ended, the gap before it
    m = lane == 0
    kk = (kind[m].astype(np.int64) << 8) | bounce[m]
    st_, en_ = t0[m].astype(np.float64), (t0[m] + dur[m]).astype(np.float64)
    order = np.argsort(st_)
    launches = []  # (kind, bounce, first start, last end, warps)
    cur = None
    for j in order:
        if cur is not None and kk[j] == cur[0]:
            cur[2] = max(cur[2], en_[j]); cur[3] += 1
        else:
            if cur is not None: launches.append(cur)
            cur = [kk[j], st_[j], en_[j], 1]
    launches.append(cur)
    gaps = {}
    prev_end = None
    shown = 0
    for key_, a_, b_, n_ in launches:
        kd, bo = int(key_ >> 8), int(key_ & 0xFF)
        if prev_end is not None:
            g = (a_ - prev_end) / 1e3
            gaps.setdefault(KINDS[kd], []).append(g)
            if 40 <= shown < 75:
                print(f"   lane 0: {KINDS[kd]:6s} bounce {bo:2d}: starts {g:8.1f} us after the previous launch ended, runs {(b_ - a_) / 1e3:8.1f} us, {n_} warps")
        shown += 1
        prev_end = b_
    for kname, g in gaps.items():
        g = np.array(g)
        print(f"lane 0: gap before a {kname} launch: mean {g.mean():7.1f} us, median {np.median(g):7.1f}, total {g.sum() / 1e3:.2f} ms over {len(g)} launches")
    # balance between the CTAs of one launch: mean against max of the CTAs' busy time (extend: one CTA per SM per launch)
    for k_id, k_name in ((1, "extend"), (2, "shade")):
        m = kind == k_id
        key = (lane[m].astype(np.int64) << 8) | bounce[m]
        starts, ends, sms = t0[m].astype(np.float64), (t0[m] + dur[m]).astype(np.float64), sm[m]
        rows = []
        for kk in np.unique(key):
            idx = np.nonzero(key == kk)[0]
            order = idx[np.argsort(starts[idx])]
            cuts = np.nonzero(np.diff(starts[order]) > 50e3)[0] + 1
            for grp in np.split(order, cuts):
                if len(grp) < 512:
                    continue
                launch_begin = starts[grp].min()
                per_sm_end = np.zeros(n_sm)
                np.maximum.at(per_sm_end, sms[grp], ends[grp] - launch_begin)
                busy = per_sm_end[per_sm_end > 0]
                rows.append((int(kk & 0xFF), busy.mean() / 1e3, busy.max() / 1e3, (ends[grp] - starts[grp]).sum() / 1e3))
        if rows:
            rows = np.array(rows)
            for b in sorted(set(rows[:, 0].astype(int)))[:6]:
                r = rows[rows[:, 0] == b]
                print(f"{k_name} bounce {b}: {len(r)} launches, last warp of an SM ends after mean {r[:, 1].mean():7.1f} us, "
                      f"of the launch after {r[:, 2].mean():7.1f} us (balance {r[:, 1].mean() / r[:, 2].mean():.2f}), "
                      f"warp-time {r[:, 3].mean() / 1e3:.2f} ms")
    if spreads:
        print(f"extend launches: {len(spreads)}, launch duration us mean {np.mean(lives):.1f}; time between the 10th-percentile "
              f"warp's end and the last warp's end: mean {np.mean(spreads):.1f} us ({np.mean(spreads) / np.mean(lives) * 100:.0f} % of the launch)")


if __name__ == "__main__":
    main(sys.argv[1])
