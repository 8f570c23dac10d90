#!/usr/bin/env python
"""Kernel timeline of replayed Agent.update graphs (CUPTI through torch.profiler; no nsys in this image).
Prints, for one critic-only and one actor step, every kernel with its start offset, duration and stream, and a per-name
summary: which kernels the step's wall time is made of, how much of it no kernel of the step was running (gaps), and how
many kernels ran concurrently.  usage: python tools/timeline.py [--batch B] [--morph M] [--full]"""
import argparse
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgrl_b200 import graph as G, morphologies as M, synth  # noqa: E402
from sgrl_b200.agent import Agent  # noqa: E402
from sgrl_b200.config import default_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--morph", default="3d_humanoid_9_full")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--full", action="store_true", help="print every kernel, not only the summary")
a = ap.parse_args()

torch.cuda.set_device(0)
ag = Agent(default_args())
par = M.ALL[a.morph]
ag.change_morphology(G.build_graph(par, device="cuda"))
b = {k: v.cuda() for k, v in synth.make_batch(a.batch, len(par), seed=1).items()}
for it in range(6):
    ag.update(b, it)
torch.cuda.synchronize()


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*", "", n)
    return n.replace("sgrl::", "")[:64]


for it, label in ((7, "critic-only step"), (8, "actor step")):
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        ag.update(b, it)
        torch.cuda.synchronize()
    class Ev:      # kineto activity: name, [start, end) in us, stream id
        def __init__(self, k):
            self.name, self.stream = k.name(), k.device_resource_id()
            st = k.start_ns() / 1e3
            self.time_range = type("R", (), {"start": st, "end": st + k.duration_ns() / 1e3})
    ev = [Ev(k) for k in prof.profiler.kineto_results.events()
          if k.device_type() == torch.autograd.DeviceType.CUDA and "memcpy" not in k.name().lower() and k.duration_ns() > 0]
    ev.sort(key=lambda e: e.time_range.start)
    if not ev:
        print("no CUDA events recorded (CUPTI unavailable?)")
        sys.exit(1)
    t0 = ev[0].time_range.start
    t1 = max(e.time_range.end for e in ev)
    print(f"# {label}: {len(ev)} kernels, {t1 - t0:.1f} us from first start to last end")
    # busy time (union of kernel intervals) and concurrency
    pts = sorted([(e.time_range.start, 1) for e in ev] + [(e.time_range.end, -1) for e in ev])
    busy, depth, last, conc = 0.0, 0, t0, collections.Counter()
    for t, d in pts:
        if depth > 0:
            busy += t - last
        conc[depth] += t - last
        depth += d
        last = t
    print(f"# some kernel running {busy:.1f} us, idle gaps {t1 - t0 - busy:.1f} us; time at concurrency k: " +
          ", ".join(f"{k}:{v:.0f}" for k, v in sorted(conc.items()) if v > 0.5))
    agg = collections.OrderedDict()
    for e in ev:
        k = short(e.name)
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += e.time_range.end - e.time_range.start
    tot = sum(v[1] for v in agg.values())
    print(f"# {'sum us':>9} {'n':>4} {'avg us':>7} {'share':>6}  kernel   (sum of durations {tot:.0f} us = {tot / (t1 - t0):.2f} x wall)")
    for k, (n, s) in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
        print(f"{s:11.1f} {n:4d} {s / n:7.1f} {100 * s / tot:5.1f}%  {k}")
    if a.full:      # per stream: start, end, end - previous end on that stream (the kernel's share of the stream's chain: with
        # programmatic dependent launch a kernel starts early and waits inside griddepcontrol.wait, so durations overlap)
        last_end = {}
        print("#   start      end   d(end) stream  kernel")
        for e in ev:
            le = last_end.get(e.stream, e.time_range.start)
            print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:8.1f} {e.time_range.end - max(le, e.time_range.start if le < e.time_range.start else le):8.1f} {e.stream:6d}  {short(e.name)}")
            last_end[e.stream] = e.time_range.end
