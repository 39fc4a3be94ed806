"""Developer probe: why does the render launch take longer inside the step loop than alone?

    python tools/render_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.environ.get("PPB_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # PPB_ROOT: another build (A/B)
from pyphy_engine_b200 import sampler  # noqa: E402
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

N, NB = 131072, 8


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def main():
    kw = sampler.scaled64_params(NB)
    W = sampler.sample_rect_worlds(N, NB, seed=1, **kw)
    wb = WorldBatch(N, NB, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * NB, [(0, (1, 0, 0, 1))] * 4, int(W["rad"].min()), int(W["rad"].max()))
    frames = torch.empty((2, N, 64, 64, 4), dtype=torch.uint8, device="cuda")
    wb.step_async(20)
    wb.render(out=frames[0])
    print("A  render x20 back to back, same slot        : %.1f us" % timed(lambda: wb.render(out=frames[0]), 20))
    k = [0]

    def alt():
        wb.render(out=frames[k[0] & 1])
        k[0] += 1
    print("A2 render x20 back to back, alternating slots: %.1f us" % timed(alt, 20))
    phys = timed(lambda: wb.step_async(1), 20)
    print("P  physics step alone                        : %.1f us" % phys)

    def serial():
        wb.step_async(1)
        wb.render(out=frames[0])
    t = timed(serial, 20)
    print("B  physics + render, one stream              : %.1f us (render = %.1f)" % (t, t - phys))
    t = timed(lambda: wb.step_async(1, None, frames[:1], 1), 20)
    print("B2 step_async(1, frames) serial path         : %.1f us (render = %.1f)" % (t, t - phys))
    t = timed(lambda: wb.step_ring_async(20, frames, 1), 3) / 20
    print("C  step_ring_async pipelined                 : %.1f us/step" % t)
    t = timed(lambda: wb.step_ring_async(100, frames, 1), 2) / 100
    print("C2 step_ring_async pipelined, 100 steps/call : %.1f us/step" % t)
    if os.environ.get("PROBE_SHORT"):
        wb.close()
        return
    s2 = torch.cuda.Stream()

    def side():
        wb.render(out=frames[0], stream=s2)
    print("G  render x20 back to back on a side stream  : %.1f us" % timed(side, 20))

    def both():
        wb.step_async(1)
        wb.render(out=frames[0], stream=s2)
    t = timed(both, 40)
    print("H  render (side stream) || physics (main), no dependencies: %.1f us per pair" % t)

    def both2():
        wb.step_async(1)
        wb.step_async(1)
        wb.render(out=frames[0], stream=s2)
    t = timed(both2, 40)
    print("H2 render || 2 x physics, no dependencies    : %.1f us per group" % t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, with_phys in (("I  one render, nothing else", False), ("I2 one render while physics steps run", True)):
        acc = []
        for _ in range(10):
            torch.cuda.synchronize()
            if with_phys:
                wb.step_async(4)
            e0.record(s2)
            wb.render(out=frames[0], stream=s2)
            e1.record(s2)
            torch.cuda.synchronize()
            acc.append(1e3 * e0.elapsed_time(e1))
        print("%s: median %.1f us  min %.1f" % (label, float(np.median(acc)), min(acc)))
    wb.close()


if __name__ == "__main__":
    main()
