#!/usr/bin/env python
"""Where does the streamed (host buffers in / out) step of config 2 lose time against the resident one?  Wraps the tokenizer with
CUDA events on the compute stream: per batch the device-side duration of its kernels and the idle gap before the next batch's
first kernel, for the resident loop and for pipeline.stream_tokenize.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from setok_b200.pipeline import stream_tokenize


class Timed:
    def __init__(self, tok):
        self.tok, self.ev = tok, []
        self.device = tok.device

    def __call__(self, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = self.tok(*a, **k)
        e.record()
        self.ev.append((s, e))
        return out

    def report(self, what):
        torch.cuda.synchronize()
        ev = self.ev[4:]                                     # skip the warm-up batches
        dur = [s.elapsed_time(e) for s, e in ev]
        gap = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)]
        total = ev[0][0].elapsed_time(ev[-1][1]) / len(ev)
        print(f"{what:28s}: kernels of a batch {sum(dur) / len(dur):7.3f} ms (min {min(dur):.3f} max {max(dur):.3f}), idle gap before the next batch "
              f"{sum(gap) / len(gap):6.3f} ms (max {max(gap):.3f}), per batch overall {total:7.3f} ms", flush=True)
        self.ev = []


def main():
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    images, noise = imgs_h.to(dev), noise_h.to(dev)
    h_u8, h_noise = u8.pin_memory(), noise_h.pin_memory()
    T = Timed(tok)
    n = 16
    for rnd in range(2):
        for _ in range(n):
            T(images, k=bench.KNN_K, noise=noise)
        T.report("resident, float images")
        d_u8 = u8.to(dev)
        for _ in range(n):
            T(d_u8, k=bench.KNN_K, noise=noise)
        T.report("resident, uint8 images")
        for _ in stream_tokenize(T, ((h_u8, h_noise) for _ in range(n)), k=bench.KNN_K):
            pass
        T.report("streamed, uint8 images")


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def variants():
    """Which ingredient of the streamed loop opens the ~0.3 ms gap at a batch boundary?  The resident loop plus one ingredient at a time."""
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    noise = noise_h.to(dev)
    d_u8 = u8.to(dev)
    h_u8 = u8.pin_memory()
    side, copy = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    T = Timed(tok)
    n = 14
    pinned = [torch.empty(256, 256, dtype=torch.int64).pin_memory() for _ in range(3)]

    def run(name, after=None, before=None):
        prev = None
        for i in range(n):
            img = before(i) if before else d_u8
            out = T(img, k=bench.KNN_K, noise=noise)
            if after:
                prev = after(i, out, prev)
        T.report(name)

    run("resident")

    def a1(i, out, prev):
        e = torch.cuda.Event(); e.record(main); return e
    run("+ event record", a1)

    def a2(i, out, prev):
        e = torch.cuda.Event(); e.record(main)
        side.wait_event(e)
        with torch.cuda.stream(side):
            pinned[i % 3].copy_(out[1], non_blocking=True)
        return e
    run("+ side-stream D2H after it", a2)

    def a3(i, out, prev):
        e = torch.cuda.Event(); e.record(main)
        side.wait_event(e)
        with torch.cuda.stream(side):
            pinned[i % 3].copy_(out[1], non_blocking=True)
            e2 = torch.cuda.Event(); e2.record(side)
        if prev is not None:
            prev.synchronize()
        return e2
    run("+ host sync one batch behind", a3)

    ups = {}

    def b4(i):
        with torch.cuda.stream(copy):
            d = h_u8.to(dev, non_blocking=True)
            e = torch.cuda.Event(); e.record(copy)
        ups[i] = (d, e)
        if i == 0:
            main.wait_event(e)
            return d
        d0, e0 = ups.pop(i - 1)
        if not e0.query():
            main.wait_event(e0)
        d0.record_stream(main)
        return d0
    run("+ H2D upload one batch ahead", None, b4)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "variants":
    variants()


def variants2():
    """The real stream_tokenize with pieces of its read-back switched off (monkey-patched), to see which piece opens the gap."""
    from setok_b200 import pipeline
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    h_u8, h_noise = u8.pin_memory(), noise_h.pin_memory()
    T = Timed(tok)
    n = 14

    def run(name):
        for _ in stream_tokenize(T, ((h_u8, h_noise) for _ in range(n)), k=bench.KNN_K):
            pass
        T.report(name)

    run("stream_tokenize as shipped")
    orig_init, orig_result, orig_rows = pipeline._Readback.__init__, pipeline._Readback.result, pipeline._Readback._rows

    # (a) no second hop: the rows are not copied at all
    def result_no_rows(self):
        self.ev.synchronize()
        self._release()
        return pipeline.HostResult(self.h_off, self.h_off, self.h_idx, self.h_score)
    pipeline._Readback.result = result_no_rows
    run("no row read-back (second hop)")

    # (b) additionally no first-hop copies (only the event)
    def init_no_copies(self, out, d2h, after):
        self.rt, self.idx, self.score = out
        self.d2h = d2h
        d2h.wait_event(after)
        self.h_off = self.h_idx = self.h_score = torch.zeros(1)
        self.h_tok = None
        self.ev = torch.cuda.Event()
        self.ev.record(d2h)
    pipeline._Readback.__init__ = init_no_copies
    run("no read-back copies at all")

    # (c) and no host wait either: results are handed out without synchronising
    def result_no_sync(self):
        self._release()
        return pipeline.HostResult(self.h_off, self.h_off, self.h_off, self.h_off)
    pipeline._Readback.result = result_no_sync
    run("... and no host sync")
    pipeline._Readback.__init__, pipeline._Readback.result, pipeline._Readback._rows = orig_init, orig_result, orig_rows


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "variants2":
    variants2()


def variants3():
    """Second hop of the read-back: is it the pinned allocation or the copy that opens the gap?"""
    from setok_b200 import pipeline
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    h_u8, h_noise = u8.pin_memory(), noise_h.pin_memory()
    T = Timed(tok)
    n = 14

    def run(name):
        for _ in stream_tokenize(T, ((h_u8, h_noise) for _ in range(n)), k=bench.KNN_K):
            pass
        T.report(name)

    orig_rows = pipeline._Readback._rows
    ring = [torch.empty(4096, 1024, dtype=torch.float32).pin_memory() for _ in range(3)]
    state = {"i": 0}

    def rows_prealloc(self, total):
        buf = ring[state["i"] % 3][:total]
        state["i"] += 1
        buf.copy_(self.rt.data[:total], non_blocking=True)
        return buf
    pipeline._Readback._rows = rows_prealloc
    run("rows into a preallocated ring")

    def rows_alloc_only(self, total):
        return torch.empty((total, self.rt.data.shape[1]), dtype=self.rt.data.dtype, pin_memory=True)
    pipeline._Readback._rows = rows_alloc_only
    run("pinned allocation, no copy")
    pipeline._Readback._rows = orig_rows
    run("as shipped")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "variants3":
    variants3()


def hosttrace():
    """Host-side timeline of stream_tokenize: when is each batch's launch sequence complete, when does each result come back?"""
    import time
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    h_u8, h_noise = u8.pin_memory(), noise_h.pin_memory()
    launched, returned = [], []

    class W:
        device = tok.device

        def __call__(self, *a, **k):
            t0 = time.perf_counter()
            out = tok(*a, **k)
            launched.append((t0, time.perf_counter()))
            return out
    for rnd in range(2):
        launched.clear(); returned.clear()
        torch.cuda.synchronize()
        for _ in stream_tokenize(W(), ((h_u8, h_noise) for _ in range(10)), k=bench.KNN_K):
            returned.append(time.perf_counter())
        base = launched[0][0]
        for j in range(10):
            print(f"round {rnd} batch {j}: launch begins {1e3 * (launched[j][0] - base):8.2f} ms, ends {1e3 * (launched[j][1] - base):8.2f} ms; "
                  f"result returned {1e3 * (returned[j] - base):8.2f} ms", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "hosttrace":
    hosttrace()
