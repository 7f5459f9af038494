#!/usr/bin/env python
"""pcie_peak.py -- what the host link sustains for exactly the bytes one bench.py e2e step moves
(pinned host memory): H2D alone, D2H alone, and both at once on two streams (the shape of
grimb_impute_host's pipeline).  The e2e leg of bench.py cannot be faster than the duplex figure.
Measurement tool only (torch is plumbing here: pinned buffers, streams, events).

  python tools/pcie_peak.py [h2d_bytes d2h_bytes]   -> one JSON line
"""
import json
import sys

import torch


def main():
    h2d = int(sys.argv[1]) if len(sys.argv) > 1 else 52428812
    d2h = int(sys.argv[2]) if len(sys.argv) > 2 else 134937152
    dev = torch.device("cuda", 0)
    hin = torch.empty(h2d, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d, dtype=torch.uint8, device=dev)
    dout = torch.zeros(d2h, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_in, do_out, reps=20):
        best = 1e30
        for _ in range(reps + 3):
            torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            s1.wait_event(a)
            s2.wait_event(a)
            if do_in:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
    print(json.dumps({
        "gpu": torch.cuda.get_device_name(0), "h2d_bytes": h2d, "d2h_bytes": d2h,
        "h2d_alone_ms": t_in, "h2d_alone_gbs": h2d / t_in * 1e-6,
        "d2h_alone_ms": t_out, "d2h_alone_gbs": d2h / t_out * 1e-6,
        "duplex_ms": t_both, "duplex_gbs_total": (h2d + d2h) / t_both * 1e-6,
        "note": "best of 20, CUDA events; one copy per direction (no chunking, no kernels)"}))


if __name__ == "__main__":
    main()
