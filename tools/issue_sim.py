#!/usr/bin/env python3
"""Toy issue model of one SM sub-partition running W copies of a kernel's inner loop.

usage: issue_sim.py <object-or-so> <mangled-kernel-substring> <lo_hex> <hi_hex> [warps ...]

Reads the SASS of [lo, hi] with the control words (cuobjdump prints the two 64-bit halves of every 128-bit
instruction), takes every instruction's STALL COUNT (bits 41..44 of the high word: cycles before the same warp may
issue its next instruction — that is how ptxas encodes the fixed latencies of dependent DFMAs) and replays the
straight-line stream on W warps that share one issue slot per cycle and one FP64 pipe that is busy two cycles per
DFMA / DMUL / DADD.  Variable latencies (loads, MUFU) and branches are ignored: the result is an UPPER bound on the
FP64-pipe utilisation the static schedule allows at that occupancy."""
import re
import subprocess
import sys


def load(obj, kern, lo, hi):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    body = next(b for b in re.split(r"\n\s*Function : ", txt) if kern in b.split("\n")[0])
    lines = body.splitlines()
    out = []
    for i, ln in enumerate(lines[:-1]):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", ln)
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        if m and m2:
            a = int(m.group(1), 16)
            if lo <= a <= hi:
                op = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip()).split()[0].split(".")[0]
                stall = (int(m2.group(1), 16) >> 41) & 0xF
                out.append((op in ("DFMA", "DMUL", "DADD"), max(1, stall)))
    return out


def simulate(stream, warps, iters=6):
    n = len(stream)
    pc = [0] * warps
    done_iter = [0] * warps
    ready = [w * 97 for w in range(warps)]  # staggered start
    pipe_free, cycle, fp_busy = 0, 0, 0
    last = 0
    while min(done_iter) < iters:
        # oldest-ready-first among the warps that may issue this cycle
        cand = [w for w in range(warps) if ready[w] <= cycle and done_iter[w] < iters]
        issued = False
        for w in sorted(cand, key=lambda w: ready[w]):
            fp, st = stream[pc[w]]
            if fp and pipe_free > cycle:
                continue  # math-pipe throttle: try another warp
            if fp:
                pipe_free = cycle + 2
                fp_busy += 2
            ready[w] = cycle + st
            pc[w] += 1
            if pc[w] == n:
                pc[w] = 0
                done_iter[w] += 1
            issued = True
            last = (last + 1) % warps
            break
        cycle += 1
    return fp_busy / cycle, cycle / (iters * warps)


def main():
    obj, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    stream = load(obj, kern, lo, hi)
    nfp = sum(1 for fp, _ in stream if fp)
    print(f"{len(stream)} instructions, {nfp} FP64, sum of stall counts {sum(s for _, s in stream)} "
          f"(one warp alone: FP64 pipe {2 * nfp / sum(s for _, s in stream):.2f} busy)")
    for w in [int(v) for v in sys.argv[5:]] or [1, 2, 3, 4, 6, 8]:
        util, cyc = simulate(stream, w)
        print(f"{w} warps per scheduler: FP64 pipe busy {util:.3f}, {cyc:.0f} cycles per warp and loop body")


if __name__ == "__main__":
    main()
