"""Splits the tile kernel's executed instructions and stall samples by phase from an `ncu --set full --import-source on` capture.
usage: ncu -i gpurun_out/raster_full.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv; python tools/tile_kernel_phases.py /tmp/src.csv 256"""
import csv
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "dfpsr_b200", "csrc", "raster.cu")).read().splitlines()


def line_of(text, after=0):
    for i in range(after, len(src)):
        if text in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + text)


kernel = line_of("raster_kernel(FrameDev frame, TexTable textures) {")
tex0, tex1 = line_of("__device__ __forceinline__ uint32_t weight_colors("), line_of("struct Rec {")
batch = line_of("for (uint32_t batchStart = 0;", kernel)
cover = line_of("uint32_t cover = 0;", batch)
rounds = line_of("while (__any_sync(0xffffffffu, cover != 0u))", cover)
chain = line_of("const int32_t outerStart = min(upperRow.x, lowerRow.x)", rounds)
weights = line_of("fillerTemplates.h:196-243", chain)
shade = line_of("RgbaMultiply.h:75-106", weights) - 1
epilogue = line_of("if (dirty) {", shade)


def bucket(path, line):
    if path.endswith("common.cuh"):
        return "shading: saturate + pack (common.cuh)"
    if not path.endswith("raster.cu"):
        return "other"
    if tex0 <= line < tex1:
        return "shading: u/v interpolation, mip level, bilinear sample, byte -> float"
    if kernel <= line < batch:
        return "per tile: prologue (tile lookup, target loads or clear, list sort)"
    if batch <= line < cover:
        return "per batch: checkpoints of 16 commands x 2 row pairs"
    if cover <= line < rounds:
        return "per batch: coverage masks -> per-lane command sets"
    if rounds <= line < chain:
        return "per round: command pick, record loads"
    if chain <= line < weights:
        return "per round: replay of the addition chains to the lane's quad"
    if weights <= line < shade:
        return "per round: 1/W, barycentric weights, coverage + depth test"
    if shade <= line < epilogue:
        return "per round: shader variants, alpha filter, register write-back"
    if line >= epilogue:
        return "per tile: epilogue (stores)"
    return "other"


rows = list(csv.reader(open(sys.argv[1])))
views = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cur, hdr, acc = None, None, defaultdict(lambda: [0, 0, 0])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        b = acc[bucket(cur, int(r[0]))]
        b[0] += int(d["Instructions Executed"]); b[1] += int(d["Thread Instructions Executed"]); b[2] += int(d["# Samples"])
total, samples = sum(v[0] for v in acc.values()), sum(v[2] for v in acc.values())
print("| phase | warp instructions per frame | share | active lanes | share of stall samples |")
print("|---|---|---|---|---|")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"| {k} | {v[0] / views / 1e6:.2f} M | {100 * v[0] / total:.1f} % | {v[1] / max(v[0], 1):.1f} | {100 * v[2] / samples:.1f} % |")
