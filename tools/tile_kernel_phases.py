"""Splits the tile kernel's executed instructions and stall samples by phase from an `ncu --set full --import-source on` capture.
usage: ncu -i gpurun_out/r2_tile_exact.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv
       python tools/tile_kernel_phases.py /tmp/src.csv 256 [path/to/raster.cu as it was when the capture was taken]"""
import csv, sys, re
from collections import defaultdict
rows=list(csv.reader(open(sys.argv[1])))
views=int(sys.argv[2]) if len(sys.argv)>2 else 256
srcpath=sys.argv[3] if len(sys.argv)>3 else 'dfpsr_b200/csrc/raster.cu'
src=open(srcpath).read().split('\n')
def line_of(text, after=0):
    for i in range(after, len(src)):
        if text in src[i]: return i+1
    raise SystemExit("marker not found: "+text)
marks=[
 ("sampling: load_tex/bilinear/mip/interp/unpack", line_of("__device__ __forceinline__ TexDev load_tex")),
 ("reciprocal_w", line_of("__device__ __forceinline__ float reciprocal_w")),
 ("structs/sort helpers", line_of("struct Rec {")),
 ("chain_to_quad (replay)", line_of("__device__ __forceinline__ void chain_to_quad")),
 ("shade_pixel (S-phase)", line_of("__device__ __forceinline__ uint32_t shade_pixel")),
 ("kernel prologue (tile lookup, loads/clear, sort)", line_of("raster_kernel(FrameDev frame) {")),
 ("batch prologue (checkpoints)", line_of("for (uint32_t batchStart = 0;")),
 ("cover masks", line_of("uint32_t cover = 0;")),
 ("round: pick + record loads", line_of("while (__any_sync(0xffffffffu, cover != 0u))")),
 ("round: row geometry + plane loads", line_of("const int32_t outerStart = min(upperRow.x, lowerRow.x), outerEnd = max(upperRow.y, lowerRow.y);", line_of("while (__any_sync(0xffffffffu, cover != 0u))"))),
 ("round: visibility", line_of("fillerTemplates.h:93-138")),
 ("round: deferred mip + take", line_of("if constexpr (DEFERRED) {", line_of("fillerTemplates.h:93-138"))),
 ("round: immediate shading", line_of("fillerTemplates.h:196-243 — weights", line_of("fillerTemplates.h:93-138"))),
 ("S-phase loop", line_of("---- shading pass")),
 ("epilogue stores", line_of("if (dirty) {", line_of("---- shading pass"))),
 ("host", line_of("} // namespace dfpsr", line_of("---- shading pass"))),
]
def bucket(path,line):
    if path.endswith('common.cuh'): return "saturate+pack (common.cuh)"
    if not path.endswith('raster.cu'): return "intrinsics (shfl etc.)"
    name="before"
    for n,l in marks:
        if line>=l: name=n
    return name
cur=None;hdr=None;acc=defaultdict(lambda:[0,0,0])
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1]; continue
    if len(r)>2 and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].isdigit() and r[2]=="-":
        d=dict(zip(hdr,r)); b=acc[bucket(cur,int(r[0]))]
        b[0]+=int(d["Instructions Executed"]); b[1]+=int(d["Thread Instructions Executed"]); b[2]+=int(d["# Samples"])
tot=sum(v[0] for v in acc.values()); ts=sum(v[2] for v in acc.values())
print("| phase | warp instructions per frame | share | active lanes | share of stall samples |\n|---|---|---|---|---|")
for k,v in sorted(acc.items(), key=lambda kv:-kv[1][0]):
    print(f"| {k} | {v[0]/views/1e6:.2f} M | {100*v[0]/tot:.1f} % | {v[1]/max(v[0],1):.1f} | {100*v[2]/ts:.1f} % |")
print("total %.2f M"%(tot/views/1e6))
